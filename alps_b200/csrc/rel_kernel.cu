// alps_b200: relativistic species (src/ALPS_fns_rel.f90), compiled with -fmad=false.
//
// One CTA per (omega, relativistic species, |n|) evaluates both signs and all six tensor
// components of full_integrate (src/ALPS_fns.f90:750-792) for that harmonic:
//   determine_resonances, relativistic branch       src/ALPS_fns.f90:683-701
//   no resonance: integrate() with gamma in resU    src/ALPS_fns.f90:799-864, 1560-1596
//   resonance:    integrate_res_rel / integrate_resU_rel / principal_integral_rel / funct_g_rel
//                                                   src/ALPS_fns_rel.f90:460-999
//                 landau_integrate_rel, int_T_res_rel, CBESSJ, Gamma, Fact   :1005-1092, 1374-1574
// The Bessel factors of int_T_rel depend on both grid indices, so J_{n-1}, J_n, J_{n+1} are
// evaluated (literal BESSJ) once per grid point and shared by the two signs' six components.
#include <stdlib.h>

#include "bessel.cuh"
#include "kernels.h"

namespace alps {

constexpr double PI_ = 3.14159265358979323846;

struct Six2 {
  cd v[6];
};
__device__ __forceinline__ void zero6(Six2& s) {
#pragma unroll
  for (int q = 0; q < 6; q++) s.v[q] = mk(0.0, 0.0);
}

// the six T components from bessel = J_n, besselP = J_n' (signs for n < 0 already applied),
// perpendicular factor pp (p_perp or pperpbar), parallel factor pz, zbar (= z of int_T), signed n
__device__ __forceinline__ void modes_real(double bj, double bp, double pp, double pz, double zbar, double nn,
                                           double kf1, double kf2, Six2& T) {
  T.v[0] = mk(1.0 * (nn * nn) * bj * bj / (zbar * zbar), 0.0);
  T.v[1] = mk(kf2 * (bp * bp * pp * pp), 0.0);
  T.v[2] = mk(kf2 * (bj * bj * (pz * pz)), 0.0);
  T.v[3] = mk(0.0, kf1 * ((1.0 * nn) * bj * bp * pp / zbar));
  T.v[4] = mk(kf1 * ((1.0 * nn) * bj * bj * pz / zbar), 0.0);
  T.v[5] = mk(0.0, -kf2 * (bj * bp * pz * pp));
}

// ---- Gamma (Lanczos), Fact, CBESSJ  (src/ALPS_fns_rel.f90:1738-1791, 1560-1574, 1502-1555)
__device__ inline double gamma_ref(double xx) {
  const double cof[6] = {76.18009173, -86.50532033, 24.01409822, -1.231739516, 0.120858003e-2, -0.536382e-5};
  const double stp = 2.50662827465;
  double x = xx - 1.0, tmp = x + 5.5, ser = 1.0;
  tmp = (x + 0.5) * log(tmp) - tmp;
  for (int j = 0; j < 6; j++) {
    x = x + 1.0;
    ser = ser + cof[j] / x;
  }
  return exp(tmp + log(stp * ser));
}
__device__ inline cd cpowi(cd z, int k) {
  cd r = mk(1.0, 0.0), b = z;
  unsigned n = (unsigned)(k < 0 ? -k : k);
  while (n) {
    if (n & 1u) r = r * b;
    n >>= 1;
    if (n) b = b * b;
  }
  return k < 0 ? mk(1.0, 0.0) / r : r;
}
// cold paths (principal-value window, Landau term) stay out of line: the hot loop then fits the
// instruction cache (ncu: no_instruction stalls 1.6 -> per issue before)
#define REL_NOINLINE __noinline__
// CBESSJ (src/ALPS_fns_rel.f90:1502-1555): J_{nu0}, J_{nu0+1}, J_{nu0+2} at complex z by its 21-term series
// sum_k (-z^2/4)^k / (k! Gamma(nu+k+1)) (z/2)^nu, sharing the powers
// (-z^2/4)^k / k! between the three orders; rfact[k] = 1/k!, rgam[i] = 1/Gamma(nu0 + 1 + i) from the
// literal Lanczos Gamma (tables per CTA).  nu0 = -1 (|n| = 0): the first order is skipped.
__device__ __forceinline__ void cbessj3(cd z, int nu0, const double* rfact, const double* rgam, cd& jm, cd& j0, cd& jp) {
  const cd mz2 = (-(z * z)) * 0.25;
  cd pw = mk(1.0, 0.0), sm_ = mk(0.0, 0.0), s0 = mk(0.0, 0.0), sp_ = mk(0.0, 0.0);
#pragma unroll 3
  for (int k = 0; k <= 20; k++) {
    const cd t = rfact[k] * pw;
    if (nu0 >= 0) sm_ += rgam[k] * t;
    s0 += rgam[k + 1] * t;
    sp_ += rgam[k + 2] * t;
    pw = pw * mz2;
  }
  const cd zh = 0.5 * z;
  const cd p0 = cpowi(zh, nu0 + 1);
  j0 = p0 * s0;
  jp = (p0 * zh) * sp_;
  jm = nu0 >= 0 ? cpowi(zh, nu0) * sm_ : mk(0.0, 0.0);
}
__device__ inline cd csqrt_(cd z) {
  double m = hypot(z.x, z.y);
  if (m == 0.0) return mk(0.0, 0.0);
  double a = sqrt(0.5 * (m + fabs(z.x)));
  double b = 0.5 * z.y / a;
  return z.x >= 0.0 ? mk(a, b) : mk(fabs(b), copysign(a, z.y));
}

// eval_fit for relativistic species: fit types 4 and 5 (src/ALPS_analyt.f90:85-117, 222-232)
__device__ inline cd eval_fit_rel(const GlobalDev& g, const SpeciesDev& sp, int igamma, cd p) {
  cd f = mk(0.0, 0.0);
  const double gam = sp.grel[igamma];
  for (int ifit = 0; ifit < sp.n_fits; ifit++) {
    const double* pf = sp.param_fit + ((size_t)igamma * g.maxfits + ifit) * 5;
    const double pc = sp.perp_correction[ifit];
    if (sp.fit_type[ifit] == 4) {
      f += mk(pf[0] * exp(-pc * gam), 0.0);
    } else if (sp.fit_type[ifit] == 5) {
      cd d = p - mk(pf[2], 0.0);
      cd e = -(pf[1] * (d * d));
      double ex = exp(e.x), sn, cs;
      sincos(e.y, &sn, &cs);
      f += (pf[0] * exp(-pc * gam)) * mk(ex * cs, ex * sn);
    }
  }
  return f;
}

// BESSJ triple at z for |n|: bessel and besselP of int_T_rel with the sign rules for n < 0
__device__ __forceinline__ void bessel_pair(int nabs, int sg, double z, double jm, double j0, double jp, double& bj,
                                            double& bp) {
  // jm = J_{|n|-1} (|n| >= 1), j0 = J_|n|, jp = J_{|n|+1}
  const double par = (nabs & 1) ? -1.0 : 1.0;
  if (nabs == 0) {
    bj = j0;
    bp = -jp;   // -J_1
  } else if (!sg) {
    bj = j0;
    bp = 0.5 * (jm - jp);
  } else {
    // n < 0: bessel = (-1)^n J_|n| ; besselP = 0.5((-1)^(n-1) J_{|n|+1} - (-1)^(n+1) J_{|n|-1}), n = -1: 0.5 (J_2 - J_0)
    bj = par * j0;
    bp = (nabs == 1) ? 0.5 * (jp - jm) : 0.5 * ((-par) * jp - (-par) * jm);
  }
  (void)z;
}

// weight of node ip in the one-sided trapezoid pieces of integrate_resU_rel (lines 663-697)
__device__ __forceinline__ double piece_w(int ip, int a, int b) {
  if (a > b || ip < a || ip > b) return 0.0;
  if (a == b) return 1.0;
  return (ip == a || ip == b) ? 1.0 : 2.0;
}

struct RelCtx {
  const GlobalDev* g;
  const SpeciesDev* sp;
  cd om;
  int nabs;
  double pref;    // -2 pi (ms/vA)^3 (qs vA/(kpar ms))
  double zfac;    // kperp ms/(vA qs)
  double zbar;
  double kf1, kf2;
};

// Node (ig, ip), sign sg: num = pref (om dfg + (kpar/vA) dfp) (numerator of resU_rel) and the six real Bessel
// moments M = {J^2, J^2 p, J^2 p^2, J J' pp, J J' pp p, J'^2 pp^2} (J, J' with the sign rules for n < 0, pp =
// pperpbar, p = pparbar) of which the six components of int_T_rel are constant multiples (moments_to_modes).
// (the literal BESSJ per point is the fallback without the per-k tables: out of line, one copy, so that the kernels that
// inline node_moments several times stay small -- their code runs from L2 when it runs once per launch)
__device__ __noinline__ double bessj_cold(int n, double z) { return bessj_ref(n, z); }
__device__ inline void node_moments(const RelCtx& c, int ig, int ip, int sg, double* M, cd& num) {
  const SpeciesDev& sp = *c.sp;
  const int ldr = c.g->npparbar + 1;
  const double gam = sp.grel[ig], pb = sp.pbrel[ip];
  const double pperpbar = sqrt(gam * gam - 1.0 - pb * pb);
  const double z = c.zfac * pperpbar;
  double j0, jp, jm;
  if (sp.Jrel) {
    // the same BESSJ values, tabulated once per k (k_rel_bessel_table)
    const size_t plane = (size_t)(c.g->ngamma + 1) * ldr, o = (size_t)ig * ldr + ip;
    j0 = sp.Jrel[(size_t)c.nabs * plane + o];
    jp = sp.Jrel[(size_t)(c.nabs + 1) * plane + o];
    jm = c.nabs >= 1 ? sp.Jrel[(size_t)(c.nabs - 1) * plane + o] : 0.0;
  } else {
    j0 = bessj_cold(c.nabs, z);
    jp = bessj_cold(c.nabs + 1, z);
    jm = c.nabs >= 1 ? bessj_cold(c.nabs - 1, z) : 0.0;
  }
  double bj, bp;
  bessel_pair(c.nabs, sg, z, jm, j0, jp, bj, bp);
  const double b2 = bj * bj, bb = bj * bp * pperpbar, q2 = (bp * pperpbar) * (bp * pperpbar);
  M[0] = b2;
  M[1] = b2 * pb;
  M[2] = (b2 * pb) * pb;
  M[3] = bb;
  M[4] = bb * pb;
  M[5] = q2;
  const double dfg = sp.dfg_rel[(size_t)ig * ldr + ip], dfp = sp.dfp_rel[(size_t)ig * ldr + ip];
  num = c.pref * (c.om * dfg + mk((c.g->kpar / c.g->vA) * dfp, 0.0));
}

// moment sums -> tensor components (int_T_rel, src/ALPS_fns_rel.f90:1256-1369), nn = signed harmonic
__device__ __forceinline__ void moments_to_modes(const RelCtx& c, double nn, const Six2& S, Six2& T) {
  const double c0 = (nn * nn) / (c.zbar * c.zbar), c3 = c.kf1 * nn / c.zbar;
  T.v[0] = c0 * S.v[0];
  T.v[1] = c.kf2 * S.v[5];
  T.v[2] = c.kf2 * S.v[2];
  T.v[3] = cmul_i(c3 * S.v[3]);
  T.v[4] = c3 * S.v[1];
  T.v[5] = -cmul_i(c.kf2 * S.v[4]);
}

// node selection of funct_g_rel (src/ALPS_fns_rel.f90:918-999)
__device__ __forceinline__ int funct_g_node(const SpeciesDev& sp, int npb, int ig, double p, double inv_dpb) {
  const double* pbv = sp.pbrel;
  const double* f0r = sp.f0_rel + (size_t)ig * (npb + 1);
  int ic = -2;
  {
    int i0 = (int)floor((p - pbv[0]) * inv_dpb);
    for (int q = min(i0 + 2, npb - 1); q >= max(i0 - 2, 0); q--)
      if (pbv[q + 1] > p && pbv[q] <= p) {
        ic = q;
        break;
      }
  }
  if (ic + 1 >= 0 && ic + 1 <= npb && f0r[ic + 1] <= -1.0) ic = ic - 1;
  if (ic - 1 >= 0 && ic - 1 <= npb && f0r[ic - 1] <= -1.0) ic = ic + 1;
  if (p == pbv[npb]) ic = npb - 2;
  if (ic >= npb - 1) ic = npb - 2;
  if (ic <= 1) ic = 2;
  return ic;
}

// funct_g_rel for one sign, moment form, nodes evaluated on the spot (fallback of funct_g_win)
__device__ REL_NOINLINE void funct_g_rel6(const RelCtx& c, int sg, double p, int ig, int ic, double inv_dpb, Six2& out) {
  double Mm[6], M0[6], Mp[6];
  cd nm, n0, np_;
  node_moments(c, ig, ic - 1, sg, Mm, nm);
  node_moments(c, ig, ic, sg, M0, n0);
  node_moments(c, ig, ic + 1, sg, Mp, np_);
  const double sx = (0.5 * inv_dpb) * (p - c.sp->pbrel[ic]);
#pragma unroll
  for (int q = 0; q < 6; q++) {
    const cd a = M0[q] * n0, lo = Mm[q] * nm, hi = Mp[q] * np_;
    out.v[q] = mk(fma(sx, hi.x - lo.x, a.x), fma(sx, hi.y - lo.y, a.y));
  }
}

// funct_g_rel for the principal-value window: the node values num M (moment form) of the 2 M_I + 7 nodes
// around the resonance are evaluated once per (row, sign) into shared memory (win[node - W0][6]); every
// quadrature point then only selects its node (the same search and cone rules as funct_g_rel6) and
// interpolates.  Falls back to funct_g_rel6 if a node outside the window is asked for.
constexpr int REL_WIN = 24;      // 2 M_I + 7 window nodes: positions_principal up to 8 (larger: no window cache)
// per-warp cache of one principal-value window (shared memory): node values, central differences, the node
// coordinates (one more than the window: the search tests pb(q + 1)) and the outside-the-cone flags
struct RelWin {
  cd val[REL_WIN][6];
  cd dif[REL_WIN][6];          // val[k + 1] - val[k - 1], k = 1 .. nwin - 2
  double pb[REL_WIN + 1];
  unsigned char out[REL_WIN + 1];
  int anyout;                  // any out[] flag set
};
// Window slot k (node W0 + k) that funct_g_rel interpolates around for the point p, when the cached search decides it
// (see funct_g_win); -1: the caller takes the general path.
__device__ __forceinline__ int win_locate(const SpeciesDev& sp, int npb, double p, const RelWin& w, int W0, int nwin,
                                          double inv_dpb) {
  if (w.anyout) return -1;
  const int i0 = (int)floor((p - sp.pbrel[0]) * inv_dpb);
  const int k0 = i0 - W0;
  if (!(k0 >= 2 && k0 + 3 <= nwin && i0 >= 4 && i0 + 4 <= npb)) return -1;
  int q = k0;
  if (w.pb[q] > p) q--;
  else if (w.pb[q + 1] <= p) q++;
  if (!(w.pb[q + 1] > p && w.pb[q] <= p)) return -1;
  return (q >= 1 && q + 1 < nwin) ? q : -1;
}
__device__ __forceinline__ void funct_g_win(const RelCtx& c, int sg, double p, int ig, const RelWin& w, int W0,
                                            int nwin, double inv_dpb, Six2& out) {
  const SpeciesDev& sp = *c.sp;
  const int npb = c.g->npparbar;
  // node selection of funct_g_rel (funct_g_node) on the cached coordinates / cone flags when everything it touches lies
  // in the window -- the normal case; same comparisons on the same values, so the same node.  The cell [pb(q), pb(q+1))
  // that holds p is unique, so instead of funct_g_node's scan from i0 + 2 downwards the index guess i0 is corrected by
  // at most one step and then checked with the reference's two comparisons.
  int ic = -2;
  bool cached = false;
  {
    const int i0 = (int)floor((p - sp.pbrel[0]) * inv_dpb);
    const int k0 = i0 - W0;
    if (k0 >= 2 && k0 + 3 <= nwin && i0 >= 4 && i0 + 4 <= npb) {
      int q = k0;
      if (w.pb[q] > p) q--;
      else if (w.pb[q + 1] <= p) q++;
      if (w.pb[q + 1] > p && w.pb[q] <= p) {
        ic = q + W0;
        cached = !w.anyout;   // a node of the window outside the sub-luminal cone (rare: the window lies inside it):
                              // funct_g_node applies the cone rules
        // (p == pb(npb) and the clamps ic >= npb - 1, ic <= 1 cannot trigger: 3 <= i0 - 1 <= ic <= i0 + 1 <= npb - 3)
      }
    }
  }
  if (!cached) ic = funct_g_node(sp, npb, ig, p, inv_dpb);
  const int k = ic - W0;
  if (k < 1 || k + 1 >= nwin) {
    Six2 tmp;     // (a temporary: the caller's accumulators must not escape to an out-of-line call)
    funct_g_rel6(c, sg, p, ig, ic, inv_dpb, tmp);
#pragma unroll
    for (int q = 0; q < 6; q++) out.v[q] = tmp.v[q];
    return;
  }
  const double sx = (0.5 * inv_dpb) * (p - w.pb[k]);   // central-difference slope factor times the offset
#pragma unroll
  for (int q = 0; q < 6; q++) {
    const cd a = w.val[k][q], d = w.dif[k][q];
    out.v[q] = mk(fma(sx, d.x, a.x), fma(sx, d.y, a.y));
  }
}

__device__ __forceinline__ cd warp_sum_cd(cd v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

// principal_integral_rel for one Gamma row (src/ALPS_fns_rel.f90:724-913): window of 2 M_I + 7 nodes around the resonance
// cached in shared memory (one per warp), symmetric pairing / linearised branch, tiny rest; adds wg smdelta (sum) to Sd.
// Called by a whole warp with warp-uniform arguments.
__device__ __forceinline__ void pv_row(const RelCtx& c, const GlobalDev& g, const SpeciesDev& sp, cd omc, int sg, double nn,
                                       int ig, int ires, int upperlimit, double wg, RelWin& win_, int lane, Six2& Sd) {
  const int npb = g.npparbar, M_I = g.M_I, M_P = g.M_P, ldr = npb + 1;
  const double qs = sp.qs, ms = sp.ms, vA = g.vA, kpar = g.kpar, dpb = sp.dpparbar;
  const double* pbv = sp.pbrel;
  const double gres = sp.grel[ig];   // gamma_rel(sproc_rel,igamma,ipparbar_res): separable grid
  const double denomR = (gres * omc.x * vA / kpar) - (1.0 * nn) * (qs / ms) * vA / kpar;
  const double denomI = gres * omc.y * vA / kpar;
  const double capDelta = denomR - pbv[ires - M_I];
  const double smdelta = capDelta / (1.0 * M_P);
  Six2 pr;
  zero6(pr);
  // node values of the window [ires - M_I - 2, ires + M_I + 4]: one node per lane
  const int W0 = ires - M_I - 2, nwin = (2 * M_I + 7 <= REL_WIN) ? 2 * M_I + 7 : 0;
  RelWin& win = win_;
  const double inv_dpb = 1.0 / dpb;
  __syncwarp();
  if (lane <= nwin && nwin > 0) {
    const int idx = W0 + lane;
    const bool in = idx >= 0 && idx <= npb;
    // coordinates outside the table never match a comparison (NaN): the cached search then misses and the
    // global one decides
    win.pb[lane] = in ? pbv[idx] : __longlong_as_double(0x7ff8000000000000LL);
    win.out[lane] = (in && sp.f0_rel[(size_t)ig * ldr + idx] <= -1.0) ? 1 : 0;
    const unsigned any = __ballot_sync(__activemask(), win.out[lane] != 0);
    if (lane == 0) win.anyout = any != 0;
    if (lane < nwin) {
      double M[6];
      cd num = mk(0.0, 0.0);
#pragma unroll
      for (int q = 0; q < 6; q++) M[q] = 0.0;
      if (in) node_moments(c, ig, idx, sg, M, num);
#pragma unroll
      for (int q = 0; q < 6; q++) win.val[lane][q] = M[q] * num;
    }
  }
  __syncwarp();
  if (lane >= 1 && lane + 1 < nwin) {
#pragma unroll
    for (int q = 0; q < 6; q++) win.dif[lane][q] = win.val[lane + 1][q] - win.val[lane - 1][q];
  }
  __syncwarp();
  if (fabs(denomI) > g.Tlim) {
    for (int j = lane; j <= M_P; j += 32) {
      const double wj = (j == 0 || j == M_P) ? 1.0 : 2.0;
      const double p = (j == 0) ? denomR : (j == M_P ? denomR + capDelta : denomR + smdelta * j);
      const double p2 = 2.0 * denomR - p;
      // wj / d1 and wj / d2 with d2 = conj(d1): one reciprocal for the twelve quotients
      const double dx = p - denomR, tt = wj * fast_rcp(fma(dx, dx, denomI * denomI));
      // f1 r1 - f2 conj(r1), r1 = (x, y)
      const double rx = dx * tt, ry = denomI * tt;
      const int k1 = win_locate(sp, npb, p, win, W0, nwin, inv_dpb);
      const int k2 = win_locate(sp, npb, p2, win, W0, nwin, inv_dpb);
      if (k1 >= 0 && k2 >= 0) {
        // both points interpolate inside the cached window (the normal case): component by component, nothing but
        // the running sums stays live
        const double sx1 = (0.5 * inv_dpb) * (p - win.pb[k1]), sx2 = (0.5 * inv_dpb) * (p2 - win.pb[k2]);
#pragma unroll
        for (int q = 0; q < 6; q++) {
          const cd a1 = win.val[k1][q], d1 = win.dif[k1][q], a2 = win.val[k2][q], d2 = win.dif[k2][q];
          const double f1x = fma(sx1, d1.x, a1.x), f1y = fma(sx1, d1.y, a1.y);
          const double f2x = fma(sx2, d2.x, a2.x), f2y = fma(sx2, d2.y, a2.y);
          pr.v[q].x += fma(f1x - f2x, rx, -((f1y + f2y) * ry));
          pr.v[q].y += fma(f1x + f2x, ry, (f1y - f2y) * rx);
        }
      } else {
        Six2 f1, f2;
        funct_g_win(c, sg, p, ig, win, W0, nwin, inv_dpb, f1);
        funct_g_win(c, sg, p2, ig, win, W0, nwin, inv_dpb, f2);
#pragma unroll
        for (int q = 0; q < 6; q++) {
          const double sr = f1.v[q].x - f2.v[q].x, si = f1.v[q].y + f2.v[q].y;
          const double tr = f1.v[q].x + f2.v[q].x, ti = f1.v[q].y - f2.v[q].y;
          pr.v[q].x += fma(sr, rx, -(si * ry));
          pr.v[q].y += fma(tr, ry, ti * rx);
        }
      }
    }
  } else {
    Six2 fp_, fm_;
    funct_g_win(c, sg, denomR + dpb, ig, win, W0, nwin, inv_dpb, fp_);
    funct_g_win(c, sg, denomR - dpb, ig, win, W0, nwin, inv_dpb, fm_);
    // sum_j 2 wj g' x^2 / (x^2 + denomI^2): g' does not depend on j
    double sj = 0.0;
    for (int j = 1 + lane; j <= M_P; j += 32) {
      const double wj = (j == M_P) ? 1.0 : 2.0;
      const double p = (j == M_P) ? denomR + capDelta : denomR + smdelta * j;
      const double x2 = (p - denomR) * (p - denomR);
      sj += ((wj * 2.0) * x2) * fast_rcp(x2 + denomI * denomI);
    }
    const double h2 = 0.5 * inv_dpb;
#pragma unroll
    for (int q = 0; q < 6; q++) pr.v[q] += (sj * h2) * (fp_.v[q] - fm_.v[q]);
    if (lane == 0 && denomI != 0.0) {
      Six2 f0_;
      funct_g_win(c, sg, denomR, ig, win, W0, nwin, inv_dpb, f0_);
      const double sgn = denomI > 0.0 ? 1.0 : -1.0;
#pragma unroll
      for (int q = 0; q < 6; q++) pr.v[q] += sgn * (cmul_i((2.0 * PI_) * f0_.v[q]) / smdelta);
    }
  }
  const double rest = pbv[upperlimit] - denomR - capDelta;
  const int ntiny = (int)(rest / smdelta);
  if (ntiny > 0) {
    const double correction = (rest / (1.0 * ntiny)) / smdelta;
    for (int j = lane; j <= ntiny; j += 32) {
      const double wj = (j == 0 || j == ntiny) ? 1.0 : 2.0;
      const double p = (j == 0) ? denomR + capDelta : denomR + capDelta + correction * smdelta * j;
      Six2 f1;
      funct_g_win(c, sg, p, ig, win, W0, nwin, inv_dpb, f1);
      const double dx = p - denomR, tt = (wj * correction) * fast_rcp(fma(dx, dx, denomI * denomI));
      const cd r1 = mk(dx * tt, denomI * tt);
#pragma unroll
      for (int q = 0; q < 6; q++) pr.v[q] += f1.v[q] * r1;
    }
  }
#pragma unroll
  for (int q = 0; q < 6; q++) Sd.v[q] += (wg * smdelta) * pr.v[q];
}

// landau_integrate_rel (src/ALPS_fns_rel.f90:1005-1092) for Im(om) <= 0: this thread's Gamma rows first + 1, first + 1 +
// stride, ...; adds i L mult to acc
__device__ __forceinline__ void landau_rows(const RelCtx& c, const GlobalDev& g, const SpeciesDev& sp, cd omc, int sg,
                                            double nn, int first, int stride, const double* s_rfact, const double* s_rgam,
                                            Six2& acc) {
  const int ng = g.ngamma, nabs = c.nabs;
  const double qs = sp.qs, ms = sp.ms, vA = g.vA, kpar = g.kpar, dpb = sp.dpparbar, dgam = sp.dgamma;
  Six2 L;
  zero6(L);
  for (int ig = 1 + first; ig <= ng - 1; ig += stride) {
    const double g1 = sp.grel[ig];
    const cd pres = (g1 * omc) * vA / kpar - mk((1.0 * nn) * qs * vA / (kpar * ms), 0.0);
    if (!(pres.x * pres.x <= g1 * g1 - 1.0)) continue;
    const double h = (ig == ng - 1) ? 0.5 : 1.0;
    // the four fit evaluations of the two central differences: one copy of eval_fit_rel in the code
    cd fv[4];
#pragma unroll 1
    for (int k = 0; k < 4; k++) {
      const int row = k == 0 ? ig + 1 : (k == 1 ? (ig == 1 ? ig : ig - 1) : ig);
      const cd pk = k < 2 ? pres : (k == 2 ? pres + mk(dpb, 0.0) : pres - mk(dpb, 0.0));
      fv[k] = eval_fit_rel(g, sp, row, pk);
    }
    const cd dfg = (ig == 1) ? (fv[0] - fv[1]) / dgam : (fv[0] - fv[1]) / (2.0 * dgam);
    const cd dfp = (fv[2] - fv[3]) / (2.0 * dpb);
    const cd fac = -h * (omc * dfg + (kpar / vA) * dfp);
    // int_T_res_rel with complex-argument Bessel functions
    const cd pperpbar = csqrt_(mk(g1 * g1 - 1.0, 0.0) - pres * pres);
    const cd z = c.zfac * pperpbar;
    const double par = (nabs & 1) ? -1.0 : 1.0;
    cd bj, bp, b1, b2;
    cbessj3(z, nabs - 1, s_rfact, s_rgam, b1, bj, b2);   // J_{|n|-1}, J_|n|, J_{|n|+1}
    if (sg) bj = par * bj;
    if (nabs == 0) {
      bp = -b2;
    } else {
      if (!sg) bp = 0.5 * (b1 - b2);
      else bp = (nabs == 1) ? 0.5 * (b2 - b1) : 0.5 * ((-par) * b2 - (-par) * b1);
    }
    cd T[6];
    T[0] = ((nn * nn) / (c.zbar * c.zbar)) * (bj * bj);
    T[1] = c.kf2 * (bp * bp * pperpbar * pperpbar);
    T[2] = c.kf2 * (bj * bj * (pres * pres));
    T[3] = cmul_i((c.kf1 * nn / c.zbar) * (bj * bp * pperpbar));
    T[4] = (c.kf1 * nn / c.zbar) * (bj * bj * pres);
    T[5] = -cmul_i(c.kf2 * (bj * bp * pres * pperpbar));
#pragma unroll
    for (int q = 0; q < 6; q++) L.v[q] += fac * T[q];
  }
  const double mult = (omc.y < 0.0 ? 2.0 : 1.0) * dgam * PI_ * 2.0 * PI_ * (qs * vA / (kpar * ms)) *
                      ((ms / vA) * (ms / vA) * (ms / vA));
#pragma unroll
  for (int q = 0; q < 6; q++) acc.v[q] += cmul_i(L.v[q]) * mult;
}

// One Gamma row of integrate_res_rel (src/ALPS_fns_rel.f90:580-720) for a warp: the limits of integrate_resU_rel, the
// direct part over the p_par-bar nodes (lanes) and the principal part (pv_row); adds the row's Bessel-moment sums to Sd.
__device__ __forceinline__ void rel_res_row(const RelCtx& c, const GlobalDev& g, const SpeciesDev& sp, cd omc, int sg,
                                            double nn, int ig, int lane, RelWin& win, Six2& Sd, int* __restrict__ err_flag) {
  const int ng = g.ngamma, npb = g.npparbar, M_I = g.M_I, nabs = c.nabs;
  const double qs = sp.qs, ms = sp.ms, vA = g.vA, kpar = g.kpar;
  const double dpb = sp.dpparbar;
  const double* pbv = sp.pbrel;
  const int ldr = npb + 1;
    const double wg = (ig == ng - 1) ? 1.0 : 2.0;
    const double g1 = sp.grel[ig];
    const cd pres = (g1 * omc - mk(nn * qs / ms, 0.0)) * vA / kpar;
    int ires = 0, found = 0;
    if (pres.x * pres.x <= g1 * g1 - 1.0) {
      if (pres.x >= pbv[1] && pres.x < pbv[npb - 1]) {
        int lo = 1, hi = npb - 2;
        while (lo < hi) {
          int mid = (lo + hi + 1) >> 1;
          if (pbv[mid] <= pres.x) lo = mid; else hi = mid - 1;
        }
        if (pbv[lo + 1] > pres.x && pbv[lo] <= pres.x) {
          ires = lo;
          found = 1;
        }
      }
    }
    for (int ip = 0; ip <= M_I; ip++) {
      if (pres.x >= pbv[0] - dpb * ip && pres.x < pbv[0] - dpb * (ip - 1)) {
        ires = -ip;
        found = 1;
      }
      if (pres.x >= pbv[npb - 1] + dpb * ip && pres.x < pbv[npb - 1] + dpb * (ip + 1)) {
        ires = npb - 1 + ip;
        found = 1;
      }
    }
    const int lo_c = sp.cone_lo[ig], up_c = sp.cone_up[ig];
    int int_start, int_end, lowerlimit, upperlimit;
    if (found) {
      int_start = lo_c;
      int_end = up_c;
      lowerlimit = ires - M_I;
      upperlimit = ires + M_I + 1;
      if (ires >= 0 && ires <= npb)
        if (fabs(pres.x - pbv[ires]) > 0.5 * dpb) upperlimit = upperlimit + 1;
      if (lowerlimit < lo_c && upperlimit > up_c) {
        if (lane == 0) err_flag[0] = 8;   // alps_error(8)
        return;
      } else if (lowerlimit <= lo_c) {
        int_start = 1;
        lowerlimit = 0;
        upperlimit = lo_c;
      } else if (upperlimit >= up_c) {
        lowerlimit = up_c;
        upperlimit = npb;
        int_end = npb - 1;
      }
    } else {
      int_start = lo_c;
      lowerlimit = up_c;
      int_end = npb - 1;
      upperlimit = npb;
    }
    // direct part
    if (sp.Jrel) {
      // hot loop: the six T components are real multiples of six Bessel moments (like the table species'
      // p_par moments), so the loop accumulates sum U {J^2, J^2 p, J^2 p^2, J J' pp, J J' pp p, J'^2 pp^2}
      // with one reciprocal per node; Bessel factors and pperpbar come from the per-k tables
      const size_t plane = (size_t)(ng + 1) * ldr;
      const double* __restrict__ J0 = sp.Jrel + (size_t)nabs * plane + (size_t)ig * ldr;
      const double* __restrict__ JP = J0 + plane;
      const double* __restrict__ JM = nabs >= 1 ? J0 - plane : J0;
      const double* __restrict__ PP = sp.Jrel + (size_t)(sp.nhi + 2) * plane + (size_t)ig * ldr;
      const double* __restrict__ DG = sp.dfg_rel + (size_t)ig * ldr;
      const double* __restrict__ DP = sp.dfp_rel + (size_t)ig * ldr;
      const cd gom = (g1 * omc) * vA / kpar;
      const double nqv = nn * qs * vA / (kpar * ms), kv = kpar / vA, cw = wg * dpb;
      const double par = (nabs & 1) ? -1.0 : 1.0;
      // the node range with non-zero weight: the union of [int_start, lowerlimit] and [upperlimit, int_end];
      // nodes in the gap between them have weight 0 and are predicated off, so the loads of two
      // iterations can be in flight together
      int a0 = npb, b0 = 0;
      if (int_start <= lowerlimit) {
        a0 = int_start;
        b0 = lowerlimit;
      }
      if (upperlimit <= int_end) {
        a0 = min(a0, upperlimit);
        b0 = max(b0, int_end);
      }
      a0 = max(a0, 1);
      b0 = min(b0, npb - 1);
#pragma unroll 2
      for (int ip = a0 + lane; ip <= b0; ip += 32) {
        const double w = piece_w(ip, int_start, lowerlimit) + piece_w(ip, upperlimit, int_end);
        const double j0 = J0[ip], jp = JP[ip], jm = nabs >= 1 ? JM[ip] : 0.0;
        double bj, bp;
        if (nabs == 0) {
          bj = j0;
          bp = -jp;
        } else if (!sg) {
          bj = j0;
          bp = 0.5 * (jm - jp);
        } else {
          bj = par * j0;
          bp = (nabs == 1) ? 0.5 * (jp - jm) : 0.5 * ((-par) * jp - (-par) * jm);
        }
        const double pb = pbv[ip], pq = PP[ip], dfg = DG[ip], dfp = DP[ip];
        const double nr = c.pref * fma(omc.x, dfg, kv * dfp), ni = c.pref * (omc.y * dfg);
        const double dr = pb - gom.x + nqv, di = -gom.y;
        const double t = (w != 0.0) ? (cw * w) * fast_rcp(fma(dr, dr, di * di)) : 0.0;
        const double ur = fma(nr, dr, ni * di) * t, ui = fma(ni, dr, -(nr * di)) * t;
        const double b2 = bj * bj, bb = bj * bp * pq, q2 = (bp * pq) * (bp * pq);
        const double b2p = b2 * pb, b2pp = b2p * pb, bbp = bb * pb;
        Sd.v[0].x = fma(ur, b2, Sd.v[0].x);   Sd.v[0].y = fma(ui, b2, Sd.v[0].y);
        Sd.v[1].x = fma(ur, b2p, Sd.v[1].x);  Sd.v[1].y = fma(ui, b2p, Sd.v[1].y);
        Sd.v[2].x = fma(ur, b2pp, Sd.v[2].x); Sd.v[2].y = fma(ui, b2pp, Sd.v[2].y);
        Sd.v[3].x = fma(ur, bb, Sd.v[3].x);   Sd.v[3].y = fma(ui, bb, Sd.v[3].y);
        Sd.v[4].x = fma(ur, bbp, Sd.v[4].x);  Sd.v[4].y = fma(ui, bbp, Sd.v[4].y);
        Sd.v[5].x = fma(ur, q2, Sd.v[5].x);   Sd.v[5].y = fma(ui, q2, Sd.v[5].y);
      }
    } else {
      for (int ip = 1 + lane; ip <= npb - 1; ip += 32) {
        const double w = piece_w(ip, int_start, lowerlimit) + piece_w(ip, upperlimit, int_end);
        if (w == 0.0) continue;
        double M[6];
        cd num;
        node_moments(c, ig, ip, sg, M, num);
        const cd den = mk(pbv[ip], 0.0) - (g1 * omc) * vA / kpar + mk(nn * qs * vA / (kpar * ms), 0.0);
        const cd U = (wg * dpb * w) * (num / den);
#pragma unroll
        for (int q = 0; q < 6; q++) Sd.v[q] += M[q] * U;
      }
    }
    // principal part
    if (found && lowerlimit >= int_start && upperlimit <= int_end)
      pv_row(c, g, sp, omc, sg, nn, ig, ires, upperlimit, wg, win, lane, Sd);
}

constexpr int REL_THREADS = 256;

// nsplit > 1 (few omegas in flight: sequential root finding is latency bound): the gamma rows / grid points
// of one (omega, species, |n|) are dealt round-robin to nsplit CTAs; each leaves a partial row in Mpart and
// the last one to finish (ticket counter) adds them up in a fixed order.
// MINB = 2 (throughput batches): 128 registers, two CTAs per SM; MINB = 1 (latency batches): 255 registers
template <int MINB>
__global__ void __launch_bounds__(REL_THREADS, MINB) k_rel(const GlobalDev* __restrict__ gp, const double* __restrict__ om,
                                                     int n_om, const RelTile* __restrict__ tiles, int ntiles,
                                                     double* __restrict__ Mrel, int* __restrict__ err_flag, int nsplit,
                                                     double* __restrict__ Mpart, int* __restrict__ tickets,
                                                     const unsigned char* __restrict__ rflag, int skip_res) {
  // rflag (from k_rel_plan): the resonance flags of every (omega, tile); skip_res: the resonant signs are left to
  // k_rel_rows, which spreads their Gamma rows over the whole GPU
  const GlobalDev& g = *gp;
  pdl_trigger();
  pdl_wait();
  const int js = blockIdx.x % nsplit;
  const int iom = (blockIdx.x / nsplit) / ntiles;
  const int tile_id = (blockIdx.x / nsplit) % ntiles;
  const RelTile tl = tiles[tile_id];
  const SpeciesDev& sp = g.sp[tl.s];
  const int nabs = tl.nabs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = REL_THREADS / 32;
  const cd omc = mk(om[2 * iom], om[2 * iom + 1]);
  const int nperp = g.nperp, npar = g.npar, ng = g.ngamma, npb = g.npparbar, M_I = g.M_I, M_P = g.M_P;
  const double qs = sp.qs, ms = sp.ms, vA = g.vA, kpar = g.kpar, kperp = g.kperp;
  __shared__ int s_found[2];
  __shared__ cd s_red[REL_THREADS / 32][6];
  __shared__ RelWin s_win[REL_THREADS / 32];           // principal-value window of the row a warp works on
  __shared__ double s_rfact[21], s_rgam[23];           // 1/k!, 1/Gamma(|n| + i): series of the Landau term
  if (tid < 21) {
    double fact = 1.0;
    for (int k = 2; k <= tid; k++) fact = fact * (1.0 * k);   // Fact, src/ALPS_fns_rel.f90:1560-1574
    s_rfact[tid] = 1.0 / fact;
  } else if (tid >= 32 && tid < 32 + 23) {
    const int m = nabs + (tid - 32);                          // Gamma(m), m = |n| .. |n| + 22
    s_rgam[tid - 32] = m >= 1 ? 1.0 / gamma_ref(1.0 * m) : 0.0;
  }
  if (tid < 2) s_found[tid] = 0;
  __syncthreads();

  // ---- determine_resonances, relativistic branch: any (iperp, ipar) cell containing Re p_res
  if (rflag) {
    if (tid < 2) s_found[tid] = (rflag[(size_t)iom * ntiles + tile_id] >> tid) & 1;
  } else {
    int fp = 0, fm = 0;
    for (int idx = tid; idx < (nperp + 1) * npar; idx += REL_THREADS) {
      const int iperp = idx / npar, ipar = idx - iperp * npar;
      const double pp1 = sp.pperp[iperp], pp2 = sp.ppar[ipar];
      const double gamma = sqrt((pp1 * pp1 + pp2 * pp2) * (vA * vA) / (ms * ms) + 1.0);
      const double prp = (gamma * ms * omc.x - 1.0 * nabs * qs) / kpar;
      const double prm = (gamma * ms * omc.x + 1.0 * nabs * qs) / kpar;
      if (sp.ppar[ipar] <= prp && sp.ppar[ipar + 1] > prp) fp = 1;
      if (sp.ppar[ipar] <= prm && sp.ppar[ipar + 1] > prm) fm = 1;
    }
    if (fp) s_found[0] = 1;
    if (fm) s_found[1] = 1;
  }
  __syncthreads();

  RelCtx c;
  c.g = gp;
  c.sp = &sp;
  c.om = omc;
  c.nabs = nabs;
  c.pref = -2.0 * PI_ * ((ms / vA) * (ms / vA) * (ms / vA)) * (qs * vA / (kpar * ms));
  c.zfac = kperp * ms / (vA * qs);
  c.zbar = g.kperp_norm ? kperp * ms / (vA * qs) : ms / (vA * qs);
  c.kf1 = g.kperp_norm ? 1.0 : kperp;
  c.kf2 = g.kperp_norm ? 1.0 : kperp * kperp;

  for (int sg = 0; sg < 2; sg++) {
    if (nabs == 0 && sg == 1) break;
    const double nn = sg ? -(double)nabs : (double)nabs;
    if (s_found[sg] && skip_res) continue;
    Six2 acc;
    zero6(acc);
    if (!s_found[sg]) {
      // ---- integrate() on the (p_perp, p_par) grid with gamma in resU
      const double zb = g.kperp_norm ? kperp / qs : 1.0 / qs;
      const double* Jn = sp.J + (size_t)(nabs + 1) * sp.ldj;
      const double* Jm = sp.J + (size_t)nabs * sp.ldj;
      const double* Jp = sp.J + (size_t)(nabs + 2) * sp.ldj;
      for (int idx = tid + REL_THREADS * js; idx < (nperp - 1) * (npar - 1); idx += REL_THREADS * nsplit) {
        const int iperp = idx / (npar - 1) + 1, ipar = idx % (npar - 1) + 1;
        const double wperp = (iperp == nperp - 1) ? 1.0 : 2.0;
        const double wpar = (ipar == 1 || ipar == npar - 1) ? 1.0 : 2.0;
        const double pp1 = sp.pperp[iperp], pp2 = sp.ppar[ipar];
        const double gamma = sqrt((pp1 * pp1 + pp2 * pp2) * (vA * vA) / (ms * ms) + 1.0);
        const size_t o = (size_t)(iperp - 1) * sp.ldp + (ipar - 1);
        // resU = (om A' + (kpar/gamma) C0) / (gamma ms om - kpar p_par - n qs)
        const cd num = omc * sp.A[o] + mk((kpar / gamma) * sp.C0[o], 0.0);
        const cd den = mk(gamma * ms * omc.x - kpar * pp2 - nn * qs, gamma * ms * omc.y);
        const cd U = (wperp * wpar) * (num / den);
        double bj, bp;
        bessel_pair(nabs, sg, 0.0, nabs >= 1 ? Jm[iperp] : 0.0, Jn[iperp], Jp[iperp], bj, bp);
        Six2 T;
        modes_real(bj, bp, pp1, pp2, zb, nn, c.kf1, c.kf2, T);
#pragma unroll
        for (int q = 0; q < 6; q++) acc.v[q] += U * T.v[q];
      }
      const double fac = 2.0 * PI_ * sp.dpperp * sp.dppar_abs * 0.25;
#pragma unroll
      for (int q = 0; q < 6; q++) acc.v[q] = fac * acc.v[q];
    } else {
      // ---- integrate_res_rel: outer trapezoid over gamma, one warp per igamma
      const double dpb = sp.dpparbar, dgam = sp.dgamma;
      const double* pbv = sp.pbrel;
      const int ldr = npb + 1;
      Six2 Sd;   // Bessel-moment sums of the direct and principal parts
      zero6(Sd);
      for (int ig = 1 + warp + nwarps * js; ig <= ng - 1; ig += nwarps * nsplit)
        rel_res_row(c, g, sp, omc, sg, nn, ig, lane, s_win[warp], Sd, err_flag);
      // moment sums of the direct and principal parts -> tensor components
      moments_to_modes(c, nn, Sd, acc);
#pragma unroll
      for (int q = 0; q < 6; q++) acc.v[q] = (dgam * 0.25) * acc.v[q];

      // ---- landau_integrate_rel (Im om <= 0), all threads stride over igamma
      if (omc.y <= 0.0) landau_rows(c, g, sp, omc, sg, nn, tid + REL_THREADS * js, REL_THREADS * nsplit, s_rfact, s_rgam, acc);
    }
    // ---- block reduction and store
#pragma unroll
    for (int q = 0; q < 6; q++) {
      cd v = warp_sum_cd(acc.v[q]);
      if (lane == 0) s_red[warp][q] = v;
    }
    __syncthreads();
    if (tid < 6) {
      cd t = mk(0.0, 0.0);
      for (int w = 0; w < nwarps; w++) t += s_red[w][tid];
      const size_t item = (size_t)iom * g.NI + sp.item_base + 2 * nabs + sg;
      double* o = (nsplit == 1) ? Mrel + item * 12 : Mpart + (item * nsplit + js) * 12;
      o[2 * tid] = t.x;
      o[2 * tid + 1] = t.y;
    }
    __syncthreads();
  }
  if (nsplit > 1) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&tickets[iom * ntiles + tile_id], 1) == nsplit - 1);
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (tid < 24) {
        const int sg = tid / 12, q = tid % 12;
        if (!(nabs == 0 && sg == 1) && !(s_found[sg] && skip_res)) {
          const size_t item = (size_t)iom * g.NI + sp.item_base + 2 * nabs + sg;
          double t = 0.0;
          for (int j = 0; j < nsplit; j++) t += __ldcg(Mpart + (item * nsplit + j) * 12 + q);
          Mrel[item * 12 + q] = t;
        }
      }
      if (tid == 0) tickets[iom * ntiles + tile_id] = 0;   // ready for the next launch
    }
  }
}

// =====================================================================================================================
// Throughput class (batches of more than 64 omegas): k_rel_plan -> k_rel_pv -> k_rel_direct -> k_rel_tiled.
//
// k_rel above gives one CTA to every (omega, species, |n|) and streams the (Gamma, pbar_par) tables -- 56 B per node for
// ~45 FP64 operations -- from L2 again for every omega: L2-bound at a fifth of the FP64 pipe (profiles/r01_k_rel_ncu_full.csv).
// Here the bulk of the work -- the regular quadrature of the non-resonant harmonics on the (p_perp, p_par) grid and the
// direct part of integrate_resU_rel on the (Gamma, pbar_par) grid -- is tiled over OMEGA instead: one CTA = 128 omegas
// (one per thread) x one (species, |n|).  The omega-independent half of every grid node -- numerator coefficients,
// denominator coefficients and the six real Bessel moments of which the T components are constant multiples; identical
// for +n and -n because J_-n = (-1)^n J_n enters them squared or as J J' -- is computed ONCE per node by one thread into a
// shared-memory packet, and every thread then applies its own omega to the packet (warp-uniform LDS broadcasts): per
// (node, omega, sign) one reciprocal and 12 FMAs.  The principal-value window and the Landau term of the resonant
// harmonics keep their warp-per-Gamma-row form (k_rel_pv) but run only for the resonant (omega, species, n, sign)
// entries that k_rel_plan lists.
// ---------------------------------------------------------------------------------------------------------------------

// determine_resonances, relativistic branch (src/ALPS_fns.f90:683-701): one warp per (omega, tile); flags + work list
__global__ void __launch_bounds__(256) k_rel_plan(const GlobalDev* __restrict__ gp, const double* __restrict__ om, int n_om,
                                                  const RelTile* __restrict__ tiles, int ntiles,
                                                  unsigned char* __restrict__ rflag, int* __restrict__ rwork,
                                                  int* __restrict__ rcount, int* __restrict__ rpos) {
  const GlobalDev& g = *gp;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  // (small batches launch it programmatically inside the captured chain: completion of this kernel must imply that of
  // its predecessors)
  pdl_trigger();
  pdl_wait();
  if (w >= n_om * ntiles) return;
  const int iom = w / ntiles, tile_id = w % ntiles;
  const RelTile tl = tiles[tile_id];
  const SpeciesDev& sp = g.sp[tl.s];
  const int nabs = tl.nabs, nperp = g.nperp, npar = g.npar;
  const double omr = om[2 * iom], qs = sp.qs, ms = sp.ms, vA = g.vA, kpar = g.kpar;
  int fp = 0, fm = 0;
  for (int idx = lane; idx < (nperp + 1) * npar; idx += 32) {
    const int iperp = idx / npar, ipar = idx - iperp * npar;
    const double pp1 = sp.pperp[iperp], pp2 = sp.ppar[ipar];
    const double gamma = sqrt((pp1 * pp1 + pp2 * pp2) * (vA * vA) / (ms * ms) + 1.0);
    const double prp = (gamma * ms * omr - 1.0 * nabs * qs) / kpar;
    const double prm = (gamma * ms * omr + 1.0 * nabs * qs) / kpar;
    if (sp.ppar[ipar] <= prp && sp.ppar[ipar + 1] > prp) fp = 1;
    if (sp.ppar[ipar] <= prm && sp.ppar[ipar + 1] > prm) fm = 1;
  }
  fp = __any_sync(0xffffffffu, fp);
  fm = nabs > 0 ? __any_sync(0xffffffffu, fm) : 0;
  if (lane == 0) {
    rflag[w] = (unsigned char)(fp | (fm << 1));
    // the resonant (omega, sign) entries of a tile: one segment of 2 n_om slots per tile, filled in any order (every
    // entry is evaluated on its own), and where each entry sits (k_rel_tiled collects its partial rows from there)
    int* seg = rwork + (size_t)tile_id * 2 * n_om;
    int pp = -1, pm = -1;
    if (fp) seg[pp = atomicAdd(rcount + tile_id, 1)] = (iom << 1);
    if (fm) seg[pm = atomicAdd(rcount + tile_id, 1)] = (iom << 1) | 1;
    rpos[2 * (size_t)w] = pp;
    rpos[2 * (size_t)w + 1] = pm;
  }
}

// The same flags and lists with one 256-thread block per (omega, tile): small batches have few (omega, tile) pairs and
// the 1860-cell scan of a warp is then a chain of its own (12 us on C3)
__global__ void __launch_bounds__(256) k_rel_plan_blk(const GlobalDev* __restrict__ gp, const double* __restrict__ om,
                                                      int n_om, const RelTile* __restrict__ tiles, int ntiles,
                                                      unsigned char* __restrict__ rflag, int* __restrict__ rwork,
                                                      int* __restrict__ rcount, int* __restrict__ rpos) {
  const GlobalDev& g = *gp;
  pdl_trigger();
  pdl_wait();
  const int w = blockIdx.x, tid = threadIdx.x;
  const int iom = w / ntiles, tile_id = w % ntiles;
  const RelTile tl = tiles[tile_id];
  const SpeciesDev& sp = g.sp[tl.s];
  const int nabs = tl.nabs, nperp = g.nperp, npar = g.npar;
  const double omr = om[2 * iom], qs = sp.qs, ms = sp.ms, vA = g.vA, kpar = g.kpar;
  int fp = 0, fm = 0;
  for (int idx = tid; idx < (nperp + 1) * npar; idx += 256) {
    const int iperp = idx / npar, ipar = idx - iperp * npar;
    const double pp1 = sp.pperp[iperp], pp2 = sp.ppar[ipar];
    const double gamma = sqrt((pp1 * pp1 + pp2 * pp2) * (vA * vA) / (ms * ms) + 1.0);
    const double prp = (gamma * ms * omr - 1.0 * nabs * qs) / kpar;
    const double prm = (gamma * ms * omr + 1.0 * nabs * qs) / kpar;
    if (sp.ppar[ipar] <= prp && sp.ppar[ipar + 1] > prp) fp = 1;
    if (sp.ppar[ipar] <= prm && sp.ppar[ipar + 1] > prm) fm = 1;
  }
  fp = __syncthreads_or(fp);
  fm = __syncthreads_or(fm);
  if (nabs == 0) fm = 0;
  if (tid == 0) {
    rflag[w] = (unsigned char)((fp ? 1 : 0) | (fm ? 2 : 0));
    int* seg = rwork + (size_t)tile_id * 2 * n_om;
    int pp = -1, pm = -1;
    if (fp) seg[pp = atomicAdd(rcount + tile_id, 1)] = (iom << 1);
    if (fm) seg[pm = atomicAdd(rcount + tile_id, 1)] = (iom << 1) | 1;
    rpos[2 * (size_t)w] = pp;
    rpos[2 * (size_t)w + 1] = pm;
  }
}

// limits of integrate_resU_rel for one Gamma row (src/ALPS_fns_rel.f90:591-672); returns 0 on alps_error(8)
struct RelRow {
  int int_start, int_end, lowerlimit, upperlimit;
  int ires, pv;      // cell of the resonance; principal_integral_rel applies (resonance well inside the cone)
};
__device__ __forceinline__ int rel_row_limits(const SpeciesDev& sp, int npb, int M_I, double presx, double g1, int lo_c,
                                              int up_c, RelRow& r) {
  const double* __restrict__ pbv = sp.pbrel;
  const double dpb = sp.dpparbar;
  int ires = 0, found = 0;
  r.ires = 0;
  r.pv = 0;
  if (presx * presx <= g1 * g1 - 1.0) {
    if (presx >= pbv[1] && presx < pbv[npb - 1]) {
      // the cell [pbv(lo), pbv(lo+1)) that holds Re p_res, lo in [1, npb-2]: index guess on the uniform grid, then the
      // reference's comparisons on the actual node values
      int lo = (int)floor((presx - pbv[0]) / dpb);
      lo = min(max(lo, 1), npb - 2);
      while (lo > 1 && pbv[lo] > presx) lo--;
      while (lo < npb - 2 && pbv[lo + 1] <= presx) lo++;
      if (pbv[lo + 1] > presx && pbv[lo] <= presx) {
        ires = lo;
        found = 1;
      }
    }
  }
  for (int ip = 0; ip <= M_I; ip++) {
    if (presx >= pbv[0] - dpb * ip && presx < pbv[0] - dpb * (ip - 1)) {
      ires = -ip;
      found = 1;
    }
    if (presx >= pbv[npb - 1] + dpb * ip && presx < pbv[npb - 1] + dpb * (ip + 1)) {
      ires = npb - 1 + ip;
      found = 1;
    }
  }
  if (found) {
    r.ires = ires;
    r.int_start = lo_c;
    r.int_end = up_c;
    r.lowerlimit = ires - M_I;
    r.upperlimit = ires + M_I + 1;
    if (ires >= 0 && ires <= npb)
      if (fabs(presx - pbv[ires]) > 0.5 * dpb) r.upperlimit = r.upperlimit + 1;
    if (r.lowerlimit < lo_c && r.upperlimit > up_c) return 0;
    if (r.lowerlimit <= lo_c) {
      r.int_start = 1;
      r.lowerlimit = 0;
      r.upperlimit = lo_c;
    } else if (r.upperlimit >= up_c) {
      r.lowerlimit = up_c;
      r.upperlimit = npb;
      r.int_end = npb - 1;
    }
  } else {
    r.int_start = lo_c;
    r.lowerlimit = up_c;
    r.int_end = npb - 1;
    r.upperlimit = npb;
  }
  r.pv = found && r.lowerlimit >= r.int_start && r.upperlimit <= r.int_end;
  return 1;
}

constexpr int RT_THREADS = 128, RT_CH = 128;
constexpr int RT_WA = 12;   // packet of a (p_perp, p_par) node: w, a, c, dgm, e, pad, q0..q5
constexpr int RT_WB = 10;   // packet of a (Gamma, pbar_par) node: pb, dfg, kv dfp, pad, M0..M5

__global__ void __launch_bounds__(RT_THREADS, 3)
    k_rel_tiled(const GlobalDev* __restrict__ gp, const double* __restrict__ om, int n_om,
                const RelTile* __restrict__ tiles, int ntiles, const unsigned char* __restrict__ rflag,
                double* __restrict__ Mrel, const int* __restrict__ rpos, const double* __restrict__ dpart,
                int nsplitB) {
  __shared__ __align__(16) double s_pk[2][RT_CH * RT_WA];
  const GlobalDev& g = *gp;
  const RelTile tl = tiles[blockIdx.y];
  const SpeciesDev& sp = g.sp[tl.s];
  const int nabs = tl.nabs, tid = threadIdx.x;
  const int iom = blockIdx.x * RT_THREADS + tid;
  const bool live = iom < n_om;
  const double omr = live ? om[2 * iom] : 1.0, omi = live ? om[2 * iom + 1] : 1.0;
  const int flags = live ? rflag[(size_t)iom * ntiles + blockIdx.y] : 0;
  const int nperp = g.nperp, npar = g.npar, ng = g.ngamma, npb = g.npparbar, M_I = g.M_I;
  const double qs = sp.qs, ms = sp.ms, vA = g.vA, kpar = g.kpar, kperp = g.kperp;
  const double nq = (double)nabs * qs;
  // moment sums per sign: {q0 .. q5} (re, im)
  double SP[12], SM[12];
#pragma unroll
  for (int q = 0; q < 12; q++) SP[q] = SM[q] = 0.0;
  const bool resP = (flags & 1) != 0, resM = (flags & 2) != 0;
  const bool nonP = live && !resP, nonM = live && nabs > 0 && !resM;

  // ---------------------------------------------------------------- (p_perp, p_par) grid: integrate() with gamma in resU
  if (__syncthreads_or(nonP || nonM)) {
    const double zb = g.kperp_norm ? kperp / qs : 1.0 / qs;
    (void)zb;
    const double* __restrict__ Jn = sp.J + (size_t)(nabs + 1) * sp.ldj;
    const double* __restrict__ Jm = sp.J + (size_t)nabs * sp.ldj;
    const double* __restrict__ Jp = sp.J + (size_t)(nabs + 2) * sp.ldj;
    const int nnode = (nperp - 1) * (npar - 1);
    const int nch = (nnode + RT_CH - 1) / RT_CH;
    auto stage = [&](int c, double* pk) {
      const int idx = c * RT_CH + tid;
      double* o = pk + tid * RT_WA;
      if (idx < nnode) {
        const int iperp = idx / (npar - 1) + 1, ipar = idx % (npar - 1) + 1;
        const double wperp = (iperp == nperp - 1) ? 1.0 : 2.0;
        const double wpar = (ipar == 1 || ipar == npar - 1) ? 1.0 : 2.0;
        const double pp1 = sp.pperp[iperp], pp2 = sp.ppar[ipar];
        const double gamma = sqrt((pp1 * pp1 + pp2 * pp2) * (vA * vA) / (ms * ms) + 1.0);
        const size_t oo = (size_t)(iperp - 1) * sp.ldp + (ipar - 1);
        double bj, bp;
        bessel_pair(nabs, 0, 0.0, nabs >= 1 ? Jm[iperp] : 0.0, Jn[iperp], Jp[iperp], bj, bp);
        o[0] = wperp * wpar;
        o[1] = sp.A[oo];
        o[2] = (kpar / gamma) * sp.C0[oo];
        o[3] = gamma * ms;
        o[4] = kpar * pp2;
        o[5] = 0.0;
        o[6] = bj * bj;                       // xx: n^2 J^2 / z^2
        o[7] = bp * bp * pp1 * pp1;           // yy
        o[8] = bj * bj * (pp2 * pp2);         // zz
        o[9] = bj * bp * pp1;                 // xy (times i n / z)
        o[10] = bj * bj * pp2;                // xz (times n / z)
        o[11] = bj * bp * pp2 * pp1;          // yz (times -i)
      } else {
#pragma unroll
        for (int q = 0; q < RT_WA; q++) o[q] = 0.0;
        o[3] = 1.0;   // den = om: finite for the padding nodes (weight 0)
      }
    };
    stage(0, s_pk[0]);
    __syncthreads();
    for (int c = 0; c < nch; c++) {
      if (c + 1 < nch) stage(c + 1, s_pk[(c + 1) & 1]);
      const double* __restrict__ pk = s_pk[c & 1];
      if (nonP || nonM) {
#pragma unroll 2
        for (int j = 0; j < RT_CH; j++) {
          const double2* t2 = reinterpret_cast<const double2*>(pk + j * RT_WA);
          const double2 h0 = t2[0], h1 = t2[1], h2 = t2[2];      // (w, a) (c, dgm) (e, -)
          const double nr = fma(omr, h0.y, h1.x), ni = omi * h0.y;
          const double x = fma(h1.y, omr, -h2.x), di = h1.y * omi, di2 = di * di;
          const double drp = x - nq, drm = x + nq;
          // both reciprocals from one (a non-resonant denominator is never 0; padding nodes have den = om)
          // a resonant sign (k_rel's and the (Gamma, pbar_par) part's business) may have den = 0 on this grid: it must not
          // poison the shared reciprocal
          const double dp = nonP ? fma(drp, drp, di2) : 1.0, dm = nonM ? fma(drm, drm, di2) : 1.0;
          const double inv = h0.x * fast_rcp(dp * dm);
          const double tp = nonP ? dm * inv : 0.0, tm = nonM ? dp * inv : 0.0;
          const double nrdi = nr * di, nidi = ni * di;
          const double upr = fma(nr, drp, nidi) * tp, upi = fma(ni, drp, -nrdi) * tp;
          const double umr = fma(nr, drm, nidi) * tm, umi = fma(ni, drm, -nrdi) * tm;
#pragma unroll
          for (int q = 0; q < 3; q++) {
            const double2 m = t2[3 + q];
            SP[4 * q] = fma(upr, m.x, SP[4 * q]);         SP[4 * q + 1] = fma(upi, m.x, SP[4 * q + 1]);
            SP[4 * q + 2] = fma(upr, m.y, SP[4 * q + 2]); SP[4 * q + 3] = fma(upi, m.y, SP[4 * q + 3]);
            SM[4 * q] = fma(umr, m.x, SM[4 * q]);         SM[4 * q + 1] = fma(umi, m.x, SM[4 * q + 1]);
            SM[4 * q + 2] = fma(umr, m.y, SM[4 * q + 2]); SM[4 * q + 3] = fma(umi, m.y, SM[4 * q + 3]);
          }
        }
      }
      __syncthreads();
    }
  }
  // non-resonant signs: moment sums -> tensor components (modes_real), trapezoid factor
  cd TP[6], TM[6];
  {
    const double zb = g.kperp_norm ? kperp / qs : 1.0 / qs;
    const double kf1 = g.kperp_norm ? 1.0 : kperp, kf2 = g.kperp_norm ? 1.0 : kperp * kperp;
    const double fac = 2.0 * PI_ * sp.dpperp * sp.dppar_abs * 0.25;
#pragma unroll
    for (int sgn = 0; sgn < 2; sgn++) {
      const double* S = sgn ? SM : SP;
      cd* T = sgn ? TM : TP;
      const double nn = sgn ? -(double)nabs : (double)nabs;
      const double c0 = 1.0 * (nn * nn) / (zb * zb), c3 = kf1 * (1.0 * nn) / zb;
      T[0] = fac * mk(c0 * S[0], c0 * S[1]);
      T[1] = fac * mk(kf2 * S[2], kf2 * S[3]);
      T[2] = fac * mk(kf2 * S[4], kf2 * S[5]);
      T[3] = fac * cmul_i(mk(c3 * S[6], c3 * S[7]));
      T[4] = fac * mk(c3 * S[8], c3 * S[9]);
      T[5] = fac * (-cmul_i(mk(kf2 * S[10], kf2 * S[11])));
    }
  }

  if (!live) return;
  // non-resonant signs: done.  Resonant signs: k_rel_pv has left the principal-value and Landau parts in Mrel; add the
  // direct part, i.e. the partial rows of the Gamma splits of k_rel_direct in their fixed order
#pragma unroll
  for (int sgn = 0; sgn < 2; sgn++) {
    if (sgn == 1 && nabs == 0) break;
    const bool res = sgn ? resM : resP;
    double* o = Mrel + ((size_t)iom * g.NI + sp.item_base + 2 * nabs + sgn) * 12;
    if (!res) {
      const cd* T = sgn ? TM : TP;
#pragma unroll
      for (int q = 0; q < 6; q++) {
        o[2 * q] = T[q].x;
        o[2 * q + 1] = T[q].y;
      }
    } else {
      const int pos = rpos[2 * ((size_t)iom * ntiles + blockIdx.y) + sgn];
      const double* pr = dpart + (((size_t)blockIdx.y * 2 * n_om + pos) * nsplitB) * 12;
      double t[12];
#pragma unroll
      for (int q = 0; q < 12; q++) t[q] = 0.0;
      for (int js = 0; js < nsplitB; js++)
#pragma unroll
        for (int q = 0; q < 12; q++) t[q] += pr[(size_t)js * 12 + q];
#pragma unroll
      for (int q = 0; q < 12; q++) o[q] += t[q];
    }
  }
}


// Direct part of integrate_resU_rel for the resonant entries of one tile: one CTA = 128 entries (one per thread: its
// omega and sign) x the Gamma rows js, js + nsplitB, ... ; the row's node packets are staged once per CTA.  Leaves the
// tensor-component partial row of (entry, js) in dpart; k_rel_tiled adds the rows of an entry in order.
__global__ void __launch_bounds__(RT_THREADS, 3)
    k_rel_direct(const GlobalDev* __restrict__ gp, const double* __restrict__ om, int n_om,
                 const RelTile* __restrict__ tiles, int ntiles, const int* __restrict__ rwork,
                 const int* __restrict__ rcount, double* __restrict__ dpart, int nsplitB, int* __restrict__ err_flag) {
  const int tile_id = blockIdx.y, js = blockIdx.z;
  const int cnt_e = rcount[tile_id];
  if (blockIdx.x * RT_THREADS >= cnt_e) return;
  __shared__ __align__(16) double s_pk[RT_CH * RT_WB];
  const GlobalDev& g = *gp;
  const RelTile tl = tiles[tile_id];
  const SpeciesDev& sp = g.sp[tl.s];
  const int nabs = tl.nabs, tid = threadIdx.x;
  const int e = blockIdx.x * RT_THREADS + tid;
  const bool live = e < cnt_e;
  const int ecode = live ? rwork[(size_t)tile_id * 2 * n_om + e] : 0;
  const int iom = ecode >> 1, sg = ecode & 1;
  const double omr = om[2 * iom], omi = om[2 * iom + 1];
  const int ng = g.ngamma, npb = g.npparbar, M_I = g.M_I;
  const double qs = sp.qs, ms = sp.ms, vA = g.vA, kpar = g.kpar, kperp = g.kperp;
  double S[12];
#pragma unroll
  for (int q = 0; q < 12; q++) S[q] = 0.0;
  const double pref = -2.0 * PI_ * ((ms / vA) * (ms / vA) * (ms / vA)) * (qs * vA / (kpar * ms));
  const double dpb = sp.dpparbar, kv = kpar / vA;
  const double* __restrict__ pbv = sp.pbrel;
  const int ldr = npb + 1;
  const size_t plane = (size_t)(ng + 1) * ldr;
  const double nn = sg ? -(double)nabs : (double)nabs;
  const double nqv = nn * qs * vA / (kpar * ms);
  for (int ig = 1 + js; ig <= ng - 1; ig += nsplitB) {
    const double wg = (ig == ng - 1) ? 1.0 : 2.0, g1 = sp.grel[ig], cw = wg * dpb;
    const int lo_c = sp.cone_lo[ig], up_c = sp.cone_up[ig];
    const double gomx = (g1 * omr) * vA / kpar, gomy = (g1 * omi) * vA / kpar;
    RelRow r;
    r.int_start = 1; r.lowerlimit = 0; r.upperlimit = npb; r.int_end = npb - 1;
    bool ok = live;
    if (live && !rel_row_limits(sp, npb, M_I, gomx - nqv, g1, lo_c, up_c, r)) {
      err_flag[0] = 8;   // alps_error(8)
      ok = false;
    }
    // every weighted node of the row lies in the cone [lo_c, up_c], whatever the omega: the pieces
    // [int_start, lowerlimit] and [upperlimit, int_end] are cut out of it (or empty)
    const int ra = max(1, lo_c), rb = min(npb - 1, up_c);
    const double* __restrict__ J0 = sp.Jrel + (size_t)nabs * plane + (size_t)ig * ldr;
    const double* __restrict__ JP = J0 + plane;
    const double* __restrict__ JM = nabs >= 1 ? J0 - plane : J0;
    const double* __restrict__ PPq = sp.Jrel + (size_t)(sp.nhi + 2) * plane + (size_t)ig * ldr;
    const double* __restrict__ DG = sp.dfg_rel + (size_t)ig * ldr;
    const double* __restrict__ DP = sp.dfp_rel + (size_t)ig * ldr;
    for (int base = ra; base <= rb; base += RT_CH) {
      __syncthreads();      // the previous chunk has been consumed
      {
        const int ip = base + tid;
        double* o = s_pk + tid * RT_WB;
        if (ip <= rb) {
          const double j0 = J0[ip], jp = JP[ip], jm = nabs >= 1 ? JM[ip] : 0.0;
          double bj, bp;
          bessel_pair(nabs, 0, 0.0, jm, j0, jp, bj, bp);     // the moments are the same for +n and -n
          const double pb = pbv[ip], pq = PPq[ip];
          const double b2 = bj * bj, bb = bj * bp * pq, q2 = (bp * pq) * (bp * pq);
          o[0] = pb;
          o[1] = DG[ip];
          o[2] = kv * DP[ip];
          o[3] = 0.0;
          o[4] = b2;
          o[5] = b2 * pb;
          o[6] = (b2 * pb) * pb;
          o[7] = bb;
          o[8] = bb * pb;
          o[9] = q2;
        }
      }
      __syncthreads();
      if (!ok) continue;
      // this thread's weighted nodes of the chunk: [int_start, lowerlimit] and [upperlimit, int_end]
      const int c1 = min(RT_CH, rb - base + 1);
#pragma unroll
      for (int piece = 0; piece < 2; piece++) {
        const int pa = piece ? r.upperlimit : r.int_start, pb_ = piece ? r.int_end : r.lowerlimit;
        if (pa > pb_) continue;
        const int j0 = max(pa - base, 0), j1 = min(pb_ - base, c1 - 1);
        for (int j = j0; j <= j1; j++) {
          const int ip = base + j;
          const double w = (pa == pb_) ? 1.0 : ((ip == pa || ip == pb_) ? 1.0 : 2.0);     // piece_w
          const double2* t2 = reinterpret_cast<const double2*>(s_pk + j * RT_WB);
          const double2 h0 = t2[0], h1 = t2[1];      // (pb, dfg) (kv dfp, -)
          const double nr = pref * fma(omr, h0.y, h1.x), ni = pref * (omi * h0.y);
          const double di = -gomy, dr = h0.x - gomx + nqv;
          const double t = (cw * w) * fast_rcp(fma(dr, dr, di * di));
          const double ur = fma(nr, dr, ni * di) * t, ui = fma(ni, dr, -(nr * di)) * t;
#pragma unroll
          for (int q = 0; q < 3; q++) {
            const double2 m = t2[2 + q];
            S[4 * q] = fma(ur, m.x, S[4 * q]);         S[4 * q + 1] = fma(ui, m.x, S[4 * q + 1]);
            S[4 * q + 2] = fma(ur, m.y, S[4 * q + 2]); S[4 * q + 3] = fma(ui, m.y, S[4 * q + 3]);
          }
        }
      }
    }
  }
  if (!live) return;
  // moment sums {J^2, J^2 p, J^2 p^2, J J' pp, J J' pp p, J'^2 pp^2} -> tensor components (moments_to_modes)
  const double zbar = g.kperp_norm ? kperp * ms / (vA * qs) : ms / (vA * qs);
  const double kf1 = g.kperp_norm ? 1.0 : kperp, kf2 = g.kperp_norm ? 1.0 : kperp * kperp;
  const double rowfac = sp.dgamma * 0.25;
  const double c0 = (nn * nn) / (zbar * zbar), c3 = kf1 * nn / zbar;
  cd T[6];
  T[0] = rowfac * (c0 * mk(S[0], S[1]));
  T[1] = rowfac * (kf2 * mk(S[10], S[11]));
  T[2] = rowfac * (kf2 * mk(S[4], S[5]));
  T[3] = rowfac * cmul_i(c3 * mk(S[6], S[7]));
  T[4] = rowfac * (c3 * mk(S[2], S[3]));
  T[5] = rowfac * (-cmul_i(kf2 * mk(S[8], S[9])));
  double* o = dpart + ((((size_t)tile_id * 2 * n_om + e) * nsplitB) + js) * 12;
#pragma unroll
  for (int q = 0; q < 6; q++) {
    o[2 * q] = T[q].x;
    o[2 * q + 1] = T[q].y;
  }
}

// Principal-value window, tiny rest and Landau term of the resonant entries (persistent CTAs over the per-tile lists of
// k_rel_plan).  The row bookkeeping of integrate_resU_rel (cell of the resonance, limits, whether the principal part
// applies) is computed for 32 Gamma rows at a time, one row per lane, and handed to the warp by shuffles; the warp then
// works through the rows that have a principal part with pv_row.  Leaves the entry's tensor components in Mrel
// (k_rel_tiled adds the direct part).
template <int MINB>
__global__ void __launch_bounds__(REL_THREADS, MINB)
    k_rel_pv(const GlobalDev* __restrict__ gp, const double* __restrict__ om, int n_om, const RelTile* __restrict__ tiles,
             int ntiles, double* __restrict__ Mrel, int* __restrict__ err_flag, const int* __restrict__ rwork,
             const int* __restrict__ rcount) {
  const GlobalDev& g = *gp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = REL_THREADS / 32;
  __shared__ cd s_red[REL_THREADS / 32][6];
  __shared__ RelWin s_win[REL_THREADS / 32];
  __shared__ double s_rfact[21], s_rgam[23];
  const int ng = g.ngamma, npb = g.npparbar, M_I = g.M_I;
  const double vA = g.vA, kpar = g.kpar, kperp = g.kperp;
  int last_nabs = -1;
  for (int etile = 0; etile < ntiles; etile++) {
    const int cnt = rcount[etile];
    if (blockIdx.x >= cnt) continue;
    const RelTile tl = tiles[etile];
    const SpeciesDev& sp = g.sp[tl.s];
    const int nabs = tl.nabs;
    const double qs = sp.qs, ms = sp.ms;
    if (nabs != last_nabs) {      // series coefficients of the Landau term: 1/k!, 1/Gamma(|n| + i)
      __syncthreads();
      if (tid < 21) {
        double fact = 1.0;
        for (int k = 2; k <= tid; k++) fact = fact * (1.0 * k);
        s_rfact[tid] = 1.0 / fact;
      } else if (tid >= 32 && tid < 32 + 23) {
        const int m = nabs + (tid - 32);
        s_rgam[tid - 32] = m >= 1 ? 1.0 / gamma_ref(1.0 * m) : 0.0;
      }
      __syncthreads();
      last_nabs = nabs;
    }
    RelCtx c;
    c.g = gp;
    c.sp = &sp;
    c.nabs = nabs;
    c.pref = -2.0 * PI_ * ((ms / vA) * (ms / vA) * (ms / vA)) * (qs * vA / (kpar * ms));
    c.zfac = kperp * ms / (vA * qs);
    c.zbar = g.kperp_norm ? kperp * ms / (vA * qs) : ms / (vA * qs);
    c.kf1 = g.kperp_norm ? 1.0 : kperp;
    c.kf2 = g.kperp_norm ? 1.0 : kperp * kperp;
    for (int entry = blockIdx.x; entry < cnt; entry += gridDim.x) {
      const int ecode = rwork[(size_t)etile * 2 * n_om + entry];
      const int iom = ecode >> 1, sg = ecode & 1;
      const cd omc = mk(om[2 * iom], om[2 * iom + 1]);
      c.om = omc;
      const double nn = sg ? -(double)nabs : (double)nabs;
      const double nqv = nn * qs * vA / (kpar * ms);
      Six2 Sd, acc;
      zero6(Sd);
      // rows interleaved over the warps (warp w: rows 1 + w, 1 + w + nwarps, ...): the rows that need a principal part
      // cluster in Gamma, so consecutive blocks per warp would leave the warps unevenly loaded
      for (int b = 0; 1 + warp + nwarps * b <= ng - 1; b += 32) {
        // one row per lane: limits of integrate_resU_rel
        const int myig = 1 + warp + nwarps * (b + lane);
        RelRow r;
        r.ires = 0; r.pv = 0; r.upperlimit = 0;
        if (myig <= ng - 1) {
          const double g1 = sp.grel[myig];
          const double presx = (g1 * omc.x) * vA / kpar - nqv;
          if (!rel_row_limits(sp, npb, M_I, presx, g1, sp.cone_lo[myig], sp.cone_up[myig], r)) {
            err_flag[0] = 8;   // alps_error(8)
            r.pv = 0;
          }
        }
        unsigned todo = __ballot_sync(0xffffffffu, r.pv != 0);
        while (todo) {
          const int l = __ffs(todo) - 1;
          todo &= todo - 1;
          const int ig = 1 + warp + nwarps * (b + l);
          const int ires = __shfl_sync(0xffffffffu, r.ires, l), upper = __shfl_sync(0xffffffffu, r.upperlimit, l);
          pv_row(c, g, sp, omc, sg, nn, ig, ires, upper, (ig == ng - 1) ? 1.0 : 2.0, s_win[warp], lane, Sd);
        }
      }
      moments_to_modes(c, nn, Sd, acc);
#pragma unroll
      for (int q = 0; q < 6; q++) acc.v[q] = (sp.dgamma * 0.25) * acc.v[q];
      if (omc.y <= 0.0) landau_rows(c, g, sp, omc, sg, nn, tid, REL_THREADS, s_rfact, s_rgam, acc);
#pragma unroll
      for (int q = 0; q < 6; q++) {
        cd v = warp_sum_cd(acc.v[q]);
        if (lane == 0) s_red[warp][q] = v;
      }
      __syncthreads();
      if (tid < 6) {
        cd t = mk(0.0, 0.0);
        for (int w = 0; w < nwarps; w++) t += s_red[w][tid];
        double* o = Mrel + ((size_t)iom * g.NI + sp.item_base + 2 * nabs + sg) * 12;
        o[2 * tid] = t.x;
        o[2 * tid + 1] = t.y;
      }
      __syncthreads();
    }
  }
}

// The non-resonant signs of a small batch: integrate() on the (p_perp, p_par) grid with gamma in resU, exactly the
// non-resonant branch of k_rel (same operations, same split of the nodes over nsplit CTAs and the same order of the
// partial rows) as a kernel of its own -- a few hundred instructions instead of k_rel's thousands, which matters for
// code that runs once per launch, from L2.
__global__ void __launch_bounds__(REL_THREADS) k_rel_nonres(const GlobalDev* __restrict__ gp, const double* __restrict__ om,
                                                            int n_om, const RelTile* __restrict__ tiles, int ntiles,
                                                            double* __restrict__ Mrel, int nsplit, double* __restrict__ Mpart,
                                                            int* __restrict__ tickets, const unsigned char* __restrict__ rflag) {
  const GlobalDev& g = *gp;
  // launched after k_rel_rows has seen k_rel_plan_blk complete (see there): the flags are final; the wait at the end
  // makes this kernel's completion imply that of k_rel_rows
  pdl_trigger();
  const int js = blockIdx.x % nsplit;
  const int iom = (blockIdx.x / nsplit) / ntiles;
  const int tile_id = (blockIdx.x / nsplit) % ntiles;
  const RelTile tl = tiles[tile_id];
  const SpeciesDev& sp = g.sp[tl.s];
  const int nabs = tl.nabs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = REL_THREADS / 32;
  const cd omc = mk(om[2 * iom], om[2 * iom + 1]);
  const int nperp = g.nperp, npar = g.npar;
  const double qs = sp.qs, ms = sp.ms, vA = g.vA, kpar = g.kpar, kperp = g.kperp;
  const int found = rflag[(size_t)iom * ntiles + tile_id];
  __shared__ cd s_red[REL_THREADS / 32][6];
  __shared__ int s_last;
  const double kf1 = g.kperp_norm ? 1.0 : kperp, kf2 = g.kperp_norm ? 1.0 : kperp * kperp;
  const double zb = g.kperp_norm ? kperp / qs : 1.0 / qs;
  const double* Jn = sp.J + (size_t)(nabs + 1) * sp.ldj;
  const double* Jm = sp.J + (size_t)nabs * sp.ldj;
  const double* Jp = sp.J + (size_t)(nabs + 2) * sp.ldj;
#pragma unroll 1
  for (int sg = 0; sg < 2; sg++) {
    if (nabs == 0 && sg == 1) break;
    if ((found >> sg) & 1) continue;       // resonant: k_rel_rows
    const double nn = sg ? -(double)nabs : (double)nabs;
    Six2 acc;
    zero6(acc);
#pragma unroll 1
    for (int idx = tid + REL_THREADS * js; idx < (nperp - 1) * (npar - 1); idx += REL_THREADS * nsplit) {
      const int iperp = idx / (npar - 1) + 1, ipar = idx % (npar - 1) + 1;
      const double wperp = (iperp == nperp - 1) ? 1.0 : 2.0;
      const double wpar = (ipar == 1 || ipar == npar - 1) ? 1.0 : 2.0;
      const double pp1 = sp.pperp[iperp], pp2 = sp.ppar[ipar];
      const double gamma = sqrt((pp1 * pp1 + pp2 * pp2) * (vA * vA) / (ms * ms) + 1.0);
      const size_t o = (size_t)(iperp - 1) * sp.ldp + (ipar - 1);
      // resU = (om A' + (kpar/gamma) C0) / (gamma ms om - kpar p_par - n qs)
      const cd num = omc * sp.A[o] + mk((kpar / gamma) * sp.C0[o], 0.0);
      const cd den = mk(gamma * ms * omc.x - kpar * pp2 - nn * qs, gamma * ms * omc.y);
      const cd U = (wperp * wpar) * (num / den);
      double bj, bp;
      bessel_pair(nabs, sg, 0.0, nabs >= 1 ? Jm[iperp] : 0.0, Jn[iperp], Jp[iperp], bj, bp);
      Six2 T;
      modes_real(bj, bp, pp1, pp2, zb, nn, kf1, kf2, T);
#pragma unroll
      for (int q = 0; q < 6; q++) acc.v[q] += U * T.v[q];
    }
    const double fac = 2.0 * PI_ * sp.dpperp * sp.dppar_abs * 0.25;
#pragma unroll
    for (int q = 0; q < 6; q++) acc.v[q] = fac * acc.v[q];
#pragma unroll
    for (int q = 0; q < 6; q++) {
      cd v = warp_sum_cd(acc.v[q]);
      if (lane == 0) s_red[warp][q] = v;
    }
    __syncthreads();
    if (tid < 6) {
      cd t = mk(0.0, 0.0);
      for (int w = 0; w < nwarps; w++) t += s_red[w][tid];
      const size_t item = (size_t)iom * g.NI + sp.item_base + 2 * nabs + sg;
      double* o = (nsplit == 1) ? Mrel + item * 12 : Mpart + (item * nsplit + js) * 12;
      o[2 * tid] = t.x;
      o[2 * tid + 1] = t.y;
    }
    __syncthreads();
  }
  if (nsplit > 1) {
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&tickets[iom * ntiles + tile_id], 1) == nsplit - 1);
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (tid < 24) {
        const int sg = tid / 12, q = tid % 12;
        if (!(nabs == 0 && sg == 1) && !((found >> sg) & 1)) {
          const size_t item = (size_t)iom * g.NI + sp.item_base + 2 * nabs + sg;
          double t = 0.0;
          for (int j = 0; j < nsplit; j++) t += __ldcg(Mpart + (item * nsplit + j) * 12 + q);
          Mrel[item * 12 + q] = t;
        }
      }
      if (tid == 0) tickets[iom * ntiles + tile_id] = 0;   // ready for the next launch
    }
  }
  pdl_wait();
}

// Small batches (n <= 64 omegas: sequential root finding): the resonant (omega, species, |n|, sign) entries that
// k_rel_plan lists are few, and each is a chain of ~500 Gamma rows per warp when it is left to the CTAs of its own tile
// (k_rel) -- the critical path of a relativistic disp().  Here their rows are dealt in chunks of RR_CH to ALL the CTAs of
// a persistent grid: work item = (entry, chunk); a CTA leaves the chunk's tensor components (direct + principal part of
// its rows, Landau term of its rows) in Mpart and the last chunk of an entry to finish (ticket) adds the chunks in order.
constexpr int RR_CH = 8;           // Gamma rows per work item: one per warp
constexpr int RR_MAXTILES = 1024;  // (species, |n|) tiles the prefix table holds
template <int MINB>
__global__ void __launch_bounds__(REL_THREADS, MINB)
    k_rel_rows(const GlobalDev* __restrict__ gp, const double* __restrict__ om, int n_om, const RelTile* __restrict__ tiles,
               int ntiles, double* __restrict__ Mrel, int* __restrict__ err_flag, const int* __restrict__ rwork,
               const int* __restrict__ rcount, double* __restrict__ Mpart, int* __restrict__ tickets) {
  const GlobalDev& g = *gp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = REL_THREADS / 32;
  __shared__ int s_pref[RR_MAXTILES + 1];
  __shared__ cd s_red[REL_THREADS / 32][6];
  __shared__ RelWin s_win[REL_THREADS / 32];
  __shared__ double s_rfact[21], s_rgam[23];
  __shared__ int s_last;
  // (its dependent, k_rel_nonres, needs k_rel_plan_blk's flags and nothing of this kernel: it may start once this kernel
  // has seen k_rel_plan_blk complete, and runs beside it)
  pdl_wait();
  pdl_trigger();
  for (int t = tid; t < ntiles; t += REL_THREADS) s_pref[t + 1] = rcount[t];
  if (tid == 0) s_pref[0] = 0;
  __syncthreads();
  if (tid == 0)
    for (int t = 0; t < ntiles; t++) s_pref[t + 1] += s_pref[t];
  __syncthreads();
  const int ng = g.ngamma;
  const int nch = (ng - 1 + RR_CH - 1) / RR_CH;
  const long long total = (long long)s_pref[ntiles] * nch;
  const double vA = g.vA, kpar = g.kpar, kperp = g.kperp;
  int last_nabs = -1;
  for (long long w = blockIdx.x; w < total; w += gridDim.x) {
    const int e = (int)(w / nch), ch = (int)(w - (long long)e * nch);
    int lo = 0, hi = ntiles - 1;      // the tile whose segment holds entry e
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_pref[mid + 1] > e) hi = mid; else lo = mid + 1;
    }
    const int tile_id = lo;
    const int ecode = rwork[(size_t)tile_id * 2 * n_om + (e - s_pref[tile_id])];
    const int iom = ecode >> 1, sg = ecode & 1;
    const RelTile tl = tiles[tile_id];
    const SpeciesDev& sp = g.sp[tl.s];
    const int nabs = tl.nabs;
    const double qs = sp.qs, ms = sp.ms;
    const cd omc = mk(om[2 * iom], om[2 * iom + 1]);
    if (nabs != last_nabs) {      // series coefficients of the Landau term: 1/k!, 1/Gamma(|n| + i)
      __syncthreads();
      if (tid < 21) {
        double fact = 1.0;
        for (int k = 2; k <= tid; k++) fact = fact * (1.0 * k);
        s_rfact[tid] = 1.0 / fact;
      } else if (tid >= 32 && tid < 32 + 23) {
        const int m = nabs + (tid - 32);
        s_rgam[tid - 32] = m >= 1 ? 1.0 / gamma_ref(1.0 * m) : 0.0;
      }
      __syncthreads();
      last_nabs = nabs;
    }
    RelCtx c;
    c.g = gp;
    c.sp = &sp;
    c.om = omc;
    c.nabs = nabs;
    c.pref = -2.0 * PI_ * ((ms / vA) * (ms / vA) * (ms / vA)) * (qs * vA / (kpar * ms));
    c.zfac = kperp * ms / (vA * qs);
    c.zbar = g.kperp_norm ? kperp * ms / (vA * qs) : ms / (vA * qs);
    c.kf1 = g.kperp_norm ? 1.0 : kperp;
    c.kf2 = g.kperp_norm ? 1.0 : kperp * kperp;
    const double nn = sg ? -(double)nabs : (double)nabs;
    Six2 Sd, acc;
    zero6(Sd);
    const int row0 = 1 + ch * RR_CH, row1 = min(ng - 1, row0 + RR_CH - 1);
    for (int ig = row0 + warp; ig <= row1; ig += nwarps) rel_res_row(c, g, sp, omc, sg, nn, ig, lane, s_win[warp], Sd, err_flag);
    moments_to_modes(c, nn, Sd, acc);
#pragma unroll
    for (int q = 0; q < 6; q++) acc.v[q] = (sp.dgamma * 0.25) * acc.v[q];
    // landau_integrate_rel (Im om <= 0): one thread per row of the chunk
    if (omc.y <= 0.0 && tid < RR_CH && row0 + tid <= row1)
      landau_rows(c, g, sp, omc, sg, nn, row0 - 1 + tid, ng, s_rfact, s_rgam, acc);
#pragma unroll
    for (int q = 0; q < 6; q++) {
      cd v = warp_sum_cd(acc.v[q]);
      if (lane == 0) s_red[warp][q] = v;
    }
    __syncthreads();
    const size_t item = (size_t)iom * g.NI + sp.item_base + 2 * nabs + sg;
    if (tid < 6) {
      cd t = mk(0.0, 0.0);
      for (int ww = 0; ww < nwarps; ww++) t += s_red[ww][tid];
      double* o = Mpart + (item * nch + ch) * 12;
      o[2 * tid] = t.x;
      o[2 * tid + 1] = t.y;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&tickets[item], 1) == nch - 1);
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (tid < 12) {
        double t = 0.0;
        for (int j = 0; j < nch; j++) t += __ldcg(Mpart + (item * nch + j) * 12 + tid);
        Mrel[item * 12 + tid] = t;
      }
      if (tid == 0) tickets[item] = 0;   // ready for the next launch
    }
    __syncthreads();
  }
}
int rel_rows_chunks(int ngamma) { return (ngamma - 1 + RR_CH - 1) / RR_CH; }

// Bessel factors of int_T_rel (src/ALPS_fns_rel.f90:1297-1322) depend on (igamma, ipparbar) through
// pperpbar = sqrt(gamma^2 - 1 - pparbar^2) but not on omega: tabulated once per k with the same literal BESSJ.
// Planes 0..nmaxord hold J_n, plane nmaxord + 1 holds pperpbar.
__global__ void k_rel_bessel_table(const double* __restrict__ grel, const double* __restrict__ pbrel, int ng, int npb,
                                   double zfac, int nmaxord, double* __restrict__ Jrel) {
  const int ldr = npb + 1;
  const size_t plane = (size_t)(ng + 1) * ldr;
  const size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (o >= plane || n > nmaxord + 1) return;
  const int ig = (int)(o / ldr), ip = (int)(o % ldr);
  const double gam = grel[ig], pb = pbrel[ip];
  const double a = gam * gam - 1.0 - pb * pb;
  double v = 0.0;
  if (n == nmaxord + 1) {
    if (a >= 0.0) v = sqrt(a);   // plane nmaxord + 1: pperpbar itself
  } else if (a >= 0.0) {
    v = bessj_ref(n, zfac * sqrt(a));
  }
  Jrel[(size_t)n * plane + o] = v;
}

// int_ee_rel, src/ALPS_fns_rel.f90:1097-1215: one block
__global__ void k_int_ee_rel(const double* __restrict__ pbv, const double* __restrict__ dfp, const int* __restrict__ lo,
                             const int* __restrict__ up, int ng, int npb, double qs, double ms, double vA, double dgam,
                             double dpb, double* __restrict__ out) {
  double acc = 0.0;
  const int ldr = npb + 1;
  for (int ig = 1 + threadIdx.x; ig <= ng - 1; ig += blockDim.x) {
    const double wg = (ig == ng - 1) ? 1.0 : 2.0;
    for (int ip = lo[ig]; ip <= up[ig]; ip++) {
      // ends count once even when lower == upper? the reference adds both end terms: twice (lines 1131-1134)
      double w = (ip == lo[ig] ? 1.0 : 0.0) + (ip == up[ig] ? 1.0 : 0.0) + ((ip > lo[ig] && ip < up[ig]) ? 2.0 : 0.0);
      acc += wg * w * pbv[ip] * dfp[(size_t)ig * ldr + ip];
    }
  }
  __shared__ double sm[256];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double r = sm[0] * 2.0 * PI_ * qs / ms;
    out[0] = r * dgam * dpb * 0.25 * ((ms / vA) * (ms / vA) * (ms / vA));
  }
}

// small batches: resonance flags + entry lists (k_rel_plan), the non-resonant signs per tile (k_rel), the rows of the
// resonant entries over the whole GPU (k_rel_rows).  rcount must be zero when k_rel_plan starts: zero_rcount, or the
// chain's k_plan has cleared it (launch_plan).
void launch_rel_small(const GlobalDev* g, const double* om, int n_om, const RelTile* tiles, int ntiles, double* Mrel,
                      int* err_flag, unsigned char* rflag, int* rwork, int* rcount, int* rpos, int nsplit, double* Mpart,
                      int* tickets, size_t rows_part_offset, size_t rows_ticket_offset, int sm_count, bool zero_rcount,
                      cudaStream_t st) {
  if (n_om <= 0 || ntiles <= 0) return;
  if (zero_rcount) cudaMemsetAsync(rcount, 0, (size_t)ntiles * sizeof(int), st);
  launch_chain(k_rel_plan_blk, dim3(n_om * ntiles), dim3(256), 0, st, g, om, n_om, tiles, ntiles, rflag, rwork, rcount, rpos);
  if (nsplit < 1 || !Mpart || !tickets) nsplit = 1;
  // k_rel_rows and k_rel_nonres run side by side: separate partial rows and tickets (the caller sizes both for two users)
  const int sms = sm_count > 0 ? sm_count : 148;
  // one or two omegas: few work items, 255 registers and one CTA per SM; more: two CTAs of 128 registers per SM (the same
  // operations in the same order either way)
  static const char* rb = getenv("ALPS_B200_REL_ROWS_MINB");   // A/B knob: "1" / "2" forces the variant
  const bool two = rb ? rb[0] == '2' : n_om > 2;
  if (two)
    launch_chain(k_rel_rows<2>, dim3(2 * sms), dim3(REL_THREADS), 0, st, g, om, n_om, tiles, ntiles, Mrel, err_flag,
                 (const int*)rwork, (const int*)rcount, Mpart + rows_part_offset, tickets + rows_ticket_offset);
  else
    launch_chain(k_rel_rows<1>, dim3(sms), dim3(REL_THREADS), 0, st, g, om, n_om, tiles, ntiles, Mrel, err_flag,
                 (const int*)rwork, (const int*)rcount, Mpart + rows_part_offset, tickets + rows_ticket_offset);
  launch_chain(k_rel_nonres, dim3(n_om * ntiles * nsplit), dim3(REL_THREADS), 0, st, g, om, n_om, tiles, ntiles, Mrel, nsplit,
               Mpart, tickets, (const unsigned char*)rflag);
}
void launch_rel(const GlobalDev* g, const double* om, int n_om, const RelTile* tiles, int ntiles, double* Mrel,
                int* err_flag, int nsplit, double* Mpart, int* tickets, cudaStream_t st) {
  if (n_om <= 0 || ntiles <= 0) return;
  if (nsplit < 1 || !Mpart || !tickets) nsplit = 1;
  static const char* force = getenv("ALPS_B200_REL_MINB");   // A/B knob: "1" = 255-register variant always
  // one CTA of 255 registers per SM up to about two waves of them (measured: C3, one omega); beyond that two CTAs of 128 registers per SM
  // finish sooner (the same operations in the same order: the results do not depend on the choice)
  if ((long long)n_om * ntiles * nsplit <= 3 * 148 || (force && force[0] == '1'))
    launch_chain(k_rel<1>, dim3(n_om * ntiles * nsplit), dim3(REL_THREADS), 0, st, g, om, n_om, tiles, ntiles, Mrel,
                 err_flag, nsplit, Mpart, tickets, (const unsigned char*)nullptr, 0);
  else
    launch_chain(k_rel<2>, dim3(n_om * ntiles * nsplit), dim3(REL_THREADS), 0, st, g, om, n_om, tiles, ntiles, Mrel,
                 err_flag, nsplit, Mpart, tickets, (const unsigned char*)nullptr, 0);
}
// throughput class: resonance flags + work list, principal-value / Landau parts of the listed entries, omega-tiled rest
// (three launches; rflag: n_om * ntiles bytes, rwork: 2 * n_om * ntiles ints, rcount: one int)
void launch_rel_tiled(const GlobalDev* g, const double* om, int n_om, const RelTile* tiles, int ntiles, double* Mrel,
                      int* err_flag, unsigned char* rflag, int* rwork, int* rcount, int* rpos, double* dpart, int nsplitB,
                      int sm_count, cudaStream_t st) {
  if (n_om <= 0 || ntiles <= 0) return;
  cudaMemsetAsync(rcount, 0, (size_t)ntiles * sizeof(int), st);
  const int warps = n_om * ntiles;
  k_rel_plan<<<(warps + 7) / 8, 256, 0, st>>>(g, om, n_om, tiles, ntiles, rflag, rwork, rcount, rpos);
  static const char* pvb = getenv("ALPS_B200_REL_PV_MINB");   // A/B knob: "1" = 255 registers, one CTA per SM
  const int sms = sm_count > 0 ? sm_count : 148;
  if (pvb && pvb[0] == '1')
    k_rel_pv<1><<<sms, REL_THREADS, 0, st>>>(g, om, n_om, tiles, ntiles, Mrel, err_flag, rwork, rcount);
  else
    k_rel_pv<2><<<2 * sms, REL_THREADS, 0, st>>>(g, om, n_om, tiles, ntiles, Mrel, err_flag, rwork, rcount);
  k_rel_direct<<<dim3((2 * n_om + RT_THREADS - 1) / RT_THREADS, ntiles, nsplitB), RT_THREADS, 0, st>>>(
      g, om, n_om, tiles, ntiles, rwork, rcount, dpart, nsplitB, err_flag);
  k_rel_tiled<<<dim3((n_om + RT_THREADS - 1) / RT_THREADS, ntiles), RT_THREADS, 0, st>>>(g, om, n_om, tiles, ntiles, rflag,
                                                                                        Mrel, rpos, dpart, nsplitB);
}
void launch_rel_bessel_table(const double* grel, const double* pbrel, int ng, int npb, double zfac, int nmaxord,
                             double* Jrel, cudaStream_t st) {
  const size_t plane = (size_t)(ng + 1) * (npb + 1);
  dim3 grid((unsigned)((plane + 127) / 128), nmaxord + 2);   // + the pperpbar plane
  k_rel_bessel_table<<<grid, 128, 0, st>>>(grel, pbrel, ng, npb, zfac, nmaxord, Jrel);
}
void launch_int_ee_rel(const double* pbv, const double* dfp, const int* lo, const int* up, int ng, int npb, double qs,
                       double ms, double vA, double dgam, double dpb, double* out, cudaStream_t st) {
  k_int_ee_rel<<<1, 256, 0, st>>>(pbv, dfp, lo, up, ng, npb, qs, ms, vA, dgam, dpb, out);
}

}  // namespace alps
