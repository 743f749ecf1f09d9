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
    j0 = bessj_ref(c.nabs, z);
    jp = bessj_ref(c.nabs + 1, z);
    jm = c.nabs >= 1 ? bessj_ref(c.nabs - 1, z) : 0.0;
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
constexpr int REL_WIN = 32;
__device__ __forceinline__ void funct_g_win(const RelCtx& c, int sg, double p, int ig, const cd (*win)[6], int W0,
                                            int nwin, double inv_dpb, Six2& out) {
  const SpeciesDev& sp = *c.sp;
  const int ic = funct_g_node(sp, c.g->npparbar, ig, p, inv_dpb);
  const int k = ic - W0;
  if (k < 1 || k + 1 >= nwin) {
    funct_g_rel6(c, sg, p, ig, ic, inv_dpb, out);
    return;
  }
  const double sx = (0.5 * inv_dpb) * (p - sp.pbrel[ic]);   // central-difference slope factor times the offset
#pragma unroll
  for (int q = 0; q < 6; q++) {
    const cd a = win[k][q], lo = win[k - 1][q], hi = win[k + 1][q];
    out.v[q] = mk(fma(sx, hi.x - lo.x, a.x), fma(sx, hi.y - lo.y, a.y));
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

constexpr int REL_THREADS = 256;

// nsplit > 1 (few omegas in flight: sequential root finding is latency bound): the gamma rows / grid points
// of one (omega, species, |n|) are dealt round-robin to nsplit CTAs; each leaves a partial row in Mpart and
// the last one to finish (ticket counter) adds them up in a fixed order.
// MINB = 2 (throughput batches): 128 registers, two CTAs per SM; MINB = 1 (latency batches): 255 registers
template <int MINB>
__global__ void __launch_bounds__(REL_THREADS, MINB) k_rel(const GlobalDev* __restrict__ gp, const double* __restrict__ om,
                                                     int n_om, const RelTile* __restrict__ tiles, int ntiles,
                                                     double* __restrict__ Mrel, int* __restrict__ err_flag, int nsplit,
                                                     double* __restrict__ Mpart, int* __restrict__ tickets) {
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
  __shared__ cd s_win[REL_THREADS / 32][REL_WIN][6];   // node values of the principal-value window, per warp
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
  {
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
      for (int ig = 1 + warp + nwarps * js; ig <= ng - 1; ig += nwarps * nsplit) {
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
            continue;
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
        if (found && lowerlimit >= int_start && upperlimit <= int_end) {
          const double gres = sp.grel[ig];   // gamma_rel(sproc_rel,igamma,ipparbar_res): separable grid
          const double denomR = (gres * omc.x * vA / kpar) - (1.0 * nn) * (qs / ms) * vA / kpar;
          const double denomI = gres * omc.y * vA / kpar;
          const double capDelta = denomR - pbv[ires - M_I];
          const double smdelta = capDelta / (1.0 * M_P);
          Six2 pr;
          zero6(pr);
          // node values of the window [ires - M_I - 2, ires + M_I + 4]: one node per lane
          const int W0 = ires - M_I - 2, nwin = (2 * M_I + 7 <= REL_WIN) ? 2 * M_I + 7 : 0;
          const cd(*win)[6] = s_win[warp];
          const double inv_dpb = 1.0 / dpb;
          __syncwarp();
          if (lane < nwin && W0 + lane >= 0 && W0 + lane <= npb) {
            double M[6];
            cd num;
            node_moments(c, ig, W0 + lane, sg, M, num);
#pragma unroll
            for (int q = 0; q < 6; q++) s_win[warp][lane][q] = M[q] * num;
          }
          __syncwarp();
          if (fabs(denomI) > g.Tlim) {
            for (int j = lane; j <= M_P; j += 32) {
              const double wj = (j == 0 || j == M_P) ? 1.0 : 2.0;
              const double p = (j == 0) ? denomR : (j == M_P ? denomR + capDelta : denomR + smdelta * j);
              Six2 f1, f2;
              funct_g_win(c, sg, p, ig, win, W0, nwin, inv_dpb, f1);
              funct_g_win(c, sg, 2.0 * denomR - p, ig, win, W0, nwin, inv_dpb, f2);
              // wj / d1 and wj / d2 with d2 = conj(d1): one reciprocal for the twelve quotients
              const double dx = p - denomR, tt = wj * fast_rcp(fma(dx, dx, denomI * denomI));
              const cd r1 = mk(dx * tt, denomI * tt), r2 = mk(dx * tt, -(denomI * tt));
#pragma unroll
              for (int q = 0; q < 6; q++) pr.v[q] += f1.v[q] * r1 - f2.v[q] * r2;
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
      }
      // moment sums of the direct and principal parts -> tensor components
      moments_to_modes(c, nn, Sd, acc);
#pragma unroll
      for (int q = 0; q < 6; q++) acc.v[q] = (dgam * 0.25) * acc.v[q];

      // ---- landau_integrate_rel (Im om <= 0), all threads stride over igamma
      if (omc.y <= 0.0) {
        Six2 L;
        zero6(L);
        for (int ig = 1 + tid + REL_THREADS * js; ig <= ng - 1; ig += REL_THREADS * nsplit) {
          const double g1 = sp.grel[ig];
          const cd pres = (g1 * omc) * vA / kpar - mk((1.0 * nn) * qs * vA / (kpar * ms), 0.0);
          if (!(pres.x * pres.x <= g1 * g1 - 1.0)) continue;
          const double h = (ig == ng - 1) ? 0.5 : 1.0;
          cd dfg;
          if (ig == 1) dfg = (eval_fit_rel(g, sp, ig + 1, pres) - eval_fit_rel(g, sp, ig, pres)) / dgam;
          else dfg = (eval_fit_rel(g, sp, ig + 1, pres) - eval_fit_rel(g, sp, ig - 1, pres)) / (2.0 * dgam);
          const cd dfp = (eval_fit_rel(g, sp, ig, pres + mk(dpb, 0.0)) - eval_fit_rel(g, sp, ig, pres - mk(dpb, 0.0))) /
                         (2.0 * dpb);
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
        if (!(nabs == 0 && sg == 1)) {
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

void launch_rel(const GlobalDev* g, const double* om, int n_om, const RelTile* tiles, int ntiles, double* Mrel,
                int* err_flag, int nsplit, double* Mpart, int* tickets, cudaStream_t st) {
  if (n_om <= 0 || ntiles <= 0) return;
  if (nsplit < 1 || !Mpart || !tickets) nsplit = 1;
  static const char* force = getenv("ALPS_B200_REL_MINB");   // A/B knob: "1" = 255-register variant always
  if (nsplit > 1 || n_om * ntiles <= 2 * 148 || (force && force[0] == '1'))
    launch_chain(k_rel<1>, dim3(n_om * ntiles * nsplit), dim3(REL_THREADS), 0, st, g, om, n_om, tiles, ntiles, Mrel,
                 err_flag, nsplit, Mpart, tickets);
  else
    launch_chain(k_rel<2>, dim3(n_om * ntiles * nsplit), dim3(REL_THREADS), 0, st, g, om, n_om, tiles, ntiles, Mrel,
                 err_flag, nsplit, Mpart, tickets);
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
