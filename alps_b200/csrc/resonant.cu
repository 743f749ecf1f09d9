// alps_b200: resonance handling around the regular quadrature, compiled with -fmad=false so the
// index / interval logic sees the same unfused arithmetic as the reference.
//
//   k_plan       determine_resonances (src/ALPS_fns.f90:641-745) and the interval logic of
//                integrate_res (src/ALPS_fns.f90:941-1006) for every (omega, species, n, sign)
//   k_resonant   near-pole quadrature + tiny rest of integrate_res (:1008-1233, with funct_g
//                :1243-1321) and the Landau residue term landau_integrate (:1327-1452) with
//                eval_fit (src/ALPS_analyt.f90:32-363); one warp per resonant harmonic
//   k_chi_partial  the harmonic sums of disp() (:363-514): tensor components from the moment
//                sums, chi_low for n = 0, +-1, the ee term, the ns*qs normalisation
//   k_assemble   the rank-0 part of disp() (:536-624): chi0, eps, wave, determinant
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace alps {

constexpr double PI = 3.14159265358979323846;

// ------------------------------------------------------------ complex elementary functions
__device__ inline cd c_exp(cd z) {
  double e = exp(z.x), s, c;
  sincos(z.y, &s, &c);
  return mk(e * c, e * s);
}
__device__ inline cd c_log(cd z) { return mk(log(hypot(z.x, z.y)), atan2(z.y, z.x)); }
__device__ inline cd c_pow_real(cd z, double p) {   // z**p as exp(p log z) (libm cpow)
  if (z.x == 0.0 && z.y == 0.0) return mk(p == 0.0 ? 1.0 : 0.0, 0.0);
  cd l = c_log(z);
  return c_exp(mk(p * l.x, p * l.y));
}
__device__ inline cd c_sqrt(cd z) {
  double m = hypot(z.x, z.y);
  if (m == 0.0) return mk(0.0, 0.0);
  double a = sqrt(0.5 * (m + fabs(z.x)));
  double b = 0.5 * z.y / a;
  return z.x >= 0.0 ? mk(a, b) : mk(fabs(b), copysign(a, z.y));
}

// eval_fit, src/ALPS_analyt.f90:32-258 (+ distribution_analyt, + fit_function_poly :262-363)
__device__ cd eval_fit(const GlobalDev& g, int s, int iperp, cd p) {
  const SpeciesDev& sp = g.sp[s];
  const double pperp = sp.pperp[iperp];
  if (sp.ACmethod == 0) {
    // distribution/distribution_analyt.f90:66-85: hard-coded beta=1 Maxwellians for species 1, 2
    double beta = 1.0, ms;
    if (s == 0) ms = 1.0;
    else if (s == 1) ms = 1.0 / 1836.0;
    else return mk(0.0, 0.0);
    cd e = -((p * p) / (beta * ms) + mk((pperp * pperp) / (beta * ms), 0.0));
    return (pow(PI, -1.5) / pow(ms * beta, 3.0 / 2.0)) * c_exp(e);
  }
  if (sp.ACmethod == 2) {
    if (sp.poly_kind != 1) return mk(0.0, 0.0);
    const double* co = sp.poly + (size_t)iperp * (g.maxorder + 1);
    double norm_1 = 5.e-1 * (sp.ppar[g.npar] + sp.ppar[0]);
    double norm_2 = 5.e-1 * (sp.ppar[g.npar] - sp.ppar[0]);
    cd t = (p - mk(norm_1, 0.0)) / norm_2;
    if (cabs2(t) > 1.0) return mk(0.0, 0.0);
    cd b0 = mk(1.0, 0.0), b1 = t, r = co[0] * b0;
    if (sp.poly_order >= 1) r += co[1] * b1;
    for (int n = 2; n <= sp.poly_order; n++) {
      cd b2 = mk(2.0, 0.0) * t * b1 - b0;
      r += co[n] * b2;
      b0 = b1;
      b1 = b2;
    }
    if (sp.logfit) {
      double lm = sp.poly_log_max;
      if (r.x < -lm || r.y < -lm || r.x > lm || r.y > lm) return mk(0.0, 0.0);
      const double ln10 = 2.302585092994045684;
      return c_exp(mk(ln10 * r.x, ln10 * r.y));
    }
    return r;
  }
  cd f = mk(0.0, 0.0);
  for (int ifit = 0; ifit < sp.n_fits; ifit++) {
    const double* pf = sp.param_fit + ((size_t)iperp * g.maxfits + ifit) * 5;
    const double p1 = pf[0], p2 = pf[1], p3 = pf[2], p4 = pf[3], p5 = pf[4];
    const double pc = sp.perp_correction[ifit];
    cd d = p - mk(p3, 0.0);
    cd d2 = d * d;
    switch (sp.fit_type[ifit]) {
      case 1:
        f += (p1 * exp(-pc * pperp * pperp)) * c_exp(-(p2 * d2));
        break;
      case 2: {
        cd kp = mk(1.0, 0.0) + p2 * d2 + mk(pc * p5 * pperp * pperp, 0.0);
        f += p1 * c_pow_real(kp, p4);
        break;
      }
      case 3: {
        cd sq = c_sqrt(mk(1.0, 0.0) + (mk(pperp * pperp, 0.0) + d2) * (g.vA * g.vA / (sp.ms * sp.ms)));
        f += p1 * c_exp(-(p2 * sq));
        break;
      }
      case 6: {
        cd e = mk(p4 * pc * pperp * pperp, 0.0) + p2 * d2;
        f += p1 * c_exp(0.5 * (e - c_exp(e)));
        break;
      }
      default: break;
    }
  }
  return f;
}

// ------------------------------------------------------------------------ plan
__device__ __forceinline__ void decode_item(const GlobalDev& g, int it, int& s, int& nabs, int& sg) {
  s = 0;
  for (int q = 1; q < g.nspec; q++)
    if (it >= g.sp[q].item_base) s = q;
  int r = it - g.sp[s].item_base;
  nabs = r >> 1;
  sg = r & 1;
}

// The resonance plan of item idx = (omega, species, |n|, sign): determine_resonances and the index ranges of
// integrate / integrate_res (src/ALPS_fns.f90:641-745, 941-1006).
__device__ __forceinline__ void plan_item(const GlobalDev& g, const double* om, size_t idx, PlanEntry* __restrict__ plan,
                                          int* __restrict__ work, int* __restrict__ work_count) {
  int iom = (int)(idx / g.NI), it = (int)(idx % g.NI), s, nabs, sg;
  decode_item(g, it, s, nabs, sg);
  const SpeciesDev& sp = g.sp[s];
  PlanEntry pe;
  pe.lo1 = 1; pe.hi1 = 0; pe.lo2 = 1; pe.hi2 = 0; pe.flags = 0; pe.ipar_res = 0; pe.upperlimit = 0; pe.pad = 0;
  if ((nabs == 0 && sg == 1) || nabs < sp.nlo_shard || nabs > sp.nhi_shard || sp.usebM) {
    plan[idx] = pe;
    return;
  }
  if (sp.relativistic) {   // resonance handling of relativistic species lives in k_rel
    pe.flags = PLAN_ACTIVE | PLAN_REL;
    plan[idx] = pe;
    return;
  }
  pe.flags = PLAN_ACTIVE;
  const int npar = g.npar, M_I = g.M_I;
  const double* ppar = sp.ppar;
  const double omr = om[2 * iom], omi = om[2 * iom + 1];
  const int nn = sg ? -nabs : nabs;
  const double pr = (sp.ms * omr - 1.0 * nn * sp.qs) / g.kpar;
  const double dppar = sp.dppar_signed;
  // determine_resonances: pp(ipar) <= Re p_res < pp(ipar+1) for some ipar in [1,npar-1]
  bool found = (pr >= ppar[1]) && (pr < ppar[npar]);
  if ((pr < ppar[1]) && (pr >= ppar[1] - (1.0 * M_I) * dppar)) found = true;
  if ((pr >= ppar[npar - 1]) && (pr < ppar[npar - 1] + (1.0 * M_I) * dppar)) found = true;
  if (!found) {
    pe.lo1 = 1;
    pe.hi1 = npar - 1;
    plan[idx] = pe;
    return;
  }
  pe.flags |= PLAN_RES;
  if (omi <= 0.0) pe.flags |= PLAN_LANDAU;
  // integrate_res: the step left of the resonance, ipar in [1,npar-2]
  int ipar_res = 0;
  if (pr >= ppar[1] && pr < ppar[npar - 1]) {
    // largest ipar in [1, npar-2] with ppar[ipar] <= pr: index guess on the (uniform) grid, corrected with the
    // actual node values -- the same index a search over the nodes finds, without its chain of dependent loads
    int lo = (int)floor((pr - ppar[1]) / dppar) + 1;
    lo = min(max(lo, 1), npar - 2);
    while (lo > 1 && ppar[lo] > pr) lo--;
    while (lo < npar - 2 && ppar[lo + 1] <= pr) lo++;
    if (ppar[lo + 1] > pr && ppar[lo] <= pr) ipar_res = lo;
  }
  for (int ipar = 0; ipar <= M_I; ipar++) {
    if ((pr >= (ppar[0] - dppar * ipar)) && (pr < (ppar[0] - dppar * (ipar - 1)))) ipar_res = -ipar;
    if ((pr >= (ppar[npar - 1] + dppar * ipar)) && (pr < (ppar[npar - 1] + dppar * (ipar + 1))))
      ipar_res = npar - 1 + ipar;
  }
  pe.ipar_res = ipar_res;
  if (ipar_res - M_I <= 2) {
    pe.lo1 = max(ipar_res + M_I, 1);
    pe.hi1 = npar - 1;
  } else if (ipar_res + M_I >= npar - 2) {
    pe.lo1 = 1;
    pe.hi1 = min(ipar_res - M_I, npar - 1);
  } else {
    pe.lo1 = 1;
    pe.hi1 = ipar_res - M_I;
    pe.upperlimit = (fabs(pr - ppar[ipar_res]) < 0.5 * dppar) ? ipar_res + M_I + 1 : ipar_res + M_I + 2;
    pe.lo2 = pe.upperlimit;
    pe.hi2 = npar - 1;
    pe.flags |= PLAN_NEAR;
  }
  plan[idx] = pe;
  if (pe.flags & (PLAN_NEAR | PLAN_LANDAU)) work[atomicAdd(work_count, 1)] = (int)idx;
}


// FUSED (single-block launches of the single-omega graph, api.cu): the block resets the work counter itself and
// stages the omegas -- read from pinned host memory -- into om_stage for the kernels downstream, so the chain
// needs neither a memset nor a host-to-device copy node.  plan_flag: 0 while the plan of this launch is being written,
// 1 once it is complete -- the Landau blocks of k_resonant_lat, launched programmatically while k_quad_mma still runs,
// wait for it instead of for the completion of their predecessor (the flag is cleared before the dependents are allowed
// to start, so they can never see the previous launch's value).
template <bool FUSED>
__global__ void __launch_bounds__(FUSED ? 1024 : 256)
k_plan(const GlobalDev* __restrict__ gp, const double* __restrict__ om_in, int n_om, PlanEntry* __restrict__ plan,
       int* __restrict__ work, int* __restrict__ work_count, double* __restrict__ om_stage, int* plan_flag,
       int* __restrict__ zero_ints, int nzero) {
  const GlobalDev& g = *gp;
  __shared__ double s_om[FUSED ? 2 * PLAN_FUSED_MAX_OM : 2];
  const double* om = om_in;
  if (FUSED) {
    for (int t = threadIdx.x; t < nzero; t += blockDim.x) zero_ints[t] = 0;   // (counts of k_rel_plan, relativistic species)
    if (threadIdx.x == 0) {
      lat_stamp(g, 0);
      *reinterpret_cast<volatile int*>(plan_flag) = 0;
      *reinterpret_cast<volatile int*>(plan_flag + CHAIN_QUAD) = 0;   // CTAs of k_quad_mma that have written their sums
      *reinterpret_cast<volatile int*>(plan_flag + CHAIN_RES) = 0;    // blocks of k_resonant_lat that have finished
    }
    __threadfence();
    __syncthreads();
    pdl_trigger();
    if (threadIdx.x < 2 * n_om) {
      const double v = om_in[threadIdx.x];
      s_om[threadIdx.x] = v;
      om_stage[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *work_count = 0;
    __syncthreads();
    if (threadIdx.x == 0) lat_stamp(g, 21);
    om = s_om;
  }
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx < (size_t)n_om * g.NI) plan_item(g, om, idx, plan, work, work_count);
  if (FUSED) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      *reinterpret_cast<volatile int*>(plan_flag) = 1;
      lat_stamp(g, 1);
    }
  }
}

// ------------------------------------------------------------------- resonant
struct Six {
  cd v[6];   // combos (a,0) (a,1) (a,2) (b,0) (b,1) (c,0)
};
__device__ __forceinline__ void six_zero(Six& s) {
#pragma unroll
  for (int q = 0; q < 6; q++) s.v[q] = mk(0.0, 0.0);
}

// Sum over iperp (with the p_perp trapezoid weights) of funct_g for the six (weight, moment)
// combinations at real p; linear interpolation around the nearest grid node exactly as
// funct_g does per iperp (src/ALPS_fns.f90:1284-1319).  Two steps: the node selection (grid coordinates only) and the
// interpolation of the window sums gw of k_quad -- k_resonant_lat does the first before it waits for k_quad.
struct GLoc {
  int jm, j0, jp;        // window slots of nodes ic - 1, ic, ic + 1
  int err;
  double x, pm, p0, pp;  // p - ppar(ic), ppar(ic - 1), ppar(ic), ppar(ic + 1)
};
__device__ __forceinline__ GLoc funct_g6_locate(const GlobalDev& g, const SpeciesDev& sp, int wbase, double p) {
  const int npar = g.npar;
  const double* ppar = sp.ppar;
  const double dp = sp.dppar_abs;
  int i0 = (int)floor((p - ppar[0]) / dp + 0.5);
  int ic = 0;
  {
    // nearest node among i0 + 2, ..., i0 - 2 (within [1, npar - 1]), first match in that order; the five candidates are
    // fetched together (one round trip instead of up to five dependent ones)
    double cand[5];
#pragma unroll
    for (int d = 0; d < 5; d++) cand[d] = ppar[min(max(i0 + 2 - d, 0), npar)];
#pragma unroll
    for (int d = 4; d >= 0; d--) {
      const int c = i0 + 2 - d;
      if (c <= npar - 1 && c >= 1 && fabs(cand[d] - p) <= 0.5 * dp) ic = c;
    }
  }
  if (ic >= npar - 1) ic = npar - 2;
  if (ic <= 1) ic = 2;
  GLoc L;
  L.err = 0;
  // window slots of nodes ic-1, ic, ic+1 (nodes 1..3 live behind the window, see GlobalDev::WINX)
  L.jm = ic - 1 - wbase;
  L.j0 = ic - wbase;
  L.jp = ic + 1 - wbase;
  if (L.jm < 0 || L.jp >= g.WIN) {
    if (ic == 2) {
      L.jm = g.WIN;
      L.j0 = g.WIN + 1;
      L.jp = g.WIN + 2;
    } else {
      L.err = 1 + (ic & 0xffff) + ((i0 & 0x7fff) << 16);
      L.jm = L.j0 = L.jp = 0;
    }
  }
  L.pm = ppar[ic - 1];
  L.p0 = ppar[ic];
  L.pp = ppar[ic + 1];
  L.x = p - L.p0;
  return L;
}
// (dp = |dp_par| of the species, kpar: read by the caller -- k_resonant_lat reads them before its wait)
__device__ __forceinline__ void funct_g6_at(const double dp, const double kpar, const double* __restrict__ gw,
                                            const GLoc& L, Six& out, int* err, int q0 = 0, int nq = 6) {
  if (L.err) {
    *err = L.err;
    six_zero(out);
    return;
  }
#pragma unroll
  for (int q = 0; q < 6; q++) {
    if (q < q0 || q >= q0 + nq) {   // latency variant: the six combinations are dealt to different blocks
      out.v[q] = mk(0.0, 0.0);
      continue;
    }
    const int xt = (q < 3) ? 0 : (q < 5 ? 1 : 2);   // weight type a,b,c
    const int m = (q < 3) ? q : (q < 5 ? q - 3 : 0);  // p_par power
    cd gm, g0, gp;
    {
      const double* w = gw + ((size_t)L.jm * 3 + xt) * 2;
      double pw = m == 0 ? 1.0 : (m == 1 ? L.pm : L.pm * L.pm);
      gm = -(mk(w[0], w[1]) * pw) / kpar;
    }
    {
      const double* w = gw + ((size_t)L.j0 * 3 + xt) * 2;
      double pw = m == 0 ? 1.0 : (m == 1 ? L.p0 : L.p0 * L.p0);
      g0 = -(mk(w[0], w[1]) * pw) / kpar;
    }
    {
      const double* w = gw + ((size_t)L.jp * 3 + xt) * 2;
      double pw = m == 0 ? 1.0 : (m == 1 ? L.pp : L.pp * L.pp);
      gp = -(mk(w[0], w[1]) * pw) / kpar;
    }
    out.v[q] = g0 + (0.5 * ((gp - gm) / dp)) * L.x;
  }
}
// The two combinations q0, q0 + 1 of a k_resonant_lat block (same arithmetic as funct_g6_at, a third of its code)
__device__ __forceinline__ void funct_g2_at(const double dp, const double kpar, const double* __restrict__ gw,
                                            const GLoc& L, int q0, cd out[2], int* err) {
  if (L.err) {
    *err = L.err;
    out[0] = out[1] = mk(0.0, 0.0);
    return;
  }
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int q = q0 + e;
    const int xt = (q < 3) ? 0 : (q < 5 ? 1 : 2);   // weight type a,b,c
    const int m = (q < 3) ? q : (q < 5 ? q - 3 : 0);  // p_par power
    const double* wm = gw + ((size_t)L.jm * 3 + xt) * 2;
    const double* w0 = gw + ((size_t)L.j0 * 3 + xt) * 2;
    const double* wp = gw + ((size_t)L.jp * 3 + xt) * 2;
    const cd vm = mk(wm[0], wm[1]), v0 = mk(w0[0], w0[1]), vp = mk(wp[0], wp[1]);
    const double pwm = m == 0 ? 1.0 : (m == 1 ? L.pm : L.pm * L.pm);
    const double pw0 = m == 0 ? 1.0 : (m == 1 ? L.p0 : L.p0 * L.p0);
    const double pwp = m == 0 ? 1.0 : (m == 1 ? L.pp : L.pp * L.pp);
    const cd gm = -(vm * pwm) / kpar, g0 = -(v0 * pw0) / kpar, gp = -(vp * pwp) / kpar;
    out[e] = g0 + (0.5 * ((gp - gm) / dp)) * L.x;
  }
}
__device__ inline void funct_g6(const GlobalDev& g, const SpeciesDev& sp, const double* __restrict__ gw, int wbase,
                                double p, Six& out, int* err, int q0 = 0, int nq = 6) {
  const GLoc L = funct_g6_locate(g, sp, wbase, p);
  funct_g6_at(sp.dppar_abs, g.kpar, gw, L, out, err, q0, nq);
}

__device__ __forceinline__ cd warp_sum_c(cd v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  return v;
}

constexpr int RES_THREADS = 128;
// sum over the 32 lanes cooperating on one harmonic: every lane gets the total, deterministic order
__device__ __forceinline__ void warp_sum6(Six& v) {
#pragma unroll
  for (int q = 0; q < 6; q++) v.v[q] = warp_sum_c(v.v[q]);
}

// One warp per resonant harmonic (batches > 64: throughput matters; k_resonant_lat below serves small ones).
__global__ void __launch_bounds__(RES_THREADS) k_resonant(const GlobalDev* __restrict__ gp, const double* __restrict__ om,
                                                  const PlanEntry* __restrict__ plan, const int* __restrict__ work,
                                                  const int* __restrict__ work_count,
                                                  const double* __restrict__ gwin, double* __restrict__ Sres,
                                                  int* __restrict__ err_flag) {
  const GlobalDev& g = *gp;
  constexpr int STRIDE = 32;      // threads cooperating on one harmonic
  const int wlane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int lane = wlane;         // index inside the cooperative loops
  const int nwork = *work_count;
  const int first = blockIdx.x * (RES_THREADS / 32) + wid;
  const int step = gridDim.x * (RES_THREADS / 32);
  for (int wi = first; wi < nwork; wi += step) {
    const size_t idx = (size_t)work[wi];
    const int iom = (int)(idx / g.NI), it = (int)(idx % g.NI);
    int s, nabs, sg;
    decode_item(g, it, s, nabs, sg);
    const SpeciesDev& sp = g.sp[s];
    const PlanEntry pe = plan[idx];
    const int nperp = g.nperp, M_I = g.M_I, M_P = g.M_P;
    const double* ppar = sp.ppar;
    const cd omc = mk(om[2 * iom], om[2 * iom + 1]);
    const int nn = sg ? -nabs : nabs;
    const double qs = sp.qs, ms = sp.ms, kpar = g.kpar;
    const double pR = (ms * omc.x - 1.0 * nn * qs) / kpar;
    const double pI = (ms * omc.y) / kpar;
    const cd p_res = mk(pR, pI);
    Six tot;
    six_zero(tot);
    int err = 0;

    if (pe.flags & PLAN_NEAR) {
      const double* gw = gwin + idx * (size_t)g.WINX * 6;
      const int wbase = pe.ipar_res - M_I - 2;
      const double dppar = sp.dppar_signed;
      const double capDelta = pR - ppar[pe.ipar_res - M_I];
      const double smdelta = capDelta / (1.0 * M_P);
      Six acc;
      six_zero(acc);
      if (fabs(pI) > g.Tlim) {
        // Eq. (3.5): symmetric pairing around the pole, src/ALPS_fns.f90:1026-1082
        for (int j = lane; j <= M_P; j += STRIDE) {
          const double wj = (j == 0 || j == M_P) ? 1.0 : 2.0;
          const double p = (j == 0) ? pR : (j == M_P ? pR + capDelta : pR + smdelta * j);
          Six f1, f2;
          funct_g6(g, sp, gw, wbase, p, f1, &err);
          funct_g6(g, sp, gw, wbase, 2.0 * pR - p, f2, &err);
          const cd d1 = mk(p - pR, -pI), d2 = mk(p - pR, pI);
#pragma unroll
          for (int q = 0; q < 6; q++) acc.v[q] += wj * (f1.v[q] / d1) - wj * (f2.v[q] / d2);
        }
      } else {
        // Eq. (3.6): linearised integrand + pole term, src/ALPS_fns.f90:1088-1165
        Six fp, fm, f0;
        funct_g6(g, sp, gw, wbase, pR + dppar, fp, &err);
        funct_g6(g, sp, gw, wbase, pR - dppar, fm, &err);
        Six gprime;
#pragma unroll
        for (int q = 0; q < 6; q++) gprime.v[q] = (fp.v[q] - fm.v[q]) / (2.0 * dppar);
        for (int j = 1 + lane; j <= M_P; j += STRIDE) {
          const double wj = (j == M_P) ? 1.0 : 2.0;
          const double p = (j == M_P) ? pR + capDelta : pR + smdelta * j;
          const double x2 = (p - pR) * (p - pR);
          const double lor = x2 / (x2 + pI * pI);
#pragma unroll
          for (int q = 0; q < 6; q++) acc.v[q] += (wj * 2.0) * gprime.v[q] * lor;
        }
        if (lane == 0 && pI != 0.0) {
          funct_g6(g, sp, gw, wbase, pR, f0, &err);
          const double sgn = pI > 0.0 ? 1.0 : -1.0;
#pragma unroll
          for (int q = 0; q < 6; q++) acc.v[q] += sgn * (cmul_i((2.0 * PI) * f0.v[q]) / smdelta);
        }
      }
      // tiny rest between p_res + capDelta and the first regular node, :1168-1230
      const double rest = ppar[pe.upperlimit] - pR - capDelta;
      const int ntiny = (int)(rest / smdelta);
      if (ntiny > 0) {
        const double correction = (rest / (1.0 * ntiny)) / smdelta;
        for (int j = lane; j <= ntiny; j += STRIDE) {
          const double wj = (j == 0 || j == ntiny) ? 1.0 : 2.0;
          const double p = (j == 0) ? pR + capDelta : pR + capDelta + correction * smdelta * j;
          Six f1;
          funct_g6(g, sp, gw, wbase, p, f1, &err);
          const cd d1 = mk(p - pR, -pI);
#pragma unroll
          for (int q = 0; q < 6; q++) acc.v[q] += (wj * correction) * (f1.v[q] / d1);
        }
      }
      const double fac = 2.0 * PI * smdelta * sp.dpperp * 0.25;
      warp_sum6(acc);
#pragma unroll
      for (int q = 0; q < 6; q++) tot.v[q] += fac * acc.v[q];
    }

    if (pe.flags & PLAN_LANDAU) {
      // landau_integrate, src/ALPS_fns.f90:1327-1452
      const double dpperp = sp.dpperp, dppar = sp.dppar_abs;
      const double* Jn = sp.J + (size_t)(nabs + 1) * sp.ldj;
      const double* Jm = sp.J + (size_t)nabs * sp.ldj;
      const double* Jp = sp.J + (size_t)(nabs + 2) * sp.ldj;
      cd La = mk(0.0, 0.0), Lb = La, Lc = La;
      int zero = 0;
      const cd ppl = mk(pR + dppar, pI), pmi = mk(pR - dppar, pI);
      for (int iperp = 1 + lane; iperp <= nperp - 1; iperp += STRIDE) {
        const double h = (iperp == nperp - 1) ? 0.5 : 1.0;
        const cd fpar_i = eval_fit(g, s, iperp, ppl);
        const cd fpar_f = eval_fit(g, s, iperp, pmi);
        const cd fperp_i = eval_fit(g, s, iperp + 1, p_res);
        const cd fperp_f = eval_fit(g, s, iperp - 1, p_res);
        // the reference tests fpar_f twice and never fperp_f (lines 1404-1405)
        if ((fpar_i.x == 0.0 && fpar_i.y == 0.0) || (fpar_f.x == 0.0 && fpar_f.y == 0.0) ||
            (fperp_i.x == 0.0 && fperp_i.y == 0.0))
          zero = 1;
        const cd dfperp = (fperp_i - fperp_f) / (2.0 * dpperp);
        const cd dfpar = (fpar_i - fpar_f) / (2.0 * dppar);
        const double pperp = sp.pperp[iperp];
        const cd Q = (qs / fabs(kpar)) * (((pperp * dfpar - p_res * dfperp) * kpar) / ms + omc * dfperp);
        const double bj = Jn[iperp];
        const double bp = (nabs >= 1) ? 0.5 * (Jm[iperp] - Jp[iperp]) : -Jp[iperp];
        La += (h * (bj * bj)) * Q;
        Lb += (h * (pperp * (bj * bp))) * Q;
        Lc += (h * ((pperp * pperp) * (bp * bp))) * Q;
      }
      if (lane == 0) {
        for (int e = 0; e < 2; e++) {
          const int iperp = e == 0 ? 0 : nperp;
          cd dfperp;
          if (e == 0)
            dfperp = (eval_fit(g, s, 1, p_res) - eval_fit(g, s, 0, p_res)) / dpperp;
          else
            dfperp = (eval_fit(g, s, nperp, p_res) - eval_fit(g, s, nperp - 1, p_res)) / dpperp;
          const cd dfpar = (eval_fit(g, s, iperp, ppl) - eval_fit(g, s, iperp, pmi)) / (2.0 * dppar);
          const double pperp = sp.pperp[iperp];
          const cd Q = (qs / fabs(kpar)) * (((pperp * dfpar - p_res * dfperp) * kpar) / ms + omc * dfperp);
          const double bj = Jn[iperp];
          const double bp = (nabs >= 1) ? 0.5 * (Jm[iperp] - Jp[iperp]) : -Jp[iperp];
          La += (0.5 * (bj * bj)) * Q;
          Lb += (0.5 * (pperp * (bj * bp))) * Q;
          Lc += (0.5 * ((pperp * pperp) * (bp * bp))) * Q;
        }
      }
      zero = __any_sync(0xffffffffu, zero);
      {
        Six L3;
        L3.v[0] = La; L3.v[1] = Lb; L3.v[2] = Lc;
        L3.v[3] = L3.v[4] = L3.v[5] = mk(0.0, 0.0);
        warp_sum6(L3);
        La = L3.v[0]; Lb = L3.v[1]; Lc = L3.v[2];
      }
      if (!zero) {
        // landau = -(sum) * i * dpperp * pi * 2 pi ; factor 2 (Im om < 0) or 1 (Im om == 0),
        // full_integrate src/ALPS_fns.f90:782-789
        const double mult = (omc.y < 0.0 ? 2.0 : 1.0) * dpperp * PI * 2.0 * PI;
        const cd ca = cmul_i(-La) * mult, cb = cmul_i(-Lb) * mult, cc = cmul_i(-Lc) * mult;
        tot.v[0] += ca;
        tot.v[1] += p_res * ca;
        tot.v[2] += (p_res * p_res) * ca;
        tot.v[3] += cb;
        tot.v[4] += p_res * cb;
        tot.v[5] += cc;
      }
    }
    for (int o = 16; o > 0; o >>= 1) err = max(err, __shfl_xor_sync(0xffffffffu, err, o));
    if (wlane == 0 && err) err_flag[0] = 1;
    if (lane == 0) {
      double* o = Sres + idx * 12;
#pragma unroll
      for (int q = 0; q < 6; q++) {
        o[2 * q] = tot.v[q].x;
        o[2 * q + 1] = tot.v[q].y;
      }
      if (err) {
        err_flag[0] = 1;
        err_flag[1] = (int)idx;
        err_flag[2] = pe.ipar_res;
        err_flag[3] = pe.upperlimit;
        err_flag[4] = pe.flags;
        err_flag[5] = err;
      }
    }
  }
}

// Latency variant (few omegas in flight: sequential root finding).  The independent pieces of one
// resonant harmonic go to LAT_PARTS different 256-thread blocks (blockIdx.y), i.e. to different SMs;
// each near-pole piece is further split by (weight, moment) combination, two per block:
//   parts 0..2  near-pole terms g(p)/(p - p_res)        (or the whole analytic branch)
//   parts 3..5  mirrored terms  g(2 p_R - p)/(p - p_res*)
//   parts 6..8  tiny rest
//   parts 9,10  Landau residue, alternating groups of 64 p_perp rows: 4 threads per row, one eval_fit
//               each; the two end rows (iperp = 0, nperp) are ordinary rows of the same scheme
// Each block leaves 6 complex partial sums (+ zero flag, error code, the factor of the near-pole terms) in its row of
// Spart; the harmonic sums (chi_partial_block -> lat_parts_combine) add the rows of an item in a fixed order.  Same
// arithmetic per term as k_resonant; only the (fixed) summation order differs.
// plan_flag (single-omega chain with programmatic launches): the Landau blocks need the plan but not the window sums of
// k_quad_mma, so they wait for k_plan's completion flag instead of for their predecessor and run beside k_quad_mma.
constexpr int LAT_THREADS = 256;
constexpr int LAT_PARTS = 11;
constexpr int LAT_STRIDE = 16;   // doubles per partial row: 12 sums, zero flag, error code, near-pole factor
static_assert(LAT_PARTS * LAT_STRIDE == RES_PART_DOUBLES, "partial-row buffer size");
__global__ void __launch_bounds__(LAT_THREADS) k_resonant_lat(const GlobalDev* __restrict__ gp,
                                                              const double* __restrict__ om,
                                                              const PlanEntry* __restrict__ plan,
                                                              const int* __restrict__ work,
                                                              const int* __restrict__ work_count,
                                                              const double* __restrict__ gwin,
                                                              int* __restrict__ err_flag, double* __restrict__ Spart,
                                                              int* chain, int nquad) {
  const GlobalDev& g = *gp;
  const int tid = threadIdx.x, wlane = tid & 31, wid = tid >> 5, part = blockIdx.y;
  __shared__ cd s_part[LAT_THREADS / 32][3];
  __shared__ cd s_f[3][2];   // analytic branch: g(p_R + dp), g(p_R - dp), g(p_R) for this block's two combinations
  __shared__ int s_err, s_seen;
  pdl_trigger();
  if (tid == 0) lat_stamp(g, 8);
  // quad_waited: this block has waited for its predecessor (k_quad_mma: the window sums gwin).  With plan_flag every
  // block starts on k_plan's completion flag instead, does everything that needs only the plan and the grids -- work
  // list, plan entry, species constants, the node selection of its first quadrature point; the Landau blocks all of
  // their work -- beside k_quad_mma, and waits right before the first use of gwin.
  bool quad_waited = false;
  const int* plan_flag = chain;
  if (plan_flag) {
    // bounded: if the flag does not show up (it always does), the ordinary wait below is still correct
    if (tid == 0) {
      int seen = 0;
      for (int spin = 0; spin < (1 << 16) && !seen; spin++) {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(plan_flag) : "memory");
      }
      s_seen = seen;
    }
    __syncthreads();
    if (!s_seen) {
      pdl_wait();
      quad_waited = true;
    }
  } else {
    pdl_wait();
    quad_waited = true;
  }
  if (tid == 0) { lat_stamp(g, 10); lat_stamp(g, 23); }
  // (read through L2: with the early start these are written while this kernel is resident)
  const int nwork = __ldcg(work_count);
  if (blockIdx.x == 0 && part == 0 && tid == 0) err_flag[7] = nwork;   // feedback for the host's choice of grid width
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const size_t idx = (size_t)__ldcg(work + wi);
    const int iom = (int)(idx / g.NI), it = (int)(idx % g.NI);
    int s, nabs, sg;
    decode_item(g, it, s, nabs, sg);
    const SpeciesDev& sp = g.sp[s];
    PlanEntry pe;
    {
      const int4 a = __ldcg(reinterpret_cast<const int4*>(plan + idx)), b = __ldcg(reinterpret_cast<const int4*>(plan + idx) + 1);
      pe.lo1 = a.x; pe.hi1 = a.y; pe.lo2 = a.z; pe.hi2 = a.w;
      pe.flags = b.x; pe.ipar_res = b.y; pe.upperlimit = b.z; pe.pad = b.w;
    }
    const int nperp = g.nperp, M_I = g.M_I, M_P = g.M_P;
    const double* ppar = sp.ppar;
    const cd omc = mk(__ldcg(om + 2 * iom), __ldcg(om + 2 * iom + 1));
    const int nn = sg ? -nabs : nabs;
    const double qs = sp.qs, ms = sp.ms, kpar = g.kpar;
    const double pR = (ms * omc.x - 1.0 * nn * qs) / kpar;
    const double pI = (ms * omc.y) / kpar;
    const cd p_res = mk(pR, pI);
    if (tid == 0) s_err = 0;
    __syncthreads();
    // this block's sums: combinations q0, q0 + 1 (near-pole pieces) or La, Lb, Lc (Landau blocks)
    cd acc[3];
    acc[0] = acc[1] = acc[2] = mk(0.0, 0.0);
    int err = 0, zero = 0;
    double near_fac = 0.0;   // factor of the near-pole terms (row slot 14)
    const int q0 = 2 * (part % 3);

    if ((pe.flags & PLAN_NEAR) && part <= 8) {
      const double* gw = gwin + idx * (size_t)g.WINX * 6;
      const int wbase = pe.ipar_res - M_I - 2;
      const double dppar = sp.dppar_signed;
      const double capDelta = pR - ppar[pe.ipar_res - M_I];
      const double smdelta = capDelta / (1.0 * M_P);
      const bool pairing = fabs(pI) > g.Tlim;
      const int piece = part / 3;
      // quadrature point j of this block's piece
      const double rest = ppar[pe.upperlimit] - pR - capDelta;
      const int ntiny = (int)(rest / smdelta);
      const double correction = ntiny > 0 ? (rest / (1.0 * ntiny)) / smdelta : 0.0;
      auto base = [&](int j) {   // pieces 0, 1: the point above the pole (piece 1 evaluates g at its mirror image)
        return (j == 0) ? pR : (j == M_P ? pR + capDelta : pR + smdelta * j);
      };
      auto point = [&](int j) {
        if (piece == 2) return (j == 0) ? pR + capDelta : pR + capDelta + correction * smdelta * j;
        const double p = base(j);
        return piece == 0 ? p : 2.0 * pR - p;
      };
      const int jfirst = (pairing || piece == 2) ? tid : 0;
      double pfirst = point(jfirst);
      if (!pairing && piece == 0) pfirst = wid == 0 ? pR + dppar : (wid == 1 ? pR - dppar : pR);   // analytic branch
      const GLoc Lfirst = funct_g6_locate(g, sp, wbase, pfirst);
      const double dp_abs = sp.dppar_abs;
      near_fac = 2.0 * PI * smdelta * sp.dpperp * 0.25;
      if (!quad_waited) {
        // the window sums are complete when every CTA of k_quad_mma has counted itself done (bounded; then the
        // ordinary wait for the predecessor, which is always sufficient)
        bool seen_all = false;
        if (chain && nquad > 0) {
          if (tid == 0) {
            int seen = 0;
            for (int spin = 0; spin < (1 << 16) && seen < nquad; spin++) {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(chain + CHAIN_QUAD) : "memory");
            }
            s_seen = seen >= nquad;
          }
          __syncthreads();
          seen_all = s_seen != 0;
        }
        if (!seen_all) pdl_wait();
        quad_waited = true;
      }
      // (from here to the end of the kernel the code is kept short: it runs once, after the wait, from L2)
      if (tid == 0) lat_stamp(g, 37);
      if ((pairing && piece <= 1) || piece == 2) {
        // Eq. (3.5): symmetric pairing around the pole, src/ALPS_fns.f90:1026-1082 (piece 0: g(p)/(p - p_res), piece 1:
        // -g(2 p_R - p)/(p - conj p_res)); tiny rest between p_res + capDelta and the first regular node, :1168-1230
        const int jend = piece == 2 ? (ntiny > 0 ? ntiny : -1) : M_P;
#pragma unroll 1
        for (int j = tid; j <= jend; j += LAT_THREADS) {
          const double wj = (j == 0 || j == jend) ? 1.0 : 2.0;
          const double pj = point(j);
          const double p = piece == 2 ? pj : base(j);            // the unmirrored point
          const GLoc L = (j == tid) ? Lfirst : funct_g6_locate(g, sp, wbase, pj);
          cd f[2];
          funct_g2_at(dp_abs, kpar, gw, L, q0, f, &err);
          const cd d = mk(p - pR, piece == 1 ? pI : -pI);
          if (piece == 2) {
            acc[0] += (wj * correction) * (f[0] / d);
            acc[1] += (wj * correction) * (f[1] / d);
          } else if (piece == 0) {
            acc[0] += wj * (f[0] / d);
            acc[1] += wj * (f[1] / d);
          } else {
            acc[0] -= wj * (f[0] / d);
            acc[1] -= wj * (f[1] / d);
          }
        }
      } else if (!pairing && piece == 0) {
        // Eq. (3.6): linearised integrand + pole term, src/ALPS_fns.f90:1088-1165; the three g values
        // are evaluated once, by three different warps
        if (wlane == 0 && wid < 3) {
          cd f[2];
          funct_g2_at(dp_abs, kpar, gw, Lfirst, q0, f, &err);
          s_f[wid][0] = f[0];
          s_f[wid][1] = f[1];
        }
        __syncthreads();
        cd gprime[2];
#pragma unroll
        for (int e = 0; e < 2; e++) gprime[e] = (s_f[0][e] - s_f[1][e]) / (2.0 * dppar);
#pragma unroll 1
        for (int j = 1 + tid; j <= M_P; j += LAT_THREADS) {
          const double wj = (j == M_P) ? 1.0 : 2.0;
          const double p = (j == M_P) ? pR + capDelta : pR + smdelta * j;
          const double x2 = (p - pR) * (p - pR);
          const double lor = x2 / (x2 + pI * pI);
          acc[0] += (wj * 2.0) * gprime[0] * lor;
          acc[1] += (wj * 2.0) * gprime[1] * lor;
        }
        if (tid == 0 && pI != 0.0) {
          const double sgn = pI > 0.0 ? 1.0 : -1.0;
          acc[0] += sgn * (cmul_i((2.0 * PI) * s_f[2][0]) / smdelta);
          acc[1] += sgn * (cmul_i((2.0 * PI) * s_f[2][1]) / smdelta);
        }
      }
    }

    if ((pe.flags & PLAN_LANDAU) && part >= 9) {
      // landau_integrate, src/ALPS_fns.f90:1327-1452; acc[0..2] collect La, Lb, Lc
      const double dpperp = sp.dpperp, dppar = sp.dppar_abs;
      const double* Jn = sp.J + (size_t)(nabs + 1) * sp.ldj;
      const double* Jm = sp.J + (size_t)nabs * sp.ldj;
      const double* Jp = sp.J + (size_t)(nabs + 2) * sp.ldj;
      const cd ppl = mk(pR + dppar, pI), pmi = mk(pR - dppar, pI);
      const int gidx = tid >> 2, sub = tid & 3, lbase = wlane & ~3;
#pragma unroll 1
      for (int r0 = 64 * (part - 9); r0 <= nperp; r0 += 128) {
        const int r = r0 + gidx;
        const bool valid = r <= nperp;
        const bool interior = r >= 1 && r <= nperp - 1;
        const int hi = (r == 0) ? 1 : (r == nperp ? nperp : r + 1);
        const int lo = (r == 0) ? 0 : (r == nperp ? nperp - 1 : r - 1);
        cd v = mk(0.0, 0.0);
        if (valid) {
          // one of the four values of the row's differences per thread (one copy of eval_fit in the code)
          const int row = sub < 2 ? r : (sub == 2 ? hi : lo);
          const cd pp = sub == 0 ? ppl : (sub == 1 ? pmi : p_res);
          v = eval_fit(g, s, row, pp);
        }
        cd fpar_i, fpar_f, fperp_i, fperp_f;
        fpar_i.x = __shfl_sync(0xffffffffu, v.x, lbase + 0);  fpar_i.y = __shfl_sync(0xffffffffu, v.y, lbase + 0);
        fpar_f.x = __shfl_sync(0xffffffffu, v.x, lbase + 1);  fpar_f.y = __shfl_sync(0xffffffffu, v.y, lbase + 1);
        fperp_i.x = __shfl_sync(0xffffffffu, v.x, lbase + 2); fperp_i.y = __shfl_sync(0xffffffffu, v.y, lbase + 2);
        fperp_f.x = __shfl_sync(0xffffffffu, v.x, lbase + 3); fperp_f.y = __shfl_sync(0xffffffffu, v.y, lbase + 3);
        if (valid && sub == 0) {
          // the reference tests fpar_f twice and never fperp_f (lines 1404-1405); interior rows only
          if (interior && ((fpar_i.x == 0.0 && fpar_i.y == 0.0) || (fpar_f.x == 0.0 && fpar_f.y == 0.0) ||
                           (fperp_i.x == 0.0 && fperp_i.y == 0.0)))
            zero = 1;
          const double h = (r == 0 || r >= nperp - 1) ? 0.5 : 1.0;
          const cd dfperp = (fperp_i - fperp_f) / (interior ? 2.0 * dpperp : dpperp);
          const cd dfpar = (fpar_i - fpar_f) / (2.0 * dppar);
          const double pperp = sp.pperp[r];
          const cd Q = (qs / fabs(kpar)) * (((pperp * dfpar - p_res * dfperp) * kpar) / ms + omc * dfperp);
          const double bj = Jn[r];
          const double bp = (nabs >= 1) ? 0.5 * (Jm[r] - Jp[r]) : -Jp[r];
          acc[0] += (h * (bj * bj)) * Q;
          acc[1] += (h * (pperp * (bj * bp))) * Q;
          acc[2] += (h * ((pperp * pperp) * (bp * bp))) * Q;
        }
      }
    }

#ifdef ALPS_LAT_TRACE
    if (tid == 0 && acc[0].x != 1.2345e300) lat_stamp(g, part >= 9 ? 9 : 11);
#endif
    // ---- block sum -> partial row of this part (the sums this block does not hold are zero)
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const cd t = warp_sum_c(acc[q]);
      if (wlane == 0) s_part[wid][q] = t;
    }
    if (err) atomicMax(&s_err, err);
    zero = __syncthreads_or(zero);
    if (tid == 0) lat_stamp(g, 43);
    double* prow = Spart + (idx * LAT_PARTS + part) * LAT_STRIDE;
    if (tid < 6) {
      // slot of sum tid among this block's three: near-pole blocks hold q0, q0 + 1, Landau blocks 0, 1, 2
      const int slot = part >= 9 ? (tid < 3 ? tid : -1) : ((tid == q0 || tid == q0 + 1) ? tid - q0 : -1);
      cd t = mk(0.0, 0.0);
      if (slot >= 0) {
        t = s_part[0][slot];
#pragma unroll 1
        for (int w = 1; w < LAT_THREADS / 32; w++) t += s_part[w][slot];
      }
      prow[2 * tid] = t.x;
      prow[2 * tid + 1] = t.y;
    }
    if (tid == 6) {
      prow[12] = zero ? 1.0 : 0.0;
      prow[13] = (double)s_err;
      prow[14] = near_fac;
      if (s_err) {
        err_flag[0] = 1;
        err_flag[1] = (int)idx;
        err_flag[2] = pe.ipar_res;
        err_flag[3] = pe.upperlimit;
        err_flag[4] = pe.flags;
        err_flag[5] = s_err;
      }
      lat_stamp(g, 13);
      lat_stamp(g, 15);
    }
    __syncthreads();
  }
  if (chain) {
    // k_chi_assemble, already resident, adds the resonant parts when every block has got here
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(chain + CHAIN_RES, 1);
    }
  }
}

// The resonant part of moment sum q of one item from the partial rows of k_resonant_lat (pr = the item's LAT_PARTS rows):
// near-pole pieces 0..8 (rows q/2, 3 + q/2, 6 + q/2 hold combination q) times their factor, plus the Landau residue of
// rows 9, 10 -- landau = -(sum) * i * dpperp * pi * 2 pi with factor 2 (Im om < 0) or 1 (Im om == 0),
// full_integrate src/ALPS_fns.f90:782-789 -- times p_res^m for the p_par power m of the combination.
// (mult = (Im om < 0 ? 2 : 1) dpperp pi 2 pi, formed by the caller: lat_landau_mult)
__device__ __forceinline__ double lat_landau_mult(cd omc, double dpperp) {
  return (omc.y < 0.0 ? 2.0 : 1.0) * dpperp * PI * 2.0 * PI;
}
__device__ __forceinline__ cd lat_parts_combine(const double* pr, int q, int flags, double mult, cd p_res) {
  auto ld = [&](int prt, int qq) {
    const double2 v = __ldcg(reinterpret_cast<const double2*>(pr + prt * LAT_STRIDE + 2 * qq));
    return mk(v.x, v.y);
  };
  const int x = (q < 3) ? 0 : (q < 5 ? 1 : 2), m = (q < 3) ? q : (q < 5 ? q - 3 : 0);
  // all loads first
  const cd n0 = ld(q / 2, q), n1 = ld(3 + q / 2, q), n2 = ld(6 + q / 2, q), l0 = ld(9, x), l1 = ld(10, x);
  const double near_fac = __ldcg(pr + (q / 2) * LAT_STRIDE + 14);
  const double z0 = __ldcg(pr + 9 * LAT_STRIDE + 12), z1 = __ldcg(pr + 10 * LAT_STRIDE + 12);
  cd tot = mk(0.0, 0.0);
  if (flags & PLAN_NEAR) tot += near_fac * ((n0 + n1) + n2);
  const bool zr = z0 != 0.0 || z1 != 0.0;
  if ((flags & PLAN_LANDAU) && !zr) {
    const cd cx = cmul_i(-(l0 + l1)) * mult;
    tot += (m == 0) ? cx : (m == 1 ? p_res * cx : (p_res * p_res) * cx);
  }
  return tot;
}

// ---------------------------------------------------------------- chi partial
// One block per omega: tensor components of every (n, sign) from its six moment sums, summed over the harmonics of
// this process' shard (the harmonic sums of disp, src/ALPS_fns.f90:364-516).
// Phase 1: a thread takes (item, moment sum q) units -- a tensor component needs exactly one of the six sums -- and
// issues every load of its units (plan entry, the p_par split rows of k_quad, the resonant part) before the first use:
// the whole omega costs one or two L2 round trips instead of one per row and item (the single-omega chain is a chain of
// such round trips).  The components go to shared memory.  Phase 2: for species s and component c, lane l of a warp adds
// the NON-RESONANT items r = l, l + 32, ... in increasing order and a butterfly adds the lanes; the resonant items (those
// whose sum needs k_resonant's output) are then added in increasing order: chi = butterfly(lane sums) + (resonant sum).
// The order depends on the configuration only, whatever the batch size -- and it lets the single-omega chain form
// everything but the resonant items while k_resonant_lat still runs.
// partial[(iom*nspec + s)*PARTIAL_PER_SPEC + 2*c .. ]: c = mode-1 (0..5) for chi,
// c = 6 + 3*(mode-1) + (m+1) for chi_low(mode, m), m = -1,0,1.
constexpr int CHI_WARPS = 24;
constexpr int CHI_THREADS = 32 * CHI_WARPS;
constexpr int CHI_WIN = 256;       // items (species after species) per pass
constexpr int CHI_UNROLL = 1;      // units a thread keeps in flight
constexpr int CHI_ROWS = 8;        // split rows fetched together
constexpr int CHI_PAIRS = (6 * MAXSPEC + CHI_WARPS - 1) / CHI_WARPS;   // (species, component) sums per warp
struct ChiSpec {
  double z, kf1, kf2, cbulk, ee, norm, qs, ms, dpperp;
  int base, nitems, fbase, table, ee_on, ee_low, usebM, pad;
};
// A resonant unit of the early mode: everything its completion needs once k_resonant_lat has finished
struct ChiRec {
  cd S, p_res;            // bulk sum (scaled), resonance
  double fac, mult;       // component factor, Landau factor
  const double* pr;       // the item's partial rows
  short q, rot, cm, lowm, fl, sx, flags, pad;   // lowm: chi_low slot m + 1, or -1
};
constexpr int CHI_REC = 96;
struct ChiGlobals {
  double kperp, kpar, vA;
  int kperp_norm, pad;
};
struct ChiSmem {
  cd mode[CHI_WIN][6];
  cd low[MAXSPEC][6][3];
  double partial[MAXSPEC][PARTIAL_PER_SPEC];
  ChiSpec spc[MAXSPEC];
  int iflags[CHI_WIN];       // summation class of the window's items (chi_partial_block)
  cd chinr[6 * MAXSPEC], accr[6 * MAXSPEC];   // per (species, component): lane sums added up / resonant items
  ChiRec rec[CHI_REC];
  int nrec;
  ChiGlobals gc;
  double om[2];              // this block's omega
  int total, seen;
};
// Spart != nullptr: the resonant parts come from the partial rows of k_resonant_lat (lat_parts_combine) instead of Sres.
// do_wait (fused kernel of the single-omega chain, launched programmatically while its predecessors run): the block
// waits itself -- for its predecessor before the first load of the chain's data, or, with quad_done (a counter the
// nquad CTAs of k_quad_mma bump when their sums are written; cleared by k_plan), for k_quad_mma only: the bulk sums of
// every item are then formed while k_resonant_lat still runs and only the resonant parts are added after the wait.
// Tensor component fed by moment sum q of harmonic nn: rot(fac * S) with rot = 1 (0), i (1), -i (2); cm = its index
// (src/ALPS_fns.f90:395-470).
__device__ __forceinline__ double chi_factor(int q, double nn, const double z, const double kf1, const double kf2,
                                             int& rot, int& cm) {
  rot = 0;
  switch (q) {
    case 0: cm = 0; return (nn * nn) / (z * z);        // xx: n^2 J^2 / z^2
    case 1: cm = 4; return kf1 * nn / z;               // xz: n J^2 p_par / z
    case 2: cm = 2; return kf2;                        // zz: J^2 p_par^2
    case 3: cm = 3; rot = 1; return kf1 * nn / z;      // xy: i p_perp n J J' / z
    case 4: cm = 5; rot = 2; return kf2;               // yz: -i J J' p_par p_perp
    default: cm = 1; return kf2;                       // yy: p_perp^2 J'^2
  }
}
__device__ __forceinline__ cd chi_rotate(cd v, int rot) { return rot == 0 ? v : (rot == 1 ? cmul_i(v) : -cmul_i(v)); }
__device__ __forceinline__ cd chi_component(cd S, int q, double nn, const double z, const double kf1, const double kf2,
                                            int& cm) {
  int rot;
  const double fac = chi_factor(q, nn, z, kf1, kf2, rot, cm);
  return chi_rotate(fac * S, rot);
}
__device__ __forceinline__ int chi_slot(int q) { return q == 0 ? 0 : (q == 1 ? 4 : (q == 2 ? 2 : (q == 3 ? 3 : (q == 4 ? 5 : 1)))); }
__device__ __forceinline__ void chi_partial_block(const GlobalDev& g, const double* __restrict__ om, int iom,
                                                  const PlanEntry* __restrict__ plan, const double* __restrict__ Sbulk,
                                                  int nsplit, const double* __restrict__ Sres, const double* Spart,
                                                  double* partial, ChiSmem& sm, bool do_wait = false,
                                                  const int* quad_done = nullptr, int nquad = 0,
                                                  const int* res_done = nullptr, int nres = 0) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nspec = g.nspec;
  // (locals: values of *gp used after a wait would otherwise be fetched again, one L2 round trip each)
  const int NI = g.NI;
  const double kpar = g.kpar;
  // ---- phase 0: species constants (nothing of the chain's data)
  if (tid < nspec) {
    const SpeciesDev& sp = g.sp[tid];
    ChiSpec c;
    c.table = !sp.usebM;
    c.z = g.kperp_norm ? g.kperp / sp.qs : 1.0 / sp.qs;
    c.kf1 = g.kperp_norm ? 1.0 : g.kperp;
    c.kf2 = g.kperp_norm ? 1.0 : g.kperp * g.kperp;
    c.cbulk = 2.0 * PI * sp.dpperp * sp.dppar_abs * 0.25;
    c.ee = g.kperp_norm ? sp.int_ee : g.kperp * g.kperp * sp.int_ee;
    c.ee_on = c.table && sp.nlo_shard == 0;
    c.ee_low = c.ee_on && !sp.relativistic;   // int_ee_rel goes into chi only (src/ALPS_fns.f90:481-486)
    c.norm = sp.ns * sp.qs;
    c.qs = sp.qs;
    c.ms = sp.ms;
    c.dpperp = sp.dpperp;
    c.base = sp.item_base;
    c.nitems = c.table ? 2 * (sp.nhi + 1) : 0;
    c.usebM = sp.usebM;
    int fb = 0;
    for (int t = 0; t < tid; t++) fb += g.sp[t].usebM ? 0 : 2 * (g.sp[t].nhi + 1);
    c.fbase = fb;
    sm.spc[tid] = c;
    if (tid == nspec - 1) sm.total = fb + c.nitems;
  }
  if (tid == 32) {
    sm.gc.kperp = g.kperp;
    sm.gc.kpar = g.kpar;
    sm.gc.vA = g.vA;
    sm.gc.kperp_norm = g.kperp_norm;
  }
  for (int i = tid; i < MAXSPEC * 18; i += CHI_THREADS) (&sm.low[0][0][0])[i] = mk(0.0, 0.0);
  __syncthreads();
  if (lane == 0) lat_stamp(g, 25);
  const int total = sm.total;
  bool early = false;
  if (do_wait) {
    if (quad_done && nquad > 0 && total > 0 && total <= CHI_WIN && Spart) {
      // bounded: if the count does not show up (it always does), the ordinary wait is still correct
      if (tid == 0) {
        int seen = 0;
        for (int spin = 0; spin < (1 << 16) && seen < nquad; spin++) {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(quad_done) : "memory");
        }
        sm.seen = seen >= nquad;
      }
      __syncthreads();
      early = sm.seen != 0;
    }
    if (!early) pdl_wait();
    if (tid == 0) lat_stamp(g, 18);
  }
  const cd omc = mk(__ldcg(om + 2 * iom), __ldcg(om + 2 * iom + 1));
  if (tid == 0) {
    sm.om[0] = omc.x;
    sm.om[1] = omc.y;
  }
  // per (species, component) of this warp: ordered lane sums of the non-resonant items (the ordered sum of the resonant
  // ones lives in sm.accr).  The code that runs after the last wait is kept short and rolled: it is executed once, from
  // L2 (the instruction caches hold 32 KB), and costs about 1.5 ns per instruction.
  cd acc[CHI_PAIRS];
#pragma unroll
  for (int i = 0; i < CHI_PAIRS; i++) acc[i] = mk(0.0, 0.0);
  if (tid < 6 * MAXSPEC) sm.accr[tid] = mk(0.0, 0.0);
  if (tid == 0) sm.nrec = 0;
  __syncthreads();

  for (int f0 = 0; f0 < total; f0 += CHI_WIN) {
    const int nun = min(CHI_WIN, total - f0) * 6;
    // ---- phase 1: one tensor component per unit into shared memory (early: the resonant units keep their bulk sum there)
    for (int u0 = tid; u0 < nun; u0 += CHI_THREADS * CHI_UNROLL) {
      int sK[CHI_UNROLL], rK[CHI_UNROLL];
      size_t idxK[CHI_UNROLL];
      int4 peb[CHI_UNROLL];       // flags, ipar_res, upperlimit, pad of the plan entry
      double2 row[CHI_UNROLL][CHI_ROWS], sres[CHI_UNROLL];
#pragma unroll
      for (int k = 0; k < CHI_UNROLL; k++) {
        const int u = u0 + k * CHI_THREADS;
        if (u >= nun) continue;
        const int fl = u / 6, q = u - 6 * fl, f = f0 + fl;
        int sx = 0;
        for (int t = 1; t < nspec; t++)
          if (f >= sm.spc[t].fbase) sx = t;
        sK[k] = sx;
        rK[k] = f - sm.spc[sx].fbase;
        const size_t idx = (size_t)iom * NI + sm.spc[sx].base + rK[k];
        idxK[k] = idx;
        // everything the unit may need, in flight together (rows and resonant parts of inactive items are never used);
        // through L2: in the early mode the data is written while this kernel is resident
        peb[k] = __ldcg(reinterpret_cast<const int4*>(plan + idx) + 1);
#pragma unroll
        for (int v = 0; v < CHI_ROWS; v++)
          if (v < nsplit) row[k][v] = __ldcg(reinterpret_cast<const double2*>(Sbulk + (idx * nsplit + v) * 12 + 2 * q));
        sres[k] = early ? make_double2(0.0, 0.0) : __ldcg(reinterpret_cast<const double2*>(Sres + idx * 12 + 2 * q));
      }
#ifdef ALPS_LAT_TRACE
      if (lane == 0 && (peb[0].x | 1)) lat_stamp(g, 35);   // the first plan entry has arrived
#endif
#pragma unroll
      for (int k = 0; k < CHI_UNROLL; k++) {
        const int u = u0 + k * CHI_THREADS;
        if (u >= nun) continue;
        const int fl = u / 6, q = u - 6 * fl;
        const int sx = sK[k], r = rK[k], nabs = r >> 1, sg = r & 1, flags = peb[k].x;
        const ChiSpec& c = sm.spc[sx];
        const double nn = sg ? -(double)nabs : (double)nabs;
        cd m = mk(0.0, 0.0);
        int cm = (flags & PLAN_REL) ? q : chi_slot(q);     // tensor component this sum feeds
        bool on = (flags & PLAN_ACTIVE) != 0;
        // n = 0: only yy, zz, yz are evaluated (src/ALPS_fns.f90:368-393)
        if (nabs == 0 && !(cm == 1 || cm == 2 || cm == 5)) on = false;
        bool deferred = false;
        if (on) {
          if (flags & PLAN_REL) {
            // relativistic species: k_rel already produced the six tensor components
            m = mk(sres[k].x, sres[k].y);
          } else {
            cd S = mk(0.0, 0.0);
            // partial rows of the p_par splits of k_quad, added in order
#pragma unroll
            for (int v = 0; v < CHI_ROWS; v++)
              if (v < nsplit) S += mk(row[k][v].x, row[k][v].y);
            for (int v = CHI_ROWS; v < nsplit; v++) {
              const double2 t = __ldcg(reinterpret_cast<const double2*>(Sbulk + (idxK[k] * nsplit + v) * 12 + 2 * q));
              S += mk(t.x, t.y);
            }
            S = c.cbulk * S;
            if (flags & (PLAN_NEAR | PLAN_LANDAU)) {
              const int nni = sg ? -nabs : nabs;
              const double pR = (c.ms * omc.x - 1.0 * nni * c.qs) / kpar;
              const double pI = (c.ms * omc.y) / kpar;
              if (early) {
                deferred = true;
                m = S;
                const int slot = atomicAdd(&sm.nrec, 1);
                if (slot < CHI_REC) {
                  ChiRec R;
                  int rot, cm2;
                  R.S = S;
                  R.p_res = mk(pR, pI);
                  R.fac = chi_factor(q, nn, c.z, c.kf1, c.kf2, rot, cm2);
                  R.mult = lat_landau_mult(omc, c.dpperp);
                  R.pr = Spart + idxK[k] * (size_t)(LAT_PARTS * LAT_STRIDE);
                  R.q = (short)q; R.rot = (short)rot; R.cm = (short)cm2;
                  R.lowm = (short)(nabs <= 1 ? (nabs == 0 ? 1 : (sg ? 0 : 2)) : -1);
                  R.fl = (short)fl; R.sx = (short)sx; R.flags = (short)flags; R.pad = 0;
                  sm.rec[slot] = R;
                }
              } else if (Spart) {
                S += lat_parts_combine(Spart + idxK[k] * (size_t)(LAT_PARTS * LAT_STRIDE), q, flags,
                                       lat_landau_mult(omc, c.dpperp), mk(pR, pI));
              } else {
                S += mk(sres[k].x, sres[k].y);
              }
            }
            if (!deferred) m = chi_component(S, q, nn, c.z, c.kf1, c.kf2, cm);
          }
        }
        sm.mode[fl][cm] = m;
        // the item's summation class: 0 not summed, 1 in the ordered lane sums, 2 resonant (added afterwards, in order)
        if (q == 0)
          sm.iflags[fl] = !(flags & PLAN_ACTIVE) ? 0
                          : (((flags & (PLAN_NEAR | PLAN_LANDAU)) && !(flags & PLAN_REL))
                                 ? 2 | ((flags & PLAN_NEAR) ? 4 : 0) | ((flags & PLAN_LANDAU) ? 8 : 0) : 1);
        if (on && !deferred && nabs <= 1) sm.low[sx][cm][nabs == 0 ? 1 : (sg ? 0 : 2)] = m;
      }
    }
    if (lane == 0) lat_stamp(g, 27);
    __syncthreads();
    // ---- phase 2: ordered lane sums of the non-resonant items (nothing of k_resonant_lat in them)
#pragma unroll
    for (int i = 0; i < CHI_PAIRS; i++) {
      const int p = warp + CHI_WARPS * i;
      if (p >= 6 * nspec) continue;
      const int sx = p / 6, cc = p - 6 * sx;
      const int off = sm.spc[sx].fbase - f0;                         // window slot of item 0 of the species
      const int ra = max(0, -off), rb = min(sm.spc[sx].nitems, CHI_WIN - off);
      for (int r = ra + ((lane - ra) & 31); r < rb; r += 32)
        if ((sm.iflags[off + r] & 3) == 1) acc[i] += sm.mode[off + r][cc];
    }
    if (lane == 0) lat_stamp(g, 29);
    if (early) {
      // (single window) the lane sums are complete: add the lanes now
#pragma unroll
      for (int i = 0; i < CHI_PAIRS; i++) {
        const int p = warp + CHI_WARPS * i;
        if (p >= 6 * nspec) continue;
        const cd v = warp_sum_c(acc[i]);
        if (lane == 0) sm.chinr[p] = v;
      }
      // ---- the resonant parts, once k_resonant_lat has finished: its blocks count themselves in res_done (bounded
      // wait, then the ordinary one, which is always correct)
      {
        bool seen_all = false;
        if (res_done && nres > 0) {
          if (tid == 0) {
            int seen = 0;
            for (int spin = 0; spin < (1 << 16) && seen < nres; spin++) {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(res_done) : "memory");
            }
            sm.seen = seen >= nres;
          }
          __syncthreads();
          seen_all = sm.seen != 0;
        }
        if (!seen_all) pdl_wait();
      }
      if (tid == 0) lat_stamp(g, 39);
      const int nrec = sm.nrec;
      // (short code: it runs once, after the chain's last wait, from L2)
#pragma unroll 1
      for (int i = tid; i < nrec && nrec <= CHI_REC; i += CHI_THREADS) {
        const ChiRec& R = sm.rec[i];
        const cd S = R.S + lat_parts_combine(R.pr, R.q, R.flags, R.mult, R.p_res);
        const cd m = chi_rotate(R.fac * S, R.rot);
        sm.mode[R.fl][R.cm] = m;
        if (R.lowm >= 0) sm.low[R.sx][R.cm][R.lowm] = m;
#ifdef ALPS_LAT_TRACE
        if (S.x != 1.2345e300) lat_stamp(g, 47);
#endif
      }
      // more resonant units than records: the general way
      for (int u = tid; u < nun && nrec > CHI_REC; u += CHI_THREADS) {
        const int fl = u / 6, q = u - 6 * fl;
        const int icl = sm.iflags[fl];
        if ((icl & 3) != 2) continue;
        const int flags = ((icl & 4) ? PLAN_NEAR : 0) | ((icl & 8) ? PLAN_LANDAU : 0);
        const int f = f0 + fl;
        int sx = 0;
        for (int t = 1; t < nspec; t++)
          if (f >= sm.spc[t].fbase) sx = t;
        const ChiSpec& c = sm.spc[sx];
        const int r = f - c.fbase, nabs = r >> 1, sg = r & 1;
        int cm = chi_slot(q);
        if (nabs == 0 && !(cm == 1 || cm == 2 || cm == 5)) continue;
        const size_t idx = (size_t)iom * NI + c.base + r;
        const double nn = sg ? -(double)nabs : (double)nabs;
        const int nni = sg ? -nabs : nabs;
        const double pR = (c.ms * omc.x - 1.0 * nni * c.qs) / kpar;
        const double pI = (c.ms * omc.y) / kpar;
        cd S = sm.mode[fl][cm];
        S += lat_parts_combine(Spart + idx * (size_t)(LAT_PARTS * LAT_STRIDE), q, flags, lat_landau_mult(omc, c.dpperp),
                               mk(pR, pI));
        const cd m = chi_component(S, q, nn, c.z, c.kf1, c.kf2, cm);
        sm.mode[fl][cm] = m;
        if (nabs <= 1) sm.low[sx][cm][nabs == 0 ? 1 : (sg ? 0 : 2)] = m;
      }
      __syncthreads();
    }
    // ---- the resonant items, in increasing order
#pragma unroll 1
    for (int p = warp; p < 6 * nspec; p += CHI_WARPS) {
      const int sx = p / 6, cc = p - 6 * sx;
      const int off = sm.spc[sx].fbase - f0;
      const int ra = max(0, -off), rb = min(sm.spc[sx].nitems, CHI_WIN - off);
      cd a = sm.accr[p];
#pragma unroll 1
      for (int r0 = ra; r0 < rb; r0 += 32) {
        const int r = r0 + lane;
        const bool res = r < rb && (sm.iflags[off + r] & 3) == 2;
        const cd mine = res ? sm.mode[off + r][cc] : mk(0.0, 0.0);
        unsigned mask = __ballot_sync(0xffffffffu, res);
        while (mask) {
          const int l = __ffs(mask) - 1;
          mask &= mask - 1;
          a += mk(__shfl_sync(0xffffffffu, mine.x, l), __shfl_sync(0xffffffffu, mine.y, l));
        }
      }
      __syncwarp();
      if (lane == 0) sm.accr[p] = a;
    }
    __syncthreads();
  }
  if (!early) {
#pragma unroll
    for (int i = 0; i < CHI_PAIRS; i++) {
      const int p = warp + CHI_WARPS * i;
      if (p >= 6 * nspec) continue;
      const cd v = warp_sum_c(acc[i]);
      if (lane == 0) sm.chinr[p] = v;
    }
    __syncthreads();
  }
  // chi = (lane sums of the non-resonant items) + (resonant items); int_ee and the species' normalisation
  if (tid < 6 * nspec) {
    const int sx = tid / 6, cc = tid - 6 * sx;
    const ChiSpec& c = sm.spc[sx];
    cd v = sm.chinr[tid] + sm.accr[tid];
    if (cc == 2 && c.ee_on) v.x += c.ee;
    double* o = partial + ((size_t)iom * nspec + sx) * PARTIAL_PER_SPEC;
    sm.partial[sx][2 * cc] = o[2 * cc] = c.norm * v.x;
    sm.partial[sx][2 * cc + 1] = o[2 * cc + 1] = c.norm * v.y;
  }
  for (int i = tid; i < nspec * 18; i += CHI_THREADS) {
    const int sx = i / 18, cc = (i - 18 * sx) / 3, mm = i % 3;
    const ChiSpec& c = sm.spc[sx];
    cd v = sm.low[sx][cc][mm];
    if (cc == 2 && mm == 1 && c.ee_low) v.x += c.ee;
    double* o = partial + ((size_t)iom * nspec + sx) * PARTIAL_PER_SPEC;
    sm.partial[sx][2 * (6 + 3 * cc + mm)] = o[2 * (6 + 3 * cc + mm)] = c.norm * v.x;
    sm.partial[sx][2 * (6 + 3 * cc + mm) + 1] = o[2 * (6 + 3 * cc + mm) + 1] = c.norm * v.y;
  }
}

__global__ void __launch_bounds__(CHI_THREADS) k_chi_partial(const GlobalDev* __restrict__ gp,
                                                                const double* __restrict__ om, int n_om,
                                                                const PlanEntry* __restrict__ plan,
                                                                const double* __restrict__ Sbulk, int nsplit,
                                                                const double* __restrict__ Sres,
                                                                const double* Spart, double* __restrict__ partial) {
  (void)n_om;
  __shared__ ChiSmem sm;
  chi_partial_block(*gp, om, blockIdx.x, plan, Sbulk, nsplit, Sres, Spart, partial, sm);
}

// -------------------------------------------------------------------- assemble
// One thread per omega: src/ALPS_fns.f90:536-624.
// prow: the nspec partial rows of this omega (global memory, or the block's copy in shared memory); gc, usebM (bit s:
// species s takes its chi from ext_chi): the constants of *gp, read by the caller (the fused kernel reads them before
// its waits)
__device__ __forceinline__ void assemble_one(const ChiGlobals gc, int nspec, unsigned usebM, cd omc, int iom,
                                             const double* prow, const double* __restrict__ ext_chi,
                                             double* __restrict__ D, double* __restrict__ chi0_out,
                                             double* __restrict__ chi0_low_out, double* __restrict__ wave_out) {
  const double kperp = gc.kperp, kpar = gc.kpar, vA = gc.vA;
  cd enx2, enz2, enxnz, norm2;
  if (gc.kperp_norm) {
    enx2 = mk(kperp * kperp, 0.0);
    enz2 = mk(kpar * kpar, 0.0);
    enxnz = mk(kpar * kperp, 0.0);
    norm2 = omc * omc * vA * vA;
  } else {
    enx2 = mk(kperp * kperp * kperp * kperp, 0.0);
    enz2 = mk(kpar * kpar * kperp * kperp, 0.0);
    enxnz = mk(kpar * kperp * kperp * kperp, 0.0);
    norm2 = omc * omc * vA * vA * kperp * kperp;
  }
  // mode index c -> (i,j): xx yy zz xy xz yz
  const int MI[6] = {0, 1, 2, 0, 0, 1}, MJ[6] = {0, 1, 2, 1, 2, 2};
  cd eps[6];
  for (int c = 0; c < 6; c++) eps[c] = mk(0.0, 0.0);
  for (int s = 0; s < nspec; s++) {
    const double* p = prow + (size_t)s * PARTIAL_PER_SPEC;
    const double* x = ext_chi ? ext_chi + ((size_t)iom * nspec + s) * PARTIAL_PER_SPEC : nullptr;
    const bool useext = x && ((usebM >> s) & 1u);
    for (int c = 0; c < 6; c++) {
      cd v = mk(p[2 * c], p[2 * c + 1]);
      if (useext) v += mk(x[2 * c], x[2 * c + 1]);
      eps[c] += v;
      if (chi0_out) {
        cd q = v / norm2;
        // chi0(is,i,j) Fortran order, symmetric completion lines 544-546
        double* o = chi0_out + (size_t)iom * nspec * 18;
        size_t k1 = s + (size_t)nspec * (MI[c] + 3 * MJ[c]);
        o[2 * k1] = q.x;
        o[2 * k1 + 1] = q.y;
        if (MI[c] != MJ[c]) {
          const double sgn = (c == 4) ? 1.0 : -1.0;   // (3,1)=+(1,3); (2,1)=-(1,2); (3,2)=-(2,3)
          size_t k2 = s + (size_t)nspec * (MJ[c] + 3 * MI[c]);
          o[2 * k2] = sgn * q.x;
          o[2 * k2 + 1] = sgn * q.y;
        }
      }
      if (chi0_low_out) {
        double* o = chi0_low_out + (size_t)iom * nspec * 54;
        for (int m = 0; m < 3; m++) {
          cd vl = mk(p[2 * (6 + 3 * c + m)], p[2 * (6 + 3 * c + m) + 1]);
          if (useext) vl += mk(x[2 * (6 + 3 * c + m)], x[2 * (6 + 3 * c + m) + 1]);
          cd q = vl / norm2;
          size_t k1 = s + (size_t)nspec * (MI[c] + 3 * (MJ[c] + 3 * m));
          o[2 * k1] = q.x;
          o[2 * k1 + 1] = q.y;
          if (MI[c] != MJ[c]) {
            const double sgn = (c == 4) ? 1.0 : -1.0;
            size_t k2 = s + (size_t)nspec * (MJ[c] + 3 * (MI[c] + 3 * m));
            o[2 * k2] = sgn * q.x;
            o[2 * k2 + 1] = sgn * q.y;
          }
        }
      }
    }
  }
  const cd ov = omc * vA;
  const cd unit = gc.kperp_norm ? ov * ov : (kperp * ov) * (kperp * ov);
  eps[0] += unit;
  eps[1] += unit;
  eps[2] += unit;
  const cd w11 = eps[0] - enz2, w22 = eps[1] - enz2 - enx2, w33 = eps[2] - enx2;
  const cd w13 = eps[4] + enxnz, w12 = eps[3], w23 = eps[5];
  const cd d = w11 * (w22 * w33 + w23 * w23) + mk(2.0, 0.0) * w12 * w23 * w13 - w13 * w13 * w22 + w12 * w12 * w33;
  if (D) {
    D[2 * iom] = d.x;
    D[2 * iom + 1] = d.y;
  }
  if (wave_out) {
    double* o = wave_out + (size_t)iom * 18;
    const cd W[3][3] = {{w11, w12, w13}, {-w12, w22, w23}, {w13, -w23, w33}};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        o[2 * (i + 3 * j)] = W[i][j].x;
        o[2 * (i + 3 * j) + 1] = W[i][j].y;
      }
  }
}

__global__ void k_assemble(const GlobalDev* __restrict__ gp, const double* __restrict__ om, int n_om,
                           const double* __restrict__ partial, const double* __restrict__ ext_chi,
                           double* __restrict__ D, double* __restrict__ chi0_out, double* __restrict__ chi0_low_out,
                           double* __restrict__ wave_out, const int* __restrict__ err_src, int* __restrict__ err_dst) {
  const int iom = blockIdx.x * blockDim.x + threadIdx.x;
  // single-omega graph (api.cu): D and err_dst are pinned host memory, so the chain needs no device-to-host copies
  if (err_dst && iom < 8) err_dst[iom] = err_src[iom];
  if (iom >= n_om) return;
  const GlobalDev& g = *gp;
  ChiGlobals gc;
  gc.kperp = g.kperp;
  gc.kpar = g.kpar;
  gc.vA = g.vA;
  gc.kperp_norm = g.kperp_norm;
  unsigned usebM = 0;
  for (int s = 0; s < g.nspec; s++) usebM |= g.sp[s].usebM ? (1u << s) : 0u;
  assemble_one(gc, g.nspec, usebM, mk(om[2 * iom], om[2 * iom + 1]), iom,
               partial + (size_t)iom * g.nspec * PARTIAL_PER_SPEC, ext_chi, D, chi0_out, chi0_low_out, wave_out);
}

// k_chi_partial + k_assemble of a small batch in one launch (the single-omega graph and batched roots): block =
// one omega; the same device functions in the same order of operations as the two kernels (bitwise the two-kernel
// result); thread 0 assembles after the barrier from the block's copy of the partial rows.
__global__ void __launch_bounds__(CHI_THREADS)
k_chi_assemble(const GlobalDev* __restrict__ gp, const double* __restrict__ om, const PlanEntry* __restrict__ plan,
               const double* __restrict__ Sbulk, int nsplit, const double* __restrict__ Sres, const double* Spart,
               double* partial, const double* __restrict__ ext_chi, double* __restrict__ D, double* __restrict__ chi0_out,
               double* __restrict__ chi0_low_out, double* __restrict__ wave_out, const int* err_src,
               int* __restrict__ err_dst, int* chain, int nquad, int nres, const double* nh_target) {
  const GlobalDev& g = *gp;
  const int iom = blockIdx.x, nspec = g.nspec;
  __shared__ ChiSmem sm;
  pdl_trigger();
  if (threadIdx.x == 0) lat_stamp(g, 16);
  // (the host's announcement for the k_nhds count -- a read over PCIe -- is fetched first, behind everything else)
  double nh_tgt = 0.0;
  if (nh_target && threadIdx.x == 0) nh_tgt = *reinterpret_cast<const volatile double*>(nh_target);
  chi_partial_block(g, om, iom, plan, Sbulk, nsplit, Sres, Spart, partial, sm, true, chain ? chain + CHAIN_QUAD : nullptr,
                    nquad, chain ? chain + CHAIN_RES : nullptr, nres);
  if ((threadIdx.x & 31) == 0) lat_stamp(g, 17);
  // (after the block's waits: the error words of the whole chain are final).  A D-only call is polled by the host on its
  // D slots: there the thread that stores D stores the error words first (fetched here, so that they cost it nothing)
  const bool d_only = !chi0_out && !chi0_low_out && !wave_out;
  int4 ew0 = make_int4(0, 0, 0, 0), ew1 = ew0;
  if (err_dst && iom == 0) {
    if (d_only) {
      if (threadIdx.x == 0) {
        ew0 = __ldcg(reinterpret_cast<const int4*>(err_src));
        ew1 = __ldcg(reinterpret_cast<const int4*>(err_src) + 1);
      }
    } else if (threadIdx.x >= 32 * (CHI_WARPS - 1) && threadIdx.x < 32 * (CHI_WARPS - 1) + 8) {
      err_dst[threadIdx.x & 31] = err_src[threadIdx.x & 31];
    }
  }
  __syncthreads();
  if (nh_target && chain) {
    // the closed-form chi of use_bM species comes from k_nhds on a side branch of the graph, which joins behind this
    // kernel: wait until the (never reset) count of its finished blocks has reached what the host announced for this
    // call.  k_nhds is a serial chain of ~25 us that started with the graph; no other kernel of the chain waits for it, so
    // there is nothing to fall back to -- a count that never arrives is a bug: trap.
    if (threadIdx.x == 0) {
      const double target = nh_tgt;
      const unsigned long long* ctr = reinterpret_cast<const unsigned long long*>(chain + CHAIN_NHDS64);
      unsigned long long seen = 0;
      for (long long spin = 0;; spin++) {
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(ctr) : "memory");
        if ((double)seen >= target) break;
        if (spin > (1LL << 26)) __trap();
      }
    }
    __syncthreads();
  }
  if (d_only) {
    // D only (the single-omega chain): the same operations as assemble_one in a few rolled instructions -- this code runs
    // once, after the chain's last wait, from L2.  Threads 0..5 sum a component of epsilon over the species, thread 0
    // forms the determinant (src/ALPS_fns.f90:598-624).
    if (threadIdx.x < 6) {
      const int c = threadIdx.x;
      cd e = mk(0.0, 0.0);
#pragma unroll 1
      for (int s = 0; s < nspec; s++) {
        cd v = mk(sm.partial[s][2 * c], sm.partial[s][2 * c + 1]);
        if (ext_chi && sm.spc[s].usebM) {
          const double* x = ext_chi + ((size_t)iom * nspec + s) * PARTIAL_PER_SPEC;
          v += mk(__ldcg(x + 2 * c), __ldcg(x + 2 * c + 1));   // (through L2: written while this kernel is resident)
        }
        e += v;
      }
      sm.chinr[c] = e;
    }
    __syncwarp();
    if (threadIdx.x == 0) {
      const ChiGlobals gc = sm.gc;
      const double kperp = gc.kperp, kpar = gc.kpar, vA = gc.vA;
      const cd omc = mk(sm.om[0], sm.om[1]);
      cd enx2, enz2, enxnz;
      if (gc.kperp_norm) {
        enx2 = mk(kperp * kperp, 0.0);
        enz2 = mk(kpar * kpar, 0.0);
        enxnz = mk(kpar * kperp, 0.0);
      } else {
        enx2 = mk(kperp * kperp * kperp * kperp, 0.0);
        enz2 = mk(kpar * kpar * kperp * kperp, 0.0);
        enxnz = mk(kpar * kperp * kperp * kperp, 0.0);
      }
      const cd ov = omc * vA;
      const cd unit = gc.kperp_norm ? ov * ov : (kperp * ov) * (kperp * ov);
      const cd e0 = sm.chinr[0] + unit, e1 = sm.chinr[1] + unit, e2 = sm.chinr[2] + unit;
      const cd w11 = e0 - enz2, w22 = e1 - enz2 - enx2, w33 = e2 - enx2;
      const cd w13 = sm.chinr[4] + enxnz, w12 = sm.chinr[3], w23 = sm.chinr[5];
      const cd d = w11 * (w22 * w33 + w23 * w23) + mk(2.0, 0.0) * w12 * w23 * w13 - w13 * w13 * w22 + w12 * w12 * w33;
      if (err_dst && iom == 0) {
        reinterpret_cast<int4*>(err_dst)[0] = ew0;
        reinterpret_cast<int4*>(err_dst)[1] = ew1;
      }
      // one 16-byte store: in the single-omega chain D is pinned host memory and the host polls it (api.cu)
      if (D) *reinterpret_cast<double2*>(D + 2 * iom) = make_double2(d.x, d.y);
      lat_stamp(g, 19);
    }
    return;
  }
  if (threadIdx.x == 0) {
    unsigned usebM = 0;
    for (int s = 0; s < nspec; s++) usebM |= sm.spc[s].usebM ? (1u << s) : 0u;
    assemble_one(sm.gc, nspec, usebM, mk(sm.om[0], sm.om[1]), iom, &sm.partial[0][0], ext_chi, D, chi0_out, chi0_low_out,
                 wave_out);
    lat_stamp(g, 19);
  }
}

// ------------------------------------------------------------------ launchers
void launch_plan(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, PlanEntry* plan, int* work,
                 int* work_count, cudaStream_t st, double* om_stage, int* plan_flag, int* zero_ints, int nzero) {
  size_t total = (size_t)n_om * gh.NI;
  if (om_stage) {   // fused single-block variant: caller checked plan_fused_ok()
    // (as many warps as there are items: the dependents of the chain start when all of them have passed the trigger)
    const int threads = (int)std::min<size_t>(1024, std::max<size_t>(64, (total + 31) / 32 * 32));
    k_plan<true><<<1, threads, 0, st>>>(g, om, n_om, plan, work, work_count, om_stage, plan_flag, zero_ints, nzero);
    return;
  }
  cudaMemsetAsync(work_count, 0, sizeof(int), st);
  if (!total) return;
  k_plan<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, om, n_om, plan, work, work_count, nullptr, nullptr, nullptr, 0);
}
bool plan_fused_ok(const GlobalDev& gh, int n_om) {
  return n_om >= 1 && n_om <= PLAN_FUSED_MAX_OM && (size_t)n_om * gh.NI <= 1024;
}
bool resonant_lat_class(int n_om, int class_n) { return std::max(class_n, n_om) <= 64; }
int resonant_lat_blocks(int n_om, int gx) { return std::min(148, std::max(1, gx * n_om)) * LAT_PARTS; }
void launch_resonant(const GlobalDev* g, const double* om, int n_om, const PlanEntry* plan, const int* work,
                     const int* work_count, const double* gwin, double* Sres, int* err_flag, double* Spart,
                     cudaStream_t st, int gx, int class_n, int* chain, int nquad) {
  if (n_om <= 0) return;
  if (resonant_lat_class(n_om, class_n) && Spart) {
    // gx block columns per omega loop over the list of resonant harmonics: usually only n = 0 is resonant and a
    // narrow grid saves waves of idle blocks (C1: -2.8 us per D), many resonances (large k_par) want all SMs
    launch_chain(k_resonant_lat, dim3(min(148, max(1, gx * n_om)), LAT_PARTS), dim3(LAT_THREADS), 0, st, g, om, plan,
                 work, work_count, gwin, err_flag, Spart, chain, nquad);
  }
  else
    k_resonant<<<148 * 8, RES_THREADS, 0, st>>>(g, om, plan, work, work_count, gwin, Sres, err_flag);
}
void launch_chi_partial(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, const PlanEntry* plan,
                        const double* Sbulk, int nsplit, const double* Sres, const double* Spart, double* partial,
                        cudaStream_t st) {
  if (n_om <= 0 || gh.nspec <= 0) return;
  k_chi_partial<<<n_om, CHI_THREADS, 0, st>>>(g, om, n_om, plan, Sbulk, nsplit, Sres, Spart, partial);
}
void launch_assemble(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, const double* partial,
                     const double* ext_chi, double* D, double* chi0, double* chi0_low, double* wave,
                     cudaStream_t st, const int* err_src, int* err_dst) {
  (void)gh;
  if (n_om <= 0) return;
  k_assemble<<<(n_om + 127) / 128, 128, 0, st>>>(g, om, n_om, partial, ext_chi, D, chi0, chi0_low, wave, err_src,
                                                 err_dst);
}
void launch_chi_assemble(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, const PlanEntry* plan,
                         const double* Sbulk, int nsplit, const double* Sres, const double* Spart, double* partial,
                         const double* ext_chi, double* D, double* chi0, double* chi0_low, double* wave, cudaStream_t st,
                         const int* err_src, int* err_dst, int* chain, int nquad, int nres, const double* nh_target) {
  if (n_om <= 0) return;
  launch_chain(k_chi_assemble, dim3(n_om), dim3(CHI_THREADS), 0, st, g, om, plan, Sbulk, nsplit, Sres, Spart,
               partial, ext_chi, D, chi0, chi0_low, wave, err_src, err_dst, chain, nquad, nres, nh_target);
}

}  // namespace alps
