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

// FUSED (single-block launches of the single-omega graph, api.cu): the block resets the work counter itself and
// stages the omegas -- read from pinned host memory -- into om_stage for the kernels downstream, so the chain
// needs neither a memset nor a host-to-device copy node.
template <bool FUSED>
__global__ void __launch_bounds__(FUSED ? 1024 : 256)
k_plan(const GlobalDev* __restrict__ gp, const double* __restrict__ om_in, int n_om, PlanEntry* __restrict__ plan,
       int* __restrict__ work, int* __restrict__ work_count, double* __restrict__ om_stage) {
  const GlobalDev& g = *gp;
  __shared__ double s_om[FUSED ? 2 * PLAN_FUSED_MAX_OM : 2];
  const double* om = om_in;
  if (FUSED) {
    pdl_trigger();
    if (threadIdx.x < 2 * n_om) {
      const double v = om_in[threadIdx.x];
      s_om[threadIdx.x] = v;
      om_stage[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) *work_count = 0;
    __syncthreads();
    om = s_om;
  }
  size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_om * g.NI) return;
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
// funct_g does per iperp (src/ALPS_fns.f90:1284-1319).
__device__ inline void funct_g6(const GlobalDev& g, const SpeciesDev& sp, const double* __restrict__ gw, int wbase,
                                double p, Six& out, int* err, int q0 = 0, int nq = 6) {
  const int npar = g.npar;
  const double* ppar = sp.ppar;
  const double dp = sp.dppar_abs;
  int i0 = (int)floor((p - ppar[0]) / dp + 0.5);
  int ic = 0;
  for (int c = min(i0 + 2, npar - 1); c >= max(i0 - 2, 1); c--)
    if (fabs(ppar[c] - p) <= 0.5 * dp) {
      ic = c;
      break;
    }
  if (ic >= npar - 1) ic = npar - 2;
  if (ic <= 1) ic = 2;
  // window slots of nodes ic-1, ic, ic+1 (nodes 1..3 live behind the window, see GlobalDev::WINX)
  int jm = ic - 1 - wbase, j0 = ic - wbase, jp = ic + 1 - wbase;
  if (jm < 0 || jp >= g.WIN) {
    if (ic == 2) {
      jm = g.WIN;
      j0 = g.WIN + 1;
      jp = g.WIN + 2;
    } else {
      *err = 1 + (ic & 0xffff) + ((i0 & 0x7fff) << 16);
      six_zero(out);
      return;
    }
  }
  const double x = p - ppar[ic];
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
      const double* w = gw + ((size_t)jm * 3 + xt) * 2;
      double pw = m == 0 ? 1.0 : (m == 1 ? ppar[ic - 1] : ppar[ic - 1] * ppar[ic - 1]);
      gm = -(mk(w[0], w[1]) * pw) / g.kpar;
    }
    {
      const double* w = gw + ((size_t)j0 * 3 + xt) * 2;
      double pw = m == 0 ? 1.0 : (m == 1 ? ppar[ic] : ppar[ic] * ppar[ic]);
      g0 = -(mk(w[0], w[1]) * pw) / g.kpar;
    }
    {
      const double* w = gw + ((size_t)jp * 3 + xt) * 2;
      double pw = m == 0 ? 1.0 : (m == 1 ? ppar[ic + 1] : ppar[ic + 1] * ppar[ic + 1]);
      gp = -(mk(w[0], w[1]) * pw) / g.kpar;
    }
    out.v[q] = g0 + (0.5 * ((gp - gm) / dp)) * x;
  }
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
// Each block leaves 6 complex partial sums (+ zero / error flags); the last block of an item to finish
// (ticket counter) combines them in a fixed order.  Same arithmetic per term as k_resonant; only the
// (fixed) summation order differs.
constexpr int LAT_THREADS = 256;
constexpr int LAT_PARTS = 11;
constexpr int LAT_STRIDE = 16;   // doubles per partial row: 12 sums, zero flag, error code
static_assert(LAT_PARTS * LAT_STRIDE == RES_PART_DOUBLES, "partial-row buffer size");
__global__ void __launch_bounds__(LAT_THREADS) k_resonant_lat(const GlobalDev* __restrict__ gp,
                                                              const double* __restrict__ om,
                                                              const PlanEntry* __restrict__ plan,
                                                              const int* __restrict__ work,
                                                              const int* __restrict__ work_count,
                                                              const double* __restrict__ gwin, double* __restrict__ Sres,
                                                              int* __restrict__ err_flag, double* __restrict__ Spart,
                                                              int* __restrict__ tickets) {
  const GlobalDev& g = *gp;
  const int tid = threadIdx.x, wlane = tid & 31, wid = tid >> 5, part = blockIdx.y;
  __shared__ cd s_part[LAT_THREADS / 32][6];
  __shared__ cd s_f[3][2];   // analytic branch: g(p_R + dp), g(p_R - dp), g(p_R) for this block's two combinations
  __shared__ int s_err, s_last;
  pdl_trigger();
  pdl_wait();
  const int nwork = *work_count;
  if (blockIdx.x == 0 && part == 0 && tid == 0) err_flag[7] = nwork;   // feedback for the host's choice of grid width
  for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
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
    if (tid == 0) s_err = 0;
    __syncthreads();
    Six acc;
    six_zero(acc);
    int err = 0, zero = 0;

    if ((pe.flags & PLAN_NEAR) && part <= 8) {
      const double* gw = gwin + idx * (size_t)g.WINX * 6;
      const int wbase = pe.ipar_res - M_I - 2;
      const double dppar = sp.dppar_signed;
      const double capDelta = pR - ppar[pe.ipar_res - M_I];
      const double smdelta = capDelta / (1.0 * M_P);
      const bool pairing = fabs(pI) > g.Tlim;
      const int piece = part / 3, q0 = 2 * (part % 3);   // this block: combinations q0, q0 + 1
      if (pairing && piece <= 1) {
        // Eq. (3.5): symmetric pairing around the pole, src/ALPS_fns.f90:1026-1082
        for (int j = tid; j <= M_P; j += LAT_THREADS) {
          const double wj = (j == 0 || j == M_P) ? 1.0 : 2.0;
          const double p = (j == 0) ? pR : (j == M_P ? pR + capDelta : pR + smdelta * j);
          Six f;
          if (piece == 0) {
            funct_g6(g, sp, gw, wbase, p, f, &err, q0, 2);
            const cd d1 = mk(p - pR, -pI);
#pragma unroll
            for (int q = 0; q < 6; q++)
              if (q >= q0 && q < q0 + 2) acc.v[q] += wj * (f.v[q] / d1);
          } else {
            funct_g6(g, sp, gw, wbase, 2.0 * pR - p, f, &err, q0, 2);
            const cd d2 = mk(p - pR, pI);
#pragma unroll
            for (int q = 0; q < 6; q++)
              if (q >= q0 && q < q0 + 2) acc.v[q] -= wj * (f.v[q] / d2);
          }
        }
      } else if (!pairing && piece == 0) {
        // Eq. (3.6): linearised integrand + pole term, src/ALPS_fns.f90:1088-1165; the three g values
        // are evaluated once, by three different warps
        if (wlane == 0 && wid < 3) {
          Six f;
          funct_g6(g, sp, gw, wbase, wid == 0 ? pR + dppar : (wid == 1 ? pR - dppar : pR), f, &err, q0, 2);
#pragma unroll
          for (int q = 0; q < 6; q++)
            if (q >= q0 && q < q0 + 2) s_f[wid][q - q0] = f.v[q];
        }
        __syncthreads();
        cd gprime[2];
#pragma unroll
        for (int e = 0; e < 2; e++) gprime[e] = (s_f[0][e] - s_f[1][e]) / (2.0 * dppar);
        for (int j = 1 + tid; j <= M_P; j += LAT_THREADS) {
          const double wj = (j == M_P) ? 1.0 : 2.0;
          const double p = (j == M_P) ? pR + capDelta : pR + smdelta * j;
          const double x2 = (p - pR) * (p - pR);
          const double lor = x2 / (x2 + pI * pI);
#pragma unroll
          for (int q = 0; q < 6; q++)
            if (q >= q0 && q < q0 + 2) acc.v[q] += (wj * 2.0) * gprime[q - q0] * lor;
        }
        if (tid == 0 && pI != 0.0) {
          const double sgn = pI > 0.0 ? 1.0 : -1.0;
#pragma unroll
          for (int q = 0; q < 6; q++)
            if (q >= q0 && q < q0 + 2) acc.v[q] += sgn * (cmul_i((2.0 * PI) * s_f[2][q - q0]) / smdelta);
        }
      }
      if (piece == 2) {
        // tiny rest between p_res + capDelta and the first regular node, :1168-1230
        const double rest = ppar[pe.upperlimit] - pR - capDelta;
        const int ntiny = (int)(rest / smdelta);
        if (ntiny > 0) {
          const double correction = (rest / (1.0 * ntiny)) / smdelta;
          for (int j = tid; j <= ntiny; j += LAT_THREADS) {
            const double wj = (j == 0 || j == ntiny) ? 1.0 : 2.0;
            const double p = (j == 0) ? pR + capDelta : pR + capDelta + correction * smdelta * j;
            Six f1;
            funct_g6(g, sp, gw, wbase, p, f1, &err, q0, 2);
            const cd d1 = mk(p - pR, -pI);
#pragma unroll
            for (int q = 0; q < 6; q++)
              if (q >= q0 && q < q0 + 2) acc.v[q] += (wj * correction) * (f1.v[q] / d1);
          }
        }
      }
    }

    if ((pe.flags & PLAN_LANDAU) && part >= 9) {
      // landau_integrate, src/ALPS_fns.f90:1327-1452; acc.v[0..2] collect La, Lb, Lc
      const double dpperp = sp.dpperp, dppar = sp.dppar_abs;
      const double* Jn = sp.J + (size_t)(nabs + 1) * sp.ldj;
      const double* Jm = sp.J + (size_t)nabs * sp.ldj;
      const double* Jp = sp.J + (size_t)(nabs + 2) * sp.ldj;
      const cd ppl = mk(pR + dppar, pI), pmi = mk(pR - dppar, pI);
      const int gidx = tid >> 2, sub = tid & 3, lbase = wlane & ~3;
      for (int r0 = 64 * (part - 9); r0 <= nperp; r0 += 128) {
        const int r = r0 + gidx;
        const bool valid = r <= nperp;
        const bool interior = r >= 1 && r <= nperp - 1;
        const int hi = (r == 0) ? 1 : (r == nperp ? nperp : r + 1);
        const int lo = (r == 0) ? 0 : (r == nperp ? nperp - 1 : r - 1);
        cd v = mk(0.0, 0.0);
        if (valid) {
          if (sub == 0) v = eval_fit(g, s, r, ppl);
          else if (sub == 1) v = eval_fit(g, s, r, pmi);
          else if (sub == 2) v = eval_fit(g, s, hi, p_res);
          else v = eval_fit(g, s, lo, p_res);
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
          acc.v[0] += (h * (bj * bj)) * Q;
          acc.v[1] += (h * (pperp * (bj * bp))) * Q;
          acc.v[2] += (h * ((pperp * pperp) * (bp * bp))) * Q;
        }
      }
    }

    // ---- block sum -> partial row of this part
#pragma unroll
    for (int q = 0; q < 6; q++) {
      const cd t = warp_sum_c(acc.v[q]);
      if (wlane == 0) s_part[wid][q] = t;
    }
    if (err) atomicMax(&s_err, err);
    zero = __syncthreads_or(zero);
    double* prow = Spart + (idx * LAT_PARTS + part) * LAT_STRIDE;
    if (tid < 6) {
      cd t = s_part[0][tid];
      for (int w = 1; w < LAT_THREADS / 32; w++) t += s_part[w][tid];
      prow[2 * tid] = t.x;
      prow[2 * tid + 1] = t.y;
    }
    if (tid == 6) {
      prow[12] = zero ? 1.0 : 0.0;
      prow[13] = (double)s_err;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&tickets[idx], 1) == LAT_PARTS - 1);
    __syncthreads();
    if (s_last && tid == 0) {
      __threadfence();
      tickets[idx] = 0;   // ready for the next launch
      const double* pr = Spart + idx * LAT_PARTS * LAT_STRIDE;
      auto ld = [&](int prt, int q) { return mk(__ldcg(pr + prt * LAT_STRIDE + 2 * q), __ldcg(pr + prt * LAT_STRIDE + 2 * q + 1)); };
      Six tot;
      six_zero(tot);
      int e = 0;
      for (int prt = 0; prt < LAT_PARTS; prt++) e = max(e, (int)__ldcg(pr + prt * LAT_STRIDE + 13));
      if (pe.flags & PLAN_NEAR) {
        const double capDelta = pR - ppar[pe.ipar_res - M_I];
        const double smdelta = capDelta / (1.0 * M_P);
        const double near_fac = 2.0 * PI * smdelta * sp.dpperp * 0.25;
#pragma unroll
        for (int q = 0; q < 6; q++) tot.v[q] += near_fac * ((ld(q / 2, q) + ld(3 + q / 2, q)) + ld(6 + q / 2, q));
      }
      const bool zr = __ldcg(pr + 9 * LAT_STRIDE + 12) != 0.0 || __ldcg(pr + 10 * LAT_STRIDE + 12) != 0.0;
      if ((pe.flags & PLAN_LANDAU) && !zr) {
        // landau = -(sum) * i * dpperp * pi * 2 pi ; factor 2 (Im om < 0) or 1 (Im om == 0),
        // full_integrate src/ALPS_fns.f90:782-789
        const double mult = (omc.y < 0.0 ? 2.0 : 1.0) * sp.dpperp * PI * 2.0 * PI;
        const cd ca = cmul_i(-(ld(9, 0) + ld(10, 0))) * mult, cb = cmul_i(-(ld(9, 1) + ld(10, 1))) * mult,
                 cc = cmul_i(-(ld(9, 2) + ld(10, 2))) * mult;
        tot.v[0] += ca;
        tot.v[1] += p_res * ca;
        tot.v[2] += (p_res * p_res) * ca;
        tot.v[3] += cb;
        tot.v[4] += p_res * cb;
        tot.v[5] += cc;
      }
      double* o = Sres + idx * 12;
#pragma unroll
      for (int q = 0; q < 6; q++) {
        o[2 * q] = tot.v[q].x;
        o[2 * q + 1] = tot.v[q].y;
      }
      if (e) {
        err_flag[0] = 1;
        err_flag[1] = (int)idx;
        err_flag[2] = pe.ipar_res;
        err_flag[3] = pe.upperlimit;
        err_flag[4] = pe.flags;
        err_flag[5] = e;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- chi partial
// One warp per (omega, species): tensor components of every (n, sign) from its six moment
// sums, summed over the harmonics of this process' shard.
// partial[(iom*nspec + s)*PARTIAL_PER_SPEC + 2*c .. ]: c = mode-1 (0..5) for chi,
// c = 6 + 3*(mode-1) + (m+1) for chi_low(mode, m), m = -1,0,1.
// one warp: the partial row of (omega iom, species s)
__device__ __forceinline__ void chi_partial_warp(const GlobalDev& g, int iom, int s, int lane,
                                                 const PlanEntry* __restrict__ plan, const double* __restrict__ Sbulk,
                                                 int nsplit, const double* __restrict__ Sres, double* partial) {
  const int w = iom * g.nspec + s;
  const SpeciesDev& sp = g.sp[s];
  cd chi[6], low[6][3];
#pragma unroll
  for (int c = 0; c < 6; c++) {
    chi[c] = mk(0.0, 0.0);
    low[c][0] = low[c][1] = low[c][2] = mk(0.0, 0.0);
  }
  const bool table = !sp.usebM;
  if (table) {
    const double z = g.kperp_norm ? g.kperp / sp.qs : 1.0 / sp.qs;
    const double kf1 = g.kperp_norm ? 1.0 : g.kperp, kf2 = g.kperp_norm ? 1.0 : g.kperp * g.kperp;
    const double cbulk = 2.0 * PI * sp.dpperp * sp.dppar_abs * 0.25;
    const int nitems = 2 * (sp.nhi + 1);
    for (int r = lane; r < nitems; r += 32) {
      const size_t idx = (size_t)iom * g.NI + sp.item_base + r;
      const PlanEntry pe = plan[idx];
      if (!(pe.flags & PLAN_ACTIVE)) continue;
      const int nabs = r >> 1, sg = r & 1;
      const double nn = sg ? -(double)nabs : (double)nabs;
      cd mode[6];
      if (pe.flags & PLAN_REL) {
        // relativistic species: k_rel already produced the six tensor components
        const double* sr = Sres + idx * 12;
#pragma unroll
        for (int q = 0; q < 6; q++) mode[q] = mk(sr[2 * q], sr[2 * q + 1]);
      } else {
        cd S[6];
#pragma unroll
        for (int q = 0; q < 6; q++) S[q] = mk(0.0, 0.0);
        // partial rows of the p_par splits of k_quad, added in order; four rows are fetched before they are
        // added so that a row costs one L2 round trip per four instead of one each (latency chain)
        for (int j0 = 0; j0 < nsplit; j0 += 4) {
          double2 row[4][6];
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (j0 + u < nsplit) {
              const double2* sb = reinterpret_cast<const double2*>(Sbulk + (idx * nsplit + j0 + u) * 12);
#pragma unroll
              for (int q = 0; q < 6; q++) row[u][q] = sb[q];
            }
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (j0 + u < nsplit) {
#pragma unroll
              for (int q = 0; q < 6; q++) S[q] += mk(row[u][q].x, row[u][q].y);
            }
        }
#pragma unroll
        for (int q = 0; q < 6; q++) S[q] = cbulk * S[q];
        if (pe.flags & (PLAN_NEAR | PLAN_LANDAU)) {
          const double* sr = Sres + idx * 12;
#pragma unroll
          for (int q = 0; q < 6; q++) S[q] += mk(sr[2 * q], sr[2 * q + 1]);
        }
        mode[0] = ((nn * nn) / (z * z)) * S[0];          // xx: n^2 J^2 / z^2
        mode[1] = kf2 * S[5];                            // yy: p_perp^2 J'^2
        mode[2] = kf2 * S[2];                            // zz: J^2 p_par^2
        mode[3] = cmul_i((kf1 * nn / z) * S[3]);         // xy: i p_perp n J J' / z
        mode[4] = (kf1 * nn / z) * S[1];                 // xz: n J^2 p_par / z
        mode[5] = -cmul_i(kf2 * S[4]);                   // yz: -i J J' p_par p_perp
      }
      if (nabs == 0) {
        // n = 0: only yy, zz, yz are evaluated (src/ALPS_fns.f90:368-393)
        chi[1] += mode[1]; chi[2] += mode[2]; chi[5] += mode[5];
        low[1][1] = mode[1]; low[2][1] = mode[2]; low[5][1] = mode[5];
      } else {
#pragma unroll
        for (int c = 0; c < 6; c++) chi[c] += mode[c];
        if (nabs == 1) {
#pragma unroll
          for (int c = 0; c < 6; c++) low[c][sg ? 0 : 2] = mode[c];
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 6; c++) {
    chi[c] = warp_sum_c(chi[c]);
#pragma unroll
    for (int m = 0; m < 3; m++) low[c][m] = warp_sum_c(low[c][m]);
  }
  if (lane == 0) {
    if (table && sp.nlo_shard == 0) {
      const double ee = g.kperp_norm ? sp.int_ee : g.kperp * g.kperp * sp.int_ee;
      chi[2].x += ee;
      if (!sp.relativistic) low[2][1].x += ee;   // int_ee_rel goes into chi only (src/ALPS_fns.f90:481-486)
    }
    const double norm = sp.ns * sp.qs;
    double* o = partial + (size_t)w * PARTIAL_PER_SPEC;
#pragma unroll
    for (int c = 0; c < 6; c++) {
      o[2 * c] = norm * chi[c].x;
      o[2 * c + 1] = norm * chi[c].y;
#pragma unroll
      for (int m = 0; m < 3; m++) {
        o[2 * (6 + 3 * c + m)] = norm * low[c][m].x;
        o[2 * (6 + 3 * c + m) + 1] = norm * low[c][m].y;
      }
    }
  }
}

__global__ void __launch_bounds__(128) k_chi_partial(const GlobalDev* __restrict__ gp, const double* __restrict__ om,
                                                     int n_om, const PlanEntry* __restrict__ plan,
                                                     const double* __restrict__ Sbulk, int nsplit,
                                                     const double* __restrict__ Sres, double* __restrict__ partial) {
  (void)om;
  const GlobalDev& g = *gp;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_om * g.nspec) return;
  chi_partial_warp(g, w / g.nspec, w % g.nspec, threadIdx.x & 31, plan, Sbulk, nsplit, Sres, partial);
}

// -------------------------------------------------------------------- assemble
// One thread per omega: src/ALPS_fns.f90:536-624.
__device__ __forceinline__ void assemble_one(const GlobalDev& g, const double* __restrict__ om, int iom,
                                             const double* partial, const double* __restrict__ ext_chi,
                                             double* __restrict__ D, double* __restrict__ chi0_out,
                                             double* __restrict__ chi0_low_out, double* __restrict__ wave_out) {
  const int nspec = g.nspec;
  const cd omc = mk(om[2 * iom], om[2 * iom + 1]);
  const double kperp = g.kperp, kpar = g.kpar, vA = g.vA;
  cd enx2, enz2, enxnz, norm2;
  if (g.kperp_norm) {
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
    const double* p = partial + ((size_t)iom * nspec + s) * PARTIAL_PER_SPEC;
    const double* x = ext_chi ? ext_chi + ((size_t)iom * nspec + s) * PARTIAL_PER_SPEC : nullptr;
    const bool useext = x && g.sp[s].usebM;
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
  const cd unit = g.kperp_norm ? ov * ov : (kperp * ov) * (kperp * ov);
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
  assemble_one(*gp, om, iom, partial, ext_chi, D, chi0_out, chi0_low_out, wave_out);
}

// k_chi_partial + k_assemble of a small batch in one launch (the single-omega graph and batched roots): block =
// one omega, warp s = species s; the partial rows go through global memory exactly as between the two kernels
// (same code, same order of operations: bitwise the two-kernel result), thread 0 assembles after the barrier.
__global__ void __launch_bounds__(32 * MAXSPEC)
k_chi_assemble(const GlobalDev* __restrict__ gp, const double* __restrict__ om, const PlanEntry* __restrict__ plan,
               const double* __restrict__ Sbulk, int nsplit, const double* __restrict__ Sres, double* partial,
               const double* __restrict__ ext_chi, double* __restrict__ D, double* __restrict__ chi0_out,
               double* __restrict__ chi0_low_out, double* __restrict__ wave_out, const int* __restrict__ err_src,
               int* __restrict__ err_dst) {
  const GlobalDev& g = *gp;
  const int iom = blockIdx.x, s = threadIdx.x >> 5;
  pdl_trigger();
  pdl_wait();
  if (s < g.nspec) chi_partial_warp(g, iom, s, threadIdx.x & 31, plan, Sbulk, nsplit, Sres, partial);
  if (err_dst && iom == 0 && threadIdx.x >= 32 && threadIdx.x < 40) err_dst[threadIdx.x - 32] = err_src[threadIdx.x - 32];
  __syncthreads();
  if (threadIdx.x == 0) assemble_one(g, om, iom, partial, ext_chi, D, chi0_out, chi0_low_out, wave_out);
}

// ------------------------------------------------------------------ launchers
void launch_plan(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, PlanEntry* plan, int* work,
                 int* work_count, cudaStream_t st, double* om_stage) {
  size_t total = (size_t)n_om * gh.NI;
  if (om_stage) {   // fused single-block variant: caller checked plan_fused_ok()
    k_plan<true><<<1, 1024, 0, st>>>(g, om, n_om, plan, work, work_count, om_stage);
    return;
  }
  cudaMemsetAsync(work_count, 0, sizeof(int), st);
  if (!total) return;
  k_plan<false><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, om, n_om, plan, work, work_count, nullptr);
}
bool plan_fused_ok(const GlobalDev& gh, int n_om) {
  return n_om >= 1 && n_om <= PLAN_FUSED_MAX_OM && (size_t)n_om * gh.NI <= 1024;
}
void launch_resonant(const GlobalDev* g, const double* om, int n_om, const PlanEntry* plan, const int* work,
                     const int* work_count, const double* gwin, double* Sres, int* err_flag, double* Spart,
                     int* tickets, cudaStream_t st, int gx, int class_n) {
  if (n_om <= 0) return;
  if (std::max(class_n, n_om) <= 64 && Spart && tickets) {
    // gx block columns per omega loop over the list of resonant harmonics: usually only n = 0 is resonant and a
    // narrow grid saves waves of idle blocks (C1: -2.8 us per D), many resonances (large k_par) want all SMs
    launch_chain(k_resonant_lat, dim3(min(148, max(1, gx * n_om)), LAT_PARTS), dim3(LAT_THREADS), 0, st, g, om, plan,
                 work, work_count, gwin, Sres, err_flag, Spart, tickets);
  }
  else
    k_resonant<<<148 * 8, RES_THREADS, 0, st>>>(g, om, plan, work, work_count, gwin, Sres, err_flag);
}
void launch_chi_partial(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, const PlanEntry* plan,
                        const double* Sbulk, int nsplit, const double* Sres, double* partial, cudaStream_t st) {
  int warps = n_om * gh.nspec;
  if (warps <= 0) return;
  k_chi_partial<<<(warps + 3) / 4, 128, 0, st>>>(g, om, n_om, plan, Sbulk, nsplit, Sres, partial);
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
                         const double* Sbulk, int nsplit, const double* Sres, double* partial, const double* ext_chi,
                         double* D, double* chi0, double* chi0_low, double* wave, cudaStream_t st, const int* err_src,
                         int* err_dst) {
  if (n_om <= 0) return;
  launch_chain(k_chi_assemble, dim3(n_om), dim3(32 * ((gh.nspec + 1) & ~1)), 0, st, g, om, plan, Sbulk, nsplit, Sres,
               partial, ext_chi, D, chi0, chi0_low, wave, err_src, err_dst);
}

}  // namespace alps
