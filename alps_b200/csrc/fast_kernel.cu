// alps_b200: the "map fast path" (alps_b200_set_mode(1)).
//
// For non-relativistic species the p_perp sum of the regular quadrature does not depend on omega:
//   G_x(n,ipar) = sum_iperp W_x(n,iperp) (om A' + C')(iperp,ipar) = om * GA_x(n,ipar) + GB_x(n,ipar)
// with real tables GA, GB that depend on k only ("algebraic shortcut").  set_k
// builds them once with the same TMA/FP64 pipeline as k_quad (STORE variant); this kernel then does
// the remaining O(nmax * npar) work per omega: the resonance denominators for +n and -n, the p_par
// trapezoid weights of the resonance plan, the p_par moments -- the same arithmetic as k_quad's
// epilogue -- and leaves the same six moment sums (and resonance windows) for k_resonant /
// k_chi_partial.  One thread per omega, one block column per (species, |n|): the table entries are
// warp-uniform loads, no reduction is needed.
#include "common.cuh"
#include "kernels.h"

namespace alps {

__global__ void __launch_bounds__(128) k_fast(const GlobalDev* __restrict__ gp, const double* __restrict__ om, int n_om,
                                              const FastItem* __restrict__ items, const PlanEntry* __restrict__ plan,
                                              double* __restrict__ Sbulk, double* __restrict__ gwin) {
  const GlobalDev& g = *gp;
  const int iom = blockIdx.x * blockDim.x + threadIdx.x;
  if (iom >= n_om) return;
  const FastItem it = items[blockIdx.y];
  const SpeciesDev& sp = g.sp[it.s];
  const int nabs = it.nabs, npar = g.npar, WIN = g.WIN, WINX = g.WINX, M_I = g.M_I;
  const double omr = om[2 * iom], omi = om[2 * iom + 1];
  const double qs = sp.qs, ms = sp.ms, kpar = g.kpar;
  const size_t item0 = (size_t)iom * g.NI + sp.item_base + 2 * nabs;
  PlanEntry pe[2];
  pe[0] = plan[item0];
  pe[1] = plan[item0 + 1];
  const bool act0 = (pe[0].flags & PLAN_ACTIVE) != 0, act1 = nabs > 0 && (pe[1].flags & PLAN_ACTIVE) != 0;
  if (!act0 && !act1) return;
  const double* __restrict__ G = sp.G + (size_t)nabs * (npar - 1) * 6;
  const double* __restrict__ ppar = sp.ppar;
  double S[2][12];
#pragma unroll
  for (int q = 0; q < 12; q++) S[0][q] = S[1][q] = 0.0;
  const double dre = ms * omr, dim = ms * omi, dim2 = dim * dim;
  const double nq = (double)nabs * qs;
  // non-resonant harmonics (the vast majority) use the plain trapezoid weights 1,2,...,2,1
  const bool plain0 = act0 && !(pe[0].flags & PLAN_RES), plain1 = act1 && !(pe[1].flags & PLAN_RES);
  const bool simple = (plain0 || !act0) && (plain1 || !act1);
  for (int ipar = 1; ipar <= npar - 1; ipar++) {
    const double2 t0 = __ldg(reinterpret_cast<const double2*>(G + (size_t)(ipar - 1) * 6));
    const double2 t1 = __ldg(reinterpret_cast<const double2*>(G + (size_t)(ipar - 1) * 6 + 2));
    const double2 t2 = __ldg(reinterpret_cast<const double2*>(G + (size_t)(ipar - 1) * 6 + 4));
    // G = om * GA + GB  (x = GA, y = GB)
    const cd Ga = mk(fma(omr, t0.x, t0.y), omi * t0.x);
    const cd Gb = mk(fma(omr, t1.x, t1.y), omi * t1.x);
    const cd Gc = mk(fma(omr, t2.x, t2.y), omi * t2.x);
    const double p = __ldg(ppar + ipar), p2 = p * p;
    const double x = dre - kpar * p;            // Re(den) before the -+ n qs shift
    const double drp = x - nq, drm = x + nq;    // den = ms om - kpar p_par -+ n qs (resU, src/ALPS_fns.f90:1591)
    const double dp = drp * drp + dim2, dm = drm * drm + dim2;
    double wp, wm;
    if (simple) {
      const double we = (ipar == 1 || ipar == npar - 1) ? 1.0 : 2.0;
      wp = act0 ? we : 0.0;
      wm = act1 ? we : 0.0;
    } else {
      wp = act0 ? range_w(ipar, pe[0].lo1, pe[0].hi1) + range_w(ipar, pe[0].lo2, pe[0].hi2) : 0.0;
      wm = act1 ? range_w(ipar, pe[1].lo1, pe[1].hi1) + range_w(ipar, pe[1].lo2, pe[1].hi2) : 0.0;
    }
    // both reciprocals from one: 1/dp = dm/(dp dm), 1/dm = dp/(dp dm)
    // a sign with zero weight must not poison the shared reciprocal (real omega exactly on a node)
    const double dps = wp != 0.0 ? dp : 1.0, dms = wm != 0.0 ? dm : 1.0;
    const double inv = fast_rcp(dps * dms);
    const double tp = wp * dms * inv, tm = wm * dps * inv;
    {
      const cd R = mk(drp * tp, -dim * tp);
      const cd Va = R * Ga, Vb = R * Gb, Vc = R * Gc;
      S[0][0] += Va.x;       S[0][1] += Va.y;
      S[0][2] += p * Va.x;   S[0][3] += p * Va.y;
      S[0][4] += p2 * Va.x;  S[0][5] += p2 * Va.y;
      S[0][6] += Vb.x;       S[0][7] += Vb.y;
      S[0][8] += p * Vb.x;   S[0][9] += p * Vb.y;
      S[0][10] += Vc.x;      S[0][11] += Vc.y;
    }
    {
      const cd R = mk(drm * tm, -dim * tm);
      const cd Va = R * Ga, Vb = R * Gb, Vc = R * Gc;
      S[1][0] += Va.x;       S[1][1] += Va.y;
      S[1][2] += p * Va.x;   S[1][3] += p * Va.y;
      S[1][4] += p2 * Va.x;  S[1][5] += p2 * Va.y;
      S[1][6] += Vb.x;       S[1][7] += Vb.y;
      S[1][8] += p * Vb.x;   S[1][9] += p * Vb.y;
      S[1][10] += Vc.x;      S[1][11] += Vc.y;
    }
    if (!simple) {
#pragma unroll
      for (int sg = 0; sg < 2; sg++) {
        if ((sg == 0 ? act0 : act1) && (pe[sg].flags & PLAN_NEAR)) {
          int j = ipar - (pe[sg].ipar_res - M_I - 2);
          if (j < 0 || j >= WIN) j = (ipar <= 3) ? WIN + ipar - 1 : -1;
          if (j >= 0) {
            double* gw = gwin + ((item0 + sg) * WINX + j) * 6;
            gw[0] = Ga.x; gw[1] = Ga.y; gw[2] = Gb.x; gw[3] = Gb.y; gw[4] = Gc.x; gw[5] = Gc.y;
          }
        }
      }
    }
  }
#pragma unroll
  for (int sg = 0; sg < 2; sg++) {
    if (sg == 0 ? !act0 : !act1) continue;
    double* o = Sbulk + (item0 + sg) * 12;
#pragma unroll
    for (int q = 0; q < 12; q++) o[q] = S[sg][q];
  }
}

void launch_fast(const GlobalDev* g, const double* om, int n_om, const FastItem* items, int nitems,
                 const PlanEntry* plan, double* Sbulk, double* gwin, cudaStream_t st) {
  if (n_om <= 0 || nitems <= 0) return;
  dim3 grid((n_om + 127) / 128, nitems);
  k_fast<<<grid, 128, 0, st>>>(g, om, n_om, items, plan, Sbulk, gwin);
}

}  // namespace alps
