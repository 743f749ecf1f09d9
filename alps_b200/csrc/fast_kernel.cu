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
#include "pipe.cuh"

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

// ------------------------------------------------------------------------------------------------------------------
// k_fast_tiled (default, ALPS_B200_FAST_VARIANT=1): the same sums from omega-independent REAL moment tables.
//
//   S_j(n,+-) = sum_ipar w(ipar) p^m (om GA_x + GB_x)(ipar) / den_+-(ipar)
//             = om * sum_ipar R_+-(ipar) TA_j(ipar)  +  sum_ipar R_+-(ipar) TB_j(ipar),      R = 1/den (complex),
//   TA_j = w_tab p^m GA_x,  TB_j = w_tab p^m GB_x  for the six (x, m) pairs (a,0) (a,1) (a,2) (b,0) (b,1) (c,0) and the
//   plain trapezoid weights w_tab = 1,2,...,2,1 (k_fast_tables builds T[n][ipar-1][12] once per k from the STORE tables).
// Per (|n|, ipar, omega) that is one shared reciprocal for both signs and 12 x 4 real FMAs -- 64 FP64 instructions
// against 87 in k_fast (no per-point G = om GA + GB, no complex R*G products); om is applied once per harmonic at the end.
// One CTA = 128 omegas (one per thread) x one (species, |n|): the 96 B/node table row streams through a 3-stage
// shared-memory ring filled by cp.async.bulk (SASS UBLKCP) with mbarrier completion, p_par of the species is staged once;
// all operand reads are warp-uniform LDS.128 broadcasts.  Resonant harmonics (a few omegas of a few harmonics) take a
// second loop with per-thread trapezoid factors w_eff/w_tab in {0, 1/2, 1, 2} -- exact scalings, so both loops add the
// same terms -- and leave the resonance windows for k_resonant like k_fast does.
constexpr int FT_CH = 128, FT_W = 12, FT_STAGES = 3, FT_THREADS = 128;

__global__ void __launch_bounds__(256) k_fast_tables(const double* __restrict__ G, const double* __restrict__ ppar,
                                                     int npar, int nrows, double* __restrict__ T) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // (n, ipar-1)
  const size_t tot = (size_t)nrows * (npar - 1);
  if (i >= tot) return;
  const int ipar = (int)(i % (size_t)(npar - 1)) + 1;
  const double w = (ipar == 1 || ipar == npar - 1) ? 1.0 : 2.0;
  const double p = ppar[ipar], p2 = p * p;
  const double* g = G + i * 6;     // GAa GBa GAb GBb GAc GBc
  double* t = T + i * FT_W;
#pragma unroll
  for (int ab = 0; ab < 2; ab++) {
    const double ga = w * g[0 + ab], gb = w * g[2 + ab], gc = w * g[4 + ab];
    t[6 * ab + 0] = ga;
    t[6 * ab + 1] = p * ga;
    t[6 * ab + 2] = p2 * ga;
    t[6 * ab + 3] = gb;
    t[6 * ab + 4] = p * gb;
    t[6 * ab + 5] = gc;
  }
}

void launch_fast_tables(const double* G, const double* ppar, int npar, int nrows, double* T, cudaStream_t st) {
  const size_t tot = (size_t)nrows * (npar - 1);
  if (!tot) return;
  k_fast_tables<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(G, ppar, npar, nrows, T);
}

__global__ void __launch_bounds__(FT_THREADS, 3)
    k_fast_tiled(const GlobalDev* __restrict__ gp, const double* __restrict__ om, int n_om,
                 const FastItem* __restrict__ items, const PlanEntry* __restrict__ plan, double* __restrict__ Sbulk,
                 double* __restrict__ gwin, const double kpar) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sT = reinterpret_cast<double*>(smem_raw);                    // [FT_STAGES][FT_CH * FT_W]
  double* sp = sT + FT_STAGES * FT_CH * FT_W;                          // p_par(1 .. npar-1), padded to FT_CH
  __shared__ unsigned long long full[FT_STAGES];
  const GlobalDev& g = *gp;
  const FastItem it = items[blockIdx.y];
  const SpeciesDev& sp_ = g.sp[it.s];
  const int nabs = it.nabs, npar = g.npar, np1 = npar - 1;
  const int nch = (np1 + FT_CH - 1) / FT_CH;
  const double* __restrict__ Trow = sp_.T + (size_t)nabs * np1 * FT_W;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < FT_STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = tid; j < nch * FT_CH; j += FT_THREADS) sp[j] = j < np1 ? sp_.ppar[j + 1] : 0.0;
  __syncthreads();
  auto issue = [&](int c) {      // thread 0: chunk c -> stage c % FT_STAGES
    const int s = c % FT_STAGES, cnt = min(FT_CH, np1 - c * FT_CH);
    const uint32_t bytes = (uint32_t)cnt * FT_W * sizeof(double);
    mbar_expect_tx(&full[s], bytes);
    tma_load_1d(sT + (size_t)s * FT_CH * FT_W, Trow + (size_t)c * FT_CH * FT_W, bytes, &full[s]);
  };
  if (tid == 0)
    for (int c = 0; c < FT_STAGES && c < nch; c++) issue(c);

  const int iom = blockIdx.x * FT_THREADS + tid;
  const bool live = iom < n_om;
  const double omr = live ? om[2 * iom] : 1.0, omi = live ? om[2 * iom + 1] : 1.0;
  const double qs = sp_.qs, ms = sp_.ms;
  const size_t item0 = (size_t)(live ? iom : 0) * g.NI + sp_.item_base + 2 * nabs;
  PlanEntry pe[2];
  pe[0] = plan[item0];
  pe[1] = plan[item0 + 1];
  const bool act0 = live && (pe[0].flags & PLAN_ACTIVE) != 0, act1 = live && nabs > 0 && (pe[1].flags & PLAN_ACTIVE) != 0;
  const bool plain0 = !act0 || !(pe[0].flags & PLAN_RES), plain1 = !act1 || !(pe[1].flags & PLAN_RES);
  // warp-uniform choice of the loop: every lane non-resonant -> table weights as they are
  const bool warp_simple = __all_sync(0xffffffffu, plain0 && plain1);
  const double dre = ms * omr, dim = ms * omi, dim2 = dim * dim;
  const double nq = (double)nabs * qs;
  double PR[FT_W], PI[FT_W], MR[FT_W], MI[FT_W];
#pragma unroll
  for (int q = 0; q < FT_W; q++) PR[q] = PI[q] = MR[q] = MI[q] = 0.0;
  const int WIN = g.WIN, WINX = g.WINX, M_I = g.M_I;

  for (int c = 0; c < nch; c++) {
    const int s = c % FT_STAGES, base = c * FT_CH, cnt = min(FT_CH, np1 - base);
    mbar_wait(&full[s], (uint32_t)((c / FT_STAGES) & 1));
    const double* __restrict__ tb = sT + (size_t)s * FT_CH * FT_W;
    if (warp_simple) {
      // (an explicit software pipeline of the reciprocal chain across iterations was tried: 0.81 -> 0.75 of the DFMA
      // peak -- the compiler's own interleaving of the two unrolled iterations is better)
#pragma unroll 2
      for (int j = 0; j < cnt; j++) {
        const double p = sp[base + j];
        const double x = fma(-kpar, p, dre);            // Re(den) before the -+ n qs shift
        const double drp = x - nq, drm = x + nq;        // den = ms om - kpar p_par -+ n qs (resU, src/ALPS_fns.f90:1591)
        const double dp = fma(drp, drp, dim2), dm = fma(drm, drm, dim2);
        const double inv = fast_rcp(dp * dm);           // both reciprocals from one: 1/dp = dm/(dp dm)
        const double ip = dm * inv, im = dp * inv;
        const double rpr = drp * ip, rpi = -dim * ip, rmr = drm * im, rmi = -dim * im;
        const double2* t2 = reinterpret_cast<const double2*>(tb + j * FT_W);
#pragma unroll
        for (int q = 0; q < FT_W / 2; q++) {
          const double2 t = t2[q];
          PR[2 * q] = fma(rpr, t.x, PR[2 * q]);         PI[2 * q] = fma(rpi, t.x, PI[2 * q]);
          MR[2 * q] = fma(rmr, t.x, MR[2 * q]);         MI[2 * q] = fma(rmi, t.x, MI[2 * q]);
          PR[2 * q + 1] = fma(rpr, t.y, PR[2 * q + 1]); PI[2 * q + 1] = fma(rpi, t.y, PI[2 * q + 1]);
          MR[2 * q + 1] = fma(rmr, t.y, MR[2 * q + 1]); MI[2 * q + 1] = fma(rmi, t.y, MI[2 * q + 1]);
        }
      }
    } else {
      for (int j = 0; j < cnt; j++) {
        const int ipar = base + j + 1;
        const double p = sp[base + j];
        const double x = fma(-kpar, p, dre);
        const double drp = x - nq, drm = x + nq;
        const double dp = fma(drp, drp, dim2), dm = fma(drm, drm, dim2);
        const double itab = (ipar == 1 || ipar == np1) ? 1.0 : 0.5;     // 1 / w_tab
        double fp, fm;     // w_eff / w_tab
        if (plain0) fp = act0 ? 1.0 : 0.0;
        else fp = (range_w(ipar, pe[0].lo1, pe[0].hi1) + range_w(ipar, pe[0].lo2, pe[0].hi2)) * itab;
        if (plain1) fm = act1 ? 1.0 : 0.0;
        else fm = (range_w(ipar, pe[1].lo1, pe[1].hi1) + range_w(ipar, pe[1].lo2, pe[1].hi2)) * itab;
        // a sign with zero weight must not poison the shared reciprocal (real omega exactly on a node)
        const double dps = fp != 0.0 ? dp : 1.0, dms = fm != 0.0 ? dm : 1.0;
        const double inv = fast_rcp(dps * dms);
        const double ip = fp * dms * inv, im = fm * dps * inv;
        const double rpr = drp * ip, rpi = -dim * ip, rmr = drm * im, rmi = -dim * im;
        const double* t = tb + j * FT_W;
#pragma unroll
        for (int q = 0; q < FT_W; q++) {
          const double tq = t[q];
          PR[q] = fma(rpr, tq, PR[q]); PI[q] = fma(rpi, tq, PI[q]);
          MR[q] = fma(rmr, tq, MR[q]); MI[q] = fma(rmi, tq, MI[q]);
        }
#pragma unroll
        for (int sg = 0; sg < 2; sg++) {
          if ((sg == 0 ? act0 : act1) && (pe[sg].flags & PLAN_NEAR)) {
            int jw = ipar - (pe[sg].ipar_res - M_I - 2);
            if (jw < 0 || jw >= WIN) jw = (ipar <= 3) ? WIN + ipar - 1 : -1;
            if (jw >= 0) {
              // G_x = om GA_x + GB_x at this node: the (x, m = 0) table entries without their trapezoid weight
              double* gw = gwin + ((item0 + sg) * WINX + jw) * 6;
              const double ga = t[0] * itab, gb = t[3] * itab, gc = t[5] * itab;
              gw[0] = fma(omr, ga, t[6] * itab); gw[1] = omi * ga;
              gw[2] = fma(omr, gb, t[9] * itab); gw[3] = omi * gb;
              gw[4] = fma(omr, gc, t[11] * itab); gw[5] = omi * gc;
            }
          }
        }
      }
    }
    __syncthreads();      // every warp is done with stage s
    if (tid == 0 && c + FT_STAGES < nch) issue(c + FT_STAGES);
  }
  // S_j = om * (sum R TA_j) + sum R TB_j
#pragma unroll
  for (int sg = 0; sg < 2; sg++) {
    if (sg == 0 ? !act0 : !act1) continue;
    double* o = Sbulk + (item0 + sg) * 12;
#pragma unroll
    for (int j = 0; j < 6; j++) {
      const double ar = sg == 0 ? PR[j] : MR[j], ai = sg == 0 ? PI[j] : MI[j];
      const double br = sg == 0 ? PR[6 + j] : MR[6 + j], bi = sg == 0 ? PI[6 + j] : MI[6 + j];
      o[2 * j] = fma(omr, ar, fma(-omi, ai, br));
      o[2 * j + 1] = fma(omr, ai, fma(omi, ar, bi));
    }
  }
}

void launch_fast(const GlobalDev* g, const double* om, int n_om, const FastItem* items, int nitems,
                 const PlanEntry* plan, double* Sbulk, double* gwin, int npar, double kpar, int variant,
                 cudaStream_t st) {
  if (n_om <= 0 || nitems <= 0) return;
  if (variant == 0) {
    dim3 grid((n_om + 127) / 128, nitems);
    k_fast<<<grid, 128, 0, st>>>(g, om, n_om, items, plan, Sbulk, gwin);
    return;
  }
  const int nch = (npar - 1 + FT_CH - 1) / FT_CH;
  const size_t smem = ((size_t)FT_STAGES * FT_CH * FT_W + (size_t)nch * FT_CH) * sizeof(double);
  // opt-in shared memory, per device: raised to the device limit once (the p_par axis of any grid that fits is covered)
  static PerDeviceOnce once;
  if (once.first()) {
    int dev = 0, lim = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(k_fast_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, lim - 1024);
  }
  dim3 grid((n_om + FT_THREADS - 1) / FT_THREADS, nitems);
  k_fast_tiled<<<grid, FT_THREADS, smem, st>>>(g, om, n_om, items, plan, Sbulk, gwin, kpar);
}

}  // namespace alps
