// alps_b200: the regular (p_perp, p_par) quadrature of chi_s(omega,k) -- the dominant kernel.
//
// Replaces: integrate() and its resU()/int_T() calls, src/ALPS_fns.f90:799-864, 1560-1707, for
// every (omega, species, harmonic n, sign, tensor component) of a batch of omegas.
//
// For a non-relativistic species the integrand of integrate() is
//     resU * T_mode = Num(iperp,ipar) / den(n,ipar) * t_mode(n,iperp) * p_par^m
// with Num = qs (om A + (kpar/ms)(p_perp B - p_par A))        [A = d_perp f0, B = d_par f0]
//      den = ms om - kpar p_par - n qs                         (independent of iperp)
//      t   = one of {J_n^2, p_perp J_n J_n', p_perp^2 J_n'^2} times constants (same for +n/-n).
// One CTA owns (omega, species, 16 harmonics) and walks the whole grid: for each tile of 128
// p_par columns it forms   G_x(n, ipar) = sum_iperp w_perp t_x(n,iperp) Num(iperp,ipar)
// (x = a,b,c) from the TMA-staged A' / C' tiles -- Num = om A' + C' is linear in the two real tables,
// so the kernel accumulates sum W A' and sum W C' (48 FMAs per thread and grid row) and forms
// om * (sum W A') + (sum W C') once per column -- then the
// epilogue divides by the two resonance denominators (+n, -n), applies the p_par trapezoid
// weights of the resonance plan and the p_par^m moments, and reduces over p_par with warp
// shuffles.  Six complex moment sums per (n, sign) leave the kernel; all six tensor components
// are linear in them (assemble kernel).
//
// Pipeline: one producer warp issues cp.async.bulk.tensor (TMA) loads of the A, C' and W tiles
// into a 4-stage shared-memory ring guarded by full/empty mbarriers; eight consumer warps run
// the FP64 FMA loop on 12 x 2 register tiles (48 complex-half accumulators per thread).
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "pipe.cuh"

namespace alps {

// RG = row groups (4 harmonics each; 2 consumer warps per row group), BKT = p_perp rows per
// pipeline stage, NST = stages.
template <int RG, int BKT>
struct alignas(128) QuadStage {
  double A[BKT * BN];
  double C[BKT * BN];
  double W[BKT * 12 * RG];
};

template <int RG, int BKT, int NST>
struct QuadSmem {
  QuadStage<RG, BKT> st[NST];
  double red[2 * RG][RG][4][2][12];   // [warp][row group][harmonic][sign][12]
  unsigned long long full[NST];
  unsigned long long empty[NST];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// PW = 1: dedicated TMA producer warp; PW = 0: thread 0 of consumer warp 0 issues the loads
// (lets 12 consumer warps = 3 per SM sub-partition keep 168 registers each).
template <int RG, int BKT, int NST, int PW, int UNR, bool STORE>
__global__ void __launch_bounds__((2 * RG + PW) * 32, 1) k_quad(const __grid_constant__ QuadParams P) {
  constexpr int CONSUMER_WARPS = 2 * RG;
  constexpr int BM = 12 * RG;
  constexpr int BK = BKT;
  constexpr int STAGES = NST;
  constexpr uint32_t STAGE_TX_BYTES = (2 * BK * BN + BK * BM) * sizeof(double);
  typedef QuadSmem<RG, BKT, NST> Smem;
  extern __shared__ unsigned char smem_raw[];
  // 128-byte alignment for the TMA destinations; offset arithmetic keeps the shared address space
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));

  // CTA -> (omega, tile, p_par split): with nsplit > 1 (few omegas in flight) the p_par tiles of one
  // (omega, tile) are spread over nsplit CTAs, each leaving its own partial moment sums
  const int nsplit = P.nsplit;
  const int jsplit = blockIdx.x % nsplit;
  const int tile_id = (blockIdx.x / nsplit) % P.ntiles;
  const int iom = blockIdx.x / (nsplit * P.ntiles);
  const QuadTile tile = P.tiles[tile_id];
  const GlobalDev& g = *P.g;
  const SpeciesDev& sp = g.sp[tile.s];
  const int nperp = g.nperp, npar = g.npar;
  const int KC = (nperp - 1 + BK - 1) / BK;
  const int NTall = (npar - 1 + BN - 1) / BN;
  const int NT = (NTall - jsplit + nsplit - 1) / nsplit;   // p_par tiles of this CTA: jsplit, jsplit+nsplit, ...
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < CONSUMER_WARPS * RG * 96; i += blockDim.x) (&sm.red[0][0][0][0][0])[i] = 0.0;
  __syncthreads();

  const CUtensorMap* tmA = &P.tmA[tile.s];
  const CUtensorMap* tmC = &P.tmC[tile.s];
  const CUtensorMap* tmW = &P.tmW[tile.s];
  const int T = NT * KC;   // pipeline iterations of this CTA
  auto issue = [&](int it) {
    const int stg = it % STAGES, ntl = it / KC, kc = it - ntl * KC, nt = jsplit + ntl * nsplit;
    mbar_expect_tx(&sm.full[stg], STAGE_TX_BYTES);
    tma_load_2d(sm.st[stg].A, tmA, nt * BN, kc * BK, &sm.full[stg]);
    tma_load_2d(sm.st[stg].C, tmC, nt * BN, kc * BK, &sm.full[stg]);
    tma_load_2d(sm.st[stg].W, tmW, 3 * tile.n0, kc * BK, &sm.full[stg]);
  };
  if (PW && warp == CONSUMER_WARPS) {
    // ------------------------------------------------------------ dedicated TMA producer warp
    if (lane == 0) {
      for (int it = 0; it < T; it++) {
        if (it >= STAGES) mbar_wait(&sm.empty[it % STAGES], ((it / STAGES) & 1) ^ 1);
        issue(it);
      }
    }
    return;
  }
  if (!PW && threadIdx.x == 0) {
    for (int it = 0; it < STAGES - 1 && it < T; it++) issue(it);
  }

  // ---------------------------------------------------------------- consumers
  // a warp covers all 4 row groups x 8 column groups: the A'/C' operands are then shared by 4 lanes
  // (one shared-memory wavefront per half-warp instead of two), the W operands by 8 lanes
  static_assert(RG == 4, "lane mapping assumes 4 row groups");
  const int rg = lane >> 3;                    // row group: harmonics n0 + 4 rg .. +3
  const int cg = (warp << 3) | (lane & 7);     // column group: columns 2 cg, 2 cg + 1 of the tile
  const double omr = P.om[2 * iom], omi = P.om[2 * iom + 1];
  const double qs = sp.qs, ms = sp.ms, kpar = g.kpar;
  const double* __restrict__ ppar = sp.ppar;
  const size_t item0 = (size_t)iom * g.NI + sp.item_base;
  const int WIN = g.WIN, WINX = g.WINX, M_I = g.M_I;

  int stage = 0;
  uint32_t phase = 0, ready = 0;
  int git = 0;
  for (int ntl = 0; ntl < NT; ntl++) {
    const int nt = jsplit + ntl * nsplit;
    double ar[12][2], ai[12][2];
#pragma unroll
    for (int r = 0; r < 12; r++) {
      ar[r][0] = ar[r][1] = 0.0;
      ai[r][0] = ai[r][1] = 0.0;
    }
    for (int kc = 0; kc < KC; kc++, git++) {
      if (!PW) {
        // inline producer: refill the stage released by iteration git-1
        if (threadIdx.x == 0) {
          const int nx = git + STAGES - 1;
          if (nx < T) {
            if (nx >= STAGES) mbar_wait(&sm.empty[nx % STAGES], ((nx / STAGES) & 1) ^ 1);
            issue(nx);
          }
        }
        __syncwarp();
      }
      if (!ready) mbar_wait(&sm.full[stage], phase);
      {
        // test the next stage's barrier now; the answer is only needed after this stage's FMAs
        const int ns = (stage + 1 == STAGES) ? 0 : stage + 1;
        ready = mbar_test(&sm.full[ns], (stage + 1 == STAGES) ? (phase ^ 1) : phase);
      }
      const double* sA = sm.st[stage].A + 2 * cg;
      const double* sC = sm.st[stage].C + 2 * cg;
      const double* sW = sm.st[stage].W + 12 * rg;
      // operands of row kk+1 are fetched from shared memory while row kk is in the FMA pipe
      double2 a = *reinterpret_cast<const double2*>(sA);
      double2 c = *reinterpret_cast<const double2*>(sC);
      double2 w[6];
#pragma unroll
      for (int q = 0; q < 6; q++) w[q] = reinterpret_cast<const double2*>(sW)[q];
#pragma unroll UNR
      for (int kk = 0; kk < BK; kk++) {
        // Num = om*A' + C' is linear in the two real tables, so the p_perp sums of A' and of C' are
        // accumulated separately (same 48 FMAs) and combined with om in the epilogue:
        //   sum_iperp W Num = om * sum W A' + sum W C'
        const double nr0 = c.x, ni0 = a.x;   // "real" accumulators collect C', "imaginary" ones A'
        const double nr1 = c.y, ni1 = a.y;
        double2 wc[6];
#pragma unroll
        for (int q = 0; q < 6; q++) wc[q] = w[q];
        if (kk + 1 < BK) {
          a = *reinterpret_cast<const double2*>(sA + (kk + 1) * BN);
          c = *reinterpret_cast<const double2*>(sC + (kk + 1) * BN);
#pragma unroll
          for (int q = 0; q < 6; q++) w[q] = reinterpret_cast<const double2*>(sW + (kk + 1) * BM)[q];
        }
        // numerator-major order: 12 consecutive FMAs share one multiplicand, which the operand
        // reuse cache keeps (a DFMA with three fresh 64-bit operands costs a third register-read cycle)
#pragma unroll
        for (int q = 0; q < 6; q++) {
          ar[2 * q][0] = fma(nr0, wc[q].x, ar[2 * q][0]);
          ar[2 * q + 1][0] = fma(nr0, wc[q].y, ar[2 * q + 1][0]);
        }
#pragma unroll
        for (int q = 0; q < 6; q++) {
          ai[2 * q][0] = fma(ni0, wc[q].x, ai[2 * q][0]);
          ai[2 * q + 1][0] = fma(ni0, wc[q].y, ai[2 * q + 1][0]);
        }
#pragma unroll
        for (int q = 0; q < 6; q++) {
          ar[2 * q][1] = fma(nr1, wc[q].x, ar[2 * q][1]);
          ar[2 * q + 1][1] = fma(nr1, wc[q].y, ar[2 * q + 1][1]);
        }
#pragma unroll
        for (int q = 0; q < 6; q++) {
          ai[2 * q][1] = fma(ni1, wc[q].x, ai[2 * q][1]);
          ai[2 * q + 1][1] = fma(ni1, wc[q].y, ai[2 * q + 1][1]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[stage]);
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }

    // ------------------------------------------------------------ epilogue of this p_par tile
    const int ipar0 = nt * BN + 2 * cg + 1;
    if (!STORE) {
      // G = om * GA + GB  (ai holds GA = sum W A', ar holds GB = sum W C')
#pragma unroll
      for (int r = 0; r < 12; r++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const double ga = ai[r][c];
          ar[r][c] = fma(omr, ga, ar[r][c]);
          ai[r][c] = omi * ga;
        }
      }
    }
    if (STORE) {
      // k-hoisted tables (alps_b200_set_mode(1)): the raw sums GA = sum W A' (ai) and GB = sum W C' (ar).
      // Layout [n][ipar-1][GAa, GBa, GAb, GBb, GAc, GBc].
      double* gt = P.gtab[tile.s];
#pragma unroll
      for (int nn = 0; nn < 4; nn++) {
        const int nabs = tile.n0 + 4 * rg + nn;
        if (nabs > sp.nhi_shard) continue;
#pragma unroll
        for (int c = 0; c < 2; c++) {
          const int ipar = ipar0 + c;
          if (ipar > npar - 1) continue;
          double2* o = reinterpret_cast<double2*>(gt + ((size_t)nabs * (npar - 1) + (ipar - 1)) * 6);
#pragma unroll
          for (int x = 0; x < 3; x++) o[x] = make_double2(ai[3 * nn + x][c], ar[3 * nn + x][c]);
        }
      }
      continue;
    }
    double pp_[2];
    pp_[0] = (ipar0 <= npar - 1) ? ppar[ipar0] : 0.0;
    pp_[1] = (ipar0 + 1 <= npar - 1) ? ppar[ipar0 + 1] : 0.0;
    // two chunks of two harmonics: 2 x 2 signs x 12 = 48 partial sums per thread and chunk, reduced over
    // the 32 lanes by recursive halving (93 shuffles per 96 sums instead of 480 butterfly shuffles)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double Sv[48];
#pragma unroll
      for (int q = 0; q < 48; q++) Sv[q] = 0.0;
#pragma unroll
      for (int nnl = 0; nnl < 2; nnl++) {
        const int nn = 2 * h + nnl;
        const int nabs = tile.n0 + 4 * rg + nn;
        if (nabs > sp.nhi_shard) continue;
#pragma unroll
        for (int sg = 0; sg < 2; sg++) {
          if (nabs == 0 && sg == 1) continue;
          const size_t item = item0 + 2 * nabs + sg;
          const PlanEntry pe = P.plan[item];
          if (!(pe.flags & PLAN_ACTIVE)) continue;
          const double nq = (sg ? -1.0 : 1.0) * (double)nabs * qs;
          double* S = &Sv[(nnl * 2 + sg) * 12];
#pragma unroll
          for (int c = 0; c < 2; c++) {
            const int ipar = ipar0 + c;
            if (ipar > npar - 1) continue;
            const double w = range_w(ipar, pe.lo1, pe.hi1) + range_w(ipar, pe.lo2, pe.hi2);
            const double p = pp_[c];
            if (w != 0.0) {
              // 1/den with den = ms om - kpar p_par - n qs   (resU, src/ALPS_fns.f90:1591-1592)
              const double dr = ms * omr - kpar * p - nq, di = ms * omi;
              const double t = w / (dr * dr + di * di);
              const cd R = mk(dr * t, -di * t);
              const cd Va = R * mk(ar[3 * nn + 0][c], ai[3 * nn + 0][c]);
              const cd Vb = R * mk(ar[3 * nn + 1][c], ai[3 * nn + 1][c]);
              const cd Vc = R * mk(ar[3 * nn + 2][c], ai[3 * nn + 2][c]);
              const double p2 = p * p;
              S[0] += Va.x;       S[1] += Va.y;        // sum U J^2
              S[2] += p * Va.x;   S[3] += p * Va.y;    // sum U J^2 p_par
              S[4] += p2 * Va.x;  S[5] += p2 * Va.y;   // sum U J^2 p_par^2
              S[6] += Vb.x;       S[7] += Vb.y;        // sum U p_perp J J'
              S[8] += p * Vb.x;   S[9] += p * Vb.y;    // sum U p_perp J J' p_par
              S[10] += Vc.x;      S[11] += Vc.y;       // sum U p_perp^2 J'^2
            }
            if (pe.flags & PLAN_NEAR) {
              int j = ipar - (pe.ipar_res - M_I - 2);
              if (j < 0 || j >= WIN) j = (ipar <= 3) ? WIN + ipar - 1 : -1;   // nodes 1..3: funct_g fallback
              if (j >= 0) {
                double* gw = P.gwin + (item * WINX + j) * 6;
#pragma unroll
                for (int x = 0; x < 3; x++) {
                  gw[2 * x] = ar[3 * nn + x][c];
                  gw[2 * x + 1] = ai[3 * nn + x][c];
                }
              }
            }
          }
        }
      }
      // recursive halving over the 8 lanes of a row group (lane bits 2,1,0): after the step with mask m
      // a lane keeps the half selected by its bit m
#pragma unroll
      for (int step = 0; step < 3; step++) {
        const int N = 24 >> step, mask = 4 >> step;
        const bool up = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < 24; i++) {
          if (i < N) {
            const double send = up ? Sv[i] : Sv[i + N];
            const double keep = up ? Sv[i + N] : Sv[i];
            Sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
          }
        }
      }
      {
        const int base = 24 * ((lane >> 2) & 1) + 12 * ((lane >> 1) & 1) + 6 * (lane & 1);
        double* red = &sm.red[warp][rg][0][0][0] + 48 * h + base;
#pragma unroll
        for (int i = 0; i < 6; i++) red[i] += Sv[i];
      }
    }
  }

  if (STORE) return;
  // ---------------------------------------------------------------- write the moment sums
  asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
  for (int i = threadIdx.x; i < RG * 96; i += CONSUMER_WARPS * 32) {
    const int rgq = i / 96, rem = i % 96;
    const int nn = rem / 24, sg = (rem % 24) / 12, q = rem % 12;
    const int nabs = tile.n0 + 4 * rgq + nn;
    if (nabs > sp.nhi_shard) continue;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < CONSUMER_WARPS; w++) t += sm.red[w][rgq][nn][sg][q];
    P.Sbulk[((item0 + 2 * nabs + sg) * nsplit + jsplit) * 12 + q] = t;
  }
}

// ---------------------------------------------------------------- variants / launcher
template <int RG, int BKT, int NST, int PW, int UNR, bool STORE>
static cudaError_t launch_one(const QuadParams& P, cudaStream_t st) {
  static PerDeviceOnce once;
  const size_t smem = sizeof(QuadSmem<RG, BKT, NST>) + 128;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(k_quad<RG, BKT, NST, PW, UNR, STORE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_quad<RG, BKT, NST, PW, UNR, STORE><<<P.n_om * P.ntiles * P.nsplit, (2 * RG + PW) * 32, smem, st>>>(P);
  return cudaGetLastError();
}
template <int RG, int BKT, int NST, int PW, int UNR>
static cudaError_t launch_variant(const QuadParams& P, bool store, cudaStream_t st) {
  return store ? launch_one<RG, BKT, NST, PW, UNR, true>(P, st) : launch_one<RG, BKT, NST, PW, UNR, false>(P, st);
}

QuadVariant quad_variant(int id) {
  switch (id) {
    // DMMA (quad_mma.cu)
    case 9: return QuadVariant{9, MMA_NH, 32, 2, 128};
    case 12: return QuadVariant{12, MMA_NH, 32, 2, 64};
    case 15: case 91: case 92: case 93: return QuadVariant{id, MMA_NH, 32, 2, 128};
    case 16: return QuadVariant{16, MMA_NH, 16, 4, 128};
    case 17: return QuadVariant{17, MMA_NH, 32, 2, 64};
    // DFMA
    case 5: return QuadVariant{5, 16, 16, 4, BN};    // 8 consumer warps, inline producer, BK=16
    case 8: return QuadVariant{8, 16, 32, 2, BN};    // 8 consumer warps, inline producer, BK=32, 2 stages
    default: return QuadVariant{0, 16, 8, 4, BN};    // 8 consumer warps + producer, BK=8
  }
}

cudaError_t launch_quad(const QuadParams& P, int variant, bool store, cudaStream_t st) {
  if (P.n_om <= 0 || P.ntiles <= 0) return cudaSuccess;
  switch (variant) {
    case 5: return launch_variant<4, 16, 4, 0, 16>(P, store, st);
    case 8: return launch_variant<4, 32, 2, 0, 32>(P, store, st);
    default: return launch_variant<4, 8, 4, 1, 8>(P, store, st);
  }
}

}  // namespace alps
