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
// (x = a,b,c) with Num evaluated at every grid point from the TMA-staged A / C' tiles, then the
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

namespace alps {

struct alignas(128) QuadStage {
  double A[BK * BN];
  double C[BK * BN];
  double W[BK * BM];
};
constexpr uint32_t STAGE_TX_BYTES = (2 * BK * BN + BK * BM) * sizeof(double);

struct QuadSmem {
  QuadStage st[STAGES];
  double red[CONSUMER_WARPS][4][2][12];
  unsigned long long full[STAGES];
  unsigned long long empty[STAGES];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(b)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(QUAD_THREADS, 1) k_quad(const __grid_constant__ QuadParams P) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte alignment for the TMA destinations; offset arithmetic keeps the shared address space
  QuadSmem& sm = *reinterpret_cast<QuadSmem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));

  const int tile_id = blockIdx.x % P.ntiles;
  const int iom = blockIdx.x / P.ntiles;
  const QuadTile tile = P.tiles[tile_id];
  const GlobalDev& g = *P.g;
  const SpeciesDev& sp = g.sp[tile.s];
  const int nperp = g.nperp, npar = g.npar;
  const int KC = (nperp - 1 + BK - 1) / BK;
  const int NT = (npar - 1 + BN - 1) / BN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], CONSUMER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < CONSUMER_WARPS * 96; i += blockDim.x) (&sm.red[0][0][0][0])[i] = 0.0;
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const CUtensorMap* tmA = &P.tmA[tile.s];
      const CUtensorMap* tmC = &P.tmC[tile.s];
      const CUtensorMap* tmW = &P.tmW[tile.s];
      int stage = 0;
      uint32_t phase = 0;
      for (int nt = 0; nt < NT; nt++) {
        for (int kc = 0; kc < KC; kc++) {
          mbar_wait(&sm.empty[stage], phase ^ 1);
          mbar_expect_tx(&sm.full[stage], STAGE_TX_BYTES);
          tma_load_2d(sm.st[stage].A, tmA, nt * BN, kc * BK, &sm.full[stage]);
          tma_load_2d(sm.st[stage].C, tmC, nt * BN, kc * BK, &sm.full[stage]);
          tma_load_2d(sm.st[stage].W, tmW, 3 * tile.n0, kc * BK, &sm.full[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    return;
  }

  // ---------------------------------------------------------------- consumers
  const int rg = warp >> 1;                    // row group: harmonics n0 + 4 rg .. +3
  const int cg = ((warp & 1) << 5) | lane;     // column group: columns 2 cg, 2 cg + 1 of the tile
  const double omr = P.om[2 * iom], omi = P.om[2 * iom + 1];
  const double qs = sp.qs, ms = sp.ms, kpar = g.kpar;
  const double* __restrict__ ppar = sp.ppar;
  const size_t item0 = (size_t)iom * g.NI + sp.item_base;
  const int WIN = g.WIN, WINX = g.WINX, M_I = g.M_I;

  int stage = 0;
  uint32_t phase = 0;
  for (int nt = 0; nt < NT; nt++) {
    double ar[12][2], ai[12][2];
#pragma unroll
    for (int r = 0; r < 12; r++) {
      ar[r][0] = ar[r][1] = 0.0;
      ai[r][0] = ai[r][1] = 0.0;
    }
    for (int kc = 0; kc < KC; kc++) {
      mbar_wait(&sm.full[stage], phase);
      const double* sA = sm.st[stage].A + 2 * cg;
      const double* sC = sm.st[stage].C + 2 * cg;
      const double* sW = sm.st[stage].W + 12 * rg;
#pragma unroll
      for (int kk = 0; kk < BK; kk++) {
        const double2 a = *reinterpret_cast<const double2*>(sA + kk * BN);
        const double2 c = *reinterpret_cast<const double2*>(sC + kk * BN);
        // numerator of resU at this grid point: Num = om * A' + C'
        const double nr0 = fma(omr, a.x, c.x), ni0 = omi * a.x;
        const double nr1 = fma(omr, a.y, c.y), ni1 = omi * a.y;
        const double2* wp = reinterpret_cast<const double2*>(sW + kk * BM);
#pragma unroll
        for (int q = 0; q < 6; q++) {
          const double2 w = wp[q];
          ar[2 * q][0] = fma(w.x, nr0, ar[2 * q][0]);
          ai[2 * q][0] = fma(w.x, ni0, ai[2 * q][0]);
          ar[2 * q][1] = fma(w.x, nr1, ar[2 * q][1]);
          ai[2 * q][1] = fma(w.x, ni1, ai[2 * q][1]);
          ar[2 * q + 1][0] = fma(w.y, nr0, ar[2 * q + 1][0]);
          ai[2 * q + 1][0] = fma(w.y, ni0, ai[2 * q + 1][0]);
          ar[2 * q + 1][1] = fma(w.y, nr1, ar[2 * q + 1][1]);
          ai[2 * q + 1][1] = fma(w.y, ni1, ai[2 * q + 1][1]);
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
    double pp_[2];
    pp_[0] = (ipar0 <= npar - 1) ? ppar[ipar0] : 0.0;
    pp_[1] = (ipar0 + 1 <= npar - 1) ? ppar[ipar0 + 1] : 0.0;
#pragma unroll
    for (int nn = 0; nn < 4; nn++) {
      const int nabs = tile.n0 + 4 * rg + nn;
      if (nabs > sp.nhi_shard) continue;   // warp-uniform
#pragma unroll
      for (int sg = 0; sg < 2; sg++) {
        if (nabs == 0 && sg == 1) continue;
        const size_t item = item0 + 2 * nabs + sg;
        const PlanEntry pe = P.plan[item];
        if (!(pe.flags & PLAN_ACTIVE)) continue;   // warp-uniform
        const double nq = (sg ? -1.0 : 1.0) * (double)nabs * qs;
        double S[12];
#pragma unroll
        for (int q = 0; q < 12; q++) S[q] = 0.0;
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
#pragma unroll
        for (int q = 0; q < 12; q++) S[q] = warp_sum(S[q]);
        if (lane == 0) {
#pragma unroll
          for (int q = 0; q < 12; q++) sm.red[warp][nn][sg][q] += S[q];
        }
      }
    }
  }

  // ---------------------------------------------------------------- write the moment sums
  asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory");
  for (int i = threadIdx.x; i < 4 * 96; i += CONSUMER_WARPS * 32) {
    const int rgq = i / 96, rem = i % 96;
    const int nn = rem / 24, sg = (rem % 24) / 12, q = rem % 12;
    const int nabs = tile.n0 + 4 * rgq + nn;
    if (nabs > sp.nhi_shard) continue;
    P.Sbulk[(item0 + 2 * nabs + sg) * 12 + q] = sm.red[2 * rgq][nn][sg][q] + sm.red[2 * rgq + 1][nn][sg][q];
  }
}

size_t quad_smem_bytes() { return sizeof(QuadSmem) + 128; }

cudaError_t launch_quad(const QuadParams& P, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_quad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)quad_smem_bytes());
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if (P.n_om <= 0 || P.ntiles <= 0) return cudaSuccess;
  k_quad<<<P.n_om * P.ntiles, QUAD_THREADS, quad_smem_bytes(), st>>>(P);
  return cudaGetLastError();
}

}  // namespace alps
