// alps_b200: the regular (p_perp, p_par) quadrature of chi_s(omega,k) on the FP64 tensor pipe (DMMA).
//
// Replaces: integrate() and its resU()/int_T() calls, src/ALPS_fns.f90:799-864, 1560-1707 -- the same
// formulation as quad_kernel.cu (which stays as the DFMA variants 0/5/8):
//     G_x(n, ipar) = sum_iperp W_x(n, iperp) * {A', C'}(iperp, ipar),      x = a, b, c
// is a real contraction over p_perp with omega-independent operands; omega enters once per column,
// G = om * GA + GB, in the epilogue that applies the resonance denominators, the p_par trapezoid
// weights of the resonance plan and the p_par moments.  Here the contraction is issued as
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4): the same 64 FMA/clk/SM FP64 units as DFMA, but one warp
// instruction carries 256 FMAs and reads 4 registers, so the pipe is not limited by instruction issue,
// register-file bandwidth or shared-memory operand traffic (measured: DMMA 37.07 TFLOP/s, DFMA 33.3).
//
// Operands are pre-arranged in HBM in fragment order (setup_kernels.cu: k_frag_table, k_build_Wf):
//   Xf[nt][ks][4 TW] table tile of TW = 16 NW p_par columns x 4 p_perp rows (k-step ks): element
//                    (row 4 ks + t, column TW nt + 16 w + 8 j + g) at 64 w + 2 (4 g + t) + j
//   Wf[hb][ks][192]  weights of harmonics 16 hb .. 16 hb + 15: element (row 4 ks + t, type x,
//                    harmonic 16 hb + 8 h + g) at 6 (4 g + t) + 2 x + h
// so a pipeline stage is one contiguous TMA bulk copy per operand and every thread fetches its DMMA
// fragments with conflict-free LDS.128.  Zero padding (rows to 32, columns to 128, harmonics to 16)
// lives in the buffers.
//
// CTA = (omega, species, 16 harmonics) x all p_par tiles (or every nsplit-th), NW warps; warp w owns the
// 48 x 16 block of columns 16 w .. 16 w + 15 for both tables: 6 M-tiles x (2 + 2) N-tiles = 24 DMMA per
// k-step, 48 FP64 accumulators per thread.  NW = 4 runs two CTAs per SM, so that the epilogue of one
// CTA (FP64 pipe, shuffles) overlaps the DMMA stream of the other.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "pipe.cuh"

namespace alps {

constexpr int KS_W = 4 * 3 * MMA_NH;    // doubles per k-step of one weight tile

// NW = warps per CTA = 16-column blocks per p_par tile (tile width TW = 16 NW), KSTG = k-steps per stage
template <int NW, int KSTG>
struct alignas(128) MmaStage {
  double A[KSTG * 64 * NW];
  double C[KSTG * 64 * NW];
  double W[KSTG * KS_W];
};

template <int NW, int KSTG, int NST>
struct MmaSmem {
  MmaStage<NW, KSTG> st[NST];
  double red[NW][MMA_NH][2][12];   // [warp][harmonic][sign][12]
  PlanEntry plan[MMA_NH][2];       // resonance plan of this CTA's (harmonic, sign) items: the same for every p_par tile
  unsigned long long full[NST];
  unsigned long long empty[NST];
};

// D(8x8) += A(8x4, row) * B(4x8, col); lane = 4 g + t holds A[g][t], B[t][g], C[g][2t], C[g][2t+1]
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// HS = 1: a warp owns all 16 harmonics of its 16 columns (6 M-tiles); HS = 2: two warps share the column
// block, one per group of 8 harmonics (3 M-tiles each) -- 4 warps per SM sub-partition instead of 2 hide
// the stage hand-over (barrier wait, first fragment loads) of one warp behind the DMMAs of the others.
// PK = 1: instantiation for the species' last tiles when their upper harmonic group is (nearly) empty -- see `pack`
template <int NW, int HS, int KSTG, int NST, bool STORE, int DBG = 0, int PK = 0>
__global__ void __launch_bounds__(32 * NW * HS, NW == 4 ? 2 : 1) k_quad_mma(const __grid_constant__ QuadParams P) {
  constexpr int TW = 16 * NW;     // p_par columns per tile
  constexpr int MT = 6 / HS;      // M-tiles per warp
  constexpr int HP = 2 / HS;      // harmonics per thread
  constexpr int NSUM = 12 / HS;   // moment sums a lane keeps after the lane reduction
  constexpr int KS_A = 4 * TW;    // doubles per k-step of one table tile
  constexpr uint32_t BYTES_A = KSTG * KS_A * sizeof(double);
  constexpr uint32_t BYTES_W = KSTG * KS_W * sizeof(double);
  typedef MmaSmem<NW, KSTG, NST> Smem;
  extern __shared__ unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));

  const int nsplit = P.nsplit;
  const int jsplit = blockIdx.x % nsplit;
  const int bq = blockIdx.x / nsplit;
  const int tile_id = P.tile_major ? bq / P.n_om : bq % P.ntiles;
  const int iom = P.tile_major ? bq % P.n_om : bq / P.ntiles;
  constexpr bool LATV = (NW == 2) && !STORE;   // the latency variant of the single-omega chain
  const QuadTile tile = (LATV && tile_id < P.ntiles_inline) ? P.tile_inline[tile_id] : P.tiles[tile_id];
  const GlobalDev& g = *P.g;
  const SpeciesDev& sp = g.sp[tile.s];
  const int npar = (LATV && P.ntiles_inline > 0) ? P.npar_inline : g.npar;
  const int NKS = P.nks;
  const int KC = NKS / KSTG;
  const int NTall = (npar - 1 + TW - 1) / TW;
  const int NT = (NTall - jsplit + nsplit - 1) / nsplit;
  const int lane = threadIdx.x & 31;
  const int warp = (threadIdx.x >> 5) % NW;    // column block of this warp
  const int hsel = (threadIdx.x >> 5) / NW;    // HS = 2: harmonic group of this warp
  const int gq = lane >> 2, tq = lane & 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; s++) {
      mbar_init(&sm.full[s], 1);
      mbar_init(&sm.empty[s], NW * HS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const double* __restrict__ gA = P.Af[tile.s];
  const double* __restrict__ gC = P.Cf[tile.s];
  const double* __restrict__ gW = P.Wf[tile.s] + (size_t)(tile.n0 / MMA_NH) * NKS * KS_W;
  const int T = NT * KC;
  auto issue = [&](int it) {
    const int stg = it % NST, ntl = it / KC, kc = it - ntl * KC, nt = jsplit + ntl * nsplit;
    const size_t ks0 = (size_t)nt * NKS + (size_t)kc * KSTG;
    mbar_expect_tx(&sm.full[stg], 2 * BYTES_A + BYTES_W);
    tma_load_1d(sm.st[stg].A, gA + ks0 * KS_A, BYTES_A, &sm.full[stg]);
    tma_load_1d(sm.st[stg].C, gC + ks0 * KS_A, BYTES_A, &sm.full[stg]);
    tma_load_1d(sm.st[stg].W, gW + (size_t)kc * KSTG * KS_W, BYTES_W, &sm.full[stg]);
  };
  if (threadIdx.x == 0 && DBG != 3)
    for (int it = 0; it < NST - 1 && it < T; it++) issue(it);

  // single-omega chain: launched while k_plan still runs (programmatic dependent launch).  The barrier set-up, the table
  // stages and the whole p_perp contraction do not depend on k_plan's output; the resonance plan and omega (staged by
  // k_plan) do, and are needed by the epilogue only.  LATE (the latency variant): the wait for k_plan sits in front of the
  // first epilogue, so the contraction of the first p_par tile overlaps k_plan.
  constexpr bool LATE = (NW == 2) && !STORE;
  // (locals: values of *P.g used after the wait would otherwise be fetched again, one L2 round trip each)
  const int nhi_shard = sp.nhi_shard, nlo_shard = sp.nlo_shard;
  const size_t item0 = (size_t)iom * g.NI + sp.item_base;
  pdl_trigger();
  if (LATE && threadIdx.x == 0) lat_stamp(g, 2);
  double omr = 0.0, omi = 0.0;
  auto fetch_plan = [&]() {
    pdl_wait();
    if (!STORE && threadIdx.x < 2 * MMA_NH) {
      const int nabs = tile.n0 + (threadIdx.x >> 1), sg = threadIdx.x & 1;
      PlanEntry pe;
      pe.lo1 = 1; pe.hi1 = 0; pe.lo2 = 1; pe.hi2 = 0; pe.flags = 0; pe.ipar_res = 0; pe.upperlimit = 0; pe.pad = 0;
      if (nabs <= nhi_shard && nabs >= nlo_shard && !(nabs == 0 && sg == 1))
        pe = P.plan[item0 + 2 * nabs + sg];
      sm.plan[threadIdx.x >> 1][sg] = pe;
    }
    omr = P.om[2 * iom];
    omi = P.om[2 * iom + 1];
    __syncthreads();
  };
  if (!LATE) fetch_plan();
  else __syncthreads();   // the mbarriers are initialised

  const double qs = sp.qs, ms = sp.ms, kpar = g.kpar;
  const double* __restrict__ ppar = sp.ppar;
  const int WIN = g.WIN, WINX = g.WINX, M_I = g.M_I;

  // PK = 1, HS = 2: the 8 harmonics n0 + 8 hsel .. + 7 of this warp's group.  In a species' last tile the group may
  // hold only one or two harmonics of the summed range, or none (201 harmonics = 12 1/2 tiles + 1): their 3 r weight
  // rows then go into ONE M-tile (row 3 j + x = harmonic j, weight type x; the other rows zero) instead of one M-tile
  // per weight type, and an empty group issues no DMMA at all.  Same products, same order over p_perp: the sums are
  // bitwise those of the regular layout.  A separate instantiation, so that the regular tiles keep their code.
  int pack = 0;   // 0 regular: 3 M-tiles;  1 packed: 1 M-tile;  2 empty
  int pack_hi = 0;
  if (PK == 1 && HS == 2) {
    const int g0 = tile.n0 + 8 * hsel;
    const int lo = max(nlo_shard, g0), hi = min(nhi_shard, g0 + 7);
    if (hi < lo) pack = 2;
    else if (lo == g0 && hi - g0 <= 1) pack = 1;
    pack_hi = hi - g0;
  }

  // this lane's share of the moment sums.  HS = 1: harmonic n0 + 8 (tq >> 1) + gq, sign tq & 1, all 12;
  // HS = 2: harmonic n0 + 8 hsel + gq, sign tq >> 1, sums 6 (tq & 1) .. + 5
  double mine[NSUM];
#pragma unroll
  for (int q = 0; q < NSUM; q++) mine[q] = 0.0;

  int stage = 0;
  uint32_t phase = 0, ready = 0;
  int git = 0;
  for (int ntl = 0; ntl < NT; ntl++) {
    const int nt = jsplit + ntl * nsplit;
    double acc[MT][2][2][2];   // [M-tile HP x + h][table: 0 = A', 1 = C'][N-tile j][column pair]
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int q = 0; q < 8; q++) (&acc[m][0][0][0])[q] = 0.0;

    for (int kc = 0; kc < KC; kc++, git++) {
      if (threadIdx.x == 0 && DBG != 3) {
        const int nx = git + NST - 1;
        if (nx < T) {
          if (nx >= NST) mbar_wait(&sm.empty[nx % NST], ((nx / NST) & 1) ^ 1);
          issue(nx);
        }
      }
      __syncwarp();
      if (!ready && DBG != 3) mbar_wait(&sm.full[stage], phase);
      if (DBG != 3) {
        const int ns = (stage + 1 == NST) ? 0 : stage + 1;
        ready = mbar_test(&sm.full[ns], (stage + 1 == NST) ? (phase ^ 1) : phase);
      }
      const double* sA = sm.st[stage].A + 64 * warp + 2 * lane;
      const double* sC = sm.st[stage].C + 64 * warp + 2 * lane;
      const double* sW = sm.st[stage].W + 6 * lane;
      if (PK == 1 && HS == 2 && pack != 0) {
        if (pack == 1) {
          // packed M-tile: this lane's A operand is row gq = 3 j + x, i.e. W_x of harmonic j of the group, which the
          // regular fragment layout keeps at lane (j, tq), slot 2 x + hsel
          const int j = gq / 3, x = gq - 3 * j;
          const bool on = gq < 6 && j <= pack_hi;
          const double* sWp = sm.st[stage].W + 6 * (4 * (on ? j : 0) + tq) + 2 * (on ? x : 0) + hsel;
#pragma unroll
          for (int ks = 0; ks < KSTG; ks++) {
            const double2 a = *reinterpret_cast<const double2*>(sA + ks * KS_A);
            const double2 c = *reinterpret_cast<const double2*>(sC + ks * KS_A);
            const double wv = on ? sWp[ks * KS_W] : 0.0;
            dmma(acc[0][0][0], wv, a.x);
            dmma(acc[0][0][1], wv, a.y);
            dmma(acc[0][1][0], wv, c.x);
            dmma(acc[0][1][1], wv, c.y);
          }
        }
      } else
#pragma unroll
      for (int ks = 0; ks < KSTG; ks++) {
        const int kq = (DBG == 1) ? 0 : ks;   // DBG 1: one fragment load per stage
        const double2 a = *reinterpret_cast<const double2*>(sA + kq * KS_A);
        const double2 c = *reinterpret_cast<const double2*>(sC + kq * KS_A);
        const double2 w0 = reinterpret_cast<const double2*>(sW + kq * KS_W)[0];
        const double2 w1 = reinterpret_cast<const double2*>(sW + kq * KS_W)[1];
        const double2 w2 = reinterpret_cast<const double2*>(sW + kq * KS_W)[2];
        double wv[6] = {w0.x, w0.y, w1.x, w1.y, w2.x, w2.y};
        if (HS == 2) {
          wv[0] = hsel ? w0.y : w0.x;
          wv[1] = hsel ? w1.y : w1.x;
          wv[2] = hsel ? w2.y : w2.x;
        }
#pragma unroll
        for (int m = 0; m < MT; m++) {
          dmma(acc[m][0][0], wv[m], a.x);
          dmma(acc[m][0][1], wv[m], a.y);
          dmma(acc[m][1][0], wv[m], c.x);
          dmma(acc[m][1][1], wv[m], c.y);
        }
      }
      __syncwarp();
      if (lane == 0 && DBG != 3) mbar_arrive(&sm.empty[stage]);
      if (++stage == NST) {
        stage = 0;
        phase ^= 1;
      }
    }

    if (PK == 1 && HS == 2 && pack == 1) {
      // unpack: lane (gq = j, tq) collects the three weight types of harmonic j from rows 3 j + x of the packed
      // M-tile (lanes 4 (3 j + x) + tq); lanes with gq >= 2 hold no harmonic of the summed range and are skipped
      // by the epilogue
#pragma unroll
      for (int x = 2; x >= 0; x--) {
        const int src = (4 * (3 * gq + x) + tq) & 31;
#pragma unroll
        for (int q = 0; q < 8; q++)
          (&acc[x % MT][0][0][0])[q] = __shfl_sync(0xffffffffu, (&acc[0][0][0][0])[q], src);
      }
    }
    double pp_[2][2];
    {
      const int ipar0 = nt * TW + 16 * warp + 2 * tq + 1;
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int ipar = ipar0 + 8 * j + e;
          pp_[j][e] = (ipar <= npar - 1) ? ppar[ipar] : 0.0;
        }
    }
    if (LATE && ntl == 0) {
      if (threadIdx.x == 0) lat_stamp(g, 3);
      fetch_plan();
      if (threadIdx.x == 0) lat_stamp(g, 5);
    }
    // ------------------------------------------------------------ epilogue of this p_par tile
    // this thread: harmonics n0 + 8 h + gq (HS = 1: h = 0,1; HS = 2: h = hsel), columns ipar0 + 8 j + e
    const int ipar0 = nt * TW + 16 * warp + 2 * tq + 1;
    if (STORE) {
      // k-hoisted tables (alps_b200_set_mode(1)): layout [n][ipar-1][GAa, GBa, GAb, GBb, GAc, GBc]
      double* gt = P.gtab[tile.s];
#pragma unroll
      for (int h = 0; h < HP; h++) {
        const int nabs = tile.n0 + 8 * (HS == 2 ? hsel : h) + gq;
        if (nabs > nhi_shard || nabs < nlo_shard) continue;
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int ipar = ipar0 + 8 * j + e;
            if (ipar > npar - 1) continue;
            double2* o = reinterpret_cast<double2*>(gt + ((size_t)nabs * (npar - 1) + (ipar - 1)) * 6);
#pragma unroll
            for (int x = 0; x < 3; x++) o[x] = make_double2(acc[HP * x + h][0][j][e], acc[HP * x + h][1][j][e]);
          }
      }
      continue;
    }
    if (DBG == 2) {   // no epilogue
      double t = 0.0;
#pragma unroll
      for (int m = 0; m < MT; m++)
#pragma unroll
        for (int q = 0; q < 8; q++) t += (&acc[m][0][0][0])[q];
      mine[0] += t;
      continue;
    }
    // G = om * GA + GB: real part into the C' slot, imaginary part into the A' slot
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const double ga = (&acc[m][0][0][0])[q];
        (&acc[m][1][0][0])[q] = fma(omr, ga, (&acc[m][1][0][0])[q]);
        (&acc[m][0][0][0])[q] = omi * ga;
      }
    double Sv[24 * HP];   // [h][sign][12]
#pragma unroll
    for (int q = 0; q < 24 * HP; q++) Sv[q] = 0.0;
#pragma unroll
    for (int h = 0; h < HP; h++) {
      const int nabs = tile.n0 + 8 * (HS == 2 ? hsel : h) + gq;
      if (nabs > nhi_shard || nabs < nlo_shard) continue;
#pragma unroll
      for (int sg = 0; sg < 2; sg++) {
        if (nabs == 0 && sg == 1) continue;
        const size_t item = item0 + 2 * nabs + sg;
        const PlanEntry pe = sm.plan[nabs - tile.n0][sg];
        if (!(pe.flags & PLAN_ACTIVE)) continue;
        const double nq = (sg ? -1.0 : 1.0) * (double)nabs * qs;
        double* S = &Sv[(h * 2 + sg) * 12];
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int ipar = ipar0 + 8 * j + e;
            if (ipar > npar - 1) continue;
            const double w = range_w(ipar, pe.lo1, pe.hi1) + range_w(ipar, pe.lo2, pe.hi2);
            const double p = pp_[j][e];
            if (w != 0.0) {
              // 1/den with den = ms om - kpar p_par - n qs   (resU, src/ALPS_fns.f90:1591-1592)
              const double dr = ms * omr - kpar * p - nq, di = ms * omi;
              const double t = w / (dr * dr + di * di);
              const cd R = mk(dr * t, -di * t);
              const cd Va = R * mk(acc[0 * HP + h][1][j][e], acc[0 * HP + h][0][j][e]);
              const cd Vb = R * mk(acc[1 * HP + h][1][j][e], acc[1 * HP + h][0][j][e]);
              const cd Vc = R * mk(acc[2 * HP + h][1][j][e], acc[2 * HP + h][0][j][e]);
              const double p2 = p * p;
              S[0] += Va.x;       S[1] += Va.y;        // sum U J^2
              S[2] += p * Va.x;   S[3] += p * Va.y;    // sum U J^2 p_par
              S[4] += p2 * Va.x;  S[5] += p2 * Va.y;   // sum U J^2 p_par^2
              S[6] += Vb.x;       S[7] += Vb.y;        // sum U p_perp J J'
              S[8] += p * Vb.x;   S[9] += p * Vb.y;    // sum U p_perp J J' p_par
              S[10] += Vc.x;      S[11] += Vc.y;       // sum U p_perp^2 J'^2
            }
            if (pe.flags & PLAN_NEAR) {
              int jw = ipar - (pe.ipar_res - M_I - 2);
              if (jw < 0 || jw >= WIN) jw = (ipar <= 3) ? WIN + ipar - 1 : -1;   // nodes 1..3: funct_g fallback
              if (jw >= 0) {
                double* gw = P.gwin + (item * WINX + jw) * 6;
#pragma unroll
                for (int x = 0; x < 3; x++) {
                  gw[2 * x] = acc[HP * x + h][1][j][e];
                  gw[2 * x + 1] = acc[HP * x + h][0][j][e];
                }
              }
            }
          }
      }
    }
    // recursive halving over the 4 lanes that share gq (lane bits 1, 0): the lane with tq = 2 h + sg
    // ends with the 12 sums of (harmonic h, sign sg)  [HS = 2: sign tq >> 1, half tq & 1 of its 12 sums]
#pragma unroll
    for (int step = 0; step < 2; step++) {
      const int N = (12 * HP) >> step, mask = 2 >> step;
      const bool up = (lane & mask) != 0;
#pragma unroll
      for (int i = 0; i < 12 * HP; i++) {
        if (i < N) {
          const double send = up ? Sv[i] : Sv[i + N];
          const double keep = up ? Sv[i + N] : Sv[i];
          Sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NSUM; q++) mine[q] += Sv[q];
  }

  if (STORE) return;
  // ---------------------------------------------------------------- write the moment sums
  {
    double* red = (HS == 2) ? &sm.red[warp][8 * hsel + gq][tq >> 1][6 * (tq & 1)]
                            : &sm.red[warp][8 * (tq >> 1) + gq][tq & 1][0];
#pragma unroll
    for (int q = 0; q < NSUM; q++) red[q] = mine[q];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MMA_NH * 24; i += 32 * NW * HS) {
    const int nn = i / 24, sg = (i % 24) / 12, q = i % 12;
    const int nabs = tile.n0 + nn;
    if (nabs > nhi_shard || nabs < nlo_shard) continue;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NW; w++) t += sm.red[w][nn][sg][q];
    P.Sbulk[((item0 + 2 * nabs + sg) * nsplit + jsplit) * 12 + q] = t;
  }
  if (LATE) {
    // single-omega chain: k_chi_assemble, already resident, starts its bulk sums when all CTAs have got here
    __syncthreads();
    if (threadIdx.x == 0) {
      if (P.done_ctr) {
        __threadfence();
        atomicAdd(P.done_ctr, 1);
      }
      lat_stamp(g, 7);
      lat_stamp(g, 4);
    }
  }
}

template <int NW, int HS, int KSTG, int NST, bool STORE, int DBG = 0, int PK = 0>
static cudaError_t launch_mma_one(const QuadParams& P, cudaStream_t st) {
  static PerDeviceOnce once;
  const size_t smem = sizeof(MmaSmem<NW, KSTG, NST>) + 128;
  if (once.first()) {
    cudaError_t e = cudaFuncSetAttribute(k_quad_mma<NW, HS, KSTG, NST, STORE, DBG, PK>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  return launch_chain(k_quad_mma<NW, HS, KSTG, NST, STORE, DBG, PK>, dim3(P.n_om * P.ntiles * P.nsplit),
                      dim3(32 * NW * HS), smem, st, P);
}
template <int NW, int HS, int KSTG, int NST>
static cudaError_t launch_mma_variant(const QuadParams& P, bool store, cudaStream_t st) {
  return store ? launch_mma_one<NW, HS, KSTG, NST, true>(P, st) : launch_mma_one<NW, HS, KSTG, NST, false>(P, st);
}

// variant ids >= 9; the tile width (QuadVariant::bn) fixes the fragment layout built by set_k
cudaError_t launch_quad_mma(const QuadParams& P, int variant, bool store, cudaStream_t st) {
  if (P.n_om <= 0 || P.ntiles <= 0) return cudaSuccess;
  switch (variant) {
    case 12: return launch_mma_variant<4, 1, 8, 2>(P, store, st);   // 2 CTAs of 4 warps per SM
    case 15: {   // 16 warps: 4 per SM sub-partition
      // throughput batches: the species' last tiles with a (nearly) empty upper harmonic group (the last P.ntiles_rem
      // entries of P.tiles) go through the PK = 1 instantiation in a second launch
      if (!store && P.tile_major && P.ntiles_rem > 0 && P.ntiles_rem <= P.ntiles) {
        QuadParams A = P, B = P;
        A.ntiles = P.ntiles - P.ntiles_rem;
        B.tiles = P.tiles + A.ntiles;
        B.ntiles = P.ntiles_rem;
        cudaError_t e = A.ntiles > 0 ? launch_mma_one<8, 2, 8, 2, false>(A, st) : cudaSuccess;
        if (e != cudaSuccess) return e;
        return launch_mma_one<8, 2, 8, 2, false, 0, 1>(B, st);
      }
      return launch_mma_variant<8, 2, 8, 2>(P, store, st);
    }
#ifdef ALPS_QUAD_DEBUG
    case 91: return launch_mma_one<8, 2, 8, 2, false, 1>(P, st);
    case 92: return launch_mma_one<8, 2, 8, 2, false, 2>(P, st);
    case 93: return launch_mma_one<8, 2, 8, 2, false, 3>(P, st);
#endif
    case 16: return launch_mma_variant<8, 2, 4, 4>(P, store, st);
    case 17: return launch_mma_variant<4, 2, 8, 2>(P, store, st);   // 2 CTAs of 8 warps per SM
    case 20: return launch_mma_variant<2, 2, 8, 5>(P, store, st);   // latency: 32-column tiles, 4 warps, 4 stages in flight (api.cu LAT_VARIANT)
    default: return launch_mma_variant<8, 1, 8, 2>(P, store, st);   // 9: one 8-warp CTA per SM (2 warps per sub-partition)
  }
}

}  // namespace alps
