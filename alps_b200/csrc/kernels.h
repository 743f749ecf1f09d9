// alps_b200: host-side declarations of the kernel launchers (internal; the public boundary is
// include/alps_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace alps {

struct QuadTile {
  int s;    // species (0-based)
  int n0;   // first harmonic of the tile
};

struct QuadParams {
  CUtensorMap tmA[MAXSPEC];
  CUtensorMap tmC[MAXSPEC];
  CUtensorMap tmW[MAXSPEC];
  const QuadTile* tiles;
  const GlobalDev* g;
  const double* om;        // 2 * n_om
  const PlanEntry* plan;   // n_om * NI
  double* Sbulk;           // n_om * NI * 12
  double* gwin;            // n_om * NI * WINX * 6
  double* gtab[MAXSPEC];   // k-hoisted tables (STORE launches only)
  int ntiles;
  int ntiles_rem;          // the last ntiles_rem tiles are species' last tiles whose upper 8-harmonic group holds <= 2
                           // harmonics of the summed range (quad_mma.cu, PK instantiation); 0: none / not sorted
  int n_om;
  int nsplit;              // CTAs per (omega, tile) along p_par; Sbulk then holds nsplit partial rows per item
  // DMMA variants (quad_mma.cu): fragment-ordered copies of A', C', W and their padded k-step count
  const double* Af[MAXSPEC];
  const double* Cf[MAXSPEC];
  const double* Wf[MAXSPEC];
  int nks;
  // latency variant: the first tiles and the grid size inline, so that a CTA can issue its first table loads without a
  // round trip to P.tiles / P.g (ntiles_inline = 0: not filled)
  QuadTile tile_inline[16];
  int ntiles_inline, npar_inline;
  int* done_ctr;           // latency variant: bumped by every CTA once its sums are written (nullptr: not used)
  int tile_major;          // block order: 1 = all omegas of a tile are adjacent (concurrent CTAs share the species tables
                           // and W in L2), 0 = all tiles of an omega are adjacent
};
constexpr int MMA_NH = 16;   // harmonics per CTA tile of the DMMA variants

// set by api.cu while it captures the single-omega chain: kernels after the first are launched with the
// programmatic-serialisation attribute (they call pdl_wait() before touching their predecessor's output)
extern bool g_pdl_launch;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl_launch ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Function attributes (opt-in dynamic shared memory) are per DEVICE: a launcher that raises them remembers the devices
// it has done so for (a device group launches the same kernels on several devices from several host threads).
struct PerDeviceOnce {
  unsigned mask = 0;
  // true exactly once per device (benign race: two threads never share a device)
  bool first() {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned bit = 1u << (dev & 31);
    const unsigned old = __atomic_fetch_or(&mask, bit, __ATOMIC_RELAXED);
    return !(old & bit);
  }
};

// set-up (setup_kernels.cu)
void launch_derivative_f0(const double* f0, const double* pp, double* df0, int nspec, int nperp, int npar,
                          cudaStream_t st);
void launch_build_AC(const double* df0, const double* pperp, const double* ppar, double* A, double* C0, int nspec,
                     int nperp, int npar, int is0, double qs, double ms, int ldp, cudaStream_t st);
void launch_scale(const double* in, double* out, double s, size_t n, cudaStream_t st);
void launch_bessel_max(const double* pperp, int nperp, double kperp, double qs, int n0, int count, double* out,
                       cudaStream_t st);
void launch_bessel_table(const double* pperp, int nperp, double kperp, double qs, int nhi, double* J, int ldj,
                         cudaStream_t st);
void launch_build_W(const double* pperp, const double* J, int ldj, int nperp, int nhi, double* W, int ldw,
                    cudaStream_t st);
// fragment-ordered operands of the DMMA quadrature kernel (layouts: quad_mma.cu)
void launch_frag_table(const double* X, int ldp, int nrows, int ncols, double scale, double* Xf, int nks, int tw,
                       int ntiles, cudaStream_t st);
void launch_build_Wf(const double* pperp, const double* J, int ldj, int nperp, int nhi, double* Wf, int nks, int nhb,
                     cudaStream_t st);
void launch_tps_eval(int n, const double* gc, const double* pc, const double* w, int npts, const double* gx,
                     const double* px, double* out, cudaStream_t st);
void launch_int_ee(const double* df0, const double* pperp, const double* ppar, int nspec, int nperp, int npar,
                   int is0, double qs, double ms, double dpperp, double dppar, double* out, cudaStream_t st);

// hot path (quad_kernel.cu, resonant.cu)
struct QuadVariant {
  int id;
  int NH;       // harmonics per CTA tile (rows = 3 NH)
  int BK;       // p_perp rows per pipeline stage
  int stages;
  int bn;       // p_par columns per tile
};
QuadVariant quad_variant(int id);
cudaError_t launch_quad(const QuadParams& P, int variant, bool store, cudaStream_t st);
cudaError_t launch_quad_mma(const QuadParams& P, int variant, bool store, cudaStream_t st);   // variants >= 9
double run_dmma_peak(cudaStream_t st);
// om_stage != nullptr: fused single-block variant (needs plan_fused_ok): om may be pinned host memory, the block
// copies it to om_stage (device) and resets work_count itself
constexpr int PLAN_FUSED_MAX_OM = 8;
// Single-omega chain: slots (doubles) of the pinned host block the chain's first / last kernel read / write (api.cu), and
// the chain's device words (ints): k_plan's completion flag, CTAs of k_quad_mma done, blocks of k_resonant_lat done,
// blocks of k_nhds (side branch of the graph) done (CHAIN_NHDS64).
constexpr int ZC_OM = 0, ZC_D = 16, ZC_ERR = 32, ZC_NHT = 40, ZC_DOUBLES = 64;   // ZC_NHT: see CHAIN_NHDS64
constexpr int CHAIN_PLAN = 0, CHAIN_QUAD = 1, CHAIN_RES = 2, CHAIN_INTS = 8;
// ints 4..5: a 64-bit count of finished k_nhds blocks that is never reset (k_nhds sits at the root of the graph, beside
// k_plan, so nothing could reset it safely): the host writes the count the call will reach into the pinned block
// (ZC_NHT, a double) before it launches the graph, and the determinant waits until the counter has got there.
constexpr int CHAIN_NHDS64 = 4;
static_assert(CHAIN_NHDS64 % 2 == 0 && CHAIN_NHDS64 > CHAIN_RES && CHAIN_NHDS64 + 2 <= CHAIN_INTS,
              "the 64-bit count sits on an 8-byte boundary behind the 32-bit words");
static_assert(ZC_OM + 2 * PLAN_FUSED_MAX_OM <= ZC_D && ZC_D + 2 * PLAN_FUSED_MAX_OM <= ZC_ERR && ZC_ERR + 4 <= ZC_NHT &&
              ZC_NHT < ZC_DOUBLES, "slots of the pinned host block do not overlap");
void launch_plan(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, PlanEntry* plan,
                 int* work, int* work_count, cudaStream_t st, double* om_stage = nullptr, int* plan_flag = nullptr,
                 int* zero_ints = nullptr, int nzero = 0);   // fused variant: words it clears for the chain (k_rel_plan's counts)
bool plan_fused_ok(const GlobalDev& gh, int n_om);
void launch_resonant(const GlobalDev* g, const double* om, int n_om, const PlanEntry* plan, const int* work,
                     const int* work_count, const double* gwin, double* Sres, int* err_flag, double* Spart,
                     cudaStream_t st, int gx = 148, int class_n = 0,   // class_n: omegas of the whole call
                     int* chain = nullptr,    // the chain's device words: flag-driven starts (programmatic launches)
                     int nquad = 0);          // CTAs of k_quad_mma whose count releases the near-pole blocks (0: kernel boundary)
int resonant_lat_blocks(int n_om, int gx);   // grid size of k_resonant_lat for that call
// k_resonant_lat serves the call (its partial rows, not Sres, feed the harmonic sums: pass Spart to launch_chi_*)
bool resonant_lat_class(int n_om, int class_n);
constexpr int RESLAT_GX_TINY = 4, RESLAT_GX_NARROW = 16, RESLAT_GX_WIDE = 148;   // block columns per omega of k_resonant_lat (api.cu adapts)
constexpr int RES_PART_DOUBLES = 11 * 16;   // k_resonant_lat: partial rows per item (LAT_PARTS x LAT_STRIDE)
// chi partial layout per omega: [nspec][PARTIAL_PER_SPEC] doubles (see resonant.cu)
constexpr int PARTIAL_PER_SPEC = 2 * (6 + 18);   // chi(6 modes) + chi_low(6 modes x 3) complex
void launch_chi_partial(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, const PlanEntry* plan,
                        const double* Sbulk, int nsplit, const double* Sres, const double* Spart, double* partial,
                        cudaStream_t st);
void launch_assemble(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, const double* partial,
                     const double* ext_chi, double* D, double* chi0, double* chi0_low, double* wave,
                     cudaStream_t st, const int* err_src = nullptr, int* err_dst = nullptr);
// both in one launch for small batches (one block per omega, one warp per species; bitwise the two-kernel result)
void launch_chi_assemble(const GlobalDev* g, const GlobalDev& gh, const double* om, int n_om, const PlanEntry* plan,
                         const double* Sbulk, int nsplit, const double* Sres, const double* Spart, double* partial,
                         const double* ext_chi, double* D, double* chi0, double* chi0_low, double* wave, cudaStream_t st,
                         const int* err_src = nullptr, int* err_dst = nullptr,
                         int* chain = nullptr, int nquad = 0, int nres = 0,   // early bulk sums (programmatic launches)
                         const double* nh_target = nullptr);   // pinned host word: the k_nhds block count to wait for (0: none)
struct FastItem {
  int s;
  int nabs;
};
// variant 0: k_fast (one thread per omega over the G tables); 1: k_fast_tiled (real moment tables T through a TMA ring)
void launch_fast(const GlobalDev* g, const double* om, int n_om, const FastItem* items, int nitems,
                 const PlanEntry* plan, double* Sbulk, double* gwin, int npar, double kpar, int variant,
                 cudaStream_t st);
void launch_fast_tables(const double* G, const double* ppar, int npar, int nrows, double* T, cudaStream_t st);
void launch_rel(const GlobalDev* g, const double* om, int n_om, const RelTile* tiles, int ntiles, double* Mrel,
                int* err_flag, int nsplit, double* Mpart, int* tickets, cudaStream_t st);
// throughput class of the relativistic species (rel_kernel.cu): rflag n_om*ntiles bytes; rwork 2*n_om*ntiles ints (one
// segment per tile); rcount ntiles ints; rpos 2*n_om*ntiles ints; dpart 2*n_om*ntiles*nsplitB*12 doubles would be the
// worst case -- api.cu sizes it for REL_DPART_ENTRIES resonant entries per omega and falls back to launch_rel beyond
void launch_rel_small(const GlobalDev* g, const double* om, int n_om, const RelTile* tiles, int ntiles, double* Mrel,
                      int* err_flag, unsigned char* rflag, int* rwork, int* rcount, int* rpos, int nsplit, double* Mpart,
                      int* tickets, size_t rows_part_offset, size_t rows_ticket_offset, int sm_count, bool zero_rcount,
                      cudaStream_t st);   // n <= 64 omegas (rel_kernel.cu: k_rel_rows)
int rel_rows_chunks(int ngamma);   // partial rows per resonant entry of launch_rel_small
constexpr int REL_ROWS_MAXTILES = 1024;
void launch_rel_tiled(const GlobalDev* g, const double* om, int n_om, const RelTile* tiles, int ntiles, double* Mrel,
                      int* err_flag, unsigned char* rflag, int* rwork, int* rcount, int* rpos, double* dpart, int nsplitB,
                      int sm_count, cudaStream_t st);
void launch_rel_bessel_table(const double* grel, const double* pbrel, int ng, int npb, double zfac, int nmaxord,
                             double* Jrel, cudaStream_t st);
void launch_int_ee_rel(const double* pbv, const double* dfp, const int* lo, const int* up, int ng, int npb, double qs,
                       double ms, double vA, double dgam, double dpb, double* out, cudaStream_t st);
// use_bM species on the device (nhds_kernel.cu): per-k constants of calc_chi, src/ALPS_NHDS.f90:59-242
struct NhdsSpec {
  int active;        // use_bM species with &bM_spec parameters
  int cold;          // bMbetas == 0: calc_chi_cold
  int nmaxrun;       // harmonic cut-off of the current k (bMnmaxs / bMBessel_zeros rule, :129-141)
  int pad;
  double Omega, vtherm, vdrift, al, z, zp, l2;
  const double* I;   // BESSI(n, z), n = 0 .. nmaxrun + 1
};
struct NhdsDev {
  int kperp_norm, pad;
  double kz, kperp;
  NhdsSpec sp[MAXSPEC];
};
void launch_nhds_bessel(double z, int count, double* I, cudaStream_t st);
void launch_nhds(const NhdsDev* nd, const double* om, int n_om, int nspec, int accumulate, double* ext, cudaStream_t st,
                 unsigned long long* done_ctr = nullptr);   // single-omega chain: every block bumps it (CHAIN_NHDS64)
int nhds_blocks(int n_om, int nspec);       // grid size of k_nhds for that call
// HARMONIC partition of a device group: dst += sum of the peers' partial rows, read through peer memory (multi_gpu.cu)
void launch_reduce_partials(double* dst, const double* const* src, int nsrc, size_t count, cudaStream_t st);
double run_dfma_peak(cudaStream_t st);
double run_dfma_peak_noreuse(cudaStream_t st);

}  // namespace alps
