// alps_b200: shared device-side definitions (sm_100a only).
//
// Data layout in HBM (per table species s; everything FP64):
//   pperp[s][0..nperp], ppar[s][0..npar]        separable uniform grid (validated at upload)
//   A [s][iperp-1][ipar-1]  = qs * df0/dp_perp                     row pitch ldp (ipar contiguous)
//   Cp[s][iperp-1][ipar-1]  = qs * (kpar/ms) (p_perp df0/dp_par - p_par df0/dp_perp)   (per k)
//   J [s][n+1][iperp]       = BESSJ(n, kperp p_perp/qs), n = -1..nhi+1, iperp = 0..nperp  (per k)
//   W [s][iperp-1][3*n + x] = w_perp(iperp) * {J_n^2, p_perp J_n J_n', p_perp^2 J_n'^2}     (per k)
// Per omega of a batch:
//   plan [item]             resonance descriptor of (species, |n|, sign)
//   Sbulk[item][6]          complex p_par-moment sums of the regular trapezoid quadrature
//   Sres [item][6]          complex near-pole + Landau contributions (final units)
//   gwin [item][WIN+3][3]   complex p_perp-sums at the p_par nodes around a resonance (+ nodes 1..3)
// item = omega * NI + item_base[s] + 2*|n| + sign, NI = sum_s 2 (nhi_s + 1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace alps {

constexpr int MAXSPEC = 8;
constexpr int MAXFITS = 8;

// ---- tile configuration of the quadrature kernel (see quad_kernel.cu)
constexpr int BN = 128;          // p_par columns per tile (harmonics per tile, rows per stage and
                                 // stage count are template parameters of k_quad: QuadVariant)

// ---- complex helpers (double2: x = re, y = im)
typedef double2 cd;
__host__ __device__ inline cd mk(double r, double i) { return make_double2(r, i); }
__host__ __device__ inline cd operator+(cd a, cd b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ inline cd operator-(cd a, cd b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ inline cd operator-(cd a) { return mk(-a.x, -a.y); }
__host__ __device__ inline cd operator*(cd a, cd b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__host__ __device__ inline cd operator*(double s, cd a) { return mk(s * a.x, s * a.y); }
__host__ __device__ inline cd operator*(cd a, double s) { return mk(s * a.x, s * a.y); }
__host__ __device__ inline cd operator/(cd a, double s) { return mk(a.x / s, a.y / s); }
__host__ __device__ inline cd& operator+=(cd& a, cd b) { a.x += b.x; a.y += b.y; return a; }
__host__ __device__ inline cd& operator-=(cd& a, cd b) { a.x -= b.x; a.y -= b.y; return a; }
__host__ __device__ inline cd cmul_i(cd a) { return mk(-a.y, a.x); }   // i * a
// complex division, Smith's algorithm (what libgcc's __divdc3 does in the common range)
__host__ __device__ inline cd operator/(cd a, cd b) {
  double r, den;
  if (fabs(b.x) >= fabs(b.y)) {
    r = b.y / b.x;
    den = b.x + b.y * r;
    return mk((a.x + a.y * r) / den, (a.y - a.x * r) / den);
  }
  r = b.x / b.y;
  den = b.x * r + b.y;
  return mk((a.x * r + a.y) / den, (a.y * r - a.x) / den);
}
__host__ __device__ inline cd operator/(double a, cd b) { return mk(a, 0.0) / b; }
__host__ __device__ inline double cabs2(cd a) { return hypot(a.x, a.y); }

// ---- per-species constants (device copy)
struct SpeciesDev {
  double ns, qs, ms;
  int relativistic, usebM, ACmethod, n_fits;
  int fit_type[MAXFITS];
  double perp_correction[MAXFITS];
  int logfit, poly_kind, poly_order;
  double poly_log_max;
  int nmax;        // determine_nmax result
  int nhi;         // highest harmonic actually summed (>= nmax when emulating split_processes)
  int nlo_shard, nhi_shard;   // harmonic shard owned by this process (inclusive)
  int item_base;   // first item of this species in the per-omega item list
  int ldp;         // row pitch of A / Cp (doubles)
  int ldw;         // row pitch of W (doubles)
  int ldj;         // row pitch of J (= nperp + 1)
  const double* pperp;
  const double* ppar;
  const double* A;
  const double* Cp;
  const double* J;
  const double* W;
  const double* G;           // k-hoisted p_perp sums [n][ipar-1][GAa,GBa,GAb,GBb,GAc,GBc] (mode 1 only)
  const double* T;           // their weighted p_par moments [n][ipar-1][12] (fast_kernel.cu, k_fast_tables; mode 1 only)
  const double* param_fit;   // [iperp][5][maxfits] for this species (repacked)
  const double* poly;        // [iperp][maxorder+1]
  double int_ee;             // omega-independent ee term, src/ALPS_fns.f90:1457-1555 (int_ee_rel if relativistic)
  double dpperp, dppar_abs, dppar_signed;
  // relativistic species: (Gamma, pbar_par) tables of derivative_f0_rel, [igamma][ipparbar] row-major
  const double* C0;          // (qs/ms)(p_perp d_par f0 - p_par d_perp f0), k independent
  const double* grel;        // gamma_rel(igamma), 0..ngamma
  const double* pbrel;       // pparbar_rel(ipparbar), 0..npparbar
  const double* f0_rel;      // -1 outside the sub-luminal cone
  const double* dfg_rel;     // d f0_rel / d Gamma
  const double* dfp_rel;     // d f0_rel / d pbar_par
  const int* cone_lo;        // ipparbar_lower(igamma), src/ALPS_fns_rel.f90:582-596
  const int* cone_up;
  double dgamma, dpparbar;
  // per-k table of BESSJ(n, kperp ms/(vA qs) pperpbar(igamma, ipparbar)), n = 0..nhi+1, layout
  // [n][igamma][ipparbar] (0 outside the sub-luminal cone); NULL: k_rel evaluates BESSJ per point
  const double* Jrel;
};

struct RelTile {
  int s;      // species
  int nabs;   // harmonic |n|
};

// ---- resonance plan of one (omega, species, |n|, sign), src/ALPS_fns.f90:641-745, 941-1006
struct PlanEntry {
  int lo1, hi1, lo2, hi2;    // p_par index ranges of the regular quadrature (hi < lo: empty)
  int flags;                 // PLAN_*
  int ipar_res;
  int upperlimit;
  int pad;
};
constexpr int PLAN_ACTIVE = 1;   // harmonic belongs to this process' shard and is summed
constexpr int PLAN_RES = 2;      // determine_resonances found a resonance
constexpr int PLAN_NEAR = 4;     // near-pole quadrature (not an edge fallback)
constexpr int PLAN_LANDAU = 8;   // Landau residue term needed (Im(om) <= 0)
constexpr int PLAN_REL = 16;     // relativistic species: tensor components come from k_rel

struct GlobalDev {
  int nspec, nperp, npar;
  int ngamma, npparbar;
  int NI;                    // items per omega
  int WIN;                   // nodes per resonance window = 2*M_I + 7
  int WINX;                  // gwin slots per item = WIN + 3: nodes 1..3 follow the window because
                             // funct_g falls back to node 2 when no node is within dp/2 of p
  int M_I, M_P;
  int kperp_norm;
  int maxfits, maxorder;
  double vA, Tlim, kperp, kpar;
  SpeciesDev sp[MAXSPEC];
#ifdef ALPS_LAT_TRACE
  unsigned long long* trace;   // developer build (make trace): time stamps of the single-omega chain
#endif
};

// Time stamps of the latency chain (developer build `make trace`, scripts/lat_trace.py): even slots keep the earliest
// stamp, odd slots the latest.  Compiled out of the product library.
#ifdef ALPS_LAT_TRACE
__device__ __forceinline__ void lat_stamp(const GlobalDev& g, int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  if (slot & 1) atomicMax(g.trace + slot, t);
  else atomicMin(g.trace + slot, t);
}
#else
__device__ __forceinline__ void lat_stamp(const GlobalDev&, int) {}
#endif

// trapezoid weight of node ipar inside [lo,hi]: ends 1 (a single-node range counts twice,
// exactly like integrate() with iparmin == iparmax, src/ALPS_fns.f90:838-862)
__device__ __forceinline__ double range_w(int ipar, int lo, int hi) {
  if (hi < lo || ipar < lo || ipar > hi) return 0.0;
  return (ipar == lo ? 1.0 : 0.0) + (ipar == hi ? 1.0 : 0.0) + ((ipar > lo && ipar < hi) ? 2.0 : 0.0);
}

// Programmatic dependent launch (single-omega graph, api.cu): a kernel launched with the programmatic-serialisation
// attribute may start while its predecessor still runs; pdl_wait() blocks until the predecessor grid has completed
// and its writes are visible, pdl_trigger() lets the successor grid be scheduled.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 1/x for normal, finite x: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps; avoids the
// IEEE-division slow path.  Relative error ~1e-16, far inside the 1e-9 parity budget.
__device__ __forceinline__ double fast_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(y, fma(-x, y, 1.0), y);
  y = fma(y, fma(-x, y, 1.0), y);
  return y;
}
}  // namespace alps
