// alps_b200: table set-up kernels (run at upload and whenever k changes; not the hot loop).
#include <algorithm>

#include "bessel.cuh"
#include "kernels.h"

namespace alps {

// Fortran element offsets (0-based) of the reference arrays, SURVEY.md 8.2
__device__ __forceinline__ size_t idx_pp(int nspec, int nperp, int npar, int is0, int iperp, int ipar, int c0) {
  return is0 + (size_t)nspec * (iperp + (size_t)(nperp + 1) * (ipar + (size_t)(npar + 1) * c0));
}
__device__ __forceinline__ size_t idx_df0(int nspec, int nperp, int npar, int is0, int iperp, int ipar, int c0) {
  return is0 + (size_t)nspec * ((iperp - 1) + (size_t)(nperp - 1) * ((ipar - 1) + (size_t)(npar - 1) * c0));
}
__device__ __forceinline__ size_t idx_f0(int nspec, int nperp, int is0, int iperp, int ipar) {
  return is0 + (size_t)nspec * (iperp + (size_t)(nperp + 1) * ipar);
}

// derivative_f0, src/ALPS_fns.f90:96-118: centred differences on interior nodes.
__global__ void k_derivative_f0(const double* __restrict__ f0, const double* __restrict__ pp, double* __restrict__ df0,
                                int nspec, int nperp, int npar) {
  size_t total = (size_t)nspec * (nperp - 1) * (npar - 1);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    int is0 = (int)(t % nspec);
    size_t r = t / nspec;
    int iperp = (int)(r % (nperp - 1)) + 1;
    int ipar = (int)(r / (nperp - 1)) + 1;
    double dperp = (f0[idx_f0(nspec, nperp, is0, iperp + 1, ipar)] - f0[idx_f0(nspec, nperp, is0, iperp - 1, ipar)]) /
                   (pp[idx_pp(nspec, nperp, npar, is0, iperp + 1, ipar, 0)] -
                    pp[idx_pp(nspec, nperp, npar, is0, iperp - 1, ipar, 0)]);
    double dpar = (f0[idx_f0(nspec, nperp, is0, iperp, ipar + 1)] - f0[idx_f0(nspec, nperp, is0, iperp, ipar - 1)]) /
                  (pp[idx_pp(nspec, nperp, npar, is0, iperp, ipar + 1, 1)] -
                   pp[idx_pp(nspec, nperp, npar, is0, iperp, ipar - 1, 1)]);
    df0[idx_df0(nspec, nperp, npar, is0, iperp, ipar, 0)] = dperp;
    df0[idx_df0(nspec, nperp, npar, is0, iperp, ipar, 1)] = dpar;
  }
}

// A = qs * d_perp f0 ; C0 = (qs/ms) (p_perp d_par f0 - p_par d_perp f0)   [numerator of resU,
// src/ALPS_fns.f90:1587-1590, split into its omega-proportional and kpar-proportional parts]
__global__ void k_build_AC(const double* __restrict__ df0, const double* __restrict__ pperp,
                           const double* __restrict__ ppar, double* __restrict__ A, double* __restrict__ C0,
                           int nspec, int nperp, int npar, int is0, double qs, double ms, int ldp) {
  int ipar = blockIdx.x * blockDim.x + threadIdx.x + 1;
  int iperp = blockIdx.y + 1;
  if (ipar > npar - 1 || iperp > nperp - 1) return;
  double a = df0[idx_df0(nspec, nperp, npar, is0, iperp, ipar, 0)];
  double b = df0[idx_df0(nspec, nperp, npar, is0, iperp, ipar, 1)];
  size_t o = (size_t)(iperp - 1) * ldp + (ipar - 1);
  A[o] = qs * a;
  C0[o] = (qs / ms) * (pperp[iperp] * b - ppar[ipar] * a);
}

__global__ void k_scale(const double* __restrict__ in, double* __restrict__ out, double s, size_t n) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
    out[t] = s * in[t];
}

// determine_nmax inner loop, src/ALPS_fns.f90:4017-4037: besselmax(n) = max_iperp |BESSJ(n,z)|
__global__ void k_bessel_max(const double* __restrict__ pperp, int nperp, double kperp, double qs, int n0,
                             double* __restrict__ out) {
  int n = n0 + blockIdx.x;
  double m = 0.0;
  for (int iperp = threadIdx.x; iperp <= nperp; iperp += blockDim.x) {
    double z = kperp * pperp[iperp] / qs;
    m = fmax(m, fabs(bessj_ref(n, z)));
  }
  __shared__ double sm[256];
  sm[threadIdx.x] = m;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

// determine_bessel_array, src/ALPS_fns.f90:4214-4255: J[(n+1)*ldj + iperp], n = -1..nhi+1
__global__ void k_bessel_table(const double* __restrict__ pperp, int nperp, double kperp, double qs, int nhi,
                               double* __restrict__ J, int ldj) {
  int iperp = blockIdx.x * blockDim.x + threadIdx.x;
  int n = (int)blockIdx.y - 1;
  if (iperp > nperp || n > nhi + 1) return;
  double z = kperp * pperp[iperp] / qs;
  J[(size_t)(n + 1) * ldj + iperp] = (n == -1) ? -bessj_ref(1, z) : bessj_ref(n, z);
}

// Bessel weights of the T tensor (int_T, src/ALPS_fns.f90:1601-1707) with the p_perp trapezoid
// weight folded in (integrate, src/ALPS_fns.f90:838-862: iperp=1 counts double because the
// iperp=0 row is skipped, iperp=nperp-1 is the end point).  J^2, J J', J'^2 are identical for
// +n and -n, so one row triple serves both signs.
__global__ void k_build_W(const double* __restrict__ pperp, const double* __restrict__ J, int ldj, int nperp,
                          int nhi, double* __restrict__ W, int ldw) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int iperp = blockIdx.y + 1;
  if (iperp > nperp - 1 || 3 * n + 2 >= ldw) return;
  double wa = 0.0, wb = 0.0, wc = 0.0;
  if (n <= nhi) {
    double wperp = (iperp == nperp - 1) ? 1.0 : 2.0;
    double bj = J[(size_t)(n + 1) * ldj + iperp];
    double bp = (n >= 1) ? 0.5 * (J[(size_t)n * ldj + iperp] - J[(size_t)(n + 2) * ldj + iperp])
                         : -J[(size_t)2 * ldj + iperp];
    double p = pperp[iperp];
    wa = wperp * (bj * bj);
    wb = wperp * (p * (bj * bp));
    wc = wperp * ((p * p) * (bp * bp));
  }
  double* w = W + (size_t)(iperp - 1) * ldw + 3 * n;
  w[0] = wa;
  w[1] = wb;
  w[2] = wc;
}

// Fragment-ordered operands of the DMMA kernel (quad_mma.cu).  Xf[nt][ks][4 tw]: the 4 x tw block of
// rows 4 ks .. 4 ks + 3 and columns tw nt .. tw nt + tw - 1 of scale * X, element (t, 16 w + 8 j + g) at
// 64 w + 2 (4 g + t) + j.  One thread per output element; padding elements are written as zero.
__global__ void k_frag_table(const double* __restrict__ X, int ldp, int nrows, int ncols, double scale,
                             double* __restrict__ Xf, int nks, int tw, size_t total) {
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int in = (int)(o % (size_t)(4 * tw));
    const size_t blk = o / (size_t)(4 * tw);
    const int ks = (int)(blk % nks), nt = (int)(blk / nks);
    const int w = in >> 6, l = (in >> 1) & 31, j = in & 1, g = l >> 2, t = l & 3;
    const int r = 4 * ks + t, c = tw * nt + 16 * w + 8 * j + g;
    Xf[o] = (r < nrows && c < ncols) ? scale * X[(size_t)r * ldp + c] : 0.0;
  }
}

// Wf[hb][ks][192]: weights (same values as k_build_W) of harmonics 16 hb .. 16 hb + 15 on rows
// iperp - 1 = 4 ks + t: element (t, type x, harmonic 16 hb + 8 h + g) at 6 (4 g + t) + 2 x + h.
__global__ void k_build_Wf(const double* __restrict__ pperp, const double* __restrict__ J, int ldj, int nperp,
                           int nhi, double* __restrict__ Wf, int nks, size_t total) {
  for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
    const int in = (int)(o % 192);
    const size_t blk = o / 192;
    const int ks = (int)(blk % nks), hb = (int)(blk / nks);
    const int l = in / 6, m = in % 6, x = m >> 1, h = m & 1, g = l >> 2, t = l & 3;
    const int iperp = 4 * ks + t + 1, n = 16 * hb + 8 * h + g;
    double v = 0.0;
    if (iperp <= nperp - 1 && n <= nhi) {
      double wperp = (iperp == nperp - 1) ? 1.0 : 2.0;
      double bj = J[(size_t)(n + 1) * ldj + iperp];
      double bp = (n >= 1) ? 0.5 * (J[(size_t)n * ldj + iperp] - J[(size_t)(n + 2) * ldj + iperp])
                           : -J[(size_t)2 * ldj + iperp];
      double p = pperp[iperp];
      v = (x == 0) ? wperp * (bj * bj) : (x == 1) ? wperp * (p * (bj * bp)) : wperp * ((p * p) * (bp * bp));
    }
    Wf[o] = v;
  }
}

// int_ee, src/ALPS_fns.f90:1457-1555 (omega independent): one block per species, read from
// the Fortran-layout df0.  Weights: iperp=1 -> 2, interior -> 2, nperp-1 -> 1 ; ipar ends -> 1,
// interior -> 2; the (1,1) corner multiplies d_perp f0 by p_perp where every other term uses
// p_par (lines 1483-1486).
__global__ void k_int_ee(const double* __restrict__ df0, const double* __restrict__ pperp,
                         const double* __restrict__ ppar, int nspec, int nperp, int npar, int is0, double qs,
                         double ms, double dpperp, double dppar, double* __restrict__ out) {
  double acc = 0.0;
  size_t total = (size_t)(nperp - 1) * (npar - 1);
  for (size_t t = threadIdx.x; t < total; t += blockDim.x) {
    int iperp = (int)(t % (nperp - 1)) + 1;
    int ipar = (int)(t / (nperp - 1)) + 1;
    double wperp = (iperp == nperp - 1) ? 1.0 : 2.0;
    double wpar = (ipar == 1 || ipar == npar - 1) ? 1.0 : 2.0;
    double dperp = df0[idx_df0(nspec, nperp, npar, is0, iperp, ipar, 0)];
    double dpar = df0[idx_df0(nspec, nperp, npar, is0, iperp, ipar, 1)];
    double pq = (iperp == 1 && ipar == 1) ? pperp[iperp] : ppar[ipar];
    acc += wperp * wpar * (ppar[ipar] * (dpar * pperp[iperp] - pq * dperp));
  }
  __shared__ double sm[1024];
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double r = sm[0] * 2.0 * 3.14159265358979323846 * qs / ms;
    out[0] = r * dpperp * dppar * 0.25;
  }
}

// --------------------------------------------------------------- launchers
void launch_derivative_f0(const double* f0, const double* pp, double* df0, int nspec, int nperp, int npar,
                          cudaStream_t st) {
  size_t total = (size_t)nspec * (nperp - 1) * (npar - 1);
  int blocks = (int)((total + 255) / 256 < 65535 ? (total + 255) / 256 : 65535);
  k_derivative_f0<<<blocks, 256, 0, st>>>(f0, pp, df0, nspec, nperp, npar);
}
void launch_build_AC(const double* df0, const double* pperp, const double* ppar, double* A, double* C0, int nspec,
                     int nperp, int npar, int is0, double qs, double ms, int ldp, cudaStream_t st) {
  dim3 g((npar - 1 + 127) / 128, nperp - 1);
  k_build_AC<<<g, 128, 0, st>>>(df0, pperp, ppar, A, C0, nspec, nperp, npar, is0, qs, ms, ldp);
}
void launch_scale(const double* in, double* out, double s, size_t n, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  k_scale<<<blocks, 256, 0, st>>>(in, out, s, n);
}
void launch_bessel_max(const double* pperp, int nperp, double kperp, double qs, int n0, int count, double* out,
                       cudaStream_t st) {
  k_bessel_max<<<count, 256, 0, st>>>(pperp, nperp, kperp, qs, n0, out);
}
void launch_bessel_table(const double* pperp, int nperp, double kperp, double qs, int nhi, double* J, int ldj,
                         cudaStream_t st) {
  dim3 g((nperp + 1 + 127) / 128, nhi + 3);
  k_bessel_table<<<g, 128, 0, st>>>(pperp, nperp, kperp, qs, nhi, J, ldj);
}
void launch_build_W(const double* pperp, const double* J, int ldj, int nperp, int nhi, double* W, int ldw,
                    cudaStream_t st) {
  int ncols = ldw / 3 + 1;
  dim3 g((ncols + 63) / 64, nperp - 1);
  k_build_W<<<g, 64, 0, st>>>(pperp, J, ldj, nperp, nhi, W, ldw);
}
void launch_frag_table(const double* X, int ldp, int nrows, int ncols, double scale, double* Xf, int nks, int tw,
                       int ntiles, cudaStream_t st) {
  const size_t total = (size_t)ntiles * nks * 4 * tw;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 32);
  k_frag_table<<<blocks, 256, 0, st>>>(X, ldp, nrows, ncols, scale, Xf, nks, tw, total);
}
void launch_build_Wf(const double* pperp, const double* J, int ldj, int nperp, int nhi, double* Wf, int nks, int nhb,
                     cudaStream_t st) {
  const size_t total = (size_t)nhb * nks * 192;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 32);
  k_build_Wf<<<blocks, 256, 0, st>>>(pperp, J, ldj, nperp, nhi, Wf, nks, total);
}
void launch_int_ee(const double* df0, const double* pperp, const double* ppar, int nspec, int nperp, int npar,
                   int is0, double qs, double ms, double dpperp, double dppar, double* out, cudaStream_t st) {
  k_int_ee<<<1, 1024, 0, st>>>(df0, pperp, ppar, nspec, nperp, npar, is0, qs, ms, dpperp, dppar, out);
}

// polyharmonic_spline evaluation of derivative_f0_rel, src/ALPS_fns_rel.f90:300-331, 407-423: for every point of the
// (Gamma, pbar_par) grid  sum_i w_i phi(r_i) + w_n + w_{n+1} Gamma + w_{n+2} pbar_par  over the n table nodes, with the
// reference's kernel (r >= 1: r^2 log r; 0 < r < 1: r log(r**r); r = 0: 0, lines 385-391), summed in node order.
// One thread per grid point, the nodes staged through shared memory in chunks.
constexpr int TPS_CHUNK = 256;
__global__ void __launch_bounds__(TPS_CHUNK) k_tps_eval(int n, const double* __restrict__ gc,
                                                        const double* __restrict__ pc, const double* __restrict__ w,
                                                        int npts, const double* __restrict__ gx,
                                                        const double* __restrict__ px, double* __restrict__ out) {
  __shared__ double sg[TPS_CHUNK], sp[TPS_CHUNK], sw[TPS_CHUNK];
  const int j = blockIdx.x * TPS_CHUNK + threadIdx.x;
  const double x = j < npts ? gx[j] : 0.0, y = j < npts ? px[j] : 0.0;
  double acc = 0.0;
  for (int i0 = 0; i0 < n; i0 += TPS_CHUNK) {
    const int i = i0 + threadIdx.x;
    __syncthreads();
    sg[threadIdx.x] = i < n ? gc[i] : 0.0;
    sp[threadIdx.x] = i < n ? pc[i] : 0.0;
    sw[threadIdx.x] = i < n ? w[i] : 0.0;
    __syncthreads();
    const int m = min(TPS_CHUNK, n - i0);
    for (int k = 0; k < m; k++) {
      const double dx = x - sg[k], dy = y - sp[k];
      const double r = sqrt(dx * dx + dy * dy);
      double phi = 0.0;
      if (r >= 1.0) phi = r * r * log(r);
      else if (r > 0.0) phi = r * log(pow(r, r));
      acc += sw[k] * phi;
    }
  }
  if (j < npts) out[j] = acc + w[n] + w[n + 1] * x + w[n + 2] * y;
}
void launch_tps_eval(int n, const double* gc, const double* pc, const double* w, int npts, const double* gx,
                     const double* px, double* out, cudaStream_t st) {
  if (npts <= 0) return;
  k_tps_eval<<<(npts + TPS_CHUNK - 1) / TPS_CHUNK, TPS_CHUNK, 0, st>>>(n, gc, pc, w, npts, gx, px, out);
}

}  // namespace alps
