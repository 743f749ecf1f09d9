// alps_b200: closed-form susceptibility of bi-Maxwellian / cold species (use_bM) on the device.
//
// Replaces the reference's NHDS module as disp() uses it (src/ALPS_fns.f90:344-362):
//   calc_chi      src/ALPS_NHDS.f90:59-242      calc_ypsilon  :250-375     calc_chi_cold :379-464
//   dispfunct     :492-533                      WOFZ          :536-745 (ACM Algorithm 680)
//   BESSI/BESSI0/BESSI1  :750-865 (exp(-x)-scaled modified Bessel functions)
// Not a table quadrature: O(nmax) closed-form terms per (omega, species).  The exp(-z)-scaled I_n(z) depend
// on k only, so set_k tabulates them once (k_nhds_bessel, the literal BESSI recurrences) and finds the
// harmonic cut-off; k_nhds gives one warp to each (omega, species), lanes striding over n in
// [-nmaxrun, nmaxrun], and writes the chi partial rows that k_assemble adds (kernels.h: PARTIAL_PER_SPEC).
// Compiled with -fmad=false like the other literal restatements.
#include "common.cuh"
#include "kernels.h"

namespace alps {

__device__ inline double nh_bessi0(double X) {
  const double P1 = 1.0, P2 = 3.5156229, P3 = 3.0899424, P4 = 1.2067492, P5 = 0.2659732, P6 = 0.360768e-1, P7 = 0.45813e-2;
  const double Q1 = 0.39894228, Q2 = 0.1328592e-1, Q3 = 0.225319e-2, Q4 = -0.157565e-2, Q5 = 0.916281e-2,
               Q6 = -0.2057706e-1, Q7 = 0.2635537e-1, Q8 = -0.1647633e-1, Q9 = 0.392377e-2;
  double AX = fabs(X);
  if (AX < 3.75) {
    const double Y = (X / 3.75) * (X / 3.75);
    return (P1 + Y * (P2 + Y * (P3 + Y * (P4 + Y * (P5 + Y * (P6 + Y * P7)))))) * exp(-AX);
  }
  const double Y = 3.75 / AX, BX = 1.0 / sqrt(AX);
  AX = Q1 + Y * (Q2 + Y * (Q3 + Y * (Q4 + Y * (Q5 + Y * (Q6 + Y * (Q7 + Y * (Q8 + Y * Q9)))))));
  return AX * BX;
}

__device__ inline double nh_bessi1(double X) {
  const double P1 = 0.5, P2 = 0.87890594, P3 = 0.51498869, P4 = 0.15084934, P5 = 0.2658733e-1, P6 = 0.301532e-2,
               P7 = 0.32411e-3;
  const double Q1 = 0.39894228, Q2 = -0.3988024e-1, Q3 = -0.362018e-2, Q4 = 0.163801e-2, Q5 = -0.1031555e-1,
               Q6 = 0.2282967e-1, Q7 = -0.2895312e-1, Q8 = 0.1787654e-1, Q9 = -0.420059e-2;
  double AX = fabs(X);
  if (AX < 3.75) {
    const double Y = (X / 3.75) * (X / 3.75);
    return X * (P1 + Y * (P2 + Y * (P3 + Y * (P4 + Y * (P5 + Y * (P6 + Y * P7)))))) * exp(-AX);
  }
  const double Y = 3.75 / AX, BX = 1.0 / sqrt(AX);
  AX = Q1 + Y * (Q2 + Y * (Q3 + Y * (Q4 + Y * (Q5 + Y * (Q6 + Y * (Q7 + Y * (Q8 + Y * Q9)))))));
  return AX * BX;
}

// BESSI(N, X), src/ALPS_NHDS.f90:750-800: Miller's downward recurrence, rescaled by 2^-512 whenever the
// exponent passes 512, normalised with BESSI0
__device__ inline double nh_bessi(int N, double X) {
  const int IACC = 40, IBIGNO = 1024 / 2;
  if (N == 0) return nh_bessi0(X);
  if (N == 1) return nh_bessi1(X);
  if (X == 0.0) return 0.0;
  const double TOX = 2.0 / X;
  double BIP = 0.0, BI = 1.0, R = 0.0, BIM;
  const int M = 2 * (N + (int)sqrtf((float)(IACC * N)));
  for (int J = M; J >= 1; J--) {
    BIM = BIP + (double)J * TOX * BI;
    BIP = BI;
    BI = BIM;
    int ex;
    frexp(BI, &ex);
    if (ex > IBIGNO) {
      BI = ldexp(BI, -IBIGNO);
      BIP = ldexp(BIP, -IBIGNO);
      R = ldexp(R, -IBIGNO);
    }
    if (J == N) R = BIP;
  }
  return nh_bessi0(X) * (R / BI);
}

// I[n] = BESSI(n, z), n = 0 .. count-1
__global__ void k_nhds_bessel(double z, int count, double* __restrict__ I) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < count) I[n] = nh_bessi(n, z);
}

// WOFZ (Faddeeva function, ACM Algorithm 680) as the reference carries it, including its default-REAL
// literals 6.3, 4.4, 0.85, 1.88 and the default-REAL division 1.0/J
__device__ inline void nh_wofz(double XI, double YI, double& U, double& V, bool& FLAG) {
  const double FACTOR = 1.12837916709551257388, RMAXREAL = 0.5e+154, RMAXEXP = 708.503061461606,
               RMAXGONI = 3.53711887601422e+15;
  FLAG = false;
  U = V = 0.0;
  const double XABS = fabs(XI), YABS = fabs(YI);
  const double X = XABS / (double)6.3f, Y = YABS / (double)4.4f;
  if (XABS > RMAXREAL || YABS > RMAXREAL) {
    FLAG = true;
    return;
  }
  double QRHO = X * X + Y * Y;
  const double XABSQ = XABS * XABS;
  double XQUAD = XABSQ - YABS * YABS;
  const double YQUAD = 2 * XABS * YABS;
  double U2 = 0.0, V2 = 0.0;
  const bool A = QRHO < 0.085264;
  if (A) {
    QRHO = (1 - (double)0.85f * Y) * sqrt(QRHO);
    const int N = (int)llround(6 + 72 * QRHO);
    int J = 2 * N + 1;
    double XSUM = (double)(1.0f / (float)J), YSUM = 0.0, XAUX;
    for (int I = N; I >= 1; I--) {
      J = J - 2;
      XAUX = (XSUM * XQUAD - YSUM * YQUAD) / I;
      YSUM = (XSUM * YQUAD + YSUM * XQUAD) / I;
      XSUM = XAUX + (double)(1.0f / (float)J);
    }
    const double U1 = -FACTOR * (XSUM * YABS + YSUM * XABS) + 1.0, V1 = FACTOR * (XSUM * XABS - YSUM * YABS);
    const double DAUX = exp(-XQUAD);
    U2 = DAUX * cos(YQUAD);
    V2 = -DAUX * sin(YQUAD);
    U = U1 * U2 - V1 * V2;
    V = U1 * V2 + V1 * U2;
  } else {
    double H = 0.0, H2 = 0.0, QLAMBDA = 0.0;
    int KAPN = 0, NU;
    if (QRHO > 1.0) {
      QRHO = sqrt(QRHO);
      NU = (int)(3 + (1442 / (26 * QRHO + 77)));
    } else {
      QRHO = (1 - Y) * sqrt(1 - QRHO);
      H = (double)1.88f * QRHO;
      H2 = 2 * H;
      KAPN = (int)llround(7 + 34 * QRHO);
      NU = (int)llround(16 + 26 * QRHO);
    }
    const bool B = H > 0.0;
    if (B) QLAMBDA = pow(H2, (double)KAPN);
    double RX = 0.0, RY = 0.0, SX = 0.0, SY = 0.0, TX, TY, C;
    for (int N = NU; N >= 0; N--) {
      const int NP1 = N + 1;
      TX = YABS + H + NP1 * RX;
      TY = XABS - NP1 * RY;
      C = 0.5 / (TX * TX + TY * TY);
      RX = C * TX;
      RY = C * TY;
      if (B && N <= KAPN) {
        TX = QLAMBDA + SX;
        SX = RX * TX - RY * SY;
        SY = RY * TX + RX * SY;
        QLAMBDA = QLAMBDA / H2;
      }
    }
    if (H == 0.0) {
      U = FACTOR * RX;
      V = FACTOR * RY;
    } else {
      U = FACTOR * SX;
      V = FACTOR * SY;
    }
    if (YABS == 0.0) U = exp(-XABS * XABS);
  }
  if (YI < 0.0) {
    if (A) {
      U2 = 2 * U2;
      V2 = 2 * V2;
    } else {
      XQUAD = -XQUAD;
      if (YQUAD > RMAXGONI || XQUAD > RMAXEXP) {
        FLAG = true;
        return;
      }
      const double W1 = 2 * exp(XQUAD);
      U2 = W1 * cos(YQUAD);
      V2 = -W1 * sin(YQUAD);
    }
    U = U2 - U;
    V = V2 - V;
    if (XI > 0.0) V = -V;
  } else if (XI < 0.0) {
    V = -V;
  }
}

// plasma dispersion function Z(zeta), src/ALPS_NHDS.f90:492-533
__device__ inline cd nh_dispfunct(cd zeta, bool kpos) {
  const double sqpi = sqrt(4.0 * atan(1.0));
  double U, V;
  bool flag;
  if (kpos) {
    nh_wofz(zeta.x, zeta.y, U, V, flag);
    return cmul_i(mk(sqpi * U, sqpi * V));
  }
  nh_wofz(-zeta.x, -zeta.y, U, V, flag);
  return -cmul_i(mk(sqpi * U, sqpi * V));
}

// the six tensor entries k_assemble needs of Y_n (calc_ypsilon, src/ALPS_NHDS.f90:250-375), in the order of
// the partial rows: (1,1) (2,2) (3,3) (1,2) (1,3) (2,3)
__device__ inline void nh_ypsilon(cd* Y, const NhdsSpec& p, int n, double kz, double kperp, cd x, bool kperp_norm) {
  const bool kpos = !(kz < 0.0);
  const double Omega = p.Omega, vtherm = p.vtherm, vdrift = p.vdrift, al = p.al;
  const double dn = (double)n, nOm = 1.0 * n * Omega;
  const cd resfac = mk(x.x - kz * vdrift - nOm, x.y);
  const cd zeta = resfac / (kz * vtherm);
  const double z = p.z, zp = p.zp;
  const cd Z = nh_dispfunct(zeta, kpos);
  cd An = mk(al - 1.0, 0.0);
  An = An + ((1.0 / (kz * vtherm)) * (al * resfac + mk(nOm, 0.0))) * Z;
  const cd xmn = mk(x.x - nOm, x.y);
  cd Bn = (al * xmn - mk(kz * vdrift - nOm, 0.0)) / kz;
  Bn = Bn + ((xmn * (al * resfac + mk(nOm, 0.0))) / (kz * kz * vtherm)) * Z;
  const int na = n >= 0 ? n : -n;
  const double BInz = p.I[na];
  const double dB = 5.e-1 * (p.I[n + 1 >= 0 ? n + 1 : -(n + 1)] + p.I[n - 1 >= 0 ? n - 1 : -(n - 1)]);
  const double nn2 = 1.0 * (n * n);
  if (kperp_norm) {
    Y[0] = (nn2 * BInz) * An / z;
    Y[1] = (nn2 * BInz / z + 2.0 * z * BInz - 2.0 * z * dB) * An;
    Y[2] = (2.0 * xmn * BInz) * Bn / (kz * vtherm * vtherm * al);
    Y[3] = -cmul_i((dn * (BInz - dB)) * An);
    Y[4] = (kperp * dn * BInz) * Bn / (Omega * z);
    Y[5] = cmul_i((kperp * (BInz - dB)) * Bn) / Omega;
  } else {
    const double k2 = kperp * kperp;
    Y[0] = (nn2 * BInz) * An / zp;
    Y[1] = (nn2 * BInz / zp + k2 * 2.0 * z * BInz - k2 * 2.0 * z * dB) * An;
    Y[2] = ((k2 * 2.0) * xmn * BInz) * Bn / (kz * vtherm * vtherm * al);
    Y[3] = -cmul_i((dn * (BInz - dB)) * An) * k2;
    Y[4] = (kperp * dn * BInz) * Bn / (Omega * zp);
    Y[5] = cmul_i((k2 * kperp * (BInz - dB)) * Bn) / Omega;
  }
}

constexpr int NHDS_WARPS = 4;

// one warp per (omega, species): rows of non-bM species are zeroed (or left alone when accumulating on top of
// caller-supplied external chi)
__device__ __forceinline__ void nhds_warp(const NhdsDev* __restrict__ nd, const double* __restrict__ om, int w, int nspec,
                                          int accumulate, double* __restrict__ ext, int lane) {
  const int iom = w / nspec, s = w % nspec;
  double* o = ext + (size_t)w * PARTIAL_PER_SPEC;
  const NhdsSpec& p = nd->sp[s];
  if (!p.active) {
    if (!accumulate)
      for (int q = lane; q < PARTIAL_PER_SPEC; q += 32) o[q] = 0.0;
    return;
  }
  const cd x = mk(om[2 * iom], om[2 * iom + 1]);
  const double kz = nd->kz, kperp = nd->kperp;
  const bool kperp_norm = nd->kperp_norm != 0;
  const double l2 = p.l2;
  cd tot[6], low[6][3];   // low[c][m + 1], m = -1, 0, 1
#pragma unroll
  for (int c = 0; c < 6; c++) {
    tot[c] = mk(0.0, 0.0);
    low[c][0] = low[c][1] = low[c][2] = mk(0.0, 0.0);
  }
  if (p.cold) {
    // calc_chi_cold, src/ALPS_NHDS.f90:379-464 (kperp_norm=.false. is refused by set_k)
    const double Omega = p.Omega, vdrift = p.vdrift;
    const cd xd = mk(x.x - kz * vdrift, x.y);
    const cd dispR = -(1.0 / l2) * (xd / (xd + mk(Omega, 0.0))), dispL = -(1.0 / l2) * (xd / (xd - mk(Omega, 0.0)));
    const cd den = xd * xd - mk(Omega * Omega, 0.0);
    cd dispP = ((x * x) / (xd * xd)) + (((kperp * vdrift) * (kperp * vdrift)) / den);
    dispP = -(1.0 / l2) * dispP;
    const cd dispJ = (-(1.0 / l2) * kperp * vdrift) * (xd / den);
    const cd dispM = cmul_i(((1.0 / l2) * kperp * vdrift * Omega) / den);
    tot[0] = (dispR + dispL) / 2.0;
    tot[1] = (dispR + dispL) / 2.0;
    tot[2] = dispP;
    tot[3] = -cmul_i(dispR - dispL) / 2.0;
    tot[4] = dispJ;
    tot[5] = dispM;
  } else {
    const int nmaxrun = p.nmaxrun;
    for (int n = -nmaxrun + lane; n <= nmaxrun; n += 32) {
      cd Y[6];
      nh_ypsilon(Y, p, n, kz, kperp, x, kperp_norm);
#pragma unroll
      for (int c = 0; c < 6; c++) {
        tot[c] += Y[c];
        if (n >= -1 && n <= 1) low[c][n + 1] = Y[c];
      }
    }
#pragma unroll
    for (int c = 0; c < 6; c++) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        tot[c].x += __shfl_xor_sync(0xffffffffu, tot[c].x, off);
        tot[c].y += __shfl_xor_sync(0xffffffffu, tot[c].y, off);
#pragma unroll
        for (int m = 0; m < 3; m++) {   // exactly one lane holds a non-zero value
          low[c][m].x += __shfl_xor_sync(0xffffffffu, low[c][m].x, off);
          low[c][m].y += __shfl_xor_sync(0xffffffffu, low[c][m].y, off);
        }
      }
      tot[c] = tot[c] / l2;
#pragma unroll
      for (int m = 0; m < 3; m++) low[c][m] = low[c][m] / l2;
    }
    // drift term of chi_zz, src/ALPS_NHDS.f90:222-232
    const cd drift33 = ((kperp_norm ? 1.0 : kperp * kperp) * 2.0 * p.vdrift / (l2 * kz * p.vtherm * p.vtherm * p.al)) * x;
    tot[2] = drift33 + tot[2];
    low[2][1] = low[2][1] + drift33;
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < 6; c++) {
      if (accumulate) {
        o[2 * c] += tot[c].x;
        o[2 * c + 1] += tot[c].y;
      } else {
        o[2 * c] = tot[c].x;
        o[2 * c + 1] = tot[c].y;
      }
#pragma unroll
      for (int m = 0; m < 3; m++) {
        const int q = 2 * (6 + 3 * c + m);
        if (accumulate) {
          o[q] += low[c][m].x;
          o[q + 1] += low[c][m].y;
        } else {
          o[q] = low[c][m].x;
          o[q + 1] = low[c][m].y;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(32 * NHDS_WARPS)
k_nhds(const NhdsDev* __restrict__ nd, const double* __restrict__ om, int n_om, int nspec, int accumulate,
       double* __restrict__ ext, unsigned long long* done_ctr) {
  const int w = blockIdx.x * NHDS_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w < n_om * nspec) nhds_warp(nd, om, w, nspec, accumulate, ext, lane);
  if (done_ctr) {
    // single-omega chain: this kernel is a branch of its own in the graph; the determinant (k_chi_assemble) waits for the
    // count of finished blocks instead of for a graph edge, so that it can be resident and working meanwhile
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(done_ctr, 1ULL);
    }
  }
}
int nhds_blocks(int n_om, int nspec) { return (n_om * nspec + NHDS_WARPS - 1) / NHDS_WARPS; }

void launch_nhds_bessel(double z, int count, double* I, cudaStream_t st) {
  if (count <= 0) return;
  k_nhds_bessel<<<(count + 63) / 64, 64, 0, st>>>(z, count, I);
}

void launch_nhds(const NhdsDev* nd, const double* om, int n_om, int nspec, int accumulate, double* ext, cudaStream_t st,
                 unsigned long long* done_ctr) {
  if (n_om <= 0) return;
  k_nhds<<<nhds_blocks(n_om, nspec), 32 * NHDS_WARPS, 0, st>>>(nd, om, n_om, nspec, accumulate, ext, done_ctr);
}

}  // namespace alps
