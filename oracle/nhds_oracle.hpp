// TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's NHDS module for species flagged
// use_bM -- closed-form susceptibility of bi-Maxwellian / cold species, O(nmax) per omega:
//   calc_chi      src/ALPS_NHDS.f90:59-242      calc_ypsilon  :250-375     calc_chi_cold :379-464
//   dispfunct     :492-533                      WOFZ          :536-745 (ACM Algorithm 680)
//   BESSI/BESSI0/BESSI1  :750-865 (exp(-x)-scaled modified Bessel functions)
// The product computes the same thing on the device (alps_b200/csrc/nhds_kernel.cu); this file is the
// checker: tests/test_nhds.py pins it against scipy's Faddeeva / Bessel functions, the GPU parity tests
// compare k_nhds with it.  Parity status: the reference ships no NHDS golden vector -> unpinned by the
// reference itself, cross-checked only (scipy, 2e-6: BESSI is the 1e-7 Numerical-Recipes polynomial).
#pragma once
#include <cmath>
#include <complex>

namespace nhds {

typedef std::complex<double> cplx;

struct Params {   // &bM_spec_j (src/ALPS_io.f90:342-372) + the species constants
  int bMnmaxs = 500;
  double bMBessel_zeros = 1.e-50, bMbetas = 1.0, bMalphas = 1.0, bMpdrifts = 0.0;
  double ns = 1.0, qs = 1.0, ms = 1.0;
  bool set = false;
};

inline double bessi0(double X) {
  const double P1 = 1.0, P2 = 3.5156229, P3 = 3.0899424, P4 = 1.2067492, P5 = 0.2659732, P6 = 0.360768e-1, P7 = 0.45813e-2;
  const double Q1 = 0.39894228, Q2 = 0.1328592e-1, Q3 = 0.225319e-2, Q4 = -0.157565e-2, Q5 = 0.916281e-2,
               Q6 = -0.2057706e-1, Q7 = 0.2635537e-1, Q8 = -0.1647633e-1, Q9 = 0.392377e-2;
  double AX = std::fabs(X);
  if (AX < 3.75) {
    double Y = (X / 3.75) * (X / 3.75);
    return (P1 + Y * (P2 + Y * (P3 + Y * (P4 + Y * (P5 + Y * (P6 + Y * P7)))))) * std::exp(-AX);
  }
  double Y = 3.75 / AX, BX = 1.0 / std::sqrt(AX);
  AX = Q1 + Y * (Q2 + Y * (Q3 + Y * (Q4 + Y * (Q5 + Y * (Q6 + Y * (Q7 + Y * (Q8 + Y * Q9)))))));
  return AX * BX;
}

inline double bessi1(double X) {
  const double P1 = 0.5, P2 = 0.87890594, P3 = 0.51498869, P4 = 0.15084934, P5 = 0.2658733e-1, P6 = 0.301532e-2,
               P7 = 0.32411e-3;
  const double Q1 = 0.39894228, Q2 = -0.3988024e-1, Q3 = -0.362018e-2, Q4 = 0.163801e-2, Q5 = -0.1031555e-1,
               Q6 = 0.2282967e-1, Q7 = -0.2895312e-1, Q8 = 0.1787654e-1, Q9 = -0.420059e-2;
  double AX = std::fabs(X);
  if (AX < 3.75) {
    double Y = (X / 3.75) * (X / 3.75);
    return X * (P1 + Y * (P2 + Y * (P3 + Y * (P4 + Y * (P5 + Y * (P6 + Y * P7)))))) * std::exp(-AX);
  }
  double Y = 3.75 / AX, BX = 1.0 / std::sqrt(AX);
  AX = Q1 + Y * (Q2 + Y * (Q3 + Y * (Q4 + Y * (Q5 + Y * (Q6 + Y * (Q7 + Y * (Q8 + Y * Q9)))))));
  return AX * BX;
}

inline double bessi(int N, double X) {
  const int IACC = 40, IBIGNO = 1024 / 2;   // maxexponent(x)/2
  if (N == 0) return bessi0(X);
  if (N == 1) return bessi1(X);
  if (X == 0.0) return 0.0;
  double TOX = 2.0 / X, BIP = 0.0, BI = 1.0, R = 0.0, BIM;
  int M = 2 * (N + (int)sqrtf((float)(IACC * N)));
  for (int J = M; J >= 1; J--) {
    BIM = BIP + (double)J * TOX * BI;
    BIP = BI;
    BI = BIM;
    int ex;
    std::frexp(BI, &ex);
    if (ex > IBIGNO) {
      BI = std::ldexp(BI, -IBIGNO);
      BIP = std::ldexp(BIP, -IBIGNO);
      R = std::ldexp(R, -IBIGNO);
    }
    if (J == N) R = BIP;
  }
  return bessi0(X) * (R / BI);
}
inline double besselI(int n, double x) { return n < 0 ? bessi(-n, x) : bessi(n, x); }

// WOFZ: Faddeeva function, ACM Algorithm 680 as the reference carries it (including its
// default-REAL literals 6.3, 4.4, 0.85, 1.88, which are single precision there).
inline void wofz(double XI, double YI, double& U, double& V, bool& FLAG) {
  const double FACTOR = 1.12837916709551257388, RMAXREAL = 0.5e+154, RMAXEXP = 708.503061461606,
               RMAXGONI = 3.53711887601422e+15;
  FLAG = false;
  U = V = 0.0;
  double XABS = std::fabs(XI), YABS = std::fabs(YI);
  double X = XABS / (double)6.3f, Y = YABS / (double)4.4f;
  if (XABS > RMAXREAL || YABS > RMAXREAL) {
    FLAG = true;
    return;
  }
  double QRHO = X * X + Y * Y, XABSQ = XABS * XABS, XQUAD = XABSQ - YABS * YABS, YQUAD = 2 * XABS * YABS;
  double U2 = 0.0, V2 = 0.0;
  const bool A = QRHO < 0.085264;
  if (A) {
    QRHO = (1 - (double)0.85f * Y) * std::sqrt(QRHO);
    int N = (int)std::lround(6 + 72 * QRHO);
    int J = 2 * N + 1;
    double XSUM = (double)(1.0f / (float)J), YSUM = 0.0, XAUX;   // 1.0/J is a default-REAL division
    for (int I = N; I >= 1; I--) {
      J = J - 2;
      XAUX = (XSUM * XQUAD - YSUM * YQUAD) / I;
      YSUM = (XSUM * YQUAD + YSUM * XQUAD) / I;
      XSUM = XAUX + (double)(1.0f / (float)J);
    }
    double U1 = -FACTOR * (XSUM * YABS + YSUM * XABS) + 1.0, V1 = FACTOR * (XSUM * XABS - YSUM * YABS);
    double DAUX = std::exp(-XQUAD);
    U2 = DAUX * std::cos(YQUAD);
    V2 = -DAUX * std::sin(YQUAD);
    U = U1 * U2 - V1 * V2;
    V = U1 * V2 + V1 * U2;
  } else {
    double H = 0.0, H2 = 0.0, QLAMBDA = 0.0;
    int KAPN = 0, NU;
    if (QRHO > 1.0) {
      QRHO = std::sqrt(QRHO);
      NU = (int)(3 + (1442 / (26 * QRHO + 77)));
    } else {
      QRHO = (1 - Y) * std::sqrt(1 - QRHO);
      H = (double)1.88f * QRHO;
      H2 = 2 * H;
      KAPN = (int)std::lround(7 + 34 * QRHO);
      NU = (int)std::lround(16 + 26 * QRHO);
    }
    const bool B = H > 0.0;
    if (B) QLAMBDA = std::pow(H2, KAPN);
    double RX = 0.0, RY = 0.0, SX = 0.0, SY = 0.0, TX, TY, C;
    for (int N = NU; N >= 0; N--) {
      int NP1 = N + 1;
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
    if (YABS == 0.0) U = std::exp(-XABS * XABS);
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
      double W1 = 2 * std::exp(XQUAD);
      U2 = W1 * std::cos(YQUAD);
      V2 = -W1 * std::sin(YQUAD);
    }
    U = U2 - U;
    V = V2 - V;
    if (XI > 0.0) V = -V;
  } else if (XI < 0.0) {
    V = -V;
  }
}

inline cplx dispfunct(cplx zeta, bool kpos) {
  const cplx uniti(0.0, 1.0);
  const double M_PI_ = 4.0 * std::atan(1.0);
  double U, V;
  bool flag;
  if (kpos) {
    wofz(zeta.real(), zeta.imag(), U, V, flag);
    return uniti * std::sqrt(M_PI_) * (U + uniti * V);
  }
  wofz(-zeta.real(), -zeta.imag(), U, V, flag);
  return -uniti * std::sqrt(M_PI_) * (U + uniti * V);
}

// Y(3,3) column-major: Y[i + 3*k]
inline void calc_ypsilon(cplx* Y, const Params& p, int n, double kz, double kperp, cplx x, bool kperp_norm) {
  const cplx uniti(0.0, 1.0);
  const bool kpos = !(kz < 0.0);
  const double Omega = p.qs / p.ms, vtherm = std::sqrt(p.bMbetas / (p.ns * p.ms)), vdrift = p.bMpdrifts / p.ms;
  const double al = p.bMalphas;
  const cplx zeta = (x - kz * vdrift - 1.0 * n * Omega) / (kz * vtherm);
  const cplx resfac = x - kz * vdrift - 1.0 * n * Omega;
  const double z = 0.5 * (kperp * vtherm / Omega) * (kperp * vtherm / Omega) * al;
  const double zp = 0.5 * (vtherm / Omega) * (vtherm / Omega) * al;
  const cplx Z = dispfunct(zeta, kpos);
  cplx An = (al - 1.0);
  An = An + (1.0 / (kz * vtherm)) * (al * resfac + 1.0 * n * Omega) * Z;
  cplx Bn = (al * (x - 1.0 * n * Omega) - (kz * vdrift - 1.0 * n * Omega)) / kz;
  Bn = Bn + ((x - 1.0 * n * Omega) * (al * resfac + 1.0 * n * Omega) / (kz * kz * vtherm)) * Z;
  const double BInz = 1.0 * besselI(n >= 0 ? n : -n, z);
  const double dB = 5.e-1 * (besselI(n + 1, z) + besselI(n - 1, z));
  const double dn = (double)n, nn2 = 1.0 * (n * n);
  auto y = [&](int i, int k) -> cplx& { return Y[(i - 1) + 3 * (k - 1)]; };
  if (kperp_norm) {
    y(1, 1) = nn2 * BInz * An / z;
    y(1, 2) = -uniti * dn * (BInz - dB) * An;
    y(1, 3) = kperp * dn * BInz * Bn / (Omega * z);
    y(2, 1) = uniti * dn * (BInz - dB) * An;
    y(2, 2) = (nn2 * BInz / z + 2.0 * z * BInz - 2.0 * z * dB) * An;
    y(2, 3) = uniti * kperp * (BInz - dB) * Bn / Omega;
    y(3, 1) = kperp * BInz * dn * Bn / (Omega * z);
    y(3, 2) = -uniti * kperp * (BInz - dB) * Bn / Omega;
    y(3, 3) = 2.0 * (x - 1.0 * n * Omega) * BInz * Bn / (kz * vtherm * vtherm * al);
  } else {
    const double k2 = kperp * kperp;
    y(1, 1) = nn2 * BInz * An / zp;
    y(1, 2) = -uniti * dn * (BInz - dB) * An * k2;
    y(1, 3) = kperp * dn * BInz * Bn / (Omega * zp);
    y(2, 1) = uniti * dn * (BInz - dB) * An * k2;
    y(2, 2) = (nn2 * BInz / zp + k2 * 2.0 * z * BInz - k2 * 2.0 * z * dB) * An;
    y(2, 3) = uniti * k2 * kperp * (BInz - dB) * Bn / Omega;
    y(3, 1) = kperp * BInz * dn * Bn / (Omega * zp);
    y(3, 2) = -uniti * k2 * kperp * (BInz - dB) * Bn / Omega;
    y(3, 3) = k2 * 2.0 * (x - 1.0 * n * Omega) * BInz * Bn / (kz * vtherm * vtherm * al);
  }
}

// chi(3,3), chi_low(3,3,-1:1) column-major.  Returns 0, or 1 if kperp_norm=.false. with a cold species
// (the reference stops there).
inline int calc_chi(cplx* chi, cplx* chi_low, const Params& p, double kz, double kperp, cplx x, bool kperp_norm) {
  const cplx uniti(0.0, 1.0);
  for (int i = 0; i < 9; i++) chi[i] = 0.0;
  for (int i = 0; i < 27; i++) chi_low[i] = 0.0;
  const double Omega = p.qs / p.ms, ell = std::sqrt(p.ms / (p.ns * p.qs * p.qs));
  const double vtherm = std::sqrt(p.bMbetas / (p.ns * p.ms)), vdrift = p.bMpdrifts / p.ms;
  auto c = [&](int i, int k) -> cplx& { return chi[(i - 1) + 3 * (k - 1)]; };
  if (p.bMbetas == 0.0) {
    if (!kperp_norm) return 1;
    const cplx xd = x - kz * vdrift;
    const cplx dispR = -(1.0 / (ell * ell)) * xd / (xd + Omega), dispL = -(1.0 / (ell * ell)) * xd / (xd - Omega);
    cplx dispP = (x * x / (xd * xd)) + ((kperp * vdrift) * (kperp * vdrift) / (xd * xd - Omega * Omega));
    dispP = -(1.0 / (ell * ell)) * dispP;
    const cplx dispJ = -(1.0 / (ell * ell)) * kperp * vdrift * xd / (xd * xd - Omega * Omega);
    const cplx dispM = uniti * (1.0 / (ell * ell)) * kperp * vdrift * Omega / (xd * xd - Omega * Omega);
    c(1, 1) = (dispR + dispL) / 2.0;
    c(1, 2) = -uniti * (dispR - dispL) / 2.0;
    c(1, 3) = dispJ;
    c(2, 1) = uniti * (dispR - dispL) / 2.0;
    c(2, 2) = (dispR + dispL) / 2.0;
    c(2, 3) = dispM;
    c(3, 1) = dispJ;
    c(3, 2) = -dispM;
    c(3, 3) = dispP;
    return 0;
  }
  const double z = 0.5 * (kperp * vtherm / Omega) * (kperp * vtherm / Omega) * p.bMalphas;
  cplx Y[9], Y0[9], Y1[9], Yn1[9], Ynew[9];
  for (int i = 0; i < 9; i++) Y[i] = Y0[i] = Y1[i] = Yn1[i] = 0.0;
  int nmaxrun = p.bMnmaxs, n = 0;
  for (bool run = true; run; n++)
    if (n >= p.bMnmaxs || besselI(n, z) < p.bMBessel_zeros) {
      nmaxrun = n;
      run = false;
    }
  for (n = -nmaxrun; n <= nmaxrun; n++) {
    calc_ypsilon(Ynew, p, n, kz, kperp, x, kperp_norm);
    for (int i = 0; i < 9; i++) {
      Y[i] += Ynew[i];
      if (n == 0) Y0[i] = Ynew[i];
      if (n == 1) Y1[i] += Ynew[i];
      if (n == -1) Yn1[i] += Ynew[i];
    }
  }
  const double l2 = ell * ell;
  const cplx drift33 = (kperp_norm ? 1.0 : kperp * kperp) * 2.0 * x * vdrift / (l2 * kz * vtherm * vtherm * p.bMalphas);
  for (int i = 0; i < 9; i++) {
    chi[i] = Y[i] / l2;
    chi_low[i + 9 * 1] = Y0[i] / l2;    // m = 0
    chi_low[i + 9 * 2] = Y1[i] / l2;    // m = +1
    chi_low[i + 9 * 0] = Yn1[i] / l2;   // m = -1
  }
  chi[8] = drift33 + Y[8] / l2;
  chi_low[8 + 9] = Y0[8] / l2 + drift33;
  return 0;
}

}  // namespace nhds
