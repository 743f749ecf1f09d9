// Device restatement of the reference's Bessel routines.  Parity at 1e-9 requires the same
// ~1e-8-accurate approximation the reference uses, not a better Bessel function:
//   BESSJ  src/ALPS_fns_rel.f90:1575-1628   BESSJ0 :1633-1679   BESSJ1 :1684-1721
#pragma once
#include "common.cuh"

namespace alps {

__device__ inline double bessj0_ref(double X) {
  const double P1 = 1.0, P2 = -.1098628627e-2, P3 = .2734510407e-4, P4 = -.2073370639e-5, P5 = .2093887211e-6;
  const double Q1 = -.1562499995e-1, Q2 = .1430488765e-3, Q3 = -.6911147651e-5, Q4 = .7621095161e-6,
               Q5 = -.9349451520e-7;
  const double R1 = 57568490574.0, R2 = -13362590354.0, R3 = 651619640.7, R4 = -11214424.18, R5 = 77392.33017,
               R6 = -184.9052456;
  const double S1 = 57568490411.0, S2 = 1029532985.0, S3 = 9494680.718, S4 = 59272.64853, S5 = 267.8532712,
               S6 = 1.0;
  if (X == 0.0) return 1.0;
  double AX = fabs(X);
  if (AX < 8.0) {
    double Y = X * X;
    double FR = R1 + Y * (R2 + Y * (R3 + Y * (R4 + Y * (R5 + Y * R6))));
    double FS = S1 + Y * (S2 + Y * (S3 + Y * (S4 + Y * (S5 + Y * S6))));
    return FR / FS;
  }
  double Z = 8.0 / AX;
  double Y = Z * Z;
  double XX = AX - .785398164;
  double FP = P1 + Y * (P2 + Y * (P3 + Y * (P4 + Y * P5)));
  double FQ = Q1 + Y * (Q2 + Y * (Q3 + Y * (Q4 + Y * Q5)));
  return sqrt(.636619772 / AX) * (FP * cos(XX) - Z * FQ * sin(XX));
}

__device__ inline double bessj1_ref(double X) {
  const double P1 = 1.0, P2 = .183105e-2, P3 = -.3516396496e-4, P4 = .2457520174e-5, P5 = -.240337019e-6,
               P6 = .636619772;
  const double Q1 = .04687499995, Q2 = -.2002690873e-3, Q3 = .8449199096e-5, Q4 = -.88228987e-6,
               Q5 = .105787412e-6;
  const double R1 = 72362614232.0, R2 = -7895059235.0, R3 = 242396853.1, R4 = -2972611.439, R5 = 15704.48260,
               R6 = -30.16036606;
  const double S1 = 144725228442.0, S2 = 2300535178.0, S3 = 18583304.74, S4 = 99447.43394, S5 = 376.9991397,
               S6 = 1.0;
  double AX = fabs(X);
  if (AX < 8.0) {
    double Y = X * X;
    double FR = R1 + Y * (R2 + Y * (R3 + Y * (R4 + Y * (R5 + Y * R6))));
    double FS = S1 + Y * (S2 + Y * (S3 + Y * (S4 + Y * (S5 + Y * S6))));
    return X * (FR / FS);
  }
  double Z = 8.0 / AX;
  double Y = Z * Z;
  double XX = AX - (double)2.35619491f;   // REAL*4 literal in the reference (line 1713)
  double FP = P1 + Y * (P2 + Y * (P3 + Y * (P4 + Y * P5)));
  double FQ = Q1 + Y * (Q2 + Y * (Q3 + Y * (Q4 + Y * Q5)));
  return sqrt(P6 / AX) * (cos(XX) * FP - Z * sin(XX) * FQ) * copysign(S6, X);
}

__device__ inline double bessj_ref(int N, double X) {
  const int IACC = 40;
  const double BIGNO = 1.e10, BIGNI = 1.e-10;
  if (N == 0) return bessj0_ref(X);
  if (N == 1) return bessj1_ref(X);
  if (X == 0.0) return 0.0;
  double TOX = 2.0 / X;
  if (X > (double)(float)N) {
    double BJM = bessj0_ref(X), BJ = bessj1_ref(X), BJP;
    for (int J = 1; J <= N - 1; J++) {
      BJP = J * TOX * BJ - BJM;
      BJM = BJ;
      BJ = BJP;
    }
    return BJ;
  }
  int M = 2 * ((N + (int)sqrtf((float)(IACC * N))) / 2);
  double R = 0.0, SUM = 0.0, BJP = 0.0, BJ = 1.0, BJM;
  int JSUM = 0;
  for (int J = M; J >= 1; J--) {
    BJM = J * TOX * BJ - BJP;
    BJP = BJ;
    BJ = BJM;
    if (fabs(BJ) > BIGNO) {
      BJ = BJ * BIGNI;
      BJP = BJP * BIGNI;
      R = R * BIGNI;
      SUM = SUM * BIGNI;
    }
    if (JSUM != 0) SUM = SUM + BJ;
    JSUM = 1 - JSUM;
    if (J == N) R = BJP;
  }
  SUM = 2.0 * SUM - BJ;
  return R / SUM;
}

}  // namespace alps
