// TEST INFRASTRUCTURE ONLY (oracle): C entry point of the NHDS restatement (nhds_oracle.hpp).
#include "nhds_oracle.hpp"

extern "C" int oracle_nhds_calc_chi(double ns, double qs, double ms, int bM_nmaxs, double bM_Bessel_zeros,
                                    double bM_betas, double bM_alphas, double bM_pdrifts, double kz, double kperp,
                                    const double x[2], int kperp_norm, double* chi, double* chi_low) {
  nhds::Params p;
  p.ns = ns; p.qs = qs; p.ms = ms; p.bMnmaxs = bM_nmaxs; p.bMBessel_zeros = bM_Bessel_zeros; p.bMbetas = bM_betas;
  p.bMalphas = bM_alphas; p.bMpdrifts = bM_pdrifts; p.set = true;
  nhds::cplx c[9], l[27];
  if (nhds::calc_chi(c, l, p, kz, kperp, nhds::cplx(x[0], x[1]), kperp_norm != 0)) return 1;
  for (int i = 0; i < 9 && chi; i++) { chi[2 * i] = c[i].real(); chi[2 * i + 1] = c[i].imag(); }
  for (int i = 0; i < 27 && chi_low; i++) { chi_low[2 * i] = l[i].real(); chi_low[2 * i + 1] = l[i].imag(); }
  return 0;
}
