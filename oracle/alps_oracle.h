/*
 * TEST INFRASTRUCTURE ONLY -- NOT PART OF THE PRODUCT PATH.
 *
 * CPU restatement ("oracle") of the reference ALPS algorithm for the hot path
 * disp(om): chi_s(omega,k), the dispersion tensor and its determinant.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; alps_b200/ never does.
 *
 * Parity status: the Fortran/MPI reference cannot be compiled in this image
 * (no gfortran / MPI), so this restatement is pinned to the reference only by
 * the reference's own 5-digit golden files (tests/golden/test_kpar_fast.*).
 * At the 1e-9 level: PARITY UNPINNED by the reference itself.
 */
#ifndef ALPS_ORACLE_H
#define ALPS_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* All arrays use the reference's Fortran layout (column-major, species index
 * fastest), see src/ALPS_var.f90:174-246 and src/ALPS_com.f90:183-202. */

typedef struct {
  int nspec, nperp, npar;
  int ngamma, npparbar;       /* relativistic grid (unused unless relativistic) */
  double vA;
  double Bessel_zero;
  double Tlim;
  int positions_principal;
  int n_resonance_interval;
  int kperp_norm;
  int nproc;                  /* emulated MPI size (0 = one worker per species, n in [0,nmax]) */
  int maxfits;                /* maxval(n_fits) */
  int maxorder;               /* max poly_order */
  int nmax_force;             /* >0: nmax(is) = nmax_force for table species (synthetic configs) */
} oracle_cfg;

int  oracle_init(const oracle_cfg *cfg);
void oracle_finalize(void);

/* per species (is is 1-based like the reference) */
int oracle_set_species(int is, double ns, double qs, double ms, int relativistic,
                       int usebM, int ACmethod, int n_fits, const int *fit_type,
                       const double *perp_correction, int logfit, int poly_kind,
                       int poly_order, double poly_log_max);

/* pp(nspec,0:nperp,0:npar,2), df0(nspec,1:nperp-1,1:npar-1,2),
 * param_fit(nspec,0:max(nperp,ngamma),5,maxfits), poly_fit_coeffs(nspec,0:nperp,0:maxorder) */
int oracle_upload(const double *pp, const double *df0, const double *param_fit,
                  const double *poly_fit_coeffs);

/* relativistic tables of derivative_f0_rel (src/ALPS_fns_rel.f90:36-276):
 * f0_rel, gamma_rel, pparbar_rel (nspec_rel,0:ngamma,0:npparbar), df0_rel (...,2) */
int oracle_upload_rel(int nspec_rel, const double *f0_rel, const double *df0_rel, const double *gamma_rel,
                      const double *pparbar_rel);
double oracle_int_ee_rel(int is);

/* derivative_f0: f0(nspec,0:nperp,0:npar) -> df0 (src/ALPS_fns.f90:96-118) */
int oracle_derivative_f0(const double *f0, const double *pp, double *df0_out,
                         int nspec, int nperp, int npar);

/* determine_nmax + split_processes + determine_bessel_array for every worker */
int oracle_set_k(double kperp, double kpar, int *nmax_out);

/* D = disp(om); outputs chi0(nspec,3,3), chi0_low(nspec,3,3,-1:1), wave(3,3) as
 * interleaved (re,im) doubles in Fortran element order. Any pointer may be NULL. */
int oracle_disp(const double om[2], double D[2], double *chi0, double *chi0_low,
                double *wave);

/* Restrict every worker to harmonics |n| <= ncap (bench sampling only; <0 = off). */
/* chi(3,3), chi_low(3,3,-1:1) of a use_bM species (NHDS calc_chi output) for the following oracle_disp calls */
void oracle_set_external_chi(int is, const double *chi, const double *chi_low);
void oracle_set_ncap(int ncap);
/* Evaluate only the harmonics with |n| % stride == offset (bench sampling only: a strided sample holds the expensive
 * resonant harmonics in proportion, unlike the |n| <= ncap cut; stride <= 1 = off).  D is then NOT the dispersion
 * determinant. */
void oracle_set_sample(int stride, int offset);
void oracle_set_threads(int nthreads);

/* exposed pieces for unit tests */
double oracle_bessj(int n, double x);
void   oracle_eval_fit(int is, int iperp, const double p[2], double out[2]);
void   oracle_full_integrate(int is, int nn, int mode, const double om[2], double out[2],
                             int *found_res);
double oracle_int_ee(int is);
void   oracle_get_nlim(int *nworkers, int *sproc, int *nlim1, int *nlim2, int cap);

#ifdef __cplusplus
}
#endif
#endif
