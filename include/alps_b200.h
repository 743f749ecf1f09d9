/*
 * alps_b200 -- C ABI of the B200-native replacement for the ALPS hot path
 *
 *     double complex function disp(om)            (reference: src/ALPS_fns.f90:252-636)
 *
 * i.e. chi_s(omega,k), the dispersion tensor and its determinant D(omega,k) for tabulated
 * gyrotropic f0(p_perp,p_par), plus the k-dependent set-up that path needs
 * (determine_nmax / split_processes / determine_bessel_array, src/ALPS_fns.f90:3971-4255) and
 * the f0 derivatives (derivative_f0, src/ALPS_fns.f90:34-248).
 *
 * The reference has no plugin/FFI interface: the boundary is that Fortran module function and
 * the alps_var module globals it reads (src/ALPS_var.f90:174-246).  A ~60 line iso_c_binding
 * shim keeps disp()'s signature and forwards to the entry points below.
 * All arrays are handed over exactly as the Fortran side holds them: column-major, species
 * index fastest, interleaved (re,im) doubles for complex values.  The library copies what it is
 * given; the caller keeps ownership.  One host thread, not re-entrant, one CUDA stream.
 *
 * Return value: 0 on success, a positive ALPS error id as used by alps_error()
 * (src/ALPS_io.f90:937-1012) or a negative value for CUDA / usage errors
 * (see alps_b200_last_error()).  There is no CPU fallback: every call fails if no sm_100 device
 * is usable.
 */
#ifndef ALPS_B200_H
#define ALPS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ALPS_B200_ERR_CUDA (-1)      /* CUDA runtime / driver failure            */
#define ALPS_B200_ERR_USAGE (-2)     /* bad argument or call order                */
#define ALPS_B200_ERR_UNSUPPORTED (-3) /* feature of the reference not built yet  */
#define ALPS_B200_ERR_GRID (-4)      /* (p_perp,p_par) grid not separable/uniform */
#define ALPS_B200_ERR_NMAX (-5)      /* determine_nmax exceeded nmax_cap          */

/* Scalars of namelist &system (src/ALPS_io.f90:79-85) that the path reads. */
typedef struct {
  int nspec, nperp, npar;
  int ngamma, npparbar;          /* relativistic grid (src/ALPS_var.f90:60-64)               */
  double vA;
  double Bessel_zero;            /* determine_nmax threshold, src/ALPS_fns.f90:4017           */
  double Tlim;                   /* analytic switch, src/ALPS_fns.f90:1026                    */
  int positions_principal;       /* M_I, src/ALPS_fns.f90:981-1006                            */
  int n_resonance_interval;      /* M_P, src/ALPS_fns.f90:1019                                */
  int kperp_norm;                /* src/ALPS_fns.f90:322-328                                  */
  int emulate_nproc;             /* MPI size to emulate in determine_nmax/split_processes
                                    (nmax and the summed harmonic range depend on it,
                                    src/ALPS_fns.f90:4048-4064, 4185-4200); 0 = n in [0,nmax]  */
  int maxfits;                   /* maxval(n_fits): extent of param_fit, src/ALPS_io.f90:181  */
  int maxorder;                  /* max poly_order: extent of poly_fit_coeffs                 */
  int device;                    /* CUDA device ordinal; -1 = current device                  */
  int nmax_cap;                  /* upper bound for nmax (0 = 2000)                           */
  int batch_max;                 /* omegas processed per internal chunk (0 = auto)            */
  int nmax_force;                /* >0: skip determine_nmax and sum n in [0,nmax_force] for every
                                    table species (synthetic benchmark configs only)          */
  int ngpu;                      /* devices this PROCESS drives: device, device+1, ... (0 or 1 = one GPU).  With
                                    ngpu > 1 the library keeps the tables on every device and partitions the work
                                    itself (alps_b200_set_partition) -- the replacement for `mpirun -np N` when only
                                    one process calls (the Fortran shim)                  */
} alps_b200_cfg;

/* replaces: allocation + pass_instructions (src/ALPS_com.f90:28-170) for the path's scalars */
int alps_b200_init(const alps_b200_cfg *cfg);
void alps_b200_finalize(void);
const char *alps_b200_last_error(void);

/* replaces: the per-species namelists &spec_j/&ffit_j_i/&poly_spec_j as broadcast by
 * pass_instructions (src/ALPS_com.f90:60-170).  is is 1-based. */
int alps_b200_set_species(int is, double ns, double qs, double ms, int relativistic, int usebM,
                          int ACmethod, int n_fits, const int *fit_type,
                          const double *perp_correction, int logfit, int poly_kind,
                          int poly_order, double poly_log_max);

/* replaces: pass_distribution (src/ALPS_com.f90:175-284).  Fortran arrays as they are:
 *   pp(nspec,0:nperp,0:npar,2), df0(nspec,1:nperp-1,1:npar-1,2),
 *   param_fit(nspec,0:max(nperp,ngamma),5,maxfits), poly_fit_coeffs(nspec,0:nperp,0:maxorder).
 * df0 may be NULL if alps_b200_derivative_f0 is called instead.  param_fit / poly may be NULL
 * when no species needs them. */
int alps_b200_upload(const double *pp, const double *df0, const double *param_fit,
                     const double *poly_fit_coeffs);

/* replaces: the relativistic part of pass_distribution (src/ALPS_com.f90:252-278): the tables of
 * derivative_f0_rel, f0_rel / gamma_rel / pparbar_rel (nspec_rel,0:ngamma,0:npparbar) and
 * df0_rel(nspec_rel,0:ngamma,0:npparbar,2), Fortran layout.  Call after alps_b200_set_species. */
int alps_b200_upload_rel(int nspec_rel, const double *f0_rel, const double *df0_rel,
                         const double *gamma_rel, const double *pparbar_rel);

/* replaces: derivative_f0 (src/ALPS_fns.f90:96-118) -- centred differences on the device from
 * f0(nspec,0:nperp,0:npar); optional df0_out (host, Fortran layout of df0) receives the result. */
int alps_b200_derivative_f0(const double *f0, double *df0_out);

/* replaces: determine_nmax + split_processes + determine_bessel_array
 * (src/ALPS_fns.f90:3971-4255), called whenever k changes (src/ALPS_fns.f90:2465-2472). */
int alps_b200_set_k(double kperp, double kpar, int *nmax_out /* nspec, may be NULL */);

/* replaces: disp(om) for one omega.  Outputs (any may be NULL): D[2]; chi0(nspec,3,3);
 * chi0_low(nspec,3,3,-1:1); wave(3,3) -- the rank-0 globals calc_eigen consumes
 * (src/ALPS_fns.f90:2611, 2687-2702). */
int alps_b200_disp(const double om[2], double D[2], double *chi0, double *chi0_low, double *wave);

/* replaces: the serial nr x ni loop of map_search (src/ALPS_fns.f90:3697-3757) and any batch of
 * independent root iterations: n omegas in, n D's out (host buffers; H2D/D2H inside).
 * chi0_opt (may be NULL): n x chi0(nspec,3,3). */
int alps_b200_disp_batch(int n, const double *om, double *D, double *chi0_opt);
/* the same with every side output of disp() per omega (any may be NULL): n x chi0(nspec,3,3), n x chi0_low(nspec,3,3,-1:1),
 * n x wave(3,3) -- what calc_eigen would read after each of the n disp calls (src/ALPS_fns.f90:2687-2702) */
int alps_b200_disp_batch_full(int n, const double *om, double *D, double *chi0_opt, double *chi0_low_opt,
                              double *wave_opt);

/* Solver aid: evaluate up to 8 omegas that are about to be requested one by one (the start pair of secant / rtsec, the
 * om, om(1 +- delta) triple of secant_osc's Newton step, src/ALPS_fns.f90:2046) as one small batch and keep the results in
 * the memo of alps_b200_disp; the single calls that follow return the bitwise same values without a launch. */
int alps_b200_disp_prefetch(int n, const double *om);

/* Same with device-resident buffers (om, D: 2n doubles in HBM) on the library's stream. */
int alps_b200_disp_batch_dev(int n, const double *d_om, double *d_D);

/* replaces: calc_chi (NHDS) results being summed into chi for use_bM species
 * (src/ALPS_fns.f90:344-362): chi(3,3) and chi_low(3,3,-1:1) for the next alps_b200_disp call. */
int alps_b200_add_external_chi(int is, const double *chi, const double *chi_low);

/* replaces: calc_chi of module alps_nhds (src/ALPS_NHDS.f90:59-464) for use_bM species: with the
 * &bM_spec_j parameters set, every disp call computes the closed-form bi-Maxwellian / cold chi on the
 * device (k_nhds: one warp per (omega, species), O(nmax) terms; the BESSI(n,z) table and the harmonic
 * cut-off are rebuilt by alps_b200_set_k) and sums it in like src/ALPS_fns.f90:344-362. */
int alps_b200_set_bm_species(int is, int bM_nmaxs, double bM_Bessel_zeros, double bM_betas,
                             double bM_alphas, double bM_pdrifts);
/* the same device kernels for one species and one omega, stateless (needs a GPU like every entry point):
 * chi(3,3), chi_low(3,3,-1:1), column-major complex */
int alps_b200_nhds_calc_chi(double ns, double qs, double ms, int bM_nmaxs, double bM_Bessel_zeros,
                            double bM_betas, double bM_alphas, double bM_pdrifts, double kz, double kperp,
                            const double x[2], int kperp_norm, double *chi, double *chi_low);

/* replaces: the evaluation loop of polyharmonic_spline inside derivative_f0_rel (src/ALPS_fns_rel.f90:300-331,
 * 407-423): out[j] = sum_i w[i] phi(|(gx[j],px[j]) - (gc[i],pc[i])|) + w[n] + w[n+1] gx[j] + w[n+2] px[j]
 * with the reference's thin-plate kernel phi, over the n table nodes in order.  Host buffers; stateless; needs a GPU.
 * The dense (n+3)^2 solve for w stays with LAPACK on the host like in the reference (line 402). */
int alps_b200_tps_eval(int n, const double *gc, const double *pc, const double *w, int npts, const double *gx,
                       const double *px, double *out);

/* ------------------------------------------------------------------------------------------
 * Partition over the GPUs of one box -- replaces ALPS_com's MPI harmonic split (split_processes + the two MPI_REDUCEs
 * of disp(), src/ALPS_fns.f90:4079-4207, 519-523).  Two ways to have several GPUs:
 *   (a) device group: alps_b200_cfg.ngpu = N, ONE process; nothing else changes in the caller;
 *   (b) one process per GPU (mpirun / torchrun, like the reference's ranks): every rank initialises its own device and
 *       joins a library-owned NCCL communicator (alps_b200_comm_unique_id on rank 0, the 128 bytes broadcast by the
 *       caller -- MPI_Bcast / torch.distributed -- then alps_b200_comm_init on every rank).  disp / disp_batch /
 *       map_search are then COLLECTIVE like the reference's disp(): every rank calls them with the same arguments and
 *       every rank gets the full result.
 * and two partitions (alps_b200_set_partition; call alps_b200_set_k afterwards):
 *   OMEGA     (default) batches of more than 8 omegas -- the nr x ni loop of map_search, batches of roots -- are cut
 *             into contiguous slices, one per GPU, tables replicated, no communication but the final gather of D
 *             ((a): the slices land in the caller's host array; (b): one ncclAllGather).  Single omegas run on one GPU.
 *   HARMONIC  every GPU sums a contiguous block of |n| per species for ALL omegas of the call; the un-normalised chi
 *             partials (48 complex per species and omega) are summed over NVLink -- (a) by one kernel on device 0 that
 *             reads its peers' buffers (peer memory; ALPS_B200_REDUCE=nccl: ncclAllReduce), (b) ncclAllReduce on the
 *             library's stream -- and D is assembled from the sum.  Pays when one D is much more than ~100 us of GPU
 *             work (C5-sized tables); for the shipped configs the OMEGA partition is the faster one.
 *             Calls of at most 8 omegas on a configuration with fewer than 2e8 point-harmonics per D are therefore
 *             evaluated UNSHARDED (device 0 of a group / redundantly on every rank, no exchange): a single disp() under
 *             this partition is never slower than on one GPU, and bitwise the one-GPU value.
 * The result of a call does not depend on how it was cut: every piece sums in the order of the whole call's batch
 * class, so ngpu = N and ngpu = 1 (and N ranks vs one) give bitwise identical D under the OMEGA partition. */
#define ALPS_B200_PARTITION_OMEGA 0
#define ALPS_B200_PARTITION_HARMONIC 1
int alps_b200_set_partition(int kind);
int alps_b200_comm_unique_id(char id[128]);                        /* ncclGetUniqueId (rank 0)               */
int alps_b200_comm_init(int rank, int nranks, const char id[128]); /* ncclCommInitRank on this rank's device */
int alps_b200_comm_finalize(void);
/* host-only: the slice [lo, hi) of n omegas that part `rank` of `nparts` evaluates under the OMEGA partition */
int alps_b200_omega_slice(int n, int rank, int nparts, int *lo, int *hi);

/* Low-level harmonic sharding for callers that do their own reduction (kept from round 1):
 * restrict this process to harmonics |n| in [nlo,nhi] of species is (is=0: all species),
 * produce un-normalised partial sums (caller all-reduces them over NCCL), then assemble. */
int alps_b200_set_harmonic_shard(int rank, int nranks);
int alps_b200_chi_partial_len(void);                       /* doubles per omega            */
int alps_b200_chi_partial_dev(int n, const double *d_om, double *d_partial);
int alps_b200_assemble_dev(int n, const double *d_om, const double *d_partial, double *d_D);

/* 0 = direct quadrature per omega (default); 1 = k-hoisted tables (p_perp sums precomputed in
 * set_k, O(nmax*npar) per omega). */
int alps_b200_set_mode(int mode);

/* plumbing */
int alps_b200_set_stream(void *cuda_stream);   /* cudaStream_t to launch on (NULL = legacy default
                                                  stream); until called, a library-owned stream */
int alps_b200_sync(void);
int alps_b200_get_info(int what, double *out); /* see ALPS_B200_INFO_* */
#define ALPS_B200_INFO_POINT_HARMONICS 0  /* sum_s (2 nmax_s+1)(nperp-1)(npar-1) for current k */
#define ALPS_B200_INFO_LAUNCHES 1         /* kernels launched since init                        */
#define ALPS_B200_INFO_SM_COUNT 2
#define ALPS_B200_INFO_LAST_KERNEL_MS 3   /* device time of the last quadrature kernel batch    */
#define ALPS_B200_INFO_BATCH 4            /* internal omega chunk size                          */
#define ALPS_B200_INFO_DFMA_NOREUSE 5     /* DFMA micro-benchmark with three fresh operands per FMA,
                                             TFLOP/s (register-read limit of tiled FP64 kernels)  */
#define ALPS_B200_INFO_DMMA_PEAK 6        /* FP64 tensor-pipe micro-benchmark (mma.sync.m8n8k4.f64), TFLOP/s */
#define ALPS_B200_INFO_QUAD_VARIANT 7     /* id of the quadrature kernel variant in use (>= 9: DMMA) */
#define ALPS_B200_INFO_D_EVALS 8          /* D(omega,k) evaluations since init (every entry point)             */
#define ALPS_B200_INFO_SET_K_CALLS 9      /* alps_b200_set_k calls since init                                   */
#define ALPS_B200_INFO_PREFETCHED 11      /* omegas evaluated ahead of their request by alps_b200_disp_prefetch (part of D_EVALS) */
#define ALPS_B200_INFO_NGPU 12            /* devices of this process' group x ranks of the communicator       */
#define ALPS_B200_INFO_MEMO_HITS 10       /* alps_b200_disp calls answered from the memo of the last omegas (same
                                             omega bits, same state: no launch; not counted in D_EVALS)          */

/* ------------------------------------------------------------------------------------------
 * Host-side twins of the reference's omega-point generators (alps_b200/csrc/drivers.cpp).  They
 * only produce omegas, call alps_b200_disp / alps_b200_disp_batch, and write the reference's
 * output files; all paths may be NULL (no file output).
 * ------------------------------------------------------------------------------------------ */
typedef struct {      /* &system entries read by the solvers, src/ALPS_io.f90:79-85 */
  int numiter;
  double D_threshold, D_prec, D_tol, D_gap;
  int secant_method;  /* 0 secant, 1 rtsec, 2 secant_osc (src/ALPS_fns.f90:2504-2516) */
} alps_b200_solver_opts;

typedef struct {      /* &maps_1, src/ALPS_io.f90:233-256 */
  double omi, omf, gami, gamf;
  int nr, ni, loggridw, loggridg, determine_minima;
} alps_b200_map;

typedef struct {      /* type scanner, src/ALPS_var.f90:354-386 */
  int type;           /* 0 k1_k2, 1 theta, 2 |k| at constant theta, 3 kperp, 4 kpar */
  int n_out, n_res, log_scan, eigen, heat;
  double diff, diff2;
} alps_b200_scan;

/* replaces: secant (src/ALPS_fns.f90:1815-1917), secant_osc (:1919-2101), rtsec (:2105-2195) */
int alps_b200_secant(double om[2], const alps_b200_solver_opts *o, int *iters);
int alps_b200_secant_osc(double om[2], const alps_b200_solver_opts *o, int *iters);
int alps_b200_rtsec(double om[2], const alps_b200_solver_opts *o, int *iflag);
/* replaces: refine_guess (:3793-3856); writes <runname>.roots (i4,5es14.4e3) */
int alps_b200_refine_guess(int nroots, double *wroots, const alps_b200_solver_opts *o,
                           const char *roots_path, double *D_out);
/* replaces: map_search (:3595-3788) + find_minima (:3860-3966); the nr x ni loop is one GPU batch;
 * writes <runname>.map (5es16.6e3).  om/cal: nr*ni complex (ir fastest), val: nr*ni, iroots(2,numroots) */
int alps_b200_map_search(const alps_b200_map *m, const char *map_path, double *om_out, double *val_out,
                         double *cal_out, int numroots, int *iroots, int *nroots_found);
/* Formulation map_search evaluates its nr x ni batch in: 1 (default) = k-hoisted tables (alps_b200_set_mode(1) for the
 * duration of the batch: one table build per k, then O(nmax npar) per omega; the same D up to rounding, ~1e-12), 0 =
 * whatever alps_b200_set_mode says (direct quadrature unless changed).  The mode in force before the map is restored
 * after it, so the root refinement that follows runs the direct quadrature.  alps_b200_map_eval is the batch call
 * map_search makes (n omegas in, n D out, host buffers). */
int alps_b200_set_map_mode(int mode);
int alps_b200_map_eval(int n, const double *om, double *D);
/* Multi-GPU map_search (omega sharding: replaces the MPI harmonic split for maps).  Host-only
 * halves of alps_b200_map_search so that the nr x ni grid can be evaluated in slices by several processes
 * (one per GPU, alps_b200_disp_batch on each slice, slices gathered by the caller -- NCCL / torch.distributed):
 *   alps_b200_map_grid    the omega grid of :3684-3712, nr*ni complex, ir fastest;
 *   alps_b200_map_finish  everything after the loop of disp calls (:3722-3786) on the gathered D: cal_io
 *                         (nr*ni complex, D in, D with the NaN / infinity sentinels out), val = log10|D|,
 *                         the .map file (map_path may be NULL) and find_minima. */
int alps_b200_map_grid(const alps_b200_map *m, double *om_out);
int alps_b200_map_finish(const alps_b200_map *m, double *cal_io, const char *map_path, double *val_out,
                         int numroots, int *iroots, int *nroots_found);
/* replaces: calc_eigen (:2605-2899).  current_int(nspec) from derivative_f0 (may be NULL = 0).
 * Outputs: ef(3), bf(3), Us(3,nspec), ds(nspec) complex; Ps(nspec), Ps_split(4,nspec), W_EM real. */
int alps_b200_calc_eigen(const double om[2], int nspec, const double *ns, const double *qs,
                         const double *current_int, double kperp, double kpar, double vA, int eigen,
                         int heat, double *ef, double *bf, double *Us, double *ds, double *Ps,
                         double *Ps_split, double *W_EM);
/* Root batching: with on != 0, refine_guess and om_scan advance all roots of a k step concurrently --
 * each root runs the unchanged serial algorithm on a host thread and the D requests of all waiting
 * roots are served by alps_b200_disp_batch launches of at most 8 omegas (the batch class of a single disp() call),
 * so the results per root are bit-identical to the serial order for any number of roots. */
int alps_b200_set_root_batching(int on);

/* replaces: scan_read step sizes (src/ALPS_io.f90:472-549); updates kperp_last / kpar_last */
int alps_b200_scan_setup(int scan_type, double swi, double swf, int swlog, int ns, int nres, int eigen,
                         int heat, double *kperp_last, double *kpar_last, alps_b200_scan *out);
/* replaces: om_scan (:2198-2600); writes <prefix>.scan_<id><ik>.root_<in> (+ .eigen_, .heat_,
 * .heat_mech_).  rows_out (may be NULL): (n_out+1) x nroots x 4 doubles (kperp,kpar,Re om,Im om). */
int alps_b200_om_scan(const alps_b200_scan *sc, int nroots, double *wroots,
                      const alps_b200_solver_opts *o, int nspec, const double *ns, const double *qs,
                      const double *current_int, double vA, double *kperp_io, double *kpar_io,
                      const char *prefix, int ik, double *rows_out);

/* replaces: om_double_scan (:2904-3591, scan_option=2): for every output step of scan 1 a full scan 2;
 * writes <prefix>.scan_<id1>_<id2>.root_<in> (+ .eigen_, .heat_, .heat_mech_).  The outer guesses are kept
 * in single precision like the reference's `complex :: omlast`.  rows_out (may be NULL):
 * (n_out1+1) x (n_out2+1) x nroots x 4 doubles. */
int alps_b200_om_double_scan(const alps_b200_scan *sc1, const alps_b200_scan *sc2, int nroots,
                             double *wroots, const alps_b200_solver_opts *o, int nspec, const double *ns,
                             const double *qs, const double *current_int, double vA, double *kperp_io,
                             double *kpar_io, const char *prefix, double *rows_out);

/* Host-only helper (no GPU needed): determine_nmax's "more processes than harmonics" adjustment
 * and split_processes (src/ALPS_fns.f90:4048-4064, 4079-4207) for an emulated MPI size nproc.
 * nmax[nspec] is updated in place; nhi[nspec] receives the highest harmonic any rank sums. */
int alps_b200_emulate_split(int nproc, int nspec, const int *usebM, int *nmax, int *nhi);

/* FP64 FMA micro-benchmark (roofline denominator; not part of the path): returns TFLOP/s. */
int alps_b200_dfma_peak(double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* ALPS_B200_H */
