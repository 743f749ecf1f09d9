// alps_b200: host-side twins of the reference's omega-point generators.  Every D comes from the
// GPU through the C ABI (alps_b200_disp / alps_b200_disp_batch); nothing here integrates anything.
//
//   secant        src/ALPS_fns.f90:1815-1917      secant_osc   :1919-2101     rtsec :2105-2195
//   refine_guess  :3793-3856 (.roots writer)      map_search   :3595-3788 (.map writer)
//   find_minima   :3860-3966                      calc_eigen   :2605-2899
//   om_scan       :2198-2600 (.scan_* / .eigen_* / .heat_* / .heat_mech_* writers)
// File formats are the reference's edit descriptors (es14.4e3, es16.6e3, i4).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <complex>
#include <condition_variable>
#include <functional>
#include <initializer_list>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/alps_b200.h"

namespace {

typedef std::complex<double> cplx;
const cplx II(0.0, 1.0);

struct DispError {
  int code;
};

// ---------------------------------------------------------------------------- root batching
// The reference refines its roots one after the other; every iteration of every root is one
// latency-bound disp() call.  With batching on, each root's *unchanged serial algorithm* runs on its
// own host thread, and a broker collects the D requests of all roots that are waiting and serves
// them with alps_b200_disp_batch launches of at most 8 omegas (the batch class of a single disp() call).  Per-root
// results are bit-identical to the serial order for any number of roots (disp_batch and disp agree bitwise within
// that class); only the latency is shared.
struct Broker {
  struct Req {
    cplx om, D;
    double *chi0 = nullptr, *chi0_low = nullptr, *wave = nullptr;   // full request if chi0 != nullptr
    bool pending = false;
    int rc = 0;
  };
  std::mutex m;
  std::condition_variable cv_broker, cv_worker;
  std::vector<Req> reqs;
  int active = 0;

  cplx request(int id, cplx om, double* chi0, double* chi0_low, double* wave) {
    std::unique_lock<std::mutex> lk(m);
    Req& r = reqs[id];
    r.om = om; r.chi0 = chi0; r.chi0_low = chi0_low; r.wave = wave; r.pending = true; r.rc = 0;
    cv_broker.notify_one();
    cv_worker.wait(lk, [&] { return !r.pending; });
    if (r.rc) throw DispError{r.rc};
    return r.D;
  }
  void finish() {
    std::lock_guard<std::mutex> lk(m);
    active--;
    cv_broker.notify_one();
  }
  // run the jobs (one per root) concurrently, serving their disp requests in batches
  void run(const std::vector<std::function<void()> >& jobs);
};
thread_local Broker* tl_broker = nullptr;
thread_local int tl_id = 0;
bool g_root_batching = false;

cplx disp1(cplx om) {
  if (tl_broker) return tl_broker->request(tl_id, om, nullptr, nullptr, nullptr);
  double o[2] = {om.real(), om.imag()}, D[2] = {0, 0};
  int rc = alps_b200_disp(o, D, nullptr, nullptr, nullptr);
  if (rc) throw DispError{rc};
  return cplx(D[0], D[1]);
}

// omegas the serial algorithm is about to evaluate one after the other: one small batch into the memo of
// alps_b200_disp (bitwise the same values); the algorithm itself is untouched
void prefetch(std::initializer_list<cplx> oms) {
  if (tl_broker) return;
  double o[16];
  int n = 0;
  for (const cplx& w : oms) {
    if (n == 8) break;
    o[2 * n] = w.real();
    o[2 * n + 1] = w.imag();
    n++;
  }
  int rc = alps_b200_disp_prefetch(n, o);
  if (rc) throw DispError{rc};
}

void Broker::run(const std::vector<std::function<void()> >& jobs) {
  const int n = (int)jobs.size();
  reqs.assign(n, Req());
  active = n;
  std::vector<std::thread> th;
  std::vector<int> err(n, 0);
  for (int i = 0; i < n; i++)
    th.emplace_back([&, i] {
      tl_broker = this;
      tl_id = i;
      try {
        jobs[i]();
      } catch (const DispError& e) {
        err[i] = e.code;
      }
      tl_broker = nullptr;
      finish();
    });
  std::vector<double> om, D;
  std::vector<int> ids;
  for (;;) {
    std::unique_lock<std::mutex> lk(m);
    cv_broker.wait(lk, [&] {
      int pend = 0;
      for (auto& r : reqs) pend += r.pending ? 1 : 0;
      return active == 0 || pend == active;
    });
    if (active == 0) break;
    om.clear();
    ids.clear();
    for (int i = 0; i < n; i++)
      if (reqs[i].pending && !reqs[i].chi0) {
        ids.push_back(i);
        om.push_back(reqs[i].om.real());
        om.push_back(reqs[i].om.imag());
      }
    int rc = 0;
    if (!ids.empty()) {
      D.assign(om.size(), 0.0);
      // chunks of at most 8 omegas (LAT_BATCH in api.cu): every evaluation stays in the batch class of a single
      // disp() call, so a root's D bits -- and its iteration path -- do not depend on how many other roots are pending
      const int CHUNK = 8;
      for (int q0 = 0; q0 < (int)ids.size() && !rc; q0 += CHUNK)
        rc = alps_b200_disp_batch(std::min(CHUNK, (int)ids.size() - q0), om.data() + 2 * q0, D.data() + 2 * q0, nullptr);
      for (size_t q = 0; q < ids.size(); q++) {
        reqs[ids[q]].D = cplx(D[2 * q], D[2 * q + 1]);
        reqs[ids[q]].rc = rc;
      }
    }
    for (int i = 0; i < n; i++)
      if (reqs[i].pending && reqs[i].chi0) {   // calc_eigen needs chi0, chi0_low, wave: served one by one
        double o[2] = {reqs[i].om.real(), reqs[i].om.imag()}, d[2] = {0, 0};
        reqs[i].rc = alps_b200_disp(o, d, reqs[i].chi0, reqs[i].chi0_low, reqs[i].wave);
        reqs[i].D = cplx(d[0], d[1]);
      }
    for (auto& r : reqs) r.pending = false;
    cv_worker.notify_all();
  }
  for (auto& t : th) t.join();
  for (int i = 0; i < n; i++)
    if (err[i]) throw DispError{err[i]};
}

// run one job per root, concurrently through the broker when batching is on
void for_each_root(const std::vector<std::function<void()> >& jobs) {
  if (g_root_batching && jobs.size() > 1 && !tl_broker) {
    Broker b;
    b.run(jobs);
  } else {
    for (auto& j : jobs) j();
  }
}

// Fortran ESw.dEe edit descriptor (e.g. es14.4e3 -> "   1.0000E-002")
std::string es(double v, int w, int d, int e) {
  char buf[128];
  if (std::isnan(v)) {
    snprintf(buf, sizeof(buf), "%*s", w, "NaN");
    return buf;
  }
  if (std::isinf(v)) {
    snprintf(buf, sizeof(buf), "%*s", w, v > 0 ? "Infinity" : "-Infinity");
    return buf;
  }
  char m[64];
  snprintf(m, sizeof(m), "%.*E", d, v);
  char* ep = strchr(m, 'E');
  int ex = atoi(ep + 1);
  *ep = 0;
  char body[96];
  snprintf(body, sizeof(body), "%sE%c%0*d", m, ex < 0 ? '-' : '+', e, ex < 0 ? -ex : ex);
  snprintf(buf, sizeof(buf), "%*s", w, body);
  return buf;
}
std::string es14(double v) { return es(v, 14, 4, 3); }
std::string es16(double v) { return es(v, 16, 6, 3); }

const double F01 = (double)0.1f;       // REAL*4 literals of secant_osc promoted to double
const double F1EM3 = (double)1.0e-3f;

// ----------------------------------------------------------------------------- root finders
int secant_impl(cplx& om, const alps_b200_solver_opts& o) {
  cplx prevom = om * (1.0 - o.D_prec);
  prefetch({prevom, om});
  cplx Dprev = disp1(prevom);
  cplx minD = disp1(om), minom = om, D, jump;
  if (std::abs(Dprev) < std::abs(minD)) {
    minom = prevom;
    minD = Dprev;
  }
  int iter = 0;
  bool go = true;
  while (iter <= o.numiter - 1 && go) {
    iter++;
    D = disp1(om);
    if (std::abs(D - Dprev) < 1.e-80) {
      prevom = prevom + 1.e-8;
      Dprev = disp1(prevom);
    }
    if (std::abs(D) < o.D_threshold) {
      jump = 0.0;
      go = false;
    } else {
      jump = D * (om - prevom) / (D - Dprev);
    }
    prevom = om;
    om = om - jump;
    Dprev = D;
    if (std::abs(D) < std::abs(minD)) {
      minom = om;
      minD = D;
    }
  }
  if (iter >= o.numiter) om = minom;
  return iter;
}

int secant_osc_impl(cplx& om, const alps_b200_solver_opts& o) {
  const cplx delta(1.e-6, 1.e-8);
  const double lambda = F01, osc_threshold = F1EM3;
  prefetch({om, om * (1.0 - o.D_prec)});
  cplx D = disp1(om), minom = om, minD = D;
  cplx prevom = om * (1.0 - o.D_prec), prev2om = om, prev3om = om, prev4om = om;
  cplx prevD = disp1(prevom), jump;
  if (std::abs(prevD) < std::abs(minD)) {
    minom = prevom;
    minD = prevD;
  }
  int iter = 0, oscillation_count = 0;
  bool go = true;
  double damping_factor = 1.0;
  auto close = [&](const cplx& a) {
    return (std::fabs(om.real() - a.real()) < std::fabs(om.real()) * osc_threshold) &&
           (std::fabs(om.imag() - a.imag()) < std::fabs(om.imag()) * osc_threshold);
  };
  while (iter <= o.numiter - 1 && go) {
    iter++;
    // the oscillation test below depends on the omegas only: when this iteration will take the finite-difference
    // Newton step, its three evaluations are independent and go out together
    if (oscillation_count + ((iter > 4 && (close(prevom) || close(prev2om) || close(prev3om) || close(prev4om))) ? 1 : 0) > 1)
      prefetch({om, om * (1.0 + delta), om * (1.0 - delta)});
    D = disp1(om);
    if (std::abs(D - prevD) < 1.e-80) {
      prevom = prevom + 1.e-8;
      prevD = disp1(prevom);
    }
    if (std::abs(D) < o.D_threshold) {
      go = false;
    } else {
      if (iter > 4) {
        if (close(prevom) || close(prev2om) || close(prev3om) || close(prev4om)) {
          oscillation_count++;
          damping_factor = std::min(0.5, damping_factor * 0.75);
        }
      }
      if (oscillation_count > 1) {
        cplx Dprime = (disp1(om * (1.0 + delta)) - disp1(om * (1.0 - delta))) / (2.0 * om * delta);
        jump = D / (Dprime + lambda * D);
      } else {
        jump = damping_factor * D * (om - prevom) / (D - prevD);
      }
      if (std::abs(jump) > F01 * std::abs(om)) jump = (F01 * std::abs(om)) * (jump / std::abs(jump));
      if (std::abs(D) > std::abs(prevD)) jump = 0.5 * jump;
      prev4om = prev3om;
      prev3om = prev2om;
      prev2om = prevom;
      prevom = om;
      prevD = D;
      if (std::abs(D) < std::abs(minD)) {
        minom = om;
        minD = D;
      }
      om = om - jump;
    }
  }
  if (iter >= o.numiter) om = minom;
  return iter;
}

cplx rtsec_impl(cplx xin, const alps_b200_solver_opts& o, int* iflag) {
  cplx x1 = xin * 1.0, x2 = xin * (1.0 + o.D_prec);
  prefetch({x1, x2});
  cplx fl = disp1(x1), f = disp1(x2), xl, r, dx;
  if (std::abs(fl) < std::abs(f)) {
    r = x1;
    xl = x2;
    std::swap(fl, f);
  } else {
    xl = x1;
    r = x2;
  }
  for (int j = 1; j <= o.numiter - 1; j++) {
    *iflag = j;
    if (std::abs(f - fl) > 1.e-40)
      dx = (xl - r) * f / (f - fl);
    else
      dx = (x2 - x1) / 25.0;
    xl = r;
    fl = f;
    r = r + dx / 2.0;
    f = disp1(r);
    if (std::abs(dx) < o.D_tol || std::abs(f) == 0.0) return r;
  }
  return r;
}

cplx solve_root(cplx om, const alps_b200_solver_opts& o) {
  int iflag = 0;
  switch (o.secant_method) {
    case 0: secant_impl(om, o); break;
    case 1: om = rtsec_impl(om, o, &iflag); break;
    default: secant_osc_impl(om, o); break;
  }
  return om;
}

// ------------------------------------------------------------------------------- calc_eigen
struct Eigen {
  cplx ef[3], bf[3];
  std::vector<cplx> Us, ds;           // Us(3,nspec) column-major, ds(nspec)
  std::vector<double> Ps, Ps_split;   // Ps(nspec), Ps_split(4,nspec) column-major
  double W_EM = 0.0;
};

struct Tensors {
  int nspec;
  std::vector<cplx> chi0, chi0_low;   // Fortran order (is,i,j) / (is,i,j,m)
  cplx wave[9];
  cplx& c0(int is, int i, int j) { return chi0[is + (size_t)nspec * (i + 3 * j)]; }
  cplx& cl(int is, int i, int j, int m) { return chi0_low[is + (size_t)nspec * (i + 3 * (j + 3 * (m + 1)))]; }
  cplx& w(int i, int j) { return wave[i + 3 * j]; }
};

void disp_full(cplx om, Tensors& t) {
  if (tl_broker) {
    tl_broker->request(tl_id, om, reinterpret_cast<double*>(t.chi0.data()),
                       reinterpret_cast<double*>(t.chi0_low.data()), reinterpret_cast<double*>(t.wave));
    return;
  }
  double o[2] = {om.real(), om.imag()}, D[2];
  int rc = alps_b200_disp(o, D, reinterpret_cast<double*>(t.chi0.data()),
                          reinterpret_cast<double*>(t.chi0_low.data()), reinterpret_cast<double*>(t.wave));
  if (rc) throw DispError{rc};
}

// calc_eigen, src/ALPS_fns.f90:2605-2899.  The caller has just evaluated disp(omega) in the
// reference; here that call is made explicitly.
void calc_eigen_impl(cplx omega, int nspec, const double* ns, const double* qs, const double* current_int,
                     double kperp, double kpar, double vA, bool eigen_L, bool heat_L, Eigen& E) {
  Tensors t;
  t.nspec = nspec;
  t.chi0.assign((size_t)nspec * 9, 0.0);
  t.chi0_low.assign((size_t)nspec * 27, 0.0);
  disp_full(omega, t);
  cplx* e = E.ef;
  e[0] = cplx(1.0, 0.0);
  e[2] = -e[0] * (t.w(1, 0) * t.w(2, 1) - t.w(2, 0) * t.w(1, 1));
  e[2] = e[2] / (t.w(1, 2) * t.w(2, 1) - t.w(2, 2) * t.w(1, 1));
  if (std::abs(t.w(2, 1)) != 0.0) {
    e[1] = -e[2] * t.w(2, 2) - e[0] * t.w(2, 0);
    e[1] = e[1] / t.w(2, 1);
  } else {
    e[1] = t.w(1, 0) * t.w(0, 2) - t.w(0, 0) * t.w(1, 2);
    e[1] = e[1] / (t.w(1, 2) * t.w(0, 1) - t.w(1, 1) * t.w(0, 2));
  }
  cplx* b = E.bf;
  b[0] = -1.0 * kpar * e[1] / (omega * vA);
  b[1] = -1.0 * (kperp * e[2] - kpar * e[0]) / (omega * vA);
  b[2] = kperp * e[1] / (omega * vA);
  E.Us.assign((size_t)3 * nspec, 0.0);
  E.ds.assign(nspec, 0.0);
  E.Ps.assign(nspec, 0.0);
  E.Ps_split.assign((size_t)4 * nspec, 0.0);
  std::vector<double> pflow(nspec, 0.0);
  for (int jj = 0; jj < nspec; jj++) pflow[jj] = current_int ? current_int[jj] / (ns[jj] * qs[jj]) : 0.0;
  if (eigen_L) {
    auto vm = [&](int j, int jj) -> cplx& { return E.Us[j + 3 * (size_t)jj]; };
    for (int jj = 0; jj < nspec; jj++) {
      auto base = [&](int j) {
        cplx s = 0.0;
        for (int q = 0; q < 3; q++) s += e[q] * t.c0(jj, j, q);
        return -(vA * vA / (qs[jj] * ns[jj])) * II * omega * s;
      };
      if (pflow[jj] == 0.0) {
        for (int j = 0; j < 3; j++) vm(j, jj) = base(j);
      } else {
        for (int j = 0; j < 2; j++) vm(j, jj) = base(j);
        vm(2, jj) = base(2) - pflow[jj] * kperp * vm(0, jj) / (omega - kpar * pflow[jj]);
        vm(2, jj) = vm(2, jj) / (1.0 + (kpar * pflow[jj]) / (omega - kpar * pflow[jj]));
      }
    }
    for (int jj = 0; jj < nspec; jj++)
      E.ds[jj] = (1.0 / vA) * (vm(0, jj) * kperp + vm(2, jj) * kpar) / (omega - kpar * pflow[jj]);
  }
  if (!heat_L) return;
  // heating: chi0 at real omega (the Im(om) == 0 branch of full_integrate)
  disp_full(cplx(omega.real(), 0.0), t);
  std::vector<cplx> chia((size_t)nspec * 9), term((size_t)nspec * 3);
  cplx chihold[9], chih[9], dchih[9], term1[3];
  auto ca = [&](int jj, int i, int j) -> cplx& { return chia[jj + (size_t)nspec * (i + 3 * j)]; };
  for (int ii = 0; ii < 3; ii++)
    for (int j = 0; j < 3; j++) {
      cplx s1 = 0.0, s2 = 0.0;
      for (int jj = 0; jj < nspec; jj++) {
        ca(jj, ii, j) = -0.5 * II * (t.c0(jj, ii, j) - std::conj(t.c0(jj, j, ii)));
        s1 += t.c0(jj, ii, j);
        s2 += std::conj(t.c0(jj, j, ii));
      }
      chihold[ii + 3 * j] = 0.5 * (s1 + s2);
    }
  for (int ii = 0; ii < 3; ii++)
    for (int jj = 0; jj < nspec; jj++) {
      cplx s = 0.0;
      for (int q = 0; q < 3; q++) s += std::conj(e[q]) * ca(jj, q, ii);
      term[jj + (size_t)nspec * ii] = s;
    }
  std::vector<cplx> Psc(nspec);
  for (int jj = 0; jj < nspec; jj++) {
    cplx s = 0.0;
    for (int q = 0; q < 3; q++) s += term[jj + (size_t)nspec * q] * e[q];
    Psc[jj] = s;
  }
  disp_full(cplx((omega * 1.000001).real(), 0.0), t);
  for (int ii = 0; ii < 3; ii++)
    for (int j = 0; j < 3; j++) {
      cplx s1 = 0.0, s2 = 0.0;
      for (int jj = 0; jj < nspec; jj++) {
        s1 += t.c0(jj, ii, j);
        s2 += std::conj(t.c0(jj, j, ii));
      }
      chih[ii + 3 * j] = 0.5 * (s1 + s2);
      dchih[ii + 3 * j] = (1.000001 * chih[ii + 3 * j] - chihold[ii + 3 * j]) / 0.000001;
    }
  cplx ew = 0.0;
  for (int ii = 0; ii < 3; ii++) {
    cplx s = 0.0;
    for (int q = 0; q < 3; q++) s += std::conj(e[q]) * dchih[q + 3 * ii];
    term1[ii] = s;
  }
  for (int q = 0; q < 3; q++) ew += term1[q] * e[q];
  for (int q = 0; q < 3; q++) ew += b[q] * std::conj(b[q]);
  E.W_EM = ew.real();   // ewave is double precision: the assignment keeps the real part
  for (int jj = 0; jj < nspec; jj++) E.Ps[jj] = Psc[jj].real() / E.W_EM;   // Ps is real: real part first
  // LD / TTD (n = 0) and cyclotron (n = +-1) split; chi0_low is from the last disp call (lines 2800-2890)
  auto PS = [&](int k, int jj) -> double& { return E.Ps_split[k + 4 * (size_t)jj]; };
  for (int jj = 0; jj < nspec; jj++) {
    cplx p1 = -0.5 * II * std::conj(e[1]) * e[1] * (t.cl(jj, 1, 1, 0) - std::conj(t.cl(jj, 1, 1, 0)));
    double r1 = p1.real();   // each assignment to the real array keeps the real part
    cplx p1b = -0.5 * II * (e[2] * std::conj(e[1]) * t.cl(jj, 1, 2, 0) - std::conj(e[2]) * e[1] * std::conj(t.cl(jj, 1, 2, 0)));
    r1 = (cplx(r1, 0.0) + p1b).real();
    PS(0, jj) = r1;
    cplx p2 = -0.5 * II * (e[1] * std::conj(e[2]) * t.cl(jj, 2, 1, 0) - std::conj(e[1]) * e[2] * std::conj(t.cl(jj, 2, 1, 0)));
    double r2 = p2.real();
    cplx p2b = -0.5 * II * std::conj(e[2]) * e[2] * (t.cl(jj, 2, 2, 0) - std::conj(t.cl(jj, 2, 2, 0)));
    r2 = (cplx(r2, 0.0) + p2b).real();
    PS(1, jj) = r2;
  }
  cplx exy[3] = {e[0], e[1], cplx(0.0, 0.0)};
  for (int pass = 0; pass < 2; pass++) {
    const int m = pass == 0 ? 1 : -1;
    for (int ii = 0; ii < 3; ii++)
      for (int j = 0; j < 3; j++)
        for (int jj = 0; jj < nspec; jj++) ca(jj, ii, j) = -0.5 * II * (t.cl(jj, ii, j, m) - std::conj(t.cl(jj, j, ii, m)));
    for (int jj = 0; jj < nspec; jj++) {
      cplx tot = 0.0;
      for (int ii = 0; ii < 3; ii++) {
        cplx s = 0.0;
        for (int q = 0; q < 3; q++) s += std::conj(exy[q]) * ca(jj, q, ii);
        tot += s * exy[ii];
      }
      PS(2 + pass, jj) = tot.real();
    }
  }
  for (size_t i = 0; i < E.Ps_split.size(); i++) E.Ps_split[i] /= E.W_EM;
}

// ------------------------------------------------------------------------------ find_minima
// src/ALPS_fns.f90:3860-3966: strict minima w.r.t. the 4 neighbours, scanned ii = ni..1, ir = 1..nr
int find_minima(const std::vector<double>& val, int nr, int ni, int numroots, int* iroots) {
  auto V = [&](int ir, int ii) { return val[(ir - 1) + (size_t)nr * (ii - 1)]; };
  int nroots = 0;
  for (int ii = ni; ii >= 1; ii--)
    for (int ir = 1; ir <= nr; ir++) {
      bool okr;
      if (ir == 1) okr = V(ir, ii) < V(ir + 1, ii);
      else if (ir == nr) okr = V(ir, ii) < V(ir - 1, ii);
      else okr = V(ir, ii) < V(ir - 1, ii) && V(ir, ii) < V(ir + 1, ii);
      if (!okr) continue;
      bool oki;
      if (ii == 1) oki = V(ir, ii) < V(ir, ii + 1);
      else if (ii == ni) oki = V(ir, ii) < V(ir, ii - 1);
      else oki = V(ir, ii) < V(ir, ii - 1) && V(ir, ii) < V(ir, ii + 1);
      if (!oki) continue;
      if (nroots < numroots) {
        iroots[2 * nroots] = ir;
        iroots[2 * nroots + 1] = ii;
      }
      nroots++;
    }
  return nroots;
}

void append_line(const std::string& path, const std::string& line, bool replace = false) {
  FILE* f = fopen(path.c_str(), replace ? "w" : "a");
  if (!f) return;
  fputs(line.c_str(), f);
  fclose(f);
}

template <typename F>
int guarded(F f) {
  try {
    f();
  } catch (const DispError& e) {
    return e.code;
  }
  return 0;
}

}  // namespace

extern "C" {

int alps_b200_secant(double om[2], const alps_b200_solver_opts* o, int* iters) {
  return guarded([&] {
    cplx w(om[0], om[1]);
    int it = secant_impl(w, *o);
    om[0] = w.real();
    om[1] = w.imag();
    if (iters) *iters = it;
  });
}

int alps_b200_secant_osc(double om[2], const alps_b200_solver_opts* o, int* iters) {
  return guarded([&] {
    cplx w(om[0], om[1]);
    int it = secant_osc_impl(w, *o);
    om[0] = w.real();
    om[1] = w.imag();
    if (iters) *iters = it;
  });
}

int alps_b200_rtsec(double om[2], const alps_b200_solver_opts* o, int* iflag) {
  return guarded([&] {
    int fl = 0;
    cplx w = rtsec_impl(cplx(om[0], om[1]), *o, &fl);
    om[0] = w.real();
    om[1] = w.imag();
    if (iflag) *iflag = fl;
  });
}

int alps_b200_refine_guess(int nroots, double* wroots, const alps_b200_solver_opts* o, const char* roots_path,
                           double* D_out) {
  return guarded([&] {
    if (roots_path) append_line(roots_path, "", true);
    std::vector<cplx> sol(nroots), dsol(nroots);
    std::vector<std::function<void()> > jobs;
    for (int iw = 0; iw < nroots; iw++)
      jobs.push_back([&, iw] {
        sol[iw] = solve_root(cplx(wroots[2 * iw], wroots[2 * iw + 1]), *o);
        dsol[iw] = disp1(sol[iw]);
      });
    for_each_root(jobs);
    for (int iw = 0; iw < nroots; iw++) {
      cplx om = sol[iw];
      wroots[2 * iw] = om.real();
      wroots[2 * iw + 1] = om.imag();
      cplx d = dsol[iw];
      if (D_out) {
        D_out[2 * iw] = d.real();
        D_out[2 * iw + 1] = d.imag();
      }
      if (roots_path && std::abs(d) != 0.0) {
        char head[16];
        snprintf(head, sizeof(head), "%4d", iw + 1);
        append_line(roots_path, std::string(head) + es14(om.real()) + es14(om.imag()) + es14(log10(std::abs(d))) +
                                    es14(d.real()) + es14(d.imag()) + "\n");
      }
    }
  });
}

// map grid of map_search (src/ALPS_fns.f90:3684-3712), ir fastest
static void map_grid(const alps_b200_map* m, std::vector<cplx>& om) {
  const int nr = m->nr, ni = m->ni;
  om.resize((size_t)nr * ni);
  double dr = m->omf - m->omi, di = m->gamf - m->gami;
  if (nr > 1) dr = (m->omf - m->omi) / (1.0 * (nr - 1));
  if (ni > 1) di = (m->gamf - m->gami) / (1.0 * (ni - 1));
  for (int ir = 1; ir <= nr; ir++) {
    double wr;
    if (m->loggridw) {
      wr = m->omi;
      if (nr > 1) wr = m->omi * pow(m->omf / m->omi, (1.0 * (ir - 1)) / (1.0 * (nr - 1)));
    } else {
      wr = m->omi + dr * (1.0 * (ir - 1));
    }
    for (int ii = 1; ii <= ni; ii++) {
      double wi;
      if (m->loggridg) {
        wi = m->gami;
        if (ni > 1) wi = m->gami * pow(m->gamf / m->gami, (1.0 * (ii - 1)) / (1.0 * (ni - 1)));
      } else {
        wi = m->gami + di * (1.0 * (ii - 1));
      }
      om[(ir - 1) + (size_t)nr * (ii - 1)] = cplx(wr, wi);
    }
  }
}

// everything map_search does after the nr x ni loop of disp calls (:3722-3786): val = log10|D|, the NaN /
// infinity sentinels, the .map file, find_minima
static void map_finish(const alps_b200_map* m, const std::vector<cplx>& om, std::vector<cplx>& cal,
                       std::vector<double>& val, const char* map_path, int numroots, int* iroots,
                       int* nroots_found) {
  const int nr = m->nr, ni = m->ni;
  const size_t n = (size_t)nr * ni;
  val.resize(n);
  for (size_t i = 0; i < n; i++) {
    const double tmp = cal[i].real();
    val[i] = log10(std::abs(cal[i]));
    // NaN / infinity sentinels exactly as written in the reference (lines 3726-3742)
    if (cal[i].imag() != 0.0) {
      if (!(cplx(tmp, 0.0) != cal[i])) {
        cal[i] = 999999.0;
        val[i] = 999999.0;
      }
    } else if (cplx(tmp, 0.0) != cal[i]) {
      cal[i] = 999999.0;
      val[i] = 999999.0;
    }
    if (std::fabs(tmp) > 1.e100) {
      cal[i] = 899999.0;
      val[i] = 899999.0;
    }
  }
  if (map_path) {
    std::string out;
    for (int ir = 1; ir <= nr; ir++) {
      for (int ii = 1; ii <= ni; ii++) {
        size_t i = (ir - 1) + (size_t)nr * (ii - 1);
        out += es16(om[i].real()) + es16(om[i].imag()) + es16(val[i]) + es16(cal[i].real()) + es16(cal[i].imag()) + "\n";
      }
      out += "\n";
    }
    append_line(map_path, out, true);
  }
  if (iroots && nroots_found) {
    *nroots_found = 0;
    if (m->determine_minima && nr > 1 && ni > 1) *nroots_found = find_minima(val, nr, ni, numroots, iroots);
  }
}

static void map_export(const std::vector<cplx>& om, const std::vector<cplx>& cal, const std::vector<double>& val,
                       double* om_out, double* val_out, double* cal_out) {
  for (size_t i = 0; i < om.size(); i++) {
    if (om_out) {
      om_out[2 * i] = om[i].real();
      om_out[2 * i + 1] = om[i].imag();
    }
    if (val_out) val_out[i] = val[i];
    if (cal_out) {
      cal_out[2 * i] = cal[i].real();
      cal_out[2 * i + 1] = cal[i].imag();
    }
  }
}

int alps_b200_map_search(const alps_b200_map* m, const char* map_path, double* om_out, double* val_out,
                         double* cal_out, int numroots, int* iroots, int* nroots_found) {
  return guarded([&] {
    std::vector<cplx> om, cal;
    std::vector<double> val;
    map_grid(m, om);
    cal.resize(om.size());
    // the nr x ni serial loop of the reference becomes one batch on the GPU
    // (in the map mode of alps_b200_set_map_mode: k-hoisted tables by default; over every GPU of the device group or
    // of the communicator under the OMEGA partition)
    int rc = alps_b200_map_eval((int)om.size(), reinterpret_cast<const double*>(om.data()),
                                reinterpret_cast<double*>(cal.data()));
    if (rc) throw DispError{rc};
    map_finish(m, om, cal, val, map_path, numroots, iroots, nroots_found);
    map_export(om, cal, val, om_out, val_out, cal_out);
  });
}

// Multi-GPU map_search (omega sharding): every rank generates the grid, evaluates its slice of it
// with alps_b200_disp_batch, the slices are gathered by the caller, and rank 0 (or every rank) finishes.
int alps_b200_map_grid(const alps_b200_map* m, double* om_out) {
  if (!m || !om_out || m->nr < 1 || m->ni < 1) return ALPS_B200_ERR_USAGE;
  std::vector<cplx> om;
  map_grid(m, om);
  map_export(om, om, std::vector<double>(), om_out, nullptr, nullptr);
  return 0;
}

int alps_b200_map_finish(const alps_b200_map* m, double* cal_io, const char* map_path, double* val_out,
                         int numroots, int* iroots, int* nroots_found) {
  if (!m || !cal_io || m->nr < 1 || m->ni < 1) return ALPS_B200_ERR_USAGE;
  std::vector<cplx> om, cal((size_t)m->nr * m->ni);
  std::vector<double> val;
  map_grid(m, om);
  for (size_t i = 0; i < cal.size(); i++) cal[i] = cplx(cal_io[2 * i], cal_io[2 * i + 1]);
  map_finish(m, om, cal, val, map_path, numroots, iroots, nroots_found);
  map_export(om, cal, val, nullptr, val_out, cal_io);
  return 0;
}

int alps_b200_calc_eigen(const double om[2], int nspec, const double* ns, const double* qs,
                         const double* current_int, double kperp, double kpar, double vA, int eigen, int heat,
                         double* ef, double* bf, double* Us, double* ds, double* Ps, double* Ps_split,
                         double* W_EM) {
  return guarded([&] {
    Eigen E;
    calc_eigen_impl(cplx(om[0], om[1]), nspec, ns, qs, current_int, kperp, kpar, vA, eigen != 0, heat != 0, E);
    for (int q = 0; q < 3; q++) {
      if (ef) { ef[2 * q] = E.ef[q].real(); ef[2 * q + 1] = E.ef[q].imag(); }
      if (bf) { bf[2 * q] = E.bf[q].real(); bf[2 * q + 1] = E.bf[q].imag(); }
    }
    for (size_t i = 0; i < E.Us.size() && Us; i++) { Us[2 * i] = E.Us[i].real(); Us[2 * i + 1] = E.Us[i].imag(); }
    for (size_t i = 0; i < E.ds.size() && ds; i++) { ds[2 * i] = E.ds[i].real(); ds[2 * i + 1] = E.ds[i].imag(); }
    for (size_t i = 0; i < E.Ps.size() && Ps; i++) Ps[i] = E.Ps[i];
    for (size_t i = 0; i < E.Ps_split.size() && Ps_split; i++) Ps_split[i] = E.Ps_split[i];
    if (W_EM) *W_EM = E.W_EM;
  });
}

int alps_b200_set_root_batching(int on) {
  g_root_batching = on != 0;
  return 0;
}

int alps_b200_scan_setup(int scan_type, double swi, double swf, int swlog, int ns, int nres, int eigen, int heat,
                         double* kperp_last, double* kpar_last, alps_b200_scan* out) {
  if (!out || !kperp_last || !kpar_last || scan_type < 0 || scan_type > 4) return ALPS_B200_ERR_USAGE;
  const double pi = 4.0 * atan(1.0), den = 1.0 * ns * nres;
  const double kpl = *kperp_last, kql = *kpar_last;
  out->type = scan_type; out->n_out = ns; out->n_res = nres; out->log_scan = swlog; out->eigen = eigen;
  out->heat = heat; out->diff = 0.0; out->diff2 = 0.0;
  switch (scan_type) {
    case 0:
      if (swlog) {
        out->diff = (log10(swi) - log10(kpl)) / den;
        out->diff2 = (log10(swf) - log10(kql)) / den;
      } else {
        out->diff = (swi - kpl) / den;
        out->diff2 = (swf - kql) / den;
      }
      *kperp_last = swi;
      *kpar_last = swf;
      break;
    case 1: {
      const double theta_0 = atan(kpl / kql), k_0 = sqrt(kpl * kpl + kql * kql);
      // the reference divides by the REAL*4 literal 180. in the log branch (line 506)
      if (swlog) out->diff = (log10(swf * pi / 180.0) - log10(theta_0)) / den;
      else out->diff = ((swf * pi / 180.0) - theta_0) / den;
      *kpar_last = k_0 * cos(swf * pi / 180.0);
      *kperp_last = k_0 * sin(swf * pi / 180.0);
      break;
    }
    case 2: {
      const double theta_0 = atan(kpl / kql), k_0 = sqrt(kpl * kpl + kql * kql);
      out->diff = swlog ? (log10(swf) - log10(k_0)) / den : (swf - k_0) / den;
      *kpar_last = k_0 * cos(theta_0);
      *kperp_last = k_0 * sin(theta_0);
      break;
    }
    case 3:
      out->diff = swlog ? (log10(swf) - log10(kpl)) / den : (swf - kpl) / den;
      *kperp_last = swf;
      break;
    default:
      out->diff = swlog ? (log10(swf) - log10(kql)) / den : (swf - kql) / den;
      *kpar_last = swf;
      break;
  }
  return 0;
}

int alps_b200_om_scan(const alps_b200_scan* sc, int nroots, double* wroots, const alps_b200_solver_opts* o,
                      int nspec, const double* ns, const double* qs, const double* current_int, double vA,
                      double* kperp_io, double* kpar_io, const char* prefix, int ik, double* rows_out) {
  return guarded([&] {
    static const char* ids[5] = {"k1_k2_", "theta_", "kcstq_", "kperp_", "kpara_"};
    double kperp = *kperp_io, kpar = *kpar_io;
    const int nt = sc->n_out * sc->n_res;
    std::vector<bool> jump(nroots, true);
    auto fname = [&](const char* kind, int in) {
      char b[1024];
      snprintf(b, sizeof(b), "%s.%s_%s%d.root_%d", prefix, kind, ids[sc->type], ik, in + 1);
      return std::string(b);
    };
    auto head = [&](int in) {
      return es14(kperp) + es14(kpar) + es14(wroots[2 * in]) + es14(wroots[2 * in + 1]);
    };
    auto eigen_line = [&](int in, const Eigen& E) {
      std::string s = head(in);
      for (int q = 0; q < 3; q++) s += es14(E.ef[q].real()) + es14(E.ef[q].imag());
      for (int q = 0; q < 3; q++) s += es14(E.bf[q].real()) + es14(E.bf[q].imag());
      for (auto& u : E.Us) s += es14(u.real()) + es14(u.imag());
      for (auto& d : E.ds) s += es14(d.real()) + es14(d.imag());
      return s + "\n";
    };
    auto heat_line = [&](int in, const Eigen& E) {
      std::string s = head(in);
      for (double p : E.Ps) s += es14(p);
      return s + es14(E.W_EM) + "\n";
    };
    auto mech_line = [&](int in, const Eigen& E) {
      std::string s = head(in);
      for (double p : E.Ps_split) s += es14(p);
      return s + "\n";
    };
    const bool want = sc->eigen || sc->heat;
    size_t row = 0;
    auto record = [&](int in) {
      if (!rows_out) return;
      double* r = rows_out + 4 * ((size_t)row * nroots + in);
      r[0] = kperp; r[1] = kpar; r[2] = wroots[2 * in]; r[3] = wroots[2 * in + 1];
    };
    for (int in = 0; in < nroots; in++) {
      if (prefix) append_line(fname("scan", in), head(in) + "\n", true);
      record(in);
      if (want) {
        Eigen E;
        calc_eigen_impl(cplx(wroots[2 * in], wroots[2 * in + 1]), nspec, ns, qs, current_int, kperp, kpar, vA,
                        sc->eigen != 0, sc->heat != 0, E);
        if (prefix && sc->eigen) append_line(fname("eigen", in), eigen_line(in, E), true);
        if (prefix && sc->heat) {
          append_line(fname("heat", in), heat_line(in, E), true);
          append_line(fname("heat_mech", in), mech_line(in, E), true);
        }
      }
    }
    row = 1;
    const double kperp_last = kperp, kpar_last = kpar;
    const double theta_0 = atan(kperp_last / kpar_last), k_0 = sqrt(kperp_last * kperp_last + kpar_last * kpar_last);
    for (int it = 1; it <= nt; it++) {
      switch (sc->type) {
        case 0:
          if (sc->log_scan) {
            kperp = pow(10.0, log10(kperp_last) + sc->diff * it);
            kpar = pow(10.0, log10(kpar_last) + sc->diff2 * it);
          } else {
            kperp = kperp_last + sc->diff * it;
            kpar = kpar_last + sc->diff2 * it;
          }
          break;
        case 1: {
          double theta_1 = sc->log_scan ? pow(10.0, log10(theta_0) + sc->diff * it) : theta_0 + sc->diff * it;
          kperp = k_0 * sin(theta_1);
          kpar = k_0 * cos(theta_1);
          break;
        }
        case 2: {
          double k_tmp = sc->log_scan ? pow(10.0, log10(k_0) + sc->diff * it) : k_0 + sc->diff * it;
          kperp = k_tmp * sin(theta_0);
          kpar = k_tmp * cos(theta_0);
          break;
        }
        case 3: kperp = sc->log_scan ? pow(10.0, log10(kperp_last) + sc->diff * it) : kperp_last + sc->diff * it; break;
        default: kpar = sc->log_scan ? pow(10.0, log10(kpar_last) + sc->diff * it) : kpar_last + sc->diff * it; break;
      }
      int rc = alps_b200_set_k(kperp, kpar, nullptr);
      if (rc) throw DispError{rc};
      bool alljump = false;
      for (int in = 0; in < nroots; in++) alljump = alljump || jump[in];
      if (!alljump) throw DispError{9};   // alps_error(9)
      const bool out_step = (it % sc->n_res) == 0;
      std::vector<Eigen> Es(nroots);
      std::vector<cplx> sol(nroots);
      {
        std::vector<std::function<void()> > jobs;
        for (int in = 0; in < nroots; in++) {
          if (!jump[in]) continue;
          jobs.push_back([&, in] {
            sol[in] = solve_root(cplx(wroots[2 * in], wroots[2 * in + 1]), *o);
            if (out_step && want)
              calc_eigen_impl(sol[in], nspec, ns, qs, current_int, kperp, kpar, vA, sc->eigen != 0, sc->heat != 0, Es[in]);
          });
        }
        for_each_root(jobs);
      }
      for (int in = 0; in < nroots; in++) {
        if (!jump[in]) continue;
        const cplx omega = sol[in];
        wroots[2 * in] = omega.real();
        wroots[2 * in + 1] = omega.imag();
        const Eigen& E = Es[in];
        if (std::isnan(omega.real())) jump[in] = false;
        for (int imm = 0; imm < in; imm++) {
          cplx a(wroots[2 * in], wroots[2 * in + 1]), b(wroots[2 * imm], wroots[2 * imm + 1]);
          if (std::abs(a - b) < o->D_gap) {
            wroots[2 * in] = wroots[2 * in + 1] = 0.0;
            jump[in] = false;
          }
        }
        if (out_step) {
          if (prefix) {
            append_line(fname("scan", in), head(in) + "\n");
            if (sc->eigen) append_line(fname("eigen", in), eigen_line(in, E));
            if (sc->heat) {
              append_line(fname("heat", in), heat_line(in, E));
              append_line(fname("heat_mech", in), mech_line(in, E));
            }
          }
          record(in);
        }
      }
      if ((it % sc->n_res) == 0) row++;
    }
    *kperp_io = kperp;
    *kpar_io = kpar;
  });
}

// k of scan step `it` starting from (kperp0, kpar0): the select case blocks of om_scan / om_double_scan
static void scan_k_at(const alps_b200_scan* sc, int it, double kperp0, double kpar0, double* kperp, double* kpar) {
  const double theta_0 = atan(kperp0 / kpar0), k_0 = sqrt(kperp0 * kperp0 + kpar0 * kpar0);
  switch (sc->type) {
    case 0:
      if (sc->log_scan) {
        *kperp = pow(10.0, log10(kperp0) + sc->diff * it);
        *kpar = pow(10.0, log10(kpar0) + sc->diff2 * it);
      } else {
        *kperp = kperp0 + sc->diff * it;
        *kpar = kpar0 + sc->diff2 * it;
      }
      break;
    case 1: {
      double theta_1 = sc->log_scan ? pow(10.0, log10(theta_0) + sc->diff * it) : theta_0 + sc->diff * it;
      *kperp = k_0 * sin(theta_1);
      *kpar = k_0 * cos(theta_1);
      break;
    }
    case 2: {
      double k_tmp = sc->log_scan ? pow(10.0, log10(k_0) + sc->diff * it) : k_0 + sc->diff * it;
      *kperp = k_tmp * sin(theta_0);
      *kpar = k_tmp * cos(theta_0);
      break;
    }
    case 3: *kperp = sc->log_scan ? pow(10.0, log10(kperp0) + sc->diff * it) : kperp0 + sc->diff * it; break;
    default: *kpar = sc->log_scan ? pow(10.0, log10(kpar0) + sc->diff * it) : kpar0 + sc->diff * it; break;
  }
}

// replaces: om_double_scan, src/ALPS_fns.f90:2904-3591
int alps_b200_om_double_scan(const alps_b200_scan* sc1, const alps_b200_scan* sc2, int nroots, double* wroots,
                             const alps_b200_solver_opts* o, int nspec, const double* ns, const double* qs,
                             const double* current_int, double vA, double* kperp_io, double* kpar_io,
                             const char* prefix, double* rows_out) {
  return guarded([&] {
    static const char* ids[5] = {"k1_k2", "theta", "kcstq", "kperp", "kpara"};
    if (sc1->type == sc2->type) throw DispError{5};          // alps_error(5)
    if (sc1->type == 0 || sc2->type == 0) throw DispError{6};  // alps_error(6)
    double kperp = *kperp_io, kpar = *kpar_io;
    const int nt = sc1->n_out * sc1->n_res, nt2 = sc2->n_out * sc2->n_res;
    const bool want = sc1->eigen || sc1->heat;
    std::vector<bool> jump(nroots, true);
    auto fname = [&](const char* kind, int in) {
      char b[1024];
      snprintf(b, sizeof(b), "%s.%s_%s_%s.root_%d", prefix, kind, ids[sc1->type], ids[sc2->type], in + 1);
      return std::string(b);
    };
    auto head = [&](int in) { return es14(kperp) + es14(kpar) + es14(wroots[2 * in]) + es14(wroots[2 * in + 1]); };
    auto lines = [&](int in, const Eigen& E, std::string& le, std::string& lh, std::string& lm) {
      le = lh = lm = head(in);
      for (int q = 0; q < 3; q++) le += es14(E.ef[q].real()) + es14(E.ef[q].imag());
      for (int q = 0; q < 3; q++) le += es14(E.bf[q].real()) + es14(E.bf[q].imag());
      for (auto& u : E.Us) le += es14(u.real()) + es14(u.imag());
      for (auto& d : E.ds) le += es14(d.real()) + es14(d.imag());
      for (double p : E.Ps) lh += es14(p);
      lh += es14(E.W_EM);
      for (double p : E.Ps_split) lm += es14(p);
      le += "\n"; lh += "\n"; lm += "\n";
    };
    if (prefix)
      for (int in = 0; in < nroots; in++) {
        append_line(fname("scan", in), "", true);
        if (sc1->eigen) append_line(fname("eigen", in), "", true);
        if (sc1->heat) {
          append_line(fname("heat", in), "", true);
          append_line(fname("heat_mech", in), "", true);
        }
      }
    const double kperp_last = kperp, kpar_last = kpar;
    std::vector<std::complex<float> > omlast(nroots);   // single-precision COMPLEX in the reference (line 2979)
    for (int in = 0; in < nroots; in++) omlast[in] = std::complex<float>((float)wroots[2 * in], (float)wroots[2 * in + 1]);
    size_t row = 0;
    const int cols = sc2->n_out + 1;
    for (int it = 0; it <= nt; it++) {
      scan_k_at(sc1, it, kperp_last, kpar_last, &kperp, &kpar);
      int rc = alps_b200_set_k(kperp, kpar, nullptr);
      if (rc) throw DispError{rc};
      bool alljump = false;
      for (int in = 0; in < nroots; in++) alljump = alljump || jump[in];
      if (!alljump) throw DispError{9};
      for (int in = 0; in < nroots; in++) {
        if (!jump[in]) continue;
        cplx omega((double)omlast[in].real(), (double)omlast[in].imag());
        wroots[2 * in] = omega.real();
        wroots[2 * in + 1] = omega.imag();
        omega = solve_root(omega, *o);
        wroots[2 * in] = omega.real();
        wroots[2 * in + 1] = omega.imag();
        omlast[in] = std::complex<float>((float)omega.real(), (float)omega.imag());
        if (std::isnan(omega.real())) jump[in] = false;
        for (int imm = 0; imm < in; imm++)
          if (std::abs(cplx(wroots[2 * in], wroots[2 * in + 1]) - cplx(wroots[2 * imm], wroots[2 * imm + 1])) < o->D_gap) {
            wroots[2 * in] = wroots[2 * in + 1] = 0.0;
            jump[in] = false;
          }
      }
      if (it % sc1->n_res != 0) continue;
      double kperpi = kperp, kpari = kpar;
      size_t col = 0;
      for (int it2 = 0; it2 <= nt2; it2++) {
        if (it2 == 0)
          for (int in = 0; in < nroots; in++) {
            wroots[2 * in] = (double)omlast[in].real();
            wroots[2 * in + 1] = (double)omlast[in].imag();
          }
        scan_k_at(sc2, it2, kperpi, kpari, &kperp, &kpar);
        rc = alps_b200_set_k(kperp, kpar, nullptr);
        if (rc) throw DispError{rc};
        const bool out_step = (it2 % sc2->n_res) == 0;
        for (int in = 0; in < nroots; in++) {
          if (!jump[in]) continue;
          cplx omega = solve_root(cplx(wroots[2 * in], wroots[2 * in + 1]), *o);
          wroots[2 * in] = omega.real();
          wroots[2 * in + 1] = omega.imag();
          const cplx tmp = disp1(omega);
          Eigen E;
          if (out_step && want)
            calc_eigen_impl(omega, nspec, ns, qs, current_int, kperp, kpar, vA, sc1->eigen != 0, sc1->heat != 0, E);
          if (std::isnan(omega.real())) jump[in] = false;
          if (std::abs(tmp) > 1.e100) jump[in] = false;
          for (int imm = 0; imm < in; imm++)
            if (std::abs(cplx(wroots[2 * in], wroots[2 * in + 1]) - cplx(wroots[2 * imm], wroots[2 * imm + 1])) < o->D_gap) {
              wroots[2 * in] = wroots[2 * in + 1] = 0.0;
              jump[in] = false;
            }
          if (out_step) {
            if (prefix) {
              append_line(fname("scan", in), head(in) + "\n");
              if (want) {
                std::string le, lh, lm;
                lines(in, E, le, lh, lm);
                if (sc1->eigen) append_line(fname("eigen", in), le);
                if (sc1->heat) {
                  append_line(fname("heat", in), lh);
                  append_line(fname("heat_mech", in), lm);
                }
              }
            }
            if (rows_out) {
              double* r = rows_out + 4 * ((row * cols + col) * nroots + in);
              r[0] = kperp; r[1] = kpar; r[2] = wroots[2 * in]; r[3] = wroots[2 * in + 1];
            }
          }
        }
        if (out_step) col++;
      }
      if (prefix)
        for (int in = 0; in < nroots; in++) {
          append_line(fname("scan", in), "\n");
          if (sc1->eigen) append_line(fname("eigen", in), "\n");
          if (sc1->heat) {
            append_line(fname("heat", in), "\n");
            append_line(fname("heat_mech", in), "\n");
          }
        }
      row++;
      // reset the second scan's variable to its start value (lines 3535-3580)
      scan_k_at(sc2, 0, kperp_last, kpar_last, &kperp, &kpar);
    }
    *kperp_io = kperp;
    *kpar_io = kpar;
  });
}

}  // extern "C"
