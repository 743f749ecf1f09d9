"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def tensor_err(a, b):
    """max |a-b| relative to the largest entry of the reference tensor b (entries of chi span
    many decades; tiny off-diagonals are judged against the tensor scale)."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / scale) if scale > 0 else float(np.max(np.abs(a)))


def det_scale(wave):
    """scale of the summed products in D = det-like expression (src/ALPS_fns.f90:622-624): near a
    root D is a cancellation, so |dD| is judged against the size of its terms."""
    w = np.abs(np.asarray(wave))
    return float(w[0, 0] * (w[1, 1] * w[2, 2] + w[1, 2] ** 2) + 2 * w[0, 1] * w[1, 2] * w[0, 2]
                 + w[0, 2] ** 2 * w[1, 1] + w[0, 1] ** 2 * w[2, 2])


def omega_samples(seed, n, re_range, im_range):
    rng = np.random.default_rng(seed)
    re = rng.uniform(re_range[0], re_range[1], n)
    im = rng.uniform(im_range[0], im_range[1], n)
    return re + 1j * im


def chi_err(a, b):
    """Mixed criterion for one species' chi0 tensor: element-wise relative error, with an absolute
    floor of 1e-3 * tol-scale for entries that are themselves cancellations between the +n and -n
    harmonics (e.g. electron chi_xx at k_perp rho << 1): returns max |a-b| / (|b| + 1e-3 max|b|)."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / (np.abs(b) + 1e-3 * np.max(np.abs(b)))))


def wave_scale(chi0, om, vA, kperp, kpar):
    """Magnitude of the terms summed into each wave(i,j) (src/ALPS_fns.f90:566-612, kperp_norm):
    sum_s |chi_s| + |om vA|^2 on the diagonal + the refraction terms.  Near a root wave(1,1) is
    itself a cancellation (eps_xx - kpar^2), so its error is judged against this scale."""
    n2 = abs(om * om * vA * vA)
    ws = np.sum(np.abs(np.asarray(chi0)), axis=0) * n2
    u = abs(om * vA) ** 2
    ws[0, 0] += u + kpar ** 2
    ws[1, 1] += u + kpar ** 2 + kperp ** 2
    ws[2, 2] += u + kperp ** 2
    ws[0, 2] += kperp * kpar
    ws[2, 0] += kperp * kpar
    return ws


def scaled_err(a, b, scale):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / scale))
