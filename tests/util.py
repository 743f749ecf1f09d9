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
