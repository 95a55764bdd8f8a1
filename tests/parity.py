"""Helpers shared by the parity tests: build the same seeded case on the oracle (host) and on the
CUDA library (device) and compare field by field."""
import numpy as np

import ne_b200
from numericalearth_jl_b200 import synthetic


def build_pair(config, oracle_lib, cuda_backend, **kw):
    host = ne_b200.NumpyHostBackend()
    ref = synthetic.build_case(config, host, lib=oracle_lib, with_iterations=True, **kw)
    dev = synthetic.build_case(config, cuda_backend, with_iterations=True, **kw)
    return ref, dev


def rel_err(a, b):
    """max |a-b| / max(|b|, floor) with the floor at 1e-300 relative to the field scale."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), np.abs(a))
    d = np.abs(a - b)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(scale > 0, d / scale, 0.0)
    return float(np.nanmax(r)) if r.size else 0.0


def field_rel_err(a, b):
    """Error relative to the field's own magnitude scale (robust near sign changes)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = float(np.nanmax(np.abs(b))) or 1.0
    return float(np.nanmax(np.abs(a - b))) / s


def _window(a, grid, with_halo_ring):
    if with_halo_ring:
        return grid.interior(a)
    return a[grid.hy:grid.hy + grid.ny, grid.hx:grid.hx + grid.nx]


def converged_mask(iterations, grid, maxiter, with_halo_ring=True, dilate=True):
    """True where the oracle's solve stopped before maxiter.  Points that hit maxiter sit on a limit
    cycle of the fixed-point map, which amplifies last-ulp libm differences; they are compared
    separately with a loose tolerance.  `dilate` also drops their (i-1, j-1) stencil neighbours
    for the net-flux fields."""
    bad = _window(np.asarray(iterations), grid, True) >= maxiter
    if dilate:
        b = bad.copy()
        b[:, 1:] |= bad[:, :-1]
        b[1:, :] |= bad[:-1, :]
        b[:, :-1] |= bad[:, 1:]
        b[:-1, :] |= bad[1:, :]
        bad = b
    if not with_halo_ring:
        bad = bad[1:-1, 1:-1]
    return ~bad


def compare_fields(ref_bag, dev_bag, grid, backend, names=None, with_halo_ring=True, mask=None):
    """name -> (max pointwise relative error, error relative to the field scale, bit-exact?).
    NaNs must sit at identical points."""
    out = {}
    for n in (names or ref_bag.names()):
        r = _window(getattr(ref_bag, n), grid, with_halo_ring)
        d = _window(backend.to_numpy(getattr(dev_bag, n)), grid, with_halo_ring)
        exact = bool(np.array_equal(d, r, equal_nan=True))
        assert np.array_equal(np.isnan(d), np.isnan(r)), f"{n}: NaN pattern differs"
        if mask is not None:
            r, d = r[mask], d[mask]
        out[n] = (rel_err(d, r), field_rel_err(d, r), exact)
    return out
