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


def pointwise_rel(a, b, floor=1e-6):
    """The criterion of the BASELINE north_star read pointwise: max over points of |a - b| / max(|b|, floor * scale),
    scale = max |b| over the field.  A point passes `tol` when |a - b| <= tol * max(|b|, floor * scale): relative
    everywhere except within `floor` of a sign change of the field, where the denominator stops shrinking."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = float(np.nanmax(np.abs(b))) or 1.0
    den = np.maximum(np.abs(b), floor * scale)
    with np.errstate(invalid="ignore"):
        r = np.abs(a - b) / den
    return float(np.nanmax(r))


def pointwise_exceed(a, b, tol, floor=1e-6, near_zero=1e-4):
    """(number of points that miss `tol` under the pointwise criterion, how many of those lie where the field is smaller
    than `near_zero` times its scale, i.e. next to a sign change)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0, 0
    scale = float(np.nanmax(np.abs(b))) or 1.0
    den = np.maximum(np.abs(b), floor * scale)
    with np.errstate(invalid="ignore"):
        bad = np.abs(a - b) > tol * den
    return int(bad.sum()), int((bad & (np.abs(b) < near_zero * scale)).sum())


class ParityLog:
    """Collects what the parity tests measured (max pointwise relative error per field, trip-count mismatch rate, number
    of points that ran into maxiter) — printed, and written to gpurun_out/parity_r02.jsonl when that directory exists,
    so the numbers behind each assertion are on record."""
    rows = []

    @classmethod
    def add(cls, test, **kw):
        import json
        import os
        row = dict(test=test, **kw)
        cls.rows.append(row)
        print("PARITY", json.dumps(row))
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        out = os.path.join(root, "gpurun_out")
        if os.path.isdir(out):
            with open(os.path.join(out, "parity_r02.jsonl"), "a") as f:
                f.write(json.dumps(row) + "\n")


def trip_statistics(ref_it, dev_it, maxiter):
    """(mismatch rate over solved points, max |difference|, points at maxiter in the oracle, in the device result)."""
    r = np.asarray(ref_it).astype(np.int64)
    d = np.asarray(dev_it).astype(np.int64)
    solved = r > 0
    n = max(int(solved.sum()), 1)
    return dict(trip_mismatch_rate=float((r != d)[solved].sum()) / n, trip_max_abs_diff=int(np.abs(r - d).max()) if r.size else 0,
                maxiter_points_ref=int((r >= maxiter).sum()), maxiter_points_dev=int((d >= maxiter).sum()),
                solved_points=int(solved.sum()), mean_trips=float(r[solved].mean()) if solved.any() else 0.0)


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


def compare_pointwise(ref_bag, dev_bag, grid, backend, names=None, with_halo_ring=True, mask=None, tol=1e-10, floor=1e-6):
    """name -> dict(pw = max pointwise relative error (pointwise_rel), exceed = points missing `tol`, n = points compared,
    exact = bit-identical).  NaNs must sit at identical points."""
    out = {}
    for n in (names or ref_bag.names()):
        r = _window(getattr(ref_bag, n), grid, with_halo_ring)
        d = _window(backend.to_numpy(getattr(dev_bag, n)), grid, with_halo_ring)
        exact = bool(np.array_equal(d, r, equal_nan=True))
        assert np.array_equal(np.isnan(d), np.isnan(r)), f"{n}: NaN pattern differs"
        if mask is not None:
            r, d = r[mask], d[mask]
        ex, ex_near_zero = pointwise_exceed(d, r, tol, floor)
        out[n] = dict(pw=pointwise_rel(d, r, floor), exceed=ex, exceed_near_zero=ex_near_zero, n=int(np.asarray(r).size), exact=exact)
    return out


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
