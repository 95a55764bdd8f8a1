"""FieldTimeSeries window on the device: which time index lives in which slot of the device
ring, and what to load ahead of time.

The reference keeps `Nt_mem` consecutive slices of each prescribed series in memory
(`JRA55NetCDFBackend(start, length)`); `update_state!(::PrescribedAtmosphere)`
(src/Atmospheres/prescribed_atmosphere.jl:154-162) calls Oceananigans' `update_field_time_series!` for every
series, and when the interpolating indices (n1, n2) are no longer both inside the window, the whole window is
re-read and re-set (`set!(fts)`, src/DataWrangling/JRA55/JRA55_field_time_series.jl:60-76, :78-124) before the step
can go on: a periodic host stall.

Here the device array of each series is a ring of `n_slots` slices (csrc/ne_series_ring.cu).  After the
interpolation kernel of a step has been enqueued, the slices the NEXT steps will need are loaded one at a time into
slots that hold neither n1 nor n2, on the ring's own copy stream, behind the step's kernels; when the clock crosses
into the next interval its slice is already resident and the step only waits for an event that has long fired.
`NeTimeInterp.m1/m2` carry the slots (the reference's `memory_index`).  Values interpolated through a window equal
those interpolated from the fully in-memory series bit for bit (same slices, same kernel).

`WindowPolicy` is the bookkeeping alone (no device; unit-tested on CPU); `SeriesWindow` owns the rings, the pinned
host store and the native handle.
"""
import ctypes as C

import numpy as np

from . import abi as A

CONVERSIONS = {   # convert_units (src/DataWrangling/metadata_field.jl:486-525): name -> (kind, a, b)
    None: (A.NE_CONV_NONE, 0.0, 0.0),
    "InverseSign": (A.NE_CONV_NEGATE, 0.0, 0.0),
    "Kelvin": (A.NE_CONV_SUB, 273.15, 0.0),
    "Celsius": (A.NE_CONV_ADD, 273.15, 0.0),
    "Millibar": (A.NE_CONV_MUL, 100.0, 0.0),
    "MillimetersPerHour": (A.NE_CONV_DIV, 3600.0, 0.0),
    "MetersPerHour": (A.NE_CONV_MUL_DIV, 1000.0, 3600.0),
    "JoulesPerSquareMeterPerHour": (A.NE_CONV_DIV, 3600.0, 0.0),
    "InverseGravity": (A.NE_CONV_DIV, 9.80665, 0.0),
    "CentimetersPerSecond": (A.NE_CONV_DIV, 100.0, 0.0),
    "GramPerKilogram": (A.NE_CONV_DIV, 1e3, 0.0),
    "GramPerKilogramMinus35": (A.NE_CONV_ADD, 35.0, 0.0),
}


class WindowPolicy:
    """Slot bookkeeping of one ring: `resident[s]` is the (1-based) time index slot s holds, or None."""

    def __init__(self, n_times, n_slots, time_indexing="cyclical", lookahead=None):
        if n_slots < 2:
            raise ValueError("a window needs at least 2 slots (both interpolating slices are read by one kernel)")
        self.nt, self.n_slots, self.time_indexing = int(n_times), int(n_slots), time_indexing
        self.lookahead = max(0, self.n_slots - 2) if lookahead is None else int(lookahead)
        self.resident = [None] * self.n_slots
        self.where = {}

    def upcoming(self, n2, count):
        """The next `count` time indices after n2 in the order the clock will need them."""
        out = []
        n = n2
        for _ in range(count):
            n += 1
            if n > self.nt:
                if self.time_indexing != "cyclical":
                    break
                n = 1
            if n in out or n == n2:
                break
            out.append(n)
        return out

    def _distance_ahead(self, n, n1):
        if self.time_indexing == "cyclical":
            return (n - n1) % self.nt
        return n - n1 if n >= n1 else self.nt + (n1 - n)   # the past is the best victim

    def _victim(self, protected, n1):
        free = [s for s in range(self.n_slots) if self.resident[s] is None]
        if free:
            return free[0]
        cands = [s for s in range(self.n_slots) if self.resident[s] not in protected]
        if not cands:
            return None
        return max(cands, key=lambda s: self._distance_ahead(self.resident[s], n1))

    def _assign(self, n, slot):
        old = self.resident[slot]
        if old is not None:
            self.where.pop(old, None)
        self.resident[slot] = n
        self.where[n] = slot

    def demand(self, n1, n2):
        """Loads that MUST happen before the step: [(n, slot)] for those of n1, n2 that are not resident."""
        loads = []
        for n in dict.fromkeys((n1, n2)):
            if n not in self.where:
                slot = self._victim({n1, n2}, n1)
                self._assign(n, slot)
                loads.append((n, slot))
        return loads

    def prefetch(self, n1, n2):
        """Loads worth starting now for the next intervals: never into the slots of n1, n2 or of a nearer index."""
        ahead = self.upcoming(n2, self.lookahead)
        protected = {n1, n2}
        loads = []
        for n in ahead:
            protected.add(n)
            if n in self.where:
                continue
            slot = self._victim(set(protected) | set(ahead), n1)
            if slot is None:
                break
            self._assign(n, slot)
            loads.append((n, slot))
        return loads


def infer_longitudinal_period(lon_centers):
    """360 if the file's longitudes span the full globe (cyclic), else None (set_region_data.jl:121-126)."""
    lam = np.asarray(lon_centers, dtype=np.float64)
    if lam.size < 2:
        return None
    delta = lam[1] - lam[0]
    span = lam[-1] - lam[0] + delta
    return 360 if np.isclose(span, 360.0) else None


def bracket_with_weight(coords, x, period=None):
    """Cyclic-aware bracketing of `x` by cell centres (set_region_data.jl:130-150): 1-based (i_minus, i_plus, w) as the
    reference returns them; with `period`, the cell between coords[end] and coords[1] + period is the wrap cell (n, 1, w)."""
    c = np.asarray(coords, dtype=np.float64)
    n = c.size
    if n <= 1:
        return 1, 1, 0.0
    x = float(x)
    if period is not None:
        x = c[0] + np.mod(x - c[0], period)
        if x > c[-1]:
            delta = (c[0] + period) - c[-1]
            return n, 1, float(np.clip((x - c[-1]) / delta, 0.0, 1.0))
    ip = int(np.searchsorted(c, x, side="left")) + 1        # searchsortedfirst, 1-based
    ip = min(max(ip, 2), n)
    im = ip - 1
    delta = c[ip - 1] - c[im - 1]
    w = 0.0 if delta == 0 else (x - c[im - 1]) / delta
    return im, ip, float(np.clip(w, 0.0, 1.0))


class ColumnRegion:
    """DataWrangling.Column(longitude, latitude; interpolation = Linear() | Nearest()) resolved against a file's cell centres
    (region_info(::Column), set_region_data.jl:113-118): what the ring's slot fill needs to blend four file cells into the one
    cell of a column series."""

    def __init__(self, lon_centers, lat_centers, longitude, latitude, interpolation="linear"):
        if interpolation not in ("linear", "nearest"):
            raise ValueError("Column interpolation is 'linear' or 'nearest'")
        self.i_minus, self.i_plus, self.wx = bracket_with_weight(lon_centers, longitude, period=infer_longitudinal_period(lon_centers))
        self.j_minus, self.j_plus, self.wy = bracket_with_weight(lat_centers, latitude)   # latitude is never cyclic
        self.interpolation = interpolation


class SeriesWindow:
    """Device rings of the series of one prescribed component (all on one source grid and one time axis).

    raw: dict name -> host array (nt, ny_file, nx_file) of the source element type, the decoded file variable (no halos,
    x fastest; what `ds[name][:, :, nn]` returns in JRA55_field_time_series.jl:67, transposed to row-major); kept in
    pinned memory.  `region_offset` = (di, dj) of a BoundingBox region inside the file (region_info, set_region_data.jl:88-111);
    a whole-globe file with one latitude less / more than the grid is read with ShiftSouth / AverageNorthSouth mangling.
    `series[name]` is the device ring (n_slots, ny + 2hy, nx + 2hx) a PrescribedAtmosphere / PrescribedRadiation /
    PrescribedLand field points to.
    """

    def __init__(self, backend, lib, grid, times, raw, n_slots=4, time_indexing="cyclical", conversions=None,
                 missing_values=None, periodic_x=True, lookahead=None, region_offset=(0, 0), column=None):
        if not backend.is_device:
            raise RuntimeError("SeriesWindow needs the CUDA library and device arrays (there is no CPU fallback)")
        self.backend, self.lib, self.grid = backend, lib, grid
        self.times, self.time_indexing = np.asarray(times, dtype=np.float64), time_indexing
        self.names = list(raw)
        if not 1 <= len(self.names) <= A.NE_RING_MAX_SERIES:
            raise ValueError(f"a window holds 1 to {A.NE_RING_MAX_SERIES} series")
        nt = len(self.times)
        self.n_slots = int(min(n_slots, nt)) if nt >= 2 else 2
        self.policy = WindowPolicy(nt, self.n_slots, time_indexing, lookahead)
        npd = np.float64 if grid.FT == "f64" else np.float32
        torch = backend.torch
        self.host = {}
        raw_shape = None
        for k in self.names:
            a = np.ascontiguousarray(np.asarray(raw[k], dtype=npd))
            if a.ndim != 3 or a.shape[0] != nt or (raw_shape is not None and a.shape != raw_shape):
                raise ValueError(f"raw series {k}: expected {nt} slices of one common (ny_file, nx_file) shape, got {a.shape}")
            raw_shape = a.shape
            self.host[k] = torch.from_numpy(a).pin_memory()
        raw_ny, raw_nx = raw_shape[1:]
        # mangling_for (src/DataWrangling/set_region_data.jl:153-158): a file with one latitude less / more than the grid
        whole = tuple(region_offset) == (0, 0)
        mangling = A.NE_MANGLE_SHIFT_SOUTH if whole and column is None and raw_ny == grid.ny - 1 else \
            A.NE_MANGLE_AVERAGE_NORTH_SOUTH if whole and column is None and raw_ny == grid.ny + 1 else A.NE_MANGLE_NONE
        self.mangling = mangling
        self.series = {k: backend.zeros((self.n_slots,) + tuple(grid.shape), grid.FT) for k in self.names}
        d = self.desc = A.NeSeriesRingDesc()
        d.n_series, d.n_slots, d.dtype, d.periodic_x = len(self.names), self.n_slots, A.NE_F64 if grid.FT == "f64" else A.NE_F32, int(periodic_x)
        d.nx, d.ny, d.hx, d.hy = grid.nx, grid.ny, grid.hx, grid.hy
        d.raw_nx, d.raw_ny, d.di, d.dj = raw_nx, raw_ny, int(region_offset[0]), int(region_offset[1])
        if column is not None:    # a ColumnRegion: the series is 1 x 1, its value the blend of four file cells
            if (grid.nx, grid.ny) != (1, 1):
                raise ValueError("a Column region fills a 1 x 1 series")
            d.region_kind = A.NE_REGION_COLUMN
            d.column_interpolation = A.NE_COLUMN_NEAREST if column.interpolation == "nearest" else A.NE_COLUMN_LINEAR
            d.col_i_minus, d.col_i_plus = column.i_minus - 1, column.i_plus - 1
            d.col_j_minus, d.col_j_plus = column.j_minus - 1, column.j_plus - 1
            npf = np.float64 if grid.FT == "f64" else np.float32
            d.col_wx, d.col_wy = float(npf(column.wx)), float(npf(column.wy))   # ColumnInfo holds FT(wx), FT(wy)
        for i, k in enumerate(self.names):
            d.ring[i] = backend.ptr(self.series[k])
            d.mangling[i] = mangling
            kind, a, b = CONVERSIONS[(conversions or {}).get(k)]
            d.conv_kind[i], d.conv_a[i], d.conv_b[i] = kind, a, b
            if missing_values and k in missing_values:
                d.has_missing[i], d.missing_value[i] = 1, float(missing_values[k])
        self.handle = C.c_void_p()
        self._check(lib.dll.ne_series_ring_create(C.byref(self.handle), C.byref(d)))
        self._slice_bytes = raw_ny * raw_nx * (8 if grid.FT == "f64" else 4)
        self.demand_loads = 0      # slices a step had to wait for (the reference's stall, per slice)
        self.prefetched = 0        # slices loaded behind a step's kernels
        self._current = None

    def __getitem__(self, name):
        return self.series[name]

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.last_error())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.dll.ne_series_ring_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _load(self, n, slot):
        ptrs = (C.c_void_p * len(self.names))(*[self.host[k].data_ptr() + (n - 1) * self._slice_bytes for k in self.names])
        self._check(self.lib.dll.ne_series_ring_load(self.handle, C.c_int32(slot), ptrs))

    def time_interp(self, t, stream):
        """(ñ, m1, m2, same) for the interpolation descriptor at time t (m: 1-based ring slots).  Loads whichever of
        the two slices is not resident (only the first step, or a jump of the clock) and makes `stream` wait for both."""
        from .interface import interpolating_time_indices
        frac, n1, n2 = interpolating_time_indices(self.times, t, self.time_indexing)
        for n, slot in self.policy.demand(n1, n2):
            self._load(n, slot)
            self.demand_loads += 1
        s1, s2 = self.policy.where[n1], self.policy.where[n2]
        for s in {s1, s2}:
            self._check(self.lib.dll.ne_series_ring_acquire(self.handle, C.c_int32(s), C.c_void_p(stream)))
        self._current = (n1, n2)
        return frac, s1 + 1, s2 + 1, int(n1 == n2)

    def after_launch(self, stream):
        """Call once the step's interpolation kernel(s) have been enqueued on `stream`: later loads wait for them, and
        the slices of the coming intervals start loading now."""
        if self._current is None:
            return
        self._check(self.lib.dll.ne_series_ring_release(self.handle, C.c_void_p(stream)))
        for n, slot in self.policy.prefetch(*self._current):
            self._load(n, slot)
            self.prefetched += 1
