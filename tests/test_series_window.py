"""the FieldTimeSeries window on the device.

Reference behaviour restated: update_state!(::PrescribedAtmosphere) (src/Atmospheres/prescribed_atmosphere.jl:154-162) ->
update_field_time_series! -> set!(fts) (src/DataWrangling/JRA55/JRA55_field_time_series.jl:60-76): raw file slices go
through _set_region_kernel! (src/DataWrangling/set_region_data.jl:200-205) into the interior and fill_halo_regions!
fills the halos (test/test_jra55.jl:40-47 checks fts[Nx+1, ...] == fts[1, ...]).

CPU: the slot bookkeeping (WindowPolicy) and the oracle's slot fill against numpy.  GPU: the ring's slot-fill kernel against
the oracle bit for bit, and a windowed PrescribedAtmosphere + PrescribedRadiation stepped across many intervals (and the
cyclical wrap) against the same series fully in memory, bit for bit.
"""
import ctypes as C

import numpy as np
import pytest

import ne_b200
from ne_b200 import abi as A
from numericalearth_jl_b200 import synthetic
from numericalearth_jl_b200.series_window import (CONVERSIONS, ColumnRegion, SeriesWindow, WindowPolicy, bracket_with_weight,
                                                   infer_longitudinal_period)

NPD = {"f64": np.float64, "f32": np.float32}


# ------------------------------------------------------------------------------------------------- policy (CPU)
def _walk(policy, times, ts, indexing):
    """Drive the policy the way SeriesWindow does; returns (demand loads per step, prefetch loads per step)."""
    dem, pre = [], []
    for t in ts:
        _, n1, n2 = ne_b200.interpolating_time_indices(times, t, indexing)
        before = dict(policy.where)
        d = policy.demand(n1, n2)
        assert policy.resident[policy.where[n1]] == n1 and policy.resident[policy.where[n2]] == n2
        assert n1 == n2 or policy.where[n1] != policy.where[n2]
        for n, _slot in d:
            assert n not in before
        p = policy.prefetch(n1, n2)
        # a prefetch never replaces the two slices the kernel that was just enqueued reads
        assert policy.resident[policy.where[n1]] == n1 and policy.resident[policy.where[n2]] == n2
        for n, slot in p:
            assert slot not in (before.get(n1, -1), before.get(n2, -1)) or before.get(n1) is None
        # the map and the slot table stay each other's inverse
        assert {n: s for s, n in enumerate(policy.resident) if n is not None} == policy.where
        dem.append(len(d))
        pre.append(len(p))
    return dem, pre


@pytest.mark.parametrize("nt,n_slots", [(8, 4), (8, 3), (5, 5), (24, 6), (3, 3)])
def test_cyclical_clock_never_waits_after_the_first_step(nt, n_slots):
    times = np.arange(nt) * 10800.0
    policy = WindowPolicy(nt, n_slots, "cyclical")
    ts = np.arange(0, 3.2 * nt * 10800.0, 1800.0)   # 6 steps per interval, three times around the year
    dem, pre = _walk(policy, times, ts, "cyclical")
    assert dem[0] == 2 and sum(dem[1:]) == 0          # only the very first step loads on demand
    n_intervals = int(ts[-1] // 10800.0)
    if nt > n_slots:
        assert sum(pre) >= n_intervals - 1            # one new slice per interval, each ahead of its interval
    else:
        assert sum(pre) == nt - 2                     # the whole series fits: every slice is loaded exactly once


def test_two_slots_degrade_to_one_demand_load_per_interval():
    """n_slots = 2 leaves nowhere to prefetch into: every new interval waits for ONE slice — still not for a whole
    window like the reference's reload."""
    nt = 6
    times = np.arange(nt) * 10800.0
    policy = WindowPolicy(nt, 2, "cyclical")
    ts = np.arange(0, 2 * nt * 10800.0, 5400.0)
    dem, pre = _walk(policy, times, ts, "cyclical")
    assert sum(pre) == 0 and dem[0] == 2
    assert all(d == (1 if k % 2 == 0 else 0) for k, d in enumerate(dem[1:], start=1))


def test_linear_and_clamped_series_stop_prefetching_at_the_end():
    nt = 6
    times = np.arange(nt) * 10800.0
    for indexing in ("linear", "clamp"):
        policy = WindowPolicy(nt, 3, indexing)
        ts = np.arange(600.0, (nt - 1) * 10800.0, 3600.0)
        dem, pre = _walk(policy, times, ts, indexing)
        assert dem[0] == 2 and sum(dem[1:]) == 0
        assert sum(pre) == nt - 2                      # every later slice exactly once, nothing past the last
    policy = WindowPolicy(nt, 3, "clamp")
    _walk(policy, times, [1e9], "clamp")               # clamped past the end: n1 == n2 == nt, one slot
    assert policy.where == {nt: policy.where[nt]}


def test_clock_jumps_reload_on_demand_and_recover():
    nt = 12
    times = np.arange(nt) * 10800.0
    rng = np.random.default_rng(3)
    policy = WindowPolicy(nt, 4, "cyclical")
    ts = list(rng.uniform(0, 5 * nt * 10800.0, 40))    # random access: every step may need both slices
    dem, _ = _walk(policy, times, ts, "cyclical")
    assert max(dem) <= 2
    t0 = ts[-1]
    dem, _ = _walk(policy, times, t0 + np.arange(1, 60) * 1800.0, "cyclical")   # marching again from wherever it landed
    assert sum(dem) == 0


def test_policy_invariants_under_arbitrary_clocks():
    """Any sequence of clock values, any window size, any indexing: the two interpolating slices are resident together
    after `demand`, a prefetch never displaces them, and the slot table stays consistent (checked inside _walk)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(nt=st.integers(2, 20), n_slots=st.integers(2, 8), indexing=st.sampled_from(["cyclical", "linear", "clamp"]),
           ts=st.lists(st.floats(-5e4, 4e5, allow_nan=False), min_size=1, max_size=40), lookahead=st.one_of(st.none(), st.integers(0, 9)))
    def run(nt, n_slots, indexing, ts, lookahead):
        times = np.arange(nt) * 10800.0
        policy = WindowPolicy(nt, min(n_slots, nt), indexing, lookahead)
        dem, pre = _walk(policy, times, ts, indexing)
        assert max(dem) <= 2
        assert all(p <= policy.lookahead for p in pre)
        assert len(policy.where) <= policy.n_slots

    run()


def test_upcoming_wraps_only_for_cyclical_series():
    assert WindowPolicy(5, 4, "cyclical").upcoming(4, 3) == [5, 1, 2]
    assert WindowPolicy(5, 4, "linear").upcoming(4, 3) == [5]
    assert WindowPolicy(2, 2, "cyclical").upcoming(2, 3) == [1]


# ------------------------------------------------------------------------------------------------- call order (CPU, fakes)
class _FakeDll:
    """Records the ring entry points SeriesWindow calls, in order."""

    def __init__(self):
        self.calls = []

    def ne_series_ring_create(self, handle_ref, desc_ref):
        self.calls.append(("create",))
        handle_ref._obj.value = 0xB200    # a non-null handle
        return 0

    def ne_series_ring_destroy(self, handle):
        self.calls.append(("destroy",))
        return 0

    def ne_series_ring_load(self, handle, slot, ptrs):
        self.calls.append(("load", slot.value, tuple(int(p) for p in ptrs)))
        return 0

    def ne_series_ring_acquire(self, handle, slot, stream):
        self.calls.append(("acquire", slot.value))
        return 0

    def ne_series_ring_release(self, handle, stream):
        self.calls.append(("release",))
        return 0


class _FakePinned:
    def __init__(self, a):
        self.a = a

    def pin_memory(self):
        return self

    def data_ptr(self):
        return self.a.ctypes.data


class _FakeTorch:
    @staticmethod
    def from_numpy(a):
        return _FakePinned(a)


class _FakeDeviceBackend(ne_b200.NumpyHostBackend):
    is_device = True
    torch = _FakeTorch()


class _FakeLib:
    def __init__(self):
        self.dll = _FakeDll()

    def last_error(self):
        return ""


def test_series_window_call_order_and_host_slices():
    """What SeriesWindow asks of the native ring, step by step: demand loads before acquire, acquire of both slots before the
    launch, release after it, then the prefetch loads — each load handing over the host pointers of THAT time index."""
    src = ne_b200.LatLonSourceGrid(nx=16, ny=8, FT="f32")
    nt = 6
    raw = {"a": np.arange(nt * 8 * 16, dtype=np.float32).reshape(nt, 8, 16), "b": np.ones((nt, 8, 16), np.float32)}
    lib = _FakeLib()
    w = SeriesWindow(_FakeDeviceBackend(), lib, src, np.arange(nt) * 100.0, raw, n_slots=4)
    calls = lib.dll.calls
    assert calls == [("create",)]
    assert w.series["a"].shape == (4, 8 + 2 * src.hy, 16 + 2 * src.hx)
    assert (w.desc.n_series, w.desc.n_slots, w.desc.raw_nx, w.desc.raw_ny) == (2, 4, 16, 8)
    slice_bytes = 8 * 16 * 4

    def host_ptrs(n):
        return tuple(w.host[k].data_ptr() + (n - 1) * slice_bytes for k in ("a", "b"))

    del calls[:]
    frac, m1, m2, same = w.time_interp(150.0, 0)           # n1 = 2, n2 = 3
    assert (frac, same) == (0.5, 0) and m1 != m2
    assert calls == [("load", m1 - 1, host_ptrs(2)), ("load", m2 - 1, host_ptrs(3)), ("acquire", m1 - 1), ("acquire", m2 - 1)] or \
        calls == [("load", m1 - 1, host_ptrs(2)), ("load", m2 - 1, host_ptrs(3)), ("acquire", m2 - 1), ("acquire", m1 - 1)]
    del calls[:]
    w.after_launch(0)
    assert calls[0] == ("release",)
    loads = calls[1:]
    assert [c[0] for c in loads] == ["load", "load"]         # lookahead = n_slots - 2: time indices 4 and 5
    assert [c[2] for c in loads] == [host_ptrs(4), host_ptrs(5)]
    assert {c[1] for c in loads}.isdisjoint({m1 - 1, m2 - 1})
    del calls[:]
    # same interval again: nothing to load, only the two waits and the release
    w.time_interp(180.0, 0)
    w.after_launch(0)
    assert sorted(c[0] for c in calls) == ["acquire", "acquire", "release"]
    del calls[:]
    # next interval (n1 = 3, n2 = 4): both resident; the prefetch replaces the slot of time index 2 with index 6
    slot_of_2 = w.policy.where[2]
    w.time_interp(250.0, 0)
    w.after_launch(0)
    assert [c for c in calls if c[0] == "load"] == [("load", slot_of_2, host_ptrs(6))]
    assert (w.demand_loads, w.prefetched) == (2, 3)
    w.close()
    assert calls[-1] == ("destroy",)


class _HostRing:
    """A window whose rings live in host memory and whose slot loads are the oracle's slot fill: the interface wiring
    (ring slots in NeTimeInterp, src_nt = n_slots, release after the launches) checked on the CPU against the oracle."""

    def __init__(self, oracle_lib, grid, times, raw, n_slots, indexing="cyclical"):
        self.lib, self.grid, self.times, self.raw, self.n_slots, self.time_indexing = oracle_lib, grid, times, raw, n_slots, indexing
        self.policy = WindowPolicy(len(times), n_slots, indexing)
        self.series = {k: np.full((n_slots,) + tuple(grid.shape), np.nan, dtype=NPD[grid.FT]) for k in raw}
        self.desc = _ring_desc(grid.FT, grid.nx, grid.ny, grid.hx, grid.hy)
        self.loads, self.releases, self._current = 0, 0, None

    def _load(self, n, slot):
        for k, v in self.raw.items():
            self.series[k][slot] = _oracle_fill(self.lib, self.desc, v[n - 1], self.grid.FT)
        self.loads += 1

    def time_interp(self, t, stream):
        frac, n1, n2 = ne_b200.interpolating_time_indices(self.times, t, self.time_indexing)
        for n, slot in self.policy.demand(n1, n2):
            self._load(n, slot)
        self._current = (n1, n2)
        return frac, self.policy.where[n1] + 1, self.policy.where[n2] + 1, int(n1 == n2)

    def after_launch(self, stream):
        self.releases += 1
        for n, slot in self.policy.prefetch(*self._current):
            self._load(n, slot)


@pytest.mark.parametrize("FT,atm_FT", [("f64", "f32"), ("f32", "f32")])
def test_interface_reads_ring_slots_like_the_in_memory_series_on_the_oracle(oracle_lib, host_backend, FT, atm_FT):
    cfg = dict(nx=48, ny=20, latitude=(-60.0, 60.0), src_nx=32, src_ny=16)
    nt = 6
    full = synthetic.build_case(cfg, host_backend, FT=FT, atm_FT=atm_FT, nt=nt, lib=oracle_lib)
    win = synthetic.build_case(cfg, host_backend, FT=FT, atm_FT=atm_FT, nt=nt, lib=oracle_lib)
    src = full.atmosphere.grid
    a = full._host_inputs["atmosphere"]
    raw = {k: np.ascontiguousarray(v[:, src.hy:src.hy + src.ny, src.hx:src.hx + src.nx]) for k, v in a.items()}
    padded = {k: np.stack([_numpy_fill(v[n], src.hx, src.hy, True, atm_FT) for n in range(nt)]) for k, v in raw.items()}
    ring = _HostRing(oracle_lib, src, full.atmosphere.times, raw, n_slots=3)
    for ci, s in ((full, padded), (win, ring.series)):
        atm, rad = ci.atmosphere, ci.radiation
        atm.u, atm.v, atm.T, atm.q, atm.p = s["u"], s["v"], s["T"], s["q"], s["p"]
        atm.rain, atm.snow = (s["rain"],), (s["snow"],)
        rad.downwelling_shortwave, rad.downwelling_longwave = s["sw"], s["lw"]
    win.atmosphere.window = win.radiation.window = ring      # one ring shared by both components
    full.initialize()
    win.initialize()
    steps = 0
    for k in range(2 * nt * 2 + 3):
        t = 50.0 + k * 10800.0 / 2
        full.update_state(t)
        win.update_state(t)
        steps += 1
        for bag_f, bag_w in ((full.atmos_state, win.atmos_state), (full.rad_state, win.rad_state), (full.ao_fluxes, win.ao_fluxes),
                             (full.net_ocean, win.net_ocean)):
            for name in bag_f.names():
                assert np.array_equal(getattr(bag_f, name), getattr(bag_w, name), equal_nan=True), (name, k)
    assert ring.releases == steps                            # once per step although two components share the ring
    crossed = int((50.0 + (steps - 1) * 10800.0 / 2) // 10800.0)
    assert ring.loads == 2 + 1 + crossed                     # two on demand, one look-ahead, then one prefetch per interval entered


# ------------------------------------------------------------------------------------------------- slot fill (oracle, CPU)
def _ring_desc(FT, nx, ny, hx, hy, n_series=1, periodic=True, conv=None, missing=None):
    d = A.NeSeriesRingDesc()
    d.n_series, d.n_slots, d.dtype, d.periodic_x = n_series, 2, A.NE_F64 if FT == "f64" else A.NE_F32, int(periodic)
    d.nx, d.ny, d.hx, d.hy = nx, ny, hx, hy
    for k in range(n_series):
        kind, a, b = CONVERSIONS[conv]
        d.conv_kind[k], d.conv_a[k], d.conv_b[k] = kind, a, b
        if missing is not None:
            d.has_missing[k], d.missing_value[k] = 1, missing
    return d


def _oracle_fill(oracle_lib, d, raw, FT):
    out = np.full((d.ny + 2 * d.hy, d.nx + 2 * d.hx), -777.0, dtype=NPD[FT])
    raw = np.ascontiguousarray(raw, dtype=NPD[FT])
    assert oracle_lib.dll.neo_series_slot_fill(C.addressof(d), 0, raw.ctypes.data, out.ctypes.data) == 0
    return out


def _numpy_fill(raw, hx, hy, periodic, FT, conv=None, missing=None):
    v = np.array(raw, dtype=NPD[FT])
    if missing is not None:
        v[v == NPD[FT](missing)] = np.nan
    kind, a, b = CONVERSIONS[conv]
    a, b = NPD[FT](a), NPD[FT](b)
    v = {A.NE_CONV_NONE: lambda: v, A.NE_CONV_NEGATE: lambda: -v, A.NE_CONV_ADD: lambda: v + a, A.NE_CONV_SUB: lambda: v - a,
         A.NE_CONV_MUL: lambda: v * a, A.NE_CONV_DIV: lambda: v / a, A.NE_CONV_MUL_DIV: lambda: (v * a) / b}[kind]()
    v = np.pad(v, ((0, 0), (hx, hx)), mode="wrap" if periodic else "symmetric")
    return np.pad(v, ((hy, hy), (0, 0)), mode="symmetric")


@pytest.mark.parametrize("FT", ["f32", "f64"])
@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("conv", [None, "Celsius", "Kelvin", "Millibar", "MillimetersPerHour", "MetersPerHour", "InverseSign"])
def test_oracle_slot_fill_is_set_region_plus_halo_fill(oracle_lib, FT, periodic, conv):
    rng = np.random.default_rng(11)
    nx, ny, hx, hy = 37, 19, 3, 2
    raw = rng.normal(280.0, 30.0, (ny, nx)).astype(NPD[FT])
    raw[rng.random((ny, nx)) < 0.05] = NPD[FT](-9999.0)
    d = _ring_desc(FT, nx, ny, hx, hy, periodic=periodic, conv=conv, missing=-9999.0)
    got = _oracle_fill(oracle_lib, d, raw, FT)
    want = _numpy_fill(raw, hx, hy, periodic, FT, conv, -9999.0)
    assert np.array_equal(got, want, equal_nan=True)
    assert np.isnan(got).any()
    if periodic:   # test/test_jra55.jl:40-47: fts[Nx+1, j] == fts[1, j], data[1, :] == data[Nx+1, :]
        assert np.array_equal(got[:, hx], got[:, hx + nx], equal_nan=True)
        assert np.array_equal(got[:, hx - 1], got[:, hx + nx - 1], equal_nan=True)


def test_mangle_known_answers_of_the_reference(oracle_lib):
    """test/test_mangling.jl:8-41: data = reshape(Float32[1 2 3; 4 5 6; 7 8 9], 3, 3, 1), data[i, j]; our raw[j, i]."""
    data = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9]], dtype=np.float32)    # data[i-1, j-1]
    raw = np.ascontiguousarray(data.T)

    def mangle(i, j, mangling):   # 1-based grid index (i, j) on a grid large enough to hold it; returns the interior value
        nx, ny = max(i, 3), max(j, 3)
        d = _ring_desc("f32", nx, ny, 0, 0)
        d.raw_nx, d.raw_ny = 3, 3
        d.mangling[0] = mangling
        out = np.zeros((ny, nx), dtype=np.float32)
        assert oracle_lib.dll.neo_series_slot_fill(C.addressof(d), 0, raw.ctypes.data, out.ctypes.data) == 0
        return out[j - 1, i - 1]

    N, S, AV = A.NE_MANGLE_NONE, A.NE_MANGLE_SHIFT_SOUTH, A.NE_MANGLE_AVERAGE_NORTH_SOUTH
    assert mangle(2, 2, N) == 5
    assert mangle(2, 2, S) == 4 and mangle(2, 1, S) == 4            # j - 1, clamped at the southern edge
    assert mangle(2, 1, AV) == np.float32(4.5) and mangle(2, 2, AV) == np.float32(5.5)
    assert mangle(4, 2, N) == data[2, 1] and mangle(2, 4, N) == data[1, 2]   # past the east / north edge: nearest cell
    assert mangle(2, 5, S) == data[1, 2]                             # j - 1 past north
    assert mangle(2, 3, AV) == data[1, 2]                            # j + 1 past north: (6 + 6) / 2


@pytest.mark.parametrize("FT", ["f32", "f64"])
def test_oracle_slot_fill_reads_a_bounding_box_region(oracle_lib, FT):
    """read_data(data, i, j, k, b::BoundingBoxOffset, …) = mangle(i + di, j + dj, …) (set_region_data.jl:163)."""
    rng = np.random.default_rng(2)
    file = rng.normal(0, 1, (50, 90)).astype(NPD[FT])
    nx, ny, di, dj = 20, 12, 31, 7
    d = _ring_desc(FT, nx, ny, 2, 2, periodic=False)
    d.raw_nx, d.raw_ny, d.di, d.dj = 90, 50, di, dj
    got = _oracle_fill(oracle_lib, d, file, FT)
    want = _numpy_fill(file[dj:dj + ny, di:di + nx], 2, 2, False, FT)
    assert np.array_equal(got, want)


def test_oracle_slot_fill_against_numpy_for_random_shapes_regions_and_manglings(oracle_lib):
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=120, deadline=None)
    @given(nx=st.integers(1, 24), ny=st.integers(1, 16), hx=st.integers(0, 3), hy=st.integers(0, 3), periodic=st.booleans(),
           mangling=st.sampled_from([A.NE_MANGLE_NONE, A.NE_MANGLE_SHIFT_SOUTH, A.NE_MANGLE_AVERAGE_NORTH_SOUTH]),
           di=st.integers(0, 5), dj=st.integers(0, 5), ex=st.integers(-1, 6), ey=st.integers(-1, 6), FT=st.sampled_from(["f32", "f64"]),
           seed=st.integers(0, 2**16))
    def run(nx, ny, hx, hy, periodic, mangling, di, dj, ex, ey, FT, seed):
        hx, hy = min(hx, nx), min(hy, ny)
        rnx, rny = max(1, nx + di + ex), max(1, ny + dj + ey)      # the file may be larger OR smaller than region + grid
        rng = np.random.default_rng(seed)
        file = rng.normal(0, 1, (rny, rnx)).astype(NPD[FT])
        d = _ring_desc(FT, nx, ny, hx, hy, periodic=periodic)
        d.raw_nx, d.raw_ny, d.di, d.dj = rnx, rny, di, dj
        d.mangling[0] = mangling
        got = _oracle_fill(oracle_lib, d, file, FT)
        # numpy restatement of read_data + mangle (set_region_data.jl:48-53, :163): clamped file indices
        ii = np.clip(np.arange(nx) + di, 0, rnx - 1)
        jj = np.arange(ny) + dj
        if mangling == A.NE_MANGLE_SHIFT_SOUTH:
            interior = file[np.clip(jj - 1, 0, rny - 1)][:, ii]
        elif mangling == A.NE_MANGLE_AVERAGE_NORTH_SOUTH:
            interior = (file[np.clip(jj, 0, rny - 1)][:, ii] + file[np.clip(jj + 1, 0, rny - 1)][:, ii]) / NPD[FT](2)
        else:
            interior = file[np.clip(jj, 0, rny - 1)][:, ii]
        want = _numpy_fill(interior, hx, hy, periodic, FT)
        assert np.array_equal(got, want)

    run()


# ------------------------------------------------------------------------------------------------- device

# ------------------------------------------------------------------------------------------------- Column regions (CPU)
def test_bracket_with_weight_known_answers_of_the_reference():
    """test/test_column_field.jl:18-66 (non-cyclic and cyclic wrap)."""
    coords = [0.5, 1.5, 2.5, 3.5]
    im, ip, w = bracket_with_weight(coords, 2.0)
    assert (im, ip) == (2, 3) and w == pytest.approx(0.5)
    assert bracket_with_weight(coords, -1.0) == (1, 2, 0.0)       # off-grid below: first interval
    assert bracket_with_weight(coords, 5.0) == (3, 4, 1.0)        # off-grid above: last interval
    im, ip, w = bracket_with_weight(coords, 3.5)
    assert (im, ip) == (3, 4) and w == pytest.approx(1.0)
    assert bracket_with_weight([7.5], 7.5) == (1, 1, 0.0)         # single-cell axis
    assert bracket_with_weight([7.5], 99.0) == (1, 1, 0.0)
    coords = np.arange(0.5, 360.0, 1.0)
    n = len(coords)
    im, ip, w = bracket_with_weight(coords, 180.0, period=360)
    assert (im, ip) == (180, 181) and w == pytest.approx(0.5)
    im, ip, w = bracket_with_weight(coords, 359.99, period=360)    # the wrap cell
    assert (im, ip) == (n, 1) and 0 < w < 1
    im, ip, w = bracket_with_weight(coords, 360.5, period=360)     # past the period: wrapped back
    assert im == 1 and ip == 2


def test_infer_longitudinal_period_known_answers_of_the_reference():
    """test/test_column_field.jl:68-73."""
    assert infer_longitudinal_period(np.arange(0.5, 360.0, 1.0)) == 360
    assert infer_longitudinal_period(np.arange(-179.75, 180.0, 0.5)) == 360
    assert infer_longitudinal_period([10.0, 11.0, 12.0]) is None
    assert infer_longitudinal_period([100.0]) is None


def _column_desc(FT, raw_nx, raw_ny, im, ip, jm, jp, wx, wy, kind="linear", hx=0, hy=0, missing=None, conv=None):
    d = _ring_desc(FT, 1, 1, hx, hy, conv=conv, missing=missing)
    d.raw_nx, d.raw_ny = raw_nx, raw_ny
    d.region_kind = A.NE_REGION_COLUMN
    d.column_interpolation = A.NE_COLUMN_NEAREST if kind == "nearest" else A.NE_COLUMN_LINEAR
    d.col_i_minus, d.col_i_plus, d.col_j_minus, d.col_j_plus = im - 1, ip - 1, jm - 1, jp - 1   # the ABI is 0-based
    d.col_wx, d.col_wy = float(NPD[FT](wx)), float(NPD[FT](wy))
    return d


def test_column_blend_known_answers_of_the_reference(oracle_lib):
    """test/test_column_field.jl:75-115: data = reshape(Float32[1 2; 3 4], 2, 2, 1), data[i, j]; our raw[j, i]."""
    full = np.array([[1, 2], [3, 4]], np.float32).T
    part = np.array([[1, 2], [3, np.nan]], np.float32).T
    allnan = np.full((2, 2), np.nan, np.float32)
    lin = _column_desc("f32", 2, 2, 1, 2, 1, 2, 0.5, 0.5)
    assert _oracle_fill(oracle_lib, lin, full, "f32")[0, 0] == np.float32(2.5)
    assert np.isnan(_oracle_fill(oracle_lib, lin, allnan, "f32")[0, 0])
    assert _oracle_fill(oracle_lib, lin, part, "f32")[0, 0] == pytest.approx(2.0, rel=1e-6)     # renormalised over three corners
    miss = part.copy()
    miss[np.isnan(miss)] = -9999.0                                                                # `missing` in the file
    assert _oracle_fill(oracle_lib, _column_desc("f32", 2, 2, 1, 2, 1, 2, 0.5, 0.5, missing=-9999.0), miss, "f32")[0, 0] == \
        pytest.approx(2.0, rel=1e-6)
    assert _oracle_fill(oracle_lib, _column_desc("f32", 2, 2, 1, 2, 1, 2, 0.7, 0.7, "nearest"), full, "f32")[0, 0] == 4.0
    assert _oracle_fill(oracle_lib, _column_desc("f32", 2, 2, 1, 2, 1, 2, 0.3, 0.3, "nearest"), full, "f32")[0, 0] == 1.0
    # the closest corner is land: Nearest falls back to the NaN-aware linear blend (set_region_data.jl:189-195)
    want = _oracle_fill(oracle_lib, _column_desc("f32", 2, 2, 1, 2, 1, 2, 0.7, 0.7), part, "f32")[0, 0]
    assert _oracle_fill(oracle_lib, _column_desc("f32", 2, 2, 1, 2, 1, 2, 0.7, 0.7, "nearest"), part, "f32")[0, 0] == want
    assert np.isfinite(want)


def _numpy_blend(raw, c, FT, kind, missing=None):
    """blend(::Linear / ::Nearest) written from the reference's docstring, one numpy scalar operation per rounding."""
    T = NPD[FT]
    v = np.array(raw, dtype=T)
    if missing is not None:
        v[v == T(missing)] = np.nan
    wx, wy = T(c.wx), T(c.wy)
    corners = [(c.i_minus, c.j_minus, (T(1) - wx) * (T(1) - wy)), (c.i_plus, c.j_minus, wx * (T(1) - wy)),
               (c.i_minus, c.j_plus, (T(1) - wx) * wy), (c.i_plus, c.j_plus, wx * wy)]
    if kind == "nearest":
        near = v[(c.j_plus if wy >= T(0.5) else c.j_minus) - 1, (c.i_plus if wx >= T(0.5) else c.i_minus) - 1]
        if not np.isnan(near):
            return near
    sw, num = None, None
    for i, j, w in corners:
        d = v[j - 1, i - 1]
        w = T(0) if np.isnan(d) else w
        t = w * (T(0) if np.isnan(d) else d)
        sw = w if sw is None else sw + w
        num = t if num is None else num + t
    return T(np.nan) if sw == 0 else num / sw


@pytest.mark.parametrize("FT", ["f32", "f64"])
@pytest.mark.parametrize("kind", ["linear", "nearest"])
def test_oracle_column_fill_against_numpy_for_random_points_and_coasts(oracle_lib, FT, kind):
    rng = np.random.default_rng(5)
    nx, ny = 48, 24
    lon, lat = (np.arange(nx) + 0.5) * 360.0 / nx, -90.0 + (np.arange(ny) + 0.5) * 180.0 / ny
    seen_nan = seen_wrap = 0
    for trial in range(200):
        raw = rng.normal(12.0, 9.0, (ny, nx)).astype(NPD[FT])
        raw[rng.random((ny, nx)) < 0.35] = NPD[FT](-9999.0)
        c = ColumnRegion(lon, lat, rng.uniform(-400.0, 400.0), rng.uniform(-95.0, 95.0), interpolation=kind)
        seen_wrap += (c.i_minus, c.i_plus) == (nx, 1)
        d = _column_desc(FT, nx, ny, c.i_minus, c.i_plus, c.j_minus, c.j_plus, c.wx, c.wy, kind, hx=2, hy=1, missing=-9999.0,
                         conv="Celsius")
        got = _oracle_fill(oracle_lib, d, raw, FT)
        b = _numpy_blend(raw, c, FT, kind, -9999.0)
        want = b + NPD[FT](273.15)
        assert np.array_equal(got, np.full_like(got, want), equal_nan=True), trial     # one value: interior and every halo cell
        seen_nan += bool(np.isnan(want))
    assert seen_nan > 0 and seen_wrap > 0


def _window_case(backend, lib, FT, atm_FT, nt, n_slots, seed_offset=0):
    """Two identically seeded interfaces: `full` reads the series fully in memory (halos filled the reference's way),
    `win` reads device rings fed from the raw slices."""
    cfg = dict(nx=96, ny=40, latitude=(-70.0, 70.0), src_nx=64, src_ny=32)
    full = synthetic.build_case(cfg, backend, FT=FT, atm_FT=atm_FT, nt=nt, seed_offset=seed_offset)
    win = synthetic.build_case(cfg, backend, FT=FT, atm_FT=atm_FT, nt=nt, seed_offset=seed_offset)
    src = full.atmosphere.grid
    a = full._host_inputs["atmosphere"]
    raw = {k: np.ascontiguousarray(v[:, src.hy:src.hy + src.ny, src.hx:src.hx + src.nx]) for k, v in a.items()}
    padded = {k: np.stack([_numpy_fill(v[n], src.hx, src.hy, True, atm_FT) for n in range(nt)]) for k, v in raw.items()}
    dev = {k: backend.from_numpy(v) for k, v in padded.items()}
    w = SeriesWindow(backend, lib, src, full.atmosphere.times, raw, n_slots=n_slots)
    for ci, s in ((full, dev), (win, w.series)):
        atm, rad = ci.atmosphere, ci.radiation
        atm.u, atm.v, atm.T, atm.q, atm.p = s["u"], s["v"], s["T"], s["q"], s["p"]
        atm.rain, atm.snow = (s["rain"],), (s["snow"],)
        rad.downwelling_shortwave, rad.downwelling_longwave = s["sw"], s["lw"]
    win.atmosphere.window = win.radiation.window = w
    full.initialize()
    win.initialize()
    return full, win, w, padded


@pytest.mark.gpu
@pytest.mark.parametrize("FT", ["f32", "f64"])
@pytest.mark.parametrize("periodic", [True, False])
def test_ring_slot_fill_kernel_matches_the_oracle(cuda_backend, cuda_lib, oracle_lib, FT, periodic):
    import torch
    rng = np.random.default_rng(5)
    src = ne_b200.LatLonSourceGrid(nx=640, ny=320, FT=FT)
    names = ["T", "p", "rain"]
    convs = {"T": "Celsius", "p": "Millibar", "rain": "MillimetersPerHour"}
    nt = 3
    raw = {k: rng.normal(20.0, 9.0, (nt, src.ny, src.nx)).astype(NPD[FT]) for k in names}
    raw["rain"][raw["rain"] < 8.0] = NPD[FT](1e20)
    w = SeriesWindow(cuda_backend, cuda_lib, src, np.arange(nt) * 10800.0, raw, n_slots=3, conversions=convs,
                     missing_values={"rain": 1e20}, periodic_x=periodic)
    s = cuda_backend.stream()
    w.time_interp(0.5 * 10800.0, s)      # demand-loads slices 1 and 2
    w.after_launch(s)                    # prefetches slice 3
    torch.cuda.synchronize()
    assert w.demand_loads == 2 and w.prefetched == 1
    for k in names:
        d = _ring_desc(FT, src.nx, src.ny, src.hx, src.hy, periodic=periodic, conv=convs[k], missing=1e20 if k == "rain" else None)
        ring = cuda_backend.to_numpy(w[k])
        for n in range(1, nt + 1):
            want = _oracle_fill(oracle_lib, d, raw[k][n - 1], FT)
            assert np.array_equal(ring[w.policy.where[n]], want, equal_nan=True), (k, n)
    assert np.isnan(cuda_backend.to_numpy(w["rain"])).any()
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize("FT,atm_FT,n_slots", [("f64", "f32", 4), ("f64", "f64", 3), ("f32", "f32", 5), ("f64", "f32", 2)])
def test_windowed_series_interpolate_like_the_series_fully_in_memory(cuda_backend, cuda_lib, FT, atm_FT, n_slots):
    import torch
    nt = 8
    full, win, w, _ = _window_case(cuda_backend, cuda_lib, FT, atm_FT, nt, n_slots)
    dt = 10800.0 / 5
    steps = int(2.3 * nt * 5)            # more than twice around the cyclical series
    for k in range(steps):
        t = 400.0 + k * dt
        full.interpolate_state(t)
        win.interpolate_state(t)
        if k % 7 == 0 or k == steps - 1:
            torch.cuda.synchronize()
            for name in ("u", "v", "T", "q", "p", "Jrn", "Jsn"):
                x, y = cuda_backend.to_numpy(getattr(full.atmos_state, name)), cuda_backend.to_numpy(getattr(win.atmos_state, name))
                assert np.array_equal(x, y, equal_nan=True), (name, k)
            for name in ("sw", "lw"):
                x, y = cuda_backend.to_numpy(getattr(full.rad_state, name)), cuda_backend.to_numpy(getattr(win.rad_state, name))
                assert np.array_equal(x, y, equal_nan=True), (name, k)
    torch.cuda.synchronize()
    n_intervals = int((400.0 + (steps - 1) * dt) // 10800.0)
    if n_slots >= 3:   # after the first step nothing was ever waited for; each later slice arrived ahead of its interval
        assert w.demand_loads == 2
        assert w.prefetched >= n_intervals
    else:
        assert w.prefetched == 0 and w.demand_loads == 2 + n_intervals
    w.close()


@pytest.mark.gpu
def test_windowed_fused_step_and_host_pipeline_match_the_in_memory_step(cuda_backend, cuda_lib):
    """The one-call interface step and the host-buffer pipeline read the rings too (ring slots in NeTimeInterp.m1/m2,
    src_nt = n_slots), including the merged 9-series interpolation launch."""
    import torch
    nt = 6
    full, win, w, _ = _window_case(cuda_backend, cuda_lib, "f64", "f32", nt, 4, seed_offset=2)
    assert win.shared_frac
    pipe = ne_b200.HostPipelinedStep(win, n_chunks=3)
    host = {k: torch.from_numpy(cuda_backend.to_numpy(getattr(win.ocean_state, k))).pin_memory() for k in ("T", "S", "u", "v")}
    for k in range(3 * nt):
        t = 100.0 + k * 10800.0 * 0.75
        full.fused_interface_step(t)
        if k % 2:
            pipe.step(t, host)
        else:
            win.fused_interface_step(t)
        torch.cuda.synchronize()
        for name in ("latent_heat", "sensible_heat", "water_vapor", "x_momentum", "y_momentum"):
            x, y = cuda_backend.to_numpy(getattr(full.ao_fluxes, name)), cuda_backend.to_numpy(getattr(win.ao_fluxes, name))
            assert np.array_equal(x, y, equal_nan=True), (name, k)
        for name in ("T", "S", "u", "v"):
            x, y = cuda_backend.to_numpy(getattr(full.net_ocean, name)), cuda_backend.to_numpy(getattr(win.net_ocean, name))
            assert np.array_equal(x, y, equal_nan=True), (name, k)
    assert w.demand_loads == 2
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize("FT", ["f32", "f64"])
@pytest.mark.parametrize("case", ["shift_south", "average_north_south", "bounding_box"])
def test_ring_slot_fill_with_mangling_and_regions_matches_the_oracle(cuda_backend, cuda_lib, oracle_lib, FT, case):
    import torch
    rng = np.random.default_rng(8)
    src = ne_b200.LatLonSourceGrid(nx=72, ny=36, FT=FT)
    raw_ny = {"shift_south": 35, "average_north_south": 37, "bounding_box": 60}[case]
    raw_nx = 120 if case == "bounding_box" else 72
    off = (17, 9) if case == "bounding_box" else (0, 0)
    raw = {"v": rng.normal(0, 3, (2, raw_ny, raw_nx)).astype(NPD[FT])}
    raw["v"][rng.random(raw["v"].shape) < 0.03] = NPD[FT](-9999.0)
    w = SeriesWindow(cuda_backend, cuda_lib, src, np.arange(2) * 3600.0, raw, n_slots=2, conversions={"v": "CentimetersPerSecond"},
                     missing_values={"v": -9999.0}, periodic_x=case != "bounding_box", region_offset=off)
    assert w.mangling == {"shift_south": A.NE_MANGLE_SHIFT_SOUTH, "average_north_south": A.NE_MANGLE_AVERAGE_NORTH_SOUTH,
                          "bounding_box": A.NE_MANGLE_NONE}[case]
    w.time_interp(1800.0, cuda_backend.stream())
    torch.cuda.synchronize()
    ring = cuda_backend.to_numpy(w["v"])
    for n in (1, 2):
        want = _oracle_fill(oracle_lib, w.desc, raw["v"][n - 1], FT)
        assert np.array_equal(ring[w.policy.where[n]], want, equal_nan=True), n
        assert np.isnan(want).any() and np.isfinite(want).any()
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize("FT", ["f32", "f64"])
@pytest.mark.parametrize("kind", ["linear", "nearest"])
def test_ring_slot_fill_of_a_column_region_matches_the_oracle(cuda_backend, cuda_lib, oracle_lib, FT, kind):
    """Column(lon, lat; interpolation): a 1 x 1 series whose value is the NaN-aware blend of four file cells
    (set_region_data.jl:113-118, :164-195), through the ring like any other region."""
    import torch
    rng = np.random.default_rng(21)
    nx, ny, nt = 90, 45, 3
    lon, lat = (np.arange(nx) + 0.5) * 4.0, -90.0 + (np.arange(ny) + 0.5) * 4.0
    seen_nan = seen_finite = 0
    for point in [(12.0, -50.0), (359.3, 10.0), (-61.5, 18.0), (181.0, 89.9), (77.7, -33.3), (3.0, 0.2)]:
        raw = {"T": rng.normal(10, 8, (nt, ny, nx)).astype(NPD[FT]), "S": rng.normal(35, 1, (nt, ny, nx)).astype(NPD[FT])}
        for v in raw.values():
            v[rng.random(v.shape) < 0.4] = NPD[FT](-9999.0)
        col = ColumnRegion(lon, lat, *point, interpolation=kind)
        src = ne_b200.LatLonSourceGrid(nx=1, ny=1, hx=1, hy=1, FT=FT, lam_nodes=[point[0]], phi_nodes=[point[1]])
        w = SeriesWindow(cuda_backend, cuda_lib, src, np.arange(nt) * 3600.0, raw, n_slots=2, conversions={"T": "Celsius"},
                         missing_values={"T": -9999.0, "S": -9999.0}, column=col)
        w.time_interp(1800.0, cuda_backend.stream())
        torch.cuda.synchronize()
        for k, name in enumerate(("T", "S")):
            ring = cuda_backend.to_numpy(w[name])
            for n in (1, 2):
                want = np.full((3, 3), -777.0, NPD[FT])
                assert oracle_lib.dll.neo_series_slot_fill(C.addressof(w.desc), k, np.ascontiguousarray(raw[name][n - 1]).ctypes.data,
                                                           want.ctypes.data) == 0
                assert np.array_equal(ring[w.policy.where[n]], want, equal_nan=True), (point, name, n)
                seen_nan += bool(np.isnan(want).any())
                seen_finite += bool(np.isfinite(want).all())
        w.close()
    assert seen_finite > 0 and seen_nan > 0


@pytest.mark.gpu
def test_windowed_land_runoff_matches_the_in_memory_series(cuda_backend, cuda_lib):
    """PrescribedLand (daily JRA55 river + iceberg runoff, Lands/interpolate_land_state.jl:6-61) behind its own ring."""
    import torch
    cfg = dict(nx=96, ny=40, latitude=(-70.0, 70.0), src_nx=64, src_ny=32, land_nx=120, land_ny=60)
    nt = 7
    full = synthetic.build_case(cfg, cuda_backend, FT="f64", atm_FT="f32", nt=nt, land=True)
    win = synthetic.build_case(cfg, cuda_backend, FT="f64", atm_FT="f32", nt=nt, land=True)
    lsrc = full.land.grid
    series = [cuda_backend.to_numpy(x) for x in full.land.freshwater_flux]
    raw = {f"runoff{k}": np.ascontiguousarray(v[:, lsrc.hy:lsrc.hy + lsrc.ny, lsrc.hx:lsrc.hx + lsrc.nx]) for k, v in enumerate(series)}
    padded = [np.stack([_numpy_fill(v[n], lsrc.hx, lsrc.hy, True, "f32") for n in range(nt)]) for v in raw.values()]
    full.land.freshwater_flux = tuple(cuda_backend.from_numpy(v) for v in padded)
    w = SeriesWindow(cuda_backend, cuda_lib, lsrc, full.land.times, raw, n_slots=3)
    win.land.freshwater_flux = tuple(w[k] for k in raw)
    win.land.window = w
    full.initialize()
    win.initialize()
    for k in range(40):
        t = 1000.0 + k * 86400.0 * 0.45
        full.interpolate_state(t)
        win.interpolate_state(t)
        torch.cuda.synchronize()
        x, y = cuda_backend.to_numpy(full.land_state.freshwater_flux), cuda_backend.to_numpy(win.land_state.freshwater_flux)
        assert np.array_equal(x, y, equal_nan=True), k
        assert (x > 0).any()
    assert w.demand_loads == 2 and w.prefetched >= 15
    w.close()


@pytest.mark.gpu
def test_ring_rejects_bad_descriptors(cuda_backend, cuda_lib):
    d = _ring_desc("f32", 8, 8, 3, 3)
    h = C.c_void_p()
    assert cuda_lib.dll.ne_series_ring_create(C.byref(h), C.byref(d)) == A.NE_E_INVALID   # null ring pointer
    assert "ring" in cuda_lib.last_error()
    d.n_slots = 1
    assert cuda_lib.dll.ne_series_ring_create(C.byref(h), C.byref(d)) == A.NE_E_INVALID
    with pytest.raises(RuntimeError):   # acquire of a slot nothing was ever loaded into
        src = ne_b200.LatLonSourceGrid(nx=16, ny=8, FT="f32")
        w = SeriesWindow(cuda_backend, cuda_lib, src, np.arange(3) * 1.0, {"a": np.zeros((3, 8, 16), np.float32)}, n_slots=3)
        w._check(cuda_lib.dll.ne_series_ring_acquire(w.handle, C.c_int32(1), C.c_void_p(cuda_backend.stream())))
