"""Latitude-band sharding of the exchange grid over the GPUs of one box.

Every kernel on the path is pointwise apart from 2-point stencils, and — exactly like the reference,
which computes fluxes on (0:N+1) so that no halo exchange is needed afterwards
(EarthSystemModels/InterfaceComputations/InterfaceComputations.jl:108-112) — each band overcomputes
one ring, so there is NO data-path collective.  The only collective is the global flux /
conservation diagnostics sum (src/Diagnostics/interface_fluxes.jl:90-195): per-band partial sums on
the device (ne_diag_reduce, fixed summation order) followed by one small all-reduce (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  The reference itself issues no collective on this
path: each MPI rank holds the whole atmosphere
(DataWrangling/JRA55/JRA55_prescribed_atmosphere.jl:7-8) — as does each band here.
"""
import numpy as np

from . import abi as A
from .interface import ExchangeGrid


def latitude_bands(ny, world_size, weights=None):
    """Split rows 1..ny into `world_size` contiguous bands [(j0, j1)] (1-based, inclusive).
    With `weights` (per-row expected cost: active points x expected iterations) the split balances
    cumulative weight instead of row count."""
    if world_size < 1 or ny < world_size:
        raise ValueError("need at least one row per band")
    if weights is None:
        edges = [round(r * ny / world_size) for r in range(world_size + 1)]
    else:
        w = np.asarray(weights, dtype=np.float64)
        if w.shape != (ny,):
            raise ValueError("weights must have one entry per row")
        c = np.concatenate([[0.0], np.cumsum(w)])
        edges = [0]
        for r in range(1, world_size):
            e = int(np.searchsorted(c, c[-1] * r / world_size))
            e = min(max(e, edges[-1] + 1), ny - (world_size - r))
            edges.append(e)
        edges.append(ny)
    return [(edges[r] + 1, edges[r + 1]) for r in range(world_size)]


def measured_row_weights(nx, warp_trips_per_row, per_point=2.6):
    """Per-row cost from MEASURED trip counts (the `iterations` output of the previous coupled step: the trip count
    of a point changes slowly from step to step, so last step's field predicts this step's cost).  The solve costs
    one unit per lane SLOT of a warp-trip — a warp runs until its slowest lane converges, so the row's cost is
    32 x the sum over its 32-point groups of the group's maximum trip count (fits the measured per-band kernel times
    to 4 %; the plain sum of trips to 8.5 %) — plus `per_point` units for the HBM-bound kernels every point pays
    (ratio measured on B200, profiles/r01_notes.md)."""
    return np.asarray(warp_trips_per_row, dtype=np.float64) + per_point * float(nx)


def ordered_row_weights(nx, active_per_row, trips_per_row, per_active=1.5, per_point=3.4):
    """Per-row cost for the TRIP-ORDERED solve (ne_flux_tab2.cu): the lanes of a warp leave the loop together (lane
    efficiency 97 %), so a row costs the plain sum of its trips, plus `per_active` trip-equivalents per solved point for
    the prologue / epilogue (~600 of ~400 instructions per trip) and `per_point` for the HBM-bound kernels every launch
    point pays (C4 on B200: solve 1.28 ms for 69.1 M thread-trips = 18.5 ps per trip, interpolation + post-solve 0.46 ms
    for 7.27 M points = 63 ps per point)."""
    return (np.asarray(trips_per_row, dtype=np.float64) + per_active * np.asarray(active_per_row, dtype=np.float64) +
            per_point * float(nx))


def gather_row_statistics(iterations_interior, grid, group=None):
    """All ranks: per-row (active points, trips, warp-slot trips) of the GLOBAL grid from each band's
    iteration-count field (rows 1..ny of the band's launch window; one small all_gather_object)."""
    import torch.distributed as dist
    it = np.asarray(iterations_interior)[1:-1, 1:-1]           # drop the overcomputed ring
    n32 = it.shape[1] // 32 * 32
    groups = it[:, :n32].reshape(it.shape[0], -1, 32).max(axis=2).sum(axis=1) * 32
    if n32 < it.shape[1]:
        groups = groups + it[:, n32:].max(axis=1) * 32
    mine = (int(grid.j_offset), (it > 0).sum(axis=1).astype(np.int64), it.sum(axis=1).astype(np.int64), groups.astype(np.int64))
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, mine, group=group)
    ny = grid.ny_global or grid.ny
    active, trips, wtrips = (np.zeros(ny, dtype=np.int64) for _ in range(3))
    for j0, a, t, w in parts:
        active[j0:j0 + len(a)] = a
        trips[j0:j0 + len(t)] = t
        wtrips[j0:j0 + len(w)] = w
    return active, trips, wtrips


def band_grid(nx, ny, latitude, rank, world_size, FT="f64", hx=7, hy=7, weights=None):
    j0, j1 = latitude_bands(ny, world_size, weights)[rank]
    return ExchangeGrid(nx=nx, ny=j1 - j0 + 1, hx=hx, hy=hy, latitude=latitude, FT=FT, j_offset=j0 - 1, ny_global=ny)


def band_rows(global_parent, grid: ExchangeGrid):
    """Rows of a GLOBAL exchange-layout parent array (ny_global+2hy, nx+2hx) that make up the band's
    parent (its interior + hy halo rows each side, i.e. the neighbours' data: the one-row ocean halo
    the overcomputed ring reads)."""
    return global_parent[..., grid.j_offset:grid.j_offset + grid.ny + 2 * grid.hy, :]


class FluxDiagnostics:
    """Area-weighted global integrals of flux fields: local deterministic two-stage sum on the device
    + one all-reduce of n_fields doubles."""

    def __init__(self, interfaces, fields, n_blocks=1184):   # 8 resident 256-thread blocks per SM on a 148-SM B200
        self.ci = interfaces
        b, g = interfaces.backend, interfaces.grid
        self.fields = fields
        phi = np.deg2rad(g.phi.astype(np.float64))
        area = (np.cos(phi)[:, None] * np.ones((1, g.shape[1]))).astype(np.float64 if g.FT == "f64" else np.float32)
        self.area = b.from_numpy(area)
        self.partial = b.zeros((n_blocks * len(fields),), "f64")
        # two result vectors: with an asynchronous all-reduce the next step's local sums go to the other one
        self.results = [b.zeros((len(fields),), "f64"), b.zeros((len(fields),), "f64")]
        self._cur = 0
        self._work = [None, None]
        d = A.NeDiagDesc()
        d.grid = g.pod(False)
        d.n_fields = len(fields)
        for k, f in enumerate(fields):
            d.fields[k] = b.ptr(f)
        d.area = b.ptr(self.area)
        d.inactive = b.ptr(interfaces.inactive) if interfaces.inactive is not None else None
        d.partial, d.n_blocks, d.result = b.ptr(self.partial), n_blocks, b.ptr(self.results[0])
        self.desc = d

    @property
    def result(self):
        """The result vector of the current step (global sums once its all-reduce has been waited for)."""
        return self.results[self._cur]

    def flip(self, *descs):
        """Switch to the other result vector before enqueuing the next step, so that an all-reduce still in flight
        (all_reduce(async_op=True)) is not overwritten.  `descs`: cached NeFusedStepDesc / NeHostStepDesc / NeDiagDesc
        copies of this object's descriptor whose `result` pointer has to follow."""
        self._cur ^= 1
        w = self._work[self._cur]
        if w is not None:          # the all-reduce issued two steps ago into this vector: long done, orders the streams
            w.wait()
            self._work[self._cur] = None
        ptr = self.ci.backend.ptr(self.results[self._cur])
        self.desc.result = ptr
        for d in descs:
            if hasattr(d, "step"):
                d.step.diag.result = ptr
            elif hasattr(d, "diag"):
                d.diag.result = ptr
            else:
                d.result = ptr

    def wait(self):
        """Make the current stream (or the host, gloo) wait for every all-reduce still in flight."""
        for k, w in enumerate(self._work):
            if w is not None:
                w.wait()
                self._work[k] = None

    def reduce(self, group=None):
        """Enqueue the local reduction and (world_size > 1) the all-reduce; returns the device/host
        result array without synchronising."""
        self.ci.lib.call("diag_reduce", self.ci.grid.FT, self.desc, self.ci.backend.stream())
        return self.all_reduce(group)

    def all_reduce(self, group=None, async_op=False):
        """The collective alone: for local sums already produced by the fused interface step
        (NeFusedStepDesc.diag).  async_op: the collective runs on the communicator's own stream behind the kernels
        enqueued so far and the current stream does NOT wait for it — the next step's kernels overlap it; call
        flip() before enqueuing that step and wait() before reading `result`."""
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                r = self.result
                if isinstance(r, np.ndarray):
                    import torch
                    r = torch.from_numpy(r)
                if async_op:
                    self.wait_current()
                    self._work[self._cur] = dist.all_reduce(r, group=group, async_op=True)
                else:
                    dist.all_reduce(r, group=group)
        except ImportError:
            pass
        return self.result

    def wait_current(self):
        w = self._work[self._cur]
        if w is not None:
            w.wait()
            self._work[self._cur] = None
