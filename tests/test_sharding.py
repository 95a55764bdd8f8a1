"""Latitude-band sharding: band bookkeeping, band-vs-global equivalence on the oracle,
and the diagnostics all-reduce over a 2-rank gloo group on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ne_b200
from numericalearth_jl_b200 import sharding, synthetic


def test_latitude_bands_cover_rows_exactly():
    for ny, w in ((150, 1), (150, 2), (560, 8), (1680, 8), (7, 7)):
        bands = sharding.latitude_bands(ny, w)
        assert bands[0][0] == 1 and bands[-1][1] == ny
        for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
            assert b0 == a1 + 1 and a1 >= a0
    with pytest.raises(ValueError):
        sharding.latitude_bands(3, 4)
    w = np.ones(100); w[:10] = 10.0                      # costly polar rows -> thinner first band
    bands = sharding.latitude_bands(100, 4, weights=w)
    assert bands[0][1] - bands[0][0] + 1 < 25 and bands[-1][1] == 100


def _band_run(oracle_lib, rank, world):
    host = ne_b200.NumpyHostBackend()
    cfg = synthetic.CONFIGS["tiny"]
    grid = sharding.band_grid(cfg["nx"], cfg["ny"], cfg["latitude"], rank, world)
    ci = synthetic.build_case("tiny", host, lib=oracle_lib, grid=grid)
    ci.initialize()
    ci.update_state(4000.0)
    return ci


def test_bands_reproduce_the_global_step_bit_for_bit(oracle_lib, host_backend):
    """Pointwise kernels + one-ring overcompute: every band's interior equals the global result exactly,
    with no halo exchange of fluxes (InterfaceComputations.jl:108-112)."""
    glob = synthetic.build_case("tiny", host_backend, lib=oracle_lib)
    glob.initialize(); glob.update_state(4000.0)
    g = glob.grid
    for world in (2, 3):
        for rank in range(world):
            ci = _band_run(oracle_lib, rank, world)
            b = ci.grid
            rows = slice(b.hy, b.hy + b.ny)
            grows = slice(g.hy + b.j_offset, g.hy + b.j_offset + b.ny)
            for bag_b, bag_g in ((ci.ao_fluxes, glob.ao_fluxes), (ci.net_ocean, glob.net_ocean), (ci.atmos_state, glob.atmos_state)):
                for n in bag_b.names():
                    assert np.array_equal(getattr(bag_b, n)[rows, b.hx:b.hx + b.nx], getattr(bag_g, n)[grows, g.hx:g.hx + g.nx]), (world, rank, n)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    lib = oracle.load()
    ci = _band_run(lib, rank, world)
    f = ci.ao_fluxes
    diag = sharding.FluxDiagnostics(ci, [f.latent_heat, f.sensible_heat, f.x_momentum, ci.net_ocean.T], n_blocks=4)
    # the oracle has no diag entry point: local partial sums on the host, then the same all-reduce
    g = ci.grid
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    act = ci.inactive[rows, cols] == 0
    local = np.array([(np.asarray(x)[rows, cols] * diag.area[rows, cols])[act].sum() for x in diag.fields])
    t = torch.from_numpy(local.copy())
    dist.all_reduce(t)
    # the product's collective, asynchronous and double-buffered the way bench.py steps: three "steps" whose local sums
    # are k * local; each result is read only after wait(), while the next step's sums already sit in the other vector
    seen = []
    for k in (1, 2, 3):
        diag.flip()
        assert diag.desc.result == ci.backend.ptr(diag.result)
        diag.result[:] = k * local
        diag.all_reduce(async_op=True)
        if k > 1:
            prev = diag.results[diag._cur ^ 1]
            w = diag._work[diag._cur ^ 1]
            if w is not None:
                w.wait()
            seen.append(prev.copy())
    diag.wait()
    seen.append(diag.result.copy())
    q.put((rank, t.numpy().copy(), local, seen))
    dist.destroy_process_group()


def test_diagnostics_all_reduce_two_ranks_gloo(oracle_lib, host_backend):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda x: x[0])
    assert np.array_equal(res[0][1], res[1][1])                       # every rank holds the same global sums
    assert np.allclose(res[0][1], res[0][2] + res[1][2], rtol=1e-15)
    for r in res:                                                     # the asynchronous, double-buffered path: step k sums to k x global
        assert len(r[3]) == 3
        for k, got in enumerate(r[3], start=1):
            assert np.allclose(got, k * (res[0][2] + res[1][2]), rtol=1e-15), k
    # and they equal the single-rank (global grid) integrals
    glob = synthetic.build_case("tiny", host_backend, lib=oracle_lib)
    glob.initialize(); glob.update_state(4000.0)
    g = glob.grid
    rows, cols = slice(g.hy, g.hy + g.ny), slice(g.hx, g.hx + g.nx)
    area = np.cos(np.deg2rad(g.phi.astype(np.float64)))[:, None] * np.ones((1, g.shape[1]))
    act = glob.inactive[rows, cols] == 0
    f = glob.ao_fluxes
    ref = np.array([(x[rows, cols] * area[rows, cols])[act].sum() for x in (f.latent_heat, f.sensible_heat, f.x_momentum, glob.net_ocean.T)])
    assert np.allclose(res[0][1], ref, rtol=1e-12)


def _rebalance_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    lib = oracle.load()
    host = ne_b200.NumpyHostBackend()
    cfg = synthetic.CONFIGS["tiny"]
    grid = sharding.band_grid(cfg["nx"], cfg["ny"], cfg["latitude"], rank, world)
    ci = synthetic.build_case("tiny", host, lib=lib, grid=grid, with_iterations=True)
    ci.initialize(); ci.update_state(4000.0)
    active, trips, wtrips = sharding.gather_row_statistics(grid.interior(ci.ao_iterations), grid)
    q.put((rank, active, trips, wtrips))
    dist.destroy_process_group()


def test_rebalancing_from_measured_trip_counts_two_ranks_gloo(oracle_lib, host_backend):
    """Every rank assembles the same global per-row (active, trips) table from the bands' iteration-count fields; the
    table equals the single-rank one; bands dealt by it are contiguous, cover every row once and even out the cost."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rebalance_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    glob = synthetic.build_case("tiny", host_backend, lib=oracle_lib, with_iterations=True)
    glob.initialize(); glob.update_state(4000.0)
    g = glob.grid
    it = g.interior(glob.ao_iterations)[1:-1, 1:-1]
    assert np.array_equal(res[0][1], (it > 0).sum(axis=1)) and np.array_equal(res[0][2], it.sum(axis=1))
    assert np.array_equal(res[0][3], res[1][3]) and (res[0][3] >= res[0][2]).all()   # a warp waits for its slowest lane
    w = sharding.measured_row_weights(g.nx, res[0][3])
    assert (w > 0).all()
    for world in (2, 3, 4):
        bands = sharding.latitude_bands(g.ny, world, w)
        assert bands[0][0] == 1 and bands[-1][1] == g.ny
        assert all(bands[k][1] + 1 == bands[k + 1][0] for k in range(world - 1))
        cost = np.array([w[j0 - 1:j1].sum() for j0, j1 in bands])
        equal = np.array([w[j0 - 1:j1].sum() for j0, j1 in sharding.latitude_bands(g.ny, world)])
        assert cost.max() <= equal.max() + w.max()      # never worse than equal rows by more than one row
