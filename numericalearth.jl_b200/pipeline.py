"""Host-buffer interface step: the ocean surface state (T, S, u, v) arrives in pinned HOST memory
every coupled step (an ocean model that lives on the host, or on another device behind a host
bounce buffer), the interface fluxes are wanted back as global diagnostics on the host.

The reference does this serially (copy, then `update_state!`).  Here the exchange grid is cut into
latitude chunks: chunk c's rows of the four fields are copied host->device on a COPY stream
(`ne_memcpy_h2d`) while the COMPUTE stream runs the a–o solve, the net-flux assembly and the
radiation kernel on the launch rows whose inputs (including the (j+1) row of the v stencil,
atmosphere_ocean_fluxes.jl:64-65) have already landed — the descriptors' launch range
(`NeExchangeGrid.j_lo/j_hi`) restricts each call to its rows.  The atmosphere interpolation does not
depend on the ocean and runs up front, hidden behind the first copy.  PCIe time and kernel time
overlap instead of adding up; results are identical to the unchunked step (pointwise kernels, and
the (i-1, j-1) stencils of the assembly only reach back into rows an earlier chunk computed).
"""
import ctypes as C

import numpy as np

from . import abi as A
from .interface import interpolating_time_indices


class HostPipelinedStep:
    FIELDS = ("T", "S", "u", "v")

    def __init__(self, interfaces, n_chunks=8, diagnostics=None):
        ci = self.ci = interfaces
        if not ci.backend.is_device:
            raise RuntimeError("HostPipelinedStep needs the CUDA library and device arrays (there is no CPU fallback)")
        import torch
        self.torch = torch
        g = ci.grid
        self.lib = ci.lib
        self.FT = g.FT
        self.diagnostics = diagnostics
        self.copy_stream = torch.cuda.Stream(device=ci.backend.device)
        self.dev = {"T": ci.ocean_state.T, "S": ci.ocean_state.S, "u": ci.ocean_state.u, "v": ci.ocean_state.v}
        rows_total = g.ny + 2 * g.hy
        self.row_bytes = (g.nx + 2 * g.hx) * (8 if g.FT == "f64" else 4)
        n_chunks = max(1, min(int(n_chunks), g.ny))
        # copy chunks: a partition of the parent rows; compute bands: launch rows whose j and j+1 parent rows have landed
        edges = np.linspace(0, rows_total, n_chunks + 1).round().astype(int)
        self.copy_rows = [(int(edges[c]), int(edges[c + 1])) for c in range(n_chunks)]
        self.bands = []
        j_next = 0
        for c, (r0, r1) in enumerate(self.copy_rows):
            b = (r1 - 1) - g.hy          # last launch row whose (j+1) parent row (j + hy) is inside [0, r1)
            if c == n_chunks - 1:
                b = g.ny + 1
            b = min(b, g.ny + 1)
            self.bands.append((j_next, b) if b >= j_next else None)
            j_next = max(j_next, b + 1)
        self.events = [torch.cuda.Event() for _ in range(n_chunks)]
        self.done = torch.cuda.Event()
        # descriptors: interpolation over the whole grid, then one (ao, assemble, apply) triple per band
        self.fused = ci.fused_step_desc(0.0)
        self.band_descs = []
        for band in self.bands:
            if band is None:
                self.band_descs.append(None)
                continue
            j0, j1 = band
            ao = ci.atmosphere_ocean_desc()
            ao.grid.j_lo, ao.grid.j_hi = j0, j1
            asm = ci.assemble_ocean_desc()
            a0, a1 = max(j0, 1), min(j1, g.ny)
            asm.grid.j_lo, asm.grid.j_hi = a0, a1
            rad = None
            if ci.radiation is not None:
                rad = ci.apply_radiation_desc(False)
                rad.grid.j_lo, rad.grid.j_hi = a0, a1
            self.band_descs.append((ao, asm if a1 >= a0 else None, rad if a1 >= a0 else None))

    def _set_time(self, t):
        ci = self.ci
        for d, src in ((self.fused.atmosphere, ci.atmosphere), (self.fused.radiation, ci.radiation)):
            if src is None:
                continue
            nt, n1, n2 = interpolating_time_indices(src.times, t, src.time_indexing)
            d.time.frac, d.time.m1, d.time.m2, d.time.same = nt, n1, n2, int(n1 == n2)

    def step(self, t, host_ocean):
        """host_ocean: dict of pinned host arrays (torch CPU tensors or numpy) named T, S, u, v with the exchange
        layout.  Enqueues everything and returns without synchronising."""
        torch, lib, FT = self.torch, self.lib, self.FT
        compute = torch.cuda.current_stream(self.ci.backend.device)
        cs, ks = self.copy_stream.cuda_stream, compute.cuda_stream
        # the copies of this step must not overtake the kernels of the previous one that still read the device arrays
        self.copy_stream.wait_event(self.done)
        host_ptr = {k: (v.data_ptr() if hasattr(v, "data_ptr") else v.ctypes.data) for k, v in host_ocean.items()}
        dev_ptr = {k: v.data_ptr() for k, v in self.dev.items()}
        for c, (r0, r1) in enumerate(self.copy_rows):
            off, nbytes = r0 * self.row_bytes, (r1 - r0) * self.row_bytes
            for k in self.FIELDS:
                rc = lib.dll.ne_memcpy_h2d(C.c_void_p(dev_ptr[k] + off), C.c_void_p(host_ptr[k] + off), C.c_uint64(nbytes), C.c_void_p(cs))
                if rc != 0:
                    raise RuntimeError(lib.last_error())
            self.events[c].record(self.copy_stream)
        self._set_time(t)
        if self.ci.radiation is not None:
            lib.call("interp_state", FT, self.fused.radiation, ks)
        lib.call("interp_state", FT, self.fused.atmosphere, ks)
        for c, descs in enumerate(self.band_descs):
            compute.wait_event(self.events[c])
            if descs is None:
                continue
            ao, asm, rad = descs
            lib.call("atmosphere_ocean_fluxes", FT, ao, ks)
            if asm is not None:
                lib.call("assemble_net_ocean_fluxes", FT, asm, ks)
            if rad is not None:
                lib.call("apply_radiative_fluxes", FT, rad, ks)
        if self.diagnostics is not None:
            self.diagnostics.reduce()
        self.done.record(compute)

    def h2d_bytes_per_step(self):
        g = self.ci.grid
        return len(self.FIELDS) * (g.ny + 2 * g.hy) * self.row_bytes

    def launches_per_step(self):
        n = 1 + (self.ci.radiation is not None)
        for d in self.band_descs:
            if d is not None:
                n += 1 + (d[1] is not None) + (d[2] is not None)
        return n + (2 if self.diagnostics is not None else 0)
