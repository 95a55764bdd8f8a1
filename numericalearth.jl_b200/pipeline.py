"""Host-buffer interface step: the ocean surface state (T, S, u, v) arrives in pinned HOST memory
every coupled step (an ocean model that lives on the host, or on another device behind a host
bounce buffer), the interface fluxes are wanted back as global diagnostics on the host.

The reference does this serially (copy, then `update_state!`).  Here the exchange grid is cut into
latitude chunks and the whole sequence is enqueued by ONE native call (`ne_host_pipelined_step_*`,
csrc/ne_pipeline.cu): chunk c's rows of the four fields go host->device on the pipeline's copy stream
while the compute stream runs the a–o solve and the post-solve kernel (net-flux assembly + radiation)
on the launch rows whose inputs (including the (j+1) row of the v stencil,
atmosphere_ocean_fluxes.jl:64-65) have already landed — the descriptors' launch range
(`NeExchangeGrid.j_lo/j_hi`) restricts each call to its rows.  The atmosphere interpolation does not
depend on the ocean and runs up front, hidden behind the first copy; the diagnostics sums run once
after the last band.  PCIe time and kernel time overlap instead of adding up; results are identical
to the unchunked step (pointwise kernels, and the (i-1, j-1) stencils of the assembly only reach back
into rows an earlier band computed).
"""
import ctypes as C

from . import abi as A


class HostPipelinedStep:
    FIELDS = ("T", "S", "u", "v")

    def __init__(self, interfaces, n_chunks=8, diagnostics=None, return_fields=None):
        """return_fields: dict name -> (device array, pinned host array) of exchange-layout results (e.g. the six net
        ocean fluxes) copied back to the host band by band behind the kernels; synchronising the compute stream then
        also means the results are on the host."""
        ci = self.ci = interfaces
        if not ci.backend.is_device:
            raise RuntimeError("HostPipelinedStep needs the CUDA library and device arrays (there is no CPU fallback)")
        g = ci.grid
        self.lib = ci.lib
        self.FT = g.FT
        self.diagnostics = diagnostics
        self.dev = {"T": ci.ocean_state.T, "S": ci.ocean_state.S, "u": ci.ocean_state.u, "v": ci.ocean_state.v}
        self.row_bytes = (g.nx + 2 * g.hx) * (8 if g.FT == "f64" else 4)
        self.n_chunks = max(1, min(int(n_chunks), g.ny))
        self.handle = C.c_void_p()
        rc = self.lib.dll.ne_host_pipeline_create(C.byref(self.handle), C.c_int32(self.n_chunks))
        if rc != 0:
            raise RuntimeError(self.lib.last_error())
        d = self.desc = A.NeHostStepDesc()
        d.step = ci.fused_step_desc(0.0, diagnostics=diagnostics)
        d.n_fields, d.n_chunks, d.row_bytes = len(self.FIELDS), self.n_chunks, self.row_bytes
        for k, name in enumerate(self.FIELDS):
            d.fields[k].device = ci.backend.ptr(self.dev[name])
        self.return_fields = dict(return_fields or {})
        if len(self.return_fields) > A.NE_HOST_MAX_FIELDS:
            raise ValueError(f"at most {A.NE_HOST_MAX_FIELDS} fields can be returned to the host")
        d.n_out_fields = len(self.return_fields)
        for k, (dev, host) in enumerate(self.return_fields.values()):
            d.out_fields[k].device = ci.backend.ptr(dev)
            d.out_fields[k].host = host.data_ptr() if hasattr(host, "data_ptr") else host.ctypes.data
        self._fn = getattr(self.lib.dll, "ne_host_pipelined_step_" + self.FT)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.dll.ne_host_pipeline_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _set_time(self, t):
        ci = self.ci
        # clock-dependent surface properties (TabulatedAlbedo: seconds in the day, solar declination) are host scalars
        # inside the radiation PODs: refresh them every step, as fused_interface_step does by rebuilding its descriptor
        ci.clock_time = t
        if ci.radiation is not None:
            self.desc.step.ao.radiation = ci._surface_radiation("ocean")
            self.desc.step.apply_radiation.radiation = ci._surface_radiation("ocean")
        for d, src in ((self.desc.step.atmosphere, ci.atmosphere), (self.desc.step.radiation, ci.radiation)):
            if src is None:
                continue
            ti = ci._time_interp(src, t)
            d.time.frac, d.time.m1, d.time.m2, d.time.same = ti.frac, ti.m1, ti.m2, ti.same

    def step(self, t, host_ocean, ocean_column=None):
        """host_ocean: dict of pinned host arrays (torch CPU tensors or numpy) named T, S, u, v with the exchange
        layout.  Enqueues everything on the current stream (+ the pipeline's copy stream) and returns without
        synchronising."""
        d = self.desc
        for k, name in enumerate(self.FIELDS):
            v = host_ocean[name]
            d.fields[k].host = v.data_ptr() if hasattr(v, "data_ptr") else v.ctypes.data
        self._set_time(t)
        if self.ci.land is not None:   # PrescribedLand runoff: independent of the ocean state, ahead of the pipeline
            self.lib.call("interp_state", self.FT, self.ci.land_interp_desc(t), self.ci.backend.stream())
        rc = self._fn(self.handle, C.byref(d), C.c_void_p(self.ci.backend.stream()))
        if rc != 0:
            if rc == A.NE_E_NO_VARIANT:
                from .formulations import NoKernelVariantError
                raise NoKernelVariantError(self.lib.last_error())
            raise RuntimeError(self.lib.last_error())
        self.ci.release_windows()
        self.ci.compute_sea_ice_ocean_fluxes(ocean_column)   # FreezingLimitedOceanTemperature clamp (device-resident column)
        if self.diagnostics is not None:
            self.diagnostics.all_reduce()

    def h2d_bytes_per_step(self):
        g = self.ci.grid
        return len(self.FIELDS) * (g.ny + 2 * g.hy) * self.row_bytes

    def d2h_bytes_per_step(self):
        g = self.ci.grid
        return len(self.return_fields) * (g.ny + 2 * g.hy) * self.row_bytes

    def launches_per_step(self):
        """Kernel launches of one step: 2 interpolations, (solve + post-solve) per band, 2 diagnostics stages."""
        return 1 + (self.ci.radiation is not None) + 2 * self.n_chunks + (2 if self.diagnostics is not None else 0)
