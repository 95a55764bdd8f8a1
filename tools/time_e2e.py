"""Host-pipelined step timing with the copy / compute sides switched off in turn (development tool)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ne_b200  # noqa: E402
from numericalearth_jl_b200 import sharding, synthetic  # noqa: E402

FT = sys.argv[1] if len(sys.argv) > 1 else "f64"
chunks = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "8,16").split(",")]
backend = ne_b200.TorchCudaBackend("cuda:0")
ci = synthetic.build_case("C4", backend, FT=FT, atm_FT="f32")
ci.initialize()
f = ci.ao_fluxes
diag = sharding.FluxDiagnostics(ci, [f.latent_heat, f.sensible_heat, f.water_vapor, f.x_momentum, f.y_momentum, ci.net_ocean.T, ci.net_ocean.eta])
o = ci._host_inputs["ocean"]
pinned = {k: torch.from_numpy(np.ascontiguousarray(o[k])).pin_memory() for k in ("T", "S", "u", "v")}
use_side_stream = os.environ.get("SIDE_STREAM") == "1"
side = torch.cuda.Stream()
for nch in chunks:
    pipe = ne_b200.HostPipelinedStep(ci, n_chunks=nch, diagnostics=diag)
    for env in ({}, {"NE_B200_PIPE_NO_COPY": "1"}, {"NE_B200_PIPE_NO_COMPUTE": "1"}):
        os.environ.update(env)
        def run():
            for _ in range(3):
                pipe.step(1000.0, pinned)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            e0.record()
            for _ in range(10):
                pipe.step(1000.0, pinned)
            e1.record()
            w1 = time.perf_counter()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 10, 1e3 * (w1 - w0) / 10
        if use_side_stream:
            with torch.cuda.stream(side):
                ms, host = run()
        else:
            ms, host = run()
        print(FT, "chunks", nch, env, f"{ms:.3f} ms/step, host enqueue {host:.3f} ms/step", flush=True)
        for k in env:
            os.environ.pop(k)
