"""Build libne_b200.so (sm_100a) in-tree with nvcc.  Called by __graft_entry__.build().

    python numericalearth.jl_b200/csrc/build.py [--force]

Every translation unit is compiled with
    -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
and never with --use_fast_math.  ne_interp_kernels.cu additionally gets -fmad=false: the
interpolation indices/weights must be bit-exact with the reference, which never contracts a*b+c.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "libne_b200.so")
OBJ = os.path.join(HERE, "_obj")

SOURCES = {
    "ne_api.cu": [],
    "ne_flux_kernels.cu": [],
    "ne_flux_tab2.cu": [],
    "ne_flux_generic_ao_f64.cu": [],
    "ne_flux_generic_ao_f32.cu": [],
    "ne_flux_generic_asi_f64.cu": [],
    "ne_flux_generic_asi_f32.cu": [],
    "ne_flux_generic_al.cu": [],
    "ne_flux_generic_al_f32.cu": [],
    "ne_flux_queue_ao.cu": [],
    "ne_flux_queue_asi.cu": [],
    "ne_flux_queue_land.cu": [],
    "ne_interp_kernels.cu": ["-fmad=false"],
    "ne_surface_kernels.cu": ["-fmad=false"],
    "ne_fused.cu": [],
    "ne_pipeline.cu": [],
    "ne_series_ring.cu": [],
}
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
DEPS = ["ne_common.cuh", "ne_physics.cuh", os.path.join("..", "..", "include", "ne_b200.h")]
EXTRA_DEPS = {
    "ne_flux_kernels.cu": ["ne_flux_fast.cuh", "ne_flux_tab.cuh", "ne_flux_queue.cuh", "ne_flux_asi_fast.cuh", "ne_flux_land_fast.cuh", "ne_queue_host.cuh", "ne_fastmath.cuh", "ne_interp_device.cuh"],
    "ne_interp_kernels.cu": ["ne_interp_device.cuh"],
    "ne_flux_tab2.cu": ["ne_flux_fast.cuh", "ne_flux_tab.cuh", "ne_flux_tab2.cuh", "ne_fastmath.cuh", "ne_queue_host.cuh"],
    "ne_flux_generic_ao_f64.cu": ["ne_flux_generic.cuh"],
    "ne_flux_generic_ao_f32.cu": ["ne_flux_generic.cuh"],
    "ne_flux_generic_asi_f64.cu": ["ne_flux_generic.cuh"],
    "ne_flux_generic_asi_f32.cu": ["ne_flux_generic.cuh"],
    "ne_flux_generic_al.cu": ["ne_flux_generic.cuh"],
    "ne_flux_generic_al_f32.cu": ["ne_flux_generic.cuh"],
    "ne_flux_queue_ao.cu": ["ne_flux_fast.cuh", "ne_flux_tab.cuh", "ne_fastmath.cuh", "ne_flux_queue.cuh", "ne_queue_host.cuh"],
    "ne_flux_queue_land.cu": ["ne_flux_fast.cuh", "ne_flux_tab.cuh", "ne_fastmath.cuh", "ne_flux_queue.cuh", "ne_flux_land_fast.cuh", "ne_queue_host.cuh"],
    "ne_flux_queue_asi.cu": ["ne_flux_fast.cuh", "ne_flux_tab.cuh", "ne_fastmath.cuh", "ne_flux_queue.cuh", "ne_flux_asi_fast.cuh", "ne_queue_host.cuh"],
    "ne_fused.cu": ["ne_flux_fast.cuh"],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    deps = [os.path.join(HERE, d) for d in DEPS]
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + deps + [os.path.join(HERE, d) for d in EXTRA_DEPS.get(src, [])]):
            jobs.append((src, [nvcc] + COMMON + extra + ["-c", s, "-o", o]))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(OBJ, src + ".log"), "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 2)) as ex:
            for src in ex.map(run, jobs):
                if verbose:
                    print("compiled", src)
    if jobs or force or _stale(OUT, objs):
        cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
