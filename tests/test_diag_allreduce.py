"""ne_diag_allreduce_f64: the one collective of the path behind the C-ABI (include/ne_b200.h).  The communicator is the
host's: the test creates it the way NCCL.jl / an MPI host would (ncclGetUniqueId + ncclCommInitRank on libnccl) and hands the
raw ncclComm_t to the library."""
import ctypes as C
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def _nccl():
    import torch  # noqa: F401  (maps torch's libnccl.so.2 into the process)
    cands = []
    for p in sys.path:
        cands += glob.glob(os.path.join(p, "nvidia", "nccl", "lib", "libnccl.so.2"))
    for name in cands + ["libnccl.so.2"]:
        try:
            return C.CDLL(name, mode=C.RTLD_GLOBAL)
        except OSError:
            continue
    pytest.skip("libnccl.so.2 not found")


def _worker(rank, world, uid_bytes, ret):
    import torch
    import ne_b200
    nccl = _nccl()
    torch.cuda.set_device(rank)
    uid = _UniqueId()
    C.memmove(C.byref(uid), uid_bytes, 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
    rc = nccl.ncclCommInitRank(C.byref(comm), world, uid, rank)
    assert rc == 0, f"ncclCommInitRank rc={rc}"
    lib = ne_b200.get_library()
    sums = torch.arange(7, dtype=torch.float64, device=f"cuda:{rank}") * (rank + 1) + 0.25
    stream = torch.cuda.current_stream().cuda_stream
    lib.diag_allreduce(comm.value, sums.data_ptr(), 7, stream)
    torch.cuda.synchronize()
    expect = sum(np.arange(7, dtype=np.float64) * (r + 1) + 0.25 for r in range(world))
    ok = bool(np.array_equal(sums.cpu().numpy(), expect))
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    ret[rank] = ok


def _run(world):
    import torch
    import torch.multiprocessing as mp
    nccl = _nccl()
    uid = _UniqueId()
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    uid_bytes = bytes(uid.internal.raw if hasattr(uid.internal, "raw") else C.string_at(C.byref(uid), 128))
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        ret = m.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, C.string_at(C.byref(uid), 128), ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert all(ret.get(r) for r in range(world)), dict(ret)


@pytest.mark.gpu
def test_allreduce_single_rank_communicator():
    _run(1)


@pytest.mark.gpu
def test_allreduce_two_ranks_over_nvlink():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    _run(2)


def test_allreduce_rejects_null_communicator():
    import ne_b200
    lib = ne_b200.Library()
    with pytest.raises(ne_b200.NeError) as e:
        lib.diag_allreduce(0, 0, 7, 0)
    assert e.value.code == ne_b200.abi.NE_E_INVALID
