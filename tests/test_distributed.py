"""The N>1 path on CPU: gloo, world_size 2.  The device computation is replaced by a deterministic stand-in
(the product has no CPU compute path); what is tested is sharding, gather and reassembly order."""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _fake_compute(seqs):
    out = []
    for s in seqs:
        L = len(s)
        h = float(sum(s[:8])) if L else 0.0
        out.append((np.arange(L, dtype=np.float32) + h, np.full(L, h, dtype=np.float32)))
    return out


def _seqs():
    from priblast_b200 import workloads
    return workloads.cfg2(first=40)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from priblast_b200.distributed import run_sharded
    img = run_sharded(_seqs(), _fake_compute, rank, world)
    t = torch.tensor([float(len(_seqs()))])
    dist.all_reduce(t)  # the barrier + reduction pattern bench.py uses
    if rank == 0:
        q.put(img)
    dist.destroy_process_group()


def test_two_rank_gloo_reassembly_matches_single_rank():
    from priblast_b200.distributed import run_sharded
    want = run_sharded(_seqs(), _fake_compute, 0, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(got, want)


def test_python_and_cpp_partitioners_agree():
    import subprocess
    from priblast_b200.distributed import lpt_shard
    subprocess.run(["make", "-C", os.path.join(ROOT, "priblast_b200", "csrc", "host")], check=True,
                   stdout=subprocess.DEVNULL)
    lib = ctypes.CDLL(os.path.join(ROOT, "priblast_b200", "libprib_dbformat.so"))
    lens = np.array([len(s) for s in _seqs()], dtype=np.int32)
    for world in (1, 2, 4, 8):
        part = np.full(len(lens), -1, np.int32)
        lib.prib_lpt_partition(len(lens), lens.ctypes.data_as(ctypes.c_void_p), world, part.ctypes.data_as(ctypes.c_void_p))
        py = lpt_shard(lens, world)
        for d, ids in enumerate(py):
            assert set(ids.tolist()) == set(np.nonzero(part == d)[0].tolist()), (world, d)
        loads = [int(lens[p].sum()) for p in py]
        assert max(loads) - min(loads) <= int(lens.max())
