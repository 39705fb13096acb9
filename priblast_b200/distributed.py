"""One-process-per-GPU sharding of the `db` accessibility step (the torchrun path; the C++ front-end does
the same with one thread per GPU).  Sequences are independent, so there is NO data-path collective: each
rank computes its length-balanced shard, and a host-side gather on rank 0 reassembles the `.acc` image in
FASTA order.  Replaces the MPI distributors of fastafile_reader.cpp:135-314 (+ the RMA work counter of
db_construction.cpp:85-95, 191-197) for this step."""
from __future__ import annotations

import heapq
from typing import Callable, Sequence

import numpy as np

from .raccess import packed_layout


def lpt_shard(lens: Sequence[int], world: int) -> list[np.ndarray]:
    """Longest-processing-time greedy on cost = length (same rule as csrc/host/db_format.cpp lpt_partition
    and, in spirit, the reference's heap distributor fastafile_reader.cpp:270-283).  Deterministic."""
    lens = np.asarray(lens, dtype=np.int64)
    order = np.argsort(-lens, kind="stable")
    heap = [(0, d) for d in range(world)]
    heapq.heapify(heap)
    parts: list[list[int]] = [[] for _ in range(world)]
    for idx in order:
        load, d = heapq.heappop(heap)
        parts[d].append(int(idx))
        heapq.heappush(heap, (load + int(lens[idx]), d))
    return [np.asarray(p, dtype=np.int64) for p in parts]


def run_sharded(seqs: Sequence[bytes], compute: Callable[[list], list], rank: int, world: int, gather=None):
    """compute(list_of_sequences) -> [(acc, cond), ...] on this rank's device.  Returns, on rank 0, the
    packed float32 image [acc L | cond L] per sequence in input order (None elsewhere).
    gather(obj) -> list of objects on rank 0: defaults to torch.distributed.gather_object."""
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    parts = lpt_shard(lens, world)
    mine = parts[rank]
    res = compute([seqs[k] for k in mine]) if len(mine) else []
    payload = (mine, [np.asarray(a, np.float32) for a, _ in res], [np.asarray(c, np.float32) for _, c in res])
    if world == 1:
        gathered = [payload]
    elif gather is not None:
        gathered = gather(payload)
    else:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return None
    acc_off, cond_off, total = packed_layout(lens)
    image = np.zeros(max(total, 1), dtype=np.float32)
    seen = np.zeros(len(seqs), dtype=bool)
    for ids, accs, conds in gathered:
        for k, a, c in zip(ids, accs, conds):
            image[acc_off[k]:acc_off[k] + lens[k]] = a
            image[cond_off[k]:cond_off[k] + lens[k]] = c
            seen[k] = True
    assert seen.all(), "a sequence was not computed by any rank"
    return image
