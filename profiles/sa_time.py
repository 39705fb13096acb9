"""Suffix-array construction of encoded database pages: GPU (prib_suffix_array) vs the host checker."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from priblast_b200 import suffix_array, workloads
rng = np.random.default_rng(5)
host = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "priblast_b200", "libprib_dbformat.so"))
for n in (2000, 20000, 100000):
    lens = workloads.cfg2_lengths(100_000)[:n]
    parts = []
    for L in lens:
        parts.append(np.asarray((2, 3, 4, 5), np.uint8)[rng.integers(0, 4, int(L))]); parts.append(np.zeros(1, np.uint8))
    text = np.concatenate(parts)
    suffix_array(text[:1000])
    t0 = time.perf_counter(); sa = suffix_array(text); dt = time.perf_counter() - t0
    line = f"{n} transcripts, {len(text)} symbols: GPU {dt:.3f} s ({len(text)/dt:.3e} symbols/s)"
    if n <= 20000:
        ref = np.zeros(len(text), np.int32)
        t0 = time.perf_counter(); host.prib_suffix_array_host(text.ctypes.data_as(ctypes.c_void_p), len(text), ref.ctypes.data_as(ctypes.c_void_p)); dh = time.perf_counter() - t0
        line += f"; host checker {dh:.2f} s ({len(text)/dh:.3e}/s); equal={bool(np.array_equal(sa, ref))}"
    print(line, flush=True)
