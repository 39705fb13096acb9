"""TEST INFRASTRUCTURE — ctypes bindings for the checkers under oracle/.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.  The
product package (priblast_b200) never does.

  * RefLib    -> oracle/_ref/libpriblast_ref*.so : the UNMODIFIED reference Raccess (raccess.cpp:34-50)
                 behind oracle/ref_shim.cpp.
  * OracleLib -> oracle/liboracle.so             : the plain-C restatement (oracle/raccess_oracle.c).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def _as(ptr_t, arr):
    return arr.ctypes.data_as(ptr_t)


def build(target: str = "all") -> None:
    """(Re)build the checkers; `ref` is a no-op when /root/reference is absent (GPU box)."""
    subprocess.run(["make", "-C", HERE, target], check=True, stdout=subprocess.DEVNULL)


def acc_layout(lens):
    """Float offsets of (acc, cond) per sequence in a packed [acc L | cond L] image."""
    lens = np.asarray(lens, dtype=np.int64)
    base = np.concatenate([[0], np.cumsum(2 * lens)])[:-1]
    return base.astype(np.int64), (base + lens).astype(np.int64), int((2 * lens).sum())


class _BatchMixin:
    _batch_fn = None

    def run(self, seq: str | bytes, W: int = 70, delta: int = 5):
        b = seq.encode() if isinstance(seq, str) else bytes(seq)
        L = len(b)
        acc = np.zeros(max(L, 1), dtype=np.float32)
        cond = np.zeros(max(L, 1), dtype=np.float32)
        rc = self._run_fn(b, L, W, delta, _as(_f32p, acc), _as(_f32p, cond))
        if rc != 0:
            raise RuntimeError(f"oracle run failed rc={rc}")
        return acc[:L], cond[:L]

    def run_batch(self, seqs, W: int = 70, delta: int = 5, nthreads: int = 0):
        bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
        n = len(bs)
        lens = np.array([len(b) for b in bs], dtype=np.int32)
        acc_off, cond_off, total = acc_layout(lens)
        out = np.zeros(max(total, 1), dtype=np.float32)
        arr = (ctypes.c_char_p * n)(*bs)
        used = self._batch_fn(n, arr, _as(_i32p, lens), W, delta, _as(_f32p, out),
                              _as(_i64p, acc_off), _as(_i64p, cond_off), nthreads)
        if used < 0:
            raise RuntimeError(f"oracle batch failed rc={used}")
        res = [(out[a:a + l], out[c:c + l]) for a, c, l in zip(acc_off, cond_off, lens)]
        return res, used


class RefLib(_BatchMixin):
    """The compiled, unmodified reference.  fast=False: bit oracle (-ffp-contract=off)."""

    def __init__(self, fast: bool = False):
        name = "libpriblast_ref_fast.so" if fast else "libpriblast_ref.so"
        self.path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(self.path):
            raise FileNotFoundError(self.path)
        self.lib = ctypes.CDLL(self.path)
        self._run_fn = self.lib.ref_raccess_run
        self._run_fn.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _f32p]
        self._batch_fn = self.lib.ref_raccess_batch
        self._batch_fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), _i32p, ctypes.c_int,
                                   ctypes.c_int, _f32p, _i64p, _i64p, ctypes.c_int]
        self.lib.ref_raccess_dump.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              _f64p, _f64p, _f64p]
        self.lib.ref_raccess_run_file.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_char_p, ctypes.c_int]

    @staticmethod
    def available(fast: bool = False) -> bool:
        name = "libpriblast_ref_fast.so" if fast else "libpriblast_ref.so"
        return os.path.exists(os.path.join(HERE, "_ref", name))

    def dump(self, seq: str, W: int = 70, delta: int = 5):
        """Reference DP state: dict name -> (L+1, W+2) float64, plus alpha_outer/beta_outer."""
        b = seq.encode()
        L = len(b)
        band = np.zeros((12, L + 1, W + 2), dtype=np.float64)
        ao = np.zeros(L + 1, dtype=np.float64)
        bo = np.zeros(L + 1, dtype=np.float64)
        self.lib.ref_raccess_dump(b, L, W, delta, _as(_f64p, band), _as(_f64p, ao), _as(_f64p, bo))
        names = ["a_stem", "a_stemend", "a_multi", "a_multibif", "a_multi1", "a_multi2",
                 "b_stem", "b_stemend", "b_multi", "b_multibif", "b_multi1", "b_multi2"]
        d = {n: band[k] for k, n in enumerate(names)}
        d["alpha_outer"] = ao
        d["beta_outer"] = bo
        return d

    def run_file_bytes(self, seq: str, W: int, delta: int, tmpdir: str, idx: int = 0) -> bytes:
        """Bytes of the per-sequence temp .acc file the reference writes (raccess.cpp:447-481)."""
        b = seq.encode()
        self.lib.ref_raccess_run_file(b, len(b), W, delta, tmpdir.encode(), idx)
        p = os.path.join(tmpdir, f"priblast_tmp_acc0_{idx}.acc")
        with open(p, "rb") as f:
            data = f.read()
        os.unlink(p)
        return data


class OracleLib(_BatchMixin):
    """The plain-C restatement (oracle/raccess_oracle.c)."""

    def __init__(self):
        self.path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(self.path):
            build("oracle")
        self.lib = ctypes.CDLL(self.path)
        self._run_fn = self.lib.oracle_raccess_run
        self._run_fn.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p, _f32p]
        self._batch_fn = self.lib.oracle_raccess_batch
        self._batch_fn.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), _i32p, ctypes.c_int,
                                   ctypes.c_int, _f32p, _i64p, _i64p, ctypes.c_int]
        self.lib.oracle_raccess_count.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, _i64p]
        self.lib.oracle_raccess_exact.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, _f32p, _f32p]

    def count_terms(self, seq: str, W: int = 70, delta: int = 5):
        """Algorithmic work of one sequence: dict of counters."""
        b = seq.encode() if isinstance(seq, str) else bytes(seq)
        c = np.zeros(8, dtype=np.int64)
        self.lib.oracle_raccess_count(b, len(b), W, delta, _as(_i64p, c))
        keys = ["lse_inside", "lse_outside", "lse_access", "expd_access", "loop_energy", "cells",
                "final_logs", "reserved"]
        return dict(zip(keys, (int(x) for x in c)))

    def run_exact(self, seq: str, W: int = 70, delta: int = 5):
        """Same recurrences with exact libm log1p(exp()) in place of the fmath tables (noise probe)."""
        b = seq.encode() if isinstance(seq, str) else bytes(seq)
        L = len(b)
        acc = np.zeros(max(L, 1), dtype=np.float32)
        cond = np.zeros(max(L, 1), dtype=np.float32)
        self.lib.oracle_raccess_exact(b, L, W, delta, _as(_f32p, acc), _as(_f32p, cond))
        return acc[:L], cond[:L]
