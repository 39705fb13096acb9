"""Synthetic transcript sets of BASELINE.json's configs.  numpy `default_rng` (PCG64) with
fixed seeds, so every box regenerates identical inputs."""
from __future__ import annotations

import numpy as np

_ALPHA = np.frombuffer(b"ACGU", dtype=np.uint8)


def _bases_uniform(rng, total: int) -> np.ndarray:
    return _ALPHA[rng.integers(0, 4, total)]


def _bases_gc(rng, total: int, gc: float) -> np.ndarray:
    u = rng.random(total)
    half = gc / 2.0
    # A | C | G | U with P(C)=P(G)=gc/2
    code = np.where(u < half, 1, np.where(u < gc, 2, np.where(u < gc + (1 - gc) / 2, 0, 3)))
    return _ALPHA[code]


def _split(bases: np.ndarray, lens: np.ndarray) -> list[bytes]:
    out, p = [], 0
    raw = bases.tobytes()
    for L in lens:
        out.append(raw[p:p + int(L)])
        p += int(L)
    return out


def cfg1(n: int = 1000, L: int = 500, seed: int = 1) -> list[bytes]:
    """1,000 random RNAs of 500 nt, uniform ACGU (config 0 of BASELINE.json)."""
    rng = np.random.default_rng(seed)
    lens = np.full(n, L, dtype=np.int64)
    return _split(_bases_uniform(rng, int(lens.sum())), lens)


def cfg2_lengths(n: int = 100_000, seed: int = 2) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(rng.lognormal(np.log(1500.0), 0.75, n)), 200, 5000).astype(np.int64)


def cfg2(n: int = 100_000, seed_len: int = 2, seed_base: int = 3, gc: float = 0.45, first: int | None = None):
    """GENCODE-like: length = clip(round(LogNormal(ln 1500, 0.75)), 200, 5000), GC 0.45 (config 1).
    `first` keeps only the first k sequences of the same stream (bounded samples)."""
    lens = cfg2_lengths(n, seed_len)
    if first is not None:
        lens = lens[:first]
    rng = np.random.default_rng(seed_base)
    return _split(_bases_gc(rng, int(lens.sum()), gc), lens)


def cfg3(n: int = 2000, seed_len: int = 4, seed_base: int = 5, first: int | None = None):
    """long-lncRNA stress: length = round(exp(U(ln 1e4, ln 1e5))) (config 2)."""
    rng = np.random.default_rng(seed_len)
    lens = np.rint(np.exp(rng.uniform(np.log(1e4), np.log(1e5), n))).astype(np.int64)
    if first is not None:
        lens = lens[:first]
    rng = np.random.default_rng(seed_base)
    return _split(_bases_uniform(rng, int(lens.sum())), lens)


def cfg4(n: int = 10_000, L: int = 2000, seed: int = 6, first: int | None = None):
    """span sweep: 10k uniform sequences of 2 kb (config 3); W is chosen by the caller."""
    if first is not None:
        n = first
    rng = np.random.default_rng(seed)
    lens = np.full(n, L, dtype=np.int64)
    return _split(_bases_uniform(rng, int(lens.sum())), lens)


def write_fasta(path: str, seqs, prefix: str = "r", width: int = 60) -> None:
    with open(path, "w") as f:
        for k, s in enumerate(seqs):
            s = s.decode() if isinstance(s, bytes) else s
            f.write(f">{prefix}{k}\n")
            for p in range(0, len(s), width):
                f.write(s[p:p + width] + "\n")
