#!/usr/bin/env python
"""TEST INFRASTRUCTURE — generate tests/golden/raccess_golden.npz from the compiled reference.

The reference ships no golden vectors, so the fixtures are outputs of the UNMODIFIED
reference `Raccess::Run(seq, acc, cond)` (raccess.cpp:42-50) built by `make -C oracle ref` with
`-O3 -ffp-contract=off` (the bit oracle).  Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

Every case is reproducible from (kind, L, seed) below; sequences are stored too so the GPU box needs
neither numpy's RNG stream nor the reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle_py import RefLib, build  # noqa: E402


def rand_seq(L, seed, alphabet="ACGU"):
    rng = np.random.default_rng(seed)
    return "".join(alphabet[k] for k in rng.integers(0, len(alphabet), L))


def cases():
    out = []
    # random ACGU over the (L, W, delta) grid of SURVEY §8c
    for L in (4, 5, 6, 30, 72, 100, 300, 500):
        for W in (20, 70, 150):
            for delta in (2, 5, 10):
                out.append((f"rand_L{L}_W{W}_d{delta}", rand_seq(L, 1000 + L), W, delta))
    # clamp / regime boundaries at W=70 (Z ~ 0.22-0.26 per nt: 88.72 near 380-420 nt, 690 near 2.7-3.1 kb)
    for L in (360, 390, 420, 2600, 2900, 3100):
        out.append((f"rand_L{L}_W70_d5", rand_seq(L, 2000 + L), 70, 5))
    out.append(("rand_L3000_W20_d5", rand_seq(3000, 5000), 20, 5))
    out.append(("rand_L1500_W150_d5", rand_seq(1500, 5001), 150, 5))
    # alphabet edge cases: N and lower case, T for U
    out.append(("mixed_case_N_L300", rand_seq(300, 3001, "ACGUNacgutT"), 70, 5))
    out.append(("allN_L50", "N" * 50, 70, 5))
    out.append(("polyA_L100", "A" * 100, 70, 5))
    out.append(("polyGC_L120", "GC" * 60, 70, 5))
    # GC stem / poly-A loop repeat: Q3 (direct path, missing-b) and Q4 (log path, Z > 690)
    unit = "GGGGGCCCCC" + "A" * 6
    out.append(("gcstem_polyA_L576", unit * 36, 70, 5))
    out.append(("gcstem_polyA_L3360", unit * 210, 70, 5))
    # perfect long hairpin (large local Boltzmann weights: range stress for any linear-domain method)
    out.append(("perfect_hairpin_L70", "G" * 33 + "AAAA" + "C" * 33, 70, 5))
    out.append(("perfect_hairpin_L150_W150", "GC" * 36 + "GAAA" + "GC" * 36 + "AA", 150, 5))
    # L around delta and TURN
    for L in (1, 2, 3, 7, 8, 9, 10, 11):
        out.append((f"tiny_L{L}", rand_seq(L, 4000 + L), 70, 5))
    return out


def main():
    build("ref")
    ref = RefLib(fast=False)
    names, seqs, Ws, deltas, accs, conds = [], [], [], [], [], []
    for name, seq, W, delta in cases():
        if len(seq) < delta:
            continue  # reference writes a negative count for L < delta (SURVEY §8a edge); excluded
        a, c = ref.run(seq, W, delta)
        names.append(name)
        seqs.append(seq)
        Ws.append(W)
        deltas.append(delta)
        accs.append(a.copy())
        conds.append(c.copy())
        print(f"{name:32s} L={len(seq):5d} W={W:3d} d={delta:2d} acc[0]={a[0] if len(a) else float('nan'):.6f}")
    out = os.path.join(HERE, "..", "tests", "golden", "raccess_golden.npz")
    np.savez_compressed(
        out,
        names=np.array(names),
        seqs=np.array(seqs),
        W=np.array(Ws, dtype=np.int32),
        delta=np.array(deltas, dtype=np.int32),
        acc=np.concatenate(accs),
        cond=np.concatenate(conds),
        lens=np.array([len(s) for s in seqs], dtype=np.int64),
    )
    print("wrote", os.path.abspath(out), os.path.getsize(out), "bytes;", len(names), "cases")


if __name__ == "__main__":
    main()
