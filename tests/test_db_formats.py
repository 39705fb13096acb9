"""Byte formats of the database files written by the C++ `db` front-end (priblast_b200/pRIblast_b200)
against the UNMODIFIED reference program (oracle/_ref/pRIblast_ref = reference sources + single-rank mpi.h
stand-in).  `.seq/.ind/.nam/.bas` must be byte-identical; `.acc` needs the GPU (tests/test_gpu_db.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

FRONT = os.path.join(ROOT, "priblast_b200", "pRIblast_b200")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "pRIblast_ref")


def _fasta(path, n=18, seed=3, crlf=False):
    rng = np.random.default_rng(seed)
    eol = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        for k in range(n):
            L = int(rng.integers(20, 260))
            alpha = "ACGU" if k % 4 else "ACGTNacgu"
            s = "".join(alpha[i] for i in rng.integers(0, len(alpha), L))
            f.write(f">seq{k} some description{eol}")
            for p in range(0, L, 60):
                f.write(s[p:p + 60] + eol)


def _build_front():
    subprocess.run(["make", "-C", os.path.join(ROOT, "priblast_b200", "csrc", "host")], check=True,
                   stdout=subprocess.DEVNULL)


@pytest.mark.parametrize("extra", [[], ["-c", "7"], ["-r", "1", "-s", "5", "-c", "5"], ["-r", "2", "-s", "3"]])
def test_seq_ind_nam_bas_bytes_match_reference(tmp_path, extra):
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref/pRIblast_ref not built")
    _build_front()
    fa = str(tmp_path / "in.fa")
    _fasta(fa)
    subprocess.run([REFBIN, "db", "-i", fa, "-o", str(tmp_path / "ref"), "-w", "30"] + extra, check=True,
                   env=dict(os.environ, OMP_NUM_THREADS="4"), cwd=tmp_path)
    env = dict(os.environ, PRIB_DB_FORMATS_ONLY="1")
    subprocess.run([FRONT, "db", "-i", fa, "-o", str(tmp_path / "new"), "-w", "30"] + extra, check=True, env=env)
    for ext in (".seq", ".ind", ".nam", ".bas"):
        a = open(str(tmp_path / "ref") + ext, "rb").read()
        b = open(str(tmp_path / "new") + ext, "rb").read()
        assert a == b, f"{ext} differs ({len(a)} vs {len(b)} bytes)"


def test_crlf_fasta_matches_reference(tmp_path):
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref/pRIblast_ref not built")
    _build_front()
    fa = str(tmp_path / "in.fa")
    _fasta(fa, n=6, crlf=True)
    subprocess.run([REFBIN, "db", "-i", fa, "-o", str(tmp_path / "ref"), "-w", "20"], check=True, cwd=tmp_path)
    subprocess.run([FRONT, "db", "-i", fa, "-o", str(tmp_path / "new"), "-w", "20"], check=True,
                   env=dict(os.environ, PRIB_DB_FORMATS_ONLY="1"))
    for ext in (".seq", ".ind", ".nam"):
        assert open(str(tmp_path / "ref") + ext, "rb").read() == open(str(tmp_path / "new") + ext, "rb").read(), ext


def test_cli_errors_mirror_reference(tmp_path):
    _build_front()
    fa = str(tmp_path / "in.fa")
    _fasta(fa, n=3)
    env = dict(os.environ, PRIB_DB_FORMATS_ONLY="1")

    def run(*args):
        p = subprocess.run([FRONT, "db", *args], env=env, capture_output=True, text=True)
        return p.returncode, p.stderr

    assert run("-i", fa) == (1, "Error: -o option is required\n")                      # raccess.hpp:42-45
    assert run("-i", fa, "-o", "x", "-d", "1") == (1, "Error: -d option must be greater than 1\n")
    assert run("-i", fa, "-o", "x", "-r", "3") == (1, "Error: -r option must be 0, 1, or 2\n")
    assert run("-i", fa, "-o", "x", "-a", "area") == (1, "Error: parallel algorithm not supported\n")
    rc, err = run("-i", str(tmp_path / "missing.fa"), "-o", "x")
    assert rc == 1 and "can't open input_file" in err
    rc, err = run("-i", fa, "-o", "x", "-t", "4")
    assert rc == 1 and err.strip().endswith("Error: invalid argument")


def test_suffix_array_and_partitioner():
    _build_front()
    lib = ctypes.CDLL(os.path.join(ROOT, "priblast_b200", "libprib_dbformat.so"))
    rng = np.random.default_rng(0)
    for n in (1, 2, 17, 500, 3000):
        text = rng.integers(0, 4, n).astype(np.uint8) + 2
        text[rng.integers(0, n, max(1, n // 50))] = 0   # sentinels
        text[-1] = 0
        sa = np.zeros(n, np.int32)
        lib.prib_suffix_array_host(text.ctypes.data_as(ctypes.c_void_p), n, sa.ctypes.data_as(ctypes.c_void_p))
        raw = text.tobytes()
        want = sorted(range(n), key=lambda i: raw[i:])
        assert sa.tolist() == want
    # homopolymer (worst case for naive sorting)
    text = np.full(2000, 2, np.uint8)
    text[-1] = 0
    sa = np.zeros(2000, np.int32)
    lib.prib_suffix_array_host(text.ctypes.data_as(ctypes.c_void_p), 2000, sa.ctypes.data_as(ctypes.c_void_p))
    assert sa.tolist() == list(range(1999, -1, -1))
    # LPT: every sequence assigned once, loads balanced within the longest item
    lens = np.clip(rng.lognormal(np.log(1500), 0.75, 4000), 200, 5000).astype(np.int32)
    for parts in (1, 2, 8):
        part = np.full(len(lens), -1, np.int32)
        lib.prib_lpt_partition(len(lens), lens.ctypes.data_as(ctypes.c_void_p), parts, part.ctypes.data_as(ctypes.c_void_p))
        assert part.min() >= 0 and part.max() == parts - 1
        loads = np.bincount(part, weights=lens, minlength=parts)
        assert loads.max() - loads.min() <= lens.max()
