"""Host side of the Python mirror: how sequences reach the C ABI (no GPU, no compute call)."""
import ctypes

import numpy as np
import pytest

from priblast_b200 import raccess


class _NoCtx(raccess.Raccess):
    def __init__(self):  # the marshalling needs no device context
        pass

    def close(self):
        pass


def _check(seqs, want):
    keep, n, lens, arr = _NoCtx()._marshal(seqs)
    assert n == len(want) and lens.dtype == np.int32 and list(lens) == [len(w) for w in want]
    ptrs = keep[1]
    for k, w in enumerate(want):
        assert ctypes.string_at(int(ptrs[k]), len(w)) == w
    return keep


@pytest.mark.parametrize("zero_copy", [True, False])
def test_pointer_array_addresses_the_sequences(monkeypatch, zero_copy):
    if zero_copy:
        assert raccess._BYTES_OFFSET == bytes.__basicsize__ - 1  # the import-time probe accepted this interpreter
    else:
        monkeypatch.setattr(raccess, "_BYTES_OFFSET", None)      # the fallback: one joined buffer
    rng = np.random.default_rng(5)
    seqs = [bytes(rng.choice(list(b"ACGU"), size=int(L)).astype(np.uint8)) for L in rng.integers(1, 400, size=300)]
    _check(seqs, seqs)
    _check(["ACGU", b"GGGAAACCC", "acgun"], [b"ACGU", b"GGGAAACCC", b"acgun"])
    _check([b"", b"A", b""], [b"", b"A", b""])
    keep, n, lens, arr = _NoCtx()._marshal([])
    assert n == 0 and len(lens) == 0


def test_zero_copy_points_into_the_callers_objects():
    seqs = [b"ACGUACGU" * 10, b"GGGG" * 7]
    keep, n, lens, arr = _NoCtx()._marshal(seqs)
    off = raccess._BYTES_OFFSET
    assert off is not None
    assert [int(p) for p in keep[1][:n]] == [id(s) + off for s in seqs]
