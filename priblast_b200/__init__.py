"""priblast_b200 — B200-native accessibility path of pRIblast's `db` step.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C ABI), data/ (energy parameters),
the ctypes binding, the `Raccess` host mirror and the synthetic workload generators.
"""
from .raccess import Raccess, packed_layout, suffix_array  # noqa: F401

__all__ = ["Raccess", "packed_layout", "suffix_array"]
__version__ = "0.1.0"
