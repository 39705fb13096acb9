"""What would a tensor-core (TF32 3-product split) generic-loop stencil do to parity?  CPU emulation.

tests/hostemu runs the device engine's per-thread bodies in plain loops.  With hostemu_set_tf32_split(1) every product
of the generic-loop stencils of the two tile passes (acc_tile.h in_rows / out_rows) is computed the way
profiles/tc_probe.cu's tensor-core formulation computes it: both operands split into TF32 halves, x_hi w_hi +
x_lo w_hi + x_hi w_lo, FP32 accumulation.  Everything else stays the FP32 engine.  Compared against the reference
values of the committed fixtures (tests/golden/), next to the unmodified FP32 engine.

python profiles/tc_parity.py  > profiles/r2/tc_parity.txt      (CPU only, ~1 min)
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLDEN  # noqa: E402
from test_hostemu import _tiled  # noqa: E402

d = os.path.join(ROOT, "tests", "hostemu")
subprocess.run(["make", "-C", d], check=True, stdout=subprocess.DEVNULL)
lib = ctypes.CDLL(os.path.join(d, "libhostemu.so"))
SCALE = (0.45, 4.0, 16.0)  # the product's span scaling (acc_tables.h default_scale_fp32)

rows = []
for case in GOLDEN:
    L = len(case["seq"])
    if L < 200 or case["W"] != 70:
        continue
    ref = np.concatenate([np.asarray(case["acc"], np.float64), np.asarray(case["cond"], np.float64)])
    errs = []
    for split in (0, 1):
        lib.hostemu_set_tf32_split(split)
        out, flags, _ = _tiled(lib, [case["seq"]], case["W"], case["delta"], 256, SCALE, f32=True)
        if flags[0]:
            errs = None
            break
        errs.append(np.abs(out[:2 * L].astype(np.float64) - ref))
    lib.hostemu_set_tf32_split(0)
    if errs is None:
        continue
    rows.append((case["name"], L, errs[0].max(), errs[0].mean(), errs[1].max(), errs[1].mean()))

print("fixture                         L    FP32 engine max / mean      TF32-split stencil max / mean   (kcal/mol vs reference)")
for r in rows:
    print(f"{r[0]:<28} {r[1]:>5}    {r[2]:.2e} / {r[3]:.2e}        {r[4]:.2e} / {r[5]:.2e}")
a = np.array([r[2:] for r in rows])
print(f"{'all ' + str(len(rows)) + ' fixtures (W = 70, L >= 200)':<34}    {a[:, 0].max():.2e} / {a[:, 1].mean():.2e}        "
      f"{a[:, 2].max():.2e} / {a[:, 3].mean():.2e}")
print("gate of the parity tests: max <= 1e-4 + 1e-6 |x|, mean <= 5e-6")
