"""Builds experiment variants of libpriblast_acc.so (same sources, different -D flags) side by side.

  python profiles/build_variants.py name1:-DPRIB_TC32=640 name2:-DFOO=1,-DBAR=2 ...

Each variant is priblast_b200/variants/libpriblast_acc_<name>.so (git-ignored, travels to the GPU box); select it
with PRIB_ACC_LIB=<path>.  Only acc_kernels.cu is recompiled; the other objects come from the normal build.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from priblast_b200 import build as b  # noqa: E402


def main():
    b.build_library()
    vdir = os.path.join(b.PKG, "variants")
    os.makedirs(vdir, exist_ok=True)
    nvcc = b.nvcc_path()
    others = [os.path.join(b.CSRC, os.path.splitext(s)[0] + ".o") for s in b.SOURCES if s != "acc_kernels.cu"]
    others.append(os.path.join(b.CSRC, "turner_blob.o"))

    def one(spec):
        name, _, flags = spec.partition(":")
        flags = [f for f in flags.split(",") if f]
        obj = os.path.join(vdir, f"acc_kernels_{name}.o")
        lib = os.path.join(vdir, f"libpriblast_acc_{name}.so")
        log = subprocess.run([nvcc, *b.NVCC_FLAGS, "-Xptxas", "-v", *flags, "-c", os.path.join(b.CSRC, "acc_kernels.cu"),
                              "-o", obj], check=True, capture_output=True, text=True).stderr
        subprocess.run([nvcc, *b.NVCC_FLAGS, "-shared", obj, *others, "-o", lib], check=True)
        keep = []
        lines = log.splitlines()
        for i, ln in enumerate(lines):
            if "Compiling entry function" in ln and ("tileIf" in ln or "biloop" in ln and "If" in ln):
                keep.append(ln.split("'")[1][-60:] + " | " + lines[i + 2].strip() + " | " + lines[i + 3].strip())
        return name, lib, keep

    with ThreadPoolExecutor(max_workers=4) as ex:
        for name, lib, keep in ex.map(one, sys.argv[1:]):
            print(name, lib)
            for k in keep:
                print("   ", k)


if __name__ == "__main__":
    main()
