#!/usr/bin/env python
"""Kernel experiments: rebuild ONE translation unit with extra -D flags and link a variant library next to the shipped one.

    python tools/build_variant.py <name> <source.cu> [-DFLAG ...]      -> diffrax_b200/lib/variants/libdiffrax_b200_<name>.so
    DFX_LIB=diffrax_b200/lib/variants/libdiffrax_b200_<name>.so python bench.py ...

The variant reuses every other object of the normal build (run `python -m diffrax_b200.build` first)."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffrax_b200 import build as B  # noqa: E402


def main():
    name, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
    B.build(verbose=False)
    vdir = os.path.join(B.LIBDIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    src = os.path.join(B.CSRC, os.path.basename(src))
    obj = os.path.join(vdir, f"{os.path.basename(src)[:-3]}_{name}.o")
    r = subprocess.run([B.NVCC, *B.ARCH, *B.FLAGS, *flags, "-c", src, "-o", obj], capture_output=True, text=True)
    open(obj + ".log", "w").write(r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[-4000:])
        raise SystemExit(1)
    objs = [o for o in sorted(glob.glob(os.path.join(B.OBJ, "*.o"))) if os.path.basename(o) != os.path.basename(src)[:-3] + ".o"]
    lib = os.path.join(vdir, f"libdiffrax_b200_{name}.so")
    r = subprocess.run([B.NVCC, *B.ARCH, "-shared", "-o", lib, *objs, obj, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr)
        raise SystemExit(1)
    print(lib)


if __name__ == "__main__":
    main()
