#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle/): extract the reference's OpenCL-C kernel string and make it g++-compilable.

What this does (nothing is copied into git; every output goes to oracle/_ref/, which is git-ignored):
  1. compile <reference>/core/cfd_core/FluidX3D/src/kernel.cpp *where it lies* together with dump_kernel.cpp
     (10 lines, ours) that prints get_opencl_c_code() (FX/kernel.hpp:6-17), the exact string the reference
     hands to the OpenCL JIT (FX/opencl.hpp:297-316);
  2. re-join the one-token-per-line text into statements (pure whitespace change);
  3. apply ONE mechanical rewrite so that a C++ compiler accepts OpenCL-C vector literals:
        (float3)(a,b,c)  ->  cl_make_float3(a,b,c)      (same for uint3/int3/float2/uchar4 ...)
     In C++ `(T)(a,b,c)` would parse as a cast of a comma expression, so this rewrite is unavoidable.
     No arithmetic, no identifier, no control flow is touched.
The result `oracle/_ref/kernel_cl.inc` is #included by ref_unit.cpp under clshim.hpp (our emulation of the
OpenCL-C built-ins the LBM part uses).  See oracle/README.md for how the result is used to pin the oracle.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")


def main(reference_root: str) -> int:
    fx = os.path.join(reference_root, "core", "cfd_core", "FluidX3D", "src")
    if not os.path.isfile(os.path.join(fx, "kernel.cpp")):
        print(f"reference kernel not found under {fx}", file=sys.stderr)
        return 2
    os.makedirs(OUT, exist_ok=True)
    dump = os.path.join(OUT, "dump_kernel")
    subprocess.check_call(["g++", "-std=c++17", "-O0", "-w", "-pthread", f"-I{fx}",
                           os.path.join(HERE, "dump_kernel.cpp"), os.path.join(fx, "kernel.cpp"), "-o", dump])
    text = subprocess.check_output([dump]).decode()
    os.remove(dump)

    lines, cur = [], []
    for tok in text.split("\n"):
        if tok.startswith("#"):
            if cur:
                lines.append(" ".join(cur)); cur = []
            lines.append(tok)
        else:
            cur.append(tok)
            if tok.endswith((";", "{", "}")):
                lines.append(" ".join(cur)); cur = []
    if cur:
        lines.append(" ".join(cur))
    joined = "\n".join(lines) + "\n"

    vec = r"(?:float|uint|int|uchar|ulong|ushort|char|short|long)(?:2|3|4)"
    rewritten, n = re.subn(r"\(\s*(" + vec + r")\s*\)\s*\(", r"cl_make_\1(", joined)
    with open(os.path.join(OUT, "kernel_cl.inc"), "w") as f:
        f.write("// GENERATED from the reference's get_opencl_c_code() by oracle/ref_shim/make_ref_kernel.py -- do not commit\n")
        f.write(rewritten)
    print(f"kernel_cl.inc: {len(lines)} statements, {n} vector-literal rewrites")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
