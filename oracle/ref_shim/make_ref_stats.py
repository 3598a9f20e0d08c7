#!/usr/bin/env python3
"""TEST INFRASTRUCTURE (oracle/): extract the reference's own averaging loop -- the lambda `accumulate_from_buffers` of FX/setup.cpp:4441-4488 --
verbatim from the source where it lies and write it to oracle/_ref/stats_lambda.inc (git-ignored). ref_stats.cpp supplies the few names the text
refers to (lbm.u.x/.y/.z, lbm.rho, avg_u, ..., parallel_for) and exports it as one C function, so that the C restatement (luwo_stats_accumulate) and
the CUDA kernel are pinned against the reference text itself. Nothing is copied into git."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")


def main(reference_root):
    src = os.path.join(reference_root, "core", "cfd_core", "FluidX3D", "src", "setup.cpp")
    lines = open(src, encoding="utf-8", errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if "auto accumulate_from_buffers = [&]() {" in l)
    indent = len(lines[start]) - len(lines[start].lstrip())
    end = next(i for i in range(start + 1, len(lines)) if lines[i].rstrip() == " " * indent + "};")
    body = [l for l in lines[start + 1:end]]
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "stats_lambda.inc"), "w") as fh:
        fh.write("// extracted verbatim from FX/setup.cpp:%d-%d by make_ref_stats.py\n" % (start + 2, end))
        fh.write("\n".join(body) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
