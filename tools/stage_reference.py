#!/usr/bin/env python
"""Stage the UNMODIFIED reference package next to the repository so that it travels to the GPU box.

    python tools/stage_reference.py

Copies /root/reference/pygraphblas (the pure-Python package) and /root/reference/tests into the git-ignored
baseline/_ref/ (never into the history: reference sources are not product source).  `gpurun` snapshots carry
baseline/_ref/ along, so on a B200 the reference's own Matrix.mxm / Matrix.mxv / Vector.vxm run, unchanged, over
suitesparse_graphblas/ (the binding stub) -> libb200grb.so: tests/test_reference_gpu.py and bench.py's e2e leg.
Called by __graft_entry__.build() whenever /root/reference is present.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(verbose=True):
    if not os.path.isdir(os.path.join(REF, "pygraphblas")):
        if verbose:
            print(f"stage_reference: {REF} not present, keeping {DST} as it is")
        return os.path.isdir(os.path.join(DST, "pygraphblas"))
    os.makedirs(DST, exist_ok=True)
    for sub in ("pygraphblas", "tests"):
        dst = os.path.join(DST, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copy of /root/reference/{pygraphblas,tests} (Graphegon/pygraphblas), staged by tools/stage_reference.py.\n"
                "Git-ignored; present only so that the reference's own code can run on the GPU box over libb200grb.so.\n")
    if verbose:
        print(f"stage_reference: staged the reference package and its tests into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
