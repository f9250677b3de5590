"""Static checks on the SASS of the built library (no GPU needed: cuobjdump reads the cubin in libb200grb.so).

1. The TMA-staged SpMV kernel (spmv_run.cuh) must keep its WAR guard: in every instantiation, between the mbarrier wait of a run
   (SYNCS.PHASECHK...TRYWAIT) and the bulk copy of the NEXT run into the same stage (UBLKCP) there is a warp vote that reads the
   words just loaded from the stage -- without it the copy can overtake shared-memory loads that were only issued.
2. The library is sm_100a code using the bulk-copy / mbarrier instructions DESIGN.md claims (UBLKCP, SYNCS.ARRIVE.TRANS64)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pygraphblas_b200", "libb200grb.so")


@pytest.fixture(scope="module")
def hot2_sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name and "spmv_run_hot2_kernel" in name:
            m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                funcs.setdefault(name, []).append(m.group(1))
    if not funcs:
        pytest.skip("no SASS found in the library")
    return funcs


def test_every_hot2_instantiation_keeps_the_stage_guard(hot2_sass):
    assert len(hot2_sass) >= 20                       # FP32 / FP64 / integer / BOOL semirings x {plain, pipelined}
    for name, ops in hot2_sass.items():
        waits = [i for i, o in enumerate(ops) if o.startswith("SYNCS.PHASECHK")]
        copies = [i for i, o in enumerate(ops) if o.startswith("UBLKCP")]
        votes = [i for i, o in enumerate(ops) if o.startswith("VOTE.")]
        assert waits and copies, name
        guarded = 0
        for u in copies:
            before = [w for w in waits if w < u]
            if not before:
                continue                              # prologue: the table and the first run, nothing has been read yet
            w = before[-1]
            assert any(w < v < u for v in votes), f"{name}: bulk copy at instruction {u} follows the wait at {w} without the vote"
            guarded += 1
        assert guarded >= 1, name


def test_the_kernel_uses_bulk_copies_and_transaction_barriers(hot2_sass):
    for name, ops in hot2_sass.items():
        assert any(o.startswith("UBLKCP") for o in ops) and any(o.startswith("SYNCS.ARRIVE.TRANS64") for o in ops), name
