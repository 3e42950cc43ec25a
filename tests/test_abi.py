"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls) -- CPU only."""
import ctypes
import os
import re

import pytest

from arap_flow_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"^ARAPB200_API [^;(]*?\b((?:Opt_|arapb200_)\w+)\(", txt, flags=re.M)))


def test_library_exports_every_declared_symbol():
    L = lib.load()
    opt, arap = _declared("Opt.h"), _declared("arapb200.h")
    assert len(opt) == 10 and sorted(opt) == sorted(lib.OPT_SYMBOLS)  # the 10 entry points of Opt.h:34-70
    assert sorted(arap) == sorted(lib.ARAP_SYMBOLS)
    for s in opt + arap:
        assert getattr(L, s) is not None, s


def test_opt_struct_layout_matches_reference_abi():
    # Opt.h:10-30: four ints, passed by value
    assert ctypes.sizeof(lib.OptInitializationParameters) == 16
    assert [f[0] for f in lib.OptInitializationParameters._fields_] == [
        "doublePrecision", "verbosityLevel", "collectPerKernelTimingInfo", "threadsPerBlock"]


def test_problem_define_refuses_foreign_plans(tmp_path):
    """Opt_NewState / Opt_ProblemDefine are host-only: exercise the refusal paths without a GPU."""
    L = lib.load()
    st = L.Opt_NewState(lib.OptInitializationParameters(0, 0, 0, 0))
    assert st
    assert not L.Opt_NewState(lib.OptInitializationParameters(1, 0, 0, 0))  # doublePrecision unsupported
    bad = tmp_path / "other.t"
    bad.write_text("local X = Unknown('X', float, {W,H}, 0)\n")
    assert not L.Opt_ProblemDefine(st, str(bad).encode(), b"gaussNewtonGPU")
    assert not L.Opt_ProblemDefine(st, b"/nonexistent/arap_plan.t", b"gaussNewtonGPU")
    good = os.path.join(ROOT, "arap_flow_b200", "arap_plan.t")
    assert not L.Opt_ProblemDefine(st, good.encode(), b"lbfgsGPU")      # o.t:121-124: gaussNewtonGPU or LMGPU
    lm = L.Opt_ProblemDefine(st, good.encode(), b"LMGPU")
    assert lm
    L.Opt_ProblemDelete(st, lm)
    p = L.Opt_ProblemDefine(st, good.encode(), b"gaussNewtonGPU")
    assert p
    L.Opt_ProblemDelete(st, p)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under arap_flow_b200/ (Python or native sources) may import, link or
    execute it, and the shared library must not depend on liboracle."""
    import re
    import subprocess
    pkg = os.path.join(ROOT, "arap_flow_b200")
    offenders = []
    for dp, dn, fn in os.walk(pkg):
        if os.path.basename(dp) in ("build", "bin", "__pycache__"):
            continue
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|pyoracle|oracle/_ref", txt, re.M):
                    offenders.append(os.path.join(dp, f))
    assert not offenders, offenders
    needed = subprocess.run(["readelf", "-d", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed


def test_no_cpu_fallback_without_a_gpu():
    """Without a CUDA device every compute entry point fails loudly (non-zero code / exit), it never computes on the host."""
    import subprocess
    import sys
    probe = subprocess.run([sys.executable, "-c", "import torch,sys; sys.exit(0 if torch.cuda.is_available() else 3)"],
                           capture_output=True)
    if probe.returncode == 0:
        pytest.skip("a GPU is present")
    code = ("import numpy as np\n"
            "from arap_flow_b200 import lib, synth\n"
            "sp = synth.synth(32, 32, 1, 1, 0)\n"
            "lib.deform(sp.rgb, sp.masks[0], sp.matches, nCont=1, nGN=1, nPCG=1)\n"
            "print('COMPUTED')\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode != 0 and "COMPUTED" not in r.stdout
    assert "CUDA" in (r.stderr + r.stdout) or "failed" in (r.stderr + r.stdout)
