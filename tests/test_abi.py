"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls) -- CPU only."""
import ctypes
import os
import re

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
    assert not L.Opt_ProblemDefine(st, good.encode(), b"LMGPU")
    p = L.Opt_ProblemDefine(st, good.encode(), b"gaussNewtonGPU")
    assert p
    L.Opt_ProblemDelete(st, p)
