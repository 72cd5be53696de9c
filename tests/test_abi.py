"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/neompc.h declares,
its records have the sizes the numpy mirrors assume, and it refuses to run without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from neo_mpc_planner2_b200 import _lib, abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "neompc.h")).read()
    declared = set(re.findall(r"\b(neompc_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_header_is_plain_c():
    """include/neompc.h is the drop-in boundary: it must compile as C99 (cgo / ctypes / plain C callers) and as C++."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "neompc.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])
    # and a C caller links against the library using nothing but the header
    src = os.path.join(ROOT, "tests", "c_caller.c")
    exe = os.path.join(ROOT, "tests", "hostsim", "_build", "c_caller")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                           "-L", libdir, "-lneompc", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "abi ok" in out.stdout


@pytest.mark.gpu
def test_c_caller_solves_on_gpu():
    """The same plain-C program on a box with a GPU: creates a handle, solves the known-answer problem, rolls the path."""
    import subprocess
    test_header_is_plain_c()                       # builds tests/hostsim/_build/c_caller
    out = subprocess.run([os.path.join(ROOT, "tests", "hostsim", "_build", "c_caller")], capture_output=True, text=True)
    assert out.returncode == 0 and "abi ok (twist 0.0833" in out.stdout, out.stdout + out.stderr


def test_record_sizes(lib):
    sz = (ctypes.c_size_t * 7)()
    assert lib.neompc_abi_sizes(sz) == 0
    assert list(sz) == [abi.REQUEST_DTYPE.itemsize, abi.RESPONSE_DTYPE.itemsize, abi.PARAMS_DTYPE.itemsize,
                        abi.MSG_DTYPE.itemsize, abi.TICK_DTYPE.itemsize, abi.CARROT_INFO_DTYPE.itemsize,
                        abi.PLAN_POSE_DTYPE.itemsize] \
        == [64, 32, 128, 240, 48, 16, 32]
    assert lib.neompc_version() == 100


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from neo_mpc_planner2_b200.solver import BatchSolver, NeompcError
    with pytest.raises(NeompcError, match="no CPU fallback"):
        BatchSolver(abi.README_SAMPLE)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "neo_mpc_planner2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "scipy.optimize" not in src, f


def test_params_record_keeps_solver_knobs():
    """A record rebuilt from a record (what BatchSolver.set_params(**changes) does) keeps every field that is not named —
    the solver knobs included (round-1 advisor finding: they were reset to 0)."""
    rec = abi.params_record(abi.README_SAMPLE, footprint_mode=abi.FOOTPRINT_MOVING, costmap_mode=abi.COSTMAP_BILINEAR,
                            lbfgs_memory=3, control_smoothing=0.005, max_iterations=50, lanes_per_instance=8,
                            costmap_guidance=abi.GUIDANCE_OFF)
    again = abi.params_record(rec, w_trans=0.3)
    assert again["w_trans"] == np.float32(0.3)
    for k in abi.KNOB_NAMES:
        assert again[k] == rec[k], k
    for k in abi.REFERENCE_PARAM_NAMES:
        if k != "w_trans":
            assert again[k] == rec[k], k
    with pytest.raises(KeyError):
        abi.params_record(abi.README_SAMPLE, w_tranz=1.0)
