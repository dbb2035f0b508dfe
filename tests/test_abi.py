"""The drop-in boundary on CPU: libconex_b200.so loads and exports every symbol include/*.h
declares, the ABI structs have the reference layout, argument validation follows the reference
(interfaces/test/interface_test.cc), and the host-side logic of the product (mu rule, Jacobi-matrix
extremes) agrees with the oracle. No compute call needs a GPU here.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from harness import PRODUCT_SO, ROOT, IterationStats, SolverConfiguration, dptr, oracle


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(PRODUCT_SO):
        import __graft_entry__
        __graft_entry__.build()
    return C.CDLL(PRODUCT_SO)


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:CONEX|CONEXB200|cxb)_\w+)\s*\(", text)))


@pytest.mark.parametrize("header", ["conex.h", "conex_b200.h", "conex_b200_device.h"])
def test_every_declared_symbol_is_exported(lib, header):
    names = declared_symbols(header)
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_reference_abi_has_all_entry_points():
    """The 21 functions of the reference header (interfaces/conex.h:41-99)."""
    expected = {
        "CONEX_CreateConeProgram", "CONEX_DeleteConeProgram", "CONEX_AddDenseLinearConstraint",
        "CONEX_AddLinearInequalities", "CONEX_AddQuadraticCost", "CONEX_AddDenseLMIConstraint",
        "CONEX_AddSparseLMIConstraint", "CONEX_Maximize", "CONEX_Solve", "CONEX_GetDualVariable",
        "CONEX_GetDualVariableSize", "CONEX_SetDefaultOptions", "CONEX_GetIterationStats",
        "CONEX_UpdateLinearOperator", "CONEX_NewLinearMatrixInequality", "CONEX_UpdateAffineTerm",
        "CONEX_NewLorentzConeConstraint", "CONEX_NewLinearInequality", "CONEX_NewQuadraticCost",
        "CONEX_UpdateQuadraticCostMatrix", "CONEX_SetNumberOfVariables"}
    assert set(declared_symbols("conex.h")) == expected


def test_struct_layouts_match_reference():
    # reference conex.h:10-35: 19 fields, natural alignment
    assert C.sizeof(SolverConfiguration) == 120
    assert SolverConfiguration.inv_sqrt_mu_max.offset == 8
    assert SolverConfiguration.enable_line_search.offset == 40
    assert SolverConfiguration.max_iterations.offset == 88
    assert SolverConfiguration.kkt_solver.offset == 116
    assert C.sizeof(IterationStats) == 16


def test_default_options(lib):
    lib.CONEX_SetDefaultOptions.argtypes = [C.POINTER(SolverConfiguration)]
    cfg = SolverConfiguration()
    C.memset(C.byref(cfg), 0xFF, C.sizeof(cfg))
    lib.CONEX_SetDefaultOptions(C.byref(cfg))
    # reference cone_program.h:17-38
    assert (cfg.prepare_dual_variables, cfg.initialization_mode) == (0, 0)
    assert cfg.inv_sqrt_mu_max == 1000 and cfg.minimum_mu == 1e-15 and cfg.maximum_mu == 1e4
    assert cfg.divergence_upper_bound == 1 and cfg.enable_line_search == 0 and cfg.dinf_upper_bound == 1
    assert cfg.final_centering_steps == 5 and cfg.final_centering_tolerance == .01
    assert cfg.warmstart_abort_threshold == 2 and cfg.max_iterations == 25
    assert cfg.infeasibility_threshold == 1e5 and cfg.kkt_error_tolerance == 1e10
    assert cfg.enable_rescaling == 1
    # the two fields the reference leaves uninitialised are zeroed here
    assert cfg.kkt_solver == 0 and cfg.iterative_refinement_iterations == 0
    lib.CONEX_SetDefaultOptions(None)  # null pointer: message, no crash (conex.cc:232-235)


def test_argument_validation_without_gpu(lib):
    """interfaces/test/interface_test.cc:5-120: failures are reported, never crash."""
    cid = C.c_int(-7)
    lib.CONEX_NewLinearMatrixInequality.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
    assert lib.CONEX_NewLinearMatrixInequality(None, 2, 2, C.byref(cid)) == 1      # null program
    dummy = C.c_void_p(1)
    assert lib.CONEX_NewLinearMatrixInequality(dummy, 2, 3, C.byref(cid)) == 1     # bad complex dim
    assert lib.CONEX_NewLinearMatrixInequality(dummy, 0, 2, C.byref(cid)) == 1     # bad order
    assert lib.CONEX_NewLinearMatrixInequality(dummy, 2, 2, None) == 1             # null output
    lib.CONEX_SetNumberOfVariables.argtypes = [C.c_void_p, C.c_int]
    assert lib.CONEX_SetNumberOfVariables(None, 4) == 1
    assert lib.CONEX_SetNumberOfVariables(dummy, 0) == 1
    lib.CONEX_NewLorentzConeConstraint.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    assert lib.CONEX_NewLorentzConeConstraint(dummy, 0, C.byref(cid)) == 1
    lib.CONEX_GetIterationStats.argtypes = [C.c_void_p, C.POINTER(IterationStats), C.c_int]
    lib.CONEX_GetIterationStats(None, None, 0)


def test_no_cpu_fallback_without_device(lib):
    """On a box without a GPU the library must refuse to create a program (no silent CPU path)."""
    lib.CONEXB200_DeviceAvailable.restype = C.c_int
    if lib.CONEXB200_DeviceAvailable():
        pytest.skip("a B200 is visible")
    lib.CONEX_CreateConeProgram.restype = C.c_void_p
    assert lib.CONEX_CreateConeProgram() is None


def test_mu_rule_matches_oracle(lib):
    """Host logic: DivergenceUpperBoundInverse of the product vs the oracle (bit-for-bit paths of
    divergence.cc are scalar closed forms, so 1e-12 relative)."""
    O = oracle().lib
    f = lib.CONEXB200_DivergenceUpperBoundInverse
    f.argtypes = [C.c_double] * 6
    f.restype = C.c_double
    rng = np.random.default_rng(0)
    for _ in range(500):
        n = int(rng.integers(1, 30))
        lam = rng.uniform(-0.5, 4.0, n)
        args = (float(rng.uniform(0.1, 4.0)) * n, float((lam ** 2).sum()), float(lam.sum()),
                float(lam.min()), float(lam.max()), float(n))
        a, b = f(*args), O.ORACLE_DivergenceUpperBoundInverse(*args)
        if np.isnan(a) or np.isnan(b):
            assert np.isnan(a) and np.isnan(b)
        else:
            assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), args


def test_tridiagonal_extremes_match_oracle(lib):
    O = oracle().lib
    g = lib.CONEXB200_TridiagonalExtremes
    g.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 7, 50, 400, 1000):
        a = rng.standard_normal(n)
        b = np.abs(rng.standard_normal(max(n - 1, 1))) * rng.choice([1.0, 1e-3, 1e-9])
        out = np.zeros(2)
        g(n, dptr(a), dptr(b), dptr(out))
        ev = np.zeros(n)
        O.ORACLE_TridiagonalEigenvalues(n, dptr(a), dptr(b), dptr(ev))
        scale = max(1.0, np.abs(ev).max())
        assert abs(out[0] - ev[0]) < 1e-13 * scale and abs(out[1] - ev[-1]) < 1e-13 * scale


def test_python_front_end_imports():
    import conex_b200
    assert os.path.exists(conex_b200.LIBRARY_PATH)
    assert hasattr(conex_b200.Conex, "AddDenseLinearMatrixInequality")
    assert hasattr(conex_b200.Conex, "Maximize")


def test_plain_c_caller_compiles_links_and_runs(tmp_path):
    """include/conex.h is plain C (MATLAB loadlibrary, SWIG) and every symbol a C caller uses resolves
    against libconex_b200.so; the program solves when a device is present and is refused cleanly when
    none is (reference: interfaces/test/test_app.cc, interfaces/Makefile:51-52)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "conex_b200", "lib")
    exe = str(tmp_path / "c_caller")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "capp", "c_caller.c"), "-o", exe, "-L", libdir,
                           "-lconex_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "solved" in out.stdout or "no device" in out.stdout
