"""Loader of the CPU oracle (test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this). The oracle speaks the same CONEX_* C ABI as the
product, so it is driven through the very same ctypes binding (conex_b200.binding.ConexLib)."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from conex_b200.binding import ConexLib, c_double_p, c_int_p  # noqa: E402

ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libconex_oracle.so")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return ORACLE_SO


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        _oracle = ConexLib(ORACLE_SO, "oracle")
        L = _oracle.lib
        L.ORACLE_SchurDenseLMI.argtypes = [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                           C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]
        L.ORACLE_NegativeSlack.argtypes = [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                           C.c_double, c_double_p]
        L.ORACLE_PsdStep.argtypes = [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p,
                                     C.c_double, C.c_double, C.c_int, C.c_int, c_double_p, c_double_p]
        L.ORACLE_PsdWeightedSlackEigenvalues.argtypes = [C.c_int, C.c_int, c_double_p, c_double_p,
                                                         c_double_p, c_double_p, C.c_double, c_double_p]
        L.ORACLE_PadeExpm.argtypes = [C.c_int, c_double_p, c_double_p]
        L.ORACLE_ApproximateEigenvalues.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p,
                                                    C.c_int, c_double_p]
        L.ORACLE_SymmetricLanczos.argtypes = [C.c_int, c_double_p, c_double_p, C.c_int, c_double_p]
        L.ORACLE_SymmetricEigenvalues.argtypes = [C.c_int, c_double_p, c_double_p]
        L.ORACLE_TridiagonalEigenvalues.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p]
        L.ORACLE_CholeskyLower.argtypes = [C.c_int, c_double_p]
        L.ORACLE_SolveLower.argtypes = [C.c_int, c_double_p, c_double_p, C.c_int]
        L.ORACLE_SolveSpd.argtypes = [C.c_int, c_double_p, c_double_p]
        L.ORACLE_LuSolve.argtypes = [C.c_int, c_double_p, C.c_int, c_double_p]
        L.ORACLE_Gemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, c_double_p,
                                  C.c_int, c_double_p, C.c_int, C.c_double, c_double_p, C.c_int]
        L.ORACLE_DivergenceUpperBoundInverse.argtypes = [C.c_double] * 6
        L.ORACLE_DivergenceUpperBoundInverse.restype = C.c_double
        L.ORACLE_DivergenceUpperBound.argtypes = [C.c_double] * 6
        L.ORACLE_DivergenceUpperBound.restype = C.c_double
        L.ORACLE_SetGramVariant.argtypes = [C.c_void_p, C.c_int]
        L.ORACLE_ForcePlainLoops.argtypes = [C.c_int]
        L.ORACLE_SetBlasThreads.argtypes = [C.c_int]
        L.ORACLE_LdltLower.argtypes = [C.c_int, c_double_p, c_int_p]
        L.ORACLE_SolveLdlt.argtypes = [C.c_int, c_double_p, c_int_p, c_double_p]
    return _oracle
