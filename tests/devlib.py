"""ctypes bindings of the cxb_* device layer (include/conex_b200_device.h) for the GPU tests.

torch is used only to own device memory: a column-major r x c matrix is a torch tensor of shape
(c, r), so `tensor.data_ptr()` is exactly the `double*` the C ABI expects.
"""
import ctypes as C
import os

import numpy as np

from harness import PRODUCT_SO, ConexLib

_dev = None
vp = C.c_void_p


def product():
    """The product library as a CONEX_* speaker plus its cxb_* kernels. Fails loudly if missing."""
    global _dev
    if _dev is None:
        if not os.path.exists(PRODUCT_SO):
            raise RuntimeError(f"{PRODUCT_SO} is missing: run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` — conex-b200 has no CPU fallback")
        _dev = ConexLib(PRODUCT_SO, "b200")
        L = _dev.lib
        L.cxb_dgemm.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, vp, C.c_long,
                                C.c_long, vp, C.c_long, C.c_long, C.c_double, vp, C.c_long, C.c_long,
                                C.c_int, C.c_int]
        L.cxb_dgemm_ex.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_double, vp, C.c_long, C.c_long, vp, C.c_long, C.c_long, C.c_double,
                                   vp, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int]
        L.cxb_set_default_gemm_config.argtypes = [C.c_int]
        L.cxb_set_default_gemm_config.restype = None
        L.cxb_schur_dense_lmi.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, C.c_long]
        L.cxb_schur_dense_lmi_streamed.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, C.c_long]
        L.CONEXB200_SetAssemblyMode.argtypes = [vp, C.c_int]
        L.CONEXB200_SetAssemblyMode.restype = None
        L.cxb_potrf_lower.argtypes = [vp, C.c_int, vp, C.c_long, vp, vp]
        L.cxb_potrs_lower.argtypes = [vp, C.c_int, vp, C.c_long, vp, C.c_long, C.c_int]
        L.cxb_gemv_n.argtypes = [vp, C.c_long, C.c_int, vp, vp, vp]
        L.cxb_lanczos_worksize.argtypes = [C.c_int]
        L.cxb_lanczos_worksize.restype = C.c_size_t
        L.cxb_lanczos_two_sided.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp]
        L.cxb_ws_reductions.argtypes = [vp, C.c_int, vp, vp]
        L.cxb_geodesic_worksize.argtypes = [C.c_int]
        L.cxb_geodesic_worksize.restype = C.c_size_t
        L.cxb_geodesic_update.argtypes = [vp, C.c_int, vp, vp, C.c_double, C.c_double, vp, vp, vp]
        L.cxb_pade_expm.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
        L.cxb_lu_solve.argtypes = [vp, C.c_int, vp, C.c_long, C.c_int, vp, C.c_long, vp, vp]
        L.cxb_set_identity.argtypes = [vp, C.c_int, vp]
        L.cxb_dot.argtypes = [vp, C.c_long, vp, vp, vp]
        L.cxb_axpbypcz.argtypes = [vp, C.c_long, C.c_double, vp, C.c_double, vp, C.c_double, vp]
        L.cxb_copy_strided.argtypes = [vp, C.c_long, vp, C.c_long, vp, C.c_long]
        L.cxb_scatter_add_lower.argtypes = [vp, C.c_int, vp, C.c_long, vp, vp, C.c_long]
        L.cxb_fill.argtypes = [vp, C.c_long, C.c_double, vp]
        L.cxb_scatter_add_vec.argtypes = [vp, C.c_int, vp, vp, vp]
        L.cxb_gather_vec.argtypes = [vp, C.c_int, vp, vp, vp]
        L.cxb_affine_update.argtypes = [vp, C.c_int, vp, vp, C.c_double]
        L.CONEXB200_AddDenseLMIConstraintDevice.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        L.CONEXB200_AddDenseLMIConstraintShard.argtypes = [vp, vp, C.c_int, C.c_int, vp]
        L.CONEXB200_NewDenseLMIConstraintStorage.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp)]
        L.CONEXB200_CommGetUniqueId.argtypes = [C.c_char_p]
        L.CONEXB200_CommInitRank.argtypes = [C.c_int, C.c_int, C.c_char_p]
        L.CONEXB200_CommDestroy.restype = None
        L.CONEXB200_ShardRange.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.CONEXB200_ShardRange.restype = None
        L.CONEXB200_GetIterationMilliseconds.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
        L.CONEXB200_SetTiming.argtypes = [vp, C.c_int]
        L.CONEXB200_GetIterationPhaseMilliseconds.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
        L.CONEXB200_DivergenceUpperBoundInverse.argtypes = [C.c_double] * 6
        L.CONEXB200_DivergenceUpperBoundInverse.restype = C.c_double
        L.CONEXB200_TridiagonalExtremes.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.CONEXB200_LaunchCount.restype = C.c_long
        L.CONEXB200_SetCollective.argtypes = [vp, C.c_int]
        L.CONEXB200_SetCollective.restype = None
        L.CONEXB200_SetDistributedCholesky.argtypes = [C.c_int, C.c_int]
        L.CONEXB200_SetDistributedCholesky.restype = None
        L.CONEXB200_DistributedPotrf.argtypes = [C.c_int, vp, C.c_long, C.c_int, C.POINTER(C.c_int)]
        L.cxb_potrf_begin.argtypes = [vp, vp]
        L.cxb_potrf_panel.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_long, vp]
    return _dev


def to_dev(a):
    """numpy (r x c or 1-D) -> torch cuda tensor whose memory is the column-major matrix."""
    import torch
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()


def from_dev(t, rows=None, cols=None):
    import torch
    torch.cuda.synchronize()
    h = t.cpu().numpy()
    if h.ndim == 1 and rows is None:
        return h.copy()
    if rows is not None:
        return h.reshape(-1)[: rows * cols].reshape((cols, rows)).T.copy()
    return h.T.copy()


def dzeros(*shape):
    import torch
    return torch.zeros(*shape, dtype=torch.float64, device="cuda")


def izeros(n):
    import torch
    return torch.zeros(n, dtype=torch.int32, device="cuda")


def ptr(t):
    return C.c_void_p(t.data_ptr())


def init_communicator(dev, rank, world):
    """Rendezvous of the library's NCCL communicator over an initialised torch.distributed group:
    rank 0 creates the unique id, torch broadcasts its 128 bytes, every rank joins."""
    import torch
    import torch.distributed as dist
    L = dev.lib
    buf = C.create_string_buffer(128)
    if rank == 0:
        assert L.CONEXB200_CommGetUniqueId(buf) == 0
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    ident = bytes(t.cpu().tolist())
    assert L.CONEXB200_CommInitRank(world, rank, ident) == 0


def shard_range(dev, m, world, rank):
    b, c = C.c_int(), C.c_int()
    dev.lib.CONEXB200_ShardRange(m, world, rank, C.byref(b), C.byref(c))
    return b.value, c.value
