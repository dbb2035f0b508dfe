"""N > 1 on real GPUs: the sharded Newton step (one process per GPU, NCCL) against the single-GPU
product and the oracle. Needs >= 2 visible GPUs (`gpurun --gpus 2 -- python -m pytest
tests/test_multi_gpu.py -m gpu`); skipped on a one-GPU box. The host logic of the same path is
covered on CPU by tests/test_sharding.py (gloo, world_size 2 and 3)."""
import os
import socket

import numpy as np
import pytest

from harness import maxcut_lmi, oracle, pack_matrices, random_dense_lmi

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(kind):
    if kind == "random":
        mats, Cm = random_dense_lmi(40, 23, 11)
        return mats, Cm, None
    mats, Cm, b = maxcut_lmi(48, 2)
    return mats, Cm, b


def _rank_main(rank, world, port, kind, out_dir, chol_block=0):
    import ctypes as C

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        import devlib
        dev = devlib.product()
        devlib.init_communicator(dev, rank, world)
        if chol_block:
            # force the multi-GPU Cholesky on these small systems (default: order >= 4096 only)
            dev.lib.CONEXB200_SetDistributedCholesky(0, chol_block)
        mats, Cm, b = _problem(kind)
        n, m = Cm.shape[0], len(mats)
        rb, rc = devlib.shard_range(dev, m, world, rank)
        A_local = torch.from_numpy(pack_matrices(mats[rb:rb + rc])).cuda()
        Cd = devlib.to_dev(Cm)
        P = dev.program()
        assert dev.lib.CONEXB200_AddDenseLMIConstraintShard(P.h, C.c_void_p(A_local.data_ptr()), n, m,
                                                            C.c_void_p(Cd.data_ptr())) == 0
        P.m = m
        P.cone_shapes.append((n, n))
        H, AW, AQc, sc = P.newton_system(coldstart=True)
        if b is None:
            b = P.feasible_objective()
        solved, y = P.maximize(b, dev.default_config(prepare_dual_variables=1))
        X = P.dual_variable(0)
        log = P.iteration_log()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), H=H, AW=AW, AQc=AQc, sc=sc, y=y, X=X, solved=solved,
                 by=[r["by"] for r in log], cx=[r["cx"] for r in log], k=[r["inv_sqrt_mu"] for r in log])
        dev.lib.CONEXB200_CommDestroy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,chol_block", [("random", 0), ("maxcut", 0), ("random", 8), ("maxcut", 16)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_solve_matches_single_gpu_and_oracle(tmp_path, world, kind, chol_block):
    """chol_block != 0: the Schur complement is also FACTORED across the ranks (block columns of
    chol_block columns, panel broadcasts) instead of replicated."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mp.spawn(_rank_main, args=(world, _free_port(), kind, str(tmp_path), chol_block), nprocs=world, join=True)
    ranks = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    # replicated state stays bit-identical on every rank
    for r in ranks[1:]:
        for key in ("H", "AW", "AQc", "y", "X", "k"):
            assert np.array_equal(ranks[0][key], r[key]), key
    # single-GPU product and oracle on the same bytes
    import devlib
    mats, Cm, b = _problem(kind)
    res = {}
    for name, L in (("oracle", oracle()), ("single", devlib.product())):
        P = L.program()
        P.add_dense_lmi(mats, Cm)
        H, AW, AQc, sc = P.newton_system(coldstart=True)
        bb = P.feasible_objective() if b is None else b
        solved, y = P.maximize(bb, L.default_config(prepare_dual_variables=1))
        res[name] = dict(H=H, AW=AW, AQc=AQc, sc=sc, y=y, solved=solved, log=P.iteration_log())
    got = ranks[0]
    for name in ("oracle", "single"):
        ref = res[name]
        scale = np.sqrt(np.outer(np.diag(ref["H"]), np.diag(ref["H"])))
        assert (np.abs(got["H"] - ref["H"]) / scale).max() < 1e-10, name      # BASELINE: 1e-10 relative
        assert np.abs(got["AW"] - ref["AW"]).max() <= 1e-10 * max(1.0, np.abs(ref["AW"]).max())
        assert np.abs(got["AQc"] - ref["AQc"]).max() <= 1e-10 * max(1.0, np.abs(ref["AQc"]).max())
        assert int(got["solved"]) == ref["solved"] == 1
        assert abs(len(got["by"]) - len(ref["log"])) <= 1                      # iteration count +-1
        assert abs(got["by"][-1] - ref["log"][-1]["by"]) <= 1e-7 * max(1.0, abs(ref["log"][-1]["by"]))
        assert np.abs(got["y"] - ref["y"]).max() <= 1e-6 * max(1.0, np.abs(ref["y"]).max())


def _spd(N, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((N, N + 8))
    return M @ M.T / N + np.eye(N)


def _potrf_rank(rank, world, port, cases, out_dir):
    import ctypes as C

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        import devlib
        dev = devlib.product()
        devlib.init_communicator(dev, rank, world)
        for idx, (N, block, spoil) in enumerate(cases):
            K = _spd(N, 100 + idx)
            if spoil >= 0:
                K[spoil, spoil] = -1.0                     # not positive definite from this pivot on
            ld = N + 2                                     # like the augmented Schur complement
            Hd = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
            Hd[:, :N] = torch.from_numpy(np.ascontiguousarray(np.tril(K).T)).cuda()   # column-major, lower
            info = C.c_int(-7)
            assert dev.lib.CONEXB200_DistributedPotrf(N, C.c_void_p(Hd.data_ptr()), ld, block, C.byref(info)) == 0
            torch.cuda.synchronize()
            np.savez(os.path.join(out_dir, f"potrf{idx}_rank{rank}.npz"),
                     L=np.tril(Hd[:, :N].cpu().numpy().T), info=info.value)
        dev.lib.CONEXB200_CommDestroy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_distributed_cholesky_matches_lapack_on_every_rank(tmp_path, world):
    """K3 across GPUs (host/distributed_cholesky.cc): every rank ends with the complete factor,
    bit-identical across ranks and within 1e-11 of LAPACK; a non-positive pivot is reported on every
    rank (block_triangular_operations.cc:193-196)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    cases = [(97, 8, -1), (700, 128, -1), (1500, 256, -1), (2200, 512, -1), (700, 128, 400)]
    mp.spawn(_potrf_rank, args=(world, _free_port(), cases, str(tmp_path)), nprocs=world, join=True)
    for idx, (N, block, spoil) in enumerate(cases):
        got = [np.load(tmp_path / f"potrf{idx}_rank{r}.npz") for r in range(world)]
        if spoil >= 0:
            assert all(int(g["info"]) != 0 for g in got)
            continue
        assert all(int(g["info"]) == 0 for g in got)
        for g in got[1:]:
            assert np.array_equal(got[0]["L"], g["L"])
        ref = np.linalg.cholesky(_spd(N, 100 + idx))
        assert np.abs(got[0]["L"] - ref).max() <= 1e-11 * np.abs(ref).max(), (N, block)
