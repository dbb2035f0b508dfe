"""Multi-GPU host logic on CPU: the 1-D row partition and the symmetric block-pair plan of the
sharded Schur assembly (conex_b200/csrc/host/communicator.{h,cc}), and — with two gloo ranks — the
whole sharded assembly algorithm (local diagonal block, peer matrices exchanged point to point,
off-diagonal blocks contracted into place, one all-reduce) restated in numpy around the product's own
plan functions and checked against the oracle's single-process Newton system. No GPU is touched:
CONEXB200_ShardRange / CONEXB200_ShardPlan are pure host code."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

from harness import PRODUCT_SO, oracle, random_dense_lmi


def product_lib():
    L = C.CDLL(PRODUCT_SO)
    L.CONEXB200_ShardRange.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.CONEXB200_ShardRange.restype = None
    L.CONEXB200_ShardPlan.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    return L


def shard_range(L, m, world, rank):
    b, c = C.c_int(), C.c_int()
    L.CONEXB200_ShardRange(m, world, rank, C.byref(b), C.byref(c))
    return b.value, c.value


def shard_plan(L, m, world, rank):
    buf = (C.c_int * (5 * 16))()
    k = L.CONEXB200_ShardPlan(m, world, rank, buf, 16)
    assert k <= 16
    return [tuple(buf[5 * i:5 * i + 5]) for i in range(k)]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("m", [8, 37, 100, 2000])
def test_plan_covers_lower_triangle_exactly_once(world, m):
    L = product_lib()
    cover = np.zeros((m, m), dtype=np.int32)
    ranges = [shard_range(L, m, world, r) for r in range(world)]
    assert ranges[0][0] == 0 and sum(c for _, c in ranges) == m
    for (b0, c0), (b1, _) in zip(ranges[:-1], ranges[1:]):
        assert b0 + c0 == b1
    work = []
    for r in range(world):
        rb, rc = ranges[r]
        idx = np.arange(rb, rb + rc)
        cover[np.ix_(idx, idx)] += np.tril(np.ones((rc, rc), dtype=np.int32))
        w = rc * (rc + 1) // 2
        for peer, row_b, row_c, col_b, col_c in shard_plan(L, m, world, r):
            assert peer != r and row_c > 0 and col_c > 0
            assert rb <= row_b and row_b + row_c <= rb + rc                     # own rows
            pb, pc = ranges[peer]
            assert pb <= col_b and col_b + col_c <= pb + pc                     # peer's matrices
            rows, cols = np.arange(row_b, row_b + row_c), np.arange(col_b, col_b + col_c)
            if peer < r:
                cover[np.ix_(rows, cols)] += 1
            else:
                cover[np.ix_(cols, rows)] += 1                                  # stored transposed
            w += row_c * col_c
        work.append(w)
    assert np.array_equal(cover, np.tril(np.ones((m, m), dtype=np.int32)))
    if m >= 100:
        assert max(work) <= 1.1 * min(work) + 2 * m, work                          # balanced


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_rank(rank, world, port, n, m, seed, out):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        L = product_lib()
        mats, Cm = random_dense_lmi(n, m, seed)
        rng = np.random.default_rng(seed + 1)
        Wh = rng.standard_normal((n, n))
        W = Wh @ Wh.T / n + np.eye(n)
        rb, rc = shard_range(L, m, world, rank)
        local = mats[rb:rb + rc]                      # the only constraint matrices this rank holds
        B = [W @ A @ W for A in local]                # K1 on the local shard
        H = np.zeros((m + 2, m + 1))
        for a in range(rc):                           # local diagonal block + residual rows
            for b in range(a + 1):
                H[rb + a, rb + b] = np.sum(B[a] * local[b])
            H[m, rb + a] = np.sum((W @ Cm @ W) * local[a])     # AQc
            H[m + 1, rb + a] = np.sum(W * local[a])            # AW
        if rank == 0:
            H[m, m] = np.sum((W @ Cm @ W) * Cm)
            H[m + 1, m] = np.sum(W * Cm)
        # exchange at every cyclic distance: send what the receiver's task needs, receive ours
        plan = {t[0]: t for t in shard_plan(L, m, world, rank)}
        for d in range(1, world // 2 + 1):
            to = (rank - d) % world
            frm = (rank + d) % world
            theirs = {t[0]: t for t in shard_plan(L, m, world, to)}.get(rank)
            reqs = []
            if theirs is not None:
                _, _, _, cb, cc = theirs
                send = torch.from_numpy(np.stack(mats[cb:cb + cc]).copy())
                assert rb <= cb and cb + cc <= rb + rc
                reqs.append(dist.isend(send, to))
            mine = plan.get(frm)
            if mine is not None:
                _, row_b, row_c, col_b, col_c = mine
                recv = torch.empty((col_c, n, n), dtype=torch.float64)
                reqs.append(dist.irecv(recv, frm))
            for q in reqs:
                q.wait()
            if mine is not None:
                peer_mats = recv.numpy()
                for i in range(row_b, row_b + row_c):
                    for j in range(col_c):
                        v = np.sum(B[i - rb] * peer_mats[j])
                        if frm < rank:
                            H[i, col_b + j] = v
                        else:
                            H[col_b + j, i] = v
        Ht = torch.from_numpy(H)
        dist.all_reduce(Ht)
        if rank == 0:
            np.save(out, Ht.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_assembly_over_gloo_matches_oracle(tmp_path, world):
    import torch.multiprocessing as mp
    n, m, seed = 6, 11, 3
    out = str(tmp_path / "H.npy")
    mp.spawn(_gloo_rank, args=(world, _free_port(), n, m, seed, out), nprocs=world, join=True)
    H = np.load(out)
    # oracle: the reference's assembly on one process with every matrix (dense_lmi_constraint.cc:62-103)
    from harness import dptr, fmat, pack_matrices
    O = oracle()
    mats, Cm = random_dense_lmi(n, m, seed)
    rng = np.random.default_rng(seed + 1)
    Wh = rng.standard_normal((n, n))
    W = Wh @ Wh.T / n + np.eye(n)
    A = pack_matrices(mats)
    G = np.zeros((m, m), order="F")
    AW, AQc, sc = np.zeros(m), np.zeros(m), np.zeros(2)
    O.lib.ORACLE_SchurDenseLMI(n, m, dptr(A), dptr(fmat(Cm)), dptr(fmat(W)), 0, dptr(G), dptr(AW), dptr(AQc),
                               dptr(sc))
    scale = np.sqrt(np.outer(np.diag(G), np.diag(G)))
    assert (np.abs(np.tril(H[:m, :m]) - np.tril(G)) / scale).max() < 1e-12
    assert np.allclose(H[m, :m], AQc, rtol=1e-12, atol=1e-12)
    assert np.allclose(H[m + 1, :m], AW, rtol=1e-12, atol=1e-12)
    assert np.allclose([H[m + 1, m], H[m, m]], sc, rtol=1e-12)


# ---- distributed Cholesky (conex_b200/csrc/host/distributed_cholesky.{h,cc}) --------------------------
FACTOR, BROADCAST, WAIT, UPDATE = 0, 1, 2, 3


def cholesky_schedule(L, N, block, world, rank):
    L.CONEXB200_CholeskySchedule.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_int), C.c_int]
    k = L.CONEXB200_CholeskySchedule(N, block, world, rank, None, 0)
    buf = (C.c_int * (3 * max(k, 1)))()
    assert L.CONEXB200_CholeskySchedule(N, block, world, rank, buf, k) == k
    return [tuple(buf[3 * i:3 * i + 3]) for i in range(k)]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("N,block", [(1, 4), (7, 8), (64, 8), (100, 16), (20000, 512)])
def test_cholesky_schedule_is_a_valid_right_looking_order(world, N, block):
    """Every panel is factored exactly once, by its owner, after its block column has received the
    update of every earlier panel (each of which had arrived before it was used); every rank issues
    the same broadcasts in the same order; every block column is updated only by its owner."""
    L = product_lib()
    nblk = (N + block - 1) // block
    bcasts = None
    factored_by = {}
    for rank in range(world):
        ops = cholesky_schedule(L, N, block, world, rank)
        arrived, updates = set(), {}
        mine = []
        for kind, panel, target in ops:
            if kind == BROADCAST:
                assert target == panel % world
                if target == rank:
                    assert factored_by.get(panel) == rank                    # factored before it is sent
                mine.append((panel, target))
            elif kind == WAIT:
                arrived.add(panel)
            elif kind == UPDATE:
                assert target % world == rank and target > panel
                assert panel in arrived
                assert updates.setdefault(target, []) == list(range(panel))  # panels applied in order
                updates[target].append(panel)
            elif kind == FACTOR:
                assert panel % world == rank and panel not in factored_by
                assert updates.get(panel, []) == list(range(panel))          # fully updated
                factored_by[panel] = rank
        assert arrived == set(range(nblk))
        if bcasts is None:
            bcasts = mine
        assert mine == bcasts == [(J, J % world) for J in range(nblk)]
    assert sorted(factored_by) == list(range(nblk))


def _gloo_cholesky_rank(rank, world, port, N, block, out):
    import scipy.linalg as sla
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        L = product_lib()
        rng = np.random.default_rng(5)
        M = rng.standard_normal((N, N))
        H = np.tril(M @ M.T + N * np.eye(N))          # replicated input, lower triangle only
        if rank != 0:
            # block columns a rank does not own may hold anything before their panel arrives
            for J in range((N + block - 1) // block):
                if J % world != rank:
                    H[:, J * block:(J + 1) * block] = np.nan
        for kind, panel, target in cholesky_schedule(L, N, block, world, rank):
            j0, j1 = panel * block, min(N, (panel + 1) * block)
            if kind == FACTOR:
                H[j0:j1, j0:j1] = np.linalg.cholesky(np.tril(H[j0:j1, j0:j1]) + np.tril(H[j0:j1, j0:j1], -1).T)
                if j1 < N:
                    H[j1:, j0:j1] = sla.solve_triangular(H[j0:j1, j0:j1], H[j1:, j0:j1].T, lower=True).T
            elif kind == BROADCAST:
                buf = torch.from_numpy(np.ascontiguousarray(H[j0:, j0:j1]))
                dist.broadcast(buf, src=target)
                H[j0:, j0:j1] = buf.numpy()
            elif kind == UPDATE:
                k0, k1 = target * block, min(N, (target + 1) * block)
                upd = H[k0:, j0:j1] @ H[k0:k1, j0:j1].T
                blockc = H[k0:, k0:k1]
                blockc -= np.where(np.arange(k0, N)[:, None] >= np.arange(k0, k1)[None, :], upd, 0.0)
        np.save(out + f".{rank}.npy", H)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_cholesky_schedule_over_gloo(tmp_path, world):
    """The product's own schedule executed with numpy blocks and gloo broadcasts: every rank ends
    with the complete factor of the replicated matrix."""
    import torch.multiprocessing as mp
    N, block = 45, 8
    out = str(tmp_path / "L")
    mp.spawn(_gloo_cholesky_rank, args=(world, _free_port(), N, block, out), nprocs=world, join=True)
    rng = np.random.default_rng(5)
    M = rng.standard_normal((N, N))
    ref = np.linalg.cholesky(M @ M.T + N * np.eye(N))
    for r in range(world):
        Lr = np.tril(np.load(out + f".{r}.npy"))
        assert np.abs(Lr - ref).max() < 1e-11
