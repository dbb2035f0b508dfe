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
