"""Parity of the small-cone arithmetic (LP cone, second-order cone, small dense LMI block, small
KKT Cholesky) with the CPU oracle, cone function by cone function, on batches of seeded problems.

The same checks run twice: on the CPU through the host stand-in of the CTA (tests/emul, checks the
header's index arithmetic where no GPU exists) and, marked gpu, through the real cxb_small_*
kernels. Tolerances: 1e-10 relative on Newton-system entries (BASELINE.json), 1e-9 on the
eigenvalue / norm outputs that pass through Lanczos.
"""
import ctypes as C

import numpy as np
import pytest

from harness import dptr, oracle
from small_backend import LP, PSD, SOC, Backend, rows_of

c_double_p = C.POINTER(C.c_double)


@pytest.fixture(scope="module", params=["emul", pytest.param("device", marks=pytest.mark.gpu)])
def be(request):
    return Backend(request.param)


def oracle_cone_api():
    O = oracle().lib
    O.ORACLE_ConeCreate.restype = C.c_void_p
    O.ORACLE_ConeCreate.argtypes = [C.c_int, C.c_int, C.c_int, c_double_p]
    O.ORACLE_ConeDelete.argtypes = [C.c_void_p]
    O.ORACLE_ConeStateSize.argtypes = [C.c_void_p]
    O.ORACLE_ConeGetState.argtypes = [C.c_void_p, c_double_p]
    O.ORACLE_ConeSetState.argtypes = [C.c_void_p, c_double_p]
    O.ORACLE_ConeSchur.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]
    O.ORACLE_ConeEigen.argtypes = [C.c_void_p, c_double_p, C.c_double, c_double_p]
    O.ORACLE_ConePrepare.argtypes = [C.c_void_p, c_double_p, C.c_int, C.c_double, C.c_double, c_double_p]
    O.ORACLE_ConeTakeStep.argtypes = [C.c_void_p, C.c_double, C.c_double]
    return O


class OracleCone:
    def __init__(self, kind, n, m, data):
        self.O = oracle_cone_api()
        self.m = m
        self.data = np.ascontiguousarray(data)
        self.h = C.c_void_p(self.O.ORACLE_ConeCreate(kind, n, m, dptr(self.data)))
        self.sz = self.O.ORACLE_ConeStateSize(self.h)

    def __del__(self):
        self.O.ORACLE_ConeDelete(self.h)

    def state(self):
        w = np.zeros(self.sz)
        self.O.ORACLE_ConeGetState(self.h, dptr(w))
        return w

    def schur(self):
        m = self.m
        G, AW, AQc, sc = np.zeros((m, m), order="F"), np.zeros(m), np.zeros(m), np.zeros(2)
        self.O.ORACLE_ConeSchur(self.h, dptr(G), dptr(AW), dptr(AQc), dptr(sc))
        return np.tril(G), AW, AQc, sc

    def eigen(self, y, cw):
        out = np.zeros(4)
        self.O.ORACLE_ConeEigen(self.h, dptr(np.ascontiguousarray(y)), cw, dptr(out))
        return out

    def prepare(self, y, cw, ew=1.0, affine=False):
        out = np.zeros(2)
        self.O.ORACLE_ConePrepare(self.h, dptr(np.ascontiguousarray(y)), int(affine), cw, ew, dptr(out))
        return out

    def take_step(self, step, ew=1.0):
        self.O.ORACLE_ConeTakeStep(self.h, step, ew)


def random_cone_data(kind, n, m, rng):
    """rows x (m + 1) column-major block: operator columns then a strictly feasible affine term."""
    rows = rows_of(kind, n)
    cols = []
    if kind == PSD:
        for _ in range(m):
            R = rng.uniform(-1, 1, size=(n, n))
            cols.append((0.5 * (R + R.T)).ravel(order="F"))
        cols.append(np.eye(n).ravel(order="F"))
    elif kind == SOC:
        for _ in range(m):
            cols.append(rng.uniform(-1, 1, size=rows))
        c = np.zeros(rows)
        c[0] = 1.0
        cols.append(c)
    else:
        for _ in range(m):
            cols.append(rng.uniform(-1, 1, size=rows))
        cols.append(np.ones(rows))
    return np.concatenate(cols)


def close(a, b, tol, what):
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err < tol, (what, err)


SHAPES = [(LP, 1, 1), (LP, 7, 3), (LP, 40, 40), (SOC, 1, 2), (SOC, 10, 40), (SOC, 5, 3),
          (PSD, 1, 2), (PSD, 2, 1), (PSD, 5, 3), (PSD, 20, 40), (PSD, 9, 4),
          # orders the DMMA Schur kernel takes (n % 4 == 0, n <= 32): every pitch class and ragged 8-wide tiles
          (PSD, 4, 3), (PSD, 8, 5), (PSD, 12, 7), (PSD, 16, 30), (PSD, 24, 10), (PSD, 28, 6), (PSD, 32, 9)]


@pytest.mark.parametrize("kind,n,m", SHAPES)
def test_cone_trajectory_matches_oracle(be, kind, n, m):
    """Schur system, eigen-bounds, step norms and the updated scaling point along three damped
    Newton-like steps, for a batch of 9 independent problems (from 8 programs on the device kernels run one WARP per
    program, below that one CTA per program: test_both_thread_layouts_give_the_same_results covers the other one)."""
    rng = np.random.Generator(np.random.PCG64(1000 * kind + 10 * n + m))
    B = 9
    data = np.stack([random_cone_data(kind, n, m, rng) for _ in range(B)])
    cone = be.cone(kind, n, m, data)
    refs = [OracleCone(kind, n, m, data[p]) for p in range(B)]
    for it in range(3):
        G, AW, AQc, sc = cone.schur()
        for p in range(B):
            Go, AWo, AQco, sco = refs[p].schur()
            dscale = np.sqrt(np.outer(np.diag(Go), np.diag(Go))) + 1e-300
            assert (np.abs(G[p] - Go) / dscale).max() < 1e-10, ("H", it, p)
            close(AW[p], AWo, 1e-10, "AW")
            close(AQc[p], AQco, 1e-10, "AQc")
            close(sc[p], sco, 1e-10, "scalars")
        y = rng.uniform(-1, 1, size=(B, m)) * 0.3 / np.sqrt(m)
        cw = rng.uniform(0.5, 1.5, size=B)
        ev = cone.eigen(y, cw)
        for p in range(B):
            close(ev[p], refs[p].eigen(y[p], cw[p]), 1e-9, "eigen")
        out = cone.prepare(y, cw, 1.0)
        steps = np.zeros(B)
        for p in range(B):
            ro = refs[p].prepare(y[p], cw[p], 1.0)
            close(out[p], ro, 1e-9, "prepare")
            steps[p] = min(1.0, 2.0 / ro[0] ** 2)
        info = cone.take_step(steps, 1.0)
        assert (info == 0).all()
        W = cone.get_state()
        for p in range(B):
            refs[p].take_step(steps[p], 1.0)
            close(W[p], refs[p].state(), 1e-10, "W after step")


@pytest.mark.parametrize("kind,n,m", [(LP, 6, 4), (PSD, 6, 4)])
def test_affine_update_matches_oracle(be, kind, n, m):
    # dual recovery: PrepareStep with affine = true, c_weight = e_weight = 0 (cone_program.cc:505-515)
    rng = np.random.Generator(np.random.PCG64(77 + kind))
    data = np.stack([random_cone_data(kind, n, m, rng) for _ in range(2)])
    cone = be.cone(kind, n, m, data)
    refs = [OracleCone(kind, n, m, data[p]) for p in range(2)]
    y = rng.uniform(-1, 1, size=(2, m)) * 0.1
    cone.prepare(y, 0.0, 0.0, affine=True)
    W = cone.get_state()
    for p in range(2):
        refs[p].prepare(y[p], 0.0, 0.0, affine=True)
        close(W[p], refs[p].state(), 1e-12, "W after affine update")


def test_schur_accumulates_over_cones(be):
    # several cones of one program add into the same H (supernodal_assembler.cc:144-164)
    rng = np.random.Generator(np.random.PCG64(5))
    m, B = 6, 2
    kinds = [(PSD, 4), (SOC, 3), (LP, 5)]
    total = None
    ref = [None] * B
    for kind, n in kinds:
        data = np.stack([random_cone_data(kind, n, m, rng) for _ in range(B)])
        cone = be.cone(kind, n, m, data)
        got = cone.schur(accumulate_into=total)
        total = cone.last
        for p in range(B):
            r = OracleCone(kind, n, m, data[p]).schur()
            ref[p] = r if ref[p] is None else tuple(a + b for a, b in zip(ref[p], r))
    for p in range(B):
        for a, b, name in zip([g[p] for g in got], ref[p], ["H", "AW", "AQc", "scalars"]):
            close(a, b, 1e-12, name)


@pytest.mark.parametrize("N", [1, 2, 17, 40, 100])
def test_small_cholesky_and_solve(be, N):
    rng = np.random.Generator(np.random.PCG64(N))
    B = 9
    H = []
    for _ in range(B):
        R = rng.uniform(-1, 1, size=(N, N + 3))
        H.append(R @ R.T + 0.1 * np.eye(N))
    H = np.stack(H)
    L, info, buf = be.potrf(H)
    assert (info == 0).all()
    for p in range(B):
        close(L[p], np.linalg.cholesky(H[p]), 1e-12, "L")
    X = rng.uniform(-1, 1, size=(B, N))
    sol = be.potrs(buf, B, N, X)
    for p in range(B):
        close(sol[p], np.linalg.solve(H[p], X[p]), 1e-9, "solve")


def test_small_cholesky_reports_first_bad_pivot(be):
    # block_triangular_operations.cc:193-196: a non-positive pivot makes Factor() fail
    H = np.stack([np.diag([1.0, 2.0, -1.0, 3.0]), np.eye(4)])
    _, info, _ = be.potrf(H)
    assert info.tolist() == [3, 0]


@pytest.mark.parametrize("n,m", [(7, 5), (8, 5), (20, 40)])
def test_psd_schur_with_indefinite_scaling_point_uses_the_classic_form(be, n, m):
    """The batched PSD Schur kernel takes the symmetric form (W = L L^T, packed L^T A_i L) and falls
    back to the reference's formula H_ij = tr(A_i W A_j W) (dense_lmi_constraint.cc:62-103) when W
    does not factor; the formula itself is defined for any symmetric W."""
    rng = np.random.Generator(np.random.PCG64(99))
    B = 2
    data = np.stack([random_cone_data(PSD, n, m, rng) for _ in range(B)])
    cone = be.cone(PSD, n, m, data)
    Ws = []
    for p in range(B):
        R = rng.uniform(-1, 1, size=(n, n))
        W = 0.5 * (R + R.T)                      # indefinite
        if p == 1:
            W = W @ W.T + 0.1 * np.eye(n)        # positive definite: symmetric form
        Ws.append(W)
    cone.set_state(np.stack([W.ravel(order="F") for W in Ws]))
    G, AW, AQc, sc = cone.schur()
    for p in range(B):
        mats = [data[p][i * n * n:(i + 1) * n * n].reshape(n, n, order="F") for i in range(m + 1)]
        W, Cm = Ws[p], mats[m]
        Href = np.array([[np.trace(mats[i] @ W @ mats[j] @ W) for j in range(m)] for i in range(m)])
        close(G[p], np.tril(Href), 1e-11, "H")
        close(AW[p], np.array([np.trace(W @ mats[j]) for j in range(m)]), 1e-11, "AW")
        close(AQc[p], np.array([np.trace(W @ Cm @ W @ mats[j]) for j in range(m)]), 1e-11, "AQc")
        close(sc[p], np.array([np.trace(W @ Cm), np.trace(W @ Cm @ W @ Cm)]), 1e-11, "scalars")


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(20, 40), (8, 3), (32, 12), (12, 9), (16, 15), (24, 16), (4, 8), (8, 60)])
def test_dmma_schur_kernel_against_the_dfma_team_kernel(n, m):
    """A/B of the two device kernels behind cxb_small_schur for dense LMI blocks (cxb_set_small_psd_mma): the DMMA
    kernel with the scaled matrices in shared memory against the DFMA team kernel, on a random positive definite
    scaling point, with and without accumulation into an existing system. The shapes cover the leftover matrices of
    the slot-per-warp layout ((m + 1) mod 8 = 1 or 2: scaled by the whole CTA; else by their warps), one to four DMMA
    tiles per side, and Gram tile lists of one and two passes."""
    be = Backend("device")
    rng = np.random.Generator(np.random.PCG64(7 * n + m))
    B = 5
    data = np.stack([random_cone_data(PSD, n, m, rng) for _ in range(B)])
    Ws = []
    for p in range(B):
        R = rng.uniform(-1, 1, size=(n, n))
        Ws.append((R @ R.T / n + 0.5 * np.eye(n)).ravel(order="F"))
    out = []
    for enabled in (2, 1, 0):
        be.lib.cxb_set_small_psd_mma(enabled)
        try:
            cone = be.cone(PSD, n, m, data)
            cone.set_state(np.stack(Ws))
            first = cone.schur()
            second = cone.schur(accumulate_into=cone.last)   # G <- 2 G
            out.append((first, second))
        finally:
            be.lib.cxb_set_small_psd_mma(2)
    for variant in (0, 1):   # both DMMA layouts against the DFMA team kernel
        for (a, b) in zip(out[variant], out[2]):
            for x, y, name in zip(a, b, ("H", "AW", "AQc", "scalars")):
                close(x, y, 1e-12, name)
    for x, y in zip(out[0][0], out[0][1]):
        close(2 * x, y, 1e-12, "accumulate")


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,m", [(LP, 40, 40), (SOC, 10, 40), (PSD, 20, 40), (PSD, 9, 4)])
def test_both_thread_layouts_give_the_same_results(kind, n, m):
    """cxb_set_small_team_mode: one warp per program (default for batches) against one CTA per program, on the same
    batch: Schur system, eigen-bounds, step norms, updated scaling point."""
    be = Backend("device")
    rng = np.random.Generator(np.random.PCG64(31 * kind + n + m))
    B = 11
    data = np.stack([random_cone_data(kind, n, m, rng) for _ in range(B)])
    y = rng.uniform(-0.05, 0.05, size=(B, m))
    out = []
    for mode in (1, 0):
        be.lib.cxb_set_small_team_mode(mode)
        try:
            cone = be.cone(kind, n, m, data)
            res = [cone.schur()]
            res.append(cone.eigen(y, 1.0))
            res.append(cone.prepare(y, 1.0))
            cone.take_step(0.5)
            res.append(cone.get_state())
            res.append(cone.schur())
            out.append(res)
        finally:
            be.lib.cxb_set_small_team_mode(2)
    for a, b in zip(out[0], out[1]):
        if isinstance(a, tuple):
            for x, z in zip(a, b):
                close(x, z, 1e-11, "schur")
        else:
            close(a, b, 1e-8, "eigen / prepare / state")



@pytest.mark.gpu
@pytest.mark.parametrize("threads", [128, 64, 32])
def test_fused_launch_over_the_cones_of_a_program_is_bit_identical(threads):
    """cxb_small_*_multi (one launch, grid = programs x cones — what the batched lock step uses) against the per-cone
    entry points on the same batch of {PSD 20, PSD 8, SOC 10, LP 40, PSD 12 ... } cones: eigen-bounds, step norms and
    the updated scaling points must agree bit for bit, for every CTA size of cxb_set_small_cone_threads. Ten cones, so
    the list is also split over two launches."""
    from small_backend import ConeDesc
    be = Backend("device")
    vp = C.c_void_p
    rng = np.random.Generator(np.random.PCG64(2024))
    B, m = 13, 12
    shapes = [(PSD, 20), (PSD, 8), (SOC, 10), (LP, 40), (PSD, 12), (SOC, 3), (LP, 5), (PSD, 4), (PSD, 16), (LP, 17)]
    data = [np.stack([random_cone_data(kind, n, m, rng) for _ in range(B)]) for kind, n in shapes]
    y = rng.uniform(-0.05, 0.05, size=(B, m))
    cw = rng.uniform(0.5, 1.5, size=B)
    steps = rng.uniform(0.3, 1.0, size=B)
    nc = len(shapes)

    def run(fused):
        cones = [be.cone(kind, n, m, d) for (kind, n), d in zip(shapes, data)]
        if not fused:
            ev = np.stack([c.eigen(y, cw) for c in cones], axis=1)                  # B x nc x 4
            pr = np.stack([c.prepare(y, cw, 1.0) for c in cones], axis=1)           # B x nc x 2
            for c in cones:
                assert (c.take_step(steps, 1.0) == 0).all()
            return ev, pr, [c.get_state() for c in cones]
        descs = (ConeDesc * nc)(*[c.desc for c in cones])
        lib = be.lib
        yd, cwd, sd = be.upload(y), be.upload(cw), be.upload(steps)
        out = be.alloc(B * nc * 4)
        info = be.alloc(B, np.int32)
        for f in (lib.cxb_small_set_identity_multi, lib.cxb_small_eigen_multi, lib.cxb_small_prepare_multi,
                  lib.cxb_small_take_step_multi):
            f.restype = C.c_int
        assert lib.cxb_small_set_identity_multi(vp(None), C.c_int(B), C.c_int(nc), descs, vp(None)) == 0
        assert lib.cxb_small_eigen_multi(vp(None), C.c_int(B), C.c_int(nc), descs, be.ptr(yd), C.c_long(m),
                                         C.c_double(0.0), be.ptr(cwd), be.ptr(out), C.c_long(4 * nc), vp(None)) == 0
        ev = be.download(out).reshape(B, nc, 4).copy()
        assert lib.cxb_small_prepare_multi(vp(None), C.c_int(B), C.c_int(nc), descs, be.ptr(yd), C.c_long(m),
                                           C.c_int(0), C.c_double(0.0), be.ptr(cwd), C.c_double(1.0), be.ptr(out),
                                           C.c_long(4 * nc), vp(None)) == 0
        pr = be.download(out).reshape(B, nc, 4)[:, :, :2].copy()
        assert lib.cxb_small_take_step_multi(vp(None), C.c_int(B), C.c_int(nc), descs, C.c_double(1.0), be.ptr(sd),
                                             C.c_double(1.0), be.ptr(info), vp(None)) == 0
        assert (be.download(info) == 0).all()
        return ev, pr, [c.get_state() for c in cones]

    be.lib.cxb_set_small_cone_threads.restype = None
    try:
        be.lib.cxb_set_small_cone_threads(128)
        reference = run(False)                      # per-cone launches, 128 threads
        be.lib.cxb_set_small_cone_threads(threads)
        fused = run(True)
        single = run(False) if threads != 128 else reference
    finally:
        be.lib.cxb_set_small_cone_threads(64)       # the default
    for got in (fused, single):
        # the reductions of a 128-thread team and of a smaller one combine in different orders: bitwise only at 128
        if threads == 128:
            assert np.array_equal(got[0], reference[0]) and np.array_equal(got[1], reference[1])
            for a, b in zip(got[2], reference[2]):
                assert np.array_equal(a, b)
        else:
            close(got[0], reference[0], 1e-9, "eigen")
            close(got[1], reference[1], 1e-9, "prepare")
            for a, b in zip(got[2], reference[2]):
                close(a, b, 1e-12, "state")
    assert np.array_equal(fused[0], single[0]) and np.array_equal(fused[1], single[1])


def _packed_vs_full(be, n, m, B, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    data = np.stack([random_cone_data(PSD, n, m, rng) for _ in range(B)])
    y = rng.uniform(-0.05, 0.05, size=(B, m))
    cw = rng.uniform(0.5, 1.5, size=B)
    out = []
    for packed in (False, True):
        cone = be.cone(PSD, n, m, data)
        if packed:
            assert cone.pack() is False
            kp = n * (n + 1) // 2
            got = cone.packed_host().reshape(B, -1)[:, :(m + 1) * kp].reshape(B, m + 1, kp)
            for p in (0, B - 1):
                for j in (0, m):
                    A = data[p][j * n * n:(j + 1) * n * n].reshape(n, n, order="F")
                    assert np.array_equal(got[p, j], np.concatenate([A[c:, c] for c in range(n)]))
        res = [cone.eigen(y, cw), cone.prepare(y, cw, 1.0)]
        cone.take_step(np.full(B, 0.7), 1.0)
        res.append(cone.get_state())
        res.append(cone.prepare(y, cw, 0.0, affine=True))
        res.append(cone.get_state())
        out.append(res)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)          # same sums in the same order: bit for bit
    # a matrix that is not symmetric: the copy is refused
    bad = data.copy()
    bad[B - 1][(m // 2) * n * n + 1] += 1e-9      # entry (1, 0) of one matrix
    assert be.cone(PSD, n, m, bad).pack() is True


@pytest.mark.parametrize("n,m", [(6, 4), (5, 3), (20, 7)])
def test_slack_from_the_packed_lower_triangles_is_bit_identical_host_emulation(n, m):
    """The slack passes of the LMI blocks read the packed lower triangles of the (symmetric) matrices when the batch
    has made that copy (cxb_small_cone.packed): eigen-bounds, step norms, dual-recovery update and the scaling point
    after a step must not change by a bit; index arithmetic checked here through the host stand-in of the team."""
    _packed_vs_full(Backend("emul"), n, m, 3, 100 * n + m)


@pytest.mark.gpu
@pytest.mark.parametrize("n,m", [(6, 4), (5, 3), (20, 40), (32, 9)])
def test_slack_from_the_packed_lower_triangles_is_bit_identical(n, m):
    _packed_vs_full(Backend("device"), n, m, 7, 100 * n + m)


def test_index_divider_is_exact():
    """small_cone_math.cuh replaces the integer division of the flat element index by a float reciprocal with a
    one-step correction; the element loops rely on it being exact for every quotient below 2^22."""
    lib = Backend("emul").lib
    lib.emul_divider_mismatches.restype = C.c_long
    assert lib.emul_divider_mismatches(C.c_int(700), C.c_int(2500)) == 0
    assert lib.emul_divider_mismatches(C.c_int(9), C.c_int(4000000)) == 0    # quotients up to 4e6 ~ 2^22
