"""Synthetic problem generators of the BASELINE.json configurations, shared by bench.py and
the tests. All PCG64-seeded so that the oracle and the device see the same bytes."""
import numpy as np


def random_sym(rng, n):
    R = rng.uniform(-1.0, 1.0, size=(n, n))
    return 0.5 * (R + R.T)


def random_dense_lmi(n, m, seed):
    """conex/test/test_sdp.cc:170-185 + test_util.cc:19,67-73: A_i = sym(U[-1,1]), C = I."""
    rng = np.random.Generator(np.random.PCG64(seed))
    mats = [random_sym(rng, n) for _ in range(m)]
    return mats, np.eye(n)


def maxcut_lmi(n, seed, p=0.5):
    """MaxCut dual: A_i = -e_i e_i^T, C = -L/4, b = -1."""
    rng = np.random.Generator(np.random.PCG64(seed))
    upper = np.triu(rng.random((n, n)) < p, 1).astype(np.float64)
    adj = upper + upper.T
    lap = np.diag(adj.sum(axis=1)) - adj
    mats = []
    for i in range(n):
        E = np.zeros((n, n))
        E[i, i] = -1.0
        mats.append(E)
    return mats, -lap / 4.0, -np.ones(n)


def lovasz_theta_lmi(n, num_edges, seed):
    """Lovász-theta dual: vars (t, y_e); A_0 = -I, A_e = E_ij + E_ji, C = -J,
    maximise -t."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pairs = [(i, j) for i in range(n) for j in range(i + 1, n)]
    idx = rng.choice(len(pairs), size=num_edges, replace=False)
    mats = [-np.eye(n)]
    for k in sorted(idx.tolist()):
        i, j = pairs[k]
        E = np.zeros((n, n))
        E[i, j] = 1.0
        E[j, i] = 1.0
        mats.append(E)
    b = np.zeros(num_edges + 1)
    b[0] = -1.0
    return mats, -np.ones((n, n)), b


def small_multicone_problem(seed, m=40, psd_blocks=3, psd_order=20, soc_cones=2, soc_order=10, lp_rows=40):
    """One program of BASELINE config 3: `psd_blocks` dense LMI blocks
    (A_i = sym(U[-1,1]), C = I), `soc_cones` Lorentz cones of order `soc_order` (A uniform,
    c = (1, 0, ...), interfaces/python/test/run_tests.py:23-34) and one LP block (A uniform, c = 1),
    all on the same m variables; b = sum over cones of the feasible objective AW/2 at W = I
    (cone_program.cc:535-545), so both primal and dual are strictly feasible."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cones = []
    b = np.zeros(m)
    for _ in range(psd_blocks):
        mats = [random_sym(rng, psd_order) for _ in range(m)]
        cones.append(("psd", mats, np.eye(psd_order)))
        b += 0.5 * np.array([np.trace(M) for M in mats])
    for _ in range(soc_cones):
        A = rng.uniform(-1.0, 1.0, size=(soc_order + 1, m))
        c = np.zeros(soc_order + 1)
        c[0] = 1.0
        cones.append(("soc", A, c))
        b += 0.5 * 2.0 * A[0, :]
    if lp_rows:
        A = rng.uniform(-1.0, 1.0, size=(lp_rows, m))
        cones.append(("lp", A, np.ones(lp_rows)))
        b += 0.5 * A.sum(axis=0)
    return cones, b


def add_cones(P, cones):
    for kind, A, c in cones:
        if kind == "psd":
            P.add_dense_lmi(A, c)
        elif kind == "soc":
            P.add_soc(A, c)
        else:
            P.add_linear(A, c)


def block_arrow_program(blocks, private, shared, order, seed):
    """`blocks` LMI cones of order `order`, cone k on its own `private` variables plus the same
    `shared` variables: H is block-arrow."""
    rng = np.random.default_rng(seed)
    m = blocks * private + shared
    cones = []
    for k in range(blocks):
        variables = list(range(k * private, (k + 1) * private)) + list(range(blocks * private, m))
        mats = [random_sym(rng, order) for _ in variables]
        cones.append((mats, np.eye(order), variables))
    return m, cones
