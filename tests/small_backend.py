"""Two ways to run the small-cone arithmetic (conex_b200/csrc/device/small_cone_math.cuh):

* "device": the real CUDA kernels through the cxb_small_* entry points of libconex_b200.so
  (device memory owned by torch tensors) — used by the `-m gpu` parity tests;
* "emul":  the same header compiled for the host with a serial stand-in for the CTA
  (tests/emul/emul_small.cc) — test infrastructure that lets the CPU suite check the index
  arithmetic of the header against the oracle. It is never part of the product library.

Both expose: cone(type, n, m, data_batch) -> SmallCone with schur / eigen / prepare / take_step /
get_state / set_state, and potrf / potrs on batches of small KKT matrices.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL_SRC = os.path.join(ROOT, "tests", "emul", "emul_small.cc")
EMUL_SO = os.path.join(ROOT, "tests", "emul", "_build", "libemul_small.so")
HEADER = os.path.join(ROOT, "conex_b200", "csrc", "device", "small_cone_math.cuh")

LP, SOC, PSD = 0, 1, 2
vp = C.c_void_p


class ConeDesc(C.Structure):
    _fields_ = [("type", C.c_int), ("n", C.c_int), ("m", C.c_int), ("data", vp), ("data_stride", C.c_long),
                ("state", vp), ("state_stride", C.c_long), ("work", vp), ("work_stride", C.c_long),
                ("packed", vp), ("packed_stride", C.c_long)]


def align4(n):
    return (n + 3) & ~3


def rows_of(kind, n):
    return {LP: n, SOC: n + 1, PSD: n * n}[kind]


def state_size(kind, n):
    return {LP: 3 * align4(n), SOC: 2 * align4(n + 1), PSD: 3 * align4(n * n)}[kind]


def work_size(kind, n, m):
    return {LP: 0, SOC: (n + 1) * (m + 4), PSD: (m + 2) * n * n}[kind]


def w_size(kind, n):
    return rows_of(kind, n)


def build_emul():
    os.makedirs(os.path.dirname(EMUL_SO), exist_ok=True)
    newest = max(os.path.getmtime(EMUL_SRC), os.path.getmtime(HEADER))
    if not os.path.exists(EMUL_SO) or os.path.getmtime(EMUL_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", EMUL_SO, EMUL_SRC])
    return EMUL_SO


class Backend:
    """kind = "emul" | "device"."""

    def __init__(self, kind):
        self.kind = kind
        if kind == "emul":
            self.lib = C.CDLL(build_emul())
            self.prefix = "emul_small_"
        else:
            import devlib
            self.dev = devlib
            self.lib = devlib.product().lib
            self.prefix = "cxb_small_"

    # ---- memory ------------------------------------------------------------------------------
    def alloc(self, count, dtype=np.float64):
        if self.kind == "emul":
            return np.zeros(max(count, 1), dtype=dtype)
        import torch
        return torch.zeros(max(count, 1), dtype=torch.float64 if dtype == np.float64 else torch.int32,
                           device="cuda")

    def upload(self, a, dtype=np.float64):
        a = np.ascontiguousarray(np.asarray(a, dtype=dtype).ravel())
        if self.kind == "emul":
            return a.copy()
        import torch
        return torch.from_numpy(a).cuda()

    def download(self, t):
        if self.kind == "emul":
            return np.array(t, copy=True)
        import torch
        torch.cuda.synchronize()
        return t.cpu().numpy()

    def ptr(self, t):
        if t is None:
            return vp(None)
        if self.kind == "emul":
            return vp(t.ctypes.data)
        return vp(t.data_ptr())

    def call(self, name, *args):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        if self.kind == "device":
            args = (vp(None),) + args  # stream
            if name in ("set_identity", "schur", "eigen", "prepare", "take_step", "potrf", "potrs"):
                args = args + (vp(None),)  # active mask
        rc = f(*args)
        assert rc == 0, (name, rc)

    # ---- cones -------------------------------------------------------------------------------
    def cone(self, kind, n, m, data):
        """data: (batch, rows * (m + 1)) array — per problem the column-major rows x (m+1) block."""
        return SmallCone(self, kind, n, m, np.asarray(data, dtype=np.float64))

    def potrf(self, H):
        """H: (batch, N, N) symmetric matrices. Returns (L batch x N x N lower, info batch)."""
        batch, N, _ = H.shape
        buf = self.upload(np.stack([h.T for h in H]))  # column-major
        info = self.alloc(batch, np.int32)
        self.call("potrf", C.c_int(batch), C.c_int(N), self.ptr(buf), C.c_long(N), C.c_long(N * N),
                  self.ptr(info))
        out = self.download(buf).reshape(batch, N, N).transpose(0, 2, 1)
        return np.tril(out), self.download(info), buf

    def potrs(self, Lbuf, batch, N, X):
        x = self.upload(X)
        self.call("potrs", C.c_int(batch), C.c_int(N), self.ptr(Lbuf), C.c_long(N), C.c_long(N * N),
                  self.ptr(x), C.c_long(N))
        return self.download(x).reshape(batch, N)


class SmallCone:
    def __init__(self, be, kind, n, m, data):
        self.be, self.kind, self.n, self.m = be, kind, n, m
        self.batch = data.shape[0]
        self.rows = rows_of(kind, n)
        assert data.shape[1] == self.rows * (m + 1)
        self.ss, self.ws = state_size(kind, n), work_size(kind, n, m)
        self.data = be.upload(data)
        self.state = be.alloc(self.batch * self.ss)
        self.work = be.alloc(self.batch * self.ws) if self.ws else None
        self.desc = ConeDesc(kind, n, m, be.ptr(self.data), self.rows * (m + 1), be.ptr(self.state), self.ss,
                             be.ptr(self.work), self.ws)
        be.call("set_identity", C.c_int(self.batch), C.byref(self.desc))

    def pack(self):
        """PSD: attach the packed copy of the operator (lower triangles of the m + 1 matrices) to the descriptor.
        Returns True when some matrix is not symmetric (the copy is then left detached, as the library does)."""
        be, n, m, B = self.be, self.n, self.m, self.batch
        kp = n * (n + 1) // 2
        stride = align4((m + 1) * kp)
        if be.kind == "emul":
            full = np.asarray(self.data).reshape(B, m + 1, n, n)          # [p, j, col, row]
            packed = np.zeros((B, stride))
            idx = [(c, r) for c in range(n) for r in range(c, n)]
            cols, rows = np.array([c for c, _ in idx]), np.array([r for _, r in idx])
            packed[:, :(m + 1) * kp] = full[:, :, cols, rows].reshape(B, (m + 1) * kp)
            asymmetric = not np.array_equal(full, full.transpose(0, 1, 3, 2))
            self.packed = packed.ravel().copy()
        else:
            self.packed = be.alloc(B * stride)
            flag = be.alloc(1, np.int32)
            f = be.lib.cxb_small_pack_symmetric
            f.restype = C.c_int
            assert f(vp(None), C.c_int(B), C.byref(self.desc), be.ptr(self.packed), C.c_long(stride),
                     be.ptr(flag)) == 0
            asymmetric = bool(be.download(flag)[0])
        if not asymmetric:
            self.desc.packed = be.ptr(self.packed)
            self.desc.packed_stride = stride
        return asymmetric

    def packed_host(self):
        return self.be.download(self.packed)

    def get_state(self):
        s = self.be.download(self.state).reshape(self.batch, self.ss)
        return s[:, :w_size(self.kind, self.n)].copy()

    def set_state(self, W):
        s = self.be.download(self.state).reshape(self.batch, self.ss)
        s[:, :w_size(self.kind, self.n)] = W
        if self.be.kind == "emul":
            self.state[:] = s.ravel()
        else:
            import torch
            self.state.copy_(torch.from_numpy(s.ravel()).cuda())

    def schur(self, accumulate_into=None):
        """Returns (G batch x m x m lower, AW, AQc, scal batch x 2)."""
        be, m, B = self.be, self.m, self.batch
        ld = m + 2
        if accumulate_into is None:
            G, AW, AQc, sc = be.alloc(B * ld * m), be.alloc(B * m), be.alloc(B * m), be.alloc(B * 2)
            acc = 0
        else:
            G, AW, AQc, sc = accumulate_into
            acc = 1
        be.call("schur", C.c_int(B), C.byref(self.desc), be.ptr(G), C.c_long(ld), C.c_long(ld * m),
                be.ptr(AW), be.ptr(AQc), C.c_long(m), be.ptr(sc), C.c_long(2), C.c_int(acc))
        self.last = (G, AW, AQc, sc)
        Gh = be.download(G).reshape(B, m, ld)[:, :, :m].transpose(0, 2, 1)
        return (np.tril(Gh), be.download(AW).reshape(B, m), be.download(AQc).reshape(B, m),
                be.download(sc).reshape(B, 2))

    def eigen(self, y, cw):
        be, B = self.be, self.batch
        yd, out = be.upload(y), be.alloc(B * 4)
        cwd = be.upload(np.broadcast_to(np.asarray(cw, dtype=np.float64), (B,)))
        be.call("eigen", C.c_int(B), C.byref(self.desc), be.ptr(yd), C.c_long(self.m), C.c_double(0.0),
                be.ptr(cwd), be.ptr(out), C.c_long(4))
        return be.download(out).reshape(B, 4)

    def prepare(self, y, cw, ew=1.0, affine=False):
        be, B = self.be, self.batch
        yd, out = be.upload(y), be.alloc(B * 2)
        cwd = be.upload(np.broadcast_to(np.asarray(cw, dtype=np.float64), (B,)))
        be.call("prepare", C.c_int(B), C.byref(self.desc), be.ptr(yd), C.c_long(self.m), C.c_int(int(affine)),
                C.c_double(0.0), be.ptr(cwd), C.c_double(ew), be.ptr(out), C.c_long(2))
        return be.download(out).reshape(B, 2)

    def take_step(self, step, ew=1.0):
        be, B = self.be, self.batch
        sd = be.upload(np.broadcast_to(np.asarray(step, dtype=np.float64), (B,)))
        info = be.alloc(B, np.int32)
        be.call("take_step", C.c_int(B), C.byref(self.desc), C.c_double(1.0), be.ptr(sd), C.c_double(ew),
                be.ptr(info))
        return be.download(info)
