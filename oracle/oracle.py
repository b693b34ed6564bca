"""ctypes binding of oracle/liblstm_oracle*.so — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  See oracle/lstm_oracle.cc for what the oracle restates (R/lstm.cc:142-356)
and how it is pinned.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
W, U, B_, WHY, BY = 0, 1, 2, 3, 4
PARAM, GRAD, MEM = 0, 1, 2
NAMES = ["W", "U", "b", "Why", "by"]


def build():
    """Compile the oracle with oracle/Makefile (building the checker is not using it)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


def _load(mt):
    name = "liblstm_oracle_mt.so" if mt else "liblstm_oracle.so"
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


class Oracle:
    """One CPU LSTM (params + per-timestep state), float32 (`dtype='f32'`) or float64."""

    def __init__(self, M, N, S, B, dtype="f32", mt=False, threads=1):
        self.lib = _load(mt)
        self.M, self.N, self.S, self.B = M, N, S, B
        self.pfx = "oracle32" if dtype == "f32" else "oracle64"
        self.np = np.float32 if dtype == "f32" else np.float64
        self.cr = C.c_float if dtype == "f32" else C.c_double
        f = self._f
        f("create").restype = C.c_void_p
        self.lib.oracle_set_threads.restype = C.c_int
        self.threads = self.lib.oracle_set_threads(int(threads))
        self.o = C.c_void_p(f("create")(M, N, S, B))
        f("forward").restype = C.c_double
        f("train").restype = C.c_double
        f("eval_bpc").restype = C.c_double
        f("tensor_size").restype = C.c_long

    def _f(self, name):
        return getattr(self.lib, f"{self.pfx}_{name}")

    def __del__(self):
        try:
            self._f("destroy")(self.o)
        except Exception:
            pass

    def shape(self, which):
        M, N = self.M, self.N
        return [(4 * N, M), (4 * N, N), (4 * N, 1), (M, N), (M, 1)][which]

    def set_options(self, softmax_shift=0, dense_onehot=0):
        self._f("set_options")(self.o, int(softmax_shift), int(dense_onehot))

    # tensors are exchanged as numpy arrays of the mathematical shape (rows, cols); the
    # buffer handed to C is column-major (Fortran order) like Eigen's.
    def set(self, kind, which, arr):
        a = np.asfortranarray(np.asarray(arr, dtype=self.np).reshape(self.shape(which)))
        assert self._f("set_tensor")(self.o, kind, which, a.ctypes.data_as(C.c_void_p)) == 0

    def get(self, kind, which):
        a = np.empty(self.shape(which), dtype=self.np, order="F")
        assert self._f("get_tensor")(self.o, kind, which, a.ctypes.data_as(C.c_void_p)) == 0
        return a

    def set_params(self, params):
        for i, p in enumerate(params):
            self.set(PARAM, i, p)

    def params(self):
        return [self.get(PARAM, i) for i in range(5)]

    def grads(self):
        return [self.get(GRAD, i) for i in range(5)]

    def state(self, what, t):
        """what: 'h','c' (N,B) | 'g','dg' (4N,B) | 'probs' (M,B)."""
        code = {"h": 0, "c": 1, "g": 2, "probs": 3, "dg": 5}[what]
        rows = {0: self.N, 1: self.N, 2: 4 * self.N, 3: self.M, 5: 4 * self.N}[code]
        a = np.empty((rows, self.B), dtype=self.np, order="F")
        assert self._f("get_state")(self.o, code, t, a.ctypes.data_as(C.c_void_p)) == 0
        return a

    def set_state(self, what, t, arr):
        code = {"h": 0, "c": 1}[what]
        a = np.asfortranarray(np.asarray(arr, dtype=self.np).reshape(self.N, self.B))
        assert self._f("set_state")(self.o, code, t, a.ctypes.data_as(C.c_void_p)) == 0

    def set_window(self, x_idx, t_idx):
        """x_idx, t_idx: int arrays [S][B]; row 0 of x is unused; -1 = all-zero column."""
        x = np.ascontiguousarray(np.asarray(x_idx, dtype=np.int32).reshape(self.S, self.B))
        t = np.ascontiguousarray(np.asarray(t_idx, dtype=np.int32).reshape(self.S, self.B))
        self._f("set_window")(self.o, x.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))

    def window(self):
        x = np.empty((self.S, self.B), dtype=np.int32)
        t = np.empty((self.S, self.B), dtype=np.int32)
        self._f("get_window")(self.o, x.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
        return x, t

    def set_positions(self, pos):
        p = np.ascontiguousarray(np.asarray(pos, dtype=np.uint64))
        assert p.size == self.B
        self._f("set_positions")(self.o, p.ctypes.data_as(C.c_void_p))

    def positions(self):
        p = np.empty(self.B, dtype=np.uint64)
        self._f("get_positions")(self.o, p.ctypes.data_as(C.c_void_p))
        return p

    def forward(self):
        return float(self._f("forward")(self.o))

    def backward(self):
        self._f("backward")(self.o)

    def adagrad(self, lr):
        self._f("adagrad")(self.o, C.c_double(lr))

    def carry(self, stride):
        self._f("carry")(self.o, int(stride))

    def advance(self, data, stride):
        d = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        self._f("advance")(self.o, d.ctypes.data_as(C.c_void_p), C.c_size_t(d.size), int(stride))

    def train(self, data, iters, stride=1, lr=0.1):
        """Run `iters` iterations of R/lstm.cc:151-272; returns (losses[iters], wall seconds)."""
        d = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        losses = np.zeros(iters, dtype=np.float64)
        secs = self._f("train")(self.o, d.ctypes.data_as(C.c_void_p), C.c_size_t(d.size), int(iters),
                                int(stride), C.c_double(lr), losses.ctypes.data_as(C.c_void_p))
        return losses, float(secs)

    def eval_bpc(self, data):
        d = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        return float(self._f("eval_bpc")(self.o, d.ctypes.data_as(C.c_void_p), C.c_size_t(d.size)))

    def sample(self, h0, c0, seed, n, greedy=False):
        h = np.ascontiguousarray(np.asarray(h0, dtype=self.np).reshape(self.N))
        c = np.ascontiguousarray(np.asarray(c0, dtype=self.np).reshape(self.N))
        out = np.zeros(n, dtype=np.uint8)
        self._f("sample")(self.o, h.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                          C.c_uint64(seed), out.ctypes.data_as(C.c_void_p), C.c_size_t(n), int(bool(greedy)))
        return out


def randn(rows, cols, mean, sd, seed, dtype=np.float32):
    """R/lstm.cc:364-380 with an explicit seed: mt19937 + normal_distribution<double>, (i,j) fill order."""
    lib = _load(False)
    a = np.empty((rows, cols), dtype=dtype, order="F")
    fn = lib.oracle32_randn if dtype == np.float32 else lib.oracle64_randn
    fn(a.ctypes.data_as(C.c_void_p), rows, cols, C.c_float(mean), C.c_float(sd), C.c_uint64(seed))
    return a


def randn_d(rows, cols, mean, sd, seed):
    """The double-precision snapshots' randn(MatrixXd&, double, double) (OV/lstm_eigen_class_batch/lstm.cc:513-529)."""
    lib = _load(False)
    a = np.empty((rows, cols), dtype=np.float64, order="F")
    lib.oracle64_randn_d(a.ctypes.data_as(C.c_void_p), rows, cols, C.c_double(mean), C.c_double(sd), C.c_uint64(seed))
    return a


def uniform01(seed, n):
    """n draws of mt19937(seed) + uniform_real_distribution<double>(0, 1)."""
    lib = _load(False)
    a = np.empty(n, dtype=np.float64)
    lib.oracle_uniform01(C.c_uint64(seed), C.c_size_t(n), a.ctypes.data_as(C.c_void_p))
    return a


def init_params(M, N, seed, sd=0.01, forget_bias=0.0, dtype=np.float32):
    """R/lstm.cc:113-119 (k-th randn call seeded seed+k); forget bias per OV/lstm_eigen_class_batch/lstm.cc:81."""
    Wm = randn(4 * N, M, 0, sd, seed + 0, dtype)
    Um = randn(4 * N, N, 0, sd, seed + 1, dtype)
    Why = randn(M, N, 0, sd, seed + 2, dtype)
    b = np.zeros((4 * N, 1), dtype=dtype)
    b[2 * N:3 * N] = forget_bias
    by = np.zeros((M, 1), dtype=dtype)
    return [Wm, Um, b, Why, by]
