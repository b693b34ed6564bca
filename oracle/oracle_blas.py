"""BLAS-backed restatement of the batched CPU path — TEST INFRASTRUCTURE / CPU BASELINE, NOT PRODUCT CODE.

The north-star asks for the reference's `lstm_eigen_BLAS` CPU path timed next to the GPU numbers.  That program
(OV/lstm_eigen_BLAS/lstm.cc) routes its eight per-timestep contractions through `BLAS_mmul` -> `cblas_sgemm`
(:548-573, alpha = beta = 1, column-major) and leaves the element-wise work to Eigen.  OpenBLAS and Eigen are not
installed here, so this module restates exactly that structure on the best BLAS the image has — numpy's bundled
multi-threaded OpenBLAS (`@` on float32 Fortran-ordered arrays) — statement by statement:

    forward   :228-270   g = W x + U h + b (dense one-hot x, like the reference), sigmoid / tanh, c = tanh(i u + f c'),
                         h = o c, y = Why h + by, probs = exp(y) / colsum, loss += sum(-log2 probs * target) / B
    backward  :275-343   dy = probs - target, dWhy += dy h^T, dby += rowsum(dy), dh = Why^T dy + dhnext, gate gradients,
                         dU += dg h'^T, dW += dg x^T, db += rowsum(dg), dhnext = U^T dg, dcnext = dc f
    adagrad   :346-357   m += d*d ; p -= lr * d / sqrt(m + 1e-10)

It is checked against the C++ oracle (tests/test_oracle_blas.py: same losses and weights to float32 summation-order
noise) and used by bench.py's CPU arm for the large configurations, where GEMM throughput decides the CPU's speed.
Only tests/ and bench.py's cpu_baseline / --impl reference leg may import it.
"""
import time

import numpy as np

EPS = 1e-10  # OV/lstm_eigen_BLAS/lstm.cc:26


def _f(a):
    return np.asfortranarray(a, dtype=np.float32)


class BlasOracle:
    """float32, B streams; state arrays indexed by timestep like the reference's `Eigen::MatrixXf h[S]` etc."""

    def __init__(self, M, N, S, B):
        self.M, self.N, self.S, self.B = M, N, S, B
        z = lambda r, c: np.zeros((r, c), dtype=np.float32, order="F")
        self.W, self.U, self.b = z(4 * N, M), z(4 * N, N), z(4 * N, 1)
        self.Why, self.by = z(M, N), z(M, 1)
        self.mW, self.mU, self.mb, self.mWhy, self.mby = z(4 * N, M), z(4 * N, N), z(4 * N, 1), z(M, N), z(M, 1)
        self.h = [z(N, B) for _ in range(S)]
        self.c = [z(N, B) for _ in range(S)]
        self.g = [z(4 * N, B) for _ in range(S)]
        self.x = [z(M, B) for _ in range(S)]
        self.target = [z(M, B) for _ in range(S)]
        self.probs = [z(M, B) for _ in range(S)]
        self.grads = None
        self.dg = [None] * S        # per-timestep gate gradients of the last backward (differential tests)

    def set_params(self, params):
        self.W, self.U, self.b, self.Why, self.by = [_f(np.array(p, copy=True)) for p in params]

    def params(self):
        return [self.W, self.U, self.b, self.Why, self.by]

    def set_window(self, x_idx, t_idx):
        """int [S][B]; -1 = all-zero column.  Builds the dense one-hot matrices the reference multiplies."""
        x_idx = np.asarray(x_idx).reshape(self.S, self.B)
        t_idx = np.asarray(t_idx).reshape(self.S, self.B)
        cols = np.arange(self.B)
        for t in range(self.S):
            for dst, idx in ((self.x[t], x_idx[t]), (self.target[t], t_idx[t])):
                dst[...] = 0
                ok = idx >= 0
                dst[idx[ok], cols[ok]] = 1.0

    def forward(self):
        N, B = self.N, self.B
        loss = 0.0
        for t in range(1, self.S):
            g = self.W @ self.x[t]                       # BLAS_mmul(g[t], W, x[t])          :229-231
            g += self.U @ self.h[t - 1]                  # BLAS_mmul(g[t], U, h[t-1])        :232
            g += self.b                                  # colwise + b                       :233
            g[:3 * N] = 1.0 / (1.0 + np.exp(-g[:3 * N]))  # logistic on i, o, f              :238
            g[3 * N:] = np.tanh(g[3 * N:])               # tanh on u                         :240
            c = np.tanh(g[:N] * g[3 * N:] + g[2 * N:3 * N] * self.c[t - 1])   # :242-246
            h = g[N:2 * N] * c                           # :249
            y = self.Why @ h                             # BLAS_mmul(y[t], Why, h[t])        :254
            y += self.by
            p = np.exp(y)                                # :261 (no max shift in this snapshot)
            p /= p.sum(axis=0, keepdims=True)            # :262-263
            self.g[t], self.c[t], self.h[t], self.probs[t] = _f(g), _f(c), _f(h), _f(p)
            with np.errstate(divide="ignore", invalid="ignore"):
                s = np.float32(np.sum(-np.log2(p) * self.target[t], dtype=np.float32))   # :267
            loss += float(s) / np.float32(B)             # :270
        return loss

    def backward(self):
        N = self.N
        dWhy, dby = np.zeros_like(self.Why), np.zeros_like(self.by)
        dU, dW, db = np.zeros_like(self.U), np.zeros_like(self.W), np.zeros_like(self.b)
        dhnext = np.zeros_like(self.h[0])
        dcnext = np.zeros_like(self.c[0])
        for t in range(self.S - 1, 0, -1):
            g, c = self.g[t], self.c[t]
            dy = self.probs[t] - self.target[t]                      # :290
            dWhy += dy @ self.h[t].T                                 # BLAS_mmul(dWhy, dy, h, false, true)  :292
            dby += dy.sum(axis=1, keepdims=True)                     # :296
            dh = self.Why.T @ dy + dhnext                            # :298-303
            dc = (dh * g[N:2 * N] + dcnext) * (1.0 - c * c)          # :308-310
            dg = np.empty_like(g)
            dg[N:2 * N] = dh * c                                     # do   :313
            dg[:N] = dc * g[3 * N:]                                  # di   :314
            dg[2 * N:3 * N] = dc * self.c[t - 1]                     # df   :315
            dg[3 * N:] = dc * g[:N]                                  # du   :316
            dg[:3 * N] *= g[:3 * N] * (1.0 - g[:3 * N])              # :319-320
            dg[3 * N:] *= 1.0 - g[3 * N:] * g[3 * N:]                # :323-324
            self.dg[t] = dg
            dU += dg @ self.h[t - 1].T                               # :327
            dW += dg @ self.x[t].T                                   # :328 (dense, like the reference)
            db += dg.sum(axis=1, keepdims=True)                      # :335
            dhnext = self.U.T @ dg                                   # :339-340
            dcnext = dc * g[2 * N:3 * N]                             # :343
        self.grads = [dW, dU, db, dWhy, dby]

    def adagrad(self, lr):
        lr = np.float32(lr)
        for p, d, m in zip(self.params(), self.grads, [self.mW, self.mU, self.mb, self.mWhy, self.mby]):
            m += d * d
            p -= lr * (d / np.sqrt(m + np.float32(EPS)))

    def carry(self, stride):
        stride = min(stride, self.S - 1)
        self.h[0] = self.h[stride].copy(order="F")
        self.c[0] = self.c[stride].copy(order="F")

    def train_windows(self, text, positions, iters, stride, lr):
        """`iters` iterations over the byte text: stream b's window ends at positions[b] and moves by `stride` per
        iteration (positions wrap to S like OV/lstm_eigen_BLAS/lstm.cc:196-197).  Returns (losses, wall seconds)."""
        S, B = self.S, self.B
        data = np.frombuffer(bytes(text), dtype=np.uint8)
        L = data.size
        pos = np.asarray(positions, dtype=np.int64).copy()
        losses = np.zeros(iters)
        t0 = time.perf_counter()
        for it in range(iters):
            self.carry(stride)
            pos = S + (pos - S + stride) % (L - S)
            idx = pos[None, :] - S + np.arange(S)[:, None]          # target_t = data[pos - S + t], x_t = the byte before it
            self.set_window(data[(idx - 1) % L].astype(np.int64), data[idx % L].astype(np.int64))
            losses[it] = self.forward()
            self.backward()
            self.adagrad(lr)
        return losses, time.perf_counter() - t0
