"""Multi-threaded CPU restatement of the batched reference path on torch's CPU backend — TEST INFRASTRUCTURE / CPU BASELINE,
NOT PRODUCT CODE.

Same statements, in the same order, as oracle/oracle_blas.py (the structure of OV/lstm_eigen_BLAS/lstm.cc:228-357: every
contraction one BLAS GEMM with dense one-hot inputs, element-wise work between them), but on torch CPU tensors: the GEMMs go
to torch's bundled multi-threaded BLAS (MKL / OpenBLAS) and — unlike numpy — the element-wise statements (logistic, tanh, the
gate gradients, Adagrad) are themselves multi-threaded and vectorised.  This is the strongest CPU stand-in for the
`lstm_eigen_BLAS` program this image can offer (Eigen and OpenBLAS are not installed); bench.py's CPU arm uses it for the
large configurations with an explicit thread count so that the number does not depend on how the process was launched.
Checked against oracle_blas.py / the C++ oracle in tests/test_oracle_blas.py.  Only tests/ and bench.py's cpu_baseline /
--impl reference leg may import it.
"""
import time

import numpy as np
import torch

EPS = 1e-10  # OV/lstm_eigen_BLAS/lstm.cc:26


class TorchOracle:
    def __init__(self, M, N, S, B, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.threads = torch.get_num_threads()
        self.M, self.N, self.S, self.B = M, N, S, B
        z = lambda r, c: torch.zeros((r, c), dtype=torch.float32)
        self.W, self.U, self.b, self.Why, self.by = z(4 * N, M), z(4 * N, N), z(4 * N, 1), z(M, N), z(M, 1)
        self.m = [z(4 * N, M), z(4 * N, N), z(4 * N, 1), z(M, N), z(M, 1)]
        self.h = [z(N, B) for _ in range(S)]
        self.c = [z(N, B) for _ in range(S)]
        self.g = [z(4 * N, B) for _ in range(S)]
        self.x = [z(M, B) for _ in range(S)]
        self.target = [z(M, B) for _ in range(S)]
        self.probs = [z(M, B) for _ in range(S)]
        self.grads = None

    def set_params(self, params):
        self.W, self.U, self.b, self.Why, self.by = [torch.from_numpy(np.ascontiguousarray(np.asarray(p, dtype=np.float32))).clone()
                                                     for p in params]

    def params(self):
        return [self.W, self.U, self.b, self.Why, self.by]

    def set_window(self, x_idx, t_idx):
        x_idx = torch.as_tensor(np.asarray(x_idx).reshape(self.S, self.B), dtype=torch.int64)
        t_idx = torch.as_tensor(np.asarray(t_idx).reshape(self.S, self.B), dtype=torch.int64)
        cols = torch.arange(self.B)
        for t in range(self.S):
            for dst, idx in ((self.x[t], x_idx[t]), (self.target[t], t_idx[t])):
                dst.zero_()
                ok = idx >= 0
                dst[idx[ok], cols[ok]] = 1.0

    def forward(self, t0=1, t1=None):
        N, B = self.N, self.B
        loss = 0.0
        for t in range(t0, t1 or self.S):
            g = self.W @ self.x[t]                              # :229-231 (dense one-hot product, like the reference)
            g.addmm_(self.U, self.h[t - 1])                     # :232
            g += self.b                                         # :233
            g[:3 * N].sigmoid_()                                # :238
            g[3 * N:].tanh_()                                   # :240
            c = torch.tanh(g[:N] * g[3 * N:] + g[2 * N:3 * N] * self.c[t - 1])   # :242-246
            h = g[N:2 * N] * c                                  # :249
            y = self.Why @ h                                    # :254
            y += self.by
            p = torch.exp(y)                                    # :261
            p /= p.sum(dim=0, keepdim=True)                     # :262-263
            self.g[t], self.c[t], self.h[t], self.probs[t] = g, c, h, p
            s = torch.sum(-torch.log2(p) * self.target[t])      # :267 (log of ALL entries, like the reference)
            loss += float(s) / B                                # :270
        return loss

    def backward(self, t0=1, t1=None):
        N = self.N
        dWhy, dby = torch.zeros_like(self.Why), torch.zeros_like(self.by)
        dU, dW, db = torch.zeros_like(self.U), torch.zeros_like(self.W), torch.zeros_like(self.b)
        dhnext = torch.zeros_like(self.h[0])
        dcnext = torch.zeros_like(self.c[0])
        for t in range((t1 or self.S) - 1, t0 - 1, -1):
            g, c = self.g[t], self.c[t]
            dy = self.probs[t] - self.target[t]                      # :290
            dWhy.addmm_(dy, self.h[t].T)                             # :292
            dby += dy.sum(dim=1, keepdim=True)                       # :296
            dh = torch.addmm(dhnext, self.Why.T, dy)                 # :298-303
            dc = (dh * g[N:2 * N] + dcnext) * (1.0 - c * c)          # :308-310
            dg = torch.empty_like(g)
            dg[N:2 * N] = dh * c                                     # :313
            dg[:N] = dc * g[3 * N:]                                  # :314
            dg[2 * N:3 * N] = dc * self.c[t - 1]                     # :315
            dg[3 * N:] = dc * g[:N]                                  # :316
            dg[:3 * N] *= g[:3 * N] * (1.0 - g[:3 * N])              # :319-320
            dg[3 * N:] *= 1.0 - g[3 * N:] * g[3 * N:]                # :323-324
            dU.addmm_(dg, self.h[t - 1].T)                           # :327
            dW.addmm_(dg, self.x[t].T)                               # :328 (dense, like the reference)
            db += dg.sum(dim=1, keepdim=True)                        # :335
            dhnext = self.U.T @ dg                                   # :339-340
            dcnext = dc * g[2 * N:3 * N]                             # :343
        self.grads = [dW, dU, db, dWhy, dby]

    def adagrad(self, lr):
        for p, d, m in zip(self.params(), self.grads, self.m):       # :346-357
            m.addcmul_(d, d)
            p.sub_(lr * (d / torch.sqrt(m + EPS)))

    def carry(self, stride):
        stride = min(stride, self.S - 1)
        self.h[0] = self.h[stride].clone()
        self.c[0] = self.c[stride].clone()


def time_parts(M, N, B, Ts, params, text, threads=None, repeats=1):
    """Seconds per TIMESTEP of B streams (forward + backward, everything the reference executes per timestep) and seconds per
    Adagrad sweep, measured separately on a window of Ts timesteps — so that a full window of T timesteps can be composed as
    T * t_step + t_adagrad without running all T on the CPU."""
    q = TorchOracle(M, N, Ts + 1, B, threads)
    q.set_params(params)
    data = np.frombuffer(bytes(text), dtype=np.uint8)
    pos = (Ts + 1 + np.arange(B) * 997) % (data.size - Ts - 2)
    idx = pos[None, :] + np.arange(Ts + 1)[:, None]
    q.set_window(data[idx].astype(np.int64), data[idx + 1].astype(np.int64))
    q.forward(); q.backward(); q.adagrad(0.002)               # warm-up (thread pools, allocator)
    best_step, best_ada = float("inf"), float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        q.forward(); q.backward()
        t1 = time.perf_counter()
        q.adagrad(0.002)
        t2 = time.perf_counter()
        best_step = min(best_step, (t1 - t0) / Ts)
        best_ada = min(best_ada, t2 - t1)
    return best_step, best_ada, q.threads
