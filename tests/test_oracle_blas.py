"""The BLAS-backed CPU baseline (oracle/oracle_blas.py: the structure of OV/lstm_eigen_BLAS/lstm.cc on numpy's OpenBLAS)
against the C++ oracle: same windows and weights in, same losses, gradients and updated weights out, up to float32
summation order."""
import numpy as np

from oracle import oracle as orc
from oracle import oracle_blas as ob
from tests.util import rel_err


def test_blas_oracle_matches_the_cpp_oracle(enwik6):
    M, N, S, B = 256, 24, 6, 5
    params = orc.init_params(M, N, seed=11, sd=0.08, forget_bias=1.0)
    o = orc.Oracle(M, N, S, B, "f32")
    o.set_options(dense_onehot=1)
    o.set_params(params)
    o.set_positions([S + 40 * b for b in range(B)])
    q = ob.BlasOracle(M, N, S, B)
    q.set_params(params)
    rng = np.random.default_rng(0)
    h0 = rng.normal(0, 0.1, (N, B)).astype(np.float32)
    c0 = rng.normal(0, 0.1, (N, B)).astype(np.float32)
    o.set_state("h", 0, h0); o.set_state("c", 0, c0)
    q.h[0][...] = h0; q.c[0][...] = c0
    for it in range(6):
        if it:
            o.carry(2); q.carry(2)
        o.advance(enwik6, 2)
        x, t = o.window()
        q.set_window(x, t)
        lo, lq = o.forward(), q.forward()
        assert abs(lo - lq) <= 2e-6 * abs(lo), (it, lo, lq)
        o.backward(); q.backward()
        for go, gq in zip(o.grads(), q.grads):
            assert rel_err(gq, go) < 2e-5
        o.adagrad(0.1); q.adagrad(0.1)
    for po, pq in zip(o.params(), q.params()):
        assert rel_err(pq, po) < 1e-3      # six Adagrad steps of +-lr amplify the summation-order noise of the gradients
    for t in range(1, S):
        assert rel_err(q.h[t], o.state("h", t)) < 1e-4 and rel_err(q.c[t], o.state("c", t)) < 1e-4


def test_blas_oracle_train_windows_runs_and_learns(enwik6):
    q = ob.BlasOracle(256, 32, 9, 8)
    q.set_params(orc.init_params(256, 32, seed=3, sd=0.01, forget_bias=1.0))
    losses, secs = q.train_windows(enwik6[:20000], [9 + 900 * b for b in range(8)], 60, 8, 0.1)
    assert secs > 0 and np.all(np.isfinite(losses))
    assert losses[-10:].mean() < 0.8 * losses[:3].mean()      # 8 bits/char at the start, falling


def test_torch_oracle_matches_the_blas_oracle(enwik6):
    """oracle/oracle_torch.py (bench.py's multi-threaded CPU arm) against oracle_blas.py: same statements, same numbers."""
    from oracle import oracle_torch as ot
    M, N, S, B = 256, 24, 6, 5
    params = orc.init_params(M, N, seed=11, sd=0.08, forget_bias=1.0)
    q = ob.BlasOracle(M, N, S, B); q.set_params(params)
    p = ot.TorchOracle(M, N, S, B, threads=2); p.set_params(params)
    rng = np.random.default_rng(0)
    h0 = rng.normal(0, 0.1, (N, B)).astype(np.float32); c0 = rng.normal(0, 0.1, (N, B)).astype(np.float32)
    q.h[0][...] = h0; q.c[0][...] = c0
    import torch
    p.h[0] = torch.from_numpy(h0.copy()); p.c[0] = torch.from_numpy(c0.copy())
    d = np.frombuffer(enwik6, dtype=np.uint8)
    for it in range(4):
        idx = (100 * np.arange(B))[None, :] + it * (S - 1) + np.arange(S)[:, None]
        x, t = d[idx].astype(np.int64), d[idx + 1].astype(np.int64)
        q.set_window(x, t); p.set_window(x, t)
        lq, lp = q.forward(), p.forward()
        assert abs(lq - lp) <= 2e-6 * abs(lq), (it, lq, lp)
        q.backward(); p.backward()
        for gq, gp in zip(q.grads, p.grads):
            assert rel_err(gp.numpy(), gq) < 2e-5
        q.adagrad(0.1); p.adagrad(0.1); q.carry(S - 1); p.carry(S - 1)
    step, ada, th = ot.time_parts(M, N, B, 3, params, enwik6[:5000], threads=2)
    assert step > 0 and ada > 0 and th == 2
