"""Pin the oracle against the reference's own known-answer artefacts."""
import os

import numpy as np

from oracle import oracle as orc


def _fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "enwik5_test.npz"))
    return [z["W"], z["U"], z["b"], z["Why"], z["by"]], z["test_bytes"].tobytes(), float(z["logged_test_bpc"])


def test_enwik5_known_answer_f32(golden_dir):
    """test() recipe of OV/lstm_eigen_class_CUDA/lstm.cc:661-720 on models/enwik5_test_* gives the
    logged 3.24396 bits/char (models/enwik5_test.txt:1).  The checkpoint text keeps 6 significant
    digits, which moves the result by < 2e-4."""
    params, test_bytes, logged = _fixture(golden_dir)
    o = orc.Oracle(256, 32, 2, 1, "f32")
    o.set_params(params)
    bpc = o.eval_bpc(test_bytes)
    assert abs(bpc - logged) < 2e-4, (bpc, logged)


def test_enwik5_known_answer_f64(golden_dir):
    params, test_bytes, logged = _fixture(golden_dir)
    o = orc.Oracle(256, 32, 2, 1, "f64")
    o.set_params(params)
    bpc = o.eval_bpc(test_bytes)
    assert abs(bpc - logged) < 2e-4, (bpc, logged)


def test_untrained_bpc_is_8(enwik6):
    """Untrained model: 7.99998 bits/char (OV/lstm_eigen_class_batch/enwik8_n32_s3_initial_b4_test_1pc.txt:1)."""
    o = orc.Oracle(256, 32, 3, 4, "f32")
    o.set_params(orc.init_params(256, 32, seed=1, forget_bias=1.0))
    bpc = o.eval_bpc(enwik6[:3000])
    assert abs(bpc - 8.0) < 5e-3, bpc


def test_forward_step_matches_eval_recipe(golden_dir):
    """The training forward (R/lstm.cc:173-209) and test() share the cell: with h0=c0=0 and a full
    window, the summed training loss over t equals the eval recipe's surprisals."""
    params, test_bytes, _ = _fixture(golden_dir)
    S = 9
    o = orc.Oracle(256, 32, S, 1, "f32")
    o.set_params(params)
    d = np.frombuffer(test_bytes, dtype=np.uint8).astype(np.int32)
    x = np.full((S, 1), -1, np.int32)
    t = np.full((S, 1), -1, np.int32)
    x[1:, 0] = d[0:S - 1]
    t[1:, 0] = d[1:S]
    o.set_window(x, t)
    loss = o.forward()
    o2 = orc.Oracle(256, 32, 2, 1, "f32")
    o2.set_params(params)
    ref = o2.eval_bpc(test_bytes[:S]) * (S - 1)
    assert abs(loss - ref) < 1e-4 * ref


def test_window_semantics_match_reference_shift(alice):
    """R/lstm.cc:155-170: after the warm-up, x_t = data[i-S+t], target_t = data[i-S+t+1]; before it the
    not-yet-filled columns are all-zero (-1)."""
    S = 4
    o = orc.Oracle(256, 8, S, 1, "f32")
    o.set_positions([S])
    d = np.frombuffer(alice, dtype=np.uint8)
    o.advance(alice, 1)  # iteration i = S
    x, t = o.window()
    assert t[S - 1, 0] == d[S] and (t[:S - 1] == -1).all() and (x == -1).all()
    o.advance(alice, 1)  # i = S + 1
    x, t = o.window()
    assert t[S - 1, 0] == d[S + 1] and t[S - 2, 0] == d[S] and x[S - 1, 0] == d[S]
    for _ in range(10):
        o.advance(alice, 1)
    i = S + 11
    x, t = o.window()
    for tt in range(1, S):
        assert x[tt, 0] == d[i - S + tt] and t[tt, 0] == d[i - S + tt + 1]


def test_dense_onehot_equals_gather(alice):
    """W*x with one-hot x (R/lstm.cc:176) is exactly column x of W; dW += dg*x^T (:251) is a column add."""
    M, N, S, B = 256, 16, 5, 3
    params = orc.init_params(M, N, seed=3, sd=0.1)
    res = []
    for dense in (0, 1):
        o = orc.Oracle(M, N, S, B, "f32")
        o.set_options(dense_onehot=dense)
        o.set_params(params)
        o.set_positions([5, 100, 300])
        losses, _ = o.train(alice, 6, stride=1, lr=0.1)
        res.append((losses, o.params()))
    assert np.array_equal(res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        assert np.array_equal(a, b)
