"""The one-kernel training path (eigen_lstm_b200/csrc/train_small.cu, adagrad_batch) skips the reference's double-precision
`m + 1e-10` (R/lstm.cc:25,46-48: eps is added in double, the sum is rounded to float for sqrtf) whenever m >= m_exact, claiming
that the sum then rounds back to m.  That is a statement about IEEE arithmetic, checkable here without a GPU: the threshold the
launcher computes (smallest power of two whose half ulp exceeds eps) is 2^-9, the claim holds for every float at or above it
— and fails right below it, so the threshold is not conservative by more than one binade."""
import numpy as np

EPS = 1e-10


def launcher_threshold(eps):
    # launch_train_small(): smallest power of two 2^e with 2^(e-24) > eps
    for e in range(-100, 25):
        if np.ldexp(1.0, e - 24) > eps * 1.0000001:
            return np.float32(np.ldexp(1.0, e))
    return np.float32(np.inf)


def detour(m):
    return (m.astype(np.float64) + EPS).astype(np.float32)


def test_threshold_is_two_to_the_minus_nine():
    assert launcher_threshold(EPS) == np.float32(2.0 ** -9)


def test_sum_rounds_back_to_m_at_and_above_the_threshold():
    t = launcher_threshold(EPS)
    rng = np.random.default_rng(0)
    # log-uniform magnitudes from the threshold up to 2^40, plus every power of two in the range and its neighbours
    m = np.exp(rng.uniform(np.log(float(t)), np.log(2.0 ** 40), 4_000_000)).astype(np.float32)
    m = m[m >= t]
    edges = []
    for e in range(-9, 41):
        p = np.float32(2.0 ** e)
        edges += [p, np.nextafter(p, np.float32(np.inf)), np.nextafter(p, np.float32(0))]
    edges = np.array([x for x in edges if x >= t], dtype=np.float32)
    # the whole first binade above the threshold, exhaustively (2^23 floats)
    first = (np.arange(1 << 23, dtype=np.uint32) + np.float32(t).view(np.uint32)).view(np.float32)
    for arr in (m, edges, first):
        assert np.array_equal(detour(arr), arr)


def test_the_detour_matters_below_the_threshold():
    t = launcher_threshold(EPS)
    below = (np.arange(1 << 23, dtype=np.uint32) + np.float32(t / 2).view(np.uint32)).view(np.float32)   # the binade below
    changed = detour(below) != below
    assert changed.all()            # there an ulp is 2^-33 = 1.16e-10: 1e-10 is more than half of it, EVERY sum moves to the next float
