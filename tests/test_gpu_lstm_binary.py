"""The drop-in `lstm` host program (eigen_lstm_b200/csrc/lstm_main.cc) against the stdout of the reference program.

tests/golden/ref_lstm_cc_run.json is what the unmodified R/lstm.cc printed (built against oracle/eigen_shim, see
tests/golden/make_ref_run.py) for seed 1234 on the first 3000 bytes of alice29.txt.  `lstm --seed 1234` seeds its
generators in the same order, so it starts from the SAME weights and state and must print the same text structure
(R/lstm.cc:274-291, 352-356, 398) and — the fp32 trajectory being chaotic in its last bits — epoch
losses that agree to a few percent rather than to the digit."""
import json
import os
import re
import subprocess
import tempfile

import pytest

from tests.conftest import GOLDEN, ROOT

pytestmark = pytest.mark.gpu
FIX = json.load(open(os.path.join(GOLDEN, "ref_lstm_cc_run.json")))
BIN = os.path.join(ROOT, "eigen_lstm_b200", "lstm")


def test_lstm_binary_prints_what_the_reference_program_prints(tmp_path, alice):
    assert os.path.exists(BIN), "eigen_lstm_b200/lstm is not built (python -m eigen_lstm_b200.build)"
    (tmp_path / "alice29.txt").write_bytes(alice[: FIX["corpus_bytes"]])
    epochs = 3
    out = subprocess.run([BIN, "--seed", str(FIX["seed"]), "--epochs", str(epochs)], cwd=tmp_path, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, timeout=300)
    assert out.returncode == 0, out.stderr.decode(errors="replace")
    so = out.stdout
    # line 1: rawread's report (R/lstm.cc:398)
    assert so.split(b"\n", 1)[0].decode() == FIX["read_line"]
    # progress fields of epoch 1: same count, same text (:274-279)
    first = so[: so.find(b"====")]
    assert [m.decode() for m in re.findall(rb"([ 0-9.]{7})%\r", first)] == FIX["progress_fields"]
    # epoch lines: same format as the reference's, epoch counter out of the same total, loss close (:284-291)
    lines = [m.decode() for m in re.findall(rb"Epoch \d+/\d+, t = [^\n]*", so)]
    assert len(lines) == epochs
    pat = r"Epoch (\d+)/(\d+), t = \d+\.\d{3} s, est GFLOP/s = \d+\.\d{3}, avg loss = (\d+\.\d{3}) bits/char"
    for e, (line, ref_line) in enumerate(zip(lines, FIX["epoch_lines"])):
        m, mr = re.fullmatch(pat, line), re.fullmatch(pat, ref_line)
        assert m and mr, (line, ref_line)
        assert int(m.group(1)) == e + 1
        got, want = float(m.group(3)), float(mr.group(3))
        print(f"epoch {e + 1}: lstm {got:.3f} reference {want:.3f} bits/char")
        assert abs(got - want) / want < 0.08, (e, got, want)
    # the separator and the sample block (:284, :352-356): 1000 characters between the reference's markers
    assert so.count(b"\n" + b"=" * 84 + b"\n") == epochs
    gen = re.findall(rb"\n\n\*{12} Generated text \|(.*?)\| Generated text END \*{12}\n", so, flags=re.S)
    assert len(gen) == epochs and all(len(g) == 1000 for g in gen)


def test_lstm_binary_heldout_eval_results_log_and_checkpoint_roundtrip(tmp_path, enwik6):
    """The long-run workflow of the last snapshot (OV/lstm_eigen_class_CUDA/lstm.cc:73-86,188-238): train/test split,
    timed held-out evaluation, 5-column results log, text checkpoint + sample file; then --load resumes from the
    checkpoint (io.h:16-81) and scores the same held-out bits/char the log recorded (to the 6 digits the text keeps)."""
    import numpy as np
    (tmp_path / "corpus.txt").write_bytes(enwik6[:20000])
    prefix = str(tmp_path / "run")
    common = [BIN, "--file", "corpus.txt", "--hidden", "48", "--seq", "12", "--batch", "8", "--stride", "11", "--seed", "3",
              "--forget-bias", "1", "--state-std", "0", "--train-percent", "95"]
    out = subprocess.run(common + ["--epochs", "2", "--max-iters", "3000", "--test-every", "0.05", "--save", prefix],
                         cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert out.returncode == 0, out.stderr.decode(errors="replace")
    so = out.stdout.decode(errors="replace")
    assert "Train set size: 19000, Test set size: 1000, Total: 20000" in so
    assert "Train error: " in so and ", Test error: " in so
    rows = np.loadtxt(prefix + ".txt", ndmin=2)
    assert rows.shape[1] == 5 and rows.shape[0] >= 2
    assert list(rows[:, 0]) == list(range(rows.shape[0]))            # running index
    assert np.all(rows[:, 1] >= 0.05) and np.all(rows[:, 4] > 0)      # seconds since the last evaluation, GFlOP/s
    assert rows[-1, 3] < rows[0, 3] < 8.5                             # held-out bits/char falls while training
    for name, shape in [("W", (192, 256)), ("U", (192, 48)), ("Why", (256, 48)), ("b", (192,)), ("by", (256,))]:
        assert np.loadtxt(f"{prefix}_{name}.txt").shape == shape
    assert os.path.getsize(prefix + "_sample.txt") == 5000
    # resume: zero further iterations, immediate evaluation -> the held-out error of the saved weights
    out2 = subprocess.run(common + ["--epochs", "1", "--max-iters", "1", "--lr", "0", "--test-every", "0.000001", "--load", prefix],
                          cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert out2.returncode == 0, out2.stderr.decode(errors="replace")
    m = re.search(r"Test error: ([0-9.eE+-]+)", out2.stdout.decode(errors="replace"))
    assert m, out2.stdout.decode(errors="replace")
    # the checkpoint was written at the LAST evaluation or at the end of the epoch (later): at least as good as the last row
    assert float(m.group(1)) < rows[-1, 3] + 0.3        # (an unloaded, untrained model scores 8.0)


def test_reference_text_checkpoint_through_the_product_reader(golden_dir):
    """the reference's own checkpoint files models/enwik5_test_{W,U,Why,b,by}.txt (written by its
    Parameters::save_to_disk, OV/lstm_eigen_class_CUDA/io.h:16-32), committed byte for byte under tests/golden/, are parsed by
    lstm_load_text_ckpt — not by numpy — and evaluated on the GPU with the test() recipe: the reference logged 3.24396."""
    import numpy as np
    import eigen_lstm_b200 as el
    z = np.load(os.path.join(golden_dir, "enwik5_test.npz"))
    g = el.LSTM(256, 32, 3, 1)
    g.load_from_disk(os.path.join(golden_dir, "enwik5_test"))
    bpc = g.test(z["test_bytes"].tobytes())
    assert abs(bpc - float(z["logged_test_bpc"])) < 2e-4, bpc
    for name, got in zip(["W", "U", "b", "Why", "by"], g.params()):
        assert np.array_equal(got, z[name].reshape(got.shape)), name      # same numbers numpy parsed from the same files
    # and through the drop-in binary: `lstm --load` + a held-out evaluation prints the same test error
    with tempfile.TemporaryDirectory() as d:
        corpus = os.path.join(d, "enwik5_like.txt")
        text = open(os.path.join(golden_dir, "enwik6.txt"), "rb").read()[:100000]
        open(corpus, "wb").write(text)
        out = subprocess.run([BIN, "--file", corpus, "--hidden", "32", "--seq", "25", "--batch", "16", "--lr", "0", "--epochs", "1",
                              "--max-iters", "300", "--train-percent", "99", "--test-every", "0.001", "--sample", "0", "--seed", "1",
                              "--load", os.path.join(golden_dir, "enwik5_test")], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        errs = [float(l.split("Test error: ")[1]) for l in out.stdout.splitlines() if "Test error: " in l]
        assert errs and abs(errs[0] - 3.24396) < 2e-4, out.stdout[-400:]


def test_last_snapshot_workflow_flags(golden_dir):
    """lr = 0 while the window warms up (OV/lstm_eigen_class_CUDA/lstm.cc:364-367), the eta progress line (:239-268), --clip and the
    class_batch loss report through the binary."""
    with tempfile.TemporaryDirectory() as d:
        corpus = os.path.join(d, "c.txt")
        open(corpus, "wb").write(open(os.path.join(golden_dir, "enwik6_head.bin"), "rb").read()[:20000])
        base = [BIN, "--file", corpus, "--hidden", "32", "--seq", "8", "--batch", "4", "--epochs", "1", "--sample", "0", "--seed", "3",
                "--stride", "7", "--max-iters", "400"]
        a = subprocess.run(base + ["--save", os.path.join(d, "a"), "--lr-warmup", "400"], capture_output=True, text=True, timeout=120)
        b = subprocess.run(base + ["--save", os.path.join(d, "b"), "--lr", "0"], capture_output=True, text=True, timeout=120)
        assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
        for n in ("W", "U", "Why", "b", "by"):                               # 400 warm-up iterations = 400 iterations at lr 0
            assert open(os.path.join(d, f"a_{n}.txt")).read() == open(os.path.join(d, f"b_{n}.txt")).read(), n
        c = subprocess.run(base + ["--progress", "eta", "--clip", "0.5", "--loss", "last-ln", "--softmax-shift", "global"],
                           capture_output=True, text=True, timeout=120)
        assert c.returncode == 0, c.stderr
        import re
        assert re.search(r"\[Epoch 1/1\]\s+\d+\.\d\d%\s+\(eta\s+\d+ h \d\d m \d\d s\)\s+loss = \d+\.\d{6}\s+\d+\.\d\d GFlOP/s", c.stdout), c.stdout[:300]
