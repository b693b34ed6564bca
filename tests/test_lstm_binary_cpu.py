"""Host-side behaviour of the drop-in `lstm` program that needs no GPU: corpus reading (rawread, R/lstm.cc:382-420) and
the loud failure without a CUDA device (there is no CPU fallback)."""
import os
import subprocess

import pytest

from tests.conftest import ROOT

BIN = os.path.join(ROOT, "eigen_lstm_b200", "lstm")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(not os.path.exists(BIN), reason="lstm binary not built")
def test_multi_corpus_reading_and_no_cpu_fallback(tmp_path):
    d = tmp_path / "calgary"
    d.mkdir()
    (d / "b.txt").write_bytes(b"B" * 30)
    (d / "a.txt").write_bytes(b"A" * 20)
    (tmp_path / "single.txt").write_bytes(b"xyz" * 5)
    out = subprocess.run([BIN, "--file", str(d), "--file", str(tmp_path / "single.txt"), "--file", str(tmp_path / "missing.txt"),
                          "--epochs", "0"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    lines = out.stdout.decode().splitlines()
    # a directory contributes its files in name order; every file is announced the way rawread announces it (:398,416)
    assert lines[0] == f"Read 20 bytes ({d}/a.txt)"
    assert lines[1] == f"Read 30 bytes ({d}/b.txt)"
    assert lines[2] == f"Read 15 bytes ({tmp_path}/single.txt)"
    assert lines[3] == f"fopen error: ({tmp_path}/missing.txt)"
    if not _has_gpu():
        assert out.returncode == 1
        assert "lstm_create failed" in out.stderr.decode() and "no CPU fallback" in out.stderr.decode()


@pytest.mark.skipif(not os.path.exists(BIN), reason="lstm binary not built")
def test_empty_and_missing_default_file(tmp_path):
    # no arguments: the reference's default corpus name; a missing file prints rawread's message and the program ends quietly
    out = subprocess.run([BIN], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert out.stdout.decode().splitlines()[0] == "fopen error: (alice29.txt)"
    assert out.returncode == 0
    (tmp_path / "alice29.txt").write_bytes(b"")
    out = subprocess.run([BIN], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert out.stdout.decode().splitlines()[0] == "Empty file! (alice29.txt)"
    assert out.returncode == 0
