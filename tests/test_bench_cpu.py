"""bench.py pieces that run without a GPU: the flop model of SURVEY §8d, the synthetic corpus, and the CPU
(`--impl reference`) arm printing the contract's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np

from tests.conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_flop_model_matches_survey():
    # cfg1 458 752 / 196 608; cfg2 9 175 040 / 7 077 888; cfg4 112 197 632 / 103 809 024
    for n, dense, alg in [(64, 458752, 196608), (512, 9175040, 7077888), (2048, 112197632, 103809024)]:
        f = bench.flops_per_charstep(n)
        assert f["dense"] == dense and f["alg"] == alg


def test_synthetic_text_is_deterministic_and_follows_the_histogram():
    a, b = bench.synthetic_text(200000, seed=0), bench.synthetic_text(200000, seed=0)
    assert np.array_equal(a, b) and a.dtype == np.uint8
    hist = np.load(os.path.join(ROOT, "tests", "golden", "enwik6_hist.npy")).astype(np.float64)
    emp = np.bincount(a, minlength=256) / a.size
    assert np.abs(emp - hist / hist.sum()).max() < 5e-3
    assert (emp[hist == 0] == 0).all()


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "3", "--cpu-seconds", "0.3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "chars/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_non_rank0_reference_arm_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_uses_the_blas_port_for_large_configurations():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2",
                          "--steps", "1", "--warmup", "3", "--cpu-seconds", "0.5"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    cb = line["cpu_baseline"]
    assert line["impl"] == "reference" and line["value"] > 0 and cb["value"] == line["value"]
    assert cb["kind"] == "port" and "oracle_blas" in cb["sample"] and "64 streams" in cb["sample"] and cb["cores"] >= 1
