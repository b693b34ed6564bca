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


def test_reference_arm_uses_the_multithreaded_blas_port_for_large_configurations():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2",
                          "--steps", "1", "--warmup", "3", "--cpu-seconds", "0.5"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    cb = line["cpu_baseline"]
    assert line["impl"] == "reference" and line["value"] > 0 and cb["value"] == line["value"]
    assert cb["kind"] == "port" and "oracle_torch" in cb["sample"] and "64 streams" in cb["sample"] and cb["cores"] >= 1
    # a full window is composed from separately timed parts: T x t_timestep + one Adagrad sweep
    assert abs(cb["value"] - 64 * 100 / (100 * cb["seconds_per_timestep"] + cb["seconds_per_adagrad_sweep"])) < 1e-6 * cb["value"]
    assert cb["gflops_dense"] > 0


def test_reference_arm_thread_count_does_not_depend_on_the_launcher():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU arm sets its thread count explicitly."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2", "--gpus", "2",
                          "--steps", "1", "--warmup", "3", "--cpu-seconds", "0.3"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-500:]
    cb = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
    assert cb["cores"] == len(os.sched_getaffinity(0))
