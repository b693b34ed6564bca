#!/usr/bin/env python
"""bench.py — training chars/sec (fwd + BPTT + Adagrad) of the B200-native char-LSTM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg4|cfg3|cfg2|cfg1] [--dtype bf16|f32]

One JSON line on stdout (rank 0).  A "step" is one training iteration over one window of every
stream: B streams x T = S-1 timesteps, stride T (non-overlapping truncated BPTT, precedent
OV/lstm_eigen_class_batch/lstm_segment.cc:130,187), so one step consumes B*T new characters.

  value      chars/s with the corpus, windows and weights resident in HBM (lstm_train_text)
  e2e        chars/s through the public per-step API with HOST index buffers: pinned host ->
             device copy of the step's window and a device -> host read of its loss every step
  roofline   the dominant kernel's achieved TFLOP/s (live CUDA-event phase timing) vs the measured
             dense bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (port of R/lstm.cc / OV/lstm_eigen_BLAS; the reference needs Eigen and
             OpenBLAS and cannot be built here) on the box's host cores, on a bounded sample of the same
             workload: BLAS-backed (numpy's OpenBLAS) for the large configurations, the single-thread
             C++ port for the reference's own tiny default

--impl reference runs only the CPU arm (the reference's own algorithm on host cores).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3]: the configuration the chars/sec metric is quoted on
    "cfg4": dict(N=2048, B=256, S=257, desc="synthetic char-LSTM H=2048, batch 256 per GPU, seq 256"),
    "cfg3": dict(N=1024, B=128, S=101, desc="char-LSTM H=1024, batch 128, seq 100"),
    "cfg2": dict(N=512, B=64, S=101, desc="class_batch-style char-LSTM H=512, batch 64, seq 100"),
    "cfg1": dict(N=64, B=1, S=3, desc="lstm.cc default char-LSTM H=64, batch 1, S=3"),
    # BASELINE.json configs[4]: a different metric (sampled chars/s); handled by bench_sampling()
    "cfg5": dict(N=1024, B=1, S=3, desc="latency-bound sampling: 1M chars at batch 1 from an H=1024 model (persistent recurrent kernel)"),
}
M = 256
LR = 0.002   # Adagrad's first steps move every weight by +-lr whatever the gradient size; 0.1 (R/lstm.cc:59) saturates a 2048-wide net


def flops_per_charstep(N):
    """GEMM flops per (sequence, timestep).  dense = what the reference executes;
    alg = one-hot aware (W*x and dW are a gather / scatter)."""
    return dict(dense=22 * N * M + 24 * N * N, alg=24 * N * N + 6 * M * N)


def synthetic_text(nbytes, seed=0):
    """i.i.d. bytes from the empirical byte histogram of enwik6."""
    hist = np.load(os.path.join(ROOT, "tests", "golden", "enwik6_hist.npy")).astype(np.float64)
    rng = np.random.default_rng(seed)
    return rng.choice(256, size=nbytes, p=hist / hist.sum()).astype(np.uint8)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's algorithm) on host cores, bounded sample
# -------------------------------------------------------------------------------------------------
def host_threads():
    """Threads the CPU arm uses: every core the process may run on — set EXPLICITLY, so that the number does not depend on how
    the process was launched (torch.distributed.run exports OMP_NUM_THREADS=1 to its workers)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_arm_torch(cfg, text, target_seconds, steps=1, warmup=0):
    """Large configurations: the structure of OV/lstm_eigen_BLAS/lstm.cc (every contraction one multi-threaded BLAS GEMM with
    dense one-hot inputs, like the reference; element-wise work between them) on torch's CPU backend — oracle/oracle_torch.py —
    with all B streams.  A full window is COMPOSED from separately timed parts, T * t_timestep + t_adagrad: the per-timestep cost
    is measured over a bounded number of timesteps, the Adagrad sweep once, and one sweep is charged per window of T timesteps
    exactly as in the real iteration."""
    from oracle import oracle as orc
    from oracle import oracle_torch as ot
    N, S, B = cfg["N"], cfg["S"], cfg["B"]
    T = S - 1
    th = host_threads()
    params = orc.init_params(M, N, seed=0, sd=0.01)
    t_step, t_ada, th = ot.time_parts(M, N, B, 2, params, text, threads=th)          # probe
    Ts = int(max(2, min(T, target_seconds / max(t_step, 1e-9))))
    vals = []
    for _ in range(max(1, steps)):
        t_step, t_ada, th = ot.time_parts(M, N, B, Ts, params, text, threads=th)
        vals.append((t_step, t_ada))
    t_step = sum(v[0] for v in vals) / len(vals)
    t_ada = sum(v[1] for v in vals) / len(vals)
    window_s = T * t_step + t_ada
    fl = flops_per_charstep(N)["dense"] * B
    return dict(value=B * T / window_s, unit="chars/s", cores=th, kind="port",
                gflops_dense=fl / t_step / 1e9, seconds_per_timestep=t_step, seconds_per_adagrad_sweep=t_ada,
                sample=f"{B} streams x {Ts} timesteps timed (forward + BPTT), one Adagrad sweep timed separately; a full window "
                       f"composed as {T} x t_timestep + t_adagrad.  oracle/oracle_torch.py: the lstm_eigen_BLAS structure (every "
                       f"contraction one BLAS sgemm, dense one-hot products like the reference, multi-threaded element-wise math) "
                       f"on torch's CPU backend, {th} threads set explicitly"), window_s


def cpu_arm(cfg, text, target_seconds, steps=1, warmup=0):
    from oracle import oracle as orc
    N, S = cfg["N"], cfg["S"]
    if cfg["B"] * N >= 4096 and not os.environ.get("BENCH_CPU_PORT"):
        return cpu_arm_torch(cfg, text, target_seconds, steps, warmup)   # GEMM throughput decides: use the best BLAS present
    threads = 1                                       # tiny models: fork/join costs more than the loops (the reference is 1 thread)
    Bs = min(cfg["B"], threads)
    # the reference's own default (config 1): the single-thread C++ port compiled with the reference's flags, full windows
    iters = 2000
    o = orc.Oracle(M, N, S, Bs, "f32", mt=False, threads=1)
    o.set_options(dense_onehot=1)                     # the reference multiplies the one-hot densely (R/lstm.cc:176,251)
    o.set_params(orc.init_params(M, N, seed=0, sd=0.01))
    o.set_positions([S + i * 1000 for i in range(Bs)])
    _, secs = o.train(text, iters, stride=S - 1, lr=0.1)
    iters = int(max(200, min(2_000_000, iters * target_seconds / max(secs, 1e-9))))
    t = []
    for _ in range(max(1, steps)):
        _, secs = o.train(text, iters, stride=S - 1, lr=0.1)
        t.append(secs / iters)
    per_iter = sum(t) / len(t)
    return dict(value=Bs * (S - 1) / per_iter, unit="chars/s", cores=1, kind="port",
                gflops_dense=flops_per_charstep(N)["dense"] * Bs * (S - 1) / per_iter / 1e9,
                sample=f"{iters} full training iterations (window of {S - 1} timesteps, forward + BPTT + Adagrad) of the N={N} model, "
                       f"oracle/liblstm_oracle.so: the scalar port compiled like R/Makefile (g++ -O3, one thread, dense one-hot "
                       f"products like the reference)"), per_iter


def dp_check(el, dist, rank, world, local_rank, net, dtype_code):
    """Correctness of the data-parallel path, run before the timed region on every multi-GPU bench (the records then carry it):
      (1) `world` contexts with b streams each == ONE context with world*b streams (fp32: the exact same sums up to order),
          on a small shape, after 4 full training iterations incl. graph capture and replay;
      (2) at the BENCH shape: the gradients the library allreduced itself == the sum over ranks of the gradients of a second,
          communicator-less context fed the same local window (summed here with torch.distributed);
      (3) replicas bit-identical (min == max over ranks of every parameter) — checked again after the timed region."""
    import torch
    from eigen_lstm_b200 import dp
    out = {}
    dev = torch.device("cuda", local_rank)
    # (1) small-shape equivalence
    Ms, Ns, Ss, Bs, chunk = 256, 64, 6, 4, 700
    text = synthetic_text(chunk * Bs * world + 64, seed=3).tobytes()
    g = el.LSTM(Ms, Ns, Ss, Bs, device=local_rank, dtype=el.F32)
    g.dp_init(rank, world, dp.broadcast_unique_id(dist, el.dp_unique_id, rank))
    g.init_params(7, 0.05, 1.0)
    g.load_text(text); g.set_positions(dp.stream_positions(Bs * world, rank, world, Ss, chunk))
    g.train_text(4, stride=Ss - 1, lr=0.01)
    flat = np.concatenate([p.ravel(order="F") for p in g.params()])
    err = 0.0
    if rank == 0:
        one = el.LSTM(Ms, Ns, Ss, Bs * world, device=local_rank, dtype=el.F32)
        one.init_params(7, 0.05, 1.0)
        one.load_text(text); one.set_positions(dp.stream_positions(Bs * world, 0, 1, Ss, chunk))
        one.train_text(4, stride=Ss - 1, lr=0.01)
        ref = np.concatenate([p.ravel(order="F") for p in one.params()])
        err = float(np.max(np.abs(flat - ref)) / np.max(np.abs(ref)))
        one.close()
    out["f32_vs_one_gpu_with_concatenated_batch_max_rel_err"] = err
    t = torch.from_numpy(flat).to(dev)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["small_replicas_identical"] = bool(torch.equal(lo, hi))
    g.close()
    # (2) bench shape: library allreduce vs torch.distributed sum of local gradients (one forward + backward, no update)
    M_, N_, S_, B_ = net.M, net.N, net.S, net.B
    x, tg = net.window()
    if (x[1:] < 0).any():                      # the window has not been built yet: advance once
        net.train_text(1, stride=S_ - 1, lr=0.0, want_losses=False)
        x, tg = net.window()
    h0, c0 = net.get_state()
    net.forward(x, tg); net.backward(); net.sync()
    got = [a.copy() for a in net.grads()]
    loc = el.LSTM(M_, N_, S_, B_, device=local_rank, dtype=dtype_code)
    loc.set_params(net.params()); loc.set_state(h0, c0)
    loc.forward(x, tg); loc.backward()
    worst = 0.0
    for a, b in zip(got, loc.grads()):
        tb = torch.from_numpy(np.ascontiguousarray(b)).to(dev)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        want = tb.cpu().numpy()
        worst = max(worst, float(np.max(np.abs(a - want)) / max(float(np.max(np.abs(want))), 1e-30)))
    loc.close()
    net.set_state(h0, c0)
    tw = torch.tensor([worst], device=dev, dtype=torch.float64)
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    out["bench_shape_allreduced_grads_vs_sum_of_local_grads_max_rel_err"] = float(tw.item())
    out["ok"] = bool(out["small_replicas_identical"] and out["bench_shape_allreduced_grads_vs_sum_of_local_grads_max_rel_err"] < 1e-4
                     and (rank != 0 or err < 2e-4))
    return out


def replicas_identical(dist, net, local_rank):
    import torch
    flat = np.concatenate([p.ravel(order="F") for p in net.params()])
    t = torch.from_numpy(flat).to(torch.device("cuda", local_rank))
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool(torch.equal(lo, hi))


def bench_sampling(args, cfg):
    """BASELINE.json configs[4]: generate characters at batch 1 from an H=1024 model with the persistent batch-1 kernel (K9),
    and the companion held-out evaluation (test()).  One JSON line; the metric here is SAMPLED chars/s."""
    import torch
    import eigen_lstm_b200 as el
    N = cfg["N"]
    n = 1_000_000
    net = el.LSTM(M, N, 3, 1, device=0, dtype=el.F32)
    net.init_params(seed=0, std=0.05, forget_bias=0.0)
    stream = torch.cuda.ExternalStream(net.stream(), device=torch.device("cuda", 0))
    net.sample(20000, seed=1)                                    # warm-up
    sampler = ClockSampler(0)
    t = []
    for k in range(max(1, args.steps)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = net.sample(n, seed=2 + k)
        e1.record(stream)
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    clocks = sampler.stop()
    ms = sum(t) / len(t)
    text = synthetic_text(200_001, seed=5).tobytes()
    net.test(text[:20000])
    ev = []
    for _ in range(3):                                           # device time of the whole test() call, median of three
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        bpc = net.test(text)
        e1.record(stream)
        torch.cuda.synchronize()
        ev.append(e0.elapsed_time(e1) * 1e-3)
    ev_s = sorted(ev)[1]                                         # (the first full-length call also pays its allocations)
    pk = peaks()
    U_bytes = 4.0 * N * N * 4                                    # fp32 U resident in shared memory, read once per character
    line = {"metric": "sampled chars/sec at batch 1 (persistent recurrent kernel)", "value": n / (ms * 1e-3), "unit": "chars/s",
            "n_gpus": 1, "steps": args.steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg5: {cfg['desc']}", "N": N, "M": M, "chars_per_step": n,
                       "l2": "weights are resident in shared memory; per-step traffic is h(t) only"},
            "us_per_sampled_char": ms * 1e3 / n, "us_per_evaluated_char": ev_s * 1e6 / (len(text) - 1),
            "us_per_evaluated_char_runs": [x * 1e6 / (len(text) - 1) for x in ev], "eval_bits_per_char": bpc,
            "e2e": {"value": n / (ms * 1e-3), "unit": "chars/s", "h2d_bytes_per_step": 4 * n, "d2h_bytes_per_step": n,
                    "api": "lstm_sample: host uniforms -> device, sampled bytes -> host, inside the timed region"},
            "gpu_launches": int(args.steps), "clocks": clocks,
            "roofline": {"bound": "latency", "kernel": "k_recur_persist (one grid barrier pair per character)",
                         "achieved": U_bytes * n / (ms * 1e-3) / 1e9, "peak": None, "unit": "GB/s (shared-memory weight reads)",
                         "frac": None, "traffic": None,
                         "note": "serial GEMV chain: bounded by two grid-wide barriers per character, not by bandwidth; reported in us/char"},
            "sampled_text_head": bytes(out[:60]).decode("latin1")}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=list(WORKLOADS))
    ap.add_argument("--dtype", default="auto", choices=["auto", "bf16", "f32"])
    ap.add_argument("--cpu-seconds", type=float, default=None, help="CPU work per sample (default 12 s for cpu_baseline; --impl reference sizes it from --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-f32", action="store_true", help="skip the fp32-path measurement of the same workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = WORKLOADS[args.workload]
    N, B, S = cfg["N"], cfg["B"], cfg["S"]
    T = S - 1
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    fl = flops_per_charstep(N)
    config = {"workload": f"{args.workload}: {cfg['desc']}", "N": N, "M": M, "B_per_gpu": B, "T": T, "stride": T,
              "global_batch": B * world, "parallelism": f"dp{world}",
              "flops_per_charstep_dense": fl["dense"], "flops_per_charstep_alg": fl["alg"]}

    if args.workload == "cfg5" and args.impl != "reference":
        return bench_sampling(args, cfg) if rank == 0 else 0
    if args.impl == "reference":
        if rank != 0:
            return 0
        text = synthetic_text(1 << 20)
        per_step = args.cpu_seconds if args.cpu_seconds else max(2.0, min(20.0, 90.0 / (args.steps + 1)))
        cb, secs = cpu_arm(cfg, text.tobytes(), per_step, steps=args.steps, warmup=args.warmup)
        line = {"impl": "reference", "metric": "training chars/sec (fwd+BPTT+Adagrad)", "value": cb["value"], "unit": "chars/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "chars/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "R/lstm.cc and OV/lstm_eigen_BLAS need Eigen / OpenBLAS, which are not installed (no network): the CPU arm is "
                        "the oracle's port of the algorithm on the best BLAS present (torch CPU), all host threads set explicitly "
                        "(same count at every N); the reference sources compiled against oracle/eigen_shim pin the oracle's "
                        "semantics but are not a timing baseline"}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import torch.distributed as dist
    import eigen_lstm_b200 as el

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dtype = args.dtype
    if dtype == "auto":
        dtype = "bf16"
    try:
        net = el.LSTM(M, N, S, B, device=local_rank, dtype=el.BF16 if dtype == "bf16" else el.F32)
    except el.LstmError as ex:
        if args.dtype == "auto" and "not built" in str(ex):
            dtype = "f32"
            net = el.LSTM(M, N, S, B, device=local_rank, dtype=el.F32)
        else:
            raise
    if world > 1:
        from eigen_lstm_b200 import dp
        net.dp_init(rank, world, dp.broadcast_unique_id(dist, el.dp_unique_id, rank))
    net.init_params(seed=0, std=0.01, forget_bias=1.0)       # identical replicas on every rank
    net.reset_state(0, 0.0)
    total_steps = 2 * (args.steps + args.warmup) + 8
    chunk = T * total_steps + S + 16
    text = synthetic_text(chunk * B * world + 2 * S + 16, seed=0)
    net.load_text(text.tobytes())
    net.set_positions([S + (rank * B + b) * chunk for b in range(B)])
    stream = torch.cuda.ExternalStream(net.stream(), device=torch.device("cuda", local_rank))
    dpc = None
    if world > 1:
        dpc = dp_check(el, dist, rank, world, local_rank, net, el.BF16 if dtype == "bf16" else el.F32)
        net.init_params(seed=0, std=0.01, forget_bias=1.0)
        net.reset_state(0, 0.0)
        net.set_positions([S + (rank * B + b) * chunk for b in range(B)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (the iteration replays as one CUDA graph) ----
    net.train_text(args.warmup, stride=T, lr=LR, want_losses=False)
    net.sync()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = net.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    net.train_text(args.steps, stride=T, lr=LR, want_losses=False)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = net.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    # one more, profiled iteration (plain stream launches with CUDA events between the phases): the live
    # per-kernel durations behind `roofline` and `us_per_recurrent_timestep`
    net_variant = net.variant()
    net.set_profiling(True)
    losses = net.train_text(2, stride=T, lr=LR, want_losses=True)
    phases = net.phase_ms()
    net.set_profiling(False)
    if world > 1:
        tm = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    value = B * T * world * args.steps / (ms * 1e-3)

    # ---- end to end: host index windows in pinned memory, loss read back every step ----
    xw = torch.empty((S, B), dtype=torch.int32).pin_memory()
    tw = torch.empty((S, B), dtype=torch.int32).pin_memory()
    xn, tn = xw.numpy(), tw.numpy()
    pos = np.array([S + (rank * B + b) * chunk for b in range(B)], dtype=np.int64) + T * (args.steps + args.warmup + 2)
    tnp = text

    def host_window(p):
        # x_t = text[p - S + t], target_t = text[p - S + t + 1]   for every stream
        idx = (p[None, :] - S + np.arange(S)[:, None])
        xn[...] = tnp[idx]
        tn[...] = tnp[idx + 1]

    for _ in range(3):
        pos += T; host_window(pos); net.train_step(xn, tn, stride=T, lr=LR)
    barrier()
    e2e_losses = np.zeros(args.steps, dtype=np.float64)
    t0 = time.perf_counter()
    for i in range(args.steps):
        pos += T
        host_window(pos)                                          # the step's inputs are produced on the host ...
        net.train_step_async(xn, tn, e2e_losses, i, stride=T, lr=LR)   # ... copied in, and the step's loss is copied out, every step
    net.sync()                                                    # all K losses are on the host when the clock stops
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.all(np.isfinite(e2e_losses)) and np.all(e2e_losses > 0), e2e_losses
    if world > 1:
        tm = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = float(tm.item())
    e2e = {"value": B * T * world * args.steps / e2e_s, "unit": "chars/s", "h2d_bytes_per_step": int(2 * S * B * 4),
           "d2h_bytes_per_step": 8, "ms_per_step": e2e_s * 1e3 / args.steps,
           "api": "lstm_train_step (host int32 windows -> pinned staging -> device, loss -> host, every step; steps are "
                  "asynchronous, the clock stops after lstm_sync has delivered all losses)"}
    if net_variant.get("train_small"):
        e2e["api"] = ("lstm_train_step on the one-kernel path: the host window travels as a kernel argument (its bytes are the "
                      "h2d bytes), the kernel writes the step's loss into pinned host memory; one launch per step, asynchronous, "
                      "the clock stops after lstm_sync has delivered all losses")
    same = replicas_identical(dist, net, local_rank) if world > 1 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    # dominant kernel = the recurrent timestep kernel (forward + backward recurrences are 2/3 of the flops).
    # Durations are LIVE: CUDA events around each phase of the profiled iteration, divided by its launch count.
    fwd_us = phases["fwd_recurrence"] * 1e3 / T
    bwd_us = phases["bwd_recurrence"] * 1e3 / T
    fwd_flops = 2.0 * 4 * N * N * B                         # U*h(t-1) for B streams
    bwd_flops = 2.0 * N * (4 * N + (M if dtype == "bf16" else 0)) * B   # U^T*dg (+ Why^T*dy folded into K5's K range)
    dom = "fwd" if phases["fwd_recurrence"] >= phases["bwd_recurrence"] else "bwd"
    dom_us, dom_flops = (fwd_us, fwd_flops) if dom == "fwd" else (bwd_us, bwd_flops)
    var = net_variant if dtype == "bf16" else {}
    persistent = bool(var.get("fwd_persistent" if dom == "fwd" else "bwd_persistent"))
    if dtype != "bf16":
        dom_kernel = {"fwd": "k_step_fwd_f32", "bwd": "k_step_bwd_f32"}[dom]
    elif persistent:
        dom_kernel = {"fwd": "k_fwd_recur", "bwd": "k_bwd_recur"}[dom]
    else:
        dom_kernel = {"fwd": "k_fwd_step", "bwd": "k_bwd_step"}[dom]
    # one LAUNCH of the dominant kernel: the whole recurrence of a window (T timesteps) for the persistent kernels, one timestep
    # for the per-timestep ones; flops, duration and ncu's DRAM traffic are all per launch
    per_launch = T if persistent else 1
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    if dtype == "bf16" and args.workload == "cfg4" and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom_kernel, {}).get("dram_bytes_per_launch")
    achieved = dom_flops / (dom_us * 1e-6) / 1e12 if dom_us > 0 else 0.0
    roofline = {"bound": "tensor",
                "kernel": f"{dom_kernel} ({'all ' + str(T) + ' timesteps of a window in one persistent launch' if persistent else 'one recurrent timestep'}, "
                          f"{'tcgen05 bf16' if dtype == 'bf16' else 'SIMT fp32'})",
                "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                "traffic": traffic, "traffic_source": "ncu --set full dram__bytes_read+write per launch (profiles/r02f_kernels.md)" if traffic else None,
                "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside a long step)",
                "flops_per_launch": dom_flops * per_launch, "us_per_launch": dom_us * per_launch, "us_per_timestep": dom_us,
                "note": "the recurrence is a serial chain per timestep (contraction -> exchange -> gate math -> publish): the tensor "
                        "pipe idles during the hand-offs, see DESIGN.md section 5 for the measured anatomy"}
    if net_variant.get("train_small"):
        # the timed region ran the one-kernel path (train_small.cu): every iteration inside one persistent launch on one SM;
        # the phase times above come from the profiled iteration, which runs the launch-per-kernel path (same bits)
        it_us = ms * 1e3 / args.steps
        roofline = {"bound": "latency", "kernel": "k_train_small (whole training iterations in one persistent single-CTA launch, SIMT fp32)",
                    "achieved": fl["dense"] * B * T / (it_us * 1e-6) / 1e12, "peak": None, "unit": "TFLOP/s", "frac": None,
                    "traffic": None, "us_per_iteration": it_us,
                    "note": "~250 K multiply-adds per iteration: bounded by the dependent chains of one iteration and by 33 K exact "
                            "sqrt + divisions of the Adagrad update on ONE SM, not by any throughput roofline; "
                            "profiles/r02w_small_clocks.txt has the in-kernel anatomy"}
    BT = B * T
    wg_flops = 2.0 * 4 * N * (M + N + 1) * BT               # dW|dU|db as one GEMM (bf16 path)
    pre_flops = 2.0 * M * (N + 1) * BT                      # dWhy|dby
    lg_flops = 2.0 * M * N * BT
    P = 4 * N * M + 4 * N * N + 4 * N + M * N + M

    def rate(flops, ms_):
        return {"ms": ms_, "tflops": flops / (ms_ * 1e-3) / 1e12 if ms_ > 0 else None,
                "frac_of_peak": flops / (ms_ * 1e-3) / 1e12 / pk["tf_sustained"] if ms_ > 0 else None}
    kernels = {"fwd_recurrence(K2 x T)": rate(fwd_flops * T, phases["fwd_recurrence"]),
               "bwd_recurrence(K5 x T)": rate(bwd_flops * T, phases["bwd_recurrence"]),
               "weight_grads(K6a dW|dU|db)": rate(wg_flops, phases["weight_grads"]),
               "logits_softmax(K3)": rate(lg_flops, phases["logits_softmax"]),
               "adagrad(K7 + operand refresh)": {"ms": phases["adagrad"], "algorithmic_GBps": 20.0 * P / (phases["adagrad"] * 1e-3) / 1e9
                                                  if phases["adagrad"] > 0 else None, "hbm_peak_GBps": pk["hbm_gbs"]}}
    e2e_tf = fl["alg"] * value / 1e12
    line = {"metric": "training chars/sec (fwd+BPTT+Adagrad)", "value": value, "unit": "chars/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": dict(config, l2="per-step working set (activations) is >> the 126 MB L2; no flush needed"
                           if N * B * T * 24 > 4e8 else "small working set: L2-resident by design (latency-bound config)"),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "us_per_recurrent_timestep": {"forward": fwd_us, "backward": bwd_us}, "kernels": kernels,
            "phases_ms_last_step": phases,
            "end_to_end_tflops": {"alg": e2e_tf, "dense": fl["dense"] * value / 1e12,
                                  "frac_of_peak_alg": e2e_tf / (pk["tf_sustained"] * world)},
            "final_loss_bits_per_char": float(losses[-1] / T) if len(losses) else None, "learning_rate": LR,
            "kernel_variant": net_variant,
            "launch_mode": ("all timed iterations in ONE persistent kernel launch (train_small.cu); " if net_variant.get("train_small") else
                            "one CUDA graph per training iteration (timed region); ") + "plain stream launches for the profiled iteration"}
    if dpc is not None:
        dpc["replicas_identical_after_timed_region"] = same
        dpc["ok"] = bool(dpc["ok"] and same)
        line["dp_check"] = dpc
    if world == 1 and dtype == "bf16" and not args.no_f32:
        # the same configuration on the fp32 path (SIMT FFMA, the reference's arithmetic): same-precision throughput
        net.close()
        f = el.LSTM(M, N, S, B, device=local_rank, dtype=el.F32)
        f.init_params(seed=0, std=0.01, forget_bias=1.0); f.reset_state(0, 0.0)
        f.load_text(text.tobytes()); f.set_positions([S + b * chunk for b in range(B)])
        fs = torch.cuda.ExternalStream(f.stream(), device=torch.device("cuda", local_rank))
        f.train_text(2, stride=T, lr=LR, want_losses=False); f.sync()
        k = 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(fs); f.train_text(k, stride=T, lr=LR, want_losses=False); e1.record(fs)
        torch.cuda.synchronize()
        fms = e0.elapsed_time(e1) / k
        line["f32"] = {"value": B * T / (fms * 1e-3), "unit": "chars/s", "ms_per_step": fms, "steps": k,
                       "tflops_dense": fl["dense"] * B * T / (fms * 1e-3) / 1e12,
                       "note": "same workload on the LSTM_F32 path (every contraction fp32 SIMT FFMA, the reference's rounding): "
                               "the same-precision number; the headline `value` is the bf16-GEMM / fp32-accumulate path"}
        f.close()
    if not args.no_cpu_baseline and world == 1:
        cb, _ = cpu_arm(cfg, text.tobytes()[: 1 << 20], args.cpu_seconds or 12.0)
        line["cpu_baseline"] = cb
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
