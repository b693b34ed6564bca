"""Turn gpurun_out/launches_<tag>.csv and gpurun_out/prof_<kernel>_<tag>.ncu-rep into committed summaries:
profiles/<tag>_launches.md (per-kernel totals and shares of one training iteration) and
profiles/<tag>_kernels.md (the ncu --set full headline metrics per hot kernel).  Run here (no GPU needed)."""
import collections
import csv
import glob
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)


def launches():
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[ui], v)
        name = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of the bench command (config 4: N=2048, B=256, T=256, bf16; every training iteration it runs)\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --workload cfg4 --steps 1 --warmup 3` "
                "(LSTM_NO_GRAPH=1: same kernels as plain launches; the first lines are the one-time parameter set-up).  "
                "Times are cold-cache and serialised — compare SHARES.\n\n"
                "| kernel | launches | total ms | avg µs | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.3f} | {v[1] / v[0]:.2f} | {100 * v[1] / tot:.1f} % |\n")
        f.write(f"| **total** | {sum(v[0] for v in agg.values())} | {tot / 1e3:.3f} | | |\n")


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "sm__cycles_elapsed.max", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic",
]


def kernels():
    reps = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_*_{tag}.ncu-rep")))
    if not reps:
        return
    with open(os.path.join(OUT, f"{tag}_kernels.md"), "w") as f:
        f.write(f"# {tag}: ncu --set full, one launch per hot kernel (config 4)\n\n"
                "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1` over the bench command.\n")
        for rep in reps:
            out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(out.splitlines()))
            if len(rows) < 3:
                continue
            hdr, units, vals = rows[0], rows[1], rows[2]
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else os.path.basename(rep)
            f.write(f"\n## `{name.split('(')[0]}`  ({os.path.basename(rep)})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"| {w} | {vals[i]} | {units[i]} |\n")


launches()
kernels()
print(open(os.path.join(OUT, f"{tag}_launches.md")).read() if os.path.exists(os.path.join(OUT, f"{tag}_launches.md")) else "no launch list")
