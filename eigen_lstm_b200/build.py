"""Build the CUDA shared library (and the drop-in `lstm` host binary) in-tree for sm_100a.

    python -m eigen_lstm_b200.build            # incremental
    python -m eigen_lstm_b200.build --force

Outputs: eigen_lstm_b200/liblstm_b200.so, eigen_lstm_b200/lstm  (git-ignored, shipped by gpurun).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "liblstm_b200.so")
BIN = os.path.join(HERE, "lstm")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
          "--expt-extended-lambda", "-Xptxas", "-v"]
CU = ["capi.cu", "kernels_f32.cu", "train_small.cu", "gradcheck.cu", "tc_path.cu", "tc_kernels.cu", "tc_steps.cu", "tc_recur.cu"]


def _newer(src_list, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "lstm_b200.h"))
    objs = []
    for f in CU:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f + ".o")
        objs.append(obj)
        if force or _newer([src] + headers, obj):
            cmd = [NVCC] + ARCH + CFLAGS + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(obj + ".log", "w") as lf:
                lf.write(r.stdout + r.stderr)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {f}")
    if force or _newer(objs, LIB):
        subprocess.check_call([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"])
    main_src = os.path.join(CSRC, "lstm_main.cc")
    if os.path.exists(main_src) and (force or _newer([main_src, LIB] + headers, BIN)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", main_src, "-o", BIN, "-L" + HERE, "-llstm_b200",
                               "-Wl,-rpath,$ORIGIN", "-Wl,-rpath-link," + HERE])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
