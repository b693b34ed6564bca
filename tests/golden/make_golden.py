"""Generate the committed golden fixtures from the reference's own artefacts.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Nothing here is reference SOURCE: the inputs are the reference's trained-model text checkpoints,
its logged result, and slices of its public corpora.

Outputs (tests/golden/):
  enwik5_test.npz   W,U,b,Why,by parsed from OV/lstm_eigen_class_CUDA/models/enwik5_test_{W,U,Why,b,by}.txt
                    (N=32, M=256; Eigen `operator<<` text, 6 significant digits, io.h:16-32),
                    test_bytes = enwik5.txt[99000:100000] (the 1 % split of lstm.cc:73-86),
                    logged_test_bpc = column 4 of models/enwik5_test.txt:1 (3.24396)
  alice29_head.bin  first 8192 bytes of R/alice29.txt (config-1 loss-trace parity input)
  enwik6_head.bin   first 65536 bytes of R/enwik6.txt (batched-config parity input)
  enwik5_test_{W,U,Why,b,by}.txt   the checkpoint files themselves, byte for byte (590 KB): the product's own reader
                    (lstm_load_text_ckpt / `lstm --load`) must parse what the reference's save_to_disk wrote (io.h:16-32)
  enwik6.txt        R/enwik6.txt in full (10^6 bytes): the corpus of BASELINE config 2 and the stand-in for the missing
                    enwik7 of config 3 (bf16-vs-fp32 bits-per-char comparison, scripts/bpc_bf16_vs_f32.py)
  enwik6_hist.npy   byte histogram of R/enwik6.txt (int64[256]) for the synthetic bench corpus
"""
import os
import numpy as np

R = "/root/reference"
MD = f"{R}/optimized-obsfuscated_versions/lstm_eigen_class_CUDA/models"
OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    t = {}
    for name in ["W", "U", "Why", "b", "by"]:
        t[name] = np.loadtxt(f"{MD}/enwik5_test_{name}.txt", dtype=np.float64, ndmin=2).astype(np.float32)
    N, M = 32, 256
    assert t["W"].shape == (4 * N, M) and t["U"].shape == (4 * N, N) and t["Why"].shape == (M, N)
    t["b"] = t["b"].reshape(4 * N, 1)
    t["by"] = t["by"].reshape(M, 1)
    data = open(f"{R}/enwik5.txt", "rb").read()
    pct = len(data) // 100
    test_bytes = np.frombuffer(data[99 * pct:], dtype=np.uint8)
    logged = float(open(f"{MD}/enwik5_test.txt").read().split()[3])
    np.savez_compressed(f"{OUT}/enwik5_test.npz", test_bytes=test_bytes, logged_test_bpc=logged, **t)
    open(f"{OUT}/alice29_head.bin", "wb").write(open(f"{R}/alice29.txt", "rb").read()[:8192])
    e6 = open(f"{R}/enwik6.txt", "rb").read()
    open(f"{OUT}/enwik6_head.bin", "wb").write(e6[:65536])
    import shutil
    for name in ["W", "U", "Why", "b", "by"]:
        shutil.copyfile(f"{MD}/enwik5_test_{name}.txt", f"{OUT}/enwik5_test_{name}.txt")
        os.chmod(f"{OUT}/enwik5_test_{name}.txt", 0o644)
    open(f"{OUT}/enwik6.txt", "wb").write(e6)
    np.save(f"{OUT}/enwik6_hist.npy", np.bincount(np.frombuffer(e6, dtype=np.uint8), minlength=256).astype(np.int64))
    print("logged", logged, "test bytes", test_bytes.size)


if __name__ == "__main__":
    main()
