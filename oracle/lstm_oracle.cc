// oracle/lstm_oracle.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Plain-C++ CPU restatement of the krocki/Eigen-LSTM character-LSTM hot path.  It exists only
// as the checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
// --impl reference leg).  Nothing under eigen_lstm_b200/ may include, link or call it.
//
// Parity status: PINNED.
//   * Against the reference's own SOURCE: all six UNMODIFIED CPU programs of the reference — R/lstm.cc, OV/lstm_eigen_opt,
//     OV/lstm_eigen_class, OV/lstm_eigen_BLAS/lstm.cc (its
//     pure-Eigen branch, B = 4), OV/lstm_eigen_class_batch/lstm.cc + lstm.h (double precision, softmax shift, gradient
//     check) and OV/lstm_eigen_class_batch/lstm_segment.cc (window stride > 1) — are compiled by `make -C oracle ref` into
//     oracle/_ref/.  Their one external dependency, Eigen, is not installed in this image (no network), so they are built
//     against oracle/eigen_shim/ — a from-scratch stand-in for the Eigen members they use — with std::random_device
//     replaced by a seed counter.  Driven with the same seeds, this oracle reproduces those programs' output EXACTLY: every
//     printed loss and every sampled character after thousands of Adagrad iterations
//     (tests/test_oracle_vs_reference_source.py; committed outputs: tests/golden/ref_*_run.json).  That pins window
//     construction, state carry, loss normalisation, batch summation, Adagrad order, seeding order and sampling in float
//     and double.  What a shim cannot pin is Eigen's internal float summation order (the shim, like this file, sums
//     sequentially).
//   * forward semantics (gate order [i,o,f,u], tanh'd carried cell, softmax, log2 loss) by the reference's
//     known-answer fixture models/enwik5_test_{W,U,Why,b,by}.txt -> 3.24396 bits/char (tests/test_oracle_golden.py),
//   * backward semantics by the reference's own acceptance rule for its numerical gradient check (max rel.err
//     < 1e-1, mean < 1e-3; we hold 1e-6) in double precision (tests/test_oracle_gradcheck.py).
//
// Reference shorthand: R/ = /root/reference/, OV/ = R/optimized-obsfuscated_versions/.
// All matrices are column-major like Eigen's default (R/lstm.cc:65-84): element (r,c) of an
// (rows x cols) matrix lives at [r + rows*c].
//
// Build: see oracle/Makefile (g++ -O3 -std=c++11, the flags of R/Makefile:8-13, plus -fPIC
// -shared; -mavx2 -mfma -pthread for the multi-threaded "all host cores" baseline build).

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
#include <chrono>
#include <thread>
#include <functional>

#define ORACLE_EPS 1e-10  // R/lstm.cc:25  (double literal: the add happens in double)

namespace {

// Threading for the "all host cores" CPU baseline (the reference itself is single-threaded;
// OV/lstm_eigen_BLAS gets threads only from OpenBLAS).  libgomp is not in this image, so a plain
// std::thread fork/join.  g_threads == 1 (default) runs inline, in index order.
int g_threads = 1;
template <typename F>
void parallel_for(int n, F fn) {
  const int nt = g_threads < n ? g_threads : n;
  if (nt <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
  std::vector<std::thread> th;
  for (int w = 0; w < nt; w++)
    th.emplace_back([=]() { for (int i = (int)((long)n * w / nt); i < (int)((long)n * (w + 1) / nt); i++) fn(i); });
  for (auto& t : th) t.join();
}

// R/lstm.cc:31-33
template <typename R> inline R logistic(R x);
template <> inline float logistic<float>(float x) { return 1.0f / (1.0f + ::expf(-x)); }
template <> inline double logistic<double>(double x) { return 1.0 / (1.0 + ::exp(-x)); }
template <typename R> inline R tanh_r(R x);
template <> inline float tanh_r<float>(float x) { return ::tanhf(x); }
template <> inline double tanh_r<double>(double x) { return ::tanh(x); }
template <typename R> inline R exp_r(R x);
template <> inline float exp_r<float>(float x) { return ::expf(x); }
template <> inline double exp_r<double>(double x) { return ::exp(x); }
template <typename R> inline R log2_r(R x);
template <> inline float log2_r<float>(float x) { return ::log2f(x); }
template <> inline double log2_r<double>(double x) { return ::log2(x); }
// R/lstm.cc:36-38, 41-43
template <typename R> inline R tanh_prime(R x) { return R(1) - x * x; }
template <typename R> inline R logistic_prime(R x) { return x * (R(1) - x); }
// R/lstm.cc:46-48: sqrtf(x + eps) with eps a double literal
inline float sqrt_eps(float x) { return sqrtf((float)((double)x + ORACLE_EPS)); }
inline double sqrt_eps(double x) { return sqrt(x + ORACLE_EPS); }

// R/lstm.cc:364-380: a fresh mt19937 per call, normal_distribution<double>, filled in
// (row i, col j) iteration order into a column-major matrix.  The reference seeds from
// std::random_device; we take the seed as an argument so runs are reproducible.
template <typename R>
void randn_fill(R* m, int rows, int cols, float mean, float stddev, uint64_t seed) {  // float args like :364
  std::mt19937 mt((uint32_t)seed);
  std::normal_distribution<> dist(mean, stddev);
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) m[i + (size_t)rows * j] = (R)dist(mt);
}

template <typename R>
struct Oracle {
  int M, N, S, B;
  int softmax_shift = 0;  // 1: subtract the global max of the M x B logits (OV/lstm_eigen_class_batch/lstm.h:175)
  int dense_onehot = 0;   // 1: do W*x and dW += dg*x^T as dense products like R/lstm.cc:176,251 (timing honesty)
  std::vector<R> W, U, b, Why, by;
  std::vector<R> dW, dU, db, dWhy, dby;
  std::vector<R> mW, mU, mb, mWhy, mby;
  // per-timestep state, t = 0..S-1 (R/lstm.cc:75-84)
  std::vector<R> h, c;      // [S][N*B]
  std::vector<R> g;         // [S][4N*B]   activated gates [i o f u]
  std::vector<R> probs;     // [S][M*B]
  std::vector<R> dgs;       // [S][4N*B]   (kept for differential tests; the reference keeps only one)
  std::vector<int> xi, ti;  // [S][B] input / target byte per column, -1 = all-zero column
  std::vector<size_t> pos;  // stream positions (OV/lstm_eigen_opt/lstm.cc:140-144)

  Oracle(int M_, int N_, int S_, int B_) : M(M_), N(N_), S(S_), B(B_) {
    size_t n4 = 4 * (size_t)N;
    W.assign(n4 * M, 0); U.assign(n4 * N, 0); b.assign(n4, 0);
    Why.assign((size_t)M * N, 0); by.assign(M, 0);
    dW = W; dU = U; db = b; dWhy = Why; dby = by;
    mW = W; mU = U; mb = b; mWhy = Why; mby = by;
    h.assign((size_t)S * N * B, 0); c = h;
    g.assign((size_t)S * n4 * B, 0); dgs = g;
    probs.assign((size_t)S * M * B, 0);
    xi.assign((size_t)S * B, -1); ti.assign((size_t)S * B, -1);
    pos.assign(B, (size_t)S);
  }
  R* ht(int t) { return &h[(size_t)t * N * B]; }
  R* ct(int t) { return &c[(size_t)t * N * B]; }
  R* gt(int t) { return &g[(size_t)t * 4 * N * B]; }
  R* pt(int t) { return &probs[(size_t)t * M * B]; }

  // One LSTM cell step on column vectors (R/lstm.cc:176-192).  x < 0 means x is the zero vector.
  void cell(int x, const R* hprev, const R* cprev, R* gcol, R* cnew, R* hnew) const {
    const int n4 = 4 * N;
    for (int r = 0; r < n4; r++) gcol[r] = 0;
    // U * h(t-1)
    for (int k = 0; k < N; k++) {
      const R hk = hprev[k];
      const R* Uk = &U[(size_t)n4 * k];
      for (int r = 0; r < n4; r++) gcol[r] += Uk[r] * hk;
    }
    // W * x(t): x is one-hot, so the product is column x of W (exactly; adding zeros is exact)
    if (dense_onehot) {
      std::vector<R> acc(n4, 0);
      for (int m = 0; m < M; m++) {
        const R xm = (m == x) ? R(1) : R(0);
        const R* Wm = &W[(size_t)n4 * m];
        for (int r = 0; r < n4; r++) acc[r] += Wm[r] * xm;
      }
      for (int r = 0; r < n4; r++) gcol[r] = acc[r] + gcol[r];
    } else if (x >= 0) {
      const R* Wx = &W[(size_t)n4 * x];
      for (int r = 0; r < n4; r++) gcol[r] = Wx[r] + gcol[r];
    }
    for (int r = 0; r < n4; r++) gcol[r] += b[r];
    for (int r = 0; r < 3 * N; r++) gcol[r] = logistic<R>(gcol[r]);          // :179
    for (int r = 3 * N; r < n4; r++) gcol[r] = tanh_r<R>(gcol[r]);           // :182
    for (int j = 0; j < N; j++) {
      R cc = gcol[j] * gcol[3 * N + j] + gcol[2 * N + j] * cprev[j];         // :185-186
      cc = tanh_r<R>(cc);                                                    // :189  carried value is tanh'd
      cnew[j] = cc;
      hnew[j] = gcol[N + j] * cc;                                            // :192
    }
  }

  // logits + softmax for one column (R/lstm.cc:195-201); shift = value subtracted before exp
  void softmax_col(const R* hcol, R* p, R shift) const {
    for (int m = 0; m < M; m++) p[m] = 0;
    for (int n = 0; n < N; n++) {
      const R hn = hcol[n];
      const R* Wn = &Why[(size_t)M * n];
      for (int m = 0; m < M; m++) p[m] += Wn[m] * hn;
    }
    R sum = 0;
    for (int m = 0; m < M; m++) { p[m] = exp_r<R>(p[m] + by[m] - shift); sum += p[m]; }
    for (int m = 0; m < M; m++) p[m] = p[m] / sum;
  }

  // Forward over t = 1..S-1 (R/lstm.cc:173-209; batched OV/lstm_eigen_opt/lstm.cc:216-251).
  // Returns the iteration loss the reference accumulates: sum_t (sum_b -log2 p[target]) / B.
  double forward() {
    double loss = 0;
    for (int t = 1; t < S; t++) {
      parallel_for(B, [&](int bb) {
        cell(xi[(size_t)t * B + bb], ht(t - 1) + (size_t)N * bb, ct(t - 1) + (size_t)N * bb,
             gt(t) + (size_t)4 * N * bb, ct(t) + (size_t)N * bb, ht(t) + (size_t)N * bb); });
      R shift = 0;
      if (softmax_shift) {
        // OV/lstm_eigen_class_batch/lstm.h:175 subtracts the max over the whole M x B logit matrix
        R mx = -INFINITY;
        std::vector<R> y(M);
        for (int bb = 0; bb < B; bb++) {   // y = Why * h + by: the product first, then the bias (same association as softmax_col)
          for (int m = 0; m < M; m++) y[m] = 0;
          for (int n = 0; n < N; n++)
            for (int m = 0; m < M; m++) y[m] += Why[(size_t)M * n + m] * ht(t)[(size_t)N * bb + n];
          for (int m = 0; m < M; m++) { const R v = y[m] + by[m]; mx = v > mx ? v : mx; }
        }
        shift = mx;
      }
      parallel_for(B, [&](int bb) { softmax_col(ht(t) + (size_t)N * bb, pt(t) + (size_t)M * bb, shift); });
      R s = 0;  // surprisals.sum() in the working precision (:204-207)
      for (int bb = 0; bb < B; bb++) {
        const int k = ti[(size_t)t * B + bb];
        if (k >= 0) s += -log2_r<R>(pt(t)[(size_t)M * bb + k]);
      }
      loss += s / (R)B;  // OV/lstm_eigen_opt/lstm.cc:249 (B = 1 in R/lstm.cc)
    }
    return loss;
  }

  // Backward t = S-1..1 (R/lstm.cc:213-257; batched OV/lstm_eigen_opt/lstm.cc:255-303)
  void backward() {
    const int n4 = 4 * N;
    std::fill(dWhy.begin(), dWhy.end(), R(0)); std::fill(dby.begin(), dby.end(), R(0));
    std::fill(dU.begin(), dU.end(), R(0)); std::fill(dW.begin(), dW.end(), R(0));
    std::fill(db.begin(), db.end(), R(0));
    std::vector<R> dhnext((size_t)N * B, 0), dcnext((size_t)N * B, 0);
    std::vector<R> dy((size_t)M * B), dh((size_t)N * B);
    for (int t = S - 1; t > 0; t--) {
      R* dg = &dgs[(size_t)t * n4 * B];
      // dy = probs - target (:225)
      for (int bb = 0; bb < B; bb++) {
        const int k = ti[(size_t)t * B + bb];
        for (int m = 0; m < M; m++) dy[(size_t)M * bb + m] = pt(t)[(size_t)M * bb + m] - (m == k ? R(1) : R(0));
      }
      // dWhy += dy * h(t)^T (:226).  Batched (OV/lstm_eigen_BLAS/lstm.cc:295): the product over the B streams is
      // evaluated first (summed over b from zero) and THEN added — the association `dWhy += (dy * h^T)` has.
      parallel_for(N, [&](int n) {
        std::vector<R> tmp(M, R(0));
        for (int bb = 0; bb < B; bb++) {
          const R hn = ht(t)[(size_t)N * bb + n];
          for (int m = 0; m < M; m++) tmp[m] += dy[(size_t)M * bb + m] * hn;
        }
        for (int m = 0; m < M; m++) dWhy[(size_t)M * n + m] += tmp[m]; });
      // dby += dy.rowwise().sum() (:227; batched OV/lstm_eigen_BLAS/lstm.cc:297)
      {
        std::vector<R> tmp(M, R(0));
        for (int bb = 0; bb < B; bb++)
          for (int m = 0; m < M; m++) tmp[m] += dy[(size_t)M * bb + m];
        for (int m = 0; m < M; m++) dby[m] += tmp[m];
      }
      parallel_for(B, [&](int bb) {
        const R* gcol = gt(t) + (size_t)n4 * bb;
        const R* ccol = ct(t) + (size_t)N * bb;
        const R* cprev = ct(t - 1) + (size_t)N * bb;
        R* dgc = dg + (size_t)n4 * bb;
        for (int n = 0; n < N; n++) {
          // dh = Why^T * dy + dhnext (:228)
          R acc = 0;
          const R* Wn = &Why[(size_t)M * n];
          for (int m = 0; m < M; m++) acc += Wn[m] * dy[(size_t)M * bb + m];
          const R dhv = acc + dhnext[(size_t)N * bb + n];
          dh[(size_t)N * bb + n] = dhv;
          // dc = (dh .* o + dcnext) .* tanh'(c(t)) (:233-235)
          R dcv = dhv * gcol[N + n] + dcnext[(size_t)N * bb + n];
          dcv = dcv * tanh_prime<R>(ccol[n]);
          // gates (:238-241), through the nonlinearities (:244-247)
          dgc[N + n] = (dhv * ccol[n]) * logistic_prime<R>(gcol[N + n]);                  // do
          dgc[n] = (dcv * gcol[3 * N + n]) * logistic_prime<R>(gcol[n]);                  // di
          dgc[2 * N + n] = (dcv * cprev[n]) * logistic_prime<R>(gcol[2 * N + n]);         // df
          dgc[3 * N + n] = (dcv * gcol[n]) * tanh_prime<R>(gcol[3 * N + n]);              // du
          dcnext[(size_t)N * bb + n] = dcv * gcol[2 * N + n];                             // :256
        }
      });
      // dU += dg * h(t-1)^T (:250): product over the streams first, then the add (see dWhy above)
      parallel_for(N, [&](int k) {
        std::vector<R> tmp(n4, R(0));
        for (int bb = 0; bb < B; bb++) {
          const R hk = ht(t - 1)[(size_t)N * bb + k];
          const R* dgc = dg + (size_t)n4 * bb;
          for (int r = 0; r < n4; r++) tmp[r] += dgc[r] * hk;
        }
        R* dUk = &dU[(size_t)n4 * k];
        for (int r = 0; r < n4; r++) dUk[r] += tmp[r]; });
      // dW += dg * x(t)^T (:251) — x one-hot: column m receives the sum of dg over the streams whose input is m
      if (dense_onehot) {
        parallel_for(M, [&](int m) {
          std::vector<R> tmp(n4, R(0));
          for (int bb = 0; bb < B; bb++) {
            const R xm = (xi[(size_t)t * B + bb] == m) ? R(1) : R(0);
            const R* dgc = dg + (size_t)n4 * bb;
            for (int r = 0; r < n4; r++) tmp[r] += dgc[r] * xm;
          }
          R* dWm = &dW[(size_t)n4 * m];
          for (int r = 0; r < n4; r++) dWm[r] += tmp[r]; });
      } else {
        std::vector<R> tmp(n4);
        for (int bb = 0; bb < B; bb++) {
          const int x = xi[(size_t)t * B + bb];
          if (x < 0) continue;
          bool seen = false;   // column x was already handled by an earlier stream with the same input byte
          for (int b2 = 0; b2 < bb && !seen; b2++) seen = (xi[(size_t)t * B + b2] == x);
          if (seen) continue;
          std::fill(tmp.begin(), tmp.end(), R(0));
          for (int b2 = bb; b2 < B; b2++) {
            if (xi[(size_t)t * B + b2] != x) continue;
            const R* dgc = dg + (size_t)n4 * b2;
            for (int r = 0; r < n4; r++) tmp[r] += dgc[r];
          }
          R* dWm = &dW[(size_t)n4 * x];
          for (int r = 0; r < n4; r++) dWm[r] += tmp[r];
        }
      }
      // db += dg.rowwise().sum() (:252; batched OV/lstm_eigen_BLAS/lstm.cc:335)
      {
        std::vector<R> tmp(n4, R(0));
        for (int bb = 0; bb < B; bb++)
          for (int r = 0; r < n4; r++) tmp[r] += dg[(size_t)n4 * bb + r];
        for (int r = 0; r < n4; r++) db[r] += tmp[r];
      }
      // dhnext = U^T * dg (:255)
      parallel_for(B, [&](int bb) {
        for (int k = 0; k < N; k++) {
          const R* Uk = &U[(size_t)n4 * k];
          const R* dgc = dg + (size_t)n4 * bb;
          R acc = 0;
          for (int r = 0; r < n4; r++) acc += Uk[r] * dgc[r];
          dhnext[(size_t)N * bb + k] = acc;
        } });
    }
  }

  // R/lstm.cc:259-272
  static void adagrad_one(std::vector<R>& p, const std::vector<R>& d, std::vector<R>& m, R lr) {
    const size_t n = p.size();
    const int chunks = 64;
    parallel_for(chunks, [&](int ch) {
      for (size_t i = n * ch / chunks; i < n * (ch + 1) / chunks; i++) {
        m[i] += d[i] * d[i];
        p[i] -= lr * (d[i] / sqrt_eps(m[i]));
      } });
  }
  void adagrad(R lr) {
    adagrad_one(Why, dWhy, mWhy, lr); adagrad_one(by, dby, mby, lr);
    adagrad_one(U, dU, mU, lr); adagrad_one(W, dW, mW, lr); adagrad_one(b, db, mb, lr);
  }

  // Window advance by one event per stream — the literal restatement of
  // OV/lstm_eigen_opt/lstm.cc:190-213 (for B = 1 and pos = i: R/lstm.cc:155-170).
  void advance_one(const uint8_t* data, size_t length) {
    for (int bb = 0; bb < B; bb++) {
      const int event = data[pos[bb]];
      pos[bb]++;
      if (pos[bb] >= length) pos[bb] = S;
      for (int s = 1; s < S; s++) {
        xi[(size_t)(s - 1) * B + bb] = xi[(size_t)s * B + bb];
        ti[(size_t)(s - 1) * B + bb] = ti[(size_t)s * B + bb];
      }
      ti[(size_t)(S - 1) * B + bb] = event;
      xi[(size_t)(S - 1) * B + bb] = ti[(size_t)(S - 2) * B + bb];
    }
  }
  // State carry for a window shift of `stride` timesteps: h(0) <- h(stride) (R/lstm.cc:163-164
  // for stride 1; OV/lstm_eigen_class_batch/lstm_segment.cc:183-184 for stride > 1).
  void carry(int stride) {
    if (stride <= 0) return;
    if (stride > S - 1) stride = S - 1;
    memcpy(ht(0), ht(stride), sizeof(R) * (size_t)N * B);
    memcpy(ct(0), ct(stride), sizeof(R) * (size_t)N * B);
  }

  // One full training iteration on device-less text: advance window, fwd, bwd, Adagrad.
  double train_iter(const uint8_t* data, size_t length, int stride, R lr) {
    carry(stride);
    for (int s = 0; s < stride; s++) advance_one(data, length);
    const double loss = forward();
    backward();
    adagrad(lr);
    return loss;
  }

  // Held-out evaluation, OV/lstm_eigen_class_CUDA/lstm.cc:661-720 (h = c = 0 since reset_std = 0, :45)
  double eval_bpc(const uint8_t* data, size_t n) const {
    std::vector<R> hh(N, 0), cc(N, 0), gg(4 * N), p(M), hn(N), cn(N);
    double err = 0;
    for (size_t ii = 0; ii + 1 < n; ii++) {
      cell(data[ii], hh.data(), cc.data(), gg.data(), cn.data(), hn.data());
      hh = hn; cc = cn;
      softmax_col(hh.data(), p.data(), 0);
      err += -log2((double)p[data[ii + 1]]);
    }
    return err / (double)(n - 1);
  }

  // Sampling, R/lstm.cc:293-356: sample from the current h first, then advance.
  // h0/c0 are supplied by the caller (the reference draws them N(0, 0.1), :306-307).
  void sample(const R* h0, const R* c0, uint64_t seed, uint8_t* out, size_t n, int greedy) const {
    std::vector<R> hh(h0, h0 + N), cc(c0, c0 + N), gg(4 * N), p(M), cdf(M), hn(N), cn(N);
    std::mt19937 gen((uint32_t)seed);
    std::uniform_real_distribution<> dis(0, 1);
    for (size_t i = 0; i < n; i++) {
      softmax_col(hh.data(), p.data(), 0);
      cdf[0] = p[0];
      for (int ii = 1; ii < M; ii++) cdf[ii] = cdf[ii - 1] + p[ii];
      int index = 0;
      if (greedy) {
        for (int ii = 1; ii < M; ii++) if (p[ii] > p[index]) index = ii;
      } else {
        const R r = (R)dis(gen);  // `float r = dis(gen)` R/lstm.cc:326; `double r` in the double snapshots (OV/lstm_eigen_class_batch/lstm.cc:378)
        for (int ii = 0; ii < M; ii++) if (r < cdf[ii]) { index = ii; break; }
      }
      out[i] = (uint8_t)index;
      cell(index, hh.data(), cc.data(), gg.data(), cn.data(), hn.data());
      hh = hn; cc = cn;
    }
  }
};

template <typename R>
std::vector<R>* pick(Oracle<R>* o, int kind, int which) {
  // kind 0 = param, 1 = grad, 2 = adagrad memory; which: 0 W, 1 U, 2 b, 3 Why, 4 by
  std::vector<R>* tab[3][5] = {{&o->W, &o->U, &o->b, &o->Why, &o->by},
                               {&o->dW, &o->dU, &o->db, &o->dWhy, &o->dby},
                               {&o->mW, &o->mU, &o->mb, &o->mWhy, &o->mby}};
  if (kind < 0 || kind > 2 || which < 0 || which > 4) return nullptr;
  return tab[kind][which];
}

}  // namespace

#define ORACLE_API(PFX, R)                                                                            \
  extern "C" void* PFX##_create(int M, int N, int S, int B) { return new Oracle<R>(M, N, S, B); }     \
  extern "C" void PFX##_destroy(void* o) { delete (Oracle<R>*)o; }                                    \
  extern "C" void PFX##_set_options(void* o, int softmax_shift, int dense_onehot) {                   \
    ((Oracle<R>*)o)->softmax_shift = softmax_shift; ((Oracle<R>*)o)->dense_onehot = dense_onehot; }   \
  extern "C" long PFX##_tensor_size(void* o, int kind, int which) {                                   \
    auto* v = pick<R>((Oracle<R>*)o, kind, which); return v ? (long)v->size() : -1; }                 \
  extern "C" int PFX##_set_tensor(void* o, int kind, int which, const R* src) {                       \
    auto* v = pick<R>((Oracle<R>*)o, kind, which); if (!v) return -1;                                 \
    memcpy(v->data(), src, sizeof(R) * v->size()); return 0; }                                        \
  extern "C" int PFX##_get_tensor(void* o, int kind, int which, R* dst) {                             \
    auto* v = pick<R>((Oracle<R>*)o, kind, which); if (!v) return -1;                                 \
    memcpy(dst, v->data(), sizeof(R) * v->size()); return 0; }                                        \
  extern "C" void PFX##_randn(R* m, int rows, int cols, float mean, float sd, uint64_t seed) {        \
    randn_fill<R>(m, rows, cols, mean, sd, seed); }                                                   \
  /* state access: what 0 h, 1 c (N*B), 2 g, 5 dg (4N*B), 3 probs (M*B); column-major per t */        \
  extern "C" int PFX##_get_state(void* o_, int what, int t, R* dst) {                                 \
    auto* o = (Oracle<R>*)o_; if (t < 0 || t >= o->S) return -1;                                      \
    size_t n; const R* src;                                                                           \
    switch (what) {                                                                                   \
      case 0: n = (size_t)o->N * o->B; src = o->ht(t); break;                                         \
      case 1: n = (size_t)o->N * o->B; src = o->ct(t); break;                                         \
      case 2: n = (size_t)4 * o->N * o->B; src = o->gt(t); break;                                     \
      case 3: n = (size_t)o->M * o->B; src = o->pt(t); break;                                         \
      case 5: n = (size_t)4 * o->N * o->B; src = &o->dgs[(size_t)t * 4 * o->N * o->B]; break;         \
      default: return -1; }                                                                           \
    memcpy(dst, src, sizeof(R) * n); return 0; }                                                      \
  extern "C" int PFX##_set_state(void* o_, int what, int t, const R* src) {                           \
    auto* o = (Oracle<R>*)o_; if (t < 0 || t >= o->S || what < 0 || what > 1) return -1;              \
    memcpy(what == 0 ? o->ht(t) : o->ct(t), src, sizeof(R) * (size_t)o->N * o->B); return 0; }        \
  extern "C" void PFX##_set_window(void* o_, const int* x, const int* tg) {                           \
    auto* o = (Oracle<R>*)o_; memcpy(o->xi.data(), x, sizeof(int) * o->xi.size());                    \
    memcpy(o->ti.data(), tg, sizeof(int) * o->ti.size()); }                                           \
  extern "C" void PFX##_get_window(void* o_, int* x, int* tg) {                                       \
    auto* o = (Oracle<R>*)o_; memcpy(x, o->xi.data(), sizeof(int) * o->xi.size());                    \
    memcpy(tg, o->ti.data(), sizeof(int) * o->ti.size()); }                                           \
  extern "C" void PFX##_set_positions(void* o_, const uint64_t* p) {                                  \
    auto* o = (Oracle<R>*)o_; for (int i = 0; i < o->B; i++) o->pos[i] = (size_t)p[i]; }              \
  extern "C" void PFX##_get_positions(void* o_, uint64_t* p) {                                        \
    auto* o = (Oracle<R>*)o_; for (int i = 0; i < o->B; i++) p[i] = (uint64_t)o->pos[i]; }            \
  extern "C" double PFX##_forward(void* o) { return ((Oracle<R>*)o)->forward(); }                     \
  extern "C" void PFX##_backward(void* o) { ((Oracle<R>*)o)->backward(); }                            \
  extern "C" void PFX##_adagrad(void* o, double lr) { ((Oracle<R>*)o)->adagrad((R)lr); }              \
  extern "C" void PFX##_carry(void* o, int stride) { ((Oracle<R>*)o)->carry(stride); }                \
  extern "C" void PFX##_advance(void* o, const uint8_t* d, size_t len, int stride) {                  \
    for (int s = 0; s < stride; s++) ((Oracle<R>*)o)->advance_one(d, len); }                          \
  /* run `iters` training iterations; losses[i] = the reference's per-iteration `loss`.  */           \
  /* returns wall seconds. */                                                                         \
  extern "C" double PFX##_train(void* o_, const uint8_t* d, size_t len, int iters, int stride,        \
                                double lr, double* losses) {                                          \
    auto* o = (Oracle<R>*)o_;                                                                         \
    auto t0 = std::chrono::steady_clock::now();                                                       \
    for (int i = 0; i < iters; i++) {                                                                 \
      const double l = o->train_iter(d, len, stride, (R)lr);                                          \
      if (losses) losses[i] = l; }                                                                    \
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }            \
  extern "C" double PFX##_eval_bpc(void* o, const uint8_t* d, size_t n) {                             \
    return ((Oracle<R>*)o)->eval_bpc(d, n); }                                                         \
  extern "C" void PFX##_sample(void* o, const R* h0, const R* c0, uint64_t seed, uint8_t* out,        \
                               size_t n, int greedy) {                                                \
    ((Oracle<R>*)o)->sample(h0, c0, seed, out, n, greedy); }

ORACLE_API(oracle32, float)
ORACLE_API(oracle64, double)

// the double-precision snapshots call randn(MatrixXd&, double mean, double stddev) (OV/lstm_eigen_class_batch/lstm.cc:513)
extern "C" void oracle64_randn_d(double* m, int rows, int cols, double mean, double sd, uint64_t seed) {
  std::mt19937 mt((uint32_t)seed);
  std::normal_distribution<> dist(mean, sd);
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) m[i + (size_t)rows * j] = dist(mt);
}
// n draws of mt19937(seed) + uniform_real_distribution<double>(0,1): the generator of the reference's sampled gradient
// check (OV/lstm_eigen_class_batch/lstm.h:211-219) and of its sampling loop
extern "C" void oracle_uniform01(uint64_t seed, size_t n, double* out) {
  std::mt19937 gen((uint32_t)seed);
  std::uniform_real_distribution<> dis(0, 1);
  for (size_t i = 0; i < n; i++) out[i] = dis(gen);
}

extern "C" int oracle_set_threads(int n) {
  if (n <= 0) n = (int)std::thread::hardware_concurrency();
  if (n <= 0) n = 1;
  g_threads = n;
  return g_threads;
}
extern "C" int oracle_num_threads() { return g_threads; }
