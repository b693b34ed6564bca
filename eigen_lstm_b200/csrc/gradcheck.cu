// gradcheck.cu — the reference's numerical gradient check, on the device, in double precision.
//
// OV/lstm_eigen_class_batch/lstm.h:203-261 (numerical_grads / compute_all_numerical_grads) perturbs ~100 randomly
// chosen entries of each parameter tensor by +/- delta = 1e-5, re-runs forward_loss() for each, and compares the central
// difference with the analytic gradient (check_gradient_error, lstm.cc:440-510).  The reference does this in double
// (OV/lstm_eigen_class*), because a float forward cannot resolve a 1e-5 perturbation.  Here the analytic gradients are
// whatever the context's training path produces (fp32 SIMT or bf16 tensor core); the perturbed forward passes run in a
// dedicated fp64 kernel: ONE CTA per (probe, sign) evaluates the window's loss with a single parameter entry shifted —
// the parameters are read as double(fp32 master) and the shift is applied on the fly, so nothing is copied or restored.
//
// The loss differentiated is the one the backward pass differentiates: sum over streams and timesteps of
// -ln p[target] (natural log, no 1/B), with the carried-in h(0), c(0) constant.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"

namespace lstm {

namespace {

struct Offsets { unsigned long long w, u, b, why, by; };

__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

// grid (n_probes, 2): blockIdx.y == 0 evaluates the loss at p - delta, 1 at p + delta.  out[2 * probe + y].
__global__ void __launch_bounds__(256)
k_window_loss_f64(const float* __restrict__ params, Offsets off, const float* __restrict__ h0, const float* __restrict__ c0,
                  const int* __restrict__ xs, const int* __restrict__ tg, int M, int N, int S, int B,
                  const unsigned long long* __restrict__ probe_pos, double delta, double* __restrict__ out) {
  extern __shared__ double sm[];
  double* h = sm;                 // [N]
  double* c = h + N;              // [N]
  double* g = c + N;              // [4N]
  double* y = g + 4 * (size_t)N;  // [M]
  const int tid = threadIdx.x, nt = blockDim.x;
  const unsigned long long ppos = probe_pos[blockIdx.x];      // index into the flat parameter vector
  const double shift = blockIdx.y ? delta : -delta;
  const size_t N4 = 4 * (size_t)N;
  // parameter entry i of the flat vector, as a double, with the probe's shift applied
  auto P = [&](unsigned long long i) -> double {
    const double v = (double)params[i];
    return i == ppos ? v + shift : v;
  };
  double loss = 0.0;   // thread 0 only
  for (int b = 0; b < B; b++) {
    for (int n = tid; n < N; n += nt) { h[n] = (double)h0[(size_t)b * N + n]; c[n] = (double)c0[(size_t)b * N + n]; }
    __syncthreads();
    for (int t = 1; t < S; t++) {
      const int x = xs[(size_t)t * B + b], k = tg[(size_t)t * B + b];
      // g = W x + U h(t-1) + b                                   (R/lstm.cc:176)
      for (size_t r = tid; r < N4; r += nt) {
        double acc = 0.0;
        for (int kk = 0; kk < N; kk++) acc += P(off.u + (size_t)kk * N4 + r) * h[kk];
        if (x >= 0) acc += P(off.w + (size_t)x * N4 + r);
        g[r] = acc + P(off.b + r);
      }
      __syncthreads();
      // gates, cell, hidden                                      (R/lstm.cc:179-192)
      for (int n = tid; n < N; n += nt) {
        const double gi = sigmoid_d(g[n]), go = sigmoid_d(g[N + n]), gf = sigmoid_d(g[2 * (size_t)N + n]);
        const double gu = tanh(g[3 * (size_t)N + n]);
        const double cc = tanh(gi * gu + gf * c[n]);              // the carried cell value is the tanh'd one
        c[n] = cc;
        h[n] = go * cc;
      }
      __syncthreads();
      // y = Why h + by                                           (R/lstm.cc:195)
      for (int m = tid; m < M; m += nt) {
        double acc = 0.0;
        for (int n = 0; n < N; n++) acc += P(off.why + (size_t)n * M + m) * h[n];
        y[m] = acc + P(off.by + m);
      }
      __syncthreads();
      // -ln softmax(y)[k] = ln sum exp(y) - y[k]                 (R/lstm.cc:199-207, in nats)
      if (tid == 0 && k >= 0) {
        double mx = y[0];
        for (int m = 1; m < M; m++) mx = fmax(mx, y[m]);
        double s = 0.0;
        for (int m = 0; m < M; m++) s += exp(y[m] - mx);
        loss += log(s) + mx - y[k];
      }
      __syncthreads();
    }
  }
  if (tid == 0) out[2 * (size_t)blockIdx.x + blockIdx.y] = loss;
}

}  // namespace

cudaError_t launch_window_loss_f64(const float* params, const size_t off[5], const float* h0, const float* c0, const int* xs,
                                   const int* tg, int M, int N, int S, int B, const unsigned long long* probe_pos,
                                   int n_probes, double delta, double* out, cudaStream_t st) {
  const size_t smem = (6 * (size_t)N + (size_t)M) * sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(k_window_loss_f64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  Offsets o{off[0], off[1], off[2], off[3], off[4]};
  k_window_loss_f64<<<dim3((unsigned)n_probes, 2), 256, smem, st>>>(params, o, h0, c0, xs, tg, M, N, S, B, probe_pos, delta, out);
  return cudaGetLastError();
}

}  // namespace lstm
