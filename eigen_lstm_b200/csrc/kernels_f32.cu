// kernels_f32.cu — fp32 SIMT kernels: the parity path (per-step loss within 1e-4 relative of the
// reference's Eigen CPU path) and the elementwise / data-pipeline kernels shared by both dtypes.
// Every kernel cites the reference statement(s) it computes (R/ = reference root).
#include "kernels.h"
#include "scalar_f32.cuh"

#include <math.h>

namespace lstm {

// ------------------------------------------------------------------------------------------------
// 64x64x16 SIMT tile engine: 256 threads, 4x4 register micro-tile per thread.
//   thread (ty = tid/16, tx = tid%16) owns rows ty*4+v (v<4) and columns u*16+tx (u<4)
//   (column interleave u*16+tx: the 4 columns of a thread are the 4 GATES of one hidden unit in
//   the recurrent kernels, and consecutive tx give coalesced stores everywhere)
// la(ii,k) / lb(k,jj) return the operand element (0 outside the matrix).
// ------------------------------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 16, TPAD = 4;

template <bool A_KCONTIG, bool B_KCONTIG, typename LA, typename LB>
__device__ __forceinline__ void tile_mainloop(float (&acc)[4][4], int K, LA la, LB lb) {
  __shared__ __align__(16) float As[TK][TM + TPAD];
  __shared__ __align__(16) float Bs[TK][TN + TPAD];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
#pragma unroll
  for (int v = 0; v < 4; v++)
#pragma unroll
    for (int u = 0; u < 4; u++) acc[v][u] = 0.f;
  float ar[4], br[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (A_KCONTIG) ar[r] = la((tid >> 4) + 16 * r, k0 + (tid & 15));
      else           ar[r] = la(tid & 63, k0 + (tid >> 6) + 4 * r);
      if (B_KCONTIG) br[r] = lb(k0 + (tid & 15), (tid >> 4) + 16 * r);
      else           br[r] = lb(k0 + (tid >> 6) + 4 * r, tid & 63);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if (A_KCONTIG) As[tid & 15][(tid >> 4) + 16 * r] = ar[r];
      else           As[(tid >> 6) + 4 * r][tid & 63] = ar[r];
      if (B_KCONTIG) Bs[tid & 15][(tid >> 4) + 16 * r] = br[r];
      else           Bs[(tid >> 6) + 4 * r][tid & 63] = br[r];
    }
    __syncthreads();
    if (k0 + TK < K) fetch(k0 + TK);
#pragma unroll
    for (int kk = 0; kk < TK; kk++) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      float b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) b[u] = Bs[kk][u * 16 + tx];
#pragma unroll
      for (int v = 0; v < 4; v++)
#pragma unroll
        for (int u = 0; u < 4; u++) acc[v][u] = fmaf(a[v], b[u], acc[v][u]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// K2 (fp32): one recurrent timestep, fused.  R/lstm.cc:176-192
//   g = W*x + U*h(t-1) + b ; sigmoid on [i o f], tanh on u ; c = tanh(i*u + f*c(t-1)) ; h = o*c
// grid (ceil(N/16), ceil(B/64)); tile rows = streams b, tile cols = 4 gates x 16 hidden units.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_step_fwd_f32(const float* __restrict__ U, const float* __restrict__ W,
                                                      const float* __restrict__ bias, const float* __restrict__ hprev,
                                                      const float* __restrict__ cprev, const int* __restrict__ x,
                                                      float* __restrict__ g, float* __restrict__ c,
                                                      float* __restrict__ h, int B, int N) {
  const int j0 = blockIdx.x * 16, b0 = blockIdx.y * 64;
  const int N4 = 4 * N;
  float acc[4][4];
  auto la = [&](int ii, int k) -> float {  // A(b,k) = h(t-1)[b][k]
    const int b = b0 + ii;
    return (b < B && k < N) ? hprev[(size_t)b * N + k] : 0.f;
  };
  auto lb = [&](int k, int jj) -> float {  // B(k, gate*16+unit) = U(gate*N + j, k)
    const int j = j0 + (jj & 15);
    return (k < N && j < N) ? U[(size_t)k * N4 + (jj >> 4) * N + j] : 0.f;
  };
  tile_mainloop<true, false>(acc, N, la, lb);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int j = j0 + tx;
  if (j >= N) return;
#pragma unroll
  for (int v = 0; v < 4; v++) {
    const int b = b0 + ty * 4 + v;
    if (b >= B) continue;
    const int xb = x[b];
    float pre[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const float wx = (xb >= 0) ? W[(size_t)xb * N4 + u * N + j] : 0.f;  // W*x for one-hot x = column x
      pre[u] = __fadd_rn(__fadd_rn(wx, acc[v][u]), bias[u * N + j]);
    }
    const float gi = logistic_f(pre[0]), go = logistic_f(pre[1]), gf = logistic_f(pre[2]);
    const float gu = tanhf(pre[3]);
    const float cp = cprev[(size_t)b * N + j];
    const float cc = tanhf(__fadd_rn(__fmul_rn(gi, gu), __fmul_rn(gf, cp)));  // carried value is tanh'd (:189)
    float* gb = g + (size_t)b * N4;
    gb[j] = gi; gb[N + j] = go; gb[2 * N + j] = gf; gb[3 * N + j] = gu;
    c[(size_t)b * N + j] = cc;
    h[(size_t)b * N + j] = __fmul_rn(go, cc);
  }
}

void launch_step_fwd_f32(const float* U, const float* W, const float* bias, const float* hprev,
                         const float* cprev, const int* x, float* g, float* c, float* h, int B, int N,
                         cudaStream_t st) {
  dim3 grid((N + 15) / 16, (B + 63) / 64);
  k_step_fwd_f32<<<grid, 256, 0, st>>>(U, W, bias, hprev, cprev, x, g, c, h, B, N);
}

// ------------------------------------------------------------------------------------------------
// K5 (fp32): one BPTT timestep, fused.  R/lstm.cc:228-256
//   dh = Why^T dy (precomputed dHy) + U^T dg(t+1) ; dc = (dh*o + dcnext) * (1 - c^2) ;
//   do, di, df, du through sigmoid'/tanh' ; dcnext = dc * f
// grid (ceil(N/64), ceil(B/64)); tile rows = streams, cols = hidden units; K = 4N.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_step_bwd_f32(const float* __restrict__ U, const float* __restrict__ dg_next,
                                                      const float* __restrict__ dHy, const float* __restrict__ g,
                                                      const float* __restrict__ c, const float* __restrict__ cprev,
                                                      float* __restrict__ dcnext, float* __restrict__ dg, int B,
                                                      int N, int first) {
  const int j0 = blockIdx.x * 64, b0 = blockIdx.y * 64;
  const int N4 = 4 * N;
  float acc[4][4];
  if (!first) {
    auto la = [&](int ii, int r) -> float {  // A(b,r) = dg(t+1)[b][r]
      const int b = b0 + ii;
      return (b < B && r < N4) ? dg_next[(size_t)b * N4 + r] : 0.f;
    };
    auto lb = [&](int r, int jj) -> float {  // B(r,j) = U(r,j)
      const int j = j0 + jj;
      return (r < N4 && j < N) ? U[(size_t)j * N4 + r] : 0.f;
    };
    tile_mainloop<true, true>(acc, N4, la, lb);
  } else {
#pragma unroll
    for (int v = 0; v < 4; v++)
#pragma unroll
      for (int u = 0; u < 4; u++) acc[v][u] = 0.f;
  }
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int v = 0; v < 4; v++) {
    const int b = b0 + ty * 4 + v;
    if (b >= B) continue;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = j0 + u * 16 + tx;
      if (j >= N) continue;
      const size_t bj = (size_t)b * N + j;
      const float* gb = g + (size_t)b * N4;
      const float gi = gb[j], go = gb[N + j], gf = gb[2 * N + j], gu = gb[3 * N + j];
      const float ct = c[bj], cp = cprev[bj];
      const float dh = __fadd_rn(dHy[bj], acc[v][u]);                             // :228
      float dc = __fadd_rn(__fmul_rn(dh, go), first ? 0.f : dcnext[bj]);          // :233
      dc = __fmul_rn(dc, tanh_prime_f(ct));                                       // :235
      float* dgb = dg + (size_t)b * N4;
      dgb[N + j] = __fmul_rn(__fmul_rn(dh, ct), logistic_prime_f(go));            // do :238,244
      dgb[j] = __fmul_rn(__fmul_rn(dc, gu), logistic_prime_f(gi));                // di :239,244
      dgb[2 * N + j] = __fmul_rn(__fmul_rn(dc, cp), logistic_prime_f(gf));        // df :240,244
      dgb[3 * N + j] = __fmul_rn(__fmul_rn(dc, gi), tanh_prime_f(gu));            // du :241,247
      dcnext[bj] = __fmul_rn(dc, gf);                                             // :256
    }
  }
}

void launch_step_bwd_f32(const float* U, const float* dg_next, const float* dHy_t, const float* g_t,
                         const float* c_t, const float* c_prev, float* dcnext, float* dg_t, int B, int N,
                         int first, cudaStream_t st) {
  dim3 grid((N + 63) / 64, (B + 63) / 64);
  k_step_bwd_f32<<<grid, 256, 0, st>>>(U, dg_next, dHy_t, g_t, c_t, c_prev, dcnext, dg_t, B, N, first);
}

// ------------------------------------------------------------------------------------------------
// generic strided GEMM (K3 logits, K4 dHy, K6a dU, K6c dWhy in the fp32 path)
// ------------------------------------------------------------------------------------------------
template <bool AK, bool BK>
__global__ void __launch_bounds__(256) k_gemm_f32(const float* __restrict__ A, long a_si, long a_sk,
                                                  const float* __restrict__ Bm, long b_sk, long b_sj,
                                                  float* __restrict__ C, long c_si, long c_sj,
                                                  const float* __restrict__ bias_j, int I, int J, int K) {
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  float acc[4][4];
  auto la = [&](int ii, int k) -> float {
    const int i = i0 + ii;
    return (i < I && k < K) ? A[(size_t)i * a_si + (size_t)k * a_sk] : 0.f;
  };
  auto lb = [&](int k, int jj) -> float {
    const int j = j0 + jj;
    return (j < J && k < K) ? Bm[(size_t)k * b_sk + (size_t)j * b_sj] : 0.f;
  };
  tile_mainloop<AK, BK>(acc, K, la, lb);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int v = 0; v < 4; v++) {
    const int i = i0 + ty * 4 + v;
    if (i >= I) continue;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = j0 + u * 16 + tx;
      if (j >= J) continue;
      float r = acc[v][u];
      if (bias_j) r = __fadd_rn(r, bias_j[j]);
      C[(size_t)i * c_si + (size_t)j * c_sj] = r;
    }
  }
}

void launch_gemm_f32(const float* A, long a_si, long a_sk, const float* Bm, long b_sk, long b_sj, float* C,
                     long c_si, long c_sj, const float* bias_j, int I, int J, int K, cudaStream_t st) {
  dim3 grid((J + 63) / 64, (I + 63) / 64);
  const bool ak = (a_sk == 1), bk = (b_sk == 1);
  if (ak && bk)       k_gemm_f32<true, true><<<grid, 256, 0, st>>>(A, a_si, a_sk, Bm, b_sk, b_sj, C, c_si, c_sj, bias_j, I, J, K);
  else if (ak && !bk) k_gemm_f32<true, false><<<grid, 256, 0, st>>>(A, a_si, a_sk, Bm, b_sk, b_sj, C, c_si, c_sj, bias_j, I, J, K);
  else if (!ak && bk) k_gemm_f32<false, true><<<grid, 256, 0, st>>>(A, a_si, a_sk, Bm, b_sk, b_sj, C, c_si, c_sj, bias_j, I, J, K);
  else                k_gemm_f32<false, false><<<grid, 256, 0, st>>>(A, a_si, a_sk, Bm, b_sk, b_sj, C, c_si, c_sj, bias_j, I, J, K);
}

// ------------------------------------------------------------------------------------------------
// K3 epilogue (fp32 path): softmax + loss + dy.  R/lstm.cc:199-207,225
//   p = exp(y) / sum(exp(y))  (no max shift, like the reference) ; surp = -log2 p[target] ;
//   dy = p - onehot(target).  One warp per (t,b) row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_softmax_ce_f32(float* __restrict__ y, const int* __restrict__ tg,
                                                        float* __restrict__ surp, int rows, int M,
                                                        const float* __restrict__ shift, int B) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* yr = y + (size_t)row * M;
  // shift != NULL: OV/lstm_eigen_class_batch/lstm.h:175 subtracts the maximum of the timestep's whole M x B logit matrix
  const float sh = shift ? shift[row / B] : 0.f;
  float s = 0.f;
  for (int m = lane; m < M; m += 32) {
    const float e = expf(shift ? __fsub_rn(yr[m], sh) : yr[m]);
    yr[m] = e;
    s += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const int k = tg[row];
  for (int m = lane; m < M; m += 32) {
    const float p = __fdiv_rn(yr[m], s);
    if (m == k) surp[row] = -log2f(p);
    yr[m] = (m == k) ? __fsub_rn(p, 1.0f) : p;
  }
  if (k < 0 && lane == 0) surp[row] = 0.f;
}

// maximum logit of every timestep's [B][M] block (the "global" softmax shift of OV/lstm_eigen_class_batch/lstm.h:175)
__global__ void __launch_bounds__(256) k_logit_max_f32(const float* __restrict__ y, float* __restrict__ shift, int n) {
  __shared__ float red[8];
  const float* yt = y + (size_t)blockIdx.x * n;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) mx = fmaxf(mx, yt[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, red[w]);
    shift[blockIdx.x] = mx;
  }
}

void launch_softmax_ce_f32(float* y, const int* tg, float* surp, int rows, int M, const float* shift, int B, cudaStream_t st) {
  k_softmax_ce_f32<<<(rows + 7) / 8, 256, 0, st>>>(y, tg, surp, rows, M, shift, B);
}
void launch_logit_max_f32(const float* y, float* shift, int T, int per_t, cudaStream_t st) {
  k_logit_max_f32<<<T, 256, 0, st>>>(y, shift, per_t);
}

// loss = sum_t (float)(sum_b surp[t][b]) / (float)B    (OV/lstm_eigen_opt/lstm.cc:246-249)
// The result goes to ring[*iter % cap] and the kernel bumps *iter itself, so a captured CUDA graph of one training
// iteration can be replayed unchanged.
// mode 1 (OV/lstm_eigen_class_batch/lstm.cc:308-319): the LAST timestep only, in nats: sum_b -ln p[target] / B
__global__ void __launch_bounds__(1024) k_loss_reduce(const float* __restrict__ surp, int T, int B,
                                                      double* __restrict__ ring, unsigned long long cap,
                                                      unsigned long long* __restrict__ iter, int mode) {
  extern __shared__ float st_sum[];  // [T]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < T; t += 32) {
    float s = 0.f;
    for (int b = lane; b < B; b += 32) s += surp[(size_t)t * B + b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) st_sum[t] = s / (float)B;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0;
    if (mode == 1) l = (double)st_sum[T - 1] * 0.6931471805599453;
    else for (int t = 0; t < T; t++) l += (double)st_sum[t];
    const unsigned long long it = iter[0];
    ring[it % cap] = l;
    iter[0] = it + 1;
  }
}

void launch_loss_reduce(const float* surp, int T, int B, double* ring, size_t cap, unsigned long long* iter, int mode, cudaStream_t st) {
  k_loss_reduce<<<1, 1024, T * sizeof(float), st>>>(surp, T, B, ring, (unsigned long long)cap, iter, mode);
}

// ------------------------------------------------------------------------------------------------
// K6b (fp32 path): dW += dg * x^T for one-hot x is a segmented column sum (R/lstm.cc:251).
// One CTA per (symbol m, 256-row chunk of the 4N gate rows); deterministic (fixed (t,b) order).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dw_scatter_f32(const float* __restrict__ dG, const int* __restrict__ xs,
                                                        float* __restrict__ dW, int rows, int N4) {
  const int m = blockIdx.y;
  const int r = blockIdx.x * 256 + threadIdx.x;
  __shared__ int sx[256];
  float acc = 0.f;
  for (int base = 0; base < rows; base += 256) {
    const int i = base + threadIdx.x;
    sx[threadIdx.x] = (i < rows) ? xs[i] : -1;
    __syncthreads();
    const int lim = min(256, rows - base);
    if (r < N4)
      for (int q = 0; q < lim; q++)
        if (sx[q] == m) acc += dG[(size_t)(base + q) * N4 + r];
    __syncthreads();
  }
  if (r < N4) dW[(size_t)m * N4 + r] = acc;
}

void launch_dw_scatter_f32(const float* dG, const int* xs, float* dW, int rows, int N4, int M, cudaStream_t st) {
  dim3 grid((N4 + 255) / 256, M);
  k_dw_scatter_f32<<<grid, 256, 0, st>>>(dG, xs, dW, rows, N4);
}

// db += dg (R/lstm.cc:252), dby += dy (:227): column sums of a [I][J] row-major matrix
__global__ void __launch_bounds__(1024) k_colsum_f32(const float* __restrict__ X, float* __restrict__ out, int I, int J) {
  __shared__ float red[32][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (j < J)
    for (int i = threadIdx.y; i < I; i += 32) s += X[(size_t)i * J + j];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && j < J) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 32; q++) t += red[q][threadIdx.x];
    out[j] = t;
  }
}

void launch_colsum_f32(const float* X, float* out, int I, int J, cudaStream_t st) {
  k_colsum_f32<<<(J + 31) / 32, dim3(32, 32), 0, st>>>(X, out, I, J);
}

// ------------------------------------------------------------------------------------------------
// K7 Adagrad, R/lstm.cc:259-272:  m += d*d ; p -= lr * d / sqrtf(m + eps)   (eps added in double,
// :25,46-48).  HBM-bound: 20 B/param (read d, read+write m, read+write p); 128-bit accesses.
// clip > 0 clamps d first (north-star addition; 0 = the reference's behaviour).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_adagrad_f32(float* __restrict__ p, const float* __restrict__ d,
                                                     float* __restrict__ m, size_t n, float lr, double eps,
                                                     float clip) {
  const size_t n4 = n >> 2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 dv = reinterpret_cast<const float4*>(d)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    adagrad_one(pv.x, dv.x, mv.x, lr, eps, clip);
    adagrad_one(pv.y, dv.y, mv.y, lr, eps, clip);
    adagrad_one(pv.z, dv.z, mv.z, lr, eps, clip);
    adagrad_one(pv.w, dv.w, mv.w, lr, eps, clip);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
  }
  // tail
  for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adagrad_one(p[i], d[i], m[i], lr, eps, clip);
}

void launch_adagrad_f32(float* p, const float* d, float* m, size_t n, float lr, double eps, float clip,
                        cudaStream_t st) {
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;  // 8 resident CTAs per SM on 148 SMs, grid-stride
  if (blocks < 1) blocks = 1;
  k_adagrad_f32<<<(unsigned)blocks, 256, 0, st>>>(p, d, m, n, lr, eps, clip);
}

__global__ void k_fill_f32(float* p, float v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
void launch_fill_f32(float* p, float v, size_t n, cudaStream_t st) {
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_fill_f32<<<(unsigned)blocks, 256, 0, st>>>(p, v, n);
}

// ------------------------------------------------------------------------------------------------
// K10 data pipeline: text bytes + stream positions -> window index tensors.
// Closed form of the reference's shift loop (R/lstm.cc:155-170; OV/lstm_eigen_opt/lstm.cc:190-213):
// stream b has consumed v events; event q of stream b is E_b[q] = text[S + (pos0_b - S + q) mod (len - S)]
// (positions wrap to S at the end of the text); target column s holds E_b[v-1-(S-1-s)], input column
// s holds target column s-1; not-yet-filled columns are -1 (the all-zero one-hot).
// The kernel owns the event counter (so a captured CUDA graph can replay it): every CTA reads v_counter[0]; the LAST CTA to
// finish (ticket in v_counter[1]) advances it and re-arms the ticket.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_window_advance(const uint8_t* __restrict__ text, unsigned long long len,
                                                        const unsigned long long* __restrict__ pos0,
                                                        unsigned long long* __restrict__ v_counter, int stride,
                                                        int S, int B, int* __restrict__ xs, int* __restrict__ tg) {
  const long long v = (long long)(v_counter[0] + (unsigned long long)stride);
  const unsigned long long span = len - (unsigned long long)S;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < S * B; e += gridDim.x * blockDim.x) {
    const int s = e / B, b = e - s * B;
    const long long qt = v - 1 - (S - 1 - s);  // event index held by target column s
    const long long qx = qt - 1;               // input column s = target column s-1
    const unsigned long long off = pos0[b] - (unsigned long long)S;
    tg[e] = (qt >= 0) ? (int)text[S + (off + (unsigned long long)qt) % span] : -1;
    xs[e] = (qx >= 0) ? (int)text[S + (off + (unsigned long long)qx) % span] : -1;
  }
  __syncthreads();                             // every thread of this CTA has read the counter
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(v_counter + 1, 1ull) == (unsigned long long)gridDim.x - 1) {   // all CTAs have read it
      v_counter[0] = (unsigned long long)v;
      v_counter[1] = 0ull;
    }
  }
}

void launch_window_advance(const uint8_t* text, size_t len, const unsigned long long* pos0,
                           unsigned long long* v_counter, int stride, int S, int B, int* xs, int* tg,
                           cudaStream_t st) {
  int grid = (S * B + 255) / 256;
  if (grid > 148) grid = 148;
  k_window_advance<<<grid, 256, 0, st>>>(text, (unsigned long long)len, pos0, v_counter, stride, S, B, xs, tg);
}

// ------------------------------------------------------------------------------------------------
// batch-1 serial recurrence, single CTA (first version of K9 / test()).
//   eval  : for ii in [0, n-1): step on text[ii], bits += -log2 p[text[ii+1]]   (class_CUDA/lstm.cc:661-720)
//   sample: for i in [0, n): p from current h, draw, emit, step on the drawn byte (R/lstm.cc:313-350)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_recur_b1_f32(const float* __restrict__ W, const float* __restrict__ U,
                                                       const float* __restrict__ bias, const float* __restrict__ Why,
                                                       const float* __restrict__ by, int M, int N, int mode,
                                                       const uint8_t* __restrict__ text, size_t n,
                                                       const float* __restrict__ uniforms, const float* __restrict__ h0,
                                                       const float* __restrict__ c0, uint8_t* __restrict__ out,
                                                       double* __restrict__ bits_out) {
  extern __shared__ float sm[];
  float* h = sm;              // [N]
  float* c = h + N;           // [N]
  float* g = c + N;           // [4N]
  float* p = g + 4 * N;       // [M]
  float* part = p + M;        // [4][M] partial logits
  __shared__ float s_sum;
  __shared__ int s_index;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N4 = 4 * N;
  for (int j = tid; j < N; j += nt) { h[j] = h0 ? h0[j] : 0.f; c[j] = c0 ? c0[j] : 0.f; }
  __syncthreads();
  double bits = 0.0;

  auto cell = [&](int x) {
    for (int r = tid; r < N4; r += nt) {
      float acc = 0.f;
      for (int k = 0; k < N; k++) acc = fmaf(U[(size_t)k * N4 + r], h[k], acc);
      const float wx = (x >= 0) ? W[(size_t)x * N4 + r] : 0.f;
      const float pre = __fadd_rn(__fadd_rn(wx, acc), bias[r]);
      g[r] = (r < 3 * N) ? logistic_f(pre) : tanhf(pre);
    }
    __syncthreads();
    for (int j = tid; j < N; j += nt) {
      const float cc = tanhf(__fadd_rn(__fmul_rn(g[j], g[3 * N + j]), __fmul_rn(g[2 * N + j], c[j])));
      c[j] = cc;
      h[j] = __fmul_rn(g[N + j], cc);
    }
    __syncthreads();
  };
  auto softmax = [&]() {
    // y = Why*h + by with the N range split in 4 chunks over the 1024 threads (needs M <= 256)
    const int chunk = tid / 256, m = tid % 256;
    if (m < M) {
      float acc = 0.f;
      const int n0 = (int)((long)N * chunk / 4), n1 = (int)((long)N * (chunk + 1) / 4);
      for (int q = n0; q < n1; q++) acc = fmaf(Why[(size_t)q * M + m], h[q], acc);
      part[chunk * M + m] = acc;
    }
    __syncthreads();
    if (tid < M) p[tid] = expf(__fadd_rn(((part[tid] + part[M + tid]) + part[2 * M + tid]) + part[3 * M + tid], by[tid]));
    __syncthreads();
    if (tid < 32) {
      float s = 0.f;
      for (int q = tid; q < M; q += 32) s += p[q];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (tid == 0) s_sum = s;
    }
    __syncthreads();
    if (tid < M) p[tid] = __fdiv_rn(p[tid], s_sum);
    __syncthreads();
  };

  if (mode == 0) {
    for (size_t ii = 0; ii + 1 < n; ii++) {
      cell((int)text[ii]);
      softmax();
      if (tid == 0) bits += -log2((double)p[text[ii + 1]]);
    }
    if (tid == 0) bits_out[0] = bits;
  } else {
    for (size_t i = 0; i < n; i++) {
      softmax();
      if (tid == 0) {
        int index = 0;
        if (mode == 2) {
          for (int q = 1; q < M; q++) if (p[q] > p[index]) index = q;
        } else {
          const float r = uniforms[i];
          float cdf = 0.f;
          for (int q = 0; q < M; q++) {  // cdf(ii) = cdf(ii-1) + probs(ii), first r < cdf (R/lstm.cc:321-338)
            cdf = (q == 0) ? p[0] : __fadd_rn(cdf, p[q]);
            if (r < cdf) { index = q; break; }
          }
        }
        s_index = index;
        out[i] = (uint8_t)index;
      }
      __syncthreads();
      cell(s_index);
    }
  }
}

void launch_recur_b1_f32(const float* W, const float* U, const float* bias, const float* Why, const float* by,
                         int M, int N, int mode, const uint8_t* text, size_t n, const float* uniforms,
                         const float* h0, const float* c0, uint8_t* out, double* bits_out, cudaStream_t st) {
  const size_t smem = sizeof(float) * ((size_t)6 * N + (size_t)5 * M);
  cudaFuncSetAttribute(k_recur_b1_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_recur_b1_f32<<<1, 1024, smem, st>>>(W, U, bias, Why, by, M, N, mode, text, n, uniforms, h0, c0, out, bits_out);
}

// ------------------------------------------------------------------------------------------------
// K9: PERSISTENT batch-1 recurrence (sampling R/lstm.cc:293-356, test() class_CUDA/lstm.cc:661-720).
// Latency-bound GEMV chain: one cooperative grid of G CTAs lives for the whole sequence; CTA g owns
// hidden units [g*UPC, (g+1)*UPC): their 4*UPC rows of U stay RESIDENT IN SHARED MEMORY across all
// timesteps (when they fit), as do its rows of Why.  Per character: h(t) is exchanged through a
// double-buffered global vector and one grid-wide barrier; sampling needs the full softmax, i.e. a second
// barrier; evaluation defers the softmax normalisation to the end (per-step partial sums) and needs one.
// ------------------------------------------------------------------------------------------------
// One arriving thread (release: cumulative over the __syncthreads-ordered stores of the CTA) and one polling WARP per CTA:
// scripts/gridbar_bench.cu measures 1.17 us per barrier over 128 co-resident CTAs for this scheme against 4.4 us for
// atomicAdd + last-arriver flag.  Bounded: a protocol bug must not hang the GPU.
__device__ __forceinline__ void grid_arrive(unsigned int* counter) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
}
__device__ __forceinline__ void grid_wait(unsigned int* counter, unsigned int target) {
  if (threadIdx.x < 32) {
    const long long t0 = clock64();
    for (;;) {
      unsigned int v = target;
      if (threadIdx.x == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (__all_sync(0xffffffffu, (int)(v - target) >= 0)) break;
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  }
  __syncthreads();
}
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  grid_arrive(counter);
  grid_wait(counter, target);
}

__host__ __device__ inline int recur_pitch(int N) { return (N + 31) / 32 * 32 + 8; }

struct RecurArgs {
  const float *W, *U, *bias, *Why, *by;   // fp32 masters, column-major like the reference
  int M, N, mode;                          // mode 0 eval, 1 sample, 2 greedy
  int UPC, rows_resident;                  // hidden units per CTA; 1 if this CTA's U rows live in smem
  int w_resident;                          // 1 if this CTA's W rows ([4*UPC][M]) live in smem too
  const uint8_t* text; size_t n;
  const float* uniforms; const float *h0, *c0;
  uint8_t* out;
  float* hbuf;                             // [2][N]
  float* ebuf;                             // [2][M]   exp(y) of the current step (sampling)
  float* sum_part;                         // [n][G]   eval: per-step per-CTA sum of exp(y)
  float* y_tgt;                            // [n]      eval: logit of the target byte
  float* c_out;                            // [N] final cell state (optional)
  unsigned int* bar;
};

__global__ void __launch_bounds__(256, 1) k_recur_persist(const RecurArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int N = a.N, M = a.M, N4 = 4 * N, G = gridDim.x, g = blockIdx.x, tid = threadIdx.x;
  const int UPC = a.UPC, R = 4 * UPC;                 // gate rows owned by this CTA (unit-major: r = 4*u + gate)
  const int MPC = (M + G - 1) / G;                    // logit rows owned by this CTA
  float* sh = sm;                                      // [N]   h(t)
  float* sU = sh + N;                                  // [R][N+1] resident rows of U (pitch N+1: conflict-free)
  // row pitch of the resident rows: == 8 (mod 32) words, so that the 4 rows x 8 lanes of a warp pass hit 32 distinct banks
  // (pitch N + 1 made rows r and lanes l with equal r + l collide: up to 4-way conflicts on every load of the matvec)
  const int UP = recur_pitch(N);
  float* sWhy = sU + (a.rows_resident ? (size_t)R * UP : 0);   // [MPC][UP]
  float* sW = sWhy + (size_t)MPC * UP;                 // [R][M] resident rows of W: the one-hot product W x is one smem read per row
  float* sg = sW + (a.w_resident ? (size_t)R * M : 0); // [R] gate pre-activations
  float* sc = sg + R;                                  // [UPC] cell state of the owned units
  float* se = sc + UPC;                                // [M] exp(y) of the current step
  __shared__ float s_sum;
  __shared__ int s_index;
  const int j0 = g * UPC;
  // one-time: make the weights resident
  if (a.rows_resident)
    for (int idx = tid; idx < R * N; idx += blockDim.x) {
      const int r = idx / N, k = idx - r * N;
      const int u = r >> 2, gate = r & 3, j = j0 + u;
      sU[(size_t)r * UP + k] = (j < N) ? a.U[(size_t)k * N4 + (size_t)gate * N + j] : 0.f;
    }
  for (int idx = tid; idx < MPC * N; idx += blockDim.x) {
    const int mm = idx / N, k = idx - mm * N, m = g * MPC + mm;
    sWhy[(size_t)mm * UP + k] = (m < M) ? a.Why[(size_t)k * M + m] : 0.f;
  }
  if (a.w_resident)
    for (int idx = tid; idx < R * M; idx += blockDim.x) {
      const int r = idx / M, m = idx - r * M;
      const int u = r >> 2, gate = r & 3, j = j0 + u;
      sW[idx] = (j < N) ? a.W[(size_t)m * N4 + (size_t)gate * N + j] : 0.f;
    }
  for (int u = tid; u < UPC; u += blockDim.x) sc[u] = (a.c0 && j0 + u < N) ? a.c0[j0 + u] : 0.f;
  if (g == 0)
    for (int k = tid; k < N; k += blockDim.x) a.hbuf[k] = a.h0 ? a.h0[k] : 0.f;
  unsigned int epoch = 0;
  grid_barrier(a.bar, ++epoch * G);

  auto load_h = [&](int buf) {
    if ((N & 3) == 0) {
      const float4* src = reinterpret_cast<const float4*>(a.hbuf + (size_t)buf * N);
      for (int k = tid; k < (N >> 2); k += blockDim.x) reinterpret_cast<float4*>(sh)[k] = __ldcg(src + k);
    } else {
      for (int k = tid; k < N; k += blockDim.x) sh[k] = __ldcg(a.hbuf + (size_t)buf * N + k);
    }
    __syncthreads();
  };
  // 8 lanes per gate row (4 rows per warp and pass; every lane runs the shuffles): dot(U_row, h) + W[x][row] + b[row];
  // then the cell update for the owned units
  const int warp = tid >> 5, lane = tid & 31, lane8 = lane & 7, nwarps = blockDim.x >> 5;
  // U h for the owned gate rows: 8 lanes per row (4 rows per warp and pass; every lane runs the shuffles).  Does NOT depend on
  // the input byte, so when sampling it runs BEFORE the byte is drawn, off the critical path.
  auto matvec = [&]() {
    for (int base = warp * 4; base < R; base += nwarps * 4) {
      const int r = base + (lane >> 3);
      const int u = r >> 2, gate = r & 3, j = j0 + u;
      const bool ok = r < R && j < N;
      float acc = 0.f;
      if (ok) {
        if (a.rows_resident && (N & 31) == 0) {
          // 128-bit shared-memory loads: a quarter-warp (the 8 lanes of one row) reads 32 consecutive words = all 32 banks, and
          // the four rows of a warp read the SAME h words (one broadcast wavefront): 5 wavefronts per 512 B of U instead of 8
          const float4* ur4 = reinterpret_cast<const float4*>(sU + (size_t)r * UP);
          const float4* sh4 = reinterpret_cast<const float4*>(sh);
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          for (int k4 = lane8; k4 < (N >> 2); k4 += 8) {
            const float4 u = ur4[k4], hv = sh4[k4];
            a0 = fmaf(u.x, hv.x, a0); a1 = fmaf(u.y, hv.y, a1); a2 = fmaf(u.z, hv.z, a2); a3 = fmaf(u.w, hv.w, a3);
          }
          acc = (a0 + a1) + (a2 + a3);
        } else if (a.rows_resident) {
          const float* ur = sU + (size_t)r * UP;
          for (int k = lane8; k < N; k += 8) acc = fmaf(ur[k], sh[k], acc);
        } else {
          const float* ug = a.U + (size_t)gate * N + j;
          for (int k = lane8; k < N; k += 8) acc = fmaf(__ldg(ug + (size_t)k * N4), sh[k], acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      if (lane8 == 0 && ok) sg[r] = acc;
    }
  };
  // g = W[:, x] + U h + b (R/lstm.cc:176), gates, cell update for the owned units; needs a __syncthreads() after matvec()
  auto finish = [&](int x, int buf_out) {
    for (int r = tid; r < R; r += blockDim.x) {
      const int u = r >> 2, gate = r & 3, j = j0 + u;
      if (j < N) {
        const size_t row = (size_t)gate * N + j;
        const float wx = (x >= 0) ? (a.w_resident ? sW[(size_t)r * M + x] : a.W[(size_t)x * N4 + row]) : 0.f;
        const float pre = __fadd_rn(__fadd_rn(wx, sg[r]), a.bias[row]);
        sg[r] = (gate < 3) ? logistic_f(pre) : tanhf(pre);
      }
    }
    __syncthreads();
    for (int u = tid; u < UPC; u += blockDim.x) {
      const int j = j0 + u;
      if (j < N) {
        const float gi = sg[4 * u], go = sg[4 * u + 1], gf = sg[4 * u + 2], gu = sg[4 * u + 3];   // gate order i o f u
        const float cc = tanhf(__fadd_rn(__fmul_rn(gi, gu), __fmul_rn(gf, sc[u])));
        sc[u] = cc;
        a.hbuf[(size_t)buf_out * N + j] = __fmul_rn(go, cc);
      }
    }
  };
  // exp(Why*h + by) for the owned logit rows: one full warp per row (there are only ceil(M/G) ~ 2 rows per CTA)
  auto logits = [&](float* e_out, float* y_out) {
    for (int mm = warp; mm < MPC; mm += nwarps) {
      const int m = g * MPC + mm;
      float acc = 0.f;
      if (m < M) {
        const float* wr = sWhy + (size_t)mm * UP;
        for (int k = lane; k < N; k += 32) acc = fmaf(wr[k], sh[k], acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0 && m < M) {
        const float y = __fadd_rn(acc, a.by[m]);
        e_out[mm] = expf(y);
        if (y_out) y_out[mm] = y;
      }
    }
  };

  if (a.mode == 0) {
    // evaluation: ONE phase and ONE barrier per character.  With h(i) in shared memory a phase scores text[i] against the
    // output of the previous step (logits of h(i), normalisation deferred) AND steps on text[i] (cell) — the two do not depend on
    // each other.
    float* my_e = se;          // [MPC]
    float* my_y = se + MPC;    // [MPC]
    for (size_t i = 0; i < a.n; i++) {
      const int buf = (int)(i & 1);
      load_h(buf);
      // both products read h(i) only: the gate rows (all warps) and the owned logit rows (one more row for the first warps) run
      // back to back, one barrier for both
      if (i + 1 < a.n) matvec();
      if (i > 0) logits(my_e, my_y);
      __syncthreads();
      if (i > 0) {
        if (tid == 0) {
          float s = 0.f;
          const int tgt = (int)a.text[i];
          for (int mm = 0; mm < MPC; mm++) {
            const int m = g * MPC + mm;
            if (m < M) { s += my_e[mm]; if (m == tgt) a.y_tgt[i - 1] = my_y[mm]; }
          }
          a.sum_part[(i - 1) * G + g] = s;
        }
      }
      if (i + 1 < a.n) {
        finish((int)a.text[i], buf ^ 1);
        grid_barrier(a.bar, ++epoch * G);
      }
    }
  } else {
    // sampling: p from the current h, draw, emit, step on the drawn byte; TWO barriers per character.  U h — the expensive part
    // of the step — does not depend on the drawn byte: it is computed next to the logits, before the first barrier.
    for (size_t i = 0; i < a.n; i++) {
      const int buf = (int)(i & 1);
      const float r_draw = (tid == 0 && a.mode == 1) ? a.uniforms[i] : 0.f;   // requested early, used after the softmax barrier
      load_h(buf);
      logits(se, nullptr);                                           // se[0..MPC) = owned exp(y)
      __syncthreads();
      for (int mm = tid; mm < MPC; mm += blockDim.x)
        if (g * MPC + mm < M) a.ebuf[(size_t)buf * M + g * MPC + mm] = se[mm];
      grid_arrive(a.bar);                                            // split barrier: U h is computed while the other CTAs arrive
      matvec();
      grid_wait(a.bar, ++epoch * G);
      for (int m = tid; m < M; m += blockDim.x) se[m] = __ldcg(a.ebuf + (size_t)buf * M + m);
      __syncthreads();
      if (tid < 32) {                                                // same reduction shape as the 1-CTA kernel
        float s = 0.f;
        for (int q = tid; q < M; q += 32) s += se[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) s_sum = s;
      }
      __syncthreads();
      {
        const float sum = s_sum;
        for (int m = tid; m < M; m += blockDim.x) se[m] = __fdiv_rn(se[m], sum);   // probs = exp(y) / sum, in parallel
      }
      __syncthreads();
      if (tid == 0) {
        int index = 0;
        if (a.mode == 2) {
          float best = se[0];
          for (int q = 1; q < M; q++) if (se[q] > best) { best = se[q]; index = q; }
        } else {
          // sequential cdf, first r < cdf (R/lstm.cc:321-338): the fp32 adds keep the reference's order; eight probabilities
          // are loaded per round so that the shared-memory latency is paid once per eight dependent adds instead of per add
          const float r = r_draw;
          float cdf = 0.f;
          bool found = false;
          for (int q0 = 0; q0 < M && !found; q0 += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = (q0 + k < M) ? se[q0 + k] : 0.f;
#pragma unroll
            for (int k = 0; k < 8; k++) {
              cdf = (q0 + k == 0) ? v[0] : __fadd_rn(cdf, v[k]);
              if (!found && q0 + k < M && r < cdf) { index = q0 + k; found = true; }
            }
          }
        }
        s_index = index;
        if (g == 0) a.out[i] = (uint8_t)index;
      }
      __syncthreads();
      finish(s_index, buf ^ 1);
      grid_barrier(a.bar, ++epoch * G);
    }
  }
  // final state: h is in hbuf[(steps) & 1]; c of the owned units goes to c_out (lets the host chain chunks)
  if (a.c_out)
    for (int u = tid; u < UPC; u += blockDim.x)
      if (j0 + u < N) a.c_out[j0 + u] = sc[u];
}

// host side: sizes, cooperative launch.  Returns cudaSuccess or the launch error.
size_t recur_persist_smem(int M, int N, int G, int UPC, int resident, int w_resident) {
  const int MPC = (M + G - 1) / G;
  const size_t UP = (size_t)recur_pitch(N);
  return sizeof(float) * ((size_t)N + (resident ? (size_t)4 * UPC * UP : 0) + (size_t)MPC * UP +
                          (w_resident ? (size_t)4 * UPC * M : 0) + 4 * UPC + UPC + M + 2 * MPC + 64);
}

cudaError_t launch_recur_persist(const float* W, const float* U, const float* bias, const float* Why, const float* by, int M,
                                 int N, int mode, const uint8_t* text, size_t n, const float* uniforms, const float* h0,
                                 const float* c0, uint8_t* out, float* hbuf, float* ebuf, float* sum_part, float* y_tgt,
                                 float* c_out, unsigned int* bar, int num_sms, int* G_out, cudaStream_t st) {
  int G = num_sms;
  while (G > 1 && (N + G - 1) / G * (G - 1) >= N) G--;       // no CTA without units
  if (N % 128 == 0 && num_sms >= 128) G = 128;                // even split: 128 CTAs x N/128 units
  const int UPC = (N + G - 1) / G;
  int resident = 1, w_resident = 1;
  size_t smem = recur_persist_smem(M, N, G, UPC, 1, 1);
  if (smem > 227 * 1024) { w_resident = 0; smem = recur_persist_smem(M, N, G, UPC, 1, 0); }
  if (smem > 227 * 1024) { resident = 0; smem = recur_persist_smem(M, N, G, UPC, 0, 0); }
  RecurArgs a;
  a.W = W; a.U = U; a.bias = bias; a.Why = Why; a.by = by; a.M = M; a.N = N; a.mode = mode;
  a.UPC = UPC; a.rows_resident = resident; a.w_resident = w_resident; a.text = text; a.n = n; a.uniforms = uniforms; a.h0 = h0; a.c0 = c0;
  a.out = out; a.hbuf = hbuf; a.ebuf = ebuf; a.sum_part = sum_part; a.y_tgt = y_tgt; a.c_out = c_out; a.bar = bar;
  cudaError_t e = cudaFuncSetAttribute(k_recur_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&a};
  *G_out = G;
  return cudaLaunchCooperativeKernel((void*)k_recur_persist, dim3(G), dim3(256), args, smem, st);
}

// eval epilogue: bits = sum_i -log2( exp(y_tgt[i]) / sum_g sum_part[i][g] )
__global__ void k_eval_finish(const float* __restrict__ sum_part, const float* __restrict__ y_tgt, size_t steps, int G,
                              double* __restrict__ bits_out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (size_t i = threadIdx.x; i < steps; i += blockDim.x) {
    float s = 0.f;
    for (int g = 0; g < G; g++) s += sum_part[i * G + g];
    const float p = __fdiv_rn(expf(y_tgt[i]), s);
    acc += -log2((double)p);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) bits_out[0] += red[0];
}
void launch_eval_finish(const float* sum_part, const float* y_tgt, size_t steps, int G, double* bits_out, cudaStream_t st) {
  k_eval_finish<<<1, 256, 0, st>>>(sum_part, y_tgt, steps, G, bits_out);
}

}  // namespace lstm
