// kernels.h — launcher prototypes of the hand-written sm_100a kernels (host-callable, C++).
// Layouts (all device memory, T = S-1 active timesteps, "slot" = timestep index):
//   params/grads/adagrad memory : column-major float32 like the reference's Eigen matrices:
//        W[m*4N + r], U[k*4N + r], b[r], Why[n*M + m], by[m]   (R/lstm.cc:65-70)
//   Hs, Cs : [(T+1)][B][N]   h_t / c_t of stream b at slot t (slot 0 = carried-in state)
//   Gs     : [T][B][4N]      activated gates [i o f u] of timestep t at slot t-1 (R/lstm.cc:77)
//   dY     : [T][B][M]       logits -> probs - onehot(target) (R/lstm.cc:195-201,225)
//   dHy    : [T][B][N]       Why^T dy (R/lstm.cc:228 without dhnext)
//   dG     : [T][B][4N]      gate pre-activation gradients (R/lstm.cc:238-247)
//   xs, tg : [S][B] int32    input / target byte, -1 = all-zero column; row 0 unused
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lstm {

// ---- K10 data pipeline (R/lstm.cc:155-170, OV/lstm_eigen_opt/lstm.cc:190-213) ----
void launch_window_advance(const uint8_t* text, size_t len, const unsigned long long* pos0,
                           unsigned long long* v_counter, int stride, int S, int B, int* xs, int* tg,
                           cudaStream_t st);

// ---- fp32 SIMT path ----
void launch_step_fwd_f32(const float* U, const float* W, const float* bias, const float* hprev,
                         const float* cprev, const int* x, float* g, float* c, float* h, int B, int N,
                         cudaStream_t st);
void launch_step_bwd_f32(const float* U, const float* dg_next, const float* dHy_t, const float* g_t,
                         const float* c_t, const float* c_prev, float* dcnext, float* dg_t, int B, int N,
                         int first, cudaStream_t st);
// C(i,j) = sum_k A(i,k) B(k,j) (+ bias[j]); element strides given explicitly
void launch_gemm_f32(const float* A, long a_si, long a_sk, const float* Bm, long b_sk, long b_sj, float* C,
                     long c_si, long c_sj, const float* bias_j, int I, int J, int K, cudaStream_t st);
// rows of `y` ([rows][M] logits) -> p - onehot(tg) in place; surp[row] = -log2 p[tg] (0 if tg < 0)
// shift (optional): [rows / B] value subtracted from every logit of the timestep before exp (global-max softmax shift)
void launch_softmax_ce_f32(float* y, const int* tg, float* surp, int rows, int M, const float* shift, int B, cudaStream_t st);
// shift[t] = max over the timestep's per_t = B*M logits (OV/lstm_eigen_class_batch/lstm.h:175)
void launch_logit_max_f32(const float* y, float* shift, int T, int per_t, cudaStream_t st);
// loss = sum_t (float)(sum_b surp[t][b]) / B  -> ring[*iter % cap] (double); *iter += 1
// mode 1: last timestep only, natural log (OV/lstm_eigen_class_batch/lstm.cc:308-319)
void launch_loss_reduce(const float* surp, int T, int B, double* ring, size_t cap, unsigned long long* iter, int mode, cudaStream_t st);
// dW[m*4N + r] = sum over (t,b) with xs[t][b] == m of dG[t][b][r]   (R/lstm.cc:251 for one-hot x)
void launch_dw_scatter_f32(const float* dG, const int* xs, float* dW, int rows, int N4, int M, cudaStream_t st);
// out[j] = sum_i X[i*J + j]
void launch_colsum_f32(const float* X, float* out, int I, int J, cudaStream_t st);
// K7 Adagrad (R/lstm.cc:259-272) over a flat parameter vector
void launch_adagrad_f32(float* p, const float* d, float* m, size_t n, float lr, double eps, float clip,
                        cudaStream_t st);
// batch-1 serial recurrence: mode 0 = eval (test(), OV/lstm_eigen_class_CUDA/lstm.cc:661-720),
// mode 1 = sample (R/lstm.cc:293-356), mode 2 = greedy sample
void launch_recur_b1_f32(const float* W, const float* U, const float* bias, const float* Why, const float* by,
                         int M, int N, int mode, const uint8_t* text, size_t n, const float* uniforms,
                         const float* h0, const float* c0, uint8_t* out, double* bits_out, cudaStream_t st);
void launch_fill_f32(float* p, float v, size_t n, cudaStream_t st);
// K9 persistent multi-CTA version (cooperative launch; weights resident in shared memory; grid barrier per step).
// hbuf [2][N], ebuf [2][M], sum_part [n][G] + y_tgt [n] (eval only), bar: zeroed counter.  The state after the call is
// NOT returned (sampling / test() start from h0, c0 each time, like the reference).
cudaError_t launch_recur_persist(const float* W, const float* U, const float* bias, const float* Why, const float* by, int M,
                                 int N, int mode, const uint8_t* text, size_t n, const float* uniforms, const float* h0,
                                 const float* c0, uint8_t* out, float* hbuf, float* ebuf, float* sum_part, float* y_tgt,
                                 float* c_out, unsigned int* bar, int num_sms, int* G_out, cudaStream_t st);
// gradient check (gradcheck.cu): loss of the current window in double with ONE flat-parameter entry shifted by -/+ delta;
// out[2 * probe + {0: minus, 1: plus}] = sum_b sum_t -ln p[target]   (OV/lstm_eigen_class_batch/lstm.h:203-245)
cudaError_t launch_window_loss_f64(const float* params, const size_t off[5], const float* h0, const float* c0, const int* xs,
                                   const int* tg, int M, int N, int S, int B, const unsigned long long* probe_pos,
                                   int n_probes, double delta, double* out, cudaStream_t st);
void launch_eval_finish(const float* sum_part, const float* y_tgt, size_t steps, int G, double* bits_out, cudaStream_t st);

// ---- the reference's default shape as one persistent kernel (train_small.cu): `iters` whole training iterations per launch ----
struct TrainSmallArgs {
  float *W, *U, *b, *Why, *by;         // parameters            (updated in place)
  float *mW, *mU, *mb, *mWhy, *mby;    // Adagrad memory        (updated in place)
  float *gW, *gU, *gb, *gWhy, *gby;    // gradients             (written for the LAST iteration of the call)
  float *Hs, *Cs, *Gs, *dY, *dHy, *dG, *surp;   // activations  (slot 0 of Hs / Cs in and out; the rest written for the last iteration)
  int *xs, *tg;                        // the last iteration's window is written here (lstm_get_window)
  const uint8_t* text; unsigned long long len; const unsigned long long* pos0; unsigned long long* vcount;   // mode 0
  double* ring; unsigned long long cap; unsigned long long* iter;                                             // loss ring
  int S, T, iters, stride, mode, loss_mode, shift;   // mode 0: windows from the device text; 1: ONE iteration on the window in win_x / win_t
  int win_x[8], win_t[8];              // mode 1: the window itself travels as a kernel argument (S <= 5), no copy is enqueued
  double* host_loss;                   // optional: pinned, device-accessible host double that also receives the last iteration's loss
  float lr, clip; double eps;
  float m_exact;                       // set by the launcher: from this Adagrad memory value on, (float)(m + eps) == m
  long long* dbg;                      // optional [32]: SM-clock stamps of the last iteration
};
bool train_small_eligible(int M, int N, int S, int B);
cudaError_t launch_train_small(const TrainSmallArgs& a, int N, cudaStream_t st);

}  // namespace lstm
