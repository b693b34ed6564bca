/* lstm_b200.h — C ABI of the B200-native character-LSTM training path.
 *
 * This is the drop-in boundary for the hot path of krocki/Eigen-LSTM (R/ = the reference root,
 * OV/ = R/optimized-obsfuscated_versions/).  The reference has no FFI of its own; the boundary is
 * cut where its last snapshots cut their class layer:
 *
 *   Parameters / cuParameters      OV/lstm_eigen_class_CUDA/lstm.h:43-112, cu_lstm.h:20-76
 *   LSTM<S>::forward/backward      OV/lstm_eigen_class_CUDA/lstm.h:137-366, cu_lstm.h:162-275
 *   adagrad()/cuda_adagrad()       OV/lstm_eigen_class_CUDA/lstm.cc:397-417, cu_lstm.h:417-432
 *   rawread()                      R/lstm.cc:382-420
 *   test(), sample()               OV/lstm_eigen_class_CUDA/lstm.cc:661-720, :578-659
 *   save_to_disk/load_from_disk    OV/lstm_eigen_class_CUDA/lstm.h:83-101, io.h:16-81
 *
 * Conventions
 *   - plain pointers and sizes only; every matrix crossing the boundary is COLUMN-MAJOR float32
 *     exactly like the reference's Eigen matrices (element (r,c) of rows x cols at [r + rows*c]).
 *   - gate order inside W, U, b is [i, o, f, u] in row blocks [0,N) [N,2N) [2N,3N) [3N,4N)
 *     (R/lstm.cc:77,179-192).
 *   - all state lives in the context; one CUDA stream (+ one NCCL communicator when data-parallel)
 *     per context.  A context is not thread-safe; independent contexts are.
 *   - every function returns 0 on success and a negative code on error; lstm_last_error() gives
 *     the text.  There is no CPU fallback: without a CUDA device lstm_create() fails.
 */
#ifndef LSTM_B200_H_
#define LSTM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lstm_ctx lstm_ctx;

/* tensor selectors (order of R/lstm.cc:65-70) */
enum { LSTM_W = 0, LSTM_U = 1, LSTM_B = 2, LSTM_WHY = 3, LSTM_BY = 4 };
/* tensor kinds */
enum { LSTM_PARAM = 0, LSTM_GRAD = 1, LSTM_ADAGRAD_MEM = 2 };
/* arithmetic of the contractions */
enum { LSTM_F32 = 0,      /* fp32 SIMT FFMA everywhere: the parity path (1e-4 rel. loss vs the Eigen path) */
       LSTM_BF16 = 1 };   /* bf16 operands on tcgen05 tensor cores, fp32 accumulate, fp32 master weights */
/* activation selectors for lstm_get_activation (differential tests, like compare_lstm_states,
 * OV/lstm_eigen_class_CUDA/cu_lstm.h:398-415) */
enum { LSTM_ACT_H = 0, LSTM_ACT_C = 1, LSTM_ACT_G = 2, LSTM_ACT_PROBS = 3, LSTM_ACT_DHY = 4, LSTM_ACT_DG = 5 };

/* option keys for lstm_set_option */
enum { LSTM_OPT_CLIP = 0,           /* gradient clipping fused into Adagrad for lstm_train_step / lstm_train_text: each gradient
                                     * entry is clamped to [-clip, clip] first; 0 = off = the reference (R/lstm.cc:259-272) */
       LSTM_OPT_LOSS_MODE = 1,      /* reported loss: 0 = sum_t (sum_b -log2 p_t[k_t]) / B (R/lstm.cc:204-207, OV/lstm_eigen_opt/
                                     * lstm.cc:246-249); 1 = the LAST timestep only, natural log: (sum_b -ln p_{S-1}[k]) / B
                                     * (OV/lstm_eigen_class_batch/lstm.cc:308-319).  Gradients flow from all timesteps either way. */
       LSTM_OPT_SOFTMAX_SHIFT = 2 };/* 0 = exp(y) unshifted (R/lstm.cc:199-201); 1 = the maximum of the timestep's whole M x B logit
                                     * matrix is subtracted first (OV/lstm_eigen_class_batch/lstm.h:175).  LSTM_F32 only: the bf16
                                     * path always subtracts each column's own maximum (the same probabilities up to rounding). */

/* error codes */
enum { LSTM_OK = 0, LSTM_ERR_ARG = -1, LSTM_ERR_CUDA = -2, LSTM_ERR_STATE = -3, LSTM_ERR_IO = -4,
       LSTM_ERR_NCCL = -5, LSTM_ERR_UNSUPPORTED = -6 };

/* ---- lifetime ------------------------------------------------------------------------------ */
/* M vocabulary (256 for raw bytes), N hidden size, S window length (S-1 active timesteps,
 * R/lstm.cc:53-57), B streams (batch, OV/lstm_eigen_opt/lstm.cc:116).  Replaces the
 * Parameters/LSTM<S> constructors.  dtype: LSTM_F32 | LSTM_BF16. */
int lstm_create(lstm_ctx** out, int M, int N, int S, int B, int device, int dtype);
int lstm_destroy(lstm_ctx* ctx);
const char* lstm_last_error(const lstm_ctx* ctx);   /* ctx may be NULL: error of the last failed lstm_create */
int lstm_sync(lstm_ctx* ctx);                       /* wait for the context's stream */
/* options of the training path (see LSTM_OPT_*); drops captured iteration graphs */
int lstm_set_option(lstm_ctx* ctx, int key, double value);
/* elements of tensor `which` (rows*cols), or <0 */
long lstm_tensor_size(const lstm_ctx* ctx, int which);

/* ---- parameters (Parameters{W,U,b,Why,by}) -------------------------------------------------- */
int lstm_set_tensor(lstm_ctx* ctx, int kind, int which, const float* colmajor, size_t n);
int lstm_get_tensor(lstm_ctx* ctx, int kind, int which, float* colmajor, size_t n);
/* R/lstm.cc:113-119: W,U,Why ~ N(0,std) with mt19937 + normal_distribution<double> in (row,col)
 * fill order, k-th tensor seeded seed+k; b,by = 0; forget-gate bias rows [2N,3N) = forget_bias
 * (OV/lstm_eigen_class_batch/lstm.cc:81).  Adagrad memory is zeroed. */
int lstm_init_params(lstm_ctx* ctx, uint64_t seed, float std, float forget_bias);

/* ---- carried state h(0), c(0): N x B column-major (R/lstm.cc:146-147,163-164) ---------------- */
int lstm_set_state(lstm_ctx* ctx, const float* h0, const float* c0);
int lstm_get_state(lstm_ctx* ctx, float* h0, float* c0);
/* h(0),c(0) ~ N(mean 0, std) like randn(h,0,0.1) at the top of each epoch; std = 0 zeroes them */
int lstm_reset_state(lstm_ctx* ctx, uint64_t seed, float std);

/* ---- one window: LSTM<S>::forward / backward / adagrad -------------------------------------- */
/* x_idx, t_idx: int32 [S][B] host arrays; row t holds the input byte x_t and target byte k_t of
 * every stream for timestep t (row 0 is ignored); -1 = the all-zero column the reference has while
 * its window warms up (R/lstm.cc:84,124,169-170).  loss_out (optional) receives the reference's
 * per-iteration `loss` = sum_t (sum_b -log2 p_t[k_t]) / B  (R/lstm.cc:204-207,
 * OV/lstm_eigen_opt/lstm.cc:246-249).  Passing loss_out forces a stream sync. */
int lstm_forward(lstm_ctx* ctx, const int32_t* x_idx, const int32_t* t_idx, double* loss_out);
int lstm_backward(lstm_ctx* ctx);   /* BPTT over the window of the last forward; gradients summed over t and b */
/* m += d*d; p -= lr * d / sqrtf(m + eps)  (R/lstm.cc:259-272; eps = 1e-10 is a double there, :25).  clip > 0 clamps each gradient
 * entry to [-clip, clip] first — an ADDITION of the north-star; 0 = off = the reference. */
int lstm_adagrad(lstm_ctx* ctx, float lr, double eps, float clip);
/* after a step: h(0),c(0) <- h(stride),c(stride) — the state the shifted window's first input needs (stride 1:
 * R/lstm.cc:163-164).  The reference's one stride > 1 program, OV/lstm_eigen_class_batch/lstm_segment.cc:183-184, copies
 * h[stride-1] instead, i.e. carries a state one character short of its window shift; a caller that wants exactly that
 * reads slot stride-1 with lstm_get_activation and writes it back with lstm_set_state
 * (tests/test_oracle_vs_reference_source.py replays that program this way). */
int lstm_carry_state(lstm_ctx* ctx, int stride);
/* forward + backward + (data-parallel gradient allreduce) + adagrad + carry(stride) in one call: the window
 * starts from the state in slot 0 (lstm_set_state / the previous step's carry) and leaves h(stride),c(stride) there.
 * ASYNCHRONOUS: x_idx / t_idx are copied into pinned staging before the call returns (the caller may reuse them at once);
 * *loss_out (optional; must stay valid) is written at the next lstm_sync() — or any other call that synchronises the
 * context — so that consecutive steps never wait for one another. */
int lstm_train_step(lstm_ctx* ctx, const int32_t* x_idx, const int32_t* t_idx, int stride, float lr,
                    double* loss_out);

/* ---- device-resident text pipeline (rawread + the window shift of R/lstm.cc:155-170) --------- */
/* copies the corpus to the device; positions reset to S (R/lstm.cc:151) */
int lstm_load_text(lstm_ctx* ctx, const uint8_t* bytes, size_t n);
/* per-stream text positions (OV/lstm_eigen_opt/lstm.cc:140-144); also clears the window */
int lstm_set_positions(lstm_ctx* ctx, const uint64_t* pos);
int lstm_get_positions(lstm_ctx* ctx, uint64_t* pos);
/* `iters` full training iterations on the loaded text, windows built on the device: each iteration
 * consumes `stride` new bytes per stream (1 = the reference's sliding window), runs fwd + BPTT + allreduce +
 * Adagrad and carries h(stride),c(stride) into slot 0 for the next iteration.  losses (optional, host, [iters]) gets every iteration's
 * loss.  Asynchronous unless losses != NULL. */
int lstm_train_text(lstm_ctx* ctx, int iters, int stride, float lr, double* losses);
/* copy out the current window indices ([S][B] each) — for tests */
int lstm_get_window(lstm_ctx* ctx, int32_t* x_idx, int32_t* t_idx);

/* ---- evaluation and sampling ----------------------------------------------------------------- */
/* test(): mean -log2 p over adjacent byte pairs, h = c = 0 at the start */
int lstm_eval_bpc(lstm_ctx* ctx, const uint8_t* bytes, size_t n, double* bpc_out);
/* sample(): draw n bytes; h0,c0 (N floats each, may be NULL = zeros); the k-th uniform comes from
 * mt19937(seed) + uniform_real_distribution<double> exactly as R/lstm.cc:309-311,326 (generated on
 * the host and streamed to the device); greedy != 0 takes the arg-max instead. */
int lstm_sample(lstm_ctx* ctx, uint64_t seed, const float* h0, const float* c0, uint8_t* out, size_t n, int greedy);

/* ---- introspection for differential tests ----------------------------------------------------- */
/* what: LSTM_ACT_*; t in [0,S): column-major (rows x B) like the reference's per-timestep matrices */
int lstm_get_activation(lstm_ctx* ctx, int what, int t, float* out, size_t n);

/* ---- gradient check (compute_all_numerical_grads + check_gradients, OV/lstm_eigen_class_batch/lstm.h:203-261,
 * lstm.cc:440-510; GPU snapshot: OV/lstm_eigen_class_CUDA/lstm.cc:420-500) ----------------------------------------- */
/* Runs forward + backward on the given window (same arguments as lstm_forward) from the current h(0), c(0), then, for
 * `per_tensor` randomly chosen entries of each of the five tensors (mt19937_64(seed + which), without replacement; all
 * entries if the tensor is smaller), evaluates the window's loss in DOUBLE precision on the device at p - delta and
 * p + delta and compares the central difference n with the analytic gradient a of the context's own arithmetic:
 * error = |a - n| / |a + n|, the reference's rule; it counts 0 where a + n == 0 and — the analytic side being fp32
 * sums here, not doubles — also where |a + n| <= 1e-4 * max|n| over the tensor's probes.
 *   report[which][0..5] = max rel. error, mean rel. error, min n, max n, min a, max a     (which = LSTM_W .. LSTM_BY)
 *   probe_idx / numeric / analytic (optional, each [5][per_tensor]): the sampled column-major indices (-1 = unused
 *   slot) and both gradient values, for callers that want their own statistics.
 * Returns LSTM_OK when the check RAN; the verdict (reference thresholds: max <= 1e-1 and mean <= 1e-3 for every tensor)
 * goes to *passed.  Targets of timesteps 1..S-1 must be real bytes: for an all-zero target column (-1) the reference's
 * dy = probs - target (R/lstm.cc:225) is not the derivative of its loss (:204), so LSTM_ERR_ARG is returned; -1 INPUT
 * columns are fine.  The loss differentiated is sum_b sum_t -ln p[target] (the gradients carry no 1/B, R/lstm.cc:225). */
int lstm_gradcheck(lstm_ctx* ctx, const int32_t* x_idx, const int32_t* t_idx, int per_tensor, uint64_t seed, double delta,
                   double report[5][6], long long* probe_idx, double* numeric, double* analytic, int* passed);

/* ---- checkpoints ------------------------------------------------------------------------------ */
/* Eigen-text format of Parameters::save_to_disk: <prefix>_{W,U,Why,b,by}.txt, one matrix row per
 * line, 6 significant digits (OV/lstm_eigen_class_CUDA/io.h:16-32) — interchangeable with the
 * reference's models/ files */
int lstm_save_text_ckpt(lstm_ctx* ctx, const char* prefix);
int lstm_load_text_ckpt(lstm_ctx* ctx, const char* prefix);
/* lossless: params + Adagrad memory + carried state + positions + iteration counter */
int lstm_save_bin(lstm_ctx* ctx, const char* path);
int lstm_load_bin(lstm_ctx* ctx, const char* path);

/* ---- data parallel (north-star item 5; the reference has none) -------------------- */
/* 128-byte NCCL unique id, created on rank 0 and distributed by the caller */
int lstm_dp_unique_id(uint8_t id[128]);
/* join a communicator: this context becomes rank `rank` of `world`; afterwards backward's five
 * gradient tensors are summed over ranks (no 1/world scaling: the reference sums over the batch,
 * R/lstm.cc:250-252) before Adagrad, so `world` contexts with B streams each equal one with world*B. */
int lstm_dp_init(lstm_ctx* ctx, int rank, int world, const uint8_t id[128]);

/* ---- measurement ------------------------------------------------------------------------------ */
/* CUDA-event time in ms of the phases of the LAST lstm_train_step / last iteration of
 * lstm_train_text when profiling is enabled: [0] window, [1] forward recurrence, [2] logits+softmax,
 * [3] dH_y (fp32 path) + dWhy|dby GEMM, [4] backward recurrence, [5] weight gradients, [6] allreduce wait, [7] adagrad, [8] total.
 * Profiling launches the kernels one after the other with plain stream launches: the kernels that normally run BESIDE the
 * persistent recurrences (logits, dWhy|dby, the late share of the weight-gradient GEMM) and the one-kernel training path of
 * the reference's default shape show up as separate phases (same results, bit for bit).
 * Data parallel: [9 + b] = ms since the start of the iteration at which gradient bucket b (0 = last column panel of [W|U|b],
 * 1 = [Why|by], 2 / 3 = leading panels) had been summed over the ranks, [13..15] = when buckets 2, 3, 0 were handed to the
 * communication stream (0 = bucket not used). */
int lstm_set_profiling(lstm_ctx* ctx, int on);
int lstm_get_phase_ms(lstm_ctx* ctx, float ms[16]);
/* diagnostics (LSTM_TC_DEBUG=1 in the environment at lstm_create, bf16 contexts): SM clock stamps taken inside the
 * last forward-step kernel ([0..8]) and the last BPTT-step kernel ([16..24]) by CTA 0: [0] entry, [4] prologue done,
 * [2] first operand stage landed, [1] last TMA issued, [3] last MMA issued, [5] accumulator complete, [6] tile
 * re-mapped through shared memory / cluster reduce done, [7] LSTM math + stores done, [8] exit */
int lstm_debug_kernel_clocks(lstm_ctx* ctx, long long out[32]);
/* which kernel instantiations this context's shape selects ([0..6]: bf16 contexts, zeros otherwise), so that parity tests can assert
 * that they cover the variants the benchmark runs: [0] K2 gate columns per CTA (BN), [1] K2 as cta_group::2 pairs,
 * [2] K5 hidden units per tile, [3] K5 variant (0 = split-K cluster of 4, 1 = pairs + 8-way split-K), [4] K6a tile width,
 * [5] K2 persistent, [6] K5 persistent (tile width, 0 = per-timestep launches); [7] fp32 contexts: 1 = lstm_train_text /
 * lstm_train_step run the whole iteration inside one persistent kernel (batch 1, N = 32 or 64: the reference's default shape) */
int lstm_debug_variant(lstm_ctx* ctx, int out[8]);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
long lstm_launch_count(const lstm_ctx* ctx);
/* raw CUDA stream (cudaStream_t) of the context, for callers that time with their own events */
void* lstm_stream(lstm_ctx* ctx);
const char* lstm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LSTM_B200_H_ */
