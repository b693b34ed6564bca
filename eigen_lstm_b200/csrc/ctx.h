// ctx.h — the context behind the C ABI (include/lstm_b200.h): owns all device memory, one compute
// stream, one communication stream and (optionally) one NCCL communicator.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/lstm_b200.h"

struct Bf16State;  // bf16 tensor-core path state (tc_path.cu)

struct lstm_ctx {
  int M = 0, N = 0, S = 0, B = 0, T = 0, device = 0, dtype = 0;
  cudaStream_t st = nullptr, comm_st = nullptr;
  static constexpr int NBUCKET = 4;   // 0 = [W|U|b] (or its last column panel), 1 = [Why|by], 2 / 3 = first / second column panel of [W|U|b]
  cudaEvent_t ev_bucket[NBUCKET] = {}, ev_comm[NBUCKET] = {};
  size_t panel_end[2] = {0, 0};       // end offsets (floats) of the panels summed as buckets 2 and 3 (0 = no such panel): set by the path that launches K6a in panels
  // flat parameter / gradient / Adagrad-memory vectors, tensor order W,U,b,Why,by (column-major each)
  size_t P = 0, off[5] = {0, 0, 0, 0, 0}, sz[5] = {0, 0, 0, 0, 0};
  float *params = nullptr, *grads = nullptr, *mem = nullptr;
  // activations (layouts in kernels.h)
  float *Hs = nullptr, *Cs = nullptr, *Gs = nullptr, *dY = nullptr, *dHy = nullptr, *dG = nullptr;
  float *dcnext = nullptr, *surp = nullptr;
  int *xs = nullptr, *tg = nullptr;
  // options (lstm_set_option): clip <= 0 = off (the reference); loss_mode 0 = log2 over all timesteps, 1 = last timestep in nats;
  // softmax_shift 0 = none (R/lstm.cc:199-201), 1 = the timestep's global maximum (OV/lstm_eigen_class_batch/lstm.h:175)
  float clip = 0.f;
  int loss_mode = 0, softmax_shift = 0;
  float* logit_shift = nullptr;   // [T] (fp32 path, softmax_shift = 1)
  double* d_loss = nullptr;  // [loss_cap] ring of per-iteration losses, slot = iteration % loss_cap
  size_t loss_cap = 0;
  unsigned long long* d_iter = nullptr;  // device-side forward counter (owned by k_loss_reduce)
  uint64_t fwd_count = 0;                // host mirror of *d_iter
  // one training iteration captured as a CUDA graph, keyed by (mode, stride, lr)
  // Data-parallel contexts replay the iteration as SEGMENTS: the graph is cut at every gradient bucket and the
  // ncclAllReduce is issued eagerly on the communication stream between two segment launches (NCCL stays outside
  // the graphs).  seg_bucket[i] = bucket to sum after segment i, -1 = none.
  struct IterGraph {
    cudaGraphExec_t exec = nullptr;
    std::vector<cudaGraphExec_t> segs;
    std::vector<int> seg_bucket;
    std::vector<unsigned> seg_wait;      // buckets (bit mask) whose sums segment i waits for before it is launched
    unsigned open_wait = 0;              // capture pass: the mask of the segment being captured
    long open_l0 = 0;                    //               and the launch count when it was opened
    int stride = 0; float lr = 0.f; float clip = 0.f; long launches = 0; int warm = 0;
  } graph[2];
  IterGraph* seg_capture = nullptr;      // non-null while run_iteration captures a segmented graph
  // lstm_train_step / lstm_forward stage the caller's windows through TWO pinned buffers ([x | t], 2*S*B ints each) so that a
  // step never waits for the previous step's DMA; ev_win[k] marks the last copy that read buffer k.
  int32_t* h_win[2] = {nullptr, nullptr};
  cudaEvent_t ev_win[2] = {nullptr, nullptr};
  int win_k = 0;
  // losses requested through loss_out travel device -> pinned ring -> caller's double at the next synchronisation
  static constexpr int LOSS_RING = 256;
  double* h_loss_ring = nullptr;
  std::vector<std::pair<int, double*>> pending_loss;
  // device text pipeline
  uint8_t* text = nullptr;
  size_t text_len = 0;
  unsigned long long *pos0 = nullptr, *vcount = nullptr;
  std::vector<uint64_t> h_pos0;
  uint64_t v_host = 0;
  // data parallel
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // bookkeeping
  bool fwd_done = false;
  long launches = 0;
  long enqueued = 0;                     // kernels + copies put on the compute stream (tells an empty graph segment from a used one)
  long iteration = 0;
  std::string err;
  // profiling
  bool profiling = false;
  cudaEvent_t pev[16] = {};
  float phase_ms[16] = {};
  Bf16State* tc = nullptr;
  long long* small_dbg = nullptr;   // LSTM_TC_DEBUG=1, fp32 one-kernel training: clock stamps (train_small.cu)

  float* p(int which) const { return params + off[which]; }
  float* g(int which) const { return grads + off[which]; }
  float* m(int which) const { return mem + off[which]; }
  float* Hslot(int t) const { return Hs + (size_t)t * B * N; }
  float* Cslot(int t) const { return Cs + (size_t)t * B * N; }
};

int lstm_fail(lstm_ctx* c, int code, const std::string& msg);

#define LSTM_CUDA(call)                                                                               \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return lstm_fail(ctx, LSTM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
  } while (0)
#define LSTM_NCCL(call)                                                                               \
  do {                                                                                                \
    ncclResult_t r_ = (call);                                                                         \
    if (r_ != ncclSuccess)                                                                            \
      return lstm_fail(ctx, LSTM_ERR_NCCL, std::string(#call) + ": " + ncclGetErrorString(r_));       \
  } while (0)
#define LSTM_LAUNCHED(n)                                                                              \
  do {                                                                                                \
    ctx->launches += (n);                                                                             \
    ctx->enqueued += (n);                                                                             \
    cudaError_t e_ = cudaGetLastError();                                                              \
    if (e_ != cudaSuccess)                                                                            \
      return lstm_fail(ctx, LSTM_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_));  \
  } while (0)

// bf16 tensor-core path (tc_path.cu); all return 0 or a negative LSTM_ERR_*
int tc_create(lstm_ctx* ctx);
void tc_destroy(lstm_ctx* ctx);
int tc_params_changed(lstm_ctx* ctx);  // refresh bf16 operand copies from the fp32 masters
int tc_forward(lstm_ctx* ctx);
int tc_backward(lstm_ctx* ctx);
int tc_get_activation(lstm_ctx* ctx, int what, int t, float* out, size_t n);
int tc_state_from_f32(lstm_ctx* ctx);  // Hs slot 0 (fp32) -> bf16 operand copies
int tc_state_to_f32(lstm_ctx* ctx);    // bf16 h(0) -> Hs slot 0
int tc_carry(lstm_ctx* ctx, int stride);
int tc_debug_read(lstm_ctx* ctx, long long out[32]);
void tc_variant(lstm_ctx* ctx, int out[8]);   // which kernel instantiations this context's shape selects
// sum gradient bucket (0 = [W,U,b] after the last panel end, 1 = [Why,by], 2 / 3 = the column panels ending at ctx->panel_end[0 / 1]) over
// the data-parallel ranks on the communication stream
int lstm_allreduce_bucket(lstm_ctx* ctx, int bucket);
// make the compute stream wait until the buckets in `mask` (bit b = bucket b) have been summed
int lstm_wait_buckets(lstm_ctx* ctx, unsigned mask);
