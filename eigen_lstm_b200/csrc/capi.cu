// capi.cu — implementation of the C ABI declared in include/lstm_b200.h.
// Host-side orchestration only: every arithmetic statement of the reference path runs in a CUDA
// kernel (kernels_f32.cu for LSTM_F32, tc_*.cu for LSTM_BF16).  There is no CPU fallback.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <iomanip>
#include <random>
#include <sstream>

#include <dlfcn.h>

#include "ctx.h"
#include "kernels.h"

using namespace lstm;

// NCCL is bound at run time, only when data parallelism is requested: the library then shares the
// NCCL instance already loaded in the process (torch ships its own libnccl.so.2) instead of pulling
// a second copy in at load time.
namespace {
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
bool nccl_load() {
  if (g_nccl.ok) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return false;
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy && g_nccl.GetErrorString;
  return g_nccl.ok;
}
}  // namespace
#undef LSTM_NCCL
#define LSTM_NCCL(call)                                                                               \
  do {                                                                                                \
    ncclResult_t r_ = (call);                                                                         \
    if (r_ != ncclSuccess)                                                                            \
      return lstm_fail(ctx, LSTM_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_));    \
  } while (0)

static thread_local std::string g_create_error;

static void free_iter_graph_keep_key(lstm_ctx::IterGraph& g) {
  if (g.exec) cudaGraphExecDestroy(g.exec);
  for (cudaGraphExec_t e : g.segs) if (e) cudaGraphExecDestroy(e);
  g.exec = nullptr;
  g.segs.clear();
  g.seg_bucket.clear();
  g.seg_wait.clear();
  g.open_wait = 0;
}
static void free_iter_graph(lstm_ctx::IterGraph& g) {
  free_iter_graph_keep_key(g);
  g = lstm_ctx::IterGraph();
}

// captured iteration graphs bake in device pointers, sizes and the launch structure: anything that changes one of
// those (a new corpus, a resized loss ring, joining a communicator) must drop them
static void drop_graphs(lstm_ctx* ctx);

int lstm_fail(lstm_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

static inline void shape_of(const lstm_ctx* c, int which, int* rows, int* cols) {
  const int M = c->M, N = c->N;
  switch (which) {
    case LSTM_W: *rows = 4 * N; *cols = M; break;
    case LSTM_U: *rows = 4 * N; *cols = N; break;
    case LSTM_B: *rows = 4 * N; *cols = 1; break;
    case LSTM_WHY: *rows = M; *cols = N; break;
    default: *rows = M; *cols = 1; break;
  }
}

// R/lstm.cc:364-380 with an explicit seed: fresh mt19937, normal_distribution<double>, (i,j) order
static void randn_colmajor(float* m, int rows, int cols, double mean, double stddev, uint64_t seed) {
  std::mt19937 mt((uint32_t)seed);
  std::normal_distribution<> dist(mean, stddev);
  for (int i = 0; i < rows; i++)
    for (int j = 0; j < cols; j++) m[i + (size_t)rows * j] = (float)dist(mt);
}

extern "C" const char* lstm_version(void) { return "eigen_lstm_b200 0.1 (sm_100a)"; }

extern "C" const char* lstm_last_error(const lstm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int lstm_create(lstm_ctx** out, int M, int N, int S, int B, int device, int dtype) {
  lstm_ctx* ctx = nullptr;
  if (!out) return lstm_fail(nullptr, LSTM_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (M < 1 || M > 256 || N < 1 || S < 2 || B < 1)
    return lstm_fail(nullptr, LSTM_ERR_ARG, "need 1 <= M <= 256 (bytes), N >= 1, S >= 2, B >= 1");
  if (dtype != LSTM_F32 && dtype != LSTM_BF16) return lstm_fail(nullptr, LSTM_ERR_ARG, "dtype must be LSTM_F32 or LSTM_BF16");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return lstm_fail(nullptr, LSTM_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return lstm_fail(nullptr, LSTM_ERR_ARG, "device index out of range");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return lstm_fail(nullptr, LSTM_ERR_CUDA, cudaGetErrorString(e));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10)
    return lstm_fail(nullptr, LSTM_ERR_UNSUPPORTED, "this library is built for sm_100a (B200) only");

  ctx = new lstm_ctx();
  ctx->M = M; ctx->N = N; ctx->S = S; ctx->B = B; ctx->T = S - 1; ctx->device = device; ctx->dtype = dtype;
  const size_t N4 = 4 * (size_t)N;
  ctx->sz[LSTM_W] = N4 * M; ctx->sz[LSTM_U] = N4 * N; ctx->sz[LSTM_B] = N4;
  ctx->sz[LSTM_WHY] = (size_t)M * N; ctx->sz[LSTM_BY] = M;
  size_t o = 0;
  for (int i = 0; i < 5; i++) { ctx->off[i] = o; o += (ctx->sz[i] + 3) & ~(size_t)3; }  // 16-byte aligned tensors
  ctx->P = o;
#define CREATE_CUDA(call)                                                                    \
  do {                                                                                       \
    cudaError_t e2 = (call);                                                                 \
    if (e2 != cudaSuccess) {                                                                 \
      std::string msg = std::string(#call) + ": " + cudaGetErrorString(e2);                  \
      lstm_destroy(ctx);                                                                     \
      return lstm_fail(nullptr, LSTM_ERR_CUDA, msg);                                         \
    }                                                                                        \
  } while (0)
  CREATE_CUDA(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
  {
    // the communication stream outranks the compute stream: when an SM frees up while a weight-gradient panel still has tiles
    // queued, NCCL's few CTAs are placed first, so an allreduce never waits for a whole GEMM wave
    int prio_lo = 0, prio_hi = 0;
    CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CREATE_CUDA(cudaStreamCreateWithPriority(&ctx->comm_st, cudaStreamNonBlocking, prio_hi));
  }
  for (int i = 0; i < lstm_ctx::NBUCKET; i++) {
    CREATE_CUDA(cudaEventCreate(&ctx->ev_bucket[i]));   // timed: lstm_get_phase_ms reports when each bucket was handed over and summed
    CREATE_CUDA(cudaEventCreate(&ctx->ev_comm[i]));
  }
  for (int i = 0; i < 16; i++) CREATE_CUDA(cudaEventCreate(&ctx->pev[i]));
  const size_t T = ctx->T, BN = (size_t)B * N;
  CREATE_CUDA(cudaMalloc(&ctx->params, ctx->P * sizeof(float)));
  CREATE_CUDA(cudaMalloc(&ctx->grads, ctx->P * sizeof(float)));
  CREATE_CUDA(cudaMalloc(&ctx->mem, ctx->P * sizeof(float)));
  CREATE_CUDA(cudaMemsetAsync(ctx->params, 0, ctx->P * sizeof(float), ctx->st));
  CREATE_CUDA(cudaMemsetAsync(ctx->grads, 0, ctx->P * sizeof(float), ctx->st));
  CREATE_CUDA(cudaMemsetAsync(ctx->mem, 0, ctx->P * sizeof(float), ctx->st));
  const size_t h_slots = (dtype == LSTM_BF16) ? 1 : T + 1;  // the bf16 path keeps h in bf16 (tc_path.cu); fp32 slot 0 is the API copy
  CREATE_CUDA(cudaMalloc(&ctx->Hs, h_slots * BN * sizeof(float)));
  CREATE_CUDA(cudaMalloc(&ctx->Cs, (T + 1) * BN * sizeof(float)));
  CREATE_CUDA(cudaMemsetAsync(ctx->Hs, 0, h_slots * BN * sizeof(float), ctx->st));
  CREATE_CUDA(cudaMemsetAsync(ctx->Cs, 0, (T + 1) * BN * sizeof(float), ctx->st));
  if (dtype == LSTM_F32) {
    CREATE_CUDA(cudaMalloc(&ctx->Gs, T * B * N4 * sizeof(float)));
    CREATE_CUDA(cudaMalloc(&ctx->dG, T * B * N4 * sizeof(float)));
    CREATE_CUDA(cudaMalloc(&ctx->dY, T * B * (size_t)M * sizeof(float)));
    CREATE_CUDA(cudaMalloc(&ctx->dHy, T * BN * sizeof(float)));
    CREATE_CUDA(cudaMalloc(&ctx->dcnext, BN * sizeof(float)));
    CREATE_CUDA(cudaMalloc(&ctx->logit_shift, T * sizeof(float)));
  }
  CREATE_CUDA(cudaMalloc(&ctx->surp, T * B * sizeof(float)));
  CREATE_CUDA(cudaMalloc(&ctx->xs, (size_t)S * B * sizeof(int)));
  CREATE_CUDA(cudaMalloc(&ctx->tg, (size_t)S * B * sizeof(int)));
  CREATE_CUDA(cudaMemsetAsync(ctx->xs, 0xff, (size_t)S * B * sizeof(int), ctx->st));
  CREATE_CUDA(cudaMemsetAsync(ctx->tg, 0xff, (size_t)S * B * sizeof(int), ctx->st));
  ctx->loss_cap = 1024;
  CREATE_CUDA(cudaMalloc(&ctx->d_loss, ctx->loss_cap * sizeof(double)));
  CREATE_CUDA(cudaMallocHost(&ctx->h_loss_ring, lstm_ctx::LOSS_RING * sizeof(double)));
  CREATE_CUDA(cudaMalloc(&ctx->d_iter, sizeof(unsigned long long)));
  CREATE_CUDA(cudaMemsetAsync(ctx->d_iter, 0, sizeof(unsigned long long), ctx->st));
  for (int k = 0; k < 2; k++) {
    CREATE_CUDA(cudaMallocHost(&ctx->h_win[k], 2 * (size_t)S * B * sizeof(int32_t)));
    CREATE_CUDA(cudaEventCreateWithFlags(&ctx->ev_win[k], cudaEventDisableTiming));
  }
  CREATE_CUDA(cudaMalloc(&ctx->pos0, (size_t)B * sizeof(unsigned long long)));
  if (getenv("LSTM_TC_DEBUG") && dtype == LSTM_F32 && train_small_eligible(M, N, S, B)) {
    CREATE_CUDA(cudaMalloc(&ctx->small_dbg, 32 * sizeof(long long)));
    CREATE_CUDA(cudaMemsetAsync(ctx->small_dbg, 0, 32 * sizeof(long long), ctx->st));
  }
  CREATE_CUDA(cudaMalloc(&ctx->vcount, 2 * sizeof(unsigned long long)));
  CREATE_CUDA(cudaMemsetAsync(ctx->vcount, 0, 2 * sizeof(unsigned long long), ctx->st));
  ctx->h_pos0.assign(B, (uint64_t)S);
  CREATE_CUDA(cudaMemcpyAsync(ctx->pos0, ctx->h_pos0.data(), (size_t)B * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->st));
  if (dtype == LSTM_BF16) {
    int rc = tc_create(ctx);
    if (rc != 0) {
      std::string msg = ctx->err;
      lstm_destroy(ctx);
      return lstm_fail(nullptr, rc, msg);
    }
  }
  CREATE_CUDA(cudaStreamSynchronize(ctx->st));
#undef CREATE_CUDA
  *out = ctx;
  return LSTM_OK;
}

extern "C" int lstm_destroy(lstm_ctx* ctx) {
  if (!ctx) return LSTM_OK;
  cudaSetDevice(ctx->device);
  if (ctx->st) cudaStreamSynchronize(ctx->st);
  if (ctx->comm_st) cudaStreamSynchronize(ctx->comm_st);
  if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(ctx->comm);
  if (ctx->tc) tc_destroy(ctx);
  void* bufs[] = {ctx->params, ctx->grads, ctx->mem, ctx->Hs, ctx->Cs, ctx->Gs, ctx->dY, ctx->dHy, ctx->dG,
                  ctx->dcnext, ctx->surp, ctx->xs, ctx->tg, ctx->d_loss, ctx->text, ctx->pos0, ctx->vcount, ctx->logit_shift, ctx->small_dbg};
  for (void* b : bufs) if (b) cudaFree(b);
  if (ctx->h_loss_ring) cudaFreeHost(ctx->h_loss_ring);
  for (int k = 0; k < 2; k++) {
    if (ctx->h_win[k]) cudaFreeHost(ctx->h_win[k]);
    if (ctx->ev_win[k]) cudaEventDestroy(ctx->ev_win[k]);
  }
  if (ctx->d_iter) cudaFree(ctx->d_iter);
  for (auto& g : ctx->graph) free_iter_graph(g);
  for (int i = 0; i < lstm_ctx::NBUCKET; i++) {
    if (ctx->ev_bucket[i]) cudaEventDestroy(ctx->ev_bucket[i]);
    if (ctx->ev_comm[i]) cudaEventDestroy(ctx->ev_comm[i]);
  }
  for (int i = 0; i < 16; i++) if (ctx->pev[i]) cudaEventDestroy(ctx->pev[i]);
  if (ctx->st) cudaStreamDestroy(ctx->st);
  if (ctx->comm_st) cudaStreamDestroy(ctx->comm_st);
  delete ctx;
  return LSTM_OK;
}

extern "C" int lstm_sync(lstm_ctx* ctx) {
  if (!ctx) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->comm_st));
  for (auto& pl : ctx->pending_loss) *pl.second = ctx->h_loss_ring[pl.first];   // losses requested by lstm_train_step
  ctx->pending_loss.clear();
  return LSTM_OK;
}

extern "C" long lstm_tensor_size(const lstm_ctx* ctx, int which) {
  if (!ctx || which < 0 || which > 4) return -1;
  return (long)ctx->sz[which];
}
extern "C" long lstm_launch_count(const lstm_ctx* ctx) { return ctx ? ctx->launches : -1; }
extern "C" void* lstm_stream(lstm_ctx* ctx) { return ctx ? (void*)ctx->st : nullptr; }

extern "C" int lstm_set_option(lstm_ctx* ctx, int key, double value) {
  if (!ctx) return LSTM_ERR_ARG;
  switch (key) {
    case LSTM_OPT_CLIP:
      if (!(value >= 0.0)) return lstm_fail(ctx, LSTM_ERR_ARG, "clip must be >= 0 (0 = off)");
      ctx->clip = (float)value;
      break;
    case LSTM_OPT_LOSS_MODE:
      if (value != 0.0 && value != 1.0) return lstm_fail(ctx, LSTM_ERR_ARG, "loss mode: 0 = log2 over all timesteps, 1 = last timestep in nats");
      ctx->loss_mode = (int)value;
      break;
    case LSTM_OPT_SOFTMAX_SHIFT:
      if (value != 0.0 && value != 1.0) return lstm_fail(ctx, LSTM_ERR_ARG, "softmax shift: 0 = none, 1 = global maximum of the timestep");
      ctx->softmax_shift = (int)value;
      break;
    default:
      return lstm_fail(ctx, LSTM_ERR_ARG, "unknown option key");
  }
  LSTM_CUDA(cudaSetDevice(ctx->device));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  drop_graphs(ctx);   // options are baked into the captured iteration
  return LSTM_OK;
}

extern "C" int lstm_debug_variant(lstm_ctx* ctx, int out[8]) {
  if (!ctx || !out) return LSTM_ERR_ARG;
  for (int i = 0; i < 8; i++) out[i] = 0;
  if (ctx->tc) tc_variant(ctx, out);
  out[7] = (ctx->dtype == LSTM_F32 && ctx->world == 1 && train_small_eligible(ctx->M, ctx->N, ctx->S, ctx->B)) ? 1 : 0;
  return LSTM_OK;
}

static float* tensor_ptr(lstm_ctx* ctx, int kind, int which) {
  if (which < 0 || which > 4) return nullptr;
  switch (kind) {
    case LSTM_PARAM: return ctx->p(which);
    case LSTM_GRAD: return ctx->g(which);
    case LSTM_ADAGRAD_MEM: return ctx->m(which);
    default: return nullptr;
  }
}

extern "C" int lstm_set_tensor(lstm_ctx* ctx, int kind, int which, const float* src, size_t n) {
  if (!ctx || !src) return LSTM_ERR_ARG;
  float* d = tensor_ptr(ctx, kind, which);
  if (!d || n != ctx->sz[which]) return lstm_fail(ctx, LSTM_ERR_ARG, "lstm_set_tensor: bad kind/which/size");
  LSTM_CUDA(cudaSetDevice(ctx->device));
  LSTM_CUDA(cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyHostToDevice, ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  if (kind == LSTM_PARAM && ctx->tc) return tc_params_changed(ctx);
  return LSTM_OK;
}

extern "C" int lstm_get_tensor(lstm_ctx* ctx, int kind, int which, float* dst, size_t n) {
  if (!ctx || !dst) return LSTM_ERR_ARG;
  float* d = tensor_ptr(ctx, kind, which);
  if (!d || n != ctx->sz[which]) return lstm_fail(ctx, LSTM_ERR_ARG, "lstm_get_tensor: bad kind/which/size");
  LSTM_CUDA(cudaSetDevice(ctx->device));
  LSTM_CUDA(cudaStreamSynchronize(ctx->comm_st));
  LSTM_CUDA(cudaMemcpyAsync(dst, d, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  return LSTM_OK;
}

extern "C" int lstm_init_params(lstm_ctx* ctx, uint64_t seed, float std, float forget_bias) {
  if (!ctx) return LSTM_ERR_ARG;
  const int order[3] = {LSTM_W, LSTM_U, LSTM_WHY};  // R/lstm.cc:113-115
  for (int k = 0; k < 3; k++) {
    int rows, cols;
    shape_of(ctx, order[k], &rows, &cols);
    std::vector<float> h((size_t)rows * cols);
    randn_colmajor(h.data(), rows, cols, 0.0, std, seed + k);
    int rc = lstm_set_tensor(ctx, LSTM_PARAM, order[k], h.data(), h.size());
    if (rc) return rc;
  }
  std::vector<float> b(4 * (size_t)ctx->N, 0.f), by(ctx->M, 0.f);
  for (int j = 0; j < ctx->N; j++) b[2 * (size_t)ctx->N + j] = forget_bias;
  int rc = lstm_set_tensor(ctx, LSTM_PARAM, LSTM_B, b.data(), b.size());
  if (rc) return rc;
  rc = lstm_set_tensor(ctx, LSTM_PARAM, LSTM_BY, by.data(), by.size());
  if (rc) return rc;
  LSTM_CUDA(cudaMemsetAsync(ctx->mem, 0, ctx->P * sizeof(float), ctx->st));
  LSTM_CUDA(cudaMemsetAsync(ctx->grads, 0, ctx->P * sizeof(float), ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  return LSTM_OK;
}

extern "C" int lstm_set_state(lstm_ctx* ctx, const float* h0, const float* c0) {
  if (!ctx) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->B * ctx->N * sizeof(float);
  if (h0) LSTM_CUDA(cudaMemcpyAsync(ctx->Hslot(0), h0, bytes, cudaMemcpyHostToDevice, ctx->st));
  else LSTM_CUDA(cudaMemsetAsync(ctx->Hslot(0), 0, bytes, ctx->st));
  if (c0) LSTM_CUDA(cudaMemcpyAsync(ctx->Cslot(0), c0, bytes, cudaMemcpyHostToDevice, ctx->st));
  else LSTM_CUDA(cudaMemsetAsync(ctx->Cslot(0), 0, bytes, ctx->st));
  if (ctx->tc) {
    int rc = tc_state_from_f32(ctx);
    if (rc) return rc;
  }
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  return LSTM_OK;
}

extern "C" int lstm_get_state(lstm_ctx* ctx, float* h0, float* c0) {
  if (!ctx) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->B * ctx->N * sizeof(float);
  if (ctx->tc && h0) {
    int rc = tc_state_to_f32(ctx);
    if (rc) return rc;
  }
  if (h0) LSTM_CUDA(cudaMemcpyAsync(h0, ctx->Hslot(0), bytes, cudaMemcpyDeviceToHost, ctx->st));
  if (c0) LSTM_CUDA(cudaMemcpyAsync(c0, ctx->Cslot(0), bytes, cudaMemcpyDeviceToHost, ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  return LSTM_OK;
}

extern "C" int lstm_reset_state(lstm_ctx* ctx, uint64_t seed, float std) {
  if (!ctx) return LSTM_ERR_ARG;
  if (std == 0.f) return lstm_set_state(ctx, nullptr, nullptr);
  // R/lstm.cc:146-147 draws the whole N x S matrix; the column that becomes h(0) after the first
  // shift is column 1.  For B streams the batched snapshots draw N x B per timestep
  // (OV/lstm_eigen_opt/lstm.cc:176-181); we draw the carried-in column only: N x B, (i,j) order.
  std::vector<float> h((size_t)ctx->N * ctx->B), c((size_t)ctx->N * ctx->B);
  randn_colmajor(h.data(), ctx->N, ctx->B, 0.0, std, seed);
  randn_colmajor(c.data(), ctx->N, ctx->B, 0.0, std, seed + 1);
  return lstm_set_state(ctx, h.data(), c.data());
}

// ------------------------------------------------------------------------------------------------
// forward / backward / adagrad
// ------------------------------------------------------------------------------------------------
static void drop_graphs(lstm_ctx* ctx) {
  for (auto& g : ctx->graph) {
    free_iter_graph(g);
  }
}

static int ensure_loss_cap(lstm_ctx* ctx, size_t need) {
  if (need <= ctx->loss_cap) return LSTM_OK;
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  drop_graphs(ctx);   // the ring pointer and capacity are baked into captured graphs
  LSTM_CUDA(cudaFree(ctx->d_loss));
  ctx->d_loss = nullptr;
  LSTM_CUDA(cudaMalloc(&ctx->d_loss, need * sizeof(double)));
  ctx->loss_cap = need;
  return LSTM_OK;
}

#define PROF(i) do { if (ctx->profiling) LSTM_CUDA(cudaEventRecord(ctx->pev[i], ctx->st)); } while (0)

static int forward_device(lstm_ctx* ctx) {
  const int B = ctx->B, N = ctx->N, M = ctx->M, T = ctx->T;
  PROF(1);
  if (ctx->dtype == LSTM_BF16) {
    int rc = tc_forward(ctx);
    if (rc) return rc;
  } else {
    const size_t N4 = 4 * (size_t)N;
    for (int t = 1; t <= T; t++) {
      launch_step_fwd_f32(ctx->p(LSTM_U), ctx->p(LSTM_W), ctx->p(LSTM_B), ctx->Hslot(t - 1), ctx->Cslot(t - 1),
                          ctx->xs + (size_t)t * B, ctx->Gs + (size_t)(t - 1) * B * N4, ctx->Cslot(t), ctx->Hslot(t),
                          B, N, ctx->st);
    }
    LSTM_LAUNCHED(T);
    PROF(2);
    // K3: Y[(t,b)][m] = sum_n H[(t,b)][n] * Why(m,n) + by[m]            (R/lstm.cc:195)
    launch_gemm_f32(ctx->Hslot(1), N, 1, ctx->p(LSTM_WHY), M, 1, ctx->dY, M, 1, ctx->p(LSTM_BY), T * B, M, N, ctx->st);
    if (ctx->softmax_shift) {
      launch_logit_max_f32(ctx->dY, ctx->logit_shift, T, B * M, ctx->st);
      LSTM_LAUNCHED(1);
    }
    launch_softmax_ce_f32(ctx->dY, ctx->tg + B, ctx->surp, T * B, M, ctx->softmax_shift ? ctx->logit_shift : nullptr, B, ctx->st);
    LSTM_LAUNCHED(2);
  }
  launch_loss_reduce(ctx->surp, T, B, ctx->d_loss, ctx->loss_cap, ctx->d_iter, ctx->loss_mode, ctx->st);
  LSTM_LAUNCHED(1);
  PROF(3);
  ctx->fwd_done = true;
  return LSTM_OK;
}

// Segmented capture (data parallel): close the graph segment being captured on ctx->st, remember that `bucket` is
// summed after it, and open the next segment.
static int cut_segment(lstm_ctx* ctx, int bucket) {
  lstm_ctx::IterGraph* g = ctx->seg_capture;
  cudaGraph_t graph = nullptr;
  LSTM_CUDA(cudaStreamEndCapture(ctx->st, &graph));
  cudaGraphExec_t exec = nullptr;
  cudaError_t ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess) return lstm_fail(ctx, LSTM_ERR_CUDA, std::string("cudaGraphInstantiate (segment): ") + cudaGetErrorString(ce));
  g->segs.push_back(exec);
  g->seg_bucket.push_back(bucket);
  g->seg_wait.push_back(g->open_wait);
  g->open_wait = 0;
  g->open_l0 = ctx->enqueued;
  LSTM_CUDA(cudaStreamBeginCapture(ctx->st, cudaStreamCaptureModeThreadLocal));
  return LSTM_OK;
}

int lstm_allreduce_bucket(lstm_ctx* ctx, int bucket) {
  // bucket 0 = [W,U,b] (or its last panel), bucket 1 = [Why,by]; each contiguous in the flat gradient vector.
  if (ctx->world <= 1) return LSTM_OK;
  if (ctx->seg_capture) return cut_segment(ctx, bucket);   // capture pass: the graph ends here, NCCL runs between replays
  // K6a may be launched as up to three column panels of the column-major [W|U|b] matrix = contiguous ranges of the flat gradient
  // vector: panel 0 (bucket 2) and panel 1 (bucket 3) are summed while the panels after them are still being computed
  const size_t e0 = ctx->panel_end[0], e1 = ctx->panel_end[1] ? ctx->panel_end[1] : e0, end = ctx->off[LSTM_WHY];
  float* ptr; size_t cnt;
  switch (bucket) {
    case 1: ptr = ctx->g(LSTM_WHY); cnt = ctx->P - end; break;
    case 2: ptr = ctx->g(LSTM_W); cnt = e0; break;
    case 3: ptr = ctx->g(LSTM_W) + e0; cnt = e1 - e0; break;
    default: ptr = ctx->g(LSTM_W) + e1; cnt = end - e1; break;
  }
  if (cnt == 0) return LSTM_OK;
  LSTM_CUDA(cudaEventRecord(ctx->ev_bucket[bucket], ctx->st));
  LSTM_CUDA(cudaStreamWaitEvent(ctx->comm_st, ctx->ev_bucket[bucket], 0));
  LSTM_NCCL(g_nccl.AllReduce(ptr, ptr, cnt, ncclFloat, ncclSum, ctx->comm, ctx->comm_st));
  LSTM_CUDA(cudaEventRecord(ctx->ev_comm[bucket], ctx->comm_st));
  return LSTM_OK;
}

int lstm_wait_buckets(lstm_ctx* ctx, unsigned mask) {
  if (ctx->world <= 1 || mask == 0) return LSTM_OK;
  if (ctx->seg_capture) {   // capture pass: the waits are issued between two segment launches, before the segment that needs them
    lstm_ctx::IterGraph* g = ctx->seg_capture;
    if (ctx->enqueued != g->open_l0) {   // the open segment already holds work that must not wait: close it
      int rc = cut_segment(ctx, -1);
      if (rc) return rc;
    }
    g->open_wait |= mask;
    return LSTM_OK;
  }
  for (int b = 0; b < lstm_ctx::NBUCKET; b++)
    if (mask & (1u << b)) LSTM_CUDA(cudaStreamWaitEvent(ctx->st, ctx->ev_comm[b], 0));
  return LSTM_OK;
}

static int backward_device(lstm_ctx* ctx) {
  if (!ctx->fwd_done) return lstm_fail(ctx, LSTM_ERR_STATE, "lstm_backward before lstm_forward");
  const int B = ctx->B, N = ctx->N, M = ctx->M, T = ctx->T;
  if (ctx->dtype == LSTM_BF16) {
    int rc = tc_backward(ctx);  // records its own phase events and launches the allreduce buckets
    if (rc) return rc;
    return LSTM_OK;
  }
  const size_t N4 = 4 * (size_t)N;
  const int BT = T * B;
  // K4: dHy[(t,b)][n] = sum_m dY[(t,b)][m] * Why(m,n)                      (R/lstm.cc:228)
  launch_gemm_f32(ctx->dY, M, 1, ctx->p(LSTM_WHY), 1, M, ctx->dHy, N, 1, nullptr, BT, N, M, ctx->st);
  // K6c first (its inputs are complete), so its allreduce bucket overlaps the BPTT recurrence:
  // dWhy(m,n) = sum_(t,b) dY[(t,b)][m] * H_t[(t,b)][n]  (:226) ; dby = sum dY (:227)
  launch_gemm_f32(ctx->Hslot(1), 1, N, ctx->dY, M, 1, ctx->g(LSTM_WHY), M, 1, nullptr, N, M, BT, ctx->st);
  launch_colsum_f32(ctx->dY, ctx->g(LSTM_BY), BT, M, ctx->st);
  LSTM_LAUNCHED(3);
  PROF(4);
  int rc = lstm_allreduce_bucket(ctx, 1);
  if (rc) return rc;
  // K5: BPTT recurrence t = T..1
  for (int t = T; t >= 1; t--) {
    launch_step_bwd_f32(ctx->p(LSTM_U), t < T ? ctx->dG + (size_t)t * B * N4 : nullptr,
                        ctx->dHy + (size_t)(t - 1) * B * N, ctx->Gs + (size_t)(t - 1) * B * N4, ctx->Cslot(t),
                        ctx->Cslot(t - 1), ctx->dcnext, ctx->dG + (size_t)(t - 1) * B * N4, B, N, t == T, ctx->st);
  }
  LSTM_LAUNCHED(T);
  PROF(5);
  // K6a: dU(r,k) = sum_(t,b) dG[(t,b)][r] * H_{t-1}[(t,b)][k]              (:250)
  launch_gemm_f32(ctx->Hslot(0), 1, N, ctx->dG, (long)N4, 1, ctx->g(LSTM_U), (long)N4, 1, nullptr, N, (int)N4, BT, ctx->st);
  // K6b: dW(:,m) = sum of dG rows whose input byte is m                    (:251)
  launch_dw_scatter_f32(ctx->dG, ctx->xs + B, ctx->g(LSTM_W), BT, (int)N4, M, ctx->st);
  launch_colsum_f32(ctx->dG, ctx->g(LSTM_B), BT, (int)N4, ctx->st);      // (:252)
  LSTM_LAUNCHED(3);
  PROF(6);
  rc = lstm_allreduce_bucket(ctx, 0);
  if (rc) return rc;
  return LSTM_OK;
}

static int adagrad_device(lstm_ctx* ctx, float lr, double eps, float clip) {
  // Data parallel: every bucket must have been summed.  (Updating the ranges that are already summed UNDER the last allreduce
  // was measured on 2 GPUs and lost: next to the HBM-bound update the 25 MB allreduce takes 0.19 ms instead of 0.10.)
  int rc = lstm_wait_buckets(ctx, (1u << lstm_ctx::NBUCKET) - 1u);
  if (rc) return rc;
  PROF(7);
  launch_adagrad_f32(ctx->params, ctx->grads, ctx->mem, ctx->P, lr, eps, clip, ctx->st);
  LSTM_LAUNCHED(1);
  if (ctx->tc) {
    rc = tc_params_changed(ctx);
    if (rc) return rc;
  }
  PROF(8);
  return LSTM_OK;
}

static int finish_profile(lstm_ctx* ctx) {
  if (!ctx->profiling) return LSTM_OK;
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  // events: 0 start(window) 1 fwd-recur start 2 logits start 3 fwd end 4 dHy end 5 bwd-recur end
  //         6 wgrad end 7 allreduce-wait end 8 adagrad end
  const int pairs[9][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {4, 5}, {5, 6}, {6, 7}, {7, 8}, {0, 8}};
  for (int i = 0; i < 9; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->pev[pairs[i][0]], ctx->pev[pairs[i][1]]) != cudaSuccess) { ms = -1.f; cudaGetLastError(); }
    ctx->phase_ms[i] = ms;
  }
  // data parallel: phase_ms[9 + b] = when bucket b's allreduce had finished, phase_ms[13..15] = when buckets 2, 3, 0 were handed
  // to the communication stream — all in ms since the start of the iteration (0 = bucket not used)
  for (int i = 9; i < 16; i++) ctx->phase_ms[i] = 0.f;
  if (ctx->world > 1) {
    LSTM_CUDA(cudaStreamSynchronize(ctx->comm_st));
    const int handed[3] = {2, 3, 0};
    for (int b = 0; b < lstm_ctx::NBUCKET; b++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->pev[0], ctx->ev_comm[b]) == cudaSuccess && ms > 0.f) ctx->phase_ms[9 + b] = ms;
      else cudaGetLastError();
    }
    for (int i = 0; i < 3; i++) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ctx->pev[0], ctx->ev_bucket[handed[i]]) == cudaSuccess && ms > 0.f) ctx->phase_ms[13 + i] = ms;
      else cudaGetLastError();
    }
  }
  return LSTM_OK;
}

// loss of the most recent forward: device ring -> pinned ring now (stream-ordered), pinned ring -> *out at the next lstm_sync
static int queue_loss(lstm_ctx* ctx, double* out) {
  if ((int)ctx->pending_loss.size() == lstm_ctx::LOSS_RING) {
    int rc = lstm_sync(ctx);
    if (rc) return rc;
  }
  const int k = (int)ctx->pending_loss.size();
  const size_t slot = (size_t)((ctx->fwd_count - 1) % ctx->loss_cap);
  LSTM_CUDA(cudaMemcpyAsync(ctx->h_loss_ring + k, ctx->d_loss + slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  ctx->pending_loss.emplace_back(k, out);
  return LSTM_OK;
}
static int fetch_loss(lstm_ctx* ctx, double* out) {
  int rc = queue_loss(ctx, out);
  return rc ? rc : lstm_sync(ctx);
}

// Stage the caller's window in the pinned buffer that is not in flight and copy it to the device (asynchronously: the caller's
// arrays may be reused as soon as this returns).
static int upload_window(lstm_ctx* ctx, const int32_t* x_idx, const int32_t* t_idx) {
  const size_t n = (size_t)ctx->S * ctx->B, bytes = n * sizeof(int);
  if (!x_idx || !t_idx) return lstm_fail(ctx, LSTM_ERR_ARG, "x_idx / t_idx is NULL");
  for (size_t i = 0; i < n; i++)
    if (x_idx[i] < -1 || x_idx[i] >= ctx->M || t_idx[i] < -1 || t_idx[i] >= ctx->M)
      return lstm_fail(ctx, LSTM_ERR_ARG, "window index outside [-1, M)");
  const int k = ctx->win_k;
  ctx->win_k ^= 1;
  LSTM_CUDA(cudaEventSynchronize(ctx->ev_win[k]));   // the copy issued two steps ago has read this buffer (usually long done)
  memcpy(ctx->h_win[k], x_idx, bytes);
  memcpy(ctx->h_win[k] + n, t_idx, bytes);
  LSTM_CUDA(cudaMemcpyAsync(ctx->xs, ctx->h_win[k], bytes, cudaMemcpyHostToDevice, ctx->st));
  LSTM_CUDA(cudaMemcpyAsync(ctx->tg, ctx->h_win[k] + n, bytes, cudaMemcpyHostToDevice, ctx->st));
  LSTM_CUDA(cudaEventRecord(ctx->ev_win[k], ctx->st));
  return LSTM_OK;
}

// One full training iteration on the context's stream.  mode 0: window built on the device from the loaded text;
// mode 1: the window is already on its way (upload_window).
static int iteration_body(lstm_ctx* ctx, int mode, int stride, float lr) {
  int rc;
  if (mode == 0) {
    launch_window_advance(ctx->text, ctx->text_len, ctx->pos0, ctx->vcount, stride, ctx->S, ctx->B, ctx->xs, ctx->tg, ctx->st);
    LSTM_LAUNCHED(1);
  }
  rc = forward_device(ctx);
  if (rc) return rc;
  rc = backward_device(ctx);
  if (rc) return rc;
  rc = adagrad_device(ctx, lr, 1e-10, ctx->clip);
  if (rc) return rc;
  // slot 0 <- the state the next window starts from.  (Issued before the update — under the last allreduce on several GPUs — the
  // copies did not pay: the allreduce next to them took 0.25 instead of 0.13 ms on 2 GPUs.)
  return lstm_carry_state(ctx, stride);
}

// Launch-bound inner loop (hundreds of short kernels per iteration): the iteration is captured once into a CUDA graph
// and replayed.  Profiling (per-phase events) and LSTM_NO_GRAPH=1 use plain stream launches.
//
// Data parallel: NCCL is kept OUT of the graphs (capturing the allreduce on the side stream into the iteration graph
// hung at 4 and 8 ranks on B200 / NCCL 2.28).  The iteration is captured as segments cut at each gradient bucket
// (lstm_allreduce_bucket -> cut_segment); a replay launches segment, eager allreduce on the communication stream,
// next segment, ..., and joins the communication stream before the last segment (Adagrad).
static int run_iteration(lstm_ctx* ctx, int mode, int stride, float lr) {
  static const bool no_graph = getenv("LSTM_NO_GRAPH") != nullptr;   // diagnostics (per-kernel profiling): plain stream launches
  lstm_ctx::IterGraph& g = ctx->graph[mode];
  const bool dp = ctx->world > 1;
  if ((g.exec || !g.segs.empty()) && (g.stride != stride || g.lr != lr)) free_iter_graph(g);
  // the host mirrors of the device counters (fwd_count ~ *d_iter, iteration) move only once the iteration is enqueued
  struct Bump { lstm_ctx* c; bool ok = false; ~Bump() { if (ok) { c->fwd_count++; c->iteration++; } } } bump{ctx};
  if (ctx->profiling || no_graph) {
    PROF(0);
    int rc = iteration_body(ctx, mode, stride, lr);
    bump.ok = rc == LSTM_OK;
    return rc;
  }
  if (!g.exec && g.segs.empty()) {
    if (g.warm == 0 || g.stride != stride || g.lr != lr) {   // first time: plain launches (also sets kernel attributes)
      g.warm = 1; g.stride = stride; g.lr = lr;
      int rc = iteration_body(ctx, mode, stride, lr);
      bump.ok = rc == LSTM_OK;
      return rc;
    }
    const long l0 = ctx->launches;
    LSTM_CUDA(cudaStreamBeginCapture(ctx->st, cudaStreamCaptureModeThreadLocal));
    if (dp) { ctx->seg_capture = &g; g.open_wait = 0; g.open_l0 = ctx->enqueued; }
    int rc = iteration_body(ctx, mode, stride, lr);
    ctx->seg_capture = nullptr;
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamEndCapture(ctx->st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); free_iter_graph_keep_key(g); return rc; }
    if (ce != cudaSuccess) {
      free_iter_graph_keep_key(g);
      return lstm_fail(ctx, LSTM_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    }
    g.launches = ctx->launches - l0;
    ctx->launches = l0;
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) {
      free_iter_graph_keep_key(g);
      return lstm_fail(ctx, LSTM_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
    }
    if (dp) { g.segs.push_back(exec); g.seg_bucket.push_back(-1); g.seg_wait.push_back(g.open_wait); g.open_wait = 0; }
    else g.exec = exec;
  }
  if (!dp) {
    LSTM_CUDA(cudaGraphLaunch(g.exec, ctx->st));
  } else {
    for (size_t i = 0; i < g.segs.size(); i++) {
      for (int b = 0; b < lstm_ctx::NBUCKET; b++)             // Adagrad reads the summed gradients
        if (g.seg_wait[i] & (1u << b)) LSTM_CUDA(cudaStreamWaitEvent(ctx->st, ctx->ev_comm[b], 0));
      LSTM_CUDA(cudaGraphLaunch(g.segs[i], ctx->st));
      if (g.seg_bucket[i] >= 0) {
        int rc = lstm_allreduce_bucket(ctx, g.seg_bucket[i]);
        if (rc) return rc;
      }
    }
  }
  ctx->launches += g.launches;
  bump.ok = true;
  return LSTM_OK;
}

// The reference's own default shape (batch 1, N = 64) trains inside ONE persistent kernel (train_small.cu): `iters` whole
// iterations per launch, bit-identical to the launch-per-kernel path above.  Profiling (per-phase events) uses the general path.
static bool small_path(const lstm_ctx* ctx) {
  return ctx->dtype == LSTM_F32 && ctx->world == 1 && !ctx->profiling && train_small_eligible(ctx->M, ctx->N, ctx->S, ctx->B);
}

static int train_small_run(lstm_ctx* ctx, int iters, int stride, float lr, int mode, const int32_t* x_idx = nullptr,
                           const int32_t* t_idx = nullptr, double* host_loss = nullptr) {
  if (iters <= 0) return LSTM_OK;
  TrainSmallArgs a;
  a.host_loss = host_loss;
  for (int i = 0; i < 8; i++) a.win_x[i] = a.win_t[i] = -1;
  if (mode == 1) {                     // the window is a kernel argument: nothing is staged or copied
    if (!x_idx || !t_idx) return lstm_fail(ctx, LSTM_ERR_ARG, "x_idx / t_idx is NULL");
    for (int i = 0; i < ctx->S; i++) {
      if (x_idx[i] < -1 || x_idx[i] >= ctx->M || t_idx[i] < -1 || t_idx[i] >= ctx->M)
        return lstm_fail(ctx, LSTM_ERR_ARG, "window index outside [-1, M)");
      a.win_x[i] = x_idx[i]; a.win_t[i] = t_idx[i];
    }
  }
  a.W = ctx->p(LSTM_W); a.U = ctx->p(LSTM_U); a.b = ctx->p(LSTM_B); a.Why = ctx->p(LSTM_WHY); a.by = ctx->p(LSTM_BY);
  a.mW = ctx->m(LSTM_W); a.mU = ctx->m(LSTM_U); a.mb = ctx->m(LSTM_B); a.mWhy = ctx->m(LSTM_WHY); a.mby = ctx->m(LSTM_BY);
  a.gW = ctx->g(LSTM_W); a.gU = ctx->g(LSTM_U); a.gb = ctx->g(LSTM_B); a.gWhy = ctx->g(LSTM_WHY); a.gby = ctx->g(LSTM_BY);
  a.Hs = ctx->Hs; a.Cs = ctx->Cs; a.Gs = ctx->Gs; a.dY = ctx->dY; a.dHy = ctx->dHy; a.dG = ctx->dG; a.surp = ctx->surp;
  a.xs = ctx->xs; a.tg = ctx->tg;
  a.text = ctx->text; a.len = ctx->text_len; a.pos0 = ctx->pos0; a.vcount = ctx->vcount;
  a.ring = ctx->d_loss; a.cap = ctx->loss_cap; a.iter = ctx->d_iter;
  a.S = ctx->S; a.T = ctx->T; a.iters = iters; a.stride = stride; a.mode = mode;
  a.loss_mode = ctx->loss_mode; a.shift = ctx->softmax_shift;
  a.lr = lr; a.clip = ctx->clip; a.eps = 1e-10;
  a.dbg = ctx->small_dbg;
  LSTM_CUDA(launch_train_small(a, ctx->N, ctx->st));
  ctx->launches += 1;
  ctx->fwd_count += (uint64_t)iters;
  ctx->iteration += iters;
  ctx->fwd_done = true;
  return LSTM_OK;
}

extern "C" int lstm_forward(lstm_ctx* ctx, const int32_t* x_idx, const int32_t* t_idx, double* loss_out) {
  if (!ctx) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  PROF(0);
  int rc = upload_window(ctx, x_idx, t_idx);
  if (rc) return rc;
  rc = forward_device(ctx);
  if (rc) return rc;
  ctx->fwd_count++;
  if (loss_out) return fetch_loss(ctx, loss_out);
  return LSTM_OK;
}

extern "C" int lstm_backward(lstm_ctx* ctx) {
  if (!ctx) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  int rc = backward_device(ctx);
  if (rc) return rc;
  if (ctx->world > 1) {  // standalone backward: make the summed gradients visible to the caller
    for (int i = 0; i < lstm_ctx::NBUCKET; i++) LSTM_CUDA(cudaStreamWaitEvent(ctx->st, ctx->ev_comm[i], 0));
  }
  return LSTM_OK;
}

extern "C" int lstm_adagrad(lstm_ctx* ctx, float lr, double eps, float clip) {
  if (!ctx) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  ctx->iteration++;
  return adagrad_device(ctx, lr, eps, clip);
}

extern "C" int lstm_carry_state(lstm_ctx* ctx, int stride) {
  if (!ctx) return LSTM_ERR_ARG;
  if (stride <= 0) return LSTM_OK;
  if (stride > ctx->T) stride = ctx->T;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->B * ctx->N * sizeof(float);
  if (ctx->tc) {
    int rc = tc_carry(ctx, stride);
    if (rc) return rc;
  } else {
    LSTM_CUDA(cudaMemcpyAsync(ctx->Hslot(0), ctx->Hslot(stride), bytes, cudaMemcpyDeviceToDevice, ctx->st));
  }
  LSTM_CUDA(cudaMemcpyAsync(ctx->Cslot(0), ctx->Cslot(stride), bytes, cudaMemcpyDeviceToDevice, ctx->st));
  ctx->enqueued += 1;
  return LSTM_OK;
}

extern "C" int lstm_train_step(lstm_ctx* ctx, const int32_t* x_idx, const int32_t* t_idx, int stride, float lr,
                               double* loss_out) {
  if (!ctx) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  if (small_path(ctx)) {
    // one-kernel path: the window (S <= 5 bytes pairs) is a kernel argument and the loss lands in the pinned ring directly —
    // the step is ONE kernel launch, no copies
    double* slot = nullptr;
    if (loss_out) {
      if ((int)ctx->pending_loss.size() == lstm_ctx::LOSS_RING) { int rc0 = lstm_sync(ctx); if (rc0) return rc0; }
      const int k = (int)ctx->pending_loss.size();
      slot = ctx->h_loss_ring + k;
      ctx->pending_loss.emplace_back(k, loss_out);
    }
    int rc = train_small_run(ctx, 1, stride, lr, 1, x_idx, t_idx, slot);
    if (rc && loss_out) ctx->pending_loss.pop_back();
    return rc;
  }
  int rc = upload_window(ctx, x_idx, t_idx);
  if (rc) return rc;
  rc = run_iteration(ctx, 1, stride, lr);
  if (rc) return rc;
  rc = finish_profile(ctx);
  if (rc) return rc;
  if (loss_out) return queue_loss(ctx, loss_out);   // delivered at the next lstm_sync: the step itself never blocks
  return LSTM_OK;
}

// ------------------------------------------------------------------------------------------------
// device text pipeline
// ------------------------------------------------------------------------------------------------
extern "C" int lstm_load_text(lstm_ctx* ctx, const uint8_t* bytes, size_t n) {
  if (!ctx || !bytes) return LSTM_ERR_ARG;
  if (n < (size_t)ctx->S + 2) return lstm_fail(ctx, LSTM_ERR_ARG, "text shorter than the window");
  LSTM_CUDA(cudaSetDevice(ctx->device));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  drop_graphs(ctx);   // the text pointer and length are kernel arguments inside the captured iteration
  if (ctx->text) { LSTM_CUDA(cudaFree(ctx->text)); ctx->text = nullptr; }
  LSTM_CUDA(cudaMalloc(&ctx->text, n));
  LSTM_CUDA(cudaMemcpyAsync(ctx->text, bytes, n, cudaMemcpyHostToDevice, ctx->st));
  ctx->text_len = n;
  std::vector<uint64_t> pos(ctx->B, (uint64_t)ctx->S);
  return lstm_set_positions(ctx, pos.data());
}

extern "C" int lstm_set_positions(lstm_ctx* ctx, const uint64_t* pos) {
  if (!ctx || !pos) return LSTM_ERR_ARG;
  if (!ctx->text) return lstm_fail(ctx, LSTM_ERR_STATE, "lstm_set_positions before lstm_load_text");
  for (int b = 0; b < ctx->B; b++)
    if (pos[b] < (uint64_t)ctx->S || pos[b] >= ctx->text_len)
      return lstm_fail(ctx, LSTM_ERR_ARG, "position outside [S, length)");
  LSTM_CUDA(cudaSetDevice(ctx->device));
  ctx->h_pos0.assign(pos, pos + ctx->B);
  ctx->v_host = 0;
  LSTM_CUDA(cudaMemcpyAsync(ctx->pos0, ctx->h_pos0.data(), (size_t)ctx->B * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->st));
  LSTM_CUDA(cudaMemsetAsync(ctx->vcount, 0, 2 * sizeof(unsigned long long), ctx->st));
  LSTM_CUDA(cudaMemsetAsync(ctx->xs, 0xff, (size_t)ctx->S * ctx->B * sizeof(int), ctx->st));
  LSTM_CUDA(cudaMemsetAsync(ctx->tg, 0xff, (size_t)ctx->S * ctx->B * sizeof(int), ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  return LSTM_OK;
}

extern "C" int lstm_get_positions(lstm_ctx* ctx, uint64_t* pos) {
  if (!ctx || !pos) return LSTM_ERR_ARG;
  if (!ctx->text) return lstm_fail(ctx, LSTM_ERR_STATE, "no text loaded");
  const uint64_t span = ctx->text_len - ctx->S;
  for (int b = 0; b < ctx->B; b++) pos[b] = ctx->S + (ctx->h_pos0[b] - ctx->S + ctx->v_host) % span;
  return LSTM_OK;
}

extern "C" int lstm_get_window(lstm_ctx* ctx, int32_t* x_idx, int32_t* t_idx) {
  if (!ctx || !x_idx || !t_idx) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = (size_t)ctx->S * ctx->B * sizeof(int);
  LSTM_CUDA(cudaMemcpyAsync(x_idx, ctx->xs, bytes, cudaMemcpyDeviceToHost, ctx->st));
  LSTM_CUDA(cudaMemcpyAsync(t_idx, ctx->tg, bytes, cudaMemcpyDeviceToHost, ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  return LSTM_OK;
}

extern "C" int lstm_train_text(lstm_ctx* ctx, int iters, int stride, float lr, double* losses) {
  if (!ctx || iters < 0 || stride < 1) return LSTM_ERR_ARG;
  if (!ctx->text) return lstm_fail(ctx, LSTM_ERR_STATE, "lstm_train_text before lstm_load_text");
  LSTM_CUDA(cudaSetDevice(ctx->device));
  int rc = ensure_loss_cap(ctx, (size_t)iters);
  if (rc) return rc;
  const uint64_t first = ctx->fwd_count;
  if (small_path(ctx)) {
    rc = train_small_run(ctx, iters, stride, lr, 0);
    if (rc) return rc;
    ctx->v_host += (uint64_t)stride * (uint64_t)iters;
  } else {
    for (int it = 0; it < iters; it++) {
      rc = run_iteration(ctx, 0, stride, lr);
      if (rc) return rc;
      ctx->v_host += stride;   // host mirror of *vcount: moves only once the iteration is enqueued
    }
  }
  rc = finish_profile(ctx);
  if (rc) return rc;
  if (losses && iters > 0) {
    LSTM_CUDA(cudaStreamSynchronize(ctx->st));
    std::vector<double> ring(ctx->loss_cap);
    LSTM_CUDA(cudaMemcpy(ring.data(), ctx->d_loss, ctx->loss_cap * sizeof(double), cudaMemcpyDeviceToHost));
    for (int it = 0; it < iters; it++) losses[it] = ring[(size_t)((first + it) % ctx->loss_cap)];
  }
  return LSTM_OK;
}

// ------------------------------------------------------------------------------------------------
// evaluation / sampling
// ------------------------------------------------------------------------------------------------
// Batch-1 recurrences run on the persistent multi-CTA kernel (K9) when the model is big enough to need more than one
// SM; tiny models (the reference's N = 64 default) stay on the single-CTA kernel, which has no grid barriers.
static bool use_persistent(const lstm_ctx* ctx) { return ctx->N >= 128; }

struct ScopedFree {
  std::vector<void*> p;
  ~ScopedFree() { for (void* q : p) if (q) cudaFree(q); }
  template <typename T> cudaError_t alloc(T** out, size_t bytes) { cudaError_t e = cudaMalloc(out, bytes); if (e == cudaSuccess) p.push_back(*out); return e; }
};

extern "C" int lstm_eval_bpc(lstm_ctx* ctx, const uint8_t* bytes, size_t n, double* bpc_out) {
  if (!ctx || !bytes || !bpc_out || n < 2) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  ScopedFree mem;
  uint8_t* d_text = nullptr;
  double* d_bits = nullptr;
  LSTM_CUDA(mem.alloc(&d_text, n));
  LSTM_CUDA(mem.alloc(&d_bits, sizeof(double)));
  LSTM_CUDA(cudaMemcpyAsync(d_text, bytes, n, cudaMemcpyHostToDevice, ctx->st));
  LSTM_CUDA(cudaMemsetAsync(d_bits, 0, sizeof(double), ctx->st));
  if (!use_persistent(ctx)) {
    launch_recur_b1_f32(ctx->p(LSTM_W), ctx->p(LSTM_U), ctx->p(LSTM_B), ctx->p(LSTM_WHY), ctx->p(LSTM_BY), ctx->M, ctx->N,
                        0, d_text, n, nullptr, nullptr, nullptr, nullptr, d_bits, ctx->st);
    LSTM_LAUNCHED(1);
  } else {
    // chunks of CH characters: the per-step softmax partial sums are [CH][G] floats; state is chained on the device
    const size_t CH = 16384;
    int sms = 0;
    LSTM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    float *hbuf = nullptr, *ebuf = nullptr, *sum_part = nullptr, *y_tgt = nullptr, *h_in = nullptr, *c_in = nullptr, *c_out = nullptr;
    unsigned int* bar = nullptr;
    LSTM_CUDA(mem.alloc(&hbuf, 2 * (size_t)ctx->N * sizeof(float)));
    LSTM_CUDA(mem.alloc(&ebuf, 2 * (size_t)ctx->M * sizeof(float)));
    LSTM_CUDA(mem.alloc(&sum_part, CH * (size_t)sms * sizeof(float)));
    LSTM_CUDA(mem.alloc(&y_tgt, CH * sizeof(float)));
    LSTM_CUDA(mem.alloc(&h_in, (size_t)ctx->N * sizeof(float)));
    LSTM_CUDA(mem.alloc(&c_in, (size_t)ctx->N * sizeof(float)));
    LSTM_CUDA(mem.alloc(&c_out, (size_t)ctx->N * sizeof(float)));
    LSTM_CUDA(mem.alloc(&bar, sizeof(unsigned int)));
    LSTM_CUDA(cudaMemsetAsync(h_in, 0, (size_t)ctx->N * sizeof(float), ctx->st));
    LSTM_CUDA(cudaMemsetAsync(c_in, 0, (size_t)ctx->N * sizeof(float), ctx->st));
    for (size_t off = 0; off + 1 < n; off += CH) {
      const size_t len = std::min(CH + 1, n - off);   // len bytes -> len-1 scored pairs
      int G = 0;
      LSTM_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned int), ctx->st));
      LSTM_CUDA(launch_recur_persist(ctx->p(LSTM_W), ctx->p(LSTM_U), ctx->p(LSTM_B), ctx->p(LSTM_WHY), ctx->p(LSTM_BY), ctx->M,
                                     ctx->N, 0, d_text + off, len, nullptr, h_in, c_in, nullptr, hbuf, ebuf, sum_part, y_tgt,
                                     c_out, bar, sms, &G, ctx->st));
      launch_eval_finish(sum_part, y_tgt, len - 1, G, d_bits, ctx->st);
      LSTM_LAUNCHED(2);
      LSTM_CUDA(cudaMemcpyAsync(h_in, hbuf + ((len - 1) & 1) * (size_t)ctx->N, (size_t)ctx->N * sizeof(float), cudaMemcpyDeviceToDevice, ctx->st));
      LSTM_CUDA(cudaMemcpyAsync(c_in, c_out, (size_t)ctx->N * sizeof(float), cudaMemcpyDeviceToDevice, ctx->st));
    }
  }
  double bits = 0;
  LSTM_CUDA(cudaMemcpyAsync(&bits, d_bits, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  *bpc_out = bits / (double)(n - 1);
  return LSTM_OK;
}

extern "C" int lstm_sample(lstm_ctx* ctx, uint64_t seed, const float* h0, const float* c0, uint8_t* out, size_t n,
                           int greedy) {
  if (!ctx) return LSTM_ERR_ARG;
  if (n == 0) return LSTM_OK;
  if (!out) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  // R/lstm.cc:309-311,326: mt19937 + uniform_real_distribution<double>(0,1), narrowed to float
  std::vector<float> u(n);
  {
    std::mt19937 gen((uint32_t)seed);
    std::uniform_real_distribution<> dis(0, 1);
    for (size_t i = 0; i < n; i++) u[i] = (float)dis(gen);
  }
  ScopedFree mem;
  float *d_u = nullptr, *d_h = nullptr, *d_c = nullptr;
  uint8_t* d_out = nullptr;
  LSTM_CUDA(mem.alloc(&d_u, n * sizeof(float)));
  LSTM_CUDA(mem.alloc(&d_out, n));
  LSTM_CUDA(cudaMemcpyAsync(d_u, u.data(), n * sizeof(float), cudaMemcpyHostToDevice, ctx->st));
  if (h0) { LSTM_CUDA(mem.alloc(&d_h, ctx->N * sizeof(float))); LSTM_CUDA(cudaMemcpyAsync(d_h, h0, ctx->N * sizeof(float), cudaMemcpyHostToDevice, ctx->st)); }
  if (c0) { LSTM_CUDA(mem.alloc(&d_c, ctx->N * sizeof(float))); LSTM_CUDA(cudaMemcpyAsync(d_c, c0, ctx->N * sizeof(float), cudaMemcpyHostToDevice, ctx->st)); }
  if (!use_persistent(ctx)) {
    launch_recur_b1_f32(ctx->p(LSTM_W), ctx->p(LSTM_U), ctx->p(LSTM_B), ctx->p(LSTM_WHY), ctx->p(LSTM_BY), ctx->M, ctx->N,
                        greedy ? 2 : 1, nullptr, n, d_u, d_h, d_c, d_out, nullptr, ctx->st);
    LSTM_LAUNCHED(1);
  } else {
    int sms = 0, G = 0;
    LSTM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    float *hbuf = nullptr, *ebuf = nullptr;
    unsigned int* bar = nullptr;
    LSTM_CUDA(mem.alloc(&hbuf, 2 * (size_t)ctx->N * sizeof(float)));
    LSTM_CUDA(mem.alloc(&ebuf, 2 * (size_t)ctx->M * sizeof(float)));
    LSTM_CUDA(mem.alloc(&bar, sizeof(unsigned int)));
    LSTM_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned int), ctx->st));
    LSTM_CUDA(launch_recur_persist(ctx->p(LSTM_W), ctx->p(LSTM_U), ctx->p(LSTM_B), ctx->p(LSTM_WHY), ctx->p(LSTM_BY), ctx->M,
                                   ctx->N, greedy ? 2 : 1, nullptr, n, d_u, d_h, d_c, d_out, hbuf, ebuf, nullptr, nullptr, nullptr,
                                   bar, sms, &G, ctx->st));
    LSTM_LAUNCHED(1);
  }
  LSTM_CUDA(cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  return LSTM_OK;
}

// ------------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------------
extern "C" int lstm_get_activation(lstm_ctx* ctx, int what, int t, float* out, size_t n) {
  if (!ctx || !out) return LSTM_ERR_ARG;
  LSTM_CUDA(cudaSetDevice(ctx->device));
  const int B = ctx->B, N = ctx->N, M = ctx->M, T = ctx->T;
  if (t < 0 || t > T) return lstm_fail(ctx, LSTM_ERR_ARG, "t outside [0, S)");
  if (ctx->dtype == LSTM_BF16 && what != LSTM_ACT_C) return tc_get_activation(ctx, what, t, out, n);
  const float* src = nullptr;
  size_t cnt = 0;
  switch (what) {
    case LSTM_ACT_H: src = ctx->Hslot(t); cnt = (size_t)B * N; break;
    case LSTM_ACT_C: src = ctx->Cslot(t); cnt = (size_t)B * N; break;
    case LSTM_ACT_G: if (t < 1) return LSTM_ERR_ARG; src = ctx->Gs + (size_t)(t - 1) * B * 4 * N; cnt = (size_t)B * 4 * N; break;
    case LSTM_ACT_DG: if (t < 1) return LSTM_ERR_ARG; src = ctx->dG + (size_t)(t - 1) * B * 4 * N; cnt = (size_t)B * 4 * N; break;
    case LSTM_ACT_DHY: if (t < 1) return LSTM_ERR_ARG; src = ctx->dHy + (size_t)(t - 1) * B * N; cnt = (size_t)B * N; break;
    case LSTM_ACT_PROBS: if (t < 1) return LSTM_ERR_ARG; src = ctx->dY + (size_t)(t - 1) * B * M; cnt = (size_t)B * M; break;
    default: return lstm_fail(ctx, LSTM_ERR_ARG, "unknown activation selector");
  }
  if (n != cnt) return lstm_fail(ctx, LSTM_ERR_ARG, "lstm_get_activation: wrong size");
  LSTM_CUDA(cudaMemcpyAsync(out, src, cnt * sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  if (what == LSTM_ACT_PROBS) {  // stored as p - onehot(target): add the one-hot back
    std::vector<int> tgt(B);
    LSTM_CUDA(cudaMemcpy(tgt.data(), ctx->tg + (size_t)t * B, B * sizeof(int), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; b++)
      if (tgt[b] >= 0) out[(size_t)b * M + tgt[b]] += 1.0f;
  }
  return LSTM_OK;
}

// ------------------------------------------------------------------------------------------------
// gradient check
// ------------------------------------------------------------------------------------------------
extern "C" int lstm_gradcheck(lstm_ctx* ctx, const int32_t* x_idx, const int32_t* t_idx, int per_tensor, uint64_t seed,
                              double delta, double report[5][6], long long* probe_idx, double* numeric, double* analytic,
                              int* passed) {
  if (!ctx) return LSTM_ERR_ARG;
  if (per_tensor < 1 || !(delta > 0.0) || !report) return lstm_fail(ctx, LSTM_ERR_ARG, "lstm_gradcheck: need per_tensor >= 1, delta > 0, report");
  if (ctx->world > 1)
    return lstm_fail(ctx, LSTM_ERR_STATE, "lstm_gradcheck on a data-parallel context: backward sums gradients over ranks, the numeric side sees the local window only");
  LSTM_CUDA(cudaSetDevice(ctx->device));
  // An all-zero TARGET column (-1, the window warm-up of R/lstm.cc:124,169) contributes nothing to the loss (:204) but
  // dy = probs - 0 still enters the reference's gradients (:225): there the analytic gradient is, by the reference's own
  // construction, not the derivative of the loss, and a gradient check is meaningless.  (-1 INPUT columns are fine.)
  if (t_idx)
    for (size_t i = (size_t)ctx->B; i < (size_t)ctx->S * ctx->B; i++)
      if (t_idx[i] < 0) return lstm_fail(ctx, LSTM_ERR_ARG, "lstm_gradcheck: every target of timesteps 1..S-1 must be a real byte (not -1)");
  // analytic gradients: the context's own forward + backward (OV/lstm_eigen_class_batch/lstm.cc:291-293)
  int rc = lstm_forward(ctx, x_idx, t_idx, nullptr);
  if (rc) return rc;
  rc = lstm_backward(ctx);
  if (rc) return rc;
  if (ctx->tc) {                                   // bf16 contexts keep h(0) as bf16: hand the kernel the fp32 view of it
    rc = tc_state_to_f32(ctx);
    if (rc) return rc;
  }
  std::vector<float> grads(ctx->P);
  LSTM_CUDA(cudaMemcpyAsync(grads.data(), ctx->grads, ctx->P * sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
  // probes: per tensor, distinct uniformly drawn entries (the reference keeps each entry with probability 100 / size)
  std::vector<unsigned long long> pos;
  std::vector<long long> idx((size_t)5 * per_tensor, -1);
  for (int w = 0; w < 5; w++) {
    const size_t size = ctx->sz[w];
    const size_t want = std::min((size_t)per_tensor, size);
    std::mt19937_64 gen(seed + (uint64_t)w);
    std::vector<long long> chosen;
    if (want == size) {
      for (size_t i = 0; i < size; i++) chosen.push_back((long long)i);
    } else {
      while (chosen.size() < want) {
        const long long i = (long long)(gen() % size);
        if (std::find(chosen.begin(), chosen.end(), i) == chosen.end()) chosen.push_back(i);
      }
    }
    for (size_t q = 0; q < chosen.size(); q++) {
      idx[(size_t)w * per_tensor + q] = chosen[q];
      pos.push_back((unsigned long long)(ctx->off[w] + (size_t)chosen[q]));
    }
  }
  const int np = (int)pos.size();
  unsigned long long* d_pos = nullptr;
  double* d_out = nullptr;
  LSTM_CUDA(cudaMalloc(&d_pos, (size_t)np * sizeof(unsigned long long)));
  cudaError_t ce = cudaMalloc(&d_out, (size_t)np * 2 * sizeof(double));
  if (ce != cudaSuccess) { cudaFree(d_pos); return lstm_fail(ctx, LSTM_ERR_CUDA, cudaGetErrorString(ce)); }
  std::vector<double> loss((size_t)np * 2);
  ce = cudaMemcpyAsync(d_pos, pos.data(), (size_t)np * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->st);
  if (ce == cudaSuccess)
    ce = launch_window_loss_f64(ctx->params, ctx->off, ctx->Hslot(0), ctx->Cslot(0), ctx->xs, ctx->tg, ctx->M, ctx->N, ctx->S,
                                ctx->B, d_pos, np, delta, d_out, ctx->st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(loss.data(), d_out, loss.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->st);
  cudaFree(d_pos);
  cudaFree(d_out);
  if (ce != cudaSuccess) return lstm_fail(ctx, LSTM_ERR_CUDA, std::string("lstm_gradcheck: ") + cudaGetErrorString(ce));
  ctx->launches += 1;
  // check_gradient_error (OV/lstm_eigen_class_batch/lstm.cc:440-492).  The reference's analytic gradients are doubles and
  // it exempts only a + n == 0; ours are fp32 sums over B*T terms, which cannot resolve entries far below the tensor's
  // typical gradient: entries with |a + n| <= 1e-4 * max|n| of the tensor's probes count as error 0 as well.
  bool ok = true;
  size_t q0 = 0;
  for (int w = 0; w < 5; w++) {
    double mx = 0.0, sum = 0.0, nmin = INFINITY, nmax = -INFINITY, amin = INFINITY, amax = -INFINITY, nabs = 0.0;
    size_t cnt = 0;
    const size_t qbase = q0;
    for (int q = 0; q < per_tensor; q++) {
      if (idx[(size_t)w * per_tensor + q] < 0) continue;
      nabs = std::max(nabs, fabs((loss[2 * q0 + 1] - loss[2 * q0]) / (2.0 * delta)));
      q0++;
    }
    q0 = qbase;
    for (int q = 0; q < per_tensor; q++) {
      const long long i = idx[(size_t)w * per_tensor + q];
      if (i < 0) continue;
      const double n = (loss[2 * q0 + 1] - loss[2 * q0]) / (2.0 * delta);
      const double a = (double)grads[ctx->off[w] + (size_t)i];
      q0++;
      const double den = fabs(a + n);
      const double err = den > 1e-4 * nabs ? fabs(a - n) / den : 0.0;
      mx = std::max(mx, err); sum += err; cnt++;
      nmin = std::min(nmin, n); nmax = std::max(nmax, n); amin = std::min(amin, a); amax = std::max(amax, a);
      if (numeric) numeric[(size_t)w * per_tensor + q] = n;
      if (analytic) analytic[(size_t)w * per_tensor + q] = a;
    }
    const double mean = cnt ? sum / (double)cnt : 0.0;
    report[w][0] = mx; report[w][1] = mean; report[w][2] = nmin; report[w][3] = nmax; report[w][4] = amin; report[w][5] = amax;
    if (mx > 1e-1 || mean > 1e-3) ok = false;
  }
  if (probe_idx) memcpy(probe_idx, idx.data(), idx.size() * sizeof(long long));
  if (passed) *passed = ok ? 1 : 0;
  return LSTM_OK;
}

// ------------------------------------------------------------------------------------------------
// checkpoints
// ------------------------------------------------------------------------------------------------
static const char* kNames[5] = {"W", "U", "b", "Why", "by"};

// Eigen's default IOFormat (OV/lstm_eigen_class_CUDA/io.h:24 `file << m`): every coefficient printed
// with the stream's default precision (6 significant digits), right-aligned to the widest one,
// separated by one space, rows separated by '\n', no trailing newline.
static bool write_eigen_text(const std::string& path, const float* colmajor, int rows, int cols) {
  std::ofstream f(path.c_str());
  if (!f.is_open()) return false;
  size_t width = 0;
  std::vector<std::string> cell((size_t)rows * cols);
  for (int j = 0; j < cols; j++)
    for (int i = 0; i < rows; i++) {
      std::ostringstream s;
      s << colmajor[i + (size_t)rows * j];
      cell[i + (size_t)rows * j] = s.str();
      width = std::max(width, s.str().size());
    }
  for (int i = 0; i < rows; i++) {
    if (i) f << "\n";
    for (int j = 0; j < cols; j++) {
      if (j) f << " ";
      f << std::setw((int)width) << cell[i + (size_t)rows * j];
    }
  }
  return f.good();
}

// readMatrix (io.h:36-74): whitespace-separated numbers, one matrix row per line, into a pre-sized matrix
static bool read_eigen_text(const std::string& path, float* colmajor, int rows, int cols) {
  std::ifstream f(path.c_str());
  if (!f.is_open()) return false;
  std::string line;
  int r = 0;
  while (r < rows && std::getline(f, line)) {
    std::stringstream s(line);
    int c = 0;
    double v;
    while (c < cols && (s >> v)) colmajor[r + (size_t)rows * c++] = (float)v;
    if (c == 0) continue;  // blank line
    if (c != cols) return false;
    r++;
  }
  return r == rows;
}

extern "C" int lstm_save_text_ckpt(lstm_ctx* ctx, const char* prefix) {
  if (!ctx || !prefix) return LSTM_ERR_ARG;
  for (int w = 0; w < 5; w++) {
    int rows, cols;
    shape_of(ctx, w, &rows, &cols);
    std::vector<float> h((size_t)rows * cols);
    int rc = lstm_get_tensor(ctx, LSTM_PARAM, w, h.data(), h.size());
    if (rc) return rc;
    const std::string path = std::string(prefix) + "_" + kNames[w] + ".txt";
    if (!write_eigen_text(path, h.data(), rows, cols)) return lstm_fail(ctx, LSTM_ERR_IO, "file save error: (" + path + ")");
  }
  return LSTM_OK;
}

extern "C" int lstm_load_text_ckpt(lstm_ctx* ctx, const char* prefix) {
  if (!ctx || !prefix) return LSTM_ERR_ARG;
  std::vector<std::vector<float>> all(5);
  for (int w = 0; w < 5; w++) {  // read everything first: a failed load leaves the parameters untouched
    int rows, cols;
    shape_of(ctx, w, &rows, &cols);
    all[w].resize((size_t)rows * cols);
    const std::string path = std::string(prefix) + "_" + kNames[w] + ".txt";
    if (!read_eigen_text(path, all[w].data(), rows, cols)) return lstm_fail(ctx, LSTM_ERR_IO, "file read error: (" + path + ")");
  }
  for (int w = 0; w < 5; w++) {
    int rc = lstm_set_tensor(ctx, LSTM_PARAM, w, all[w].data(), all[w].size());
    if (rc) return rc;
  }
  return LSTM_OK;
}

struct BinHeader {
  char magic[8];
  int32_t M, N, S, B;
  int64_t iteration;
  uint64_t v;
};

extern "C" int lstm_save_bin(lstm_ctx* ctx, const char* path) {
  if (!ctx || !path) return LSTM_ERR_ARG;
  int rc = lstm_sync(ctx);
  if (rc) return rc;
  // everything is copied to the host first, so no error path leaves a file handle open
  std::vector<float> pbuf(ctx->P), mbuf(ctx->P);
  LSTM_CUDA(cudaMemcpy(pbuf.data(), ctx->params, ctx->P * sizeof(float), cudaMemcpyDeviceToHost));
  LSTM_CUDA(cudaMemcpy(mbuf.data(), ctx->mem, ctx->P * sizeof(float), cudaMemcpyDeviceToHost));
  const size_t bn = (size_t)ctx->B * ctx->N;
  std::vector<float> hb(bn), cb(bn);
  rc = lstm_get_state(ctx, hb.data(), cb.data());   // bf16 contexts carry h(0) in bf16: this converts it to the fp32 view
  if (rc) return rc;
  FILE* f = fopen(path, "wb");
  if (!f) return lstm_fail(ctx, LSTM_ERR_IO, std::string("cannot open ") + path);
  BinHeader h;
  memcpy(h.magic, "LSTMB200", 8);
  h.M = ctx->M; h.N = ctx->N; h.S = ctx->S; h.B = ctx->B; h.iteration = ctx->iteration; h.v = ctx->v_host;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  ok = ok && fwrite(pbuf.data(), sizeof(float), ctx->P, f) == ctx->P;
  ok = ok && fwrite(mbuf.data(), sizeof(float), ctx->P, f) == ctx->P;
  ok = ok && fwrite(hb.data(), sizeof(float), bn, f) == bn;
  ok = ok && fwrite(cb.data(), sizeof(float), bn, f) == bn;
  ok = ok && fwrite(ctx->h_pos0.data(), sizeof(uint64_t), ctx->B, f) == (size_t)ctx->B;
  ok = (fclose(f) == 0) && ok;
  return ok ? LSTM_OK : lstm_fail(ctx, LSTM_ERR_IO, std::string("short write to ") + path);
}

extern "C" int lstm_load_bin(lstm_ctx* ctx, const char* path) {
  if (!ctx || !path) return LSTM_ERR_ARG;
  FILE* f = fopen(path, "rb");
  if (!f) return lstm_fail(ctx, LSTM_ERR_IO, std::string("cannot open ") + path);
  BinHeader h;
  if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "LSTMB200", 8) != 0) { fclose(f); return lstm_fail(ctx, LSTM_ERR_IO, "not an LSTMB200 checkpoint"); }
  if (h.M != ctx->M || h.N != ctx->N || h.S != ctx->S || h.B != ctx->B) { fclose(f); return lstm_fail(ctx, LSTM_ERR_ARG, "checkpoint shape differs from the context"); }
  std::vector<float> pbuf(ctx->P), mbuf(ctx->P);
  const size_t bn = (size_t)ctx->B * ctx->N;
  std::vector<float> hb(bn), cb(bn);
  std::vector<uint64_t> pos(ctx->B);
  bool ok = fread(pbuf.data(), sizeof(float), ctx->P, f) == ctx->P && fread(mbuf.data(), sizeof(float), ctx->P, f) == ctx->P &&
            fread(hb.data(), sizeof(float), bn, f) == bn && fread(cb.data(), sizeof(float), bn, f) == bn &&
            fread(pos.data(), sizeof(uint64_t), ctx->B, f) == (size_t)ctx->B;
  fclose(f);
  if (!ok) return lstm_fail(ctx, LSTM_ERR_IO, "truncated checkpoint");
  LSTM_CUDA(cudaSetDevice(ctx->device));
  {  // in-flight iterations (lstm_train_text is asynchronous) and allreduces must not land on the restored tensors
    int src = lstm_sync(ctx);
    if (src) return src;
  }
  LSTM_CUDA(cudaMemcpy(ctx->params, pbuf.data(), ctx->P * sizeof(float), cudaMemcpyHostToDevice));
  LSTM_CUDA(cudaMemcpy(ctx->mem, mbuf.data(), ctx->P * sizeof(float), cudaMemcpyHostToDevice));
  int rc = lstm_set_state(ctx, hb.data(), cb.data());
  if (rc) return rc;
  ctx->iteration = h.iteration;
  if (ctx->text) {
    rc = lstm_set_positions(ctx, pos.data());
    if (rc) return rc;
    ctx->v_host = h.v;
    LSTM_CUDA(cudaMemcpy(ctx->vcount, &h.v, sizeof(uint64_t), cudaMemcpyHostToDevice));
    if (h.v > 0) {  // rebuild the window the saved run was looking at
      launch_window_advance(ctx->text, ctx->text_len, ctx->pos0, ctx->vcount, 0, ctx->S, ctx->B, ctx->xs, ctx->tg, ctx->st);
      LSTM_LAUNCHED(1);
    }
  } else {
    ctx->h_pos0 = pos;
    ctx->v_host = h.v;
  }
  if (ctx->tc) return tc_params_changed(ctx);
  return LSTM_OK;
}

// ------------------------------------------------------------------------------------------------
// data parallel
// ------------------------------------------------------------------------------------------------
extern "C" int lstm_dp_unique_id(uint8_t id[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  if (!nccl_load()) return lstm_fail(nullptr, LSTM_ERR_NCCL, "libnccl.so.2 not found");
  if (g_nccl.GetUniqueId(&u) != ncclSuccess) return LSTM_ERR_NCCL;
  memcpy(id, &u, 128);
  return LSTM_OK;
}

extern "C" int lstm_dp_init(lstm_ctx* ctx, int rank, int world, const uint8_t id[128]) {
  if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return LSTM_ERR_ARG;
  if (ctx->comm) return lstm_fail(ctx, LSTM_ERR_STATE, "communicator already initialised");
  LSTM_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId u;
  memcpy(&u, id, 128);
  if (!nccl_load()) return lstm_fail(ctx, LSTM_ERR_NCCL, "libnccl.so.2 not found");
  LSTM_NCCL(g_nccl.CommInitRank(&ctx->comm, world, u, rank));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  drop_graphs(ctx);   // a graph captured before joining has no gradient buckets; data-parallel iterations replay as segments
  ctx->rank = rank;
  ctx->world = world;
  return LSTM_OK;
}

// ------------------------------------------------------------------------------------------------
// measurement
// ------------------------------------------------------------------------------------------------
extern "C" int lstm_debug_kernel_clocks(lstm_ctx* ctx, long long out[32]) {
  if (!ctx || !out) return LSTM_ERR_ARG;
  if (ctx->small_dbg) {
    LSTM_CUDA(cudaStreamSynchronize(ctx->st));
    LSTM_CUDA(cudaMemcpy(out, ctx->small_dbg, 32 * sizeof(long long), cudaMemcpyDeviceToHost));
    return LSTM_OK;
  }
  return tc_debug_read(ctx, out);
}

extern "C" int lstm_set_profiling(lstm_ctx* ctx, int on) {
  if (!ctx) return LSTM_ERR_ARG;
  ctx->profiling = on != 0;
  return LSTM_OK;
}
extern "C" int lstm_get_phase_ms(lstm_ctx* ctx, float ms[16]) {
  if (!ctx || !ms) return LSTM_ERR_ARG;
  memcpy(ms, ctx->phase_ms, sizeof(float) * 16);
  return LSTM_OK;
}
