// tc_path.cu — host orchestration of the bf16 tensor-core path (LSTM_BF16): device buffers in the
// operand layouts of tc_kernels.cuh, TMA tensor maps, and the per-iteration kernel sequence.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ctx.h"
#include "kernels.h"
#include "tc_kernels.cuh"

using bf16 = __nv_bfloat16;

struct Bf16State {
  int Bp = 0, N4 = 0, RZ = 0, BN2 = 0, BN5 = 0;
  int bn2r = 0;                   // persistent forward recurrence (tc_recur.cu): gate columns per tile, 0 = per-timestep kernels
  bf16* Wb2 = nullptr;            // its blocked copy of U
  CUtensorMap tmWb2;
  int bnj5 = 0;                   // persistent BPTT recurrence (tc_recur.cu): hidden units per tile, 0 = per-timestep kernels
  bf16* Wb5 = nullptr;            // its blocked weight copy
  float* red5 = nullptr;
  float* kc_parts = nullptr;      // split-K partials of the dWhy|dby GEMM
  float* k6_parts = nullptr;      // [dW|dU|db] summed over the LAST timesteps of the window (computed beside the BPTT recurrence)
  int k6_late = 0;                // how many timesteps that is (0 = the weight-gradient GEMM is not split)
  CUtensorMap tmWb5;
  long LDZ = 0, LDT = 0;
  bf16 *Hbf = nullptr, *Urk = nullptr, *Ukr = nullptr, *Wmn = nullptr, *Wnm = nullptr;
  bf16 *dYbf = nullptr, *dYT = nullptr, *dGbf = nullptr, *dGT = nullptr, *ZT = nullptr;
  float *Wp = nullptr, *bp = nullptr, *Gp = nullptr, *dcnext = nullptr, *scratch = nullptr, *red = nullptr;
  unsigned int* gbar = nullptr;   // arrival counters of the persistent recurrences' grid barriers
  unsigned int* xcnt = nullptr;   // per-tile arrival counters of the persistent BPTT recurrence's split-K exchange
  size_t xcnt_bytes = 0;
  long long* dbg = nullptr;   // [32] kernel-internal clock stamps (LSTM_TC_DEBUG=1)
  // The persistent recurrences occupy 128 of the 148 SMs for 2.9 + 3.9 ms.  The small tensor-core kernels next to them are launched
  // as their PROGRAMMATIC DEPENDENTS (same stream; the recurrence signals griddepcontrol.launch_dependents once all its CTAs are
  // resident, the dependent never waits for its completion), so they run on the SMs the recurrence leaves free: K3 (logits,
  // softmax, dy) follows the forward recurrence timestep by timestep, K6c (dWhy | dby) runs beside the BPTT recurrence.
  // side_ctas = SMs the recurrences leave free (0 = no overlap).
  int side_ctas = 0;
  size_t scratch_elems = 0;
  CUtensorMap tmH, tmUrk, tmUkr, tmWmn, tmWnm, tmdY, tmdYT, tmdG, tmdGT, tmZT, tmZT256;
};

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

bool load_encode() {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return false;
  g_encode = (EncodeTiledFn)fn;
  return true;
}

// K-major bf16 matrix [rows][cols] (cols contiguous); box = 64 columns (128 B, SWIZZLE_128B) x box_rows
bool make_tmap(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {cols * sizeof(bf16)};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int pick_bn(int total_cols, int m_tiles, int min_ctas) {
  for (int bn : {128, 64, 32})
    if (total_cols % bn == 0 && (total_cols / bn) * m_tiles >= min_ctas) return bn;
  return 32;
}

}  // namespace

#define TC_ALLOC(ptr, bytes)                                          \
  do {                                                                \
    LSTM_CUDA(cudaMalloc(&(ptr), (bytes)));                           \
    LSTM_CUDA(cudaMemsetAsync((ptr), 0, (bytes), ctx->st));           \
  } while (0)

int tc_create(lstm_ctx* ctx) {
  const int M = ctx->M, N = ctx->N, B = ctx->B, T = ctx->T;
  if (M != 256) return lstm_fail(ctx, LSTM_ERR_UNSUPPORTED, "LSTM_BF16 needs M == 256 (raw bytes)");
  if (N % 64 != 0) return lstm_fail(ctx, LSTM_ERR_UNSUPPORTED, "LSTM_BF16 needs N to be a multiple of 64 (one 128-byte swizzled K block)");
  if (!load_encode()) return lstm_fail(ctx, LSTM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  Bf16State* s = new Bf16State();
  ctx->tc = s;
  s->Bp = (B + 127) / 128 * 128;
  s->N4 = 4 * N;
  s->RZ = M + N + 16;
  s->LDZ = (long)(T + 1) * s->Bp;
  s->LDT = (long)T * s->Bp;
  s->BN2 = pick_bn(s->N4, s->Bp / 128, 96);
  s->BN5 = pick_bn(N, (s->Bp / 128) * 4, 96);   // K5 runs 4 split-K CTAs per output tile
  const size_t Bp = s->Bp, N4 = s->N4;
  TC_ALLOC(s->Hbf, (size_t)(T + 1) * Bp * N * sizeof(bf16));
  TC_ALLOC(s->Urk, N4 * N * sizeof(bf16));
  TC_ALLOC(s->Ukr, N4 * N * sizeof(bf16));
  TC_ALLOC(s->Wmn, (size_t)M * N * sizeof(bf16));
  TC_ALLOC(s->Wnm, (size_t)M * N * sizeof(bf16));
  TC_ALLOC(s->dYbf, (size_t)T * Bp * M * sizeof(bf16));
  TC_ALLOC(s->dYT, (size_t)M * s->LDT * sizeof(bf16));
  TC_ALLOC(s->dGbf, (size_t)T * Bp * N4 * sizeof(bf16));
  TC_ALLOC(s->dGT, N4 * (size_t)s->LDT * sizeof(bf16));
  TC_ALLOC(s->ZT, (size_t)s->RZ * s->LDZ * sizeof(bf16));
  TC_ALLOC(s->Wp, (size_t)M * N4 * sizeof(float));
  TC_ALLOC(s->bp, N4 * sizeof(float));
  TC_ALLOC(s->Gp, (size_t)T * B * N4 * sizeof(float));
  TC_ALLOC(s->dcnext, (size_t)B * N * sizeof(float));
  TC_ALLOC(s->red, (size_t)(N / s->BN5) * (Bp / 128) * 16 * 128 * (s->BN5 / 4) * sizeof(float));   // K5 split-K exchange
  s->scratch_elems = (size_t)B * N4;
  TC_ALLOC(s->scratch, s->scratch_elems * sizeof(float));
  TC_ALLOC(s->gbar, std::max<size_t>((size_t)(Bp / 128) * 8 * 32, 2 * 256) * sizeof(unsigned int));   // persistent kernels: 2 x 256 flags
  s->xcnt_bytes = std::max<size_t>((size_t)(N / s->BN5) * (Bp / 128), (size_t)(N / 32) * 2) * sizeof(unsigned int);
  TC_ALLOC(s->xcnt, s->xcnt_bytes);
  if (getenv("LSTM_TC_DEBUG")) TC_ALLOC(s->dbg, 32 * sizeof(long long));
  TC_ALLOC(s->kc_parts, 8 * (ctx->P - ctx->off[LSTM_WHY]) * sizeof(float));
  // Persistent recurrences: the shape policy proposes a tile width, the occupancy query decides (every CTA of the launch must be
  // co-resident on THIS device).  Decided once, here: the operand copies that are kept fresh depend on it (tc_params_changed),
  // and a launch that fails later is an error, never a silent switch to kernels whose operands were not refreshed.
  s->bn2r = tc::fwd_recur_bn(N, s->Bp, M);
  s->bnj5 = tc::bwd_recur_bnj(N, s->Bp, M);
  {
    tc::FwdRecurArgs fa = {}; fa.B = B; fa.Bp = s->Bp; fa.N = N; fa.M = M; fa.T = T;
    tc::BwdRecurArgs ba = {}; ba.B = B; ba.Bp = s->Bp; ba.N = N; ba.M = M; ba.T = T;
    CUtensorMap none = {};
    if (s->bn2r && !tc::launch_fwd_recur(s->bn2r, none, none, fa, ctx->st, true)) s->bn2r = 0;
    if (s->bnj5 && !tc::launch_bwd_recur(s->bnj5, none, none, none, ba, ctx->st, true)) s->bnj5 = 0;
  }
  if (s->bn2r || s->bnj5) {
    // the SMs the persistent recurrences leave free run K3 / K6c beside them (tc_forward, tc_backward)
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device) != cudaSuccess) sms = 0;
    int busy = 0;
    if (s->bn2r) busy = tc::fwd_recur_ctas(s->bn2r, N, s->Bp);
    if (s->bnj5 && tc::bwd_recur_ctas(s->bnj5, N, s->Bp) > busy) busy = tc::bwd_recur_ctas(s->bnj5, N, s->Bp);
    s->side_ctas = sms - busy >= 8 ? sms - busy : 0;
    // The weight-gradient GEMM is split along time: BPTT runs from the last timestep to the first, so the last 3/16 of the window
    // are final early; their share of the GEMM (2.7 ms on the 20 free SMs at config 4) fits beside the remaining 13/16 of the
    // recurrence, and the GEMM that follows the recurrence is 3/16 shorter.
    // (2/16 ... 5/16 measured, profiles/r02ag_*: 3/16 is the best; the GPU runs at its power cap, so work moved beside the
    // recurrence lowers the clock of everything and only a part of the 3/16 shows up as time saved)
    if (s->bnj5 && s->side_ctas) s->k6_late = (T * 3) / 16;
    if (s->k6_late + 1 > T) s->k6_late = 0;
    if (s->k6_late) TC_ALLOC(s->k6_parts, ctx->off[LSTM_WHY] * sizeof(float));
  }
  if (s->bn2r) TC_ALLOC(s->Wb2, N4 * N * sizeof(bf16));
  if (s->bnj5) {
    TC_ALLOC(s->Wb5, (size_t)N * (N4 + M) * sizeof(bf16));
    TC_ALLOC(s->red5, tc::bwd_recur_red_floats(N, s->bnj5) * sizeof(float));
    if (s->xcnt_bytes < (size_t)(N / s->bnj5) * 2 * sizeof(unsigned int)) return lstm_fail(ctx, LSTM_ERR_STATE, "xcnt too small");
  }
  tc::launch_fill_bf16(s->ZT + (size_t)(M + N) * s->LDZ, 1.0f, (size_t)s->LDZ, ctx->st);  // the ones row (db, dby)
  LSTM_LAUNCHED(1);
  bool ok = true;
  ok &= make_tmap(&s->tmH, s->Hbf, (uint64_t)(T + 1) * Bp, N, 128);
  ok &= make_tmap(&s->tmUrk, s->Urk, N4, N, tc::step_pair(s->Bp) ? s->BN2 / 2 : s->BN2);   // pairs stage half of the U tile each
  ok &= make_tmap(&s->tmUkr, s->Ukr, N, N4, s->BN5);
  ok &= make_tmap(&s->tmWmn, s->Wmn, M, N, 256);
  ok &= make_tmap(&s->tmWnm, s->Wnm, N, M, s->BN5);
  ok &= make_tmap(&s->tmdY, s->dYbf, (uint64_t)T * Bp, M, 128);
  ok &= make_tmap(&s->tmdYT, s->dYT, M, s->LDT, 128);
  ok &= make_tmap(&s->tmdG, s->dGbf, (uint64_t)T * Bp, N4, 128);
  ok &= make_tmap(&s->tmdGT, s->dGT, N4, s->LDT, 128);
  ok &= make_tmap(&s->tmZT, s->ZT, s->RZ, s->LDZ, 128);
  ok &= make_tmap(&s->tmZT256, s->ZT, s->RZ, s->LDZ, 256);
  if (s->bn2r) ok &= make_tmap(&s->tmWb2, s->Wb2, (uint64_t)N4 * (N / 64), 64, (uint32_t)tc::fwd_recur_box_rows(s->bn2r, s->Bp));
  if (s->bnj5) ok &= make_tmap(&s->tmWb5, s->Wb5, (uint64_t)N * ((N4 + M) / 64), 64, (uint32_t)tc::bwd_recur_box_rows(s->bnj5, s->Bp));
  if (!ok) return lstm_fail(ctx, LSTM_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return LSTM_OK;
}

void tc_destroy(lstm_ctx* ctx) {
  Bf16State* s = ctx->tc;
  if (!s) return;
  void* bufs[] = {s->Hbf, s->Urk, s->Ukr, s->Wmn, s->Wnm, s->dYbf, s->dYT, s->dGbf, s->dGT, s->ZT, s->Wp, s->bp, s->Gp,
                  s->dcnext, s->scratch, s->red, s->dbg, s->gbar, s->xcnt, s->Wb5, s->red5, s->Wb2, s->kc_parts, s->k6_parts};
  for (void* b : bufs) if (b) cudaFree(b);
  delete s;
  ctx->tc = nullptr;
}

// fp32 masters (column-major like the reference) -> bf16 operand copies
int tc_params_changed(lstm_ctx* ctx) {
  Bf16State* s = ctx->tc;
  const int M = ctx->M, N = ctx->N;
  // Every bf16 copy of U from one pass over the fp32 master.  The row-major copies feed the per-timestep kernels only: a context
  // whose shape runs a recurrence persistently (tc_recur.cu) reads the blocked copy instead and skips the other.
  tc::launch_refresh_u(ctx->p(LSTM_U), s->bn2r ? nullptr : s->Urk, s->bnj5 ? nullptr : s->Ukr, s->Wb2, s->bn2r, s->Wb5, s->bnj5, M, N,
                       ctx->st);
  tc::launch_permute_rows_f32(ctx->p(LSTM_W), s->Wp, M, N, ctx->st);            // Wp[m][r']
  tc::launch_permute_rows_f32(ctx->p(LSTM_B), s->bp, 1, N, ctx->st);
  tc::launch_transpose_cast(ctx->p(LSTM_WHY), s->Wmn, N, M, 0, ctx->st);       // Wmn[m][n]
  LSTM_LAUNCHED(4);
  if (s->bnj5) tc::launch_block_why(ctx->p(LSTM_WHY), s->Wb5, N, M, s->bnj5, ctx->st);
  else tc::launch_cast_bf16(ctx->p(LSTM_WHY), s->Wnm, (size_t)M * N, ctx->st);  // Wnm[n][m]
  LSTM_LAUNCHED(1);
  return LSTM_OK;
}

int tc_state_from_f32(lstm_ctx* ctx) {
  Bf16State* s = ctx->tc;
  tc::launch_state_to_bf16(ctx->Hslot(0), s->Hbf, s->ZT + (size_t)ctx->M * s->LDZ, s->LDZ, ctx->B, ctx->N, ctx->st);
  LSTM_LAUNCHED(1);
  return LSTM_OK;
}

int tc_state_to_f32(lstm_ctx* ctx) {
  Bf16State* s = ctx->tc;
  tc::launch_state_to_f32(s->Hbf, ctx->Hslot(0), ctx->B, ctx->N, ctx->st);
  LSTM_LAUNCHED(1);
  return LSTM_OK;
}

int tc_carry(lstm_ctx* ctx, int stride) {
  Bf16State* s = ctx->tc;
  const size_t Bp = s->Bp;
  LSTM_CUDA(cudaMemcpyAsync(s->Hbf, s->Hbf + (size_t)stride * Bp * ctx->N, Bp * ctx->N * sizeof(bf16), cudaMemcpyDeviceToDevice, ctx->st));
  bf16* zh = s->ZT + (size_t)ctx->M * s->LDZ;
  LSTM_CUDA(cudaMemcpy2DAsync(zh, s->LDZ * sizeof(bf16), zh + (size_t)stride * Bp, s->LDZ * sizeof(bf16), Bp * sizeof(bf16), ctx->N,
                              cudaMemcpyDeviceToDevice, ctx->st));
  return LSTM_OK;
}

#define PROF(i) do { if (ctx->profiling) LSTM_CUDA(cudaEventRecord(ctx->pev[i], ctx->st)); } while (0)

int tc_forward(lstm_ctx* ctx) {
  Bf16State* s = ctx->tc;
  const int M = ctx->M, N = ctx->N, B = ctx->B, T = ctx->T;
  const size_t Bp = s->Bp, N4 = s->N4;
  bool persistent = false;
  // per-phase profiling keeps the kernels one after the other on the compute stream
  const bool overlap = s->bn2r && s->side_ctas > 0 && !ctx->profiling;
  // the one-hot rows of ZT feed the weight-gradient GEMM only: next to a persistent forward recurrence they are built beside it
  if (!overlap) {
    tc::launch_build_xt(ctx->xs + B, s->ZT, s->LDZ, M, T, B, s->Bp, ctx->st);
    LSTM_LAUNCHED(1);
  }
  if (s->bn2r) {                                         // the whole forward recurrence in one persistent launch (tc_recur.cu)
    tc::FwdRecurArgs pa;
    pa.B = B; pa.Bp = s->Bp; pa.N = N; pa.M = M; pa.T = T;
    pa.xs = ctx->xs; pa.Wp = s->Wp; pa.bp = s->bp; pa.Cs = ctx->Cs; pa.Gp = s->Gp; pa.Hbf = s->Hbf;
    pa.ZT_h0 = s->ZT + (size_t)M * s->LDZ; pa.ldz = s->LDZ; pa.gbar = s->gbar; pa.dbg = s->dbg;
    persistent = tc::launch_fwd_recur(s->bn2r, s->tmH, s->tmWb2, pa, ctx->st);
    if (!persistent) return lstm_fail(ctx, LSTM_ERR_CUDA, "the persistent forward recurrence could not be launched");
    LSTM_LAUNCHED(1);
  }
  for (int t = 1; t <= T && !persistent; t++) {
    tc::FwdStepArgs a;
    a.B = B; a.Bp = s->Bp; a.N = N; a.M = M;
    a.a_row0 = (t - 1) * s->Bp;
    a.x = ctx->xs + (size_t)t * B;
    a.Wp = s->Wp; a.bp = s->bp;
    a.c_prev = ctx->Cslot(t - 1); a.c_out = ctx->Cslot(t);
    a.Gp_t = s->Gp + (size_t)(t - 1) * B * N4;
    a.Hbf_t = s->Hbf + (size_t)t * Bp * N;
    a.ZT_h = s->ZT + (size_t)M * s->LDZ + (size_t)t * Bp;
    a.ldz = s->LDZ;
    a.dbg = s->dbg;
    tc::launch_fwd_step(s->BN2, s->tmH, s->tmUrk, a, ctx->st);
  }
  if (!persistent) LSTM_LAUNCHED(T);
  PROF(2);
  tc::LogitsArgs la;
  la.B = B; la.Bp = s->Bp; la.N = N; la.M = M; la.T = T;
  la.by = ctx->p(LSTM_BY); la.tg = ctx->tg + B;
  la.dYbf = s->dYbf; la.dYT = s->dYT; la.surp = ctx->surp;
  if (overlap) {
    // K3 beside the running recurrence: a programmatic dependent on the SMs the recurrence leaves free; it contracts the tile of
    // timestep t as soon as the recurrence's arrival counters say that h(t) is complete.  The next launch in the stream (the
    // loss reduction) waits for both kernels.
    la.progress = s->gbar;
    la.per_slot = (unsigned int)(4 * N / s->bn2r / tc::R_SLOTS);
    tc::launch_build_xt(ctx->xs + B, s->ZT, s->LDZ, M, T, B, s->Bp, ctx->st, true);
    LSTM_LAUNCHED(1);
    tc::launch_logits(s->tmH, s->tmWmn, la, ctx->st, s->side_ctas, true);
  } else {
    tc::launch_logits(s->tmH, s->tmWmn, la, ctx->st);
  }
  LSTM_LAUNCHED(1);
  return LSTM_OK;
}

// the whole BPTT recurrence in one persistent launch (tc_recur.cu)
static int launch_bptt_persistent(lstm_ctx* ctx) {
  Bf16State* s = ctx->tc;
  tc::BwdRecurArgs pa;
  pa.B = ctx->B; pa.Bp = s->Bp; pa.N = ctx->N; pa.M = ctx->M; pa.T = ctx->T;
  pa.Gp = s->Gp; pa.Cs = ctx->Cs; pa.dGbf = s->dGbf; pa.dGT = s->dGT; pa.ldg = s->LDT; pa.red = s->red5;
  pa.xcnt = s->xcnt; pa.gbar = s->gbar; pa.dbg = s->dbg ? s->dbg + 16 : nullptr;
  if (const char* e = getenv("LSTM_TC_DEBUG_STEP")) pa.dbg_s = atoi(e);
  if (!tc::launch_bwd_recur(s->bnj5, s->tmdG, s->tmWb5, s->tmdY, pa, ctx->st))
    return lstm_fail(ctx, LSTM_ERR_CUDA, "the persistent BPTT recurrence could not be launched");
  LSTM_LAUNCHED(1);
  return LSTM_OK;
}

// [dW | dU | db] over the LAST k6_late timesteps of the window -> k6_parts.  beside = true: as a programmatic dependent right behind
// K6c, next to the running BPTT recurrence; it starts contracting once the recurrence's arrival counters show that those timesteps
// (and one more: a timestep's dG^T columns are stored after its arrival) are done.
static int launch_k6a_late(lstm_ctx* ctx, bool beside) {
  Bf16State* s = ctx->tc;
  const int M = ctx->M, N = ctx->N, T = ctx->T;
  const int N4 = 4 * N;
  tc::GemmArgs g;
  const int bn = (M + N + 1 >= 1024) ? 256 : 128;
  g.rows = N4; g.cols = M + N + 1; g.nkb = s->k6_late * s->Bp / 64;
  g.a_k0 = g.b_k0 = (T - s->k6_late) * s->Bp; g.b_row0 = 0;
  g.C = s->k6_parts; g.ldc = (long)N4; g.tiles_m = N4 / 128; g.tiles_n = (M + N + 1 + bn - 1) / bn;
  g.splits = 1; g.split_stride = 0;
  if (beside) {
    g.wait_slots = s->gbar;
    g.wait_n = (s->Bp / 128) * tc::R_SLOTS;
    g.wait_target = (unsigned int)(s->k6_late + 1) * (unsigned int)tc::bwd_recur_per_slot(s->bnj5, N, s->Bp);
  }
  const bool pair = bn == 256 && g.tiles_m % 2 == 0;   // cta_group::2 pairs on 256 x 256 tiles
  tc::launch_gemm_nt(bn, s->tmdGT, pair ? s->tmZT : (bn == 256 ? s->tmZT256 : s->tmZT), g, ctx->st, beside, pair);
  LSTM_LAUNCHED(1);
  return LSTM_OK;
}

int tc_backward(lstm_ctx* ctx) {
  Bf16State* s = ctx->tc;
  const int M = ctx->M, N = ctx->N, B = ctx->B, T = ctx->T;
  const size_t Bp = s->Bp, N4 = s->N4;
  // K6c: [dWhy | dby](m, n) = sum_(s,b) dY^T[m][(s,b)] * [H^T ; 1][n][(s+1,b)] — inputs complete after K3.  With a persistent
  // BPTT recurrence it is launched right AFTER it as its programmatic dependent: it runs beside the recurrence on the SMs that
  // leaves free (it does not read anything the recurrence writes).
  const bool overlap = s->bnj5 && s->side_ctas > 0 && !ctx->profiling;
  if (overlap) {
    int rc0 = launch_bptt_persistent(ctx);
    if (rc0) return rc0;
  }
  {
    tc::GemmArgs g;
    g.rows = M; g.cols = N + 1; g.nkb = (int)(s->LDT / 64); g.a_k0 = 0; g.b_k0 = s->Bp; g.b_row0 = M;
    g.C = ctx->g(LSTM_WHY); g.ldc = M; g.tiles_m = M / 128; g.tiles_n = (N + 1 + 127) / 128;
    g.splits = 1; g.split_stride = 0;
    // few output tiles and a long K: split K over enough CTAs to fill the SMs, partials summed in a fixed order
    const int tiles = g.tiles_m * g.tiles_n;
    int splits = 1;
    while (splits < 8 && tiles * splits * 2 <= 148 && g.nkb % (splits * 2) == 0) splits *= 2;
    const size_t out_n = ctx->P - ctx->off[LSTM_WHY];        // [Why | by] incl. alignment padding: a multiple of 4 floats
    if (splits > 1 && s->kc_parts) {
      g.nkb /= splits; g.splits = splits; g.split_stride = out_n; g.C = s->kc_parts;
      tc::launch_gemm_nt(128, s->tmdYT, s->tmZT, g, ctx->st, overlap);
      LSTM_LAUNCHED(1);
      if (overlap && s->k6_late) { int rc1 = launch_k6a_late(ctx, true); if (rc1) return rc1; }
      tc::launch_sum_splits(s->kc_parts, ctx->g(LSTM_WHY), out_n, out_n, splits, ctx->st);   // plain launch: after all of the above
      LSTM_LAUNCHED(1);
    } else {
      tc::launch_gemm_nt(128, s->tmdYT, s->tmZT, g, ctx->st, overlap);
      LSTM_LAUNCHED(1);
      if (overlap && s->k6_late) { int rc1 = launch_k6a_late(ctx, true); if (rc1) return rc1; }
    }
  }
  PROF(4);
  // The [Why|by] bucket (2 MB) is summed under the BPTT recurrence when that is a chain of per-timestep launches; a PERSISTENT
  // recurrence needs all of its CTAs co-resident, so NCCL's CTAs must not sit on its SMs: the bucket then goes out right after
  // the recurrence and hides under the weight-gradient GEMM instead.
  int rc = s->bnj5 ? 0 : lstm_allreduce_bucket(ctx, 1);
  if (rc) return rc;
  bool persistent = overlap;
  if (s->bnj5 && !overlap) {                             // the whole BPTT recurrence in one persistent launch (tc_recur.cu)
    rc = launch_bptt_persistent(ctx);
    if (rc) return rc;
    persistent = true;
  }
  for (int t = T; t >= 1 && !persistent; t--) {
    tc::BwdStepArgs a;
    a.B = B; a.Bp = s->Bp; a.N = N; a.M = M;
    a.first = (t == T);
    a.dg_row0 = t * s->Bp;
    a.dy_row0 = (t - 1) * s->Bp;
    a.Gp_t = s->Gp + (size_t)(t - 1) * B * N4;
    a.c_t = ctx->Cslot(t); a.c_prev = ctx->Cslot(t - 1);
    a.dcnext = s->dcnext;
    a.dGbf_t = s->dGbf + (size_t)(t - 1) * Bp * N4;
    a.dGT_t = s->dGT + (size_t)(t - 1) * Bp;
    a.ldg = s->LDT;
    a.red = s->red;
    a.dbg = s->dbg ? s->dbg + 16 : nullptr;
    tc::launch_bwd_step(s->BN5, s->tmdG, s->tmUkr, s->tmdY, s->tmWnm, a, ctx->st);
  }
  if (!persistent) LSTM_LAUNCHED(T);
  PROF(5);
  if (s->bnj5) {
    rc = lstm_allreduce_bucket(ctx, 1);
    if (rc) return rc;
  }
  // K6a+b: [dW | dU | db](r, col) = sum_(s,b) dG^T[r][(s,b)] * [X^T ; H^T ; 1][col][(s,b)]  — the flat gradient
  // vector's first three tensors are exactly this column-major 4N x (M+N+1) matrix.
  {
    tc::GemmArgs g;
    g.rows = (int)N4; g.cols = M + N + 1; g.nkb = (int)(s->LDT / 64); g.a_k0 = 0; g.b_k0 = 0; g.b_row0 = 0;
    if (s->k6_late) {                                  // the last timesteps' share is already in k6_parts (or is computed now)
      if (!overlap) { int rc1 = launch_k6a_late(ctx, false); if (rc1) return rc1; }
      g.nkb = (T - s->k6_late) * (int)s->Bp / 64;
      g.addend = s->k6_parts;
    }
    const int bn = (M + N + 1 >= 1024) ? 256 : 128;   // wide tiles once there are enough of them to fill the SMs
    g.C = ctx->g(LSTM_W); g.ldc = (long)N4; g.tiles_m = (int)N4 / 128; g.tiles_n = (M + N + 1 + bn - 1) / bn;
    g.splits = 1; g.split_stride = 0;
    // Data parallel: two column panels (6 + 4 tile columns of 256 at N = 2048: 2.6 + 1.7 waves = the same five tile times as one
    // launch of 4.3 waves).  Each panel is a contiguous range of the flat gradient vector; the leading panel's allreduce (50 MB)
    // runs under the trailing panel's GEMM, the trailing 25 MB under the Adagrad update of everything else (adagrad_device).
    // Measured on 2 GPUs (profiles/r02t_*): 4+4+2 panels (equal on 8 GPUs, profiles/r02ad_*) and a cap on NCCL's CTAs (8 / 16 / 24 / 32) were all slower —
    // every extra graph segment costs more than the shorter tail saves, and NCCL under a GEMM runs at ~110 GB/s whatever its CTA count.
    int widths[3] = {g.tiles_n, 0, 0};
    if (ctx->world > 1 && g.tiles_n >= 10) { widths[0] = (g.tiles_n * 6) / 10; widths[1] = g.tiles_n - widths[0]; }
    ctx->panel_end[0] = ctx->panel_end[1] = 0;
    int col0 = 0;
    for (int i = 0; i < 3 && widths[i] > 0; i++) {
      tc::GemmArgs gp = g;
      const bool last = (i == 2 || widths[i + 1] == 0);
      gp.tiles_n = widths[i];
      gp.cols = last ? g.cols - col0 * bn : widths[i] * bn;
      gp.b_row0 = col0 * bn;
      gp.C = g.C + (size_t)col0 * bn * g.ldc;
      if (g.addend) gp.addend = g.addend + (size_t)col0 * bn * g.ldc;
      const bool pair = bn == 256 && g.tiles_m % 2 == 0;   // cta_group::2 pairs on 256 x 256 tiles
      tc::launch_gemm_nt(bn, s->tmdGT, pair ? s->tmZT : (bn == 256 ? s->tmZT256 : s->tmZT), gp, ctx->st, false, pair);
      LSTM_LAUNCHED(1);
      col0 += widths[i];
      if (!last) {
        ctx->panel_end[i] = (size_t)col0 * bn * (size_t)N4;
        int rc2 = lstm_allreduce_bucket(ctx, 2 + i);
        if (rc2) return rc2;
      }
    }
  }
  PROF(6);
  return lstm_allreduce_bucket(ctx, 0);
}

void tc_variant(lstm_ctx* ctx, int out[8]) {
  Bf16State* s = ctx->tc;
  out[0] = s->BN2;
  out[1] = tc::step_pair(s->Bp) ? 1 : 0;
  out[2] = s->BN5;
  out[3] = 0;
  out[4] = (ctx->M + ctx->N + 1 >= 1024) ? 256 : 128;
  out[5] = s->bn2r;
  out[6] = s->bnj5;
}

int tc_debug_read(lstm_ctx* ctx, long long out[32]) {
  Bf16State* s = ctx->tc;
  if (!s || !s->dbg) return lstm_fail(ctx, LSTM_ERR_STATE, "set LSTM_TC_DEBUG=1 before lstm_create (bf16 contexts only)");
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  LSTM_CUDA(cudaMemcpy(out, s->dbg, 32 * sizeof(long long), cudaMemcpyDeviceToHost));
  return LSTM_OK;
}

int tc_get_activation(lstm_ctx* ctx, int what, int t, float* out, size_t n) {
  Bf16State* s = ctx->tc;
  const int M = ctx->M, N = ctx->N, B = ctx->B, T = ctx->T;
  const size_t Bp = s->Bp, N4 = s->N4;
  size_t cnt = 0;
  if (t < 0 || t > T) return lstm_fail(ctx, LSTM_ERR_ARG, "t outside [0, S)");
  switch (what) {
    case LSTM_ACT_H:
      cnt = (size_t)B * N;
      if (n != cnt) break;
      tc::launch_bf16_to_f32(s->Hbf + (size_t)t * Bp * N, s->scratch, cnt, ctx->st);
      break;
    case LSTM_ACT_G:
      cnt = (size_t)B * N4;
      if (t < 1 || n != cnt) break;
      tc::launch_unpermute_f32(s->Gp + (size_t)(t - 1) * B * N4, s->scratch, B, N, ctx->st);
      break;
    case LSTM_ACT_DG:
      cnt = (size_t)B * N4;
      if (t < 1 || n != cnt) break;
      tc::launch_unpermute_bf16(s->dGbf + (size_t)(t - 1) * Bp * N4, s->scratch, B, N, ctx->st);
      break;
    case LSTM_ACT_PROBS:
      cnt = (size_t)B * M;
      if (t < 1 || n != cnt) break;
      tc::launch_bf16_to_f32(s->dYbf + (size_t)(t - 1) * Bp * M, s->scratch, cnt, ctx->st);
      break;
    default:
      return lstm_fail(ctx, LSTM_ERR_UNSUPPORTED, "activation not materialised by the bf16 path (dHy is fused into K5)");
  }
  if (n != cnt || (t < 1 && what != LSTM_ACT_H)) return lstm_fail(ctx, LSTM_ERR_ARG, "lstm_get_activation: wrong size or t");
  LSTM_LAUNCHED(1);
  LSTM_CUDA(cudaMemcpyAsync(out, s->scratch, cnt * sizeof(float), cudaMemcpyDeviceToHost, ctx->st));
  LSTM_CUDA(cudaStreamSynchronize(ctx->st));
  if (what == LSTM_ACT_PROBS) {
    std::vector<int> tgt(B);
    LSTM_CUDA(cudaMemcpy(tgt.data(), ctx->tg + (size_t)t * B, B * sizeof(int), cudaMemcpyDeviceToHost));
    for (int b = 0; b < B; b++)
      if (tgt[b] >= 0) out[(size_t)b * M + tgt[b]] += 1.0f;
  }
  return LSTM_OK;
}
