// tc_kernels.cuh — launcher prototypes of the bf16 tcgen05 kernels (tc_kernels.cu).
//
// Device layouts of the bf16 path (Bp = B rounded up to 128: the tensor-core tile height; padded
// rows/columns stay zero).  r' = 4*j + gate is the UNIT-MAJOR gate-row order used wherever the four
// gates of one hidden unit must sit next to each other (gate in {i,o,f,u} = {0,1,2,3}):
//   Hbf   [(T+1)][Bp][N]  bf16   h_t, slot t            (A operand of K2 and K3)
//   Urk   [4N r'][N]      bf16   U(gate*N+j, k)         (B operand of K2)
//   Ukr   [N k][4N r']    bf16   same matrix, k-major   (B operand of K5)
//   Wp    [M][4N r']      fp32   W permuted (gathered by K2's epilogue), bp [4N r'] fp32
//   Wmn   [M][N]          bf16   Why(m,n)               (B operand of K3)
//   Wnm   [N][M]          bf16   Why(m,n) transposed    (B operand of K5's second K segment)
//   Gp    [T][B][4N r']   fp32   activated gates of timestep t at slot t-1
//   Cs    [(T+1)][B][N]   fp32   c_t (tanh'd), slot t   (shared with the fp32 path)
//   dYbf  [T][Bp][M]      bf16   p - onehot at slot t-1 (A operand of K5's second K segment)
//   dYT   [M][T*Bp]       bf16   same, transposed       (A operand of K6c)
//   dGbf  [T][Bp][4N r']  bf16   dg_t at slot t-1       (A operand of K5)
//   dGT   [4N r][T*Bp]    bf16   dg, MASTER row order r = gate*N+j, transposed (A operand of K6a)
//   ZT    [M+N+16][(T+1)*Bp] bf16  rows [0,M): onehot(x_{s+1}) at slot s; rows [M,M+N): h_s at slot s;
//                                  row M+N: ones  (B operand of K6a / K6c: one GEMM yields dW|dU|db)
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

struct FwdStepArgs {
  int B, Bp, N, M;
  int a_row0;                  // first Hbf row of h_{t-1}: (t-1)*Bp
  const int* x;                // [B] input byte of this timestep (-1 = zero column)
  const float* Wp;             // [M][4N r']
  const float* bp;             // [4N r']
  const float* c_prev;         // [B][N]
  float* c_out;                // [B][N]
  float* Gp_t;                 // [B][4N r']
  __nv_bfloat16* Hbf_t;        // [Bp][N]
  __nv_bfloat16* ZT_h;         // ZT + M*ldz + t*Bp : h rows of this slot
  long ldz;                    // ZT leading dimension (columns)
  long long* dbg;              // optional: clock64 stamps of CTA (0,0) for diagnostics (NULL = off)
};

// persistent forward recurrence (tc_recur.cu): all T timesteps in one launch
struct FwdRecurArgs {
  int B, Bp, N, M, T;
  const int* xs;               // [S][B] input bytes; timestep t reads row t
  const float* Wp;             // [M][4N r']
  const float* bp;             // [4N r']
  float* Cs;                   // [(T+1)][B][N]: slot 0 read once, slot t written by timestep t
  float* Gp;                   // [T][B][4N r']
  __nv_bfloat16* Hbf;          // [(T+1)][Bp][N]: slot t-1 read (TMA), slot t written by timestep t
  __nv_bfloat16* ZT_h0;        // ZT + M*ldz: the h rows; timestep t writes columns [t*Bp, (t+1)*Bp)
  long ldz;
  unsigned int* gbar;          // [2][8] per-batch-half arrival counters (zeroed by the launcher)
  long long* dbg;
};

// persistent BPTT recurrence (tc_recur.cu): all T timesteps in one launch, cta_group::2 pairs + KS-way split-K
struct BwdRecurArgs {
  int B, Bp, N, M, T;
  const float* Gp;             // [T][B][4N r'] activated gates
  const float* Cs;             // [(T+1)][B][N]
  __nv_bfloat16* dGbf;         // [T][Bp][4N r']: slot t (= dg(t+1)) read by TMA, slot t-1 written by timestep t
  __nv_bfloat16* dGT;          // [4N][T*Bp]: timestep t writes columns [(t-1)*Bp, t*Bp)
  long ldg;
  float* red;                  // split-K exchange scratch [tile][dst][src][8][128][4] fp32 (bwd_recur_red_floats)
  unsigned int* xcnt;          // [N/BNJ * 2] per-(tile, batch half) arrival counters of the exchange (zeroed by the launcher)
  unsigned int* gbar;          // [2][8] per-batch-half arrival counters of the dg barrier (zeroed by the launcher)
  long long* dbg;              // optional clock64 stamps of CTA 0 around one timestep (NULL = off)
  int dbg_s = 0;               // which timestep is stamped (0-based from the end of the window; 0 = the fifth)
};

constexpr int R_SLOTS = 8;     // arrival counters per batch half of the persistent recurrences (spreads the same-address atomics)

struct LogitsArgs {
  int B, Bp, N, M, T;
  // optional: the arrival counters of a forward recurrence that is STILL RUNNING (k_fwd_recur, [Bp/128][R_SLOTS]): the tile of
  // timestep t and batch half mb is contracted once all R_SLOTS counters of mb have reached t * per_slot
  const unsigned int* progress = nullptr;
  unsigned int per_slot = 0;
  const float* by;             // [M]
  const int* tg;               // [T][B] targets of timesteps 1..T
  __nv_bfloat16* dYbf;         // [T*Bp][M]
  __nv_bfloat16* dYT;          // [M][T*Bp]
  float* surp;                 // [T][B]
};

struct BwdStepArgs {
  int B, Bp, N, M;
  int first;                   // t == T: no dg_{t+1}, dcnext = 0
  int dg_row0;                 // first dGbf row of dg_{t+1}: t*Bp
  int dy_row0;                 // first dYbf row of dy_t: (t-1)*Bp
  const float* Gp_t;           // [B][4N r'] gates of timestep t
  const float* c_t;            // [B][N]
  const float* c_prev;         // [B][N]
  float* dcnext;               // [B][N] in/out
  __nv_bfloat16* dGbf_t;       // [Bp][4N r']
  __nv_bfloat16* dGT_t;        // dGT + (t-1)*Bp
  long ldg;                    // dGT leading dimension (columns)
  float* red;                  // split-K exchange scratch: [tiles][4 dst][4 src][128][BN/4] fp32 (L2-resident)
  long long* dbg;              // optional: clock64 stamps of CTA (0,0,0) for diagnostics (NULL = off)
};

struct GemmArgs {
  int rows, cols;              // valid output extent: C(row, col) for row < rows, col < cols
  int nkb;                     // K / 64
  int a_k0, b_k0;              // starting K coordinate (elements) in A / B
  int b_row0;                  // first B row
  float* C;                    // C[col*ldc + row]
  long ldc;
  int tiles_m, tiles_n;
  int splits;                  // split-K: nkb is the k-blocks PER split; split s writes to C + s * split_stride (1 = off)
  size_t split_stride;
  const float* addend = nullptr;   // optional: C = product + addend (same layout as C): the partial sum of another K range
  // optional: contract only once the first wait_n arrival counters at wait_slots have reached wait_target — the kernel runs
  // beside the persistent BPTT recurrence that is still producing the rest of dG (tc_backward)
  const unsigned int* wait_slots = nullptr;
  int wait_n = 0;
  unsigned int wait_target = 0;
};

// K2 runs as cta_group::2 CTA pairs over the batch tiles when Bp/128 is even: each CTA then stages only half of the weight tile,
// so the weight map needs a box of BN/2 rows.
bool step_pair(int Bp);
// K2: one recurrent timestep.  BN in {32, 64, 128} gate columns per CTA.
void launch_fwd_step(int BN, const CUtensorMap& tmH, const CUtensorMap& tmUrk, const FwdStepArgs& a, cudaStream_t st);
int fwd_recur_bn(int N, int Bp, int M);
int fwd_recur_box_rows(int bn, int Bp);   // rows of the blocked-U TMA box: bn/2 for pairs (Bp = 256), bn for single CTAs (Bp = 128)
// tmWb = blocked U, Wb2[(tile*N/64 + kb)*bn + row][c] = U(r' = tile*bn + row, k = kb*64 + c) (r' = 4*unit + gate), as a 2D map
// with a box of bn/2 rows; tmH box = 128 rows
// dry = true: launch nothing, only answer whether all CTAs would be co-resident on this device
bool launch_fwd_recur(int bn, const CUtensorMap& tmH, const CUtensorMap& tmWb, const FwdRecurArgs& a, cudaStream_t st, bool dry = false);
int bwd_recur_ctas(int bnj, int N, int Bp);
int bwd_recur_per_slot(int bnj, int N, int Bp);               // arrivals per counter and timestep of k_bwd_recur's dg barrier
int fwd_recur_ctas(int bn, int N, int Bp);                     // CTAs of that launch; arrivals per counter and timestep = tiles / R_SLOTS
// tmWb = blocked BPTT weights, Wb5[(tile*NKBG + kbg)*bnj + row][c]: kbg < 4N/64: U(r' = kbg*64 + c, j = tile*bnj + row), else
// Why(m = (kbg - 4N/64)*64 + c, j); NKBG = 4N/64 + M/64; box of bnj/2 rows
int bwd_recur_bnj(int N, int Bp, int M);
size_t bwd_recur_red_floats(int N, int bnj);
int bwd_recur_box_rows(int bnj, int Bp);
bool launch_bwd_recur(int bnj, const CUtensorMap& tmdG, const CUtensorMap& tmWb, const CUtensorMap& tmdY, const BwdRecurArgs& a,
                      cudaStream_t st, bool dry = false);
// K3: logits + softmax + loss + dy for all timesteps
// beside = true: launched as the programmatic dependent of the forward recurrence that precedes it in the stream, on max_ctas CTAs
void launch_logits(const CUtensorMap& tmH, const CUtensorMap& tmWmn, const LogitsArgs& a, cudaStream_t st, int max_ctas = 0, bool beside = false);
// K5: one BPTT timestep.  BN in {32, 64, 128} hidden units per tile (4 split-K CTAs each).
void launch_bwd_step(int BN, const CUtensorMap& tmdG, const CUtensorMap& tmUkr, const CUtensorMap& tmdY,
                     const CUtensorMap& tmWnm, const BwdStepArgs& a, cudaStream_t st);
// K6: C = A * B^T, both K-major bf16, fp32 out, 128 x bn tiles (bn = 128 | 256; tmB box = bn rows)
// beside = true: launched as the programmatic dependent of the kernel that precedes it in the stream (runs next to it)
// pair = true (bn = 256, tiles_m even, no split-K): cta_group::2 pairs on 256 x 256 tiles; tmB must then have a box of 128 rows
void launch_gemm_nt(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st, bool beside = false,
                    bool pair = false);

// every bf16 operand copy of U from ONE pass over the fp32 master (null outputs skipped); N % 64 == 0
void launch_refresh_u(const float* U, __nv_bfloat16* Urk, __nv_bfloat16* Ukr, __nv_bfloat16* Wb2, int bn2, __nv_bfloat16* Wb5,
                      int bn5, int M, int N, cudaStream_t st);
void launch_block_why(const float* Why, __nv_bfloat16* Wb5, int N, int M, int bn5, cudaStream_t st);
// out[i] = sum over the splits of parts[s * stride + i], fixed order; n and stride multiples of 4
void launch_sum_splits(const float* parts, float* out, size_t n, size_t stride, int splits, cudaStream_t st);
// parameter / state conversion kernels
void launch_permute_rows_f32(const float* in, float* out, int rows, int N, cudaStream_t st);           // out[k][4j+g] = in[k][gN+j]
void launch_permute_rows_bf16(const float* in, __nv_bfloat16* out, int rows, int N, cudaStream_t st);  // same, cast
// out[c'][r] = bf16(in[r][src(c')]), src(c') = (c'%4)*N4/4 + c'/4 if permute else c'   (in: [R][C] fp32)
void launch_transpose_cast(const float* in, __nv_bfloat16* out, int R, int C, int permute, cudaStream_t st);
void launch_cast_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t st);
// ZT rows [0,M): one-hot of xs (slot s <- xs[(s+1)*B + b]); cols = T*Bp
void launch_build_xt(const int* xs1, __nv_bfloat16* ZT, long ldz, int M, int T, int B, int Bp, cudaStream_t st, bool beside = false);
void launch_fill_bf16(__nv_bfloat16* p, float v, size_t n, cudaStream_t st);
// h state fp32 [B][N] <-> Hbf slot [Bp][N] + ZT h-rows of one slot
void launch_state_to_bf16(const float* h, __nv_bfloat16* Hbf_slot, __nv_bfloat16* ZT_h, long ldz, int B, int N, cudaStream_t st);
void launch_state_to_f32(const __nv_bfloat16* Hbf_slot, float* h, int B, int N, cudaStream_t st);
// un-permute helpers for introspection: out[b][g*N+j] = in[b][4j+g]
void launch_unpermute_f32(const float* in, float* out, int rows, int N, cudaStream_t st);
void launch_unpermute_bf16(const __nv_bfloat16* in, float* out, int rows, int N, cudaStream_t st);
void launch_bf16_to_f32(const __nv_bfloat16* in, float* out, size_t n, cudaStream_t st);

}  // namespace tc
