// tc_recur.cu — the BPTT recurrence of a whole window as ONE persistent kernel (K5).
//
//   R/lstm.cc:223-257   for t = S-1 .. 1:  dh = Why^T dy(t) + U^T dg(t+1);  gate gradients;  dcnext
//
// Per timestep the contraction is D[b][j] = sum_r' dg(t+1)[b][r'] U[r'][j] + sum_m dy(t)[b][m] Why[m][j]: a 256 x N output
// over K = 4N + M.  The output is tiny and K is long, so the work is split over K:
//
//   * tile = 256 streams x BNJ hidden units, computed by a cta_group::2 PAIR (each CTA owns 128 streams = 128 TMEM lanes and
//     stages its own dg tile plus HALF of the weight tile: (128 + BNJ/2) operand rows per k-block instead of 128 + BNJ);
//   * KS split-K ranks per tile; rank ks contracts NKB_U/KS k-blocks of U (+ one k-block of Why for ks < M/64);
//   * grid = N/BNJ tiles x KS ranks pairs = 128 CTAs at N = 2048 for <256, 8> and <128, 4>, all co-resident for the whole
//     window (one launch; every spin is bounded and traps instead of hanging).
//
// One timestep of a CTA (tile jt, rank ks, batch half mb):
//   producer warp   weight tiles (and the dy / Why k-block, which no timestep produces) of the NEXT timestep are prefetched
//                   while the epilogue still runs; the dg(t+1) tiles wait for the batch half's grid barrier (global arrival
//                   counters, red.release.gpu / ld.acquire.gpu + fence.proxy.async: generic-proxy stores -> TMA reads)
//   MMA warp        (pair leader) M = 256, N = BNJ tcgen05.mma into the pair's TMEM accumulator
//   16 epilogue warps
//     1. drain      TMEM -> the KS column slices of 32 hidden units: the own slice to shared memory, the other KS-1 to an
//                   L2-resident exchange buffer laid out so that every warp store is 512 contiguous bytes
//     2. exchange   one release-add on the tile's arrival counter; acquire-poll until all KS ranks have arrived
//     3. reduce     every thread sums the KS partials of its float4 positions in a FIXED order (deterministic), coalesced
//                   512-byte warp loads, all issued before the first use
//     4. math       lane = hidden unit: gate gradients from the stashed activations; dcnext never leaves its registers;
//                   dg(t) goes to the bf16 operand buffer (+ fence) and the batch half's barrier is released
//     5. off the critical path: dg^T rows for the weight-gradient GEMM (K6a), coalesced through shared memory
//
// Weights are read from a BLOCKED copy (tc_path.cu: Wb5[tile][k-block][BNJ rows][64]) so that each TMA box is one contiguous
// 8-16 KB run of memory instead of 64-128 rows 16 KB apart.
#include <stdlib.h>

#include "tc_kernels.cuh"
#include "tc_tile.cuh"

namespace tc {

namespace {

constexpr int R_EPI_WARPS = 16;
constexpr int R_EPI_THREADS = R_EPI_WARPS * 32;
constexpr int R_CTA_THREADS = 64 + R_EPI_THREADS;
constexpr int R_HT_LD = 130;             // bf16 row pitch of the transposed staging tile: 65 words -> conflict-free

__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t cta_rank) {
  const uint32_t addr = mapa_u32(smem_u32(local_bar), cta_rank);
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// PAIR = true: two batch tiles (128 < B <= 256) as a cta_group::2 pair; PAIR = false: one batch tile (B <= 128), cta_group::1.
template <bool PAIR> __device__ __forceinline__ void r_tma(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  if constexpr (PAIR) tma_load_2d_pair(dst, tm, bar, c0, c1);
  else tma_load_2d(dst, tm, bar, c0, c1);
}
template <bool PAIR> __device__ __forceinline__ void r_tma_w(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  if constexpr (PAIR) tma_load_2d_pair_hint(dst, tm, bar, c0, c1, L2_EVICT_LAST);   // weights: re-read every timestep
  else tma_load_2d_hint(dst, tm, bar, c0, c1, L2_EVICT_LAST);
}
template <bool PAIR> __device__ __forceinline__ void r_mma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if constexpr (PAIR) umma_bf16_pair(d, ad, bd, idesc, acc);
  else umma_bf16(d, ad, bd, idesc, acc);
}
template <bool PAIR> __device__ __forceinline__ void r_commit(uint64_t* bar) {
  if constexpr (PAIR) umma_commit_pair(bar, (uint16_t)0x3);
  else umma_commit(bar);
}
template <int BN, int STAGES, bool PAIR> __device__ __forceinline__ TileCtx r_prologue(uint8_t* raw) {
  if constexpr (PAIR) return pair_prologue<BN, STAGES>(raw);
  else return tile_prologue<BN, STAGES>(raw);
}
template <int BN, int STAGES, bool PAIR> __device__ __forceinline__ void r_epilogue_end(const TileCtx& c) {
  if constexpr (PAIR) pair_epilogue_end<BN, STAGES>(c);
  else tile_epilogue_end<BN, STAGES>(c);
}

// ------------------------------------------------------------------------------------------------------------------
// Forward recurrence (K2) of a whole window as one persistent kernel.
//   R/lstm.cc:173-192   for t = 1 .. S-1:  g = W x(t) + U h(t-1) + b;  gates;  c(t) = tanh(i u + f c(t-1));  h(t) = o c(t)
// 4N/BN gate-column tiles, each computed for all 256 streams by a cta_group::2 pair (one M = 256 MMA per k-step; each CTA
// stages its own h tile and half of the U tile).  The U tiles of the NEXT timestep's first k-blocks are prefetched during the
// epilogue; h(t) travels through global memory (generic stores -> fence.proxy.async -> red.release.gpu on the batch half's
// arrival counters -> producers acquire-poll -> fence.proxy.async -> TMA); c(t) stays in the registers of the thread that owns
// the (stream, unit) pair; the gate stash, c(t) and the h^T rows for K6 are stored after h(t) has been announced.
// U is read from a blocked copy (Wb2[tile][k-block][BN rows][64]): every TMA box is one contiguous 8 KB run.
// ------------------------------------------------------------------------------------------------------------------
constexpr int SMEM_LIMIT = 227 * 1024;   // opt-in dynamic shared memory per CTA on sm_100

template <int BN, bool PAIR>
struct FwdRecurCfg {
  // The mainloop is bound by bytes in flight (L2 latency under load ~1.7 k cycles): ring depth comes first.  The accumulator is
  // therefore drained in CHUNKS of 64 gate columns (16 hidden units) through a 35 KB staging tile instead of a 67 KB one, which
  // pays for a sixth stage.
  static constexpr int STAGES = 6;
  static constexpr int UT = BN / 4;                        // hidden units per tile
  static constexpr int CW = BN < 64 ? BN : 64;             // gate columns per drain chunk
  static constexpr int CH = BN / CW;                       // chunks
  static constexpr int ACC_LD = CW + 4;
  static constexpr int ACC_BYTES = 128 * ACC_LD * 4;
  static constexpr int HT_BYTES = UT * R_HT_LD * 2;
  static constexpr int X_BYTES = 128 * 4;
  static constexpr int CC_BYTES = 128 * UT * 4;            // c(t-1) of the tile's (stream, unit) pairs, carried across timesteps
  static constexpr int EPI_BYTES = (ACC_BYTES + HT_BYTES + X_BYTES + CC_BYTES + 127) / 128 * 128;
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;        // weight-tile rows staged by each CTA
  static constexpr int B_HALF_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_HALF_BYTES;
  static constexpr int FIXED_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_BYTES + 1024;
  // Recurrent weights RESIDENT in shared memory: whatever is left after the operand ring and the epilogue tiles holds the
  // first RES k-blocks of this CTA's U half-tile for the whole window (loaded once); only the remaining k-blocks are streamed
  // from L2 every timestep.  (U in bf16 is 33.5 MB at N = 2048, the SMs have 33.6 MB of shared memory in total: full residency
  // is impossible there; at N <= 1024 the whole tile fits.)
  static constexpr int RES = (SMEM_LIMIT - FIXED_BYTES) / B_HALF_BYTES;
  static constexpr int SMEM_BYTES = FIXED_BYTES + RES * B_HALF_BYTES;
};

// PAIR: grid (2 * n_tiles), cluster (2,1,1): blockIdx.x & 1 = pair member = batch half.  Otherwise grid (n_tiles).
template <int BN, bool PAIR>
__global__ void __launch_bounds__(R_CTA_THREADS, 1)
k_fwd_recur(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmWb, const FwdRecurArgs a) {
  using F = FwdRecurCfg<BN, PAIR>;
  constexpr uint32_t NCTA = PAIR ? 2u : 1u;                  // CTAs whose loads complete on the (leader's) full barrier
  constexpr int STAGES = F::STAGES, UT = F::UT, ACC_LD = F::ACC_LD, CH = F::CH;
  constexpr int UC = F::CW / 4, RG = R_EPI_THREADS / UC, RPC = 128 / RG;   // per chunk: 16 units x 32 row groups, 4 rows per thread
  extern __shared__ uint8_t smem_raw[];
  TileCtx c = r_prologue<BN, STAGES, PAIR>(smem_raw);
  // this CTA is resident: a kernel launched as our programmatic dependent (k_logits, following the arrival counters) may be
  // scheduled on the SMs this grid leaves free once every CTA has said so
  if (threadIdx.x == 0) pdl_launch_dependents();
  c.epi = c.tiles + STAGES * F::STAGE_BYTES + 256;
  uint64_t* tmem_free = c.accum_full + 2;
  uint64_t* res_full = c.accum_full + 3;                     // (leader) the CTAs' resident U k-blocks have landed
  if (threadIdx.x == 0) { mbar_init(tmem_free, NCTA); mbar_init(res_full, 1); fence_barrier_init(); }
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  uint8_t* res;                                              // resident U k-blocks: [nres][B_ROWS][128 B], 1024-byte aligned
  {
    const uint32_t e0 = smem_u32(c.epi) + (uint32_t)F::EPI_BYTES;
    res = c.epi + F::EPI_BYTES + (((e0 + 1023u) & ~1023u) - e0);
  }
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int nb = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int mb = (int)rank;
  const int n_tiles = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int nkb = a.N / BK;
  const unsigned int per_slot = (unsigned int)(n_tiles / R_SLOTS);
  unsigned int* my_slots = a.gbar + (size_t)mb * R_SLOTS;
  const int N = a.N, N4 = 4 * a.N, B = a.B;
  const int wrow0 = nb * nkb * BN + (int)rank * F::B_ROWS;   // + kb * BN
  const int nres = F::RES < nkb ? F::RES : nkb;              // k-blocks [0, nres) of U never leave shared memory
  constexpr int DBG_T = 4;
  long long* dbg = (a.dbg && blockIdx.x == 0 && a.T > DBG_T) ? a.dbg : nullptr;

  if (c.warp == 0) {
    // ---------------- producer ----------------
    if (elect_one()) { tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmWb); }
    __syncwarp();
    if (elect_one()) {                                       // once: the resident k-blocks
      if (rank == 0) mbar_expect_tx(res_full, NCTA * (uint32_t)nres * (uint32_t)F::B_HALF_BYTES);
      for (int kb = 0; kb < nres; kb++)
        r_tma<PAIR>(res + (size_t)kb * F::B_HALF_BYTES, &tmWb, res_full, 0, wrow0 + kb * BN);
    }
    __syncwarp();
    const int pre = nkb < STAGES ? nkb : STAGES;
    int g = 0;
    for (int t = 1; t <= a.T; t++) {
      const int a_row = (t - 1) * a.Bp + mb * BM;
      for (int kb = 0; kb < pre; kb++) {                     // streamed weights first: they do not depend on h(t-1)
        const int st = (g + kb) % STAGES;
        const uint32_t ph = (uint32_t)((g + kb) / STAGES) & 1u;
        mbar_wait(&c.empty[st], ph ^ 1u);
        if (elect_one()) {
          uint8_t* bdst = c.tiles + (size_t)st * F::STAGE_BYTES + A_TILE_BYTES;
          if (rank == 0) mbar_expect_tx(&c.full[st], NCTA * (uint32_t)(kb < nres ? A_TILE_BYTES : F::STAGE_BYTES));
          if (kb >= nres) r_tma_w<PAIR>(bdst, &tmWb, &c.full[st], 0, wrow0 + kb * BN);
        }
        __syncwarp();
      }
      if (t > 1) {                                           // h(t-1) of this batch half is complete in global memory
        if (dbg && c.lane == 0 && (t == DBG_T || t == DBG_T + 1)) dbg[t == DBG_T ? 0 : 8] = clock64();
        counters_wait(my_slots, R_SLOTS, (unsigned int)(t - 1) * per_slot, c.lane);
        fence_proxy_async_global();
        if (dbg && c.lane == 0 && t == DBG_T) dbg[1] = clock64();
      }
      for (int kb = 0; kb < nkb; kb++) {
        const int st = (g + kb) % STAGES;
        const uint32_t ph = (uint32_t)((g + kb) / STAGES) & 1u;
        if (kb >= pre) mbar_wait(&c.empty[st], ph ^ 1u);
        if (elect_one()) {
          uint8_t* adst = c.tiles + (size_t)st * F::STAGE_BYTES;
          if (kb >= pre) {
            if (rank == 0) mbar_expect_tx(&c.full[st], NCTA * (uint32_t)(kb < nres ? A_TILE_BYTES : F::STAGE_BYTES));
            if (kb >= nres) r_tma_w<PAIR>(adst + A_TILE_BYTES, &tmWb, &c.full[st], 0, wrow0 + kb * BN);
          }
          r_tma<PAIR>(adst, &tmH, &c.full[st], kb * BK, a_row);
        }
        __syncwarp();
      }
      g += nkb;
    }
  } else if (c.warp == 1) {
    // ---------------- MMA issuer (leader CTA of the pair) ----------------
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, BN);
      const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(c.tiles));
      const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(c.tiles) + A_TILE_BYTES);
      const uint64_t r_desc0 = make_smem_desc_sw128(smem_u32(res));
      mbar_wait(res_full, 0);                                // the resident k-blocks of both CTAs are in shared memory
      int g = 0;
      for (int t = 1; t <= a.T; t++) {
        if (t > 1) {
          mbar_wait_cluster(tmem_free, (uint32_t)(t - 2) & 1u);
          tcgen05_after_sync();
        }
        for (int kb = 0; kb < nkb; kb++) {
          const int st = (g + kb) % STAGES;
          const uint32_t ph = (uint32_t)((g + kb) / STAGES) & 1u;
          mbar_wait(&c.full[st], ph);
          if (dbg && c.lane == 0 && t == DBG_T && kb == 0) dbg[2] = clock64();
          tcgen05_after_sync();
          if (elect_one()) {
            const uint64_t soff = (uint64_t)((uint32_t)st * (uint32_t)(F::STAGE_BYTES >> 4));
            const uint64_t bd = kb < nres ? r_desc0 + (uint64_t)((uint32_t)kb * (uint32_t)(F::B_HALF_BYTES >> 4)) : b_desc0 + soff;
#pragma unroll
            for (int k = 0; k < BK / 16; k++)
              r_mma<PAIR>(c.tmem_d, a_desc0 + soff + 2 * k, bd + 2 * k, idesc, (uint32_t)((kb | k) != 0));
            r_commit<PAIR>(&c.empty[st]);
          }
          __syncwarp();
        }
        if (elect_one()) r_commit<PAIR>(c.accum_full);
        __syncwarp();
        if (dbg && c.lane == 0 && t == DBG_T) dbg[3] = clock64();
        g += nkb;
      }
    }
  } else {
    // ---------------- epilogue: LSTM math, lane = hidden unit ----------------
    float* acc = reinterpret_cast<float*>(c.epi);
    __nv_bfloat16* hT = reinterpret_cast<__nv_bfloat16*>(c.epi + F::ACC_BYTES);
    int* sx = reinterpret_cast<int*>(c.epi + F::ACC_BYTES + F::HT_BYTES);
    const int e = threadIdx.x - 64;
    const int l2 = e % UC, rg = e / UC;                      // unit inside a chunk, row group
    // c(t-1) of this thread's (stream, unit) pairs never leaves the SM: [ch][q][thread] in shared memory (registers are the
    // scarce resource of this 576-thread CTA: 96 per thread)
    float* cc_s = reinterpret_cast<float*>(c.epi + F::ACC_BYTES + F::HT_BYTES + F::X_BYTES);
#pragma unroll
    for (int ch = 0; ch < CH; ch++) {
      const int j = nb * UT + ch * UC + l2;
#pragma unroll
      for (int q = 0; q < RPC; q++) {
        const int b = mb * BM + rg + RG * q;
        cc_s[(ch * RPC + q) * R_EPI_THREADS + e] = b < B ? a.Cs[(size_t)b * N + j] : 0.f;  // slot 0 = carried-in state
      }
    }
    for (int t = 1; t <= a.T; t++) {
      const int* x_t = a.xs + (size_t)t * B;
      float* c_out = a.Cs + (size_t)t * B * N;
      float* Gp_t = a.Gp + (size_t)(t - 1) * B * N4;
      __nv_bfloat16* Hbf_t = a.Hbf + (size_t)t * a.Bp * N;
      __nv_bfloat16* ZT_h = a.ZT_h0 + (size_t)t * a.Bp;
      if (e < 128) {
        const int b = mb * BM + e;
        sx[e] = (b < B) ? x_t[b] : -2;                       // -2 = padding row, -1 = all-zero input column
      }
      named_bar_sync(1, R_EPI_THREADS);
      float4 w[CH][RPC];
#pragma unroll
      for (int ch = 0; ch < CH; ch++) {                      // W*x for one-hot x = a row gather; in flight during the contraction
        const float4 bias = __ldg(reinterpret_cast<const float4*>(a.bp + 4 * (nb * UT + ch * UC + l2)));
#pragma unroll
        for (int q = 0; q < RPC; q++) {
          const int x = sx[rg + RG * q];
          w[ch][q] = bias;
          if (x >= 0) {
            const float4 wr = __ldg(reinterpret_cast<const float4*>(a.Wp + (size_t)x * N4 + 4 * (nb * UT + ch * UC + l2)));
            w[ch][q] = make_float4(wr.x + bias.x, wr.y + bias.y, wr.z + bias.z, wr.w + bias.w);
          }
        }
      }
      // Only h(t) is on the critical path of the other CTAs: it is computed and stored chunk by chunk, fenced and announced; the
      // gate stash and c(t) (registers until then) follow after the arrival.
#pragma unroll
      for (int ch = 0; ch < CH; ch++) {
        if (c.warp < 6) {                                    // TMEM (lane = stream) -> shared memory, 64 gate columns
          const int quarter = c.warp & 3;
          const int row = quarter * 32 + c.lane;
          if (ch == 0) {
            mbar_wait(c.accum_full, (uint32_t)(t - 1) & 1u);
            if (dbg && e == 0 && t == DBG_T) dbg[4] = clock64();
            tcgen05_after_sync();
          }
          static_assert(F::CW % 32 == 0, "drain chunks are moved 32 columns at a time");
#pragma unroll 1
          for (int c0 = 0; c0 < F::CW; c0 += 32) {
            float v[32];
            tmem_ld32(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ch * F::CW + c0), v);
            float4* dst = reinterpret_cast<float4*>(acc + (size_t)row * ACC_LD + c0);
#pragma unroll
            for (int u = 0; u < 8; u++) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
          }
          if (ch == CH - 1) tcgen05_before_sync();
        }
        named_bar_sync(1, R_EPI_THREADS);
        if (ch == CH - 1) {
          if (dbg && e == 0 && t == DBG_T) dbg[5] = clock64();
          if (e == 0) { if constexpr (PAIR) mbar_arrive_remote(tmem_free, 0); else mbar_arrive(tmem_free); }   // accumulator drained
        }
        const int l = ch * UC + l2;                          // unit inside the tile
        const int j = nb * UT + l;
#pragma unroll
        for (int q = 0; q < RPC; q++) {
          const int r = rg + RG * q;
          float hval = 0.f;
          if (sx[r] >= -1) {
            const int b = mb * BM + r;
            const float4 pre = *reinterpret_cast<const float4*>(acc + (size_t)r * ACC_LD + 4 * l2);
            const float gi = sigmoid_fast(pre.x + w[ch][q].x);
            const float go = sigmoid_fast(pre.y + w[ch][q].y);
            const float gf = sigmoid_fast(pre.z + w[ch][q].z);
            const float gu = tanh_fast(pre.w + w[ch][q].w);
            float* ccp = cc_s + (ch * RPC + q) * R_EPI_THREADS + e;
            const float cc = tanh_fast(gi * gu + gf * *ccp);   // the carried cell value is the tanh'd one (R/lstm.cc:185-189)
            hval = go * cc;
            *ccp = cc;
            w[ch][q] = make_float4(gi, go, gf, gu);          // the W row is dead: its registers carry the activated gates
            Hbf_t[(size_t)b * N + j] = __float2bfloat16_rn(hval);
          }
          hT[l * R_HT_LD + r] = __float2bfloat16_rn(hval);
        }
        if (ch < CH - 1) named_bar_sync(1, R_EPI_THREADS);   // the staging tile is rewritten by the next chunk's drain
      }
      fence_proxy_async_global();                            // h(t) is read by other CTAs' TMA loads
      named_bar_sync(1, R_EPI_THREADS);
      if (e == 0) red_release_gpu_add(my_slots + (nb % R_SLOTS), 1u);   // release: cumulative over the barrier-ordered stores
      if (dbg && e == 0 && t == DBG_T) dbg[6] = clock64();
#pragma unroll
      for (int ch = 0; ch < CH; ch++) {
        const int j = nb * UT + ch * UC + l2;
#pragma unroll
        for (int q = 0; q < RPC; q++) {
          if (sx[rg + RG * q] >= -1) {
            const int b = mb * BM + rg + RG * q;
            __stcs(reinterpret_cast<float4*>(Gp_t + (size_t)b * N4 + 4 * j), w[ch][q]);   // streamed: read once, in BPTT
            c_out[(size_t)b * N + j] = cc_s[(ch * RPC + q) * R_EPI_THREADS + e];
          }
        }
      }
      {                                                      // h^T rows of ZT (K6 operand): off the critical path
        const int w4 = e >> 5, lane = e & 31;
        for (int u = w4; u < UT; u += R_EPI_WARPS) {
          const uint32_t* src = reinterpret_cast<const uint32_t*>(hT + u * R_HT_LD);
          uint32_t* dst = reinterpret_cast<uint32_t*>(ZT_h + (size_t)(nb * UT + u) * a.ldz + mb * BM);
          __stcs(dst + lane, src[lane]);
          __stcs(dst + lane + 32, src[lane + 32]);
        }
      }
      if (dbg && e == 0 && t == DBG_T) dbg[7] = clock64();
    }
  }
  r_epilogue_end<BN, STAGES, PAIR>(c);
}

// dry = true: only answer whether every CTA of the launch would be co-resident on this device (asked once, at context creation)
template <int BN, bool PAIR>
bool launch_fwd_recur_t(const CUtensorMap& tmH, const CUtensorMap& tmWb, const FwdRecurArgs& a, cudaStream_t st, bool dry) {
  using F = FwdRecurCfg<BN, PAIR>;
  const int n_tiles = 4 * a.N / BN;
  auto kernel = k_fwd_recur<BN, PAIR>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES) != cudaSuccess) { cudaGetLastError(); return false; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((PAIR ? 2 : 1) * n_tiles, 1, 1);
  cfg.blockDim = dim3(R_CTA_THREADS);
  cfg.dynamicSmemBytes = (size_t)F::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
  if (max_clusters < n_tiles) return false;                  // every CTA must be resident: the grid barrier spins
  if (dry) return true;
  if (cudaMemsetAsync(a.gbar, 0, 2 * R_SLOTS * sizeof(unsigned int), st) != cudaSuccess) { cudaGetLastError(); return false; }
  return cudaLaunchKernelEx(&cfg, kernel, tmH, tmWb, a) == cudaSuccess;
}

template <int BNJ, int KS, bool PAIR>
struct BwdRecurCfg {
  static constexpr int STAGES = BNJ == 256 ? 4 : 6;        // ring depth first (see FwdRecurCfg)
  static constexpr int UO = BNJ / KS;                      // hidden units finalised by each CTA
  static_assert(UO == 8 || UO == 16 || UO == 32, "one tcgen05.ld of 8, 16 or 32 columns per exchanged slice");
  static_assert(KS % 4 == 0, "four drain warps per TMEM lane quarter share the KS slices");
  static constexpr int B_ROWS = PAIR ? BNJ / 2 : BNJ;      // weight-tile rows staged by each CTA
  static constexpr int B_HALF_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_HALF_BYTES;
  static constexpr int DH_LD = UO + 4;                     // fp32 row pitch of the dh tile (16-byte aligned rows)
  static constexpr int DH_BYTES = 128 * DH_LD * 4;
  static constexpr int GT_BYTES = 4 * UO * R_HT_LD * 2;    // dg^T staging [gate*UO + unit][row]
  static constexpr int ST_BYTES = 128 * UO * 4;            // per (stream, unit): dcnext, carried across timesteps
  static constexpr int EPI_BYTES = (DH_BYTES + GT_BYTES + ST_BYTES + 127) / 128 * 128;
  static constexpr int FIXED_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_BYTES + 1024;
  static constexpr int RES = (SMEM_LIMIT - FIXED_BYTES) / B_HALF_BYTES;   // resident U^T k-blocks, see FwdRecurCfg
  static constexpr int SMEM_BYTES = FIXED_BYTES + RES * B_HALF_BYTES;
  static constexpr int CHUNKS_PER_WARP = KS / 4;           // UO-column slices each drain warp moves
  static constexpr int U4 = UO / 4;                        // float4 groups per slice row
  static constexpr size_t RED_TILE_FLOATS = (size_t)KS * KS * U4 * 128 * 4;   // [dst][src][u][row][4]
};

// PAIR: grid (2 * JT * KS), cluster (2,1,1): blockIdx.x & 1 = pair member = batch half.  Otherwise grid (JT * KS), one batch tile.
template <int BNJ, int KS, bool PAIR>
__global__ void __launch_bounds__(R_CTA_THREADS, 1)
k_bwd_recur(const __grid_constant__ CUtensorMap tmdG, const __grid_constant__ CUtensorMap tmWb,
            const __grid_constant__ CUtensorMap tmdY, const BwdRecurArgs a) {
  using F = BwdRecurCfg<BNJ, KS, PAIR>;
  constexpr int STAGES = F::STAGES, UO = F::UO, RG = R_EPI_THREADS / UO, ROWS = 128 / RG, DH_LD = F::DH_LD, U4 = F::U4;
  constexpr uint32_t NCTA = PAIR ? 2u : 1u;
  extern __shared__ uint8_t smem_raw[];
  TileCtx c = r_prologue<BNJ, STAGES, PAIR>(smem_raw);
  if (threadIdx.x == 0) pdl_launch_dependents();   // see k_fwd_recur: the dWhy|dby GEMM runs beside this grid
  c.epi = c.tiles + STAGES * F::STAGE_BYTES + 256;
  uint64_t* tmem_free = c.accum_full + 2;                    // (leader) the CTAs have read the accumulator out of TMEM
  uint64_t* res_full = c.accum_full + 3;                     // (leader) the CTAs' resident weight k-blocks have landed
  if (threadIdx.x == 0) { mbar_init(tmem_free, NCTA); mbar_init(res_full, 1); fence_barrier_init(); }
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  uint8_t* res;                                              // resident U^T k-blocks: [nres][BNJ/2 rows][128 B], 1024-byte aligned
  {
    const uint32_t e0 = smem_u32(c.epi) + (uint32_t)F::EPI_BYTES;
    res = c.epi + F::EPI_BYTES + (((e0 + 1023u) & ~1023u) - e0);
  }
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;       // pair member = batch half
  const int mb = (int)rank;
  const int pairi = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int jt = pairi / KS, ks = pairi - jt * KS;
  const int N = a.N, N4 = 4 * a.N, B = a.B, T = a.T;
  const int nkbu = N4 / BK / KS;                             // U k-blocks per rank and timestep
  const int nkbw = (ks < a.M / BK) ? 1 : 0;                  // one Why k-block for the first M/64 ranks
  const int nres = F::RES < nkbu ? F::RES : nkbu;            // this rank's first nres U k-blocks never leave shared memory
  const int NKBG = N4 / BK + a.M / BK;                       // k-blocks per tile in the blocked weight copy
  const int wrow0 = jt * NKBG * BNJ + (int)rank * F::B_ROWS; // + kbg * BNJ
  const int tile = jt * 2 + mb;
  unsigned int* my_slots = a.gbar + (size_t)mb * R_SLOTS;
  unsigned int* xcnt = a.xcnt + tile;
  const unsigned int per_slot = (unsigned int)(gridDim.x / NCTA / R_SLOTS);   // arrivals per counter and timestep
  float* red_tile = a.red + (size_t)tile * F::RED_TILE_FLOATS;             // [dst][src][u][row][4]
  long long* dbg = (a.dbg && blockIdx.x == 0) ? a.dbg : nullptr;
  const int DBG_S = a.dbg_s > 0 ? a.dbg_s : 4;               // the timestep (0-based from the end of the window) that is stamped

  if (c.warp == 0) {
    // ---------------- producer ----------------
    if (elect_one()) { tma_prefetch_desc(&tmdG); tma_prefetch_desc(&tmWb); tma_prefetch_desc(&tmdY); }
    __syncwarp();
    if (elect_one()) {                                       // once: the resident k-blocks
      if (rank == 0 && nres > 0) mbar_expect_tx(res_full, NCTA * (uint32_t)nres * (uint32_t)F::B_HALF_BYTES);
      for (int iu = 0; iu < nres; iu++)
        r_tma<PAIR>(res + (size_t)iu * F::B_HALF_BYTES, &tmWb, res_full, 0, wrow0 + (ks * nkbu + iu) * BNJ);
    }
    __syncwarp();
    int g = 0;                                               // k-blocks issued so far (ring position)
    for (int s = 0; s < T; s++) {
      const int t = T - s;
      const int nu = s == 0 ? 0 : nkbu;                      // dg(T+1) = 0: the first timestep has no U part
      const int total = nkbw + nu;
      // Phase A: everything of this timestep that no other CTA produces: the dy / Why k-block completely, and the streamed
      // weight tiles of the first U k-blocks (as many as fit in the ring)
      const int pre = total < STAGES ? total : STAGES;
      for (int i = 0; i < pre; i++) {
        const int st = (g + i) % STAGES;
        const uint32_t ph = (uint32_t)((g + i) / STAGES) & 1u;
        const bool resident = i >= nkbw && i - nkbw < nres;
        mbar_wait(&c.empty[st], ph ^ 1u);
        if (elect_one()) {
          uint8_t* adst = c.tiles + (size_t)st * F::STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&c.full[st], NCTA * (uint32_t)(resident ? A_TILE_BYTES : F::STAGE_BYTES));
          if (i < nkbw) {
            r_tma_w<PAIR>(adst + A_TILE_BYTES, &tmWb, &c.full[st], 0, wrow0 + (N4 / BK + ks) * BNJ);
            r_tma<PAIR>(adst, &tmdY, &c.full[st], ks * BK, (t - 1) * a.Bp + mb * BM);
          } else if (!resident) {
            const int kbg = ks * nkbu + (i - nkbw);
            r_tma_w<PAIR>(adst + A_TILE_BYTES, &tmWb, &c.full[st], 0, wrow0 + kbg * BNJ);
          }
        }
        __syncwarp();
      }
      if (s > 0) {                                           // dg(t+1) of this batch half is complete in global memory
        if (dbg && c.lane == 0 && (s == DBG_S || s == DBG_S + 1)) dbg[s == DBG_S ? 0 : 8] = clock64();
        counters_wait(my_slots, R_SLOTS, (unsigned int)s * per_slot, c.lane);
        fence_proxy_async_global();                          // generic-proxy writes (observed via acquire) -> async-proxy reads
        if (dbg && c.lane == 0 && s == DBG_S) dbg[1] = clock64();
      }
      for (int i = nkbw; i < total; i++) {
        const int st = (g + i) % STAGES;
        const uint32_t ph = (uint32_t)((g + i) / STAGES) & 1u;
        const int kbg = ks * nkbu + (i - nkbw);
        const bool resident = i - nkbw < nres;
        if (i >= pre) mbar_wait(&c.empty[st], ph ^ 1u);
        if (elect_one()) {
          uint8_t* adst = c.tiles + (size_t)st * F::STAGE_BYTES;
          if (i >= pre) {
            if (rank == 0) mbar_expect_tx(&c.full[st], NCTA * (uint32_t)(resident ? A_TILE_BYTES : F::STAGE_BYTES));
            if (!resident) r_tma_w<PAIR>(adst + A_TILE_BYTES, &tmWb, &c.full[st], 0, wrow0 + kbg * BNJ);
          }
          r_tma<PAIR>(adst, &tmdG, &c.full[st], kbg * BK, t * a.Bp + mb * BM);
        }
        __syncwarp();
      }
      g += total;
    }
  } else if (c.warp == 1) {
    // ---------------- MMA issuer (leader CTA of the pair) ----------------
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, BNJ);
      const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(c.tiles));
      const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(c.tiles) + A_TILE_BYTES);
      const uint64_t r_desc0 = make_smem_desc_sw128(smem_u32(res));
      if (nres > 0) mbar_wait(res_full, 0);                  // the resident k-blocks of both CTAs are in shared memory
      int g = 0, used = 0;                                   // used = timesteps that produced an accumulator so far
      for (int s = 0; s < T; s++) {
        const int total = nkbw + (s == 0 ? 0 : nkbu);
        if (total == 0) continue;
        if (used > 0) {                                      // both CTAs have drained the previous accumulator
          mbar_wait_cluster(tmem_free, (uint32_t)(used - 1) & 1u);
          tcgen05_after_sync();
        }
        for (int i = 0; i < total; i++) {
          const int st = (g + i) % STAGES;
          const uint32_t ph = (uint32_t)((g + i) / STAGES) & 1u;
          mbar_wait(&c.full[st], ph);
          if (dbg && c.lane == 0 && s == DBG_S && i == nkbw) dbg[2] = clock64();
          tcgen05_after_sync();
          if (elect_one()) {
            const uint64_t soff = (uint64_t)((uint32_t)st * (uint32_t)(F::STAGE_BYTES >> 4));
            const bool resident = i >= nkbw && i - nkbw < nres;
            const uint64_t bd = resident ? r_desc0 + (uint64_t)((uint32_t)(i - nkbw) * (uint32_t)(F::B_HALF_BYTES >> 4)) : b_desc0 + soff;
#pragma unroll
            for (int k = 0; k < BK / 16; k++)
              r_mma<PAIR>(c.tmem_d, a_desc0 + soff + 2 * k, bd + 2 * k, idesc, (uint32_t)((i | k) != 0));
            r_commit<PAIR>(&c.empty[st]);
          }
          __syncwarp();
        }
        if (elect_one()) r_commit<PAIR>(c.accum_full);
        __syncwarp();
        if (dbg && c.lane == 0 && s == DBG_S) dbg[3] = clock64();
        g += total;
        used++;
      }
    }
  } else {
    // ---------------- epilogue ----------------
    float* dh = reinterpret_cast<float*>(c.epi);             // [128][DH_LD]: own partial slice, then the summed dh
    __nv_bfloat16* gT = reinterpret_cast<__nv_bfloat16*>(c.epi + F::DH_BYTES);
    const int e = threadIdx.x - 64;
    const int w = e >> 5, lane = e & 31;
    const int quarter = c.warp & 3;                          // the TMEM lane quarter this warp may read
    const int cgrp = w >> 2;                                 // which of the four column groups of this quarter's warps
    const int l = e % UO, rg = e / UO;
    const int j = jt * BNJ + ks * UO + l;                    // the hidden unit this thread finalises
    // dcnext of this thread's (stream, unit) pairs lives in shared memory for the whole window: [q][thread]
    float* carry = reinterpret_cast<float*>(c.epi + F::DH_BYTES + F::GT_BYTES);
#pragma unroll
    for (int q = 0; q < ROWS; q++) carry[q * R_EPI_THREADS + e] = 0.f;
    int used = 0;
    for (int s = 0; s < T; s++) {
      const int t = T - s;
      const bool has_acc = nkbw + (s == 0 ? 0 : nkbu) > 0;
      const float* Gp_t = a.Gp + (size_t)(t - 1) * B * N4;
      const float* c_prev = a.Cs + (size_t)(t - 1) * B * N;
      const float* c_cur = a.Cs + (size_t)t * B * N;
      __nv_bfloat16* dGbf_t = a.dGbf + (size_t)(t - 1) * a.Bp * N4;
      __nv_bfloat16* dGT_t = a.dGT + (size_t)(t - 1) * a.Bp;
      // 1. drain: TMEM (lane = stream) -> KS slices of 32 hidden units
      {
        const int row = quarter * 32 + lane;
        if (has_acc) {
          mbar_wait(c.accum_full, (uint32_t)used & 1u);
          if (dbg && e == 0 && s == DBG_S) dbg[4] = clock64();
          tcgen05_after_sync();
        }
#pragma unroll 1
        for (int ch = 0; ch < F::CHUNKS_PER_WARP; ch++) {
          const int d = cgrp * F::CHUNKS_PER_WARP + ch;      // destination rank of this UO-column slice
          float v[UO];
          if (has_acc) {
            tmem_ldw<UO>(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(d * UO), v);
          } else {
#pragma unroll
            for (int i = 0; i < UO; i++) v[i] = 0.f;
          }
          if (d == ks) {
            float4* dst = reinterpret_cast<float4*>(dh + (size_t)row * DH_LD);
#pragma unroll
            for (int u = 0; u < U4; u++) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
          } else {
            float4* dst = reinterpret_cast<float4*>(red_tile + ((size_t)(d * KS + ks) * U4 * 128 + row) * 4);
#pragma unroll
            for (int u = 0; u < U4; u++) dst[(size_t)u * 128] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
          }
        }
        if (has_acc) tcgen05_before_sync();
      }
      named_bar_sync(1, R_EPI_THREADS);
      if (dbg && e == 0 && s == DBG_S) dbg[5] = clock64();
      // 2. exchange: announce our slices, free the accumulator, wait for the other ranks
      if (e == 0) {
        red_release_gpu_add(xcnt, 1u);                       // release: cumulative over the barrier-ordered stores
        if (dbg && s == DBG_S) dbg[9] = clock64();
        if (has_acc) { if constexpr (PAIR) mbar_arrive_remote(tmem_free, 0); else mbar_arrive(tmem_free); }
      }
      // operands of the gate math that no timestep of this launch produces: in flight while the exchange is awaited
      float4 gv[ROWS];
      float cpp[ROWS], ctt[ROWS];
#pragma unroll
      for (int q = 0; q < ROWS; q++) {
        const int b = mb * BM + rg + RG * q;
        gv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        cpp[q] = ctt[q] = 0.f;
        if (b < B) {
          gv[q] = __ldcs(reinterpret_cast<const float4*>(Gp_t + (size_t)b * N4 + 4 * (size_t)j));   // i o f u (last use)
          cpp[q] = c_prev[(size_t)b * N + j];
          ctt[q] = c_cur[(size_t)b * N + j];
        }
      }
      if (w == 1) {
        counters_wait(xcnt, 1, (unsigned int)KS * (unsigned int)(s + 1), lane);
        if (dbg && lane == 0 && s == DBG_S) dbg[10] = clock64();
      }
      named_bar_sync(1, R_EPI_THREADS);                      // every reader is ordered after the acquire
      if (dbg && e == 0 && s == DBG_S) dbg[6] = clock64();
      // 3. reduce: float4 position p = (u, row); sum over the KS sources in a fixed order (both positions' loads in flight at once)
#pragma unroll
      for (int p = e; p < U4 * 128; p += R_EPI_THREADS) {
        const int u = p >> 7, row = p & 127;
        float4 pv[KS];
#pragma unroll
        for (int sr = 0; sr < KS; sr++)
          if (sr != ks) pv[sr] = __ldcg(reinterpret_cast<const float4*>(red_tile + ((size_t)(ks * KS + sr) * U4 * 128 + (size_t)u * 128 + row) * 4));
        float4* own = reinterpret_cast<float4*>(dh + (size_t)row * DH_LD + 4 * u);
        const float4 o = *own;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int sr = 0; sr < KS; sr++) {
          const float4 x = (sr == ks) ? o : pv[sr];
          acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
        *own = acc;
      }
      named_bar_sync(1, R_EPI_THREADS);
      if (dbg && e == 0 && s == DBG_S) dbg[11] = clock64();
      // 4. gate gradients, lane = hidden unit (R/lstm.cc:233-256)
#pragma unroll
      for (int q = 0; q < ROWS; q++) {
        const int r = rg + RG * q;
        const int b = mb * BM + r;
        float d_i = 0.f, d_o = 0.f, d_f = 0.f, d_u = 0.f;
        if (b < B) {
          const float dhv = dh[(size_t)r * DH_LD + l];
          const float4 gg = gv[q];
          const float ct = ctt[q];
          const float dc = (dhv * gg.y + carry[q * R_EPI_THREADS + e]) * (1.0f - ct * ct);   // :233-235
          d_o = dhv * ct * (gg.y * (1.0f - gg.y));                             // :238,244
          d_i = dc * gg.w * (gg.x * (1.0f - gg.x));                            // :239,244
          d_f = dc * cpp[q] * (gg.z * (1.0f - gg.z));                          // :240,244
          d_u = dc * gg.x * (1.0f - gg.w * gg.w);                              // :241,247
          carry[q * R_EPI_THREADS + e] = dc * gg.z;                            // :256
          uint2 pk;
          pk.x = pack_bf16x2(d_i, d_o);
          pk.y = pack_bf16x2(d_f, d_u);
          *reinterpret_cast<uint2*>(dGbf_t + (size_t)b * N4 + 4 * (size_t)j) = pk;
        }
        gT[(0 * UO + l) * R_HT_LD + r] = __float2bfloat16_rn(d_i);
        gT[(1 * UO + l) * R_HT_LD + r] = __float2bfloat16_rn(d_o);
        gT[(2 * UO + l) * R_HT_LD + r] = __float2bfloat16_rn(d_f);
        gT[(3 * UO + l) * R_HT_LD + r] = __float2bfloat16_rn(d_u);
      }
      if (dbg && e == 0 && s == DBG_S) dbg[12] = clock64();
      fence_proxy_async_global();                            // dg(t) is read by other CTAs' TMA loads
      named_bar_sync(1, R_EPI_THREADS);
      if (e == 0) red_release_gpu_add(my_slots + (pairi % R_SLOTS), 1u);
      if (dbg && e == 0 && s == DBG_S) dbg[7] = clock64();
      // 5. dg^T rows (master row order gate*N + unit) for K6a: off the critical path
      {
        const int jbase = jt * BNJ + ks * UO;
        for (int q = w; q < 4 * UO; q += R_EPI_WARPS) {
          const int gate = q / UO, u = q - gate * UO;
          const uint32_t* src = reinterpret_cast<const uint32_t*>(gT + q * R_HT_LD);
          uint32_t* dst = reinterpret_cast<uint32_t*>(dGT_t + (size_t)(gate * N + jbase + u) * a.ldg + mb * BM);
          __stcs(dst + lane, src[lane]);
          __stcs(dst + lane + 32, src[lane + 32]);
        }
      }
      named_bar_sync(1, R_EPI_THREADS);                      // gT and dh are rewritten by the next timestep
      if (has_acc) used++;
    }
  }
  r_epilogue_end<BNJ, STAGES, PAIR>(c);
}

template <int BNJ, int KS, bool PAIR>
bool launch_bwd_recur_t(const CUtensorMap& tmdG, const CUtensorMap& tmWb, const CUtensorMap& tmdY, const BwdRecurArgs& a,
                        cudaStream_t st, bool dry) {
  using F = BwdRecurCfg<BNJ, KS, PAIR>;
  const int JT = a.N / BNJ;
  auto kernel = k_bwd_recur<BNJ, KS, PAIR>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES) != cudaSuccess) { cudaGetLastError(); return false; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((PAIR ? 2 : 1) * JT * KS, 1, 1);
  cfg.blockDim = dim3(R_CTA_THREADS);
  cfg.dynamicSmemBytes = (size_t)F::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // every CTA must be resident at once (the grid barrier spins): refuse unless all of them fit on an idle GPU
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
  if (max_clusters < JT * KS) return false;
  if (dry) return true;
  if (cudaMemsetAsync(a.gbar, 0, 2 * R_SLOTS * sizeof(unsigned int), st) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cudaMemsetAsync(a.xcnt, 0, (size_t)JT * 2 * sizeof(unsigned int), st) != cudaSuccess) { cudaGetLastError(); return false; }
  return cudaLaunchKernelEx(&cfg, kernel, tmdG, tmWb, tmdY, a) == cudaSuccess;
}

}  // namespace

// Shape policy of the persistent recurrences.  One launch must hold every CTA (co-resident, <= 148) and the barrier counters
// want a multiple of 8 CTAs per batch half:
//   128 < B <= 256 (Bp = 256): cta_group::2 pairs, one per tile, each CTA one batch half;
//   B <= 128       (Bp = 128): cta_group::1, one CTA per tile (small models: U fully resident in shared memory).
// Returns the tile width (gate columns / hidden units per tile), 0 = run the per-timestep kernels.
int fwd_recur_bn(int N, int Bp, int M) {
  if ((Bp != 256 && Bp != 128) || M != 256) return 0;
  static const int force = getenv("LSTM_FWD_RECUR") ? atoi(getenv("LSTM_FWD_RECUR")) : -1;
  if (force == 0) return 0;
  if (Bp == 256) {
    for (int bn : {128, 64}) {
      const int tiles = 4 * N / bn;
      if ((4 * N) % bn == 0 && tiles <= 74 && tiles % R_SLOTS == 0) return bn;
    }
    return 0;
  }
  for (int bn : {32, 64}) {                                  // one batch tile: narrow tiles, so that more SMs share the timestep
    const int tiles = 4 * N / bn;
    if ((4 * N) % bn == 0 && tiles <= 148 && tiles % R_SLOTS == 0) return bn;
  }
  return 0;
}
bool launch_fwd_recur(int bn, const CUtensorMap& tmH, const CUtensorMap& tmWb, const FwdRecurArgs& a, cudaStream_t st, bool dry) {
  if (a.Bp == 256) {
    if (bn == 128) return launch_fwd_recur_t<128, true>(tmH, tmWb, a, st, dry);
    if (bn == 64) return launch_fwd_recur_t<64, true>(tmH, tmWb, a, st, dry);
  } else if (a.Bp == 128) {
    if (bn == 64) return launch_fwd_recur_t<64, false>(tmH, tmWb, a, st, dry);
    if (bn == 32) return launch_fwd_recur_t<32, false>(tmH, tmWb, a, st, dry);
  }
  return false;
}
int fwd_recur_ctas(int bn, int N, int Bp) { return (Bp == 256 ? 2 : 1) * (4 * N / bn); }
int fwd_recur_box_rows(int bn, int Bp) { return Bp == 256 ? bn / 2 : bn; }

// BPTT: tile = BNJ hidden units, KS split-K ranks, each rank finalises BNJ/KS units.
//   Bp = 256: <128, 4> pairs (measured at N = 2048, profiles/r02d_*: 12.9 us per timestep against 16.8 us for <256, 8>, whose 8-way
//             exchange moves 14 MB through L2 per timestep instead of 6 MB); LSTM_BWD_RECUR=256 selects the latter for tests;
//   Bp = 128: <64, 4> or <32, 4> single CTAs, whichever yields 64..148 CTAs.
int bwd_recur_bnj(int N, int Bp, int M) {
  if ((Bp != 256 && Bp != 128) || M != 256) return 0;
  static const int force = getenv("LSTM_BWD_RECUR") ? atoi(getenv("LSTM_BWD_RECUR")) : -1;
  if (force == 0) return 0;
  const int nkbu = 4 * N / BK;
  auto fits = [&](int bnj, int ks) {
    const int ctas = (N / bnj) * ks;                         // per batch half
    return N % bnj == 0 && nkbu % ks == 0 && ctas <= (Bp == 256 ? 74 : 148) && ctas % R_SLOTS == 0;
  };
  if (Bp == 256) {
    if (force == 256) return fits(256, 8) ? 256 : 0;
    if (fits(128, 4)) return 128;
    if (fits(256, 8)) return 256;
    return 0;
  }
  if (fits(32, 4)) return 32;
  if (fits(64, 4)) return 64;
  return 0;
}
size_t bwd_recur_red_floats(int N, int bnj) {
  const int ks = bnj == 256 ? 8 : 4;
  return (size_t)(N / bnj) * 2 * ks * ks * (bnj / ks / 4) * 128 * 4;
}
int bwd_recur_box_rows(int bnj, int Bp) { return Bp == 256 ? bnj / 2 : bnj; }
int bwd_recur_ctas(int bnj, int N, int Bp) { return (Bp == 256 ? 2 : 1) * (N / bnj) * (bnj == 256 ? 8 : 4); }
int bwd_recur_per_slot(int bnj, int N, int Bp) { return (N / bnj) * (bnj == 256 ? 8 : 4) / R_SLOTS; }
bool launch_bwd_recur(int bnj, const CUtensorMap& tmdG, const CUtensorMap& tmWb, const CUtensorMap& tmdY, const BwdRecurArgs& a,
                      cudaStream_t st, bool dry) {
  if (a.Bp == 256) {
    if (bnj == 256) return launch_bwd_recur_t<256, 8, true>(tmdG, tmWb, tmdY, a, st, dry);
    if (bnj == 128) return launch_bwd_recur_t<128, 4, true>(tmdG, tmWb, tmdY, a, st, dry);
  } else if (a.Bp == 128) {
    if (bnj == 64) return launch_bwd_recur_t<64, 4, false>(tmdG, tmWb, tmdY, a, st, dry);
    if (bnj == 32) return launch_bwd_recur_t<32, 4, false>(tmdG, tmWb, tmdY, a, st, dry);
  }
  return false;
}

}  // namespace tc
