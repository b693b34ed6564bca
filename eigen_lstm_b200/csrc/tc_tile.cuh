// tc_tile.cuh — the shared skeleton of every tcgen05 kernel here: one 128 x BN output tile per CTA,
// warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5 = epilogue.
#pragma once
#include "tc_common.cuh"

namespace tc {

constexpr int BM = 128;                 // tile rows = TMEM lanes
constexpr int BK = 64;                  // bf16 per 128-byte swizzled row
constexpr int A_TILE_BYTES = BM * BK * 2;

struct KSeg {                           // one contiguous K range of the contraction
  const CUtensorMap* ta;
  const CUtensorMap* tb;
  int a_row, b_row, a_k0, b_k0, nkb;
};

template <int BN, int STAGES, int EPI_BYTES = 0>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int EPI = (EPI_BYTES + 127) / 128 * 128;   // epilogue staging region after the barriers
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + EPI;
  static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

struct TileCtx {
  long long* dbg;   // diagnostics (NULL = off): [1] last TMA issued, [2] first stage landed, [3] last MMA issued
  uint8_t* epi;     // epilogue staging region (Cfg::EPI bytes)
  uint8_t* tiles;
  uint64_t *full, *empty, *accum_full;
  uint32_t tmem_d;
  int warp, lane;
  uint64_t hint_a, hint_b;   // L2 eviction-priority hints of the A / B operand loads (0 = none)
};
__device__ __forceinline__ void tma_ld(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint64_t hint) {
  if (hint) tma_load_2d_hint(dst, tm, bar, c0, c1, hint);
  else tma_load_2d(dst, tm, bar, c0, c1);
}
__device__ __forceinline__ void tma_ld_pair(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint64_t hint) {
  if (hint) tma_load_2d_pair_hint(dst, tm, bar, c0, c1, hint);
  else tma_load_2d_pair(dst, tm, bar, c0, c1);
}

// Common prologue: carve shared memory, init barriers, allocate TMEM.
template <int BN, int STAGES>
__device__ __forceinline__ TileCtx tile_prologue(uint8_t* raw) {
  using C = Cfg<BN, STAGES>;
  TileCtx c;
  c.dbg = nullptr;
  c.hint_a = c.hint_b = 0;
  const uint32_t base = smem_u32(raw);
  const uint32_t pad = ((base + 1023u) & ~1023u) - base;   // SWIZZLE_128B atoms need 1024-byte alignment
  c.tiles = raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(c.tiles + C::TILE_BYTES);
  c.full = bars;
  c.empty = bars + STAGES;
  c.accum_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  c.epi = c.tiles + C::TILE_BYTES + 256;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], 1); }
    mbar_init(c.accum_full, 1);
    fence_barrier_init();
  }
  if (c.warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  c.tmem_d = *tmem_slot;
  return c;
}

template <int BN, int STAGES>
__device__ __forceinline__ void tile_epilogue_end(const TileCtx& c) {
  tcgen05_before_sync();
  __syncthreads();
  if (c.warp == 1) {
    __syncwarp();
    tmem_dealloc<Cfg<BN, STAGES>::TMEM_COLS>(c.tmem_d);
  }
}

// Producer (warp 0) and MMA issuer (warp 1) of one output tile.
// Both role loops run WARP-UNIFORM (all 32 lanes take the loop and the waits) with the single-thread instructions under
// elect.sync: issued from a divergent `if (lane == 0)` region, every UTMALDG / UTCHMMA / UTCBAR gets wrapped by ptxas in a
// vote-and-branch loop, and the MMA issue thread — not the tensor pipe or the operand delivery — became the bottleneck
// (~530 cycles per k-block regardless of tile size).
template <int BN, int STAGES>
__device__ __forceinline__ void tile_mainloop(const TileCtx& c, const KSeg& s0, const KSeg& s1) {
  using C = Cfg<BN, STAGES>;
  const int total = s0.nkb + s1.nkb;
  if (c.warp == 0) {
    for (int kb = 0; kb < total; kb++) {
      const int st = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&c.empty[st], ph ^ 1u);
      if (elect_one()) {
        const bool first = kb < s0.nkb;
        const KSeg& s = first ? s0 : s1;
        const int k = first ? kb : kb - s0.nkb;
        uint8_t* a = c.tiles + (size_t)st * C::STAGE_BYTES;
        uint8_t* b = a + A_TILE_BYTES;
        mbar_expect_tx(&c.full[st], (uint32_t)C::STAGE_BYTES);
        tma_ld(a, s.ta, &c.full[st], s.a_k0 + k * BK, s.a_row, c.hint_a);
        tma_ld(b, s.tb, &c.full[st], s.b_k0 + k * BK, s.b_row, c.hint_b);
      }
      __syncwarp();
    }
    if (c.dbg && c.lane == 0) c.dbg[1] = clock64();
  } else if (c.warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(c.tiles));
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(c.tiles) + A_TILE_BYTES);
    for (int kb = 0; kb < total; kb++) {
      const int st = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&c.full[st], ph);
      if (c.dbg && kb == 0 && c.lane == 0) c.dbg[2] = clock64();
      tcgen05_after_sync();
      if (elect_one()) {
        const uint64_t soff = (uint64_t)((uint32_t)st * (uint32_t)(C::STAGE_BYTES >> 4));   // descriptors address smem in 16-byte units
#pragma unroll
        for (int k = 0; k < BK / 16; k++)   // UMMA_K = 16 bf16 = 32 bytes inside the swizzled row
          umma_bf16(c.tmem_d, a_desc0 + soff + 2 * k, b_desc0 + soff + 2 * k, idesc, (uint32_t)((kb | k) != 0));
        umma_commit(&c.empty[st]);          // frees the stage once these MMAs have read it
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(c.accum_full);   // accumulator complete -> epilogue
    __syncwarp();
    if (c.dbg && c.lane == 0) c.dbg[3] = clock64();
  }
}

// ---- CTA-pair variant (tcgen05 cta_group::2): D = 256 rows x BN, the pair's two CTAs are the two 128-row halves ----
// Each CTA stages its own A tile (128 rows) and HALF of the B tile (BN/2 rows): 25 % fewer operand bytes per MMA than
// two independent 128 x BN tiles.  All loads of a stage complete on the leader's full barrier; the leader issues the
// M = 256 MMAs and its commit frees the stage in both CTAs.
template <int BN, int STAGES>
struct PairCfg {
  static constexpr int B_HALF_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_HALF_BYTES;
  static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
};

template <int BN, int STAGES>
__device__ __forceinline__ TileCtx pair_prologue(uint8_t* raw) {
  using C = PairCfg<BN, STAGES>;
  TileCtx c;
  c.dbg = nullptr;
  c.hint_a = c.hint_b = 0;
  const uint32_t base = smem_u32(raw);
  const uint32_t pad = ((base + 1023u) & ~1023u) - base;
  c.tiles = raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(c.tiles + C::TILE_BYTES);
  c.full = bars;
  c.empty = bars + STAGES;
  c.accum_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  c.epi = c.tiles + C::TILE_BYTES + 256;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&c.full[s], 1); mbar_init(&c.empty[s], 1); }
    mbar_init(c.accum_full, 1);
    fence_barrier_init();
  }
  if (c.warp == 1) tmem_alloc_pair<(BN < 32 ? 32 : BN)>(tmem_slot);
  tcgen05_before_sync();
  __syncthreads();
  cluster_sync_all();            // both CTAs' barriers exist before any remote completion / commit arrives
  tcgen05_after_sync();
  c.tmem_d = *tmem_slot;
  return c;
}

template <int BN, int STAGES>
__device__ __forceinline__ void pair_epilogue_end(const TileCtx& c) {
  tcgen05_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (c.warp == 1) {
    __syncwarp();
    tmem_dealloc_pair<(BN < 32 ? 32 : BN)>(c.tmem_d);
  }
}

// s.a_row = this CTA's own 128 A rows; s.b_row = first row of the pair's BN-row B tile (each CTA loads rows
// [b_row + rank*BN/2, +BN/2)).  rank = cluster rank within the pair (0 = leader).
template <int BN, int STAGES>
__device__ __forceinline__ void pair_mainloop(const TileCtx& c, const KSeg& s0, const KSeg& s1, uint32_t rank, uint16_t pair_mask) {
  using C = PairCfg<BN, STAGES>;
  const int total = s0.nkb + s1.nkb;
  if (c.warp == 0) {
    for (int kb = 0; kb < total; kb++) {
      const int st = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&c.empty[st], ph ^ 1u);
      if (elect_one()) {
        const bool first = kb < s0.nkb;
        const KSeg& s = first ? s0 : s1;
        const int k = first ? kb : kb - s0.nkb;
        uint8_t* a = c.tiles + (size_t)st * C::STAGE_BYTES;
        uint8_t* b = a + A_TILE_BYTES;
        if (rank == 0) mbar_expect_tx(&c.full[st], 2u * (uint32_t)C::STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
        tma_ld_pair(a, s.ta, &c.full[st], s.a_k0 + k * BK, s.a_row, c.hint_a);
        tma_ld_pair(b, s.tb, &c.full[st], s.b_k0 + k * BK, s.b_row + (int)rank * (BN / 2), c.hint_b);
      }
      __syncwarp();
    }
    if (c.dbg && c.lane == 0) c.dbg[1] = clock64();
  } else if (c.warp == 1 && rank == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(c.tiles));
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(c.tiles) + A_TILE_BYTES);
    for (int kb = 0; kb < total; kb++) {
      const int st = kb % STAGES;
      const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
      mbar_wait(&c.full[st], ph);
      if (c.dbg && kb == 0 && c.lane == 0) c.dbg[2] = clock64();
      tcgen05_after_sync();
      if (elect_one()) {
        const uint64_t soff = (uint64_t)((uint32_t)st * (uint32_t)(C::STAGE_BYTES >> 4));
#pragma unroll
        for (int k = 0; k < BK / 16; k++)
          umma_bf16_pair(c.tmem_d, a_desc0 + soff + 2 * k, b_desc0 + soff + 2 * k, idesc, (uint32_t)((kb | k) != 0));
        umma_commit_pair(&c.empty[st], pair_mask);     // stage free in both CTAs
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit_pair(c.accum_full, pair_mask);   // accumulator halves complete in both CTAs
    __syncwarp();
    if (c.dbg && c.lane == 0) c.dbg[3] = clock64();
  }
}

// Gate nonlinearities of the bf16 path: ONE MUFU operation each (tanh.approx.f32, max. relative error 2^-11 — a quarter of the
// rounding h(t) gets when it is stored as bf16).  The exp + reciprocal forms cost two MUFU operations per activation, and with
// five activations per (stream, unit) the 16 MUFU lanes of an SM bounded the epilogue of the timestep kernels (2.6 k cycles per
// 128 x 32 tile).  LSTM_F32 contexts use expf / tanhf like the reference (kernels_f32.cu).
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(tanh_fast(0.5f * x), 0.5f, 0.5f); }


template <typename K>
static void set_smem(K kernel, int bytes) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace tc
