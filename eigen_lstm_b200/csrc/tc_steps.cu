// tc_steps.cu — the two SERIAL kernels of the bf16 path: one forward timestep (K2) and one BPTT
// timestep (K5).  Both run the contraction on tcgen05 (tc_tile.cuh) and then re-map the 128 x BN
// accumulator tile through shared memory so that the LSTM math runs with LANE = HIDDEN UNIT: every
// global access of the epilogue (W row gather, gates stash, c, h, dg) is then a contiguous 128-512 B
// warp access instead of 32 scattered rows.
//
//   k_fwd_step  R/lstm.cc:176-192   g = W x + U h(t-1) + b, gates, c = tanh(i u + f c(t-1)), h = o c
//   k_bwd_step  R/lstm.cc:228-256   dh = Why^T dy + U^T dg(t+1), gate gradients, dcnext
//
// These are the GENERAL kernels (any B, N % 64 == 0), one launch per timestep, linked by programmatic dependent launch inside
// the iteration's CUDA graph.  Shapes with one pair of batch tiles and N <= 2048 run both recurrences as persistent kernels
// instead (tc_recur.cu).
//
// K5's output (B x N) is 4x smaller than K2's (B x 4N) for the same flops, so its K range
// (4N + M) is split over a cluster of 4 CTAs; the partial accumulators are reduce-scattered (each CTA
// finalises a quarter of the tile's hidden units).  The exchange goes through an L2-resident global
// scratch buffer, ordered by the hardware cluster barrier: the first version used distributed shared
// memory (st.shared::cluster) and measured 7.8 k cycles per step for it — DSMEM stores run at ~20 B/clk/SM,
// the L2 path at the SM's 64 B/clk fill rate.
#include <stdlib.h>

#include "tc_kernels.cuh"
#include "tc_tile.cuh"

namespace tc {

constexpr int EPI_WARPS = 16;            // the LSTM math is latency-bound: many warps, few rows each
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int CTA_THREADS = 64 + EPI_THREADS;
constexpr int HT_LD = 130;   // bf16 row pitch of the transposed staging tiles: 65 words -> conflict-free

// ------------------------------------------------------------------------------------------------
// K2: forward timestep.  D[b][r'] = sum_k h(t-1)[b][k] * U[r'][k], r' = 4*unit + gate.
// grid (4N/BN, Bp/128)
// ------------------------------------------------------------------------------------------------
template <int BN, bool PAIR = false>
struct FwdCfg {
  static constexpr int STAGES = PAIR ? (BN == 128 ? 5 : 8) : (BN == 128 ? 4 : (BN == 64 ? 6 : 8));
  static constexpr int UT = BN / 4;                        // hidden units per tile
  static constexpr int ACC_LD = BN + 4;                    // fp32 row pitch (16-byte aligned, odd multiple of 16 B)
  static constexpr int ACC_BYTES = 128 * ACC_LD * 4;
  static constexpr int HT_BYTES = UT * HT_LD * 2;
  static constexpr int X_BYTES = 128 * 4;
  static constexpr int EPI_BYTES = ACC_BYTES + HT_BYTES + X_BYTES;
  using C = Cfg<BN, STAGES, EPI_BYTES>;
  static constexpr int SMEM_BYTES = PAIR ? PairCfg<BN, STAGES>::TILE_BYTES + 1024 + 256 + (EPI_BYTES + 127) / 128 * 128 : C::SMEM_BYTES;
};

// PAIR: the two batch tiles of one gate-column tile form a cta_group::2 pair (cluster 2 x 1): one M = 256 MMA,
// each CTA stages its own h tile and half of the U tile.
template <int BN, bool PAIR = false>
__global__ void __launch_bounds__(CTA_THREADS, 1)
k_fwd_step(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmU, const FwdStepArgs a) {
  using F = FwdCfg<BN, PAIR>;
  constexpr int STAGES = F::STAGES, UT = F::UT, RG = EPI_THREADS / UT, ROWS = 128 / RG, ACC_LD = F::ACC_LD;
  extern __shared__ uint8_t smem_raw[];
  const long long t_entry = clock64();
  TileCtx c;
  if constexpr (PAIR) c = pair_prologue<BN, STAGES>(smem_raw);
  else c = tile_prologue<BN, STAGES>(smem_raw);
  pdl_launch_dependents();   // the next timestep's CTAs may take SMs as ours drain; they block in pdl_wait()
  pdl_wait();                // everything below reads what the previous timestep's kernel wrote
  c.dbg = (a.dbg && blockIdx.x == 0 && blockIdx.y == 0) ? a.dbg : nullptr;
  c.hint_b = L2_EVICT_LAST;  // U is re-read by every timestep: keep it in L2 ahead of the streamed stash
  const bool stamp = c.dbg && threadIdx.x == 64;
  if (stamp) { c.dbg[0] = t_entry; c.dbg[4] = clock64(); }
  float* acc = reinterpret_cast<float*>(c.epi);
  __nv_bfloat16* hT = reinterpret_cast<__nv_bfloat16*>(c.epi + F::ACC_BYTES);
  int* sx = reinterpret_cast<int*>(c.epi + F::ACC_BYTES + F::HT_BYTES);
  // PAIR: the two CTAs of a pair must be neighbours along cluster x: grid (2 * n_tiles, m_tiles / 2), cluster (2,1,1)
  const int nb = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int mb = PAIR ? (int)(blockIdx.y * 2 + (blockIdx.x & 1)) : (int)blockIdx.y;
  const KSeg s0{&tmH, &tmU, a.a_row0 + mb * BM, nb * BN, 0, 0, a.N / BK};
  const KSeg s1{&tmH, &tmU, 0, 0, 0, 0, 0};
  if constexpr (PAIR) pair_mainloop<BN, STAGES>(c, s0, s1, cluster_ctarank(), (uint16_t)0x3);
  else tile_mainloop<BN, STAGES>(c, s0, s1);
  if (c.warp >= 2) {
    const int e = threadIdx.x - 64;                        // 0..EPI_THREADS-1
    const int N = a.N, N4 = 4 * a.N;
    const int l = e % UT, rg = e / UT;                     // phase-2 mapping: lane = hidden unit
    const int j = nb * UT + l;                             // hidden unit
    const int rp = 4 * j;                                  // first of its 4 gate rows (unit-major order)
    if (e < 128) {                                         // input bytes of the tile's 128 streams
      const int b = mb * BM + e;
      sx[e] = (b < a.B) ? a.x[b] : -2;                     // -2 = padding row, -1 = all-zero input column
    }
    const float4 bias = *reinterpret_cast<const float4*>(a.bp + rp);
    named_bar_sync(1, EPI_THREADS);
    // Operands of the LSTM math (W row gather, c(t-1)) are loaded into registers NOW, while the tensor core is still
    // busy with the contraction: the phase-2 math then starts with its inputs already on chip.
    float4 w[ROWS];
    float cpv[ROWS];
    int xv[ROWS];
#pragma unroll
    for (int q = 0; q < ROWS; q++) {
      const int r = rg + RG * q;
      const int x = sx[r];
      xv[q] = x;
      w[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      cpv[q] = 0.f;
      if (x >= 0) w[q] = __ldg(reinterpret_cast<const float4*>(a.Wp + (size_t)x * N4 + rp));   // W*x, one-hot x
      if (x >= -1) cpv[q] = a.c_prev[(size_t)(mb * BM + r) * N + j];
    }
    // phase 1 (warps 2-5, one per TMEM lane quarter): TMEM (lane = stream) -> shared memory tile acc[row][col]
    if (c.warp < 6) {
      const int quarter = c.warp & 3;
      const int row = quarter * 32 + c.lane;
      mbar_wait(c.accum_full, 0);
      if (stamp) c.dbg[5] = clock64();
      tcgen05_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
        float4* dst = reinterpret_cast<float4*>(acc + (size_t)row * ACC_LD + c0);
#pragma unroll
        for (int u = 0; u < 8; u++) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
      }
    }
    named_bar_sync(1, EPI_THREADS);
    if (stamp) c.dbg[6] = clock64();
    // phase 2: lane = hidden unit; one warp instruction touches one stream's contiguous row segment
    {
      constexpr int RB = ROWS, i0 = 0;
#pragma unroll
      for (int q = 0; q < RB; q++) {
        const int r = rg + RG * (i0 + q);
        float hval = 0.f;
        if (xv[q] >= -1) {
          const int b = mb * BM + r;
          const float4 pre = *reinterpret_cast<const float4*>(acc + (size_t)r * ACC_LD + 4 * l);
          const float gi = sigmoid_fast(pre.x + w[q].x + bias.x);
          const float go = sigmoid_fast(pre.y + w[q].y + bias.y);
          const float gf = sigmoid_fast(pre.z + w[q].z + bias.z);
          const float gu = tanh_fast(pre.w + w[q].w + bias.w);
          const float cc = tanh_fast(gi * gu + gf * cpv[q]);   // the carried cell value is the tanh'd one
          hval = go * cc;
          __stcs(reinterpret_cast<float4*>(a.Gp_t + (size_t)b * N4 + rp), make_float4(gi, go, gf, gu));   // streamed: read once, in BPTT
          a.c_out[(size_t)b * N + j] = cc;
          a.Hbf_t[(size_t)b * N + j] = __float2bfloat16_rn(hval);
        }
        hT[l * HT_LD + r] = __float2bfloat16_rn(hval);
      }
    }
    named_bar_sync(1, EPI_THREADS);
    if (stamp) c.dbg[7] = clock64();
    // phase 3: h^T rows of ZT (the K6 operand), lanes along the stream index
    {
      const int w4 = e >> 5, lane = e & 31;
      for (int u = w4; u < UT; u += EPI_WARPS) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(hT + u * HT_LD);
        uint32_t* dst = reinterpret_cast<uint32_t*>(a.ZT_h + (size_t)(nb * UT + u) * a.ldz + mb * BM);
        __stcs(dst + lane, src[lane]);            // read again only by the K6 GEMM: do not displace U in L2
        __stcs(dst + lane + 32, src[lane + 32]);
      }
    }
  }
  if constexpr (PAIR) pair_epilogue_end<BN, STAGES>(c);
  else tile_epilogue_end<BN, STAGES>(c);
  if (stamp) c.dbg[8] = clock64();
}

// cluster launch with programmatic stream serialization (PDL): the successor's CTAs are resident and past their prologue when
// the predecessor grid drains
template <typename Kern, typename... Args>
static void launch_cluster(Kern kernel, dim3 grid, dim3 cluster, int smem, cudaStream_t st, Args... args) {
  set_smem(kernel, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(CTA_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster.x;
  attr[0].val.clusterDim.y = cluster.y;
  attr[0].val.clusterDim.z = cluster.z;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}

// K2 runs as cta_group::2 CTA pairs over the batch tiles when there is an even number of them: each CTA then stages only half
// of the weight tile, so the weight map needs a box of BN/2 rows.
bool step_pair(int Bp) { return (Bp / 128) % 2 == 0; }

template <int BN>
static void launch_fwd_t(const CUtensorMap& tmH, const CUtensorMap& tmUrk, const FwdStepArgs& a, cudaStream_t st) {
  dim3 grid(4 * a.N / BN, a.Bp / BM);
  if (step_pair(a.Bp))
    launch_cluster(k_fwd_step<BN, true>, dim3(2 * grid.x, grid.y / 2), dim3(2, 1, 1), FwdCfg<BN, true>::SMEM_BYTES, st, tmH, tmUrk, a);
  else
    launch_cluster(k_fwd_step<BN, false>, grid, dim3(1, 1, 1), FwdCfg<BN>::C::SMEM_BYTES, st, tmH, tmUrk, a);
}
// tmH must have a box of 128 rows and tmUrk one of BN rows (BN/2 when step_pair(Bp))
void launch_fwd_step(int BN, const CUtensorMap& tmH, const CUtensorMap& tmUrk, const FwdStepArgs& a, cudaStream_t st) {
  if (BN == 128) launch_fwd_t<128>(tmH, tmUrk, a, st);
  else if (BN == 64) launch_fwd_t<64>(tmH, tmUrk, a, st);
  else launch_fwd_t<32>(tmH, tmUrk, a, st);
}

// ------------------------------------------------------------------------------------------------
// K5: backward timestep.  D[b][j] = sum_r' dg(t+1)[b][r'] U[r'][j]  +  sum_m dy(t)[b][m] Why[m][j]
// (one contraction over the concatenated K range of 4N + M), split-K over a cluster of 4 CTAs.
// grid (N/BN, Bp/128, 4), cluster (1,1,4)
// ------------------------------------------------------------------------------------------------
constexpr int SPLIT = 4;
template <int BN>
struct BwdCfg {
  static constexpr int STAGES = BN == 128 ? 4 : (BN == 64 ? 6 : 8);
  static constexpr int UO = BN / SPLIT;                    // hidden units finalised by each CTA of the cluster
  static constexpr int RV_LD = UO + 4;                     // fp32 row pitch of this CTA's own partial slice
  static constexpr int RV_BYTES = 128 * RV_LD * 4;         // [row][UO]
  static constexpr int GT_BYTES = 4 * UO * HT_LD * 2;      // dg^T staging [gate*UO + unit][row]
  static constexpr int EPI_BYTES = RV_BYTES + GT_BYTES;
  using C = Cfg<BN, STAGES, EPI_BYTES>;
};

template <int BN>
__global__ void __launch_bounds__(CTA_THREADS, 1)
k_bwd_step(const __grid_constant__ CUtensorMap tmdG, const __grid_constant__ CUtensorMap tmU,
           const __grid_constant__ CUtensorMap tmdY, const __grid_constant__ CUtensorMap tmW, const BwdStepArgs a) {
  using F = BwdCfg<BN>;
  constexpr int STAGES = F::STAGES, UO = F::UO, RG = EPI_THREADS / UO, ROWS = 128 / RG, RV_LD = F::RV_LD;
  extern __shared__ uint8_t smem_raw[];
  const long long t_entry = clock64();
  TileCtx c = tile_prologue<BN, STAGES>(smem_raw);
  pdl_launch_dependents();
  pdl_wait();
  c.dbg = (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? a.dbg : nullptr;
  c.hint_b = L2_EVICT_LAST;  // U^T / Why^T are re-read by every timestep (measured: -0.9 us per step at N = 2048)
  const bool stamp = c.dbg && threadIdx.x == 64;
  if (stamp) { c.dbg[0] = t_entry; c.dbg[4] = clock64(); }
  float* recv = reinterpret_cast<float*>(c.epi);
  __nv_bfloat16* gT = reinterpret_cast<__nv_bfloat16*>(c.epi + F::RV_BYTES);
  const int nb = (int)blockIdx.x, mb = (int)blockIdx.y;
  const uint32_t rank = cluster_ctarank();                   // split-K rank
  float* red_tile = a.red + (size_t)(mb * (a.N / BN) + nb) * (SPLIT * SPLIT * 128 * UO);
  // this CTA's quarter of the concatenated K range [0, nkb0) ++ [0, nkb1)
  const int nkb0 = a.first ? 0 : (4 * a.N) / BK, nkb1 = a.M / BK;
  const int per = (nkb0 + nkb1) / SPLIT;
  const int lo = (int)rank * per, hi = lo + per;
  const int lo0 = min(lo, nkb0), hi0 = min(hi, nkb0);
  const int lo1 = max(lo, nkb0) - nkb0, hi1 = max(hi, nkb0) - nkb0;
  const KSeg s0{&tmdG, &tmU, a.dg_row0 + mb * BM, nb * BN, lo0 * BK, lo0 * BK, hi0 - lo0};
  const KSeg s1{&tmdY, &tmW, a.dy_row0 + mb * BM, nb * BN, lo1 * BK, lo1 * BK, hi1 - lo1};
  tile_mainloop<BN, STAGES>(c, s0, s1);
  const int e = threadIdx.x - 64;
  const int N = a.N, N4 = 4 * a.N;
  const int l = e >= 0 ? e % UO : 0, rg = e >= 0 ? e / UO : 0;
  const int j = nb * BN + (int)rank * UO + l;              // the hidden unit this thread finalises
  if (c.warp >= 2) {
#pragma unroll
    for (int i = 0; i < ROWS; i++) {                       // warm L2 for phase 2
      const int b = mb * BM + rg + RG * i;
      if (b < a.B) {
        if ((l & 7) == 0) prefetch_l2(a.Gp_t + (size_t)b * N4 + 4 * (size_t)j);
        if (l == 0) { prefetch_l2(a.c_t + (size_t)b * N + j); prefetch_l2(a.c_prev + (size_t)b * N + j); }
      }
    }
    // phase 1 (warps 2-5): reduce-scatter.  Column slice q of this CTA's partial accumulator goes to CTA q:
    // the own slice stays in shared memory, the other three go to red[tile][q][rank][row][UO] in global memory (L2)
    if (c.warp < 6) {
      const int quarter = c.warp & 3;
      const int row = quarter * 32 + c.lane;
      mbar_wait(c.accum_full, 0);
      if (stamp) c.dbg[5] = clock64();
      tcgen05_after_sync();
#pragma unroll 1
      for (int q = 0; q < SPLIT; q++) {
        float v[UO];
        tmem_ldw<UO>(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(q * UO), v);
        float4* dst = (q == (int)rank)
                          ? reinterpret_cast<float4*>(recv + (size_t)row * RV_LD)
                          : reinterpret_cast<float4*>(red_tile + (((size_t)q * SPLIT + rank) * 128 + row) * UO);
#pragma unroll
        for (int u = 0; u < UO / 4; u++) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
      }
    }
  }
  cluster_sync_all();                                      // all partial slices are written (cluster-scope release/acquire)
  if (stamp) c.dbg[6] = clock64();
  if (c.warp >= 2) {
    // phase 2: lane = hidden unit; batches of RB rows with all global loads issued up front (latency-bound phase)
    constexpr int RB = ROWS < 4 ? ROWS : 4;
#pragma unroll 1
    for (int i0 = 0; i0 < ROWS; i0 += RB) {
      float4 gv[RB];
      float ctv[RB], cpv[RB], dnv[RB], pv[RB][SPLIT];
#pragma unroll
      for (int q = 0; q < RB; q++) {
        const int r = rg + RG * (i0 + q);
        const int b = mb * BM + r;
        gv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        ctv[q] = cpv[q] = dnv[q] = 0.f;
#pragma unroll
        for (int sr = 0; sr < SPLIT; sr++) pv[q][sr] = 0.f;
        if (b < a.B) {
          const size_t bj = (size_t)b * N + j;
#pragma unroll
          for (int sr = 0; sr < SPLIT; sr++)
            if (sr != (int)rank) pv[q][sr] = __ldcg(red_tile + (((size_t)rank * SPLIT + sr) * 128 + r) * UO + l);
          gv[q] = __ldcs(reinterpret_cast<const float4*>(a.Gp_t + (size_t)b * N4 + 4 * (size_t)j));   // i o f u (last use)
          ctv[q] = a.c_t[bj];
          cpv[q] = a.c_prev[bj];
          if (!a.first) dnv[q] = a.dcnext[bj];
        }
      }
#pragma unroll
      for (int q = 0; q < RB; q++) {
        const int r = rg + RG * (i0 + q);
        const int b = mb * BM + r;
        float d_i = 0.f, d_o = 0.f, d_f = 0.f, d_u = 0.f;
        if (b < a.B) {
          float dh = 0.f;
#pragma unroll
          for (int sr = 0; sr < SPLIT; sr++) dh += (sr == (int)rank) ? recv[(size_t)r * RV_LD + l] : pv[q][sr];   // fixed order: deterministic
          const float4 g = gv[q];
          const float ct = ctv[q];
          const float dc = (dh * g.y + dnv[q]) * (1.0f - ct * ct);           // :233-235
          d_o = dh * ct * (g.y * (1.0f - g.y));                              // :238,244
          d_i = dc * g.w * (g.x * (1.0f - g.x));                             // :239,244
          d_f = dc * cpv[q] * (g.z * (1.0f - g.z));                          // :240,244
          d_u = dc * g.x * (1.0f - g.w * g.w);                               // :241,247
          a.dcnext[(size_t)b * N + j] = dc * g.z;                            // :256
          uint2 pk;
          pk.x = pack_bf16x2(d_i, d_o);
          pk.y = pack_bf16x2(d_f, d_u);
          *reinterpret_cast<uint2*>(a.dGbf_t + (size_t)b * N4 + 4 * (size_t)j) = pk;
        }
        gT[(0 * UO + l) * HT_LD + r] = __float2bfloat16_rn(d_i);
        gT[(1 * UO + l) * HT_LD + r] = __float2bfloat16_rn(d_o);
        gT[(2 * UO + l) * HT_LD + r] = __float2bfloat16_rn(d_f);
        gT[(3 * UO + l) * HT_LD + r] = __float2bfloat16_rn(d_u);
      }
    }
    named_bar_sync(1, EPI_THREADS);
    if (stamp) c.dbg[7] = clock64();
    // phase 3: dg^T rows (master row order gate*N + unit), lanes along the stream index
    {
      const int w4 = e >> 5, lane = e & 31;
      const int jbase = nb * BN + (int)rank * UO;
      for (int q = w4; q < 4 * UO; q += EPI_WARPS) {
        const int gate = q / UO, u = q - gate * UO;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(gT + q * HT_LD);
        uint32_t* dst = reinterpret_cast<uint32_t*>(a.dGT_t + (size_t)(gate * N + jbase + u) * a.ldg + mb * BM);
        __stcs(dst + lane, src[lane]);
        __stcs(dst + lane + 32, src[lane + 32]);
      }
    }
  }
  tile_epilogue_end<BN, STAGES>(c);
  if (stamp) c.dbg[8] = clock64();
}

template <int BN>
static void launch_bwd_t(const CUtensorMap& tmdG, const CUtensorMap& tmUkr, const CUtensorMap& tmdY,
                         const CUtensorMap& tmWnm, const BwdStepArgs& a, cudaStream_t st) {
  using F = BwdCfg<BN>;
  launch_cluster(k_bwd_step<BN>, dim3(a.N / BN, a.Bp / BM, SPLIT), dim3(1, 1, SPLIT), F::C::SMEM_BYTES, st, tmdG, tmUkr, tmdY, tmWnm, a);
}
void launch_bwd_step(int BN, const CUtensorMap& tmdG, const CUtensorMap& tmUkr, const CUtensorMap& tmdY,
                     const CUtensorMap& tmWnm, const BwdStepArgs& a, cudaStream_t st) {
  if (BN == 128) launch_bwd_t<128>(tmdG, tmUkr, tmdY, tmWnm, a, st);
  else if (BN == 64) launch_bwd_t<64>(tmdG, tmUkr, tmdY, tmWnm, a, st);
  else launch_bwd_t<32>(tmdG, tmUkr, tmdY, tmWnm, a, st);
}

}  // namespace tc
