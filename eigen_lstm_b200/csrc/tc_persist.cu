// tc_persist.cu — EXPERIMENTAL (off unless LSTM_PERSIST_FWD=1 / LSTM_PERSIST_BWD=1): the whole forward (and, further
// down, BPTT) recurrence of a window as ONE persistent kernel instead of T dependent launches.
//
// Why: with one launch per timestep a CTA is done with its own work ≈ 9.6 µs after
// griddepcontrol.wait returns, but the step costs 12.9 µs — ≈ 3.3 µs per boundary go to tail skew, grid-completion
// detection and the flush the wait implies, and neither PDL nor prefetching before the wait recovers them.  Here the
// 128 CTAs (64 gate-column tiles x 2 batch tiles, cta_group::2 pairs exactly as in k_fwd_step<.., PAIR>) stay resident
// for all T timesteps:
//   * warp 0 (producer) runs ahead: the U tiles of the next timestep's first STAGES k-blocks are in flight while the
//     epilogue of the current one is still running; only the h tiles wait for the grid barrier;
//   * h(t) is exchanged through global memory: the epilogue writes its [128 x BN/4] slice of Hbf slot t with generic
//     stores, fences them towards the async proxy (fence.proxy.async), and one thread per CTA release-adds to one of 8
//     arrival counters of its batch tile; the producers of the 64 CTAs that read this batch tile's h acquire-poll the
//     8 counters, fence again, and only then issue their TMA loads of slot t;
//   * the cell state c(t) of a (stream, unit) stays in the REGISTERS of the thread that owns it for the whole window
//     (it is still written to the Cs stash for BPTT, but never read back);
//   * the TMEM accumulator is reused every timestep: the leader's MMA warp waits on a pair-wide "accumulator drained"
//     mbarrier that both CTAs' epilogues arrive on after their tcgen05.ld.
// Every spin is bounded (trap after ~2 s) so a protocol error fails the launch instead of hanging the GPU.
//
// STATUS (end of round 1): runs on the B200 with a loss trajectory BITWISE identical to the launch-per-step kernels
// (config 4, profiles/r01f_bench_cfg4_persist_fwd_v*.json), i.e. the cross-proxy / cross-SM protocol is sound, but it is
// not yet faster: ~14.0 us per timestep against 12.9 us.  The grid barrier is not the reason (scripts/gridbar_bench.cu:
// 1.17 us per 128-CTA barrier, independent of the counter layout); the suspects are the 512 per-thread proxy fences and
// the MEMBAR inside the release (LSTM_PERSIST_WFENCE=0|2, scripts/persist_clocks.py).  See DESIGN.md sections 5 and 9.
#include <stdlib.h>

#include "tc_kernels.cuh"
#include "tc_tile.cuh"

namespace tc {

namespace {

constexpr int P_EPI_WARPS = 16;
constexpr int P_EPI_THREADS = P_EPI_WARPS * 32;
constexpr int P_CTA_THREADS = 64 + P_EPI_THREADS;
constexpr int P_HT_LD = 130;
constexpr int GBAR_SLOTS = 8;            // arrival counters per batch tile (spreads the same-address atomics)

template <int BN>
struct PersistCfg {
  static constexpr int STAGES = BN == 128 ? 5 : 8;
  static constexpr int UT = BN / 4;
  static constexpr int ACC_LD = BN + 4;
  static constexpr int ACC_BYTES = 128 * ACC_LD * 4;
  static constexpr int HT_BYTES = UT * P_HT_LD * 2;
  static constexpr int X_BYTES = 128 * 4;
  static constexpr int EPI_BYTES = ACC_BYTES + HT_BYTES + X_BYTES;
  static constexpr int SMEM_BYTES = PairCfg<BN, STAGES>::TILE_BYTES + 1024 + 256 + (EPI_BYTES + 127) / 128 * 128;
};

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
// generic-proxy <-> async-proxy ordering for GLOBAL memory only (FENCE.VIEW.ASYNC.G): the all-space form adds a
// MEMBAR.ALL.GPU per executing thread, which 512 epilogue threads would pay on the critical path
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// whole warp: wait until all GBAR_SLOTS counters of this batch tile have reached `target` (bounded)
__device__ __forceinline__ void grid_wait(const unsigned int* slots, int stride, unsigned int target, int lane) {
  const long long t0 = clock64();
  for (;;) {
    const unsigned int v = lane < GBAR_SLOTS ? ld_acquire_gpu(slots + (size_t)lane * stride) : target;
    if (__all_sync(0xffffffffu, (int)(v - target) >= 0)) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// grid (2 * n_tiles, m_tiles / 2), cluster (2,1,1): blockIdx.x & 1 = pair member = batch-tile parity
template <int BN>
__global__ void __launch_bounds__(P_CTA_THREADS, 1)
k_fwd_persist(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmU, const FwdPersistArgs a) {
  using F = PersistCfg<BN>;
  using PC = PairCfg<BN, F::STAGES>;
  constexpr int STAGES = F::STAGES, UT = F::UT, RG = P_EPI_THREADS / UT, ROWS = 128 / RG, ACC_LD = F::ACC_LD;
  extern __shared__ uint8_t smem_raw[];
  TileCtx c = pair_prologue<BN, STAGES>(smem_raw);
  uint64_t* tmem_free = c.accum_full + 2;                    // barrier slot after accum_full and the TMEM-address word
  if (threadIdx.x == 0) { mbar_init(tmem_free, 2); fence_barrier_init(); }
  __syncthreads();
  cluster_sync_all();
  const uint32_t rank = cluster_ctarank();
  const int nb = (int)(blockIdx.x >> 1);
  const int mb = (int)(blockIdx.y * 2 + (blockIdx.x & 1));
  const int n_tiles = (int)(gridDim.x >> 1);
  const int nkb = a.N / BK;
  const unsigned int per_slot = (unsigned int)(n_tiles / GBAR_SLOTS);   // arrivals per counter per timestep
  // bar_stride = 1: the 8 counters of a batch tile share one 32-byte sector (the variant verified on the B200);
  // bar_stride = 32 (LSTM_PERSIST_SPREAD=1, untested): one 128-byte line per counter, so the 64 pollers and 64 `red`s of a
  // batch tile spread over 8 L2 lines instead of serialising on one.
  const int bstride = a.bar_stride;
  unsigned int* my_slots = a.bar + (size_t)mb * GBAR_SLOTS * bstride;
  const int N = a.N, N4 = 4 * a.N, B = a.B;
  // diagnostics (LSTM_TC_DEBUG=1): SM-clock stamps of CTA (0,0) around timestep DBG_T, read with lstm_debug_kernel_clocks:
  // [0] producer reaches the grid barrier  [1] barrier passed  [2] first operand stage landed  [3] last MMA issued
  // [4] accumulator complete  [5] accumulator in shared memory  [6] h(t) announced  [7] timestep's stores done
  // [8] producer reaches the NEXT timestep's grid barrier ([8] - [0] = one timestep period)
  constexpr int DBG_T = 4;
  long long* dbg = (a.dbg && blockIdx.x == 0 && blockIdx.y == 0 && a.T > DBG_T) ? a.dbg : nullptr;

  if (c.warp == 0) {
    // ---------------- producer ----------------
    if (elect_one()) { tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmU); }
    __syncwarp();
    const int pre = nkb < STAGES ? nkb : STAGES;
    int g = 0;                                               // k-blocks issued so far (ring position)
    for (int t = 1; t <= a.T; t++) {
      const int a_row = (t - 1) * a.Bp + mb * BM;
      for (int kb = 0; kb < pre; kb++) {                     // weights first: they do not depend on h(t-1)
        const int st = (g + kb) % STAGES;
        const uint32_t ph = (uint32_t)((g + kb) / STAGES) & 1u;
        mbar_wait(&c.empty[st], ph ^ 1u);
        if (elect_one()) {
          uint8_t* bdst = c.tiles + (size_t)st * PC::STAGE_BYTES + A_TILE_BYTES;
          if (rank == 0) mbar_expect_tx(&c.full[st], 2u * (uint32_t)PC::STAGE_BYTES);
          tma_load_2d_pair(bdst, &tmU, &c.full[st], kb * BK, nb * BN + (int)rank * (BN / 2));
        }
        __syncwarp();
      }
      if (t > 1) {                                           // h(t-1) of this batch tile is complete in global memory
        if (dbg && c.lane == 0 && (t == DBG_T || t == DBG_T + 1)) dbg[t == DBG_T ? 0 : 8] = clock64();
        grid_wait(my_slots, bstride, (unsigned int)(t - 1) * per_slot, c.lane);
        fence_proxy_async_global();                          // generic-proxy writes (observed via acquire) -> async-proxy reads
        if (dbg && c.lane == 0 && t == DBG_T) dbg[1] = clock64();
      }
      for (int kb = 0; kb < nkb; kb++) {
        const int st = (g + kb) % STAGES;
        const uint32_t ph = (uint32_t)((g + kb) / STAGES) & 1u;
        if (kb >= pre) mbar_wait(&c.empty[st], ph ^ 1u);
        if (elect_one()) {
          uint8_t* adst = c.tiles + (size_t)st * PC::STAGE_BYTES;
          if (kb >= pre) {
            if (rank == 0) mbar_expect_tx(&c.full[st], 2u * (uint32_t)PC::STAGE_BYTES);
            tma_load_2d_pair(adst + A_TILE_BYTES, &tmU, &c.full[st], kb * BK, nb * BN + (int)rank * (BN / 2));
          }
          tma_load_2d_pair(adst, &tmH, &c.full[st], kb * BK, a_row);
        }
        __syncwarp();
      }
      g += nkb;
    }
  } else if (c.warp == 1) {
    // ---------------- MMA issuer (leader CTA of the pair) ----------------
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(c.tiles));
      const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(c.tiles) + A_TILE_BYTES);
      int g = 0;
      for (int t = 1; t <= a.T; t++) {
        if (t > 1) {                                         // both CTAs have read timestep t-1's accumulator out of TMEM
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
            const uint64_t soff = (uint64_t)((uint32_t)st * (uint32_t)(PC::STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BK / 16; k++)
              umma_bf16_pair(c.tmem_d, a_desc0 + soff + 2 * k, b_desc0 + soff + 2 * k, idesc, (uint32_t)((kb | k) != 0));
            umma_commit_pair(&c.empty[st], (uint16_t)0x3);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit_pair(c.accum_full, (uint16_t)0x3);
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
    const int l = e % UT, rg = e / UT;
    const int j = nb * UT + l;
    const int rp = 4 * j;
    const float4 bias = *reinterpret_cast<const float4*>(a.bp + rp);
    float cpv[ROWS];                                         // c(t-1) of this thread's (stream, unit) pairs: register-resident
#pragma unroll
    for (int q = 0; q < ROWS; q++) {
      const int b = mb * BM + rg + RG * q;
      cpv[q] = b < B ? a.Cs[(size_t)b * N + j] : 0.f;        // slot 0 = carried-in state
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
      named_bar_sync(1, P_EPI_THREADS);
      float4 w[ROWS];
      int xv[ROWS];
#pragma unroll
      for (int q = 0; q < ROWS; q++) {                       // W*x for one-hot x = a row gather; in flight during the contraction
        const int x = sx[rg + RG * q];
        xv[q] = x;
        w[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x >= 0) w[q] = __ldg(reinterpret_cast<const float4*>(a.Wp + (size_t)x * N4 + rp));
      }
      if (c.warp < 6) {                                      // TMEM (lane = stream) -> shared memory tile
        const int quarter = c.warp & 3;
        const int row = quarter * 32 + c.lane;
        mbar_wait(c.accum_full, (uint32_t)(t - 1) & 1u);
        if (dbg && e == 0 && t == DBG_T) dbg[4] = clock64();
        tcgen05_after_sync();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          float v[32];
          tmem_ld32(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
          float4* dst = reinterpret_cast<float4*>(acc + (size_t)row * ACC_LD + c0);
#pragma unroll
          for (int u = 0; u < 8; u++) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
        }
        tcgen05_before_sync();
      }
      named_bar_sync(1, P_EPI_THREADS);
      if (dbg && e == 0 && t == DBG_T) dbg[5] = clock64();
      if (e == 0) mbar_arrive_remote(tmem_free, 0);          // this CTA's half of the accumulator is drained
      // Only h(t) is on the critical path of the other CTAs: it is stored FIRST, fenced and announced; the gate stash and
      // c(t) (registers until then) follow after the arrival, so the release does not wait for their 10x larger traffic.
#pragma unroll
      for (int q = 0; q < ROWS; q++) {
        const int r = rg + RG * q;
        float hval = 0.f;
        if (xv[q] >= -1) {
          const int b = mb * BM + r;
          const float4 pre = *reinterpret_cast<const float4*>(acc + (size_t)r * ACC_LD + 4 * l);
          const float gi = sigmoid_fast(pre.x + w[q].x + bias.x);
          const float go = sigmoid_fast(pre.y + w[q].y + bias.y);
          const float gf = sigmoid_fast(pre.z + w[q].z + bias.z);
          const float gu = tanh_fast(pre.w + w[q].w + bias.w);
          const float cc = tanh_fast(gi * gu + gf * cpv[q]);   // the carried cell value is the tanh'd one (R/lstm.cc:185-189)
          hval = go * cc;
          cpv[q] = cc;
          w[q] = make_float4(gi, go, gf, gu);                  // the W row is dead: its registers carry the activated gates
          Hbf_t[(size_t)b * N + j] = __float2bfloat16_rn(hval);
        }
        hT[l * P_HT_LD + r] = __float2bfloat16_rn(hval);
      }
      // writer-side proxy fence: 1 (default, the verified variant) = every writing thread; 2 = only the announcing thread,
      // after the barrier; 0 = none (the readers fence after their acquire in any case).  LSTM_PERSIST_WFENCE selects.
      if (a.writer_fence == 1) fence_proxy_async_global();
      named_bar_sync(1, P_EPI_THREADS);
      if (a.writer_fence == 2 && e == 0) fence_proxy_async_global();
      if (e == 0) red_release_gpu_add(my_slots + (size_t)(nb % GBAR_SLOTS) * bstride, 1u);   // release: cumulative over the barrier-ordered stores
      if (dbg && e == 0 && t == DBG_T) dbg[6] = clock64();
#pragma unroll
      for (int q = 0; q < ROWS; q++) {
        if (xv[q] >= -1) {
          const int b = mb * BM + rg + RG * q;
          __stcs(reinterpret_cast<float4*>(Gp_t + (size_t)b * N4 + rp), w[q]);   // streamed: read once, in BPTT
          c_out[(size_t)b * N + j] = cpv[q];
        }
      }
      {                                                      // h^T rows of ZT (K6 operand): off the critical path
        const int w4 = e >> 5, lane = e & 31;
        for (int u = w4; u < UT; u += P_EPI_WARPS) {
          const uint32_t* src = reinterpret_cast<const uint32_t*>(hT + u * P_HT_LD);
          uint32_t* dst = reinterpret_cast<uint32_t*>(ZT_h + (size_t)(nb * UT + u) * a.ldz + mb * BM);
          __stcs(dst + lane, src[lane]);
          __stcs(dst + lane + 32, src[lane + 32]);
        }
      }
      if (dbg && e == 0 && t == DBG_T) dbg[7] = clock64();
    }
  }
  pair_epilogue_end<BN, STAGES>(c);
}

template <int BN>
bool launch_persist_t(const CUtensorMap& tmH, const CUtensorMap& tmUrk, const FwdPersistArgs& a, cudaStream_t st) {
  using F = PersistCfg<BN>;
  const int n_tiles = 4 * a.N / BN, m_tiles = a.Bp / BM;
  if (n_tiles % GBAR_SLOTS != 0 || m_tiles % 2 != 0) return false;
  auto kernel = k_fwd_persist<BN>;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F::SMEM_BYTES) != cudaSuccess) { cudaGetLastError(); return false; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * n_tiles, m_tiles / 2, 1);
  cfg.blockDim = dim3(P_CTA_THREADS);
  cfg.dynamicSmemBytes = (size_t)F::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // every CTA must be resident at once (the grid barrier spins): refuse unless all clusters fit on an idle GPU
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
  if (max_clusters < n_tiles * m_tiles / 2) return false;
  if (cudaMemsetAsync(a.bar, 0, (size_t)m_tiles * GBAR_SLOTS * a.bar_stride * sizeof(unsigned int), st) != cudaSuccess) { cudaGetLastError(); return false; }
  return cudaLaunchKernelEx(&cfg, kernel, tmH, tmUrk, a) == cudaSuccess;
}


// ------------------------------------------------------------------------------------------------------------------
// Persistent BPTT recurrence (LSTM_PERSIST_BWD=1; written at the end of round 1, compiled, NOT yet run).
// The per-timestep kernel k_bwd_step (tc_steps.cu) in a loop over t = T .. 1, non-pair variant: grid (N/BN, Bp/128, 4),
// cluster (1,1,4) = the four split-K ranks of one output tile.  What changes against one launch per timestep:
//   * the grid barrier on dg(t+1) replaces the kernel boundary (same counters and fences as k_fwd_persist: every CTA
//     finalises [128 streams x BN/4 units x 4 gates] of dGbf slot t, 64 CTAs per batch tile);
//   * the weight tiles (U^T, Why^T) of the next timestep's first STAGES k-blocks are prefetched during the epilogue;
//   * the split-K exchange is ordered by a shared-memory mbarrier in every CTA of the cluster that the four ranks arrive
//     on remotely (mapa + mbarrier.arrive.release.cluster) instead of barrier.cluster, which would need the producer and
//     MMA warps — busy with the next timestep — to take part;
//   * dcnext lives in the registers of the thread that owns the (stream, unit) pair for the whole window.
// Reuse of the exchange scratch across timesteps is safe: a CTA writes its slices of timestep t-1 only after its MMAs of
// that timestep, which wait for the grid barrier, which every CTA passes only after it has read the slices of timestep t.
template <int BN>
struct PersistBwdCfg {
  static constexpr int STAGES = BN == 128 ? 4 : (BN == 64 ? 6 : 8);
  static constexpr int UO = BN / 4;
  static constexpr int RV_LD = UO + 4;
  static constexpr int RV_BYTES = 128 * RV_LD * 4;
  static constexpr int GT_BYTES = 4 * UO * P_HT_LD * 2;
  static constexpr int EPI_BYTES = RV_BYTES + GT_BYTES;
  using C = Cfg<BN, STAGES, EPI_BYTES>;
};

template <int BN>
__global__ void __launch_bounds__(P_CTA_THREADS, 1)
k_bwd_persist(const __grid_constant__ CUtensorMap tmdG, const __grid_constant__ CUtensorMap tmU,
              const __grid_constant__ CUtensorMap tmdY, const __grid_constant__ CUtensorMap tmW, const BwdPersistArgs a) {
  using F = PersistBwdCfg<BN>;
  using C = typename F::C;
  constexpr int STAGES = F::STAGES, UO = F::UO, RG = P_EPI_THREADS / UO, ROWS = 128 / RG, RV_LD = F::RV_LD, SPLITK = 4;
  extern __shared__ uint8_t smem_raw[];
  TileCtx c = tile_prologue<BN, STAGES>(smem_raw);
  uint64_t* tmem_free = c.accum_full + 2;                    // this CTA's accumulator has been read out of TMEM
  uint64_t* xch = c.accum_full + 3;                          // the four ranks' partial slices of this timestep are written
  if (threadIdx.x == 0) { mbar_init(tmem_free, 1); mbar_init(xch, SPLITK); fence_barrier_init(); }
  __syncthreads();
  cluster_sync_all();                                        // every rank's xch barrier exists before anyone arrives on it
  const int nb = (int)blockIdx.x, mb = (int)blockIdx.y;
  const uint32_t rank = cluster_ctarank();                   // split-K rank = blockIdx.z
  const int N = a.N, N4 = 4 * a.N, B = a.B;
  const int n_tiles = (int)gridDim.x;
  const unsigned int per_slot = (unsigned int)(n_tiles * SPLITK / GBAR_SLOTS);   // arrivals per counter per timestep
  const int bstride = a.bar_stride;
  unsigned int* my_slots = a.bar + (size_t)mb * GBAR_SLOTS * bstride;
  float* red_tile = a.red + (size_t)(mb * n_tiles + nb) * (SPLITK * SPLITK * 128 * UO);

  // this CTA's quarter of the concatenated K range [0, nkb0) ++ [0, nkb1) of timestep t (nkb0 = 0 for t == T)
  auto ksegs = [&](int t, KSeg& s0, KSeg& s1) {
    const int nkb0 = (t == a.T) ? 0 : N4 / BK, nkb1 = a.M / BK;
    const int per = (nkb0 + nkb1) / SPLITK;
    const int lo = (int)rank * per, hi = lo + per;
    const int lo0 = min(lo, nkb0), hi0 = min(hi, nkb0);
    const int lo1 = max(lo, nkb0) - nkb0, hi1 = max(hi, nkb0) - nkb0;
    s0 = KSeg{&tmdG, &tmU, t * a.Bp + mb * BM, nb * BN, lo0 * BK, lo0 * BK, hi0 - lo0};
    s1 = KSeg{&tmdY, &tmW, (t - 1) * a.Bp + mb * BM, nb * BN, lo1 * BK, lo1 * BK, hi1 - lo1};
  };

  if (c.warp == 0) {
    // ---------------- producer ----------------
    if (elect_one()) { tma_prefetch_desc(&tmdG); tma_prefetch_desc(&tmU); tma_prefetch_desc(&tmdY); tma_prefetch_desc(&tmW); }
    __syncwarp();
    int g = 0;
    for (int t = a.T; t >= 1; t--) {
      KSeg s0, s1;
      ksegs(t, s0, s1);
      const int total = s0.nkb + s1.nkb;
      const int pre = total < STAGES ? total : STAGES;
      for (int kb = 0; kb < pre; kb++) {                     // weight tiles first
        const int st = (g + kb) % STAGES;
        const uint32_t ph = (uint32_t)((g + kb) / STAGES) & 1u;
        mbar_wait(&c.empty[st], ph ^ 1u);
        if (elect_one()) {
          const bool first = kb < s0.nkb;
          const KSeg& s = first ? s0 : s1;
          const int k = first ? kb : kb - s0.nkb;
          mbar_expect_tx(&c.full[st], (uint32_t)C::STAGE_BYTES);
          tma_load_2d(c.tiles + (size_t)st * C::STAGE_BYTES + A_TILE_BYTES, s.tb, &c.full[st], s.b_k0 + k * BK, s.b_row);
        }
        __syncwarp();
      }
      if (t < a.T) {                                         // dg(t+1) of this batch tile is complete in global memory
        grid_wait(my_slots, bstride, (unsigned int)(a.T - t) * per_slot, c.lane);
        fence_proxy_async_global();
      }
      for (int kb = 0; kb < total; kb++) {
        const int st = (g + kb) % STAGES;
        const uint32_t ph = (uint32_t)((g + kb) / STAGES) & 1u;
        if (kb >= pre) mbar_wait(&c.empty[st], ph ^ 1u);
        if (elect_one()) {
          const bool first = kb < s0.nkb;
          const KSeg& s = first ? s0 : s1;
          const int k = first ? kb : kb - s0.nkb;
          uint8_t* adst = c.tiles + (size_t)st * C::STAGE_BYTES;
          if (kb >= pre) {
            mbar_expect_tx(&c.full[st], (uint32_t)C::STAGE_BYTES);
            tma_load_2d(adst + A_TILE_BYTES, s.tb, &c.full[st], s.b_k0 + k * BK, s.b_row);
          }
          tma_load_2d(adst, s.ta, &c.full[st], s.a_k0 + k * BK, s.a_row);
        }
        __syncwarp();
      }
      g += total;
    }
  } else if (c.warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(c.tiles));
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(c.tiles) + A_TILE_BYTES);
    int g = 0;
    for (int t = a.T; t >= 1; t--) {
      KSeg s0, s1;
      ksegs(t, s0, s1);
      const int total = s0.nkb + s1.nkb;
      if (t < a.T) {
        mbar_wait(tmem_free, (uint32_t)(a.T - t - 1) & 1u);
        tcgen05_after_sync();
      }
      for (int kb = 0; kb < total; kb++) {
        const int st = (g + kb) % STAGES;
        const uint32_t ph = (uint32_t)((g + kb) / STAGES) & 1u;
        mbar_wait(&c.full[st], ph);
        tcgen05_after_sync();
        if (elect_one()) {
          const uint64_t soff = (uint64_t)((uint32_t)st * (uint32_t)(C::STAGE_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BK / 16; k++)
            umma_bf16(c.tmem_d, a_desc0 + soff + 2 * k, b_desc0 + soff + 2 * k, idesc, (uint32_t)((kb | k) != 0));
          umma_commit(&c.empty[st]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(c.accum_full);
      __syncwarp();
      g += total;
    }
  } else {
    // ---------------- epilogue: split-K reduce + gate gradients, lane = hidden unit ----------------
    float* recv = reinterpret_cast<float*>(c.epi);
    __nv_bfloat16* gT = reinterpret_cast<__nv_bfloat16*>(c.epi + F::RV_BYTES);
    const int e = threadIdx.x - 64;
    const int l = e % UO, rg = e / UO;
    const int j = nb * BN + (int)rank * UO + l;              // the hidden unit this thread finalises
    float dcn[ROWS];                                         // dcnext of this thread's (stream, unit) pairs: register-resident
#pragma unroll
    for (int q = 0; q < ROWS; q++) dcn[q] = 0.f;
    for (int t = a.T; t >= 1; t--) {
      const int step = a.T - t;                              // 0-based timestep counter of this launch
      const float* Gp_t = a.Gp + (size_t)(t - 1) * B * N4;
      const float* c_t = a.Cs + (size_t)t * B * N;
      const float* c_prev = a.Cs + (size_t)(t - 1) * B * N;
      __nv_bfloat16* dGbf_t = a.dGbf + (size_t)(t - 1) * a.Bp * N4;
      __nv_bfloat16* dGT_t = a.dGT + (size_t)(t - 1) * a.Bp;
      // phase 1 (warps 2-5): reduce-scatter of the partial accumulator (own slice -> smem, foreign slices -> L2 scratch)
      if (c.warp < 6) {
        const int quarter = c.warp & 3;
        const int row = quarter * 32 + c.lane;
        mbar_wait(c.accum_full, (uint32_t)step & 1u);
        tcgen05_after_sync();
#pragma unroll 1
        for (int q = 0; q < SPLITK; q++) {
          float v[UO];
          tmem_ldw<UO>(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(q * UO), v);
          float4* dst = (q == (int)rank)
                            ? reinterpret_cast<float4*>(recv + (size_t)row * RV_LD)
                            : reinterpret_cast<float4*>(red_tile + (((size_t)q * SPLITK + rank) * 128 + row) * UO);
#pragma unroll
          for (int u = 0; u < UO / 4; u++) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
        }
        tcgen05_before_sync();
        named_bar_sync(2, 128);                              // the four writer warps
        if (e == 0) {
          mbar_arrive(tmem_free);                            // accumulator drained: the next timestep's MMAs may overwrite it
#pragma unroll
          for (uint32_t q = 0; q < (uint32_t)SPLITK; q++) mbar_arrive_remote(xch, q);   // release.cluster: our slices are written
        }
      }
      mbar_wait_cluster(xch, (uint32_t)step & 1u);           // all four ranks' slices are visible (acquire.cluster)
      // phase 2: lane = hidden unit; batches of RB rows with all global loads issued up front (the phase is latency-bound,
      // and the stores of one row must not serialise the loads of the next behind a possible alias)
      constexpr int RB = ROWS < 4 ? ROWS : 4;
#pragma unroll
      for (int i0 = 0; i0 < ROWS; i0 += RB) {
        float4 gv[RB];
        float ctv[RB], cpv[RB], pv[RB][SPLITK];
#pragma unroll
        for (int q = 0; q < RB; q++) {
          const int r = rg + RG * (i0 + q);
          const int b = mb * BM + r;
          gv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          ctv[q] = cpv[q] = 0.f;
#pragma unroll
          for (int sr = 0; sr < SPLITK; sr++) pv[q][sr] = 0.f;
          if (b < B) {
            const size_t bj = (size_t)b * N + j;
#pragma unroll
            for (int sr = 0; sr < SPLITK; sr++)
              if (sr != (int)rank) pv[q][sr] = __ldcg(red_tile + (((size_t)rank * SPLITK + sr) * 128 + r) * UO + l);
            gv[q] = __ldcs(reinterpret_cast<const float4*>(Gp_t + (size_t)b * N4 + 4 * (size_t)j));   // i o f u (last use)
            ctv[q] = c_t[bj];
            cpv[q] = c_prev[bj];
          }
        }
#pragma unroll
        for (int q = 0; q < RB; q++) {
          const int r = rg + RG * (i0 + q);
          const int b = mb * BM + r;
          float d_i = 0.f, d_o = 0.f, d_f = 0.f, d_u = 0.f;
          if (b < B) {
            float dh = 0.f;
#pragma unroll
            for (int sr = 0; sr < SPLITK; sr++) dh += (sr == (int)rank) ? recv[(size_t)r * RV_LD + l] : pv[q][sr];   // fixed order: deterministic
            const float4 gg = gv[q];
            const float ct = ctv[q];
            const float dc = (dh * gg.y + dcn[i0 + q]) * (1.0f - ct * ct);       // R/lstm.cc:233-235
            d_o = dh * ct * (gg.y * (1.0f - gg.y));                              // :238,244
            d_i = dc * gg.w * (gg.x * (1.0f - gg.x));                            // :239,244
            d_f = dc * cpv[q] * (gg.z * (1.0f - gg.z));                          // :240,244
            d_u = dc * gg.x * (1.0f - gg.w * gg.w);                              // :241,247
            dcn[i0 + q] = dc * gg.z;                                             // :256
            uint2 pk;
            pk.x = pack_bf16x2(d_i, d_o);
            pk.y = pack_bf16x2(d_f, d_u);
            *reinterpret_cast<uint2*>(dGbf_t + (size_t)b * N4 + 4 * (size_t)j) = pk;
          }
          gT[(0 * UO + l) * P_HT_LD + r] = __float2bfloat16_rn(d_i);
          gT[(1 * UO + l) * P_HT_LD + r] = __float2bfloat16_rn(d_o);
          gT[(2 * UO + l) * P_HT_LD + r] = __float2bfloat16_rn(d_f);
          gT[(3 * UO + l) * P_HT_LD + r] = __float2bfloat16_rn(d_u);
        }
      }
      if (a.writer_fence == 1) fence_proxy_async_global();
      named_bar_sync(1, P_EPI_THREADS);
      if (a.writer_fence == 2 && e == 0) fence_proxy_async_global();
      if (e == 0) red_release_gpu_add(my_slots + (size_t)((nb * SPLITK + (int)rank) % GBAR_SLOTS) * bstride, 1u);
      {                                                      // dg^T rows (master row order gate*N + unit) for K6
        const int w4 = e >> 5, lane = e & 31;
        const int jbase = nb * BN + (int)rank * UO;
        for (int q = w4; q < 4 * UO; q += P_EPI_WARPS) {
          const int gate = q / UO, u = q - gate * UO;
          const uint32_t* src = reinterpret_cast<const uint32_t*>(gT + q * P_HT_LD);
          uint32_t* dst = reinterpret_cast<uint32_t*>(dGT_t + (size_t)(gate * N + jbase + u) * a.ldg + mb * BM);
          __stcs(dst + lane, src[lane]);
          __stcs(dst + lane + 32, src[lane + 32]);
        }
      }
      named_bar_sync(1, P_EPI_THREADS);                      // gT and recv are rewritten by the next timestep
    }
  }
  tcgen05_before_sync();
  __syncthreads();
  cluster_sync_all();                                        // no rank may still arrive on an exited CTA's barrier
  if (c.warp == 1) {
    __syncwarp();
    tmem_dealloc<C::TMEM_COLS>(c.tmem_d);
  }
}

template <int BN>
bool launch_persist_bwd_t(const CUtensorMap& tmdG, const CUtensorMap& tmUkr, const CUtensorMap& tmdY, const CUtensorMap& tmWnm,
                          const BwdPersistArgs& a, cudaStream_t st) {
  using F = PersistBwdCfg<BN>;
  const int n_tiles = a.N / BN, m_tiles = a.Bp / BM;
  if ((n_tiles * 4) % GBAR_SLOTS != 0 || ((4 * a.N / BK) + a.M / BK) % 4 != 0 || (a.M / BK) % 4 != 0) return false;
  auto kernel = k_bwd_persist<BN>;
  const int smem = F::C::SMEM_BYTES;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { cudaGetLastError(); return false; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_tiles, m_tiles, 4);
  cfg.blockDim = dim3(P_CTA_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 4;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
  if (max_clusters < n_tiles * m_tiles) return false;        // every CTA must be resident: the grid barrier spins
  if (cudaMemsetAsync(a.bar, 0, (size_t)m_tiles * GBAR_SLOTS * a.bar_stride * sizeof(unsigned int), st) != cudaSuccess) { cudaGetLastError(); return false; }
  return cudaLaunchKernelEx(&cfg, kernel, tmdG, tmUkr, tmdY, tmWnm, a) == cudaSuccess;
}

}  // namespace

bool fwd_persist_enabled() {
  static const bool on = getenv("LSTM_PERSIST_FWD") != nullptr && atoi(getenv("LSTM_PERSIST_FWD")) != 0;
  return on;
}

// Returns false when the shape cannot run persistently (the caller then launches one kernel per timestep).
// tmH must have a 128-row box and tmUrk a BN/2-row box — the maps the pair variant of k_fwd_step uses.
bool launch_fwd_persist(int BN, const CUtensorMap& tmH, const CUtensorMap& tmUrk, const FwdPersistArgs& a0, cudaStream_t st) {
  if (!step_pair(a0.Bp)) return false;
  static const bool spread = getenv("LSTM_PERSIST_SPREAD") != nullptr && atoi(getenv("LSTM_PERSIST_SPREAD")) != 0;
  FwdPersistArgs a = a0;
  a.bar_stride = spread ? 32 : 1;                            // the counter buffer holds [Bp/128][8][32] words either way
  static const int wfence = getenv("LSTM_PERSIST_WFENCE") ? atoi(getenv("LSTM_PERSIST_WFENCE")) : 1;
  a.writer_fence = wfence;
  if (BN == 128) return launch_persist_t<128>(tmH, tmUrk, a, st);
  if (BN == 64) return launch_persist_t<64>(tmH, tmUrk, a, st);
  return false;
}

bool bwd_persist_enabled() {
  static const bool on = getenv("LSTM_PERSIST_BWD") != nullptr && atoi(getenv("LSTM_PERSIST_BWD")) != 0;
  return on;
}

// Returns false when the shape cannot run persistently.  tmUkr / tmWnm must have BN-row boxes (the non-pair maps).
bool launch_bwd_persist(int BN, const CUtensorMap& tmdG, const CUtensorMap& tmUkr, const CUtensorMap& tmdY, const CUtensorMap& tmWnm,
                        const BwdPersistArgs& a0, cudaStream_t st) {
  if (bwd_pair(a0.Bp)) return false;                         // the pair variants use half-height weight boxes
  static const bool spread = getenv("LSTM_PERSIST_SPREAD") != nullptr && atoi(getenv("LSTM_PERSIST_SPREAD")) != 0;
  static const int wfence = getenv("LSTM_PERSIST_WFENCE") ? atoi(getenv("LSTM_PERSIST_WFENCE")) : 1;
  BwdPersistArgs a = a0;
  a.bar_stride = spread ? 32 : 1;
  a.writer_fence = wfence;
  if (BN == 128) return launch_persist_bwd_t<128>(tmdG, tmUkr, tmdY, tmWnm, a, st);
  if (BN == 64) return launch_persist_bwd_t<64>(tmdG, tmUkr, tmdY, tmWnm, a, st);
  return false;
}

}  // namespace tc
