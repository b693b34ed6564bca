// tc_kernels.cu — the bf16 tensor-core path: TMA-fed tcgen05.mma kernels with TMEM accumulators and
// the LSTM math fused into their epilogues.  One output tile (128 rows x BN columns) per CTA;
// warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2-5 = epilogue
// (each owns the 32 TMEM lanes of its warp-id % 4 quarter).  Layouts: tc_kernels.cuh.
//
// Reference statements computed here (R/ = reference root):
//   k_fwd_step  R/lstm.cc:176-192   g = W x + U h(t-1) + b, gates, c = tanh(i u + f c(t-1)), h = o c
//   k_logits    R/lstm.cc:195-207,225   y = Why h + by, softmax, -log2 p[target], dy = p - target
//   k_bwd_step  R/lstm.cc:228-256   dh = Why^T dy + U^T dg(t+1), gate gradients, dcnext
//   k_gemm_nt   R/lstm.cc:226-227,250-252   dWhy|dby and dW|dU|db as two batched-over-time GEMMs
#include "tc_common.cuh"
#include "tc_kernels.cuh"
#include "tc_tile.cuh"

namespace tc {

// ------------------------------------------------------------------------------------------------
// K3: logits for all (t,b) rows + fused softmax / loss / dy.  D[(s,b)][m] = sum_n h[(s+1,b)][n] Why[m][n]
// PERSISTENT: one CTA per SM loops over the 128-row tiles; the accumulator is DOUBLE-BUFFERED in TMEM (2 x 256 columns) so that
// the softmax epilogue of tile i (each epilogue thread owns one (t,b) row with all M = 256 logits) runs under the contraction
// of tile i+1.
// ------------------------------------------------------------------------------------------------
constexpr int K3_BN = 256, K3_STAGES = 4;
__global__ void __launch_bounds__(192, 1)
k_logits(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW, const LogitsArgs a) {
  using C = Cfg<K3_BN, K3_STAGES>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_by[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_by[i] = i < a.M ? a.by[i] : -1e30f;
  // prologue (tile_prologue with two accumulators): barriers full/empty per stage, acc_full/acc_empty per accumulator
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* tiles = smem_raw + (((base + 1023u) & ~1023u) - base);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + C::TILE_BYTES);
  uint64_t *full = bars, *empty = bars + K3_STAGES, *acc_full = bars + 2 * K3_STAGES, *acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < K3_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_d = *tmem_slot;
  const int n_tiles = a.T * a.Bp / BM;
  const int nkb = a.N / BK;
  if (warp == 0) {
    int g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int row0 = tile * BM;                          // row in (slot s, padded b) space; h of slot s+1
      if (a.progress) {                                    // beside a running forward recurrence: wait for h of that timestep
        const int s = row0 / a.Bp, mb = (row0 - s * a.Bp) / BM;
        counters_wait(a.progress + (size_t)mb * R_SLOTS, R_SLOTS, (unsigned int)(s + 1) * a.per_slot, lane, 500u);
        fence_proxy_async_global();
      }
      for (int kb = 0; kb < nkb; kb++, g++) {
        const int st = g % K3_STAGES;
        mbar_wait(&empty[st], ((uint32_t)(g / K3_STAGES) & 1u) ^ 1u);
        if (elect_one()) {
          uint8_t* sa = tiles + (size_t)st * C::STAGE_BYTES;
          mbar_expect_tx(&full[st], (uint32_t)C::STAGE_BYTES);
          tma_load_2d(sa, &tmH, &full[st], kb * BK, row0 + a.Bp);
          tma_load_2d_hint(sa + A_TILE_BYTES, &tmW, &full[st], kb * BK, 0, L2_EVICT_LAST);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(BM, K3_BN);
    const uint64_t a_desc0 = make_smem_desc_sw128(smem_u32(tiles));
    const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(tiles) + A_TILE_BYTES);
    int g = 0, it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int slot = it & 1;
      mbar_wait(&acc_empty[slot], ((uint32_t)(it >> 1) & 1u) ^ 1u);   // the epilogue has read this accumulator's previous tile
      tcgen05_after_sync();
      for (int kb = 0; kb < nkb; kb++, g++) {
        const int st = g % K3_STAGES;
        mbar_wait(&full[st], (uint32_t)(g / K3_STAGES) & 1u);
        tcgen05_after_sync();
        if (elect_one()) {
          const uint64_t soff = (uint64_t)((uint32_t)st * (uint32_t)(C::STAGE_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BK / 16; k++)
            umma_bf16(tmem_d + (uint32_t)(slot * K3_BN), a_desc0 + soff + 2 * k, b_desc0 + soff + 2 * k, idesc, (uint32_t)((kb | k) != 0));
          umma_commit(&empty[st]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_full[slot]);
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    constexpr float LOG2E = 1.4426950408889634f;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
      const int slot = it & 1;
      const int q = tile * BM + quarter * 32 + lane;
      const int s = q / a.Bp, b = q - s * a.Bp;
      const bool valid = b < a.B;
      const int k = valid ? a.tg[(size_t)s * a.B + b] : -1;
      const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * K3_BN);
      mbar_wait(&acc_full[slot], (uint32_t)(it >> 1) & 1u);
      tcgen05_after_sync();
      float mx = -1e30f;
#pragma unroll 1
      for (int c0 = 0; c0 < K3_BN; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
#pragma unroll
        for (int i = 0; i < 32; i++) mx = fmaxf(mx, v[i] + s_by[c0 + i]);
      }
      float sum = 0.f, yk = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < K3_BN; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const float y = v[i] + s_by[c0 + i];
          sum += exp2f((y - mx) * LOG2E);
          if (c0 + i == k) yk = y;
        }
      }
      const float inv = __fdividef(1.0f, sum);
#pragma unroll 1
      for (int c0 = 0; c0 < K3_BN; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (valid) {
          float dy[32];
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const float p = exp2f((v[i] + s_by[c0 + i] - mx) * LOG2E) * inv;
            dy[i] = (c0 + i == k) ? p - 1.0f : p;          // dy = probs - target (R/lstm.cc:225)
          }
          uint4* row = reinterpret_cast<uint4*>(a.dYbf + (size_t)q * a.M + c0);
#pragma unroll
          for (int w = 0; w < 4; w++) {
            uint4 o;
            o.x = pack_bf16x2(dy[8 * w + 0], dy[8 * w + 1]); o.y = pack_bf16x2(dy[8 * w + 2], dy[8 * w + 3]);
            o.z = pack_bf16x2(dy[8 * w + 4], dy[8 * w + 5]); o.w = pack_bf16x2(dy[8 * w + 6], dy[8 * w + 7]);
            if (c0 + 8 * w < a.M) row[w] = o;
          }
          const size_t ldt = (size_t)a.T * a.Bp;
#pragma unroll
          for (int i = 0; i < 32; i++)
            if (c0 + i < a.M) a.dYT[(size_t)(c0 + i) * ldt + q] = __float2bfloat16_rn(dy[i]);
        }
      }
      tcgen05_before_sync();
      mbar_arrive(&acc_empty[slot]);                       // this thread's TMEM lane of the accumulator is free again
      if (valid) a.surp[(size_t)s * a.B + b] = (k >= 0) ? -((yk - mx) * LOG2E - log2f(sum)) : 0.f;  // -log2 p[k]
    }
  }
  tcgen05_before_sync();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<512>(tmem_d);
  }
}

// ------------------------------------------------------------------------------------------------
// K6: C(row, col) = sum_k A[row][k] B[col][k], 128 x 128 tiles, fp32 result stored column-major
// (C[col*ldc + row]: the TMEM lane dimension is the contiguous one, so every warp store is 128 B).
// Tiles are rasterised in groups of 16 row-tiles so concurrently resident CTAs share operand rows in L2.
// ------------------------------------------------------------------------------------------------
// K6_BN = 256 (one N = 256 MMA per k-step) moves 48 KB of operands per 512 MMA cycles instead of 32 KB per 256: the
// big dW|dU|db GEMM uses it; the small dWhy|dby GEMM keeps 128-wide tiles so that more CTAs share its long K loop.
template <int K6_BN>
__global__ void __launch_bounds__(192, 1)
k_gemm_nt(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs a) {
  constexpr int K6_STAGES = K6_BN == 256 ? 4 : 6;
  extern __shared__ uint8_t smem_raw[];
  TileCtx c = tile_prologue<K6_BN, K6_STAGES>(smem_raw);
  constexpr int GROUP = 16;
  const int pid = blockIdx.x;
  const int per_group = GROUP * a.tiles_n;
  const int group = pid / per_group;
  const int first_m = group * GROUP;
  const int gsize = min(a.tiles_m - first_m, GROUP);
  const int tm = first_m + (pid % per_group) % gsize;
  const int tn = (pid % per_group) / gsize;
  // split-K (blockIdx.y): every split contracts its own nkb k-blocks into its own partial output (deterministic: the partials are
  // summed in a fixed order by k_sum_splits)
  const int split = blockIdx.y;
  float* C = a.C + (size_t)split * a.split_stride;
  const KSeg s0{&tmA, &tmB, tm * BM, a.b_row0 + tn * K6_BN, a.a_k0 + split * a.nkb * BK, a.b_k0 + split * a.nkb * BK, a.nkb};
  const KSeg s1{&tmA, &tmB, 0, 0, 0, 0, 0};
  if (a.wait_slots && c.warp == 0) {   // beside a running BPTT recurrence: this K range of dG is complete once the counters say so
    counters_wait(a.wait_slots, a.wait_n, a.wait_target, c.lane, 1000u);
    fence_proxy_async_global();
  }
  tile_mainloop<K6_BN, K6_STAGES>(c, s0, s1);
  if (c.warp >= 2) {
    const int quarter = c.warp & 3;
    const int row = tm * BM + quarter * 32 + c.lane;
    mbar_wait(c.accum_full, 0);
    tcgen05_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < K6_BN; c0 += 32) {
      float v[32];
      tmem_ld32(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      if (row < a.rows) {
        if (a.addend) {
          float ad[32];                                        // all 32 loads in flight before the first store (C may alias nothing,
#pragma unroll                                                 // but the compiler cannot know: one load per store would serialise)
          for (int i = 0; i < 32; i++) {
            const int col = tn * K6_BN + c0 + i;
            ad[i] = (col < a.cols) ? __ldg(a.addend + (size_t)col * a.ldc + row) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const int col = tn * K6_BN + c0 + i;
            if (col < a.cols) C[(size_t)col * a.ldc + row] = v[i] + ad[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const int col = tn * K6_BN + c0 + i;
            if (col < a.cols) C[(size_t)col * a.ldc + row] = v[i];
          }
        }
      }
    }
  }
  tile_epilogue_end<K6_BN, K6_STAGES>(c);
}

// The same GEMM as cta_group::2 PAIRS: a 256 x 256 output tile per pair of CTAs, every CTA stages its own 128 A rows and HALF of the
// 256-row B tile — 32 KB of operands per k-block and SM instead of 48 KB for the same 512 tensor-core cycles.  The single-CTA
// kernel above is bound by what one SM can take in (~61 B per clock measured: 790 cycles per k-block instead of 512).
constexpr int K6P_BN = 256, K6P_STAGES = 6;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
k_gemm_nt_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs a) {
  extern __shared__ uint8_t smem_raw[];
  TileCtx c = pair_prologue<K6P_BN, K6P_STAGES>(smem_raw);
  const uint32_t rank = cluster_ctarank();
  constexpr int GROUP = 8;                                   // pair-rows per raster group (16 row tiles, like k_gemm_nt)
  const int pid = blockIdx.x >> 1;
  const int pair_rows = a.tiles_m >> 1;
  const int per_group = GROUP * a.tiles_n;
  const int group = pid / per_group;
  const int first_m = group * GROUP;
  const int gsize = min(pair_rows - first_m, GROUP);
  const int tm = 2 * (first_m + (pid % per_group) % gsize) + (int)rank;   // this CTA's 128-row tile
  const int tn = (pid % per_group) / gsize;
  const KSeg s0{&tmA, &tmB, tm * BM, a.b_row0 + tn * K6P_BN, a.a_k0, a.b_k0, a.nkb};
  const KSeg s1{&tmA, &tmB, 0, 0, 0, 0, 0};
  if (a.wait_slots && c.warp == 0) {   // beside a running BPTT recurrence: this K range of dG is complete once the counters say so
    counters_wait(a.wait_slots, a.wait_n, a.wait_target, c.lane, 1000u);
    fence_proxy_async_global();
  }
  pair_mainloop<K6P_BN, K6P_STAGES>(c, s0, s1, rank, (uint16_t)0x3);
  if (c.warp >= 2) {
    const int quarter = c.warp & 3;
    const int row = tm * BM + quarter * 32 + c.lane;
    float* C = a.C;
    mbar_wait(c.accum_full, 0);
    tcgen05_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < K6P_BN; c0 += 32) {
      float v[32];
      tmem_ld32(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      if (row < a.rows) {
        if (a.addend) {
          float ad[32];
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const int col = tn * K6P_BN + c0 + i;
            ad[i] = (col < a.cols) ? __ldg(a.addend + (size_t)col * a.ldc + row) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const int col = tn * K6P_BN + c0 + i;
            if (col < a.cols) C[(size_t)col * a.ldc + row] = v[i] + ad[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const int col = tn * K6P_BN + c0 + i;
            if (col < a.cols) C[(size_t)col * a.ldc + row] = v[i];
          }
        }
      }
    }
  }
  pair_epilogue_end<K6P_BN, K6P_STAGES>(c);
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------

// launch `kernel` as the programmatic dependent of the kernel before it in `st`: it may start as soon as every CTA of that kernel
// has executed griddepcontrol.launch_dependents (it never calls griddepcontrol.wait: the two run side by side)
template <typename K, typename... Args>
static void launch_beside(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}

void launch_logits(const CUtensorMap& tmH, const CUtensorMap& tmWmn, const LogitsArgs& a, cudaStream_t st, int max_ctas, bool beside) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int tiles = a.T * a.Bp / BM;
  int grid = tiles < sms ? tiles : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  set_smem(k_logits, Cfg<K3_BN, K3_STAGES>::SMEM_BYTES);
  if (beside) launch_beside(k_logits, dim3(grid), dim3(192), (size_t)Cfg<K3_BN, K3_STAGES>::SMEM_BYTES, st, tmH, tmWmn, a);
  else k_logits<<<grid, 192, Cfg<K3_BN, K3_STAGES>::SMEM_BYTES, st>>>(tmH, tmWmn, a);
}

// bn = 128 or 256 = tile width; tmB must have a box of bn rows and a.tiles_n = ceil(cols / bn)
void launch_gemm_nt(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st, bool beside, bool pair) {
  const dim3 grid(a.tiles_m * a.tiles_n, a.splits > 1 ? a.splits : 1);
  if (bn == 256 && pair) {
    // tmB must have a box of 128 rows here: each CTA of a pair loads half of the 256-row B tile
    constexpr int SMEM = PairCfg<K6P_BN, K6P_STAGES>::TILE_BYTES + 1024 + 256;
    set_smem(k_gemm_nt_pair, SMEM);
    if (beside) launch_beside(k_gemm_nt_pair, grid, dim3(192), (size_t)SMEM, st, tmA, tmB, a);
    else k_gemm_nt_pair<<<grid, 192, SMEM, st>>>(tmA, tmB, a);
  } else if (bn == 256) {
    set_smem(k_gemm_nt<256>, Cfg<256, 4>::SMEM_BYTES);
    if (beside) launch_beside(k_gemm_nt<256>, grid, dim3(192), (size_t)Cfg<256, 4>::SMEM_BYTES, st, tmA, tmB, a);
    else k_gemm_nt<256><<<grid, 192, Cfg<256, 4>::SMEM_BYTES, st>>>(tmA, tmB, a);
  } else {
    set_smem(k_gemm_nt<128>, Cfg<128, 6>::SMEM_BYTES);
    if (beside) launch_beside(k_gemm_nt<128>, grid, dim3(192), (size_t)Cfg<128, 6>::SMEM_BYTES, st, tmA, tmB, a);
    else k_gemm_nt<128><<<grid, 192, Cfg<128, 6>::SMEM_BYTES, st>>>(tmA, tmB, a);
  }
}

// out[i] = parts[0][i] + parts[1][i] + ... in that order (split-K partials of k_gemm_nt)
__global__ void k_sum_splits(const float* __restrict__ parts, float* __restrict__ out, size_t n4, size_t stride4, int splits) {
  const float4* p = reinterpret_cast<const float4*>(parts);
  float4* o = reinterpret_cast<float4*>(out);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = p[i];
    for (int s = 1; s < splits; s++) {
      const float4 x = p[i + (size_t)s * stride4];
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    o[i] = acc;
  }
}
void launch_sum_splits(const float* parts, float* out, size_t n, size_t stride, int splits, cudaStream_t st) {
  k_sum_splits<<<148 * 4, 256, 0, st>>>(parts, out, n / 4, stride / 4, splits);
}

// ------------------------------------------------------------------------------------------------
// conversion kernels (fp32 masters -> bf16 operand copies in the layouts above)
// ------------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ float cvt_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void k_permute_rows(const float* __restrict__ in, T* __restrict__ out, int N) {
  const int N4 = 4 * N;
  const int rp = blockIdx.x * blockDim.x + threadIdx.x;
  if (rp >= N4) return;
  const size_t row = blockIdx.y;
  out[row * N4 + rp] = cvt_out<T>(in[row * N4 + (size_t)(rp & 3) * N + (rp >> 2)]);
}
void launch_permute_rows_f32(const float* in, float* out, int rows, int N, cudaStream_t st) {
  k_permute_rows<float><<<dim3((4 * N + 255) / 256, rows), 256, 0, st>>>(in, out, N);
}
void launch_permute_rows_bf16(const float* in, __nv_bfloat16* out, int rows, int N, cudaStream_t st) {
  k_permute_rows<__nv_bfloat16><<<dim3((4 * N + 255) / 256, rows), 256, 0, st>>>(in, out, N);
}

__global__ void k_transpose_cast(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C, int permute) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  {
    const int cp = c0 + tx;
    const int src = permute ? (cp & 3) * (C >> 2) + (cp >> 2) : cp;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int r = r0 + ty + 8 * i;
      tile[ty + 8 * i][tx] = (r < R && cp < C) ? in[(size_t)r * C + src] : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int cp = c0 + ty + 8 * i, r = r0 + tx;
    if (cp < C && r < R) out[(size_t)cp * R + r] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
  }
}
void launch_transpose_cast(const float* in, __nv_bfloat16* out, int R, int C, int permute, cudaStream_t st) {
  k_transpose_cast<<<dim3((C + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, st>>>(in, out, R, C, permute);
}

__global__ void k_cast_bf16(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
static unsigned grid_for(size_t n) {
  size_t b = (n + 255) / 256;
  if (b > 148 * 8) b = 148 * 8;
  return (unsigned)(b ? b : 1);
}
void launch_cast_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t st) {
  k_cast_bf16<<<grid_for(n), 256, 0, st>>>(in, out, n);
}
__global__ void k_bf16_to_f32(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}
void launch_bf16_to_f32(const __nv_bfloat16* in, float* out, size_t n, cudaStream_t st) {
  k_bf16_to_f32<<<grid_for(n), 256, 0, st>>>(in, out, n);
}
__global__ void k_fill_bf16(__nv_bfloat16* p, float v, size_t n) {
  const __nv_bfloat16 b = __float2bfloat16_rn(v);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = b;
}
void launch_fill_bf16(__nv_bfloat16* p, float v, size_t n, cudaStream_t st) { k_fill_bf16<<<grid_for(n), 256, 0, st>>>(p, v, n); }

// ZT rows [0,M): one-hot of the window's input bytes (W*x / dW += dg x^T as tensor-core contractions)
__global__ void __launch_bounds__(256) k_build_xt(const int* __restrict__ xs1, __nv_bfloat16* __restrict__ ZT, long ldz, int T,
                                                  int B, int Bp) {
  // one thread = 8 consecutive (timestep, stream) columns of symbol row m: one 16-byte store (Bp is a multiple of 8)
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (col >= T * Bp) return;
  const int m = blockIdx.y;
  const int s = col / Bp, b0 = col - s * Bp;
  const unsigned short one = 0x3F80;   // bf16 1.0
  unsigned short v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = (b0 + i < B && xs1[(size_t)s * B + b0 + i] == m) ? one : (unsigned short)0;
  uint4 o;
  o.x = v[0] | ((unsigned)v[1] << 16); o.y = v[2] | ((unsigned)v[3] << 16);
  o.z = v[4] | ((unsigned)v[5] << 16); o.w = v[6] | ((unsigned)v[7] << 16);
  *reinterpret_cast<uint4*>(ZT + (size_t)m * ldz + col) = o;
}
void launch_build_xt(const int* xs1, __nv_bfloat16* ZT, long ldz, int M, int T, int B, int Bp, cudaStream_t st, bool beside) {
  const dim3 grid((T * Bp / 8 + 255) / 256, M);
  if (beside) launch_beside(k_build_xt, grid, dim3(256), (size_t)0, st, xs1, ZT, ldz, T, B, Bp);
  else k_build_xt<<<grid, 256, 0, st>>>(xs1, ZT, ldz, T, B, Bp);
}

__global__ void k_state_to_bf16(const float* __restrict__ h, __nv_bfloat16* __restrict__ Hbf, __nv_bfloat16* __restrict__ ZT_h,
                                long ldz, int B, int N) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * N) return;
  const int b = (int)(i / N), n = (int)(i - (size_t)b * N);
  const __nv_bfloat16 v = __float2bfloat16_rn(h[i]);
  Hbf[i] = v;
  ZT_h[(size_t)n * ldz + b] = v;
}
void launch_state_to_bf16(const float* h, __nv_bfloat16* Hbf_slot, __nv_bfloat16* ZT_h, long ldz, int B, int N, cudaStream_t st) {
  k_state_to_bf16<<<(unsigned)(((size_t)B * N + 255) / 256), 256, 0, st>>>(h, Hbf_slot, ZT_h, ldz, B, N);
}
void launch_state_to_f32(const __nv_bfloat16* Hbf_slot, float* h, int B, int N, cudaStream_t st) {
  launch_bf16_to_f32(Hbf_slot, h, (size_t)B * N, st);
}

// All bf16 operand copies of U in ONE pass over the fp32 master (read once, coalesced; every write a 128-byte row segment):
// one CTA transposes the tile [64 k] x [16 units x 4 gates] through shared memory.
//   Urk [4N r'][N k]   and its blocked form Wb2[(tile*N/64 + kb)*bn2 + row][64]   (r' = 4*unit + gate; K2 operands)
//   Ukr [N k][4N r']   and its blocked form Wb5[(tile*NKBG + kbg)*bn5 + row][64]  (K5 operands; NKBG = 4N/64 + M/64)
// Null outputs are skipped.  grid (N/16, N/64), 256 threads.
__global__ void __launch_bounds__(256) k_refresh_u(const float* __restrict__ U, __nv_bfloat16* __restrict__ Urk,
                                                   __nv_bfloat16* __restrict__ Ukr, __nv_bfloat16* __restrict__ Wb2, int bn2,
                                                   __nv_bfloat16* __restrict__ Wb5, int bn5, int nkbg, int N) {
  __shared__ float tile[64][65];                             // [k][r' local]
  const int N4 = 4 * N;
  const int j0 = blockIdx.x * 16, kb = blockIdx.y, k0 = kb * 64;
  const int rp0 = 4 * j0;                                    // a multiple of 64: one k-block of the r' axis
  // 128-bit loads: 4 consecutive hidden units of one (k, gate)
  for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
    const int q = idx & 3, g = (idx >> 2) & 3, k = idx >> 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(U + (size_t)(k0 + k) * N4 + (size_t)g * N + j0 + 4 * q));
    tile[k][4 * (4 * q + 0) + g] = v.x; tile[k][4 * (4 * q + 1) + g] = v.y;
    tile[k][4 * (4 * q + 2) + g] = v.z; tile[k][4 * (4 * q + 3) + g] = v.w;
  }
  __syncthreads();
  // every output row is 64 bf16 = 128 bytes; one thread writes 8 of them (16 bytes), 8 threads a whole row
  // rows = r', 64 consecutive k each (Urk / Wb2)
  for (int idx = threadIdx.x; idx < 64 * 8; idx += 256) {
    const int c8 = idx & 7, r = idx >> 3;
    uint4 o;
    o.x = pack_bf16x2(tile[8 * c8 + 0][r], tile[8 * c8 + 1][r]); o.y = pack_bf16x2(tile[8 * c8 + 2][r], tile[8 * c8 + 3][r]);
    o.z = pack_bf16x2(tile[8 * c8 + 4][r], tile[8 * c8 + 5][r]); o.w = pack_bf16x2(tile[8 * c8 + 6][r], tile[8 * c8 + 7][r]);
    const int rp = rp0 + r;
    if (Urk) *reinterpret_cast<uint4*>(Urk + (size_t)rp * N + k0 + 8 * c8) = o;
    if (Wb2) *reinterpret_cast<uint4*>(Wb2 + (((size_t)(rp / bn2) * (N / 64) + kb) * bn2 + rp % bn2) * 64 + 8 * c8) = o;
  }
  // rows = k, 64 consecutive r' each (Ukr / Wb5)
  for (int idx = threadIdx.x; idx < 64 * 8; idx += 256) {
    const int c8 = idx & 7, kk = idx >> 3;
    const float* t = &tile[kk][8 * c8];
    uint4 o;
    o.x = pack_bf16x2(t[0], t[1]); o.y = pack_bf16x2(t[2], t[3]); o.z = pack_bf16x2(t[4], t[5]); o.w = pack_bf16x2(t[6], t[7]);
    const int k = k0 + kk;
    if (Ukr) *reinterpret_cast<uint4*>(Ukr + (size_t)k * N4 + rp0 + 8 * c8) = o;
    if (Wb5) *reinterpret_cast<uint4*>(Wb5 + (((size_t)(k / bn5) * nkbg + rp0 / 64) * bn5 + k % bn5) * 64 + 8 * c8) = o;
  }
}
void launch_refresh_u(const float* U, __nv_bfloat16* Urk, __nv_bfloat16* Ukr, __nv_bfloat16* Wb2, int bn2, __nv_bfloat16* Wb5,
                      int bn5, int M, int N, cudaStream_t st) {
  k_refresh_u<<<dim3(N / 16, N / 64), 256, 0, st>>>(U, Urk, Ukr, Wb2, bn2 ? bn2 : 64, Wb5, bn5 ? bn5 : 64, (4 * N + M) / 64, N);
}
// the Why k-blocks of the blocked BPTT weight copy: Wb5[(tile*NKBG + 4N/64 + mkb)*bn5 + row][c] = Why(m = mkb*64 + c, j = tile*bn5 + row)
__global__ void k_block_why(const float* __restrict__ Why, __nv_bfloat16* __restrict__ Wb5, int N, int M, int bn5) {
  const int nkbu = 4 * N / 64, nkbg = nkbu + M / 64;
  const size_t total = (size_t)N * M;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx % M), j = (int)(idx / M);
    Wb5[(((size_t)(j / bn5) * nkbg + nkbu + m / 64) * bn5 + j % bn5) * 64 + (m & 63)] = __float2bfloat16_rn(Why[idx]);
  }
}
void launch_block_why(const float* Why, __nv_bfloat16* Wb5, int N, int M, int bn5, cudaStream_t st) {
  k_block_why<<<148 * 2, 256, 0, st>>>(Why, Wb5, N, M, bn5);
}

template <typename T>
__global__ void k_unpermute(const T* __restrict__ in, float* __restrict__ out, int N) {
  const int N4 = 4 * N;
  const int rp = blockIdx.x * blockDim.x + threadIdx.x;
  if (rp >= N4) return;
  const size_t row = blockIdx.y;
  out[row * N4 + (size_t)(rp & 3) * N + (rp >> 2)] = (float)in[row * N4 + rp];
}
void launch_unpermute_f32(const float* in, float* out, int rows, int N, cudaStream_t st) {
  k_unpermute<float><<<dim3((4 * N + 255) / 256, rows), 256, 0, st>>>(in, out, N);
}
void launch_unpermute_bf16(const __nv_bfloat16* in, float* out, int rows, int N, cudaStream_t st) {
  k_unpermute<__nv_bfloat16><<<dim3((4 * N + 255) / 256, rows), 256, 0, st>>>(in, out, N);
}

}  // namespace tc
