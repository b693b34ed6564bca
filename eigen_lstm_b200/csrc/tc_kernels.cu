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

namespace tc {

constexpr int BM = 128;                 // tile rows = TMEM lanes
constexpr int BK = 64;                  // bf16 per 128-byte swizzled row
constexpr int A_TILE_BYTES = BM * BK * 2;

struct KSeg {                           // one contiguous K range of the contraction
  const CUtensorMap* ta;
  const CUtensorMap* tb;
  int a_row, b_row, a_k0, b_k0, nkb;
};

template <int BN, int STAGES>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
};

struct TileCtx {
  uint8_t* tiles;
  uint64_t *full, *empty, *accum_full;
  uint32_t tmem_d;
  int warp, lane;
};

// Common prologue: carve shared memory, init barriers, allocate TMEM.
template <int BN, int STAGES>
__device__ __forceinline__ TileCtx tile_prologue(uint8_t* raw) {
  using C = Cfg<BN, STAGES>;
  TileCtx c;
  const uint32_t base = smem_u32(raw);
  const uint32_t pad = ((base + 1023u) & ~1023u) - base;   // SWIZZLE_128B atoms need 1024-byte alignment
  c.tiles = raw + pad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(c.tiles + C::TILE_BYTES);
  c.full = bars;
  c.empty = bars + STAGES;
  c.accum_full = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
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
  if (c.warp == 1) tmem_dealloc<Cfg<BN, STAGES>::TMEM_COLS>(c.tmem_d);
}

// Producer (warp 0, one lane) and MMA issuer (warp 1, one lane) of one output tile.
template <int BN, int STAGES>
__device__ __forceinline__ void tile_mainloop(const TileCtx& c, const KSeg& s0, const KSeg& s1) {
  using C = Cfg<BN, STAGES>;
  const int total = s0.nkb + s1.nkb;
  if (c.warp == 0) {
    if (c.lane == 0) {
      for (int kb = 0; kb < total; kb++) {
        const int st = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&c.empty[st], ph ^ 1u);
        const bool first = kb < s0.nkb;
        const KSeg& s = first ? s0 : s1;
        const int k = first ? kb : kb - s0.nkb;
        uint8_t* a = c.tiles + (size_t)st * C::STAGE_BYTES;
        uint8_t* b = a + A_TILE_BYTES;
        mbar_expect_tx(&c.full[st], (uint32_t)C::STAGE_BYTES);
        tma_load_2d(a, s.ta, &c.full[st], s.a_k0 + k * BK, s.a_row);
        tma_load_2d(b, s.tb, &c.full[st], s.b_k0 + k * BK, s.b_row);
      }
    }
  } else if (c.warp == 1) {
    if (c.lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      for (int kb = 0; kb < total; kb++) {
        const int st = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&c.full[st], ph);
        tcgen05_after_sync();
        const uint32_t a_addr = smem_u32(c.tiles + (size_t)st * C::STAGE_BYTES);
        const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; k++)   // UMMA_K = 16 bf16 = 32 bytes inside the swizzled row
          umma_bf16(c.tmem_d, make_smem_desc_sw128(a_addr + k * 32), make_smem_desc_sw128(b_addr + k * 32), idesc,
                    (uint32_t)((kb | k) != 0));
        umma_commit(&c.empty[st]);          // frees the smem stage once these MMAs have read it
      }
      umma_commit(c.accum_full);            // accumulator complete -> epilogue
    }
  }
}

__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }

// ------------------------------------------------------------------------------------------------
// K2: forward timestep.  D[b][r'] = sum_k h(t-1)[b][k] * U[r'][k]; epilogue = R/lstm.cc:176-192.
// grid (4N/BN, Bp/128)
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
k_fwd_step(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmU, const FwdStepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  TileCtx c = tile_prologue<BN, STAGES>(smem_raw);
  const int nb = blockIdx.x, mb = blockIdx.y;
  const KSeg s0{&tmH, &tmU, a.a_row0 + mb * BM, nb * BN, 0, 0, a.N / BK};
  const KSeg s1{&tmH, &tmU, 0, 0, 0, 0, 0};
  tile_mainloop<BN, STAGES>(c, s0, s1);
  if (c.warp >= 2) {
    const int quarter = c.warp & 3;
    const int b = mb * BM + quarter * 32 + c.lane;
    const bool valid = b < a.B;
    const int N = a.N, N4 = 4 * a.N;
    const int x = valid ? a.x[b] : -1;
    mbar_wait(c.accum_full, 0);
    tcgen05_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      if (valid) {
        const int r0 = nb * BN + c0, j0 = r0 >> 2;
        const float4* bq = reinterpret_cast<const float4*>(a.bp + r0);
        const float4* wq = reinterpret_cast<const float4*>(a.Wp + (size_t)(x >= 0 ? x : 0) * N4 + r0);
        const float4 cpa = *reinterpret_cast<const float4*>(a.c_prev + (size_t)b * N + j0);
        const float4 cpb = *reinterpret_cast<const float4*>(a.c_prev + (size_t)b * N + j0 + 4);
        const float cp[8] = {cpa.x, cpa.y, cpa.z, cpa.w, cpb.x, cpb.y, cpb.z, cpb.w};
        float hv[8], cv[8];
        float4* gp = reinterpret_cast<float4*>(a.Gp_t + (size_t)b * N4 + r0);
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const float4 bb = bq[u];
          float4 ww = make_float4(0.f, 0.f, 0.f, 0.f);
          if (x >= 0) ww = wq[u];                         // W*x for one-hot x: row x of the permuted W
          const float gi = sigmoid_fast(v[4 * u + 0] + ww.x + bb.x);
          const float go = sigmoid_fast(v[4 * u + 1] + ww.y + bb.y);
          const float gf = sigmoid_fast(v[4 * u + 2] + ww.z + bb.z);
          const float gu = tanh_fast(v[4 * u + 3] + ww.w + bb.w);
          const float cc = tanh_fast(gi * gu + gf * cp[u]);   // the carried cell value is the tanh'd one
          cv[u] = cc;
          hv[u] = go * cc;
          gp[u] = make_float4(gi, go, gf, gu);
        }
        float4* co = reinterpret_cast<float4*>(a.c_out + (size_t)b * N + j0);
        co[0] = make_float4(cv[0], cv[1], cv[2], cv[3]);
        co[1] = make_float4(cv[4], cv[5], cv[6], cv[7]);
        uint4 hb;
        hb.x = pack_bf16x2(hv[0], hv[1]); hb.y = pack_bf16x2(hv[2], hv[3]);
        hb.z = pack_bf16x2(hv[4], hv[5]); hb.w = pack_bf16x2(hv[6], hv[7]);
        *reinterpret_cast<uint4*>(a.Hbf_t + (size_t)b * N + j0) = hb;
#pragma unroll
        for (int u = 0; u < 8; u++) a.ZT_h[(size_t)(j0 + u) * a.ldz + b] = __float2bfloat16_rn(hv[u]);
      }
    }
  }
  tile_epilogue_end<BN, STAGES>(c);
}

// ------------------------------------------------------------------------------------------------
// K3: logits for all (t,b) rows + fused softmax / loss / dy.  D[(s,b)][m] = sum_n h[(s+1,b)][n] Why[m][n]
// One CTA per 128 rows; each epilogue thread owns one (t,b) row with all M = 256 logits in TMEM.
// ------------------------------------------------------------------------------------------------
constexpr int K3_BN = 256, K3_STAGES = 4;
__global__ void __launch_bounds__(192, 1)
k_logits(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW, const LogitsArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_by[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_by[i] = i < a.M ? a.by[i] : -1e30f;
  TileCtx c = tile_prologue<K3_BN, K3_STAGES>(smem_raw);
  const int row0 = blockIdx.x * BM;                      // row in (slot s, padded b) space
  const KSeg s0{&tmH, &tmW, row0 + a.Bp, 0, 0, 0, a.N / BK};   // h of slot s+1
  const KSeg s1{&tmH, &tmW, 0, 0, 0, 0, 0};
  tile_mainloop<K3_BN, K3_STAGES>(c, s0, s1);
  if (c.warp >= 2) {
    const int quarter = c.warp & 3;
    const int q = row0 + quarter * 32 + c.lane;
    const int s = q / a.Bp, b = q - s * a.Bp;
    const bool valid = b < a.B;
    const int k = valid ? a.tg[(size_t)s * a.B + b] : -1;
    const uint32_t taddr = c.tmem_d + ((uint32_t)(quarter * 32) << 16);
    constexpr float LOG2E = 1.4426950408889634f;
    mbar_wait(c.accum_full, 0);
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
    if (valid) a.surp[(size_t)s * a.B + b] = (k >= 0) ? -((yk - mx) * LOG2E - log2f(sum)) : 0.f;  // -log2 p[k]
  }
  tile_epilogue_end<K3_BN, K3_STAGES>(c);
}

// ------------------------------------------------------------------------------------------------
// K5: backward timestep.  D[b][j] = sum_r' dg(t+1)[b][r'] U[r'][j]  +  sum_m dy(t)[b][m] Why[m][j]
// (the contraction's K range is the concatenation of the two); epilogue = R/lstm.cc:233-256.
// grid (N/BN, Bp/128)
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
k_bwd_step(const __grid_constant__ CUtensorMap tmdG, const __grid_constant__ CUtensorMap tmU,
           const __grid_constant__ CUtensorMap tmdY, const __grid_constant__ CUtensorMap tmW, const BwdStepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  TileCtx c = tile_prologue<BN, STAGES>(smem_raw);
  const int nb = blockIdx.x, mb = blockIdx.y;
  const KSeg s0{&tmdG, &tmU, a.dg_row0 + mb * BM, nb * BN, 0, 0, a.first ? 0 : (4 * a.N) / BK};
  const KSeg s1{&tmdY, &tmW, a.dy_row0 + mb * BM, nb * BN, 0, 0, a.M / BK};
  tile_mainloop<BN, STAGES>(c, s0, s1);
  if (c.warp >= 2) {
    const int quarter = c.warp & 3;
    const int b = mb * BM + quarter * 32 + c.lane;
    const bool valid = b < a.B;
    const int N = a.N, N4 = 4 * a.N;
    mbar_wait(c.accum_full, 0);
    tcgen05_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(c.tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      if (valid) {
        const int j0 = nb * BN + c0;
        const size_t bj = (size_t)b * N + j0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
          const float4 ct4 = *reinterpret_cast<const float4*>(a.c_t + bj + 4 * q);
          const float4 cp4 = *reinterpret_cast<const float4*>(a.c_prev + bj + 4 * q);
          float4 dn4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!a.first) dn4 = *reinterpret_cast<const float4*>(a.dcnext + bj + 4 * q);
          const float ct[4] = {ct4.x, ct4.y, ct4.z, ct4.w}, cp[4] = {cp4.x, cp4.y, cp4.z, cp4.w};
          const float dn[4] = {dn4.x, dn4.y, dn4.z, dn4.w};
          float dno[4];
          uint32_t packed[8];
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int j = j0 + 4 * q + e;
            const float4 g = *reinterpret_cast<const float4*>(a.Gp_t + (size_t)b * N4 + 4 * (size_t)j);  // i o f u
            const float dh = v[4 * q + e];
            const float dc = (dh * g.y + dn[e]) * (1.0f - ct[e] * ct[e]);      // :233-235
            const float d_o = dh * ct[e] * (g.y * (1.0f - g.y));               // :238,244
            const float d_i = dc * g.w * (g.x * (1.0f - g.x));                 // :239,244
            const float d_f = dc * cp[e] * (g.z * (1.0f - g.z));               // :240,244
            const float d_u = dc * g.x * (1.0f - g.w * g.w);                   // :241,247
            dno[e] = dc * g.z;                                                 // :256
            packed[2 * e] = pack_bf16x2(d_i, d_o);
            packed[2 * e + 1] = pack_bf16x2(d_f, d_u);
            __nv_bfloat16* gt = a.dGT_t + b;
            gt[(size_t)j * a.ldg] = __float2bfloat16_rn(d_i);
            gt[(size_t)(N + j) * a.ldg] = __float2bfloat16_rn(d_o);
            gt[(size_t)(2 * N + j) * a.ldg] = __float2bfloat16_rn(d_f);
            gt[(size_t)(3 * N + j) * a.ldg] = __float2bfloat16_rn(d_u);
          }
          uint4* drow = reinterpret_cast<uint4*>(a.dGbf_t + (size_t)b * N4 + 4 * (size_t)(j0 + 4 * q));
          drow[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          drow[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
          *reinterpret_cast<float4*>(a.dcnext + bj + 4 * q) = make_float4(dno[0], dno[1], dno[2], dno[3]);
        }
      }
    }
  }
  tile_epilogue_end<BN, STAGES>(c);
}

// ------------------------------------------------------------------------------------------------
// K6: C(row, col) = sum_k A[row][k] B[col][k], 128 x 128 tiles, fp32 result stored column-major
// (C[col*ldc + row]: the TMEM lane dimension is the contiguous one, so every warp store is 128 B).
// Tiles are rasterised in groups of 16 row-tiles so concurrently resident CTAs share operand rows in L2.
// ------------------------------------------------------------------------------------------------
constexpr int K6_BN = 128, K6_STAGES = 6;
__global__ void __launch_bounds__(192, 1)
k_gemm_nt(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs a) {
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
  const KSeg s0{&tmA, &tmB, tm * BM, a.b_row0 + tn * K6_BN, a.a_k0, a.b_k0, a.nkb};
  const KSeg s1{&tmA, &tmB, 0, 0, 0, 0, 0};
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
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const int col = tn * K6_BN + c0 + i;
          if (col < a.cols) a.C[(size_t)col * a.ldc + row] = v[i];
        }
      }
    }
  }
  tile_epilogue_end<K6_BN, K6_STAGES>(c);
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
template <typename K>
static void set_smem(K kernel, int bytes) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

void launch_fwd_step(int BN, const CUtensorMap& tmH, const CUtensorMap& tmUrk, const FwdStepArgs& a, cudaStream_t st) {
  dim3 grid(4 * a.N / BN, a.Bp / BM);
  if (BN == 128) {
    set_smem(k_fwd_step<128, 6>, Cfg<128, 6>::SMEM_BYTES);
    k_fwd_step<128, 6><<<grid, 192, Cfg<128, 6>::SMEM_BYTES, st>>>(tmH, tmUrk, a);
  } else if (BN == 64) {
    set_smem(k_fwd_step<64, 8>, Cfg<64, 8>::SMEM_BYTES);
    k_fwd_step<64, 8><<<grid, 192, Cfg<64, 8>::SMEM_BYTES, st>>>(tmH, tmUrk, a);
  } else {
    set_smem(k_fwd_step<32, 8>, Cfg<32, 8>::SMEM_BYTES);
    k_fwd_step<32, 8><<<grid, 192, Cfg<32, 8>::SMEM_BYTES, st>>>(tmH, tmUrk, a);
  }
}

void launch_logits(const CUtensorMap& tmH, const CUtensorMap& tmWmn, const LogitsArgs& a, cudaStream_t st) {
  set_smem(k_logits, Cfg<K3_BN, K3_STAGES>::SMEM_BYTES);
  k_logits<<<a.T * a.Bp / BM, 192, Cfg<K3_BN, K3_STAGES>::SMEM_BYTES, st>>>(tmH, tmWmn, a);
}

void launch_bwd_step(int BN, const CUtensorMap& tmdG, const CUtensorMap& tmUkr, const CUtensorMap& tmdY,
                     const CUtensorMap& tmWnm, const BwdStepArgs& a, cudaStream_t st) {
  dim3 grid(a.N / BN, a.Bp / BM);
  if (BN == 128) {
    set_smem(k_bwd_step<128, 6>, Cfg<128, 6>::SMEM_BYTES);
    k_bwd_step<128, 6><<<grid, 192, Cfg<128, 6>::SMEM_BYTES, st>>>(tmdG, tmUkr, tmdY, tmWnm, a);
  } else if (BN == 64) {
    set_smem(k_bwd_step<64, 8>, Cfg<64, 8>::SMEM_BYTES);
    k_bwd_step<64, 8><<<grid, 192, Cfg<64, 8>::SMEM_BYTES, st>>>(tmdG, tmUkr, tmdY, tmWnm, a);
  } else {
    set_smem(k_bwd_step<32, 8>, Cfg<32, 8>::SMEM_BYTES);
    k_bwd_step<32, 8><<<grid, 192, Cfg<32, 8>::SMEM_BYTES, st>>>(tmdG, tmUkr, tmdY, tmWnm, a);
  }
}

void launch_gemm_nt(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& a, cudaStream_t st) {
  set_smem(k_gemm_nt, Cfg<K6_BN, K6_STAGES>::SMEM_BYTES);
  k_gemm_nt<<<a.tiles_m * a.tiles_n, 192, Cfg<K6_BN, K6_STAGES>::SMEM_BYTES, st>>>(tmA, tmB, a);
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
__global__ void k_build_xt(const int* __restrict__ xs1, __nv_bfloat16* __restrict__ ZT, long ldz, int T, int B, int Bp) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= T * Bp) return;
  const int m = blockIdx.y;
  const int s = col / Bp, b = col - s * Bp;
  const int x = (b < B) ? xs1[(size_t)s * B + b] : -1;
  ZT[(size_t)m * ldz + col] = __float2bfloat16_rn(x == m ? 1.0f : 0.0f);
}
void launch_build_xt(const int* xs1, __nv_bfloat16* ZT, long ldz, int M, int T, int B, int Bp, cudaStream_t st) {
  k_build_xt<<<dim3((T * Bp + 255) / 256, M), 256, 0, st>>>(xs1, ZT, ldz, T, B, Bp);
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
