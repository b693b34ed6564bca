// gridbar_bench.cu — stand-alone microbenchmark for round 2: what does one grid barrier among G co-resident CTAs cost on
// the B200, for the arrival/poll layouts the persistent recurrence kernels (tc_persist.cu) can use?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gridbar scripts/gridbar_bench.cu && /tmp/gridbar [G] [rounds]
//
// Every CTA runs `rounds` barriers back to back (one arriving thread, one polling warp, like the kernels); the time per
// barrier is (clock64 at the end - at the start) / rounds on CTA 0.  Variants:
//   0  one counter word: G `red.release.gpu` on it, G warps polling it
//   1  8 counters in ONE 32-byte sector (tc_persist.cu default): CTA i arrives on counter i % 8, 8 lanes poll
//   2  8 counters on 8 separate 128-byte lines (LSTM_PERSIST_SPREAD=1)
//   3  `atom.add.acq_rel.gpu` on one counter; the LAST arriver publishes the round number to 8 replicated flag lines, CTA i
//      polls flag i % 8 (8 pollers per line, no poller ever touches the counter)
//   4  variant 2 with a 64 ns nanosleep between polls
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned atom_acqrel(unsigned* p, unsigned v) {
  unsigned o;
  asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory");
  return o;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void k_bar(unsigned* mem, int variant, int rounds, long long* cycles) {
  const int G = gridDim.x, cta = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned* counters = mem;                 // variants 0,1: words 0..7;  2,4: word 32*i;  3: counter at word 0
  unsigned* flags = mem + 1024;             // variant 3: flag i at word 32*i
  const int stride = (variant == 2 || variant == 4) ? 32 : 1;
  const int slots = variant == 0 ? 1 : 8;
  long long t0 = 0;
  for (int r = 1; r <= rounds + 8; r++) {
    if (r == 9 && threadIdx.x == 0) t0 = clock64();          // 8 warm-up rounds
    if (warp == 1 && lane == 0) {                            // the arriving thread
      if (variant == 3) {
        const unsigned old = atom_acqrel(counters, 1u);
        if (old + 1 == (unsigned)G * (unsigned)r)
          for (int i = 0; i < 8; i++) st_release(flags + 32 * i, (unsigned)r);
      } else {
        red_release(counters + (size_t)(cta % slots) * stride, 1u);
      }
    }
    if (warp == 0) {                                         // the polling warp
      const long long ts = clock64();
      for (;;) {
        bool ok;
        if (variant == 3) {
          const unsigned v = lane == 0 ? ld_acquire(flags + 32 * (cta % 8)) : (unsigned)r;
          ok = __all_sync(0xffffffffu, (int)(v - (unsigned)r) >= 0);
        } else {
          // slot s gets ceil/floor(G/slots) arrivals per round: CTAs with cta % slots == s
          const unsigned per = (unsigned)((G - lane + slots - 1) / slots);
          const unsigned v = lane < slots ? ld_acquire(counters + (size_t)lane * stride) : 0u;
          ok = __all_sync(0xffffffffu, lane >= slots || (int)(v - per * (unsigned)r) >= 0);
        }
        if (ok) break;
        if (variant == 4) __nanosleep(64);
        if (clock64() - ts > 2000000000LL) __trap();
      }
    }
    __syncthreads();
  }
  if (cta == 0 && threadIdx.x == 0) *cycles = clock64() - t0;
}

int main(int argc, char** argv) {
  const int G = argc > 1 ? atoi(argv[1]) : 128, rounds = argc > 2 ? atoi(argv[2]) : 2000;
  unsigned* mem;
  long long* cyc;
  cudaMalloc(&mem, 8192 * sizeof(unsigned));
  cudaMallocManaged(&cyc, sizeof(long long));
  int dev = 0, mhz = 0, sms = 0;
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (G > sms) { printf("G = %d exceeds the %d SMs: the CTAs would not be co-resident\n", G, sms); return 1; }
  cudaFuncSetAttribute(k_bar, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[5] = {"one word", "8 words, one sector", "8 words, 8 lines", "atom + last-arriver flags", "8 lines + nanosleep(64)"};
  for (int v = 0; v < 5; v++) {
    cudaMemset(mem, 0, 8192 * sizeof(unsigned));
    k_bar<<<G, 64, 200 * 1024>>>(mem, v, rounds, cyc);      // 200 KB dynamic smem: one CTA per SM, like the real kernels
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
    printf("variant %d (%-26s): %7.0f cycles = %6.3f us per barrier (G = %d)\n", v, names[v], (double)*cyc / rounds,
           (double)*cyc / rounds / (mhz * 1e-3), G);
  }
  return 0;
}
