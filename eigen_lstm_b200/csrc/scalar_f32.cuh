// scalar_f32.cuh — the scalar functions of the fp32 path, written to round like the reference's scalar code (no FMA contraction
// where the reference has a separate multiply and add: g++ -O3 on x86-64 baseline does not fuse).  Shared by kernels_f32.cu and
// train_small.cu so that both evaluate bit-identical expressions.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace lstm {

__device__ __forceinline__ float logistic_f(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }  // R/lstm.cc:31-33
__device__ __forceinline__ float tanh_prime_f(float x) { return __fsub_rn(1.0f, __fmul_rn(x, x)); }          // :36-38
__device__ __forceinline__ float logistic_prime_f(float x) { return __fmul_rn(x, __fsub_rn(1.0f, x)); }      // :41-43

// Adagrad, R/lstm.cc:259-272:  m += d*d ; p -= lr * d / sqrtf(m + eps)   (eps added in double, :25,46-48);
// clip > 0 clamps d first (north-star addition; 0 = the reference's behaviour)
__device__ __forceinline__ void adagrad_one(float& p, float d, float& m, float lr, double eps, float clip) {
  if (clip > 0.f) d = fminf(fmaxf(d, -clip), clip);
  m = __fadd_rn(m, __fmul_rn(d, d));
  const float s = sqrtf((float)((double)m + eps));
  p = __fsub_rn(p, __fmul_rn(lr, __fdiv_rn(d, s)));
}

}  // namespace lstm
