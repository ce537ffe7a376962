// shx device math: fixed-point conversions and the kernels' own erf.
//
// The whole library is compiled with --fmad=false -prec-div=true -prec-sqrt=true, so
// every fp32 operation below is a single correctly-rounded IEEE operation in the order
// written.  That is what makes (a) the sequential mode reproduce the reference's g++ -O2
// arithmetic and (b) the batched mode comparable bit for bit with the CPU lock-step
// oracle (tests/test_gpu_batched.py).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace shx {

constexpr int kHeightFracBits = 26;                  // heights are Q5.26 in an int32
constexpr float kHeightScale = 67108864.0f;          // 2^26
constexpr float kHeightInv = 1.490116119384765625e-8f;  // 2^-26
constexpr float kTrackScale = 262144.0f;             // 2^18: tracks are Q13.18 in an int32 (|track| < 8192)
constexpr float kTrackInv = 3.814697265625e-6f;      // 2^-18
constexpr int kTrackLimit = 1 << 30;                 // a discharge track beyond 4096 is reported as overflow
constexpr float kLedgerScale = 4294967296.0f;        // 2^32: sediment ledgers are Q31.32 in an int64

__device__ __forceinline__ float h_to_float(int32_t v) { return (float)v * kHeightInv; }
__device__ __forceinline__ int32_t h_quantize(float h) { return __float2int_rn(h * kHeightScale); }
__device__ __forceinline__ int32_t t_quantize(float v) { return __float2int_rn(v * kTrackScale); }
__device__ __forceinline__ float t_to_float(int32_t v) { return (float)v * kTrackInv; }
// power-of-two scaling is exact in fp32, so this equals llrint((double)v * 2^32)
__device__ __forceinline__ long long l_quantize(float v) { return __float2ll_rn(v * kLedgerScale); }

// exp(y) for y in [-17, 0]: k = rint(y*log2e), r = y - k*ln2 (two-piece), degree-6 Taylor, scale by 2^k.
__device__ __forceinline__ float exp_neg(float y) {
  const float k = rintf(y * 0x1.715476p+0f);
  float r = y - k * 0x1.62e400p-1f;
  r = r - k * 0x1.7f7d1cp-20f;
  float p = 0x1.6c16c2p-10f;
  p = p * r + 0x1.111112p-7f;
  p = p * r + 0x1.555556p-5f;
  p = p * r + 0x1.555556p-3f;
  p = p * r + 0.5f;
  p = p * r + 1.0f;
  p = p * r + 1.0f;
  return p * __int_as_float(((int)k + 127) << 23);
}

// erf(x) for the sediment-capacity term erf(0.4*discharge) (reference cellpool.h:242-244 calls
// libm erf).  Own implementation so that the GPU and the CPU oracle agree to the bit; constants
// and error (max 1.47 ulp against the exact function) come from tools/fit_erf.py.
//   |x| < 0.875 : x + x*q(x^2)
//   |x| < 4     : 1 - exp(-x^2) * g(1/(1+|x|))
//   else        : 1
__device__ __forceinline__ float shx_erff(float x) {
  const float ax = fabsf(x);
  const float t = ax * ax;
  float r;
  if (ax < 0.875f) {
    float q = 0x1.6cae4ap-14f;
    q = q * t + -0x1.aebb60p-11f;
    q = q * t + 0x1.554138p-8f;
    q = q * t + -0x1.b81a4ep-6f;
    q = q * t + 0x1.ce2e8cp-4f;
    q = q * t + -0x1.812744p-2f;
    q = q * t + 0x1.06eba8p-3f;
    r = ax + ax * q;
  } else if (ax < 4.0f) {
    const float z = 1.0f / (ax + 1.0f);
    float g = 0x1.e2ce40p-1f;
    g = g * z + -0x1.9f1d12p+1f;
    g = g * z + 0x1.1f8fc4p+2f;
    g = g * z + -0x1.65195ep+1f;
    g = g * z + 0x1.1f65a8p-2f;
    g = g * z + 0x1.7e0a80p-3f;
    g = g * z + 0x1.25f36ep-1f;
    g = g * z + 0x1.2092dap-1f;
    g = g * z + 0x1.bfc2c6p-17f;
    r = 1.0f - exp_neg(-t) * g;
  } else {
    r = 1.0f;
  }
  return copysignf(r, x);
}

// splitmix64 finaliser: the counter-based replacement for rand() in the spawn (world.h:69)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

}  // namespace shx
