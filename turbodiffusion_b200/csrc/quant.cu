// a1. Per-128x128-block symmetric INT8 quantisation (reference: turbodiffusion/ops/quant/quant.hpp:86-99,122-164;
// scale layout common/store.hpp:46).  HBM-bound: 2 B read + 1 B written per element.
//
// One CTA (256 threads) per 128x128 tile.  A warp-wide load instruction covers two full 256-byte row
// segments (lane -> 16-byte chunk), so every global access is a fully used 128-byte line; each thread keeps
// 8 independent 16-byte loads in flight (8 rows x 8 elements) before the first use.
#include "common.cuh"
#include "host_common.h"

namespace {

using namespace tdb;

constexpr int kBlk = 128;
constexpr int kThreads = 256;
constexpr int kRowsPerPass = kThreads / 16;      // 16 rows per pass (16 lanes x 16 B cover one 256-byte row segment)
constexpr int kPasses = kBlk / kRowsPerPass;     // 8

// kGelu: the input is the FFN's pre-activation; y = T(gelu_tanh(x)) (nn.GELU(approximate="tanh") on a 16-bit tensor: fp32
// math, one rounding, rcm/networks/wan2pt1.py:375) is quantised instead of x.  gelu_tanh(x) = x * sigmoid(2u),
// u = sqrt(2/pi)(x + 0.044715 x^3), evaluated as x / (1 + 2^z) with ex2.approx + rcp.approx (relative error ~2^-21 everywhere,
// including the negative tail where 1 + tanh(u) cancels).
// sigmoid(2u(x)) = 1 / (1 + 2^z), z = -2 log2(e) sqrt(2/pi) (x + 0.044715 x^3): ex2.approx + rcp.approx (relative error ~2^-21
// everywhere, no cancellation in the negative tail).  Measured alternatives (profiles/r02_microbench_prologue.jsonl, shape A
// 32760 x 8960): this form 0.285 ms; reciprocal by three Newton steps on the FMA pipe instead of the second MUFU 0.371 ms
// (the kernel is bound by issue slots, not by the XU pipe).
__device__ __forceinline__ float sigmoid_2u(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + fast_exp2(x * fmaf(-0.10294324f, x * x, -2.3022082f))));
  return r;
}

template <typename T, bool kGelu = false>
__global__ void __launch_bounds__(kThreads) quant_int8_block128_kernel(const T* __restrict__ x, int8_t* __restrict__ q,
                                                                       float* __restrict__ s, int64_t m, int64_t k,
                                                                       int k_blocks) {
  __shared__ float warp_amax[kThreads / 32];
  const int tid = threadIdx.x;
  const int blk_n = blockIdx.x, blk_m = blockIdx.y;
  const int col = blk_n * kBlk + (tid & 15) * 8;  // first of this thread's 8 consecutive columns
  const int row0 = blk_m * kBlk + (tid >> 4);
  const bool col_ok = col < k;  // k % 8 == 0, so a chunk is entirely in or out

  uint4 raw[kPasses];
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const int64_t row = row0 + p * kRowsPerPass;
    raw[p] = make_uint4(0u, 0u, 0u, 0u);
    if (col_ok && row < m) raw[p] = ldg_nc_v4(x + row * k + col);
  }

  float v[kPasses][8];
  float amax = 1e-8f;
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const uint32_t w[4] = {raw[p].x, raw[p].y, raw[p].z, raw[p].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = F16Traits<T>::lo(w[j]), b = F16Traits<T>::hi(w[j]);
      if (kGelu) {
        a = F16Traits<T>::round(a * sigmoid_2u(a));
        b = F16Traits<T>::round(b * sigmoid_2u(b));
      }
      v[p][2 * j] = a;
      v[p][2 * j + 1] = b;
      amax = fmaxf(amax, fmaxf(fabsf(a), fabsf(b)));
    }
  }
  amax = warp_max(amax);
  if ((tid & 31) == 0) warp_amax[tid >> 5] = amax;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) amax = fmaxf(amax, warp_amax[w]);

  const float r = __fdiv_rn(128.0f, amax);  // IEEE division: codes reproduce on any host (the reference's
                                            // --use_fast_math div.approx differs from this only at .5 ties)
  if (tid == 0) s[static_cast<int64_t>(blk_m) * k_blocks + blk_n] = amax * 0.0078125f;  // amax / 128, exact

#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const int64_t row = row0 + p * kRowsPerPass;
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int a, b;
      asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(a) : "f"(__fmul_rn(v[p][j], r)));
      asm("cvt.rni.sat.s8.f32 %0, %1;" : "=r"(b) : "f"(__fmul_rn(v[p][4 + j], r)));
      lo |= (static_cast<uint32_t>(a) & 0xFFu) << (8 * j);
      hi |= (static_cast<uint32_t>(b) & 0xFFu) << (8 * j);
    }
    if (col_ok && row < m) *reinterpret_cast<uint2*>(q + row * k + col) = make_uint2(lo, hi);
  }
}

}  // namespace

static int quant_impl(const void* x, int x_dtype, int64_t m, int64_t k, int8_t* q, float* s, bool gelu, void* stream) {
  using namespace tdb;
  if (!x || !q || !s) return fail(TDB200_ERR_INVALID_ARG, "quant_int8_block128: null pointer");
  if (m < 0 || k < 0) return fail(TDB200_ERR_INVALID_ARG, "quant_int8_block128: negative size");
  if (m == 0 || k == 0) return TDB200_OK;
  if (k % 8 != 0) return fail(TDB200_ERR_UNSUPPORTED, "quant_int8_block128: k=%lld must be a multiple of 8", (long long)k);
  if (!aligned16(x) || (reinterpret_cast<uintptr_t>(q) & 7u))
    return fail(TDB200_ERR_INVALID_ARG, "quant_int8_block128: x must be 16-byte and q 8-byte aligned");
  if (int rc = require_sm100()) return rc;
  const int64_t kb = cdiv64(k, kBlk), mb = cdiv64(m, kBlk);
  if (mb > 65535) return fail(TDB200_ERR_UNSUPPORTED, "quant_int8_block128: m too large (%lld)", (long long)m);
  dim3 grid(static_cast<unsigned>(kb), static_cast<unsigned>(mb));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_dtype == TDB200_DTYPE_BF16 && !gelu)
    quant_int8_block128_kernel<__nv_bfloat16>
        <<<grid, kThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x), q, s, m, k, static_cast<int>(kb));
  else if (x_dtype == TDB200_DTYPE_FP16 && !gelu)
    quant_int8_block128_kernel<__half><<<grid, kThreads, 0, st>>>(static_cast<const __half*>(x), q, s, m, k,
                                                                  static_cast<int>(kb));
  else if (x_dtype == TDB200_DTYPE_BF16)
    quant_int8_block128_kernel<__nv_bfloat16, true>
        <<<grid, kThreads, 0, st>>>(static_cast<const __nv_bfloat16*>(x), q, s, m, k, static_cast<int>(kb));
  else if (x_dtype == TDB200_DTYPE_FP16)
    quant_int8_block128_kernel<__half, true><<<grid, kThreads, 0, st>>>(static_cast<const __half*>(x), q, s, m, k,
                                                                        static_cast<int>(kb));
  else
    return fail(TDB200_ERR_UNSUPPORTED, "quant_int8_block128: dtype tag %d (only bf16/fp16, like quant.cu:64-67)", x_dtype);
  return check_launch("quant_int8_block128_kernel");
}

extern "C" int tdb200_quant_int8_block128(const void* x, int x_dtype, int64_t m, int64_t k, int8_t* q, float* s,
                                          void* stream) {
  return quant_impl(x, x_dtype, m, k, q, s, false, stream);
}

extern "C" int tdb200_gelu_quant_int8_block128(const void* x, int x_dtype, int64_t m, int64_t k, int8_t* q, float* s,
                                               void* stream) {
  return quant_impl(x, x_dtype, m, k, q, s, true, stream);
}
