/* C part of the CPU oracle (TEST INFRASTRUCTURE ONLY — never linked into the product).
 *
 * td_w8a8_gemm_f32: restatement of the reference W8A8 GEMM main loop,
 *   turbodiffusion/ops/gemm/kernel.hpp:391-427 (one 128-deep K-block at a time, scales advanced per block) and
 *   turbodiffusion/ops/gemm/utils.hpp:116-121 (`accf += __int2float_rn(acci) * scale`, contracted to one fp32 FMA
 *   under the reference's --use_fast_math build, setup.py:34).
 * The int32 dot of a K-block is exact; fmaf() is correctly rounded, so this reproduces the device result bit for bit.
 *   a_q [m,k] int8, a_s [ceil(m/128), k/128], b_q [n,k] int8, b_s [ceil(n/128), k/128], out [m,n] fp32.
 */
#include <math.h>
#include <stdint.h>

void td_w8a8_gemm_f32(const int8_t* a_q, const float* a_s, const int8_t* b_q, const float* b_s, float* out, int64_t m,
                      int64_t n, int64_t k) {
  const int64_t kb_n = k / 128;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < m; ++i) {
    const int8_t* a = a_q + i * k;
    const float* as = a_s + (i / 128) * kb_n;
    for (int64_t j = 0; j < n; ++j) {
      const int8_t* b = b_q + j * k;
      const float* bs = b_s + (j / 128) * kb_n;
      float acc = 0.0f;
      for (int64_t kb = 0; kb < kb_n; ++kb) {
        int32_t dot = 0;
        for (int t = 0; t < 128; ++t) dot += (int32_t)a[kb * 128 + t] * (int32_t)b[kb * 128 + t];
        const float scale = as[kb] * bs[kb];
        acc = fmaf((float)dot, scale, acc);
      }
      out[i * n + j] = acc;
    }
  }
}
