#include <cstdio>
#include "msq_quant_core.cuh"
using namespace msq;
__global__ void k(const float* in, float* out) {
  float x[16], t[16];
  for (int i = 0; i < 16; ++i) x[i] = in[i];
#pragma unroll
  for (int i = 0; i + 1 < 16; i += 2) t[i] = __fadd_rn(fabsf(x[i]), fabsf(x[i + 1]));
  const float s = tree_sum_regs2<16, 2>(t);
  const float mean = __fmul_rn(s, 1.0f / 16);
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float d = __fsub_rn(fabsf(x[i]), mean); t[i] = __fmul_rn(d, d); }
  const float q = tree_sum_regs2<16>(t);
  const float sd = __fsqrt_rn(__fmul_rn(q, 1.0f / 16));
  const float w = __fmul_rn(2.0f, sd);
  out[0] = s; out[1] = mean; out[2] = q; out[3] = sd; out[4] = __fsub_rn(mean, w);
  // scalar tree
  float u[16];
  for (int i = 0; i < 16; ++i) u[i] = fabsf(x[i]);
  for (int st = 1; st < 16; st <<= 1) for (int i = 0; i + st < 16; i += 2 * st) u[i] = __fadd_rn(u[i], u[i + st]);
  out[5] = u[0];
}
int main() {
  float h[16] = {-0.012943864800035954f, 0.007606775965541601f, -0.006049501709640026f, -0.003153566736727953f, 0.0262801144272089f, -0.02341817505657673f, -0.028241610154509544f, 0.025167010724544525f, 0.007884628139436245f, 0.06395717710256577f, -0.030391091480851173f, 0.02073526754975319f, -0.0056556290946900845f, 0.0020561297424137592f, 0.0028521972708404064f, -0.03499250486493111f};
  float *d, *o, r[6];
  cudaMalloc(&d, 64); cudaMalloc(&o, 24);
  cudaMemcpy(d, h, 64, cudaMemcpyHostToDevice);
  k<<<1, 1>>>(d, o);
  cudaMemcpy(r, o, 24, cudaMemcpyDeviceToHost);
  printf("s %.9g mean %.9g q %.9g sd %.9g l %.9g scalar_s %.9g x0 %.9g\n", r[0], r[1], r[2], r[3], r[4], r[5], h[0]);
  return 0;
}
