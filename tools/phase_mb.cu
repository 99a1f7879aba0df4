// Microbenchmark of the skeleton's phase loops (debug aid, not product): cycles per frame of phase_piece for
// 1..8 active warps per CTA (lane 0 only), piece length 64 vs 1024.
#include <cstdio>
#include <cstdint>
#include "../phonic_b200/csrc/voice.cuh"  // build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -o /tmp/phase_mb tools/phase_mb.cu
using namespace pb;
template <bool UNI>
__global__ void mb(float ratio, uint32_t span, uint32_t pieces, float* out, long long* cyc) {
  if ((threadIdx.x & 31) != 0) return;
  const uint32_t w = threadIdx.x >> 5;
  float s = 0.1f * (w + 1), p = 0.0f, o = 0.0f;
  const PhaseK k = phase_consts(ratio + 0.001f * w);
  uint32_t np = 0;
  bool first = true;
  const long long t0 = clock64();
  for (uint32_t i = 0; i < pieces; ++i) {
    np += UNI ? phase_piece_uniform(s, p, k, span, first, o, 0.0f) : phase_piece<false>(s, p, k, span, first, o, 0.0f);
    first = false;
  }
  const long long t1 = clock64();
  out[blockIdx.x * 8 + w] = s + (float)np + o;
  cyc[blockIdx.x * 8 + w] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8192);
  long long h[8];
  for (int uni = 0; uni < 2; ++uni)
  for (float ratio : {0.243f, 0.7f, 1.7f, 3.3f}) {
    for (uint32_t span : {64u, 1024u}) {
      for (int warps : {1, 2, 4, 8}) {
        const uint32_t frames = 1 << 20, pieces = frames / span;
        for (int rep = 0; rep < 2; ++rep) {
          if (uni) mb<true><<<1, warps * 32>>>(ratio, span, pieces, out, cyc); else mb<false><<<1, warps * 32>>>(ratio, span, pieces, out, cyc);
          cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%s ratio %.3f span %4u warps %d: %.2f cycles/frame (warp 0)\n", uni ? "uniform" : "plain  ", ratio, span, warps, (double)h[0] / frames);
      }
    }
  }
  return 0;
}
