// tie_mix.cu -- developer microbenchmark: the instruction mix of one DP column of the packed fill
// (1 LOP3, 5 VIMNMX.U16x2 of which 4 with predicate outputs, 2 integer adds, 8 predicated
// tie-bit sets) with the bit-sets issued as integer adds, float adds, or half and half.
// Question: do predicated FADDs relieve the integer pipes?   nvcc -arch=sm_100a -O3 tie_mix.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int MODE>
__device__ __forceinline__ uint32_t vmax_tie(uint32_t a, uint32_t b, uint32_t& ia, uint32_t& ib, float& fa, float& fb, uint32_t IBIT) {
  uint32_t r;
  if (MODE == 0) {
    asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, a0, a1;\n\t"
        "max.u16x2 %0, %3, %4;\n\tmov.b32 {r0, r1}, %0;\n\tmov.b32 {a0, a1}, %3;\n\t"
        "setp.eq.u16 pl, r0, a0;\n\tsetp.eq.u16 ph, r1, a1;\n\t"
        "@pl add.u32 %1, %1, %5;\n\t@ph add.u32 %2, %2, %5;\n\t}"
        : "=r"(r), "+r"(ia), "+r"(ib) : "r"(a), "r"(b), "r"(IBIT));
  } else {
    asm("{\n\t.reg .pred ph, pl;\n\t.reg .u16 r0, r1, a0, a1;\n\t"
        "max.u16x2 %0, %3, %4;\n\tmov.b32 {r0, r1}, %0;\n\tmov.b32 {a0, a1}, %3;\n\t"
        "setp.eq.u16 pl, r0, a0;\n\tsetp.eq.u16 ph, r1, a1;\n\t"
        "@pl add.f32 %1, %1, 0f41800000;\n\t@ph add.f32 %2, %2, 0f41800000;\n\t}"
        : "=r"(r), "+f"(fa), "+f"(fb) : "r"(a), "r"(b));
  }
  return r;
}

// FMODE: 0 all integer, 1 compares 2,3 float, 2 all float
template <int FMODE>
__global__ void __launch_bounds__(32) mix(uint32_t* out, const uint32_t* in, int iters) {
  constexpr int K = 8;
  uint32_t H[K], F[K], q[K];
  for (int c = 0; c < K; ++c) { H[c] = in[c] + threadIdx.x; F[c] = in[8 + c]; q[c] = in[16 + c]; }
  const uint32_t pen = in[24], cm = in[25], open = in[26];
  uint32_t d = in[27], hd = in[28], E = in[29];
  uint32_t ia = 0, ib = 0;
  float fa = 8388608.0f, fb = 8388608.0f;
  for (int it = 0; it < iters; ++it) {
    uint32_t hdiag = hd;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      const uint32_t hup = H[c];
      const uint32_t m = __vminu2(q[c] ^ d, pen);
      const uint32_t M = hdiag + cm - m;
      const uint32_t t = vmax_tie<(FMODE == 2)>(E, M, ia, ib, fa, fb, 1u << (c * 4));
      const uint32_t Hn = vmax_tie<(FMODE == 2)>(F[c], t, ia, ib, fa, fb, 2u << (c * 4));
      const uint32_t Mo = M - open;
      E = vmax_tie<(FMODE >= 1)>(Mo, E, ia, ib, fa, fb, 4u << (c * 4));
      F[c] = vmax_tie<(FMODE >= 1)>(Mo, F[c], ia, ib, fa, fb, 8u << (c * 4));
      H[c] = Hn;
      hdiag = hup;
    }
    d = d * 5 + 1;
    hd = H[K - 1];
  }
  uint32_t acc = ia ^ ib ^ __float_as_uint(fa) ^ __float_as_uint(fb) ^ E;
  for (int c = 0; c < K; ++c) acc ^= H[c] ^ F[c];
  out[blockIdx.x * 32 + threadIdx.x] = acc;
}

template <int FMODE>
void run(uint32_t* d_out, uint32_t* d_in, int sms, int per_sm = 14) {
  const int iters = 20000, grid = sms * per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix<FMODE><<<grid, 32>>>(d_out, d_in, iters);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    mix<FMODE><<<grid, 32>>>(d_out, d_in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double cols = (double)grid * iters * 8;
  printf("warps/SM %d fmode %d: %.3f ms, %.2f G column-steps/s (x2 cells), %.2f T lane-instr/s at 16 instr/column\n", per_sm, FMODE, best,
         cols / best / 1e6, cols * 16 * 32 / best / 1e9);
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  uint32_t *d_out, *d_in;
  cudaMalloc(&d_out, (size_t)prop.multiProcessorCount * 32 * 32 * 4);
  uint32_t h[32];
  for (int i = 0; i < 32; ++i) h[i] = 0x4000u + 37u * i + ((0x4100u + 11u * i) << 16);
  h[24] = 18 | (18 << 16); h[25] = 34 | (34 << 16); h[26] = 16 | (16 << 16);
  cudaMalloc(&d_in, sizeof(h));
  cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int w : {8, 12, 14, 16, 20, 24, 32}) run<0>(d_out, d_in, prop.multiProcessorCount, w);
  run<1>(d_out, d_in, prop.multiProcessorCount);
  run<2>(d_out, d_in, prop.multiProcessorCount);
  return 0;
}
