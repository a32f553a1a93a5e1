// int_peak.cu -- integer-pipe throughput microbenchmark for sm_100a (B200).
//
// The affine-NW fill is bound by integer issue, not HBM or tensor cores, and
// MEASURED_PEAKS.json has no integer figure, so the roofline denominator for the DP kernels
// is measured here: independent dependency chains of one SASS opcode (8 per thread),
// 8 warps x 4 CTAs per SM, timed with clock64() inside the kernel (cycles) and CUDA events
// outside (seconds).  Prints one JSON object; bench.py embeds it as "int_peak".
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, \
              __LINE__);                                                            \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

constexpr int CHAINS = 8;
constexpr int UNROLL = 16;
constexpr int ITERS = 2048;

// opaque register barrier: keeps the value live and un-foldable without emitting SASS
#define OPAQUE(x) asm volatile("" : "+r"(x))

enum Op {
  OP_IADD = 0,      // IADD3 / VIADD
  OP_LOP3,          // LOP3.LUT
  OP_IMAD,          // IMAD (fma pipe)
  OP_VIMNMX_S32,    // max.s32
  OP_VIMNMX_S16X2,  // VIMNMX.S16x2
  OP_VIBMAX_S16X2,  // VIMNMX.S16x2 with two predicate outputs + 2 predicated VIADD
  OP_VIADDMNMX_S16X2,
  OP_VIMNMX3_S16X2,
  OP_VIADDMNMX_S32,
  OP_VIMNMX3_S32,
  OP_MIX_VIMNMX_IMAD,  // 1 VIMNMX.S16x2 + 1 IMAD per step, independent: dual-pipe issue test
  OP_MIX_LOP_IMAD,
  OP_PRED_VIADD,       // @P VIADD
  OP_SHFL,             // SHFL.UP
  OP_COUNT
};

static const char* kNames[OP_COUNT] = {"iadd",           "lop3",           "imad",
                                       "vimnmx_s32",     "vimnmx_s16x2",   "vibmax_s16x2_2pred_2viadd",
                                       "viaddmnmx_s16x2", "vimnmx3_s16x2", "viaddmnmx_s32",
                                       "vimnmx3_s32",    "mix_vimnmx_imad", "mix_lop3_imad",
                                       "pred_viadd",     "shfl_up"};
// SASS instructions issued per chain step
static const int kInstrPerStep[OP_COUNT] = {1, 1, 1, 1, 1, 3, 1, 1, 1, 1, 2, 2, 1, 1};

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t* out, const uint32_t* in, long long* cycles) {
  uint32_t r[CHAINS], s[CHAINS];
  const uint32_t b = in[threadIdx.x & 31], c = in[32 + (threadIdx.x & 31)];
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) {
    r[k] = in[64 + k] + threadIdx.x;
    s[k] = in[72 + k] ^ threadIdx.x;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int k = 0; k < CHAINS; ++k) {
        if (OP == OP_IADD) {
          r[k] = r[k] + b;
          OPAQUE(r[k]);
        } else if (OP == OP_LOP3) {
          r[k] = (r[k] ^ b) | c;  // one LOP3
          OPAQUE(r[k]);
        } else if (OP == OP_IMAD) {
          r[k] = r[k] * b + c;
          OPAQUE(r[k]);
        } else if (OP == OP_VIMNMX_S32) {
          r[k] = (uint32_t)max((int)r[k], (int)b);
          OPAQUE(r[k]);
        } else if (OP == OP_VIMNMX_S16X2) {
          r[k] = __vmaxs2(r[k], b);
          OPAQUE(r[k]);
        } else if (OP == OP_VIBMAX_S16X2) {
          bool ph, pl;
          r[k] = __vibmax_s16x2(r[k], b, &ph, &pl);
          OPAQUE(r[k]);
          if (ph) s[k] += 0x10;
          OPAQUE(s[k]);
          if (pl) s[k] += 0x1;
          OPAQUE(s[k]);
        } else if (OP == OP_VIADDMNMX_S16X2) {
          r[k] = __viaddmax_s16x2(r[k], b, c);
          OPAQUE(r[k]);
        } else if (OP == OP_VIMNMX3_S16X2) {
          r[k] = __vimax3_s16x2(r[k], b, c);
          OPAQUE(r[k]);
        } else if (OP == OP_VIADDMNMX_S32) {
          r[k] = (uint32_t)__viaddmax_s32((int)r[k], (int)b, (int)c);
          OPAQUE(r[k]);
        } else if (OP == OP_VIMNMX3_S32) {
          r[k] = (uint32_t)__vimax3_s32((int)r[k], (int)b, (int)c);
          OPAQUE(r[k]);
        } else if (OP == OP_MIX_VIMNMX_IMAD) {
          r[k] = __vmaxs2(r[k], b);
          OPAQUE(r[k]);
          s[k] = s[k] * b + c;
          OPAQUE(s[k]);
        } else if (OP == OP_MIX_LOP_IMAD) {
          r[k] = (r[k] ^ b) | c;
          OPAQUE(r[k]);
          s[k] = s[k] * b + c;
          OPAQUE(s[k]);
        } else if (OP == OP_PRED_VIADD) {
          if (b & (1u << k)) r[k] += 0x10;
          OPAQUE(r[k]);
        } else if (OP == OP_SHFL) {
          r[k] = __shfl_up_sync(0xffffffffu, r[k], 1);
        }
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) acc ^= r[k] ^ s[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int sms, uint32_t* d_out, uint32_t* d_in, long long* d_cyc, bool first) {
  const int ctas_per_sm = 4, threads = 256;
  const int grid = sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  bench<OP><<<grid, threads>>>(d_out, d_in, d_cyc);  // warm-up
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f;
  double best_cyc = 0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    bench<OP><<<grid, threads>>>(d_out, d_in, d_cyc);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cyc(grid);
    CK(cudaMemcpy(cyc.data(), d_cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (long long v : cyc) mean += (double)v;
    mean /= grid;
    if (ms < best_ms) {
      best_ms = ms;
      best_cyc = mean;
    }
  }
  // per-SM: ctas_per_sm * 8 warps, each issuing steps * instr warp-instructions
  const double steps = (double)ITERS * UNROLL * CHAINS;
  const double warp_instr_per_sm = steps * kInstrPerStep[OP] * ctas_per_sm * (threads / 32);
  const double ipc_sm = warp_instr_per_sm / best_cyc;  // warp-instr / clk / SM
  const double lane_ops_per_s = warp_instr_per_sm * 32.0 * sms / (best_ms * 1e-3);
  const double mhz = best_cyc / (best_ms * 1e-3) / 1e6;
  printf("%s\n  \"%s\": {\"warp_instr_per_clk_per_sm\": %.3f, \"lane_ops_per_s\": %.4e, "
         "\"ms\": %.3f, \"sm_mhz_effective\": %.0f}",
         first ? "" : ",", kNames[OP], ipc_sm, lane_ops_per_s, best_ms, mhz);
  CK(cudaEventDestroy(e0));
  CK(cudaEventDestroy(e1));
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  uint32_t *d_out, *d_in;
  long long* d_cyc;
  CK(cudaMalloc(&d_out, (size_t)sms * 4 * 256 * sizeof(uint32_t)));
  CK(cudaMalloc(&d_cyc, (size_t)sms * 4 * sizeof(long long)));
  std::vector<uint32_t> h(128);
  for (int i = 0; i < 128; ++i) h[i] = 0x01230456u * (i + 1) + 0x9e3779b9u;
  h[0] |= 1;
  CK(cudaMalloc(&d_in, 128 * sizeof(uint32_t)));
  CK(cudaMemcpy(d_in, h.data(), 128 * sizeof(uint32_t), cudaMemcpyHostToDevice));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"results\": {", prop.name, sms,
         prop.clockRate);
  run<OP_IADD>(sms, d_out, d_in, d_cyc, true);
  run<OP_LOP3>(sms, d_out, d_in, d_cyc, false);
  run<OP_IMAD>(sms, d_out, d_in, d_cyc, false);
  run<OP_VIMNMX_S32>(sms, d_out, d_in, d_cyc, false);
  run<OP_VIMNMX_S16X2>(sms, d_out, d_in, d_cyc, false);
  run<OP_VIBMAX_S16X2>(sms, d_out, d_in, d_cyc, false);
  run<OP_VIADDMNMX_S16X2>(sms, d_out, d_in, d_cyc, false);
  run<OP_VIMNMX3_S16X2>(sms, d_out, d_in, d_cyc, false);
  run<OP_VIADDMNMX_S32>(sms, d_out, d_in, d_cyc, false);
  run<OP_VIMNMX3_S32>(sms, d_out, d_in, d_cyc, false);
  run<OP_MIX_VIMNMX_IMAD>(sms, d_out, d_in, d_cyc, false);
  run<OP_MIX_LOP_IMAD>(sms, d_out, d_in, d_cyc, false);
  run<OP_PRED_VIADD>(sms, d_out, d_in, d_cyc, false);
  run<OP_SHFL>(sms, d_out, d_in, d_cyc, false);
  printf("\n}}\n");
  return 0;
}
