// Micro-probe for DESIGN.md section 10: how fast is a 4-lane transpose through warp shuffles compared
// with the shared-memory round trip the tile kernel uses between two fused ops?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/shfl_probe tools/probes/shfl_probe.cu
// Prints one JSON line per variant: bytes moved per SM per clock and time per "tile op" equivalent.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 256
#define THREADS 256

// A: what the tile kernel does between ops: 16 x STS.64, barrier, 16 x LDS.64 (conflict-free, other thread's data)
__global__ void __launch_bounds__(THREADS, 3) k_smem_roundtrip(double* out, int iters) {
  extern __shared__ double sm[];
  double v[16];
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = t * 16 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[i * THREADS + t] = v[i];
    __syncthreads();
    const int u = (t + 32 * (it & 7) + 1) & (THREADS - 1);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = sm[i * THREADS + u] + 1.0;
    __syncthreads();
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * THREADS + t] = s;
}

// B: 4x4 block transpose over the 4 lanes of a lane-bit pair: 3 rounds x 4 doubles = 24 SHFL.32 per thread,
// register selection with compile-time indices (rotation trick: round r exchanges block (lane ^ r))
__device__ __forceinline__ double shfl_xor_d(double x, int m) {
  int lo = __double2loint(x), hi = __double2hiint(x);
  lo = __shfl_xor_sync(0xffffffffu, lo, m);
  hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return __hiloint2double(hi, lo);
}
template <int SHIFT>
__global__ void __launch_bounds__(THREADS, 3) k_shfl_transpose(double* out, int iters) {
  double v[4][4];                       // v[c][k]: block destined for the lane whose digit is c
  const int t = threadIdx.x;
  const int me = (t >> SHIFT) & 3;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) v[c][k] = t * 16 + c * 4 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 1; r < 4; ++r) {
      // lane `me` sends block v[me ^ r] to lane me ^ r and receives that lane's block v[me]... into slot me ^ r
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double send = (me ^ r) == 0 ? v[0][k] : (me ^ r) == 1 ? v[1][k] : (me ^ r) == 2 ? v[2][k] : v[3][k];
        double got = shfl_xor_d(send, r << SHIFT);
        if ((me ^ r) == 0) v[0][k] = got; else if ((me ^ r) == 1) v[1][k] = got; else if ((me ^ r) == 2) v[2][k] = got; else v[3][k] = got;
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) v[c][k] += 1.0;
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 4; ++k) s += v[c][k];
  out[blockIdx.x * THREADS + t] = s;
}

// C: raw SHFL.32 throughput (no selects)
__global__ void __launch_bounds__(THREADS, 3) k_shfl_raw(double* out, int iters) {
  int v[24];
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 24; ++i) v[i] = t + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 24; ++i) v[i] = __shfl_xor_sync(0xffffffffu, v[i], 1 + (i & 3)) + 1;
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += v[i];
  out[blockIdx.x * THREADS + t] = s;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount, grid = sms * 3;
  double* out;
  cudaMalloc(&out, sizeof(double) * grid * THREADS);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const size_t smem = 16 * THREADS * sizeof(double);      // 32 KiB, as one tile
  cudaFuncSetAttribute(k_smem_roundtrip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // one iteration = one "tile op" worth of data exchange for 4096 doubles (256 threads x 16)
  const double tile_ops = (double)grid * ITERS;
  float ms = time_ms([&] { k_smem_roundtrip<<<grid, THREADS, smem>>>(out, ITERS); });
  printf("{\"probe\": \"smem_roundtrip_16x64bit\", \"ms\": %.4f, \"ns_per_tile_exchange_per_sm\": %.1f, \"bytes_per_clk_per_sm\": %.1f}\n",
         ms, ms * 1e6 / (tile_ops / sms), 2.0 * 32768.0 * (tile_ops / sms) / (ms * 1e-3 * khz * 1e3));
  ms = time_ms([&] { k_shfl_transpose<0><<<grid, THREADS>>>(out, ITERS); });
  printf("{\"probe\": \"shfl_transpose_lane_bits_0_1\", \"ms\": %.4f, \"ns_per_tile_exchange_per_sm\": %.1f, \"shfl32_per_clk_per_sm\": %.2f}\n",
         ms, ms * 1e6 / (tile_ops / sms), 24.0 * 8 * (tile_ops / sms) / (ms * 1e-3 * khz * 1e3));
  ms = time_ms([&] { k_shfl_transpose<2><<<grid, THREADS>>>(out, ITERS); });
  printf("{\"probe\": \"shfl_transpose_lane_bits_2_3\", \"ms\": %.4f, \"ns_per_tile_exchange_per_sm\": %.1f, \"shfl32_per_clk_per_sm\": %.2f}\n",
         ms, ms * 1e6 / (tile_ops / sms), 24.0 * 8 * (tile_ops / sms) / (ms * 1e-3 * khz * 1e3));
  ms = time_ms([&] { k_shfl_raw<<<grid, THREADS>>>(out, ITERS); });
  printf("{\"probe\": \"shfl_raw_24_per_thread\", \"ms\": %.4f, \"ns_per_tile_exchange_per_sm\": %.1f, \"shfl32_per_clk_per_sm\": %.2f, \"sm_clock_khz\": %d, \"sms\": %d}\n",
         ms, ms * 1e6 / (tile_ops / sms), 24.0 * 8 * (tile_ops / sms) / (ms * 1e-3 * khz * 1e3), khz, sms);
  return 0;
}
