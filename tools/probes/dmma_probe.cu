// DMMA probe with a kill criterion (BASELINE north_star: "tensor cores are used only if fused k-qubit blocks are
// applied as dense fp64 DMMA products and ncu shows that wins").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/dmma_probe tools/probes/dmma_probe.cu
//
// The op phase of k_tile_pass6 applies  v <- (pa (x) pb) v  to every 16-block of a 4096-coefficient tile: 2 x 4
// products of a 3x3 (or 3x4) real matrix with a 4-vector per block = 72 DFMA per block, 18 432 per tile = 576 warp
// instructions on the FP64 pipe.  The only fp64 tensor-core shape is mma.sync.m8n8k4 (D[8x8] += A[8x4] B[4x8]):
// a 4x4 map fills HALF of the 8-row operand, so one DMMA transforms 8 four-vectors with 256 FMA slots of which 128
// do wanted work (72 essential); a tile needs 2 maps x 128 = 256 DMMAs.
//
// Measured here, all with 3 CTAs x 256 threads per SM like the tile kernel:
//   dfma_peak   independent DFMA chains                                    -> FP64 pipe rate
//   dmma_peak   independent mma.sync.m8n8k4.f64 accumulators               -> tensor fp64 rate
//   map_dfma    the tile kernel's arithmetic for both maps of a block from registers (72 DFMA per 16-block)
//   map_dmma    the same result for 8 four-vectors per DMMA, operands ALREADY in fragment layout (best case for
//               DMMA: no shuffles, no shared-memory re-layout)
// Prints one JSON line per kernel with ns per 4096-coefficient tile-op per SM.  Kill criterion: map_dmma must beat
// 345 ns per tile-op per SM (what the whole op, shared-memory round trip included, costs today).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define THREADS 256
#define CTAS 3

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(THREADS, CTAS) k_dfma_peak(double* out, int iters, double m) {
  double v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fma(v[i], m, 1.0);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * THREADS + threadIdx.x] = s;
}

__global__ void __launch_bounds__(THREADS, CTAS) k_dmma_peak(double* out, int iters, double m) {
  double d[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { d[i][0] = threadIdx.x + i; d[i][1] = i; }
  const double a = m, b = 1.0 / (1 + (threadIdx.x & 3));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(d[i][0], d[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1];
  out[blockIdx.x * THREADS + threadIdx.x] = s;
}

// both maps of one 16-block, the tile kernel's statements (maps without I admixture: 9 DFMA per 3-vector)
__device__ __forceinline__ void mat3(const double* m, double& x1, double& x2, double& x3) {
  const double y1 = m[0] * x1 + m[1] * x2 + m[2] * x3;
  const double y2 = m[3] * x1 + m[4] * x2 + m[5] * x3;
  const double y3 = m[6] * x1 + m[7] * x2 + m[8] * x3;
  x1 = y1; x2 = y2; x3 = y3;
}
struct maps { double pa[9], pb[9]; };
__global__ void __launch_bounds__(THREADS, CTAS) k_map_dfma(double* out, int iters, const __grid_constant__ maps M) {
  double v[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = 1e-3 * (threadIdx.x + 4 * i + j);
  for (int it = 0; it < iters; ++it) {        // one iteration = one 16-block = 1/256 of a tile-op
#pragma unroll
    for (int j = 0; j < 4; ++j) mat3(M.pa, v[1][j], v[2][j], v[3][j]);
#pragma unroll
    for (int i = 0; i < 4; ++i) mat3(M.pb, v[i][1], v[i][2], v[i][3]);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += v[i][j];
  out[blockIdx.x * THREADS + threadIdx.x] = s;
}

// DMMA form, best case: a warp holds 8 x 8 four-vectors as B fragments (one double per thread per fragment);
// map a:  D = [pa; 0] (8x4) x X (4x8), then the D fragment (2 doubles, rows = lane / 4) is taken as the next B
// fragment without any re-layout (a real kernel would need one: D is row-distributed, B is k-distributed).
// One iteration = 8 DMMAs = 64 four-vectors through ONE map = 8 DMMAs x 2 maps / (2 x 128) -> 1/16 of a tile-op
// per warp, i.e. per 8-warp CTA iteration = half a tile-op.
__global__ void __launch_bounds__(THREADS, CTAS) k_map_dmma(double* out, int iters, const __grid_constant__ maps M) {
  const int lane = threadIdx.x & 31;
  const int r = lane >> 2, k = lane & 3;
  // A fragment: row r of [I + map; 0], column k
  const double a = r == 0 ? (k == 0 ? 1.0 : 0.0) : (r < 4 && k > 0 ? M.pa[(r - 1) * 3 + (k - 1)] : 0.0);
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 1e-3 * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double d0 = 0.0, d1 = 0.0;
      dmma884(d0, d1, a, x[i]);
      x[i] = d0 + d1;                        // stand-in for the (free, in this best case) D -> B re-layout
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * THREADS + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch();
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, grid = sms * CTAS;
  double* out;
  cudaMalloc(&out, sizeof(double) * grid * THREADS);
  const int iters = 4096;
  maps M;
  for (int i = 0; i < 9; ++i) { M.pa[i] = (i % 4 == 0) ? 0.999 : 1e-3 * i; M.pb[i] = (i % 4 == 0) ? 0.998 : -1e-3 * i; }
  const double threads_total = (double)grid * THREADS;

  float ms = time_ms([&] { k_dfma_peak<<<grid, THREADS>>>(out, iters, 0.999999); });
  const double dfma_tflops = 2.0 * 16 * iters * threads_total / (ms * 1e-3) / 1e12;
  printf("{\"kernel\": \"dfma_peak\", \"ms\": %.4f, \"tflops\": %.2f, \"sms\": %d}\n", ms, dfma_tflops, sms);

  ms = time_ms([&] { k_dmma_peak<<<grid, THREADS>>>(out, iters, 0.999999); });
  const double dmma_per_s = 8.0 * iters * (threads_total / 32) / (ms * 1e-3);
  printf("{\"kernel\": \"dmma_peak\", \"ms\": %.4f, \"tflops\": %.2f, \"ns_per_dmma_per_sm\": %.3f}\n", ms,
         dmma_per_s * 512 / 1e12, 1e9 / (dmma_per_s / sms));

  // map_dfma: CTAS x 256 threads per SM, each iteration = one 16-block; a tile-op = 256 blocks
  ms = time_ms([&] { k_map_dfma<<<grid, THREADS>>>(out, iters, M); });
  const double blocks_per_sm = (double)CTAS * THREADS * iters;
  printf("{\"kernel\": \"map_dfma\", \"ms\": %.4f, \"ns_per_tile_op_per_sm\": %.1f, \"note\": \"72 DFMA per 16-block, registers only\"}\n",
         ms, ms * 1e6 / (blocks_per_sm / 256.0));

  // map_dmma: per warp and iteration 8 DMMAs = 64 four-vectors through one map; a tile-op = 2 x 1024 four-vectors
  ms = time_ms([&] { k_map_dmma<<<grid, THREADS>>>(out, iters, M); });
  const double vecs_per_sm = (double)CTAS * (THREADS / 32) * iters * 64.0;
  printf("{\"kernel\": \"map_dmma\", \"ms\": %.4f, \"ns_per_tile_op_per_sm\": %.1f, \"note\": \"256 DMMA per tile-op, fragments already in place (best case)\", \"kill_if_above_ns\": 345}\n",
         ms, ms * 1e6 / (vecs_per_sm / 2048.0));
  cudaFree(out);
  return 0;
}
