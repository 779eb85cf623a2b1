// Probe for the fused-pair ("triple") op phase: does keeping a digit TRIPLE's 64 coefficients in registers
// for two consecutive ops that share one digit -- one shared-memory round trip for two ops -- beat two passes of
// the shipped 16-block body?  (ncu, profiles/r02_tile_pass.md: the shipped kernel runs at 73-80 % of the
// shared-memory wavefront peak; halving the round trips is the only lever left on that pipe.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -Iqiskit-aakash_b200/csrc \
//        -o /tmp/pair_probe tools/probes/pair_probe.cu -ldl
//   /tmp/pair_probe qiskit-aakash_b200/libdmb200.so
// Baseline = the library's own k_tile_pass6 through the C ABI on the same pass; the probe kernel applies the same
// ops as fused pairs (64 threads per tile, 64 coefficients per thread) and the two results are compared.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "dm_device.h"

struct smem_mem {
  uint32_t base;
  __device__ __forceinline__ double ld64(uint32_t off) const {
    double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + off) : "memory"); return v; }
  __device__ __forceinline__ dmb_d2 ld128(uint32_t off) const {
    dmb_d2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + off) : "memory"); return v; }
  __device__ __forceinline__ void st64(uint32_t off, double v) const {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + off), "d"(v) : "memory"); }
  __device__ __forceinline__ void st128(uint32_t off, dmb_d2 v) const {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(base + off), "d"(v.x), "d"(v.y) : "memory"); }
};

// per fused pair: canonical block u[p][q][r], p = value of op1's digit a, q = of op1's digit b, r = of op2's other digit
struct pair_info {
  uint32_t sr[4];      // swizzled byte offsets of the r digit's values
  uint32_t sh[3];      // shifts placing the thread's three index digits on the free tile digits
  int32_t ld_mode;     // 0: 64-bit accesses; 1 / 2 / 3: tile digit 0 is p / q / r -> 128-bit accesses along it
  int32_t m2;          // bit 0: the shared digit is q (else p); bit 1: op2's control (digit a) is the shared digit
  int32_t pad_[3];
};
struct probe_params {
  int32_t n_pairs;
  int32_t pad_[3];
  pair_info pi[DMB_MAX_OPS / 2];
};

template <int LD>
__device__ __forceinline__ void pair_load(uint32_t sb, const dmb_lean_op& o1, const pair_info& pi, const smem_mem& mem,
                                          double (&u)[4][4][4]) {
  const uint32_t* sp = o1.sa;
  const uint32_t* sq = o1.sj;
  if constexpr (LD == 0) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t a = sb ^ sp[p] ^ sq[q];
#pragma unroll
        for (int r = 0; r < 4; ++r) u[p][q][r] = mem.ld64(a ^ pi.sr[r]);
      }
  } else if constexpr (LD == 1) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t a = sb ^ sq[q] ^ pi.sr[r];
        const dmb_d2 x = mem.ld128(a), y = mem.ld128(a ^ 16u);
        u[0][q][r] = x.x; u[1][q][r] = x.y; u[2][q][r] = y.x; u[3][q][r] = y.y;
      }
  } else if constexpr (LD == 2) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t a = sb ^ sp[p] ^ pi.sr[r];
        const dmb_d2 x = mem.ld128(a), y = mem.ld128(a ^ 16u);
        u[p][0][r] = x.x; u[p][1][r] = x.y; u[p][2][r] = y.x; u[p][3][r] = y.y;
      }
  } else {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t a = sb ^ sp[p] ^ sq[q];
        const dmb_d2 x = mem.ld128(a), y = mem.ld128(a ^ 16u);
        u[p][q][0] = x.x; u[p][q][1] = x.y; u[p][q][2] = y.x; u[p][q][3] = y.y;
      }
  }
}

template <int LD>
__device__ __forceinline__ void pair_store(uint32_t sb, const dmb_lean_op& o1, const pair_info& pi, const smem_mem& mem,
                                           double (&u)[4][4][4]) {
  const uint32_t* sp = o1.sa;
  const uint32_t* sq = o1.sj;
  if constexpr (LD == 0) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t a = sb ^ sp[p] ^ sq[q];
#pragma unroll
        for (int r = 0; r < 4; ++r) mem.st64(a ^ pi.sr[r], u[p][q][r]);
      }
  } else if constexpr (LD == 1) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t a = sb ^ sq[q] ^ pi.sr[r];
        dmb_d2 x, y;
        x.x = u[0][q][r]; x.y = u[1][q][r]; y.x = u[2][q][r]; y.y = u[3][q][r];
        mem.st128(a, x); mem.st128(a ^ 16u, y);
      }
  } else if constexpr (LD == 2) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t a = sb ^ sp[p] ^ pi.sr[r];
        dmb_d2 x, y;
        x.x = u[p][0][r]; x.y = u[p][1][r]; y.x = u[p][2][r]; y.y = u[p][3][r];
        mem.st128(a, x); mem.st128(a ^ 16u, y);
      }
  } else {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t a = sb ^ sp[p] ^ sq[q];
        dmb_d2 x, y;
        x.x = u[p][q][0]; x.y = u[p][q][1]; y.x = u[p][q][2]; y.y = u[p][q][3];
        mem.st128(a, x); mem.st128(a ^ 16u, y);
      }
  }
}

// op1 on (p, q) for every r, then op2 on (shared, r)
template <int KX, int MC, int M2>
__device__ __forceinline__ void pair_math(const dmb_lean_op& o1, const dmb_lean_op& o2, double (&u)[4][4][4]) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    double w[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[i][j] = u[i][j][r];
    dmb_spec_math<KX, MC, MC>(o1, w);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) u[i][j][r] = w[i][j];
  }
  constexpr bool SHQ = (M2 & 1) != 0, CTL_SH = (M2 & 2) != 0;
#pragma unroll
  for (int o = 0; o < 4; ++o) {          // o: value of the digit op2 does not touch (p if the shared digit is q)
    double w[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int s = CTL_SH ? i : j, r = CTL_SH ? j : i;       // i: op2's control value, j: its target value
        w[i][j] = SHQ ? u[o][s][r] : u[s][o][r];
      }
    dmb_spec_math<KX, MC, MC>(o2, w);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int s = CTL_SH ? i : j, r = CTL_SH ? j : i;
        if (SHQ) u[o][s][r] = w[i][j]; else u[s][o][r] = w[i][j];
      }
  }
}

template <int KX, int MC>
__device__ __forceinline__ void pair_phase(uint32_t t, const dmb_lean_op& o1, const dmb_lean_op& o2, const pair_info& pi,
                                           const smem_mem& mem) {
  const uint32_t bl = ((t & 3u) << pi.sh[0]) | (((t >> 2) & 3u) << pi.sh[1]) | (((t >> 4) & 3u) << pi.sh[2]);
  const uint32_t sb = dmb_swz(bl) << 3;
  double u[4][4][4];
  switch (pi.ld_mode) {
    case 0: pair_load<0>(sb, o1, pi, mem, u); break;
    case 1: pair_load<1>(sb, o1, pi, mem, u); break;
    case 2: pair_load<2>(sb, o1, pi, mem, u); break;
    default: pair_load<3>(sb, o1, pi, mem, u); break;
  }
  switch (pi.m2) {
    case 0: pair_math<KX, MC, 0>(o1, o2, u); break;
    case 1: pair_math<KX, MC, 1>(o1, o2, u); break;
    case 2: pair_math<KX, MC, 2>(o1, o2, u); break;
    default: pair_math<KX, MC, 3>(o1, o2, u); break;
  }
  switch (pi.ld_mode) {
    case 0: pair_store<0>(sb, o1, pi, mem, u); break;
    case 1: pair_store<1>(sb, o1, pi, mem, u); break;
    case 2: pair_store<2>(sb, o1, pi, mem, u); break;
    default: pair_store<3>(sb, o1, pi, mem, u); break;
  }
}

template <int MAXREG>
__global__ void __launch_bounds__(64) __maxnreg__(MAXREG)
k_pair_probe(double* __restrict__ state, const __grid_constant__ dmb_lean_pass L, const __grid_constant__ probe_params Q) {
  extern __shared__ __align__(128) unsigned char sm[];
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(sm);
  smem_mem mem;
  mem.base = smem0;
  const uint32_t t = threadIdx.x;
  dmb_lean_thread T0;
  dmb_lean_thread_init((int)t, L, T0);
  // virtual threads t + 64 h: offsets differ by uniform terms (2 (t + 64 h) = 2 t | 128 h, disjoint bits)
  uint64_t vg[4];
  uint32_t vs[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    vg[h] = dmb_tile_off(128u * h, L.td, DMB_LEAN_K);
    vs[h] = dmb_swz(128u * h) << 3;
  }
  for (uint64_t tile = blockIdx.x; tile < L.n_tiles; tile += gridDim.x) {
    const uint64_t tb = dmb_tile_base(tile, L.td, DMB_LEAN_K);
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
      for (int i = 0; i < DMB_LEAN_PAIRS; ++i) {
        const double* src = state + tb + (T0.goff | vg[h] | L.pair_goff[i]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem0 + (T0.soff ^ vs[h] ^ L.pair_soff[i])), "l"(src) : "memory");
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int k = 0; k < Q.n_pairs; ++k) {
      pair_phase<DMB_KIND_TSP0, 1>(t, L.ops[2 * k], L.ops[2 * k + 1], Q.pi[k], mem);
      __syncthreads();
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      dmb_d2 w[DMB_LEAN_PAIRS];
#pragma unroll
      for (int i = 0; i < DMB_LEAN_PAIRS; ++i) w[i] = mem.ld128(T0.soff ^ vs[h] ^ L.pair_soff[i]);
#pragma unroll
      for (int i = 0; i < DMB_LEAN_PAIRS; ++i)
        *reinterpret_cast<dmb_d2*>(state + tb + (T0.goff | vg[h] | L.pair_goff[i])) = w[i];
    }
    __syncthreads();
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------
static void lane_order(int a, int b, int8_t* fd) {        // schedule.lane_order for K = 6
  int free_d[4], nf = 0;
  for (int d = 0; d < 6; ++d) if (d != a && d != b) free_d[nf++] = d;
  auto has = [&](int d) { for (int i = 0; i < nf; ++i) if (free_d[i] == d) return true; return false; };
  int order[4], w = 0;
  if (has(0)) order[w++] = 0;
  else {
    const int o1 = has(1) ? 1 : (has(2) ? 2 : free_d[0]);
    order[w++] = o1;
    int o2 = -1;
    if (has(3) && o1 != 3) o2 = 3; else if (has(5) && o1 != 5) o2 = 5;
    if (o2 < 0) for (int i = 0; i < nf; ++i) if (free_d[i] != o1) { o2 = free_d[i]; break; }
    order[w++] = o2;
  }
  for (int i = 0; i < nf; ++i) { bool used = false; for (int k = 0; k < w; ++k) if (order[k] == free_d[i]) used = true; if (!used) order[w++] = free_d[i]; }
  for (int i = 0; i < 4; ++i) fd[i] = (int8_t)order[i];
}

static bool triple_lanes(int x, int y, int z, uint32_t* sh) {      // lane order of the three free digits
  int free_d[3], nf = 0;
  for (int d = 0; d < 6; ++d) if (d != x && d != y && d != z) free_d[nf++] = d;
  auto has = [&](int d) { return free_d[0] == d || free_d[1] == d || free_d[2] == d; };
  int order[3];
  if (has(0)) {
    order[0] = 0;
    int w = 1;
    for (int i = 0; i < 3; ++i) if (free_d[i] != 0) order[w++] = free_d[i];
  } else {
    const int o1 = has(1) ? 1 : (has(2) ? 2 : (has(4) ? 4 : free_d[0]));
    int o2 = -1;
    if (has(3) && o1 != 3) o2 = 3; else if (has(5) && o1 != 5) o2 = 5;
    if (o2 < 0) return false;
    order[0] = o1; order[1] = o2;
    for (int i = 0; i < 3; ++i) if (free_d[i] != o1 && free_d[i] != o2) order[2] = free_d[i];
  }
  for (int i = 0; i < 3; ++i) sh[i] = 2u * order[i];
  return true;
}

typedef int (*create_fn)(int, void**);
typedef int (*apply_fn)(void*, double*, int, const dmb_pass*, size_t);
typedef int (*sync_fn)(void*);
typedef const char* (*err_fn)(void);

template <int MAXREG>
static float run_probe(double* d_state, const dmb_lean_pass& L, const probe_params& Q, int sms, int ctas, int reps) {
  cudaFuncSetAttribute(k_pair_probe<MAXREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, DMB_LEAN_TILE_BYTES);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) k_pair_probe<MAXREG><<<sms * ctas, 64, DMB_LEAN_TILE_BYTES>>>(d_state, L, Q);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) k_pair_probe<MAXREG><<<sms * ctas, 64, DMB_LEAN_TILE_BYTES>>>(d_state, L, Q);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main(int argc, char** argv) {
  const char* libpath = argc > 1 ? argv[1] : "qiskit-aakash_b200/libdmb200.so";
  void* lib = dlopen(libpath, RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen %s failed: %s\n", libpath, dlerror()); return 1; }
  create_fn dmb_create = (create_fn)dlsym(lib, "dmb_create");
  apply_fn dmb_apply_passes = (apply_fn)dlsym(lib, "dmb_apply_passes");
  sync_fn dmb_sync = (sync_fn)dlsym(lib, "dmb_sync");
  err_fn dmb_last_error = (err_fn)dlsym(lib, "dmb_last_error");
  void* ctx = nullptr;
  if (dmb_create(0, &ctx)) { fprintf(stderr, "dmb_create: %s\n", dmb_last_error()); return 1; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int n_bits = 28;
  const size_t count = 1ull << n_bits;
  std::vector<double> h(count);
  srand(7);
  for (size_t i = 0; i < count; ++i) h[i] = (rand() / (double)RAND_MAX - 0.5) * 1e-3;
  double *d_ref, *d_new;
  cudaMalloc(&d_ref, count * 8); cudaMalloc(&d_new, count * 8);
  std::vector<double> out_ref(count), out_new(count);

  // scenarios: chains of op pairs sharing one digit; {a1,b1,a2,b2} per pair
  struct scen { const char* name; int n_pairs; int d[8][4]; };
  const scen scens[] = {
      {"digit0_free_chain", 5, {{1, 2, 2, 3}, {3, 4, 4, 5}, {2, 1, 3, 2}, {4, 3, 5, 4}, {1, 2, 3, 2}}},
      {"digit0_in_triples", 5, {{0, 1, 1, 2}, {2, 3, 0, 3}, {1, 0, 4, 1}, {5, 4, 4, 0}, {0, 2, 2, 1}}},
      {"mixed_brickwall", 5, {{0, 1, 1, 2}, {3, 4, 4, 5}, {2, 3, 1, 2}, {4, 5, 3, 4}, {0, 1, 2, 1}}},
  };
  for (const scen& S : scens) {
    dmb_pass P;
    memset(&P, 0, sizeof(P));
    P.n_tile_digits = 6;
    for (int j = 0; j < 6; ++j) P.tile_digit[j] = j == 0 ? 0 : (j == 1 ? 1 : 2 * j);     // digits {0,1,4,6,8,10}
    P.n_ops = 2 * S.n_pairs;
    probe_params Q;
    memset(&Q, 0, sizeof(Q));
    Q.n_pairs = S.n_pairs;
    bool ok = true;
    for (int k = 0; k < P.n_ops; ++k) {
      dmb_op& o = P.ops[k];
      o.kind = DMB_OP_CX_TSP;
      o.flags = DMB_HAS_PA | DMB_HAS_PB;
      o.a = (int8_t)S.d[k / 2][(k & 1) * 2];
      o.b = (int8_t)S.d[k / 2][(k & 1) * 2 + 1];
      lane_order(o.a, o.b, o.fd);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) {
          o.pa[4 * r + c] = c == 0 ? 0.0 : (c == r + 1 ? 0.97 : 0.05 * (r - c + 0.5 * k));
          o.pb[4 * r + c] = c == 0 ? 0.0 : (c == r + 1 ? 0.96 : -0.04 * (r + c - 0.25 * k));
        }
      const double cav = 0.999, c2av = 4 * cav - 3;
      o.coef[0] = cav; o.coef[1] = 0.0; o.coef[2] = 0.5 * (1 + c2av); o.coef[3] = 0.5 * (1 - c2av); o.coef[4] = 0.0;
    }
    static dmb_lean_pass L;
    dmb_make_lean_pass(P, n_bits, L, false);
    for (int k = 0; k < S.n_pairs; ++k) {
      const dmb_op &o1 = P.ops[2 * k], &o2 = P.ops[2 * k + 1];
      pair_info& pi = Q.pi[k];
      const bool a2_shared = (o2.a == o1.a || o2.a == o1.b);
      const int shared = a2_shared ? o2.a : o2.b;
      const int other = a2_shared ? o2.b : o2.a;
      pi.m2 = (shared == o1.b ? 1 : 0) | (a2_shared ? 2 : 0);
      for (int r = 0; r < 4; ++r) pi.sr[r] = dmb_swz((uint32_t)r << (2 * other)) << 3;
      pi.ld_mode = o1.a == 0 ? 1 : (o1.b == 0 ? 2 : (other == 0 ? 3 : 0));
      if (!triple_lanes(o1.a, o1.b, other, pi.sh)) ok = false;
    }
    if (!ok) { printf("{\"scenario\": \"%s\", \"error\": \"no conflict-free lane order\"}\n", S.name); continue; }
    // baseline: the library
    cudaMemcpy(d_ref, h.data(), count * 8, cudaMemcpyHostToDevice);
    if (dmb_apply_passes(ctx, d_ref, n_bits, &P, 1)) { fprintf(stderr, "apply: %s\n", dmb_last_error()); return 1; }
    dmb_sync(ctx);
    cudaMemcpy(out_ref.data(), d_ref, count * 8, cudaMemcpyDeviceToHost);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) dmb_apply_passes(ctx, d_ref, n_bits, &P, 1);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms_ref;
    cudaEventElapsedTime(&ms_ref, a, b);
    ms_ref /= 10;
    // fused
    cudaMemcpy(d_new, h.data(), count * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_pair_probe<200>, cudaFuncAttributeMaxDynamicSharedMemorySize, DMB_LEAN_TILE_BYTES);
    k_pair_probe<200><<<prop.multiProcessorCount * 5, 64, DMB_LEAN_TILE_BYTES>>>(d_new, L, Q);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "probe kernel: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out_new.data(), d_new, count * 8, cudaMemcpyDeviceToHost);
    double worst = 0.0, big = 0.0;
    for (size_t i = 0; i < count; ++i) { worst = fmax(worst, fabs(out_new[i] - out_ref[i])); big = fmax(big, fabs(out_ref[i])); }
    const float ms200 = run_probe<200>(d_new, L, Q, prop.multiProcessorCount, 5, 10);
    const float ms168 = run_probe<168>(d_new, L, Q, prop.multiProcessorCount, 6, 10);
    const float ms248 = run_probe<248>(d_new, L, Q, prop.multiProcessorCount, 4, 10);
    printf("{\"scenario\": \"%s\", \"ops\": %d, \"library_ms\": %.4f, \"fused_200regs_5ctas_ms\": %.4f, \"fused_168regs_6ctas_ms\": %.4f, "
           "\"fused_248regs_4ctas_ms\": %.4f, \"max_abs_diff\": %.3e, \"max_abs_value\": %.3e}\n",
           S.name, P.n_ops, ms_ref, ms200, ms168, ms248, worst, big);
  }
  return 0;
}
