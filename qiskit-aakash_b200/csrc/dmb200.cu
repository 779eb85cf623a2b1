// dmb200.cu -- libdmb200.so: CUDA kernels (sm_100a) + the C ABI declared in include/dmb200.h.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC
//             -Iinclude qiskit-aakash_b200/csrc/dmb200.cu -o qiskit-aakash_b200/libdmb200.so
//
// Hot kernel: k_tile_pass -- one HBM round trip of the Pauli-coefficient vector per launch,
// with up to DMB_MAX_OPS fused two-digit ops (pre-multiplied single-qubit maps + CNOT / TSP
// CNOT / diagonal mask / digit swap) applied to each 4^K-coefficient tile while it sits in
// shared memory.  HBM-bound by design: 16 algorithmic bytes per coefficient per launch.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include <mutex>
#include <atomic>
#include <string>
#include <vector>

#include "dm_device.h"
#include "dm_schedule.h"

// ---------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(const char* what, const char* detail = nullptr) {
  g_err = what;
  if (detail) { g_err += ": "; g_err += detail; }
  return 1;
}

#define CU_TRY(expr)                                                        \
  do {                                                                      \
    cudaError_t e_ = (expr);                                                \
    if (e_ != cudaSuccess) return fail(#expr, cudaGetErrorString(e_));      \
  } while (0)

// Every entry point runs on the context's device and puts the caller's current device back afterwards: the
// host (PyTorch) keeps its own notion of the current device, and a backend built on device 3 must not
// redirect the caller's later torch.*(device="cuda") work there.
struct dmb_device_guard {
  int prev = -1, dev;
  cudaError_t err = cudaSuccess;
  explicit dmb_device_guard(int d) : dev(d) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~dmb_device_guard() {
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};
#define DMB_ON_DEVICE(ctx)                                                            \
  dmb_device_guard dev_guard_((ctx)->device);                                         \
  if (dev_guard_.err != cudaSuccess) return fail("cudaSetDevice", cudaGetErrorString(dev_guard_.err))

struct dmb_ctx {
  int device;
  cudaStream_t stream;
  dmb_stats stats;
  int tile_variant;
  int sm_count;
  double* d_scratch;        // small device scratch (gathers, reduction partials)
  double* h_scratch;        // pinned host mirror
  uint64_t* d_idx;
  size_t scratch_elems;
  // small-state path (whole pass list in one launch): device copy of the pre-converted passes + pinned staging
  unsigned char* d_plan;
  unsigned char* h_plan;
  cudaEvent_t plan_copied;
};

static const size_t kScratchElems = 1 << 16;

// ---------------------------------------------------------------------------------------
// tile pass, generic K (2..6): register-staged 128-bit loads/stores, one tile per CTA
// iteration.  Used for states smaller than 4^6 coefficients (K < 6) and kept selectable
// for K = 6 (dmb_set_tile_variant(ctx, 1)) as the A/B baseline of the lean kernel below.
// ---------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(DMB_TILE_THREADS, (K == 6 ? 3 : 1))
k_tile_pass(double* __restrict__ state, const __grid_constant__ dmb_pass P, uint64_t n_tiles,
            const __grid_constant__ dmb_remote_src S, const __grid_constant__ dmb_remote_src D) {
  extern __shared__ __align__(16) double smem[];
  constexpr int MAXPAIRS = ((1 << (2 * K - 1)) + DMB_TILE_THREADS - 1) / DMB_TILE_THREADS;
  const int t = threadIdx.x;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t tbase = dmb_tile_base(tile, P.tile_digit, K);
    dmb_tile_load_thread<MAXPAIRS>(t, state, tbase, smem, P.tile_digit, K, S);
    __syncthreads();
    for (int i = 0; i < P.n_ops; ++i) {
      dmb_tile_op_thread(t, P.ops[i], smem, K);
      __syncthreads();
    }
    dmb_tile_store_thread<MAXPAIRS>(t, state, tbase, smem, P.tile_digit, K, D);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// tile pass, K = 6: the hot kernel.  128 threads per 4096-coefficient tile, one 32 KiB stage per CTA,
// 5 CTAs per SM (persistent: CTA c handles tiles c, c + grid, ...).  The 16-byte pairs of a tile stream in
// with cp.async.cg (LDGSTS, no register staging, L1 bypass), the fused ops run in place in shared memory
// (one barrier after each), and the tile is written back with LDS.128 + STG.128 -- the other four CTAs of
// the SM cover the HBM latency of this CTA's load, so no prefetch ring is needed (measured: 223 ms for
// BASELINE config 3 against 231 ms with a 2-stage ring x 3 CTAs and 252 ms for round 1's 256-thread kernel,
// profiles/r02_tile_variants.md).  The control flow is dmb_half_kernel_body in dm_device.h, the function
// the CPU tests run with real host threads; this is its CUDA execution context.
// ---------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared-window accessor: 32-bit addresses, one LDS/STS per access
struct dmb_smem_mem {
  uint32_t base;
  __device__ __forceinline__ double ld64(uint32_t off) const {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + off) : "memory");
    return v;
  }
  __device__ __forceinline__ dmb_d2 ld128(uint32_t off) const {
    dmb_d2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + off) : "memory");
    return v;
  }
  __device__ __forceinline__ void st64(uint32_t off, double v) const {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + off), "d"(v) : "memory");
  }
  __device__ __forceinline__ void st128(uint32_t off, dmb_d2 v) const {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(base + off), "d"(v.x), "d"(v.y) : "memory");
  }
};

struct dmb_cuda_cta {
  uint32_t smem0;
  __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
  __device__ __forceinline__ uint64_t block() const { return blockIdx.x; }
  __device__ __forceinline__ uint64_t grid() const { return gridDim.x; }
  __device__ __forceinline__ void copy16(uint32_t off, const double* src) const {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem0 + off), "l"(src) : "memory");
  }
  __device__ __forceinline__ void commit() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
  template <int N>
  __device__ __forceinline__ void wait() const { cp_async_wait<N>(); }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ dmb_smem_mem mem(uint32_t off) const {
    dmb_smem_mem m;
    m.base = smem0 + off;
    return m;
  }
};

// REMOTE: 0 in place, 1 pull (remote loads), 2 push (remote stores); STMODE: DMB_ST_PLAIN, or a relabelling
// store that realises the pass's trailing digit swaps
// CHAINS: the instantiation that also carries the chained op bodies (passes with L.n_chained > 0, dmb_chain_ops); the
// common one is kept free of them
template <int CTAS, int REMOTE, int STMODE, bool CHAINS = false>
__global__ void __launch_bounds__(DMB_HALF_THREADS, CTAS)
k_tile_pass6(double* __restrict__ state, const __grid_constant__ dmb_lean_pass L,
             const __grid_constant__ dmb_remote_src S) {
  extern __shared__ __align__(128) unsigned char lean_smem[];
  dmb_cuda_cta cx;
  cx.smem0 = (uint32_t)__cvta_generic_to_shared(lean_smem);
  dmb_half_kernel_body<STMODE, true, 1, REMOTE, CHAINS>(cx, state, L, S);
}

// ---------------------------------------------------------------------------------------
// Small states (at most 4^8 coefficients = 512 KiB, i.e. 1, 4 or 16 tiles): the WHOLE pass list of a flush in ONE
// launch.  A circuit on <= 8 qubits is launch-bound when every pass is a kernel (QFT-8: five passes on a state
// that lives in L2); here the tiles' CTAs form one thread-block cluster (16 CTAs: non-portable size, opted in),
// each pass is "load my tile, run the ops, write it back" exactly as in k_tile_pass6, and between passes the
// CTAs meet at the hardware cluster barrier (release / acquire: the write-backs of pass p are visible to the loads
// of pass p + 1, which fetch through L2).  The pre-converted passes (dmb_lean_pass) are read from global memory.
// States below one 4^6 tile (n_digits = K < 6): a single CTA keeps the one tile in shared memory across all passes.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(DMB_HALF_THREADS)
k_small_passes6(double* __restrict__ state, const dmb_lean_pass* __restrict__ plan, int n_passes) {
  extern __shared__ __align__(128) unsigned char lean_smem[];
  dmb_cuda_cta cx;
  cx.smem0 = (uint32_t)__cvta_generic_to_shared(lean_smem);
  const auto mem = cx.mem(0);
  const int t = (int)threadIdx.x;
  dmb_remote_src none;
  none.enabled = 0;
  for (int p = 0; p < n_passes; ++p) {
    const dmb_lean_pass& L = plan[p];
    dmb_lean_thread S0, S1, P0;
    dmb_lean_thread_init(t, L, S0);
    dmb_lean_thread_init(t + DMB_HALF_THREADS, L, S1);
    dmb_lean_thread_init(2 * t, L, P0);
    const uint64_t tb = dmb_tile_base(blockIdx.x, L.td, DMB_LEAN_K);
#pragma unroll
    for (int i = 0; i < DMB_LEAN_PAIRS; ++i) {
      cx.copy16(S0.soff ^ L.pair_soff[i], state + tb + (S0.goff | L.pair_goff[i]));
      cx.copy16(S1.soff ^ L.pair_soff[i], state + tb + (S1.goff | L.pair_goff[i]));
    }
    cx.commit();
    cx.wait<0>();
    __syncthreads();
    for (int i = 0; i < L.n_ops;) {
      i += dmb_lean_ops_step<true>(P0, S0, &L.ops[i], mem);
      __syncthreads();
    }
    if (L.st_mode == DMB_ST_PERM128) {
      dmb_lean_store_thread<false, DMB_ST_PERM128>(S0, L, state, tb, none, mem);
      dmb_lean_store_thread<false, DMB_ST_PERM128>(S1, L, state, tb, none, mem);
    } else if (L.st_mode == DMB_ST_SPLIT64) {
      dmb_lean_store_thread<false, DMB_ST_SPLIT64>(S0, L, state, tb, none, mem);
      dmb_lean_store_thread<false, DMB_ST_SPLIT64>(S1, L, state, tb, none, mem);
    } else {
      dmb_lean_store_thread<false, DMB_ST_PLAIN>(S0, L, state, tb, none, mem);
      dmb_lean_store_thread<false, DMB_ST_PLAIN>(S1, L, state, tb, none, mem);
    }
    cluster_barrier();       // also a CTA barrier: the stage is free for the next pass's copies
  }
}

template <int K>
__global__ void __launch_bounds__(DMB_TILE_THREADS)
k_small_passes(double* __restrict__ state, const dmb_pass* __restrict__ plan, int n_passes) {
  extern __shared__ __align__(16) double smem[];
  constexpr int MAXPAIRS = ((1 << (2 * K - 1)) + DMB_TILE_THREADS - 1) / DMB_TILE_THREADS;
  const int t = threadIdx.x;
  dmb_remote_src none;
  none.enabled = 0;
  // n_digits == K: every pass's tile is the whole state in ascending digit order -- stage it once
  dmb_tile_load_thread<MAXPAIRS>(t, state, 0, smem, plan[0].tile_digit, K, none);
  __syncthreads();
  for (int p = 0; p < n_passes; ++p)
    for (int i = 0; i < plan[p].n_ops; ++i) {
      dmb_tile_op_thread(t, plan[p].ops[i], smem, K);
      __syncthreads();
    }
  dmb_tile_store_thread<MAXPAIRS>(t, state, 0, smem, plan[0].tile_digit, K, none);
}

// ---------------------------------------------------------------------------------------
// element-wise kernels
// ---------------------------------------------------------------------------------------
__global__ void k_init_product(double* __restrict__ state, uint64_t count, const __grid_constant__ dmb_init_params p) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    state[i] = dmb_init_value(i, p);
}

__global__ void k_marginal(const double* __restrict__ state, double* __restrict__ out, uint64_t count,
                           const __grid_constant__ dmb_marginal_params p) {
  for (uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; c < count; c += (uint64_t)gridDim.x * blockDim.x)
    out[c] = dmb_marginal_value(c, state, p);
}

__global__ void k_fwht_stage(double* vec, uint64_t pairs, int stage) {
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < pairs; t += (uint64_t)gridDim.x * blockDim.x)
    dmb_fwht_pair(t, stage, vec);
}

// all stages of a <= 2^11-point transform (or the low 11 stages of a larger one) in shared memory
__global__ void __launch_bounds__(1024) k_fwht_smem(double* vec, int n_stages) {
  __shared__ double s[2048];
  const uint64_t base = (uint64_t)blockIdx.x << n_stages;
  const uint32_t len = 1u << n_stages;
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) s[i] = vec[base + i];
  __syncthreads();
  for (int st = 0; st < n_stages; ++st) {
    for (uint32_t t = threadIdx.x; t < len / 2; t += blockDim.x) dmb_fwht_pair(t, st, s);
    __syncthreads();
  }
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) vec[base + i] = s[i];
}

__global__ void k_contract_digit(const double* __restrict__ in, double* __restrict__ out, uint64_t count,
                                 uint64_t L, double n0, double n1, double n2) {
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < count; t += (uint64_t)gridDim.x * blockDim.x)
    dmb_contract_elem(t, in, out, L, n0, n1, n2);
}

__global__ void k_gather(const double* __restrict__ state, const uint64_t* __restrict__ idx, size_t k,
                         double* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < k; i += (size_t)gridDim.x * blockDim.x)
    out[i] = state[idx[i]];
}

__global__ void k_tomatrix_digit(const double* __restrict__ state, dmb_d2* work, uint64_t count, int pos, int first) {
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < count; t += (uint64_t)gridDim.x * blockDim.x)
    dmb_tomatrix_digit(t, pos, first, state, work);
}

__global__ void k_tomatrix_scatter(const dmb_d2* __restrict__ work, dmb_d2* __restrict__ out, uint64_t count, int n) {
  for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < count; t += (uint64_t)gridDim.x * blockDim.x)
    dmb_tomatrix_scatter(t, n, work, out);
}

__global__ void k_chop(double* state, uint64_t count, double thr) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
    const double x = state[i];
    if (fabs(x) < thr) state[i] = 0.0;
  }
}

// deterministic two-stage dot product: fixed grid, fixed per-thread order, tree in the block
__global__ void __launch_bounds__(256) k_dot_partial(const double* __restrict__ a, const double* __restrict__ b,
                                                     uint64_t count, double* __restrict__ partial) {
  __shared__ double red[256];
  double acc = 0.0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x)
    acc += a[i] * b[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
static inline unsigned grid_for(uint64_t count, int block, int sm_count) {
  uint64_t g = (count + block - 1) / block;
  const uint64_t cap = (uint64_t)sm_count * 32;
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (unsigned)g;
}

static dmb_remote_src g_no_remote;      // zero-initialised: enabled == 0

template <int K>
static int launch_tile_pass(dmb_ctx* ctx, double* state, int n_bits, const dmb_pass& P,
                            const dmb_remote_src& S = g_no_remote, const dmb_remote_src& D = g_no_remote) {
  const uint64_t n_tiles = 1ull << (n_bits - 2 * K);
  const size_t smem = sizeof(double) << (2 * K);
  static std::atomic<uint64_t> attr_done[2];          // one bit per device: the attribute is per device
  const int dev = ctx->device & 127;
  if (!((attr_done[dev >> 6].load() >> (dev & 63)) & 1ull)) {
    CU_TRY(cudaFuncSetAttribute(k_tile_pass<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev >> 6].fetch_or(1ull << (dev & 63));
  }
  const uint64_t grid = n_tiles < 0x7fffffffull ? n_tiles : 0x7fffffffull;
  k_tile_pass<K><<<(unsigned)grid, DMB_TILE_THREADS, smem, ctx->stream>>>(state, P, n_tiles, S, D);
  CU_TRY(cudaGetLastError());
  return 0;
}

#define DMB_TILE_CTAS 5      // CTAs per SM of k_tile_pass6 (measured on config 3: 4 -> 232.0 ms, 5 -> 223.0 ms, 6 -> 226.7 ms)
template <int CTAS, int REMOTE, int STMODE, bool CHAINS>
static int launch_tile6_inst(dmb_ctx* ctx, double* state, const dmb_lean_pass& L, const dmb_remote_src& S) {
  const size_t smem = DMB_LEAN_TILE_BYTES;
  static std::atomic<uint64_t> attr_done[2];          // one bit per device: the attribute is per device
  const int dev = ctx->device & 127;
  if (!((attr_done[dev >> 6].load() >> (dev & 63)) & 1ull)) {
    CU_TRY(cudaFuncSetAttribute(k_tile_pass6<CTAS, REMOTE, STMODE, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev >> 6].fetch_or(1ull << (dev & 63));
  }
  uint64_t grid = (uint64_t)ctx->sm_count * CTAS;
  if (grid > L.n_tiles) grid = L.n_tiles;
  k_tile_pass6<CTAS, REMOTE, STMODE, CHAINS><<<(unsigned)grid, DMB_HALF_THREADS, smem, ctx->stream>>>(state, L, S);
  CU_TRY(cudaGetLastError());
  return 0;
}

template <int CTAS, int REMOTE, int STMODE>
static int launch_tile6(dmb_ctx* ctx, double* state, const dmb_lean_pass& L, const dmb_remote_src& S = g_no_remote) {
  return L.n_chained > 0 ? launch_tile6_inst<CTAS, REMOTE, STMODE, true>(ctx, state, L, S)
                         : launch_tile6_inst<CTAS, REMOTE, STMODE, false>(ctx, state, L, S);
}

template <int CTAS>
static int launch_tile6_any(dmb_ctx* ctx, double* state, const dmb_lean_pass& L) {
  if (L.st_mode == DMB_ST_PERM128) return launch_tile6<CTAS, 0, DMB_ST_PERM128>(ctx, state, L);
  if (L.st_mode == DMB_ST_SPLIT64) return launch_tile6<CTAS, 0, DMB_ST_SPLIT64>(ctx, state, L);
  return launch_tile6<CTAS, 0, DMB_ST_PLAIN>(ctx, state, L);
}

static int launch_tile_pass6(dmb_ctx* ctx, double* state, int n_bits, const dmb_pass& P) {
  static thread_local dmb_lean_pass L;    // 6.5 KB: keep it off the stack; one host thread drives a ctx
  dmb_make_lean_pass(P, n_bits, L, dmb_fold_swaps_enabled());
  ctx->stats.folded_swaps += (uint64_t)(P.n_ops - L.n_ops);
  ctx->stats.chained_ops += (uint64_t)L.n_chained;
  return launch_tile6_any<DMB_TILE_CTAS>(ctx, state, L);
}

static int validate_pass(const dmb_pass& P, int n_bits) {
  const int K = P.n_tile_digits;
  if (K < 2 || K > DMB_MAX_TILE_DIGITS) return fail("dmb_apply_passes", "n_tile_digits out of range");
  if (2 * K > n_bits) return fail("dmb_apply_passes", "tile larger than the state");
  if (P.n_ops < 0 || P.n_ops > DMB_MAX_OPS) return fail("dmb_apply_passes", "n_ops out of range");
  if (P.tile_digit[0] != 0) return fail("dmb_apply_passes", "tile_digit[0] must be 0");
  for (int j = 0; j < K; ++j) {
    if (j && P.tile_digit[j] <= P.tile_digit[j - 1]) return fail("dmb_apply_passes", "tile digits not ascending");
    if (2 * P.tile_digit[j] + 1 >= n_bits) return fail("dmb_apply_passes", "tile digit outside the state");
  }
  for (int i = 0; i < P.n_ops; ++i) {
    const dmb_op& op = P.ops[i];
    if (op.kind < DMB_OP_MATS || op.kind > DMB_OP_SWAP) return fail("dmb_apply_passes", "unknown op kind");
    if (op.a < 0 || op.a >= K || op.b < 0 || op.b >= K || op.a == op.b)
      return fail("dmb_apply_passes", "op digits invalid");
    unsigned seen = (1u << op.a) | (1u << op.b);
    for (int m = 0; m < K - 2; ++m) {
      if (op.fd[m] < 0 || op.fd[m] >= K || (seen >> op.fd[m]) & 1u)
        return fail("dmb_apply_passes", "op free-digit list is not a permutation");
      seen |= 1u << op.fd[m];
    }
    if (op.reserved_[0] || op.reserved_[1]) return fail("dmb_apply_passes", "dmb_op.reserved_ must be 0");
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int dmb_abi_version(void) { return DMB_ABI_VERSION; }
size_t dmb_sizeof_op(void) { return sizeof(dmb_op); }
size_t dmb_sizeof_pass(void) { return sizeof(dmb_pass); }
const char* dmb_last_error(void) { return g_err.c_str(); }

size_t dmb_sizeof_qop(void) { return sizeof(dmb_qop); }
int dmb_schedule(const dmb_qop* ops, size_t n_ops, int32_t* pos, int n_qubits, int n_digits, int max_tile,
                 int max_ops, int window, int strategy, size_t min_tail, int32_t* moves, int32_t* n_moves,
                 dmb_pass* out, size_t out_cap, size_t* n_out, int32_t* left, size_t* n_left) {
  if ((!ops && n_ops) || !pos || !out || !n_out) return fail("dmb_schedule", "null argument");
  std::string err;
  if (dmb_sched::schedule_impl(ops, n_ops, pos, n_qubits, n_digits, max_tile, max_ops, window, strategy, min_tail, moves,
                               n_moves, out, out_cap, n_out, left, n_left, err))
    return fail("dmb_schedule", err.c_str());
  return 0;
}

int dmb_create(int device, dmb_ctx** out) {
  if (!out) return fail("dmb_create", "null out pointer");
  int count = 0;
  CU_TRY(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail("dmb_create", "no such CUDA device");
  dmb_device_guard dev_guard_(device);
  if (dev_guard_.err != cudaSuccess) return fail("cudaSetDevice", cudaGetErrorString(dev_guard_.err));
  cudaDeviceProp prop;
  CU_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("dmb_create", "libdmb200 is built for sm_100a (B200) only");
  dmb_ctx* ctx = new dmb_ctx();
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  ctx->stream = 0;
  ctx->tile_variant = 0;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->scratch_elems = kScratchElems;
  cudaError_t e = cudaMalloc(&ctx->d_scratch, kScratchElems * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_idx, kScratchElems * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_scratch, kScratchElems * sizeof(double));
  if (e != cudaSuccess) {                  // do not leak a half-built context
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->d_idx);
    delete ctx;
    return fail("dmb_create: scratch allocation", cudaGetErrorString(e));
  }
  *out = ctx;
  return 0;
}

int dmb_destroy(dmb_ctx* ctx) {
  if (!ctx) return 0;
  dmb_device_guard dev_guard_(ctx->device);
  cudaFree(ctx->d_scratch);
  cudaFree(ctx->d_idx);
  cudaFreeHost(ctx->h_scratch);
  if (ctx->d_plan) {
    cudaFree(ctx->d_plan);
    cudaFreeHost(ctx->h_plan);
    cudaEventDestroy(ctx->plan_copied);
  }
  delete ctx;
  return 0;
}

int dmb_set_stream(dmb_ctx* ctx, void* cuda_stream) {
  if (!ctx) return fail("dmb_set_stream", "null context");
  ctx->stream = (cudaStream_t)cuda_stream;
  return 0;
}

int dmb_sync(dmb_ctx* ctx) {
  if (!ctx) return fail("dmb_sync", "null context");
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dmb_get_stats(dmb_ctx* ctx, dmb_stats* out) {
  if (!ctx || !out) return fail("dmb_get_stats", "null argument");
  *out = ctx->stats;
  return 0;
}

int dmb_reset_stats(dmb_ctx* ctx) {
  if (!ctx) return fail("dmb_reset_stats", "null context");
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  return 0;
}

int dmb_set_tile_variant(dmb_ctx* ctx, int variant) {
  if (!ctx) return fail("dmb_set_tile_variant", "null context");
  if (variant < 0 || variant > 1) return fail("dmb_set_tile_variant", "variant must be 0 or 1");
  ctx->tile_variant = variant;
  return 0;
}

int dmb_init_product(dmb_ctx* ctx, double* state, int n_bits, uint64_t rank_bits, int n_qubits,
                     const int32_t* hi, const int32_t* lo, const double* v, double scale) {
  if (!ctx || !state || !hi || !lo || !v) return fail("dmb_init_product", "null argument");
  if (n_qubits < 1 || n_qubits > DMB_MAX_QUBITS) return fail("dmb_init_product", "n_qubits out of range");
  if (n_bits < 0 || n_bits > 62) return fail("dmb_init_product", "n_bits out of range");
  DMB_ON_DEVICE(ctx);
  dmb_init_params p;
  memset(&p, 0, sizeof(p));
  p.map.n_qubits = n_qubits;
  for (int q = 0; q < n_qubits; ++q) {
    p.map.hi[q] = hi[q];
    p.map.lo[q] = lo[q];
    for (int d = 0; d < 4; ++d) p.v[q][d] = v[4 * q + d];
  }
  p.scale = scale;
  p.rank_bits = rank_bits;
  p.n_bits = n_bits;
  const uint64_t count = 1ull << n_bits;
  k_init_product<<<grid_for(count, 256, ctx->sm_count), 256, 0, ctx->stream>>>(state, count, p);
  CU_TRY(cudaGetLastError());
  ctx->stats.other_launches++;
  return 0;
}

static const size_t kPlanBytes = 1 << 20;          // 1 MiB: 157 pre-converted K = 6 passes per launch

static bool small_path_enabled() {                 // DMB_SMALL_PATH=0: one launch per pass also for small states (A/B)
  static const bool on = [] { const char* e = getenv("DMB_SMALL_PATH"); return !(e && e[0] == '0'); }();
  return on;
}

// n_passes >= 2 passes on a state of at most 16 tiles, all validated: one launch for all of them
static int apply_passes_small(dmb_ctx* ctx, double* state, int n_bits, const dmb_pass* passes, size_t n_passes) {
  const int K = passes[0].n_tile_digits;
  if (!ctx->d_plan) {
    CU_TRY(cudaMalloc(&ctx->d_plan, kPlanBytes));
    CU_TRY(cudaMallocHost(&ctx->h_plan, kPlanBytes));
    CU_TRY(cudaEventCreateWithFlags(&ctx->plan_copied, cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(ctx->plan_copied, ctx->stream));
  }
  const size_t item = K == DMB_LEAN_K ? sizeof(dmb_lean_pass) : sizeof(dmb_pass);
  const size_t per_launch = kPlanBytes / item;
  for (size_t done = 0; done < n_passes; done += per_launch) {
    const size_t m = (n_passes - done) < per_launch ? (n_passes - done) : per_launch;
    CU_TRY(cudaEventSynchronize(ctx->plan_copied));        // the staging buffer is free again
    if (K == DMB_LEAN_K) {
      dmb_lean_pass* h = reinterpret_cast<dmb_lean_pass*>(ctx->h_plan);
      for (size_t i = 0; i < m; ++i) {
        dmb_make_lean_pass(passes[done + i], n_bits, h[i], dmb_fold_swaps_enabled());
        ctx->stats.folded_swaps += (uint64_t)(passes[done + i].n_ops - h[i].n_ops);
        ctx->stats.chained_ops += (uint64_t)h[i].n_chained;
      }
    } else {
      memcpy(ctx->h_plan, passes + done, m * item);
    }
    CU_TRY(cudaMemcpyAsync(ctx->d_plan, ctx->h_plan, m * item, cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(cudaEventRecord(ctx->plan_copied, ctx->stream));
    if (K == DMB_LEAN_K) {
      const unsigned n_tiles = 1u << (n_bits - 2 * DMB_LEAN_K);
      static std::atomic<uint64_t> attr_done[2];
      const int dev = ctx->device & 127;
      if (!((attr_done[dev >> 6].load() >> (dev & 63)) & 1ull)) {
        CU_TRY(cudaFuncSetAttribute(k_small_passes6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DMB_LEAN_TILE_BYTES));
        CU_TRY(cudaFuncSetAttribute(k_small_passes6, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr_done[dev >> 6].fetch_or(1ull << (dev & 63));
      }
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(n_tiles, 1, 1);
      cfg.blockDim = dim3(DMB_HALF_THREADS, 1, 1);
      cfg.dynamicSmemBytes = DMB_LEAN_TILE_BYTES;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = n_tiles;
      attr.val.clusterDim.y = 1;
      attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      CU_TRY(cudaLaunchKernelEx(&cfg, k_small_passes6, state, reinterpret_cast<const dmb_lean_pass*>(ctx->d_plan), (int)m));
    } else {
      const dmb_pass* d = reinterpret_cast<const dmb_pass*>(ctx->d_plan);
      const size_t smem = sizeof(double) << (2 * K);
      switch (K) {
        case 2: k_small_passes<2><<<1, DMB_TILE_THREADS, smem, ctx->stream>>>(state, d, (int)m); break;
        case 3: k_small_passes<3><<<1, DMB_TILE_THREADS, smem, ctx->stream>>>(state, d, (int)m); break;
        case 4: k_small_passes<4><<<1, DMB_TILE_THREADS, smem, ctx->stream>>>(state, d, (int)m); break;
        default: k_small_passes<5><<<1, DMB_TILE_THREADS, smem, ctx->stream>>>(state, d, (int)m); break;
      }
      CU_TRY(cudaGetLastError());
    }
    ctx->stats.tile_pass_launches++;
    ctx->stats.small_plan_launches++;
  }
  for (size_t i = 0; i < n_passes; ++i) {
    ctx->stats.fused_ops += (uint64_t)passes[i].n_ops;
    ctx->stats.state_bytes_moved += 16ull << n_bits;
  }
  return 0;
}

int dmb_apply_passes(dmb_ctx* ctx, double* state, int n_bits, const dmb_pass* passes, size_t n_passes) {
  if (!ctx || !state || (!passes && n_passes)) return fail("dmb_apply_passes", "null argument");
  DMB_ON_DEVICE(ctx);
  if (n_passes >= 2 && n_bits <= 16 && ctx->tile_variant == 0 && small_path_enabled()) {
    bool same = true;
    for (size_t i = 0; i < n_passes; ++i) {
      if (validate_pass(passes[i], n_bits)) return 1;
      if (passes[i].n_tile_digits != passes[0].n_tile_digits) same = false;
    }
    const int K = passes[0].n_tile_digits;
    if (same && (K == DMB_LEAN_K || 2 * K == n_bits)) return apply_passes_small(ctx, state, n_bits, passes, n_passes);
  }
  for (size_t i = 0; i < n_passes; ++i) {
    if (validate_pass(passes[i], n_bits)) return 1;
    const dmb_pass& P = passes[i];
    int rc = 0;
    switch (P.n_tile_digits) {
      case 2: rc = launch_tile_pass<2>(ctx, state, n_bits, P); break;
      case 3: rc = launch_tile_pass<3>(ctx, state, n_bits, P); break;
      case 4: rc = launch_tile_pass<4>(ctx, state, n_bits, P); break;
      case 5: rc = launch_tile_pass<5>(ctx, state, n_bits, P); break;
      case 6:
        rc = ctx->tile_variant == 1 ? launch_tile_pass<6>(ctx, state, n_bits, P)
                                    : launch_tile_pass6(ctx, state, n_bits, P);
        break;
      default: return fail("dmb_apply_passes", "unsupported tile size");
    }
    if (rc) return rc;
    ctx->stats.tile_pass_launches++;
    ctx->stats.fused_ops += (uint64_t)P.n_ops;
    ctx->stats.state_bytes_moved += 16ull << n_bits;
  }
  return 0;
}

int dmb_apply_pass_remote(dmb_ctx* ctx, double* dst_state, int n_bits, const dmb_pass* pass,
                          const uint64_t* src_tab, int tab_bits, int block_shift, int push) {
  if (tab_bits < 0 || tab_bits > DMB_REMOTE_BITS) return fail("dmb_apply_pass_remote", "table too large");
  if (block_shift + tab_bits != n_bits) return fail("dmb_apply_pass_remote", "block_shift + tab_bits != n_bits");
  int32_t sel[DMB_REMOTE_BITS];
  for (int j = 0; j < tab_bits; ++j) sel[j] = block_shift + j;       // the table index = the high bits of the element index
  return dmb_apply_pass_remote_sel(ctx, dst_state, n_bits, pass, src_tab, tab_bits, sel, push);
}

int dmb_apply_pass_remote_sel(dmb_ctx* ctx, double* dst_state, int n_bits, const dmb_pass* pass,
                              const uint64_t* src_tab, int n_sel, const int32_t* sel_bits, int push) {
  if (!ctx || !dst_state || !pass || !src_tab || (!sel_bits && n_sel)) return fail("dmb_apply_pass_remote", "null argument");
  if (n_sel < 0 || n_sel > DMB_REMOTE_BITS) return fail("dmb_apply_pass_remote", "table too large");
  for (int j = 0; j < n_sel; ++j)
    if (sel_bits[j] < 0 || sel_bits[j] >= n_bits) return fail("dmb_apply_pass_remote", "selected bit outside the local index");
  DMB_ON_DEVICE(ctx);
  if (validate_pass(*pass, n_bits)) return 1;
  const dmb_pass& P = *pass;
  dmb_remote_src S;
  memset(&S, 0, sizeof(S));
  for (int i = 0; i < (1 << n_sel); ++i) S.tab[i] = src_tab[i];
  S.n_sel = n_sel;
  for (int j = 0; j < n_sel; ++j) S.sel[j] = sel_bits[j];
  S.enabled = 1;
  int rc = 0;
  const dmb_remote_src& ld = push ? g_no_remote : S;     // pull: table drives the loads
  const dmb_remote_src& st = push ? S : g_no_remote;     // push: table drives the stores
  switch (P.n_tile_digits) {
    case 2: rc = launch_tile_pass<2>(ctx, dst_state, n_bits, P, ld, st); break;
    case 3: rc = launch_tile_pass<3>(ctx, dst_state, n_bits, P, ld, st); break;
    case 4: rc = launch_tile_pass<4>(ctx, dst_state, n_bits, P, ld, st); break;
    case 5: rc = launch_tile_pass<5>(ctx, dst_state, n_bits, P, ld, st); break;
    case 6: {
      static thread_local dmb_lean_pass L;
      dmb_make_lean_pass(P, n_bits, L);
      ctx->stats.chained_ops += (uint64_t)L.n_chained;
      rc = push ? launch_tile6<DMB_TILE_CTAS, 2, DMB_ST_PLAIN>(ctx, dst_state, L, S)
                : launch_tile6<DMB_TILE_CTAS, 1, DMB_ST_PLAIN>(ctx, dst_state, L, S);
      break;
    }
    default: return fail("dmb_apply_pass_remote", "unsupported tile size");
  }
  if (rc) return rc;
  ctx->stats.tile_pass_launches++;
  ctx->stats.fused_ops += (uint64_t)P.n_ops;
  ctx->stats.state_bytes_moved += 16ull << n_bits;
  return 0;
}

int dmb_ipc_export(dmb_ctx* ctx, const void* dev_ptr, unsigned char* handle64, uint64_t* offset) {
  if (!ctx || !dev_ptr || !handle64 || !offset) return fail("dmb_ipc_export", "null argument");
  DMB_ON_DEVICE(ctx);
  // the handle names the whole allocation: find its base through the driver API, resolved at
  // run time so that the library still loads on machines without libcuda (build/CI boxes)
  typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
  static range_fn get_range = nullptr;
  if (!get_range) {
    void* drv = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (drv) get_range = (range_fn)dlsym(drv, "cuMemGetAddressRange_v2");
    if (!get_range) return fail("dmb_ipc_export", "cuMemGetAddressRange_v2 not available");
  }
  unsigned long long base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0)
    return fail("dmb_ipc_export", "cuMemGetAddressRange failed");
  cudaIpcMemHandle_t h;
  CU_TRY(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *offset = (uint64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
  return 0;
}

// An allocation can be mapped only once per process, and PyTorch's caching allocator hands the same allocation to
// successive engines: mappings are reference counted per handle (dmb_ipc_open / dmb_ipc_close) under a mutex.
struct dmb_ipc_mapping { void* base; int refs; int device; };
static std::mutex g_ipc_mutex;
static std::map<std::string, dmb_ipc_mapping> g_ipc_open;

int dmb_ipc_open(dmb_ctx* ctx, const unsigned char* handle64, uint64_t offset, void** out_ptr) {
  if (!ctx || !handle64 || !out_ptr) return fail("dmb_ipc_open", "null argument");
  DMB_ON_DEVICE(ctx);
  std::lock_guard<std::mutex> lock(g_ipc_mutex);
  const std::string key((const char*)handle64, 64);
  auto it = g_ipc_open.find(key);
  if (it == g_ipc_open.end()) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* base = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    it = g_ipc_open.emplace(key, dmb_ipc_mapping{base, 0, ctx->device}).first;
  }
  it->second.refs++;
  *out_ptr = (void*)((char*)it->second.base + offset);
  return 0;
}

int dmb_ipc_close(dmb_ctx* ctx, const unsigned char* handle64) {
  if (!ctx || !handle64) return fail("dmb_ipc_close", "null argument");
  std::lock_guard<std::mutex> lock(g_ipc_mutex);
  const std::string key((const char*)handle64, 64);
  auto it = g_ipc_open.find(key);
  if (it == g_ipc_open.end()) return fail("dmb_ipc_close", "handle is not open");
  if (--it->second.refs > 0) return 0;
  dmb_device_guard guard(it->second.device);
  const cudaError_t e = cudaIpcCloseMemHandle(it->second.base);
  g_ipc_open.erase(it);
  if (e != cudaSuccess) return fail("cudaIpcCloseMemHandle", cudaGetErrorString(e));
  return 0;
}

int dmb_marginal(dmb_ctx* ctx, const double* state, int n_bits, uint64_t rank_bits, int n_qubits,
                 const int32_t* hi, const int32_t* lo, const double* wt, double* out) {
  if (!ctx || !state || !hi || !lo || !wt || !out) return fail("dmb_marginal", "null argument");
  if (n_qubits < 1 || n_qubits > DMB_MAX_QUBITS) return fail("dmb_marginal", "n_qubits out of range");
  DMB_ON_DEVICE(ctx);
  dmb_marginal_params p;
  memset(&p, 0, sizeof(p));
  p.map.n_qubits = n_qubits;
  p.rank_bits = rank_bits;
  p.n_bits = n_bits;
  for (int k = 0; k < n_qubits; ++k) {
    p.map.hi[k] = hi[k];
    p.map.lo[k] = lo[k];
    bool multi = false;
    for (int cb = 0; cb < 2; ++cb) {
      int nz = 0, last = 0;
      for (int d = 0; d < 4; ++d) {
        const double w = wt[(k * 2 + cb) * 4 + d];
        p.wt[k][cb][d] = w;
        if (w != 0.0) { ++nz; last = d; }
      }
      p.simple_digit[k][cb] = (int8_t)last;
      if (nz > 1) multi = true;
    }
    if (multi) {
      if (p.n_multi >= DMB_MAX_MULTI) return fail("dmb_marginal", "too many multi-weight qubits");
      p.simple_digit[k][0] = p.simple_digit[k][1] = -1;
      p.multi[p.n_multi++] = k;
    }
  }
  const uint64_t count = 1ull << n_qubits;
  k_marginal<<<grid_for(count, 256, ctx->sm_count), 256, 0, ctx->stream>>>(state, out, count, p);
  CU_TRY(cudaGetLastError());
  ctx->stats.other_launches++;
  return 0;
}

int dmb_fwht(dmb_ctx* ctx, double* vec, int n_qubits) {
  if (!ctx || !vec) return fail("dmb_fwht", "null argument");
  if (n_qubits < 0 || n_qubits > 40) return fail("dmb_fwht", "n_qubits out of range");
  DMB_ON_DEVICE(ctx);
  const int low = n_qubits < 11 ? n_qubits : 11;
  if (low > 0) {
    k_fwht_smem<<<(unsigned)(1ull << (n_qubits - low)), 1024, 0, ctx->stream>>>(vec, low);
    CU_TRY(cudaGetLastError());
    ctx->stats.other_launches++;
  }
  const uint64_t pairs = (1ull << n_qubits) >> 1;
  for (int st = low; st < n_qubits; ++st) {
    k_fwht_stage<<<grid_for(pairs, 256, ctx->sm_count), 256, 0, ctx->stream>>>(vec, pairs, st);
    CU_TRY(cudaGetLastError());
    ctx->stats.other_launches++;
  }
  return 0;
}

int dmb_contract_digit(dmb_ctx* ctx, const double* in, double* out, uint64_t H, uint64_t L, const double nv[3]) {
  if (!ctx || !in || !out || !nv) return fail("dmb_contract_digit", "null argument");
  DMB_ON_DEVICE(ctx);
  const uint64_t count = H * L;
  k_contract_digit<<<grid_for(count, 256, ctx->sm_count), 256, 0, ctx->stream>>>(in, out, count, L, nv[0], nv[1], nv[2]);
  CU_TRY(cudaGetLastError());
  ctx->stats.other_launches++;
  return 0;
}

int dmb_read_coeffs(dmb_ctx* ctx, const double* state, const uint64_t* idx, size_t k, double* out_host) {
  if (!ctx || !state || !idx || !out_host) return fail("dmb_read_coeffs", "null argument");
  if (k == 0) return 0;
  DMB_ON_DEVICE(ctx);
  for (size_t done = 0; done < k; done += ctx->scratch_elems) {       // the context's scratch bounds one round
    const size_t m = (k - done) < ctx->scratch_elems ? (k - done) : ctx->scratch_elems;
    CU_TRY(cudaMemcpyAsync(ctx->d_idx, idx + done, m * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    k_gather<<<grid_for(m, 256, ctx->sm_count), 256, 0, ctx->stream>>>(state, ctx->d_idx, m, ctx->d_scratch);
    CU_TRY(cudaGetLastError());
    ctx->stats.other_launches++;
    CU_TRY(cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, m * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    memcpy(out_host + done, ctx->h_scratch, m * sizeof(double));
  }
  return 0;
}

int dmb_to_matrix(dmb_ctx* ctx, const double* state, int n_qubits, double* work, double* out) {
  if (!ctx || !state || !work || !out) return fail("dmb_to_matrix", "null argument");
  if (n_qubits < 1 || n_qubits > 15) return fail("dmb_to_matrix", "n_qubits out of range");
  DMB_ON_DEVICE(ctx);
  const uint64_t vecs = 1ull << (2 * n_qubits - 2);
  for (int pos = 0; pos < n_qubits; ++pos) {
    k_tomatrix_digit<<<grid_for(vecs, 256, ctx->sm_count), 256, 0, ctx->stream>>>(state, (dmb_d2*)work, vecs, pos,
                                                                               pos == 0 ? 1 : 0);
    CU_TRY(cudaGetLastError());
    ctx->stats.other_launches++;
  }
  const uint64_t count = 1ull << (2 * n_qubits);
  k_tomatrix_scatter<<<grid_for(count, 256, ctx->sm_count), 256, 0, ctx->stream>>>((const dmb_d2*)work, (dmb_d2*)out,
                                                                                 count, n_qubits);
  CU_TRY(cudaGetLastError());
  ctx->stats.other_launches++;
  return 0;
}

int dmb_dot(dmb_ctx* ctx, const double* a, const double* b, uint64_t count, double* out_host) {
  if (!ctx || !a || !b || !out_host) return fail("dmb_dot", "null argument");
  DMB_ON_DEVICE(ctx);
  const unsigned blocks = 1024;
  k_dot_partial<<<blocks, 256, 0, ctx->stream>>>(a, b, count, ctx->d_scratch);
  CU_TRY(cudaGetLastError());
  ctx->stats.other_launches++;
  CU_TRY(cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, blocks * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  double acc = 0.0;
  for (unsigned i = 0; i < blocks; ++i) acc += ctx->h_scratch[i];
  *out_host = acc;
  return 0;
}

int dmb_chop(dmb_ctx* ctx, double* state, uint64_t count, double thr) {
  if (!ctx || !state) return fail("dmb_chop", "null argument");
  DMB_ON_DEVICE(ctx);
  k_chop<<<grid_for(count, 256, ctx->sm_count), 256, 0, ctx->stream>>>(state, count, thr);
  CU_TRY(cudaGetLastError());
  ctx->stats.other_launches++;
  return 0;
}

int dmb_upload(dmb_ctx* ctx, double* state, const double* host, uint64_t offset, uint64_t count) {
  if (!ctx || !state || !host) return fail("dmb_upload", "null argument");
  DMB_ON_DEVICE(ctx);
  CU_TRY(cudaMemcpyAsync(state + offset, host, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}

int dmb_download(dmb_ctx* ctx, const double* state, double* host, uint64_t offset, uint64_t count) {
  if (!ctx || !state || !host) return fail("dmb_download", "null argument");
  DMB_ON_DEVICE(ctx);
  CU_TRY(cudaMemcpyAsync(host, state + offset, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dmb_download_async(dmb_ctx* ctx, const double* state, double* host, uint64_t offset, uint64_t count) {
  if (!ctx || !state || !host) return fail("dmb_download_async", "null argument");
  DMB_ON_DEVICE(ctx);
  CU_TRY(cudaMemcpyAsync(host, state + offset, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return 0;
}

}  // extern "C"
