// TEST INFRASTRUCTURE ONLY -- CPU thread-by-thread emulation of the dmb200 kernels.
//
// Implements the C ABI of include/dmb200.h over HOST memory by running the very same
// per-thread bodies (qiskit-aakash_b200/csrc/dm_device.h) that the CUDA kernels in
// dmb200.cu wrap, one "thread" at a time, phase by phase.  It exists so that the index
// arithmetic of the kernels (tile addressing, swizzle, lane mapping, Pauli-pair maths) and
// the host-side scheduler can be unit-tested in the GPU-less build container.  It is built
// into tests/emu/libdmb200_emu.so by tests/emu/build_emu.py and loaded only by the
// `-m "not gpu"` tests through an explicit injection hook; the product never loads it.
#include <math.h>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "dm_device.h"
#include "dm_schedule.h"

static thread_local std::string g_err;
static int fail(const char* what, const char* detail = nullptr) {
  g_err = what;
  if (detail) { g_err += ": "; g_err += detail; }
  return 1;
}

struct dmb_ctx {
  dmb_stats stats;
};

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

static dmb_remote_src g_no_remote;   // zero-initialised: in place

template <int K>
static void run_tile_pass(double* state, int n_bits, const dmb_pass& P, const dmb_remote_src& S = g_no_remote,
                          const dmb_remote_src& D = g_no_remote) {
  constexpr int MAXPAIRS = ((1 << (2 * K - 1)) + DMB_TILE_THREADS - 1) / DMB_TILE_THREADS;
  const uint64_t n_tiles = 1ull << (n_bits - 2 * K);
  alignas(16) static thread_local double smem[1 << 12];
  for (uint64_t tile = 0; tile < n_tiles; ++tile) {
    const uint64_t tbase = dmb_tile_base(tile, P.tile_digit, K);
    for (int t = 0; t < DMB_TILE_THREADS; ++t)
      dmb_tile_load_thread<MAXPAIRS>(t, state, tbase, smem, P.tile_digit, K, S);
    for (int i = 0; i < P.n_ops; ++i)
      for (int t = 0; t < DMB_TILE_THREADS; ++t) dmb_tile_op_thread(t, P.ops[i], smem, K);
    for (int t = 0; t < DMB_TILE_THREADS; ++t)
      dmb_tile_store_thread<MAXPAIRS>(t, state, tbase, smem, P.tile_digit, K, D);
  }
}

// K = 6 path: the per-thread bodies of k_tile_pass6 (paired dispatch, folded swaps), run one virtual
// thread after the other, tiles one after another
static int g_variant = 0;
static thread_local long g_folded_swaps = 0;
static thread_local long g_chained_ops = 0;
static long g_paired_ops = 0;     // thread-ops that went through dmb_lean_op_pair (test hook)
extern "C" long dmb_emu_paired_ops(void) { return g_paired_ops; }
static void run_tile_pass6(double* state, int n_bits, const dmb_pass& P, const dmb_remote_src& S = g_no_remote,
                           const dmb_remote_src& D = g_no_remote) {
  static thread_local dmb_lean_pass L;
  dmb_make_lean_pass(P, n_bits, L, dmb_fold_swaps_enabled() && !S.enabled && !D.enabled);
  g_folded_swaps += P.n_ops - L.n_ops;
  g_chained_ops += L.n_chained;
  alignas(128) static thread_local unsigned char stage[DMB_LEAN_TILE_BYTES];
  static thread_local dmb_lean_thread T[DMB_TILE_THREADS];
  for (int t = 0; t < DMB_TILE_THREADS; ++t) dmb_lean_thread_init(t, L, T[t]);
  dmb_host_mem mem;
  mem.base = stage;
  for (uint64_t tile = 0; tile < L.n_tiles; ++tile) {
    const uint64_t tbase = dmb_tile_base(tile, L.td, DMB_LEAN_K);
    for (int t = 0; t < DMB_TILE_THREADS; ++t) dmb_lean_load_thread(T[t], L, state, tbase, S, mem);
    for (int i = 0; i < L.n_ops;) {
      int used = 1;
      for (int u = 0; u < DMB_TILE_THREADS / 2; ++u) {   // real thread u plays virtual threads 2u, 2u + 1 or u, u + 128
        if (dmb_lean_op_is_paired(L.ops[i])) ++g_paired_ops;
        used = dmb_lean_ops_step<true>(T[2 * u], T[u], &L.ops[i], mem);
      }
      i += used;
    }
    for (int t = 0; t < DMB_TILE_THREADS; ++t) {
      if (D.enabled) dmb_lean_store_thread<true, DMB_ST_PLAIN>(T[t], L, state, tbase, D, mem);
      else if (L.st_mode == DMB_ST_PERM128) dmb_lean_store_thread<false, DMB_ST_PERM128>(T[t], L, state, tbase, D, mem);
      else if (L.st_mode == DMB_ST_SPLIT64) dmb_lean_store_thread<false, DMB_ST_SPLIT64>(T[t], L, state, tbase, D, mem);
      else dmb_lean_store_thread<false, DMB_ST_PLAIN>(T[t], L, state, tbase, D, mem);
    }
  }
}

// ---------------------------------------------------------------------------------------
// The tile kernel's REAL control flow (dmb_half_kernel_body, the function the CUDA kernel
// k_tile_pass6 wraps) on host threads: 128 std::threads per CTA, a CTA barrier, and cp.async emulated as
// deferred copies that only land at wait<N>() -- so a read before the wait, a missing barrier or a wrong stage
// index shows up as a wrong result here, without a GPU.
// ---------------------------------------------------------------------------------------
struct emu_barrier {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, generation = 0, parties;
  explicit emu_barrier(int n) : parties(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    const int gen = generation;
    if (++count == parties) { count = 0; ++generation; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != generation; });
  }
};

struct emu_cta_thread {
  int tid_;
  uint64_t block_, grid_;
  unsigned char* stages;                    // STAGES x 32 KiB, shared by the CTA's threads
  emu_barrier* bar;
  std::vector<std::vector<std::pair<uint32_t, const double*>>> groups;   // open group = groups.back()
  emu_cta_thread() { groups.emplace_back(); }
  int tid() const { return tid_; }
  uint64_t block() const { return block_; }
  uint64_t grid() const { return grid_; }
  void copy16(uint32_t off, const double* src) { groups.back().emplace_back(off, src); }
  void commit() { groups.emplace_back(); }
  template <int N>
  void wait() {                             // all but the N most recently committed groups complete
    while ((int)groups.size() - 1 > N) {
      for (auto& c : groups.front()) memcpy(stages + c.first, c.second, 16);
      groups.erase(groups.begin());
    }
  }
  void sync() { bar->wait(); }
  dmb_host_mem mem(uint32_t off) const {
    dmb_host_mem m;
    m.base = stages + off;
    return m;
  }
};

template <int STMODE, bool PAIRED, int STAGES, int REMOTE = 0>
static void run_half_kernel_cta(double* state, const dmb_lean_pass& L, uint64_t block, uint64_t grid,
                                const dmb_remote_src& R = g_no_remote) {
  std::vector<unsigned char> stages((size_t)STAGES * DMB_LEAN_TILE_BYTES + 128);
  unsigned char* base = stages.data() + (128 - (reinterpret_cast<uintptr_t>(stages.data()) & 127)) % 128;
  emu_barrier bar(DMB_HALF_THREADS);
  std::vector<std::thread> threads;
  for (int t = 0; t < DMB_HALF_THREADS; ++t)
    threads.emplace_back([&, t] {
      emu_cta_thread cx;
      cx.tid_ = t; cx.block_ = block; cx.grid_ = grid; cx.stages = base; cx.bar = &bar;
      dmb_half_kernel_body<STMODE, PAIRED, STAGES, REMOTE>(cx, state, L, R);
    });
  for (auto& th : threads) th.join();
}

template <bool PAIRED, int STAGES>
static void run_half_kernel(double* state, const dmb_lean_pass& L, uint64_t grid) {
  for (uint64_t block = 0; block < grid; ++block) {
    if (L.st_mode == DMB_ST_PERM128) run_half_kernel_cta<DMB_ST_PERM128, PAIRED, STAGES>(state, L, block, grid);
    else if (L.st_mode == DMB_ST_SPLIT64) run_half_kernel_cta<DMB_ST_SPLIT64, PAIRED, STAGES>(state, L, block, grid);
    else run_half_kernel_cta<DMB_ST_PLAIN, PAIRED, STAGES>(state, L, block, grid);
  }
}

// test hook: run `n_passes` K = 6 passes through the threaded kernel body (paired: 0/1, stages: 1/2, grid CTAs)
extern "C" int dmb_emu_run_half_kernel(double* state, int n_bits, const dmb_pass* passes, size_t n_passes, int paired,
                                       int stages, int grid) {
  static thread_local dmb_lean_pass L;
  for (size_t i = 0; i < n_passes; ++i) {
    if (passes[i].n_tile_digits != DMB_LEAN_K) return 1;
    dmb_make_lean_pass(passes[i], n_bits, L, dmb_fold_swaps_enabled());
    uint64_t g = (uint64_t)grid < L.n_tiles ? (uint64_t)grid : L.n_tiles;
    if (paired && stages == 2) run_half_kernel<true, 2>(state, L, g);
    else if (paired) run_half_kernel<true, 1>(state, L, g);
    else if (stages == 2) run_half_kernel<false, 2>(state, L, g);
    else run_half_kernel<false, 1>(state, L, g);
  }
  return 0;
}

// test hook: ONE K = 6 pass through the threaded kernel body in pull mode (REMOTE = 1: every 16-byte pair is read through
// the source table, whose index is gathered from the selected index bits and composed from per-tile / per-thread /
// per-pair parts) -- the path the CUDA kernel takes for a fused exchange; compared with dmb_apply_pass_remote_sel
extern "C" int dmb_emu_run_half_kernel_pull(double* dst_state, int n_bits, const dmb_pass* pass, const uint64_t* tab,
                                            int n_sel, const int32_t* sel, int grid) {
  static thread_local dmb_lean_pass L;
  if (pass->n_tile_digits != DMB_LEAN_K || n_sel < 0 || n_sel > DMB_REMOTE_BITS) return 1;
  dmb_make_lean_pass(*pass, n_bits, L);
  dmb_remote_src R;
  memset(&R, 0, sizeof(R));
  for (int i = 0; i < (1 << n_sel); ++i) R.tab[i] = tab[i];
  R.n_sel = n_sel;
  for (int j = 0; j < n_sel; ++j) R.sel[j] = sel[j];
  R.enabled = 1;
  const uint64_t g = (uint64_t)grid < L.n_tiles ? (uint64_t)grid : L.n_tiles;
  for (uint64_t block = 0; block < g; ++block) run_half_kernel_cta<DMB_ST_PLAIN, true, 1, 1>(dst_state, L, block, g, R);
  return 0;
}

// test hook: how many ops of a K = 6 pass the library chains to their predecessor (dmb_chain_ops), and the op order it runs
extern "C" int dmb_emu_pass_chained(const dmb_pass* pass, int n_bits, int32_t* order_out) {
  static thread_local dmb_lean_pass L;
  if (pass->n_tile_digits != DMB_LEAN_K) return -1;
  dmb_make_lean_pass(*pass, n_bits, L, dmb_fold_swaps_enabled());
  if (order_out) {
    bool chain[DMB_MAX_OPS];
    dmb_chain_ops(*pass, L.n_ops, order_out, chain);
  }
  return L.n_chained;
}

// test hook: thread order chosen for a SPLIT64 store + worst number of half-warp lanes per bank slot
extern "C" int dmb_emu_split_order(const int32_t* perm, int32_t* tbit_out) {
  int ibit[3];
  for (int k = 0; k < 8; ++k) tbit_out[k] = k + 1;
  dmb_choose_split_order(perm, tbit_out, ibit);
  int worst = 0;
  for (int half = 0; half < 2; ++half)
    for (int lo = 0; lo < 2; ++lo) {
      int count[16] = {0};
      for (int lane = 0; lane < 16; ++lane) {
        uint32_t l2 = (uint32_t)lo | ((uint32_t)half << tbit_out[4]);
        for (int k = 0; k < 4; ++k) l2 |= (((uint32_t)lane >> k) & 1u) << tbit_out[k];
        ++count[dmb_swz(dmb_st_source(l2, perm)) & 15u];
      }
      for (int sl = 0; sl < 16; ++sl) if (count[sl] > worst) worst = count[sl];
    }
  return worst;
}

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

int dmb_create(int, dmb_ctx** out) {
  if (!out) return fail("dmb_create", "null out pointer");
  *out = new dmb_ctx();
  memset(&(*out)->stats, 0, sizeof(dmb_stats));
  return 0;
}
int dmb_destroy(dmb_ctx* ctx) { delete ctx; return 0; }
int dmb_set_stream(dmb_ctx*, void*) { return 0; }
int dmb_sync(dmb_ctx*) { return 0; }
int dmb_get_stats(dmb_ctx* ctx, dmb_stats* out) { *out = ctx->stats; return 0; }
int dmb_reset_stats(dmb_ctx* ctx) { memset(&ctx->stats, 0, sizeof(dmb_stats)); return 0; }
int dmb_set_tile_variant(dmb_ctx*, int variant) {
  if (variant < 0 || variant > 1) return fail("dmb_set_tile_variant", "variant must be 0 or 1");
  g_variant = variant;
  return 0;
}

int dmb_init_product(dmb_ctx* ctx, double* state, int n_bits, uint64_t rank_bits, int n_qubits,
                     const int32_t* hi, const int32_t* lo, const double* v, double scale) {
  if (n_qubits < 1 || n_qubits > DMB_MAX_QUBITS) return fail("dmb_init_product", "n_qubits out of range");
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
  for (uint64_t i = 0; i < (1ull << n_bits); ++i) state[i] = dmb_init_value(i, p);
  ctx->stats.other_launches++;
  return 0;
}

int dmb_apply_passes(dmb_ctx* ctx, double* state, int n_bits, const dmb_pass* passes, size_t n_passes) {
  for (size_t i = 0; i < n_passes; ++i) {
    if (validate_pass(passes[i], n_bits)) return 1;
    const dmb_pass& P = passes[i];
    switch (P.n_tile_digits) {
      case 2: run_tile_pass<2>(state, n_bits, P); break;
      case 3: run_tile_pass<3>(state, n_bits, P); break;
      case 4: run_tile_pass<4>(state, n_bits, P); break;
      case 5: run_tile_pass<5>(state, n_bits, P); break;
      case 6:
        if (g_variant == 1) run_tile_pass<6>(state, n_bits, P);
        else { const long before = g_folded_swaps, chained_before = g_chained_ops; run_tile_pass6(state, n_bits, P);
               ctx->stats.folded_swaps += (uint64_t)(g_folded_swaps - before);
               ctx->stats.chained_ops += (uint64_t)(g_chained_ops - chained_before); }
        break;
      default: return fail("dmb_apply_passes", "unsupported tile size");
    }
    ctx->stats.tile_pass_launches++;
    ctx->stats.fused_ops += (uint64_t)P.n_ops;
    ctx->stats.state_bytes_moved += 16ull << n_bits;
  }
  return 0;
}

int dmb_apply_pass_remote_sel(dmb_ctx* ctx, double* dst_state, int n_bits, const dmb_pass* pass,
                              const uint64_t* src_tab, int n_sel, const int32_t* sel_bits, int push);

int dmb_apply_pass_remote(dmb_ctx* ctx, double* dst_state, int n_bits, const dmb_pass* pass,
                          const uint64_t* src_tab, int tab_bits, int block_shift, int push) {
  if (tab_bits < 0 || tab_bits > DMB_REMOTE_BITS) return fail("dmb_apply_pass_remote", "table too large");
  if (block_shift + tab_bits != n_bits) return fail("dmb_apply_pass_remote", "block_shift + tab_bits != n_bits");
  int32_t sel[DMB_REMOTE_BITS];
  for (int j = 0; j < tab_bits; ++j) sel[j] = block_shift + j;
  return dmb_apply_pass_remote_sel(ctx, dst_state, n_bits, pass, src_tab, tab_bits, sel, push);
}

int dmb_apply_pass_remote_sel(dmb_ctx* ctx, double* dst_state, int n_bits, const dmb_pass* pass,
                              const uint64_t* src_tab, int n_sel, const int32_t* sel_bits, int push) {
  if (n_sel < 0 || n_sel > DMB_REMOTE_BITS) return fail("dmb_apply_pass_remote", "table too large");
  for (int j = 0; j < n_sel; ++j)
    if (sel_bits[j] < 0 || sel_bits[j] >= n_bits) return fail("dmb_apply_pass_remote", "selected bit outside the local index");
  if (validate_pass(*pass, n_bits)) return 1;
  const dmb_pass& P = *pass;
  dmb_remote_src S;
  memset(&S, 0, sizeof(S));
  for (int i = 0; i < (1 << n_sel); ++i) S.tab[i] = src_tab[i];
  S.n_sel = n_sel;
  for (int j = 0; j < n_sel; ++j) S.sel[j] = sel_bits[j];
  S.enabled = 1;
  const dmb_remote_src& ld = push ? g_no_remote : S;
  const dmb_remote_src& st = push ? S : g_no_remote;
  switch (P.n_tile_digits) {
    case 2: run_tile_pass<2>(dst_state, n_bits, P, ld, st); break;
    case 3: run_tile_pass<3>(dst_state, n_bits, P, ld, st); break;
    case 4: run_tile_pass<4>(dst_state, n_bits, P, ld, st); break;
    case 5: run_tile_pass<5>(dst_state, n_bits, P, ld, st); break;
    case 6: run_tile_pass6(dst_state, n_bits, P, ld, st); break;
    default: return fail("dmb_apply_pass_remote", "unsupported tile size");
  }
  ctx->stats.tile_pass_launches++;
  ctx->stats.fused_ops += (uint64_t)P.n_ops;
  ctx->stats.state_bytes_moved += 16ull << n_bits;
  return 0;
}

// in the emulation "peers" are threads of one process: a handle is just the pointer
int dmb_ipc_export(dmb_ctx*, const void* dev_ptr, unsigned char* handle64, uint64_t* offset) {
  memset(handle64, 0, 64);
  memcpy(handle64, &dev_ptr, sizeof(dev_ptr));
  *offset = 0;
  return 0;
}

int dmb_ipc_open(dmb_ctx*, const unsigned char* handle64, uint64_t offset, void** out_ptr) {
  void* p = nullptr;
  memcpy(&p, handle64, sizeof(p));
  *out_ptr = (char*)p + offset;
  return 0;
}

int dmb_ipc_close(dmb_ctx*, const unsigned char*) { return 0; }

int dmb_marginal(dmb_ctx* ctx, const double* state, int n_bits, uint64_t rank_bits, int n_qubits,
                 const int32_t* hi, const int32_t* lo, const double* wt, double* out) {
  if (n_qubits < 1 || n_qubits > DMB_MAX_QUBITS) return fail("dmb_marginal", "n_qubits out of range");
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
  for (uint64_t c = 0; c < (1ull << n_qubits); ++c) out[c] = dmb_marginal_value(c, state, p);
  ctx->stats.other_launches++;
  return 0;
}

int dmb_fwht(dmb_ctx* ctx, double* vec, int n_qubits) {
  const uint64_t pairs = (1ull << n_qubits) >> 1;
  for (int st = 0; st < n_qubits; ++st)
    for (uint64_t t = 0; t < pairs; ++t) dmb_fwht_pair(t, st, vec);
  ctx->stats.other_launches++;
  return 0;
}

int dmb_contract_digit(dmb_ctx* ctx, const double* in, double* out, uint64_t H, uint64_t L, const double nv[3]) {
  for (uint64_t t = 0; t < H * L; ++t) dmb_contract_elem(t, in, out, L, nv[0], nv[1], nv[2]);
  ctx->stats.other_launches++;
  return 0;
}

int dmb_read_coeffs(dmb_ctx* ctx, const double* state, const uint64_t* idx, size_t k, double* out_host) {
  for (size_t i = 0; i < k; ++i) out_host[i] = state[idx[i]];
  ctx->stats.other_launches++;
  return 0;
}

int dmb_to_matrix(dmb_ctx* ctx, const double* state, int n_qubits, double* work, double* out) {
  const uint64_t vecs = 1ull << (2 * n_qubits - 2);
  for (int pos = 0; pos < n_qubits; ++pos)
    for (uint64_t t = 0; t < vecs; ++t) dmb_tomatrix_digit(t, pos, pos == 0, state, (dmb_d2*)work);
  for (uint64_t t = 0; t < (1ull << (2 * n_qubits)); ++t)
    dmb_tomatrix_scatter(t, n_qubits, (const dmb_d2*)work, (dmb_d2*)out);
  ctx->stats.other_launches++;
  return 0;
}

int dmb_dot(dmb_ctx* ctx, const double* a, const double* b, uint64_t count, double* out_host) {
  double acc = 0.0;
  for (uint64_t i = 0; i < count; ++i) acc += a[i] * b[i];
  *out_host = acc;
  ctx->stats.other_launches++;
  return 0;
}

int dmb_chop(dmb_ctx* ctx, double* state, uint64_t count, double thr) {
  for (uint64_t i = 0; i < count; ++i)
    if (fabs(state[i]) < thr) state[i] = 0.0;
  ctx->stats.other_launches++;
  return 0;
}

int dmb_upload(dmb_ctx*, double* state, const double* host, uint64_t offset, uint64_t count) {
  memcpy(state + offset, host, count * sizeof(double));
  return 0;
}

int dmb_download(dmb_ctx*, const double* state, double* host, uint64_t offset, uint64_t count) {
  memcpy(host, state + offset, count * sizeof(double));
  return 0;
}

int dmb_download_async(dmb_ctx*, const double* state, double* host, uint64_t offset, uint64_t count) {
  memcpy(host, state + offset, count * sizeof(double));
  return 0;
}

}  // extern "C"
