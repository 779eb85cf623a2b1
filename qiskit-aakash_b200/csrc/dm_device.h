// dm_device.h -- per-thread bodies of the dmb200 kernels.
//
// Every kernel in dmb200.cu is "phase structured": a phase is a function of the thread
// index that touches shared/global memory, phases are separated by block barriers.  The
// bodies live here as __host__ __device__ functions so that the index arithmetic (tile
// addressing, shared-memory swizzle, per-op thread mapping, Pauli-pair arithmetic) is one
// piece of code: dmb200.cu wraps it in __global__ kernels for sm_100a; the CPU unit tests
// (tests/emu) run the same bodies thread-by-thread to check the arithmetic where no GPU is
// available.  Nothing here is a fallback path of the product.
#pragma once
#include <stdint.h>
#include <string.h>
#include "dmb200.h"
#include <cstdlib>

#if defined(__CUDACC__)
#define DMB_HD __host__ __device__ __forceinline__
#else
#define DMB_HD inline
#endif

#define DMB_TILE_THREADS 256

struct alignas(16) dmb_d2 { double x, y; };

// ---------------------------------------------------------------------------------------
// Tile addressing
// ---------------------------------------------------------------------------------------
// A tile is addressed by a 2K-bit local index l: digit j of l (bits 2j,2j+1) is the value
// of state digit tile_digit[j].  In shared memory element l lives at word dmb_swz(l).  The
// swizzle keeps 16-byte pairs (l0) intact -- so tiles can be filled by 128-bit loads or
// 16-byte cp.async -- and XORs the 16-byte chunk index inside each 128-byte row (bits 3:1)
// with a linear function of the higher digits D2..D5:
//     chunk[2:1] ^= D2 ^ D3 ^ D4 ^ D5          chunk[0] ^= parity(D3) ^ parity(D5)
// With it every op is bank-conflict free for a suitable lane order (dmb_op.fd, chosen by
// the host, see schedule.py):
//   (A) op digits exclude digit 0: 64-bit accesses, lanes[1:0] = digit 0, lanes[3:2] = any
//       other free digit -> the 16 lanes of a half-warp hit 16 distinct 8-byte slots;
//   (B) op digits include digit 0: the thread's elements along digit 0 are two 16-byte
//       chunks, accessed as 128-bit words; lanes[1:0] = digit 1 (or 2 if 1 is an op digit),
//       lane[2] = low bit of digit 3 (or 5) -> the 8 lanes of a quarter-warp hit 8 distinct
//       chunks.
DMB_HD uint32_t dmb_swz(uint32_t l) {
  const uint32_t d2 = (l >> 4) & 3u, d3 = (l >> 6) & 3u, d4 = (l >> 8) & 3u, d5 = (l >> 10) & 3u;
  const uint32_t hi = d2 ^ d3 ^ d4 ^ d5;
  const uint32_t odd = d3 ^ d5;
  return l ^ (hi << 2) ^ (((odd ^ (odd >> 1)) & 1u) << 1);
}

// tile number -> offset of the tile's element 0: insert a zero digit at every tile digit.
DMB_HD uint64_t dmb_tile_base(uint64_t tile, const int32_t* td, int K) {
  uint64_t x = tile;
  for (int j = 0; j < K; ++j) {
    const int sh = 2 * td[j];
    const uint64_t low = x & ((1ull << sh) - 1ull);
    x = ((x >> sh) << (sh + 2)) | low;
  }
  return x;
}

// local index -> offset relative to the tile base.
DMB_HD uint64_t dmb_tile_off(uint32_t l, const int32_t* td, int K) {
  uint64_t off = 0;
  for (int j = 0; j < K; ++j) off |= (uint64_t)((l >> (2 * j)) & 3u) << (2 * td[j]);
  return off;
}

// Fused exchange ("pull"): in the pass that follows a global<->local slot swap, element idx of
// the NEW local layout is read straight from the rank that holds it in the OLD layout --
// tab[dmb_remote_index(idx)] is that rank's (peer-mapped) buffer address, pre-offset so that the
// element sits at tab[..] + idx -- and the result is written to the local destination buffer.
// One kernel does the all-to-all and the first fused ops; enabled == 0 means in place.
// The table index is gathered from up to DMB_REMOTE_BITS SELECTED bits of the element index (sel[j] = position of the
// bit that becomes bit j of the table index): the high bits when the global slots are swapped with the top local slots,
// scattered bits when a global slot is swapped with an arbitrary local slot (no parking pass, distributed.py).
#define DMB_REMOTE_MAX 32
#define DMB_REMOTE_BITS 5
struct dmb_remote_src {
  uint64_t tab[DMB_REMOTE_MAX];
  int32_t n_sel;
  int32_t enabled;
  int32_t sel[DMB_REMOTE_BITS];
  int32_t pad_;
};

DMB_HD uint32_t dmb_remote_index(const dmb_remote_src& S, uint64_t idx) {
  uint32_t k = 0;
#pragma unroll
  for (int j = 0; j < DMB_REMOTE_BITS; ++j)
    if (j < S.n_sel) k |= (uint32_t)((idx >> S.sel[j]) & 1ull) << j;
  return k;
}

DMB_HD const double* dmb_src_ptr(const dmb_remote_src& S, const double* local, uint64_t idx) {
  if (!S.enabled) return local + idx;
  return reinterpret_cast<const double*>(S.tab[dmb_remote_index(S, idx)]) + idx;
}

// Load phase: thread t moves the 16-byte pairs p = t, t+256, ... (tile_digit[0] == 0, so the
// two elements of a pair are adjacent in global memory and share a 16-byte chunk of smem).
template <int MAXPAIRS>
DMB_HD void dmb_tile_load_thread(int t, const double* __restrict__ state, uint64_t tile_base, double* smem,
                                 const int32_t* td, int K, const dmb_remote_src& S) {
  const uint32_t npairs = 1u << (2 * K - 1);
  dmb_d2 v[MAXPAIRS];
#pragma unroll
  for (int i = 0; i < MAXPAIRS; ++i) {
    const uint32_t p = (uint32_t)t + (uint32_t)i * DMB_TILE_THREADS;
    if (p < npairs)
      v[i] = *reinterpret_cast<const dmb_d2*>(dmb_src_ptr(S, state, tile_base + dmb_tile_off(2u * p, td, K)));
  }
#pragma unroll
  for (int i = 0; i < MAXPAIRS; ++i) {
    const uint32_t p = (uint32_t)t + (uint32_t)i * DMB_TILE_THREADS;
    if (p < npairs) {
      *reinterpret_cast<dmb_d2*>(smem + dmb_swz(2u * p)) = v[i];
    }
  }
}

// "push" flavour of the fused exchange: the pass BEFORE a slot swap writes every tile straight
// into the buffer of the rank that owns it in the new layout (same table convention).
DMB_HD double* dmb_dst_ptr(const dmb_remote_src& D, double* local, uint64_t idx) {
  if (!D.enabled) return local + idx;
  return reinterpret_cast<double*>(D.tab[dmb_remote_index(D, idx)]) + idx;
}

template <int MAXPAIRS>
DMB_HD void dmb_tile_store_thread(int t, double* __restrict__ state, uint64_t tile_base, const double* smem,
                                  const int32_t* td, int K, const dmb_remote_src& D) {
  const uint32_t npairs = 1u << (2 * K - 1);
#pragma unroll
  for (int i = 0; i < MAXPAIRS; ++i) {
    const uint32_t p = (uint32_t)t + (uint32_t)i * DMB_TILE_THREADS;
    if (p < npairs) {
      const dmb_d2 w = *reinterpret_cast<const dmb_d2*>(smem + dmb_swz(2u * p));
      *reinterpret_cast<dmb_d2*>(dmb_dst_ptr(D, state, tile_base + dmb_tile_off(2u * p, td, K))) = w;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Pauli-pair arithmetic on one 16-element block  v[i][j]  (i: digit a, j: digit b)
// ---------------------------------------------------------------------------------------
// rows 1..3 of a single-qubit map along the first index (x0 is the I component, unchanged)
DMB_HD void dmb_mat3(const double* __restrict__ m, double& x0, double& x1, double& x2, double& x3) {
  const double y1 = m[0] * x0 + m[1] * x1 + m[2] * x2 + m[3] * x3;
  const double y2 = m[4] * x0 + m[5] * x1 + m[6] * x2 + m[7] * x3;
  const double y3 = m[8] * x0 + m[9] * x1 + m[10] * x2 + m[11] * x3;
  x1 = y1; x2 = y2; x3 = y3;
}

// CNOT with the transition-selective-pulse error model, control = first index, target =
// second index.  Restates basicaertools.py:346-363 (the q_2 > q_1 branch; the other branch
// is the same map with the axes exchanged).  I,X,Y,Z = 0,1,2,3.
DMB_HD void dmb_cx_tsp(double (&v)[4][4], double c, double s, double c2, double s2, double cs) {
  const double iy = v[0][2], zy = v[3][2], iz = v[0][3], zz = v[3][3];
  const double dz = iz - zz, dy = iy - zy;
  v[0][2] = s2 * iy + c2 * zy - cs * dz;
  v[3][2] = c2 * iy + s2 * zy + cs * dz;
  v[0][3] = s2 * iz + c2 * zz + cs * dy;
  v[3][3] = c2 * iz + s2 * zz - cs * dy;
  const double xi = v[1][0], xx = v[1][1], xy = v[1][2], xz = v[1][3];
  const double yi = v[2][0], yx = v[2][1], yy = v[2][2], yz = v[2][3];
  v[1][0] = c * xx - s * yi;
  v[1][1] = c * xi - s * yx;
  v[1][2] = -s * yy + c * yz;
  v[1][3] = -c * yy - s * yz;
  v[2][0] = s * xi + c * yx;
  v[2][1] = s * xx + c * yi;
  v[2][2] = s * xy - c * xz;
  v[2][3] = c * xy + s * xz;
}

// TSP CNOT with zero mean angle error (s = cs = 0): half the arithmetic.
DMB_HD void dmb_cx_tsp0(double (&v)[4][4], double c, double c2, double s2) {
  const double iy = v[0][2], zy = v[3][2], iz = v[0][3], zz = v[3][3];
  v[0][2] = s2 * iy + c2 * zy;
  v[3][2] = c2 * iy + s2 * zy;
  v[0][3] = s2 * iz + c2 * zz;
  v[3][3] = c2 * iz + s2 * zz;
  const double xi = v[1][0], xx = v[1][1], xy = v[1][2], xz = v[1][3];
  const double yi = v[2][0], yx = v[2][1], yy = v[2][2], yz = v[2][3];
  v[1][0] = c * xx; v[1][1] = c * xi; v[1][2] = c * yz; v[1][3] = -(c * yy);
  v[2][0] = c * yx; v[2][1] = c * yi; v[2][2] = -(c * xz); v[2][3] = c * xy;
}

// Ideal CNOT: the same map with (c,s,c2,s2,cs) = (1,0,1,0,0) -- a signed permutation.
DMB_HD void dmb_cx_ideal(double (&v)[4][4]) {
  double t;
  t = v[0][2]; v[0][2] = v[3][2]; v[3][2] = t;      // IY <-> ZY
  t = v[0][3]; v[0][3] = v[3][3]; v[3][3] = t;      // IZ <-> ZZ
  t = v[1][0]; v[1][0] = v[1][1]; v[1][1] = t;      // XI <-> XX
  t = v[2][0]; v[2][0] = v[2][1]; v[2][1] = t;      // YI <-> YX
  const double xy = v[1][2], xz = v[1][3], yy = v[2][2], yz = v[2][3];
  v[1][2] = yz; v[1][3] = -yy; v[2][2] = -xz; v[2][3] = xy;
}

// Op phase: thread t owns the 16-block number t of the tile (4^(K-2) blocks).
DMB_HD void dmb_tile_op_thread(int t, const dmb_op& op, double* smem, int K) {
  const int nfree = K - 2;
  if (t >= (1 << (2 * nfree))) return;
  uint32_t bl = 0;
  for (int m = 0; m < nfree; ++m) bl |= (uint32_t)((t >> (2 * m)) & 3) << (2 * op.fd[m]);
  const uint32_t sb = dmb_swz(bl);
  uint32_t sa[4], sj[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sa[i] = dmb_swz((uint32_t)i << (2 * op.a));
    sj[i] = dmb_swz((uint32_t)i << (2 * op.b));
  }
  double v[4][4];
  if (op.a == 0) {            // mode (B): pairs along i
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const dmb_d2 p0 = *reinterpret_cast<const dmb_d2*>(smem + (sb ^ sj[j]));
      const dmb_d2 p1 = *reinterpret_cast<const dmb_d2*>(smem + (sb ^ sj[j] ^ 2u));
      v[0][j] = p0.x; v[1][j] = p0.y; v[2][j] = p1.x; v[3][j] = p1.y;
    }
  } else if (op.b == 0) {     // mode (B): pairs along j
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const dmb_d2 p0 = *reinterpret_cast<const dmb_d2*>(smem + (sb ^ sa[i]));
      const dmb_d2 p1 = *reinterpret_cast<const dmb_d2*>(smem + (sb ^ sa[i] ^ 2u));
      v[i][0] = p0.x; v[i][1] = p0.y; v[i][2] = p1.x; v[i][3] = p1.y;
    }
  } else {                    // mode (A)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = smem[sb ^ sa[i] ^ sj[j]];
  }

  if (op.flags & DMB_HAS_PA) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dmb_mat3(op.pa, v[0][j], v[1][j], v[2][j], v[3][j]);
  }
  if (op.flags & DMB_HAS_PB) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmb_mat3(op.pb, v[i][0], v[i][1], v[i][2], v[i][3]);
  }
  switch (op.kind) {
    case DMB_OP_CX:
      dmb_cx_ideal(v);
      break;
    case DMB_OP_CX_TSP:
      dmb_cx_tsp(v, op.coef[0], op.coef[1], op.coef[2], op.coef[3], op.coef[4]);
      break;
    case DMB_OP_DIAG2:
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] *= op.coef[4 * i + j];
      break;
    case DMB_OP_SWAP:
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) { const double tmp = v[i][j]; v[i][j] = v[j][i]; v[j][i] = tmp; }
      break;
    default:
      break;
  }
  if (op.a == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dmb_d2 p0, p1;
      p0.x = v[0][j]; p0.y = v[1][j]; p1.x = v[2][j]; p1.y = v[3][j];
      *reinterpret_cast<dmb_d2*>(smem + (sb ^ sj[j])) = p0;
      *reinterpret_cast<dmb_d2*>(smem + (sb ^ sj[j] ^ 2u)) = p1;
    }
  } else if (op.b == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dmb_d2 p0, p1;
      p0.x = v[i][0]; p0.y = v[i][1]; p1.x = v[i][2]; p1.y = v[i][3];
      *reinterpret_cast<dmb_d2*>(smem + (sb ^ sa[i])) = p0;
      *reinterpret_cast<dmb_d2*>(smem + (sb ^ sa[i] ^ 2u)) = p1;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) smem[sb ^ sa[i] ^ sj[j]] = v[i][j];
  }
}

// ---------------------------------------------------------------------------------------
// Lean K = 6 path (the hot kernel, k_tile_pass6 in dmb200.cu)
// ---------------------------------------------------------------------------------------
// The generic bodies above recompute every address from the digit lists; ncu showed that
// version issue-bound (69 % issue-slot utilisation, FP64 pipe 27 %, DRAM 40 %).  Everything
// that does not depend on the thread is therefore precomputed on the host, once per pass,
// into dmb_lean_pass (a kernel parameter, read through the uniform datapath):
//   * load/store: element l = 2t + 512 i of a tile splits into a thread part (2t, bits 0-8)
//     and a uniform part (i << 9, bits 9-11); both the global offset and the swizzle are
//     linear in those disjoint bit fields, so address = thread_part (+|^) table[i];
//   * op: byte offsets sa[i] / sj[j] of the digit values (already swizzled; the swizzle is
//     XOR-linear), the shifts that scatter the thread's four index digits onto the free
//     tile digits, the access mode, and "column 0 is zero" flags that drop 3 of 12 DFMAs
//     per 4-vector when a map has no I-component admixture (no amplitude damping / reset).
#define DMB_LEAN_K 6
#define DMB_LEAN_TILE (1u << (2 * DMB_LEAN_K))          // 4096 elements
#define DMB_LEAN_TILE_BYTES (DMB_LEAN_TILE * 8u)        // 32 KiB
#define DMB_LEAN_PAIRS 8                                 // 16-byte pairs per thread per tile
enum { DMB_MODE_A = 0, DMB_MODE_PAIR_A = 1, DMB_MODE_PAIR_B = 2 };
enum { DMB_PA_COL0 = 4, DMB_PB_COL0 = 8, DMB_TSP_ZERO_MEAN = 16, DMB_PAIRABLE = 32,
       DMB_CHAIN = 64 };   // extra flag bits (library-internal); DMB_CHAIN: the NEXT op acts on the same digit pair and runs in
                           // the same shared-memory round trip (dmb_chain_ops below)

struct alignas(16) dmb_lean_op {
  uint32_t sa[4];      // swizzled BYTE offset of digit-a value i
  uint32_t sj[4];      // swizzled BYTE offset of digit-b value j
  uint32_t sh[4];      // left shift (bits) placing thread digit m at its tile digit
  int32_t kind, flags, mode, variant;   // variant: compile-time specialisation id, -1 = generic
  double pa[12], pb[12], coef[16];
};

struct alignas(16) dmb_lean_pass {
  uint64_t n_tiles;
  int32_t n_ops;
  int32_t td[DMB_LEAN_K];
  int32_t n_chained;                    // ops that run chained to their predecessor (DMB_CHAIN): round trips saved
  uint64_t pair_goff[DMB_LEAN_PAIRS];   // element offset of the uniform part (i << 9)
  uint32_t pair_soff[DMB_LEAN_PAIRS];   // swizzled byte offset of the uniform part
  // relabelling store (trailing SWAP ops folded into the write-back, see dmb_make_lean_pass):
  // store-index digit k is read from tile digit st_perm[k]
  int32_t st_mode;                      // DMB_ST_PLAIN / DMB_ST_PERM128 / DMB_ST_SPLIT64
  int32_t st_perm[DMB_LEAN_K];
  uint32_t st_delta;                    // SPLIT64: byte-address XOR between the two elements of a pair
  uint32_t st_pair_soff[DMB_LEAN_PAIRS];
  // SPLIT64 walks the tile in its own thread order (bit k of the thread index is bit st_tbit[k] of
  // the store index, the pair counter supplies the rest): chosen by the host so that the 16 lanes
  // of a half-warp read 16 distinct 8-byte bank slots while a warp still writes whole 128-byte lines
  int32_t st_tbit[8];
  uint64_t st_pair_goff[DMB_LEAN_PAIRS];
  dmb_lean_op ops[DMB_MAX_OPS];
};
enum { DMB_ST_PLAIN = 0, DMB_ST_PERM128 = 1, DMB_ST_SPLIT64 = 2 };

// tile-local index of the element that store index l2 (digits in NEW global order) reads
DMB_HD uint32_t dmb_st_source(uint32_t l2, const int32_t* perm) {
  uint32_t l = 0;
  for (int k = 0; k < DMB_LEAN_K; ++k) l |= ((l2 >> (2 * k)) & 3u) << (2 * perm[k]);
  return l;
}

// Specialisation ids: the hot (kind, matrix class, access mode) combinations get straight-line
// code (no flag branches, no register moves at control-flow merges: 302 -> ~190 issued
// instructions per op); everything else runs the generic body.
#define DMB_KIND_TSP0 5
DMB_HD int dmb_variant_id(int kindx, int ma, int mb, int mode) { return ((kindx * 3 + ma) * 3 + mb) * 3 + mode; }
inline bool dmb_variant_is_specialised(int kindx, int ma, int mb) {
  if (ma == mb && (ma == 1 || ma == 2))
    return kindx == DMB_OP_MATS || kindx == DMB_OP_CX || kindx == DMB_OP_CX_TSP || kindx == DMB_KIND_TSP0;
  if (ma == 0 && mb == 0) return kindx == DMB_OP_CX || kindx == DMB_OP_SWAP;
  return false;
}

// host-side conversion (runs once per pass, before the launch)
// With `fold_swaps` the SWAP ops at the END of the op list are not executed in shared memory:
// a swap of two tile digits only renames which global digit position each of them is written
// back to, so it is folded into the store addressing (the tile covers the same addresses).
// The store then walks the tile in the NEW global order (coalescing unchanged: same global
// offset tables) and gathers from the permuted shared-memory addresses -- as 128-bit pairs while
// tile digit 0 stays in place, as two 64-bit reads per pair otherwise.  Saves one 64 KiB
// shared-memory round trip per swap (0.13 ms at n = 14).
// Thread order of the SPLIT64 store.  Store-index bits 1..3 (high bit of the new digit 0 and the
// new digit 1) must be lane bits so that a warp's 16-byte stores fill whole 128-byte lines; the
// other two lane bits are free.  Among the orders that satisfy this, take one whose half-warps
// (lanes 0-15, 64-bit loads) hit the fewest equal bank slots -- the swizzle spreads the OLD digits
// 0 and 1 over the slots, and after the swap old digit 0 is a high digit of the store index.
inline void dmb_choose_split_order(const int32_t* perm, int32_t* tbit, int* ibit) {
  int best_cost = 1 << 30, best[5] = {1, 2, 3, 4, 5};
  for (int e0 = 4; e0 <= 11; ++e0)
    for (int e1 = e0 + 1; e1 <= 11; ++e1) {
      const int five[5] = {1, 2, 3, e0, e1};
      for (int top = 0; top < 5; ++top) {              // which of the five is lane bit 4
        int lanes[5], w = 0;
        for (int k = 0; k < 5; ++k) if (k != top) lanes[w++] = five[k];
        lanes[4] = five[top];
        int cost = 0;
        for (int lo = 0; lo < 2; ++lo) {
          int count[16] = {0};
          for (int lane = 0; lane < 16; ++lane) {
            uint32_t l2 = (uint32_t)lo;
            for (int k = 0; k < 4; ++k) l2 |= (((uint32_t)lane >> k) & 1u) << lanes[k];
            ++count[dmb_swz(dmb_st_source(l2, perm)) & 15u];
          }
          for (int sl = 0; sl < 16; ++sl) if (count[sl] > cost) cost = count[sl];
        }
        if (cost < best_cost) { best_cost = cost; for (int k = 0; k < 5; ++k) best[k] = lanes[k]; }
      }
    }
  bool used[12] = {false};
  used[0] = true;
  for (int k = 0; k < 5; ++k) { tbit[k] = best[k]; used[best[k]] = true; }
  int next = 5;
  for (int b = 1; b <= 11; ++b) {
    if (used[b]) continue;
    if (next < 8) tbit[next++] = b; else ibit[next++ - 8] = b;
  }
}

inline bool dmb_fold_swaps_enabled() {         // DMB_FOLD_SWAPS=0 keeps trailing swaps as shared-memory ops (A/B switch)
  static const bool on = [] { const char* e = getenv("DMB_FOLD_SWAPS"); return !(e && e[0] == '0'); }();
  return on;
}

inline bool dmb_chain_ops_enabled() {          // DMB_CHAIN_OPS=0 runs every op in its own shared-memory round trip (A/B switch)
  static const bool on = [] { const char* e = getenv("DMB_CHAIN_OPS"); return !(e && e[0] == '0'); }();
  return on;
}

inline int dmb_op_kindx(const dmb_op& o) {     // kind with the zero-mean TSP CNOT as a kind of its own
  return (o.kind == DMB_OP_CX_TSP && o.coef[1] == 0.0 && o.coef[4] == 0.0) ? DMB_KIND_TSP0 : o.kind;
}

// kinds whose chained bodies are instantiated (DMB_CHAIN_VARIANTS): the ideal CNOT and both TSP CNOTs
inline bool dmb_kind_is_chainable(int kindx) { return kindx == DMB_OP_CX || kindx == DMB_KIND_TSP0 || kindx == DMB_OP_CX_TSP; }

inline int dmb_op_map_class(const dmb_op& o, int which) {     // 0 no map, 1 map without column 0, 2 full map
  if (!(o.flags & (which ? DMB_HAS_PB : DMB_HAS_PA))) return 0;
  const double* m = which ? o.pb : o.pa;
  return (m[0] != 0.0 || m[4] != 0.0 || m[8] != 0.0) ? 2 : 1;
}

// Which ops of a pass share a round trip (DMB_CHAIN).  order[k] = index into P.ops of the op at position k;
// chain[k] = the op at position k + 1 runs chained to the one at k.  A later op on the same ordered digit pair and of
// the same CNOT kind is hoisted next to its partner when no op in between touches either digit (ops on disjoint
// digits commute exactly).
inline void dmb_chain_ops(const dmb_pass& P, int n_ops, int* order, bool* chain) {
  for (int k = 0; k < n_ops; ++k) { order[k] = k; chain[k] = false; }
  if (!dmb_chain_ops_enabled()) return;
  int k = 0;
  while (k + 1 < n_ops) {
    const dmb_op& o = P.ops[order[k]];
    const int kx = dmb_op_kindx(o);
    int partner = -1;
    if (dmb_kind_is_chainable(kx)) {
      for (int t = k + 1; t < n_ops; ++t) {
        const dmb_op& x = P.ops[order[t]];
        if (x.a == o.a || x.a == o.b || x.b == o.a || x.b == o.b) {
          if (x.a == o.a && x.b == o.b && dmb_op_kindx(x) == kx) partner = t;
          break;
        }
      }
    }
    if (partner >= 0) {        // at least one of the four maps must exist: the chained bodies are the both-maps variants
      const dmb_op& x = P.ops[order[partner]];
      if (dmb_op_map_class(o, 0) + dmb_op_map_class(o, 1) + dmb_op_map_class(x, 0) + dmb_op_map_class(x, 1) == 0) partner = -1;
    }
    if (partner < 0) { ++k; continue; }
    const int moved = order[partner];
    for (int t = partner; t > k + 1; --t) order[t] = order[t - 1];
    order[k + 1] = moved;
    chain[k] = true;
    k += 2;
  }
}

inline void dmb_make_lean_pass(const dmb_pass& P, int n_bits, dmb_lean_pass& L, bool fold_swaps = false) {
  L.n_tiles = 1ull << (n_bits - 2 * DMB_LEAN_K);
  L.n_ops = P.n_ops;
  L.n_chained = 0;
  for (int j = 0; j < DMB_LEAN_K; ++j) L.td[j] = P.tile_digit[j];
  for (int i = 0; i < DMB_LEAN_PAIRS; ++i) {
    const uint32_t l = (uint32_t)i << 9;
    L.pair_goff[i] = dmb_tile_off(l, P.tile_digit, DMB_LEAN_K);
    L.pair_soff[i] = dmb_swz(l) << 3;
  }
  int dest[DMB_LEAN_K];                       // tile digit j's content ends up at tile digit dest[j]
  for (int j = 0; j < DMB_LEAN_K; ++j) dest[j] = j;
  if (fold_swaps) {
    int first = P.n_ops;
    while (first > 0 && P.ops[first - 1].kind == DMB_OP_SWAP && (P.ops[first - 1].flags & (DMB_HAS_PA | DMB_HAS_PB)) == 0)
      --first;
    for (int k = first; k < P.n_ops; ++k) {
      const int a = P.ops[k].a, b = P.ops[k].b;
      for (int j = 0; j < DMB_LEAN_K; ++j) dest[j] = dest[j] == a ? b : (dest[j] == b ? a : dest[j]);
    }
    L.n_ops = first;
  }
  bool identity = true;
  for (int j = 0; j < DMB_LEAN_K; ++j) {
    L.st_perm[dest[j]] = j;
    if (dest[j] != j) identity = false;
  }
  L.st_mode = identity ? DMB_ST_PLAIN : (L.st_perm[0] == 0 ? DMB_ST_PERM128 : DMB_ST_SPLIT64);
  L.st_delta = dmb_swz(dmb_st_source(1u, L.st_perm)) << 3;
  int ibit[3] = {9, 10, 11};
  for (int k = 0; k < 8; ++k) L.st_tbit[k] = k + 1;
  if (L.st_mode == DMB_ST_SPLIT64) dmb_choose_split_order(L.st_perm, L.st_tbit, ibit);
  for (int i = 0; i < DMB_LEAN_PAIRS; ++i) {
    uint32_t l2 = 0;
    for (int k = 0; k < 3; ++k) l2 |= (((uint32_t)i >> k) & 1u) << ibit[k];
    L.st_pair_soff[i] = dmb_swz(dmb_st_source(l2, L.st_perm)) << 3;
    L.st_pair_goff[i] = dmb_tile_off(l2, P.tile_digit, DMB_LEAN_K);
  }
  int order[DMB_MAX_OPS];
  bool chain[DMB_MAX_OPS];
  dmb_chain_ops(P, L.n_ops, order, chain);
  L.n_chained = 0;
  for (int k = 0; k < L.n_ops; ++k) {
    const dmb_op& o = P.ops[order[k]];
    dmb_lean_op& q = L.ops[k];
    for (int i = 0; i < 4; ++i) {
      q.sa[i] = dmb_swz((uint32_t)i << (2 * o.a)) << 3;
      q.sj[i] = dmb_swz((uint32_t)i << (2 * o.b)) << 3;
      q.sh[i] = (uint32_t)(2 * o.fd[i]);
    }
    q.kind = o.kind;
    q.flags = o.flags & (DMB_HAS_PA | DMB_HAS_PB);
    q.mode = o.a == 0 ? DMB_MODE_PAIR_A : (o.b == 0 ? DMB_MODE_PAIR_B : DMB_MODE_A);
    for (int i = 0; i < 12; ++i) { q.pa[i] = o.pa[i]; q.pb[i] = o.pb[i]; }
    for (int i = 0; i < 16; ++i) q.coef[i] = o.coef[i];
    int ma = dmb_op_map_class(o, 0), mb = dmb_op_map_class(o, 1);
    const bool leads = chain[k], follows = k > 0 && chain[k - 1];
    if (leads || follows) {
      // both ops of a chain run the same both-maps variant: the widest map class of the four maps, a missing map
      // becomes the identity (x * 1 + 0 * ..: exact)
      const dmb_op& other = P.ops[order[leads ? k + 1 : k - 1]];
      int c = ma > mb ? ma : mb;
      const int oa = dmb_op_map_class(other, 0), ob = dmb_op_map_class(other, 1);
      if (oa > c) c = oa;
      if (ob > c) c = ob;
      static const double ident[12] = {0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
      if (ma == 0) for (int i = 0; i < 12; ++i) q.pa[i] = ident[i];
      if (mb == 0) for (int i = 0; i < 12; ++i) q.pb[i] = ident[i];
      q.flags |= DMB_HAS_PA | DMB_HAS_PB;
      ma = mb = c;
      if (leads) { q.flags |= DMB_CHAIN; ++L.n_chained; }
    }
    if (ma == 2) q.flags |= DMB_PA_COL0;
    if (mb == 2) q.flags |= DMB_PB_COL0;
    if (o.kind == DMB_OP_CX_TSP && o.coef[1] == 0.0 && o.coef[4] == 0.0) q.flags |= DMB_TSP_ZERO_MEAN;
    const int kx = dmb_op_kindx(o);
    q.variant = dmb_variant_is_specialised(kx, ma, mb) ? dmb_variant_id(kx, ma, mb, q.mode) : -1;
    // thread digit 0 sits on tile digit 0: virtual threads 2u and 2u+1 own the two halves of every 16-byte pair
    if (q.mode == DMB_MODE_A && o.fd[0] == 0) q.flags |= DMB_PAIRABLE;
  }
}

// rows 1..3 of a map whose column 0 is zero (X,Y,Z mix among themselves only)
DMB_HD void dmb_mat3_nocol0(const double* __restrict__ m, double& x1, double& x2, double& x3) {
  const double y1 = m[1] * x1 + m[2] * x2 + m[3] * x3;
  const double y2 = m[5] * x1 + m[6] * x2 + m[7] * x3;
  const double y3 = m[9] * x1 + m[10] * x2 + m[11] * x3;
  x1 = y1; x2 = y2; x3 = y3;
}

struct dmb_lean_thread {        // per-thread constants, computed once per launch
  uint32_t tq[4];               // the thread's four base-4 index digits
  uint64_t goff;                // global element offset of pair 0 (l = 2t)
  uint32_t soff;                // swizzled byte offset of pair 0
  uint32_t st_soff;             // relabelling store: swizzled byte offset of the source of pair 0
  uint64_t st_goff;             // SPLIT64 store: global element offset of the thread part
};

DMB_HD void dmb_lean_thread_init(int t, const dmb_lean_pass& L, dmb_lean_thread& T) {
  T.tq[0] = (uint32_t)t & 3u; T.tq[1] = ((uint32_t)t >> 2) & 3u;
  T.tq[2] = ((uint32_t)t >> 4) & 3u; T.tq[3] = ((uint32_t)t >> 6) & 3u;
  T.goff = dmb_tile_off(2u * (uint32_t)t, L.td, DMB_LEAN_K);
  T.soff = dmb_swz(2u * (uint32_t)t) << 3;
  uint32_t l2 = 0;
  for (int k = 0; k < 8; ++k) l2 |= (((uint32_t)t >> k) & 1u) << L.st_tbit[k];
  T.st_soff = L.st_mode == DMB_ST_PLAIN ? T.soff : dmb_swz(dmb_st_source(l2, L.st_perm)) << 3;
  T.st_goff = dmb_tile_off(l2, L.td, DMB_LEAN_K);
}

// arithmetic of one op on the thread's 16-block (shared by all access modes)
DMB_HD void dmb_lean_math(const dmb_lean_op& op, double (&v)[4][4]) {
  if (op.flags & DMB_HAS_PA) {
    if (op.flags & DMB_PA_COL0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dmb_mat3(op.pa, v[0][j], v[1][j], v[2][j], v[3][j]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) dmb_mat3_nocol0(op.pa, v[1][j], v[2][j], v[3][j]);
    }
  }
  if (op.flags & DMB_HAS_PB) {
    if (op.flags & DMB_PB_COL0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dmb_mat3(op.pb, v[i][0], v[i][1], v[i][2], v[i][3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) dmb_mat3_nocol0(op.pb, v[i][1], v[i][2], v[i][3]);
    }
  }
  switch (op.kind) {
    case DMB_OP_CX: dmb_cx_ideal(v); break;
    case DMB_OP_CX_TSP: dmb_cx_tsp(v, op.coef[0], op.coef[1], op.coef[2], op.coef[3], op.coef[4]); break;
    case DMB_OP_DIAG2:
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) v[i][j] *= op.coef[4 * i + j];
      break;
    case DMB_OP_SWAP:
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) { const double tmp = v[i][j]; v[i][j] = v[j][i]; v[j][i] = tmp; }
      break;
    default: break;
  }
}

// Memory accessor for a stage: the CUDA kernel passes a 32-bit shared-window address and
// inline ld.shared/st.shared (one LDS/STS per access, no 64-bit generic address arithmetic);
// the CPU tests pass a plain pointer.  Mem::ld64/ld128/st64/st128 take a BYTE offset.
struct dmb_host_mem {
  unsigned char* base;
  double ld64(uint32_t off) const { return *reinterpret_cast<const double*>(base + off); }
  dmb_d2 ld128(uint32_t off) const { return *reinterpret_cast<const dmb_d2*>(base + off); }
  void st64(uint32_t off, double v) const { *reinterpret_cast<double*>(base + off) = v; }
  void st128(uint32_t off, dmb_d2 v) const { *reinterpret_cast<dmb_d2*>(base + off) = v; }
};

template <class Mem>
DMB_HD void dmb_lean_op_thread(const dmb_lean_thread& T, const dmb_lean_op& op, const Mem& mem) {
  const uint32_t bl = (T.tq[0] << op.sh[0]) | (T.tq[1] << op.sh[1]) | (T.tq[2] << op.sh[2]) | (T.tq[3] << op.sh[3]);
  const uint32_t sb = dmb_swz(bl) << 3;
  const int mode = op.mode;
  double v[4][4];
  if (mode == DMB_MODE_A) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = mem.ld64(sb ^ op.sa[i] ^ op.sj[j]);
  } else if (mode == DMB_MODE_PAIR_A) {      // digit 0 is digit a: rows i come in 16-byte pairs
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t o = sb ^ op.sj[j];
      const dmb_d2 p0 = mem.ld128(o);
      const dmb_d2 p1 = mem.ld128(o ^ 16u);
      v[0][j] = p0.x; v[1][j] = p0.y; v[2][j] = p1.x; v[3][j] = p1.y;
    }
  } else {                                   // digit 0 is digit b: columns j come in pairs
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t o = sb ^ op.sa[i];
      const dmb_d2 p0 = mem.ld128(o);
      const dmb_d2 p1 = mem.ld128(o ^ 16u);
      v[i][0] = p0.x; v[i][1] = p0.y; v[i][2] = p1.x; v[i][3] = p1.y;
    }
  }
  dmb_lean_math(op, v);
  if (mode == DMB_MODE_A) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) mem.st64(sb ^ op.sa[i] ^ op.sj[j], v[i][j]);
  } else if (mode == DMB_MODE_PAIR_A) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t o = sb ^ op.sj[j];
      dmb_d2 p0, p1;
      p0.x = v[0][j]; p0.y = v[1][j]; p1.x = v[2][j]; p1.y = v[3][j];
      mem.st128(o, p0);
      mem.st128(o ^ 16u, p1);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t o = sb ^ op.sa[i];
      dmb_d2 p0, p1;
      p0.x = v[i][0]; p0.y = v[i][1]; p1.x = v[i][2]; p1.y = v[i][3];
      mem.st128(o, p0);
      mem.st128(o ^ 16u, p1);
    }
  }
}

// arithmetic of a compile-time specialised op on one 16-block (same statements as in dmb_lean_op_spec)
template <int KINDX, int MA, int MB>
DMB_HD void dmb_spec_math(const dmb_lean_op& op, double (&v)[4][4]) {
  if constexpr (MA == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dmb_mat3_nocol0(op.pa, v[1][j], v[2][j], v[3][j]);
  } else if constexpr (MA == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dmb_mat3(op.pa, v[0][j], v[1][j], v[2][j], v[3][j]);
  }
  if constexpr (MB == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmb_mat3_nocol0(op.pb, v[i][1], v[i][2], v[i][3]);
  } else if constexpr (MB == 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmb_mat3(op.pb, v[i][0], v[i][1], v[i][2], v[i][3]);
  }
  if constexpr (KINDX == DMB_OP_CX) dmb_cx_ideal(v);
  else if constexpr (KINDX == DMB_KIND_TSP0) dmb_cx_tsp0(v, op.coef[0], op.coef[2], op.coef[3]);
  else if constexpr (KINDX == DMB_OP_CX_TSP) dmb_cx_tsp(v, op.coef[0], op.coef[1], op.coef[2], op.coef[3], op.coef[4]);
  else if constexpr (KINDX == DMB_OP_SWAP) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i + 1; j < 4; ++j) { const double tmp = v[i][j]; v[i][j] = v[j][i]; v[j][i] = tmp; }
  }
}

// PAIRED op body: one real thread plays the virtual threads 2u and 2u+1 of an op in
// access mode A (tile digit 0 free).  Their 16-blocks differ only in the low bit of digit 0, i.e. they are
// the two halves of the same sixteen 16-byte pairs: 16 LDS.128 + 16 STS.128 move both blocks (instead of
// 2 x 16 LDS.64 + 2 x 16 STS.64), with one address computation per pair.  T is the EVEN virtual thread.
// Bank conflicts: a quarter-warp of real lanes is a half-warp of virtual lanes, whose 16 8-byte slots are 8
// distinct 16-byte chunks (tests/test_host_logic.py::test_lane_order_is_bank_conflict_free).
template <int KINDX, int MA, int MB, bool CHAIN = false, class Mem>
DMB_HD void dmb_lean_op_pair(const dmb_lean_thread& T, const dmb_lean_op& op, const Mem& mem, const dmb_lean_op* op2 = nullptr) {
  const uint32_t bl = (T.tq[0] << op.sh[0]) | (T.tq[1] << op.sh[1]) | (T.tq[2] << op.sh[2]) | (T.tq[3] << op.sh[3]);
  const uint32_t sb = dmb_swz(bl) << 3;
  double v0[4][4], v1[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const dmb_d2 p = mem.ld128(sb ^ op.sa[i] ^ op.sj[j]);
      v0[i][j] = p.x; v1[i][j] = p.y;
    }
  dmb_spec_math<KINDX, MA, MB>(op, v0);
  dmb_spec_math<KINDX, MA, MB>(op, v1);
  if constexpr (CHAIN) {           // the next op: same digit pair, same variant -- the blocks stay in registers
    dmb_spec_math<KINDX, MA, MB>(*op2, v0);
    dmb_spec_math<KINDX, MA, MB>(*op2, v1);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dmb_d2 p;
      p.x = v0[i][j]; p.y = v1[i][j];
      mem.st128(sb ^ op.sa[i] ^ op.sj[j], p);
    }
}

// Compile-time specialised op body: KINDX in {MATS, CX, CX_TSP, SWAP, DMB_KIND_TSP0},
// MA/MB: 0 = no map, 1 = map without column 0, 2 = full map; MODE as in dmb_lean_op.mode.
template <int KINDX, int MA, int MB, int MODE, bool CHAIN = false, class Mem>
DMB_HD void dmb_lean_op_spec(const dmb_lean_thread& T, const dmb_lean_op& op, const Mem& mem, const dmb_lean_op* op2 = nullptr) {
  const uint32_t bl = (T.tq[0] << op.sh[0]) | (T.tq[1] << op.sh[1]) | (T.tq[2] << op.sh[2]) | (T.tq[3] << op.sh[3]);
  const uint32_t sb = dmb_swz(bl) << 3;
  double v[4][4];
  uint32_t ad[4][4];
  if constexpr (MODE == DMB_MODE_A) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { ad[i][j] = sb ^ op.sa[i] ^ op.sj[j]; v[i][j] = mem.ld64(ad[i][j]); }
  } else if constexpr (MODE == DMB_MODE_PAIR_A) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      ad[0][j] = sb ^ op.sj[j];
      const dmb_d2 p0 = mem.ld128(ad[0][j]);
      const dmb_d2 p1 = mem.ld128(ad[0][j] ^ 16u);
      v[0][j] = p0.x; v[1][j] = p0.y; v[2][j] = p1.x; v[3][j] = p1.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ad[i][0] = sb ^ op.sa[i];
      const dmb_d2 p0 = mem.ld128(ad[i][0]);
      const dmb_d2 p1 = mem.ld128(ad[i][0] ^ 16u);
      v[i][0] = p0.x; v[i][1] = p0.y; v[i][2] = p1.x; v[i][3] = p1.y;
    }
  }
  if constexpr (MA == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dmb_mat3_nocol0(op.pa, v[1][j], v[2][j], v[3][j]);
  } else if constexpr (MA == 2) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dmb_mat3(op.pa, v[0][j], v[1][j], v[2][j], v[3][j]);
  }
  if constexpr (MB == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmb_mat3_nocol0(op.pb, v[i][1], v[i][2], v[i][3]);
  } else if constexpr (MB == 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmb_mat3(op.pb, v[i][0], v[i][1], v[i][2], v[i][3]);
  }
  if constexpr (KINDX == DMB_OP_CX) dmb_cx_ideal(v);
  else if constexpr (KINDX == DMB_KIND_TSP0) dmb_cx_tsp0(v, op.coef[0], op.coef[2], op.coef[3]);
  else if constexpr (KINDX == DMB_OP_CX_TSP) dmb_cx_tsp(v, op.coef[0], op.coef[1], op.coef[2], op.coef[3], op.coef[4]);
  else if constexpr (KINDX == DMB_OP_SWAP) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i + 1; j < 4; ++j) { const double tmp = v[i][j]; v[i][j] = v[j][i]; v[j][i] = tmp; }
  }
  if constexpr (CHAIN) dmb_spec_math<KINDX, MA, MB>(*op2, v);    // the next op on the same digit pair (DMB_CHAIN)
  if constexpr (MODE == DMB_MODE_A) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) mem.st64(ad[i][j], v[i][j]);
  } else if constexpr (MODE == DMB_MODE_PAIR_A) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dmb_d2 p0, p1;
      p0.x = v[0][j]; p0.y = v[1][j]; p1.x = v[2][j]; p1.y = v[3][j];
      mem.st128(ad[0][j], p0);
      mem.st128(ad[0][j] ^ 16u, p1);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dmb_d2 p0, p1;
      p0.x = v[i][0]; p0.y = v[i][1]; p1.x = v[i][2]; p1.y = v[i][3];
      mem.st128(ad[i][0], p0);
      mem.st128(ad[i][0] ^ 16u, p1);
    }
  }
}

#define DMB_SPEC_MODES(K, A, B)                                                                        \
  case ((K * 3 + A) * 3 + B) * 3 + 0: dmb_lean_op_spec<K, A, B, 0>(T, op, mem); break;                 \
  case ((K * 3 + A) * 3 + B) * 3 + 1: dmb_lean_op_spec<K, A, B, 1>(T, op, mem); break;                 \
  case ((K * 3 + A) * 3 + B) * 3 + 2: dmb_lean_op_spec<K, A, B, 2>(T, op, mem); break;

template <class Mem>
DMB_HD void dmb_lean_op_dispatch(const dmb_lean_thread& T, const dmb_lean_op& op, const Mem& mem) {
  switch (op.variant) {
    DMB_SPEC_MODES(DMB_OP_MATS, 1, 1)
    DMB_SPEC_MODES(DMB_OP_MATS, 2, 2)
    DMB_SPEC_MODES(DMB_OP_CX, 1, 1)
    DMB_SPEC_MODES(DMB_OP_CX, 2, 2)
    DMB_SPEC_MODES(DMB_OP_CX, 0, 0)
    DMB_SPEC_MODES(DMB_OP_CX_TSP, 1, 1)
    DMB_SPEC_MODES(DMB_OP_CX_TSP, 2, 2)
    DMB_SPEC_MODES(DMB_KIND_TSP0, 1, 1)
    DMB_SPEC_MODES(DMB_KIND_TSP0, 2, 2)
    DMB_SPEC_MODES(DMB_OP_SWAP, 0, 0)
    default: dmb_lean_op_thread(T, op, mem); break;
  }
}

// virtual threads u and u + 128 one after the other (u + 128 = same index digits with tq[3] + 2); a real loop,
// not unrolled, so that the specialised bodies exist once in the instruction stream
template <class Mem>
DMB_HD void dmb_lean_op_dispatch_twice(const dmb_lean_thread& S0, const dmb_lean_op& op, const Mem& mem) {
  dmb_lean_thread T = S0;
#pragma unroll 1
  for (uint32_t h = 0; h < 2; ++h) {
    T.tq[3] = S0.tq[3] + 2u * h;
    dmb_lean_op_dispatch(T, op, mem);
  }
}

#define DMB_PAIR_CASE(K, A, B) \
  case ((K * 3 + A) * 3 + B) * 3 + 0: dmb_lean_op_pair<K, A, B>(P0, op, mem); return;

// Dispatch for a real thread u of the paired kernel.  Ops in access mode A run the paired body as virtual
// threads 2u / 2u + 1 (P0 = thread struct of 2u); every other op runs the ordinary body twice, as virtual
// threads u and u + 128 (from S0), so that consecutive lanes stay consecutive virtual threads -- the
// arrangement the bank-conflict analysis of the 128-bit modes was made for.  Either way all 256 virtual
// threads of the tile run exactly once per op.
template <class Mem>
DMB_HD void dmb_lean_op_dispatch_pair(const dmb_lean_thread& P0, const dmb_lean_thread& S0, const dmb_lean_op& op,
                                      const Mem& mem) {
  if (op.flags & DMB_PAIRABLE) {
    switch (op.variant) {
      DMB_PAIR_CASE(DMB_OP_MATS, 1, 1)
      DMB_PAIR_CASE(DMB_OP_MATS, 2, 2)
      DMB_PAIR_CASE(DMB_OP_CX, 1, 1)
      DMB_PAIR_CASE(DMB_OP_CX, 2, 2)
      DMB_PAIR_CASE(DMB_OP_CX, 0, 0)
      DMB_PAIR_CASE(DMB_OP_CX_TSP, 1, 1)
      DMB_PAIR_CASE(DMB_OP_CX_TSP, 2, 2)
      DMB_PAIR_CASE(DMB_KIND_TSP0, 1, 1)
      DMB_PAIR_CASE(DMB_KIND_TSP0, 2, 2)
      DMB_PAIR_CASE(DMB_OP_SWAP, 0, 0)
      default: break;
    }
  }
  dmb_lean_op_dispatch_twice(S0, op, mem);
}

// true when dmb_lean_op_dispatch_pair takes the paired body for this op
DMB_HD bool dmb_lean_op_is_paired(const dmb_lean_op& op) {
  return (op.flags & DMB_PAIRABLE) && op.variant >= 0 && (op.variant % 3) == 0;
}

// ---------------------------------------------------------------------------------------
// Chained ops (DMB_CHAIN): two consecutive ops on the SAME ordered digit pair and of the same specialised variant --
// the two CNOTs of a controlled-phase / controlled-rotation gate, of an rzz, ... -- keep their 16-blocks in registers:
// one shared-memory round trip and one barrier for both.  Same statements in the same order as running them one after
// the other, so the result is bit-identical.  Instantiated for the ideal CNOT and the TSP CNOTs with both maps present
// (the host pads a missing map with the identity, dmb_chain_ops).
// ---------------------------------------------------------------------------------------
#define DMB_CHAIN_VARIANTS(X) \
  X(DMB_OP_CX, 1, 1) X(DMB_OP_CX, 2, 2) X(DMB_KIND_TSP0, 1, 1) X(DMB_KIND_TSP0, 2, 2) X(DMB_OP_CX_TSP, 1, 1) X(DMB_OP_CX_TSP, 2, 2)


#define DMB_CHAIN_SPEC_CASE(K, A, B)                                                                            \
  case ((K * 3 + A) * 3 + B) * 3 + 0: dmb_lean_op_spec<K, A, B, 0, true>(T, op, mem, &op2); break;              \
  case ((K * 3 + A) * 3 + B) * 3 + 1: dmb_lean_op_spec<K, A, B, 1, true>(T, op, mem, &op2); break;              \
  case ((K * 3 + A) * 3 + B) * 3 + 2: dmb_lean_op_spec<K, A, B, 2, true>(T, op, mem, &op2); break;
#define DMB_CHAIN_PAIR_CASE(K, A, B) \
  case ((K * 3 + A) * 3 + B) * 3 + 0: dmb_lean_op_pair<K, A, B, true>(P0, op, mem, &op2); return;

// virtual threads u and u + 128 one after the other, like dmb_lean_op_dispatch_twice
template <class Mem>
DMB_HD void dmb_lean_op_chain_twice(const dmb_lean_thread& S0, const dmb_lean_op& op, const dmb_lean_op& op2, const Mem& mem) {
  dmb_lean_thread T = S0;
#pragma unroll 1
  for (uint32_t h = 0; h < 2; ++h) {
    T.tq[3] = S0.tq[3] + 2u * h;
    switch (op.variant) {
      DMB_CHAIN_VARIANTS(DMB_CHAIN_SPEC_CASE)
      default: break;       // the host chains these variants only
    }
  }
}

template <class Mem>
DMB_HD void dmb_lean_op_chain_pair(const dmb_lean_thread& P0, const dmb_lean_thread& S0, const dmb_lean_op& op,
                                   const dmb_lean_op& op2, const Mem& mem) {
  if (op.flags & DMB_PAIRABLE) {
    switch (op.variant) {
      DMB_CHAIN_VARIANTS(DMB_CHAIN_PAIR_CASE)
      default: break;
    }
  }
  dmb_lean_op_chain_twice(S0, op, op2, mem);
}

// One step of a pass's op loop for real thread u: runs ops[0] (and ops[1] with it when they are chained) and returns
// the number of ops consumed.  PAIRED: the shipped kernel's dispatch (dmb_lean_op_dispatch_pair).
// CHAINS = false: an instantiation without the chained bodies, for passes whose L.n_chained is 0 (the host picks; keeps
// the instruction footprint of the common kernel what it was).
template <bool PAIRED, bool CHAINS = true, class Mem>
DMB_HD int dmb_lean_ops_step(const dmb_lean_thread& P0, const dmb_lean_thread& S0, const dmb_lean_op* ops, const Mem& mem) {
  if constexpr (CHAINS) {
    if (ops[0].flags & DMB_CHAIN) {
      if (PAIRED) dmb_lean_op_chain_pair(P0, S0, ops[0], ops[1], mem);
      else dmb_lean_op_chain_twice(S0, ops[0], ops[1], mem);
      return 2;
    }
  }
  if (PAIRED) dmb_lean_op_dispatch_pair(P0, S0, ops[0], mem);
  else dmb_lean_op_dispatch_twice(S0, ops[0], mem);
  return 1;
}

// load / store of one tile (the CUDA kernel replaces the load by cp.async.cg of the same
// addresses; the store is used as is)
template <class Mem>
DMB_HD void dmb_lean_load_thread(const dmb_lean_thread& T, const dmb_lean_pass& L, const double* state,
                                 uint64_t tile_base, const dmb_remote_src& S, const Mem& mem) {
#pragma unroll
  for (int i = 0; i < DMB_LEAN_PAIRS; ++i)
    mem.st128(T.soff ^ L.pair_soff[i],
              *reinterpret_cast<const dmb_d2*>(dmb_src_ptr(S, state, tile_base + (T.goff | L.pair_goff[i]))));
}

template <bool PUSH, int STMODE, class Mem>
DMB_HD void dmb_lean_store_thread(const dmb_lean_thread& T, const dmb_lean_pass& L, double* state,
                                  uint64_t tile_base, const dmb_remote_src& D, const Mem& mem) {
  dmb_d2 w[DMB_LEAN_PAIRS];
#pragma unroll
  for (int i = 0; i < DMB_LEAN_PAIRS; ++i) {
    if constexpr (STMODE == DMB_ST_PLAIN) {
      w[i] = mem.ld128(T.soff ^ L.pair_soff[i]);
    } else if constexpr (STMODE == DMB_ST_PERM128) {
      w[i] = mem.ld128(T.st_soff ^ L.st_pair_soff[i]);
    } else {
      const uint32_t a = T.st_soff ^ L.st_pair_soff[i];
      w[i].x = mem.ld64(a);
      w[i].y = mem.ld64(a ^ L.st_delta);
    }
  }
#pragma unroll
  for (int i = 0; i < DMB_LEAN_PAIRS; ++i) {
    const uint64_t idx = tile_base + (STMODE == DMB_ST_SPLIT64 ? (T.st_goff | L.st_pair_goff[i]) : (T.goff | L.pair_goff[i]));
    double* dst = PUSH ? reinterpret_cast<double*>(D.tab[dmb_remote_index(D, idx)]) + idx : state + idx;
    *reinterpret_cast<dmb_d2*>(dst) = w[i];
  }
}

// ---------------------------------------------------------------------------------------
// Body of the tile kernel (k_tile_pass6 in dmb200.cu), written against an execution-context policy so that the
// SAME control flow -- tile loop, stage ring, asynchronous staging, barriers -- runs on the GPU and, with real host
// threads, in the CPU tests (tests/emu).
//   Ctx: tid() / block() / grid()       thread index in the CTA (0..127), CTA index, number of CTAs
//        copy16(stage byte offset, src) asynchronous 16-byte copy global -> stage memory (cp.async)
//        commit() / wait<N>()           close the current copy group / wait for all but the N newest groups
//        sync()                         CTA barrier
//        mem(stage byte offset)         accessor (ld64 / ld128 / st64 / st128) of a stage
// 128 threads per 4096-coefficient tile.  Staging (tile load / write-back) and ops that touch tile digit 0 walk the
// tile as virtual threads u and u + 128; ops that leave tile digit 0 free (PAIRED) run as virtual threads 2u / 2u + 1,
// whose two 16-blocks are the halves of the same sixteen 16-byte pairs (dmb_lean_op_pair).
// STAGES = 1: load, wait, ops, store per tile (the other CTAs of the SM hide the latency; what ships: 5 CTAs per
// SM); STAGES = 2: the next tile of this CTA streams in during the op phase (measured slower on B200, kept for the
// CPU control-flow test of the ring).
// REMOTE (fused exchange, distributed.py): 0 in place; 1 "pull": pair idx of the tile is loaded from the peer buffer
// R.tab[dmb_remote_index(R, idx)] that holds it in the old layout; 2 "push": the write-back goes to the peer that owns idx.
// ---------------------------------------------------------------------------------------
#define DMB_HALF_THREADS 128
template <int STMODE, bool PAIRED, int STAGES, int REMOTE, bool CHAINS = true, class Ctx>
DMB_HD void dmb_half_kernel_body(Ctx& cx, double* state, const dmb_lean_pass& L, const dmb_remote_src& R) {
  dmb_lean_thread S0, S1;
  dmb_lean_thread_init(cx.tid(), L, S0);
  dmb_lean_thread_init(cx.tid() + DMB_HALF_THREADS, L, S1);
  dmb_lean_thread P0;                                   // PAIRED: virtual thread 2u (only its index digits are used)
  dmb_lean_thread_init(2 * cx.tid(), L, P0);
  const uint64_t first = cx.block(), stride = cx.grid();
  if (first >= L.n_tiles) return;
  auto src_of = [&](uint64_t idx) -> const double* {
    if (REMOTE == 1) return reinterpret_cast<const double*>(R.tab[dmb_remote_index(R, idx)]) + idx;
    return state + idx;
  };
  if (STAGES == 2) {
    const uint64_t tb = dmb_tile_base(first, L.td, DMB_LEAN_K);
#pragma unroll
    for (int i = 0; i < DMB_LEAN_PAIRS; ++i) {
      cx.copy16(S0.soff ^ L.pair_soff[i], src_of(tb + (S0.goff | L.pair_goff[i])));
      cx.copy16(S1.soff ^ L.pair_soff[i], src_of(tb + (S1.goff | L.pair_goff[i])));
    }
    cx.commit();
  }
  uint32_t cur = 0;
  for (uint64_t tile = first; tile < L.n_tiles; tile += stride) {
    const uint64_t fetch_tile = STAGES == 2 ? tile + stride : tile;
    if (fetch_tile < L.n_tiles) {
      const uint64_t tb = dmb_tile_base(fetch_tile, L.td, DMB_LEAN_K);
      const uint32_t dst = (STAGES == 2 ? (cur ^ 1u) : 0u) * DMB_LEAN_TILE_BYTES;
#pragma unroll
      for (int i = 0; i < DMB_LEAN_PAIRS; ++i) {
        cx.copy16(dst + (S0.soff ^ L.pair_soff[i]), src_of(tb + (S0.goff | L.pair_goff[i])));
        cx.copy16(dst + (S1.soff ^ L.pair_soff[i]), src_of(tb + (S1.goff | L.pair_goff[i])));
      }
    }
    cx.commit();
    cx.template wait<STAGES - 1>();
    cx.sync();
    const auto mem = cx.mem(cur * DMB_LEAN_TILE_BYTES);
    for (int i = 0; i < L.n_ops;) {
      i += dmb_lean_ops_step<PAIRED, CHAINS>(P0, S0, &L.ops[i], mem);
      cx.sync();
    }
    const uint64_t tb = dmb_tile_base(tile, L.td, DMB_LEAN_K);
    dmb_lean_store_thread<REMOTE == 2, STMODE>(S0, L, state, tb, R, mem);
    dmb_lean_store_thread<REMOTE == 2, STMODE>(S1, L, state, tb, R, mem);
    cx.sync();
    if (STAGES == 2) cur ^= 1u;
  }
}

// ---------------------------------------------------------------------------------------
// Element-wise kernels (one logical thread per output element)
// ---------------------------------------------------------------------------------------
struct dmb_qubit_map {              // where each qubit's digit lives in the global index
  int32_t n_qubits;
  int32_t hi[DMB_MAX_QUBITS];
  int32_t lo[DMB_MAX_QUBITS];
};

DMB_HD int dmb_digit_of(uint64_t g, int hi, int lo) {
  return (int)(((g >> hi) & 1ull) << 1 | ((g >> lo) & 1ull));
}

struct dmb_init_params {
  dmb_qubit_map map;
  double v[DMB_MAX_QUBITS][4];
  double scale;
  uint64_t rank_bits;
  int32_t n_bits;
};

DMB_HD double dmb_init_value(uint64_t idx, const dmb_init_params& p) {
  const uint64_t g = (p.rank_bits << p.n_bits) | idx;
  double x = p.scale;
  for (int q = 0; q < p.map.n_qubits; ++q) x *= p.v[q][dmb_digit_of(g, p.map.hi[q], p.map.lo[q])];
  return x;
}

// Marginal gather (see dmb_marginal in dmb200.h).  Qubits are split by the host into
// "simple" ones (at most one non-zero weight per result bit value: the usual I / B pick)
// and up to DMB_MAX_MULTI "multi" ones whose weights mix several digit values (pending
// single-qubit maps: on sharded qubits, and the few a readout-only job leaves on local ones).
#define DMB_MAX_MULTI 4
struct dmb_marginal_params {
  dmb_qubit_map map;
  double wt[DMB_MAX_QUBITS][2][4];
  int8_t simple_digit[DMB_MAX_QUBITS][2];   // digit picked for result bit 0 / 1, -1 = multi
  int32_t n_multi;
  int32_t multi[DMB_MAX_MULTI];
  uint64_t rank_bits;
  int32_t n_bits;
};

DMB_HD uint64_t dmb_place_digit(int d, int hi, int lo) {
  return ((uint64_t)((d >> 1) & 1) << hi) | ((uint64_t)(d & 1) << lo);
}

DMB_HD double dmb_marginal_value(uint64_t c, const double* __restrict__ state,
                                 const dmb_marginal_params& p) {
  uint64_t g0 = 0;
  double w0 = 1.0;
  for (int k = 0; k < p.map.n_qubits; ++k) {
    const int cb = (int)((c >> k) & 1ull);
    const int d = p.simple_digit[k][cb];
    if (d < 0) continue;
    w0 *= p.wt[k][cb][d];
    g0 |= dmb_place_digit(d, p.map.hi[k], p.map.lo[k]);
  }
  if (w0 == 0.0) return 0.0;
  const uint64_t mask = (1ull << p.n_bits) - 1ull;
  double total = 0.0;
  const int combos = 1 << (2 * p.n_multi);
  for (int combo = 0; combo < combos; ++combo) {
    uint64_t g = g0;
    double w = w0;
    for (int m = 0; m < p.n_multi; ++m) {
      const int k = p.multi[m];
      const int cb = (int)((c >> k) & 1ull);
      const int d = (combo >> (2 * m)) & 3;
      w *= p.wt[k][cb][d];
      g |= dmb_place_digit(d, p.map.hi[k], p.map.lo[k]);
    }
    if (w == 0.0) continue;
    if ((g >> p.n_bits) != p.rank_bits) continue;
    total += w * state[g & mask];
  }
  return total;
}

// One butterfly of the Walsh-Hadamard transform: pair number `t` of stage `stage`.
DMB_HD void dmb_fwht_pair(uint64_t t, int stage, double* vec) {
  const uint64_t low = t & ((1ull << stage) - 1ull);
  const uint64_t i = ((t >> stage) << (stage + 1)) | low;
  const uint64_t j = i | (1ull << stage);
  const double a = vec[i], b = vec[j];
  vec[i] = a + b;
  vec[j] = a - b;
}

// 'N'-basis contraction step: in [H][4][L] -> out [H][2][L]; t enumerates (h, l).
DMB_HD void dmb_contract_elem(uint64_t t, const double* __restrict__ in, double* __restrict__ out,
                              uint64_t L, double n0, double n1, double n2) {
  const uint64_t h = t / L, l = t - h * L;
  const double* src = in + h * 4 * L + l;
  out[h * 2 * L + l] = src[0];
  out[h * 2 * L + L + l] = n0 * src[L] + n1 * src[2 * L] + n2 * src[3 * L];
}

// Pauli -> matrix basis, one digit: [I,X,Y,Z] -> [I+Z, X-iY, X+iY, I-Z]
// (dm_simulator.py:1223-1245).  t enumerates the 4-vectors of digit position `pos`.
// `first` != 0 reads the real state, otherwise the complex work buffer in place.
DMB_HD void dmb_tomatrix_digit(uint64_t t, int pos, int first, const double* __restrict__ state,
                               dmb_d2* work) {
  const int sh = 2 * pos;
  const uint64_t low = t & ((1ull << sh) - 1ull);
  const uint64_t base = ((t >> sh) << (sh + 2)) | low;
  const uint64_t st = 1ull << sh;
  dmb_d2 a[4];
  for (int d = 0; d < 4; ++d) {
    if (first) { a[d].x = state[base + d * st]; a[d].y = 0.0; }
    else a[d] = work[base + d * st];
  }
  dmb_d2 r0, r1, r2, r3;
  r0.x = a[0].x + a[3].x; r0.y = a[0].y + a[3].y;
  r1.x = a[1].x + a[2].y; r1.y = a[1].y - a[2].x;     // X - iY
  r2.x = a[1].x - a[2].y; r2.y = a[1].y + a[2].x;     // X + iY
  r3.x = a[0].x - a[3].x; r3.y = a[0].y - a[3].y;
  work[base] = r0; work[base + st] = r1; work[base + 2 * st] = r2; work[base + 3 * st] = r3;
}

// Final scatter: out[row][col] = work[interleave(row, col)], digit mu = 2*rowbit + colbit
// (dm_simulator.py:1229-1253).  t = row * 2^n + col.
DMB_HD void dmb_tomatrix_scatter(uint64_t t, int n, const dmb_d2* __restrict__ work, dmb_d2* out) {
  const uint64_t row = t >> n, col = t & ((1ull << n) - 1ull);
  uint64_t idx = 0;
  for (int b = 0; b < n; ++b)
    idx |= (((row >> b) & 1ull) << (2 * b + 1)) | (((col >> b) & 1ull) << (2 * b));
  out[t] = work[idx];
}
