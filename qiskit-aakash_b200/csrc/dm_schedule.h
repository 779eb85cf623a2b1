// dm_schedule.h -- host-side gate-fusion scheduler of libdmb200 (plain C++, no CUDA).
//
// Packs a stream of two-qubit ops (every single-qubit map already rides on one of them, see
// qiskit-aakash_b200/engine.py) into tile passes for dmb_apply_passes: greedy list scheduling
// with qubit blocking plus dynamic relabelling of digit positions 0/1.  This generalises the
// reference's only fusion, the per-qubit U3 merge (qiskit/providers/basicaer/basicaertools.py:
// 251-307), and is the native form of schedule.build_passes_relabel (the Python function stays
// as the executable specification: tests/test_native_schedule.py checks that both produce
// byte-identical dmb_pass arrays).  Qubit sets are 64-bit masks, so one greedy selection over a
// 256-op window costs a few hundred nanoseconds; the whole of BASELINE config 3 (1300 CNOTs,
// n = 14) schedules in well under a millisecond.
//
// Included by dmb200.cu and by the CPU emulation library of the tests (tests/emu/emu_api.cpp);
// both wrap dmb_schedule_impl() in the extern "C" entry point dmb_schedule (include/dmb200.h).
#ifndef DM_SCHEDULE_H
#define DM_SCHEDULE_H

#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "dmb200.h"

namespace dmb_sched {

typedef uint64_t qmask;

static inline int popc(qmask m) { return __builtin_popcountll(m); }

static inline qmask op_mask(const dmb_qop& op) {
  qmask m = (qmask)1 << op.qa;
  if (op.qb >= 0) m |= (qmask)1 << op.qb;
  return m;
}

// One round of list scheduling over qubit ids (schedule._greedy_select): walks `rem` in program
// order, admits an op when its qubits fit a tile of `cap` qubits and none of them is blocked by an
// earlier skipped op (ops on disjoint qubits commute, so hoisting is exact).
static inline int greedy_select(const dmb_qop* ops, const std::vector<int32_t>& rem, qmask start, int cap,
                                int max_ops, int window, int n_qubits, qmask* tile_out,
                                std::vector<int32_t>* chosen) {
  qmask tile = start, blocked = 0;
  int n_chosen = 0;
  const size_t lim = rem.size() < (size_t)window ? rem.size() : (size_t)window;
  for (size_t idx = 0; idx < lim; ++idx) {
    if (n_chosen >= max_ops || popc(blocked) >= n_qubits) break;
    const qmask dg = op_mask(ops[rem[idx]]);
    if (dg & blocked) { blocked |= dg; continue; }
    const qmask need = tile | dg;
    if (popc(need) <= cap) {
      tile = need;
      if (chosen) chosen->push_back((int32_t)idx);
      ++n_chosen;
    } else {
      blocked |= dg;
    }
  }
  if (tile_out) *tile_out = tile;
  return n_chosen;
}

// Ops of `rem` that run when the tile is exactly the qubit set S: an op inside S runs unless one of
// its qubits is blocked; every op that does not run blocks its qubits for the rest of the scan.
static inline int count_fixed(const dmb_qop* ops, const std::vector<int32_t>& rem, qmask S, int max_ops, int window,
                              int n_qubits, std::vector<int32_t>* chosen) {
  qmask blocked = 0;
  int n_chosen = 0;
  const size_t lim = rem.size() < (size_t)window ? rem.size() : (size_t)window;
  for (size_t idx = 0; idx < lim; ++idx) {
    if (n_chosen >= max_ops || popc(blocked) >= n_qubits) break;
    const qmask dg = op_mask(ops[rem[idx]]);
    if ((dg & blocked) || (dg & ~S)) { blocked |= dg; continue; }
    if (chosen) chosen->push_back((int32_t)idx);
    ++n_chosen;
  }
  return n_chosen;
}

// Strategy 1, "tile search": instead of letting program order decide which qubits join the tile
// (greedy_select), grow the tile from `start` by the qubits of one ready op at a time, always taking
// the op whose inclusion lets the most ops run (ties: program order).  On nearest-neighbour circuits
// this finds the time-skewed windows that a layer-by-layer scan misses: 130 instead of 201 passes
// for BASELINE configs[2] (n = 14, depth 200).
static inline int select_tile(const dmb_qop* ops, const std::vector<int32_t>& rem, qmask start, int cap, int max_ops,
                              int window, int n_qubits, qmask* tile_out, std::vector<int32_t>* chosen) {
  qmask S = start;
  const size_t lim = rem.size() < (size_t)window ? rem.size() : (size_t)window;
  while (popc(S) < cap) {
    qmask blocked = 0, best_S = 0;
    int best_cnt = -1;
    for (size_t idx = 0; idx < lim; ++idx) {
      if (popc(blocked) >= n_qubits) break;
      const qmask dg = op_mask(ops[rem[idx]]);
      if (dg & blocked) { blocked |= dg; continue; }
      if (dg & ~S) {
        blocked |= dg;
        if (popc(S | dg) <= cap) {
          const int cnt = count_fixed(ops, rem, S | dg, max_ops, window, n_qubits, nullptr);
          if (cnt > best_cnt) { best_cnt = cnt; best_S = S | dg; }
        }
      }
    }
    if (best_cnt < 0) break;
    S = best_S;
  }
  if (tile_out) *tile_out = S;
  return count_fixed(ops, rem, S, max_ops, window, n_qubits, chosen);
}

// schedule.lane_order: order in which thread-index bit pairs are dealt to the free tile digits
// (conflict-free shared memory under dmb_swz).
static inline void lane_order(int K, int a, int b, int8_t* fd) {
  int free_[DMB_MAX_TILE_DIGITS], nf = 0;
  for (int d = 0; d < K; ++d)
    if (d != a && d != b) free_[nf++] = d;
  if (!nf) return;
  auto has = [&](int d, int skip) {
    for (int i = 0; i < nf; ++i)
      if (free_[i] == d && d != skip) return true;
    return false;
  };
  int first[2], n_first = 0;
  if (has(0, -1)) {
    first[n_first++] = 0;
  } else {
    const int o1 = has(1, -1) ? 1 : (has(2, -1) ? 2 : free_[0]);
    first[n_first++] = o1;
    int o2 = -1;
    if (has(3, o1)) o2 = 3;
    else if (has(5, o1)) o2 = 5;
    else
      for (int i = 0; i < nf; ++i)
        if (free_[i] != o1) { o2 = free_[i]; break; }
    if (o2 >= 0) first[n_first++] = o2;
  }
  int k = 0;
  for (int i = 0; i < n_first; ++i) fd[k++] = (int8_t)first[i];
  for (int i = 0; i < nf; ++i) {
    bool in_first = false;
    for (int j = 0; j < n_first; ++j) in_first |= (first[j] == free_[i]);
    if (!in_first) fd[k++] = (int8_t)free_[i];
  }
}

struct PlannedOp {
  int32_t src;      // index into the op stream, or -1 for a SWAP emitted by the scheduler
  int32_t da, db;   // digit positions (db = -1: lone single-digit op)
};

static inline void encode_pass(const dmb_qop* ops, qmask tile_d, const std::vector<PlannedOp>& plan, dmb_pass* out) {
  memset(out, 0, sizeof(dmb_pass));
  int local[64];
  int K = 0;
  for (int d = 0; d < 64; ++d)
    if ((tile_d >> d) & 1) { local[d] = K; out->tile_digit[K++] = d; }
  out->n_tile_digits = K;
  out->n_ops = (int32_t)plan.size();
  for (size_t oi = 0; oi < plan.size(); ++oi) {
    const PlannedOp& p = plan[oi];
    dmb_op& o = out->ops[oi];
    const int a = local[p.da];
    const int b = p.db < 0 ? (a == 0 ? 1 : 0) : local[p.db];
    o.a = (int8_t)a;
    o.b = (int8_t)b;
    if (p.src < 0) {
      o.kind = DMB_OP_SWAP;
    } else {
      const dmb_qop& q = ops[p.src];
      o.kind = q.kind;
      o.flags = q.flags & (DMB_HAS_PA | DMB_HAS_PB);
      if (q.flags & DMB_HAS_PA) memcpy(o.pa, q.pa, sizeof(o.pa));
      if (q.flags & DMB_HAS_PB) memcpy(o.pb, q.pb, sizeof(o.pb));
      memcpy(o.coef, q.coef, sizeof(o.coef));
    }
    lane_order(K, a, b, o.fd);
  }
}

static inline int real_cap_of(int K, int max_ops) {          // room for two relabelling swaps
  return K >= 4 ? (max_ops - 2 > 1 ? max_ops - 2 : 1) : max_ops;
}

// See dmb_schedule in include/dmb200.h for the contract.
static inline int schedule_impl(const dmb_qop* ops, size_t n_ops, int32_t* pos, int n_qubits, int n_digits,
                                int max_tile, int max_ops, int window, int strategy, size_t min_tail, int32_t* moves,
                                int32_t* n_moves, dmb_pass* out, size_t out_cap, size_t* n_out, int32_t* left,
                                size_t* n_left, std::string& err) {
  if (n_qubits < 1 || n_qubits > 64 || n_digits < 2 || n_digits > 64) { err = "qubit / digit count out of range"; return 1; }
  if (max_tile > DMB_MAX_TILE_DIGITS) max_tile = DMB_MAX_TILE_DIGITS;
  const int K = max_tile < n_digits ? max_tile : n_digits;
  if (K < 2) { err = "state must have at least 2 digit positions"; return 1; }
  if (max_ops > DMB_MAX_OPS) max_ops = DMB_MAX_OPS;
  if (max_ops < 1 || window < 1) { err = "max_ops / window must be positive"; return 1; }
  if (strategy != DMB_SCHED_PROGRAM_ORDER && strategy != DMB_SCHED_TILE_SEARCH) { err = "unknown strategy"; return 1; }
  auto select = [&](const std::vector<int32_t>& r, qmask start, qmask* tile, std::vector<int32_t>* ch) {
    return strategy == DMB_SCHED_TILE_SEARCH ? select_tile(ops, r, start, K, real_cap_of(K, max_ops), window, n_qubits, tile, ch)
                                             : greedy_select(ops, r, start, K, real_cap_of(K, max_ops), window, n_qubits, tile, ch);
  };
  const int n_final = n_moves ? *n_moves : -1;                                   // -1: no final moves requested
  int32_t owner[64];
  for (int d = 0; d < 64; ++d) owner[d] = -1;
  for (int q = 0; q < n_qubits; ++q) {
    if (pos[q] < 0 || pos[q] >= 64 || owner[pos[q]] >= 0) { err = "pos is not an injective map into digit positions"; return 1; }
    owner[pos[q]] = q;
  }
  for (size_t i = 0; i < n_ops; ++i) {
    const dmb_qop& op = ops[i];
    if (op.qa < 0 || op.qa >= n_qubits || op.qb >= n_qubits || op.qb == op.qa || op.qb < -1) { err = "op qubits invalid"; return 1; }
    if (op.kind < DMB_OP_MATS || op.kind > DMB_OP_SWAP) { err = "unknown op kind"; return 1; }
    if (pos[op.qa] >= n_digits || (op.qb >= 0 && pos[op.qb] >= n_digits)) { err = "op on a qubit outside the local digits"; return 1; }
  }
  std::vector<int32_t> rem(n_ops), chosen, scratch_rem;
  for (size_t i = 0; i < n_ops; ++i) rem[i] = (int32_t)i;
  std::vector<PlannedOp> plan;
  size_t produced = 0;

  auto lows = [&](int32_t* cur) {
    int n = 0;
    for (int d = 0; d < 2; ++d)
      if (d < n_digits && owner[d] >= 0) cur[n++] = owner[d];
    return n;
  };
  auto mask_of = [](const int32_t* q, int n) {
    qmask m = 0;
    for (int i = 0; i < n; ++i) m |= (qmask)1 << q[i];
    return m;
  };
  auto emit_swap = [&](int q, int target) {      // move qubit q to digit `target`, displacing its owner
    const int src = pos[q], other = owner[target];
    owner[target] = q;
    pos[q] = target;
    owner[src] = other;
    if (other >= 0) pos[other] = src;
  };

  while (rem.size() > min_tail) {
    int32_t cur[2];
    int n_cur = lows(cur);
    qmask tile_q = 0;
    chosen.clear();
    select(rem, mask_of(cur, n_cur), &tile_q, &chosen);
    if (chosen.empty()) { err = "scheduler made no progress"; return 1; }
    plan.clear();
    for (int32_t idx : chosen) {
      const dmb_qop& op = ops[rem[idx]];
      plan.push_back(PlannedOp{rem[idx], pos[op.qa], op.qb < 0 ? -1 : pos[op.qb]});
    }
    {   // remaining = everything not chosen, order kept
      scratch_rem.clear();
      size_t c = 0;
      for (size_t i = 0; i < rem.size(); ++i) {
        if (c < chosen.size() && (size_t)chosen[c] == i) { ++c; continue; }
        scratch_rem.push_back(rem[i]);
      }
      rem.swap(scratch_rem);
    }
    qmask tile_d = 3;
    for (int q = 0; q < n_qubits; ++q)
      if ((tile_q >> q) & 1) tile_d |= (qmask)1 << pos[q];
    int32_t moves_here[64][2];
    int n_moves_here = 0;
    if (n_final > 0 && rem.empty()) {
      for (int i = 0; i < n_final; ++i) {
        const int q = moves[2 * i], target = moves[2 * i + 1];
        if (q < 0 || q >= n_qubits || target < 0 || target >= n_digits) { err = "final move invalid"; return 1; }
        if (pos[q] == target) continue;
        const qmask with = tile_d | ((qmask)1 << pos[q]) | ((qmask)1 << target);
        if (popc(with) <= K) {
          tile_d = with;
          moves_here[n_moves_here][0] = q;
          moves_here[n_moves_here][1] = target;
          ++n_moves_here;
        }
      }
    }
    for (int d = 0; popc(tile_d) < K; ++d) tile_d |= (qmask)1 << d;    // spare slots: longer contiguous runs
    // Look ahead: which two qubits of this tile should sit in digit positions 0/1 for the next pass?
    // Every pair (and "leave as is") is tried; the one whose greedy next pass runs the most ops wins,
    // ties prefer fewer swaps.  The swaps ride at the end of this pass (folded into its write-back).
    if (!rem.empty() && K >= 4) {
      int32_t cands[64];
      int n_cands = 0;
      qmask active = 0;
      const size_t lim = rem.size() < (size_t)window ? rem.size() : (size_t)window;
      for (size_t i = 0; i < lim; ++i) active |= op_mask(ops[rem[i]]);
      qmask in_tile = 0;
      for (int d = 0; d < 64; ++d)
        if (((tile_d >> d) & 1) && owner[d] >= 0) in_tile |= (qmask)1 << owner[d];
      for (int q = 0; q < n_qubits; ++q)
        if (((in_tile & active) >> q) & 1) cands[n_cands++] = q;
      n_cur = lows(cur);
      const qmask cur_mask = mask_of(cur, n_cur);
      int best_cnt = select(rem, cur_mask, nullptr, nullptr);
      int best_swaps = 0, best_i = -1, best_j = -1;
      for (int i = 0; i < n_cands; ++i)
        for (int j = i + 1; j < n_cands; ++j) {
          const qmask pair = ((qmask)1 << cands[i]) | ((qmask)1 << cands[j]);
          const int swaps = popc(pair & ~cur_mask);
          const int cnt = select(rem, pair, nullptr, nullptr);
          if (cnt > best_cnt || (cnt == best_cnt && swaps < best_swaps)) {
            best_cnt = cnt; best_swaps = swaps; best_i = i; best_j = j;
          }
        }
      if (best_i >= 0) {
        const int32_t targets[2] = {cands[best_i], cands[best_j]};
        const qmask tmask = mask_of(targets, 2);
        int free_low[2], n_free = 0;
        for (int d = 0; d < 2; ++d) {
          const bool kept = owner[d] >= 0 && ((cur_mask >> owner[d]) & 1) && ((tmask >> owner[d]) & 1);
          if (!kept) free_low[n_free++] = d;
        }
        int used = 0;
        for (int t = 0; t < 2; ++t) {
          const int q = targets[t];
          if ((cur_mask >> q) & 1) continue;
          if (used >= n_free) break;
          const int dd = free_low[used++];
          plan.push_back(PlannedOp{-1, dd, pos[q]});
          emit_swap(q, dd);
        }
      }
    }
    for (int i = 0; i < n_moves_here; ++i) {       // trailing swaps of the last pass
      const int q = moves_here[i][0], target = moves_here[i][1];
      if (pos[q] == target) continue;
      plan.push_back(PlannedOp{-1, pos[q], target});
      emit_swap(q, target);
    }
    if (plan.size() > DMB_MAX_OPS) { err = "pass exceeds DMB_MAX_OPS"; return 1; }
    if (produced >= out_cap) { err = "output buffer too small"; return 1; }
    encode_pass(ops, tile_d, plan, &out[produced++]);
  }
  *n_out = produced;
  if (n_left) {
    *n_left = rem.size();
    if (left)
      for (size_t i = 0; i < rem.size(); ++i) left[i] = rem[i];
  }
  if (n_final >= 0) {                              // report the moves that did not fit the last tile
    int k = 0;
    for (int i = 0; i < n_final; ++i) {
      const int q = moves[2 * i], target = moves[2 * i + 1];
      if (pos[q] != target) { moves[2 * k] = q; moves[2 * k + 1] = target; ++k; }
    }
    *n_moves = k;
  }
  return 0;
}

}  // namespace dmb_sched

#endif  // DM_SCHEDULE_H
