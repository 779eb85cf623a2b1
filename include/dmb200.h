/*
 * dmb200.h -- C ABI of libdmb200.so, the B200 (sm_100a) implementation of the
 * qiskit-aakash dm_simulator hot path (Pauli-basis density-matrix simulator).
 *
 * Everything numeric in the reference mutates one array, DmSimulatorPy._densitymatrix
 * (a real float64 vector of 4^n Pauli coefficients, qubit 0 = most significant base-4
 * digit; reference: qiskit/providers/basicaer/dm_simulator.py:61-68,376-377).  The entry
 * points below are what a binding for that path needs; each cites the reference method(s)
 * it replaces.  Conventions:
 *   - plain C, no torch / C++ types; every function returns 0 on success, non-zero on
 *     error, with the message available from dmb_last_error() (thread-local);
 *   - the CALLER owns all state / scratch buffers (device pointers; PyTorch tensors in
 *     the Python host) -- the library owns only a small per-context scratch area;
 *   - all work is enqueued on the context's CUDA stream (dmb_set_stream); only the
 *     functions documented as synchronous wait for it;
 *   - "digit position" p means the base-4 digit with stride 4^p in the *local* buffer
 *     (bits 2p, 2p+1 of the element index).  In the reference layout qubit q of an
 *     n-qubit register sits at digit position n-1-q.
 */
#ifndef DMB200_H
#define DMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMB_ABI_VERSION 5   /* 2: dmb_stats.folded_swaps; 3: dmb_schedule; 4: dmb_op.post_swap and
                               dmb_stats.r3_phases removed (measured slower, round 2), dmb_ipc_close;
                               5: dmb_download_async, dmb_stats.chained_ops,
                               dmb_apply_pass_remote_sel */
#define DMB_MAX_TILE_DIGITS 6   /* a tile holds 4^6 = 4096 doubles = 32 KiB of shared memory */
#define DMB_MAX_OPS 16          /* fused ops per tile pass */
#define DMB_MAX_QUBITS 32

typedef struct dmb_ctx dmb_ctx;

/* ---- fused two-digit op, applied inside a tile pass --------------------------------
 * Acts on the 16-element blocks spanned by tile-local digits a and b (v[i][j], i = value
 * of digit a, j = value of digit b).  Order of application inside one op:
 *   1. if (flags & DMB_HAS_PA): rows 1..3 of a 4x4 real matrix along digit a
 *      (row 0 of every trace-preserving single-qubit map is (1,0,0,0) and is implied);
 *      this is where u1/u3 rotations (basicaertools.py:69-126), memory noise
 *      (dm_simulator.py:397-425), projective measurements (:574-704), Pauli-string
 *      projections (:553-566) and reset (:810-823) end up, pre-multiplied on the host;
 *   2. if (flags & DMB_HAS_PB): the same along digit b;
 *   3. the two-digit kernel selected by `kind`. */
enum {
  DMB_OP_MATS   = 0, /* nothing further (single-qubit work only)                         */
  DMB_OP_CX     = 1, /* ideal CNOT, a = control digit, b = target digit: signed permutation
                        of the 16 Pauli pairs (basicaertools.py:310-392 with [1,0])       */
  DMB_OP_CX_TSP = 2, /* CNOT under the transition-selective-pulse error model; coef[0..4] =
                        c, s, c2, s2, cs of basicaertools.py:329-336                      */
  DMB_OP_DIAG2  = 3, /* v[i][j] *= coef[4*i+j]  (Bell-basis measurement mask,
                        dm_simulator.py:749-756)                                          */
  DMB_OP_SWAP   = 4  /* v[i][j] <-> v[j][i]: exchanges the two digits (layout remap)      */
};
enum { DMB_HAS_PA = 1, DMB_HAS_PB = 2 };

typedef struct dmb_op {
  int32_t kind;
  int32_t flags;
  int8_t a, b;      /* tile-local digit indices (indices into dmb_pass.tile_digit), a != b */
  int8_t fd[4];     /* the other tile-local digits, in the order thread-index bit pairs are
                       dealt to them (chosen by the host for conflict-free shared memory)  */
  int8_t reserved_[2]; /* must be 0 */
  double pa[12];    /* rows 1..3 of the matrix on digit a, row-major [3][4]                */
  double pb[12];    /* rows 1..3 of the matrix on digit b                                  */
  double coef[16];
} dmb_op;           /* 336 bytes */

/* One tile pass = one HBM round trip of the whole (local) state: every tile (all values
 * of the n_tile_digits listed digits, for one value of all other index bits) is staged in
 * shared memory, the ops are applied in order, and the tile is written back. */
typedef struct dmb_pass {
  int32_t n_tile_digits;                    /* K, 2..DMB_MAX_TILE_DIGITS                   */
  int32_t n_ops;                            /* 1..DMB_MAX_OPS                              */
  int32_t tile_digit[DMB_MAX_TILE_DIGITS];  /* ascending digit positions, tile_digit[0]==0 */
  dmb_op ops[DMB_MAX_OPS];
} dmb_pass;

/* Per-context counters for roofline reporting. */
typedef struct dmb_stats {
  uint64_t tile_pass_launches;   /* launches of the fused tile kernel                      */
  uint64_t other_launches;       /* every other kernel of this library                     */
  uint64_t fused_ops;            /* dmb_op entries executed                                */
  uint64_t state_bytes_moved;    /* algorithmic HBM bytes of tile passes: 16 B x elements  */
  uint64_t folded_swaps;         /* trailing SWAP ops realised by the relabelling write-back instead */
  uint64_t small_plan_launches;  /* launches that ran a whole pass list (states of <= 16 tiles), counted in
                                    tile_pass_launches too                                   */
  uint64_t chained_ops;          /* ops that ran in the shared-memory round trip of the op before them (two
                                    consecutive CNOT-kind ops on the same ordered digit pair, e.g. the two CNOTs
                                    of a controlled-phase gate; environment DMB_CHAIN_OPS=0 disables this)  */
} dmb_stats;

/* ---- context ------------------------------------------------------------------------ */
int dmb_abi_version(void);
size_t dmb_sizeof_op(void);     /* layout checks for foreign-language bindings */
size_t dmb_sizeof_pass(void);
const char* dmb_last_error(void);
/* Binds to CUDA device `device`; fails (non-zero) if no such device / not sm_100. */
int dmb_create(int device, dmb_ctx** out);
int dmb_destroy(dmb_ctx* ctx);
/* cudaStream_t to enqueue on (e.g. torch.cuda.current_stream().cuda_stream); 0 = default. */
int dmb_set_stream(dmb_ctx* ctx, void* cuda_stream);
int dmb_sync(dmb_ctx* ctx);                                   /* synchronous */
int dmb_get_stats(dmb_ctx* ctx, dmb_stats* out);
int dmb_reset_stats(dmb_ctx* ctx);
/* Tile-kernel variant for 4^6-coefficient tiles (A/B switch for measurements, not a feature):
 *   0 = the shipped kernel k_tile_pass6: 128 threads per tile, one 32 KiB cp.async stage per CTA, 5 CTAs per SM;
 *       ops that leave tile digit 0 free move two 16-blocks per thread with 128-bit shared-memory accesses
 *   1 = the generic register-staged kernel k_tile_pass<6> (one tile per CTA iteration), the A/B baseline. */
int dmb_set_tile_variant(dmb_ctx* ctx, int variant);

/* ---- state initialisation (replaces DmSimulatorPy._initialize_densitymatrix,
 *      dm_simulator.py:284-349: every built-in initial state is a product state) --------
 * state[idx] = scale * prod_q v[q][digit_q(g)],  g = (rank_bits << n_bits) | idx, where
 * digit_q(g) = 2*bit(g, hi[q]) + bit(g, lo[q]).  n_bits = log2(#local elements).         */
int dmb_init_product(dmb_ctx* ctx, double* state, int n_bits, uint64_t rank_bits,
                     int n_qubits, const int32_t* hi, const int32_t* lo,
                     const double* v /* [n_qubits][4] */, double scale);

/* ---- gate / noise / projection application (replaces _add_unitary_single :367-384,
 *      _add_unitary_two :386-395 + basicaertools.cx_gate_dm_matrix, _add_decoherence_and_
 *      amp_decay :397-425, _add_qasm_measure_X/Y/Z/N :574-704, the projections of
 *      _pauli_string_expectation :553-566, the mask of _add_bell_basis_measure :749-756,
 *      _add_qasm_reset :810-823) -- pre-scheduled into tile passes by the host.
 * `passes` is HOST memory; it is consumed before the call returns (kernel parameters, or a
 * staged copy).  For states of at most 4^8 coefficients a call with several passes is ONE launch:
 * the tiles' CTAs form a thread-block cluster and meet at the cluster barrier between passes
 * (dmb_stats.small_plan_launches; environment DMB_SMALL_PATH=0 disables this).
 * DMB_OP_SWAP ops WITHOUT maps at the end of a pass's op list are not executed in shared
 * memory: they are realised by the tile's write-back addressing (same result, counted in
 * dmb_stats.folded_swaps; environment DMB_FOLD_SWAPS=0 disables this).                   */
int dmb_apply_passes(dmb_ctx* ctx, double* state, int n_bits,
                     const dmb_pass* passes, size_t n_passes);

/* ---- host-side gate-fusion scheduler (no device work; generalises the reference's per-qubit
 *      U3 merge, basicaertools.py:251-307 single_gate_merge, from "adjacent single-qubit gates"
 *      to "every op whose qubits fit one 4^6-coefficient tile") ---------------------------------
 * Input: the op stream in program order on QUBIT ids -- every single-qubit map (rotations, noise,
 * projections, resets) has already been multiplied into the pa / pb of the next two-qubit op of
 * its qubit by the caller.  pos[q] = digit position of qubit q (positions >= n_digits are allowed
 * for qubits no op touches: the global slots of a sharded state); it is advanced to the layout
 * after the last emitted pass.  The scheduler packs the stream into tile passes of <= max_tile
 * digits and <= max_ops ops (greedy list scheduling over the next `window` ops; an op is hoisted
 * past skipped ops only when it shares no qubit with them, which is exact because ops on disjoint
 * qubits commute).  Digit positions 0 and 1 belong to every tile (128-byte runs in HBM), so after
 * each pass it chooses which two qubits of the tile should occupy them next -- the pair whose
 * greedy next pass runs the most ops -- and appends the DMB_OP_SWAP ops that move them there
 * (trailing swaps are free: dmb_apply_passes folds them into the write-back).
 *   min_tail > 0: stop as soon as no more than min_tail ops are left and return their indices in
 *     left[0..*n_left) (program order) -- a caller that is still producing ops keeps them queued,
 *     so pass boundaries do not depend on where the stream was cut;
 *   moves / n_moves (may be NULL): *n_moves pairs (qubit, digit position) to realise after the
 *     last op (the sharded engine parks the qubits it is about to send away in the top local
 *     slots).  Moves whose digits fit the last pass's tile ride along as trailing swaps; the rest
 *     are written back to moves[] and counted in *n_moves.
 * out must have room for out_cap passes (n_ops always suffices).                                 */
typedef struct dmb_qop {
  int32_t kind;      /* DMB_OP_*                                                            */
  int32_t flags;     /* DMB_HAS_PA | DMB_HAS_PB                                             */
  int32_t qa, qb;    /* qubit ids; qb = -1: lone single-qubit map (kind DMB_OP_MATS)        */
  double pa[12];     /* rows 1..3 of the pending map of qubit qa                            */
  double pb[12];     /* rows 1..3 of the pending map of qubit qb                            */
  double coef[16];
} dmb_qop;           /* 336 bytes */
size_t dmb_sizeof_qop(void);
/* strategy: how the qubits of a pass's tile are chosen.
 *   DMB_SCHED_PROGRAM_ORDER  walk the stream in program order and let the first ops that fit claim
 *                            the tile (list scheduling as it is usually done);
 *   DMB_SCHED_TILE_SEARCH    grow the tile by the qubits of one ready op at a time, keeping the
 *                            choice that lets the most ops run: finds time-skewed windows on
 *                            nearest-neighbour circuits (fewer, fuller passes).                     */
enum { DMB_SCHED_PROGRAM_ORDER = 0, DMB_SCHED_TILE_SEARCH = 1 };
int dmb_schedule(const dmb_qop* ops, size_t n_ops, int32_t* pos, int n_qubits, int n_digits,
                 int max_tile, int max_ops, int window, int strategy, size_t min_tail,
                 int32_t* moves, int32_t* n_moves,
                 dmb_pass* out, size_t out_cap, size_t* n_out, int32_t* left, size_t* n_left);

/* ---- multi-GPU: fused exchange + tile pass over NVLink peer memory -----------------------
 * In the pass that follows a global<->local slot swap (qiskit-aakash_b200/distributed.py) the
 * tile kernel PULLS its input straight from the ranks that hold it in the old layout: element
 * idx of the new local layout is read from  (const double*)src_tab[idx >> block_shift] + idx
 * (peer-mapped addresses, 2^tab_bits entries, block_shift + tab_bits == n_bits), the fused ops
 * are applied, and the tile is written to dst_state[idx].  One kernel = all-to-all + ops; a pass
 * with n_ops == 0 is a pure exchange.  All ranks must have finished writing their old buffers
 * (host barrier) before the call, and must not overwrite them until every rank has finished.
 * dmb_ipc_export / dmb_ipc_open turn a device pointer of one process into a mapped pointer of
 * another (cudaIpcGetMemHandle / cudaIpcOpenMemHandle).  Mappings are reference counted per handle:
 * every dmb_ipc_open is paired with one dmb_ipc_close, the last close unmaps the peer allocation.
 * push != 0 selects the mirror image: the pass runs in place on `dst_state` (this rank's
 * current buffer, old layout) and STORES every tile to  (double*)src_tab[idx >> block_shift] + idx,
 * i.e. into the buffers of the ranks that own the data after the swap (remote stores). */
int dmb_apply_pass_remote(dmb_ctx* ctx, double* dst_state, int n_bits, const dmb_pass* pass,
                          const uint64_t* src_tab, int tab_bits, int block_shift, int push);
/* The same with the table index gathered from n_sel <= 5 SELECTED bits of the element index (bit sel_bits[j] of idx is
 * bit j of the table index) instead of its high bits: a global slot swapped with an ARBITRARY local slot, so that the
 * evicted qubit need not be parked on the top local slots by a pass of its own (distributed.py).                   */
int dmb_apply_pass_remote_sel(dmb_ctx* ctx, double* dst_state, int n_bits, const dmb_pass* pass,
                              const uint64_t* src_tab, int n_sel, const int32_t* sel_bits, int push);
int dmb_ipc_export(dmb_ctx* ctx, const void* dev_ptr, unsigned char* handle64, uint64_t* offset);
int dmb_ipc_open(dmb_ctx* ctx, const unsigned char* handle64, uint64_t offset, void** out_ptr);
int dmb_ipc_close(dmb_ctx* ctx, const unsigned char* handle64);

/* ---- readout ------------------------------------------------------------------------
 * I/B-marginal of _add_ensemble_measure (:427-481) for basis X/Y/Z: for c in [0,2^n_qubits)
 *   out[c] = sum over the listed digit values of  w * state[...],
 * generalised so that pending single-qubit maps and sharded (global) qubits can be folded
 * in: qubit k (result bit k of c) reads digit (hi[k], lo[k]) of the global index
 * g = (rank_bits << n_bits) | idx and contributes weight wt[k][c_k][digit]  (wt is
 * [n_qubits][2][4]); digits with zero weight are skipped, bits >= n_bits must match
 * rank_bits or the term is dropped.  Then dmb_fwht turns the marginal into the 2^n
 * probabilities  P(r) = sum_c (-1)^(r.c) out[c]. `out` is a device buffer of 2^n_qubits. */
int dmb_marginal(dmb_ctx* ctx, const double* state, int n_bits, uint64_t rank_bits,
                 int n_qubits, const int32_t* hi, const int32_t* lo,
                 const double* wt /* [n_qubits][2][4] */, double* out);
int dmb_fwht(dmb_ctx* ctx, double* vec, int n_qubits);
/* One step of the 'N'-basis ensemble contraction (:457-462): in viewed as [H][4][L],
 * out[h][0][l] = in[h][0][l], out[h][1][l] = nv . in[h][1..3][l].                        */
int dmb_contract_digit(dmb_ctx* ctx, const double* in, double* out, uint64_t H, uint64_t L,
                       const double nv[3]);
/* Gather of a few coefficients to HOST memory (synchronous): expectation values
 * (:569-570), Bell reduced matrix (:744-747, :758-760), trace check (:363).              */
int dmb_read_coeffs(dmb_ctx* ctx, const double* state, const uint64_t* idx, size_t k,
                    double* out_host);
/* Pauli -> matrix basis (_compute_densitymatrix :1198-1255).  `work` and `out` are device
 * buffers of 4^n_qubits complex doubles (interleaved re,im); out is the row-major
 * 2^n x 2^n matrix, qubit 0 = MSB of row and column index.                               */
int dmb_to_matrix(dmb_ctx* ctx, const double* state, int n_qubits, double* work, double* out);
/* dot(a, b) (_state_overlap :1277-1282, without the 2^n factor); synchronous.            */
int dmb_dot(dmb_ctx* ctx, const double* a, const double* b, uint64_t count, double* out_host);
/* vec[|x| < thr] = 0 (_get_densitymatrix :1263).                                         */
int dmb_chop(dmb_ctx* ctx, double* state, uint64_t count, double thr);

/* ---- host <-> device transfer of coefficients ('stored_density_matrix' :337-343,
 *      'coeffmatrix' :1181-1183); host memory should be pinned for full PCIe speed ------ */
int dmb_upload(dmb_ctx* ctx, double* state, const double* host, uint64_t offset, uint64_t count);
int dmb_download(dmb_ctx* ctx, const double* state, double* host, uint64_t offset, uint64_t count);
/* The same without waiting: `host` (pinned) is valid after the next dmb_sync.  Lets a caller queue several
 * readouts of one job and wait once.                                                                    */
int dmb_download_async(dmb_ctx* ctx, const double* state, double* host, uint64_t offset, uint64_t count);

#ifdef __cplusplus
}
#endif
#endif /* DMB200_H */
