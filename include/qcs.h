/* libqcs -- C ABI of the B200 (sm_100a) state-vector engine behind the qcircuits Python API.
 *
 * The reference (grey-area/qcircuits v0.6.0) is pure Python/NumPy and has no FFI of its own;
 * the drop-in boundary is its Python class API.  Each entry point below replaces the NumPy
 * call site(s) named in its comment (file:line into the reference tree) and is what
 * qcircuits_b200/_lib.py binds with ctypes (see INTEGRATION.md for the binding a maintainer
 * of the reference would add).
 *
 * Conventions
 *  - A state of n qubits is a flat device array of 2^n complex amplitudes, interleaved (re, im),
 *    float (QCS_C64) or double (QCS_C128).  Flat bit b (0 = least significant) is qcircuits
 *    qubit n-1-b (C-order axis 0 = qubit 0 = most significant bit; reference state.py:60,111).
 *  - Gate matrices are HOST pointers to 2^k x 2^k complex doubles, row-major, interleaved
 *    (re, im), exactly Operator.to_matrix() (operators.py:75-93).  Operator qubit j (MSB-first
 *    in the matrix index) acts on flat bit bitpos[j].
 *  - Lazy normalisation.  The reference renormalises after every Operator(State)
 *    (operators.py:275-276 -> state.py:97).  Here a buffer is paired with a device double
 *    "norm2" holding the squared norm of its contents; the represented state is
 *    buf / sqrt(norm2).  Kernels take `norm_in` (device pointer or NULL = 1.0), fold
 *    1/sqrt(*norm_in) into their arithmetic, and write the squared norm of what they produced
 *    to `norm_out` (device pointer or NULL), so no extra pass over the amplitudes is needed.
 *  - libqcs never allocates or frees amplitude memory.  `ws` is a caller-owned, zero-initialised
 *    device scratch of at least qcs_workspace_bytes() bytes; calls that share a `ws` must be
 *    stream-ordered.  All calls are asynchronous on `stream` (a cudaStream_t).
 *  - Every function returns 0 on success or a negative QCS_ERR_* code; qcs_last_error() gives
 *    a thread-local message.  Nothing throws across the boundary.
 */
#ifndef QCS_H
#define QCS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QCS_OK 0
#define QCS_ERR_ARG (-1)
#define QCS_ERR_CUDA (-2)
#define QCS_ERR_UNSUPPORTED (-3)
#define QCS_ERR_NOMEM (-4) /* qcs_dist_alloc: the device allocation failed                       */
#define QCS_ERR_PEER (-5)  /* multi-GPU: peer access / CUDA IPC between the devices is unavailable */

#define QCS_C64 0  /* complex64  amplitudes (float2)  */
#define QCS_C128 1 /* complex128 amplitudes (double2) */
#define QCS_F64 2  /* real float64 input (marginals of a probability vector) */

#define QCS_MAX_OP_QUBITS 4 /* largest gate the tile engine applies in one op   */
#define QCS_MAX_OPS 128     /* ops per fused pass                                */

#define QCS_OP_DENSE 0
#define QCS_OP_DIAG 1

/* One gate of a fused pass. */
typedef struct qcs_op {
    int32_t kind;                       /* QCS_OP_DENSE or QCS_OP_DIAG                                  */
    int32_t k;                          /* operator qubits, 1..QCS_MAX_OP_QUBITS                         */
    int32_t bitpos[QCS_MAX_OP_QUBITS];  /* flat bit each operator qubit acts on, MSB-first              */
    uint64_t ctrl_mask;                 /* op acts only on indices i with (i & ctrl_mask) == ctrl_mask  */
    const double *matrix;               /* host; DENSE: 2^k x 2^k complex, DIAG: the 2^k diagonal       */
} qcs_op;

typedef void *qcs_stream; /* cudaStream_t */

int qcs_version(void);
const char *qcs_last_error(void);
int64_t qcs_workspace_bytes(void);

/* Tile geometry for a fused pass (used by the host-side fusion scheduler).
 * low_bits  : flat bits 0..low_bits-1 are always inside a tile (contiguous run staged in shared memory)
 * tile_bits : at most tile_bits bits (low ones included) can be touched by dense ops of one pass */
int qcs_tile_limits(int dtype, int *low_bits, int *tile_bits);

/* Tuning hook (not a replacement for any reference call): 0 = cp.async ring with padded shared-memory
 * layout (default), 1 = TMA bulk-copy ring.  Also settable with QCS_TILE_MODE=pipe|bulk. */
int qcs_set_tile_mode(int mode);

/* K1/K3: Operator._apply on a State (operators.py:225-278): dst = M applied to src on bitpos,
 * scaled by 1/sqrt(*norm_in); squared norm of dst to *norm_out.  k <= QCS_MAX_OP_QUBITS.
 * dst may equal src (in-place). */
int qcs_apply_gate(void *dst, const void *src, int n, int dtype, const double *matrix, int k,
                   const int32_t *bitpos, const double *norm_in, double *norm_out, void *ws,
                   qcs_stream stream);

/* Same contraction for an operator of any size (k up to n), matrix already on the device in the
 * amplitude dtype (2^k x 2^k, row-major).  Used for the "operator spans many qubits" case, e.g.
 * the Grover iterate of examples/grover_algorithm.py:41,49.  dst must not alias src. */
int qcs_apply_dense(void *dst, const void *src, int n, int dtype, const void *dev_matrix, int k,
                    const int32_t *bitpos, const double *norm_in, double *norm_out, void *ws,
                    qcs_stream stream);

/* A run of gates fused into ONE pass over the state (north_star (1): "runs of adjacent gates fuse
 * into one pass").  All dense ops' target bits >= low_bits must number at most
 * tile_bits - low_bits (see qcs_tile_limits), else QCS_ERR_UNSUPPORTED.  Diagonal ops and control
 * bits may sit anywhere.  Ops are applied in order. */
int qcs_apply_fused(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops,
                    const double *norm_in, double *norm_out, void *ws, qcs_stream stream);

/* The same pass for a caller that does not keep the lazy norm but knows a bound (the sharded engine: the norm
 * belongs to the whole state, not to a shard).  bound2_in: device scalar >= the squared 2-norm of src, or NULL.
 * It only RANGES the fp16 operands of the tensor-core sweeps (complex64; amplitudes are scaled by a power of two so
 * that the bound maps to 2^15; the error of a sweep is about 2^-40 of the bound, so a bound within a few powers of
 * two of the largest amplitude times 2^(n/2) keeps the 1e-5 tolerance); the result is not scaled by it.  With
 * norm_in == NULL and bound2_in == NULL the library measures the norm itself (one extra read of src). */
int qcs_apply_fused_bounded(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops,
                            const double *norm_in, double *norm_out, const double *bound2_in, void *ws,
                            qcs_stream stream);

/* Planning only, no launch and no device needed (diagnostics for the fusion scheduler and the tests):
 * how the tile engine would run this pass.  stats[0..5] = tile bits, sweeps, register ops, shared-memory
 * dense ops, tile-level predicates, scale mode; stats[6] = number of input ops that run in the last sweep and
 * stats[116..127] the first 12 of their indices (-1 terminated); stats[8 + a] = number of register ops of arm
 * a (the arm table of csrc/gen_regops.py, a < 108).  `stats` has 128 entries. */
int qcs_plan_stats(int n, int dtype, const qcs_op *ops, int nops, int32_t *stats);

/* Planning only: as qcs_plan_stats, plus the sweep that executes each input op: sweep_of_op[i] = index of the sweep
 * (execution order), + 1000 when that sweep applies a tensor-core block.  The fusion scheduler uses it to cut a pass
 * after its last tensor-core sweep (ops of later sweeps are last in execution order and may move to the next pass). */
int qcs_plan_order(int n, int dtype, const qcs_op *ops, int nops, int32_t *stats, int32_t *sweep_of_op);

/* Planning only, complex64: the tensor-core sweeps of the pass (north_star (1): "fused blocks ... run on tensor
 * cores").  A sweep multiplies a run of gates into one dense 2^5 x 2^5 block on five tile bits (the contraction
 * of operators.py:261 for all of them at once) and applies it with tcgen05.mma kind::f16 as a two-term split of
 * both operands (hi.hi + hi.lo + lo.hi, fp32 accumulation); passes with a non-unitary op are not taken.
 * stats as qcs_plan_stats, with stats[7] = sweeps with a block | (input ops inside blocks) << 8.
 * images: 8 slots x 8192 bytes, the shared-memory operand images ("hi" fp16 image then "lo" image; K-major, no
 * swizzle: row r, column c at byte (r%8)*16 + (r/8)*512 + (c/8)*128 + (c%8)*2; rows 0..31 = Re U, 32..63 = Im U).
 * info: 8 int32 per block: flat bits of block-index bits 0..4, register ops before the block, slot, sweep index.
 * Returns QCS_ERR_UNSUPPORTED when the pass has no dense block (it then runs on the register-sweep kernel). */
int qcs_plan_blocks(int n, const qcs_op *ops, int nops, int32_t *stats, void *images, int32_t *info);

/* K5: State.probabilities (state.py:200): probs[i] = |src[i]|^2 / *norm_in ; sum to *sum_out. */
int qcs_abs2(double *probs, const void *src, int n, int dtype, const double *norm_in,
             double *sum_out, void *ws, qcs_stream stream);

/* State.dot (state.py:89): out[0..1] = sum conj(a) * b (raw buffers, no lazy scale applied). */
int qcs_dot(const void *a, const void *b, int n, int dtype, double *out2, void *ws, qcs_stream stream);

/* State.renormalize_ / scalar * and / (state.py:97,174-185): dst = (re + i im)/sqrt(*norm_in) * src. */
int qcs_scale(void *dst, const void *src, int n, int dtype, double re, double im,
              const double *norm_in, double *norm_out, void *ws, qcs_stream stream);

/* State + State, State - State, DensityOperator.from_ensemble accumulation (state.py:165-169):
 * dst = alpha/sqrt(*norm_a) * a + beta/sqrt(*norm_b) * b  (alpha, beta complex, host). */
int qcs_lincomb(void *dst, const void *a, const void *b, int n, int dtype, const double *alpha2,
                const double *beta2, const double *norm_a, const double *norm_b, double *norm_out,
                void *ws, qcs_stream stream);

/* K9: Tensor.tensor_product (tensors.py:76): dst[(i << nb) | j] = a[i] * b[j] (a's qubits high). */
int qcs_kron(void *dst, const void *a, int na, const void *b, int nb, int dtype,
             const double *norm_a, const double *norm_b, double *norm_out, void *ws, qcs_stream stream);

/* K4: np.transpose relabelling (state.py:117,162): dst bit src_bit_of_dst_bit[b] <- ...:
 * dst[i] = src[j] where bit src_bit[b] of j equals bit b of i. */
int qcs_permute_bits(void *dst, const void *src, int n, int dtype, const int32_t *src_bit,
                     qcs_stream stream);

/* OperatorBase.adj (operators.py:95-114: np.conj + transpose of every (out, in) axis pair) and the conjugated matrix of
 * U rho U^dagger (operators.py:317-321), for operator tensors that live on the device:
 * dst[i] = conj(src[j]) with the same index map as qcs_permute_bits.  The adjoint of a d-qubit operator tensor
 * (2d flat bits, out-wire of qubit q on bit 2(d-1-q)+1) is src_bit[b] = b ^ 1. */
int qcs_conj_permute_bits(void *dst, const void *src, int n, int dtype, const int32_t *src_bit,
                          qcs_stream stream);

/* U_f applied to a State (operators.py:809-853 builds the 2^d x 2^d permutation tensor with a Python loop and contracts
 * it with np.tensordot, :261): dst[i] = src[i XOR (f(x(i)) << tbit)] / sqrt(*norm_in), where x(i) gathers the flat bits
 * xbits[0..nx-1] of i (xbits[0] = most significant bit of x) and f is a packed truth table on the device (bit x of
 * dev_table: word x >> 5, bit x & 31).  A permutation: the squared norm of dst is that of src (1 after the scale).
 * dst must not alias src. */
int qcs_apply_uf(void *dst, const void *src, int n, int dtype, const uint32_t *dev_table, const int32_t *xbits,
                 int nx, int tbit, const double *norm_in, qcs_stream stream);

/* State factories zeros/ones/bitstring (state.py:455-551): buf = 0 except buf[index] = 1. */
int qcs_set_basis(void *buf, int n, int dtype, uint64_t index, qcs_stream stream);

/* K6: State._measurement_probabilities (state.py:262-269): float64 marginals over the m listed
 * bits (bitpos[0] = MSB of the outcome), divided by *norm_in.  dtype may be QCS_F64 (src is a
 * real probability vector; used for DensityOperator marginals, density_operator.py:140-147).
 * probs_out has 2^m doubles and is overwritten. */
int qcs_marginal_probs(double *probs_out, const void *src, int n, int dtype, const int32_t *bitpos,
                       int m, const double *norm_in, void *ws, int64_t ws_bytes, qcs_stream stream);

/* K8: the collapse of State.measure (state.py:368-378): keep the amplitudes whose bits at bitpos
 * equal `outcome` (bitpos[0] <-> MSB of outcome), times `scale`.  remove != 0: dst has n-m qubits
 * (surviving bits keep their order); remove == 0: dst has n qubits, zero elsewhere.
 * Squared norm of dst to *norm_out.  dst must not alias src when remove != 0. */
int qcs_collapse(void *dst, const void *src, int n, int dtype, const int32_t *bitpos, int m,
                 uint64_t outcome, int remove, double scale, double *norm_out, void *ws,
                 qcs_stream stream);

/* K11: DensityOperator.probabilities (density_operator.py:60): probs[i] = Re rho[i, i] for the
 * interleaved rank-2n tensor (row bit of qubit q = flat bit 2(n-1-q)+1, column bit 2(n-1-q)). */
int qcs_diag_probs(double *probs, const void *rho, int n, int dtype, double *sum_out, void *ws,
                   qcs_stream stream);

/* K13: DensityOperator.from_ensemble (density_operator.py:106-125):
 * rho[interleave(r, c)] (+)= p / *norm_psi * psi[r] * conj(psi[c]);  accumulate != 0 adds. */
int qcs_outer_accumulate(void *rho, const void *psi, int n, int dtype, double p,
                         const double *norm_psi, int accumulate, qcs_stream stream);

/* DensityOperator._reduced_tensor (density_operator.py:127-138: transposes + np.einsum('ii...->...')):
 * out = partial trace of the d-qubit rho over every qubit not listed; output qubit j is qubit keep[j]
 * (same interleaved row / column layout on 2 * nkeep bits).  out must not alias rho. */
int qcs_partial_trace(void *out, const void *rho, int d, int dtype, const int32_t *keep, int nkeep,
                      qcs_stream stream);

/* ---- multi-GPU: the qubit exchange of a sharded state over NVLink peer memory (north star (4)) ---------------------------
 * A state of n_local + log2(P) qubits lives on P = 2, 4, 8 or 16 GPUs, shard r holding the amplitudes whose top log2(P)
 * flat bits ("rank bits") spell r.  The reference has no counterpart (one NumPy array, state.py:20); the operation is
 * the np.transpose that would bring a high axis down (state.py:117) when that axis is spread over devices.  The
 * exchange of g (rank bit, local bit) pairs:
 *     dst_r[i] = src_r'[i']  with  bit local_bits[j] of i' = rank bit global_bits[j] of r,
 *                                  rank bit global_bits[j] of r' = bit local_bits[j] of i,  all other bits unchanged
 * (global_bits index the rank bits 0..log2(P)-1, local_bits the flat bits of a shard; complex64 needs local_bits >= 1).
 * One kernel per GPU moves its 16-byte vectors straight between its own and the peers' HBM (P2P stores, or loads, over
 * NVLink) -- no NCCL, no pack / unpack pass.  Out of place: dst shards must not alias src shards.  All gate / measurement entry points above
 * are per device and are simply called once per shard. */
typedef struct qcs_dist_ctx *qcs_dist_t;

/* One process driving all GPUs: enables peer access between devs[0..ndev-1] (QCS_ERR_PEER if the topology refuses). */
int qcs_dist_init(int ndev, const int *devs, qcs_dist_t *handle);
/* dst_shards[r] / src_shards[r] / streams[r] belong to devs[r].  Ordered after the work already queued on every
 * stream; on return every stream also waits until all peers have finished reading its source shard. */
int qcs_dist_swap(qcs_dist_t handle, void *const *dst_shards, const void *const *src_shards, int n_local, int dtype,
                  const int32_t *global_bits, const int32_t *local_bits, int g, const qcs_stream *streams);
int qcs_dist_destroy(qcs_dist_t handle);

/* One process per GPU (torchrun / MPI): this rank's half of the same exchange, on pointers valid on THIS device -- the
 * caller's own buffer at index `rank`, mappings made with qcs_ipc_open for the other ranks.
 *   push = 1 (preferred: remote stores, 682 GB/s per GPU and direction between two B200s): `local` = this rank's SOURCE
 *            shard, peers[r] = rank r's DESTINATION shard;
 *   push = 0 (remote loads, 627 GB/s): `local` = this rank's DESTINATION shard, peers[r] = rank r's SOURCE shard.
 * The caller orders the ranks: every source complete and every destination free before the call, nothing reused before
 * every rank's call has finished (qcircuits_b200/dist.py brackets it with two one-element NCCL all-reduces on the same
 * stream).  qcs_dist_swap (one process) pushes unless QCS_DIST_PUSH=0. */
int qcs_dist_swap_rank(void *local, const void *const *peers, int nranks, int rank, int n_local, int dtype,
                       const int32_t *global_bits, const int32_t *local_bits, int g, int push, qcs_stream stream);
/* CUDA IPC plumbing for that mode.  Shards that other processes map must be plain cudaMalloc allocations (a caching
 * allocator hands out interior pointers): qcs_dist_alloc / qcs_dist_free are the one place libqcs allocates amplitude
 * memory.  handle64 = the 64 bytes of a cudaIpcMemHandle_t, to be sent to the other ranks by any means. */
int qcs_dist_alloc(void **ptr, int64_t bytes);
int qcs_dist_free(void *ptr);
int qcs_ipc_get_handle(const void *dev_ptr, void *handle64);
int qcs_ipc_open(const void *handle64, void **mapped);
int qcs_ipc_close(void *mapped);

#ifdef __cplusplus
}
#endif
#endif /* QCS_H */
