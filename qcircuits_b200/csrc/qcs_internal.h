// Internal declarations shared by the libqcs translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/qcs.h"

namespace qcs {

// ---- tile engine geometry ----------------------------------------------------------------------
constexpr int TILE_LOW_BITS_C64 = 10;    // 2^10 complex64  = 8 KB contiguous chunk
constexpr int TILE_LOW_BITS_C128 = 9;    // 2^9  complex128 = 8 KB contiguous chunk
constexpr int TILE_MIN_BITS_C64 = 12;    // tiles are padded with free bits up to 32 KB (bytes in flight)
constexpr int TILE_MIN_BITS_C128 = 11;
constexpr int TILE_MAX_BITS_C64 = 13;    // 64 KB tile
constexpr int TILE_MAX_BITS_C128 = 12;   // 64 KB tile
constexpr int TILE_MAX_H = 8;
constexpr int TILE_THREADS = 256;
constexpr int TILE_THREADS_BULK = 512;
constexpr int TILE_MAX_STAGES = 4;
constexpr size_t TILE_SMEM_BUDGET = 220 * 1024;   // dynamic shared memory per SM we plan against
constexpr int TILE_MODE_PIPE = 0;        // cp.async multi-stage pipeline, padded shared-memory layout
constexpr int TILE_MODE_BULK = 1;        // TMA bulk-copy pipeline (cp.async.bulk + mbarrier), linear layout
constexpr int TILE_MODE_DEFAULT = TILE_MODE_PIPE;

constexpr int REG_BITS_C64 = 4;          // amplitudes a thread holds in registers per sweep: 2^4 float2
constexpr int REG_BITS_C128 = 3;         //                                                   2^3 double2
constexpr int REG_BITS_MAX = 4;

// ops that cannot run in registers (dense on >= 2 targets) go through shared memory one at a time
struct TileOp {
    int32_t kind, k, mat, flags;
    int8_t lpos[QCS_MAX_OP_QUBITS];  // tile-local bit of operator qubit j (MSB-first)
    int8_t gpos[QCS_MAX_OP_QUBITS];  // flat bit of operator qubit j
    uint32_t ctrl_local;             // control bits inside the tile (tile-local mask)
    uint32_t pad;
    uint64_t ctrl_outer;             // control bits outside the tile (flat-index mask)
};

// register-level ops of a sweep
constexpr int MAX_OUTER_PREDS = 63;      // slot 63 of the per-tile predicate mask means "no predicate"
struct RegOp {                   // 16 bytes, read with one 128-bit shared-memory load
    uint16_t code;               // arm of the interpreter (gen_regops.py)
    uint8_t reg_care, reg_val;   // masks over the sweep's register bits
    uint16_t base_care, base_val;// tile-local index bits that are not register bits (per-thread predicate)
    int32_t mat;                 // byte offset of the op's coefficients in the pool (16-byte aligned)
    int32_t outer_slot;          // index of the tile-uniform predicate this op depends on, 63 = none
};
struct OuterPred { uint64_t care, val; };   // flat-index bits outside the tile
constexpr int SWEEP_REG = 0, SWEEP_GENERIC = 1;
struct alignas(16) Sweep {
    int32_t kind, first, count, nreg;
    int8_t rpos[REG_BITS_MAX];       // tile-local positions of the register bits, ascending
    int16_t thread_phase;            // some PHASE op of the sweep is uniform over a thread's amplitudes
    int16_t has_dense2;              // the sweep holds a dense two-target register op (GEN2 arms)
    // tensor-core sweeps (tile_kernel_tc): after the register ops the 2^5 x 2^5 block matrix of slot tc_slot is
    // applied on the four register bits + one more tile bit; that bit is thread-index bit 7 of the 256-thread
    // group and is inserted at rank `hrank` among the 8 non-register tile bits (hrank = 7: plain register sweep)
    int16_t tc_slot;                 // -1: no tensor-core block in this sweep
    int16_t hrank;
    // everything the tensor-core kernel needs of the record in ONE 8-byte word (one broadcast shared-memory load per
    // thread and sweep instead of ~19 dependent ones): rpos[j] at bits 4j (j < 4), hrank at 16, tc_slot + 1 at 20,
    // flags at 24 (1: has register ops, 2: dense2 interpreter, 4: thread phase), first register op at 32
    uint64_t packed;                 // (8-byte aligned: read with one 64-bit shared-memory load)
    uint64_t pad_;
    // host-precomputed addressing of the compile-time-geometry path: inserting a zero at rpos[j] is
    // x += x & himask[j]; stride[j] = padded shared-memory distance in bytes between the two values of bit j
    uint32_t stride[REG_BITS_MAX];   // (16-byte aligned: the tensor-core kernel reads the four with one 128-bit load)
    uint32_t himask[REG_BITS_MAX];
};

template <int NSW, int NREG, int NGEN, int MATBYTES>
struct TileProgram {
    static constexpr int MAX_SWEEPS = NSW, MAX_REGOPS = NREG, MAX_GENERIC = NGEN, MAT_BYTES = MATBYTES;
    int32_t n, L, H, nsweeps, nregops, ngen, nouter, stages, mat_bytes, scale_mode;
    double gphase[2];                // pass-level complex factor (global phases of diagonal gates)
    int32_t hbits[TILE_MAX_H];
    uint64_t ntiles;
    const void *src;
    void *dst;
    const double *norm_in;
    double *norm_out;
    void *ws;
    const void *tc_mats;             // device: n_tc block-matrix images of TC_SLOT_BYTES each (tile_kernel_tc)
    const double *tc_norm;           // device: squared norm of src (or an upper bound): ranges the fp16 tensor-core operands
    int32_t n_tc, tc_pad;
    Sweep sweeps[NSW];
    RegOp regops[NREG];
    OuterPred outer[MAX_OUTER_PREDS];
    TileOp gen[NGEN];
    alignas(16) unsigned char mats[MATBYTES];
};
using SmallProg = TileProgram<4, 34, 2, 4096 + 64>;
using LargeProg = TileProgram<72, 420, 64, 12288>;
// passes that run on the tensor-core kernel: its shared memory also holds the block matrices, so the tables are smaller
using TcProg = TileProgram<24, 192, 8, 6144>;

// ---- tensor-core sweeps (complex64, 12-bit tiles) ------------------------------------------------
// A sweep's dense block acts on 5 tile bits: a (2^5 x 2^5 complex) matrix applied to the 128 rows x 32 amplitudes
// of a tile, as 3 x 6 tcgen05.mma kind::f16 instructions (fp16 operands, fp32 accumulation; two-term split of both
// operands: hi*hi + hi*lo + lo*hi, the amplitudes ranged by a power of two taken from the norm of the state).
constexpr int TC_BLOCK_BITS = 5;
constexpr int TC_MAX_SLOTS = 8;                      // block matrices resident in shared memory per pass
constexpr int TC_HALF_BYTES = 64 * 32 * 2;           // one fp16 image: rows 0..31 = Re U, rows 32..63 = Im U; K-major, no swizzle
constexpr int TC_SLOT_BYTES = 2 * TC_HALF_BYTES;     // "hi" image = fp16(U), then the "lo" image = fp16(U - hi)
constexpr int TC_DEFAULT_SLOTS = 6;                  // what the planner uses unless QCS_TC_SLOTS says otherwise
constexpr int64_t WS_TC_NORM_OFFSET = 32640;         // squared norm of src when the caller gave none (tile_apply computes it)
constexpr int TC_DEFAULT_GROUPS = 4;                 // groups per CTA (2 x 256 threads with two ring stages each, or 4 x 128 with one)
constexpr int64_t WS_TC_OFFSET = 32768;              // block-matrix images live in the caller's workspace from here
constexpr int TC_UPLOAD_RING = 8;                    // pinned host staging sets for their upload (one per launch in flight)
static_assert(sizeof(Sweep) % 16 == 0 && offsetof(Sweep, packed) % 8 == 0 && offsetof(Sweep, stride) % 16 == 0,
              "Sweep::packed is read with one 64-bit load, Sweep::stride with one 128-bit load");
static_assert(sizeof(LargeProg) < 32000, "kernel parameter space is 32764 bytes");

struct DeviceInfo { int sms; size_t smem_per_sm; };
const DeviceInfo &device_info();
int tile_mode();
int tile_max_bits(int dtype);
int tune_int(const char *name, int fallback);
int set_error(int code, const char *msg);

int tile_apply(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops, const double *norm_in,
               double *norm_out, const double *bound_in, void *ws, cudaStream_t stream);

int tile_plan_stats(int n, int dtype, const qcs_op *ops, int nops, int32_t *stats);
int tile_plan_order(int n, int dtype, const qcs_op *ops, int nops, int32_t *stats, int32_t *sweep_of_op);
int tile_plan_blocks(int n, const qcs_op *ops, int nops, int32_t *stats, void *images, int32_t *info);

int misc_set_basis(void *buf, int n, int dtype, uint64_t index, cudaStream_t st);
int misc_abs2(double *probs, const void *src, int n, int dtype, const double *norm_in, double *sum_out, void *ws,
              cudaStream_t st);
int misc_dot(const void *a, const void *b, int n, int dtype, double *out2, void *ws, cudaStream_t st);
int misc_lincomb(void *dst, const void *a, const void *b, int n, int dtype, const double *alpha2, const double *beta2,
                 const double *norm_a, const double *norm_b, double *norm_out, void *ws, cudaStream_t st);
int misc_kron(void *dst, const void *a, int na, const void *b, int nb, int dtype, const double *norm_a,
              const double *norm_b, double *norm_out, void *ws, cudaStream_t st);
int misc_permute(void *dst, const void *src, int n, int dtype, const int32_t *src_bit, bool conj, cudaStream_t st);
int misc_marginal(double *out, const void *src, int n, int dtype, const int32_t *bitpos, int m, const double *norm_in,
                  void *ws, int64_t ws_bytes, cudaStream_t st);
int misc_collapse(void *dst, const void *src, int n, int dtype, const int32_t *bitpos, int m, uint64_t outcome,
                  int remove, double scale, double *norm_out, void *ws, cudaStream_t st);
int misc_diag_probs(double *probs, const void *rho, int n, int dtype, double *sum_out, void *ws, cudaStream_t st);
int misc_outer(void *rho, const void *psi, int n, int dtype, double p, const double *norm_psi, int accumulate,
               cudaStream_t st);
int misc_gate1(void *dst, const void *src, int n, int dtype, const double *matrix, int tbit, uint64_t ctrl,
               const double *norm_in, double *norm_out, void *ws, cudaStream_t st);
int misc_partial_trace(void *out, const void *rho, int d, int dtype, const int32_t *keep, int nkeep, cudaStream_t st);
int misc_apply_uf(void *dst, const void *src, int n, int dtype, const uint32_t *dev_table, const int32_t *xbits, int nx,
                  int tbit, const double *norm_in, cudaStream_t st);
int misc_dense(void *dst, const void *src, int n, int dtype, const void *dev_matrix, int k, const int32_t *bitpos,
               const double *norm_in, double *norm_out, void *ws, cudaStream_t st);

struct DistCtx;
int dist_init(int ndev, const int *devs, DistCtx **out);
int dist_destroy(DistCtx *h);
int dist_swap(DistCtx *h, void *const *dst_shards, const void *const *src_shards, int n_local, int dtype,
              const int32_t *global_bits, const int32_t *local_bits, int g, const cudaStream_t *streams);
int dist_swap_rank(void *dst, const void *const *peer_src, int nranks, int rank, int n_local, int dtype,
                   const int32_t *global_bits, const int32_t *local_bits, int g, int push, cudaStream_t st);
bool dist_push_mode();
int ipc_get_handle(const void *dev_ptr, void *handle64);
int ipc_open(const void *handle64, void **mapped);
int ipc_close(void *mapped);
int dist_alloc(void **ptr, int64_t bytes);
int dist_free(void *ptr);

}  // namespace qcs
