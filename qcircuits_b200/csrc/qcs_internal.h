// Internal declarations shared by the libqcs translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/qcs.h"

namespace qcs {

// ---- tile engine geometry ----------------------------------------------------------------------
constexpr int TILE_LOW_BITS_C64 = 10;    // 2^10 complex64  = 8 KB contiguous chunk
constexpr int TILE_LOW_BITS_C128 = 9;    // 2^9  complex128 = 8 KB contiguous chunk
constexpr int TILE_MAX_BITS_C64 = 13;    // 64 KB tile
constexpr int TILE_MAX_BITS_C128 = 12;   // 64 KB tile
constexpr int TILE_MAX_H = 8;
constexpr int TILE_THREADS = 256;
constexpr int TILE_THREADS_BULK = 512;
constexpr int TILE_MAX_STAGES = 4;
constexpr size_t TILE_SMEM_BUDGET = 200 * 1024;   // dynamic shared memory per SM we plan against
constexpr size_t TILE_STAGE_BUDGET = 96 * 1024;   // per CTA, pipelined variant
constexpr int TILE_MODE_PLAIN = 0;
constexpr int TILE_MODE_BULK = 1;
constexpr int TILE_MODE_DEFAULT = TILE_MODE_PLAIN;
constexpr int TILEOP_SCALE = 1;

struct TileOp {
    int32_t kind, k, mat, flags;
    int8_t lpos[QCS_MAX_OP_QUBITS];  // tile-local bit of operator qubit j (MSB-first); -1 = outside the tile (diag only)
    int8_t gpos[QCS_MAX_OP_QUBITS];  // flat bit of operator qubit j
    uint32_t ctrl_local;             // control bits inside the tile (tile-local mask)
    uint32_t pad;
    uint64_t ctrl_outer;             // control bits outside the tile (flat-index mask)
};

template <int NOPS, int MATBYTES>
struct TileProgram {
    static constexpr int MAXOPS = NOPS;
    int32_t n, L, H, nops, stages, mat_bytes;
    int32_t hbits[TILE_MAX_H];
    uint64_t ntiles;
    const void *src;
    void *dst;
    const double *norm_in;
    double *norm_out;
    void *ws;
    TileOp ops[NOPS];
    alignas(16) unsigned char mats[MATBYTES];
};
using SmallProg = TileProgram<4, 4096 + 64>;
using LargeProg = TileProgram<QCS_MAX_OPS + 2, 20480>;
static_assert(sizeof(LargeProg) < 32000, "kernel parameter space is 32764 bytes");

struct DeviceInfo { int sms; size_t smem_per_sm; };
const DeviceInfo &device_info();
int tile_mode();
int set_error(int code, const char *msg);

int tile_apply(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops, const double *norm_in,
               double *norm_out, void *ws, cudaStream_t stream);

int misc_set_basis(void *buf, int n, int dtype, uint64_t index, cudaStream_t st);
int misc_abs2(double *probs, const void *src, int n, int dtype, const double *norm_in, double *sum_out, void *ws,
              cudaStream_t st);
int misc_dot(const void *a, const void *b, int n, int dtype, double *out2, void *ws, cudaStream_t st);
int misc_lincomb(void *dst, const void *a, const void *b, int n, int dtype, const double *alpha2, const double *beta2,
                 const double *norm_a, const double *norm_b, double *norm_out, void *ws, cudaStream_t st);
int misc_kron(void *dst, const void *a, int na, const void *b, int nb, int dtype, const double *norm_a,
              const double *norm_b, double *norm_out, void *ws, cudaStream_t st);
int misc_permute(void *dst, const void *src, int n, int dtype, const int32_t *src_bit, cudaStream_t st);
int misc_marginal(double *out, const void *src, int n, int dtype, const int32_t *bitpos, int m, const double *norm_in,
                  void *ws, int64_t ws_bytes, cudaStream_t st);
int misc_collapse(void *dst, const void *src, int n, int dtype, const int32_t *bitpos, int m, uint64_t outcome,
                  int remove, double scale, double *norm_out, void *ws, cudaStream_t st);
int misc_diag_probs(double *probs, const void *rho, int n, int dtype, double *sum_out, void *ws, cudaStream_t st);
int misc_outer(void *rho, const void *psi, int n, int dtype, double p, const double *norm_psi, int accumulate,
               cudaStream_t st);
int misc_dense(void *dst, const void *src, int n, int dtype, const void *dev_matrix, int k, const int32_t *bitpos,
               const double *norm_in, double *norm_out, void *ws, cudaStream_t st);

}  // namespace qcs
