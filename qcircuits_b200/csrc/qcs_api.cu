// extern "C" surface of libqcs (include/qcs.h): argument checks, error strings, dispatch.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "qcs_common.cuh"
#include "qcs_internal.h"

namespace qcs {

static thread_local char g_err[256] = "";

int set_error(int code, const char *msg) {
    std::snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

const DeviceInfo &device_info() {
    static thread_local DeviceInfo info[16];
    static thread_local bool have[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) dev = 0;
    if (!have[dev]) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int smem = 0;
        cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        info[dev].sms = sms > 0 ? sms : 148;
        info[dev].smem_per_sm = (size_t)smem;
        have[dev] = true;
    }
    return info[dev];
}

int tune_int(const char *name, int fallback) {
    const char *e = std::getenv(name);
    if (!e || !*e) return fallback;
    return std::atoi(e);
}

// Largest tile in bits.  By default 32 KB tiles (two 256-thread CTAs per SM): measured faster than 64 KB tiles
// (one 512-thread CTA per SM, every sweep in lockstep across the SM) even though a 64 KB tile holds one more
// window bit -- 300 vs 326 ms on the headline circuit.  QCS_TILE_SHRINK=0 allows the 64 KB tiles.
int tile_max_bits(int dtype) {
    const int full = dtype == QCS_C64 ? TILE_MAX_BITS_C64 : TILE_MAX_BITS_C128;
    int shrink = tune_int("QCS_TILE_SHRINK", 1);       // 2: 16 KB tiles, four 128-thread CTAs per SM
    if (shrink < 0) shrink = 0;
    if (shrink > 2) shrink = 2;
    return full - shrink;
}

static int g_tile_mode = -1;
int tile_mode() {
    if (g_tile_mode < 0) {
        const char *e = std::getenv("QCS_TILE_MODE");
        g_tile_mode = (e && std::strcmp(e, "pipe") == 0)   ? TILE_MODE_PIPE
                      : (e && std::strcmp(e, "bulk") == 0) ? TILE_MODE_BULK
                                                           : TILE_MODE_DEFAULT;
    }
    return g_tile_mode;
}

static int check_state(const void *p, int n, int dtype) {
    if (!p) return set_error(QCS_ERR_ARG, "null amplitude pointer");
    if (n < 1 || n > 40) return set_error(QCS_ERR_ARG, "qubit count must be in 1..40");
    if (dtype != QCS_C64 && dtype != QCS_C128) return set_error(QCS_ERR_ARG, "dtype must be QCS_C64 or QCS_C128");
    return QCS_OK;
}

}  // namespace qcs

using namespace qcs;

extern "C" {

int qcs_version(void) { return 100; }
const char *qcs_last_error(void) { return g_err; }
int64_t qcs_workspace_bytes(void) { return WS_BYTES; }

int qcs_set_tile_mode(int mode) {
    if (mode != TILE_MODE_PIPE && mode != TILE_MODE_BULK) return set_error(QCS_ERR_ARG, "tile mode must be 0 or 1");
    g_tile_mode = mode;
    return QCS_OK;
}

int qcs_tile_limits(int dtype, int *low_bits, int *tile_bits) {
    if (dtype != QCS_C64 && dtype != QCS_C128) return set_error(QCS_ERR_ARG, "bad dtype");
    if (low_bits) *low_bits = dtype == QCS_C64 ? TILE_LOW_BITS_C64 : TILE_LOW_BITS_C128;
    if (tile_bits) *tile_bits = tile_max_bits(dtype);
    return QCS_OK;
}

int qcs_plan_stats(int n, int dtype, const qcs_op *ops, int nops, int32_t *stats) {
    if (!ops || !stats) return set_error(QCS_ERR_ARG, "null argument");
    if (nops < 1 || nops > QCS_MAX_OPS) return set_error(QCS_ERR_ARG, "nops must be in 1..QCS_MAX_OPS");
    return tile_plan_stats(n, dtype, ops, nops, stats);
}

int qcs_plan_order(int n, int dtype, const qcs_op *ops, int nops, int32_t *stats, int32_t *sweep_of_op) {
    if (!ops || !stats || !sweep_of_op) return set_error(QCS_ERR_ARG, "null argument");
    if (nops < 1 || nops > QCS_MAX_OPS) return set_error(QCS_ERR_ARG, "nops must be in 1..QCS_MAX_OPS");
    return tile_plan_order(n, dtype, ops, nops, stats, sweep_of_op);
}

int qcs_plan_blocks(int n, const qcs_op *ops, int nops, int32_t *stats, void *images, int32_t *info) {
    if (!ops || !stats || !images || !info) return set_error(QCS_ERR_ARG, "null argument");
    if (nops < 1 || nops > QCS_MAX_OPS) return set_error(QCS_ERR_ARG, "nops must be in 1..QCS_MAX_OPS");
    return tile_plan_blocks(n, ops, nops, stats, images, info);
}

int qcs_apply_fused_bounded(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops, const double *norm_in,
                            double *norm_out, const double *bound2_in, void *ws, qcs_stream stream) {
    if (int e = check_state(dst, n, dtype)) return e;
    if (int e = check_state(src, n, dtype)) return e;
    if (!ops || nops < 1 || nops > QCS_MAX_OPS) return set_error(QCS_ERR_ARG, "op count must be in 1..QCS_MAX_OPS");
    if (norm_out && !ws) return set_error(QCS_ERR_ARG, "norm_out needs a workspace");
    for (int i = 0; i < nops; ++i)
        if (!ops[i].matrix) return set_error(QCS_ERR_ARG, "op without a matrix");
    return tile_apply(dst, src, n, dtype, ops, nops, norm_in, norm_out, bound2_in, ws, (cudaStream_t)stream);
}

int qcs_apply_fused(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops, const double *norm_in,
                    double *norm_out, void *ws, qcs_stream stream) {
    return qcs_apply_fused_bounded(dst, src, n, dtype, ops, nops, norm_in, norm_out, nullptr, ws, stream);
}

int qcs_apply_gate(void *dst, const void *src, int n, int dtype, const double *matrix, int k, const int32_t *bitpos,
                   const double *norm_in, double *norm_out, void *ws, qcs_stream stream) {
    if (!matrix || !bitpos) return set_error(QCS_ERR_ARG, "null matrix or bit positions");
    if (k < 1 || k > QCS_MAX_OP_QUBITS) return set_error(QCS_ERR_ARG, "qcs_apply_gate handles 1..4 operator qubits");
    qcs_op op;
    op.kind = QCS_OP_DENSE;
    op.k = k;
    for (int j = 0; j < QCS_MAX_OP_QUBITS; ++j) op.bitpos[j] = j < k ? bitpos[j] : 0;
    op.ctrl_mask = 0;
    op.matrix = matrix;
    return qcs_apply_fused(dst, src, n, dtype, &op, 1, norm_in, norm_out, ws, stream);
}

int qcs_apply_dense(void *dst, const void *src, int n, int dtype, const void *dev_matrix, int k, const int32_t *bitpos,
                    const double *norm_in, double *norm_out, void *ws, qcs_stream stream) {
    if (int e = check_state(dst, n, dtype)) return e;
    if (int e = check_state(src, n, dtype)) return e;
    if (!dev_matrix || !bitpos) return set_error(QCS_ERR_ARG, "null matrix or bit positions");
    if (norm_out && !ws) return set_error(QCS_ERR_ARG, "norm_out needs a workspace");
    return misc_dense(dst, src, n, dtype, dev_matrix, k, bitpos, norm_in, norm_out, ws, (cudaStream_t)stream);
}

int qcs_abs2(double *probs, const void *src, int n, int dtype, const double *norm_in, double *sum_out, void *ws,
             qcs_stream stream) {
    if (int e = check_state(src, n, dtype)) return e;
    if (sum_out && !ws) return set_error(QCS_ERR_ARG, "sum_out needs a workspace");
    return misc_abs2(probs, src, n, dtype, norm_in, sum_out, ws, (cudaStream_t)stream);
}

int qcs_dot(const void *a, const void *b, int n, int dtype, double *out2, void *ws, qcs_stream stream) {
    if (int e = check_state(a, n, dtype)) return e;
    if (int e = check_state(b, n, dtype)) return e;
    if (!out2 || !ws) return set_error(QCS_ERR_ARG, "null output or workspace");
    return misc_dot(a, b, n, dtype, out2, ws, (cudaStream_t)stream);
}

int qcs_scale(void *dst, const void *src, int n, int dtype, double re, double im, const double *norm_in,
              double *norm_out, void *ws, qcs_stream stream) {
    if (int e = check_state(dst, n, dtype)) return e;
    if (int e = check_state(src, n, dtype)) return e;
    const double alpha[2] = {re, im};
    return misc_lincomb(dst, src, nullptr, n, dtype, alpha, nullptr, norm_in, nullptr, norm_out, ws, (cudaStream_t)stream);
}

int qcs_lincomb(void *dst, const void *a, const void *b, int n, int dtype, const double *alpha2, const double *beta2,
                const double *norm_a, const double *norm_b, double *norm_out, void *ws, qcs_stream stream) {
    if (int e = check_state(dst, n, dtype)) return e;
    if (int e = check_state(a, n, dtype)) return e;
    if (int e = check_state(b, n, dtype)) return e;
    if (!alpha2 || !beta2) return set_error(QCS_ERR_ARG, "null coefficient");
    return misc_lincomb(dst, a, b, n, dtype, alpha2, beta2, norm_a, norm_b, norm_out, ws, (cudaStream_t)stream);
}

int qcs_kron(void *dst, const void *a, int na, const void *b, int nb, int dtype, const double *norm_a,
             const double *norm_b, double *norm_out, void *ws, qcs_stream stream) {
    if (int e = check_state(a, na, dtype)) return e;
    if (int e = check_state(b, nb, dtype)) return e;
    if (int e = check_state(dst, na + nb, dtype)) return e;
    return misc_kron(dst, a, na, b, nb, dtype, norm_a, norm_b, norm_out, ws, (cudaStream_t)stream);
}

int qcs_permute_bits(void *dst, const void *src, int n, int dtype, const int32_t *src_bit, qcs_stream stream) {
    if (int e = check_state(dst, n, dtype)) return e;
    if (int e = check_state(src, n, dtype)) return e;
    if (!src_bit) return set_error(QCS_ERR_ARG, "null permutation");
    if (dst == src) return set_error(QCS_ERR_ARG, "permute_bits cannot run in place");
    return misc_permute(dst, src, n, dtype, src_bit, false, (cudaStream_t)stream);
}

int qcs_conj_permute_bits(void *dst, const void *src, int n, int dtype, const int32_t *src_bit, qcs_stream stream) {
    if (int e = check_state(dst, n, dtype)) return e;
    if (int e = check_state(src, n, dtype)) return e;
    if (!src_bit) return set_error(QCS_ERR_ARG, "null permutation");
    if (dst == src) return set_error(QCS_ERR_ARG, "conj_permute_bits cannot run in place");
    return misc_permute(dst, src, n, dtype, src_bit, true, (cudaStream_t)stream);
}

int qcs_apply_uf(void *dst, const void *src, int n, int dtype, const uint32_t *dev_table, const int32_t *xbits, int nx,
                 int tbit, const double *norm_in, qcs_stream stream) {
    if (int e = check_state(dst, n, dtype)) return e;
    if (int e = check_state(src, n, dtype)) return e;
    if (!dev_table || !xbits) return set_error(QCS_ERR_ARG, "null argument");
    return misc_apply_uf(dst, src, n, dtype, dev_table, xbits, nx, tbit, norm_in, (cudaStream_t)stream);
}

int qcs_set_basis(void *buf, int n, int dtype, uint64_t index, qcs_stream stream) {
    if (int e = check_state(buf, n, dtype)) return e;
    return misc_set_basis(buf, n, dtype, index, (cudaStream_t)stream);
}

int qcs_marginal_probs(double *probs_out, const void *src, int n, int dtype, const int32_t *bitpos, int m,
                       const double *norm_in, void *ws, int64_t ws_bytes, qcs_stream stream) {
    if (!probs_out || !src || !bitpos || !ws) return set_error(QCS_ERR_ARG, "null pointer");
    if (n < 1 || n > 40) return set_error(QCS_ERR_ARG, "qubit count must be in 1..40");
    return misc_marginal(probs_out, src, n, dtype, bitpos, m, norm_in, ws, ws_bytes, (cudaStream_t)stream);
}

int qcs_collapse(void *dst, const void *src, int n, int dtype, const int32_t *bitpos, int m, uint64_t outcome,
                 int remove, double scale, double *norm_out, void *ws, qcs_stream stream) {
    if (int e = check_state(src, n, dtype)) return e;
    if (!dst || !bitpos) return set_error(QCS_ERR_ARG, "null pointer");
    if (norm_out && !ws) return set_error(QCS_ERR_ARG, "norm_out needs a workspace");
    return misc_collapse(dst, src, n, dtype, bitpos, m, outcome, remove, scale, norm_out, ws, (cudaStream_t)stream);
}

int qcs_diag_probs(double *probs, const void *rho, int n, int dtype, double *sum_out, void *ws, qcs_stream stream) {
    if (!probs || !rho) return set_error(QCS_ERR_ARG, "null pointer");
    if (dtype != QCS_C64 && dtype != QCS_C128) return set_error(QCS_ERR_ARG, "bad dtype");
    return misc_diag_probs(probs, rho, n, dtype, sum_out, ws, (cudaStream_t)stream);
}

int qcs_partial_trace(void *out, const void *rho, int d, int dtype, const int32_t *keep, int nkeep, qcs_stream stream) {
    if (!out || !rho || !keep) return set_error(QCS_ERR_ARG, "null argument");
    if (dtype != QCS_C64 && dtype != QCS_C128) return set_error(QCS_ERR_ARG, "dtype must be QCS_C64 or QCS_C128");
    return misc_partial_trace(out, rho, d, dtype, keep, nkeep, (cudaStream_t)stream);
}

int qcs_outer_accumulate(void *rho, const void *psi, int n, int dtype, double p, const double *norm_psi,
                         int accumulate, qcs_stream stream) {
    if (!rho || !psi) return set_error(QCS_ERR_ARG, "null pointer");
    if (dtype != QCS_C64 && dtype != QCS_C128) return set_error(QCS_ERR_ARG, "bad dtype");
    return misc_outer(rho, psi, n, dtype, p, norm_psi, accumulate, (cudaStream_t)stream);
}

// ---- multi-GPU exchange (qcs_dist.cu) ------------------------------------------------------------------------------
int qcs_dist_init(int ndev, const int *devs, qcs_dist_t *handle) {
    return dist_init(ndev, devs, reinterpret_cast<DistCtx **>(handle));
}
int qcs_dist_swap(qcs_dist_t handle, void *const *dst_shards, const void *const *src_shards, int n_local, int dtype,
                  const int32_t *global_bits, const int32_t *local_bits, int g, const qcs_stream *streams) {
    if (!global_bits || !local_bits) return set_error(QCS_ERR_ARG, "null bit list");
    return dist_swap(reinterpret_cast<DistCtx *>(handle), dst_shards, src_shards, n_local, dtype, global_bits, local_bits, g,
                     reinterpret_cast<const cudaStream_t *>(streams));
}
int qcs_dist_destroy(qcs_dist_t handle) { return dist_destroy(reinterpret_cast<DistCtx *>(handle)); }
int qcs_dist_swap_rank(void *local, const void *const *peers, int nranks, int rank, int n_local, int dtype,
                       const int32_t *global_bits, const int32_t *local_bits, int g, int push, qcs_stream stream) {
    if (!local || !peers || !global_bits || !local_bits) return set_error(QCS_ERR_ARG, "null argument");
    return dist_swap_rank(local, peers, nranks, rank, n_local, dtype, global_bits, local_bits, g, push, (cudaStream_t)stream);
}
int qcs_dist_alloc(void **ptr, int64_t bytes) { return dist_alloc(ptr, bytes); }
int qcs_dist_free(void *ptr) { return dist_free(ptr); }
int qcs_ipc_get_handle(const void *dev_ptr, void *handle64) {
    if (!dev_ptr || !handle64) return set_error(QCS_ERR_ARG, "null argument");
    return ipc_get_handle(dev_ptr, handle64);
}
int qcs_ipc_open(const void *handle64, void **mapped) {
    if (!handle64 || !mapped) return set_error(QCS_ERR_ARG, "null argument");
    return ipc_open(handle64, mapped);
}
int qcs_ipc_close(void *mapped) { return ipc_close(mapped); }

}  // extern "C"
