// Tile engine: k-qubit gate application and fused gate runs as ONE pass over the state.
//
// Replaces  np.tensordot(self._t, arg_t, ...) + np.array copy + renormalize_
//           (reference operators.py:259-276, state.py:89-97).
//
// Data layout in HBM: flat array of 2^n interleaved complex amplitudes (float2 / double2).
// A pass picks a set of "tile bits": the low L flat bits (a contiguous 2^L-amplitude chunk, 8 KB)
// plus H arbitrary high bits.  One tile = 2^H chunks = every amplitude that differs only in tile
// bits, so any gate whose targets are tile bits is local to a tile.  A CTA stages a tile in shared
// memory (coalesced 16-byte loads, or TMA bulk copies in the pipelined variant), applies the whole
// op list in shared memory / registers, accumulates the squared norm, and streams the tile out:
// each amplitude is read from HBM once and written once per pass however many gates are fused.
// Diagonal ops and control bits may address ANY bit (their outer bits are tile-uniform).
#include "qcs_common.cuh"
#include "qcs_internal.h"

namespace qcs {

// ------------------------------------------------------------------------------------------------
// op application on a tile held in shared memory
// ------------------------------------------------------------------------------------------------
template <typename R, int K>
__device__ __forceinline__ void apply_dense(typename CxT<R>::T *tile, int m, const TileOp &op,
                                            const typename CxT<R>::T *__restrict__ M, R scale,
                                            int tid, int nt) {
    using C = typename CxT<R>::T;
    constexpr int D = 1 << K;
    int sp[K];  // target positions, ascending
#pragma unroll
    for (int j = 0; j < K; ++j) sp[j] = op.lpos[j];
#pragma unroll
    for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = a + 1; b < K; ++b)
            if (sp[b] < sp[a]) { int t = sp[a]; sp[a] = sp[b]; sp[b] = t; }
    uint32_t off[D];
#pragma unroll
    for (int x = 0; x < D; ++x) {
        uint32_t o = 0;
#pragma unroll
        for (int j = 0; j < K; ++j)
            if ((x >> (K - 1 - j)) & 1) o |= 1u << op.lpos[j];
        off[x] = o;
    }
    const bool scaled = (op.flags & TILEOP_SCALE) != 0;
    const uint32_t ctrl = op.ctrl_local;
    const uint32_t ngroups = 1u << (m - K);
    // small matrices live in registers for the whole sweep; larger ones are broadcast-read from smem
    constexpr bool MREG = (K == 1);
    constexpr int DR = MREG ? D * D : 1;
    C Mr[DR];
    if (MREG) {
#pragma unroll
        for (int e = 0; e < DR; ++e) {
            Mr[e] = M[e];
            if (scaled) { Mr[e].x *= scale; Mr[e].y *= scale; }
        }
    }
    for (uint32_t g = tid; g < ngroups; g += nt) {
        uint32_t base = g;
#pragma unroll
        for (int j = 0; j < K; ++j) base = insert_zero32(base, sp[j]);
        if ((base & ctrl) != ctrl) continue;
        C a[D];
#pragma unroll
        for (int x = 0; x < D; ++x) a[x] = tile[base + off[x]];
#pragma unroll
        for (int y = 0; y < D; ++y) {
            C acc = cx<R>(R(0), R(0));
            if (MREG) {
#pragma unroll
                for (int x = 0; x < D; ++x) cfma(acc, Mr[(y * D + x) % DR], a[x]);
            } else {
#pragma unroll
                for (int x = 0; x < D; ++x) cfma(acc, M[y * D + x], a[x]);
                if (scaled) { acc.x *= scale; acc.y *= scale; }
            }
            tile[base + off[y]] = acc;
        }
    }
}

// A run of consecutive diagonal ops [i0, i1) is applied in one sweep over the tile.
template <typename R>
__device__ __forceinline__ void apply_diag_run(typename CxT<R>::T *tile, int m, const TileOp *ops,
                                               int i0, int i1, const typename CxT<R>::T *__restrict__ pool,
                                               uint64_t tile_base, R scale, int tid, int nt) {
    using C = typename CxT<R>::T;
    const uint32_t namp = 1u << m;
    for (uint32_t l = tid; l < namp; l += nt) {
        C a = tile[l];
        for (int i = i0; i < i1; ++i) {
            const TileOp &op = ops[i];
            if ((l & op.ctrl_local) != op.ctrl_local) continue;
            uint32_t x = 0;
            for (int j = 0; j < op.k; ++j) {
                const int lp = op.lpos[j];
                const uint32_t bit = lp >= 0 ? ((l >> lp) & 1u) : (uint32_t)((tile_base >> op.gpos[j]) & 1ull);
                x = (x << 1) | bit;
            }
            a = cmul(pool[op.mat + x], a);
            if (op.flags & TILEOP_SCALE) { a.x *= scale; a.y *= scale; }
        }
        tile[l] = a;
    }
}

template <typename R, int KMAX>
__device__ __forceinline__ void apply_ops(typename CxT<R>::T *tile, const TileOp *ops, int nops,
                                          const typename CxT<R>::T *pool, int m, uint64_t tile_base,
                                          R scale, int tid, int nt) {
    using C = typename CxT<R>::T;
    int i = 0;
    while (i < nops) {
        const TileOp &op = ops[i];
        if ((tile_base & op.ctrl_outer) != op.ctrl_outer) { ++i; continue; }  // tile-uniform
        if (op.kind == QCS_OP_DIAG) {
            int j = i + 1;
            while (j < nops && ops[j].kind == QCS_OP_DIAG &&
                   (tile_base & ops[j].ctrl_outer) == ops[j].ctrl_outer) ++j;
            apply_diag_run<R>(tile, m, ops, i, j, pool, tile_base, scale, tid, nt);
            i = j;
        } else {
            const C *M = pool + op.mat;
            if (KMAX <= 2) {
                if (op.k == 1) apply_dense<R, 1>(tile, m, op, M, scale, tid, nt);
                else apply_dense<R, 2>(tile, m, op, M, scale, tid, nt);
            } else {
                switch (op.k) {
                    case 1: apply_dense<R, 1>(tile, m, op, M, scale, tid, nt); break;
                    case 2: apply_dense<R, 2>(tile, m, op, M, scale, tid, nt); break;
                    case 3: apply_dense<R, 3>(tile, m, op, M, scale, tid, nt); break;
                    default: apply_dense<R, 4>(tile, m, op, M, scale, tid, nt); break;
                }
            }
            ++i;
        }
        __syncthreads();
    }
}

__host__ __device__ __forceinline__ uint32_t align128(uint32_t x) { return (x + 127u) & ~127u; }

// copy the op list and matrix pool from kernel-parameter space into shared memory (once per CTA)
template <typename Prog>
__device__ __forceinline__ uint32_t stage_program(unsigned char *smem, const Prog &P, int tid, int nt) {
    const uint32_t ops_words = (uint32_t)(P.nops * sizeof(TileOp) / 4);
    const uint32_t ops_bytes = align128(ops_words * 4);
    const uint32_t *po = reinterpret_cast<const uint32_t *>(P.ops);
    uint32_t *so = reinterpret_cast<uint32_t *>(smem);
    for (uint32_t i = tid; i < ops_words; i += nt) so[i] = po[i];
    const uint32_t mat_words = (uint32_t)P.mat_bytes / 4;
    const uint32_t *pm = reinterpret_cast<const uint32_t *>(P.mats);
    uint32_t *sm = reinterpret_cast<uint32_t *>(smem + ops_bytes);
    for (uint32_t i = tid; i < mat_words; i += nt) sm[i] = pm[i];
    return ops_bytes;
}
template <typename Prog>
static uint32_t program_smem_bytes(const Prog &P) {
    return align128((uint32_t)(P.nops * sizeof(TileOp))) + align128((uint32_t)P.mat_bytes);
}

template <typename Prog>
__device__ __forceinline__ uint64_t tile_base_of(uint64_t t, const Prog &P) {
    uint64_t x = t;
    for (int j = 0; j < P.H; ++j) x = insert_zero(x, P.hbits[j] - P.L);
    return x << P.L;
}
template <typename Prog>
__device__ __forceinline__ uint64_t chunk_offset(uint32_t c, const Prog &P) {
    uint64_t o = 0;
    for (int j = 0; j < P.H; ++j)
        if ((c >> j) & 1u) o |= 1ull << P.hbits[j];
    return o;
}

template <typename R> __device__ __forceinline__ double vec_abs2(const uint4 v);
template <> __device__ __forceinline__ double vec_abs2<float>(const uint4 v) {
    const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z),
                d = __uint_as_float(v.w);
    return (double)(fmaf(a, a, b * b) + fmaf(c, c, d * d));
}
template <> __device__ __forceinline__ double vec_abs2<double>(const uint4 v) {
    const double a = __hiloint2double(v.y, v.x), b = __hiloint2double(v.w, v.z);
    return a * a + b * b;
}

// ------------------------------------------------------------------------------------------------
// Variant 1: cooperative loads/stores, one tile per CTA iteration, several CTAs per SM
// ------------------------------------------------------------------------------------------------
template <typename R, typename Prog, int KMAX>
__global__ void __launch_bounds__(TILE_THREADS, KMAX <= 2 ? 4 : 2) tile_kernel(const __grid_constant__ Prog P) {
    using C = typename CxT<R>::T;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const uint32_t ops_bytes = stage_program(smem_raw, P, tid, nt);
    const TileOp *s_ops = reinterpret_cast<const TileOp *>(smem_raw);
    const C *s_pool = reinterpret_cast<const C *>(smem_raw + ops_bytes);
    unsigned char *tile_raw = smem_raw + ops_bytes + align128((uint32_t)P.mat_bytes);
    C *tile = reinterpret_cast<C *>(tile_raw);
    uint4 *tile_v = reinterpret_cast<uint4 *>(tile_raw);
    const int m = P.L + P.H;
    constexpr int VSHIFT = sizeof(C) == 8 ? 1 : 0;       // amplitudes per 16-byte vector = 1 << VSHIFT
    const uint32_t tile_vecs = (1u << m) >> VSHIFT;
    const int chunk_vec_bits = P.L - VSHIFT;             // log2(vectors per chunk)
    const R scale = (R)load_inv_norm(P.norm_in);
    const uint4 *src = reinterpret_cast<const uint4 *>(P.src);
    uint4 *dst = reinterpret_cast<uint4 *>(P.dst);
    double nacc[1] = {0.0};

    for (uint64_t t = blockIdx.x; t < P.ntiles; t += gridDim.x) {
        const uint64_t base = tile_base_of(t, P);
        // stage in
        for (uint32_t v = tid; v < tile_vecs; v += nt) {
            const uint32_t c = v >> chunk_vec_bits;
            const uint64_t g = ((base + chunk_offset(c, P)) >> VSHIFT) + (v & ((1u << chunk_vec_bits) - 1u));
            tile_v[v] = src[g];
        }
        __syncthreads();
        apply_ops<R, KMAX>(tile, s_ops, P.nops, s_pool, m, base, scale, tid, nt);
        // stage out (+ squared norm)
        for (uint32_t v = tid; v < tile_vecs; v += nt) {
            const uint32_t c = v >> chunk_vec_bits;
            const uint64_t g = ((base + chunk_offset(c, P)) >> VSHIFT) + (v & ((1u << chunk_vec_bits) - 1u));
            const uint4 val = tile_v[v];
            nacc[0] += vec_abs2<R>(val);
            dst[g] = val;
        }
        __syncthreads();
    }
    if (P.norm_out) grid_sum_publish<1>(nacc, P.ws, P.norm_out);
}

// ------------------------------------------------------------------------------------------------
// Variant 2: TMA bulk-copy pipeline (cp.async.bulk + mbarrier), S stages per CTA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename R, typename Prog, int KMAX>
__global__ void __launch_bounds__(TILE_THREADS_BULK, KMAX <= 2 ? 2 : 1) tile_kernel_bulk(const __grid_constant__ Prog P) {
    using C = typename CxT<R>::T;
    extern __shared__ __align__(128) unsigned char smem_all[];
    __shared__ __align__(8) uint64_t full_bar[TILE_MAX_STAGES];

    const int tid = threadIdx.x, nt = blockDim.x;
    const uint32_t ops_bytes = stage_program(smem_all, P, tid, nt);
    const TileOp *s_ops = reinterpret_cast<const TileOp *>(smem_all);
    const C *s_pool = reinterpret_cast<const C *>(smem_all + ops_bytes);
    unsigned char *smem_raw = smem_all + ops_bytes + align128((uint32_t)P.mat_bytes);
    const int m = P.L + P.H;
    const int S = P.stages;
    const uint32_t tile_bytes = (uint32_t)((sizeof(C)) << m);
    const uint32_t chunk_bytes = (uint32_t)((sizeof(C)) << P.L);
    const uint32_t nchunks = 1u << P.H;
    const R scale = (R)load_inv_norm(P.norm_in);
    const unsigned char *src = reinterpret_cast<const unsigned char *>(P.src);
    unsigned char *dst = reinterpret_cast<unsigned char *>(P.dst);
    double nacc[1] = {0.0};

    // tiles owned by this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint64_t my_tiles = (P.ntiles > blockIdx.x) ? (P.ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_load = [&](uint64_t it) {
        const int st = (int)(it % S);
        const uint64_t base = tile_base_of(blockIdx.x + it * gridDim.x, P);
        unsigned char *sdst = smem_raw + (size_t)st * tile_bytes;
        mbar_expect_tx(&full_bar[st], tile_bytes);
        for (uint32_t c = 0; c < nchunks; ++c)
            bulk_g2s(sdst + (size_t)c * chunk_bytes, src + (base + chunk_offset(c, P)) * sizeof(C), chunk_bytes,
                     &full_bar[st]);
    };

    if (tid == 0) {
        const uint64_t pre = my_tiles < (uint64_t)(S - 1) ? my_tiles : (uint64_t)(S - 1);
        for (uint64_t it = 0; it < pre; ++it) issue_load(it);
    }

    for (uint64_t it = 0; it < my_tiles; ++it) {
        const int st = (int)(it % S);
        const uint32_t parity = (uint32_t)((it / S) & 1);
        const uint64_t base = tile_base_of(blockIdx.x + it * gridDim.x, P);
        C *tile = reinterpret_cast<C *>(smem_raw + (size_t)st * tile_bytes);
        mbar_wait(&full_bar[st], parity);
        apply_ops<R, KMAX>(tile, s_ops, P.nops, s_pool, m, base, scale, tid, nt);
        if (P.norm_out) {
            const uint4 *tv = reinterpret_cast<const uint4 *>(tile);
            const uint32_t tile_vecs = tile_bytes >> 4;
            for (uint32_t v = tid; v < tile_vecs; v += nt) nacc[0] += vec_abs2<R>(tv[v]);
        }
        fence_proxy_async();   // generic-proxy writes of this stage -> visible to the bulk store
        __syncthreads();
        if (tid == 0) {
            for (uint32_t c = 0; c < nchunks; ++c)
                bulk_s2g(dst + (base + chunk_offset(c, P)) * sizeof(C),
                         reinterpret_cast<unsigned char *>(tile) + (size_t)c * chunk_bytes, chunk_bytes);
            bulk_commit();
            const uint64_t nxt = it + (uint64_t)(S - 1);
            if (nxt < my_tiles) {
                bulk_wait_read<1>();   // the store that last read stage (it-1)%S has drained it
                issue_load(nxt);
            }
        }
    }
    if (tid == 0) bulk_wait_all<0>();
    if (P.norm_out) grid_sum_publish<1>(nacc, P.ws, P.norm_out);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <typename R, typename Prog, int KMAX>
static int launch_tile(const Prog &P, const DeviceInfo &dev, int mode, cudaStream_t stream) {
    using C = typename CxT<R>::T;
    const int m = P.L + P.H;
    const size_t tile_bytes = sizeof(C) << m;
    if (mode == TILE_MODE_BULK && tile_bytes >= 16 && (sizeof(C) << P.L) >= 16) {
        auto kern = tile_kernel_bulk<R, Prog, KMAX>;
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BUDGET);
            attr_done = true;
        }
        const size_t smem = (size_t)P.stages * tile_bytes + program_smem_bytes(P);
        int per_sm = (int)(TILE_SMEM_BUDGET / (smem + 1024));
        if (per_sm > (KMAX <= 2 ? 2 : 1)) per_sm = (KMAX <= 2 ? 2 : 1);
        if (per_sm < 1) per_sm = 1;
        uint64_t grid = (uint64_t)dev.sms * per_sm;
        if (grid > P.ntiles) grid = P.ntiles;
        if (grid > MAX_GRID) grid = MAX_GRID;
        int threads = TILE_THREADS_BULK;
        while (threads > 32 && (1u << m) < (unsigned)threads) threads >>= 1;
        kern<<<(unsigned)grid, threads, smem, stream>>>(P);
    } else {
        auto kern = tile_kernel<R, Prog, KMAX>;
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BUDGET);
            attr_done = true;
        }
        const size_t smem = (tile_bytes < 16 ? 16 : tile_bytes) + program_smem_bytes(P);
        int per_sm = (int)(TILE_SMEM_BUDGET / (smem + 1024));
        if (per_sm > (KMAX <= 2 ? 4 : 2)) per_sm = (KMAX <= 2 ? 4 : 2);
        if (per_sm < 1) per_sm = 1;
        uint64_t grid = (uint64_t)dev.sms * per_sm;
        if (grid > P.ntiles) grid = P.ntiles;
        if (grid > MAX_GRID) grid = MAX_GRID;
        int threads = TILE_THREADS;
        while (threads > 32 && (1u << m) < (unsigned)threads) threads >>= 1;
        kern<<<(unsigned)grid, threads, smem, stream>>>(P);
    }
    return cudaGetLastError() == cudaSuccess ? QCS_OK : QCS_ERR_CUDA;
}

template <typename Prog>
static int build_and_launch(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops,
                            const double *norm_in, double *norm_out, void *ws, cudaStream_t stream) {
    const bool c64 = dtype == QCS_C64;
    const int vshift = c64 ? 1 : 0;
    Prog *Pp = new Prog();
    Prog &P = *Pp;
    struct Guard { Prog *p; ~Guard() { delete p; } } guard{Pp};

    // ---- tile geometry: low L bits + the dense targets above them --------------------------------
    uint64_t dense_mask = 0;
    for (int i = 0; i < nops; ++i) {
        const qcs_op &o = ops[i];
        if (o.k < 1 || o.k > QCS_MAX_OP_QUBITS) return set_error(QCS_ERR_ARG, "op with k outside 1..4");
        uint64_t seen = 0;
        for (int j = 0; j < o.k; ++j) {
            if (o.bitpos[j] < 0 || o.bitpos[j] >= n) return set_error(QCS_ERR_ARG, "op bit position out of range");
            if (seen & (1ull << o.bitpos[j])) return set_error(QCS_ERR_ARG, "repeated bit position in op");
            seen |= 1ull << o.bitpos[j];
        }
        if (o.ctrl_mask & seen) return set_error(QCS_ERR_ARG, "control bit is also a target");
        if (n < 64 && (o.ctrl_mask >> n)) return set_error(QCS_ERR_ARG, "control bit out of range");
        if (o.kind == QCS_OP_DENSE) dense_mask |= seen;
        else if (o.kind != QCS_OP_DIAG) return set_error(QCS_ERR_ARG, "unknown op kind");
    }
    const int lpref = c64 ? TILE_LOW_BITS_C64 : TILE_LOW_BITS_C128;
    const int mmax = c64 ? TILE_MAX_BITS_C64 : TILE_MAX_BITS_C128;
    int L = n < lpref ? n : lpref;
    int H = 0;
    for (;; --L) {
        H = __builtin_popcountll(dense_mask >> L);
        if (L + H <= mmax) break;
        if (L <= vshift) return set_error(QCS_ERR_UNSUPPORTED, "fused ops touch too many high bits for one tile");
    }
    if (H > TILE_MAX_H) return set_error(QCS_ERR_UNSUPPORTED, "too many high tile bits");
    P.n = n; P.L = L; P.H = H; P.nops = nops;
    {
        int h = 0;
        for (int b = L; b < n; ++b)
            if ((dense_mask >> b) & 1ull) P.hbits[h++] = b;
    }
    const int m = L + H;
    P.ntiles = 1ull << (n - m);
    P.src = src; P.dst = dst; P.norm_in = norm_in; P.norm_out = norm_out; P.ws = ws;
    uint64_t tile_mask = (L >= 64 ? ~0ull : ((1ull << L) - 1ull));
    for (int h = 0; h < H; ++h) tile_mask |= 1ull << P.hbits[h];
    auto local_of = [&](int b) -> int {
        if (b < L) return b;
        for (int h = 0; h < H; ++h) if (P.hbits[h] == b) return L + h;
        return -1;
    };

    // ---- ops + matrix pool ------------------------------------------------------------------------
    const size_t csize = c64 ? 8 : 16;
    const size_t pool_cap = sizeof(P.mats) / csize;
    size_t used = 0;
    bool scale_placed = (norm_in == nullptr);
    for (int i = 0; i < nops; ++i) {
        const qcs_op &o = ops[i];
        TileOp &t = P.ops[i];
        t.kind = o.kind; t.k = o.k; t.flags = 0;
        const size_t entries = o.kind == QCS_OP_DENSE ? (size_t)1 << (2 * o.k) : (size_t)1 << o.k;
        if (used + entries > pool_cap) return set_error(QCS_ERR_UNSUPPORTED, "fused pass matrix pool exhausted");
        t.mat = (int32_t)used;
        if (c64) {
            float *p = reinterpret_cast<float *>(P.mats) + 2 * used;
            for (size_t e = 0; e < 2 * entries; ++e) p[e] = (float)o.matrix[e];
        } else {
            double *p = reinterpret_cast<double *>(P.mats) + 2 * used;
            for (size_t e = 0; e < 2 * entries; ++e) p[e] = o.matrix[e];
        }
        used += entries;
        for (int j = 0; j < QCS_MAX_OP_QUBITS; ++j) { t.lpos[j] = -1; t.gpos[j] = 0; }
        for (int j = 0; j < o.k; ++j) { t.lpos[j] = (int8_t)local_of(o.bitpos[j]); t.gpos[j] = (int8_t)o.bitpos[j]; }
        t.ctrl_local = 0;
        t.ctrl_outer = o.ctrl_mask & ~tile_mask;
        for (int b = 0; b < n; ++b)
            if (((o.ctrl_mask & tile_mask) >> b) & 1ull) t.ctrl_local |= 1u << local_of(b);
        if (!scale_placed && o.ctrl_mask == 0) { t.flags |= TILEOP_SCALE; scale_placed = true; }
    }
    if (!scale_placed) {
        // every op is controlled: append an uncontrolled identity diagonal that carries the scale
        if (nops >= (int)(sizeof(P.ops) / sizeof(P.ops[0])) || used + 2 > pool_cap)
            return set_error(QCS_ERR_UNSUPPORTED, "no room for the scale op");
        TileOp &t = P.ops[nops];
        t.kind = QCS_OP_DIAG; t.k = 1; t.flags = TILEOP_SCALE; t.mat = (int32_t)used;
        for (int j = 0; j < QCS_MAX_OP_QUBITS; ++j) { t.lpos[j] = -1; t.gpos[j] = 0; }
        t.lpos[0] = 0; t.gpos[0] = 0; t.ctrl_local = 0; t.ctrl_outer = 0;
        if (c64) { float *p = reinterpret_cast<float *>(P.mats) + 2 * used; p[0] = 1; p[1] = 0; p[2] = 1; p[3] = 0; }
        else { double *p = reinterpret_cast<double *>(P.mats) + 2 * used; p[0] = 1; p[1] = 0; p[2] = 1; p[3] = 0; }
        P.nops = nops + 1;
    }

    P.mat_bytes = (int32_t)((used + 2) * csize);
    if (P.mat_bytes > (int32_t)sizeof(P.mats)) P.mat_bytes = (int32_t)sizeof(P.mats);
    const DeviceInfo &dev = device_info();
    const int mode = tile_mode();
    const size_t tile_bytes = csize << m;
    int S = (int)(TILE_STAGE_BUDGET / (tile_bytes ? tile_bytes : 1));
    if (S > TILE_MAX_STAGES) S = TILE_MAX_STAGES;
    if (S < 2) S = 2;
    P.stages = S;
    int kmax = 1;
    for (int i = 0; i < nops; ++i)
        if (ops[i].kind == QCS_OP_DENSE && ops[i].k > kmax) kmax = ops[i].k;
    if (kmax <= 2)
        return c64 ? launch_tile<float, Prog, 2>(P, dev, mode, stream) : launch_tile<double, Prog, 2>(P, dev, mode, stream);
    return c64 ? launch_tile<float, Prog, 4>(P, dev, mode, stream) : launch_tile<double, Prog, 4>(P, dev, mode, stream);
}

int tile_apply(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops, const double *norm_in,
               double *norm_out, void *ws, cudaStream_t stream) {
    if (nops < 1) return set_error(QCS_ERR_ARG, "empty op list");
    size_t entries = 0;
    for (int i = 0; i < nops; ++i)
        entries += ops[i].kind == QCS_OP_DENSE ? (size_t)1 << (2 * ops[i].k) : (size_t)1 << ops[i].k;
    const size_t csize = dtype == QCS_C64 ? 8 : 16;
    if (nops + 1 <= SmallProg::MAXOPS && (entries + 2) * csize <= sizeof(SmallProg::mats))
        return build_and_launch<SmallProg>(dst, src, n, dtype, ops, nops, norm_in, norm_out, ws, stream);
    if (nops + 1 > LargeProg::MAXOPS) return set_error(QCS_ERR_UNSUPPORTED, "too many ops in one fused pass");
    return build_and_launch<LargeProg>(dst, src, n, dtype, ops, nops, norm_in, norm_out, ws, stream);
}

}  // namespace qcs
