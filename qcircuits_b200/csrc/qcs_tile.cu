// Tile engine: k-qubit gate application and fused gate runs as ONE pass over the state.
//
// Replaces  np.tensordot(self._t, arg_t, ...) + np.array copy + renormalize_
//           (reference operators.py:259-276, state.py:89-97).
//
// Data layout in HBM: flat array of 2^n interleaved complex amplitudes (float2 / double2).
// A pass picks a set of "tile bits": the low L flat bits (a contiguous 2^L-amplitude chunk, 8 KB)
// plus H arbitrary high bits.  One tile = 2^H chunks = every amplitude that differs only in tile
// bits, so any gate whose targets are tile bits is local to a tile.  A CTA streams tiles through a
// multi-stage shared-memory ring (cp.async, or TMA bulk copies in the alternative variant), and for
// each tile runs the op list as a few *register sweeps*: a thread pulls the 2^r amplitudes spanned by
// r tile bits into registers, applies every gate of the sweep that acts on those bits (2x2 butterflies,
// controlled or not, and phase multiplications on any bits) and writes them back.  The squared norm is
// accumulated and the lazy 1/sqrt(norm) folded in while the tile streams out: each amplitude is read
// from HBM once and written once per pass however many gates are fused.
#include <cstdio>
#include <cmath>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "qcs_common.cuh"
#include "qcs_internal.h"

namespace qcs {

// ------------------------------------------------------------------------------------------------
// shared-memory layouts (indices in units of one amplitude / one 16-byte vector)
// ------------------------------------------------------------------------------------------------
// Padded: one 16-byte pad after every 128-byte row, one more after every 8 rows and one more after every 64 rows.
// Tile bit b then moves an amplitude by 8 * 2^b bytes plus pads whose sum is 16, 32 or 64 (mod 128) for EVERY b >= 4
// (complex64; b >= 3 for complex128), cycling with period three: any three consecutive tile bits that land on the lanes
// of a quarter warp address eight different 16-byte slots of a 128-byte bank row -- conflict-free 128-bit accesses
// whichever bits a sweep keeps in registers.  (With the single-level pad of round 1 every bit above 6 moved a lane by a
// multiple of 128 bytes: 23 % of the shared-memory wavefronts of a tensor-core pass were conflict replays.)
// All maps are additive over disjoint bit sets (no carries), which the sweeps and the staging loops use.
template <typename C> struct PaddedLayout {
    static constexpr int ROW = 128 / sizeof(C);                 // amplitudes per row
    static constexpr int PADA = 16 / sizeof(C);                 // pad, in amplitudes
    static __host__ __device__ __forceinline__ constexpr uint32_t amp(uint32_t i) {
        return i + ((i / ROW) + (i / (8 * ROW)) + (i / (64 * ROW))) * PADA;
    }
    static __host__ __device__ __forceinline__ constexpr uint32_t vec(uint32_t v) { return v + (v >> 3) + (v >> 6) + (v >> 9); }
    static __host__ __device__ __forceinline__ size_t bytes(int m) {
        const size_t raw = sizeof(C) << m;
        const size_t rows = (raw + 127) / 128;
        return raw + (rows + (rows + 7) / 8 + (rows + 63) / 64) * 16;
    }
};
template <typename C> struct LinearLayout {
    static __host__ __device__ __forceinline__ uint32_t amp(uint32_t i) { return i; }
    static __host__ __device__ __forceinline__ uint32_t vec(uint32_t v) { return v; }
    static __host__ __device__ __forceinline__ size_t bytes(int m) { return sizeof(C) << m; }
};

// ------------------------------------------------------------------------------------------------
// generic ops in shared memory (dense on 2..4 targets)
// ------------------------------------------------------------------------------------------------
template <typename R, int K, typename Layout>
__device__ __forceinline__ void apply_dense(typename CxT<R>::T *tile, int m, const TileOp &op,
                                            const typename CxT<R>::T *__restrict__ M, int tid, int nt) {
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
        off[x] = Layout::amp(o);
    }
    const uint32_t ctrl = op.ctrl_local;
    const uint32_t ngroups = 1u << (m - K);
    for (uint32_t g = tid; g < ngroups; g += nt) {
        uint32_t base = g;
#pragma unroll
        for (int j = 0; j < K; ++j) base = insert_zero32(base, sp[j]);
        if ((base & ctrl) != ctrl) continue;
        const uint32_t pb = Layout::amp(base);
        C a[D];
#pragma unroll
        for (int x = 0; x < D; ++x) a[x] = tile[pb + off[x]];
#pragma unroll
        for (int y = 0; y < D; ++y) {
            C acc = cx<R>(R(0), R(0));
#pragma unroll
            for (int x = 0; x < D; ++x) cfma(acc, M[y * D + x], a[x]);
            tile[pb + off[y]] = acc;
        }
    }
}

template <typename R, typename Layout>
__device__ __forceinline__ void apply_generic(typename CxT<R>::T *tile, int m, const TileOp &op,
                                              const typename CxT<R>::T *pool, uint64_t tile_base, int tid, int nt) {
    using C = typename CxT<R>::T;
    if ((tile_base & op.ctrl_outer) != op.ctrl_outer) return;   // tile-uniform control
    const C *M = pool + op.mat;
    switch (op.k) {
        case 1: apply_dense<R, 1, Layout>(tile, m, op, M, tid, nt); break;
        case 2: apply_dense<R, 2, Layout>(tile, m, op, M, tid, nt); break;
        case 3: apply_dense<R, 3, Layout>(tile, m, op, M, tid, nt); break;
        default: apply_dense<R, 4, Layout>(tile, m, op, M, tid, nt); break;
    }
}

// ------------------------------------------------------------------------------------------------
// register sweeps
// ------------------------------------------------------------------------------------------------
// A fused pass is instruction-issue-bound, so the op interpreter is written for the fewest instructions
// per op (gen_regops.py -> qcs_regops.inc, one inline-PTX block per amplitude type):
//  * dispatch is one `brx.idx` through a jump table (a C++ switch became a 6-level compare/branch tree and
//    the compiler copied all 2^RB loop-carried amplitudes between two register sets once per op);
//  * all arithmetic is an in-place real 2x2 map on a pair of registers; gates are classified on the host
//    into the cheapest exact form, and common scalar factors (1/sqrt2 of H, cos of RX) are deferred into the
//    pass-level factor that the final store applies anyway;
//  * every position that selects registers (target bit, control bit, phase bits) is baked into the arm.
// Instructions per thread (complex64, 16 amplitudes): HAD 32, REALU / IMAGU 48, REAL / IMAG 64, GEN 128, GEN2 256,
// PH1 32, SGN1 16, PH2 16, SGN2 8, CX 24 (moves), PHT 4; + ~9 per op for fetch and dispatch.  The loop is
// software-pipelined: the record of op i+1 and the coefficients of op i are in flight during the dispatch.
// Predicates that do not involve register bits (control bits elsewhere in the tile / outside the tile) are
// the "_P" entries: they substitute identity coefficients instead of branching around the body.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
#include "qcs_regops.inc"

// Compile-time geometry (MT tile bits, 2^(MT - RB) threads): exactly one group of 2^RB amplitudes per thread
// and sweep.  The zero insertions and the byte strides come precomputed in the Sweep record, so the 2^RB
// shared-memory addresses cost ~27 integer instructions (they were 80 when derived from the bit positions).
template <typename R, int RB, typename Layout>
__device__ __forceinline__ void reg_sweep_fast(typename CxT<R>::T *tile, const Sweep &sw, const RegOp *rops,
                                               const typename CxT<R>::T *pool, uint64_t outer_ok, int tid) {
    using C = typename CxT<R>::T;
    constexpr int NX = 1 << RB;
    // the record as one 8-byte broadcast load (Sweep::packed); masks and strides are recomputed from the register-bit
    // positions instead of being fetched with dependent shared-memory loads
    uint32_t plo, phi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(plo), "=r"(phi) : "r"(smem_u32(&sw.packed)) : "memory");
    const uint32_t flags = plo >> 24, first = phi;
    const bool VEC = sizeof(C) == 8 && (plo & 15u) == 0;     // complex64 pairs move as one 128-bit access
    uint32_t base = (uint32_t)tid;
#pragma unroll
    for (int j = 0; j < RB; ++j) base += base & (0xFFFFFFFFu << ((plo >> (4 * j)) & 15u));
    const uint32_t tile_u32 = smem_u32(tile) + Layout::amp(base) * (uint32_t)sizeof(C);
    uint32_t addr[NX];
    addr[0] = tile_u32;
#pragma unroll
    for (int j = 0; j < RB; ++j) {
        const uint32_t st = Layout::amp(1u << ((plo >> (4 * j)) & 15u)) * (uint32_t)sizeof(C);
#pragma unroll
        for (int x = 0; x < (1 << j); ++x) addr[x | (1 << j)] = addr[x] + st;
    }
    C a[NX];
    if (VEC) {
#pragma unroll
        for (int x = 0; x < NX; x += 2)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(reinterpret_cast<float &>(a[x].x)), "=f"(reinterpret_cast<float &>(a[x].y)),
                           "=f"(reinterpret_cast<float &>(a[x + 1].x)), "=f"(reinterpret_cast<float &>(a[x + 1].y))
                         : "r"(addr[x])
                         : "memory");
    } else {
#pragma unroll
        for (int x = 0; x < NX; ++x) a[x] = *reinterpret_cast<const C *>(__cvta_shared_to_generic(addr[x]));
    }
    C ph = cx<R>(R(1), R(0));     // product of the phases that are uniform over this thread's amplitudes
    if (flags & 2u) run_sweep_ops_dense2(a, ph, smem_u32(rops + first), smem_u32(pool), base, outer_ok);
    else run_sweep_ops(a, ph, smem_u32(rops + first), smem_u32(pool), base, outer_ok);
    if (flags & 4u) {
#pragma unroll
        for (int x = 0; x < NX; ++x) a[x] = cmul(ph, a[x]);
    }
    if (VEC) {
#pragma unroll
        for (int x = 0; x < NX; x += 2)
            asm volatile("st.shared.v4.f32 [%4], {%0, %1, %2, %3};" ::"f"(reinterpret_cast<float &>(a[x].x)),
                         "f"(reinterpret_cast<float &>(a[x].y)), "f"(reinterpret_cast<float &>(a[x + 1].x)),
                         "f"(reinterpret_cast<float &>(a[x + 1].y)), "r"(addr[x])
                         : "memory");
    } else {
#pragma unroll
        for (int x = 0; x < NX; ++x) *reinterpret_cast<C *>(__cvta_shared_to_generic(addr[x])) = a[x];
    }
}

// MT > 0: the tile has exactly MT bits, every sweep uses RB register bits (compile-time trip counts).
template <typename R, int RB, typename Layout, int MT>
__device__ __forceinline__ void reg_sweep(typename CxT<R>::T *tile, int m_rt, const Sweep &sw, const RegOp *rops,
                                          const typename CxT<R>::T *pool, uint64_t outer_ok, int tid, int nt) {
    using C = typename CxT<R>::T;
    constexpr int NX = 1 << RB;
    constexpr bool FAST = MT > 0;
    // complex64 with tile bit 0 among the register bits: amplitude pairs move as one 128-bit access
    const bool VEC = sizeof(C) == 8 && sw.nreg >= 1 && sw.rpos[0] == 0;
    const int m = FAST ? MT : m_rt;
    const int nreg = FAST ? RB : sw.nreg;
    const uint32_t nx_used = 1u << nreg;
    uint32_t ps[RB];      // shared-memory stride of each register bit (layout maps are additive)
    int rp[RB];
#pragma unroll
    for (int j = 0; j < RB; ++j) {
        rp[j] = sw.rpos[j];
        ps[j] = j < nreg ? Layout::amp(1u << rp[j]) : 0u;
    }
    const uint32_t ngroups = 1u << (m - nreg);
    const uint32_t ops_smem = smem_u32(rops + sw.first), pool_smem = smem_u32(pool);
#pragma unroll
    for (uint32_t g = tid; g < ngroups; g += nt) {
        uint32_t base = g;
#pragma unroll
        for (int j = 0; j < RB; ++j)
            if (j < nreg) base = insert_zero32(base, rp[j]);
        const uint32_t pb = Layout::amp(base);
        uint32_t addr[NX];
#pragma unroll
        for (uint32_t x = 0; x < NX; ++x) {
            uint32_t o = pb;
#pragma unroll
            for (int j = 0; j < RB; ++j)
                if (x & (1u << j)) o += ps[j];
            addr[x] = o;
        }
        C a[NX];
        if (!FAST) {
#pragma unroll
            for (uint32_t x = 0; x < NX; ++x) a[x] = cx<R>(R(0), R(0));     // slots beyond 2^nreg stay defined
        }
        if (VEC) {
#pragma unroll
            for (uint32_t x = 0; x < NX; x += 2) {
                if (FAST || x < nx_used) {
                    const float4 v = *reinterpret_cast<const float4 *>(&tile[addr[x]]);
                    a[x].x = v.x; a[x].y = v.y; a[x + 1].x = v.z; a[x + 1].y = v.w;
                }
            }
        } else {
#pragma unroll
            for (uint32_t x = 0; x < NX; ++x)
                if (FAST || x < nx_used) a[x] = tile[addr[x]];
        }
        C ph = cx<R>(R(1), R(0));     // product of the phases that are uniform over this thread's amplitudes
        if (sw.has_dense2) run_sweep_ops_dense2(a, ph, ops_smem, pool_smem, base, outer_ok);
        else run_sweep_ops(a, ph, ops_smem, pool_smem, base, outer_ok);
        if (sw.thread_phase) {
#pragma unroll
            for (uint32_t x = 0; x < NX; ++x) a[x] = cmul(ph, a[x]);
        }
        if (VEC) {
#pragma unroll
            for (uint32_t x = 0; x < NX; x += 2) {
                if (FAST || x < nx_used) {
                    float4 v;
                    v.x = a[x].x; v.y = a[x].y; v.z = a[x + 1].x; v.w = a[x + 1].y;
                    *reinterpret_cast<float4 *>(&tile[addr[x]]) = v;
                }
            }
        } else {
#pragma unroll
            for (uint32_t x = 0; x < NX; ++x)
                if (FAST || x < nx_used) tile[addr[x]] = a[x];
        }
    }
}

// The program as staged in shared memory.
struct ProgView {
    const Sweep *sweeps;
    const RegOp *regops;
    const OuterPred *outer;
    const TileOp *gen;
    const void *pool;
    uint64_t *outer_ok;     // [4]: per-tile bitmask of the outer predicates, double-buffered by tile parity (x 2 groups)
    int nsweeps, nouter;
};

__host__ __device__ __forceinline__ uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }
__host__ __device__ __forceinline__ uint32_t align128(uint32_t x) { return (x + 127u) & ~127u; }

template <typename Prog>
static uint32_t program_smem_bytes(const Prog &P) {
    return align128(align16(P.nsweeps * (uint32_t)sizeof(Sweep)) + align16(P.nregops * (uint32_t)sizeof(RegOp)) +
                    align16(P.nouter * (uint32_t)sizeof(OuterPred)) + align16(P.ngen * (uint32_t)sizeof(TileOp)) +
                    align16((uint32_t)P.mat_bytes) + 64u + (uint32_t)sizeof(uint64_t) * ((1u << P.H) + 8u));
}

__device__ __forceinline__ void copy_words(void *dst, const void *src, uint32_t bytes, int tid, int nt) {
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
    uint32_t *d = reinterpret_cast<uint32_t *>(dst);
    for (uint32_t i = tid; i < bytes / 4; i += nt) d[i] = s[i];
}

// copy the program from kernel-parameter space into shared memory (once per CTA); returns its size
template <typename Prog>
__device__ __forceinline__ uint32_t stage_program(unsigned char *smem, const Prog &P, ProgView &pv,
                                                  const uint64_t *&chunk_off, int tid, int nt) {
    uint32_t o = 0;
    pv.sweeps = reinterpret_cast<const Sweep *>(smem + o);
    copy_words(smem + o, P.sweeps, P.nsweeps * (uint32_t)sizeof(Sweep), tid, nt);
    o += align16(P.nsweeps * (uint32_t)sizeof(Sweep));
    pv.regops = reinterpret_cast<const RegOp *>(smem + o);
    copy_words(smem + o, P.regops, P.nregops * (uint32_t)sizeof(RegOp), tid, nt);
    o += align16(P.nregops * (uint32_t)sizeof(RegOp));
    pv.outer = reinterpret_cast<const OuterPred *>(smem + o);
    copy_words(smem + o, P.outer, P.nouter * (uint32_t)sizeof(OuterPred), tid, nt);
    o += align16(P.nouter * (uint32_t)sizeof(OuterPred));
    pv.gen = reinterpret_cast<const TileOp *>(smem + o);
    copy_words(smem + o, P.gen, P.ngen * (uint32_t)sizeof(TileOp), tid, nt);
    o += align16(P.ngen * (uint32_t)sizeof(TileOp));
    pv.pool = smem + o;
    copy_words(smem + o, P.mats, (uint32_t)P.mat_bytes, tid, nt);
    o += align16((uint32_t)P.mat_bytes);
    pv.outer_ok = reinterpret_cast<uint64_t *>(smem + o);      // [2 tile parities] x [up to 4 groups of the tensor-core kernel]
    o += 64u;
    uint64_t *co = reinterpret_cast<uint64_t *>(smem + o);
    for (uint32_t c = tid; c < (1u << P.H); c += nt) {
        uint64_t off = 0;
        for (int j = 0; j < P.H; ++j)
            if ((c >> j) & 1u) off |= 1ull << P.hbits[j];
        co[c] = off;
    }
    chunk_off = co;
    o += (uint32_t)sizeof(uint64_t) * (1u << P.H);
    // flat offsets of the 8 combinations of the top three tile bits (compile-time-geometry staging loops)
    uint64_t *hi = reinterpret_cast<uint64_t *>(smem + o);
    if (tid < 8) {
        const int mt = P.L + P.H;
        uint64_t off = 0;
        for (int j = 0; j < 3; ++j) {
            const int i = mt - 3 + j;
            if (i >= 0 && ((tid >> j) & 1)) off |= 1ull << (i < P.L ? i : P.hbits[i - P.L]);
        }
        hi[tid] = off;
    }
    o += (uint32_t)sizeof(uint64_t) * 8u;
    pv.nsweeps = P.nsweeps;
    pv.nouter = P.nouter;
    return align128(o);
}

// warp 0 evaluates the tile-uniform predicates of tile `tile_base` into pv.outer_ok[parity]
__device__ __forceinline__ void eval_outer(const ProgView &pv, uint64_t tile_base, int parity, int tid) {
    if (pv.nouter == 0 || tid >= 32) return;
    uint64_t mask = 0;
    for (int r = 0; r < pv.nouter; r += 32) {
        const int i = r + tid;
        const bool ok = i < pv.nouter && (tile_base & pv.outer[i].care) == pv.outer[i].val;
        mask |= (uint64_t)__ballot_sync(0xffffffffu, ok) << r;
    }
    if (tid == 0) pv.outer_ok[parity] = mask;
}

template <typename R, typename Layout, int MT>
__device__ __forceinline__ void run_program(typename CxT<R>::T *tile, const ProgView &pv, int m, uint64_t tile_base,
                                            int parity, int tid, int nt) {
    using C = typename CxT<R>::T;
    constexpr int RB = sizeof(C) == 8 ? REG_BITS_C64 : REG_BITS_C128;
    const C *pool = reinterpret_cast<const C *>(pv.pool);
    // bit OUTER_SLOT_NONE is always set: ops without a tile-level predicate name that slot
    const uint64_t outer_ok = (pv.nouter ? pv.outer_ok[parity] : 0ull) | (1ull << OUTER_SLOT_NONE);
    for (int s = 0; s < pv.nsweeps; ++s) {
        const Sweep &sw = pv.sweeps[s];
        if (sw.kind == SWEEP_REG) {
            if (MT > 0) reg_sweep_fast<R, RB, Layout>(tile, sw, pv.regops, pool, outer_ok, tid);
            else reg_sweep<R, RB, Layout, MT>(tile, m, sw, pv.regops, pool, outer_ok, tid, nt);
        }
        else apply_generic<R, Layout>(tile, m, pv.gen[sw.first], pool, tile_base, tid, nt);
        __syncthreads();
    }
}

template <typename Prog>
__device__ __forceinline__ uint64_t tile_base_of(uint64_t t, const Prog &P) {
    uint64_t x = t;
    for (int j = 0; j < P.H; ++j) x = insert_zero(x, P.hbits[j] - P.L);
    return x << P.L;
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
template <typename R> __device__ __forceinline__ uint4 vec_scale(const uint4 v, R s);
template <> __device__ __forceinline__ uint4 vec_scale<float>(const uint4 v, float s) {
    uint4 r;
    r.x = __float_as_uint(__uint_as_float(v.x) * s); r.y = __float_as_uint(__uint_as_float(v.y) * s);
    r.z = __float_as_uint(__uint_as_float(v.z) * s); r.w = __float_as_uint(__uint_as_float(v.w) * s);
    return r;
}
template <> __device__ __forceinline__ uint4 vec_scale<double>(const uint4 v, double s) {
    const double a = __hiloint2double(v.y, v.x) * s, b = __hiloint2double(v.w, v.z) * s;
    uint4 r;
    r.x = __double2loint(a); r.y = __double2hiint(a); r.z = __double2loint(b); r.w = __double2hiint(b);
    return r;
}


// ------------------------------------------------------------------------------------------------
// Variant 0: cp.async ring, padded layout, every warp both copies and computes
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n) {   // n = groups allowed to remain in flight
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}

// complex factor applied to a 16-byte vector of amplitudes while it streams out
template <typename R> __device__ __forceinline__ uint4 vec_cscale(const uint4 v, R fr, R fi);
template <> __device__ __forceinline__ uint4 vec_cscale<float>(const uint4 v, float fr, float fi) {
    const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z), d = __uint_as_float(v.w);
    uint4 r;
    r.x = __float_as_uint(fmaf(fr, a, -fi * b)); r.y = __float_as_uint(fmaf(fr, b, fi * a));
    r.z = __float_as_uint(fmaf(fr, c, -fi * d)); r.w = __float_as_uint(fmaf(fr, d, fi * c));
    return r;
}
template <> __device__ __forceinline__ uint4 vec_cscale<double>(const uint4 v, double fr, double fi) {
    const double a = __hiloint2double(v.y, v.x), b = __hiloint2double(v.w, v.z);
    const double re = fma(fr, a, -fi * b), im = fma(fr, b, fi * a);
    uint4 r;
    r.x = __double2loint(re); r.y = __double2hiint(re); r.z = __double2loint(im); r.w = __double2hiint(im);
    return r;
}
// squared magnitude of a vector in the accumulation type of the tile-local partial sum
__device__ __forceinline__ void vec_norm_acc(float &acc, const uint4 v, float) {
    const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z), d = __uint_as_float(v.w);
    acc = fmaf(a, a, acc); acc = fmaf(b, b, acc); acc = fmaf(c, c, acc); acc = fmaf(d, d, acc);
}
__device__ __forceinline__ void vec_norm_acc(double &acc, const uint4 v, double) {
    const double a = __hiloint2double(v.y, v.x), b = __hiloint2double(v.w, v.z);
    acc = fma(a, a, acc); acc = fma(b, b, acc);
}

// MT > 0: compile-time geometry -- the tile has MT bits, its low LSTD bits are contiguous, the CTA has
// TILE_THREADS threads; every staging loop has a constant trip count.  MT == 0: everything at run time
// (small states, unusual tiles).
// threads per CTA of an instantiation: the 64 KB tiles (one CTA per SM) get 512 threads so that the SM
// still holds 16 warps
template <typename R, int MT> struct PipeThreads {
    static constexpr int MSTD = sizeof(typename CxT<R>::T) == 8 ? TILE_MIN_BITS_C64 : TILE_MIN_BITS_C128;
    static constexpr int value = MT == MSTD + 1 ? 2 * TILE_THREADS : (MT == MSTD - 1 ? TILE_THREADS / 2 : TILE_THREADS);
};

template <typename R, typename Prog, int MT, int OCC>
__global__ void __launch_bounds__((PipeThreads<R, MT>::value), OCC) tile_kernel_pipe(const __grid_constant__ Prog P) {
    using C = typename CxT<R>::T;
    using Layout = PaddedLayout<C>;
    constexpr bool FAST = MT > 0;
    constexpr int NT = PipeThreads<R, MT>::value;
    constexpr int VSHIFT = sizeof(C) == 8 ? 1 : 0;       // amplitudes per 16-byte vector = 1 << VSHIFT
    extern __shared__ __align__(128) unsigned char smem_all[];
    const int tid = threadIdx.x;
    const int nt = FAST ? NT : (int)blockDim.x;
    ProgView pv;
    const uint64_t *chunk_off;
    const uint32_t prog_bytes = stage_program(smem_all, P, pv, chunk_off, tid, nt);
    const uint64_t *hi_off = chunk_off + (1u << P.H);     // the 8 offsets of the top three tile bits
    unsigned char *ring = smem_all + prog_bytes;
    const int L = P.L;
    const int m = FAST ? MT : P.L + P.H;
    const int S = P.stages;
    const uint32_t stage_bytes = (uint32_t)Layout::bytes(m);
    const uint32_t nchunks = 1u << (m - L);
    const uint32_t chunk_vecs = (1u << L) >> VSHIFT;     // 16-byte vectors per contiguous chunk
    // this thread's vectors: v = tid, tid + nt, ... ; the padded map is additive over the disjoint bit
    // ranges of tid and the stride, so the shared-memory address advances by a constant
    const uint32_t svec_tid = Layout::vec((uint32_t)tid);
    const int scale_mode = P.scale_mode;                 // 0 none, 1 real, 2 complex
    R fr = R(1), fi = R(0);
    if (scale_mode) {
        const double s = load_inv_norm(P.norm_in);
        fr = (R)(s * P.gphase[0]); fi = (R)(s * P.gphase[1]);
    }
    // Compile-time geometry: the tile has 8 * NT vectors; vector v = tid + k * NT has its low bits from tid
    // and its top three tile bits from k.  Tile bit i sits at flat bit i (i < L) or hbits[i - L], so the
    // flat offset of a vector is (tile base) + (a per-thread constant) + (one of 8 per-pass constants),
    // whatever the split between contiguous low bits and high bits (L is a run-time property of the pass).
    constexpr int NK = 8;
    uint64_t low_tid = 0;
    if (FAST) {
        for (int b = 0; b < MT - 3 - VSHIFT; ++b)
            if ((tid >> b) & 1) {
                const int i = b + VSHIFT;
                low_tid |= 1ull << (i < L ? i : P.hbits[i - L]);
            }
    }
    const uint4 *src = reinterpret_cast<const uint4 *>(P.src) + (FAST ? (low_tid >> VSHIFT) : (uint64_t)tid);
    uint4 *dst = reinterpret_cast<uint4 *>(P.dst) + (FAST ? (low_tid >> VSHIFT) : (uint64_t)tid);
    double nacc[1] = {0.0};
    const uint64_t my_tiles = (P.ntiles > blockIdx.x) ? (P.ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    __syncthreads();   // program + offset tables visible

    auto issue_load = [&](uint64_t it, int st) {
        const uint64_t base = tile_base_of(blockIdx.x + it * gridDim.x, P);
        const uint32_t stage_u32 = smem_u32(ring + (size_t)st * stage_bytes) + svec_tid * 16u;
        if (FAST) {
            uint32_t sp = stage_u32;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const uint4 *gp = src + ((base + hi_off[k]) >> VSHIFT);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp + Layout::vec((uint32_t)(k * NT)) * 16u), "l"(gp)
                             : "memory");
            }
        } else {
            for (uint32_t c = 0; c < nchunks; ++c) {
                const uint4 *gp = src + ((base + chunk_off[c]) >> VSHIFT);
                const uint32_t sp = stage_u32 + Layout::vec(c * chunk_vecs) * 16u;
                for (uint32_t w = 0; w < chunk_vecs; w += nt) {
                    if (w + tid < chunk_vecs)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp + Layout::vec(w) * 16u), "l"(gp) : "memory");
                    gp += nt;
                }
            }
        }
    };

    for (int s = 0; s < S - 1; ++s) {
        if ((uint64_t)s < my_tiles) issue_load(s, s);
        cp_async_commit();
    }
    int st = 0, st_next = S - 1;       // stage of tile `it`, stage the prefetch goes to
    for (uint64_t it = 0; it < my_tiles; ++it) {
        const uint64_t base = tile_base_of(blockIdx.x + it * gridDim.x, P);
        eval_outer(pv, base, (int)(it & 1), tid);
        cp_async_wait_pending(S - 2);      // this thread's copies of tile `it` have landed
        __syncthreads();                   // ... everyone's have, and stage (it-1)%S is drained
        const uint64_t nxt = it + (uint64_t)(S - 1);
        if (nxt < my_tiles) issue_load(nxt, st_next);
        cp_async_commit();
        unsigned char *stage = ring + (size_t)st * stage_bytes;
        run_program<R, Layout, MT>(reinterpret_cast<C *>(stage), pv, m, base, (int)(it & 1), tid, nt);
        // stream the tile out: fold the lazy scale / global phase, accumulate the squared norm
        R tacc = R(0);
        const uint4 *sbase = reinterpret_cast<const uint4 *>(stage) + svec_tid;
        auto emit = [&](const uint4 *sp, uint4 *gp) {
            uint4 val = *sp;
            if (scale_mode == 1) val = vec_scale<R>(val, fr);
            else if (scale_mode == 2) val = vec_cscale<R>(val, fr, fi);
            vec_norm_acc(tacc, val, R(0));
            *gp = val;
        };
        if (FAST) {
#pragma unroll
            for (int k = 0; k < NK; ++k) emit(sbase + Layout::vec((uint32_t)(k * NT)), dst + ((base + hi_off[k]) >> VSHIFT));
        } else {
            for (uint32_t c = 0; c < nchunks; ++c) {
                uint4 *gp = dst + ((base + chunk_off[c]) >> VSHIFT);
                const uint4 *sp = sbase + Layout::vec(c * chunk_vecs);
                for (uint32_t w = 0; w < chunk_vecs; w += nt) {
                    if (w + tid < chunk_vecs) emit(sp + Layout::vec(w), gp);
                    gp += nt;
                }
            }
        }
        nacc[0] += (double)tacc;
        st = st + 1 == S ? 0 : st + 1;
        st_next = st_next + 1 == S ? 0 : st_next + 1;
    }
    cp_async_wait_pending(0);
    if (P.norm_out) grid_sum_publish<1>(nacc, P.ws, P.norm_out);
}

// ------------------------------------------------------------------------------------------------
// Variant 1: TMA bulk-copy ring (cp.async.bulk + mbarrier), linear layout
// ------------------------------------------------------------------------------------------------
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

template <typename R, typename Prog>
__global__ void __launch_bounds__(TILE_THREADS_BULK, 1) tile_kernel_bulk(const __grid_constant__ Prog P) {
    using C = typename CxT<R>::T;
    using Layout = LinearLayout<C>;
    extern __shared__ __align__(128) unsigned char smem_all[];
    __shared__ __align__(8) uint64_t full_bar[TILE_MAX_STAGES];

    const int tid = threadIdx.x, nt = blockDim.x;
    ProgView pv;
    const uint64_t *chunk_off;
    const uint32_t prog_bytes = stage_program(smem_all, P, pv, chunk_off, tid, nt);
    unsigned char *ring = smem_all + prog_bytes;
    const int m = P.L + P.H;
    const int S = P.stages;
    const uint32_t tile_bytes = (uint32_t)((sizeof(C)) << m);
    const uint32_t chunk_bytes = (uint32_t)((sizeof(C)) << P.L);
    const uint32_t nchunks = 1u << P.H;
    const int scale_mode = P.scale_mode;
    R fr = R(1), fi = R(0);
    if (scale_mode) {
        const double s = load_inv_norm(P.norm_in);
        fr = (R)(s * P.gphase[0]); fi = (R)(s * P.gphase[1]);
    }
    const unsigned char *src = reinterpret_cast<const unsigned char *>(P.src);
    unsigned char *dst = reinterpret_cast<unsigned char *>(P.dst);
    double nacc[1] = {0.0};
    const uint64_t my_tiles = (P.ntiles > blockIdx.x) ? (P.ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_load = [&](uint64_t it) {
        const int st = (int)(it % S);
        const uint64_t base = tile_base_of(blockIdx.x + it * gridDim.x, P);
        unsigned char *sdst = ring + (size_t)st * tile_bytes;
        mbar_expect_tx(&full_bar[st], tile_bytes);
        for (uint32_t c = 0; c < nchunks; ++c)
            bulk_g2s(sdst + (size_t)c * chunk_bytes, src + (base + chunk_off[c]) * sizeof(C), chunk_bytes, &full_bar[st]);
    };

    if (tid == 0) {
        const uint64_t pre = my_tiles < (uint64_t)(S - 1) ? my_tiles : (uint64_t)(S - 1);
        for (uint64_t it = 0; it < pre; ++it) issue_load(it);
    }
    for (uint64_t it = 0; it < my_tiles; ++it) {
        const int st = (int)(it % S);
        const uint32_t parity = (uint32_t)((it / S) & 1);
        const uint64_t base = tile_base_of(blockIdx.x + it * gridDim.x, P);
        C *tile = reinterpret_cast<C *>(ring + (size_t)st * tile_bytes);
        eval_outer(pv, base, (int)(it & 1), tid);
        mbar_wait(&full_bar[st], parity);
        if (pv.nouter) __syncthreads();
        run_program<R, Layout, 0>(tile, pv, m, base, (int)(it & 1), tid, nt);
        if (P.norm_out || scale_mode) {
            uint4 *tv = reinterpret_cast<uint4 *>(tile);
            const uint32_t tile_vecs = tile_bytes >> 4;
            R tacc = R(0);
            for (uint32_t v = tid; v < tile_vecs; v += nt) {
                uint4 val = tv[v];
                if (scale_mode == 1) { val = vec_scale<R>(val, fr); tv[v] = val; }
                else if (scale_mode == 2) { val = vec_cscale<R>(val, fr, fi); tv[v] = val; }
                vec_norm_acc(tacc, val, R(0));
            }
            nacc[0] += (double)tacc;
        }
        fence_proxy_async();   // generic-proxy writes of this stage -> visible to the bulk store
        __syncthreads();
        if (tid == 0) {
            for (uint32_t c = 0; c < nchunks; ++c)
                bulk_s2g(dst + (base + chunk_off[c]) * sizeof(C),
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
// Variant 2: tensor-core sweeps (tcgen05 + tensor memory), complex64
// ------------------------------------------------------------------------------------------------
#include "qcs_tc.cuh"

// ------------------------------------------------------------------------------------------------
// host side: geometry, sweep planning, launch
// ------------------------------------------------------------------------------------------------
template <typename R, typename Prog>
static int launch_tile(Prog &P, const DeviceInfo &dev, int mode, cudaStream_t stream) {
    using C = typename CxT<R>::T;
    const int m = P.L + P.H;
    const size_t prog = program_smem_bytes(P);
    if (mode == TILE_MODE_BULK && (sizeof(C) << P.L) >= 16) {
        auto kern = tile_kernel_bulk<R, Prog>;
        static bool attr_done[64] = {};         // the opt-in is per device
        int device = 0;
        cudaGetDevice(&device);
        if (device >= 64 || !attr_done[device]) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_BUDGET);
            if (device < 64) attr_done[device] = true;
        }
        const size_t tile_bytes = sizeof(C) << m;
        int S = (int)((TILE_SMEM_BUDGET - prog - 1024) / tile_bytes);
        S = S > TILE_MAX_STAGES ? TILE_MAX_STAGES : S;
        S = tune_int("QCS_BULK_STAGES", S);
        if (S < 2) S = 2;
        if (S > TILE_MAX_STAGES) S = TILE_MAX_STAGES;
        P.stages = S;
        const size_t smem = (size_t)S * tile_bytes + prog;
        if (smem > TILE_SMEM_BUDGET) return set_error(QCS_ERR_UNSUPPORTED, "tile does not fit shared memory");
        uint64_t grid = (uint64_t)dev.sms;
        if (grid > P.ntiles) grid = P.ntiles;
        int threads = tune_int("QCS_BULK_THREADS", TILE_THREADS_BULK);
        if (threads > TILE_THREADS_BULK) threads = TILE_THREADS_BULK;
        while (threads > 32 && (1u << m) < (unsigned)threads) threads >>= 1;
        kern<<<(unsigned)grid, threads, smem, stream>>>(P);
    } else {
        constexpr int MSTD = sizeof(C) == 8 ? TILE_MIN_BITS_C64 : TILE_MIN_BITS_C128;
        constexpr int LSTD = sizeof(C) == 8 ? TILE_LOW_BITS_C64 : TILE_LOW_BITS_C128;
        const bool fast_ok = tune_int("QCS_PIPE_GENERIC", 0) == 0;
        int ctas = tune_int("QCS_PIPE_CTAS", 2);
        if (ctas < 1) ctas = 1;
        if (ctas > 2) ctas = 2;
        const bool fast12 = fast_ok && m == MSTD, fast13 = fast_ok && m == MSTD + 1, fast11 = fast_ok && m == MSTD - 1;
        void (*kern)(const Prog) = tile_kernel_pipe<R, Prog, 0, 2>;
        if (fast12) kern = tile_kernel_pipe<R, Prog, MSTD, 2>;
        else if (fast13) { kern = tile_kernel_pipe<R, Prog, MSTD + 1, 1>; ctas = 1; }
        else if (fast11) { kern = tile_kernel_pipe<R, Prog, MSTD - 1, 4>; ctas = tune_int("QCS_PIPE_CTAS11", 4) == 3 ? 3 : 4; }
        const bool is_fast = fast12 || fast13 || fast11;
        static bool attr_done[64] = {};         // the opt-in is per device
        int device = 0;
        cudaGetDevice(&device);
        if (device >= 64 || !attr_done[device]) {
            const int lim = (int)TILE_SMEM_BUDGET;
            cudaFuncSetAttribute(tile_kernel_pipe<R, Prog, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            cudaFuncSetAttribute(tile_kernel_pipe<R, Prog, MSTD, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            cudaFuncSetAttribute(tile_kernel_pipe<R, Prog, MSTD + 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            cudaFuncSetAttribute(tile_kernel_pipe<R, Prog, MSTD - 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            if (device < 64) attr_done[device] = true;
        }
        const size_t stage_bytes = PaddedLayout<C>::bytes(m);
        const size_t sm_smem = 227 * 1024;      // opt-in shared memory per SM; each resident CTA reserves 1 KB
        int S = (int)((sm_smem / ctas - prog - 1024) / stage_bytes);
        while (S < 2 && ctas > 1) {
            --ctas;
            S = (int)((sm_smem / ctas - prog - 1024) / stage_bytes);
        }
        S = S > TILE_MAX_STAGES ? TILE_MAX_STAGES : S;
        while (S > 2 && (size_t)S * stage_bytes + prog > TILE_SMEM_BUDGET) --S;
        const int tuned = tune_int("QCS_PIPE_STAGES", S);
        if (tuned < S) S = tuned;
        if (S < 2) S = 2;
        P.stages = S;
        const size_t smem = (size_t)S * stage_bytes + prog;
        if (smem > TILE_SMEM_BUDGET) return set_error(QCS_ERR_UNSUPPORTED, "tile does not fit shared memory");
        uint64_t grid = (uint64_t)dev.sms * ctas;
        if (grid > P.ntiles) grid = P.ntiles;
        if (grid > MAX_GRID) grid = MAX_GRID;
        int threads = fast13 ? 2 * TILE_THREADS : (fast11 ? TILE_THREADS / 2 : TILE_THREADS);
        if (!is_fast)
            while (threads > 32 && (1u << m) < (unsigned)threads * 2u) threads >>= 1;
        kern<<<(unsigned)grid, threads, smem, stream>>>(P);
    }
    return cudaGetLastError() == cudaSuccess ? QCS_OK : set_error(QCS_ERR_CUDA, "tile kernel launch failed");
}

namespace {
struct HostOp {            // one entry of the caller's list, classified
    int kind;              // 0 = U1 (one dense target), 1 = phases (any diagonal), 2 = generic dense
    uint64_t nd_mask;      // flat bits acted on non-diagonally
    uint64_t d_mask;       // flat bits acted on diagonally (controls, diagonal targets)
};
}  // namespace

// ---- tensor-core blocks: host side ----------------------------------------------------------------
static thread_local int32_t *g_sweep_of_op = nullptr;           // set by tile_plan_order: sweep index of every input op
static thread_local unsigned char *g_block_dump = nullptr;     // set by tile_plan_blocks (diagnostics / tests)
static thread_local int32_t *g_block_info = nullptr;

// round to nearest even, to fp16 and back / the 16 bits (host)
static inline uint16_t half_bits_host(float x) {
    const __half h = __float2half_rn(x);
    uint16_t u;
    std::memcpy(&u, &h, 2);
    return u;
}
static inline float half_value_host(float x) { return __half2float(__float2half_rn(x)); }

// Is the op a unitary (within 1e-4, entry-wise on U^H U - 1)?  The tensor-core kernel ranges its fp16 operands with the
// norm of the state, which only unitaries keep; anything else runs on the register-sweep kernel in fp32.
static bool op_is_unitary(const qcs_op &o) {
    const int dim = 1 << o.k;
    const double tol = 1e-4;
    if (o.kind == QCS_OP_DIAG) {
        for (int y = 0; y < dim; ++y) {
            const double m2 = o.matrix[2 * y] * o.matrix[2 * y] + o.matrix[2 * y + 1] * o.matrix[2 * y + 1];
            if (!(std::fabs(m2 - 1.0) <= tol)) return false;
        }
        return true;
    }
    for (int a = 0; a < dim; ++a)
        for (int b = a; b < dim; ++b) {
            double sr = 0.0, si = 0.0;                    // sum_y conj(U[y][a]) U[y][b]
            for (int y = 0; y < dim; ++y) {
                const double ar = o.matrix[2 * (y * dim + a)], ai = o.matrix[2 * (y * dim + a) + 1];
                const double br = o.matrix[2 * (y * dim + b)], bi = o.matrix[2 * (y * dim + b) + 1];
                sr += ar * br + ai * bi;
                si += ar * bi - ai * br;
            }
            if (!(std::fabs(sr - (a == b ? 1.0 : 0.0)) <= tol) || !(std::fabs(si) <= tol)) return false;
        }
    return true;
}

// Product of the listed ops (in order) as a 2^5 x 2^5 complex matrix over the block index, written as the two
// shared-memory images of a slot (qcs_tc.cuh: fp16, K-major, no swizzle; rows 0..31 = Re U, rows 32..63 = Im U; "hi"
// image = fp16(U), "lo" image = fp16(U - hi)).  block_bit(flat bit) = bit of the block index that flat bit maps to.
template <typename F>
static void tc_build_block(const qcs_op *ops, const int *list, int count, F block_bit, unsigned char *image) {
    constexpr int D = 1 << TC_BLOCK_BITS;
    static thread_local double U[D][D][2], V[D][D][2];
    for (int r = 0; r < D; ++r)
        for (int c = 0; c < D; ++c) { U[r][c][0] = r == c ? 1.0 : 0.0; U[r][c][1] = 0.0; }
    for (int t = 0; t < count; ++t) {
        const qcs_op &o = ops[list[t]];
        uint32_t cmask = 0, tbit[QCS_MAX_OP_QUBITS], tmask = 0;
        for (int b = 0; b < 64; ++b)
            if ((o.ctrl_mask >> b) & 1ull) cmask |= 1u << block_bit(b);
        for (int j = 0; j < o.k; ++j) { tbit[j] = 1u << block_bit(o.bitpos[j]); tmask |= tbit[j]; }
        const int dim = 1 << o.k;
        auto offs = [&](int x) {                       // matrix index x (operator qubit 0 = its most significant bit)
            uint32_t off = 0;
            for (int j = 0; j < o.k; ++j)
                if ((x >> (o.k - 1 - j)) & 1) off |= tbit[j];
            return off;
        };
        for (int r = 0; r < D; ++r)
            for (int c = 0; c < D; ++c) { V[r][c][0] = U[r][c][0]; V[r][c][1] = U[r][c][1]; }
        for (uint32_t sidx = 0; sidx < (uint32_t)D; ++sidx) {
            if ((sidx & tmask) || (sidx & cmask) != cmask) continue;      // group base: targets 0, controls 1
            for (int y = 0; y < dim; ++y) {
                const uint32_t ry = sidx | offs(y);
                for (int c = 0; c < D; ++c) {
                    double ar = 0.0, ai = 0.0;
                    if (o.kind == QCS_OP_DIAG) {
                        const double mr = o.matrix[2 * y], mi = o.matrix[2 * y + 1];
                        ar = mr * V[ry][c][0] - mi * V[ry][c][1];
                        ai = mr * V[ry][c][1] + mi * V[ry][c][0];
                    } else {
                        for (int x = 0; x < dim; ++x) {
                            const double mr = o.matrix[2 * (y * dim + x)], mi = o.matrix[2 * (y * dim + x) + 1];
                            const uint32_t rx = sidx | offs(x);
                            ar += mr * V[rx][c][0] - mi * V[rx][c][1];
                            ai += mr * V[rx][c][1] + mi * V[rx][c][0];
                        }
                    }
                    U[ry][c][0] = ar; U[ry][c][1] = ai;
                }
            }
        }
    }
    // element (row n, column k) of an image sits at (n % 8) * 16 + (n / 8) * 512 + (k / 8) * 128 + (k % 8) * 2 bytes:
    // 8 x 16-byte core matrices, LBO = 128 bytes between the K chunks, SBO = 512 bytes between the 8-row groups
    uint16_t *hi = reinterpret_cast<uint16_t *>(image), *lo = reinterpret_cast<uint16_t *>(image + TC_HALF_BYTES);
    for (int part = 0; part < 2; ++part)
        for (int r = 0; r < D; ++r)
            for (int c = 0; c < D; ++c) {
                const int nrow = part * D + r;
                const size_t off = ((size_t)(nrow % 8) * 16 + (size_t)(nrow / 8) * 512 + (size_t)(c / 8) * 128 + (size_t)(c % 8) * 2) / 2;
                const double v = U[r][c][part];
                const float h = half_value_host((float)v);
                hi[off] = half_bits_host((float)v);
                lo[off] = half_bits_host((float)(v - (double)h));
            }
}

// The images travel host -> device through a pinned staging ring (one set per launch in flight), so the copy is
// asynchronous and the host buffer is not reused before its copy has run.
static int tc_upload(const unsigned char *images, int nslots, void *ws, cudaStream_t stream) {
    static std::mutex mu;
    static unsigned char *pinned = nullptr;
    static cudaEvent_t ev[TC_UPLOAD_RING];
    static bool used[TC_UPLOAD_RING];
    static int next = 0;
    std::lock_guard<std::mutex> lock(mu);
    const size_t set_bytes = (size_t)TC_MAX_SLOTS * TC_SLOT_BYTES;
    if (!pinned) {
        if (cudaHostAlloc(reinterpret_cast<void **>(&pinned), set_bytes * TC_UPLOAD_RING, cudaHostAllocDefault) != cudaSuccess)
            return set_error(QCS_ERR_CUDA, "pinned staging allocation failed");
        for (int i = 0; i < TC_UPLOAD_RING; ++i) { cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); used[i] = false; }
    }
    const int slot = next;
    next = (next + 1) % TC_UPLOAD_RING;
    if (used[slot]) cudaEventSynchronize(ev[slot]);
    unsigned char *stage = pinned + (size_t)slot * set_bytes;
    std::memcpy(stage, images, (size_t)nslots * TC_SLOT_BYTES);
    if (cudaMemcpyAsync(reinterpret_cast<unsigned char *>(ws) + WS_TC_OFFSET, stage, (size_t)nslots * TC_SLOT_BYTES,
                        cudaMemcpyHostToDevice, stream) != cudaSuccess)
        return set_error(QCS_ERR_CUDA, "block-matrix upload failed");
    cudaEventRecord(ev[slot], stream);
    used[slot] = true;
    return QCS_OK;
}

template <typename Prog, int GT>
static int launch_tc_groups(Prog &P, const DeviceInfo &dev, cudaStream_t stream) {
    using Layout = PaddedLayout<float2>;
    auto kern = tile_kernel_tc<Prog, GT>;
    constexpr int NG = 512 / GT, STAGES = GT == 256 ? 2 : 1;
    static bool attr_done[64] = {};
    int device = 0;
    cudaGetDevice(&device);
    constexpr int SMEM_LIMIT = 227 * 1024 - 2048;      // opt-in maximum minus the kernel's static shared memory
    if (device >= 64 || !attr_done[device]) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess)
            return set_error(QCS_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
        if (device < 64) attr_done[device] = true;
    }
    const size_t smem = program_smem_bytes(P) + (size_t)P.nsweeps * 256 * sizeof(uint32_t) + (size_t)P.n_tc * TC_SLOT_BYTES +
                        (size_t)NG * STAGES * Layout::bytes(TILE_MIN_BITS_C64);
    if (smem > (size_t)SMEM_LIMIT) return set_error(QCS_ERR_UNSUPPORTED, "tensor-core pass does not fit shared memory");
    P.stages = STAGES;
    uint64_t grid = (uint64_t)dev.sms;
    if (grid * NG > P.ntiles) grid = (P.ntiles + NG - 1) / NG;
    kern<<<(unsigned)grid, 512, smem, stream>>>(P);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        static thread_local char msg[160];
        std::snprintf(msg, sizeof msg, "tensor-core tile kernel launch failed: %s (grid %u, smem %zu)", cudaGetErrorString(err),
                      (unsigned)grid, smem);
        return set_error(QCS_ERR_CUDA, msg);
    }
    return QCS_OK;
}

template <typename Prog>
static int launch_tc(Prog &P, const DeviceInfo &dev, cudaStream_t stream) {
    // four groups of 128 threads (one stage each) or two of 256 (two stages each): QCS_TC_GROUPS
    if (tune_int("QCS_TC_GROUPS", TC_DEFAULT_GROUPS) == 4) return launch_tc_groups<Prog, 128>(P, dev, stream);
    return launch_tc_groups<Prog, 256>(P, dev, stream);
}

template <typename Prog>
static int build_and_launch(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops,
                            const double *norm_in, double *norm_out, void *ws, cudaStream_t stream,
                            int32_t *stats = nullptr, const double *bound_in = nullptr) {
    const bool c64 = dtype == QCS_C64;
    const int vshift = c64 ? 1 : 0;
    Prog *Pp = new Prog();
    Prog &P = *Pp;
    struct Guard { Prog *p; ~Guard() { delete p; } } guard{Pp};

    // ---- validate + classify ------------------------------------------------------------------------
    HostOp hop[QCS_MAX_OPS];
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
        if (o.kind == QCS_OP_DENSE) {
            dense_mask |= seen;
            // 0: one target (register op), 3: two targets, few controls (register op on a pair of register
            // bits when both fit), 2: everything else dense (shared-memory op)
            hop[i].kind = o.k == 1 ? 0 : (o.k == 2 && __builtin_popcountll(o.ctrl_mask) <= 2 ? 3 : 2);
            hop[i].nd_mask = seen;
            hop[i].d_mask = o.ctrl_mask;
        } else if (o.kind == QCS_OP_DIAG) {
            hop[i].kind = 1;
            hop[i].nd_mask = 0;
            hop[i].d_mask = seen | o.ctrl_mask;
        } else {
            return set_error(QCS_ERR_ARG, "unknown op kind");
        }
    }

    // ---- tile geometry: low L bits + the dense targets above them, padded to the minimum tile -------
    const int lpref = c64 ? TILE_LOW_BITS_C64 : TILE_LOW_BITS_C128;
    const int mmax = tile_max_bits(dtype);
    const int mmin = (c64 ? TILE_MIN_BITS_C64 : TILE_MIN_BITS_C128) < mmax ? (c64 ? TILE_MIN_BITS_C64 : TILE_MIN_BITS_C128) : mmax;
    int L = n < lpref ? n : lpref;
    int H = 0;
    for (;; --L) {
        H = __builtin_popcountll(dense_mask >> L);
        if (L + H <= mmax) break;
        if (L <= vshift) return set_error(QCS_ERR_UNSUPPORTED, "fused ops touch too many high bits for one tile");
    }
    uint64_t high_mask = (dense_mask >> L) << L;
    for (int b = L; b < n && L + H < mmin; ++b)          // free bits: bigger tiles, more bytes in flight
        if (!((high_mask >> b) & 1ull)) { high_mask |= 1ull << b; ++H; }
    if (H > TILE_MAX_H) return set_error(QCS_ERR_UNSUPPORTED, "too many high tile bits");
    P.n = n; P.L = L; P.H = H;
    {
        int h = 0;
        for (int b = L; b < n; ++b)
            if ((high_mask >> b) & 1ull) P.hbits[h++] = b;
    }
    const int m = L + H;
    P.ntiles = 1ull << (n - m);
    P.src = src; P.dst = dst; P.norm_in = norm_in; P.norm_out = norm_out; P.ws = ws;
    const uint64_t tile_mask = ((1ull << L) - 1ull) | high_mask;
    auto local_of = [&](int b) -> int {
        if (b < L) return b;
        for (int h = 0; h < H; ++h) if (P.hbits[h] == b) return L + h;
        return -1;
    };
    auto local_mask = [&](uint64_t flat) -> uint32_t {   // flat-index mask (inside the tile) -> tile-local mask
        uint32_t r = 0;
        for (int b = 0; b < n; ++b)
            if ((flat >> b) & 1ull) r |= 1u << local_of(b);
        return r;
    };

    // ---- tensor-core sweeps: which ops could join a dense block ------------------------------------------
    constexpr bool TCPROG = std::is_same<Prog, TcProg>::value;
    const bool use_tc = TCPROG && c64 && m == TILE_MIN_BITS_C64 && n >= TILE_MIN_BITS_C64 && tune_int("QCS_TC", 1) != 0;
    if (TCPROG && !use_tc) return set_error(QCS_ERR_UNSUPPORTED, "not a tensor-core pass");
    int tc_max_slots = tune_int("QCS_TC_SLOTS", TC_DEFAULT_SLOTS);
    if (tc_max_slots > TC_MAX_SLOTS) tc_max_slots = TC_MAX_SLOTS;
    if (TCPROG)
        for (int i = 0; i < nops; ++i)
            if (!op_is_unitary(ops[i])) return set_error(QCS_ERR_UNSUPPORTED, "non-unitary op: not a tensor-core pass");
    static thread_local unsigned char tc_images[TC_MAX_SLOTS * TC_SLOT_BYTES];
    int tc_block_ops = 0;
    uint32_t op_lmask[QCS_MAX_OPS];        // tile-local mask of every bit the op touches inside the tile
    bool op_inside[QCS_MAX_OPS];           // ... and whether it touches nothing else (and fits a block)
    for (int i = 0; i < nops; ++i) {
        uint64_t all = ops[i].ctrl_mask;
        for (int j = 0; j < ops[i].k; ++j) all |= 1ull << ops[i].bitpos[j];
        op_lmask[i] = local_mask(all & tile_mask);
        op_inside[i] = (all & ~tile_mask) == 0 && __builtin_popcountll(all) <= TC_BLOCK_BITS;
    }
    P.n_tc = 0; P.tc_mats = nullptr; P.tc_norm = nullptr; P.tc_pad = tune_int("QCS_TC_DEBUG", 0);   // (read by scratch builds with -DQCS_TC_DBG only)

    // ---- matrix pool -----------------------------------------------------------------------------------
    const size_t csize = c64 ? 8 : 16;
    const size_t pool_cap = sizeof(P.mats) / csize;
    size_t used = 0;
    auto push_entries = [&](const double *re_im, size_t entries) -> int {
        if (used + entries > pool_cap) return -1;
        const int at = (int)used;
        if (c64) {
            float *p = reinterpret_cast<float *>(P.mats) + 2 * used;
            for (size_t e = 0; e < 2 * entries; ++e) p[e] = (float)re_im[e];
        } else {
            double *p = reinterpret_cast<double *>(P.mats) + 2 * used;
            for (size_t e = 0; e < 2 * entries; ++e) p[e] = re_im[e];
        }
        used += entries;
        return at;
    };
    // `count` reals, 16-byte aligned (read with 128-bit shared-memory loads); returns the entry index
    auto push_reals = [&](const double *v, int count) -> int {
        const size_t per = c64 ? 2 : 1;                          // entries per 16 bytes
        used = (used + per - 1) / per * per;
        const size_t entries = c64 ? (size_t)(count + 1) / 2 : (size_t)(count + 1) / 2;
        const size_t padded = (entries + per - 1) / per * per;
        if (used + padded > pool_cap) return -1;
        const int at = (int)used;
        if (c64) {
            float *p = reinterpret_cast<float *>(P.mats) + 2 * used;
            for (size_t e = 0; e < 2 * padded; ++e) p[e] = e < (size_t)count ? (float)v[e] : 0.0f;
        } else {
            double *p = reinterpret_cast<double *>(P.mats) + 2 * used;
            for (size_t e = 0; e < 2 * padded; ++e) p[e] = e < (size_t)count ? v[e] : 0.0;
        }
        used += padded;
        return at;
    };
    auto gphase_mul = [&](double re, double im) {
        const double gr = P.gphase[0] * re - P.gphase[1] * im, gi = P.gphase[0] * im + P.gphase[1] * re;
        P.gphase[0] = gr; P.gphase[1] = gi;
    };

    // ---- sweeps: greedy, commuting ops may overtake ops that wait for a later sweep --------------------
    const int rb = c64 ? REG_BITS_C64 : REG_BITS_C128;
    const int nreg = rb < m ? rb : m;
    bool any_dense2 = false;
    for (int i = 0; i < nops; ++i) {
        // a two-target op runs in registers only if its control bits inside the tile can all stay thread bits
        if (hop[i].kind == 3 && (nreg < 2 || tune_int("QCS_NO_DENSE2_REG", 0) ||
                                 __builtin_popcountll(ops[i].ctrl_mask & tile_mask) > m - nreg))
            hop[i].kind = 2;
        any_dense2 |= hop[i].kind == 3;
    }
    int pending[QCS_MAX_OPS];
    int npend = nops;
    for (int i = 0; i < nops; ++i) pending[i] = i;
    P.nsweeps = P.nregops = P.ngen = P.nouter = 0;
    P.gphase[0] = 1.0; P.gphase[1] = 0.0;
    int last_sweep_ops[QCS_MAX_OPS], n_last_sweep = 0;     // input ops of the most recent sweep (diagnostics)
    while (npend > 0) {
        if (P.nsweeps >= Prog::MAX_SWEEPS) return set_error(QCS_ERR_UNSUPPORTED, "too many sweeps in one pass");
        Sweep &sw = P.sweeps[P.nsweeps];
        const int first = pending[0];
        sw.tc_slot = -1; sw.hrank = 7;
        // ---- tensor-core sweep: register ops ("pre") followed by one dense block on five tile bits -------------
        bool tc_sweep = false;
        int blk_ops[QCS_MAX_OPS], nblk = 0, tc_hbit = -1;
        uint32_t S_local = 0;                  // tile-local mask of the register bits
        int S_count = 0;
        int taken[QCS_MAX_OPS], ntaken = 0, rest[QCS_MAX_OPS], nrest = 0;
        uint32_t S_forbid = 0;                 // control bits of the taken two-target ops: must stay thread bits
        if (use_tc && P.n_tc < tc_max_slots) {
            // walk the pending ops for a given block B (tile-local mask of 5 bits) whose bit `hb` is the thread bit
            auto walk = [&](uint32_t B, int hb, int *bk, int &nbk, int *pr, int &npr, int *rs, int &nrs) -> bool {
                const uint32_t R = B & ~(1u << hb);
                uint64_t skip_nd = 0, skip_d = 0, mm_nd = 0, mm_d = 0;
                nbk = npr = nrs = 0;
                bool first_taken = false;
                for (int i = 0; i < npend; ++i) {
                    const int idx = pending[i];
                    const HostOp &h = hop[idx];
                    bool take = false;
                    if ((h.nd_mask & (skip_nd | skip_d)) == 0 && (h.d_mask & skip_nd) == 0) {
                        if (op_inside[idx] && (op_lmask[idx] & ~B) == 0) {
                            bk[nbk++] = idx; mm_nd |= h.nd_mask; mm_d |= h.d_mask; take = true;
                        } else if ((h.nd_mask & (mm_nd | mm_d)) == 0 && (h.d_mask & mm_nd) == 0) {
                            // runs before the block: must commute with everything already in it
                            if (h.kind == 1) take = true;
                            else if (h.kind == 0 && ((R >> local_of(ops[idx].bitpos[0])) & 1u)) take = true;
                            if (take) pr[npr++] = idx;
                        }
                    }
                    if (take) { if (i == 0) first_taken = true; }
                    else { rs[nrs++] = idx; skip_nd |= h.nd_mask; skip_d |= h.d_mask; }
                }
                return first_taken;
            };
            // candidate bits: the tile bits of the first pending ops, in order of appearance
            int W[16], nW = 0;
            uint32_t Wmask = 0;
            const int cand_bits = tune_int("QCS_TC_CAND_BITS", 9), cand_ops = tune_int("QCS_TC_CAND_OPS", 48);
            for (int i = 0; i < npend && i < cand_ops && nW < cand_bits; ++i) {
                const uint32_t lm = op_lmask[pending[i]];
                for (int b = 0; b < m && nW < cand_bits; ++b)
                    if (((lm >> b) & 1u) && !((Wmask >> b) & 1u)) { W[nW++] = b; Wmask |= 1u << b; }
            }
            for (int b = m - 1; b >= 0 && nW < TC_BLOCK_BITS; --b)      // fewer than five wanted bits: pad
                if (!((Wmask >> b) & 1u)) { W[nW++] = b; Wmask |= 1u << b; }
            // the first pending op must make progress: its bits (block op) or its target (register op) are required
            uint32_t required = 0, reg_required = 0;
            if (op_inside[first] && __builtin_popcount(op_lmask[first]) <= TC_BLOCK_BITS) required = op_lmask[first];
            else if (hop[first].kind == 0) { required = 1u << local_of(ops[first].bitpos[0]); reg_required = required; }
            int best_score = -1, best_nblk = 0;
            uint32_t best_B = 0;
            int best_hb = -1;
            int bk2[QCS_MAX_OPS], pr2[QCS_MAX_OPS], rs2[QCS_MAX_OPS];
            if (hop[first].kind == 0 || hop[first].kind == 1 || required) {
                for (uint32_t sub = 0; sub < (1u << nW); ++sub) {
                    if (__builtin_popcount(sub) != TC_BLOCK_BITS) continue;
                    uint32_t B = 0;
                    for (int c = 0; c < nW; ++c)
                        if ((sub >> c) & 1u) B |= 1u << W[c];
                    if ((B & required) != required) continue;
                    int hb = -1;                                  // thread bit of the block: the highest non-required one, not bit 0
                    for (int b = m - 1; b >= 1; --b)
                        if (((B >> b) & 1u) && !((reg_required >> b) & 1u)) { hb = b; break; }
                    if (hb < 0) continue;
                    int nb, np, nr;
                    if (!walk(B, hb, bk2, nb, pr2, np, rs2, nr)) continue;
                    const int score = nb + np;
                    if (score > best_score || (score == best_score && nb > best_nblk)) {
                        best_score = score; best_nblk = nb; best_B = B; best_hb = hb;
                    }
                }
            }
            if (best_score > 0 && best_nblk >= 1) {
                // the plain register sweep the existing planner would form, for comparison (ops per unit of cost:
                // a sweep costs about as much as 15 register ops, a block about 12)
                int nb, np, nr;
                walk(best_B, best_hb, bk2, nb, pr2, np, rs2, nr);
                tc_sweep = true;
                tc_hbit = best_hb;
                S_local = best_B & ~(1u << best_hb); S_count = TC_BLOCK_BITS - 1;
                nblk = nb; ntaken = np; nrest = nr;
                for (int i = 0; i < nb; ++i) blk_ops[i] = bk2[i];
                for (int i = 0; i < np; ++i) taken[i] = pr2[i];
                for (int i = 0; i < nr; ++i) rest[i] = rs2[i];
            }
        }
        if (!tc_sweep && hop[first].kind == 2) {            // generic dense op: its own shared-memory sweep
            const qcs_op &o = ops[first];
            if (P.ngen >= Prog::MAX_GENERIC) return set_error(QCS_ERR_UNSUPPORTED, "too many generic ops");
            TileOp &t = P.gen[P.ngen];
            t.kind = o.kind; t.k = o.k; t.flags = 0;
            const int at = push_entries(o.matrix, (size_t)1 << (2 * o.k));
            if (at < 0) return set_error(QCS_ERR_UNSUPPORTED, "fused pass matrix pool exhausted");
            t.mat = at;
            for (int j = 0; j < QCS_MAX_OP_QUBITS; ++j) { t.lpos[j] = -1; t.gpos[j] = 0; }
            for (int j = 0; j < o.k; ++j) { t.lpos[j] = (int8_t)local_of(o.bitpos[j]); t.gpos[j] = (int8_t)o.bitpos[j]; }
            t.ctrl_local = local_mask(o.ctrl_mask & tile_mask);
            t.ctrl_outer = o.ctrl_mask & ~tile_mask;
            sw.kind = SWEEP_GENERIC; sw.first = P.ngen; sw.count = 1; sw.nreg = 0;
            ++P.ngen; ++P.nsweeps;
            n_last_sweep = 1; last_sweep_ops[0] = first;        // this op IS the most recent sweep
            if (g_sweep_of_op) g_sweep_of_op[first] = P.nsweeps - 1;
            for (int i = 1; i < npend; ++i) pending[i - 1] = pending[i];
            --npend;
            continue;
        }
        // choose the ops of this register sweep.  select(fixed, ...) walks the pending ops in order: an op is taken
        // if nothing it must follow is left behind and (one-target dense op) its target is a register bit; with
        // fixed == 0 the register set grows greedily with the first distinct targets met.
        auto select = [&](uint32_t fixed, uint32_t &S_out, int &S_cnt, int *tk, int &ntk, int *rs, int &nrs) -> int {
            uint64_t blk_nd = 0, blk_d = 0;
            uint32_t S = fixed, forbid = 0;
            int cnt = __builtin_popcount(fixed), score = 0;
            ntk = nrs = 0;
            for (int i = 0; i < npend; ++i) {
                const int idx = pending[i];
                const HostOp &h = hop[idx];
                bool can = h.kind != 2 && (h.nd_mask & (blk_nd | blk_d)) == 0 && (h.d_mask & blk_nd) == 0;
                if (can && h.kind == 0) {
                    const uint32_t tbit = 1u << local_of(ops[idx].bitpos[0]);
                    if (!(S & tbit)) {
                        if (!fixed && cnt < nreg && !(forbid & tbit)) { S |= tbit; ++cnt; }
                        else can = false;
                    }
                } else if (can && h.kind == 3) {
                    const uint32_t tb = (1u << local_of(ops[idx].bitpos[0])) | (1u << local_of(ops[idx].bitpos[1]));
                    const uint32_t cl = local_mask(ops[idx].ctrl_mask & tile_mask);
                    const uint32_t need = tb & ~S;
                    if ((cl & (S | tb)) || (need & forbid) || __builtin_popcount(forbid | cl) > m - nreg) can = false;
                    else if (need) {
                        if (!fixed && cnt + __builtin_popcount(need) <= nreg) { S |= need; cnt += __builtin_popcount(need); }
                        else can = false;
                    }
                    if (can) forbid |= cl;
                }
                if (can) { tk[ntk++] = idx; score += h.kind == 1 ? 1 : 16; }
                else { rs[nrs++] = idx; blk_nd |= h.nd_mask; blk_d |= h.d_mask; }
            }
            S_out = S; S_cnt = cnt; S_forbid = forbid;
            return score;
        };
        int best_score = tc_sweep ? 0 : select(0u, S_local, S_count, taken, ntaken, rest, nrest);
        if (!tc_sweep && nreg >= 2 && !any_dense2 && tune_int("QCS_SWEEP_SEARCH", 1)) {
            // alternatives: register sets drawn from the first distinct dense targets still pending (always with
            // the first one, so that the sweep makes progress); keep the set that absorbs the most ops
            int cand[8], ncand = 0;
            for (int i = 0; i < npend && ncand < 8; ++i) {
                const HostOp &h = hop[pending[i]];
                if (h.kind == 2) break;
                if (h.kind != 0) continue;
                const int lb = local_of(ops[pending[i]].bitpos[0]);
                bool dup = false;
                for (int c = 0; c < ncand; ++c) dup |= cand[c] == lb;
                if (!dup) cand[ncand++] = lb;
            }
            if (ncand > nreg) {
                int tk2[QCS_MAX_OPS], rs2[QCS_MAX_OPS];
                for (uint32_t sub = 0; sub < (1u << (ncand - 1)); ++sub) {
                    if (__builtin_popcount(sub) != nreg - 1) continue;
                    uint32_t fixed = 1u << cand[0];
                    for (int c = 1; c < ncand; ++c)
                        if ((sub >> (c - 1)) & 1u) fixed |= 1u << cand[c];
                    uint32_t S2;
                    int c2, n2, r2;
                    const int sc = select(fixed, S2, c2, tk2, n2, rs2, r2);
                    if (sc > best_score) {
                        best_score = sc; S_local = S2; S_count = c2; ntaken = n2; nrest = r2;
                        for (int i = 0; i < n2; ++i) taken[i] = tk2[i];
                        for (int i = 0; i < r2; ++i) rest[i] = rs2[i];
                    }
                }
            }
        }
        // complete the register set: first the aligned blocks that already hold a register bit, then from the top
        for (int blk = 0; blk * nreg < m && S_count < nreg; ++blk) {
            const uint32_t bm = ((1u << nreg) - 1u) << (blk * nreg);
            if (!(S_local & bm)) continue;
            for (int b = blk * nreg; b < (blk + 1) * nreg && b < m && S_count < nreg; ++b)
                if (!((S_local | S_forbid) & (1u << b))) { S_local |= 1u << b; ++S_count; }
        }
        for (int b = m - 1; b >= 0 && S_count < nreg; --b)
            if (!((S_local | S_forbid) & (1u << b))) { S_local |= 1u << b; ++S_count; }
        if (S_count < nreg) return set_error(QCS_ERR_UNSUPPORTED, "no free tile bit left for the register set");
        if (ntaken + nblk == 0) return set_error(QCS_ERR_UNSUPPORTED, "pass planner made no progress");   // (never loop)
        sw.kind = SWEEP_REG; sw.first = P.nregops; sw.count = 0; sw.nreg = nreg; sw.thread_phase = 0; sw.has_dense2 = 0;
        if (tc_sweep) {
            // thread bit 7 of the group goes to tile bit tc_hbit: its rank among the 8 non-register bits
            sw.hrank = (int16_t)(tc_hbit - __builtin_popcount(S_local & ((1u << tc_hbit) - 1u)));
            sw.tc_slot = (int16_t)P.n_tc;
            int idx5[32];                                          // tile-local bit -> bit of the block index
            for (int b = 0, j = 0; b < 32; ++b) { idx5[b] = -1; if ((S_local >> b) & 1u) idx5[b] = j++; }
            idx5[tc_hbit] = TC_BLOCK_BITS - 1;
            tc_build_block(ops, blk_ops, nblk, [&](int flat_bit) { return idx5[local_of(flat_bit)]; },
                           tc_images + (size_t)P.n_tc * TC_SLOT_BYTES);
            ++P.n_tc;
            tc_block_ops += nblk;
        }
        int8_t reg_index_of_local[32];
        for (int j = 0; j < REG_BITS_MAX; ++j) { sw.rpos[j] = 0; sw.himask[j] = 0; sw.stride[j] = 0; }
        for (int b = 0, j = 0; b < 32; ++b) {
            reg_index_of_local[b] = -1;
            if ((S_local >> b) & 1u) {
                if (j < REG_BITS_MAX) {
                    sw.rpos[j] = (int8_t)b;
                    sw.himask[j] = ~((1u << b) - 1u);
                    sw.stride[j] = c64 ? (uint32_t)(PaddedLayout<float2>::amp(1u << b) * sizeof(float2))
                                       : (uint32_t)(PaddedLayout<double2>::amp(1u << b) * sizeof(double2));
                }
                reg_index_of_local[b] = (int8_t)j;
                ++j;
            }
        }
        // flat-index (care, val) -> register / thread / tile parts of a RegOp; false if out of slots
        auto split_masks = [&](uint64_t care, uint64_t val, RegOp &r) -> bool {
            r.reg_care = r.reg_val = 0; r.base_care = r.base_val = 0; r.outer_slot = OUTER_SLOT_NONE;
            const uint64_t oc = care & ~tile_mask, ov = val & ~tile_mask;
            if (oc) {
                int slot = -1;
                for (int i = 0; i < P.nouter; ++i)
                    if (P.outer[i].care == oc && P.outer[i].val == ov) { slot = i; break; }
                if (slot < 0) {
                    if (P.nouter >= MAX_OUTER_PREDS) return false;
                    slot = P.nouter++;
                    P.outer[slot].care = oc; P.outer[slot].val = ov;
                }
                r.outer_slot = slot;
            }
            for (int b = 0; b < n; ++b) {
                if (!(((care & tile_mask) >> b) & 1ull)) continue;
                const int lb = local_of(b);
                const bool one = (val >> b) & 1ull;
                if (reg_index_of_local[lb] >= 0) {
                    r.reg_care |= (uint8_t)(1u << reg_index_of_local[lb]);
                    if (one) r.reg_val |= (uint8_t)(1u << reg_index_of_local[lb]);
                } else {
                    r.base_care |= (uint16_t)(1u << lb);
                    if (one) r.base_val |= (uint16_t)(1u << lb);
                }
            }
            return true;
        };
        n_last_sweep = ntaken + nblk;
        for (int t = 0; t < ntaken; ++t) last_sweep_ops[t] = taken[t];
        for (int t = 0; t < nblk; ++t) last_sweep_ops[ntaken + t] = blk_ops[t];
        if (g_sweep_of_op) {                    // (tensor-core sweeps are reported as 1000 + index)
            for (int t = 0; t < ntaken + nblk; ++t) g_sweep_of_op[last_sweep_ops[t]] = P.nsweeps + (tc_sweep ? 1000 : 0);
        }
        for (int t = 0; t < ntaken; ++t) {
            const qcs_op &o = ops[taken[t]];
            if (P.nregops >= Prog::MAX_REGOPS) return set_error(QCS_ERR_UNSUPPORTED, "too many register ops");
            if (hop[taken[t]].kind == 3) {
                // dense two-target op on the register pair (j1 < j2); the arm indexes the 4x4 matrix by
                // (bit j2 : bit j1), so swap the index bits when the matrix's first qubit sits on j1
                RegOp &r = P.regops[P.nregops++];
                const int r0 = reg_index_of_local[local_of(o.bitpos[0])], r1 = reg_index_of_local[local_of(o.bitpos[1])];
                if (r0 < 0 || r1 < 0) return set_error(QCS_ERR_UNSUPPORTED, "two-target op outside the register set");
                if (!split_masks(o.ctrl_mask, o.ctrl_mask, r))
                    return set_error(QCS_ERR_UNSUPPORTED, "too many distinct tile-level predicates");
                if (r.reg_care) return set_error(QCS_ERR_UNSUPPORTED, "two-target op with a register-bit control");
                const int j1 = r0 < r1 ? r0 : r1, j2 = r0 < r1 ? r1 : r0;
                static const int pair_index2[4][4] = {{-1, 0, 1, 2}, {-1, -1, 3, 4}, {-1, -1, -1, 5}, {-1, -1, -1, -1}};
                const bool swap_idx = r0 == j1;
                double cf2[32];
                for (int y = 0; y < 4; ++y)
                    for (int x = 0; x < 4; ++x) {
                        const int ys = swap_idx ? ((y & 1) << 1) | (y >> 1) : y, xs = swap_idx ? ((x & 1) << 1) | (x >> 1) : x;
                        cf2[2 * (4 * y + x)] = o.matrix[2 * (4 * ys + xs)];
                        cf2[2 * (4 * y + x) + 1] = o.matrix[2 * (4 * ys + xs) + 1];
                    }
                const bool pred = r.base_care != 0 || r.outer_slot != OUTER_SLOT_NONE;
                r.code = (uint16_t)((pred ? ARM_GEN2_P : ARM_GEN2) + pair_index2[j1][j2]);
                sw.has_dense2 = 1;
                const int at = push_reals(cf2, 32);
                if (at < 0) return set_error(QCS_ERR_UNSUPPORTED, "fused pass matrix pool exhausted");
                r.mat = at * (int)csize;
                ++sw.count;
            } else if (hop[taken[t]].kind == 0) {
                RegOp &r = P.regops[P.nregops++];
                const int j = reg_index_of_local[local_of(o.bitpos[0])];
                const double *M = o.matrix;   // m00 m01 m10 m11 as (re, im)
                if (!split_masks(o.ctrl_mask, o.ctrl_mask, r))
                    return set_error(QCS_ERR_UNSUPPORTED, "too many distinct tile-level predicates");
                // the cheapest exact form of the matrix (arm table: gen_regops.py)
                const int nrc = __builtin_popcount(r.reg_care);
                const bool free_op = o.ctrl_mask == 0;
                const bool pred = r.base_care != 0 || r.outer_slot != OUTER_SLOT_NONE;
                const bool real = M[1] == 0.0 && M[3] == 0.0 && M[5] == 0.0 && M[7] == 0.0;
                const bool imag = M[1] == 0.0 && M[7] == 0.0 && M[2] == 0.0 && M[4] == 0.0;   // real diagonal, imaginary off-diagonal
                const bool is_x = real && M[0] == 0.0 && M[6] == 0.0 && M[2] == 1.0 && M[4] == 1.0;
                double cf[8];
                int ncf = 0;
                if (nrc == 1 && is_x && !pred) {
                    r.code = (uint16_t)cx_arm(j, __builtin_ctz(r.reg_care));
                } else if (nrc >= 1) {
                    r.code = (uint16_t)(ARM_GENM + j);
                    for (int e = 0; e < 8; ++e) cf[e] = M[e];
                    ncf = 8;
                } else if (free_op && real && M[0] != 0.0 && M[0] == M[2] && M[0] == M[4] && M[0] == -M[6]) {
                    r.code = (uint16_t)(ARM_HAD + j);
                    gphase_mul(M[0], 0.0);
                } else if (free_op && real && M[0] != 0.0 && M[0] == M[6] && fabs(M[2]) <= 2.0 * fabs(M[0]) &&
                           fabs(M[4]) <= 2.0 * fabs(M[0])) {
                    r.code = (uint16_t)(ARM_REALU + j);
                    cf[0] = M[2] / M[0]; cf[1] = M[4] / M[0]; cf[2] = 1.0 - cf[0] * cf[1]; ncf = 3;
                    gphase_mul(M[0], 0.0);
                } else if (free_op && imag && M[0] != 0.0 && M[0] == M[6] && fabs(M[3]) <= 2.0 * fabs(M[0]) &&
                           fabs(M[5]) <= 2.0 * fabs(M[0])) {
                    // out0 = a0 + i b a1, out1 = a1 + i g a0 (b = Im m01 / m00, g = Im m10 / m00):
                    // (re a0, im a1) -> (u - b v, v + g u)
                    r.code = (uint16_t)(ARM_IMAGU + j);
                    cf[0] = -M[3] / M[0]; cf[1] = M[5] / M[0]; cf[2] = 1.0 - cf[0] * cf[1]; ncf = 3;
                    gphase_mul(M[0], 0.0);
                } else if (real) {
                    r.code = (uint16_t)((pred ? ARM_REAL_P : ARM_REAL) + j);
                    cf[0] = M[0]; cf[1] = M[2]; cf[2] = M[4]; cf[3] = M[6]; ncf = 4;
                } else if (imag) {
                    r.code = (uint16_t)((pred ? ARM_IMAG_P : ARM_IMAG) + j);
                    cf[0] = M[0]; cf[1] = -M[3]; cf[2] = M[5]; cf[3] = M[6]; ncf = 4;
                } else {
                    r.code = (uint16_t)((pred ? ARM_GEN_P : ARM_GEN) + j);
                    for (int e = 0; e < 8; ++e) cf[e] = M[e];
                    ncf = 8;
                }
                r.mat = 0;
                if (ncf) {
                    const int at = push_reals(cf, ncf);
                    if (at < 0) return set_error(QCS_ERR_UNSUPPORTED, "fused pass matrix pool exhausted");
                    r.mat = at * (int)csize;          // byte offset into the pool
                }
                ++sw.count;
            } else {
                // diagonal on k bits.  An uncontrolled diagonal is d[0] * diag(1, d[x]/d[0], ...): the common
                // factor goes to the pass-level phase, and patterns whose entry is exactly 1 cost nothing.
                double d0r = 1.0, d0i = 0.0;
                if (o.ctrl_mask == 0 && !(o.matrix[0] == 1.0 && o.matrix[1] == 0.0) &&
                    (o.matrix[0] != 0.0 || o.matrix[1] != 0.0)) {
                    d0r = o.matrix[0]; d0i = o.matrix[1];
                    gphase_mul(d0r, d0i);
                }
                const double d0n = d0r * d0r + d0i * d0i;
                for (int x = 0; x < (1 << o.k); ++x) {
                    double re = o.matrix[2 * x], im = o.matrix[2 * x + 1];
                    if (d0r != 1.0 || d0i != 0.0) {
                        if (x == 0) continue;
                        const double qr = (re * d0r + im * d0i) / d0n, qi = (im * d0r - re * d0i) / d0n;   // d[x] / d[0]
                        re = qr; im = qi;
                    }
                    if (re == 1.0 && im == 0.0) continue;
                    if (P.nregops >= Prog::MAX_REGOPS) return set_error(QCS_ERR_UNSUPPORTED, "too many register ops");
                    RegOp &r = P.regops[P.nregops++];
                    uint64_t care = o.ctrl_mask, val = o.ctrl_mask;
                    for (int j = 0; j < o.k; ++j) {
                        care |= 1ull << o.bitpos[j];
                        if ((x >> (o.k - 1 - j)) & 1) val |= 1ull << o.bitpos[j];
                    }
                    if (!split_masks(care, val, r))
                        return set_error(QCS_ERR_UNSUPPORTED, "too many distinct tile-level predicates");
                    const int nrc = __builtin_popcount(r.reg_care);
                    const bool pred = r.base_care != 0 || r.outer_slot != OUTER_SLOT_NONE;
                    const bool all_ones = r.reg_val == r.reg_care, is_real = im == 0.0;
                    static const int pair_index[4][4] = {{-1, 0, 1, 2}, {-1, -1, 3, 4}, {-1, -1, -1, 5}, {-1, -1, -1, -1}};
                    double cf[4] = {re, -im, im, re};       // the factor as a real 2x2 map on (re, im)
                    int ncf = 4;
                    if (nrc == 0) {
                        r.code = (uint16_t)ARM_PHT;
                        sw.thread_phase = 1;
                    } else if (nrc == 1 && all_ones) {
                        const int jj = __builtin_ctz(r.reg_care);
                        r.code = (uint16_t)((is_real ? (pred ? ARM_SGN1_P : ARM_SGN1) : (pred ? ARM_PH1_P : ARM_PH1)) + jj);
                        if (is_real) ncf = 1;
                    } else if (nrc == 2 && all_ones) {
                        const int j1 = __builtin_ctz(r.reg_care), j2 = 31 - __builtin_clz((unsigned)r.reg_care);
                        const int pi = pair_index[j1][j2];
                        r.code = (uint16_t)((is_real ? (pred ? ARM_SGN2_P : ARM_SGN2) : (pred ? ARM_PH2_P : ARM_PH2)) + pi);
                        if (is_real) ncf = 1;
                    } else {
                        r.code = (uint16_t)ARM_PHM;
                    }
                    const int at = push_reals(cf, ncf);
                    if (at < 0) return set_error(QCS_ERR_UNSUPPORTED, "fused pass matrix pool exhausted");
                    r.mat = at * (int)csize;
                    ++sw.count;
                }
            }
        }
        {
            const uint32_t lo = (uint32_t)sw.rpos[0] | ((uint32_t)sw.rpos[1] << 4) | ((uint32_t)sw.rpos[2] << 8) |
                                ((uint32_t)sw.rpos[3] << 12) | ((uint32_t)sw.hrank << 16) | ((uint32_t)(sw.tc_slot + 1) << 20) |
                                ((sw.count > 0 ? 1u : 0u) | (sw.has_dense2 ? 2u : 0u) | (sw.thread_phase ? 4u : 0u)) << 24;
            sw.packed = (uint64_t)lo | ((uint64_t)(uint32_t)sw.first << 32);
        }
        if (sw.count > 0 || tc_sweep) {    // (a run of exact identities produces no work)
            if (P.nregops >= Prog::MAX_REGOPS) return set_error(QCS_ERR_UNSUPPORTED, "too many register ops");
            RegOp &endop = P.regops[P.nregops++];          // terminates the interpreter loop of this sweep
            endop.code = (uint16_t)ARM_END; endop.reg_care = endop.reg_val = 0; endop.base_care = endop.base_val = 0;
            endop.mat = 0; endop.outer_slot = OUTER_SLOT_NONE;
            ++P.nsweeps;
        }
        for (int i = 0; i < nrest; ++i) pending[i] = rest[i];
        npend = nrest;
    }
    P.mat_bytes = (int32_t)(used * csize);
    // 2: complex factor, 1: real factor (lazy norm and / or deferred real scalars), 0: none
    P.scale_mode = P.gphase[1] != 0.0 ? 2 : ((norm_in || P.gphase[0] != 1.0) ? 1 : 0);
    if (TCPROG && P.n_tc == 0) return set_error(QCS_ERR_UNSUPPORTED, "no dense block in this pass");
    if (stats && tune_int("QCS_PLAN_DUMP", 0)) {      // diagnostics: the register bits and op count of every sweep
        for (int i = 0; i < P.nsweeps; ++i) {
            const Sweep &w = P.sweeps[i];
            if (w.kind == SWEEP_REG)
                std::fprintf(stderr, "sweep %d regs %d %d %d %d ops %d block slot %d (thread-bit rank %d)\n", i, w.rpos[0], w.rpos[1],
                             w.rpos[2], w.rpos[3], w.count, w.tc_slot, w.hrank);
            else std::fprintf(stderr, "sweep %d dense\n", i);
        }
    }
    if (stats && g_block_dump) {        // qcs_plan_blocks: the images and geometry of the tensor-core sweeps
        std::memcpy(g_block_dump, tc_images, (size_t)P.n_tc * TC_SLOT_BYTES);
        int w = 0;
        for (int i = 0; i < P.nsweeps; ++i) {
            const Sweep &sp = P.sweeps[i];
            if (sp.kind != SWEEP_REG || sp.tc_slot < 0) continue;
            int32_t *o = g_block_info + 8 * (w++);
            auto flat_of = [&](int lb) { return lb < P.L ? lb : P.hbits[lb - P.L]; };
            for (int j = 0; j < 4; ++j) o[j] = flat_of(sp.rpos[j]);
            // the thread bit: the tile bit of rank hrank among the non-register bits
            int rank = 0, hb = -1;
            for (int b = 0; b < m; ++b) {
                bool isreg = false;
                for (int j = 0; j < 4; ++j) isreg |= sp.rpos[j] == b;
                if (isreg) continue;
                if (rank == sp.hrank) { hb = b; break; }
                ++rank;
            }
            o[4] = flat_of(hb); o[5] = sp.count; o[6] = sp.tc_slot; o[7] = i;
        }
    }
    if (stats) {       // planning only (qcs_plan_stats): tile bits, sweeps, register ops, per-arm counts
        stats[0] = m; stats[1] = P.nsweeps; stats[2] = P.nregops; stats[3] = P.ngen; stats[4] = P.nouter;
        stats[5] = P.scale_mode;
        stats[7] = P.n_tc | (tc_block_ops << 8);       // tensor-core sweeps, input ops multiplied into their blocks
        for (int i = 0; i < 108; ++i) stats[8 + i] = 0;
        // the input ops that run in the last sweep (a short tail sweep is a candidate for deferral to the next
        // pass): count in stats[6], the first 12 indices in stats[116..127], -1 terminated
        stats[6] = n_last_sweep;
        for (int i = 0; i < 12; ++i) stats[116 + i] = i < n_last_sweep ? last_sweep_ops[i] : -1;
        for (int i = 0; i < P.nregops; ++i)
            if (P.regops[i].code < 108) stats[8 + P.regops[i].code]++;
        return QCS_OK;
    }
    const DeviceInfo &dev = device_info();
    if constexpr (TCPROG) {
        if (!ws) return set_error(QCS_ERR_UNSUPPORTED, "tensor-core pass needs the workspace");
        const int rc = tc_upload(tc_images, P.n_tc, ws, stream);
        if (rc != QCS_OK) return rc;
        P.tc_mats = reinterpret_cast<const unsigned char *>(ws) + WS_TC_OFFSET;
        // the fp16 operands are ranged with the norm of the state: the caller's lazy norm, its bound, or one reduction here
        P.tc_norm = bound_in ? bound_in : norm_in;
        if (!P.tc_norm) {
            double *scratch = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(ws) + WS_TC_NORM_OFFSET);
            const int rn = misc_abs2(nullptr, src, n, dtype, nullptr, scratch, ws, stream);
            if (rn != QCS_OK) return rn;
            P.tc_norm = scratch;
        }
        return launch_tc(P, dev, stream);
    } else {
        const int mode = tile_mode();
        return c64 ? launch_tile<float, Prog>(P, dev, mode, stream) : launch_tile<double, Prog>(P, dev, mode, stream);
    }
}

int tile_apply(void *dst, const void *src, int n, int dtype, const qcs_op *ops, int nops, const double *norm_in,
               double *norm_out, const double *bound_in, void *ws, cudaStream_t stream) {
    if (nops < 1) return set_error(QCS_ERR_ARG, "empty op list");
    // Alternative for a pass that is one one-qubit gate: stream straight through registers with strided 128-bit
    // loads, no shared-memory tile (qcs_misc.cu: gate1_stream_kernel).  Parity-green but measured SLOWER than the
    // cp.async tile ring (3.07 vs 2.92 ms at 30 qubits complex64, 1.54 vs 1.48 ms at 28 qubits complex128), so
    // it is off by default and stays selectable (QCS_STREAM1=1) as the recorded evidence for that choice.
    if (nops == 1 && ops[0].kind == QCS_OP_DENSE && ops[0].k == 1 && n >= 12 && tune_int("QCS_STREAM1", 0) &&
        ops[0].bitpos[0] >= 0 && ops[0].bitpos[0] < n && !(ops[0].ctrl_mask & (1ull << ops[0].bitpos[0])) &&
        (n >= 64 || !(ops[0].ctrl_mask >> n)))
        return misc_gate1(dst, src, n, dtype, ops[0].matrix, ops[0].bitpos[0], ops[0].ctrl_mask, norm_in, norm_out, ws,
                          stream);
    if (nops <= 2) {
        const int rc = build_and_launch<SmallProg>(dst, src, n, dtype, ops, nops, norm_in, norm_out, ws, stream);
        if (rc != QCS_ERR_UNSUPPORTED) return rc;
    }
    if (dtype == QCS_C64 && ws && n >= TILE_MIN_BITS_C64 && nops >= tune_int("QCS_TC_MIN_OPS", 3)) {
        // passes with dense blocks run on the tensor-core kernel (the planner says "unsupported" when it finds none)
        const int rc = build_and_launch<TcProg>(dst, src, n, dtype, ops, nops, norm_in, norm_out, ws, stream, nullptr, bound_in);
        if (rc != QCS_ERR_UNSUPPORTED) return rc;
    }
    return build_and_launch<LargeProg>(dst, src, n, dtype, ops, nops, norm_in, norm_out, ws, stream);
}

int tile_plan_stats(int n, int dtype, const qcs_op *ops, int nops, int32_t *stats) {
    if (nops < 1) return set_error(QCS_ERR_ARG, "empty op list");
    if (dtype == QCS_C64 && n >= TILE_MIN_BITS_C64 && nops >= tune_int("QCS_TC_MIN_OPS", 3)) {
        const int rc = build_and_launch<TcProg>(nullptr, nullptr, n, dtype, ops, nops, nullptr, nullptr, nullptr, nullptr, stats);
        if (rc != QCS_ERR_UNSUPPORTED) return rc;
    }
    return build_and_launch<LargeProg>(nullptr, nullptr, n, dtype, ops, nops, nullptr, nullptr, nullptr, nullptr, stats);
}

int tile_plan_order(int n, int dtype, const qcs_op *ops, int nops, int32_t *stats, int32_t *sweep_of_op) {
    if (nops < 1) return set_error(QCS_ERR_ARG, "empty op list");
    for (int i = 0; i < nops; ++i) sweep_of_op[i] = -1;
    g_sweep_of_op = sweep_of_op;
    const int rc = tile_plan_stats(n, dtype, ops, nops, stats);
    g_sweep_of_op = nullptr;
    return rc;
}

int tile_plan_blocks(int n, const qcs_op *ops, int nops, int32_t *stats, void *images, int32_t *info) {
    if (nops < 1) return set_error(QCS_ERR_ARG, "empty op list");
    g_block_dump = reinterpret_cast<unsigned char *>(images);
    g_block_info = info;
    const int rc = build_and_launch<TcProg>(nullptr, nullptr, n, QCS_C64, ops, nops, nullptr, nullptr, nullptr, nullptr, stats);
    g_block_dump = nullptr;
    g_block_info = nullptr;
    return rc;
}

}  // namespace qcs

#ifdef QCS_TC_DBG
// timing experiments only (scratch builds): the clock stamps of the last tensor-core launch
extern "C" int qcs_debug_tc_trace(unsigned long long *out, int max_pairs) {
    unsigned int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, qcs::g_tc_trace_n, sizeof n);
    if ((int)n > max_pairs) n = (unsigned)max_pairs;
    cudaMemcpyFromSymbol(out, qcs::g_tc_trace, (size_t)n * 16);
    return (int)n;
}
#endif
