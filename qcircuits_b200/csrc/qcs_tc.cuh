// Tensor-core sweeps of the tile engine (complex64, 12-bit tiles): tcgen05.mma kind::f16, accumulators in
// tensor memory.  Included by qcs_tile.cu after the register-sweep machinery.
//
// The contraction replaced is the same np.tensordot as everywhere (reference operators.py:261), for a run of
// gates that the planner has multiplied into ONE dense 2^5 x 2^5 complex block U acting on five tile bits.
//
// Mapping.  A tile holds 2^12 amplitudes = 128 rows x 32 amplitudes, the 32 spanned by the block's five bits.
//   A (M = 128 rows, K): the tile, in TENSOR MEMORY, fp16 packed two per 32-bit column -- lane = row; thread (row, h)
//       writes its 16 amplitudes as 8 columns of real parts, then 8 of imaginary parts, with ONE tcgen05.st.x16 per
//       image straight from the registers of the sweep (no second shared-memory image of the tile).
//       Two images: "hi" = fp16(s a), "lo" = fp16(s a - hi); s = a power of two that maps the norm of the state to
//       [2^14, 2^15) -- every amplitude is at most the norm, so nothing overflows, and an amplitude below the fp16
//       normal range still carries an absolute error of only 2^-25 / s ~ 2^-40 of the norm (subnormals are exact to 2^-24).
//   B (N, K): the block matrix in SHARED MEMORY, fp16, K-major, no swizzle, rows 0..31 = Re U, rows 32..63 = Im U; a
//       "hi" and a "lo" image, built on the host from the float64 product of the gates (8 KB per block).
//   D (M = 128, N = 64, fp32) in tensor memory: columns [Re out_0..31 | Im out_0..31], read back with tcgen05.ld by the
//       same thread into the same registers, which then go back to the tile in shared memory (the tile stays ranged
//       for the whole pass -- register ops are linear -- and 1 / s is folded into the factor of the store-out).
//   Re out = Re a . Re U^T - Im a . Im U^T ,  Im out = Re a . Im U^T + Im a . Re U^T :
//       per split term 2 instructions M128 N64 K16 (Re a against both halves of B) + 2 M128 N32 K16 with the b-negate
//       flag (Im a against Im U, into the Re columns) + 2 M128 N32 K16 (Im a against Re U, into the Im columns);
//       three terms, small ones first: lo.hi + hi.lo + hi.hi (the dropped lo.lo is 2^-24).  18 instructions per tile
//       and sweep, half the tensor-pipe time and half the shared-memory operand reads of the tf32 form it replaced.
// One elected lane issues the instructions after the group's barrier; completion comes back through
// tcgen05.commit on an mbarrier.  A CTA is four groups of 128 threads (or two of 256) that stream different tiles, so
// that the groups' shared-memory and split work overlaps each other's MMAs (tile_kernel_tc below).
#pragma once
#include <cuda_fp16.h>

#ifdef QCS_TC_DBG
// timing experiments only: clock stamps of group 0, CTA 0 (thread 0) for a few steady-state tiles
__device__ unsigned long long g_tc_trace[4096];
__device__ unsigned int g_tc_trace_n;
#define TC_STAMP(ctx, tag) do { if ((ctx).trace_on && (ctx).trace_i < 2040u) { \
    g_tc_trace[2 * (ctx).trace_i] = (unsigned long long)(tag); g_tc_trace[2 * (ctx).trace_i + 1] = clock64(); ++(ctx).trace_i; } } while (0)
#else
#define TC_STAMP(ctx, tag) do { } while (0)
#endif

namespace tc {

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 consecutive columns of this thread's lane
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                 "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T
__device__ __forceinline__ void mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle; core matrix = 8 rows x 16 bytes stored contiguously,
// LBO = bytes between the two K chunks of an instruction, SBO = bytes between 8-row groups; bit 46 = sm_100 version
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = f32 (bit 4), A and B = f16 (0 at bits 7 and 10), b-negate bit 14,
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int b_neg) {
    return (1u << 4) | ((uint32_t)b_neg << 14) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// bounded wait (a lost completion traps instead of hanging the GPU)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    asm volatile("trap;");
}
template <int GT>
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(GT) : "memory"); }

// tensor-memory columns of one group (base = (512 / groups) * group): D, then the hi and lo images of A.  In an image the thread
// with top block bit h owns columns 16 h .. 16 h + 15: 8 of packed real parts (K chunk "Re, h"), 8 of imaginary parts
constexpr uint32_t COL_D = 0, COL_AHI = 64, COL_ALO = 96;
constexpr uint32_t B_LBO = 128, B_SBO = 512, B_IM_ROWS = 4 * B_SBO;      // rows 32..63 start 4 row groups in

// all 18 instructions of one tile and sweep; `bimg` = shared-memory address of the slot's images (hi, then lo)
__device__ __forceinline__ void issue_block(uint32_t tbase, uint32_t bimg) {
    constexpr uint32_t I64 = idesc_f16(128, 64, 0), I32 = idesc_f16(128, 32, 0), I32N = idesc_f16(128, 32, 1);
    const uint64_t hi = smem_desc(bimg, B_LBO, B_SBO), lo = smem_desc(bimg + TC_HALF_BYTES, B_LBO, B_SBO);
    constexpr uint64_t KSTEP = (2 * B_LBO) >> 4, IM = B_IM_ROWS >> 4;     // 16 columns of U per instruction
    const uint32_t d = tbase + COL_D;
#pragma unroll
    for (int term = 0; term < 3; ++term) {
        const uint32_t a = tbase + (term == 0 ? COL_ALO : COL_AHI);
        const uint64_t b = term == 1 ? lo : hi;
#pragma unroll
        for (int h = 0; h < 2; ++h) mma(d, a + 16 * h, b + h * KSTEP, I64, (term | h) ? 1u : 0u);
#pragma unroll
        for (int h = 0; h < 2; ++h) mma(d, a + 16 * h + 8, b + IM + h * KSTEP, I32N, 1u);
#pragma unroll
        for (int h = 0; h < 2; ++h) mma(d + 32, a + 16 * h + 8, b + h * KSTEP, I32, 1u);
    }
}

struct Ctx {
    uint32_t tbase;        // tensor-memory base of this group
    uint32_t lane_addr;    // ... with this warp's lane quarter
    uint32_t bimg0;        // shared-memory address of block-matrix slot 0
    uint64_t *bar;         // this group's completion barrier
    uint32_t phase;
    float scale, inv_scale;   // power of two that ranges the fp16 operands, and its inverse
    int group, gtid;
#ifdef QCS_TC_DBG
    bool trace_on;
    uint32_t trace_i;
    uint32_t dbg;          // timing experiments only (scratch builds): 1 no MMA, 2 no split / tcgen05.st, 4 no tcgen05.ld, 8 no tile LDS / STS, 16 trivial addresses, 32 every warp polls, 64 no commit / wait (with 1), 128 no register ops, 256 no barriers around the MMA
#endif
};
#ifdef QCS_TC_DBG
#define TC_DBG(ctx, bit) ((ctx).dbg & (bit))
#else
#define TC_DBG(ctx, bit) false
#endif

// two-term fp16 split of a pair of (scaled) reals: hi = fp16(v) (round to nearest), lo = fp16(v - hi); packed with
// the first value in the low half (K even)
__device__ __forceinline__ void split2(float v0, float v1, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - f.x, v1 - f.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// One sweep of the tensor-core kernel: register ops (the generated interpreter), then -- if the sweep has a block --
// the dense 2^5 x 2^5 matrix on tensor cores.  A row of the tile (32 amplitudes) is two "virtual threads" (rho, h), each
// owning the 16 amplitudes whose top block bit is h: with 256-thread groups (HALVES = 1) they are two threads, with
// 128-thread groups (HALVES = 2) thread rho does both, one after the other.
template <typename Layout, int HALVES>
__device__ __forceinline__ void sweep(float2 *tile, const Sweep &sw, const uint32_t *tab, const RegOp *rops, const float2 *pool,
                                      uint64_t outer_ok, Ctx &ctx, bool first_touch) {
    constexpr int NX = 16;
    constexpr int GT = 256 / HALVES;
    // the record as one 8-byte broadcast load, the four register-bit strides as one 16-byte load
    TC_STAMP(ctx, 1);
    uint32_t plo, phi;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(plo), "=r"(phi) : "r"(smem_u32(&sw.packed)) : "memory");
    const uint32_t flags = plo >> 24, first = phi;
    const int tc_slot = (int)((plo >> 20) & 15u) - 1;
    const bool VEC = (plo & 15u) == 0;
    uint32_t st4[4];
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(st4[0]), "=r"(st4[1]), "=r"(st4[2]), "=r"(st4[3]) : "r"(smem_u32(&sw.stride[0])) : "memory");
    // the virtual thread's first amplitude (tile-local index, shared-memory byte offset) comes from the per-pass table
    // built at kernel start -- it is the same for every tile
    auto addresses = [&](uint32_t vt, uint32_t (&addr)[NX]) -> uint32_t {
        const uint32_t ent = tab[vt];
        addr[0] = smem_u32(tile) + (ent & 0xFFFFu);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int x = 0; x < (1 << j); ++x) addr[x | (1 << j)] = addr[x] + st4[j];
        }
        return ent >> 16;
    };
    auto store_tile = [&](const float2 (&a)[NX], const uint32_t (&addr)[NX]) {
        if (TC_DBG(ctx, 8u)) {
            if (a[3].x == 12345.678f) tile[0] = a[5];       // keep the values alive
        } else if (VEC) {
#pragma unroll
            for (int x = 0; x < NX; x += 2)
                asm volatile("st.shared.v4.f32 [%4], {%0, %1, %2, %3};" ::"f"(a[x].x), "f"(a[x].y), "f"(a[x + 1].x), "f"(a[x + 1].y),
                             "r"(addr[x])
                             : "memory");
        } else {
#pragma unroll
            for (int x = 0; x < NX; ++x)
                asm volatile("st.shared.v2.f32 [%2], {%0, %1};" ::"f"(a[x].x), "f"(a[x].y), "r"(addr[x]) : "memory");
        }
    };
#pragma unroll 1
    for (int hh = 0; hh < HALVES; ++hh) {
        const uint32_t vt = (uint32_t)ctx.gtid + (uint32_t)hh * GT, h = vt >> 7;
        uint32_t addr[NX];
        const uint32_t base = addresses(vt, addr);
        float2 a[NX];
        if (TC_DBG(ctx, 8u)) {
#pragma unroll
            for (int x = 0; x < NX; ++x) a[x] = make_float2((float)x, (float)ctx.gtid);
        } else if (VEC) {
#pragma unroll
            for (int x = 0; x < NX; x += 2)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(a[x].x), "=f"(a[x].y), "=f"(a[x + 1].x), "=f"(a[x + 1].y) : "r"(addr[x]) : "memory");
        } else {
#pragma unroll
            for (int x = 0; x < NX; ++x)
                asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a[x].x), "=f"(a[x].y) : "r"(addr[x]) : "memory");
        }
        if (first_touch) {     // the tile stays ranged (times s, a power of two) until the store-out folds 1 / s back in
#pragma unroll
            for (int x = 0; x < NX; ++x) { a[x].x *= ctx.scale; a[x].y *= ctx.scale; }
        }
        TC_STAMP(ctx, 2);      // tile in registers (loads issued)
        if ((flags & 1u) && !TC_DBG(ctx, 128u)) {
            float2 ph = make_float2(1.0f, 0.0f);
            if (flags & 2u) run_sweep_ops_dense2(a, ph, smem_u32(rops + first), smem_u32(pool), base, outer_ok);
            else run_sweep_ops(a, ph, smem_u32(rops + first), smem_u32(pool), base, outer_ok);
            if (flags & 4u) {
#pragma unroll
                for (int x = 0; x < NX; ++x) a[x] = cmul(ph, a[x]);
            }
        }
        TC_STAMP(ctx, 3);      // register ops done
        if (tc_slot >= 0) {
            if (!TC_DBG(ctx, 2u)) {
                const uint32_t col = ctx.lane_addr + 16u * h;
                uint32_t hi[NX], lo[NX];       // 8 columns of packed real parts, then 8 of imaginary parts
#pragma unroll
                for (int x = 0; x < NX; x += 2) {
                    split2(a[x].x, a[x + 1].x, hi[x >> 1], lo[x >> 1]);
                    split2(a[x].y, a[x + 1].y, hi[8 + (x >> 1)], lo[8 + (x >> 1)]);
                }
                st16(col + COL_AHI, hi);
                st16(col + COL_ALO, lo);
            }
            TC_STAMP(ctx, 4);  // split + tcgen05.st issued
        } else {
            store_tile(a, addr);
        }
    }
    if (tc_slot < 0) return;
    wait_st();
    fence_before();
    TC_STAMP(ctx, 5);  // stores complete
    if (!TC_DBG(ctx, 256u)) group_sync<GT>(ctx.group);
    TC_STAMP(ctx, 6);  // barrier passed
    if ((ctx.gtid >> 5) == 0 && !TC_DBG(ctx, 64u)) {
        // only this warp watches the completion barrier; the others sleep in the group barrier below
        // (a spinning try_wait loop in every warp took issue slots from the other groups' register work)
        fence_after();
        if (elect_one()) {
            if (!TC_DBG(ctx, 1u)) issue_block(ctx.tbase, ctx.bimg0 + (uint32_t)tc_slot * (uint32_t)TC_SLOT_BYTES);
            commit(ctx.bar);
        }
        __syncwarp();
        TC_STAMP(ctx, 7);  // MMAs issued
        mbar_wait_bounded(ctx.bar, ctx.phase);
        TC_STAMP(ctx, 8);  // MMAs complete
    }
    if (!TC_DBG(ctx, 64u)) ctx.phase ^= 1u;
    if (TC_DBG(ctx, 32u)) { if ((ctx.gtid >> 5) != 0) mbar_wait_bounded(ctx.bar, ctx.phase ^ 1u); }   // every warp watches the barrier itself
    else if (!TC_DBG(ctx, 256u)) group_sync<GT>(ctx.group);
    fence_after();
    TC_STAMP(ctx, 9);  // second barrier passed
#pragma unroll 1
    for (int hh = 0; hh < HALVES; ++hh) {
        const uint32_t vt = (uint32_t)ctx.gtid + (uint32_t)hh * GT, h = vt >> 7;
        const uint32_t col = ctx.lane_addr + 16u * h;
        uint32_t addr[NX];
        (void)addresses(vt, addr);
        float2 a[NX];
        if (!TC_DBG(ctx, 4u)) {
            uint32_t re[NX], im[NX];
            ld16(col + COL_D, re);
            ld16(col + COL_D + 32u, im);
            wait_ld();
#pragma unroll
            for (int x = 0; x < NX; ++x) { a[x].x = __uint_as_float(re[x]); a[x].y = __uint_as_float(im[x]); }
        } else {
#pragma unroll
            for (int x = 0; x < NX; ++x) a[x] = make_float2((float)x, (float)ctx.gtid);
        }
        TC_STAMP(ctx, 10); // D in registers
        store_tile(a, addr);
    }
    fence_before();        // orders these tensor-memory reads before the next sweep's barrier + MMA
    TC_STAMP(ctx, 11);     // tile stores issued
}

}  // namespace tc

// Persistent CTAs of 512 threads in NG = 512 / GT groups that stream different tiles (group g of CTA b is "virtual CTA"
// NG b + g of the tile loop of tile_kernel_pipe: same cp.async staging, same store epilogue), so that one group's
// shared-memory and split work overlaps the others' MMAs.  GT = 256: two groups, a 2-stage ring each (the next tile is
// prefetched).  GT = 128: four groups with ONE stage each -- a group waits for its own tile, the other three keep the
// SM busy -- and every thread owns a whole row of the tile (two halves of 16 amplitudes, one after the other).
template <typename Prog, int GT>
__global__ void __launch_bounds__(512, 1) tile_kernel_tc(const __grid_constant__ Prog P) {
    using C = float2;
    using Layout = PaddedLayout<C>;
    constexpr int MT = TILE_MIN_BITS_C64, NT = GT, NG = 512 / GT, HALVES = 256 / GT, VSHIFT = 1;
    constexpr int NKB = GT == 256 ? 3 : 4, NK = 1 << NKB;       // vectors per thread of a tile's staging and store-out
    constexpr int S = GT == 256 ? 2 : 1;
    constexpr uint32_t TCOLS = 512 / NG;
    extern __shared__ __align__(128) unsigned char smem_all[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t mma_bar[4];
    const int tid = threadIdx.x, g = tid / GT, gtid = tid % GT;
    ProgView pv;
    const uint64_t *chunk_off;
    const uint32_t prog_bytes = stage_program(smem_all, P, pv, chunk_off, tid, 512);
    // flat offsets (in bytes) of the top tile bits, in registers: vector k of a thread sits at the sum of those whose
    // bit of k is set -- no table look-ups in the load and store loops
    uint64_t top_off[NKB];
    for (int j = 0; j < NKB; ++j) {
        const int i = TILE_MIN_BITS_C64 - NKB + j;
        top_off[j] = (1ull << (i < P.L ? i : P.hbits[i - P.L])) * sizeof(float2);
    }
    (void)chunk_off;
    // per-sweep addressing table of the 256 virtual threads: (tile-local index of the first amplitude) << 16 | its byte
    // offset in a stage.  Virtual thread (rho, h) owns half h of row rho: h is the block's thread bit (rank `hrank` among
    // the 8 non-register tile bits), the 7 bits of rho are the other non-register bits -- in an order chosen so that
    // the lanes of one shared-memory wavefront (16 lanes of 8-byte accesses, or 8 lanes of 16-byte accesses when tile bit
    // 0 is a register bit) fall on different banks: in the padded layout tile bit b moves an amplitude by 8 bytes (b =
    // 0) or 16 / 32 / 64 bytes mod 128 (b = 1, 2, 3 mod 3), so the low lane bits take one tile bit of each class.  Rows
    // are independent in the MMA and the register ops see the index through `base`, so any order is legal.
    uint32_t *swtab = reinterpret_cast<uint32_t *>(smem_all + prog_bytes);
    for (uint32_t i = tid; i < (uint32_t)P.nsweeps * 256u; i += 512u) {
        const Sweep &sw = P.sweeps[i >> 8];
        if (sw.kind != SWEEP_REG) continue;
        uint32_t regmask = 0;
        for (int j = 0; j < 4; ++j) regmask |= 1u << sw.rpos[j];
        int tb[8], nt = 0;                             // the non-register tile bits, ascending
        for (int b = 0; b < MT; ++b)
            if (!((regmask >> b) & 1u)) tb[nt++] = b;
        const int hbit = tb[sw.hrank];
        int order[7], no = 0;
        uint32_t used = 1u << hbit;
        if (!(regmask & 1u) && hbit != 0) { order[no++] = 0; used |= 1u; }          // 8-byte accesses: bit 0 first
        for (int cls = 0; cls < 3; ++cls)                                             // then one bit of each class
            for (int k = 0; k < 8; ++k) {
                const int b = tb[k];
                if (b >= 1 && !((used >> b) & 1u) && (b - 1) % 3 == cls) { order[no++] = b; used |= 1u << b; break; }
            }
        for (int k = 0; k < 8; ++k)                                                   // the rest, ascending
            if (!((used >> tb[k]) & 1u)) { order[no++] = tb[k]; used |= 1u << tb[k]; }
        const uint32_t t = i & 255u, rho = t & 127u, hh = t >> 7;
        uint32_t base = hh << hbit;
        for (int j = 0; j < 7; ++j) base |= ((rho >> j) & 1u) << order[j];
        swtab[i] = (base << 16) | (Layout::amp(base) * (uint32_t)sizeof(float2));
    }
    // block-matrix images
    unsigned char *bimg = smem_all + prog_bytes + (size_t)P.nsweeps * 256u * sizeof(uint32_t);
    {
        const uint4 *srcm = reinterpret_cast<const uint4 *>(P.tc_mats);
        uint4 *dstm = reinterpret_cast<uint4 *>(bimg);
        const uint32_t nvec = (uint32_t)P.n_tc * (TC_SLOT_BYTES / 16);
        for (uint32_t i = tid; i < nvec; i += 512u) dstm[i] = srcm[i];
    }
    unsigned char *ring = bimg + (size_t)P.n_tc * TC_SLOT_BYTES + (size_t)g * S * Layout::bytes(MT);
    if (tid < 32) tc::tmem_alloc(&tmem_slot, 512);
    if (tid == 32) {
        for (int i = 0; i < 4; ++i) mbar_init(&mma_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the images were written through the generic proxy
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    tc::Ctx ctx;
    ctx.group = g; ctx.gtid = gtid; ctx.phase = 0; ctx.bar = &mma_bar[g];
    ctx.tbase = tmem_slot + TCOLS * (uint32_t)g;
    ctx.lane_addr = ctx.tbase + ((uint32_t)(((gtid >> 5) & 3) * 32) << 16);
    ctx.bimg0 = smem_u32(bimg);
    {
        // s = 2^(15 - e) with sqrt(norm^2) = m 2^e, m in [0.5, 1): every |s a| < 2^15 (fp16 holds 65504)
        int e = 0;
        const double n2 = P.tc_norm ? *P.tc_norm : 1.0;
        if (n2 > 0.0 && n2 < 1e300) (void)frexp(sqrt(n2), &e);
        e = e < -100 ? -100 : (e > 100 ? 100 : e);
        ctx.scale = exp2f((float)(15 - e));
        ctx.inv_scale = exp2f((float)(e - 15));
    }
#ifdef QCS_TC_DBG
    ctx.dbg = (uint32_t)P.tc_pad;
    ctx.trace_on = false; ctx.trace_i = 0;
#endif

    const int L = P.L;
    const uint32_t stage_bytes = (uint32_t)Layout::bytes(MT);
    const uint32_t svec_tid = Layout::vec((uint32_t)gtid);
    // the store-out always multiplies: lazy norm / global phase of the pass, and 1 / s of the fp16 ranging
    const int scale_mode = P.scale_mode == 2 ? 2 : 1;
    float fr = ctx.inv_scale, fi = 0.0f;
    if (P.scale_mode) {
        const double s = load_inv_norm(P.norm_in) * (double)ctx.inv_scale;
        fr = (float)(s * P.gphase[0]); fi = (float)(s * P.gphase[1]);
    }
    uint64_t low_tid = 0;
    for (int b = 0; b < MT - NKB - VSHIFT; ++b)
        if ((gtid >> b) & 1) {
            const int i = b + VSHIFT;
            low_tid |= 1ull << (i < L ? i : P.hbits[i - L]);
        }
    const uint4 *src = reinterpret_cast<const uint4 *>(P.src) + (low_tid >> VSHIFT);
    uint4 *dst = reinterpret_cast<uint4 *>(P.dst) + (low_tid >> VSHIFT);
    double nacc[1] = {0.0};
    const uint64_t vcta = (uint64_t)NG * blockIdx.x + (uint64_t)g, vgrid = (uint64_t)NG * gridDim.x;
    const uint64_t my_tiles = (P.ntiles > vcta) ? (P.ntiles - 1 - vcta) / vgrid + 1 : 0;
    const float2 *pool = reinterpret_cast<const float2 *>(pv.pool);

    uint64_t next_base = 0;            // flat index of the tile whose copies were issued last (computed once per tile)
    auto issue_load = [&](uint64_t it, int st) {
        const uint64_t base = tile_base_of(vcta + it * vgrid, P);
        next_base = base;
        const uint32_t sp = smem_u32(ring + (size_t)st * stage_bytes) + svec_tid * 16u;
        // the vectors of a thread: one pointer + the top-bit offsets (NK - 1 additions, no per-vector index arithmetic)
        const unsigned char *gptr[NK];
        gptr[0] = reinterpret_cast<const unsigned char *>(src) + base * sizeof(float2);
#pragma unroll
        for (int j = 0; j < NKB; ++j) {
#pragma unroll
            for (int k = 0; k < (1 << j); ++k) gptr[k | (1 << j)] = gptr[k] + top_off[j];
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp + Layout::vec((uint32_t)(k * NT)) * 16u), "l"(gptr[k]) : "memory");
        }
    };
    if (my_tiles > 0) issue_load(0, 0);
    cp_async_commit();
    int st = 0;
    for (uint64_t it = 0; it < my_tiles; ++it) {
        const uint64_t base = next_base;
        const int parity = (int)(it & 1) + 2 * g;
#ifdef QCS_TC_DBG
        ctx.trace_on = tid == 0 && blockIdx.x == 0 && it >= 100 && it < 104;
#endif
        TC_STAMP(ctx, 20);
        eval_outer(pv, base, parity, gtid);
        cp_async_wait_pending(0);          // this thread's copies of tile `it` have landed
        TC_STAMP(ctx, 21);
        tc::group_sync<GT>(g);             // ... everyone's have, and the other stage is drained
        TC_STAMP(ctx, 22);
        // (issuing the next tile's copies later -- while the tensor core works on this tile's first block -- was measured
        // slower: 202 vs 182 ms on the headline circuit)
        if (S > 1) {
            if (it + 1 < my_tiles) issue_load(it + 1, st ^ 1);
            cp_async_commit();
        }
        unsigned char *stage = ring + (size_t)st * stage_bytes;
        const uint64_t outer_ok = (pv.nouter ? pv.outer_ok[parity] : 0ull) | (1ull << OUTER_SLOT_NONE);
        bool ranged = false;
        for (int s = 0; s < pv.nsweeps; ++s) {
            if (TC_DBG(ctx, 512u)) break;                          // (timing experiment: staging and store-out only)
            const Sweep &sw = pv.sweeps[s];
            if (sw.kind == SWEEP_REG) {
                tc::sweep<Layout, HALVES>(reinterpret_cast<C *>(stage), sw, swtab + s * 256, pv.regops, pool, outer_ok, ctx, !ranged);
                ranged = true;
            }
            else apply_generic<float, Layout>(reinterpret_cast<C *>(stage), MT, pv.gen[sw.first], pool, base, gtid, NT);
            tc::group_sync<GT>(g);
            TC_STAMP(ctx, 12);     // end-of-sweep barrier passed
        }
        TC_STAMP(ctx, 23);
        float tacc = 0.0f;
        const uint4 *sbase = reinterpret_cast<const uint4 *>(stage) + svec_tid;
        unsigned char *optr[NK];
        optr[0] = reinterpret_cast<unsigned char *>(dst) + base * sizeof(float2);
#pragma unroll
        for (int j = 0; j < NKB; ++j) {
#pragma unroll
            for (int k = 0; k < (1 << j); ++k) optr[k | (1 << j)] = optr[k] + top_off[j];
        }
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            uint4 val = sbase[Layout::vec((uint32_t)(k * NT))];
            if (scale_mode == 1) val = vec_scale<float>(val, fr);
            else if (scale_mode == 2) val = vec_cscale<float>(val, fr, fi);
            vec_norm_acc(tacc, val, 0.0f);
            *reinterpret_cast<uint4 *>(optr[k]) = val;
        }
        nacc[0] += (double)tacc;
        if (S > 1) st ^= 1;
        else {
            // one stage: the next tile is copied over this one as soon as it has been read out (a thread's copies land in
            // the slots the same thread has just read)
            if (it + 1 < my_tiles) issue_load(it + 1, 0);
            cp_async_commit();
        }
        TC_STAMP(ctx, 24);         // store-out issued
    }
    cp_async_wait_pending(0);
#ifdef QCS_TC_DBG
    if (tid == 0 && blockIdx.x == 0) g_tc_trace_n = ctx.trace_i;
#endif
    tc::fence_before();
    __syncthreads();
    if (tid < 32) tc::tmem_dealloc(tmem_slot, 512);
    if (P.norm_out) grid_sum_publish<1>(nacc, P.ws, P.norm_out);
}
