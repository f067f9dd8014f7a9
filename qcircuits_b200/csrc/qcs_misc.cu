// Streaming / reduction kernels around the gate pass: probabilities, marginals, collapse,
// tensor product, linear combinations, bit permutation, density-operator helpers, and the
// dense fallback for operators wider than the tile engine's 4 qubits.
// Each kernel cites the reference call site it replaces (see include/qcs.h).
#include "qcs_common.cuh"
#include "qcs_internal.h"

namespace qcs {

constexpr int EW_THREADS = 256;

struct BitList {
    int32_t m;
    int32_t pos[64];     // as given (MSB-first for outcomes / operator qubits)
    int32_t sorted[64];  // ascending
};

struct BinMap {
    int32_t nbits;           // index bits of the source
    uint64_t contrib[40];    // outcome bit(s) contributed by index bit b (0 if b is not measured)
};
template <typename T> struct VecOf;
template <> struct VecOf<float2> { static constexpr int VE = 2, LV = 1; };
template <> struct VecOf<double2> { static constexpr int VE = 1, LV = 0; };
template <> struct VecOf<double> { static constexpr int VE = 2, LV = 1; };
constexpr int MS_VECS = 8;                          // vectors per thread and segment
constexpr int MS_SEG_VECS = 256 * MS_VECS;          // 2048
template <typename T> __device__ __forceinline__ void vec_probs(const uint4 v, double (&p)[2]);
template <> __device__ __forceinline__ void vec_probs<float2>(const uint4 v, double (&p)[2]) {
    const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z), d = __uint_as_float(v.w);
    p[0] = (double)fmaf(a, a, b * b); p[1] = (double)fmaf(c, c, d * d);
}
template <> __device__ __forceinline__ void vec_probs<double2>(const uint4 v, double (&p)[2]) {
    const double a = __hiloint2double(v.y, v.x), b = __hiloint2double(v.w, v.z);
    p[0] = fma(a, a, b * b); p[1] = 0.0;
}
template <> __device__ __forceinline__ void vec_probs<double>(const uint4 v, double (&p)[2]) {
    p[0] = __hiloint2double(v.y, v.x); p[1] = __hiloint2double(v.w, v.z);
}

static unsigned ew_grid(uint64_t work_items, int per_sm = 8) {
    const DeviceInfo &dev = device_info();
    uint64_t g = (work_items + EW_THREADS - 1) / EW_THREADS;
    const uint64_t cap = (uint64_t)dev.sms * per_sm;
    if (g > cap) g = cap;
    if (g > MAX_GRID) g = MAX_GRID;
    if (g < 1) g = 1;
    return (unsigned)g;
}

template <typename C> __device__ __forceinline__ C cscale(const C a, const double s) {
    C r; r.x = (decltype(r.x))(a.x * s); r.y = (decltype(r.y))(a.y * s); return r;
}

// ---- set_basis (state.py:482-484, 515-517, 548-550) ---------------------------------------------
template <typename C> __global__ void set_one_kernel(C *buf, uint64_t index) {
    C one; one.x = 1; one.y = 0;
    buf[index] = one;
}

// ---- abs2 (state.py:200) --------------------------------------------------------------------------
template <typename C>
__global__ void __launch_bounds__(EW_THREADS) abs2_kernel(double *__restrict__ probs, const C *__restrict__ src,
                                                          uint64_t N, const double *norm_in, double *sum_out,
                                                          void *ws) {
    const double inv = norm_in ? 1.0 / *norm_in : 1.0;
    double acc[1] = {0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const double p = cabs2(src[i]) * inv;
        if (probs) probs[i] = p;
        acc[0] += p;
    }
    if (sum_out) grid_sum_publish<1>(acc, ws, sum_out);
}
// same, one 16-byte vector (VE amplitudes) per thread and step, four steps in flight; N >= 2 amplitudes
template <typename C>
__global__ void __launch_bounds__(EW_THREADS) abs2_vec_kernel(double *__restrict__ probs, const C *__restrict__ src,
                                                              uint64_t NV, const double *norm_in, double *sum_out,
                                                              void *ws) {
    constexpr int VE = VecOf<C>::VE;
    const double inv = norm_in ? 1.0 / *norm_in : 1.0;
    double acc[1] = {0.0};
    const uint4 *vs = reinterpret_cast<const uint4 *>(src);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
#pragma unroll 4
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < NV; v += stride) {
        double p[2];
        vec_probs<C>(__ldcs(vs + v), p);
        p[0] *= inv; p[1] *= inv;
        if (probs) {
            if (VE == 2) reinterpret_cast<double2 *>(probs)[v] = make_double2(p[0], p[1]);
            else probs[v] = p[0];
        }
        acc[0] += p[0] + p[1];
    }
    if (sum_out) grid_sum_publish<1>(acc, ws, sum_out);
}

// ---- dot (state.py:89) ----------------------------------------------------------------------------
template <typename C>
__global__ void __launch_bounds__(EW_THREADS) dot_kernel(const C *__restrict__ a, const C *__restrict__ b, uint64_t N,
                                                         double *out2, void *ws) {
    double acc[2] = {0.0, 0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const C x = a[i], y = b[i];
        acc[0] += (double)x.x * (double)y.x + (double)x.y * (double)y.y;   // conj(x) * y
        acc[1] += (double)x.x * (double)y.y - (double)x.y * (double)y.x;
    }
    grid_sum_publish<2>(acc, ws, out2);
}

// ---- scale / lincomb (state.py:97, 165-185) -------------------------------------------------------
template <typename C>
__global__ void __launch_bounds__(EW_THREADS) lincomb_kernel(C *__restrict__ dst, const C *a, const C *b, uint64_t N,
                                                             double are, double aim, double bre, double bim,
                                                             const double *norm_a, const double *norm_b,
                                                             double *norm_out, void *ws) {
    const double sa = load_inv_norm(norm_a), sb = b ? load_inv_norm(norm_b) : 0.0;
    are *= sa; aim *= sa; bre *= sb; bim *= sb;
    double acc[1] = {0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const C x = a[i];
        double re = are * (double)x.x - aim * (double)x.y;
        double im = are * (double)x.y + aim * (double)x.x;
        if (b) {
            const C y = b[i];
            re += bre * (double)y.x - bim * (double)y.y;
            im += bre * (double)y.y + bim * (double)y.x;
        }
        C r; r.x = (decltype(r.x))re; r.y = (decltype(r.y))im;
        dst[i] = r;
        acc[0] += cabs2(r);
    }
    if (norm_out) grid_sum_publish<1>(acc, ws, norm_out);
}

// ---- kron (tensors.py:76) --------------------------------------------------------------------------
template <typename C>
__global__ void __launch_bounds__(EW_THREADS) kron_kernel(C *__restrict__ dst, const C *__restrict__ a, const C *__restrict__ b,
                                                          int na, int nb, const double *norm_a, const double *norm_b,
                                                          double *norm_out, void *ws) {
    const double s = load_inv_norm(norm_a) * load_inv_norm(norm_b);
    const uint64_t N = 1ull << (na + nb), maskb = (1ull << nb) - 1ull;
    double acc[1] = {0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const C x = a[i >> nb], y = b[i & maskb];
        const double re = ((double)x.x * (double)y.x - (double)x.y * (double)y.y) * s;
        const double im = ((double)x.x * (double)y.y + (double)x.y * (double)y.x) * s;
        C r; r.x = (decltype(r.x))re; r.y = (decltype(r.y))im;
        dst[i] = r;
        acc[0] += cabs2(r);
    }
    if (norm_out) grid_sum_publish<1>(acc, ws, norm_out);
}

// kron with nb >= 1 (complex64: the two amplitudes of a 16-byte vector share the `a` factor) -- one vector
// store per thread and step, arithmetic in the amplitude type
template <typename R>
__global__ void __launch_bounds__(EW_THREADS) kron_vec_kernel(typename CxT<R>::T *__restrict__ dst,
                                                              const typename CxT<R>::T *__restrict__ a,
                                                              const typename CxT<R>::T *__restrict__ b, int na, int nb,
                                                              const double *norm_a, const double *norm_b,
                                                              double *norm_out, void *ws) {
    using C = typename CxT<R>::T;
    constexpr int VE = VecOf<C>::VE, LV = VecOf<C>::LV;
    const R s = (R)(load_inv_norm(norm_a) * load_inv_norm(norm_b));
    const uint64_t NV = (1ull << (na + nb)) >> LV, maskb = (1ull << nb) - 1ull;
    double acc[1] = {0.0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
#pragma unroll 4
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < NV; v += stride) {
        const uint64_t i = v << LV;
        C x = a[i >> nb];
        x.x *= s; x.y *= s;
        C r[VE], yv[VE];
        if (VE == 2) {      // the pair of `b` factors is one aligned 16-byte load
            const float4 yy = *reinterpret_cast<const float4 *>(b + (i & maskb));
            yv[0].x = yy.x; yv[0].y = yy.y; yv[VE - 1].x = yy.z; yv[VE - 1].y = yy.w;
        } else {
            yv[0] = b[i & maskb];
        }
        R part = R(0);
#pragma unroll
        for (int u = 0; u < VE; ++u) {
            const C y = yv[u];
            r[u] = cmul(x, y);
            part += r[u].x * r[u].x + r[u].y * r[u].y;
        }
        if (VE == 2) {
            float4 o;
            o.x = (float)r[0].x; o.y = (float)r[0].y; o.z = (float)r[VE - 1].x; o.w = (float)r[VE - 1].y;
            __stcs(reinterpret_cast<float4 *>(dst) + v, o);
        } else {
            dst[v] = r[0];
        }
        acc[0] += (double)part;
    }
    if (norm_out) grid_sum_publish<1>(acc, ws, norm_out);
}

// kron with a small left factor (na <= 6): a thread loads one 16-byte vector of `b` once and writes it times
// every amplitude of `a`, so `b` is read once instead of 2^na times (it does not fit L2 for large states)
template <typename R>
__global__ void __launch_bounds__(EW_THREADS) kron_small_a_kernel(typename CxT<R>::T *__restrict__ dst,
                                                                  const typename CxT<R>::T *__restrict__ a,
                                                                  const typename CxT<R>::T *__restrict__ b, int na, int nb,
                                                                  const double *norm_a, const double *norm_b,
                                                                  double *norm_out, void *ws) {
    using C = typename CxT<R>::T;
    constexpr int VE = VecOf<C>::VE, LV = VecOf<C>::LV;
    __shared__ C as[64];
    const R s = (R)(load_inv_norm(norm_a) * load_inv_norm(norm_b));
    const uint32_t NA = 1u << na;
    if (threadIdx.x < NA) { C x = a[threadIdx.x]; x.x *= s; x.y *= s; as[threadIdx.x] = x; }
    __syncthreads();
    const uint64_t NVB = (1ull << nb) >> LV;
    double acc[1] = {0.0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < NVB; v += stride) {
        C yv[VE];
        if (VE == 2) {
            const float4 yy = __ldcs(reinterpret_cast<const float4 *>(b) + v);
            yv[0].x = yy.x; yv[0].y = yy.y; yv[VE - 1].x = yy.z; yv[VE - 1].y = yy.w;
        } else {
            yv[0] = b[v];
        }
        R part = R(0);
#pragma unroll 4
        for (uint32_t ai = 0; ai < NA; ++ai) {
            const C x = as[ai];
            C r[VE];
#pragma unroll
            for (int u = 0; u < VE; ++u) {
                r[u] = cmul(x, yv[u]);
                part += r[u].x * r[u].x + r[u].y * r[u].y;
            }
            const uint64_t o = ((uint64_t)ai << (nb - LV)) + v;
            if (VE == 2) {
                float4 w;
                w.x = (float)r[0].x; w.y = (float)r[0].y; w.z = (float)r[VE - 1].x; w.w = (float)r[VE - 1].y;
                __stcs(reinterpret_cast<float4 *>(dst) + o, w);
            } else {
                dst[o] = r[0];
            }
        }
        acc[0] += (double)part;
    }
    if (norm_out) grid_sum_publish<1>(acc, ws, norm_out);
}

// ---- permute_bits (state.py:117,162) --------------------------------------------------------------
// conjugation for the adjoint / conjugate forms of the bit permutation (operators.py:95-114): a sign flip of the
// imaginary parts, on typed amplitudes and on raw 16-byte vectors (complex64: two amplitudes, complex128: one)
template <bool CONJ, typename C> __device__ __forceinline__ C maybe_conj(C v) { if (CONJ) v.y = -v.y; return v; }
template <bool CONJ, typename C> __device__ __forceinline__ uint4 maybe_conj_vec(uint4 v) {
    if (CONJ) {
        if (sizeof(C) == 8) { v.y ^= 0x80000000u; v.w ^= 0x80000000u; }
        else v.w ^= 0x80000000u;
    }
    return v;
}
template <typename C, bool CONJ>
__global__ void __launch_bounds__(EW_THREADS) permute_kernel(C *__restrict__ dst, const C *__restrict__ src, int n,
                                                             const __grid_constant__ BitList bl) {
    const uint64_t N = 1ull << n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t j = 0;
        for (int b = 0; b < n; ++b) j |= ((i >> b) & 1ull) << bl.pos[b];
        dst[i] = maybe_conj<CONJ>(src[j]);
    }
}
// Large states: the source index of destination element e = ((seg * 8 + j) * 256 + tid) * VE + u is an OR of
// per-bit contributions, P(e) = Ps(seg) | Pj(j) | Pt(tid, u), so the n-step bit loop runs once per segment
// instead of once per amplitude; the destination is written with 16-byte vectors, the source is read with
// 16-byte vectors when flat bit 0 stays in place (complex64) and with 8-byte gathers otherwise.
template <typename C, bool CONJ>
__global__ void __launch_bounds__(EW_THREADS) permute_stream_kernel(C *__restrict__ dst, const C *__restrict__ src,
                                                                   const __grid_constant__ BinMap pm) {
    constexpr int VE = VecOf<C>::VE, LV = VecOf<C>::LV;
    const int tid = threadIdx.x;
    uint64_t Pt = 0;
    for (int k = 0; k < 8; ++k)
        if ((tid >> k) & 1) Pt |= pm.contrib[k + LV];
    const uint64_t p_u = LV ? pm.contrib[0] : 0ull;             // source offset of the second amplitude of a vector
    const uint64_t pj0 = pm.contrib[8 + LV], pj1 = pm.contrib[9 + LV], pj2 = pm.contrib[10 + LV];
    const int seg_shift = 11 + LV;
    const uint64_t nseg = (1ull << pm.nbits) >> seg_shift;
    const bool pair_contiguous = VE == 2 && p_u == 1ull;
    for (uint64_t sgm = blockIdx.x; sgm < nseg; sgm += gridDim.x) {
        uint64_t ps = 0;
        for (int k = 0; k + seg_shift < pm.nbits; ++k)
            if ((sgm >> k) & 1ull) ps |= pm.contrib[k + seg_shift];
        const uint64_t base = ps | Pt;
        uint4 *vd = reinterpret_cast<uint4 *>(dst) + sgm * (uint64_t)MS_SEG_VECS + tid;
        uint4 val[MS_VECS];
#pragma unroll
        for (int j = 0; j < MS_VECS; ++j) {
            const uint64_t e = base | ((j & 1) ? pj0 : 0ull) | ((j & 2) ? pj1 : 0ull) | ((j & 4) ? pj2 : 0ull);
            if (VE == 1 || pair_contiguous) {
                val[j] = *reinterpret_cast<const uint4 *>(src + e);
            } else {
                const uint2 lo = *reinterpret_cast<const uint2 *>(src + e);
                const uint2 hi = *reinterpret_cast<const uint2 *>(src + (e | p_u));
                val[j] = make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
        }
#pragma unroll
        for (int j = 0; j < MS_VECS; ++j) __stcs(vd + j * EW_THREADS, maybe_conj_vec<CONJ, C>(val[j]));
    }
}

// Flat bit 0 of the destination fed by another source bit (pm.contrib[0] = 2^j, j > 0), e.g. swapping the last qubit
// with any other: a destination pair (i, i + 1) gathers from two source locations whose neighbours belong to a
// destination pair far away -- every 32-byte source sector is fetched twice (0.63-0.66 of the copy peak).  Here a thread
// takes the 2 x 2 block instead: destination pairs at i and i | 2^d0 (d0 = where source bit 0 goes) <- source pairs at
// e and e | 2^j, transposed in registers; every access is a full 16-byte (complex64) or 32-byte (complex128) unit.
// q enumerates the blocks: bit k of q is destination bit k + 1 (k + 1 < d0) or k + 2.
template <typename C, bool CONJ>
__global__ void __launch_bounds__(EW_THREADS) permute_quad_kernel(C *__restrict__ dst, const C *__restrict__ src,
                                                                  const __grid_constant__ BinMap pm, int d0) {
    constexpr int NVQ = sizeof(C) == 8 ? MS_VECS : MS_VECS / 2;     // blocks per thread and segment (same bytes in flight)
    constexpr int SEGB = sizeof(C) == 8 ? 11 : 10;
    const int tid = threadIdx.x;
    auto contrib_q = [&](int k) -> uint64_t { return pm.contrib[k + 1 < d0 ? k + 1 : k + 2]; };
    uint64_t Pt = 0;
    for (int k = 0; k < 8; ++k)
        if ((tid >> k) & 1) Pt |= contrib_q(k);
    const uint64_t p_u = pm.contrib[0];                 // source offset of destination bit 0
    const uint64_t pj0 = contrib_q(8), pj1 = contrib_q(9), pj2 = contrib_q(10);
    const int nq = pm.nbits - 2;                        // bits of q
    const uint64_t nseg = (1ull << nq) >> SEGB;
    const uint64_t lowmask = (1ull << (d0 - 1)) - 1ull;
    for (uint64_t sgm = blockIdx.x; sgm < nseg; sgm += gridDim.x) {
        uint64_t ps = 0;
        for (int k = 0; k + SEGB < nq; ++k)
            if ((sgm >> k) & 1ull) ps |= contrib_q(k + SEGB);
        const uint64_t base = ps | Pt;
        C a0[NVQ], a1[NVQ], b0[NVQ], b1[NVQ];
#pragma unroll
        for (int j = 0; j < NVQ; ++j) {
            const uint64_t e = base | ((j & 1) ? pj0 : 0ull) | ((j & 2) ? pj1 : 0ull) | ((j & 4) ? pj2 : 0ull);
            if (sizeof(C) == 8) {
                const uint4 va = *reinterpret_cast<const uint4 *>(src + e), vb = *reinterpret_cast<const uint4 *>(src + (e | p_u));
                a0[j] = *reinterpret_cast<const C *>(&va.x); a1[j] = *reinterpret_cast<const C *>(&va.z);
                b0[j] = *reinterpret_cast<const C *>(&vb.x); b1[j] = *reinterpret_cast<const C *>(&vb.z);
            } else {
                a0[j] = src[e]; a1[j] = src[e + 1];               // one 32-byte sector
                b0[j] = src[e | p_u]; b1[j] = src[(e | p_u) + 1];
            }
        }
#pragma unroll
        for (int j = 0; j < NVQ; ++j) {
            if (CONJ) { a0[j].y = -a0[j].y; a1[j].y = -a1[j].y; b0[j].y = -b0[j].y; b1[j].y = -b1[j].y; }
            const uint64_t q = (sgm << SEGB) | ((uint64_t)j << 8) | (uint64_t)tid;
            const uint64_t i = (((q >> (d0 - 1)) << d0) | (q & lowmask)) << 1;    // destination index, bits 0 and d0 clear
            if (sizeof(C) == 8) {
                uint4 lo, hi;
                *reinterpret_cast<C *>(&lo.x) = a0[j]; *reinterpret_cast<C *>(&lo.z) = b0[j];
                *reinterpret_cast<C *>(&hi.x) = a1[j]; *reinterpret_cast<C *>(&hi.z) = b1[j];
                __stcs(reinterpret_cast<uint4 *>(dst + i), lo);
                __stcs(reinterpret_cast<uint4 *>(dst + (i | (1ull << d0))), hi);
            } else {
                dst[i] = a0[j]; dst[i + 1] = b0[j];
                dst[i | (1ull << d0)] = a1[j]; dst[(i | (1ull << d0)) + 1] = b1[j];
            }
        }
    }
}

// ---- marginal probabilities (state.py:262-269; density_operator.py:140-147) ----------------------
template <typename T> __device__ __forceinline__ double prob_of(const T v) { return cabs2(v); }
template <> __device__ __forceinline__ double prob_of<double>(const double v) { return v; }

__device__ __forceinline__ uint32_t extract_bits(uint64_t i, const BitList &bl) {
    uint32_t x = 0;
    for (int j = 0; j < bl.m; ++j) x = (x << 1) | (uint32_t)((i >> bl.pos[j]) & 1ull);
    return x;
}

constexpr int MARG_SEG = 16;          // elements per thread per segment
constexpr int MARG_SMALL_BITS = 10;   // shared-memory histogram up to 2^10 outcomes

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) marginal_small_kernel(double *out, const T *__restrict__ src, uint64_t N,
                                                                   const __grid_constant__ BitList bl,
                                                                   const double *norm_in, double *hists, void *ws) {
    extern __shared__ double hist[];
    __shared__ int last_flag;
    const uint32_t nbins = 1u << bl.m;
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) hist[b] = 0.0;
    __syncthreads();
    const uint64_t seg = (uint64_t)blockDim.x * MARG_SEG;
    for (uint64_t s0 = (uint64_t)blockIdx.x * seg; s0 < N; s0 += (uint64_t)gridDim.x * seg) {
        uint32_t cur = 0xffffffffu;
        double acc = 0.0;
        for (int j = 0; j < MARG_SEG; ++j) {
            const uint64_t i = s0 + (uint64_t)j * blockDim.x + threadIdx.x;
            if (i >= N) break;
            const uint32_t bin = extract_bits(i, bl);
            const double p = prob_of(src[i]);
            if (bin != cur) {
                if (cur != 0xffffffffu) atomicAdd(&hist[cur], acc);
                cur = bin; acc = 0.0;
            }
            acc += p;
        }
        if (cur != 0xffffffffu) atomicAdd(&hist[cur], acc);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) hists[(size_t)blockIdx.x * nbins + b] = hist[b];
    __threadfence();
    __syncthreads();
    unsigned *ticket = reinterpret_cast<unsigned *>(ws);
    if (threadIdx.x == 0) last_flag = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (last_flag) {
        __threadfence();
        const double inv = norm_in ? 1.0 / *norm_in : 1.0;
        for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) {
            double a = 0.0;
            for (unsigned g = 0; g < gridDim.x; ++g) a += __ldcg(&hists[(size_t)g * nbins + b]);
            out[b] = a * inv;
        }
        if (threadIdx.x == 0) { *ticket = 0u; __threadfence(); }
    }
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS) marginal_large_kernel(double *out, const T *__restrict__ src, uint64_t N,
                                                                   const __grid_constant__ BitList bl, int exact,
                                                                   const double *norm_in) {
    const double inv = norm_in ? 1.0 / *norm_in : 1.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = 0;
        for (int j = 0; j < bl.m; ++j) x = (x << 1) | ((i >> bl.pos[j]) & 1ull);
        const double p = prob_of(src[i]) * inv;
        if (exact) out[x] = p;          // every qubit measured: one contribution per outcome
        else atomicAdd(&out[x], p);
    }
}

// Streaming marginal kernel for states of at least one segment (2048 16-byte vectors).  Element index
// e = ((seg * 8 + j) * 256 + tid) * VE + u, and the outcome ("bin") of an index is an OR of per-bit
// contributions, so bin(e) = Bs(seg) | Bj(j) | Bt(tid, u): Bt is a per-thread constant, Bj a compile-time
// selection of three per-pass masks, Bs is recomputed once per segment.  A CTA owns a contiguous range of
// segments; a thread keeps 8 * VE float64 accumulators (one per (j, u)) for as long as Bs does not change
// and flushes them with atomics into the shared-memory histogram (<= 2^10 outcomes) or the global one.
// Per element: one 16-byte load shared by VE elements, 2 FMA, 1 add -- the kernel is HBM-bound.
template <typename T, bool SMALL>
__global__ void __launch_bounds__(EW_THREADS) marginal_stream_kernel(double *out, const T *__restrict__ src,
                                                                    const __grid_constant__ BinMap bm, int m,
                                                                    const double *norm_in, double *hists, void *ws) {
    constexpr int VE = VecOf<T>::VE, LV = VecOf<T>::LV;
    extern __shared__ double hist[];
    __shared__ int last_flag;
    const uint32_t nbins = SMALL ? (1u << m) : 0u;
    if (SMALL) {
        for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) hist[b] = 0.0;
        __syncthreads();
    }
    const int tid = threadIdx.x;
    const double inv = norm_in ? 1.0 / *norm_in : 1.0;
    uint64_t Bt[VE];
#pragma unroll
    for (int u = 0; u < VE; ++u) {
        const uint32_t e = ((uint32_t)tid << LV) | (uint32_t)u;
        uint64_t b = 0;
        for (int k = 0; k < 8 + LV; ++k)
            if ((e >> k) & 1u) b |= bm.contrib[k];
        Bt[u] = b;
    }
    const uint64_t bj0 = bm.contrib[8 + LV], bj1 = bm.contrib[9 + LV], bj2 = bm.contrib[10 + LV];
    const int seg_shift = 11 + LV;                              // element bits below the segment index
    const uint64_t nseg = (1ull << bm.nbits) >> seg_shift;
    const uint64_t s_begin = nseg * blockIdx.x / gridDim.x, s_end = nseg * (blockIdx.x + 1) / gridDim.x;
    const uint4 *vsrc = reinterpret_cast<const uint4 *>(src) + tid;
    double acc[MS_VECS][VE];
#pragma unroll
    for (int j = 0; j < MS_VECS; ++j)
#pragma unroll
        for (int u = 0; u < VE; ++u) acc[j][u] = 0.0;
    uint64_t cur = 0;
    auto flush = [&]() {
#pragma unroll
        for (int j = 0; j < MS_VECS; ++j) {
            const uint64_t bj = ((j & 1) ? bj0 : 0ull) | ((j & 2) ? bj1 : 0ull) | ((j & 4) ? bj2 : 0ull);
#pragma unroll
            for (int u = 0; u < VE; ++u) {
                if (acc[j][u] != 0.0) {
                    const uint64_t bin = cur | bj | Bt[u];
                    if (SMALL) atomicAdd(&hist[bin], acc[j][u]);
                    else atomicAdd(&out[bin], acc[j][u] * inv);
                    acc[j][u] = 0.0;
                }
            }
        }
    };
    for (uint64_t sgm = s_begin; sgm < s_end; ++sgm) {
        uint4 v[MS_VECS];
        const uint4 *p = vsrc + sgm * (uint64_t)MS_SEG_VECS;
#pragma unroll
        for (int j = 0; j < MS_VECS; ++j) v[j] = __ldcs(p + j * EW_THREADS);
        uint64_t bs = 0;
        for (int k = 0; k + seg_shift < bm.nbits; ++k)
            if ((sgm >> k) & 1ull) bs |= bm.contrib[k + seg_shift];
        if (bs != cur) { flush(); cur = bs; }
#pragma unroll
        for (int j = 0; j < MS_VECS; ++j) {
            double pr[2];
            vec_probs<T>(v[j], pr);
#pragma unroll
            for (int u = 0; u < VE; ++u) acc[j][u] += pr[u];
        }
    }
    flush();
    if (!SMALL) return;
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) hists[(size_t)blockIdx.x * nbins + b] = hist[b];
    __threadfence();
    __syncthreads();
    unsigned *ticket = reinterpret_cast<unsigned *>(ws);
    if (threadIdx.x == 0) last_flag = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (last_flag) {
        __threadfence();
        for (uint32_t b = threadIdx.x; b < nbins; b += blockDim.x) {
            double a = 0.0;
            for (unsigned g = 0; g < gridDim.x; ++g) a += __ldcg(&hists[(size_t)g * nbins + b]);
            out[b] = a * inv;
        }
        if (threadIdx.x == 0) { *ticket = 0u; __threadfence(); }
    }
}

// ---- collapse (state.py:368-378; density_operator.py:266-279) ------------------------------------
template <typename C>
__global__ void __launch_bounds__(EW_THREADS) collapse_kernel(C *__restrict__ dst, const C *src, int n,
                                                              const __grid_constant__ BitList bl, uint64_t outbits,
                                                              uint64_t mask, int remove, double scale, double *norm_out,
                                                              void *ws) {
    double acc[1] = {0.0};
    if (remove) {
        const uint64_t N = 1ull << (n - bl.m);
        for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (uint64_t)gridDim.x * blockDim.x) {
            uint64_t i = j;
            for (int b = 0; b < bl.m; ++b) i = insert_zero(i, bl.sorted[b]);
            const C r = cscale(src[i | outbits], scale);
            dst[j] = r;
            acc[0] += cabs2(r);
        }
    } else {
        const uint64_t N = 1ull << n;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
            C r; r.x = 0; r.y = 0;
            if ((i & mask) == outbits) r = cscale(src[i], scale);
            dst[i] = r;
            acc[0] += cabs2(r);
        }
    }
    if (norm_out) grid_sum_publish<1>(acc, ws, norm_out);
}

// ---- density helpers (density_operator.py:60, 106-125) -------------------------------------------
__device__ __forceinline__ uint64_t spread_bits(uint64_t x) {   // bit b -> bit 2b
    x &= 0xffffffffull;
    x = (x | (x << 16)) & 0x0000ffff0000ffffull;
    x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
    x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}
__device__ __forceinline__ uint64_t compact_bits(uint64_t x) {  // bit 2b -> bit b
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return x;
}

template <typename C>
__global__ void __launch_bounds__(EW_THREADS) diag_probs_kernel(double *__restrict__ probs, const C *__restrict__ rho, int n,
                                                                double *sum_out, void *ws) {
    const uint64_t N = 1ull << n;
    double acc[1] = {0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const double p = (double)rho[3ull * spread_bits(i)].x;
        probs[i] = p;
        acc[0] += p;
    }
    if (sum_out) grid_sum_publish<1>(acc, ws, sum_out);
}

template <typename C>
__global__ void __launch_bounds__(EW_THREADS) outer_kernel(C *__restrict__ rho, const C *__restrict__ psi, int n, double p,
                                                           const double *norm_psi, int accumulate) {
    // one 16-byte vector of rho per thread and step: for complex64 its two entries share the row amplitude
    // and take the column amplitudes c, c | 1 (one 16-byte load of psi)
    constexpr int VE = VecOf<C>::VE, LV = VecOf<C>::LV;
    using R = decltype(C().x);
    const R w = (R)(p * (norm_psi ? 1.0 / *norm_psi : 1.0));
    const uint64_t NV = (1ull << (2 * n)) >> LV;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
#pragma unroll 2
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < NV; v += stride) {
        const uint64_t i = v << LV;
        C a = psi[compact_bits(i >> 1)];   // row index: odd bits
        a.x *= w; a.y *= w;
        const uint64_t c = compact_bits(i);   // column index: even bits
        C b[VE], r[VE];
        if (VE == 2) {
            const float4 bb = *reinterpret_cast<const float4 *>(psi + c);
            b[0].x = bb.x; b[0].y = bb.y; b[VE - 1].x = bb.z; b[VE - 1].y = bb.w;
        } else {
            b[0] = psi[c];
        }
#pragma unroll
        for (int u = 0; u < VE; ++u) {      // a * conj(b)
            r[u].x = a.x * b[u].x + a.y * b[u].y;
            r[u].y = a.y * b[u].x - a.x * b[u].y;
        }
        if (VE == 2) {
            float4 *pr = reinterpret_cast<float4 *>(rho) + v;
            float4 o;
            o.x = (float)r[0].x; o.y = (float)r[0].y; o.z = (float)r[VE - 1].x; o.w = (float)r[VE - 1].y;
            if (accumulate) { const float4 old = *pr; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            *pr = o;
        } else {
            if (accumulate) { const C old = rho[v]; r[0].x += old.x; r[0].y += old.y; }
            rho[v] = r[0];
        }
    }
}

// Large rho (>= 2^14 vectors): the vector index is ((seg * 8 + j) * 256 + tid); de-interleaving the index into (row,
// column) is an OR of per-field contributions, so the two 6-step bit compactions run once per thread and once per
// segment instead of once per vector, every thread keeps 8 independent 16-byte stores in flight, and the stores bypass
// L1 (__stcs): the kernel is a pure write stream (0.66-0.71 of the copy peak before, with the compactions per vector).
template <typename C>
__global__ void __launch_bounds__(EW_THREADS) outer_stream_kernel(C *__restrict__ rho, const C *__restrict__ psi, int n, double p,
                                                                  const double *norm_psi, int accumulate) {
    constexpr int VE = VecOf<C>::VE, LV = VecOf<C>::LV;
    using R = decltype(C().x);
    const R w = (R)(p * (norm_psi ? 1.0 / *norm_psi : 1.0));
    const int tid = threadIdx.x;
    const uint64_t nseg = ((1ull << (2 * n)) >> LV) >> 11;
    const uint64_t it = (uint64_t)tid << LV;                       // amplitude-index bits owned by the thread
    const uint64_t row_t = compact_bits(it >> 1), col_t = compact_bits(it);
    for (uint64_t sgm = blockIdx.x; sgm < nseg; sgm += gridDim.x) {
        const uint64_t is = sgm << (11 + LV);
        const uint64_t row_s = compact_bits(is >> 1) | row_t, col_s = compact_bits(is) | col_t;
        uint4 *vd = reinterpret_cast<uint4 *>(rho) + sgm * (uint64_t)MS_SEG_VECS + tid;
        uint4 out[MS_VECS];
#pragma unroll
        for (int j = 0; j < MS_VECS; ++j) {
            const uint64_t ij = (uint64_t)j << (8 + LV);           // compile-time per j
            const uint64_t row = row_s | compact_bits(ij >> 1), col = col_s | compact_bits(ij);
            C a = __ldg(psi + row);
            a.x *= w; a.y *= w;
            if (VE == 2) {
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(psi + col));
                float4 o;
                o.x = (float)(a.x * bb.x + a.y * bb.y); o.y = (float)(a.y * bb.x - a.x * bb.y);
                o.z = (float)(a.x * bb.z + a.y * bb.w); o.w = (float)(a.y * bb.z - a.x * bb.w);
                if (accumulate) { const float4 old = *reinterpret_cast<const float4 *>(vd + j * EW_THREADS); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                out[j] = make_uint4(__float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w));
            } else {
                const C bb = __ldg(psi + col);
                double2 o;
                o.x = (double)(a.x * bb.x + a.y * bb.y); o.y = (double)(a.y * bb.x - a.x * bb.y);
                if (accumulate) { const double2 old = *reinterpret_cast<const double2 *>(vd + j * EW_THREADS); o.x += old.x; o.y += old.y; }
                out[j] = make_uint4(__double2loint(o.x), __double2hiint(o.x), __double2loint(o.y), __double2hiint(o.y));
            }
        }
#pragma unroll
        for (int j = 0; j < MS_VECS; ++j) __stcs(vd + j * EW_THREADS, out[j]);
    }
}

// ---- one-qubit gate as a pure streaming pass (operators.py:261 for a 2x2 operator) --------------------
// North star (1): "high target qubits handled by strided 128-bit loads".  A pass that holds a single one-qubit
// gate needs no shared-memory tile: a thread loads the two 16-byte vectors that differ in the target bit,
// applies the 2x2 matrix (identity, still scaled by the lazy 1/sqrt(norm), where a control bit is 0) and stores
// both -- four pairs in flight per thread.  For complex64 and target bit 0 the pair sits inside one vector.
template <typename R> struct Gate1Args {
    R m[8];                 // m00 m01 m10 m11 as (re, im)
    uint64_t ctrl;          // flat-index control mask
    int32_t n, tbit;
};
template <typename C> __device__ __forceinline__ void unpack_vec(const uint4 v, C (&a)[2]);
template <> __device__ __forceinline__ void unpack_vec<float2>(const uint4 v, float2 (&a)[2]) {
    a[0] = make_float2(__uint_as_float(v.x), __uint_as_float(v.y));
    a[1] = make_float2(__uint_as_float(v.z), __uint_as_float(v.w));
}
template <> __device__ __forceinline__ void unpack_vec<double2>(const uint4 v, double2 (&a)[2]) {
    a[0] = make_double2(__hiloint2double(v.y, v.x), __hiloint2double(v.w, v.z));
    a[1] = a[0];
}
__device__ __forceinline__ uint4 pack_vec(const float2 (&a)[2]) {
    return make_uint4(__float_as_uint(a[0].x), __float_as_uint(a[0].y), __float_as_uint(a[1].x), __float_as_uint(a[1].y));
}
__device__ __forceinline__ uint4 pack_vec(const double2 (&a)[2]) {
    return make_uint4(__double2loint(a[0].x), __double2hiint(a[0].x), __double2loint(a[0].y), __double2hiint(a[0].y));
}
template <typename R, bool INVEC>
__global__ void __launch_bounds__(EW_THREADS) gate1_stream_kernel(typename CxT<R>::T *dst,
                                                                 const typename CxT<R>::T *src,
                                                                 const __grid_constant__ Gate1Args<R> g,
                                                                 const double *norm_in, double *norm_out, void *ws) {
    using C = typename CxT<R>::T;
    constexpr int VE = VecOf<C>::VE, LV = VecOf<C>::LV;
    const R s = (R)load_inv_norm(norm_in);
    const C m00 = cx<R>(g.m[0] * s, g.m[1] * s), m01 = cx<R>(g.m[2] * s, g.m[3] * s);
    const C m10 = cx<R>(g.m[4] * s, g.m[5] * s), m11 = cx<R>(g.m[6] * s, g.m[7] * s);
    const uint4 *vs = reinterpret_cast<const uint4 *>(src);
    uint4 *vd = reinterpret_cast<uint4 *>(dst);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    R part = R(0);
    double acc[1] = {0.0};
    auto apply = [&](C &a, C &b, uint64_t index_of_a) {          // (a, b) = the amplitudes with target bit 0 / 1
        C r0, r1;
        if ((index_of_a & g.ctrl) == g.ctrl) {
            r0 = cmul(m00, a); cfma(r0, m01, b);
            r1 = cmul(m10, a); cfma(r1, m11, b);
        } else {
            r0 = cx<R>(a.x * s, a.y * s); r1 = cx<R>(b.x * s, b.y * s);
        }
        part += r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y;
        a = r0; b = r1;
    };
    if (INVEC) {            // complex64, target bit 0: the pair is one vector
        const uint64_t NV = (1ull << g.n) >> LV;
#pragma unroll 4
        for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < NV; v += stride) {
            C a[2];
            unpack_vec<C>(__ldcs(vs + v), a);
            apply(a[0], a[1], v << LV);
            __stcs(vd + v, pack_vec(a));
            acc[0] += (double)part; part = R(0);
        }
    } else {
        const int tv = g.tbit - LV;                       // target bit in vector units
        const uint64_t NP = ((1ull << g.n) >> LV) >> 1;   // pairs of vectors
        const uint64_t himask = ~((1ull << tv) - 1ull);
#pragma unroll 4
        for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < NP; p += stride) {
            const uint64_t v0 = p + (p & himask), v1 = v0 | (1ull << tv);
            C a[2], b[2];
            unpack_vec<C>(__ldcs(vs + v0), a);
            unpack_vec<C>(__ldcs(vs + v1), b);
#pragma unroll
            for (int u = 0; u < VE; ++u) apply(a[u], b[u], (v0 << LV) | (uint64_t)u);
            __stcs(vd + v0, pack_vec(a));
            __stcs(vd + v1, pack_vec(b));
            acc[0] += (double)part; part = R(0);
        }
    }
    if (norm_out) grid_sum_publish<1>(acc, ws, norm_out);
}

// ---- partial trace (density_operator.py:127-138, np.einsum('ii...->...')) ------------------------
// out[o] = sum_t rho[base(o) | diag(t)]: base(o) scatters the 2k bits of the output index (row / column bits of
// the kept qubits, in the caller's order) to their flat bits in rho, diag(t) sets the row AND the column bit
// of traced qubit j to bit j of t.  Only 4^k 2^(d-k) of the 4^d entries are read.  One CTA per output entry
// (strided over the grid): its threads split the 2^(d-k) terms, float64 accumulation, block reduction.
struct TraceMap {
    int32_t nout_bits, ntraced;
    uint64_t out_contrib[64];     // flat bit of rho set by output-index bit b
    uint64_t tr_contrib[32];      // the two flat bits of rho set by bit j of the traced configuration
};
template <typename C>
__global__ void __launch_bounds__(EW_THREADS) partial_trace_kernel(C *__restrict__ out, const C *__restrict__ rho,
                                                                  const __grid_constant__ TraceMap tm) {
    __shared__ double red[2][EW_THREADS / 32];
    const uint64_t nout = 1ull << tm.nout_bits, nterms = 1ull << tm.ntraced;
    for (uint64_t o = blockIdx.x; o < nout; o += gridDim.x) {
        uint64_t base = 0;
        for (int b = 0; b < tm.nout_bits; ++b)
            if ((o >> b) & 1ull) base |= tm.out_contrib[b];
        double re = 0.0, im = 0.0;
        for (uint64_t t = threadIdx.x; t < nterms; t += blockDim.x) {
            uint64_t idx = base;
            for (int j = 0; j < tm.ntraced; ++j)
                if ((t >> j) & 1ull) idx |= tm.tr_contrib[j];
            const C v = rho[idx];
            re += (double)v.x; im += (double)v.y;
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) {
            re += __shfl_down_sync(0xffffffffu, re, sh);
            im += __shfl_down_sync(0xffffffffu, im, sh);
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { red[0][warp] = re; red[1][warp] = im; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double sr = 0.0, si = 0.0;
            for (int w = 0; w < EW_THREADS / 32; ++w) { sr += red[0][w]; si += red[1][w]; }
            C r; r.x = (decltype(r.x))sr; r.y = (decltype(r.y))si;
            out[o] = r;
        }
        __syncthreads();
    }
}

// ---- dense fallback: operator on k > 4 qubits (operators.py:261 with a wide operator) ------------
constexpr int DT = 16;
template <typename R>
__global__ void __launch_bounds__(DT * DT) dense_kernel(typename CxT<R>::T *__restrict__ dst,
                                                        const typename CxT<R>::T *__restrict__ src,
                                                        const typename CxT<R>::T *__restrict__ M, int n, int k,
                                                        const __grid_constant__ BitList bl, const double *norm_in) {
    using C = typename CxT<R>::T;
    __shared__ C Ms[DT][DT + 1];
    __shared__ C Xs[DT][DT + 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const uint64_t D = 1ull << k, NR = 1ull << (n - k);
    const uint64_t o = (uint64_t)blockIdx.y * DT + ty;
    const uint64_t r = (uint64_t)blockIdx.x * DT + tx;
    // flat index of "rest" configuration r: zeros inserted at the (ascending) target bits
    uint64_t rbase = r;
    for (int b = 0; b < k; ++b) rbase = insert_zero(rbase, bl.sorted[b]);
    auto deposit = [&](uint64_t x) {
        uint64_t v = 0;
        for (int j = 0; j < k; ++j)
            if ((x >> (k - 1 - j)) & 1ull) v |= 1ull << bl.pos[j];
        return v;
    };
    const R s = (R)load_inv_norm(norm_in);
    C acc = cx<R>(R(0), R(0));
    for (uint64_t j0 = 0; j0 < D; j0 += DT) {
        Ms[ty][tx] = M[o * D + j0 + tx];
        Xs[ty][tx] = r < NR ? src[rbase | deposit(j0 + ty)] : cx<R>(R(0), R(0));
        __syncthreads();
#pragma unroll
        for (int j = 0; j < DT; ++j) cfma(acc, Ms[ty][j], Xs[j][tx]);
        __syncthreads();
    }
    if (r < NR) {
        acc.x *= s; acc.y *= s;
        dst[rbase | deposit(o)] = acc;
    }
}

// ================================================================================================
// host wrappers
// ================================================================================================
static int check_launch() { return cudaGetLastError() == cudaSuccess ? QCS_OK : set_error(QCS_ERR_CUDA, "kernel launch failed"); }

static int make_bitlist(BitList &bl, const int32_t *pos, int m, int n, bool need_sorted) {
    if (m < 0 || m > 64 || m > n) return set_error(QCS_ERR_ARG, "bad bit count");
    bl.m = m;
    uint64_t seen = 0;
    for (int j = 0; j < m; ++j) {
        if (pos[j] < 0 || pos[j] >= n) return set_error(QCS_ERR_ARG, "bit position out of range");
        if (seen & (1ull << pos[j])) return set_error(QCS_ERR_ARG, "repeated bit position");
        seen |= 1ull << pos[j];
        bl.pos[j] = pos[j];
    }
    if (need_sorted) {
        int c = 0;
        for (int b = 0; b < n; ++b) if ((seen >> b) & 1ull) bl.sorted[c++] = b;
    }
    return QCS_OK;
}

#define DISPATCH_C(dtype, CALL)                         \
    do {                                                \
        if ((dtype) == QCS_C64) { using C = float2; CALL; }        \
        else if ((dtype) == QCS_C128) { using C = double2; CALL; } \
        else return set_error(QCS_ERR_ARG, "bad dtype"); \
    } while (0)

int misc_set_basis(void *buf, int n, int dtype, uint64_t index, cudaStream_t st) {
    if (n < 0 || n > 40 || index >> n) return set_error(QCS_ERR_ARG, "bad basis index");
    const size_t bytes = (dtype == QCS_C64 ? 8ull : 16ull) << n;
    if (cudaMemsetAsync(buf, 0, bytes, st) != cudaSuccess) return set_error(QCS_ERR_CUDA, "memset failed");
    DISPATCH_C(dtype, (set_one_kernel<C><<<1, 1, 0, st>>>((C *)buf, index)));
    return check_launch();
}

int misc_abs2(double *probs, const void *src, int n, int dtype, const double *norm_in, double *sum_out, void *ws,
              cudaStream_t st) {
    const uint64_t N = 1ull << n;
    if (n >= 1) {
        DISPATCH_C(dtype, (abs2_vec_kernel<C><<<ew_grid((N >> VecOf<C>::LV) / 4 + 1), EW_THREADS, 0, st>>>(
                              probs, (const C *)src, N >> VecOf<C>::LV, norm_in, sum_out, ws)));
        return check_launch();
    }
    DISPATCH_C(dtype, (abs2_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>(probs, (const C *)src, N, norm_in,
                                                                                sum_out, ws)));
    return check_launch();
}

int misc_dot(const void *a, const void *b, int n, int dtype, double *out2, void *ws, cudaStream_t st) {
    const uint64_t N = 1ull << n;
    DISPATCH_C(dtype, (dot_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>((const C *)a, (const C *)b, N, out2, ws)));
    return check_launch();
}

int misc_lincomb(void *dst, const void *a, const void *b, int n, int dtype, const double *alpha2, const double *beta2,
                 const double *norm_a, const double *norm_b, double *norm_out, void *ws, cudaStream_t st) {
    const uint64_t N = 1ull << n;
    const double bre = beta2 ? beta2[0] : 0.0, bim = beta2 ? beta2[1] : 0.0;
    DISPATCH_C(dtype, (lincomb_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>(
                          (C *)dst, (const C *)a, (const C *)b, N, alpha2[0], alpha2[1], bre, bim, norm_a, norm_b,
                          norm_out, ws)));
    return check_launch();
}

int misc_kron(void *dst, const void *a, int na, const void *b, int nb, int dtype, const double *norm_a,
              const double *norm_b, double *norm_out, void *ws, cudaStream_t st) {
    const uint64_t N = 1ull << (na + nb);
    if (nb >= 12 && na <= 6) {
        const uint64_t NVB = (1ull << nb) >> (dtype == QCS_C64 ? 1 : 0);
        if (dtype == QCS_C64)
            kron_small_a_kernel<float><<<ew_grid(NVB), EW_THREADS, 0, st>>>((float2 *)dst, (const float2 *)a, (const float2 *)b,
                                                                           na, nb, norm_a, norm_b, norm_out, ws);
        else if (dtype == QCS_C128)
            kron_small_a_kernel<double><<<ew_grid(NVB), EW_THREADS, 0, st>>>((double2 *)dst, (const double2 *)a,
                                                                            (const double2 *)b, na, nb, norm_a, norm_b,
                                                                            norm_out, ws);
        else return set_error(QCS_ERR_ARG, "bad dtype");
        return check_launch();
    }
    if (nb >= 1) {
        if (dtype == QCS_C64)
            kron_vec_kernel<float><<<ew_grid(N / 8 + 1), EW_THREADS, 0, st>>>((float2 *)dst, (const float2 *)a, (const float2 *)b,
                                                                             na, nb, norm_a, norm_b, norm_out, ws);
        else if (dtype == QCS_C128)
            kron_vec_kernel<double><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>((double2 *)dst, (const double2 *)a,
                                                                              (const double2 *)b, na, nb, norm_a, norm_b,
                                                                              norm_out, ws);
        else return set_error(QCS_ERR_ARG, "bad dtype");
        return check_launch();
    }
    DISPATCH_C(dtype, (kron_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>((C *)dst, (const C *)a, (const C *)b, na,
                                                                                nb, norm_a, norm_b, norm_out, ws)));
    return check_launch();
}

template <bool CONJ>
static int permute_impl(void *dst, const void *src, int n, int dtype, const int32_t *src_bit, cudaStream_t st) {
    BitList bl;
    if (int e = make_bitlist(bl, src_bit, n, n, false)) return e;
    const uint64_t N = 1ull << n;
    if (dst == src) return set_error(QCS_ERR_ARG, "permute_bits cannot run in place");
    if (n >= 15 && n <= 40) {
        BinMap pm;
        pm.nbits = n;
        for (int b = 0; b < 40; ++b) pm.contrib[b] = b < n ? 1ull << src_bit[b] : 0ull;
        const int lv = dtype == QCS_C64 ? 1 : 0;
        if (src_bit[0] != 0 && n >= 16) {
            int d0 = -1;                                 // the destination bit that takes source bit 0
            for (int b = 0; b < n; ++b)
                if (src_bit[b] == 0) d0 = b;
            uint64_t gridq = (uint64_t)device_info().sms * 8;
            const uint64_t nsegq = (N >> 2) >> (lv ? 11 : 10);
            if (gridq > nsegq) gridq = nsegq;
            DISPATCH_C(dtype, (permute_quad_kernel<C, CONJ><<<(unsigned)gridq, EW_THREADS, 0, st>>>((C *)dst, (const C *)src, pm, d0)));
            return check_launch();
        }
        uint64_t grid = (uint64_t)device_info().sms * 8;
        const uint64_t nseg = N >> (11 + lv);
        if (grid > nseg) grid = nseg;
        DISPATCH_C(dtype, (permute_stream_kernel<C, CONJ><<<(unsigned)grid, EW_THREADS, 0, st>>>((C *)dst, (const C *)src, pm)));
        return check_launch();
    }
    DISPATCH_C(dtype, (permute_kernel<C, CONJ><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>((C *)dst, (const C *)src, n, bl)));
    return check_launch();
}

int misc_permute(void *dst, const void *src, int n, int dtype, const int32_t *src_bit, bool conj, cudaStream_t st) {
    return conj ? permute_impl<true>(dst, src, n, dtype, src_bit, st) : permute_impl<false>(dst, src, n, dtype, src_bit, st);
}

template <typename T>
static int marginal_impl(double *out, const T *src, int n, const BitList &bl, const double *norm_in, void *ws,
                         int64_t ws_bytes, cudaStream_t st) {
    const uint64_t N = 1ull << n;
    const int m = bl.m;
    const int64_t avail = ws_bytes - WS_BYTES;
    double *hists = reinterpret_cast<double *>(reinterpret_cast<char *>(ws) + WS_BYTES);
    const bool small = m <= MARG_SMALL_BITS;
    const size_t hist_bytes = small ? sizeof(double) << m : 0;
    if (small && avail < (int64_t)hist_bytes) return set_error(QCS_ERR_ARG, "workspace too small for marginals");
    const int seg_bits = 11 + VecOf<T>::LV;
    if (n >= seg_bits + 3 && m < n && n <= 40) {
        // streaming kernel: per-thread register accumulators, contiguous segment ranges per CTA
        BinMap bm;
        bm.nbits = n;
        for (int b = 0; b < 40; ++b) bm.contrib[b] = 0;
        for (int j = 0; j < m; ++j) bm.contrib[bl.pos[j]] = 1ull << (m - 1 - j);
        const uint64_t nseg = N >> seg_bits;
        uint64_t grid = (uint64_t)device_info().sms * 8;
        if (grid > nseg) grid = nseg;
        if (grid > MAX_GRID) grid = MAX_GRID;
        if (small) {
            if ((int64_t)grid * (int64_t)hist_bytes > avail) grid = (uint64_t)(avail / (int64_t)hist_bytes);
            marginal_stream_kernel<T, true><<<(unsigned)grid, EW_THREADS, hist_bytes, st>>>(out, src, bm, m, norm_in, hists, ws);
        } else {
            if (cudaMemsetAsync(out, 0, sizeof(double) << m, st) != cudaSuccess) return set_error(QCS_ERR_CUDA, "memset failed");
            marginal_stream_kernel<T, false><<<(unsigned)grid, EW_THREADS, 0, st>>>(out, src, bm, m, norm_in, nullptr, ws);
        }
        return check_launch();
    }
    if (small) {
        unsigned grid = ew_grid(N / MARG_SEG + 1, 4);
        if ((int64_t)grid * (int64_t)hist_bytes > avail) grid = (unsigned)(avail / (int64_t)hist_bytes);
        marginal_small_kernel<T><<<grid, EW_THREADS, hist_bytes, st>>>(out, src, N, bl, norm_in, hists, ws);
    } else {
        const int exact = (m == n);
        if (!exact && cudaMemsetAsync(out, 0, sizeof(double) << m, st) != cudaSuccess)
            return set_error(QCS_ERR_CUDA, "memset failed");
        marginal_large_kernel<T><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>(out, src, N, bl, exact, norm_in);
    }
    return check_launch();
}

int misc_marginal(double *out, const void *src, int n, int dtype, const int32_t *bitpos, int m, const double *norm_in,
                  void *ws, int64_t ws_bytes, cudaStream_t st) {
    BitList bl;
    if (m < 1) return set_error(QCS_ERR_ARG, "must measure at least one bit");
    if (int e = make_bitlist(bl, bitpos, m, n, false)) return e;
    if (dtype == QCS_C64) return marginal_impl<float2>(out, (const float2 *)src, n, bl, norm_in, ws, ws_bytes, st);
    if (dtype == QCS_C128) return marginal_impl<double2>(out, (const double2 *)src, n, bl, norm_in, ws, ws_bytes, st);
    if (dtype == QCS_F64) return marginal_impl<double>(out, (const double *)src, n, bl, norm_in, ws, ws_bytes, st);
    return set_error(QCS_ERR_ARG, "bad dtype");
}

int misc_collapse(void *dst, const void *src, int n, int dtype, const int32_t *bitpos, int m, uint64_t outcome,
                  int remove, double scale, double *norm_out, void *ws, cudaStream_t st) {
    BitList bl;
    if (m < 1) return set_error(QCS_ERR_ARG, "must measure at least one bit");
    if (int e = make_bitlist(bl, bitpos, m, n, true)) return e;
    if (m < 64 && (outcome >> m)) return set_error(QCS_ERR_ARG, "outcome out of range");
    if (remove && dst == src) return set_error(QCS_ERR_ARG, "collapse with remove cannot run in place");
    uint64_t outbits = 0, mask = 0;
    for (int j = 0; j < m; ++j) {
        mask |= 1ull << bitpos[j];
        if ((outcome >> (m - 1 - j)) & 1ull) outbits |= 1ull << bitpos[j];
    }
    const uint64_t N = remove ? 1ull << (n - m) : 1ull << n;
    DISPATCH_C(dtype, (collapse_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>((C *)dst, (const C *)src, n, bl, outbits,
                                                                                    mask, remove, scale, norm_out, ws)));
    return check_launch();
}

int misc_diag_probs(double *probs, const void *rho, int n, int dtype, double *sum_out, void *ws, cudaStream_t st) {
    if (n < 1 || n > 31) return set_error(QCS_ERR_ARG, "bad qubit count");
    const uint64_t N = 1ull << n;
    DISPATCH_C(dtype, (diag_probs_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>(probs, (const C *)rho, n, sum_out, ws)));
    return check_launch();
}

int misc_outer(void *rho, const void *psi, int n, int dtype, double p, const double *norm_psi, int accumulate,
               cudaStream_t st) {
    if (n < 1 || n > 31) return set_error(QCS_ERR_ARG, "bad qubit count");
    const uint64_t N = 1ull << (2 * n);
    if (2 * n >= 15) {          // the segment-decomposed write stream
        const int lv = dtype == QCS_C64 ? 1 : 0;
        uint64_t grid = (uint64_t)device_info().sms * 8;
        const uint64_t nseg = (N >> lv) >> 11;
        if (grid > nseg) grid = nseg;
        DISPATCH_C(dtype, (outer_stream_kernel<C><<<(unsigned)grid, EW_THREADS, 0, st>>>((C *)rho, (const C *)psi, n, p, norm_psi,
                                                                                        accumulate)));
        return check_launch();
    }
    DISPATCH_C(dtype, (outer_kernel<C><<<ew_grid(N / 8 + 1), EW_THREADS, 0, st>>>((C *)rho, (const C *)psi, n, p, norm_psi,
                                                                                 accumulate)));
    return check_launch();
}

template <typename R>
static int gate1_launch(void *dst, const void *src, int n, const double *matrix, int tbit, uint64_t ctrl,
                        const double *norm_in, double *norm_out, void *ws, cudaStream_t st) {
    using C = typename CxT<R>::T;
    Gate1Args<R> g;
    for (int e = 0; e < 8; ++e) g.m[e] = (R)matrix[e];
    g.ctrl = ctrl; g.n = n; g.tbit = tbit;
    const uint64_t NV = (1ull << n) >> VecOf<C>::LV;
    const bool invec = tbit < VecOf<C>::LV;
    const unsigned grid = ew_grid((invec ? NV : NV / 2) / 4 + 1);
    if (invec) gate1_stream_kernel<R, true><<<grid, EW_THREADS, 0, st>>>((C *)dst, (const C *)src, g, norm_in, norm_out, ws);
    else gate1_stream_kernel<R, false><<<grid, EW_THREADS, 0, st>>>((C *)dst, (const C *)src, g, norm_in, norm_out, ws);
    return check_launch();
}

int misc_gate1(void *dst, const void *src, int n, int dtype, const double *matrix, int tbit, uint64_t ctrl,
               const double *norm_in, double *norm_out, void *ws, cudaStream_t st) {
    if (n < 2 || tbit < 0 || tbit >= n) return set_error(QCS_ERR_ARG, "bad target bit");
    if (dtype == QCS_C64) return gate1_launch<float>(dst, src, n, matrix, tbit, ctrl, norm_in, norm_out, ws, st);
    if (dtype == QCS_C128) return gate1_launch<double>(dst, src, n, matrix, tbit, ctrl, norm_in, norm_out, ws, st);
    return set_error(QCS_ERR_ARG, "bad dtype");
}

int misc_partial_trace(void *out, const void *rho, int d, int dtype, const int32_t *keep, int nkeep, cudaStream_t st) {
    if (d < 1 || d > 20) return set_error(QCS_ERR_ARG, "bad qubit count");
    if (nkeep < 1 || nkeep > d) return set_error(QCS_ERR_ARG, "must keep between 1 and d qubits");
    if (out == rho) return set_error(QCS_ERR_ARG, "partial trace cannot run in place");
    TraceMap tm;
    tm.nout_bits = 2 * nkeep;
    tm.ntraced = d - nkeep;
    uint32_t kept = 0;
    for (int j = 0; j < nkeep; ++j) {
        if (keep[j] < 0 || keep[j] >= d) return set_error(QCS_ERR_ARG, "qubit index out of range");
        if (kept & (1u << keep[j])) return set_error(QCS_ERR_ARG, "repeated qubit index");
        kept |= 1u << keep[j];
        // output qubit j (0 = most significant pair of the output index) <- qubit keep[j] of rho
        const int ob = 2 * (nkeep - 1 - j), ib = 2 * (d - 1 - keep[j]);
        tm.out_contrib[ob] = 1ull << ib;              // column bit
        tm.out_contrib[ob + 1] = 1ull << (ib + 1);    // row bit
    }
    int t = 0;
    for (int q = 0; q < d; ++q)
        if (!((kept >> q) & 1u)) tm.tr_contrib[t++] = 3ull << (2 * (d - 1 - q));
    const uint64_t nout = 1ull << tm.nout_bits;
    uint64_t grid = nout < (uint64_t)MAX_GRID ? nout : (uint64_t)MAX_GRID;
    DISPATCH_C(dtype, (partial_trace_kernel<C><<<(unsigned)grid, EW_THREADS, 0, st>>>((C *)out, (const C *)rho, tm)));
    return check_launch();
}

int misc_dense(void *dst, const void *src, int n, int dtype, const void *dev_matrix, int k, const int32_t *bitpos,
               const double *norm_in, double *norm_out, void *ws, cudaStream_t st) {
    if (k < 4 || k > n) return set_error(QCS_ERR_ARG, "dense path needs 4 <= k <= n");
    if (dst == src) return set_error(QCS_ERR_ARG, "dense path cannot run in place");
    BitList bl;
    if (int e = make_bitlist(bl, bitpos, k, n, true)) return e;
    const uint64_t NR = 1ull << (n - k), D = 1ull << k;
    const uint64_t gx = (NR + DT - 1) / DT, gy = D / DT;
    if (gy > 65535 || gx > 0x7fffffffull) return set_error(QCS_ERR_UNSUPPORTED, "operator too large for the dense path");
    dim3 grid((unsigned)gx, (unsigned)gy), block(DT, DT);
    if (dtype == QCS_C64)
        dense_kernel<float><<<grid, block, 0, st>>>((float2 *)dst, (const float2 *)src, (const float2 *)dev_matrix, n, k, bl, norm_in);
    else if (dtype == QCS_C128)
        dense_kernel<double><<<grid, block, 0, st>>>((double2 *)dst, (const double2 *)src, (const double2 *)dev_matrix, n, k, bl, norm_in);
    else return set_error(QCS_ERR_ARG, "bad dtype");
    if (int e = check_launch()) return e;
    if (norm_out) return misc_abs2(nullptr, dst, n, dtype, nullptr, norm_out, ws, st);
    return QCS_OK;
}

// ------------------------------------------------------------------------------------------------
// U_f (reference operators.py:809-853): |x>|y> -> |x>|y XOR f(x)> as an index permutation
// ------------------------------------------------------------------------------------------------
// dst[i] = src[i XOR (f(x(i)) << tbit)], x(i) = the input bits of i gathered MSB-first; f as a packed truth table.
// Reads and writes are 16-byte vectors wherever the target bit is not inside a vector; the permutation only moves
// amplitudes within pairs, so both streams stay coalesced.  The operator tensor (2^(2d) entries) is never built.
struct UfMap { int32_t xbit[32]; int32_t nx, tbit; };

template <typename C>
__global__ void __launch_bounds__(EW_THREADS) uf_kernel(C *__restrict__ dst, const C *__restrict__ src, uint64_t N, UfMap mp,
                                                        const uint32_t *__restrict__ table, const double *norm_in) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const double sc = load_inv_norm(norm_in);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
        uint32_t x = 0;
        for (int j = 0; j < mp.nx; ++j) x = (x << 1) | (uint32_t)((i >> mp.xbit[j]) & 1ull);
        const uint64_t flip = (uint64_t)((__ldg(&table[x >> 5]) >> (x & 31u)) & 1u) << mp.tbit;
        C v = src[i ^ flip];
        if (norm_in) { v.x = (decltype(v.x))(v.x * sc); v.y = (decltype(v.y))(v.y * sc); }
        dst[i] = v;
    }
}

int misc_apply_uf(void *dst, const void *src, int n, int dtype, const uint32_t *dev_table, const int32_t *xbits, int nx,
                  int tbit, const double *norm_in, cudaStream_t st) {
    if (nx < 1 || nx > 31 || nx >= n + 1) return set_error(QCS_ERR_ARG, "U_f needs 1..31 input qubits");
    if (tbit < 0 || tbit >= n) return set_error(QCS_ERR_ARG, "U_f target bit out of range");
    if (dst == src) return set_error(QCS_ERR_ARG, "U_f cannot run in place");
    UfMap mp;
    mp.nx = nx; mp.tbit = tbit;
    uint64_t seen = 1ull << tbit;
    for (int j = 0; j < nx; ++j) {
        if (xbits[j] < 0 || xbits[j] >= n || ((seen >> xbits[j]) & 1ull)) return set_error(QCS_ERR_ARG, "bad U_f input bit");
        seen |= 1ull << xbits[j];
        mp.xbit[j] = xbits[j];
    }
    const uint64_t N = 1ull << n;
    DISPATCH_C(dtype, (uf_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>((C *)dst, (const C *)src, N, mp, dev_table, norm_in)));
    return check_launch();
}

}  // namespace qcs
