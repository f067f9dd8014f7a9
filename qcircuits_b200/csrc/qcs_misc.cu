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

// ---- permute_bits (state.py:117,162) --------------------------------------------------------------
template <typename C>
__global__ void __launch_bounds__(EW_THREADS) permute_kernel(C *__restrict__ dst, const C *__restrict__ src, int n,
                                                             const __grid_constant__ BitList bl) {
    const uint64_t N = 1ull << n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t j = 0;
        for (int b = 0; b < n; ++b) j |= ((i >> b) & 1ull) << bl.pos[b];
        dst[i] = src[j];
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
    const double w = p * (norm_psi ? 1.0 / *norm_psi : 1.0);
    const uint64_t N = 1ull << (2 * n);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (uint64_t)gridDim.x * blockDim.x) {
        const C a = psi[compact_bits(i >> 1)];   // row index: odd bits
        const C b = psi[compact_bits(i)];        // column index: even bits
        double re = ((double)a.x * (double)b.x + (double)a.y * (double)b.y) * w;   // a * conj(b)
        double im = ((double)a.y * (double)b.x - (double)a.x * (double)b.y) * w;
        if (accumulate) { const C o = rho[i]; re += (double)o.x; im += (double)o.y; }
        C r; r.x = (decltype(r.x))re; r.y = (decltype(r.y))im;
        rho[i] = r;
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
    DISPATCH_C(dtype, (kron_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>((C *)dst, (const C *)a, (const C *)b, na,
                                                                                nb, norm_a, norm_b, norm_out, ws)));
    return check_launch();
}

int misc_permute(void *dst, const void *src, int n, int dtype, const int32_t *src_bit, cudaStream_t st) {
    BitList bl;
    if (int e = make_bitlist(bl, src_bit, n, n, false)) return e;
    const uint64_t N = 1ull << n;
    DISPATCH_C(dtype, (permute_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>((C *)dst, (const C *)src, n, bl)));
    return check_launch();
}

template <typename T>
static int marginal_impl(double *out, const T *src, int n, const BitList &bl, const double *norm_in, void *ws,
                         int64_t ws_bytes, cudaStream_t st) {
    const uint64_t N = 1ull << n;
    const int m = bl.m;
    if (m <= MARG_SMALL_BITS) {
        const size_t hist_bytes = sizeof(double) << m;
        const int64_t avail = ws_bytes - WS_BYTES;
        unsigned grid = ew_grid(N / MARG_SEG + 1, 4);
        if (avail < (int64_t)hist_bytes) return set_error(QCS_ERR_ARG, "workspace too small for marginals");
        if ((int64_t)grid * (int64_t)hist_bytes > avail) grid = (unsigned)(avail / (int64_t)hist_bytes);
        double *hists = reinterpret_cast<double *>(reinterpret_cast<char *>(ws) + WS_BYTES);
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
    DISPATCH_C(dtype, (outer_kernel<C><<<ew_grid(N / 4 + 1), EW_THREADS, 0, st>>>((C *)rho, (const C *)psi, n, p, norm_psi,
                                                                                 accumulate)));
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

}  // namespace qcs
