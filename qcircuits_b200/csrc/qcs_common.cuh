// Shared device helpers for libqcs (sm_100a).  See include/qcs.h for the ABI contract.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qcs {

template <typename R> struct CxT;
template <> struct CxT<float>  { using T = float2;  };
template <> struct CxT<double> { using T = double2; };

template <typename R> __device__ __forceinline__ typename CxT<R>::T cx(R re, R im) {
    typename CxT<R>::T v; v.x = re; v.y = im; return v;
}
// acc += m * a
template <typename C> __device__ __forceinline__ void cfma(C &acc, const C m, const C a) {
    acc.x = fma(m.x, a.x, acc.x); acc.x = fma(-m.y, a.y, acc.x);
    acc.y = fma(m.x, a.y, acc.y); acc.y = fma(m.y, a.x, acc.y);
}
template <typename C> __device__ __forceinline__ C cmul(const C m, const C a) {
    C r; r.x = m.x * a.x - m.y * a.y; r.y = m.x * a.y + m.y * a.x; return r;
}
template <typename C> __device__ __forceinline__ double cabs2(const C a) {
    return (double)a.x * (double)a.x + (double)a.y * (double)a.y;
}

// insert a 0 bit at position p of x (bits >= p move up by one)
__device__ __forceinline__ uint64_t insert_zero(uint64_t x, int p) {
    const uint64_t lo = x & ((1ull << p) - 1ull);
    return ((x >> p) << (p + 1)) | lo;
}
__device__ __forceinline__ uint32_t insert_zero32(uint32_t x, int p) {
    const uint32_t lo = x & ((1u << p) - 1u);
    return ((x >> p) << (p + 1)) | lo;
}

// ---- workspace layout -------------------------------------------------------------------------
// ws[0]                : uint32 ticket counter of the last-block reduction (self-resetting)
// ws + 256             : double partials[2 * MAX_GRID]
// ws + PROGRAM_OFFSET  : (unused, reserved)
constexpr int WS_PARTIALS_OFFSET = 256;
constexpr int MAX_GRID = 148 * 8;
// ws + 32640          : squared norm of the source of a tensor-core pass whose caller gave none
// ws + 32768          : tensor-core block-matrix images of the pass in flight, 8 slots x 8 KB (qcs_internal.h)
constexpr int64_t WS_BYTES = 32768 + 4 * 16384;
static_assert(WS_PARTIALS_OFFSET + 2 * MAX_GRID * 8 + 1024 <= 32640, "reduction scratch overlaps the matrix images");

// Deterministic grid-wide sum of NV doubles per thread: per-CTA partials are written to the
// workspace, the CTA that takes the last ticket adds them in a fixed order and stores the result.
// Must be called by all threads of all CTAs of the grid (gridDim.x <= MAX_GRID, blockDim.x <= 1024).
template <int NV>
__device__ __forceinline__ void grid_sum_publish(double (&acc)[NV], void *ws, double *out) {
    __shared__ double red[NV][32];
    __shared__ int last_flag;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        double a = acc[v];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
        if (lane == 0) red[v][warp] = a;
    }
    __syncthreads();
    unsigned *ticket = reinterpret_cast<unsigned *>(ws);
    double *partials = reinterpret_cast<double *>(reinterpret_cast<char *>(ws) + WS_PARTIALS_OFFSET);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double a = 0.0;
            for (int w = 0; w < nwarps; ++w) a += red[v][w];
            partials[(size_t)blockIdx.x * NV + v] = a;
        }
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        last_flag = (t == gridDim.x - 1u);
    }
    __syncthreads();
    if (last_flag) {
        __threadfence();
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double a = 0.0;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
                a += __ldcg(&partials[(size_t)b * NV + v]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
            __syncthreads();
            if (lane == 0) red[v][warp] = a;
            __syncthreads();
            if (threadIdx.x == 0) {
                double s = 0.0;
                for (int w = 0; w < nwarps; ++w) s += red[v][w];
                out[v] = s;
            }
        }
        if (threadIdx.x == 0) { *ticket = 0u; __threadfence(); }
    }
}

__device__ __forceinline__ double load_inv_norm(const double *norm_in) {
    return norm_in ? rsqrt(*norm_in) : 1.0;
}

}  // namespace qcs
