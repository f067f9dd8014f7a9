// tcgen05 probe (development tool, not part of libqcs): checks, on a real B200, the operand layouts, the
// instruction-descriptor flags and the issue rates that the tensor-core sweep of qcs_tc.cu relies on:
//   * A operand in tensor memory (lane = row, column = K element), written with tcgen05.st 32x32b
//   * B operand in shared memory, K-major, no swizzle: (n, k) at (n%8)*16 + (n/8)*SBO + (k/T)*LBO + (k%T)*esize
//   * D in tensor memory read back with tcgen05.ld 32x32b
//   * kind::tf32 (K = 8 per instruction) and kind::f16 (K = 16), the a_negate flag
//   * 3-term split accuracy against a float64 host product
//   * cycles per instruction for N = 32 .. 256, and tcgen05.ld / st throughput
// Every wait is bounded; a timeout sets an error flag instead of hanging the GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

#define R32(a, o) "r"(a[o + 0]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3]), "r"(a[o + 4]), "r"(a[o + 5]), "r"(a[o + 6]), "r"(a[o + 7]), \
    "r"(a[o + 8]), "r"(a[o + 9]), "r"(a[o + 10]), "r"(a[o + 11]), "r"(a[o + 12]), "r"(a[o + 13]), "r"(a[o + 14]), "r"(a[o + 15])
#define W32(a, o) "=r"(a[o + 0]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3]), "=r"(a[o + 4]), "=r"(a[o + 5]), "=r"(a[o + 6]), "=r"(a[o + 7]), \
    "=r"(a[o + 8]), "=r"(a[o + 9]), "=r"(a[o + 10]), "=r"(a[o + 11]), "=r"(a[o + 12]), "=r"(a[o + 13]), "=r"(a[o + 14]), "=r"(a[o + 15])

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), R32(v, 0) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : W32(v, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// bounded wait: returns false on timeout
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t it = 0; it < (1u << 22); ++it) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
template <int KIND>   // 0 = tf32, 1 = f16
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int KIND>
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ inline uint32_t make_idesc(int kind, int M, int N, int a_neg, int b_neg) {
    const uint32_t fmt = kind == 0 ? 2u : 0u;   // TF32 = 2, F16 = 0
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_neg << 13) | ((uint32_t)b_neg << 14) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

struct Params {
    const uint32_t *A;   // [128][KW] 32-bit words per row (KW = K for tf32, K/2 for f16 packed pairs)
    const uint32_t *B;   // raw bytes of the B operand image, copied verbatim to shared memory
    float *D;            // [128][N]
    int kind, N, K, KW, b_bytes, lbo, sbo, a_neg, b_neg, ss;   // ss: A from shared memory instead (image in A, a_bytes)
    int a_bytes, a_lbo, a_sbo;
    int *err;
    long long *cycles;
    int rounds;          // timing: repeat the whole K loop this many times (accumulating)
    int nacc;            // timing: cycle over this many independent accumulators (column base = i * N)
    int nthreads_issue;  // 1, or 2: warps 0 and 1 each issue half of the instructions
};

// one CTA, 128 threads.  dynamic smem: B image (and A image in SS mode)
template <int KIND>
__global__ void __launch_bounds__(128, 1) probe_kernel(Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char *sB = smem;
    unsigned char *sA = smem + ((p.b_bytes + 1023) / 1024) * 1024;
    for (int i = tid; i < p.b_bytes / 4; i += 128) reinterpret_cast<uint32_t *>(sB)[i] = p.B[i];
    if (p.ss) for (int i = tid; i < p.a_bytes / 4; i += 128) reinterpret_cast<uint32_t *>(sA)[i] = p.A[i];
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy smem writes -> async proxy (MMA)
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tbase = tmem_slot;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t colD = 0, colA = 256;
    if (!p.ss) {
        // thread = row: write this row's KW words to columns colA ..
        for (int c = 0; c < p.KW; c += 16) {
            uint32_t v[16];
            for (int j = 0; j < 16; ++j) v[j] = p.A[(size_t)tid * p.KW + c + j];
            tmem_st16(lane_base + colA + c, v);
        }
        tmem_wait_st();
    }
    fence_before();
    __syncthreads();
    fence_after();
    const int kstep = KIND == 0 ? 8 : 16;           // K per instruction
    const int a_cols_per_step = 8;                  // 32 bytes of K per row = 8 columns
    long long t0 = 0, t1 = 0;
    uint32_t elected = 0;
    if (warp == 0) asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(elected));
    if (warp == 0 && elected) {
        const uint32_t idesc = make_idesc(KIND, 128, p.N, p.a_neg, p.b_neg);
        t0 = clock64();
        int acc = 0;
        if (p.rounds > 1 && !p.ss && p.K / kstep == 8) {
            // timing path: descriptors are base + constant, the K loop is fully unrolled
            const uint64_t bd0 = make_desc(smem_u32(sB), p.lbo, p.sbo);
            const uint64_t bstep = (uint64_t)((2 * p.lbo) >> 4);
            const uint32_t d0 = tbase + colD, a0 = tbase + colA;
            const uint32_t dstep = p.nacc > 1 ? (uint32_t)p.N : 0u;
            for (int r = 0; r < p.rounds; ++r) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    mma_ts<KIND>(d0 + (ks & 1) * dstep, a0 + ks * 8, bd0 + ks * bstep, idesc, 1u);
            }
        } else
        for (int r = 0; r < p.rounds; ++r)
            for (int ks = 0; ks < p.K / kstep; ++ks) {
                const uint64_t bd = make_desc(smem_u32(sB) + ks * 2 * p.lbo, p.lbo, p.sbo);
                const uint32_t dcol = tbase + colD + acc * p.N;
                if (p.ss) {
                    const uint64_t ad = make_desc(smem_u32(sA) + ks * 2 * p.a_lbo, p.a_lbo, p.a_sbo);
                    mma_ss<KIND>(dcol, ad, bd, idesc, (r | ks) ? 1u : 0u);
                } else {
                    mma_ts<KIND>(dcol, tbase + colA + ks * a_cols_per_step, bd, idesc, (r | ks) ? 1u : 0u);
                }
                acc = acc + 1 == p.nacc ? 0 : acc + 1;
            }
        mma_commit(&bar);
    }
    const bool ok = mbar_wait(&bar, 0);
    if (warp == 0 && elected) { t1 = clock64(); p.cycles[0] = t1 - t0; }
    if (!ok) { if (tid == 0) atomicExch(p.err, 1); }
    fence_after();
    if (ok) {
        for (int c = 0; c < p.N; c += 16) {
            uint32_t v[16];
            tmem_ld16(lane_base + colD + c, v);
            tmem_wait_ld();
            for (int j = 0; j < 16; ++j) p.D[(size_t)tid * p.N + c + j] = __uint_as_float(v[j]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

// tcgen05.ld / st throughput: NW warps, each moving x16 blocks `iters` times
__global__ void __launch_bounds__(256, 1) ldst_kernel(int iters, long long *cycles, int *sink, int mode) {
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t lane_base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) v[j] = tid + j;
    for (int c = 0; c < 128; c += 16) tmem_st16(lane_base + c, v);
    tmem_wait_st();
    __syncthreads();
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
            for (int c = 0; c < 128; c += 16) { tmem_ld16(lane_base + c, v); tmem_wait_ld(); acc += v[3]; }
        } else if (mode == 1) {
            uint32_t w[16];
            tmem_ld16(lane_base + 0, v); tmem_ld16(lane_base + 16, w); tmem_wait_ld(); acc += v[3] + w[5];
            tmem_ld16(lane_base + 32, v); tmem_ld16(lane_base + 48, w); tmem_wait_ld(); acc += v[3] + w[5];
            tmem_ld16(lane_base + 64, v); tmem_ld16(lane_base + 80, w); tmem_wait_ld(); acc += v[3] + w[5];
            tmem_ld16(lane_base + 96, v); tmem_ld16(lane_base + 112, w); tmem_wait_ld(); acc += v[3] + w[5];
        } else {
            for (int c = 0; c < 128; c += 16) { v[0] = acc + c; tmem_st16(lane_base + c, v); }
            tmem_wait_st();
            acc += 1;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
    sink[tid] = (int)acc;
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

static float to_tf32(float x) {   // round to nearest, ties away (cvt.rna.tf32.f32)
    uint32_t u; memcpy(&u, &x, 4);
    u += 0x1000u; u &= 0xFFFFE000u;
    float r; memcpy(&r, &u, 4); return r;
}
static double urand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

struct Run { std::vector<float> D; long long cycles; int err; };

static Run run(int kind, int N, int K, const std::vector<uint32_t> &Aimg, int KW, const std::vector<uint32_t> &Bimg, int lbo, int sbo,
               int a_neg, int b_neg, int rounds = 1, int ss = 0, int a_lbo = 0, int a_sbo = 0, int nacc = 1) {
    Params p{};
    uint32_t *dA, *dB; float *dD; int *derr; long long *dcyc;
    CK(cudaMalloc(&dA, Aimg.size() * 4)); CK(cudaMalloc(&dB, Bimg.size() * 4)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMalloc(&derr, 4)); CK(cudaMalloc(&dcyc, 8));
    CK(cudaMemcpy(dA, Aimg.data(), Aimg.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bimg.data(), Bimg.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(derr, 0, 4)); CK(cudaMemset(dD, 0, 128 * N * 4));
    p.A = dA; p.B = dB; p.D = dD; p.kind = kind; p.N = N; p.K = K; p.KW = KW; p.b_bytes = (int)Bimg.size() * 4; p.lbo = lbo; p.sbo = sbo;
    p.a_neg = a_neg; p.b_neg = b_neg; p.err = derr; p.cycles = dcyc; p.rounds = rounds; p.ss = ss; p.a_bytes = (int)Aimg.size() * 4;
    p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.nacc = nacc;
    const size_t smem = ((p.b_bytes + 1023) / 1024) * 1024 + (ss ? p.a_bytes : 0) + 1024;
    if (kind == 0) { CK(cudaFuncSetAttribute(probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); probe_kernel<0><<<1, 128, smem>>>(p); }
    else { CK(cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); probe_kernel<1><<<1, 128, smem>>>(p); }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    Run r; r.D.resize(128 * N);
    CK(cudaMemcpy(r.D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&r.err, derr, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&r.cycles, dcyc, 8, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(derr); cudaFree(dcyc);
    return r;
}

// B image, K-major no swizzle.  esize = 4 (tf32) or 2 (f16); T = 16 / esize elements per 16-byte chunk
template <typename F>
static std::vector<uint32_t> b_image(int N, int K, int esize, int lbo, int sbo, F value_bits) {
    const int T = 16 / esize;
    size_t bytes = 0;
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        size_t o = (size_t)(n % 8) * 16 + (size_t)(n / 8) * sbo + (size_t)(k / T) * lbo + (size_t)(k % T) * esize;
        if (o + esize > bytes) bytes = o + esize;
    }
    bytes = (bytes + 15) / 16 * 16;
    std::vector<uint32_t> img(bytes / 4, 0);
    unsigned char *b = reinterpret_cast<unsigned char *>(img.data());
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        size_t o = (size_t)(n % 8) * 16 + (size_t)(n / 8) * sbo + (size_t)(k / T) * lbo + (size_t)(k % T) * esize;
        uint32_t v = value_bits(n, k);
        memcpy(b + o, &v, esize);
    }
    return img;
}
static uint32_t fbits(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
static uint32_t hbits(float x) { __half h = __float2half_rn(x); uint16_t u; memcpy(&u, &h, 2); return u; }
static float hval(float x) { return __half2float(__float2half_rn(x)); }

static void timing_probe() {
    for (int kind = 0; kind < 2; ++kind)
        for (int ss = 0; ss < 2; ++ss)
            for (int N : {32, 64, 128})
                for (int nacc : {1, 2, 4}) {
                    if (nacc * N > 256) continue;
                    const int K = 64, esz = kind == 0 ? 4 : 2, T = 16 / esz, lbo = 128, sbo = 128 * (K / T);
                    const int KW = kind == 0 ? K : K / 2;
                    auto Bimg = b_image(N, K, esz, lbo, sbo, [&](int, int) { return 0u; });
                    std::vector<uint32_t> Aimg = ss ? b_image(128, K, esz, lbo, sbo, [&](int, int) { return 0u; }) : std::vector<uint32_t>(128 * KW, 0);
                    const int rounds = 200;
                    Run r = run(kind, N, K, Aimg, KW, Bimg, lbo, sbo, 0, 0, rounds, ss, lbo, sbo, nacc);
                    const int ninstr = rounds * K / (kind == 0 ? 8 : 16);
                    printf("T7 %s %s M=128 N=%d accumulators %d: %.1f cycles/instr (timeout %d)\n", kind == 0 ? "tf32" : "f16", ss ? "SS" : "TS", N, nacc,
                           (double)r.cycles / ninstr, r.err);
                }
}

int main(int argc, char **argv) {
    srand(7);
    if (argc > 1 && !strcmp(argv[1], "timing")) { CK(cudaSetDevice(0)); timing_probe(); return 0; }
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);

    // ---------------- T1: tf32 TS correctness, two (LBO, SBO) conventions ----------------
    for (int conv = 0; conv < 2; ++conv) {
        const int N = 32, K = 32;
        const int lbo = conv == 0 ? 128 : 128 * (N / 8), sbo = conv == 0 ? 128 * (K / 4) : 128;
        std::vector<float> A(128 * K), B(N * K);
        for (auto &x : A) x = to_tf32((float)urand());
        for (auto &x : B) x = to_tf32((float)urand());
        std::vector<uint32_t> Aimg(128 * K);
        for (int i = 0; i < 128 * K; ++i) Aimg[i] = fbits(A[i]);
        auto Bimg = b_image(N, K, 4, lbo, sbo, [&](int n, int k) { return fbits(B[n * K + k]); });
        Run r = run(0, N, K, Aimg, K, Bimg, lbo, sbo, 0, 0);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(ref - r.D[m * N + n]));
        }
        printf("T1 tf32 TS conv%d (lbo %d sbo %d): timeout %d maxerr %.3e cycles %lld\n", conv, lbo, sbo, r.err, maxerr, r.cycles);
    }
    // ---------------- T1b: layout dumps with coded operands (tf32) ----------------
    {
        const int N = 32, K = 32, lbo = 128, sbo = 128 * (K / 4);
        // B coded, A one-hot column k0 -> D[m][n] = code(B element the hardware pairs with (n, k0))
        auto Bimg = b_image(N, K, 4, lbo, sbo, [&](int n, int k) { return fbits((float)(n * 32 + k)); });
        int bad = 0;
        for (int k0 = 0; k0 < K; ++k0) {
            std::vector<uint32_t> Aimg(128 * K, 0);
            for (int m = 0; m < 128; ++m) Aimg[m * K + k0] = fbits(1.0f);
            Run r = run(0, N, K, Aimg, K, Bimg, lbo, sbo, 0, 0);
            for (int n = 0; n < N; ++n) {
                const int code = (int)lrintf(r.D[5 * N + n]);
                if (code != n * 32 + k0) { if (bad < 24) printf("  Bmap: hw (n=%d,k=%d) read stored (n=%d,k=%d)\n", n, k0, code / 32, code % 32); ++bad; }
            }
        }
        printf("T1b B-layout mismatches %d\n", bad);
        // A coded, B one-hot: B[n][k] = (k == k0 && n == 0) -> D[m][0] = A[m][k0]
        bad = 0;
        std::vector<uint32_t> Aimg(128 * K);
        for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) Aimg[m * K + k] = fbits((float)(m + 128 * (k % 16)));
        for (int k0 = 0; k0 < K; ++k0) {
            auto Bi = b_image(N, K, 4, lbo, sbo, [&](int n, int k) { return fbits((n == 0 && k == k0) ? 1.0f : 0.0f); });
            Run r = run(0, N, K, Aimg, K, Bi, lbo, sbo, 0, 0);
            for (int m = 0; m < 128; ++m) {
                const float expect = (float)(m + 128 * (k0 % 16));
                if (r.D[m * N] != expect) { if (bad < 24) printf("  Amap: (m=%d,k=%d) expect %g got %g\n", m, k0, expect, r.D[m * N]); ++bad; }
            }
        }
        printf("T1b A-layout mismatches %d\n", bad);
    }
    // ---------------- T2: a_negate / b_negate ----------------
    for (int which = 0; which < 2; ++which) {
        const int N = 32, K = 32, lbo = 128, sbo = 128 * (K / 4);
        std::vector<float> A(128 * K), B(N * K);
        for (auto &x : A) x = to_tf32((float)urand());
        for (auto &x : B) x = to_tf32((float)urand());
        std::vector<uint32_t> Aimg(128 * K);
        for (int i = 0; i < 128 * K; ++i) Aimg[i] = fbits(A[i]);
        auto Bimg = b_image(N, K, 4, lbo, sbo, [&](int n, int k) { return fbits(B[n * K + k]); });
        Run r = run(0, N, K, Aimg, K, Bimg, lbo, sbo, which == 0, which == 1);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(-ref - r.D[m * N + n]));
        }
        printf("T2 negate %s: timeout %d maxerr vs -A.B %.3e\n", which == 0 ? "A" : "B", r.err, maxerr);
    }
    // ---------------- T3: f16 TS correctness (K = 16 per instruction, A packed 2 per column) ----------------
    for (int order = 0; order < 2; ++order) {
        const int N = 32, K = 32, lbo = 128, sbo = 128 * (K / 8);
        std::vector<float> A(128 * K), B(N * K);
        for (auto &x : A) x = hval((float)urand());
        for (auto &x : B) x = hval((float)urand());
        std::vector<uint32_t> Aimg(128 * K / 2);
        for (int m = 0; m < 128; ++m) for (int k = 0; k < K; k += 2) {
            const uint32_t lo = hbits(A[m * K + k + (order ? 1 : 0)]), hi = hbits(A[m * K + k + (order ? 0 : 1)]);
            Aimg[m * (K / 2) + k / 2] = lo | (hi << 16);
        }
        auto Bimg = b_image(N, K, 2, lbo, sbo, [&](int n, int k) { return hbits(B[n * K + k]); });
        Run r = run(1, N, K, Aimg, K / 2, Bimg, lbo, sbo, 0, 0);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(ref - r.D[m * N + n]));
        }
        printf("T3 f16 TS (low half = k %s): timeout %d maxerr %.3e\n", order ? "odd" : "even", r.err, maxerr);
    }
    // ---------------- T4: 3-term split accuracy, K = 64, N = 64 (tf32) and f16 with scaled low part ----------------
    {
        const int N = 64, K = 64, lbo = 128, sbo = 128 * (3 * K / 4);
        std::vector<float> A(128 * K), B(N * K);
        for (auto &x : A) x = (float)urand();
        for (auto &x : B) x = (float)(urand() * 0.5);
        // stacked along K: [A_lo | A_hi | A_hi] x [B_hi | B_lo | B_hi]  (small terms first)
        const int K3 = 3 * K;
        std::vector<uint32_t> Aimg(128 * K3);
        for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) {
            const float hi = to_tf32(A[m * K + k]), lo = A[m * K + k] - hi;
            Aimg[m * K3 + k] = fbits(lo); Aimg[m * K3 + K + k] = fbits(hi); Aimg[m * K3 + 2 * K + k] = fbits(hi);
        }
        auto Bimg = b_image(N, K3, 4, lbo, sbo, [&](int n, int k) {
            const int kk = k % K, part = k / K;
            const float hi = to_tf32(B[n * K + kk]), lo = B[n * K + kk] - hi;
            return fbits(part == 1 ? lo : hi);
        });
        Run r = run(0, N, K3, Aimg, K3, Bimg, lbo, sbo, 0, 0);
        auto Bimg1 = b_image(N, K, 4, 128, 128 * (K / 4), [&](int n, int k) { return fbits(B[n * K + k]); });
        std::vector<uint32_t> Aimg1(128 * K);
        for (int i = 0; i < 128 * K; ++i) Aimg1[i] = fbits(A[i]);
        Run r1 = run(0, N, K, Aimg1, K, Bimg1, 128, 128 * (K / 4), 0, 0);
        double e3 = 0, e1 = 0, ef = 0, scale = 0, bias3 = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; float f = 0;
            for (int k = 0; k < K; ++k) { ref += (double)A[m * K + k] * B[n * K + k]; f = fmaf(A[m * K + k], B[n * K + k], f); }
            e3 = fmax(e3, fabs(ref - r.D[m * N + n])); e1 = fmax(e1, fabs(ref - r1.D[m * N + n])); ef = fmax(ef, fabs(ref - f));
            scale = fmax(scale, fabs(ref));
            bias3 += (fabs((double)r.D[m * N + n]) - fabs(ref));
        }
        printf("T4 K=64 N=64: max|ref| %.3f  err 3xTF32 %.3e  plain TF32 %.3e  fp32 FMA %.3e  (timeouts %d %d) mean |.| bias %.3e\n", scale, e3, e1,
               ef, r.err, r1.err, bias3 / (128 * N));
    }
    {
        const int N = 64, K = 64, K3 = 3 * K, lbo = 128, sbo = 128 * (K3 / 8);
        std::vector<float> A(128 * K), B(N * K);
        for (auto &x : A) x = (float)urand();
        for (auto &x : B) x = (float)(urand() * 0.5);
        const float S = 2048.0f;   // 2^11
        // [A_lo*S | A_hi | A_hi] x [B_hi/S | B_lo | B_hi]
        std::vector<uint32_t> Aimg(128 * K3 / 2);
        auto a_part = [&](int m, int k3) {
            const int k = k3 % K, part = k3 / K;
            const float hi = hval(A[m * K + k]);
            return part == 0 ? (A[m * K + k] - hi) * S : hi;
        };
        for (int m = 0; m < 128; ++m) for (int k = 0; k < K3; k += 2)
            Aimg[m * (K3 / 2) + k / 2] = hbits(a_part(m, k)) | (hbits(a_part(m, k + 1)) << 16);
        auto Bimg = b_image(N, K3, 2, lbo, sbo, [&](int n, int k3) {
            const int k = k3 % K, part = k3 / K;
            const float hi = hval(B[n * K + k]);
            return hbits(part == 0 ? hi / S : (part == 1 ? B[n * K + k] - hi : hi));
        });
        Run r = run(1, N, K3, Aimg, K3 / 2, Bimg, lbo, sbo, 0, 0);
        double e3 = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            e3 = fmax(e3, fabs(ref - r.D[m * N + n]));
        }
        printf("T4 f16 split (x2^11 low part): err %.3e timeout %d\n", e3, r.err);
    }
    // ---------------- T5: cycles per instruction ----------------
    for (int kind = 0; kind < 2; ++kind)
        for (int N : {16, 32, 64, 128, 256}) {
            const int K = 64, esz = kind == 0 ? 4 : 2, T = 16 / esz, lbo = 128, sbo = 128 * (K / T);
            const int KW = kind == 0 ? K : K / 2;
            std::vector<uint32_t> Aimg(128 * KW, 0);
            auto Bimg = b_image(N, K, esz, lbo, sbo, [&](int, int) { return 0u; });
            const int rounds = 200;
            Run r = run(kind, N, K, Aimg, KW, Bimg, lbo, sbo, 0, 0, rounds);
            const int ninstr = rounds * K / (kind == 0 ? 8 : 16);
            printf("T5 %s TS M=128 N=%d: %d instr in %lld cycles = %.1f cycles/instr (timeout %d)\n", kind == 0 ? "tf32" : "f16", N, ninstr,
                   r.cycles, (double)r.cycles / ninstr, r.err);
        }
    // SS mode timing + correctness (tf32): A image K-major no swizzle in smem
    for (int N : {32, 64}) {
        const int K = 64, lbo = 128, sbo = 128 * (K / 4);
        std::vector<float> A(128 * K), B(N * K);
        for (auto &x : A) x = to_tf32((float)urand());
        for (auto &x : B) x = to_tf32((float)urand());
        auto Aimg = b_image(128, K, 4, lbo, sbo, [&](int m, int k) { return fbits(A[m * K + k]); });
        auto Bimg = b_image(N, K, 4, lbo, sbo, [&](int n, int k) { return fbits(B[n * K + k]); });
        Run r = run(0, N, K, Aimg, K, Bimg, lbo, sbo, 0, 0, 1, 1, lbo, sbo);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
            double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(ref - r.D[m * N + n]));
        }
        Run rt = run(0, N, K, Aimg, K, Bimg, lbo, sbo, 0, 0, 200, 1, lbo, sbo);
        printf("T5 tf32 SS N=%d: maxerr %.3e timeout %d; %.1f cycles/instr\n", N, maxerr, r.err, (double)rt.cycles / (200 * K / 8));
    }
    // ---------------- T6: tcgen05.ld / st throughput ----------------
    {
        long long *dcyc; int *sink; CK(cudaMalloc(&dcyc, 8)); CK(cudaMalloc(&sink, 1024));
        for (int nt : {128, 256})
            for (int mode = 0; mode < 3; ++mode) {
                const int iters = 200;
                ldst_kernel<<<1, nt>>>(iters, dcyc, sink, mode);
                CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
                long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
                const double bytes = (double)iters * nt * 128 * 4;
                printf("T6 %s %d threads: %.1f bytes/cycle (%lld cycles)\n", mode == 0 ? "ld x16 wait each" : (mode == 1 ? "ld x16 pairs" : "st x16"), nt,
                       bytes / c, c);
            }
        cudaFree(dcyc); cudaFree(sink);
    }
    printf("done\n");
    return 0;
}
