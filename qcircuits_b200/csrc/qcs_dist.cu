// Multi-GPU qubit exchange over NVLink peer memory (include/qcs.h: qcs_dist_*).
//
// A state of n_local + log2(P) qubits is sharded by its top log2(P) flat bits: shard r holds the amplitudes whose
// "rank bits" spell r.  A gate on a rank bit needs that qubit on a local bit first (north star (4): "a gate on a global
// qubit triggers an ... all-to-all qubit swap over NVLink").  The exchange of g (rank bit, local bit) pairs is
//     dst_r[i] = src_{r'}[i']   with   bit lb_j of i' = rank bit gb_j of r,   rank bit gb_j of r' = bit lb_j of i,
// every other bit unchanged.  ONE kernel per GPU does it: each thread computes the peer and the offset of its 16-byte
// vector and loads it straight from the peer's HBM over NVLink (P2P load), eight loads in flight per thread, and
// stores it locally with a streaming store -- no pack / unpack pass, no staging buffer, no NCCL on the data path.
// The 2^-g of the data that stays on its GPU is a local copy in the same kernel.
//
// Two hosts are served: a single process that drives all GPUs (qcs_dist_init / qcs_dist_swap / qcs_dist_destroy:
// peer access enabled between the devices, cross-device ordering with events), and one process per GPU
// (qcs_dist_swap_rank on pointers that the caller mapped with qcs_ipc_open; the caller orders the ranks, e.g. with a
// one-element all-reduce before and after -- qcircuits_b200/dist.py).
#include <cstdio>
#include <cstring>

#include "qcs_common.cuh"
#include "qcs_internal.h"

namespace qcs {

constexpr int DIST_MAX_RANKS = 16, DIST_MAX_PAIRS = 4, DIST_THREADS = 256, DIST_UNROLL = 8;

struct DistMap {
    const uint4 *peer[DIST_MAX_RANKS];   // pull: source shard of every rank; push: destination shard of every rank (addressable from this device)
    int32_t lb[DIST_MAX_PAIRS];          // local bit of pair j, in 16-byte vector units
    int32_t gb[DIST_MAX_PAIRS];          // rank bit of pair j
    int32_t g, rank;
    uint64_t nvec;                       // vectors per shard
};

// PUSH = false: dst is this rank's destination shard and the loads go to the peers (P2P loads); PUSH = true: `dst` is
// this rank's SOURCE shard, read locally, and the stores go to the peers' destination shards (the map is an involution,
// so the same index arithmetic serves both directions).
template <int UNROLL, bool PUSH>
__global__ void __launch_bounds__(DIST_THREADS) dist_pull_kernel(uint4 *__restrict__ dst, const __grid_constant__ DistMap mp) {
    const uint64_t chunk = (uint64_t)DIST_THREADS * UNROLL;
    const uint64_t nchunks = mp.nvec / chunk;        // (a power of two: the host guarantees it)
    // The low g bits of the loop counter become the TOP g bits of the chunk index: when the exchanged local bits are the
    // top ones (the usual case) consecutive iterations alternate between the block that stays on this GPU (an HBM copy)
    // and the blocks that come over NVLink, so the local copy hides under the remote reads instead of preceding them.
    const int cb = 63 - __clzll((long long)nchunks);
    const int rot = mp.g < cb ? mp.g : 0;
    for (uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const uint64_t cc = rot ? ((c & ((1ull << rot) - 1ull)) << (cb - rot)) | (c >> rot) : c;
        const uint64_t v0 = cc * chunk + threadIdx.x;
        uint4 val[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t v = v0 + (uint64_t)u * DIST_THREADS;
            uint64_t sv = v;
            uint32_t sr = (uint32_t)mp.rank;
            for (int j = 0; j < mp.g; ++j) {
                const uint32_t diff = ((uint32_t)(v >> mp.lb[j]) ^ ((uint32_t)mp.rank >> mp.gb[j])) & 1u;
                sv ^= (uint64_t)diff << mp.lb[j];
                sr ^= diff << mp.gb[j];
            }
            if (PUSH) {
                const uint4 x = __ldcs(dst + v);
                uint4 *q = const_cast<uint4 *>(mp.peer[sr]) + sv;
                if (sr == (uint32_t)mp.rank) __stcs(q, x);
                else *q = x;
            } else {
                val[u] = __ldcs(mp.peer[sr] + sv);       // read once: streaming
            }
        }
        if (!PUSH) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) __stcs(dst + v0 + (uint64_t)u * DIST_THREADS, val[u]);
        }
    }
}

static int fill_map(DistMap &mp, int nranks, int rank, int n_local, int dtype, const int32_t *global_bits,
                    const int32_t *local_bits, int g) {
    if (dtype != QCS_C64 && dtype != QCS_C128) return set_error(QCS_ERR_ARG, "bad dtype");
    int G = 0;
    while ((1 << G) < nranks) ++G;
    if (nranks < 2 || nranks > DIST_MAX_RANKS || (1 << G) != nranks) return set_error(QCS_ERR_ARG, "the number of ranks must be a power of two in 2..16");
    if (rank < 0 || rank >= nranks) return set_error(QCS_ERR_ARG, "rank out of range");
    if (g < 1 || g > G || g > DIST_MAX_PAIRS) return set_error(QCS_ERR_ARG, "bad number of exchanged bit pairs");
    const int lv = dtype == QCS_C64 ? 1 : 0;         // amplitudes per 16-byte vector, log2
    if (n_local - lv < 12 || n_local > 40) return set_error(QCS_ERR_UNSUPPORTED, "shard too small (or too large) for the exchange kernel");
    uint32_t seen_l = 0, seen_g = 0;
    uint64_t seen_l64 = 0;
    (void)seen_l;
    for (int j = 0; j < g; ++j) {
        if (global_bits[j] < 0 || global_bits[j] >= G || ((seen_g >> global_bits[j]) & 1u)) return set_error(QCS_ERR_ARG, "bad or repeated rank bit");
        if (local_bits[j] < lv || local_bits[j] >= n_local || ((seen_l64 >> local_bits[j]) & 1ull)) return set_error(QCS_ERR_ARG, "bad or repeated local bit");
        seen_g |= 1u << global_bits[j];
        seen_l64 |= 1ull << local_bits[j];
        mp.lb[j] = local_bits[j] - lv;
        mp.gb[j] = global_bits[j];
    }
    mp.g = g; mp.rank = rank;
    mp.nvec = (1ull << n_local) >> lv;
    return QCS_OK;
}

bool dist_push_mode() { return tune_int("QCS_DIST_PUSH", 1) != 0; }   // measured: remote stores 682 GB/s, remote loads 627 GB/s (2 x B200)

static int launch_pull(void *dst, const DistMap &mp, bool push, cudaStream_t st) {
    const int unroll = tune_int("QCS_DIST_UNROLL", DIST_UNROLL) == 16 ? 16 : 8;
    const uint64_t nchunks = mp.nvec / ((uint64_t)DIST_THREADS * unroll);
    uint64_t grid = (uint64_t)device_info().sms * (uint64_t)tune_int("QCS_DIST_CTAS", 8);
    if (grid > nchunks) grid = nchunks;
    if (push) dist_pull_kernel<8, true><<<(unsigned)grid, DIST_THREADS, 0, st>>>(reinterpret_cast<uint4 *>(dst), mp);
    else if (unroll == 16) dist_pull_kernel<16, false><<<(unsigned)grid, DIST_THREADS, 0, st>>>(reinterpret_cast<uint4 *>(dst), mp);
    else dist_pull_kernel<8, false><<<(unsigned)grid, DIST_THREADS, 0, st>>>(reinterpret_cast<uint4 *>(dst), mp);
    return cudaGetLastError() == cudaSuccess ? QCS_OK : set_error(QCS_ERR_CUDA, "exchange kernel launch failed");
}

int dist_swap_rank(void *dst, const void *const *peer_src, int nranks, int rank, int n_local, int dtype,
                   const int32_t *global_bits, const int32_t *local_bits, int g, int push, cudaStream_t st) {
    DistMap mp;
    std::memset(&mp, 0, sizeof(mp));
    if (int e = fill_map(mp, nranks, rank, n_local, dtype, global_bits, local_bits, g)) return e;
    for (int r = 0; r < nranks; ++r) {
        if (!peer_src[r]) return set_error(QCS_ERR_ARG, "null peer shard");
        if (peer_src[r] == dst) return set_error(QCS_ERR_ARG, "the exchange cannot run in place");
        mp.peer[r] = reinterpret_cast<const uint4 *>(peer_src[r]);
    }
    return launch_pull(dst, mp, push != 0, st);
}

// ---- one process, all GPUs -------------------------------------------------------------------------------------------
struct DistCtx {
    int ndev;
    int dev[DIST_MAX_RANKS];
    cudaEvent_t ready[DIST_MAX_RANKS], done[DIST_MAX_RANKS];
};

int dist_init(int ndev, const int *devs, DistCtx **out) {
    int G = 0;
    while ((1 << G) < ndev) ++G;
    if (!devs || !out || ndev < 2 || ndev > DIST_MAX_RANKS || (1 << G) != ndev) return set_error(QCS_ERR_ARG, "ndev must be a power of two in 2..16");
    int prev = 0;
    cudaGetDevice(&prev);
    DistCtx *h = new DistCtx();
    h->ndev = ndev;
    for (int a = 0; a < ndev; ++a) h->dev[a] = devs[a];
    for (int a = 0; a < ndev; ++a) {
        if (cudaSetDevice(devs[a]) != cudaSuccess) { delete h; cudaSetDevice(prev); return set_error(QCS_ERR_CUDA, "cudaSetDevice failed"); }
        for (int b = 0; b < ndev; ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[a], devs[b]);
            if (!can) { delete h; cudaSetDevice(prev); return set_error(QCS_ERR_PEER, "devices cannot access each other's memory"); }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { delete h; cudaSetDevice(prev); return set_error(QCS_ERR_PEER, "cudaDeviceEnablePeerAccess failed"); }
            cudaGetLastError();
        }
        cudaEventCreateWithFlags(&h->ready[a], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->done[a], cudaEventDisableTiming);
    }
    cudaSetDevice(prev);
    *out = h;
    return QCS_OK;
}

int dist_destroy(DistCtx *h) {
    if (!h) return QCS_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    for (int a = 0; a < h->ndev; ++a) {
        cudaSetDevice(h->dev[a]);
        cudaEventDestroy(h->ready[a]);
        cudaEventDestroy(h->done[a]);
    }
    cudaSetDevice(prev);
    delete h;
    return QCS_OK;
}

int dist_swap(DistCtx *h, void *const *dst_shards, const void *const *src_shards, int n_local, int dtype,
              const int32_t *global_bits, const int32_t *local_bits, int g, const cudaStream_t *streams) {
    if (!h || !dst_shards || !src_shards || !streams) return set_error(QCS_ERR_ARG, "null argument");
    int prev = 0;
    cudaGetDevice(&prev);
    int rc = QCS_OK;
    // every source shard is complete before any peer reads it
    for (int a = 0; a < h->ndev; ++a) {
        cudaSetDevice(h->dev[a]);
        cudaEventRecord(h->ready[a], streams[a]);
    }
    for (int a = 0; a < h->ndev && rc == QCS_OK; ++a) {
        cudaSetDevice(h->dev[a]);
        for (int b = 0; b < h->ndev; ++b)
            if (b != a) cudaStreamWaitEvent(streams[a], h->ready[b], 0);
        if (dist_push_mode())
            rc = dist_swap_rank(const_cast<void *>(src_shards[a]), const_cast<const void *const *>(dst_shards), h->ndev, a, n_local,
                                dtype, global_bits, local_bits, g, 1, streams[a]);
        else
            rc = dist_swap_rank(dst_shards[a], src_shards, h->ndev, a, n_local, dtype, global_bits, local_bits, g, 0, streams[a]);
        cudaEventRecord(h->done[a], streams[a]);
    }
    // ... and nobody reuses its source shard before every peer has read it
    for (int a = 0; a < h->ndev; ++a) {
        cudaSetDevice(h->dev[a]);
        for (int b = 0; b < h->ndev; ++b)
            if (b != a) cudaStreamWaitEvent(streams[a], h->done[b], 0);
    }
    cudaSetDevice(prev);
    return rc;
}

// ---- one process per GPU: CUDA IPC mapping of the peers' shards ----------------------------------------------------------
int ipc_get_handle(const void *dev_ptr, void *handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t hd;
    if (cudaIpcGetMemHandle(&hd, const_cast<void *>(dev_ptr)) != cudaSuccess) {
        cudaGetLastError();
        return set_error(QCS_ERR_PEER, "cudaIpcGetMemHandle failed (the pointer must be the base of a cudaMalloc allocation)");
    }
    std::memcpy(handle64, &hd, 64);
    return QCS_OK;
}

int ipc_open(const void *handle64, void **mapped) {
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handle64, 64);
    if (cudaIpcOpenMemHandle(mapped, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return set_error(QCS_ERR_PEER, "cudaIpcOpenMemHandle failed");
    }
    return QCS_OK;
}

int ipc_close(void *mapped) {
    if (cudaIpcCloseMemHandle(mapped) != cudaSuccess) {
        cudaGetLastError();
        return set_error(QCS_ERR_PEER, "cudaIpcCloseMemHandle failed");
    }
    return QCS_OK;
}

int dist_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return set_error(QCS_ERR_ARG, "bad allocation request");
    if (cudaMalloc(ptr, (size_t)bytes) != cudaSuccess) {
        cudaGetLastError();
        return set_error(QCS_ERR_NOMEM, "cudaMalloc failed");
    }
    return QCS_OK;
}

int dist_free(void *ptr) {
    if (ptr && cudaFree(ptr) != cudaSuccess) {
        cudaGetLastError();
        return set_error(QCS_ERR_CUDA, "cudaFree failed");
    }
    return QCS_OK;
}

}  // namespace qcs
