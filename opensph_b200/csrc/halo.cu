// Multi-GPU step inside the library: ghost (halo) exchange and the global time-step reduction with NCCL, queued on the
// context's stream between the kernels of one PredictorCorrector step, so a step costs ONE host synchronisation.
//
// The reference has no distributed path (core/thread/Scheduler.h:27 rules out MPI); this is the B200-native plumbing
// of SURVEY 8(e): slab neighbours exchange the dynamic neighbour inputs of their boundary bands with grouped
// ncclSend/ncclRecv over NVLink after the prediction and before the derivative evaluation; ncclAllReduce(min) combines
// the per-criterion time-step minima. NCCL is resolved at run time (dlopen of libnccl.so.2 -- inside a PyTorch process
// that is PyTorch's own copy), so libsphgpu has no link-time dependency on it.
#include "sphgpu_internal.h"
#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>

namespace sph {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* ncclApi() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) {
#define SPH_NCCL_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name))
            SPH_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
            SPH_NCCL_SYM(CommInitRank, "ncclCommInitRank");
            SPH_NCCL_SYM(CommDestroy, "ncclCommDestroy");
            SPH_NCCL_SYM(Send, "ncclSend");
            SPH_NCCL_SYM(Recv, "ncclRecv");
            SPH_NCCL_SYM(AllReduce, "ncclAllReduce");
            SPH_NCCL_SYM(GroupStart, "ncclGroupStart");
            SPH_NCCL_SYM(GroupEnd, "ncclGroupEnd");
            SPH_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef SPH_NCCL_SYM
        }
    }
    const bool ok = api.handle && api.GetUniqueId && api.CommInitRank && api.Send && api.Recv && api.AllReduce && api.GroupStart &&
                    api.GroupEnd;
    return ok ? &api : nullptr;
}

#define SPH_NCCL_CHECK(expr)                                                                                          \
    do {                                                                                                              \
        ncclResult_t _r = (expr);                                                                                     \
        if (_r != ncclSuccess) {                                                                                      \
            sph::setError(std::string(#expr) + ": " + (api->GetErrorString ? api->GetErrorString(_r) : "NCCL error")); \
            return SPHGPU_E_CUDA;                                                                                     \
        }                                                                                                             \
    } while (0)

// ---- exchange over peer memory (NVLink) instead of NCCL ---------------------------------------------------------------
// Every rank exports CUDA IPC handles of the 16 planes a halo record consists of and of a small MAILBOX; after
// sphgpu_peer_connect a rank holds mapped pointers to the planes of its two slab neighbours and to every rank's mailbox.
// One kernel then packs the send bands and writes them straight into the neighbours' ghost slots (no staging buffers, no
// unpack pass, both directions at once), and the four time-step minima are combined by every rank writing its values
// into every mailbox. Ordering is by sequence numbers with system-scope release / acquire:
//   ack[side]   -- "I have finished reading the ghosts you sent last time" (written into the neighbour's mailbox before a
//                  rank pushes; a rank pushes only after both neighbours' acks have arrived)
//   halo[side]  -- "my push number seq has landed in your ghost slots"
//   redSeq/redVal[parity][rank] -- contribution of `rank` to all-reduce number seq (double-buffered by parity: a rank can
//                  be at most one all-reduce ahead of another, because finishing one needs everybody's contribution)
constexpr int PEER_MAX_WORLD = 64;
struct PeerMailbox {
    unsigned long long ack[2];
    unsigned long long halo[2];
    unsigned long long redSeq[2][PEER_MAX_WORLD];
    unsigned long long redVal[2][PEER_MAX_WORLD][4];
};
struct PeerBlob { // what sphgpu_peer_export hands to the other ranks
    cudaIpcMemHandle_t planes[16];
    cudaIpcMemHandle_t mailbox;
    uint32_t n, recvLeft, recvRight, rank;
};
static_assert(sizeof(PeerBlob) <= SPHGPU_PEER_BLOB_BYTES, "SPHGPU_PEER_BLOB_BYTES too small");
struct PeerPlanes {
    double* f[16];
};
static const int HALO_PLANES[16] = { F_X, F_Y, F_Z, F_H, F_VX, F_VY, F_VZ, F_VH, F_RHO, F_U, F_S0, F_S1, F_S2, F_S3, F_S4, F_D };

struct HaloState {
    // peer-memory path (sphgpu_peer_connect); unused (nullptr) on the NCCL path
    PeerMailbox* mailbox = nullptr;                 // mine (device memory, exported)
    PeerMailbox* peerMailbox[PEER_MAX_WORLD] = {};  // everybody's, mapped (own entry = mailbox)
    void* mapped[2][16] = {};                       // planes of the left / right neighbour, mapped
    PeerPlanes peerPlanes[2] = {};
    uint32_t peerGhostFirst[2] = { 0, 0 };          // slot in the neighbour's arrays where my band goes
    // exchange overlapped with the interior part of k_correct_predict (sphgpu_run_pc): its own stream and two events
    cudaStream_t xstream = nullptr;
    cudaEvent_t evBands = nullptr, evArrived = nullptr;
    bool peerReady = false;
    unsigned long long haloSeq = 0, redSeqNo = 0;
    uint32_t* pushCounter = nullptr;                // device: blocks of the push kernel that have finished
    int guardAxis = -1;            // cut planes perpendicular to this axis (sphgpu_halo_set_guard); -1: no guard
    double guardLo = 0., guardHi = 0.;
    bool guardHasLo = false, guardHasHi = false;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int left = -1, right = -1;
    uint32_t sendLeft = 0, sendRight = 0, recvLeft = 0, recvRight = 0;
    double *bufSendL = nullptr, *bufSendR = nullptr, *bufRecvL = nullptr, *bufRecvR = nullptr;
};

// The bands a rank sends are fixed slot ranges chosen at decomposition time. An INTERIOR particle (neither band) that
// comes within reach R (h_i + h_max) / 2 of a cut plane would need neighbours the other rank never sends -- and would be
// needed by them. The guard measures the smallest head-room every exchange; a negative one raises a sticky flag that
// turns the next synchronising call into SPHGPU_E_STATE ("repartition"), so neighbours are never lost silently.
// h_max bound: the grid's h_max at the last list build times (1 + skin) -- the list reuse rebuilds before h grows more.
__global__ void __launch_bounds__(256) k_halo_guard(DevicePointers d, uint32_t first, uint32_t last, int axis, double lo, double hi, bool hasLo,
    bool hasHi, double kernelRadius, double skin) {
    const double hmax = d.grid->hmaxAll * (1. + skin);
    float margin = 3.0e38f;
    for (uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x; i < last; i += gridDim.x * blockDim.x) {
        const double x = d.f[F_X + axis][i];
        const double reach = 0.5 * kernelRadius * (d.f[F_H][i] + hmax);
        if (hasLo) {
            margin = fminf(margin, (float)((x - lo) / reach - 1.));
        }
        if (hasHi) {
            margin = fminf(margin, (float)((hi - x) / reach - 1.));
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        margin = fminf(margin, __shfl_xor_sync(0xffffffffu, margin, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (margin < 0.f) {
            d.listCtl->haloViolation = 1u;
            margin = 0.f;
        }
        atomicMin(&d.listCtl->haloMarginBits, __float_as_int(margin));
    }
}

__device__ __forceinline__ void storeRelease(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long loadAcquire(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

/// Spins until *p >= seq. A peer that never arrives (it failed, or the ranks' call sequences differ) must not hang the
/// GPU: after 20 s the wait gives up and raises ListCtlDev::peerTimeout, which fails the next synchronising call.
__device__ __forceinline__ void waitForSeq(const unsigned long long* p, unsigned long long seq, uint32_t* timeoutFlag) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t spins = 0;
    while (loadAcquire(p) < seq) {
        if ((++spins & 0x3ffu) == 0u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > 20000000000ull) {
                *timeoutFlag = 1u;
                return;
            }
        }
    }
}

/// One thread: tells both neighbours that their previous push has been consumed, then waits for their acks.
__global__ void k_peer_ack(PeerMailbox* mine, PeerMailbox* left, PeerMailbox* right, unsigned long long seq, uint32_t* timeoutFlag) {
    if (threadIdx.x != 0) {
        return;
    }
    // my left neighbour sees me on its right side (index 1) and vice versa
    if (left) storeRelease(&left->ack[1], seq);
    if (right) storeRelease(&right->ack[0], seq);
    if (left) waitForSeq(&mine->ack[0], seq, timeoutFlag);
    if (right) waitForSeq(&mine->ack[1], seq, timeoutFlag);
}

/// Packs both send bands and writes them into the neighbours' ghost slots over NVLink; the last block to finish publishes
/// the sequence number in both neighbours' mailboxes.
__global__ void __launch_bounds__(256) k_halo_push(DevicePointers d, uint32_t n, uint32_t sendLeft, uint32_t sendRight, PeerPlanes toLeft,
    PeerPlanes toRight, uint32_t firstLeft, uint32_t firstRight, bool solid, bool damage, PeerMailbox* left, PeerMailbox* right,
    unsigned long long seq, uint32_t* counter) {
    const uint32_t total = sendLeft + sendRight;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x) {
        const bool toL = k < sendLeft;
        const uint32_t src = toL ? k : n - sendRight + (k - sendLeft);
        const uint32_t dst = toL ? firstLeft + k : firstRight + (k - sendLeft);
        const PeerPlanes& out = toL ? toLeft : toRight;
        const int planes[16] = { F_X, F_Y, F_Z, F_H, F_VX, F_VY, F_VZ, F_VH, F_RHO, F_U, F_S0, F_S1, F_S2, F_S3, F_S4, F_D };
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const bool use = (c < 10) || (c < 15 && solid) || (c == 15 && damage);
            if (use) {
                out.f[c][dst] = d.f[planes[c]][src];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(counter, 1u) == gridDim.x - 1u) {
            *counter = 0u;
            __threadfence_system();
            if (left) storeRelease(&left->halo[1], seq);
            if (right) storeRelease(&right->halo[0], seq);
        }
    }
}

/// One thread: waits until both neighbours' pushes number seq have landed.
__global__ void k_halo_arrived(PeerMailbox* mine, bool hasLeft, bool hasRight, unsigned long long seq, uint32_t* timeoutFlag) {
    if (threadIdx.x != 0) {
        return;
    }
    if (hasLeft) waitForSeq(&mine->halo[0], seq, timeoutFlag);
    if (hasRight) waitForSeq(&mine->halo[1], seq, timeoutFlag);
}

struct MailboxList {
    PeerMailbox* p[PEER_MAX_WORLD];
};

/// All-reduce (min) of the four time-step minima: lane r writes this rank's values into rank r's mailbox, then waits for
/// rank r's values in its own; a warp reduction combines them. One warp.
__global__ void k_peer_allreduce_min(unsigned long long* vals, MailboxList boxes, PeerMailbox* mine, int rank, int world, unsigned long long seq,
    uint32_t* timeoutFlag) {
    const int r = threadIdx.x;
    const int par = (int)(seq & 1ull);
    unsigned long long v[4] = { ~0ull, ~0ull, ~0ull, ~0ull };
    if (r < world) {
        PeerMailbox* box = boxes.p[r];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            box->redVal[par][rank][k] = vals[k];
        }
        __threadfence_system();
        storeRelease(&box->redSeq[par][rank], seq);
        waitForSeq(&mine->redSeq[par][r], seq, timeoutFlag);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = mine->redVal[par][r][k];
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, v[k], o);
            v[k] = other < v[k] ? other : v[k];
        }
    }
    __syncwarp();
    if (r == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            vals[k] = v[k];
        }
    }
}

static int exchangePeer(sphgpu_ctx* ctx, HaloState* h) {
    const uint32_t n = ctx->n;
    const unsigned long long seq = ++h->haloSeq;
    PeerMailbox* left = h->left >= 0 ? h->peerMailbox[h->left] : nullptr;
    PeerMailbox* right = h->right >= 0 ? h->peerMailbox[h->right] : nullptr;
    k_peer_ack<<<1, 32, 0, ctx->stream>>>(h->mailbox, left, right, seq, &ctx->d.listCtl->peerTimeout);
    if (h->sendLeft + h->sendRight > 0) {
        const uint32_t blocks = std::min<uint32_t>((h->sendLeft + h->sendRight + 255) / 256, 592u);
        k_halo_push<<<blocks, 256, 0, ctx->stream>>>(ctx->d, n, h->sendLeft, h->sendRight, h->peerPlanes[0], h->peerPlanes[1],
            h->peerGhostFirst[0], h->peerGhostFirst[1], ctx->solid, ctx->hasDamage, h->sendLeft ? left : nullptr, h->sendRight ? right : nullptr,
            seq, h->pushCounter);
    }
    k_halo_arrived<<<1, 32, 0, ctx->stream>>>(h->mailbox, h->left >= 0 && h->recvLeft > 0, h->right >= 0 && h->recvRight > 0, seq,
        &ctx->d.listCtl->peerTimeout);
    ctx->launches += 3;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

/// Global minimum of the four per-criterion time steps in d.tsd (bit patterns of positive doubles order like the values).
static int allReduceTimestep(sphgpu_ctx* ctx, HaloState* h, NcclApi* api) {
    if (h->peerReady) {
        MailboxList boxes;
        for (int r = 0; r < PEER_MAX_WORLD; ++r) {
            boxes.p[r] = r < h->world ? h->peerMailbox[r] : nullptr;
        }
        k_peer_allreduce_min<<<1, 32, 0, ctx->stream>>>(ctx->d.tsd->minBits, boxes, h->mailbox, h->rank, h->world, ++h->redSeqNo,
            &ctx->d.listCtl->peerTimeout);
        ctx->launches += 1;
        SPH_CUDA_CHECK(cudaGetLastError());
        return SPHGPU_OK;
    }
    if (api->AllReduce(ctx->d.tsd, ctx->d.tsd, 4, ncclUint64, ncclMin, h->comm, ctx->stream) != ncclSuccess) {
        setError("ncclAllReduce failed");
        return SPHGPU_E_CUDA;
    }
    return SPHGPU_OK;
}

/// The halo guard: reads the INTERIOR particles (everything between the send bands).
static int exchangeGuard(sphgpu_ctx* ctx, HaloState* h) {
    const uint32_t n = ctx->n;
    if (h->guardAxis >= 0 && n > h->sendLeft + h->sendRight) {
        const int inf = 0x7f7fffff; // FLT_MAX
        SPH_CUDA_CHECK(cudaMemcpyAsync(&ctx->d.listCtl->haloMarginBits, &inf, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        k_halo_guard<<<296, 256, 0, ctx->stream>>>(ctx->d, h->sendLeft, n - h->sendRight, h->guardAxis, h->guardLo, h->guardHi, h->guardHasLo,
            h->guardHasHi, ctx->prm.kernel_radius, ctx->listSkin > 0. ? ctx->listSkin : 0.);
        ctx->launches += 1;
    }
    return SPHGPU_OK;
}

/// The transfer: reads the send bands, writes the ghost slots (of this rank: NCCL unpack; of the neighbours: peer push).
/// withGuard = false when the caller runs exchangeGuard itself (overlapped exchange).
static int exchange(sphgpu_ctx* ctx, bool withGuard = true) {
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    NcclApi* api = ncclApi();
    const uint32_t n = ctx->n;
    if (h->sendLeft + h->sendRight > n || (uint64_t)n + h->recvLeft + h->recvRight > ctx->capacity) {
        setError("the halo configuration does not fit the current particle count: call sphgpu_halo_configure again");
        return SPHGPU_E_STATE;
    }
    int rc;
    if (!h->peerReady) {
        if ((rc = launchHalo(ctx, true, 0, h->sendLeft, h->bufSendL)) != SPHGPU_OK) return rc;
        if ((rc = launchHalo(ctx, true, n - h->sendRight, h->sendRight, h->bufSendR)) != SPHGPU_OK) return rc;
    }
    if (withGuard && (rc = exchangeGuard(ctx, h)) != SPHGPU_OK) {
        return rc;
    }
    if (h->peerReady) {
        return exchangePeer(ctx, h);
    }
    const size_t w = SPHGPU_HALO_DOUBLES;
    // (every operation of the group is attempted and the group is always closed, also after an error)
    ncclResult_t nr = api->GroupStart();
    auto keep = [&](ncclResult_t r) {
        if (nr == ncclSuccess) {
            nr = r;
        }
    };
    if (nr == ncclSuccess) {
        if (h->left >= 0) {
            if (h->sendLeft) keep(api->Send(h->bufSendL, w * h->sendLeft, ncclFloat64, h->left, h->comm, ctx->stream));
            if (h->recvLeft) keep(api->Recv(h->bufRecvL, w * h->recvLeft, ncclFloat64, h->left, h->comm, ctx->stream));
        }
        if (h->right >= 0) {
            if (h->sendRight) keep(api->Send(h->bufSendR, w * h->sendRight, ncclFloat64, h->right, h->comm, ctx->stream));
            if (h->recvRight) keep(api->Recv(h->bufRecvR, w * h->recvRight, ncclFloat64, h->right, h->comm, ctx->stream));
        }
        keep(api->GroupEnd());
    }
    if (nr != ncclSuccess) {
        setError(std::string("NCCL halo exchange: ") + (api->GetErrorString ? api->GetErrorString(nr) : "NCCL error"));
        return SPHGPU_E_CUDA;
    }
    if ((rc = launchHalo(ctx, false, n, h->recvLeft, h->bufRecvL)) != SPHGPU_OK) return rc;
    if ((rc = launchHalo(ctx, false, n + h->recvLeft, h->recvRight, h->bufRecvR)) != SPHGPU_OK) return rc;
    ctx->launches += 4;
    return SPHGPU_OK;
}

static void closePeers(HaloState* h) {
    for (int side = 0; side < 2; ++side) {
        for (int c = 0; c < 16; ++c) {
            if (h->mapped[side][c]) {
                cudaIpcCloseMemHandle(h->mapped[side][c]);
                h->mapped[side][c] = nullptr;
            }
        }
    }
    for (int r = 0; r < PEER_MAX_WORLD; ++r) {
        if (h->peerMailbox[r] && h->peerMailbox[r] != h->mailbox) {
            cudaIpcCloseMemHandle(h->peerMailbox[r]);
        }
        h->peerMailbox[r] = nullptr;
    }
    h->peerReady = false;
}

void invalidateHalo(sphgpu_ctx* ctx) {
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    if (h) {
        h->peerReady = false;
        h->sendLeft = h->sendRight = 0xffffffffu; // fails the range check of exchange()
        h->recvLeft = h->recvRight = 0u;
    }
}

void destroyHalo(sphgpu_ctx* ctx) {
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    if (!h) {
        return;
    }
    cudaFree(h->bufSendL);
    cudaFree(h->bufSendR);
    cudaFree(h->bufRecvL);
    cudaFree(h->bufRecvR);
    closePeers(h);
    cudaFree(h->mailbox);
    cudaFree(h->pushCounter);
    if (h->xstream) cudaStreamDestroy(h->xstream);
    if (h->evBands) cudaEventDestroy(h->evBands);
    if (h->evArrived) cudaEventDestroy(h->evArrived);
    NcclApi* api = ncclApi();
    if (api && h->comm && api->CommDestroy) {
        api->CommDestroy(h->comm);
    }
    delete h;
    ctx->halo = nullptr;
}

} // namespace sph

using namespace sph;

extern "C" {

int sphgpu_comm_unique_id(void* out128) {
    NcclApi* api = ncclApi();
    if (!api || !out128) {
        setError("NCCL (libnccl.so.2) is not available");
        return SPHGPU_E_INVALID;
    }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    SPH_NCCL_CHECK(api->GetUniqueId(static_cast<ncclUniqueId*>(out128)));
    return SPHGPU_OK;
}

int sphgpu_comm_init(sphgpu_ctx* ctx, const void* id128, int rank, int world) {
    NcclApi* api = ncclApi();
    if (!ctx || !id128 || !api) {
        setError("NCCL (libnccl.so.2) is not available or null argument");
        return SPHGPU_E_INVALID;
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    destroyHalo(ctx);
    HaloState* h = new HaloState();
    h->rank = rank;
    h->world = world;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    {
        const ncclResult_t r = api->CommInitRank(&h->comm, world, id, rank);
        if (r != ncclSuccess) {
            delete h;
            setError(std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "NCCL error"));
            return SPHGPU_E_CUDA;
        }
    }
    ctx->halo = h;
    return SPHGPU_OK;
}

int sphgpu_halo_configure(sphgpu_ctx* ctx, int left_rank, int right_rank, uint32_t send_left, uint32_t send_right, uint32_t recv_left,
    uint32_t recv_right) {
    if (!ctx || !ctx->halo) {
        setError("sphgpu_comm_init must be called first");
        return SPHGPU_E_STATE;
    }
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    if (ctx->balsara || ctx->xsph || ctx->deltasph || ctx->stressAv) {
        // these terms read results of the PREVIOUS evaluation of their neighbours (div v / rot v, the velocity correction, the
        // density gradient) or per-particle constants of their own (the spacing kernel), which the ghost bands do not carry
        setError("the Balsara switch, the XSph term, the delta-SPH terms and the artificial stress are not available on decomposed runs");
        return SPHGPU_E_INVALID;
    }
    if ((uint64_t)ctx->n + recv_left + recv_right > ctx->capacity || send_left + send_right > ctx->n) {
        setError("halo sizes exceed the context capacity");
        return SPHGPU_E_INVALID;
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    h->left = left_rank;
    h->right = right_rank;
    h->sendLeft = left_rank >= 0 ? send_left : 0;
    h->sendRight = right_rank >= 0 ? send_right : 0;
    h->recvLeft = left_rank >= 0 ? recv_left : 0;
    h->recvRight = right_rank >= 0 ? recv_right : 0;
    const size_t w = SPHGPU_HALO_DOUBLES * sizeof(double);
    cudaFree(h->bufSendL); cudaFree(h->bufSendR); cudaFree(h->bufRecvL); cudaFree(h->bufRecvR);
    SPH_CUDA_CHECK(cudaMalloc(&h->bufSendL, std::max<size_t>(w * h->sendLeft, 16)));
    SPH_CUDA_CHECK(cudaMalloc(&h->bufSendR, std::max<size_t>(w * h->sendRight, 16)));
    SPH_CUDA_CHECK(cudaMalloc(&h->bufRecvL, std::max<size_t>(w * h->recvLeft, 16)));
    SPH_CUDA_CHECK(cudaMalloc(&h->bufRecvR, std::max<size_t>(w * h->recvRight, 16)));
    ctx->nActive = ctx->n + h->recvLeft + h->recvRight;
    h->peerReady = false; // the neighbours' ghost offsets are stale: sphgpu_peer_export / sphgpu_peer_connect again
    ctx->listsDirty = true; // another set of ghosts
    SPH_CUDA_CHECK(cudaMemsetAsync(&ctx->d.listCtl->haloViolation, 0, sizeof(uint32_t), ctx->stream));
    return SPHGPU_OK;
}

int sphgpu_peer_export(sphgpu_ctx* ctx, void* blob) {
    if (!ctx || !ctx->halo || !blob) {
        setError("sphgpu_comm_init / sphgpu_halo_configure must be called first");
        return SPHGPU_E_STATE;
    }
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!h->mailbox) {
        SPH_CUDA_CHECK(cudaMalloc(&h->mailbox, sizeof(PeerMailbox)));
        SPH_CUDA_CHECK(cudaMemset(h->mailbox, 0, sizeof(PeerMailbox)));
        SPH_CUDA_CHECK(cudaMalloc(&h->pushCounter, sizeof(uint32_t)));
        SPH_CUDA_CHECK(cudaMemset(h->pushCounter, 0, sizeof(uint32_t)));
    }
    PeerBlob b;
    std::memset(&b, 0, sizeof(b));
    for (int c = 0; c < 16; ++c) {
        SPH_CUDA_CHECK(cudaIpcGetMemHandle(&b.planes[c], ctx->d.f[HALO_PLANES[c]]));
    }
    SPH_CUDA_CHECK(cudaIpcGetMemHandle(&b.mailbox, h->mailbox));
    b.n = ctx->n;
    b.recvLeft = h->recvLeft;
    b.recvRight = h->recvRight;
    b.rank = (uint32_t)h->rank;
    std::memset(blob, 0, SPHGPU_PEER_BLOB_BYTES);
    std::memcpy(blob, &b, sizeof(b));
    return SPHGPU_OK;
}

int sphgpu_peer_connect(sphgpu_ctx* ctx, const void* blobs, int world) {
    if (!ctx || !ctx->halo || !blobs) {
        setError("sphgpu_comm_init / sphgpu_halo_configure must be called first");
        return SPHGPU_E_STATE;
    }
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    if (world != h->world || world > PEER_MAX_WORLD || !h->mailbox) {
        setError("sphgpu_peer_connect: world size mismatch, too many ranks, or sphgpu_peer_export was not called");
        return SPHGPU_E_INVALID;
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    closePeers(h);
    const char* base = static_cast<const char*>(blobs);
    auto blobOf = [&](int r) {
        PeerBlob b;
        std::memcpy(&b, base + (size_t)r * SPHGPU_PEER_BLOB_BYTES, sizeof(b));
        return b;
    };
    for (int r = 0; r < world; ++r) {
        if (r == h->rank) {
            h->peerMailbox[r] = h->mailbox;
            continue;
        }
        const PeerBlob b = blobOf(r);
        void* p = nullptr;
        SPH_CUDA_CHECK(cudaIpcOpenMemHandle(&p, b.mailbox, cudaIpcMemLazyEnablePeerAccess));
        h->peerMailbox[r] = static_cast<PeerMailbox*>(p);
    }
    const int nb[2] = { h->left, h->right };
    for (int side = 0; side < 2; ++side) {
        if (nb[side] < 0) {
            continue;
        }
        const PeerBlob b = blobOf(nb[side]);
        for (int c = 0; c < 16; ++c) {
            SPH_CUDA_CHECK(cudaIpcOpenMemHandle(&h->mapped[side][c], b.planes[c], cudaIpcMemLazyEnablePeerAccess));
            h->peerPlanes[side].f[c] = static_cast<double*>(h->mapped[side][c]);
        }
        // my band arrives behind the neighbour's own particles; the left neighbour files me as ITS right neighbour, i.e.
        // behind the ghosts it receives from its left
        h->peerGhostFirst[side] = side == 0 ? b.n + b.recvLeft : b.n;
        const uint32_t expect = side == 0 ? b.recvRight : b.recvLeft;
        const uint32_t mine = side == 0 ? h->sendLeft : h->sendRight;
        if (expect != mine) {
            setError("sphgpu_peer_connect: the neighbour expects a different number of ghosts than this rank sends");
            return SPHGPU_E_INVALID;
        }
    }
    h->peerReady = true;
    return SPHGPU_OK;
}

int sphgpu_halo_set_guard(sphgpu_ctx* ctx, int axis, double lo_plane, double hi_plane, int has_lo, int has_hi) {
    if (!ctx || !ctx->halo) {
        setError("sphgpu_comm_init must be called first");
        return SPHGPU_E_STATE;
    }
    if (axis < -1 || axis > 2) {
        setError("axis must be 0, 1, 2 or -1 (guard off)");
        return SPHGPU_E_INVALID;
    }
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    h->guardAxis = axis;
    h->guardLo = lo_plane;
    h->guardHi = hi_plane;
    h->guardHasLo = has_lo != 0;
    h->guardHasHi = has_hi != 0;
    return SPHGPU_OK;
}

int sphgpu_halo_margin(sphgpu_ctx* ctx, double* margin) {
    if (!ctx || !margin) {
        setError("null argument");
        return SPHGPU_E_INVALID;
    }
    *margin = ctx->haloMargin;
    return SPHGPU_OK;
}

int sphgpu_halo_exchange(sphgpu_ctx* ctx) {
    if (!ctx || !ctx->halo) {
        setError("halo exchange is not configured");
        return SPHGPU_E_STATE;
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    return exchange(ctx);
}

int sphgpu_run_pc(sphgpu_ctx* ctx, uint32_t steps, double dt, double max_dt, sphgpu_stats* stats, sphgpu_timestep* history) {
    if (!ctx || (steps > 0 && !history)) {
        setError("null argument");
        return SPHGPU_E_INVALID;
    }
    if (!ctx->stateUploaded && ctx->n > 0) {
        setError("run called before any state was uploaded");
        return SPHGPU_E_STATE;
    }
    if (steps == 0) {
        return SPHGPU_OK;
    }
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    NcclApi* api = h ? ncclApi() : nullptr;
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    StepRecordDev* histDev = nullptr;
    SPH_CUDA_CHECK(cudaMalloc(&histDev, sizeof(StepRecordDev) * steps));
    const StepStateDev init = { dt, ctx->lastDt, ctx->lastDtInit ? 1u : 0u, 0u, dt };
    int rc = SPHGPU_OK;
    cudaError_t ce = cudaMemcpyAsync(ctx->d.stepState, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream);
    ctx->d.dtDev = &ctx->d.stepState->dt;
    ctx->launches = 0;
    // Only the derivative criterion reads what the corrector changes (values and clamped derivatives). Without it the
    // criteria are evaluated BEFORE the corrector -- bit-identical results -- so the next time step is known when the
    // corrector runs and the corrector of step s and the predictor of step s + 1 are one kernel (k_correct_predict).
    const bool fused = (ctx->prm.criteria & SPHGPU_CRIT_DERIVATIVES) == 0u;
    // SPHGPU_HALO_OVERLAP=1: the exchange of step s + 1 runs beside the interior half of k_correct_predict. Measured at
    // 10.6 M particles: 6.63 against 6.64 ms per step on 2 GPUs, but 2.35 against 2.21 ms on 8 -- the three short kernels,
    // two events and the earlier acknowledgement cost more than the 0.1 ms of exchange they hide -- so it is off by default.
    bool overlap = false, exchanged = false;
    if (h && fused) {
        const char* env = std::getenv("SPHGPU_HALO_OVERLAP");
        overlap = env && env[0] == '1' && h->peerReady; // (NCCL calls of one communicator stay on one stream)
        if (overlap && !h->xstream) {
            if (cudaStreamCreateWithFlags(&h->xstream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&h->evBands, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&h->evArrived, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                overlap = false;
            }
        }
    }
    auto beginLastStep = [&]() { // the events time the last step
        ce = cudaEventRecord(ctx->ev[4], ctx->stream);
        ctx->launches = 0;
    };
    if (fused) {
        if (steps == 1) {
            beginLastStep();
        }
        rc = launchPredict(ctx, 0.);
    }
    for (uint32_t s = 0; s < steps && rc == SPHGPU_OK && ce == cudaSuccess; ++s) {
        if (!fused) {
            if (s + 1 == steps) {
                beginLastStep();
            }
            rc = launchPredict(ctx, 0.);
        }
        if (rc == SPHGPU_OK && h && !exchanged) {
            cudaEventRecord(ctx->ev[6], ctx->stream);
            rc = exchange(ctx); // ghosts carry the PREDICTED state
            cudaEventRecord(ctx->ev[7], ctx->stream);
        }
        exchanged = false;
        if (rc == SPHGPU_OK) rc = enqueueIntegrate(ctx);
        if (!fused) {
            if (rc == SPHGPU_OK) rc = launchCorrect(ctx, 0.);
        }
        if (rc == SPHGPU_OK) rc = launchCriteria(ctx);
        if (rc == SPHGPU_OK && h) {
            rc = allReduceTimestep(ctx, h, api);
        }
        if (rc == SPHGPU_OK) rc = launchFinishTimestep(ctx, max_dt, histDev, s);
        if (fused && rc == SPHGPU_OK) {
            if (s + 1 < steps) {
                if (s + 2 == steps) {
                    beginLastStep(); // (the timed step then holds the corrector of the step before instead of its own)
                }
                if (h && overlap) {
                    // Decomposed run: the send bands first, then their exchange on a second stream WHILE the interior is
                    // corrected / predicted. The neighbours' pushes land in the ghost slots, which the interior kernel
                    // does not touch; the guard reads interior particles and follows them on the main stream.
                    const uint32_t n = ctx->n;
                    rc = launchCorrectPredictRange(ctx, 0u, h->sendLeft);
                    if (rc == SPHGPU_OK) rc = launchCorrectPredictRange(ctx, n - h->sendRight, n);
                    if (ce == cudaSuccess) ce = cudaEventRecord(h->evBands, ctx->stream);
                    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(h->xstream, h->evBands, 0);
                    if (rc == SPHGPU_OK && ce == cudaSuccess) {
                        cudaStream_t mainStream = ctx->stream;
                        ctx->stream = h->xstream;
                        cudaEventRecord(ctx->ev[6], ctx->stream);
                        rc = exchange(ctx, false);
                        cudaEventRecord(ctx->ev[7], ctx->stream);
                        ce = cudaEventRecord(h->evArrived, ctx->stream);
                        ctx->stream = mainStream;
                    }
                    if (rc == SPHGPU_OK) rc = launchCorrectPredictRange(ctx, h->sendLeft, n - h->sendRight);
                    if (rc == SPHGPU_OK) rc = exchangeGuard(ctx, h);
                    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->stream, h->evArrived, 0);
                    exchanged = true;
                } else {
                    rc = launchCorrectPredict(ctx);
                }
            } else {
                ctx->d.dtDev = &ctx->d.stepState->dtPrev; // k_finish_timestep has already moved on to the next step
                rc = launchCorrect(ctx, 0.);
                ctx->d.dtDev = &ctx->d.stepState->dt;
            }
        }
    }
    ctx->d.dtDev = nullptr;
    if (ce != cudaSuccess) {
        setError(std::string("CUDA: ") + cudaGetErrorString(ce));
        rc = SPHGPU_E_CUDA;
    }
    if (rc == SPHGPU_OK) {
        ce = cudaEventRecord(ctx->ev[5], ctx->stream);
        rc = collectStats(ctx, stats, ctx->ev[4], ctx->ev[5]); // the one host synchronisation
    } else {
        cudaStreamSynchronize(ctx->stream);
    }
    if (rc == SPHGPU_OK) {
        std::vector<StepRecordDev> hist(steps);
        StepStateDev fin;
        ce = cudaMemcpy(hist.data(), histDev, sizeof(StepRecordDev) * steps, cudaMemcpyDeviceToHost);
        if (ce == cudaSuccess) ce = cudaMemcpy(&fin, ctx->d.stepState, sizeof(fin), cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) {
            setError(std::string("CUDA: ") + cudaGetErrorString(ce));
            rc = SPHGPU_E_CUDA;
        } else {
            for (uint32_t s = 0; s < steps; ++s) {
                history[s].dt = hist[s].dt;
                history[s].criterion = hist[s].criterion;
                history[s].reserved0 = 0;
            }
            ctx->lastDt = fin.lastDt;
            ctx->lastDtInit = fin.lastDtInit != 0u;
            if (h) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
                ctx->lastHaloMs = ms;
            }
        }
    }
    cudaFree(histDev);
    return rc;
}

int sphgpu_step_pc_mgpu(sphgpu_ctx* ctx, double t, double dt, double max_dt, sphgpu_stats* stats, sphgpu_timestep* out) {
    (void)t;
    if (!ctx || !ctx->halo) {
        setError("sphgpu_comm_init / sphgpu_halo_configure must be called first");
        return SPHGPU_E_STATE;
    }
    NcclApi* api = ncclApi();
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[4], ctx->stream));
    int rc;
    if ((rc = launchPredict(ctx, dt)) != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[6], ctx->stream));
    if ((rc = exchange(ctx)) != SPHGPU_OK) return rc; // ghosts carry the PREDICTED state
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[7], ctx->stream));
    if ((rc = enqueueIntegrate(ctx)) != SPHGPU_OK) return rc;
    if ((rc = launchCorrect(ctx, dt)) != SPHGPU_OK) return rc;
    if ((rc = launchCriteria(ctx)) != SPHGPU_OK) return rc;
    // global time step: the bit patterns of positive doubles order like the values, so min over ranks is a u64 min
    if ((rc = allReduceTimestep(ctx, h, api)) != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[5], ctx->stream));
    if ((rc = collectStats(ctx, stats, ctx->ev[4], ctx->ev[5])) != SPHGPU_OK) return rc;
    float ms = 0.f;
    SPH_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
    ctx->lastHaloMs = ms;
    return finishTimestep(ctx, max_dt, out);
}

} // extern "C"
