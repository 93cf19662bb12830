// Multi-GPU step inside the library: ghost (halo) exchange and the global time-step reduction with NCCL, queued on the
// context's stream between the kernels of one PredictorCorrector step, so a step costs ONE host synchronisation.
//
// The reference has no distributed path (core/thread/Scheduler.h:27 rules out MPI); this is the B200-native plumbing
// of SURVEY 8(e): slab neighbours exchange the dynamic neighbour inputs of their boundary bands with grouped
// ncclSend/ncclRecv over NVLink after the prediction and before the derivative evaluation; ncclAllReduce(min) combines
// the per-criterion time-step minima. NCCL is resolved at run time (dlopen of libnccl.so.2 -- inside a PyTorch process
// that is PyTorch's own copy), so libsphgpu has no link-time dependency on it.
#include "sphgpu_internal.h"
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

struct HaloState {
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
    const double hmax = d.grid->hmax * (1. + skin);
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

static int exchange(sphgpu_ctx* ctx) {
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    NcclApi* api = ncclApi();
    const uint32_t n = ctx->n;
    if (h->sendLeft + h->sendRight > n || (uint64_t)n + h->recvLeft + h->recvRight > ctx->capacity) {
        setError("the halo configuration does not fit the current particle count: call sphgpu_halo_configure again");
        return SPHGPU_E_STATE;
    }
    int rc;
    if ((rc = launchHalo(ctx, true, 0, h->sendLeft, h->bufSendL)) != SPHGPU_OK) return rc;
    if ((rc = launchHalo(ctx, true, n - h->sendRight, h->sendRight, h->bufSendR)) != SPHGPU_OK) return rc;
    if (h->guardAxis >= 0 && n > h->sendLeft + h->sendRight) {
        const int inf = 0x7f7fffff; // FLT_MAX
        SPH_CUDA_CHECK(cudaMemcpyAsync(&ctx->d.listCtl->haloMarginBits, &inf, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        k_halo_guard<<<296, 256, 0, ctx->stream>>>(ctx->d, h->sendLeft, n - h->sendRight, h->guardAxis, h->guardLo, h->guardHi, h->guardHasLo,
            h->guardHasHi, ctx->prm.kernel_radius, ctx->listSkin > 0. ? ctx->listSkin : 0.);
        ctx->launches += 1;
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

void invalidateHalo(sphgpu_ctx* ctx) {
    HaloState* h = static_cast<HaloState*>(ctx->halo);
    if (h) {
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
    ctx->listsDirty = true; // another set of ghosts
    SPH_CUDA_CHECK(cudaMemsetAsync(&ctx->d.listCtl->haloViolation, 0, sizeof(uint32_t), ctx->stream));
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
    const StepStateDev init = { dt, ctx->lastDt, ctx->lastDtInit ? 1u : 0u, 0u };
    int rc = SPHGPU_OK;
    cudaError_t ce = cudaMemcpyAsync(ctx->d.stepState, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream);
    ctx->d.dtDev = &ctx->d.stepState->dt;
    ctx->launches = 0;
    for (uint32_t s = 0; s < steps && rc == SPHGPU_OK && ce == cudaSuccess; ++s) {
        if (s + 1 == steps) { // the events time the last step
            ce = cudaEventRecord(ctx->ev[4], ctx->stream);
            ctx->launches = 0;
        }
        rc = launchPredict(ctx, 0.);
        if (rc == SPHGPU_OK && h) {
            cudaEventRecord(ctx->ev[6], ctx->stream);
            rc = exchange(ctx); // ghosts carry the PREDICTED state
            cudaEventRecord(ctx->ev[7], ctx->stream);
        }
        if (rc == SPHGPU_OK) rc = enqueueIntegrate(ctx);
        if (rc == SPHGPU_OK) rc = launchCorrect(ctx, 0.);
        if (rc == SPHGPU_OK) rc = launchCriteria(ctx);
        if (rc == SPHGPU_OK && h) {
            if (api->AllReduce(ctx->d.tsd, ctx->d.tsd, 4, ncclUint64, ncclMin, h->comm, ctx->stream) != ncclSuccess) {
                setError("ncclAllReduce failed");
                rc = SPHGPU_E_CUDA;
            }
        }
        if (rc == SPHGPU_OK) rc = launchFinishTimestep(ctx, max_dt, histDev, s);
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
    SPH_NCCL_CHECK(api->AllReduce(ctx->d.tsd, ctx->d.tsd, 4, ncclUint64, ncclMin, h->comm, ctx->stream));
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[5], ctx->stream));
    if ((rc = collectStats(ctx, stats, ctx->ev[4], ctx->ev[5])) != SPHGPU_OK) return rc;
    float ms = 0.f;
    SPH_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
    ctx->lastHaloMs = ms;
    return finishTimestep(ctx, max_dt, out);
}

} // extern "C"
