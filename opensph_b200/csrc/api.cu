// C-ABI entry points of libsphgpu (see include/sphgpu.h for the reference interface each one replaces).
#include "sphgpu_internal.h"
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

namespace sph {

static thread_local std::string g_lastError;

void setError(const std::string& msg) {
    g_lastError = msg;
}


static int fail(int code, const std::string& msg) {
    setError(msg);
    return code;
}

template <typename T>
static int devAlloc(T** ptr, size_t count) {
    SPH_CUDA_CHECK(cudaMalloc((void**)ptr, sizeof(T) * std::max<size_t>(count, 1)));
    SPH_CUDA_CHECK(cudaMemset(*ptr, 0, sizeof(T) * std::max<size_t>(count, 1)));
    return SPHGPU_OK;
}

static int validate(const sphgpu_config* cfg, const sphgpu_material* mats, uint32_t nmat, uint32_t n, uint32_t cap) {
    if (!cfg || !mats) {
        return fail(SPHGPU_E_INVALID, "null config or materials");
    }
    if (cfg->abi_version != SPHGPU_ABI_VERSION) {
        return fail(SPHGPU_E_INVALID, "ABI version mismatch");
    }
    if (nmat == 0 || nmat > (uint32_t)MAX_MATERIALS) {
        return fail(SPHGPU_E_INVALID, "material count must be in [1, 32]");
    }
    if (cap < n) {
        return fail(SPHGPU_E_INVALID, "capacity smaller than particle count");
    }
    if (!cfg->lut_grad || cfg->lut_entries < 2 || !(cfg->kernel_radius > 0.)) {
        return fail(SPHGPU_E_INVALID, "kernel LUT missing");
    }
    // Terms without a GPU implementation are rejected; there is no CPU fallback (hard parts).
    if (cfg->discretization != SPHGPU_DISCR_STANDARD) {
        return fail(SPHGPU_E_INVALID, "only DiscretizationEnum::STANDARD is implemented on the GPU path");
    }
    if (cfg->flags & SPHGPU_FLAG_XSPH) {
        if (!cfg->lut_value) {
            return fail(SPHGPU_E_INVALID, "the XSph term needs the kernel value table (lut_value)");
        }
        if (cfg->flags & SPHGPU_FLAG_BALSARA) {
            return fail(SPHGPU_E_INVALID, "the XSph term together with the Balsara switch is not implemented on the GPU path");
        }
    }
    if (cfg->flags & SPHGPU_FLAG_STRESS_AV) {
        if (!(cfg->forces & SPHGPU_FORCE_SOLID_STRESS)) {
            return fail(SPHGPU_E_INVALID, "the artificial stress needs ForceEnum::SOLID_STRESS (it is built from the deviatoric stress, Stress.cpp:91-109)");
        }
        if (!cfg->lut_value) {
            return fail(SPHGPU_E_INVALID, "the artificial stress needs the kernel value table (lut_value)");
        }
        if (cfg->flags & (SPHGPU_FLAG_BALSARA | SPHGPU_FLAG_XSPH | SPHGPU_FLAG_DELTASPH)) {
            return fail(SPHGPU_E_INVALID, "the artificial stress together with the Balsara switch, the XSph term or the delta-SPH terms is not implemented on the GPU path");
        }
    }
    if (cfg->flags & SPHGPU_FLAG_DELTASPH) {
        if (cfg->flags & (SPHGPU_FLAG_BALSARA | SPHGPU_FLAG_XSPH)) {
            return fail(SPHGPU_E_INVALID, "the delta-SPH terms together with the Balsara switch or the XSph term are not implemented on the GPU path");
        }
        if ((cfg->flags & SPHGPU_FLAG_CORRECTION_TENSOR) && !(cfg->forces & SPHGPU_FORCE_SOLID_STRESS)) {
            return fail(SPHGPU_E_INVALID, "the delta-SPH terms with the correction tensor need ForceEnum::SOLID_STRESS on the GPU path");
        }
    }
    if (!(cfg->forces & SPHGPU_FORCE_PRESSURE)) {
        return fail(SPHGPU_E_INVALID, "ForceEnum::PRESSURE is required (SolidStressForce is only added with it, StandardSets.cpp:24-33)");
    }
    uint32_t expect = 0;
    for (uint32_t m = 0; m < nmat; ++m) {
        if (mats[m].begin != expect || mats[m].end < mats[m].begin) {
            return fail(SPHGPU_E_INVALID, "materials must own contiguous, ascending particle ranges");
        }
        expect = mats[m].end;
        if (mats[m].eos != SPHGPU_EOS_TILLOTSON && mats[m].eos != SPHGPU_EOS_IDEAL_GAS) {
            return fail(SPHGPU_E_INVALID, "unsupported equation of state (Tillotson and ideal gas are implemented)");
        }
        if (mats[m].yielding != SPHGPU_YIELD_NONE && mats[m].yielding != SPHGPU_YIELD_ELASTIC &&
            mats[m].yielding != SPHGPU_YIELD_VON_MISES) {
            return fail(SPHGPU_E_INVALID, "unsupported rheology (none, elastic and von Mises are implemented)");
        }
        if (mats[m].fracture != SPHGPU_FRACTURE_NONE && mats[m].fracture != SPHGPU_FRACTURE_SCALAR_GRADY_KIPP) {
            return fail(SPHGPU_E_INVALID, "unsupported fracture model (none and scalar Grady-Kipp are implemented)");
        }
    }
    if (expect != n) {
        return fail(SPHGPU_E_INVALID, "material ranges do not cover [0, n_particles)");
    }
    return SPHGPU_OK;
}

} // namespace sph

using namespace sph;

extern "C" {

uint32_t sphgpu_abi_version(void) {
    return SPHGPU_ABI_VERSION;
}

const char* sphgpu_last_error(void) {
    return g_lastError.c_str();
}

int sphgpu_create(const sphgpu_config* cfg, const sphgpu_material* materials, uint32_t n_materials, uint32_t n_particles,
    uint32_t capacity, int device, sphgpu_ctx** out) {
    if (!out) {
        return fail(SPHGPU_E_INVALID, "null output pointer");
    }
    *out = nullptr;
    if (capacity < n_particles) {
        capacity = n_particles;
    }
    int rc = validate(cfg, materials, n_materials, n_particles, capacity);
    if (rc != SPHGPU_OK) {
        return rc;
    }
    if (capacity >= (1u << 30)) {
        return fail(SPHGPU_E_INVALID, "capacity must be below 2^30 particles per device");
    }
    int devCount = 0;
    if (cudaGetDeviceCount(&devCount) != cudaSuccess || devCount == 0) {
        cudaGetLastError();
        return fail(SPHGPU_E_NO_DEVICE, "no CUDA device available; libsphgpu has no CPU path");
    }
    if (device < 0 || device >= devCount) {
        return fail(SPHGPU_E_NO_DEVICE, "device index out of range");
    }
    SPH_CUDA_CHECK(cudaSetDevice(device));
    sphgpu_ctx* ctx = new (std::nothrow) sphgpu_ctx();
    if (!ctx) {
        return fail(SPHGPU_E_OOM, "host allocation failed");
    }
    ctx->device = device;
    ctx->n = n_particles;
    ctx->capacity = capacity;
    ctx->nActive = n_particles;
    ctx->nMaterials = n_materials;

    ParamsDev& p = ctx->prm;
    p.forces = cfg->forces;
    p.flags = cfg->flags;
    p.continuity_mode = cfg->continuity_mode;
    p.lut_entries = cfg->lut_entries;
    p.kernel_radius = cfg->kernel_radius;
    p.radius_sqr = cfg->kernel_radius * cfg->kernel_radius;
    p.q_sqr_to_idx = (double)cfg->lut_entries * (1. / (cfg->kernel_radius * cfg->kernel_radius)); // Kernel.h:88-90
    p.av_alpha = cfg->av_alpha;
    p.av_beta = cfg->av_beta;
    p.av_minus_half_alpha = -0.5 * cfg->av_alpha;
    p.av_eps_over_radius_sqr = 1.e-2 / (cfg->kernel_radius * cfg->kernel_radius);
    p.h_min = cfg->h_min;
    p.h_max = cfg->h_max;
    p.neigh_enforcing = cfg->neigh_enforcing;
    p.neigh_lower = cfg->neigh_lower;
    p.neigh_upper = cfg->neigh_upper;
    p.criteria = cfg->criteria;
    p.n_materials = n_materials;
    p.courant = cfg->courant;
    p.derivative_factor = cfg->derivative_factor;
    p.divergence_factor = cfg->divergence_factor;
    ctx->maxChange = cfg->max_change;

    ctx->solid = (cfg->forces & SPHGPU_FORCE_SOLID_STRESS) != 0;
    ctx->corrected = ctx->solid && (cfg->flags & SPHGPU_FLAG_CORRECTION_TENSOR);
    ctx->balsara = (cfg->flags & SPHGPU_FLAG_BALSARA) != 0;
    ctx->xsph = (cfg->flags & SPHGPU_FLAG_XSPH) != 0;
    p.xsph_eps = 1.; // SPH_XSPH_EPSILON default (core/system/Settings.cpp); sphgpu_set_xsph_epsilon
    ctx->deltasph = (cfg->flags & SPHGPU_FLAG_DELTASPH) != 0;
    p.deltasph_half_delta = 0.5 * 0.01; // SPH_DENSITY_DIFFUSION_DELTA, SPH_VELOCITY_DIFFUSION_ALPHA defaults
    p.deltasph_half_alpha = 0.5 * 0.01; // (core/system/Settings.cpp:541-544); sphgpu_set_deltasph
    ctx->stressAv = (cfg->flags & SPHGPU_FLAG_STRESS_AV) != 0;
    p.stress_av_exponent = 4.; // SPH_AV_STRESS_EXPONENT, SPH_AV_STRESS_FACTOR defaults (core/system/Settings.cpp:579-582);
    p.stress_av_factor = 0.04; // sphgpu_set_stress_av
    p.stress_av_int_exponent = stressAvIntExponent(p.stress_av_exponent);
    ctx->recDoubles = recordDoubles(ctx->solid, ctx->balsara, ctx->deltasph, ctx->stressAv);
    ctx->hasReduce = false;
    ctx->hasDamage = false;
    std::memset(ctx->matsHost, 0, sizeof(ctx->matsHost));
    for (uint32_t m = 0; m < n_materials; ++m) {
        const sphgpu_material& a = materials[m];
        ctx->matsApi[m] = a;
        MaterialDev& b = ctx->matsHost[m];
        b.eos = a.eos;
        b.yielding = a.yielding;
        b.fracture = (a.yielding == SPHGPU_YIELD_VON_MISES) ? a.fracture : (uint32_t)SPHGPU_FRACTURE_NONE;
        b.til_u0 = a.til_u0; b.til_uiv = a.til_uiv; b.til_ucv = a.til_ucv; b.til_a = a.til_a; b.til_b = a.til_b;
        b.rho0 = a.rho0; b.til_A = a.til_A; b.til_B = a.til_B; b.til_alpha = a.til_alpha; b.til_beta = a.til_beta;
        b.gamma = a.gamma;
        b.shear_modulus = a.shear_modulus; b.elasticity_limit = a.elasticity_limit; b.melt_energy = a.melt_energy;
        b.young_modulus = a.young_modulus;
        b.rho_min = a.rho_min; b.rho_max = a.rho_max; b.u_min = a.u_min; b.u_max = a.u_max; b.d_min = a.d_min; b.d_max = a.d_max;
        b.rho_small = a.rho_small; b.u_small = a.u_small; b.d_small = a.d_small; b.s_small = a.s_small;
        // STRESS_REDUCING exists when the material has a rheology (VonMises/Elastic ::create, Rheology.cpp:31,224)
        ctx->hasReduce |= (a.yielding == SPHGPU_YIELD_VON_MISES || a.yielding == SPHGPU_YIELD_ELASTIC);
        ctx->hasDamage |= (b.fracture != SPHGPU_FRACTURE_NONE);
    }
    ctx->filter = ctx->solid && (cfg->flags & SPHGPU_FLAG_SUM_ONLY_UNDAMAGED) && ctx->hasReduce;
    if (ctx->solid && !ctx->hasReduce && cfg->continuity_mode == SPHGPU_CONTINUITY_SUM_ONLY_UNDAMAGED) {
        delete ctx;
        return fail(SPHGPU_E_INVALID, "ContinuityEnum::SUM_ONLY_UNDAMAGED needs STRESS_REDUCING (a rheology)");
    }

    const size_t cap = capacity;
    ctx->maxCells = std::max<uint32_t>(capacity / 4, 4096u);
    ctx->scanBlocks = (ctx->maxCells + 1 + SCAN_ITEMS - 1) / SCAN_ITEMS;
#define SPH_TRY(expr)                                                                                                 \
    do {                                                                                                              \
        rc = (expr);                                                                                                  \
        if (rc != SPHGPU_OK) {                                                                                        \
            sphgpu_destroy(ctx);                                                                                      \
            return rc;                                                                                                \
        }                                                                                                             \
    } while (0)
    auto wrap = [&](cudaError_t e, const char* what) -> int {
        if (e != cudaSuccess) {
            setError(std::string(what) + ": " + cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? SPHGPU_E_OOM : SPHGPU_E_CUDA;
        }
        return SPHGPU_OK;
    };
    SPH_TRY(wrap(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "cudaStreamCreate"));
    ctx->privateStream = ctx->stream;
    for (int k = 0; k < 8; ++k) {
        SPH_TRY(wrap(cudaEventCreate(&ctx->ev[k]), "cudaEventCreate"));
    }
    for (int k = 0; k < 4; ++k) {
        SPH_TRY(wrap(cudaEventCreate(&ctx->evPair[k]), "cudaEventCreate"));
    }
    for (int f = 0; f < F_COUNT; ++f) {
        SPH_TRY(devAlloc(&ctx->d.f[f], cap));
    }
    for (int u = 0; u < U_COUNT; ++u) {
        SPH_TRY(devAlloc(&ctx->d.u[u], cap));
    }
    SPH_TRY(devAlloc(&ctx->d.rec, cap * (size_t)ctx->recDoubles));
    ctx->maxSegs = capacity / 16 + ctx->maxCells + ctx->maxCells / 16 + 2; // units <= N / tile + double rows + x-range cuts
    SPH_TRY(devAlloc(&ctx->d.segStart, (size_t)ctx->maxCells + 2));
    SPH_TRY(devAlloc(&ctx->d.unitDesc, (size_t)ctx->maxSegs));
    SPH_TRY(devAlloc(&ctx->d.unitAux, (size_t)ctx->maxSegs));
    SPH_TRY(devAlloc(&ctx->d.unitLane, cap));
    SPH_TRY(devAlloc(&ctx->d.unitList, (size_t)ctx->maxSegs * 4));
    // list pool: ~0.8 rows of 256 B per particle on the eta = 1.3 lattice (68 neighbours); units that do not fit fall
    // back to building their lists inside the pair kernel
    ctx->poolRows = (uint32_t)std::min<size_t>(cap + cap / 2 + 8192, 0xfffffff0u);
    SPH_TRY(devAlloc(&ctx->d.listPool, (size_t)ctx->poolRows * 256));
    SPH_TRY(devAlloc(&ctx->d.listCursor, 1));
    SPH_TRY(devAlloc(&ctx->d.stepState, 1));
    ctx->d.dtDev = nullptr;
    SPH_TRY(devAlloc(&ctx->d.sCell, cap));
    SPH_TRY(devAlloc(&ctx->d.posF, cap));
    SPH_TRY(devAlloc(&ctx->d.pos0, cap));
    SPH_TRY(devAlloc(&ctx->d.accLarge, cap));
    SPH_TRY(devAlloc(&ctx->d.largePartial, (size_t)LARGE_MAX));
    SPH_TRY(devAlloc(&ctx->d.largeCounter, (size_t)LARGE_MAX));
    SPH_TRY(devAlloc(&ctx->d.listCtl, 1));
    SPH_TRY(devAlloc(&ctx->d.cellHmax, (size_t)ctx->maxCells + 1));
    SPH_TRY(devAlloc(&ctx->d.order, cap));
    SPH_TRY(devAlloc(&ctx->d.cellOf, cap));
    SPH_TRY(devAlloc(&ctx->d.dispA, 6 * ((size_t)ctx->maxCells + 1)));
    SPH_TRY(devAlloc(&ctx->d.dispGlobal, 8));
    SPH_TRY(devAlloc(&ctx->d.rank, cap));
    SPH_TRY(devAlloc(&ctx->d.cellStart, (size_t)ctx->maxCells + 2));
    SPH_TRY(devAlloc(&ctx->d.cellCount, (size_t)ctx->maxCells + 2));
    SPH_TRY(devAlloc(&ctx->d.scanBlock, (size_t)ctx->scanBlocks + 1));
    SPH_TRY(devAlloc(&ctx->d.boundsPartial, (size_t)BOUNDS_BLOCKS * BOUNDS_STRIDE));
    SPH_TRY(devAlloc(&ctx->d.grid, 1));
    SPH_TRY(devAlloc(&ctx->d.stats, 1));
    SPH_TRY(devAlloc(&ctx->d.statsInit, 1));
    {
        const StatsDev init = { 0xffffffffu, 0u, 0ull, 0u, 0u };
        SPH_TRY(wrap(cudaMemcpy(ctx->d.statsInit, &init, sizeof(init), cudaMemcpyHostToDevice), "init"));
    }
    SPH_TRY(devAlloc(&ctx->d.tsd, 1));
    double* lut = nullptr;
    SPH_TRY(devAlloc(&lut, (size_t)cfg->lut_entries + 2));
    ctx->d.lut = lut;
    SPH_TRY(wrap(cudaMemcpy(lut, cfg->lut_grad, sizeof(double) * ((size_t)cfg->lut_entries + 1), cudaMemcpyHostToDevice), "LUT upload"));
    {
        std::vector<LutPair> pairs((size_t)cfg->lut_entries + 1);
        buildLutPairs(cfg->lut_grad, cfg->lut_entries, pairs.data());
        LutPair* lut2 = nullptr;
        SPH_TRY(devAlloc(&lut2, pairs.size()));
        ctx->d.lut2 = lut2;
        SPH_TRY(wrap(cudaMemcpy(lut2, pairs.data(), sizeof(LutPair) * pairs.size(), cudaMemcpyHostToDevice), "LUT upload"));
    }
    if (ctx->xsph || ctx->stressAv) { // kernel values for the XSph term and the artificial stress: LutKernel::valueImpl interpolates the same way (Kernel.h:111-127)
        double* lutW = nullptr;
        SPH_TRY(devAlloc(&lutW, (size_t)cfg->lut_entries + 2));
        ctx->d.lutW = lutW;
        SPH_TRY(wrap(cudaMemcpy(lutW, cfg->lut_value, sizeof(double) * ((size_t)cfg->lut_entries + 1), cudaMemcpyHostToDevice), "LUT upload"));
        std::vector<LutPair> pairs((size_t)cfg->lut_entries + 1);
        buildLutPairs(cfg->lut_value, cfg->lut_entries, pairs.data());
        LutPair* lutW2 = nullptr;
        SPH_TRY(devAlloc(&lutW2, pairs.size()));
        ctx->d.lutW2 = lutW2;
        SPH_TRY(wrap(cudaMemcpy(lutW2, pairs.data(), sizeof(LutPair) * pairs.size(), cudaMemcpyHostToDevice), "LUT upload"));
    }
    SPH_TRY(wrap(cudaMalloc(&ctx->staging, std::max<size_t>(cap, 1) * 64), "staging"));
    // identity correction tensor and reduce = 1 by default (SolidStressForce::create, EquationTerm.cpp:215-218)
    {
        std::vector<double> ones(std::max<size_t>(cap, 1), 1.);
        for (int f : { (int)F_C0, (int)F_C1, (int)F_C2, (int)F_REDUCE }) {
            SPH_TRY(wrap(cudaMemcpy(ctx->d.f[f], ones.data(), sizeof(double) * cap, cudaMemcpyHostToDevice), "init"));
        }
        std::vector<uint32_t> matid(std::max<size_t>(cap, 1), 0u);
        for (uint32_t m = 0; m < n_materials; ++m) {
            for (uint32_t i = materials[m].begin; i < materials[m].end; ++i) {
                matid[i] = m;
            }
        }
        SPH_TRY(wrap(cudaMemcpy(ctx->d.u[U_MATID], matid.data(), sizeof(uint32_t) * cap, cudaMemcpyHostToDevice), "init"));
    }
    SPH_TRY(ensureConstants(ctx));
#undef SPH_TRY
    *out = ctx;
    return SPHGPU_OK;
}

int sphgpu_destroy(sphgpu_ctx* ctx) {
    if (!ctx) {
        return SPHGPU_OK;
    }
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    forgetConstants(ctx);
    destroyHalo(ctx);
    destroyGravity(ctx);
    destroySymmetric(ctx);
    for (int f = 0; f < F_COUNT; ++f) cudaFree(ctx->d.f[f]);
    for (int u = 0; u < U_COUNT; ++u) cudaFree(ctx->d.u[u]);
    cudaFree(ctx->d.rec); cudaFree(ctx->d.segStart); cudaFree(ctx->d.unitDesc); cudaFree(ctx->d.unitAux); cudaFree(ctx->d.unitLane); cudaFree(ctx->d.unitList);
    cudaFree(ctx->d.listPool); cudaFree(ctx->d.listCursor); cudaFree(ctx->d.stepState);
    cudaFree(ctx->d.posF); cudaFree(ctx->d.pos0); cudaFree(ctx->d.accLarge); cudaFree(ctx->d.largePartial); cudaFree(ctx->d.largeCounter); cudaFree(ctx->d.listCtl); cudaFree(ctx->d.cellHmax);
    cudaFree(ctx->d.sCell); cudaFree(ctx->d.order); cudaFree(ctx->d.cellOf); cudaFree(ctx->d.rank);
    cudaFree(ctx->d.dispA); cudaFree(ctx->d.dispGlobal);
    cudaFree(ctx->d.cellStart); cudaFree(ctx->d.cellCount); cudaFree(ctx->d.scanBlock); cudaFree(ctx->d.boundsPartial);
    cudaFree(ctx->d.grid); cudaFree(ctx->d.stats); cudaFree(ctx->d.statsInit); cudaFree(ctx->d.tsd); cudaFree((void*)ctx->d.lut); cudaFree((void*)ctx->d.lut2); cudaFree((void*)ctx->d.lutW); cudaFree((void*)ctx->d.lutW2); cudaFree(ctx->staging);
    for (int k = 0; k < 8; ++k) {
        if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
        if (k < 4 && ctx->evPair[k]) cudaEventDestroy(ctx->evPair[k]);
    }
    if (ctx->privateStream) cudaStreamDestroy(ctx->privateStream);
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    if (ctx->evPacked) cudaEventDestroy(ctx->evPacked);
    if (ctx->evCopied) cudaEventDestroy(ctx->evCopied);
    cudaFree(ctx->stagingDown);
    cudaGetLastError();
    delete ctx;
    return SPHGPU_OK;
}

static int checkRange(sphgpu_ctx* ctx, uint32_t first, uint32_t count, const void* ptr) {
    if (!ctx) {
        return fail(SPHGPU_E_INVALID, "null context");
    }
    if ((uint64_t)first + count > ctx->capacity) {
        return fail(SPHGPU_E_INVALID, "particle range exceeds the context capacity");
    }
    if (count > 0 && !ptr) {
        return fail(SPHGPU_E_INVALID, "null buffer");
    }
    return SPHGPU_OK;
}

int sphgpu_upload(sphgpu_ctx* ctx, int q, int order, int layout, const void* host, uint32_t first, uint32_t count) {
    int rc = checkRange(ctx, first, count, host);
    if (rc != SPHGPU_OK) return rc;
    const size_t eb = elementBytes(q, layout);
    if (eb == 0) return fail(SPHGPU_E_INVALID, "unknown quantity id");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaMemcpyAsync(ctx->staging, host, eb * count, cudaMemcpyHostToDevice, ctx->stream));
    rc = launchUnpack(ctx, q, order, layout, ctx->staging, first, count);
    if (rc != SPHGPU_OK) return rc;
    // the staging buffer is reused by the next call
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->stateUploaded = true;
    return SPHGPU_OK;
}

int sphgpu_download(sphgpu_ctx* ctx, int q, int order, int layout, void* host, uint32_t first, uint32_t count) {
    int rc = checkRange(ctx, first, count, host);
    if (rc != SPHGPU_OK) return rc;
    const size_t eb = elementBytes(q, layout);
    if (eb == 0) return fail(SPHGPU_E_INVALID, "unknown quantity id");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    rc = launchPack(ctx, q, order, layout, ctx->staging, first, count);
    if (rc != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaMemcpyAsync(host, ctx->staging, eb * count, cudaMemcpyDeviceToHost, ctx->stream));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return SPHGPU_OK;
}

int sphgpu_upload_async(sphgpu_ctx* ctx, int q, int order, int layout, const void* host, uint32_t first, uint32_t count) {
    int rc = checkRange(ctx, first, count, host);
    if (rc != SPHGPU_OK) return rc;
    const size_t eb = elementBytes(q, layout);
    if (eb == 0) return fail(SPHGPU_E_INVALID, "unknown quantity id");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    // stream order protects the staging buffer: the unpack kernel of the previous call runs before this copy
    SPH_CUDA_CHECK(cudaMemcpyAsync(ctx->staging, host, eb * count, cudaMemcpyHostToDevice, ctx->stream));
    rc = launchUnpack(ctx, q, order, layout, ctx->staging, first, count);
    if (rc != SPHGPU_OK) return rc;
    ctx->stateUploaded = true;
    return SPHGPU_OK;
}

int sphgpu_download_async(sphgpu_ctx* ctx, int q, int order, int layout, void* host, uint32_t first, uint32_t count) {
    int rc = checkRange(ctx, first, count, host);
    if (rc != SPHGPU_OK) return rc;
    const size_t eb = elementBytes(q, layout);
    if (eb == 0) return fail(SPHGPU_E_INVALID, "unknown quantity id");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!ctx->copyStream) {
        SPH_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        SPH_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->evPacked, cudaEventDisableTiming));
        SPH_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->evCopied, cudaEventDisableTiming));
    }
    const size_t bytes = (eb * count + 255) & ~(size_t)255;
    if (ctx->downOffset + bytes > ctx->stagingDownBytes) { // (first use, or a batch larger than any before)
        SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->copyStream));
        if (ctx->downOffset > 0) {
            return fail(SPHGPU_E_STATE, "download batch larger than the staging buffer: call sphgpu_download_batch_end between batches");
        }
        cudaFree(ctx->stagingDown);
        ctx->stagingDown = nullptr;
        ctx->stagingDownBytes = std::max<size_t>((size_t)ctx->capacity * 256 + 4096, bytes * 4);
        SPH_CUDA_CHECK(cudaMalloc(&ctx->stagingDown, ctx->stagingDownBytes));
    }
    if (ctx->downOffset == 0 && ctx->copiesPending) {
        // a new batch overwrites the staging area: the copies of the previous batch must have left it
        SPH_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->evCopied, 0));
    }
    void* slot = static_cast<char*>(ctx->stagingDown) + ctx->downOffset;
    rc = launchPack(ctx, q, order, layout, slot, first, count);
    if (rc != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->evPacked, ctx->stream));
    SPH_CUDA_CHECK(cudaStreamWaitEvent(ctx->copyStream, ctx->evPacked, 0));
    SPH_CUDA_CHECK(cudaMemcpyAsync(host, slot, eb * count, cudaMemcpyDeviceToHost, ctx->copyStream));
    SPH_CUDA_CHECK(cudaEventRecord(ctx->evCopied, ctx->copyStream));
    ctx->copiesPending = true;
    ctx->downOffset += bytes;
    return SPHGPU_OK;
}

int sphgpu_download_batch_end(sphgpu_ctx* ctx) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    ctx->downOffset = 0;
    return SPHGPU_OK;
}

int sphgpu_host_alloc(void** out, size_t bytes) {
    if (!out) return fail(SPHGPU_E_INVALID, "null argument");
    *out = nullptr;
    SPH_CUDA_CHECK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return SPHGPU_OK;
}

int sphgpu_host_free(void* ptr) {
    if (ptr) {
        cudaFreeHost(ptr);
        cudaGetLastError();
    }
    return SPHGPU_OK;
}

int sphgpu_transfer_sync(sphgpu_ctx* ctx) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (ctx->copyStream) {
        SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->copyStream));
    }
    ctx->copiesPending = false;
    ctx->downOffset = 0;
    return SPHGPU_OK;
}

int sphgpu_upload_device(sphgpu_ctx* ctx, int q, int order, const void* dev, uint32_t first, uint32_t count) {
    int rc = checkRange(ctx, first, count, dev);
    if (rc != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    rc = launchUnpack(ctx, q, order, SPHGPU_LAYOUT_PACKED, dev, first, count);
    if (rc == SPHGPU_OK) ctx->stateUploaded = true;
    return rc;
}

int sphgpu_download_device(sphgpu_ctx* ctx, int q, int order, void* dev, uint32_t first, uint32_t count) {
    int rc = checkRange(ctx, first, count, dev);
    if (rc != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    return launchPack(ctx, q, order, SPHGPU_LAYOUT_PACKED, dev, first, count);
}

int sphgpu_halo_pack(sphgpu_ctx* ctx, uint32_t first, uint32_t count, void* dev_buffer) {
    int rc = checkRange(ctx, first, count, dev_buffer);
    if (rc != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    return launchHalo(ctx, true, first, count, dev_buffer);
}

int sphgpu_halo_unpack(sphgpu_ctx* ctx, uint32_t first, uint32_t count, const void* dev_buffer) {
    int rc = checkRange(ctx, first, count, dev_buffer);
    if (rc != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    rc = launchHalo(ctx, false, first, count, const_cast<void*>(dev_buffer));
    if (rc == SPHGPU_OK) ctx->stateUploaded = true;
    return rc;
}

int sphgpu_set_active(sphgpu_ctx* ctx, uint32_t n_active) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (n_active < ctx->n || n_active > ctx->capacity) {
        return fail(SPHGPU_E_INVALID, "active count must be in [n_particles, capacity]");
    }
    if (ctx->nActive != n_active) {
        ctx->listsDirty = true; // the candidate lists were built for another set of particles
    }
    ctx->nActive = n_active;
    return SPHGPU_OK;
}

int sphgpu_set_particle_count(sphgpu_ctx* ctx, uint32_t n_particles) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (n_particles > ctx->capacity) {
        return fail(SPHGPU_E_INVALID, "particle count exceeds the capacity of the context");
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->n = n_particles;
    ctx->nActive = n_particles; // ghosts are dropped; the caller appends them again (sphgpu_set_active)
    ctx->listsDirty = true;
    invalidateHalo(ctx); // the band slot ranges referred to the old particle set
    return SPHGPU_OK;
}

int sphgpu_set_variant(sphgpu_ctx* ctx, int variant) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (variant < 0 || variant > 4) return fail(SPHGPU_E_INVALID, "pair-kernel variant must be 0 .. 4");
    if (variant == 4 && (ctx->corrected || ctx->balsara || ctx->xsph || ctx->deltasph || ctx->stressAv)) {
        return fail(SPHGPU_E_INVALID, "the symmetric formulation (variant 4) offers neither the correction tensor (like SymmetricSolver, "
                                      "SymmetricSolver.cpp:41-44) nor the Balsara switch / XSph");
    }
    if (ctx->variant != variant) {
        ctx->listsDirty = true; // variant 3 builds its lists against a different pool size
    }
    ctx->variant = variant;
    return SPHGPU_OK;
}

int sphgpu_set_list_skin(sphgpu_ctx* ctx, double skin) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (!(skin >= 0.) || skin > 0.5) return fail(SPHGPU_E_INVALID, "list skin must be in [0, 0.5]");
    ctx->listSkin = skin;
    ctx->listsDirty = true;
    return SPHGPU_OK;
}

int sphgpu_list_stats(sphgpu_ctx* ctx, uint32_t* rebuilds, uint32_t* age, double* metric) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (rebuilds) *rebuilds = ctx->listRebuilds;
    if (age) *age = ctx->listAge;
    if (metric) *metric = ctx->listMetric;
    return SPHGPU_OK;
}

int sphgpu_set_stream(sphgpu_ctx* ctx, void* cuda_stream) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)cuda_stream;
    return SPHGPU_OK;
}

int sphgpu_use_private_stream(sphgpu_ctx* ctx) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = ctx->privateStream;
    return SPHGPU_OK;
}

int sphgpu_synchronize(sphgpu_ctx* ctx) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return SPHGPU_OK;
}

} // extern "C"

namespace sph {

// Queues one integrate() on the stream; events 0..3 bracket grid build / prologue / pair kernel.
int enqueueIntegrate(sphgpu_ctx* ctx) {
    int rc;
    if ((rc = ensureConstants(ctx)) != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[0], ctx->stream));
    SPH_CUDA_CHECK(cudaMemcpyAsync(ctx->d.stats, ctx->d.statsInit, sizeof(StatsDev), cudaMemcpyDeviceToDevice, ctx->stream));
    if ((rc = launchGridBuild(ctx)) != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[1], ctx->stream));
    if ((rc = launchProloguePack(ctx)) != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[2], ctx->stream));
    if ((rc = launchPair(ctx)) != SPHGPU_OK) return rc;
    // GravitySolver::loop (GravitySolver.cpp:64-99): the gravitational accelerations join the SPH ones
    if (ctx->gravity != nullptr && (rc = launchGravity(ctx, 1)) != SPHGPU_OK) return rc;
    // afterLoop: boundary conditions see the finished derivatives (AsymmetricSolver.cpp:216)
    if (ctx->hasFrozen && (rc = launchFrozen(ctx)) != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[3], ctx->stream));
    return SPHGPU_OK;
}

int collectStats(sphgpu_ctx* ctx, sphgpu_stats* stats, cudaEvent_t begin, cudaEvent_t end) {
    StatsDev sd;
    ListCtlDev lc;
    SPH_CUDA_CHECK(cudaMemcpyAsync(&sd, ctx->d.stats, sizeof(sd), cudaMemcpyDeviceToHost, ctx->stream));
    SPH_CUDA_CHECK(cudaMemcpyAsync(&lc, ctx->d.listCtl, sizeof(lc), cudaMemcpyDeviceToHost, ctx->stream));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    sd.fallbackUnits = lc.fallbackUnits;
    ctx->listRebuilds = lc.rebuilds;
    ctx->listAge = lc.age;
    ctx->listMetric = lc.lastMetric;
    {
        float m;
        std::memcpy(&m, &lc.haloMarginBits, sizeof(m));
        ctx->haloMargin = m;
    }
    if (lc.peerTimeout != 0u) {
        return fail(SPHGPU_E_STATE, "a peer rank did not arrive within 20 s (peer-memory halo exchange / all-reduce): the ranks are out of step");
    }
    if (lc.haloViolation != 0u) {
        return fail(SPHGPU_E_STATE, "halo band outgrown: an interior particle has come within the kernel reach of a cut plane "
                                    "(neighbours on the other rank would be missed); repartition and configure the halo again");
    }
    float ms = 0.f;
    for (int k = 0; k < 3; ++k) {
        SPH_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[k], ctx->ev[k + 1]));
        ctx->lastMs[k] = ms;
    }
    for (int k = 0; k < 3; ++k) {
        ctx->lastPairMs[k] = 0.;
        if (ctx->pairTimed) {
            SPH_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->evPair[k], ctx->evPair[k + 1]));
            ctx->lastPairMs[k] = ms;
        }
    }
    SPH_CUDA_CHECK(cudaEventElapsedTime(&ms, begin, end));
    ctx->lastMs[3] = ms - (ctx->lastMs[0] + ctx->lastMs[1] + ctx->lastMs[2]);
    // Units whose candidate lists did not fit the list pool were handled by the (slower) fused kernel; give the next
    // steps a larger pool (the stream is idle here). Variant 3 keeps its deliberately tiny pool.
    if (sd.fallbackUnits > 0 && ctx->variant != 3 && ctx->poolRows < 0x7ffffff0u) {
        const uint32_t grown = (uint32_t)std::min<uint64_t>((uint64_t)ctx->poolRows * 2, 0xfffffff0ull);
        unsigned char* bigger = nullptr;
        if (cudaMalloc(&bigger, (size_t)grown * 256) == cudaSuccess) {
            cudaFree(ctx->d.listPool);
            ctx->d.listPool = bigger;
            ctx->poolRows = grown;
            ctx->listsDirty = true; // the lists lived in the old pool
        } else {
            cudaGetLastError(); // out of memory: keep the pool, the fused path stays correct
        }
    }
    if (sd.badFlags > 0) {
        return fail(SPHGPU_E_INVALID, "body flags must be below 254 on the solid GPU path (group field of the neighbour record)");
    }
    if (stats) {
        stats->neigh_min = ctx->n ? sd.neighMin : 0;
        stats->neigh_max = sd.neighMax;
        stats->pair_count = sd.pairCount;
        stats->neigh_mean = ctx->n ? (double)sd.pairCount / (double)ctx->n : 0.;
        stats->gpu_ms = ms;
        stats->kernel_launches = ctx->launches;
        stats->reserved0 = sd.fallbackUnits; // units whose candidate lists did not fit the list pool
    }
    return SPHGPU_OK;
}

} // namespace sph

extern "C" {

int sphgpu_integrate(sphgpu_ctx* ctx, double t, sphgpu_stats* stats) {
    (void)t;
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (!ctx->stateUploaded && ctx->n > 0) return fail(SPHGPU_E_STATE, "integrate called before any state was uploaded");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    int rc = enqueueIntegrate(ctx);
    if (rc != SPHGPU_OK) return rc;
    return collectStats(ctx, stats, ctx->ev[0], ctx->ev[3]);
}

int sphgpu_set_frozen(sphgpu_ctx* ctx, const sphgpu_frozen* cfg) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (cfg && cfg->has_domain && !(cfg->radius > 0. && cfg->freeze_radius >= 0.)) {
        return fail(SPHGPU_E_INVALID, "frozen particles: the domain needs a positive radius and a non-negative freeze radius");
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->hasFrozen = cfg != nullptr && (cfg->flag_mask != 0ull || cfg->has_domain != 0);
    if (cfg) ctx->frozen = *cfg;
    return SPHGPU_OK;
}

int sphgpu_set_stress_av(sphgpu_ctx* ctx, double exponent, double factor) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (!ctx->stressAv) return fail(SPHGPU_E_STATE, "the context was created without SPHGPU_FLAG_STRESS_AV");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->prm.stress_av_exponent = exponent;
    ctx->prm.stress_av_factor = factor;
    ctx->prm.stress_av_int_exponent = stressAvIntExponent(exponent);
    forgetConstants(ctx); // uploaded again by the next call
    return SPHGPU_OK;
}

int sphgpu_set_deltasph(sphgpu_ctx* ctx, double delta, double alpha) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (!ctx->deltasph) return fail(SPHGPU_E_STATE, "the context was created without SPHGPU_FLAG_DELTASPH");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->prm.deltasph_half_delta = 0.5 * delta;
    ctx->prm.deltasph_half_alpha = 0.5 * alpha;
    forgetConstants(ctx); // uploaded again by the next call
    return SPHGPU_OK;
}

int sphgpu_set_xsph_epsilon(sphgpu_ctx* ctx, double epsilon) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (!ctx->xsph) return fail(SPHGPU_E_STATE, "the context was created without SPHGPU_FLAG_XSPH");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ctx->prm.xsph_eps = epsilon;
    forgetConstants(ctx); // uploaded again by the next call
    return SPHGPU_OK;
}

int sphgpu_step_predict(sphgpu_ctx* ctx, double dt) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    return launchPredict(ctx, dt);
}

int sphgpu_step_correct(sphgpu_ctx* ctx, double dt) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    return launchCorrect(ctx, dt);
}

int sphgpu_step_euler(sphgpu_ctx* ctx, double dt) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    return launchEuler(ctx, dt);
}

} // extern "C"

namespace sph {

// MultiCriterion::compute (TimeStepCriterion.cpp:389-419) from the four per-criterion minima.
int finishTimestep(sphgpu_ctx* ctx, double max_dt, sphgpu_timestep* out) {
    TimestepDev td;
    SPH_CUDA_CHECK(cudaMemcpyAsync(&td, ctx->d.tsd, sizeof(td), cudaMemcpyDeviceToHost, ctx->stream));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    static const uint32_t bits[4] = { SPHGPU_CRIT_COURANT, SPHGPU_CRIT_DERIVATIVES, SPHGPU_CRIT_ACCELERATION, SPHGPU_CRIT_DIVERGENCE };
    static const uint32_t ids[4] = { SPHGPU_CRITID_CFL_CONDITION, SPHGPU_CRITID_DERIVATIVE, SPHGPU_CRITID_ACCELERATION,
        SPHGPU_CRITID_DIVERGENCE };
    double minStep = INFTY_REF;
    uint32_t minId = SPHGPU_CRITID_INITIAL_VALUE;
    for (int k = 0; k < 4; ++k) {
        if (!(ctx->prm.criteria & bits[k])) continue;
        double step;
        std::memcpy(&step, &td.minBits[k], 8);
        uint32_t id = ids[k];
        if (step > max_dt) {
            step = max_dt;
            id = SPHGPU_CRITID_MAXIMAL_VALUE;
        }
        if (step < minStep) {
            minStep = step;
            minId = id;
        }
    }
    if (ctx->maxChange < 1.e300) {
        if (!ctx->lastDtInit) {
            ctx->lastDt = minStep;
            ctx->lastDtInit = true;
        }
        const double maxStep = ctx->lastDt * (1. + ctx->maxChange);
        if (minStep > maxStep) {
            minStep = maxStep;
            minId = SPHGPU_CRITID_MAX_CHANGE;
        }
        ctx->lastDt = minStep;
    }
    if (out) {
        out->dt = minStep;
        out->criterion = minId;
        out->reserved0 = 0;
    }
    return SPHGPU_OK;
}

} // namespace sph

extern "C" {

int sphgpu_compute_timestep(sphgpu_ctx* ctx, double max_dt, sphgpu_timestep* out) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc = launchCriteria(ctx);
    if (rc != SPHGPU_OK) return rc;
    return finishTimestep(ctx, max_dt, out);
}

int sphgpu_step_pc(sphgpu_ctx* ctx, double t, double dt, double max_dt, sphgpu_stats* stats, sphgpu_timestep* out) {
    (void)t;
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    if (!ctx->stateUploaded && ctx->n > 0) return fail(SPHGPU_E_STATE, "step called before any state was uploaded");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    ctx->launches = 0;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[4], ctx->stream));
    int rc;
    if ((rc = launchPredict(ctx, dt)) != SPHGPU_OK) return rc;
    if ((rc = enqueueIntegrate(ctx)) != SPHGPU_OK) return rc;
    if ((rc = launchCorrect(ctx, dt)) != SPHGPU_OK) return rc;
    if ((rc = launchCriteria(ctx)) != SPHGPU_OK) return rc;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[5], ctx->stream));
    if ((rc = collectStats(ctx, stats, ctx->ev[4], ctx->ev[5])) != SPHGPU_OK) return rc;
    return finishTimestep(ctx, max_dt, out);
}

int sphgpu_set_last_timestep(sphgpu_ctx* ctx, double dt) {
    if (!ctx) return fail(SPHGPU_E_INVALID, "null context");
    ctx->lastDt = dt;
    ctx->lastDtInit = true;
    return SPHGPU_OK;
}

int sphgpu_last_timings(sphgpu_ctx* ctx, double* ms4) {
    if (!ctx || !ms4) return fail(SPHGPU_E_INVALID, "null argument");
    for (int k = 0; k < 4; ++k) ms4[k] = ctx->lastMs[k];
    return SPHGPU_OK;
}

int sphgpu_last_pair_timings(sphgpu_ctx* ctx, double* ms3) {
    if (!ctx || !ms3) return fail(SPHGPU_E_INVALID, "null argument");
    for (int k = 0; k < 3; ++k) ms3[k] = ctx->lastPairMs[k];
    return SPHGPU_OK;
}

int sphgpu_measure_fp64_peak(sphgpu_ctx* ctx, double* fma_per_second) {
    if (!ctx || !fma_per_second) return fail(SPHGPU_E_INVALID, "null argument");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    return measureFp64Peak(ctx, fma_per_second);
}

int sphgpu_last_halo_ms(sphgpu_ctx* ctx, double* ms) {
    if (!ctx || !ms) return fail(SPHGPU_E_INVALID, "null argument");
    *ms = ctx->lastHaloMs;
    return SPHGPU_OK;
}

int sphgpu_neighbour_dump(sphgpu_ctx* ctx, uint64_t* offsets, uint32_t* idx, uint64_t idx_capacity) {
    if (!ctx || !offsets) return fail(SPHGPU_E_INVALID, "null argument");
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc;
    // uses the cell list and sorted planes of the current positions
    ctx->listsDirty = true;
    if ((rc = launchGridBuild(ctx)) != SPHGPU_OK) return rc;
    if ((rc = launchProloguePackPositionsOnly(ctx)) != SPHGPU_OK) return rc;
    ctx->listsDirty = true; // the work units and candidate lists no longer match the rebuilt cell list
    const uint32_t n = ctx->n;
    uint32_t* countsDev = nullptr;
    SPH_CUDA_CHECK(cudaMalloc(&countsDev, sizeof(uint32_t) * std::max<uint32_t>(n, 1)));
    rc = launchNeighbourCount(ctx, countsDev);
    std::vector<uint32_t> counts(n);
    if (rc == SPHGPU_OK && n > 0) {
        cudaError_t e = cudaMemcpyAsync(counts.data(), countsDev, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            setError(cudaGetErrorString(e));
            rc = SPHGPU_E_CUDA;
        }
    }
    cudaFree(countsDev);
    if (rc != SPHGPU_OK) return rc;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; ++i) {
        offsets[i] = total;
        total += counts[i];
    }
    offsets[n] = total;
    if (!idx) return SPHGPU_OK;
    if (total > idx_capacity) return fail(SPHGPU_E_INVALID, "index buffer too small");
    unsigned long long* offDev = nullptr;
    uint32_t* idxDev = nullptr;
    SPH_CUDA_CHECK(cudaMalloc(&offDev, sizeof(unsigned long long) * ((size_t)n + 1)));
    cudaError_t e = cudaMalloc(&idxDev, sizeof(uint32_t) * std::max<uint64_t>(total, 1));
    if (e != cudaSuccess) {
        cudaFree(offDev);
        return fail(SPHGPU_E_OOM, "neighbour index buffer");
    }
    cudaMemcpyAsync(offDev, offsets, sizeof(unsigned long long) * ((size_t)n + 1), cudaMemcpyHostToDevice, ctx->stream);
    rc = launchNeighbourFill(ctx, offDev, idxDev);
    if (rc == SPHGPU_OK) {
        e = cudaMemcpyAsync(idx, idxDev, sizeof(uint32_t) * total, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            setError(cudaGetErrorString(e));
            rc = SPHGPU_E_CUDA;
        }
    }
    cudaFree(offDev);
    cudaFree(idxDev);
    if (rc != SPHGPU_OK) return rc;
    for (uint32_t i = 0; i < n; ++i) {
        std::sort(idx + offsets[i], idx + offsets[i + 1]);
    }
    return SPHGPU_OK;
}

} // extern "C"
