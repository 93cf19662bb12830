// Host AoS <-> device SoA repack. The host array (packed or in the reference's in-memory layout) is copied to a
// device staging buffer in one cudaMemcpy and (un)packed there, so the PCIe transfer is a single contiguous copy.
// Replaces direct Array<T> access through Storage::getValue/getDt/getD2t (core/quantities/Storage.h:291-608) and
// Accumulated::store moving result buffers into the Storage (core/sph/equations/Accumulated.cpp:87-100).
#include "sphgpu_internal.h"

namespace sph {

struct QuantityMap {
    int ncomp;        // components per particle (0 = invalid combination)
    int field[6];     // Field ids (double planes) or UField ids (u32 planes); -1 = constant zero lane
    bool isU32;
};

static QuantityMap quantityMap(int q, int order) {
    QuantityMap m{ 0, { -1, -1, -1, -1, -1, -1 }, false };
    auto set = [&](int n, std::initializer_list<int> f) {
        m.ncomp = n;
        int k = 0;
        for (int v : f) {
            m.field[k++] = v;
        }
    };
    switch (q) {
    case SPHGPU_Q_POSITION:
        if (order == 0) set(4, { F_X, F_Y, F_Z, F_H });
        if (order == 1) set(4, { F_VX, F_VY, F_VZ, F_VH });
        if (order == 2) set(4, { F_AX, F_AY, F_AZ, -1 });
        break;
    case SPHGPU_Q_MASS: if (order == 0) set(1, { F_M }); break;
    case SPHGPU_Q_DENSITY: if (order == 0) set(1, { F_RHO }); if (order == 1) set(1, { F_DRHO }); break;
    case SPHGPU_Q_ENERGY: if (order == 0) set(1, { F_U }); if (order == 1) set(1, { F_DU }); break;
    case SPHGPU_Q_PRESSURE: if (order == 0) set(1, { F_P }); break;
    case SPHGPU_Q_SOUND_SPEED: if (order == 0) set(1, { F_CS }); break;
    case SPHGPU_Q_DEVIATORIC_STRESS:
        if (order == 0) set(5, { F_S0, F_S1, F_S2, F_S3, F_S4 });
        if (order == 1) set(5, { F_DS0, F_DS1, F_DS2, F_DS3, F_DS4 });
        break;
    case SPHGPU_Q_DAMAGE: if (order == 0) set(1, { F_D }); if (order == 1) set(1, { F_DD }); break;
    case SPHGPU_Q_STRESS_REDUCING: if (order == 0) set(1, { F_REDUCE }); break;
    case SPHGPU_Q_VELOCITY_DIVERGENCE: if (order == 0) set(1, { F_DIVV }); break;
    case SPHGPU_Q_VELOCITY_GRADIENT: if (order == 0) set(6, { F_GV0, F_GV1, F_GV2, F_GV3, F_GV4, F_GV5 }); break;
    case SPHGPU_Q_CORRECTION_TENSOR: if (order == 0) set(6, { F_C0, F_C1, F_C2, F_C3, F_C4, F_C5 }); break;
    case SPHGPU_Q_EPS_MIN: if (order == 0) set(1, { F_EPSMIN }); break;
    case SPHGPU_Q_M_ZERO: if (order == 0) set(1, { F_MZERO }); break;
    case SPHGPU_Q_EXPLICIT_GROWTH: if (order == 0) set(1, { F_GROWTH }); break;
    case SPHGPU_Q_N_FLAWS: if (order == 0) { set(1, { U_NFLAWS }); m.isU32 = true; } break;
    case SPHGPU_Q_FLAG: if (order == 0) { set(1, { U_FLAG }); m.isU32 = true; } break;
    case SPHGPU_Q_NEIGHBOR_CNT: if (order == 0) { set(1, { U_NCNT }); m.isU32 = true; } break;
    case SPHGPU_Q_MATERIAL_ID: if (order == 0) { set(1, { U_MATID }); m.isU32 = true; } break;
    case SPHGPU_Q_VELOCITY_ROTATION: if (order == 0) set(4, { F_ROTX, F_ROTY, F_ROTZ, -1 }); break;
    case SPHGPU_Q_XSPH_VELOCITIES: if (order == 0) set(4, { F_XSX, F_XSY, F_XSZ, -1 }); break;
    case SPHGPU_Q_DELTASPH_DENSITY_GRADIENT: if (order == 0) set(4, { F_DGX, F_DGY, F_DGZ, -1 }); break;
    case SPHGPU_Q_AV_STRESS: if (order == 0) set(6, { F_AS0, F_AS1, F_AS2, F_AS3, F_AS4, F_AS5 }); break;
    case SPHGPU_Q_INTERPARTICLE_SPACING_KERNEL: if (order == 0) set(1, { F_WP }); break;
    default: break;
    }
    return m;
}

struct HostLayout {
    int stride;     // in 8-byte words (f64 quantities) or 4-byte words (u32)
    int offset[6];
};

static HostLayout hostLayout(const QuantityMap& m, int layout) {
    HostLayout h{ m.ncomp, { 0, 1, 2, 3, 4, 5 } };
    if (layout == SPHGPU_LAYOUT_OPENSPH && !m.isU32) {
        if (m.ncomp == 5) { // TracelessTensor: Vector{xx,yy,xy,xz} + yz, 64 B (TracelessTensor.h:36-45)
            h.stride = 8;
        } else if (m.ncomp == 6) { // SymmetricTensor: diag Vector + off-diagonal Vector, 64 B (SymmetricTensor.h:18-21)
            h.stride = 8;
            h.offset[3] = 4;
            h.offset[4] = 5;
            h.offset[5] = 6;
        }
    }
    return h;
}

size_t elementBytes(int q, int layout) {
    for (int order = 0; order < 3; ++order) {
        const QuantityMap m = quantityMap(q, order);
        if (m.ncomp > 0) {
            const HostLayout h = hostLayout(m, layout);
            return (size_t)h.stride * (m.isU32 ? 4 : 8);
        }
    }
    return 0;
}

struct PlaneSet {
    double* f[6];
    uint32_t* u;
    int ncomp, stride;
    int offset[6];
};

__global__ void __launch_bounds__(256) k_unpack_f64(PlaneSet ps, const double* __restrict__ src, uint32_t first, uint32_t count) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) {
        return;
    }
    for (int c = 0; c < ps.ncomp; ++c) {
        if (ps.f[c]) {
            ps.f[c][first + k] = src[(size_t)k * ps.stride + ps.offset[c]];
        }
    }
}

__global__ void __launch_bounds__(256) k_pack_f64(PlaneSet ps, double* __restrict__ dst, uint32_t first, uint32_t count) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) {
        return;
    }
    for (int c = 0; c < ps.stride; ++c) {
        dst[(size_t)k * ps.stride + c] = 0.; // padding lanes are defined
    }
    for (int c = 0; c < ps.ncomp; ++c) {
        dst[(size_t)k * ps.stride + ps.offset[c]] = ps.f[c] ? ps.f[c][first + k] : 0.;
    }
}

static int makePlaneSet(sphgpu_ctx* ctx, int q, int order, int layout, PlaneSet& ps, bool& isU32) {
    const QuantityMap m = quantityMap(q, order);
    if (m.ncomp == 0) {
        setError("quantity " + std::to_string(q) + " has no buffer of order " + std::to_string(order));
        return SPHGPU_E_INVALID;
    }
    if (layout != SPHGPU_LAYOUT_PACKED && layout != SPHGPU_LAYOUT_OPENSPH) {
        setError("unknown host layout");
        return SPHGPU_E_INVALID;
    }
    const HostLayout h = hostLayout(m, layout);
    ps.ncomp = m.ncomp;
    ps.stride = h.stride;
    isU32 = m.isU32;
    ps.u = nullptr;
    for (int c = 0; c < 6; ++c) {
        ps.offset[c] = h.offset[c];
        ps.f[c] = nullptr;
    }
    if (m.isU32) {
        ps.u = ctx->d.u[m.field[0]];
    } else {
        for (int c = 0; c < m.ncomp; ++c) {
            ps.f[c] = m.field[c] >= 0 ? ctx->d.f[m.field[c]] : nullptr;
        }
    }
    return SPHGPU_OK;
}

// ---- halo records: {x,y,z,h, vx,vy,vz,vh, rho, u, S0..S4, D} ----------------------------------------------------
template <bool PACK>
__global__ void __launch_bounds__(256) k_halo(DevicePointers d, uint32_t first, uint32_t count, double* __restrict__ buf, bool solid,
    bool damage) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) {
        return;
    }
    const uint32_t i = first + k;
    const int planes[16] = { F_X, F_Y, F_Z, F_H, F_VX, F_VY, F_VZ, F_VH, F_RHO, F_U, F_S0, F_S1, F_S2, F_S3, F_S4, F_D };
    double2* rec = reinterpret_cast<double2*>(buf + (size_t)k * SPHGPU_HALO_DOUBLES);
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
        const bool use0 = (c < 10) || (c < 15 && solid) || (c == 15 && damage);
        const bool use1 = (c + 1 < 10) || (c + 1 < 15 && solid) || (c + 1 == 15 && damage);
        if (PACK) {
            rec[c / 2] = make_double2(use0 ? d.f[planes[c]][i] : 0., use1 ? d.f[planes[c + 1]][i] : 0.);
        } else {
            const double2 v = rec[c / 2];
            if (use0) {
                d.f[planes[c]][i] = v.x;
            }
            if (use1) {
                d.f[planes[c + 1]][i] = v.y;
            }
        }
    }
}

int launchHalo(sphgpu_ctx* ctx, bool pack, uint32_t first, uint32_t count, void* buf) {
    if (count == 0) {
        return SPHGPU_OK;
    }
    const uint32_t blocks = (count + 255) / 256;
    if (pack) {
        k_halo<true><<<blocks, 256, 0, ctx->stream>>>(ctx->d, first, count, (double*)buf, ctx->solid, ctx->hasDamage);
    } else {
        k_halo<false><<<blocks, 256, 0, ctx->stream>>>(ctx->d, first, count, (double*)buf, ctx->solid, ctx->hasDamage);
    }
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchUnpack(sphgpu_ctx* ctx, int q, int order, int layout, const void* stagingDev, uint32_t first, uint32_t count) {
    PlaneSet ps;
    bool isU32;
    const int rc = makePlaneSet(ctx, q, order, layout, ps, isU32);
    if (rc != SPHGPU_OK || count == 0) {
        return rc;
    }
    if (isU32) {
        SPH_CUDA_CHECK(cudaMemcpyAsync(ps.u + first, stagingDev, sizeof(uint32_t) * count, cudaMemcpyDeviceToDevice, ctx->stream));
    } else if (ps.ncomp == 1) {
        SPH_CUDA_CHECK(cudaMemcpyAsync(ps.f[0] + first, stagingDev, sizeof(double) * count, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        k_unpack_f64<<<(count + 255) / 256, 256, 0, ctx->stream>>>(ps, (const double*)stagingDev, first, count);
        SPH_CUDA_CHECK(cudaGetLastError());
    }
    return SPHGPU_OK;
}

int launchPack(sphgpu_ctx* ctx, int q, int order, int layout, void* stagingDev, uint32_t first, uint32_t count) {
    PlaneSet ps;
    bool isU32;
    const int rc = makePlaneSet(ctx, q, order, layout, ps, isU32);
    if (rc != SPHGPU_OK || count == 0) {
        return rc;
    }
    if (isU32) {
        SPH_CUDA_CHECK(cudaMemcpyAsync(stagingDev, ps.u + first, sizeof(uint32_t) * count, cudaMemcpyDeviceToDevice, ctx->stream));
    } else if (ps.ncomp == 1) {
        SPH_CUDA_CHECK(cudaMemcpyAsync(stagingDev, ps.f[0] + first, sizeof(double) * count, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        k_pack_f64<<<(count + 255) / 256, 256, 0, ctx->stream>>>(ps, (double*)stagingDev, first, count);
        SPH_CUDA_CHECK(cudaGetLastError());
    }
    return SPHGPU_OK;
}

} // namespace sph
