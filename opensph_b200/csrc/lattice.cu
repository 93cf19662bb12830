// Initial conditions on the device: the hexagonal close packing of a spherical body, its smoothing
// lengths, masses and body flag written straight into the slot planes -- what InitialConditions::addMonolithicBody
// (core/sph/initial/Initial.cpp:100-125) does with HexagonalPacking::generate (core/sph/initial/Distribution.cpp:126-200) and
// InitialConditions::setQuantities / getMasses (Initial.cpp:288-333) for a SphericalDomain
// (core/objects/geometry/Domain.cpp:10-33, Domain.h:134-136). At 10^7 .. 10^8 particles this removes the host-side O(N)
// set-up and the upload of positions and masses.
//
// The reference walks the bounding box with three nested loops that ADD the step to a running coordinate
// (Box::iterateWithIndices, core/objects/geometry/Box.h:182-195), so the coordinate of row k is not lower + k * step but the
// k-fold rounded sum. The three coordinate tables are therefore produced by the same repeated addition on the host (a few
// hundred to a few thousand entries) and the device only does the O(N) part: membership test in the reference's operation
// order, order-preserving compaction (z outermost, x innermost, like the reference's raster), the centring shift and the
// masses. Lattice points are bit-identical to the reference's; the centring shift and the mass normalisation involve sums
// over all particles, which the reference adds up sequentially and the device as a tree: equal to rounding.
#include "sphgpu_internal.h"
#include <cub/device/device_scan.cuh>
#include <vector>

namespace sph {

struct LatticeDev {
    const double *xs, *ys, *zs; // coordinate tables
    uint32_t nx, ny, nz;
    double deltaX, deltaY;
    double cx, cy, cz, radiusSqr;
    uint32_t* rowCount; // [ny * nz + 1]
    uint32_t* rowStart; // exclusive prefix
};

/// Lattice point (i, j, k) of HexagonalPacking::generate (Distribution.cpp:163-175) and the SphericalDomain test.
__device__ __forceinline__ bool latticePoint(const LatticeDev& l, uint32_t i, uint32_t j, uint32_t k, double& x, double& y, double& z) {
    x = l.xs[i];
    y = l.ys[j];
    z = l.zs[k];
    if (k % 2u == 0u) {
        if (j % 2u == 1u) {
            x = __dadd_rn(x, l.deltaX);
        }
    } else {
        if (j % 2u == 0u) {
            x = __dadd_rn(x, l.deltaX);
        }
        y = __dadd_rn(y, l.deltaY);
    }
    const double dx = x - l.cx, dy = y - l.cy, dz = z - l.cz;
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return d2 <= l.radiusSqr;
}

/// One warp per lattice row (j, k): FILL = false counts the points inside the domain, FILL = true writes them, in the order
/// of i, behind the row's prefix.
template <bool FILL>
__global__ void __launch_bounds__(256) k_lattice_rows(LatticeDev l, DevicePointers d, uint32_t first, double h) {
    const uint32_t lane = threadIdx.x & 31u, warpsPerBlock = blockDim.x >> 5;
    const uint32_t rows = l.ny * l.nz;
    for (uint32_t row = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); row < rows; row += gridDim.x * warpsPerBlock) {
        const uint32_t j = row % l.ny, k = row / l.ny;
        uint32_t base = FILL ? first + l.rowStart[row] : 0u, count = 0u;
        for (uint32_t i0 = 0; i0 < l.nx; i0 += 32u) {
            const uint32_t i = i0 + lane;
            double x = 0., y = 0., z = 0.;
            const bool in = i < l.nx && latticePoint(l, i, j, k, x, y, z);
            const uint32_t mask = __ballot_sync(0xffffffffu, in);
            if (FILL && in) {
                const uint32_t slot = base + __popc(mask & ((1u << lane) - 1u));
                d.f[F_X][slot] = x;
                d.f[F_Y][slot] = y;
                d.f[F_Z][slot] = z;
                d.f[F_H][slot] = h;
            }
            base += __popc(mask);
            count += __popc(mask);
        }
        if (!FILL && lane == 0u) {
            l.rowCount[row] = count;
        }
    }
}

/// Deterministic sums over the slots [first, first + count): per block in a fixed order, then the block results in order.
constexpr int LATTICE_BLOCKS = 296;
__global__ void __launch_bounds__(256) k_lattice_sums(DevicePointers d, uint32_t first, uint32_t count, double* partial) {
    double s[4] = { 0., 0., 0., 0. }; // x, y, z, h^3
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        const uint32_t i = first + t;
        const double h = d.f[F_H][i];
        s[0] += d.f[F_X][i];
        s[1] += d.f[F_Y][i];
        s[2] += d.f[F_Z][i];
        s[3] += h * h * h;
    }
    __shared__ double sm[8][4];
    for (int k = 0; k < 4; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 4; ++k) {
            sm[threadIdx.x >> 5][k] = s[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double r = 0.;
        for (int w = 0; w < 8; ++w) {
            r += sm[w][threadIdx.x];
        }
        partial[blockIdx.x * 4 + threadIdx.x] = r;
    }
}

/// HexagonalPacking's CENTER option (Distribution.cpp:186-198), getMasses (m ~ h^3 scaled to the total mass; the rows were
/// written with h * eta, Initial.cpp:308-313), the body flag, zero velocities and accelerations (Initial.cpp:288-323).
__global__ void __launch_bounds__(256) k_lattice_finish(DevicePointers d, uint32_t first, uint32_t count, const double* partial, int blocks, bool centre,
    double cx, double cy, double cz, double totalMass, uint32_t bodyFlag) {
    __shared__ double tot[4];
    if (threadIdx.x < 4) {
        double r = 0.;
        for (int b = 0; b < blocks; ++b) {
            r += partial[b * 4 + threadIdx.x];
        }
        tot[threadIdx.x] = r;
    }
    __syncthreads();
    const double inv = 1. / (double)count;
    const double sx = centre ? cx - tot[0] * inv : 0., sy = centre ? cy - tot[1] * inv : 0., sz = centre ? cz - tot[2] * inv : 0.;
    const double normalization = totalMass / tot[3]; // getMasses (Initial.cpp:288-306): m = h^3, scaled to the total mass
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += gridDim.x * blockDim.x) {
        const uint32_t i = first + t;
        d.f[F_X][i] += sx;
        d.f[F_Y][i] += sy;
        d.f[F_Z][i] += sz;
        const double h = d.f[F_H][i];
        d.f[F_M][i] = (h * h * h) * normalization;
        d.f[F_VX][i] = d.f[F_VY][i] = d.f[F_VZ][i] = d.f[F_VH][i] = 0.;
        d.f[F_AX][i] = d.f[F_AY][i] = d.f[F_AZ][i] = 0.;
        d.u[U_FLAG][i] = bodyFlag;
    }
}

struct LatticeHost {
    std::vector<double> xs, ys, zs;
    double h, deltaX, deltaY, volume;
};

/// The scalars and coordinate tables of HexagonalPacking::generate for a SphericalDomain (Distribution.cpp:133-152).
static bool latticeTables(const sphgpu_lattice& cfg, LatticeHost& out) {
    if (!(cfg.radius > 0.) || cfg.particle_count == 0) {
        return false;
    }
    const double PI = 3.14159265358979323846264338327950288;
    out.volume = 1.3333333333333333333333 * PI * (cfg.radius * cfg.radius * cfg.radius); // sphereVolume (MathUtils.h:393-395)
    const double particleDensity = (double)cfg.particle_count / out.volume;
    out.h = 1. / cbrt(particleDensity);
    const double dx = 1.1 * out.h;
    const double dy = sqrt(3.) * 0.5 * dx;
    const double dz = sqrt(6.) / 3. * dx;
    out.deltaX = 0.5 * dx;
    out.deltaY = sqrt(3.) / 6. * dx;
    const double step[3] = { dx, dy, dz };
    std::vector<double>* tab[3] = { &out.xs, &out.ys, &out.zs };
    for (int a = 0; a < 3; ++a) {
        const double lower = (cfg.center[a] - cfg.radius) + 0.5 * step[a], upper = cfg.center[a] + cfg.radius;
        for (double c = lower; c <= upper; c += step[a]) { // Box::iterateWithIndices: a running sum, not lower + k * step
            tab[a]->push_back(c);
            if (tab[a]->size() > (1u << 24)) {
                return false;
            }
        }
        if (tab[a]->empty()) {
            return false;
        }
    }
    return true;
}

/// Counts (ctx == null) or generates the lattice. Temporary device memory lives for the duration of the call.
static int runLattice(int device, sphgpu_ctx* ctx, const sphgpu_lattice& cfg, uint32_t first, uint32_t* countOut) {
    LatticeHost host;
    if (!latticeTables(cfg, host)) {
        setError("lattice: radius and particle count must be positive (and the lattice below 2^24 points per axis)");
        return SPHGPU_E_INVALID;
    }
    SPH_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st = ctx ? ctx->stream : nullptr;
    LatticeDev l{};
    l.nx = (uint32_t)host.xs.size();
    l.ny = (uint32_t)host.ys.size();
    l.nz = (uint32_t)host.zs.size();
    const uint64_t rows64 = (uint64_t)l.ny * l.nz;
    if (rows64 >= 0x7fffffffull) {
        setError("lattice too large");
        return SPHGPU_E_INVALID;
    }
    const uint32_t rows = (uint32_t)rows64;
    double* tables = nullptr;
    uint32_t* counts = nullptr;
    void* cubTemp = nullptr;
    double* partial = nullptr;
    int rc = SPHGPU_OK;
    auto cleanup = [&]() {
        cudaFree(tables);
        cudaFree(counts);
        cudaFree(cubTemp);
        cudaFree(partial);
    };
#define LATTICE_CHECK(expr)                                                                                           \
    do {                                                                                                              \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess) {                                                                                      \
            setError(std::string(#expr) + ": " + cudaGetErrorString(_e));                                            \
            cleanup();                                                                                                \
            return _e == cudaErrorMemoryAllocation ? SPHGPU_E_OOM : SPHGPU_E_CUDA;                                    \
        }                                                                                                             \
    } while (0)
    const size_t nTab = (size_t)l.nx + l.ny + l.nz;
    LATTICE_CHECK(cudaMalloc(&tables, nTab * sizeof(double)));
    LATTICE_CHECK(cudaMalloc(&counts, 2 * ((size_t)rows + 1) * sizeof(uint32_t)));
    LATTICE_CHECK(cudaMalloc(&partial, (size_t)LATTICE_BLOCKS * 4 * sizeof(double)));
    LATTICE_CHECK(cudaMemcpyAsync(tables, host.xs.data(), l.nx * sizeof(double), cudaMemcpyHostToDevice, st));
    LATTICE_CHECK(cudaMemcpyAsync(tables + l.nx, host.ys.data(), l.ny * sizeof(double), cudaMemcpyHostToDevice, st));
    LATTICE_CHECK(cudaMemcpyAsync(tables + l.nx + l.ny, host.zs.data(), l.nz * sizeof(double), cudaMemcpyHostToDevice, st));
    l.xs = tables;
    l.ys = tables + l.nx;
    l.zs = tables + l.nx + l.ny;
    l.deltaX = host.deltaX;
    l.deltaY = host.deltaY;
    l.cx = cfg.center[0];
    l.cy = cfg.center[1];
    l.cz = cfg.center[2];
    l.radiusSqr = cfg.radius * cfg.radius;
    l.rowCount = counts;
    l.rowStart = counts + rows + 1;
    LATTICE_CHECK(cudaMemsetAsync(counts, 0, 2 * ((size_t)rows + 1) * sizeof(uint32_t), st));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((uint64_t)sms * 8, ((uint64_t)rows + 7) / 8);
    DevicePointers none{};
    k_lattice_rows<false><<<blocks, 256, 0, st>>>(l, ctx ? ctx->d : none, 0u, 0.);
    size_t tempBytes = 0;
    LATTICE_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, l.rowCount, l.rowStart, (int)rows + 1, st));
    LATTICE_CHECK(cudaMalloc(&cubTemp, tempBytes + 16));
    LATTICE_CHECK(cub::DeviceScan::ExclusiveSum(cubTemp, tempBytes, l.rowCount, l.rowStart, (int)rows + 1, st));
    uint32_t total = 0;
    LATTICE_CHECK(cudaMemcpyAsync(&total, l.rowStart + rows, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    LATTICE_CHECK(cudaStreamSynchronize(st));
    if (countOut) {
        *countOut = total;
    }
    if (ctx && total > 0) {
        if ((uint64_t)first + total > ctx->capacity) {
            cleanup();
            setError("lattice: the generated particles do not fit the context's capacity (sphgpu_lattice_count tells how many there are)");
            return SPHGPU_E_INVALID;
        }
        k_lattice_rows<true><<<blocks, 256, 0, st>>>(l, ctx->d, first, host.h * cfg.eta);
        const int sumBlocks = (int)std::min<uint32_t>(LATTICE_BLOCKS, (total + 255) / 256);
        k_lattice_sums<<<sumBlocks, 256, 0, st>>>(ctx->d, first, total, partial);
        k_lattice_finish<<<(uint32_t)std::min<uint64_t>((uint64_t)sms * 8, ((uint64_t)total + 255) / 256), 256, 0, st>>>(ctx->d, first, total, partial,
            sumBlocks, (cfg.flags & SPHGPU_LATTICE_CENTER) != 0, cfg.center[0], cfg.center[1], cfg.center[2], host.volume * cfg.density,
            cfg.body_flag);
        LATTICE_CHECK(cudaGetLastError());
        LATTICE_CHECK(cudaStreamSynchronize(st));
        ctx->launches += 4;
        ctx->listsDirty = true;
        ctx->stateUploaded = true;
    }
#undef LATTICE_CHECK
    cleanup();
    return rc;
}

} // namespace sph

extern "C" {

int sphgpu_lattice_count(int device, const sphgpu_lattice* cfg, uint32_t* count) {
    if (!cfg || !count) {
        sph::setError("null argument");
        return SPHGPU_E_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        sph::setError("no CUDA device available; libsphgpu has no CPU path");
        return SPHGPU_E_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) {
        sph::setError("device index out of range");
        return SPHGPU_E_NO_DEVICE;
    }
    return sph::runLattice(device, nullptr, *cfg, 0u, count);
}

int sphgpu_lattice_generate(sphgpu_ctx* ctx, const sphgpu_lattice* cfg, uint32_t first, uint32_t* count) {
    if (!ctx || !cfg) {
        sph::setError("null argument");
        return SPHGPU_E_INVALID;
    }
    if (!(cfg->eta > 0.) || !(cfg->density > 0.)) {
        sph::setError("lattice: eta and density must be positive");
        return SPHGPU_E_INVALID;
    }
    return sph::runLattice(ctx->device, ctx, *cfg, first, count);
}

} // extern "C"
