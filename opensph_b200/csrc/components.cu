// Connected components of the particle cloud on the device: Post::findComponents of the reference
// (core/post/Analysis.cpp:36-75,115-128) with ComponentFlag::OVERLAP and, optionally, SEPARATE_BY_FLAG.
//
// The reference visits the particles in index order; an unassigned one opens the next component and floods it through a
// stack over the DIRECTED relation "j lies within h_index * radius of index" (findAll(index, r[index][H] * radius)), for
// SEPARATE_BY_FLAG restricted to equal body flags. With unequal smoothing lengths the relation is not symmetric, so the result
// is not the undirected connected components. It is, however, order-free: the root of a particle p -- the particle that opened
// p's component -- is the ancestor of p with the LOWEST INDEX (ancestor: anything with a directed path to p, p included).
// Proof: let a be that ancestor. Nothing with a lower index reaches a (it would reach p as well), so no earlier flood took a
// and a opens a component; that flood reaches p, and no earlier root does. Hence
//     label(p) = min(p, min over q -> p of label(q))
// iterated to the fixed point gives the roots, and the component indices are the ranks of the roots in index order. Labels
// are pushed along the edges with atomicMin; since an ancestor of an ancestor is an ancestor, label(p) may also jump to
// label(label(p)) (pointer jumping), which takes the iteration count from the graph's diameter to a few dozen.
//
// The sweep reuses the cell list and the sorted position records of the SPH step (grid.cu, k_pack_positions): one thread per
// particle walks the cells its reach h * radius touches. Integer work on exact FP64 predicates: the result is bit-identical to
// the reference's.
#include "sphgpu_internal.h"

#include <algorithm>
#include <cub/device/device_scan.cuh>
#include <vector>

namespace sph {

/// One sweep. lab[t]: label (an original particle index) of the particle at sorted position t.
template <bool BY_FLAG>
__global__ void __launch_bounds__(128) k_comp_sweep(DevicePointers d, uint32_t n, int recDoubles, double radius, uint32_t* lab, uint32_t* changed) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) {
        return;
    }
    const GridDev g = *d.grid;
    double2 ixy, izh;
    loadSortedPosition(d.rec, t, recDoubles, ixy, izh);
    const double xi = ixy.x, yi = ixy.y, zi = izh.x;
    const double reach = izh.y * radius;
    const double reachSqr = reach * reach;
    // pointer jumping first: the label of my label's owner is an ancestor of mine as well
    uint32_t L = lab[t];
    {
        const uint32_t up = lab[d.rank[L]];
        if (up < L) {
            L = up;
            atomicMin(&lab[t], L);
            *changed = 1u;
        }
    }
    const uint32_t myFlag = BY_FLAG ? d.u[U_FLAG][d.order[t]] : 0u;
    bool any = false;
    auto visit = [&](uint32_t k) {
        double2 pxy, pzh;
        loadSortedPosition(d.rec, k, recDoubles, pxy, pzh);
        const double dx = pxy.x - xi, dy = pxy.y - yi, dz = pzh.x - zi;
        // getSqrLength(r_j - r_i) < (h_i radius)^2, products and sums rounded one by one like the reference's
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (!(d2 < reachSqr)) {
            return;
        }
        if (BY_FLAG && d.u[U_FLAG][d.order[k]] != myFlag) {
            return;
        }
        if (lab[k] > L) {
            atomicMin(&lab[k], L);
            any = true;
        }
    };
    if (g.nLarge > 0u && t >= g.largeBegin) { // a large particle (two-level radii) is in no cell: it may reach anything
        for (uint32_t k = 0; k < n; ++k) {
            visit(k);
        }
    } else {
        const int x0 = max((int)floor((xi - reach - g.lo[0]) * g.cellInv), 0), x1 = min((int)floor((xi + reach - g.lo[0]) * g.cellInv), g.dim[0] - 1);
        const int y0 = max((int)floor((yi - reach - g.lo[1]) * g.cellInv), 0), y1 = min((int)floor((yi + reach - g.lo[1]) * g.cellInv), g.dim[1] - 1);
        const int z0 = max((int)floor((zi - reach - g.lo[2]) * g.cellZInv), 0), z1 = min((int)floor((zi + reach - g.lo[2]) * g.cellZInv), g.dim[2] - 1);
        for (int z = z0; z <= z1; ++z) {
            for (int y = y0; y <= y1; ++y) {
                const uint32_t row = (uint32_t)((z * g.dim[1] + y) * g.dim[0]);
                const uint32_t s = d.cellStart[row + x0], e = d.cellStart[row + x1 + 1];
                for (uint32_t k = s; k < e; ++k) {
                    visit(k);
                }
            }
        }
        if (g.nLarge > 0u) { // the large particles are candidates of everyone
            for (uint32_t k = g.largeBegin; k < n; ++k) {
                visit(k);
            }
        }
    }
    if (any) {
        *changed = 1u;
    }
}

__global__ void __launch_bounds__(256) k_comp_init(DevicePointers d, uint32_t n, uint32_t* lab) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        lab[t] = d.order[t];
    }
}

/// isRoot[i] = 1 when particle i (original index) kept its own label.
__global__ void __launch_bounds__(256) k_comp_roots(DevicePointers d, uint32_t n, const uint32_t* lab, uint32_t* isRoot) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        isRoot[i] = lab[d.rank[i]] == i ? 1u : 0u;
    }
}

/// indices[i] = rank of the root of i among the roots in index order.
__global__ void __launch_bounds__(256) k_comp_indices(DevicePointers d, uint32_t n, const uint32_t* lab, const uint32_t* rootRank, uint32_t* indices) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        indices[i] = rootRank[lab[d.rank[i]]];
    }
}

} // namespace sph

using namespace sph;

extern "C" int sphgpu_find_components(sphgpu_ctx* ctx, double radius, uint32_t flags, uint32_t* indices, uint32_t* component_count, uint32_t* sweeps) {
    if (!ctx || !indices || !component_count) {
        setError("null argument");
        return SPHGPU_E_INVALID;
    }
    if (!(radius > 0.)) {
        setError("the component radius must be positive");
        return SPHGPU_E_INVALID;
    }
    if (flags & ~(uint32_t)SPHGPU_COMPONENTS_SEPARATE_BY_FLAG) {
        setError("unknown component flag (ESCAPE_VELOCITY and SORT_BY_MASS are applied by the caller to the result)");
        return SPHGPU_E_INVALID;
    }
    if (ctx->nActive != ctx->n) {
        setError("components are searched in a single domain (the context holds ghost particles)");
        return SPHGPU_E_STATE;
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->n;
    *component_count = 0;
    if (sweeps) {
        *sweeps = 0;
    }
    if (n == 0) {
        return SPHGPU_OK;
    }
    int rc;
    // cell list and sorted positions of the current state (the next evaluation rebuilds its own lists)
    ctx->listsDirty = true;
    if ((rc = launchGridBuild(ctx)) != SPHGPU_OK) return rc;
    if ((rc = launchProloguePackPositionsOnly(ctx)) != SPHGPU_OK) return rc;
    ctx->listsDirty = true;
    cudaStream_t st = ctx->stream;
    uint32_t *lab = nullptr, *aux = nullptr, *rootRank = nullptr, *changedDev = nullptr;
    void* scanTmp = nullptr;
    size_t scanBytes = 0;
    auto release = [&]() {
        cudaFree(lab);
        cudaFree(aux);
        cudaFree(rootRank);
        cudaFree(changedDev);
        cudaFree(scanTmp);
    };
    auto check = [&](cudaError_t e) {
        if (e != cudaSuccess) {
            setError(cudaGetErrorString(e));
            release();
            return false;
        }
        return true;
    };
    if (!check(cudaMalloc(&lab, sizeof(uint32_t) * n)) || !check(cudaMalloc(&aux, sizeof(uint32_t) * n)) ||
        !check(cudaMalloc(&rootRank, sizeof(uint32_t) * n)) || !check(cudaMalloc(&changedDev, sizeof(uint32_t)))) {
        return SPHGPU_E_OOM;
    }
    cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, aux, rootRank, (int)n, st);
    if (!check(cudaMalloc(&scanTmp, std::max<size_t>(scanBytes, 16)))) {
        return SPHGPU_E_OOM;
    }
    const uint32_t blocks256 = (n + 255) / 256, blocks128 = (n + 127) / 128;
    k_comp_init<<<blocks256, 256, 0, st>>>(ctx->d, n, lab);
    const bool byFlag = (flags & SPHGPU_COMPONENTS_SEPARATE_BY_FLAG) != 0;
    uint32_t done = 0;
    for (;;) {
        uint32_t changed = 0;
        if (!check(cudaMemsetAsync(changedDev, 0, sizeof(uint32_t), st))) return SPHGPU_E_CUDA;
        if (byFlag) {
            k_comp_sweep<true><<<blocks128, 128, 0, st>>>(ctx->d, n, ctx->recDoubles, radius, lab, changedDev);
        } else {
            k_comp_sweep<false><<<blocks128, 128, 0, st>>>(ctx->d, n, ctx->recDoubles, radius, lab, changedDev);
        }
        ++done;
        if (!check(cudaMemcpyAsync(&changed, changedDev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st)) || !check(cudaStreamSynchronize(st))) {
            return SPHGPU_E_CUDA;
        }
        if (!changed) {
            break;
        }
        if (done > n + 2u) { // (cannot happen: every sweep that reports a change lowers at least one label)
            setError("component labels did not converge");
            release();
            return SPHGPU_E_STATE;
        }
    }
    k_comp_roots<<<blocks256, 256, 0, st>>>(ctx->d, n, lab, aux);
    cub::DeviceScan::ExclusiveSum(scanTmp, scanBytes, aux, rootRank, (int)n, st);
    k_comp_indices<<<blocks256, 256, 0, st>>>(ctx->d, n, lab, rootRank, aux);
    uint32_t lastRank = 0, lastIsRoot = 0;
    // (aux now holds the indices; the last root flag is recomputed from the label of the last particle)
    std::vector<uint32_t> tail(2);
    if (!check(cudaMemcpyAsync(indices, aux, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, st)) ||
        !check(cudaMemcpyAsync(&lastRank, rootRank + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st)) || !check(cudaStreamSynchronize(st))) {
        return SPHGPU_E_CUDA;
    }
    // number of roots = exclusive rank of the last particle + whether it is a root itself (it is one iff its index is its own
    // component's highest rank, i.e. its component index equals its exclusive rank)
    lastIsRoot = indices[n - 1] == lastRank ? 1u : 0u;
    *component_count = lastRank + lastIsRoot;
    if (sweeps) {
        *sweeps = done;
    }
    ctx->launches += done + 4;
    release();
    return SPHGPU_OK;
}
