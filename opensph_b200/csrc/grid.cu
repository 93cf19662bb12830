// Cell-list build: bounds reduction -> grid parameters (on device, no host round trip) -> cell histogram ->
// exclusive scan -> scatter -> per-cell ordering. Replaces the finder build of the reference
// (ISymmetricFinder::build, core/objects/finders/NeighborFinder.cpp:7-26; KdTree::buildImpl,
// core/objects/finders/KdTree.inl.h:10-41; LookupMap::update, core/objects/containers/LookupMap.h:31-45) and
// IAsymmetricSolver::getMaxSearchRadius (core/sph/solvers/AsymmetricSolver.cpp:104-111).
//
// The cell edge is a = R * h_max (x 1+1e-6) in x and y and a/2 in z, so every neighbour j of i
// (|r_i - r_j| < R (h_i + h_j)/2 <= R h_max) lies within +-1 cell in x,y and +-2 cells in z. The half-height cells let
// the tiled pair kernel pair the z-layers (-2,+3), (-1,+2), (0,+1) of a double row so that its warps stay balanced.
// If that would need more than maxCells cells the edge is enlarged, like the reference's UniformGridFinder caps its grid at (cbrt(N)+1)^3 cells (UniformGrid.cpp:16).
#include "sphgpu_internal.h"

namespace sph {

__device__ __forceinline__ double warpMin(double v) {
    for (int o = 16; o > 0; o >>= 1) {
        v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;
}
__device__ __forceinline__ double warpMax(double v) {
    for (int o = 16; o > 0; o >>= 1) {
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;
}

// AdaptiveSmoothingLength::initialize (h clamp, EquationTerm.cpp:356-364) fused with the bounding-box / h_max pass and
// with the displacement check of the list reuse (ListCtlDev): how far every particle has moved, relative to R h, and how
// much its h has grown since the lists were built.
// Displacement boxes of the list reuse are kept per block of 2^DISP_BX x 2^DISP_BY x 2^DISP_BZ cells.
constexpr int DISP_BX = 2, DISP_BY = 1, DISP_BZ = 2;

/// Order-preserving map of a float to an unsigned key (atomicMin / atomicMax on floats of either sign) and back.
__device__ __forceinline__ uint32_t floatKey(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float keyFloat(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(256) k_bounds(DevicePointers d, uint32_t nActive, bool clampH, double hMin, double hMax, double kernelRadius,
    bool trackCells, uint32_t cellStride) {
    double lo[3] = { INFTY_REF, INFTY_REF, INFTY_REF }, hi[3] = { -INFTY_REF, -INFTY_REF, -INFTY_REF }, hm = 0.;
    double grow = 0., hsum = 0., hNegMin0 = -INFTY_REF;
    const uint32_t overflowCell = d.grid->ncells;
    const uint32_t dimx = (uint32_t)max(d.grid->dim[0], 1), dimy = (uint32_t)max(d.grid->dim[1], 1);
    const uint32_t nbx = (dimx + (1u << DISP_BX) - 1u) >> DISP_BX, nby = (dimy + (1u << DISP_BY) - 1u) >> DISP_BY;
    const double gx = d.grid->lo[0], gy = d.grid->lo[1], gz = d.grid->lo[2]; // origin of the grid the lists were built on
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nActive; i += gridDim.x * blockDim.x) {
        double h = d.f[F_H][i];
        if (clampH) {
            const double hc = fmax(hMin, fmin(h, hMax));
            if (hc != h) {
                d.f[F_H][i] = hc;
            }
            h = hc;
        }
        const double x = d.f[F_X][i], y = d.f[F_Y][i], z = d.f[F_Z][i];
        lo[0] = fmin(lo[0], x);
        lo[1] = fmin(lo[1], y);
        lo[2] = fmin(lo[2], z);
        hi[0] = fmax(hi[0], x);
        hi[1] = fmax(hi[1], y);
        hi[2] = fmax(hi[2], z);
        hm = fmax(hm, h);
        hsum += h;
        const float4 p0 = d.pos0[i];
        const double ex = (x - gx) - (double)p0.x, ey = (y - gy) - (double)p0.y, ez = (z - gz) - (double)p0.z;
        grow = fmax(grow, h / (double)p0.w - 1.);
        // (cellOf belongs to the lists in use; large particles are paired directly every step)
        uint32_t blk = 0xffffffffu;
        if (trackCells) {
            const uint32_t c = d.cellOf[i];
            if (c < overflowCell) {
                hNegMin0 = fmax(hNegMin0, -(double)p0.w);
                const uint32_t cx = c % dimx, cy = (c / dimx) % dimy, cz = c / (dimx * dimy);
                blk = ((cz >> DISP_BZ) * nby + (cy >> DISP_BY)) * nbx + (cx >> DISP_BX);
            }
        }
        // consecutive slots are neighbours along x on a lattice: the lanes of a warp fall into a few blocks; one set of
        // atomics per block and warp
        const unsigned active = __activemask();
        const unsigned same = __match_any_sync(active, blk);
        if (blk != 0xffffffffu) {
            const uint32_t lx = __reduce_min_sync(same, floatKey(__double2float_rd(ex))), ly = __reduce_min_sync(same, floatKey(__double2float_rd(ey))),
                           lz = __reduce_min_sync(same, floatKey(__double2float_rd(ez)));
            const uint32_t ux = __reduce_max_sync(same, floatKey(__double2float_ru(ex))), uy = __reduce_max_sync(same, floatKey(__double2float_ru(ey))),
                           uz = __reduce_max_sync(same, floatKey(__double2float_ru(ez)));
            if ((threadIdx.x & 31u) == (uint32_t)(__ffs((int)same) - 1)) {
                atomicMin(&d.dispA[blk], lx);
                atomicMin(&d.dispA[cellStride + blk], ly);
                atomicMin(&d.dispA[2 * cellStride + blk], lz);
                atomicMax(&d.dispA[3 * cellStride + blk], ux);
                atomicMax(&d.dispA[4 * cellStride + blk], uy);
                atomicMax(&d.dispA[5 * cellStride + blk], uz);
            }
        }
    }
    hNegMin0 = warpMax(hNegMin0);
    for (int o = 16; o > 0; o >>= 1) {
        hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
    }
    __shared__ double sm[8][11];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[11] = { warpMin(lo[0]), warpMin(lo[1]), warpMin(lo[2]), warpMax(hi[0]), warpMax(hi[1]), warpMax(hi[2]), warpMax(hm),
        0. /* (was: largest absolute displacement) */, warpMax(grow), hsum, hNegMin0 };
    if (lane == 0) {
        for (int k = 0; k < 11; ++k) {
            sm[warp][k] = v[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 11) {
        const int k = threadIdx.x;
        double r = sm[0][k];
        for (int w = 1; w < 8; ++w) {
            r = (k < 3) ? fmin(r, sm[w][k]) : (k == 9 ? r + sm[w][k] : fmax(r, sm[w][k]));
        }
        // ([10], [11] of a block's row belong to k_hmax_small; minus the smallest build-time h goes to [12])
        d.boundsPartial[blockIdx.x * BOUNDS_STRIDE + (k < 10 ? k : 12)] = r;
    }
}

// ---- relative displacement since the list build (ListCtlDev) ------------------------------------------------------------
/// Bound of |u_i - u_j| over the particles i of a block of cells and j of the 27 blocks around it -> dispGlobal[6] (largest
/// over all blocks), and the global box of u -> dispGlobal[0..5]. Blocks are 4 x 2 x 4 cells (DISP_B*): one more block in
/// every direction covers the candidate stencil (+-1 cell in x, y, +-2 in z) plus one more cell edge.
__global__ void __launch_bounds__(256) k_disp_check(DevicePointers d, uint32_t cellStride, bool trackCells) {
    if (!trackCells) {
        return;
    }
    const GridDev g = *d.grid;
    const int nbx = (g.dim[0] + (1 << DISP_BX) - 1) >> DISP_BX, nby = (g.dim[1] + (1 << DISP_BY) - 1) >> DISP_BY,
              nbz = (g.dim[2] + (1 << DISP_BZ) - 1) >> DISP_BZ;
    const uint32_t nBlocks = (uint32_t)(nbx * nby * nbz);
    uint32_t glo[3] = { 0xffffffffu, 0xffffffffu, 0xffffffffu }, ghi[3] = { 0u, 0u, 0u };
    float worst = 0.f;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nBlocks; b += gridDim.x * blockDim.x) {
        if (d.dispA[3 * cellStride + b] == 0u) {
            continue; // no particle in this block
        }
        const int bx = (int)(b % (uint32_t)nbx), by = (int)((b / (uint32_t)nbx) % (uint32_t)nby), bz = (int)(b / (uint32_t)(nbx * nby));
        uint32_t wlo[3] = { 0xffffffffu, 0xffffffffu, 0xffffffffu }, whi[3] = { 0u, 0u, 0u };
        for (int z = max(bz - 1, 0); z <= min(bz + 1, nbz - 1); ++z) {
            for (int y = max(by - 1, 0); y <= min(by + 1, nby - 1); ++y) {
                for (int x = max(bx - 1, 0); x <= min(bx + 1, nbx - 1); ++x) {
                    const uint32_t e = (uint32_t)((z * nby + y) * nbx + x);
#pragma unroll
                    for (int k = 0; k < 3; ++k) { // (untouched blocks hold the neutral keys ~0 / 0)
                        wlo[k] = min(wlo[k], d.dispA[k * cellStride + e]);
                        whi[k] = max(whi[k], d.dispA[(3 + k) * cellStride + e]);
                    }
                }
            }
        }
        float s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t ol = d.dispA[k * cellStride + b], oh = d.dispA[(3 + k) * cellStride + b];
            const float e = fmaxf(__fsub_ru(keyFloat(oh), keyFloat(wlo[k])), __fsub_ru(keyFloat(whi[k]), keyFloat(ol)));
            s2 = __fmaf_ru(e, e, s2);
            glo[k] = min(glo[k], ol);
            ghi[k] = max(ghi[k], oh);
        }
        worst = fmaxf(worst, __fsqrt_ru(s2));
    }
    for (int o = 16; o > 0; o >>= 1) {
        worst = fmaxf(worst, __shfl_xor_sync(0xffffffffu, worst, o));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            glo[k] = min(glo[k], __shfl_xor_sync(0xffffffffu, glo[k], o));
            ghi[k] = max(ghi[k], __shfl_xor_sync(0xffffffffu, ghi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&d.dispGlobal[6], __float_as_uint(worst)); // non-negative floats order like their bit patterns
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&d.dispGlobal[k], glo[k]);
            atomicMax(&d.dispGlobal[3 + k], ghi[k]);
        }
    }
}

/// Largest h among the particles with h <= hSplit and the number of particles above it (only when the lists are being
/// rebuilt and some particle exceeds the split).
__global__ void __launch_bounds__(256) k_hmax_small(DevicePointers d, uint32_t nActive) {
    const ListCtlDev ctl = *d.listCtl;
    if (ctl.rebuild == 0u || !(ctl.hmaxAll > ctl.hSplit)) {
        return;
    }
    double hm = 0., cnt = 0.;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nActive; i += gridDim.x * blockDim.x) {
        const double h = d.f[F_H][i];
        if (h > ctl.hSplit) {
            cnt += 1.;
        } else {
            hm = fmax(hm, h);
        }
    }
    hm = warpMax(hm);
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ double sm[8][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        sm[warp][0] = hm;
        sm[warp][1] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            hm = fmax(hm, sm[w][0]);
            cnt += sm[w][1];
        }
        d.boundsPartial[blockIdx.x * BOUNDS_STRIDE + 10] = fmax(sm[0][0], hm);
        d.boundsPartial[blockIdx.x * BOUNDS_STRIDE + 11] = cnt;
    }
}

/// Reduces the partial bounds and decides whether this integrate() rebuilds the cell list / units / candidate lists (force,
/// or the displacement metric has used up the skin; the margin covers the FP32 rounding of pos0). Every later build
/// kernel reads ListCtlDev::rebuild and returns at once when it is 0. Also fixes the split of the search radii.
__global__ void __launch_bounds__(256) k_grid_decide(DevicePointers d, int nPartials, uint32_t nActive, bool force, double skin,
    double kernelRadius) {
    __shared__ double sm[8][11];
    double v[11] = { INFTY_REF, INFTY_REF, INFTY_REF, -INFTY_REF, -INFTY_REF, -INFTY_REF, 0., 0., 0., 0., -INFTY_REF };
    for (int b = threadIdx.x; b < nPartials; b += blockDim.x) {
        for (int k = 0; k < 11; ++k) {
            const double p = d.boundsPartial[b * BOUNDS_STRIDE + (k < 10 ? k : 12)];
            v[k] = (k < 3) ? fmin(v[k], p) : (k == 9 ? v[k] + p : fmax(v[k], p));
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < 11; ++k) {
        if (k == 9) {
            for (int o = 16; o > 0; o >>= 1) {
                v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            }
        } else {
            v[k] = (k < 3) ? warpMin(v[k]) : warpMax(v[k]);
        }
    }
    if (lane == 0) {
        for (int k = 0; k < 11; ++k) {
            sm[warp][k] = v[k];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 11; ++k) {
            double r = sm[0][k];
            for (int w = 1; w < 8; ++w) {
                r = (k < 3) ? fmin(r, sm[w][k]) : (k == 9 ? r + sm[w][k] : fmax(r, sm[w][k]));
            }
            // the first slots of the partial array carry the totals to k_grid_params
            if (k < 10) {
                d.boundsPartial[k] = r;
            }
            v[k] = r;
        }
        ListCtlDev ctl = *d.listCtl;
        // Relative displacement of particles that can reach each other (k_disp_check), in units of the smallest reach
        // R h_min0, plus the largest growth of h. The FP32 copies pos0 are exact to 2^-24 of the grid extent per coordinate.
        double metric = INFTY_REF;
        bool farSafe = false;
        if (!force) {
            const GridDev g = *d.grid;
            const double slack = 4.e-7 * g.extent;
            const double hMin0 = -v[10];
            const double rel = (double)__uint_as_float(d.dispGlobal[6]) + slack;
            metric = (hMin0 > 0. ? rel / (kernelRadius * hMin0) : 0.) + fmax(v[8], 0.);
            // pairs beyond the window were at least two cell edges apart: safe while all of u fits into 0.9 cell edges
            double ext2 = 0.;
            for (int k = 0; k < 3; ++k) {
                const uint32_t kl = d.dispGlobal[k], kh = d.dispGlobal[3 + k];
                const double e = kh >= kl ? (double)keyFloat(kh) - (double)keyFloat(kl) : 0.;
                ext2 += e * e;
            }
            farSafe = sqrt(ext2) + slack < 0.9 * g.cell;
        }
        ctl.lastMetric = metric;
        const bool rebuild = force || !(metric < 0.9 * skin) || !farSafe;
        ctl.rebuild = rebuild ? 1u : 0u;
        if (rebuild) {
            ctl.age = 0u;
            ctl.rebuilds++;
            ctl.fallbackUnits = 0u;
            ctl.hmaxAll = v[6];
            ctl.hSplit = LARGE_FACTOR * v[9] / fmax((double)nActive, 1.);
        } else {
            ctl.age++;
        }
        *d.listCtl = ctl;
    }
}

/// Sets up the grid of a rebuild: cell edge R h_max (1 + skin) with h_max the largest SMALL h (GridDev::hSplit).
__global__ void k_grid_params(DevicePointers d, int nPartials, double kernelRadius, uint32_t maxCells, double skin) {
    if (threadIdx.x != 0 || d.listCtl->rebuild == 0u) {
        return;
    }
    const ListCtlDev ctl = *d.listCtl;
    double v[7];
    for (int k = 0; k < 7; ++k) {
        v[k] = d.boundsPartial[k];
    }
    GridDev g;
    g.hmaxAll = ctl.hmaxAll;
    g.hmax = ctl.hmaxAll;
    g.hSplit = INFTY_REF;
    g.nLarge = 0u;
    g.largeBegin = 0u;
    if (ctl.hmaxAll > ctl.hSplit) { // some particles are much larger than the mean: keep them out of the cell list
        double hSmall = 0., nLarge = 0.;
        for (int b = 0; b < nPartials; ++b) {
            hSmall = fmax(hSmall, d.boundsPartial[b * BOUNDS_STRIDE + 10]);
            nLarge += d.boundsPartial[b * BOUNDS_STRIDE + 11];
        }
        if (nLarge <= (double)LARGE_MAX && hSmall > 0.) {
            g.hmax = hSmall;
            g.hSplit = ctl.hSplit;
            g.nLarge = (uint32_t)nLarge;
        }
    }
    {
        double cell = kernelRadius * g.hmax * (1. + 1.e-6) * (1. + skin);
        if (!(cell > 0.)) {
            cell = 1.;
        }
        double ext[3];
        for (int k = 0; k < 3; ++k) {
            g.lo[k] = v[k];
            ext[k] = fmax(v[3 + k] - v[k], 0.);
        }
        for (int iter = 0; iter < 64; ++iter) {
            double total = 1.;
            for (int k = 0; k < 3; ++k) {
                total *= floor(ext[k] / (k == 2 ? 0.5 * cell : cell)) + 1.;
            }
            if (total <= (double)(maxCells - 1u)) { // (one more cell, the overflow cell of the large particles, must fit)
                break;
            }
            cell *= fmax(cbrt(total / (double)(maxCells - 1u)), 1.0) * 1.02;
        }
        g.cell = cell;
        g.cellInv = 1. / cell;
        g.cellZ = 0.5 * cell;
        g.cellZInv = 1. / g.cellZ;
        uint32_t n = 1;
        for (int k = 0; k < 3; ++k) {
            g.dim[k] = (int)floor(ext[k] / (k == 2 ? g.cellZ : cell)) + 1;
            n *= (uint32_t)g.dim[k];
        }
        g.ncells = n;
        g.extent = fmax(fmax(ext[0], ext[1]), ext[2]) + cell;
        g.unsorted = 0u;
        *d.grid = g;
    }
}

__device__ __forceinline__ uint32_t cellIndex(const GridDev& g, double x, double y, double z) {
    int cx = (int)floor((x - g.lo[0]) * g.cellInv);
    int cy = (int)floor((y - g.lo[1]) * g.cellInv);
    int cz = (int)floor((z - g.lo[2]) * g.cellZInv);
    cx = min(max(cx, 0), g.dim[0] - 1);
    cy = min(max(cy, 0), g.dim[1] - 1);
    cz = min(max(cz, 0), g.dim[2] - 1);
    return (uint32_t)((cz * g.dim[1] + cy) * g.dim[0] + cx);
}

__global__ void __launch_bounds__(256) k_cell_count(DevicePointers d, uint32_t nActive) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nActive || d.listCtl->rebuild == 0u) {
        return;
    }
    const GridDev g = *d.grid;
    // large particles (two-level radii) sit in the overflow cell behind the last real one: never targets or candidates of
    // the tiled kernels
    const uint32_t c = d.f[F_H][i] > g.hSplit ? g.ncells : cellIndex(g, d.f[F_X][i], d.f[F_Y][i], d.f[F_Z][i]);
    d.cellOf[i] = c;
    d.rank[i] = atomicAdd(&d.cellCount[c], 1u);
}

// ---- exclusive scan of cellCount[0..total) into cellStart, three passes -----------------------------------
__global__ void __launch_bounds__(512) k_scan_block(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
    uint32_t* __restrict__ blockSums, uint32_t total, const ListCtlDev* ctl) {
    // each thread owns 8 consecutive items
    __shared__ uint32_t warpSums[16];
    if (ctl->rebuild == 0u) {
        return;
    }
    const uint32_t base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 8;
    uint32_t v[8];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = (base + k < total) ? in[base + k] : 0u;
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) {
            incl += t;
        }
    }
    if (lane == 31) {
        warpSums[warp] = incl;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < 16) ? warpSums[lane] : 0u;
        for (int o = 1; o < 16; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) {
                w += t;
            }
        }
        if (lane < 16) {
            warpSums[lane] = w; // inclusive over warps
        }
    }
    __syncthreads();
    uint32_t excl = incl - sum + (warp > 0 ? warpSums[warp - 1] : 0u);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (base + k < total) {
            out[base + k] = excl;
        }
        excl += v[k];
    }
    if (threadIdx.x == 511) {
        blockSums[blockIdx.x] = warpSums[15];
    }
}

__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* blockSums, uint32_t nBlocks, const ListCtlDev* ctl) {
    // single block, sequential over chunks of 1024
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    if (ctl->rebuild == 0u) {
        return;
    }
    if (threadIdx.x == 0) {
        carry = 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < nBlocks; base += 1024) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t val = idx < nBlocks ? blockSums[idx] : 0u;
        uint32_t incl = val;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) {
                incl += t;
            }
        }
        if (lane == 31) {
            warpSums[warp] = incl;
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warpSums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) {
                    w += t;
                }
            }
            warpSums[lane] = w;
        }
        __syncthreads();
        const uint32_t excl = incl - val + (warp > 0 ? warpSums[warp - 1] : 0u) + carry;
        if (idx < nBlocks) {
            blockSums[idx] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) {
            carry = excl + val;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(512) k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ blockSums, uint32_t total,
    const ListCtlDev* ctl) {
    if (ctl->rebuild == 0u) {
        return;
    }
    const uint32_t base = blockIdx.x * SCAN_ITEMS + threadIdx.x * 8;
    const uint32_t add = blockSums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (base + k < total) {
            out[base + k] += add;
        }
    }
}

__global__ void __launch_bounds__(256) k_scatter(DevicePointers d, uint32_t nActive) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nActive || d.listCtl->rebuild == 0u) {
        return;
    }
    d.order[d.cellStart[d.cellOf[i]] + d.rank[i]] = i;
}

// The histogram ranks come from atomics and are not reproducible; order every cell by (x, slot index) so that the
// summation order (and therefore every bit of the result) is the same from run to run. Cells of one row are ordered
// in x already, so every cell ROW ends up sorted by x: the pair kernel finds the x-window a target can reach in a
// candidate row by bisection. Also records the largest h of every cell (bound of the conservative FP32 pre-filter).
constexpr int SORT_LOCAL = 48;
__device__ __forceinline__ void sortCell(const DevicePointers& d, uint32_t c) {
    const uint32_t s = d.cellStart[c], e = d.cellStart[c + 1];
    const uint32_t n = e - s;
    float hm = 0.f;
    for (uint32_t a = s; a < e; ++a) {
        hm = fmaxf(hm, __double2float_ru(d.f[F_H][d.order[a]]));
    }
    d.cellHmax[c] = __float_as_uint(hm);
    if (n < 2) {
        return;
    }
    if (n <= (uint32_t)SORT_LOCAL) { // the usual case: insertion sort on a private copy of the keys
        double kx[SORT_LOCAL];
        uint32_t ks[SORT_LOCAL];
        for (uint32_t a = 0; a < n; ++a) {
            const uint32_t slot = d.order[s + a];
            const double x = d.f[F_X][slot];
            uint32_t b = a;
            while (b > 0 && (kx[b - 1] > x || (kx[b - 1] == x && ks[b - 1] > slot))) {
                kx[b] = kx[b - 1];
                ks[b] = ks[b - 1];
                --b;
            }
            kx[b] = x;
            ks[b] = slot;
        }
        for (uint32_t a = 0; a < n; ++a) {
            d.order[s + a] = ks[a];
        }
        return;
    }
    if (n > 4096) { // degenerate grid (one huge h): give up the order, the pair kernel then scans whole rows
        d.grid->unsorted = 1u;
        return;
    }
    for (uint32_t a = s + 1; a < e; ++a) {
        const uint32_t slot = d.order[a];
        const double x = d.f[F_X][slot];
        uint32_t b = a;
        while (b > s) {
            const uint32_t prev = d.order[b - 1];
            const double xp = d.f[F_X][prev];
            if (!(xp > x || (xp == x && prev > slot))) {
                break;
            }
            d.order[b] = prev;
            --b;
        }
        d.order[b] = slot;
    }
}

__global__ void __launch_bounds__(128) k_sort_cells(DevicePointers d, uint32_t maxCells) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= maxCells || d.listCtl->rebuild == 0u) {
        return;
    }
    const uint32_t ncells = d.grid->ncells;
    if (c > ncells) {
        d.cellHmax[c] = 0u;
        return;
    }
    if (c == ncells) { // the overflow cell of the large particles
        d.cellHmax[c] = 0u;
        d.grid->largeBegin = d.cellStart[c]; // first sorted index of the large particles
    } else {
        sortCell(d, c);
    }
    // slot -> sorted index (the slot's rank inside its cell is no longer needed): the prologue runs over the SLOTS, reads
    // the planes coalesced and scatters whole records
    const uint32_t s = d.cellStart[c], e = d.cellStart[c + 1];
    for (uint32_t a = s; a < e; ++a) {
        d.rank[d.order[a]] = a;
    }
}

int launchGridBuild(sphgpu_ctx* ctx) {
    const uint32_t n = ctx->nActive;
    cudaStream_t st = ctx->stream;
    const bool clampH = (ctx->prm.flags & SPHGPU_FLAG_ADAPTIVE_H) != 0;
    const bool force = ctx->listsDirty || !(ctx->listSkin > 0.);
    const bool track = !force; // (the first build has no cells to track)
    const uint32_t cellStride = ctx->maxCells + 1;
    if (track) { // neutral keys: lower bounds ~0, upper bounds 0; dispGlobal likewise, [6] = 0.f
        SPH_CUDA_CHECK(cudaMemsetAsync(ctx->d.dispA, 0xff, sizeof(uint32_t) * 3 * (size_t)cellStride, st));
        SPH_CUDA_CHECK(cudaMemsetAsync(ctx->d.dispA + 3 * (size_t)cellStride, 0, sizeof(uint32_t) * 3 * (size_t)cellStride, st));
        SPH_CUDA_CHECK(cudaMemsetAsync(ctx->d.dispGlobal, 0xff, sizeof(uint32_t) * 3, st));
        SPH_CUDA_CHECK(cudaMemsetAsync(ctx->d.dispGlobal + 3, 0, sizeof(uint32_t) * 5, st));
    }
    k_bounds<<<BOUNDS_BLOCKS, 256, 0, st>>>(ctx->d, n, clampH, ctx->prm.h_min, ctx->prm.h_max, ctx->prm.kernel_radius, track, cellStride);
    if (track) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        const uint32_t cb = (uint32_t)std::min<uint64_t>((uint64_t)sms * 2, ((uint64_t)ctx->maxCells / 8 + 255) / 256 + 1);
        k_disp_check<<<cb, 256, 0, st>>>(ctx->d, cellStride, track);
        ctx->launches += 1;
    }
    k_grid_decide<<<1, 256, 0, st>>>(ctx->d, BOUNDS_BLOCKS, n, force, ctx->listSkin, ctx->prm.kernel_radius);
    k_hmax_small<<<BOUNDS_BLOCKS, 256, 0, st>>>(ctx->d, n);
    k_grid_params<<<1, 32, 0, st>>>(ctx->d, BOUNDS_BLOCKS, ctx->prm.kernel_radius, ctx->maxCells, ctx->listSkin > 0. ? ctx->listSkin : 0.);
    ctx->listsDirty = false;
    SPH_CUDA_CHECK(cudaMemsetAsync(ctx->d.cellCount, 0, sizeof(uint32_t) * (ctx->maxCells + 1), st));
    const uint32_t blocks = (n + 255) / 256;
    if (blocks > 0) {
        k_cell_count<<<blocks, 256, 0, st>>>(ctx->d, n);
    }
    const uint32_t total = ctx->maxCells + 1;
    k_scan_block<<<ctx->scanBlocks, 512, 0, st>>>(ctx->d.cellCount, ctx->d.cellStart, ctx->d.scanBlock, total, ctx->d.listCtl);
    k_scan_sums<<<1, 1024, 0, st>>>(ctx->d.scanBlock, ctx->scanBlocks, ctx->d.listCtl);
    k_scan_add<<<ctx->scanBlocks, 512, 0, st>>>(ctx->d.cellStart, ctx->d.scanBlock, total, ctx->d.listCtl);
    if (blocks > 0) {
        k_scatter<<<blocks, 256, 0, st>>>(ctx->d, n);
    }
    k_sort_cells<<<(ctx->maxCells + 127) / 128, 128, 0, st>>>(ctx->d, ctx->maxCells);
    ctx->launches += 10;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

} // namespace sph
