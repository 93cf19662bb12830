// Tiled pair kernel: the production variant of the fused neighbour search + pair sums.
//
// Work unit  = up to 128 consecutive sorted targets of ONE cell row (same cy, cz), one target per thread.
// Candidates = the three cell rows of a z-slab (dy = -1,0,1) restricted to the x-range the unit's targets can reach,
//              staged in shared memory as 144-byte FP64 records (double2 moves, odd 16-byte stride => consecutive
//              records fall into different bank groups) plus an FP32 {x,y,z,h} copy relative to a unit-local origin.
// Phase 1    : every thread scans the candidates of ITS OWN 3 cells per row with a conservative FP32 distance test
//              (cheap, runs on the FP32/ALU pipes) and appends the survivors to a private list in shared memory.
// Phase 2    : every thread walks its list; the exact FP64 predicate (bit-identical neighbour sets) and the FP64 pair
//              arithmetic run with (nearly) full lanes -- no divergence on the expensive path.
// Per-target sums stay in registers; nothing is accumulated with atomics (asymmetric formulation).
// Staging uses the TMA bulk-copy engine: one thread issues cp.async.bulk (global -> shared, completion on an mbarrier)
// for each contiguous candidate row, the rest of the CTA only waits on the barrier's phase.
//
// Replaces the reference hot loop AsymmetricSolver.cpp:174-201 (finder.findAll + filter + kernel.grad +
// derivatives.eval) -- see pair.cu for the epilogue it shares with the direct variant.
#include "sphgpu_internal.h"

namespace sph {

constexpr int TILE_T = 128;   // targets (threads) per work unit
constexpr int TILE_C = 576;   // staged candidates per chunk
constexpr int LIST_CAP = 64;  // private list entries per round

constexpr int TILE_X = 20;    // cell-range entries cached in shared memory per candidate row

template <bool SOLID>
struct TileLayout {
    static constexpr int G = SOLID ? REC_SOLID : REC_FLUID; // record stride (doubles), same in global and shared memory:
    static constexpr int S = G;                              // 9 or 7 (odd) x 16 B => conflict-free consecutive records
    static constexpr size_t bytes = (size_t)TILE_C * S * 8 + (size_t)TILE_C * 16 + (size_t)LIST_CAP * TILE_T * 2;
};

// ---- TMA bulk copy + mbarrier (raw PTX; sm_90+ / sm_100a) -----------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCopyG2S(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smemAddr(bar)), "r"(parity)
                     : "memory");
    }
}
__device__ __forceinline__ void fenceProxyAsync() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- work list: segments of <= 128 targets per cell row ------------------------------------------------------
__global__ void __launch_bounds__(256) k_row_segments(DevicePointers d, uint32_t maxCells) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > maxCells) {
        return;
    }
    const GridDev g = *d.grid;
    const uint32_t rows = (uint32_t)(g.dim[1] * g.dim[2]);
    uint32_t nseg = 0;
    if (r < rows) {
        const uint32_t cnt = d.cellStart[(r + 1) * (uint32_t)g.dim[0]] - d.cellStart[r * (uint32_t)g.dim[0]];
        nseg = (cnt + TILE_T - 1) / TILE_T;
    }
    d.cellCount[r] = nseg; // cellCount is free once cellStart has been built
}

__global__ void __launch_bounds__(256) k_fill_segments(DevicePointers d, uint32_t maxCells) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= maxCells) {
        return;
    }
    const uint32_t s = d.segStart[r], e = d.segStart[r + 1];
    for (uint32_t k = s; k < e; ++k) {
        d.segRow[k] = r;
    }
}

struct ChunkState {
    uint32_t beg[3], end[3], base[3]; // per dy row: global sorted range staged in this chunk and its smem offset
    uint32_t used;
    int z;                            // slab (absolute cell z) of this chunk
    uint32_t cells[3][TILE_X + 2];    // cellStart[row base + x0 ...] of the three candidate rows (x0 .. x1+1)
};

struct ChunkCursor { // iteration state of the chunk builder (thread 0 only)
    int dz, dy;
    uint32_t pos;
    int posValid;
};

/// Next chunk of the unit: as many whole / partial candidate rows of the current z-slab as fit into TILE_C records.
__device__ __forceinline__ void nextChunk(const DevicePointers& d, ChunkCursor& cur, ChunkState& cs, int cy, int cz, int x0, int x1,
    int dimx, int dimy, int dimz) {
    uint32_t used = 0;
    for (int r = 0; r < 3; ++r) {
        cs.beg[r] = cs.end[r] = cs.base[r] = 0;
    }
    int z = 0;
    while (cur.dz <= 1) {
        z = cz + cur.dz;
        if (z < 0 || z >= dimz) {
            cur.dz++;
            cur.dy = 0;
            cur.posValid = 0;
            continue;
        }
        bool full = false;
        while (cur.dy < 3) {
            const int y = cy + cur.dy - 1;
            if (y < 0 || y >= dimy) {
                cur.dy++;
                cur.posValid = 0;
                continue;
            }
            const uint32_t rb = (uint32_t)((z * dimy + y) * dimx);
            const uint32_t rowEnd = d.cellStart[rb + x1 + 1];
            if (!cur.posValid) {
                cur.pos = d.cellStart[rb + x0];
                cur.posValid = 1;
            }
            const uint32_t take = min(rowEnd - cur.pos, (uint32_t)TILE_C - used);
            if (take > 0) {
                cs.beg[cur.dy] = cur.pos;
                cs.end[cur.dy] = cur.pos + take;
                cs.base[cur.dy] = used;
                used += take;
                cur.pos += take;
            }
            if (cur.pos >= rowEnd) {
                cur.dy++;
                cur.posValid = 0;
            } else {
                full = true;
                break;
            }
        }
        if (full) {
            break;
        }
        cur.dz++; // slab finished; a chunk never mixes slabs
        cur.dy = 0;
        cur.posValid = 0;
        if (used > 0) {
            break;
        }
    }
    cs.used = used;
    cs.z = z;
    if (used > 0 && x1 - x0 + 2 <= TILE_X + 2) {
        for (int r = 0; r < 3; ++r) {
            const int y = cy + r - 1;
            if (y < 0 || y >= dimy) {
                continue;
            }
            const uint32_t rb = (uint32_t)((z * dimy + y) * dimx);
            for (int c = x0; c <= x1 + 1; ++c) {
                cs.cells[r][c - x0] = d.cellStart[rb + c];
            }
        }
    }
}

template <bool SOLID, bool CORRECTED, bool FILTER>
__global__ void __launch_bounds__(TILE_T, 2) k_pair_tiled(DevicePointers d, uint32_t nOwned, uint32_t maxCells) {
    using L = TileLayout<SOLID>;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    double* recS = reinterpret_cast<double*>(smemRaw);
    float4* f4 = reinterpret_cast<float4*>(recS + (size_t)TILE_C * L::S);
    uint16_t* list = reinterpret_cast<uint16_t*>(f4 + TILE_C);
    __shared__ ChunkState csBuf[2];
    __shared__ float sKey[TILE_T];
    __shared__ uint16_t sPerm[TILE_T];
    __shared__ __align__(8) uint64_t stageBar;

    const GridDev g = *d.grid;
    const int dimx = g.dim[0], dimy = g.dim[1], dimz = g.dim[2];
    const uint32_t totalSegs = d.segStart[maxCells];
    const int tid = threadIdx.x;
    const float Rhalf = (float)(0.5 * c_prm.kernel_radius * (1. + 2.e-5));
    ChunkCursor cur;
    uint32_t stagePhase = 0;
    if (tid == 0) {
        mbarInit(&stageBar, 1);
    }
    __syncthreads();

    for (uint32_t unit = blockIdx.x; unit < totalSegs; unit += gridDim.x) {
        const uint32_t row = d.segRow[unit];
        const uint32_t seg = unit - d.segStart[row];
        const int cy = (int)(row % (uint32_t)dimy), cz = (int)(row / (uint32_t)dimy);
        const uint32_t rowBase = row * (uint32_t)dimx;
        const uint32_t tBeg = d.cellStart[rowBase] + seg * TILE_T;
        const uint32_t tEnd = min(tBeg + (uint32_t)TILE_T, d.cellStart[rowBase + dimx]);
        // x-range of cells the unit's targets occupy, and the unit-local origin of the FP32 copies
        const int cxA = (int)(d.sCell[tBeg] - rowBase), cxB = (int)(d.sCell[tEnd - 1] - rowBase);
        const int x0 = max(cxA - 1, 0), x1 = min(cxB + 1, dimx - 1);
        const double2 oxy = reinterpret_cast<const double2*>(d.rec + (size_t)tBeg * L::G)[0];
        const double oz = d.rec[(size_t)tBeg * L::G + R_Z];
        // absolute error bound of the FP32 relative coordinates (2^-24 * extent per coordinate) with a wide safety factor
        const float slack = (float)((double)(x1 - x0 + 2) * g.cell * 1.e-6);

        // ---- lane assignment: order the unit's targets by z so that the lanes of a warp see similar numbers of
        // neighbours in every z-slab (balanced private lists => full lanes in phase 2)
        {
            const uint32_t t0 = tBeg + tid;
            sKey[tid] = (t0 < tEnd) ? (float)(d.rec[(size_t)t0 * L::G + R_Z] - oz) : 3.0e38f;
        }
        __syncthreads();
        {
            const float key = sKey[tid];
            int rank = 0;
            for (int j = 0; j < TILE_T; ++j) {
                const float kj = sKey[j];
                rank += (kj < key || (kj == key && j < tid)) ? 1 : 0;
            }
            sPerm[rank] = (uint16_t)tid;
        }
        __syncthreads();
        const uint32_t t = tBeg + sPerm[tid];
        const bool live = t < tEnd;
        const uint32_t slot = live ? d.order[t] : 0xffffffffu;
        const bool target = live && slot < nOwned; // ghosts are neighbours only

        Particle pi;
        int cx = 0;
        float fxi = 0.f, fyi = 0.f, fzi = 0.f, fhi = 0.f;
        double reach = 0.;
        if (live) {
            loadRecord<SOLID>(d.rec + (size_t)t * L::G, pi);
            cx = (int)(d.sCell[t] - rowBase);
            fxi = (float)(pi.x - oxy.x);
            fyi = (float)(pi.y - oxy.y);
            fzi = (float)(pi.z - oz);
            fhi = (float)pi.h;
            reach = 0.5 * c_prm.kernel_radius * (pi.h + g.hmax) * (1. + 1.e-9); // >= R * hbar for every neighbour
        }
        // distance of the target to the faces of its own cell in y and z (lower bounds, shrunk for rounding safety)
        const double tol = 1.e-9 * g.cell;
        const double yLo = fmax(pi.y - (g.lo[1] + cy * g.cell) - tol, 0.), yHi = fmax((g.lo[1] + (cy + 1) * g.cell) - pi.y - tol, 0.);
        const double zLo = fmax(pi.z - (g.lo[2] + cz * g.cell) - tol, 0.), zHi = fmax((g.lo[2] + (cz + 1) * g.cell) - pi.z - tol, 0.);
        Accum acc;
        accumZero(acc);

        if (tid == 0) {
            cur.dz = -1;
            cur.dy = 0;
            cur.posValid = 0;
            nextChunk(d, cur, csBuf[0], cy, cz, x0, x1, dimx, dimy, dimz);
        }
        int buf = 0;
        while (true) {
            __syncthreads(); // chunk descriptor published; everyone is done with the previous chunk's shared memory
            const ChunkState& cs = csBuf[buf];
            if (cs.used == 0) {
                break;
            }
            // ---- stage the chunk: TMA bulk copies of the FP64 records (one per candidate row), then the FP32 copies ----
            if (tid == 0) {
                fenceProxyAsync(); // the buffer was last read through the generic proxy
                mbarExpectTx(&stageBar, cs.used * (uint32_t)(L::G * 8));
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const uint32_t n = cs.end[r] - cs.beg[r];
                    if (n > 0) {
                        bulkCopyG2S(recS + (size_t)cs.base[r] * L::S, d.rec + (size_t)cs.beg[r] * L::G, n * (uint32_t)(L::G * 8), &stageBar);
                    }
                }
                nextChunk(d, cur, csBuf[buf ^ 1], cy, cz, x0, x1, dimx, dimy, dimz); // overlaps with the copies
            }
            mbarWait(&stageBar, stagePhase);
            stagePhase ^= 1;
            for (uint32_t c = tid; c < cs.used; c += TILE_T) {
                const double2* src = reinterpret_cast<const double2*>(recS + (size_t)c * L::S);
                const double2 a0 = src[0], a1 = src[1];
                f4[c] = make_float4((float)(a0.x - oxy.x), (float)(a0.y - oxy.y), (float)(a1.x - oz), (float)a1.y);
            }
            __syncthreads();
            buf ^= 1;
            if (!target) {
                continue;
            }
            // ---- private rounds: phase 1 (FP32 filter -> list), phase 2 (FP64 pairs) ----
            const int z = cs.z;
            const double dzMin = (z == cz) ? 0. : (z < cz ? zLo : zHi);
            int r = 0;
            uint32_t gpos = 0, ghi = 0, gbase = 0;
            bool open = false;
            while (r < 3) {
                int cnt = 0;
                while (r < 3) {
                    if (!open) {
                        const uint32_t b = cs.beg[r], e = cs.end[r];
                        gpos = ghi = 0;
                        if (e > b) {
                            // x-interval of this candidate row the target can reach => cell sub-range [c0, c1]
                            const double dyMin = (r == 1) ? 0. : (r == 0 ? yLo : yHi);
                            const double rem = reach * reach - dyMin * dyMin - dzMin * dzMin;
                            if (rem > 0.) {
                                const double ext = sqrt(rem) * (1. + 1.e-9);
                                int c0 = (int)floor((pi.x - ext - g.lo[0]) * g.cellInv);
                                int c1 = (int)floor((pi.x + ext - g.lo[0]) * g.cellInv);
                                c0 = max(max(c0, cx - 1), 0);
                                c1 = min(min(c1, cx + 1), dimx - 1);
                                if (c0 <= c1) {
                                    if (x1 - x0 + 2 <= TILE_X + 2) {
                                        gpos = max(cs.cells[r][c0 - x0], b);
                                        ghi = min(cs.cells[r][c1 + 1 - x0], e);
                                    } else {
                                        const uint32_t rb = (uint32_t)((z * dimy + cy + r - 1) * dimx);
                                        gpos = max(d.cellStart[rb + c0], b);
                                        ghi = min(d.cellStart[rb + c1 + 1], e);
                                    }
                                    gbase = cs.base[r] - b; // smem index = gbase + global index (mod 2^32)
                                }
                            }
                        }
                        open = true;
                    }
                    // four candidates per trip: the loads are independent, only the list append is serial. The target
                    // itself is not excluded here (672 compares) but masked in phase 2 (68 compares).
                    {
                        uint16_t* lp = list + cnt * TILE_T + tid;
#define SPH_F32_TEST(C, K)                                                                                            \
    {                                                                                                                 \
        const float ddx = fxi - C.x, ddy = fyi - C.y, ddz = fzi - C.z;                                                \
        const float dd2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));                                                  \
        const float lim = fmaf(Rhalf, fhi + C.w, slack);                                                              \
        if (dd2 <= lim * lim) {                                                                                       \
            *lp = (uint16_t)(K);                                                                                      \
            lp += TILE_T;                                                                                             \
            cnt++;                                                                                                    \
        }                                                                                                             \
    }
                        while (gpos + 4 <= ghi && cnt + 4 <= LIST_CAP) {
                            const uint32_t k0 = gbase + gpos;
                            const float4 ca = f4[k0], cb = f4[k0 + 1], cc = f4[k0 + 2], cd = f4[k0 + 3];
                            SPH_F32_TEST(ca, k0)
                            SPH_F32_TEST(cb, k0 + 1)
                            SPH_F32_TEST(cc, k0 + 2)
                            SPH_F32_TEST(cd, k0 + 3)
                            gpos += 4;
                        }
                        while (gpos < ghi && ghi - gpos < 4 && cnt < LIST_CAP) {
                            const uint32_t k0 = gbase + gpos;
                            const float4 ca = f4[k0];
                            SPH_F32_TEST(ca, k0)
                            gpos++;
                        }
#undef SPH_F32_TEST
                    }
                    if (gpos >= ghi) {
                        r++;
                        open = false;
                    } else {
                        break; // list (nearly) full: drain it, then resume the scan
                    }
                }
                // phase 2: two list entries per trip so the long per-pair chains overlap
                const double* self = recS + ((t >= cs.beg[1] && t < cs.end[1]) ? (size_t)(cs.base[1] + (t - cs.beg[1])) * L::S : (size_t)TILE_C * L::S);
                int q = 0;
                for (; q + 1 < cnt; q += 2) {
                    const double* rp0 = recS + (size_t)list[q * TILE_T + tid] * L::S;
                    const double* rp1 = recS + (size_t)list[(q + 1) * TILE_T + tid] * L::S;
                    Particle pj0, pj1;
                    loadRecord<SOLID>(rp0, pj0);
                    loadRecord<SOLID>(rp1, pj1);
                    const double dx0 = pi.x - pj0.x, dy0 = pi.y - pj0.y, dz0 = pi.z - pj0.z;
                    const double dx1 = pi.x - pj1.x, dy1 = pi.y - pj1.y, dz1 = pi.z - pj1.z;
                    double d20, hb0, d21, hb1;
                    const bool v0 = isNeighbour(dx0, dy0, dz0, pi.h, pj0.h, c_prm.kernel_radius, d20, hb0) && rp0 != self;
                    const bool v1 = isNeighbour(dx1, dy1, dz1, pi.h, pj1.h, c_prm.kernel_radius, d21, hb1) && rp1 != self;
                    pairAccumulateMasked<SOLID, CORRECTED, FILTER>(c_prm, d.lut, pi, pj0, dx0, dy0, dz0, d20, hb0, v0, acc);
                    pairAccumulateMasked<SOLID, CORRECTED, FILTER>(c_prm, d.lut, pi, pj1, dx1, dy1, dz1, d21, hb1, v1, acc);
                }
                if (q < cnt) {
                    const double* rp0 = recS + (size_t)list[q * TILE_T + tid] * L::S;
                    Particle pj0;
                    loadRecord<SOLID>(rp0, pj0);
                    const double dx0 = pi.x - pj0.x, dy0 = pi.y - pj0.y, dz0 = pi.z - pj0.z;
                    double d20, hb0;
                    const bool v0 = isNeighbour(dx0, dy0, dz0, pi.h, pj0.h, c_prm.kernel_radius, d20, hb0) && rp0 != self;
                    pairAccumulateMasked<SOLID, CORRECTED, FILTER>(c_prm, d.lut, pi, pj0, dx0, dy0, dz0, d20, hb0, v0, acc);
                }
            }
        }
        // ---- epilogue: finalizers + stores (shared with the direct variant) ----
        if (target) {
            const MaterialDev& mat = c_mats[d.u[U_MATID][slot]];
            double S[5] = { 0., 0., 0., 0., 0. };
            if (SOLID) {
                for (int k = 0; k < 5; ++k) {
                    S[k] = d.f[F_S0 + k][slot];
                }
            }
            Derivs out;
            finalizeParticle<SOLID, CORRECTED>(c_prm, mat, acc, pi.h, pi.rho, d.f[F_P][slot], pi.cs, SOLID ? d.f[F_REDUCE][slot] : 1., S, out);
            storeDerivs<SOLID, CORRECTED>(d, slot, out);
        }
        neighbourStats(d, acc.cnt, target);
    }
}

int launchSegments(sphgpu_ctx* ctx) {
    cudaStream_t st = ctx->stream;
    const uint32_t total = ctx->maxCells + 1;
    k_row_segments<<<(total + 255) / 256, 256, 0, st>>>(ctx->d, ctx->maxCells);
    k_scan_block<<<ctx->scanBlocks, 512, 0, st>>>(ctx->d.cellCount, ctx->d.segStart, ctx->d.scanBlock, total);
    k_scan_sums<<<1, 1024, 0, st>>>(ctx->d.scanBlock, ctx->scanBlocks);
    k_scan_add<<<ctx->scanBlocks, 512, 0, st>>>(ctx->d.segStart, ctx->d.scanBlock, total);
    k_fill_segments<<<(ctx->maxCells + 255) / 256, 256, 0, st>>>(ctx->d, ctx->maxCells);
    ctx->launches += 5;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

template <bool SOLID, bool CORRECTED, bool FILTER>
static int launchTiledVariant(sphgpu_ctx* ctx) {
    auto kernel = k_pair_tiled<SOLID, CORRECTED, FILTER>;
    static bool configured = false; // per instantiation; the attribute is per device function
    const size_t smem = TileLayout<SOLID>::bytes;
    if (!configured) {
        SPH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const uint32_t upper = ctx->nActive / TILE_T + ctx->maxCells + 1; // >= number of segments
    const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)sms * 2 * 8, std::max<uint32_t>(upper, 1u));
    kernel<<<grid, TILE_T, smem, ctx->stream>>>(ctx->d, ctx->n, ctx->maxCells);
    ctx->launches += 1;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchPairTiled(sphgpu_ctx* ctx) {
    int rc = launchSegments(ctx);
    if (rc != SPHGPU_OK) {
        return rc;
    }
    if (!ctx->solid) {
        return launchTiledVariant<false, false, false>(ctx);
    }
    if (ctx->corrected) {
        return ctx->filter ? launchTiledVariant<true, true, true>(ctx) : launchTiledVariant<true, true, false>(ctx);
    }
    return ctx->filter ? launchTiledVariant<true, false, true>(ctx) : launchTiledVariant<true, false, false>(ctx);
}

} // namespace sph
