// Tiled pair kernel: the production variant of the fused neighbour search + pair sums.
//
// Grid       : cells are a x a x a/2 (a = R h_max); a "double row" is the two half-height cell rows (cy, 2k), (cy, 2k+1).
// Work unit  = up to 128 targets of ONE double row restricted to a cell range [cA, cB] in x, one target per thread,
//              per-target sums in registers, no atomics (asymmetric formulation).
// Candidates = the six z-layers 2k-2 .. 2k+3 (three cell rows dy = -1,0,1 each) restricted to cells [cA-1, cB+1]. They
//              are processed as three CHUNKS that pair layers symmetrically: far (2k-2, 2k+3), near (2k-1, 2k+2),
//              centre (2k, 2k+1). With the lanes ordered by z, every warp then sees about the same number of
//              neighbours in every chunk (measured on the hex lattice: 5/19/45 per chunk for all four z-bands, versus
//              25/43/0 ... 0/43/25 when whole layers are processed one after the other), so the CTA-wide barrier at the
//              chunk boundaries costs little.
// Staging    : one thread issues TMA bulk copies (cp.async.bulk global -> shared, completion on an mbarrier), one per
//              contiguous candidate row; records are 144-byte FP64 structures (odd 16-byte stride => consecutive
//              records fall into different bank groups); an FP32 {x,y,z,h} copy relative to a unit-local origin is
//              derived from them.
// Phase 1    : every thread scans the candidates of the cells IT can reach (row-wise interval culling in y, z and x)
//              with a conservative FP32 distance test (FP32/ALU pipes) and appends survivors to a private u16 list.
// Phase 2    : every thread walks its list two entries at a time; the exact FP64 predicate (bit-identical neighbour
//              sets) enters the branch-free FP64 pair body as a mask -- full lanes, no divergence on the expensive path.
//
// Replaces the reference hot loop AsymmetricSolver.cpp:174-201 (finder.findAll + filter + kernel.grad +
// derivatives.eval) -- see pair.cu for the epilogue it shares with the direct variant.
#include "sphgpu_internal.h"

namespace sph {

constexpr int TILE_T = 128;   // targets (threads) per work unit
constexpr int TILE_C = 576;   // staged candidates per chunk
constexpr int LIST_CAP = 64;  // private list entries per round
constexpr int TILE_X = 20;    // cell-range entries cached in shared memory per candidate row
constexpr int CHUNK_ROWS = 6; // candidate rows per chunk: 2 z-layers x 3 y-rows

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

// ---- work list: units of <= 128 targets per double row (next-fit packing of cell columns) ----------------------
// One thread per double row walks its cells in x; FILL = false counts the units, FILL = true writes the descriptors
// {double row, cA, skip, cB}: the unit takes targets [skip, skip + 128) of the concatenation
// (lower row cells cA..cB) ++ (upper row cells cA..cB).
template <bool FILL>
__global__ void __launch_bounds__(128) k_units(DevicePointers d, uint32_t maxCells) {
    const uint32_t dr = blockIdx.x * blockDim.x + threadIdx.x;
    if (dr > maxCells || (FILL && dr >= maxCells)) {
        return;
    }
    const GridDev g = *d.grid;
    const int dimx = g.dim[0], dimy = g.dim[1], dimz = g.dim[2];
    const uint32_t doubleRows = (uint32_t)dimy * (uint32_t)((dimz + 1) / 2);
    if (dr >= doubleRows) {
        if (!FILL) {
            d.cellCount[dr] = 0; // cellCount is free once cellStart has been built
        }
        return;
    }
    const int cy = (int)(dr % (uint32_t)dimy), k = (int)(dr / (uint32_t)dimy);
    const uint32_t rbL = (uint32_t)(((2 * k) * dimy + cy) * dimx);
    const bool hasU = 2 * k + 1 < dimz;
    const uint32_t rbU = hasU ? (uint32_t)(((2 * k + 1) * dimy + cy) * dimx) : 0u;
    uint32_t units = 0;
    const uint32_t out = FILL ? d.segStart[dr] : 0u;
    int cA = 0, cLast = 0;
    uint32_t acc = 0;
    auto emit = [&](int a, int b, uint32_t total) {
        const uint32_t parts = (total + TILE_T - 1) / TILE_T;
        if (FILL) {
            for (uint32_t p = 0; p < parts; ++p) {
                d.unitDesc[out + units + p] = make_uint4(dr, (uint32_t)a, p * TILE_T, (uint32_t)b);
            }
        }
        units += parts;
    };
    for (int c = 0; c < dimx; ++c) {
        uint32_t cnt = d.cellStart[rbL + c + 1] - d.cellStart[rbL + c];
        if (hasU) {
            cnt += d.cellStart[rbU + c + 1] - d.cellStart[rbU + c];
        }
        if (cnt == 0) {
            continue;
        }
        // close the open unit when the column does not fit or the x-range would outgrow the cached cell table
        if (acc > 0 && (acc + cnt > (uint32_t)TILE_T || c - cA + 3 > TILE_X)) {
            emit(cA, cLast, acc);
            acc = 0;
        }
        if (acc == 0) {
            cA = c;
        }
        acc += cnt;
        cLast = c;
        if (acc >= (uint32_t)TILE_T) { // a single column with more than 128 targets is split by `skip`
            emit(cA, cLast, acc);
            acc = 0;
        }
    }
    if (acc > 0) {
        emit(cA, cLast, acc);
    }
    if (!FILL) {
        d.cellCount[dr] = units;
    }
}

struct ChunkState {
    uint32_t beg[CHUNK_ROWS], end[CHUNK_ROWS], base[CHUNK_ROWS]; // staged global range per candidate row + smem offset
    uint32_t used;
    int chunk;                                  // 0 far, 1 near, 2 centre
    uint32_t cells[CHUNK_ROWS][TILE_X + 2];     // cellStart[row base + x0 ...] of the candidate rows (x0 .. x1+1)
};

struct ChunkCursor { // iteration state of the chunk builder (thread 0 only)
    int chunk, row;
    uint32_t pos;
    int posValid;
};

/// z-layer (absolute half-height cell index) of candidate row `r` (0..5) of chunk `c` for the double row k.
__device__ __forceinline__ int chunkLayer(int c, int r, int k) {
    const int lo = (c == 0) ? 2 * k - 2 : (c == 1 ? 2 * k - 1 : 2 * k);
    const int hi = (c == 0) ? 2 * k + 3 : (c == 1 ? 2 * k + 2 : 2 * k + 1);
    return r < 3 ? lo : hi;
}

/// Next chunk of the unit: as many whole / partial candidate rows of the current layer pair as fit into TILE_C records.
/// rowBeg/rowEnd hold the global sorted ranges of the unit's 18 candidate rows (3 chunks x 6 rows; empty if outside).
__device__ __forceinline__ void nextChunk(const uint32_t* rowBeg, const uint32_t* rowEnd, ChunkCursor& cur, ChunkState& cs) {
    uint32_t used = 0;
    for (int r = 0; r < CHUNK_ROWS; ++r) {
        cs.beg[r] = cs.end[r] = cs.base[r] = 0;
    }
    int chunk = 0;
    while (cur.chunk < 3) {
        chunk = cur.chunk;
        bool full = false;
        while (cur.row < CHUNK_ROWS) {
            const uint32_t rb = rowBeg[chunk * CHUNK_ROWS + cur.row], re = rowEnd[chunk * CHUNK_ROWS + cur.row];
            if (!cur.posValid) {
                cur.pos = rb;
                cur.posValid = 1;
            }
            const uint32_t take = re > cur.pos ? min(re - cur.pos, (uint32_t)TILE_C - used) : 0u;
            if (take > 0) {
                cs.beg[cur.row] = cur.pos;
                cs.end[cur.row] = cur.pos + take;
                cs.base[cur.row] = used;
                used += take;
                cur.pos += take;
            }
            if (cur.pos >= re) {
                cur.row++;
                cur.posValid = 0;
            } else {
                full = true;
                break;
            }
        }
        if (full) {
            break;
        }
        cur.chunk++; // layer pair finished; a chunk never mixes layer pairs
        cur.row = 0;
        cur.posValid = 0;
        if (used > 0) {
            break;
        }
    }
    cs.used = used;
    cs.chunk = chunk;
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
    __shared__ uint32_t sRowBeg[3 * CHUNK_ROWS], sRowEnd[3 * CHUNK_ROWS];

    const GridDev g = *d.grid;
    const int dimx = g.dim[0], dimy = g.dim[1], dimz = g.dim[2];
    const uint32_t totalUnits = d.segStart[maxCells];
    const int tid = threadIdx.x;
    const float Rhalf = (float)(0.5 * c_prm.kernel_radius * (1. + 2.e-5));
    ChunkCursor cur;
    uint32_t stagePhase = 0;
    if (tid == 0) {
        mbarInit(&stageBar, 1);
    }
    __syncthreads();

    for (uint32_t unit = blockIdx.x; unit < totalUnits; unit += gridDim.x) {
        const uint4 desc = d.unitDesc[unit];
        const uint32_t dr = desc.x;
        const int cA = (int)desc.y, cB = (int)desc.w;
        const int cy = (int)(dr % (uint32_t)dimy), k = (int)(dr / (uint32_t)dimy);
        const uint32_t rbL = (uint32_t)(((2 * k) * dimy + cy) * dimx);
        const bool hasU = 2 * k + 1 < dimz;
        const uint32_t rbU = hasU ? (uint32_t)(((2 * k + 1) * dimy + cy) * dimx) : 0u;
        const uint32_t lBeg = d.cellStart[rbL + cA], lCnt = d.cellStart[rbL + cB + 1] - lBeg;
        const uint32_t uBeg = hasU ? d.cellStart[rbU + cA] : 0u, uCnt = hasU ? d.cellStart[rbU + cB + 1] - uBeg : 0u;
        const uint32_t total = lCnt + uCnt, skip = desc.z;
        const uint32_t nLive = min((uint32_t)TILE_T, total - skip);
        const int x0 = max(cA - 1, 0), x1 = min(cB + 1, dimx - 1);
        // unit-local origin of the FP32 copies: corner of the unit's first cell
        const double ox = g.lo[0] + cA * g.cell, oy = g.lo[1] + cy * g.cell, oz = g.lo[2] + (2 * k) * g.cellZ;
        // absolute error bound of the FP32 relative coordinates (2^-24 * extent per coordinate) with a wide safety factor
        const float slack = (float)((double)(x1 - x0 + 3) * g.cell * 1.e-6);
        const float cellF = (float)g.cell, cellZF = (float)g.cellZ, cellInvF = (float)g.cellInv;
        const float guard = (float)((double)(x1 - x0 + 3) * g.cell * 1.e-5);

        // ---- lane assignment: order the unit's targets by z so that the lanes of a warp see similar numbers of
        // neighbours in every chunk (balanced private lists => full lanes in phase 2)
        auto targetIndex = [&](uint32_t j) { return j < lCnt ? lBeg + j : uBeg + (j - lCnt); };
        sKey[tid] = ((uint32_t)tid < nLive) ? (float)(d.rec[(size_t)targetIndex(skip + tid) * L::G + R_Z] - oz) : 3.0e38f;
        if (tid < 3 * CHUNK_ROWS) { // global ranges of the unit's 18 candidate rows
            const int z = chunkLayer(tid / CHUNK_ROWS, tid % CHUNK_ROWS, k), y = cy + (tid % 3) - 1;
            uint32_t rb = 0, re = 0;
            if (z >= 0 && z < dimz && y >= 0 && y < dimy) {
                const uint32_t base = (uint32_t)((z * dimy + y) * dimx);
                rb = d.cellStart[base + x0];
                re = d.cellStart[base + x1 + 1];
            }
            sRowBeg[tid] = rb;
            sRowEnd[tid] = re;
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
        const bool live = (uint32_t)tid < nLive;
        const uint32_t jOwn = skip + sPerm[tid];
        const uint32_t t = live ? targetIndex(jOwn) : 0u;
        const bool upper = live && jOwn >= lCnt;
        const uint32_t slot = live ? d.order[t] : 0xffffffffu;
        const bool target = live && slot < nOwned; // ghosts are neighbours only
        if (!__syncthreads_or(target ? 1 : 0)) {
            continue; // a unit made of ghost particles only (halo band of a decomposed run)
        }

        Particle pi;
        int cx = 0;
        float fxi = 0.f, fyi = 0.f, fzi = 0.f, fhi = 0.f;
        float reachF = 0.f;
        if (live) {
            loadRecord<SOLID>(d.rec + (size_t)t * L::G, pi);
            cx = (int)(d.sCell[t] - (upper ? rbU : rbL));
            fxi = (float)(pi.x - ox);
            fyi = (float)(pi.y - oy);
            fzi = (float)(pi.z - oz);
            fhi = (float)pi.h;
            reachF = (float)(0.5 * c_prm.kernel_radius * (pi.h + g.hmax) * (1. + 1.e-5)); // >= R * hbar for every neighbour
        } else {
            pi.x = pi.y = pi.z = 0.;
            pi.h = 1.;
        }
        Accum acc;
        accumZero(acc);

        if (tid == 0) {
            cur.chunk = 0;
            cur.row = 0;
            cur.posValid = 0;
            nextChunk(sRowBeg, sRowEnd, cur, csBuf[0]);
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
                for (int r = 0; r < CHUNK_ROWS; ++r) {
                    const uint32_t n = cs.end[r] - cs.beg[r];
                    if (n > 0) {
                        bulkCopyG2S(recS + (size_t)cs.base[r] * L::S, d.rec + (size_t)cs.beg[r] * L::G, n * (uint32_t)(L::G * 8), &stageBar);
                    }
                }
                nextChunk(sRowBeg, sRowEnd, cur, csBuf[buf ^ 1]); // overlaps with the copies
            }
            // cell ranges of the candidate rows, one entry per thread (issued before waiting for the copies)
            ChunkState& csw = csBuf[buf];
            for (int e = tid; e < CHUNK_ROWS * (TILE_X + 2); e += TILE_T) {
                const int r = e / (TILE_X + 2), c = x0 + e % (TILE_X + 2);
                const int z = chunkLayer(cs.chunk, r, k), y = cy + (r % 3) - 1;
                if (c <= x1 + 1 && z >= 0 && z < dimz && y >= 0 && y < dimy) {
                    csw.cells[r][c - x0] = d.cellStart[(uint32_t)((z * dimy + y) * dimx) + c];
                }
            }
            mbarWait(&stageBar, stagePhase);
            stagePhase ^= 1;
            for (uint32_t c = tid; c < cs.used; c += TILE_T) {
                const double2* src = reinterpret_cast<const double2*>(recS + (size_t)c * L::S);
                const double2 a0 = src[0], a1 = src[1];
                f4[c] = make_float4((float)(a0.x - ox), (float)(a0.y - oy), (float)(a1.x - oz), (float)a1.y);
            }
            __syncthreads();
            buf ^= 1;
            if (!target) {
                continue;
            }
            // ---- private rounds: phase 1 (FP32 filter -> list), phase 2 (FP64 pairs) ----
            const int chunk = cs.chunk;
            // the target's own record, if it is staged in this chunk (centre chunk, own layer, dy = 0)
            const int selfRow = upper ? 4 : 1;
            const double* self = recS + ((chunk == 2 && t >= cs.beg[selfRow] && t < cs.end[selfRow])
                                                ? (size_t)(cs.base[selfRow] + (t - cs.beg[selfRow])) * L::S
                                                : (size_t)TILE_C * L::S);
            int r = 0;
            uint32_t kpos = 0, khi = 0;
            bool open = false;
            while (r < CHUNK_ROWS) {
                int cnt = 0;
                while (r < CHUNK_ROWS) {
                    if (!open) {
                        const uint32_t b = cs.beg[r], e = cs.end[r];
                        kpos = khi = 0;
                        if (e > b) {
                            // intervals of this candidate row in y and z, and the x-interval the target can reach in it:
                            // FP32 relative to the unit origin with a guard band (cells were assigned in FP64)
                            const int zrel = chunkLayer(chunk, r, 0), yrel = (r % 3) - 1; // relative to (2k, cy)
                            const float yl = (float)yrel * cellF, zl = (float)zrel * cellZF;
                            const float dyMin = fmaxf(fmaxf(yl - fyi, fyi - (yl + cellF)) - guard, 0.f);
                            const float dzMin = fmaxf(fmaxf(zl - fzi, fzi - (zl + cellZF)) - guard, 0.f);
                            const float rem = reachF * reachF - dyMin * dyMin - dzMin * dzMin;
                            if (rem > 0.f) {
                                const float ext = sqrtf(rem) * (1.f + 1.e-5f) + guard;
                                int c0 = cA + (int)floorf((fxi - ext) * cellInvF);
                                int c1 = cA + (int)floorf((fxi + ext) * cellInvF);
                                c0 = max(max(c0, cx - 1), x0);
                                c1 = min(min(c1, cx + 1), x1);
                                if (c0 <= c1) {
                                    // smem index = base + (global index - beg); the staged piece may be a part of the row
                                    const uint32_t glo = max(cs.cells[r][c0 - x0], b), ghi = min(cs.cells[r][c1 + 1 - x0], e);
                                    if (ghi > glo) {
                                        kpos = cs.base[r] + (glo - b);
                                        khi = cs.base[r] + (ghi - b);
                                    }
                                }
                            }
                        }
                        open = true;
                    }
                    // four candidates per trip: the loads are independent, only the list append is serial. The target
                    // itself is not excluded here (~600 compares) but masked in phase 2 (~70 compares).
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
                        while (kpos + 8 <= khi && cnt + 8 <= LIST_CAP) {
                            const float4 ca = f4[kpos], cb = f4[kpos + 1], cc = f4[kpos + 2], cd = f4[kpos + 3];
                            const float4 ce = f4[kpos + 4], cf = f4[kpos + 5], cg = f4[kpos + 6], ch = f4[kpos + 7];
                            SPH_F32_TEST(ca, kpos)
                            SPH_F32_TEST(cb, kpos + 1)
                            SPH_F32_TEST(cc, kpos + 2)
                            SPH_F32_TEST(cd, kpos + 3)
                            SPH_F32_TEST(ce, kpos + 4)
                            SPH_F32_TEST(cf, kpos + 5)
                            SPH_F32_TEST(cg, kpos + 6)
                            SPH_F32_TEST(ch, kpos + 7)
                            kpos += 8;
                        }
                        while (kpos + 4 <= khi && cnt + 4 <= LIST_CAP) {
                            const float4 ca = f4[kpos], cb = f4[kpos + 1], cc = f4[kpos + 2], cd = f4[kpos + 3];
                            SPH_F32_TEST(ca, kpos)
                            SPH_F32_TEST(cb, kpos + 1)
                            SPH_F32_TEST(cc, kpos + 2)
                            SPH_F32_TEST(cd, kpos + 3)
                            kpos += 4;
                        }
                        while (kpos < khi && khi - kpos < 4 && cnt < LIST_CAP) {
                            const float4 ca = f4[kpos];
                            SPH_F32_TEST(ca, kpos)
                            kpos++;
                        }
#undef SPH_F32_TEST
                    }
                    if (kpos >= khi) {
                        r++;
                        open = false;
                    } else {
                        break; // list (nearly) full: drain it, then resume the scan
                    }
                }
                // phase 2: two list entries per trip so the long per-pair chains overlap
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
                for (int q = 0; q < 5; ++q) {
                    S[q] = d.f[F_S0 + q][slot];
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
    k_units<false><<<(total + 127) / 128, 128, 0, st>>>(ctx->d, ctx->maxCells);
    k_scan_block<<<ctx->scanBlocks, 512, 0, st>>>(ctx->d.cellCount, ctx->d.segStart, ctx->d.scanBlock, total);
    k_scan_sums<<<1, 1024, 0, st>>>(ctx->d.scanBlock, ctx->scanBlocks);
    k_scan_add<<<ctx->scanBlocks, 512, 0, st>>>(ctx->d.segStart, ctx->d.scanBlock, total);
    k_units<true><<<(ctx->maxCells + 127) / 128, 128, 0, st>>>(ctx->d, ctx->maxCells);
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
    const uint32_t upper = ctx->nActive / (TILE_T / 2) + ctx->maxCells + 1; // >= number of units
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
