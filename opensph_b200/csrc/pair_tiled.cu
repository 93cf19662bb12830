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
//              records fall into different bank groups); the FP32 {x,y,z,h} copies relative to the grid origin
//              (written by the prologue) are staged the same way.
// Phase 1    : every thread scans, per candidate row, the x-window IT can reach (interval culling in y and z, then
//              bisection on x: cell rows are sorted by x, see k_sort_cells) with a conservative FP32 distance test
//              (FP32/ALU pipes) and appends survivors to a private u16 list.
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
constexpr int TILE_X = 20;    // widest unit in cells (bounds the per-unit loops over candidate cells)
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
/// 16-bit store to shared memory through a 32-bit shared-window address (keeps the list cursor a single register).
__device__ __forceinline__ void storeSharedU16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void fenceProxyAsync() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- work list: units of <= 128 targets per double row ---------------------------------------------------------
// The targets of a double row are taken in COLUMN order: for every cell column c the lower cell's particles, then the
// upper cell's. One thread per double row cuts this sequence into pieces of 128 (a column may be split between two
// units, so units are full except at the end of a row or where the x-range would outgrow TILE_X). FILL = false counts
// the units, FILL = true writes the descriptors {double row, cA, skip, (cB - cA) | targets << 8}: the unit takes
// `targets` entries of the sequence of columns cA.., starting `skip` entries into column cA.
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
    auto columnCount = [&](int c) {
        uint32_t cnt = d.cellStart[rbL + c + 1] - d.cellStart[rbL + c];
        if (hasU) {
            cnt += d.cellStart[rbU + c + 1] - d.cellStart[rbU + c];
        }
        return cnt;
    };
    int c = 0;
    uint32_t skip = 0; // entries of column c already handed out
    while (c < dimx) {
        if (columnCount(c) - skip == 0) {
            c++;
            skip = 0;
            continue;
        }
        const int cA = c;
        const uint32_t skipA = skip;
        uint32_t taken = 0;
        int cLast = c;
        while (c < dimx && taken < (uint32_t)TILE_T && c - cA + 3 <= TILE_X) {
            const uint32_t avail = columnCount(c) - skip;
            if (avail == 0) {
                c++;
                skip = 0;
                continue;
            }
            const uint32_t take = min(avail, (uint32_t)TILE_T - taken);
            taken += take;
            cLast = c;
            if (take == avail) {
                c++;
                skip = 0;
            } else {
                skip += take;
                break;
            }
        }
        if (FILL) {
            d.unitDesc[out + units] = make_uint4(dr, (uint32_t)cA, skipA, (uint32_t)(cLast - cA) | (taken << 8));
        }
        units++;
    }
    if (!FILL) {
        d.cellCount[dr] = units;
    }
}

struct ChunkState {
    uint32_t beg[CHUNK_ROWS], end[CHUNK_ROWS], base[CHUNK_ROWS]; // staged global range per candidate row + smem offset
    uint32_t used;
    int chunk;                                  // 0 far, 1 near, 2 centre
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
    __shared__ uint32_t sHmax;
    __shared__ uint32_t sColL[TILE_X + 2], sColU[TILE_X + 2], sTarget[TILE_T];

    const GridDev g = *d.grid;
    const int dimx = g.dim[0], dimy = g.dim[1], dimz = g.dim[2];
    const uint32_t totalUnits = d.segStart[maxCells];
    const int tid = threadIdx.x;
    const float Rhalf = (float)(0.5 * c_prm.kernel_radius * (1. + 2.e-5));
    // the FP32 coordinates are relative to the grid origin: absolute rounding error <= 2^-24 * extent per coordinate
    const float slack = (float)(g.extent * 1.e-6), guard = (float)(g.extent * 1.e-5);
    const float cellF = (float)g.cell, cellZF = (float)g.cellZ;
    const bool rowsSorted = g.unsorted == 0u;
    ChunkCursor cur;
    uint32_t stagePhase = 0;
    if (tid == 0) {
        mbarInit(&stageBar, 1);
    }
    __syncthreads();

    for (uint32_t unit = blockIdx.x; unit < totalUnits; unit += gridDim.x) {
        const uint4 desc = d.unitDesc[unit];
        const uint32_t dr = desc.x, skip = desc.z, nLive = desc.w >> 8;
        const int cA = (int)desc.y, span = (int)(desc.w & 0xffu), cB = cA + span;
        const int cy = (int)(dr % (uint32_t)dimy), k = (int)(dr / (uint32_t)dimy);
        const uint32_t rbL = (uint32_t)(((2 * k) * dimy + cy) * dimx);
        const bool hasU = 2 * k + 1 < dimz;
        const uint32_t rbU = hasU ? (uint32_t)(((2 * k + 1) * dimy + cy) * dimx) : 0u;
        const int x0 = max(cA - 1, 0), x1 = min(cB + 1, dimx - 1);

        // ---- sorted ranges of the unit's cell columns (targets) and of its 18 candidate rows ----
        if (tid <= span + 1) {
            sColL[tid] = d.cellStart[rbL + cA + tid];
            sColU[tid] = hasU ? d.cellStart[rbU + cA + tid] : 0u;
        }
        if (tid == 0) {
            sHmax = 0u;
        }
        if (tid >= 32 && tid < 32 + 3 * CHUNK_ROWS) {
            const int row = tid - 32;
            const int z = chunkLayer(row / CHUNK_ROWS, row % CHUNK_ROWS, k), y = cy + (row % 3) - 1;
            uint32_t rb = 0, re = 0;
            if (z >= 0 && z < dimz && y >= 0 && y < dimy) {
                const uint32_t base = (uint32_t)((z * dimy + y) * dimx);
                rb = d.cellStart[base + x0];
                re = d.cellStart[base + x1 + 1];
            }
            sRowBeg[row] = rb;
            sRowEnd[row] = re;
        }
        __syncthreads();
        // ---- lane assignment: the unit's targets in column order (lower cell, then upper cell of every column), then
        // ordered by z so that the lanes of a warp see similar numbers of neighbours in every chunk (balanced private
        // lists => full lanes in phase 2)
        {
            uint32_t tIdx = 0u; // sorted index of target `tid` of the unit, bit 31 = upper row
            if ((uint32_t)tid < nLive) {
                uint32_t pos = skip + (uint32_t)tid;
                for (int c = 0; c <= span; ++c) {
                    const uint32_t nl = sColL[c + 1] - sColL[c], nu = sColU[c + 1] - sColU[c];
                    if (pos < nl) {
                        tIdx = sColL[c] + pos;
                        break;
                    }
                    pos -= nl;
                    if (pos < nu) {
                        tIdx = (sColU[c] + pos) | 0x80000000u;
                        break;
                    }
                    pos -= nu;
                }
            }
            sTarget[tid] = tIdx;
            sKey[tid] = ((uint32_t)tid < nLive) ? d.posF[tIdx & 0x7fffffffu].z : 3.0e38f;
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
            // largest h among the unit's candidates (cells x0..x1 of the 18 rows): bound of the FP32 filter radius
            uint32_t hm = 0u;
            for (int e = tid; e < 3 * CHUNK_ROWS * (TILE_X + 2); e += TILE_T) {
                const int row = e / (TILE_X + 2), c = x0 + e % (TILE_X + 2);
                const int z = chunkLayer(row / CHUNK_ROWS, row % CHUNK_ROWS, k), y = cy + (row % 3) - 1;
                if (c <= x1 && z >= 0 && z < dimz && y >= 0 && y < dimy) {
                    hm = max(hm, d.cellHmax[(uint32_t)((z * dimy + y) * dimx) + c]); // bit patterns of floats >= 0 order like the values
                }
            }
            hm = __reduce_max_sync(0xffffffffu, hm);
            if ((tid & 31) == 0) {
                atomicMax(&sHmax, hm);
            }
        }
        __syncthreads();
        const bool live = (uint32_t)tid < nLive; // the idle lanes sort last (key 3e38)
        const uint32_t tOwn = sTarget[sPerm[tid]];
        const uint32_t t = live ? (tOwn & 0x7fffffffu) : 0u;
        const bool upper = live && (tOwn >> 31) != 0u;
        const uint32_t slot = live ? d.order[t] : 0xffffffffu;
        const bool target = live && slot < nOwned; // ghosts are neighbours only
        if (!__syncthreads_or(target ? 1 : 0)) {
            continue; // a unit made of ghost particles only (halo band of a decomposed run)
        }

        Particle pi;
        float fxi = 0.f, fyi = 0.f, fzi = 0.f;
        float lim = 0.f;
        if (live) {
            loadRecord<SOLID>(d.rec + (size_t)t * L::G, pi);
            const float4 pf = d.posF[t];
            fxi = pf.x;
            fyi = pf.y;
            fzi = pf.z;
            // >= R * hbar + the rounding of the FP32 coordinates, for every candidate of the unit
            lim = fmaf(Rhalf, pf.w * (1.f + 1.e-6f) + __uint_as_float(sHmax), slack);
        } else {
            pi.x = pi.y = pi.z = 0.;
            pi.h = 1.;
        }
        const float lim2 = lim * lim;
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
            // ---- stage the chunk: TMA bulk copies of the FP64 records and of the FP32 positions, one pair per row ----
            if (tid == 0) {
                fenceProxyAsync(); // the buffers were last read through the generic proxy
                mbarExpectTx(&stageBar, cs.used * (uint32_t)(L::G * 8 + 16));
#pragma unroll
                for (int r = 0; r < CHUNK_ROWS; ++r) {
                    const uint32_t n = cs.end[r] - cs.beg[r];
                    if (n > 0) {
                        bulkCopyG2S(recS + (size_t)cs.base[r] * L::S, d.rec + (size_t)cs.beg[r] * L::G, n * (uint32_t)(L::G * 8), &stageBar);
                        bulkCopyG2S(f4 + cs.base[r], d.posF + cs.beg[r], n * 16u, &stageBar);
                    }
                }
                nextChunk(sRowBeg, sRowEnd, cur, csBuf[buf ^ 1]); // overlaps with the copies
            }
            mbarWait(&stageBar, stagePhase);
            stagePhase ^= 1;
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
            const uint32_t listOwn = smemAddr(list + tid); // byte address of this lane's first list slot
            constexpr uint32_t LIST_STRIDE = TILE_T * 2u;
            int r = 0;
            uint32_t kpos = 0, khi = 0;
            bool open = false;
            while (r < CHUNK_ROWS) {
                uint32_t lp = listOwn;
                while (r < CHUNK_ROWS) {
                    if (!open) {
                        const uint32_t b = cs.beg[r], len = cs.end[r] - b;
                        kpos = khi = 0;
                        if (len > 0) {
                            // y- and z-intervals of this candidate row (FP32 relative to the grid origin, with a guard band:
                            // cells were assigned in FP64) and the x-window [xlo, xhi] the target can reach in it
                            const int zabs = chunkLayer(chunk, r, k), yabs = cy + (r % 3) - 1;
                            const float yl = (float)yabs * cellF, zl = (float)zabs * cellZF;
                            const float dyMin = fmaxf(fmaxf(yl - fyi, fyi - (yl + cellF)) - guard, 0.f);
                            const float dzMin = fmaxf(fmaxf(zl - fzi, fzi - (zl + cellZF)) - guard, 0.f);
                            const float rem = lim2 - dyMin * dyMin - dzMin * dzMin;
                            const float ext = sqrtf(fmaxf(rem, 0.f)) * (1.f + 1.e-5f) + guard;
                            const float xlo = rem > 0.f ? fxi - ext : 3.0e38f, xhi = rem > 0.f ? fxi + ext : -3.0e38f;
                            const uint32_t pieceBase = cs.base[r];
                            uint32_t pl = 0, ph = len;
                            if (rowsSorted) {
                                // bisection for both ends at once: pl = #{x < xlo}, ph = #{x <= xhi}; the trip count depends on
                                // the row only, so the warp stays converged
                                const float* fx = reinterpret_cast<const float*>(f4 + pieceBase);
                                ph = 0;
                                for (uint32_t step = 1u << (31 - __clz(len)); step > 0; step >>= 1) {
                                    const uint32_t tl = pl + step, th = ph + step;
                                    const float vl = fx[4 * (min(tl, len) - 1)], vh = fx[4 * (min(th, len) - 1)];
                                    pl = (tl <= len && vl < xlo) ? tl : pl;
                                    ph = (th <= len && vh <= xhi) ? th : ph;
                                }
                            } else if (!(rem > 0.f)) {
                                ph = 0;
                            }
                            if (ph > pl) {
                                kpos = pieceBase + pl;
                                khi = pieceBase + ph;
                            }
                        }
                        open = true;
                    }
                    // eight candidates per trip: the loads are independent, only the list append is serial. The target
                    // itself is not excluded here (~300 compares) but masked in phase 2 (~70 compares).
                    {
#define SPH_F32_TEST(C, K)                                                                                            \
    {                                                                                                                 \
        const float ddx = fxi - C.x, ddy = fyi - C.y, ddz = fzi - C.z;                                                \
        const float dd2 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));                                                  \
        if (dd2 <= lim2) {                                                                                            \
            storeSharedU16(lp, K);                                                                                    \
            lp += LIST_STRIDE;                                                                                        \
        }                                                                                                             \
    }
                        const uint32_t lpMax8 = listOwn + (LIST_CAP - 8) * LIST_STRIDE;
                        const uint32_t lpMax4 = listOwn + (LIST_CAP - 4) * LIST_STRIDE;
                        while (kpos + 8 <= khi && lp <= lpMax8) {
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
                        while (kpos + 4 <= khi && lp <= lpMax4) {
                            const float4 ca = f4[kpos], cb = f4[kpos + 1], cc = f4[kpos + 2], cd = f4[kpos + 3];
                            SPH_F32_TEST(ca, kpos)
                            SPH_F32_TEST(cb, kpos + 1)
                            SPH_F32_TEST(cc, kpos + 2)
                            SPH_F32_TEST(cd, kpos + 3)
                            kpos += 4;
                        }
                        while (kpos < khi && khi - kpos < 4 && lp < listOwn + LIST_CAP * LIST_STRIDE) {
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
                const int cnt = (int)((lp - listOwn) / LIST_STRIDE);
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
