// The pair stage: fused neighbour search + pair sums of the asymmetric solver, as a chain of kernels over WORK UNITS.
//
// Grid       : cells are a x a x a/2 (a = R h_max); a "double row" is the two half-height cell rows (cy, 2k), (cy, 2k+1).
// Work unit  = up to 128 targets of ONE double row (column order, a cell column may be split between units), one
//              target per thread, per-target sums in registers, no atomics (asymmetric formulation). k_units builds the
//              units, k_unit_prep orders the targets of a unit by z and stores the lane assignment.
// Candidates = the six z-layers 2k-2 .. 2k+3 (three cell rows dy = -1,0,1 each) restricted to cells [cA-1, cB+1]. They
//              are processed as three CHUNKS that pair layers symmetrically: far (2k-2, 2k+3), near (2k-1, 2k+2),
//              centre (2k, 2k+1). With the lanes ordered by z, every warp then sees about the same number of
//              neighbours in every chunk (measured on the hex lattice: 5/19/45 per chunk for all four z-bands, versus
//              25/43/0 ... 0/43/25 when whole layers are processed one after the other), so the CTA-wide barrier at the
//              chunk boundaries costs little.
// Phase 1    = k_pair_lists (FP32/ALU pipes, 32 warps per SM): TMA bulk copies (cp.async.bulk global -> shared,
//              completion on an mbarrier) stage the FP32 {x,y,z,h} copies of a chunk (relative to the grid origin, written
//              by the prologue); every thread scans, per candidate row, the x-window IT can reach (interval culling in y
//              and z, then bisection on x: cell rows are sorted by x, see k_sort_cells) with a conservative FP32 distance
//              test and appends survivors to its u16 list. The list is then SCHEDULED against shared-memory bank
//              conflicts of phase 2 (entry q of lane l prefers a candidate whose staged index is q + l mod 8, see
//              scheduleList) and leaves as one block of the list pool; blocks are chained per unit through descriptors.
// Phase 2    = k_pair_sum (FP64 pipe, 12 warps per SM): follows the chains; per block TMA copies of the chunk's records
//              (128-byte FP64 structures with XOR-swizzled 16-byte pieces, sphgpu_internal.h) into the single stage;
//              every thread walks its list (read from the pool through L2, prefetched four entries ahead) two entries
//              at a time; the exact FP64 predicate (bit-identical neighbour sets) enters the branch-free FP64 pair body
//              as a mask.
// k_pair_fallback evaluates the units whose lists overflowed the pool with the direct per-target loop of pair.cu (and
// every unit in variant 2, the cross-check of the unit bookkeeping).
//
// Replaces the reference hot loop AsymmetricSolver.cpp:174-201 (finder.findAll + filter + kernel.grad +
// derivatives.eval) -- see pair.cu for the prologue, the epilogue helpers and the direct variant.
#include "sphgpu_internal.h"

namespace sph {

#ifndef SPH_TILE_T
#define SPH_TILE_T 128
#endif
#ifndef SPH_TILE_C
#define SPH_TILE_C 576
#endif
constexpr int TILE_T = SPH_TILE_T;   // targets (threads) per work unit
constexpr int TILE_C = SPH_TILE_C;   // staged candidate slots per chunk (incl. the alignment gaps between row pieces)
constexpr int LIST_CAP = 60;  // list entries per lane and block (15 quads; 63 is the null link of scheduleList)
constexpr int UNIT_KBLOCK = 4;    // double rows (in z) interleaved in the unit order, see k_units
constexpr int PAIRS_PER_TRIP = 2; // list entries the pair-sum kernel processes together
constexpr int TILE_X = 20;    // widest unit in cells (bounds the per-unit loops over candidate cells)
constexpr int CHUNK_ROWS = 6; // candidate rows per chunk: 2 z-layers x 3 y-rows

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
/// Asynchronous copies global -> shared (LDGSTS): the data never occupies a register while it is in flight.
__device__ __forceinline__ void copyAsync4(void* dstSmem, const void* srcGlobal) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal) : "memory");
}
__device__ __forceinline__ void copyAsync8(void* dstSmem, const void* srcGlobal) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal) : "memory");
}
__device__ __forceinline__ void copyAsyncWaitAll() {
    asm volatile("cp.async.wait_all;" ::: "memory");
}
/// Predicated 8-byte read-only global load (keeps the caller's code free of branches).
__device__ __forceinline__ void loadGlobalU2If(bool p, const uint2* ptr, uint2& v) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.global.nc.v2.u32 {%0, %1}, [%2];\n\t}"
                 : "+r"(v.x), "+r"(v.y)
                 : "l"(ptr), "r"((uint32_t)p));
}
/// 16-byte load from shared memory through a 32-bit shared-window address (LDS.128).
__device__ __forceinline__ double2 loadSharedD2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t loadSharedU16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t loadSharedU8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void storeSharedU8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// ---- work list: units of <= 128 targets per double row ---------------------------------------------------------
// The targets of a double row are taken in COLUMN order: for every cell column c the lower cell's particles, then the
// upper cell's. One thread per double row cuts this sequence into pieces of 128 (a column may be split between two
// units, so units are full except at the end of a row or where the x-range would outgrow TILE_X). FILL = false counts
// the units, FILL = true writes the descriptors {double row, cA, skip, (cB - cA) | targets << 8}: the unit takes
// `targets` entries of the sequence of columns cA.., starting `skip` entries into column cA.
template <bool FILL>
__global__ void __launch_bounds__(128) k_units(DevicePointers d, uint32_t maxCells) {
    // one WARP per double row: the lanes fetch the column counts of 256 columns at a time into shared memory (the loads
    // overlap), lane 0 then walks them; a thread-per-row walk would pay one L2 round trip per column
    constexpr int CHUNK = 256;
    __shared__ uint32_t sCnt[4][CHUNK];
    if (d.listCtl->rebuild == 0u) {
        return; // the units of the last build are reused
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GridDev g = *d.grid;
    const int dimx = g.dim[0], dimy = g.dim[1], dimz = g.dim[2];
    const uint32_t doubleRows = min((uint32_t)dimy * (uint32_t)((dimz + 1) / 2), maxCells);
    // (cellCount, free once cellStart has been built, receives the unit counts; the host zeroed it beyond the double rows)
    // Units are numbered in the order the kernels process them. Double rows are taken in blocks of UNIT_KBLOCK
    // consecutive k (z) per cy, so that the CTAs running at the same time work on z-neighbours as well as y-neighbours
    // and find each other's candidate rows in L2 (plain k-major order re-reads every record from DRAM for k - 1, k, k + 1).
    const int nk = (dimz + 1) / 2;
    for (uint32_t p = blockIdx.x * 4 + warp; p < doubleRows; p += gridDim.x * 4) {
        const int kb = min((int)(p / (uint32_t)(UNIT_KBLOCK * dimy)), (nk - 1) / UNIT_KBLOCK);
        const int hB = min(UNIT_KBLOCK, nk - kb * UNIT_KBLOCK);
        const uint32_t rem = p - (uint32_t)(kb * UNIT_KBLOCK * dimy);
        const int cy = (int)(rem / (uint32_t)hB), k = kb * UNIT_KBLOCK + (int)(rem % (uint32_t)hB);
        const uint32_t dr = (uint32_t)(k * dimy + cy);
        const uint32_t rbL = (uint32_t)(((2 * k) * dimy + cy) * dimx);
        const bool hasU = 2 * k + 1 < dimz;
        const uint32_t rbU = hasU ? (uint32_t)(((2 * k + 1) * dimy + cy) * dimx) : 0u;
        uint32_t units = 0;
        const uint32_t out = FILL ? d.segStart[p] : 0u;
        // walk state (lane 0): the open unit [cA, cLast] with `taken` targets, `skipA` entries into column cA
        int cA = 0, cLast = 0;
        uint32_t skipA = 0, taken = 0;
        auto emit = [&]() {
            if (FILL) {
                d.unitDesc[out + units] = make_uint4(dr, (uint32_t)cA, skipA, (uint32_t)(cLast - cA) | (taken << 8));
            }
            units++;
            taken = 0;
        };
        for (int c0 = 0; c0 < dimx; c0 += CHUNK) {
            const int n = min(CHUNK, dimx - c0);
            for (int i = lane; i < n; i += 32) {
                uint32_t cnt = d.cellStart[rbL + c0 + i + 1] - d.cellStart[rbL + c0 + i];
                if (hasU) {
                    cnt += d.cellStart[rbU + c0 + i + 1] - d.cellStart[rbU + c0 + i];
                }
                sCnt[warp][i] = cnt;
            }
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < n; ++i) {
                    const int c = c0 + i;
                    uint32_t avail = sCnt[warp][i];
                    uint32_t skip = 0; // entries of column c already handed out
                    while (avail > 0) {
                        if (taken > 0 && c - cA + 3 > TILE_X) {
                            emit(); // the x-range would outgrow the tables of the unit preparation
                        }
                        if (taken == 0) {
                            cA = c;
                            skipA = skip;
                        }
                        const uint32_t take = min(avail, (uint32_t)TILE_T - taken);
                        taken += take;
                        skip += take;
                        avail -= take;
                        cLast = c;
                        if (taken == (uint32_t)TILE_T) {
                            emit();
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            if (taken > 0) {
                emit();
            }
            if (!FILL) {
                d.cellCount[p] = units;
            }
        }
        __syncwarp();
    }
}

/// Where the lane words of a unit start in unitLane: units partition the particles, so the number of targets of all
/// earlier units (earlier double rows, then earlier columns of this one) is a sum of cellStart differences.
__device__ __forceinline__ uint32_t unitLaneBase(const DevicePointers& d, int dimx, int dimy, int dimz, int k, int cy, int cA, uint32_t skip) {
    const uint32_t layerL = (uint32_t)((2 * k) * dimy * dimx), rowL = layerL + (uint32_t)(cy * dimx);
    uint32_t base = d.cellStart[rowL + cA]; // all layers below 2k + earlier rows of layer 2k + earlier columns of its row cy
    if (2 * k + 1 < dimz) {
        const uint32_t layerU = (uint32_t)((2 * k + 1) * dimy * dimx), rowU = layerU + (uint32_t)(cy * dimx);
        base += d.cellStart[rowU + cA] - d.cellStart[layerU]; // earlier rows of layer 2k+1 + earlier columns of its row cy
    }
    return base + skip;
}

// ---- unit preparation: lane order, filter radius bound, ghost-only flag -----------------------------------------
// One CTA per unit (grid-stride). Orders the unit's targets by z (stable counting sort on 32 height bins): the lanes of
// a warp then see similar numbers of neighbours in every chunk => full lanes in phase 2. The order is stored once, so
// that the register- and shared-memory-heavy pair kernels start every unit with a single coalesced load.
__global__ void __launch_bounds__(TILE_T) k_unit_prep(DevicePointers d, uint32_t nOwned, uint32_t maxCells) {
    constexpr int ZBINS = 32; // z resolution of the lane order: 1/32 of the double row's height
    __shared__ uint32_t sColL[TILE_X + 2], sColU[TILE_X + 2], sHmax;
    __shared__ uint32_t sBinW[TILE_T / 32][ZBINS], sBase[ZBINS];
    if (d.listCtl->rebuild == 0u) {
        return;
    }
    const GridDev g = *d.grid;
    const int dimx = g.dim[0], dimy = g.dim[1], dimz = g.dim[2];
    const uint32_t totalUnits = d.segStart[maxCells];
    const int tid = threadIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float binScale = (float)(ZBINS / (2. * g.cellZ));
    for (uint32_t unit = blockIdx.x; unit < totalUnits; unit += gridDim.x) {
        const uint4 desc = d.unitDesc[unit];
        const uint32_t dr = desc.x, skip = desc.z, nLive = desc.w >> 8;
        const int cA = (int)desc.y, span = (int)(desc.w & 0xffu), cB = cA + span;
        const int cy = (int)(dr % (uint32_t)dimy), k = (int)(dr / (uint32_t)dimy);
        const uint32_t rbL = (uint32_t)(((2 * k) * dimy + cy) * dimx);
        const bool hasU = 2 * k + 1 < dimz;
        const uint32_t rbU = hasU ? (uint32_t)(((2 * k + 1) * dimy + cy) * dimx) : 0u;
        const int x0 = max(cA - 1, 0), x1 = min(cB + 1, dimx - 1);
        __syncthreads(); // the previous unit is done with the shared tables
        if (tid <= span + 1) {
            sColL[tid] = d.cellStart[rbL + cA + tid];
            sColU[tid] = hasU ? d.cellStart[rbU + cA + tid] : 0u;
        }
        if (tid == 0) {
            sHmax = 0u;
        }
        sBinW[warp][lane] = 0u;
        __syncthreads();
        // the unit's targets in column order: lower cell, then upper cell of every column
        uint32_t word = 0xffffffffu; // idle lane
        bool owned = false;
        uint32_t bin = ZBINS + (uint32_t)lane; // idle lanes: a key of their own
        if ((uint32_t)tid < nLive) {
            uint32_t pos = skip + (uint32_t)tid, tIdx = 0u;
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
            const uint32_t t = tIdx & 0x3fffffffu;
            owned = d.order[t] < nOwned; // ghosts are neighbours only
            word = tIdx | (owned ? 0u : 0x40000000u);
            const float zrel = d.posF[t].z - (float)(2 * k) * (float)g.cellZ; // height above the bottom of the double row
            bin = (uint32_t)min(max((int)(zrel * binScale), 0), ZBINS - 1);
        }
        // stable counting sort by z bin: lanes of equal bin keep the column (x) order
        const uint32_t same = __match_any_sync(0xffffffffu, bin);
        const uint32_t lower = __popc(same & ((1u << lane) - 1u));
        if (bin < (uint32_t)ZBINS && lower == 0u) {
            sBinW[warp][bin] = __popc(same);
        }
        // largest h among the unit's candidates (cells x0..x1 of the 18 rows): bound of the FP32 filter radius
        uint32_t hm = 0u;
        for (int e = tid; e < 18 * (TILE_X + 2); e += TILE_T) {
            const int row = e / (TILE_X + 2), c = x0 + e % (TILE_X + 2);
            const int z = 2 * k - 2 + row / 3, y = cy + (row % 3) - 1;
            if (c <= x1 && z >= 0 && z < dimz && y >= 0 && y < dimy) {
                hm = max(hm, d.cellHmax[(uint32_t)((z * dimy + y) * dimx) + c]); // bit patterns of floats >= 0 order like the values
            }
        }
        hm = __reduce_max_sync(0xffffffffu, hm);
        if ((tid & 31) == 0) {
            atomicMax(&sHmax, hm);
        }
        const int anyOwned = __syncthreads_or(owned ? 1 : 0);
        if (warp == 0) { // exclusive prefix of the bin totals
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < TILE_T / 32; ++w) {
                tot += sBinW[w][lane];
            }
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                incl += lane >= o ? v : 0u;
            }
            sBase[lane] = incl - tot;
        }
        __syncthreads();
        const uint32_t base = unitLaneBase(d, dimx, dimy, dimz, k, cy, cA, skip);
        if ((uint32_t)tid < nLive) {
            uint32_t rank = sBase[bin] + lower;
            for (int w = 0; w < warp; ++w) {
                rank += sBinW[w][bin];
            }
            d.unitLane[base + rank] = word;
        }
        if (tid == 0) {
            d.unitAux[unit] = make_uint4(base, sHmax, (uint32_t)anyOwned, 0u);
        }
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
            // the staged index of a record is congruent to its sorted index mod 8 (swizzle of the solid records, bank
            // schedule of the lists): up to 7 slots are skipped in front of a row piece
            const uint32_t base = used + ((cur.pos - used) & 7u);
            const uint32_t take = (re > cur.pos && base < (uint32_t)TILE_C) ? min(re - cur.pos, (uint32_t)TILE_C - base) : 0u;
            if (take > 0) {
                cs.beg[cur.row] = cur.pos;
                cs.end[cur.row] = cur.pos + take;
                cs.base[cur.row] = base;
                used = base + take;
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

// ---- pieces shared by the list kernel and the pair-sum kernel ---------------------------------------------------
constexpr uint32_t LIST_STRIDE = TILE_T * 2u;        // bytes between consecutive entries of one lane's list
constexpr uint32_t LIST_END = 0xffffffffu;           // unitList / block header: no (further) block
constexpr uint32_t LIST_FALLBACK = 0xfffffffeu;      // unitList: the pool was exhausted, the unit goes to k_pair_fallback
constexpr uint32_t LIST_ROW_BYTES = TILE_T * 2u;     // one row of the list pool (and of the list in shared memory): 256 B
constexpr uint32_t LIST_QUAD_BYTES = TILE_T * 8u;    // one quad row of a block: four u16 entries per lane
constexpr uint32_t LIST_NULL = 63u;                  // scheduleList: end of a residue chain
// A list block in the pool is 1 + 4 Q rows of 256 B. Row 0: bytes 0..127 = per-lane entry counts (u8), bytes 128..187 =
// the descriptor of the NEXT block of the unit (word 0 = LIST_END terminates the chain). Then Q quad rows of 1 KB:
// [quad][lane][4] u16, entry q of a lane in quad q / 4, each entry the staged (shared-memory) index of a candidate of the
// chunk. unitList[4 * unit ..] holds the descriptor of the unit's first block.

/// What a thread knows about its target and its unit.
struct UnitLane {
    int k, cy;          // double row
    uint32_t t;         // sorted index of the target
    uint32_t slot;      // particle slot of the target (0xffffffff for ghosts / idle lanes)
    bool live, upper, target;
    float fx, fy, fz;   // grid-relative FP32 position
    float lim2;         // squared FP32 filter radius
};

/// Decodes the lane of thread `tid` in a unit (no synchronisation).
__device__ __forceinline__ void loadLane(const DevicePointers& d, const GridDev& g, const uint4& desc, const uint4& aux, int tid, float Rhalf,
    float slack, UnitLane& u) {
    const uint32_t dr = desc.x, nLive = desc.w >> 8;
    u.cy = (int)(dr % (uint32_t)g.dim[1]);
    u.k = (int)(dr / (uint32_t)g.dim[1]);
    // lane order prepared by k_unit_prep: sorted index | upper row << 31 | ghost << 30
    u.live = (uint32_t)tid < nLive;
    const uint32_t word = u.live ? d.unitLane[aux.x + tid] : 0xffffffffu;
    u.t = u.live ? (word & 0x3fffffffu) : 0u;
    u.upper = u.live && (word >> 31) != 0u;
    u.target = u.live && (word & 0x40000000u) == 0u; // ghosts are neighbours only
    u.slot = u.target ? d.order[u.t] : 0xffffffffu;
    u.fx = u.fy = u.fz = 0.f;
    u.lim2 = 0.f;
    if (u.live) {
        const float4 pf = d.posF[u.t];
        u.fx = pf.x;
        u.fy = pf.y;
        u.fz = pf.z;
        // >= R * hbar + the rounding of the FP32 coordinates, for every candidate of the unit
        const float lim = fmaf(Rhalf, pf.w * (1.f + 1.e-6f) + __uint_as_float(aux.y), slack);
        u.lim2 = lim * lim;
    }
}

/// Decodes unit `unit` for thread `tid`, publishes the unit's 18 candidate row ranges and the first chunk (thread 0).
/// Returns false (for the whole CTA) if the unit has no owned target. Contains one CTA-wide barrier.
__device__ __forceinline__ bool beginUnit(const DevicePointers& d, const GridDev& g, uint32_t unit, int tid, float Rhalf, float slack,
    uint32_t* sRowBeg, uint32_t* sRowEnd, ChunkCursor& cur, ChunkState& cs0, UnitLane& u) {
    const uint4 desc = d.unitDesc[unit];
    const uint4 aux = d.unitAux[unit];
    if (aux.z == 0u) {
        return false; // a unit made of ghost particles only (halo band of a decomposed run)
    }
    const int dimx = g.dim[0], dimy = g.dim[1], dimz = g.dim[2];
    const int cA = (int)desc.y, cB = cA + (int)(desc.w & 0xffu);
    const int x0 = max(cA - 1, 0), x1 = min(cB + 1, dimx - 1);
    loadLane(d, g, desc, aux, tid, Rhalf, slack, u);
    __syncthreads(); // everyone has left the previous unit's loop
    if (tid < 3 * CHUNK_ROWS) { // global sorted ranges of the unit's 18 candidate rows
        const int z = chunkLayer(tid / CHUNK_ROWS, tid % CHUNK_ROWS, u.k), y = u.cy + (tid % 3) - 1;
        uint32_t rb = 0, re = 0;
        if (z >= 0 && z < dimz && y >= 0 && y < dimy) {
            const uint32_t base = (uint32_t)((z * dimy + y) * dimx);
            rb = d.cellStart[base + x0];
            re = d.cellStart[base + x1 + 1];
        }
        sRowBeg[tid] = rb;
        sRowEnd[tid] = re;
    }
    if (tid < 32) {
        __syncwarp();
        if (tid == 0) {
            cur.chunk = 0;
            cur.row = 0;
            cur.posValid = 0;
            nextChunk(sRowBeg, sRowEnd, cur, cs0);
        }
    }
    return true;
}

struct ScanGeometry {
    float cellF, cellZF, guard;
    bool rowsSorted;
};

struct ScanState { // where a lane stands in the candidate rows of the staged chunk
    int r;
    uint32_t kpos, khi;
    bool open;
};

/// Phase 1: scans the candidate rows of the staged chunk from `st` on, appending the shared-memory index of every
/// candidate within the FP32 filter radius to the lane's list (u16 at listOwn + n * LIST_STRIDE), until the chunk is
/// exhausted (st.r == CHUNK_ROWS) or the list is (nearly) full. Returns the list cursor.
__device__ __forceinline__ uint32_t scanRows(const ChunkState& cs, const UnitLane& u, const ScanGeometry& sg, const float4* f4,
    uint32_t listOwn, ScanState& st) {
    uint32_t lp = listOwn;
    const float fxi = u.fx, fyi = u.fy, fzi = u.fz, lim2 = u.lim2;
    while (st.r < CHUNK_ROWS) {
        if (!st.open) {
            const int r = st.r;
            const uint32_t len = cs.end[r] - cs.beg[r];
            st.kpos = st.khi = 0;
            if (len > 0) {
                // y- and z-intervals of this candidate row (FP32 relative to the grid origin, with a guard band: cells
                // were assigned in FP64) and the x-window [xlo, xhi] the target can reach in it
                const int zabs = chunkLayer(cs.chunk, r, u.k), yabs = u.cy + (r % 3) - 1;
                const float yl = (float)yabs * sg.cellF, zl = (float)zabs * sg.cellZF;
                const float dyMin = fmaxf(fmaxf(yl - fyi, fyi - (yl + sg.cellF)) - sg.guard, 0.f);
                const float dzMin = fmaxf(fmaxf(zl - fzi, fzi - (zl + sg.cellZF)) - sg.guard, 0.f);
                const float rem = lim2 - dyMin * dyMin - dzMin * dzMin;
                if (rem > 0.f) {
                    const float ext = sqrtf(rem) * (1.f + 1.e-5f) + sg.guard;
                    const float xlo = fxi - ext, xhi = fxi + ext;
                    const uint32_t pieceBase = cs.base[r];
                    uint32_t pl = 0, ph = len;
                    if (sg.rowsSorted) {
                        // bisection for both ends at once (cell rows are sorted by x): pl = #{x < xlo}, ph = #{x <= xhi}
                        const float* fx = reinterpret_cast<const float*>(f4 + pieceBase);
                        ph = 0;
                        for (uint32_t step = 1u << (31 - __clz(len)); step > 0; step >>= 1) {
                            const uint32_t tl = pl + step, th = ph + step;
                            const float vl = fx[4 * (min(tl, len) - 1)], vh = fx[4 * (min(th, len) - 1)];
                            pl = (tl <= len && vl < xlo) ? tl : pl;
                            ph = (th <= len && vh <= xhi) ? th : ph;
                        }
                    }
                    if (ph > pl) {
                        st.kpos = pieceBase + pl;
                        st.khi = pieceBase + ph;
                    }
                }
            }
            st.open = true;
        }
        // eight candidates per trip: the loads are independent, only the list append is serial. The target itself is
        // not excluded here (~300 compares) but masked in phase 2 (~70 compares).
        uint32_t kpos = st.kpos;
        const uint32_t khi = st.khi;
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
        st.kpos = kpos;
        if (kpos >= khi) {
            st.r++;
            st.open = false;
        } else {
            break; // list (nearly) full: drain it, then resume the scan
        }
    }
    return lp;
}

// ---- list schedule: ordering a lane's list against bank conflicts in phase 2 -------------------------------------------
// Phase 2 gathers, per list entry, one staged record per lane with LDS.128; the 8 lanes of a quarter warp are served
// together and collide when their records' staged indices agree mod 8 (equal indices broadcast). Entry q of lane l
// therefore PREFERS a candidate of residue class (q + l) mod 8: as long as every lane of a quarter finds its preferred
// class, the quarter reads 8 different bank groups. A lane whose preferred class has run out takes the next non-empty
// one. Measured with profiles/conflict_model.py on the bench lattice: 1.92 -> 1.36 wavefronts per ideal wavefront (a
// jittered lattice: 2.27 -> 1.41).
//
// chainList (backwards over the lane's list) threads the entries of every class into a chain: bits 10..15 of an entry
// receive the list position of the next entry of its class, byte b of `heads` the first one of class b. emitList pops
// the chains in the preferred order and writes the quads {4 x u16} straight to the block in the pool. Both loops run in
// lockstep over the warp (every lane is at the same list row, so the u16 accesses of two lanes that share a 32-bit word
// do not collide); the eight chain heads live in two registers.
struct ChainHeads {
    uint32_t lo, hi; // byte b: list position of the first entry of class b (LIST_NULL: none)
};

__device__ __forceinline__ uint32_t headGet(const ChainHeads& h, uint32_t b) {
    return __byte_perm(h.lo, h.hi, b) & 0xffu;
}
__device__ __forceinline__ void headSet(ChainHeads& h, uint32_t b, uint32_t v) {
    // byte (b & 3) of the selected word <- v: PRMT selector 0x3210 with nibble (b & 3) replaced by 4 (= byte 0 of v)
    const uint32_t n = b & 3u;
    const uint32_t sel = 0x3210u + ((4u - n) << (4u * n));
    const uint32_t lo = __byte_perm(h.lo, v, sel), hi = __byte_perm(h.hi, v, sel);
    h.lo = b < 4u ? lo : h.lo;
    h.hi = b < 4u ? h.hi : hi;
}

__device__ __forceinline__ void chainList(uint32_t listOwn, uint32_t cnt, uint32_t warpMax, ChainHeads& heads) {
    heads.lo = heads.hi = LIST_NULL * 0x01010101u;
    for (uint32_t e = warpMax; e-- > 0;) {
        if (e < cnt) {
            const uint32_t addr = listOwn + e * LIST_STRIDE;
            const uint32_t v = loadSharedU16(addr);
            const uint32_t b = v & 7u;
            storeSharedU16(addr, v | (headGet(heads, b) << 10));
            headSet(heads, b, e);
        }
    }
}

__device__ __forceinline__ void emitList(uint32_t listOwn, uint32_t cnt, uint32_t warpMax, uint32_t lane, ChainHeads heads, uint2* quadOut) {
    uint32_t lo = 0u, hi = 0u;
    uint32_t live = 0u; // bit b: class b still has entries
#pragma unroll
    for (uint32_t b = 0; b < 8; ++b) {
        live |= headGet(heads, b) != LIST_NULL ? 1u << b : 0u;
    }
    for (uint32_t q = 0; q < warpMax; ++q) {
        if (q < cnt) {
            // the preferred class (q + lane) mod 8 or, if it has run out, the next live one (entries remain: q < cnt)
            const uint32_t want = (q + lane) & 7u;
            const uint32_t rot = ((live | (live << 8)) >> want) & 0xffu;
            const uint32_t b = (want + (uint32_t)__ffs((int)rot) - 1u) & 7u;
            const uint32_t v = loadSharedU16(listOwn + headGet(heads, b) * LIST_STRIDE);
            const uint32_t next = v >> 10;
            headSet(heads, b, next);
            live = next == LIST_NULL ? live & ~(1u << b) : live;
            const uint32_t idx = (v & 0x3ffu) << ((q & 1u) * 16u);
            if (q & 2u) {
                hi |= idx;
            } else {
                lo |= idx;
            }
            if ((q & 3u) == 3u || q + 1u == cnt) {
                quadOut[(size_t)(q >> 2) * TILE_T] = make_uint2(lo, hi);
                lo = hi = 0u;
            }
        }
    }
}

/// Finalizers + stores of one target (shared with the direct variant) and the neighbour statistics of the warp.
template <bool SOLID, bool CORRECTED, bool XSPH = false, bool DELTA = false>
__device__ __forceinline__ void finishTarget(const DevicePointers& d, const UnitLane& u, const Particle& pi, const Accum& acc) {
    if (XSPH && u.target) { // (compile-time: the sums of the XSph term stay dead registers in the other instantiations)
        storeXsph(d, u.slot, acc.xs);
    }
    if (u.target) {
        const uint32_t slot = u.slot;
        const MaterialDev& mat = c_mats[d.u[U_MATID][slot]];
        double S[5] = { 0., 0., 0., 0., 0. };
        if (SOLID) {
            for (int q = 0; q < 5; ++q) {
                S[q] = d.f[F_S0 + q][slot];
            }
        }
        Derivs out;
        finalizeParticle<SOLID, CORRECTED>(c_prm, mat, acc, pi.h, pi.rho, d.f[F_P][slot], d.f[F_CS][slot], SOLID ? d.f[F_REDUCE][slot] : 1., S,
            out);
        if (DELTA) { // (compile-time, like XSPH: the four sums of the delta-SPH terms are dead registers elsewhere)
            finalizeDeltaSph<SOLID>(acc, out);
        }
        storeDerivs<SOLID, CORRECTED>(d, slot, out);
    }
    neighbourStats(d, acc.cnt, u.target);
}

// ---- kernel A: candidate lists ------------------------------------------------------------------------------------
// Phase 1 of every unit at high occupancy (26 KB of shared memory, 64 registers: 8 CTAs per SM): stages the FP32
// positions of each chunk, runs the conservative filter, schedules the per-lane lists (above) and writes them as one
// contiguous block of the list pool. Every block carries the DESCRIPTOR of the unit's next block (pool offset, quad rows,
// chunk and the six staged record ranges), so the pair-sum kernel only follows the chain: it stages records and spends
// its warps on FP64 work.
constexpr size_t LISTS_SMEM = (size_t)TILE_C * 16 + (size_t)(LIST_CAP + 1) * LIST_ROW_BYTES;
// descriptor of a block: {pool row offset, quad rows, chunk, 6 x {first sorted index, count | stage offset << 16}}, word 15 =
// the pool row offset of the block AFTER it in the chain (or LIST_END): the pair-sum kernel prefetches one block ahead
constexpr int DESC_WORDS = 16;

__global__ void __launch_bounds__(TILE_T, 1024 / TILE_T) k_pair_lists(DevicePointers d, uint32_t maxCells, uint32_t poolRows, double skin) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float4* f4 = reinterpret_cast<float4*>(smemRaw);
    uint16_t* list = reinterpret_cast<uint16_t*>(f4 + TILE_C); // row 0 = counts + "no next block", rows 1.. = entries [entry][lane]
    __shared__ ChunkState csBuf[2];
    __shared__ __align__(8) uint64_t stageBar;
    __shared__ uint32_t sRowBeg[3 * CHUNK_ROWS], sRowEnd[3 * CHUNK_ROWS];
    __shared__ uint32_t sWarp[TILE_T / 32], sOff;

    if (d.listCtl->rebuild == 0u) {
        return; // the lists of the last build are reused
    }
    const GridDev g = *d.grid;
    const uint32_t totalUnits = d.segStart[maxCells];
    const int tid = threadIdx.x;
    // search radius enlarged by the skin of the list reuse (ListCtlDev)
    const float Rhalf = (float)(0.5 * c_prm.kernel_radius * (1. + 2.e-5) * (1. + skin));
    // the FP32 coordinates are relative to the grid origin: absolute rounding error <= 2^-24 * extent per coordinate
    const float slack = (float)(g.extent * 1.e-6);
    const ScanGeometry sg = { (float)g.cell, (float)g.cellZ, (float)(g.extent * 1.e-5), g.unsorted == 0u };
    ChunkCursor cur;
    uint32_t stagePhase = 0;
    if (tid == 0) {
        mbarInit(&stageBar, 1);
    }
    __syncthreads();
    const uint32_t listOwn = smemAddr(list + TILE_T + tid); // this lane's first entry (row 1)
    uint4* pool = reinterpret_cast<uint4*>(d.listPool);
    constexpr uint32_t ROW_Q = LIST_ROW_BYTES / 16; // uint4 per row

    for (uint32_t unit = blockIdx.x; unit < totalUnits; unit += gridDim.x) {
        UnitLane u;
        uint32_t* const first = reinterpret_cast<uint32_t*>(d.unitList + (size_t)unit * 4);
        if (!beginUnit(d, g, unit, tid, Rhalf, slack, sRowBeg, sRowEnd, cur, csBuf[0], u)) {
            if (tid == 0) {
                first[0] = LIST_END;
            }
            continue;
        }
        // thread 0 chains the blocks of the unit: the descriptor of a block goes into the header of its predecessor
        uint32_t* link = first;
        uint32_t* prev = nullptr; // descriptor of the chain's current last block (its word 15 names the block after it)
        bool failed = false;
        int buf = 0;
        while (true) {
            __syncthreads(); // chunk descriptor published; everyone is done with the previous chunk's shared memory
            const ChunkState& cs = csBuf[buf];
            if (cs.used == 0) {
                break;
            }
            if (tid == 0) {
                fenceProxyAsync(); // the buffer was last read through the generic proxy
                uint32_t bytes = 0;
#pragma unroll
                for (int r = 0; r < CHUNK_ROWS; ++r) {
                    bytes += (cs.end[r] - cs.beg[r]) * 16u;
                }
                mbarExpectTx(&stageBar, bytes);
#pragma unroll
                for (int r = 0; r < CHUNK_ROWS; ++r) {
                    const uint32_t n = cs.end[r] - cs.beg[r];
                    if (n > 0) {
                        bulkCopyG2S(f4 + cs.base[r], d.posF + cs.beg[r], n * 16u, &stageBar);
                    }
                }
                nextChunk(sRowBeg, sRowEnd, cur, csBuf[buf ^ 1]); // overlaps with the copies
            }
            mbarWait(&stageBar, stagePhase);
            stagePhase ^= 1;
            buf ^= 1;
            ScanState st = { u.target ? 0 : CHUNK_ROWS, 0u, 0u, false };
            while (true) { // rounds: one block per round (a second round only if some lane's list overflowed)
                const uint32_t lp = scanRows(cs, u, sg, f4, listOwn, st);
                const uint32_t cnt = (lp - listOwn) / LIST_STRIDE;
                reinterpret_cast<unsigned char*>(list)[tid] = (unsigned char)cnt;
                const uint32_t wmax = __reduce_max_sync(0xffffffffu, cnt);
                ChainHeads heads;
                chainList(listOwn, cnt, wmax, heads);
                const bool wmore = __any_sync(0xffffffffu, st.r < CHUNK_ROWS);
                if ((tid & 31) == 0) {
                    sWarp[tid >> 5] = wmax | (wmore ? 0x100u : 0u);
                }
                __syncthreads(); // counts and the warp summaries are complete
                uint32_t rows = 0;
                bool anyMore = false;
#pragma unroll
                for (int w = 0; w < TILE_T / 32; ++w) {
                    rows = max(rows, sWarp[w] & 0xffu);
                    anyMore = anyMore || (sWarp[w] & 0x100u) != 0u;
                }
                if (rows > 0) {
                    const uint32_t quads = (rows + 3u) / 4u;
                    const uint32_t need = 1u + 4u * quads; // pool rows of the block
                    if (tid == 0) {
                        uint32_t off = LIST_FALLBACK;
                        if (!failed) {
                            off = atomicAdd(d.listCursor, need);
                            if (off > poolRows || need > poolRows - off) {
                                failed = true;
                                off = LIST_FALLBACK;
                            }
                        }
                        reinterpret_cast<uint32_t*>(list)[TILE_T / 4] = LIST_END; // byte 128 of row 0: no next block (yet)
                        sOff = off;
                    }
                    __syncthreads();
                    const uint32_t off = sOff;
                    if (off != LIST_FALLBACK) {
                        emitList(listOwn, cnt, wmax, (uint32_t)tid, heads, reinterpret_cast<uint2*>(pool + (size_t)(off + 1u) * ROW_Q) + tid);
                        if (tid < (int)ROW_Q) { // row 0: the counts and the end-of-chain mark
                            pool[(size_t)off * ROW_Q + tid] = reinterpret_cast<const uint4*>(list)[tid];
                        }
                        if (tid == 0) { // descriptor of this block -> unitList (first block) or the header of the previous
                                        // block, which was written in an earlier round
                            link[1] = quads;
                            link[2] = (uint32_t)cs.chunk;
#pragma unroll
                            for (int r = 0; r < CHUNK_ROWS; ++r) {
                                link[3 + 2 * r] = cs.beg[r];
                                link[4 + 2 * r] = (cs.end[r] - cs.beg[r]) | (cs.base[r] << 16);
                            }
                            link[15] = LIST_END; // no block after this one (yet)
                            link[0] = off;
                            if (prev != nullptr) {
                                prev[15] = off;
                            }
                            prev = link;
                            link = reinterpret_cast<uint32_t*>(pool + (size_t)off * ROW_Q) + TILE_T / 4;
                        }
                    }
                }
                if (!anyMore) {
                    break;
                }
                __syncthreads(); // row 0 has been copied out: the lanes may overwrite the counts
            }
        }
        if (tid == 0) {
            if (failed) {
                atomicAdd(&d.listCtl->fallbackUnits, 1u);
                first[0] = LIST_FALLBACK;
            } else if (link == first) {
                first[0] = LIST_END; // no candidate at all
            }
        }
    }
}

// ---- kernel B: pair sums, list-driven --------------------------------------------------------------------------------
// Follows the block chains written by k_pair_lists. The only shared memory is the record stage (72 KB solid: three CTAs
// per SM); the lists stay in the pool and every lane reads its own quads through L2, one quad (four entries) ahead.
// Per block: thread 0 issues the TMA copies of the block's record rows, everyone fetches its entry count and first quads
// (threads 0..15 also the descriptor of the next block), waits on the mbarrier and walks its list; one CTA barrier frees
// the stage. At the end of a chain the first block of the CTA's next unit is issued BEFORE the finalizers / stores of the
// finished unit and the loads of the next unit's targets, so those overlap with the copies.
template <bool SOLID, bool BALSARA, bool DELTA = false, bool STRESSAV = false>
struct SumLayout {
    static constexpr int S = STRESSAV ? REC_SOLID_STRESSAV
                             : DELTA  ? (SOLID ? REC_SOLID_DELTA : REC_FLUID_DELTA)
                                   : SOLID ? (BALSARA ? REC_SOLID_BALSARA : REC_SOLID) : REC_FLUID; // doubles per staged record
    static constexpr uint32_t REC_BYTES = (uint32_t)S * 8u;
    static constexpr bool SWIZZLED = S == REC_SOLID;
    static constexpr size_t bytes = (size_t)TILE_C * REC_BYTES; // 73 728 B solid (three CTAs per SM), 64 512 B fluid, 82 944 B
                                                                // solid with the Balsara factor, 101 376 B solid / 82 944 B
                                                                // fluid with the delta-SPH gradient (two CTAs per SM)
    // what fits: 228 KB of shared memory per SM (per CTA 1 KB reserved + 16 B per thread + 280 B of static arrays), 64 K registers
    static constexpr int BY_SMEM = (int)((228 * 1024) / (bytes + 1024 + 16 * TILE_T + 280));
    static constexpr int BY_REGS = 65536 / (TILE_T * (DELTA || STRESSAV ? 255 : BALSARA && SOLID ? 208 : 168));
    static constexpr int CTAS_PER_SM = BY_SMEM < BY_REGS ? BY_SMEM : BY_REGS;
};

/// One list entry between the two stages of the pair body (sph_math.cuh).
struct PairSlot {
    PairGeom g;
    double vz, rho;  // second half of the {vz, rho} piece stage A loads for the density
    double2 lut;     // table entry {G[k], G[k+1] - G[k]}, in flight between the stages
    uint32_t rec;    // shared address of the staged record (solid: swizzle applied)
};

/// Stage A of staged record k: the position pieces and {vz, rho} (LDS.128 through 32-bit shared addresses), geometry,
/// and the table load. `stage` is 128-byte aligned, so piece c of a swizzled (128-byte) record sits at
/// (record start + swizzle) ^ (c << 4); the other layouts have an odd stride and plain offsets.
template <typename P>
__device__ __forceinline__ void stageA(uint32_t stage, uint32_t k, uint32_t selfIdx, const Particle& pi, const LutPair* lut2, PairSlot& s) {
    double2 a, b, e;
    if (P::SWIZZLED) {
        s.rec = stage + k * 128u + ((k & 7u) << 4);
        a = loadSharedD2(s.rec);
        b = loadSharedD2(s.rec ^ 16u);
        e = loadSharedD2(s.rec ^ 48u);
    } else {
        s.rec = stage + k * P::REC_BYTES;
        a = loadSharedD2(s.rec);
        b = loadSharedD2(s.rec + 16u);
        e = loadSharedD2(s.rec + 48u);
    }
    s.vz = e.x;
    s.rho = e.y;
    pairGeometry(c_prm, pi.x, pi.y, pi.z, pi.h, pi.rho, a.x, a.y, b.x, b.y, e.y, s.g);
    s.g.valid = s.g.valid && k != selfIdx;
    s.lut = __ldg(reinterpret_cast<const double2*>(lut2) + s.g.k);
}

/// Stage B: the rest of the record and the sums.
template <bool SOLID, bool CORRECTED, bool FILTER, bool BALSARA, bool XSPH, bool DELTA, bool STRESSAV>
__device__ __forceinline__ void stageB(const PairSlot& s, const Particle& pi, Accum& acc, const LutPair* lutW2) {
    using P = SumLayout<SOLID, BALSARA, DELTA, STRESSAV>;
    Particle pj;
    pj.vz = s.vz;
    pj.rho = s.rho;
    pj.bal = 0.;
    if (SOLID) {
        double2 c, f, g, s1, s2;
        if (P::SWIZZLED) {
            c = loadSharedD2(s.rec ^ 32u);
            f = loadSharedD2(s.rec ^ 64u);
            g = loadSharedD2(s.rec ^ 80u);
            s1 = loadSharedD2(s.rec ^ 96u);
            s2 = loadSharedD2(s.rec ^ 112u);
        } else {
            c = loadSharedD2(s.rec + 32u);
            f = loadSharedD2(s.rec + 64u);
            g = loadSharedD2(s.rec + 80u);
            s1 = loadSharedD2(s.rec + 96u);
            s2 = loadSharedD2(s.rec + 112u);
            if (BALSARA) {
                pj.bal = loadSharedD2(s.rec + 128u).x;
            }
        }
        pj.vx = c.x; pj.vy = c.y;
        pj.P = f.x;
        unpackCsGroup(f.y, pj.cs, pj.grp);
        pj.vol = g.x;
        pj.Sr[0] = g.y; pj.Sr[1] = s1.x; pj.Sr[2] = s1.y; pj.Sr[3] = s2.x; pj.Sr[4] = s2.y;
    } else {
        const double2 c = loadSharedD2(s.rec + 32u), f = loadSharedD2(s.rec + 64u), g = loadSharedD2(s.rec + 80u);
        pj.vx = c.x; pj.vy = c.y;
        pj.P = f.x; pj.cs = f.y; pj.vol = g.x;
        pj.grp = 0;
        if (BALSARA) {
            pj.bal = loadSharedD2(s.rec + 96u).x;
        }
    }
    pj.m = pj.vol * pj.rho;
    if (DELTA) { // the density gradient of the previous evaluation: the two pieces behind the regular ones (odd stride, no swizzle)
        const double2 ga = loadSharedD2(s.rec + (SOLID ? 128u : 96u)), gb = loadSharedD2(s.rec + (SOLID ? 144u : 112u));
        pj.gr[0] = ga.x; pj.gr[1] = ga.y; pj.gr[2] = gb.x;
    }
    if (STRESSAV) { // as / rho^2: the three pieces behind the regular ones
        const double2 aa = loadSharedD2(s.rec + 128u), ab = loadSharedD2(s.rec + 144u), ac = loadSharedD2(s.rec + 160u);
        pj.as[0] = aa.x; pj.as[1] = aa.y; pj.as[2] = ab.x; pj.as[3] = ab.y; pj.as[4] = ac.x; pj.as[5] = ac.y;
    }
    double W = 0.;
    if (XSPH || STRESSAV) { // kernel value (XSph term, artificial stress): same index and weight as the gradient table
        const double2 w = __ldg(reinterpret_cast<const double2*>(lutW2) + s.g.k);
        W = fma(s.g.ratio, w.y, w.x);
    }
    pairSums<SOLID, CORRECTED, FILTER, BALSARA, XSPH, DELTA, STRESSAV>(c_prm, pi, pj, s.g, fma(s.g.ratio, s.lut.y, s.lut.x), acc, W);
}

template <bool SOLID, bool CORRECTED, bool FILTER, bool BALSARA, bool XSPH, bool DELTA, bool STRESSAV>
__global__ void __launch_bounds__(TILE_T, (SumLayout<SOLID, BALSARA, DELTA, STRESSAV>::CTAS_PER_SM)) k_pair_sum(DevicePointers d, uint32_t maxCells) {
    using P = SumLayout<SOLID, BALSARA, DELTA, STRESSAV>;
    extern __shared__ __align__(128) unsigned char smemStage[];
    __shared__ __align__(8) uint64_t stageBar;
    __shared__ uint32_t sFirst[2][16]; // first-block descriptors of the CTA's current and next unit
    __shared__ uint32_t sDesc[2][16];  // descriptor of the chain's next block, double-buffered over the blocks
    // per-thread landing slots of the asynchronous prefetches (no registers are held while the loads are in flight):
    __shared__ uint2 sPfQuad[TILE_T];     // first quad of the block this thread processes next
    __shared__ uint32_t sPfCnt[TILE_T];   // the word of that block's count row that holds this thread's entry count
    __shared__ uint32_t sPfWord[TILE_T];  // this thread's lane word in the CTA's next unit

    const GridDev g = *d.grid;
    const uint32_t totalUnits = d.segStart[maxCells];
    const uint32_t stride = gridDim.x;
    const int tid = threadIdx.x;
    const unsigned char* pool = reinterpret_cast<const unsigned char*>(d.listPool);
    const uint32_t stage = smemAddr(smemStage);

    uint32_t unit = blockIdx.x;
    if (unit >= totalUnits) {
        return;
    }
    auto loadFirst = [&](uint32_t v, int ring) { // threads 0..15: descriptor of unit v's first block -> sFirst[ring]
        if (tid < 16) {
            sFirst[ring][tid] = v < totalUnits ? reinterpret_cast<const uint32_t*>(d.unitList + (size_t)v * 4)[tid] : LIST_END;
        }
    };
    auto issue = [&](const uint32_t* dsc) { // thread 0: TMA copies of the record rows of the block described by dsc
        fenceProxyAsync(); // the stage was last read through the generic proxy
        uint32_t bytes = 0;
#pragma unroll
        for (int r = 0; r < CHUNK_ROWS; ++r) {
            bytes += (dsc[4 + 2 * r] & 0xffffu) * P::REC_BYTES;
        }
        mbarExpectTx(&stageBar, bytes);
#pragma unroll
        for (int r = 0; r < CHUNK_ROWS; ++r) {
            const uint32_t nb = dsc[4 + 2 * r], n = nb & 0xffffu;
            if (n > 0) {
                bulkCopyG2S(smemStage + (size_t)(nb >> 16) * P::REC_BYTES, d.rec + (size_t)dsc[3 + 2 * r] * P::S, n * P::REC_BYTES, &stageBar);
            }
        }
    };
    auto laneWord = [&](uint32_t v) -> uint32_t { // lane order word of thread tid in unit v (k_unit_prep), idle: ~0
        const uint4 desc = d.unitDesc[v];
        return (uint32_t)tid < (desc.w >> 8) ? d.unitLane[d.unitAux[v].x + tid] : 0xffffffffu;
    };
    auto prefetchLaneWord = [&](uint32_t v) { // the same word for the CTA's NEXT unit, into sPfWord[tid]
        sPfWord[tid] = 0xffffffffu;
        if (v < totalUnits) {
            const uint4 desc = d.unitDesc[v];
            if ((uint32_t)tid < (desc.w >> 8)) {
                copyAsync4(&sPfWord[tid], d.unitLane + d.unitAux[v].x + tid);
            }
        }
    };
    auto setupUnit = [&](uint32_t v, uint32_t word, UnitLane& u, Particle& pi, Accum& acc) {
        const uint32_t dr = d.unitDesc[v].x;
        u.cy = (int)(dr % (uint32_t)g.dim[1]);
        u.k = (int)(dr / (uint32_t)g.dim[1]);
        u.live = word != 0xffffffffu;
        u.t = u.live ? (word & 0x3fffffffu) : 0u;
        u.upper = u.live && (word >> 31) != 0u;
        u.target = u.live && (word & 0x40000000u) == 0u; // ghosts are neighbours only
        u.slot = u.target ? d.order[u.t] : 0xffffffffu;
        if (u.live) {
            loadRecord<SOLID>(d.rec, u.t, P::S, (DELTA ? SPHGPU_FLAG_DELTASPH : 0u) | (STRESSAV ? SPHGPU_FLAG_STRESS_AV : 0u), pi);
            if (STRESSAV) {
                pi.wpInv = u.target ? 1. / d.f[F_WP][u.slot] : 0.;
            }
        } else {
            pi.x = pi.y = pi.z = 0.;
            pi.h = 1.;
        }
        if (g.nLarge > 0u && u.target) {
            acc = d.accLarge[u.t]; // sums over the large neighbours (two-level radii, pair.cu: k_large_neighbours)
        } else {
            accumZero(acc);
        }
    };

    if (tid == 0) {
        mbarInit(&stageBar, 1);
    }
    int ring = 0, par = 0;
    loadFirst(unit, 0);
    loadFirst(unit + stride, 1);
    UnitLane u;
    Particle pi;
    Accum acc;
    setupUnit(unit, laneWord(unit), u, pi, acc);
    prefetchLaneWord(unit + stride); // one unit ahead
    __syncthreads();
    // This lane's share of the block it processes NEXT is fetched from the pool (through L2) as soon as that block's
    // offset is known -- one block ahead: entry count and first quad.
    auto prefetchBlock = [&](uint32_t off) {
        const unsigned char* blk = pool + (size_t)off * LIST_ROW_BYTES;
        copyAsync4(&sPfCnt[tid], blk + (tid & ~3));
        copyAsync8(&sPfQuad[tid], reinterpret_cast<const uint2*>(blk + LIST_ROW_BYTES) + tid);
    };
    if (sFirst[0][0] < LIST_FALLBACK) {
        prefetchBlock(sFirst[0][0]);
    }
    uint32_t stagePhase = 0;
    bool firstOfUnit = true; // the next block is the first one of `unit` (its descriptor is sFirst[ring])
    bool issued = false;     // ... and its copies are already in flight
    // Invariant at the top of the loop: a CTA barrier has passed since the descriptor read here was written and since
    // every thread finished with the stage; pfCnt / pfQuad belong to the block that descriptor names.
    while (true) {
        const uint32_t* dsc = firstOfUnit ? sFirst[ring] : sDesc[par];
        const uint32_t off = dsc[0];
        if (off >= LIST_FALLBACK) { // the chain of `unit` has ended (or the unit has no chain)
            const bool fallback = firstOfUnit && off == LIST_FALLBACK; // left to k_pair_fallback
            const uint32_t nextUnit = unit + stride;
            const int nextRing = ring ^ 1;
            __syncthreads(); // everyone has read sFirst[ring]: its slot may be refilled below
            const uint32_t nextOff = nextUnit < totalUnits ? sFirst[nextRing][0] : LIST_END;
            issued = nextOff < LIST_FALLBACK;
            if (issued) {
                if (tid == 0) {
                    issue(sFirst[nextRing]); // the next unit's first block, right away
                }
                prefetchBlock(nextOff);
            }
            if (!fallback) {
                finishTarget<SOLID, CORRECTED, XSPH, DELTA>(d, u, pi, acc); // overlaps with the copies
            }
            if (nextUnit >= totalUnits) {
                break;
            }
            loadFirst(nextUnit + stride, ring); // ring slot of the finished unit; visible after the next barrier
            unit = nextUnit;
            ring = nextRing;
            copyAsyncWaitAll(); // (also completes the block prefetch issued above; it has had the finalizers' time)
            setupUnit(unit, sPfWord[tid], u, pi, acc);
            prefetchLaneWord(unit + stride);
            firstOfUnit = true;
            continue;
        }
        // descriptor words every thread needs: chunk and the staged row that may hold the target's own record
        const int selfRow = u.upper ? 4 : 1;
        const uint32_t chunk = dsc[2], selfBeg = dsc[3 + 2 * selfRow], selfNB = dsc[4 + 2 * selfRow];
        if (!issued && tid == 0) {
            issue(dsc);
        }
        issued = false;
        firstOfUnit = false;
        copyAsyncWaitAll();
        const uint32_t cnt = (sPfCnt[tid] >> (8 * (tid & 3))) & 0xffu;
        uint2 cur = sPfQuad[tid];
        if (dsc[15] < LIST_FALLBACK) {
            prefetchBlock(dsc[15]); // the chain's next block: in flight during the copies and the walk
        }
        if (tid < 16) { // the descriptor of that block, from the header of this one
            copyAsync4(&sDesc[par ^ 1][tid], reinterpret_cast<const uint32_t*>(pool + (size_t)off * LIST_ROW_BYTES) + TILE_T / 4 + tid);
        }
        const uint2* quads = reinterpret_cast<const uint2*>(pool + (size_t)(off + 1u) * LIST_ROW_BYTES) + tid;
        mbarWait(&stageBar, stagePhase);
        stagePhase ^= 1;
        if (cnt > 0u) {
            const bool selfHere = chunk == 2u && u.t >= selfBeg && u.t < selfBeg + (selfNB & 0xffffu);
            const uint32_t selfIdx = selfHere ? (selfNB >> 16) + (u.t - selfBeg) : 0xffffffffu;
            // One quad (four entries) per trip of the loop, the next quad fetched at its top and first used by the last
            // stage A of the trip. Stage A runs one entry ahead of stage B and is unconditional (past the end of the list
            // it runs on whatever the exhausted quads hold and its result is dropped): every [A; B] pair below is ONE
            // basic block, so that the scheduler interleaves the two stages.
            PairSlot s0, s1;
            stageA<P>(stage, cur.x & 0xffffu, selfIdx, pi, d.lut2, s0);
            uint32_t q = 0;
            while (true) {
                quads += TILE_T;
                uint2 nxt = make_uint2(0u, 0u);
                loadGlobalU2If(q + 4u < cnt, quads, nxt);
                stageA<P>(stage, cur.x >> 16, selfIdx, pi, d.lut2, s1);
                stageB<SOLID, CORRECTED, FILTER, BALSARA, XSPH, DELTA, STRESSAV>(s0, pi, acc, d.lutW2);
                if (++q >= cnt) {
                    break;
                }
                stageA<P>(stage, cur.y & 0xffffu, selfIdx, pi, d.lut2, s0);
                stageB<SOLID, CORRECTED, FILTER, BALSARA, XSPH, DELTA, STRESSAV>(s1, pi, acc, d.lutW2);
                if (++q >= cnt) {
                    break;
                }
                stageA<P>(stage, cur.y >> 16, selfIdx, pi, d.lut2, s1);
                stageB<SOLID, CORRECTED, FILTER, BALSARA, XSPH, DELTA, STRESSAV>(s0, pi, acc, d.lutW2);
                if (++q >= cnt) {
                    break;
                }
                stageA<P>(stage, nxt.x & 0xffffu, selfIdx, pi, d.lut2, s0);
                stageB<SOLID, CORRECTED, FILTER, BALSARA, XSPH, DELTA, STRESSAV>(s1, pi, acc, d.lutW2);
                if (++q >= cnt) {
                    break;
                }
                cur = nxt;
            }
        }
        copyAsyncWaitAll(); // the next descriptor (threads 0..15) has landed
        par ^= 1;
        __syncthreads(); // the stage is free and the next descriptor is visible
    }
}

// ---- fallback: direct evaluation of whole units ------------------------------------------------------------------------
// Units whose candidate lists did not fit the list pool (the pool is then enlarged for the following steps, api.cu), and
// every unit in variant 2: one thread per target of the unit, candidates streamed from the sorted records (pair.cu).
template <bool SOLID, bool CORRECTED, bool FILTER>
__global__ void __launch_bounds__(TILE_T) k_pair_fallback(DevicePointers d, uint32_t maxCells, bool allUnits) {
    if (!allUnits && d.listCtl->fallbackUnits == 0u) {
        return;
    }
    const uint32_t totalUnits = d.segStart[maxCells];
    const int tid = threadIdx.x;
    for (uint32_t unit = blockIdx.x; unit < totalUnits; unit += gridDim.x) {
        if (!allUnits && d.unitList[(size_t)unit * 4].x != LIST_FALLBACK) {
            continue;
        }
        const uint4 desc = d.unitDesc[unit];
        const uint4 aux = d.unitAux[unit];
        if (aux.z == 0u || (uint32_t)tid >= (desc.w >> 8)) {
            continue; // ghost-only unit / idle lane
        }
        const uint32_t word = d.unitLane[aux.x + tid];
        if (word & 0x40000000u) {
            continue; // ghosts are neighbours only
        }
        const uint32_t t = word & 0x3fffffffu;
        directTarget<SOLID, CORRECTED, FILTER>(d, t, d.order[t]);
    }
}

int launchSegments(sphgpu_ctx* ctx) {
    cudaStream_t st = ctx->stream;
    const uint32_t total = ctx->maxCells + 1;
    int smsU = 148;
    cudaDeviceGetAttribute(&smsU, cudaDevAttrMultiProcessorCount, ctx->device);
    SPH_CUDA_CHECK(cudaMemsetAsync(ctx->d.cellCount, 0, sizeof(uint32_t) * total, st));
    k_units<false><<<smsU * 16, 128, 0, st>>>(ctx->d, ctx->maxCells);
    k_scan_block<<<ctx->scanBlocks, 512, 0, st>>>(ctx->d.cellCount, ctx->d.segStart, ctx->d.scanBlock, total, ctx->d.listCtl);
    k_scan_sums<<<1, 1024, 0, st>>>(ctx->d.scanBlock, ctx->scanBlocks, ctx->d.listCtl);
    k_scan_add<<<ctx->scanBlocks, 512, 0, st>>>(ctx->d.segStart, ctx->d.scanBlock, total, ctx->d.listCtl);
    k_units<true><<<smsU * 16, 128, 0, st>>>(ctx->d, ctx->maxCells);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const uint32_t upper = ctx->nActive / (TILE_T / 2) + ctx->maxCells + 1; // >= number of units
    k_unit_prep<<<(uint32_t)std::min<uint64_t>((uint64_t)sms * 16, upper), TILE_T, 0, st>>>(ctx->d, ctx->n, ctx->maxCells);
    ctx->launches += 6;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

static uint32_t unitGrid(sphgpu_ctx* ctx, int ctasPerSm, int waves) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const uint32_t upper = ctx->nActive / (TILE_T / 2) + ctx->maxCells + 1; // >= number of units
    return (uint32_t)std::min<uint64_t>((uint64_t)sms * ctasPerSm * waves, std::max<uint32_t>(upper, 1u));
}

static int launchLists(sphgpu_ctx* ctx) {
    SPH_CUDA_CHECK(cudaMemsetAsync(ctx->d.listCursor, 0, sizeof(uint32_t), ctx->stream));
    // variant 3 (tests): a pool of a few blocks only, so that most units take the fallback path
    const uint32_t poolRows = ctx->variant == 3 ? std::min<uint32_t>(ctx->poolRows, 4096u) : ctx->poolRows;
    static bool configured[64] = {}; // per device: the attribute is per device function
    if (!configured[ctx->device & 63]) {
        SPH_CUDA_CHECK(cudaFuncSetAttribute(k_pair_lists, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LISTS_SMEM));
        configured[ctx->device & 63] = true;
    }
    k_pair_lists<<<unitGrid(ctx, 1024 / TILE_T, 8), TILE_T, LISTS_SMEM, ctx->stream>>>(ctx->d, ctx->maxCells, poolRows, ctx->listSkin > 0. ? ctx->listSkin : 0.);
    ctx->launches += 1;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

template <bool SOLID, bool CORRECTED, bool FILTER, bool BALSARA, bool XSPH, bool DELTA = false, bool STRESSAV = false>
static int launchSumVariantB(sphgpu_ctx* ctx) {
    auto kernel = k_pair_sum<SOLID, CORRECTED, FILTER, BALSARA, XSPH, DELTA, STRESSAV>;
    static bool configured[64] = {}; // per instantiation and device; the attribute is per device function
    const size_t smem = SumLayout<SOLID, BALSARA, DELTA, STRESSAV>::bytes;
    if (!configured[ctx->device & 63]) {
        SPH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SPH_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured[ctx->device & 63] = true;
    }
    // many more CTAs than fit at once: the block scheduler balances the load (measured: 32 waves beat 8 by 1 %, 1 by 4 %)
    kernel<<<unitGrid(ctx, SumLayout<SOLID, BALSARA, DELTA, STRESSAV>::CTAS_PER_SM, 32), TILE_T, smem, ctx->stream>>>(ctx->d, ctx->maxCells);
    ctx->launches += 1;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

template <bool SOLID, bool CORRECTED, bool FILTER>
static int launchSumVariant(sphgpu_ctx* ctx) {
    if (ctx->xsph) { // (not together with the Balsara switch: rejected by sphgpu_create)
        return launchSumVariantB<SOLID, CORRECTED, FILTER, false, true>(ctx);
    }
    if (ctx->deltasph) { // (not together with the Balsara switch or the XSph term either)
        return launchSumVariantB<SOLID, CORRECTED, FILTER, false, false, true>(ctx);
    }
    if (SOLID && ctx->stressAv) { // (solids only, and on its own as well)
        return launchSumVariantB<SOLID, CORRECTED, FILTER, false, false, false, SOLID>(ctx);
    }
    return ctx->balsara ? launchSumVariantB<SOLID, CORRECTED, FILTER, true, false>(ctx) : launchSumVariantB<SOLID, CORRECTED, FILTER, false, false>(ctx);
}

template <bool SOLID, bool CORRECTED, bool FILTER>
static int launchFallback(sphgpu_ctx* ctx, bool allUnits) {
    // (all units: a full grid; otherwise a small one -- the kernel returns at once unless the list pool overflowed)
    k_pair_fallback<SOLID, CORRECTED, FILTER><<<unitGrid(ctx, allUnits ? 8 : 2, allUnits ? 4 : 1), TILE_T, 0, ctx->stream>>>(ctx->d, ctx->maxCells,
        allUnits);
    ctx->launches += 1;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

template <bool SOLID, bool CORRECTED, bool FILTER>
static int launchPairKernels(sphgpu_ctx* ctx) {
    if (ctx->variant == 2) { // every unit through the direct per-target loop (cross-check of the unit bookkeeping)
        return launchFallback<SOLID, CORRECTED, FILTER>(ctx, true);
    }
    SPH_CUDA_CHECK(cudaEventRecord(ctx->evPair[1], ctx->stream));
    int rc = launchLists(ctx);
    SPH_CUDA_CHECK(cudaEventRecord(ctx->evPair[2], ctx->stream));
    if (rc == SPHGPU_OK) {
        rc = launchSumVariant<SOLID, CORRECTED, FILTER>(ctx);
    }
    if (rc == SPHGPU_OK) { // returns at once unless the list pool overflowed
        rc = launchFallback<SOLID, CORRECTED, FILTER>(ctx, false);
    }
    SPH_CUDA_CHECK(cudaEventRecord(ctx->evPair[3], ctx->stream));
    ctx->pairTimed = true;
    return rc;
}

/// variant 0: candidate lists (k_pair_lists) + list-driven pair sums (k_pair_sum); variant 2: k_pair_fallback only;
/// variant 3: as 0 with a deliberately tiny list pool (most units overflow).
int launchPairTiled(sphgpu_ctx* ctx) {
    ctx->pairTimed = false;
    SPH_CUDA_CHECK(cudaEventRecord(ctx->evPair[0], ctx->stream));
    int rc = launchSegments(ctx);
    if (rc != SPHGPU_OK) {
        return rc;
    }
    if (!ctx->solid) {
        return launchPairKernels<false, false, false>(ctx);
    }
    if (ctx->corrected) {
        return ctx->filter ? launchPairKernels<true, true, true>(ctx) : launchPairKernels<true, true, false>(ctx);
    }
    return ctx->filter ? launchPairKernels<true, false, true>(ctx) : launchPairKernels<true, false, false>(ctx);
}

} // namespace sph
