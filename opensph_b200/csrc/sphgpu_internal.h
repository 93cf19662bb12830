// Internal declarations of libsphgpu: device-resident SoA state, grid, and kernel launchers.
#pragma once
#include "../../include/sphgpu.h"
#include "sph_math.cuh"
#include <cuda_runtime.h>
#include <string>

namespace sph {

// ---- device-resident particle state (structure of arrays, FP64) ------------------------------------------
// One plane of `capacity` doubles per field. Particles keep the caller's ("slot") order; the per-step cell sort
// only produces an index permutation plus compact, sorted neighbour-input planes (see Sorted below).
enum Field : int {
    F_X, F_Y, F_Z, F_H,           // POSITION value {x,y,z,h}
    F_VX, F_VY, F_VZ, F_VH,       // POSITION dt {v, dh/dt}
    F_AX, F_AY, F_AZ,             // POSITION d2t (h lane is identically 0)
    F_M,
    F_RHO, F_DRHO,
    F_U, F_DU,
    F_P, F_CS,
    F_S0, F_S1, F_S2, F_S3, F_S4,       // DEVIATORIC_STRESS {xx,yy,xy,xz,yz}
    F_DS0, F_DS1, F_DS2, F_DS3, F_DS4,
    F_REDUCE,
    F_D, F_DD,
    F_EPSMIN, F_MZERO, F_GROWTH,
    F_DIVV,
    F_GV0, F_GV1, F_GV2, F_GV3, F_GV4, F_GV5, // VELOCITY_GRADIENT {xx,yy,zz,xy,xz,yz}
    F_C0, F_C1, F_C2, F_C3, F_C4, F_C5,       // STRAIN_RATE_CORRECTION_TENSOR
    // PredictorCorrector "predictions": copies of the highest derivatives (TimeStepping.cpp:272-282)
    F_AXP, F_AYP, F_AZP, F_DRHOP, F_DUP, F_DSP0, F_DSP1, F_DSP2, F_DSP3, F_DSP4, F_DDP,
    F_ROTX, F_ROTY, F_ROTZ,       // VELOCITY_ROTATION (Balsara switch): result of the last evaluation, input of the next
    F_XSX, F_XSY, F_XSZ,          // XSPH_VELOCITIES (XSph term): the correction currently contained in the velocities
    F_DGX, F_DGY, F_DGZ,          // DELTASPH_DENSITY_GRADIENT (delta-SPH terms): result of the last evaluation, input of the next
    F_AS0, F_AS1, F_AS2, F_AS3, F_AS4, F_AS5, // AV_STRESS {xx,yy,zz,xy,xz,yz} (artificial stress): written by every prologue
    F_WP,                         // INTERPARTICLE_SPACING_KERNEL (artificial stress): uploaded once
    F_COUNT
};

enum UField : int { U_NFLAWS, U_FLAG, U_MATID, U_NCNT, U_COUNT };

// Sorted neighbour-input records (array of structures, rebuilt every integrate() in cell order). One record is what a
// particle contributes as a neighbour, as 16-byte PIECES {x,y | z,h | vx,vy | vz,rho | P,cs* | vol,Sr0 | Sr1,Sr2 |
// Sr3,Sr4} with P = p/rho^2, Sr = S/rho^2, vol = m/rho (m = vol * rho is rebuilt by the loaders) and cs* = the sound
// speed with the group id (body flag + 1, 0 = fully damaged) in its 8 low mantissa bits (sph_math.cuh: packCsGroup).
// Solid: 8 pieces = 128 bytes; piece c of the record with sorted index t sits at position c ^ (t & 7) inside the record
// (XOR swizzle), so that the lanes of a quarter warp that gather records with different t mod 8 hit different bank
// groups of shared memory although the stride is a multiple of 128 bytes. Staged copies keep t mod 8 (the chunk
// builder aligns every staged row piece accordingly), so the same loader serves global and shared memory.
// Fluid: the first 6 pieces + 16 bytes of padding = 7 pieces (odd stride, no swizzle needed).
// With the Balsara switch every particle also contributes its factor f: a fluid record keeps it in the padding piece
// {f, -}; a solid record grows to nine pieces = 144 bytes {... | f, -} with an odd stride and therefore no swizzle.
// With the delta-SPH terms every particle contributes the density gradient G of the previous evaluation, two more pieces
// {Gx,Gy | Gz,-} behind the regular ones plus one piece of padding to keep the stride odd: solid 11 pieces = 176 bytes,
// fluid 9 pieces = 144 bytes.
// With the artificial stress (solids only) every particle contributes as / rho^2, three more pieces {xx,yy | zz,xy | xz,yz}:
// 11 pieces = 176 bytes as well (the two layouts never coexist; the run flags tell them apart).
constexpr int REC_SOLID = 16;         // doubles per record, solid
constexpr int REC_FLUID = 14;         // doubles per record, fluid
constexpr int REC_SOLID_BALSARA = 18; // doubles per record, solid with the Balsara switch
constexpr int REC_SOLID_DELTA = 22;   // doubles per record, solid with the delta-SPH terms
constexpr int REC_FLUID_DELTA = 18;   // doubles per record, fluid with the delta-SPH terms
constexpr int REC_SOLID_STRESSAV = 22; // doubles per record, solid with the artificial stress

__host__ __device__ inline int recordDoubles(bool solid, bool balsara, bool delta = false, bool stressAv = false) {
    return stressAv ? REC_SOLID_STRESSAV
                    : delta ? (solid ? REC_SOLID_DELTA : REC_FLUID_DELTA) : solid ? (balsara ? REC_SOLID_BALSARA : REC_SOLID) : REC_FLUID;
}
/// Record size from the run flags (SPHGPU_FLAG_BALSARA / SPHGPU_FLAG_DELTASPH exclude each other).
__host__ __device__ inline int recordDoublesOf(bool solid, uint32_t flags) {
    return recordDoubles(solid, (flags & SPHGPU_FLAG_BALSARA) != 0, (flags & SPHGPU_FLAG_DELTASPH) != 0, (flags & SPHGPU_FLAG_STRESS_AV) != 0);
}
/// XOR swizzle of the record with sorted / staged index t (0 for the layouts with an odd stride).
__host__ __device__ inline uint32_t recordSwizzle(int recDoubles, uint32_t t) {
    return recDoubles == REC_SOLID ? (t & 7u) : 0u;
}

struct GridDev {
    double lo[3];
    double cell, cellInv;   // cell edge in x and y: R * h_max (x 1+1e-6), possibly enlarged to cap the cell count
    double cellZ, cellZInv; // cells are HALF as high in z (cell / 2): neighbours span +-1 cell in x,y and +-2 in z
    int dim[3];
    uint32_t ncells;
    double hmax;       // largest h of the particles in the cell list (the "small" ones when the radii are split, see hSplit)
    double hmaxAll;    // largest h of all particles
    // Two-level search radii (the reference's answer to a few huge particles is RadiiHashMap, AsymmetricSolver.cpp:14-56):
    // particles with h > hSplit ("large": at most LARGE_MAX of them, else +inf and everything is small) are kept out of
    // the cell list -- they sit in one overflow cell behind the last real cell -- so the cell edge follows the largest
    // SMALL h; every pair with a large particle is evaluated by k_large_targets / k_large_neighbours (pair.cu).
    double hSplit;
    uint32_t nLarge, largeBegin; // number of large particles and their first sorted index
    double extent;     // largest edge of the grid box: scale of the FP32 rounding of the grid-relative coordinates
    uint32_t unsorted; // set by k_sort_cells when a cell was too large to be ordered by x (windows then span whole rows)
};

struct StatsDev {
    unsigned int neighMin, neighMax;
    unsigned long long pairCount;
    unsigned int fallbackUnits; // units whose candidate lists did not fit the list pool (handled by k_pair_fallback)
    unsigned int badFlags;      // particles whose body flag does not fit the record's group field
};

// Reuse of the cell list, the work units and the candidate lists over several steps (Verlet lists). The lists are built
// with the search radius enlarged by (1 + skin); they stay valid (conservative supersets; the exact predicate is applied
// by the pair-sum kernel every step) as long as no pair that was left out can have become a neighbour:
//     max_(i,j) |u_i - u_j| / (R h_min0)  +  max_i (h_i / h_i0 - 1)  <  skin,      u_i = r_i - r_i0,
// with r_i0, h_i0 the values at build time (pos0) and the first maximum over all pairs that could reach each other. What
// matters is the RELATIVE displacement of nearby particles -- a body that translates, rotates or expands smoothly keeps
// its lists for a long time. k_bounds records the bounding box of u over the particles of every block of 4 x 2 x 4
// (build-time) cells; k_disp_check bounds |u_i - u_j| for every block against the 27 blocks around it -- which contain the
// candidate stencil plus at least one more cell edge --; pairs further apart than that are covered by the global extent
// of u staying below 0.9 cell edges. k_grid_decide decides.
struct ListCtlDev {
    uint32_t rebuild;       // decision of the current integrate(): the build kernels run (1) or return at once (0)
    uint32_t age;           // integrate() calls served by the current lists after the one that built them
    uint32_t rebuilds;      // builds so far
    uint32_t fallbackUnits; // units whose candidate lists did not fit the list pool at the last build
    double lastMetric;      // 2 max ratio + max growth seen by the last call
    double hSplit, hmaxAll; // k_grid_decide -> k_hmax_small -> k_grid_params (split of the search radii, GridDev)
    // halo guard (halo.cu): smallest head-room of an interior particle towards a cut plane, in units of its reach
    // R (h_i + h_max) / 2; negative = an interior particle has come within reach of the other rank: neighbours are missing
    int haloMarginBits;     // bit pattern of a non-negative float (atomicMin), reset to +inf by every exchange
    uint32_t haloViolation; // sticky until the halo is configured again
    uint32_t peerTimeout;   // a wait on a peer's mailbox gave up (halo.cu: waitForSeq)
    uint32_t pad;
};

struct TimestepDev {
    unsigned long long minBits[4]; // Courant, Derivative, Acceleration, Divergence (bit patterns of positive doubles)
};

struct StepStateDev { // time-step bookkeeping of sphgpu_run_pc (MultiCriterion::compute on the device)
    double dt;          // step the next predict / correct will use
    double lastDt;      // MultiCriterion::lastStep
    uint32_t lastDtInit, pad;
    double dtPrev;      // the step before `dt`: what the corrector half of k_correct_predict uses
};

struct StepRecordDev {
    double dt;          // step chosen after this PredictorCorrector step
    uint32_t criterion, pad;
};

struct DevicePointers {
    double* f[F_COUNT];
    uint32_t* u[U_COUNT];
    double* rec;        // sorted neighbour-input records, REC_SOLID or REC_FLUID doubles each
    uint32_t* sCell;    // sorted: linear cell index
    float4* posF;       // sorted: FP32 {x, y, z} relative to the grid origin and h (conservative pre-filter of the pair kernel)
    float4* pos0;       // by slot: the same at the time the lists were built (displacement check of the list reuse)
    Accum* accLarge;    // sorted: partial sums of the small targets over their LARGE neighbours (two-level radii)
    Accum* largePartial;    // [LARGE_MAX] per-slice partial sums of the large targets (k_large_targets)
    uint32_t* largeCounter; // [LARGE_MAX] slices of a large target that have finished
    ListCtlDev* listCtl;
    uint32_t* cellHmax; // [maxCells] bit pattern of the largest (float) h inside each cell
    // displacement boxes of the list reuse (ListCtlDev), one per block of 4 x 2 x 4 cells: order-preserving uint keys of
    // floats, plane k of (maxCells + 1) entries = lower (k = 0..2) / upper (k = 3..5) bound of u_x, u_y, u_z
    uint32_t* dispA;
    uint32_t* dispGlobal; // [8] global lower / upper keys of u, [6] float bits of the largest relative displacement bound
    uint32_t* order;    // sorted position -> slot
    uint32_t* cellOf;   // slot -> cell
    uint32_t* rank;     // slot -> rank inside its cell during the build, afterwards slot -> sorted position (inverse of order)
    uint32_t* cellStart; // [maxCells + 1] exclusive prefix of counts
    uint32_t* cellCount; // [maxCells + 1]
    uint32_t* scanBlock; // block sums of the scan
    uint32_t* segStart;  // [maxCells + 1] exclusive prefix of the number of work units per double row
    uint4* unitDesc;     // [maxSegs] work units of the tiled pair kernel: {double row, cA, skip, (cB - cA) | targets << 8}
    uint4* unitAux;      // [maxSegs] {first entry in unitLane, bits of the largest candidate h, any owned target, -}
    uint4* unitList;     // [4 * maxSegs] descriptor of every unit's first list block (pair_tiled.cu)
    unsigned char* listPool; // candidate lists written by k_pair_lists, blocks of 256-byte rows
    uint32_t* listCursor;    // bump allocator of the pool (rows)
    uint32_t* unitLane;  // [capacity] lane order of every unit: sorted index | upper row << 31 | ghost << 30
    double* boundsPartial; // [BOUNDS_BLOCKS * 8]
    const double* lut;          // (dW/dq)/q table of the reference, lut_entries + 2 doubles (direct variant)
    const LutPair* lut2;        // the same table as {G[k], G[k+1] - G[k]} pairs, lut_entries + 1 entries (tiled kernels)
    const double* lutW;         // kernel values W(q^2), lut_entries + 2 doubles (XSph term, direct variant)
    const LutPair* lutW2;       // the same as pairs (XSph term, tiled kernels)
    GridDev* grid;
    StatsDev* stats;
    StatsDev* statsInit;       // constant {min = ~0, 0, ...}: copied over stats at the start of every integrate()
    TimestepDev* tsd;
    StepStateDev* stepState;   // allocated once
    const double* dtDev;       // != null: k_predict / k_correct take dt from here (sphgpu_run_pc), not from their argument
};

constexpr int BOUNDS_BLOCKS = 592;   // 148 SMs x 4
constexpr int BOUNDS_STRIDE = 16;    // doubles per block in boundsPartial: lo[3], hi[3], hmax, max ratio, max growth, sum h,
                                     // [10] largest small h, [11] number of large particles (k_hmax_small)
constexpr double LARGE_FACTOR = 2.;  // particles with h > LARGE_FACTOR * mean h are "large" ...
constexpr uint32_t LARGE_MAX = 1024; // ... if there are at most this many of them
constexpr int SCAN_ITEMS = 4096;     // items per scan block

} // namespace sph

struct sphgpu_ctx {
    int device = 0;
    uint32_t n = 0, capacity = 0, nActive = 0, maxCells = 0, scanBlocks = 0, maxSegs = 0, poolRows = 0;
    sph::ParamsDev prm{};
    sph::MaterialDev matsHost[sph::MAX_MATERIALS];
    sphgpu_material matsApi[sph::MAX_MATERIALS];
    uint32_t nMaterials = 0;
    bool solid = false, corrected = false, filter = false, hasReduce = false, hasDamage = false, balsara = false, xsph = false, deltasph = false, stressAv = false;
    int recDoubles = sph::REC_FLUID; // doubles per sorted neighbour record of this context
    sph::DevicePointers d{};
    void* staging = nullptr;   // device staging for AoS <-> SoA repack (capacity * 64 B)
    // asynchronous downloads (sphgpu_download_async): results are packed into stagingDown on `stream` and leave on
    // copyStream, so the device -> host DMA of one step overlaps with the host -> device DMA of the next one
    void* stagingDown = nullptr;
    size_t stagingDownBytes = 0, downOffset = 0;
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evPacked = nullptr, evCopied = nullptr;
    bool copiesPending = false;
    cudaStream_t stream = nullptr;        // stream all work is queued on (private or caller-provided)
    cudaStream_t privateStream = nullptr;
    double maxChange = 1.e308;            // TIMESTEPPING_MAX_INCREASE
    cudaEvent_t ev[8] = {};
    cudaEvent_t evPair[4] = {}; // unit preparation | k_pair_lists | k_pair_sum (+ fallback) boundaries
    double lastPairMs[3] = { 0, 0, 0 };
    bool pairTimed = false;     // the last pair stage recorded evPair (variant 0)
    double lastMs[4] = { 0, 0, 0, 0 };
    double lastHaloMs = 0.;    // device time of the last halo exchange (pack + NCCL + unpack, includes waiting for peers)
    double lastDt = 0.;        // MultiCriterion::lastStep
    bool lastDtInit = false;
    int variant = 0;
    double listSkin = 0.03;    // relative enlargement of the search radius of the candidate lists (0: rebuild every step)
    bool listsDirty = true;    // the next integrate() must rebuild (first call, particle counts changed, pool resized)
    uint32_t listRebuilds = 0, listAge = 0; // ListCtlDev as of the last call that synchronised (collectStats)
    double listMetric = 0.;
    double haloMargin = 0.;    // head-room seen by the halo guard at the last exchange (collectStats)
    uint32_t launches = 0;
    bool stateUploaded = false;
    void* halo = nullptr;      // sph::HaloState (halo.cu): NCCL communicator + exchange buffers
    void* symmetric = nullptr; // sph::SymState (pair_symmetric.cu): buffers of the symmetric formulation (variant 4)
    bool hasFrozen = false;    // FrozenParticles boundary condition (sphgpu_set_frozen)
    sphgpu_frozen frozen{};
    void* gravity = nullptr;   // sph::GravState (gravity.cu): self-gravity, when configured
    double gravityConstant = 0.;
};

namespace sph {

void setError(const std::string& msg);

#define SPH_CUDA_CHECK(expr)                                                                                          \
    do {                                                                                                              \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess) {                                                                                      \
            sph::setError(std::string(#expr) + ": " + cudaGetErrorString(_e));                                       \
            return _e == cudaErrorMemoryAllocation ? SPHGPU_E_OOM : SPHGPU_E_CUDA;                                    \
        }                                                                                                             \
    } while (0)

// grid.cu
int launchGridBuild(sphgpu_ctx* ctx);
int launchSegments(sphgpu_ctx* ctx);
// pair.cu
int ensureConstants(const sphgpu_ctx* ctx);  // this context's ParamsDev / MaterialDev are the ones in constant memory
void forgetConstants(const sphgpu_ctx* ctx);
int launchProloguePack(sphgpu_ctx* ctx);
int launchProloguePackPositionsOnly(sphgpu_ctx* ctx);
int launchPair(sphgpu_ctx* ctx);
int launchPairSymmetric(sphgpu_ctx* ctx); // pair_symmetric.cu
void destroySymmetric(sphgpu_ctx* ctx);
int launchNeighbourCount(sphgpu_ctx* ctx, uint32_t* countsDev);
int launchNeighbourFill(sphgpu_ctx* ctx, const unsigned long long* offsetsDev, uint32_t* idxDev);
// stepping.cu
int launchPredict(sphgpu_ctx* ctx, double dt);
int launchCorrect(sphgpu_ctx* ctx, double dt);
int launchCorrectPredict(sphgpu_ctx* ctx); // corrector of the step that ends + predictor of the next one (sphgpu_run_pc)
int launchCorrectPredictRange(sphgpu_ctx* ctx, uint32_t first, uint32_t end); // the same for the slots [first, end)
int launchEuler(sphgpu_ctx* ctx, double dt);
int launchCriteria(sphgpu_ctx* ctx);
int launchFrozen(sphgpu_ctx* ctx); // FrozenParticles::finalize, when configured
int measureFp64Peak(sphgpu_ctx* ctx, double* fmaPerSecond);
int launchFinishTimestep(sphgpu_ctx* ctx, double maxDt, StepRecordDev* history, uint32_t index);
// transfer.cu
int launchUnpack(sphgpu_ctx* ctx, int q, int order, int layout, const void* stagingDev, uint32_t first, uint32_t count);
int launchPack(sphgpu_ctx* ctx, int q, int order, int layout, void* stagingDev, uint32_t first, uint32_t count);
size_t elementBytes(int q, int layout);
int launchHalo(sphgpu_ctx* ctx, bool pack, uint32_t first, uint32_t count, void* buf);
// api.cu / halo.cu
int enqueueIntegrate(sphgpu_ctx* ctx);
int collectStats(sphgpu_ctx* ctx, sphgpu_stats* stats, cudaEvent_t begin, cudaEvent_t end);
int finishTimestep(sphgpu_ctx* ctx, double max_dt, sphgpu_timestep* out);
void destroyHalo(sphgpu_ctx* ctx);
void invalidateHalo(sphgpu_ctx* ctx); // the exchange refuses to run until sphgpu_halo_configure is called again
// gravity.cu
int launchGravity(sphgpu_ctx* ctx, int accumulate); // no-op unless self-gravity is configured
void destroyGravity(sphgpu_ctx* ctx);

} // namespace sph
