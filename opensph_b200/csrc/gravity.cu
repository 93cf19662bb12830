// Self-gravity on the device: Barnes-Hut with multipoles up to the octupole, replacing the reference's host-side
// BarnesHut (core/gravity/BarnesHut.cpp:50-501: build, buildLeaf / buildInner, evalNode, evalParticleList,
// evalNodeList) and BruteForceGravity (core/gravity/BruteForceGravity.h:38-45,103-121) as GravitySolver<TSphSolver>::loop
// composes them with the SPH part (core/sph/solvers/GravitySolver.cpp:64-99). SURVEY section 8(f) #1.
//
// The reference walks a k-d tree with one task per node and per-node interaction lists. Here:
//   1. k_grav_bbox_*   bounding cube of the owned particles
//   2. k_grav_keys     63-bit Morton keys (21 bits per axis), sorted with their particle slots by cub::DeviceRadixSort
//                      (the one library call of the library: a support step, not the hot loop)
//   3. k_grav_gather   sorted copies {x, y, z, h}, G m
//   4. k_grav_tree     binary radix tree over the sorted keys (Karras 2012): every internal node finds its key range and
//                      its split independently; a node covers a contiguous range of sorted particles
//   5. k_grav_moments  bottom-up (second arrival computes the parent): nodes of <= leafSize particles (the analogue of the
//                      k-d tree's leaves) sum their particles directly like buildLeaf, larger nodes combine their two
//                      children like buildInner; bounding boxes, centres of mass, opening radii, traceless moments
//   6. target groups   the maximal nodes of <= 32 particles (marked by k_grav_moments, compacted in the order of the sorted
//                      particles by cub::DeviceSelect)
//   7. k_grav_walk     one warp per group, one target per lane: depth-first walk of the tree with a warp-wide stack in
//                      shared memory, 32 nodes classified per trip with the reference's criterion for an evaluated leaf
//                      (BarnesHut::evalNode): the opening ball of the node does not reach the group's box -> multipole
//                      approximation, otherwise open it (or, for a leaf, sum its particles exactly). Accepted nodes and
//                      particle ranges are staged in shared memory 32 at a time and applied to all lanes. The exact
//                      particle pairs are FP64 throughout; the multipole expansions of the accepted (distant) nodes are
//                      evaluated in single precision relative to the group's centre and added to FP64 sums -- their
//                      rounding (6e-8) is three orders below the error of the approximation itself, and the FP64 pipe,
//                      which bounds this kernel, is left to the pairs.
// Summation order is fixed by the tree, so results are reproducible from run to run.
#include "grav_math.cuh"
#include "sphgpu_internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

namespace sph {

constexpr int GRAV_GROUP = 32;         // targets per group (one per lane)
constexpr int GRAV_WARPS = 4;          // warps (groups in flight) per CTA of the walk
constexpr int GRAV_STACK = 1024;       // stack entries per warp (a nearly full stack is drained one node at a time, see the walk)
constexpr uint32_t GRAV_NONE = 0xffffffffu;

struct GravDev {
    uint32_t n;
    unsigned long long *keys, *keysSorted;
    uint32_t *slot, *slotSorted; // sorted position -> particle slot
    double4* spos;               // sorted {x, y, z, h}
    double* smass;               // sorted G m
    int4* meta;                  // internal nodes [n - 1]: {left, right, first, last}; child >= 0 internal, < 0 particle ~child
    int* parent;                 // [n - 1]
    int* parentLeaf;             // [n]
    uint32_t* arrived;           // [n - 1]
    double4* sphere;             // {centre of mass, opening radius}
    double* box;                 // [6 (n - 1)] lower, upper
    GravNode* node;              // source record
    GravRaw* raw;                // raw moments about the centre of mass
    uint32_t* groupAt;           // [n] node of the group that starts at this sorted particle (GRAV_NONE: none)
    uint32_t* groups;            // compacted, in the order of the sorted particles
    uint32_t* groupCount;
    double* bounds;              // [8] lo[3], scale (key cells per unit length), then scratch
    double* boundsPartial;       // [blocks * 6]
    unsigned long long* counters; // [0] approximated node-target-group interactions, [1] exact particle ranges, [2] stack overflow
    const LutPair* lut;
};

struct GravState {
    GravDev d{};
    GravParams prm{};
    uint32_t capacity = 0;
    void* cubTemp = nullptr;
    size_t cubTempBytes = 0;
    cudaEvent_t ev[2] = {};
    double lastMs = 0.;
    unsigned long long lastCounters[3] = { 0, 0, 0 };
    uint32_t lastGroups = 0;
};

// ---- 1. bounds ----------------------------------------------------------------------------------------------------
constexpr int GRAV_BBOX_BLOCKS = 296;

__global__ void __launch_bounds__(256) k_grav_bbox_partial(DevicePointers p, GravDev g) {
    double lo[3] = { INFTY_REF, INFTY_REF, INFTY_REF }, hi[3] = { -INFTY_REF, -INFTY_REF, -INFTY_REF };
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += gridDim.x * blockDim.x) {
        const double x = p.f[F_X][i], y = p.f[F_Y][i], z = p.f[F_Z][i];
        lo[0] = fmin(lo[0], x); lo[1] = fmin(lo[1], y); lo[2] = fmin(lo[2], z);
        hi[0] = fmax(hi[0], x); hi[1] = fmax(hi[1], y); hi[2] = fmax(hi[2], z);
    }
    __shared__ double s[8][6];
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; ++k) {
            s[threadIdx.x >> 5][k] = lo[k];
            s[threadIdx.x >> 5][3 + k] = hi[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = s[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) {
            v = threadIdx.x < 3 ? fmin(v, s[w][threadIdx.x]) : fmax(v, s[w][threadIdx.x]);
        }
        g.boundsPartial[blockIdx.x * 6 + threadIdx.x] = v;
    }
}

__global__ void k_grav_bbox_final(GravDev g, int blocks) {
    if (threadIdx.x < 6) {
        double v = g.boundsPartial[threadIdx.x];
        for (int b = 1; b < blocks; ++b) {
            const double w = g.boundsPartial[b * 6 + threadIdx.x];
            v = threadIdx.x < 3 ? fmin(v, w) : fmax(v, w);
        }
        g.bounds[4 + threadIdx.x] = v;
    }
    __syncwarp();
    if (threadIdx.x == 0) {
        const double ext = fmax(fmax(g.bounds[7] - g.bounds[4], g.bounds[8] - g.bounds[5]), g.bounds[9] - g.bounds[6]);
        g.bounds[0] = g.bounds[4];
        g.bounds[1] = g.bounds[5];
        g.bounds[2] = g.bounds[6];
        g.bounds[3] = ext > 0. ? 2097151. / ext : 0.; // 2^21 - 1 key cells along the longest edge
        g.counters[0] = g.counters[1] = g.counters[2] = 0ull;
        *g.groupCount = 0u;
    }
}

// ---- 2. Morton keys -----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long spreadBits21(unsigned long long v) { // bit k -> bit 3k
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void __launch_bounds__(256) k_grav_keys(DevicePointers p, GravDev g) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n) {
        return;
    }
    const double s = g.bounds[3];
    const double qx = fmin(fmax((p.f[F_X][i] - g.bounds[0]) * s, 0.), 2097151.);
    const double qy = fmin(fmax((p.f[F_Y][i] - g.bounds[1]) * s, 0.), 2097151.);
    const double qz = fmin(fmax((p.f[F_Z][i] - g.bounds[2]) * s, 0.), 2097151.);
    g.keys[i] = spreadBits21((unsigned long long)qx) | (spreadBits21((unsigned long long)qy) << 1) | (spreadBits21((unsigned long long)qz) << 2);
    g.slot[i] = i;
}

// ---- 3. sorted copies -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_grav_gather(DevicePointers p, GravDev g, double G) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.n) {
        return;
    }
    const uint32_t i = g.slotSorted[t];
    g.spos[t] = make_double4(p.f[F_X][i], p.f[F_Y][i], p.f[F_Z][i], p.f[F_H][i]);
    g.smass[t] = G * p.f[F_M][i];
    g.groupAt[t] = GRAV_NONE;
    if (t + 1 < g.n) {
        g.arrived[t] = 0u;
    }
}

// ---- 4. binary radix tree (Karras 2012) -----------------------------------------------------------------------------
/// Length of the common prefix of the keys at sorted positions i and j (equal keys are told apart by their positions);
/// -1 outside the array.
__device__ __forceinline__ int gravDelta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) {
        return -1;
    }
    const unsigned long long a = keys[i], b = keys[j];
    return a != b ? __clzll((long long)(a ^ b)) : 64 + __clz(i ^ j);
}

__global__ void __launch_bounds__(256) k_grav_tree(GravDev g) {
    const int n = (int)g.n;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) {
        return;
    }
    const unsigned long long* keys = g.keysSorted;
    const int d = gravDelta(keys, n, i, i + 1) - gravDelta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dMin = gravDelta(keys, n, i, i - d);
    int lMax = 2;
    while (gravDelta(keys, n, i, i + lMax * d) > dMin) {
        lMax *= 2;
    }
    int l = 0;
    for (int t = lMax / 2; t >= 1; t /= 2) {
        if (gravDelta(keys, n, i, i + (l + t) * d) > dMin) {
            l += t;
        }
    }
    const int j = i + l * d;
    const int dNode = gravDelta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (gravDelta(keys, n, i, i + (s + t) * d) > dNode) {
            s += t;
        }
        if (t == 1) {
            break;
        }
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = first == gamma ? ~gamma : gamma;
    const int right = last == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    g.meta[i] = make_int4(left, right, first, last);
    if (left >= 0) {
        g.parent[left] = i;
    } else {
        g.parentLeaf[gamma] = i;
    }
    if (right >= 0) {
        g.parent[right] = i;
    } else {
        g.parentLeaf[gamma + 1] = i;
    }
    if (i == 0) {
        g.parent[0] = -1;
    }
}

// ---- 5. moments, bottom-up ------------------------------------------------------------------------------------------
struct GravSummary { // what a parent needs from a child
    double com[3], m, lo[3], hi[3];
    GravRaw raw;
};

__device__ __forceinline__ void gravStoreNode(const GravDev& g, int node, const GravSummary& s, const GravParams& prm, uint32_t count) {
    GravNode out;
    out.cx = s.com[0];
    out.cy = s.com[1];
    out.cz = s.com[2];
    out.m = s.m;
    gravReduce(s.raw, out);
    g.node[node] = out;
    g.raw[node] = s.raw;
    for (int k = 0; k < 3; ++k) {
        g.box[6 * (size_t)node + k] = s.lo[k];
        g.box[6 * (size_t)node + 3 + k] = s.hi[k];
    }
    // a node whose particles coincide has r_open = 0 like the reference's single-particle leaves: always summed exactly
    const double rOpen = prm.exact ? INFTY_REF : gravOpeningRadius(s.com, s.lo, s.hi, prm.thetaInv);
    g.sphere[node] = make_double4(s.com[0], s.com[1], s.com[2], count > 1 ? rOpen : 0.);
}

/// buildLeaf (BarnesHut.cpp:371-432): centre of mass, box and moments of the sorted particles [first, last].
__device__ void gravLeafSummary(const GravDev& g, int first, int last, GravSummary& s) {
    s.m = 0.;
    for (int k = 0; k < 3; ++k) {
        s.com[k] = 0.;
        s.lo[k] = INFTY_REF;
        s.hi[k] = -INFTY_REF;
    }
    for (int t = first; t <= last; ++t) {
        const double4 r = g.spos[t];
        const double m = g.smass[t];
        s.com[0] += m * r.x;
        s.com[1] += m * r.y;
        s.com[2] += m * r.z;
        s.m += m;
        s.lo[0] = fmin(s.lo[0], r.x); s.lo[1] = fmin(s.lo[1], r.y); s.lo[2] = fmin(s.lo[2], r.z);
        s.hi[0] = fmax(s.hi[0], r.x); s.hi[1] = fmax(s.hi[1], r.y); s.hi[2] = fmax(s.hi[2], r.z);
    }
    if (s.m > 0.) {
        const double inv = 1. / s.m;
        s.com[0] *= inv;
        s.com[1] *= inv;
        s.com[2] *= inv;
    } else { // massless particles only: any centre serves, the moments are zero
        for (int k = 0; k < 3; ++k) {
            s.com[k] = 0.5 * (s.lo[k] + s.hi[k]);
        }
    }
    gravRawZero(s.raw);
    for (int t = first; t <= last; ++t) {
        const double4 r = g.spos[t];
        gravRawAddPoint(s.raw, g.smass[t], r.x - s.com[0], r.y - s.com[1], r.z - s.com[2]);
    }
}

__device__ __forceinline__ void gravChildSummary(const GravDev& g, int child, GravSummary& s) {
    if (child < 0) { // a single particle
        const double4 r = g.spos[~child];
        s.com[0] = s.lo[0] = s.hi[0] = r.x;
        s.com[1] = s.lo[1] = s.hi[1] = r.y;
        s.com[2] = s.lo[2] = s.hi[2] = r.z;
        s.m = g.smass[~child];
        gravRawZero(s.raw);
        return;
    }
    // written by another thread of this kernel: read through L2
    const double* nd = reinterpret_cast<const double*>(g.node + child);
    s.com[0] = __ldcg(nd + 0);
    s.com[1] = __ldcg(nd + 1);
    s.com[2] = __ldcg(nd + 2);
    s.m = __ldcg(nd + 3);
    for (int k = 0; k < 3; ++k) {
        s.lo[k] = __ldcg(g.box + 6 * (size_t)child + k);
        s.hi[k] = __ldcg(g.box + 6 * (size_t)child + 3 + k);
    }
    const double* rw = reinterpret_cast<const double*>(g.raw + child);
    for (int k = 0; k < 6; ++k) {
        s.raw.m2[k] = __ldcg(rw + k);
    }
    for (int k = 0; k < 10; ++k) {
        s.raw.m3[k] = __ldcg(rw + 6 + k);
    }
}

__global__ void __launch_bounds__(128) k_grav_moments(GravDev g, GravParams prm) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.n) {
        return;
    }
    const int leaf = (int)prm.leafSize;
    int cur = g.parentLeaf[t];
    while (cur >= 0) {
        if (atomicAdd(&g.arrived[cur], 1u) == 0u) {
            return; // the sibling subtree is not finished yet: its last thread will come back here
        }
        __threadfence();
        const int4 meta = g.meta[cur];
        const int count = meta.w - meta.z + 1;
        const int par = g.parent[cur];
        int parCount = 0x7fffffff;
        if (par >= 0) {
            const int4 pm = g.meta[par];
            parCount = pm.w - pm.z + 1;
        }
        // target groups: the maximal nodes of at most GRAV_GROUP particles (one per lane of the walk)
        const bool isGroup = count <= GRAV_GROUP && parCount > GRAV_GROUP;
        if (isGroup) {
            g.groupAt[meta.z] = (uint32_t)cur;
        }
        if (count <= leaf) {
            if (parCount > leaf || isGroup) { // a leaf of the gravity tree (or a group inside one: it needs its box)
                GravSummary s;
                gravLeafSummary(g, meta.z, meta.w, s);
                gravStoreNode(g, cur, s, prm, (uint32_t)count);
            }
        } else { // buildInner (BarnesHut.cpp:434-489)
            GravSummary a, b, s;
            gravChildSummary(g, meta.x, a);
            gravChildSummary(g, meta.y, b);
            s.m = a.m + b.m;
            for (int k = 0; k < 3; ++k) {
                s.lo[k] = fmin(a.lo[k], b.lo[k]);
                s.hi[k] = fmax(a.hi[k], b.hi[k]);
                s.com[k] = s.m > 0. ? (a.m * a.com[k] + b.m * b.com[k]) / s.m : 0.5 * (s.lo[k] + s.hi[k]);
            }
            gravRawZero(s.raw);
            gravRawAddShifted(s.raw, a.raw, a.m, a.com[0] - s.com[0], a.com[1] - s.com[1], a.com[2] - s.com[2]);
            gravRawAddShifted(s.raw, b.raw, b.m, b.com[0] - s.com[0], b.com[1] - s.com[1], b.com[2] - s.com[2]);
            gravStoreNode(g, cur, s, prm, (uint32_t)count);
        }
        if (count > GRAV_GROUP) { // single particles hanging off a larger node are groups of their own
            if (meta.x < 0) {
                g.groupAt[~meta.x] = 0x80000000u | (uint32_t)(~meta.x);
            }
            if (meta.y < 0) {
                g.groupAt[~meta.y] = 0x80000000u | (uint32_t)(~meta.y);
            }
        }
        __threadfence();
        cur = par;
    }
}

struct GravIsGroup {
    __device__ __forceinline__ bool operator()(const uint32_t& v) const {
        return v != GRAV_NONE;
    }
};

// ---- 7. the walk ----------------------------------------------------------------------------------------------------
struct GravWarpShared {
    uint32_t stack[GRAV_STACK];
    float nodeStage[32][16];      // 32 accepted nodes (GravNodeF: relative to the group's centre, single precision)
    double partStage[5][32];      // x, y, z, h, m of a particle range
    uint32_t approx[64];          // accepted nodes waiting for a full batch
    uint2 exact[96];              // particle ranges {first, count} waiting
};

/// Applies `count` (<= 32) accepted nodes to all lanes. Lane l stages node l: centre of mass relative to the group's centre
/// (FP64 subtraction, then single precision) and the moments; every lane then evaluates the expansion in single precision
/// at its own offset from the group's centre and adds the batch to its FP64 sums.
template <int ORDER>
__device__ __forceinline__ void gravApplyNodes(const GravDev& g, GravWarpShared& w, int count, int lane, bool live, double gcx, double gcy, double gcz,
    double invL, double invM, double accUnit, float ox, float oy, float oz, double& ax, double& ay, double& az) {
    if (lane < count) {
        const double* nd = reinterpret_cast<const double*>(g.node + w.approx[lane]); // GravNode: com, m, q2[5], q3[7]
        float* dst = &w.nodeStage[lane][0];
        dst[0] = (float)((__ldg(nd + 0) - gcx) * invL);
        dst[1] = (float)((__ldg(nd + 1) - gcy) * invL);
        dst[2] = (float)((__ldg(nd + 2) - gcz) * invL);
        dst[3] = (float)(__ldg(nd + 3) * invM);
        if (ORDER >= 2) {
            const double s2 = invM * invL * invL, s3 = s2 * invL;
#pragma unroll
            for (int q = 4; q < 9; ++q) {
                dst[q] = (float)(__ldg(nd + q) * s2);
            }
            if (ORDER >= 3) {
#pragma unroll
                for (int q = 9; q < 16; ++q) {
                    dst[q] = (float)(__ldg(nd + q) * s3);
                }
            }
        }
    }
    __syncwarp();
    if (live) {
        float fx = 0.f, fy = 0.f, fz = 0.f;
        for (int k = 0; k < count; ++k) {
            const GravNodeF& nd = *reinterpret_cast<const GravNodeF*>(&w.nodeStage[k][0]);
            gravNodeAccelF<ORDER>(nd, ox, oy, oz, fx, fy, fz);
        }
        ax += (double)fx * accUnit;
        ay += (double)fy * accUnit;
        az += (double)fz * accUnit;
    }
    __syncwarp();
}

__device__ __forceinline__ void gravApplyRange(const GravDev& g, const GravParams& prm, GravWarpShared& w, uint2 range, int lane, bool live,
    uint32_t self, double x, double y, double z, double h, double& ax, double& ay, double& az) {
    if ((uint32_t)lane < range.y) {
        const double4 r = g.spos[range.x + lane];
        w.partStage[0][lane] = r.x;
        w.partStage[1][lane] = r.y;
        w.partStage[2][lane] = r.z;
        w.partStage[3][lane] = r.w;
        w.partStage[4][lane] = g.smass[range.x + lane];
    }
    __syncwarp();
    if (live) {
        for (uint32_t k = 0; k < range.y; ++k) {
            if (range.x + k != self) {
                gravPairAccel(prm, g.lut, x, y, z, h, w.partStage[0][k], w.partStage[1][k], w.partStage[2][k], w.partStage[3][k], w.partStage[4][k],
                    ax, ay, az);
            }
        }
    }
    __syncwarp();
}

template <int ORDER>
__global__ void __launch_bounds__(GRAV_WARPS * 32) k_grav_walk(DevicePointers p, GravDev g, GravParams prm, int accumulate) {
    extern __shared__ __align__(16) unsigned char gravSmem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    GravWarpShared& w = reinterpret_cast<GravWarpShared*>(gravSmem)[warp];
    const uint32_t nGroups = *g.groupCount;
    const uint32_t ltMask = (1u << lane) - 1u;
    // units of the single-precision far field: the bounding cube and the total mass (GravNodeF)
    const double invL = g.bounds[3] * (1. / 2097151.);
    const double rootM = g.node[0].m;
    const double invM = rootM > 0. ? 1. / rootM : 0.;
    const double accUnit = rootM * invL * invL;
    unsigned long long nApprox = 0, nExact = 0;
    for (uint32_t gi = blockIdx.x * GRAV_WARPS + warp; gi < nGroups; gi += gridDim.x * GRAV_WARPS) {
        const uint32_t code = g.groups[gi];
        uint32_t first, count;
        double lo[3], hi[3];
        if (code & 0x80000000u) { // a single particle
            first = code & 0x7fffffffu;
            count = 1u;
            const double4 r = g.spos[first];
            lo[0] = hi[0] = r.x;
            lo[1] = hi[1] = r.y;
            lo[2] = hi[2] = r.z;
        } else {
            const int4 meta = g.meta[code];
            first = (uint32_t)meta.z;
            count = (uint32_t)(meta.w - meta.z + 1);
            for (int k = 0; k < 3; ++k) {
                lo[k] = g.box[6 * (size_t)code + k];
                hi[k] = g.box[6 * (size_t)code + 3 + k];
            }
        }
        double ax = 0., ay = 0., az = 0.;
        const uint32_t nPasses = (count + GRAV_GROUP - 1) / GRAV_GROUP; // 1 unless leafSize > 32
        for (uint32_t pass = 0; pass < nPasses; ++pass) {
            const uint32_t self = first + pass * GRAV_GROUP + (uint32_t)lane;
            const bool live = self < first + count;
            double4 ri = make_double4(0., 0., 0., 1.);
            if (live) {
                ri = g.spos[self];
            }
            const double gcx = 0.5 * (lo[0] + hi[0]), gcy = 0.5 * (lo[1] + hi[1]), gcz = 0.5 * (lo[2] + hi[2]);
            const float ox = (float)((ri.x - gcx) * invL), oy = (float)((ri.y - gcy) * invL), oz = (float)((ri.z - gcz) * invL);
            ax = ay = az = 0.;
            int sp = 0, nA = 0, nE = 0;
            if (lane == 0) {
                w.stack[0] = 0u; // the root
            }
            sp = 1;
            __syncwarp();
            while (sp > 0) {
                // 32 nodes per trip push at most 64 children; when that might not fit, one node per trip (at most two
                // children for one popped node): the depth-first walk then needs no more entries than the tree is deep
                const int take = sp + 64 > GRAV_STACK ? 1 : min(sp, 32);
                sp -= take;
                int4 meta = make_int4(0, 0, 0, -1);
                uint32_t nodeId = 0;
                bool isApprox = false, isRange = false, pushL = false, pushR = false, oneL = false, oneR = false;
                if (lane < take) {
                    nodeId = w.stack[sp + lane];
                    meta = g.meta[nodeId];
                    const double4 ball = g.sphere[nodeId];
                    const int cnt = meta.w - meta.z + 1;
                    const bool leafLike = cnt <= (int)prm.leafSize;
                    const bool open = ball.w == 0. || gravBallOverlapsBox(ball.x, ball.y, ball.z, ball.w, lo, hi);
                    if (!open) {
                        isApprox = true;
                    } else if (leafLike) {
                        isRange = true;
                    } else {
                        pushL = meta.x >= 0;
                        pushR = meta.y >= 0;
                        oneL = meta.x < 0;
                        oneR = meta.y < 0;
                    }
                }
                __syncwarp();
                // children of the opened nodes -> stack
                const uint32_t bL = __ballot_sync(0xffffffffu, pushL), bR = __ballot_sync(0xffffffffu, pushR);
                const int nPush = __popc(bL) + __popc(bR);
                if (sp + nPush > GRAV_STACK) {
                    if (lane == 0) {
                        atomicAdd(&g.counters[2], 1ull);
                    }
                    sp = 0;
                    break;
                }
                int off = sp + __popc(bL & ltMask) + __popc(bR & ltMask);
                if (pushL) {
                    w.stack[off++] = (uint32_t)meta.x;
                }
                if (pushR) {
                    w.stack[off] = (uint32_t)meta.y;
                }
                sp += nPush;
                // accepted nodes
                const uint32_t bA = __ballot_sync(0xffffffffu, isApprox);
                if (isApprox) {
                    w.approx[nA + __popc(bA & ltMask)] = nodeId;
                }
                nA += __popc(bA);
                // particle ranges: leaves summed exactly, single particles below opened nodes
                const uint32_t bE = __ballot_sync(0xffffffffu, isRange), b1 = __ballot_sync(0xffffffffu, oneL), b2 = __ballot_sync(0xffffffffu, oneR);
                int eoff = nE + __popc(bE & ltMask) + __popc(b1 & ltMask) + __popc(b2 & ltMask);
                if (isRange) {
                    w.exact[eoff] = make_uint2((uint32_t)meta.z, (uint32_t)(meta.w - meta.z + 1));
                }
                if (oneL) {
                    w.exact[eoff++] = make_uint2((uint32_t)(~meta.x), 1u);
                }
                if (oneR) {
                    w.exact[eoff] = make_uint2((uint32_t)(~meta.y), 1u);
                }
                nE += __popc(bE) + __popc(b1) + __popc(b2);
                __syncwarp();
                if (nA >= 32) {
                    gravApplyNodes<ORDER>(g, w, 32, lane, live, gcx, gcy, gcz, invL, invM, accUnit, ox, oy, oz, ax, ay, az);
                    nApprox += 32;
                    const uint32_t keep = w.approx[32 + lane];
                    __syncwarp();
                    w.approx[lane] = keep;
                    nA -= 32;
                    __syncwarp();
                }
                if (nE >= 32) { // (at most 64 new ranges per trip: the buffer holds 96)
                    for (int k = 0; k < nE; ++k) {
                        uint2 range = w.exact[k];
                        // ranges longer than a warp (leafSize > 32) are taken in pieces
                        for (uint32_t o = 0; o < range.y; o += 32) {
                            gravApplyRange(g, prm, w, make_uint2(range.x + o, min(range.y - o, 32u)), lane, live, self, ri.x, ri.y, ri.z, ri.w, ax, ay,
                                az);
                        }
                    }
                    nExact += (unsigned long long)nE;
                    nE = 0;
                }
            }
            if (nA > 0) {
                gravApplyNodes<ORDER>(g, w, nA, lane, live, gcx, gcy, gcz, invL, invM, accUnit, ox, oy, oz, ax, ay, az);
                nApprox += (unsigned long long)nA;
            }
            for (int k = 0; k < nE; ++k) {
                const uint2 range = w.exact[k];
                for (uint32_t o = 0; o < range.y; o += 32) {
                    gravApplyRange(g, prm, w, make_uint2(range.x + o, min(range.y - o, 32u)), lane, live, self, ri.x, ri.y, ri.z, ri.w, ax, ay, az);
                }
            }
            nExact += (unsigned long long)nE;
            if (live) {
                const uint32_t slot = g.slotSorted[self];
                if (accumulate) {
                    p.f[F_AX][slot] += ax;
                    p.f[F_AY][slot] += ay;
                    p.f[F_AZ][slot] += az;
                } else {
                    p.f[F_AX][slot] = ax;
                    p.f[F_AY][slot] = ay;
                    p.f[F_AZ][slot] = az;
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0 && (nApprox | nExact) != 0ull) {
        atomicAdd(&g.counters[0], nApprox);
        atomicAdd(&g.counters[1], nExact);
    }
}

__global__ void k_grav_single(DevicePointers p, int accumulate) { // one particle: no gravity
    if (!accumulate) {
        p.f[F_AX][0] = p.f[F_AY][0] = p.f[F_AZ][0] = 0.;
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
static GravState* gravState(sphgpu_ctx* ctx) {
    return static_cast<GravState*>(ctx->gravity);
}

void destroyGravity(sphgpu_ctx* ctx) {
    GravState* s = gravState(ctx);
    if (!s) {
        return;
    }
    GravDev& d = s->d;
    cudaFree(d.keys); cudaFree(d.keysSorted); cudaFree(d.slot); cudaFree(d.slotSorted); cudaFree(d.spos); cudaFree(d.smass);
    cudaFree(d.meta); cudaFree(d.parent); cudaFree(d.parentLeaf); cudaFree(d.arrived); cudaFree(d.sphere); cudaFree(d.box);
    cudaFree(d.node); cudaFree(d.raw); cudaFree(d.groupAt); cudaFree(d.groups); cudaFree(d.groupCount); cudaFree(d.bounds);
    cudaFree(d.boundsPartial); cudaFree(d.counters); cudaFree((void*)d.lut); cudaFree(s->cubTemp);
    for (int k = 0; k < 2; ++k) {
        if (s->ev[k]) cudaEventDestroy(s->ev[k]);
    }
    cudaGetLastError();
    delete s;
    ctx->gravity = nullptr;
}

template <typename T>
static int gravAlloc(T** p, size_t count) {
    SPH_CUDA_CHECK(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    return SPHGPU_OK;
}

int configureGravity(sphgpu_ctx* ctx, const sphgpu_gravity* cfg) {
    destroyGravity(ctx);
    if (!cfg) {
        return SPHGPU_OK;
    }
    GravState* s = new (std::nothrow) GravState();
    if (!s) {
        setError("host allocation failed");
        return SPHGPU_E_OOM;
    }
    ctx->gravity = s;
    s->capacity = ctx->capacity;
    const size_t cap = ctx->capacity, inner = cap > 0 ? cap : 1;
    GravDev& d = s->d;
    int rc = SPHGPU_OK;
#define GRAV_TRY(expr)                                                                                                \
    if (rc == SPHGPU_OK) {                                                                                            \
        rc = (expr);                                                                                                  \
    }
    GRAV_TRY(gravAlloc(&d.keys, cap));
    GRAV_TRY(gravAlloc(&d.keysSorted, cap));
    GRAV_TRY(gravAlloc(&d.slot, cap));
    GRAV_TRY(gravAlloc(&d.slotSorted, cap));
    GRAV_TRY(gravAlloc(&d.spos, cap));
    GRAV_TRY(gravAlloc(&d.smass, cap));
    GRAV_TRY(gravAlloc(&d.meta, inner));
    GRAV_TRY(gravAlloc(&d.parent, inner));
    GRAV_TRY(gravAlloc(&d.parentLeaf, cap));
    GRAV_TRY(gravAlloc(&d.arrived, inner));
    GRAV_TRY(gravAlloc(&d.sphere, inner));
    GRAV_TRY(gravAlloc(&d.box, 6 * inner));
    GRAV_TRY(gravAlloc(&d.node, inner));
    GRAV_TRY(gravAlloc(&d.raw, inner));
    GRAV_TRY(gravAlloc(&d.groupAt, cap));
    GRAV_TRY(gravAlloc(&d.groups, cap));
    GRAV_TRY(gravAlloc(&d.groupCount, 1));
    GRAV_TRY(gravAlloc(&d.bounds, 16));
    GRAV_TRY(gravAlloc(&d.boundsPartial, (size_t)GRAV_BBOX_BLOCKS * 6));
    GRAV_TRY(gravAlloc(&d.counters, 4));
    if (rc == SPHGPU_OK) {
        size_t a = 0, b = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, a, d.keys, d.keysSorted, d.slot, d.slotSorted, (int)cap, 0, 63, ctx->stream);
        cub::DeviceSelect::If(nullptr, b, d.groupAt, d.groups, d.groupCount, (int)cap, GravIsGroup(), ctx->stream);
        s->cubTempBytes = std::max(a, b) + 256;
        cudaError_t e = cudaMalloc(&s->cubTemp, s->cubTempBytes);
        if (e != cudaSuccess) {
            setError(std::string("cudaMalloc (sort workspace): ") + cudaGetErrorString(e));
            rc = SPHGPU_E_OOM;
        }
    }
    // the softening kernel's gradient table as {G[k], G[k + 1] - G[k]} pairs; the last node G[entries] is the Newtonian
    // value at the kernel's edge (LutKernel stores NEntries + 1 values, Kernel.h:92-99)
    GravParams& prm = s->prm;
    prm.exact = cfg->opening_angle > 0. ? 0 : 1;
    prm.thetaInv = cfg->opening_angle > 0. ? 1. / cfg->opening_angle : 0.;
    prm.order = cfg->multipole_order;
    prm.leafSize = cfg->leaf_size > 0 ? cfg->leaf_size : 25u; // FINDER_LEAF_SIZE default
    prm.radiusSqr = 0.;
    prm.qSqrToIdx = 0.;
    prm.lutEntries = 0;
    if (rc == SPHGPU_OK && cfg->kernel_radius > 0.) {
        std::vector<LutPair> pairs((size_t)cfg->lut_entries + 1);
        for (uint32_t k = 0; k <= cfg->lut_entries; ++k) {
            const double next = k < cfg->lut_entries ? cfg->lut_grad[k + 1] : cfg->lut_grad[k];
            pairs[k].g = cfg->lut_grad[k];
            pairs[k].dg = next - cfg->lut_grad[k];
        }
        LutPair* dev = nullptr;
        rc = gravAlloc(&dev, pairs.size());
        if (rc == SPHGPU_OK) {
            d.lut = dev;
            cudaError_t e = cudaMemcpy(dev, pairs.data(), pairs.size() * sizeof(LutPair), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) {
                setError(std::string("cudaMemcpy (gravity table): ") + cudaGetErrorString(e));
                rc = SPHGPU_E_CUDA;
            }
        }
        prm.radiusSqr = cfg->kernel_radius * cfg->kernel_radius;
        prm.qSqrToIdx = (double)cfg->lut_entries / prm.radiusSqr;
        prm.lutEntries = cfg->lut_entries;
    }
    for (int k = 0; k < 2 && rc == SPHGPU_OK; ++k) {
        if (cudaEventCreate(&s->ev[k]) != cudaSuccess) {
            setError("cudaEventCreate failed");
            rc = SPHGPU_E_CUDA;
        }
    }
#undef GRAV_TRY
    s->lastMs = 0.;
    ctx->gravityConstant = cfg->constant;
    if (rc != SPHGPU_OK) {
        destroyGravity(ctx);
    }
    return rc;
}

template <int ORDER>
static void launchWalk(sphgpu_ctx* ctx, GravState* s, int accumulate, int sms) {
    auto kernel = k_grav_walk<ORDER>;
    const size_t smem = sizeof(GravWarpShared) * GRAV_WARPS;
    static bool configured[64] = {};
    if (!configured[ctx->device & 63]) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured[ctx->device & 63] = true;
    }
    const uint32_t maxGroups = std::max<uint32_t>(s->d.n, 1u);
    const uint32_t blocks = (uint32_t)std::min<uint64_t>((uint64_t)sms * 16, (maxGroups + GRAV_WARPS - 1) / GRAV_WARPS);
    kernel<<<blocks, GRAV_WARPS * 32, smem, ctx->stream>>>(ctx->d, s->d, s->prm, accumulate);
}

/// Queues the whole gravity evaluation on the context's stream. accumulate: add to the accelerations (what
/// GravitySolver::loop does) or overwrite them (IGravity::evalSelfGravity on a zeroed buffer).
int launchGravity(sphgpu_ctx* ctx, int accumulate) {
    GravState* s = gravState(ctx);
    if (!s) {
        return SPHGPU_OK;
    }
    if (ctx->nActive != ctx->n || ctx->halo != nullptr) {
        setError("self-gravity is global and not available on a decomposed run (halo / ghost particles configured)");
        return SPHGPU_E_INVALID;
    }
    cudaStream_t st = ctx->stream;
    const uint32_t n = ctx->n;
    s->d.n = n;
    SPH_CUDA_CHECK(cudaEventRecord(s->ev[0], st));
    if (n == 0) {
        SPH_CUDA_CHECK(cudaEventRecord(s->ev[1], st));
        return SPHGPU_OK;
    }
    if (n == 1) {
        k_grav_single<<<1, 1, 0, st>>>(ctx->d, accumulate);
        SPH_CUDA_CHECK(cudaEventRecord(s->ev[1], st));
        ctx->launches += 1;
        return SPHGPU_OK;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const GravDev& d = s->d;
    const uint32_t blocks = (n + 255) / 256;
    const int bboxBlocks = (int)std::min<uint32_t>(GRAV_BBOX_BLOCKS, blocks);
    k_grav_bbox_partial<<<bboxBlocks, 256, 0, st>>>(ctx->d, d);
    k_grav_bbox_final<<<1, 32, 0, st>>>(d, bboxBlocks);
    k_grav_keys<<<blocks, 256, 0, st>>>(ctx->d, d);
    size_t bytes = s->cubTempBytes;
    SPH_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(s->cubTemp, bytes, d.keys, d.keysSorted, d.slot, d.slotSorted, (int)n, 0, 63, st));
    k_grav_gather<<<blocks, 256, 0, st>>>(ctx->d, d, ctx->gravityConstant);
    k_grav_tree<<<(n - 1 + 255) / 256, 256, 0, st>>>(d);
    k_grav_moments<<<(n + 127) / 128, 128, 0, st>>>(d, s->prm);
    bytes = s->cubTempBytes;
    SPH_CUDA_CHECK(cub::DeviceSelect::If(s->cubTemp, bytes, d.groupAt, d.groups, d.groupCount, (int)n, GravIsGroup(), st));
    switch (s->prm.order) {
    case 0:
        launchWalk<0>(ctx, s, accumulate, sms);
        break;
    case 2:
        launchWalk<2>(ctx, s, accumulate, sms);
        break;
    default:
        launchWalk<3>(ctx, s, accumulate, sms);
        break;
    }
    SPH_CUDA_CHECK(cudaEventRecord(s->ev[1], st));
    ctx->launches += 8; // own kernels; the two cub calls launch a few more
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int gravityStats(sphgpu_ctx* ctx, sphgpu_gravity_stats* out) {
    GravState* s = gravState(ctx);
    if (!s) {
        setError("self-gravity is not configured");
        return SPHGPU_E_STATE;
    }
    unsigned long long c[4] = { 0, 0, 0, 0 };
    uint32_t groups = 0;
    SPH_CUDA_CHECK(cudaMemcpyAsync(c, s->d.counters, sizeof(unsigned long long) * 3, cudaMemcpyDeviceToHost, ctx->stream));
    SPH_CUDA_CHECK(cudaMemcpyAsync(&groups, s->d.groupCount, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]) != cudaSuccess) {
        cudaGetLastError();
        ms = 0.f;
    }
    if (c[2] != 0ull) {
        setError("gravity tree walk: the per-warp stack overflowed (degenerate tree); results are incomplete");
        return SPHGPU_E_STATE;
    }
    if (out) {
        out->nodes = ctx->n > 1 ? ctx->n - 1 : 0;
        out->groups = groups;
        out->approximated = c[0];
        out->exact = c[1];
        out->gpu_ms = ms;
    }
    return SPHGPU_OK;
}

} // namespace sph

extern "C" {

int sphgpu_gravity_configure(sphgpu_ctx* ctx, const sphgpu_gravity* cfg) {
    if (!ctx) {
        sph::setError("null context");
        return SPHGPU_E_INVALID;
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    SPH_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (cfg) {
        if (cfg->multipole_order != 0 && cfg->multipole_order != 2 && cfg->multipole_order != 3) {
            sph::setError("multipole order must be 0 (monopole), 2 (quadrupole) or 3 (octupole) as in MultipoleOrder (Moments.h:307-312)");
            return SPHGPU_E_INVALID;
        }
        if (cfg->opening_angle > 1.) {
            sph::setError("opening angle must be at most 1: beyond 2/sqrt(3) a node's opening ball no longer contains the node itself");
            return SPHGPU_E_INVALID;
        }
        if (cfg->kernel_radius > 0. && (cfg->lut_grad == nullptr || cfg->lut_entries == 0)) {
            sph::setError("a softening kernel needs its gradient table (lut_grad, lut_entries)");
            return SPHGPU_E_INVALID;
        }
        if (!(cfg->constant > 0.)) {
            sph::setError("the gravitational constant must be positive");
            return SPHGPU_E_INVALID;
        }
        if (ctx->halo != nullptr) {
            sph::setError("self-gravity is global and not available on a decomposed run");
            return SPHGPU_E_INVALID;
        }
    }
    return sph::configureGravity(ctx, cfg);
}

int sphgpu_gravity_eval(sphgpu_ctx* ctx, int accumulate, sphgpu_gravity_stats* stats) {
    if (!ctx) {
        sph::setError("null context");
        return SPHGPU_E_INVALID;
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (!ctx->gravity) {
        sph::setError("self-gravity is not configured (sphgpu_gravity_configure)");
        return SPHGPU_E_STATE;
    }
    const int rc = sph::launchGravity(ctx, accumulate);
    if (rc != SPHGPU_OK) {
        return rc;
    }
    return sph::gravityStats(ctx, stats);
}

int sphgpu_gravity_last_stats(sphgpu_ctx* ctx, sphgpu_gravity_stats* stats) {
    if (!ctx) {
        sph::setError("null context");
        return SPHGPU_E_INVALID;
    }
    SPH_CUDA_CHECK(cudaSetDevice(ctx->device));
    return sph::gravityStats(ctx, stats);
}

} // extern "C"
