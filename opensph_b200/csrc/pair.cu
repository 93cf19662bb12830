// The fused neighbour-search + pair-sum kernels. Replaces the hot loop of the reference
// (AsymmetricSolver::loop functor, core/sph/solvers/AsymmetricSolver.cpp:174-201: finder.findAll, the neighbour
// filter, kernel.grad, derivatives.eval) together with material->initialize (AsymmetricSolver.cpp:75-79),
// accumulated.store + equations.finalize (:204-216) and material->finalize (:90-95).
#include "sphgpu_internal.h"
#include <algorithm>
#include <cstddef>
#include <mutex>

namespace sph {

__constant__ ParamsDev c_prm;
__constant__ MaterialDev c_mats[MAX_MATERIALS];

int uploadConstants(const sphgpu_ctx* ctx) {
    SPH_CUDA_CHECK(cudaMemcpyToSymbol(c_prm, &ctx->prm, sizeof(ParamsDev)));
    SPH_CUDA_CHECK(cudaMemcpyToSymbol(c_mats, ctx->matsHost, sizeof(MaterialDev) * MAX_MATERIALS));
    return SPHGPU_OK;
}

// The run-level constants live in __constant__ symbols, which are per device, not per context. Every entry point that
// launches kernels calls ensureConstants first: if another context of this device launched last, the device is drained
// (that context's kernels may still be reading the symbols) and this context's constants are uploaded. Contexts with
// different settings can therefore coexist on one device; alternating between them costs a device synchronisation.
static std::mutex g_constMutex;
static const sphgpu_ctx* g_constOwner[64] = {};

int ensureConstants(const sphgpu_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_constMutex);
    const int dev = ctx->device & 63;
    if (g_constOwner[dev] == ctx) {
        return SPHGPU_OK;
    }
    if (g_constOwner[dev] != nullptr) {
        SPH_CUDA_CHECK(cudaDeviceSynchronize());
    }
    const int rc = uploadConstants(ctx);
    g_constOwner[dev] = rc == SPHGPU_OK ? ctx : nullptr;
    return rc;
}

void forgetConstants(const sphgpu_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_constMutex);
    const int dev = ctx->device & 63;
    if (g_constOwner[dev] == ctx) {
        g_constOwner[dev] = nullptr;
    }
}

/// FP32 copy of a position relative to the grid origin (the pre-filter of the tiled pair kernel compares these; targets
/// and candidates must be converted by this very expression so that equal positions stay equal).
__device__ __forceinline__ float4 gridRelative(const GridDev& g, double x, double y, double z, double h) {
    return make_float4((float)(x - g.lo[0]), (float)(y - g.lo[1]), (float)(z - g.lo[2]), (float)h);
}

// 64 registers, 4 CTAs per SM: the kernel waits on its loads (ncu: long scoreboard), occupancy is what hides them. Measured
// at 10.6 M particles: 1.11 ms with 3 CTAs, 1.01 ms with 4, 1.06 ms with 5 (spills); hoisting every load above the first
// store made it slower (1.10 ms with 4 CTAs).
#ifndef PROLOGUE_MIN_CTAS
#define PROLOGUE_MIN_CTAS 4
#endif
// ---- prologue: EoS + rheology + damage growth, and packing of the sorted neighbour-input planes -----------
// One thread per SLOT i; its sorted position t = rank[i] (the inverse of `order`, k_sort_cells). The slot planes are read
// and written coalesced; the neighbour inputs (with p/rho^2, S/rho^2, m/rho precomputed) leave as whole 112..144-byte
// records scattered to the sorted array. (Round 1 and most of round 2 ran one thread per sorted position and GATHERED the
// planes through `order`: a cell's particles sit on several lattice rows, so only about 60 % of every 32-byte sector was
// used -- 0.99 ms against 0.6 ms of compulsory traffic at 10.6 M particles.)
/// StressAV::initialize (Stress.cpp:91-109) for the total stress sigma {xx,yy,zz,xy,xz,yz}. Not inlined, and called with
/// scalars: neither the eigen-solver's registers nor an address-taken array may weigh on the prologue of the runs without the
/// term (passing the kernel's S[] or the DevicePointers by reference moved them to local memory: 0.94 -> 2.5 ms, measured).
__device__ __noinline__ void avStressCold(double sxx, double syy, double szz, double sxy, double sxz, double syz, double* as) {
    const double sigma[6] = { sxx, syy, szz, sxy, sxz, syz };
    avStressOf(sigma, as);
}

template <bool SOLID>
__global__ void __launch_bounds__(256, PROLOGUE_MIN_CTAS) k_prologue_pack(DevicePointers d, uint32_t nActive, uint32_t nOwned, bool hasReduce,
    bool hasDamage) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nActive) {
        return;
    }
    const uint32_t t = d.rank[i];
    const MaterialDev& mat = c_mats[d.u[U_MATID][i]];
    const double rho = d.f[F_RHO][i], u = d.f[F_U][i];
    double p = d.f[F_P][i], cs = d.f[F_CS][i];
    evalEos(mat, rho, u, p, cs);
    double S[5] = { 0., 0., 0., 0., 0. };
    double reduce = 1.;
    if (SOLID) {
        for (int k = 0; k < 5; ++k) {
            S[k] = d.f[F_S0 + k][i];
        }
    }
    if (hasReduce) {
        reduce = d.f[F_REDUCE][i];
    }
    if (mat.yielding == SPHGPU_YIELD_VON_MISES) {
        const bool dmg = hasDamage && mat.fracture != SPHGPU_FRACTURE_NONE;
        const double D = dmg ? d.f[F_D][i] : 0.;
        reduce = vonMises(mat, u, D, dmg, p, S);
        d.f[F_REDUCE][i] = reduce;
        if (SOLID) {
            for (int k = 0; k < 5; ++k) {
                d.f[F_S0 + k][i] = S[k];
            }
        }
        if (dmg && mat.fracture == SPHGPU_FRACTURE_SCALAR_GRADY_KIPP && i < nOwned) {
            d.f[F_DD][i] = damageRate(mat, p, S, D, d.f[F_EPSMIN][i], d.f[F_MZERO][i], d.f[F_GROWTH][i], d.u[U_NFLAWS][i]);
        }
    }
    d.f[F_P][i] = p;
    d.f[F_CS][i] = cs;

    const double rhoInv2 = 1. / (rho * rho);
    const double m = d.f[F_M][i];
    const bool balsara = (c_prm.flags & SPHGPU_FLAG_BALSARA) != 0;
    const int RD = recordDoublesOf(SOLID, c_prm.flags);
    double2* rec = reinterpret_cast<double2*>(d.rec + (size_t)t * RD);
    const uint32_t sw = recordSwizzle(RD, t); // XOR swizzle of the 128-byte solid record's pieces (sphgpu_internal.h)
    const double x = d.f[F_X][i], y = d.f[F_Y][i], z = d.f[F_Z][i], h = d.f[F_H][i];
    rec[0 ^ sw] = make_double2(x, y);
    rec[1 ^ sw] = make_double2(z, h);
    const bool rebuild = d.listCtl->rebuild != 0u; // the lists are being rebuilt: new FP32 copies for the search ...
    if (rebuild) {
        const float4 pf = gridRelative(*d.grid, x, y, z, h);
        d.posF[t] = pf;
        d.pos0[i] = pf; // ... and the reference point of the displacement check (k_bounds)
    }
    double vx = d.f[F_VX][i], vy = d.f[F_VY][i], vz = d.f[F_VZ][i];
    if (c_prm.flags & SPHGPU_FLAG_XSPH) { // XSph::initialize (XSph.h:69-79): take the previous correction out of the velocities
        vx -= d.f[F_XSX][i];
        vy -= d.f[F_XSY][i];
        vz -= d.f[F_XSZ][i];
        d.f[F_VX][i] = vx;
        d.f[F_VY][i] = vy;
        d.f[F_VZ][i] = vz;
    }
    rec[2 ^ sw] = make_double2(vx, vy);
    rec[3 ^ sw] = make_double2(vz, rho);
    if (SOLID) {
        const uint32_t flag = d.u[U_FLAG][i];
        if (flag >= GROUP_FLAG_LIMIT) {
            atomicAdd(&d.stats->badFlags, 1u);
        }
        const int grp = (hasReduce && reduce == 0.) ? -1 : (int)flag;
        rec[4 ^ sw] = make_double2(p * rhoInv2, packCsGroup(cs, grp));
        rec[5 ^ sw] = make_double2(m / rho, S[0] * rhoInv2);
        rec[6 ^ sw] = make_double2(S[1] * rhoInv2, S[2] * rhoInv2);
        rec[7 ^ sw] = make_double2(S[3] * rhoInv2, S[4] * rhoInv2);
    } else {
        rec[4] = make_double2(p * rhoInv2, cs);
        rec[5] = make_double2(m / rho, 0.);
    }
    if (balsara) { // the factor of the PREVIOUS evaluation's div v / rot v with the current sound speed (Balsara.h:76-80)
        const double f = balsaraFactor(d.f[F_DIVV][i], d.f[F_ROTX][i], d.f[F_ROTY][i], d.f[F_ROTZ][i], cs, h);
        rec[SOLID ? 8 : 6] = make_double2(f, 0.);
    }
    if (SOLID && (c_prm.flags & SPHGPU_FLAG_STRESS_AV)) { // the artificial stress goes to its planes and, over rho^2, into the record
        double as[6];
        avStressCold(S[0] - p, S[1] - p, (-S[0] - S[1]) - p, S[2], S[3], S[4], as);
        for (int k = 0; k < 6; ++k) {
            d.f[F_AS0 + k][i] = as[k];
        }
        rec[8] = make_double2(as[0] * rhoInv2, as[1] * rhoInv2);
        rec[9] = make_double2(as[2] * rhoInv2, as[3] * rhoInv2);
        rec[10] = make_double2(as[4] * rhoInv2, as[5] * rhoInv2);
    }
    if (c_prm.flags & SPHGPU_FLAG_DELTASPH) { // the density gradient the PREVIOUS evaluation stored (DeltaSph.h:71-74)
        rec[SOLID ? 8 : 6] = make_double2(d.f[F_DGX][i], d.f[F_DGY][i]);
        rec[SOLID ? 9 : 7] = make_double2(d.f[F_DGZ][i], 0.);
    }
    if (rebuild) {
        d.sCell[t] = d.cellOf[i];
    }
}

// Sorted positions only (neighbour-list inspection; does not touch the particle state).
__global__ void __launch_bounds__(256) k_pack_positions(DevicePointers d, uint32_t nActive, int recDoubles) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nActive) {
        return;
    }
    const uint32_t i = d.order[t];
    double2* rec = reinterpret_cast<double2*>(d.rec + (size_t)t * recDoubles);
    const uint32_t sw = recordSwizzle(recDoubles, t);
    rec[0 ^ sw] = make_double2(d.f[F_X][i], d.f[F_Y][i]);
    rec[1 ^ sw] = make_double2(d.f[F_Z][i], d.f[F_H][i]);
    d.sCell[t] = d.cellOf[i];
}

/// Unpacks the neighbour-input record with sorted index t from global memory (layouts: sphgpu_internal.h).
template <bool SOLID>
__device__ __forceinline__ void loadRecord(const double* __restrict__ recBase, uint32_t t, int recDoubles, uint32_t flags, Particle& p) {
    const double2* r = reinterpret_cast<const double2*>(recBase + (size_t)t * recDoubles);
    const uint32_t sw = recordSwizzle(recDoubles, t);
    p.bal = 0.;
    if (SOLID) {
        const double2 a = r[0 ^ sw], b = r[1 ^ sw], c = r[2 ^ sw], e = r[3 ^ sw], f = r[4 ^ sw], g = r[5 ^ sw], s1 = r[6 ^ sw],
                      s2 = r[7 ^ sw];
        p.x = a.x; p.y = a.y; p.z = b.x; p.h = b.y;
        p.vx = c.x; p.vy = c.y; p.vz = e.x; p.rho = e.y;
        p.P = f.x;
        unpackCsGroup(f.y, p.cs, p.grp);
        p.vol = g.x;
        p.Sr[0] = g.y; p.Sr[1] = s1.x; p.Sr[2] = s1.y; p.Sr[3] = s2.x; p.Sr[4] = s2.y;
        if (recDoubles == REC_SOLID_BALSARA) {
            p.bal = r[8].x;
        }
        if (flags & SPHGPU_FLAG_DELTASPH) { // (the 176-byte layouts are told apart by the run flags)
            const double2 ga = r[8], gb = r[9];
            p.gr[0] = ga.x; p.gr[1] = ga.y; p.gr[2] = gb.x;
        }
        if (flags & SPHGPU_FLAG_STRESS_AV) {
            const double2 aa = r[8], ab = r[9], ac = r[10];
            p.as[0] = aa.x; p.as[1] = aa.y; p.as[2] = ab.x; p.as[3] = ab.y; p.as[4] = ac.x; p.as[5] = ac.y;
        }
    } else {
        const double2 a = r[0], b = r[1], c = r[2], e = r[3], f = r[4], g = r[5];
        p.x = a.x; p.y = a.y; p.z = b.x; p.h = b.y;
        p.vx = c.x; p.vy = c.y; p.vz = e.x; p.rho = e.y;
        p.P = f.x; p.cs = f.y; p.vol = g.x;
        p.grp = 0;
        p.bal = r[6].x; // (the padding piece: the Balsara factor when the switch is on, unused otherwise)
        if (flags & SPHGPU_FLAG_DELTASPH) {
            const double2 ga = r[6], gb = r[7];
            p.gr[0] = ga.x; p.gr[1] = ga.y; p.gr[2] = gb.x;
        }
    }
    p.m = p.vol * p.rho;
}

template <bool SOLID>
__device__ __forceinline__ void loadSorted(const DevicePointers& d, uint32_t t, Particle& p) {
    loadRecord<SOLID>(d.rec, t, recordDoublesOf(SOLID, c_prm.flags), c_prm.flags, p);
    if (c_prm.flags & SPHGPU_FLAG_STRESS_AV) { // (read for the target only; the direct kernels are not the hot path)
        p.wpInv = 1. / d.f[F_WP][d.order[t]];
    }
}

/// Position pieces {x, y}, {z, h} of the sorted record t.
__device__ __forceinline__ void loadSortedPosition(const double* __restrict__ rec, uint32_t t, int recDoubles, double2& pxy, double2& pzh) {
    const double2* r = reinterpret_cast<const double2*>(rec + (size_t)t * recDoubles);
    const uint32_t sw = recordSwizzle(recDoubles, t);
    pxy = r[0 ^ sw];
    pzh = r[1 ^ sw];
}

/// XSph::finalize (XSph.h:81-90): the new correction is stored and added to the velocity (which the prologue left pure).
__device__ __forceinline__ void storeXsph(const DevicePointers& d, uint32_t i, const double xs[3]) {
    d.f[F_XSX][i] = xs[0];
    d.f[F_XSY][i] = xs[1];
    d.f[F_XSZ][i] = xs[2];
    d.f[F_VX][i] += xs[0];
    d.f[F_VY][i] += xs[1];
    d.f[F_VZ][i] += xs[2];
}

template <bool SOLID, bool CORRECTED>
__device__ __forceinline__ void storeDerivs(const DevicePointers& d, uint32_t i, const Derivs& o) {
    d.f[F_AX][i] = o.ax;
    d.f[F_AY][i] = o.ay;
    d.f[F_AZ][i] = o.az;
    d.f[F_VH][i] = o.vh;
    d.f[F_DU][i] = o.du;
    d.f[F_DRHO][i] = o.drho;
    d.f[F_DIVV][i] = o.divv;
    d.u[U_NCNT][i] = o.ncnt;
    if (c_prm.flags & SPHGPU_FLAG_BALSARA) {
        d.f[F_ROTX][i] = o.rot[0];
        d.f[F_ROTY][i] = o.rot[1];
        d.f[F_ROTZ][i] = o.rot[2];
    }
    if (c_prm.flags & SPHGPU_FLAG_DELTASPH) { // read by the prologue of the NEXT evaluation only (the records hold the old one)
        d.f[F_DGX][i] = o.dg[0];
        d.f[F_DGY][i] = o.dg[1];
        d.f[F_DGZ][i] = o.dg[2];
    }
    if (SOLID) {
        for (int k = 0; k < 5; ++k) {
            d.f[F_DS0 + k][i] = o.dS[k];
        }
        for (int k = 0; k < 6; ++k) {
            d.f[F_GV0 + k][i] = o.gradv[k];
        }
        if (CORRECTED) {
            for (int k = 0; k < 6; ++k) {
                d.f[F_C0 + k][i] = o.corr[k];
            }
        }
    }
}

__device__ __forceinline__ void neighbourStats(const DevicePointers& d, uint32_t cnt, bool valid) {
    // warp-aggregated min / max / sum of NEIGHBOR_CNT (AsymmetricSolver.cpp:218-225)
    const unsigned mask = __activemask();
    uint32_t mn = valid ? cnt : 0xffffffffu, mx = valid ? cnt : 0u, sm = valid ? cnt : 0u;
    mn = __reduce_min_sync(mask, mn);
    mx = __reduce_max_sync(mask, mx);
    sm = __reduce_add_sync(mask, sm);
    if ((threadIdx.x & 31) == (__ffs(mask) - 1)) {
        atomicMin(&d.stats->neighMin, mn);
        atomicMax(&d.stats->neighMax, mx);
        atomicAdd(&d.stats->pairCount, (unsigned long long)sm);
    }
}

// ---- direct evaluation of one target: candidates streamed from the sorted records through L1 / L2 -------------------
// Simple and obviously correct: the cross-check for the tiled kernels (variant 1) and the path of the work units whose
// candidate lists did not fit the list pool (k_pair_fallback, pair_tiled.cu).
template <bool SOLID, bool CORRECTED, bool FILTER>
__device__ __forceinline__ void directTarget(const DevicePointers& d, uint32_t t, uint32_t i) {
    Accum acc;
    const GridDev g = *d.grid;
    if (g.nLarge > 0u) {
        acc = d.accLarge[t]; // sums over the large neighbours (two-level radii), k_large_neighbours
    } else {
        accumZero(acc);
    }
    Particle pi;
    loadSorted<SOLID>(d, t, pi);
    const uint32_t c = d.sCell[t];
    const int cx = (int)(c % (uint32_t)g.dim[0]);
    const int cy = (int)((c / (uint32_t)g.dim[0]) % (uint32_t)g.dim[1]);
    const int cz = (int)(c / ((uint32_t)g.dim[0] * (uint32_t)g.dim[1]));
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
    const int RD = recordDoublesOf(SOLID, c_prm.flags);
    for (int z = max(cz - 2, 0); z <= min(cz + 2, g.dim[2] - 1); ++z) {
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
            const uint32_t row = (uint32_t)((z * g.dim[1] + y) * g.dim[0]);
            const uint32_t s = d.cellStart[row + x0], e = d.cellStart[row + x1 + 1];
            for (uint32_t k = s; k < e; ++k) {
                if (k == t) {
                    continue;
                }
                double2 pxy, pzh;
                loadSortedPosition(d.rec, k, RD, pxy, pzh);
                const double dx = pi.x - pxy.x, dy = pi.y - pxy.y, dz = pi.z - pzh.x;
                double d2, hbar;
                if (!isNeighbour(dx, dy, dz, pi.h, pzh.y, c_prm.kernel_radius, d2, hbar)) {
                    continue;
                }
                Particle pj;
                loadSorted<SOLID>(d, k, pj);
                pairAccumulate<SOLID, CORRECTED, FILTER>(c_prm, d.lut, d.lutW, pi, pj, dx, dy, dz, d2, hbar, acc);
            }
        }
    }
    const MaterialDev& mat = c_mats[d.u[U_MATID][i]];
    double S[5] = { 0., 0., 0., 0., 0. };
    if (SOLID) {
        for (int k = 0; k < 5; ++k) {
            S[k] = d.f[F_S0 + k][i];
        }
    }
    Derivs out;
    finalizeParticle<SOLID, CORRECTED>(c_prm, mat, acc, pi.h, pi.rho, d.f[F_P][i], d.f[F_CS][i], SOLID ? d.f[F_REDUCE][i] : 1., S, out);
    if (c_prm.flags & SPHGPU_FLAG_DELTASPH) {
        finalizeDeltaSph<SOLID>(acc, out);
    }
    storeDerivs<SOLID, CORRECTED>(d, i, out);
    if (c_prm.flags & SPHGPU_FLAG_XSPH) {
        storeXsph(d, i, acc.xs);
    }
    // neighbour statistics (AsymmetricSolver.cpp:218-225); the callers' warps are divergent here, so plain atomics
    atomicMin(&d.stats->neighMin, acc.cnt);
    atomicMax(&d.stats->neighMax, acc.cnt);
    atomicAdd(&d.stats->pairCount, (unsigned long long)acc.cnt);
}

template <bool SOLID, bool CORRECTED, bool FILTER>
__global__ void __launch_bounds__(128) k_pair_direct(DevicePointers d, uint32_t nActive, uint32_t nOwned) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nActive) {
        return;
    }
    const uint32_t i = d.order[t];
    if (i < nOwned && !(d.grid->nLarge > 0u && t >= d.grid->largeBegin)) { // ghosts are neighbours only; large targets: k_large_targets
        directTarget<SOLID, CORRECTED, FILTER>(d, t, i);
    }
}

// ---- two-level search radii: every pair that involves a LARGE particle (GridDev::hSplit) ---------------------------------
// k_large_neighbours: one thread per small target, loops over the (few) large particles and leaves the partial sums in
// accLarge, from which the tiled / direct kernels start instead of from zero. k_large_targets: one CTA per large target,
// the threads share ALL other particles, a fixed-order reduction combines the partial sums. Both apply the exact
// predicate of AsymmetricSolver.cpp:186-191; the pair arithmetic is pairAccumulate (as in the direct variant).
template <bool SOLID, bool CORRECTED, bool FILTER>
__global__ void __launch_bounds__(128) k_large_neighbours(DevicePointers d, uint32_t nActive) {
    const GridDev g = *d.grid;
    if (g.nLarge == 0u) {
        return;
    }
    const int RD = recordDoublesOf(SOLID, c_prm.flags);
    // (a small grid with a stride loop: in the usual case -- no large particles -- the launch must cost next to nothing)
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < g.largeBegin; t += gridDim.x * blockDim.x) {
        Accum acc;
        accumZero(acc);
        Particle pi;
        loadSorted<SOLID>(d, t, pi);
        for (uint32_t k = g.largeBegin; k < nActive; ++k) {
            double2 pxy, pzh;
            loadSortedPosition(d.rec, k, RD, pxy, pzh);
            const double dx = pi.x - pxy.x, dy = pi.y - pxy.y, dz = pi.z - pzh.x;
            double d2, hbar;
            if (!isNeighbour(dx, dy, dz, pi.h, pzh.y, c_prm.kernel_radius, d2, hbar)) {
                continue;
            }
            Particle pj;
            loadSorted<SOLID>(d, k, pj);
            pairAccumulate<SOLID, CORRECTED, FILTER>(c_prm, d.lut, d.lutW, pi, pj, dx, dy, dz, d2, hbar, acc);
        }
        d.accLarge[t] = acc;
    }
}

template <bool SOLID, bool CORRECTED, bool FILTER>
__global__ void __launch_bounds__(256) k_large_targets(DevicePointers d, uint32_t nActive, uint32_t nOwned) {
    const GridDev g = *d.grid;
    if (g.nLarge == 0u) {
        return;
    }
    // The grid's CTAs are dealt to the large targets: target L gets S = gridDim / nLarge CTAs, each summing one slice of
    // the particles; the CTA that finishes last adds the slices' partial sums in slice order (the result does not depend
    // on which one that is) and runs the finalizers.
    const int RD = recordDoublesOf(SOLID, c_prm.flags);
    constexpr int NV = (int)(offsetof(Accum, cnt) / sizeof(double)); // doubles of Accum in front of the counter
    __shared__ double red[8][NV];
    __shared__ uint32_t redCnt[8];
    __shared__ bool isLast;
    const uint32_t nL = g.nLarge;
    const uint32_t S = max(gridDim.x / nL, 1u);
    const uint32_t per = (nActive + S - 1u) / S;
    for (uint32_t w = blockIdx.x; w < nL * S; w += gridDim.x) {
        const uint32_t L = w % nL, slice = w / nL;
        const uint32_t t = g.largeBegin + L;
        const uint32_t i = d.order[t];
        if (i >= nOwned) {
            continue; // ghosts are neighbours only (uniform for the CTA)
        }
        Particle pi;
        loadSorted<SOLID>(d, t, pi);
        Accum acc;
        accumZero(acc);
        const uint32_t kEnd = min((slice + 1u) * per, nActive);
        for (uint32_t k = slice * per + threadIdx.x; k < kEnd; k += blockDim.x) {
            if (k == t) {
                continue;
            }
            double2 pxy, pzh;
            loadSortedPosition(d.rec, k, RD, pxy, pzh);
            const double dx = pi.x - pxy.x, dy = pi.y - pxy.y, dz = pi.z - pzh.x;
            double d2, hbar;
            if (!isNeighbour(dx, dy, dz, pi.h, pzh.y, c_prm.kernel_radius, d2, hbar)) {
                continue;
            }
            Particle pj;
            loadSorted<SOLID>(d, k, pj);
            pairAccumulate<SOLID, CORRECTED, FILTER>(c_prm, d.lut, d.lutW, pi, pj, dx, dy, dz, d2, hbar, acc);
        }
        // fixed-order reduction inside the CTA: lanes by shuffle, warps through shared memory
        double* a = reinterpret_cast<double*>(&acc);
        uint32_t cnt = acc.cnt;
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                a[q] += __shfl_down_sync(0xffffffffu, a[q], o);
            }
            cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __syncthreads(); // (the previous reduction has been read)
        if (lane == 0) {
            for (int q = 0; q < NV; ++q) {
                red[warp][q] = a[q];
            }
            redCnt[warp] = cnt;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int wp = 1; wp < 8; ++wp) {
                for (int q = 0; q < NV; ++q) {
                    a[q] += red[wp][q];
                }
                cnt += redCnt[wp];
            }
            acc.cnt = cnt;
            d.largePartial[slice * nL + L] = acc;
            __threadfence();
            isLast = atomicAdd(&d.largeCounter[L], 1u) == S - 1u;
            if (isLast) {
                d.largeCounter[L] = 0u;
                __threadfence();
                Accum sum;
                accumZero(sum);
                double* b = reinterpret_cast<double*>(&sum);
                for (uint32_t sl = 0; sl < S; ++sl) {
                    const Accum part = d.largePartial[sl * nL + L];
                    const double* p = reinterpret_cast<const double*>(&part);
                    for (int q = 0; q < NV; ++q) {
                        b[q] += p[q];
                    }
                    sum.cnt += part.cnt;
                }
                const MaterialDev& mat = c_mats[d.u[U_MATID][i]];
                double Sv[5] = { 0., 0., 0., 0., 0. };
                if (SOLID) {
                    for (int q = 0; q < 5; ++q) {
                        Sv[q] = d.f[F_S0 + q][i];
                    }
                }
                Derivs out;
                finalizeParticle<SOLID, CORRECTED>(c_prm, mat, sum, pi.h, pi.rho, d.f[F_P][i], d.f[F_CS][i], SOLID ? d.f[F_REDUCE][i] : 1., Sv, out);
                if (c_prm.flags & SPHGPU_FLAG_DELTASPH) {
                    finalizeDeltaSph<SOLID>(sum, out);
                }
                storeDerivs<SOLID, CORRECTED>(d, i, out);
                if (c_prm.flags & SPHGPU_FLAG_XSPH) {
                    storeXsph(d, i, sum.xs);
                }
                atomicMin(&d.stats->neighMin, sum.cnt);
                atomicMax(&d.stats->neighMax, sum.cnt);
                atomicAdd(&d.stats->pairCount, (unsigned long long)sum.cnt);
            }
        }
    }
}

template <bool SOLID, bool CORRECTED, bool FILTER>
static int launchLargeVariant(sphgpu_ctx* ctx) {
    const uint32_t n = ctx->nActive;
    k_large_neighbours<SOLID, CORRECTED, FILTER><<<std::min<uint32_t>((n + 127) / 128, 148u * 8u), 128, 0, ctx->stream>>>(ctx->d, n);
    k_large_targets<SOLID, CORRECTED, FILTER><<<(unsigned)LARGE_MAX, 256, 0, ctx->stream>>>(ctx->d, n, ctx->n);
    ctx->launches += 2;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

/// Pairs with large particles (returns at once on the device when there are none).
int launchLargePairs(sphgpu_ctx* ctx) {
    if (ctx->nActive == 0) {
        return SPHGPU_OK;
    }
    if (!ctx->solid) {
        return launchLargeVariant<false, false, false>(ctx);
    }
    if (ctx->corrected) {
        return ctx->filter ? launchLargeVariant<true, true, true>(ctx) : launchLargeVariant<true, true, false>(ctx);
    }
    return ctx->filter ? launchLargeVariant<true, false, true>(ctx) : launchLargeVariant<true, false, false>(ctx);
}

// ---- neighbour lists for the tests ---------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(128) k_neighbour_lists(DevicePointers d, uint32_t nActive, uint32_t nOwned, uint32_t* counts,
    const unsigned long long* offsets, uint32_t* idx, int recDoubles) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nActive) {
        return;
    }
    const uint32_t i = d.order[t];
    if (i >= nOwned) {
        return;
    }
    const GridDev g = *d.grid;
    double2 ixy, izh;
    loadSortedPosition(d.rec, t, recDoubles, ixy, izh);
    const double xi = ixy.x, yi = ixy.y, zi = izh.x, hi = izh.y;
    uint32_t cnt = 0;
    const unsigned long long base = FILL ? offsets[i] : 0ull;
    auto test = [&](uint32_t k) {
        double2 pxy, pzh;
        loadSortedPosition(d.rec, k, recDoubles, pxy, pzh);
        double d2, hbar;
        if (isNeighbour(xi - pxy.x, yi - pxy.y, zi - pzh.x, hi, pzh.y, c_prm.kernel_radius, d2, hbar)) {
            if (FILL) {
                idx[base + cnt] = d.order[k];
            }
            cnt++;
        }
    };
    const bool large = g.nLarge > 0u && t >= g.largeBegin;
    if (large) { // a large particle (two-level radii): everything can be its neighbour
        for (uint32_t k = 0; k < nActive; ++k) {
            if (k != t) {
                test(k);
            }
        }
    } else {
        const uint32_t c = d.sCell[t];
        const int cx = (int)(c % (uint32_t)g.dim[0]);
        const int cy = (int)((c / (uint32_t)g.dim[0]) % (uint32_t)g.dim[1]);
        const int cz = (int)(c / ((uint32_t)g.dim[0] * (uint32_t)g.dim[1]));
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
        for (int z = max(cz - 2, 0); z <= min(cz + 2, g.dim[2] - 1); ++z) {
            for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
                const uint32_t row = (uint32_t)((z * g.dim[1] + y) * g.dim[0]);
                const uint32_t s = d.cellStart[row + x0], e = d.cellStart[row + x1 + 1];
                for (uint32_t k = s; k < e; ++k) {
                    if (k != t) {
                        test(k);
                    }
                }
            }
        }
        if (g.nLarge > 0u) {
            for (uint32_t k = g.largeBegin; k < nActive; ++k) {
                test(k);
            }
        }
    }
    if (!FILL) {
        counts[i] = cnt;
    }
}

int launchProloguePack(sphgpu_ctx* ctx) {
    const uint32_t n = ctx->nActive;
    if (n == 0) {
        return SPHGPU_OK;
    }
    const uint32_t blocks = (n + 255) / 256;
    if (ctx->solid) {
        k_prologue_pack<true><<<blocks, 256, 0, ctx->stream>>>(ctx->d, n, ctx->n, ctx->hasReduce, ctx->hasDamage);
    } else {
        k_prologue_pack<false><<<blocks, 256, 0, ctx->stream>>>(ctx->d, n, ctx->n, ctx->hasReduce, ctx->hasDamage);
    }
    ctx->launches += 1;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchProloguePackPositionsOnly(sphgpu_ctx* ctx) {
    const uint32_t n = ctx->nActive;
    if (n == 0) {
        return SPHGPU_OK;
    }
    k_pack_positions<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d, n, ctx->recDoubles);
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchPairTiled(sphgpu_ctx* ctx); // pair_tiled.cu

int launchPair(sphgpu_ctx* ctx) {
    const uint32_t n = ctx->nActive; // (the statistics were reset at the start of enqueueIntegrate)
    if (n == 0) {
        return SPHGPU_OK;
    }
    ctx->pairTimed = false;
    {
        const int rcLarge = launchLargePairs(ctx);
        if (rcLarge != SPHGPU_OK) {
            return rcLarge;
        }
    }
    if (ctx->variant == 4) { // every pair once, both particles (SymmetricSolver)
        return launchPairSymmetric(ctx);
    }
    if (ctx->variant != 1) {
        return launchPairTiled(ctx);
    }
    const uint32_t blocks = (n + 127) / 128;
    cudaStream_t st = ctx->stream;
    if (!ctx->solid) {
        k_pair_direct<false, false, false><<<blocks, 128, 0, st>>>(ctx->d, n, ctx->n);
    } else if (ctx->corrected) {
        if (ctx->filter) {
            k_pair_direct<true, true, true><<<blocks, 128, 0, st>>>(ctx->d, n, ctx->n);
        } else {
            k_pair_direct<true, true, false><<<blocks, 128, 0, st>>>(ctx->d, n, ctx->n);
        }
    } else {
        if (ctx->filter) {
            k_pair_direct<true, false, true><<<blocks, 128, 0, st>>>(ctx->d, n, ctx->n);
        } else {
            k_pair_direct<true, false, false><<<blocks, 128, 0, st>>>(ctx->d, n, ctx->n);
        }
    }
    ctx->launches += 1;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchNeighbourCount(sphgpu_ctx* ctx, uint32_t* countsDev) {
    const uint32_t n = ctx->nActive;
    if (n == 0) {
        return SPHGPU_OK;
    }
    k_neighbour_lists<false><<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d, n, ctx->n, countsDev, nullptr, nullptr, ctx->recDoubles);
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchNeighbourFill(sphgpu_ctx* ctx, const unsigned long long* offsetsDev, uint32_t* idxDev) {
    const uint32_t n = ctx->nActive;
    if (n == 0) {
        return SPHGPU_OK;
    }
    k_neighbour_lists<true><<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d, n, ctx->n, nullptr, offsetsDev, idxDev, ctx->recDoubles);
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

} // namespace sph
