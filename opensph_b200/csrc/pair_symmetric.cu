// The symmetric formulation, variant 4 of the pair stage: every pair is evaluated ONCE
// and contributes to both particles, as SymmetricSolver::loop does it (core/sph/solvers/SymmetricSolver.cpp:104-143):
//   * rank of every particle in the order of the smoothing lengths -- makeRankH (core/objects/finders/Order.h:46-58,
//     NeighborFinder.cpp:11-26): here a radix sort of the bit patterns of h (stable: equal h keep their order, any strict
//     order consistent with h assigns every pair to exactly one of its two particles, which is all findLowerRank needs);
//   * particle i takes the neighbours of LOWER rank within R h_i (ISymmetricFinder::findLowerRank), drops i == j and
//     d^2 >= (R hbar)^2 (SymmetricSolver.cpp:125-133), computes one kernel gradient and calls evalSymmetric: the sums of i
//     AND of j receive the pair (VelocityDivergence / VelocityGradient: DerivativeHelpers.h:100-146,292-304;
//     PressureGradient / StressDivergence through AccelerationTemplate::evalSymmetric, DerivativeHelpers.h:165-285;
//     StandardAV, av/Standard.h:63-83; NeighborCountTerm, HelperTerms.h:16-47: ++cnt_i, ++cnt_j);
//   * the reference gives every thread its own Accumulated and adds them up afterwards (SymmetricSolver.cpp:145-163); here
//     the partner's share goes to a per-particle array with FP64 atomics and a second kernel finalizes.
// This is the faithful counterpart, not the fast path: the asymmetric kernels (pair_tiled.cu) evaluate both directions of
// a pair without atomics and are several times faster; results agree to rounding (the reference's own test,
// core/sph/solvers/test/Solvers.cpp:178-216, asks for 1e-12). The atomics make the summation order -- not the result
// beyond rounding -- vary from run to run. Like SymmetricSolver it does not offer the strain-rate correction tensor
// (SymmetricSolver.cpp:41-44); the Balsara switch and XSph are not wired to it either.
#include "sphgpu_internal.h"
#include <cub/device/device_radix_sort.cuh>

namespace sph {

constexpr int SYM_DOUBLES = 14; // ax ay az du divv T[9]

struct SymState {
    unsigned long long *keys = nullptr, *keysSorted = nullptr;
    uint32_t *vals = nullptr, *valsSorted = nullptr, *rank = nullptr, *cnt = nullptr;
    double* acc = nullptr; // [SYM_DOUBLES * capacity], plane q at q * capacity
    void* cubTemp = nullptr;
    size_t cubTempBytes = 0;
    uint32_t capacity = 0;
};

/// Sort key of the sorted particle t: the bit pattern of its (positive) smoothing length orders like the value.
__global__ void __launch_bounds__(256) k_sym_keys(DevicePointers d, uint32_t nActive, int recDoubles, unsigned long long* keys, uint32_t* vals) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nActive) {
        return;
    }
    double2 pxy, pzh;
    loadSortedPosition(d.rec, t, recDoubles, pxy, pzh);
    keys[t] = (unsigned long long)__double_as_longlong(pzh.y);
    vals[t] = t;
}

/// rank = inverse of the sorted order (Order::getInverted), and the start values of the sums: zero, or the partial sums over
/// the LARGE neighbours (two-level radii: those pairs are evaluated by k_large_neighbours / k_large_targets).
__global__ void __launch_bounds__(256) k_sym_rank(DevicePointers d, uint32_t nActive, uint32_t capacity, const uint32_t* valsSorted, uint32_t* rank,
    double* acc, uint32_t* cnt) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nActive) {
        return;
    }
    rank[valsSorted[p]] = p;
    const bool large = d.grid->nLarge > 0u;
    Accum a;
    if (large) {
        a = d.accLarge[p];
    } else {
        accumZero(a);
    }
    double* out = acc + p;
    out[0] = a.ax;
    out[(size_t)capacity] = a.ay;
    out[2 * (size_t)capacity] = a.az;
    out[3 * (size_t)capacity] = a.du;
    out[4 * (size_t)capacity] = a.divv;
    for (int q = 0; q < 9; ++q) {
        out[(5 + q) * (size_t)capacity] = a.T[q];
    }
    cnt[p] = a.cnt;
}

/// One pair, both directions (evalSymmetric). `i` receives the sums of particle i, `j` the share of its partner.
template <bool SOLID, bool FILTER>
__device__ __forceinline__ void pairSymmetric(const ParamsDev& prm, const double* __restrict__ lut, const Particle& pi, const Particle& pj, double dx,
    double dy, double dz, double d2, double hbar, Accum& i, Accum& j) {
    const double hInv = 1. / hbar;
    const double hInv2 = hInv * hInv;
    const double qSqr = d2 * hInv2;
    double G = 0.;
    if (qSqr < prm.radius_sqr) {
        const double fidx = prm.q_sqr_to_idx * qSqr;
        const uint32_t k = (uint32_t)fidx;
        const double ratio = fidx - (double)k;
        G = __ldg(lut + k) * (1. - ratio) + __ldg(lut + k + 1) * ratio;
    }
    const double s = hInv2 * hInv2 * hInv * G;
    const double gx = dx * s, gy = dy * s, gz = dz * s; // grad W_ij, shared by both particles
    const double dvx = pj.vx - pi.vx, dvy = pj.vy - pi.vy, dvz = pj.vz - pi.vz;
    const double dvg = dvx * gx + dvy * gy + dvz * gz;
    i.cnt++;
    j.cnt++;
    i.divv += pj.m * dvg;
    j.divv += pi.m * dvg;
    // StandardAV (Standard.h:63-83): the same Pi for both; heating 1/2 Pi (v_i - v_j).grad is added to both energies
    const double w = -(dvx * dx + dvy * dy + dvz * dz);
    double Pi = 0.;
    if (w < 0.) {
        const double rhobar = 0.5 * (pi.rho + pj.rho);
        const double csbar = 0.5 * (pi.cs + pj.cs);
        const double mu = hbar * w / (d2 + 1.e-2 * hbar * hbar);
        Pi = (-prm.av_alpha * csbar * mu + prm.av_beta * mu * mu) / rhobar;
        const double heating = -0.5 * Pi * dvg;
        i.du += pj.m * heating;
        j.du += pi.m * heating;
    }
    // f = -(P_i + P_j + Pi) grad:  dv_i += m_j f,  dv_j -= m_i f
    const double c = pi.P + pj.P + Pi;
    double fx = -c * gx, fy = -c * gy, fz = -c * gz;
    if (SOLID) {
        bool ok = true;
        if (FILTER) {
            ok = (pi.grp == pj.grp) && (pi.grp >= 0);
        }
        if (ok) {
            const double sxx = pi.Sr[0] + pj.Sr[0], syy = pi.Sr[1] + pj.Sr[1], sxy = pi.Sr[2] + pj.Sr[2], sxz = pi.Sr[3] + pj.Sr[3],
                         syz = pi.Sr[4] + pj.Sr[4];
            const double szz = -sxx - syy;
            fx += sxx * gx + sxy * gy + sxz * gz;
            fy += sxy * gx + syy * gy + syz * gz;
            fz += sxz * gx + syz * gy + szz * gz;
            // velocity gradient: m/rho sym(dv (x) grad) for both (the division by rho_i / rho_j happens in the finalizer)
            const double t[9] = { dvx * gx, dvx * gy, dvx * gz, dvy * gx, dvy * gy, dvy * gz, dvz * gx, dvz * gy, dvz * gz };
            for (int q = 0; q < 9; ++q) {
                i.T[q] += pj.m * t[q];
                j.T[q] += pi.m * t[q];
            }
        }
    }
    i.ax += pj.m * fx;
    i.ay += pj.m * fy;
    i.az += pj.m * fz;
    j.ax -= pi.m * fx;
    j.ay -= pi.m * fy;
    j.az -= pi.m * fz;
}

__device__ __forceinline__ void symAtomicAdd(double* acc, uint32_t* cnt, uint32_t capacity, uint32_t t, const Accum& a, bool solid) {
    double* out = acc + t;
    atomicAdd(out, a.ax);
    atomicAdd(out + (size_t)capacity, a.ay);
    atomicAdd(out + 2 * (size_t)capacity, a.az);
    atomicAdd(out + 3 * (size_t)capacity, a.du);
    atomicAdd(out + 4 * (size_t)capacity, a.divv);
    if (solid) {
        for (int q = 0; q < 9; ++q) {
            atomicAdd(out + (5 + q) * (size_t)capacity, a.T[q]);
        }
    }
    atomicAdd(cnt + t, a.cnt);
}

/// SymmetricSolver::loop (SymmetricSolver.cpp:120-140): one thread per particle (ghosts included: a ghost of higher rank
/// owns the pair and delivers the share of its owned partner).
template <bool SOLID, bool FILTER>
__global__ void __launch_bounds__(128) k_pair_symmetric(DevicePointers d, uint32_t nActive, uint32_t capacity, const uint32_t* __restrict__ rank,
    double* acc, uint32_t* cnt) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nActive) {
        return;
    }
    const GridDev g = *d.grid;
    if (g.nLarge > 0u && t >= g.largeBegin) {
        return; // large particles: every pair of theirs is evaluated by the direct kernels of the two-level radii
    }
    Particle pi;
    loadSorted<SOLID>(d, t, pi);
    const uint32_t myRank = rank[t];
    const uint32_t c = d.sCell[t];
    const int cx = (int)(c % (uint32_t)g.dim[0]);
    const int cy = (int)((c / (uint32_t)g.dim[0]) % (uint32_t)g.dim[1]);
    const int cz = (int)(c / ((uint32_t)g.dim[0] * (uint32_t)g.dim[1]));
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
    const int RD = recordDoubles(SOLID, false);
    Accum mine;
    accumZero(mine);
    for (int z = max(cz - 2, 0); z <= min(cz + 2, g.dim[2] - 1); ++z) {
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
            const uint32_t row = (uint32_t)((z * g.dim[1] + y) * g.dim[0]);
            const uint32_t s = d.cellStart[row + x0], e = d.cellStart[row + x1 + 1];
            for (uint32_t k = s; k < e; ++k) {
                if (k == t || rank[k] >= myRank) {
                    continue; // findLowerRank: the pair belongs to the particle of higher rank
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
                Accum theirs;
                accumZero(theirs);
                pairSymmetric<SOLID, FILTER>(c_prm, d.lut, pi, pj, dx, dy, dz, d2, hbar, mine, theirs);
                symAtomicAdd(acc, cnt, capacity, k, theirs, SOLID);
            }
        }
    }
    symAtomicAdd(acc, cnt, capacity, t, mine, SOLID);
}

/// accumulated.store + equations.finalize for the owned particles (afterLoop, SymmetricSolver.cpp:165-203).
template <bool SOLID>
__global__ void __launch_bounds__(128) k_sym_finalize(DevicePointers d, uint32_t nActive, uint32_t nOwned, uint32_t capacity, const double* acc,
    const uint32_t* cnt) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nActive) {
        return;
    }
    const uint32_t i = d.order[t];
    if (i >= nOwned || (d.grid->nLarge > 0u && t >= d.grid->largeBegin)) {
        return;
    }
    Accum a;
    accumZero(a);
    const double* in = acc + t;
    a.ax = in[0];
    a.ay = in[(size_t)capacity];
    a.az = in[2 * (size_t)capacity];
    a.du = in[3 * (size_t)capacity];
    a.divv = in[4 * (size_t)capacity];
    for (int q = 0; q < 9; ++q) {
        a.T[q] = in[(5 + q) * (size_t)capacity];
    }
    a.cnt = cnt[t];
    Particle pi;
    loadSorted<SOLID>(d, t, pi);
    const MaterialDev& mat = c_mats[d.u[U_MATID][i]];
    double S[5] = { 0., 0., 0., 0., 0. };
    if (SOLID) {
        for (int k = 0; k < 5; ++k) {
            S[k] = d.f[F_S0 + k][i];
        }
    }
    Derivs out;
    finalizeParticle<SOLID, false>(c_prm, mat, a, pi.h, pi.rho, d.f[F_P][i], d.f[F_CS][i], SOLID ? d.f[F_REDUCE][i] : 1., S, out);
    storeDerivs<SOLID, false>(d, i, out);
    atomicMin(&d.stats->neighMin, a.cnt);
    atomicMax(&d.stats->neighMax, a.cnt);
    atomicAdd(&d.stats->pairCount, (unsigned long long)a.cnt);
}

void destroySymmetric(sphgpu_ctx* ctx) {
    SymState* s = static_cast<SymState*>(ctx->symmetric);
    if (!s) {
        return;
    }
    cudaFree(s->keys); cudaFree(s->keysSorted); cudaFree(s->vals); cudaFree(s->valsSorted); cudaFree(s->rank); cudaFree(s->cnt);
    cudaFree(s->acc); cudaFree(s->cubTemp);
    cudaGetLastError();
    delete s;
    ctx->symmetric = nullptr;
}

static int ensureSymmetric(sphgpu_ctx* ctx) {
    if (ctx->symmetric) {
        return SPHGPU_OK;
    }
    SymState* s = new (std::nothrow) SymState();
    if (!s) {
        setError("host allocation failed");
        return SPHGPU_E_OOM;
    }
    ctx->symmetric = s;
    const size_t cap = std::max<size_t>(ctx->capacity, 1);
    s->capacity = (uint32_t)cap;
    SPH_CUDA_CHECK(cudaMalloc(&s->keys, cap * 8));
    SPH_CUDA_CHECK(cudaMalloc(&s->keysSorted, cap * 8));
    SPH_CUDA_CHECK(cudaMalloc(&s->vals, cap * 4));
    SPH_CUDA_CHECK(cudaMalloc(&s->valsSorted, cap * 4));
    SPH_CUDA_CHECK(cudaMalloc(&s->rank, cap * 4));
    SPH_CUDA_CHECK(cudaMalloc(&s->cnt, cap * 4));
    SPH_CUDA_CHECK(cudaMalloc(&s->acc, cap * 8 * SYM_DOUBLES));
    cub::DeviceRadixSort::SortPairs(nullptr, s->cubTempBytes, s->keys, s->keysSorted, s->vals, s->valsSorted, (int)cap, 0, 64, ctx->stream);
    SPH_CUDA_CHECK(cudaMalloc(&s->cubTemp, s->cubTempBytes + 256));
    return SPHGPU_OK;
}

template <bool SOLID, bool FILTER>
static int launchSymmetricVariant(sphgpu_ctx* ctx, SymState* s) {
    const uint32_t n = ctx->nActive;
    cudaStream_t st = ctx->stream;
    k_sym_keys<<<(n + 255) / 256, 256, 0, st>>>(ctx->d, n, ctx->recDoubles, s->keys, s->vals);
    size_t bytes = s->cubTempBytes;
    SPH_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(s->cubTemp, bytes, s->keys, s->keysSorted, s->vals, s->valsSorted, (int)n, 0, 64, st));
    k_sym_rank<<<(n + 255) / 256, 256, 0, st>>>(ctx->d, n, s->capacity, s->valsSorted, s->rank, s->acc, s->cnt);
    k_pair_symmetric<SOLID, FILTER><<<(n + 127) / 128, 128, 0, st>>>(ctx->d, n, s->capacity, s->rank, s->acc, s->cnt);
    k_sym_finalize<SOLID><<<(n + 127) / 128, 128, 0, st>>>(ctx->d, n, ctx->n, s->capacity, s->acc, s->cnt);
    ctx->launches += 4;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

/// Variant 4 of the pair stage (sphgpu_set_variant): called after the large-particle kernels.
int launchPairSymmetric(sphgpu_ctx* ctx) {
    if (ctx->corrected || ctx->balsara || ctx->xsph || ctx->deltasph || ctx->stressAv) {
        setError("the symmetric formulation (variant 4) offers neither the correction tensor (like SymmetricSolver, SymmetricSolver.cpp:41-44) nor "
                 "the Balsara switch / XSph");
        return SPHGPU_E_INVALID;
    }
    int rc = ensureSymmetric(ctx);
    if (rc != SPHGPU_OK) {
        return rc;
    }
    SymState* s = static_cast<SymState*>(ctx->symmetric);
    if (!ctx->solid) {
        return launchSymmetricVariant<false, false>(ctx, s);
    }
    return ctx->filter ? launchSymmetricVariant<true, true>(ctx, s) : launchSymmetricVariant<true, false>(ctx, s);
}

} // namespace sph
