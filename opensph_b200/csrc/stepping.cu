// Device-resident time integration: elementwise kernels over the SoA state, no PCIe traffic.
// Replaces PredictorCorrector::makePredictions / makeCorrections and EulerExplicit::stepParticles
// (core/timestepping/TimeStepping.cpp:234-346, stepFirstOrder/stepSecondOrder + clampWithDerivative :81-228) and the
// four criteria of MultiCriterion (core/timestepping/TimeStepCriterion.cpp:117-419).
#include "sphgpu_internal.h"

namespace sph {

__device__ __forceinline__ void clampFirstOrder(const MaterialDev& m, bool hasDamage, double& rho, double& drho, double& u,
    double& du, double& D, double& dD) {
    if (rangeBounded(m.rho_min, m.rho_max)) {
        clampWithDerivative(rho, drho, m.rho_min, m.rho_max);
    }
    if (rangeBounded(m.u_min, m.u_max)) {
        clampWithDerivative(u, du, m.u_min, m.u_max);
    }
    if (hasDamage && rangeBounded(m.d_min, m.d_max)) {
        clampWithDerivative(D, dD, m.d_min, m.d_max);
    }
}

/// v dt + a dt^2/2 of the second-order predictor with the contraction spelled out: k_predict and k_correct_predict must
/// round alike (left to the compiler, the choice of which product joins the FMA differed between the two kernels).
__device__ __forceinline__ double secondOrderStep(double v, double a, double dt, double dt2) {
    return fma(v, dt, __dmul_rn(a, dt2));
}

// makePredictions (TimeStepping.cpp:286-300) + storage->swap(predictions) + zeroHighestDerivatives (:331-334).
// The derivative planes are not zeroed: integrate() overwrites every one of them.
template <bool SOLID>
__global__ void __launch_bounds__(256) k_predict(DevicePointers d, uint32_t n, double dtArg, bool hasDamage) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const double dt = d.dtDev ? *d.dtDev : dtArg;
    const MaterialDev& m = c_mats[d.u[U_MATID][i]];
    const double dt2 = 0.5 * dt * dt;
    const double ax = d.f[F_AX][i], ay = d.f[F_AY][i], az = d.f[F_AZ][i];
    const double vx = d.f[F_VX][i], vy = d.f[F_VY][i], vz = d.f[F_VZ][i], vh = d.f[F_VH][i];
    d.f[F_X][i] += secondOrderStep(vx, ax, dt, dt2);
    d.f[F_Y][i] += secondOrderStep(vy, ay, dt, dt2);
    d.f[F_Z][i] += secondOrderStep(vz, az, dt, dt2);
    d.f[F_H][i] += secondOrderStep(vh, 0., dt, dt2);
    d.f[F_VX][i] = vx + ax * dt;
    d.f[F_VY][i] = vy + ay * dt;
    d.f[F_VZ][i] = vz + az * dt;
    d.f[F_AXP][i] = ax;
    d.f[F_AYP][i] = ay;
    d.f[F_AZP][i] = az;
    const bool dmg = hasDamage && m.fracture != SPHGPU_FRACTURE_NONE;
    double rho = d.f[F_RHO][i], drho = d.f[F_DRHO][i], u = d.f[F_U][i], du = d.f[F_DU][i];
    double D = dmg ? d.f[F_D][i] : 0., dD = dmg ? d.f[F_DD][i] : 0.;
    rho += drho * dt;
    u += du * dt;
    D += dD * dt;
    clampFirstOrder(m, dmg, rho, drho, u, du, D, dD);
    d.f[F_RHO][i] = rho;
    d.f[F_U][i] = u;
    d.f[F_DRHOP][i] = drho;
    d.f[F_DUP][i] = du;
    if (dmg) {
        d.f[F_D][i] = D;
        d.f[F_DDP][i] = dD;
    }
    if (SOLID) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double ds = d.f[F_DS0 + k][i];
            d.f[F_S0 + k][i] += ds * dt;
            d.f[F_DSP0 + k][i] = ds;
        }
    }
}

// makeCorrections (TimeStepping.cpp:302-322): storage1 = *storage (p*), storage2 = predictions (c*).
template <bool SOLID>
__global__ void __launch_bounds__(256) k_correct(DevicePointers d, uint32_t n, double dtArg, bool hasDamage) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const double dt = d.dtDev ? *d.dtDev : dtArg;
    const MaterialDev& m = c_mats[d.u[U_MATID][i]];
    const double dt2 = 0.5 * dt * dt;
    const double a = 1. / 3., b = 0.5;
    const double ex = d.f[F_AXP][i] - d.f[F_AX][i], ey = d.f[F_AYP][i] - d.f[F_AY][i], ez = d.f[F_AZP][i] - d.f[F_AZ][i];
    d.f[F_X][i] -= a * ex * dt2;
    d.f[F_Y][i] -= a * ey * dt2;
    d.f[F_Z][i] -= a * ez * dt2;
    d.f[F_VX][i] -= b * ex * dt;
    d.f[F_VY][i] -= b * ey * dt;
    d.f[F_VZ][i] -= b * ez * dt;
    const bool dmg = hasDamage && m.fracture != SPHGPU_FRACTURE_NONE;
    double rho = d.f[F_RHO][i], drho = d.f[F_DRHO][i], u = d.f[F_U][i], du = d.f[F_DU][i];
    double D = dmg ? d.f[F_D][i] : 0., dD = dmg ? d.f[F_DD][i] : 0.;
    rho -= 0.5 * (d.f[F_DRHOP][i] - drho) * dt;
    u -= 0.5 * (d.f[F_DUP][i] - du) * dt;
    if (dmg) {
        D -= 0.5 * (d.f[F_DDP][i] - dD) * dt;
    }
    const double drho0 = drho, du0 = du, dD0 = dD;
    clampFirstOrder(m, dmg, rho, drho, u, du, D, dD);
    d.f[F_RHO][i] = rho;
    d.f[F_U][i] = u;
    if (drho != drho0) {
        d.f[F_DRHO][i] = drho;
    }
    if (du != du0) {
        d.f[F_DU][i] = du;
    }
    if (dmg) {
        d.f[F_D][i] = D;
        if (dD != dD0) {
            d.f[F_DD][i] = dD;
        }
    }
    if (SOLID) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            d.f[F_S0 + k][i] -= 0.5 * (d.f[F_DSP0 + k][i] - d.f[F_DS0 + k][i]) * dt;
        }
    }
}

// makeCorrections of the step that ends (time step StepStateDev::dtPrev) followed by makePredictions of the next one
// (StepStateDev::dt) in one pass: the bodies of k_correct and k_predict back to back on registers, so the state is read and
// written once instead of twice (512 instead of 824 bytes per solid particle). sphgpu_run_pc uses it between two steps when
// no time-step criterion reads what the corrector changes (everything but the derivative criterion), because then the
// criteria can be evaluated before the corrector and the next step is known when this kernel starts.
#ifndef INTEG_MIN_CTAS
#define INTEG_MIN_CTAS 4
#endif
template <bool SOLID>
__global__ void __launch_bounds__(256, INTEG_MIN_CTAS) k_correct_predict(DevicePointers d, uint32_t first, uint32_t n, bool hasDamage) {
    // particles [first, n): the whole owned range, or -- decomposed runs -- the send bands first and the interior while the
    // bands are already on their way to the neighbours (halo.cu)
    const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const MaterialDev& m = c_mats[d.u[U_MATID][i]];
    const bool dmg = hasDamage && m.fracture != SPHGPU_FRACTURE_NONE;
    const double dtc = d.stepState->dtPrev, dtp = d.stepState->dt; // corrector / predictor time step
    const double dtc2 = 0.5 * dtc * dtc, dtp2 = 0.5 * dtp * dtp;
    // Three groups, each loaded, advanced and stored before the next one starts (a store ends what the compiler may hoist,
    // so the groups bound the number of live registers): positions and velocities, first-order scalars, stress.
    {
        const double a = 1. / 3., b = 0.5;
        const double ax = d.f[F_AX][i], ay = d.f[F_AY][i], az = d.f[F_AZ][i];
        const double ex = d.f[F_AXP][i] - ax, ey = d.f[F_AYP][i] - ay, ez = d.f[F_AZP][i] - az;
        double x = d.f[F_X][i], y = d.f[F_Y][i], z = d.f[F_Z][i];
        double vx = d.f[F_VX][i], vy = d.f[F_VY][i], vz = d.f[F_VZ][i];
        const double vh = d.f[F_VH][i], h = d.f[F_H][i];
        x -= a * ex * dtc2; // corrector, as k_correct
        y -= a * ey * dtc2;
        z -= a * ez * dtc2;
        vx -= b * ex * dtc;
        vy -= b * ey * dtc;
        vz -= b * ez * dtc;
        x += secondOrderStep(vx, ax, dtp, dtp2); // predictor, as k_predict
        y += secondOrderStep(vy, ay, dtp, dtp2);
        z += secondOrderStep(vz, az, dtp, dtp2);
        d.f[F_X][i] = x;
        d.f[F_Y][i] = y;
        d.f[F_Z][i] = z;
        d.f[F_H][i] = h + secondOrderStep(vh, 0., dtp, dtp2);
        d.f[F_VX][i] = vx + ax * dtp;
        d.f[F_VY][i] = vy + ay * dtp;
        d.f[F_VZ][i] = vz + az * dtp;
        d.f[F_AXP][i] = ax;
        d.f[F_AYP][i] = ay;
        d.f[F_AZP][i] = az;
    }
    {
        double rho = d.f[F_RHO][i], drho = d.f[F_DRHO][i], u = d.f[F_U][i], du = d.f[F_DU][i];
        double D = dmg ? d.f[F_D][i] : 0., dD = dmg ? d.f[F_DD][i] : 0.;
        rho -= 0.5 * (d.f[F_DRHOP][i] - drho) * dtc;
        u -= 0.5 * (d.f[F_DUP][i] - du) * dtc;
        if (dmg) {
            D -= 0.5 * (d.f[F_DDP][i] - dD) * dtc;
        }
        clampFirstOrder(m, dmg, rho, drho, u, du, D, dD);
        rho += drho * dtp; // (the clamped derivatives of the corrector are the ones the predictor extrapolates with)
        u += du * dtp;
        D += dD * dtp;
        clampFirstOrder(m, dmg, rho, drho, u, du, D, dD);
        d.f[F_RHO][i] = rho;
        d.f[F_U][i] = u;
        d.f[F_DRHOP][i] = drho;
        d.f[F_DUP][i] = du;
        if (dmg) {
            d.f[F_D][i] = D;
            d.f[F_DDP][i] = dD;
        }
    }
    if (SOLID) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const double dS = d.f[F_DS0 + k][i];
            double S = d.f[F_S0 + k][i];
            S -= 0.5 * (d.f[F_DSP0 + k][i] - dS) * dtc;
            d.f[F_S0 + k][i] = S + dS * dtp;
            d.f[F_DSP0 + k][i] = dS;
        }
    }
}

// EulerExplicit::stepParticles after solver.integrate (TimeStepping.cpp:243-264).
template <bool SOLID>
__global__ void __launch_bounds__(256) k_euler(DevicePointers d, uint32_t n, double dt, bool hasDamage) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const MaterialDev& m = c_mats[d.u[U_MATID][i]];
    const double vx = d.f[F_VX][i] + d.f[F_AX][i] * dt;
    const double vy = d.f[F_VY][i] + d.f[F_AY][i] * dt;
    const double vz = d.f[F_VZ][i] + d.f[F_AZ][i] * dt;
    d.f[F_VX][i] = vx;
    d.f[F_VY][i] = vy;
    d.f[F_VZ][i] = vz;
    d.f[F_X][i] += vx * dt;
    d.f[F_Y][i] += vy * dt;
    d.f[F_Z][i] += vz * dt;
    d.f[F_H][i] += d.f[F_VH][i] * dt;
    const bool dmg = hasDamage && m.fracture != SPHGPU_FRACTURE_NONE;
    double rho = d.f[F_RHO][i], drho = d.f[F_DRHO][i], u = d.f[F_U][i], du = d.f[F_DU][i];
    double D = dmg ? d.f[F_D][i] : 0., dD = dmg ? d.f[F_DD][i] : 0.;
    rho += drho * dt;
    u += du * dt;
    D += dD * dt;
    const double drho0 = drho, du0 = du, dD0 = dD;
    clampFirstOrder(m, dmg, rho, drho, u, du, D, dD);
    d.f[F_RHO][i] = rho;
    d.f[F_U][i] = u;
    if (drho != drho0) {
        d.f[F_DRHO][i] = drho;
    }
    if (du != du0) {
        d.f[F_DU][i] = du;
    }
    if (dmg) {
        d.f[F_D][i] = D;
        if (dD != dD0) {
            d.f[F_DD][i] = dD;
        }
    }
    if (SOLID) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            d.f[F_S0 + k][i] += d.f[F_DS0 + k][i] * dt;
        }
    }
}

__device__ __forceinline__ double warpMinD(double v) {
    for (int o = 16; o > 0; o >>= 1) {
        v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;
}

// Courant / Derivative / Acceleration / Divergence criteria in one pass; per-criterion minima are combined with
// atomicMin on the bit pattern (all candidates are positive doubles, whose bit patterns order like the values).
template <bool SOLID>
__global__ void __launch_bounds__(256) k_criteria(DevicePointers d, uint32_t n, bool hasDamage) {
    double mins[4] = { INFTY_REF, INFTY_REF, INFTY_REF, INFTY_REF };
    const uint32_t crit = c_prm.criteria;
#pragma unroll 4 // (no stores in the loop: the loads of four particles are in flight together)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const MaterialDev& m = c_mats[d.u[U_MATID][i]];
        const double h = d.f[F_H][i];
        if (crit & SPHGPU_CRIT_COURANT) {
            const double cs = d.f[F_CS][i];
            if (cs > 0.) {
                mins[0] = fmin(mins[0], c_prm.courant * h / cs);
            }
        }
        if (crit & SPHGPU_CRIT_DERIVATIVES) {
            const double f = c_prm.derivative_factor;
            double s = derivativeStep(fabs(d.f[F_RHO][i]), fabs(d.f[F_DRHO][i]), m.rho_small, f);
            s = fmin(s, derivativeStep(fabs(d.f[F_U][i]), fabs(d.f[F_DU][i]), m.u_small, f));
            if (hasDamage && m.fracture != SPHGPU_FRACTURE_NONE) {
                s = fmin(s, derivativeStep(fabs(d.f[F_D][i]), fabs(d.f[F_DD][i]), m.d_small, f));
            }
            if (SOLID) {
                double S[5], dS[5];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    S[k] = d.f[F_S0 + k][i];
                    dS[k] = d.f[F_DS0 + k][i];
                    s = fmin(s, derivativeStep(fabs(S[k]), fabs(dS[k]), m.s_small, f));
                }
                s = fmin(s, derivativeStep(fabs(-S[0] - S[1]), fabs(-dS[0] - dS[1]), m.s_small, f));
            }
            mins[1] = fmin(mins[1], s);
        }
        if (crit & SPHGPU_CRIT_ACCELERATION) {
            const double ax = d.f[F_AX][i], ay = d.f[F_AY][i], az = d.f[F_AZ][i];
            const double dvNorm = ax * ax + ay * ay + az * az;
            if (dvNorm > EPS_REF) {
                mins[2] = fmin(mins[2], c_prm.derivative_factor * sqrt(sqrt(h * h / dvNorm)));
            }
        }
        if (crit & SPHGPU_CRIT_DIVERGENCE) {
            const double dv = fabs(d.f[F_DIVV][i]);
            if (dv > EPS_REF) {
                mins[3] = fmin(mins[3], c_prm.divergence_factor / dv);
            }
        }
    }
    __shared__ double sm[8][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < 4; ++k) {
        mins[k] = warpMinD(mins[k]);
    }
    if (lane == 0) {
        for (int k = 0; k < 4; ++k) {
            sm[warp][k] = mins[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double r = sm[0][threadIdx.x];
        for (int w = 1; w < 8; ++w) {
            r = fmin(r, sm[w][threadIdx.x]);
        }
        if (r < 0.) {
            r = 0.;
        }
        atomicMin(&d.tsd->minBits[threadIdx.x], (unsigned long long)__double_as_longlong(r));
    }
}

// FP64 FMA throughput probe (the secondary roofline of the pair kernel; SURVEY 8d asks for a measured DFMA peak).
__global__ void __launch_bounds__(256) k_dfma_probe(double* out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double b = 1.0000001, c = 1.e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) {
        out[0] = r; // never true; keeps the loop alive
    }
}

#define SPH_DISPATCH_SOLID(KERNEL, ...)                                                                               \
    do {                                                                                                              \
        if (ctx->solid) {                                                                                             \
            KERNEL<true><<<blocks, 256, 0, ctx->stream>>>(__VA_ARGS__);                                               \
        } else {                                                                                                      \
            KERNEL<false><<<blocks, 256, 0, ctx->stream>>>(__VA_ARGS__);                                              \
        }                                                                                                             \
        ctx->launches += 1;                                                                                           \
    } while (0)

int measureFp64Peak(sphgpu_ctx* ctx, double* fmaPerSecond) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int blocks = sms * 8, threads = 256, iters = 1 << 16;
    cudaStream_t st = ctx->stream;
    k_dfma_probe<<<blocks, threads, 0, st>>>(ctx->d.boundsPartial, 1024, 0.5); // warm-up
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[6], st));
    k_dfma_probe<<<blocks, threads, 0, st>>>(ctx->d.boundsPartial, iters, 0.5);
    SPH_CUDA_CHECK(cudaEventRecord(ctx->ev[7], st));
    SPH_CUDA_CHECK(cudaStreamSynchronize(st));
    float ms = 0.f;
    SPH_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]));
    *fmaPerSecond = (double)blocks * threads * 8. * iters / (ms * 1.e-3);
    return SPHGPU_OK;
}

// MultiCriterion::compute (TimeStepCriterion.cpp:389-419) on the device, the same operations as finishTimestep() in
// api.cu: lets sphgpu_run_pc queue several steps without a host round trip for the time step.
__global__ void k_finish_timestep(DevicePointers d, double maxDt, double maxChange, uint32_t criteria, StepRecordDev* history,
    uint32_t index) {
    if (blockIdx.x != 0 || threadIdx.x != 0) {
        return;
    }
    const uint32_t bits[4] = { SPHGPU_CRIT_COURANT, SPHGPU_CRIT_DERIVATIVES, SPHGPU_CRIT_ACCELERATION, SPHGPU_CRIT_DIVERGENCE };
    const uint32_t ids[4] = { SPHGPU_CRITID_CFL_CONDITION, SPHGPU_CRITID_DERIVATIVE, SPHGPU_CRITID_ACCELERATION,
        SPHGPU_CRITID_DIVERGENCE };
    double minStep = INFTY_REF;
    uint32_t minId = SPHGPU_CRITID_INITIAL_VALUE;
    for (int k = 0; k < 4; ++k) {
        if (!(criteria & bits[k])) {
            continue;
        }
        double step = __longlong_as_double((long long)d.tsd->minBits[k]);
        uint32_t id = ids[k];
        if (step > maxDt) {
            step = maxDt;
            id = SPHGPU_CRITID_MAXIMAL_VALUE;
        }
        if (step < minStep) {
            minStep = step;
            minId = id;
        }
    }
    StepStateDev& st = *d.stepState;
    if (maxChange < 1.e300) {
        if (!st.lastDtInit) {
            st.lastDt = minStep;
            st.lastDtInit = 1u;
        }
        const double maxStep = __dmul_rn(st.lastDt, 1. + maxChange);
        if (minStep > maxStep) {
            minStep = maxStep;
            minId = SPHGPU_CRITID_MAX_CHANGE;
        }
        st.lastDt = minStep;
    }
    st.dtPrev = st.dt;
    st.dt = minStep;
    history[index].dt = minStep;
    history[index].criterion = minId;
    history[index].pad = 0u;
}

int launchFinishTimestep(sphgpu_ctx* ctx, double maxDt, StepRecordDev* history, uint32_t index) {
    k_finish_timestep<<<1, 32, 0, ctx->stream>>>(ctx->d, maxDt, ctx->maxChange, ctx->prm.criteria, history, index);
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchPredict(sphgpu_ctx* ctx, double dt) {
    {
        const int rcConst = ensureConstants(ctx);
        if (rcConst != SPHGPU_OK) {
            return rcConst;
        }
    }
    const uint32_t n = ctx->n;
    if (n == 0) {
        return SPHGPU_OK;
    }
    const uint32_t blocks = (n + 255) / 256;
    SPH_DISPATCH_SOLID(k_predict, ctx->d, n, dt, ctx->hasDamage);
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchCorrect(sphgpu_ctx* ctx, double dt) {
    {
        const int rcConst = ensureConstants(ctx);
        if (rcConst != SPHGPU_OK) {
            return rcConst;
        }
    }
    const uint32_t n = ctx->n;
    if (n == 0) {
        return SPHGPU_OK;
    }
    const uint32_t blocks = (n + 255) / 256;
    SPH_DISPATCH_SOLID(k_correct, ctx->d, n, dt, ctx->hasDamage);
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchCorrectPredictRange(sphgpu_ctx* ctx, uint32_t first, uint32_t end) {
    {
        const int rcConst = ensureConstants(ctx);
        if (rcConst != SPHGPU_OK) {
            return rcConst;
        }
    }
    if (end <= first) {
        return SPHGPU_OK;
    }
    const uint32_t blocks = (end - first + 255) / 256;
    SPH_DISPATCH_SOLID(k_correct_predict, ctx->d, first, end, ctx->hasDamage);
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchCorrectPredict(sphgpu_ctx* ctx) {
    return launchCorrectPredictRange(ctx, 0u, ctx->n);
}

int launchEuler(sphgpu_ctx* ctx, double dt) {
    {
        const int rcConst = ensureConstants(ctx);
        if (rcConst != SPHGPU_OK) {
            return rcConst;
        }
    }
    const uint32_t n = ctx->n;
    if (n == 0) {
        return SPHGPU_OK;
    }
    const uint32_t blocks = (n + 255) / 256;
    SPH_DISPATCH_SOLID(k_euler, ctx->d, n, dt, ctx->hasDamage);
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

// ---- FrozenParticles::finalize (core/sph/boundary/Boundary.cpp:221-258) ------------------------------------------------
// Particles outside the spherical domain are projected onto its surface (SphericalDomain::project, Domain.cpp:66-85:
// centre + unit vector * (1 - EPS) * radius); particles within `freezeRadius` smoothing lengths of the surface and particles
// of the frozen bodies get their highest derivatives zeroed. The damage derivative belongs to material->finalize, which the
// reference runs after the boundary condition: it stays.
template <bool SOLID>
__global__ void __launch_bounds__(256) k_frozen(DevicePointers d, uint32_t n, sphgpu_frozen f) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) {
        return;
    }
    const uint32_t flag = d.u[U_FLAG][i];
    bool frozen = flag < 64u && ((f.flag_mask >> flag) & 1ull) != 0ull;
    if (f.has_domain) {
        double x = d.f[F_X][i] - f.center[0], y = d.f[F_Y][i] - f.center[1], z = d.f[F_Z][i] - f.center[2];
        double len2 = x * x + y * y + z * z;
        if (!(len2 <= f.radius * f.radius)) { // SphericalDomain::isInsideImpl (Domain.h:134-136)
            const double len = sqrt(len2);
            const double s = 1. - (double)1.e-12f; // getNormalized(v - centre) * (1 - EPS) * radius + centre, in this order
            x = x / len * s * f.radius + f.center[0];
            y = y / len * s * f.radius + f.center[1];
            z = z / len * s * f.radius + f.center[2];
            d.f[F_X][i] = x;
            d.f[F_Y][i] = y;
            d.f[F_Z][i] = z;
            x -= f.center[0];
            y -= f.center[1];
            z -= f.center[2];
            len2 = x * x + y * y + z * z;
        }
        // getDistanceToBoundary (Domain.cpp:58-64): radius - |r - centre|
        frozen = frozen || (f.radius - sqrt(len2) < f.freeze_radius * d.f[F_H][i]);
    }
    if (!frozen) {
        return;
    }
    d.f[F_AX][i] = d.f[F_AY][i] = d.f[F_AZ][i] = 0.;
    d.f[F_DRHO][i] = 0.;
    d.f[F_DU][i] = 0.;
    if (SOLID) {
        for (int k = 0; k < 5; ++k) {
            d.f[F_DS0 + k][i] = 0.;
        }
    }
}

int launchFrozen(sphgpu_ctx* ctx) {
    const uint32_t n = ctx->n;
    if (n == 0) {
        return SPHGPU_OK;
    }
    if (ctx->solid) {
        k_frozen<true><<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d, n, ctx->frozen);
    } else {
        k_frozen<false><<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d, n, ctx->frozen);
    }
    ctx->launches += 1;
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

int launchCriteria(sphgpu_ctx* ctx) {
    {
        const int rcConst = ensureConstants(ctx);
        if (rcConst != SPHGPU_OK) {
            return rcConst;
        }
    }
    const uint32_t n = ctx->n;
    TimestepDev init;
    const double inf = INFTY_REF;
    for (int k = 0; k < 4; ++k) {
        memcpy(&init.minBits[k], &inf, 8);
    }
    SPH_CUDA_CHECK(cudaMemcpyAsync(ctx->d.tsd, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    if (n == 0) {
        return SPHGPU_OK;
    }
    const uint32_t blocks = min((n + 255) / 256, (uint32_t)BOUNDS_BLOCKS * 2);
    SPH_DISPATCH_SOLID(k_criteria, ctx->d, n, ctx->hasDamage);
    SPH_CUDA_CHECK(cudaGetLastError());
    return SPHGPU_OK;
}

} // namespace sph
