// Per-particle and per-pair arithmetic of the SPH hot path, shared by every kernel variant.
// All functions are __host__ __device__ so that tests/csrc/host_math_check.cu can run the very same code on the
// CPU against the oracle (a compile-time check of the formulas, NOT a product path).
//
// Reference formulas (paths relative to the reference root):
//   kernel gradient      core/sph/kernel/Kernel.h:32-36,129-144,640-643
//   VelocityDivergence   core/sph/equations/DerivativeHelpers.h:292-304,355-371
//   VelocityGradient     core/sph/equations/DerivativeHelpers.h:100-146,382-398
//   CorrectionTensor     core/sph/equations/Derivative.cpp:36-74
//   PressureGradient     core/sph/equations/EquationTerm.cpp:12-24,41-67
//   StressDivergence     core/sph/equations/EquationTerm.cpp:113-140
//   StandardAV           core/sph/equations/av/Standard.h:63-83
//   finalizers           core/sph/equations/EquationTerm.cpp:90-99,177-203,289-316,366-434
//   Tillotson / ideal gas core/physics/Eos.cpp:42-45,198-238
//   von Mises            core/physics/Rheology.cpp:36-83
//   Grady-Kipp           core/physics/Damage.cpp:128-170, core/objects/geometry/SymmetricTensor.h:377-399
#pragma once
#include "../../include/sphgpu.h"
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SPH_HD __host__ __device__ __forceinline__
#else
#define SPH_HD inline
#endif

namespace sph {

constexpr double EPS_REF = (double)1.e-12f;  // math/MathUtils.h:26-30 (float literal)
constexpr double LARGE_REF = (double)1.e20f; // math/MathUtils.h:33
constexpr double INFTY_REF = 1.7976931348623157e308;

/// Per-material constants in the form the kernels consume.
struct MaterialDev {
    uint32_t eos, yielding, fracture, pad;
    double til_u0, til_uiv, til_ucv, til_a, til_b, rho0, til_A, til_B, til_alpha, til_beta, gamma;
    double shear_modulus, elasticity_limit, melt_energy, young_modulus;
    double rho_min, rho_max, u_min, u_max, d_min, d_max;
    double rho_small, u_small, d_small, s_small;
};

constexpr int MAX_MATERIALS = 32;

/// Run-level constants.
struct ParamsDev {
    uint32_t forces, flags, continuity_mode, lut_entries;
    double kernel_radius, radius_sqr, q_sqr_to_idx;
    double av_alpha, av_beta;
    double av_minus_half_alpha, av_eps_over_radius_sqr; // -alpha / 2 and 0.01 / R^2 (masked pair body)
    double h_min, h_max, neigh_enforcing, neigh_lower, neigh_upper;
    uint32_t criteria, n_materials;
    double courant, derivative_factor, divergence_factor;
    double xsph_eps; // SPH_XSPH_EPSILON (SPHGPU_FLAG_XSPH)
    double deltasph_half_delta, deltasph_half_alpha; // SPH_DENSITY_DIFFUSION_DELTA / 2, SPH_VELOCITY_DIFFUSION_ALPHA / 2
    double stress_av_exponent, stress_av_factor;     // SPH_AV_STRESS_EXPONENT, SPH_AV_STRESS_FACTOR (SPHGPU_FLAG_STRESS_AV)
    uint32_t stress_av_int_exponent, pad0;           // the exponent if it is 1, 2, 3 or 4 (stressAvIntExponent), else 0
};

/// Small integer exponents of the artificial stress' weighting function are evaluated by multiplication.
SPH_HD uint32_t stressAvIntExponent(double n) {
    return n == 1. ? 1u : n == 2. ? 2u : n == 3. ? 3u : n == 4. ? 4u : 0u;
}

SPH_HD double sqr(double x) {
    return x * x;
}

/// c ? a : b as a single predicated move. Written in PTX so that the compiler cannot turn the select (and the
/// arithmetic feeding it) into a divergent branch; keeps the pair body straight-line code.
SPH_HD double selectD(bool c, double a, double b) {
#ifdef __CUDA_ARCH__
    double r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b), "r"((unsigned)c));
    return r;
#else
    return c ? a : b;
#endif
}

/// Reciprocal of a positive normal double: hardware seed (MUFU.RCP64H, >= 20 bits) + one cubically convergent step
/// x (1 + e + e^2), e = 1 - a x (3 FMA; error e^3 <= 2^-60 plus rounding). Within 1-2 ulp, and -- unlike the compiler's
/// IEEE division -- free of the denormal/overflow fix-up branches, which would keep the scheduler from overlapping two
/// pair bodies. On the host (formula tests) it is a plain division.
SPH_HD double fastRcp(double a) {
#ifdef __CUDA_ARCH__
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    const double e = fma(-a, x, 1.);
    const double t = fma(e, e, e);
    return fma(x, t, x);
#else
    return 1. / a;
#endif
}

// ---- equation of state ---------------------------------------------------------------------------------

SPH_HD void eosTillotson(const MaterialDev& m, double rho, double u, double& pOut, double& csOut) {
    const double eta = rho / m.rho0;
    const double mu = eta - 1.;
    const double denom = u / (m.til_u0 * eta * eta) + 1.;
    const double denom2 = sqr(denom);
    const double rho2 = rho * rho;
    // compressed phase
    const double pc = (m.til_a + m.til_b / denom) * rho * u + m.til_A * mu + m.til_B * mu * mu;
    double dpdu = m.til_a * rho + m.til_b * rho / denom2;
    double dpdrho = m.til_a * u + m.til_b * u * (3. * denom - 2.) / denom2 + m.til_A / m.rho0 + 2. * m.til_B * mu / m.rho0;
    const double csc = dpdrho + dpdu * pc / rho2;
    // expanded phase
    const double rhoExp = m.rho0 / rho - 1.;
    const double betaExp = exp(-fmin(m.til_beta * rhoExp, 70.));
    const double alphaExp = exp(-fmin(m.til_alpha * sqr(rhoExp), 70.));
    const double pe = m.til_a * rho * u + (m.til_b * rho * u / denom + m.til_A * mu * betaExp) * alphaExp;
    dpdu = m.til_a * rho + alphaExp * m.til_b * rho / denom2;
    dpdrho = m.til_a * u + alphaExp * (m.til_b * u * (3. * denom - 2.) / denom2) +
             alphaExp * (m.til_b * u * rho / denom) * m.rho0 * (2. * m.til_alpha * rhoExp) / rho2 +
             alphaExp * m.til_A * betaExp * (1. / m.rho0 + m.rho0 * mu / rho2 * (2. * m.til_alpha * rhoExp + m.til_beta));
    double cse = dpdrho + dpdu * pe / rho2;
    cse = fmax(cse, 0.);
    double p = pc, cs = csc;
    if (rho <= m.rho0 && u > m.til_ucv) {
        p = pe;
        cs = cse;
    } else if (rho <= m.rho0 && u > m.til_uiv && u <= m.til_ucv) {
        p = ((u - m.til_uiv) * pe + (m.til_ucv - u) * pc) / (m.til_ucv - m.til_uiv);
        cs = ((u - m.til_uiv) * cse + (m.til_ucv - u) * csc) / (m.til_ucv - m.til_uiv);
    }
    cs = fmax(cs, 0.25 * m.til_A / m.rho0);
    pOut = p;
    csOut = sqrt(cs);
}

SPH_HD void evalEos(const MaterialDev& m, double rho, double u, double& p, double& cs) {
    if (m.eos == SPHGPU_EOS_TILLOTSON) {
        eosTillotson(m, rho, u, p, cs);
    } else if (m.eos == SPHGPU_EOS_IDEAL_GAS) {
        p = (m.gamma - 1.) * u * rho;
        cs = sqrt(m.gamma * p / rho);
    }
}

// ---- rheology + damage ---------------------------------------------------------------------------------

/// VonMisesRheology::initialize for one particle. S = {xx,yy,xy,xz,yz} is scaled in place; returns the reducing factor.
SPH_HD double vonMises(const MaterialDev& m, double u, double D, bool hasDamage, double& p, double S[5]) {
    const double d = hasDamage ? D * D * D : 0.;
    if (p < 0.) {
        p = (1. - d) * p;
    }
    const double unorm = u / m.melt_energy;
    double Y = unorm < 1.e-5 ? m.elasticity_limit : m.elasticity_limit * fmax(1. - unorm, 0.);
    Y = (1. - d) * Y;
    if (Y < EPS_REF) {
        S[0] = S[1] = S[2] = S[3] = S[4] = 0.;
        return 0.;
    }
    const double szz = -S[0] - S[1];
    const double ddot = (S[0] * S[0] + S[1] * S[1] + szz * szz) + 2. * (S[2] * S[2] + S[3] * S[3] + S[4] * S[4]);
    const double J2 = 0.5 * ddot + 1.e-15;
    const double red = fmin(Y / sqrt(3. * J2), 1.);
    for (int k = 0; k < 5; ++k) {
        S[k] = S[k] * red;
    }
    return red;
}

/// sqrtApprox (math/MathUtils.h:40-62): float bit trick + one Newton iteration. Only scales the eigenvalue problem.
SPH_HD double sqrtApproxRef(double f) {
    if (f == 0.) {
        return 0.;
    }
    float y = (float)f;
    const float x2 = y * 0.5f;
#ifdef __CUDA_ARCH__
    int i = __float_as_int(y);
    i = 0x5f3759df - (i >> 1);
    y = __int_as_float(i);
#else
    int i;
    memcpy(&i, &y, sizeof(float));
    i = 0x5f3759df - (i >> 1);
    memcpy(&y, &i, sizeof(float));
#endif
    const float r = y * (1.5f - (x2 * y * y));
    return (double)(1.f / r);
}

/// Largest eigenvalue of the symmetric tensor t = {xx,yy,zz,xy,xz,yz} as findEigenvalues computes it.
SPH_HD double maxEigenvalue(const double t[6]) {
    const double v0 = fmax(t[0], t[3]), v1 = fmax(t[1], t[4]), v2 = fmax(t[2], t[5]);
    const double n = sqrtApproxRef(v0 * v0 + v1 * v1 + v2 * v2);
    if (n < 1.e-12) {
        return 0.;
    }
    const double inv1 = t[0] + t[1] + t[2];
    const double inv2 = (t[3] * t[3] + t[4] * t[4] + t[5] * t[5]) - (t[1] * t[2] + t[2] * t[0] + t[0] * t[1]);
    const double inv3 = t[0] * t[1] * t[2] + 2. * t[3] * t[4] * t[5] - (t[3] * t[3] * t[2] + t[4] * t[4] * t[1] + t[5] * t[5] * t[0]);
    const double p = -inv1 / n;
    const double q = -inv2 / sqr(n);
    const double r = -inv3 / (n * n * n);
    const double a = q - p * p / 3.;
    const double b = (2. * p * p * p - 9. * p * q + 27. * r) / 27.;
    const double aCub = a * a * a / 27.;
    if (0.25 * b * b + aCub >= 0.) {
        return 0.;
    }
    const double t1 = 2. * sqrt(-a / 3.);
    const double phi = acos(-0.5 * b / sqrt(-aCub));
    const double PI = 3.14159265358979323846;
    const double s0 = (t1 * cos(phi / 3.) - p / 3.) * n;
    const double s1 = (t1 * cos((phi + 2 * PI) / 3.) - p / 3.) * n;
    const double s2 = (t1 * cos((phi + 4 * PI) / 3.) - p / 3.) * n;
    return fmax(fmax(s0, s1), s2);
}

/// ScalarGradyKippModel::integrate for one particle; returns dD/dt (0 when the flaw is not activated).
SPH_HD double damageRate(const MaterialDev& m, double p, const double S[5], double D, double epsMin, double mZero,
    double growth, uint32_t nFlaws) {
    if (D >= m.d_max) {
        return LARGE_REF;
    }
    const double sigma[6] = { S[0] - p, S[1] - p, (-S[0] - S[1]) - p, S[2], S[3], S[4] };
    const double sigMax = maxEigenvalue(sigma);
    const double youngRed = fmax((1. - D * D * D) * m.young_modulus, 1.e-20);
    const double strain = sigMax / youngRed;
    const double ratio = strain / epsMin;
    if (ratio <= 1.) {
        return 0.;
    }
    return growth * cbrt(fmin(pow(ratio, mZero), (double)nFlaws));
}

/// BalsaraSwitch::Derivative::factor (core/sph/equations/av/Balsara.h:76-80), from the divergence and rotation of the
/// PREVIOUS evaluation (the Storage values at the time of the call) and the current sound speed.
SPH_HD double balsaraFactor(double divv, double rx, double ry, double rz, double cs, double h) {
    const double dv = fabs(divv);
    const double rv = sqrt(rx * rx + ry * ry + rz * rz);
    return dv / (dv + rv + 1.e-4 * cs / h);
}

/// One Jacobi rotation annihilating the off-diagonal element a_pq of a symmetric 3x3 matrix; r is the third index.
/// vp, vq: the columns p and q of the accumulated rotation.
SPH_HD void jacobiRotate(double& app, double& aqq, double& apq, double& arp, double& arq, double vp[3], double vq[3]) {
    if (apq == 0.) {
        return;
    }
    const double theta = (aqq - app) / (2. * apq);
    const double t = (theta >= 0. ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1.));
    const double c = 1. / sqrt(t * t + 1.), sn = t * c;
    app -= t * apq;
    aqq += t * apq;
    apq = 0.;
    const double rp = arp, rq = arq;
    arp = c * rp - sn * rq;
    arq = sn * rp + c * rq;
    for (int k = 0; k < 3; ++k) {
        const double a = vp[k], b = vq[k];
        vp[k] = c * a - sn * b;
        vq[k] = sn * a + c * b;
    }
}

/// StressAV::initialize for one particle (core/sph/equations/av/Stress.cpp:91-109): as = -V max(Lambda, 0) V^T for the
/// eigen-decomposition sigma = V Lambda V^T of the total stress sigma = S - p I, i.e. minus its positive (tensile) part. Being
/// a function of the tensor, the result does not depend on the eigen-solver beyond rounding: the reference runs JAMA's
/// tred2 / tql2 (SymmetricTensor.cpp:110-272), this is cyclic Jacobi (quadratically convergent; the sweeps stop when the
/// off-diagonal part is exactly zero, 12 at most). sigma, as: {xx,yy,zz,xy,xz,yz}.
SPH_HD void avStressOf(const double sigma[6], double as[6]) {
    double a00 = sigma[0], a11 = sigma[1], a22 = sigma[2], a01 = sigma[3], a02 = sigma[4], a12 = sigma[5];
    double v0[3] = { 1., 0., 0. }, v1[3] = { 0., 1., 0. }, v2[3] = { 0., 0., 1. }; // columns of V
    for (int sweep = 0; sweep < 12; ++sweep) {
        if (fabs(a01) + fabs(a02) + fabs(a12) == 0.) {
            break;
        }
        jacobiRotate(a00, a11, a01, a02, a12, v0, v1); // (p, q, r) = (0, 1, 2)
        jacobiRotate(a00, a22, a02, a01, a12, v0, v2); // (0, 2, 1)
        jacobiRotate(a11, a22, a12, a01, a02, v1, v2); // (1, 2, 0)
    }
    const double l0 = fmax(a00, 0.), l1 = fmax(a11, 0.), l2 = fmax(a22, 0.);
    as[0] = -(l0 * v0[0] * v0[0] + l1 * v1[0] * v1[0] + l2 * v2[0] * v2[0]);
    as[1] = -(l0 * v0[1] * v0[1] + l1 * v1[1] * v1[1] + l2 * v2[1] * v2[1]);
    as[2] = -(l0 * v0[2] * v0[2] + l1 * v1[2] * v1[2] + l2 * v2[2] * v2[2]);
    as[3] = -(l0 * v0[0] * v0[1] + l1 * v1[0] * v1[1] + l2 * v2[0] * v2[1]);
    as[4] = -(l0 * v0[0] * v0[2] + l1 * v1[0] * v1[2] + l2 * v2[0] * v2[2]);
    as[5] = -(l0 * v0[1] * v0[2] + l1 * v1[1] * v1[2] + l2 * v2[1] * v2[2]);
}

/// The weighting function (W_ij / W_0)^n of the artificial stress (Stress.cpp:46-47): small integer exponents (the default is 4)
/// by multiplication, anything else through pow.
SPH_HD double stressAvWeight(double x, double n, uint32_t intN) {
    const double x2 = x * x;
    if (intN == 4u) {
        return x2 * x2;
    }
    if (intN == 0u) {
        return pow(x, n);
    }
    return intN == 2u ? x2 : intN == 3u ? x2 * x : x;
}

// ---- pair interaction ------------------------------------------------------------------------------------

/// What one particle contributes as a neighbour. P = p/rho^2, Sr = S/rho^2, vol = m/rho, grp = body flag or -1 if the
/// particle is fully damaged (reduce == 0), i.e. excluded by the SUM_ONLY_UNDAMAGED filter. The sorted records do not
/// carry m: the loaders set m = vol * rho (one rounding, 1e-16 relative).
struct Particle {
    double x, y, z, h;
    double vx, vy, vz;
    double m, rho, P, cs, vol;
    double Sr[5];
    double bal; // Balsara factor |div v| / (|div v| + |rot v| + 1e-4 cs / h) (Balsara.h:76-80); unused without the switch
    double gr[3]; // DELTASPH_DENSITY_GRADIENT of the previous evaluation (DeltaSph.h:71-74); delta-SPH terms only
    double as[6]; // AV_STRESS / rho^2 {xx,yy,zz,xy,xz,yz} (Stress.cpp:68-79); artificial stress only
    double wpInv; // 1 / INTERPARTICLE_SPACING_KERNEL (Stress.cpp:113-121); artificial stress only, targets only
    int grp;
};

/// Per-target running sums. T = sum m_j (v_j - v_i) (x) gradW (full 3x3, row-major), Cm = sum V_j (r_j - r_i) (x) gradW.
struct Accum {
    double ax, ay, az, du, divv;
    double T[9];
    double Cm[6];
    double F[3]; // sum of the stress-weighted kernel gradients m_j gradW (pairSums): the target's own Sr is applied once
    double rot[3]; // sum m_j gradW x (v_j - v_i)  (VelocityRotation, DerivativeHelpers.h:374-395; Balsara switch only)
    double xs[3];  // sum m_j eps (v_j - v_i) W_ij / rhobar  (XSph::Derivative, XSph.h:55-63; XSph term only)
    double dg[3];  // sum V_j (rho_j - rho_i) gradW  (DeltaSph::RenormalizedDensityGradient, DeltaSph.h:37-44; delta-SPH only)
    double ddrho;  // sum V_j delta hbar cbar psi_ij . gradW  (DeltaSph::DensityDiffusion, DeltaSph.h:81-93; delta-SPH only)
    uint32_t cnt;
};

SPH_HD void accumZero(Accum& a) {
    a.ax = a.ay = a.az = a.du = a.divv = 0.;
    for (int k = 0; k < 9; ++k) {
        a.T[k] = 0.;
    }
    for (int k = 0; k < 6; ++k) {
        a.Cm[k] = 0.;
    }
    a.F[0] = a.F[1] = a.F[2] = 0.;
    a.rot[0] = a.rot[1] = a.rot[2] = 0.;
    a.xs[0] = a.xs[1] = a.xs[2] = 0.;
    a.dg[0] = a.dg[1] = a.dg[2] = 0.;
    a.ddrho = 0.;
    a.cnt = 0;
}

/// Exact neighbour predicate of AsymmetricSolver::loop (AsymmetricSolver.cpp:186-191): d^2 < (R * hbar)^2, evaluated
/// without FMA contraction and in the reference's operation order so that the neighbour SETS are bit-identical.
/// sq receives {dx^2, dy^2, dz^2, (R hbar)^2}, by-products the masked pair body reuses.
SPH_HD bool isNeighbour(double dx, double dy, double dz, double hi, double hj, double R, double& d2, double& hbar, double sq[4]) {
#ifdef __CUDA_ARCH__
    sq[0] = __dmul_rn(dx, dx);
    sq[1] = __dmul_rn(dy, dy);
    sq[2] = __dmul_rn(dz, dz);
    d2 = __dadd_rn(__dadd_rn(sq[0], sq[1]), sq[2]);
    hbar = __dmul_rn(0.5, __dadd_rn(hi, hj));
    const double rh = __dmul_rn(R, hbar);
    sq[3] = __dmul_rn(rh, rh);
    return d2 < sq[3];
#else
    sq[0] = dx * dx;
    sq[1] = dy * dy;
    sq[2] = dz * dz;
    d2 = sq[0] + sq[1] + sq[2];
    hbar = 0.5 * (hi + hj);
    const double rh = R * hbar;
    sq[3] = rh * rh;
    return d2 < sq[3];
#endif
}

SPH_HD bool isNeighbour(double dx, double dy, double dz, double hi, double hj, double R, double& d2, double& hbar) {
    double sq[4];
    return isNeighbour(dx, dy, dz, hi, hj, R, d2, hbar, sq);
}

/// One entry of the interleaved gradient table: {G[k], G[k + 1] - G[k]}, so that the interpolation is one 16-byte load
/// and one FMA. Entry lut_entries is the zero guard {0, 0}.
struct LutPair {
    double g, dg;
};

/// Fills out[0 .. entries] from the reference table lutGrad[0 .. entries] (Kernel.h:85-101); G[entries + 1] = 0.
inline void buildLutPairs(const double* lutGrad, uint32_t entries, LutPair* out) {
    for (uint32_t k = 0; k <= entries; ++k) {
        const double next = k < entries ? lutGrad[k + 1] : 0.;
        out[k].g = lutGrad[k];
        out[k].dg = next - lutGrad[k];
    }
}

/// The solid record keeps the group id (body flag + 1, 0 = fully damaged, see Particle::grp) in the 8 low mantissa bits
/// of the sound speed: cs only enters the artificial viscosity through csbar, a relative change of 2^-44 is far below
/// the 1e-10 tolerance, and the record shrinks to 16 doubles = 128 bytes. Body flags must be below GROUP_FLAG_LIMIT.
constexpr uint32_t GROUP_FLAG_LIMIT = 254;

SPH_HD double packCsGroup(double cs, int grp) {
    uint64_t bits;
#ifdef __CUDA_ARCH__
    bits = (uint64_t)__double_as_longlong(cs);
#else
    memcpy(&bits, &cs, 8);
#endif
    bits = (bits & ~0xffull) | (uint64_t)((uint32_t)(grp + 1) & 0xffu);
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)bits);
#else
    double r;
    memcpy(&r, &bits, 8);
    return r;
#endif
}

SPH_HD void unpackCsGroup(double packed, double& cs, int& grp) {
#ifdef __CUDA_ARCH__
    const int lo = __double2loint(packed), hi = __double2hiint(packed);
    grp = (lo & 0xff) - 1;
    cs = __hiloint2double(hi, lo & ~0xff);
#else
    uint64_t bits;
    memcpy(&bits, &packed, 8);
    grp = (int)(bits & 0xffu) - 1;
    bits &= ~0xffull;
    memcpy(&cs, &bits, 8);
#endif
}

/// One directed pair i <- j. `lut` is the (dW/dq)/q table. SOLID adds velocity gradient + stress divergence (with the
/// undamaged filter when FILTER), CORRECTED adds the correction-tensor sums.
template <bool SOLID, bool CORRECTED, bool FILTER>
SPH_HD void pairAccumulate(const ParamsDev& prm, const double* __restrict__ lut, const double* __restrict__ lutW, const Particle& pi,
    const Particle& pj, double dx, double dy, double dz, double d2, double hbar, Accum& acc) {
    acc.cnt++;
    const double hInv = 1. / hbar;
    const double hInv2 = hInv * hInv;
    const double qSqr = d2 * hInv2;
    double G = 0., W = 0.;
    if (qSqr < prm.radius_sqr) {
        const double fidx = prm.q_sqr_to_idx * qSqr;
        const uint32_t k = (uint32_t)fidx;
        const double ratio = fidx - (double)k;
#ifdef __CUDA_ARCH__
        const double g0 = __ldg(lut + k), g1 = __ldg(lut + k + 1);
#else
        const double g0 = lut[k], g1 = lut[k + 1];
#endif
        G = g0 * (1. - ratio) + g1 * ratio;
        if (prm.flags & (SPHGPU_FLAG_XSPH | SPHGPU_FLAG_STRESS_AV)) { // LutKernel::valueImpl (Kernel.h:111-127)
            W = lutW[k] * (1. - ratio) + lutW[k + 1] * ratio;
        }
    }
    if (prm.flags & SPHGPU_FLAG_XSPH) {
        // XSph::Derivative::eval (XSph.h:55-63): f = eps (v_j - v_i) / rhobar W(r_i, r_j), W = hbar^-3 W(q^2); dr_i += m_j f
        const double c = pj.m * (prm.xsph_eps * (hInv2 * hInv * W) / (0.5 * (pi.rho + pj.rho)));
        acc.xs[0] += c * (pj.vx - pi.vx);
        acc.xs[1] += c * (pj.vy - pi.vy);
        acc.xs[2] += c * (pj.vz - pi.vz);
    }
    const double s = hInv2 * hInv2 * hInv * G; // h^-5 * (dW/dq)/q
    const double gx = dx * s, gy = dy * s, gz = dz * s;
    const double dvx = pj.vx - pi.vx, dvy = pj.vy - pi.vy, dvz = pj.vz - pi.vz;
    const double dvg = dvx * gx + dvy * gy + dvz * gz;
    const double mgx = pj.m * gx, mgy = pj.m * gy, mgz = pj.m * gz;
    acc.divv += pj.m * dvg;

    // artificial viscosity: w = (v_i - v_j).(r_i - r_j) = -(dv.d)
    const double w = -(dvx * dx + dvy * dy + dvz * dz);
    const bool balsara = (prm.flags & SPHGPU_FLAG_BALSARA) != 0;
    double Pi = 0.;
    if (w < 0.) {
        const double rhobar = 0.5 * (pi.rho + pj.rho);
        const double csbar = 0.5 * (pi.cs + pj.cs);
        const double mu = hbar * w / (d2 + 1.e-2 * hbar * hbar);
        Pi = (-prm.av_alpha * csbar * mu + prm.av_beta * mu * mu) / rhobar;
        if (balsara) {
            Pi *= 0.5 * (pi.bal + pj.bal);
        }
        acc.du += 0.5 * Pi * (-pj.m * dvg); // m_j * 0.5 * Pi * (v_i - v_j).gradW
    }
    if (balsara) {
        // rot v: m_j gradW x (v_j - v_i)
        acc.rot[0] += pj.m * (gy * dvz - gz * dvy);
        acc.rot[1] += pj.m * (gz * dvx - gx * dvz);
        acc.rot[2] += pj.m * (gx * dvy - gy * dvx);
        // BalsaraSwitch::Derivative::eval returns +Pi gradW (Balsara.h:69-74) where StandardAV returns -Pi gradW
        // (Standard.h:63-70); the drop-in reproduces the reference as it is
        Pi = -Pi;
    }
    // pressure gradient + AV: dv_i -= m_j (P_i + P_j + Pi) gradW
    const double c = pi.P + pj.P + Pi;
    acc.ax -= c * mgx;
    acc.ay -= c * mgy;
    acc.az -= c * mgz;

    if (prm.flags & SPHGPU_FLAG_STRESS_AV) { // StressAV::Derivative::eval (Stress.cpp:44-57) in the reference's own form
        bool ok = true;
        if (SOLID && FILTER) {
            ok = (pi.grp == pj.grp) && (pi.grp >= 0);
        }
        if (ok) {
            const double phi = prm.stress_av_factor * stressAvWeight((hInv2 * hInv * W) * pi.wpInv, prm.stress_av_exponent, prm.stress_av_int_exponent);
            double Pi[6];
            for (int q = 0; q < 6; ++q) {
                Pi[q] = phi * (pi.as[q] + pj.as[q]);
            }
            const double fx = Pi[0] * gx + Pi[3] * gy + Pi[4] * gz;
            const double fy = Pi[3] * gx + Pi[1] * gy + Pi[5] * gz;
            const double fz = Pi[4] * gx + Pi[5] * gy + Pi[2] * gz;
            acc.ax += pj.m * fx;
            acc.ay += pj.m * fy;
            acc.az += pj.m * fz;
            // heating = 1/2 (Pi (v_i - v_j)) . gradW = -1/2 (v_j - v_i) . (Pi gradW)  (Pi is symmetric)
            acc.du -= pj.m * (0.5 * (dvx * fx + dvy * fy + dvz * fz));
        }
    }

    if (prm.flags & SPHGPU_FLAG_DELTASPH) { // DeltaSph.h:37-44, 81-93, 147-163 in the reference's own form (dr = r_j - r_i = -d)
        bool ok = true;
        if (SOLID && FILTER) {
            ok = (pi.grp == pj.grp) && (pi.grp >= 0);
        }
        if (ok) {
            const double drh = pj.rho - pi.rho;
            acc.dg[0] += pj.vol * (drh * gx);
            acc.dg[1] += pj.vol * (drh * gy);
            acc.dg[2] += pj.vol * (drh * gz);
            const double cbar = 0.5 * (pi.cs + pj.cs);
            const double psx = 2. * drh * (-dx) / d2 - (pi.gr[0] + pj.gr[0]);
            const double psy = 2. * drh * (-dy) / d2 - (pi.gr[1] + pj.gr[1]);
            const double psz = 2. * drh * (-dz) / d2 - (pi.gr[2] + pj.gr[2]);
            acc.ddrho += pj.vol * ((2. * prm.deltasph_half_delta) * hbar * cbar * (psx * gx + psy * gy + psz * gz));
            const double pij = (dvx * (-dx) + dvy * (-dy) + dvz * (-dz)) / d2;
            const double f = (2. * prm.deltasph_half_alpha) * hbar * cbar * pij;
            acc.ax += pj.vol * (f * gx);
            acc.ay += pj.vol * (f * gy);
            acc.az += pj.vol * (f * gz);
        }
    }

    if (SOLID) {
        bool ok = true;
        if (FILTER) {
            ok = (pi.grp == pj.grp) && (pi.grp >= 0);
        }
        if (ok) {
            // stress divergence: dv_i += m_j (S_i/rho_i^2 + S_j/rho_j^2) gradW
            const double sxx = pi.Sr[0] + pj.Sr[0], syy = pi.Sr[1] + pj.Sr[1], sxy = pi.Sr[2] + pj.Sr[2],
                         sxz = pi.Sr[3] + pj.Sr[3], syz = pi.Sr[4] + pj.Sr[4];
            const double szz = -sxx - syy;
            acc.ax += sxx * mgx + sxy * mgy + sxz * mgz;
            acc.ay += sxy * mgx + syy * mgy + syz * mgz;
            acc.az += sxz * mgx + syz * mgy + szz * mgz;
            // velocity gradient (uncorrected outer product; C_i is applied once in the epilogue)
            acc.T[0] += dvx * mgx;
            acc.T[1] += dvx * mgy;
            acc.T[2] += dvx * mgz;
            acc.T[3] += dvy * mgx;
            acc.T[4] += dvy * mgy;
            acc.T[5] += dvy * mgz;
            acc.T[6] += dvz * mgx;
            acc.T[7] += dvz * mgy;
            acc.T[8] += dvz * mgz;
            if (CORRECTED) {
                // (r_j - r_i) (x) gradW = -s d (x) d  (symmetric)
                const double vs = -pj.vol * s;
                const double vdx = vs * dx, vdy = vs * dy, vdz = vs * dz;
                acc.Cm[0] += vdx * dx;
                acc.Cm[1] += vdy * dy;
                acc.Cm[2] += vdz * dz;
                acc.Cm[3] += vdx * dy;
                acc.Cm[4] += vdx * dz;
                acc.Cm[5] += vdy * dz;
            }
        }
    }
}

/// floor(x) and x - floor(x) of 0 <= x < 2^32 without the (quarter-rate) F2I / I2F conversions: adding 2^52 - 1/2 rounds
/// x - 1/2 to the nearest integer, which then sits in the low mantissa bits. At an exact integer x the tie may go down
/// (k = x - 1, frac = 1) instead of (k = x, frac = 0): the interpolated value is the same table entry either way.
SPH_HD void floorFrac(double x, uint32_t& k, double& frac) {
    const double MAGIC = 4503599627370496. - 0.5; // 2^52 - 1/2 (exact)
#ifdef __CUDA_ARCH__
    const double t = __dadd_rn(x, MAGIC);
    k = (uint32_t)__double2loint(t);
    frac = x - __dadd_rn(t, -4503599627370496.);
#else
    volatile double t = x + MAGIC;
    uint64_t bits;
    double tt = t;
    memcpy(&bits, &tt, 8);
    k = (uint32_t)bits;
    frac = x - (tt - 4503599627370496.);
#endif
}

// ---- the branch-free pair body of the tiled kernels, in two stages -----------------------------------------------------
// Stage A (pairGeometry) needs only the positions, h and the densities of the pair: exact predicate, the merged
// reciprocal, the kernel argument and the table index. Stage B (pairSums) needs the rest of the neighbour's record and
// the table entry. k_pair_sum runs stage A of entry q + 1 before stage B of entry q, so that the table load (L2 latency)
// and the remaining record loads of a pair are in flight while the previous pair is summed.
// `valid` (the exact neighbour predicate) scales the neighbour's mass / volume to zero instead of branching, and the AV
// condition is a min: straight-line code. About 95 FP64 instructions per pair (solid, corrected), no conversions.

/// What stage A hands to stage B.
struct PairGeom {
    double dx, dy, dz; // r_i - r_j
    double hbar;
    double hInv5;      // hbar^-5: gradW = hInv5 G(q^2) (r_i - r_j)
    double ratio;      // interpolation weight inside table entry k
    double iD, irs;    // 1 / (d^2 + 0.01 hbar^2) and 1 / (rho_i + rho_j)
    uint32_t k;
    bool valid;
};

SPH_HD void pairGeometry(const ParamsDev& prm, double xi, double yi, double zi, double hi, double rhoi, double xj, double yj, double zj,
    double hj, double rhoj, PairGeom& g) {
    g.dx = xi - xj;
    g.dy = yi - yj;
    g.dz = zi - zj;
    double d2, sq[4];
    g.valid = isNeighbour(g.dx, g.dy, g.dz, hi, hj, prm.kernel_radius, d2, g.hbar, sq);
    // one reciprocal serves 1/hbar (kernel) and 1/(D rhobar) (viscosity); the factors 1/2 of rhobar and csbar are folded:
    // rs = 2 rhobar, inv = 1 / (hbar D rs)  =>  1/hbar = D rs inv,  1/(D rs) = hbar inv
    const double rs = rhoi + rhoj;
    const double D = fma(prm.av_eps_over_radius_sqr, sq[3], d2); // d^2 + 0.01 hbar^2, from (R hbar)^2
    const double A = D * rs;
    const double inv = fastRcp(g.hbar * A);
    const double hInv = A * inv;
    const double invA = g.hbar * inv;
    const double hInv2 = hInv * hInv;
    const double qSqr = d2 * hInv2;
    // branch-free table lookup: the index is clamped to the zero guard entry behind the table; rejected candidates only
    // ever read a clamped (finite) slot and are masked in stage B. A valid pair has qSqr < R^2 up to rounding, where the
    // table goes to zero quadratically, so the reference's extra `qSqr < R^2` test changes nothing above 1e-30.
    uint32_t k;
    floorFrac(prm.q_sqr_to_idx * qSqr, k, g.ratio);
    g.k = k < prm.lut_entries ? k : prm.lut_entries;
    g.hInv5 = hInv2 * hInv2 * hInv;
    g.iD = rs * invA;
    g.irs = D * invA;
}

/// Stage B. G = the interpolated table value g + ratio dg of entry g.k. Of pi only v, P, cs, grp are read; of pj v, m,
/// P, cs, vol, Sr, grp. The stress sum is split: sum_j (Sr_i + Sr_j) f_j = Sr_i F + sum_j Sr_j f_j with F = sum_j f_j
/// (acc.F, applied by finalizeParticle), which saves the target's Sr registers and two additions per pair.
template <bool SOLID, bool CORRECTED, bool FILTER, bool BALSARA = false, bool XSPH = false, bool DELTA = false, bool STRESSAV = false>
SPH_HD void pairSums(const ParamsDev& prm, const Particle& pi, const Particle& pj, const PairGeom& g, double G, Accum& acc, double W = 0.) {
    acc.cnt += g.valid ? 1u : 0u;
    const double mj = selectD(g.valid, pj.m, 0.);
    if (XSPH) { // W = table value of the kernel at q^2; hbar^-3 = hbar^-5 hbar^2; 1 / rhobar = 2 / (rho_i + rho_j)
        const double c = mj * ((2. * prm.xsph_eps) * W) * ((g.hInv5 * g.hbar) * g.hbar) * g.irs;
        acc.xs[0] = fma(c, pj.vx - pi.vx, acc.xs[0]);
        acc.xs[1] = fma(c, pj.vy - pi.vy, acc.xs[1]);
        acc.xs[2] = fma(c, pj.vz - pi.vz, acc.xs[2]);
    }
    const double s = g.hInv5 * G; // gradW = s d
    const double dx = g.dx, dy = g.dy, dz = g.dz;
    const double dvx = pj.vx - pi.vx, dvy = pj.vy - pi.vy, dvz = pj.vz - pi.vz;
    const double t = dvx * dx + dvy * dy + dvz * dz; // (v_j - v_i).(r_i - r_j)
    const double dvg = t * s;                        // (v_j - v_i).gradW
    const double ms = mj * s;
    const double mgx = dx * ms, mgy = dy * ms, mgz = dz * ms;
    const double mdvg = mj * dvg;
    acc.divv += mdvg;
    // StandardAV: mu = hbar w / D, Pi = (-alpha csbar mu + beta mu^2) / rhobar for approaching pairs (w < 0); with
    // w clamped to min(w, 0) the receding pairs give mu = 0 and Pi = 0 exactly, without a branch. Q = Pi / 2.
    const double w = fmin(-t, 0.);
    const double mu = (g.hbar * w) * g.iD;
    double Q = mu * fma(prm.av_beta, mu, prm.av_minus_half_alpha * (pi.cs + pj.cs)) * g.irs;
    if (BALSARA) {
        Q *= 0.5 * (pi.bal + pj.bal);
        // rot v: m_j gradW x (v_j - v_i) = (m_j s) d x dv
        acc.rot[0] = fma(ms, dy * dvz - dz * dvy, acc.rot[0]);
        acc.rot[1] = fma(ms, dz * dvx - dx * dvz, acc.rot[1]);
        acc.rot[2] = fma(ms, dx * dvy - dy * dvx, acc.rot[2]);
    }
    acc.du -= Q * mdvg;
    // (the Balsara-switched viscosity enters the acceleration with the opposite sign, as in Balsara.h:69-74)
    const double c = fma(BALSARA ? -2. : 2., Q, pi.P + pj.P);
    acc.ax -= c * mgx;
    acc.ay -= c * mgy;
    acc.az -= c * mgz;
    if (STRESSAV) {
        // StressAV::Derivative (Stress.cpp:44-57), SUM_ONLY_UNDAMAGED: Pi = phi (A_i + A_j) with A = as / rho^2 and
        // phi = xi (W_ij / W_0i)^n, W_ij = hbar^-3 W(q^2) (W: the table value, 0 for a rejected candidate);
        // dv_i += m_j Pi gradW, du_i += m_j (Pi (v_i - v_j)) . gradW / 2 = -m_j (v_j - v_i) . (Pi gradW) / 2
        bool ok = g.valid;
        if (SOLID && FILTER) {
            ok = g.valid && (pi.grp == pj.grp) && (pi.grp >= 0);
        }
        const double w = W * ((g.hInv5 * g.hbar) * g.hbar);
        const double phi = prm.stress_av_factor * stressAvWeight(w * pi.wpInv, prm.stress_av_exponent, prm.stress_av_int_exponent);
        const double ps = (selectD(ok, pj.m, 0.) * s) * phi; // m_j phi gradW = ps d
        const double a0 = pi.as[0] + pj.as[0], a1 = pi.as[1] + pj.as[1], a2 = pi.as[2] + pj.as[2];
        const double a3 = pi.as[3] + pj.as[3], a4 = pi.as[4] + pj.as[4], a5 = pi.as[5] + pj.as[5];
        const double fx = ps * fma(a4, dz, fma(a3, dy, a0 * dx));
        const double fy = ps * fma(a5, dz, fma(a1, dy, a3 * dx));
        const double fz = ps * fma(a2, dz, fma(a5, dy, a4 * dx));
        acc.ax += fx;
        acc.ay += fy;
        acc.az += fz;
        acc.du -= 0.5 * (dvx * fx + dvy * fy + dvz * fz);
    }
    if (DELTA) {
        // The three delta-SPH derivatives carry SUM_ONLY_UNDAMAGED (DeltaSph.h:22-24,63-64,135-136), which acts where the
        // Storage has STRESS_REDUCING, i.e. for solids (DerivativeHelpers.h:84-91).
        bool ok = g.valid;
        if (SOLID && FILTER) {
            ok = g.valid && (pi.grp == pj.grp) && (pi.grp >= 0);
        }
        const double vs = selectD(ok, pj.vol, 0.) * s; // V_j gradW = vs d, with d = r_i - r_j
        const double drh = pj.rho - pi.rho;
        // RenormalizedDensityGradient: V_j (rho_j - rho_i) gradW; the correction tensor C_i is applied once to the sum
        const double cg = vs * drh;
        acc.dg[0] = fma(cg, dx, acc.dg[0]);
        acc.dg[1] = fma(cg, dy, acc.dg[1]);
        acc.dg[2] = fma(cg, dz, acc.dg[2]);
        // DensityDiffusion: psi = 2 (rho_j - rho_i) dr / dr^2 - (G_i + G_j) with dr = -d, so that
        // psi . gradW = -s (2 (rho_j - rho_i) + (G_i + G_j) . d); the sum gets V_j delta hbar cbar psi . gradW
        const double gd = (pi.gr[0] + pj.gr[0]) * dx + (pi.gr[1] + pj.gr[1]) * dy + (pi.gr[2] + pj.gr[2]) * dz;
        const double hc = g.hbar * (pi.cs + pj.cs); // 2 hbar cbar
        acc.ddrho -= (vs * (prm.deltasph_half_delta * hc)) * fma(2., drh, gd);
        // VelocityDiffusion: pi_ij = (v_j - v_i) . dr / dr^2 = -t / d^2; dv_i += V_j alpha hbar cbar pi_ij gradW
        // (the self pair is a masked candidate with d = 0: keep its reciprocal finite, vs = 0 removes it)
        const double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
        const double cv = (vs * (prm.deltasph_half_alpha * hc)) * (t * fastRcp(g.valid ? d2 : 1.));
        acc.ax -= cv * dx;
        acc.ay -= cv * dy;
        acc.az -= cv * dz;
    }
    if (SOLID) {
        bool ok = g.valid;
        if (FILTER) {
            ok = g.valid && (pi.grp == pj.grp) && (pi.grp >= 0);
        }
        // masked mass / volume instead of a branch around the tensor sums
        const double fs = selectD(ok, pj.m, 0.) * s;
        const double fx = dx * fs, fy = dy * fs, fz = dz * fs;
        acc.F[0] += fx;
        acc.F[1] += fy;
        acc.F[2] += fz;
        const double sxx = pj.Sr[0], syy = pj.Sr[1], sxy = pj.Sr[2], sxz = pj.Sr[3], syz = pj.Sr[4];
        const double szz = -sxx - syy;
        acc.ax = fma(sxz, fz, fma(sxy, fy, fma(sxx, fx, acc.ax)));
        acc.ay = fma(syz, fz, fma(syy, fy, fma(sxy, fx, acc.ay)));
        acc.az = fma(szz, fz, fma(syz, fy, fma(sxz, fx, acc.az)));
        acc.T[0] += dvx * fx;
        acc.T[1] += dvx * fy;
        acc.T[2] += dvx * fz;
        acc.T[3] += dvy * fx;
        acc.T[4] += dvy * fy;
        acc.T[5] += dvy * fz;
        acc.T[6] += dvz * fx;
        acc.T[7] += dvz * fy;
        acc.T[8] += dvz * fz;
        if (CORRECTED) {
            // (r_j - r_i) (x) gradW = -s d (x) d  (symmetric)
            const double vs = -selectD(ok, pj.vol, 0.) * s;
            const double vdx = vs * dx, vdy = vs * dy, vdz = vs * dz;
            acc.Cm[0] += vdx * dx;
            acc.Cm[1] += vdy * dy;
            acc.Cm[2] += vdz * dz;
            acc.Cm[3] += vdx * dy;
            acc.Cm[4] += vdx * dz;
            acc.Cm[5] += vdy * dz;
        }
    }
}

/// Both stages for one pair (host formula tests).
template <bool SOLID, bool CORRECTED, bool FILTER>
SPH_HD void pairAccumulateMasked(const ParamsDev& prm, const LutPair* __restrict__ lut2, const Particle& pi, const Particle& pj,
    Accum& acc, const LutPair* __restrict__ lutW2 = nullptr) {
    PairGeom g;
    pairGeometry(prm, pi.x, pi.y, pi.z, pi.h, pi.rho, pj.x, pj.y, pj.z, pj.h, pj.rho, g);
    const double G = fma(g.ratio, lut2[g.k].dg, lut2[g.k].g);
    if (prm.flags & SPHGPU_FLAG_STRESS_AV) {
        pairSums<SOLID, CORRECTED, FILTER, false, false, false, true>(prm, pi, pj, g, G, acc, fma(g.ratio, lutW2[g.k].dg, lutW2[g.k].g));
    } else if (prm.flags & SPHGPU_FLAG_DELTASPH) {
        pairSums<SOLID, CORRECTED, FILTER, false, false, true>(prm, pi, pj, g, G, acc);
    } else if (prm.flags & SPHGPU_FLAG_XSPH) {
        pairSums<SOLID, CORRECTED, FILTER, false, true>(prm, pi, pj, g, G, acc, fma(g.ratio, lutW2[g.k].dg, lutW2[g.k].g));
    } else if (prm.flags & SPHGPU_FLAG_BALSARA) {
        pairSums<SOLID, CORRECTED, FILTER, true>(prm, pi, pj, g, G, acc);
    } else {
        pairSums<SOLID, CORRECTED, FILTER, false>(prm, pi, pj, g, G, acc);
    }
}

/// Derivatives of one particle (everything IAsymmetricSolver::afterLoop leaves in the Storage for particle i).
struct Derivs {
    double ax, ay, az, vh, du, drho, divv;
    double rot[3];
    double xs[3];
    double dg[3]; // DELTASPH_DENSITY_GRADIENT of this evaluation (delta-SPH terms only)
    double dS[5];
    double gradv[6];
    double corr[6];
    uint32_t ncnt;
};

/// accumulated.store + equations.finalize for one particle.
/// S = yielded stress of i (unscaled), p = reduced pressure.
template <bool SOLID, bool CORRECTED>
SPH_HD void finalizeParticle(const ParamsDev& prm, const MaterialDev& mat, const Accum& acc, double h, double rho, double p,
    double cs, double reduce, const double S[5], Derivs& out) {
    const double rhoInv = 1. / rho;
    out.ncnt = acc.cnt;
    out.ax = acc.ax;
    out.ay = acc.ay;
    out.az = acc.az;
    if (SOLID) { // the target's half of the stress divergence, (S_i / rho_i^2) . F  (zero F for the unsplit variants)
        const double r2 = rhoInv * rhoInv;
        const double sxx = S[0] * r2, syy = S[1] * r2, sxy = S[2] * r2, sxz = S[3] * r2, syz = S[4] * r2;
        out.ax += sxx * acc.F[0] + sxy * acc.F[1] + sxz * acc.F[2];
        out.ay += sxy * acc.F[0] + syy * acc.F[1] + syz * acc.F[2];
        out.az += sxz * acc.F[0] + syz * acc.F[1] + (-sxx - syy) * acc.F[2];
    }
    out.divv = acc.divv * rhoInv;
    out.rot[0] = acc.rot[0] * rhoInv;
    out.rot[1] = acc.rot[1] * rhoInv;
    out.rot[2] = acc.rot[2] * rhoInv;
    double du = acc.du;
    double trGradv = 0.;
    if (SOLID) {
        double C[6] = { 1., 1., 1., 0., 0., 0. };
        if (CORRECTED) {
            const double* c = acc.Cm;
            const bool isNull = c[0] == 0. && c[1] == 0. && c[2] == 0. && c[3] == 0. && c[4] == 0. && c[5] == 0.;
            if (!isNull) {
                const double det = c[0] * c[1] * c[2] + 2. * c[3] * c[4] * c[5] -
                                   (c[3] * c[3] * c[2] + c[4] * c[4] * c[1] + c[5] * c[5] * c[0]);
                if (det > 0.01) {
                    C[0] = (c[1] * c[2] - c[5] * c[5]) / det;
                    C[1] = (c[2] * c[0] - c[4] * c[4]) / det;
                    C[2] = (c[0] * c[1] - c[3] * c[3]) / det;
                    C[3] = (c[4] * c[5] - c[2] * c[3]) / det;
                    C[4] = (c[5] * c[3] - c[1] * c[4]) / det;
                    C[5] = (c[3] * c[4] - c[0] * c[5]) / det;
                }
            }
        }
        for (int k = 0; k < 6; ++k) {
            out.corr[k] = C[k];
        }
        // gradv = sym(T * C) / rho ; C symmetric: rows {C0,C3,C4},{C3,C1,C5},{C4,C5,C2}
        const double* T = acc.T;
        double M[9];
        for (int a = 0; a < 3; ++a) {
            M[3 * a + 0] = T[3 * a] * C[0] + T[3 * a + 1] * C[3] + T[3 * a + 2] * C[4];
            M[3 * a + 1] = T[3 * a] * C[3] + T[3 * a + 1] * C[1] + T[3 * a + 2] * C[5];
            M[3 * a + 2] = T[3 * a] * C[4] + T[3 * a + 1] * C[5] + T[3 * a + 2] * C[2];
        }
        out.gradv[0] = M[0] * rhoInv;
        out.gradv[1] = M[4] * rhoInv;
        out.gradv[2] = M[8] * rhoInv;
        out.gradv[3] = 0.5 * (M[1] + M[3]) * rhoInv;
        out.gradv[4] = 0.5 * (M[2] + M[6]) * rhoInv;
        out.gradv[5] = 0.5 * (M[5] + M[7]) * rhoInv;
        trGradv = out.gradv[0] + out.gradv[1] + out.gradv[2];
    }
    out.dg[0] = out.dg[1] = out.dg[2] = 0.; // (finalizeDeltaSph)
    // smoothing length (AdaptiveSmoothingLength::finalize / ConstSmoothingLength::finalize)
    double vh = 0.;
    if (prm.flags & SPHGPU_FLAG_ADAPTIVE_H) {
        if (h > 2. * prm.h_min) {
            vh = h / 3. * out.divv;
        }
        if ((prm.flags & SPHGPU_FLAG_SOUND_SPEED_ENFORCING) && !(prm.neigh_enforcing <= -1.e2)) {
            const double dn1 = (double)acc.cnt - prm.neigh_upper;
            if (dn1 > 0.) {
                vh -= exp(prm.neigh_enforcing * dn1) * cs;
            } else {
                const double dn2 = prm.neigh_lower - (double)acc.cnt;
                if (dn2 > 0.) {
                    vh += exp(prm.neigh_enforcing * dn2) * cs;
                }
            }
        }
    }
    out.vh = vh;
    // continuity equation
    if (SOLID && prm.continuity_mode == SPHGPU_CONTINUITY_SUM_ONLY_UNDAMAGED && reduce > 0.) {
        out.drho = -rho * trGradv;
    } else {
        out.drho = -rho * out.divv;
    }
    // solid stress: du += S:gradv / rho ; dS = 2 mu (gradv - tr/3 I)
    for (int k = 0; k < 5; ++k) {
        out.dS[k] = 0.;
    }
    if (SOLID && mat.yielding != SPHGPU_YIELD_NONE && mat.yielding != SPHGPU_YIELD_DUST) {
        const double* gv = out.gradv;
        const double ddot = (S[0] * gv[0] + S[1] * gv[1] + (-S[0] - S[1]) * gv[2]) + 2. * (S[2] * gv[3] + S[3] * gv[4] + S[4] * gv[5]);
        du += rhoInv * ddot;
        const double tr3 = trGradv / 3.;
        const double mu2 = 2. * mat.shear_modulus;
        out.dS[0] = mu2 * (gv[0] - tr3);
        out.dS[1] = mu2 * (gv[1] - tr3);
        out.dS[2] = mu2 * gv[3];
        out.dS[3] = mu2 * gv[4];
        out.dS[4] = mu2 * gv[5];
    }
    if (prm.forces & SPHGPU_FORCE_PRESSURE) {
        du -= p * rhoInv * out.divv;
    }
    out.du = du;
}

/// The delta-SPH part of the epilogue, after finalizeParticle (kept apart so that the gradient and diffusion sums are dead
/// registers in the kernels without the terms). RenormalizedDensityGradient is a CORRECTED derivative: C_i gradW summed =
/// C_i applied to the sum (out.corr: the identity without the tensor); DensityDiffusion shares the density-derivative
/// buffer with the continuity equation (EquationTerm.cpp:289-316).
template <bool SOLID>
SPH_HD void finalizeDeltaSph(const Accum& acc, Derivs& out) {
    if (SOLID) {
        const double* C = out.corr;
        out.dg[0] = C[0] * acc.dg[0] + C[3] * acc.dg[1] + C[4] * acc.dg[2];
        out.dg[1] = C[3] * acc.dg[0] + C[1] * acc.dg[1] + C[5] * acc.dg[2];
        out.dg[2] = C[4] * acc.dg[0] + C[5] * acc.dg[1] + C[2] * acc.dg[2];
    } else {
        out.dg[0] = acc.dg[0];
        out.dg[1] = acc.dg[1];
        out.dg[2] = acc.dg[2];
    }
    out.drho += acc.ddrho;
}

// ---- time stepping ----------------------------------------------------------------------------------------

SPH_HD bool rangeBounded(double lo, double hi) {
    return !(lo <= -INFTY_REF && hi >= INFTY_REF);
}

/// clampWithDerivative<Float>, core/objects/wrappers/Interval.h:159-162
SPH_HD void clampWithDerivative(double& v, double& dv, double lo, double hi) {
    const bool zeroDeriv = (v >= hi && dv > 0.) || (v <= lo && dv < 0.);
    v = fmax(lo, fmin(v, hi));
    if (zeroDeriv) {
        dv = 0.;
    }
}

/// One component of DerivativeCriterion::computeImpl, core/timestepping/TimeStepCriterion.cpp:150-176
SPH_HD double derivativeStep(double absv, double absdv, double minValue, double factor) {
    if (absv < 2. * minValue) {
        return INFTY_REF;
    }
    return factor * (absv + minValue) / (absdv + EPS_REF);
}

} // namespace sph
