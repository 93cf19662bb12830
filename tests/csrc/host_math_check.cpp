// Test-only harness: runs the product's per-particle / per-pair arithmetic (opensph_b200/csrc/sph_math.cuh, the very
// functions the CUDA kernels call) on the CPU over neighbour lists supplied by the caller, so the formulas can be
// checked against the oracle without a GPU. Not part of the product; built by tests/test_host_math.py.
#include "../../opensph_b200/csrc/sph_math.cuh"
#include "../../oracle/sph_oracle.h"
#include <vector>

using namespace sph;

template <bool SOLID, bool CORRECTED, bool FILTER, bool MASKED>
static void run(orc_state* s, const ParamsDev& prm, const std::vector<MaterialDev>& mats, const std::vector<uint32_t>& matid,
    const double* lut, const LutPair* lut2, const double* lutW, const LutPair* lutW2, const uint64_t* off, const uint32_t* idx, bool hasReduce,
    bool hasDamage) {
    const uint32_t n = s->n;
    std::vector<Particle> P(n);
    for (uint32_t i = 0; i < n; ++i) {
        const MaterialDev& mat = mats[matid[i]];
        double p = s->p[i], cs = s->cs[i];
        evalEos(mat, s->rho[i], s->u[i], p, cs);
        double S[5] = { 0, 0, 0, 0, 0 };
        double reduce = hasReduce ? s->reduce[i] : 1.;
        if (SOLID) {
            for (int k = 0; k < 5; ++k) S[k] = s->S[5 * (size_t)i + k];
        }
        if (mat.yielding == SPHGPU_YIELD_VON_MISES) {
            const bool dmg = hasDamage && mat.fracture != SPHGPU_FRACTURE_NONE;
            const double D = dmg ? s->damage[i] : 0.;
            reduce = vonMises(mat, s->u[i], D, dmg, p, S);
            s->reduce[i] = reduce;
            if (SOLID) {
                for (int k = 0; k < 5; ++k) s->S[5 * (size_t)i + k] = S[k];
            }
            if (dmg) {
                s->ddamage[i] = damageRate(mat, p, S, D, s->eps_min[i], s->m_zero[i], s->growth[i], s->n_flaws[i]);
            }
        }
        s->p[i] = p;
        s->cs[i] = cs;
        if (prm.flags & SPHGPU_FLAG_ADAPTIVE_H) {
            s->pos[4 * (size_t)i + 3] = fmax(prm.h_min, fmin(s->pos[4 * (size_t)i + 3], prm.h_max));
        }
        Particle& q = P[i];
        q.x = s->pos[4 * (size_t)i]; q.y = s->pos[4 * (size_t)i + 1]; q.z = s->pos[4 * (size_t)i + 2]; q.h = s->pos[4 * (size_t)i + 3];
        if (prm.flags & SPHGPU_FLAG_XSPH) { // k_prologue_pack: XSph::initialize
            for (int k = 0; k < 3; ++k) s->vel[4 * (size_t)i + k] -= s->xsph[4 * (size_t)i + k];
        }
        q.vx = s->vel[4 * (size_t)i]; q.vy = s->vel[4 * (size_t)i + 1]; q.vz = s->vel[4 * (size_t)i + 2];
        q.m = s->mass[i]; q.rho = s->rho[i]; q.cs = cs;
        const double r2 = 1. / (q.rho * q.rho);
        q.P = p * r2; q.vol = q.m / q.rho;
        for (int k = 0; k < 5; ++k) q.Sr[k] = S[k] * r2;
        q.grp = (hasReduce && reduce == 0.) ? -1 : (int)s->flag[i];
        if (prm.flags & SPHGPU_FLAG_STRESS_AV) { // k_prologue_pack: StressAV::initialize on the yielded stress and reduced pressure
            const double sigma[6] = { S[0] - p, S[1] - p, (-S[0] - S[1]) - p, S[2], S[3], S[4] };
            double as[6];
            avStressOf(sigma, as);
            for (int k = 0; k < 6; ++k) {
                s->av_stress[6 * (size_t)i + k] = as[k];
                q.as[k] = as[k] * r2;
            }
            q.wpInv = 1. / s->wp[i];
        }
        if (prm.flags & SPHGPU_FLAG_DELTASPH) { // k_prologue_pack: the gradient of the previous evaluation rides in the record
            for (int k = 0; k < 3; ++k) q.gr[k] = s->drho_grad[4 * (size_t)i + k];
        }
        if (MASKED) { // what the device loaders see: cs carries the group id, m is rebuilt from vol * rho
            unpackCsGroup(packCsGroup(q.cs, q.grp), q.cs, q.grp);
            q.m = q.vol * q.rho;
        }
    }
    std::vector<double> xsNew((prm.flags & SPHGPU_FLAG_XSPH) ? 3 * (size_t)n : 0);
    for (uint32_t i = 0; i < n; ++i) {
        Accum acc;
        accumZero(acc);
        for (uint64_t k = off[i]; k < off[i + 1]; ++k) {
            const Particle& pj = P[idx[k]];
            const double dx = P[i].x - pj.x, dy = P[i].y - pj.y, dz = P[i].z - pj.z;
            double d2, hbar, sq[4];
            const bool valid = isNeighbour(dx, dy, dz, P[i].h, pj.h, prm.kernel_radius, d2, hbar, sq);
            if (MASKED) {
                // the tiled kernel's branch-free body; also fed one non-neighbour per target to exercise the masking
                pairAccumulateMasked<SOLID, CORRECTED, FILTER>(prm, lut2, P[i], pj, acc, lutW2);
            } else if (valid) {
                pairAccumulate<SOLID, CORRECTED, FILTER>(prm, lut, lutW, P[i], pj, dx, dy, dz, d2, hbar, acc);
            }
        }
        if (MASKED) {
            const Particle& pj = P[(i + n / 2) % n]; // an arbitrary (almost surely non-neighbour) candidate, masked out
            const double dx = P[i].x - pj.x, dy = P[i].y - pj.y, dz = P[i].z - pj.z;
            double d2, hbar, sq[4];
            const bool valid = isNeighbour(dx, dy, dz, P[i].h, pj.h, prm.kernel_radius, d2, hbar, sq);
            if (!valid) pairAccumulateMasked<SOLID, CORRECTED, FILTER>(prm, lut2, P[i], pj, acc, lutW2);
        }
        double S[5] = { 0, 0, 0, 0, 0 };
        if (SOLID) {
            for (int k = 0; k < 5; ++k) S[k] = s->S[5 * (size_t)i + k];
        }
        Derivs o;
        finalizeParticle<SOLID, CORRECTED>(prm, mats[matid[i]], acc, P[i].h, P[i].rho, s->p[i], P[i].cs, hasReduce ? s->reduce[i] : 1., S, o);
        if (prm.flags & SPHGPU_FLAG_DELTASPH) finalizeDeltaSph<SOLID>(acc, o);
        s->acc[4 * (size_t)i] = o.ax; s->acc[4 * (size_t)i + 1] = o.ay; s->acc[4 * (size_t)i + 2] = o.az; s->acc[4 * (size_t)i + 3] = 0.;
        s->vel[4 * (size_t)i + 3] = o.vh;
        s->du[i] = o.du; s->drho[i] = o.drho; s->divv[i] = o.divv; s->ncnt[i] = o.ncnt;
        if (prm.flags & SPHGPU_FLAG_XSPH) { // storeXsph: XSph::finalize
            for (int k = 0; k < 3; ++k) {
                s->xsph[4 * (size_t)i + k] = acc.xs[k];
                xsNew[3 * (size_t)i + k] = acc.xs[k];
            }
        }
        if (prm.flags & SPHGPU_FLAG_DELTASPH) { // storeDerivs (the records P[] keep the old gradient for the other targets)
            for (int k = 0; k < 3; ++k) s->drho_grad[4 * (size_t)i + k] = o.dg[k];
            s->drho_grad[4 * (size_t)i + 3] = 0.;
        }
        if (SOLID) {
            for (int k = 0; k < 5; ++k) s->dS[5 * (size_t)i + k] = o.dS[k];
            for (int k = 0; k < 6; ++k) s->gradv[6 * (size_t)i + k] = o.gradv[k];
            if (CORRECTED) {
                for (int k = 0; k < 6; ++k) s->corr[6 * (size_t)i + k] = o.corr[k];
            }
        }
    }
    for (size_t k = 0; k < xsNew.size(); ++k) {
        s->vel[4 * (k / 3) + k % 3] += xsNew[k];
    }
}

template <bool MASKED>
static int dispatch(orc_state* s, const sphgpu_config* cfg, const sphgpu_material* mats, uint32_t nmat,
    const uint64_t* off, const uint32_t* idx) {
    ParamsDev prm{};
    prm.forces = cfg->forces; prm.flags = cfg->flags; prm.continuity_mode = cfg->continuity_mode; prm.lut_entries = cfg->lut_entries;
    prm.kernel_radius = cfg->kernel_radius; prm.radius_sqr = cfg->kernel_radius * cfg->kernel_radius;
    prm.q_sqr_to_idx = (double)cfg->lut_entries * (1. / (cfg->kernel_radius * cfg->kernel_radius));
    prm.av_alpha = cfg->av_alpha; prm.av_beta = cfg->av_beta;
    prm.av_minus_half_alpha = -0.5 * cfg->av_alpha; prm.av_eps_over_radius_sqr = 1.e-2 / (cfg->kernel_radius * cfg->kernel_radius); prm.h_min = cfg->h_min; prm.h_max = cfg->h_max;
    prm.neigh_enforcing = cfg->neigh_enforcing; prm.neigh_lower = cfg->neigh_lower; prm.neigh_upper = cfg->neigh_upper;
    std::vector<MaterialDev> md(nmat);
    std::vector<uint32_t> matid(s->n, 0);
    bool hasReduce = false, hasDamage = false;
    for (uint32_t m = 0; m < nmat; ++m) {
        const sphgpu_material& a = mats[m];
        MaterialDev& b = md[m];
        b.eos = a.eos; b.yielding = a.yielding;
        b.fracture = a.yielding == SPHGPU_YIELD_VON_MISES ? a.fracture : (uint32_t)SPHGPU_FRACTURE_NONE;
        b.til_u0 = a.til_u0; b.til_uiv = a.til_uiv; b.til_ucv = a.til_ucv; b.til_a = a.til_a; b.til_b = a.til_b;
        b.rho0 = a.rho0; b.til_A = a.til_A; b.til_B = a.til_B; b.til_alpha = a.til_alpha; b.til_beta = a.til_beta; b.gamma = a.gamma;
        b.shear_modulus = a.shear_modulus; b.elasticity_limit = a.elasticity_limit; b.melt_energy = a.melt_energy;
        b.young_modulus = a.young_modulus; b.d_min = a.d_min; b.d_max = a.d_max;
        hasReduce |= (a.yielding == SPHGPU_YIELD_VON_MISES || a.yielding == SPHGPU_YIELD_ELASTIC);
        hasDamage |= b.fracture != SPHGPU_FRACTURE_NONE;
        for (uint32_t i = a.begin; i < a.end; ++i) matid[i] = m;
    }
    // same guard entry behind the table as the device copy (api.cu allocates lut_entries + 2 zero-initialised doubles)
    std::vector<double> lutGuard(cfg->lut_grad, cfg->lut_grad + cfg->lut_entries + 1);
    lutGuard.push_back(0.);
    const double* lutPtr = lutGuard.data();
    std::vector<LutPair> lutPairs(cfg->lut_entries + 1);
    buildLutPairs(cfg->lut_grad, cfg->lut_entries, lutPairs.data());
    const LutPair* lut2 = lutPairs.data();
    // kernel values for the XSph term (api.cu uploads them the same way)
    std::vector<double> lutWGuard;
    std::vector<LutPair> lutWPairs;
    if (cfg->flags & (SPHGPU_FLAG_XSPH | SPHGPU_FLAG_STRESS_AV)) {
        lutWGuard.assign(cfg->lut_value, cfg->lut_value + cfg->lut_entries + 1);
        lutWGuard.push_back(0.);
        lutWPairs.resize(cfg->lut_entries + 1);
        buildLutPairs(cfg->lut_value, cfg->lut_entries, lutWPairs.data());
    }
    const double* lutWPtr = lutWGuard.data();
    const LutPair* lutW2 = lutWPairs.data();
    prm.xsph_eps = s->xsph_eps;
    prm.deltasph_half_delta = 0.5 * s->deltasph_delta;
    prm.deltasph_half_alpha = 0.5 * s->deltasph_alpha;
    prm.stress_av_exponent = s->stress_av_exponent;
    prm.stress_av_factor = s->stress_av_factor;
    prm.stress_av_int_exponent = stressAvIntExponent(s->stress_av_exponent);
    const bool solid = cfg->forces & SPHGPU_FORCE_SOLID_STRESS;
    const bool corrected = solid && (cfg->flags & SPHGPU_FLAG_CORRECTION_TENSOR);
    const bool filter = solid && (cfg->flags & SPHGPU_FLAG_SUM_ONLY_UNDAMAGED) && hasReduce;
    if (!solid) run<false, false, false, MASKED>(s, prm, md, matid, lutPtr, lut2, lutWPtr, lutW2, off, idx, hasReduce, hasDamage);
    else if (corrected && filter) run<true, true, true, MASKED>(s, prm, md, matid, lutPtr, lut2, lutWPtr, lutW2, off, idx, hasReduce, hasDamage);
    else if (corrected) run<true, true, false, MASKED>(s, prm, md, matid, lutPtr, lut2, lutWPtr, lutW2, off, idx, hasReduce, hasDamage);
    else if (filter) run<true, false, true, MASKED>(s, prm, md, matid, lutPtr, lut2, lutWPtr, lutW2, off, idx, hasReduce, hasDamage);
    else run<true, false, false, MASKED>(s, prm, md, matid, lutPtr, lut2, lutWPtr, lutW2, off, idx, hasReduce, hasDamage);
    return 0;
}

extern "C" int hostcheck_integrate(orc_state* s, const sphgpu_config* cfg, const sphgpu_material* mats, uint32_t nmat,
    const uint64_t* off, const uint32_t* idx, int masked) {
    return masked ? dispatch<true>(s, cfg, mats, nmat, off, idx) : dispatch<false>(s, cfg, mats, nmat, off, idx);
}

/// The product's artificial-stress tensor function (sph_math.cuh: avStressOf) for n tensors {xx,yy,zz,xy,xz,yz}.
extern "C" void hostcheck_av_stress(uint32_t n, const double* sigma, double* as) {
    for (uint32_t i = 0; i < n; ++i) {
        avStressOf(sigma + 6 * (size_t)i, as + 6 * (size_t)i);
    }
}

// ---- self-gravity arithmetic (opensph_b200/csrc/grav_math.cuh) ------------------------------------------------------
#include "../../opensph_b200/csrc/grav_math.cuh"

/// Moments of the particle set formed the way the device tree forms them: the set is cut into `pieces` consecutive
/// parts, each summed directly about its own centre of mass (k_grav_moments: leaf nodes), then merged pairwise with the
/// raw-moment shift (inner nodes). Outputs the centre of mass, {M0, Q2[5], Q3[7]} and the acceleration at every probe.
extern "C" int hostcheck_gravity_moments(uint32_t n, const double* pos, const double* mass, uint32_t pieces, double* com, double* mom,
    uint32_t nProbes, const double* probes, int order, double* probeAcc) {
    struct Part {
        double c[3], m;
        GravRaw raw;
    };
    std::vector<Part> parts;
    for (uint32_t p = 0; p < pieces; ++p) {
        const uint32_t b = (uint64_t)n * p / pieces, e = (uint64_t)n * (p + 1) / pieces;
        if (b == e) continue;
        Part q{};
        for (uint32_t i = b; i < e; ++i) {
            for (int k = 0; k < 3; ++k) q.c[k] += mass[i] * pos[4 * i + k];
            q.m += mass[i];
        }
        for (int k = 0; k < 3; ++k) q.c[k] /= q.m;
        gravRawZero(q.raw);
        for (uint32_t i = b; i < e; ++i) {
            gravRawAddPoint(q.raw, mass[i], pos[4 * i] - q.c[0], pos[4 * i + 1] - q.c[1], pos[4 * i + 2] - q.c[2]);
        }
        parts.push_back(q);
    }
    while (parts.size() > 1) {
        std::vector<Part> next;
        for (size_t k = 0; k + 1 < parts.size(); k += 2) {
            const Part &a = parts[k], &b = parts[k + 1];
            Part s{};
            s.m = a.m + b.m;
            for (int d = 0; d < 3; ++d) s.c[d] = (a.m * a.c[d] + b.m * b.c[d]) / s.m;
            gravRawZero(s.raw);
            gravRawAddShifted(s.raw, a.raw, a.m, a.c[0] - s.c[0], a.c[1] - s.c[1], a.c[2] - s.c[2]);
            gravRawAddShifted(s.raw, b.raw, b.m, b.c[0] - s.c[0], b.c[1] - s.c[1], b.c[2] - s.c[2]);
            next.push_back(s);
        }
        if (parts.size() & 1) next.push_back(parts.back());
        parts.swap(next);
    }
    GravNode node;
    node.cx = parts[0].c[0]; node.cy = parts[0].c[1]; node.cz = parts[0].c[2]; node.m = parts[0].m;
    gravReduce(parts[0].raw, node);
    com[0] = node.cx; com[1] = node.cy; com[2] = node.cz;
    mom[0] = node.m;
    for (int k = 0; k < 5; ++k) mom[1 + k] = node.q2[k];
    for (int k = 0; k < 7; ++k) mom[6 + k] = node.q3[k];
    for (uint32_t p = 0; p < nProbes; ++p) {
        double ax = 0., ay = 0., az = 0.;
        if (order == 0) gravNodeAccel<0>(node, probes[3 * p], probes[3 * p + 1], probes[3 * p + 2], ax, ay, az);
        else if (order == 2) gravNodeAccel<2>(node, probes[3 * p], probes[3 * p + 1], probes[3 * p + 2], ax, ay, az);
        else gravNodeAccel<3>(node, probes[3 * p], probes[3 * p + 1], probes[3 * p + 2], ax, ay, az);
        probeAcc[3 * p] = ax; probeAcc[3 * p + 1] = ay; probeAcc[3 * p + 2] = az;
    }
    return 0;
}

/// All pairs with gravPairAccel (the walk's exact branch): acc [n*3]; masses are multiplied by G like k_grav_gather does.
extern "C" int hostcheck_gravity_pairs(uint32_t n, const double* pos, const double* mass, double G, const double* lutGrad, uint32_t entries,
    double radius, double* acc) {
    GravParams prm{};
    prm.radiusSqr = radius * radius;
    prm.qSqrToIdx = radius > 0. ? (double)entries / prm.radiusSqr : 0.;
    std::vector<LutPair> pairs(entries + 1);
    if (radius > 0.) {
        for (uint32_t k = 0; k <= entries; ++k) {
            const double next = k < entries ? lutGrad[k + 1] : lutGrad[k];
            pairs[k].g = lutGrad[k];
            pairs[k].dg = next - lutGrad[k];
        }
    }
    for (uint32_t i = 0; i < n; ++i) {
        double ax = 0., ay = 0., az = 0.;
        for (uint32_t j = 0; j < n; ++j) {
            if (j != i) {
                gravPairAccel(prm, pairs.data(), pos[4 * i], pos[4 * i + 1], pos[4 * i + 2], pos[4 * i + 3], pos[4 * j], pos[4 * j + 1],
                    pos[4 * j + 2], pos[4 * j + 3], G * mass[j], ax, ay, az);
            }
        }
        acc[3 * i] = ax; acc[3 * i + 1] = ay; acc[3 * i + 2] = az;
    }
    return 0;
}
