/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of OpenSPH's per-step SPH evaluation (see sph_oracle.h).
 * Every function cites the reference file:line it follows (paths relative to the reference root, core/...).
 * Arithmetic is written in the reference's operation order so that differences against the compiled
 * reference (oracle/_ref) are pure rounding noise. Compiled with -fno-fast-math -ffp-contract=off.
 */
#include "sph_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EPS ((double)1.e-12f)  /* math/MathUtils.h:26-30: EPS is a float literal */
#define ORC_LARGE ((double)1.e20f) /* math/MathUtils.h:33 */
#define ORC_INFTY 1.7976931348623157e308

static double sqr(double x) {
    return x * x;
}
static double dmin(double a, double b) {
    return a < b ? a : b;
}
static double dmax(double a, double b) {
    return a > b ? a : b;
}
static double clampd(double v, double lo, double hi) {
    return dmax(lo, dmin(v, hi)); /* objects/wrappers/Interval.h: Interval::clamp */
}

/* ---- kernel --------------------------------------------------------------------------------------------- */

/* CubicSpline<3>::valueImpl / gradImpl, core/sph/kernel/Kernel.h:160-187 (normalization 1/pi). */
static double cubic_value(double qSqr) {
    const double norm = 1. / M_PI;
    const double q = sqrt(qSqr);
    if (q < 1.) {
        return norm * (0.25 * (2. - q) * (2. - q) * (2. - q) - (1. - q) * (1. - q) * (1. - q));
    }
    if (q < 2.) {
        return norm * (0.25 * (2. - q) * (2. - q) * (2. - q));
    }
    return 0.;
}
static double cubic_grad(double qSqr) {
    const double norm = 1. / M_PI;
    const double q = sqrt(qSqr);
    if (q == 0.) {
        return -3. * norm;
    }
    if (q < 1.) {
        return (1. / q) * norm * (-0.75 * sqr(2. - q) + 3. * sqr(1. - q));
    }
    if (q < 2.) {
        return (1. / q) * norm * (-0.75 * sqr(2. - q));
    }
    return 0.;
}

/* LutKernel constructor, core/sph/kernel/Kernel.h:85-101 */
void orc_build_lut(double* grad, double* value, uint32_t entries, double radius) {
    const double radInvSqr = 1. / (radius * radius);
    const double qSqrToIdx = (double)entries * radInvSqr;
    for (uint32_t i = 0; i < entries + 1; ++i) {
        const double qSqr = (double)i / qSqrToIdx;
        if (value) {
            value[i] = cubic_value(qSqr);
        }
        if (grad) {
            grad[i] = cubic_grad(qSqr);
        }
    }
}

/* LutKernel::gradImpl, core/sph/kernel/Kernel.h:129-144 */
static double lut_grad(const sphgpu_config* cfg, double qSqr) {
    const double rad = cfg->kernel_radius;
    if (qSqr >= sqr(rad)) {
        return 0.;
    }
    const double qSqrToIdx = (double)cfg->lut_entries * (1. / (rad * rad));
    const double floatIdx = qSqrToIdx * qSqr;
    const uint32_t idx1 = (uint32_t)floatIdx;
    const uint32_t idx2 = idx1 + 1;
    const double ratio = floatIdx - (double)idx1;
    return cfg->lut_grad[idx1] * (1. - ratio) + cfg->lut_grad[idx2] * ratio;
}

/* Kernel::grad + SymmetrizeSmoothingLengths::grad, core/sph/kernel/Kernel.h:32-36,640-643 */
static void kernel_grad(const sphgpu_config* cfg, const double* ri, const double* rj, double g[3]) {
    const double h = 0.5 * (ri[3] + rj[3]);
    const double d[3] = { ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2] };
    const double hInv = 1. / h;
    const double hInv2 = hInv * hInv;
    const double hInv5 = hInv2 * hInv2 * hInv; /* pow<5> */
    const double G = lut_grad(cfg, (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * sqr(hInv));
    for (int k = 0; k < 3; ++k) {
        g[k] = d[k] * hInv5 * G;
    }
}

/* LutKernel::valueImpl (Kernel.h:111-127) and Kernel::value with symmetrised smoothing lengths (Kernel.h:26-30,635-638) */
static double kernel_value(const sphgpu_config* cfg, const double* ri, const double* rj) {
    const double h = 0.5 * (ri[3] + rj[3]);
    const double d[3] = { ri[0] - rj[0], ri[1] - rj[1], ri[2] - rj[2] };
    const double hInv = 1. / h;
    const double qSqr = (d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * sqr(hInv);
    const double rad = cfg->kernel_radius;
    if (qSqr >= sqr(rad)) {
        return 0.;
    }
    const double floatIdx = (double)cfg->lut_entries * (1. / (rad * rad)) * qSqr;
    const uint32_t idx1 = (uint32_t)floatIdx;
    const double ratio = floatIdx - (double)idx1;
    return hInv * hInv * hInv * (cfg->lut_value[idx1] * (1. - ratio) + cfg->lut_value[idx1 + 1] * ratio);
}

/* ---- neighbour search ------------------------------------------------------------------------------------ */

typedef struct {
    int dim[3];
    double lo[3], cell;
    uint32_t* start; /* [ncell+1] */
    uint32_t* items; /* [n] */
} orc_grid;

/* Uniform grid in the spirit of UniformGridFinder / LookupMap (core/objects/finders/UniformGrid.cpp:13-79,
 * core/objects/containers/LookupMap.h:31-97): the number of cells per axis is capped at cbrt(N)+1. */
static void grid_build(orc_grid* g, const orc_state* s, double cell) {
    const uint32_t n = s->n;
    double lo[3] = { ORC_INFTY, ORC_INFTY, ORC_INFTY }, hi[3] = { -ORC_INFTY, -ORC_INFTY, -ORC_INFTY };
    for (uint32_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = dmin(lo[k], s->pos[4 * i + k]);
            hi[k] = dmax(hi[k], s->pos[4 * i + k]);
        }
    }
    const int cap = (int)cbrt((double)n) + 1;
    double ext = 0.;
    for (int k = 0; k < 3; ++k) {
        ext = dmax(ext, hi[k] - lo[k]);
    }
    if (ext / cell > cap) {
        cell = ext / cap;
    }
    size_t ncell = 1;
    for (int k = 0; k < 3; ++k) {
        g->lo[k] = lo[k];
        g->dim[k] = (int)floor((hi[k] - lo[k]) / cell) + 1;
        ncell *= (size_t)g->dim[k];
    }
    g->cell = cell;
    g->start = (uint32_t*)calloc(ncell + 1, sizeof(uint32_t));
    g->items = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t* cellOf = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (uint32_t i = 0; i < n; ++i) {
        int c[3];
        for (int k = 0; k < 3; ++k) {
            c[k] = (int)floor((s->pos[4 * i + k] - lo[k]) / cell);
            c[k] = c[k] < 0 ? 0 : (c[k] >= g->dim[k] ? g->dim[k] - 1 : c[k]);
        }
        cellOf[i] = (uint32_t)((c[0] * g->dim[1] + c[1]) * g->dim[2] + c[2]);
        g->start[cellOf[i] + 1]++;
    }
    for (size_t c = 0; c < ncell; ++c) {
        g->start[c + 1] += g->start[c];
    }
    uint32_t* cursor = (uint32_t*)malloc(sizeof(uint32_t) * (ncell + 1));
    memcpy(cursor, g->start, sizeof(uint32_t) * (ncell + 1));
    for (uint32_t i = 0; i < n; ++i) {
        g->items[cursor[cellOf[i]]++] = i; /* ascending particle index inside each cell */
    }
    free(cursor);
    free(cellOf);
}

static void grid_free(orc_grid* g) {
    free(g->start);
    free(g->items);
}

static int cmp_u32(const void* a, const void* b) {
    const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* Neighbours of i: finder->findAll(i, 0.5*(h_i*R + maxRadius)) followed by the filter
 * `i == j || distanceSqr >= sqr(R * hbar)`, core/sph/solvers/AsymmetricSolver.cpp:177-191; the finder's own
 * test is `distSqr < sqr(radius)` (core/objects/finders/UniformGrid.cpp:69-71). Returns the count, writes
 * ascending indices into out (capacity cap; entries beyond cap are counted but not stored). */
static uint32_t neighbours_of(const orc_state* s, const sphgpu_config* cfg, const orc_grid* g, double maxRadius,
    uint32_t i, uint32_t* out, uint32_t cap) {
    const double* ri = s->pos + 4 * (size_t)i;
    const double R = cfg->kernel_radius;
    const double radius = 0.5 * (ri[3] * R + maxRadius);
    int lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = (int)floor((ri[k] - radius - g->lo[k]) / g->cell);
        hi[k] = (int)floor((ri[k] + radius - g->lo[k]) / g->cell);
        lo[k] = lo[k] < 0 ? 0 : lo[k];
        hi[k] = hi[k] >= g->dim[k] ? g->dim[k] - 1 : hi[k];
    }
    uint32_t cnt = 0;
    for (int x = lo[0]; x <= hi[0]; ++x) {
        for (int y = lo[1]; y <= hi[1]; ++y) {
            for (int z = lo[2]; z <= hi[2]; ++z) {
                const size_t c = ((size_t)x * g->dim[1] + y) * g->dim[2] + z;
                for (uint32_t k = g->start[c]; k < g->start[c + 1]; ++k) {
                    const uint32_t j = g->items[k];
                    const double* rj = s->pos + 4 * (size_t)j;
                    const double dx = rj[0] - ri[0], dy = rj[1] - ri[1], dz = rj[2] - ri[2];
                    const double distSqr = dx * dx + dy * dy + dz * dz;
                    if (!(distSqr < sqr(radius))) {
                        continue;
                    }
                    const double hbar = 0.5 * (ri[3] + rj[3]);
                    if (i == j || distSqr >= sqr(R * hbar)) {
                        continue;
                    }
                    if (cnt < cap) {
                        out[cnt] = j;
                    }
                    cnt++;
                }
            }
        }
    }
    if (cnt <= cap) {
        qsort(out, cnt, sizeof(uint32_t), cmp_u32);
    }
    return cnt;
}

/* IAsymmetricSolver::getMaxSearchRadius, core/sph/solvers/AsymmetricSolver.cpp:104-111 */
static double max_search_radius(const orc_state* s, const sphgpu_config* cfg) {
    double maxH = 0.;
    for (uint32_t i = 0; i < s->n; ++i) {
        maxH = dmax(maxH, s->pos[4 * (size_t)i + 3]);
    }
    return maxH * cfg->kernel_radius;
}

uint64_t orc_find_neighbours(const orc_state* s, const sphgpu_config* cfg, uint64_t* offsets, uint32_t* idx,
    uint64_t capacity) {
    const double maxRadius = max_search_radius(s, cfg);
    orc_grid g;
    grid_build(&g, s, maxRadius);
    uint64_t total = 0;
    uint32_t* tmp = (uint32_t*)malloc(sizeof(uint32_t) * (s->n ? s->n : 1));
    for (uint32_t i = 0; i < s->n; ++i) {
        offsets[i] = total;
        const uint32_t cnt = neighbours_of(s, cfg, &g, maxRadius, i, tmp, s->n);
        if (idx && total + cnt <= capacity) {
            memcpy(idx + total, tmp, sizeof(uint32_t) * cnt);
        }
        total += cnt;
    }
    offsets[s->n] = total;
    free(tmp);
    grid_free(&g);
    return total;
}

/* ---- materials -------------------------------------------------------------------------------------------- */

/* TillotsonEos::evaluate, core/physics/Eos.cpp:198-238 */
static void eos_tillotson(const sphgpu_material* m, double rho, double u, double* pOut, double* csOut) {
    const double u0 = m->til_u0, uiv = m->til_uiv, ucv = m->til_ucv, a = m->til_a, b = m->til_b, rho0 = m->rho0,
                 A = m->til_A, B = m->til_B, alpha = m->til_alpha, beta = m->til_beta;
    const double eta = rho / rho0;
    const double mu = eta - 1.;
    const double denom = u / (u0 * eta * eta) + 1.;
    const double pc = (a + b / denom) * rho * u + A * mu + B * mu * mu;
    double dpdu = a * rho + b * rho / sqr(denom);
    double dpdrho = a * u + b * u * (3. * denom - 2.) / sqr(denom) + A / rho0 + 2. * B * mu / rho0;
    const double csc = dpdrho + dpdu * pc / (rho * rho);

    const double rhoExp = rho0 / rho - 1.;
    const double betaExp = exp(-dmin(beta * rhoExp, 70.));
    const double alphaExp = exp(-dmin(alpha * sqr(rhoExp), 70.));
    const double pe = a * rho * u + (b * rho * u / denom + A * mu * betaExp) * alphaExp;
    dpdu = a * rho + alphaExp * b * rho / sqr(denom);
    dpdrho = a * u + alphaExp * (b * u * (3. * denom - 2.) / sqr(denom)) +
             alphaExp * (b * u * rho / denom) * rho0 * (2. * alpha * rhoExp) / sqr(rho) +
             alphaExp * A * betaExp * (1. / rho0 + rho0 * mu / sqr(rho) * (2. * alpha * rhoExp + beta));
    double cse = dpdrho + dpdu * pe / (rho * rho);
    cse = dmax(cse, 0.);

    double p = pc, cs = csc;
    if (rho <= rho0 && u > ucv) {
        p = pe;
        cs = cse;
    } else if (rho <= rho0 && u > uiv && u <= ucv) {
        p = ((u - uiv) * pe + (ucv - u) * pc) / (ucv - uiv);
        cs = ((u - uiv) * cse + (ucv - u) * csc) / (ucv - uiv);
    }
    cs = dmax(cs, 0.25 * A / rho0);
    *pOut = p;
    *csOut = sqrt(cs);
}

/* IdealGasEos::evaluate, core/physics/Eos.cpp:42-45 */
static void eos_ideal_gas(const sphgpu_material* m, double rho, double u, double* p, double* cs) {
    *p = (m->gamma - 1.) * u * rho;
    *cs = sqrt(m->gamma * *p / rho);
}

/* material->initialize for every material: EosMaterial::initialize (core/sph/Materials.cpp:47-58) then
 * VonMisesRheology::initialize (core/physics/Rheology.cpp:36-83). */
static void materials_initialize(orc_state* s, const sphgpu_material* mats, uint32_t nmat) {
    for (uint32_t mi = 0; mi < nmat; ++mi) {
        const sphgpu_material* m = &mats[mi];
        for (uint32_t i = m->begin; i < m->end; ++i) {
            if (m->eos == SPHGPU_EOS_TILLOTSON) {
                eos_tillotson(m, s->rho[i], s->u[i], &s->p[i], &s->cs[i]);
            } else if (m->eos == SPHGPU_EOS_IDEAL_GAS) {
                eos_ideal_gas(m, s->rho[i], s->u[i], &s->p[i], &s->cs[i]);
            }
        }
        if (m->yielding != SPHGPU_YIELD_VON_MISES) {
            continue;
        }
        const double limit = m->elasticity_limit;
        const double u_melt = m->melt_energy;
        const double eps = 1.e-15;
        const int hasD = (s->damage != NULL) && m->fracture != SPHGPU_FRACTURE_NONE;
        for (uint32_t i = m->begin; i < m->end; ++i) {
            double* S = s->S + 5 * (size_t)i;
            const double d = hasD ? s->damage[i] * s->damage[i] * s->damage[i] : 0.;
            if (s->p[i] < 0.) {
                s->p[i] = (1. - d) * s->p[i];
            }
            const double unorm = s->u[i] / u_melt;
            double Y = unorm < 1.e-5 ? limit : limit * dmax(1. - unorm, 0.);
            Y = (1. - d) * Y;
            if (Y < ORC_EPS) {
                s->reduce[i] = 0.;
                memset(S, 0, 5 * sizeof(double));
                continue;
            }
            /* ddot(S,S) = dot(diag,diag) + 2 dot(off,off), objects/geometry/TracelessTensor.h:388-390 */
            const double szz = -S[0] - S[1];
            const double ddot = (S[0] * S[0] + S[1] * S[1] + szz * szz) + 2. * (S[2] * S[2] + S[3] * S[3] + S[4] * S[4]);
            const double J2 = 0.5 * ddot + eps;
            const double red = dmin(Y / sqrt(3. * J2), 1.);
            s->reduce[i] = red;
            for (int k = 0; k < 5; ++k) {
                S[k] = S[k] * red;
            }
        }
    }
}

/* sqrtInv / sqrtApprox, core/math/MathUtils.h:40-62 (float bit trick, one Newton iteration) */
static double sqrt_approx(double f) {
    if (f == 0.) {
        return 0.;
    }
    int i;
    float x2 = (float)f * 0.5f;
    float y = (float)f;
    memcpy(&i, &y, sizeof(float));
    i = 0x5f3759df - (i >> 1);
    memcpy(&y, &i, sizeof(float));
    const float r = y * (1.5f - (x2 * y * y));
    return (double)(1.f / r);
}

/* findEigenvalues, core/objects/geometry/SymmetricTensor.h:377-399; t = {xx,yy,zz,xy,xz,yz} */
static void find_eigenvalues(const double t[6], double sig[3]) {
    /* norm(SymmetricTensor) = norm(max(diag, offdiag)) with the approximate sqrt, SymmetricTensor.h:290-294 */
    const double v0 = dmax(t[0], t[3]), v1 = dmax(t[1], t[4]), v2 = dmax(t[2], t[5]);
    const double n = sqrt_approx(v0 * v0 + v1 * v1 + v2 * v2);
    sig[0] = sig[1] = sig[2] = 0.;
    if (n < 1.e-12) {
        return;
    }
    const double inv1 = t[0] + t[1] + t[2];
    const double inv2 = (t[3] * t[3] + t[4] * t[4] + t[5] * t[5]) - (t[1] * t[2] + t[2] * t[0] + t[0] * t[1]);
    /* determinant, SymmetricTensor.h:193-196: dot(sqr(off), (zz,yy,xx)) */
    const double inv3 = t[0] * t[1] * t[2] + 2. * t[3] * t[4] * t[5] - (t[3] * t[3] * t[2] + t[4] * t[4] * t[1] + t[5] * t[5] * t[0]);
    const double p = -inv1 / n;
    const double q = -inv2 / sqr(n);
    const double r = -inv3 / (n * n * n);
    const double a = q - p * p / 3.;
    const double b = (2. * p * p * p - 9. * p * q + 27. * r) / 27.;
    const double aCub = a * a * a / 27.;
    if (0.25 * b * b + aCub >= 0.) {
        return;
    }
    const double t1 = 2. * sqrt(-a / 3.);
    const double phi = acos(-0.5 * b / sqrt(-aCub));
    const double ang[3] = { phi / 3., (phi + 2 * M_PI) / 3., (phi + 4 * M_PI) / 3. };
    for (int k = 0; k < 3; ++k) {
        sig[k] = (t1 * cos(ang[k]) - p / 3.) * n;
    }
}

/* material->finalize: ScalarGradyKippModel::integrate, core/physics/Damage.cpp:128-170 */
static void materials_finalize(orc_state* s, const sphgpu_material* mats, uint32_t nmat) {
    for (uint32_t mi = 0; mi < nmat; ++mi) {
        const sphgpu_material* m = &mats[mi];
        if (m->yielding != SPHGPU_YIELD_VON_MISES || m->fracture != SPHGPU_FRACTURE_SCALAR_GRADY_KIPP) {
            continue;
        }
        for (uint32_t i = m->begin; i < m->end; ++i) {
            if (s->damage[i] >= m->d_max) {
                s->ddamage[i] = ORC_LARGE;
                continue;
            }
            const double* S = s->S + 5 * (size_t)i;
            const double sigma[6] = { S[0] - s->p[i], S[1] - s->p[i], (-S[0] - S[1]) - s->p[i], S[2], S[3], S[4] };
            double sig[3];
            find_eigenvalues(sigma, sig);
            const double sigMax = dmax(dmax(sig[0], sig[1]), sig[2]);
            const double D = s->damage[i];
            const double young_red = dmax((1. - D * D * D) * m->young_modulus, 1.e-20);
            const double strain = sigMax / young_red;
            const double ratio = strain / s->eps_min[i];
            if (ratio <= 1.) {
                continue;
            }
            s->ddamage[i] = s->growth[i] * cbrt(dmin(pow(ratio, s->m_zero[i]), (double)s->n_flaws[i]));
        }
    }
}

/* ---- derivatives ------------------------------------------------------------------------------------------ */

/* StressAV::initialize for one particle (core/sph/equations/av/Stress.cpp:91-109): sigma = S - p I is diagonalised, the positive
 * principal stresses are negated, the negative ones dropped, and the result is rotated back: as = -V max(Lambda, 0) V^T, i.e.
 * minus the positive part of sigma -- a function of the tensor, so any convergent symmetric eigen-solver gives the same result
 * to rounding. The reference uses JAMA's tred2 / tql2 (SymmetricTensor.cpp:110-272); this restatement uses cyclic Jacobi
 * rotations and is pinned on the golden vectors of the reference run. sigma, as: {xx,yy,zz,xy,xz,yz}. */
static void av_stress_of(const double sigma[6], double as[6]) {
    double A[3][3] = { { sigma[0], sigma[3], sigma[4] }, { sigma[3], sigma[1], sigma[5] }, { sigma[4], sigma[5], sigma[2] } };
    double V[3][3] = { { 1., 0., 0. }, { 0., 1., 0. }, { 0., 0., 1. } };
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off == 0.) {
            break;
        }
        for (int p = 0; p < 2; ++p) {
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.) {
                    continue;
                }
                const double theta = (A[q][q] - A[p][p]) / (2. * A[p][q]);
                const double t = (theta >= 0. ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1.));
                const double c = 1. / sqrt(t * t + 1.), sn = t * c;
                const int r = 3 - p - q;
                const double app = A[p][p], aqq = A[q][q], apq = A[p][q], arp = A[r][p], arq = A[r][q];
                A[p][p] = app - t * apq;
                A[q][q] = aqq + t * apq;
                A[p][q] = A[q][p] = 0.;
                A[r][p] = A[p][r] = c * arp - sn * arq;
                A[r][q] = A[q][r] = sn * arp + c * arq;
                for (int k = 0; k < 3; ++k) {
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = c * vp - sn * vq;
                    V[k][q] = sn * vp + c * vq;
                }
            }
        }
    }
    static const int IX[6][2] = { { 0, 0 }, { 1, 1 }, { 2, 2 }, { 0, 1 }, { 0, 2 }, { 1, 2 } };
    for (int c = 0; c < 6; ++c) {
        double v = 0.;
        for (int k = 0; k < 3; ++k) {
            const double lam = A[k][k] > 0. ? A[k][k] : 0.;
            v -= lam * V[IX[c][0]][k] * V[IX[c][1]][k];
        }
        as[c] = v;
    }
}

/* DerivativeTemplate::sum / AccelerationTemplate::sum filter, core/sph/equations/DerivativeHelpers.h:132-139 */
static int undamaged_skip(const orc_state* s, uint32_t i, uint32_t j) {
    return s->flag[i] != s->flag[j] || s->reduce[i] == 0. || s->reduce[j] == 0.;
}

void orc_integrate(orc_state* s, const sphgpu_config* cfg, const sphgpu_material* mats, uint32_t nmat) {
    const uint32_t n = s->n;
    const int solid = (cfg->forces & SPHGPU_FORCE_SOLID_STRESS) != 0;
    const int hasReduce = s->reduce != NULL;
    const int filter = (cfg->flags & SPHGPU_FLAG_SUM_ONLY_UNDAMAGED) && hasReduce;
    const int corrected = solid && (cfg->flags & SPHGPU_FLAG_CORRECTION_TENSOR);
    const int adaptive = (cfg->flags & SPHGPU_FLAG_ADAPTIVE_H) != 0;

    /* IAsymmetricSolver::integrate, AsymmetricSolver.cpp:71-96 */
    materials_initialize(s, mats, nmat);

    /* beforeLoop -> equations.initialize: AdaptiveSmoothingLength::initialize clamps h, EquationTerm.cpp:356-364 */
    if (adaptive) {
        for (uint32_t i = 0; i < n; ++i) {
            s->pos[4 * (size_t)i + 3] = clampd(s->pos[4 * (size_t)i + 3], cfg->h_min, cfg->h_max);
        }
    }

    /* XSph::initialize (XSph.h:69-79; the first term of getStandardEquations): remove the previous correction */
    const int xsph = (cfg->flags & SPHGPU_FLAG_XSPH) != 0;
    if (xsph) {
        for (uint32_t i = 0; i < n; ++i) {
            for (int q = 0; q < 3; ++q) {
                s->vel[4 * (size_t)i + q] -= s->xsph[4 * (size_t)i + q];
            }
        }
    }

    /* StressAV::initialize (Stress.cpp:91-109) on the pressure and stress the materials just left in the Storage */
    const int stressAv = (cfg->flags & SPHGPU_FLAG_STRESS_AV) != 0;
    if (stressAv) {
        for (uint32_t i = 0; i < n; ++i) {
            const double* S = s->S + 5 * (size_t)i;
            const double sigma[6] = { S[0] - s->p[i], S[1] - s->p[i], (-S[0] - S[1]) - s->p[i], S[2], S[3], S[4] };
            av_stress_of(sigma, s->av_stress + 6 * (size_t)i);
        }
    }

    /* DeltaSph terms (core/sph/equations/DeltaSph.h; StandardSets.cpp:64-67). The new density gradient goes to the
     * Accumulated buffer while DensityDiffusion reads the Storage values of the previous evaluation (DeltaSph.h:71-74),
     * hence the separate array. */
    const int deltasph = (cfg->flags & SPHGPU_FLAG_DELTASPH) != 0;
    double* newGrad = deltasph ? (double*)calloc(4 * (size_t)n + 4, sizeof(double)) : NULL;

    /* Accumulated buffers start zeroed (Accumulated::initialize, core/sph/equations/Accumulated.cpp:40-60) */
    for (uint32_t i = 0; i < n; ++i) {
        s->acc[4 * (size_t)i] = s->acc[4 * (size_t)i + 1] = s->acc[4 * (size_t)i + 2] = s->acc[4 * (size_t)i + 3] = 0.;
        s->du[i] = 0.;
        s->drho[i] = 0.; /* storage derivative zeroed by Storage::zeroHighestDerivatives, Storage.cpp:594-602 */
        s->divv[i] = 0.;
        if (solid) {
            memset(s->gradv + 6 * (size_t)i, 0, 6 * sizeof(double));
            memset(s->dS + 5 * (size_t)i, 0, 5 * sizeof(double));
        }
        if (s->ddamage) {
            s->ddamage[i] = 0.;
        }
    }

    /* AsymmetricSolver::loop, AsymmetricSolver.cpp:153-202 */
    const double maxRadius = max_search_radius(s, cfg);
    orc_grid g;
    grid_build(&g, s, maxRadius);
    const double alpha = cfg->av_alpha, beta = cfg->av_beta, avEps = 1.e-2;

#pragma omp parallel
    {
        uint32_t cap = 1024;
        uint32_t* neighs = (uint32_t*)malloc(sizeof(uint32_t) * cap);
        double* grads = (double*)malloc(sizeof(double) * 3 * cap);
#pragma omp for schedule(dynamic, 256)
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t cnt = neighbours_of(s, cfg, &g, maxRadius, i, neighs, cap);
            if (cnt > cap) {
                cap = cnt * 2;
                neighs = (uint32_t*)realloc(neighs, sizeof(uint32_t) * cap);
                grads = (double*)realloc(grads, sizeof(double) * 3 * cap);
                cnt = neighbours_of(s, cfg, &g, maxRadius, i, neighs, cap);
            }
            const double* ri = s->pos + 4 * (size_t)i;
            const double* vi = s->vel + 4 * (size_t)i;
            for (uint32_t k = 0; k < cnt; ++k) {
                kernel_grad(cfg, ri, s->pos + 4 * (size_t)neighs[k], grads + 3 * k);
            }
            /* derivatives.eval(i, idxs, grads): PRECOMPUTE phase first, Derivative.cpp:113-118, Derivative.h:177-186 */
            double C[6] = { 1., 1., 1., 0., 0., 0. };
            if (corrected) {
                /* CorrectionTensor::evalNeighs, Derivative.cpp:36-74 */
                double c[6] = { 0., 0., 0., 0., 0., 0. };
                for (uint32_t k = 0; k < cnt; ++k) {
                    const uint32_t j = neighs[k];
                    if (filter && undamaged_skip(s, i, j)) {
                        continue;
                    }
                    const double* rj = s->pos + 4 * (size_t)j;
                    const double* gr = grads + 3 * k;
                    const double a[3] = { rj[0] - ri[0], rj[1] - ri[1], rj[2] - ri[2] };
                    /* symmetricOuter, SymmetricTensor.h:362-368 */
                    const double t[6] = { a[0] * gr[0], a[1] * gr[1], a[2] * gr[2], 0.5 * (a[0] * gr[1] + a[1] * gr[0]),
                        0.5 * (a[0] * gr[2] + a[2] * gr[0]), 0.5 * (a[1] * gr[2] + a[2] * gr[1]) };
                    const double w = s->mass[j] / s->rho[j];
                    for (int q = 0; q < 6; ++q) {
                        c[q] += w * t[q];
                    }
                }
                if (c[0] == 0. && c[1] == 0. && c[2] == 0. && c[3] == 0. && c[4] == 0. && c[5] == 0.) {
                    /* identity */
                } else {
                    /* determinant / inverse, SymmetricTensor.h:193-196,219-231 (off = {xy,xz,yz}) */
                    const double det = c[0] * c[1] * c[2] + 2. * c[3] * c[4] * c[5] -
                                       (c[3] * c[3] * c[2] + c[4] * c[4] * c[1] + c[5] * c[5] * c[0]);
                    if (det > 0.01) {
                        C[0] = (c[1] * c[2] - sqr(c[5])) / det;
                        C[1] = (c[2] * c[0] - sqr(c[4])) / det;
                        C[2] = (c[0] * c[1] - sqr(c[3])) / det;
                        C[3] = (c[4] * c[5] - c[2] * c[3]) / det;
                        C[4] = (c[5] * c[3] - c[1] * c[4]) / det;
                        C[5] = (c[3] * c[4] - c[0] * c[5]) / det;
                    }
                }
                memcpy(s->corr + 6 * (size_t)i, C, sizeof(C));
            }
            /* VelocityDivergence<CenterDensityDiscr> (no flags), DerivativeHelpers.h:292-304,355-371 */
            double divv = 0.;
            for (uint32_t k = 0; k < cnt; ++k) {
                const uint32_t j = neighs[k];
                const double* vj = s->vel + 4 * (size_t)j;
                const double* gr = grads + 3 * k;
                const double dvg = (vj[0] - vi[0]) * gr[0] + (vj[1] - vi[1]) * gr[1] + (vj[2] - vi[2]) * gr[2];
                divv += s->mass[j] / s->rho[i] * dvg;
            }
            s->divv[i] = divv;
            /* VelocityGradient<CenterDensityDiscr> with SUM_ONLY_UNDAMAGED | CORRECTED, DerivativeHelpers.h:100-146,382-398 */
            if (solid) {
                double gv[6] = { 0., 0., 0., 0., 0., 0. };
                for (uint32_t k = 0; k < cnt; ++k) {
                    const uint32_t j = neighs[k];
                    if (filter && undamaged_skip(s, i, j)) {
                        continue;
                    }
                    const double* vj = s->vel + 4 * (size_t)j;
                    const double* g0 = grads + 3 * k;
                    double gr[3] = { g0[0], g0[1], g0[2] };
                    if (corrected) {
                        /* C[i] * grad, SymmetricTensor.h:102-107 */
                        gr[0] = C[0] * g0[0] + C[3] * g0[1] + C[4] * g0[2];
                        gr[1] = C[3] * g0[0] + C[1] * g0[1] + C[5] * g0[2];
                        gr[2] = C[4] * g0[0] + C[5] * g0[1] + C[2] * g0[2];
                    }
                    const double a[3] = { vj[0] - vi[0], vj[1] - vi[1], vj[2] - vi[2] };
                    const double t[6] = { a[0] * gr[0], a[1] * gr[1], a[2] * gr[2], 0.5 * (a[0] * gr[1] + a[1] * gr[0]),
                        0.5 * (a[0] * gr[2] + a[2] * gr[0]), 0.5 * (a[1] * gr[2] + a[2] * gr[1]) };
                    const double w = s->mass[j] / s->rho[i];
                    for (int q = 0; q < 6; ++q) {
                        gv[q] += w * t[q];
                    }
                }
                memcpy(s->gradv + 6 * (size_t)i, gv, sizeof(gv));
            }
            double dv[3] = { 0., 0., 0. };
            double du = 0.;
            /* PressureGradient<StandardForceDiscr>, EquationTerm.cpp:12-24,41-67 */
            if (cfg->forces & SPHGPU_FORCE_PRESSURE) {
                for (uint32_t k = 0; k < cnt; ++k) {
                    const uint32_t j = neighs[k];
                    const double* gr = grads + 3 * k;
                    const double c = s->p[i] / sqr(s->rho[i]) + s->p[j] / sqr(s->rho[j]);
                    for (int q = 0; q < 3; ++q) {
                        dv[q] += s->mass[j] * (-(c * gr[q]));
                    }
                }
            }
            /* StressDivergence<StandardForceDiscr>, SUM_ONLY_UNDAMAGED, EquationTerm.cpp:113-140 */
            if (solid) {
                const double* Si = s->S + 5 * (size_t)i;
                for (uint32_t k = 0; k < cnt; ++k) {
                    const uint32_t j = neighs[k];
                    if (filter && undamaged_skip(s, i, j)) {
                        continue;
                    }
                    const double* Sj = s->S + 5 * (size_t)j;
                    const double* gr = grads + 3 * k;
                    const double ri2 = sqr(s->rho[i]), rj2 = sqr(s->rho[j]);
                    double T[5];
                    for (int q = 0; q < 5; ++q) {
                        T[q] = Si[q] / ri2 + Sj[q] / rj2;
                    }
                    /* TracelessTensor * Vector, TracelessTensor.h:159-163 */
                    const double f[3] = { T[0] * gr[0] + T[2] * gr[1] + T[3] * gr[2],
                        T[2] * gr[0] + T[1] * gr[1] + T[4] * gr[2], T[3] * gr[0] + T[4] * gr[1] + (-T[0] - T[1]) * gr[2] };
                    for (int q = 0; q < 3; ++q) {
                        dv[q] += s->mass[j] * f[q];
                    }
                }
            }
            /* StandardAV::Derivative, core/sph/equations/av/Standard.h:63-83 */
            for (uint32_t k = 0; k < cnt; ++k) {
                const uint32_t j = neighs[k];
                const double* rj = s->pos + 4 * (size_t)j;
                const double* vj = s->vel + 4 * (size_t)j;
                const double* gr = grads + 3 * k;
                const double dvx = vi[0] - vj[0], dvy = vi[1] - vj[1], dvz = vi[2] - vj[2];
                const double dx = ri[0] - rj[0], dy = ri[1] - rj[1], dz = ri[2] - rj[2];
                const double dvdr = dvx * dx + dvy * dy + dvz * dz;
                double av = 0.;
                if (!(dvdr >= 0.)) {
                    const double hbar = 0.5 * (ri[3] + rj[3]);
                    const double rhobar = 0.5 * (s->rho[i] + s->rho[j]);
                    const double csbar = 0.5 * (s->cs[i] + s->cs[j]);
                    const double mu = hbar * dvdr / ((dx * dx + dy * dy + dz * dz) + avEps * sqr(hbar));
                    av = 1. / rhobar * (-alpha * csbar * mu + beta * sqr(mu));
                }
                const double heating = 0.5 * av * (dvx * gr[0] + dvy * gr[1] + dvz * gr[2]);
                for (int q = 0; q < 3; ++q) {
                    dv[q] += s->mass[j] * (-(av * gr[q]));
                }
                du += s->mass[j] * heating;
            }
            /* XSph::Derivative::eval, XSph.h:55-63 (written to its own buffer: the velocities stay pure during the loop) */
            if (xsph) {
                double xs[3] = { 0., 0., 0. };
                for (uint32_t k = 0; k < cnt; ++k) {
                    const uint32_t j = neighs[k];
                    const double* vj = s->vel + 4 * (size_t)j;
                    const double w = kernel_value(cfg, ri, s->pos + 4 * (size_t)j);
                    for (int q = 0; q < 3; ++q) {
                        const double f = s->xsph_eps * (vj[q] - vi[q]) / (0.5 * (s->rho[i] + s->rho[j])) * w;
                        xs[q] += s->mass[j] * f;
                    }
                }
                for (int q = 0; q < 3; ++q) {
                    s->xsph[4 * (size_t)i + q] = xs[q];
                }
                s->xsph[4 * (size_t)i + 3] = 0.;
            }
            /* StressAV::Derivative::eval with the STANDARD discretisation (Stress.cpp:44-57,68-79), SUM_ONLY_UNDAMAGED;
             * AccelerationTemplate: dv_i += m_j f, du_i += m_j heating (DerivativeHelpers.h:216-227) */
            if (stressAv) {
                const double* ai = s->av_stress + 6 * (size_t)i;
                for (uint32_t k = 0; k < cnt; ++k) {
                    const uint32_t j = neighs[k];
                    if (filter && undamaged_skip(s, i, j)) {
                        continue;
                    }
                    const double* aj = s->av_stress + 6 * (size_t)j;
                    const double* vj = s->vel + 4 * (size_t)j;
                    const double* gr = grads + 3 * k;
                    const double w = kernel_value(cfg, ri, s->pos + 4 * (size_t)j);
                    const double phi = s->stress_av_factor * pow(w / s->wp[i], s->stress_av_exponent);
                    const double ri2 = sqr(s->rho[i]), rj2 = sqr(s->rho[j]);
                    double Pi[6];
                    for (int q = 0; q < 6; ++q) {
                        Pi[q] = phi * (ai[q] / ri2 + aj[q] / rj2);
                    }
                    const double f[3] = { Pi[0] * gr[0] + Pi[3] * gr[1] + Pi[4] * gr[2], Pi[3] * gr[0] + Pi[1] * gr[1] + Pi[5] * gr[2],
                        Pi[4] * gr[0] + Pi[5] * gr[1] + Pi[2] * gr[2] };
                    const double a[3] = { vi[0] - vj[0], vi[1] - vj[1], vi[2] - vj[2] };
                    const double Pa[3] = { Pi[0] * a[0] + Pi[3] * a[1] + Pi[4] * a[2], Pi[3] * a[0] + Pi[1] * a[1] + Pi[5] * a[2],
                        Pi[4] * a[0] + Pi[5] * a[1] + Pi[2] * a[2] };
                    const double heating = 0.5 * (Pa[0] * gr[0] + Pa[1] * gr[1] + Pa[2] * gr[2]);
                    for (int q = 0; q < 3; ++q) {
                        dv[q] += s->mass[j] * f[q];
                    }
                    du += s->mass[j] * heating;
                }
            }
            if (deltasph) {
                const double* Gi = s->drho_grad + 4 * (size_t)i;
                double G[3] = { 0., 0., 0. };
                double diff = 0.;
                for (uint32_t k = 0; k < cnt; ++k) {
                    const uint32_t j = neighs[k];
                    if (filter && undamaged_skip(s, i, j)) { /* all three derivatives: SUM_ONLY_UNDAMAGED */
                        continue;
                    }
                    const double* rj = s->pos + 4 * (size_t)j;
                    const double* vj = s->vel + 4 * (size_t)j;
                    const double* Gj = s->drho_grad + 4 * (size_t)j;
                    const double* g0 = grads + 3 * k;
                    const double vol = s->mass[j] / s->rho[j];
                    const double drh = s->rho[j] - s->rho[i];
                    /* RenormalizedDensityGradient::eval (DeltaSph.h:37-44) with C[i] * grad when CORRECTED
                     * (DerivativeHelpers.h:100-107) */
                    double gr[3] = { g0[0], g0[1], g0[2] };
                    if (corrected) {
                        const double* C = s->corr + 6 * (size_t)i;
                        gr[0] = C[0] * g0[0] + C[3] * g0[1] + C[4] * g0[2];
                        gr[1] = C[3] * g0[0] + C[1] * g0[1] + C[5] * g0[2];
                        gr[2] = C[4] * g0[0] + C[5] * g0[1] + C[2] * g0[2];
                    }
                    for (int q = 0; q < 3; ++q) {
                        G[q] += vol * (drh * gr[q]);
                    }
                    /* DensityDiffusion::Derivative::eval (DeltaSph.h:81-93) */
                    const double dr[3] = { rj[0] - ri[0], rj[1] - ri[1], rj[2] - ri[2] };
                    const double dr2 = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2];
                    const double hbar = 0.5 * (ri[3] + rj[3]);
                    const double cbar = 0.5 * (s->cs[i] + s->cs[j]);
                    double psig = 0.;
                    for (int q = 0; q < 3; ++q) {
                        const double psi = 2. * drh * dr[q] / dr2 - (Gi[q] + Gj[q]);
                        psig += psi * g0[q];
                    }
                    diff += vol * (s->deltasph_delta * hbar * cbar * psig);
                    /* VelocityDiffusion::Derivative::eval (DeltaSph.h:147-163) */
                    const double pi = ((vj[0] - vi[0]) * dr[0] + (vj[1] - vi[1]) * dr[1] + (vj[2] - vi[2]) * dr[2]) / dr2;
                    for (int q = 0; q < 3; ++q) {
                        dv[q] += vol * (s->deltasph_alpha * hbar * cbar * pi * g0[q]);
                    }
                }
                newGrad[4 * (size_t)i + 0] = G[0];
                newGrad[4 * (size_t)i + 1] = G[1];
                newGrad[4 * (size_t)i + 2] = G[2];
                s->drho[i] += diff; /* the SHARED density-derivative buffer; ContinuityEquation::finalize adds to it */
            }
            s->acc[4 * (size_t)i + 0] = dv[0];
            s->acc[4 * (size_t)i + 1] = dv[1];
            s->acc[4 * (size_t)i + 2] = dv[2];
            s->du[i] = du;
            s->ncnt[i] = cnt; /* AsymmetricSolver.cpp:199 */
        }
        free(neighs);
        free(grads);
    }
    grid_free(&g);
    if (deltasph) { /* Accumulated::store: the new gradient replaces the Storage values */
        memcpy(s->drho_grad, newGrad, sizeof(double) * 4 * (size_t)n);
        free(newGrad);
    }

    /* XSph::finalize (XSph.h:81-90): the new correction joins the velocities (the loop above read pure velocities only:
     * every thread wrote xsph of its own particle and nobody read it) */
    if (xsph) {
        for (uint32_t i = 0; i < n; ++i) {
            for (int q = 0; q < 3; ++q) {
                s->vel[4 * (size_t)i + q] += s->xsph[4 * (size_t)i + q];
            }
        }
    }
    /* afterLoop -> equations.finalize in REVERSE term order (EquationTerm.h:293-297); term order from
     * getStandardEquations (StandardSets.cpp:24-92): Pressure, SolidStress, Continuity, AV, SmoothingLength. */
    for (uint32_t i = 0; i < n; ++i) {
        double* r = s->pos + 4 * (size_t)i;
        double* v = s->vel + 4 * (size_t)i;
        double* dv = s->acc + 4 * (size_t)i;
        if (adaptive) {
            /* AdaptiveSmoothingLength::finalize + enforce, EquationTerm.cpp:366-418 */
            if (r[3] > 2. * cfg->h_min) {
                v[3] = r[3] / 3. * s->divv[i];
            } else {
                v[3] = 0.;
            }
            dv[3] = 0.;
            if (cfg->flags & SPHGPU_FLAG_SOUND_SPEED_ENFORCING) {
                const double strength = cfg->neigh_enforcing;
                if (!(strength <= -1.e2)) {
                    const double dn1 = (double)s->ncnt[i] - cfg->neigh_upper;
                    if (dn1 > 0.) {
                        v[3] -= exp(strength * dn1) * s->cs[i];
                    } else {
                        const double dn2 = cfg->neigh_lower - (double)s->ncnt[i];
                        if (dn2 > 0.) {
                            v[3] += exp(strength * dn2) * s->cs[i];
                        }
                    }
                }
            }
        } else {
            /* ConstSmoothingLength::finalize, EquationTerm.cpp:427-434 */
            v[3] = 0.;
            dv[3] = 0.;
        }
    }
    /* ContinuityEquation::finalize, EquationTerm.cpp:289-316 */
    for (uint32_t i = 0; i < n; ++i) {
        if (cfg->continuity_mode == SPHGPU_CONTINUITY_SUM_ONLY_UNDAMAGED && hasReduce && solid && s->reduce[i] > 0.) {
            const double* gv = s->gradv + 6 * (size_t)i;
            s->drho[i] += -s->rho[i] * (gv[0] + gv[1] + gv[2]);
        } else {
            s->drho[i] += -s->rho[i] * s->divv[i];
        }
    }
    /* SolidStressForce::finalize, EquationTerm.cpp:177-203 */
    if (solid) {
        for (uint32_t mi = 0; mi < nmat; ++mi) {
            const sphgpu_material* m = &mats[mi];
            if (m->yielding == SPHGPU_YIELD_NONE || m->yielding == SPHGPU_YIELD_DUST) {
                continue;
            }
            const double mu = m->shear_modulus;
            for (uint32_t i = m->begin; i < m->end; ++i) {
                const double* S = s->S + 5 * (size_t)i;
                const double* gv = s->gradv + 6 * (size_t)i;
                double* dS = s->dS + 5 * (size_t)i;
                const double ddot = (S[0] * gv[0] + S[1] * gv[1] + (-S[0] - S[1]) * gv[2]) +
                                    2. * (S[2] * gv[3] + S[3] * gv[4] + S[4] * gv[5]);
                s->du[i] += 1. / s->rho[i] * ddot;
                const double tr3 = (gv[0] + gv[1] + gv[2]) / 3.;
                const double dev[5] = { gv[0] - tr3, gv[1] - tr3, gv[3], gv[4], gv[5] };
                for (int q = 0; q < 5; ++q) {
                    dS[q] += 2. * mu * dev[q];
                }
            }
        }
    }
    /* PressureForce::finalize, EquationTerm.cpp:90-99 */
    if (cfg->forces & SPHGPU_FORCE_PRESSURE) {
        for (uint32_t i = 0; i < n; ++i) {
            s->du[i] -= s->p[i] / s->rho[i] * s->divv[i];
        }
    }
    materials_finalize(s, mats, nmat);
}

/* ---- time stepping ---------------------------------------------------------------------------------------- */

/* clampWithDerivative<Float>, core/objects/wrappers/Interval.h:159-162 */
static void clamp_with_derivative(double* v, double* dv, double lo, double hi) {
    const int zeroDeriv = (*v >= hi && *dv > 0.) || (*v <= lo && *dv < 0.);
    *v = clampd(*v, lo, hi);
    if (zeroDeriv) {
        *dv = 0.;
    }
}

static int bounded(double lo, double hi) {
    return !(lo <= -ORC_INFTY && hi >= ORC_INFTY); /* range != Interval::unbounded() */
}

typedef struct {
    double lo, hi;
} orc_range;

static void step_first_order_clamp(const sphgpu_material* m, orc_state* s, uint32_t i) {
    if (bounded(m->rho_min, m->rho_max)) {
        clamp_with_derivative(&s->rho[i], &s->drho[i], m->rho_min, m->rho_max);
    }
    if (bounded(m->u_min, m->u_max)) {
        clamp_with_derivative(&s->u[i], &s->du[i], m->u_min, m->u_max);
    }
    if (s->damage && s->ddamage && m->fracture != SPHGPU_FRACTURE_NONE && bounded(m->d_min, m->d_max)) {
        clamp_with_derivative(&s->damage[i], &s->ddamage[i], m->d_min, m->d_max);
    }
}

/* PredictorCorrector::makePredictions (TimeStepping.cpp:286-300) followed by storage->swap(predictions,
 * HIGHEST_DERIVATIVES) and zeroHighestDerivatives (TimeStepping.cpp:331-334). */
void orc_predict(orc_state* s, const sphgpu_config* cfg, const sphgpu_material* mats, uint32_t nmat, double dt) {
    (void)cfg;
    const double dt2 = 0.5 * sqr(dt);
    const int solid = s->S != NULL;
    for (uint32_t mi = 0; mi < nmat; ++mi) {
        const sphgpu_material* m = &mats[mi];
        for (uint32_t i = m->begin; i < m->end; ++i) {
            double* r = s->pos + 4 * (size_t)i;
            double* v = s->vel + 4 * (size_t)i;
            double* dv = s->acc + 4 * (size_t)i;
            for (int k = 0; k < 4; ++k) {
                r[k] += v[k] * dt + dv[k] * dt2;
                v[k] += dv[k] * dt;
            }
            s->rho[i] += s->drho[i] * dt;
            s->u[i] += s->du[i] * dt;
            if (s->damage && s->ddamage) {
                s->damage[i] += s->ddamage[i] * dt;
            }
            if (solid) {
                for (int k = 0; k < 5; ++k) {
                    s->S[5 * (size_t)i + k] += s->dS[5 * (size_t)i + k] * dt;
                }
            }
            step_first_order_clamp(m, s, i);
            /* swap into predictions + zero */
            for (int k = 0; k < 4; ++k) {
                s->acc_pred[4 * (size_t)i + k] = dv[k];
                dv[k] = 0.;
            }
            s->drho_pred[i] = s->drho[i];
            s->drho[i] = 0.;
            s->du_pred[i] = s->du[i];
            s->du[i] = 0.;
            if (s->damage && s->ddamage) {
                s->ddamage_pred[i] = s->ddamage[i];
                s->ddamage[i] = 0.;
            }
            if (solid) {
                for (int k = 0; k < 5; ++k) {
                    s->dS_pred[5 * (size_t)i + k] = s->dS[5 * (size_t)i + k];
                    s->dS[5 * (size_t)i + k] = 0.;
                }
            }
        }
    }
}

/* PredictorCorrector::makeCorrections, TimeStepping.cpp:302-322: storage1 = *storage (p*), storage2 = predictions (c*) */
void orc_correct(orc_state* s, const sphgpu_config* cfg, const sphgpu_material* mats, uint32_t nmat, double dt) {
    (void)cfg;
    const double dt2 = 0.5 * sqr(dt);
    const double a = 1. / 3., b = 0.5;
    const int solid = s->S != NULL;
    for (uint32_t mi = 0; mi < nmat; ++mi) {
        const sphgpu_material* m = &mats[mi];
        for (uint32_t i = m->begin; i < m->end; ++i) {
            double* r = s->pos + 4 * (size_t)i;
            double* v = s->vel + 4 * (size_t)i;
            const double* pdv = s->acc + 4 * (size_t)i;
            const double* cdv = s->acc_pred + 4 * (size_t)i;
            for (int k = 0; k < 4; ++k) {
                r[k] -= a * (cdv[k] - pdv[k]) * dt2;
                v[k] -= b * (cdv[k] - pdv[k]) * dt;
            }
            s->rho[i] -= 0.5 * (s->drho_pred[i] - s->drho[i]) * dt;
            s->u[i] -= 0.5 * (s->du_pred[i] - s->du[i]) * dt;
            if (s->damage && s->ddamage) {
                s->damage[i] -= 0.5 * (s->ddamage_pred[i] - s->ddamage[i]) * dt;
            }
            if (solid) {
                for (int k = 0; k < 5; ++k) {
                    s->S[5 * (size_t)i + k] -= 0.5 * (s->dS_pred[5 * (size_t)i + k] - s->dS[5 * (size_t)i + k]) * dt;
                }
            }
            step_first_order_clamp(m, s, i);
        }
    }
}

/* EulerExplicit::stepParticles after solver.integrate, TimeStepping.cpp:243-264 */
void orc_euler(orc_state* s, const sphgpu_config* cfg, const sphgpu_material* mats, uint32_t nmat, double dt) {
    (void)cfg;
    const int solid = s->S != NULL;
    for (uint32_t mi = 0; mi < nmat; ++mi) {
        const sphgpu_material* m = &mats[mi];
        for (uint32_t i = m->begin; i < m->end; ++i) {
            double* r = s->pos + 4 * (size_t)i;
            double* v = s->vel + 4 * (size_t)i;
            const double* dv = s->acc + 4 * (size_t)i;
            for (int k = 0; k < 4; ++k) {
                v[k] += dv[k] * dt;
            }
            for (int k = 0; k < 4; ++k) {
                r[k] += v[k] * dt;
            }
            s->rho[i] += s->drho[i] * dt;
            s->u[i] += s->du[i] * dt;
            if (s->damage && s->ddamage) {
                s->damage[i] += s->ddamage[i] * dt;
            }
            if (solid) {
                for (int k = 0; k < 5; ++k) {
                    s->S[5 * (size_t)i + k] += s->dS[5 * (size_t)i + k] * dt;
                }
            }
            step_first_order_clamp(m, s, i);
        }
    }
}

/* One first-order component of DerivativeCriterion::computeImpl, TimeStepCriterion.cpp:150-176 */
static double derivative_step(double absv, double absdv, double minValue, double factor) {
    if (fabs(absv) < 2. * minValue) {
        return ORC_INFTY;
    }
    return factor * (absv + minValue) / (absdv + ORC_EPS);
}

/* MultiCriterion::compute and the four criteria, core/timestepping/TimeStepCriterion.cpp:117-419 */
double orc_timestep(const orc_state* s, const sphgpu_config* cfg, const sphgpu_material* mats, uint32_t nmat,
    double max_dt, double* last_dt, uint32_t* criterion) {
    double minStep = ORC_INFTY;
    uint32_t minId = SPHGPU_CRITID_INITIAL_VALUE;
    if (cfg->criteria & SPHGPU_CRIT_COURANT) {
        /* CourantCriterion::compute, :327-365 */
        double step = ORC_INFTY;
        for (uint32_t i = 0; i < s->n; ++i) {
            if (s->cs[i] > 0.) {
                step = dmin(step, cfg->courant * s->pos[4 * (size_t)i + 3] / s->cs[i]);
            }
        }
        uint32_t id = SPHGPU_CRITID_CFL_CONDITION;
        if (step > max_dt) {
            step = max_dt;
            id = SPHGPU_CRITID_MAXIMAL_VALUE;
        }
        if (step < minStep) {
            minStep = step;
            minId = id;
        }
    }
    if (cfg->criteria & SPHGPU_CRIT_DERIVATIVES) {
        /* DerivativeCriterion::computeImpl<MinimalStepTls>, :139-214 (mean power -INFTY => minimum) */
        double step = ORC_INFTY;
        const double f = cfg->derivative_factor;
        for (uint32_t mi = 0; mi < nmat; ++mi) {
            const sphgpu_material* m = &mats[mi];
            for (uint32_t i = m->begin; i < m->end; ++i) {
                step = dmin(step, derivative_step(fabs(s->rho[i]), fabs(s->drho[i]), m->rho_small, f));
                step = dmin(step, derivative_step(fabs(s->u[i]), fabs(s->du[i]), m->u_small, f));
                if (s->damage && s->ddamage && m->fracture != SPHGPU_FRACTURE_NONE) {
                    step = dmin(step, derivative_step(fabs(s->damage[i]), fabs(s->ddamage[i]), m->d_small, f));
                }
                if (s->S) {
                    /* abs(TracelessTensor) -> SymmetricTensor(abs(diag), abs(off)), TracelessTensor.h:325-327;
                     * getComponents: xx,yy,zz,xy,xz,yz */
                    const double* S = s->S + 5 * (size_t)i;
                    const double* dS = s->dS + 5 * (size_t)i;
                    const double vs[6] = { fabs(S[0]), fabs(S[1]), fabs(-S[0] - S[1]), fabs(S[2]), fabs(S[3]), fabs(S[4]) };
                    const double dvs[6] = { fabs(dS[0]), fabs(dS[1]), fabs(-dS[0] - dS[1]), fabs(dS[2]), fabs(dS[3]),
                        fabs(dS[4]) };
                    for (int k = 0; k < 6; ++k) {
                        step = dmin(step, derivative_step(vs[k], dvs[k], m->s_small, f));
                    }
                }
            }
        }
        uint32_t id = SPHGPU_CRITID_DERIVATIVE;
        if (step > max_dt) {
            step = max_dt;
            id = SPHGPU_CRITID_MAXIMAL_VALUE;
        }
        if (step < minStep) {
            minStep = step;
            minId = id;
        }
    }
    if (cfg->criteria & SPHGPU_CRIT_ACCELERATION) {
        /* AccelerationCriterion::compute, :228-266 */
        double step = ORC_INFTY;
        for (uint32_t i = 0; i < s->n; ++i) {
            const double* dv = s->acc + 4 * (size_t)i;
            const double dvNorm = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
            if (dvNorm > ORC_EPS) {
                step = dmin(step, cfg->derivative_factor * sqrt(sqrt(sqr(s->pos[4 * (size_t)i + 3]) / dvNorm)));
            }
        }
        uint32_t id = SPHGPU_CRITID_ACCELERATION;
        if (step > max_dt) {
            step = max_dt;
            id = SPHGPU_CRITID_MAXIMAL_VALUE;
        }
        if (step < minStep) {
            minStep = step;
            minId = id;
        }
    }
    if (cfg->criteria & SPHGPU_CRIT_DIVERGENCE) {
        /* DivergenceCriterion::compute, :276-317 */
        double step = ORC_INFTY;
        for (uint32_t i = 0; i < s->n; ++i) {
            const double dv = fabs(s->divv[i]);
            if (dv > ORC_EPS) {
                step = dmin(step, cfg->divergence_factor / dv);
            }
        }
        uint32_t id = SPHGPU_CRITID_DIVERGENCE;
        if (step > max_dt) {
            step = max_dt;
            id = SPHGPU_CRITID_MAXIMAL_VALUE;
        }
        if (step < minStep) {
            minStep = step;
            minId = id;
        }
    }
    /* MultiCriterion smoothing, :403-413 */
    if (cfg->max_change < 1.e300) {
        const double maxStep = *last_dt * (1. + cfg->max_change);
        if (minStep > maxStep) {
            minStep = maxStep;
            minId = SPHGPU_CRITID_MAX_CHANGE;
        }
        *last_dt = minStep;
    }
    if (criterion) {
        *criterion = minId;
    }
    return minStep;
}

/* ---- self-gravity ----------------------------------------------------------------------------------------------- */

/* GravityKernel<CubicSpline<3>>::gradImpl, core/sph/kernel/GravityKernel.h:108-121. */
static double gravity_cubic_grad(double qSqr) {
    const double q = sqrt(qSqr);
    if (q == 0.) {
        return 4. / 3.;
    } else if (q < 1.) {
        return 1. / q * (4. / 3. * q - 6. / 5. * q * q * q + 1. / 2. * qSqr * qSqr);
    }
    return 1. / q * (8. / 3. * q - 3. * qSqr + 6. / 5. * q * q * q - 1. / 6. * qSqr * qSqr - 1. / (15. * qSqr));
}

/* LutKernel<3>::LutKernel(GravityKernel<CubicSpline<3>>), Kernel.h:85-101: NEntries + 1 node values over q^2. */
void orc_build_gravity_lut(double* grad, uint32_t entries) {
    const double radInvSqr = 1. / (2. * 2.);
    const double qSqrToIdx = (double)entries * radInvSqr;
    for (uint32_t i = 0; i <= entries; ++i) {
        grad[i] = gravity_cubic_grad((double)i / qSqrToIdx);
    }
}

/* GravityLutKernel::grad(r, h), GravityKernel.h:72-86, with LutKernel<3>::gradImpl (Kernel.h:129-144). */
static void gravity_kernel_grad(const double r[3], double h, const double* lut, uint32_t entries, double radius, double out[3]) {
    const double hInv = 1. / h;
    const double sx = r[0] * hInv, sy = r[1] * hInv, sz = r[2] * hInv;
    const double qSqr = sx * sx + sy * sy + sz * sz;
    if (qSqr + ORC_EPS >= sqr(radius)) {
        const double len = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        const double len3 = len * len * len;
        out[0] = r[0] / len3;
        out[1] = r[1] / len3;
        out[2] = r[2] / len3;
    } else {
        const double floatIdx = (double)entries / sqr(radius) * qSqr;
        const uint32_t idx1 = (uint32_t)floatIdx;
        const double ratio = floatIdx - (double)idx1;
        const double grad = lut[idx1] * (1. - ratio) + lut[idx1 + 1] * ratio;
        const double f = hInv * hInv * hInv;
        out[0] = f * r[0] * grad;
        out[1] = f * r[1] * grad;
        out[2] = f * r[2] * grad;
    }
}

/* BruteForceGravity::evalImpl, BruteForceGravity.h:103-121: a_i = sum_{j != i} G m_j grad(r_j, r_i), the kernel taken
 * at the mean smoothing length (SymmetrizeSmoothingLengths::grad, Kernel.h:640-643). */
void orc_gravity_brute(uint32_t n, const double* pos, const double* mass, double G, const double* lut_grad, uint32_t entries,
    double radius, double* acc) {
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n; ++i) {
        double a[3] = { 0., 0., 0. };
        for (uint32_t j = 0; j < n; ++j) {
            if (j == i) {
                continue;
            }
            const double r[3] = { pos[4 * j] - pos[4 * i], pos[4 * j + 1] - pos[4 * i + 1], pos[4 * j + 2] - pos[4 * i + 2] };
            double g[3];
            gravity_kernel_grad(r, 0.5 * (pos[4 * j + 3] + pos[4 * i + 3]), lut_grad, entries, radius, g);
            a[0] += mass[j] * g[0];
            a[1] += mass[j] * g[1];
            a[2] += mass[j] * g[2];
        }
        acc[3 * i] = G * a[0];
        acc[3 * i + 1] = G * a[1];
        acc[3 * i + 2] = G * a[2];
    }
}

/* BarnesHut::buildLeaf, BarnesHut.cpp:393-431: com = sum m r / sum m; M2, M3 = computeMultipole<2>, <3> about com
 * (Moments.h:166-179); reduced (traceless) multipoles computeReducedMultipole (Moments.h:92-111):
 * Q2 = M2 - delta tr(M2) / 3, Q3_ijk = M3_ijk - (delta_ij T_k + delta_ik T_j + delta_jk T_i) / 5, T_k = M3_llk. */
void orc_gravity_moments(uint32_t n, const double* pos, const double* mass, double* com, double* mom) {
    double c[3] = { 0., 0., 0. }, m0 = 0.;
    for (uint32_t i = 0; i < n; ++i) {
        for (int k = 0; k < 3; ++k) {
            c[k] += mass[i] * pos[4 * i + k];
        }
        m0 += mass[i];
    }
    for (int k = 0; k < 3; ++k) {
        c[k] /= m0;
        com[k] = c[k];
    }
    double M2[3][3], M3[3][3][3];
    memset(M2, 0, sizeof(M2));
    memset(M3, 0, sizeof(M3));
    for (uint32_t i = 0; i < n; ++i) {
        const double d[3] = { pos[4 * i] - c[0], pos[4 * i + 1] - c[1], pos[4 * i + 2] - c[2] };
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) {
                M2[a][b] += d[a] * d[b] * mass[i];
                for (int e = 0; e < 3; ++e) {
                    M3[a][b][e] += d[a] * d[b] * d[e] * mass[i];
                }
            }
        }
    }
    const double tr = M2[0][0] + M2[1][1] + M2[2][2];
    double T[3];
    for (int k = 0; k < 3; ++k) {
        T[k] = M3[0][0][k] + M3[1][1][k] + M3[2][2][k];
    }
    double Q2[3][3], Q3[3][3][3];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) {
            Q2[a][b] = M2[a][b] - (a == b ? tr / 3. : 0.);
            for (int e = 0; e < 3; ++e) {
                Q3[a][b][e] = M3[a][b][e] - ((a == b ? T[e] : 0.) + (a == e ? T[b] : 0.) + (b == e ? T[a] : 0.)) / 5.;
            }
        }
    }
    mom[0] = m0;
    mom[1] = Q2[0][0]; mom[2] = Q2[1][1]; mom[3] = Q2[0][1]; mom[4] = Q2[0][2]; mom[5] = Q2[1][2];
    mom[6] = Q3[0][0][0]; mom[7] = Q3[0][0][1]; mom[8] = Q3[0][0][2]; mom[9] = Q3[0][1][1]; mom[10] = Q3[0][1][2];
    mom[11] = Q3[1][1][1]; mom[12] = Q3[1][1][2];
}

/* evaluateGravity, Moments.h:315-340: gamma from computeGreenGamma (:22-29); per order M the acceleration
 * gamma[M+1] dr Q0 + gamma[M] Q1 with Q0 = q . dr^M / M!, Q1 = q . dr^(M-1) / (M-1)! (computeMultipoleAcceleration
 * :292-304, computeMultipolePotential :128-136), evaluated at dr = -(point - com). */
void orc_gravity_multipole(const double* com, const double* mom, int order, const double* point, double* acc) {
    double Q2[3][3], Q3[3][3][3];
    Q2[0][0] = mom[1]; Q2[1][1] = mom[2]; Q2[2][2] = -mom[1] - mom[2];
    Q2[0][1] = Q2[1][0] = mom[3]; Q2[0][2] = Q2[2][0] = mom[4]; Q2[1][2] = Q2[2][1] = mom[5];
    /* the ten symmetric components from the seven independent ones (traceless in every index pair) */
    double s[10];
    s[0] = mom[6]; s[1] = mom[7]; s[2] = mom[8]; s[3] = mom[9]; s[4] = mom[10]; s[6] = mom[11]; s[7] = mom[12];
    s[5] = -s[0] - s[3]; /* xzz */
    s[8] = -s[1] - s[6]; /* yzz */
    s[9] = -s[2] - s[7]; /* zzz */
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) {
            for (int e = 0; e < 3; ++e) {
                const int nx = (a == 0) + (b == 0) + (e == 0), ny = (a == 1) + (b == 1) + (e == 1);
                /* (nx, ny) -> xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz */
                int idx;
                if (nx == 3) idx = 0;
                else if (nx == 2) idx = ny == 1 ? 1 : 2;
                else if (nx == 1) idx = ny == 2 ? 3 : (ny == 1 ? 4 : 5);
                else idx = ny == 3 ? 6 : (ny == 2 ? 7 : (ny == 1 ? 8 : 9));
                Q3[a][b][e] = s[idx];
            }
        }
    }
    const double dr0[3] = { point[0] - com[0], point[1] - com[1], point[2] - com[2] };
    const double invDistSqr = 1. / (dr0[0] * dr0[0] + dr0[1] * dr0[1] + dr0[2] * dr0[2]);
    double gamma[5];
    gamma[0] = -sqrt(invDistSqr);
    for (int i = 1; i < 5; ++i) {
        gamma[i] = -(2. * i - 1.) * invDistSqr * gamma[i - 1];
    }
    const double dr[3] = { -dr0[0], -dr0[1], -dr0[2] };
    double a[3] = { 0., 0., 0. };
    if (order >= 3) {
        double q0 = 0., q1[3] = { 0., 0., 0. };
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) {
                for (int k = 0; k < 3; ++k) {
                    q0 += Q3[i][j][k] * dr[i] * dr[j] * dr[k] / 6.;
                    q1[i] += Q3[i][j][k] * dr[j] * dr[k] / 2.;
                }
            }
        }
        for (int k = 0; k < 3; ++k) {
            a[k] += gamma[4] * dr[k] * q0 + gamma[3] * q1[k];
        }
    }
    if (order >= 2) {
        double q0 = 0., q1[3] = { 0., 0., 0. };
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) {
                q0 += Q2[i][j] * dr[i] * dr[j] / 2.;
                q1[i] += Q2[i][j] * dr[j];
            }
        }
        for (int k = 0; k < 3; ++k) {
            a[k] += gamma[3] * dr[k] * q0 + gamma[2] * q1[k];
        }
    }
    for (int k = 0; k < 3; ++k) {
        a[k] += gamma[1] * dr[k] * mom[0];
    }
    acc[0] = a[0];
    acc[1] = a[1];
    acc[2] = a[2];
}

/* ---- initial conditions --------------------------------------------------------------------------------------------- */

/* HexagonalPacking::generate (Distribution.cpp:126-200) inside a SphericalDomain (Domain.h:134-136), the loops of
 * Box::iterateWithIndices (Box.h:182-195) with their running coordinate sums, the CENTER option (:186-198), then
 * InitialConditions::setQuantities / getMasses (Initial.cpp:288-318). */
uint32_t orc_hexagonal_sphere(uint32_t n, const double* center, double radius, int centred, double eta, double rho0, double* pos,
    double* mass, uint32_t capacity) {
    const double volume = 1.3333333333333333333333 * M_PI * (radius * radius * radius);
    const double particleDensity = (double)n / volume;
    const double h = 1. / cbrt(particleDensity);
    const double dx = 1.1 * h;
    const double dy = sqrt(3.) * 0.5 * dx;
    const double dz = sqrt(6.) / 3. * dx;
    const double lo[3] = { center[0] - radius + 0.5 * dx, center[1] - radius + 0.5 * dy, center[2] - radius + 0.5 * dz };
    const double hi[3] = { center[0] + radius, center[1] + radius, center[2] + radius };
    const double deltaX = 0.5 * dx;
    const double deltaY = sqrt(3.) / 6. * dx;
    uint32_t cnt = 0;
    uint32_t i, j, k = 0;
    for (double z = lo[2]; z <= hi[2]; z += dz, k++) {
        j = 0;
        for (double y = lo[1]; y <= hi[1]; y += dy, j++) {
            i = 0;
            for (double x = lo[0]; x <= hi[0]; x += dx, i++) {
                double v[3] = { x, y, z };
                if (k % 2 == 0) {
                    if (j % 2 == 1) {
                        v[0] += deltaX;
                    }
                } else {
                    if (j % 2 == 0) {
                        v[0] += deltaX;
                    }
                    v[1] += deltaY;
                }
                const double d[3] = { v[0] - center[0], v[1] - center[1], v[2] - center[2] };
                if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] <= sqr(radius)) {
                    if (pos && cnt < capacity) {
                        pos[4 * cnt] = v[0];
                        pos[4 * cnt + 1] = v[1];
                        pos[4 * cnt + 2] = v[2];
                        pos[4 * cnt + 3] = h;
                    }
                    cnt++;
                }
            }
        }
    }
    if (!pos || cnt > capacity) {
        return cnt;
    }
    if (centred) {
        double com[3] = { 0., 0., 0. };
        for (uint32_t p = 0; p < cnt; ++p) {
            for (int a = 0; a < 3; ++a) {
                com[a] += pos[4 * p + a];
            }
        }
        for (int a = 0; a < 3; ++a) {
            com[a] /= (double)cnt;
        }
        for (uint32_t p = 0; p < cnt; ++p) {
            for (int a = 0; a < 3; ++a) {
                pos[4 * p + a] += center[a] - com[a];
            }
        }
    }
    const double totalM = volume * rho0;
    double prelimM = 0.;
    for (uint32_t p = 0; p < cnt; ++p) {
        pos[4 * p + 3] *= eta;
        const double hp = pos[4 * p + 3];
        mass[p] = hp * hp * hp;
        prelimM += mass[p];
    }
    const double normalization = totalM / prelimM;
    for (uint32_t p = 0; p < cnt; ++p) {
        mass[p] *= normalization;
    }
    return cnt;
}

/* ---- boundary condition ------------------------------------------------------------------------------------------------ */

/* FrozenParticles::finalize, Boundary.cpp:221-258; SphericalDomain::getSubset / project / getDistanceToBoundary,
 * Domain.cpp:35-85. */
void orc_frozen(orc_state* s, int solid, uint64_t flag_mask, int has_domain, const double* center, double radius, double freeze_radius) {
    for (uint32_t i = 0; i < s->n; ++i) {
        double* r = s->pos + 4 * (size_t)i;
        int frozen = 0;
        if (has_domain) {
            double d[3] = { r[0] - center[0], r[1] - center[1], r[2] - center[2] };
            if (!(d[0] * d[0] + d[1] * d[1] + d[2] * d[2] <= sqr(radius))) {
                const double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                for (int k = 0; k < 3; ++k) {
                    r[k] = d[k] / len * (1. - ORC_EPS) * radius + center[k];
                    d[k] = r[k] - center[k];
                }
            }
            const double dist = radius - sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (dist < freeze_radius * r[3]) {
                frozen = 1;
            }
        }
        if (s->flag && s->flag[i] < 64u && ((flag_mask >> s->flag[i]) & 1ull)) {
            frozen = 1;
        }
        if (!frozen) {
            continue;
        }
        /* iterate<HIGHEST_DERIVATIVES>: POSITION d2t, DENSITY dt, ENERGY dt, DEVIATORIC_STRESS dt (and DAMAGE dt, which
         * material->finalize overwrites afterwards) */
        for (int k = 0; k < 4; ++k) {
            s->acc[4 * (size_t)i + k] = 0.;
        }
        s->drho[i] = 0.;
        s->du[i] = 0.;
        if (solid && s->dS) {
            memset(s->dS + 5 * (size_t)i, 0, 5 * sizeof(double));
        }
    }
}

/* ---- connected components ------------------------------------------------------------------- */

/* Post::findComponents without ESCAPE_VELOCITY / SORT_BY_MASS = findComponentsImpl (core/post/Analysis.cpp:36-75,115-128):
 * particles are visited in index order; an unassigned one opens the next component and floods it through a stack over the
 * DIRECTED relation "j lies within h_index * radius of index" (IBasicFinder::findAll(index, r[index][H] * radius): distSqr <
 * radius^2, KdTree.inl.h), restricted to equal body flags when flag != NULL (FlagComponentChecker, Analysis.cpp:84-94).
 * A uniform grid with cells of the largest reach replaces the reference's k-d tree (same candidate sets). */
uint32_t orc_find_components(const double* pos, uint32_t n, double radius, const uint32_t* flag, uint32_t* indices) {
    if (n == 0) {
        return 0;
    }
    double lo[3] = { pos[0], pos[1], pos[2] }, hi[3] = { pos[0], pos[1], pos[2] }, reachMax = 0.;
    for (uint32_t i = 0; i < n; ++i) {
        for (int q = 0; q < 3; ++q) {
            lo[q] = dmin(lo[q], pos[4 * (size_t)i + q]);
            hi[q] = dmax(hi[q], pos[4 * (size_t)i + q]);
        }
        reachMax = dmax(reachMax, pos[4 * (size_t)i + 3] * radius);
    }
    int dim[3];
    double cell = reachMax * (1. + 1.e-9);
    for (int q = 0; q < 3; ++q) {
        cell = dmax(cell, (hi[q] - lo[q]) / 200.); /* at most 200 cells per axis */
    }
    if (!(cell > 0.)) {
        cell = 1.;
    }
    for (int q = 0; q < 3; ++q) {
        dim[q] = (int)((hi[q] - lo[q]) / cell) + 1;
    }
    const size_t ncell = (size_t)dim[0] * dim[1] * dim[2];
    uint32_t* start = (uint32_t*)calloc(ncell + 1, sizeof(uint32_t));
    uint32_t* items = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint32_t* cellOf = (uint32_t*)malloc(sizeof(uint32_t) * n);
    for (uint32_t i = 0; i < n; ++i) {
        int c[3];
        for (int q = 0; q < 3; ++q) {
            c[q] = (int)((pos[4 * (size_t)i + q] - lo[q]) / cell);
            c[q] = c[q] < 0 ? 0 : (c[q] >= dim[q] ? dim[q] - 1 : c[q]);
        }
        cellOf[i] = (uint32_t)((c[2] * dim[1] + c[1]) * dim[0] + c[0]);
        start[cellOf[i] + 1]++;
    }
    for (size_t c = 0; c < ncell; ++c) {
        start[c + 1] += start[c];
    }
    uint32_t* fill = (uint32_t*)malloc(sizeof(uint32_t) * (ncell + 1));
    memcpy(fill, start, sizeof(uint32_t) * (ncell + 1));
    for (uint32_t i = 0; i < n; ++i) {
        items[fill[cellOf[i]]++] = i;
    }
    free(fill);
    const uint32_t unassigned = 0xffffffffu;
    for (uint32_t i = 0; i < n; ++i) {
        indices[i] = unassigned;
    }
    uint32_t* stack = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint32_t componentIdx = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (indices[i] != unassigned) {
            continue;
        }
        uint32_t sp = 0;
        indices[i] = componentIdx;
        stack[sp++] = i;
        while (sp > 0) {
            const uint32_t index = stack[--sp];
            const double* ri = pos + 4 * (size_t)index;
            const double reach = ri[3] * radius;
            const double reachSqr = reach * reach;
            const uint32_t c = cellOf[index];
            const int cx = (int)(c % (uint32_t)dim[0]), cy = (int)((c / (uint32_t)dim[0]) % (uint32_t)dim[1]),
                      cz = (int)(c / ((uint32_t)dim[0] * (uint32_t)dim[1]));
            for (int z = cz > 0 ? cz - 1 : 0; z <= (cz + 1 < dim[2] ? cz + 1 : dim[2] - 1); ++z) {
                for (int y = cy > 0 ? cy - 1 : 0; y <= (cy + 1 < dim[1] ? cy + 1 : dim[1] - 1); ++y) {
                    for (int x = cx > 0 ? cx - 1 : 0; x <= (cx + 1 < dim[0] ? cx + 1 : dim[0] - 1); ++x) {
                        const size_t cc = ((size_t)z * dim[1] + y) * dim[0] + x;
                        for (uint32_t k = start[cc]; k < start[cc + 1]; ++k) {
                            const uint32_t j = items[k];
                            if (indices[j] != unassigned) {
                                continue;
                            }
                            const double* rj = pos + 4 * (size_t)j;
                            const double dx = rj[0] - ri[0], dy = rj[1] - ri[1], dz = rj[2] - ri[2];
                            if (!(dx * dx + dy * dy + dz * dz < reachSqr)) {
                                continue;
                            }
                            if (flag && flag[index] != flag[j]) {
                                continue;
                            }
                            indices[j] = componentIdx;
                            stack[sp++] = j;
                        }
                    }
                }
            }
        }
        componentIdx++;
    }
    free(stack);
    free(cellOf);
    free(items);
    free(start);
    return componentIdx;
}
