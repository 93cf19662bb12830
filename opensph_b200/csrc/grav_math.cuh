// Arithmetic of the self-gravity path: multipole moments of a particle set about its centre of mass (monopole,
// traceless quadrupole and octupole), their shift to another centre, the acceleration they produce at a point, and the
// softened particle-particle attraction. Host + device (tests/test_host_math.py builds these functions for the host).
//
// Reference: core/gravity/Moments.h (computeMultipole :166-179, computeReducedMultipole :72-111, parallelAxisTheorem
// :181-290, computeGreenGamma :22-29, computeMultipoleAcceleration :292-304, evaluateGravity :315-340),
// core/gravity/BarnesHut.cpp (buildLeaf :371-432, buildInner :434-489), core/sph/kernel/GravityKernel.h:58-86
// (GravityLutKernel::grad) and core/sph/kernel/Kernel.h:640-643 (SymmetrizeSmoothingLengths::grad).
//
// The reference carries the traceless tensors through the tree with its parallel-axis formulas; here every node keeps
// the RAW moments M2_ij = sum m x_i x_j, M3_ijk = sum m x_i x_j x_k about its centre of mass (16 doubles), which shift by
//   M2'_ij  = M2_ij + m d_i d_j,     M3'_ijk = M3_ijk + d_i M2_jk + d_j M2_ik + d_k M2_ij + m d_i d_j d_k
// (the first moment vanishes about the centre of mass), and the traceless tensors are formed once per node:
//   Q2_ij = M2_ij - delta_ij tr(M2) / 3,   Q3_ijk = M3_ijk - (delta_ij T_k + delta_ik T_j + delta_jk T_i) / 5,  T_k = M3_llk
// (reducedFactor<2,1> = -1/3, reducedFactor<3,1> = -1/5). Mathematically the same tensors as the reference's.
#pragma once
#include "sph_math.cuh"

namespace sph {

/// Raw second and third moments about the centre of mass. m2: xx yy zz xy xz yz; m3: xxx xxy xxz xyy xyz xzz yyy yyz yzz zzz.
struct GravRaw {
    double m2[6];
    double m3[10];
};

/// What a node contributes as a source: centre of mass, mass (times the gravitational constant) and the independent
/// components of the traceless tensors. q2: xx yy xy xz yz (zz = -xx - yy); q3: xxx xxy xxz xyy xyz yyy yyz
/// (xzz = -xxx - xyy, yzz = -xxy - yyy, zzz = -xxz - yyz). 16 doubles = 128 bytes.
struct GravNode {
    double cx, cy, cz, m;
    double q2[5];
    double q3[7];
};

SPH_HD void gravRawZero(GravRaw& r) {
    for (int k = 0; k < 6; ++k) {
        r.m2[k] = 0.;
    }
    for (int k = 0; k < 10; ++k) {
        r.m3[k] = 0.;
    }
}

/// Adds a point mass m at offset (x, y, z) from the centre (computeMultipole<2>, <3>).
SPH_HD void gravRawAddPoint(GravRaw& r, double m, double x, double y, double z) {
    const double mx = m * x, my = m * y, mz = m * z;
    r.m2[0] += mx * x;
    r.m2[1] += my * y;
    r.m2[2] += mz * z;
    r.m2[3] += mx * y;
    r.m2[4] += mx * z;
    r.m2[5] += my * z;
    const double mxx = mx * x, mxy = mx * y, myy = my * y, mzz = mz * z;
    r.m3[0] += mxx * x;
    r.m3[1] += mxx * y;
    r.m3[2] += mxx * z;
    r.m3[3] += mxy * y;
    r.m3[4] += mxy * z;
    r.m3[5] += mzz * x;
    r.m3[6] += myy * y;
    r.m3[7] += myy * z;
    r.m3[8] += mzz * y;
    r.m3[9] += mzz * z;
}

/// Adds the moments `c` of a child of mass m whose centre of mass sits at offset d from the new centre.
SPH_HD void gravRawAddShifted(GravRaw& r, const GravRaw& c, double m, double dx, double dy, double dz) {
    const double xx = c.m2[0], yy = c.m2[1], zz = c.m2[2], xy = c.m2[3], xz = c.m2[4], yz = c.m2[5];
    r.m3[0] += c.m3[0] + 3. * dx * xx;
    r.m3[1] += c.m3[1] + 2. * dx * xy + dy * xx;
    r.m3[2] += c.m3[2] + 2. * dx * xz + dz * xx;
    r.m3[3] += c.m3[3] + dx * yy + 2. * dy * xy;
    r.m3[4] += c.m3[4] + dx * yz + dy * xz + dz * xy;
    r.m3[5] += c.m3[5] + dx * zz + 2. * dz * xz;
    r.m3[6] += c.m3[6] + 3. * dy * yy;
    r.m3[7] += c.m3[7] + 2. * dy * yz + dz * yy;
    r.m3[8] += c.m3[8] + dy * zz + 2. * dz * yz;
    r.m3[9] += c.m3[9] + 3. * dz * zz;
    for (int k = 0; k < 6; ++k) {
        r.m2[k] += c.m2[k];
    }
    gravRawAddPoint(r, m, dx, dy, dz);
}

/// Traceless tensors of the raw moments (computeReducedMultipole).
SPH_HD void gravReduce(const GravRaw& r, GravNode& n) {
    const double tr = (r.m2[0] + r.m2[1] + r.m2[2]) * (1. / 3.);
    n.q2[0] = r.m2[0] - tr;
    n.q2[1] = r.m2[1] - tr;
    n.q2[2] = r.m2[3];
    n.q2[3] = r.m2[4];
    n.q2[4] = r.m2[5];
    const double tx = (r.m3[0] + r.m3[3] + r.m3[5]) * 0.2, ty = (r.m3[1] + r.m3[6] + r.m3[8]) * 0.2, tz = (r.m3[2] + r.m3[7] + r.m3[9]) * 0.2;
    n.q3[0] = r.m3[0] - 3. * tx; // xxx
    n.q3[1] = r.m3[1] - ty;      // xxy
    n.q3[2] = r.m3[2] - tz;      // xxz
    n.q3[3] = r.m3[3] - tx;      // xyy
    n.q3[4] = r.m3[4];           // xyz
    n.q3[5] = r.m3[6] - 3. * ty; // yyy
    n.q3[6] = r.m3[7] - tz;      // yyz
}

/// Opening radius of a node, Eq. (2.36) of Stadel's thesis as BarnesHut::buildLeaf / buildInner use it:
/// 2 / sqrt(3) / theta * |max(com - lower, upper - com)|.
SPH_HD double gravOpeningRadius(const double com[3], const double lo[3], const double hi[3], double thetaInv) {
    const double ex = fmax(com[0] - lo[0], hi[0] - com[0]), ey = fmax(com[1] - lo[1], hi[1] - com[1]), ez = fmax(com[2] - lo[2], hi[2] - com[2]);
    return 2. / sqrt(3.) * thetaInv * sqrt(ex * ex + ey * ey + ez * ez);
}

/// Sphere::overlaps (core/objects/geometry/Sphere.h:72-79): the opening ball of a node reaches into the box.
SPH_HD bool gravBallOverlapsBox(double cx, double cy, double cz, double rOpen, const double lo[3], const double hi[3]) {
    const double lx = fmax(lo[0] - cx, 0.), ly = fmax(lo[1] - cy, 0.), lz = fmax(lo[2] - cz, 0.);
    const double rx = fmax(cx - hi[0], 0.), ry = fmax(cy - hi[1], 0.), rz = fmax(cz - hi[2], 0.);
    return rOpen * rOpen - (lx * lx + ly * ly + lz * lz) - (rx * rx + ry * ry + rz * rz) > 0.;
}

/// Acceleration at (x, y, z) due to the node, evaluateGravity(r0 - com, moments, order): order 0 monopole, 2 adds the
/// quadrupole, 3 the octupole. Masses / moments already carry the gravitational constant.
template <int ORDER>
SPH_HD void gravNodeAccel(const GravNode& n, double x, double y, double z, double& ax, double& ay, double& az) {
    const double rx = n.cx - x, ry = n.cy - y, rz = n.cz - z; // -(r0 - com)
    const double inv2 = 1. / (rx * rx + ry * ry + rz * rz);
    const double g0 = -sqrt(inv2);
    const double g1 = -inv2 * g0;
    double fr = g1 * n.m; // coefficient of r
    double fx = 0., fy = 0., fz = 0.;
    if (ORDER >= 2) {
        const double g2 = -3. * inv2 * g1, g3 = -5. * inv2 * g2;
        const double qzz = -n.q2[0] - n.q2[1];
        const double q1x = n.q2[0] * rx + n.q2[2] * ry + n.q2[3] * rz;
        const double q1y = n.q2[2] * rx + n.q2[1] * ry + n.q2[4] * rz;
        const double q1z = n.q2[3] * rx + n.q2[4] * ry + qzz * rz;
        const double q0 = 0.5 * (q1x * rx + q1y * ry + q1z * rz);
        fr += g3 * q0;
        fx = g2 * q1x;
        fy = g2 * q1y;
        fz = g2 * q1z;
        if (ORDER >= 3) {
            const double g4 = -7. * inv2 * g3;
            const double xxx = n.q3[0], xxy = n.q3[1], xxz = n.q3[2], xyy = n.q3[3], xyz = n.q3[4], yyy = n.q3[5], yyz = n.q3[6];
            const double xzz = -xxx - xyy, yzz = -xxy - yyy, zzz = -xxz - yyz;
            const double xx = rx * rx, yy = ry * ry, zz = rz * rz, xy = 2. * rx * ry, xz = 2. * rx * rz, yz = 2. * ry * rz;
            const double o1x = 0.5 * (xxx * xx + xyy * yy + xzz * zz + xxy * xy + xxz * xz + xyz * yz);
            const double o1y = 0.5 * (xxy * xx + yyy * yy + yzz * zz + xyy * xy + xyz * xz + yyz * yz);
            const double o1z = 0.5 * (xxz * xx + yyz * yy + zzz * zz + xyz * xy + xzz * xz + yzz * yz);
            const double o0 = (1. / 3.) * (o1x * rx + o1y * ry + o1z * rz);
            fr += g4 * o0;
            fx += g3 * o1x;
            fy += g3 * o1y;
            fz += g3 * o1z;
        }
    }
    ax += fr * rx + fx;
    ay += fr * ry + fy;
    az += fr * rz + fz;
}

/// The same expansion in single precision for the tree walk. The multipole approximation itself is good to 1e-3 .. 1e-5 of
/// the field; rounding of 6e-8 per node does not show. FP32 runs at 64 times the FP64 rate on this device, so the far field
/// costs next to nothing and the FP64 pipe is left to the exact particle pairs. To stay inside the FP32 range the walk works
/// in scaled units (lengths in units of the bounding cube L, masses in units of the total mass M: every moment is <= 1)
/// and with the unit vector n = r / d instead of high powers of r:
///   a = m n / d^2  +  (7.5 (n.Q2 n) n - 3 Q2 n) / d^4  +  (7.5 Q3 n n - 17.5 (Q3 n n n) n) / d^5,   r = com - target.
struct GravNodeF {
    float rx, ry, rz, m; // centre of mass relative to the group's centre
    float q2[5];
    float q3[7];
};

template <int ORDER>
SPH_HD void gravNodeAccelF(const GravNodeF& n, float ox, float oy, float oz, float& ax, float& ay, float& az) {
    const float rx = n.rx - ox, ry = n.ry - oy, rz = n.rz - oz;
    const float d2 = rx * rx + ry * ry + rz * rz;
#ifdef __CUDA_ARCH__
    const float inv = rsqrtf(d2);
#else
    const float inv = 1.f / sqrtf(d2);
#endif
    const float nx = rx * inv, ny = ry * inv, nz = rz * inv;
    const float inv2 = inv * inv;
    float cn = n.m * inv2; // coefficient of the unit vector
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (ORDER >= 2) {
        const float inv4 = inv2 * inv2;
        const float qzz = -n.q2[0] - n.q2[1];
        const float qx = n.q2[0] * nx + n.q2[2] * ny + n.q2[3] * nz;
        const float qy = n.q2[2] * nx + n.q2[1] * ny + n.q2[4] * nz;
        const float qz = n.q2[3] * nx + n.q2[4] * ny + qzz * nz;
        cn += 7.5f * inv4 * (qx * nx + qy * ny + qz * nz);
        const float c2 = -3.f * inv4;
        fx = c2 * qx;
        fy = c2 * qy;
        fz = c2 * qz;
        if (ORDER >= 3) {
            const float inv5 = inv4 * inv;
            const float xxx = n.q3[0], xxy = n.q3[1], xxz = n.q3[2], xyy = n.q3[3], xyz = n.q3[4], yyy = n.q3[5], yyz = n.q3[6];
            const float xzz = -xxx - xyy, yzz = -xxy - yyy, zzz = -xxz - yyz;
            const float xx = nx * nx, yy = ny * ny, zz = nz * nz, xy = 2.f * nx * ny, xz = 2.f * nx * nz, yz = 2.f * ny * nz;
            const float tx = xxx * xx + xyy * yy + xzz * zz + xxy * xy + xxz * xz + xyz * yz;
            const float ty = xxy * xx + yyy * yy + yzz * zz + xyy * xy + xyz * xz + yyz * yz;
            const float tz = xxz * xx + yyz * yy + zzz * zz + xyz * xy + xzz * xz + yzz * yz;
            cn -= 17.5f * inv5 * (tx * nx + ty * ny + tz * nz);
            const float c3 = 7.5f * inv5;
            fx += c3 * tx;
            fy += c3 * ty;
            fz += c3 * tz;
        }
    }
    ax += cn * nx + fx;
    ay += cn * ny + fy;
    az += cn * nz + fz;
}

/// Run-wide constants of the gravity path.
struct GravParams {
    double thetaInv;       // 1 / opening angle
    double radiusSqr;      // squared radius of the softening kernel (in units of h); 0: point particles
    double qSqrToIdx;      // table index per unit of q^2
    uint32_t lutEntries;
    int order;             // 0, 2, 3
    int exact;             // every node is opened: all pairs exactly (BruteForceGravity)
    uint32_t leafSize;
};

/// m_j * SymmetrizeSmoothingLengths<GravityLutKernel>::grad(r_j, r_i): attraction of particle i by particle j (the mass
/// carries the gravitational constant). Newton's law outside the softening kernel (q^2 + EPS >= R^2, EPS = 1e-12f),
/// h^-3 r G(q^2) inside, G interpolated from the reference's own table.
SPH_HD void gravPairAccel(const GravParams& p, const LutPair* __restrict__ lut, double xi, double yi, double zi, double hi, double xj, double yj,
    double zj, double hj, double mj, double& ax, double& ay, double& az) {
    const double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    const double d2 = dx * dx + dy * dy + dz * dz;
    const double hbar = 0.5 * (hj + hi);
    // The reference forms q^2 = |r / hbar|^2 (GravityKernel.h:73-75). On the device the two divisions of a pair -- 1 / hbar
    // here and 1 / |r|^3 below -- are a seeded reciprocal and a reciprocal square root with Newton steps (1-2 ulp) instead
    // of the IEEE sequences, which cost about as much as the rest of the pair together; 1e-16 against a 1e-10 tolerance.
#ifdef __CUDA_ARCH__
    const double hInv = fastRcp(hbar);
#else
    const double hInv = 1. / hbar;
#endif
    const double qSqr = d2 * (hInv * hInv);
    double f;
    if (qSqr + (double)1.e-12f >= p.radiusSqr) {
#ifdef __CUDA_ARCH__
        const double rInv = rsqrt(d2);
        f = mj * (rInv * rInv * rInv);
#else
        const double d = sqrt(d2);
        f = mj / (d2 * d);
#endif
    } else {
        const double fidx = p.qSqrToIdx * qSqr;
        const uint32_t k = (uint32_t)fidx;
        const double ratio = fidx - (double)k;
#ifdef __CUDA_ARCH__
        const double2 e = __ldg(reinterpret_cast<const double2*>(lut) + k);
        const double G = fma(ratio, e.y, e.x);
#else
        const double G = lut[k].g + ratio * lut[k].dg;
#endif
        f = mj * (hInv * hInv * hInv) * G;
    }
    ax += f * dx;
    ay += f * dy;
    az += f * dz;
}

} // namespace sph
