#include "GpuSolver.h"
#include "../../include/sphgpu.h"
#include "objects/Exceptions.h"
#include "objects/geometry/Domain.h"
#include "physics/Constants.h"
#include "quantities/IMaterial.h"
#include "quantities/Storage.h"
#include "sph/kernel/GravityKernel.h"
#include "sph/equations/av/Balsara.h"
#include "sph/equations/av/Standard.h"
#include "sph/equations/DeltaSph.h"
#include "sph/equations/XSph.h"
#include "sph/equations/av/Stress.h"
#include "system/Factory.h"
#include "system/Statistics.h"
#include "system/Timer.h"
#include "thread/Scheduler.h"
#include <algorithm>
#include <set>
#include <typeinfo>
#include <string>
#include <vector>

NAMESPACE_SPH_BEGIN

namespace {

/// Non-zero C-ABI status -> the exception the reference would have thrown (InvalidSetup from constructors / create /
/// sanityCheck, Exception otherwise; core/objects/Exceptions.h).
void check(const int rc) {
    if (rc == SPHGPU_OK) {
        return;
    }
    const String msg = String::fromAscii(sphgpu_last_error());
    if (rc == SPHGPU_E_INVALID || rc == SPHGPU_E_NO_DEVICE) {
        throw InvalidSetup("GpuSolver: " + msg);
    }
    throw Exception("GpuSolver: " + msg);
}

struct QuantityBinding {
    QuantityId id;
    int q;
};

// first-order scalar quantities with their derivative
const QuantityBinding SCALARS_FIRST[] = { { QuantityId::DENSITY, SPHGPU_Q_DENSITY }, { QuantityId::ENERGY, SPHGPU_Q_ENERGY },
    { QuantityId::DAMAGE, SPHGPU_Q_DAMAGE } };
// zero-order scalar inputs
const QuantityBinding SCALARS_ZERO[] = { { QuantityId::MASS, SPHGPU_Q_MASS }, { QuantityId::PRESSURE, SPHGPU_Q_PRESSURE },
    { QuantityId::SOUND_SPEED, SPHGPU_Q_SOUND_SPEED }, { QuantityId::STRESS_REDUCING, SPHGPU_Q_STRESS_REDUCING },
    { QuantityId::EPS_MIN, SPHGPU_Q_EPS_MIN }, { QuantityId::M_ZERO, SPHGPU_Q_M_ZERO },
    { QuantityId::EXPLICIT_GROWTH, SPHGPU_Q_EXPLICIT_GROWTH },
    { QuantityId::INTERPARTICLE_SPACING_KERNEL, SPHGPU_Q_INTERPARTICLE_SPACING_KERNEL } };
const QuantityBinding INDICES[] = { { QuantityId::FLAG, SPHGPU_Q_FLAG }, { QuantityId::N_FLAWS, SPHGPU_Q_N_FLAWS } };

} // namespace

GpuSolver::GpuSolver(IScheduler& scheduler, const RunSettings& settings, const EquationHolder& eqs, const int device)
    : scheduler(scheduler)
    , settings(settings)
    , device(device) {
    kernel = Factory::getKernel<DIMENSIONS>(settings);
    equations += eqs;

    // Terms with a device implementation; anything else cannot be evaluated (no CPU fallback by design).
    Size known = 0;
    known += equations.contains<PressureForce>() ? 1 : 0;
    known += equations.contains<SolidStressForce>() ? 1 : 0;
    known += equations.contains<ContinuityEquation>() ? 1 : 0;
    known += equations.contains<StandardAV>() ? 1 : 0;
    known += equations.contains<BalsaraSwitch<StandardAV>>() ? 1 : 0;
    known += equations.contains<XSph>() ? 1 : 0;
    known += equations.contains<DeltaSph::DensityDiffusion>() ? 1 : 0;
    known += equations.contains<DeltaSph::VelocityDiffusion>() ? 1 : 0;
    known += equations.contains<StressAV>() ? 1 : 0;
    known += equations.contains<AdaptiveSmoothingLength>() ? 1 : 0;
    known += equations.contains<ConstSmoothingLength>() ? 1 : 0;
    if (known != equations.getTermCnt()) {
        throw InvalidSetup("GpuSolver: the equation set contains a term without a GPU implementation");
    }
    if (!equations.contains<PressureForce>() || !equations.contains<ContinuityEquation>() ||
        !(equations.contains<StandardAV>() || equations.contains<BalsaraSwitch<StandardAV>>())) {
        throw InvalidSetup("GpuSolver needs PressureForce, ContinuityEquation and StandardAV, plain or with the Balsara switch");
    }
    if (equations.contains<XSph>() && equations.contains<BalsaraSwitch<StandardAV>>()) {
        throw InvalidSetup("GpuSolver: the XSph term together with the Balsara switch is not implemented on the device");
    }
    // the delta-SPH terms come as a pair (StandardSets.cpp:64-67) and are evaluated as one device variant
    if (equations.contains<DeltaSph::DensityDiffusion>() != equations.contains<DeltaSph::VelocityDiffusion>()) {
        throw InvalidSetup("GpuSolver: DeltaSph::DensityDiffusion and DeltaSph::VelocityDiffusion are only implemented together");
    }
    if (equations.contains<DeltaSph::DensityDiffusion>() &&
        (equations.contains<XSph>() || equations.contains<BalsaraSwitch<StandardAV>>())) {
        throw InvalidSetup("GpuSolver: the delta-SPH terms together with XSph or the Balsara switch are not implemented on the device");
    }
    if (equations.contains<StressAV>()) {
        if (equations.contains<XSph>() || equations.contains<BalsaraSwitch<StandardAV>>() || equations.contains<DeltaSph::DensityDiffusion>()) {
            throw InvalidSetup("GpuSolver: the artificial stress together with XSph, the Balsara switch or the delta-SPH terms is not implemented on the device");
        }
        if (!equations.contains<SolidStressForce>()) {
            throw InvalidSetup("GpuSolver: the artificial stress needs SolidStressForce (it is built from the deviatoric stress)");
        }
    }
    if (equations.contains<BalsaraSwitch<StandardAV>>() && settings.get<bool>(RunSettingsId::SPH_AV_BALSARA_STORE)) {
        throw InvalidSetup("GpuSolver: SPH_AV_BALSARA_STORE (the AV_BALSARA output quantity) is not implemented on the device");
    }
    // same check as AsymmetricSolver::sanityCheck (AsymmetricSolver.cpp:228-234)
    if (!equations.contains<AdaptiveSmoothingLength>() && !equations.contains<ConstSmoothingLength>()) {
        throw InvalidSetup("No solver of smoothing length specified; add either ConstSmootingLength or "
                           "AdaptiveSmootingLength into the list of equations");
    }
    // settings the device path would silently ignore are rejected instead
    if (settings.get<Float>(RunSettingsId::TIMESTEPPING_MEAN_POWER) > -1.e3_f) {
        throw InvalidSetup("GpuSolver: TIMESTEPPING_MEAN_POWER selects a generalised mean of the time steps "
                           "(TimeStepCriterion.cpp:117-148); the device criteria implement the minimum only");
    }
}

GpuSolver::GpuSolver(IScheduler& scheduler, const RunSettings& settings, const EquationHolder& eqs, AutoPtr<IBoundaryCondition>&& bc,
    const int device)
    : GpuSolver(scheduler, settings, eqs, device) {
    if (!bc || dynamic_cast<NullBoundaryCondition*>(&*bc)) {
        return;
    }
    // FrozenParticles itself (not WindTunnel, which derives from it and does more): its frozen bodies, domain and radius
    // are protected members; a pointer to member named through a derived class reads them without touching the reference
    struct Access : FrozenParticles {
        static std::set<Size> FrozenParticles::*flags() {
            return &Access::frozen;
        }
        static SharedPtr<IDomain> FrozenParticles::*dom() {
            return &Access::domain;
        }
        static Float FrozenParticles::*rad() {
            return &Access::radius;
        }
    };
    if (typeid(*bc) == typeid(KillEscapersBoundary)) {
        hostBc = std::move(bc);
        return;
    }
    FrozenParticles* frozen = dynamic_cast<FrozenParticles*>(&*bc);
    if (!frozen || typeid(*frozen) != typeid(FrozenParticles)) {
        throw InvalidSetup("GpuSolver: this boundary condition is not implemented (NullBoundaryCondition, FrozenParticles and "
                           "KillEscapersBoundary are)");
    }
    for (const Size flag : frozen->*Access::flags()) {
        if (flag >= 64) {
            throw InvalidSetup("GpuSolver: FrozenParticles on the device freezes body flags below 64");
        }
        frozenMask |= 1ull << flag;
    }
    const SharedPtr<IDomain>& domain = frozen->*Access::dom();
    if (domain) {
        const SphericalDomain* sphere = dynamic_cast<const SphericalDomain*>(&*domain);
        if (!sphere) {
            throw InvalidSetup("GpuSolver: FrozenParticles on the device needs a SphericalDomain");
        }
        frozenDomain = true;
        const Vector c = sphere->getCenter();
        frozenCenter[0] = c[X];
        frozenCenter[1] = c[Y];
        frozenCenter[2] = c[Z];
        frozenRadius = 0.5 * sphere->getBoundingBox().size()[X];
        frozenFreezeRadius = frozen->*Access::rad();
    }
    frozenActive = frozenMask != 0 || frozenDomain;
}

void GpuSolver::configureVariant() {
    check(sphgpu_set_variant(ctx, pairVariant));
}

void GpuSolver::configureFrozen() {
    sphgpu_frozen f{};
    f.flag_mask = frozenMask;
    f.has_domain = frozenDomain ? 1 : 0;
    for (int k = 0; k < 3; ++k) {
        f.center[k] = frozenCenter[k];
    }
    f.radius = frozenRadius;
    f.freeze_radius = frozenFreezeRadius;
    check(sphgpu_set_frozen(ctx, &f));
}

/// Attached to the Storage (Storage::setUserData): Storage::remove tells us that particle indices changed, so the device
/// mirror is stale. A device state that is newer than the Storage cannot be reconciled with a host-side removal.
class GpuSolver::DeviceMirror : public IStorageUserData {
private:
    GpuSolver& owner;

public:
    explicit DeviceMirror(GpuSolver& owner)
        : owner(owner) {}

    virtual void remove(ArrayView<const Size> idxs) override {
        if (idxs.empty()) {
            return;
        }
        if (owner.hostStale) {
            throw Exception("GpuSolver: particles removed from the Storage while the device holds the newer state; call "
                            "GpuPredictorCorrector::syncToHost() first");
        }
        owner.generation++;
        owner.staticGeneration = Size(-1);
    }
};

void GpuSolver::attachMirror(Storage& storage) {
    if (!storage.getUserData()) { // (a slot used by someone else is left alone; count changes are still detected)
        storage.setUserData(makeShared<DeviceMirror>(*this));
    }
}

GpuSolver::~GpuSolver() {
    sphgpu_destroy(ctx);
}

void GpuSolver::create(Storage& storage, IMaterial& material) const {
    storage.insert<Size>(QuantityId::NEIGHBOR_CNT, OrderEnum::ZERO, 0);
    equations.create(storage, material);
}

sphgpu_ctx* GpuSolver::context(const Storage& storage) {
    const Size n = storage.getParticleCnt();
    if (ctx && ctxParticleCnt == n) {
        return ctx;
    }
    sphgpu_destroy(ctx);
    ctx = nullptr;
    generation++;
    staticGeneration = Size(-1);

    sphgpu_config cfg{};
    cfg.abi_version = SPHGPU_ABI_VERSION;
    cfg.forces = SPHGPU_FORCE_PRESSURE | (equations.contains<SolidStressForce>() ? SPHGPU_FORCE_SOLID_STRESS : 0);
    const Flags<SmoothingLengthEnum> hflags =
        settings.getFlags<SmoothingLengthEnum>(RunSettingsId::SPH_ADAPTIVE_SMOOTHING_LENGTH);
    cfg.flags = 0;
    if (equations.contains<SolidStressForce>() && settings.get<bool>(RunSettingsId::SPH_STRAIN_RATE_CORRECTION_TENSOR)) {
        cfg.flags |= SPHGPU_FLAG_CORRECTION_TENSOR;
    }
    if (settings.get<bool>(RunSettingsId::SPH_SUM_ONLY_UNDAMAGED)) {
        cfg.flags |= SPHGPU_FLAG_SUM_ONLY_UNDAMAGED;
    }
    if (equations.contains<BalsaraSwitch<StandardAV>>()) {
        cfg.flags |= SPHGPU_FLAG_BALSARA;
    }
    if (equations.contains<XSph>()) {
        cfg.flags |= SPHGPU_FLAG_XSPH;
    }
    if (equations.contains<DeltaSph::DensityDiffusion>()) {
        cfg.flags |= SPHGPU_FLAG_DELTASPH;
    }
    if (equations.contains<StressAV>()) {
        cfg.flags |= SPHGPU_FLAG_STRESS_AV;
    }
    if (equations.contains<AdaptiveSmoothingLength>()) {
        cfg.flags |= SPHGPU_FLAG_ADAPTIVE_H;
        if (hflags.has(SmoothingLengthEnum::SOUND_SPEED_ENFORCING)) {
            cfg.flags |= SPHGPU_FLAG_SOUND_SPEED_ENFORCING;
        }
    }
    cfg.discretization = uint32_t(settings.get<DiscretizationEnum>(RunSettingsId::SPH_DISCRETIZATION));
    cfg.continuity_mode = uint32_t(settings.get<ContinuityEnum>(RunSettingsId::SPH_CONTINUITY_MODE));

    // kernel LUT exactly as the reference tabulates it (LutKernel, core/sph/kernel/Kernel.h:85-101)
    const Size entries = 40000;
    const Float qSqrToIdx = Float(entries) / sqr(kernel.radius());
    std::vector<double> lutGrad(entries + 1, 0.), lutValue(entries + 1, 0.);
    for (Size i = 0; i < entries; ++i) {
        lutGrad[i] = kernel.gradImpl(Float(i) / qSqrToIdx);
        lutValue[i] = kernel.valueImpl(Float(i) / qSqrToIdx);
    }
    cfg.lut_entries = entries;
    cfg.lut_grad = lutGrad.data();
    cfg.lut_value = lutValue.data();
    cfg.kernel_radius = kernel.radius();
    cfg.av_alpha = settings.get<Float>(RunSettingsId::SPH_AV_ALPHA);
    cfg.av_beta = settings.get<Float>(RunSettingsId::SPH_AV_BETA);
    const Interval hRange = settings.get<Interval>(RunSettingsId::SPH_SMOOTHING_LENGTH_RANGE);
    cfg.h_min = hRange.lower();
    cfg.h_max = hRange.upper();
    cfg.neigh_enforcing = settings.get<Float>(RunSettingsId::SPH_NEIGHBOR_ENFORCING);
    const Interval nRange = settings.get<Interval>(RunSettingsId::SPH_NEIGHBOR_RANGE);
    cfg.neigh_lower = nRange.lower();
    cfg.neigh_upper = nRange.upper();
    cfg.criteria = uint32_t(settings.getFlags<TimeStepCriterionEnum>(RunSettingsId::TIMESTEPPING_CRITERION).value());
    cfg.courant = settings.get<Float>(RunSettingsId::TIMESTEPPING_COURANT_NUMBER);
    cfg.derivative_factor = settings.get<Float>(RunSettingsId::TIMESTEPPING_DERIVATIVE_FACTOR);
    cfg.divergence_factor = settings.get<Float>(RunSettingsId::TIMESTEPPING_DIVERGENCE_FACTOR);
    cfg.max_change = settings.get<Float>(RunSettingsId::TIMESTEPPING_MAX_INCREASE);

    std::vector<sphgpu_material> mats(storage.getMaterialCnt());
    for (Size i = 0; i < storage.getMaterialCnt(); ++i) {
        MaterialView mat = storage.getMaterial(i);
        const BodySettings& b = mat->getParams();
        sphgpu_material& m = mats[i];
        m = sphgpu_material{};
        m.begin = *mat.sequence().begin();
        m.end = *mat.sequence().end();
        m.eos = uint32_t(b.get<EosEnum>(BodySettingsId::EOS));
        m.yielding = uint32_t(b.get<YieldingEnum>(BodySettingsId::RHEOLOGY_YIELDING));
        m.fracture = (m.yielding == SPHGPU_YIELD_NONE) ? uint32_t(SPHGPU_FRACTURE_NONE)
                                                       : uint32_t(b.get<FractureEnum>(BodySettingsId::RHEOLOGY_DAMAGE));
        m.til_u0 = b.get<Float>(BodySettingsId::TILLOTSON_SUBLIMATION);
        m.til_uiv = b.get<Float>(BodySettingsId::TILLOTSON_ENERGY_IV);
        m.til_ucv = b.get<Float>(BodySettingsId::TILLOTSON_ENERGY_CV);
        m.til_a = b.get<Float>(BodySettingsId::TILLOTSON_SMALL_A);
        m.til_b = b.get<Float>(BodySettingsId::TILLOTSON_SMALL_B);
        m.rho0 = b.get<Float>(BodySettingsId::DENSITY);
        m.til_A = b.get<Float>(BodySettingsId::BULK_MODULUS);
        m.til_B = b.get<Float>(BodySettingsId::TILLOTSON_NONLINEAR_B);
        m.til_alpha = b.get<Float>(BodySettingsId::TILLOTSON_ALPHA);
        m.til_beta = b.get<Float>(BodySettingsId::TILLOTSON_BETA);
        m.gamma = b.get<Float>(BodySettingsId::ADIABATIC_INDEX);
        m.shear_modulus = b.get<Float>(BodySettingsId::SHEAR_MODULUS);
        m.elasticity_limit = b.get<Float>(BodySettingsId::ELASTICITY_LIMIT);
        m.melt_energy = b.get<Float>(BodySettingsId::MELT_ENERGY);
        m.young_modulus = b.get<Float>(BodySettingsId::YOUNG_MODULUS);
        const Interval rhoRange = mat->range(QuantityId::DENSITY), uRange = mat->range(QuantityId::ENERGY),
                       dRange = mat->range(QuantityId::DAMAGE);
        m.rho_min = rhoRange.lower();
        m.rho_max = rhoRange.upper();
        m.u_min = uRange.lower();
        m.u_max = uRange.upper();
        m.d_min = dRange.lower();
        m.d_max = dRange.upper();
        m.rho_small = mat->minimal(QuantityId::DENSITY);
        m.u_small = mat->minimal(QuantityId::ENERGY);
        m.d_small = mat->minimal(QuantityId::DAMAGE);
        m.s_small = mat->minimal(QuantityId::DEVIATORIC_STRESS);
        if (storage.has(QuantityId::DEVIATORIC_STRESS) && mat->range(QuantityId::DEVIATORIC_STRESS) != Interval::unbounded()) {
            throw InvalidSetup("GpuSolver: a bounded range of the deviatoric stress is not implemented on the device");
        }
    }
    check(sphgpu_create(&cfg, mats.data(), uint32_t(mats.size()), n, n, device, &ctx));
    ctxParticleCnt = n;
    if (cfg.flags & SPHGPU_FLAG_XSPH) {
        check(sphgpu_set_xsph_epsilon(ctx, settings.get<Float>(RunSettingsId::SPH_XSPH_EPSILON)));
    }
    if (cfg.flags & SPHGPU_FLAG_STRESS_AV) {
        check(sphgpu_set_stress_av(ctx, settings.get<Float>(RunSettingsId::SPH_AV_STRESS_EXPONENT),
            settings.get<Float>(RunSettingsId::SPH_AV_STRESS_FACTOR)));
    }
    if (cfg.flags & SPHGPU_FLAG_DELTASPH) {
        check(sphgpu_set_deltasph(ctx, settings.get<Float>(RunSettingsId::SPH_DENSITY_DIFFUSION_DELTA),
            settings.get<Float>(RunSettingsId::SPH_VELOCITY_DIFFUSION_ALPHA)));
    }
    if (deviceGravity) {
        this->configureGravity();
    }
    if (frozenActive) {
        this->configureFrozen();
    }
    if (pairVariant != 0) {
        this->configureVariant();
    }
    return ctx;
}

namespace {

/// The table GravityLutKernel holds for the run's SPH kernel (Factory::getGravityKernel, Factory.cpp:712-735): LutKernel<3>
/// built from the exact gravity kernel; its interpolation returns the entries themselves at the nodes, and the last
/// node (the Newtonian value at the kernel's edge) comes from the exact kernel.
template <typename TGravityKernel>
void tabulateGravityKernel(const TGravityKernel& exact, std::vector<double>& lut, double& radius) {
    const LutKernel<3> lk(exact);
    const Size entries = 40000;
    radius = lk.radius();
    const Float qSqrToIdx = Float(entries) / sqr(lk.radius());
    lut.resize(entries + 1);
    for (Size i = 0; i < entries; ++i) {
        lut[i] = lk.gradImpl(Float(i) / qSqrToIdx);
    }
    lut[entries] = exact.gradImpl(Float(entries) / qSqrToIdx);
}

bool gravityTable(const RunSettings& settings, std::vector<double>& lut, double& radius) {
    switch (settings.get<KernelEnum>(RunSettingsId::SPH_KERNEL)) {
    case KernelEnum::CUBIC_SPLINE:
        tabulateGravityKernel(GravityKernel<CubicSpline<3>>(), lut, radius);
        return true;
    case KernelEnum::THOMAS_COUCHMAN:
        tabulateGravityKernel(GravityKernel<ThomasCouchmanKernel<3>>(), lut, radius);
        return true;
    case KernelEnum::FOURTH_ORDER_SPLINE:
        tabulateGravityKernel(getAssociatedGravityKernel(FourthOrderSpline<3>()), lut, radius);
        return true;
    case KernelEnum::GAUSSIAN:
        tabulateGravityKernel(getAssociatedGravityKernel(Gaussian<3>()), lut, radius);
        return true;
    case KernelEnum::CORE_TRIANGLE:
        tabulateGravityKernel(getAssociatedGravityKernel(CoreTriangle()), lut, radius);
        return true;
    case KernelEnum::WENDLAND_C2:
        tabulateGravityKernel(getAssociatedGravityKernel(WendlandC2()), lut, radius);
        return true;
    case KernelEnum::WENDLAND_C4:
        tabulateGravityKernel(getAssociatedGravityKernel(WendlandC4()), lut, radius);
        return true;
    case KernelEnum::WENDLAND_C6:
        tabulateGravityKernel(getAssociatedGravityKernel(WendlandC6()), lut, radius);
        return true;
    default:
        return false;
    }
}

} // namespace

bool GpuSolver::enableDeviceGravity() {
    const GravityEnum id = settings.get<GravityEnum>(RunSettingsId::GRAVITY_SOLVER);
    const GravityKernelEnum kernelId = settings.get<GravityKernelEnum>(RunSettingsId::GRAVITY_KERNEL);
    if (id != GravityEnum::BARNES_HUT && id != GravityEnum::BRUTE_FORCE) {
        return false;
    }
    if (kernelId != GravityKernelEnum::POINT_PARTICLES && kernelId != GravityKernelEnum::SPH_KERNEL) {
        return false;
    }
    if (settings.get<BoundaryEnum>(RunSettingsId::DOMAIN_BOUNDARY) == BoundaryEnum::SYMMETRIC ||
        settings.get<Float>(RunSettingsId::GRAVITY_RECOMPUTATION_PERIOD) > 0._f) {
        return false; // SymmetricGravity / CachedGravity wrappers (Factory.cpp:401-410)
    }
    if (id == GravityEnum::BARNES_HUT) {
        const int order = settings.get<int>(RunSettingsId::GRAVITY_MULTIPOLE_ORDER);
        if (settings.get<Float>(RunSettingsId::GRAVITY_OPENING_ANGLE) > 1._f || (order != 0 && order != 2 && order != 3)) {
            return false;
        }
    }
    if (kernelId == GravityKernelEnum::SPH_KERNEL) {
        std::vector<double> lut;
        double radius;
        if (!gravityTable(settings, lut, radius)) {
            return false;
        }
    }
    deviceGravity = true;
    if (ctx) {
        this->configureGravity();
    }
    return true;
}

void GpuSolver::configureGravity() {
    sphgpu_gravity g{};
    const bool brute = settings.get<GravityEnum>(RunSettingsId::GRAVITY_SOLVER) == GravityEnum::BRUTE_FORCE;
    g.opening_angle = brute ? 0. : double(settings.get<Float>(RunSettingsId::GRAVITY_OPENING_ANGLE));
    g.multipole_order = brute ? 3 : settings.get<int>(RunSettingsId::GRAVITY_MULTIPOLE_ORDER);
    g.leaf_size = uint32_t(settings.get<int>(RunSettingsId::FINDER_LEAF_SIZE));
    // Factory::getGravity hands GRAVITY_CONSTANT to BarnesHut only; BruteForceGravity is built with its default,
    // Constants::gravity (Factory.cpp:384-395). The drop-in reproduces the reference as it is.
    g.constant = brute ? double(Constants::gravity) : double(settings.get<Float>(RunSettingsId::GRAVITY_CONSTANT));
    std::vector<double> lut;
    double radius = 0.;
    if (settings.get<GravityKernelEnum>(RunSettingsId::GRAVITY_KERNEL) == GravityKernelEnum::SPH_KERNEL) {
        gravityTable(settings, lut, radius);
        g.lut_grad = lut.data();
        g.lut_entries = uint32_t(lut.size() - 1);
    }
    g.kernel_radius = radius;
    check(sphgpu_gravity_configure(ctx, &g));
}

void GpuSolver::uploadQuantities(const Storage& storage, const bool derivatives) {
    sphgpu_ctx* c = this->context(storage);
    const Size n = storage.getParticleCnt();
    const int L = SPHGPU_LAYOUT_OPENSPH;
    check(sphgpu_upload(c, SPHGPU_Q_POSITION, 0, L, &storage.getValue<Vector>(QuantityId::POSITION)[0], 0, n));
    check(sphgpu_upload(c, SPHGPU_Q_POSITION, 1, L, &storage.getDt<Vector>(QuantityId::POSITION)[0], 0, n));
    if (derivatives) {
        check(sphgpu_upload(c, SPHGPU_Q_POSITION, 2, L, &storage.getD2t<Vector>(QuantityId::POSITION)[0], 0, n));
    }
    // per-particle constants (and p, cs, reduce, which the device recomputes or keeps): once per device context
    const bool statics = staticGeneration != generation;
    for (const QuantityBinding& b : SCALARS_ZERO) {
        if (statics && storage.has(b.id)) {
            check(sphgpu_upload(c, b.q, 0, L, &storage.getValue<Float>(b.id)[0], 0, n));
        }
    }
    for (const QuantityBinding& b : SCALARS_FIRST) {
        if (storage.has(b.id)) {
            check(sphgpu_upload(c, b.q, 0, L, &storage.getValue<Float>(b.id)[0], 0, n));
            if (derivatives) {
                check(sphgpu_upload(c, b.q, 1, L, &storage.getDt<Float>(b.id)[0], 0, n));
            }
        }
    }
    for (const QuantityBinding& b : INDICES) {
        if (statics && storage.has(b.id)) {
            check(sphgpu_upload(c, b.q, 0, L, &storage.getValue<Size>(b.id)[0], 0, n));
        }
    }
    if (storage.has(QuantityId::DELTASPH_DENSITY_GRADIENT)) {
        // the gradient the previous evaluation stored: input of this one's density diffusion (DeltaSph.h:71-74)
        check(sphgpu_upload(c, SPHGPU_Q_DELTASPH_DENSITY_GRADIENT, 0, L,
            &storage.getValue<Vector>(QuantityId::DELTASPH_DENSITY_GRADIENT)[0], 0, n));
    }
    if (storage.has(QuantityId::XSPH_VELOCITIES)) {
        // the correction the previous evaluation left in the velocities (XSph::initialize takes it out again, XSph.h:69-79)
        check(sphgpu_upload(c, SPHGPU_Q_XSPH_VELOCITIES, 0, L, &storage.getValue<Vector>(QuantityId::XSPH_VELOCITIES)[0], 0, n));
    }
    if (storage.has(QuantityId::VELOCITY_ROTATION)) {
        // the Balsara factor is built from the divergence and rotation of the previous evaluation (Balsara.h:76-80)
        check(sphgpu_upload(c, SPHGPU_Q_VELOCITY_ROTATION, 0, L, &storage.getValue<Vector>(QuantityId::VELOCITY_ROTATION)[0], 0, n));
        check(sphgpu_upload(c, SPHGPU_Q_VELOCITY_DIVERGENCE, 0, L, &storage.getValue<Float>(QuantityId::VELOCITY_DIVERGENCE)[0], 0, n));
    }
    if (storage.has(QuantityId::DEVIATORIC_STRESS)) {
        check(sphgpu_upload(c, SPHGPU_Q_DEVIATORIC_STRESS, 0, L, &storage.getValue<TracelessTensor>(QuantityId::DEVIATORIC_STRESS)[0], 0, n));
        if (derivatives) {
            check(sphgpu_upload(c, SPHGPU_Q_DEVIATORIC_STRESS, 1, L, &storage.getDt<TracelessTensor>(QuantityId::DEVIATORIC_STRESS)[0], 0, n));
        }
    }
    staticGeneration = generation;
}

void GpuSolver::downloadQuantities(Storage& storage, const bool stateToo) {
    sphgpu_ctx* c = this->context(storage);
    const Size n = storage.getParticleCnt();
    const int L = SPHGPU_LAYOUT_OPENSPH;
    ArrayView<Vector> r, v, dv;
    tie(r, v, dv) = storage.getAll<Vector>(QuantityId::POSITION);
    // h may have been clamped (AdaptiveSmoothingLength::initialize) and dh/dt is a result of the evaluation
    check(sphgpu_download(c, SPHGPU_Q_POSITION, 0, L, &r[0], 0, n));
    check(sphgpu_download(c, SPHGPU_Q_POSITION, 1, L, &v[0], 0, n));
    check(sphgpu_download(c, SPHGPU_Q_POSITION, 2, L, &dv[0], 0, n));
    // p, cs, reduce and S are (re)computed by material->initialize
    for (const QuantityBinding& b : { QuantityBinding{ QuantityId::PRESSURE, SPHGPU_Q_PRESSURE },
             QuantityBinding{ QuantityId::SOUND_SPEED, SPHGPU_Q_SOUND_SPEED },
             QuantityBinding{ QuantityId::STRESS_REDUCING, SPHGPU_Q_STRESS_REDUCING },
             QuantityBinding{ QuantityId::VELOCITY_DIVERGENCE, SPHGPU_Q_VELOCITY_DIVERGENCE } }) {
        if (storage.has(b.id)) {
            check(sphgpu_download(c, b.q, 0, L, &storage.getValue<Float>(b.id)[0], 0, n));
        }
    }
    for (const QuantityBinding& b : SCALARS_FIRST) {
        if (storage.has(b.id)) {
            if (stateToo) {
                check(sphgpu_download(c, b.q, 0, L, &storage.getValue<Float>(b.id)[0], 0, n));
            }
            check(sphgpu_download(c, b.q, 1, L, &storage.getDt<Float>(b.id)[0], 0, n));
        }
    }
    if (storage.has(QuantityId::DEVIATORIC_STRESS)) {
        check(sphgpu_download(c, SPHGPU_Q_DEVIATORIC_STRESS, 0, L, &storage.getValue<TracelessTensor>(QuantityId::DEVIATORIC_STRESS)[0], 0, n));
        check(sphgpu_download(c, SPHGPU_Q_DEVIATORIC_STRESS, 1, L, &storage.getDt<TracelessTensor>(QuantityId::DEVIATORIC_STRESS)[0], 0, n));
    }
    if (storage.has(QuantityId::DELTASPH_DENSITY_GRADIENT)) {
        check(sphgpu_download(c, SPHGPU_Q_DELTASPH_DENSITY_GRADIENT, 0, L,
            &storage.getValue<Vector>(QuantityId::DELTASPH_DENSITY_GRADIENT)[0], 0, n));
    }
    if (storage.has(QuantityId::XSPH_VELOCITIES)) {
        check(sphgpu_download(c, SPHGPU_Q_XSPH_VELOCITIES, 0, L, &storage.getValue<Vector>(QuantityId::XSPH_VELOCITIES)[0], 0, n));
    }
    if (storage.has(QuantityId::VELOCITY_ROTATION)) {
        check(sphgpu_download(c, SPHGPU_Q_VELOCITY_ROTATION, 0, L, &storage.getValue<Vector>(QuantityId::VELOCITY_ROTATION)[0], 0, n));
    }
    if (storage.has(QuantityId::AV_STRESS)) { // StressAV::initialize leaves it in the Storage (Stress.cpp:91-109)
        check(sphgpu_download(c, SPHGPU_Q_AV_STRESS, 0, L, &storage.getValue<SymmetricTensor>(QuantityId::AV_STRESS)[0], 0, n));
    }
    if (storage.has(QuantityId::VELOCITY_GRADIENT)) {
        check(sphgpu_download(c, SPHGPU_Q_VELOCITY_GRADIENT, 0, L, &storage.getValue<SymmetricTensor>(QuantityId::VELOCITY_GRADIENT)[0], 0, n));
    }
    if (storage.has(QuantityId::STRAIN_RATE_CORRECTION_TENSOR)) {
        check(sphgpu_download(c, SPHGPU_Q_CORRECTION_TENSOR, 0, L,
            &storage.getValue<SymmetricTensor>(QuantityId::STRAIN_RATE_CORRECTION_TENSOR)[0], 0, n));
    }
    check(sphgpu_download(c, SPHGPU_Q_NEIGHBOR_CNT, 0, L, &storage.getValue<Size>(QuantityId::NEIGHBOR_CNT)[0], 0, n));
}

void GpuSolver::upload(const Storage& storage) {
    this->uploadQuantities(storage, true);
    hostStale = false;
}

void GpuSolver::download(Storage& storage) {
    this->downloadQuantities(storage, true);
    hostStale = false;
}

namespace {

/// The checker of the second pass of Post::findComponents with ESCAPE_VELOCITY (Analysis.cpp:97-112): components whose
/// relative speed is below their mutual escape speed belong together.
struct BoundComponents : public Post::IComponentChecker {
    ArrayView<const Vector> r, v;
    ArrayView<const Float> m;

    virtual bool belong(const Size i, const Size j) const override {
        const Float dv = getLength(v[i] - v[j]);
        const Float dr = getLength(r[i] - r[j]);
        return dv < sqrt(2._f * Constants::gravity * (m[i] + m[j]) / dr);
    }
};

} // namespace

Size GpuSolver::findComponents(const Storage& storage, const Float particleRadius, const Flags<Post::ComponentFlag> flags,
    Array<Size>& indices) {
    sphgpu_ctx* c = this->context(storage);
    const Size n = storage.getParticleCnt();
    ArrayView<const Vector> r = storage.getValue<Vector>(QuantityId::POSITION);
    check(sphgpu_upload(c, SPHGPU_Q_POSITION, 0, SPHGPU_LAYOUT_OPENSPH, &r[0], 0, n));
    uint32_t deviceFlags = 0;
    if (flags.has(Post::ComponentFlag::SEPARATE_BY_FLAG)) {
        check(sphgpu_upload(c, SPHGPU_Q_FLAG, 0, SPHGPU_LAYOUT_OPENSPH, &storage.getValue<Size>(QuantityId::FLAG)[0], 0, n));
        deviceFlags |= SPHGPU_COMPONENTS_SEPARATE_BY_FLAG;
    }
    indices.resize(n);
    static_assert(sizeof(Size) == sizeof(uint32_t), "Size is a 32-bit index");
    uint32_t count = 0;
    check(sphgpu_find_components(c, particleRadius, deviceFlags, reinterpret_cast<uint32_t*>(&indices[0]), &count, nullptr));
    Size componentCnt = count;

    if (flags.has(Post::ComponentFlag::ESCAPE_VELOCITY)) {
        // mass, barycentre, mean velocity and equivalent radius of every component, summed in particle order
        Array<Float> masses(componentCnt), volumes(componentCnt);
        Array<Vector> positions(componentCnt), velocities(componentCnt);
        masses.fill(0._f);
        volumes.fill(0._f);
        positions.fill(Vector(0._f));
        velocities.fill(Vector(0._f));
        ArrayView<const Float> m = storage.getValue<Float>(QuantityId::MASS);
        ArrayView<const Vector> v = storage.getDt<Vector>(QuantityId::POSITION);
        for (Size i = 0; i < n; ++i) {
            const Size k = indices[i];
            masses[k] += m[i];
            positions[k] += m[i] * r[i];
            velocities[k] += m[i] * v[i];
            volumes[k] += pow<3>(r[i][H]);
        }
        for (Size k = 0; k < componentCnt; ++k) {
            positions[k] /= masses[k];
            positions[k][H] = cbrt(3._f * volumes[k] / (4._f * PI));
            velocities[k] /= masses[k];
        }
        BoundComponents bound;
        bound.r = positions;
        bound.v = velocities;
        bound.m = masses;
        Storage bodies;
        bodies.insert<Vector>(QuantityId::POSITION, OrderEnum::ZERO, positions.clone());
        Array<Size> merged;
        componentCnt = Post::findComponents(bodies, 50._f, bound, merged);
        for (Size i = 0; i < n; ++i) {
            indices[i] = merged[indices[i]];
        }
    }
    if (flags.has(Post::ComponentFlag::SORT_BY_MASS)) {
        Array<Float> componentMass(componentCnt);
        componentMass.fill(0._f);
        ArrayView<const Float> m = storage.getValue<Float>(QuantityId::MASS);
        for (Size i = 0; i < n; ++i) {
            componentMass[indices[i]] += m[i];
        }
        Order byMass(componentCnt);
        byMass.shuffle([&componentMass](const Size a, const Size b) { return componentMass[a] > componentMass[b]; });
        const Order rank = byMass.getInverted();
        for (Size i = 0; i < n; ++i) {
            indices[i] = rank[indices[i]];
        }
    }
    // the context now holds these positions; the next integrate() uploads its own state anyway
    return componentCnt;
}

void GpuSolver::integrate(Storage& storage, Statistics& stats) {
    // ISolver contract: highest derivatives are zero on entry (ISolver.h:33-34); the device overwrites them, which is
    // what the reference's accumulate-into-zero amounts to.
    Timer timer;
    this->attachMirror(storage);
    if (hostBc) { // beforeLoop: bc->initialize (AsymmetricSolver.cpp:143); removed particles re-create the device mirror
        hostBc->initialize(storage);
    }
    this->uploadQuantities(storage, false);
    sphgpu_stats st{};
    const Float t = stats.getOr<Float>(StatisticsId::RUN_TIME, 0._f);
    check(sphgpu_integrate(ctx, t, &st));
    this->downloadQuantities(storage, false);
    if (hostBc) {
        hostBc->finalize(storage);
    }
    stats.set(StatisticsId::SPH_EVAL_TIME, int(timer.elapsed(TimerUnit::MILLISECOND))); // as AsymmetricSolver.cpp:92

    // neighbour statistics as AsymmetricSolver::afterLoop stores them (AsymmetricSolver.cpp:218-225)
    ArrayView<const Size> neighs = storage.getValue<Size>(QuantityId::NEIGHBOR_CNT);
    MinMaxMean neighsStats;
    for (Size i = 0; i < neighs.size(); ++i) {
        neighsStats.accumulate(neighs[i]);
    }
    stats.set(StatisticsId::NEIGHBOR_COUNT, neighsStats);
}

//-----------------------------------------------------------------------------------------------------------
// GpuGravitySolver
//-----------------------------------------------------------------------------------------------------------

GpuGravitySolver::GpuGravitySolver(IScheduler& scheduler,
    const RunSettings& settings,
    const EquationHolder& eqs,
    AutoPtr<IGravity>&& gravityImpl,
    const int device)
    : scheduler(scheduler)
    , settings(settings)
    , sph(scheduler, settings, eqs, device)
    , gravity(std::move(gravityImpl)) {
    if (!gravity && !sph.enableDeviceGravity()) {
        gravity = Factory::getGravity(settings);
    }
}

void GpuGravitySolver::ensureHostGravity() {
    if (!gravity) {
        gravity = Factory::getGravity(settings);
    }
}

GpuGravitySolver::~GpuGravitySolver() = default;

void GpuGravitySolver::create(Storage& storage, IMaterial& material) const {
    sph.create(storage, material);
}

void GpuGravitySolver::integrate(Storage& storage, Statistics& stats) {
    ArrayView<Attractor> attractors = storage.getAttractors();
    if (sph.hasDeviceGravity()) {
        // SPH and gravity in one device call (sphgpu_integrate adds the gravitational accelerations, api.cu)
        Timer timer;
        sph.integrate(storage, stats);
        sphgpu_gravity_stats gs{};
        if (sphgpu_gravity_last_stats(sph.context(storage), &gs) == SPHGPU_OK) {
            stats.set(StatisticsId::GRAVITY_EVAL_TIME, int(gs.gpu_ms));
            stats.set(StatisticsId::GRAVITY_NODES_APPROX, int(std::min<uint64_t>(gs.approximated, 0x7fffffff)));
            stats.set(StatisticsId::GRAVITY_NODES_EXACT, int(std::min<uint64_t>(gs.exact, 0x7fffffff)));
            stats.set(StatisticsId::GRAVITY_NODE_COUNT, int(gs.nodes));
        }
        stats.set(StatisticsId::SPH_EVAL_TIME, int(timer.elapsed(TimerUnit::MILLISECOND)));
        if (!attractors.empty()) {
            // attractors are few: their interaction with the particles stays on the host (BarnesHut::evalAttractors,
            // BarnesHut.cpp:101-124, needs only positions and masses, which the Storage holds)
            this->ensureHostGravity();
            gravity->build(scheduler, storage);
            ArrayView<Vector> dv = storage.getD2t<Vector>(QuantityId::POSITION);
            gravity->evalAttractors(scheduler, attractors, dv);
            for (Size i = 0; i < dv.size(); ++i) {
                dv[i][H] = 0._f;
            }
        }
        return;
    }
    // SPH part on the device; it overwrites the (zeroed) highest derivatives of the Storage
    Timer timer;
    sph.integrate(storage, stats);
    stats.set(StatisticsId::SPH_EVAL_TIME, int(timer.elapsed(TimerUnit::MILLISECOND)));

    // gravity as GravitySolver::loop does it (GravitySolver.cpp:68-88): build the tree, add the accelerations of the
    // particles and of the attractors to the acceleration buffer. The reference adds gravity first and the SPH terms
    // afterwards; the sum differs by rounding only.
    timer.restart();
    gravity->build(scheduler, storage);
    stats.set(StatisticsId::GRAVITY_BUILD_TIME, int(timer.elapsed(TimerUnit::MILLISECOND)));
    ArrayView<Vector> dv = storage.getD2t<Vector>(QuantityId::POSITION);
    timer.restart();
    gravity->evalSelfGravity(scheduler, dv, stats);
    stats.set(StatisticsId::GRAVITY_EVAL_TIME, int(timer.elapsed(TimerUnit::MILLISECOND)));
    gravity->evalAttractors(scheduler, attractors, dv);
    // the gravity kernels work on whole Vectors and leave a value in the H lane; the smoothing length is a first-order
    // quantity (AdaptiveSmoothingLength::finalize / ConstSmoothingLength::finalize zero it AFTER gravity in the
    // reference's order of operations, EquationTerm.cpp:381-383,427-434)
    for (Size i = 0; i < dv.size(); ++i) {
        dv[i][H] = 0._f;
    }
}

//-----------------------------------------------------------------------------------------------------------
// GpuPredictorCorrector
//-----------------------------------------------------------------------------------------------------------

/// Hands the time step computed on the device to ITimeStepping::step (TimeStepping.cpp:55-62).
class GpuPredictorCorrector::DeviceCriterion : public ITimeStepCriterion {
private:
    const TimeStep& step;

public:
    explicit DeviceCriterion(const TimeStep& step)
        : step(step) {}

    virtual TimeStep compute(IScheduler& UNUSED(scheduler),
        Storage& UNUSED(storage),
        Float UNUSED(maxStep),
        Statistics& UNUSED(stats),
        ArrayView<TimeStep> UNUSED(dts)) override {
        return step;
    }
};

GpuPredictorCorrector::GpuPredictorCorrector(const SharedPtr<Storage>& storage, const RunSettings& settings, GpuSolver& solver)
    : ITimeStepping(storage, settings, makeAuto<DeviceCriterion>(lastStep))
    , gpu(solver) {
    if (solver.hasHostBoundary()) {
        throw InvalidSetup("GpuPredictorCorrector: the solver's boundary condition works on the host Storage every step; use the "
                           "reference's PredictorCorrector with this GpuSolver");
    }
    lastStep.value = settings.get<Float>(RunSettingsId::TIMESTEPPING_INITIAL_TIMESTEP);
    lastStep.id = CriterionId::INITIAL_VALUE;
    // PredictorCorrector's constructor clears the derivatives before the first step (TimeStepping.cpp:280-281)
    storage->zeroHighestDerivatives(SEQUENTIAL);
}

void GpuPredictorCorrector::syncToHost() {
    if (gpu.isHostStale()) {
        gpu.download(*storage);
    }
}

void GpuPredictorCorrector::stepParticles(IScheduler& UNUSED(scheduler), ISolver& solver, Statistics& stats) {
    if (&solver != &gpu) {
        throw InvalidSetup("GpuPredictorCorrector must be used with the GpuSolver it was constructed with");
    }
    sphgpu_ctx* c = gpu.context(*storage);
    if (uploadedGeneration != gpu.getGeneration()) { // first step, or the device context / particle set was replaced
        gpu.upload(*storage);
        check(sphgpu_set_last_timestep(c, timeStep));
        uploadedGeneration = gpu.getGeneration();
    }
    sphgpu_stats st{};
    sphgpu_timestep ts{};
    const Float t = stats.getOr<Float>(StatisticsId::RUN_TIME, 0._f);
    check(sphgpu_step_pc(c, t, timeStep, maxTimeStep, &st, &ts));
    gpu.setHostStale(true);
    lastStep.value = ts.dt;
    lastStep.id = CriterionId(ts.criterion);

    // min / max / mean as AsymmetricSolver::afterLoop reports them (AsymmetricSolver.cpp:218-225); the per-particle counts
    // stay on the device, so the mean enters as SAMPLES copies of the value that, with min and max, averages to it
    MinMaxMean neighsStats;
    const Float lo = Float(st.neigh_min), hi = Float(st.neigh_max);
    neighsStats.accumulate(lo);
    neighsStats.accumulate(hi);
    constexpr int SAMPLES = 1022;
    const Float rest = clamp((Float(SAMPLES + 2) * Float(st.neigh_mean) - lo - hi) / Float(SAMPLES), lo, hi);
    for (int k = 0; k < SAMPLES; ++k) {
        neighsStats.accumulate(rest);
    }
    stats.set(StatisticsId::NEIGHBOR_COUNT, neighsStats);
    stats.set(StatisticsId::SPH_EVAL_TIME, int(st.gpu_ms));
}

GpuSyncOutput::GpuSyncOutput(AutoPtr<IOutput>&& inner, GpuPredictorCorrector& stepping)
    : IOutput(OutputFile())
    , inner(std::move(inner))
    , stepping(stepping) {}

Expected<Path> GpuSyncOutput::dump(const Storage& storage, const Statistics& stats) {
    stepping.syncToHost(); // the stepping object owns the same Storage (ITimeStepping::storage)
    return inner->dump(storage, stats);
}

NAMESPACE_SPH_END
