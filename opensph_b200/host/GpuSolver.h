#pragma once

/// \file GpuSolver.h
/// \brief Drop-in replacements for OpenSPH's AsymmetricSolver (ISolver) and PredictorCorrector (ITimeStepping)
///        that run the per-step SPH evaluation on a B200 through the C ABI of libsphgpu (include/sphgpu.h).
///
/// This is the reference-side binding: it is compiled AGAINST the reference's headers and linked with the
/// reference's core library plus libsphgpu.so. It owns no physics; it moves Storage arrays (the reference's own
/// AoS memory layout is understood by sphgpu_upload/download) and forwards to the C ABI.
///
///   reference class replaced                     file:line in the reference
///   AsymmetricSolver (ISolver)                   core/sph/solvers/AsymmetricSolver.h:124-129, .cpp:58-238
///   PredictorCorrector (ITimeStepping)           core/timestepping/TimeStepping.cpp:272-346
///   MultiCriterion (ITimeStepCriterion)          core/timestepping/TimeStepCriterion.cpp:389-419
///
/// Usage inside IRun::setUp (cf. examples/04_simple_collision/SimpleCollision.cpp and
/// core/run/jobs/SimulationJobs.cpp:173-180, where the protected members `solver` / `timeStepping` are assigned):
///
///     solver = makeAuto<GpuSolver>(*scheduler, settings, getStandardEquations(settings));
///     // optional: keep the whole step on the device
///     timeStepping = makeAuto<GpuPredictorCorrector>(storage, settings, static_cast<GpuSolver&>(*solver));

#include "gravity/IGravity.h"
#include "io/Output.h"
#include "post/Analysis.h"
#include "sph/boundary/Boundary.h"
#include "sph/equations/EquationTerm.h"
#include "system/Settings.h"
#include "sph/kernel/Kernel.h"
#include "timestepping/ISolver.h"
#include "timestepping/TimeStepCriterion.h"
#include "timestepping/TimeStepping.h"

struct sphgpu_ctx;

NAMESPACE_SPH_BEGIN

class GpuSolver : public ISolver {
private:
    IScheduler& scheduler;
    RunSettings settings;
    EquationHolder equations;
    LutKernel<DIMENSIONS> kernel;

    sphgpu_ctx* ctx = nullptr;
    Size ctxParticleCnt = 0;
    int device;

    /// Incremented whenever the device context is (re)created or the Storage reports removed particles: everything
    /// that mirrors device state (GpuPredictorCorrector, the static-quantity cache below) compares against it.
    Size generation = 0;
    /// Generation for which the per-particle constants (mass, flaw parameters, flags) are already on the device;
    /// integrate() then moves only the time-dependent quantities.
    Size staticGeneration = Size(-1);

    /// True while the device holds a newer state than the Storage (after GpuPredictorCorrector steps).
    bool hostStale = false;

    /// 0 = asymmetric kernels (default), 4 = the symmetric formulation of SymmetricSolver (sphgpu_set_variant)
    int pairVariant = 0;

    /// KillEscapersBoundary handed to the constructor: it only removes particles from the Storage
    /// (Boundary.cpp:445-454), which it does on the host at the start of integrate(); the device mirror follows through
    /// IStorageUserData. Needs the Storage every call, so it cannot be combined with GpuPredictorCorrector.
    AutoPtr<IBoundaryCondition> hostBc;

    /// FrozenParticles boundary condition taken over from the reference object handed to the constructor
    bool frozenActive = false;
    unsigned long long frozenMask = 0;
    bool frozenDomain = false;
    double frozenCenter[3] = { 0., 0., 0. }, frozenRadius = 0., frozenFreezeRadius = 0.;
    void configureFrozen();
    void configureVariant();

    /// Self-gravity on the device (enableDeviceGravity): applied to every context this solver creates.
    bool deviceGravity = false;
    void configureGravity();

    class DeviceMirror; ///< IStorageUserData attached to the Storage: sees Storage::remove (Storage.h:126-133)
    friend class DeviceMirror;

public:
    /// \throw InvalidSetup if the equation set contains a term without a GPU implementation (there is no CPU
    ///        fallback), mirroring AsymmetricSolver::sanityCheck (AsymmetricSolver.cpp:228-238).
    GpuSolver(IScheduler& scheduler, const RunSettings& settings, const EquationHolder& eqs, const int device = 0);

    /// The signature Factory::getSolver uses (AsymmetricSolver.h:126-129). Boundary conditions with a device
    /// implementation: none (nullptr / NullBoundaryCondition) and FrozenParticles (core/sph/boundary/Boundary.h:162-197) with
    /// frozen bodies and / or a SphericalDomain -- its settings are read from the object and applied on the device after
    /// every evaluation (sphgpu_set_frozen) -- and KillEscapersBoundary (Boundary.h), whose removal of particles from the
    /// Storage runs on the host at the start of integrate(). Anything else throws InvalidSetup.
    GpuSolver(IScheduler& scheduler, const RunSettings& settings, const EquationHolder& eqs, AutoPtr<IBoundaryCondition>&& bc,
        const int device = 0);

    ~GpuSolver() override;

    /// ISolver::integrate: uploads the state, runs sphgpu_integrate, stores derivatives back into the Storage.
    virtual void integrate(Storage& storage, Statistics& stats) override;

    /// Same quantities as IAsymmetricSolver::create (AsymmetricSolver.cpp:98-102).
    virtual void create(Storage& storage, IMaterial& material) const override;

    /// Copies the whole time-dependent state host -> device / device -> host.
    void upload(const Storage& storage);
    void download(Storage& storage);

    /// Device context (created lazily for the storage's particle count and materials).
    sphgpu_ctx* context(const Storage& storage);

    /// Adds the self-gravity Factory::getGravity(settings) describes (core/system/Factory.cpp:361-413: BarnesHut with
    /// GRAVITY_OPENING_ANGLE / GRAVITY_MULTIPOLE_ORDER / FINDER_LEAF_SIZE or BruteForceGravity, GRAVITY_KERNEL point
    /// particles or the softening kernel of the SPH kernel, GRAVITY_CONSTANT) to every integrate() on the device, the way
    /// GravitySolver<TSphSolver>::loop composes them (core/sph/solvers/GravitySolver.cpp:64-99).
    /// \return false -- and nothing changes -- if the settings ask for something the device does not implement
    ///         (SphericalGravity, solid-sphere kernel, symmetric boundary, cached gravity, opening angle above 1).
    bool enableDeviceGravity();

    /// Evaluates every pair once and adds it to both particles, like SymmetricSolver<3>::loop
    /// (core/sph/solvers/SymmetricSolver.cpp:104-163: rank by smoothing length, findLowerRank, evalSymmetric), instead of
    /// the asymmetric kernels. Same results to rounding, several times slower; not with the correction tensor (the
    /// reference's SymmetricSolver rejects it as well), the Balsara switch, XSph or the delta-SPH terms (InvalidSetup at the first integrate).
    /// Setups with SolverEnum::SYMMETRIC_SOLVER are served by the asymmetric kernels unless this is switched on.
    void useSymmetricFormulation(const bool symmetric) {
        pairVariant = symmetric ? 4 : 0;
        if (ctx) {
            this->configureVariant();
        }
    }

    /// Post::findComponents (core/post/Analysis.cpp:115-217) with the flood over the particles on the device
    /// (sphgpu_find_components: same components, same numbering). ComponentFlag::ESCAPE_VELOCITY merges the components found --
    /// a problem of the size of their number -- through the reference's own public overload with a checker, SORT_BY_MASS
    /// renumbers them by mass, both on the host exactly like the reference does.
    Size findComponents(const Storage& storage, const Float particleRadius, const Flags<Post::ComponentFlag> flags, Array<Size>& indices);

    /// True if a boundary condition runs on the host inside integrate() (KillEscapersBoundary).
    bool hasHostBoundary() const {
        return bool(hostBc);
    }

    bool hasDeviceGravity() const {
        return deviceGravity;
    }

    bool isHostStale() const {
        return hostStale;
    }
    void setHostStale(const bool stale) {
        hostStale = stale;
    }

    /// Changes when the device context was re-created or particles were removed from the Storage.
    Size getGeneration() const {
        return generation;
    }

    /// The per-particle constants (mass, flaw parameters, flags) changed on the host: upload them again.
    void invalidateStatic() {
        staticGeneration = Size(-1);
    }

private:
    void attachMirror(Storage& storage);
    void uploadQuantities(const Storage& storage, const bool derivatives);
    void downloadQuantities(Storage& storage, const bool stateToo);
};

/// \brief GpuSolver plus self-gravity, the composition GravitySolver<TSphSolver> performs
///        (core/sph/solvers/GravitySolver.cpp:64-99, selected by ForceEnum::SELF_GRAVITY in Factory.cpp:300-312), so that
///        setups with self-gravity -- the GUI collision preset, SimulationJobs.cpp:209-223 -- run on the device.
///
/// By default gravity is evaluated ON THE DEVICE (GpuSolver::enableDeviceGravity: Barnes-Hut / brute force as the run
/// settings describe it, sphgpu_gravity_configure); no particle data crosses PCIe for it and GpuPredictorCorrector keeps
/// whole steps on the device. When the caller passes its own IGravity, when the settings ask for something the device does
/// not implement, or when the Storage holds attractors, the accelerations come from the reference's IGravity on the host
/// cores and are added to the acceleration buffer (the round-1 behaviour).
class GpuGravitySolver : public ISolver {
private:
    IScheduler& scheduler;
    RunSettings settings;
    GpuSolver sph;
    AutoPtr<IGravity> gravity; ///< host-side gravity; null while the device evaluates it

    void ensureHostGravity();

public:
    /// \param gravity Host-side gravity implementation to use instead of the device's; nullptr selects the device path
    ///        (falling back to Factory::getGravity(settings) on the host where the device cannot serve the settings).
    GpuGravitySolver(IScheduler& scheduler,
        const RunSettings& settings,
        const EquationHolder& eqs,
        AutoPtr<IGravity>&& gravity = nullptr,
        const int device = 0);

    /// True if the gravitational accelerations are computed by the device.
    bool gravityOnDevice() const {
        return sph.hasDeviceGravity();
    }

    ~GpuGravitySolver() override;

    virtual void integrate(Storage& storage, Statistics& stats) override;

    virtual void create(Storage& storage, IMaterial& material) const override;

    GpuSolver& sphSolver() {
        return sph;
    }
};

/// \brief PredictorCorrector whose whole step (predict, derivatives, correct, time-step criteria) runs on the device.
///
/// The Storage on the host is refreshed only when syncToHost() is called (before output->dump, log writers or
/// callbacks that read particle data); between calls no particle data crosses PCIe.
class GpuPredictorCorrector : public ITimeStepping {
private:
    GpuSolver& gpu;
    Size uploadedGeneration = Size(-1); ///< GpuSolver::getGeneration() the device state was uploaded for
    Float runTime = 0._f;

    class DeviceCriterion;
    TimeStep lastStep;

public:
    GpuPredictorCorrector(const SharedPtr<Storage>& storage, const RunSettings& settings, GpuSolver& solver);

    /// Brings the host Storage up to date with the device state.
    void syncToHost();

protected:
    virtual void stepParticles(IScheduler& scheduler, ISolver& solver, Statistics& stats) override;
};

/// \brief IOutput decorator for runs stepped by GpuPredictorCorrector: IRun::run calls output->dump(storage, stats) at the
///        output times (core/run/IRun.cpp:251-257); this brings the host Storage up to date first, then forwards.
///
///     output = makeAuto<GpuSyncOutput>(std::move(output), static_cast<GpuPredictorCorrector&>(*timeStepping));
class GpuSyncOutput : public IOutput {
private:
    AutoPtr<IOutput> inner;
    GpuPredictorCorrector& stepping;

public:
    GpuSyncOutput(AutoPtr<IOutput>&& inner, GpuPredictorCorrector& stepping);

    virtual Expected<Path> dump(const Storage& storage, const Statistics& stats) override;
};

NAMESPACE_SPH_END
