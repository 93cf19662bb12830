// Drop-in test (run on the GPU box): the reference's own AsymmetricSolver / PredictorCorrector next to GpuSolver /
// GpuPredictorCorrector on identical Storages, all through the reference's ISolver / ITimeStepping interfaces.
// Modelled on the reference's solver cross-check (core/sph/solvers/test/Solvers.cpp:178-216: two solvers, same
// storage, compare every quantity) and its Impact smoke test (core/sph/solvers/test/Impact.cpp:41-85).
//
// Built by opensph_b200/host/Makefile against /root/reference (headers + oracle/_ref/libopensph_core_strict.a) and
// libsphgpu.so; the binary goes to oracle/_ref/dropin_test because it embeds reference code.
#include "../../opensph_b200/host/GpuSolver.h"
#include "../../opensph_b200/host/GpuOutput.h"
#include "Sph.h"
#include "physics/Constants.h"
#include "sph/solvers/GravitySolver.h"
#include <cstdio>
#include <fstream>
#include <iterator>
#include <vector>
#include <random>

using namespace Sph;

namespace {

int failures = 0;

void expect(const bool cond, const char* what) {
    printf("%s  %s\n", cond ? "ok  " : "FAIL", what);
    if (!cond) {
        failures++;
    }
}

template <typename T, typename TGet>
double relErr(ArrayView<const T> a, ArrayView<const T> b, const int comps, const TGet& get) {
    double scale = 0.;
    for (Size i = 0; i < a.size(); ++i) {
        for (int k = 0; k < comps; ++k) {
            scale = std::max(scale, std::max(std::abs(get(a[i], k)), std::abs(get(b[i], k))));
        }
    }
    const double floorV = 1.e-4 * std::max(scale, 1.e-300);
    double err = 0.;
    for (Size i = 0; i < a.size(); ++i) {
        for (int k = 0; k < comps; ++k) {
            const double x = get(a[i], k), y = get(b[i], k);
            err = std::max(err, std::abs(x - y) / std::max(std::max(std::abs(x), std::abs(y)), floorV));
        }
    }
    return err;
}

double cmpVector(ArrayView<const Vector> a, ArrayView<const Vector> b, const int comps = 4) {
    return relErr<Vector>(a, b, comps, [](const Vector& v, int k) { return double(v[k]); });
}
double cmpFloat(ArrayView<const Float> a, ArrayView<const Float> b) {
    return relErr<Float>(a, b, 1, [](const Float& v, int) { return double(v); });
}
double cmpTraceless(ArrayView<const TracelessTensor> a, ArrayView<const TracelessTensor> b) {
    static const int I[5] = { 0, 1, 0, 0, 1 }, J[5] = { 0, 1, 1, 2, 2 };
    return relErr<TracelessTensor>(a, b, 5, [](const TracelessTensor& t, int k) { return double(t(I[k], J[k])); });
}
double cmpSymmetric(ArrayView<const SymmetricTensor> a, ArrayView<const SymmetricTensor> b) {
    static const int I[6] = { 0, 1, 2, 0, 0, 1 }, J[6] = { 0, 1, 2, 1, 2, 2 };
    return relErr<SymmetricTensor>(a, b, 6, [](const SymmetricTensor& t, int k) { return double(t(I[k], J[k])); });
}

/// Compares every quantity the solvers write; returns the largest relative error.
double compareStorages(const Storage& a, const Storage& b, const bool derivatives, const char* label) {
    double worst = 0.;
    auto note = [&](const char* name, const double e) {
        printf("    %-28s %.3e\n", name, e);
        worst = std::max(worst, e);
    };
    printf("  [%s]\n", label);
    note("POSITION", cmpVector(a.getValue<Vector>(QuantityId::POSITION), b.getValue<Vector>(QuantityId::POSITION)));
    note("POSITION dt", cmpVector(a.getDt<Vector>(QuantityId::POSITION), b.getDt<Vector>(QuantityId::POSITION)));
    if (derivatives) {
        note("POSITION d2t", cmpVector(a.getD2t<Vector>(QuantityId::POSITION), b.getD2t<Vector>(QuantityId::POSITION)));
    }
    for (QuantityId id : { QuantityId::DENSITY, QuantityId::ENERGY, QuantityId::DAMAGE }) {
        if (a.has(id)) {
            note(getMetadata(id).quantityName.toAscii(), cmpFloat(a.getValue<Float>(id), b.getValue<Float>(id)));
            if (derivatives) {
                note((getMetadata(id).quantityName + " dt").toAscii(), cmpFloat(a.getDt<Float>(id), b.getDt<Float>(id)));
            }
        }
    }
    for (QuantityId id : { QuantityId::PRESSURE, QuantityId::SOUND_SPEED, QuantityId::STRESS_REDUCING, QuantityId::VELOCITY_DIVERGENCE }) {
        if (a.has(id) && derivatives) {
            note(getMetadata(id).quantityName.toAscii(), cmpFloat(a.getValue<Float>(id), b.getValue<Float>(id)));
        }
    }
    if (a.has(QuantityId::DEVIATORIC_STRESS)) {
        note("DEVIATORIC_STRESS", cmpTraceless(a.getValue<TracelessTensor>(QuantityId::DEVIATORIC_STRESS), b.getValue<TracelessTensor>(QuantityId::DEVIATORIC_STRESS)));
        if (derivatives) {
            note("DEVIATORIC_STRESS dt", cmpTraceless(a.getDt<TracelessTensor>(QuantityId::DEVIATORIC_STRESS), b.getDt<TracelessTensor>(QuantityId::DEVIATORIC_STRESS)));
        }
    }
    if (derivatives && a.has(QuantityId::VELOCITY_GRADIENT)) {
        note("VELOCITY_GRADIENT", cmpSymmetric(a.getValue<SymmetricTensor>(QuantityId::VELOCITY_GRADIENT), b.getValue<SymmetricTensor>(QuantityId::VELOCITY_GRADIENT)));
    }
    if (derivatives && a.has(QuantityId::STRAIN_RATE_CORRECTION_TENSOR)) {
        note("CORRECTION_TENSOR", cmpSymmetric(a.getValue<SymmetricTensor>(QuantityId::STRAIN_RATE_CORRECTION_TENSOR), b.getValue<SymmetricTensor>(QuantityId::STRAIN_RATE_CORRECTION_TENSOR)));
    }
    return worst;
}

bool sameNeighbourCounts(const Storage& a, const Storage& b) {
    ArrayView<const Size> x = a.getValue<Size>(QuantityId::NEIGHBOR_CNT), y = b.getValue<Size>(QuantityId::NEIGHBOR_CNT);
    for (Size i = 0; i < x.size(); ++i) {
        if (x[i] != y[i]) {
            return false;
        }
    }
    return true;
}

RunSettings presetSettings() {
    RunSettings settings;
    settings.set(RunSettingsId::TIMESTEPPING_INTEGRATOR, TimesteppingEnum::PREDICTOR_CORRECTOR)
        .set(RunSettingsId::TIMESTEPPING_INITIAL_TIMESTEP, 0.01_f)
        .set(RunSettingsId::TIMESTEPPING_MAX_TIMESTEP, 10._f)
        .set(RunSettingsId::SPH_SOLVER_TYPE, SolverEnum::ASYMMETRIC_SOLVER)
        .set(RunSettingsId::SPH_SOLVER_FORCES, ForceEnum::PRESSURE | ForceEnum::SOLID_STRESS)
        .set(RunSettingsId::SPH_FINDER, FinderEnum::KD_TREE)
        .set(RunSettingsId::FINDER_LEAF_SIZE, 20)
        .set(RunSettingsId::SPH_ADAPTIVE_SMOOTHING_LENGTH, SmoothingLengthEnum::CONTINUITY_EQUATION)
        .set(RunSettingsId::SPH_STRAIN_RATE_CORRECTION_TENSOR, true)
        .set(RunSettingsId::TIMESTEPPING_CRITERION, TimeStepCriterionEnum::COURANT | TimeStepCriterionEnum::DIVERGENCE)
        .set(RunSettingsId::RUN_THREAD_CNT, 0);
    return settings;
}

/// Target + impactor (examples/04_simple_collision) with jittered state so that every branch is exercised.
SharedPtr<Storage> makeStorage(const RunSettings& settings, ISolver& creator, const Size n) {
    SharedPtr<Storage> storage = makeShared<Storage>();
    InitialConditions ic(settings);
    BodySettings body;
    body.set(BodySettingsId::PARTICLE_COUNT, int(n));
    ic.addMonolithicBody(*storage, SphericalDomain(Vector(0._f), 1.e5_f), body);
    body.set(BodySettingsId::PARTICLE_COUNT, int(std::max<Size>(n / 100, 10)));
    BodyView impactor = ic.addMonolithicBody(*storage, SphericalDomain(Vector(1.4e5_f, 0._f, 0._f), 2.e4_f), body);
    impactor.addVelocity(Vector(-5.e3_f, 0._f, 0._f));
    for (Size i = 0; i < storage->getMaterialCnt(); ++i) {
        creator.create(*storage, storage->getMaterial(i));
    }
    std::mt19937_64 gen(4321);
    std::uniform_real_distribution<double> uni(0., 1.);
    ArrayView<Vector> r, v, dv;
    tie(r, v, dv) = storage->getAll<Vector>(QuantityId::POSITION);
    ArrayView<Float> rho = storage->getValue<Float>(QuantityId::DENSITY), u = storage->getValue<Float>(QuantityId::ENERGY),
                     D = storage->getValue<Float>(QuantityId::DAMAGE);
    ArrayView<TracelessTensor> s = storage->getValue<TracelessTensor>(QuantityId::DEVIATORIC_STRESS);
    for (Size i = 0; i < r.size(); ++i) {
        const Float h = r[i][H];
        r[i] += Vector(0.2_f * h * (2 * uni(gen) - 1), 0.2_f * h * (2 * uni(gen) - 1), 0.2_f * h * (2 * uni(gen) - 1));
        r[i][H] = h * (0.9_f + 0.2_f * uni(gen));
        v[i] += Vector(30._f * (2 * uni(gen) - 1), 30._f * (2 * uni(gen) - 1), 30._f * (2 * uni(gen) - 1));
        v[i][H] = 0._f;
        rho[i] *= 0.95_f + 0.1_f * uni(gen);
        u[i] = uni(gen) < 0.7 ? 1.e5_f * uni(gen) : 8.e6_f * uni(gen);
        const Float a = 5.e8_f;
        s[i] = TracelessTensor(a * (2 * uni(gen) - 1), a * (2 * uni(gen) - 1), a * (2 * uni(gen) - 1), a * (2 * uni(gen) - 1), a * (2 * uni(gen) - 1));
        const Float c = uni(gen);
        D[i] = c < 0.5 ? 0._f : (c < 0.55 ? 1._f : uni(gen));
    }
    return storage;
}

} // namespace

int main(int argc, char** argv) {
    const Size n = argc > 1 ? Size(atoi(argv[1])) : 20000;
    const int steps = argc > 2 ? atoi(argv[2]) : 3;
    try {
        RunSettings settings = presetSettings();
        SharedPtr<IScheduler> scheduler = Factory::getScheduler(settings);
        const EquationHolder eqs = getStandardEquations(settings);

        AsymmetricSolver refSolver(*scheduler, settings, eqs);
        GpuSolver gpuSolver(*scheduler, settings, eqs);

        // ---- 1. one integrate() through ISolver ----
        SharedPtr<Storage> base = makeStorage(settings, refSolver, n);
        printf("particles: %u, materials: %u\n", unsigned(base->getParticleCnt()), unsigned(base->getMaterialCnt()));
        Storage a = base->clone(VisitorEnum::ALL_BUFFERS), b = base->clone(VisitorEnum::ALL_BUFFERS);
        Statistics statsA, statsB;
        statsA.set(StatisticsId::RUN_TIME, 0._f);
        statsB.set(StatisticsId::RUN_TIME, 0._f);
        a.zeroHighestDerivatives(*scheduler);
        b.zeroHighestDerivatives(*scheduler);
        refSolver.integrate(a, statsA);
        gpuSolver.integrate(b, statsB);
        expect(sameNeighbourCounts(a, b), "NEIGHBOR_CNT identical (AsymmetricSolver vs GpuSolver)");
        const MinMaxMean na = statsA.get<MinMaxMean>(StatisticsId::NEIGHBOR_COUNT), nb = statsB.get<MinMaxMean>(StatisticsId::NEIGHBOR_COUNT);
        expect(na.min() == nb.min() && na.max() == nb.max() && std::abs(na.mean() - nb.mean()) < 1.e-9, "StatisticsId::NEIGHBOR_COUNT identical");
        expect(compareStorages(a, b, true, "integrate()") <= 1.e-10, "all quantities within 1e-10 after ISolver::integrate");

        // ---- 2. reference PredictorCorrector driving GpuSolver (host integrator, device derivatives) ----
        SharedPtr<Storage> sa = makeShared<Storage>(base->clone(VisitorEnum::ALL_BUFFERS));
        SharedPtr<Storage> sb = makeShared<Storage>(base->clone(VisitorEnum::ALL_BUFFERS));
        SharedPtr<Storage> sc = makeShared<Storage>(base->clone(VisitorEnum::ALL_BUFFERS));
        AutoPtr<ITimeStepping> ta = Factory::getTimeStepping(settings, sa);
        AutoPtr<ITimeStepping> tb = Factory::getTimeStepping(settings, sb);
        GpuSolver gpuSolver2(*scheduler, settings, eqs);
        GpuPredictorCorrector tc(sc, settings, gpuSolver2);
        bool dtOk = true;
        for (int s = 0; s < steps; ++s) {
            ta->step(*scheduler, refSolver, statsA);
            tb->step(*scheduler, gpuSolver, statsB);
            Statistics statsC;
            statsC.set(StatisticsId::RUN_TIME, 0._f);
            tc.step(*scheduler, gpuSolver2, statsC);
            dtOk &= std::abs(ta->getTimeStep() - tb->getTimeStep()) <= 1.e-9 * ta->getTimeStep();
            dtOk &= std::abs(ta->getTimeStep() - tc.getTimeStep()) <= 1.e-9 * ta->getTimeStep();
            printf("  step %d: dt ref %.12e  host-PC+GpuSolver %.12e  GpuPredictorCorrector %.12e\n", s, double(ta->getTimeStep()),
                double(tb->getTimeStep()), double(tc.getTimeStep()));
        }
        expect(dtOk, "time steps agree to 1e-9");
        {
            // the statistics GpuPredictorCorrector reports without moving per-particle data (AsymmetricSolver.cpp:218-225)
            Statistics statsC;
            statsC.set(StatisticsId::RUN_TIME, 0._f);
            tc.step(*scheduler, gpuSolver2, statsC);
            ta->step(*scheduler, refSolver, statsA);
            const MinMaxMean ra = statsA.get<MinMaxMean>(StatisticsId::NEIGHBOR_COUNT), rc = statsC.get<MinMaxMean>(StatisticsId::NEIGHBOR_COUNT);
            printf("  NEIGHBOR_COUNT ref %g/%g/%.6f  device %g/%g/%.6f\n", double(ra.min()), double(ra.max()), double(ra.mean()), double(rc.min()),
                double(rc.max()), double(rc.mean()));
            expect(ra.min() == rc.min() && ra.max() == rc.max() && std::abs(ra.mean() - rc.mean()) < 1.e-6 * ra.mean(),
                "NEIGHBOR_COUNT min / max / mean of GpuPredictorCorrector equal the reference's");
            expect(statsC.has(StatisticsId::SPH_EVAL_TIME), "SPH_EVAL_TIME is reported");
            tb->step(*scheduler, gpuSolver, statsB);
            expect(statsB.has(StatisticsId::SPH_EVAL_TIME), "SPH_EVAL_TIME is reported by GpuSolver::integrate");
        }
        expect(compareStorages(*sa, *sb, false, "PredictorCorrector + GpuSolver") <= 1.e-9, "state after PC steps (host integrator + GpuSolver) within 1e-9");
        {
            // ---- 2b. snapshot straight from the device planes (the host Storage is stale here) against the reference's
            // BinaryOutput of the synchronised Storage: identical files, loadable by the reference's BinaryInput ----
            Statistics dumpStats;
            dumpStats.set(StatisticsId::RUN_TIME, 1.5_f);
            dumpStats.set(StatisticsId::TIMESTEP_VALUE, tc.getTimeStep());
            const Path pathGpu("dropin_gpu_dump.ssf"), pathRef("dropin_ref_dump.ssf");
            GpuBinaryOutput gpuOut(OutputFile(pathGpu), gpuSolver2);
            expect(gpuSolver2.isHostStale(), "the host Storage is stale when the device snapshot is written");
            Expected<Path> wa = gpuOut.dump(*sc, dumpStats);
            expect(bool(wa), "GpuBinaryOutput::dump succeeds");
            expect(gpuSolver2.isHostStale(), "GpuBinaryOutput does not synchronise the host Storage");
            tc.syncToHost();
            BinaryOutput refOut{ OutputFile(pathRef) };
            Expected<Path> wb = refOut.dump(*sc, dumpStats);
            expect(bool(wb), "BinaryOutput::dump succeeds");
            std::ifstream fa(pathGpu.native(), std::ios::binary), fb(pathRef.native(), std::ios::binary);
            std::vector<char> ba((std::istreambuf_iterator<char>(fa)), std::istreambuf_iterator<char>());
            std::vector<char> bb((std::istreambuf_iterator<char>(fb)), std::istreambuf_iterator<char>());
            bool same = ba.size() == bb.size() && !ba.empty();
            size_t firstDiff = 0;
            for (size_t k = 0; same && k < ba.size(); ++k) {
                if (ba[k] != bb[k] && !(k >= 68 && k < 84)) { // (bytes 69-84: build date of the writing translation unit)
                    same = false;
                    firstDiff = k;
                }
            }
            printf("  snapshot sizes: device %zu bytes, reference %zu bytes, first difference at %zu\n", ba.size(), bb.size(), firstDiff);
            expect(same, "snapshot written from the device planes is byte-identical to BinaryOutput's");
            BinaryInput input;
            Storage loaded;
            Statistics loadedStats;
            Outcome res = input.load(pathGpu, loaded, loadedStats);
            expect(bool(res), "the reference's BinaryInput loads the device snapshot");
            if (res) {
                expect(loaded.getParticleCnt() == sc->getParticleCnt() && loaded.getMaterialCnt() == sc->getMaterialCnt(),
                    "loaded snapshot has the same particles and materials");
                expect(compareStorages(loaded, *sc, true, "snapshot round trip") == 0., "loaded snapshot equals the Storage bit for bit");
            }
        }
        expect(compareStorages(*sa, *sc, false, "GpuPredictorCorrector") <= 1.e-9, "state after device-resident PC steps within 1e-9");

        // ---- 3. self-gravity: GravitySolver<AsymmetricSolver> (Factory.cpp:300-312) next to GpuGravitySolver ----
        {
            RunSettings gs = settings;
            gs.set(RunSettingsId::SPH_SOLVER_FORCES, ForceEnum::PRESSURE | ForceEnum::SOLID_STRESS | ForceEnum::SELF_GRAVITY)
                .set(RunSettingsId::GRAVITY_SOLVER, GravityEnum::BARNES_HUT)
                .set(RunSettingsId::GRAVITY_OPENING_ANGLE, 0.8_f)
                .set(RunSettingsId::GRAVITY_MULTIPOLE_ORDER, 3)
                // the test body is small: a constant large enough for gravity to matter next to the SPH accelerations
                .set(RunSettingsId::GRAVITY_CONSTANT, Float(Constants::gravity * 1.e4));
            const EquationHolder geqs = getStandardEquations(gs);
            auto integrateWith = [&](ISolver& solver, Storage& st) {
                st.zeroHighestDerivatives(*scheduler);
                Statistics ss;
                ss.set(StatisticsId::RUN_TIME, 0._f);
                solver.integrate(st, ss);
            };
            GravitySolver<AsymmetricSolver> refGravity(*scheduler, gs, geqs);
            Storage ga = base->clone(VisitorEnum::ALL_BUFFERS);
            integrateWith(refGravity, ga);
            const double gravShare = cmpVector(ga.getD2t<Vector>(QuantityId::POSITION), b.getD2t<Vector>(QuantityId::POSITION), 3);
            expect(gravShare > 1.e-3, "gravity contributes to the accelerations");

            // (a) host-side gravity handed in by the caller: the reference's own IGravity, identical tree
            {
                GpuGravitySolver hostGravity(*scheduler, gs, geqs, Factory::getGravity(gs));
                expect(!hostGravity.gravityOnDevice(), "a caller-supplied IGravity stays on the host");
                Storage gb = base->clone(VisitorEnum::ALL_BUFFERS);
                integrateWith(hostGravity, gb);
                expect(sameNeighbourCounts(ga, gb), "NEIGHBOR_CNT identical with self-gravity");
                expect(compareStorages(ga, gb, true, "integrate() with host Barnes-Hut") <= 1.e-10,
                    "all quantities within 1e-10 (GravitySolver<AsymmetricSolver> vs GpuGravitySolver with host gravity)");
            }
            // (b) gravity on the device, every pair exactly. (Factory::getGravity builds BruteForceGravity without
            // GRAVITY_CONSTANT, Factory.cpp:385, so the enlarged constant would not reach it; an opening angle of 1e-3
            // makes BarnesHut open every node instead -- the same exact pair sums on both sides.)
            {
                RunSettings bs = gs;
                bs.set(RunSettingsId::GRAVITY_OPENING_ANGLE, 1.e-3_f);
                GravitySolver<AsymmetricSolver> refBrute(*scheduler, bs, geqs);
                GpuGravitySolver gpuBrute(*scheduler, bs, geqs);
                expect(gpuBrute.gravityOnDevice(), "gravity runs on the device");
                {
                    RunSettings bf = gs;
                    bf.set(RunSettingsId::GRAVITY_SOLVER, GravityEnum::BRUTE_FORCE);
                    GpuGravitySolver gpuBf(*scheduler, bf, geqs);
                    expect(gpuBf.gravityOnDevice(), "brute-force gravity runs on the device");
                }
                Storage gr = base->clone(VisitorEnum::ALL_BUFFERS), gb = base->clone(VisitorEnum::ALL_BUFFERS);
                integrateWith(refBrute, gr);
                integrateWith(gpuBrute, gb);
                expect(compareStorages(gr, gb, true, "integrate() with exact gravity on the device") <= 1.e-10,
                    "all quantities within 1e-10 (GravitySolver<AsymmetricSolver> vs device gravity, exact sums)");
                // (c) device Barnes-Hut: a different tree, so it is measured like the reference's own BarnesHut tests
                // (core/gravity/test/BarnesHut.cpp): against the exact sums, and not worse than the reference's tree
                GpuGravitySolver gpuBh(*scheduler, gs, geqs);
                expect(gpuBh.gravityOnDevice(), "Barnes-Hut gravity runs on the device");
                Storage gd = base->clone(VisitorEnum::ALL_BUFFERS);
                integrateWith(gpuBh, gd);
                auto rms = [](ArrayView<const Vector> x, ArrayView<const Vector> y) {
                    double num = 0., den = 0.;
                    for (Size i = 0; i < x.size(); ++i) {
                        for (int k = 0; k < 3; ++k) {
                            num += sqr(x[i][k] - y[i][k]);
                            den += sqr(y[i][k]);
                        }
                    }
                    return std::sqrt(num / den);
                };
                const double errRef = rms(ga.getD2t<Vector>(QuantityId::POSITION), gr.getD2t<Vector>(QuantityId::POSITION));
                const double errGpu = rms(gd.getD2t<Vector>(QuantityId::POSITION), gr.getD2t<Vector>(QuantityId::POSITION));
                printf("  [Barnes-Hut, opening angle 0.8, octupole] rms error of the accelerations against the exact sums: reference %.3e, device %.3e\n",
                    errRef, errGpu);
                expect(errGpu > 0. && errGpu <= 2. * errRef, "device Barnes-Hut is as accurate as the reference's");
                expect(sameNeighbourCounts(ga, gd), "NEIGHBOR_CNT identical with device gravity");
            }
        }

        // ---- 3b. particle removal: the Storage tells the solver (IStorageUserData), the device mirror is rebuilt ----
        {
            Storage c = base->clone(VisitorEnum::ALL_BUFFERS), r = base->clone(VisitorEnum::ALL_BUFFERS);
            GpuSolver gpu3(*scheduler, settings, eqs);
            Statistics s3, s4;
            s3.set(StatisticsId::RUN_TIME, 0._f);
            s4.set(StatisticsId::RUN_TIME, 0._f);
            c.zeroHighestDerivatives(*scheduler);
            r.zeroHighestDerivatives(*scheduler);
            gpu3.integrate(c, s3);
            refSolver.integrate(r, s4); // (both storages have been through one evaluation: S is yielded in place)
            const Size gen = gpu3.getGeneration();
            Array<Size> toRemove;
            for (Size i = 0; i < c.getParticleCnt(); i += 7) {
                toRemove.push(i);
            }
            c.remove(toRemove, Storage::IndicesFlag::INDICES_SORTED);
            r.remove(toRemove, Storage::IndicesFlag::INDICES_SORTED);
            expect(gpu3.getGeneration() != gen, "Storage::remove is seen by the solver (generation changes)");
            c.zeroHighestDerivatives(*scheduler);
            r.zeroHighestDerivatives(*scheduler);
            gpu3.integrate(c, s3);
            refSolver.integrate(r, s4);
            expect(sameNeighbourCounts(r, c), "NEIGHBOR_CNT identical after removing every 7th particle");
            expect(compareStorages(r, c, true, "integrate() after Storage::remove") <= 1.e-10, "all quantities within 1e-10 after removal");
        }

        // ---- 4. unsupported setups throw InvalidSetup instead of silently running elsewhere ----
        {
            // the constructor Factory::getSolver calls (AsymmetricSolver.h:126-129)
            GpuSolver viaFactorySignature(*scheduler, settings, eqs, makeAuto<NullBoundaryCondition>());
            bool bcThrown = false;
            try {
                GpuSolver withBc(*scheduler, settings, eqs, makeAuto<GhostParticles>(makeAuto<SphericalDomain>(Vector(0._f), 1._f), settings));
            } catch (const InvalidSetup&) {
                bcThrown = true;
            }
            expect(bcThrown, "a boundary condition without an implementation (GhostParticles) is rejected with InvalidSetup");
            // KillEscapersBoundary removes the particles outside its domain from the Storage at the start of integrate()
            {
                Float rmax = 0._f;
                ArrayView<const Vector> rr = base->getValue<Vector>(QuantityId::POSITION);
                for (Size i = 0; i < rr.size(); ++i) {
                    rmax = max(rmax, getLength(rr[i]));
                }
                AsymmetricSolver refK(*scheduler, settings, eqs, makeAuto<KillEscapersBoundary>(makeShared<SphericalDomain>(Vector(0._f), 0.93_f * rmax)));
                GpuSolver gpuK(*scheduler, settings, eqs, makeAuto<KillEscapersBoundary>(makeShared<SphericalDomain>(Vector(0._f), 0.93_f * rmax)));
                Storage ka = base->clone(VisitorEnum::ALL_BUFFERS), kb = base->clone(VisitorEnum::ALL_BUFFERS);
                const Size before = ka.getParticleCnt();
                ka.zeroHighestDerivatives(*scheduler);
                kb.zeroHighestDerivatives(*scheduler);
                refK.integrate(ka, statsA);
                gpuK.integrate(kb, statsA);
                printf("  [KillEscapersBoundary] %u -> %u particles (reference), %u (GpuSolver)\n", unsigned(before), unsigned(ka.getParticleCnt()),
                    unsigned(kb.getParticleCnt()));
                expect(ka.getParticleCnt() < before && ka.getParticleCnt() == kb.getParticleCnt(), "escapers are removed from both Storages");
                expect(sameNeighbourCounts(ka, kb), "NEIGHBOR_CNT identical after the removal");
                expect(compareStorages(ka, kb, true, "integrate() with KillEscapersBoundary") <= 1.e-10, "all quantities within 1e-10 with KillEscapersBoundary");
            }
        }
        // ---- the XSph term (SPH_USE_XSPH): two consecutive evaluations, so that the second one starts from velocities that
        // contain the correction of the first (XSph::initialize / finalize, XSph.h:69-90)
        {
            RunSettings xs = settings;
            xs.set(RunSettingsId::SPH_USE_XSPH, true).set(RunSettingsId::SPH_XSPH_EPSILON, 0.7_f);
            const EquationHolder xeqs = getStandardEquations(xs);
            AsymmetricSolver refX(*scheduler, xs, xeqs);
            GpuSolver gpuX(*scheduler, xs, xeqs);
            Storage xa = base->clone(VisitorEnum::ALL_BUFFERS), xb = base->clone(VisitorEnum::ALL_BUFFERS);
            for (Size m = 0; m < xa.getMaterialCnt(); ++m) {
                refX.create(xa, xa.getMaterial(m));
                gpuX.create(xb, xb.getMaterial(m));
            }
            for (int pass = 0; pass < 2; ++pass) {
                xa.zeroHighestDerivatives(*scheduler);
                xb.zeroHighestDerivatives(*scheduler);
                refX.integrate(xa, statsA);
                gpuX.integrate(xb, statsA);
            }
            expect(compareStorages(xa, xb, true, "integrate() twice with the XSph term") <= 1.e-10, "all quantities within 1e-10 with XSph");
            const double xe = cmpVector(xa.getValue<Vector>(QuantityId::XSPH_VELOCITIES), xb.getValue<Vector>(QuantityId::XSPH_VELOCITIES), 3);
            printf("    %-28s %.3e\n", "XSPH_VELOCITIES", xe);
            expect(xe <= 1.e-10, "XSPH_VELOCITIES within 1e-10");
        }
        // ---- the delta-SPH terms (SPH_USE_DELTASPH): three consecutive evaluations, so that the density diffusion of the later
        // ones reads the gradient stored by the one before (DeltaSph.h:37-44, 71-93)
        {
            RunSettings ds = settings;
            ds.set(RunSettingsId::SPH_USE_DELTASPH, true)
                .set(RunSettingsId::SPH_DENSITY_DIFFUSION_DELTA, 0.1_f)
                .set(RunSettingsId::SPH_VELOCITY_DIFFUSION_ALPHA, 0.05_f);
            const EquationHolder deqs = getStandardEquations(ds);
            AsymmetricSolver refD(*scheduler, ds, deqs);
            GpuSolver gpuD(*scheduler, ds, deqs);
            Storage da = base->clone(VisitorEnum::ALL_BUFFERS), db = base->clone(VisitorEnum::ALL_BUFFERS);
            for (Size m = 0; m < da.getMaterialCnt(); ++m) {
                refD.create(da, da.getMaterial(m));
                gpuD.create(db, db.getMaterial(m));
            }
            for (int pass = 0; pass < 3; ++pass) {
                da.zeroHighestDerivatives(*scheduler);
                db.zeroHighestDerivatives(*scheduler);
                refD.integrate(da, statsA);
                gpuD.integrate(db, statsA);
            }
            expect(compareStorages(da, db, true, "integrate() three times with the delta-SPH terms") <= 1.e-10,
                "all quantities within 1e-10 with the delta-SPH terms");
            const double ge = cmpVector(da.getValue<Vector>(QuantityId::DELTASPH_DENSITY_GRADIENT),
                db.getValue<Vector>(QuantityId::DELTASPH_DENSITY_GRADIENT), 3);
            printf("    %-28s %.3e\n", "DELTASPH_DENSITY_GRADIENT", ge);
            expect(ge <= 1.e-10, "DELTASPH_DENSITY_GRADIENT within 1e-10");
        }
        // ---- the symmetric formulation: SymmetricSolver<3> of the reference next to GpuSolver::useSymmetricFormulation ----
        {
            RunSettings ss = settings;
            ss.set(RunSettingsId::SPH_SOLVER_TYPE, SolverEnum::SYMMETRIC_SOLVER).set(RunSettingsId::SPH_STRAIN_RATE_CORRECTION_TENSOR, false);
            const EquationHolder seqs = getStandardEquations(ss);
            SymmetricSolver<3> refS(*scheduler, ss, seqs);
            GpuSolver gpuS(*scheduler, ss, seqs);
            gpuS.useSymmetricFormulation(true);
            Storage sa2 = base->clone(VisitorEnum::ALL_BUFFERS), sb2 = base->clone(VisitorEnum::ALL_BUFFERS);
            sa2.zeroHighestDerivatives(*scheduler);
            sb2.zeroHighestDerivatives(*scheduler);
            refS.integrate(sa2, statsA);
            gpuS.integrate(sb2, statsA);
            expect(sameNeighbourCounts(sa2, sb2), "NEIGHBOR_CNT identical (NeighborCountTerm of the SymmetricSolver vs device)");
            expect(compareStorages(sa2, sb2, true, "SymmetricSolver<3> vs the device's symmetric formulation") <= 1.e-10,
                "all quantities within 1e-10 (symmetric formulation)");
        }
        // ---- FrozenParticles boundary condition handed to the solvers' Factory constructor ----
        {
            auto makeBc = [&]() {
                Float rmax = 0._f;
                ArrayView<const Vector> rr = base->getValue<Vector>(QuantityId::POSITION);
                for (Size i = 0; i < rr.size(); ++i) {
                    rmax = max(rmax, getLength(rr[i]));
                }
                AutoPtr<FrozenParticles> bc = makeAuto<FrozenParticles>(makeShared<SphericalDomain>(Vector(0._f), 0.9_f * rmax), 0.5_f);
                bc->freeze(1);
                return bc;
            };
            AsymmetricSolver refF(*scheduler, settings, eqs, makeBc());
            GpuSolver gpuF(*scheduler, settings, eqs, makeBc());
            Storage fa = base->clone(VisitorEnum::ALL_BUFFERS), fb = base->clone(VisitorEnum::ALL_BUFFERS);
            fa.zeroHighestDerivatives(*scheduler);
            fb.zeroHighestDerivatives(*scheduler);
            refF.integrate(fa, statsA);
            gpuF.integrate(fb, statsA);
            Size frozenCnt = 0;
            ArrayView<const Vector> dvF = fa.getD2t<Vector>(QuantityId::POSITION);
            for (Size i = 0; i < dvF.size(); ++i) {
                frozenCnt += (dvF[i] == Vector(0._f)) ? 1 : 0;
            }
            printf("  [FrozenParticles] %u of %u particles frozen\n", unsigned(frozenCnt), unsigned(dvF.size()));
            expect(frozenCnt > 0 && frozenCnt < dvF.size(), "the boundary condition freezes some particles");
            expect(compareStorages(fa, fb, true, "integrate() with FrozenParticles") <= 1.e-10, "all quantities within 1e-10 with FrozenParticles");
        }
        // ---- Post::findComponents: the reference's flood against the device's label propagation, every flag combination ----
        {
            GpuSolver gpuC(*scheduler, settings, getStandardEquations(settings));
            Storage cs = base->clone(VisitorEnum::ALL_BUFFERS);
            for (Size m = 0; m < cs.getMaterialCnt(); ++m) {
                gpuC.create(cs, cs.getMaterial(m));
            }
            using Post::ComponentFlag;
            const Flags<ComponentFlag> combos[] = { ComponentFlag::OVERLAP, ComponentFlag::SEPARATE_BY_FLAG, ComponentFlag::SORT_BY_MASS,
                ComponentFlag::SEPARATE_BY_FLAG | ComponentFlag::SORT_BY_MASS, ComponentFlag::ESCAPE_VELOCITY,
                ComponentFlag::ESCAPE_VELOCITY | ComponentFlag::SORT_BY_MASS };
            for (const Float radius : { 0.45_f, 0.6_f, 1._f }) {
                for (const Flags<ComponentFlag> f : combos) {
                    Array<Size> ia, ib;
                    const Size ca = Post::findComponents(cs, radius, f, ia);
                    const Size cb = gpuC.findComponents(cs, radius, f, ib);
                    bool same = ca == cb && ia.size() == ib.size();
                    for (Size i = 0; same && i < ia.size(); ++i) {
                        same = ia[i] == ib[i];
                    }
                    printf("  [findComponents] radius %.2f flags %u: %u components (reference), %u (GpuSolver)\n", double(radius), unsigned(f.value()),
                        unsigned(ca), unsigned(cb));
                    expect(same, "component indices identical to Post::findComponents");
                }
            }
        }
        // ---- the artificial stress (SPH_AV_USE_STRESS): two consecutive evaluations against the reference ----
        {
            RunSettings as = settings;
            as.set(RunSettingsId::SPH_AV_USE_STRESS, true).set(RunSettingsId::SPH_AV_STRESS_FACTOR, 0.2_f);
            const EquationHolder aeqs = getStandardEquations(as);
            AsymmetricSolver refA(*scheduler, as, aeqs);
            GpuSolver gpuA(*scheduler, as, aeqs);
            Storage aa = base->clone(VisitorEnum::ALL_BUFFERS), ab = base->clone(VisitorEnum::ALL_BUFFERS);
            for (Size m = 0; m < aa.getMaterialCnt(); ++m) {
                refA.create(aa, aa.getMaterial(m));
                gpuA.create(ab, ab.getMaterial(m));
            }
            for (int pass = 0; pass < 2; ++pass) {
                aa.zeroHighestDerivatives(*scheduler);
                ab.zeroHighestDerivatives(*scheduler);
                refA.integrate(aa, statsA);
                gpuA.integrate(ab, statsA);
            }
            expect(compareStorages(aa, ab, true, "integrate() twice with the artificial stress") <= 1.e-10,
                "all quantities within 1e-10 with the artificial stress");
            const double se = cmpSymmetric(aa.getValue<SymmetricTensor>(QuantityId::AV_STRESS), ab.getValue<SymmetricTensor>(QuantityId::AV_STRESS));
            printf("    %-28s %.3e\n", "AV_STRESS", se);
            expect(se <= 1.e-10, "AV_STRESS within 1e-10");
        }
        bool thrown = false;
        try {
            RunSettings s2 = settings;
            s2.set(RunSettingsId::SPH_AV_USE_STRESS, true).set(RunSettingsId::SPH_USE_XSPH, true);
            GpuSolver bad(*scheduler, s2, getStandardEquations(s2));
        } catch (const InvalidSetup&) {
            thrown = true;
        }
        expect(thrown, "the artificial stress together with XSph is rejected with InvalidSetup");
        thrown = false;
        try {
            RunSettings s3 = settings;
            s3.set(RunSettingsId::SPH_DISCRETIZATION, DiscretizationEnum::BENZ_ASPHAUG);
            GpuSolver bad(*scheduler, s3, getStandardEquations(s3));
            Storage c = base->clone(VisitorEnum::ALL_BUFFERS);
            bad.integrate(c, statsA);
        } catch (const InvalidSetup&) {
            thrown = true;
        }
        expect(thrown, "BENZ_ASPHAUG discretisation is rejected with InvalidSetup");
    } catch (const std::exception& e) {
        printf("FAIL  exception: %s\n", e.what());
        failures++;
    }
    printf(failures == 0 ? "DROPIN PASS\n" : "DROPIN FAILED (%d)\n", failures);
    return failures == 0 ? 0 : 1;
}
