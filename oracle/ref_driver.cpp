// TEST INFRASTRUCTURE ONLY (oracle) -- never linked or executed by the product path.
//
// Driver around the UNMODIFIED reference library (pavelsevecek/OpenSPH core/, compiled by oracle/Makefile into
// oracle/_ref/libopensph_core*.a). It only calls the reference's public API:
//   InitialConditions::addMonolithicBody (core/sph/initial/Initial.cpp:100-125), Tests::getSolidStorage
//   (core/tests/Setup.cpp:57-88), getStandardEquations (core/sph/solvers/StandardSets.cpp:14-95),
//   AsymmetricSolver / SymmetricSolver<3> ::create/integrate (core/sph/solvers/AsymmetricSolver.cpp:71-102),
//   Factory::getTimeStepping / ITimeStepping::step (core/timestepping/TimeStepping.cpp:34-75),
//   Factory::getFinder + IBasicFinder::findAll (core/objects/finders/NeighborFinder.h:39-133).
//
// Commands
//   sph_ref snapshot --config C --n N [--jitter SEED] [--threads T] [--finder kd|grid] [--solver asym|sym]
//                    [--neighbours] [--steps K] [--integrator pc|euler] [--no-lut] [--lut LUT.snap]
//                    [--frozen-flag F] [--frozen-domain RADIUS [--frozen-radius r]]
//                    [--corrected 0|1] [--const-h] [--enforcing] [--continuity-undamaged] [--sum-all] [--balsara] [--xsph [EPS]] [--deltasph [--deltasph-delta D] [--deltasph-alpha A]] [--stress-av [--stress-av-exponent N] [--stress-av-factor X]] [--criteria MASK]
//                    --in IN.snap --out OUT.snap
//       builds the Storage of config C, writes it (state BEFORE integrate) to IN.snap, then either runs one
//       solver.integrate() on zeroed highest derivatives (K == 0) or K time steps, and writes OUT.snap.
//   sph_ref gravity  --config C --n N [--jitter SEED] [--gravity bh|brute] [--theta X] [--order 0|2|3] [--leaf L] [--point]
//                    [--no-lut] [--threads T] --out OUT.snap
//       builds the Storage of config C and evaluates the reference's self-gravity on it: Factory::getGravity
//       (core/system/Factory.cpp:361-413) -> IGravity::build + evalSelfGravity (core/gravity/BarnesHut.cpp:50-99,
//       core/gravity/BruteForceGravity.h:38-45) on a zeroed acceleration buffer. OUT.snap holds pos, mass, grav_acc, the
//       softening kernel's gradient table (grav_lut), grav_params {theta, order, G, kernel radius, leaf size, seconds}
//       and, for Barnes-Hut, the moments of the root node (BarnesHut::getMoments) with the accelerations
//       evaluateGravity (core/gravity/Moments.h:315-340) derives from them at a few distant probe points.
//   sph_ref components --config C --n N [--jitter SEED] --radius X [--separate-by-flag] [--sort-by-mass] --out OUT.snap
//       Post::findComponents (core/post/Analysis.cpp:36-75,115-217) on the particles of config C: OUT.snap holds pos, mass, flag,
//       comp_idx and comp_params {radius, flags, component count, seconds}.
//   sph_ref bench    --config C --n N --steps K --warmup W [--threads T] [--finder kd|grid] [--integrate-only] [--fixed-dt X]
//       prints one JSON line with seconds per step of the reference CPU path.
//
// Snapshot format "SPHSNAP1": u32 count, then per array {char name[32]; u32 dtype(0=f64,1=u32); u32 ncomp;
// u64 rows; payload}. All particle arrays are in the reference's own particle order.
#include "Sph.h"
#include "gravity/BarnesHut.h"
#include "gravity/Moments.h"
#include "post/Analysis.h"
#include "sph/kernel/GravityKernel.h"
#include "tests/Setup.h"
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <string>
#include <vector>

using namespace Sph;

namespace {

struct SnapWriter {
    struct Item {
        std::string name;
        uint32_t dtype, ncomp;
        uint64_t rows;
        std::vector<char> data;
    };
    std::vector<Item> items;

    void addF64(const std::string& name, const std::vector<double>& v, uint32_t ncomp) {
        Item it{ name, 0, ncomp, v.size() / ncomp, {} };
        it.data.resize(v.size() * 8);
        memcpy(it.data.data(), v.data(), it.data.size());
        items.push_back(std::move(it));
    }
    void addU32(const std::string& name, const std::vector<uint32_t>& v, uint32_t ncomp) {
        Item it{ name, 1, ncomp, v.size() / ncomp, {} };
        it.data.resize(v.size() * 4);
        memcpy(it.data.data(), v.data(), it.data.size());
        items.push_back(std::move(it));
    }
    void write(const std::string& path) {
        std::ofstream f(path, std::ios::binary);
        f.write("SPHSNAP1", 8);
        uint32_t cnt = items.size();
        f.write((char*)&cnt, 4);
        for (auto& it : items) {
            char name[32] = { 0 };
            strncpy(name, it.name.c_str(), 31);
            f.write(name, 32);
            f.write((char*)&it.dtype, 4);
            f.write((char*)&it.ncomp, 4);
            f.write((char*)&it.rows, 8);
            f.write(it.data.data(), it.data.size());
        }
    }
};

std::vector<double> vec4(ArrayView<const Vector> a) {
    std::vector<double> out(a.size() * 4);
    for (Size i = 0; i < a.size(); ++i) {
        out[4 * i + 0] = a[i][X];
        out[4 * i + 1] = a[i][Y];
        out[4 * i + 2] = a[i][Z];
        out[4 * i + 3] = a[i][H];
    }
    return out;
}
std::vector<double> scal(ArrayView<const Float> a) {
    return std::vector<double>(a.begin(), a.end());
}
std::vector<uint32_t> uscal(ArrayView<const Size> a) {
    return std::vector<uint32_t>(a.begin(), a.end());
}
std::vector<double> tt5(ArrayView<const TracelessTensor> a) {
    std::vector<double> out(a.size() * 5);
    for (Size i = 0; i < a.size(); ++i) {
        out[5 * i + 0] = a[i](0, 0);
        out[5 * i + 1] = a[i](1, 1);
        out[5 * i + 2] = a[i](0, 1);
        out[5 * i + 3] = a[i](0, 2);
        out[5 * i + 4] = a[i](1, 2);
    }
    return out;
}
std::vector<double> st6(ArrayView<const SymmetricTensor> a) {
    std::vector<double> out(a.size() * 6);
    for (Size i = 0; i < a.size(); ++i) {
        out[6 * i + 0] = a[i](0, 0);
        out[6 * i + 1] = a[i](1, 1);
        out[6 * i + 2] = a[i](2, 2);
        out[6 * i + 3] = a[i](0, 1);
        out[6 * i + 4] = a[i](0, 2);
        out[6 * i + 5] = a[i](1, 2);
    }
    return out;
}

struct Args {
    std::map<std::string, std::string> kv;
    bool has(const std::string& k) const {
        return kv.count(k) > 0;
    }
    std::string str(const std::string& k, const std::string& def = "") const {
        auto it = kv.find(k);
        return it == kv.end() ? def : it->second;
    }
    long num(const std::string& k, long def) const {
        auto it = kv.find(k);
        return it == kv.end() ? def : atol(it->second.c_str());
    }
};

/// Run settings of the named config.
RunSettings makeSettings(const std::string& config, const Args& args) {
    RunSettings settings; // library defaults (core/system/Settings.cpp:486-715)
    if (config == "preset" || config == "preset_const_h" || config == "fluid" || config == "gas" || config == "collision_preset") {
        // GUI "collision" preset, SphJob::getDefaultSettings (core/run/jobs/SimulationJobs.cpp:195-232),
        // SELF_GRAVITY removed (out of scope), adaptive h per BASELINE.json configs 3-4.
        settings.set(RunSettingsId::TIMESTEPPING_INTEGRATOR, TimesteppingEnum::PREDICTOR_CORRECTOR)
            .set(RunSettingsId::TIMESTEPPING_INITIAL_TIMESTEP, 0.01_f)
            .set(RunSettingsId::TIMESTEPPING_MAX_TIMESTEP, 10._f)
            .set(RunSettingsId::TIMESTEPPING_COURANT_NUMBER, 0.2_f)
            .set(RunSettingsId::SPH_SOLVER_TYPE, SolverEnum::ASYMMETRIC_SOLVER)
            .set(RunSettingsId::SPH_SOLVER_FORCES, ForceEnum::PRESSURE | ForceEnum::SOLID_STRESS)
            .set(RunSettingsId::SPH_DISCRETIZATION, DiscretizationEnum::STANDARD)
            .set(RunSettingsId::SPH_FINDER, FinderEnum::KD_TREE)
            .set(RunSettingsId::SPH_AV_TYPE, ArtificialViscosityEnum::STANDARD)
            .set(RunSettingsId::SPH_AV_ALPHA, 1.5_f)
            .set(RunSettingsId::SPH_AV_BETA, 3._f)
            .set(RunSettingsId::SPH_KERNEL, KernelEnum::CUBIC_SPLINE)
            .set(RunSettingsId::FINDER_LEAF_SIZE, 20)
            .set(RunSettingsId::RUN_THREAD_GRANULARITY, 1000)
            .set(RunSettingsId::SPH_ADAPTIVE_SMOOTHING_LENGTH, SmoothingLengthEnum::CONTINUITY_EQUATION)
            .set(RunSettingsId::SPH_ASYMMETRIC_COMPUTE_RADII_HASH_MAP, false)
            .set(RunSettingsId::SPH_STRAIN_RATE_CORRECTION_TENSOR, true)
            // Presets::makeAsteroidCollision (core/run/jobs/Presets.cpp:96-99)
            .set(RunSettingsId::TIMESTEPPING_CRITERION,
                TimeStepCriterionEnum::COURANT | TimeStepCriterionEnum::DIVERGENCE);
        if (config == "preset_const_h") {
            settings.set(RunSettingsId::SPH_ADAPTIVE_SMOOTHING_LENGTH, EMPTY_FLAGS);
        }
        if (config == "fluid" || config == "gas") {
            settings.set(RunSettingsId::SPH_SOLVER_FORCES, ForceEnum::PRESSURE)
                .set(RunSettingsId::SPH_STRAIN_RATE_CORRECTION_TENSOR, false);
        }
    } else if (config == "collision") {
        // examples/04_simple_collision/SimpleCollision.cpp:12-41
        settings.set(RunSettingsId::TIMESTEPPING_CRITERION, TimeStepCriterionEnum::COURANT)
            .set(RunSettingsId::TIMESTEPPING_MAX_TIMESTEP, 0.1_f);
    }
    // "hello", "solid_test": library defaults (examples/01_hello_asteroid/HelloAsteroid.cpp:11-36)
    if (args.str("solver") == "asym") {
        settings.set(RunSettingsId::SPH_SOLVER_TYPE, SolverEnum::ASYMMETRIC_SOLVER);
    } else if (args.str("solver") == "sym") {
        settings.set(RunSettingsId::SPH_SOLVER_TYPE, SolverEnum::SYMMETRIC_SOLVER);
    }
    if (args.str("finder") == "grid") {
        settings.set(RunSettingsId::SPH_FINDER, FinderEnum::UNIFORM_GRID);
    } else if (args.str("finder") == "kd") {
        settings.set(RunSettingsId::SPH_FINDER, FinderEnum::KD_TREE);
    }
    if (args.has("corrected")) {
        settings.set(RunSettingsId::SPH_STRAIN_RATE_CORRECTION_TENSOR, args.num("corrected", 1) != 0);
    }
    if (args.has("const-h")) {
        settings.set(RunSettingsId::SPH_ADAPTIVE_SMOOTHING_LENGTH, EMPTY_FLAGS);
    }
    if (args.has("enforcing")) { // AdaptiveSmoothingLength::enforce (EquationTerm.cpp:396-418)
        settings.set(RunSettingsId::SPH_ADAPTIVE_SMOOTHING_LENGTH,
            SmoothingLengthEnum::CONTINUITY_EQUATION | SmoothingLengthEnum::SOUND_SPEED_ENFORCING);
    }
    if (args.has("continuity-undamaged")) { // ContinuityEnum::SUM_ONLY_UNDAMAGED (EquationTerm.cpp:302-313)
        settings.set(RunSettingsId::SPH_CONTINUITY_MODE, ContinuityEnum::SUM_ONLY_UNDAMAGED);
    }
    if (args.has("balsara")) { // BalsaraSwitch<StandardAV> (core/sph/equations/av/Balsara.h)
        settings.set(RunSettingsId::SPH_AV_USE_BALSARA, true);
    }
    if (args.has("xsph")) { // the XSph term (core/sph/equations/XSph.h), first term of getStandardEquations
        settings.set(RunSettingsId::SPH_USE_XSPH, true);
        if (args.str("xsph") != "1") {
            settings.set(RunSettingsId::SPH_XSPH_EPSILON, Float(atof(args.str("xsph").c_str())));
        }
    }
    if (args.has("deltasph")) { // DeltaSph::DensityDiffusion + VelocityDiffusion (core/sph/equations/DeltaSph.h; StandardSets.cpp:64-67)
        settings.set(RunSettingsId::SPH_USE_DELTASPH, true);
        if (args.has("deltasph-delta")) {
            settings.set(RunSettingsId::SPH_DENSITY_DIFFUSION_DELTA, Float(atof(args.str("deltasph-delta").c_str())));
        }
        if (args.has("deltasph-alpha")) {
            settings.set(RunSettingsId::SPH_VELOCITY_DIFFUSION_ALPHA, Float(atof(args.str("deltasph-alpha").c_str())));
        }
    }
    if (args.has("stress-av")) { // the artificial stress StressAV (core/sph/equations/av/Stress.cpp; StandardSets.cpp:72-74)
        settings.set(RunSettingsId::SPH_AV_USE_STRESS, true);
        if (args.has("stress-av-exponent")) {
            settings.set(RunSettingsId::SPH_AV_STRESS_EXPONENT, Float(atof(args.str("stress-av-exponent").c_str())));
        }
        if (args.has("stress-av-factor")) {
            settings.set(RunSettingsId::SPH_AV_STRESS_FACTOR, Float(atof(args.str("stress-av-factor").c_str())));
        }
    }
    if (args.has("sum-all")) { // SPH_SUM_ONLY_UNDAMAGED = false: no undamaged filter
        settings.set(RunSettingsId::SPH_SUM_ONLY_UNDAMAGED, false);
    }
    if (args.str("integrator") == "euler") {
        settings.set(RunSettingsId::TIMESTEPPING_INTEGRATOR, TimesteppingEnum::EULER_EXPLICIT);
    } else if (args.str("integrator") == "pc") {
        settings.set(RunSettingsId::TIMESTEPPING_INTEGRATOR, TimesteppingEnum::PREDICTOR_CORRECTOR);
    }
    if (args.has("criteria")) {
        settings.set(RunSettingsId::TIMESTEPPING_CRITERION, Flags<TimeStepCriterionEnum>::fromValue(args.num("criteria", 0)));
    }
    settings.set(RunSettingsId::RUN_THREAD_CNT, int(args.num("threads", 0)));
    return settings;
}

/// Builds the particle storage of the named config using the reference's own initial conditions.
void makeStorage(const std::string& config, const Size n, const RunSettings& settings, Storage& storage) {
    if (config == "solid_test") {
        // core/sph/solvers/benchmark/Solvers.cpp:20 -- the author's own benchmark input
        storage = Tests::getSolidStorage(n, BodySettings::getDefaults(), 1.e3_f);
        return;
    }
    InitialConditions ic(settings);
    BodySettings body;
    if (config == "hello") {
        body.set(BodySettingsId::PARTICLE_COUNT, int(n));
        ic.addMonolithicBody(storage, SphericalDomain(Vector(0._f), 1.e3_f), body);
    } else if (config == "collision" || config == "collision_preset") {
        body.set(BodySettingsId::PARTICLE_COUNT, int(n));
        ic.addMonolithicBody(storage, SphericalDomain(Vector(0._f), 1.e5_f), body);
        body.set(BodySettingsId::PARTICLE_COUNT, int(max<Size>(n / 100, 10)));
        BodyView impactor =
            ic.addMonolithicBody(storage, SphericalDomain(Vector(1.4e5_f, 0._f, 0._f), 2.e4_f), body);
        impactor.addVelocity(Vector(-5.e3_f, 0._f, 0._f));
    } else if (config == "preset" || config == "preset_const_h") {
        body.set(BodySettingsId::PARTICLE_COUNT, int(n));
        if (n >= 5000000) {
            body.set(BodySettingsId::WEIBULL_SAMPLE_DISTRIBUTIONS, true);
        }
        ic.addMonolithicBody(storage, SphericalDomain(Vector(0._f), 5.e4_f), body);
    } else if (config == "fluid") {
        body.set(BodySettingsId::PARTICLE_COUNT, int(n))
            .set(BodySettingsId::RHEOLOGY_YIELDING, YieldingEnum::NONE)
            .set(BodySettingsId::RHEOLOGY_DAMAGE, FractureEnum::NONE);
        ic.addMonolithicBody(storage, SphericalDomain(Vector(0._f), 5.e4_f), body);
    } else if (config == "gas") {
        // ideal gas ball (IdealGasEos::evaluate, core/physics/Eos.cpp:42-45), cf. the reference's gas-ball solver tests
        // (core/sph/solvers/test/Solvers.cpp:41-76: Tests::getGassStorage)
        body.set(BodySettingsId::PARTICLE_COUNT, int(n))
            .set(BodySettingsId::EOS, EosEnum::IDEAL_GAS)
            .set(BodySettingsId::ADIABATIC_INDEX, 1.4_f)
            .set(BodySettingsId::DENSITY, 1._f)
            .set(BodySettingsId::DENSITY_RANGE, Interval(1.e-3_f, INFTY))
            .set(BodySettingsId::ENERGY, 1._f)
            .set(BodySettingsId::ENERGY_RANGE, Interval(1.e-3_f, INFTY))
            .set(BodySettingsId::RHEOLOGY_YIELDING, YieldingEnum::NONE)
            .set(BodySettingsId::RHEOLOGY_DAMAGE, FractureEnum::NONE);
        ic.addMonolithicBody(storage, SphericalDomain(Vector(0._f), 1._f), body);
    } else {
        throw std::runtime_error("unknown config " + config);
    }
}

/// Smooth analytic velocity field so that AV / stress / continuity terms are non-trivial.
void seedVelocity(Storage& storage, const Float scale) {
    ArrayView<Vector> r, v, dv;
    tie(r, v, dv) = storage.getAll<Vector>(QuantityId::POSITION);
    Float rmax = 0._f;
    for (Size i = 0; i < r.size(); ++i) {
        rmax = max(rmax, getLength(r[i]));
    }
    for (Size i = 0; i < r.size(); ++i) {
        const Float x = r[i][X] / rmax, y = r[i][Y] / rmax, z = r[i][Z] / rmax;
        Vector w(-0.8_f * x + 0.3_f * y * z, 0.5_f * std::sin(3._f * x) - 0.6_f * y, 0.4_f * z * x - 0.7_f * z + 0.2_f * y);
        const Float hOld = v[i][H];
        v[i] += scale * w;
        v[i][H] = hOld;
    }
}

/// Random perturbation of the state to exercise every branch (EoS phases, yielding, damage, undamaged filter).
void jitter(Storage& storage, const unsigned seed) {
    std::mt19937_64 gen(seed);
    std::uniform_real_distribution<double> uni(0., 1.);
    ArrayView<Vector> r, v, dv;
    tie(r, v, dv) = storage.getAll<Vector>(QuantityId::POSITION);
    ArrayView<Float> rho = storage.getValue<Float>(QuantityId::DENSITY);
    ArrayView<Float> u = storage.getValue<Float>(QuantityId::ENERGY);
    Float cs0 = 3000._f;
    for (Size i = 0; i < r.size(); ++i) {
        const Float h = r[i][H];
        r[i][X] += 0.25_f * h * (2 * uni(gen) - 1);
        r[i][Y] += 0.25_f * h * (2 * uni(gen) - 1);
        r[i][Z] += 0.25_f * h * (2 * uni(gen) - 1);
        r[i][H] = h * (0.85_f + 0.3_f * uni(gen));
        const Float vh = v[i][H];
        v[i] += Vector(0.02_f * cs0 * (2 * uni(gen) - 1), 0.02_f * cs0 * (2 * uni(gen) - 1), 0.02_f * cs0 * (2 * uni(gen) - 1));
        v[i][H] = vh;
        rho[i] *= 0.9_f + 0.2_f * uni(gen);
        const Float c = uni(gen);
        u[i] = c < 0.5 ? 1.e5_f * uni(gen) : (c < 0.8 ? 5.e6_f * uni(gen) : 3.e7_f * uni(gen));
    }
    if (storage.has(QuantityId::DEVIATORIC_STRESS)) {
        ArrayView<TracelessTensor> s = storage.getValue<TracelessTensor>(QuantityId::DEVIATORIC_STRESS);
        for (Size i = 0; i < s.size(); ++i) {
            const Float a = (uni(gen) < 0.3 ? 4.e9_f : 4.e8_f);
            s[i] = TracelessTensor(a * (2 * uni(gen) - 1), a * (2 * uni(gen) - 1), a * (2 * uni(gen) - 1), a * (2 * uni(gen) - 1), a * (2 * uni(gen) - 1));
        }
    }
    if (storage.has(QuantityId::DAMAGE)) {
        ArrayView<Float> d = storage.getValue<Float>(QuantityId::DAMAGE);
        for (Size i = 0; i < d.size(); ++i) {
            const Float c = uni(gen);
            d[i] = c < 0.4 ? 0._f : (c < 0.47 ? 1._f : uni(gen));
        }
    }
}

void dumpLut(const RunSettings& settings, SnapWriter& w);

void dumpState(const Storage& storage, const RunSettings& settings, SnapWriter& w, const bool withLut = true) {
    ArrayView<const Vector> r, v, dv;
    tie(r, v, dv) = storage.getAll<Vector>(QuantityId::POSITION);
    w.addF64("pos", vec4(r), 4);
    w.addF64("vel", vec4(v), 4);
    w.addF64("acc", vec4(dv), 4);
    w.addF64("mass", scal(storage.getValue<Float>(QuantityId::MASS)), 1);
    auto first = [&](const char* name, const char* dname, QuantityId id) {
        if (storage.has(id)) {
            w.addF64(name, scal(storage.getValue<Float>(id)), 1);
            if (dname && storage.getQuantity(id).getOrderEnum() != OrderEnum::ZERO) {
                w.addF64(dname, scal(storage.getDt<Float>(id)), 1);
            }
        }
    };
    first("rho", "drho", QuantityId::DENSITY);
    first("u", "du", QuantityId::ENERGY);
    first("p", nullptr, QuantityId::PRESSURE);
    first("cs", nullptr, QuantityId::SOUND_SPEED);
    first("reduce", nullptr, QuantityId::STRESS_REDUCING);
    first("damage", "ddamage", QuantityId::DAMAGE);
    first("eps_min", nullptr, QuantityId::EPS_MIN);
    first("m_zero", nullptr, QuantityId::M_ZERO);
    first("growth", nullptr, QuantityId::EXPLICIT_GROWTH);
    first("divv", nullptr, QuantityId::VELOCITY_DIVERGENCE);
    if (storage.has(QuantityId::DEVIATORIC_STRESS)) {
        w.addF64("S", tt5(storage.getValue<TracelessTensor>(QuantityId::DEVIATORIC_STRESS)), 5);
        w.addF64("dS", tt5(storage.getDt<TracelessTensor>(QuantityId::DEVIATORIC_STRESS)), 5);
    }
    if (storage.has(QuantityId::VELOCITY_GRADIENT)) {
        w.addF64("gradv", st6(storage.getValue<SymmetricTensor>(QuantityId::VELOCITY_GRADIENT)), 6);
    }
    if (storage.has(QuantityId::XSPH_VELOCITIES)) {
        w.addF64("xsph", vec4(storage.getValue<Vector>(QuantityId::XSPH_VELOCITIES)), 4);
    }
    if (storage.has(QuantityId::AV_STRESS)) {
        w.addF64("av_stress", st6(storage.getValue<SymmetricTensor>(QuantityId::AV_STRESS)), 6);
    }
    first("wp", nullptr, QuantityId::INTERPARTICLE_SPACING_KERNEL);
    if (storage.has(QuantityId::DELTASPH_DENSITY_GRADIENT)) {
        w.addF64("drho_grad", vec4(storage.getValue<Vector>(QuantityId::DELTASPH_DENSITY_GRADIENT)), 4);
    }
    if (storage.has(QuantityId::STRAIN_RATE_CORRECTION_TENSOR)) {
        w.addF64("corr", st6(storage.getValue<SymmetricTensor>(QuantityId::STRAIN_RATE_CORRECTION_TENSOR)), 6);
    }
    auto uq = [&](const char* name, QuantityId id) {
        if (storage.has(id)) {
            w.addU32(name, uscal(storage.getValue<Size>(id)), 1);
        }
    };
    uq("flag", QuantityId::FLAG);
    uq("n_flaws", QuantityId::N_FLAWS);
    uq("ncnt", QuantityId::NEIGHBOR_CNT);

    // materials: index ranges + every constant the device path needs
    std::vector<uint32_t> matRange;
    std::vector<double> matParams;
    for (Size i = 0; i < storage.getMaterialCnt(); ++i) {
        MaterialView mat = storage.getMaterial(i);
        IndexSequence seq = mat.sequence();
        matRange.push_back(*seq.begin());
        matRange.push_back(*seq.end());
        const BodySettings& b = mat->getParams();
        auto f = [&](BodySettingsId id) { return double(b.get<Float>(id)); };
        const Interval rhoRange = mat->range(QuantityId::DENSITY);
        const Interval uRange = mat->range(QuantityId::ENERGY);
        const Interval dRange = mat->range(QuantityId::DAMAGE);
        const double row[32] = {
            double(int(b.get<EosEnum>(BodySettingsId::EOS))),
            f(BodySettingsId::TILLOTSON_SUBLIMATION), f(BodySettingsId::TILLOTSON_ENERGY_IV),
            f(BodySettingsId::TILLOTSON_ENERGY_CV), f(BodySettingsId::TILLOTSON_SMALL_A),
            f(BodySettingsId::TILLOTSON_SMALL_B), f(BodySettingsId::DENSITY), f(BodySettingsId::BULK_MODULUS),
            f(BodySettingsId::TILLOTSON_NONLINEAR_B), f(BodySettingsId::TILLOTSON_ALPHA),
            f(BodySettingsId::TILLOTSON_BETA), f(BodySettingsId::ADIABATIC_INDEX),
            double(int(b.get<YieldingEnum>(BodySettingsId::RHEOLOGY_YIELDING))),
            double(int(b.get<FractureEnum>(BodySettingsId::RHEOLOGY_DAMAGE))),
            f(BodySettingsId::SHEAR_MODULUS), f(BodySettingsId::ELASTICITY_LIMIT), f(BodySettingsId::MELT_ENERGY),
            f(BodySettingsId::YOUNG_MODULUS),
            rhoRange.lower(), rhoRange.upper(), uRange.lower(), uRange.upper(), dRange.lower(), dRange.upper(),
            mat->minimal(QuantityId::DENSITY), mat->minimal(QuantityId::ENERGY), mat->minimal(QuantityId::DAMAGE),
            mat->minimal(QuantityId::DEVIATORIC_STRESS), 0, 0, 0, 0 };
        matParams.insert(matParams.end(), row, row + 32);
    }
    w.addU32("mat_range", matRange, 2);
    w.addF64("mat_params", matParams, 32);

    LutKernel<3> kernel = Factory::getKernel<3>(settings);
    const Flags<SmoothingLengthEnum> hflags =
        settings.getFlags<SmoothingLengthEnum>(RunSettingsId::SPH_ADAPTIVE_SMOOTHING_LENGTH);
    const Flags<ForceEnum> forces = settings.getFlags<ForceEnum>(RunSettingsId::SPH_SOLVER_FORCES);
    const Interval hRange = settings.get<Interval>(RunSettingsId::SPH_SMOOTHING_LENGTH_RANGE);
    const Interval nRange = settings.get<Interval>(RunSettingsId::SPH_NEIGHBOR_RANGE);
    std::vector<double> run = {
        kernel.radius(),
        settings.get<Float>(RunSettingsId::SPH_AV_ALPHA),
        settings.get<Float>(RunSettingsId::SPH_AV_BETA),
        double(forces.has(ForceEnum::PRESSURE)),
        double(forces.has(ForceEnum::SOLID_STRESS)),
        double(settings.get<bool>(RunSettingsId::SPH_STRAIN_RATE_CORRECTION_TENSOR)),
        double(settings.get<bool>(RunSettingsId::SPH_SUM_ONLY_UNDAMAGED)),
        double(hflags.has(SmoothingLengthEnum::CONTINUITY_EQUATION)),
        double(hflags.has(SmoothingLengthEnum::SOUND_SPEED_ENFORCING)),
        double(int(settings.get<ContinuityEnum>(RunSettingsId::SPH_CONTINUITY_MODE))),
        double(int(settings.get<DiscretizationEnum>(RunSettingsId::SPH_DISCRETIZATION))),
        hRange.lower(), hRange.upper(),
        settings.get<Float>(RunSettingsId::SPH_NEIGHBOR_ENFORCING), nRange.lower(), nRange.upper(),
        settings.get<Float>(RunSettingsId::TIMESTEPPING_COURANT_NUMBER),
        settings.get<Float>(RunSettingsId::TIMESTEPPING_DERIVATIVE_FACTOR),
        settings.get<Float>(RunSettingsId::TIMESTEPPING_DIVERGENCE_FACTOR),
        double(settings.getFlags<TimeStepCriterionEnum>(RunSettingsId::TIMESTEPPING_CRITERION).value()),
        settings.get<Float>(RunSettingsId::TIMESTEPPING_MAX_TIMESTEP),
        settings.get<Float>(RunSettingsId::TIMESTEPPING_INITIAL_TIMESTEP),
        settings.get<Float>(RunSettingsId::TIMESTEPPING_MAX_INCREASE),
        double(int(settings.get<SolverEnum>(RunSettingsId::SPH_SOLVER_TYPE))),
        double(int(settings.get<TimesteppingEnum>(RunSettingsId::TIMESTEPPING_INTEGRATOR))),
        double(settings.get<bool>(RunSettingsId::SPH_AV_USE_BALSARA)),
        double(settings.get<bool>(RunSettingsId::SPH_USE_XSPH)),
        settings.get<Float>(RunSettingsId::SPH_XSPH_EPSILON),
        double(settings.get<bool>(RunSettingsId::SPH_USE_DELTASPH)),
        settings.get<Float>(RunSettingsId::SPH_DENSITY_DIFFUSION_DELTA),
        settings.get<Float>(RunSettingsId::SPH_VELOCITY_DIFFUSION_ALPHA),
        double(settings.get<bool>(RunSettingsId::SPH_AV_USE_STRESS)),
        settings.get<Float>(RunSettingsId::SPH_AV_STRESS_EXPONENT),
        settings.get<Float>(RunSettingsId::SPH_AV_STRESS_FACTOR),
    };
    w.addF64("run_params", run, 1);

    if (withLut) {
        dumpLut(settings, w);
    }
}

void dumpLut(const RunSettings& settings, SnapWriter& w) {
    // the LUT exactly as the reference builds it (core/sph/kernel/Kernel.h:85-101)
    LutKernel<3> kernel = Factory::getKernel<3>(settings);
    const Size entries = 40000;
    const Float qSqrToIdx = Float(entries) / sqr(kernel.radius());
    std::vector<double> lutGrad(entries + 1), lutVal(entries + 1);
    for (Size i = 0; i <= entries; ++i) {
        const Float qSqr = Float(i) / qSqrToIdx;
        // exact node values: at the nodes the linear interpolation returns the table entries themselves
        lutGrad[i] = (i < entries) ? kernel.gradImpl(qSqr) : 0.;
        lutVal[i] = (i < entries) ? kernel.valueImpl(qSqr) : 0.;
    }
    w.addF64("lut_grad", lutGrad, 1);
    w.addF64("lut_val", lutVal, 1);
}

/// Neighbour lists exactly as AsymmetricSolver::loop selects them (core/sph/solvers/AsymmetricSolver.cpp:174-199),
/// obtained from the reference finder; sorted by index for set comparison.
void dumpNeighbours(const Storage& storage, const RunSettings& settings, IScheduler& scheduler, SnapWriter& w) {
    ArrayView<const Vector> r = storage.getValue<Vector>(QuantityId::POSITION);
    AutoPtr<ISymmetricFinder> finder = Factory::getFinder(settings);
    finder->build(scheduler, r);
    LutKernel<3> kernel = Factory::getKernel<3>(settings);
    Float maxH = 0._f;
    for (Size i = 0; i < r.size(); ++i) {
        maxH = max(maxH, r[i][H]);
    }
    const Float maxRadius = maxH * kernel.radius();
    std::vector<uint32_t> offsets(2 * (r.size() + 1)); // u64 stored as two u32 (little endian)
    std::vector<uint32_t> idxs;
    Array<NeighborRecord> neighs;
    uint64_t total = 0;
    for (Size i = 0; i < r.size(); ++i) {
        offsets[2 * i] = uint32_t(total & 0xffffffffu);
        offsets[2 * i + 1] = uint32_t(total >> 32);
        const Float radius = 0.5_f * (r[i][H] * kernel.radius() + maxRadius);
        finder->findAll(i, radius, neighs);
        std::vector<uint32_t> mine;
        for (auto& n : neighs) {
            const Size j = n.index;
            const Float hbar = 0.5_f * (r[i][H] + r[j][H]);
            if (i == j || n.distanceSqr >= sqr(kernel.radius() * hbar)) {
                continue;
            }
            mine.push_back(j);
        }
        std::sort(mine.begin(), mine.end());
        idxs.insert(idxs.end(), mine.begin(), mine.end());
        total += mine.size();
    }
    offsets[2 * r.size()] = uint32_t(total & 0xffffffffu);
    offsets[2 * r.size() + 1] = uint32_t(total >> 32);
    w.addU32("nbr_offsets", offsets, 2);
    w.addU32("nbr_idx", idxs, 1);
}

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        std::cerr << "usage: sph_ref snapshot|bench --config C --n N ..." << std::endl;
        return 2;
    }
    const std::string cmd = argv[1];
    Args args;
    for (int i = 2; i < argc; ++i) {
        std::string k = argv[i];
        if (k.rfind("--", 0) != 0) {
            continue;
        }
        k = k.substr(2);
        if (i + 1 < argc && std::string(argv[i + 1]).rfind("--", 0) != 0) {
            args.kv[k] = argv[++i];
        } else {
            args.kv[k] = "1";
        }
    }
    try {
        const std::string config = args.str("config", "hello");
        const Size n = Size(args.num("n", 10000));
        RunSettings settings = makeSettings(config, args);
        SharedPtr<IScheduler> scheduler = Factory::getScheduler(settings);
        SharedPtr<Storage> storage = makeShared<Storage>();
        makeStorage(config, n, settings, *storage);

        // --frozen-flag F / --frozen-domain RADIUS [--frozen-radius r]: FrozenParticles boundary condition
        // (core/sph/boundary/Boundary.cpp:203-258) handed to the solver like IRun::setUp would
        AutoPtr<IBoundaryCondition> bc;
        if (args.has("frozen-flag") || args.has("frozen-domain")) {
            AutoPtr<FrozenParticles> frozen;
            if (args.has("frozen-domain")) {
                frozen = makeAuto<FrozenParticles>(makeShared<SphericalDomain>(Vector(0._f), Float(atof(args.str("frozen-domain").c_str()))),
                    Float(atof(args.str("frozen-radius", "2").c_str())));
            } else {
                frozen = makeAuto<FrozenParticles>();
            }
            if (args.has("frozen-flag")) {
                frozen->freeze(Size(args.num("frozen-flag", 0)));
            }
            bc = std::move(frozen);
        }
        AutoPtr<ISolver> solver = bc ? Factory::getSolver(*scheduler, settings, std::move(bc)) : Factory::getSolver(*scheduler, settings);
        for (Size i = 0; i < storage->getMaterialCnt(); ++i) {
            solver->create(*storage, storage->getMaterial(i));
        }
        if (args.num("velocity", 1) != 0) {
            seedVelocity(*storage, Float(args.num("vscale", 50)));
        }
        if (args.has("jitter")) {
            jitter(*storage, unsigned(args.num("jitter", 1)));
        }
        const Size N = storage->getParticleCnt();

        if (cmd == "snapshot") {
            const bool withLut = !args.has("no-lut");
            if (args.has("lut")) {
                SnapWriter w;
                dumpLut(settings, w);
                w.write(args.str("lut"));
            }
            if (args.has("in")) {
                SnapWriter w;
                dumpState(*storage, settings, w, withLut);
                w.write(args.str("in"));
            }
            Statistics stats;
            stats.set(StatisticsId::RUN_TIME, 0._f);
            const long steps = args.num("steps", 0);
            std::vector<double> dts;
            if (steps == 0) {
                storage->zeroHighestDerivatives(*scheduler);
                solver->integrate(*storage, stats);
            } else {
                AutoPtr<ITimeStepping> stepping = Factory::getTimeStepping(settings, storage);
                for (long s = 0; s < steps; ++s) {
                    dts.push_back(stepping->getTimeStep());
                    stepping->step(*scheduler, *solver, stats);
                }
                dts.push_back(stepping->getTimeStep());
            }
            SnapWriter w;
            dumpState(*storage, settings, w, withLut);
            if (!dts.empty()) {
                w.addF64("dt_history", dts, 1);
            }
            if (args.has("neighbours")) {
                dumpNeighbours(*storage, settings, *scheduler, w);
            }
            w.write(args.str("out", "out.snap"));
            std::cout << "{\"particles\": " << N << ", \"materials\": " << storage->getMaterialCnt() << "}"
                      << std::endl;
        } else if (cmd == "gravity") {
            const bool brute = args.str("gravity", "bh") == "brute";
            const double theta = atof(args.str("theta", "0.5").c_str());
            const int order = int(args.num("order", 3));
            settings.set(RunSettingsId::GRAVITY_SOLVER, brute ? GravityEnum::BRUTE_FORCE : GravityEnum::BARNES_HUT)
                .set(RunSettingsId::GRAVITY_OPENING_ANGLE, Float(theta))
                .set(RunSettingsId::GRAVITY_MULTIPOLE_ORDER, order)
                .set(RunSettingsId::GRAVITY_KERNEL,
                    args.has("point") ? GravityKernelEnum::POINT_PARTICLES : GravityKernelEnum::SPH_KERNEL);
            if (args.has("leaf")) {
                settings.set(RunSettingsId::FINDER_LEAF_SIZE, int(args.num("leaf", 25)));
            }
            AutoPtr<IGravity> gravity = Factory::getGravity(settings);
            Array<Vector> dv(N);
            dv.fill(Vector(0._f));
            Statistics stats;
            const double t0 = now();
            gravity->build(*scheduler, *storage);
            gravity->evalSelfGravity(*scheduler, dv, stats);
            const double seconds = now() - t0;
            SnapWriter w;
            ArrayView<const Vector> r = storage->getValue<Vector>(QuantityId::POSITION);
            ArrayView<const Float> m = storage->getValue<Float>(QuantityId::MASS);
            w.addF64("pos", vec4(r), 4);
            w.addF64("mass", scal(m), 1);
            w.addF64("grav_acc", vec4(dv), 4);
            const Float G = settings.get<Float>(RunSettingsId::GRAVITY_CONSTANT);
            double radius = 0.;
            if (!args.has("point")) {
                radius = GravityKernel<CubicSpline<3>>().radius();
            }
            if (!args.has("point") && !args.has("no-lut")) {
                // the table GravityLutKernel holds: LutKernel<3> built from the exact gravity kernel (GravityKernel.h:40-50,
                // Kernel.h:85-101); sampled at its own nodes, where the interpolation returns the entries themselves
                GravityKernel<CubicSpline<3>> exact;
                LutKernel<3> lk(exact);
                const Size entries = 40000;
                radius = lk.radius();
                const Float qSqrToIdx = Float(entries) / sqr(lk.radius());
                std::vector<double> lut(entries + 1);
                for (Size i = 0; i <= entries; ++i) {
                    const Float qSqr = Float(i) / qSqrToIdx;
                    lut[i] = (i < entries) ? lk.gradImpl(qSqr) : exact.gradImpl(qSqr);
                }
                w.addF64("grav_lut", lut, 1);
            }
            w.addF64("grav_params",
                { theta, double(order), double(G), radius, double(settings.get<int>(RunSettingsId::FINDER_LEAF_SIZE)), seconds,
                    brute ? 1. : 0. },
                1);
            if (!brute) {
                const BarnesHut* bh = dynamic_cast<const BarnesHut*>(&*gravity);
                if (bh) {
                    const MultipoleExpansion<3> ms = bh->getMoments(); // divided by G
                    const TracelessMultipole<2>& q2 = ms.order<2>();
                    const TracelessMultipole<3>& q3 = ms.order<3>();
                    std::vector<double> mom = { ms.order<0>().value(), q2.value<0, 0>(), q2.value<1, 1>(), q2.value<0, 1>(),
                        q2.value<0, 2>(), q2.value<1, 2>(), q3.value<0, 0, 0>(), q3.value<0, 0, 1>(), q3.value<0, 0, 2>(),
                        q3.value<0, 1, 1>(), q3.value<0, 1, 2>(), q3.value<1, 1, 1>(), q3.value<1, 1, 2>() };
                    w.addF64("root_moments", mom, 1);
                    // centre of mass as BarnesHut::buildLeaf / buildInner define it (mass-weighted mean)
                    Vector com(0._f);
                    Float mtot = 0._f;
                    for (Size i = 0; i < N; ++i) {
                        com += m[i] * r[i];
                        mtot += m[i];
                    }
                    com /= mtot;
                    Float rmax = 0._f;
                    for (Size i = 0; i < N; ++i) {
                        rmax = max(rmax, getLength(r[i] - com));
                    }
                    std::vector<double> probes, field;
                    const Vector dirs[4] = { Vector(1._f, 0.2_f, -0.3_f), Vector(-0.5_f, 1._f, 0.4_f), Vector(0.1_f, -0.7_f, 1._f),
                        Vector(-1._f, -1._f, -1._f) };
                    const MultipoleExpansion<3> msG = ms.multiply(G);
                    for (int o : { 0, 2, 3 }) {
                        for (int k = 0; k < 4; ++k) {
                            const Vector dr = dirs[k] / getLength(dirs[k]) * (3._f + k) * rmax;
                            const Vector a = evaluateGravity(dr, msG, MultipoleOrder(o));
                            probes.insert(probes.end(), { com[X] + dr[X], com[Y] + dr[Y], com[Z] + dr[Z], double(o) });
                            field.insert(field.end(), { a[X], a[Y], a[Z] });
                        }
                    }
                    w.addF64("probe_pos", probes, 4);
                    w.addF64("probe_acc", field, 3);
                }
                std::vector<double> st = { double(stats.get<int>(StatisticsId::GRAVITY_NODES_APPROX)),
                    double(stats.get<int>(StatisticsId::GRAVITY_NODES_EXACT)), double(stats.get<int>(StatisticsId::GRAVITY_NODE_COUNT)) };
                w.addF64("grav_stats", st, 1);
            }
            w.write(args.str("out", "gravity.snap"));
            printf("{\"particles\": %u, \"gravity\": \"%s\", \"theta\": %g, \"order\": %d, \"seconds\": %.4f, \"threads\": %d}\n", N,
                brute ? "brute" : "bh", theta, order, seconds, int(scheduler->getThreadCnt()));
        } else if (cmd == "components") {
            // Post::findComponents (core/post/Analysis.cpp:115-217) on the positions of the config: --radius X (in units of h),
            // --separate-by-flag, --sort-by-mass
            const Float radius = Float(atof(args.str("radius", "1").c_str()));
            Flags<Post::ComponentFlag> flags = Post::ComponentFlag::OVERLAP;
            if (args.has("separate-by-flag")) {
                flags.set(Post::ComponentFlag::SEPARATE_BY_FLAG);
            }
            if (args.has("sort-by-mass")) {
                flags.set(Post::ComponentFlag::SORT_BY_MASS);
            }
            Array<Size> indices;
            const double t0 = now();
            const Size count = Post::findComponents(*storage, radius, flags, indices);
            const double seconds = now() - t0;
            SnapWriter w;
            w.addF64("pos", vec4(storage->getValue<Vector>(QuantityId::POSITION)), 4);
            w.addF64("mass", scal(storage->getValue<Float>(QuantityId::MASS)), 1);
            if (storage->has(QuantityId::FLAG)) {
                w.addU32("flag", uscal(storage->getValue<Size>(QuantityId::FLAG)), 1);
            }
            w.addU32("comp_idx", uscal(indices), 1);
            w.addF64("comp_params", { double(radius), double(flags.value()), double(count), seconds }, 1);
            w.write(args.str("out", "out.snap"));
            std::cout << "{\"particles\": " << N << ", \"components\": " << count << ", \"seconds\": " << seconds << "}" << std::endl;
        } else if (cmd == "bench") {
            const long steps = args.num("steps", 3), warmup = args.num("warmup", 1);
            Statistics stats;
            stats.set(StatisticsId::RUN_TIME, 0._f);
            double total = 0.;
            if (args.has("integrate-only")) {
                for (long s = 0; s < warmup + steps; ++s) {
                    storage->zeroHighestDerivatives(*scheduler);
                    const double t0 = now();
                    solver->integrate(*storage, stats);
                    if (s >= warmup) {
                        total += now() - t0;
                    }
                }
            } else {
                // default: the time step the config's criteria choose (initial / maximal step of the config), so the
                // particles move; --fixed-dt X pins the step (X = 1e-6 keeps the lattice intact: identical work per step)
                if (args.has("fixed-dt")) {
                    const Float fixedDt = Float(atof(args.str("fixed-dt", "1e-6").c_str()));
                    settings.set(RunSettingsId::TIMESTEPPING_INITIAL_TIMESTEP, fixedDt)
                        .set(RunSettingsId::TIMESTEPPING_MAX_TIMESTEP, fixedDt);
                }
                AutoPtr<ITimeStepping> stepping = Factory::getTimeStepping(settings, storage);
                for (long s = 0; s < warmup + steps; ++s) {
                    const double t0 = now();
                    stepping->step(*scheduler, *solver, stats);
                    if (s >= warmup) {
                        total += now() - t0;
                    }
                }
            }
            const MinMaxMean nc = stats.get<MinMaxMean>(StatisticsId::NEIGHBOR_COUNT);
            printf("{\"particles\": %u, \"steps\": %ld, \"seconds_per_step\": %.6f, \"particle_updates_per_s\": %.6e, "
                   "\"threads\": %d, \"neigh_mean\": %.2f, \"neigh_max\": %.0f, \"mode\": \"%s\"}\n",
                N, steps, total / steps, double(N) * steps / total, int(scheduler->getThreadCnt()), nc.mean(), nc.max(),
                args.has("integrate-only") ? "integrate" : "full_step");
        } else {
            std::cerr << "unknown command " << cmd << std::endl;
            return 2;
        }
    } catch (const std::exception& e) {
        std::cerr << "sph_ref error: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
