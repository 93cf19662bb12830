/*
 * sphgpu.h -- C ABI of the B200-native SPH evaluation engine (libsphgpu.so).
 *
 * This is the drop-in boundary for OpenSPH's per-step hot path. The reference has no C ABI (it is one C++
 * address space); each entry point below names the reference interface it replaces (paths relative to the
 * reference root). The reference-side binding (a C++ `GpuSolver : ISolver` that forwards to these calls) is
 * in opensph_b200/host/GpuSolver.{h,cpp} and described in INTEGRATION.md.
 *
 * Conventions
 *  - every function returns 0 on success, a negative SPHGPU_E_* code otherwise; sphgpu_last_error() returns a
 *    human-readable message for the calling thread (reference convention: C++ exceptions such as InvalidSetup,
 *    core/sph/solvers/AsymmetricSolver.cpp:228-234 -- the C++ wrapper turns non-zero into those exceptions);
 *  - one context per device, calls on one context are serialised by the caller (ISolver::integrate is not
 *    re-entrant, core/timestepping/ISolver.h:28-31);
 *  - plain pointers and sizes only; no CUDA or torch types. Pointers are HOST pointers unless the name says
 *    `_device`;
 *  - there is NO CPU fallback: if no CUDA device is usable, sphgpu_create fails with SPHGPU_E_NO_DEVICE.
 */
#ifndef SPHGPU_H
#define SPHGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPHGPU_ABI_VERSION 1

#if defined(__GNUC__)
#define SPHGPU_API __attribute__((visibility("default")))
#else
#define SPHGPU_API
#endif

/* ---- status codes ------------------------------------------------------------------------------------- */
enum {
    SPHGPU_OK = 0,
    SPHGPU_E_INVALID = -1,     /* bad argument / unsupported setup  (reference: InvalidSetup)              */
    SPHGPU_E_NO_DEVICE = -2,   /* no usable CUDA device -- there is no CPU path                            */
    SPHGPU_E_CUDA = -3,        /* a CUDA runtime call or kernel failed                                      */
    SPHGPU_E_OOM = -4,         /* device allocation failed                                                  */
    SPHGPU_E_STATE = -5        /* call order violated (e.g. integrate before the state was uploaded)       */
};

/* ---- enumerations mirroring the reference's settings (core/system/Settings.h) ------------------------ */
enum { SPHGPU_FORCE_PRESSURE = 1u << 0, SPHGPU_FORCE_SOLID_STRESS = 1u << 1 }; /* ForceEnum subset         */
enum {
    SPHGPU_FLAG_CORRECTION_TENSOR = 1u << 0,    /* RunSettingsId::SPH_STRAIN_RATE_CORRECTION_TENSOR          */
    SPHGPU_FLAG_SUM_ONLY_UNDAMAGED = 1u << 1,   /* RunSettingsId::SPH_SUM_ONLY_UNDAMAGED                     */
    SPHGPU_FLAG_ADAPTIVE_H = 1u << 2,           /* SmoothingLengthEnum::CONTINUITY_EQUATION                  */
    SPHGPU_FLAG_SOUND_SPEED_ENFORCING = 1u << 3, /* SmoothingLengthEnum::SOUND_SPEED_ENFORCING               */
    SPHGPU_FLAG_BALSARA = 1u << 4,              /* RunSettingsId::SPH_AV_USE_BALSARA: BalsaraSwitch<StandardAV>,
                                                   core/sph/equations/av/Balsara.h:36-153. Not on decomposed runs
                                                   (sphgpu_halo_configure refuses: the ghosts do not carry the div v /
                                                   rot v of the previous evaluation)                            */
    SPHGPU_FLAG_XSPH = 1u << 5,                 /* RunSettingsId::SPH_USE_XSPH: the XSph term, core/sph/equations/XSph.h:20-97;
                                                   its epsilon is set with sphgpu_set_xsph_epsilon. Not together with
                                                   SPHGPU_FLAG_BALSARA, not on decomposed runs                 */
    SPHGPU_FLAG_DELTASPH = 1u << 6,             /* RunSettingsId::SPH_USE_DELTASPH: DeltaSph::DensityDiffusion and
                                                   DeltaSph::VelocityDiffusion, core/sph/equations/DeltaSph.h:12-189
                                                   (StandardSets.cpp:64-67); coefficients: sphgpu_set_deltasph. Not together
                                                   with SPHGPU_FLAG_BALSARA or SPHGPU_FLAG_XSPH, not on decomposed runs; with
                                                   SPHGPU_FLAG_CORRECTION_TENSOR only for solids                  */
    SPHGPU_FLAG_STRESS_AV = 1u << 7             /* RunSettingsId::SPH_AV_USE_STRESS: the artificial stress StressAV,
                                                   core/sph/equations/av/Stress.h:10-36, Stress.cpp:8-123 (added by
                                                   getStandardEquations, StandardSets.cpp:72-74); exponent and factor:
                                                   sphgpu_set_stress_av. Needs ForceEnum::SOLID_STRESS and the kernel value
                                                   table (lut_value). Not together with SPHGPU_FLAG_BALSARA, _XSPH or
                                                   _DELTASPH, not on decomposed runs                              */
};
enum { SPHGPU_DISCR_STANDARD = 0, SPHGPU_DISCR_BENZ_ASPHAUG = 1 };        /* DiscretizationEnum            */
enum { SPHGPU_CONTINUITY_STANDARD = 0, SPHGPU_CONTINUITY_SUM_ONLY_UNDAMAGED = 1 }; /* ContinuityEnum       */
enum { SPHGPU_EOS_NONE = 0, SPHGPU_EOS_IDEAL_GAS = 1, SPHGPU_EOS_TILLOTSON = 4 };  /* EosEnum values       */
enum { SPHGPU_YIELD_NONE = 0, SPHGPU_YIELD_ELASTIC = 1, SPHGPU_YIELD_VON_MISES = 2, SPHGPU_YIELD_DUST = 4 };
enum { SPHGPU_FRACTURE_NONE = 0, SPHGPU_FRACTURE_SCALAR_GRADY_KIPP = 1 };
/* TimeStepCriterionEnum bit values (core/system/Settings.h:577-592) */
enum {
    SPHGPU_CRIT_COURANT = 1u << 1,
    SPHGPU_CRIT_DERIVATIVES = 1u << 2,
    SPHGPU_CRIT_ACCELERATION = 1u << 3,
    SPHGPU_CRIT_DIVERGENCE = 1u << 4
};
/* CriterionId (core/timestepping/TimeStepCriterion.h:18-26) */
enum {
    SPHGPU_CRITID_INITIAL_VALUE = 0,
    SPHGPU_CRITID_MAXIMAL_VALUE = 1,
    SPHGPU_CRITID_DERIVATIVE = 2,
    SPHGPU_CRITID_CFL_CONDITION = 3,
    SPHGPU_CRITID_ACCELERATION = 4,
    SPHGPU_CRITID_DIVERGENCE = 5,
    SPHGPU_CRITID_MAX_CHANGE = 6
};

/* Quantities of the Storage schema the path touches (QuantityId in
 * core/quantities/QuantityIds.h). `order` selects value (0), first (1) or second (2) time derivative. */
enum {
    SPHGPU_Q_POSITION = 0,            /* Vector {x,y,z,h}; order 1 = {v, dh/dt}; order 2 = {a, 0}           */
    SPHGPU_Q_MASS = 1,                /* f64                                                                 */
    SPHGPU_Q_DENSITY = 2,             /* f64, orders 0..1                                                    */
    SPHGPU_Q_ENERGY = 3,              /* f64, orders 0..1                                                    */
    SPHGPU_Q_PRESSURE = 4,            /* f64                                                                 */
    SPHGPU_Q_SOUND_SPEED = 5,         /* f64                                                                 */
    SPHGPU_Q_DEVIATORIC_STRESS = 6,   /* TracelessTensor {xx,yy,xy,xz,yz}, orders 0..1                       */
    SPHGPU_Q_DAMAGE = 7,              /* f64, orders 0..1                                                    */
    SPHGPU_Q_STRESS_REDUCING = 8,     /* f64                                                                 */
    SPHGPU_Q_VELOCITY_DIVERGENCE = 9, /* f64                                                                 */
    SPHGPU_Q_VELOCITY_GRADIENT = 10,  /* SymmetricTensor {xx,yy,zz,xy,xz,yz}                                 */
    SPHGPU_Q_CORRECTION_TENSOR = 11,  /* SymmetricTensor (STRAIN_RATE_CORRECTION_TENSOR)                     */
    SPHGPU_Q_EPS_MIN = 12,            /* f64                                                                 */
    SPHGPU_Q_M_ZERO = 13,             /* f64                                                                 */
    SPHGPU_Q_EXPLICIT_GROWTH = 14,    /* f64                                                                 */
    SPHGPU_Q_N_FLAWS = 15,            /* u32                                                                 */
    SPHGPU_Q_FLAG = 16,               /* u32  body index                                                     */
    SPHGPU_Q_NEIGHBOR_CNT = 17,       /* u32                                                                 */
    SPHGPU_Q_MATERIAL_ID = 18,        /* u32  index into the materials passed to sphgpu_create; initialised from
                                         their [begin,end) ranges, must be uploaded for ghost particles      */
    SPHGPU_Q_VELOCITY_ROTATION = 19,  /* Vector {x,y,z,0}: nabla x v, input and output of the Balsara switch  */
    SPHGPU_Q_XSPH_VELOCITIES = 20,    /* Vector {x,y,z,0}: the velocity correction the XSph term left in the velocities */
    SPHGPU_Q_DELTASPH_DENSITY_GRADIENT = 21, /* Vector {x,y,z,0}: the renormalised density gradient, output of one
                                         evaluation and input of the next (DeltaSph.h:16-44,71-74)          */
    SPHGPU_Q_AV_STRESS = 22,          /* SymmetricTensor: the artificial stress -(S - p I)+ every evaluation computes
                                         (StressAV::initialize, Stress.cpp:91-109); output                   */
    SPHGPU_Q_INTERPARTICLE_SPACING_KERNEL = 23, /* f64: W(h, h) of StressAV::create (Stress.cpp:113-121), set once from
                                         the initial smoothing lengths; must be uploaded with SPHGPU_FLAG_STRESS_AV */
    SPHGPU_Q_COUNT = 24
};

/* Host memory layouts understood by upload/download. */
enum {
    SPHGPU_LAYOUT_PACKED = 0, /* components tightly packed: Vector 4 f64, TracelessTensor 5 f64 {xx,yy,xy,xz,yz},
                                 SymmetricTensor 6 f64 {xx,yy,zz,xy,xz,yz}, scalars 1 f64 / 1 u32              */
    SPHGPU_LAYOUT_OPENSPH = 1 /* the reference's in-memory AoS: Vector 32 B {x,y,z,h} (geometry/Vector.h:378-395),
                                 TracelessTensor 64 B {xx,yy,xy,xz|yz,pad} (TracelessTensor.h:36-45),
                                 SymmetricTensor 64 B {xx,yy,zz,pad|xy,xz,yz,pad} (SymmetricTensor.h:18-21)    */
};

/* ---- configuration ------------------------------------------------------------------------------------ */

/* Run-level configuration: what AsymmetricSolver's constructor reads from RunSettings
 * (core/sph/solvers/AsymmetricSolver.cpp:58-69,124-136; core/sph/solvers/StandardSets.cpp:14-95). */
typedef struct sphgpu_config {
    uint32_t abi_version;        /* SPHGPU_ABI_VERSION                                                        */
    uint32_t forces;             /* SPHGPU_FORCE_*                                                            */
    uint32_t flags;              /* SPHGPU_FLAG_*                                                             */
    uint32_t discretization;     /* SPHGPU_DISCR_*  (only STANDARD is implemented)                            */
    uint32_t continuity_mode;    /* SPHGPU_CONTINUITY_*                                                       */
    uint32_t lut_entries;        /* LutKernel::NEntries = 40000 (tables hold lut_entries+1 values)            */
    const double* lut_grad;      /* (dW/dq)/q sampled in q^2, core/sph/kernel/Kernel.h:85-101                 */
    const double* lut_value;     /* W sampled in q^2 (only W(0) is needed: density floor); may be NULL        */
    double kernel_radius;        /* LutKernel::radius(), 2 for the cubic spline                               */
    double av_alpha, av_beta;    /* StandardAV, core/sph/equations/av/Standard.h:44-46                        */
    double h_min, h_max;         /* SPH_SMOOTHING_LENGTH_RANGE, EquationTerm.cpp:349,356-364                  */
    double neigh_enforcing;      /* SPH_NEIGHBOR_ENFORCING (used only with SOUND_SPEED_ENFORCING)             */
    double neigh_lower, neigh_upper; /* SPH_NEIGHBOR_RANGE                                                    */
    /* time-step criteria, core/timestepping/TimeStepCriterion.cpp:117-419 */
    uint32_t criteria;           /* SPHGPU_CRIT_* mask                                                        */
    uint32_t reserved0;
    double courant;              /* TIMESTEPPING_COURANT_NUMBER                                               */
    double derivative_factor;    /* TIMESTEPPING_DERIVATIVE_FACTOR                                            */
    double divergence_factor;    /* TIMESTEPPING_DIVERGENCE_FACTOR                                            */
    double max_change;           /* TIMESTEPPING_MAX_INCREASE (>= 1e300 means unlimited)                      */
} sphgpu_config;

/* Per-material constants (one per body; materials own contiguous particle index ranges,
 * MaterialView::sequence(), core/quantities/IMaterial.h:110-194). */
typedef struct sphgpu_material {
    uint32_t begin, end;         /* particle index range [begin,end)                                          */
    uint32_t eos;                /* SPHGPU_EOS_*                                                              */
    uint32_t yielding;           /* SPHGPU_YIELD_*                                                            */
    uint32_t fracture;           /* SPHGPU_FRACTURE_*                                                         */
    uint32_t reserved0;
    /* TillotsonEos (core/physics/Eos.cpp:184-238) */
    double til_u0, til_uiv, til_ucv, til_a, til_b, rho0, til_A, til_B, til_alpha, til_beta;
    double gamma;                /* IdealGasEos (Eos.cpp:42-45)                                               */
    double shear_modulus;        /* mu, SolidStressForce::finalize (EquationTerm.cpp:190-201)                 */
    double elasticity_limit;     /* von Mises Y0 (Rheology.cpp:47)                                            */
    double melt_energy;          /* u_melt (Rheology.cpp:50)                                                  */
    double young_modulus;        /* set by ScalarGradyKippModel::setFlaws (Damage.cpp:44-48)                  */
    /* ranges / minimals used by the integrators and the derivative criterion (IMaterial::range/minimal) */
    double rho_min, rho_max, u_min, u_max, d_min, d_max;
    double rho_small, u_small, d_small, s_small;
} sphgpu_material;

typedef struct sphgpu_stats {
    uint32_t neigh_min, neigh_max; /* StatisticsId::NEIGHBOR_COUNT (AsymmetricSolver.cpp:218-225)            */
    double neigh_mean;
    uint64_t pair_count;           /* sum of NEIGHBOR_CNT                                                    */
    double gpu_ms;                 /* device time of the call, CUDA events                                   */
    uint32_t kernel_launches;      /* kernels launched by this call                                          */
    uint32_t reserved0;            /* work units whose candidate lists overflowed the list pool (slow path)  */
} sphgpu_stats;

typedef struct sphgpu_timestep {
    double dt;                 /* new time step                                                              */
    uint32_t criterion;        /* SPHGPU_CRITID_*                                                            */
    uint32_t reserved0;
} sphgpu_timestep;

typedef struct sphgpu_ctx sphgpu_ctx;

/* ---- lifetime ------------------------------------------------------------------------------------------ */

/* Replaces AsymmetricSolver::AsymmetricSolver + Factory::getKernel/getFinder (AsymmetricSolver.cpp:58-69,124-136).
 * `capacity` >= n_particles reserves room for ghost (halo) particles appended after the owned ones. */
SPHGPU_API int sphgpu_create(const sphgpu_config* cfg, const sphgpu_material* materials, uint32_t n_materials,
    uint32_t n_particles, uint32_t capacity, int device, sphgpu_ctx** out);
SPHGPU_API int sphgpu_destroy(sphgpu_ctx* ctx);
SPHGPU_API const char* sphgpu_last_error(void);
SPHGPU_API uint32_t sphgpu_abi_version(void);

/* ---- state transfer (replaces Storage::getValue/getDt/getD2t array access, core/quantities/Storage.h:291-608) */

/* Copies `count` particles starting at particle index `first` of quantity `q`/`order` between host memory in
 * `layout` and the device-resident SoA mirror. */
SPHGPU_API int sphgpu_upload(sphgpu_ctx* ctx, int q, int order, int layout, const void* host, uint32_t first, uint32_t count);
SPHGPU_API int sphgpu_download(sphgpu_ctx* ctx, int q, int order, int layout, void* host, uint32_t first, uint32_t count);
/* Same, PACKED layout only, with a DEVICE pointer on the context's device (used by the multi-GPU halo plumbing). */
/* Asynchronous variants for drivers that keep the Storage on the host and move it every step: the calls only queue work.
 * Uploads are ordered with the kernels on the context's stream; downloads are packed on that stream and copied out on a
 * second stream, so that the device -> host transfer of one step overlaps with the host -> device transfer of the next
 * (PCIe is full duplex). The host buffers must stay valid -- and, to really be asynchronous, be page-locked -- until
 * sphgpu_transfer_sync returns. The downloads of one step form a batch, closed by sphgpu_download_batch_end. */
SPHGPU_API int sphgpu_upload_async(sphgpu_ctx* ctx, int q, int order, int layout, const void* host, uint32_t first, uint32_t count);
SPHGPU_API int sphgpu_download_async(sphgpu_ctx* ctx, int q, int order, int layout, void* host, uint32_t first, uint32_t count);
SPHGPU_API int sphgpu_download_batch_end(sphgpu_ctx* ctx);
SPHGPU_API int sphgpu_transfer_sync(sphgpu_ctx* ctx);
/* Page-locked host memory for the asynchronous transfers (so that host code needs no CUDA headers). */
SPHGPU_API int sphgpu_host_alloc(void** out, size_t bytes);
SPHGPU_API int sphgpu_host_free(void* ptr);
SPHGPU_API int sphgpu_upload_device(sphgpu_ctx* ctx, int q, int order, const void* dev, uint32_t first, uint32_t count);
SPHGPU_API int sphgpu_download_device(sphgpu_ctx* ctx, int q, int order, void* dev, uint32_t first, uint32_t count);
/* Halo exchange helpers (multi-GPU): pack / unpack the dynamic neighbour inputs {r,h | v,dh/dt | rho | u | S[5] | D}
 * of `count` particles starting at slot `first` as records of SPHGPU_HALO_DOUBLES doubles into / from a DEVICE buffer
 * in one kernel each. The reference has no counterpart (single address space); ghosts play the role of
 * GhostParticles (core/sph/boundary/Boundary.h:73-). */
#define SPHGPU_HALO_DOUBLES 16
SPHGPU_API int sphgpu_halo_pack(sphgpu_ctx* ctx, uint32_t first, uint32_t count, void* dev_buffer);
SPHGPU_API int sphgpu_halo_unpack(sphgpu_ctx* ctx, uint32_t first, uint32_t count, const void* dev_buffer);
/* Multi-GPU step inside the library (one process per GPU, slab neighbours left/right; NCCL is resolved at run time).
 *  - sphgpu_comm_unique_id: rank 0 creates the 128-byte NCCL id, the caller distributes it (e.g. torch.distributed);
 *  - sphgpu_comm_init: every rank joins the communicator (collective);
 *  - sphgpu_halo_configure: neighbour ranks (-1 = none) and the sizes of the slot bands [0, send_left) and
 *    [n - send_right, n) that are sent, and of the ghost ranges [n, n + recv_left), [.., + recv_right) that are filled;
 *    SPHGPU_E_INVALID for contexts with SPHGPU_FLAG_BALSARA / _XSPH / _DELTASPH / _STRESS_AV, whose pair terms need per-particle
 *    results of the previous evaluation (or constants) that the bands do not carry;
 *  - sphgpu_halo_exchange: pack -> grouped ncclSend/ncclRecv -> unpack, queued on the context's stream;
 *  - sphgpu_step_pc_mgpu: predict -> halo exchange -> integrate -> correct -> criteria -> ncclAllReduce(min) of the time
 *    step, one host synchronisation per step. */
SPHGPU_API int sphgpu_comm_unique_id(void* out128);
SPHGPU_API int sphgpu_comm_init(sphgpu_ctx* ctx, const void* id128, int rank, int world);
SPHGPU_API int sphgpu_halo_configure(sphgpu_ctx* ctx, int left_rank, int right_rank, uint32_t send_left, uint32_t send_right,
    uint32_t recv_left, uint32_t recv_right);
SPHGPU_API int sphgpu_halo_exchange(sphgpu_ctx* ctx);
/* Exchange over peer memory (NVLink) instead of NCCL point-to-point, for ranks on one node (one process per GPU):
 * sphgpu_peer_export fills SPHGPU_PEER_BLOB_BYTES with the CUDA IPC handles of this rank's halo planes and mailbox; the
 * caller gathers the blobs of all ranks (rank order) and passes them to sphgpu_peer_connect. From then on
 * sphgpu_halo_exchange / sphgpu_step_pc_mgpu / sphgpu_run_pc write the send bands straight into the neighbours' ghost
 * slots with one kernel (no staging, no unpack) and combine the time-step minima through the mailboxes. Call both again
 * after every sphgpu_halo_configure. Without them the NCCL path is used. */
#define SPHGPU_PEER_BLOB_BYTES 1280
SPHGPU_API int sphgpu_peer_export(sphgpu_ctx* ctx, void* blob);
SPHGPU_API int sphgpu_peer_connect(sphgpu_ctx* ctx, const void* blobs, int world);
/* Guard of the fixed send bands: the cut planes of this rank's domain perpendicular to `axis` (has_lo / has_hi: there is
 * a neighbour behind that plane). Every exchange then measures the smallest head-room of an interior particle (one that
 * is in neither band) towards a plane, (distance / (R (h_i + h_max) / 2)) - 1. Once it would be negative the next
 * synchronising call fails with SPHGPU_E_STATE instead of silently missing cross-rank neighbours (the caller repartitions
 * and calls sphgpu_halo_configure again). sphgpu_halo_margin returns the head-room seen at the last synchronisation, so a
 * driver can repartition before that happens. The reference's analogue of stale ghosts is GhostParticles being
 * regenerated in every IBoundaryCondition::initialize (core/sph/boundary/Boundary.cpp). */
SPHGPU_API int sphgpu_halo_set_guard(sphgpu_ctx* ctx, int axis, double lo_plane, double hi_plane, int has_lo, int has_hi);
SPHGPU_API int sphgpu_halo_margin(sphgpu_ctx* ctx, double* margin);
SPHGPU_API int sphgpu_step_pc_mgpu(sphgpu_ctx* ctx, double t, double dt, double max_dt, sphgpu_stats* stats, sphgpu_timestep* out);
/* `steps` PredictorCorrector steps queued back to back: the time step chosen by the criteria (MultiCriterion::compute,
 * TimeStepCriterion.cpp:389-419, evaluated by a device kernel) stays on the device and feeds the next step, so the host
 * synchronises once for the whole batch instead of once per step (what IRun::run's loop does between output times,
 * core/run/IRun.cpp:232-250). Uses the halo exchange + ncclAllReduce of the time step when a communicator is configured.
 * `dt` is the first step, history[s] the step chosen AFTER step s (history[steps-1].dt is the next dt); stats and the
 * timings refer to the last step. Identical results to `steps` calls of sphgpu_step_pc / sphgpu_step_pc_mgpu. */
SPHGPU_API int sphgpu_run_pc(sphgpu_ctx* ctx, uint32_t steps, double dt, double max_dt, sphgpu_stats* stats, sphgpu_timestep* history);
/* Number of particles that take part as neighbours: owned + ghosts (ghosts occupy [n_particles, n_active)). */
SPHGPU_API int sphgpu_set_active(sphgpu_ctx* ctx, uint32_t n_active);

/* ---- the hot path --------------------------------------------------------------------------------------- */

/* Replaces IAsymmetricSolver::integrate (AsymmetricSolver.cpp:71-96): material->initialize (EoS + rheology),
 * equations.initialize (h clamp), finder build + findAll, derivatives.eval over all pairs, accumulated.store,
 * equations.finalize, material->finalize (damage growth). Highest derivatives are OVERWRITTEN, i.e. the call
 * behaves as the reference does after Storage::zeroHighestDerivatives (TimeStepping.cpp:238,334). */
SPHGPU_API int sphgpu_integrate(sphgpu_ctx* ctx, double t, sphgpu_stats* stats);

/* Replaces PredictorCorrector::makePredictions + swap + zeroHighestDerivatives (TimeStepping.cpp:286-300,331-334). */
SPHGPU_API int sphgpu_step_predict(sphgpu_ctx* ctx, double dt);
/* Replaces PredictorCorrector::makeCorrections (TimeStepping.cpp:302-322). */
SPHGPU_API int sphgpu_step_correct(sphgpu_ctx* ctx, double dt);
/* Replaces EulerExplicit::stepParticles after solver.integrate (TimeStepping.cpp:243-264). */
SPHGPU_API int sphgpu_step_euler(sphgpu_ctx* ctx, double dt);
/* Replaces MultiCriterion::compute (TimeStepCriterion.cpp:389-419). */
SPHGPU_API int sphgpu_compute_timestep(sphgpu_ctx* ctx, double max_dt, sphgpu_timestep* out);
/* Seeds MultiCriterion::lastStep (TimeStepCriterion.cpp:387) used when max_change limits the growth of dt. */
SPHGPU_API int sphgpu_set_last_timestep(sphgpu_ctx* ctx, double dt);
/* One whole PredictorCorrector step on the device (ITimeStepping::step, TimeStepping.cpp:34-75):
 * predict(dt) -> integrate -> correct(dt) -> criteria. No host<->device traffic except the returned scalars. */
SPHGPU_API int sphgpu_step_pc(sphgpu_ctx* ctx, double t, double dt, double max_dt, sphgpu_stats* stats, sphgpu_timestep* out);

/* ---- self-gravity --------------------------------------------------------------- */

/* Replaces the IGravity object GravitySolver<TSphSolver> owns (core/sph/solvers/GravitySolver.cpp:18-30,64-99; built by
 * Factory::getGravity, core/system/Factory.cpp:361-413): BarnesHut (core/gravity/BarnesHut.cpp:50-501) with the opening
 * angle GRAVITY_OPENING_ANGLE and the multipole order GRAVITY_MULTIPOLE_ORDER, or -- opening_angle <= 0 --
 * BruteForceGravity (core/gravity/BruteForceGravity.h:38-45). Close pairs use the softening kernel GravityLutKernel
 * (core/sph/kernel/GravityKernel.h:58-86) with symmetrised smoothing lengths: lut_grad holds the lut_entries + 1 node
 * values of its gradient table over q^2 in [0, kernel_radius^2] (Kernel.h:85-101; the last one is the Newtonian value at
 * the edge); kernel_radius = 0 selects point particles (GravityKernelEnum::POINT_PARTICLES). leaf_size is
 * FINDER_LEAF_SIZE (0 = the reference's default 25): nodes of at most that many particles are summed exactly when opened.
 * Once configured, sphgpu_integrate / sphgpu_step_pc / sphgpu_run_pc add the gravitational accelerations to the SPH
 * ones like GravitySolver::loop does. The device tree is a binary radix tree over Morton-sorted particles, not the
 * reference's k-d tree: with opening_angle <= 0 the result equals the reference's to rounding, with an opening angle it
 * agrees within the error of the multipole approximation (the reference's own BarnesHut tests compare against
 * BruteForceGravity the same way, core/gravity/test/BarnesHut.cpp). Attractors (Storage::getAttractors) are not
 * handled. Not available on decomposed runs. cfg == NULL switches gravity off. */
typedef struct sphgpu_gravity {
    double opening_angle;
    int multipole_order;   /* 0, 2 or 3 (MultipoleOrder, core/gravity/Moments.h:307-312) */
    uint32_t leaf_size;
    double constant;       /* GRAVITY_CONSTANT */
    double kernel_radius;
    const double* lut_grad;
    uint32_t lut_entries;
    uint32_t reserved;
} sphgpu_gravity;

typedef struct sphgpu_gravity_stats {
    uint64_t approximated; /* node x target-group interactions evaluated by multipoles (GRAVITY_NODES_APPROX) */
    uint64_t exact;        /* particle ranges x target groups summed exactly (GRAVITY_NODES_EXACT) */
    uint32_t nodes;        /* internal nodes of the tree (GRAVITY_NODE_COUNT) */
    uint32_t groups;       /* target groups (the leaves the walk is run for) */
    double gpu_ms;         /* device time of the last evaluation: keys + sort + tree + moments + walk */
} sphgpu_gravity_stats;

SPHGPU_API int sphgpu_gravity_configure(sphgpu_ctx* ctx, const sphgpu_gravity* cfg);
/* IGravity::build + evalSelfGravity (BarnesHut.cpp:50-99) on the particles as they are on the device. accumulate != 0
 * adds to the acceleration planes, 0 overwrites them (evalSelfGravity on a zeroed buffer). Synchronises. */
SPHGPU_API int sphgpu_gravity_eval(sphgpu_ctx* ctx, int accumulate, sphgpu_gravity_stats* stats);
/* Statistics of the last evaluation, e.g. the one inside sphgpu_integrate. Synchronises. */
SPHGPU_API int sphgpu_gravity_last_stats(sphgpu_ctx* ctx, sphgpu_gravity_stats* stats);

/* ---- initial conditions on the device --------------------------------------------------- */

/* Replaces InitialConditions::addMonolithicBody (core/sph/initial/Initial.cpp:100-125) for a SphericalDomain
 * (core/objects/geometry/Domain.cpp:10-33) filled by HexagonalPacking (core/sph/initial/Distribution.cpp:126-200,
 * BodySettings defaults: not sorted, SPHGPU_LATTICE_CENTER = BodySettingsId::CENTER_PARTICLES) with
 * InitialConditions::setQuantities / getMasses (Initial.cpp:288-333): positions, h = eta * lattice spacing, masses
 * proportional to h^3 that sum to density * volume, the body flag, zero velocities and accelerations are written into
 * the slots [first, first + count) of the context without touching host memory. Lattice points are bit-identical to the
 * reference's; the centring shift and the mass normalisation are sums over all particles (sequential in the reference, a
 * tree here) and agree to rounding. The remaining quantities (density, energy, stress, flaws ...) are uploaded or filled by
 * the caller as IMaterial::create would. sphgpu_lattice_count tells how many particles the lattice has (about 6 % more
 * than particle_count), so that the context can be created with the right capacity. */
#define SPHGPU_LATTICE_CENTER 1u
typedef struct sphgpu_lattice {
    double center[3];
    double radius;
    uint32_t particle_count; /* BodySettingsId::PARTICLE_COUNT */
    uint32_t flags;
    double eta;              /* BodySettingsId::SMOOTHING_LENGTH_ETA */
    double density;          /* BodySettingsId::DENSITY */
    uint32_t body_flag;      /* value of QuantityId::FLAG (InitialConditions::bodyIndex) */
    uint32_t reserved;
} sphgpu_lattice;
SPHGPU_API int sphgpu_lattice_count(int device, const sphgpu_lattice* cfg, uint32_t* count);
SPHGPU_API int sphgpu_lattice_generate(sphgpu_ctx* ctx, const sphgpu_lattice* cfg, uint32_t first, uint32_t* count);

/* ---- boundary condition: frozen particles ---------------------------------------------- */

/* Replaces FrozenParticles::finalize (core/sph/boundary/Boundary.cpp:221-258), which the solver calls after
 * equations.finalize and before material->finalize (AsymmetricSolver.cpp:204-224): particles whose body flag is in
 * flag_mask (bit f = QuantityId::FLAG value f, f < 64), and -- with has_domain -- particles of a SphericalDomain that are
 * closer to its surface than freeze_radius smoothing lengths, get all highest derivatives set to zero (acceleration,
 * density, energy and stress derivatives; the damage derivative is the material's and stays); particles outside the
 * domain are first projected onto its surface (SphericalDomain::project, Domain.cpp:66-85). Applied by every
 * sphgpu_integrate / sphgpu_step_pc / sphgpu_run_pc after the derivatives are complete. cfg == NULL switches it off. */
typedef struct sphgpu_frozen {
    uint64_t flag_mask;
    int has_domain;
    int reserved;
    double center[3];
    double radius;
    double freeze_radius; /* in units of the smoothing length */
} sphgpu_frozen;
SPHGPU_API int sphgpu_set_frozen(sphgpu_ctx* ctx, const sphgpu_frozen* cfg);

/* ---- inspection (tests) --------------------------------------------------------------------------------- */

/* Neighbour lists exactly as AsymmetricSolver::loop selects them (AsymmetricSolver.cpp:174-199), CSR:
 * offsets[n_particles+1], indices ascending per particle. Pass idx == NULL to obtain only offsets. */
/* Post::findComponents of the reference (core/post/Analysis.cpp:36-75,115-128) on the positions and smoothing lengths in the
 * context: particles are joined along the directed relation |r_j - r_i| < h_i * radius, with SPHGPU_COMPONENTS_SEPARATE_BY_FLAG
 * (ComponentFlag::SEPARATE_BY_FLAG) only between equal body flags. indices (host, particle_count entries) receives the
 * component of every particle, numbered like the reference numbers them (in the order of the lowest particle index of each
 * component), *component_count their number; bit-identical to the reference. ESCAPE_VELOCITY and SORT_BY_MASS post-process
 * this result on the host (GpuPost::findComponents in opensph_b200/host). sweeps (may be NULL): label-propagation sweeps used.
 * Single domain only; rebuilds the cell list (the next evaluation rebuilds its neighbour lists). */
enum { SPHGPU_COMPONENTS_SEPARATE_BY_FLAG = 1u << 0 };
SPHGPU_API int sphgpu_find_components(sphgpu_ctx* ctx, double radius, uint32_t flags, uint32_t* indices, uint32_t* component_count,
    uint32_t* sweeps);
SPHGPU_API int sphgpu_neighbour_dump(sphgpu_ctx* ctx, uint64_t* offsets, uint32_t* idx, uint64_t idx_capacity);
/* Device-time breakdown of the last integrate call in milliseconds: [0] grid build + sort, [1] prologue + pack,
 * [2] pair kernel, [3] rest. */
SPHGPU_API int sphgpu_last_timings(sphgpu_ctx* ctx, double* ms4);
/* Device times (ms) of the parts of the last pair stage: {work units + lane order, candidate lists (k_pair_lists),
 * pair sums (k_pair_sum)}; zeros for the other variants. */
SPHGPU_API int sphgpu_last_pair_timings(sphgpu_ctx* ctx, double* ms3);
/* Measures the device's FP64 FMA throughput (fused multiply-adds per second, all SMs) with a register-resident DFMA
 * loop; the pair kernel is bound by this pipe, not by HBM. */
SPHGPU_API int sphgpu_measure_fp64_peak(sphgpu_ctx* ctx, double* fma_per_second);
/* Device time of the halo exchange of the last sphgpu_step_pc_mgpu call (pack + NCCL send/recv + unpack, including the
 * time spent waiting for the neighbour ranks), milliseconds. */
SPHGPU_API int sphgpu_last_halo_ms(sphgpu_ctx* ctx, double* ms);
/* Changes the number of owned particles (<= capacity) after particles were added, removed or migrated between the
 * ranks of a decomposed run (the analogue of Storage::remove / merge, core/quantities/Storage.h:560-). Slots [0, n) must
 * be uploaded again, including MATERIAL_ID for more than one material; ghosts are dropped (n_active = n). */
SPHGPU_API int sphgpu_set_particle_count(sphgpu_ctx* ctx, uint32_t n_particles);
/* Selects the pair-kernel variant: 0 = default (candidate lists in their own kernel + tiled pair sums), 1 = direct
 * per-thread kernel, 2 = every work unit through the direct per-target loop, 3 = as 0 with a tiny list pool (exercises the
 * overflow path) -- all of them the asymmetric formulation of AsymmetricSolver::loop --, 4 = the SYMMETRIC formulation of
 * SymmetricSolver::loop (core/sph/solvers/SymmetricSolver.cpp:104-163): particles ranked by smoothing length (makeRankH,
 * core/objects/finders/Order.h:46-58), every pair evaluated once by its particle of higher rank (findLowerRank) and added to
 * both particles (evalSymmetric; NeighborCountTerm, HelperTerms.h:16-47). Same results to rounding; slower (FP64 atomics),
 * without the correction tensor (as in the reference), the Balsara switch and XSph. */
SPHGPU_API int sphgpu_set_variant(sphgpu_ctx* ctx, int variant);
/* Reuse of the cell list / work units / candidate lists over several steps (the reference rebuilds its finder in every
 * ISolver::integrate, AsymmetricSolver.cpp:81-84; results do not depend on this setting beyond summation order, because
 * the exact neighbour predicate of AsymmetricSolver.cpp:186-191 is evaluated for every listed candidate in every step).
 * The lists are built with the search radius enlarged by (1 + skin) and rebuilt -- decided on the device, no host round
 * trip -- once 2 max|r - r_build| / (R h_build) + max(h / h_build - 1) reaches 0.9 skin. skin = 0 rebuilds every call.
 * Default 0.03. sphgpu_list_stats: builds so far, calls served by the current lists, the metric seen by the last call
 * (values as of the last call that synchronised with the device). */
SPHGPU_API int sphgpu_set_list_skin(sphgpu_ctx* ctx, double skin);
/* RunSettingsId::SPH_XSPH_EPSILON of the XSph term (SPHGPU_FLAG_XSPH; XSph.h:36-38). XSph::initialize subtracts the
 * correction of the previous evaluation from the velocities before the derivatives are evaluated, XSph::finalize adds the
 * new one (XSph.h:69-90): sphgpu_integrate does both, so POSITION dt holds the corrected velocities between calls and
 * SPHGPU_Q_XSPH_VELOCITIES the correction, exactly like the Storage of the reference. Default 1 (Settings.cpp). */
SPHGPU_API int sphgpu_set_xsph_epsilon(sphgpu_ctx* ctx, double epsilon);
/* RunSettingsId::SPH_DENSITY_DIFFUSION_DELTA and SPH_VELOCITY_DIFFUSION_ALPHA of the delta-SPH terms (SPHGPU_FLAG_DELTASPH;
 * DeltaSph.h:63-65,135-137). Every evaluation stores the renormalised density gradient sum_j m_j/rho_j (rho_j - rho_i)
 * C_i grad W_ij in SPHGPU_Q_DELTASPH_DENSITY_GRADIENT; the density diffusion of the NEXT evaluation reads it, as the
 * reference does through its Storage (zero before the first evaluation). Defaults 0.01, 0.01 (Settings.cpp:541-544). */
SPHGPU_API int sphgpu_set_deltasph(sphgpu_ctx* ctx, double delta, double alpha);
/* RunSettingsId::SPH_AV_STRESS_EXPONENT and SPH_AV_STRESS_FACTOR of the artificial stress (SPHGPU_FLAG_STRESS_AV;
 * Stress.cpp:24-27): Pi_ij = factor (W_ij / W(h_i, h_i)_0)^exponent (as_i / rho_i^2 + as_j / rho_j^2), dv_i += m_j Pi_ij grad W_ij,
 * du_i += m_j Pi_ij (v_i - v_j) . grad W_ij / 2 over the undamaged neighbours of the same body. Defaults 4, 0.04
 * (Settings.cpp:579-582). */
SPHGPU_API int sphgpu_set_stress_av(sphgpu_ctx* ctx, double exponent, double factor);
SPHGPU_API int sphgpu_list_stats(sphgpu_ctx* ctx, uint32_t* rebuilds, uint32_t* age, double* metric);
/* Runs all subsequent work of the context on the caller's CUDA stream (a cudaStream_t passed as void*; NULL is the
 * legacy default stream 0, which is what torch.cuda.current_stream() is unless the caller changed it), so that the
 * caller's own work -- NCCL halo exchange, torch.cuda.Event timing -- is ordered with the engine's kernels.
 * sphgpu_use_private_stream() returns to the context's own non-blocking stream (the initial state). */
SPHGPU_API int sphgpu_set_stream(sphgpu_ctx* ctx, void* cuda_stream);
SPHGPU_API int sphgpu_use_private_stream(sphgpu_ctx* ctx);
/* Blocks until all work queued on the context's stream has finished. */
SPHGPU_API int sphgpu_synchronize(sphgpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SPHGPU_H */
