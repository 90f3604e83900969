/* psim_b200 - C ABI of the B200 phonon Monte Carlo particle loop.
 *
 * The reference (GwGibson/Psim) has no plugin / FFI layer.  The seam this library sits behind is the
 * `ModelSimulator` class as `Model` uses it (reference psim/include/psim/modelSimulator.h:12-41, called from
 * psim/src/model.cpp:160-161) together with the sensor tally accessors (psim/include/psim/sensor.h:50-55,
 * 73-74).  Each entry point below names the reference interface it replaces.  INTEGRATION.md shows the
 * binding a maintainer of the reference would add.
 *
 * Conventions: plain C, caller-owned buffers, no C++ or torch types.  Every function returns 0 on success
 * or a negative PSIM_E_* code; psim_gpu_last_error() gives the message.  Nothing throws across the ABI.
 * A handle is bound to ONE CUDA device and may be used from one thread at a time.  There is no CPU path:
 * without a usable CUDA device psim_gpu_create fails with PSIM_E_NO_DEVICE.
 */
#ifndef PSIM_B200_H
#define PSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSIM_NUM_FREQ_BINS 1000 /* reference Material::NUM_FREQ_BINS, material.h:14 */

enum {
    PSIM_OK = 0,
    PSIM_E_INVALID = -1,    /* bad argument / inconsistent model description */
    PSIM_E_NO_DEVICE = -2,  /* no CUDA device, or the device is not usable */
    PSIM_E_CUDA = -3,       /* a CUDA call failed */
    PSIM_E_OVERFLOW = -4,   /* phonon pool capacity exceeded (never silently dropped) */
    PSIM_E_STATE = -5,      /* call order violated (e.g. run before set_sources) */
    PSIM_E_IO = -6,         /* host layer: file could not be read / written */
    PSIM_E_MODEL = -7,      /* host layer: model file rejected (same conditions the reference rejects) */
    PSIM_E_RNG = -8         /* a phonon exhausted the random-number blocks of one measurement interval (never a silent reuse) */
};

/* ---- flat model description -------------------------------------------------------------------------
 * What `Model` hands to `ModelSimulator` through Cell / Sensor / Material references
 * (cell.h:77-82, sensor.h:66-74, material.h:85-110), flattened to arrays. */

typedef struct psim_material {          /* RelaxationData + DispersionData cut-offs, material.h:113-139 */
    double b_l, b_tn, b_tu, b_i, w;     /* Holland-type relaxation constants; w = TA Umklapp cut-off (rad/s) */
    double w_max_la, w_max_ta;          /* rad/s */
    double freq_width;                  /* bin width, max(w_max_la, w_max_ta) / 1000 (material.cpp:29) */
} psim_material;

typedef struct psim_sensor {            /* Sensor + SensorController, sensorController.h:41-55 */
    uint32_t material;                  /* index into materials */
    uint32_t base_table;                /* table sampled for phonons born inside its cells (base_table_) */
    uint32_t scatter_table;             /* table sampled on intrinsic scattering (scatter_table_) */
    uint32_t reserved;
    double temperature;                 /* getSteadyTemp(): temperature the relaxation rates are evaluated at */
} psim_sensor;

#define PSIM_SURF_TRANSITION 1u
#define PSIM_SURF_EMIT 2u

typedef struct psim_subsurface {        /* TransitionSurface / EmitSurface inside a CompositeSurface */
    uint32_t kind;                      /* PSIM_SURF_TRANSITION or PSIM_SURF_EMIT */
    uint32_t target;                    /* neighbour cell index, or emitter index */
    uint32_t target_edge;               /* transition: which edge (0..2) of the neighbour it lies on */
    uint32_t reserved;
    double s0, s1;                      /* extent on this cell's edge, as fractions from its first vertex */
    double t0, t1;                      /* transition: the same two end points as fractions along the neighbour's edge */
} psim_subsurface;

typedef struct psim_cell {              /* Cell, cell.h:10-82.  Edge k joins vertex k to vertex (k+1)%3. */
    double x[3], y[3];                  /* triangle vertices p1, p2, p3 (nm) */
    double specularity;                 /* of its boundary surfaces */
    uint32_t sensor;                    /* index into sensors */
    uint32_t sub_first[3];              /* per edge: first entry in subsurfaces ... */
    uint32_t sub_count[3];              /* ... and how many (0 = plain boundary); transitions before emitters */
    uint32_t reserved;
} psim_cell;

typedef struct psim_emitter {           /* EmitSurface, surface.h:73-98 */
    uint32_t cell, edge;                /* where it sits */
    uint32_t table;                     /* velocity-weighted table at its temperature (emit_table_) */
    uint32_t reserved;
    double s_p1, s_p2;                  /* its two end points as fractions along that edge */
    double start_time, duration;        /* emission / absorption window (ns) */
} psim_emitter;

typedef struct psim_table {             /* Material::Table, material.h:17: 1000 x (cumulative, LA fraction) */
    const double* cumulative;           /* [1000] */
    const double* la_fraction;          /* [1000] */
} psim_table;

typedef struct psim_model_desc {
    uint32_t num_materials, num_sensors, num_cells, num_subsurfaces, num_emitters, num_tables;
    const psim_material* materials;
    const double* velocities;           /* [num_materials][2][1000]: LA then TA group velocity, m/s */
    const psim_sensor* sensors;
    const psim_cell* cells;
    const psim_subsurface* subsurfaces;
    const psim_emitter* emitters;
    const psim_table* tables;
    uint32_t measurement_steps;         /* ModelSimulator ctor, modelSimulator.cpp:29-37 */
    uint32_t step_adjustment;           /* setStepAdjustment(), modelSimulator.h:24-26 */
    double simulation_time;             /* ns */
    uint32_t full_simulation;           /* t_eq == 0 (Material::setFullSimulation) */
    uint32_t phasor_sim;
    /* Optional (NULL = the records above hold for every step): [num_sensors][measurement_steps] per-step records of a
     * transient run that re-iterates (TransientController after reset(false), sensorController.cpp:101-113: the relaxation
     * rates are evaluated at steady_temps_[step] and scatter_tables_[step] is sampled, :80-88).  base_table is ignored. */
    const psim_sensor* step_sensors;
} psim_model_desc;

#define PSIM_SRC_CELL 0u
#define PSIM_SRC_SURFACE 1u

typedef struct psim_source {            /* one phonon builder, phononBuilder.h:38-72 */
    uint32_t kind;                      /* PSIM_SRC_CELL or PSIM_SRC_SURFACE */
    uint32_t index;                     /* cell index or emitter index */
    int32_t sign;                       /* +1: hotter than t_eq, -1: colder */
    uint32_t reserved;
    uint64_t count;                     /* phonons this source emits over the whole run */
} psim_source;

typedef struct psim_stats {
    uint64_t total_phonons;             /* over all shards */
    uint64_t shard_phonons;             /* created by this handle */
    uint64_t drift_steps;               /* (phonon, measurement interval) advances performed by this handle */
    uint64_t events;                    /* free-flight segments: impacts + scatters + interval ends */
    uint64_t peak_alive;                /* largest pool population seen at a launch boundary */
    double kernel_ms;                   /* CUDA-event time of the drift launches of the last run */
    uint32_t launches;                  /* drift-kernel launches of the last run */
    uint32_t steps_per_launch;
    uint32_t warps;                     /* resident warps = pool segments */
    uint32_t tally_in_shared;           /* last launch: 0 tallies straight to global memory, 1 staged in shared memory as
                                           32-bit halves, 2 staged as 64-bit sums, 3 global memory in difference form,
                                           4 staged as three 32-bit parts (pools too large for 1) */
    uint64_t image_bytes;               /* host->device bytes of the model image (psim_gpu_create) */
    uint64_t plan_bytes;                /* host->device bytes of the sources + birth plan (psim_gpu_set_sources) */
    uint64_t tally_bytes;               /* device->host bytes of psim_gpu_get_tallies */
    uint32_t kernel;                    /* drift kernel in use: 2 work queues, 0 lane-bound slots, 1 lock step */
    uint32_t flight_cells;              /* cells the per-phonon loop flies through: num_cells, less one for every pair of
                                           triangles that flies as one parallelogram (option "merge_cells") */
    uint32_t lattice_cells;             /* cells of the lattice image flown by the launches that record nothing (blocks of identical
                                           parallelograms of one rate class as one cell each); 0: the model has none, or
                                           "merge_cells" < 2 */
    uint32_t lattice_recorded;          /* 1: the recorded launches of the last run flew the lattice image too (option "lattice_recorded") */
} psim_stats;

typedef struct psim_gpu psim_gpu;

/* Replaces ModelSimulator::ModelSimulator (modelSimulator.cpp:29-37) plus the references to cells, sensors
 * and materials the builders keep (phononBuilder.h:48,61).  Everything is copied; no pointer is retained.
 * device < 0 selects the current CUDA device. */
int psim_gpu_create(const psim_model_desc* desc, int device, psim_gpu** out);

/* Replaces ModelSimulator::initPhononBuilders (modelSimulator.cpp:43-85).  The per-source counts are computed
 * by the caller (host layer: seed-keyed stochastic rounding) so that they do not depend on the number of GPUs.
 * Phonon ids are assigned in source order; this handle simulates ids with id % num_shards == shard. */
int psim_gpu_set_sources(psim_gpu* h, const psim_source* sources, size_t n, uint64_t seed, uint32_t shard,
                         uint32_t num_shards);

/* Replaces ModelSimulator::runSimulation (modelSimulator.cpp:39-41): all measurement steps, blocking. */
int psim_gpu_run(psim_gpu* h);

/* The same, a range of measurement steps at a time and asynchronous on `cuda_stream` (a cudaStream_t, NULL =
 * the handle's own stream), so that a caller can overlap its per-step tally all-reduce.  Steps must be
 * submitted in order starting at 0; the last useful step is measurement_steps - 2 (the reference's final
 * interval records nothing, modelSimulator.cpp:183-188). */
int psim_gpu_run_steps(psim_gpu* h, uint32_t step_begin, uint32_t step_end, void* cuda_stream);
int psim_gpu_synchronize(psim_gpu* h);

/* Where the launch window that starts at measurement step `step_begin` would end if nothing cut it short: a caller that
 * exchanges tallies between groups of steps (bench.py:one_job) cuts its groups at these boundaries, so that its
 * grouping adds no launch.  Needs psim_gpu_set_sources. */
int psim_gpu_next_window(psim_gpu* h, uint32_t step_begin, uint32_t* step_end);

/* Replaces reading Sensor::inc_energy_ / inc_flux_ (sensor.h:50-55,73-74).  energy: [num_sensors][R] signed
 * phonon counts; flux: [num_sensors][R][2] sum of sign * velocity (m/s); R = measurement_steps - step_adjustment.
 * flux_fixed (optional, may be NULL) receives the exact integer sums in units of 1/256 m/s. */
int psim_gpu_get_tallies(psim_gpu* h, int32_t* energy, double* flux, int64_t* flux_fixed);

/* Device-resident tallies for a caller-side collective (NCCL): energy int32 [R][num_sensors], flux int64
 * [R][num_sensors][2] fixed point.  The pointers stay valid until reset / destroy. */
int psim_gpu_tally_buffers(psim_gpu* h, void** energy_dev, void** flux_dev, uint32_t* recorded_steps,
                           uint32_t* num_sensors);

/* Bookkeeping: population of the pool after the steps submitted so far, and its histogram over cells. */
int psim_gpu_alive(psim_gpu* h, uint64_t* alive);
int psim_gpu_cell_histogram(psim_gpu* h, uint64_t* per_cell /* [num_cells] */);

int psim_gpu_get_stats(psim_gpu* h, psim_stats* out);

/* Tunables: "steps_per_launch" (measurement intervals advanced per pass over the pool; 0 = automatic, the default),
 * "warps_per_sm" (0 = occupancy-derived), "kernel" (2 work queues = default, 0 lane-bound shared-memory slots, 1 lock-step first version),
 * "queue_slots" (phonons in flight per warp of the work-queue kernel, 128 or 64; default by mesh size; before set_sources),
 * "merge_cells" (1: two triangles of one sensor area whose union is a parallelogram fly as ONE cell - crossing their
 * shared edge does nothing to a phonon in the reference either, surface.cpp:71-75; 2 = default: also, in launches that
 * record nothing, every rectangular block of identical parallelograms of one material and rate class flies as one lattice
 * cell - where no sensor is read, a transition between two such cells changes nothing but the cell label;
 * 0: one flight cell per triangle; before set_sources),
 * "lattice_recorded" (recorded windows of many-sensor models over the lattice cells too, the sensor area of every measurement a
 * flight segment crossed being found from the phonon's position at that instant: -1 = default, where a phonon crosses 1.25 or
 * more fine cells per measurement step at the model's largest group velocity; 0 never; 1 always; before set_sources),
 * "tally_shared" (-1 automatic = default; 0 straight to global memory, rows kept as differences along the step axis until
 * their window is complete; 1 / 4 / 2 staged per CTA in shared memory as two / three 32-bit parts / 64-bit sums - 1 and 4
 * fall back to the next form when their exactness bound does not hold; before set_sources). */
int psim_gpu_set_option(psim_gpu* h, const char* name, int64_t value);

/* Replaces ModelSimulator::reset (modelSimulator.h:20-23) + Sensor::reset (sensor.cpp:54-60). */
int psim_gpu_reset(psim_gpu* h);
void psim_gpu_destroy(psim_gpu* h);
const char* psim_gpu_last_error(const psim_gpu* h); /* h may be NULL: error of the last failed create */

/* Device memory of destroyed handles (phonon pools, tally buffers) is kept for the next handle of the process - a model is
 * usually run more than once, and cudaMalloc / cudaFree of a multi-GB pool cost more than the kernels of most models
 * (environment PSIM_DEVICE_CACHE_MB: how much is kept per process, default 16384, 0 = nothing).  This returns all of it to
 * the driver.  No counterpart in the reference (it has no device). */
void psim_gpu_release_cached(void);

/* Per-function probes used by the parity tests: run the device implementation of one reference function
 * on caller-supplied inputs.  out_bin/out_ta: Material::freqIndex (material.cpp:64-75) for uniforms u1,u2 as
 * the flight loop computes it (guided search); out_bin_bisect: the reference's plain bisection on the same
 * fp32 table - the two must be identical. */
int psim_gpu_probe_sample(psim_gpu* h, uint32_t table, const float* u1, const float* u2, size_t n,
                          uint32_t* out_bin, uint32_t* out_ta, uint32_t* out_bin_bisect);
/* rates: [n][3] = (N, U, I) in 1/s for sensor `sensor` (Material::relaxRates, material.cpp:54-57). */
int psim_gpu_probe_rates(psim_gpu* h, uint32_t sensor, const double* omega, const uint32_t* ta, size_t n,
                         double* rates);
/* One free flight + specular reflection, the geometry core of ModelSimulator::nextImpact (modelSimulator.cpp:87-122,
 * Line::getIntersection geometry.cpp:104-138) and Surface::boundaryHandlePhonon (surface.cpp:32-44): for phonon i
 * in cell cell[i] at barycentric position (b1, b2) moving with velocity (vx, vy) nm/ns, the device returns the edge
 * it reaches first, the time, the barycentric position of the hit and the mirrored direction about that edge's
 * inward normal.  in: [n][4] = b1, b2, vx, vy; out: [n][6] = edge, time, b1, b2, dx', dy' (direction of unit input). */
int psim_gpu_probe_flight(psim_gpu* h, const uint32_t* cell, const float* in, size_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif
