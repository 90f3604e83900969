/* psim_host - C ABI of the host layer that surrounds the GPU particle loop (include/psim_b200.h).
 *
 * It reproduces, on the CPU, what the reference does around ModelSimulator so that the same model files give
 * the same result tables: the loader (reference psim/src/inputManager.cpp:14-115), the run driver
 * (Model::runSimulation, psim/src/model.cpp:141-182), the tally interpreter (psim/src/sensorInterpreter.cpp:
 * 19-112) and the exporter (psim/src/outputManager.cpp:13-130).  None of these functions touches a GPU except
 * psim_model_run, which drives psim_gpu_* on one device; multi-GPU callers use psim_model_describe /
 * psim_model_sources / psim_model_set_tallies / psim_model_finish_run around their own per-rank psim_gpu handle.
 *
 * Threads: the end of an iteration, the run epilogue and the periodic exporter go over the sensors / step groups on up to
 * 16 threads of the host where the tallies are large (every sensor's sums still run over its steps in order: the numbers
 * are those of one thread, bit for bit); environment PSIM_HOST_THREADS caps the number (1 = none).  PSIM_TIMING=1 prints
 * the phase times of psim_model_run on stderr.
 *
 * Errors: 0 or a negative PSIM_E_* code; psim_host_last_error() returns the message of the calling thread's
 * last failure.  The conditions and messages for rejected model files are the reference's.
 */
#ifndef PSIM_HOST_H
#define PSIM_HOST_H

#include "psim_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct psim_model psim_model;

typedef struct psim_model_info {
    uint64_t num_runs, measurement_steps, recorded_steps, num_phonons, step_interval;
    double simulation_time, t_eq;
    uint32_t sim_type;        /* 0 steady state, 1 periodic, 2 transient (utils.h:9) */
    uint32_t phasor_sim;
    uint32_t num_materials, num_sensors, num_cells, num_emitters;
    uint32_t num_transition_links;   /* directed cell->neighbour links found */
    uint32_t num_partial_links;      /* of those, links that cover only part of an edge */
} psim_model_info;

const char* psim_host_last_error(void);

/* InputManager::deserialize (inputManager.cpp:14-115). */
int psim_model_load(const char* json_path, psim_model** out);
int psim_model_load_text(const char* json_text, psim_model** out);
void psim_model_free(psim_model* m);
int psim_model_get_info(const psim_model* m, psim_model_info* out);
/* Overrides of the file's settings, for reduced-size parity runs. */
int psim_model_set_num_phonons(psim_model* m, uint64_t n);
int psim_model_set_num_runs(psim_model* m, uint64_t n);
/* Iterations per run: the reference's MAX_ITERS (model.cpp:11), compiled in as 1 upstream - which leaves the re-iteration
 * of model.cpp:159-172 (new t_eq, sensor temperatures, tables) unreachable.  Also the optional settings key "max_iters". */
int psim_model_set_max_iters(psim_model* m, uint64_t n);

/* Start of a run: temperature bounds, material tables, energy per phonon (model.cpp:145-155). */
int psim_model_prepare(psim_model* m);
int psim_model_energy(psim_model* m, double* total_energy, double* energy_per_phonon);
/* Known-answer access to the material data (material.cpp:20-51,101-204). kind: 0 base, 1 emit, 2 scatter. */
int psim_model_material_arrays(psim_model* m, uint32_t material, double* freq, double* vel_la, double* vel_ta,
                               double* dens_la, double* dens_ta /* each [1000] */);
int psim_model_table(psim_model* m, uint32_t material, uint32_t kind, double temperature, double* cumulative,
                     double* la_fraction /* each [1000] */, double* sum);
int psim_model_cell_energies(psim_model* m, double* area, double* init_energy, double* emit_energy /* [cells] */);
int psim_model_sensor_ids(const psim_model* m, uint64_t* ids /* [sensors], model order */, double* areas);

/* Flat description for psim_gpu_create; valid until the next prepare / free. */
int psim_model_describe(psim_model* m, const psim_model_desc** out);
/* initPhononBuilders' integer bookkeeping (modelSimulator.cpp:43-85), a pure function of (model, seed).
 * Call with sources == NULL to get the count. */
int psim_model_sources(psim_model* m, uint64_t seed, psim_source* sources, size_t* n);

/* What the hot path produced (layouts of psim_gpu_get_tallies), then the run epilogue (model.cpp:163-177). */
int psim_model_set_tallies(psim_model* m, const int32_t* energy, const double* flux);
/* End of one simulated iteration (model.cpp:163-171): *again = 1 if the run has to be simulated once more - the model has
 * then been reset(false) to the new temperatures and t_eq, and psim_model_describe / psim_model_sources give the next
 * iteration's inputs.  psim_model_finish_run (storeResults) ends the iteration itself if the caller has not. */
int psim_model_end_iteration(psim_model* m, int* again);
int psim_model_finish_run(psim_model* m, uint64_t run_id, int* stable_sensors);
int psim_model_next_run(psim_model* m);  /* reset(true), model.cpp:178-180 */

/* One-GPU convenience: all runs of the model on `device` (prints the reference's progress lines to stdout
 * when verbose != 0).  seed: run r uses seed + r.  steps_per_launch <= 0 keeps the library default. */
int psim_model_run(psim_model* m, int device, uint64_t seed, int steps_per_launch, int verbose, psim_stats* stats);

/* The same over several GPUs of this process: device d simulates the phonon ids == d (mod n_devices), one host thread
 * per device, and the integer tallies are summed on the host (the payload is recorded_steps x sensors x 20 bytes).
 * The result is bit-identical to the one-GPU run with the same seed.  stats: summed over the devices (kernel_ms: max). */
int psim_model_run_devices(psim_model* m, const int* devices, int n_devices, uint64_t seed, int steps_per_launch,
                           int verbose, psim_stats* stats);

/* Results of run `run_id` (sensors sorted by id): six[S][6] = T, stdT, qx, std qx, qy, std qy;
 * temps[S][R]; fluxes[S][R][2]; any pointer may be NULL.  run_id == UINT64_MAX: average over runs. */
int psim_model_results(const psim_model* m, uint64_t run_id, double* six, double* temps, double* fluxes);
double psim_model_energy_per_phonon(const psim_model* m);
/* OutputManager::exportResults: writes ss_<stem>.txt / per_<stem>.txt next to model_path. */
int psim_model_export(const psim_model* m, const char* model_path, double seconds);
/* Same text into a caller buffer (returns needed size in *len). `when` replaces the time stamp. */
int psim_model_export_text(const psim_model* m, const char* model_filename, double seconds, const char* when,
                           char* buf, size_t* len);

#ifdef __cplusplus
}
#endif
#endif
