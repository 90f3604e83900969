// Host side of the drop-in: everything the reference does AROUND the hot path, so that `psim model.json`
// reads the same files and writes the same tables.  C++17, CPU only, no CUDA in this translation unit.
//   - model file -> materials / sensors / cells / emitting surfaces   (reference inputManager.cpp:14-115)
//   - geometry set-up: neighbour discovery, surface attachment        (model.cpp:98-138, cell.cpp:29-35,81-98,
//                                                                       compositeSurface.cpp:20-45)
//   - material tables                                                 (material.cpp:20-51,101-204,241-246)
//   - energy bookkeeping and phonons per source                       (model.cpp:148-153,196-201, cell.cpp:38-63,
//                                                                       modelSimulator.cpp:43-85)
//   - tallies -> temperatures / fluxes, run epilogue and its quirks   (model.cpp:163-177,250-272,
//                                                                       sensorInterpreter.cpp:19-112)
//   - result files                                                    (outputManager.cpp:13-130)
#ifndef PSIM_B200_HOST_MODEL_H
#define PSIM_B200_HOST_MODEL_H

#include "../../../include/psim_b200.h"
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace psim {

constexpr int kBins = PSIM_NUM_FREQ_BINS;
enum class SimType { SteadyState = 0, Periodic = 1, Transient = 2 };

struct Table {
    std::array<double, kBins> cumulative{};
    std::array<double, kBins> la_fraction{};
    double sum = 0.;  // energy / heat capacity / emitted power / scattering-weighted energy per unit volume
};

class Material {
public:
    std::string name;
    uint32_t id = 0;
    double la[3]{}, ta[3]{};
    double w_max_la = 0., w_max_ta = 0.;
    double b_l = 0., b_tn = 0., b_tu = 0., b_i = 0., w = 0.;
    double freq_width = 0.;
    bool full_simulation = false;
    std::array<double, kBins> freq{}, vel_la{}, vel_ta{}, dens_la{}, dens_ta{};

    void build_dispersion();                              // Material::Material, material.cpp:20-51
    void set_temperature_grid(double low, double high);   // initializeTables' grid, material.cpp:101-109
    size_t temp_index(double temp) const;                 // getTempIndex, material.cpp:241-246
    double grid_temperature(size_t idx) const { return temps_[idx]; }
    enum Kind { Base = 0, Emit = 1, Scatter = 2 };
    const Table& table(Kind kind, double temp);           // baseTable / emitTable / scatterTable (lazy, cached)
    double base_energy(double temp) { return table(Base, temp).sum; }
    double emit_energy(double temp) { return table(Emit, temp).sum; }
    double scatter_energy(double temp) { return table(Scatter, temp).sum; }
    std::array<double, 3> relax_rates(double temp, double omega, bool ta) const;  // material.cpp:54-57,207-239

private:
    std::vector<double> temps_;
    std::map<std::pair<int, size_t>, std::unique_ptr<Table>> cache_;
    std::array<double, kBins> phonon_dist(double temp, bool ta) const;  // material.cpp:184-204
};

struct SubSurface {
    uint32_t kind;            // PSIM_SURF_TRANSITION / PSIM_SURF_EMIT
    uint32_t target;          // neighbour cell / emitter index
    uint32_t target_edge;
    double s0, s1, t0, t1;
};

struct CellRec {
    double x[3], y[3];
    uint32_t sensor;          // index into Model::sensors
    double spec;
    double area;
    std::vector<SubSurface> transitions[3];
    std::vector<SubSurface> emits[3];
};

struct SensorRec {
    uint64_t id;
    uint32_t material;
    double t_init;
    double t_steady;          // SensorController::t_steady_
    double heat_capacity;     // heat_capacity_ (set from t_init by updateTables)
    double table_temp;        // temperature base_table_ / scatter_table_ are taken at: t_init, after a re-iteration t_steady
    double area = 0.;
    std::vector<double> steady_temps;     // transient controller only: steady_temps_[step]
    std::vector<double> heat_capacities;  // transient controller only: heat_capacities_[step]
};

struct EmitRec {
    double p1x, p1y, p2x, p2y;
    double temp, duration, start;
    uint32_t cell, edge;
    double s_p1, s_p2;
    double length;
};

struct SensorResult {  // reference SensorMeasurements, sensor.h:76-88
    uint64_t id = 0;
    double t_steady = 0., std_t_steady = 0., x_flux = 0., std_x_flux = 0., y_flux = 0., std_y_flux = 0.;
    std::vector<double> final_temps;
    std::vector<std::array<double, 2>> final_fluxes;
};

class Model {
public:
    // settings (inputManager.cpp:17-30)
    uint64_t num_runs = 1, measurement_steps = 0, num_phonons = 0;
    // Iterations per run (the reference's MAX_ITERS, model.cpp:11: compiled in as 1, which leaves its re-iteration - new
    // t_eq, new sensor temperatures and tables, model.cpp:159-172 - unreachable; here a setting, settings key "max_iters")
    uint64_t max_iters = 1;
    double simulation_time = 0., t_eq = 0.;
    bool phasor_sim = false;
    SimType sim_type = SimType::SteadyState;
    uint64_t step_interval = 0;
    uint64_t start_step = 0;        // Model::start_step_
    uint64_t step_adjustment = 0;   // ModelSimulator::step_adjustment_
    uint64_t recorded_steps = 0;    // length of each sensor's tally vectors

    std::vector<Material> materials;
    std::vector<SensorRec> sensors;
    std::vector<CellRec> cells;
    std::vector<EmitRec> emitters;

    static std::unique_ptr<Model> from_json_text(const std::string& text);   // throws std::runtime_error
    static std::unique_ptr<Model> from_file(const std::string& path);

    // --- per run, in the order Model::runSimulation uses them (model.cpp:141-182) ---
    void prepare();                                   // setTemperatureBounds + initializeMaterialTables
    double total_initial_energy();                    // getTotalInitialEnergy
    double energy_per_phonon() const { return eff_energy_; }
    void refresh();                                   // the `refresh` lambda, model.cpp:148-153
    std::vector<psim_source> source_counts(uint64_t seed);   // initPhononBuilders' integer bookkeeping
    void set_tallies(const int32_t* energy, const double* flux);   // [S][R], [S][R][2] (what the hot path produced)
    // The same without the copy: the model's own tally storage ([S][R] energies, [S][R][2] fluxes; kept from run to run, so
    // the 31 MB of a 3108-sensor model are allocated and paged in once) for the caller to fill, then tallies_written().
    std::pair<int32_t*, double*> tally_storage();
    void tallies_written();
    // End of one simulated iteration (model.cpp:163-171): resetRequired(); if the sensors or t_eq moved and max_iters
    // allows another iteration, reset(false) - tallies cleared, tables and heat capacities at the new temperatures - and the
    // new t_eq; then refresh().  Returns true if the run must be simulated again (with a fresh describe()).
    bool end_iteration(std::string* log);
    int finish_run(uint64_t run_id, std::string* log);             // model.cpp:173-177 (ends the iteration first if the caller
                                                                   // has not); returns the stable-sensor count
    void restore_file_state();                                     // t_eq and sensors as the model file gives them
    void reset_for_next_run();                                     // reset(true), model.cpp:178-180,274-283

    // flat description for psim_gpu_create (pointers stay valid until the next prepare())
    const psim_model_desc& describe();

    // results (one entry per completed run, sensors sorted by id)
    std::vector<std::vector<SensorResult>> runs;
    std::vector<SensorResult> averaged() const;       // calculateAverages, with the traces kept
    void export_results(const std::string& model_path, double seconds) const;   // outputManager.cpp:13-38
    std::string export_text(const std::string& model_filename, double seconds, const std::string& when) const;

    double temp_lo() const { return temp_lo_; }
    double temp_hi() const { return temp_hi_; }

private:
    double temp_lo_ = 0., temp_hi_ = 0., lb_ = 0., ub_ = 0.;
    double eff_energy_ = 0.;
    double t_eq_file_ = 0.;
    bool prepared_ = false;
    bool iteration_ended_ = false;  // end_iteration() has consumed the tallies set last
    uint64_t iter_ = 0;             // iterations of the current run simulated so far
    int stable_ = 0;
    std::vector<psim_sensor> d_step_sensors_;
    std::vector<int32_t> inc_energy_;   // [S][R]     Sensor::inc_energy_ of every sensor, flat
    std::vector<double> inc_flux_;      // [S][R][2]  Sensor::inc_flux_
    bool have_tallies_ = false;

    // describe() storage
    psim_model_desc desc_{};
    std::vector<psim_material> d_materials_;
    std::vector<double> d_velocities_;
    std::vector<psim_sensor> d_sensors_;
    std::vector<psim_cell> d_cells_;
    std::vector<psim_subsurface> d_subs_;
    std::vector<psim_emitter> d_emitters_;
    std::vector<psim_table> d_tables_;
    std::vector<const Table*> table_list_;
    std::map<const Table*, uint32_t> table_index_;
    uint32_t table_id(const Table& t);

    void build_geometry();
    void attach_emit_surface(double p1x, double p1y, double p2x, double p2y, double temp, double duration, double start);
    double heat_capacity_at(const SensorRec& s, size_t step) const {  // getHeatCapacity(step), sensorController.h:62,80,96
        return (sim_type == SimType::Transient && step < s.heat_capacities.size()) ? s.heat_capacities[step] : s.heat_capacity;
    }
    double steady_temp(const SensorRec& s) const { return sim_type == SimType::Transient ? s.t_init : s.t_steady; }  // getSteadyTemp(0)
    void reset_iteration();                                                   // Model::reset(false), model.cpp:274-283
    double init_temp(const SensorRec& s) const;
    std::vector<double> find_temperature(size_t sensor, size_t start_step);   // sensorInterpreter.cpp:80-112
    // f(sensor) for every sensor, on several threads where the tallies are large and f touches nothing but its own sensor
    // and the tallies (full mode inverts through the materials' lazily built tables: one thread)
    template <class F> void for_each_sensor(F&& f);
    SensorResult scale_heat_params(size_t sensor);                            // sensorInterpreter.cpp:19-66
};

}  // namespace psim
#endif
