// Host-side preparation of the device image: psim_model_desc (C ABI, include/psim_b200.h) -> arrays of the
// records in device_types.h, and the birth plan (which phonon indices each measurement step creates).
// Pure C++17, no CUDA calls: shared by the GPU library and by the CPU-side test emulation.
#ifndef PSIM_B200_FLATTEN_H
#define PSIM_B200_FLATTEN_H

#include "../../include/psim_b200.h"
#include "device_types.h"
#include <string>
#include <vector>

namespace psim {

struct HostImage {
    std::vector<DevCell> cells;       // flight cells: model triangles, or parallelograms made of two (device_types.h)
    std::vector<DevApiCell> api_cells; // [model cells]
    std::vector<DevShape> shapes;     // distinct (geometry, specularity) records
    std::vector<DevSensor> classes;   // one record per rate class (at most 255)
    std::vector<DevSensor> step_sensors;  // empty, or [sensors][steps] (psim_model_desc::step_sensors)
    std::vector<DevSub> subs;
    // The lattice image (device_types.h): the same mesh with every rectangular block of identical parallelograms of one rate
    // class as ONE flight cell; flown by launches that record nothing.  Empty when no block of more than one cell exists.
    std::vector<DevCell> lattice_cells;
    std::vector<DevApiCell> lattice_api_cells;
    std::vector<DevShape> lattice_shapes;
    std::vector<DevSub> lattice_subs;
    std::vector<DevEmitter> lattice_emitters;
    std::vector<uint32_t> lattice_sub_fine;  // DevParams::sub_fine
    std::vector<uint32_t> lattice_sub_sensor;  // DevParams::sub_sensor
    // how many fine cells a phonon at the model's largest group velocity crosses per measurement step, averaged over the
    // cells that are part of a block (0 without lattice image): where it is large, flying the lattice image pays in recorded
    // windows too (psim_gpu.cu)
    double lattice_cells_per_step = 0.;
    bool fast_links = true, lattice_fast_links = true;  // DevParams::fast_links of either image
    std::vector<DevSensor> sensors;
    std::vector<DevMaterial> materials;
    std::vector<DevEmitter> emitters;
    std::vector<float2> tables;      // [n_tables][PSIM_BINS]
    std::vector<uint32_t> guides;    // [n_tables][PSIM_GUIDE]
    std::vector<float> velocities;   // [n_materials][2][PSIM_BINS]
    std::vector<double> sensor_temperature;
    std::vector<psim_material> material_consts;
    DevParams scalars{};             // pointers left null; counts and settings filled in
};

struct BirthPlan {
    std::vector<DevSource> sources;
    std::vector<DevBirth> births;         // sorted by step, then by source
    std::vector<uint64_t> prefix;         // births.size() + 1, exclusive running sum of DevBirth::count
    std::vector<uint32_t> step_begin;     // num_steps + 1: first entry of each step
    uint64_t total_phonons = 0;           // all shards
    uint64_t shard_phonons = 0;           // created by this shard (excludes those born in the unrecorded last step)
    uint64_t shard_unrecorded = 0;        // owned by this shard but born in the last interval (never tallied)
};

// returns 0 or a PSIM_E_* code with `err` set
// merge_cells = 0 keeps one flight cell per model triangle (the per-function probes and the A/B option "merge_cells"),
// 1 pairs triangles into parallelograms, 2 (default) also builds the lattice image
int flatten_model(const psim_model_desc& d, HostImage& out, std::string& err, int merge_cells = 2);
int plan_births(const HostImage& img, const psim_source* sources, size_t n, uint32_t shard, uint32_t num_shards,
                BirthPlan& out, std::string& err);

}  // namespace psim
#endif
