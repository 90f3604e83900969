// TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// A small driver `main` of our own that is compiled TOGETHER WITH the unmodified reference
// sources where they lie under /root/reference/psim/src (see oracle/Makefile; outputs only into
// oracle/_ref/).  The reference's own main.cpp is not used: its periodic/transient export
// segfaults at HEAD (outputManager.cpp:40-70,85-86), and it never exposes raw tallies.
// Built with -fno-access-control so that the private members the hot path writes
// (Sensor::inc_energy_/inc_flux_, sensor.h:73-74) can be dumped without touching reference files.
//
// usage: psim_ref run <model.json> <out_prefix>   -> <out_prefix>.meta.json + <out_prefix>.bin
//        psim_ref kat <model.json> <out_prefix>   -> deterministic table / energy known answers
#include "psim/inputManager.h"
#include "psim/model.h"
#include "psim/timer.h"
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <set>
#include <vector>

namespace {
template<typename T> void put(std::ofstream& f, const T* p, std::size_t n) {
    f.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n * sizeof(T)));
}

int runMode(const std::string& json, const std::string& prefix) {
    auto model = InputManager::deserialize(json);
    if (!model) { return 2; }
    // same pre-run quantities Model::runSimulation computes at model.cpp:145-155 (on a second instance so
    // the run below starts from a pristine model)
    double total_energy = 0.;
    {
        auto probe = InputManager::deserialize(json);
        const auto [lo, hi] = probe->setTemperatureBounds();
        probe->initializeMaterialTables(lo, hi);
        total_energy = probe->getTotalInitialEnergy();
    }
    Timer timer;
    model->runSimulation();
    const double secs = timer.get_time_diff();

    std::vector<const Sensor*> sensors;
    for (const auto& s : model->sensors_) { sensors.push_back(&s); }
    std::sort(sensors.begin(), sensors.end(), [](auto* a, auto* b) { return a->getID() < b->getID(); });
    const auto& meas = model->outputManager_.measurements_.at(0);// sorted by id (model.cpp:247)
    const std::size_t S = sensors.size();
    const std::size_t R = sensors.front()->getEnergies().size();

    std::ofstream bin(prefix + ".bin", std::ios::binary | std::ios::trunc);
    for (auto* s : sensors) { put(bin, s->getEnergies().data(), R); }// int32 [S][R]
    for (auto* s : sensors) { put(bin, s->getFluxes().data()->data(), 2 * R); }// f64 [S][R][2]
    for (const auto& m : meas) { put(bin, m.final_temps.data(), R); }// f64 [S][R]
    for (const auto& m : meas) { put(bin, m.final_fluxes.data()->data(), 2 * R); }// f64 [S][R][2]
    for (const auto& m : meas) {// f64 [S][6]  (the six ss_*.txt columns, outputManager.cpp:72-78)
        const double row[6] = { m.t_steady, m.std_t_steady, m.x_flux, m.std_x_flux, m.y_flux, m.std_y_flux };
        put(bin, row, 6);
    }
    std::ofstream meta(prefix + ".meta.json", std::ios::trunc);
    meta << std::setprecision(17);
    meta << "{\"sensors\": " << S << ", \"recorded_steps\": " << R
         << ", \"num_phonons\": " << model->num_phonons_
         << ", \"total_phonons\": " << model->simulator_.total_phonons_
         << ", \"total_energy_pre\": " << total_energy
         << ", \"energy_per_phonon_pre\": " << total_energy / static_cast<double>(model->num_phonons_)
         << ", \"energy_per_phonon_post\": " << model->interpreter_.eff_energy_
         << ", \"t_eq\": " << model->t_eq_ << ", \"sim_type\": " << static_cast<int>(model->sim_type_)
         << ", \"seconds\": " << secs << ", \"sensor_ids\": [";
    for (std::size_t i = 0; i < S; ++i) { meta << (i ? "," : "") << sensors[i]->getID(); }
    meta << "], \"sensor_areas\": [";
    for (std::size_t i = 0; i < S; ++i) { meta << (i ? "," : "") << sensors[i]->getArea(); }
    meta << "]}\n";
    std::cout << "ref run: " << secs << " s, total_phonons " << model->simulator_.total_phonons_ << '\n';
    return 0;
}

int katMode(const std::string& json, const std::string& prefix) {
    auto model = InputManager::deserialize(json);
    if (!model) { return 2; }
    const auto [lo, hi] = model->setTemperatureBounds();
    model->initializeMaterialTables(lo, hi);
    const double total_energy = model->getTotalInitialEnergy();
    std::set<double> temps;
    for (const auto& c : model->cells_) {
        temps.insert(c.getInitTemp());
        for (const auto& b : c.getBoundaries()) {
            for (const auto& es : b.getEmitSurfaces()) { temps.insert(es.getTemp()); }
        }
    }
    // material ids are assigned in JSON order (inputManager.cpp:86-88); dump in id order
    std::vector<const Material*> mats;
    for (const auto& kv : model->materials_) { mats.push_back(&kv.second); }
    std::sort(mats.begin(), mats.end(), [](auto* a, auto* b) { return a->id() < b->id(); });

    std::ofstream bin(prefix + ".bin", std::ios::binary | std::ios::trunc);
    std::ofstream meta(prefix + ".meta.json", std::ios::trunc);
    meta << std::setprecision(17);
    meta << "{\"total_energy\": " << total_energy << ", \"temp_lo\": " << lo << ", \"temp_hi\": " << hi
         << ", \"num_materials\": " << mats.size() << ", \"temps\": [";
    bool first = true;
    for (double t : temps) { meta << (first ? "" : ",") << t; first = false; }
    meta << "], \"layout\": \"per material: freq[1000] vel_la[1000] vel_ta[1000] dens_la[1000] dens_ta[1000]; then per "
            "temp: base[1000][2] emit[1000][2] scatter[1000][2] (f64)\", \"sums\": [";
    bool firstm = true;
    for (auto* m : mats) {
        put(bin, m->frequencies_.data(), 1000);
        put(bin, m->velocities_la_.data(), 1000);
        put(bin, m->velocities_ta_.data(), 1000);
        put(bin, m->densities_la_.data(), 1000);
        put(bin, m->densities_ta_.data(), 1000);
        meta << (firstm ? "" : ",") << "[";
        firstm = false;
        bool firstt = true;
        for (double t : temps) {
            put(bin, &(*m->baseTable(t))[0].first, 2000);
            put(bin, &(*m->emitTable(t))[0].first, 2000);
            put(bin, &(*m->scatterTable(t))[0].first, 2000);
            meta << (firstt ? "" : ",") << "[" << m->baseEnergy(t) << "," << m->emitEnergy(t) << ","
                 << m->scatterEnergy(t) << "]";
            firstt = false;
        }
        meta << "]";
    }
    meta << "], \"cell_areas\": [";
    for (std::size_t i = 0; i < model->cells_.size(); ++i) { meta << (i ? "," : "") << model->cells_[i].getArea(); }
    meta << "], \"cell_init_energy\": [";
    for (std::size_t i = 0; i < model->cells_.size(); ++i) {
        meta << (i ? "," : "") << model->cells_[i].getInitEnergy(model->t_eq_);
    }
    meta << "], \"cell_emit_energy\": [";
    for (std::size_t i = 0; i < model->cells_.size(); ++i) {
        meta << (i ? "," : "") << model->cells_[i].getEmitEnergy(model->t_eq_);
    }
    meta << "]}\n";
    return 0;
}
}// namespace

int main(int argc, char** argv) {
    if (argc != 4) {
        std::cerr << "usage: psim_ref run|kat <model.json> <out_prefix>\n";
        return 1;
    }
    const std::string mode = argv[1];
    try {
        if (mode == "run") { return runMode(argv[2], argv[3]); }
        if (mode == "kat") { return katMode(argv[2], argv[3]); }
    } catch (const std::exception& e) {
        std::cerr << e.what() << '\n';
        return 3;
    }
    return 1;
}
