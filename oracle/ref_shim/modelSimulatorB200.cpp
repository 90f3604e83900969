// TEST INFRASTRUCTURE — the reference-side half of the drop-in, compiled and run, not sketched.
//
// This file implements the public methods of the reference's `ModelSimulator`
// (/root/reference/psim/include/psim/modelSimulator.h:12-41, used by Model at psim/src/model.cpp:160-161) on top of the
// C ABI of include/psim_b200.h.  oracle/Makefile (target `ref_b200`) compiles it TOGETHER WITH the unmodified reference
// sources - every file of psim/src except modelSimulator.cpp, which this replaces, including the reference's own
// main.cpp - into oracle/_ref/psim_ref_b200 and links libpsim_b200.so.  The result is the reference's `psim` program
// (its JSON loader, mesh validation, tables, run epilogue and exporter, untouched) with the particle loop on the GPU:
// tests/test_gpu_dropin.py runs models through it and compares the ss_*.txt it writes with the reference fixtures.
//
// Built with -fno-access-control (as oracle/ref_harness is) so that the flattening below can read what Model hands to
// the simulator through private members; a maintainer of the reference would add the half-dozen accessors instead.
// Nothing here is linked into libpsim_b200.so, and nothing in the product path includes it.
#include "psim/modelSimulator.h"

#include "psim/cell.h"
#include "psim/compositeSurface.h"
#include "psim/material.h"
#include "psim/sensor.h"
#include "psim/surface.h"
#include "psim/utils.h"

#include "../../include/psim_b200.h"

#include <cmath>
#include <cstdlib>
#include <map>
#include <random>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace {

// The reference's class has no member for what the GPU path must remember between initPhononBuilders and
// runSimulation (modelSimulator.h is used unmodified), so it is kept beside the object.
struct ShimState {
    std::vector<Cell>* cells = nullptr;
    std::vector<psim_source> sources;
};
std::unordered_map<const ModelSimulator*, ShimState>& states() {
    static std::unordered_map<const ModelSimulator*, ShimState> s;
    return s;
}

using Geometry::Line;
using Geometry::Point;

// position of point q along the segment a -> b, as a fraction
double fraction(const Point& a, const Point& b, const Point& q) {
    const double dx = b.x - a.x, dy = b.y - a.y;
    return ((q.x - a.x) * dx + (q.y - a.y) * dy) / (dx * dx + dy * dy);
}

template<typename T> uint32_t index_of(std::vector<const T*>& v, const T* p) {
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i] == p) { return static_cast<uint32_t>(i); }
    }
    v.push_back(p);
    return static_cast<uint32_t>(v.size() - 1);
}

}  // namespace

ModelSimulator::ModelSimulator(std::size_t measurement_steps, double simulation_time, bool phasor_sim)
    : step_time_{ simulation_time / static_cast<double>(measurement_steps) }
    , phasor_sim_{ phasor_sim } {
    step_times_.resize(measurement_steps);  // as modelSimulator.cpp:29-37; size and back() carry M and the run time
    for (std::size_t n = 0; n < measurement_steps; ++n) {
        step_times_[n] = static_cast<double>(n + 1) * simulation_time / static_cast<double>(measurement_steps);
    }
}

// modelSimulator.cpp:43-85: phonons per source by stochastic rounding, same source order (cells, their edges, the
// emitting sub-surfaces of each edge) and the same random draws (Utils::urand).  The emitter index of a sub-surface is
// its position in that enumeration - runSimulation() lists the emitters in the same order.
void ModelSimulator::initPhononBuilders(std::vector<Cell>& cells, double t_eq, double eff_energy) noexcept {
    ShimState& st = states()[this];
    st.cells = &cells;
    st.sources.clear();
    auto phonons = [&eff_energy](double energy) {
        double whole = 0;
        const double frac = std::modf(energy / eff_energy, &whole);
        auto n = static_cast<std::size_t>(whole);
        return (Utils::urand() < frac) ? n + 1 : n;
    };
    uint32_t emitter = 0;
    for (std::size_t c = 0; c < cells.size(); ++c) {
        const Cell& cell = cells[c];
        if (const auto n = phonons(cell.getInitEnergy(t_eq)); n > 0) {
            total_phonons_ += n;
            st.sources.push_back(psim_source{ PSIM_SRC_CELL, static_cast<uint32_t>(c), cell.getInitTemp() > t_eq ? 1 : -1, 0, n });
        }
        for (const auto& boundary : cell.getBoundaries()) {
            for (const auto& es : boundary.getEmitSurfaces()) {
                const double temp = es.getTemp();
                const double factor = cell.getMaterial().emitEnergy(temp) * es.getEmitDuration() * es.getLength() / 4.;
                const auto n = phonons(t_eq == 0. ? factor : factor * std::fabs(t_eq - temp));
                total_phonons_ += n;
                st.sources.push_back(psim_source{ PSIM_SRC_SURFACE, emitter, temp > t_eq ? 1 : -1, 0, n });
                ++emitter;
            }
        }
    }
}

// modelSimulator.cpp:39-41 (+ :227-254): the whole particle loop.  Flattens what Model owns into a psim_model_desc,
// runs it on the GPU and writes the tallies where Sensor::updateHeatParams (sensor.cpp:43-52) would have put them.
void ModelSimulator::runSimulation(double t_eq) {
    ShimState& st = states()[this];
    if (st.cells == nullptr) { throw std::runtime_error("runSimulation before initPhononBuilders\n"); }
    std::vector<Cell>& cells = *st.cells;

    std::vector<const Material*> mats;
    std::vector<const Sensor*> sensors;
    std::vector<const Material::Table*> tables;
    std::vector<psim_cell> pc(cells.size());
    std::vector<psim_subsurface> subs;
    std::vector<psim_emitter> emitters;
    std::vector<psim_sensor> ps;

    auto sensor_index = [&](const Sensor& s) -> uint32_t {
        const size_t before = sensors.size();
        const uint32_t i = index_of(sensors, &s);
        if (sensors.size() != before) {
            const SensorController& ctl = *s.controller_;
            psim_sensor rec{};
            rec.material = index_of(mats, &ctl.material_);
            rec.base_table = index_of(tables, ctl.base_table_);
            // transient controllers sample the table of the phonon's measurement step (sensorController.cpp:84-88); within
            // one run they are all the table of t_init (updateTables, :38-51)
            rec.scatter_table = index_of(tables, ctl.scatter_tables_.empty() ? ctl.scatter_table_ : ctl.scatter_tables_.front());
            rec.temperature = s.getSteadyTemp(0);
            ps.push_back(rec);
        }
        return i;
    };

    for (std::size_t c = 0; c < cells.size(); ++c) {
        const Cell& cell = cells[c];
        psim_cell& o = pc[c];
        o = psim_cell{};
        const Point v[3] = { cell.cell_.p1, cell.cell_.p2, cell.cell_.p3 };
        for (int k = 0; k < 3; ++k) {
            o.x[k] = v[k].x;
            o.y[k] = v[k].y;
        }
        o.specularity = cell.boundaries_[0].main_surface_.getSpecularity();
        o.sensor = sensor_index(cell.sensor_);
        for (int k = 0; k < 3; ++k) {  // edge k joins vertex k to vertex (k + 1) % 3, as Cell::buildCompositeSurfaces (cell.cpp:114-125)
            const CompositeSurface& b = cell.boundaries_[k];
            const Point &a0 = v[k], &a1 = v[(k + 1) % 3];
            o.sub_first[k] = static_cast<uint32_t>(subs.size());
            for (const TransitionSurface& ts : b.transition_sub_surfaces_) {
                const Cell& other = ts.cell_;  // the cell a phonon enters (surface.cpp:71-75)
                const auto j = static_cast<uint32_t>(&other - cells.data());
                const Line& ln = ts.getSurfaceLine();
                psim_subsurface sb{};
                sb.kind = PSIM_SURF_TRANSITION;
                sb.target = j;
                sb.s0 = fraction(a0, a1, ln.p1);
                sb.s1 = fraction(a0, a1, ln.p2);
                // which edge of the neighbour the line lies on, and where
                const Point w[3] = { other.cell_.p1, other.cell_.p2, other.cell_.p3 };
                int edge = -1;
                for (int e = 0; e < 3 && edge < 0; ++e) {
                    if (Line{ w[e], w[(e + 1) % 3] }.contains(ln)) { edge = e; }
                }
                if (edge < 0) { throw std::runtime_error("transition surface lies on no edge of its neighbour\n"); }
                sb.target_edge = static_cast<uint32_t>(edge);
                sb.t0 = fraction(w[edge], w[(edge + 1) % 3], ln.p1);
                sb.t1 = fraction(w[edge], w[(edge + 1) % 3], ln.p2);
                subs.push_back(sb);
            }
            for (const EmitSurface& es : b.emit_sub_surfaces_) {
                const Line& ln = es.getSurfaceLine();
                psim_emitter em{};
                em.cell = static_cast<uint32_t>(c);
                em.edge = static_cast<uint32_t>(k);
                em.table = index_of(tables, &es.getTable());
                em.s_p1 = fraction(a0, a1, ln.p1);
                em.s_p2 = fraction(a0, a1, ln.p2);
                em.start_time = es.start_time_;
                em.duration = es.duration_;
                psim_subsurface sb{};
                sb.kind = PSIM_SURF_EMIT;
                sb.target = static_cast<uint32_t>(emitters.size());
                sb.s0 = em.s_p1;
                sb.s1 = em.s_p2;
                emitters.push_back(em);
                subs.push_back(sb);
            }
            o.sub_count[k] = static_cast<uint32_t>(subs.size()) - o.sub_first[k];
        }
    }

    std::vector<psim_material> pm(mats.size());
    std::vector<double> vel(mats.size() * 2 * Material::NUM_FREQ_BINS);
    for (std::size_t m = 0; m < mats.size(); ++m) {
        const Material& mt = *mats[m];
        pm[m] = psim_material{ mt.b_l_, mt.b_tn_, mt.b_tu_, mt.b_i_, mt.w_, mt.w_max_la_, mt.w_max_ta_, mt.freq_width_ };
        for (std::size_t i = 0; i < Material::NUM_FREQ_BINS; ++i) {
            vel[(2 * m) * Material::NUM_FREQ_BINS + i] = mt.velocities_la_[i];
            vel[(2 * m + 1) * Material::NUM_FREQ_BINS + i] = mt.velocities_ta_[i];
        }
    }
    std::vector<std::vector<double>> cum(tables.size()), laf(tables.size());
    std::vector<psim_table> pt(tables.size());
    for (std::size_t t = 0; t < tables.size(); ++t) {
        cum[t].resize(Material::NUM_FREQ_BINS);
        laf[t].resize(Material::NUM_FREQ_BINS);
        for (std::size_t i = 0; i < Material::NUM_FREQ_BINS; ++i) {
            cum[t][i] = (*tables[t])[i].first;
            laf[t][i] = (*tables[t])[i].second;
        }
        pt[t] = psim_table{ cum[t].data(), laf[t].data() };
    }

    psim_model_desc d{};
    d.num_materials = static_cast<uint32_t>(pm.size());
    d.num_sensors = static_cast<uint32_t>(ps.size());
    d.num_cells = static_cast<uint32_t>(pc.size());
    d.num_subsurfaces = static_cast<uint32_t>(subs.size());
    d.num_emitters = static_cast<uint32_t>(emitters.size());
    d.num_tables = static_cast<uint32_t>(pt.size());
    d.materials = pm.data();
    d.velocities = vel.data();
    d.sensors = ps.data();
    d.cells = pc.data();
    d.subsurfaces = subs.data();
    d.emitters = emitters.data();
    d.tables = pt.data();
    d.measurement_steps = static_cast<uint32_t>(step_times_.size());
    d.step_adjustment = static_cast<uint32_t>(step_adjustment_);
    d.simulation_time = step_times_.back();
    d.full_simulation = (t_eq == 0.) ? 1u : 0u;
    d.phasor_sim = phasor_sim_ ? 1u : 0u;

    const char* dev_env = std::getenv("PSIM_DEVICE");
    psim_gpu* h = nullptr;
    if (psim_gpu_create(&d, dev_env ? std::atoi(dev_env) : 0, &h) != PSIM_OK) {
        throw std::runtime_error(std::string("psim_b200: ") + psim_gpu_last_error(nullptr) + '\n');
    }
    auto check = [&](int rc) {
        if (rc != PSIM_OK) {
            const std::string msg = std::string("psim_b200: ") + psim_gpu_last_error(h) + '\n';
            psim_gpu_destroy(h);
            throw std::runtime_error(msg);
        }
    };
    std::random_device rd;  // the reference seeds from std::random_device too (utils.h:16-18)
    const char* seed_env = std::getenv("PSIM_SEED");
    const uint64_t seed = seed_env ? std::strtoull(seed_env, nullptr, 10) : ((static_cast<uint64_t>(rd()) << 32) | rd());
    check(psim_gpu_set_sources(h, st.sources.data(), st.sources.size(), seed, 0, 1));
    check(psim_gpu_run(h));
    const std::size_t S = ps.size(), R = d.measurement_steps - d.step_adjustment;
    std::vector<int32_t> e(S * R);
    std::vector<double> f(2 * S * R);
    check(psim_gpu_get_tallies(h, e.data(), f.data(), nullptr));
    psim_gpu_destroy(h);

    // Sensor::updateHeatParams (sensor.cpp:43-52) adds into inc_energy_[step - step_adjustment] / inc_flux_[...]
    for (std::size_t s = 0; s < S; ++s) {
        auto* sensor = const_cast<Sensor*>(sensors[s]);
        const std::size_t n = std::min<std::size_t>(R, sensor->inc_energy_.size());
        for (std::size_t r = 0; r < n; ++r) {
            sensor->inc_energy_[r] += e[s * R + r];
            sensor->inc_flux_[r][0] += f[2 * (s * R + r)];
            sensor->inc_flux_[r][1] += f[2 * (s * R + r) + 1];
        }
    }
}

// Not used outside modelSimulator.cpp upstream; defined so that the class is complete.
std::optional<double> ModelSimulator::nextImpact(Phonon&, double) const noexcept { return std::nullopt; }
