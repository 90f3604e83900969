// TEST INFRASTRUCTURE ONLY.  Host build of psim_b200/csrc/device_core.cuh: the same per-phonon functions the
// sm_100a kernels call, driven by a plain loop, so that the ALGORITHM (barycentric flight, stratified births,
// fp32 arithmetic, Philox streams) can be checked statistically against the reference's golden data in the
// GPU-less authoring container.  It is compiled into tests/emu/libpsim_emu.so by psim_b200/build.py:build_emu,
// is loaded only by tests/, and is never part of libpsim_b200.so (which has no CPU path).
#include "../../psim_b200/csrc/device_core.cuh"
#include "../../psim_b200/csrc/flatten.h"

#include <cstring>
#include <string>
#include <vector>

// 0: one flight cell per triangle, 1: pairs of triangles as parallelograms, 2 (the library's default): also the lattice image
// for the passes that record nothing
static int g_merge_cells = 2;
extern "C" void psim_emu_set_merge_cells(int level) { g_merge_cells = level; }
// recorded passes over the lattice image: -1 the library's rule (many sensors, 1.25 or more fine cells crossed per step), 0 / 1
static int g_lattice_recorded = -1;
extern "C" void psim_emu_set_lattice_recorded(int v) { g_lattice_recorded = v; }

extern "C" int psim_emu_run(const psim_model_desc* desc, const psim_source* sources, size_t n_sources, uint64_t seed,
                            uint32_t shard, uint32_t num_shards, uint32_t steps_per_pass, int32_t* energy /*[S][R]*/,
                            double* flux /*[S][R][2]*/, int64_t* flux_fixed, uint64_t* drift_steps, uint64_t* events,
                            uint64_t* alive_per_pass /* [M] or NULL */, uint64_t* cell_hist_steps /* [M][cells] or NULL */,
                            char* err, size_t err_len) {
    psim::HostImage img;
    psim::BirthPlan plan;
    std::string e;
    int rc = psim::flatten_model(*desc, img, e, g_merge_cells);
    if (!rc) { rc = psim::plan_births(img, sources, n_sources, shard, num_shards, plan, e); }
    if (rc) {
        if (err && err_len) { std::strncpy(err, e.c_str(), err_len - 1), err[err_len - 1] = 0; }
        return rc;
    }
    DevParams P = img.scalars;
    P.cells = img.cells.data();
    P.api_cells = img.api_cells.data();
    P.shapes = img.shapes.data();
    P.classes = img.classes.data();
    P.step_sensors = img.step_sensors.empty() ? nullptr : img.step_sensors.data();
    P.subs = img.subs.data();
    P.sensors = img.sensors.data();
    P.materials = img.materials.data();
    P.emitters = img.emitters.data();
    P.tables = img.tables.data();
    P.velocities = img.velocities.data();
    P.guides = img.guides.data();
    P.sources = plan.sources.data();
    P.n_sources = static_cast<uint32_t>(n_sources);
    P.seed_lo = static_cast<uint32_t>(seed);
    P.seed_hi = static_cast<uint32_t>(seed >> 32);
    // the same parameters over the lattice image (device_types.h), for the passes that record nothing
    const bool have_lattice = !img.lattice_cells.empty();
    DevParams PL = P;
    if (have_lattice) {
        PL.cells = img.lattice_cells.data();
        PL.api_cells = img.lattice_api_cells.data();
        PL.shapes = img.lattice_shapes.data();
        PL.subs = img.lattice_subs.data();
        PL.emitters = img.lattice_emitters.data();
        PL.sub_fine = img.lattice_sub_fine.data();
        PL.sub_sensor = img.lattice_sub_sensor.data();
        PL.lattice = 1u;
        PL.fast_links = img.lattice_fast_links ? 1u : 0u;
    }
    const DevParams PF = P;
    bool pool_in_lattice = false;
    const uint32_t S = P.n_sensors, R = P.recorded_steps, M = P.num_steps;
    std::vector<long long> te(static_cast<size_t>(R) * S, 0), tf(static_cast<size_t>(R) * S * 2, 0);
    std::vector<psim::Phonon> pool, next;
    uint64_t n_steps = 0;
    uint32_t n_events = 0;
    uint64_t total_events = 0;
    const uint32_t B = steps_per_pass ? steps_per_pass : 1;
    auto run_one = [&](psim::Phonon p, uint32_t start, float t_first, uint32_t s0, uint32_t s1) {
        uint32_t steps32 = 0;
        n_events = 0;
        const bool alive = psim::advance_window(P, p, t_first, start, s1, steps32, n_events,
                                                [&](uint32_t k0, uint32_t k1, const psim::Phonon& q, const psim::Flight& f) {
            const int sg = PSIM_PACK_NEG(q.packed) ? -1 : 1;
            auto add = [&](uint32_t ka, uint32_t kb, uint32_t sensor) {
                for (uint32_t ks = ka; ks < kb; ++ks) {
                    const size_t k = static_cast<size_t>(ks + 1 - P.first_tally_step) * S + sensor;
                    te[k] += sg;
                    tf[2 * k] += static_cast<long long>(psim::flux_fixed(q.dx)) * sg;
                    tf[2 * k + 1] += static_cast<long long>(psim::flux_fixed(q.dy)) * sg;
                }
            };
            if (P.lattice) {  // the sensor area of every crossed measurement from the position at that instant
                psim::lattice_runs(P, q.cell, q.b1, q.b2, f.r1, f.r2, f.t, k0, k1, add);
            } else {
                add(k0, k1, PSIM_CELL_SENSOR(f.sensor_mat));
            }
        });
        total_events += n_events;
        n_steps += steps32;
        (void)s0;
        if (alive) { next.push_back(p); }
    };
    for (uint32_t s0 = 0; s0 + 1 < M; s0 += B) {
        const uint32_t s1 = std::min(s0 + B, M - 1);
        next.clear();
        const bool records = s1 + 1 > P.first_tally_step;
        const bool lattice_recorded = g_lattice_recorded > 0 || (g_lattice_recorded < 0 && S >= 256u && img.lattice_cells_per_step >= 1.25);
        const bool lattice_pass = have_lattice && (!records || lattice_recorded) && (pool_in_lattice || pool.empty());
        if (pool_in_lattice && !lattice_pass) {
            for (auto& p : pool) { psim::coarse_to_fine(PL.cells, PL.sub_fine, p.cell, p.b1, p.b2); }
        }
        pool_in_lattice = lattice_pass;
        P = lattice_pass ? PL : PF;
        for (const auto& p : pool) { run_one(p, s0, P.step_time, s0, s1); }
        for (uint32_t eidx = plan.step_begin[s0]; eidx < plan.step_begin[s1]; ++eidx) {
            const DevBirth& b = plan.births[eidx];
            for (uint32_t i = 0; i < b.count; ++i) {
                psim::Phonon p;
                const float t_first = psim::create_phonon(P, plan.sources[b.source], b.j0 + static_cast<uint64_t>(i) * b.stride, b.step, p);
                run_one(p, b.step, t_first, s0, s1);
            }
        }
        pool.swap(next);
        if (alive_per_pass) { alive_per_pass[s1 - 1] = pool.size(); }
        if (cell_hist_steps) {
            for (auto p : pool) {
                if (pool_in_lattice) { psim::coarse_to_fine(PL.cells, PL.sub_fine, p.cell, p.b1, p.b2); }
                ++cell_hist_steps[static_cast<size_t>(s1 - 1) * P.n_cells + psim::api_cell_of(PF, p.cell, p.b1, p.b2)];
            }
        }
    }
    const double scale = 1. / static_cast<double>(1 << PSIM_FLUX_FRAC_BITS);
    for (uint32_t r = 0; r < R; ++r) {
        for (uint32_t s = 0; s < S; ++s) {
            const size_t src = static_cast<size_t>(r) * S + s, dst = static_cast<size_t>(s) * R + r;
            if (energy) { energy[dst] = static_cast<int32_t>(te[src]); }
            if (flux) {
                flux[2 * dst] = static_cast<double>(tf[2 * src]) * scale;
                flux[2 * dst + 1] = static_cast<double>(tf[2 * src + 1]) * scale;
            }
            if (flux_fixed) {
                flux_fixed[2 * dst] = tf[2 * src];
                flux_fixed[2 * dst + 1] = tf[2 * src + 1];
            }
        }
    }
    if (drift_steps) { *drift_steps = n_steps; }
    if (events) { *events = total_events; }
    return 0;
}

// flatten_model's view of the mesh: flight cells (with and without merging triangle pairs) and distinct shapes
extern "C" int psim_emu_mesh_info(const psim_model_desc* desc, int merge_cells, uint32_t* flight_cells, uint32_t* shapes, uint32_t* classes) {
    psim::HostImage img;
    std::string e;
    const int rc = psim::flatten_model(*desc, img, e, merge_cells);
    if (rc) { return rc; }
    if (flight_cells) { *flight_cells = static_cast<uint32_t>(img.cells.size()); }
    if (shapes) { *shapes = static_cast<uint32_t>(img.shapes.size()); }
    if (classes) { *classes = static_cast<uint32_t>(img.classes.size()); }
    return 0;
}

// the lattice image: cells, how many of them stand for more than one fine flight cell, the largest block, partial-edge records
extern "C" int psim_emu_lattice_info(const psim_model_desc* desc, uint32_t* cells, uint32_t* merged, uint32_t* largest, uint32_t* subs) {
    psim::HostImage img;
    std::string e;
    const int rc = psim::flatten_model(*desc, img, e, 2);
    if (rc) { return rc; }
    uint32_t m = 0, big = 0;
    for (const DevCell& c : img.lattice_cells) {
        const uint32_t n = (c.tri[1] & 0xFFFFu) * (c.tri[1] >> 16);
        m += n > 1u;
        big = std::max(big, n);
    }
    if (cells) { *cells = static_cast<uint32_t>(img.lattice_cells.size()); }
    if (merged) { *merged = m; }
    if (largest) { *largest = big; }
    if (subs) { *subs = static_cast<uint32_t>(img.lattice_subs.size()); }
    return 0;
}
