// C ABI of the host layer (include/psim_host.h) over psim::Model.
#include "../../../include/psim_host.h"
#include "model.h"

#include <cuda_runtime_api.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

struct psim_model {
    std::unique_ptr<psim::Model> m;
};

namespace {

thread_local std::string g_error;

template<typename F> int guarded(int model_error_code, F&& f) {
    try {
        return f();
    } catch (const std::exception& e) {
        g_error = e.what();
        return model_error_code;
    } catch (...) {
        g_error = "unknown error";
        return model_error_code;
    }
}

// ---- NCCL for the tally sum of a multi-device run (psim_model_run_devices) --------------------------------------------
// Bound at run time (dlopen), only when a run really uses several devices: the library has no load-time dependency on
// NCCL, and inside a process that already carries an NCCL (a torchrun rank with torch's bundled copy) the loader hands
// back that one instead of a second copy.  Types and enumerators are the ABI-stable ones of nccl.h (ncclInt32 = 2,
// ncclInt64 = 4, ncclSum = 0, ncclSuccess = 0).
struct Nccl {
    using Comm = void*;
    int (*CommInitAll)(Comm*, int, const int*) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;

    static const Nccl& get() {
        static const Nccl instance = [] {
            Nccl n;
            void* lib = nullptr;
            for (const char* name : { "libnccl.so.2", "libnccl.so" }) {
                lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
                if (lib) { break; }
            }
            if (!lib) {
                n.why = "libnccl.so.2 not found";
                return n;
            }
            auto sym = [&](const char* name) { return dlsym(lib, name); };
            n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
            n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
            n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
            n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
            n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
            n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
            n.ok = n.CommInitAll && n.CommDestroy && n.AllReduce && n.GroupStart && n.GroupEnd && n.GetErrorString;
            if (!n.ok) { n.why = "libnccl.so.2 lacks an expected symbol"; }
            return n;
        }();
        return instance;
    }
};
constexpr int kNcclInt32 = 2, kNcclInt64 = 4, kNcclSum = 0;

// One communicator and one stream per device of a multi-device run; all_reduce() sums the device-resident tallies of all
// handles in place (int32 energy, int64 fixed-point flux: exact, so every device ends up with the bits the host sum gives).
class TallyExchange {
public:
    TallyExchange(const int* devices, size_t n) : devices_(devices, devices + n), comms_(n, nullptr), streams_(n, nullptr) {
        const Nccl& nccl = Nccl::get();
        if (!nccl.ok) {
            why_ = nccl.why;
            return;
        }
        if (std::getenv("PSIM_HOST_SUM")) {
            why_ = "PSIM_HOST_SUM is set";
            return;
        }
        if (const int e = nccl.CommInitAll(comms_.data(), static_cast<int>(n), devices)) {
            why_ = std::string("ncclCommInitAll: ") + nccl.GetErrorString(e);
            std::fill(comms_.begin(), comms_.end(), nullptr);
            return;
        }
        for (size_t d = 0; d < n; ++d) {
            if (cudaSetDevice(devices[d]) != cudaSuccess || cudaStreamCreateWithFlags(&streams_[d], cudaStreamNonBlocking) != cudaSuccess) {
                why_ = "cudaStreamCreate failed";
                return;
            }
        }
        ready_ = true;
    }
    ~TallyExchange() {
        for (size_t d = 0; d < devices_.size(); ++d) {
            if (streams_[d]) {
                cudaSetDevice(devices_[d]);
                cudaStreamDestroy(streams_[d]);
            }
            if (comms_[d]) { Nccl::get().CommDestroy(comms_[d]); }
        }
    }
    TallyExchange(const TallyExchange&) = delete;
    TallyExchange& operator=(const TallyExchange&) = delete;
    bool ready() const { return ready_; }
    const std::string& why_not() const { return why_; }

    // the handles have finished their runs (psim_gpu_run is blocking); returns an error text or ""
    std::string all_reduce(const std::vector<psim_gpu*>& gpus) {
        const Nccl& nccl = Nccl::get();
        std::vector<void*> e(gpus.size()), f(gpus.size());
        uint32_t R = 0, S = 0;
        for (size_t d = 0; d < gpus.size(); ++d) {
            if (psim_gpu_tally_buffers(gpus[d], &e[d], &f[d], &R, &S)) { return "psim_gpu_tally_buffers failed"; }
        }
        const size_t n = static_cast<size_t>(R) * S;
        int rc = nccl.GroupStart();
        for (size_t d = 0; d < gpus.size() && !rc; ++d) {
            rc = nccl.AllReduce(e[d], e[d], n, kNcclInt32, kNcclSum, comms_[d], streams_[d]);
            if (!rc) { rc = nccl.AllReduce(f[d], f[d], 2 * n, kNcclInt64, kNcclSum, comms_[d], streams_[d]); }
        }
        const int rc_end = nccl.GroupEnd();
        if (rc || rc_end) { return std::string("ncclAllReduce: ") + nccl.GetErrorString(rc ? rc : rc_end); }
        for (size_t d = 0; d < gpus.size(); ++d) {
            if (cudaSetDevice(devices_[d]) != cudaSuccess || cudaStreamSynchronize(streams_[d]) != cudaSuccess) {
                return "synchronising the tally all-reduce failed";
            }
        }
        return "";
    }

private:
    std::vector<int> devices_;
    std::vector<Nccl::Comm> comms_;
    std::vector<cudaStream_t> streams_;
    bool ready_ = false;
    std::string why_;
};

int load_with(psim_model** out, const std::function<std::unique_ptr<psim::Model>()>& make) {
    if (!out) {
        g_error = "null argument";
        return PSIM_E_INVALID;
    }
    *out = nullptr;
    return guarded(PSIM_E_MODEL, [&]() {
        auto m = make();
        *out = new psim_model{ std::move(m) };
        return PSIM_OK;
    });
}

}  // namespace

extern "C" {

const char* psim_host_last_error(void) { return g_error.c_str(); }

int psim_model_load(const char* json_path, psim_model** out) {
    if (!json_path) {
        g_error = "null path";
        return PSIM_E_INVALID;
    }
    return load_with(out, [&]() { return psim::Model::from_file(json_path); });
}

int psim_model_load_text(const char* json_text, psim_model** out) {
    if (!json_text) {
        g_error = "null text";
        return PSIM_E_INVALID;
    }
    return load_with(out, [&]() { return psim::Model::from_json_text(json_text); });
}

void psim_model_free(psim_model* m) { delete m; }

int psim_model_get_info(const psim_model* pm, psim_model_info* out) {
    if (!pm || !out) { return PSIM_E_INVALID; }
    const psim::Model& m = *pm->m;
    std::memset(out, 0, sizeof(*out));
    out->num_runs = m.num_runs;
    out->measurement_steps = m.measurement_steps;
    out->recorded_steps = m.recorded_steps;
    out->num_phonons = m.num_phonons;
    out->step_interval = m.step_interval;
    out->simulation_time = m.simulation_time;
    out->t_eq = m.t_eq;
    out->sim_type = static_cast<uint32_t>(m.sim_type);
    out->phasor_sim = m.phasor_sim ? 1u : 0u;
    out->num_materials = static_cast<uint32_t>(m.materials.size());
    out->num_sensors = static_cast<uint32_t>(m.sensors.size());
    out->num_cells = static_cast<uint32_t>(m.cells.size());
    out->num_emitters = static_cast<uint32_t>(m.emitters.size());
    for (const auto& c : m.cells) {
        for (int k = 0; k < 3; ++k) {
            for (const auto& t : c.transitions[k]) {
                ++out->num_transition_links;
                const double lo = std::min(t.s0, t.s1), hi = std::max(t.s0, t.s1);
                if (lo > 1e-9 || hi < 1. - 1e-9) { ++out->num_partial_links; }
            }
        }
    }
    return PSIM_OK;
}

int psim_model_set_num_phonons(psim_model* pm, uint64_t n) {
    if (!pm || n == 0) { return PSIM_E_INVALID; }
    pm->m->num_phonons = n;
    return PSIM_OK;
}

int psim_model_set_num_runs(psim_model* pm, uint64_t n) {
    if (!pm || n == 0) { return PSIM_E_INVALID; }
    pm->m->num_runs = n;
    return PSIM_OK;
}

int psim_model_set_max_iters(psim_model* pm, uint64_t n) {
    if (!pm || n == 0) { return PSIM_E_INVALID; }
    pm->m->max_iters = n;
    return PSIM_OK;
}

int psim_model_prepare(psim_model* pm) {
    if (!pm) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_MODEL, [&]() {
        pm->m->prepare();
        return PSIM_OK;
    });
}

int psim_model_energy(psim_model* pm, double* total, double* per_phonon) {
    if (!pm) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_STATE, [&]() {
        const double t = pm->m->total_initial_energy();
        if (total) { *total = t; }
        if (per_phonon) { *per_phonon = t / static_cast<double>(pm->m->num_phonons); }
        return PSIM_OK;
    });
}

int psim_model_material_arrays(psim_model* pm, uint32_t material, double* freq, double* vel_la, double* vel_ta,
                               double* dens_la, double* dens_ta) {
    if (!pm || material >= pm->m->materials.size()) { return PSIM_E_INVALID; }
    const psim::Material& mt = pm->m->materials[material];
    const size_t bytes = sizeof(double) * psim::kBins;
    if (freq) { std::memcpy(freq, mt.freq.data(), bytes); }
    if (vel_la) { std::memcpy(vel_la, mt.vel_la.data(), bytes); }
    if (vel_ta) { std::memcpy(vel_ta, mt.vel_ta.data(), bytes); }
    if (dens_la) { std::memcpy(dens_la, mt.dens_la.data(), bytes); }
    if (dens_ta) { std::memcpy(dens_ta, mt.dens_ta.data(), bytes); }
    return PSIM_OK;
}

int psim_model_table(psim_model* pm, uint32_t material, uint32_t kind, double temperature, double* cumulative,
                     double* la_fraction, double* sum) {
    if (!pm || material >= pm->m->materials.size() || kind > 2) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_STATE, [&]() {
        const psim::Table& t = pm->m->materials[material].table(static_cast<psim::Material::Kind>(kind), temperature);
        if (cumulative) { std::memcpy(cumulative, t.cumulative.data(), sizeof(double) * psim::kBins); }
        if (la_fraction) { std::memcpy(la_fraction, t.la_fraction.data(), sizeof(double) * psim::kBins); }
        if (sum) { *sum = t.sum; }
        return PSIM_OK;
    });
}

int psim_model_cell_energies(psim_model* pm, double* area, double* init_energy, double* emit_energy) {
    if (!pm) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_STATE, [&]() {
        psim::Model& m = *pm->m;
        for (size_t i = 0; i < m.cells.size(); ++i) {
            const auto& c = m.cells[i];
            const auto& s = m.sensors[c.sensor];
            const double t_init = (m.sim_type == psim::SimType::SteadyState) ? s.t_steady : s.t_init;
            if (area) { area[i] = c.area; }
            if (init_energy) {
                const double e = c.area * s.heat_capacity;
                init_energy[i] = (m.t_eq == 0.) ? e : e * std::fabs(t_init - m.t_eq);
            }
            if (emit_energy) {
                double sum = 0.;
                for (int k = 0; k < 3; ++k) {
                    for (const auto& sub : c.emits[k]) {
                        const auto& e = m.emitters[sub.target];
                        const double en = e.length * e.duration * m.materials[s.material].emit_energy(e.temp) / 4.;
                        sum += (m.t_eq == 0.) ? en : en * std::fabs(e.temp - m.t_eq);
                    }
                }
                emit_energy[i] = sum;
            }
        }
        return PSIM_OK;
    });
}

int psim_model_sensor_ids(const psim_model* pm, uint64_t* ids, double* areas) {
    if (!pm) { return PSIM_E_INVALID; }
    for (size_t i = 0; i < pm->m->sensors.size(); ++i) {
        if (ids) { ids[i] = pm->m->sensors[i].id; }
        if (areas) { areas[i] = pm->m->sensors[i].area; }
    }
    return PSIM_OK;
}

int psim_model_describe(psim_model* pm, const psim_model_desc** out) {
    if (!pm || !out) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_STATE, [&]() {
        *out = &pm->m->describe();
        return PSIM_OK;
    });
}

int psim_model_sources(psim_model* pm, uint64_t seed, psim_source* sources, size_t* n) {
    if (!pm || !n) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_STATE, [&]() {
        const auto v = pm->m->source_counts(seed);
        if (sources) {
            if (*n < v.size()) {
                g_error = "source buffer too small";
                return static_cast<int>(PSIM_E_INVALID);
            }
            std::memcpy(sources, v.data(), v.size() * sizeof(psim_source));
        }
        *n = v.size();
        return static_cast<int>(PSIM_OK);
    });
}

int psim_model_set_tallies(psim_model* pm, const int32_t* energy, const double* flux) {
    if (!pm || !energy || !flux) { return PSIM_E_INVALID; }
    pm->m->set_tallies(energy, flux);
    return PSIM_OK;
}

int psim_model_end_iteration(psim_model* pm, int* again) {
    if (!pm) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_STATE, [&]() {
        const bool more = pm->m->end_iteration(nullptr);
        if (again) { *again = more ? 1 : 0; }
        return PSIM_OK;
    });
}

int psim_model_finish_run(psim_model* pm, uint64_t run_id, int* stable) {
    if (!pm) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_STATE, [&]() {
        const int s = pm->m->finish_run(run_id, nullptr);
        if (stable) { *stable = s; }
        return PSIM_OK;
    });
}

int psim_model_next_run(psim_model* pm) {
    if (!pm) { return PSIM_E_INVALID; }
    pm->m->reset_for_next_run();
    return PSIM_OK;
}

int psim_model_run(psim_model* pm, int device, uint64_t seed, int steps_per_launch, int verbose, psim_stats* stats) {
    return psim_model_run_devices(pm, &device, 1, seed, steps_per_launch, verbose, stats);
}

int psim_model_run_devices(psim_model* pm, const int* devices, int n_devices, uint64_t seed, int steps_per_launch, int verbose,
                           psim_stats* stats) {
    if (!pm || !devices || n_devices < 1) { return PSIM_E_INVALID; }
    psim::Model& m = *pm->m;
    const size_t G = static_cast<size_t>(n_devices);
    std::vector<psim_gpu*> gpus(G, nullptr);
    const auto t_call = std::chrono::steady_clock::now();
    const int rc = guarded(PSIM_E_STATE, [&]() -> int {
        m.runs.clear();
        // every call starts from the state of the model file: a steady-state run leaves its final temperatures in the
        // sensors (Model::resetRequired, model.cpp:250-272), which a second call on the same handle must not inherit
        m.restore_file_state();
        if (std::getenv("PSIM_TIMING")) {
            std::cerr << "psim timing [ms]: results of the call before dropped "
                      << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count() << '\n';
        }
        // several devices: their tallies are summed with one NCCL all-reduce per run over NVLink (integers: the result is the
        // one-device result bit for bit); without a usable NCCL the same integers are summed on the host
        std::unique_ptr<TallyExchange> exchange;
        if (G > 1) {
            exchange = std::make_unique<TallyExchange>(devices, G);
            if (!exchange->ready() && verbose) { std::cerr << "psim: tallies are summed on the host (" << exchange->why_not() << ")\n"; }
        }
        const bool use_nccl = exchange && exchange->ready();
        for (uint64_t run = 0; run < m.num_runs; ++run) {
            if (verbose) { std::cout << "Run: " << run + 1 << '\n'; }
            const auto h0 = std::chrono::steady_clock::now();
            m.prepare();
            std::string log;
            // the iterations of one run (model.cpp:159-172; one, unless max_iters was raised): each starts from the sensor
            // temperatures, tables and t_eq the iteration before left behind
            for (uint64_t iter = 0;; ++iter) {
            if (iter > 0) {  // the device image holds the temperatures and tables of the iteration before: rebuild it
                for (auto& g : gpus) {
                    psim_gpu_destroy(g);
                    g = nullptr;
                }
            }
            const psim_model_desc& desc = m.describe();
            const uint64_t run_seed = seed + run + 7919u * iter;
            const auto sources = m.source_counts(run_seed);
            if (std::getenv("PSIM_TIMING")) {
                std::cerr << "psim timing [ms]: prepare+describe+sources "
                          << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count() << '\n';
            }
            const size_t n_tally = m.sensors.size() * m.recorded_steps;
            // One device, or several whose tallies NCCL sums on device 0: energies and fluxes come back from the device as
            // the host layer wants them (psim_gpu_get_tallies transposes and converts there).  Only the host-side sum of
            // several devices needs the exact fixed-point integers of each.
            const bool host_sum = G > 1 && !use_nccl;
            // They are written straight into the model's own tally storage, which lives from run to run.
            const auto storage = m.tally_storage();
            int32_t* const energy_out = storage.first;
            double* const flux_out = storage.second;
            std::vector<std::vector<int32_t>> energy(host_sum ? G : 0, std::vector<int32_t>(host_sum ? n_tally : 0));
            std::vector<std::vector<int64_t>> fixed(host_sum ? G : 0, std::vector<int64_t>(host_sum ? 2 * n_tally : 0));
            std::vector<psim_stats> st(G);
            std::vector<int> codes(G, PSIM_OK);
            std::vector<std::string> errors(G);
            auto work = [&](size_t d) {
                int e = PSIM_OK;
                const bool timing = std::getenv("PSIM_TIMING") != nullptr;  // phase times of device 0's thread on stderr
                auto now = [] { return std::chrono::steady_clock::now(); };
                auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
                const auto t0 = now();
                if (!gpus[d]) {
                    e = psim_gpu_create(&desc, devices[d], &gpus[d]);
                    if (e) {
                        errors[d] = psim_gpu_last_error(nullptr);
                    } else if (steps_per_launch > 0) {
                        e = psim_gpu_set_option(gpus[d], "steps_per_launch", steps_per_launch);
                    }
                }
                const auto t1 = now();
                if (!e) { e = psim_gpu_set_sources(gpus[d], sources.data(), sources.size(), run_seed, static_cast<uint32_t>(d), static_cast<uint32_t>(G)); }
                const auto t2 = now();
                if (!e) { e = psim_gpu_run(gpus[d]); }
                const auto t3 = now();
                if (!e && host_sum) { e = psim_gpu_get_tallies(gpus[d], energy[d].data(), nullptr, fixed[d].data()); }
                if (!e && G == 1) { e = psim_gpu_get_tallies(gpus[0], energy_out, flux_out, nullptr); }
                if (!e) { e = psim_gpu_get_stats(gpus[d], &st[d]); }
                if (timing && d == 0) {
                    std::cerr << "psim timing [ms]: create " << ms(t0, t1) << " set_sources " << ms(t1, t2) << " run " << ms(t2, t3)
                              << " tallies " << ms(t3, now()) << '\n';
                }
                if (e && errors[d].empty() && gpus[d]) { errors[d] = psim_gpu_last_error(gpus[d]); }
                codes[d] = e;
            };
            if (G == 1) {
                work(0);
            } else {
                std::vector<std::thread> threads;
                for (size_t d = 0; d < G; ++d) { threads.emplace_back(work, d); }
                for (auto& t : threads) { t.join(); }
            }
            for (size_t d = 0; d < G; ++d) {
                if (codes[d]) {
                    g_error = errors[d];
                    return codes[d];
                }
            }
            if (use_nccl) {
                const auto t4 = std::chrono::steady_clock::now();
                if (const std::string err = exchange->all_reduce(gpus); !err.empty()) {
                    g_error = err;
                    return PSIM_E_CUDA;
                }
                if (const int e = psim_gpu_get_tallies(gpus[0], energy_out, flux_out, nullptr)) {  // device 0 holds the sum
                    g_error = psim_gpu_last_error(gpus[0]);
                    return e;
                }
                if (std::getenv("PSIM_TIMING")) {
                    std::cerr << "psim timing [ms]: nccl all-reduce + tallies "
                              << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t4).count() << '\n';
                }
            }
            for (size_t i = 0; host_sum && i < n_tally; ++i) {  // integer sums: independent of the number of devices
                int64_t e = 0, fx = 0, fy = 0;
                for (size_t d = 0; d < G; ++d) {
                    e += energy[d][i];
                    fx += fixed[d][2 * i];
                    fy += fixed[d][2 * i + 1];
                }
                energy_out[i] = static_cast<int32_t>(e);
                flux_out[2 * i] = static_cast<double>(fx) / 256.;
                flux_out[2 * i + 1] = static_cast<double>(fy) / 256.;
            }
            if (stats) {
                *stats = st[0];
                for (size_t d = 1; d < G; ++d) {
                    stats->shard_phonons += st[d].shard_phonons;
                    stats->drift_steps += st[d].drift_steps;
                    stats->events += st[d].events;
                    stats->peak_alive += st[d].peak_alive;
                    stats->kernel_ms = std::max(stats->kernel_ms, st[d].kernel_ms);
                }
            }
            m.tallies_written();
            const auto h_end = std::chrono::steady_clock::now();
            const bool again = m.end_iteration(&log);
            if (std::getenv("PSIM_TIMING")) {
                std::cerr << "psim timing [ms]: end of iteration "
                          << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h_end).count() << '\n';
            }
            if (!again) { break; }
            }  // iterations
            const auto h1 = std::chrono::steady_clock::now();
            m.finish_run(run, &log);
            if (std::getenv("PSIM_TIMING")) {
                std::cerr << "psim timing [ms]: epilogue "
                          << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h1).count() << '\n';
            }
            if (verbose) { std::cout << log; }
            if (run + 1 < m.num_runs) { m.reset_for_next_run(); }
        }
        return PSIM_OK;
    });
    const auto t_destroy = std::chrono::steady_clock::now();
    for (psim_gpu* g : gpus) { psim_gpu_destroy(g); }
    if (std::getenv("PSIM_TIMING")) {
        std::cerr << "psim timing [ms]: destroy "
                  << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_destroy).count() << "  whole call "
                  << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count() << '\n';
    }
    return rc;
}

int psim_model_results(const psim_model* pm, uint64_t run_id, double* six, double* temps, double* fluxes) {
    if (!pm) { return PSIM_E_INVALID; }
    const psim::Model& m = *pm->m;
    std::vector<psim::SensorResult> avg;
    const std::vector<psim::SensorResult>* res = nullptr;
    if (run_id == std::numeric_limits<uint64_t>::max()) {
        avg = m.averaged();
        res = &avg;
    } else if (run_id < m.runs.size()) {
        res = &m.runs[run_id];
    }
    if (!res || res->empty()) {
        g_error = "no results for this run";
        return PSIM_E_STATE;
    }
    const size_t R = m.recorded_steps;
    for (size_t i = 0; i < res->size(); ++i) {
        const auto& r = (*res)[i];
        if (six) {
            const double row[6] = { r.t_steady, r.std_t_steady, r.x_flux, r.std_x_flux, r.y_flux, r.std_y_flux };
            std::memcpy(six + 6 * i, row, sizeof(row));
        }
        for (size_t k = 0; k < R; ++k) {
            if (temps) { temps[i * R + k] = r.final_temps[k]; }
            if (fluxes) {
                fluxes[2 * (i * R + k)] = r.final_fluxes[k][0];
                fluxes[2 * (i * R + k) + 1] = r.final_fluxes[k][1];
            }
        }
    }
    return PSIM_OK;
}

double psim_model_energy_per_phonon(const psim_model* pm) { return pm ? pm->m->energy_per_phonon() : 0.; }

int psim_model_export(const psim_model* pm, const char* model_path, double seconds) {
    if (!pm || !model_path) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_IO, [&]() {
        pm->m->export_results(model_path, seconds);
        return PSIM_OK;
    });
}

int psim_model_export_text(const psim_model* pm, const char* model_filename, double seconds, const char* when, char* buf,
                           size_t* len) {
    if (!pm || !model_filename || !len) { return PSIM_E_INVALID; }
    return guarded(PSIM_E_IO, [&]() {
        const std::string s = pm->m->export_text(model_filename, seconds, when ? when : "");
        if (buf && *len > s.size()) { std::memcpy(buf, s.c_str(), s.size() + 1); }
        *len = s.size() + 1;
        return PSIM_OK;
    });
}

}  // extern "C"
