#include "model.h"
#include "json.h"

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <system_error>
#include <thread>
#include <unordered_map>

namespace psim {
namespace {

constexpr double HBAR = 1.054517e-34;    // reference material.cpp:12
constexpr double BOLTZ = 1.38065e-23;    // reference material.cpp:13
constexpr double PI = 3.1415926535897932384626433832795028841971693993751058209749445923;
constexpr double GEOEPS = 2.220446049250313e-16 * 1e9;  // reference utils.h:10
constexpr double SS_STEPS_PERCENT = 0.1;  // model.cpp:22
constexpr float TEMP_INTERVAL = 0.1F;     // model.cpp:25
constexpr double TEMP_BOUND_EPS = 10.;    // model.cpp:19
constexpr double SENSOR_RESET_THRESHOLD = 0.001;    // sensorController.cpp:9
constexpr double TRANSIENT_RESET_THRESHOLD = 0.02;  // sensorController.cpp:12
constexpr size_t SYSTEM_RESET_THRESHOLD = 90;       // model.cpp:13: percentage of sensors that must be stable
constexpr double TEQ_THRESHOLD = 5.;                // model.cpp:16: |delta t_eq| / t_eq in parts per thousand
constexpr double INVERSION_EPS = 0.0001;            // sensorInterpreter.cpp:9
constexpr std::size_t INVERSION_MAX_ITERS = 40;     // sensorInterpreter.cpp:10

double get_k(double freq, const double c[3]) {  // Material::getK, material.cpp:86-91
    const double d = c[1] * c[1] - 4. * c[0] * (c[2] - freq);
    const double a = (-c[1] - std::sqrt(d)) / (2. * c[0]);
    const double b = (-c[1] + std::sqrt(d)) / (2. * c[0]);
    return (a < b) ? a : b;
}

Table build_cumulative(const std::array<double, kBins>& t1, const std::array<double, kBins>& t2) {
    // Material::buildCumulDist, material.cpp:170-180
    Table out;
    const double sum = std::accumulate(t1.begin(), t1.end(), 0.) + std::accumulate(t2.begin(), t2.end(), 0.);
    out.cumulative[0] = (t1[0] + t2[0]) / sum;
    out.la_fraction[0] = t1[0] / (t1[0] + t2[0]);
    for (int i = 1; i < kBins; ++i) {
        out.cumulative[i] = out.cumulative[i - 1] + (t1[i] + t2[i]) / sum;
        out.la_fraction[i] = t1[i] / (t1[i] + t2[i]);
    }
    out.sum = sum;
    return out;
}

struct Seg {
    double ax, ay, bx, by;
    double length() const { return std::sqrt((bx - ax) * (bx - ax) + (by - ay) * (by - ay)); }
};

bool point_on_line(const Seg& l, double px, double py) {  // isPointOnLine, geometry.cpp:295-298
    return std::fabs((l.bx - l.ax) * (py - l.ay) - (px - l.ax) * (l.by - l.ay)) < GEOEPS;
}

bool seg_contains(const Seg& a, const Seg& b) {  // Line::contains(Line), geometry.cpp:75-86
    if (!(point_on_line(a, b.ax, b.ay) && point_on_line(a, b.bx, b.by) && a.length() >= b.length())) { return false; }
    const double ax0 = std::min(a.ax, a.bx), ax1 = std::max(a.ax, a.bx), ay0 = std::min(a.ay, a.by), ay1 = std::max(a.ay, a.by);
    const double bx0 = std::min(b.ax, b.bx), bx1 = std::max(b.ax, b.bx), by0 = std::min(b.ay, b.by), by1 = std::max(b.ay, b.by);
    return ax1 >= bx1 - GEOEPS && ax0 <= bx0 + GEOEPS && ay1 >= by1 - GEOEPS && ay0 <= by0 + GEOEPS;
}

double edge_param(const Seg& e, double px, double py) {
    const double ex = e.bx - e.ax, ey = e.by - e.ay;
    return ((px - e.ax) * ex + (py - e.ay) * ey) / (ex * ex + ey * ey);
}

Seg cell_edge(const CellRec& c, int k) { return Seg{ c.x[k], c.y[k], c.x[(k + 1) % 3], c.y[(k + 1) % 3] }; }

double triangle_area(const CellRec& c) {  // Triangle::area (Heron), geometry.cpp:250-258
    const double a = cell_edge(c, 0).length(), b = cell_edge(c, 1).length(), d = cell_edge(c, 2).length();
    const double p = (a + b + d) / 2.;
    return std::sqrt(p * (p - a) * (p - b) * (p - d));
}

bool ranges_overlap(double a0, double a1, double b0, double b1) {
    const double lo = std::max(std::min(a0, a1), std::min(b0, b1)), hi = std::min(std::max(a0, a1), std::max(b0, b1));
    return hi - lo > 1e-9;
}

// uniform spatial hash over edge bounding boxes: replaces the reference's all-pairs loop (model.cpp:103-112)
class EdgeGrid {
public:
    EdgeGrid(double h, double x0, double y0) : h_(h), x0_(x0), y0_(y0) {}
    template<typename F> void visit(const Seg& s, F&& f) const {
        int ix0, iy0, ix1, iy1;
        range(s, ix0, iy0, ix1, iy1);
        for (int ix = ix0; ix <= ix1; ++ix) {
            for (int iy = iy0; iy <= iy1; ++iy) {
                auto it = map_.find(key(ix, iy));
                if (it == map_.end()) { continue; }
                for (uint32_t id : it->second) { f(id); }
            }
        }
    }
    void insert(const Seg& s, uint32_t id) {
        int ix0, iy0, ix1, iy1;
        range(s, ix0, iy0, ix1, iy1);
        for (int ix = ix0; ix <= ix1; ++ix) {
            for (int iy = iy0; iy <= iy1; ++iy) { map_[key(ix, iy)].push_back(id); }
        }
    }

private:
    double h_, x0_, y0_;
    std::unordered_map<uint64_t, std::vector<uint32_t>> map_;
    static uint64_t key(int ix, int iy) { return (static_cast<uint64_t>(static_cast<uint32_t>(ix)) << 32) | static_cast<uint32_t>(iy); }
    void range(const Seg& s, int& ix0, int& iy0, int& ix1, int& iy1) const {
        const double pad = 1e-6;
        ix0 = static_cast<int>(std::floor((std::min(s.ax, s.bx) - pad - x0_) / h_));
        ix1 = static_cast<int>(std::floor((std::max(s.ax, s.bx) + pad - x0_) / h_));
        iy0 = static_cast<int>(std::floor((std::min(s.ay, s.by) - pad - y0_) / h_));
        iy1 = static_cast<int>(std::floor((std::max(s.ay, s.by) + pad - y0_) / h_));
    }
};

uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// threads for a piece of host work that divides evenly: as many as there are units of work (a unit is what is not worth a
// thread of its own), at most 16 and at most PSIM_HOST_THREADS
size_t host_threads(size_t units) {
    size_t threads = std::min<size_t>({ std::max(1u, std::thread::hardware_concurrency()), 16, units });
    if (const char* cap = std::getenv("PSIM_HOST_THREADS")) { threads = std::min<size_t>(threads, std::strtoul(cap, nullptr, 10)); }
    return threads;
}

// part(t) for t = 0 ... parts - 1, each on a thread of its own (the calling thread takes the last part, and every part a
// thread could not be started for); the first exception of any part is rethrown once all of them have ended
template <class F> void run_parts(size_t parts, F&& part) {
    std::vector<std::thread> pool;
    std::vector<std::exception_ptr> errors(parts);
    auto guarded = [&](size_t t) {
        try {
            part(t);
        } catch (...) { errors[t] = std::current_exception(); }
    };
    size_t started = 0;
    try {
        pool.reserve(parts);
        for (; started + 1 < parts; ++started) { pool.emplace_back(guarded, started); }
    } catch (const std::system_error&) {}  // no more threads: the rest runs here
    for (size_t t = started; t < parts; ++t) { guarded(t); }
    for (auto& t : pool) { t.join(); }
    for (const auto& e : errors) {
        if (e) { std::rethrow_exception(e); }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------- Material
void Material::build_dispersion() {
    freq_width = std::max(w_max_la, w_max_ta) / kBins;
    for (int i = 0; i < kBins; ++i) { freq[i] = (2 * i + 1) * freq_width / 2.; }
    vel_la.fill(0.);
    vel_ta.fill(0.);
    dens_la.fill(0.);
    dens_ta.fill(0.);
    for (int i = 0; i < kBins; ++i) {
        const double k_la = get_k(freq[i], la);
        const double gv_la = 2. * la[0] * k_la + la[1];
        vel_la[i] = gv_la;
        dens_la[i] = k_la * k_la / 2. / (PI * PI) / gv_la;
        const double k_ta = get_k(freq[i], ta);
        const double gv_ta = 2. * ta[0] * k_ta + ta[1];
        if (!std::isnan(gv_ta)) {  // above the TA branch's top the root is complex: no TA states (material.cpp:43-48)
            vel_ta[i] = gv_ta;
            dens_ta[i] = k_ta * k_ta / (PI * PI) / gv_ta;  // doubly degenerate branch
        }
    }
}

void Material::set_temperature_grid(double low, double high) {
    const auto steps = static_cast<std::size_t>((high - low) / static_cast<double>(TEMP_INTERVAL));
    temps_.assign(steps, 0.);
    for (std::size_t n = 0; n < steps; ++n) { temps_[n] = low + static_cast<double>(TEMP_INTERVAL) * static_cast<double>(n); }
    temps_.push_back(high);
    cache_.clear();
}

size_t Material::temp_index(double temp) const {
    const auto idx = static_cast<size_t>(std::lower_bound(temps_.begin(), temps_.end(), temp) - temps_.begin());
    return std::min(idx, temps_.size() - 1);
}

std::array<double, 3> Material::relax_rates(double temp, double omega, bool is_ta) const {
    double n = 0., u = 0.;
    if (!is_ta) {
        n = u = b_l * omega * omega * std::pow(temp, 3);
    } else if (omega < w) {
        n = b_tn * omega * std::pow(temp, 4);
    } else {
        u = b_tu * omega * omega / std::sinh(HBAR * omega / (temp * BOLTZ));
    }
    return { n, u, b_i * std::pow(omega, 4) };
}

std::array<double, kBins> Material::phonon_dist(double temp, bool is_ta) const {
    std::array<double, kBins> out{};
    const auto& dens = is_ta ? dens_ta : dens_la;
    const double c = HBAR / (BOLTZ * temp);
    for (int i = 0; i < kBins; ++i) {
        const double f = freq[i];
        double d = f * HBAR / std::expm1(c * f) * freq_width * dens[i];
        if (!full_simulation) { d *= c * f * std::exp(c * f) / (std::expm1(c * f) * temp); }  // d/dT of Bose-Einstein
        out[i] = d;
    }
    return out;
}

const Table& Material::table(Kind kind, double temp) {
    if (temps_.empty()) { throw std::runtime_error("material tables requested before the temperature grid was set"); }
    const size_t idx = temp_index(temp);
    auto key = std::make_pair(static_cast<int>(kind), idx);
    auto it = cache_.find(key);
    if (it != cache_.end()) { return *it->second; }
    const double t = temps_[idx];
    auto la_d = phonon_dist(t, false), ta_d = phonon_dist(t, true);
    if (kind == Emit) {  // cumulDistEmit, material.cpp:140-150
        for (int i = 0; i < kBins; ++i) {
            la_d[i] *= vel_la[i];
            ta_d[i] *= vel_ta[i];
        }
    } else if (kind == Scatter) {  // cumulDistScatter, material.cpp:152-168
        for (int i = 0; i < kBins; ++i) {
            const auto rl = relax_rates(t, freq[i], false), rt = relax_rates(t, freq[i], true);
            la_d[i] *= rl[0] + rl[1] + rl[2];
            ta_d[i] *= rt[0] + rt[1] + rt[2];
        }
    }
    auto res = cache_.emplace(key, std::make_unique<Table>(build_cumulative(la_d, ta_d)));
    return *res.first->second;
}

// ------------------------------------------------------------------------------------------------------- Model
std::unique_ptr<Model> Model::from_file(const std::string& path) {
    std::ifstream f(path);
    if (!f.is_open()) { throw std::runtime_error("Error opening file at \"" + path + "\""); }
    std::stringstream ss;
    ss << f.rdbuf();
    return from_json_text(ss.str());
}

std::unique_ptr<Model> Model::from_json_text(const std::string& text) {
    const Json j = Json::parse(text);
    auto m = std::make_unique<Model>();
    const Json& st = j.at("settings");
    const Json& ph = st.at("phasor_sim");
    m->phasor_sim = (ph.kind == Json::Kind::Bool && ph.boolean);  // reference: dump() == "true" (inputManager.cpp:18)
    // a JSON double becomes a count only if it is one: non-negative, integral after truncation as the reference's
    // get<std::size_t>() would take it, and exactly representable (< 2^53) - the cast of anything else is undefined
    auto count = [](const Json& v, const char* what) -> uint64_t {
        const double d = v.num();
        if (!(d >= 0.) || !(d < 9007199254740992.)) { throw std::runtime_error(std::string("Invalid settings: ") + what + " is not a valid count.\n"); }
        return static_cast<uint64_t>(d);
    };
    m->measurement_steps = count(st.at("num_measurements"), "num_measurements");
    m->num_phonons = count(st.at("num_phonons"), "num_phonons");
    m->simulation_time = st.at("sim_time").num();
    m->t_eq = st.at("t_eq").num();
    if (st.contains("num_runs")) { m->num_runs = count(st.at("num_runs"), "num_runs"); }
    if (st.contains("max_iters")) { m->max_iters = std::max<uint64_t>(1, count(st.at("max_iters"), "max_iters")); }  // extension, see model.h
    m->t_eq_file_ = m->t_eq;
    switch (count(st.at("sim_type"), "sim_type")) {
    case 1: m->sim_type = SimType::Periodic; break;
    case 2: m->sim_type = SimType::Transient; break;
    default: m->sim_type = SimType::SteadyState; break;
    }
    if (m->measurement_steps < 2 || m->num_phonons == 0 || !(m->simulation_time > 0.) || m->num_runs == 0) {
        throw std::runtime_error("Invalid settings: num_measurements, num_phonons, sim_time and num_runs must be positive.\n");
    }
    // Model::setSimulationType, model.cpp:46-70
    const double M = static_cast<double>(m->measurement_steps);
    if (m->sim_type != SimType::SteadyState) {
        m->step_interval = count(st.at("step_interval"), "step_interval");
        m->start_step = static_cast<uint64_t>(M - M * SS_STEPS_PERCENT);
        if (m->step_interval == 0) { throw std::runtime_error("Step interval of 0 is invalid for transient and periodic simulations.\n"); }
    }
    if (m->sim_type == SimType::Transient && m->t_eq == 0.) {
        throw std::runtime_error("Transient simulations must be run using the deviational approach.\n");
    }
    if (m->sim_type == SimType::SteadyState) {
        m->step_adjustment = static_cast<uint64_t>(M - M * SS_STEPS_PERCENT);
        m->recorded_steps = static_cast<uint64_t>(M * SS_STEPS_PERCENT);  // model.cpp:89-91
    } else {
        m->recorded_steps = m->measurement_steps;
    }
    if (m->recorded_steps == 0) { throw std::runtime_error("Too few measurement steps to record anything.\n"); }

    std::map<std::string, uint32_t> material_ids;
    for (const Json& md : j.at("materials").items) {
        Material mat;
        mat.name = md.at("name").str();
        if (material_ids.count(mat.name)) { throw std::runtime_error("A duplicate material name was detected.\n"); }
        const Json& dd = md.at("d_data");
        const Json& rd = md.at("r_data");
        for (int i = 0; i < 3; ++i) {
            mat.la[i] = dd.at("la_data").at(i).num();
            mat.ta[i] = dd.at("ta_data").at(i).num();
        }
        mat.w_max_la = dd.at("max_freq_la").num();
        mat.w_max_ta = dd.at("max_freq_ta").num();
        mat.b_l = rd.at("b_l").num();
        mat.b_tn = rd.at("b_tn").num();
        mat.b_tu = rd.at("b_tu").num();
        mat.b_i = rd.at("b_i").num();
        mat.w = rd.at("w").num();
        mat.id = static_cast<uint32_t>(m->materials.size());
        mat.full_simulation = (m->t_eq == 0.);
        mat.build_dispersion();
        material_ids[mat.name] = mat.id;
        m->materials.push_back(std::move(mat));
    }
    std::map<uint64_t, uint32_t> sensor_index;
    for (const Json& sd : j.at("sensors").items) {
        SensorRec s{};
        s.id = count(sd.at("id"), "sensor id");
        if (sensor_index.count(s.id)) { throw std::runtime_error("Sensor with this ID already exists\n"); }
        const auto it = material_ids.find(sd.at("material").str());
        if (it == material_ids.end()) { throw std::runtime_error("Sensor refers to a material that does not exist\n"); }
        s.material = it->second;
        s.t_init = sd.at("t_init").num();
        s.t_steady = s.t_init;
        sensor_index[s.id] = static_cast<uint32_t>(m->sensors.size());
        m->sensors.push_back(s);
    }
    auto point = [](const Json& p, double& x, double& y) {
        x = p.at("x").num();
        y = p.at("y").num();
    };
    for (const Json& cd : j.at("cells").items) {
        CellRec c{};
        const Json& t = cd.at("triangle");
        point(t.at("p1"), c.x[0], c.y[0]);
        point(t.at("p2"), c.x[1], c.y[1]);
        point(t.at("p3"), c.x[2], c.y[2]);
        const auto it = sensor_index.find(count(cd.at("sensorID"), "sensorID"));
        if (it == sensor_index.end()) { throw std::runtime_error("Sensor does not exist\n"); }
        c.sensor = it->second;
        c.spec = cd.at("specularity").num();
        m->cells.push_back(c);
    }
    if (m->cells.empty() || m->sensors.empty()) { throw std::runtime_error("A model needs at least one sensor and one cell.\n"); }
    m->build_geometry();
    for (const Json& sd : j.at("emit_surfaces").items) {
        double p1x, p1y, p2x, p2y;
        point(sd.at("p1"), p1x, p1y);
        point(sd.at("p2"), p2x, p2y);
        m->attach_emit_surface(p1x, p1y, p2x, p2y, sd.at("temp").num(), sd.at("duration").num(), sd.at("start_time").num());
    }
    return m;
}

void Model::build_geometry() {
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
    std::vector<double> lengths;
    for (auto& c : cells) {
        for (int k = 0; k < 3; ++k) {
            x0 = std::min(x0, c.x[k]);
            y0 = std::min(y0, c.y[k]);
            x1 = std::max(x1, c.x[k]);
            y1 = std::max(y1, c.y[k]);
            const Seg e = cell_edge(c, k);
            if (!(e.length() > GEOEPS)) { throw std::runtime_error("Cannot create a line using 2 identical points\n"); }
            lengths.push_back(e.length());
        }
        c.area = triangle_area(c);
        const double twice = std::fabs((c.x[1] - c.x[0]) * (c.y[2] - c.y[0]) - (c.x[2] - c.x[0]) * (c.y[1] - c.y[0]));
        if (!(twice > GEOEPS)) { throw std::runtime_error("These 3 points do not allow for a valid triangle\n"); }
        sensors[c.sensor].area += c.area;  // Cell::Cell -> Sensor::addToArea, cell.cpp:13-18
    }
    std::nth_element(lengths.begin(), lengths.begin() + lengths.size() / 2, lengths.end());
    const double h = std::max(lengths[lengths.size() / 2], 1e-6);
    EdgeGrid grid(h, x0, y0);
    EdgeGrid boxes(h, x0, y0);  // bounding boxes of the cells already placed, for the validity checks
    std::vector<uint32_t> stamp(cells.size() * 3, 0xFFFFFFFFu);
    std::vector<uint32_t> cell_stamp(cells.size(), 0xFFFFFFFFu);
    auto describe = [](const CellRec& c) {
        std::ostringstream os;
        os << "Triangle: [Point (" << c.x[0] << ", " << c.y[0] << "), Point (" << c.x[1] << ", " << c.y[1] << "), Point (" << c.x[2]
           << ", " << c.y[2] << ")]";
        return os.str();
    };
    auto orient = [](double ax, double ay, double bx, double by, double px, double py) {
        return (bx - ax) * (py - ay) - (px - ax) * (by - ay);
    };
    auto strictly_inside = [&](const CellRec& c, double px, double py) {
        const double d0 = orient(c.x[0], c.y[0], c.x[1], c.y[1], px, py), d1 = orient(c.x[1], c.y[1], c.x[2], c.y[2], px, py),
                     d2 = orient(c.x[2], c.y[2], c.x[0], c.y[0], px, py);
        const double eps = 1e-9 * (std::fabs(c.x[1] - c.x[0]) + std::fabs(c.y[1] - c.y[0]) + std::fabs(c.x[2] - c.x[0]) + std::fabs(c.y[2] - c.y[0]) + 1.);
        return (d0 > eps && d1 > eps && d2 > eps) || (d0 < -eps && d1 < -eps && d2 < -eps);
    };
    for (uint32_t ci = 0; ci < cells.size(); ++ci) {
        // Model::addCell -> Cell::validate (model.cpp:98-113, cell.cpp:21-25): an incoming cell must not be a
        // duplicate of, be contained in, or contain an existing cell.  (The reference's point-in-triangle test,
        // geometry.cpp:176-201, only fires next to the first vertex; the plain test is used here.)
        {
            const CellRec& c = cells[ci];
            const Seg box{ std::min({ c.x[0], c.x[1], c.x[2] }), std::min({ c.y[0], c.y[1], c.y[2] }),
                           std::max({ c.x[0], c.x[1], c.x[2] }), std::max({ c.y[0], c.y[1], c.y[2] }) };
            boxes.visit(box, [&](uint32_t cj) {
                if (cell_stamp[cj] == ci) { return; }
                cell_stamp[cj] = ci;
                const CellRec& o = cells[cj];
                int same = 0;
                for (int a = 0; a < 3; ++a) {
                    for (int b = 0; b < 3; ++b) {
                        const double dx = c.x[a] - o.x[b], dy = c.y[a] - o.y[b];
                        if (dx * dx + dy * dy < GEOEPS * GEOEPS) { ++same; }
                    }
                }
                if (same >= 3) { throw std::runtime_error("Duplicate cell detected.\n"); }
                // Triangle::intersects (geometry.cpp:160-183): edges that cross away from their end points
                for (int a = 0; a < 3; ++a) {
                    const Seg ea = cell_edge(c, a);
                    for (int b = 0; b < 3; ++b) {
                        const Seg eb = cell_edge(o, b);
                        const double o1 = orient(ea.ax, ea.ay, ea.bx, ea.by, eb.ax, eb.ay), o2 = orient(ea.ax, ea.ay, ea.bx, ea.by, eb.bx, eb.by);
                        const double o3 = orient(eb.ax, eb.ay, eb.bx, eb.by, ea.ax, ea.ay), o4 = orient(eb.ax, eb.ay, eb.bx, eb.by, ea.bx, ea.by);
                        const double tol = 1e-9 * (ea.length() * eb.length() + 1.);
                        if (((o1 > tol && o2 < -tol) || (o1 < -tol && o2 > tol)) && ((o3 > tol && o4 < -tol) || (o3 < -tol && o4 > tol))) {
                            throw std::runtime_error("Incoming " + describe(c) + "\nintersects\nExisting" + describe(o) + "\n");
                        }
                    }
                }
                // no edges cross: one vertex strictly inside means the whole cell lies inside the other
                for (int a = 0; a < 3; ++a) {
                    if (strictly_inside(o, c.x[a], c.y[a])) { throw std::runtime_error(describe(c) + "\nis contained within\n" + describe(o) + "\n"); }
                    if (strictly_inside(c, o.x[a], o.y[a])) { throw std::runtime_error(describe(o) + "\nis contained within\n" + describe(c) + "\n"); }
                }
            });
            boxes.insert(box, ci);
        }
        for (int a = 0; a < 3; ++a) {
            const uint32_t id = ci * 3 + static_cast<uint32_t>(a);
            const Seg ea = cell_edge(cells[ci], a);
            grid.visit(ea, [&](uint32_t other) {
                const uint32_t cj = other / 3;
                const int b = static_cast<int>(other % 3);
                if (cj == ci || stamp[other] == id) { return; }
                stamp[other] = id;
                const Seg eb = cell_edge(cells[cj], b);
                // Cell::findTransitionSurface, cell.cpp:81-98: the contained edge becomes a transition on both cells
                const Seg* shared = nullptr;
                if (seg_contains(eb, ea)) {
                    shared = &ea;
                } else if (seg_contains(ea, eb)) {
                    shared = &eb;
                }
                if (!shared) { return; }
                SubSurface on_i{ PSIM_SURF_TRANSITION, cj, static_cast<uint32_t>(b), edge_param(ea, shared->ax, shared->ay),
                                 edge_param(ea, shared->bx, shared->by), edge_param(eb, shared->ax, shared->ay),
                                 edge_param(eb, shared->bx, shared->by) };
                SubSurface on_j{ PSIM_SURF_TRANSITION, ci, static_cast<uint32_t>(a), on_i.t0, on_i.t1, on_i.s0, on_i.s1 };
                for (const auto& ex : cells[ci].transitions[a]) {
                    if (ranges_overlap(ex.s0, ex.s1, on_i.s0, on_i.s1)) {
                        throw std::runtime_error("An existing surface conflicts with the location of the incoming surface.\n");
                    }
                }
                for (const auto& ex : cells[cj].transitions[b]) {
                    if (ranges_overlap(ex.s0, ex.s1, on_j.s0, on_j.s1)) {
                        throw std::runtime_error("An existing surface conflicts with the location of the incoming surface.\n");
                    }
                }
                cells[ci].transitions[a].push_back(on_i);
                cells[cj].transitions[b].push_back(on_j);
            });
        }
        for (int a = 0; a < 3; ++a) { grid.insert(cell_edge(cells[ci], a), ci * 3 + static_cast<uint32_t>(a)); }
    }
}

void Model::attach_emit_surface(double p1x, double p1y, double p2x, double p2y, double temp, double duration, double start) {
    // Model::setEmitSurface, model.cpp:125-138
    if ((start < 0. || start >= simulation_time) || (duration < 0. || duration > simulation_time - start)) {
        throw std::runtime_error("Transient Surface start_time or duration specifications are invalid.\n");
    }
    if ((start > 0. || duration < simulation_time) && sim_type != SimType::Transient) {
        throw std::runtime_error("Cannot add a transient surface to a non transient simulation.\n");
    }
    const Seg line{ p1x, p1y, p2x, p2y };
    if (!(line.length() > GEOEPS)) { throw std::runtime_error("Cannot create a line using 2 identical points\n"); }
    // first cell (in model order), first edge of it, that contains the line (cell.cpp:29-35)
    for (uint32_t ci = 0; ci < cells.size(); ++ci) {
        // cheap reject before the exact test
        const CellRec& c = cells[ci];
        const double cx0 = std::min({ c.x[0], c.x[1], c.x[2] }) - 1e-6, cx1 = std::max({ c.x[0], c.x[1], c.x[2] }) + 1e-6;
        const double cy0 = std::min({ c.y[0], c.y[1], c.y[2] }) - 1e-6, cy1 = std::max({ c.y[0], c.y[1], c.y[2] }) + 1e-6;
        if (std::max(p1x, p2x) < cx0 || std::min(p1x, p2x) > cx1 || std::max(p1y, p2y) < cy0 || std::min(p1y, p2y) > cy1) { continue; }
        for (int k = 0; k < 3; ++k) {
            const Seg e = cell_edge(c, k);
            if (!seg_contains(e, line)) { continue; }
            EmitRec er{ p1x, p1y, p2x, p2y, temp, duration, start, ci, static_cast<uint32_t>(k), edge_param(e, p1x, p1y),
                        edge_param(e, p2x, p2y), line.length() };
            auto clash = [&](const std::vector<SubSurface>& v) {
                for (const auto& ex : v) {
                    if (ranges_overlap(ex.s0, ex.s1, er.s_p1, er.s_p2)) {
                        throw std::runtime_error("An existing surface conflicts with the location of the incoming surface.\n");
                    }
                }
            };
            clash(cells[ci].transitions[k]);
            clash(cells[ci].emits[k]);
            cells[ci].emits[k].push_back(SubSurface{ PSIM_SURF_EMIT, static_cast<uint32_t>(emitters.size()), 0, er.s_p1, er.s_p2, 0., 0. });
            emitters.push_back(er);
            return;
        }
    }
    throw std::runtime_error("Unable to add emitting surface.\n");
}

double Model::init_temp(const SensorRec& s) const {
    return sim_type == SimType::SteadyState ? s.t_steady : s.t_init;  // sensorController.h:65-67,83-85,100-102
}

void Model::prepare() {
    // setTemperatureBounds, model.cpp:203-218
    double lo = 1e300, hi = -1e300;
    for (const auto& c : cells) {
        const double t = init_temp(sensors[c.sensor]);
        lo = std::min(lo, t);
        hi = std::max(hi, t);
    }
    for (const auto& e : emitters) {
        lo = std::min(lo, e.temp);
        hi = std::max(hi, e.temp);
    }
    const double bound = phasor_sim ? TEMP_BOUND_EPS * 100. : TEMP_BOUND_EPS;
    lb_ = std::max(lo - bound, 0.);
    ub_ = hi + bound;
    temp_lo_ = lo;
    temp_hi_ = hi;
    // initializeMaterialTables, model.cpp:221-227 (+ SensorController::updateTables, sensorController.cpp:38-50)
    for (auto& m : materials) { m.set_temperature_grid(lo, hi); }
    for (auto& s : sensors) {
        s.heat_capacity = materials[s.material].base_energy(s.t_init);
        s.table_temp = s.t_init;
        if (sim_type == SimType::Transient) {
            s.steady_temps.assign(measurement_steps, s.t_init);
            s.heat_capacities.assign(measurement_steps, s.heat_capacity);
        }
    }
    prepared_ = true;
    iter_ = 0;
    iteration_ended_ = false;
    refresh();
}

double Model::total_initial_energy() {
    if (!prepared_) { throw std::runtime_error("prepare() must be called first"); }
    double total = 0.;
    for (const auto& c : cells) {
        const SensorRec& s = sensors[c.sensor];
        const double init = c.area * heat_capacity_at(s, 0);
        double cell_energy = (t_eq == 0.) ? init : init * std::fabs(init_temp(s) - t_eq);  // cell.cpp:38-41
        double emit = 0.;
        for (int k = 0; k < 3; ++k) {  // cell.cpp:45-63
            double edge_sum = 0.;
            for (const auto& sub : c.emits[k]) {
                const EmitRec& e = emitters[sub.target];
                const double en = e.length * e.duration * materials[s.material].emit_energy(e.temp) / 4.;
                edge_sum += (t_eq == 0.) ? en : en * std::fabs(e.temp - t_eq);
            }
            emit += edge_sum;
        }
        total += cell_energy + emit;
    }
    return total;
}

void Model::refresh() { eff_energy_ = total_initial_energy() / static_cast<double>(num_phonons); }

std::vector<psim_source> Model::source_counts(uint64_t seed) {
    // ModelSimulator::initPhononBuilders, modelSimulator.cpp:43-85.  The fractional phonon is resolved with a
    // hash of (seed, source ordinal) instead of the reference's unseeded urand(), so the integers depend on
    // the seed only - never on how many GPUs share the work.
    // A model in which every cell and every emitting surface sits at t_eq holds no energy: the reference divides by an
    // energy per phonon of zero there (NaN -> size_t, modelSimulator.cpp:44-49) and never returns; here it is an error.
    if (!(eff_energy_ > 0.) || !std::isfinite(eff_energy_)) {
        throw std::runtime_error("The model holds no energy to distribute: every cell and emitting surface is at t_eq.");
    }
    std::vector<psim_source> out;
    uint64_t ordinal = 0;
    auto phonons_for = [&](double energy) -> uint64_t {
        double whole = 0.;
        const double frac = std::modf(energy / eff_energy_, &whole);
        const double u = static_cast<double>(splitmix64(seed ^ splitmix64(++ordinal)) >> 11) * (1. / 9007199254740992.);
        return static_cast<uint64_t>(whole) + ((u < frac) ? 1u : 0u);
    };
    for (uint32_t ci = 0; ci < cells.size(); ++ci) {
        const CellRec& c = cells[ci];
        const SensorRec& s = sensors[c.sensor];
        const double init = c.area * heat_capacity_at(s, 0);
        const double init_energy = (t_eq == 0.) ? init : init * std::fabs(init_temp(s) - t_eq);
        if (const uint64_t n = phonons_for(init_energy); n > 0) {
            out.push_back(psim_source{ PSIM_SRC_CELL, ci, (init_temp(s) > t_eq) ? 1 : -1, 0, n });
        }
        for (int k = 0; k < 3; ++k) {
            for (const auto& sub : c.emits[k]) {
                const EmitRec& e = emitters[sub.target];
                const double factor = materials[s.material].emit_energy(e.temp) * e.duration * e.length / 4.;
                const double energy = (t_eq == 0.) ? factor : factor * std::fabs(t_eq - e.temp);
                if (const uint64_t n = phonons_for(energy); n > 0) {
                    out.push_back(psim_source{ PSIM_SRC_SURFACE, sub.target, (e.temp > t_eq) ? 1 : -1, 0, n });
                }
            }
        }
    }
    return out;
}

std::pair<int32_t*, double*> Model::tally_storage() {
    const size_t n = sensors.size() * recorded_steps;
    have_tallies_ = false;
    if (inc_energy_.size() != n) {
        inc_energy_.resize(n);
        inc_flux_.resize(2 * n);
    }
    return { inc_energy_.data(), inc_flux_.data() };
}

void Model::tallies_written() {
    have_tallies_ = true;
    iteration_ended_ = false;
}

void Model::set_tallies(const int32_t* energy, const double* flux) {
    const auto [e, f] = tally_storage();
    const size_t n = sensors.size() * recorded_steps;
    std::copy(energy, energy + n, e);
    std::copy(flux, flux + 2 * n, f);
    tallies_written();
}

template <class F> void Model::for_each_sensor(F&& f) {
    const size_t S = sensors.size();
    size_t threads = host_threads(S * recorded_steps / 65536);
    if (t_eq == 0.) { threads = 1; }  // Material::table() builds its tables on first use
    if (threads <= 1) {
        for (size_t si = 0; si < S; ++si) { f(si); }
        return;
    }
    run_parts(threads, [&](size_t t) {
        for (size_t si = S * t / threads; si < S * (t + 1) / threads; ++si) { f(si); }
    });
}

std::vector<double> Model::find_temperature(size_t si, size_t start) {
    const SensorRec& s = sensors[si];
    const int32_t* energies = inc_energy_.data() + si * recorded_steps;
    std::vector<double> temps(recorded_steps - start);
    Material& mat = materials[s.material];
    size_t index = 0;
    for (size_t r = start; r < recorded_steps; ++r) {
        const double energy = eff_energy_ * energies[r];
        if (t_eq != 0.) {
            temps[r - start] = energy / (s.area * heat_capacity_at(s, index++)) + t_eq;
        } else {  // numerical inversion on the tabulated energy density
            double temp = 0., ub = ub_, lb = lb_;
            std::size_t iter = 0;
            while ((ub - lb >= INVERSION_EPS) && (++iter != INVERSION_MAX_ITERS)) {
                temp = (ub + lb) / 2.;
                const double de = mat.base_energy(temp) * s.area - energy;
                (de < 0.) ? lb = temp : ub = temp;
            }
            temps[r - start] = temp;
        }
    }
    return temps;
}

SensorResult Model::scale_heat_params(size_t si) {
    const SensorRec& s = sensors[si];
    SensorResult sm;
    sm.id = s.id;
    sm.final_temps = find_temperature(si, 0);
    sm.final_temps.front() = init_temp(s);
    const double f = eff_energy_ / s.area;
    const size_t R = recorded_steps;
    const double* flux = inc_flux_.data() + 2 * si * R;
    sm.final_fluxes.resize(R);
    for (size_t r = 0; r < R; ++r) { sm.final_fluxes[r] = { flux[2 * r] * f, flux[2 * r + 1] * f }; }
    // mean and standard error of the three columns (sensorInterpreter.cpp:34-46): every sum runs over the steps in order,
    // as the reference's does, the three of them side by side
    const double n = static_cast<double>(R);
    double st = 0., sx = 0., sy = 0.;
    for (size_t r = 0; r < R; ++r) {
        st += sm.final_temps[r];
        sx += sm.final_fluxes[r][0];
        sy += sm.final_fluxes[r][1];
    }
    const double at = st / n, ax = sx / n, ay = sy / n;
    double qt = 0., qx = 0., qy = 0.;
    for (size_t r = 0; r < R; ++r) {
        qt += (at - sm.final_temps[r]) * (at - sm.final_temps[r]);
        qx += (ax - sm.final_fluxes[r][0]) * (ax - sm.final_fluxes[r][0]);
        qy += (ay - sm.final_fluxes[r][1]) * (ay - sm.final_fluxes[r][1]);
    }
    sm.t_steady = at;
    sm.std_t_steady = std::sqrt(qt / n) / std::sqrt(n);
    sm.x_flux = ax;
    sm.std_x_flux = std::sqrt(qx / n) / std::sqrt(n);
    sm.y_flux = ay;
    sm.std_y_flux = std::sqrt(qy / n) / std::sqrt(n);
    return sm;
}

bool Model::end_iteration(std::string* log) {
    if (!have_tallies_) { throw std::runtime_error("end of an iteration without tallies"); }
    ++iter_;
    // Model::resetRequired, model.cpp:250-272 - note its side effect on every sensor's steady temperature
    std::vector<char> is_stable(sensors.size(), 0);
    for_each_sensor([&](size_t si) {
        SensorRec& s = sensors[si];
        if (sim_type != SimType::Transient) {
            double t_final = 0.;
            if (s.area != 0.) {
                const auto temps = find_temperature(si, start_step);
                t_final = std::accumulate(temps.begin(), temps.end(), 0.) / static_cast<double>(recorded_steps - start_step);
            }
            is_stable[si] = std::fabs(t_final - s.t_steady) / s.t_steady <= SENSOR_RESET_THRESHOLD;
            s.t_steady = t_final;
        } else {
            auto temps = find_temperature(si, 0);
            bool ok = true;
            for (size_t r = 0; r < temps.size(); ++r) {
                if (!(std::fabs(temps[r] - s.steady_temps[r]) / s.steady_temps[r] <= TRANSIENT_RESET_THRESHOLD)) { ok = false; }
            }
            is_stable[si] = ok;
            s.steady_temps = std::move(temps);
        }
    });
    const int stable = static_cast<int>(std::count(is_stable.begin(), is_stable.end(), 1));
    stable_ = stable;
    if (log) { *log += "Stable sensors: " + std::to_string(stable) + "\n"; }
    // avgTemp(), model.cpp:230-239: area-weighted getSteadyTemp() of the sensors
    double new_t_eq = t_eq;
    if (t_eq != 0. && sim_type != SimType::Transient) {
        double total_area = 0.;
        for (const auto& s : sensors) { total_area += s.area; }
        new_t_eq = 0.;
        for (const auto& s : sensors) { new_t_eq += steady_temp(s) * s.area / total_area; }
    }
    const bool moved = (static_cast<size_t>(stable) * 100 / sensors.size() < SYSTEM_RESET_THRESHOLD) ||
                       (std::fabs(new_t_eq - t_eq) / t_eq * 1000. > TEQ_THRESHOLD);
    const bool again = moved && iter_ < max_iters && !phasor_sim;  // model.cpp:163
    if (again) {
        reset_iteration();
        t_eq = new_t_eq;
        if (log) {
            std::ostringstream os;
            os << "system not stable\nupdated t_eq: " << t_eq << "\n";
            *log += os.str();
        }
    }
    refresh();  // model.cpp:171 - in steady state this changes the energy per phonon used for the output scaling
    if (!again && iter_ >= max_iters && log) { *log += "System did not stabilize!!\n"; }  // model.cpp:173-176
    iteration_ended_ = true;
    return again;
}

// Model::reset(false), model.cpp:274-283 -> Sensor::reset -> the controllers' reset(false), sensorController.cpp:64-113
void Model::reset_iteration() {
    for (auto& s : sensors) {
        Material& mat = materials[s.material];
        switch (sim_type) {
        case SimType::SteadyState:  // tables and heat capacity at the steady temperature of the iteration that just ended
            s.table_temp = s.t_steady;
            s.heat_capacity = mat.base_energy(s.t_steady);
            break;
        case SimType::Periodic:     // tables only
            s.table_temp = s.t_steady;
            break;
        case SimType::Transient:    // one heat capacity and one scatter table per measurement step
            s.heat_capacities.resize(s.steady_temps.size());
            for (size_t k = 0; k < s.steady_temps.size(); ++k) { s.heat_capacities[k] = mat.base_energy(s.steady_temps[k]); }
            break;
        }
    }
    have_tallies_ = false;
}

int Model::finish_run(uint64_t run_id, std::string* log) {
    if (!have_tallies_) { throw std::runtime_error("finish_run without tallies"); }
    if (!iteration_ended_) {
        // a caller that drives one iteration per run (the reference's MAX_ITERS = 1)
        const uint64_t keep = max_iters;
        max_iters = iter_ + 1;
        end_iteration(log);
        max_iters = keep;
    }
    std::vector<SensorResult> res(sensors.size());
    for_each_sensor([&](size_t si) { res[si] = scale_heat_params(si); });
    std::sort(res.begin(), res.end(), [](const SensorResult& a, const SensorResult& b) { return a.id < b.id; });
    if (runs.size() <= run_id) { runs.resize(run_id + 1); }
    runs[run_id] = std::move(res);
    return stable_;
}

void Model::reset_for_next_run() {
    for (auto& s : sensors) {
        s.t_steady = s.t_init;  // controller reset(full_reset = true), sensorController.cpp:64-78,101-113
        if (sim_type == SimType::Transient) { s.steady_temps.assign(measurement_steps, s.t_init); }
    }
    have_tallies_ = false;
    prepared_ = false;
}

void Model::restore_file_state() {
    t_eq = t_eq_file_;  // a re-iterated run moves t_eq (model.cpp:165); the reference keeps it for the runs that follow
    reset_for_next_run();
}

std::vector<SensorResult> Model::averaged() const {
    if (runs.empty()) { return {}; }
    std::vector<SensorResult> avg = runs.front();
    const double n = static_cast<double>(runs.size());
    for (size_t r = 1; r < runs.size(); ++r) {
        for (size_t i = 0; i < avg.size(); ++i) {
            const SensorResult& x = runs[r][i];
            avg[i].t_steady += x.t_steady;
            avg[i].std_t_steady += x.std_t_steady;
            avg[i].x_flux += x.x_flux;
            avg[i].std_x_flux += x.std_x_flux;
            avg[i].y_flux += x.y_flux;
            avg[i].std_y_flux += x.std_y_flux;
            for (size_t k = 0; k < avg[i].final_temps.size(); ++k) {
                avg[i].final_temps[k] += x.final_temps[k];
                avg[i].final_fluxes[k][0] += x.final_fluxes[k][0];
                avg[i].final_fluxes[k][1] += x.final_fluxes[k][1];
            }
        }
    }
    for (auto& a : avg) {
        a.t_steady /= n;
        a.std_t_steady /= n;
        a.x_flux /= n;
        a.std_x_flux /= n;
        a.y_flux /= n;
        a.std_y_flux /= n;
        if (runs.size() > 1) {
            for (size_t k = 0; k < a.final_temps.size(); ++k) {
                a.final_temps[k] /= n;
                a.final_fluxes[k][0] /= n;
                a.final_fluxes[k][1] /= n;
            }
        }
    }
    return avg;
}

std::string Model::export_text(const std::string& model_filename, double seconds, const std::string& when) const {
    std::ostringstream out;
    const bool ss = sim_type == SimType::SteadyState;
    out << (ss ? "Steady State" : "Periodic") << " Results from " << std::quoted(model_filename) << " @ " << when
        << " - Time Taken " << seconds << "[s] over " << runs.size() << " runs\n";
    std::vector<SensorResult> avg;
    if (runs.size() != 1) { avg = averaged(); }
    const std::vector<SensorResult>& ms = runs.size() == 1 ? runs.front() : avg;  // one run: nothing to average, no copy of the traces
    if (ss) {  // steadyStateExport, outputManager.cpp:72-78
        for (const auto& m : ms) {
            out << m.t_steady << ' ' << m.std_t_steady << ' ' << m.x_flux << ' ' << m.std_x_flux << ' ' << m.y_flux << ' '
                << m.std_y_flux << '\n';
        }
        return out.str();
    }
    // periodicExport as intended (outputManager.cpp:82-114; at the reference's HEAD it reads an empty vector and crashes).
    // A million numbers per file: formatted with std::to_chars, which is specified to give what the reference's
    // `output << double` gives (printf %g with the stream's default precision of 6), at a fifth of the cost.
    if (ms.empty()) { return out.str(); }
    const size_t steps = ms.back().final_temps.size();
    const size_t I = step_interval;
    const size_t S = ms.size();
    const size_t blocks = (I == 0 || steps < I) ? 0 : steps / I;  // groups of I steps: [0, I), [I, 2I), ...
    // The means of every group, sensor by sensor (a sensor's traces are contiguous; going group by group through a thousand
    // sensors' vectors was a cache miss per number) ...
    std::vector<double> mean(blocks * S * 3);
    for (size_t si = 0; si < S; ++si) {
        const SensorResult& m = ms[si];
        for (size_t b = 0; b < blocks; ++b) {
            double t = 0., fx = 0., fy = 0.;
            for (size_t k = b * I; k < (b + 1) * I; ++k) {
                t += m.final_temps[k];
                fx += m.final_fluxes[k][0];
                fy += m.final_fluxes[k][1];
            }
            double* o = &mean[(b * S + si) * 3];
            o[0] = t / static_cast<double>(I);
            o[1] = fx / static_cast<double>(I);
            o[2] = fy / static_cast<double>(I);
        }
    }
    // ... then the text, group by group; several threads write the groups of their range, the pieces are joined in order
    auto write_groups = [&](size_t b0, size_t b1, std::string& text) {
        text.reserve((b1 - b0) * (S * 40 + 24));
        char buf[40];
        auto put = [&](double v, char sep) {
            const auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::general, 6);
            text.append(buf, r.ptr);
            text.push_back(sep);
        };
        for (size_t b = b0; b < b1; ++b) {
            text += std::to_string(b * I + I / 2);
            text.push_back('\n');
            text += std::to_string(S);
            text.push_back('\n');
            for (size_t si = 0; si < S; ++si) {
                const double* o = &mean[(b * S + si) * 3];
                put(o[0], ' ');
                put(o[1], ' ');
                put(o[2], '\n');
            }
        }
    };
    const size_t threads = std::max<size_t>(1, std::min(host_threads(blocks * S * 3 / 32768), blocks));
    std::vector<std::string> pieces(threads);
    run_parts(threads, [&](size_t t) { write_groups(blocks * t / threads, blocks * (t + 1) / threads, pieces[t]); });
    std::string text = out.str();
    size_t total = text.size();
    for (const auto& piece : pieces) { total += piece.size(); }
    text.reserve(total);
    for (const auto& piece : pieces) { text += piece; }
    return text;
}

void Model::export_results(const std::string& model_path, double seconds) const {
    // adjustPath, outputManager.cpp:124-130: <dir>/<ss_|per_><stem>.txt
    const size_t slash = model_path.find_last_of('/');
    const std::string dir = slash == std::string::npos ? "" : model_path.substr(0, slash + 1);
    const std::string file = slash == std::string::npos ? model_path : model_path.substr(slash + 1);
    const size_t dot = file.find_last_of('.');
    const std::string stem = (dot == std::string::npos || dot == 0) ? file : file.substr(0, dot);
    const std::string out_path = dir + (sim_type == SimType::SteadyState ? "ss_" : "per_") + stem + ".txt";
    const auto now = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
    std::tm tm_time{};
    std::string when = "Unknown Time";
    if (localtime_r(&now, &tm_time) != nullptr) {
        std::ostringstream oss;
        oss << std::put_time(&tm_time, "%Y-%m-%d %X");
        when = oss.str();
    }
    std::ofstream f(out_path, std::ios_base::trunc);
    if (!f.is_open()) { throw std::runtime_error("cannot write " + out_path); }
    f << export_text(file, seconds, when);
}

uint32_t Model::table_id(const Table& t) {
    auto it = table_index_.find(&t);
    if (it != table_index_.end()) { return it->second; }
    const uint32_t id = static_cast<uint32_t>(table_list_.size());
    table_list_.push_back(&t);
    table_index_[&t] = id;
    return id;
}

const psim_model_desc& Model::describe() {
    if (!prepared_) { throw std::runtime_error("prepare() must be called first"); }
    table_list_.clear();
    table_index_.clear();
    d_materials_.clear();
    d_velocities_.clear();
    d_sensors_.clear();
    d_cells_.clear();
    d_subs_.clear();
    d_emitters_.clear();
    d_tables_.clear();
    for (const auto& m : materials) {
        d_materials_.push_back(psim_material{ m.b_l, m.b_tn, m.b_tu, m.b_i, m.w, m.w_max_la, m.w_max_ta, m.freq_width });
        d_velocities_.insert(d_velocities_.end(), m.vel_la.begin(), m.vel_la.end());
        d_velocities_.insert(d_velocities_.end(), m.vel_ta.begin(), m.vel_ta.end());
    }
    for (const auto& s : sensors) {
        Material& m = materials[s.material];
        psim_sensor ds{};
        ds.material = s.material;
        // tables at table_temp: t_init (SensorController::updateTables), after a re-iteration the steady temperature
        // (SteadyState / Periodic reset(false)); a transient controller keeps base_table_ / scatter_table_ at t_init and uses
        // the per-step records below.  Rates at getSteadyTemp(0).
        ds.base_table = table_id(m.table(Material::Base, s.table_temp));
        ds.scatter_table = table_id(m.table(Material::Scatter, s.table_temp));
        ds.temperature = steady_temp(s);
        d_sensors_.push_back(ds);
    }
    for (const auto& e : emitters) {
        psim_emitter de{};
        de.cell = e.cell;
        de.edge = e.edge;
        de.table = table_id(materials[sensors[cells[e.cell].sensor].material].table(Material::Emit, e.temp));
        de.s_p1 = e.s_p1;
        de.s_p2 = e.s_p2;
        de.start_time = e.start;
        de.duration = e.duration;
        d_emitters_.push_back(de);
    }
    for (const auto& c : cells) {
        psim_cell dc{};
        for (int k = 0; k < 3; ++k) {
            dc.x[k] = c.x[k];
            dc.y[k] = c.y[k];
            dc.sub_first[k] = static_cast<uint32_t>(d_subs_.size());
            for (const auto* list : { &c.transitions[k], &c.emits[k] }) {  // transitions are searched first (compositeSurface.cpp:49-62)
                for (const auto& sub : *list) {
                    d_subs_.push_back(psim_subsurface{ sub.kind, sub.target, sub.target_edge, 0, sub.s0, sub.s1, sub.t0, sub.t1 });
                }
            }
            dc.sub_count[k] = static_cast<uint32_t>(d_subs_.size()) - dc.sub_first[k];
        }
        dc.specularity = c.spec;
        dc.sensor = c.sensor;
        d_cells_.push_back(dc);
    }
    for (const Table* t : table_list_) { d_tables_.push_back(psim_table{ t->cumulative.data(), t->la_fraction.data() }); }
    desc_ = psim_model_desc{};
    desc_.num_materials = static_cast<uint32_t>(d_materials_.size());
    desc_.num_sensors = static_cast<uint32_t>(d_sensors_.size());
    desc_.num_cells = static_cast<uint32_t>(d_cells_.size());
    desc_.num_subsurfaces = static_cast<uint32_t>(d_subs_.size());
    desc_.num_emitters = static_cast<uint32_t>(d_emitters_.size());
    desc_.num_tables = static_cast<uint32_t>(d_tables_.size());
    desc_.materials = d_materials_.data();
    desc_.velocities = d_velocities_.data();
    desc_.sensors = d_sensors_.data();
    desc_.cells = d_cells_.data();
    desc_.subsurfaces = d_subs_.data();
    desc_.emitters = d_emitters_.data();
    desc_.tables = d_tables_.data();
    desc_.measurement_steps = static_cast<uint32_t>(measurement_steps);
    desc_.step_adjustment = static_cast<uint32_t>(step_adjustment);
    desc_.simulation_time = simulation_time;
    desc_.full_simulation = (t_eq == 0.) ? 1u : 0u;
    desc_.phasor_sim = phasor_sim ? 1u : 0u;
    // a transient run in its second or later iteration: TransientController::getSteadyTemp(step) and scatter_tables_[step]
    // (sensorController.cpp:80-88,101-113) differ from step to step
    d_step_sensors_.clear();
    if (sim_type == SimType::Transient && iter_ > 0) {
        d_step_sensors_.reserve(sensors.size() * measurement_steps);
        for (size_t si = 0; si < sensors.size(); ++si) {
            const SensorRec& s = sensors[si];
            Material& m = materials[s.material];
            for (size_t k = 0; k < measurement_steps; ++k) {
                psim_sensor ds = d_sensors_[si];
                ds.temperature = (k == 0) ? s.t_init : s.steady_temps[k];
                ds.scatter_table = table_id(m.table(Material::Scatter, s.steady_temps[k]));
                d_step_sensors_.push_back(ds);
            }
        }
        // (table_id may have appended tables: rebuild the table list)
        d_tables_.clear();
        for (const Table* t : table_list_) { d_tables_.push_back(psim_table{ t->cumulative.data(), t->la_fraction.data() }); }
        desc_.num_tables = static_cast<uint32_t>(d_tables_.size());
        desc_.tables = d_tables_.data();
        desc_.step_sensors = d_step_sensors_.data();
    }
    return desc_;
}

}  // namespace psim
