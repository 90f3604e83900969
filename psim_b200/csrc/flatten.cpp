#include "flatten.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <sstream>
#include <array>
#include <unordered_map>

namespace psim {
namespace {

constexpr double HBAR = 1.054517e-34;   // reference material.cpp:12
constexpr double BOLTZ = 1.38065e-23;   // reference material.cpp:13
constexpr double PER_NS = 1e-9;         // rates are carried in 1/ns (reference SCALING_FACTOR, modelSimulator.cpp:17)

uint32_t transition_word(uint32_t cell, uint32_t edge, bool same_dir) {
    return (PSIM_LINK_TRANSITION << 30) | (edge << 28) | (same_dir ? (1u << 27) : 0u) | cell;
}

bool fail(std::string& err, int code, const std::string& msg, int& rc) {
    err = msg;
    rc = code;
    return false;
}


// The hints of the flight loop's fast path (device_core.cuh: fast_impact) in the link words of an image: which walls are
// perfectly specular, which composite edges have a transition into a cell of the same material and rate class behind them,
// and which transitions lead into a cell with the SAME shape record, material and rate class (PSIM_LINK_SAME_FRAME: the
// phonon's rates of motion and relaxation rates are the same on the other side, nothing has to be loaded to go on).
// Returns whether the image has any edge the fast path can take (a transition into a cell of the same material and rate
// class, whole-edge or behind a composite edge).
bool mark_fast_links(std::vector<DevCell>& cells, const std::vector<DevShape>& shapes, std::vector<DevSub>& subs) {
    bool any_fast = false;
    for (DevCell& c : cells) {
        const bool classified = PSIM_CELL_CLASS(c.sensor_mat) != 255u;
        auto same_rates = [&](uint32_t lw) { return classified && ((cells[PSIM_LINK_TARGET(lw)].sensor_mat ^ c.sensor_mat) & 0xFFFu) == 0u; };
        auto transition = [&](uint32_t lw) {
            const bool same_frame = same_rates(lw) && cells[PSIM_LINK_TARGET(lw)].shape == c.shape;
            return (lw & ~PSIM_LINK_SAME_FRAME) | (same_frame ? PSIM_LINK_SAME_FRAME : 0u);
        };
        for (uint32_t e = 0; e < 4; ++e) {
            const uint32_t w = c.link[e];
            if (PSIM_LINK_KIND(w) == PSIM_LINK_BOUNDARY) {
                c.link[e] = (PSIM_LINK_BOUNDARY << 30) | (shapes[c.shape].spec >= 1.f ? 1u : 0u);
            } else if (PSIM_LINK_KIND(w) == PSIM_LINK_TRANSITION) {
                c.link[e] = transition(w);
                any_fast |= same_rates(w);
            } else if (PSIM_LINK_KIND(w) == PSIM_LINK_COMPOSITE) {
                const uint32_t first = (w >> 7) & 0xFFFFFu, n = w & 0x7Fu;
                bool any = false;
                for (uint32_t i = 0; i < n; ++i) {
                    uint32_t& lw = subs[first + i].link;
                    if (PSIM_LINK_KIND(lw) != PSIM_LINK_TRANSITION) { continue; }
                    any |= same_rates(lw);
                    lw = transition(lw);
                }
                c.link[e] = (w & ~(1u << 27)) | (any ? (1u << 27) : 0u);
                any_fast |= any;
            }
        }
    }
    return any_fast;
}

struct Frame {
    double ox, oy, ux, uy, vx, vy;  // origin Q0 and the edge vectors u = Q1 - Q0, v = Q3 - Q0 (triangle: P2 - P1, P3 - P1)
};

// The lattice image of the flattened mesh `img` (device_types.h).  Fine flight cells A, B are lattice neighbours along b1 when
// both are parallelograms with the SAME shape record (bit-identical fp32 frame matrix, normals and specularity: their frames
// are parallel and equally oriented), the same material and rate class (a classified one: 255 always redraws), A's edge 1
// (b1 = 1) is as a whole a plain transition onto B's edge 3 (b1 = 0) and back, and B's origin is A's origin + u; along b2
// likewise with edges 2 / 0 and v.  Blocks are grown greedily from cells without a left / lower neighbour: a row to the
// right, then rows upwards while every cell of the row above exists, is free and continues the row.  Everything else - the
// triangles, parallelograms without such neighbours - is a block of one.  Leaves the lattice arrays empty when no block has
// more than one cell or an edge would need more than 127 sub-surfaces.
void build_lattices(HostImage& img, const std::vector<Frame>& frames) {
    const uint32_t F = static_cast<uint32_t>(img.cells.size());
    constexpr uint32_t kMaxSide = 120;
    std::vector<char> quad(F, 0);
    for (const DevApiCell& a : img.api_cells) { quad[PSIM_CELL_INDEX(a.cell)] = static_cast<char>(PSIM_CELL_QUAD(a.cell)); }
    auto neighbour = [&](uint32_t c, uint32_t e_out, uint32_t e_in, bool along_u) -> int {
        if (!quad[c]) { return -1; }
        const DevCell& A = img.cells[c];
        const uint32_t w = A.link[e_out];
        if (PSIM_LINK_KIND(w) != PSIM_LINK_TRANSITION || ((w >> 28) & 3u) != e_in || (w & (1u << 27)) || !((w >> 26) & 1u)) { return -1; }
        const uint32_t t = PSIM_LINK_TARGET(w);
        if (t == c || t >= F || !quad[t]) { return -1; }
        const DevCell& B = img.cells[t];
        const uint32_t back = B.link[e_in];
        if (PSIM_LINK_KIND(back) != PSIM_LINK_TRANSITION || PSIM_LINK_TARGET(back) != c || ((back >> 28) & 3u) != e_out || (back & (1u << 27))) { return -1; }
        if (B.shape != A.shape || ((B.sensor_mat ^ A.sensor_mat) & 0xFFFu) != 0u || PSIM_CELL_CLASS(A.sensor_mat) == 255u) { return -1; }
        const Frame &fa = frames[c], &fb = frames[t];
        const double dx = along_u ? fa.ux : fa.vx, dy = along_u ? fa.uy : fa.vy;
        const double tol = 1e-9 * (std::fabs(fa.ux) + std::fabs(fa.uy) + std::fabs(fa.vx) + std::fabs(fa.vy));
        if (std::fabs(fb.ox - (fa.ox + dx)) > tol || std::fabs(fb.oy - (fa.oy + dy)) > tol) { return -1; }
        return static_cast<int>(t);
    };
    std::vector<int> right(F), up(F);
    std::vector<char> has_left(F, 0), has_down(F, 0);
    for (uint32_t c = 0; c < F; ++c) {
        right[c] = neighbour(c, 1, 3, true);
        up[c] = neighbour(c, 2, 0, false);
        if (right[c] >= 0) { has_left[right[c]] = 1; }
        if (up[c] >= 0) { has_down[up[c]] = 1; }
    }
    struct Block {
        uint32_t nx, ny, sub_off;
        std::vector<uint32_t> cells;  // [iy * nx + ix]
    };
    std::vector<Block> blocks;
    std::vector<int> block_of(F, -1);
    std::vector<uint32_t> ix_of(F, 0), iy_of(F, 0);
    bool any_merge = false;
    auto grow = [&](uint32_t c0) {
        Block b{};
        std::vector<uint32_t> row{ c0 };
        for (int r = right[c0]; r >= 0 && block_of[r] < 0 && static_cast<uint32_t>(r) != c0 && row.size() < kMaxSide; r = right[r]) {
            if (std::find(row.begin(), row.end(), static_cast<uint32_t>(r)) != row.end()) { break; }  // (a ring of cells)
            row.push_back(static_cast<uint32_t>(r));
        }
        b.nx = static_cast<uint32_t>(row.size());
        b.ny = 0;
        const int id = static_cast<int>(blocks.size());
        for (;;) {
            for (uint32_t i = 0; i < b.nx; ++i) {
                block_of[row[i]] = id;
                ix_of[row[i]] = i;
                iy_of[row[i]] = b.ny;
                b.cells.push_back(row[i]);
            }
            ++b.ny;
            if (b.ny >= kMaxSide) { break; }
            std::vector<uint32_t> next(b.nx);
            bool ok = true;
            for (uint32_t i = 0; i < b.nx && ok; ++i) {
                const int u = up[row[i]];
                ok = u >= 0 && block_of[u] < 0 && (i == 0 || right[next[i - 1]] == u);
                if (ok) { next[i] = static_cast<uint32_t>(u); }
            }
            if (!ok) { break; }
            row = next;
        }
        any_merge |= b.cells.size() > 1;
        blocks.push_back(std::move(b));
    };
    for (uint32_t c = 0; c < F; ++c) {
        if (block_of[c] < 0 && !has_left[c] && !has_down[c]) { grow(c); }
    }
    for (uint32_t c = 0; c < F; ++c) {
        if (block_of[c] < 0) { grow(c); }
    }
    if (!any_merge) { return; }
    // blocks are numbered in the order of their first fine cell, so that a mesh without lattices maps to itself
    uint32_t sub_total = 0;
    for (Block& b : blocks) {
        b.sub_off = sub_total;
        sub_total += b.nx * b.ny;
    }
    std::vector<DevCell> cells(blocks.size());
    std::vector<DevShape> shapes;
    std::vector<DevSub> subs;
    std::vector<uint32_t> sub_fine(sub_total), sub_sensor(sub_total);
    double extent_sum = 0., extent_cells = 0.;
    auto shape_id = [&](const DevShape& sh) -> uint32_t {
        for (size_t i = 0; i < shapes.size(); ++i) {
            if (std::memcmp(&shapes[i], &sh, sizeof(sh)) == 0) { return static_cast<uint32_t>(i); }
        }
        shapes.push_back(sh);
        return static_cast<uint32_t>(shapes.size() - 1);
    };
    std::unordered_map<uint64_t, uint32_t> shape_cache;  // (fine shape, nx, ny) -> lattice shape
    auto word_of = [&](uint32_t block) { return block | (quad[blocks[block].cells[0]] ? (1u << 31) : 0u); };
    // where fine cell t's edge e lies on the edge of its block: position k of n, counted along the block's edge coordinate
    auto edge_slot = [&](uint32_t t, uint32_t e, uint32_t& k, uint32_t& n) -> bool {
        const Block& b = blocks[block_of[t]];
        const uint32_t ix = ix_of[t], iy = iy_of[t];
        if (!quad[t]) {
            k = 0, n = 1;
            return true;
        }
        switch (e) {
            case 0: k = ix, n = b.nx; return iy == 0;
            case 1: k = iy, n = b.ny; return ix == b.nx - 1;
            case 2: k = b.nx - 1 - ix, n = b.nx; return iy == b.ny - 1;
            default: k = b.ny - 1 - iy, n = b.ny; return ix == 0;
        }
    };
    bool ok = true;
    for (size_t bi = 0; bi < blocks.size() && ok; ++bi) {
        const Block& b = blocks[bi];
        const uint32_t c0 = b.cells[0];
        DevCell o{};
        o.sensor_mat = img.cells[c0].sensor_mat;
        o.tri[0] = b.sub_off;
        o.tri[1] = b.nx | (b.ny << 16);
        for (size_t i = 0; i < b.cells.size(); ++i) {
            sub_fine[b.sub_off + i] = b.cells[i] | (quad[b.cells[i]] ? (1u << 31) : 0u);
            sub_sensor[b.sub_off + i] = PSIM_CELL_SENSOR(img.cells[b.cells[i]].sensor_mat);
        }
        if (b.cells.size() > 1) {  // extent of one of its parallelograms across either pair of edges = 1 / |gradient of b1 (b2)|
            const DevShape& fs = img.shapes[img.cells[c0].shape];
            const double w1 = 1. / std::hypot(fs.m00, fs.m01), w2 = 1. / std::hypot(fs.m10, fs.m11);
            extent_sum += std::min(w1, w2) * static_cast<double>(b.cells.size());
            extent_cells += static_cast<double>(b.cells.size());
        }
        const uint64_t key = (static_cast<uint64_t>(img.cells[c0].shape) << 32) | (b.nx << 16) | b.ny;
        auto it = shape_cache.find(key);
        if (it == shape_cache.end()) {
            DevShape sh = img.shapes[img.cells[c0].shape];
            sh.m00 /= static_cast<float>(b.nx);  // b1 spans nx cells
            sh.m01 /= static_cast<float>(b.nx);
            sh.m10 /= static_cast<float>(b.ny);
            sh.m11 /= static_cast<float>(b.ny);
            it = shape_cache.emplace(key, shape_id(sh)).first;
        }
        o.shape = it->second;
        const uint32_t n_edges = quad[c0] ? 4u : 3u;
        for (uint32_t e = 0; e < n_edges && ok; ++e) {
            const uint32_t n = quad[c0] ? ((e & 1u) ? b.ny : b.nx) : 1u;
            std::vector<DevSub> list;
            for (uint32_t k = 0; k < n && ok; ++k) {
                uint32_t ix = 0, iy = 0;
                if (quad[c0]) {
                    ix = e == 0 ? k : (e == 1 ? b.nx - 1 : (e == 2 ? b.nx - 1 - k : 0u));
                    iy = e == 0 ? 0u : (e == 1 ? k : (e == 2 ? b.ny - 1 : b.ny - 1 - k));
                }
                const uint32_t c = b.cells[iy * b.nx + ix];
                const uint32_t w = img.cells[c].link[e];
                // one fine sub-surface (or whole-edge link) -> a sub-surface of the block's edge: its extent [s0, s1] of the fine
                // edge becomes [(k + s0) / n, (k + s1) / n]; a transition that puts a phonon at a f s_fine + b f on the
                // neighbour's fine edge puts it at (kT + that) / nT on the edge of the neighbour's block, s_fine = n s - k
                auto add = [&](double s0, double s1, double af, double bf, uint32_t lw) {
                    DevSub ds{};
                    ds.s0 = static_cast<float>((k + s0) / n);
                    ds.s1 = static_cast<float>((k + s1) / n);
                    if (PSIM_LINK_KIND(lw) == PSIM_LINK_TRANSITION) {
                        const uint32_t t = PSIM_LINK_TARGET(lw), te = (lw >> 28) & 3u;
                        uint32_t kT = 0, nT = 1;
                        if (t >= F || !edge_slot(t, te, kT, nT)) {
                            ok = false;
                            return;
                        }
                        ds.a = static_cast<float>(af * n / nT);
                        ds.b = static_cast<float>((kT + bf - af * k) / nT);
                        const uint32_t cw = word_of(static_cast<uint32_t>(block_of[t]));
                        ds.link = (PSIM_LINK_TRANSITION << 30) | (te << 28) | (af > 0. ? (1u << 27) : 0u) | ((cw >> 31) << 26) | (cw & 0x01FFFFFFu);
                    } else {
                        ds.link = lw;  // an emitting surface keeps its index
                    }
                    list.push_back(ds);
                };
                if (PSIM_LINK_KIND(w) == PSIM_LINK_BOUNDARY) { continue; }  // uncovered = wall
                if (PSIM_LINK_KIND(w) == PSIM_LINK_COMPOSITE) {
                    const uint32_t first = (w >> 7) & 0xFFFFFu, cnt = w & 0x7Fu;
                    for (uint32_t i = 0; i < cnt; ++i) {
                        const DevSub& fs = img.subs[first + i];
                        add(fs.s0, fs.s1, fs.a, fs.b, fs.link);
                    }
                } else if (PSIM_LINK_KIND(w) == PSIM_LINK_TRANSITION) {
                    const bool same = (w & (1u << 27)) != 0u;
                    add(0., 1., same ? 1. : -1., same ? 0. : 1., w);
                } else {
                    add(0., 1., 0., 0., w);
                }
            }
            if (!ok) { break; }
            if (list.empty()) {
                o.link[e] = PSIM_LINK_BOUNDARY << 30;
            } else if (n == 1 && list.size() == 1 && list[0].s0 <= 0.f && list[0].s1 >= 1.f &&
                       (PSIM_LINK_KIND(list[0].link) == PSIM_LINK_EMIT || (std::fabs(std::fabs(list[0].a) - 1.f) < 1e-6f && (list[0].b == 0.f || list[0].b == 1.f)))) {
                o.link[e] = list[0].link;  // whole edge onto a whole edge: the plain link word
            } else if (list.size() > 127 || subs.size() + list.size() >= (1u << 20)) {
                ok = false;
            } else {
                o.link[e] = (PSIM_LINK_COMPOSITE << 30) | (static_cast<uint32_t>(subs.size()) << 7) | static_cast<uint32_t>(list.size());
                subs.insert(subs.end(), list.begin(), list.end());
            }
        }
        cells[bi] = o;
    }
    if (!ok) { return; }
    std::vector<DevApiCell> api(img.api_cells.size());
    for (size_t a = 0; a < api.size(); ++a) {
        const uint32_t c = PSIM_CELL_INDEX(img.api_cells[a].cell);
        if (ix_of[c] >= (1u << 13) || iy_of[c] >= (1u << 13)) { return; }
        api[a] = DevApiCell{ word_of(static_cast<uint32_t>(block_of[c])), (img.api_cells[a].corners & 0x3Fu) | (ix_of[c] << 6) | (iy_of[c] << 19) };
    }
    std::vector<DevEmitter> emitters = img.emitters;
    for (DevEmitter& em : emitters) {
        const uint32_t c = PSIM_CELL_INDEX(em.cell);
        uint32_t k = 0, n = 1;
        if (!edge_slot(c, em.edge, k, n)) { return; }
        em.cell = word_of(static_cast<uint32_t>(block_of[c]));
        em.s_p1 = static_cast<float>((k + static_cast<double>(em.s_p1)) / n);
        em.s_p2 = static_cast<float>((k + static_cast<double>(em.s_p2)) / n);
    }
    if (subs.empty()) { subs.push_back(DevSub{}); }
    if (emitters.empty()) { emitters.push_back(DevEmitter{}); }
    img.lattice_cells = std::move(cells);
    img.lattice_shapes = std::move(shapes);
    img.lattice_subs = std::move(subs);
    img.lattice_sub_fine = std::move(sub_fine);
    img.lattice_sub_sensor = std::move(sub_sensor);
    float vmax = img.scalars.phasor ? 1000.f : 0.f;
    for (float v : img.velocities) { vmax = std::max(vmax, std::fabs(v)); }
    img.lattice_cells_per_step = extent_cells > 0. ? static_cast<double>(vmax) * img.scalars.step_time_d / (extent_sum / extent_cells) : 0.;
    img.lattice_api_cells = std::move(api);
    img.lattice_emitters = std::move(emitters);
}

}  // namespace

int flatten_model(const psim_model_desc& d, HostImage& out, std::string& err, int merge_cells) {
    int rc = 0;
    if (!d.materials || !d.sensors || !d.cells || !d.tables || !d.velocities || d.num_materials == 0 ||
        d.num_sensors == 0 || d.num_cells == 0 || d.num_tables == 0 || (d.num_emitters > 0 && !d.emitters) ||
        (d.num_subsurfaces > 0 && !d.subsurfaces)) {
        fail(err, PSIM_E_INVALID, "model description has empty or null arrays", rc);
        return rc;
    }
    if (d.num_materials > 16 || d.num_sensors >= (1u << 20) || d.num_cells >= (1u << 27) ||
        d.num_emitters >= (1u << 27) || d.num_subsurfaces >= (1u << 27)) {
        fail(err, PSIM_E_INVALID, "model exceeds packed-index limits (16 materials, 2^20 sensors, 2^27 cells)", rc);
        return rc;
    }
    if (d.measurement_steps < 2 || d.measurement_steps > (1u << 24) || !(d.simulation_time > 0.) ||
        !(d.simulation_time < 1e300) || d.step_adjustment >= d.measurement_steps) {
        fail(err, PSIM_E_INVALID, "invalid measurement_steps (2 ... 2^24) / simulation_time / step_adjustment", rc);
        return rc;
    }
    const double step_time = d.simulation_time / static_cast<double>(d.measurement_steps);

    out = HostImage{};
    out.material_consts.assign(d.materials, d.materials + d.num_materials);
    out.materials.resize(d.num_materials);
    for (uint32_t m = 0; m < d.num_materials; ++m) {
        out.materials[m].w_max_la = static_cast<float>(d.materials[m].w_max_la * PSIM_FREQ_SCALE);
        out.materials[m].w_max_ta = static_cast<float>(d.materials[m].w_max_ta * PSIM_FREQ_SCALE);
        out.materials[m].freq_width = static_cast<float>(d.materials[m].freq_width * PSIM_FREQ_SCALE);
        out.materials[m].pad = 0.f;
    }
    out.velocities.resize(static_cast<size_t>(d.num_materials) * 2 * PSIM_BINS);
    for (size_t i = 0; i < out.velocities.size(); ++i) { out.velocities[i] = static_cast<float>(d.velocities[i]); }

    out.tables.resize(static_cast<size_t>(d.num_tables) * PSIM_BINS);
    for (uint32_t t = 0; t < d.num_tables; ++t) {
        if (!d.tables[t].cumulative || !d.tables[t].la_fraction) {
            fail(err, PSIM_E_INVALID, "null table", rc);
            return rc;
        }
        for (uint32_t i = 0; i < PSIM_BINS; ++i) {
            float2 e;
            e.x = static_cast<float>(d.tables[t].cumulative[i]);
            e.y = static_cast<float>(d.tables[t].la_fraction[i]);
            out.tables[static_cast<size_t>(t) * PSIM_BINS + i] = e;
        }
    }
    // guide: for r in [k/G, (k+1)/G), G = PSIM_GUIDE, the inverse-CDF answer lies in (low, high]; both ends come from the
    // reference's own bisection (material.cpp:64-75) evaluated on the fp32 table at the bracket's end points
    out.guides.resize(static_cast<size_t>(d.num_tables) * PSIM_GUIDE);
    for (uint32_t t = 0; t < d.num_tables; ++t) {
        const float2* tab = out.tables.data() + static_cast<size_t>(t) * PSIM_BINS;
        auto bisect = [&](float r) {
            uint32_t lo = 0, hi = PSIM_BINS - 1;
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (r < tab[mid].x) {
                    hi = mid;
                } else {
                    lo = mid;
                }
            }
            return hi;
        };
        for (uint32_t k = 0; k < PSIM_GUIDE; ++k) {
            const uint32_t low = bisect(static_cast<float>(k) / PSIM_GUIDE) - 1u;
            const uint32_t high = (k + 1 == PSIM_GUIDE) ? PSIM_BINS - 1u : bisect(static_cast<float>(k + 1) / PSIM_GUIDE);
            out.guides[static_cast<size_t>(t) * PSIM_GUIDE + k] = low | (high << 16);
        }
    }

    // sensors: fold temperature powers and unit scalings into the rate coefficients (fp64 here, fp32 on device)
    auto dev_sensor = [&](const psim_sensor& in, DevSensor& o) -> bool {
        if (in.material >= d.num_materials || in.base_table >= d.num_tables || in.scatter_table >= d.num_tables ||
            !(in.temperature > 0.)) {
            return false;
        }
        const psim_material& mt = d.materials[in.material];
        const double T = in.temperature;
        const double k = 1. / PSIM_FREQ_SCALE;  // omega = k * w
        o = DevSensor{};
        o.c_la = static_cast<float>(mt.b_l * T * T * T * k * k * PER_NS);
        o.c_tn = static_cast<float>(mt.b_tn * T * T * T * T * k * PER_NS);
        o.c_tu = static_cast<float>(mt.b_tu * k * k * PER_NS);
        o.x_t = static_cast<float>(HBAR * k / (BOLTZ * T));
        o.c_i = static_cast<float>(mt.b_i * k * k * k * k * PER_NS);
        o.w_cut = static_cast<float>(mt.w * PSIM_FREQ_SCALE);
        o.scatter_table = in.scatter_table;
        o.base_table = in.base_table;
        return true;
    };
    out.sensors.resize(d.num_sensors);
    out.sensor_temperature.resize(d.num_sensors);
    for (uint32_t s = 0; s < d.num_sensors; ++s) {
        if (!dev_sensor(d.sensors[s], out.sensors[s])) {
            fail(err, PSIM_E_INVALID, "sensor refers to a missing material/table or has a non-positive temperature", rc);
            return rc;
        }
        out.sensor_temperature[s] = d.sensors[s].temperature;
    }
    // a transient run that re-iterates: one record per (sensor, measurement step)
    if (d.step_sensors) {
        out.step_sensors.resize(static_cast<size_t>(d.num_sensors) * d.measurement_steps);
        for (uint32_t s = 0; s < d.num_sensors; ++s) {
            for (uint32_t k = 0; k < d.measurement_steps; ++k) {
                psim_sensor in = d.step_sensors[static_cast<size_t>(s) * d.measurement_steps + k];
                in.material = d.sensors[s].material;      // a sensor area does not change its material ...
                in.base_table = d.sensors[s].base_table;  // ... nor the table its initial phonons are drawn from
                if (!dev_sensor(in, out.step_sensors[static_cast<size_t>(s) * d.measurement_steps + k])) {
                    fail(err, PSIM_E_INVALID, "per-step sensor record refers to a missing table or has a non-positive temperature", rc);
                    return rc;
                }
            }
        }
    }

    // emitters
    out.emitters.resize(d.num_emitters);
    for (uint32_t e = 0; e < d.num_emitters; ++e) {
        const psim_emitter& in = d.emitters[e];
        if (in.cell >= d.num_cells || in.edge > 2 || in.table >= d.num_tables || in.duration < 0. || in.start_time < 0.) {
            fail(err, PSIM_E_INVALID, "emitter refers to a missing cell/edge/table or has a negative window", rc);
            return rc;
        }
        DevEmitter o{};
        o.cell = in.cell;
        o.edge = in.edge;
        o.s_p1 = static_cast<float>(in.s_p1);
        o.s_p2 = static_cast<float>(in.s_p2);
        o.table = in.table;
        o.start = in.start_time;
        o.duration = in.duration;
        // the reference's own test, evaluated per measurement step in fp64 (surface.cpp:61-65):
        //   reflect if  k*dt < start  ||  k*dt + dt > start + duration ; absorb otherwise
        uint32_t k_on = d.measurement_steps, k_off = 0;
        for (uint32_t k = 0; k < d.measurement_steps; ++k) {
            const double pt = static_cast<double>(k) * step_time;
            if (!(pt < in.start_time || pt + step_time > in.start_time + in.duration)) {
                k_on = std::min(k_on, k);
                k_off = k + 1;
            }
        }
        if (k_off == 0) { k_on = 0; }
        o.k_on = k_on;
        o.k_off = k_off;
        out.emitters[e] = o;
    }

    // rate classes: sensors of the same material at the same temperature (and with the same tables) have identical
    // records; the per-phonon loop reads the record of the class, so a mesh with thousands of sensor areas at a handful
    // of temperatures touches a handful of records
    std::vector<uint32_t> sensor_class(d.num_sensors, 255u);
    {
        struct Key {
            uint32_t material, base_table, scatter_table;
            double temperature;
            bool operator==(const Key& o) const {
                return material == o.material && base_table == o.base_table && scatter_table == o.scatter_table && temperature == o.temperature;
            }
        };
        std::vector<Key> classes;
        for (uint32_t s = 0; s < d.num_sensors; ++s) {
            const Key key{ d.sensors[s].material, d.sensors[s].base_table, d.sensors[s].scatter_table, d.sensors[s].temperature };
            size_t k = 0;
            while (k < classes.size() && !(classes[k] == key)) { ++k; }
            if (k == classes.size() && classes.size() < 255) {
                classes.push_back(key);
                out.classes.push_back(out.sensors[s]);
            }
            // (per-step records: the rates change from step to step even inside one sensor area, so nothing is classified)
            sensor_class[s] = (k < 255 && !d.step_sensors) ? static_cast<uint32_t>(k) : 255u;
        }
    }

    // ---- cells.  Pass 1: every model-file triangle with its links in MODEL terms (target = model cell, model edge)
    const uint32_t C = d.num_cells;
    if (C >= (1u << 25)) {
        fail(err, PSIM_E_INVALID, "model exceeds packed-index limits (2^25 cells)", rc);
        return rc;
    }
    struct Tri {
        double x[3], y[3];
        uint32_t link[3];
        float spec;
    };
    std::vector<Tri> tri(C);
    for (uint32_t c = 0; c < C; ++c) {
        const psim_cell& in = d.cells[c];
        if (in.sensor >= d.num_sensors) {
            fail(err, PSIM_E_INVALID, "cell refers to a missing sensor", rc);
            return rc;
        }
        Tri& t = tri[c];
        for (int k = 0; k < 3; ++k) {
            t.x[k] = in.x[k];
            t.y[k] = in.y[k];
        }
        const double e1x = in.x[1] - in.x[0], e1y = in.y[1] - in.y[0];
        const double e2x = in.x[2] - in.x[0], e2y = in.y[2] - in.y[0];
        if (!(std::fabs(e1x * e2y - e2x * e1y) > 0.)) {
            std::ostringstream os;
            os << "cell " << c << " is degenerate";
            fail(err, PSIM_E_INVALID, os.str(), rc);
            return rc;
        }
        t.spec = static_cast<float>(std::min(1., std::max(0., in.specularity)));
        for (int k = 0; k < 3; ++k) {
            const uint32_t n = in.sub_count[k], first = in.sub_first[k];
            if (n == 0) {
                t.link[k] = PSIM_LINK_BOUNDARY << 30;
                continue;
            }
            if (!d.subsurfaces || n > 127 || first >= d.num_subsurfaces || n > d.num_subsurfaces - first) {  // (no 32-bit wrap-around)
                fail(err, PSIM_E_INVALID, "cell edge sub-surface range is out of bounds", rc);
                return rc;
            }
            auto word = [&](const psim_subsurface& sb, float& a, float& b) -> uint32_t {
                if (sb.kind == PSIM_SURF_EMIT) {
                    a = 0.f;
                    b = 0.f;
                    return (PSIM_LINK_EMIT << 30) | sb.target;
                }
                const double aa = (sb.t1 - sb.t0) / (sb.s1 - sb.s0);
                a = static_cast<float>(aa);
                b = static_cast<float>(sb.t0 - aa * sb.s0);
                return transition_word(sb.target, sb.target_edge, aa > 0.);
            };
            for (uint32_t i = 0; i < n; ++i) {
                const psim_subsurface& sb = d.subsurfaces[first + i];
                const bool ok = (sb.kind == PSIM_SURF_EMIT && sb.target < d.num_emitters) ||
                                (sb.kind == PSIM_SURF_TRANSITION && sb.target < d.num_cells && sb.target_edge < 3);
                if (!ok || sb.s0 == sb.s1) {
                    fail(err, PSIM_E_INVALID, "invalid sub-surface record", rc);
                    return rc;
                }
            }
            const psim_subsurface& s0 = d.subsurfaces[first];
            const double lo = std::min(s0.s0, s0.s1), hi = std::max(s0.s0, s0.s1);
            const bool whole = n == 1 && lo < 1e-9 && hi > 1. - 1e-9;
            float a, b;
            if (whole && (s0.kind == PSIM_SURF_EMIT ||
                          (std::fabs(std::fabs(s0.t1 - s0.t0) - 1.) < 1e-9 && std::min(s0.t0, s0.t1) < 1e-9))) {
                t.link[k] = word(s0, a, b);  // the whole edge is one neighbour / one emitter: no sub-table lookup
                continue;
            }
            if (out.subs.size() + n >= (1u << 20)) {
                fail(err, PSIM_E_INVALID, "too many partial-edge sub-surfaces", rc);
                return rc;
            }
            t.link[k] = (PSIM_LINK_COMPOSITE << 30) | (static_cast<uint32_t>(out.subs.size()) << 7) | n;
            for (uint32_t i = 0; i < n; ++i) {
                const psim_subsurface& sb = d.subsurfaces[first + i];
                DevSub ds{};
                ds.s0 = static_cast<float>(std::min(sb.s0, sb.s1));
                ds.s1 = static_cast<float>(std::max(sb.s0, sb.s1));
                ds.link = word(sb, ds.a, ds.b);
                out.subs.push_back(ds);
            }
        }
    }

    // Pass 2: pair triangles into parallelograms where that is exactly neutral (device_types.h): same sensor area, same
    // specularity, the shared edge a plain whole-edge transition in both directions, the union a parallelogram.
    std::vector<int> partner(C, -1), shared(C, -1);
    std::vector<char> opposite_pair(C, 0);  // the pair's second triangle has the opposite orientation
    if (merge_cells) {
        for (uint32_t A = 0; A < C; ++A) {
            if (partner[A] >= 0) { continue; }
            for (uint32_t k = 0; k < 3 && partner[A] < 0; ++k) {
                const uint32_t w = tri[A].link[k];
                if (PSIM_LINK_KIND(w) != PSIM_LINK_TRANSITION) { continue; }
                const uint32_t B = PSIM_LINK_TARGET(w), j = (w >> 28) & 3u;
                if (B == A || partner[B] >= 0 || d.cells[A].sensor != d.cells[B].sensor || tri[A].spec != tri[B].spec) { continue; }
                const uint32_t back = tri[B].link[j];
                if (PSIM_LINK_KIND(back) != PSIM_LINK_TRANSITION || PSIM_LINK_TARGET(back) != A || ((back >> 28) & 3u) != k ||
                    ((back ^ w) & (1u << 27))) {
                    continue;
                }
                // Q0 = A[k+1], Q1 = A[k+2], Q2 = A[k] (A's copy of the diagonal runs Q2 -> Q0), Q3 = B[j+2]: parallelogram iff
                // Q3 - Q0 = Q2 - Q1.  B's copy of the diagonal runs Q0 -> Q2 if B has A's orientation (the usual case: the
                // link's same-direction flag is clear) and Q2 -> Q0 if it has the opposite one.
                const bool opposite = (w & (1u << 27)) != 0u;
                const Tri &ta = tri[A], &tb = tri[B];
                const uint32_t a0 = (k + 1) % 3, a1 = (k + 2) % 3, a2 = k, b3 = (j + 2) % 3;
                const uint32_t b_at_q0 = opposite ? (j + 1) % 3 : j, b_at_q2 = opposite ? j : (j + 1) % 3;
                const double scale = std::fabs(ta.x[a1] - ta.x[a0]) + std::fabs(ta.y[a1] - ta.y[a0]) + std::fabs(ta.x[a2] - ta.x[a1]) + std::fabs(ta.y[a2] - ta.y[a1]);
                const double tol = 1e-9 * scale;
                const bool same_diagonal = std::fabs(tb.x[b_at_q0] - ta.x[a0]) <= tol && std::fabs(tb.y[b_at_q0] - ta.y[a0]) <= tol &&
                                           std::fabs(tb.x[b_at_q2] - ta.x[a2]) <= tol && std::fabs(tb.y[b_at_q2] - ta.y[a2]) <= tol;
                const bool parallelogram = std::fabs((tb.x[b3] - ta.x[a0]) - (ta.x[a2] - ta.x[a1])) <= tol &&
                                           std::fabs((tb.y[b3] - ta.y[a0]) - (ta.y[a2] - ta.y[a1])) <= tol;
                if (!same_diagonal || !parallelogram) { continue; }
                partner[A] = static_cast<int>(B);
                partner[B] = static_cast<int>(A);
                opposite_pair[A] = opposite_pair[B] = opposite ? 1 : 0;
                shared[A] = static_cast<int>(k);
                shared[B] = static_cast<int>(j);
            }
        }
    }

    // Pass 3: flight cells.  (model cell, model edge) -> (flight cell word, flight edge); the shared edge of a pair vanishes.
    std::vector<uint32_t> cell_word(C, 0);
    std::vector<std::array<int, 3>> edge_map(C, std::array<int, 3>{ 0, 1, 2 });
    std::vector<std::array<char, 3>> edge_flip(C, std::array<char, 3>{ 0, 0, 0 });  // the flight edge runs against the model edge
    out.api_cells.assign(C, DevApiCell{});
    // distinct shapes, found by exact comparison of the fp32 records (hash of the bit patterns -> candidates)
    std::unordered_multimap<uint64_t, uint32_t> shape_index;
    auto shape_id = [&](const DevShape& sh) -> uint32_t {
        uint32_t w[sizeof(DevShape) / 4];
        std::memcpy(w, &sh, sizeof(sh));
        uint64_t hsh = 0xcbf29ce484222325ull;
        for (uint32_t x : w) { hsh = (hsh ^ x) * 0x100000001b3ull; }
        const auto range = shape_index.equal_range(hsh);
        for (auto it = range.first; it != range.second; ++it) {
            if (std::memcmp(&out.shapes[it->second], &sh, sizeof(sh)) == 0) { return it->second; }
        }
        out.shapes.push_back(sh);
        shape_index.emplace(hsh, static_cast<uint32_t>(out.shapes.size() - 1));
        return static_cast<uint32_t>(out.shapes.size() - 1);
    };
    auto unit = [](double x, double y, float* o) {
        const double n = std::sqrt(x * x + y * y);
        o[0] = static_cast<float>(x / n);
        o[1] = static_cast<float>(y / n);
    };
    struct Pending {
        uint32_t model_cell[4];  // which model cell / edge each flight edge comes from
        uint32_t model_edge[4];
        uint32_t n_edges;
    };
    std::vector<Pending> pending;
    std::vector<Frame> frames;  // origin and edge vectors of every flight cell (fp64), for build_lattices
    for (uint32_t A = 0; A < C; ++A) {
        if (partner[A] >= 0 && static_cast<uint32_t>(partner[A]) < A) { continue; }  // made together with its partner
        const uint32_t ic = static_cast<uint32_t>(out.cells.size());
        const Tri& ta = tri[A];
        DevCell o{};
        DevShape sh{};
        Pending pe{};
        double ox, oy, ux, uy, vx, vy;
        const bool quad = partner[A] >= 0;
        if (quad) {
            const uint32_t B = static_cast<uint32_t>(partner[A]), k = static_cast<uint32_t>(shared[A]), j = static_cast<uint32_t>(shared[B]);
            const Tri& tb = tri[B];
            const uint32_t a0 = (k + 1) % 3, a1 = (k + 2) % 3, a2 = k, b3 = (j + 2) % 3;
            ox = ta.x[a0], oy = ta.y[a0];
            ux = ta.x[a1] - ox, uy = ta.y[a1] - oy;
            vx = tb.x[b3] - ox, vy = tb.y[b3] - oy;
            const bool opposite = opposite_pair[A] != 0;
            // flight edges 2 (Q2 -> Q3) and 3 (Q3 -> Q0): B's edges j+1, j+2 in B's own direction if B has A's orientation,
            // else its edges j+2 (Q3 -> Q2) and j+1 (Q0 -> Q3), both run backwards
            const uint32_t be2 = opposite ? (j + 2) % 3 : (j + 1) % 3, be3 = opposite ? (j + 1) % 3 : (j + 2) % 3;
            pe = Pending{ { A, A, B, B }, { a0, a1, be2, be3 }, 4 };
            cell_word[A] = cell_word[B] = ic | (1u << 31);
            edge_map[A][a0] = 0, edge_map[A][a1] = 1, edge_map[A][k] = -1;
            edge_map[B][be2] = 2, edge_map[B][be3] = 3, edge_map[B][j] = -1;
            edge_flip[B][be2] = edge_flip[B][be3] = opposite ? 1 : 0;
            // corners of the model triangles in this frame, two bits (b1, b2) per vertex
            auto corners = [](uint32_t v_at_00, uint32_t v_at_second, uint32_t second, uint32_t v_at_third, uint32_t third) {
                return (0u << (2 * v_at_00)) | (second << (2 * v_at_second)) | (third << (2 * v_at_third));
            };
            out.api_cells[A] = DevApiCell{ cell_word[A], corners(a0, a1, 1u /* (1,0) */, a2, 3u /* (1,1) */) };
            out.api_cells[B] = DevApiCell{ cell_word[B], corners(opposite ? (j + 1) % 3 : j, opposite ? j : (j + 1) % 3, 3u /* (1,1) */, b3, 2u /* (0,1) */) };
            o.tri[0] = A;  // below the diagonal b1 = b2 (it holds the corner (1, 0))
            o.tri[1] = B;
        } else {
            ox = ta.x[0], oy = ta.y[0];
            ux = ta.x[1] - ox, uy = ta.y[1] - oy;
            vx = ta.x[2] - ox, vy = ta.y[2] - oy;
            pe = Pending{ { A, A, A, A }, { 0, 1, 2, 0 }, 3 };
            cell_word[A] = ic;
            out.api_cells[A] = DevApiCell{ ic, (1u << 2) | (2u << 4) };  // (0,0), (1,0), (0,1)
            o.tri[0] = o.tri[1] = A;
        }
        const double det = ux * vy - vx * uy;
        const double m00 = vy / det, m01 = -vx / det, m10 = -uy / det, m11 = ux / det;
        sh.m00 = static_cast<float>(m00);
        sh.m01 = static_cast<float>(m01);
        sh.m10 = static_cast<float>(m10);
        sh.m11 = static_cast<float>(m11);
        unit(m10, m11, sh.n[0]);                                   // edge 0 is b2 = 0: inward = grad b2
        if (quad) {
            unit(-m00, -m01, sh.n[1]);                             // edge 1 is b1 = 1
            unit(-m10, -m11, sh.n[2]);                             // edge 2 is b2 = 1
            unit(m00, m01, sh.n[3]);                               // edge 3 is b1 = 0
        } else {
            unit(-(m00 + m10), -(m01 + m11), sh.n[1]);             // edge 1 is b1 + b2 = 1
            unit(m00, m01, sh.n[2]);                               // edge 2 is b1 = 0
        }
        sh.spec = ta.spec;
        const uint32_t sensor = d.cells[A].sensor;
        o.sensor_mat = (sensor << 12) | (sensor_class[sensor] << 4) | d.sensors[sensor].material;
        o.shape = shape_id(sh);
        out.cells.push_back(o);
        pending.push_back(pe);
        frames.push_back(Frame{ ox, oy, ux, uy, vx, vy });
    }
    // Pass 4: links in flight terms.  A flight edge that runs against its model edge (edge_flip) reverses the edge
    // coordinate: s -> 1 - s on the owner's side, t -> 1 - t on the side of whoever points at it.
    auto relabel = [&](uint32_t w, bool own_flipped) -> uint32_t {
        if (PSIM_LINK_KIND(w) != PSIM_LINK_TRANSITION) { return w; }
        const uint32_t t = PSIM_LINK_TARGET(w), e = (w >> 28) & 3u;
        const uint32_t cw = cell_word[t];
        const int fe = edge_map[t][e];  // >= 0: nothing but the shared edge of a pair vanishes, and nothing else points at it
        const bool same_dir = (((w >> 27) & 1u) != 0u) != (own_flipped != (edge_flip[t][e] != 0));
        return (PSIM_LINK_TRANSITION << 30) | (static_cast<uint32_t>(fe < 0 ? 0 : fe) << 28) | (same_dir ? (1u << 27) : 0u) | ((cw >> 31) << 26) |
               (cw & 0x01FFFFFFu);
    };
    for (size_t ic = 0; ic < out.cells.size(); ++ic) {
        const Pending& pe = pending[ic];
        for (uint32_t e = 0; e < pe.n_edges; ++e) {
            const uint32_t mc = pe.model_cell[e], me = pe.model_edge[e];
            const bool flipped = edge_flip[mc][me] != 0;
            const uint32_t w = tri[mc].link[me];
            if (PSIM_LINK_KIND(w) == PSIM_LINK_COMPOSITE) {  // the partial-edge records of this edge (referenced from here only)
                const uint32_t first = (w >> 7) & 0xFFFFFu, n = w & 0x7Fu;
                for (uint32_t i = 0; i < n; ++i) {
                    DevSub& sb = out.subs[first + i];
                    if (flipped) {  // s' = 1 - s: extent [1 - s1, 1 - s0], position on the neighbour a (1 - s') + b
                        const float s0 = sb.s0, s1 = sb.s1;
                        sb.s0 = 1.f - s1;
                        sb.s1 = 1.f - s0;
                        sb.b = sb.a + sb.b;
                        sb.a = -sb.a;
                    }
                    if (PSIM_LINK_KIND(sb.link) == PSIM_LINK_TRANSITION && edge_flip[PSIM_LINK_TARGET(sb.link)][(sb.link >> 28) & 3u]) {
                        sb.a = -sb.a;  // t' = 1 - t
                        sb.b = 1.f - sb.b;
                    }
                    sb.link = relabel(sb.link, flipped);
                }
            }
            out.cells[ic].link[e] = relabel(w, flipped);
        }
    }
    for (uint32_t e = 0; e < d.num_emitters; ++e) {
        DevEmitter& em = out.emitters[e];
        const uint32_t c = em.cell, me = em.edge;  // (model cell, model edge) until here
        const int fe = edge_map[c][me];
        if (fe < 0) {
            fail(err, PSIM_E_INVALID, "emitting surface on an edge shared by two cells", rc);
            return rc;
        }
        em.cell = cell_word[c];
        em.edge = static_cast<uint32_t>(fe);
        if (edge_flip[c][me]) {
            em.s_p1 = 1.f - em.s_p1;
            em.s_p2 = 1.f - em.s_p2;
        }
    }
    if (out.subs.empty()) { out.subs.push_back(DevSub{}); }
    if (out.emitters.empty()) { out.emitters.push_back(DevEmitter{}); }

    DevParams& P = out.scalars;
    P.n_cells = d.num_cells;
    P.n_flight_cells = static_cast<uint32_t>(out.cells.size());
    P.n_shapes = static_cast<uint32_t>(out.shapes.size());
    P.n_sensors = d.num_sensors;
    P.n_materials = d.num_materials;
    P.n_tables = d.num_tables;
    P.n_emitters = d.num_emitters;
    P.n_sources = 0;
    P.num_steps = d.measurement_steps;
    P.first_tally_step = d.step_adjustment;
    P.recorded_steps = d.measurement_steps - d.step_adjustment;
    P.full_mode = d.full_simulation ? 1u : 0u;
    P.phasor = d.phasor_sim ? 1u : 0u;
    P.step_time = static_cast<float>(step_time);
    P.step_time_inv = static_cast<float>(1. / step_time);
    P.step_time_d = step_time;
    if (merge_cells >= 2) { build_lattices(out, frames); }
    out.fast_links = mark_fast_links(out.cells, out.shapes, out.subs);
    out.lattice_fast_links = mark_fast_links(out.lattice_cells, out.lattice_shapes, out.lattice_subs);
    P.fast_links = out.fast_links ? 1u : 0u;
    return 0;
}

int plan_births(const HostImage& img, const psim_source* sources, size_t n, uint32_t shard, uint32_t num_shards,
                BirthPlan& out, std::string& err) {
    out = BirthPlan{};
    if (num_shards == 0 || shard >= num_shards || (n > 0 && !sources)) {
        err = "invalid shard / source arguments";
        return PSIM_E_INVALID;
    }
    const DevParams& P = img.scalars;
    const uint32_t M = P.num_steps;
    const double dt = P.step_time_d;
    const uint32_t last = M - 1;  // the interval that starts at step M-1 records nothing: not simulated
    uint64_t next_id = 0;
    std::vector<DevBirth> births;
    auto add = [&](uint32_t src, uint32_t step, uint64_t first_id, uint64_t ja, uint64_t jb) -> bool {
        if (jb <= ja) { return true; }
        const uint64_t G = num_shards;
        const uint64_t r = (first_id + ja) % G;
        const uint64_t j0 = ja + ((shard + G - r) % G);
        if (j0 >= jb) { return true; }
        const uint64_t cnt = (jb - 1 - j0) / G + 1;
        if (step >= last) {
            out.shard_unrecorded += cnt;
            return true;
        }
        if (cnt > 0xFFFFFFFFull) { return false; }
        DevBirth b{};
        b.source = src;
        b.step = step;
        b.j0 = j0;
        b.count = static_cast<uint32_t>(cnt);
        b.stride = num_shards;
        births.push_back(b);
        out.shard_phonons += cnt;
        return true;
    };
    for (size_t i = 0; i < n; ++i) {
        const psim_source& s = sources[i];
        DevSource ds{};
        ds.kind = s.kind;
        ds.index = s.index;
        ds.sign = s.sign >= 0 ? 1 : -1;
        ds.count = s.count;
        ds.first_id = next_id;
        if (s.kind == PSIM_SRC_CELL) {
            if (s.index >= P.n_cells) {
                err = "cell source refers to a missing cell";
                return PSIM_E_INVALID;
            }
            if (!add(static_cast<uint32_t>(i), 0, next_id, 0, s.count)) {
                err = "more than 2^32 phonons in one (source, step) group";
                return PSIM_E_INVALID;
            }
        } else if (s.kind == PSIM_SRC_SURFACE) {
            if (s.index >= P.n_emitters) {
                err = "surface source refers to a missing emitter";
                return PSIM_E_INVALID;
            }
            if (P.phasor) { ds.kind = 2u; }
            const DevEmitter& em = img.emitters[s.index];
            if (s.count > 0 && !(em.duration > 0.)) {
                err = "surface source with phonons but an empty emission window";
                return PSIM_E_INVALID;
            }
            // the source's phonons are dealt over the measurement steps in proportion to the time each step
            // overlaps the emission window: indices [J(k), J(k+1)) are born (uniformly) inside step k, with
            // J(k) = ceil(count * (k * dt - start) / duration) clamped to [0, count]
            const double cnt = static_cast<double>(s.count);
            auto J = [&](uint32_t k) -> uint64_t {
                const double x = (static_cast<double>(k) * dt - em.start) / em.duration * cnt;
                if (!(x > 0.)) { return 0; }
                if (x >= cnt) { return s.count; }
                return static_cast<uint64_t>(std::ceil(x));
            };
            uint64_t ja = 0;
            for (uint32_t k = 0; k < M && ja < s.count; ++k) {
                const uint64_t jb = (k + 1 == M) ? s.count : std::max(ja, J(k + 1));
                if (!add(static_cast<uint32_t>(i), k, next_id, ja, jb)) {
                    err = "more than 2^32 phonons in one (source, step) group";
                    return PSIM_E_INVALID;
                }
                ja = jb;
            }
        } else {
            err = "unknown source kind";
            return PSIM_E_INVALID;
        }
        out.sources.push_back(ds);
        next_id += s.count;
    }
    out.total_phonons = next_id;
    if (next_id >= (1ull << 40)) {
        err = "more than 2^40 phonons";
        return PSIM_E_INVALID;
    }
    // by step, sources in their order within a step: a counting sort (the groups arrive source by source, each source's in
    // step order; a comparison sort of the 2e5 groups of the kinked wire - 42 surfaces x 5000 steps - took 10 ms)
    out.step_begin.assign(M + 1, 0);
    for (const DevBirth& b : births) { ++out.step_begin[b.step + 1]; }
    for (uint32_t k = 1; k <= M; ++k) { out.step_begin[k] += out.step_begin[k - 1]; }
    out.births.resize(births.size());
    {
        std::vector<uint32_t> next(out.step_begin.begin(), out.step_begin.end() - 1);
        for (const DevBirth& b : births) { out.births[next[b.step]++] = b; }
    }
    out.prefix.assign(births.size() + 1, 0);
    for (size_t i = 0; i < out.births.size(); ++i) { out.prefix[i + 1] = out.prefix[i] + out.births[i].count; }
    if (out.sources.empty()) { out.sources.push_back(DevSource{}); }
    return 0;
}

}  // namespace psim
