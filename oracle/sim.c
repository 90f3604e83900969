/* TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT.
 *
 * Plain-C restatement of the reference's phonon Monte Carlo hot path (GwGibson/Psim), one function per reference
 * function, each citing the file:line it follows.  It keeps the reference's own formulation: fp64 absolute
 * coordinates, slope/intercept segment intersection with GEOEPS tolerances, one phonon followed from birth to
 * death, independent uniform birth times.  (The CUDA path under psim_b200/ is formulated differently on purpose -
 * barycentric cell frames, fp32, step-wise pool, stratified births - and is compared with this and with the
 * reference statistically.)  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load it.
 *
 * Parity pinning: tests/test_oracle.py checks this restatement against the golden fixtures produced by the
 * unmodified reference (tests/golden/*.npz, 16 seeds per case) - so the oracle is pinned, not self-referential.
 *
 * The one deliberate difference: the reference seeds a thread_local mt19937 from std::random_device (utils.h:16-21)
 * and cannot be reproduced run to run; here every phonon owns a xoshiro256** stream seeded from (seed, phonon
 * index), so results are reproducible and independent of the OpenMP thread count.
 *
 * Build: make -C oracle oracle   ->  oracle/liboracle_sim.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NUM_FREQ_BINS 1000          /* material.h:14 */
#define GEOEPS (2.220446049250313e-16 * 1e9) /* utils.h:10 */
#define PI 3.14159265358979323846
#define SCALING_FACTOR 1e9          /* modelSimulator.cpp:17 */
#define VELOCITY_EPS 0.01           /* modelSimulator.cpp:23 */
#define MAX_COLLISIONS 100          /* modelSimulator.cpp:25 */
#define HBAR 1.054517e-34           /* material.cpp:12 */
#define BOLTZ 1.38065e-23           /* material.cpp:13 */

typedef struct { double x, y; } pt;

/* Geometry::Line (geometry.h:36-66, ctor geometry.cpp:45-53) */
typedef struct {
    pt p1, p2;
    double slope, intercept;
    double blx, bly, trx, try_;
    double length;
} line_t;

typedef struct {            /* Surface / EmitSurface / TransitionSurface (surface.h:13-106) */
    line_t line;
    double nx, ny;          /* normal_ */
    double spec;            /* specularity_ */
    int target;             /* transition: neighbour cell */
    int table;              /* emit: emit_table_ */
    double temp, start, duration;
} surf_t;

typedef struct {            /* Cell (cell.h:77-82) with its three CompositeSurfaces (compositeSurface.h:60-66) */
    pt v[3];
    int sensor;
    surf_t main[3];
    int trans_first[3], trans_count[3];
    int emit_first[3], emit_count[3];
} cell_t;

typedef struct {            /* Sensor + SensorController (sensorController.h:41-55) */
    int material;
    double t_steady;
    int base_table, scatter_table;
} sensor_t;

typedef struct {            /* Material (material.h:85-110) */
    double b_l, b_tn, b_tu, b_i, w, w_max_la, w_max_ta, freq_width;
    const double *freq, *vel_la, *vel_ta;
} material_t;

typedef struct {
    int n_cells, n_sensors, n_materials, n_subs, n_tables;
    cell_t* cells;
    surf_t* subs;
    sensor_t* sensors;
    material_t* materials;
    const double* tables;   /* [n_tables][1000][2] */
    int full_simulation, phasor_sim;
    int64_t measurement_steps, step_adjustment, recorded_steps;
    double step_time;
    const double* step_times;
} model_t;

/* Phonon (phonon.h:90-104) */
typedef struct {
    int sign;
    double lifetime;
    int64_t lifestep;
    double px, py, dx, dy;
    int64_t freq_index;
    double freq, velocity;
    int polar;              /* 0 LA, 1 TA */
    int cell;               /* -1 = left the system (cell_ == nullptr) */
} phonon_t;

/* ---------------------------------------------------------------------------------------------------- RNG */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix(uint64_t* x) {
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_seed(rng_t* r, uint64_t seed, uint64_t stream) {
    uint64_t x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    for (int i = 0; i < 4; ++i) { r->s[i] = splitmix(&x); }
}
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
/* Utils::urand (utils.h:16-21): uniform double on [0, 1] */
static double urand(rng_t* r) {
    uint64_t* s = r->s;
    const uint64_t result = rotl(s[1] * 5, 7) * 9;
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return (double)(result >> 11) * (1.0 / 9007199254740991.0);
}

/* ----------------------------------------------------------------------------------------------- geometry */
static int pt_eq(pt a, pt b) {                       /* operator==(Point), geometry.cpp:34-39 */
    const double dx = a.x - b.x, dy = a.y - b.y;
    return dx * dx + dy * dy < GEOEPS * GEOEPS;
}
static double get_slope(pt p1, pt p2) {              /* getSlope, geometry.cpp:320-322 (0 for a vertical line) */
    return (fabs(p1.x - p2.x) < GEOEPS) ? 0. : (p1.y - p2.y) / (p1.x - p2.x);
}
static line_t make_line(pt p1, pt p2) {              /* Line::Line, geometry.cpp:45-53; findBoundingBox :270-282 */
    line_t l;
    l.p1 = p1; l.p2 = p2;
    l.slope = get_slope(p1, p2);
    l.intercept = p1.y - l.slope * p1.x;
    double blx = p1.x, bly = p1.y, trx = p2.x, try_ = p2.y;
    if (p2.x <= p1.x) { double t = blx; blx = trx; trx = t; }
    if (p2.y <= p1.y) { double t = bly; bly = try_; try_ = t; }
    l.blx = blx; l.bly = bly; l.trx = trx; l.try_ = try_;
    l.length = sqrt((p2.x - p1.x) * (p2.x - p1.x) + (p2.y - p1.y) * (p2.y - p1.y));
    return l;
}
static double cross_pt(double ax, double ay, double bx, double by) { return ax * by - bx * ay; }
static int point_on_line(const line_t* l, pt p) {    /* isPointOnLine, geometry.cpp:295-298 (infinite line) */
    return fabs(cross_pt(l->p2.x - l->p1.x, l->p2.y - l->p1.y, p.x - l->p1.x, p.y - l->p1.y)) < GEOEPS;
}
static int point_right_of_line(const line_t* l, pt p) {  /* geometry.cpp:300-303 */
    return cross_pt(l->p2.x - l->p1.x, l->p2.y - l->p1.y, p.x - l->p1.x, p.y - l->p1.y) < 0.;
}
static int segment_crosses_line(const line_t* a, const line_t* b) {  /* doesSegmentCrossLine, geometry.cpp:305-308 */
    return point_on_line(a, b->p1) || point_on_line(a, b->p2) ||
           ((point_right_of_line(a, b->p1) ^ point_right_of_line(a, b->p2)) != 0);
}
static int boxes_intersect(const line_t* a, const line_t* b) {       /* boxesIntersect, geometry.cpp:284-292 */
    return a->blx <= b->trx - GEOEPS && a->trx >= b->blx + GEOEPS && a->bly <= b->try_ - GEOEPS && a->try_ >= b->bly + GEOEPS;
}
static int lines_intersect(const line_t* a, const line_t* b) {       /* Line::intersects, geometry.cpp:92-95 */
    return boxes_intersect(a, b) && segment_crosses_line(a, b) && segment_crosses_line(b, a);
}
static int approx_rel(double a, double b) {          /* the `equals` lambda, geometry.cpp:107-109 */
    return fabs(a - b) <= ((fabs(a) < fabs(b) ? fabs(b) : fabs(a)) * GEOEPS);
}
/* Line::getIntersection, geometry.cpp:104-138.  returns 1 and *out if the segments intersect at one point */
static int line_intersection(const line_t* t, const line_t* o, pt* out) {
    if (!lines_intersect(t, o)) { return 0; }
    if (approx_rel(t->p1.x, t->p2.x)) {              /* this line is vertical */
        if (approx_rel(o->p1.x, o->p2.x)) { return 0; }
        const double m = get_slope(o->p1, o->p2), b = o->p1.y - m * o->p1.x;
        out->x = t->p1.x; out->y = m * t->p1.x + b;
        return 1;
    }
    if (approx_rel(o->p1.x, o->p2.x)) {              /* the other line is vertical */
        const double m = get_slope(t->p1, t->p2), b = t->p1.y - m * t->p1.x;
        out->x = o->p1.x; out->y = m * o->p1.x + b;
        return 1;
    }
    if (approx_rel(t->slope, o->slope)) { return 0; }  /* parallel: by convention no intersection */
    const double x = (o->intercept - t->intercept) / (t->slope - o->slope);
    out->x = x; out->y = t->slope * x + t->intercept;
    return 1;
}

/* ----------------------------------------------------------------------------------------------- material */
/* Material::freqIndex, material.cpp:64-75 (returns `high`: bin 0 is never produced) */
static void freq_index(const double* table, rng_t* r, int64_t* index, int* polar) {
    size_t low = 0, high = NUM_FREQ_BINS - 1, mid = low + (high - low) / 2;
    const double rand = urand(r);
    while (high - low > 1) {
        if (rand < table[2 * mid]) { high = mid; } else { low = mid; }
        mid = (low + high) / 2;
    }
    *index = (int64_t)high;
    *polar = (urand(r) <= table[2 * high + 1]) ? 0 : 1;
}
/* Material::getFreq, material.cpp:77-80 */
static double get_freq(const model_t* m, const material_t* mat, int64_t index, rng_t* r) {
    return m->full_simulation ? mat->freq[index] : mat->freq[index] + (2. * urand(r) - 1.) * mat->freq_width / 2.;
}
/* Material::getVel, material.cpp:82-84 */
static double get_vel(const material_t* mat, int64_t index, int polar) { return polar == 0 ? mat->vel_la[index] : mat->vel_ta[index]; }
/* Material::relaxRates, material.cpp:54-57 with tauNInv :207-219, tauUInv :222-234, tauIInv :237-239 */
static void relax_rates(const material_t* mat, double temp, double freq, int polar, double out[3]) {
    double n = 0., u = 0.;
    if (polar == 0) {
        n = mat->b_l * freq * freq * pow(temp, 3);
        u = n;
    } else {
        if (freq < mat->w) { n = mat->b_tn * freq * pow(temp, 4); }
        if (freq >= mat->w) { u = mat->b_tu * freq * freq / sinh(HBAR * freq / (temp * BOLTZ)); }
    }
    out[0] = n; out[1] = u; out[2] = mat->b_i * pow(freq, 4);
}
/* SensorController::initialUpdate / scatterUpdate, sensorController.cpp:28-36,52-55,84-88 */
static void table_update(const model_t* m, phonon_t* p, const double* table, const material_t* mat, rng_t* r) {
    int64_t idx; int polar;
    freq_index(table, r, &idx, &polar);
    p->freq_index = idx;
    p->freq = get_freq(m, mat, idx, r);
    p->velocity = get_vel(mat, idx, polar);
    p->polar = polar;
}

/* ------------------------------------------------------------------------------------------------ phonon */
/* Phonon::setRandDirection, phonon.cpp:28-31 */
static void set_rand_direction(phonon_t* p, rng_t* r) {
    p->dx = 2. * urand(r) - 1.;
    p->dy = sqrt(1. - p->dx * p->dx) * cos(2. * PI * urand(r));
}
/* Surface::redirectPhonon, surface.cpp:23-30 */
static void redirect_phonon(const surf_t* s, phonon_t* p, rng_t* r) {
    const double rand = urand(r);
    const double ndx = sqrt(rand), ndy = sqrt(1. - rand) * cos(2. * PI * urand(r));
    p->dx = s->nx * ndx - s->ny * ndy;
    p->dy = s->ny * ndx + s->nx * ndy;
}
/* Surface::boundaryHandlePhonon, surface.cpp:32-44 */
static void boundary_handle(const surf_t* s, phonon_t* p, rng_t* r) {
    if (s->spec == 1. || urand(r) < s->spec) {
        const double ndx = -p->dx * s->nx - p->dy * s->ny, ndy = -p->dx * s->ny + p->dy * s->nx;
        p->dx = s->nx * ndx - s->ny * ndy;
        p->dy = s->ny * ndx + s->nx * ndy;
    } else {
        redirect_phonon(s, p, r);
    }
}
/* EmitSurface::handlePhonon, surface.cpp:61-65 */
static void emit_handle(const surf_t* s, phonon_t* p, double step_time, rng_t* r) {
    const double phonon_time = (double)p->lifestep * step_time;
    if (phonon_time < s->start || phonon_time + step_time > s->start + s->duration) {
        boundary_handle(s, p, r);
    } else {
        p->cell = -1;
    }
}
/* TransitionSurface::handlePhonon, surface.cpp:71-109 */
static void transition_handle(const model_t* m, const surf_t* s, phonon_t* p, rng_t* r) {
    const int cur_mat = m->sensors[m->cells[p->cell].sensor].material;
    const int new_mat = m->sensors[m->cells[s->target].sensor].material;
    if (cur_mat == new_mat) {
        p->cell = s->target;
    } else {
        const material_t* mat = &m->materials[new_mat];
        const double max_freq = (p->polar == 0) ? mat->w_max_la : mat->w_max_ta;
        if (p->freq > max_freq) {
            redirect_phonon(s, p, r);     /* back-scatter with THIS sub-surface's normal */
        } else {
            p->cell = s->target;
        }
    }
}
/* CompositeSurface::handlePhonon, compositeSurface.cpp:47-66: transitions first, then emitters, else the main surface */
static void composite_handle(const model_t* m, const cell_t* c, int k, phonon_t* p, pt poi, double step_time, rng_t* r) {
    for (int i = 0; i < c->trans_count[k]; ++i) {
        const surf_t* s = &m->subs[c->trans_first[k] + i];
        if (point_on_line(&s->line, poi)) { transition_handle(m, s, p, r); return; }
    }
    for (int i = 0; i < c->emit_count[k]; ++i) {
        const surf_t* s = &m->subs[c->emit_first[k] + i];
        if (point_on_line(&s->line, poi)) { emit_handle(s, p, step_time, r); return; }
    }
    boundary_handle(&c->main[k], p, r);
}
/* Cell::handleSurfaceCollision, cell.cpp:103-108: first boundary whose (infinite) line contains the point */
static void handle_surface_collision(const model_t* m, phonon_t* p, pt poi, double step_time, rng_t* r) {
    const cell_t* c = &m->cells[p->cell];
    for (int k = 0; k < 3; ++k) {
        if (point_on_line(&c->main[k].line, poi)) { composite_handle(m, c, k, p, poi, step_time, r); return; }
    }
}

/* ModelSimulator::nextImpact, modelSimulator.cpp:87-122.  returns 1 and *time_out on impact */
static int next_impact(const model_t* m, phonon_t* p, double time, double* time_out, rng_t* r) {
    const double vx = p->dx * p->velocity, vy = p->dy * p->velocity;
    const pt start = { p->px, p->py }, end = { p->px + time * vx, p->py + time * vy };
    if (pt_eq(start, end)) { return 0; }
    const line_t path = make_line(start, end);
    const cell_t* c = &m->cells[p->cell];
    int have = 0;
    pt impact = { 0., 0. };
    for (int k = 0; k < 3; ++k) {
        pt poi;
        if (line_intersection(&c->main[k].line, &path, &poi) && !pt_eq(poi, start)) {
            const double tx = (vx > VELOCITY_EPS || vx < -VELOCITY_EPS) ? (poi.x - start.x) / vx : time;
            const double ty = (vy > VELOCITY_EPS || vy < -VELOCITY_EPS) ? (poi.y - start.y) / vy : time;
            const double ti = (tx <= ty) ? tx : ty;
            if (ti <= time) { time = ti; impact = poi; have = 1; }
        }
    }
    if (have) {
        p->px = impact.x; p->py = impact.y;
        handle_surface_collision(m, p, impact, m->step_time, r);
        *time_out = time;
        return 1;
    }
    return 0;
}

/* Triangle::getRandPoint, geometry.cpp:234-242 */
static pt tri_rand_point(const cell_t* c, double r1, double r2) {
    if (r1 + r2 > 1.) { r1 = 1 - r1; r2 = 1 - r2; }
    pt q = { c->v[0].x + (c->v[1].x - c->v[0].x) * r1 + (c->v[2].x - c->v[0].x) * r2,
             c->v[0].y + (c->v[1].y - c->v[0].y) * r1 + (c->v[2].y - c->v[0].y) * r2 };
    return q;
}

/* ModelSimulator::handleImpacts, modelSimulator.cpp:205-225.  returns 0 if the phonon left the system */
static int handle_impacts(const model_t* m, phonon_t* p, double drift_time, int sensor_id, double* drifted_out, rng_t* r) {
    double impact_time = 0.;
    int hit = next_impact(m, p, drift_time, &impact_time, r);
    double drifted = 0.;
    size_t collisions = 0;
    while (hit) {
        if (p->cell < 0) { return 0; }
        drifted += impact_time;
        if (++collisions > MAX_COLLISIONS) {
            const double u1 = urand(r), u2 = urand(r);
            const pt q = tri_rand_point(&m->cells[p->cell], u1, u2);
            p->px = q.x; p->py = q.y;
            *drifted_out = drift_time;
            return 1;
        }
        if (sensor_id != m->cells[p->cell].sensor) { *drifted_out = drifted; return 1; }
        hit = next_impact(m, p, drift_time - drifted, &impact_time, r);
    }
    if (p->cell < 0) { return 0; }
    *drifted_out = drifted;
    return 1;
}

/* ModelSimulator::scatter, modelSimulator.cpp:124-137 */
static void scatter(const model_t* m, phonon_t* p, const double rates[3], rng_t* r) {
    const double tau_inv = rates[0] + rates[1] + rates[2];
    const double rand = urand(r);
    if (rand <= (rates[0] + rates[1]) / tau_inv) {
        const sensor_t* s = &m->sensors[m->cells[p->cell].sensor];
        table_update(m, p, m->tables + (size_t)s->scatter_table * 2 * NUM_FREQ_BINS, &m->materials[s->material], r);
        if (rand > rates[0] / tau_inv) { set_rand_direction(p, r); }
    } else if (rates[2] > 0.) {
        set_rand_direction(p, r);
    }
}

typedef struct { int64_t* energy; double* flux; int64_t drift_steps, loop_iters; } tally_t;

/* Sensor::updateHeatParams, sensor.cpp:43-52 (per-thread arrays here instead of a mutex) */
static void update_heat_params(const model_t* m, const phonon_t* p, int64_t step, tally_t* t) {
    const size_t k = (size_t)m->cells[p->cell].sensor * (size_t)m->recorded_steps + (size_t)step;
    t->energy[k] += p->sign;
    t->flux[2 * k] += p->dx * p->velocity * p->sign;
    t->flux[2 * k + 1] += p->dy * p->velocity * p->sign;
}

/* ModelSimulator::simulatePhonon, modelSimulator.cpp:139-198 */
static void simulate_phonon(const model_t* m, phonon_t p, rng_t* r, tally_t* t) {
    const int64_t M = m->measurement_steps;
    int alive = 1;
    double age = p.lifetime;
    int64_t step = (int64_t)(age / m->step_time);
    p.lifestep = step;
    double rates[3] = { 0., 0., 0. };
    double tts = 0., ttm = 0.;
    int counted_interval = 0;
    while (alive) {
        ++t->loop_iters;
        if (tts <= 0.) {   /* get_scatter_info, modelSimulator.cpp:148-153 */
            const sensor_t* s = &m->sensors[m->cells[p.cell].sensor];
            relax_rates(&m->materials[s->material], s->t_steady, p.freq, p.polar, rates);
            tts = SCALING_FACTOR * -log(urand(r)) / (rates[0] + rates[1] + rates[2]);
        }
        if (ttm <= 0.) { ttm = m->step_times[step] - age; counted_interval = 0; }
        if (!counted_interval) { ++t->drift_steps; counted_interval = 1; }  /* one drift-step per (phonon, interval) */
        double drift_time = (tts < ttm) ? tts : ttm;
        const int sensor_id = m->cells[p.cell].sensor;
        double drifted = 0.;
        if (handle_impacts(m, &p, drift_time, sensor_id, &drifted, r)) {
            if (m->cells[p.cell].sensor != sensor_id) { drift_time = drifted; }
            const double f = p.velocity * (drift_time - drifted);   /* Phonon::drift, phonon.cpp:22-26 */
            p.px += p.dx * f; p.py += p.dy * f;
            age += drift_time; ttm -= drift_time; tts -= drift_time;
            if (ttm == 0.) {
                if (++step < M) {
                    p.lifestep = step;
                    if (step >= m->step_adjustment) { update_heat_params(m, &p, step - m->step_adjustment, t); }
                } else {
                    alive = 0;
                }
            } else if (!m->phasor_sim && tts == 0.) {
                scatter(m, &p, rates, r);
            } else {
                tts = 0.;
            }
        } else {
            alive = 0;
        }
    }
}

/* sources: one row per phonon builder (phononBuilder.cpp:6-49) */
typedef struct { int kind; int cell; int sub; int sign; int64_t count; } source_t;

static void build_phonon(const model_t* m, const source_t* s, rng_t* r, phonon_t* p) {
    memset(p, 0, sizeof(*p));
    p->sign = s->sign;
    p->cell = s->cell;
    const cell_t* c = &m->cells[s->cell];
    const sensor_t* sen = &m->sensors[c->sensor];
    const material_t* mat = &m->materials[sen->material];
    if (s->kind == 0) {   /* CellOriginBuilder::operator(), phononBuilder.cpp:6-16 */
        p->lifetime = 0.;
        table_update(m, p, m->tables + (size_t)sen->base_table * 2 * NUM_FREQ_BINS, mat, r);
        const double u1 = urand(r), u2 = urand(r);
        const pt q = tri_rand_point(c, u1, u2);
        p->px = q.x; p->py = q.y;
        set_rand_direction(p, r);
    } else {              /* SurfaceOriginBuilder::operator(), phononBuilder.cpp:31-40 */
        const surf_t* es = &m->subs[s->sub];
        p->lifetime = es->start + es->duration * urand(r);       /* EmitSurface::getPhononTime, surface.cpp:67-69 */
        table_update(m, p, m->tables + (size_t)es->table * 2 * NUM_FREQ_BINS, mat, r);
        const double r1 = urand(r), r2 = 1. - r1;                /* Line::getRandPoint, geometry.cpp:140-143 */
        p->px = es->line.p1.x * r1 + es->line.p2.x * r2;
        p->py = es->line.p1.y * r1 + es->line.p2.y * r2;
        redirect_phonon(es, p, r);
        if (m->phasor_sim) {  /* PhasorBuilder::operator(), phononBuilder.cpp:42-49 */
            p->freq_index = 1; p->freq = 1.; p->velocity = 1000.; p->polar = 0;
            p->dx = es->nx; p->dy = es->ny;
        }
    }
}

/* Test hook: the reference's geometry for ONE free flight in a wall-only triangle - nextImpact (modelSimulator.cpp:
 * 87-122) with Line::getIntersection (geometry.cpp:104-138), then the specular branch of boundaryHandlePhonon
 * (surface.cpp:32-44) about the edge's inward normal (Line::normal, geometry.cpp:97-100).
 * tri[6] = x1 y1 x2 y2 x3 y3; state = px py vx vy (nm, nm/ns); out = edge, time, x, y, dx', dy' (unit input direction). */
int oracle_flight(const double* tri, const double* state, double horizon, double* out) {
    cell_t c;
    memset(&c, 0, sizeof(c));
    for (int k = 0; k < 3; ++k) { c.v[k].x = tri[2 * k]; c.v[k].y = tri[2 * k + 1]; }
    const double cw = (c.v[1].x - c.v[0].x) * (c.v[1].y + c.v[0].y) + (c.v[2].x - c.v[1].x) * (c.v[2].y + c.v[1].y) +
                      (c.v[0].x - c.v[2].x) * (c.v[0].y + c.v[2].y);   /* Triangle::isClockwise, geometry.cpp:217-222 */
    const int ns = cw >= 0. ? 1 : -1;
    for (int k = 0; k < 3; ++k) {
        surf_t* s = &c.main[k];
        s->line = make_line(c.v[k], c.v[(k + 1) % 3]);
        s->nx = ns * (s->line.p2.y - s->line.p1.y) / s->line.length;
        s->ny = -ns * (s->line.p2.x - s->line.p1.x) / s->line.length;
        s->spec = 1.;
    }
    const double speed = sqrt(state[2] * state[2] + state[3] * state[3]);
    const pt start = { state[0], state[1] }, end = { state[0] + horizon * state[2], state[1] + horizon * state[3] };
    const line_t path = make_line(start, end);
    double time = horizon;
    int edge = -1;
    pt impact = { 0., 0. };
    for (int k = 0; k < 3; ++k) {
        pt poi;
        if (line_intersection(&c.main[k].line, &path, &poi) && !pt_eq(poi, start)) {
            const double tx = (state[2] > VELOCITY_EPS || state[2] < -VELOCITY_EPS) ? (poi.x - start.x) / state[2] : time;
            const double ty = (state[3] > VELOCITY_EPS || state[3] < -VELOCITY_EPS) ? (poi.y - start.y) / state[3] : time;
            const double ti = (tx <= ty) ? tx : ty;
            if (ti <= time) { time = ti; impact = poi; edge = k; }
        }
    }
    out[0] = edge; out[1] = time; out[2] = impact.x; out[3] = impact.y;
    double dx = state[2] / speed, dy = state[3] / speed;
    if (edge >= 0) {
        const surf_t* s = &c.main[edge];
        const double ndx = -dx * s->nx - dy * s->ny, ndy = -dx * s->ny + dy * s->nx;
        dx = s->nx * ndx - s->ny * ndy;
        dy = s->ny * ndx + s->nx * ndy;
    }
    out[4] = dx; out[5] = dy;
    return 0;
}

/* Entry point.  Flat arrays in, tallies out; everything is copied into the structs above first.
 *   cell_xy[C][6], cell_sensor[C], cell_spec[C], cell_norm_sign[C]
 *   sub_*[n_subs]: kind (1 transition, 2 emit), owner cell, owner edge, target cell, x1 y1 x2 y2, nx ny, table, temp, start, duration
 *                  sorted by (cell, edge) with transitions before emitters, insertion order kept
 *   sensor_*[S], mat_consts[K][8], mat_arrays[K][3][1000] (freq, vel_la, vel_ta), tables[T][1000][2]
 *   sources[n_sources][5] as int64: kind, cell, sub, sign, count
 * Output: energy[S][R] int64, flux[S][R][2] double, counters[2] = drift steps, loop iterations. */
int oracle_run(int n_cells, const double* cell_xy, const int* cell_sensor, const double* cell_spec, const int* cell_norm_sign,
               int n_subs, const int* sub_kind, const int* sub_cell, const int* sub_edge, const int* sub_target,
               const double* sub_line, const double* sub_normal, const int* sub_table, const double* sub_window /* temp,start,duration */,
               int n_sensors, const int* sensor_material, const double* sensor_temp, const int* sensor_base, const int* sensor_scatter,
               int n_materials, const double* mat_consts, const double* mat_arrays, int n_tables, const double* tables,
               int64_t measurement_steps, int64_t step_adjustment, double simulation_time, int full_simulation, int phasor_sim,
               int n_sources, const int64_t* sources, uint64_t seed, int threads,
               int64_t* energy_out, double* flux_out, int64_t* counters) {
    model_t m;
    memset(&m, 0, sizeof(m));
    m.n_cells = n_cells; m.n_sensors = n_sensors; m.n_materials = n_materials; m.n_subs = n_subs; m.n_tables = n_tables;
    m.full_simulation = full_simulation; m.phasor_sim = phasor_sim;
    m.measurement_steps = measurement_steps; m.step_adjustment = step_adjustment;
    m.recorded_steps = measurement_steps - step_adjustment;
    m.step_time = simulation_time / (double)measurement_steps;   /* modelSimulator.cpp:30 */
    double* st = (double*)malloc(sizeof(double) * (size_t)measurement_steps);
    for (int64_t n = 1; n <= measurement_steps; ++n) { st[n - 1] = (double)n * simulation_time / (double)measurement_steps; } /* :33-36 */
    m.step_times = st;
    m.tables = tables;
    m.materials = (material_t*)calloc((size_t)n_materials, sizeof(material_t));
    for (int k = 0; k < n_materials; ++k) {
        const double* c = mat_consts + 8 * k;
        material_t* mt = &m.materials[k];
        mt->b_l = c[0]; mt->b_tn = c[1]; mt->b_tu = c[2]; mt->b_i = c[3]; mt->w = c[4]; mt->w_max_la = c[5]; mt->w_max_ta = c[6];
        mt->freq_width = c[7];
        mt->freq = mat_arrays + (size_t)k * 3 * NUM_FREQ_BINS;
        mt->vel_la = mt->freq + NUM_FREQ_BINS;
        mt->vel_ta = mt->freq + 2 * NUM_FREQ_BINS;
    }
    m.sensors = (sensor_t*)calloc((size_t)n_sensors, sizeof(sensor_t));
    for (int s = 0; s < n_sensors; ++s) {
        m.sensors[s].material = sensor_material[s]; m.sensors[s].t_steady = sensor_temp[s];
        m.sensors[s].base_table = sensor_base[s]; m.sensors[s].scatter_table = sensor_scatter[s];
    }
    m.cells = (cell_t*)calloc((size_t)n_cells, sizeof(cell_t));
    for (int c = 0; c < n_cells; ++c) {
        cell_t* cl = &m.cells[c];
        for (int k = 0; k < 3; ++k) { cl->v[k].x = cell_xy[6 * c + 2 * k]; cl->v[k].y = cell_xy[6 * c + 2 * k + 1]; }
        cl->sensor = cell_sensor[c];
        double spec = cell_spec[c];                 /* Cell::buildCompositeSurfaces, cell.cpp:114-124 */
        if (spec < 0.) { spec = 0.; } else if (spec > 1.) { spec = 1.; }
        const int ns = cell_norm_sign[c] >= 0 ? 1 : -1;
        for (int k = 0; k < 3; ++k) {
            surf_t* s = &cl->main[k];
            s->line = make_line(cl->v[k], cl->v[(k + 1) % 3]);
            s->nx = ns * (s->line.p2.y - s->line.p1.y) / s->line.length;     /* Line::normal, geometry.cpp:97-100 */
            s->ny = -ns * (s->line.p2.x - s->line.p1.x) / s->line.length;
            s->spec = spec;
        }
    }
    m.subs = (surf_t*)calloc((size_t)(n_subs > 0 ? n_subs : 1), sizeof(surf_t));
    for (int i = 0; i < n_subs; ++i) {
        surf_t* s = &m.subs[i];
        const pt a = { sub_line[4 * i], sub_line[4 * i + 1] }, b = { sub_line[4 * i + 2], sub_line[4 * i + 3] };
        s->line = make_line(a, b);
        s->nx = sub_normal[2 * i]; s->ny = sub_normal[2 * i + 1];
        s->target = sub_target[i]; s->table = sub_table[i];
        s->temp = sub_window[3 * i]; s->start = sub_window[3 * i + 1]; s->duration = sub_window[3 * i + 2];
        cell_t* cl = &m.cells[sub_cell[i]];
        const int k = sub_edge[i];
        if (sub_kind[i] == 1) {
            s->spec = 0.;                            /* compositeSurface.cpp:41 */
            if (cl->trans_count[k]++ == 0) { cl->trans_first[k] = i; }
        } else {
            s->spec = cl->main[k].spec;              /* compositeSurface.cpp:28-31 */
            if (cl->emit_count[k]++ == 0) { cl->emit_first[k] = i; }
        }
    }
    const size_t n_tally = (size_t)n_sensors * (size_t)m.recorded_steps;
    memset(energy_out, 0, sizeof(int64_t) * n_tally);
    memset(flux_out, 0, sizeof(double) * 2 * n_tally);
    int64_t total = 0;
    int64_t* first = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_sources + 1));
    for (int i = 0; i < n_sources; ++i) { first[i] = total; total += sources[5 * i + 4]; }
    first[n_sources] = total;
#ifdef _OPENMP
    if (threads > 0) { omp_set_num_threads(threads); }
#endif
    int64_t drift_steps = 0, loop_iters = 0;
#pragma omp parallel reduction(+ : drift_steps, loop_iters)
    {
        tally_t t;
        t.energy = (int64_t*)calloc(n_tally, sizeof(int64_t));
        t.flux = (double*)calloc(2 * n_tally, sizeof(double));
        t.drift_steps = 0; t.loop_iters = 0;
#pragma omp for schedule(dynamic, 256)
        for (int64_t id = 0; id < total; ++id) {
            int lo = 0, hi = n_sources;            /* which builder does phonon `id` belong to */
            while (hi - lo > 1) { const int mid = (lo + hi) / 2; if (id < first[mid]) { hi = mid; } else { lo = mid; } }
            source_t s = { (int)sources[5 * lo], (int)sources[5 * lo + 1], (int)sources[5 * lo + 2], (int)sources[5 * lo + 3], sources[5 * lo + 4] };
            rng_t r;
            rng_seed(&r, seed, (uint64_t)id);
            phonon_t p;
            build_phonon(&m, &s, &r, &p);
            simulate_phonon(&m, p, &r, &t);
        }
#pragma omp critical
        {
            for (size_t k = 0; k < n_tally; ++k) { energy_out[k] += t.energy[k]; flux_out[2 * k] += t.flux[2 * k]; flux_out[2 * k + 1] += t.flux[2 * k + 1]; }
        }
        drift_steps += t.drift_steps; loop_iters += t.loop_iters;
        free(t.energy); free(t.flux);
    }
    counters[0] = drift_steps; counters[1] = loop_iters;
    free(first); free(st); free(m.materials); free(m.sensors); free(m.cells); free(m.subs);
    return 0;
}
