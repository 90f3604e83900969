// Per-phonon physics of the hot path, written once and used by the sm_100a kernels (kernels.cu).
// (tests/emu compiles the same header with g++ so that the algorithm can be checked statistically against the
// reference's golden data in the GPU-less authoring container; that build is test-only and never shipped.)
//
// What each function replaces in the reference is cited next to it.  The structure is NOT the reference's:
// the reference follows one phonon object through its whole life (modelSimulator.cpp:139-198); here a phonon
// is advanced across ONE measurement interval at a time from a 32-byte state record, so that the pool can be
// streamed from HBM step by step, and every random draw comes from a Philox4x32-10 stream addressed by
// (seed, phonon id, measurement step, draw index) - a phonon's trajectory is therefore a pure function of
// its global id and is identical whichever GPU, warp or lane runs it.
#ifndef PSIM_B200_DEVICE_CORE_CUH
#define PSIM_B200_DEVICE_CORE_CUH

#include "device_types.h"

#if defined(__CUDACC__)
#define PSIM_HD __device__ __forceinline__
#else
#include <cmath>
#define PSIM_HD inline
#endif

// what the flight loop's fast path (fast_impact) takes on besides whole-edge transitions
#ifndef PSIM_FAST_WALLS
#define PSIM_FAST_WALLS 0
#endif
#ifndef PSIM_FAST_COMPOSITE
#define PSIM_FAST_COMPOSITE 1
#endif

namespace psim {

#if defined(__CUDA_ARCH__)
PSIM_HD uint32_t mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
PSIM_HD float f_log(float x) { return __logf(x); }
PSIM_HD float f_exp(float x) { return __expf(x); }
PSIM_HD float f_cos2pi(float u) { return __cosf(6.283185307179586f * u); }  // MUFU.COS, abs. error < 1e-6 on [0, 2 pi]
PSIM_HD float f_sqrt(float x) {                                              // MUFU.SQRT (x >= 0)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
PSIM_HD float f_div(float a, float b) {  // a / b with one MUFU.RCP (1 ulp) - ample for hit times and scatter times
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
    return a * r;
}
PSIM_HD float f_inf() { return __int_as_float(0x7f800000); }
// a b + c d with ONE association everywhere: fma(a, b, round(c d)).  Left to the compiler, which of the two products is
// fused depends on the code around the expression, and the kernel variants - which must agree bit for bit - got different
// roundings on cells whose frame is not axis-aligned (kinked wire: a few flight segments per million went another way).
PSIM_HD float dot2(float a, float b, float c, float d) { return __fmaf_rn(a, b, __fmul_rn(c, d)); }
PSIM_HD float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template<typename T> PSIM_HD T ldg(const T* p) { return __ldg(p); }
PSIM_HD uint4 load_cell_links(const DevCell* cells, uint32_t i) { return __ldg(reinterpret_cast<const uint4*>(cells + i)); }
PSIM_HD uint2 load_cell_tail(const DevCell* cells, uint32_t i) { return __ldg(reinterpret_cast<const uint2*>(&cells[i].sensor_mat)); }
PSIM_HD uint2 load_cell_tris(const DevCell* cells, uint32_t i) { return __ldg(reinterpret_cast<const uint2*>(cells[i].tri)); }
PSIM_HD float4 load_shape_matrix(const DevShape* shapes, uint32_t i) { return __ldg(reinterpret_cast<const float4*>(shapes + i)); }
PSIM_HD float2 load_shape_normal(const DevShape* shapes, uint32_t i, uint32_t e) {
    return __ldg(reinterpret_cast<const float2*>(shapes[i].n) + e);
}
PSIM_HD DevSensor load_sensor(const DevSensor* sensors, uint32_t i) {
    const float4* q = reinterpret_cast<const float4*>(sensors + i);
    union { float4 v[2]; DevSensor s; } u;
    u.v[0] = __ldg(q);
    u.v[1] = __ldg(q + 1);
    return u.s;
}
#else
PSIM_HD uint32_t mulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
PSIM_HD float f_log(float x) { return std::log(x); }
PSIM_HD float f_exp(float x) { return std::exp(x); }
PSIM_HD float f_cos2pi(float u) { return std::cos(6.283185307179586f * u); }
PSIM_HD float f_sqrt(float x) { return std::sqrt(x); }
PSIM_HD float f_div(float a, float b) { return a / b; }
PSIM_HD float f_inf() { return INFINITY; }
PSIM_HD float dot2(float a, float b, float c, float d) { return std::fmaf(a, b, c * d); }  // (built with -ffp-contract=off)
PSIM_HD float fma_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
template<typename T> PSIM_HD T ldg(const T* p) { return *p; }
PSIM_HD uint4 load_cell_links(const DevCell* cells, uint32_t i) {
    uint4 q;
    q.x = cells[i].link[0], q.y = cells[i].link[1], q.z = cells[i].link[2], q.w = cells[i].link[3];
    return q;
}
PSIM_HD uint2 load_cell_tail(const DevCell* cells, uint32_t i) {
    uint2 q;
    q.x = cells[i].sensor_mat, q.y = cells[i].shape;
    return q;
}
PSIM_HD uint2 load_cell_tris(const DevCell* cells, uint32_t i) {
    uint2 q;
    q.x = cells[i].tri[0], q.y = cells[i].tri[1];
    return q;
}
PSIM_HD float4 load_shape_matrix(const DevShape* shapes, uint32_t i) {
    float4 m;
    m.x = shapes[i].m00, m.y = shapes[i].m01, m.z = shapes[i].m10, m.w = shapes[i].m11;
    return m;
}
PSIM_HD float2 load_shape_normal(const DevShape* shapes, uint32_t i, uint32_t e) {
    float2 n;
    n.x = shapes[i].n[e][0], n.y = shapes[i].n[e][1];
    return n;
}
PSIM_HD uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
PSIM_HD uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
PSIM_HD DevSensor load_sensor(const DevSensor* sensors, uint32_t i) { return sensors[i]; }
#endif

// A phonon's cell word is TAGGED (PSIM_CELL_INDEX / PSIM_CELL_QUAD); every loader takes the tagged word.
PSIM_HD uint32_t link_of_edge(const uint4 links, uint32_t e) { return (e == 0u) ? links.x : ((e == 1u) ? links.y : ((e == 2u) ? links.z : links.w)); }
PSIM_HD uint32_t cell_sensor_word(const DevParams& P, uint32_t cell) { return ldg(&P.cells[PSIM_CELL_INDEX(cell)].sensor_mat); }
// geometry of a cell = its shape record (device_types.h: DevShape)
PSIM_HD float4 load_cell_matrix(const DevParams& P, uint32_t cell) { return load_shape_matrix(P.shapes, ldg(&P.cells[PSIM_CELL_INDEX(cell)].shape)); }
PSIM_HD float2 load_cell_normal(const DevParams& P, uint32_t cell, uint32_t e) {
    return load_shape_normal(P.shapes, ldg(&P.cells[PSIM_CELL_INDEX(cell)].shape), e);
}
PSIM_HD float load_cell_spec(const DevParams& P, uint32_t cell) { return ldg(&P.shapes[ldg(&P.cells[PSIM_CELL_INDEX(cell)].shape)].spec); }
// the model-file cell a phonon at (b1, b2) of flight cell `cell` is in: a parallelogram holds one triangle on either side of
// its diagonal b1 = b2 (device_types.h: DevCell::tri)
PSIM_HD uint32_t api_cell_of(const DevParams& P, uint32_t cell, float b1, float b2) {
    const uint2 t = load_cell_tris(P.cells, PSIM_CELL_INDEX(cell));
    return (b1 >= b2) ? t.x : t.y;
}
PSIM_HD float clamp01(float x) {
#if defined(__CUDA_ARCH__)
    return __saturatef(x);
#else
    return fminf(fmaxf(x, 0.f), 1.f);
#endif
}

// Lattice image (device_types.h): column / row of the parallelogram that holds coordinate x in [0, 1] of an n-wide lattice
// axis, and the coordinate inside it.  A point on a lattice line belongs to the parallelogram on either side: it lies on the
// edge of both, and a flight that leaves through that edge at once costs one zero-length segment.
PSIM_HD uint32_t lattice_split(float x, uint32_t n, float& local) {
    const float y = x * static_cast<float>(n);
    const uint32_t i = min(static_cast<uint32_t>(fmaxf(y, 0.f)), n - 1u);
    local = clamp01(y - static_cast<float>(i));
    return i;
}
// a phonon in lattice coordinates -> the same phonon in its fine flight cell (`cells`, `sub_fine`: of the lattice image)
PSIM_HD void coarse_to_fine(const DevCell* cells, const uint32_t* sub_fine, uint32_t& cell, float& b1, float& b2) {
    const uint2 t = load_cell_tris(cells, PSIM_CELL_INDEX(cell));  // (first sub-cell, nx | ny << 16)
    const uint32_t nx = t.y & 0xFFFFu, ny = t.y >> 16;
    float l1, l2;
    const uint32_t ix = lattice_split(b1, nx, l1), iy = lattice_split(b2, ny, l2);
    cell = ldg(&sub_fine[t.x + iy * nx + ix]);
    b1 = l1;
    b2 = l2;
}

// relaxation-rate record of the sensor area a cell word names, at measurement step `step`: the record of its rate class
// where it has one (a handful of records for the whole mesh), the sensor's own otherwise; a transient run that re-iterates
// has one record per (sensor, step) (TransientController::getSteadyTemp / scatterUpdate, sensorController.cpp:80-88)
PSIM_HD DevSensor load_rates(const DevParams& P, uint32_t cell_word, uint32_t step) {
    const uint32_t cls = PSIM_CELL_CLASS(cell_word), sensor = PSIM_CELL_SENSOR(cell_word);
    const DevSensor* rec = (cls != 255u) ? P.classes + cls : P.sensors + sensor;
    if (P.step_sensors != nullptr) { rec = P.step_sensors + static_cast<size_t>(sensor) * P.num_steps + min(step, P.num_steps - 1u); }
    return load_sensor(rec, 0u);  // one load path, the pointer is selected
}

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Replaces the reference's thread_local mt19937 seeded from
// std::random_device (utils.h:16-21).  key = seed, counter = (block, step, id_lo, id_hi): the stream of a phonon
// in a measurement step is addressed by its global id and the step, so it does not matter which lane produces it.
// A refill yields four 32-bit words; code that consumes randoms asks for what it needs at ONE place per code
// block (rng_need) so that the 10-round function is instantiated a handful of times, not once per draw.
// ---------------------------------------------------------------------------------------------------------
struct Rng {
    uint32_t block;  // next 4-word block of this (id, step) stream
    uint32_t left;   // unread words among v0..v3
    uint32_t v0, v1, v2, v3;
};

PSIM_HD void rng_begin(Rng& r) {
    r.block = 0;
    r.left = 0;
    r.v0 = r.v1 = r.v2 = r.v3 = 0;
}

PSIM_HD void rng_refill(Rng& r, const DevParams& P, uint32_t step, uint32_t id_lo, uint32_t id_hi) {
    uint32_t c0 = r.block++, c1 = step, c2 = id_lo, c3 = id_hi;
    uint32_t a = P.seed_lo, b = P.seed_hi;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < PSIM_PHILOX_ROUNDS; ++i) {
        const uint32_t hi0 = mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ a;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ b;
        c3 = lo0;
        a += 0x9E3779B9u;
        b += 0xBB67AE85u;
    }
    r.v0 = c0;
    r.v1 = c1;
    r.v2 = c2;
    r.v3 = c3;
    r.left = 4;
}

// at least n (<= 4) unread words; leftovers of the previous block are dropped
PSIM_HD void rng_need(Rng& r, uint32_t n, const DevParams& P, uint32_t step, uint32_t id_lo, uint32_t id_hi) {
    if (r.left < n) { rng_refill(r, P, step, id_lo, id_hi); }
}

// next unread 32-bit word; needs left > 0
PSIM_HD uint32_t rng_word(Rng& r) {
    const uint32_t x = r.v0;
    r.v0 = r.v1;
    r.v1 = r.v2;
    r.v2 = r.v3;
    --r.left;
    return x;
}

// uniform on (0, 1] with 24 random bits (the reference draws doubles on [0, 1], utils.h:19)
PSIM_HD float u01_from24(uint32_t x) { return static_cast<float>((x >> 8) + 1u) * 5.9604644775390625e-8f; }
// two uniforms on (0, 1] with 16 random bits each from one word (directions, branch selection)
PSIM_HD float u01_hi16(uint32_t x) { return static_cast<float>((x >> 16) + 1u) * 1.52587890625e-5f; }
PSIM_HD float u01_lo16(uint32_t x) { return static_cast<float>((x & 0xFFFFu) + 1u) * 1.52587890625e-5f; }
PSIM_HD float rng_u01(Rng& r) { return u01_from24(rng_word(r)); }

struct Phonon {
    float b1, b2;     // position in the current cell's frame
    float dx, dy;     // in-plane velocity (m/s) = group velocity * direction, |direction| <= 1   (phonon.cpp:28-31)
    float tts;        // time to the next intrinsic scatter (ns); carried from interval to interval like the
                      // reference's time_to_scatter (modelSimulator.cpp:145,155,180)
    uint32_t packed;  // see device_types.h
    uint32_t cell;
    uint32_t id_lo;
};

// ---------------------------------------------------------------------------------------------------------
// Frequency / polarisation sampling  (Material::freqIndex material.cpp:64-75, getFreq :77-80, getVel :82-84,
// called from SensorController::initialUpdate / scatterUpdate, sensorController.cpp:28-36,52-55).
// Same inverse-CDF bisection as the reference, hence the same quirk: it returns `high`, i.e. bin 0 is never
// produced.  bisect_table is the reference's loop verbatim in structure; sample_bin brackets the answer with a
// PSIM_GUIDE-entry guide first (flatten.cpp builds it with bisect_table), then bisects inside the bracket: for a
// non-decreasing table both return the unique h with cdf[h-1] <= r < cdf[h], so the results are identical.
// ---------------------------------------------------------------------------------------------------------
PSIM_HD uint32_t bisect_range(const float2* table, float r, uint32_t lo, uint32_t hi) {
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (r < ldg(&table[mid].x)) {
            hi = mid;
        } else {
            lo = mid;
        }
    }
    return hi;
}

PSIM_HD uint32_t bisect_table(const float2* table, float r) { return bisect_range(table, r, 0u, PSIM_BINS - 1u); }

PSIM_HD uint32_t sample_bin(const DevParams& P, uint32_t table_idx, float r) {
    const float2* table = P.tables + static_cast<size_t>(table_idx) * PSIM_BINS;
    const uint32_t k = min(static_cast<uint32_t>(r * static_cast<float>(PSIM_GUIDE)), static_cast<uint32_t>(PSIM_GUIDE - 1));
    const uint32_t g = ldg(&P.guides[static_cast<size_t>(table_idx) * PSIM_GUIDE + k]);  // [15:0] low bracket, [31:16] high bracket
    return bisect_range(table, r, g & 0xFFFFu, g >> 16);
}

// Angular frequency (* 1e-13) of a phonon: bin centre, plus in deviational mode the jitter inside the bin
// (Material::getFreq, material.cpp:77-80).  The jitter is kept as 8 bits in the packed word, i.e. the
// reference's uniform draw over the bin is resolved to 1/256 of a bin (0.001 % of the spectrum).
PSIM_HD float phonon_omega(const DevParams& P, uint32_t packed) {
    const float fw = ldg(&P.materials[PSIM_PACK_MAT(packed)].freq_width);
    // bin + (jitter + 1/2) / 256 bins in deviational mode, bin + 1/2 otherwise: one exact integer -> float conversion
    const uint32_t fine = P.full_mode ? (PSIM_PACK_BIN(packed) << 9) + 256u : (PSIM_PACK_BIN(packed) << 9) + (PSIM_PACK_JIT(packed) << 1) + 1u;
    const float w = static_cast<float>(fine) * (1.f / 512.f) * fw;
    return P.phasor ? static_cast<float>(PSIM_FREQ_SCALE) : w;  // PhasorBuilder: freq = 1 rad/s (phononBuilder.cpp:46)
}

// bin and polarisation from two uniforms; `jit` = position of the frequency inside the bin in 1/256ths
PSIM_HD void sample_table(const DevParams& P, uint32_t table_idx, uint32_t mat, float u_bin, float u_pol, uint32_t jit,
                          Phonon& p, float& vel) {
    const uint32_t bin = sample_bin(P, table_idx, u_bin);
    const uint32_t ta = (u_pol <= ldg(&P.tables[static_cast<size_t>(table_idx) * PSIM_BINS + bin].y)) ? 0u : 1u;
    p.packed = (p.packed & 0xFF000800u) | bin | (ta << 10) | (mat << 12) | (jit << 16);
    vel = ldg(&P.velocities[(mat * 2u + ta) * PSIM_BINS + bin]);
}

PSIM_HD float phonon_velocity(const DevParams& P, uint32_t packed) {
    if (P.phasor) { return 1000.f; }  // PhasorBuilder, phononBuilder.cpp:46
    return ldg(&P.velocities[(PSIM_PACK_MAT(packed) * 2u + PSIM_PACK_TA(packed)) * PSIM_BINS + PSIM_PACK_BIN(packed)]);
}

// Relaxation rates [N, U, I] in 1/ns (Material::relaxRates material.cpp:54-57, tauNInv :207-219,
// tauUInv :222-234, tauIInv :237-239) with the temperature powers folded into the sensor record.
PSIM_HD void relax_rates(const DevSensor& s, float w, uint32_t ta, float& rn, float& ru, float& ri) {
    const float w2 = w * w;
    if (!ta) {
        rn = ru = s.c_la * w2;
    } else if (w < s.w_cut) {
        rn = s.c_tn * w;
        ru = 0.f;
    } else {
        rn = 0.f;
        const float x = s.x_t * w;
        ru = f_div(2.f * s.c_tu * w2, f_exp(x) - f_exp(-x));
    }
    ri = s.c_i * w2 * w2;
}

// Phonon::setRandDirection (phonon.cpp:28-31)
PSIM_HD void isotropic_direction(float u1, float u2, float vel, Phonon& p) {
    const float dx = 2.f * u1 - 1.f;
    p.dx = vel * dx;
    p.dy = vel * (f_sqrt(fmaxf(1.f - dx * dx, 0.f)) * f_cos2pi(u2));
}

// Surface::redirectPhonon (surface.cpp:23-30): cosine-law direction about the inward normal n
PSIM_HD void diffuse_direction(float u1, float u2, float nx, float ny, float vel, Phonon& p) {
    const float a = f_sqrt(u1);
    const float b = f_sqrt(fmaxf(1.f - u1, 0.f)) * f_cos2pi(u2);
    p.dx = vel * dot2(nx, a, -ny, b);
    p.dy = vel * dot2(ny, a, nx, b);
}

// Surface::boundaryHandlePhonon (surface.cpp:32-44); needs up to 3 unread random words
PSIM_HD void boundary_reflect(Rng& rng, float spec, float nx, float ny, float vel, Phonon& p) {
    if (spec >= 1.f || rng_u01(rng) < spec) {
        const float dn = dot2(p.dx, nx, p.dy, ny);
        p.dx = fma_rn(-2.f * dn, nx, p.dx);
        p.dy = fma_rn(-2.f * dn, ny, p.dy);
    } else {
        const float u1 = rng_u01(rng), u2 = rng_u01(rng);
        diffuse_direction(u1, u2, nx, ny, vel, p);
    }
}

// position on edge `e` of a triangle / parallelogram at fraction s from the edge's first vertex (device_types.h)
PSIM_HD void place_on_edge(uint32_t quad, uint32_t e, float s, Phonon& p) {
    const float r = 1.f - s;
    if (quad) {  // 0: b2 = 0 (s = b1)   1: b1 = 1 (s = b2)   2: b2 = 1 (s = 1 - b1)   3: b1 = 0 (s = 1 - b2)
        p.b1 = (e == 0u) ? s : ((e == 1u) ? 1.f : ((e == 2u) ? r : 0.f));
        p.b2 = (e == 0u) ? 0.f : ((e == 1u) ? s : ((e == 2u) ? 1.f : r));
    } else {     // 0: b2 = 0 (s = b1)   1: b1 + b2 = 1 (s = b2)   2: b1 = 0 (s = 1 - b2)
        p.b1 = (e == 0u) ? s : ((e == 1u) ? r : 0.f);
        p.b2 = (e == 0u) ? 0.f : ((e == 1u) ? s : r);
    }
}

// ... and back: where on edge `e` the point (b1, b2) of that edge lies
PSIM_HD float edge_coordinate(uint32_t quad, uint32_t e, float b1, float b2) {
    if (quad) { return (e == 0u) ? b1 : ((e == 1u) ? b2 : ((e == 2u) ? 1.f - b1 : 1.f - b2)); }
    return (e == 0u) ? b1 : ((e == 1u) ? b2 : 1.f - b2);
}

// ---------------------------------------------------------------------------------------------------------
// Emission (CellOriginBuilder::operator() phononBuilder.cpp:6-16, SurfaceOriginBuilder :31-40,
// PhasorBuilder :42-49, EmitSurface::getPhononTime surface.cpp:67-69, Triangle::getRandPoint
// geometry.cpp:234-242, Line::getRandPoint :140-143).
// Birth times are STRATIFIED BY MEASUREMENT STEP: the host deals a source's `count` phonons over the steps of
// its emission window in proportion to the time each step overlaps the window (flatten.cpp:plan_births,
// J(k) = ceil(count * (k dt - start) / duration)), and a phonon is born uniformly inside its step.  Same
// expectation as the reference's independent uniform draws over the window (EmitSurface::getPhononTime), lower
// variance, and "the phonons born in step k" is a contiguous index range: no sort, no birth-time array.
// `j` is the phonon's index within its source.  Returns the time left in the birth interval.
// ---------------------------------------------------------------------------------------------------------
PSIM_HD float draw_scatter_time(const DevParams& P, const DevSensor& sen, const Phonon& p, float u);

PSIM_HD float create_phonon(const DevParams& P, const DevSource& src, uint64_t j, uint32_t step, Phonon& p) {
    const uint64_t id = src.first_id + j;
    p.id_lo = static_cast<uint32_t>(id);
    const uint32_t id_hi = static_cast<uint32_t>(id >> 32);
    p.packed = (id_hi << 24) | ((src.sign < 0) ? 0x800u : 0u);
    Rng rng;
    rng_begin(rng);
    rng_refill(rng, P, PSIM_BIRTH_STEP, p.id_lo, id_hi);
    const float u_time = rng_u01(rng), u_bin = rng_u01(rng), u_pol = rng_u01(rng);
    const uint32_t u_jit = rng_word(rng) >> 24;
    rng_refill(rng, P, PSIM_BIRTH_STEP, p.id_lo, id_hi);
    const float u_a = rng_u01(rng), u_b = rng_u01(rng), u_c = rng_u01(rng), u_d = rng_u01(rng);
    float vel;
    if (src.kind == 0u) {
        const DevApiCell ac = P.api_cells[src.index];  // the model-file triangle: its flight cell and its corners in that frame
        p.cell = ac.cell;
        const uint32_t sm = cell_sensor_word(P, p.cell);
        const DevSensor s = load_rates(P, sm, 0u);  // born at t = 0
        sample_table(P, s.base_table, PSIM_CELL_MAT(sm), u_bin, u_pol, u_jit, p, vel);
        float r1 = u_a, r2 = u_b;
        if (r1 + r2 > 1.f) {
            r1 = 1.f - r1;
            r2 = 1.f - r2;
        }
        // Triangle::getRandPoint (geometry.cpp:234-242): p1 + (p2 - p1) r1 + (p3 - p1) r2, in the frame of the flight cell
        const float x1 = static_cast<float>(ac.corners & 1u), y1 = static_cast<float>((ac.corners >> 1) & 1u);
        const float x2 = static_cast<float>((ac.corners >> 2) & 1u), y2 = static_cast<float>((ac.corners >> 3) & 1u);
        const float x3 = static_cast<float>((ac.corners >> 4) & 1u), y3 = static_cast<float>((ac.corners >> 5) & 1u);
        p.b1 = fma_rn(x3 - x1, r2, fma_rn(x2 - x1, r1, x1));
        p.b2 = fma_rn(y3 - y1, r2, fma_rn(y2 - y1, r1, y1));
        if (P.lattice) {  // the triangle's parallelogram is one of nx x ny of its lattice cell
            const uint32_t dims = ldg(&P.cells[PSIM_CELL_INDEX(p.cell)].tri[1]);
            p.b1 = clamp01((p.b1 + static_cast<float>((ac.corners >> 6) & 0x1FFFu)) / static_cast<float>(dims & 0xFFFFu));
            p.b2 = clamp01((p.b2 + static_cast<float>(ac.corners >> 19)) / static_cast<float>(dims >> 16));
        }
        isotropic_direction(u_c, u_d, vel, p);
        rng_refill(rng, P, PSIM_BIRTH_STEP, p.id_lo, id_hi);
        p.tts = draw_scatter_time(P, s, p, rng_u01(rng));
        return P.step_time;  // born at t = 0
    }
    const DevEmitter em = P.emitters[src.index];
    p.cell = em.cell;
    const uint32_t sm = cell_sensor_word(P, p.cell);
    // birth time: uniform over the part of measurement step `step` that lies inside the emission window
    const double step_lo = static_cast<double>(step) * P.step_time_d;
    const double lo = fmax(em.start - step_lo, 0.), hi = fmin(em.start + em.duration - step_lo, P.step_time_d);
    double frac = (lo + (hi - lo) * static_cast<double>(u_time)) / P.step_time_d;
    frac = frac < 0. ? 0. : (frac > 0.999999 ? 0.999999 : frac);
    (void)j;
    sample_table(P, em.table, PSIM_CELL_MAT(sm), u_bin, u_pol, u_jit, p, vel);
    place_on_edge(PSIM_CELL_QUAD(p.cell), em.edge, clamp01(dot2(em.s_p1, u_a, em.s_p2, 1.f - u_a)), p);
    const float2 n = load_cell_normal(P, p.cell, em.edge);
    if (src.kind == 2u) {  // phasor: unit frequency, 1000 m/s, straight along the normal
        p.packed = (p.packed & 0xFF000800u) | 1u | (PSIM_CELL_MAT(sm) << 12);
        p.dx = 1000.f * n.x;
        p.dy = 1000.f * n.y;
    } else {
        diffuse_direction(u_b, u_c, n.x, n.y, vel, p);
    }
    p.tts = draw_scatter_time(P, load_rates(P, sm, step), p, u_d);
    return static_cast<float>((1. - frac) * P.step_time_d);
}

// ---------------------------------------------------------------------------------------------------------
// One phonon inside one measurement interval, as a small state machine so that the kernel can keep all 32
// lanes of a warp busy (a lane that finishes its phonon fetches the next one instead of idling):
//   interval_begin   rates of the current sensor area, time to the next intrinsic scatter
//   flight_window    ONE free-flight segment: it ends at an edge (EV_IMPACT), at an intrinsic scatter (EV_SCATTER)
//                    or at the end of the launch window (EV_END), crossing measurement boundaries on the way
//   impact_event     what the edge does: reflect, transmit to the neighbour cell, absorb (EV_DEAD)
//   scatter_event    the intrinsic scatter itself
// Together they replace the body of ModelSimulator::simulatePhonon (modelSimulator.cpp:139-198) between two
// measurement events, handleImpacts (:205-225), nextImpact (:87-122), scatter (:124-137),
// Cell::handleSurfaceCollision (cell.cpp:103-108), CompositeSurface::handlePhonon (compositeSurface.cpp:47-66),
// EmitSurface::handlePhonon (surface.cpp:61-65) and TransitionSurface::handlePhonon (surface.cpp:71-109).
// The time to the next intrinsic scatter travels with the phonon (state word `tts`), so an interval in which
// nothing happens costs no random number, no logarithm and no rate evaluation; it is redrawn after a scatter
// and on every change of sensor area, as in the reference (modelSimulator.cpp:155,192-194).
// ---------------------------------------------------------------------------------------------------------
enum { EV_CONTINUE = 0, EV_END = 1, EV_DEAD = 2, EV_SCATTER = 3, EV_IMPACT = 4 };

struct Flight {
    float m00, m01, m10, m11;  // barycentric rate matrix of the current cell
    uint32_t sensor_mat;       // sensor / rate class / material word of the current cell (PSIM_CELL_*)
    float vel;                 // group velocity of the phonon
    float r1, r2;              // d(b1)/dt, d(b2)/dt
    float t;                   // time left in the measurement interval (ns)
    uint32_t ncoll;            // impacts since the last scatter / interval start (stuck-phonon guard)
    uint32_t edge;             // edge reached by the last flight segment (EV_IMPACT) ...
    float s_hit;               // ... and where on it, as a fraction from the edge's first vertex
    Rng rng;
};

PSIM_HD void set_cell_matrix(Flight& f, const float4 m) {
    f.m00 = m.x;
    f.m01 = m.y;
    f.m10 = m.z;
    f.m11 = m.w;
}

PSIM_HD void update_rates_of_motion(Flight& f, const Phonon& p) {
    f.r1 = dot2(f.m00, p.dx, f.m01, p.dy);
    f.r2 = dot2(f.m10, p.dx, f.m11, p.dy);
}

PSIM_HD float draw_scatter_time(const DevParams& P, const DevSensor& sen, const Phonon& p, float u) {
    float rn, ru, ri;
    relax_rates(sen, phonon_omega(P, p.packed), PSIM_PACK_TA(p.packed), rn, ru, ri);
    const float gam = rn + ru + ri;
    return (P.phasor || !(gam > 0.f)) ? f_inf() : f_div(-f_log(u), gam);
}

PSIM_HD void interval_begin(const DevParams& P, const Phonon& p, Flight& f, float t, uint32_t step) {
    set_cell_matrix(f, load_cell_matrix(P, p.cell));
    f.sensor_mat = cell_sensor_word(P, p.cell);
    f.vel = phonon_velocity(P, p.packed);
    update_rates_of_motion(f, p);
    f.t = t;
    f.ncoll = 0;
    rng_begin(f.rng);  // the stream of this (phonon, step) starts at block 0; nothing is drawn unless an event needs it
    (void)step;
}

// next measurement interval of a phonon whose flight state is still in registers (several steps per launch)
// One free-flight segment inside the launch window [.., step_end): the phonon flies to its next PHYSICAL event
// (edge or intrinsic scatter) or to the end of the window, whichever comes first.  Measurement events on the way
// (modelSimulator.cpp:182-186) do not interrupt the flight: `on_measure(k0, k1)` is called once with the range
// [k0, k1) of RECORDED steps that ended during this segment (the caller tallies the phonon into rows k + 1 - first
// recorded step, k0 <= k < k1; its direction, velocity and cell are the same for all of them); the step counter
// advances and the per-interval bookkeeping (impact counter, Philox block of the new (phonon, step) stream) restarts.
// A measurement wins a tie with a physical event, as in the reference.
template<class OnMeasure>
PSIM_HD int flight_window(const DevParams& P, Phonon& p, Flight& f, uint32_t& s, uint32_t step_end, uint32_t& n_steps,
                          OnMeasure&& on_measure) {
    // times to the edges, one division per axis; a position that rounding left marginally outside gives a negative time,
    // taken as 0, so that the edge reached is always the one whose time equals the minimum.
    //   b2 axis: edge 0 (b2 = 0) when moving down; a parallelogram also has edge 2 (b2 = 1) when moving up
    //   b1 axis: edge 2 of a triangle / 3 of a parallelogram (b1 = 0) when moving left; a parallelogram also edge 1 (b1 = 1)
    //   the hypotenuse of a triangle (edge 1, b1 + b2 = 1)
    const float inf = f_inf();
    const bool quad = PSIM_CELL_QUAD(p.cell) != 0u;
    const bool down = f.r2 < 0.f, left = f.r1 < 0.f;
    const float tA = (down || (quad && f.r2 > 0.f)) ? fmaxf(f_div(down ? -p.b2 : 1.f - p.b2, f.r2), 0.f) : inf;
    const float tB = (left || (quad && f.r1 > 0.f)) ? fmaxf(f_div(left ? -p.b1 : 1.f - p.b1, f.r1), 0.f) : inf;
    const float rs = f.r1 + f.r2;
    const float tC = (!quad && rs > 0.f) ? fmaxf(f_div(1.f - p.b1 - p.b2, rs), 0.f) : inf;
    const float th = fminf(tA, fminf(tC, tB));
    const bool impact = th <= p.tts;        // reference: impact_time <= time (modelSimulator.cpp:111)
    const float te = impact ? th : p.tts;   // time to the next physical event
    int ev = impact ? EV_IMPACT : EV_SCATTER;
    uint32_t mk0 = 0, mk1 = 0;              // recorded steps [mk0, mk1) that end during this segment
    float flown = te;                       // time flown in this segment
    float t_left = f.t - te;                // time left in the measurement interval in which the segment ends
    if (!(te < f.t)) {
        // Boundaries of this window lie at f.t, f.t + dt, ... (step_end - s of them).  Those at or before the event
        // are crossed first (a measurement wins a tie): n = min(left, floor((te - f.t) / dt) + 1), in closed form so
        // that a phonon that is many intervals away from its next event costs no loop.
        const uint32_t left = step_end - s;
        const float q = fminf(floorf((te - f.t) * P.step_time_inv), 1.0e6f);
        const uint32_t n = min(left, static_cast<uint32_t>(q) + 1u);
        n_steps += n;
        mk0 = max(P.first_tally_step, s + 1u) - 1u;  // first RECORDED one among them
        mk1 = s + n;
        const float to_last = f.t + static_cast<float>(n - 1u) * P.step_time;  // up to the last boundary crossed
        const bool end = n == left;  // end of the launch window: the state goes back to the pool
        s = end ? step_end - 1u : s + n;
        ev = end ? EV_END : ev;
        flown = end ? to_last : te;
        t_left = end ? P.step_time : (to_last + P.step_time) - te;
        f.ncoll = 0;
        f.rng.block = 0;
    }
    f.t = t_left;
    const float nb1 = p.b1 + f.r1 * flown, nb2 = p.b2 + f.r2 * flown;
    p.tts = (ev == EV_SCATTER) ? 0.f : p.tts - flown;
    // an impact snaps the position onto the edge it reached (ties: b2 axis, then hypotenuse, then b1 axis - for a triangle
    // the order 0, 1, 2)
    const bool isA = th == tA, isC = !isA && th == tC, hit = ev == EV_IMPACT;
    const float c1 = clamp01(nb1), c2 = clamp01(nb2);
    const float wall2 = down ? 0.f : 1.f, wall1 = left ? 0.f : 1.f;  // the coordinate line reached on either axis
    const float hb1 = isA ? c1 : (isC ? 1.f - c2 : wall1), hb2 = isA ? wall2 : c2;
    p.b1 = hit ? hb1 : nb1;
    p.b2 = hit ? hb2 : nb2;
    const uint32_t eA = down ? 0u : 2u, eB = quad ? (left ? 3u : 1u) : 2u;
    f.edge = hit ? (isA ? eA : (isC ? 1u : eB)) : 0u;
    // fraction of the edge from its first vertex (device_types.h: edges run counter to b1 on edge 2 and to b2 on the b1 = 0 edge)
    f.s_hit = isA ? (down ? c1 : 1.f - c1) : (isC ? c2 : (left ? 1.f - c2 : c2));
    // the caller's tally, with the state at the END of the segment (position p, rates of motion f.r1 / f.r2, f.t left in the
    // interval in which it ends): what a lattice cell needs to find where the phonon was at each of the measurements
    if (mk0 < mk1) { on_measure(mk0, mk1); }
    return ev;
}

// Recorded windows over the lattice image.  The measurement that ended step k, k0 <= k < k1, of a segment that ended at
// (e1, e2) of lattice cell `sub` = (first sub-cell, nx | ny << 16) with t_left of its last interval remaining lies
// (dt - t_left) + (k1 - 1 - k) dt before the end of the segment.
PSIM_HD float lattice_back(const DevParams& P, float t_left, uint32_t k, uint32_t k1) {
    return (P.step_time - t_left) + static_cast<float>(k1 - 1u - k) * P.step_time;
}
PSIM_HD uint32_t lattice_sensor_at(const DevParams& P, const uint2 sub, float e1, float e2, float r1, float r2, float back) {
    float l;
    const uint32_t nx = sub.y & 0xFFFFu, ny = sub.y >> 16;
    const uint32_t ix = lattice_split(fma_rn(-r1, back, e1), nx, l), iy = lattice_split(fma_rn(-r2, back, e2), ny, l);
    return ldg(&P.sub_sensor[sub.x + iy * nx + ix]);
}
// ... and the runs of equal sensor areas among them, in step order: post(ka, kb, sensor) for every run [ka, kb)
template<class Post>
PSIM_HD void lattice_runs(const DevParams& P, uint32_t cell, float e1, float e2, float r1, float r2, float t_left, uint32_t k0,
                          uint32_t k1, Post&& post) {
    const uint2 sub = load_cell_tris(P.cells, PSIM_CELL_INDEX(cell));
    uint32_t run0 = k0, cur = lattice_sensor_at(P, sub, e1, e2, r1, r2, lattice_back(P, t_left, k0, k1));
    for (uint32_t k = k0 + 1u; k < k1; ++k) {
        const uint32_t nxt = lattice_sensor_at(P, sub, e1, e2, r1, r2, lattice_back(P, t_left, k, k1));
        if (nxt != cur) {
            post(run0, k, cur);
            run0 = k;
            cur = nxt;
        }
    }
    post(run0, k1, cur);
}

// The reference redraws the time to scatter whenever a phonon enters another sensor area (modelSimulator.cpp:
// 167-172,192-194) because the rates may differ there.  Where they provably do not - same material, same
// temperature, i.e. the same "rate class" - the exponential law is memoryless and keeping the old time is the same
// distribution, so no random number is spent.  Class 255 = "unclassified": always redraw.
PSIM_HD bool rates_differ(uint32_t cell_word_a, uint32_t cell_word_b) {
    return ((cell_word_a ^ cell_word_b) & 0xFFFu) != 0u || PSIM_CELL_CLASS(cell_word_a) == 255u;
}

// Fast path of a surface interaction, taken inline by the flight loop - the impacts that need no random number and no
// bookkeeping beyond the phonon's own state:
//   * a transition into a neighbour cell with the same material and rate class, through a whole edge or through the
//     sub-surface of a composite edge that an even division of the edge names (the edges of lattice cells, device_types.h);
//   * the mirror reflection off a perfectly specular wall (Surface::boundaryHandlePhonon, surface.cpp:32-37).
// Everything else (diffuse walls, emitters, material interfaces, other rates, irregular partial edges, the stuck-phonon
// guard) goes to impact_event.  Where both can handle an impact they compute the same thing with the same expressions.
// `links`, `tail` = the two halves of the record of p.cell; `reflected`: the velocity changed.
// `track_sensor`: the caller keeps f.sensor_mat current across segments (false: it reloads it from the cell record).
PSIM_HD bool fast_impact(const DevParams& P, Phonon& p, Flight& f, const uint4 links, const uint2 tail, bool& reflected, const bool track_sensor) {
    reflected = false;
    if (!(P.fast_links | PSIM_FAST_WALLS) || f.ncoll >= PSIM_MAX_COLLISIONS) { return false; }
    uint32_t link = link_of_edge(links, f.edge);
    const uint32_t kind = PSIM_LINK_KIND(link);
    float s_in;  // where on the neighbour's edge the phonon enters
    if (kind == PSIM_LINK_BOUNDARY) {
        if (!PSIM_FAST_WALLS || !(link & 1u)) { return false; }  // not perfectly specular: needs random numbers
        const float2 n = load_shape_normal(P.shapes, tail.y, f.edge);
        set_cell_matrix(f, load_shape_matrix(P.shapes, tail.y));  // (a caller that only flies keeps r1, r2, not the matrix)
        const float dn = dot2(p.dx, n.x, p.dy, n.y);
        p.dx = fma_rn(-2.f * dn, n.x, p.dx);
        p.dy = fma_rn(-2.f * dn, n.y, p.dy);
        update_rates_of_motion(f, p);
        ++f.ncoll;
        reflected = true;
        return true;
    }
    if (kind == PSIM_LINK_COMPOSITE) {
        if (!PSIM_FAST_COMPOSITE || !(link & (1u << 27))) { return false; }  // nothing behind this edge that the fast path could handle
        const uint32_t first = (link >> 7) & 0xFFFFFu, n = link & 0x7Fu;
        const uint32_t j = first + min(static_cast<uint32_t>(f.s_hit * static_cast<float>(n)), n - 1u);
        const float4 q = ldg(reinterpret_cast<const float4*>(P.subs + j));
        link = ldg(&P.subs[j].link);
        if (!(f.s_hit >= q.x && f.s_hit <= q.y) || PSIM_LINK_KIND(link) != PSIM_LINK_TRANSITION) { return false; }
        s_in = clamp01(q.z * f.s_hit + q.w);
    } else if (kind == PSIM_LINK_TRANSITION) {
        s_in = (link & (1u << 27)) ? f.s_hit : 1.f - f.s_hit;
    } else {
        return false;
    }
    const uint32_t ncell = PSIM_LINK_CELL(link);
    if (link & PSIM_LINK_SAME_FRAME) {
        // the same shape, material and rates on the other side (the interior of every regular mesh): the rates of motion
        // stay what they are, only the cell label and the coordinates change - nothing is loaded
        place_on_edge(PSIM_CELL_QUAD(ncell), (link >> 28) & 3u, s_in, p);
        p.cell = ncell;
        if (track_sensor) { f.sensor_mat = cell_sensor_word(P, ncell); }
        ++f.ncoll;
        return true;
    }
    const uint2 ntail = load_cell_tail(P.cells, PSIM_CELL_INDEX(ncell));  // (sensor / class / material word, shape)
    if (rates_differ(ntail.x, f.sensor_mat)) { return false; }
    place_on_edge(PSIM_CELL_QUAD(ncell), (link >> 28) & 3u, s_in, p);
    p.cell = ncell;
    f.sensor_mat = ntail.x;
    set_cell_matrix(f, load_shape_matrix(P.shapes, ntail.y));
    update_rates_of_motion(f, p);
    ++f.ncoll;
    return true;
}
PSIM_HD bool fast_impact(const DevParams& P, Phonon& p, Flight& f) {
    bool reflected;
    if (!(P.fast_links | PSIM_FAST_WALLS)) { return false; }  // (before the cell record is loaded for it)
    return fast_impact(P, p, f, load_cell_links(P.cells, PSIM_CELL_INDEX(p.cell)), load_cell_tail(P.cells, PSIM_CELL_INDEX(p.cell)), reflected, true);
}

PSIM_HD int impact_event(const DevParams& P, Phonon& p, Flight& f, uint32_t step) {
    f.rng.left = 0;  // every event that needs random numbers starts a fresh Philox block of its (phonon, step) stream
    const uint32_t e = f.edge;
    const float s = f.s_hit;
    const uint32_t id_hi = PSIM_PACK_IDHI(p.packed);
    uint32_t link = link_of_edge(load_cell_links(P.cells, PSIM_CELL_INDEX(p.cell)), e);
    float ma = (link & (1u << 27)) ? 1.f : -1.f, mb = (link & (1u << 27)) ? 0.f : 1.f;
    if (PSIM_LINK_KIND(link) == PSIM_LINK_COMPOSITE) {
        const uint32_t first = (link >> 7) & 0xFFFFFu, n = link & 0x7Fu;
        link = 0u;  // boundary unless a sub-surface covers the hit point
        // the records are searched starting at the one an even division of the edge would give (the edges of a lattice
        // cell are divided evenly: found at once)
        uint32_t j = min(static_cast<uint32_t>(s * static_cast<float>(n)), n - 1u);
        for (uint32_t i = 0; i < n; ++i) {
            const float4 q = ldg(reinterpret_cast<const float4*>(P.subs + first + j));
            if (s >= q.x && s <= q.y) {
                ma = q.z;
                mb = q.w;
                link = ldg(&P.subs[first + j].link);
                break;
            }
            j = (j + 1u == n) ? 0u : j + 1u;
        }
    }
    const uint32_t kind = PSIM_LINK_KIND(link);
    // Random words this surface can consume - asked for at ONE place, and only if any is needed: a transition
    // into the same sensor area none, into another area 1 (new time to scatter), a blocked material interface 2
    // (back-scatter direction), a wall 3 (specular test + diffuse direction) unless it is perfectly specular.
    const float spec = (kind == PSIM_LINK_TRANSITION) ? 0.f : load_cell_spec(P, p.cell);
    uint32_t ncell = 0u, nsm = 0u, need = (spec >= 1.f) ? 0u : 3u;
    bool pass = true;
    if (kind == PSIM_LINK_TRANSITION) {
        ncell = PSIM_LINK_CELL(link);
        nsm = cell_sensor_word(P, ncell);
        const uint32_t nmat = PSIM_CELL_MAT(nsm);
        if (nmat != PSIM_CELL_MAT(f.sensor_mat)) {  // material interface: no state above the neighbour's cutoff
            const float wmax = PSIM_PACK_TA(p.packed) ? ldg(&P.materials[nmat].w_max_ta) : ldg(&P.materials[nmat].w_max_la);
            pass = !(phonon_omega(P, p.packed) > wmax);
        }
        need = pass ? (rates_differ(nsm, f.sensor_mat) ? 1u : 0u) : 2u;
    }
    rng_need(f.rng, need, P, step, p.id_lo, id_hi);
    if (kind == PSIM_LINK_TRANSITION) {
        if (pass) {
            place_on_edge(PSIM_CELL_QUAD(ncell), (link >> 28) & 3u, clamp01(ma * s + mb), p);
            p.cell = ncell;
            const bool new_sensor = rates_differ(nsm, f.sensor_mat);
            f.sensor_mat = nsm;
            set_cell_matrix(f, load_cell_matrix(P, ncell));
            if (new_sensor) {  // the old time-to-scatter is void where the rates differ (modelSimulator.cpp:167-172,192-194)
                p.tts = draw_scatter_time(P, load_rates(P, nsm, step), p, rng_u01(f.rng));
            }
        } else {  // back into the same cell, about the true inward normal
            const float2 n = load_cell_normal(P, p.cell, e);
            const float u1 = rng_u01(f.rng), u2 = rng_u01(f.rng);
            diffuse_direction(u1, u2, n.x, n.y, f.vel, p);
        }
    } else {
        if (kind == PSIM_LINK_EMIT) {
            const DevEmitter* em = P.emitters + PSIM_LINK_INDEX(link);
            if (step >= ldg(&em->k_on) && step < ldg(&em->k_off)) { return EV_DEAD; }  // absorbed
        }  // outside its window an emitting surface is an ordinary wall (surface.cpp:61-65)
        const float2 n = load_cell_normal(P, p.cell, e);
        boundary_reflect(f.rng, spec, n.x, n.y, f.vel, p);
    }
    update_rates_of_motion(f, p);
    if (++f.ncoll > PSIM_MAX_COLLISIONS) {
        // stuck in a corner: restart from a random point of the current cell (modelSimulator.cpp:215-218; the
        // reference also forfeits the rest of that free-flight segment - not reproduced, it never fires in practice)
        rng_need(f.rng, 2u, P, step, p.id_lo, id_hi);
        float q1 = rng_u01(f.rng), q2 = rng_u01(f.rng);
        if (!PSIM_CELL_QUAD(p.cell) && q1 + q2 > 1.f) {  // (a parallelogram is the whole unit square of its frame)
            q1 = 1.f - q1;
            q2 = 1.f - q2;
        }
        if (P.lattice) {  // "the current cell" is the parallelogram of the lattice the phonon is in
            const uint32_t dims = ldg(&P.cells[PSIM_CELL_INDEX(p.cell)].tri[1]);
            float l1, l2;
            const uint32_t nx = dims & 0xFFFFu, ny = dims >> 16, ix = lattice_split(p.b1, nx, l1), iy = lattice_split(p.b2, ny, l2);
            q1 = clamp01((q1 + static_cast<float>(ix)) / static_cast<float>(nx));
            q2 = clamp01((q2 + static_cast<float>(iy)) / static_cast<float>(ny));
        }
        p.b1 = q1;
        p.b2 = q2;
        f.ncoll = 0;
    }
    return EV_CONTINUE;
}

// ModelSimulator::scatter (modelSimulator.cpp:124-137) with the rates of the free flight that just ended,
// then the rates and the time to scatter of the next one (get_scatter_info, :148-153)
PSIM_HD void scatter_event(const DevParams& P, Phonon& p, Flight& f, uint32_t step) {
    const uint32_t id_hi = PSIM_PACK_IDHI(p.packed);
    const DevSensor sen = load_rates(P, f.sensor_mat, step);
    float rn, ru, ri;
    relax_rates(sen, phonon_omega(P, p.packed), PSIM_PACK_TA(p.packed), rn, ru, ri);
    // ONE Philox block per scatter: word 0 -> bin (24 bits) + position inside the bin (8 bits), word 1 -> branch
    // selection + polarisation (16 bits each), word 2 -> next time to scatter, word 3 -> new direction (2 x 16 bits)
    rng_refill(f.rng, P, step, p.id_lo, id_hi);
    const uint32_t w0 = rng_word(f.rng), w1 = rng_word(f.rng), w2 = rng_word(f.rng), w3 = rng_word(f.rng);
    const float r = u01_hi16(w1) * (rn + ru + ri);
    const float u_bin = u01_from24(w0), u_pol = u01_lo16(w1);
    const uint32_t u_jit = w0 & 0xFFu;
    const float u_d1 = u01_hi16(w3), u_d2 = u01_lo16(w3), u_tts = u01_from24(w2);
    if (r <= rn + ru) {
        const float vel_old = f.vel;
        sample_table(P, sen.scatter_table, PSIM_CELL_MAT(f.sensor_mat), u_bin, u_pol, u_jit, p, f.vel);
        if (r > rn) {  // Umklapp: new direction as well
            isotropic_direction(u_d1, u_d2, f.vel, p);
        } else {       // normal process: same direction, new group velocity
            const float scale = f_div(f.vel, vel_old);
            p.dx *= scale;
            p.dy *= scale;
        }
    } else if (ri > 0.f) {
        isotropic_direction(u_d1, u_d2, f.vel, p);
    }
    f.ncoll = 0;
    p.tts = draw_scatter_time(P, sen, p, u_tts);
    update_rates_of_motion(f, p);
}

// One phonon from its current state to the end of the launch window; false if it was absorbed.  The plain
// sequential form of the functions above (used by the lock-step kernel variant and by the CPU-side test emulation).
template<class OnMeasure>
PSIM_HD bool advance_window(const DevParams& P, Phonon& p, float t_first, uint32_t s, uint32_t step_end, uint32_t& n_steps,
                            uint32_t& events, OnMeasure&& on_measure) {
    Flight f;
    interval_begin(P, p, f, t_first, s);
    for (;;) {
        ++events;
        const int ev = flight_window(P, p, f, s, step_end, n_steps, [&](uint32_t k0, uint32_t k1) { on_measure(k0, k1, p, f); });
        if (ev == EV_IMPACT) {
            if (fast_impact(P, p, f)) { continue; }
            if (impact_event(P, p, f, s) == EV_DEAD) {
                ++n_steps;
                return false;
            }
        } else if (ev == EV_SCATTER) {
            scatter_event(P, p, f, s);
        } else {
            return true;
        }
    }
}

// fixed-point flux contribution (Sensor::updateHeatParams, sensor.cpp:43-52, accumulates doubles in arbitrary
// thread order; integers make the sum independent of order and of the number of GPUs)
PSIM_HD int32_t flux_fixed(float v) {
#if defined(__CUDA_ARCH__)
    return __float2int_rn(v * static_cast<float>(1 << PSIM_FLUX_FRAC_BITS));
#else
    return static_cast<int32_t>(std::nearbyint(v * static_cast<float>(1 << PSIM_FLUX_FRAC_BITS)));
#endif
}

}  // namespace psim

#endif
