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

namespace psim {

#if defined(__CUDA_ARCH__)
PSIM_HD uint32_t mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
PSIM_HD float f_log(float x) { return __logf(x); }
PSIM_HD float f_exp(float x) { return __expf(x); }
PSIM_HD float f_cos2pi(float u) { return cospif(2.f * u); }
PSIM_HD float f_sqrt(float x) { return sqrtf(x); }
PSIM_HD float f_div(float a, float b) { return __fdividef(a, b); }
PSIM_HD float f_inf() { return __int_as_float(0x7f800000); }
template<typename T> PSIM_HD T ldg(const T* p) { return __ldg(p); }
PSIM_HD DevCell load_cell(const DevCell* cells, uint32_t i) {
    const float4* q = reinterpret_cast<const float4*>(cells + i);
    union { float4 v[4]; DevCell c; } u;
    u.v[0] = __ldg(q);
    u.v[1] = __ldg(q + 1);
    u.v[2] = __ldg(q + 2);
    u.v[3] = __ldg(q + 3);
    return u.c;
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
template<typename T> PSIM_HD T ldg(const T* p) { return *p; }
PSIM_HD DevCell load_cell(const DevCell* cells, uint32_t i) { return cells[i]; }
PSIM_HD DevSensor load_sensor(const DevSensor* sensors, uint32_t i) { return sensors[i]; }
#endif

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Replaces the reference's thread_local mt19937 seeded from
// std::random_device (utils.h:16-21).  key = seed, counter = (block, step, id_lo, id_hi).
// ---------------------------------------------------------------------------------------------------------
struct Rng {
    uint32_t k0, k1;
    uint32_t step, id_lo, id_hi;
    uint32_t block;        // next 4-word block of this (id, step) stream
    uint32_t v0, v1, v2, v3;
    uint32_t left;         // unread words among v0..v3

    PSIM_HD void init(uint32_t seed_lo, uint32_t seed_hi, uint32_t step_, uint32_t id_lo_, uint32_t id_hi_) {
        k0 = seed_lo;
        k1 = seed_hi;
        step = step_;
        id_lo = id_lo_;
        id_hi = id_hi_;
        block = 0;
        left = 0;
        v0 = v1 = v2 = v3 = 0;
    }
    PSIM_HD void refill() {
        uint32_t c0 = block++, c1 = step, c2 = id_lo, c3 = id_hi;
        uint32_t a = k0, b = k1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ b;
            c3 = lo0;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        v0 = c0;
        v1 = c1;
        v2 = c2;
        v3 = c3;
        left = 4;
    }
    PSIM_HD uint32_t next() {
        if (left == 0) { refill(); }
        const uint32_t r = v0;
        v0 = v1;
        v1 = v2;
        v2 = v3;
        --left;
        return r;
    }
    // uniform on (0, 1] with 24 random bits (the reference draws doubles on [0, 1], utils.h:19)
    PSIM_HD float u01() { return static_cast<float>((next() >> 8) + 1u) * 5.9604644775390625e-8f; }
};

struct Phonon {
    float b1, b2;     // position in the current cell's frame
    float dx, dy;     // direction; in-plane speed = velocity * |d|   (phonon.cpp:28-31)
    float w;          // angular frequency * 1e-13
    uint32_t packed;  // see device_types.h
    uint32_t cell;
    uint32_t id_lo;
};

// ---------------------------------------------------------------------------------------------------------
// Frequency / polarisation sampling  (Material::freqIndex material.cpp:64-75, getFreq :77-80, getVel :82-84,
// called from SensorController::initialUpdate / scatterUpdate, sensorController.cpp:28-36,52-55).
// Same bisection as the reference, so the same quirk: it returns `high`, i.e. bin 0 is never produced.
// ---------------------------------------------------------------------------------------------------------
PSIM_HD uint32_t bisect_table(const float2* table, float r) {
    uint32_t lo = 0, hi = PSIM_BINS - 1;
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

PSIM_HD void sample_table(const DevParams& P, uint32_t table_idx, uint32_t mat, Rng& rng, Phonon& p, float& vel) {
    const float2* table = P.tables + static_cast<size_t>(table_idx) * PSIM_BINS;
    const uint32_t bin = bisect_table(table, rng.u01());
    const uint32_t ta = (rng.u01() <= ldg(&table[bin].y)) ? 0u : 1u;
    const float fw = ldg(&P.materials[mat].freq_width);
    float w = (2.f * static_cast<float>(bin) + 1.f) * 0.5f * fw;
    if (!P.full_mode) { w += (2.f * rng.u01() - 1.f) * 0.5f * fw; }
    p.w = w;
    p.packed = (p.packed & 0xFFFF0800u) | bin | (ta << 10) | (mat << 12);
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
PSIM_HD void isotropic_direction(Rng& rng, Phonon& p) {
    const float dx = 2.f * rng.u01() - 1.f;
    p.dx = dx;
    p.dy = f_sqrt(fmaxf(1.f - dx * dx, 0.f)) * f_cos2pi(rng.u01());
}

// Surface::redirectPhonon (surface.cpp:23-30): cosine-law direction about the inward normal n
PSIM_HD void diffuse_direction(Rng& rng, float nx, float ny, Phonon& p) {
    const float r = rng.u01();
    const float a = f_sqrt(r);
    const float b = f_sqrt(fmaxf(1.f - r, 0.f)) * f_cos2pi(rng.u01());
    p.dx = nx * a - ny * b;
    p.dy = ny * a + nx * b;
}

// Surface::boundaryHandlePhonon (surface.cpp:32-44)
PSIM_HD void boundary_reflect(Rng& rng, float spec, float nx, float ny, Phonon& p) {
    if (spec >= 1.f || rng.u01() < spec) {
        const float dn = p.dx * nx + p.dy * ny;
        p.dx -= 2.f * dn * nx;
        p.dy -= 2.f * dn * ny;
    } else {
        diffuse_direction(rng, nx, ny, p);
    }
}

PSIM_HD float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// position on edge `e` at fraction s from the edge's first vertex
PSIM_HD void place_on_edge(uint32_t e, float s, Phonon& p) {
    if (e == 0u) {
        p.b1 = s;
        p.b2 = 0.f;
    } else if (e == 1u) {
        p.b1 = 1.f - s;
        p.b2 = s;
    } else {
        p.b1 = 0.f;
        p.b2 = 1.f - s;
    }
}

PSIM_HD void edge_normal(const DevCell& c, uint32_t e, float& nx, float& ny) {
    nx = (e == 0u) ? c.n0x : ((e == 1u) ? c.n1x : c.n2x);
    ny = (e == 0u) ? c.n0y : ((e == 1u) ? c.n1y : c.n2y);
}

// ---------------------------------------------------------------------------------------------------------
// Emission (CellOriginBuilder::operator() phononBuilder.cpp:6-16, SurfaceOriginBuilder :31-40,
// PhasorBuilder :42-49, EmitSurface::getPhononTime surface.cpp:67-69, Triangle::getRandPoint
// geometry.cpp:234-242, Line::getRandPoint :140-143).
// `j` is the phonon's index within its source.  Birth times are STRATIFIED BY MEASUREMENT STEP: the host deals
// a source's `count` phonons over the steps of its emission window in proportion to the time each step overlaps
// the window (flatten.cpp:plan_births, J(k) = ceil(count * (k dt - start) / duration)), and a phonon is born
// uniformly inside its step.  Same expectation as the reference's independent uniform draws over the window
// (EmitSurface::getPhononTime), lower variance, and "the phonons born in step k" is a contiguous index range:
// no sort, no birth-time array.  Returns the time left in the birth interval.
// ---------------------------------------------------------------------------------------------------------
PSIM_HD float create_phonon(const DevParams& P, const DevSource& src, uint64_t j, uint32_t step, Phonon& p) {
    const uint64_t id = src.first_id + j;
    p.id_lo = static_cast<uint32_t>(id);
    const uint32_t id_hi = static_cast<uint32_t>(id >> 32);
    p.packed = (id_hi << 16) | ((src.sign < 0) ? 0x800u : 0u);
    Rng rng;
    rng.init(P.seed_lo, P.seed_hi, PSIM_BIRTH_STEP, p.id_lo, id_hi);
    float vel;
    if (src.kind == 0u) {
        p.cell = src.index;
        const DevCell c = load_cell(P.cells, p.cell);
        const DevSensor s = load_sensor(P.sensors, c.sensor_mat >> 8);
        sample_table(P, s.base_table, c.sensor_mat & 0xFFu, rng, p, vel);
        float r1 = rng.u01(), r2 = rng.u01();
        if (r1 + r2 > 1.f) {
            r1 = 1.f - r1;
            r2 = 1.f - r2;
        }
        p.b1 = r1;
        p.b2 = r2;
        isotropic_direction(rng, p);
        return P.step_time;  // born at t = 0
    }
    const DevEmitter em = P.emitters[src.index];
    p.cell = em.cell;
    const DevCell c = load_cell(P.cells, p.cell);
    // birth time: uniform over the part of measurement step `step` that lies inside the emission window
    const double step_lo = static_cast<double>(step) * P.step_time_d;
    const double lo = fmax(em.start - step_lo, 0.), hi = fmin(em.start + em.duration - step_lo, P.step_time_d);
    double frac = (lo + (hi - lo) * static_cast<double>(rng.u01())) / P.step_time_d;
    frac = frac < 0. ? 0. : (frac > 0.999999 ? 0.999999 : frac);
    (void)j;
    sample_table(P, em.table, c.sensor_mat & 0xFFu, rng, p, vel);
    const float u = rng.u01();
    place_on_edge(em.edge, clamp01(em.s_p1 * u + em.s_p2 * (1.f - u)), p);
    float nx, ny;
    edge_normal(c, em.edge, nx, ny);
    if (src.kind == 2u) {  // phasor: unit frequency, 1000 m/s, straight along the normal
        p.w = static_cast<float>(PSIM_FREQ_SCALE);
        p.packed = (p.packed & 0xFFFF0800u) | 1u | ((c.sensor_mat & 0xFu) << 12);
        p.dx = nx;
        p.dy = ny;
    } else {
        diffuse_direction(rng, nx, ny, p);
    }
    return static_cast<float>((1. - frac) * P.step_time_d);
}

// ---------------------------------------------------------------------------------------------------------
// One phonon across (the rest of) one measurement interval: free flight, intrinsic scattering, surface
// interaction, cell transition.  Replaces the body of ModelSimulator::simulatePhonon (modelSimulator.cpp:
// 139-198) between two measurement events, handleImpacts (:205-225), nextImpact (:87-122), scatter (:124-137),
// Cell::handleSurfaceCollision (cell.cpp:103-108), CompositeSurface::handlePhonon (compositeSurface.cpp:47-66),
// EmitSurface::handlePhonon (surface.cpp:61-65) and TransitionSurface::handlePhonon (surface.cpp:71-109).
// The time to the next intrinsic scatter is redrawn at the start of every interval instead of being carried
// in the state: the exponential law is memoryless and the rates are constant inside a sensor area (the
// reference redraws on every sensor change too, modelSimulator.cpp:192-194).
// Returns false if the phonon left the system through an absorbing (emitting) surface.
// `vel` is in/out (group velocity, changes when the phonon is resampled); `sensor_out` = sensor at the end.
// ---------------------------------------------------------------------------------------------------------
PSIM_HD bool advance_interval(const DevParams& P, Phonon& p, float t, uint32_t step, float& vel,
                              uint32_t& sensor_out, uint32_t& events) {
    Rng rng;
    rng.init(P.seed_lo, P.seed_hi, step, p.id_lo, PSIM_PACK_IDHI(p.packed));
    DevCell c = load_cell(P.cells, p.cell);
    DevSensor sen = load_sensor(P.sensors, c.sensor_mat >> 8);
    uint32_t ta = PSIM_PACK_TA(p.packed);
    float rn, ru, ri;
    relax_rates(sen, p.w, ta, rn, ru, ri);
    float gam = rn + ru + ri;
    const float inf = f_inf();
    float tts = (P.phasor || !(gam > 0.f)) ? inf : f_div(-f_log(rng.u01()), gam);
    float vx = p.dx * vel, vy = p.dy * vel;
    float r1 = c.m00 * vx + c.m01 * vy;
    float r2 = c.m10 * vx + c.m11 * vy;
    uint32_t ncoll = 0;
    for (;;) {
        ++events;
        const float dt = fminf(tts, t);
        const float t0 = (r2 < 0.f) ? f_div(-p.b2, r2) : inf;
        const float t2 = (r1 < 0.f) ? f_div(-p.b1, r1) : inf;
        const float rs = r1 + r2;
        const float t1 = (rs > 0.f) ? f_div(1.f - p.b1 - p.b2, rs) : inf;
        float th = fminf(t0, fminf(t1, t2));
        if (th <= dt) {  // reaches an edge first (reference: impact_time <= time, modelSimulator.cpp:111)
            const uint32_t e = (th == t0) ? 0u : ((th == t1) ? 1u : 2u);
            th = fmaxf(th, 0.f);
            float s;
            if (e == 0u) {
                s = clamp01(p.b1 + r1 * th);
            } else if (e == 1u) {
                s = clamp01(p.b2 + r2 * th);
            } else {
                s = clamp01(1.f - (p.b2 + r2 * th));
            }
            place_on_edge(e, s, p);
            t -= th;
            tts -= th;
            uint32_t link = (e == 0u) ? c.link[0] : ((e == 1u) ? c.link[1] : c.link[2]);
            float ma = (link & (1u << 27)) ? 1.f : -1.f, mb = (link & (1u << 27)) ? 0.f : 1.f;
            if (PSIM_LINK_KIND(link) == PSIM_LINK_COMPOSITE) {
                const uint32_t first = (link >> 7) & 0xFFFFFu, n = link & 0x7Fu;
                link = 0u;  // boundary unless a sub-surface covers the hit point
                for (uint32_t i = 0; i < n; ++i) {
                    const float4 q = ldg(reinterpret_cast<const float4*>(P.subs + first + i));
                    if (s >= q.x && s <= q.y) {
                        ma = q.z;
                        mb = q.w;
                        link = ldg(&P.subs[first + i].link);
                        break;
                    }
                }
            }
            const uint32_t kind = PSIM_LINK_KIND(link);
            float nx, ny;
            edge_normal(c, e, nx, ny);
            bool redirected = true;
            if (kind == PSIM_LINK_TRANSITION) {
                const uint32_t ncell = PSIM_LINK_INDEX(link);
                const DevCell nb = load_cell(P.cells, ncell);
                const uint32_t nmat = nb.sensor_mat & 0xFFu;
                bool pass = true;
                if (nmat != (c.sensor_mat & 0xFFu)) {  // material interface: no state above the neighbour's cutoff
                    const float wmax = ta ? ldg(&P.materials[nmat].w_max_ta) : ldg(&P.materials[nmat].w_max_la);
                    pass = !(p.w > wmax);
                }
                if (pass) {
                    place_on_edge((link >> 28) & 3u, clamp01(ma * s + mb), p);
                    p.cell = ncell;
                    const bool new_sensor = (nb.sensor_mat >> 8) != (c.sensor_mat >> 8);
                    c = nb;
                    redirected = false;
                    if (new_sensor) {  // old time-to-scatter is void in the new sensor area (modelSimulator.cpp:167-172,192-194)
                        sen = load_sensor(P.sensors, c.sensor_mat >> 8);
                        relax_rates(sen, p.w, ta, rn, ru, ri);
                        gam = rn + ru + ri;
                        tts = (P.phasor || !(gam > 0.f)) ? inf : f_div(-f_log(rng.u01()), gam);
                    }
                } else {
                    diffuse_direction(rng, nx, ny, p);  // back into the same cell, about the true inward normal
                }
            } else if (kind == PSIM_LINK_EMIT) {
                const DevEmitter* em = P.emitters + PSIM_LINK_INDEX(link);
                if (step >= ldg(&em->k_on) && step < ldg(&em->k_off)) { return false; }  // absorbed
                boundary_reflect(rng, c.spec, nx, ny, p);  // outside its window it is an ordinary wall
            } else {
                boundary_reflect(rng, c.spec, nx, ny, p);
            }
            if (redirected) {
                vx = p.dx * vel;
                vy = p.dy * vel;
            }
            r1 = c.m00 * vx + c.m01 * vy;
            r2 = c.m10 * vx + c.m11 * vy;
            if (++ncoll > PSIM_MAX_COLLISIONS) {
                // stuck in a corner: random point of the current cell, and the rest of this free flight is
                // spent (modelSimulator.cpp:215-218)
                float q1 = rng.u01(), q2 = rng.u01();
                if (q1 + q2 > 1.f) {
                    q1 = 1.f - q1;
                    q2 = 1.f - q2;
                }
                p.b1 = q1;
                p.b2 = q2;
                r1 = 0.f;
                r2 = 0.f;
            }
            continue;
        }
        p.b1 += r1 * dt;
        p.b2 += r2 * dt;
        if (!(tts < t)) { break; }  // measurement event (it wins ties, modelSimulator.cpp:182)
        t -= dt;
        ncoll = 0;
        // intrinsic scatter (ModelSimulator::scatter, modelSimulator.cpp:124-137)
        const float r = rng.u01() * gam;
        if (r <= rn + ru) {
            sample_table(P, sen.scatter_table, c.sensor_mat & 0xFFu, rng, p, vel);
            ta = PSIM_PACK_TA(p.packed);
            if (r > rn) { isotropic_direction(rng, p); }  // Umklapp
        } else if (ri > 0.f) {
            isotropic_direction(rng, p);
        }
        relax_rates(sen, p.w, ta, rn, ru, ri);
        gam = rn + ru + ri;
        tts = !(gam > 0.f) ? inf : f_div(-f_log(rng.u01()), gam);
        vx = p.dx * vel;
        vy = p.dy * vel;
        r1 = c.m00 * vx + c.m01 * vy;
        r2 = c.m10 * vx + c.m11 * vy;
    }
    sensor_out = c.sensor_mat >> 8;
    return true;
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
