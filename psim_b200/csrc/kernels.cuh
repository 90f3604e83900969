// sm_100a kernels of the particle loop.  Included by psim_gpu.cu only.
//
// drift_kernel_queues<NS, TALLY>  the drift step (default).  Persistent, one 24-warp CTA per SM, one pool segment per resident
//                          warp; a warp keeps NS phonons in flight in shared memory and one queue of slot numbers per kind
//                          of work (fly / intrinsic scatter / surface interaction / write back / fetch); a pass pops up
//                          to 32 entries of the fullest queue, lane i takes the i-th.  One instantiation per way a launch
//                          window deals with its measurements: none recorded / staged per CTA in shared memory / posted
//                          to global memory / the same over the lattice image (device_types.h).
// drift_kernel_slots<K>    the version before: every lane owns K slots, the warp votes for the kind most lanes want.
// drift_kernel_lockstep    the first version: tiles of 32 phonons, one per lane, in lock step.
//                          Both kept for A/B measurements and as cross-checks: a phonon's random stream is addressed by
//                          (id, step), so all three must produce bit-identical tallies (tests/test_gpu_parity.py).
// History of the design, with the ncu numbers that drove it, is in DESIGN.md section 6 and profiles/.
#ifndef PSIM_B200_KERNELS_CUH
#define PSIM_B200_KERNELS_CUH

#include "device_core.cuh"

namespace {

// Tuning: threads per CTA, phonons in flight per lane, CTAs per SM the kernels are compiled for.  ONE CTA of 24 warps
// per SM: the same 24 resident warps as three CTAs of 8 (80 registers per thread), but the tally staging exists once
// per SM instead of three times, so a launch can cover 36 recorded steps of a 50-sensor model instead of 18 and still
// leave 60 KB of L1 for the tables (measured: 97.4 -> 90.0 ms per job; three 8-warp CTAs with the same staging pushed
// the shared-memory carve-out to 228 KB and the L1 hit rate of the recorded windows to 65 %).
#ifndef PSIM_BLOCK
#define PSIM_BLOCK 768
#endif
#ifndef PSIM_SLOTS
#define PSIM_SLOTS 4
#define PSIM_SLOT_BLOCKS 1
#endif
#ifndef PSIM_QUEUE_SLOTS
#define PSIM_QUEUE_SLOTS 128  // phonons in flight per warp of the work-queue kernel (its second instantiation has 64)
#endif
#ifndef PSIM_SMEM_KB_PER_SM
#define PSIM_SMEM_KB_PER_SM 196  // shared-memory carve-out the launches are planned for (the rest of the 256 KB is L1)
#endif
constexpr int kBlock = PSIM_BLOCK;
constexpr int kWarpsPerBlock = kBlock / 32;
constexpr int kSlots = PSIM_SLOTS;
constexpr int kSlotBlocks = PSIM_SLOT_BLOCKS;
struct LaunchArgs {
    DevParams P;
    const float4* in_a;
    const uint4* in_b;
    float4* out_a;
    uint4* out_b;
    const uint32_t* cnt_in;
    uint32_t* cnt_out;
    uint32_t seg_cap;
    uint32_t n_warps;
    uint32_t step_begin, step_end;
    const DevBirth* births;        // (step, source) groups of this launch
    const uint64_t* birth_prefix;  // n_birth_entries + 1 running counts (absolute; birth_base = value at [0])
    uint32_t n_birth_entries;
    uint32_t birth_warp_offset;
    uint64_t birth_base;
    uint64_t n_births;
    int32_t* tally_e;
    long long* tally_f;
    uint32_t tally_shared;
    long long* tally_acc;            // difference-form accumulator of the many-sensor models: int64 (e, fx, fy, -)[R][S], one 32-byte sector per entry
    unsigned long long* stats;       // [0] drift steps [1] flight segments [2] absorbed [3] pool overflow [4] Philox block budget exceeded
    unsigned long long* alive_hist;  // [launch]: pool population after this launch
    uint32_t launch_index;
    // the pool this launch reads is in lattice coordinates (device_types.h) and the launch flies the fine image: every phonon
    // fetched from the pool is converted first (the first recording launch after the unrecorded ones)
    uint32_t convert_input;
    const DevCell* lattice_cells;
    const uint32_t* lattice_sub_fine;
};

__host__ __device__ __forceinline__ size_t tally_smem_offset_f(uint32_t nst, uint32_t S) {
    return (static_cast<size_t>(nst) * S * 4 + 15) & ~static_cast<size_t>(15);
}

// Shared-memory bytes of the per-CTA tally staging of a window of nst steps, by form (LaunchArgs::tally_shared): 1 five
// and 4 seven 32-bit words per (step, sensor), 2 an int32 array followed by an int64[2] array.
__host__ __device__ __forceinline__ size_t tally_stage_bytes(uint32_t form, uint32_t nst, uint32_t S) {
    const size_t n = static_cast<size_t>(nst) * S;
    return form == 4u ? ((n * 28 + 15) & ~static_cast<size_t>(15)) : tally_smem_offset_f(nst, S) + n * 16;
}

__device__ __forceinline__ void atomic_add_i64(long long* p, long long v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}

// Sensor::updateHeatParams (sensor.cpp:43-52): energy += sign, flux += sign * v.  Block stage in shared memory
// (flushed with one global atomic per touched (sensor, step) at the end of the kernel) or straight to global.
// Shared memory has no native 64-bit add (the compiler emits a compare-and-swap loop), so the staged flux is normally
// kept as two 32-bit sums per component - the low kStageLoBits bits of every contribution (unsigned) and the rest
// (signed) - which native ATOMS.ADD can accumulate (tally_shared == 1, see tally_range); the exact 64-bit value
// (hi << kStageLoBits) + lo is rebuilt at the flush.  The host checks the bound that keeps this exact and otherwise asks
// for the 64-bit staging (tally_shared == 2), which this function serves together with the plain global adds.
constexpr uint32_t kStageLoBits = 11u;
constexpr uint32_t kStageLoMask = (1u << kStageLoBits) - 1u;
// Pools too large for that bound (above ~1.2e8 phonons per GPU) split every contribution in THREE: two unsigned parts of
// kStageWideBits bits and the signed rest (tally_shared == 4, seven words per entry): 14 native atomics per flight
// segment instead of 10, exact up to 2^24 contributions per entry.
constexpr uint32_t kStageWideBits = 8u;
constexpr uint32_t kStageWideMask = (1u << kStageWideBits) - 1u;

__device__ __forceinline__ void tally_add(const LaunchArgs& a, int32_t* acc_e, long long* acc_f, uint32_t local_row,
                                          uint32_t sensor, int32_t e, int32_t fx, int32_t fy) {
    const uint32_t S = a.P.n_sensors;
    if (a.tally_shared == 2u) {
        const uint32_t k = local_row * S + sensor;
        atomicAdd(&acc_e[k], e);
        atomic_add_i64(&acc_f[2 * k], fx);
        atomic_add_i64(&acc_f[2 * k + 1], fy);
    } else {
        const size_t k = static_cast<size_t>(a.step_begin + local_row + 1 - a.P.first_tally_step) * S + sensor;
        atomicAdd(&a.tally_e[k], e);
        atomic_add_i64(&a.tally_f[2 * k], fx);
        atomic_add_i64(&a.tally_f[2 * k + 1], fy);
    }
}

// five consecutive 32-bit words per (step, sensor): e, fx low / high, fy low / high
__device__ __forceinline__ void stage_add_narrow(uint32_t* q, int32_t e, int32_t fx, int32_t fy) {
    atomicAdd(reinterpret_cast<int32_t*>(q), e);
    atomicAdd(q + 1, static_cast<uint32_t>(fx) & kStageLoMask);
    atomicAdd(reinterpret_cast<int32_t*>(q + 2), fx >> kStageLoBits);
    atomicAdd(q + 3, static_cast<uint32_t>(fy) & kStageLoMask);
    atomicAdd(reinterpret_cast<int32_t*>(q + 4), fy >> kStageLoBits);
}

// seven words: e, fx low / middle / high, fy low / middle / high
__device__ __forceinline__ void stage_add_wide(uint32_t* q, int32_t e, int32_t fx, int32_t fy) {
    atomicAdd(reinterpret_cast<int32_t*>(q), e);
    atomicAdd(q + 1, static_cast<uint32_t>(fx) & kStageWideMask);
    atomicAdd(q + 2, (static_cast<uint32_t>(fx) >> kStageWideBits) & kStageWideMask);
    atomicAdd(reinterpret_cast<int32_t*>(q + 3), fx >> (2u * kStageWideBits));
    atomicAdd(q + 4, static_cast<uint32_t>(fy) & kStageWideMask);
    atomicAdd(q + 5, (static_cast<uint32_t>(fy) >> kStageWideBits) & kStageWideMask);
    atomicAdd(reinterpret_cast<int32_t*>(q + 6), fy >> (2u * kStageWideBits));
}

// Tally a phonon into the recorded steps [k0, k1) that ended during one flight segment (same sensor, sign and velocity
// for all of them).  Rows are kept as DIFFERENCES along the step axis - +v at the first row, -v behind the last - so a
// segment costs a fixed number of atomics however many steps it crossed, and no loop whose trip count differs from lane
// to lane (ncu, looped form: 12 % of the instructions of a recorded window were shared atomics at 6 of 32 lanes).
//   tally_shared == 1  staged per CTA in shared memory as 32-bit halves: ten native ATOMS.ADD per segment; tally_flush
//                      turns the rows of the window into running sums (a row's difference beyond the window is dropped:
//                      the next launch starts its own sums).  Exact while 2 x (phonons a block handles) contributions
//                      fit the halves: a phonon adds to a row at most twice, once as a first row, once behind a last.
//   tally_shared == 3  no staging (models with many sensors, whose staging would not fit): six global REDs per segment;
//                      psim_gpu.cu:finalize_rows turns the rows into running sums once their window is complete.
// Integers throughout: the sums are exact and independent of the order of the adds.
__device__ __forceinline__ void tally_range(const LaunchArgs& a, int32_t* acc_e, long long* acc_f, uint32_t k0, uint32_t k1,
                                            uint32_t sensor, int32_t e, int32_t fx, int32_t fy) {
    if (a.tally_shared == 1u) {
        const uint32_t S = a.P.n_sensors;
        const uint32_t l0 = k0 - a.step_begin, l1 = k1 - a.step_begin;
        uint32_t* q = reinterpret_cast<uint32_t*>(acc_e) + 5u * (l0 * S + sensor);
        stage_add_narrow(q, e, fx, fy);
        if (l1 < a.step_end - a.step_begin) { stage_add_narrow(q + 5u * (l1 - l0) * S, -e, -fx, -fy); }
    } else if (a.tally_shared == 4u) {
        const uint32_t S = a.P.n_sensors;
        const uint32_t l0 = k0 - a.step_begin, l1 = k1 - a.step_begin;
        uint32_t* q = reinterpret_cast<uint32_t*>(acc_e) + 7u * (l0 * S + sensor);
        stage_add_wide(q, e, fx, fy);
        if (l1 < a.step_end - a.step_begin) { stage_add_wide(q + 7u * (l1 - l0) * S, -e, -fx, -fy); }
    } else if (a.tally_shared == 3u) {
        // (the lane-bound and lock-step kernels; the work-queue kernel posts these warp-cooperatively, tally_post_global)
        const uint32_t S = a.P.n_sensors;
        const uint32_t r0 = k0 + 1u - a.P.first_tally_step, r1 = k1 + 1u - a.P.first_tally_step;
        long long* q0 = a.tally_acc + 4u * (static_cast<size_t>(r0) * S + sensor);
        atomic_add_i64(q0, e);
        atomic_add_i64(q0 + 1, fx);
        atomic_add_i64(q0 + 2, fy);
        if (r1 < a.P.recorded_steps) {
            long long* q1 = a.tally_acc + 4u * (static_cast<size_t>(r1) * S + sensor);
            atomic_add_i64(q1, -static_cast<long long>(e));
            atomic_add_i64(q1 + 1, -static_cast<long long>(fx));
            atomic_add_i64(q1 + 2, -static_cast<long long>(fy));
        }
    } else {
        for (uint32_t ks = k0; ks < k1; ++ks) { tally_add(a, acc_e, acc_f, ks - a.step_begin, sensor, e, fx, fy); }
    }
}

__device__ __forceinline__ bool tally_staged(const LaunchArgs& a) { return a.tally_shared == 1u || a.tally_shared == 2u || a.tally_shared == 4u; }

__device__ __forceinline__ void tally_init(const LaunchArgs& a, int32_t* acc_e, long long* acc_f) {
    if (!tally_staged(a)) { return; }
    const uint32_t n = (a.step_end - a.step_begin) * a.P.n_sensors;
    if (a.tally_shared == 1u || a.tally_shared == 4u) {
        const uint32_t words = (a.tally_shared == 1u ? 5u : 7u) * n;
        for (uint32_t i = threadIdx.x; i < words; i += kBlock) { acc_e[i] = 0; }
    } else {
        for (uint32_t i = threadIdx.x; i < n; i += kBlock) {
            acc_e[i] = 0;
            acc_f[2 * i] = 0;
            acc_f[2 * i + 1] = 0;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void tally_flush(const LaunchArgs& a, const int32_t* acc_e, const long long* acc_f) {
    if (!tally_staged(a)) { return; }
    __syncthreads();
    const uint32_t S = a.P.n_sensors, nst = a.step_end - a.step_begin;
    if (a.tally_shared == 1u || a.tally_shared == 4u) {
        // difference rows -> running sums along the steps of the window, one thread per (sensor, component)
        const uint32_t* words = reinterpret_cast<const uint32_t*>(acc_e);
        const bool wide = a.tally_shared == 4u;
        for (uint32_t i = threadIdx.x; i < 3u * S; i += kBlock) {
            const uint32_t s = i / 3u, c = i % 3u;
            long long run = 0;
            for (uint32_t l = 0; l < nst; ++l) {
                const uint32_t* q = words + (wide ? 7u : 5u) * (l * S + s);
                if (c == 0u) {
                    run += static_cast<long long>(static_cast<int32_t>(q[0]));
                } else if (wide) {
                    const uint32_t* p = q + 3u * c - 2u;  // low, middle, high of this component
                    run += static_cast<long long>(static_cast<int32_t>(p[2])) * (1 << (2u * kStageWideBits)) +
                           static_cast<long long>(p[1]) * (1 << kStageWideBits) + static_cast<long long>(p[0]);
                } else {
                    run += static_cast<long long>(static_cast<int32_t>(q[2u * c])) * (1 << kStageLoBits) + static_cast<long long>(q[2u * c - 1u]);
                }
                const uint32_t row = a.step_begin + l + 1u;
                if (row < a.P.first_tally_step || run == 0) { continue; }
                const size_t k = static_cast<size_t>(row - a.P.first_tally_step) * S + s;
                if (c == 0u) {
                    atomicAdd(&a.tally_e[k], static_cast<int32_t>(run));
                } else {
                    atomic_add_i64(&a.tally_f[2 * k + c - 1u], run);
                }
            }
        }
        return;
    }
    const uint32_t n = nst * S;
    for (uint32_t i = threadIdx.x; i < n; i += kBlock) {
        const uint32_t row = a.step_begin + i / S + 1;
        if (row < a.P.first_tally_step) { continue; }
        const size_t k = static_cast<size_t>(row - a.P.first_tally_step) * S + (i % S);
        const int32_t e = acc_e[i];
        const long long fx = acc_f[2 * i], fy = acc_f[2 * i + 1];
        if (e) { atomicAdd(&a.tally_e[k], e); }
        if (fx) { atomic_add_i64(&a.tally_f[2 * k], fx); }
        if (fy) { atomic_add_i64(&a.tally_f[2 * k + 1], fy); }
    }
}

__device__ __forceinline__ void load_phonon(const LaunchArgs& a, size_t i, psim::Phonon& p) {
    const float4 va = __ldcs(a.in_a + i);  // streamed once: evict-first
    const uint4 vb = __ldcs(a.in_b + i);
    p.b1 = va.x;
    p.b2 = va.y;
    p.dx = va.z;
    p.dy = va.w;
    p.tts = __uint_as_float(vb.x);
    p.packed = vb.y;
    p.cell = vb.z;
    p.id_lo = vb.w;
    if (a.convert_input) { psim::coarse_to_fine(a.lattice_cells, a.lattice_sub_fine, p.cell, p.b1, p.b2); }
}

__device__ __forceinline__ void store_phonon(const LaunchArgs& a, size_t i, const psim::Phonon& p) {
    a.out_a[i] = make_float4(p.b1, p.b2, p.dx, p.dy);
    a.out_b[i] = make_uint4(__float_as_uint(p.tts), p.packed, p.cell, p.id_lo);
}

// emission: which (step, source) group does birth item `item` of this launch belong to, then build the phonon
__device__ __forceinline__ float birth_phonon(const LaunchArgs& a, uint64_t item, psim::Phonon& p, uint32_t& step) {
    const uint64_t key = item + a.birth_base;
    uint32_t lo = 0, hi = a.n_birth_entries;  // prefix[lo] <= key < prefix[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (key < __ldg(&a.birth_prefix[mid])) {
            hi = mid;
        } else {
            lo = mid;
        }
    }
    const DevBirth b = a.births[lo];
    step = b.step;
    return psim::create_phonon(a.P, a.P.sources[b.source], b.j0 + (key - __ldg(&a.birth_prefix[lo])) * b.stride, b.step, p);
}

__device__ __forceinline__ void warp_stats(const LaunchArgs& a, uint32_t lane, unsigned long long steps, unsigned long long events,
                                           unsigned long long absorbed, bool overflow, bool rng_over, uint32_t n_out) {
    for (int o = 16; o > 0; o >>= 1) {
        steps += __shfl_xor_sync(0xFFFFFFFFu, steps, o);
        events += __shfl_xor_sync(0xFFFFFFFFu, events, o);
        absorbed += __shfl_xor_sync(0xFFFFFFFFu, absorbed, o);
    }
    const bool any_overflow = __any_sync(0xFFFFFFFFu, overflow), any_rng = __any_sync(0xFFFFFFFFu, rng_over);
    if (lane == 0) {
        if (steps) { atomicAdd(&a.stats[0], steps); }
        if (events) { atomicAdd(&a.stats[1], events); }
        if (absorbed) { atomicAdd(&a.stats[2], absorbed); }
        if (any_overflow) { atomicAdd(&a.stats[3], 1ull); }
        if (any_rng) { atomicAdd(&a.stats[4], 1ull); }
        if (n_out) { atomicAdd(&a.alive_hist[a.launch_index], static_cast<unsigned long long>(n_out)); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// drift_kernel_slots<K>: emission + free flight / intrinsic scattering / surfaces / cell transitions + tally +
// compaction, with K phonons IN FLIGHT PER LANE.
//
// A warp that holds one phonon per lane in registers can only regroup 32 phonons: with five kinds of work a kind
// rarely has more than ~14 takers (ncu of that version: 14.4 of 32 threads active per instruction).  Here every lane
// owns K slots in shared memory (12 words each, laid out [field][slot][lane] so that lane l only ever touches bank l:
// no bank conflicts), i.e. a warp has 32 K phonons to choose from.  A pass picks the kind of work wanted by the most
// LANES (a lane wants a kind if any of its slots does - one bit mask per kind per lane, in a register) and every
// such lane executes it for one of its slots.  With K = 4 a lane almost always has a slot that wants to fly, and the
// rare kinds are executed when (nearly) every lane has one waiting.
// ---------------------------------------------------------------------------------------------------------------
enum : int { SF_B1 = 0, SF_B2, SF_DX, SF_DY, SF_TTS, SF_PACKED, SF_CELL, SF_ID, SF_T, SF_R1, SF_R2, SF_MISC, SF_COUNT };
// SF_MISC: [9:0] measurement step relative to the launch, [16:10] impacts since the last scatter, [18:17] edge hit,
//          [31:19] next Philox block of this (phonon, step) stream.  13 bits: a phonon may consume 8191 blocks (scatters,
//          diffuse wall hits, redraws) inside ONE measurement interval; a kernel that sees more drops the phonon and raises
//          the run's `rng budget` error (psim_gpu_synchronize: PSIM_E_RNG) instead of reusing a block - the lock-step
//          kernel, which keeps the counter in a register, has no such limit and is what such a model must be run with.
#define PSIM_MISC_STEP(m) ((m)&1023u)
#define PSIM_MISC_NCOLL(m) (((m) >> 10) & 127u)
#define PSIM_MISC_EDGE(m) (((m) >> 17) & 3u)
#define PSIM_MISC_BLOCK(m) ((m) >> 19)
#define PSIM_MISC_BLOCK_MAX 8191u
#define PSIM_MISC_PACK(step, ncoll, edge, block) ((step) | (min((ncoll), 127u) << 10) | ((edge) << 17) | (min((block), PSIM_MISC_BLOCK_MAX) << 19))

template<int K>
__global__ void __launch_bounds__(kBlock, kSlotBlocks) drift_kernel_slots(const __grid_constant__ LaunchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevParams& P = a.P;
    const uint32_t nst = a.step_end - a.step_begin;
    int32_t* acc_e = reinterpret_cast<int32_t*>(smem_raw);
    long long* acc_f = reinterpret_cast<long long*>(smem_raw + tally_smem_offset_f(nst, P.n_sensors));
    tally_init(a, acc_e, acc_f);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const size_t tally_bytes = tally_staged(a) ? tally_stage_bytes(a.tally_shared, nst, P.n_sensors) : 0;
    uint32_t* sw = reinterpret_cast<uint32_t*>(smem_raw + ((tally_bytes + 127) & ~static_cast<size_t>(127))) +
                   (threadIdx.x >> 5) * (SF_COUNT * K * 32) + lane;
    auto slot_u = [&](int field, uint32_t k) -> uint32_t& { return sw[(field * K + k) * 32]; };
    auto slot_f = [&](int field, uint32_t k) -> float& { return reinterpret_cast<float*>(sw)[(field * K + k) * 32]; };

    const uint32_t w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const uint32_t W = a.n_warps;
    const size_t seg = static_cast<size_t>(w) * a.seg_cap;
    const uint32_t n_in = a.cnt_in[w];
    const uint64_t n_chunks = (a.n_births + 31u) >> 5;
    const uint32_t c0 = (w + W - (a.birth_warp_offset % W)) % W;
    const uint32_t my_chunks = (c0 < n_chunks) ? static_cast<uint32_t>((n_chunks - 1 - c0) / W + 1) : 0u;
    const uint32_t total = n_in + my_chunks * 32u;

    uint32_t next = 0, n_out = 0;
    uint32_t n_steps = 0, n_events = 0, n_absorbed = 0;
    bool overflow = false, rng_over = false;
    uint32_t m_free = (1u << K) - 1u, m_fly = 0, m_wall = 0, m_sct = 0, m_fin = 0;  // which of my slots want what

    for (;;) {
        const bool input = next < total;
        const int c_fly = __popc(__ballot_sync(0xFFFFFFFFu, m_fly != 0u));
        const int c_wall = __popc(__ballot_sync(0xFFFFFFFFu, m_wall != 0u));
        const int c_sct = __popc(__ballot_sync(0xFFFFFFFFu, m_sct != 0u));
        const int c_fin = __popc(__ballot_sync(0xFFFFFFFFu, m_fin != 0u));
        const int c_acq = input ? __popc(__ballot_sync(0xFFFFFFFFu, m_free != 0u)) : 0;
        const int best = max(max(c_fly, c_wall), max(max(c_sct, c_fin), c_acq));
        if (best == 0) { break; }
        // the kind wanted by the most lanes runs; ties go to the rarer kinds (they waited longest to get there)
        if (c_sct == best) {
            // ---- intrinsic scatter
            if (m_sct != 0u) {
                const uint32_t k = __ffs(m_sct) - 1u;
                psim::Phonon p;
                psim::Flight f;
                p.dx = slot_f(SF_DX, k);
                p.dy = slot_f(SF_DY, k);
                p.packed = slot_u(SF_PACKED, k);
                p.cell = slot_u(SF_CELL, k);
                p.id_lo = slot_u(SF_ID, k);
                uint32_t misc = slot_u(SF_MISC, k);
                psim::set_cell_matrix(f, psim::load_cell_matrix(P, p.cell));
                f.sensor_mat = psim::cell_sensor_word(P, p.cell);
                f.vel = psim::phonon_velocity(P, p.packed);
                f.rng.block = PSIM_MISC_BLOCK(misc);
                f.rng.left = 0;
                psim::scatter_event(P, p, f, a.step_begin + PSIM_MISC_STEP(misc));
                const bool spent = f.rng.block > PSIM_MISC_BLOCK_MAX;  // dropped, the run fails (see drift_kernel_queues)
                rng_over |= spent;
                misc = PSIM_MISC_PACK(PSIM_MISC_STEP(misc), 0u, PSIM_MISC_EDGE(misc), f.rng.block);  // impacts since the last scatter := 0
                slot_f(SF_DX, k) = p.dx;
                slot_f(SF_DY, k) = p.dy;
                slot_u(SF_PACKED, k) = p.packed;
                slot_f(SF_TTS, k) = p.tts;
                slot_f(SF_R1, k) = f.r1;
                slot_f(SF_R2, k) = f.r2;
                slot_u(SF_MISC, k) = misc;
                m_sct &= ~(1u << k);
                if (spent) {
                    m_free |= 1u << k;
                } else {
                    m_fly |= 1u << k;
                }
            }
        } else if (c_wall == best) {
            // ---- surface interaction, general case: wall (specular / diffuse), emitting surface, material interface,
            //      transition into a sensor area with other rates, partial edges, stuck-phonon guard
            if (m_wall != 0u) {
                const uint32_t k = __ffs(m_wall) - 1u;
                psim::Phonon p;
                psim::Flight f;
                p.b1 = slot_f(SF_B1, k);
                p.b2 = slot_f(SF_B2, k);
                p.dx = slot_f(SF_DX, k);
                p.dy = slot_f(SF_DY, k);
                p.tts = slot_f(SF_TTS, k);
                p.packed = slot_u(SF_PACKED, k);
                p.cell = slot_u(SF_CELL, k);
                p.id_lo = slot_u(SF_ID, k);
                uint32_t misc = slot_u(SF_MISC, k);
                f.edge = PSIM_MISC_EDGE(misc);
                f.s_hit = psim::edge_coordinate(PSIM_CELL_QUAD(p.cell), f.edge, p.b1, p.b2);
                f.ncoll = PSIM_MISC_NCOLL(misc);
                f.rng.block = PSIM_MISC_BLOCK(misc);
                f.rng.left = 0;
                f.r1 = slot_f(SF_R1, k);
                f.r2 = slot_f(SF_R2, k);
                f.t = 0.f;
                psim::set_cell_matrix(f, psim::load_cell_matrix(P, p.cell));
                f.sensor_mat = psim::cell_sensor_word(P, p.cell);
                f.vel = psim::phonon_velocity(P, p.packed);
                const int ev = psim::impact_event(P, p, f, a.step_begin + PSIM_MISC_STEP(misc));
                const bool spent = f.rng.block > PSIM_MISC_BLOCK_MAX;
                rng_over |= spent;
                m_wall &= ~(1u << k);
                if (ev == psim::EV_DEAD || spent) {
                    ++n_steps;
                    ++n_absorbed;
                    m_free |= 1u << k;
                } else {
                    misc = PSIM_MISC_PACK(PSIM_MISC_STEP(misc), f.ncoll, PSIM_MISC_EDGE(misc), f.rng.block);
                    slot_f(SF_B1, k) = p.b1;
                    slot_f(SF_B2, k) = p.b2;
                    slot_f(SF_DX, k) = p.dx;
                    slot_f(SF_DY, k) = p.dy;
                    slot_f(SF_TTS, k) = p.tts;
                    slot_u(SF_CELL, k) = p.cell;
                    slot_f(SF_R1, k) = f.r1;
                    slot_f(SF_R2, k) = f.r2;
                    slot_u(SF_MISC, k) = misc;
                    m_fly |= 1u << k;
                }
            }
        } else if (c_fin == best) {
            // ---- write-back of phonons that reached the end of the launch window (compacted, coalesced)
            const bool store = m_fin != 0u;
            const unsigned storing = __ballot_sync(0xFFFFFFFFu, store);
            if (store) {
                const uint32_t k = __ffs(m_fin) - 1u;
                const uint32_t slot = n_out + __popc(storing & lt_mask);
                if (slot < a.seg_cap) {
                    a.out_a[seg + slot] = make_float4(slot_f(SF_B1, k), slot_f(SF_B2, k), slot_f(SF_DX, k), slot_f(SF_DY, k));
                    a.out_b[seg + slot] = make_uint4(slot_u(SF_TTS, k), slot_u(SF_PACKED, k), slot_u(SF_CELL, k), slot_u(SF_ID, k));
                } else {
                    overflow = true;
                }
                m_fin &= ~(1u << k);
                m_free |= 1u << k;
            }
            n_out += __popc(storing);
        } else if (c_acq == best) {
            // ---- fetch: the next items of the warp's stream (pool first, then births) into free slots
            const unsigned taking = __ballot_sync(0xFFFFFFFFu, m_free != 0u);
            if (m_free != 0u) {
                const uint32_t k = __ffs(m_free) - 1u;
                const uint32_t idx = next + __popc(taking & lt_mask);
                psim::Phonon p;
                uint32_t s = a.step_begin;
                float t_begin = P.step_time;
                bool got = false;
                if (idx < n_in) {
                    load_phonon(a, seg + idx, p);
                    got = true;
                } else if (idx < total) {
                    const uint32_t b = idx - n_in;
                    const uint64_t item = (static_cast<uint64_t>(c0) + static_cast<uint64_t>(b >> 5) * W) * 32u + (b & 31u);
                    if (item < a.n_births) {
                        t_begin = birth_phonon(a, item, p, s);
                        got = true;
                    }
                }
                if (got) {
                    psim::Flight f;
                    psim::interval_begin(P, p, f, t_begin, s);
                    slot_f(SF_B1, k) = p.b1;
                    slot_f(SF_B2, k) = p.b2;
                    slot_f(SF_DX, k) = p.dx;
                    slot_f(SF_DY, k) = p.dy;
                    slot_f(SF_TTS, k) = p.tts;
                    slot_u(SF_PACKED, k) = p.packed;
                    slot_u(SF_CELL, k) = p.cell;
                    slot_u(SF_ID, k) = p.id_lo;
                    slot_f(SF_T, k) = f.t;
                    slot_f(SF_R1, k) = f.r1;
                    slot_f(SF_R2, k) = f.r2;
                    slot_u(SF_MISC, k) = s - a.step_begin;
                    m_free &= ~(1u << k);
                    m_fly |= 1u << k;
                }
            }
            next += __popc(taking);
        } else {
            // ---- one free-flight segment: to the next edge / scatter / end of the launch window, tallying the
            //      measurement events it crosses on the way
            if (m_fly != 0u) {
                const uint32_t k = __ffs(m_fly) - 1u;
                psim::Phonon p;
                psim::Flight f;
                p.b1 = slot_f(SF_B1, k);
                p.b2 = slot_f(SF_B2, k);
                p.tts = slot_f(SF_TTS, k);
                p.cell = slot_u(SF_CELL, k);  // (its tag tells the flight whether the cell is a triangle or a parallelogram)
                f.t = slot_f(SF_T, k);
                f.r1 = slot_f(SF_R1, k);
                f.r2 = slot_f(SF_R2, k);
                uint32_t misc = slot_u(SF_MISC, k);
                f.edge = 0u;
                f.ncoll = PSIM_MISC_NCOLL(misc);
                f.rng.block = PSIM_MISC_BLOCK(misc);
                const uint32_t s0 = a.step_begin + PSIM_MISC_STEP(misc);
                uint32_t s = s0;
                const int ev = psim::flight_window(P, p, f, s, a.step_end, n_steps, [&](uint32_t k0, uint32_t k1) {
                    const int32_t sg = PSIM_PACK_NEG(slot_u(SF_PACKED, k)) ? -1 : 1;
                    const int32_t fx = psim::flux_fixed(slot_f(SF_DX, k)) * sg, fy = psim::flux_fixed(slot_f(SF_DY, k)) * sg;
                    if (P.lattice) {  // (called with the state at the end of the segment)
                        psim::lattice_runs(P, p.cell, p.b1, p.b2, f.r1, f.r2, f.t, k0, k1,
                                           [&](uint32_t ka, uint32_t kb, uint32_t sn) { tally_range(a, acc_e, acc_f, ka, kb, sn, sg, fx, fy); });
                    } else {
                        const uint32_t sensor = PSIM_CELL_SENSOR(psim::cell_sensor_word(P, slot_u(SF_CELL, k)));
                        tally_range(a, acc_e, acc_f, k0, k1, sensor, sg, fx, fy);
                    }
                });
                ++n_events;
                // a measurement boundary crossed on the way restarts the per-interval bookkeeping (impact counter, Philox
                // block): only the step survives in the packed word; otherwise only the edge changes
                misc = ((s != s0) ? (s - a.step_begin) : (misc & ~(3u << 17))) | (f.edge << 17);
                m_fly &= ~(1u << k);
                if (ev == psim::EV_IMPACT) {
                    // the frequent case - a whole-edge transition into a cell with the same material and rates - is done
                    // at once, and the slot keeps flying; anything else waits for the general surface-interaction kind
                    p.dx = slot_f(SF_DX, k);
                    p.dy = slot_f(SF_DY, k);
                    p.cell = slot_u(SF_CELL, k);
                    f.sensor_mat = psim::cell_sensor_word(P, p.cell);
                    if (psim::fast_impact(P, p, f)) {  // (it may have mirrored the velocity off a specular wall)
                        slot_f(SF_DX, k) = p.dx;
                        slot_f(SF_DY, k) = p.dy;
                        slot_u(SF_CELL, k) = p.cell;
                        slot_f(SF_R1, k) = f.r1;
                        slot_f(SF_R2, k) = f.r2;
                        misc += 1u << 10;  // one more impact since the last scatter (< PSIM_MAX_COLLISIONS here)
                        m_fly |= 1u << k;
                    } else {
                        m_wall |= 1u << k;
                    }
                } else if (ev == psim::EV_SCATTER) {
                    m_sct |= 1u << k;
                } else {
                    m_fin |= 1u << k;
                }
                slot_f(SF_B1, k) = p.b1;
                slot_f(SF_B2, k) = p.b2;
                slot_f(SF_TTS, k) = p.tts;
                slot_f(SF_T, k) = f.t;
                slot_u(SF_MISC, k) = misc;
            }
        }
    }
    n_out = min(n_out, a.seg_cap);
    if (lane == 0) { a.cnt_out[w] = n_out; }
    warp_stats(a, lane, n_steps, n_events, n_absorbed, overflow, rng_over, n_out);
    tally_flush(a, acc_e, acc_f);
}

enum : int { Q_FLY = 0, Q_SCT, Q_WALL, Q_FIN, Q_FREE, Q_COUNT };
__host__ __device__ constexpr uint32_t ring_capacity(int slots) {
    uint32_t c = 32;
    while (c < static_cast<uint32_t>(slots)) { c <<= 1; }
    return c;
}

// Slot storage of the work-queue kernel: three 16-byte groups per slot, [group][slot], so that a pass moves a slot with two
// or three LDS.128 / STS.128 (the first layout, [field][slot] words, took 7-12 scalar accesses per slot and pass: 13 % of
// the warp instructions of the bench job, profiles/r01_summary.md).
//   SG_POS   (b1, b2, r1, r2)       position in the cell frame and its rate of change
//   SG_TIME  (tts, t, misc, cell)   time to the next scatter, time left in the interval, SF_MISC word, current cell
//   SG_VEL   (dx, dy, packed, id)   in-plane velocity, packed word, low id word
enum : int { SG_POS = 0, SG_TIME, SG_VEL, SG_COUNT };
static_assert(SG_COUNT * 4 == SF_COUNT, "both layouts hold the same twelve words per slot");
// TALLY (what a launch does with the measurement events its window records)
enum : int { TALLY_NONE = 0,     // the window ends before the first recorded step: no tally code at all
             TALLY_STAGED = 1,   // per-CTA staging in shared memory (LaunchArgs::tally_shared 1 / 2 / 4) or plain global adds (0)
             TALLY_GLOBAL = 2,   // difference rows in global memory (tally_shared 3), posted warp-cooperatively: tally_post_global
             TALLY_LATTICE = 3 };// the same over the lattice image: the sensor area of every crossed measurement from the phonon's
                                 // position at that instant, the measurements of a pass dealt evenly over the lanes (tally_lattice)
constexpr uint32_t kPostScratchBytes = 64u * 16u;  // per warp: up to two posts per lane

// Shared memory is addressed with 32-bit shared-window addresses and explicit ld.shared / st.shared: with generic pointers
// the compiler rebuilt the window base (S2R SR_CgaCtaId + LEA) at most access sites of this register-tight kernel.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// Many-sensor models (no staging): a flight segment posts +v into the entry of the first recorded step it crossed and -v
// behind the last (difference form, tally_range).  One entry = one 32-byte sector of `a.tally_acc`: int64 (e, fx, fy, -).
// Posted lane by lane that is three REDs to three different words of one sector from the SAME lane, which the L2 sees as
// three sector operations (ncu, round 1: 26 sectors per RED instruction, 38 % of the L2 RED peak at 33 % issue activity).
// Here the warp first compacts its posts into shared memory, then lanes 3 j, 3 j + 1, 3 j + 2 post the three words of one
// entry with ONE RED instruction, which reaches the L2 as one sector (ten posts per instruction; lanes 30, 31 idle).
__device__ __forceinline__ void tally_post_global(const LaunchArgs& a, uint32_t scratch, uint32_t lane, uint32_t lt_mask, bool has,
                                                  uint32_t k0, uint32_t k1, uint32_t sensor, int32_t e, int32_t fx, int32_t fy) {
    const uint32_t S = a.P.n_sensors, F = a.P.first_tally_step;
    const uint32_t r0 = k0 + 1u - F, r1 = k1 + 1u - F;
    const bool has_end = has && r1 < a.P.recorded_steps;
    const unsigned m0 = __ballot_sync(0xFFFFFFFFu, has), m1 = __ballot_sync(0xFFFFFFFFu, has_end);
    if (m0 == 0u) { return; }  // warp-uniform
    const uint32_t n0 = __popc(m0), n = n0 + __popc(m1);
    if (has) { sts128u(scratch + 16u * __popc(m0 & lt_mask), make_uint4(r0 * S + sensor, static_cast<uint32_t>(e), static_cast<uint32_t>(fx), static_cast<uint32_t>(fy))); }
    if (has_end) { sts128u(scratch + 16u * (n0 + __popc(m1 & lt_mask)), make_uint4(r1 * S + sensor, static_cast<uint32_t>(-e), static_cast<uint32_t>(-fx), static_cast<uint32_t>(-fy))); }
    __syncwarp();
    const uint32_t group = (lane * 11u) >> 5, comp = lane - 3u * group;  // lane / 3 and lane % 3 for lane < 32
    if (lane < 30u) {
        for (uint32_t post = group; post < n; post += 10u) {
            const uint32_t entry = lds32(scratch + 16u * post);
            const long long val = static_cast<long long>(static_cast<int32_t>(lds32(scratch + 16u * post + 4u + 4u * comp)));
            atomic_add_i64(a.tally_acc + (static_cast<size_t>(entry) * 4u + comp), val);
        }
    }
    __syncwarp();
}

// Recorded windows over the lattice image (device_types.h).  A flight segment of a lattice cell spans several sensor areas;
// every measurement it crossed is attributed to the area the phonon was in at that instant.  The number of crossed
// measurements differs widely from lane to lane (a phonon that flies along a wire crosses many), so they are not walked
// lane by lane: the (segment, measurement) pairs of the pass are numbered through a warp scan and dealt evenly, 32 per
// round - lane i of a round takes pair i, finds the lane that owns it, fetches that lane's segment by shuffles and
// evaluates the area.  Consecutive pairs of one segment sit in consecutive lanes, so a run of equal areas is recognised by
// comparing with the neighbours: its first pair posts +v into its row, its last -v into the row behind (difference form,
// tally_range); a run cut by the end of a round is posted as two runs, which sums to the same integers.
__device__ __forceinline__ void tally_post_marks(const LaunchArgs& a, uint32_t scratch, uint32_t lane, uint32_t lt_mask, bool plus, bool minus,
                                                 uint32_t k, uint32_t sensor, int32_t e, int32_t fx, int32_t fy) {
    const uint32_t S = a.P.n_sensors, F = a.P.first_tally_step;
    const uint32_t r0 = k + 1u - F, r1 = r0 + 1u;  // the row of the measurement that ended step k, and the one behind it
    const bool has_end = minus && r1 < a.P.recorded_steps;
    const unsigned m0 = __ballot_sync(0xFFFFFFFFu, plus), m1 = __ballot_sync(0xFFFFFFFFu, has_end);
    if ((m0 | m1) == 0u) { return; }  // warp-uniform
    const uint32_t n0 = __popc(m0), n = n0 + __popc(m1);
    if (plus) { sts128u(scratch + 16u * __popc(m0 & lt_mask), make_uint4(r0 * S + sensor, static_cast<uint32_t>(e), static_cast<uint32_t>(fx), static_cast<uint32_t>(fy))); }
    if (has_end) { sts128u(scratch + 16u * (n0 + __popc(m1 & lt_mask)), make_uint4(r1 * S + sensor, static_cast<uint32_t>(-e), static_cast<uint32_t>(-fx), static_cast<uint32_t>(-fy))); }
    __syncwarp();
    const uint32_t group = (lane * 11u) >> 5, comp = lane - 3u * group;  // lane / 3 and lane % 3 for lane < 32
    if (lane < 30u) {
        for (uint32_t post = group; post < n; post += 10u) {
            const uint32_t entry = lds32(scratch + 16u * post);
            const long long val = static_cast<long long>(static_cast<int32_t>(lds32(scratch + 16u * post + 4u + 4u * comp)));
            atomic_add_i64(a.tally_acc + (static_cast<size_t>(entry) * 4u + comp), val);
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void tally_lattice(const LaunchArgs& a, uint32_t scratch, uint32_t lane, uint32_t lt_mask, bool has, uint32_t k0, uint32_t k1,
                                              uint32_t cell, float e1, float e2, float r1, float r2, float t_left, int32_t sg, int32_t fx, int32_t fy) {
    const DevParams& P = a.P;
    const uint32_t count = has ? k1 - k0 : 0u;
    uint32_t incl = count;  // inclusive scan over the lanes
#pragma unroll
    for (uint32_t d = 1u; d < 32u; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) { incl += t; }
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (total == 0u) { return; }  // warp-uniform
    const uint2 sub = has ? psim::load_cell_tris(P.cells, PSIM_CELL_INDEX(cell)) : make_uint2(0u, 1u | (1u << 16));
    const uint32_t first = incl - count;  // number of this lane's first pair
    for (uint32_t base = 0u; base < total; base += 32u) {
        const bool active = base + lane < total;
        const uint32_t i = active ? base + lane : total - 1u;
        // owner = the number of lanes whose pairs all lie before pair i
        uint32_t owner = 0u;
#pragma unroll
        for (uint32_t step = 16u; step > 0u; step >>= 1) {
            const uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, owner + step - 1u);
            if (v <= i) { owner += step; }
        }
        const uint32_t k = __shfl_sync(0xFFFFFFFFu, k0, owner) + (i - __shfl_sync(0xFFFFFFFFu, first, owner));
        const uint32_t o_k1 = __shfl_sync(0xFFFFFFFFu, k1, owner);
        const uint2 o_sub = make_uint2(__shfl_sync(0xFFFFFFFFu, sub.x, owner), __shfl_sync(0xFFFFFFFFu, sub.y, owner));
        const float o_e1 = __shfl_sync(0xFFFFFFFFu, e1, owner), o_e2 = __shfl_sync(0xFFFFFFFFu, e2, owner);
        const float o_r1 = __shfl_sync(0xFFFFFFFFu, r1, owner), o_r2 = __shfl_sync(0xFFFFFFFFu, r2, owner);
        const float o_tl = __shfl_sync(0xFFFFFFFFu, t_left, owner);
        const int32_t o_sg = __shfl_sync(0xFFFFFFFFu, sg, owner), o_fx = __shfl_sync(0xFFFFFFFFu, fx, owner), o_fy = __shfl_sync(0xFFFFFFFFu, fy, owner);
        const uint32_t sensor = psim::lattice_sensor_at(P, o_sub, o_e1, o_e2, o_r1, o_r2, psim::lattice_back(P, o_tl, k, o_k1));
        // a run begins where the lane before holds another segment or another area (or nothing: lane 0), and ends likewise
        const uint32_t p_owner = __shfl_up_sync(0xFFFFFFFFu, owner, 1), p_sensor = __shfl_up_sync(0xFFFFFFFFu, sensor, 1);
        const uint32_t n_owner = __shfl_down_sync(0xFFFFFFFFu, owner, 1), n_sensor = __shfl_down_sync(0xFFFFFFFFu, sensor, 1);
        const bool n_active = __shfl_down_sync(0xFFFFFFFFu, active ? 1u : 0u, 1) != 0u;
        const bool begins = active && (lane == 0u || p_owner != owner || p_sensor != sensor);
        const bool ends = active && (lane == 31u || !n_active || n_owner != owner || n_sensor != sensor);
        tally_post_marks(a, scratch, lane, lt_mask, begins, ends, k, sensor, o_sg, o_fx, o_fy);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// drift_kernel_queues<NS, TALLY>: a warp owns NS slots in shared memory and ONE QUEUE OF SLOT NUMBERS PER KIND of work; a pass
// pops up to 32 entries of the fullest queue - lane i takes the i-th - so every kind runs with all 32 lanes as soon as 32
// slots want it, and rare kinds wait in their queue until then.  Queue heads and counts are warp-uniform registers: choosing
// the kind costs no ballot.  (The version before, drift_kernel_slots, bound K slots to each lane: a pass ran with the lanes
// that happened to hold a slot of the chosen kind, 19.7 of 32 on average against 26 here.)
//
// Every kind of work except the write-back is followed by a free flight of the same phonon, so every pass ENDS with the
// flight segment of its lanes ("fetch, then fly", "scatter, then fly", "wall, then fly") instead of sending them through the
// flight queue: the state is still in registers, and a pass's fixed costs - choosing the queue, pop, slot loads / stores,
// the push - are paid once per two events.  Only a pass that runs with at least kFuseLanes lanes does so: a sparse pass (a
// rare kind on a mesh with few slots per warp) hands its slots to the flight queue, where they fly in a full pass (fusing
// unconditionally cost 10 % more instructions on the kinked wire, whose scatter passes run with 12 lanes on average).
// ---------------------------------------------------------------------------------------------------------------
#ifndef PSIM_FUSE_LANES
#define PSIM_FUSE_LANES 24
#endif
constexpr uint32_t kFuseLanes = PSIM_FUSE_LANES;

template<int NS, int TALLY>
__global__ void __launch_bounds__(kBlock, kSlotBlocks) drift_kernel_queues(const __grid_constant__ LaunchArgs a) {
    static_assert(NS <= 256 && NS >= 32, "slot numbers are bytes");
    constexpr uint32_t QC = ring_capacity(NS);  // the queues are power-of-two rings that can hold every slot
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevParams& P = a.P;
    const uint32_t nst = a.step_end - a.step_begin;
    int32_t* acc_e = reinterpret_cast<int32_t*>(smem_raw);
    long long* acc_f = reinterpret_cast<long long*>(smem_raw + tally_smem_offset_f(nst, P.n_sensors));
    if (TALLY == TALLY_STAGED) { tally_init(a, acc_e, acc_f); }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const size_t tally_bytes = (TALLY == TALLY_STAGED && tally_staged(a)) ? tally_stage_bytes(a.tally_shared, nst, P.n_sensors) : 0;
    const uint32_t base = smem_u32(smem_raw) + static_cast<uint32_t>((tally_bytes + 127) & ~static_cast<size_t>(127));
    const uint32_t sv = base + warp * (SG_COUNT * NS * 16u);                                                  // slot groups
    const uint32_t qb = base + kWarpsPerBlock * (SG_COUNT * NS * 16u) + warp * (Q_COUNT * QC);                // queue rings (bytes)
    const uint32_t post = base + kWarpsPerBlock * (SG_COUNT * NS * 16u + Q_COUNT * QC) + warp * kPostScratchBytes;  // tally_post_global
    auto grp = [&](int g, uint32_t k) -> uint32_t { return sv + (g * NS + k) * 16u; };
    struct Queue {
        uint32_t head, count;
    };
    Queue q_fly{ 0, 0 }, q_sct{ 0, 0 }, q_wall{ 0, 0 }, q_fin{ 0, 0 }, q_free{ 0, NS };
    // lane i takes the i-th of the first `take` entries
    auto pop = [&](Queue& q, int qi, uint32_t take, bool& active) -> uint32_t {
        active = lane < take;
        const uint32_t k = active ? lds8(qb + qi * QC + ((q.head + lane) & (QC - 1u))) : 0u;
        q.head += take;
        q.count -= take;
        return k;
    };
    for (uint32_t i = lane; i < NS; i += 32u) { sts8(qb + Q_FREE * QC + i, i); }

    const uint32_t w = blockIdx.x * kWarpsPerBlock + warp;
    const uint32_t W = a.n_warps;
    const size_t seg = static_cast<size_t>(w) * a.seg_cap;
    const uint32_t n_in = a.cnt_in[w];
    const uint64_t n_chunks = (a.n_births + 31u) >> 5;
    const uint32_t c0 = (w + W - (a.birth_warp_offset % W)) % W;
    const uint32_t my_chunks = (c0 < n_chunks) ? static_cast<uint32_t>((n_chunks - 1 - c0) / W + 1) : 0u;
    const uint32_t total = n_in + my_chunks * 32u;

    uint32_t next = 0, n_out = 0;
    uint32_t n_steps = 0, n_events = 0, n_absorbed = 0;
    bool overflow = false, rng_over = false;

    // One free-flight segment of the phonon in slot k, whose state the caller holds in registers (`have_vel`: including the
    // SG_VEL group, which the caller then also wants stored): to the next edge / scatter / end of the launch window.  The
    // frequent kind of impact - a whole-edge transition into a cell with the same material and rates - is done at once and
    // the slot keeps flying; the recorded measurement events crossed on the way are tallied after the segment, when the warp
    // has converged again.  Returns the queue the slot goes to next.  Called by all 32 lanes.
    auto fly = [&](bool act, uint32_t k, psim::Phonon& p, psim::Flight& f, uint32_t misc, const bool have_vel) -> int {
        int dest = -1;
        uint32_t tk0 = 0, tk1 = 0, sensor = 0;  // recorded steps [tk0, tk1) that ended during this segment
        int32_t fx = 0, fy = 0;                 // ... and what the segment adds to each of them (sign * fixed-point velocity)
        float le1 = 0.f, le2 = 0.f, lr1 = 0.f, lr2 = 0.f, ltl = 0.f;  // TALLY_LATTICE: the end of the segment before any impact
        uint32_t lcell = 0;
        if (act) {
            f.edge = 0u;
            f.ncoll = PSIM_MISC_NCOLL(misc);
            f.rng.block = PSIM_MISC_BLOCK(misc);
            const uint32_t s0 = a.step_begin + PSIM_MISC_STEP(misc);
            uint32_t s = s0;
            const uint32_t cell0 = p.cell;
            const int ev = psim::flight_window(P, p, f, s, a.step_end, n_steps, [&](uint32_t k0, uint32_t k1) {
                if (TALLY != TALLY_NONE) {
                    tk0 = k0;
                    tk1 = k1;
                }
            });
            ++n_events;
            // a measurement boundary crossed on the way restarts the per-interval bookkeeping (impact counter, Philox
            // block): only the step survives in the packed word; otherwise only the edge changes
            misc = ((s != s0) ? (s - a.step_begin) : (misc & ~(3u << 17))) | (f.edge << 17);
            const bool hit = ev == psim::EV_IMPACT;
            bool reflected = false;
            if (hit || tk1 > tk0) {
                const uint2 tail0 = psim::load_cell_tail(P.cells, PSIM_CELL_INDEX(cell0));  // of the cell the segment was flown in
                if (!have_vel) {
                    const float4 g2 = lds128(grp(SG_VEL, k));
                    p.dx = g2.x;
                    p.dy = g2.y;
                    p.packed = __float_as_uint(g2.z);
                }
                sensor = PSIM_CELL_SENSOR(tail0.x);
                if (TALLY != TALLY_NONE) {  // with the velocity the segment was flown at: the impact may mirror it
                    fx = psim::flux_fixed(p.dx);
                    fy = psim::flux_fixed(p.dy);
                }
                if (TALLY == TALLY_LATTICE) {
                    le1 = p.b1, le2 = p.b2, lr1 = f.r1, lr2 = f.r2, ltl = f.t;
                    lcell = cell0;
                }
                if (hit) {
                    f.sensor_mat = tail0.x;
                    // (an image without any edge the fast path could take - DevParams::fast_links - does not load the links for it)
                if ((P.fast_links | PSIM_FAST_WALLS) && psim::fast_impact(P, p, f, psim::load_cell_links(P.cells, PSIM_CELL_INDEX(cell0)), tail0, reflected, false)) {
                        misc += 1u << 10;  // one more impact since the last scatter (< PSIM_MAX_COLLISIONS here)
                        dest = Q_FLY;
                    } else {
                        dest = Q_WALL;
                    }
                }
            }
            if (!hit) { dest = (ev == psim::EV_SCATTER) ? Q_SCT : Q_FIN; }
            sts128(grp(SG_POS, k), make_float4(p.b1, p.b2, f.r1, f.r2));
            sts128(grp(SG_TIME, k), make_float4(p.tts, f.t, __uint_as_float(misc), __uint_as_float(p.cell)));
            if (have_vel) {
                sts128(grp(SG_VEL, k), make_float4(p.dx, p.dy, __uint_as_float(p.packed), __uint_as_float(p.id_lo)));
            } else if (reflected) {
                sts64(grp(SG_VEL, k), p.dx, p.dy);
            }
        }
        if (TALLY != TALLY_NONE) {
            const bool has = tk1 > tk0;
            const int32_t sg = PSIM_PACK_NEG(p.packed) ? -1 : 1;
            fx *= sg;
            fy *= sg;
            if (TALLY == TALLY_LATTICE) {
                tally_lattice(a, post, lane, lt_mask, has, tk0, tk1, lcell, le1, le2, lr1, lr2, ltl, sg, fx, fy);
            } else if (TALLY == TALLY_GLOBAL) {
                tally_post_global(a, post, lane, lt_mask, has, tk0, tk1, sensor, sg, fx, fy);
            } else if (has) {
                tally_range(a, acc_e, acc_f, tk0, tk1, sensor, sg, fx, fy);
            }
        }
        return dest;
    };

    // a sparse pass: the state goes back to the slot and the slot to the flight queue
    auto park = [&](bool act, uint32_t k, const psim::Phonon& p, const psim::Flight& f, uint32_t misc) -> int {
        if (act) {
            sts128(grp(SG_POS, k), make_float4(p.b1, p.b2, f.r1, f.r2));
            sts128(grp(SG_TIME, k), make_float4(p.tts, f.t, __uint_as_float(misc), __uint_as_float(p.cell)));
            sts128(grp(SG_VEL, k), make_float4(p.dx, p.dy, __uint_as_float(p.packed), __uint_as_float(p.id_lo)));
        }
        return act ? Q_FLY : -1;
    };

    for (;;) {
        __syncwarp();  // slot words and queue entries written by other lanes in the previous pass
        const uint32_t c_fly = min(q_fly.count, 32u), c_wall = min(q_wall.count, 32u), c_sct = min(q_sct.count, 32u);
        const uint32_t c_fin = min(q_fin.count, 32u), c_acq = min(min(q_free.count, total - next), 32u);
        const uint32_t best = max(max(c_fly, c_wall), max(max(c_sct, c_fin), c_acq));
        if (best == 0u) { break; }
        const bool fuse = best >= kFuseLanes;  // warp-uniform
        // Stage 1 - what the chosen kind does with its slots; it leaves the lanes that go on to fly with their whole state in
        // registers (`go`).  Stage 2 - ONE copy of the flight segment for all kinds (four inlined copies cost more in
        // instruction-cache misses than they saved: ncu stall no_instruction 0.32 -> 1.13 per issue).  Stage 3 - the push.
        bool act, go = false, have_vel = true, flies = fuse;
        uint32_t k, misc = 0;
        int dest = -1;
        psim::Phonon p;
        psim::Flight f;
        p.dx = p.dy = 0.f;
        p.packed = 0u;
        // the fullest queue runs; ties go to the rarer kinds (they waited longest to get there)
        if (c_sct == best) {
            // ---- intrinsic scatter
            k = pop(q_sct, Q_SCT, c_sct, act);
            if (act) {
                const float4 g0 = lds128(grp(SG_POS, k)), g1 = lds128(grp(SG_TIME, k)), g2 = lds128(grp(SG_VEL, k));
                p.b1 = g0.x;
                p.b2 = g0.y;
                f.t = g1.y;
                misc = __float_as_uint(g1.z);
                p.cell = __float_as_uint(g1.w);
                p.dx = g2.x;
                p.dy = g2.y;
                p.packed = __float_as_uint(g2.z);
                p.id_lo = __float_as_uint(g2.w);
                psim::set_cell_matrix(f, psim::load_cell_matrix(P, p.cell));
                f.sensor_mat = psim::cell_sensor_word(P, p.cell);
                f.vel = psim::phonon_velocity(P, p.packed);
                f.rng.block = PSIM_MISC_BLOCK(misc);
                f.rng.left = 0;
                psim::scatter_event(P, p, f, a.step_begin + PSIM_MISC_STEP(misc));
                // A phonon that has used up the Philox blocks of this interval is DROPPED and the run fails (PSIM_E_RNG): were it
                // to go on with a frozen block it would redraw the same time to scatter for ever - and if that time is below the
                // fp32 resolution of the time left in a long interval, the launch would never end.
                go = f.rng.block <= PSIM_MISC_BLOCK_MAX;
                rng_over |= !go;
                misc = PSIM_MISC_PACK(PSIM_MISC_STEP(misc), 0u, PSIM_MISC_EDGE(misc), f.rng.block);  // impacts since the last scatter := 0
                if (!go) { dest = Q_FREE; }
            }
        } else if (c_wall == best) {
            // ---- surface interaction, general case: wall (specular / diffuse), emitting surface, material interface,
            //      transition into a sensor area with other rates, partial edges, stuck-phonon guard
            k = pop(q_wall, Q_WALL, c_wall, act);
            if (act) {
                const float4 g0 = lds128(grp(SG_POS, k)), g1 = lds128(grp(SG_TIME, k)), g2 = lds128(grp(SG_VEL, k));
                p.b1 = g0.x;
                p.b2 = g0.y;
                f.r1 = g0.z;
                f.r2 = g0.w;
                p.tts = g1.x;
                misc = __float_as_uint(g1.z);
                p.cell = __float_as_uint(g1.w);
                p.dx = g2.x;
                p.dy = g2.y;
                p.packed = __float_as_uint(g2.z);
                p.id_lo = __float_as_uint(g2.w);
                f.edge = PSIM_MISC_EDGE(misc);
                f.s_hit = psim::edge_coordinate(PSIM_CELL_QUAD(p.cell), f.edge, p.b1, p.b2);
                f.ncoll = PSIM_MISC_NCOLL(misc);
                f.rng.block = PSIM_MISC_BLOCK(misc);
                f.rng.left = 0;
                f.t = g1.y;
                psim::set_cell_matrix(f, psim::load_cell_matrix(P, p.cell));
                f.sensor_mat = psim::cell_sensor_word(P, p.cell);
                f.vel = psim::phonon_velocity(P, p.packed);
                const int ev = psim::impact_event(P, p, f, a.step_begin + PSIM_MISC_STEP(misc));
                const bool spent = f.rng.block > PSIM_MISC_BLOCK_MAX;  // dropped, see the scatter pass
                rng_over |= spent;
                if (ev == psim::EV_DEAD || spent) {
                    ++n_steps;
                    ++n_absorbed;
                    dest = Q_FREE;
                } else {
                    go = true;
                    misc = PSIM_MISC_PACK(PSIM_MISC_STEP(misc), f.ncoll, PSIM_MISC_EDGE(misc), f.rng.block);
                }
            }
        } else if (c_fin == best) {
            // ---- write-back of phonons that reached the end of the launch window (compacted, coalesced)
            k = pop(q_fin, Q_FIN, c_fin, act);
            if (act) {
                const uint32_t slot = n_out + lane;
                if (slot < a.seg_cap) {
                    const float4 g0 = lds128(grp(SG_POS, k)), g1 = lds128(grp(SG_TIME, k)), g2 = lds128(grp(SG_VEL, k));
                    a.out_a[seg + slot] = make_float4(g0.x, g0.y, g2.x, g2.y);
                    a.out_b[seg + slot] = make_uint4(__float_as_uint(g1.x), __float_as_uint(g2.z), __float_as_uint(g1.w), __float_as_uint(g2.w));
                } else {
                    overflow = true;
                }
                dest = Q_FREE;
            }
            n_out += c_fin;
            flies = false;
        } else if (c_acq == best) {
            // ---- fetch: the next items of the warp's stream (pool first, then births) into free slots
            k = pop(q_free, Q_FREE, c_acq, act);
            if (act) {
                const uint32_t idx = next + lane;
                uint32_t s = a.step_begin;
                float t_begin = P.step_time;
                if (idx < n_in) {
                    load_phonon(a, seg + idx, p);
                    go = true;
                } else {
                    const uint32_t b = idx - n_in;
                    const uint64_t item = (static_cast<uint64_t>(c0) + static_cast<uint64_t>(b >> 5) * W) * 32u + (b & 31u);
                    if (item < a.n_births) {
                        t_begin = birth_phonon(a, item, p, s);
                        go = true;
                    }
                }
                if (go) {
                    psim::interval_begin(P, p, f, t_begin, s);
                    misc = s - a.step_begin;
                } else {
                    dest = Q_FREE;
                }
            }
            next += c_acq;
        } else {
            // ---- fly: slots that entered a neighbour cell (or were parked by a sparse pass)
            k = pop(q_fly, Q_FLY, c_fly, act);
            if (act) {
                const float4 g0 = lds128(grp(SG_POS, k)), g1 = lds128(grp(SG_TIME, k));
                p.b1 = g0.x;
                p.b2 = g0.y;
                f.r1 = g0.z;
                f.r2 = g0.w;
                p.tts = g1.x;
                f.t = g1.y;
                misc = __float_as_uint(g1.z);
                p.cell = __float_as_uint(g1.w);
                go = true;
            }
            have_vel = false;
            flies = true;
        }
        if (flies) {
            const int next_kind = fly(go, k, p, f, misc, have_vel);
            if (go) { dest = next_kind; }
        } else if (go) {
            dest = park(true, k, p, f, misc);  // a sparse pass: its slots fly with a full flight pass
        }
        // five-way push with one address computation (selects on warp-uniform values, no branches)
        const bool d_fly = dest == Q_FLY, d_sct = dest == Q_SCT, d_wall = dest == Q_WALL, d_fin = dest == Q_FIN, d_free = dest == Q_FREE;
        const unsigned m_fly = __ballot_sync(0xFFFFFFFFu, d_fly), m_sct = __ballot_sync(0xFFFFFFFFu, d_sct);
        const unsigned m_wall = __ballot_sync(0xFFFFFFFFu, d_wall), m_fin = __ballot_sync(0xFFFFFFFFu, d_fin);
        const unsigned m_free = __ballot_sync(0xFFFFFFFFu, d_free);
        const uint32_t n_fly = __popc(m_fly), n_sct = __popc(m_sct), n_wall = __popc(m_wall), n_fin = __popc(m_fin), n_free = __popc(m_free);
        const uint32_t t_fly = q_fly.head + q_fly.count;
        const uint32_t t_sct = q_sct.head + q_sct.count, t_wall = q_wall.head + q_wall.count, t_fin = q_fin.head + q_fin.count;
        const uint32_t t_free = q_free.head + q_free.count;
        const unsigned peers = d_fly ? m_fly : (d_sct ? m_sct : (d_wall ? m_wall : (d_fin ? m_fin : m_free)));
        const uint32_t tail = d_fly ? t_fly : (d_sct ? t_sct : (d_wall ? t_wall : (d_fin ? t_fin : t_free)));
        if (dest >= 0) { sts8(qb + dest * QC + ((tail + __popc(peers & lt_mask)) & (QC - 1u)), k); }
        q_fly.count += n_fly;
        q_sct.count += n_sct;
        q_wall.count += n_wall;
        q_fin.count += n_fin;
        q_free.count += n_free;
    }
    n_out = min(n_out, a.seg_cap);
    if (lane == 0) { a.cnt_out[w] = n_out; }
    warp_stats(a, lane, n_steps, n_events, n_absorbed, overflow, rng_over, n_out);
    if (TALLY == TALLY_STAGED) { tally_flush(a, acc_e, acc_f); }
}

// First version: tiles of 32 phonons in lock step (every lane waits for the slowest phonon of its tile).
__global__ void __launch_bounds__(kBlock, (kBlock <= 256 ? 2 : 1)) drift_kernel_lockstep(const __grid_constant__ LaunchArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const DevParams& P = a.P;
    int32_t* acc_e = reinterpret_cast<int32_t*>(smem_raw);
    long long* acc_f = reinterpret_cast<long long*>(smem_raw + tally_smem_offset_f(a.step_end - a.step_begin, P.n_sensors));
    tally_init(a, acc_e, acc_f);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const uint32_t W = a.n_warps;
    const size_t seg = static_cast<size_t>(w) * a.seg_cap;
    const uint32_t n_in = a.cnt_in[w];
    const uint32_t pool_tiles = (n_in + 31u) >> 5;
    const uint64_t n_chunks = (a.n_births + 31u) >> 5;
    uint64_t chunk = (w + W - (a.birth_warp_offset % W)) % W;
    uint32_t n_out = 0, n_steps = 0, n_events = 0, n_absorbed = 0;
    bool overflow = false;
    for (uint32_t tile = 0;; ++tile) {
        const bool from_pool = tile < pool_tiles;  // warp-uniform
        if (!from_pool && chunk >= n_chunks) { break; }
        psim::Phonon p;
        float t_first = P.step_time;
        uint32_t start = a.step_begin;
        bool alive = false;
        if (from_pool) {
            const uint32_t idx = tile * 32u + lane;
            if (idx < n_in) {
                load_phonon(a, seg + idx, p);
                alive = true;
            }
        } else {
            const uint64_t item = chunk * 32u + lane;
            if (item < a.n_births) {
                t_first = birth_phonon(a, item, p, start);
                alive = true;
            }
            chunk += W;
        }
        if (alive) {
            alive = psim::advance_window(P, p, t_first, start, a.step_end, n_steps, n_events,
                                         [&](uint32_t k0, uint32_t k1, const psim::Phonon& q, const psim::Flight& f) {
                const int32_t sg = PSIM_PACK_NEG(q.packed) ? -1 : 1;
                const int32_t fx = psim::flux_fixed(q.dx) * sg, fy = psim::flux_fixed(q.dy) * sg;
                if (P.lattice) {
                    psim::lattice_runs(P, q.cell, q.b1, q.b2, f.r1, f.r2, f.t, k0, k1,
                                       [&](uint32_t ka, uint32_t kb, uint32_t sn) { tally_range(a, acc_e, acc_f, ka, kb, sn, sg, fx, fy); });
                } else {
                    tally_range(a, acc_e, acc_f, k0, k1, PSIM_CELL_SENSOR(f.sensor_mat), sg, fx, fy);
                }
            });
            if (!alive) { ++n_absorbed; }
        }
        const unsigned m = __ballot_sync(0xFFFFFFFFu, alive);
        if (alive) {
            const uint32_t slot = n_out + __popc(m & ((1u << lane) - 1u));
            if (slot < a.seg_cap) {
                store_phonon(a, seg + slot, p);
            } else {
                overflow = true;
            }
        }
        n_out += __popc(m);
    }
    n_out = min(n_out, a.seg_cap);
    if (lane == 0) { a.cnt_out[w] = n_out; }
    warp_stats(a, lane, n_steps, n_events, n_absorbed, overflow, false, n_out);
    tally_flush(a, acc_e, acc_f);
}

// Rows [row_begin, row_end) of the difference-form accumulator (tally_shared == 3) become running sums in the caller-visible
// tally arrays; `carry` holds the sums at row_begin - 1 and is private to the library, so callers may all-reduce finished
// rows in place.
__global__ void finalize_rows_kernel(const long long* acc, int32_t* tally_e, long long* tally_f, int32_t* carry_e, long long* carry_f,
                                     uint32_t S, uint32_t row_begin, uint32_t row_end) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // sensor * 3 + component
    if (i >= 3u * S) { return; }
    const uint32_t s = i / 3u, c = i % 3u;
    long long v = (c == 0u) ? static_cast<long long>(carry_e[s]) : carry_f[2u * s + c - 1u];
    for (uint32_t r = row_begin; r < row_end; ++r) {
        const size_t k = static_cast<size_t>(r) * S + s;
        v += acc[4u * k + c];
        if (c == 0u) {
            tally_e[k] = static_cast<int32_t>(v);
        } else {
            tally_f[2u * k + c - 1u] = v;
        }
    }
    if (c == 0u) {
        carry_e[s] = static_cast<int32_t>(v);
    } else {
        carry_f[2u * s + c - 1u] = v;
    }
}

// device [R][S] tallies -> the caller's [S][R] layout (psim_gpu_get_tallies), through a 32 x 33 shared-memory tile so that both
// the reads and the writes are coalesced; flux also converted from fixed point to m/s
__global__ void transpose_tallies_kernel(const int32_t* tally_e, const long long* tally_f, uint32_t R, uint32_t S, double scale,
                                         int32_t* out_e, double* out_f, long long* out_x) {
    __shared__ int32_t te[32][33];
    __shared__ long long tx[32][33], ty[32][33];
    const uint32_t s0 = blockIdx.x * 32u, r0 = blockIdx.y * 32u;
    for (uint32_t j = threadIdx.y; j < 32u; j += blockDim.y) {
        const uint32_t r = r0 + j, s = s0 + threadIdx.x;
        if (r < R && s < S) {
            const size_t k = static_cast<size_t>(r) * S + s;
            te[j][threadIdx.x] = tally_e[k];
            tx[j][threadIdx.x] = tally_f[2 * k];
            ty[j][threadIdx.x] = tally_f[2 * k + 1];
        }
    }
    __syncthreads();
    for (uint32_t j = threadIdx.y; j < 32u; j += blockDim.y) {
        const uint32_t s = s0 + j, r = r0 + threadIdx.x;
        if (r < R && s < S) {
            const size_t k = static_cast<size_t>(s) * R + r;
            if (out_e) { out_e[k] = te[threadIdx.x][j]; }
            if (out_f) {
                out_f[2 * k] = static_cast<double>(tx[threadIdx.x][j]) * scale;
                out_f[2 * k + 1] = static_cast<double>(ty[threadIdx.x][j]) * scale;
            }
            if (out_x) {
                out_x[2 * k] = tx[threadIdx.x][j];
                out_x[2 * k + 1] = ty[threadIdx.x][j];
            }
        }
    }
}

// phonons per MODEL cell: a flight cell that is a parallelogram holds one model triangle on either side of its diagonal
// (lattice_cells != NULL: the pool is in lattice coordinates; P is always the fine image)
__global__ void cell_histogram_kernel(DevParams P, const float4* pool_a, const uint4* pool_b, const uint32_t* cnt, uint32_t seg_cap,
                                      uint32_t n_warps, unsigned long long* hist, const DevCell* lattice_cells, const uint32_t* lattice_sub_fine) {
    const uint32_t w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (w >= n_warps) { return; }
    const uint32_t n = cnt[w];
    for (uint32_t i = threadIdx.x & 31u; i < n; i += 32u) {
        const size_t k = static_cast<size_t>(w) * seg_cap + i;
        float4 a = pool_a[k];
        uint32_t cell = pool_b[k].z;
        if (lattice_cells != nullptr) { psim::coarse_to_fine(lattice_cells, lattice_sub_fine, cell, a.x, a.y); }
        atomicAdd(&hist[psim::api_cell_of(P, cell, a.x, a.y)], 1ull);
    }
}

// per-function probes for the parity tests
__global__ void probe_sample_kernel(DevParams P, uint32_t table, const float* u1, const float* u2, size_t n,
                                    uint32_t* out_bin, uint32_t* out_ta, uint32_t* out_bin_plain) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) { return; }
    const float2* t = P.tables + static_cast<size_t>(table) * PSIM_BINS;
    const uint32_t bin = psim::sample_bin(P, table, u1[i]);   // guided search, as the flight loop does it
    out_bin[i] = bin;
    out_bin_plain[i] = psim::bisect_table(t, u1[i]);          // the reference's plain bisection
    out_ta[i] = (u2[i] <= t[bin].y) ? 0u : 1u;
}

__global__ void probe_flight_kernel(DevParams P, const uint32_t* cell, const float* in, size_t n, float* out) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) { return; }
    psim::Phonon p;
    psim::Flight f;
    p.b1 = in[4 * i];
    p.b2 = in[4 * i + 1];
    const float vx = in[4 * i + 2], vy = in[4 * i + 3];
    const float speed = sqrtf(vx * vx + vy * vy);
    p.dx = vx;  // the state carries the velocity vector
    p.dy = vy;
    p.tts = psim::f_inf();
    p.cell = cell[i];
    p.packed = 0u;
    p.id_lo = 0u;
    psim::set_cell_matrix(f, psim::load_cell_matrix(P, p.cell));
    f.vel = speed;
    psim::update_rates_of_motion(f, p);
    f.t = 1.0e30f;  // no measurement boundary in the way
    f.ncoll = 0;
    psim::rng_begin(f.rng);
    uint32_t s = 0, n_steps = 0;
    const float tts0 = 1.0e30f;
    p.tts = tts0;
    const float b1_start = p.b1, b2_start = p.b2, r1_rate = f.r1, r2_rate = f.r2;
    const int ev = psim::flight_window(P, p, f, s, 1u, n_steps, [](uint32_t, uint32_t) {});
    // elapsed time from the displacement along the faster barycentric axis (tts0 - p.tts has no digits left)
    const float t_hit = (fabsf(r1_rate) > fabsf(r2_rate)) ? (p.b1 - b1_start) / r1_rate : (p.b2 - b2_start) / r2_rate;
    float dx = p.dx, dy = p.dy;
    if (ev == psim::EV_IMPACT) {
        const float2 nrm = psim::load_cell_normal(P, p.cell, f.edge);
        const float dn = dx * nrm.x + dy * nrm.y;
        dx -= 2.f * dn * nrm.x;
        dy -= 2.f * dn * nrm.y;
    }
    out[6 * i] = (ev == psim::EV_IMPACT) ? static_cast<float>(f.edge) : -1.f;
    out[6 * i + 1] = t_hit;
    out[6 * i + 2] = p.b1;
    out[6 * i + 3] = p.b2;
    out[6 * i + 4] = dx / speed;  // unit direction after a specular reflection
    out[6 * i + 5] = dy / speed;
}

__global__ void probe_rates_kernel(DevParams P, uint32_t sensor, const float* w, const uint32_t* ta, size_t n, float* out) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) { return; }
    const DevSensor s = psim::load_sensor(P.sensors, sensor);
    float rn, ru, ri;
    psim::relax_rates(s, w[i], ta[i], rn, ru, ri);
    out[3 * i] = rn;
    out[3 * i + 1] = ru;
    out[3 * i + 2] = ri;
}

}  // namespace
#endif
