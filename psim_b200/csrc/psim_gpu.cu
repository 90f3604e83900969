// sm_100a kernels and the C ABI of include/psim_b200.h.
//
// Data layout in HBM
//   phonon pool   two ping-pong copies of  A: float4[W][cap] (b1, b2, vx, vy)   B: uint4[W][cap] (tts, packed, cell, id)
//                 W = number of resident warps of the drift kernel.  Warp w owns segment w of both copies:
//                 it streams its segment of the input copy (two fully coalesced 16-byte loads per lane),
//                 advances every phonon across the measurement intervals of this launch, and appends the
//                 survivors - and the phonons it creates - to its segment of the output copy.  Compaction is
//                 therefore a warp ballot + popcount: no block barrier, no global atomic, no holes.
//   tallies       energy int32[R][S], flux int64[R][S][2] (fixed point), R recorded steps, S sensors.
//   model image   cells / sensors / tables ... (device_types.h), read-only, L1/L2 resident.
//
// One launch = one pass of the whole pool over a WINDOW of measurement intervals (plan_launch: up to 1023 while nothing
// is recorded, as many as the tally staging holds otherwise; psim_gpu_set_option("steps_per_launch", 1) gives the
// one-interval pass that really moves the 64 algorithmic bytes per drift-step of SURVEY.md 8d).
#include "../../include/psim_b200.h"
#include "flatten.h"
#include "kernels.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_create_error;

// -------------------------------------------------------------------------------------------------------------
// Device memory of the library goes through a small caching allocator.  cudaMalloc / cudaFree of the phonon pool (7 GB for the
// 1e8-phonon job) and of the tally staging cost 25 - 100 ms each on these boxes - more than the kernels of most shipped
// models - and a model is typically run several times by one process (the runs of a multi-run model, the iterations of a
// re-iterated one, one model after another): a freed block is kept (per device, up to PSIM_DEVICE_CACHE_MB, default 16384;
// 0 disables) and handed to the next request it fits.  Blocks return to the driver when the cache is full, when an
// allocation fails, or on psim_gpu_release_cached().  Every caller frees after the work that used the block has completed
// (synchronous copies, cudaDeviceSynchronize in destroy), so a block is never handed on while it is in use.
namespace devmem {

struct Block {
    void* p;
    size_t bytes;
    int device;
};
std::mutex mu;
std::vector<Block> cached;
std::unordered_map<void*, Block> live;
size_t cached_bytes = 0;

size_t cache_cap() {
    static const size_t cap = [] {
        const char* e = std::getenv("PSIM_DEVICE_CACHE_MB");
        const double mb = e ? std::atof(e) : 16384.;
        return mb > 0. ? static_cast<size_t>(mb * 1048576.) : size_t{ 0 };
    }();
    return cap;
}

void release_locked(int device /* < 0: every device */) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (size_t i = 0; i < cached.size();) {
        if (device < 0 || cached[i].device == device) {
            cudaSetDevice(cached[i].device);
            cudaFree(cached[i].p);
            cached_bytes -= cached[i].bytes;
            cached[i] = cached.back();
            cached.pop_back();
        } else {
            ++i;
        }
    }
    cudaSetDevice(cur);
}

cudaError_t alloc(void** out, size_t bytes) {
    const size_t need = (std::max<size_t>(bytes, 1) + 511) & ~static_cast<size_t>(511);
    int device = 0;
    cudaGetDevice(&device);
    std::lock_guard<std::mutex> lock(mu);
    size_t best = cached.size();
    // The smallest cached block of this device that fits.  A small request never takes a block more than twice its size (it
    // would leave the next pool without one); a large one (pools, tally buffers: >= 1 MB) takes whatever is there - on these
    // boxes a single cudaMalloc stalls for 25 - 100 ms every so often, whatever its size, and HBM is not what a run is short of.
    const size_t limit = need >= (1u << 20) ? ~size_t{ 0 } : 2 * need + (1u << 20);
    for (size_t i = 0; i < cached.size(); ++i) {
        const Block& b = cached[i];
        if (b.device == device && b.bytes >= need && b.bytes <= limit && (best == cached.size() || b.bytes < cached[best].bytes)) { best = i; }
    }
    if (best < cached.size()) {
        const Block b = cached[best];
        cached[best] = cached.back();
        cached.pop_back();
        cached_bytes -= b.bytes;
        live[b.p] = b;
        *out = b.p;
        return cudaSuccess;
    }
    void* p = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMalloc(&p, need);
    if (std::getenv("PSIM_TIMING")) {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 2.) { std::fprintf(stderr, "psim timing [ms]: cudaMalloc of %zu bytes took %.1f\n", need, ms); }
    }
    if (e != cudaSuccess) {  // give the cache back to the driver and try once more
        cudaGetLastError();
        release_locked(device);
        e = cudaMalloc(&p, need);
        if (e != cudaSuccess) { return e; }
    }
    live[p] = Block{ p, need, device };
    *out = p;
    return cudaSuccess;
}

template<typename T> cudaError_t alloc(T** out, size_t bytes) {
    void* p = nullptr;
    const cudaError_t e = alloc(&p, bytes);
    *out = static_cast<T*>(p);
    return e;
}

void free(void* p) {
    if (!p) { return; }
    std::lock_guard<std::mutex> lock(mu);
    const auto it = live.find(p);
    if (it == live.end()) {  // not ours (never happens): straight to the driver
        cudaFree(p);
        return;
    }
    const Block b = it->second;
    live.erase(it);
    if (cached_bytes + b.bytes <= cache_cap()) {
        cached.push_back(b);
        cached_bytes += b.bytes;
    } else {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(b.device);
        cudaFree(b.p);
        cudaSetDevice(cur);
    }
}

}  // namespace devmem

}  // namespace

// -------------------------------------------------------------------------------------------------------------
struct psim_gpu {
    int device = 0;
    int sm_count = 0;
    psim::HostImage img;
    psim::HostImage img_tri;         // the same model with one flight cell per triangle: geometry arrays only (probes, option "merge_cells")
    psim::BirthPlan plan;
    DevParams P{};  // device pointers
    DevParams PL{}; // the same over the lattice image (device_types.h); valid while have_lattice
    bool have_lattice = false;      // the model has a lattice image and option "merge_cells" is 2
    bool pool_in_lattice = false;   // the live pool holds lattice coordinates (the launches so far recorded nothing)
    void* d_lat_cells = nullptr;
    void* d_lat_api_cells = nullptr;
    void* d_lat_shapes = nullptr;
    void* d_lat_subs = nullptr;
    void* d_lat_emitters = nullptr;
    void* d_lat_sub_fine = nullptr;
    void* d_lat_sub_sensor = nullptr;
    void* d_cells = nullptr;
    void* d_api_cells = nullptr;
    void* d_shapes = nullptr;
    void* d_classes = nullptr;
    void* d_step_sensors = nullptr;
    void* d_subs = nullptr;
    void* d_sensors = nullptr;
    void* d_materials = nullptr;
    void* d_emitters = nullptr;
    void* d_tables = nullptr;
    void* d_guides = nullptr;
    void* d_velocities = nullptr;
    void* d_sources = nullptr;
    DevBirth* d_births = nullptr;
    uint64_t* d_prefix = nullptr;
    float4* pool_a[2] = { nullptr, nullptr };
    uint4* pool_b[2] = { nullptr, nullptr };
    uint32_t* cnt[2] = { nullptr, nullptr };
    int cur = 0;  // copy holding the live pool
    uint32_t seg_cap = 0, n_warps = 0;
    int32_t* tally_e = nullptr;
    long long* tally_f = nullptr;
    long long* tally_acc = nullptr;  // [R][S][4] difference-form accumulator of runs that tally straight to global memory (allocated on first use)
    int32_t* carry_e = nullptr;      // [S]   running sums behind the last finalized row (difference-form tallies)
    long long* carry_f = nullptr;    // [S][2]
    bool diff_mode = false;          // this run tallies straight to global memory, rows kept as differences until finalized
    unsigned long long* d_stats = nullptr;
    unsigned long long* d_alive_hist = nullptr;
    unsigned long long* d_hist = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool have_sources = false;
    bool timing_open = false;
    uint32_t next_step = 0;
    uint32_t launches = 0;
    uint32_t birth_offset = 0;
    double kernel_ms = 0.;
    // options
    int64_t opt_steps_per_launch = 0;  // 0: automatic (plan_launch: long windows while nothing is recorded, then what the tally staging holds)
    int64_t opt_warps_per_sm = 0;
    int64_t opt_kernel = 2;          // 2: work-queue kernel, 0: lane-bound slots kernel, 1: lock-step kernel (first version, for A/B)
    int64_t opt_queue_slots = PSIM_QUEUE_SLOTS;  // phonons in flight per warp of the work-queue kernel (128, or 64: more L1 left for the mesh)
    int64_t opt_tally_shared = -1;
    int64_t opt_lattice_recorded = -1;  // recorded windows over the lattice image: -1 where a phonon crosses kLatticeRecordedCrossings
                                        // fine cells or more per measurement step, 0 never, 1 wherever the model has the image
    int64_t opt_merge_cells = 2;     // 0: one flight cell per model triangle (A/B, per-function probes), 1: pairs of triangles as
                                     // parallelograms, 2: also blocks of parallelograms as lattice cells where nothing is recorded
    uint32_t last_tally_shared = 0;
    uint32_t last_window = 0;
    bool ran_lattice_recorded = false;  // a recording launch of this run flew the lattice image
    uint32_t max_flux_fixed = 0;     // largest |velocity| in flux fixed-point units
    std::string err;
};

namespace {

int cuda_fail(psim_gpu* h, cudaError_t e, const char* what) {
    h->err = std::string(what) + ": " + cudaGetErrorString(e);
    return PSIM_E_CUDA;
}

#define PSIM_CUDA(call)                                               \
    do {                                                              \
        const cudaError_t e_ = (call);                                \
        if (e_ != cudaSuccess) { return cuda_fail(h, e_, #call); }    \
    } while (0)

// device memory that is released on every return path (probes, temporaries)
struct DeviceBuffer {
    void* p = nullptr;
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    ~DeviceBuffer() { devmem::free(p); }
    cudaError_t alloc(size_t bytes) { return devmem::alloc(&p, bytes ? bytes : 1); }
    template<typename T> T* as() const { return static_cast<T*>(p); }
    void* release() {
        void* q = p;
        p = nullptr;
        return q;
    }
};

template<typename T> int upload(psim_gpu* h, void** dst, const std::vector<T>& v) {
    DeviceBuffer b;  // a failed copy does not leave the allocation behind
    PSIM_CUDA(b.alloc(v.size() * sizeof(T)));
    if (!v.empty()) { PSIM_CUDA(cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice)); }
    *dst = b.release();
    return 0;
}

void free_plan(psim_gpu* h) {
    devmem::free(h->d_births);
    devmem::free(h->d_prefix);
    devmem::free(h->d_sources);
    h->d_births = nullptr;
    h->d_prefix = nullptr;
    h->d_sources = nullptr;
}

void free_pool(psim_gpu* h) {
    for (int i = 0; i < 2; ++i) {
        devmem::free(h->pool_a[i]);
        devmem::free(h->pool_b[i]);
        devmem::free(h->cnt[i]);
        h->pool_a[i] = nullptr;
        h->pool_b[i] = nullptr;
        h->cnt[i] = nullptr;
    }
    devmem::free(h->d_alive_hist);
    h->d_alive_hist = nullptr;
    h->seg_cap = 0;
    h->n_warps = 0;
}

constexpr size_t kSlotBytesPerBlock = static_cast<size_t>(SF_COUNT) * kSlots * 32 * 4 * kWarpsPerBlock;


constexpr int kQueueSlots = PSIM_QUEUE_SLOTS;                // slots per warp of the queues kernel ...
constexpr int kQueueSlotsSmall = 64;                         // ... and of its second instantiation (option "queue_slots")
constexpr size_t queue_bytes_per_block(int slots) {
    return static_cast<size_t>(SG_COUNT) * slots * 16 * kWarpsPerBlock + static_cast<size_t>(Q_COUNT) * ring_capacity(slots) * kWarpsPerBlock;
}
constexpr size_t kPostBytesPerBlock = static_cast<size_t>(kPostScratchBytes) * kWarpsPerBlock;  // tally_post_global's scratch
constexpr size_t kQueueBytesPerBlock = queue_bytes_per_block(kQueueSlots);
// what a block may use so that kSlotBlocks blocks (plus 1 KB each that the driver reserves) fit the planned carve-out;
// the tally staging gets what the slot storage leaves
constexpr size_t kSmemPerBlock = static_cast<size_t>(PSIM_SMEM_KB_PER_SM) * 1024 / kSlotBlocks - 1024;
static_assert(kSmemPerBlock > kQueueBytesPerBlock + 4096, "no room for the tally staging");
constexpr size_t kTallyStageBudget = kSmemPerBlock - kSlotBytesPerBlock;

size_t stage_budget(const psim_gpu* h) {
    return h->opt_kernel == 1 ? 100 * 1024 : (h->opt_kernel == 2 ? kSmemPerBlock - queue_bytes_per_block(static_cast<int>(h->opt_queue_slots)) : kTallyStageBudget);
}
constexpr int kStatWords = 8;                   // LaunchArgs::stats
constexpr uint32_t kLongWindow = 1023;           // steps per launch while nothing is recorded (10 bits of step in the slot word)
constexpr uint32_t kGlobalTallyWindow = 128;     // steps per launch when recorded tallies go straight to global memory
// Fine cells crossed per measurement step at the largest group velocity from which recorded windows fly the lattice image
// (run_steps).  Measured on linear_sides periodic with longer and longer steps (tools/gpu_lattice_threshold.py; fine / lattice
// kernel ms): 0.46 cells per step 6.9 / 9.3, 0.93: 8.7 / 8.6, 1.85: 10.4 / 8.1, 2.8: 12.1 / 8.2, 3.7: 13.6 / 8.5.
constexpr double kLatticeRecordedCrossings = 1.25;
constexpr uint32_t kManySensors = 256;           // from here on global atomics are spread thinly enough to need no staging

// Measurement intervals a launch may cover (its "window").
uint32_t effective_steps_per_launch(const psim_gpu* h) {
    return h->opt_steps_per_launch > 0 ? static_cast<uint32_t>(h->opt_steps_per_launch) : 64u;  // auto: as many as the staging holds
}

// How a launch that starts at step s0 tallies, and how far it may reach: windows without a recorded measurement need
// no staging at all; otherwise the per-block staging must fit the budget, unless the model has so many sensors that
// tallying straight into global memory is contention-free - with few sensors the window is shortened instead.
// Which staged form is exact for a launch over the birth entries [e0, e1): a block adds at most two contributions per
// phonon it handles to one (step, sensor) entry - as the first row of one flight segment, behind the last of another.
//   1  two 32-bit parts per flux component (kStageLoBits low bits + the signed rest): the fast one
//   4  three parts (2 x kStageWideBits + rest): pools too large for form 1 (above ~1.2e8 phonons per GPU)
//   2  64-bit sums (compare-and-swap loops): whatever is left, or on request
uint32_t staged_form(const psim_gpu* h, uint32_t s0, uint32_t s1) {
    if (h->opt_tally_shared == 2) { return 2u; }
    const uint64_t births = h->plan.prefix[h->plan.step_begin[s1]] - h->plan.prefix[h->plan.step_begin[s0]];
    const uint64_t chunks_per_warp = (((births + 31) >> 5) + h->n_warps - 1) / h->n_warps;
    const uint64_t per_entry = 2ull * kWarpsPerBlock * (h->seg_cap + chunks_per_warp * 32);
    const uint64_t flux_max = h->max_flux_fixed;
    const bool narrow = per_entry < (1ull << (32 - kStageLoBits)) && per_entry * ((flux_max >> kStageLoBits) + 1) < (1ull << 31);
    const bool wide = per_entry < (1ull << (32 - kStageWideBits)) && per_entry * ((flux_max >> (2 * kStageWideBits)) + 1) < (1ull << 31);
    if (narrow && h->opt_tally_shared != 4) { return 1u; }
    return wide ? 4u : 2u;
}

void plan_launch(const psim_gpu* h, uint32_t s0, uint32_t step_end, uint32_t& s1, uint32_t& form, size_t& smem, bool& records) {
    const uint32_t B = effective_steps_per_launch(h);
    form = 0;
    smem = 0;
    records = false;
    if (h->opt_steps_per_launch <= 0 && s0 + 2 <= h->P.first_tally_step) {
        // automatic mode, nothing recorded yet (steady state: the first 90 % of the steps): long windows, cut at the
        // first recorded measurement
        s1 = std::min(std::min(s0 + kLongWindow, step_end), h->P.first_tally_step - 1);
        if (s1 > s0) { return; }
    }
    s1 = std::min(s0 + B, step_end);
    if (s1 + 1 <= h->P.first_tally_step) { return; }  // nothing recorded in this window
    records = true;
    const size_t budget = stage_budget(h);
    if (h->diff_mode) {
        if (h->opt_steps_per_launch <= 0) { s1 = std::min(s0 + kGlobalTallyWindow, step_end); }  // no staging: nothing limits the window
        return;
    }
    const uint32_t want = staged_form(h, s0, s1);  // a shorter window has no more births: the form stays exact
    while (s1 > s0 + 1 && tally_stage_bytes(want, s1 - s0, h->P.n_sensors) > budget) { --s1; }
    smem = tally_stage_bytes(want, s1 - s0, h->P.n_sensors);
    if (smem <= budget) {
        form = want;
    } else {
        smem = 0;  // not even one step fits: plain global adds
    }
}

// The geometry half of the device image - flight cells, shapes, the model-cell map, partial-edge records, emitters - of
// either form of the model (merged: h->img, one flight cell per triangle: h->img_tri).
int upload_geometry(psim_gpu* h, const psim::HostImage& g) {
    for (void** p : { &h->d_cells, &h->d_shapes, &h->d_api_cells, &h->d_subs, &h->d_emitters }) {
        devmem::free(*p);
        *p = nullptr;
    }
    if (int rc = upload(h, &h->d_cells, g.cells)) { return rc; }
    if (int rc = upload(h, &h->d_shapes, g.shapes)) { return rc; }
    if (int rc = upload(h, &h->d_api_cells, g.api_cells)) { return rc; }
    if (int rc = upload(h, &h->d_subs, g.subs)) { return rc; }
    if (int rc = upload(h, &h->d_emitters, g.emitters)) { return rc; }
    h->P.cells = static_cast<const DevCell*>(h->d_cells);
    h->P.shapes = static_cast<const DevShape*>(h->d_shapes);
    h->P.api_cells = static_cast<const DevApiCell*>(h->d_api_cells);
    h->P.subs = static_cast<const DevSub*>(h->d_subs);
    h->P.emitters = static_cast<const DevEmitter*>(h->d_emitters);
    h->P.n_flight_cells = static_cast<uint32_t>(g.cells.size());
    h->P.n_shapes = static_cast<uint32_t>(g.shapes.size());
    h->P.fast_links = g.fast_links ? 1u : 0u;
    return 0;
}

// The lattice image, if the model has one: PL = P with the geometry pointers replaced.  Called after P is complete.
int upload_lattice(psim_gpu* h) {
    h->have_lattice = false;
    if (h->img.lattice_cells.empty()) { return 0; }
    for (void** p : { &h->d_lat_cells, &h->d_lat_shapes, &h->d_lat_api_cells, &h->d_lat_subs, &h->d_lat_emitters, &h->d_lat_sub_fine, &h->d_lat_sub_sensor }) {
        devmem::free(*p);
        *p = nullptr;
    }
    if (int rc = upload(h, &h->d_lat_cells, h->img.lattice_cells)) { return rc; }
    if (int rc = upload(h, &h->d_lat_shapes, h->img.lattice_shapes)) { return rc; }
    if (int rc = upload(h, &h->d_lat_api_cells, h->img.lattice_api_cells)) { return rc; }
    if (int rc = upload(h, &h->d_lat_subs, h->img.lattice_subs)) { return rc; }
    if (int rc = upload(h, &h->d_lat_emitters, h->img.lattice_emitters)) { return rc; }
    if (int rc = upload(h, &h->d_lat_sub_fine, h->img.lattice_sub_fine)) { return rc; }
    if (int rc = upload(h, &h->d_lat_sub_sensor, h->img.lattice_sub_sensor)) { return rc; }
    h->have_lattice = true;
    return 0;
}

// P over the lattice image (the scalars, sources and seed of P as they are now)
DevParams lattice_params(const psim_gpu* h) {
    DevParams L = h->P;
    L.cells = static_cast<const DevCell*>(h->d_lat_cells);
    L.shapes = static_cast<const DevShape*>(h->d_lat_shapes);
    L.api_cells = static_cast<const DevApiCell*>(h->d_lat_api_cells);
    L.subs = static_cast<const DevSub*>(h->d_lat_subs);
    L.emitters = static_cast<const DevEmitter*>(h->d_lat_emitters);
    L.sub_fine = static_cast<const uint32_t*>(h->d_lat_sub_fine);
    L.sub_sensor = static_cast<const uint32_t*>(h->d_lat_sub_sensor);
    L.n_flight_cells = static_cast<uint32_t>(h->img.lattice_cells.size());
    L.n_shapes = static_cast<uint32_t>(h->img.lattice_shapes.size());
    L.lattice = 1u;
    L.fast_links = h->img.lattice_fast_links ? 1u : 0u;
    return L;
}

int zero_run_state(psim_gpu* h) {
    const DevParams& P = h->P;
    const size_t n = static_cast<size_t>(P.recorded_steps) * P.n_sensors;
    PSIM_CUDA(cudaMemsetAsync(h->tally_e, 0, n * sizeof(int32_t), h->stream));
    PSIM_CUDA(cudaMemsetAsync(h->tally_f, 0, n * 2 * sizeof(long long), h->stream));
    PSIM_CUDA(cudaMemsetAsync(h->carry_e, 0, P.n_sensors * sizeof(int32_t), h->stream));
    PSIM_CUDA(cudaMemsetAsync(h->carry_f, 0, P.n_sensors * 2 * sizeof(long long), h->stream));
    // Tallies go straight to global memory when the caller asks for it, or when the model has so many sensors that the
    // per-CTA staging of a useful window would not fit and global atomics are spread thinly enough
    h->diff_mode = h->opt_tally_shared == 0 || (h->opt_tally_shared < 0 && P.n_sensors >= kManySensors);
    if (h->diff_mode) {
        if (!h->tally_acc) { PSIM_CUDA(devmem::alloc(&h->tally_acc, n * 4 * sizeof(long long))); }
        PSIM_CUDA(cudaMemsetAsync(h->tally_acc, 0, n * 4 * sizeof(long long), h->stream));
    }
    PSIM_CUDA(cudaMemsetAsync(h->d_stats, 0, kStatWords * sizeof(unsigned long long), h->stream));
    if (h->d_alive_hist) {
        PSIM_CUDA(cudaMemsetAsync(h->d_alive_hist, 0, static_cast<size_t>(P.num_steps + 1) * sizeof(unsigned long long), h->stream));
    }
    if (h->cnt[0]) {
        PSIM_CUDA(cudaMemsetAsync(h->cnt[0], 0, h->n_warps * sizeof(uint32_t), h->stream));
        PSIM_CUDA(cudaMemsetAsync(h->cnt[1], 0, h->n_warps * sizeof(uint32_t), h->stream));
    }
    h->cur = 0;
    h->pool_in_lattice = false;
    h->ran_lattice_recorded = false;
    h->next_step = 0;
    h->launches = 0;
    h->birth_offset = 0;
    h->kernel_ms = 0.;
    h->timing_open = false;
    PSIM_CUDA(cudaStreamSynchronize(h->stream));  // callers may launch on a different stream next
    return 0;
}

}  // namespace

extern "C" {

int psim_gpu_create(const psim_model_desc* desc, int device, psim_gpu** out) {
    if (!desc || !out) {
        g_create_error = "null argument";
        return PSIM_E_INVALID;
    }
    *out = nullptr;
    int n_dev = 0;
    const cudaError_t e0 = cudaGetDeviceCount(&n_dev);
    if (e0 != cudaSuccess || n_dev == 0) {
        g_create_error = std::string("no usable CUDA device (this library has no CPU path): ") +
                         (e0 != cudaSuccess ? cudaGetErrorString(e0) : "device count is 0");
        return PSIM_E_NO_DEVICE;
    }
    psim_gpu* h = new psim_gpu();
    // PSIM_TIMING: which part of a create took more than 2 ms (a call into the driver stalls for 15 - 80 ms every so often)
    auto lap_t = std::chrono::steady_clock::now();
    const bool timing = std::getenv("PSIM_TIMING") != nullptr;
    auto lap = [&](const char* what) {
        const auto now = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(now - lap_t).count();
        if (timing && ms > 2.) { std::fprintf(stderr, "psim timing [ms]: create: %s took %.1f\n", what, ms); }
        lap_t = now;
    };
    auto bail = [&](int rc) {
        g_create_error = h->err;
        psim_gpu_destroy(h);
        return rc;
    };
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) { device = 0; }
    }
    if (device >= n_dev) {
        h->err = "device index out of range";
        return bail(PSIM_E_NO_DEVICE);
    }
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        h->err = "cudaSetDevice failed";
        return bail(PSIM_E_NO_DEVICE);
    }
    if (cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        h->err = "cudaDeviceGetAttribute failed";
        return bail(PSIM_E_NO_DEVICE);
    }
    lap("device query");
    try {  // nothing throws across the ABI: a description whose sizes exhaust the host's memory is an invalid description
        if (int rc = psim::flatten_model(*desc, h->img, h->err)) { return bail(rc); }
        if (int rc = psim::flatten_model(*desc, h->img_tri, h->err, 0)) { return bail(rc); }
        for (auto* v : { &h->img_tri.tables }) { std::vector<float2>().swap(*v); }  // only its geometry arrays are kept
        std::vector<uint32_t>().swap(h->img_tri.guides);
        std::vector<float>().swap(h->img_tri.velocities);
        std::vector<DevSensor>().swap(h->img_tri.step_sensors);
    } catch (const std::exception& e) {
        h->err = std::string("model description rejected: ") + e.what();
        return bail(PSIM_E_INVALID);
    }
    lap("device images (host)");
    {
        float vmax = h->img.scalars.phasor ? 1000.f : 0.f;
        for (float v : h->img.velocities) { vmax = std::max(vmax, std::fabs(v)); }
        h->max_flux_fixed = static_cast<uint32_t>(std::min(4.0e9, std::ceil(static_cast<double>(vmax) * (1 << PSIM_FLUX_FRAC_BITS))));
    }
    // The work-queue kernel for every mesh.  (Until the transition was fused into the flight pass and re-queued at the front,
    // the lane-bound slots kernel was faster on meshes whose cell records do not fit L1; measured again on the final
    // kernels, kinked wire with 6174 cells: 246 ms queues, 253 ms slots.)  psim_gpu_set_option("kernel") overrides.
    h->opt_kernel = 2;
    // Slots per warp: 128 keep the queues of the rare kinds of work full (Si/Ge bench model: 63.6 ms for the long window
    // against 76.8 ms with 64) and leave 92 KB of L1; 64 slots leave 156 KB.  With 32-byte flight-cell records every shipped
    // mesh runs fastest with 128 (round 2, same box, 64 / 128 slots: kinked wire with 3150 flight cells = 100 KB of records
    // 165.1 / 161.6 ms, linear_sides 15.7 ms with 128, Si/Ge 69.7); 64 is kept for meshes whose records exceed the L1 that 128
    // slots leave by a wide margin.  psim_gpu_set_option("queue_slots") overrides.
    h->opt_queue_slots = (h->img.cells.size() * sizeof(DevCell) > 160u * 1024u) ? kQueueSlotsSmall : kQueueSlots;
    auto setup = [&]() -> int {
        if (int rc = upload(h, &h->d_classes, h->img.classes)) { return rc; }
        if (!h->img.step_sensors.empty()) {
            if (int rc = upload(h, &h->d_step_sensors, h->img.step_sensors)) { return rc; }
        }
        if (int rc = upload(h, &h->d_sensors, h->img.sensors)) { return rc; }
        if (int rc = upload(h, &h->d_materials, h->img.materials)) { return rc; }
        if (int rc = upload(h, &h->d_tables, h->img.tables)) { return rc; }
        if (int rc = upload(h, &h->d_velocities, h->img.velocities)) { return rc; }
        if (int rc = upload(h, &h->d_guides, h->img.guides)) { return rc; }
        h->P = h->img.scalars;
        if (int rc = upload_geometry(h, h->img)) { return rc; }
        if (int rc = upload_lattice(h)) { return rc; }
        lap("uploads");
        h->P.classes = static_cast<const DevSensor*>(h->d_classes);
        h->P.step_sensors = static_cast<const DevSensor*>(h->d_step_sensors);
        h->P.sensors = static_cast<const DevSensor*>(h->d_sensors);
        h->P.materials = static_cast<const DevMaterial*>(h->d_materials);
        h->P.tables = static_cast<const float2*>(h->d_tables);
        h->P.velocities = static_cast<const float*>(h->d_velocities);
        h->P.guides = static_cast<const uint32_t*>(h->d_guides);
        const size_t n = static_cast<size_t>(h->P.recorded_steps) * h->P.n_sensors;
        PSIM_CUDA(devmem::alloc(&h->tally_e, n * sizeof(int32_t)));
        PSIM_CUDA(devmem::alloc(&h->tally_f, n * 2 * sizeof(long long)));
        PSIM_CUDA(devmem::alloc(&h->carry_e, std::max<size_t>(h->P.n_sensors, 1) * sizeof(int32_t)));
        PSIM_CUDA(devmem::alloc(&h->carry_f, std::max<size_t>(h->P.n_sensors, 1) * 2 * sizeof(long long)));
        PSIM_CUDA(devmem::alloc(&h->d_stats, kStatWords * sizeof(unsigned long long)));
        PSIM_CUDA(devmem::alloc(&h->d_hist, static_cast<size_t>(h->P.n_cells) * sizeof(unsigned long long)));
        lap("tally / counter buffers");
        PSIM_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        lap("cudaStreamCreate");
        PSIM_CUDA(cudaEventCreate(&h->ev_begin));
        PSIM_CUDA(cudaEventCreate(&h->ev_end));
        lap("cudaEventCreate");
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_lockstep, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_slots<kSlots>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerBlock)));
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_queues<kQueueSlots, TALLY_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerBlock)));
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_queues<kQueueSlots, TALLY_STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerBlock)));
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_queues<kQueueSlots, TALLY_GLOBAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerBlock)));
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_queues<kQueueSlotsSmall, TALLY_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerBlock)));
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_queues<kQueueSlotsSmall, TALLY_STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerBlock)));
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_queues<kQueueSlotsSmall, TALLY_GLOBAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerBlock)));
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_queues<kQueueSlots, TALLY_LATTICE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerBlock)));
        PSIM_CUDA(cudaFuncSetAttribute(drift_kernel_queues<kQueueSlotsSmall, TALLY_LATTICE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemPerBlock)));
        lap("cudaFuncSetAttribute");
        const int rc = zero_run_state(h);
        lap("zeroing the run state");
        return rc;
    };
    if (int rc = setup()) { return bail(rc); }
    *out = h;
    return PSIM_OK;
}

int psim_gpu_set_sources(psim_gpu* h, const psim_source* sources, size_t n, uint64_t seed, uint32_t shard,
                         uint32_t num_shards) {
    if (!h) { return PSIM_E_INVALID; }
    PSIM_CUDA(cudaSetDevice(h->device));
    PSIM_CUDA(cudaStreamSynchronize(h->stream));
    h->have_sources = false;  // a failed call leaves the handle without sources, never with a plan that is half replaced
    try {
        if (int rc = psim::plan_births(h->img, sources, n, shard, num_shards, h->plan, h->err)) { return rc; }
    } catch (const std::exception& e) {
        h->err = std::string("sources rejected: ") + e.what();
        return PSIM_E_INVALID;
    }
    free_plan(h);
    if (int rc = upload(h, &h->d_sources, h->plan.sources)) { return rc; }
    {
        void* p = nullptr;
        if (int rc = upload(h, &p, h->plan.births)) { return rc; }
        h->d_births = static_cast<DevBirth*>(p);
        p = nullptr;
        if (int rc = upload(h, &p, h->plan.prefix)) { return rc; }
        h->d_prefix = static_cast<uint64_t*>(p);
    }
    h->P.sources = static_cast<const DevSource*>(h->d_sources);
    h->P.n_sources = static_cast<uint32_t>(n);
    h->P.seed_lo = static_cast<uint32_t>(seed);
    h->P.seed_hi = static_cast<uint32_t>(seed >> 32);

    // pool geometry: one segment per resident warp
    int blocks_per_sm = 0;
    if (h->opt_kernel == 1) {
        PSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, drift_kernel_lockstep, kBlock, 0));
    } else if (h->opt_kernel == 2) {
        if (h->opt_queue_slots == kQueueSlots) {
            PSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, drift_kernel_queues<kQueueSlots, TALLY_STAGED>, kBlock, kSmemPerBlock));
        } else {
            PSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, drift_kernel_queues<kQueueSlotsSmall, TALLY_STAGED>, kBlock, kSmemPerBlock));
        }
    } else {  // shared-memory slots + the largest tally staging a launch may ask for
        const size_t dyn = kSlotBytesPerBlock + kTallyStageBudget;
        PSIM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, drift_kernel_slots<kSlots>, kBlock, dyn));
    }
    blocks_per_sm = std::max(1, blocks_per_sm);
    int warps_per_sm = blocks_per_sm * kWarpsPerBlock;
    if (h->opt_warps_per_sm > 0) {
        warps_per_sm = static_cast<int>(std::max<int64_t>(kWarpsPerBlock, h->opt_warps_per_sm / kWarpsPerBlock * kWarpsPerBlock));
    }
    const uint32_t n_warps = static_cast<uint32_t>(h->sm_count * warps_per_sm);
    const uint64_t per_warp = (h->plan.shard_phonons + n_warps - 1) / n_warps;
    const uint64_t cap = per_warp + per_warp / 8 + 1024;
    if (cap > 0xFFFFFFF0ull) {
        h->err = "shard too large for one device";
        return PSIM_E_INVALID;
    }
    const uint32_t seg_cap = static_cast<uint32_t>((cap + 31) & ~31ull);
    if (n_warps != h->n_warps || seg_cap > h->seg_cap) {
        // (re)allocate the pool; a later run of the same handle that fits (the runs of a multi-run model) keeps it
        free_pool(h);
        const size_t slots = static_cast<size_t>(seg_cap) * n_warps;
        for (int i = 0; i < 2; ++i) {
            PSIM_CUDA(devmem::alloc(&h->pool_a[i], slots * sizeof(float4)));
            PSIM_CUDA(devmem::alloc(&h->pool_b[i], slots * sizeof(uint4)));
            PSIM_CUDA(devmem::alloc(&h->cnt[i], n_warps * sizeof(uint32_t)));
        }
        PSIM_CUDA(devmem::alloc(&h->d_alive_hist, static_cast<size_t>(h->P.num_steps + 1) * sizeof(unsigned long long)));
        h->seg_cap = seg_cap;
        h->n_warps = n_warps;
    }
    h->have_sources = true;
    return zero_run_state(h);
}

int psim_gpu_run_steps(psim_gpu* h, uint32_t step_begin, uint32_t step_end, void* cuda_stream) {
    if (!h) { return PSIM_E_INVALID; }
    if (!h->have_sources) {
        h->err = "psim_gpu_set_sources must be called before running";
        return PSIM_E_STATE;
    }
    const uint32_t last = h->P.num_steps - 1;
    step_end = std::min(step_end, last);
    if (step_begin != h->next_step || step_end < step_begin) {
        h->err = "measurement steps must be submitted in order, starting at 0 after set_sources/reset";
        return PSIM_E_STATE;
    }
    PSIM_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->stream;
    for (uint32_t s0 = step_begin, s1 = 0; s0 < step_end; s0 = s1) {
        uint32_t form = 0;
        size_t smem = 0;
        bool records = false;
        plan_launch(h, s0, step_end, s1, form, smem, records);
        const bool shared = form != 0u;
        LaunchArgs a{};
        // A window that records nothing flies the lattice image (device_types.h) - from the start of a run until the first
        // window that records; that one converts the pool back as it fetches it.
        // A model whose phonons cross several fine cells per measurement step flies the lattice image in its recorded windows
        // too: the sensor area of every crossed measurement is found from the position at that instant (kernels.cuh:
        // tally_lattice).  Only where the tallies go to global memory - the many-sensor models; few sensors mean large areas.
        const bool lattice_recorded = h->diff_mode && (h->opt_lattice_recorded > 0 ||
                                                       (h->opt_lattice_recorded < 0 && h->img.lattice_cells_per_step >= kLatticeRecordedCrossings));
        const bool lattice = h->have_lattice && h->opt_merge_cells == 2 && (!records || lattice_recorded) && (h->pool_in_lattice || s0 == 0);
        a.P = lattice ? lattice_params(h) : h->P;
        a.convert_input = (h->pool_in_lattice && !lattice) ? 1u : 0u;
        a.lattice_cells = static_cast<const DevCell*>(h->d_lat_cells);
        a.lattice_sub_fine = static_cast<const uint32_t*>(h->d_lat_sub_fine);
        h->pool_in_lattice = lattice;
        h->ran_lattice_recorded |= lattice && records;
        a.in_a = h->pool_a[h->cur];
        a.in_b = h->pool_b[h->cur];
        a.out_a = h->pool_a[h->cur ^ 1];
        a.out_b = h->pool_b[h->cur ^ 1];
        a.cnt_in = h->cnt[h->cur];
        a.cnt_out = h->cnt[h->cur ^ 1];
        a.seg_cap = h->seg_cap;
        a.n_warps = h->n_warps;
        a.step_begin = s0;
        a.step_end = s1;
        const uint32_t e0 = h->plan.step_begin[s0], e1 = h->plan.step_begin[s1];
        a.births = h->d_births + e0;
        a.birth_prefix = h->d_prefix + e0;
        a.n_birth_entries = e1 - e0;
        a.birth_base = h->plan.prefix[e0];
        a.n_births = h->plan.prefix[e1] - h->plan.prefix[e0];
        a.birth_warp_offset = h->birth_offset;
        h->birth_offset = static_cast<uint32_t>((h->birth_offset + ((a.n_births + 31) >> 5)) % h->n_warps);
        a.tally_e = h->tally_e;
        a.tally_f = h->tally_f;
        a.tally_acc = h->tally_acc;
        a.tally_shared = h->diff_mode ? 3u : form;  // form 0: a staged window that did not fit even for one step (plain global adds)
        a.stats = h->d_stats;
        a.alive_hist = h->d_alive_hist;
        a.launch_index = h->launches;
        h->last_tally_shared = a.tally_shared;
        h->last_window = s1 - s0;
        if (!h->timing_open) {
            PSIM_CUDA(cudaEventRecord(h->ev_begin, st));
            h->timing_open = true;
        }
        const dim3 grid(h->n_warps / kWarpsPerBlock);
        const size_t dyn = shared ? smem : 0;
        if (h->opt_kernel == 1) {
            drift_kernel_lockstep<<<grid, kBlock, dyn, st>>>(a);
        } else if (h->opt_kernel == 2) {
            // one instantiation per (slots per warp, what the window does with its measurement events)
            const int tally = !records ? TALLY_NONE : (h->diff_mode ? (lattice ? TALLY_LATTICE : TALLY_GLOBAL) : TALLY_STAGED);
            const size_t bytes = queue_bytes_per_block(static_cast<int>(h->opt_queue_slots)) + ((tally == TALLY_GLOBAL || tally == TALLY_LATTICE) ? kPostBytesPerBlock : 0) +
                                 (shared ? ((smem + 127) & ~static_cast<size_t>(127)) : 0);
            const bool big = h->opt_queue_slots == kQueueSlots;
            if (tally == TALLY_NONE) {
                if (big) { drift_kernel_queues<kQueueSlots, TALLY_NONE><<<grid, kBlock, bytes, st>>>(a); }
                else { drift_kernel_queues<kQueueSlotsSmall, TALLY_NONE><<<grid, kBlock, bytes, st>>>(a); }
            } else if (tally == TALLY_GLOBAL) {
                if (big) { drift_kernel_queues<kQueueSlots, TALLY_GLOBAL><<<grid, kBlock, bytes, st>>>(a); }
                else { drift_kernel_queues<kQueueSlotsSmall, TALLY_GLOBAL><<<grid, kBlock, bytes, st>>>(a); }
            } else if (tally == TALLY_LATTICE) {
                if (big) { drift_kernel_queues<kQueueSlots, TALLY_LATTICE><<<grid, kBlock, bytes, st>>>(a); }
                else { drift_kernel_queues<kQueueSlotsSmall, TALLY_LATTICE><<<grid, kBlock, bytes, st>>>(a); }
            } else {
                if (big) { drift_kernel_queues<kQueueSlots, TALLY_STAGED><<<grid, kBlock, bytes, st>>>(a); }
                else { drift_kernel_queues<kQueueSlotsSmall, TALLY_STAGED><<<grid, kBlock, bytes, st>>>(a); }
            }
        } else {
            const size_t slots = kSlotBytesPerBlock + (shared ? ((smem + 127) & ~static_cast<size_t>(127)) : 0);
            drift_kernel_slots<kSlots><<<grid, kBlock, slots, st>>>(a);
        }
        PSIM_CUDA(cudaGetLastError());
        h->cur ^= 1;
        ++h->launches;
    }
    if (h->diff_mode && step_end + 1 > h->P.first_tally_step) {
        // the rows whose measurement steps are now complete turn from differences into sums
        const uint32_t F = h->P.first_tally_step;
        const uint32_t row_begin = (step_begin + 1 > F) ? step_begin + 1 - F : 0u, row_end = std::min(step_end + 1 - F, h->P.recorded_steps);
        if (row_end > row_begin) {
            const uint32_t threads = 3u * h->P.n_sensors;
            finalize_rows_kernel<<<(threads + 127) / 128, 128, 0, st>>>(h->tally_acc, h->tally_e, h->tally_f, h->carry_e, h->carry_f, h->P.n_sensors,
                                                                         row_begin, row_end);
            PSIM_CUDA(cudaGetLastError());
        }
    }
    PSIM_CUDA(cudaEventRecord(h->ev_end, st));
    h->next_step = step_end;
    return PSIM_OK;
}

int psim_gpu_synchronize(psim_gpu* h) {
    if (!h) { return PSIM_E_INVALID; }
    PSIM_CUDA(cudaSetDevice(h->device));
    PSIM_CUDA(cudaDeviceSynchronize());
    if (h->timing_open) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_begin, h->ev_end) == cudaSuccess) { h->kernel_ms = ms; }
    }
    unsigned long long st[kStatWords] = {};
    PSIM_CUDA(cudaMemcpy(st, h->d_stats, sizeof(st), cudaMemcpyDeviceToHost));
    if (st[3]) {
        h->err = "phonon pool capacity exceeded";
        return PSIM_E_OVERFLOW;
    }
    if (st[4]) {
        h->err = "a phonon needed more than 8191 random-number blocks inside one measurement interval (thousands of scatters or diffuse "
                 "wall hits per interval): shorten the measurement interval, or run with the lock-step kernel (option \"kernel\" = 1)";
        return PSIM_E_RNG;
    }
    return PSIM_OK;
}

int psim_gpu_next_window(psim_gpu* h, uint32_t step_begin, uint32_t* step_end) {
    if (!h || !step_end) { return PSIM_E_INVALID; }
    if (!h->have_sources) {
        h->err = "psim_gpu_set_sources must be called before planning windows";
        return PSIM_E_STATE;
    }
    const uint32_t last = h->P.num_steps - 1;
    *step_end = last;
    if (step_begin >= last) { return PSIM_OK; }
    uint32_t form = 0;
    size_t smem = 0;
    bool records = false;
    plan_launch(h, step_begin, last, *step_end, form, smem, records);
    return PSIM_OK;
}

int psim_gpu_run(psim_gpu* h) {
    if (!h) { return PSIM_E_INVALID; }
    if (int rc = psim_gpu_run_steps(h, h->next_step, h->P.num_steps - 1, nullptr)) { return rc; }
    return psim_gpu_synchronize(h);
}

int psim_gpu_get_tallies(psim_gpu* h, int32_t* energy, double* flux, int64_t* flux_fixed) {
    if (!h) { return PSIM_E_INVALID; }
    if (int rc = psim_gpu_synchronize(h)) { return rc; }
    const uint32_t R = h->P.recorded_steps, S = h->P.n_sensors;
    const size_t n = static_cast<size_t>(R) * S;
    if (n == 0 || (!energy && !flux && !flux_fixed)) { return PSIM_OK; }
    // device [R][S] -> caller [S][R] (the reference's per-sensor vectors): transposed and converted ON THE DEVICE, then one
    // copy per requested array straight into the caller's buffer (31 MB of tallies on the kinked wire: the host-side
    // transpose through two temporary vectors was the slowest part of a run's epilogue in round 1).  A pinned staging buffer
    // of that size was measured and dropped: on these boxes cudaMallocHost + cudaFreeHost of 25 - 50 MB cost 100 - 400 ms
    // per handle, more than the pageable copy they would have sped up.
    PSIM_CUDA(cudaSetDevice(h->device));
    const size_t off_f = (n * sizeof(int32_t) + 15) & ~static_cast<size_t>(15);
    const size_t bytes = off_f + n * (2 * sizeof(double) + 2 * sizeof(long long));
    DeviceBuffer dev;
    PSIM_CUDA(dev.alloc(bytes));
    int32_t* d_e = dev.as<int32_t>();
    double* d_f = reinterpret_cast<double*>(dev.as<unsigned char>() + off_f);
    long long* d_x = reinterpret_cast<long long*>(d_f + 2 * n);
    const dim3 block(32, 8), grid((S + 31) / 32, (R + 31) / 32);
    transpose_tallies_kernel<<<grid, block, 0, h->stream>>>(h->tally_e, h->tally_f, R, S, 1. / static_cast<double>(1 << PSIM_FLUX_FRAC_BITS),
                                                             energy ? d_e : nullptr, flux ? d_f : nullptr, flux_fixed ? d_x : nullptr);
    PSIM_CUDA(cudaGetLastError());
    PSIM_CUDA(cudaStreamSynchronize(h->stream));
    if (energy) { PSIM_CUDA(cudaMemcpy(energy, d_e, n * sizeof(int32_t), cudaMemcpyDeviceToHost)); }
    if (flux) { PSIM_CUDA(cudaMemcpy(flux, d_f, 2 * n * sizeof(double), cudaMemcpyDeviceToHost)); }
    if (flux_fixed) { PSIM_CUDA(cudaMemcpy(flux_fixed, d_x, 2 * n * sizeof(long long), cudaMemcpyDeviceToHost)); }
    return PSIM_OK;
}

int psim_gpu_tally_buffers(psim_gpu* h, void** energy_dev, void** flux_dev, uint32_t* recorded_steps, uint32_t* num_sensors) {
    if (!h) { return PSIM_E_INVALID; }
    if (energy_dev) { *energy_dev = h->tally_e; }
    if (flux_dev) { *flux_dev = h->tally_f; }
    if (recorded_steps) { *recorded_steps = h->P.recorded_steps; }
    if (num_sensors) { *num_sensors = h->P.n_sensors; }
    return PSIM_OK;
}

int psim_gpu_alive(psim_gpu* h, uint64_t* alive) {
    if (!h || !alive) { return PSIM_E_INVALID; }
    *alive = 0;
    if (!h->have_sources) { return PSIM_OK; }
    if (int rc = psim_gpu_synchronize(h)) { return rc; }
    std::vector<uint32_t> c(h->n_warps);
    PSIM_CUDA(cudaMemcpy(c.data(), h->cnt[h->cur], c.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (uint32_t v : c) { *alive += v; }
    return PSIM_OK;
}

int psim_gpu_cell_histogram(psim_gpu* h, uint64_t* per_cell) {
    if (!h || !per_cell) { return PSIM_E_INVALID; }
    if (!h->have_sources) {
        h->err = "no pool yet";
        return PSIM_E_STATE;
    }
    if (int rc = psim_gpu_synchronize(h)) { return rc; }
    PSIM_CUDA(cudaMemset(h->d_hist, 0, static_cast<size_t>(h->P.n_cells) * sizeof(unsigned long long)));
    cell_histogram_kernel<<<h->n_warps / kWarpsPerBlock, kBlock>>>(h->P, h->pool_a[h->cur], h->pool_b[h->cur], h->cnt[h->cur], h->seg_cap, h->n_warps, h->d_hist,
                                                                   h->pool_in_lattice ? static_cast<const DevCell*>(h->d_lat_cells) : nullptr,
                                                                   static_cast<const uint32_t*>(h->d_lat_sub_fine));
    PSIM_CUDA(cudaGetLastError());
    PSIM_CUDA(cudaMemcpy(per_cell, h->d_hist, static_cast<size_t>(h->P.n_cells) * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return PSIM_OK;
}

int psim_gpu_get_stats(psim_gpu* h, psim_stats* out) {
    if (!h || !out) { return PSIM_E_INVALID; }
    std::memset(out, 0, sizeof(*out));
    if (int rc = psim_gpu_synchronize(h)) { return rc; }
    unsigned long long st[kStatWords] = {};
    PSIM_CUDA(cudaMemcpy(st, h->d_stats, sizeof(st), cudaMemcpyDeviceToHost));
    out->total_phonons = h->plan.total_phonons;
    out->shard_phonons = h->plan.shard_phonons + h->plan.shard_unrecorded;
    out->drift_steps = st[0];
    out->events = st[1];
    if (h->d_alive_hist && h->launches) {
        std::vector<unsigned long long> hist(h->launches);
        PSIM_CUDA(cudaMemcpy(hist.data(), h->d_alive_hist, hist.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        out->peak_alive = *std::max_element(hist.begin(), hist.end());
    }
    out->kernel_ms = h->kernel_ms;
    out->launches = h->launches;
    out->steps_per_launch = h->last_window ? h->last_window : effective_steps_per_launch(h);  // steps of the last launch
    out->warps = h->n_warps;
    out->tally_in_shared = h->last_tally_shared;
    out->kernel = static_cast<uint32_t>(h->opt_kernel);
    out->flight_cells = h->P.n_flight_cells;
    out->lattice_cells = (h->have_lattice && h->opt_merge_cells == 2) ? static_cast<uint32_t>(h->img.lattice_cells.size()) : 0u;
    out->lattice_recorded = h->ran_lattice_recorded ? 1u : 0u;
    out->image_bytes = h->img.cells.size() * sizeof(DevCell) + h->img.api_cells.size() * sizeof(DevApiCell) + h->img.shapes.size() * sizeof(DevShape) + (h->img.classes.size() + h->img.step_sensors.size()) * sizeof(DevSensor) + h->img.subs.size() * sizeof(DevSub) +
                       h->img.sensors.size() * sizeof(DevSensor) + h->img.materials.size() * sizeof(DevMaterial) +
                       h->img.emitters.size() * sizeof(DevEmitter) + h->img.tables.size() * sizeof(float2) +
                       h->img.velocities.size() * sizeof(float);
    out->plan_bytes = h->plan.sources.size() * sizeof(DevSource) + h->plan.births.size() * sizeof(DevBirth) +
                      h->plan.prefix.size() * sizeof(uint64_t);
    out->tally_bytes = static_cast<uint64_t>(h->P.recorded_steps) * h->P.n_sensors * 20ull;
    return PSIM_OK;
}

int psim_gpu_set_option(psim_gpu* h, const char* name, int64_t value) {
    if (!h || !name) { return PSIM_E_INVALID; }
    const std::string k(name);
    if (k == "steps_per_launch") {
        if (value < 0 || value > 1023) {
            h->err = "steps_per_launch must be in [0, 1023] (0 = automatic)";
            return PSIM_E_INVALID;
        }
        h->opt_steps_per_launch = value;
    } else if (k == "warps_per_sm") {
        if (h->have_sources) {
            h->err = "warps_per_sm must be set before set_sources";
            return PSIM_E_STATE;
        }
        h->opt_warps_per_sm = value;
    } else if (k == "kernel") {
        if (h->have_sources || value < 0 || value > 2) {
            h->err = "kernel must be 0 (shared-memory slots), 1 (lock step) or 2 (work queues) and set before set_sources";
            return PSIM_E_STATE;
        }
        h->opt_kernel = value;
    } else if (k == "queue_slots") {
        if (h->have_sources || (value != kQueueSlots && value != kQueueSlotsSmall)) {
            h->err = "queue_slots must be " + std::to_string(kQueueSlots) + " or " + std::to_string(kQueueSlotsSmall) + " and set before set_sources";
            return PSIM_E_STATE;
        }
        h->opt_queue_slots = value;
    } else if (k == "merge_cells") {
        if (h->have_sources || value < 0 || value > 2) {
            h->err = "merge_cells must be 0 (one flight cell per model triangle), 1 (pairs of triangles fly as one parallelogram) or 2 (and blocks of "
                     "parallelograms as one lattice cell where nothing is recorded) and set before set_sources";
            return PSIM_E_STATE;
        }
        if ((value != 0) != (h->opt_merge_cells != 0)) {
            PSIM_CUDA(cudaSetDevice(h->device));
            PSIM_CUDA(cudaDeviceSynchronize());
            if (int rc = upload_geometry(h, value ? h->img : h->img_tri)) { return rc; }
        }
        h->opt_merge_cells = value;
    } else if (k == "lattice_recorded") {
        if (h->have_sources || value < -1 || value > 1) {
            h->err = "lattice_recorded must be -1 (automatic), 0 or 1 and set before set_sources";
            return PSIM_E_STATE;
        }
        h->opt_lattice_recorded = value;
    } else if (k == "tally_shared") {
        if (h->have_sources || value < -1 || value > 4 || value == 3) {  // the tally form of a run is fixed when it starts
            h->err = "tally_shared must be -1 (automatic), 0 (global memory), 1 / 4 / 2 (staged: two / three 32-bit parts, 64-bit) and set before set_sources";
            return PSIM_E_STATE;
        }
        h->opt_tally_shared = value;
    } else {
        h->err = "unknown option: " + k;
        return PSIM_E_INVALID;
    }
    return PSIM_OK;
}

int psim_gpu_reset(psim_gpu* h) {
    if (!h) { return PSIM_E_INVALID; }
    PSIM_CUDA(cudaSetDevice(h->device));
    PSIM_CUDA(cudaDeviceSynchronize());
    if (int rc = zero_run_state(h)) { return rc; }
    PSIM_CUDA(cudaStreamSynchronize(h->stream));
    return PSIM_OK;
}

void psim_gpu_destroy(psim_gpu* h) {
    if (!h) { return; }
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    free_pool(h);
    free_plan(h);
    for (void* p : { h->d_lat_cells, h->d_lat_api_cells, h->d_lat_shapes, h->d_lat_subs, h->d_lat_emitters, h->d_lat_sub_fine, h->d_lat_sub_sensor }) { devmem::free(p); }
    devmem::free(h->d_cells);
    devmem::free(h->d_api_cells);
    devmem::free(h->d_shapes);
    devmem::free(h->d_classes);
    devmem::free(h->d_step_sensors);
    devmem::free(h->d_subs);
    devmem::free(h->d_sensors);
    devmem::free(h->d_materials);
    devmem::free(h->d_emitters);
    devmem::free(h->d_tables);
    devmem::free(h->d_velocities);
    devmem::free(h->d_guides);
    devmem::free(h->tally_e);
    devmem::free(h->tally_f);
    devmem::free(h->tally_acc);
    devmem::free(h->carry_e);
    devmem::free(h->carry_f);
    devmem::free(h->d_stats);
    devmem::free(h->d_hist);
    if (h->ev_begin) { cudaEventDestroy(h->ev_begin); }
    if (h->ev_end) { cudaEventDestroy(h->ev_end); }
    if (h->stream) { cudaStreamDestroy(h->stream); }
    delete h;
}

void psim_gpu_release_cached(void) {
    std::lock_guard<std::mutex> lock(devmem::mu);
    devmem::release_locked(-1);
}

const char* psim_gpu_last_error(const psim_gpu* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int psim_gpu_probe_sample(psim_gpu* h, uint32_t table, const float* u1, const float* u2, size_t n, uint32_t* out_bin,
                          uint32_t* out_ta, uint32_t* out_bin_bisect) {
    if (!h || !u1 || !u2 || !out_bin || !out_ta || !out_bin_bisect || table >= h->P.n_tables) { return PSIM_E_INVALID; }
    if (n == 0) { return PSIM_OK; }
    PSIM_CUDA(cudaSetDevice(h->device));
    DeviceBuffer d1, d2, db, dt, dp;
    PSIM_CUDA(d1.alloc(n * 4));
    PSIM_CUDA(d2.alloc(n * 4));
    PSIM_CUDA(dp.alloc(n * 4));
    PSIM_CUDA(db.alloc(n * 4));
    PSIM_CUDA(dt.alloc(n * 4));
    PSIM_CUDA(cudaMemcpy(d1.p, u1, n * 4, cudaMemcpyHostToDevice));
    PSIM_CUDA(cudaMemcpy(d2.p, u2, n * 4, cudaMemcpyHostToDevice));
    probe_sample_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(h->P, table, d1.as<float>(), d2.as<float>(), n, db.as<uint32_t>(),
                                                                          dt.as<uint32_t>(), dp.as<uint32_t>());
    PSIM_CUDA(cudaGetLastError());
    PSIM_CUDA(cudaMemcpy(out_bin_bisect, dp.p, n * 4, cudaMemcpyDeviceToHost));
    PSIM_CUDA(cudaMemcpy(out_bin, db.p, n * 4, cudaMemcpyDeviceToHost));
    PSIM_CUDA(cudaMemcpy(out_ta, dt.p, n * 4, cudaMemcpyDeviceToHost));
    return PSIM_OK;
}

int psim_gpu_probe_flight(psim_gpu* h, const uint32_t* cell, const float* in, size_t n, float* out) {
    if (!h || !cell || !in || !out) { return PSIM_E_INVALID; }
    if (n == 0) { return PSIM_OK; }
    for (size_t i = 0; i < n; ++i) {
        if (cell[i] >= h->P.n_cells) {  // model cells
            h->err = "probe_flight: cell index out of range";
            return PSIM_E_INVALID;
        }
    }
    PSIM_CUDA(cudaSetDevice(h->device));
    // the probe speaks model triangles: it runs on the one-flight-cell-per-triangle form of the geometry
    DeviceBuffer dc, di, dout, cells, shapes;
    PSIM_CUDA(cells.alloc(h->img_tri.cells.size() * sizeof(DevCell)));
    PSIM_CUDA(shapes.alloc(h->img_tri.shapes.size() * sizeof(DevShape)));
    PSIM_CUDA(cudaMemcpy(cells.p, h->img_tri.cells.data(), h->img_tri.cells.size() * sizeof(DevCell), cudaMemcpyHostToDevice));
    PSIM_CUDA(cudaMemcpy(shapes.p, h->img_tri.shapes.data(), h->img_tri.shapes.size() * sizeof(DevShape), cudaMemcpyHostToDevice));
    DevParams P = h->P;
    P.cells = cells.as<DevCell>();
    P.shapes = shapes.as<DevShape>();
    PSIM_CUDA(dc.alloc(n * 4));
    PSIM_CUDA(di.alloc(n * 16));
    PSIM_CUDA(dout.alloc(n * 24));
    PSIM_CUDA(cudaMemcpy(dc.p, cell, n * 4, cudaMemcpyHostToDevice));
    PSIM_CUDA(cudaMemcpy(di.p, in, n * 16, cudaMemcpyHostToDevice));
    probe_flight_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(P, dc.as<uint32_t>(), di.as<float>(), n, dout.as<float>());
    PSIM_CUDA(cudaGetLastError());
    PSIM_CUDA(cudaMemcpy(out, dout.p, n * 24, cudaMemcpyDeviceToHost));
    return PSIM_OK;
}

int psim_gpu_probe_rates(psim_gpu* h, uint32_t sensor, const double* omega, const uint32_t* ta, size_t n, double* rates) {
    if (!h || !omega || !ta || !rates || sensor >= h->P.n_sensors) { return PSIM_E_INVALID; }
    if (n == 0) { return PSIM_OK; }
    PSIM_CUDA(cudaSetDevice(h->device));
    std::vector<float> w(n), r(3 * n);
    for (size_t i = 0; i < n; ++i) { w[i] = static_cast<float>(omega[i] * PSIM_FREQ_SCALE); }
    DeviceBuffer dw, dr, dta;
    PSIM_CUDA(dw.alloc(n * 4));
    PSIM_CUDA(dr.alloc(3 * n * 4));
    PSIM_CUDA(dta.alloc(n * 4));
    PSIM_CUDA(cudaMemcpy(dw.p, w.data(), n * 4, cudaMemcpyHostToDevice));
    PSIM_CUDA(cudaMemcpy(dta.p, ta, n * 4, cudaMemcpyHostToDevice));
    probe_rates_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(h->P, sensor, dw.as<float>(), dta.as<uint32_t>(), n, dr.as<float>());
    PSIM_CUDA(cudaGetLastError());
    PSIM_CUDA(cudaMemcpy(r.data(), dr.p, 3 * n * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < 3 * n; ++i) { rates[i] = static_cast<double>(r[i]) * 1e9; }  // 1/ns -> 1/s
    return PSIM_OK;
}

}  // extern "C"
