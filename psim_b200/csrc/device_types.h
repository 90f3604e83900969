// Plain-old-data records shared by the host flattener (flatten.cpp) and the CUDA kernels (kernels.cu).
// Everything the per-phonon loop reads is in these structs; nothing here owns memory.
//
// Coordinates: a phonon's position is stored in the AFFINE frame of its current flight cell,
//   p = Q0 + b1 u + b2 v.
// A flight cell is either one triangular cell of the model file (u = P2 - P1, v = P3 - P1; b1 >= 0, b2 >= 0, b1 + b2 <= 1;
// edges: 0 b2 = 0 (P1->P2), 1 b1 + b2 = 1 (P2->P3), 2 b1 = 0 (P3->P1)) or a PARALLELOGRAM made of two of them
// (0 <= b1, b2 <= 1; edges: 0 b2 = 0, 1 b1 = 1, 2 b2 = 1, 3 b1 = 0), and "time to hit an edge" is one division per axis.
// The reference instead intersects the path segment with three slope/intercept lines in absolute fp64 coordinates
// (reference psim/src/modelSimulator.cpp:87-122, geometry.cpp:104-138); in fp32 that is not watertight, the frame form
// is (a phonon can never be outside its own cell), and it costs a fraction of the arithmetic.
//
// Why parallelograms: the reference's builder makes every sensor area a rectangle cut into two triangles
// (builder_tools.py:addRectangularCell).  Crossing that diagonal does nothing to a phonon (same sensor, same material:
// TransitionSurface::handlePhonon, surface.cpp:71-75, only re-labels its cell), yet it ends a flight segment: 17 - 40 % of
// all segments of the shipped models are such crossings.  flatten.cpp therefore merges two triangles
// into one flight cell where that is exactly neutral - same sensor, same specularity, the shared edge a plain whole-edge
// transition, the union a parallelogram - and the per-phonon loop never sees the diagonal.  Everything the API speaks
// (cell indices of sources, the per-cell histogram) stays in triangles: DevApiCell / DevCell::tri translate.
#ifndef PSIM_B200_DEVICE_TYPES_H
#define PSIM_B200_DEVICE_TYPES_H

#include <stdint.h>
#include <vector_types.h>  // float2 / float4 / uint4 (header-only, usable from g++ too)

#define PSIM_BINS 1000           // reference Material::NUM_FREQ_BINS (material.h:14)
#define PSIM_MAX_COLLISIONS 100  // reference MAX_COLLISIONS (modelSimulator.cpp:25)
#define PSIM_FLUX_FRAC_BITS 8    // flux tallies are int64 fixed point, 1/256 m/s resolution
#define PSIM_FREQ_SCALE 1e-13    // angular frequencies are carried as omega * 1e-13 (fp32 range)
#define PSIM_BIRTH_STEP 0xFFFFFFFFu  // Philox stream selector for the draws made at emission
#ifndef PSIM_PHILOX_ROUNDS
#define PSIM_PHILOX_ROUNDS 10  // Philox4x32-10 (Salmon et al. 2011)
#endif
#ifndef PSIM_GUIDE
#define PSIM_GUIDE 1024          // entries of the per-table guide that brackets the inverse-CDF search
#endif

// edge link word: [31:30] kind, then payload
#define PSIM_LINK_BOUNDARY 0u    // [0] the wall is perfectly specular (specularity >= 1): a mirror reflection, no random number
#define PSIM_LINK_TRANSITION 1u  // [29:28] neighbour edge, [27] same-direction flag, [26] neighbour is a parallelogram,
                                 // [25] PSIM_LINK_SAME_FRAME, [24:0] neighbour flight cell
#define PSIM_LINK_EMIT 2u        // [26:0] emitter index
#define PSIM_LINK_COMPOSITE 3u   // [27] some sub-surface is a transition into a cell of the same material and rate class (worth
                                 // trying the flight loop's fast path), [26:7] first sub-surface, [6:0] number of sub-surfaces
#define PSIM_LINK_KIND(w) ((w) >> 30)
#define PSIM_LINK_INDEX(w) ((w)&0x07FFFFFFu)                          // emitter index
#define PSIM_LINK_TARGET(w) ((w)&0x01FFFFFFu)                          // neighbour flight cell
#define PSIM_LINK_SAME_FRAME (1u << 25)  // the neighbour has the same shape record, material and rate class as this cell
#define PSIM_LINK_CELL(w) (PSIM_LINK_TARGET(w) | (((w) >> 26 & 1u) << 31))  // tagged flight cell word of the neighbour

// A phonon's cell word: [24:0] flight cell, [31] that cell is a parallelogram (the flight needs to know before it has
// loaded anything)
#define PSIM_CELL_INDEX(c) ((c)&0x03FFFFFFu)
#define PSIM_CELL_QUAD(c) ((c) >> 31)

// DevCell::sensor_mat.  Rate class: sensors with the same material and the same temperature have identical
// relaxation rates; 255 = unclassified (more than 255 distinct classes in the model).
#define PSIM_CELL_SENSOR(w) ((w) >> 12)
#define PSIM_CELL_CLASS(w) (((w) >> 4) & 0xFFu)
#define PSIM_CELL_MAT(w) ((w)&0xFu)

#if defined(__CUDACC__)
#define PSIM_ALIGN(n) __align__(n)
#else
#define PSIM_ALIGN(n) alignas(n)
#endif

// What every cell transition needs, one 32-byte sector per flight cell: what lies behind its (up to four) edges, which
// sensor area / rate class / material it belongs to, its shape, and the model-file triangles it stands for.  The cell's
// GEOMETRY (frame matrix, inward normals) and its specularity live in a table of distinct SHAPES: the meshes the
// reference's builder makes are unions of rectangles, so thousands of cells share a few dozen shapes (kinked wire: 3087
// flight cells, 13 shapes) and the records a flight touches stay in L1 where the 64 B per triangle of the first layout
// did not (profiles/r01_summary.md: L1 hit rate 52 % on that mesh; 99 % now).  A mesh of arbitrary triangles simply has
// as many shapes as cells.
struct PSIM_ALIGN(16) DevCell {
    uint32_t link[4];          // what lies behind each edge (a triangle has three)
    uint32_t sensor_mat;       // [31:12] sensor index, [11:4] rate class, [3:0] material index (PSIM_CELL_*)
    uint32_t shape;            // index into DevParams::shapes
    uint32_t tri[2];           // model-file cell(s): a triangle names itself twice; a parallelogram its triangle below the
                               // diagonal Q0-Q2 (b1 >= b2) and the one above.
                               // In the LATTICE image (below) instead: [0] first entry of the cell in DevParams::sub_fine,
                               // [1] nx | ny << 16, the cell being a lattice of nx x ny identical parallelograms.
};

// The LATTICE image of a mesh (flatten.cpp: build_lattices), used by launches whose window records nothing - the first 90 % of
// the measurement steps of a steady-state run.  Where nothing is recorded a phonon's sensor area is irrelevant; what is left of
// a cell transition is the change of material / relaxation rates / wall specularity, and between the parallelograms of one
// rate class that the reference's builder lines up in rows and columns there is none.  So a rectangular block of nx x ny
// identical parallelograms (same shape record, same material and rate class, linked whole edge to whole edge) flies as ONE cell
// whose frame spans the block: b1 in [0, 1] covers the nx columns.  Its outer edges are composite surfaces (one sub-surface per
// fine edge that is not a plain wall).  The pool holds lattice coordinates while such launches run; the first launch that
// records converts every phonon it fetches back to its fine flight cell (coarse_to_fine, device_core.cuh).
// Where a phonon crosses several fine cells per measurement step (the kinked wire: 3.8 x 5.1 nm cells, 10 - 18 nm per step)
// recorded windows fly the lattice image too (psim_gpu.cu decides per model): a flight segment then spans several sensor
// areas, and the area of every measurement it crossed is found from the phonon's position at that instant
// (lattice_sensor_at); runs of equal areas are tallied in difference form.

struct PSIM_ALIGN(16) DevShape {
    float m00, m01, m10, m11;  // d(b1)/dt = m00 vx + m01 vy ; d(b2)/dt = m10 vx + m11 vy   (inverse of [u | v])
    float n[4][2];             // unit normals of edges 0..3 pointing INTO the cell (geometry.cpp:97-100)
    float spec;                // specularity of the cell's boundary surfaces, clamped to [0,1] (cell.cpp:115-119)
    uint32_t pad[3];
};

// A model-file (API) cell: the flight cell it lives in and where its three vertices sit in that cell's frame, two bits
// (b1, b2 in {0, 1}) per vertex - what a phonon born "at a random point of cell c" (CellOriginBuilder) needs.
struct DevApiCell {
    uint32_t cell;             // tagged flight cell word (PSIM_CELL_*)
    uint32_t corners;          // [1:0] vertex 1, [3:2] vertex 2, [5:4] vertex 3; bit 0 of a pair = b1, bit 1 = b2
                               // lattice image: also [18:6] column and [31:19] row of the cell's parallelogram in its lattice
};

// A part of an edge that is a transition to a neighbour or an emitting surface (compositeSurface.h:60-66).
struct PSIM_ALIGN(16) DevSub {
    float s0, s1;   // extent along the edge, as fractions of the edge from its first to its second vertex
    float a, b;     // transition: position along the neighbour's edge = a * s + b
    uint32_t link;  // PSIM_LINK_TRANSITION or PSIM_LINK_EMIT word
    uint32_t pad[3];
};

// Relaxation-rate coefficients of one sensor area, in 1/ns, for omega in units of 1e13 rad/s
// (material.cpp:207-239 evaluated at the sensor's temperature, sensorController.h:68-70,86-88).
struct PSIM_ALIGN(16) DevSensor {
    float c_la;     // LA: N = U = c_la w^2
    float c_tn;     // TA, w <  w_cut: N = c_tn w
    float c_tu;     // TA, w >= w_cut: U = c_tu w^2 / sinh(x_t w)
    float x_t;      // hbar * 1e13 / (k_B T)
    float c_i;      // impurity: I = c_i w^4
    float w_cut;
    uint32_t scatter_table;  // table sampled on an intrinsic scatter
    uint32_t base_table;     // table sampled for phonons born inside a cell
};

struct DevMaterial {
    float w_max_la, w_max_ta;  // scaled; a phonon above the neighbour's cutoff back-scatters at an interface
    float freq_width;          // scaled bin width
    float pad;
};

struct PSIM_ALIGN(16) DevEmitter {
    uint32_t cell, edge;    // tagged flight cell word and its edge
    float s_p1, s_p2;       // edge coordinate of the surface's two end points (Line::getRandPoint, geometry.cpp:140-143)
    uint32_t table;         // velocity-weighted table at the surface temperature
    uint32_t pad;
    uint32_t k_on, k_off;   // absorbing while k_on <= measurement step < k_off (surface.cpp:61-65)
    double start, duration; // emission window in ns
};

struct DevSource {
    uint32_t kind;   // 0: cell interior, 1: emitting surface, 2: phasor surface
    uint32_t index;  // cell or emitter index
    int32_t sign;    // +1 if the source is hotter than t_eq, -1 otherwise (phononBuilder.cpp:8,33)
    uint32_t pad;
    uint64_t count;  // phonons of this source over the whole run
    uint64_t first_id;
};

// One (measurement step, source) group of phonons this shard has to create.
struct DevBirth {
    uint32_t source;
    uint32_t step;
    uint64_t j0;      // first source-local phonon index owned by this shard
    uint32_t count;   // phonons created for it: j0, j0 + stride, ...
    uint32_t stride;  // = number of shards
};

// Phonon state in HBM: two 16-byte words per phonon, structure-of-arrays.
//   A = (b1, b2, vx, vy)            position in the cell frame, in-plane velocity = group velocity x projected 3-D direction
//   B = (tts, packed, cell, id)     tts = time to the next intrinsic scatter (ns); cell = tagged flight cell word
//                                   packed = [9:0] bin, [10] polarisation (1 = TA), [11] sign (1 = negative),
//                                            [15:12] material the (omega, v) pair was sampled in,
//                                            [23:16] position of omega inside the bin (1/256ths), [31:24] id bits 39:32
#define PSIM_PACK_BIN(p) ((p)&0x3FFu)
#define PSIM_PACK_TA(p) (((p) >> 10) & 1u)
#define PSIM_PACK_NEG(p) (((p) >> 11) & 1u)
#define PSIM_PACK_MAT(p) (((p) >> 12) & 0xFu)
#define PSIM_PACK_JIT(p) (((p) >> 16) & 0xFFu)
#define PSIM_PACK_IDHI(p) ((p) >> 24)

struct DevParams {
    const DevCell* cells;       // [n_flight_cells]
    const DevShape* shapes;     // [n_shapes] distinct (geometry, specularity) records
    const DevApiCell* api_cells; // [n_cells] model-file cells
    const DevSub* subs;
    const DevSensor* sensors;
    const DevSensor* classes;   // [<= 255] one record per rate class (PSIM_CELL_CLASS): the sensors of a class are identical
    const DevSensor* step_sensors;  // NULL, or [n_sensors][num_steps]: per-step records of a re-iterated transient run
    const DevMaterial* materials;
    const DevEmitter* emitters;
    const DevSource* sources;
    const float2* tables;      // [n_tables][PSIM_BINS] (cumulative probability, LA fraction)  (material.cpp:170-180)
    const uint32_t* guides;    // [n_tables][PSIM_GUIDE] search bracket for r in [k/G, (k+1)/G), G = PSIM_GUIDE: low | high << 16
    const float* velocities;   // [n_materials][2][PSIM_BINS] group velocity, LA then TA, m/s == nm/ns
    const uint32_t* sub_fine;  // lattice image only: tagged fine flight-cell word of every parallelogram of every lattice cell, rows first
    const uint32_t* sub_sensor; // lattice image only: ... and its sensor index (recorded windows flown over the lattice image)
    uint32_t n_cells, n_flight_cells, n_shapes, n_sensors, n_materials, n_tables, n_emitters, n_sources;
    uint32_t num_steps;        // measurement steps M
    uint32_t first_tally_step; // reference step_adjustment_ (modelSimulator.h:24-26)
    uint32_t recorded_steps;   // M - first_tally_step
    uint32_t full_mode;        // t_eq == 0: bin-centre frequencies, no jitter (material.cpp:77-80)
    uint32_t phasor;           // phasor_sim: no intrinsic scattering (modelSimulator.cpp:189)
    uint32_t lattice;          // cells / shapes / api_cells / subs / emitters are those of the lattice image
    uint32_t fast_links;       // some edge of this image can be handled by the flight loop's fast path (fast_impact); 0: the flight
                               // loop does not even try (the lattice image of a mesh that is ONE block per material: every impact
                               // is a wall, an emitter or a material interface)
    float step_time;           // ns
    float step_time_inv;
    double step_time_d;
    uint32_t seed_lo, seed_hi;
};

#endif
