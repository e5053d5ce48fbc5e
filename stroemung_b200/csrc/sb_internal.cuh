// sb_internal.cuh -- shared definitions of the stroemung_b200 CUDA library.
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/stroemung_b200.h"

namespace sb {

// ---- per-cell flag byte (device representation of Cell + Option<EdgeType>) -----
// bits 0-2 kind (sb_kind), bit 7 = cell exists; bits 3-6: for a boundary cell its edge
// class (sb_edge), for a fluid cell bit 3 = "has a non-fluid 4-neighbour" (CF_NEAR).
// 0x00 therefore means "outside the grid" -- what out-of-bounds reads and unowned slab
// rows produce -- and such cells are inert everywhere.
constexpr uint8_t CF_VALID = 0x80;
constexpr uint8_t CF_FLUID = CF_VALID | SB_KIND_FLUID;
constexpr uint8_t CF_NEAR = 0x08;
__host__ __device__ __forceinline__ int cf_kind(uint8_t f) { return f & 7; }
__host__ __device__ __forceinline__ bool cf_is_fluid(uint8_t f) { return (f & 0x87) == CF_FLUID; }
// edge class of a boundary cell (0 for fluid cells, whose bits 3-6 mean something else)
__host__ __device__ __forceinline__ int cf_edge(uint8_t f) {
    return cf_is_fluid(f) ? 0 : (f >> 3) & 15;
}
// a boundary cell of the grid (present and not fluid)
__host__ __device__ __forceinline__ bool cf_is_boundary(uint8_t f) {
    return (f & CF_VALID) && (f & 7) != SB_KIND_FLUID;
}

// geometry of one slab (single GPU: the whole grid), passed by value to kernels
struct Geom {
    int64_t NX, NY;   // global grid size
    int64_t nxl;      // local rows (owned + 2*halo)
    int64_t pitch;    // elements per row (multiple of 16)
    int64_t gx0;      // global x of local row 0 (may be negative in slab mode)
    int64_t own0, own1;  // owned local rows [own0, own1)
};

// boundary list (BoundaryList, src/grid/mod.rs:61-69) as device SoA, x-major sorted
struct BList {
    uint64_t n = 0, cap = 0;
    int64_t *lin = nullptr;  // local linear index lx*pitch + y
    uint8_t *ke = nullptr;   // kind | edge << 3
    double *bu = nullptr, *bv = nullptr;     // Inflow / MovingWall velocity
    double *ru = nullptr, *rv = nullptr;     // values u[b], v[b] take after set_u_and_v (u_v_restore)
    double *nu = nullptr, *nv = nullptr;     // scratch: new u[b], v[b] of the BC gather
    double *wu = nullptr, *wv = nullptr;     // scratch: values for u[west nbr], v[north nbr]
};

struct SorCtl {        // device-resident control block of one SOR solve
    int32_t active_T;  // sweeps the next pass performs (0 = solve finished / idle)
    int32_t src;       // index of the current pressure buffer
    uint32_t iters_done;
    uint32_t max_iterations;
    int32_t block_T;   // configured temporal block
    int32_t cap_hit;   // finished by reaching max_iterations
    int32_t finished;
    int32_t pad;
    double last_norm;
    double norms[8];   // norms of the last pass (diagnostics / sb_sor_sweeps)
};

// ---- row-slab decomposition (slab.cu) ---------------------------------------------------
constexpr int SB_MAX_WORLD = 16;
constexpr int SB_SLAB_HALO = 10;   // = rb_halo_rows(4): halo rows on each side of a slab

// One mailbox slot: a rank's contribution to one all-gather round (slab.cu)
struct MailSlot {
    double vals[8];
    unsigned long long seq;
    unsigned long long pad[7];
};
static_assert(sizeof(MailSlot) == 128, "one slot per 128-byte line");

// What a kernel needs to talk to the other slabs; passed by value.  All pointers are valid
// in THIS process (CUDA IPC mappings or, for handles of the same process, the peers' own
// allocations).  world <= 1: single GPU, nothing here is touched.
struct SlabLink {
    int32_t rank, world;
    int32_t H;                       // halo rows
    int32_t pad;
    MailSlot *mbox[SB_MAX_WORLD];    // [2][SB_MAX_WORLD] slots of every rank (own included)
    unsigned long long *seq;         // this rank's round counter (device memory)
    int32_t *err;                    // set to 1 when a wait timed out (device memory)
    // pressure buffers of the chain neighbours (nullptr at the ends of the chain) and the
    // local row of THEIR array that receives my first / last H owned rows
    double *lo_p[2], *hi_p[2];
    int64_t lo_row0, hi_row0;
};

// ---- red-black pass: which parts of the grid take the streaming kernel (sor_rb_stream.cu) --
// one work item of the streaming kernel: local rows [x0, x1) of the 128-column strip whose
// first column (halo included) is ty0
struct RbItem {
    int32_t x0, x1, ty0, pad;
};
// The lattice of the tile kernel (BX x BY inner regions) split into "slow" tiles -- anything
// with a wall, an obstacle, the grid ring or a slab edge in its footprint; they keep the tile
// kernel -- and runs of "plain" tiles (all-fluid footprint) along x that one warp streams
// through.  Rebuilt when the cell flags or the temporal block change.
struct RbPlan {
    uint64_t epoch = 0;      // sb_sim::flag_epoch it was built for (0 = never)
    int T = 0;               // temporal block it was built for
    int tiles_x = 0, tiles_y = 0;
    int n_slow = 0, n_items = 0;
    int32_t *d_slow = nullptr;   // tile ids of the slow tiles
    RbItem *d_items = nullptr;
    uint8_t *d_plain = nullptr;  // scratch: one byte per tile
    unsigned *d_counter = nullptr;  // CTAs of the streaming kernel done (fused finalize)
    size_t cap_tiles = 0, cap_items = 0;
    // "frozen" tiles: every cell of the inner region and its 4-neighbours is a boundary cell
    // without an edge class (deep inside an obstacle).  Nothing there changes during a
    // solve, so no pass touches them: once per solve their pressures are mirrored into the
    // other buffer and their residual sum (a constant of the solve; the reference's norm
    // counts obstacle cells too, src/simulation.rs:216-227) goes into one extra partial slot.
    int n_frozen = 0;
    int32_t *d_frozen = nullptr;     // tile ids
    double *d_frozen_part = nullptr; // per-tile residual sums
    size_t cap_frozen = 0;
    uint64_t frozen_seq = 0;         // sb_sim::solve_seq the extra slot was filled for
};

// Test hooks and diagnostics, read from the environment ONCE per handle (capi.cu:
// read_debug_knobs, the only getenv of the library) -- none of them changes results, they
// select which of several bit-identical kernels runs (so that the parity tests can drive
// every path on every shape) or print traces.
struct DebugKnobs {
    bool sor_small = true;     // SB_SOR_SMALL=0: small grids off the one-SM solve
    bool sor_mid = true;       // SB_SOR_MID=0: mid-size grids off the grid-resident solve
    int sor_mid_ctas = 0;      // SB_SOR_MID_CTAS=n: force a decomposition (0 = natural)
    int sor_mid_variant = 0;   // SB_SOR_MID_VARIANT=1|2: generic / register-window kernel
    bool rb_stream = true;     // SB_RB_STREAM=0: every tile on the tile kernel
    int rb_stream_kinds = 3;   // SB_RB_STREAM_KINDS: bit 0 wall strips, bit 1 boundary rows
    bool rb_frozen = true;     // SB_RB_FROZEN=0: frozen tiles stay on the tile kernel
    double wall_weight = 0.0;  // SB_WALL_WEIGHT: plan weight of a wall row (0 = default)
    double plan_overhead = -1.0;   // SB_PLAN_OVERHEAD: rows charged per work item (< 0 = default)
    bool trace_plan = false;   // SB_DEBUG_PLAN
    bool trace_mid = false;    // SB_MID_TRACE
    bool trace_fin = false;    // SB_FIN_TRACE
};

#define SB_CUDA(call)                                                               \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess) {                                                    \
            sb::set_error(std::string(#call) + ": " + cudaGetErrorString(_e));     \
            return SB_CUDA_ERROR;                                                   \
        }                                                                           \
    } while (0)

void set_error(const std::string &msg);

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace sb

// the opaque handle of the C ABI
struct sb_sim {
    sb_params prm;
    sb::DebugKnobs dbg;
    int device = 0;
    cudaStream_t stream = nullptr;
    sb::Geom g{};
    int halo = 0;
    // fields; p is double-buffered for the red-black passes
    double *p[2] = {nullptr, nullptr};
    int cur = 0;
    double *u = nullptr, *v = nullptr, *f = nullptr, *gq = nullptr, *rhs = nullptr;
    uint8_t *cflag = nullptr;
    uint64_t flag_epoch = 1;      // bumped by everything that writes cflag
    sb::RbPlan plan;
    size_t field_bytes = 0, flag_bytes = 0;
    sb::BList bl;
    // sparse velocity table as given by the host (global coordinates)
    std::vector<sb_boundary_velocity> velocities;
    double fluid_cells = 0.0;
    // reductions / control
    double *d_partial = nullptr;  // per-block partial sums
    size_t partial_cap = 0;
    double *d_scalars = nullptr;  // small device scratch (results of final reductions)
    double *h_scalars = nullptr;  // pinned mirror
    sb::SorCtl *d_ctl = nullptr, *h_ctl = nullptr;
    int32_t *d_lex_sync = nullptr;  // wavefront kernel: the band ticket
    size_t lex_sync_cap = 0;
    uint4 *d_lex_ll = nullptr;      // wavefront kernel: hand-over rows {lo, seq, hi, seq} per band
    size_t lex_ll_cap = 0;
    uint32_t lex_seq = 0;           // sweeps launched so far (tag of the hand-over elements)
    // classification scratch
    int64_t *d_scan = nullptr;
    size_t scan_cap = 0;
    unsigned long long *d_err = nullptr;  // first too-thin cell (min linear global index)
    // bookkeeping mirrored from the reference's structs
    double time = 0.0;
    uint32_t iterations = 0;
    int has_initial_norm = 0;
    double initial_norm_squared = 0.0;
    double pressure_range[2] = {0, 0}, speed_range[2] = {0, 0};
    double umax = 0.0, vmax = 0.0;  // max |u|, |v| over fluid cells (adaptive delt)
    bool uvmax_valid = false;       // umax / vmax describe the current u, v and fluid set
    // u_v_restore (src/grid/mod.rs:69) holds records: emptied by rebuild_boundary_list
    // (:205), filled by set_boundary_u_and_v; set_u_and_v replays nothing while it is empty
    bool restore_valid = false;
    uint32_t last_sor_iterations = 0;
    double last_norm_squared = 0.0;
    uint32_t sor_batch_hint = 0;
    uint64_t err_xy[2] = {0, 0};
    uint8_t err_kind = 0;
    uint64_t launches = 0;
    cudaEvent_t ev_sor0 = nullptr, ev_sor1 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_stage[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // sb_last_stage_ms
    double last_sor_ms = 0.0;
    // optional per-pass event profiling (sb_profile_enable)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;  // pairs: begin, end
    size_t prof_used = 0;
    // row-slab mode (prm.world > 1): links to the other slabs, see slab.cu
    bool slab = false, connected = false;
    sb::SlabLink link{};
    sb::MailSlot *d_mbox = nullptr;
    unsigned long long *d_xseq = nullptr;
    int32_t *d_xerr = nullptr;
    std::vector<void *> ipc_opened;           // bases returned by cudaIpcOpenMemHandle
    // neighbours' u, v, cflag arrays (halo puts outside the SOR passes)
    double *lo_u = nullptr, *lo_v = nullptr, *hi_u = nullptr, *hi_v = nullptr;
    uint8_t *lo_flag = nullptr, *hi_flag = nullptr;
    double *lo_rhs = nullptr, *hi_rhs = nullptr;
    double *d_hist = nullptr;                 // norm history of sb_sor_sweeps (grows only)
    size_t hist_cap = 0;
    void *d_mid = nullptr;                    // sor_mid.cu: barrier word, exchange rows, partials
    size_t mid_cap = 0;
    int last_sor_path = 0, last_sor_ctas = 0;  // sb_last_sor_path
    bool mid_unavailable = false;             // a cooperative launch was refused on this device
    uint64_t solve_seq = 0;                   // counts solve_sor calls (per-solve caches)
    uchar4 *d_img = nullptr;                  // RGBA8 frame of sb_render_rgba (lazy)
    // tensor maps for the red-black pass (built lazily per buffer)
    bool tmaps_ready = false;
    CUtensorMap tm_p[2], tm_rhs, tm_flag;
};

namespace sb {

// ---- kernels' host launchers (defined in the .cu files) ------------------------
// classify.cu
sb_status classify(sb_sim *s);                 // cflag edges + list + fluid count
sb_status apply_velocity_table(sb_sim *s);     // scatter sparse (x,y,u,v) into the list
// stages.cu
sb_status launch_velocity_bc(sb_sim *s);
sb_status launch_fg(sb_sim *s);
sb_status launch_fg_rhs(sb_sim *s, int what);  // 1: F, G; 3: F, G and RHS fused
sb_status launch_rhs(sb_sim *s);
sb_status launch_pressure_bc(sb_sim *s, int guarded);
sb_status launch_norm_partials(sb_sim *s, int guarded, int *nblocks);
sb_status launch_adapt_uv(sb_sim *s, int with_prange = 0);
sb_status launch_pressure_range(sb_sim *s);
sb_status launch_speed_range(sb_sim *s);
sb_status launch_cellop(int op, const double *u9, const double *v9, const double *scal,
                        double *out);
// sor_lex.cu
sb_status launch_sor_lex_sweep(sb_sim *s, int guarded);
// sor_rb.cu
struct RbFin;
// fin != nullptr and *fused set: the pass also did the work of launch_sor_finalize
sb_status launch_sor_rb_pass(sb_sim *s, int *ntiles, int norm_only, const RbFin *fin = nullptr,
                             int *fused = nullptr);
int rb_halo_rows(int T);
// sor_rb_stream.cu
sb_status rb_ensure_plan(sb_sim *s, int BX, int BY, int h);
sb_status launch_sor_rb_stream(sb_sim *s, int part_base, int part_stride, int h,
                               const RbFin *fin);
void rb_plan_release(sb_sim *s);
int rb_stream_slots_per_item(int T);   // partial-sum slots of one streaming work item
sb_status launch_frozen_mirror(sb_sim *s, int BX, int BY);   // frozen tiles: p[other] := p[cur]
sb_status launch_frozen_fill(sb_sim *s, int part_stride, int slot);
void preload_sor_rb_stream();
// sor_small.cu: the whole red-black solve of a grid that fits one SM's shared memory
bool sor_small_fits(const sb_sim *s);
sb_status launch_sor_small(sb_sim *s, double initial_norm, double eps2, int test_exit,
                           double *norm_hist);
void preload_sor_small();
// sor_mid.cu: the whole red-black solve of a grid that fits the shared memory of all SMs
bool sor_mid_fits(const sb_sim *s);
sb_status launch_sor_mid(sb_sim *s, double initial_norm, double eps2, int test_exit,
                         double *norm_hist);
void preload_sor_mid();
// profiling hooks (capi.cu): record an event of the current pass on the stream
void prof_mark(sb_sim *s);
// finalize (stages.cu): sum partials, exit test, update ctl
sb_status launch_sor_finalize(sb_sim *s, int nparts, double initial_norm, double eps2,
                              int test_exit, double *norm_hist);
// d_scalars[0] = (sum of partial[0..n)) / fluid_cells, read back into *out
sb_status reduce_norm(sb_sim *s, int nparts, double *out);
sb_status restore_edges_from_list(sb_sim *s);
void dump_finalize_trace(int rank);   // SB_FIN_TRACE
sb_status launch_mark_valid(sb_sim *s, const uint8_t *d_kind_rows, int keep_edges);
sb_status launch_preset(sb_sim *s, int preset, const double *args);
sb_status launch_edit_block(sb_sim *s, int64_t gx, int64_t gy, uint8_t kind, double *backup,
                            int restore, int32_t *modified);
// slab.cu -- cross-slab primitives (no-ops / identity for a single-GPU handle)
enum { XR_SUM = 0, XR_MIN = 1, XR_MAX = 2 };
// d_vals[0..n) (device, n <= 8) := reduction over all slabs, op of value i = (ops >> 2i) & 3;
// every slab gets bit-identical results (fixed rank order).  n = 0: barrier only.
sb_status slab_allreduce(sb_sim *s, double *d_vals, int n, unsigned ops);
// my first / last H owned rows of a field -> the neighbours' halo rows (no barrier)
sb_status slab_put_rows(sb_sim *s, const void *field, void *lo_field, void *hi_field,
                        size_t esize);
// p[cur], u, v (and optionally cflag) halos of all slabs made current: barrier, puts, barrier
sb_status slab_sync_halos(sb_sim *s, int with_flags);
sb_status slab_check_error(sb_sim *s);
void preload_classify();
void preload_grid();
void preload_stages();
void preload_sor_rb();
void preload_render();
// render.cu: colour-map the owned rows into d_img (speed = 0: pressure, 1: speed)
sb_status launch_render(sb_sim *s, int speed, uchar4 *d_img);
sb_status slab_prepare(sb_sim *s);   // mailbox etc. of a world > 1 handle
sb_status finish_create(sb_sim *s);  // capi.cu: try_from's work once the arrays are in place
void slab_release(sb_sim *s);

}  // namespace sb
