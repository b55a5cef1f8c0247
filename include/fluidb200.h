/*
 * fluidb200.h -- C ABI of libfluidb200.so: the B200 (sm_100a) implementation of
 * the per-step hot path of TheFellow/fluid's Go package pkg/fluid.
 *
 * This is the drop-in boundary: a cgo (or ctypes) binding of these entry points
 * backs the exported Go API of pkg/fluid.  Plain pointers
 * and sizes only; no C++/torch types.  Every call returns FB_OK (0) or a
 * negative fb_status, never throws, never aborts.  A handle is not thread-safe;
 * every entry point selects the handle's CUDA device itself, so calls may come
 * from any OS thread (Go moves goroutines between threads).
 *
 * Citations are file:line relative to the reference checkout.
 *
 * Grid: NumX = width+2, NumY = height+2 (pkg/fluid/fluid.go:42-50); host-side
 * arrays are dense row-major [NumX][NumY] float32 with index i*NumY+j exactly
 * as the reference's slices (fluid.go:189).  The device layout is private
 * (padded pitch).
 */
#ifndef FLUIDB200_H
#define FLUIDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fb_handle fb_handle;

typedef enum fb_status {
    FB_OK = 0,
    FB_ERR_INVALID = -1,     /* bad argument (the reference panics: walls.go:6-11) */
    FB_ERR_CUDA = -2,        /* CUDA runtime error; see fb_last_error */
    FB_ERR_NOMEM = -3,
    FB_ERR_UNSUPPORTED = -4, /* e.g. exact solver with nranks > 1 */
    FB_ERR_HALO = -5         /* a semi-Lagrangian trace left the ghost zone */
} fb_status;

/* Persistent arrays of `type Fluid` (fluid.go:17-23). */
typedef enum fb_field {
    FB_U = 0, FB_V = 1, FB_NEWU = 2, FB_NEWV = 3,
    FB_P = 4, FB_S = 5, FB_M = 6, FB_NEWM = 7,
    FB_NFIELDS = 8
} fb_field;

/* Pressure solver used by makeIncompressible (fluid.go:144-234). */
typedef enum fb_solver {
    /* Lexicographic in-place Gauss-Seidel/SOR, reproduced bit for bit by a
     * skewed-tile wavefront (i + j + 2*sweep ordering).  Single GPU only. */
    FB_SOLVER_EXACT = 0,
    /* The same per-cell update (fluid.go:196-229) in red-black order, all iterations fused in one
     * pass.  Order-independent, slab-decomposable.  NOT the reference's relaxation schedule on the
     * last iteration: iterations 0 .. n-2 use omega(iter) of fluid.go:169-170 on both colours, the
     * last one closes with omega = 1.0 (red) and 0.5 (black) whatever fb_params.relaxation is
     * (omega_schedule_redblack in csrc/fluidb200.cu; DESIGN.md 4.1 says why).  Parity claim of the
     * red-black modes: max|div| after the solve <= the reference's 8 lexicographic sweeps on the
     * same input (bench.py prints both); fields differ from the reference's by the size of the
     * solver residual.  Bit-exact only against this repository's restatement of the same ordering. */
    FB_SOLVER_REDBLACK = 1,
    /* The same red-black iteration and schedule in pressure form: one scalar per cell circulates
     * through the fused iterations and U, V, p are materialised once.  Algebraically identical to
     * FB_SOLVER_REDBLACK, rounding differs at the 1e-6 level.  The throughput solver. */
    FB_SOLVER_REDBLACK_PRESSURE = 2
} fb_solver;

/* fb_create arguments; replaces fluid.New(density, width, height, h) (fluid.go:42). */
typedef struct fb_config {
    int32_t width, height;   /* GLOBAL interior size */
    float density, h;
    int32_t device;          /* CUDA device ordinal */
    int32_t rank, nranks;    /* row-slab decomposition over i; 0,1 = single GPU */
    int32_t ghost;           /* ghost lines per side when nranks > 1 (0 = default) */
    int32_t flags;           /* fb_flags, 0 = default */
} fb_config;

typedef enum fb_flags {
    /* Reference-shaped kernels with the reference's physical array copies (the first,
     * unfused implementation): kept for A/B parity checks of the fused path. */
    FB_FLAG_LITERAL = 1,
    /* Keep the scratch arrays newU/newV/newM complete after every advection, as the
     * reference's copy(f.U, f.newU) does (fluid.go:331-332, 433).  Needed only when a
     * caller rewrites S directly (white-box tests); the exported API never needs it
     * because SetSolid zeroes both buffers (walls.go:19-48).  Costs 3 plane copies a step. */
    FB_FLAG_EXACT_SHADOW = 2
} fb_flags;

/* The exported knobs of `type Fluid` (fluid.go:25-39), package var Relaxation
 * (fluid.go:7-9) and numIters (fluid.go:81).  Passed by value with every step
 * because Go callers mutate the struct fields directly (main/main.go:271-278). */
typedef struct fb_params {
    float relaxation;            /* 1.9 */
    float confinement;           /* 0 */
    float viscosity_diffusion;   /* 0 */
    float pressure_damping;      /* 1 */
    float turbulence_strength;   /* 0.02 */
    float smoke_advection;       /* 1 */
    int32_t use_multigrid;       /* false (fluid.go:64).  With multigrid_levels > 1 makeIncompressible runs
                                  * solveMultigridVCycle (fluid.go:560-599): `iters` V-cycles.  Single GPU only. */
    int32_t multigrid_levels;    /* 2 */
    int32_t use_bfecc;           /* false */
    int32_t solver;              /* fb_solver */
    int32_t iters;               /* numIters; 0 = 8 (fluid.go:81) */
} fb_params;

/* Edit commands: the point edits of walls.go:5-93 and fluid.go:761-771,
 * 894-907 generalised to half-open rectangles [i0,i1) x [j0,j1) so that presets
 * on 16384^2 grids do not need 10^8 calls.  Commands apply in list order. */
typedef enum fb_edit_op {
    FB_EDIT_SET_SOLID = 0,      /* a != 0 -> solid; zeroes 4 faces in U,V,newU,newV (walls.go:19-48) */
    FB_EDIT_SET_VELOCITY = 1,   /* U=a, V=b (walls.go:62-72) */
    FB_EDIT_ADD_SMOKE = 2,      /* M += a (walls.go:74-83) */
    FB_EDIT_APPLY_FORCE = 3,    /* U+=a, V+=b unless ring or solid (fluid.go:761-771) */
    FB_EDIT_CIRCLE_OBSTACLE = 4,/* centre (i0,j0), radius i1 (fluid.go:894-907) */
    FB_EDIT_RESET = 5,          /* walls.go:85-93: zero all but S */
    FB_EDIT_SET_VELOCITY_IF_FLUID = 6, /* main/main.go:475-477 source guard */
    FB_EDIT_ADD_SMOKE_IF_FLUID = 7,
    FB_EDIT_SET_SMOKE = 8       /* M = a (fluid_bench_test.go:39 writes M directly) */
} fb_edit_op;

typedef struct fb_edit_cmd {
    int32_t op;
    int32_t i0, j0, i1, j1;
    float a, b;
} fb_edit_cmd;

/* Single phases of Simulate, for the reference's white-box tests
 * (fluid_test.go:56,117,179,239,310,332,1068-1083) and per-kernel parity. */
typedef enum fb_phase_id {
    FB_PHASE_MAKE_INCOMPRESSIBLE = 0,  /* fluid.go:144; uses `iters` */
    FB_PHASE_ADVECT_VELOCITY = 1,      /* fluid.go:291 */
    FB_PHASE_ADVECT_SMOKE = 2,         /* fluid.go:400 */
    FB_PHASE_HANDLE_BORDERS = 3,       /* fluid.go:236 */
    FB_PHASE_CONFINEMENT = 4,          /* fluid.go:449 */
    FB_PHASE_TURBULENCE = 5,           /* fluid.go:496 */
    FB_PHASE_ADVECT_VELOCITY_BFECC = 6,/* fluid.go:911 */
    FB_PHASE_ADVECT_SMOKE_BFECC = 7,   /* fluid.go:997 */
    FB_PHASE_VISCOSITY = 8,            /* fluid.go:112 */
    FB_PHASE_CLEAR_PRESSURE = 9,       /* fluid.go:83 */
    FB_PHASE_PROJECT = 10              /* fluid.go:83 + 90 as Simulate runs them: fill(p,0) then makeIncompressible(iters);
                                        * the fused solvers write p without reading it, so the fill costs no pass over HBM */
} fb_phase_id;

typedef enum fb_view_kind {
    FB_VIEW_SMOKE = 0,              /* smoke.go:5 */
    FB_VIEW_PRESSURE = 1,           /* pressure.go:5 */
    FB_VIEW_VELOCITY_MAGNITUDE = 2, /* fluid.go:841 */
    FB_VIEW_VORTICITY = 3           /* fluid.go:806 */
} fb_view_kind;

typedef enum fb_reduce_kind {
    FB_REDUCE_MAX_DIVERGENCE = 0,   /* fluid.go:876 */
    FB_REDUCE_MAX_ABS_VELOCITY = 1  /* max(|U|+|V|) of fluid.go:534-543 */
} fb_reduce_kind;

typedef struct fb_solve_stats {
    int32_t sweeps_run;       /* sweeps executed by the last solve (early exit, fluid.go:175);
                               * multigrid: V-cycles entered (early exit, fluid.go:575) */
    int32_t rolled_back;      /* exact solver: 1 if the early exit forced a re-run */
    float max_div[32];        /* max pre-update |div| seen in each sweep (fluid.go:209-216);
                               * multigrid: that of the third pre-smoothing sweep of each cycle */
} fb_solve_stats;

/* ---- lifecycle ---------------------------------------------------------- */
int fb_create(const fb_config *cfg, fb_handle **out);          /* fluid.New, fluid.go:42 */
int fb_destroy(fb_handle *h);
const char *fb_last_error(const fb_handle *h);                 /* NUL-terminated, owned by h */
int fb_default_params(fb_params *p);                           /* fluid.go:59-66, 7-9 */
int fb_dims(const fb_handle *h, int64_t *num_x, int64_t *num_y,
            int64_t *i_lo, int64_t *i_hi);                     /* global dims + owned slab */

/* ---- the hot path ------------------------------------------------------- */
/* (*Fluid).Simulate(dt) x nsteps (fluid.go:79-109).  `per_step` (may be NULL)
 * is replayed before every step: the jet / source / sink re-imposition that
 * main/main.go:233-241,474-486 performs before each Simulate. */
int fb_step(fb_handle *h, const fb_params *p, float dt, int32_t nsteps,
            const fb_edit_cmd *per_step, size_t n_per_step);
int fb_phase(fb_handle *h, int32_t phase, const fb_params *p, float dt, uint32_t iters);
/* One Simulate on this rank's slab (nranks > 1; also valid for nranks == 1).  The host
 * layer exchanges `ghost` lines of U, V, M with both neighbours BEFORE each call
 * (fb_halo_region); inside the step nothing is communicated: each phase is recomputed on
 * as many ghost lines as later phases read.  `reach` bounds how many lines a trace may
 * travel in one step, ceil(dt*max|u|/h) + 2; FB_ERR_HALO if the ghost zone cannot cover it
 * or a trace leaves the lines this rank holds.  Red-black solvers only. */
int fb_step_local(fb_handle *h, const fb_params *p, float dt, int32_t reach,
                  const fb_edit_cmd *per_step, size_t n_per_step);
int fb_get_solve_stats(fb_handle *h, fb_solve_stats *out);

/* ---- edits (walls.go, fluid.go:761-796, 894-907) ------------------------ */
int fb_edit(fb_handle *h, const fb_edit_cmd *cmds, size_t n);
/* ApplyForceRadius (fluid.go:774-796): Gaussian weights exp(-3 d^2/r^2) are
 * evaluated on the host in double precision like the reference. */
int fb_apply_force_radius(fb_handle *h, int32_t cx, int32_t cy, float fx, float fy, int32_t radius);

/* ---- field transfer (dense [NumX_global][NumY] host arrays) ------------- */
/* Copies this rank's owned lines (all lines when nranks == 1).  `host` points at
 * the start of the dense GLOBAL array. */
int fb_upload(fb_handle *h, int32_t field, const float *host);
int fb_download(fb_handle *h, int32_t field, float *host);
/* Pinned host mirror owned by the library, dense global layout; a Go binding
 * wraps it with unsafe.Slice to back the exported U/V/S/M slices. */
int fb_host_mirror(fb_handle *h, int32_t field, float **ptr, size_t *count);

/* ---- views and reductions (Q-14 semantics) ------------------------------ */
int fb_view(fb_handle *h, int32_t kind, float *out_or_null, float *min_value, float *max_value);
/* Pipelined form of fb_view for a frame loop: `begin` queues the view behind the work already
 * submitted and returns at once; the transfer into `out` (pinned host memory, dense global
 * layout) runs on a second stream and overlaps whatever is queued next (the next Simulate);
 * `end` waits for it and returns min / max.  One view may be in flight per handle.  This is
 * main/main.go's draw-frame-k-while-computing-k+1 loop: Smoke() at main/main.go:247. */
int fb_view_begin(fb_handle *h, int32_t kind, float *out);
int fb_view_end(fb_handle *h, float *min_value, float *max_value);
/* The same pipelined view, decimated and quantised for frame loops whose field is larger than any display (and for N
 * GPUs sharing one host's PCIe): min / max over the FULL field as above, then every stride-th cell of every stride-th
 * line (global indices that are multiples of stride) as one byte,
 *   index = (unsigned)(min(max((v - min) * (255 / (max - min)), 0), 255) + 0.5)      (0 when max == min),
 * written densely to `out` as [out_lines][out_cols] (this rank's lines only; pinned host memory).  The caller maps the
 * index through a 256-entry table of main/colors.go's palette.  Ends with fb_view_end.  The full-field float path
 * (fb_view / fb_view_begin) stays the parity path. */
int fb_view_u8_begin(fb_handle *h, int32_t kind, int32_t stride, uint8_t *out, int32_t *out_lines, int32_t *out_cols);
/* ---- the frame loop either side of Simulate: Draw's pixel pass and advectParticles ---------
 * fb_render*: the view through the UI's colormap (main/colors.go:8-84: getSciValue; getDivergingColor
 * for vorticity) with solid cells black (main/main.go:564-574), as the RGBA image Draw hands to the
 * renderer: [NumY rows][NumX pixels][4 bytes], row jj showing fluid column NumY-1-jj
 * (fluidToImageIndex, main/main.go:795).  `range` = {min, max} to colour with, or NULL for the view's
 * own min / max; begin / end pipeline exactly like fb_view_begin / fb_view_end (one in flight, shared). */
int fb_render_begin(fb_handle *h, int32_t kind, uint8_t *rgba_out, const float *range_or_null);
int fb_render_end(fb_handle *h, float *min_value, float *max_value);
int fb_render(fb_handle *h, int32_t kind, uint8_t *rgba_out, const float *range_or_null, float *min_value, float *max_value);
/* advectParticles (main/main.go:512-546): age, RK2 midpoint through SampleVelocity, bounds and solid
 * checks, for n particles in host memory, in place; survivors keep their order, *n_alive counts them. */
typedef struct fb_particle {   /* main/main.go:139-144 */
    float x, y;
    uint8_t r, g, b, pad;
    float age, max_age;
} fb_particle;
int fb_advect_particles(fb_handle *h, fb_particle *particles, size_t n, float dt, size_t *n_alive);
int fb_reduce(fb_handle *h, int32_t kind, float *out);
/* SampleVelocity (fluid.go:799-803) for n points; xy and uv are [n][2]. */
int fb_sample_velocity(fb_handle *h, size_t n, const float *xy, float *uv);

/* ---- slab halo exchange (nranks > 1) ------------------------------------ */
/* Device pointers to the `lines` ghost lines to RECEIVE from side (0 = lower i,
 * 1 = upper i) and to the owned lines to SEND to that side; each region is
 * `lines` contiguous rows of `pitch` floats.  The transport (NCCL send/recv,
 * peer copies) belongs to the host layer. */
int fb_halo_region(fb_handle *h, int32_t field, int32_t side, int32_t lines,
                   void **send_ptr, void **recv_ptr, size_t *bytes);
int fb_ghost_lines(const fb_handle *h, int32_t *ghost);
/* Halo exchange through peer memory (NVLink P2P), no collective and no host hand-shake.
 *   fb_halo_export   allocate this rank's send buffer for `lines` lines per side of U, V, M and
 *                    return its CUDA IPC handle (64 bytes) and its device address;
 *   fb_halo_connect  attach the send buffer of the neighbour on `side` (0 = lower i): by IPC
 *                    handle (another process on the node) or by device address (same process);
 *   fb_halo_post     pack the boundary lines and publish the exchange epoch (never waits);
 *   fb_halo_pull     copy the ghost lines out of the neighbours' buffers; the kernel waits on
 *                    their epoch flags (bounded: a neighbour that never posts -> FB_ERR_HALO);
 *   fb_halo_exchange post + pull.  Every rank calls it once before each fb_step_local.
 * All work is queued on the handle's stream.  Handles of ONE process must post all before any
 * pulls (the pull would otherwise spin in front of the post it waits for). */
int fb_halo_export(fb_handle *h, int32_t lines, void *ipc_handle64, uint64_t *device_ptr, size_t *bytes);
int fb_halo_connect(fb_handle *h, int32_t side, const void *ipc_handle64_or_null, uint64_t device_ptr);
/* neighbour handle of the same process: its post is awaited with an event instead of a spinning pull */
int fb_halo_connect_local(fb_handle *h, int32_t side, fb_handle *peer);
int fb_halo_post(fb_handle *h);
int fb_halo_pull(fb_handle *h);
int fb_halo_exchange(fb_handle *h);
/* FB_ERR_HALO if any trace since the last check read outside the lines this rank holds
 * (synchronises the stream; call it every few steps, not every step). */
int fb_check_halo(fb_handle *h);

/* ---- plumbing ----------------------------------------------------------- */
int fb_stream(fb_handle *h, void **cuda_stream);   /* the stream all work is queued on */
int fb_synchronize(fb_handle *h);
/* CUDA-event timing on the handle's stream. */
int fb_timer_start(fb_handle *h);
int fb_timer_stop(fb_handle *h, float *elapsed_ms);
/* Per-phase device time of fb_step, measured with CUDA events on the handle's
 * stream around every phase (negligible cost; off by default).  fb_profile_read
 * waits for the stream, adds up the pairs recorded since the last read into
 * ms[FB_PROF_NPHASES] / calls[FB_PROF_NPHASES] and clears them. */
typedef enum fb_prof_phase {
    FB_PROF_EDITS = 0, FB_PROF_CLEAR_PRESSURE, FB_PROF_VISCOSITY, FB_PROF_PROJECT, FB_PROF_CONFINEMENT,
    FB_PROF_TURBULENCE, FB_PROF_BORDERS, FB_PROF_ADVECT_VELOCITY, FB_PROF_ADVECT_SMOKE,
    /* single kernels, one pair per launch (they lie inside the phase pairs above): the fused pressure solve, the
     * semi-Lagrangian passes (k_advect_*_full), the BFECC back-trace + correct passes, confinement + turbulence */
    FB_PROF_K_PRESSURE_SOLVE, FB_PROF_K_ADVECT_VELOCITY, FB_PROF_K_BFECC_VELOCITY, FB_PROF_K_ADVECT_SMOKE,
    FB_PROF_K_BFECC_SMOKE, FB_PROF_K_CONFINE_TURBULENCE,
    /* slabs: fb_halo_exchange in front of a step, or -- with FB_OPT_HALO_OVERLAP -- how long the end of fb_step_local
     * waits for the second stream's exchange */
    FB_PROF_HALO,
    FB_PROF_NPHASES
} fb_prof_phase;
int fb_profile_enable(fb_handle *h, int32_t on);
int fb_profile_read(fb_handle *h, float *ms, int32_t *calls);
/* Options.  FB_OPT_SOLVE_STATS (default 1): the fused red-black solvers also track the
 * per-iteration max |div| that fb_get_solve_stats reports (a few instructions per cell
 * update); throughput runs may switch it off. */
/* FB_OPT_HALO_OVERLAP (default 0; slabs connected through fb_halo_connect only): fb_step_local itself refreshes the
 * ghost lines for the NEXT step while it finishes this one -- U, V are packed, published and pulled on a second stream
 * during the smoke passes, the last smoke pass computes the boundary strips first and M follows during its interior
 * -- so consecutive fb_step_local calls need no fb_halo_exchange between them.  The caller still exchanges once before
 * the first step and after anything else that changes U, V or M (edits from the host, fb_upload, fb_project). */
typedef enum fb_option { FB_OPT_SOLVE_STATS = 0, FB_OPT_HALO_OVERLAP = 1 } fb_option;
int fb_set_option(fb_handle *h, int32_t option, int32_t value);
/* Kernels launched by this handle since creation (bench.py's gpu_launches). */
int fb_launch_count(const fb_handle *h, uint64_t *count);
/* Diagnostics (no counterpart in the reference).  The confinement pass computes its IEEE divisions and square roots
 * (fluid.go:462-463, 474-480, 484, 509) with the instruction sequences of the hardware's own fast path written out as
 * straight-line code, valid on a stated operand range and re-done with div.rn / sqrt.rn outside it.  This runs both on
 * n random operand sets on `device`: mode 0 = random bit patterns, mode 1 = operands inside the accepted range; modes 2
 * (random bit patterns) and 3 (tiny operands, subnormals included) test the fall-back's scaled sequences, which accept everything.
 * counts6 = {quotients, accepted, accepted-and-different (must be 0), roots, accepted, accepted-and-different (must be 0)}. */
int fb_selftest_fastmath(int32_t device, uint64_t n, uint32_t seed, int32_t mode, uint64_t *counts6);
int fb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FLUIDB200_H */
