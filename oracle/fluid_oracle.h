/*
 * fluid_oracle.h -- CPU restatement of TheFellow/fluid's pkg/fluid solver.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under fluid_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * / --impl reference legs use it, and only as the checker or the CPU baseline.
 *
 * PARITY UNPINNED: the reference (pure Go) cannot be compiled here (no Go
 * toolchain) and its tests hold no golden vectors; this restatement is pinned
 * only by re-expressing every assertion of the reference's 26 tests
 * (tests/test_oracle_reference_suite.py).  See DESIGN.md.
 *
 * Semantics: Go on amd64 -- every operation rounds to float32, no fused
 * multiply-add (build with -ffp-contract=off), left-to-right evaluation.
 * All file:line citations are relative to /root/reference/.
 */
#ifndef FLUID_ORACLE_H
#define FLUID_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors `type Fluid` (pkg/fluid/fluid.go:11-40) plus package var
 * `Relaxation` (fluid.go:7-9), carried per instance. */
typedef struct fo_fluid {
    float density, h;
    int64_t NumX, NumY, numCells;
    float *U, *V, *newU, *newV, *p, *S, *M, *newM;
    float Confinement;
    float ViscosityDiffusion;
    float PressureDamping;
    float TurbulenceStrength;
    float SmokeAdvection;
    int UseMultigrid;
    int MultigridLevels;
    int UseBFECC;
    float Relaxation;
    /* bookkeeping, not in the reference */
    int threads;          /* GOMAXPROCS-equivalent for parallelRange */
    int last_iters;       /* sweeps executed by the last solveSingleGrid */
    float last_maxdiv;    /* maxDiv returned by the last executed sweep */
} fo_fluid;

enum { FO_U = 0, FO_V, FO_NEWU, FO_NEWV, FO_P, FO_S, FO_M, FO_NEWM, FO_NFIELDS };
enum { FO_FIELD_U = 0, FO_FIELD_V = 1, FO_FIELD_M = 2 };

fo_fluid *fo_new(float density, int64_t width, int64_t height, float h); /* fluid.go:42 */
void fo_free(fo_fluid *f);
float *fo_field(fo_fluid *f, int which);
void fo_set_threads(fo_fluid *f, int threads);

/* hot path */
void fo_simulate(fo_fluid *f, float dt);                       /* fluid.go:79 */
void fo_apply_viscosity(fo_fluid *f, float dt);                /* fluid.go:112 */
void fo_make_incompressible(fo_fluid *f, unsigned iters, float dt); /* fluid.go:144 */
float fo_pressure_iteration(fo_fluid *f, float relaxation, float cp); /* fluid.go:188 */
void fo_handle_borders(fo_fluid *f);                           /* fluid.go:236 */
void fo_advect_velocity(fo_fluid *f, float dt);                /* fluid.go:291 */
void fo_advect_smoke(fo_fluid *f, float dt);                   /* fluid.go:400 */
void fo_copy_border(fo_fluid *f, float *dst, const float *src);/* fluid.go:436 */
void fo_apply_vorticity_confinement(fo_fluid *f, float dt);    /* fluid.go:449 */
void fo_add_turbulence(fo_fluid *f, float dt);                 /* fluid.go:496 */
float fo_get_adaptive_time_step(fo_fluid *f, float basedt);    /* fluid.go:529 */
void fo_advect_velocity_bfecc(fo_fluid *f, float dt);          /* fluid.go:911 */
void fo_advect_smoke_bfecc(fo_fluid *f, float dt);             /* fluid.go:997 */
float fo_sample_field(const fo_fluid *f, float x, float y, int fld); /* fluid.go:357 */

/* edits (walls.go, fluid.go:761-796, 894-907); return -1 where Go panics */
int fo_set_solid(fo_fluid *f, int64_t i, int64_t j, int value);
int fo_is_solid(const fo_fluid *f, int64_t i, int64_t j);
int fo_set_velocity(fo_fluid *f, int64_t i, int64_t j, float u, float v);
int fo_add_smoke(fo_fluid *f, int64_t i, int64_t j, float smoke);
void fo_reset(fo_fluid *f);
void fo_apply_force(fo_fluid *f, int64_t i, int64_t j, float fx, float fy);
void fo_apply_force_radius(fo_fluid *f, int64_t cx, int64_t cy, float fx, float fy, int64_t radius);
void fo_set_circular_obstacle(fo_fluid *f, int64_t cx, int64_t cy, int64_t radius);

/* views (pressure.go, smoke.go, fluid.go:799-891) */
void fo_minmax(const float *a, int64_t n, float *mn, float *mx);
void fo_vorticity(const fo_fluid *f, float *vals, float *mn, float *mx);
void fo_velocity_magnitude(const fo_fluid *f, float *vals, float *mn, float *mx);
float fo_max_divergence(const fo_fluid *f);
void fo_sample_velocity(const fo_fluid *f, float x, float y, float *u, float *v);

/* ---- the frame loop either side of Simulate (main/main.go, main/colors.go): the UI's pixel pass and
 * its particle tracers, restated so that the device versions (fb_render, fb_advect_particles) can be
 * checked bit for bit.  SURVEY.md section 8(f) rank 3. */
typedef struct fo_particle {       /* main/main.go:139-144 */
    float x, y;
    uint8_t r, g, b, pad;
    float age, max_age;
} fo_particle;
/* advectParticles (main/main.go:512-546): ages, RK2 midpoint through SampleVelocity, bounds and
 * solid checks; survivors are compacted in place in their original order.  Returns how many. */
int64_t fo_advect_particles(const fo_fluid *f, fo_particle *ps, int64_t n, float dt);
/* Draw's pixel pass (main/main.go:550-574, 620-652; colors.go:8-84): view `kind` (0 smoke, 1 pressure,
 * 2 velocity magnitude, 3 vorticity) through getSciValue / getDivergingColor into an RGBA image of
 * NumX x NumY pixels (row jj of the image is fluid column NumY-1-jj: fluidToImageIndex, main.go:795),
 * solid cells painted (0,0,0,255). */
void fo_render(const fo_fluid *f, int kind, uint8_t *rgba);

/* Edit command lists with the layout of fb_edit_cmd (include/fluidb200.h): the
 * same preset description drives the oracle and the CUDA path.  Rectangles are
 * walked in lexicographic order calling the point edits above. */
typedef struct fo_edit_cmd {
    int32_t op;
    int32_t i0, j0, i1, j1;
    float a, b;
} fo_edit_cmd;
enum { FO_EDIT_SET_SOLID = 0, FO_EDIT_SET_VELOCITY = 1, FO_EDIT_ADD_SMOKE = 2, FO_EDIT_APPLY_FORCE = 3,
       FO_EDIT_CIRCLE_OBSTACLE = 4, FO_EDIT_RESET = 5, FO_EDIT_SET_VELOCITY_IF_FLUID = 6,
       FO_EDIT_ADD_SMOKE_IF_FLUID = 7, FO_EDIT_SET_SMOKE = 8 };
int fo_apply_edits(fo_fluid *f, const fo_edit_cmd *cmds, int64_t n);
/* Simulate x nsteps with `per_step` replayed before each step (main/main.go:233-245). */
int fo_run(fo_fluid *f, float dt, int64_t nsteps, const fo_edit_cmd *per_step, int64_t n);

/* NOT in the reference: CPU restatement of THIS repo's fast-mode projection
 * (red-black ordering of the same per-cell update as fluid.go:196-229), used
 * to check the CUDA fast mode bit for bit.  Returns max pre-update |div| of the
 * last iteration executed. */
float fo_project_redblack(fo_fluid *f, unsigned iters, float dt);
/* NOT in the reference: the same red-black iteration in PRESSURE FORM, restating the
 * arithmetic of this repo's fastest CUDA solver (rbq_fused.cuh) operation for
 * operation: per pass of <= 8 iterations, q accumulates the per-cell corrections
 * against the frozen initial divergence D0, and U, V, p are materialised once at
 * the end.  Algebraically identical to fo_project_redblack; rounding differs. */
float fo_project_redblack_q(fo_fluid *f, unsigned iters, float dt);
/* NOT in the reference: solveMultigridVCycle (fluid.go:560-599) with every sweep -- the six
 * fine-grid smoothing sweeps and the 40 coarse ones -- in red-black order; residual, restriction,
 * prolongation and correction as in the reference.  Checks the CUDA fast-mode V-cycle bit for bit. */
void fo_project_multigrid_redblack(fo_fluid *f, unsigned iters, float dt);
/* ... and with each block of three smoothing sweeps run as ONE pressure-form pass (fo_project_redblack_q's
 * arithmetic), which is what FB_SOLVER_REDBLACK_PRESSURE does. */
void fo_project_multigrid_redblack_q(fo_fluid *f, unsigned iters, float dt);
/* Same with an explicit omega per HALF sweep: omega[2k] red, omega[2k+1] black. */
float fo_project_redblack_sched(fo_fluid *f, const float *omega, unsigned iters, float dt);

#ifdef __cplusplus
}
#endif
#endif
