/*
 * fluid_oracle.c -- CPU restatement of TheFellow/fluid's pkg/fluid solver.
 *
 * TEST INFRASTRUCTURE ONLY (see fluid_oracle.h).  PARITY UNPINNED by reference
 * goldens (there are none); pinned by the reference tests' assertions only.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -pthread -shared -fPIC
 * (amd64-Go semantics: every op rounds to float32, no FMA contraction).
 * Every float literal below is a float32 constant, as Go's untyped constants
 * are when they meet a float32 operand.  Citations: /root/reference/<file:line>.
 */
#include "fluid_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- parallelRange (pkg/fluid/parallel.go:10-39) -------------------------
 * [start,end) split into `workers` contiguous chunks of ceil(total/workers);
 * one pooled pthread plays each goroutine (chunk 0 runs on the caller).  The
 * loops are race-free and order-independent, so results do not depend on the
 * thread count. */
typedef struct fo_ctx {
    fo_fluid *f;
    float dt;
    float *a, *b, *c, *d;
    int back;                /* 1 => BFECC backward (+dt) trace */
} fo_ctx;
typedef void (*fo_range_fn)(const fo_ctx *ctx, int64_t i);

#define FO_MAX_THREADS 256
static struct {
    pthread_mutex_t mu;
    pthread_cond_t go, done;
    pthread_t th[FO_MAX_THREADS];
    int nthreads;            /* workers created so far (ids 1..nthreads) */
    uint64_t gen;            /* job generation */
    int pending;             /* workers still running the current job */
    fo_range_fn fn;
    const fo_ctx *ctx;
    int64_t start, end, chunk;
    int workers;             /* chunks in the current job */
} g_pool = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER,
             {0}, 0, 0, 0, NULL, NULL, 0, 0, 0, 0 };

static void run_chunk(fo_range_fn fn, const fo_ctx *ctx, int64_t start, int64_t end, int64_t chunk, int w)
{
    int64_t s = start + (int64_t)w * chunk, e = s + chunk;
    if (e > end) e = end;
    for (int64_t i = s; i < e; i++) fn(ctx, i);
}

static void *pool_worker(void *arg)
{
    const int id = (int)(intptr_t)arg; /* 1-based */
    uint64_t seen = 0;
    pthread_mutex_lock(&g_pool.mu);
    for (;;) {
        while (g_pool.gen == seen) pthread_cond_wait(&g_pool.go, &g_pool.mu);
        seen = g_pool.gen;
        if (id < g_pool.workers) {
            fo_range_fn fn = g_pool.fn; const fo_ctx *ctx = g_pool.ctx;
            int64_t s = g_pool.start, e = g_pool.end, c = g_pool.chunk;
            pthread_mutex_unlock(&g_pool.mu);
            run_chunk(fn, ctx, s, e, c, id);
            pthread_mutex_lock(&g_pool.mu);
            if (--g_pool.pending == 0) pthread_cond_signal(&g_pool.done);
        }
    }
    return NULL;
}

static void parallel_range(const fo_ctx *ctx, int64_t start, int64_t end, fo_range_fn fn)
{
    int64_t total = end - start;
    if (total <= 0) return;
    int64_t workers = ctx->f->threads;
    if (workers > FO_MAX_THREADS) workers = FO_MAX_THREADS;
    if (workers > total) workers = total;
    if (workers < 1) workers = 1;
    int64_t chunk = (total + workers - 1) / workers;
    if (workers == 1) { run_chunk(fn, ctx, start, end, chunk, 0); return; }
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.nthreads < workers - 1) {
        int id = g_pool.nthreads + 1;
        if (pthread_create(&g_pool.th[id], NULL, pool_worker, (void *)(intptr_t)id) != 0) break;
        pthread_detach(g_pool.th[id]);
        g_pool.nthreads = id;
    }
    if (workers > g_pool.nthreads + 1) workers = g_pool.nthreads + 1, chunk = (total + workers - 1) / workers;
    g_pool.fn = fn; g_pool.ctx = ctx;
    g_pool.start = start; g_pool.end = end; g_pool.chunk = chunk;
    g_pool.workers = (int)workers;
    g_pool.pending = (int)workers - 1;
    g_pool.gen++;
    pthread_cond_broadcast(&g_pool.go);
    pthread_mutex_unlock(&g_pool.mu);
    run_chunk(fn, ctx, start, end, chunk, 0);
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.pending > 0) pthread_cond_wait(&g_pool.done, &g_pool.mu);
    pthread_mutex_unlock(&g_pool.mu);
}

static inline float go_minf(float a, float b) { return (a < b) ? a : ((b < a) ? b : (a != a ? a : b)); }
static inline float go_maxf(float a, float b) { return (a > b) ? a : ((b > a) ? b : (a != a ? a : b)); }
static inline int64_t min_i64(int64_t a, int64_t b) { return a < b ? a : b; }

/* ---- construction (fluid.go:42-68) -------------------------------------- */
fo_fluid *fo_new(float density, int64_t width, int64_t height, float h)
{
    fo_fluid *f = (fo_fluid *)calloc(1, sizeof(*f));
    if (!f) return NULL;
    f->density = density;
    f->h = h;
    f->NumX = width + 2;
    f->NumY = height + 2;
    f->numCells = f->NumX * f->NumY;
    float **arrs[] = { &f->U, &f->V, &f->newU, &f->newV, &f->p, &f->S, &f->M, &f->newM };
    for (int k = 0; k < 8; k++) {
        *arrs[k] = (float *)calloc((size_t)f->numCells, sizeof(float));
        if (!*arrs[k]) { fo_free(f); return NULL; }
    }
    f->Confinement = 0.0f;
    f->ViscosityDiffusion = 0.0f;
    f->PressureDamping = 1.0f;
    f->TurbulenceStrength = 0.02f;
    f->SmokeAdvection = 1.0f;
    f->UseMultigrid = 0;
    f->MultigridLevels = 2;
    f->UseBFECC = 0;
    f->Relaxation = 1.9f; /* fluid.go:8 */
    long nc = sysconf(_SC_NPROCESSORS_ONLN);
    f->threads = nc > 0 ? (int)nc : 1;
    return f;
}

void fo_free(fo_fluid *f)
{
    if (!f) return;
    free(f->U); free(f->V); free(f->newU); free(f->newV);
    free(f->p); free(f->S); free(f->M); free(f->newM);
    free(f);
}

float *fo_field(fo_fluid *f, int which)
{
    switch (which) {
    case FO_U: return f->U;       case FO_V: return f->V;
    case FO_NEWU: return f->newU; case FO_NEWV: return f->newV;
    case FO_P: return f->p;       case FO_S: return f->S;
    case FO_M: return f->M;       case FO_NEWM: return f->newM;
    default: return NULL;
    }
}

void fo_set_threads(fo_fluid *f, int threads) { f->threads = threads < 1 ? 1 : threads; }

/* ---- copyBorder (fluid.go:436-446) -------------------------------------- */
static void copy_border_i(const fo_ctx *c, int64_t i)
{
    const int64_t n = c->f->NumY, NY = c->f->NumY;
    float *dst = c->a; const float *src = c->b;
    dst[i * n + 0] = src[i * n + 0];
    dst[i * n + NY - 1] = src[i * n + NY - 1];
}
static void copy_border_j(const fo_ctx *c, int64_t j)
{
    const int64_t n = c->f->NumY, NX = c->f->NumX;
    float *dst = c->a; const float *src = c->b;
    dst[0 * n + j] = src[0 * n + j];
    dst[(NX - 1) * n + j] = src[(NX - 1) * n + j];
}
void fo_copy_border(fo_fluid *f, float *dst, const float *src)
{
    fo_ctx c = { f, 0.0f, dst, (float *)src, NULL, NULL, 0 };
    parallel_range(&c, 0, f->NumX, copy_border_i);
    parallel_range(&c, 0, f->NumY, copy_border_j);
}

/* ---- applyViscosity (fluid.go:112-142) ---------------------------------- */
static void viscosity_i(const fo_ctx *c, int64_t i)
{
    fo_fluid *f = c->f;
    const int64_t n = f->NumY, NY = f->NumY;
    const float visc = f->ViscosityDiffusion * c->dt;
    float *U = f->U, *V = f->V, *S = f->S, *nU = f->newU, *nV = f->newV;
    {
        for (int64_t j = 1; j < NY - 1; j++) {
            if (S[i * n + j] > 0.0f) {
                float c4 = 4.0f * U[i * n + j];
                float uLap = (((U[(i - 1) * n + j] + U[(i + 1) * n + j]) + U[i * n + j - 1]) + U[i * n + j + 1]) - c4;
                float t = visc * uLap;
                nU[i * n + j] = U[i * n + j] + t;
                float d4 = 4.0f * V[i * n + j];
                float vLap = (((V[(i - 1) * n + j] + V[(i + 1) * n + j]) + V[i * n + j - 1]) + V[i * n + j + 1]) - d4;
                float t2 = visc * vLap;
                nV[i * n + j] = V[i * n + j] + t2;
            }
        }
    }
}
void fo_apply_viscosity(fo_fluid *f, float dt)
{
    if (f->ViscosityDiffusion <= 0.0f) return;
    memcpy(f->newU, f->U, (size_t)f->numCells * sizeof(float));
    memcpy(f->newV, f->V, (size_t)f->numCells * sizeof(float));
    fo_ctx c = { f, dt, NULL, NULL, NULL, NULL, 0 };
    parallel_range(&c, 1, f->NumX - 1, viscosity_i);
    memcpy(f->U, f->newU, (size_t)f->numCells * sizeof(float));
    memcpy(f->V, f->newV, (size_t)f->numCells * sizeof(float));
}

/* ---- pressureJacobiIteration (fluid.go:188-234): one in-place lexicographic
 * Gauss-Seidel/SOR sweep on the face velocities (Q-1, Q-3, Q-4). */
float fo_pressure_iteration(fo_fluid *f, float relaxation, float cp)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float *U = f->U, *V = f->V, *S = f->S, *P = f->p;
    const float damping = f->PressureDamping;
    float maxDiv = 0.0f;
    for (int64_t i = 1; i < NX - 1; i++) {
        for (int64_t j = 1; j < NY - 1; j++) {
            if (S[i * n + j] == 0.0f) continue;
            float sx0 = S[(i - 1) * n + j];
            float sx1 = S[(i + 1) * n + j];
            float sy0 = S[i * n + j - 1];
            float sy1 = S[i * n + j + 1];
            float s = ((sx0 + sx1) + sy0) + sy1;
            if (s == 0.0f) continue;
            float div = ((U[(i + 1) * n + j] - U[i * n + j]) + V[i * n + j + 1]) - V[i * n + j];
            float absDiv = fabsf(div);
            if (absDiv > maxDiv) maxDiv = absDiv;
            float p = -div / s;
            p *= relaxation;
            p *= damping;
            float cpp = cp * p;
            P[i * n + j] += cpp;
            float a = sx0 * p; U[i * n + j] -= a;
            float b = sx1 * p; U[(i + 1) * n + j] += b;
            float c = sy0 * p; V[i * n + j] -= c;
            float d = sy1 * p; V[i * n + j + 1] += d;
        }
    }
    return maxDiv;
}

/* ---- solveSingleGrid (fluid.go:157-186) --------------------------------- */
static void solve_single_grid(fo_fluid *f, unsigned numIters, float dt)
{
    float cp = f->density * f->h / dt;
    const float tolerance = 1e-5f;
    float initialRelaxation = f->Relaxation;
    const float minRelaxation = 1.2f;
    f->last_iters = 0;
    f->last_maxdiv = 0.0f;
    for (unsigned iter = 0; iter < numIters; iter++) {
        float iterProgress = (float)iter / (float)numIters;
        float t = (initialRelaxation - minRelaxation) * iterProgress;
        float currentRelaxation = initialRelaxation - t;
        float maxDiv = fo_pressure_iteration(f, currentRelaxation, cp);
        f->last_iters = (int)iter + 1;
        f->last_maxdiv = maxDiv;
        if (maxDiv < tolerance) break;
        /* fluid.go:180-182 scales a local that is recomputed next iteration:
         * dead code (Q-2), intentionally not restated. */
    }
}

/* ---- multigrid V-cycle (fluid.go:560-758, 1123-1149); off by default ----- */
static float *compute_pressure_residual(fo_fluid *f)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float *U = f->U, *V = f->V, *S = f->S, *P = f->p;
    float *residual = (float *)calloc((size_t)f->numCells, sizeof(float));
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            if (S[i * n + j] == 0.0f) continue;
            float div = ((U[(i + 1) * n + j] - U[i * n + j]) + V[i * n + j + 1]) - V[i * n + j];
            float sx0 = S[(i - 1) * n + j], sx1 = S[(i + 1) * n + j];
            float sy0 = S[i * n + j - 1], sy1 = S[i * n + j + 1];
            float a = sx0 * (P[(i - 1) * n + j] - P[i * n + j]);
            float b = sx1 * (P[(i + 1) * n + j] - P[i * n + j]);
            float c = sy0 * (P[i * n + j - 1] - P[i * n + j]);
            float d = sy1 * (P[i * n + j + 1] - P[i * n + j]);
            float laplacian = ((a + b) + c) + d;
            residual[i * n + j] = -div - laplacian;
        }
    return residual;
}

static float *restrict_residual(fo_fluid *f, const float *fine)
{
    const int64_t NX = f->NumX, NY = f->NumY;
    const int64_t cNX = (NX + 1) / 2, cNY = (NY + 1) / 2;
    float *coarse = (float *)calloc((size_t)(cNX * cNY), sizeof(float));
    for (int64_t i = 1; i < cNX - 1; i++)
        for (int64_t j = 1; j < cNY - 1; j++) {
            int64_t fi = i * 2, fj = j * 2;
            if (fi < NX - 1 && fj < NY - 1) {
                float center = fine[fi * NY + fj] * 0.25f;
                float nb = (((fine[(fi - 1) * NY + fj] + fine[(fi + 1) * NY + fj]) + fine[fi * NY + fj - 1]) + fine[fi * NY + fj + 1]) * 0.125f;
                float cr = (((fine[(fi - 1) * NY + fj - 1] + fine[(fi + 1) * NY + fj - 1]) + fine[(fi - 1) * NY + fj + 1]) + fine[(fi + 1) * NY + fj + 1]) * 0.0625f;
                coarse[i * cNY + j] = (center + nb) + cr;
            }
        }
    return coarse;
}

/* redblack = 0: the reference's lexicographic sweeps.  redblack = 1 (NOT in the reference): the
 * same cell update, each sweep split into the (i+j) even cells then the odd ones -- the order of
 * this repo's CUDA fast mode. */
static float *solve_coarse_grid(fo_fluid *f, const float *rhs, int redblack)
{
    const int64_t NX = f->NumX, NY = f->NumY;
    const int64_t cNX = (NX + 1) / 2, cNY = (NY + 1) / 2;
    float *cP = (float *)calloc((size_t)(cNX * cNY), sizeof(float));
    float *cS = (float *)calloc((size_t)(cNX * cNY), sizeof(float));
    for (int64_t i = 0; i < cNX; i++)
        for (int64_t j = 0; j < cNY; j++) {
            int64_t fi = i * 2, fj = j * 2;
            cS[i * cNY + j] = (fi < NX && fj < NY) ? f->S[fi * NY + fj] : 0.0f;
        }
    const float relaxation = 1.6f;
    for (int iter = 0; iter < 40; iter++)
      for (int colour = 0; colour < (redblack ? 2 : 1); colour++)
        for (int64_t i = 1; i < cNX - 1; i++)
            for (int64_t j = 1; j < cNY - 1; j++) {
                if (redblack && ((i + j) & 1) != colour) continue;
                if (cS[i * cNY + j] == 0.0f) continue;
                float sx0 = cS[(i - 1) * cNY + j], sx1 = cS[(i + 1) * cNY + j];
                float sy0 = cS[i * cNY + j - 1], sy1 = cS[i * cNY + j + 1];
                float s = ((sx0 + sx1) + sy0) + sy1;
                if (s == 0.0f) continue;
                float a = sx0 * (cP[(i - 1) * cNY + j] - cP[i * cNY + j]);
                float b = sx1 * (cP[(i + 1) * cNY + j] - cP[i * cNY + j]);
                float c = sy0 * (cP[i * cNY + j - 1] - cP[i * cNY + j]);
                float d = sy1 * (cP[i * cNY + j + 1] - cP[i * cNY + j]);
                float laplacian = ((a + b) + c) + d;
                float residual = rhs[i * cNY + j] - laplacian;
                float correction = (-residual / s) * relaxation;
                cP[i * cNY + j] += correction;
            }
    free(cS);
    return cP;
}

static float *prolongate_correction(fo_fluid *f, const float *cc)
{
    const int64_t NX = f->NumX, NY = f->NumY;
    const int64_t cNX = (NX + 1) / 2, cNY = (NY + 1) / 2;
    float *corr = (float *)calloc((size_t)f->numCells, sizeof(float));
    for (int64_t i = 0; i < NX; i++)
        for (int64_t j = 0; j < NY; j++) {
            int64_t ci = i / 2, cj = j / 2;
            if (ci >= cNX - 1 || cj >= cNY - 1) continue;
            float fracI = (float)(i % 2) * 0.5f;
            float fracJ = (float)(j % 2) * 0.5f;
            float w00 = (1.0f - fracI) * (1.0f - fracJ);
            float w10 = fracI * (1.0f - fracJ);
            float w01 = (1.0f - fracI) * fracJ;
            float w11 = fracI * fracJ;
            float t0 = w00 * cc[ci * cNY + cj];
            float t1 = w10 * cc[(ci + 1) * cNY + cj];
            float t2 = w01 * cc[ci * cNY + cj + 1];
            float t3 = w11 * cc[(ci + 1) * cNY + cj + 1];
            corr[i * NY + j] = ((t0 + t1) + t2) + t3;
        }
    return corr;
}

static void apply_pressure_correction(fo_fluid *f, const float *correction, float cp)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float *U = f->U, *V = f->V, *S = f->S, *P = f->p;
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            if (S[i * n + j] == 0.0f) continue;
            float t = correction[i * n + j] * cp;
            P[i * n + j] += t;
            float sx0 = S[(i - 1) * n + j], sx1 = S[(i + 1) * n + j];
            float sy0 = S[i * n + j - 1], sy1 = S[i * n + j + 1];
            float pc = correction[i * n + j];
            float a = sx0 * pc; U[i * n + j] -= a;
            float b = sx1 * pc; U[(i + 1) * n + j] += b;
            float c = sy0 * pc; V[i * n + j] -= c;
            float d = sy1 * pc; V[i * n + j + 1] += d;
        }
}

static float redblack_half(fo_fluid *f, float relaxation, float cp, int colour);
static void redblack_q_pass(fo_fluid *f, const float *omega, unsigned iters, float cp);

/* Three fine-grid smoothing sweeps at one relaxation; returns max |div| of the third.
 * mode 0: pressureJacobiIteration (fluid.go:188-234), the reference.  NOT in the reference: mode 1, each
 * sweep as a red and a black half sweep; mode 2, the same three red-black sweeps in pressure form as one
 * pass (the arithmetic of rbq_fused.cuh, see redblack_q_pass). */
static float mg_smooth3(fo_fluid *f, float relaxation, float cp, int mode)
{
    float maxDiv = 0.0f;
    if (mode == 2) {
        const float om[6] = { relaxation, relaxation, relaxation, relaxation, relaxation, relaxation };
        redblack_q_pass(f, om, 3, cp);
        return f->last_maxdiv;
    }
    for (unsigned s = 0; s < 3; s++) {
        if (mode == 0) maxDiv = fo_pressure_iteration(f, relaxation, cp);
        else {
            float a = redblack_half(f, relaxation, cp, 0);
            float b = redblack_half(f, relaxation, cp, 1);
            maxDiv = a > b ? a : b;
        }
    }
    return maxDiv;
}

static void solve_multigrid_vcycle(fo_fluid *f, unsigned numIters, float dt, int mode)
{
    const int redblack = mode != 0;
    float cp = f->density * f->h / dt;
    const float tolerance = 1e-5f;
    f->last_iters = 0;
    f->last_maxdiv = 0.0f;
    for (unsigned iter = 0; iter < numIters; iter++) {
        float maxDiv = mg_smooth3(f, 1.5f, cp, mode);
        f->last_iters = (int)iter + 1;      /* cycles entered */
        f->last_maxdiv = maxDiv;            /* the value fluid.go:575 tests */
        if (maxDiv < tolerance) break;
        float *residual = compute_pressure_residual(f);
        float *coarseRHS = restrict_residual(f, residual);
        float *coarseCorr = solve_coarse_grid(f, coarseRHS, redblack);
        float *corr = prolongate_correction(f, coarseCorr);
        apply_pressure_correction(f, corr, cp);
        mg_smooth3(f, 1.2f, cp, mode);
        f->last_iters = (int)iter + 1;
        f->last_maxdiv = maxDiv;
        free(residual); free(coarseRHS); free(coarseCorr); free(corr);
    }
}

/* NOT in the reference: the V-cycle with red-black sweeps on both levels (see fluid_oracle.h). */
void fo_project_multigrid_redblack(fo_fluid *f, unsigned iters, float dt)
{
    fo_copy_border(f, f->newU, f->U);
    fo_copy_border(f, f->newV, f->V);
    solve_multigrid_vcycle(f, iters, dt, 1);
}

/* NOT in the reference: the same with the smoothing sweeps in pressure form (FB_SOLVER_REDBLACK_PRESSURE). */
void fo_project_multigrid_redblack_q(fo_fluid *f, unsigned iters, float dt)
{
    fo_copy_border(f, f->newU, f->U);
    fo_copy_border(f, f->newV, f->V);
    solve_multigrid_vcycle(f, iters, dt, 2);
}

/* ---- makeIncompressible (fluid.go:144-155) ------------------------------ */
void fo_make_incompressible(fo_fluid *f, unsigned iters, float dt)
{
    fo_copy_border(f, f->newU, f->U);
    fo_copy_border(f, f->newV, f->V);
    if (f->UseMultigrid && f->MultigridLevels > 1) solve_multigrid_vcycle(f, iters, dt, 0);
    else solve_single_grid(f, iters, dt);
}

/* ---- handleBorders (fluid.go:236-289), Q-13 ----------------------------- */
static void borders_i(const fo_ctx *c, int64_t i)
{
    fo_fluid *f = c->f;
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float *U = f->U, *S = f->S;
    {
        if (S[i * n + 0] == 0.0f || S[i * n + 1] == 0.0f) {
            U[i * n + 0] = 0.0f;
        } else if (i > 0 && i < NX - 1 && S[i * n + 2] > 0.0f) {
            float t = 2.0f * U[i * n + 1];
            U[i * n + 0] = t - U[i * n + 2];
        } else {
            U[i * n + 0] = U[i * n + 1];
        }
        if (S[i * n + NY - 1] == 0.0f || S[i * n + NY - 2] == 0.0f) {
            U[i * n + NY - 1] = 0.0f;
        } else if (i > 0 && i < NX - 1 && S[i * n + NY - 3] > 0.0f) {
            float t = 2.0f * U[i * n + NY - 2];
            U[i * n + NY - 1] = t - U[i * n + NY - 3];
        } else {
            U[i * n + NY - 1] = U[i * n + NY - 2];
        }
    }
}
static void borders_j(const fo_ctx *c, int64_t j)
{
    fo_fluid *f = c->f;
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float *V = f->V, *S = f->S;
    {
        if (S[0 * n + j] == 0.0f || S[1 * n + j] == 0.0f) {
            V[0 * n + j] = 0.0f;
        } else if (j > 0 && j < NY - 1 && S[2 * n + j] > 0.0f) {
            float t = 2.0f * V[1 * n + j];
            V[0 * n + j] = t - V[2 * n + j];
        } else {
            V[0 * n + j] = V[1 * n + j];
        }
        if (S[(NX - 1) * n + j] == 0.0f || S[(NX - 2) * n + j] == 0.0f) {
            V[(NX - 1) * n + j] = 0.0f;
        } else if (j > 0 && j < NY - 1 && S[(NX - 3) * n + j] > 0.0f) {
            float t = 2.0f * V[(NX - 2) * n + j];
            V[(NX - 1) * n + j] = t - V[(NX - 3) * n + j];
        } else {
            V[(NX - 1) * n + j] = V[(NX - 2) * n + j];
        }
    }
}
void fo_handle_borders(fo_fluid *f)
{
    fo_ctx c = { f, 0.0f, NULL, NULL, NULL, NULL, 0 };
    parallel_range(&c, 0, f->NumX, borders_i);
    parallel_range(&c, 0, f->NumY, borders_j);
}

/* ---- sampleField / sampleFieldFrom (fluid.go:357-398, 1055-1091), Q-8 ---- */
static inline float sample_from(const fo_fluid *f, float x, float y, const float *data, int fld)
{
    const int64_t n = f->NumY;
    const float h = f->h;
    const float h1 = 1.0f / h;
    const float h2 = 0.5f * h;
    x = go_maxf(go_minf(x, (float)f->NumX * h), h);
    y = go_maxf(go_minf(y, (float)f->NumY * h), h);
    float dx = 0.0f, dy = 0.0f;
    switch (fld) {
    case FO_FIELD_U: dy = h2; break;
    case FO_FIELD_V: dx = h2; break;
    default: dx = h2; dy = h2; break;
    }
    float xs = x - dx;
    float xh = xs * h1;
    int64_t x0 = min_i64((int64_t)floor((double)xh), f->NumX - 1);
    float x0h = (float)x0 * h;
    float tx = (xs - x0h) * h1;
    int64_t x1 = min_i64(x0 + 1, f->NumX - 1);

    float ys = y - dy;
    float yh = ys * h1;
    int64_t y0 = min_i64((int64_t)floor((double)yh), f->NumY - 1);
    float y0h = (float)y0 * h;
    float ty = (ys - y0h) * h1;
    int64_t y1 = min_i64(y0 + 1, f->NumY - 1);

    float sx = 1.0f - tx;
    float sy = 1.0f - ty;

    float w00 = sx * sy, w10 = tx * sy, w11 = tx * ty, w01 = sx * ty;
    float a = w00 * data[x0 * n + y0];
    float b = w10 * data[x1 * n + y0];
    float c = w11 * data[x1 * n + y1];
    float d = w01 * data[x0 * n + y1];
    return ((a + b) + c) + d;
}

float fo_sample_field(const fo_fluid *f, float x, float y, int fld)
{
    const float *data = fld == FO_FIELD_U ? f->U : (fld == FO_FIELD_V ? f->V : f->M);
    return sample_from(f, x, y, data, fld);
}

/* avgU / avgV (fluid.go:335-347) */
static inline float avg_u(const fo_fluid *f, int64_t i, int64_t j)
{
    const int64_t n = f->NumY; const float *U = f->U;
    return (((U[i * n + j - 1] + U[i * n + j]) + U[(i + 1) * n + j - 1]) + U[(i + 1) * n + j]) * 0.25f;
}
static inline float avg_v(const fo_fluid *f, int64_t i, int64_t j)
{
    const int64_t n = f->NumY; const float *V = f->V;
    return (((V[(i - 1) * n + j] + V[i * n + j]) + V[(i - 1) * n + j + 1]) + V[i * n + j + 1]) * 0.25f;
}

/* Shared body of the two velocity back-traces: advectVelocity (fluid.go:300-329,
 * sign -1, samples f->U/f->V into newU/newV) and the BFECC backward pass
 * (fluid.go:943-966, sign +1, samples fwdU/fwdV into bwdU/bwdV).  Trace
 * velocities always come from f->U / f->V.  ctx: a=srcU b=srcV c=dstU d=dstV. */
static void trace_velocity_i(const fo_ctx *c, int64_t i)
{
    fo_fluid *f = c->f;
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    const float h = f->h;
    const float h2 = h / 2.0f;
    const float *U = f->U, *V = f->V, *S = f->S;
    const float *srcU = c->a, *srcV = c->b;
    float *dstU = c->c, *dstV = c->d;
    const int back = c->back;
    const float dt = c->dt;
    for (int64_t j = 1; j < NY; j++) {
        if (S[i * n + j] != 0.0f && S[(i - 1) * n + j] != 0.0f && j < NY - 1) {
            float x = (float)i * h;
            float y = (float)j * h + h2;
            float u = U[i * n + j];
            float v = avg_v(f, i, j);
            float du = dt * u, dv = dt * v;
            if (back) { x = x + du; y = y + dv; } else { x = x - du; y = y - dv; }
            dstU[i * n + j] = sample_from(f, x, y, srcU, FO_FIELD_U);
        }
        if (S[i * n + j] != 0.0f && S[i * n + j - 1] != 0.0f && i < NX - 1) {
            float x = (float)i * h + h2;
            float y = (float)j * h;
            float u = avg_u(f, i, j);
            float v = V[i * n + j];
            float du = dt * u, dv = dt * v;
            if (back) { x = x + du; y = y + dv; } else { x = x - du; y = y - dv; }
            dstV[i * n + j] = sample_from(f, x, y, srcV, FO_FIELD_V);
        }
    }
}

/* ---- advectVelocity (fluid.go:291-333), Q-6, Q-7, Q-9 -------------------- */
void fo_advect_velocity(fo_fluid *f, float dt)
{
    fo_copy_border(f, f->newU, f->U);
    fo_copy_border(f, f->newV, f->V);
    fo_ctx c = { f, dt, f->U, f->V, f->newU, f->newV, 0 };
    parallel_range(&c, 1, f->NumX, trace_velocity_i);
    memcpy(f->U, f->newU, (size_t)f->numCells * sizeof(float));
    memcpy(f->V, f->newV, (size_t)f->numCells * sizeof(float));
}

/* Shared body of the two smoke back-traces: advectSmoke (fluid.go:408-431,
 * forward, with clamp and optional diffusion) and the BFECC backward pass
 * (fluid.go:1017-1027).  ctx: a=src b=dst. */
static void trace_smoke_i(const fo_ctx *c, int64_t i)
{
    fo_fluid *f = c->f;
    const int64_t n = f->NumY, NY = f->NumY;
    const float h = f->h;
    const float h2 = 0.5f * h;
    const float *U = f->U, *V = f->V, *S = f->S, *M = f->M;
    const float *src = c->a;
    float *dst = c->b;
    const float sa = f->SmokeAdvection;
    const float vd = f->ViscosityDiffusion;
    const int back = c->back;
    const float dt = c->dt;
    for (int64_t j = 1; j < NY - 1; j++) {
        if (S[i * n + j] != 0.0f) {
            float u = ((U[i * n + j] + U[(i + 1) * n + j]) * 0.5f) * sa;
            float v = ((V[i * n + j] + V[i * n + j + 1]) * 0.5f) * sa;
            float du = dt * u;
            float dv = dt * v;
            float x0 = (float)i * h + h2;
            float y0 = (float)j * h + h2;
            if (back) {
                dst[i * n + j] = sample_from(f, x0 + du, y0 + dv, src, FO_FIELD_M);
                continue;
            }
            float smokeValue = sample_from(f, x0 - du, y0 - dv, src, FO_FIELD_M);
            if (vd > 0.0f) {
                float smokeDiffusion = (vd * 0.3f) * dt;
                float c4 = 4.0f * M[i * n + j];
                float neighbors = (((M[(i - 1) * n + j] + M[(i + 1) * n + j]) + M[i * n + j - 1]) + M[i * n + j + 1]) - c4;
                float t = smokeDiffusion * neighbors;
                smokeValue += t;
            }
            dst[i * n + j] = go_maxf(smokeValue, 0.0f);
        }
    }
}

/* ---- advectSmoke (fluid.go:400-434), Q-9 --------------------------------- */
void fo_advect_smoke(fo_fluid *f, float dt)
{
    fo_copy_border(f, f->newM, f->M);
    fo_ctx c = { f, dt, f->M, f->newM, NULL, NULL, 0 };
    parallel_range(&c, 1, f->NumX - 1, trace_smoke_i);
    memcpy(f->M, f->newM, (size_t)f->numCells * sizeof(float));
}

/* ---- applyVorticityConfinement (fluid.go:449-493), Q-12 ------------------ */
void fo_apply_vorticity_confinement(fo_fluid *f, float dt)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    const float h = f->h;
    float *U = f->U, *V = f->V, *S = f->S;
    float *curl = (float *)calloc((size_t)f->numCells, sizeof(float));
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            if (S[i * n + j] == 0.0f) continue;
            float dvdx = ((V[(i + 1) * n + j] - V[(i - 1) * n + j]) * 0.5f) / h;
            float dudy = ((U[i * n + j + 1] - U[i * n + j - 1]) * 0.5f) / h;
            curl[i * n + j] = dvdx - dudy;
        }
    const float eps = 1e-5f;
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            if (S[i * n + j] == 0.0f) continue;
            float gx = ((fabsf(curl[(i + 1) * n + j]) - fabsf(curl[(i - 1) * n + j])) * 0.5f) / h;
            float gy = ((fabsf(curl[i * n + j + 1]) - fabsf(curl[i * n + j - 1])) * 0.5f) / h;
            float gx2 = gx * gx, gy2 = gy * gy;
            float mag = sqrtf(gx2 + gy2) + eps;
            gx /= mag;
            gy /= mag;
            float vort = curl[i * n + j];
            float uu = U[i * n + j] * U[i * n + j];
            float vv = V[i * n + j] * V[i * n + j];
            float localVel = sqrtf(uu + vv);
            float lv = localVel * 0.1f;
            float adaptiveStrength = f->Confinement * (1.0f + lv);
            float fu = ((adaptiveStrength * gy) * vort) * dt;
            float fv = ((adaptiveStrength * gx) * vort) * dt;
            U[i * n + j] += fu;
            V[i * n + j] -= fv;
        }
    free(curl);
}

/* ---- addTurbulence (fluid.go:496-526), Q-5 ------------------------------- */
static void turbulence_i(const fo_ctx *c, int64_t i)
{
    fo_fluid *f = c->f;
    const int64_t n = f->NumY, NY = f->NumY;
    const float turbStrength = f->TurbulenceStrength * c->dt;
    float *U = f->U, *V = f->V, *S = f->S;
    for (int64_t j = 1; j < NY - 1; j++) {
        if (S[i * n + j] > 0.0f) {
            float seedU = (float)(i * 137 + j * 241) * 0.01f;
            float seedV = (float)(i * 157 + j * 263) * 0.01f;
            float noiseU = (float)sin((double)seedU) * turbStrength;
            float noiseV = (float)sin((double)seedV) * turbStrength;
            float uu = U[i * n + j] * U[i * n + j];
            float vv = V[i * n + j] * V[i * n + j];
            float localVel = sqrtf(uu + vv);
            if (localVel > 0.1f) {
                float turbulenceFactor = go_minf(localVel * 0.5f, 1.0f);
                float a = noiseU * turbulenceFactor;
                float b = noiseV * turbulenceFactor;
                U[i * n + j] += a;
                V[i * n + j] += b;
            }
        }
    }
}
void fo_add_turbulence(fo_fluid *f, float dt)
{
    if (f->TurbulenceStrength <= 0.0f) return;
    fo_ctx c = { f, dt, NULL, NULL, NULL, NULL, 0 };
    parallel_range(&c, 1, f->NumX - 1, turbulence_i);
}

/* ---- GetAdaptiveTimeStep (fluid.go:529-557) ------------------------------ */
float fo_get_adaptive_time_step(fo_fluid *f, float basedt)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float maxVel = 0.0f;
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++)
            if (f->S[i * n + j] > 0.0f) {
                float vel = fabsf(f->U[i * n + j]) + fabsf(f->V[i * n + j]);
                if (vel > maxVel) maxVel = vel;
            }
    if (maxVel == 0.0f) return basedt;
    const float cflFactor = 0.8f;
    float adaptivedt = (cflFactor * f->h) / maxVel;
    adaptivedt = go_maxf(go_minf(adaptivedt, basedt * 2.0f), basedt * 0.1f);
    return adaptivedt;
}

/* ---- clampToNeighbors (fluid.go:1094-1120) ------------------------------- */
static inline float clamp_to_neighbors(const fo_fluid *f, float val, const float *src, int64_t i, int64_t j)
{
    const int64_t n = f->NumY;
    float lo = src[i * n + j], hi = src[i * n + j];
    for (int64_t di = -1; di <= 1; di++)
        for (int64_t dj = -1; dj <= 1; dj++) {
            int64_t ni = i + di, nj = j + dj;
            if (ni >= 0 && ni < f->NumX && nj >= 0 && nj < f->NumY) {
                float v = src[ni * n + nj];
                if (v < lo) lo = v;
                if (v > hi) hi = v;
            }
        }
    if (val < lo) return lo;
    if (val > hi) return hi;
    return val;
}

static float *dup_field(const fo_fluid *f, const float *src)
{
    float *d = (float *)malloc((size_t)f->numCells * sizeof(float));
    memcpy(d, src, (size_t)f->numCells * sizeof(float));
    return d;
}

/* ---- advectVelocityBFECC (fluid.go:911-994), Q-10 ------------------------ */
void fo_advect_velocity_bfecc(fo_fluid *f, float dt)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    const size_t bytes = (size_t)f->numCells * sizeof(float);
    float *origU = dup_field(f, f->U), *origV = dup_field(f, f->V);
    fo_advect_velocity(f, dt);
    float *fwdU = dup_field(f, f->U), *fwdV = dup_field(f, f->V);
    memcpy(f->U, origU, bytes);
    memcpy(f->V, origV, bytes);
    float *bwdU = (float *)calloc((size_t)f->numCells, sizeof(float));
    float *bwdV = (float *)calloc((size_t)f->numCells, sizeof(float));
    fo_copy_border(f, bwdU, fwdU);
    fo_copy_border(f, bwdV, fwdV);
    {
        fo_ctx c = { f, dt, fwdU, fwdV, bwdU, bwdV, 1 };
        parallel_range(&c, 1, NX, trace_velocity_i);
    }
    float *corrU = dup_field(f, origU), *corrV = dup_field(f, origV);
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            float eu = (bwdU[i * n + j] - origU[i * n + j]) * 0.5f;
            corrU[i * n + j] = origU[i * n + j] - eu;
            float ev = (bwdV[i * n + j] - origV[i * n + j]) * 0.5f;
            corrV[i * n + j] = origV[i * n + j] - ev;
        }
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            corrU[i * n + j] = clamp_to_neighbors(f, corrU[i * n + j], origU, i, j);
            corrV[i * n + j] = clamp_to_neighbors(f, corrV[i * n + j], origV, i, j);
        }
    memcpy(f->U, corrU, bytes);
    memcpy(f->V, corrV, bytes);
    fo_advect_velocity(f, dt);
    free(origU); free(origV); free(fwdU); free(fwdV);
    free(bwdU); free(bwdV); free(corrU); free(corrV);
}

/* ---- advectSmokeBFECC (fluid.go:997-1051), Q-11 -------------------------- */
void fo_advect_smoke_bfecc(fo_fluid *f, float dt)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    const size_t bytes = (size_t)f->numCells * sizeof(float);
    float *origM = dup_field(f, f->M);
    fo_advect_smoke(f, dt);
    float *fwdM = dup_field(f, f->M);
    float *bwdM = (float *)calloc((size_t)f->numCells, sizeof(float));
    fo_copy_border(f, bwdM, fwdM);
    {
        fo_ctx c = { f, dt, fwdM, bwdM, NULL, NULL, 1 };
        parallel_range(&c, 1, NX - 1, trace_smoke_i);
    }
    float *corrM = dup_field(f, origM);
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            float e = (bwdM[i * n + j] - origM[i * n + j]) * 0.5f;
            corrM[i * n + j] = origM[i * n + j] - e;
        }
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            corrM[i * n + j] = clamp_to_neighbors(f, corrM[i * n + j], origM, i, j);
            if (corrM[i * n + j] < 0.0f) corrM[i * n + j] = 0.0f;
        }
    memcpy(f->M, corrM, bytes);
    fo_advect_smoke(f, dt);
    free(origM); free(fwdM); free(bwdM); free(corrM);
}

/* ---- Simulate (fluid.go:79-109) ------------------------------------------ */
void fo_simulate(fo_fluid *f, float dt)
{
    const unsigned numIters = 8;
    memset(f->p, 0, (size_t)f->numCells * sizeof(float));
    if (f->ViscosityDiffusion > 0.0f) fo_apply_viscosity(f, dt);
    fo_make_incompressible(f, numIters, dt);
    if (f->Confinement != 0.0f) fo_apply_vorticity_confinement(f, dt);
    if (f->TurbulenceStrength > 0.0f) fo_add_turbulence(f, dt);
    fo_handle_borders(f);
    if (f->UseBFECC) {
        fo_advect_velocity_bfecc(f, dt);
        fo_advect_smoke_bfecc(f, dt);
    } else {
        fo_advect_velocity(f, dt);
        fo_advect_smoke(f, dt);
    }
}

/* ---- edits (walls.go:5-93) ------------------------------------------------ */
int fo_set_solid(fo_fluid *f, int64_t i, int64_t j, int value)
{
    if (i < 0 || i >= f->NumX || j < 0 || j >= f->NumY) return -1; /* Go panics */
    const int64_t n = f->NumY;
    f->S[i * n + j] = value ? 0.0f : 1.0f;
    if (value) {
        f->U[i * n + j] = 0.0f;
        if (i + 1 < f->NumX) f->U[(i + 1) * n + j] = 0.0f;
        f->V[i * n + j] = 0.0f;
        if (j + 1 < f->NumY) f->V[i * n + j + 1] = 0.0f;
        f->newU[i * n + j] = 0.0f;
        if (i + 1 < f->NumX) f->newU[(i + 1) * n + j] = 0.0f;
        f->newV[i * n + j] = 0.0f;
        if (j + 1 < f->NumY) f->newV[i * n + j + 1] = 0.0f;
    }
    return 0;
}

int fo_is_solid(const fo_fluid *f, int64_t i, int64_t j)
{
    if (i < 0 || i >= f->NumX || j < 0 || j >= f->NumY) return -1;
    return f->S[i * f->NumY + j] == 0.0f;
}

int fo_set_velocity(fo_fluid *f, int64_t i, int64_t j, float u, float v)
{
    if (i < 0 || i >= f->NumX || j < 0 || j >= f->NumY) return -1;
    f->U[i * f->NumY + j] = u;
    f->V[i * f->NumY + j] = v;
    return 0;
}

int fo_add_smoke(fo_fluid *f, int64_t i, int64_t j, float smoke)
{
    if (i < 0 || i >= f->NumX || j < 0 || j >= f->NumY) return -1;
    f->M[i * f->NumY + j] += smoke;
    return 0;
}

void fo_reset(fo_fluid *f) /* walls.go:85-93: S is NOT reset */
{
    const size_t bytes = (size_t)f->numCells * sizeof(float);
    memset(f->U, 0, bytes); memset(f->V, 0, bytes);
    memset(f->newU, 0, bytes); memset(f->newV, 0, bytes);
    memset(f->p, 0, bytes); memset(f->M, 0, bytes); memset(f->newM, 0, bytes);
}

void fo_apply_force(fo_fluid *f, int64_t i, int64_t j, float fx, float fy) /* fluid.go:761 */
{
    if (i < 1 || i >= f->NumX - 1 || j < 1 || j >= f->NumY - 1) return;
    const int64_t n = f->NumY;
    if (f->S[i * n + j] == 0.0f) return;
    f->U[i * n + j] += fx;
    f->V[i * n + j] += fy;
}

void fo_apply_force_radius(fo_fluid *f, int64_t cx, int64_t cy, float fx, float fy, int64_t radius) /* fluid.go:774 */
{
    if (radius <= 0) { fo_apply_force(f, cx, cy, fx, fy); return; }
    float r2 = (float)(radius * radius);
    for (int64_t i = cx - radius; i <= cx + radius; i++)
        for (int64_t j = cy - radius; j <= cy + radius; j++) {
            if (i < 1 || i >= f->NumX - 1 || j < 1 || j >= f->NumY - 1) continue;
            float dx = (float)(i - cx), dy = (float)(j - cy);
            float dx2 = dx * dx, dy2 = dy * dy;
            float dist2 = dx2 + dy2;
            if (dist2 > r2) continue;
            float arg = (-3.0f * dist2) / r2;
            float weight = (float)exp((double)arg);
            fo_apply_force(f, i, j, fx * weight, fy * weight);
        }
}

void fo_set_circular_obstacle(fo_fluid *f, int64_t cx, int64_t cy, int64_t radius) /* fluid.go:894 */
{
    for (int64_t i = cx - radius; i <= cx + radius; i++)
        for (int64_t j = cy - radius; j <= cy + radius; j++) {
            if (i < 0 || i >= f->NumX || j < 0 || j >= f->NumY) continue;
            float dx = (float)(i - cx), dy = (float)(j - cy);
            float dx2 = dx * dx, dy2 = dy * dy;
            if (dx2 + dy2 <= (float)(radius * radius)) fo_set_solid(f, i, j, 1);
        }
}

/* ---- views (pressure.go:5-24, smoke.go:5-24, fluid.go:799-891), Q-14 ------ */
void fo_minmax(const float *a, int64_t n, float *mn, float *mx)
{
    float lo = FLT_MAX;
    float hi = -FLT_MAX; /* float32(-(MaxFloat32-1)) rounds to -MaxFloat32 */
    for (int64_t k = 0; k < n; k++) {
        if (a[k] < lo) lo = a[k];
        if (a[k] > hi) hi = a[k];
    }
    *mn = lo; *mx = hi;
}

void fo_vorticity(const fo_fluid *f, float *vals, float *mn, float *mx)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    const float h = f->h;
    float lo = FLT_MAX, hi = -FLT_MAX;
    memset(vals, 0, (size_t)f->numCells * sizeof(float));
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            if (f->S[i * n + j] == 0.0f) continue;
            float dvdx = ((f->V[(i + 1) * n + j] - f->V[(i - 1) * n + j]) * 0.5f) / h;
            float dudy = ((f->U[i * n + j + 1] - f->U[i * n + j - 1]) * 0.5f) / h;
            float curl = dvdx - dudy;
            vals[i * n + j] = curl;
            if (curl < lo) lo = curl;
            if (curl > hi) hi = curl;
        }
    *mn = lo; *mx = hi;
}

void fo_velocity_magnitude(const fo_fluid *f, float *vals, float *mn, float *mx)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float lo = FLT_MAX, hi = -FLT_MAX;
    memset(vals, 0, (size_t)f->numCells * sizeof(float));
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            if (f->S[i * n + j] == 0.0f) continue;
            float u = (f->U[i * n + j] + f->U[(i + 1) * n + j]) * 0.5f;
            float v = (f->V[i * n + j] + f->V[i * n + j + 1]) * 0.5f;
            float uu = u * u, vv = v * v;
            float mag = sqrtf(uu + vv);
            vals[i * n + j] = mag;
            if (mag < lo) lo = mag;
            if (mag > hi) hi = mag;
        }
    *mn = lo; *mx = hi;
}

float fo_max_divergence(const fo_fluid *f)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float maxDiv = 0.0f;
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            if (f->S[i * n + j] == 0.0f) continue;
            float div = ((f->U[(i + 1) * n + j] - f->U[i * n + j]) + f->V[i * n + j + 1]) - f->V[i * n + j];
            float a = fabsf(div);
            if (a > maxDiv) maxDiv = a;
        }
    return maxDiv;
}

void fo_sample_velocity(const fo_fluid *f, float x, float y, float *u, float *v)
{
    *u = sample_from(f, x, y, f->U, FO_FIELD_U);
    *v = sample_from(f, x, y, f->V, FO_FIELD_V);
}

/* ---- the UI's pixel pass and particle tracers (main/main.go, main/colors.go) ---------------
 * Go converts float32 -> uint8 / int with CVTTSS2SL / CVTTSS2SQ on amd64: truncation toward zero,
 * the "integer indefinite" value (sign bit only) for NaN and out-of-range inputs. */
static inline uint8_t go_u8(float v)
{
    int32_t t = (v != v || v >= 2147483648.0f || v < -2147483648.0f) ? INT32_MIN : (int32_t)v;
    return (uint8_t)t;
}
static inline int64_t go_int(float v)
{
    return (v != v || v >= 9223372036854775808.0f || v < -9223372036854775808.0f) ? INT64_MIN : (int64_t)v;
}

static void sci_color(float val, float minVal, float maxVal, uint8_t *px)   /* colors.go:48-84 */
{
    val = go_minf(go_maxf(val, minVal), maxVal - 0.0001f);
    float d = maxVal - minVal;
    if (d <= 0.0f) val = 0.5f;
    else val = (val - minVal) / d;
    const float m = 0.25f;
    float num = (float)floor((double)(val / m));
    float t = num * m;
    float s = (val - t) / m;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    if (num == 0.0f) { r = 0.0f; g = s; b = 1.0f; }
    else if (num == 1.0f) { r = 0.0f; g = 1.0f; b = 1.0f - s; }
    else if (num == 2.0f) { r = s; g = 1.0f; b = 0.0f; }
    else if (num == 3.0f) { r = 1.0f; g = 1.0f - s; b = 0.0f; }
    px[0] = go_u8(255.0f * r); px[1] = go_u8(255.0f * g); px[2] = go_u8(255.0f * b); px[3] = 0xff;
}

static void diverging_color(float val, float minVal, float maxVal, uint8_t *px)   /* colors.go:8-46 */
{
    float absMax = (float)fmax(fabs((double)minVal), fabs((double)maxVal));
    if (absMax < 1e-8f) { px[0] = px[1] = px[2] = 255; px[3] = 255; return; }
    float t = val / absMax;
    if (t > 1.0f) t = 1.0f;
    if (t < -1.0f) t = -1.0f;
    float r, g, b;
    if (t >= 0.0f) { r = 1.0f; g = 1.0f - t; b = 1.0f - t; }
    else { float a = -t; r = 1.0f - a; g = 1.0f - a; b = 1.0f; }
    px[0] = go_u8(255.0f * r); px[1] = go_u8(255.0f * g); px[2] = go_u8(255.0f * b); px[3] = 0xff;
}

void fo_render(const fo_fluid *f, int kind, uint8_t *rgba)
{
    const int64_t NX = f->NumX, NY = f->NumY;
    float *vals = NULL;
    const float *src;
    float mn, mx;
    if (kind == 0) { src = f->M; fo_minmax(src, f->numCells, &mn, &mx); }
    else if (kind == 1) { src = f->p; fo_minmax(src, f->numCells, &mn, &mx); }
    else {
        vals = (float *)malloc((size_t)f->numCells * sizeof(float));
        if (kind == 2) fo_velocity_magnitude(f, vals, &mn, &mx); else fo_vorticity(f, vals, &mn, &mx);
        src = vals;
    }
    for (int64_t i = 0; i < NX; i++)
        for (int64_t jj = 0; jj < NY; jj++) {
            const int64_t j = NY - jj - 1;
            uint8_t *px = rgba + 4 * i + jj * 4 * NX;                  /* fluidToImageIndex, Stride = 4*NumX */
            if (kind == 3) diverging_color(src[i * NY + j], mn, mx, px);    /* drawVorticityField */
            else sci_color(src[i * NY + j], mn, mx, px);                    /* drawScalarField */
            if (f->S[i * NY + j] == 0.0f) { px[0] = 0; px[1] = 0; px[2] = 0; px[3] = 0xff; }   /* main.go:564-574 */
        }
    free(vals);
}

int64_t fo_advect_particles(const fo_fluid *f, fo_particle *ps, int64_t n, float dt)
{
    const float h = f->h;
    int64_t alive = 0;
    for (int64_t k = 0; k < n; k++) {
        fo_particle p = ps[k];
        p.age += dt;
        if (p.age > p.max_age) continue;
        float u1, v1, u2, v2;
        fo_sample_velocity(f, p.x, p.y, &u1, &v1);
        float hd = 0.5f * dt;
        float mu = hd * u1, mv = hd * v1;
        float midX = p.x + mu, midY = p.y + mv;
        fo_sample_velocity(f, midX, midY, &u2, &v2);
        float du = dt * u2, dv = dt * v2;
        p.x += du;
        p.y += dv;
        int64_t fi = go_int(p.x / h), fj = go_int(p.y / h);
        if (fi < 0 || fi >= f->NumX || fj < 0 || fj >= f->NumY) continue;
        if (f->S[fi * f->NumY + fj] == 0.0f) continue;
        ps[alive++] = p;
    }
    return alive;
}

/* ---- NOT in the reference: red-black ordering of the fluid.go:196-229 update.
 * One iteration = red half-sweep ((i+j) even) then black half-sweep; within a
 * half-sweep no two updated cells share a face, so the result is independent
 * of traversal order.  copyBorder as fluid.go:145-146; omega per iteration as
 * fluid.go:169-170 except for the closing iteration (see fo_project_redblack). */
static float redblack_half(fo_fluid *f, float relaxation, float cp, int colour)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float *U = f->U, *V = f->V, *S = f->S, *P = f->p;
    const float damping = f->PressureDamping;
    float maxDiv = 0.0f;
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            if (((i + j) & 1) != colour) continue;
            if (S[i * n + j] == 0.0f) continue;
            float sx0 = S[(i - 1) * n + j], sx1 = S[(i + 1) * n + j];
            float sy0 = S[i * n + j - 1], sy1 = S[i * n + j + 1];
            float s = ((sx0 + sx1) + sy0) + sy1;
            if (s == 0.0f) continue;
            float div = ((U[(i + 1) * n + j] - U[i * n + j]) + V[i * n + j + 1]) - V[i * n + j];
            float absDiv = fabsf(div);
            if (absDiv > maxDiv) maxDiv = absDiv;
            float p = -div / s;
            p *= relaxation;
            p *= damping;
            float cpp = cp * p;
            P[i * n + j] += cpp;
            float a = sx0 * p; U[i * n + j] -= a;
            float b = sx1 * p; U[(i + 1) * n + j] += b;
            float c = sy0 * p; V[i * n + j] -= c;
            float d = sy1 * p; V[i * n + j + 1] += d;
        }
    return maxDiv;
}

float fo_project_redblack_sched(fo_fluid *f, const float *omega, unsigned iters, float dt)
{
    fo_copy_border(f, f->newU, f->U);
    fo_copy_border(f, f->newV, f->V);
    float cp = f->density * f->h / dt;
    float maxDiv = 0.0f;
    f->last_iters = 0;
    for (unsigned iter = 0; iter < iters; iter++) {
        float a = redblack_half(f, omega[2 * iter], cp, 0);
        float b = redblack_half(f, omega[2 * iter + 1], cp, 1);
        maxDiv = a > b ? a : b;
        f->last_iters = (int)iter + 1;
        f->last_maxdiv = maxDiv;
    }
    return maxDiv;
}

float fo_project_redblack(fo_fluid *f, unsigned iters, float dt)
{
    float omega[128];
    const float minRelaxation = 1.2f;
    if (iters > 64) iters = 64;
    for (unsigned iter = 0; iter < iters; iter++) {   /* schedule of fluid.go:169-170 */
        float iterProgress = (float)iter / (float)iters;
        float t = (f->Relaxation - minRelaxation) * iterProgress;
        omega[2 * iter] = omega[2 * iter + 1] = f->Relaxation - t;
    }
    /* Red-black leaves the whole residual on the colour updated first; closing with
     * a plain Gauss-Seidel red half sweep and a half-relaxed black one spreads it over
     * both colours, which is what brings max|div| down to the lexicographic solver's. */
    if (iters > 0) { omega[2 * iters - 2] = 1.0f; omega[2 * iters - 1] = 0.5f; }
    return fo_project_redblack_sched(f, omega, iters, dt);
}

/* ---- NOT in the reference: pressure-form red-black (see fluid_oracle.h) ----------
 * Cell update (active colour):  nb = ((q[i-1,j] + q[i+1,j]) + q[i,j-1]) + q[i,j+1]
 *                               q' = fma(wd*rs, nb - D0, fma(-wd, q, q))
 * with wd = omega*PressureDamping, rs = 1/s (0 for cells the reference skips) and
 * D0 the divergence of the field the pass started from.  Materialisation:
 *   U[i,j] = (U0 - (S[i-1,j] ? q[i,j] : 0)) + (S[i,j] ? q[i-1,j] : 0),  V alike,
 *   p[i,j] = fma(cp, q[i,j], p[i,j]). */
static void redblack_q_pass(fo_fluid *f, const float *omega, unsigned iters, float cp)
{
    const int64_t n = f->NumY, NX = f->NumX, NY = f->NumY;
    float *U = f->U, *V = f->V, *S = f->S, *P = f->p;
    float *q = (float *)calloc((size_t)f->numCells, sizeof(float));
    float *D0 = (float *)calloc((size_t)f->numCells, sizeof(float));
    float *R = (float *)calloc((size_t)f->numCells, sizeof(float));
    float *NS = (float *)calloc((size_t)f->numCells, sizeof(float));
    static const float rs_of[5] = { 0.0f, 1.0f, 0.5f, 1.0f / 3.0f, 0.25f };
    for (int64_t i = 1; i < NX - 1; i++)
        for (int64_t j = 1; j < NY - 1; j++) {
            if (S[i * n + j] == 0.0f) continue;
            int ns = (S[(i - 1) * n + j] != 0.0f) + (S[(i + 1) * n + j] != 0.0f) + (S[i * n + j - 1] != 0.0f) +
                     (S[i * n + j + 1] != 0.0f);
            R[i * n + j] = rs_of[ns];
            NS[i * n + j] = (float)ns;
            D0[i * n + j] = ((U[(i + 1) * n + j] - U[i * n + j]) + V[i * n + j + 1]) - V[i * n + j];
        }
    float maxDiv = 0.0f;
    for (unsigned it = 0; it < iters; it++) {
        maxDiv = 0.0f;
        for (int colour = 0; colour < 2; colour++) {
            const float wd = omega[2 * it + colour] * f->PressureDamping;
            for (int64_t i = 1; i < NX - 1; i++)
                for (int64_t j = 1; j < NY - 1; j++) {
                    if (((i + j) & 1) != colour) continue;
                    const float rs = R[i * n + j];
                    const float nb = ((q[(i - 1) * n + j] + q[(i + 1) * n + j]) + q[i * n + j - 1]) + q[i * n + j + 1];
                    const float t = nb - D0[i * n + j];
                    if (rs != 0.0f) {       /* pre-update divergence, for the statistics only */
                        const float div = fmaf(NS[i * n + j], q[i * n + j], -t);
                        if (fabsf(div) > maxDiv) maxDiv = fabsf(div);
                    }
                    const float cc = wd * rs;
                    const float r0 = fmaf(-wd, q[i * n + j], q[i * n + j]);
                    q[i * n + j] = fmaf(cc, t, r0);
                }
        }
        f->last_maxdiv = maxDiv;
    }
    for (int64_t i = 0; i < NX; i++)
        for (int64_t j = 0; j < NY; j++) {
            const int c = S[i * n + j] != 0.0f;
            const int xm = i > 0 && S[(i - 1) * n + j] != 0.0f;
            const int ym = j > 0 && S[i * n + j - 1] != 0.0f;
            const float qc = q[i * n + j];
            const float qxm = i > 0 ? q[(i - 1) * n + j] : 0.0f;
            const float qym = j > 0 ? q[i * n + j - 1] : 0.0f;
            float a = xm ? qc : 0.0f, b = c ? qxm : 0.0f;
            float t1 = U[i * n + j] - a;
            U[i * n + j] = t1 + b;
            a = ym ? qc : 0.0f; b = c ? qym : 0.0f;
            t1 = V[i * n + j] - a;
            V[i * n + j] = t1 + b;
            P[i * n + j] = fmaf(cp, qc, P[i * n + j]);
        }
    free(q); free(D0); free(R); free(NS);
}

float fo_project_redblack_q(fo_fluid *f, unsigned iters, float dt)
{
    float omega[128];
    const float minRelaxation = 1.2f;
    if (iters > 64) iters = 64;
    for (unsigned iter = 0; iter < iters; iter++) {
        float iterProgress = (float)iter / (float)iters;
        float t = (f->Relaxation - minRelaxation) * iterProgress;
        omega[2 * iter] = omega[2 * iter + 1] = f->Relaxation - t;
    }
    if (iters > 0) { omega[2 * iters - 2] = 1.0f; omega[2 * iters - 1] = 0.5f; }
    fo_copy_border(f, f->newU, f->U);
    fo_copy_border(f, f->newV, f->V);
    float cp = f->density * f->h / dt;
    f->last_iters = (int)iters;
    for (unsigned done = 0; done < iters; ) {       /* passes of at most 8 iterations, like the kernel */
        unsigned k = iters - done < 8 ? iters - done : 8;
        redblack_q_pass(f, omega + 2 * done, k, cp);
        done += k;
    }
    return f->last_maxdiv;
}

/* ---- edit command lists (test convenience; semantics = the point edits) ---- */
int fo_apply_edits(fo_fluid *f, const fo_edit_cmd *cmds, int64_t n)
{
    for (int64_t q = 0; q < n; q++) {
        const fo_edit_cmd *c = &cmds[q];
        if (c->op == FO_EDIT_RESET) { fo_reset(f); continue; }
        if (c->op == FO_EDIT_CIRCLE_OBSTACLE) { fo_set_circular_obstacle(f, c->i0, c->j0, c->i1); continue; }
        for (int64_t i = c->i0; i < c->i1; i++)
            for (int64_t j = c->j0; j < c->j1; j++) {
                int rc = 0;
                switch (c->op) {
                case FO_EDIT_SET_SOLID: rc = fo_set_solid(f, i, j, c->a != 0.0f); break;
                case FO_EDIT_SET_VELOCITY: rc = fo_set_velocity(f, i, j, c->a, c->b); break;
                case FO_EDIT_ADD_SMOKE: rc = fo_add_smoke(f, i, j, c->a); break;
                case FO_EDIT_APPLY_FORCE: fo_apply_force(f, i, j, c->a, c->b); break;
                case FO_EDIT_SET_VELOCITY_IF_FLUID:
                    rc = fo_is_solid(f, i, j);
                    if (rc == 0) rc = fo_set_velocity(f, i, j, c->a, c->b); else if (rc > 0) rc = 0;
                    break;
                case FO_EDIT_ADD_SMOKE_IF_FLUID:
                    rc = fo_is_solid(f, i, j);
                    if (rc == 0) rc = fo_add_smoke(f, i, j, c->a); else if (rc > 0) rc = 0;
                    break;
                case FO_EDIT_SET_SMOKE:
                    if (i < 0 || i >= f->NumX || j < 0 || j >= f->NumY) rc = -1;
                    else f->M[i * f->NumY + j] = c->a;
                    break;
                default: return -1;
                }
                if (rc < 0) return -1;
            }
    }
    return 0;
}

int fo_run(fo_fluid *f, float dt, int64_t nsteps, const fo_edit_cmd *per_step, int64_t n)
{
    for (int64_t s = 0; s < nsteps; s++) {
        if (n > 0 && fo_apply_edits(f, per_step, n) != 0) return -1;
        fo_simulate(f, dt);
    }
    return 0;
}
