// multigrid.cuh -- the 2-level V-cycle of solveMultigridVCycle (fluid.go:560-758,
// 1123-1149) between its fine-grid smoothing sweeps (those are the ordinary projection
// sweeps of fluid.go:188-234 and reuse k_gs_wavefront / k_redblack_half).
//
// The reference materialises four temporaries per cycle (residual, coarse RHS, coarse
// correction, prolongated correction).  Residual and correction are pure functions of
// the fields they are computed from, so they are evaluated where they are consumed:
//   k_mg_restrict   = computePressureResidual + restrictResidual   (reads U,V,p,S; writes the coarse RHS)
//   k_mg_apply      = prolongateCorrection + applePressureCorrection (reads the coarse correction; updates U,V,p)
// Same float32 operations in the same order as the Go code per produced value
// (--fmad=false), so the results are bit-identical to the reference's arrays.
//
// Coarse grid: cNX = (NumX+1)/2, cNY = (NumY+1)/2 (fluid.go:632-633), line ci contiguous
// along cj with a pitch padded to 32 floats; coarse cell (ci,cj) sits on fine cell (2ci,2cj).
#pragma once
#ifdef FB_HOST_EMULATION      // tests/emul/: this header compiled by g++, kernels run thread by thread
#include "grid.cuh"
#else
#include "kernels.cuh"
#endif

struct CoarseGrid {
    int NX, NY;      // cNX, cNY
    int pitch;
    __host__ __device__ __forceinline__ size_t at(int i, int j) const { return (size_t)i * (size_t)pitch + (size_t)j; }
};

// residual[i,j] of computePressureResidual (fluid.go:603-628); 0 where the reference leaves
// its freshly allocated array untouched (ring, solid cells).
__device__ __forceinline__ float mg_residual(const Grid &g, const float *__restrict__ U, const float *__restrict__ V,
                                             const float *__restrict__ S, const float *__restrict__ P, int i, int j)
{
    if (i < 1 || i > g.NX - 2 || j < 1 || j > g.NY - 2) return 0.0f;
    const size_t a = g.at(i, j);
    if (S[a] == 0.0f) return 0.0f;
    const float div = ((U[g.at(i + 1, j)] - U[a]) + V[a + 1]) - V[a];
    const float sx0 = S[g.at(i - 1, j)], sx1 = S[g.at(i + 1, j)];
    const float sy0 = S[a - 1], sy1 = S[a + 1];
    const float pc = P[a];
    const float ta = sx0 * (P[g.at(i - 1, j)] - pc);
    const float tb = sx1 * (P[g.at(i + 1, j)] - pc);
    const float tc = sy0 * (P[a - 1] - pc);
    const float td = sy1 * (P[a + 1] - pc);
    const float laplacian = ((ta + tb) + tc) + td;
    return -div - laplacian;
}

// restrictResidual (fluid.go:631-663) with the residual evaluated in place, plus the coarse
// solid mask and the zeroed coarse pressure of solveCoarseGrid (fluid.go:666-686).
__global__ void k_mg_restrict(Grid g, CoarseGrid c, const float *__restrict__ U, const float *__restrict__ V,
                              const float *__restrict__ S, const float *__restrict__ P,
                              float *__restrict__ rhs, float *__restrict__ cS, float *__restrict__ cP)
{
    const int cj = blockIdx.x * blockDim.x + threadIdx.x;
    const int ci = blockIdx.y * blockDim.y + threadIdx.y;
    if (ci >= c.NX || cj >= c.NY) return;
    const int fi = 2 * ci, fj = 2 * cj;
    float r = 0.0f;
    if (ci >= 1 && ci <= c.NX - 2 && cj >= 1 && cj <= c.NY - 2 && fi < g.NX - 1 && fj < g.NY - 1) {
        const float center = mg_residual(g, U, V, S, P, fi, fj) * 0.25f;
        const float nb = (((mg_residual(g, U, V, S, P, fi - 1, fj) + mg_residual(g, U, V, S, P, fi + 1, fj)) +
                           mg_residual(g, U, V, S, P, fi, fj - 1)) + mg_residual(g, U, V, S, P, fi, fj + 1)) * 0.125f;
        const float cr = (((mg_residual(g, U, V, S, P, fi - 1, fj - 1) + mg_residual(g, U, V, S, P, fi + 1, fj - 1)) +
                           mg_residual(g, U, V, S, P, fi - 1, fj + 1)) + mg_residual(g, U, V, S, P, fi + 1, fj + 1)) * 0.0625f;
        r = (center + nb) + cr;
    }
    const size_t a = c.at(ci, cj);
    rhs[a] = r;
    cS[a] = (fi < g.NX && fj < g.NY) ? S[g.at(fi, fj)] : 0.0f;
    cP[a] = 0.0f;
}

// One coarse cell update of solveCoarseGrid (fluid.go:692-721).
__device__ __forceinline__ void mg_coarse_cell(const CoarseGrid &c, float *__restrict__ cP, const float *__restrict__ cS,
                                               const float *__restrict__ rhs, int i, int j, float relaxation)
{
    const size_t a = c.at(i, j);
    if (cS[a] == 0.0f) return;
    const float sx0 = cS[c.at(i - 1, j)], sx1 = cS[c.at(i + 1, j)];
    const float sy0 = cS[a - 1], sy1 = cS[a + 1];
    const float s = ((sx0 + sx1) + sy0) + sy1;
    if (s == 0.0f) return;
    const float pc = cP[a];
    const float ta = sx0 * (cP[c.at(i - 1, j)] - pc);
    const float tb = sx1 * (cP[c.at(i + 1, j)] - pc);
    const float tc = sy0 * (cP[a - 1] - pc);
    const float td = sy1 * (cP[a + 1] - pc);
    const float laplacian = ((ta + tb) + tc) + td;
    const float residual = rhs[a] - laplacian;
    const float correction = (-residual / s) * relaxation;
    cP[a] = pc + correction;
}

// Exact (lexicographic, in place) coarse solve: cell (i,j) of sweep t reads (i-1,j), (i,j-1)
// of sweep t and (i+1,j), (i,j+1) of sweep t-1, so all (t,i,j) with equal tau = i + j + 2t
// are independent and running tau in increasing order reproduces the 40 sequential sweeps
// bit for bit.  One launch per tau; blockIdx.y = sweep.
__global__ void k_mg_coarse_diag(CoarseGrid c, float *__restrict__ cP, const float *__restrict__ cS,
                                 const float *__restrict__ rhs, int tau, int nsweeps, float relaxation)
{
    const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (t >= nsweeps || i > c.NX - 2) return;
    const int j = tau - 2 * t - i;
    if (j < 1 || j > c.NY - 2) return;
    mg_coarse_cell(c, cP, cS, rhs, i, j, relaxation);
}

// Fast mode: the same coarse update in red-black order, one launch per half sweep.
__global__ void k_mg_coarse_redblack(CoarseGrid c, float *__restrict__ cP, const float *__restrict__ cS,
                                     const float *__restrict__ rhs, int colour, float relaxation)
{
    const int jj = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    if (i > c.NX - 2) return;
    const int j = 2 * jj + (((i + colour) & 1) ? 1 : 0);   // (i+j)&1 == colour
    if (j < 1 || j > c.NY - 2) return;
    mg_coarse_cell(c, cP, cS, rhs, i, j, relaxation);
}

// correction[i,j] of prolongateCorrection (fluid.go:727-758).
__device__ __forceinline__ float mg_correction(const Grid &g, const CoarseGrid &c, const float *__restrict__ cc, int i, int j)
{
    const int ci = i / 2, cj = j / 2;
    if (ci >= c.NX - 1 || cj >= c.NY - 1) return 0.0f;
    const float fracI = (float)(i % 2) * 0.5f;
    const float fracJ = (float)(j % 2) * 0.5f;
    const float w00 = (1.0f - fracI) * (1.0f - fracJ);
    const float w10 = fracI * (1.0f - fracJ);
    const float w01 = (1.0f - fracI) * fracJ;
    const float w11 = fracI * fracJ;
    const float t0 = w00 * cc[c.at(ci, cj)];
    const float t1 = w10 * cc[c.at(ci + 1, cj)];
    const float t2 = w01 * cc[c.at(ci, cj + 1)];
    const float t3 = w11 * cc[c.at(ci + 1, cj + 1)];
    return ((t0 + t1) + t2) + t3;
}

// applePressureCorrection (fluid.go:1123-1149), gathered per face.  The reference walks the
// cells in lexicographic order; face U[i,j] receives "+= S[i,j]*corr(i-1,j)" from cell
// (i-1,j) first and "-= S[i-1,j]*corr(i,j)" from cell (i,j) second, V[i,j] likewise from
// cells (i,j-1) then (i,j).  A thread owns U[i,j], V[i,j], p[i,j], so the update is in place.
__global__ void k_mg_apply(Grid g, CoarseGrid c, float *__restrict__ U, float *__restrict__ V,
                           const float *__restrict__ S, float *__restrict__ P, const float *__restrict__ cc, float cp)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < 1 || i > g.NX - 1 || j < 1 || j > g.NY - 1) return;
    const size_t a = g.at(i, j);
    const bool in_i = i <= g.NX - 2, in_j = j <= g.NY - 2;
    const float sc = S[a];
    const float sl = S[g.at(i - 1, j)];
    const float sd = S[a - 1];
    const bool self = in_i && in_j && sc != 0.0f;          // cell (i,j) is visited
    const bool left = in_j && i - 1 >= 1 && sl != 0.0f;    // cell (i-1,j) is visited
    const bool down = in_i && j - 1 >= 1 && sd != 0.0f;    // cell (i,j-1) is visited
    const float corr = self ? mg_correction(g, c, cc, i, j) : 0.0f;
    if (left || self) {
        float u = U[a];
        if (left) { const float b = sc * mg_correction(g, c, cc, i - 1, j); u += b; }
        if (self) { const float q = sl * corr; u -= q; }
        U[a] = u;
    }
    if (down || self) {
        float v = V[a];
        if (down) { const float d = sc * mg_correction(g, c, cc, i, j - 1); v += d; }
        if (self) { const float q = sd * corr; v -= q; }
        V[a] = v;
    }
    if (self) {
        const float t = corr * cp;
        P[a] += t;
    }
}
