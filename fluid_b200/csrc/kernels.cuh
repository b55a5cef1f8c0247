// kernels.cuh -- sm_100a kernels for the pkg/fluid per-step hot path.
//
// Arithmetic contract: every kernel that has a counterpart in the reference
// performs the SAME float32 operations in the SAME order as the Go code on
// amd64 (no fused multiply-add: this file is compiled with --fmad=false;
// IEEE division and square root are nvcc's defaults).  Citations are
// pkg/fluid/<file>:<line> of the reference.
//
// Layout: one float32 plane per field, line i (constant x index) contiguous
// along j with a padded pitch (multiple of 32 floats, so every line starts on a
// 128-byte boundary and float4 row access is aligned).  Lines [i_alloc0,
// i_alloc0 + lines_alloc) of the GLOBAL grid are allocated on this rank.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/fluidb200.h"
#include "grid.cuh"

struct SolveParams {
    float omega[64];   // relaxation per sweep (exact) or per half sweep (red-black)
    float damping;     // PressureDamping (fluid.go:222)
    float cp;          // density*h/dt (fluid.go:158)
    int sweeps;        // sweeps fused in this launch
    int sweep0;        // index of the first sweep of this launch (stats slot)
};

// ---- helpers ---------------------------------------------------------------
__device__ __forceinline__ float go_minf(float a, float b) { return (a < b) ? a : ((b < a) ? b : (a != a ? a : b)); }
__device__ __forceinline__ float go_maxf(float a, float b) { return (a > b) ? a : ((b > a) ? b : (a != a ? a : b)); }

__device__ __forceinline__ unsigned f2key(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide max of non-negative floats -> atomicMax on the int bit pattern.
__device__ __forceinline__ void block_atomic_max_nonneg(float v, unsigned *slot) {
    __shared__ float s_red[32];
    v = warp_max(v);
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        float w = tid < nw ? s_red[tid] : 0.0f;
        w = warp_max(w);
        if (tid == 0 && w > 0.0f) atomicMax(slot, __float_as_uint(w));
    }
    __syncthreads();
}

// ---- the per-cell projection update (fluid.go:196-229, Q-4) -----------------
// Returns pre-update |div| (0 when the cell is skipped).
struct CellS { float c, sx0, sx1, sy0, sy1; };

__device__ __forceinline__ float project_cell(float &u0, float &u1, float &v0, float &v1, float &p,
                                              const CellS &s, float omega, float damping, float cp)
{
    if (s.c == 0.0f) return 0.0f;
    float ssum = ((s.sx0 + s.sx1) + s.sy0) + s.sy1;
    if (ssum == 0.0f) return 0.0f;
    float div = ((u1 - u0) + v1) - v0;
    float pp = -div / ssum;
    pp *= omega;
    pp *= damping;
    float cpp = cp * pp;
    p += cpp;
    float a = s.sx0 * pp; u0 -= a;
    float b = s.sx1 * pp; u1 += b;
    float c = s.sy0 * pp; v0 -= c;
    float d = s.sy1 * pp; v1 += d;
    return fabsf(div);
}

// ---- copyBorder (fluid.go:436-446) ------------------------------------------
__global__ void k_copy_border(Grid g, float *__restrict__ dst, const float *__restrict__ src, int ib, int ie)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nl = ie - ib;
    if (t < nl) {
        const int i = ib + t;
        dst[g.at(i, 0)] = src[g.at(i, 0)];
        dst[g.at(i, g.NY - 1)] = src[g.at(i, g.NY - 1)];
    } else if (t < nl + g.NY) {
        const int j = t - nl;
        if (ib <= 0 && 0 < ie) dst[g.at(0, j)] = src[g.at(0, j)];
    } else if (t < nl + 2 * g.NY) {
        const int j = t - nl - g.NY;
        if (ib <= g.NX - 1 && g.NX - 1 < ie) dst[g.at(g.NX - 1, j)] = src[g.at(g.NX - 1, j)];
    }
}

// ---- handleBorders (fluid.go:236-289, Q-13) -----------------------------------
__global__ void k_handle_borders(Grid g, float *__restrict__ U, float *__restrict__ V,
                                 const float *__restrict__ S, int ib, int ie)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int nl = ie - ib;
    const int NX = g.NX, NY = g.NY;
    if (t < nl) {
        const int i = ib + t;
        if (S[g.at(i, 0)] == 0.0f || S[g.at(i, 1)] == 0.0f) {
            U[g.at(i, 0)] = 0.0f;
        } else if (i > 0 && i < NX - 1 && S[g.at(i, 2)] > 0.0f) {
            float a = 2.0f * U[g.at(i, 1)];
            U[g.at(i, 0)] = a - U[g.at(i, 2)];
        } else {
            U[g.at(i, 0)] = U[g.at(i, 1)];
        }
        if (S[g.at(i, NY - 1)] == 0.0f || S[g.at(i, NY - 2)] == 0.0f) {
            U[g.at(i, NY - 1)] = 0.0f;
        } else if (i > 0 && i < NX - 1 && S[g.at(i, NY - 3)] > 0.0f) {
            float a = 2.0f * U[g.at(i, NY - 2)];
            U[g.at(i, NY - 1)] = a - U[g.at(i, NY - 3)];
        } else {
            U[g.at(i, NY - 1)] = U[g.at(i, NY - 2)];
        }
    } else if (t < nl + NY) {
        const int j = t - nl;
        if (ib <= 0 && 0 < ie) {   // left border i == 0 (needs lines 1, 2)
            if (S[g.at(0, j)] == 0.0f || S[g.at(1, j)] == 0.0f) {
                V[g.at(0, j)] = 0.0f;
            } else if (j > 0 && j < NY - 1 && S[g.at(2, j)] > 0.0f) {
                float a = 2.0f * V[g.at(1, j)];
                V[g.at(0, j)] = a - V[g.at(2, j)];
            } else {
                V[g.at(0, j)] = V[g.at(1, j)];
            }
        }
    } else if (t < nl + 2 * NY) {
        const int j = t - nl - NY;
        if (ib <= NX - 1 && NX - 1 < ie) {
            if (S[g.at(NX - 1, j)] == 0.0f || S[g.at(NX - 2, j)] == 0.0f) {
                V[g.at(NX - 1, j)] = 0.0f;
            } else if (j > 0 && j < NY - 1 && S[g.at(NX - 3, j)] > 0.0f) {
                float a = 2.0f * V[g.at(NX - 2, j)];
                V[g.at(NX - 1, j)] = a - V[g.at(NX - 3, j)];
            } else {
                V[g.at(NX - 1, j)] = V[g.at(NX - 2, j)];
            }
        }
    }
}

// ---- red-black half sweep, unfused (reference kernel for the fused one) -----
// Same per-cell update as fluid.go:196-229 applied to all cells of one colour.
__global__ void k_redblack_half(Grid g, float *__restrict__ U, float *__restrict__ V,
                                const float *__restrict__ S, float *__restrict__ P,
                                int colour, float omega, float damping, float cp,
                                unsigned *stat_slot, int ib, int ie)
{
    const int jj = blockIdx.x * blockDim.x + threadIdx.x;   // index among cells of this colour
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    float adiv = 0.0f;
    if (i < ie && i >= 1 && i <= g.NX - 2) {
        const int j = 2 * jj + (((i + colour) & 1) ? 1 : 0);   // (i+j)&1 == colour
        if (j >= 1 && j <= g.NY - 2) {
            CellS s;
            s.c = S[g.at(i, j)];
            if (s.c != 0.0f) {
                s.sx0 = S[g.at(i - 1, j)]; s.sx1 = S[g.at(i + 1, j)];
                s.sy0 = S[g.at(i, j - 1)]; s.sy1 = S[g.at(i, j + 1)];
                float u0 = U[g.at(i, j)], u1 = U[g.at(i + 1, j)];
                float v0 = V[g.at(i, j)], v1 = V[g.at(i, j + 1)];
                float p = P[g.at(i, j)];
                float ssum = ((s.sx0 + s.sx1) + s.sy0) + s.sy1;
                if (ssum != 0.0f) {
                    adiv = project_cell(u0, u1, v0, v1, p, s, omega, damping, cp);
                    U[g.at(i, j)] = u0; U[g.at(i + 1, j)] = u1;
                    V[g.at(i, j)] = v0; V[g.at(i, j + 1)] = v1;
                    P[g.at(i, j)] = p;
                }
            }
        }
    }
    block_atomic_max_nonneg(adiv, stat_slot);
}

// ---- exact lexicographic Gauss-Seidel/SOR by skewed-tile wavefront ----------
// Cell (i,j) of sweep t depends on (i-1,j,t), (i,j-1,t), (i+1,j,t-1), (i,j+1,t-1)
// (fluid.go:192-231 is in place and lexicographic).  In skewed coordinates
// i' = i+t, j' = j+t every dependence is non-positive in (t,i',j'), so TI x TJ
// tiles of (i',j') holding all fused sweeps can run whole, ordered by their own
// anti-diagonals; inside a tile all (t,i',j') with equal i'+j' are independent.
// Result is bit-identical to the sequential sweeps, including max|div| per sweep.
#define WF_TI 32
#define WF_TJ 32
#define WF_TMAX 8
#define WF_RI (WF_TI + WF_TMAX)           // face/pressure region lines
#define WF_RJ (WF_TJ + WF_TMAX + 2)       // 42: anti-diagonal stride 41 is odd -> no bank conflicts
#define WF_SI (WF_TI + WF_TMAX + 1)
#define WF_SJ (WF_TJ + WF_TMAX + 3)

__global__ void __launch_bounds__(WF_TI * WF_TMAX)
k_gs_wavefront(Grid g, float *__restrict__ U, float *__restrict__ V, const float *__restrict__ S,
               float *__restrict__ P, SolveParams sp, const int2 *__restrict__ order, int ntiles,
               int nTb, int *tile_counter, volatile int *done, int epoch, unsigned *stats)
{
    __shared__ float sU[WF_RI][WF_RJ], sV[WF_RI][WF_RJ], sP[WF_RI][WF_RJ];
    __shared__ float sS[WF_SI][WF_SJ];
    __shared__ int s_tile;
    __shared__ float s_max[WF_TMAX][WF_TI / 32 + 1];

    const int k = threadIdx.x;        // position along i' inside the tile
    const int t = threadIdx.y;        // sweep (local to this launch)
    const int tid = t * WF_TI + k;
    const int nthreads = WF_TI * blockDim.y;
    const int T = sp.sweeps;
    const int NX = g.NX, NY = g.NY;

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) break;
        const int2 ab = order[tile];
        const int I0 = 1 + ab.x * WF_TI;   // first i' of the tile
        const int J0 = 1 + ab.y * WF_TJ;
        // wait for the two predecessor tiles
        if (tid == 0) {
            if (ab.x > 0) while (done[(ab.x - 1) * nTb + ab.y] != epoch) { }
            if (ab.y > 0) while (done[ab.x * nTb + (ab.y - 1)] != epoch) { }
            __threadfence();
        }
        __syncthreads();

        const int ri0 = I0 - (WF_TMAX - 1);   // global i of region line 0
        const int rj0 = J0 - (WF_TMAX - 1);
        for (int e = tid; e < WF_RI * WF_RJ; e += nthreads) {
            const int li = e / WF_RJ, lj = e - li * WF_RJ;
            const int i = ri0 + li, j = rj0 + lj;
            float u = 0.f, v = 0.f, p = 0.f;
            if (i >= 0 && i < NX && j >= 0 && j < NY) {
                const size_t a = g.at(i, j);
                u = __ldcg(U + a); v = __ldcg(V + a); p = __ldcg(P + a);
            }
            sU[li][lj] = u; sV[li][lj] = v; sP[li][lj] = p;
        }
        for (int e = tid; e < WF_SI * WF_SJ; e += nthreads) {
            const int li = e / WF_SJ, lj = e - li * WF_SJ;
            const int i = ri0 - 1 + li, j = rj0 - 1 + lj;
            float s = 0.f;
            if (i >= 0 && i < NX && j >= 0 && j < NY) s = S[g.at(i, j)];
            sS[li][lj] = s;
        }
        __syncthreads();

        float mymax = 0.0f;
        const float omega = sp.omega[sp.sweep0 + (t < T ? t : 0)];
        for (int s = 0; s < WF_TI + WF_TJ - 1; s++) {
            const int bq = s - k;
            if (t < T && bq >= 0 && bq < WF_TJ) {
                const int i = I0 + k - t, j = J0 + bq - t;
                if (i >= 1 && i <= NX - 2 && j >= 1 && j <= NY - 2) {
                    const int li = i - ri0, lj = j - rj0;
                    CellS c;
                    c.c = sS[li + 1][lj + 1];
                    c.sx0 = sS[li][lj + 1]; c.sx1 = sS[li + 2][lj + 1];
                    c.sy0 = sS[li + 1][lj]; c.sy1 = sS[li + 1][lj + 2];
                    float u0 = sU[li][lj], u1 = sU[li + 1][lj];
                    float v0 = sV[li][lj], v1 = sV[li][lj + 1];
                    float p = sP[li][lj];
                    float ad = project_cell(u0, u1, v0, v1, p, c, omega, sp.damping, sp.cp);
                    if (ad > mymax) mymax = ad;
                    sU[li][lj] = u0; sU[li + 1][lj] = u1;
                    sV[li][lj] = v0; sV[li][lj + 1] = v1;
                    sP[li][lj] = p;
                }
            }
            __syncthreads();
        }
        // per-sweep max |div| (threads of one sweep = one row of the block)
        mymax = warp_max(mymax);
        if ((k & 31) == 0) s_max[t][k >> 5] = mymax;

        // write back exactly the faces / pressures this tile's cells own
        for (int e = tid; e < WF_RI * WF_RJ; e += nthreads) {
            const int li = e / WF_RJ, lj = e - li * WF_RJ;
            const int i = ri0 + li, j = rj0 + lj;
            if (i < 0 || i >= NX || j < 0 || j >= NY) continue;
            bool mu = false, mv = false, mp = false;
            for (int tt = 0; tt < T; tt++) {
                const int ip = i + tt, jp = j + tt;   // skewed coords of cell (i,j) at sweep tt
                const bool jin = jp >= J0 && jp < J0 + WF_TJ && j >= 1 && j <= NY - 2;
                const bool iin = ip >= I0 && ip < I0 + WF_TI && i >= 1 && i <= NX - 2;
                const bool self = iin && jin;
                const bool left = (ip - 1) >= I0 && (ip - 1) < I0 + WF_TI && (i - 1) >= 1 && (i - 1) <= NX - 2 && jin;
                const bool down = iin && (jp - 1) >= J0 && (jp - 1) < J0 + WF_TJ && (j - 1) >= 1 && (j - 1) <= NY - 2;
                mp |= self; mu |= self | left; mv |= self | down;
            }
            const size_t a = g.at(i, j);
            if (mu) __stcg(U + a, sU[li][lj]);
            if (mv) __stcg(V + a, sV[li][lj]);
            if (mp) __stcg(P + a, sP[li][lj]);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            for (int tt = 0; tt < T; tt++) {
                float m = 0.0f;
                for (int w = 0; w < WF_TI / 32; w++) m = fmaxf(m, s_max[tt][w]);
                if (m > 0.0f) atomicMax(stats + sp.sweep0 + tt, __float_as_uint(m));
            }
            __threadfence();
            done[ab.x * nTb + ab.y] = epoch;
        }
        __syncthreads();
    }
}

// ---- vorticity confinement (fluid.go:449-493, Q-12) --------------------------
__device__ __forceinline__ float curl_at(const Grid &g, const float *U, const float *V, const float *S,
                                         int i, int j, float h)
{
    if (i < 1 || i > g.NX - 2 || j < 1 || j > g.NY - 2) return 0.0f;
    if (S[g.at(i, j)] == 0.0f) return 0.0f;
    float dvdx = ((V[g.at(i + 1, j)] - V[g.at(i - 1, j)]) * 0.5f) / h;
    float dudy = ((U[g.at(i, j + 1)] - U[g.at(i, j - 1)]) * 0.5f) / h;
    return dvdx - dudy;
}

__global__ void k_curl(Grid g, const float *__restrict__ U, const float *__restrict__ V,
                       const float *__restrict__ S, float *__restrict__ curl, float h, int ib, int ie)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= ie || j >= g.NY) return;
    curl[g.at(i, j)] = curl_at(g, U, V, S, i, j, h);
}

__global__ void k_confine(Grid g, float *__restrict__ U, float *__restrict__ V, const float *__restrict__ S,
                          const float *__restrict__ curl, float h, float dt, float confinement, int ib, int ie)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= ie || i < 1 || i > g.NX - 2 || j < 1 || j > g.NY - 2) return;
    const size_t a = g.at(i, j);
    if (S[a] == 0.0f) return;
    const float eps = 1e-5f;
    float gx = ((fabsf(curl[g.at(i + 1, j)]) - fabsf(curl[g.at(i - 1, j)])) * 0.5f) / h;
    float gy = ((fabsf(curl[g.at(i, j + 1)]) - fabsf(curl[g.at(i, j - 1)])) * 0.5f) / h;
    float gx2 = gx * gx, gy2 = gy * gy;
    float mag = sqrtf(gx2 + gy2) + eps;
    gx /= mag;
    gy /= mag;
    float vort = curl[a];
    float u = U[a], v = V[a];
    float uu = u * u, vv = v * v;
    float localVel = sqrtf(uu + vv);
    float lv = localVel * 0.1f;
    float strength = confinement * (1.0f + lv);
    float fu = ((strength * gy) * vort) * dt;
    float fv = ((strength * gx) * vort) * dt;
    U[a] = u + fu;
    V[a] = v - fv;
}

// ---- turbulence (fluid.go:496-526, Q-5) ---------------------------------------
// The noise is static in time: float32(sin(float64(float32(i*137+j*241)*0.01))).
// It is tabulated once per grid (double-precision sin, like the reference).
__global__ void k_noise_init(Grid g, float *__restrict__ nU, float *__restrict__ nV, int ib, int ie)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= ie || j >= g.NY) return;
    const long long ku = (long long)i * 137 + (long long)j * 241;
    const long long kv = (long long)i * 157 + (long long)j * 263;
    float seedU = __ll2float_rn(ku) * 0.01f;
    float seedV = __ll2float_rn(kv) * 0.01f;
    nU[g.at(i, j)] = (float)sin((double)seedU);
    nV[g.at(i, j)] = (float)sin((double)seedV);
}

__global__ void k_turbulence(Grid g, float *__restrict__ U, float *__restrict__ V, const float *__restrict__ S,
                             const float *__restrict__ nU, const float *__restrict__ nV,
                             float turbStrength, int ib, int ie)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= ie || i < 1 || i > g.NX - 2 || j < 1 || j > g.NY - 2) return;
    const size_t a = g.at(i, j);
    if (!(S[a] > 0.0f)) return;
    float u = U[a], v = V[a];
    float uu = u * u, vv = v * v;
    float localVel = sqrtf(uu + vv);
    if (localVel > 0.1f) {
        float noiseU = nU[a] * turbStrength;
        float noiseV = nV[a] * turbStrength;
        float factor = go_minf(localVel * 0.5f, 1.0f);
        float du = noiseU * factor, dv = noiseV * factor;
        U[a] = u + du;
        V[a] = v + dv;
    }
}

// ---- artificial viscosity (fluid.go:112-142) ----------------------------------
__global__ void k_viscosity(Grid g, const float *__restrict__ U, const float *__restrict__ V,
                            const float *__restrict__ S, float *__restrict__ nU, float *__restrict__ nV,
                            float visc, int ib, int ie)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= ie || i < 1 || i > g.NX - 2 || j < 1 || j > g.NY - 2) return;
    const size_t a = g.at(i, j);
    if (!(S[a] > 0.0f)) return;
    {
        float c4 = 4.0f * U[a];
        float lap = (((U[g.at(i - 1, j)] + U[g.at(i + 1, j)]) + U[g.at(i, j - 1)]) + U[g.at(i, j + 1)]) - c4;
        float t = visc * lap;
        nU[a] = U[a] + t;
    }
    {
        float c4 = 4.0f * V[a];
        float lap = (((V[g.at(i - 1, j)] + V[g.at(i + 1, j)]) + V[g.at(i, j - 1)]) + V[g.at(i, j + 1)]) - c4;
        float t = visc * lap;
        nV[a] = V[a] + t;
    }
}

// ---- sampleField / sampleFieldFrom (fluid.go:357-398, 1055-1091, Q-8) ----------
// fld: 0 = U (dy = h/2), 1 = V (dx = h/2), 2 = M (both).  `bad` is raised when a
// tap lies outside the lines this rank holds (multi-GPU ghost zone exceeded).
template <int FLD>
__device__ __forceinline__ float sample_from(const Grid &g, const float *__restrict__ data,
                                             float x, float y, float h, float h1, float h2, int *bad)
{
    x = go_maxf(go_minf(x, (float)g.NX * h), h);
    y = go_maxf(go_minf(y, (float)g.NY * h), h);
    const float dx = (FLD == 0) ? 0.0f : h2;
    const float dy = (FLD == 1) ? 0.0f : h2;
    float xs = x - dx;
    float xh = xs * h1;
    int x0 = min(max((int)floorf(xh), 0), g.NX - 1);
    float x0h = (float)x0 * h;
    float tx = (xs - x0h) * h1;
    int x1 = min(x0 + 1, g.NX - 1);
    float ys = y - dy;
    float yh = ys * h1;
    int y0 = min(max((int)floorf(yh), 0), g.NY - 1);
    float y0h = (float)y0 * h;
    float ty = (ys - y0h) * h1;
    int y1 = min(y0 + 1, g.NY - 1);
    float sx = 1.0f - tx;
    float sy = 1.0f - ty;
    if (x0 < g.i_alloc0 || x1 >= g.i_alloc0 + g.lines_alloc) { *bad = 1; return 0.0f; }
    float w00 = sx * sy, w10 = tx * sy, w11 = tx * ty, w01 = sx * ty;
    float a = w00 * data[g.at(x0, y0)];
    float b = w10 * data[g.at(x1, y0)];
    float c = w11 * data[g.at(x1, y1)];
    float d = w01 * data[g.at(x0, y1)];
    return ((a + b) + c) + d;
}

// ---- velocity back-trace: advectVelocity (fluid.go:300-329) with BACK=false
// (srcU/srcV = U/V, dst = newU/newV) and the BFECC backward pass
// (fluid.go:943-966) with BACK=true (src = fwd, dst = bwd). --------------------
template <bool BACK>
__global__ void k_trace_velocity(Grid g, const float *__restrict__ U, const float *__restrict__ V,
                                 const float *__restrict__ S, const float *__restrict__ srcU,
                                 const float *__restrict__ srcV, float *__restrict__ dstU,
                                 float *__restrict__ dstV, float dt, float h, int ib, int ie, int *bad)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= ie || i < 1 || i > g.NX - 1 || j < 1 || j > g.NY - 1) return;
    const float h1 = 1.0f / h;
    const float h2 = h / 2.0f;
    const size_t a = g.at(i, j);
    const float sc = S[a];
    if (sc == 0.0f) return;
    const float u_ij = U[a], v_ij = V[a];
    if (S[g.at(i - 1, j)] != 0.0f && j < g.NY - 1) {
        float x = (float)i * h;
        float y = (float)j * h + h2;
        // avgV (fluid.go:342-347)
        float v = (((V[g.at(i - 1, j)] + v_ij) + V[g.at(i - 1, j + 1)]) + V[g.at(i, j + 1)]) * 0.25f;
        float du = dt * u_ij, dv = dt * v;
        if (BACK) { x = x + du; y = y + dv; } else { x = x - du; y = y - dv; }
        dstU[a] = sample_from<0>(g, srcU, x, y, h, h1, h2, bad);
    }
    if (S[g.at(i, j - 1)] != 0.0f && i < g.NX - 1) {
        float x = (float)i * h + h2;
        float y = (float)j * h;
        // avgU (fluid.go:335-340)
        float u = (((U[g.at(i, j - 1)] + u_ij) + U[g.at(i + 1, j - 1)]) + U[g.at(i + 1, j)]) * 0.25f;
        float du = dt * u, dv = dt * v_ij;
        if (BACK) { x = x + du; y = y + dv; } else { x = x - du; y = y - dv; }
        dstV[a] = sample_from<1>(g, srcV, x, y, h, h1, h2, bad);
    }
}

// ---- smoke back-trace: advectSmoke (fluid.go:408-431) / BFECC backward pass
// (fluid.go:1017-1027) ---------------------------------------------------------
template <bool BACK>
__global__ void k_trace_smoke(Grid g, const float *__restrict__ U, const float *__restrict__ V,
                              const float *__restrict__ S, const float *__restrict__ M,
                              const float *__restrict__ src, float *__restrict__ dst, float dt, float h,
                              float smokeAdvection, float viscosityDiffusion, int ib, int ie, int *bad)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= ie || i < 1 || i > g.NX - 2 || j < 1 || j > g.NY - 2) return;
    const size_t a = g.at(i, j);
    if (S[a] == 0.0f) return;
    const float h1 = 1.0f / h;
    const float h2 = 0.5f * h;
    float u = ((U[a] + U[g.at(i + 1, j)]) * 0.5f) * smokeAdvection;
    float v = ((V[a] + V[g.at(i, j + 1)]) * 0.5f) * smokeAdvection;
    float du = dt * u, dv = dt * v;
    float x0 = (float)i * h + h2;
    float y0 = (float)j * h + h2;
    if (BACK) {
        dst[a] = sample_from<2>(g, src, x0 + du, y0 + dv, h, h1, h2, bad);
        return;
    }
    float val = sample_from<2>(g, src, x0 - du, y0 - dv, h, h1, h2, bad);
    if (viscosityDiffusion > 0.0f) {
        float sd = (viscosityDiffusion * 0.3f) * dt;
        float c4 = 4.0f * M[a];
        float nb = (((M[g.at(i - 1, j)] + M[g.at(i + 1, j)]) + M[g.at(i, j - 1)]) + M[g.at(i, j + 1)]) - c4;
        float t = sd * nb;
        val += t;
    }
    dst[a] = go_maxf(val, 0.0f);
}

// ---- BFECC error compensation + clamp (fluid.go:974-987, 1032-1046, 1094-1120) --
// out = clamp(orig - (bwd - orig)*0.5, 3x3 min/max of orig) on interior cells
// regardless of S (Q-10); ring cells take orig.
template <bool NONNEG>
__global__ void k_bfecc_correct(Grid g, const float *__restrict__ orig, const float *__restrict__ bwd,
                                float *__restrict__ out, int ib, int ie)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= ie || j >= g.NY) return;
    const size_t a = g.at(i, j);
    const float o = orig[a];
    if (i < 1 || i > g.NX - 2 || j < 1 || j > g.NY - 2) { out[a] = o; return; }
    float e = (bwd[a] - o) * 0.5f;
    float val = o - e;
    float lo = o, hi = o;
#pragma unroll
    for (int di = -1; di <= 1; di++)
#pragma unroll
        for (int dj = -1; dj <= 1; dj++) {
            float v = orig[g.at(i + di, j + dj)];
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
    if (val < lo) val = lo;
    else if (val > hi) val = hi;
    if (NONNEG && val < 0.0f) val = 0.0f;
    out[a] = val;
}

// ---- views and reductions (Q-14) ------------------------------------------------
// red[0] = min key, red[1] = max key (ordered-uint encoding)
__device__ __forceinline__ void block_minmax(float lo, float hi, unsigned *red)
{
    __shared__ float s_lo[32], s_hi[32];
    lo = warp_min(lo); hi = warp_max(hi);
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; }
    __syncthreads();
    if (tid < 32) {
        float l = tid < nw ? s_lo[tid] : 3.402823466e+38f;
        float m = tid < nw ? s_hi[tid] : -3.402823466e+38f;
        l = warp_min(l); m = warp_max(m);
        if (tid == 0) { atomicMin(red + 0, f2key(l)); atomicMax(red + 1, f2key(m)); }
    }
}

// min/max over ALL cells of the dense array (pressure.go:8-15, smoke.go:8-15).  A block walks whole lines with float4
// loads (every line starts 128-byte aligned, section 3 of DESIGN.md) and posts ONE pair of atomics: a block per 2 x 128
// cells (round 1) spent its time on 131 k same-address atomics at 4098^2.  Launch: 1-D grid, 256 threads.
__global__ void __launch_bounds__(256) k_minmax_all(Grid g, const float *__restrict__ A, unsigned *red, int ib, int ie)
{
    float lo = 3.402823466e+38f, hi = -3.402823466e+38f;
    const int nq = g.NY >> 2;
    for (int i = ib + blockIdx.x; i < ie; i += gridDim.x) {
        const float *row = A + g.at(i, 0);
        const float4 *row4 = reinterpret_cast<const float4 *>(row);
        for (int q = threadIdx.x; q < nq; q += 256) {
            const float4 v = row4[q];
            if (v.x < lo) lo = v.x;
            if (v.x > hi) hi = v.x;
            if (v.y < lo) lo = v.y;
            if (v.y > hi) hi = v.y;
            if (v.z < lo) lo = v.z;
            if (v.z > hi) hi = v.z;
            if (v.w < lo) lo = v.w;
            if (v.w > hi) hi = v.w;
        }
        const int j = (nq << 2) + threadIdx.x;
        if (j < g.NY) {
            const float v = row[j];
            if (v < lo) lo = v;
            if (v > hi) hi = v;
        }
    }
    block_minmax(lo, hi, red);
}
static inline int minmax_blocks(int lines) { return lines < 148 * 8 ? (lines > 0 ? lines : 1) : 148 * 8; }

// Asynchronous views: sentinels of the min/max pair without a host copy, and a compact
// (pitch -> NumY) snapshot of lines [ib, ie) fused with the min/max pass, so that the device
// to host copy is ONE contiguous transfer that may overlap the next Simulate.
__global__ void k_minmax_init(unsigned *red)
{
    red[0] = f2key(3.402823466e+38f);
    red[1] = f2key(-3.402823466e+38f);
}
__global__ void k_snapshot_minmax(Grid g, const float *__restrict__ A, float *__restrict__ snap, unsigned *red,
                                  int ib, int ie, int reduce)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    float lo = 3.402823466e+38f, hi = -3.402823466e+38f;
    if (i < ie && j < g.NY) {
        const float v = A[g.at(i, j)];
        snap[(size_t)(i - ib) * g.NY + j] = v;
        if (v < lo) lo = v;
        if (v > hi) hi = v;
    }
    if (reduce) block_minmax(lo, hi, red);
}

// Decimated 8-bit view for frame loops whose field is larger than any display: every `stride`-th cell of every
// `stride`-th line (global indices that are multiples of stride), quantised against the min / max the reduction of the
// SAME queue left in red[0..1]: index = (unsigned)(min(max((v - lo) * (255 / (hi - lo)), 0), 255) + 0.5).  The host maps
// the index through a 256-entry table of the UI's colormap (main/colors.go).  out[(i/stride - ob) * nj + j/stride].
__device__ __forceinline__ float key2f_dev_(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
__global__ void k_quantize_u8(Grid g, const float *__restrict__ A, unsigned char *__restrict__ out, const unsigned *__restrict__ red,
                              int stride, int oi0, int ni, int nj)
{
    const int jo = blockIdx.x * blockDim.x + threadIdx.x;
    const int io = blockIdx.y * blockDim.y + threadIdx.y;
    if (io >= ni || jo >= nj) return;
    const float lo = key2f_dev_(red[0]), hi = key2f_dev_(red[1]);
    const float range = hi - lo;
    const float scale = range > 0.0f ? 255.0f / range : 0.0f;
    const float v = A[g.at((oi0 + io) * stride, jo * stride)];
    const float q = fminf(fmaxf((v - lo) * scale, 0.0f), 255.0f) + 0.5f;
    out[(size_t)io * nj + jo] = (unsigned char)(unsigned)q;
}

// Vorticity (fluid.go:806-838) / VelocityMagnitude (fluid.go:841-873)
template <int KIND>
__global__ void k_view(Grid g, const float *__restrict__ U, const float *__restrict__ V,
                       const float *__restrict__ S, float *__restrict__ out, float h, unsigned *red,
                       int ib, int ie)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    float lo = 3.402823466e+38f, hi = -3.402823466e+38f;
    if (i < ie && j < g.NY) {
        float val = 0.0f;
        if (i >= 1 && i <= g.NX - 2 && j >= 1 && j <= g.NY - 2 && S[g.at(i, j)] != 0.0f) {
            if (KIND == FB_VIEW_VORTICITY) {
                val = curl_at(g, U, V, S, i, j, h);
            } else {
                float u = (U[g.at(i, j)] + U[g.at(i + 1, j)]) * 0.5f;
                float v = (V[g.at(i, j)] + V[g.at(i, j + 1)]) * 0.5f;
                float uu = u * u, vv = v * v;
                val = sqrtf(uu + vv);
            }
            if (val < lo) lo = val;
            if (val > hi) hi = val;
        }
        out[g.at(i, j)] = val;
    }
    block_minmax(lo, hi, red);
}

// MaxDivergence (fluid.go:876-891) / max(|U|+|V|) (fluid.go:534-543)
template <int KIND>
__global__ void k_reduce_max(Grid g, const float *__restrict__ U, const float *__restrict__ V,
                             const float *__restrict__ S, unsigned *slot, int ib, int ie)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = ib + blockIdx.y * blockDim.y + threadIdx.y;
    float m = 0.0f;
    if (i < ie && i >= 1 && i <= g.NX - 2 && j >= 1 && j <= g.NY - 2) {
        const size_t a = g.at(i, j);
        if (KIND == FB_REDUCE_MAX_DIVERGENCE) {
            if (S[a] != 0.0f) {
                float div = ((U[g.at(i + 1, j)] - U[a]) + V[g.at(i, j + 1)]) - V[a];
                float ad = fabsf(div);
                if (ad > m) m = ad;
            }
        } else {
            if (S[a] > 0.0f) {
                float vel = fabsf(U[a]) + fabsf(V[a]);
                if (vel > m) m = vel;
            }
        }
    }
    block_atomic_max_nonneg(m, slot);
}

__global__ void k_sample_velocity(Grid g, const float *__restrict__ U, const float *__restrict__ V,
                                  const float *__restrict__ xy, float *__restrict__ uv, size_t n, float h, int *bad)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float h1 = 1.0f / h, h2 = 0.5f * h;
    uv[2 * t + 0] = sample_from<0>(g, U, xy[2 * t], xy[2 * t + 1], h, h1, h2, bad);
    uv[2 * t + 1] = sample_from<1>(g, V, xy[2 * t], xy[2 * t + 1], h, h1, h2, bad);
}

// ---- edits (walls.go:5-93, fluid.go:761-771, 894-907) ------------------------------
struct EditFields {
    float *U, *V, *nU, *nV, *P, *S, *M, *nM;
    // Fused path only: newM is kept current just where advection skips (solid cells), so
    // a cell that turns solid latches its smoke there -- what the reference's newM holds
    // (== M after the last advectSmoke, fluid.go:433) unless the caller edited that
    // cell's smoke since the last Simulate (documented deviation).
    int latch_smoke;
};

__device__ __forceinline__ void apply_edit_cell(const Grid &g, const EditFields &f, const fb_edit_cmd &c, int i, int j)
{
    if (i < g.i_alloc0 || i >= g.i_alloc0 + g.lines_alloc) return;   // not held by this rank
    if (i < 0 || i >= g.NX || j < 0 || j >= g.NY) return;
    const size_t a = g.at(i, j);
    const bool held_ip1 = (i + 1) < g.i_alloc0 + g.lines_alloc;
    switch (c.op) {
    case FB_EDIT_CIRCLE_OBSTACLE: {
        float dx = (float)(i - c.i0), dy = (float)(j - c.j0);
        float dx2 = dx * dx, dy2 = dy * dy;
        long long r = c.i1;
        if (!(dx2 + dy2 <= (float)(r * r))) return;
    }   // fallthrough: SetSolid(i, j, true)
    case FB_EDIT_SET_SOLID: {
        const bool solid = (c.op == FB_EDIT_CIRCLE_OBSTACLE) || (c.a != 0.0f);
        if (solid && f.latch_smoke && f.S[a] != 0.0f) f.nM[a] = f.M[a];
        f.S[a] = solid ? 0.0f : 1.0f;
        if (solid) {
            f.U[a] = 0.0f; f.V[a] = 0.0f; f.nU[a] = 0.0f; f.nV[a] = 0.0f;
            if (i + 1 < g.NX && held_ip1) { f.U[g.at(i + 1, j)] = 0.0f; f.nU[g.at(i + 1, j)] = 0.0f; }
            if (j + 1 < g.NY) { f.V[g.at(i, j + 1)] = 0.0f; f.nV[g.at(i, j + 1)] = 0.0f; }
        }
        break;
    }
    case FB_EDIT_SET_VELOCITY_IF_FLUID:
        if (f.S[a] == 0.0f) break;   // fallthrough
    case FB_EDIT_SET_VELOCITY: f.U[a] = c.a; f.V[a] = c.b; break;
    case FB_EDIT_ADD_SMOKE_IF_FLUID:
        if (f.S[a] == 0.0f) break;   // fallthrough
    case FB_EDIT_ADD_SMOKE: f.M[a] += c.a; break;
    case FB_EDIT_SET_SMOKE: f.M[a] = c.a; break;
    case FB_EDIT_APPLY_FORCE:
        if (i < 1 || i >= g.NX - 1 || j < 1 || j >= g.NY - 1) break;
        if (f.S[a] == 0.0f) break;
        f.U[a] += c.a; f.V[a] += c.b;
        break;
    case FB_EDIT_RESET:
        f.U[a] = 0.f; f.V[a] = 0.f; f.nU[a] = 0.f; f.nV[a] = 0.f; f.P[a] = 0.f; f.M[a] = 0.f; f.nM[a] = 0.f;
        break;
    default: break;
    }
}

__device__ __forceinline__ void edit_rect(const Grid &g, const fb_edit_cmd &c, int &i0, int &j0, int &ni, int &nj)
{
    if (c.op == FB_EDIT_CIRCLE_OBSTACLE) {
        i0 = c.i0 - c.i1; j0 = c.j0 - c.i1; ni = 2 * c.i1 + 1; nj = 2 * c.i1 + 1;
    } else if (c.op == FB_EDIT_RESET) {
        i0 = 0; j0 = 0; ni = g.NX; nj = g.NY;
    } else {
        i0 = c.i0; j0 = c.j0; ni = c.i1 - c.i0; nj = c.j1 - c.j0;
    }
    // clip to the domain and to the lines this rank holds
    int ia = max(i0, max(0, g.i_alloc0)), ibnd = min(i0 + ni, min(g.NX, g.i_alloc0 + g.lines_alloc));
    int ja = max(j0, 0), jb = min(j0 + nj, g.NY);
    i0 = ia; j0 = ja; ni = max(ibnd - ia, 0); nj = max(jb - ja, 0);
}

// A SetSolid(true) on cell (i,j) also zeroes faces stored at (i+1,j) and (i,j+1);
// a later command in the same list may touch those, so commands are separated
// by a barrier: one CTA walks the list in order.
__global__ void k_edits_seq(Grid g, EditFields f, const fb_edit_cmd *__restrict__ cmds, int n)
{
    for (int q = 0; q < n; q++) {
        const fb_edit_cmd c = cmds[q];
        int i0, j0, ni, nj;
        edit_rect(g, c, i0, j0, ni, nj);
        const long long cells = (long long)ni * nj;
        for (long long e = threadIdx.x; e < cells; e += blockDim.x) {
            const int i = i0 + (int)(e / nj), j = j0 + (int)(e % nj);
            apply_edit_cell(g, f, c, i, j);
        }
        __syncthreads();
    }
}

// One large command over the whole grid of threads.
__global__ void k_edit_one(Grid g, EditFields f, fb_edit_cmd c)
{
    int i0, j0, ni, nj;
    edit_rect(g, c, i0, j0, ni, nj);
    const int j = j0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= j0 + nj) return;
    for (int i = i0 + blockIdx.y; i < i0 + ni; i += gridDim.y) apply_edit_cell(g, f, c, i, j);
}

// ---- halo exchange through peer memory (NVLink P2P) --------------------------------------
// Every rank PACKS the `lines` owned lines next to each slab boundary of U, V, M into a send
// buffer of its own (exported to the neighbours by CUDA IPC), PUBLISHES the exchange epoch in a
// flag next to it, and PULLS its ghost lines straight out of the neighbours' send buffers.  The
// pull waits on the neighbour's flag, so there is no host hand-shake and no collective; two
// buffers alternate so that a rank may pack epoch k+2 only after its own pull of k+1, which the
// neighbour's pack of k+1 (after ITS pull of k) precedes.
struct HaloPack { const float *src[3]; float *dst; int i_lo, i_hi, lines, pitch, i_alloc0, field0, nf; };
__global__ void k_halo_pack(HaloPack a)
{
    // grid: (pitch/4 / 256, lines, 2 * nf): z = side * nf + (field - field0); the buffer keeps room for all three fields
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y, side = blockIdx.z / a.nf, field = a.field0 + blockIdx.z % a.nf;
    if (4 * c4 >= a.pitch) return;
    const int i = side == 0 ? a.i_lo + l : a.i_hi - a.lines + l;
    const float4 v = *reinterpret_cast<const float4 *>(a.src[field] + (size_t)(i - a.i_alloc0) * a.pitch + 4 * c4);
    *reinterpret_cast<float4 *>(a.dst + ((size_t)(side * 3 + field) * a.lines + l) * a.pitch + 4 * c4) = v;
}
__global__ void k_halo_publish(unsigned *flag, unsigned epoch)
{
    __threadfence_system();
    reinterpret_cast<volatile unsigned *>(flag)[0] = epoch;
    __threadfence_system();
}
struct HaloPull { float *dst[3]; const float *peer[2]; const unsigned *peer_flag[2]; int recv_i[2]; int lines, pitch, i_alloc0; unsigned epoch; int *bad;
                  int field0, nf; };
__global__ void k_halo_pull(HaloPull a)
{
    // grid: (pitch/4 / 256, lines, 2 * nf): z = side * nf + (field - field0); side s reads the neighbour's region 1 - s
    const int side = blockIdx.z / a.nf, field = a.field0 + blockIdx.z % a.nf;
    if (!a.peer[side]) return;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const volatile unsigned *f = a.peer_flag[side];
        int good = 0;
        for (unsigned spin = 0; spin < (1u << 26); spin++) {
            if ((int)(*f - a.epoch) >= 0) { good = 1; break; }
            __nanosleep(200);
        }
        if (!good) atomicExch(a.bad, 2);          // the neighbour never published: FB_ERR_HALO at the next check
        __threadfence_system();
        ok = good;
    }
    __syncthreads();
    if (!ok) return;
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y;
    if (4 * c4 >= a.pitch) return;
    const float4 v = __ldcv(reinterpret_cast<const float4 *>(a.peer[side] + ((size_t)((1 - side) * 3 + field) * a.lines + l) * a.pitch + 4 * c4));
    *reinterpret_cast<float4 *>(a.dst[field] + (size_t)(a.recv_i[side] + l - a.i_alloc0) * a.pitch + 4 * c4) = v;
}

// ---- the frame loop either side of Simulate ------------------------
// Draw's pixel pass (main/main.go:550-574, 620-652; main/colors.go:8-84) and advectParticles
// (main/main.go:512-546) on the device, so that a frame is pixels and particle positions instead of
// float fields plus two host round trips per particle.  float32 arithmetic as the Go code; the
// float -> integer conversions follow amd64 (truncate; "integer indefinite" for NaN / out of range).
__device__ __forceinline__ float key2f_dev(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned go_u8(float v)
{
    const int t = (v != v || v >= 2147483648.0f || v < -2147483648.0f) ? (int)0x80000000 : (int)v;
    return (unsigned)t & 0xffu;
}
__device__ __forceinline__ unsigned pack_rgba(float r, float g, float b)
{
    return go_u8(255.0f * r) | (go_u8(255.0f * g) << 8) | (go_u8(255.0f * b) << 16) | 0xff000000u;
}
__device__ __forceinline__ unsigned sci_color(float val, float minVal, float maxVal)   // colors.go:48-84
{
    val = go_minf(go_maxf(val, minVal), maxVal - 0.0001f);
    const float d = maxVal - minVal;
    if (d <= 0.0f) val = 0.5f;
    else { val = val - minVal; val = val / d; }
    const float m = 0.25f;
    const float num = floorf(val / m);
    const float t = num * m;
    const float s = (val - t) / m;
    float r = 0.0f, g = 0.0f, b = 0.0f;
    if (num == 0.0f) { g = s; b = 1.0f; }
    else if (num == 1.0f) { g = 1.0f; b = 1.0f - s; }
    else if (num == 2.0f) { r = s; g = 1.0f; }
    else if (num == 3.0f) { r = 1.0f; g = 1.0f - s; }
    return pack_rgba(r, g, b);
}
__device__ __forceinline__ unsigned diverging_color(float val, float absMax)            // colors.go:8-46
{
    if (absMax < 1e-8f) return 0xffffffffu;
    float t = val / absMax;
    if (t > 1.0f) t = 1.0f;
    if (t < -1.0f) t = -1.0f;
    float r, g, b;
    if (t >= 0.0f) { r = 1.0f; g = 1.0f - t; b = 1.0f - t; }
    else { const float a = -t; r = 1.0f - a; g = 1.0f - a; b = 1.0f; }
    return pack_rgba(r, g, b);
}
__global__ void k_minmax_set(unsigned *red, float lo, float hi) { red[0] = f2key(lo); red[1] = f2key(hi); }

// 32 x 32 tile transpose: the field is [i][j] with j contiguous, the image is [NumY-1-j][i] with i
// contiguous.  Block (32, 8).  `img` holds this rank's lines only: pixel (jj, i - ib), row length ie - ib.
template <int DIVERGING>
__global__ void __launch_bounds__(256)
k_render(Grid g, const float *__restrict__ A, const float *__restrict__ S, const unsigned *__restrict__ red,
         unsigned *__restrict__ img, int ib, int ie)
{
    __shared__ unsigned tile[32][33];
    const float lo = key2f_dev(red[0]), hi = key2f_dev(red[1]);
    const float absMax = fmaxf(fabsf(lo), fabsf(hi));          // math.Max(math.Abs, math.Abs): exact in float32
    const int i0 = ib + blockIdx.y * 32, j0 = blockIdx.x * 32;
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int i = i0 + threadIdx.y + r, j = j0 + threadIdx.x;
        unsigned c = 0;
        if (i < ie && j < g.NY) {
            const size_t a = g.at(i, j);
            c = DIVERGING ? diverging_color(A[a], absMax) : sci_color(A[a], lo, hi);
            if (S[a] == 0.0f) c = 0xff000000u;                  // solid cells black (main.go:564-574)
        }
        tile[threadIdx.y + r][threadIdx.x] = c;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int i = i0 + threadIdx.x, j = j0 + threadIdx.y + r;
        if (i < ie && j < g.NY) img[(size_t)(g.NY - 1 - j) * (ie - ib) + (i - ib)] = tile[threadIdx.x][threadIdx.y + r];
    }
}

struct fb_particle_dev { float x, y; unsigned rgbx; float age, max_age; };
__global__ void k_advect_particles(Grid g, const float *__restrict__ U, const float *__restrict__ V,
                                   const float *__restrict__ S, fb_particle_dev *__restrict__ ps, int *__restrict__ alive,
                                   size_t n, float dt, float h, int *bad)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    fb_particle_dev p = ps[t];
    const float h1 = 1.0f / h, h2 = 0.5f * h;
    int ok = 0;
    p.age = p.age + dt;
    if (!(p.age > p.max_age)) {
        const float u1 = sample_from<0>(g, U, p.x, p.y, h, h1, h2, bad);
        const float v1 = sample_from<1>(g, V, p.x, p.y, h, h1, h2, bad);
        const float hd = 0.5f * dt;
        const float mu = hd * u1, mv = hd * v1;
        const float midX = p.x + mu, midY = p.y + mv;
        const float u2 = sample_from<0>(g, U, midX, midY, h, h1, h2, bad);
        const float v2 = sample_from<1>(g, V, midX, midY, h, h1, h2, bad);
        const float du = dt * u2, dv = dt * v2;
        p.x = p.x + du;
        p.y = p.y + dv;
        // int(p.X / h): truncation toward zero, so (-1, 0) maps to cell 0; NaN / huge -> MinInt64 -> dropped
        const float qx = p.x / h, qy = p.y / h;
        if (qx > -1.0f && qx < (float)g.NX && qy > -1.0f && qy < (float)g.NY) {
            const int fi = (int)qx, fj = (int)qy;
            if (fi >= g.i_alloc0 && fi < g.i_alloc0 + g.lines_alloc) ok = S[g.at(fi, fj)] != 0.0f;
            else *bad = 1;
        }
    }
    ps[t] = p;
    alive[t] = ok;
}
