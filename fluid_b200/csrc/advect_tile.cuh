// advect_tile.cuh -- semi-Lagrangian velocity advection (fluid.go:291-398) and the BFECC back-trace + correct pass
// (fluid.go:938-987, 1094-1120) on shared-memory tiles.
//
// k_advect_velocity_full / k_bfecc_velocity_correct (advect_fused.cuh) gather their bilinear taps straight from global
// memory with 64-bit address arithmetic, and a thread owns 4 consecutive cells, so every tap is a 16-byte-strided warp
// access: 4 L1 wavefronts per tap instruction (ncu round 1: LSU data pipe 83 %, issue slots 81 %, ~180 / ~260
// instructions per cell, 0.50 / 0.47 of the HBM peak).  Here a CTA owns the tile of AT_TI lines x AT_TJ columns with the
// GLOBAL index (ty, tx) -- tile origins are multiples of (AT_TI, AT_TJ) whatever line range a launch covers:
//   * the sampled planes plus a halo of AT_R + 1 lines / 8 columns are staged in shared memory by TMA: ONE 2-D tensor-map
//     copy per plane (cp.async.bulk.tensor.2d, out-of-bounds elements zero-filled, completion counted on one mbarrier).
//     The first form of these kernels issued one 1-D bulk copy per line and plane (92 - 156 copies of 576 B per tile) and
//     40 - 50 % of its stall samples were warps waiting for the tile (ncu, round 2);
//   * ONE LANE PER CELL along j: the four taps of a warp are unit-stride LDS (one wavefront each), at 32-bit shared
//     addresses with immediate offsets (+1, +pitch, +pitch+1);
//   * a tap pair that lies inside the staged region needs neither the `min(x0+1, NumX-1)` collapse nor the
//     `min(floor, NumX-1)` clamp of fluid.go:373-377 (the region ends at NumX-1 / NumY-1 at the latest), and in a tile
//     whose staged region lies inside [2, NumX-3] x [2, NumY-3] the coordinate clamps of fluid.go:363-364 are provably
//     no-ops too (an in-region tap pair means h <= x <= NumX*h), so they are skipped;
//   * a per-tile flag built with the neighbour mask (k_tile_flags: every cell of the tile is an interior fluid cell
//     with fluid on its -x and -y side) selects a straight-line body without the active / ring / stale-scratch cases
//     of fluid.go:300-317 -- all of a preset's domain except walls and obstacles;
//   * one tile per CTA, four CTAs per SM: the tile load of one CTA hides behind the arithmetic of the others.  Measured and
//     NOT adopted: persistent CTAs (two per SM, 512 threads) with two buffers, the next tile streaming in while this one is
//     computed -- 0.127 / 0.217 ms against 0.097 / 0.154 for the two kernels at 4098^2: the block-wide barrier per tile
//     idles half of an SM's warps, whereas four independent CTAs never wait for each other;
//   * a trace that leaves the staged region (|dt*u| > AT_R cells, or the domain edge) falls back to the global
//     sampler sample_fast<> on the ORIGINAL coordinates, out of line: same result, slower, rare.
// The arithmetic per face is the reference's, operation for operation (same code as sample_fast).
#pragma once
#include <cuda.h>             // CUtensorMap
#include "advect_fused.cuh"
#include "rbq_fused.cuh"      // mbarrier helpers

#ifndef AT_TI
#define AT_TI 32              // lines per tile
#endif
#define AT_TJ 128             // columns per tile
#ifndef AT_R
#define AT_R 6                // a back-trace may land up to AT_R lines / columns away (dt*|u|/h < AT_R - 1)
#endif
#define AT_CH 8               // staged columns left of the tile (>= AT_R + 1, multiple of 4: 16-byte TMA granules)
#define AT_PW (AT_TJ + 2 * AT_CH)        // staged columns: 144
#define AT_TL (AT_TI + 2 * (AT_R + 1))   // staged lines of a sampled plane
#ifndef AT_THREADS
#define AT_THREADS 256        // two line groups x 128 columns; a thread walks AT_TI / 2 consecutive lines of its column
#endif
#ifndef AT_MINB
#define AT_MINB 4             // CTAs per SM the kernels are compiled for
#endif
#define AT_LG (AT_THREADS / 128)
#define AT_BUF (2 * AT_TL * AT_PW * 4)            // the tile buffer: U and V
#define AT_SMEM (AT_BUF + 16)
// BFECC correct: U, V with a one-line halo (trace velocities, 3x3 clamp) + fwdU, fwdV with the full halo
#ifndef AT_BTI
#define AT_BTI 16
#endif
#define AT_BTL (AT_BTI + 2 * (AT_R + 1))
#define AT_BVL (AT_BTI + 2)
#define AT_BBUF ((2 * AT_BTL + 2 * AT_BVL) * AT_PW * 4)
#define AT_BSMEM (AT_BBUF + 16)
static_assert(AT_CH >= AT_R + 1 && AT_CH % 4 == 0, "column halo");
static_assert(AT_TI % AT_BTI == 0, "a BFECC tile lies inside one flag tile");

__device__ __forceinline__ int cdiv_dev(int a, int b) { return (a + b - 1) / b; }

struct ATile {
    int ls0, cs0;             // global line / column of staged element (0, 0) of the SAMPLED planes
    int vl0, vc0;             // first line / column a tap pair may START on
    unsigned nl, nc;          // ... and how many: x0 in [vl0, vl0 + nl), x0 + 1 still staged (and <= NumX-1); y alike
};

// ---- per-tile flags, rebuilt with the mask (only when S changes) ---------------------------------
// 1: every cell of tile (ty, tx) lies in 1..NumX-2 x 1..NumY-2, is resident on this rank and has MK_C, MK_XM, MK_YM
__global__ void __launch_bounds__(256) k_tile_flags(const Grid g, const unsigned char *__restrict__ mask, unsigned char *__restrict__ flags,
                                                           const int ntx)
{
    const int tx = blockIdx.x, ty = blockIdx.y;
    const int i0 = ty * AT_TI, j = tx * AT_TJ + (threadIdx.x & 127);
    int ok = 1;
    for (int i = i0 + (threadIdx.x >> 7); i < i0 + AT_TI; i += 2) {
        const bool inside = i >= 1 && i <= g.NX - 2 && j >= 1 && j <= g.NY - 2 && i >= g.i_alloc0 && i < g.i_alloc0 + g.lines_alloc;
        if (!inside || (mask[g.at(i, j)] & (MK_C | MK_XM | MK_YM)) != (MK_C | MK_XM | MK_YM)) ok = 0;
    }
    ok = __syncthreads_and(ok);
    if (threadIdx.x == 0) flags[ty * ntx + tx] = (unsigned char)ok;
}

// the rare path, kept out of line so that the tile loop stays compact
template <int FLD, bool CHECK>
__device__ __noinline__ float sample_far(const AdvCtx &c, const float *__restrict__ gdata, const float x, const float y, int *bad)
{
    return sample_fast<FLD, CHECK>(c, gdata, x, y, bad);
}

// sampleField (fluid.go:357-398) with the taps in the staged tile `sm`; `gdata` is the same plane in global memory.
template <int FLD, bool INTERIOR, bool CHECK>
__device__ __forceinline__ float sample_tile(const AdvCtx &c, const ATile &T, const float *__restrict__ sm, const float *__restrict__ gdata,
                                             const float x, const float y, int *bad)
{
    float xc = x, yc = y;
    if (!INTERIOR) {
        xc = fmaxf(fminf(x, c.xmax), c.h);
        yc = fmaxf(fminf(y, c.ymax), c.h);
    }
    const float xs = (FLD == 0) ? xc : xc - c.h2;
    const float ys = (FLD == 1) ? yc : yc - c.h2;
    const float2 s2 = make_float2(xs, ys), h1 = make_float2(c.h1, c.h1);
    const float2 q = __fmul2_rn(s2, h1);
    // floor as ONE conversion (F2I.FLOOR); inside the staged region |q| is small, so float(x0) IS floorf(q.x)
    const int x0 = __float2int_rd(q.x), y0 = __float2int_rd(q.y);
    if ((unsigned)(x0 - T.vl0) < T.nl && (unsigned)(y0 - T.vc0) < T.nc) {
        const float fx = (float)x0, fy = (float)y0;
        const float2 f0h = __fmul2_rn(make_float2(fx, fy), make_float2(c.h, c.h));
        const float2 t = __fmul2_rn(__fadd2_rn(s2, make_float2(-f0h.x, -f0h.y)), h1);
        const float2 sxy = __fadd2_rn(make_float2(1.0f, 1.0f), make_float2(-t.x, -t.y));
        const float *p = sm + (x0 - T.ls0) * AT_PW + (y0 - T.cs0);
        const float f00 = p[0], f10 = p[AT_PW], f11 = p[AT_PW + 1], f01 = p[1];
        // the four weights as scalar products (the pairs (sx, tx) a packed product needs would have to be assembled
        // with moves), the four weighted taps as two packed products
        const float w00 = sxy.x * sxy.y, w10 = t.x * sxy.y, w11 = t.x * t.y, w01 = sxy.x * t.y;
        const float2 pA = __fmul2_rn(make_float2(w00, w10), make_float2(f00, f10));
        const float2 pB = __fmul2_rn(make_float2(w01, w11), make_float2(f01, f11));
        return ((pA.x + pA.y) + pB.y) + pB.x;
    }
    return sample_far<FLD, CHECK>(c, gdata, x, y, bad);
}

// The same sample split in two for the straight-line bodies: at_tap() computes the interpolation from the staged tile
// UNCONDITIONALLY (a tap pair outside the staged region reads element 0 instead and its value is discarded) and says
// whether the pair was inside; the caller re-samples the outsiders through sample_far afterwards.  Several samples then
// sit in ONE basic block and their dependent chains interleave (with the branch inside every sample a warp issued one
// instruction every ~10 cycles, ncu round 2).  Interior tiles only (no coordinate clamps).
template <int FLD>
__device__ __forceinline__ float at_tap(const AdvCtx &c, const ATile &T, const float *__restrict__ sm, const float x, const float y, bool &inside)
{
    const float xs = (FLD == 0) ? x : x - c.h2;
    const float ys = (FLD == 1) ? y : y - c.h2;
    const float2 s2 = make_float2(xs, ys), h1 = make_float2(c.h1, c.h1);
    const float2 q = __fmul2_rn(s2, h1);
    const int x0 = __float2int_rd(q.x), y0 = __float2int_rd(q.y);
    inside = (unsigned)(x0 - T.vl0) < T.nl && (unsigned)(y0 - T.vc0) < T.nc;
    const float fx = (float)x0, fy = (float)y0;
    const float2 f0h = __fmul2_rn(make_float2(fx, fy), make_float2(c.h, c.h));
    const float2 t = __fmul2_rn(__fadd2_rn(s2, make_float2(-f0h.x, -f0h.y)), h1);
    const float2 sxy = __fadd2_rn(make_float2(1.0f, 1.0f), make_float2(-t.x, -t.y));
    const int off = inside ? (x0 - T.ls0) * AT_PW + (y0 - T.cs0) : 0;
    const float *p = sm + off;
    const float f00 = p[0], f10 = p[AT_PW], f11 = p[AT_PW + 1], f01 = p[1];
    const float w00 = sxy.x * sxy.y, w10 = t.x * sxy.y, w11 = t.x * t.y, w01 = sxy.x * t.y;
    const float2 pA = __fmul2_rn(make_float2(w00, w10), make_float2(f00, f10));
    const float2 pB = __fmul2_rn(make_float2(w01, w11), make_float2(f01, f11));
    return ((pA.x + pA.y) + pB.y) + pB.x;
}

// One 2-D tensor-map copy: the box of the map (AT_PW columns x its line count) whose first element is (line ls0, column cs0)
// of the plane -> dst; coordinates are relative to the plane's first allocated line and may lie outside it (zero fill).
__device__ __forceinline__ void at_tensor_load(float *dst, const CUtensorMap *tm, const int col, const int line, const unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(rq_s32(dst)), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(col), "r"(line), "r"(bar) : "memory");
}

__device__ __forceinline__ void at_tile_geometry(const AdvCtx &c, ATile &T, const int nlines)
{
    // lines / columns that are staged AND exist: a tap pair may start on [v0, v1 - 1]
    const int la = max(T.ls0, max(c.i_alloc0, 0)), lb = min(T.ls0 + nlines, min(c.i_alloc0 + c.lines_alloc, c.NX));
    const int ca = max(T.cs0, 0), cb = min(T.cs0 + AT_PW, c.NY);
    T.vl0 = la; T.nl = (unsigned)max(lb - 1 - la, 0);
    T.vc0 = ca; T.nc = (unsigned)max(cb - 1 - ca, 0);
}

__device__ __forceinline__ void at_wait_tiles(unsigned long long *bar, const unsigned parity, int *bad)
{
    // every thread waits for the tile; try_wait carries a suspend-time hint, so a waiting warp sleeps instead of
    // polling (without the hint the polls were 22 % of the kernel's issued instructions).  Bounded: a copy that
    // never lands latches the error flag instead of hanging.
    const unsigned b = rq_s32(bar);
    bool ok = false;
#pragma unroll 1
    for (int k = 0; k < (1 << 17) && !ok; k++) ok = rq_mbar_try_a(b, parity);
    if (!ok && bad) *bad = 3;
}

// FAST: the tile is interior (no coordinate clamps) and all its cells are active (flag of k_tile_flags): every face is
// traced, no ring, no stale-scratch fall-back.
template <bool FAST, bool CHECK>
__device__ __forceinline__ void at_velocity_cells(const AdvCtx &c, const ATile &T, const float *__restrict__ sU, const float *__restrict__ sV,
                                                  const float *__restrict__ trU, const float *__restrict__ trV,
                                                  const unsigned char *__restrict__ mask, const float *__restrict__ shU,
                                                  const float *__restrict__ shV, float *__restrict__ dstU, float *__restrict__ dstV,
                                                  const float dt, const int tl0, const int i0, const int i1, const int j, int *bad)
{
    if (j >= c.NY) return;
    const float yj = (float)j * c.h;
    const float yj2 = yj + c.h2;
    // a thread walks CONSECUTIVE lines of its column: line group g of the CTA takes lines tl0 + g * half .. of the tile.
    // Of the eight values the two averages need, four are the previous line's (carried), four are new
    const int half = AT_TI / AT_LG;
    const int ia = max(tl0 + (int)(threadIdx.x >> 7) * half, i0), ib = min(tl0 + ((int)(threadIdx.x >> 7) + 1) * half, i1);
    if (ia >= ib) return;
    const float *pu = sU + (ia - T.ls0) * AT_PW + (j - T.cs0), *pv = sV + (ia - T.ls0) * AT_PW + (j - T.cs0);
    size_t o = (size_t)(ia - c.i_alloc0) * c.pitch + j;
    float um = pu[-1], u = pu[0];                  // U[i, j-1], U[i, j]
    float vm0 = pv[-AT_PW], vm1 = pv[-AT_PW + 1];  // V[i-1, j], V[i-1, j+1]
    if (FAST) {
        // straight-line body, two lines (four samples) per trip: every face is traced, nothing is selected
        const size_t P = (size_t)c.pitch;
        int i = ia;
        for (; i + 1 < ib; i += 2, pu += 2 * AT_PW, pv += 2 * AT_PW, o += 2 * P) {
            const float upm = pu[AT_PW - 1], up = pu[AT_PW], upm2 = pu[2 * AT_PW - 1], up2 = pu[2 * AT_PW];
            const float v = pv[0], vn = pv[1], v2 = pv[AT_PW], vn2 = pv[AT_PW + 1];
            const float xi = (float)i * c.h, xi2 = (float)(i + 1) * c.h;
            const float av = (((vm0 + v) + vm1) + vn) * 0.25f, au = (((um + u) + upm) + up) * 0.25f;
            const float av2 = (((v + v2) + vn) + vn2) * 0.25f, au2 = (((upm + up) + upm2) + up2) * 0.25f;
            const float xa = xi - dt * u, ya = yj2 - dt * av, xb = (xi + c.h2) - dt * au, yb = yj - dt * v;
            const float xc = xi2 - dt * up, yc = yj2 - dt * av2, xd = (xi2 + c.h2) - dt * au2, yd = yj - dt * v2;
            bool ina, inb, inc, ind;
            float oa = at_tap<0>(c, T, sU, xa, ya, ina), ob = at_tap<1>(c, T, sV, xb, yb, inb);
            float oc = at_tap<0>(c, T, sU, xc, yc, inc), od = at_tap<1>(c, T, sV, xd, yd, ind);
            if (!(ina && inb && inc && ind)) {          // a trace left the staged region: the global sampler, same result
                if (!ina) oa = sample_far<0, CHECK>(c, trU, xa, ya, bad);
                if (!inb) ob = sample_far<1, CHECK>(c, trV, xb, yb, bad);
                if (!inc) oc = sample_far<0, CHECK>(c, trU, xc, yc, bad);
                if (!ind) od = sample_far<1, CHECK>(c, trV, xd, yd, bad);
            }
            dstU[o] = oa; dstV[o] = ob; dstU[o + P] = oc; dstV[o + P] = od;
            um = upm2; u = up2; vm0 = v2; vm1 = vn2;
        }
        if (i < ib) {
            const float upm = pu[AT_PW - 1], up = pu[AT_PW];
            const float v = pv[0], vn = pv[1];
            const float xi = (float)i * c.h;
            const float av = (((vm0 + v) + vm1) + vn) * 0.25f, au = (((um + u) + upm) + up) * 0.25f;
            const float du = dt * u, dv = dt * av, du2 = dt * au, dv2 = dt * v;
            dstU[o] = sample_tile<0, true, CHECK>(c, T, sU, trU, xi - du, yj2 - dv, bad);
            dstV[o] = sample_tile<1, true, CHECK>(c, T, sV, trV, (xi + c.h2) - du2, yj - dv2, bad);
        }
        return;
    }
#pragma unroll 2
    for (int i = ia; i < ib; i++, pu += AT_PW, pv += AT_PW, o += c.pitch) {
        const float upm = pu[AT_PW - 1], up = pu[AT_PW];      // U[i+1, j-1], U[i+1, j]
        const float v = pv[0], vn = pv[1];                    // V[i, j], V[i, j+1]
        const float xi = (float)i * c.h;
        float outU, outV;
        const unsigned m = mask[o];
        const bool in_loop = i >= 1 && j >= 1;                       // loops start at 1 (fluid.go:300-301); j < NumY holds
        const bool act_u = in_loop && (m & MK_C) && (m & MK_XM) && j < c.NY - 1;
        const bool act_v = in_loop && (m & MK_C) && (m & MK_YM) && i < c.NX - 1;
        const bool ring = i == 0 || j == 0 || i == c.NX - 1 || j == c.NY - 1;
        if (act_u) {
            // avgV (fluid.go:342-347): V[i-1,j] + V[i,j] + V[i-1,j+1] + V[i,j+1]
            const float av = (((vm0 + v) + vm1) + vn) * 0.25f;
            const float du = dt * u, dv = dt * av;
            outU = sample_tile<0, false, CHECK>(c, T, sU, trU, xi - du, yj2 - dv, bad);
        } else {
            outU = ring ? u : shU[o];
        }
        if (act_v) {
            // avgU (fluid.go:335-340): U[i,j-1] + U[i,j] + U[i+1,j-1] + U[i+1,j]
            const float au = (((um + u) + upm) + up) * 0.25f;
            const float du = dt * au, dv = dt * v;
            outV = sample_tile<1, false, CHECK>(c, T, sV, trV, (xi + c.h2) - du, yj - dv, bad);
        } else {
            outV = ring ? v : shV[o];
        }
        dstU[o] = outU;
        dstV[o] = outV;
        um = upm; u = up; vm0 = v; vm1 = vn;
    }
}

// advectVelocity writing complete planes: same contract as k_advect_velocity_full (tr*: the planes that are traced
// through AND sampled; sh*: the stale scratch values skipped faces fall back to, Q-6).  blockIdx.y counts line tiles from
// the one that holds line ib.
template <bool CHECK>
__global__ void __launch_bounds__(AT_THREADS, AT_MINB)
k_advect_velocity_tile(const AdvCtx c, const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
                       const float *__restrict__ trU, const float *__restrict__ trV,
                       const unsigned char *__restrict__ mask, const unsigned char *__restrict__ tile_flags, const int ntx,
                       const float *__restrict__ shU, const float *__restrict__ shV, float *__restrict__ dstU,
                       float *__restrict__ dstV, const float dt, const int ib, const int ie, int *bad)
{
    extern __shared__ __align__(128) unsigned char at_smem[];
    float *sU = reinterpret_cast<float *>(at_smem), *sV = sU + AT_TL * AT_PW;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(at_smem + AT_BUF);
    const int tid = threadIdx.x;
    const int ty = ib / AT_TI + blockIdx.y, tx = blockIdx.x;
    const int i0 = max(ty * AT_TI, ib), i1 = min((ty + 1) * AT_TI, ie);
    ATile T;
    T.ls0 = ty * AT_TI - (AT_R + 1); T.cs0 = tx * AT_TJ - AT_CH;
    at_tile_geometry(c, T, AT_TL);
    if (tid == 0) rq_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const unsigned b = rq_s32(bar);
        rq_mbar_expect_tx(bar, 2u * AT_TL * AT_PW * 4u);          // the whole boxes count, zero-filled parts included
        at_tensor_load(sU, &tmU, T.cs0, T.ls0 - c.i_alloc0, b);
        at_tensor_load(sV, &tmV, T.cs0, T.ls0 - c.i_alloc0, b);
    }
    // interior tile: everything staged lies in [2, NumX-3] x [2, NumY-3] (and is resident) -> the coordinate clamps
    // cannot trigger for in-region taps; all cells active: flag of k_tile_flags
    const bool fast = T.ls0 >= 2 && T.ls0 + AT_TL <= c.NX - 2 && T.cs0 >= 2 && T.cs0 + AT_PW <= c.NY - 2 &&
                      T.ls0 >= c.i_alloc0 && T.ls0 + AT_TL <= c.i_alloc0 + c.lines_alloc && tile_flags[ty * ntx + tx] != 0;
    at_wait_tiles(bar, 0, bad);
    const int j = tx * AT_TJ + (tid & 127);
    if (fast) at_velocity_cells<true, CHECK>(c, T, sU, sV, trU, trV, mask, shU, shV, dstU, dstV, dt, ty * AT_TI, i0, i1, j, bad);
    else at_velocity_cells<false, CHECK>(c, T, sU, sV, trU, trV, mask, shU, shV, dstU, dstV, dt, ty * AT_TI, i0, i1, j, bad);
}

// ---- BFECC velocity: back-trace (+dt, sampling the forward result), error compensation and clamp to the 3x3
// neighbourhood of the original field in one pass (fluid.go:938-987, 1094-1120); contract of k_bfecc_velocity_correct.
template <bool FAST, bool CHECK>
__device__ __forceinline__ void at_bfecc_cells(const AdvCtx &c, const ATile &T, const float *__restrict__ sU, const float *__restrict__ sV,
                                               const float *__restrict__ sFU, const float *__restrict__ sFV, const int vls0,
                                               const float *__restrict__ fwdU, const float *__restrict__ fwdV,
                                               const unsigned char *__restrict__ mask, float *__restrict__ corrU,
                                               float *__restrict__ corrV, const float dt, const int tl0, const int i0, const int i1, const int j, int *bad)
{
    if (j >= c.NY) return;
    const float yj = (float)j * c.h;
    const float yj2 = yj + c.h2;
    // consecutive lines per thread: the 3x3 windows of U and V slide down one line per cell, two rows are carried
    const int half = AT_BTI / AT_LG;
    const int ia = max(tl0 + (int)(threadIdx.x >> 7) * half, i0), ib = min(tl0 + ((int)(threadIdx.x >> 7) + 1) * half, i1);
    if (ia >= ib) return;
    const float *pu = sU + (ia - vls0) * AT_PW + (j - T.cs0), *pv = sV + (ia - vls0) * AT_PW + (j - T.cs0);
    size_t o = (size_t)(ia - c.i_alloc0) * c.pitch + j;
    float ua[3] = { pu[-AT_PW - 1], pu[-AT_PW], pu[-AT_PW + 1] }, ub[3] = { pu[-1], pu[0], pu[1] };      // rows i-1, i of U
    float va[3] = { pv[-AT_PW - 1], pv[-AT_PW], pv[-AT_PW + 1] }, vb[3] = { pv[-1], pv[0], pv[1] };
#pragma unroll 2
    for (int i = ia; i < ib; i++, pu += AT_PW, pv += AT_PW, o += c.pitch) {
        const float uc[3] = { pu[AT_PW - 1], pu[AT_PW], pu[AT_PW + 1] };                                  // row i+1
        const float vc[3] = { pv[AT_PW - 1], pv[AT_PW], pv[AT_PW + 1] };
        const float u = ub[1], v = vb[1];
        float cu = u, cv = v;
        // copy(corrU, origU) leaves the ring alone; correction and clamp run over ALL interior indices (Q-10)
        if (FAST || (i >= 1 && i <= c.NX - 2 && j >= 1 && j <= c.NY - 2)) {
            const float xi = (float)i * c.h;
            float bwdU = 0.0f, bwdV = 0.0f;                       // bwd arrays start as zeros (fluid.go:938-939)
            if (FAST) {
                // every face is traced; both samples unconditionally from the tile, outsiders re-sampled afterwards
                const float av = (((va[1] + v) + va[2]) + vb[2]) * 0.25f;          // V[i-1,j] + V[i,j] + V[i-1,j+1] + V[i,j+1]
                const float au = (((ub[0] + u) + uc[0]) + uc[1]) * 0.25f;          // U[i,j-1] + U[i,j] + U[i+1,j-1] + U[i+1,j]
                const float xa = xi + dt * u, ya = yj2 + dt * av, xb = (xi + c.h2) + dt * au, yb = yj + dt * v;
                bool ina, inb;
                bwdU = at_tap<0>(c, T, sFU, xa, ya, ina);
                bwdV = at_tap<1>(c, T, sFV, xb, yb, inb);
                if (!(ina && inb)) {
                    if (!ina) bwdU = sample_far<0, CHECK>(c, fwdU, xa, ya, bad);
                    if (!inb) bwdV = sample_far<1, CHECK>(c, fwdV, xb, yb, bad);
                }
            } else {
                const unsigned m = mask[o];
                if ((m & MK_C) && (m & MK_XM)) {
                    const float av = (((va[1] + v) + va[2]) + vb[2]) * 0.25f;
                    const float du = dt * u, dv = dt * av;
                    bwdU = sample_tile<0, false, CHECK>(c, T, sFU, fwdU, xi + du, yj2 + dv, bad);
                }
                if ((m & MK_C) && (m & MK_YM)) {
                    const float au = (((ub[0] + u) + uc[0]) + uc[1]) * 0.25f;
                    const float du = dt * au, dv = dt * v;
                    bwdV = sample_tile<1, false, CHECK>(c, T, sFV, fwdV, (xi + c.h2) + du, yj + dv, bad);
                }
            }
            // clampToNeighbors (fluid.go:1094-1120): min / max over the 3x3 neighbourhood of the original field
            float loU = fminf(fminf(ua[0], ua[1]), ua[2]), hiU = fmaxf(fmaxf(ua[0], ua[1]), ua[2]);
            float loV = fminf(fminf(va[0], va[1]), va[2]), hiV = fmaxf(fmaxf(va[0], va[1]), va[2]);
#pragma unroll
            for (int q = 0; q < 3; q++) {
                loU = fminf(fminf(loU, ub[q]), uc[q]); hiU = fmaxf(fmaxf(hiU, ub[q]), uc[q]);
                loV = fminf(fminf(loV, vb[q]), vc[q]); hiV = fmaxf(fmaxf(hiV, vb[q]), vc[q]);
            }
            const float eu = (bwdU - u) * 0.5f;
            const float ev = (bwdV - v) * 0.5f;
            cu = u - eu; cv = v - ev;
            cu = cu < loU ? loU : (cu > hiU ? hiU : cu);
            cv = cv < loV ? loV : (cv > hiV ? hiV : cv);
        }
        corrU[o] = cu;
        corrV[o] = cv;
#pragma unroll
        for (int q = 0; q < 3; q++) { ua[q] = ub[q]; ub[q] = uc[q]; va[q] = vb[q]; vb[q] = vc[q]; }
    }
}

template <bool CHECK>
__global__ void __launch_bounds__(AT_THREADS, AT_MINB)
k_bfecc_velocity_tile(const AdvCtx c, const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ CUtensorMap tmFU, const __grid_constant__ CUtensorMap tmFV,
                      const float *__restrict__ U, const float *__restrict__ V,
                      const unsigned char *__restrict__ mask, const unsigned char *__restrict__ tile_flags, const int ntx,
                      const float *__restrict__ fwdU, const float *__restrict__ fwdV, float *__restrict__ corrU,
                      float *__restrict__ corrV, const float dt, const int ib, const int ie, int *bad)
{
    extern __shared__ __align__(128) unsigned char at_smem[];
    float *sFU = reinterpret_cast<float *>(at_smem), *sFV = sFU + AT_BTL * AT_PW;
    float *sU = sFV + AT_BTL * AT_PW, *sV = sU + AT_BVL * AT_PW;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(at_smem + AT_BBUF);
    const int tid = threadIdx.x;
    const int ty = ib / AT_BTI + blockIdx.y, tx = blockIdx.x;
    const int i0 = max(ty * AT_BTI, ib), i1 = min((ty + 1) * AT_BTI, ie);
    ATile T;
    T.ls0 = ty * AT_BTI - (AT_R + 1); T.cs0 = tx * AT_TJ - AT_CH;
    at_tile_geometry(c, T, AT_BTL);
    const int vls0 = ty * AT_BTI - 1;
    if (tid == 0) rq_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const unsigned b = rq_s32(bar);
        rq_mbar_expect_tx(bar, (2u * AT_BTL + 2u * AT_BVL) * AT_PW * 4u);
        at_tensor_load(sFU, &tmFU, T.cs0, T.ls0 - c.i_alloc0, b);
        at_tensor_load(sFV, &tmFV, T.cs0, T.ls0 - c.i_alloc0, b);
        at_tensor_load(sU, &tmU, T.cs0, vls0 - c.i_alloc0, b);
        at_tensor_load(sV, &tmV, T.cs0, vls0 - c.i_alloc0, b);
    }
    // the per-tile flag refers to tiles of AT_TI lines: the flag of the AT_TI tile that contains this AT_BTI tile
    const bool fast = T.ls0 >= 2 && T.ls0 + AT_BTL <= c.NX - 2 && T.cs0 >= 2 && T.cs0 + AT_PW <= c.NY - 2 &&
                      T.ls0 >= c.i_alloc0 && T.ls0 + AT_BTL <= c.i_alloc0 + c.lines_alloc &&
                      tile_flags[((ty * AT_BTI) / AT_TI) * ntx + tx] != 0;
    at_wait_tiles(bar, 0, bad);
    const int j = tx * AT_TJ + (tid & 127);
    if (fast) at_bfecc_cells<true, CHECK>(c, T, sU, sV, sFU, sFV, vls0, fwdU, fwdV, mask, corrU, corrV, dt, ty * AT_BTI, i0, i1, j, bad);
    else at_bfecc_cells<false, CHECK>(c, T, sU, sV, sFU, sFV, vls0, fwdU, fwdV, mask, corrU, corrV, dt, ty * AT_BTI, i0, i1, j, bad);
}

// ======================= smoke: advectSmoke (fluid.go:400-434) and the BFECC correct pass (fluid.go:1013-1046) ==========
// Same structure: the sampled plane (M / fwdM) with the full halo, U and V (and origM for the 3x3 clamp) with a one-line
// halo, all through tensor-map TMA; a lane per cell, consecutive lines per thread, straight-line body on all-fluid tiles.
#ifndef AT_STI
#define AT_STI 16
#endif
#define AT_STL (AT_STI + 2 * (AT_R + 1))
#define AT_SVL (AT_STI + 2)
#define AT_SSMEM ((AT_STL + 2 * AT_SVL) * AT_PW * 4 + 16)
#define AT_SBSMEM ((AT_STL + 3 * AT_SVL) * AT_PW * 4 + 16)
static_assert(AT_TI % AT_STI == 0, "a smoke tile lies inside one flag tile");

template <bool FAST, bool CHECK>
__device__ __forceinline__ void at_smoke_cells(const AdvCtx &c, const ATile &T, const float *__restrict__ sM, const float *__restrict__ sU,
                                               const float *__restrict__ sV, const int vls0, const float *__restrict__ M,
                                               const unsigned char *__restrict__ mask, const float *__restrict__ shM,
                                               float *__restrict__ dst, const float dt, const float sa, const int tl0, const int i0,
                                               const int i1, const int j, int *bad)
{
    if (j >= c.NY) return;
    const float y0 = (float)j * c.h + c.h2;
    const int half = AT_STI / AT_LG;
    const int ia = max(tl0 + (int)(threadIdx.x >> 7) * half, i0), ib = min(tl0 + ((int)(threadIdx.x >> 7) + 1) * half, i1);
    if (ia >= ib) return;
    const float *pu = sU + (ia - vls0) * AT_PW + (j - T.cs0), *pv = sV + (ia - vls0) * AT_PW + (j - T.cs0);
    const float *pm = sM + (ia - T.ls0) * AT_PW + (j - T.cs0);
    size_t o = (size_t)(ia - c.i_alloc0) * c.pitch + j;
    const size_t P = (size_t)c.pitch;
    float u = pu[0];                                   // U[i, j], carried down the lines
    if (FAST) {
        int i = ia;
        for (; i + 1 < ib; i += 2, pu += 2 * AT_PW, pv += 2 * AT_PW, o += 2 * P) {
            const float up = pu[AT_PW], up2 = pu[2 * AT_PW];
            const float v = pv[0], vn = pv[1], v2 = pv[AT_PW], vn2 = pv[AT_PW + 1];
            const float uu = ((u + up) * 0.5f) * sa, vv = ((v + vn) * 0.5f) * sa;
            const float uu2 = ((up + up2) * 0.5f) * sa, vv2 = ((v2 + vn2) * 0.5f) * sa;
            const float xa = ((float)i * c.h + c.h2) - dt * uu, ya = y0 - dt * vv;
            const float xb = ((float)(i + 1) * c.h + c.h2) - dt * uu2, yb = y0 - dt * vv2;
            bool ina, inb;
            float a = at_tap<2>(c, T, sM, xa, ya, ina), b = at_tap<2>(c, T, sM, xb, yb, inb);
            if (!(ina && inb)) {
                if (!ina) a = sample_far<2, CHECK>(c, M, xa, ya, bad);
                if (!inb) b = sample_far<2, CHECK>(c, M, xb, yb, bad);
            }
            dst[o] = go_maxf(a, 0.0f);
            dst[o + P] = go_maxf(b, 0.0f);
            u = up2;
        }
        if (i < ib) {
            const float up = pu[AT_PW], v = pv[0], vn = pv[1];
            const float uu = ((u + up) * 0.5f) * sa, vv = ((v + vn) * 0.5f) * sa;
            const float du = dt * uu, dv = dt * vv;
            dst[o] = go_maxf(sample_tile<2, true, CHECK>(c, T, sM, M, ((float)i * c.h + c.h2) - du, y0 - dv, bad), 0.0f);
        }
        return;
    }
    for (int i = ia; i < ib; i++, pu += AT_PW, pv += AT_PW, pm += AT_PW, o += P) {
        const float up = pu[AT_PW];
        const float mm = pm[0];
        float out = mm;                                                       // copyBorder(newM, M) on the ring
        if (i >= 1 && i <= c.NX - 2 && j >= 1 && j <= c.NY - 2) {
            if (!(mask[o] & MK_C)) {
                out = shM[o];                                                 // solid: the stale scratch value (Q-6)
            } else {
                const float uu = ((u + up) * 0.5f) * sa, vv = ((pv[0] + pv[1]) * 0.5f) * sa;
                const float du = dt * uu, dv = dt * vv;
                out = go_maxf(sample_tile<2, false, CHECK>(c, T, sM, M, ((float)i * c.h + c.h2) - du, y0 - dv, bad), 0.0f);
            }
        }
        dst[o] = out;
        u = up;
    }
}

// contract of k_advect_smoke_full without the diffusion term (the host keeps that kernel for viscosityDiffusion > 0)
template <bool CHECK>
__global__ void __launch_bounds__(AT_THREADS, 4)
k_advect_smoke_tile(const AdvCtx c, const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmM, const float *__restrict__ M, const unsigned char *__restrict__ mask,
                    const unsigned char *__restrict__ tile_flags, const int ntx, const float *__restrict__ shM, float *__restrict__ dst,
                    const float dt, const float sa, const int ib, const int ie, int *bad)
{
    extern __shared__ __align__(128) unsigned char at_smem[];
    float *sM = reinterpret_cast<float *>(at_smem), *sU = sM + AT_STL * AT_PW, *sV = sU + AT_SVL * AT_PW;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sV + AT_SVL * AT_PW);
    const int tid = threadIdx.x;
    const int ty = ib / AT_STI + blockIdx.y, tx = blockIdx.x;
    const int i0 = max(ty * AT_STI, ib), i1 = min((ty + 1) * AT_STI, ie);
    ATile T;
    T.ls0 = ty * AT_STI - (AT_R + 1); T.cs0 = tx * AT_TJ - AT_CH;
    at_tile_geometry(c, T, AT_STL);
    const int vls0 = ty * AT_STI - 1;
    if (tid == 0) rq_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const unsigned b = rq_s32(bar);
        rq_mbar_expect_tx(bar, (AT_STL + 2u * AT_SVL) * AT_PW * 4u);
        at_tensor_load(sM, &tmM, T.cs0, T.ls0 - c.i_alloc0, b);
        at_tensor_load(sU, &tmU, T.cs0, vls0 - c.i_alloc0, b);
        at_tensor_load(sV, &tmV, T.cs0, vls0 - c.i_alloc0, b);
    }
    const bool fast = T.ls0 >= 2 && T.ls0 + AT_STL <= c.NX - 2 && T.cs0 >= 2 && T.cs0 + AT_PW <= c.NY - 2 &&
                      T.ls0 >= c.i_alloc0 && T.ls0 + AT_STL <= c.i_alloc0 + c.lines_alloc &&
                      tile_flags[((ty * AT_STI) / AT_TI) * ntx + tx] != 0;
    at_wait_tiles(bar, 0, bad);
    const int j = tx * AT_TJ + (tid & 127);
    if (fast) at_smoke_cells<true, CHECK>(c, T, sM, sU, sV, vls0, M, mask, shM, dst, dt, sa, ty * AT_STI, i0, i1, j, bad);
    else at_smoke_cells<false, CHECK>(c, T, sM, sU, sV, vls0, M, mask, shM, dst, dt, sa, ty * AT_STI, i0, i1, j, bad);
}

template <bool FAST, bool CHECK>
__device__ __forceinline__ void at_bfecc_smoke_cells(const AdvCtx &c, const ATile &T, const float *__restrict__ sF, const float *__restrict__ sO,
                                                     const float *__restrict__ sU, const float *__restrict__ sV, const int vls0,
                                                     const float *__restrict__ fwdM, const unsigned char *__restrict__ mask,
                                                     float *__restrict__ corrM, const float dt, const float sa, const int tl0,
                                                     const int i0, const int i1, const int j, int *bad)
{
    if (j >= c.NY) return;
    const float y0 = (float)j * c.h + c.h2;
    const int half = AT_STI / AT_LG;
    const int ia = max(tl0 + (int)(threadIdx.x >> 7) * half, i0), ib = min(tl0 + ((int)(threadIdx.x >> 7) + 1) * half, i1);
    if (ia >= ib) return;
    const int off = (ia - vls0) * AT_PW + (j - T.cs0);
    const float *pu = sU + off, *pv = sV + off, *po = sO + off;
    size_t o = (size_t)(ia - c.i_alloc0) * c.pitch + j;
    float u = pu[0];
    float ma[3] = { po[-AT_PW - 1], po[-AT_PW], po[-AT_PW + 1] }, mb[3] = { po[-1], po[0], po[1] };      // rows i-1, i of origM
    for (int i = ia; i < ib; i++, pu += AT_PW, pv += AT_PW, po += AT_PW, o += c.pitch) {
        const float mc[3] = { po[AT_PW - 1], po[AT_PW], po[AT_PW + 1] };                                  // row i+1
        const float up = pu[AT_PW];
        const float om = mb[1];
        float out = om;                                                       // copy(corrM, origM) leaves the ring alone
        if (FAST || (i >= 1 && i <= c.NX - 2 && j >= 1 && j <= c.NY - 2)) {
            float bwd = 0.0f;                                                 // bwdM starts as zeros (fluid.go:1013)
            const float uu = ((u + up) * 0.5f) * sa, vv = ((pv[0] + pv[1]) * 0.5f) * sa;
            const float du = dt * uu, dv = dt * vv;
            const float x = ((float)i * c.h + c.h2) + du, y = y0 + dv;
            if (FAST) {
                bool in;
                bwd = at_tap<2>(c, T, sF, x, y, in);
                if (!in) bwd = sample_far<2, CHECK>(c, fwdM, x, y, bad);
            } else if (mask[o] & MK_C) {
                bwd = sample_tile<2, false, CHECK>(c, T, sF, fwdM, x, y, bad);
            }
            float lo = fminf(fminf(ma[0], ma[1]), ma[2]), hi = fmaxf(fmaxf(ma[0], ma[1]), ma[2]);
#pragma unroll
            for (int q = 0; q < 3; q++) { lo = fminf(fminf(lo, mb[q]), mc[q]); hi = fmaxf(fmaxf(hi, mb[q]), mc[q]); }
            const float e = (bwd - om) * 0.5f;
            float val = om - e;
            val = val < lo ? lo : (val > hi ? hi : val);
            if (val < 0.0f) val = 0.0f;
            out = val;
        }
        corrM[o] = out;
        u = up;
#pragma unroll
        for (int q = 0; q < 3; q++) { ma[q] = mb[q]; mb[q] = mc[q]; }
    }
}

// contract of k_bfecc_smoke_correct
template <bool CHECK>
__global__ void __launch_bounds__(AT_THREADS, 4)
k_bfecc_smoke_tile(const AdvCtx c, const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmF, const float *__restrict__ fwdM,
                   const unsigned char *__restrict__ mask, const unsigned char *__restrict__ tile_flags, const int ntx,
                   float *__restrict__ corrM, const float dt, const float sa, const int ib, const int ie, int *bad)
{
    extern __shared__ __align__(128) unsigned char at_smem[];
    float *sF = reinterpret_cast<float *>(at_smem), *sO = sF + AT_STL * AT_PW, *sU = sO + AT_SVL * AT_PW, *sV = sU + AT_SVL * AT_PW;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sV + AT_SVL * AT_PW);
    const int tid = threadIdx.x;
    const int ty = ib / AT_STI + blockIdx.y, tx = blockIdx.x;
    const int i0 = max(ty * AT_STI, ib), i1 = min((ty + 1) * AT_STI, ie);
    ATile T;
    T.ls0 = ty * AT_STI - (AT_R + 1); T.cs0 = tx * AT_TJ - AT_CH;
    at_tile_geometry(c, T, AT_STL);
    const int vls0 = ty * AT_STI - 1;
    if (tid == 0) rq_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const unsigned b = rq_s32(bar);
        rq_mbar_expect_tx(bar, (AT_STL + 3u * AT_SVL) * AT_PW * 4u);
        at_tensor_load(sF, &tmF, T.cs0, T.ls0 - c.i_alloc0, b);
        at_tensor_load(sO, &tmO, T.cs0, vls0 - c.i_alloc0, b);
        at_tensor_load(sU, &tmU, T.cs0, vls0 - c.i_alloc0, b);
        at_tensor_load(sV, &tmV, T.cs0, vls0 - c.i_alloc0, b);
    }
    const bool fast = T.ls0 >= 2 && T.ls0 + AT_STL <= c.NX - 2 && T.cs0 >= 2 && T.cs0 + AT_PW <= c.NY - 2 &&
                      T.ls0 >= c.i_alloc0 && T.ls0 + AT_STL <= c.i_alloc0 + c.lines_alloc &&
                      tile_flags[((ty * AT_STI) / AT_TI) * ntx + tx] != 0;
    at_wait_tiles(bar, 0, bad);
    const int j = tx * AT_TJ + (tid & 127);
    if (fast) at_bfecc_smoke_cells<true, CHECK>(c, T, sF, sO, sU, sV, vls0, fwdM, mask, corrM, dt, sa, ty * AT_STI, i0, i1, j, bad);
    else at_bfecc_smoke_cells<false, CHECK>(c, T, sF, sO, sU, sV, vls0, fwdM, mask, corrM, dt, sa, ty * AT_STI, i0, i1, j, bad);
}

// ======================= vorticity confinement + turbulence (fluid.go:449-526) on tiles ==================================
// Contract of k_confine_turbulence.  U and V of the tile plus a two-cell halo arrive by tensor-map TMA; the curl of the
// tile plus a one-cell halo is computed into shared memory (the reference's `curl` array never touches HBM), a lane per
// cell, then every cell applies the force and the turbulence.  The round-1 kernel (4 cells per thread, row loads from
// global memory) issued one instruction per warp every 18 cycles behind its global loads (ncu round 2: long scoreboard
// the top stall, 0.35 of the HBM peak on a developed flow).
#ifndef AT_CTI
#define AT_CTI 16
#endif
#define AT_CVL (AT_CTI + 4)                  // staged lines of U, V
#define AT_CML (AT_CTI + 2)                  // lines of the curl tile
#define AT_CW 132                            // curl tile pitch: columns j0 - 1 .. j0 + 128 (+ padding)
#define AT_CSMEM (2 * AT_CVL * AT_PW * 4 + AT_CML * AT_CW * 4 + 16)

// curl of cell (i, j) (fluid.go:453-466) from the staged tiles; `ok` = the cell lies in 1..NumX-2 x 1..NumY-2 with both
// neighbouring lines resident
__device__ __forceinline__ float at_curl(const float *__restrict__ pu, const float *__restrict__ pv, const unsigned m, const bool ok, const float h)
{
    float cu = 0.0f;
    if (ok && (m & MK_C)) {
        const float dvdx = div0((pv[AT_PW] - pv[-AT_PW]) * 0.5f, h);
        const float dudy = div0((pu[1] - pu[-1]) * 0.5f, h);
        cu = dvdx - dudy;
    }
    return cu;
}

__global__ void __launch_bounds__(AT_THREADS, 6)
k_confine_tile(const AdvCtx c, const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
               const unsigned char *__restrict__ mask, const float *__restrict__ nU, const float *__restrict__ nV,
               float *__restrict__ dstU, float *__restrict__ dstV, const float h, const float dt, const float confinement,
               const float turbStrength, const int ib, const int ie, int *bad)
{
    extern __shared__ __align__(128) unsigned char at_smem[];
    float *sU = reinterpret_cast<float *>(at_smem), *sV = sU + AT_CVL * AT_PW;
    float *sC = sV + AT_CVL * AT_PW;                               // [AT_CML][AT_CW]: lines t0 - 1 .., columns j0 - 1 ..
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sC + AT_CML * AT_CW);
    const int tid = threadIdx.x;
    const int ty = ib / AT_CTI + blockIdx.y, tx = blockIdx.x;
    const int t0 = ty * AT_CTI, j0 = tx * AT_TJ;
    const int i0 = max(t0, ib), i1 = min(t0 + AT_CTI, ie);
    const int ls0 = t0 - 2, cs0 = j0 - AT_CH;
    if (tid == 0) rq_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        const unsigned b = rq_s32(bar);
        rq_mbar_expect_tx(bar, 2u * AT_CVL * AT_PW * 4u);
        at_tensor_load(sU, &tmU, cs0, ls0 - c.i_alloc0, b);
        at_tensor_load(sV, &tmV, cs0, ls0 - c.i_alloc0, b);
    }
    const int cl = tid & 127, grp = tid >> 7;
    const int j = j0 + cl;
    const bool colok = j >= 1 && j <= c.NY - 2;
    // the mask bytes of this thread's column for the lines it touches (curl lines t0 - 1 + grp * CL .., which contain its
    // cells), fetched while the tiles are in flight
    constexpr int CL = (AT_CML + AT_LG - 1) / AT_LG;               // curl lines per line group
    unsigned mk[CL];
#pragma unroll
    for (int q = 0; q < CL; q++) {
        const int i = t0 - 1 + grp * CL + q;
        mk[q] = (colok && i >= max(c.i_alloc0, 0) && i < min(c.i_alloc0 + c.lines_alloc, c.NX)) ? mask[(size_t)(i - c.i_alloc0) * c.pitch + j] : 0u;
    }
    unsigned mside = 0;
    if (tid < 2 * AT_CML) {
        const int side = tid >= AT_CML, i = t0 - 1 + (tid - side * AT_CML), jj = side ? j0 + AT_TJ : j0 - 1;
        if (jj >= 1 && jj <= c.NY - 2 && i >= max(c.i_alloc0, 0) && i < min(c.i_alloc0 + c.lines_alloc, c.NX)) mside = mask[(size_t)(i - c.i_alloc0) * c.pitch + jj];
    }
    at_wait_tiles(bar, 0, bad);
    if (confinement != 0.0f) {
        // curl of lines t0 - 1 .. t0 + AT_CTI, columns j0 - 1 .. j0 + 128: a lane per column, the two edge columns on the side
#pragma unroll
        for (int q = 0; q < CL; q++) {
            const int li = grp * CL + q, i = t0 - 1 + li;
            if (li < AT_CML) {
                const bool lineok = i >= 1 && i <= c.NX - 2 && i - 1 >= c.i_alloc0 && i + 1 < c.i_alloc0 + c.lines_alloc;
                const int o = (i - ls0) * AT_PW + (j - cs0);
                sC[li * AT_CW + cl + 1] = at_curl(sU + o, sV + o, mk[q], lineok && colok, h);
            }
        }
        if (tid < 2 * AT_CML) {
            const int side = tid >= AT_CML, li = tid - side * AT_CML;
            const int i = t0 - 1 + li, jj = side ? j0 + AT_TJ : j0 - 1;
            const bool ok = i >= 1 && i <= c.NX - 2 && i - 1 >= c.i_alloc0 && i + 1 < c.i_alloc0 + c.lines_alloc && jj >= 1 && jj <= c.NY - 2;
            const int o = (i - ls0) * AT_PW + (jj - cs0);
            sC[li * AT_CW + (side ? AT_TJ + 1 : 0)] = at_curl(sU + o, sV + o, mside, ok, h);
        }
        __syncthreads();
    }
    if (j >= c.NY) return;
    constexpr int half = AT_CTI / AT_LG;
    static_assert(AT_CTI % AT_LG == 0 && half + 1 <= CL, "a thread's cells lie inside its curl lines");
#pragma unroll
    for (int q = 0; q < half; q++) {
        const int i = t0 + grp * half + q;
        if (i < i0 || i >= i1) continue;
        const size_t o = (size_t)(i - c.i_alloc0) * c.pitch + j;
        float u = sU[(i - ls0) * AT_PW + (j - cs0)], v = sV[(i - ls0) * AT_PW + (j - cs0)];
        // line i is curl line (i - t0 + 1) = grp * half + q + 1; this thread's curl lines start at grp * CL
        static_assert(AT_LG == 2 && CL == half + 1, "index arithmetic below");
        const unsigned m = grp ? mk[q] : mk[q + 1];        // (static indices: the array stays in registers)
        if (i >= 1 && i <= c.NX - 2 && colok && (m & MK_C)) {
            if (confinement != 0.0f) {
                const float eps = 1e-5f;
                const float *pc = sC + (i - t0 + 1) * AT_CW + (cl + 1);
                const float c0 = pc[0];
                float gx = div0((fabsf(pc[AT_CW]) - fabsf(pc[-AT_CW])) * 0.5f, h);
                float gy = div0((fabsf(pc[1]) - fabsf(pc[-1])) * 0.5f, h);
                const float gx2 = gx * gx, gy2 = gy * gy;
                const float mag = sqrt0(gx2 + gy2) + eps;
                gx = div0(gx, mag);
                gy = div0(gy, mag);
                const float uu = u * u, vv = v * v;
                const float localVel = sqrt0(uu + vv);
                const float lv = localVel * 0.1f;
                const float strength = confinement * (1.0f + lv);
                const float fu = ((strength * gy) * c0) * dt;
                const float fv = ((strength * gx) * c0) * dt;
                u = u + fu;
                v = v - fv;
            }
            if (turbStrength > 0.0f) {
                const float uu = u * u, vv = v * v;
                const float localVel = sqrt0(uu + vv);
                if (localVel > 0.1f) {
                    const float noiseU = nU[o] * turbStrength;
                    const float noiseV = nV[o] * turbStrength;
                    const float factor = fminf(localVel * 0.5f, 1.0f);
                    const float du = noiseU * factor, dv = noiseV * factor;
                    u = u + du;
                    v = v + dv;
                }
            }
        }
        dstU[o] = u;
        dstV[o] = v;
    }
}
