// rb_fused.cuh -- all red-black iterations of one pressure solve fused in ONE pass
// over HBM (temporal blocking in shared memory).
//
// A CTA owns a block of `chunk` lines x TJ columns.  It streams lines (constant i,
// contiguous in j) through a ring of RB_NL line slots in shared memory.  Half sweep
// s (s = 0 .. 2K-1, colour s&1) runs 2 lines behind half sweep s-1, so at every step
// the 2K half sweeps work on 2K different lines that share no face: one
// __syncthreads per step, 2K half sweeps per byte loaded.  Warp w works for half
// sweep w & 15.  Halo: H = 16 lines / columns on each side are recomputed
// redundantly (dependence radius of 2K <= 16 half sweeps), so a CTA never waits for a
// neighbour and results do not depend on the decomposition.
//
// Inside shared memory even and odd columns live in separate arrays, so the active
// cells of one colour (every other column) are contiguous: conflict-free LDS/STS.
//
// The per-cell arithmetic is project_cell() of kernels.cuh: the reference's update
// (pkg/fluid/fluid.go:196-229), bit for bit; only the visiting order differs.
#pragma once
#include "kernels.cuh"

#define RB_NL 34          // line slots: 2*15 lag + current + next + being loaded + being stored
#define RB_H 16           // halo lines / columns (max 16 half sweeps per pass)
#define RB_THREADS 1024
#define RB_TJ_MAX 456     // owned columns per strip (multiple of 4); WL = TJ + 36 <= 492

struct RBFused {
    Grid g;
    const float *U, *V;          // input planes
    const float *Pin;            // nullptr: pressure known to be zero (fluid.go:83 fused away)
    const unsigned char *mask;   // bit0 cell fluid, bit1 S[i-1,j], bit2 S[i+1,j], bit3 S[i,j-1], bit4 S[i,j+1]
    float *Uo, *Vo, *Po;         // output planes (out of place)
    float omega[16];             // per half sweep
    float damping, cp;
    int nstages;                 // 2K half sweeps in this pass (<= 16)
    int stage0;                  // index of the first half sweep of this pass (stats slot = (stage0+s)>>1)
    int TJ, WL;                  // owned columns per strip, loaded columns (multiple of 4)
    int chunk;                   // owned lines per CTA
    int ib, ie;                  // global line range this rank must produce
    unsigned *stats;
    // optional fused turbulence (fluid.go:496-526) applied to the finished lines
    const float *noiseU, *noiseV;
    float turb;                  // TurbulenceStrength*dt, 0 = off
};

__device__ __forceinline__ int rb_slot(int line_rel) {   // line_rel >= 0
    return line_rel % RB_NL;
}

__global__ void __launch_bounds__(RB_THREADS, 1) k_rb_fused(const RBFused P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int WL = P.WL, WQ = WL >> 1;
    // layout: [slot][parity][q]
    float *sU = reinterpret_cast<float *>(smem_raw);
    float *sV = sU + RB_NL * WL;
    float *sP = sV + RB_NL * WL;
    unsigned char *sM = reinterpret_cast<unsigned char *>(sP + RB_NL * WL);

    const Grid g = P.g;
    const int NX = g.NX, NY = g.NY;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int strip = blockIdx.x;
    const int i0c = P.ib + blockIdx.y * P.chunk;                 // owned lines [i0c, i1c)
    const int i1c = min(i0c + P.chunk, P.ie);
    if (i0c >= i1c) return;
    const int jr0 = strip * P.TJ - RB_H;                         // global j of local column 0 (multiple of 4)
    const int e0 = i0c - RB_H;                                   // first extended line
    const int e1 = i1c + RB_H;                                   // processed lines are [e0, e1); line e1 is loaded too
    const int nst = P.nstages;

    // ---- loader: line L -> registers -> slot
    const int ngroups = WL >> 2;                                 // float4 groups per line
    // thread roles for loading: [0,ng) U, [ng,2ng) V, [2ng,3ng) mask, [3ng,4ng) Pin
    const int lrole = tid / ngroups, lgrp = tid - lrole * ngroups;
    float4 lreg = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned lmask = 0;

    auto issue_load = [&](int L) {
        lreg = make_float4(0.f, 0.f, 0.f, 0.f);
        lmask = 0;
        if (lrole > 3) return;
        const int j = jr0 + 4 * lgrp;
        const bool line_ok = (L >= g.i_alloc0) && (L < g.i_alloc0 + g.lines_alloc) && (L >= 0) && (L < NX);
        if (!line_ok || j < 0 || j >= g.pitch) return;          // j is a multiple of 4: whole group in or out
        const size_t a = g.at(L, j);
        if (lrole == 0) lreg = __ldg(reinterpret_cast<const float4 *>(P.U + a));
        else if (lrole == 1) lreg = __ldg(reinterpret_cast<const float4 *>(P.V + a));
        else if (lrole == 2) {
            unsigned m = __ldg(reinterpret_cast<const unsigned *>(P.mask + a));
            // cells that are never updated (ring, outside the interior) get mask 0
            if (L < 1 || L > NX - 2) m = 0;
            else {
                if (j < 1) m &= 0xffffff00u;
                if (j + 3 > NY - 2) {
#pragma unroll
                    for (int k = 0; k < 4; k++) if (j + k > NY - 2) m &= ~(0xffu << (8 * k));
                }
            }
            lmask = m;
        } else if (P.Pin) lreg = __ldg(reinterpret_cast<const float4 *>(P.Pin + a));
    };
    auto commit_load = [&](int L) {
        if (lrole > 3) return;
        const int sl = rb_slot(L - e0);
        const int q = 2 * lgrp;
        if (lrole == 2) {
            unsigned short ev = (unsigned short)((lmask & 0xffu) | ((lmask >> 8) & 0xff00u));
            unsigned short od = (unsigned short)(((lmask >> 8) & 0xffu) | ((lmask >> 16) & 0xff00u));
            *reinterpret_cast<unsigned short *>(sM + (sl * 2 + 0) * WQ + q) = ev;
            *reinterpret_cast<unsigned short *>(sM + (sl * 2 + 1) * WQ + q) = od;
        } else {
            float *base = lrole == 0 ? sU : (lrole == 1 ? sV : sP);
            *reinterpret_cast<float2 *>(base + (sl * 2 + 0) * WQ + q) = make_float2(lreg.x, lreg.z);
            *reinterpret_cast<float2 *>(base + (sl * 2 + 1) * WQ + q) = make_float2(lreg.y, lreg.w);
        }
    };

    // ---- prologue: lines e0 and e0+1
    issue_load(e0); commit_load(e0);
    issue_load(e0 + 1); commit_load(e0 + 1);
    __syncthreads();

    // ---- compute roles
    const int s = warp & 15;                 // half sweep of this warp
    const int sub = warp >> 4;               // which half of the line
    const int colour = (P.stage0 + s) & 1;
    const float omega = P.omega[s < nst ? s : 0];
    const int qhalf = (WQ + 1) >> 1;
    const int q_lo = sub * qhalf, q_hi = min(WQ, q_lo + qhalf);
    float mymax = 0.0f;

    // ---- store roles (finished line -> global): threads [0, 3*TJ/4)
    const int sgroups = P.TJ >> 2;
    const int srole = tid / sgroups, sgrp = tid - srole * sgroups;

    const int last_stage_lag = 2 * (nst - 1);
    const int nsteps = (e1 - e0) + last_stage_lag + 1;
    for (int t = 0; t < nsteps; t++) {
        // (1) start loading line e0+t+2 (used by half sweep 0 in step t+1)
        const int Lnext = e0 + t + 2;
        const bool do_load = Lnext <= e1;
        if (do_load) issue_load(Lnext);

        // (2) write out the line finished by the previous step
        {
            const int r = e0 + (t - 1) - last_stage_lag;
            if (t >= 1 && r >= i0c && r < i1c && srole < 3) {
                const int lj = RB_H + 4 * sgrp;                  // local column of the group (multiple of 4)
                const int j = jr0 + lj;
                if (j < NY) {
                    const int sl = rb_slot(r - e0);
                    const int q = lj >> 1;
                    const float *base = srole == 0 ? sU : (srole == 1 ? sV : sP);
                    float2 ev = *reinterpret_cast<const float2 *>(base + (sl * 2 + 0) * WQ + q);
                    float2 od = *reinterpret_cast<const float2 *>(base + (sl * 2 + 1) * WQ + q);
                    float4 out = make_float4(ev.x, od.x, ev.y, od.y);
                    const size_t a = g.at(r, j);
                    if (srole < 2 && P.turb > 0.0f && r >= 1 && r <= NX - 2) {
                        // fused addTurbulence on the finished (u,v) of fluid interior cells
                        const float *ob = srole == 0 ? sV : sU;  // the other component
                        float2 oe = *reinterpret_cast<const float2 *>(ob + (sl * 2 + 0) * WQ + q);
                        float2 oo = *reinterpret_cast<const float2 *>(ob + (sl * 2 + 1) * WQ + q);
                        float other[4] = { oe.x, oo.x, oe.y, oo.y };
                        float mine[4] = { out.x, out.y, out.z, out.w };
                        const unsigned m4 = __ldg(reinterpret_cast<const unsigned *>(P.mask + a));
                        const float *noise = srole == 0 ? P.noiseU : P.noiseV;
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int jj = j + k;
                            if (((m4 >> (8 * k)) & 1u) && jj >= 1 && jj <= NY - 2) {
                                float u = srole == 0 ? mine[k] : other[k];
                                float v = srole == 0 ? other[k] : mine[k];
                                float uu = u * u, vv = v * v;
                                float localVel = sqrtf(uu + vv);
                                if (localVel > 0.1f) {
                                    float nz = __ldg(noise + a + k) * P.turb;
                                    float factor = go_minf(localVel * 0.5f, 1.0f);
                                    float d = nz * factor;
                                    mine[k] = mine[k] + d;
                                }
                            }
                        }
                        out = make_float4(mine[0], mine[1], mine[2], mine[3]);
                    }
                    float *dst = srole == 0 ? P.Uo : (srole == 1 ? P.Vo : P.Po);
                    if (j + 3 < NY) *reinterpret_cast<float4 *>(dst + a) = out;
                    else {
                        float o[4] = { out.x, out.y, out.z, out.w };
                        for (int k = 0; k < 4 && j + k < NY; k++) dst[a + k] = o[k];
                    }
                }
            }
        }

        // (3) the half sweeps
        if (s < nst) {
            const int r = e0 + t - 2 * s;
            if (r >= e0 && r < e1 && r >= 1 && r <= NX - 2) {
                const int a = (colour + r) & 1;                  // column parity of the active cells
                const int sl = rb_slot(r - e0), sl1 = rb_slot(r + 1 - e0);
                const float *u0p = sU + (sl * 2 + a) * WQ, *u1p = sU + (sl1 * 2 + a) * WQ;
                const float *v0p = sV + (sl * 2 + a) * WQ, *v1p = sV + (sl * 2 + (1 - a)) * WQ + a;
                float *pp_ = sP + (sl * 2 + a) * WQ;
                const unsigned char *mp = sM + (sl * 2 + a) * WQ;
                const bool row_owned = (r >= i0c) && (r < i1c);
                const int q_end = a ? min(q_hi, WQ - 1) : q_hi;  // column lj+1 must exist
                // up to 4 cells per lane; all loads first, then arithmetic, then stores (ILP)
                float u0[4], u1[4], v0[4], v1[4], pr[4];
                unsigned mk[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int q = q_lo + lane + 32 * k;
                    mk[k] = 0;
                    if (q < q_end) {
                        mk[k] = mp[q];
                        u0[k] = u0p[q]; u1[k] = u1p[q]; v0[k] = v0p[q]; v1[k] = v1p[q]; pr[k] = pp_[q];
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned m = mk[k];
                    if ((m & 1u) && (m & 30u)) {
                        CellS c;
                        c.c = 1.0f;
                        c.sx0 = (m & 2u) ? 1.0f : 0.0f; c.sx1 = (m & 4u) ? 1.0f : 0.0f;
                        c.sy0 = (m & 8u) ? 1.0f : 0.0f; c.sy1 = (m & 16u) ? 1.0f : 0.0f;
                        float ad = project_cell(u0[k], u1[k], v0[k], v1[k], pr[k], c, omega, P.damping, P.cp);
                        const int q = q_lo + lane + 32 * k;
                        const_cast<float *>(u0p)[q] = u0[k]; const_cast<float *>(u1p)[q] = u1[k];
                        const_cast<float *>(v0p)[q] = v0[k]; const_cast<float *>(v1p)[q] = v1[k];
                        pp_[q] = pr[k];
                        const int lj = 2 * q + a;
                        if (row_owned && lj >= RB_H && lj < RB_H + P.TJ && ad > mymax) mymax = ad;
                    }
                }
            }
        }

        // (4) the freshly loaded line becomes visible for the next step
        if (do_load) commit_load(Lnext);
        __syncthreads();
    }

    if (s < nst) {
        mymax = warp_max(mymax);
        if (lane == 0 && mymax > 0.0f) atomicMax(P.stats + ((P.stage0 + s) >> 1), __float_as_uint(mymax));
    }
}
