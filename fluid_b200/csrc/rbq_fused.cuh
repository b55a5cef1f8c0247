// rbq_fused.cuh -- the fastest pressure solve: red-black SOR in PRESSURE FORM, every
// iteration of a pass (<= 8 iterations = 16 half sweeps) fused in one trip over HBM.
//
// Same streaming structure as rb_fused.cuh (ring of line slots in shared memory, half
// sweep s two lines behind half sweep s-1, one __syncthreads per line, 16-cell halo
// recomputed redundantly), but the state that circulates is one scalar per cell:
//   q(i,j)  = sum of the corrections the reference applies at cell (i,j)
//             (the `p` of pkg/fluid/fluid.go:218-222 accumulated over the sweeps),
// against the frozen divergence D0 of the field the pass started from:
//   div(i,j) = D0 + s*q(i,j) - (q(i-1,j)+q(i+1,j)+q(i,j-1)+q(i,j+1)),  q == 0 in solids
//   q'       = fma(wd/s, nb - D0, fma(-wd, q, q)),   wd = omega*PressureDamping.
// U, V and p are materialised once when a line leaves the window:
//   U[i,j] = (U0 - (S[i-1,j] ? q[i,j] : 0)) + (S[i,j] ? q[i-1,j] : 0), V alike, p = fma(cp,q,p).
// This is algebraically the update of fluid.go:196-229 in red-black order (Q-4);
// rounding differs at the 1e-6 level (tests/test_parity_gpu.py), and the kernel is
// checked bit for bit against its own CPU restatement fo_project_redblack_q.
//
// Why: ncu showed the face form issue-bound at ~110 instructions per cell update
// (IEEE division, 4 face read-modify-writes, mask decode).  Here an update is 4 adds,
// 1 mul, 2 fma on 7 shared-memory words, 8 cells per lane with 128-bit LDS/STS.
//
// Warp roles (768 threads): warps 0-15 = the 16 half sweeps (one line each per step),
// warps 16-19 = loader (global -> D0, 1/s, mask in a slot, prefetched one step ahead),
// warps 20-23 = writer (slot -> U, V, p in global, inputs prefetched one step ahead).
#pragma once
#include "kernels.cuh"
#include "advect_fused.cuh"

#define RQ_NL 35          // line slots (one more than rb_fused: the write-out reads line r-1 as well)
#define RQ_H 16
#define RQ_THREADS 768
#define RQ_TJ_MAX 456     // multiple of 8; WL = TJ + 40 <= 496 (WL/2 multiple of 4 for 128-bit LDS);
                          // 35 * 496 * 13 B = 220.4 KB of shared memory

struct RBQ {
    Grid g;
    const float *U, *V;          // field the pass starts from
    const float *Pin;            // nullptr: pressure known to be zero
    const unsigned char *mask;
    float *Uo, *Vo, *Po;
    float wd[16];                // omega*damping per half sweep
    float cp;
    int nstages, stage0;
    int TJ, WL, chunk, ib, ie;
    unsigned *stats;             // per-iteration max |div| (only when STATS)
    const float *noiseU, *noiseV;
    float turb;
};

struct RQLine { float4 u, u1, v; float v4; unsigned m; };

template <bool STATS>
__global__ void __launch_bounds__(RQ_THREADS, 1) k_rbq_fused(const RBQ P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int WL = P.WL, WQ = WL >> 1, ROW = WL;             // floats per slot in one plane
    float *sQ = reinterpret_cast<float *>(smem_raw);        // [slot][parity][q]
    float *sD = sQ + RQ_NL * WL;
    float *sR = sD + RQ_NL * WL;
    unsigned char *sM = reinterpret_cast<unsigned char *>(sR + RQ_NL * WL);

    const Grid g = P.g;
    const int NX = g.NX, NY = g.NY, PIT = g.pitch;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int strip = blockIdx.x;
    const int i0c = P.ib + blockIdx.y * P.chunk;
    const int i1c = min(i0c + P.chunk, P.ie);
    if (i0c >= i1c) return;
    const int jr0 = strip * P.TJ - RQ_H;
    const int e0 = i0c - RQ_H, e1 = i1c + RQ_H;
    const int nst = P.nstages;
    const int lag = 2 * (nst - 1);
    const int nsteps = (e1 - e0) + lag + 1;

    // ================= loader: lines e0 .. e1 =================
    const int ld = tid - 512;                                 // loader thread index (warps 16-19)
    const bool is_loader = ld >= 0 && ld < (WL >> 2);
    auto fetch = [&](int L, RQLine &x) {
        x.u = x.u1 = x.v = make_float4(0.f, 0.f, 0.f, 0.f);
        x.v4 = 0.0f; x.m = 0;
        const int j = jr0 + 4 * ld;
        if (L < 1 || L > NX - 2 || j < 0 || j >= PIT) return;                  // only interior lines hold updatable cells
        if (L < g.i_alloc0 || L + 1 >= g.i_alloc0 + g.lines_alloc) return;     // outside this rank's slab
        const int o = (L - g.i_alloc0) * PIT + j;
        x.m = __ldg(reinterpret_cast<const unsigned *>(P.mask + o));
        x.u = ld4(P.U + o);
        x.u1 = ld4(P.U + o + PIT);
        x.v = ld4(P.V + o);
        if (j + 4 < PIT) x.v4 = __ldg(P.V + o + 4);
    };
    auto commit = [&](int sl, const RQLine &x) {
        const int j = jr0 + 4 * ld;
        float u0[4], u1[4], v[5], d[4], r[4];
        unpack(x.u, u0); unpack(x.u1, u1); unpack(x.v, v); v[4] = x.v4;
        unsigned mk = x.m;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned m = (mk >> (8 * k)) & 0xffu;
            const int jj = j + k;
            const int ns = __popc(m & 30u);
            const bool upd = (m & MK_C) && ns > 0 && jj >= 1 && jj <= NY - 2;
            d[k] = ((u1[k] - u0[k]) + v[k + 1]) - v[k];
            r[k] = !upd ? 0.0f : (ns == 1 ? 1.0f : (ns == 2 ? 0.5f : (ns == 3 ? (1.0f / 3.0f) : 0.25f)));
            if (!upd) { d[k] = 0.0f; mk &= ~(0xffu << (8 * k)); }
        }
        const int q = 2 * ld;
        const int b0 = sl * ROW + q, b1 = b0 + WQ;
        *reinterpret_cast<float2 *>(sD + b0) = make_float2(d[0], d[2]);
        *reinterpret_cast<float2 *>(sD + b1) = make_float2(d[1], d[3]);
        *reinterpret_cast<float2 *>(sR + b0) = make_float2(r[0], r[2]);
        *reinterpret_cast<float2 *>(sR + b1) = make_float2(r[1], r[3]);
        *reinterpret_cast<float2 *>(sQ + b0) = make_float2(0.f, 0.f);
        *reinterpret_cast<float2 *>(sQ + b1) = make_float2(0.f, 0.f);
        if (STATS) {   // updatable-cell bytes, only for the residual statistics
            *reinterpret_cast<unsigned short *>(sM + b0) = (unsigned short)((mk & 0xffu) | ((mk >> 8) & 0xff00u));
            *reinterpret_cast<unsigned short *>(sM + b1) = (unsigned short)(((mk >> 8) & 0xffu) | ((mk >> 16) & 0xff00u));
        }
    };

    RQLine lnA, lnB;                                          // two lines in flight
    if (is_loader) {
        fetch(e0, lnA); commit(0, lnA);
        fetch(e0 + 1, lnA); commit(1, lnA);
        {   // slot of line e0-1 (relative -1): q must read as zero
            const int q = 2 * ld, b0 = (RQ_NL - 1) * ROW + q;
            *reinterpret_cast<float2 *>(sQ + b0) = make_float2(0.f, 0.f);
            *reinterpret_cast<float2 *>(sQ + b0 + WQ) = make_float2(0.f, 0.f);
        }
        fetch(e0 + 2, lnA);                                   // committed at the end of step 0
    }
    __syncthreads();

    // ================= half sweeps =================
    const int s = warp;                                       // warps 0..15
    const bool is_compute = warp < 16 && s < nst;
    const int colour = (P.stage0 + s) & 1;
    const float wd = P.wd[is_compute ? s : 0];
    const int g0 = lane, g1 = lane + 32;                      // 4-cell groups of this lane
    const int ngrp = WQ >> 2;
    const bool on0 = g0 < ngrp, on1 = g1 < ngrp;
    // line processed at step t: rel = t - 2s; valid while rel in [rel_lo, rel_hi)
    const int rel_lo = max(0, 1 - e0), rel_hi = min(e1 - e0, NX - 1 - e0);
    int c_rel = -2 * s;
    int c_sl = ((c_rel % RQ_NL) + RQ_NL) % RQ_NL;
    int c_a = (colour + e0 + c_rel) & 1;
    float mymax = 0.0f;

    // ================= writer =================
    const int st = tid - 640;                                 // warps 20-23
    const bool is_writer = st >= 0 && st < (P.TJ >> 2);
    const int w_lj = RQ_H + 4 * st, w_j = jr0 + w_lj;
    float4 wU = make_float4(0, 0, 0, 0), wV = wU, wP = wU;
    unsigned wM = 0;
    auto wfetch = [&](int r) {
        if (r < i0c || r >= i1c || w_j >= NY) return;
        const int o = (r - g.i_alloc0) * PIT + w_j;
        wU = ld4(P.U + o);
        wV = ld4(P.V + o);
        wM = __ldg(reinterpret_cast<const unsigned *>(P.mask + o));
        if (P.Pin) wP = ld4(P.Pin + o);
    };
    // the writer handles line r_w(t) = e0 + (t-1) - lag at step t; its inputs are fetched at step t-1
    if (is_writer) wfetch(e0 - lag);                          // for t = 1

    for (int t = 0; t < nsteps; t++) {
        if (is_compute) {
            if (c_rel >= rel_lo && c_rel < rel_hi) {
                const int slp = c_sl + 1 == RQ_NL ? 0 : c_sl + 1;
                const int slm = c_sl == 0 ? RQ_NL - 1 : c_sl - 1;
                const int own = c_sl * ROW + c_a * WQ, oth = c_sl * ROW + (WQ - c_a * WQ);
                const int upo = slp * ROW + c_a * WQ, dno = slm * ROW + c_a * WQ;
                const int xo = c_a ? 4 : -1;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const int q0 = 4 * (half ? g1 : g0);
                    if (half ? on1 : on0) {
                        float *qown = sQ + own + q0;
                        const float4 qo = *reinterpret_cast<const float4 *>(qown);
                        const float4 up = *reinterpret_cast<const float4 *>(sQ + upo + q0);
                        const float4 dn = *reinterpret_cast<const float4 *>(sQ + dno + q0);
                        const float4 ot = *reinterpret_cast<const float4 *>(sQ + oth + q0);
                        const float ox = sQ[oth + q0 + xo];
                        const float4 d0 = *reinterpret_cast<const float4 *>(sD + own + q0);
                        const float4 rs = *reinterpret_cast<const float4 *>(sR + own + q0);
                        // left / right neighbours of cell k: other-parity indices q0+k-1+a and q0+k+a
                        float l0, l1, l2, l3, r0_, r1_, r2_, r3_;
                        if (c_a) { l0 = ot.x; l1 = ot.y; l2 = ot.z; l3 = ot.w; r0_ = ot.y; r1_ = ot.z; r2_ = ot.w; r3_ = ox; }
                        else     { l0 = ox;   l1 = ot.x; l2 = ot.y; l3 = ot.z; r0_ = ot.x; r1_ = ot.y; r2_ = ot.z; r3_ = ot.w; }
                        const float nb0 = ((dn.x + up.x) + l0) + r0_;
                        const float nb1 = ((dn.y + up.y) + l1) + r1_;
                        const float nb2 = ((dn.z + up.z) + l2) + r2_;
                        const float nb3 = ((dn.w + up.w) + l3) + r3_;
                        const float t0 = nb0 - d0.x, t1 = nb1 - d0.y, t2 = nb2 - d0.z, t3 = nb3 - d0.w;
                        float4 qn;
                        qn.x = __fmaf_rn(wd * rs.x, t0, __fmaf_rn(-wd, qo.x, qo.x));
                        qn.y = __fmaf_rn(wd * rs.y, t1, __fmaf_rn(-wd, qo.y, qo.y));
                        qn.z = __fmaf_rn(wd * rs.z, t2, __fmaf_rn(-wd, qo.z, qo.z));
                        qn.w = __fmaf_rn(wd * rs.w, t3, __fmaf_rn(-wd, qo.w, qo.w));
                        *reinterpret_cast<float4 *>(qown) = qn;
                        if (STATS) {
                            const unsigned mk = *reinterpret_cast<const unsigned *>(sM + own + q0);
                            const int r = e0 + c_rel;
                            const bool row_owned = (r >= i0c) && (r < i1c);
                            const float qv[4] = { qo.x, qo.y, qo.z, qo.w }, tv[4] = { t0, t1, t2, t3 };
#pragma unroll
                            for (int k = 0; k < 4; k++) {
                                const unsigned m = (mk >> (8 * k)) & 0xffu;
                                const int lj = 2 * (q0 + k) + c_a;
                                if ((m & MK_C) && row_owned && lj >= RQ_H && lj < RQ_H + P.TJ) {
                                    const float ns = (float)__popc(m & 30u);
                                    const float ad = fabsf(__fmaf_rn(ns, qv[k], -tv[k]));
                                    if (ad > mymax) mymax = ad;
                                }
                            }
                        }
                    }
                }
            }
            c_rel++;
            c_sl = c_sl + 1 == RQ_NL ? 0 : c_sl + 1;
            c_a ^= 1;
        } else if (is_loader) {
            // line e0+t+2 (fetched a step ago) becomes visible for step t+1; start fetching e0+t+3
            const int rel = t + 2;
            if (e0 + rel <= e1) {
                const int sl = rel % RQ_NL;
                if (t & 1) { fetch(e0 + rel + 1, lnA); commit(sl, lnB); }
                else       { fetch(e0 + rel + 1, lnB); commit(sl, lnA); }
            }
        } else if (is_writer) {
            const int r = e0 + (t - 1) - lag;
            if (t >= 1 && r >= i0c && r < i1c && w_j < NY) {
                const int o = (r - g.i_alloc0) * PIT + w_j;
                const int rel = r - e0;
                const int sl = rel % RQ_NL, slm = (rel + RQ_NL - 1) % RQ_NL;
                const int q = w_lj >> 1;
                float u[4], v[4], pin[4], qc[4], qx[4], ql;
                unpack(wU, u); unpack(wV, v); unpack(wP, pin);
                const unsigned m4 = wM;
                {
                    const float2 ev = *reinterpret_cast<const float2 *>(sQ + sl * ROW + q);
                    const float2 od = *reinterpret_cast<const float2 *>(sQ + sl * ROW + WQ + q);
                    qc[0] = ev.x; qc[1] = od.x; qc[2] = ev.y; qc[3] = od.y;
                    const float2 evm = *reinterpret_cast<const float2 *>(sQ + slm * ROW + q);
                    const float2 odm = *reinterpret_cast<const float2 *>(sQ + slm * ROW + WQ + q);
                    qx[0] = evm.x; qx[1] = odm.x; qx[2] = evm.y; qx[3] = odm.y;
                    ql = sQ[sl * ROW + WQ + q - 1];                  // column lj-1 (odd parity, index q-1)
                }
                wfetch(r + 1);                                       // inputs of the next line, used next step
                const bool line_first = (r == 0);
                float pu[4], pv[4], pp[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned m = (m4 >> (8 * k)) & 0xffu;
                    const float qym = (k == 0) ? ql : qc[k - 1 < 0 ? 0 : k - 1];
                    const float a = (m & MK_XM) ? qc[k] : 0.0f;
                    const float b = ((m & MK_C) && !line_first) ? qx[k] : 0.0f;
                    const float t1 = u[k] - a;
                    pu[k] = t1 + b;
                    const float a2 = (m & MK_YM) ? qc[k] : 0.0f;
                    const float b2 = ((m & MK_C) && (w_j + k) > 0) ? qym : 0.0f;
                    const float t2 = v[k] - a2;
                    pv[k] = t2 + b2;
                    pp[k] = __fmaf_rn(P.cp, qc[k], P.Pin ? pin[k] : 0.0f);
                }
                if (P.turb > 0.0f && r >= 1 && r <= NX - 2) {        // fused addTurbulence (fluid.go:496-526)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const unsigned m = (m4 >> (8 * k)) & 0xffu;
                        const int jj = w_j + k;
                        if ((m & MK_C) && jj >= 1 && jj <= NY - 2) {
                            const float uu = pu[k] * pu[k], vv = pv[k] * pv[k];
                            const float localVel = sqrtf(uu + vv);
                            if (localVel > 0.1f) {
                                const float nu = __ldg(P.noiseU + o + k) * P.turb;
                                const float nv = __ldg(P.noiseV + o + k) * P.turb;
                                const float factor = fminf(localVel * 0.5f, 1.0f);
                                const float du = nu * factor, dv = nv * factor;
                                pu[k] = pu[k] + du;
                                pv[k] = pv[k] + dv;
                            }
                        }
                    }
                }
                store4(P.Uo + o, NY, w_j, pu);
                store4(P.Vo + o, NY, w_j, pv);
                store4(P.Po + o, NY, w_j, pp);
            } else if (t >= 1) {
                wfetch(r + 1);
            }
        }
        __syncthreads();
    }

    if (STATS && is_compute) {
        mymax = warp_max(mymax);
        if (lane == 0 && mymax > 0.0f) atomicMax(P.stats + ((P.stage0 + s) >> 1), __float_as_uint(mymax));
    }
}
