// rbq_fused.cuh -- the fastest pressure solve: red-black SOR in PRESSURE FORM, every
// iteration of a pass (<= 8 iterations = 16 half sweeps) fused in one trip over HBM.
//
// The state that circulates is one scalar per cell:
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
// Structure: a CTA owns `chunk` lines x TJ columns (+16-cell halo, recomputed).  Lines
// (constant i, contiguous in j) stream through a ring of 35 slots in 220 KB of shared
// memory.  24 warps form a software pipeline WITHOUT block-wide barriers:
//   warps 16-19  loader   global U,V,mask -> -D0, 1/s, q=0 in slot(L)       (prefetch 1 line)
//   warp  s<16   half sweep s (colour s&1): may process line r once its predecessor
//                (loader for s=0, warp s-1 otherwise) has finished line r+1
//   warps 20-23  writer   slot(r), slot(r-1) + U0,V0 -> U,V,p in global     (prefetch 1 line)
// Each role publishes the last line it finished with st.release.cta and waits on its
// predecessor with ld.acquire.cta; the loader reuses a slot once the writer is past it.
// Even and odd columns live in separate arrays so one colour is contiguous: a lane
// updates 2 x 4 consecutive same-colour cells with LDS.128 / STS.128 and packed
// FADD2 / FFMA2 (sm_100a fp32x2, bit-identical to the scalar operations).
//
// Why this shape: ncu showed the face form (rb_fused.cuh) issue-bound at ~110
// instructions per cell update and the first pressure-form version stalled on its
// per-line __syncthreads (barrier = 3.4 of 9 stall cycles per issue).
#pragma once
#include "kernels.cuh"
#include "advect_fused.cuh"

#define RQ_NL 35          // line slots
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

__device__ __forceinline__ int rq_ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void rq_st_release(int *p, int v)
{
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// all lanes of the warp return once *p >= need
__device__ __forceinline__ void rq_wait_ge(const int *p, int need, int lane)
{
    if (lane == 0) {
        while (rq_ld_acquire(p) < need) __nanosleep(20);
    }
    __syncwarp();
}
__device__ __forceinline__ void rq_publish(int *p, int v, int lane)
{
    __syncwarp();
    if (lane == 0) rq_st_release(p, v);
}

// One line of one half sweep: the active cells have column parity A.
template <int A, bool STATS>
__device__ __forceinline__ void rq_line(float *__restrict__ sQ, const float *__restrict__ sND, const float *__restrict__ sR,
                                        const unsigned char *__restrict__ sM, int own_row, int up_row, int dn_row,
                                        int WQ, int lane, float wd, bool row_owned, int TJ, float &mymax)
{
    const int ngrp = WQ >> 2;
    const float2 wd2 = make_float2(wd, wd), nwd2 = make_float2(-wd, -wd);
    const int own = own_row + A * WQ, oth = own_row + (1 - A) * WQ;
    const int upo = up_row + A * WQ, dno = dn_row + A * WQ;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int g = lane + 32 * half;
        if (g >= ngrp) continue;
        const int q0 = 4 * g;
        float *qown = sQ + own + q0;
        const float4 qo = *reinterpret_cast<const float4 *>(qown);
        const float4 up = *reinterpret_cast<const float4 *>(sQ + upo + q0);
        const float4 dn = *reinterpret_cast<const float4 *>(sQ + dno + q0);
        const float4 ot = *reinterpret_cast<const float4 *>(sQ + oth + q0);
        const float ox = sQ[oth + q0 + (A ? 4 : -1)];
        const float4 nd = *reinterpret_cast<const float4 *>(sND + own + q0);   // -D0
        const float4 rs = *reinterpret_cast<const float4 *>(sR + own + q0);
        // left / right neighbours of cell k: other-parity indices q0+k-1+A and q0+k+A
        float2 l01, l23, r01, r23;
        if (A) { l01 = make_float2(ot.x, ot.y); l23 = make_float2(ot.z, ot.w); r01 = make_float2(ot.y, ot.z); r23 = make_float2(ot.w, ox); }
        else   { l01 = make_float2(ox, ot.x);   l23 = make_float2(ot.y, ot.z); r01 = make_float2(ot.x, ot.y); r23 = make_float2(ot.z, ot.w); }
        // nb = ((q[i-1,j] + q[i+1,j]) + q[i,j-1]) + q[i,j+1];  t = nb - D0
        float2 nb01 = __fadd2_rn(__fadd2_rn(__fadd2_rn(make_float2(dn.x, dn.y), make_float2(up.x, up.y)), l01), r01);
        float2 nb23 = __fadd2_rn(__fadd2_rn(__fadd2_rn(make_float2(dn.z, dn.w), make_float2(up.z, up.w)), l23), r23);
        const float2 t01 = __fadd2_rn(nb01, make_float2(nd.x, nd.y));
        const float2 t23 = __fadd2_rn(nb23, make_float2(nd.z, nd.w));
        // q' = fma(wd*rs, t, fma(-wd, q, q))
        const float2 q01 = make_float2(qo.x, qo.y), q23 = make_float2(qo.z, qo.w);
        const float2 c01 = __fmul2_rn(wd2, make_float2(rs.x, rs.y)), c23 = __fmul2_rn(wd2, make_float2(rs.z, rs.w));
        const float2 n01 = __ffma2_rn(c01, t01, __ffma2_rn(nwd2, q01, q01));
        const float2 n23 = __ffma2_rn(c23, t23, __ffma2_rn(nwd2, q23, q23));
        *reinterpret_cast<float4 *>(qown) = make_float4(n01.x, n01.y, n23.x, n23.y);
        if (STATS) {
            const unsigned mk = *reinterpret_cast<const unsigned *>(sM + own + q0);
            const float qv[4] = { qo.x, qo.y, qo.z, qo.w }, tv[4] = { t01.x, t01.y, t23.x, t23.y };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned m = (mk >> (8 * k)) & 0xffu;
                const int lj = 2 * (q0 + k) + A;
                if ((m & MK_C) && row_owned && lj >= RQ_H && lj < RQ_H + TJ) {
                    const float ns = (float)__popc(m & 30u);
                    const float ad = fabsf(__fmaf_rn(ns, qv[k], -tv[k]));
                    if (ad > mymax) mymax = ad;
                }
            }
        }
    }
}

template <bool STATS>
__global__ void __launch_bounds__(RQ_THREADS, 1) k_rbq_fused(const RBQ P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int prog[20];     // [0] loader, [1+s] half sweep s, [17] writer: last line (relative) finished
    const int WL = P.WL, WQ = WL >> 1, ROW = WL;             // floats per slot in one plane
    float *sQ = reinterpret_cast<float *>(smem_raw);        // [slot][parity][q]
    float *sND = sQ + RQ_NL * WL;                            // -D0
    float *sR = sND + RQ_NL * WL;                            // 1/s (0: never updated)
    unsigned char *sM = reinterpret_cast<unsigned char *>(sR + RQ_NL * WL);

    const Grid g = P.g;
    const int NX = g.NX, NY = g.NY, PIT = g.pitch;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int strip = blockIdx.x;
    const int i0c = P.ib + blockIdx.y * P.chunk;
    const int i1c = min(i0c + P.chunk, P.ie);
    if (i0c >= i1c) return;
    const int jr0 = strip * P.TJ - RQ_H;
    const int e0 = i0c - RQ_H, e1 = i1c + RQ_H;               // half sweeps process lines [e0, e1); e1 is loaded too
    const int nst = P.nstages;
    const int nproc = e1 - e0;                                // lines each half sweep passes over

    if (tid < 20) prog[tid] = -1;
    __syncthreads();

    if (warp < 16) {
        // ================= half sweep `warp` =================
        const int s = warp;
        if (s >= nst) return;
        const int colour = (P.stage0 + s) & 1;
        const float wd = P.wd[s];
        const int *pred = &prog[s];                           // loader (s == 0) or half sweep s-1
        int *mine = &prog[1 + s];
        float mymax = 0.0f;
        int sl = 0;
        int a = (colour + e0) & 1;
        for (int rel = 0; rel < nproc; rel++) {
            rq_wait_ge(pred, rel + 1, lane);
            const int r = e0 + rel;
            if (r >= 1 && r <= NX - 2) {
                const int slp = sl + 1 == RQ_NL ? 0 : sl + 1;
                const int slm = sl == 0 ? RQ_NL - 1 : sl - 1;
                const bool row_owned = (r >= i0c) && (r < i1c);
                if (a) rq_line<1, STATS>(sQ, sND, sR, sM, sl * ROW, slp * ROW, slm * ROW, WQ, lane, wd, row_owned, P.TJ, mymax);
                else   rq_line<0, STATS>(sQ, sND, sR, sM, sl * ROW, slp * ROW, slm * ROW, WQ, lane, wd, row_owned, P.TJ, mymax);
            }
            rq_publish(mine, rel, lane);
            sl = sl + 1 == RQ_NL ? 0 : sl + 1;
            a ^= 1;
        }
        rq_publish(mine, nproc, lane);     // line e1 is never swept: lets the next half sweep finish its last line
        if (STATS) {
            mymax = warp_max(mymax);
            if (lane == 0 && mymax > 0.0f) atomicMax(P.stats + ((P.stage0 + s) >> 1), __float_as_uint(mymax));
        }
    } else if (warp < 20) {
        // ================= loader: lines e0 .. e1 =================
        const int ld = tid - 512;
        const bool active = ld < (WL >> 2);
        const int j = jr0 + 4 * ld;
        auto fetch = [&](int L, RQLine &x) {
            x.u = x.u1 = x.v = make_float4(0.f, 0.f, 0.f, 0.f);
            x.v4 = 0.0f; x.m = 0;
            if (!active || L < 1 || L > NX - 2 || j < 0 || j >= PIT) return;       // only interior lines hold updatable cells
            if (L < g.i_alloc0 || L + 1 >= g.i_alloc0 + g.lines_alloc) return;     // outside this rank's slab
            const int o = (L - g.i_alloc0) * PIT + j;
            x.m = __ldg(reinterpret_cast<const unsigned *>(P.mask + o));
            x.u = ld4(P.U + o);
            x.u1 = ld4(P.U + o + PIT);
            x.v = ld4(P.V + o);
            if (j + 4 < PIT) x.v4 = __ldg(P.V + o + 4);
        };
        auto commit = [&](int sl, const RQLine &x) {
            if (!active) return;
            float u0[4], u1[4], v[5], d[4], r[4];
            unpack(x.u, u0); unpack(x.u1, u1); unpack(x.v, v); v[4] = x.v4;
            unsigned mk = x.m;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned m = (mk >> (8 * k)) & 0xffu;
                const int jj = j + k;
                const int ns = __popc(m & 30u);
                const bool upd = (m & MK_C) && ns > 0 && jj >= 1 && jj <= NY - 2;
                const float dv = ((u1[k] - u0[k]) + v[k + 1]) - v[k];
                d[k] = upd ? -dv : 0.0f;
                r[k] = !upd ? 0.0f : (ns == 1 ? 1.0f : (ns == 2 ? 0.5f : (ns == 3 ? (1.0f / 3.0f) : 0.25f)));
                if (!upd) mk &= ~(0xffu << (8 * k));
            }
            const int q = 2 * ld;
            const int b0 = sl * ROW + q, b1 = b0 + WQ;
            *reinterpret_cast<float2 *>(sND + b0) = make_float2(d[0], d[2]);
            *reinterpret_cast<float2 *>(sND + b1) = make_float2(d[1], d[3]);
            *reinterpret_cast<float2 *>(sR + b0) = make_float2(r[0], r[2]);
            *reinterpret_cast<float2 *>(sR + b1) = make_float2(r[1], r[3]);
            *reinterpret_cast<float2 *>(sQ + b0) = make_float2(0.f, 0.f);
            *reinterpret_cast<float2 *>(sQ + b1) = make_float2(0.f, 0.f);
            if (STATS) {   // updatable-cell bytes, only for the residual statistics
                *reinterpret_cast<unsigned short *>(sM + b0) = (unsigned short)((mk & 0xffu) | ((mk >> 8) & 0xff00u));
                *reinterpret_cast<unsigned short *>(sM + b1) = (unsigned short)(((mk >> 8) & 0xffu) | ((mk >> 16) & 0xff00u));
            }
        };
        if (active) {   // the slot "below" line e0 (relative -1) must read as q = 0
            const int q = 2 * ld, b0 = (RQ_NL - 1) * ROW + q;
            *reinterpret_cast<float2 *>(sQ + b0) = make_float2(0.f, 0.f);
            *reinterpret_cast<float2 *>(sQ + b0 + WQ) = make_float2(0.f, 0.f);
        }
        RQLine lnA, lnB;
        fetch(e0, lnA);
        int sl = 0;
        const int *wprog = &prog[17];
        for (int rel = 0; rel <= nproc; rel++) {
            // slot(rel) last held line rel-NL, which the writer reads while writing rel-NL and rel-NL+1
            if (rel >= RQ_NL - 1) rq_wait_ge(wprog, rel - RQ_NL + 1, lane);
            if (rel & 1) { fetch(e0 + rel + 1, lnA); commit(sl, lnB); }
            else         { fetch(e0 + rel + 1, lnB); commit(sl, lnA); }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (tid == 512) rq_st_release(&prog[0], rel);
            sl = sl + 1 == RQ_NL ? 0 : sl + 1;
        }
    } else {
        // ================= writer: owned lines -> U, V, p =================
        const int st = tid - 640;
        const bool active = st < (P.TJ >> 2);
        const int w_lj = RQ_H + 4 * st, w_j = jr0 + w_lj;
        const bool col_ok = active && w_j < NY;
        float4 wU = make_float4(0, 0, 0, 0), wV = wU, wP = wU;
        unsigned wM = 0;
        auto wfetch = [&](int r) {
            if (!col_ok || r < i0c || r >= i1c) return;
            const int o = (r - g.i_alloc0) * PIT + w_j;
            wU = ld4(P.U + o);
            wV = ld4(P.V + o);
            wM = __ldg(reinterpret_cast<const unsigned *>(P.mask + o));
            if (P.Pin) wP = ld4(P.Pin + o);
        };
        const int *last = &prog[nst];                         // half sweep nst-1
        wfetch(e0);
        int sl = 0;
        for (int rel = 0; rel < nproc; rel++) {
            const int r = e0 + rel;
            rq_wait_ge(last, rel, lane);
            if (col_ok && r >= i0c && r < i1c) {
                const int o = (r - g.i_alloc0) * PIT + w_j;
                const int slm = sl == 0 ? RQ_NL - 1 : sl - 1;
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
                wfetch(r + 1);                                       // inputs of the next line
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
            } else {
                wfetch(r + 1);
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (tid == 640) rq_st_release(&prog[17], rel);
            sl = sl + 1 == RQ_NL ? 0 : sl + 1;
        }
    }
}
