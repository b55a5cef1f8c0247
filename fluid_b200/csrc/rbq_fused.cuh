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
// (constant i, contiguous in j) stream through a ring of 35 slots in shared memory.
// 26 warps form a software pipeline WITHOUT block-wide barriers:
//   thread 768   TMA producer for the loaders: cp.async.bulk of the U, V, mask segments of a
//                line into an 6-deep staging ring (mbarrier complete_tx)
//   thread 800   TMA producer for the writers: U0, V0, mask of the owned columns, 6-deep ring
//   warps 16-19  loaders  warp 16+g takes lines == g (mod 4): staging -> -D0, neighbour count, q=0
//   warp  s<16   half sweep s (colour s&1): may process line r once its predecessor
//                (loader for s=0, warp s-1 otherwise) has finished line r+1
//   warps 20-23  writers  warp 20+g takes owned lines == g (mod 4): slot(r), slot(r-1), staging
//                -> U, V, p in global (STG.128)
// Hand-offs are per-line mbarriers (arrive = release, try_wait = acquire); the loader reuses
// a slot once the writers are past it.  Four lines are in flight in each of the serial
// roles: ablation showed one iteration costing 264 us and eight 320 us, i.e. the sweeps
// were hidden behind a one-line-at-a-time loader and writer.
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
#define RQ_THREADS 832
#define RQ_TJ_MAX 448     // multiple of 16; WL = TJ + 48 <= 496 (a lane owns 2 groups of 4 cells per line;
                          // 16-byte granules for the TMA copies of the mask)
#define RQ_STG 6          // loader staging ring depth (lines in flight through TMA)
#define RQ_WSTG 6         // writer staging ring depth
#define RQ_RING 64        // hand-off barriers per role (> RQ_NL, see rq_wait_line)
#define RQ_ROLES 18
// shared memory: 35 slots * WL * 9 B (q, -D0, neighbour count) + RQ_STG * (WL*9 + 16) B
//                + RQ_WSTG * TJ*9 B + mbarriers = 156.2 + 26.9 + 24.2 + 9.3 KB at WL = 496
__host__ __device__ __forceinline__ size_t rq_stage_bytes(int WL) { return (size_t)WL * 9 + 16; }
__host__ __device__ __forceinline__ size_t rq_wstage_bytes(int TJ) { return (size_t)TJ * 9; }
__host__ __device__ __forceinline__ size_t rq_smem_bytes(int WL, int TJ)
{
    return (size_t)RQ_NL * WL * 9 + RQ_STG * rq_stage_bytes(WL) + RQ_WSTG * rq_wstage_bytes(TJ) +
           8 * (RQ_STG + RQ_WSTG) + 8 * RQ_ROLES * RQ_RING + 64;
}

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
    int *debug;                  // [0] != 0: a pipeline wait timed out, [1..5] say which
    int xflags;                  // experiments (FLUIDB200_RBQ_X): 1 skip sweeps, 2 skip writer I/O, 4 skip TMA
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
__device__ __forceinline__ unsigned rq_s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rq_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rq_s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rq_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rq_s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rq_mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rq_s32(bar)) : "memory");
}
__device__ __forceinline__ bool rq_mbar_try(unsigned long long *bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(rq_s32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline that stops making progress must never hang the GPU.  After
// ~2^22 failed polls (hundreds of milliseconds) the first waiter records who waited for
// what in P.debug and every waiter falls through; the host turns the flag into an error.
__device__ int *rq_debug;   // set per launch (RBQ::debug)
__device__ __forceinline__ void rq_mbar_wait(unsigned long long *bar, unsigned parity, int tag = 0)
{
    for (unsigned spin = 0; !rq_mbar_try(bar, parity); spin++) {
        if ((spin & 1023u) == 1023u) {
            int *d = rq_debug;
            if (d && *reinterpret_cast<volatile int *>(d) != 0) return;     // someone already gave up: drain
            if (spin > (1u << 21)) {
                if (d && atomicCAS(d, 0, 1) == 0) {
                    d[1] = tag; d[2] = (int)threadIdx.x; d[3] = (int)blockIdx.x; d[4] = (int)blockIdx.y; d[5] = (int)parity;
                    __threadfence();
                }
                return;
            }
        }
    }
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void rq_tma_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rq_s32(dst)), "l"(src), "r"(bytes), "r"(rq_s32(bar)) : "memory");
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

// Carried across consecutive lines of one half sweep (registers): the `up` vector of
// line r is the other-parity vector of line r+1, and the other-parity vector of line r
// is the `down` vector of line r+1 (nobody writes them in between), so each line costs
// 3 x LDS.128 + 2 x LDS.32 + 1 x STS.128 per 4 cells instead of 6 x LDS.128.
struct RQCarry { float4 up[2], ot[2]; };

// One line of one half sweep: the active cells have column parity A.
template <int A, bool STATS>
__device__ __forceinline__ void rq_line(float *__restrict__ sQ, const float *__restrict__ sND,
                                        const unsigned char *__restrict__ sC, int own_row, int up_row, int dn_row,
                                        int WQ, int lane, float wd, bool row_owned, int TJ, float &mymax,
                                        RQCarry &cy, bool have)
{
    const int ngrp = WQ >> 2;
    const float2 nwd2 = make_float2(-wd, -wd);
    const float c4 = wd * 0.25f;                          // wd * (1/s) for a cell with four fluid neighbours
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
        const float4 dn = have ? cy.ot[half] : *reinterpret_cast<const float4 *>(sQ + dno + q0);
        const float4 ot = have ? cy.up[half] : *reinterpret_cast<const float4 *>(sQ + oth + q0);
        cy.up[half] = up;
        cy.ot[half] = ot;
        const float ox = sQ[oth + q0 + (A ? 4 : -1)];
        const float4 nd = *reinterpret_cast<const float4 *>(sND + own + q0);   // -D0
        const unsigned code = *reinterpret_cast<const unsigned *>(sC + own + q0);   // fluid-neighbour counts, 0 = skip
        // left / right neighbours of cell k: other-parity indices q0+k-1+A and q0+k+A
        float2 l01, l23, r01, r23;
        if (A) { l01 = make_float2(ot.x, ot.y); l23 = make_float2(ot.z, ot.w); r01 = make_float2(ot.y, ot.z); r23 = make_float2(ot.w, ox); }
        else   { l01 = make_float2(ox, ot.x);   l23 = make_float2(ot.y, ot.z); r01 = make_float2(ot.x, ot.y); r23 = make_float2(ot.z, ot.w); }
        // nb = ((q[i-1,j] + q[i+1,j]) + q[i,j-1]) + q[i,j+1];  t = nb - D0
        const float2 nb01 = __fadd2_rn(__fadd2_rn(__fadd2_rn(make_float2(dn.x, dn.y), make_float2(up.x, up.y)), l01), r01);
        const float2 nb23 = __fadd2_rn(__fadd2_rn(__fadd2_rn(make_float2(dn.z, dn.w), make_float2(up.z, up.w)), l23), r23);
        const float2 t01 = __fadd2_rn(nb01, make_float2(nd.x, nd.y));
        const float2 t23 = __fadd2_rn(nb23, make_float2(nd.z, nd.w));
        // q' = fma(wd*rs, t, fma(-wd, q, q))
        const float2 q01 = make_float2(qo.x, qo.y), q23 = make_float2(qo.z, qo.w);
        float2 c01, c23;
        if (code == 0x04040404u) {                       // the common case: four interior cells
            c01 = make_float2(c4, c4); c23 = c01;
        } else {
            float c[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned ns = (code >> (8 * k)) & 0xffu;
                const float rs = ns == 0 ? 0.0f : (ns == 1 ? 1.0f : (ns == 2 ? 0.5f : (ns == 3 ? (1.0f / 3.0f) : 0.25f)));
                c[k] = wd * rs;
            }
            c01 = make_float2(c[0], c[1]); c23 = make_float2(c[2], c[3]);
        }
        const float2 n01 = __ffma2_rn(c01, t01, __ffma2_rn(nwd2, q01, q01));
        const float2 n23 = __ffma2_rn(c23, t23, __ffma2_rn(nwd2, q23, q23));
        *reinterpret_cast<float4 *>(qown) = make_float4(n01.x, n01.y, n23.x, n23.y);
        if (STATS) {
            const float qv[4] = { qo.x, qo.y, qo.z, qo.w }, tv[4] = { t01.x, t01.y, t23.x, t23.y };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned ns = (code >> (8 * k)) & 0xffu;
                const int lj = 2 * (q0 + k) + A;
                if (ns && row_owned && lj >= RQ_H && lj < RQ_H + TJ) {
                    const float ad = fabsf(__fmaf_rn((float)ns, qv[k], -tv[k]));
                    if (ad > mymax) mymax = ad;
                }
            }
        }
    }
}

// hand-off barriers: role 0 = loader, 1+s = half sweep s, 17 = writer; RQ_RING barriers per
// role, one per line (line r uses barrier r % RQ_RING, phase parity (r / RQ_RING) & 1).
// Every role arrives for EVERY line 0 .. nproc in order, and no role can be more than
// RQ_NL lines ahead of another (the loader waits for the slot), so with RQ_RING > RQ_NL a
// parity wait always refers to the current or the immediately preceding phase.
// Every lane arrives and every lane polls: measured faster than one arrive / one poller per
// warp (lane-0 polling adds a divergent branch + __syncwarp to every hand-off: 0.35 -> 0.59 ms).
__device__ __forceinline__ void rq_done(unsigned long long *bars, int role, int line)
{
    rq_mbar_arrive(bars + role * RQ_RING + (line & (RQ_RING - 1)));
}
__device__ __forceinline__ void rq_wait_line(unsigned long long *bars, int role, int line)
{
    rq_mbar_wait(bars + role * RQ_RING + (line & (RQ_RING - 1)), (unsigned)(line / RQ_RING) & 1u, (role << 20) | line);
}
__device__ __forceinline__ void rq_wait_warp(unsigned long long *bar, unsigned parity, int tag)
{
    rq_mbar_wait(bar, parity, tag);
}

template <bool STATS>
__global__ void __launch_bounds__(RQ_THREADS, 1) k_rbq_fused(const RBQ P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int WL = P.WL, WQ = WL >> 1, ROW = WL;             // elements per slot in one plane
    const int TJ = P.TJ;
    float *sQ = reinterpret_cast<float *>(smem_raw);        // [slot][parity][q]
    float *sND = sQ + RQ_NL * WL;                            // -D0
    unsigned char *sC = reinterpret_cast<unsigned char *>(sND + RQ_NL * WL);   // fluid-neighbour count, 0 = never updated
    // loader staging ring: per slot WL floats of U, WL+4 floats of V, WL mask bytes (raw global data)
    unsigned char *stg = sC + RQ_NL * WL;                    // 16-byte aligned: WL is a multiple of 16
    const int STG = (int)rq_stage_bytes(WL);
    // writer staging ring: per slot TJ floats of U0, TJ floats of V0, TJ mask bytes (owned columns)
    unsigned char *wstg = stg + RQ_STG * STG;
    const int WSTG = (int)rq_wstage_bytes(TJ);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(wstg + RQ_WSTG * WSTG);   // RQ_STG mbarriers
    unsigned long long *wfull = full + RQ_STG;               // RQ_WSTG mbarriers
    unsigned long long *bars = wfull + RQ_WSTG;              // RQ_ROLES * RQ_RING hand-off mbarriers

    const Grid g = P.g;
    const int NX = g.NX, NY = g.NY, PIT = g.pitch;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int strip = blockIdx.x;
    const int i0c = P.ib + blockIdx.y * P.chunk;
    const int i1c = min(i0c + P.chunk, P.ie);
    if (i0c >= i1c) return;
    const int jr0 = strip * TJ - RQ_H;
    const int e0 = i0c - RQ_H, e1 = i1c + RQ_H;               // half sweeps process lines [e0, e1); e1 is loaded too
    const int nst = P.nstages;
    const int nproc = e1 - e0;                                // lines each half sweep passes over
    const int first_owned = i0c - e0, last_owned = (i1c - 1) - e0;   // relative lines the writers produce

    if (tid == 0) rq_debug = P.debug;
    for (int k = tid; k < RQ_ROLES * RQ_RING; k += RQ_THREADS) rq_mbar_init(bars + k, 32);   // every role is one warp per line
    if (tid < RQ_STG) rq_mbar_init(full + tid, 1);
    if (tid >= 32 && tid < 32 + RQ_WSTG) rq_mbar_init(wfull + (tid - 32), 1);
    // the slot "below" line e0 (relative -1) must read as q = 0
    for (int k = tid; k < WL; k += RQ_THREADS) sQ[(RQ_NL - 1) * ROW + k] = 0.0f;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    if (warp < 16) {
        // ================= half sweep `warp` =================
        const int s = warp;
        if (s >= nst) return;
        const int colour = (P.stage0 + s) & 1;
        const float wd = P.wd[s];
        float mymax = 0.0f;
        int sl = 0;
        int a = (colour + e0) & 1;
        RQCarry cy;
        bool have = false;
        for (int rel = 0; rel < nproc; rel++) {
            rq_wait_line(bars, s, rel + 1);                   // predecessor (loader or half sweep s-1) is past line rel+1
            const int r = e0 + rel;
            if (r >= 1 && r <= NX - 2 && !(P.xflags & 1)) {
                const int slp = sl + 1 == RQ_NL ? 0 : sl + 1;
                const int slm = sl == 0 ? RQ_NL - 1 : sl - 1;
                const bool row_owned = (r >= i0c) && (r < i1c);
                if (a) rq_line<1, STATS>(sQ, sND, sC, sl * ROW, slp * ROW, slm * ROW, WQ, lane, wd, row_owned, TJ, mymax, cy, have);
                else   rq_line<0, STATS>(sQ, sND, sC, sl * ROW, slp * ROW, slm * ROW, WQ, lane, wd, row_owned, TJ, mymax, cy, have);
                have = true;
            } else {
                have = false;
            }
            rq_done(bars, 1 + s, rel);
            sl = sl + 1 == RQ_NL ? 0 : sl + 1;
            a ^= 1;
        }
        rq_done(bars, 1 + s, nproc);       // line e1 is never swept: lets the next half sweep finish its last line
        if (STATS) {
            mymax = warp_max(mymax);
            if (lane == 0 && mymax > 0.0f) atomicMax(P.stats + ((P.stage0 + s) >> 1), __float_as_uint(mymax));
        }
    } else if (warp < 20) {
        // ================= loaders: warp 16+g takes lines rel == g (mod 4), lines e0 .. e1 =================
        const int grp = warp - 16;
        const int ngroups = WL >> 2;                          // float4 column groups of a line
        for (int rel = grp; rel <= nproc; rel += 4) {
            const int L = e0 + rel;
            const int sl = rel % RQ_NL;
            // only interior lines inside this rank's slab hold updatable cells; line e1 is never swept
            const bool line_live = rel < nproc && L >= 1 && L <= NX - 2 && L >= g.i_alloc0 &&
                                   L + 1 < g.i_alloc0 + g.lines_alloc;
            // slot(rel) last held line rel-NL; its last readers are the writers of lines rel-NL and rel-NL+1
            // (owned lines) or the last half sweep working on line rel-NL+1 (halo lines)
            {
                const int y = rel - RQ_NL + 1;
                if (y >= 0) {
                    if (y >= first_owned && y <= last_owned) {
                        rq_wait_line(bars, 17, y);
                        if (y - 1 >= first_owned) rq_wait_line(bars, 17, y - 1);
                    } else {
                        rq_wait_line(bars, nst, y);
                        if (y - 1 >= first_owned && y - 1 <= last_owned) rq_wait_line(bars, 17, y - 1);
                    }
                }
            }
            if (rel < nproc) {
                // lines rel and rel+1 must have landed (the producer stages every line 0 .. nproc)
                rq_wait_warp(full + (rel % RQ_STG), (rel / RQ_STG) & 1, (30 << 20) | rel);
                rq_wait_warp(full + ((rel + 1) % RQ_STG), ((rel + 1) / RQ_STG) & 1, (31 << 20) | rel);
            }
            const unsigned char *s0 = stg + (rel % RQ_STG) * STG, *s1 = stg + ((rel + 1) % RQ_STG) * STG;
            const float *stU = reinterpret_cast<const float *>(s0), *stV = stU + WL;
            const float *stU1 = reinterpret_cast<const float *>(s1);
            const unsigned char *stM = s0 + (size_t)(2 * WL + 4) * 4;
#pragma unroll 4
            for (int gi = lane; gi < ngroups; gi += 32) {
                const int j = jr0 + 4 * gi;
                float d[4] = {0.f, 0.f, 0.f, 0.f};
                unsigned code = 0;
                if (line_live && j >= 0 && j < PIT) {
                    float u0[4], u1[4], v[5];
                    unpack(*reinterpret_cast<const float4 *>(stU + 4 * gi), u0);
                    unpack(*reinterpret_cast<const float4 *>(stU1 + 4 * gi), u1);
                    unpack(*reinterpret_cast<const float4 *>(stV + 4 * gi), v);
                    v[4] = (j + 4 < PIT) ? stV[4 * gi + 4] : 0.0f;
                    const unsigned mk = *reinterpret_cast<const unsigned *>(stM + 4 * gi);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const unsigned m = (mk >> (8 * k)) & 0xffu;
                        const int jj = j + k;
                        const unsigned ns = __popc(m & 30u);
                        const bool upd = (m & MK_C) && ns > 0 && jj >= 1 && jj <= NY - 2;
                        const float dv = ((u1[k] - u0[k]) + v[k + 1]) - v[k];
                        d[k] = upd ? -dv : 0.0f;
                        if (upd) code |= ns << (8 * k);
                    }
                }
                const int q = 2 * gi;
                const int b0 = sl * ROW + q, b1 = b0 + WQ;
                *reinterpret_cast<float2 *>(sND + b0) = make_float2(d[0], d[2]);
                *reinterpret_cast<float2 *>(sND + b1) = make_float2(d[1], d[3]);
                *reinterpret_cast<float2 *>(sQ + b0) = make_float2(0.f, 0.f);
                *reinterpret_cast<float2 *>(sQ + b1) = make_float2(0.f, 0.f);
                *reinterpret_cast<unsigned short *>(sC + b0) = (unsigned short)((code & 0xffu) | ((code >> 8) & 0xff00u));
                *reinterpret_cast<unsigned short *>(sC + b1) = (unsigned short)(((code >> 8) & 0xffu) | ((code >> 16) & 0xff00u));
            }
            rq_done(bars, 0, rel);
        }
    } else if (warp < 24) {
        // ================= writers: warp 20+g takes owned lines rel == g (mod 4) =================
        const int grp = warp - 20;
        const int ngroups = TJ >> 2;
        // the hand-off phases count every line: arrive for the halo lines this warp would have taken
        for (int rel = grp; rel < first_owned; rel += 4) rq_done(bars, 17, rel);
        int rel = first_owned + ((grp - first_owned) & 3);
        for (; rel <= last_owned; rel += 4) {
            const int r = e0 + rel;
            const int n = rel - first_owned;                  // n-th owned line: writer staging slot n % RQ_WSTG
            rq_wait_warp(wfull + (n % RQ_WSTG), (n / RQ_WSTG) & 1, (29 << 20) | rel);
            rq_wait_line(bars, nst, rel);                     // last half sweep is past line r
            if (!(P.xflags & 2)) {
                const int sl = rel % RQ_NL, slm = (rel + RQ_NL - 1) % RQ_NL;
                const unsigned char *w0 = wstg + (n % RQ_WSTG) * WSTG;
                const float *wU = reinterpret_cast<const float *>(w0), *wV = wU + TJ;
                const unsigned char *wM = w0 + (size_t)TJ * 8;
                const bool line_first = (r == 0);
                const bool turb_line = P.turb > 0.0f && r >= 1 && r <= NX - 2;
#pragma unroll 2
                for (int gi = lane; gi < ngroups; gi += 32) {
                    const int w_lj = RQ_H + 4 * gi, w_j = jr0 + w_lj;
                    if (w_j >= NY) continue;
                    const int o = (r - g.i_alloc0) * PIT + w_j;
                    const int q = w_lj >> 1;
                    float u[4], v[4], pin[4] = {0.f, 0.f, 0.f, 0.f}, qc[4], qx[4], ql;
                    unpack(*reinterpret_cast<const float4 *>(wU + 4 * gi), u);
                    unpack(*reinterpret_cast<const float4 *>(wV + 4 * gi), v);
                    const unsigned m4 = *reinterpret_cast<const unsigned *>(wM + 4 * gi);
                    if (P.Pin) unpack(ld4(P.Pin + o), pin);
                    {
                        const float2 ev = *reinterpret_cast<const float2 *>(sQ + sl * ROW + q);
                        const float2 od = *reinterpret_cast<const float2 *>(sQ + sl * ROW + WQ + q);
                        qc[0] = ev.x; qc[1] = od.x; qc[2] = ev.y; qc[3] = od.y;
                        const float2 evm = *reinterpret_cast<const float2 *>(sQ + slm * ROW + q);
                        const float2 odm = *reinterpret_cast<const float2 *>(sQ + slm * ROW + WQ + q);
                        qx[0] = evm.x; qx[1] = odm.x; qx[2] = evm.y; qx[3] = odm.y;
                        ql = sQ[sl * ROW + WQ + q - 1];                  // column lj-1 (odd parity, index q-1)
                    }
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
                        pp[k] = __fmaf_rn(P.cp, qc[k], pin[k]);
                    }
                    if (turb_line) {                                     // fused addTurbulence (fluid.go:496-526)
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
                }
            }
            rq_done(bars, 17, rel);
        }
    } else if (tid == 768) {
        // ================= producer for the loaders: TMA bulk copies into the staging ring =================
        // line rel goes to staging slot rel % RQ_STG once the loaders of lines rel-RQ_STG and
        // rel-RQ_STG-1 are done (slot(x) is read for lines x-1 and x).
        const int cj0 = jr0 < 0 ? 0 : jr0;                                   // first global column copied
        const int cjU = min(jr0 + WL, PIT), cjV = min(jr0 + WL + 4, PIT);    // one past the last column (U, mask / V)
        const int off = cj0 - jr0;                                            // staging column of global column cj0
        const unsigned bU = (unsigned)(cjU - cj0) * 4, bV = (unsigned)(cjV - cj0) * 4, bM = (unsigned)(cjU - cj0);
        const unsigned bytes = bU + bV + bM;
        const int lineA = max(0, g.i_alloc0), lineB = min(NX, g.i_alloc0 + g.lines_alloc);
        const int relA = lineA - e0, relB = (cjU > cj0 && !(P.xflags & 4)) ? lineB - e0 : -1;
        const long long o0 = (long long)(e0 - g.i_alloc0) * PIT + cj0;       // offset of relative line 0 (may be negative)
        const float *gU = P.U + o0, *gV = P.V + o0;
        const unsigned char *gM = P.mask + o0;
        for (int rel = 0; rel <= nproc; rel++) {
            const int k = rel % RQ_STG;
            if (rel >= RQ_STG) {
                rq_wait_line(bars, 0, rel - RQ_STG);
                if (rel - RQ_STG - 1 >= 0) rq_wait_line(bars, 0, rel - RQ_STG - 1);
            }
            unsigned char *s0 = stg + k * STG;
            const unsigned fb = rq_s32(full + k);
            if (rel >= relA && rel < relB) {
                const unsigned dU = rq_s32(reinterpret_cast<float *>(s0) + off);
                const unsigned dV = rq_s32(reinterpret_cast<float *>(s0) + WL + off);
                const unsigned dM = rq_s32(s0 + (size_t)(2 * WL + 4) * 4 + off);
                // order prior generic-proxy reads of this staging slot before the async-proxy writes
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dU), "l"(gU), "r"(bU), "r"(fb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dV), "l"(gV), "r"(bV), "r"(fb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dM), "l"(gM), "r"(bM), "r"(fb) : "memory");
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
            }
            gU += PIT; gV += PIT; gM += PIT;
        }
    } else if (tid == 800) {
        // ================= producer for the writers: U0, V0, mask of the owned columns =================
        const int c0 = strip * TJ;                                            // first owned column (multiple of 16)
        const int c1 = min(c0 + TJ, PIT);
        const bool any = c1 > c0 && !(P.xflags & 2);
        const unsigned bF = any ? (unsigned)(c1 - c0) * 4 : 0, bM = any ? (unsigned)(c1 - c0) : 0;
        const size_t o0 = (size_t)(i0c - g.i_alloc0) * PIT + c0;
        const float *gU = P.U + o0, *gV = P.V + o0;
        const unsigned char *gM = P.mask + o0;
        const int nown = i1c - i0c;
        for (int n = 0; n < nown; n++) {
            const int k = n % RQ_WSTG;
            if (n >= RQ_WSTG) rq_wait_line(bars, 17, first_owned + n - RQ_WSTG);   // that writer is done with the slot
            unsigned char *w0 = wstg + k * WSTG;
            const unsigned fb = rq_s32(wfull + k);
            if (any) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(2 * bF + bM) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(rq_s32(w0)), "l"(gU), "r"(bF), "r"(fb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(rq_s32(w0 + (size_t)TJ * 4)), "l"(gV), "r"(bF), "r"(fb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(rq_s32(w0 + (size_t)TJ * 8)), "l"(gM), "r"(bM), "r"(fb) : "memory");
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
            }
            gU += PIT; gV += PIT; gM += PIT;
        }
    }
}
