// rbq_stream.cuh -- the pressure solve of rbq_fused.cuh (red-black SOR in pressure form, fluid.go:188-234
// in red-black order, all <= 8 iterations of a pass in ONE trip over HBM) as a REGISTER pipeline.
//
// Same arithmetic, cell for cell, as k_rbq_fused and its CPU restatement fo_project_redblack_q
// (the test oracle under oracle/): q' = fma(wd/s, nb - D0, fma(-wd, q, q)); U, V, p materialised once at the end.
// What changed is where q lives between the half sweeps:
//
//   k_rbq_fused   q circulates through a 28-line ring in shared memory; one pair of warps per iteration,
//                 hand-offs through per-line mbarriers: 6.5 warp instructions and ~720 shared-memory
//                 wavefronts per line of 512 cells; issue slots 66 %, shared-memory pipe 66 %, DRAM 30 %.
//   k_rbq_stream  a strip of 256 columns is carried through all eight iterations by TWO warps (a CTA): warp 0
//                 loads and runs iterations 0-3, warp 1 runs iterations 4-7 and writes.  A lane owns 8
//                 consecutive columns (even | odd = two float4) and keeps, per iteration, the four vectors the
//                 two half sweeps of a step need (P2, P1, F2, F1 below): 4 x 16 registers.  Iteration t+1 runs
//                 two lines behind iteration t IN THE SAME THREAD, so a value handed from one iteration to the
//                 next never leaves the register file: no ring traffic for q, no per-line hand-off barriers.
//                 Left / right neighbours beyond a lane's columns come from the next lane by shuffle.  The one
//                 hand-off (iteration 3 -> 4) is two float4 per lane and tick through a 4-deep shared-memory
//                 queue guarded by named barriers (bar.arrive / bar.sync: no polling through shared memory).
//                 Shared memory otherwise holds only what every iteration re-reads, -D0 and the neighbour count
//                 (5 B per cell, a 22-line ring; a lane reads what the lane of the same index wrote), and two
//                 small TMA staging rings (U, V, mask in; U0, V0, mask again for the write-out).
//                 CTAs never talk to each other: a strip overlaps its neighbours by the 16-column dependency
//                 cone of 16 half sweeps on either side, a chunk by 16 lines.
//
// Step r of iteration t (A = parity of the active columns = parity of line r):
//   first  = colour-0 half sweep on line r     : own qo = old[r][A], up = old[r+1][A], dn = P2 = old[r-1][A],
//                                                left / right from P1 = old[r][1-A]
//   second = colour-1 half sweep on line r-1   : own P2, up = first(r), dn = F2 = first(r-2), left / right from
//                                                F1 = first(r-1)
//   then line r-1 is final for this iteration: columns A = second, columns 1-A = F1.
// "old" is the previous iteration's output: qo is ITS F2 (before it is overwritten) and up is ITS second, both
// of the same tick -- iteration t is at step k - 2t in tick k.  Names rotate instead of values moving
// (P2 <- up, F2 <- first; the roles of the a / b registers swap every tick; the loop body is two ticks).
#pragma once
#include "rbq_fused.cuh"

#define RS_W 256                     // columns per CTA
#define RS_H 16                      // halo = dependency cone of 16 half sweeps
#define RS_TJ_MAX (RS_W - 2 * RS_H)  // 224 owned columns, 28 lanes
#define RS_HS 4                      // iterations per warp
#define RS_QD 4                      // depth of the warp 0 -> warp 1 queue (ticks); power of two
#ifndef RS_NP
#define RS_NP 11                     // ring of line PAIRS: 18 lines are alive between the loader and iteration 7, + RS_QD of slack between the warps
#endif
#ifndef RS_RSF
#define RS_RSF 1                     // 1: the ring holds 1/s as a float (branch-free update), 0: the neighbour count as a byte + a fast-path branch
#endif
#if RS_RSF
#define RS_PAIRB 4096                // bytes per pair: -D0 [line parity][column parity][128] floats, then 1/s alike
#define RS_CSTRIDE 16                // bytes per lane in the second half of a pair
#else
#define RS_PAIRB 2560                // bytes per pair: -D0 [line parity][column parity][128] floats, then the counts as bytes
#define RS_CSTRIDE 4
#endif
#ifndef RS_LST
#define RS_LST 4                     // loader staging ring (lines): two in use, two in flight; power of two
#endif
#ifndef RS_WST
#define RS_WST 4                     // writer staging ring; power of two
#endif
#define RS_LSTB 2336                 // U 1024 | V 1056 (260 floats: one column beyond) | mask 256
#define RS_WSTB (9 * RS_TJ_MAX)      // U0 | V0 | mask of the owned columns
#ifndef RS_CPS
#define RS_CPS (RS_RSF ? 3 : 4)      // CTAs per SM the kernel is compiled for
#endif
static_assert((RS_LST & (RS_LST - 1)) == 0 && (RS_WST & (RS_WST - 1)) == 0 && RS_LST >= 2 && (RS_QD & (RS_QD - 1)) == 0, "ring depths");
static_assert(2 * RS_NP > 16 + RS_QD + 1, "a line's ring slot must outlive its last reader in warp 1");
static_assert(2 * RS_HS == RQ_NIT, "two warps share the iterations of a pass");

#define RS_OFF_LSTG (RS_NP * RS_PAIRB)
#define RS_OFF_WSTG (RS_OFF_LSTG + RS_LST * RS_LSTB)
#define RS_OFF_QUEUE (RS_OFF_WSTG + RS_WST * RS_WSTB)
#define RS_OFF_BARS (RS_OFF_QUEUE + RS_QD * 1024)
#define RS_OFF_TW (RS_OFF_BARS + 8 * (RS_LST + RS_WST))
#define RS_OFF_RSLUT (RS_OFF_TW + 16 * 8 * 4)
#define RS_SMEM (RS_OFF_RSLUT + 32)

__device__ __forceinline__ void rs_wait(unsigned bar, unsigned parity, int *debug, int tag)
{
#pragma unroll 1
    for (int k = 0; k < (1 << 17); k++)      // try_wait suspends the warp for up to ~microseconds per poll: well under a second in all
        if (rq_mbar_try_a(bar, parity)) return;
    // a TMA that never lands must not hang the GPU: latch a record, fall through (the host returns FB_ERR_CUDA)
    if (debug && atomicCAS(debug, 0, 1) == 0) {
        debug[1] = tag; debug[2] = (int)threadIdx.x; debug[3] = (int)blockIdx.x; debug[4] = (int)blockIdx.y; debug[5] = (int)parity;
        __threadfence();
    }
}
__device__ __forceinline__ void rs_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void rs_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

struct RSK {                         // per-CTA constants of the tick
    const unsigned char *ring_nd, *ring_c;   // ring + 16 * lane, ring + 2048 + 4 * lane
    const float *tw;                 // [half sweep][neighbours] -> wd / s
    int lane, TJ;
    int own0, last_owned, nproc;
};

__device__ __forceinline__ float4 rs_lds128(const unsigned char *a) { return *reinterpret_cast<const float4 *>(a); }
__device__ __forceinline__ unsigned rs_lds32(const unsigned char *a) { return *reinterpret_cast<const unsigned *>(a); }

// One cell update, four same-colour cells at a time: rq_update's operations on the same values.  The neighbour vector
// that is shifted by one cell against the register pairs (left for even columns, right for odd ones) is added with
// scalar FADDs -- building misaligned pairs for FADD2 costs two moves per pair.
#if RS_RSF
typedef float4 rs_code_t;            // 1/s of the four cells (0: never updated)
#else
typedef unsigned rs_code_t;          // neighbour counts of the four cells, one per byte
#endif
template <int A>
__device__ __forceinline__ float4 rs_update(const float4 qo, const float4 up, const float4 dn, const float4 ot, const float ox, const float4 nd,
                                            const rs_code_t code, const float wd, const float nwd, const float c4, const float *__restrict__ tw,
                                            float4 &t_out)
{
    // nb = ((q[i-1,j] + q[i+1,j]) + q[i,j-1]) + q[i,j+1];  t = nb - D0
    float2 s01 = __fadd2_rn(make_float2(dn.x, dn.y), make_float2(up.x, up.y));
    float2 s23 = __fadd2_rn(make_float2(dn.z, dn.w), make_float2(up.z, up.w));
    if (A == 0) {          // even columns: left = (ox, o0, o1, o2), right = (o0, o1, o2, o3)
        s01.x = s01.x + ox; s01.y = s01.y + ot.x; s23.x = s23.x + ot.y; s23.y = s23.y + ot.z;
        s01 = __fadd2_rn(s01, make_float2(ot.x, ot.y));
        s23 = __fadd2_rn(s23, make_float2(ot.z, ot.w));
    } else {               // odd columns: left = (o0, o1, o2, o3), right = (o1, o2, o3, ox)
        s01 = __fadd2_rn(s01, make_float2(ot.x, ot.y));
        s23 = __fadd2_rn(s23, make_float2(ot.z, ot.w));
        s01.x = s01.x + ot.y; s01.y = s01.y + ot.z; s23.x = s23.x + ot.w; s23.y = s23.y + ox;
    }
    const float2 t01 = __fadd2_rn(s01, make_float2(nd.x, nd.y));
    const float2 t23 = __fadd2_rn(s23, make_float2(nd.z, nd.w));
    // q' = fma(wd*rs, t, fma(-wd, q, q))
    const float2 nw = make_float2(nwd, nwd);
    const float2 q01 = make_float2(qo.x, qo.y), q23 = make_float2(qo.z, qo.w);
    const float2 b01 = __ffma2_rn(nw, q01, q01), b23 = __ffma2_rn(nw, q23, q23);
    float2 n01, n23;
#if RS_RSF
    const float2 w2 = make_float2(wd, wd);           // wd * (1/s): the product the table of the other form holds
    n01 = __ffma2_rn(__fmul2_rn(w2, make_float2(code.x, code.y)), t01, b01);
    n23 = __ffma2_rn(__fmul2_rn(w2, make_float2(code.z, code.w)), t23, b23);
#else
    if (code == 0x04040404u) {                       // the common case: four interior cells
        const float2 cc = make_float2(c4, c4);
        n01 = __ffma2_rn(cc, t01, b01);
        n23 = __ffma2_rn(cc, t23, b23);
    } else {                                         // walls, obstacles, domain edge: wd / s from this half sweep's table
        n01 = __ffma2_rn(make_float2(tw[code & 7u], tw[(code >> 8) & 7u]), t01, b01);
        n23 = __ffma2_rn(make_float2(tw[(code >> 16) & 7u], tw[(code >> 24) & 7u]), t23, b23);
    }
#endif
    t_out = make_float4(t01.x, t01.y, t23.x, t23.y);
    return make_float4(n01.x, n01.y, n23.x, n23.y);
}

__device__ __forceinline__ unsigned rs_counts(const rs_code_t code)
{
#if RS_RSF
    const float v[4] = { code.x, code.y, code.z, code.w };
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) c |= (v[k] == 0.0f ? 0u : (unsigned)__float2int_rn(1.0f / v[k])) << (8 * k);
    return c;
#else
    return code;
#endif
}

// One step of one iteration (see the header).  qo / up come from the previous iteration (zeros for iteration 0);
// returns the second half sweep's result; F2 on return is first(r), `f2old` the F2 the step started with.
template <int A, bool STATS>
__device__ __forceinline__ float4 rs_step(const RSK &K, const RBQ &P, const int t, const float4 qo, const float4 up, float4 &P2, const float4 P1,
                                          float4 &F2, const float4 F1, const unsigned char *nd1, const unsigned char *c1, const unsigned char *nd2a,
                                          const unsigned char *c2, float4 &f2old, const int r, float &mymax)
{
    const float4 nd = rs_lds128(nd1), ndb = rs_lds128(nd2a);
#if RS_RSF
    const float4 code = rs_lds128(c1), code2 = rs_lds128(c2);
#else
    const unsigned code = rs_lds32(c1), code2 = rs_lds32(c2);
#endif
    const float ox = A ? __shfl_down_sync(0xffffffffu, P1.x, 1) : __shfl_up_sync(0xffffffffu, P1.w, 1);
    const float ox2 = A ? __shfl_down_sync(0xffffffffu, F1.x, 1) : __shfl_up_sync(0xffffffffu, F1.w, 1);
    float4 tt;
    const float4 fnow = rs_update<A>(qo, up, P2, P1, ox, nd, code, P.wd[2 * t], P.nwd[2 * t], P.c4[2 * t], K.tw + 8 * (2 * t), tt);
    if (STATS) rq_stat<true>(qo, tt, rs_counts(code), 8 * K.lane + A, r >= K.own0 && r <= K.last_owned, K.TJ, mymax);
    const float4 snow = rs_update<A>(P2, fnow, F2, F1, ox2, ndb, code2, P.wd[2 * t + 1], P.nwd[2 * t + 1], P.c4[2 * t + 1], K.tw + 8 * (2 * t + 1), tt);
    if (STATS) rq_stat<true>(P2, tt, rs_counts(code2), 8 * K.lane + A, r - 1 >= K.own0 && r - 1 <= K.last_owned, K.TJ, mymax);
    f2old = F2;
    P2 = up;
    F2 = fnow;
    return snow;
}

struct RSIO {                        // loader / writer constants of a CTA
    unsigned char *ring, *lstg, *wstg, *queue;
    unsigned b_full, b_wfull;        // shared addresses of the staging mbarriers
    const float *U, *V;
    const unsigned char *mask;
    // loader
    int jw0, PIT, off, relA, relB, live_lo, live_hi;
    unsigned bU, bV, bM;
    long long o0;
    // writer
    int i0c, nown, NX, NY, i_alloc0;
    unsigned bF, bMk;
    long long ow0;
    int *debug;
};

__device__ __forceinline__ void rs_stage_line(const RSIO &Q, const int line)      // one lane: line -> loader staging slot
{
    const int k = line & (RS_LST - 1);
    const unsigned sk = rq_s32(Q.lstg + k * RS_LSTB), fb = Q.b_full + 8u * (unsigned)k;
    if (line >= Q.relA && line < Q.relB) {
        const long long o = Q.o0 + (long long)line * Q.PIT;
        // order the generic-proxy reads of this staging slot (all lanes, __syncwarp before) ahead of the async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(Q.bU + Q.bV + Q.bM) : "memory");
        rq_tma_load(sk + 4u * (unsigned)Q.off, Q.U + o, Q.bU, fb);
        rq_tma_load(sk + 1024u + 4u * (unsigned)Q.off, Q.V + o, Q.bV, fb);
        rq_tma_load(sk + 2080u + (unsigned)Q.off, Q.mask + o, Q.bM, fb);
    } else {
        rq_arrive_a(fb);
    }
}
__device__ __forceinline__ void rs_stage_wline(const RSIO &Q, const int TJ, const int n)   // one lane: owned line n -> writer staging slot
{
    const int k = n & (RS_WST - 1);
    const unsigned sb = rq_s32(Q.wstg + k * RS_WSTB), fb = Q.b_wfull + 8u * (unsigned)k;
    if (Q.bMk) {
        const long long o = Q.ow0 + (long long)n * Q.PIT;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(2u * Q.bF + Q.bMk) : "memory");
        rq_tma_load(sb, Q.U + o, Q.bF, fb);
        rq_tma_load(sb + 4u * (unsigned)TJ, Q.V + o, Q.bF, fb);
        rq_tma_load(sb + 8u * (unsigned)TJ, Q.mask + o, Q.bMk, fb);
    } else {
        rq_arrive_a(fb);
    }
}

// Tick k of warp ROLE (0: loader + iterations 0-3, 1: iterations 4-7 + writer).  Iteration t runs step k - 2t, the
// writer takes line k - 15 from iteration 7, the loader turns the staged line k + 1 into -D0 and neighbour counts.
// po[d] = byte offset of the ring pair holding lines (k0 + 2 - 2d, k0 + 3 - 2d), k0 the even tick of the body.
// Every tick runs every iteration: steps r < 0 read the zeroed ring (count 0, -D0 0) and leave q = 0, and no tick
// of the loop reaches a step r >= nproc (the last tick is the write-out of the last owned line).
template <int A, int ROLE, bool STATS>
__device__ __forceinline__ void rs_tick(const RSK &K, const RSIO &Q, const RBQ &P, const int k, const unsigned (&po)[10], float4 (&P2)[RS_HS],
                                        float4 (&P1)[RS_HS], float4 (&F2)[RS_HS], float4 (&F1)[RS_HS], float4 &wsn, float (&mymax)[RS_HS],
                                        const int klast)
{
    const int lane = K.lane;
    const int ti = k + 2;                                    // tick index from 0
    const int qs = ti & (RS_QD - 1);                         // queue slot of this tick
    unsigned char *qp = Q.queue + qs * 1024 + 16 * lane;
    float4 qo = make_float4(0.f, 0.f, 0.f, 0.f), up = qo;
    if (ROLE == 1) {                                         // what iteration 3 handed over in ITS tick k
        rs_bar_sync(1 + qs);
        qo = *reinterpret_cast<const float4 *>(qp);
        up = *reinterpret_cast<const float4 *>(qp + 512);
        if (k + RS_QD <= klast) rs_bar_arrive(1 + RS_QD + qs);
    }
    // ---------------- four iterations ----------------
    const float4 f1last = F1[RS_HS - 1];
#pragma unroll
    for (int tl = 0; tl < RS_HS; tl++) {
        const int t = tl + ROLE * RS_HS;
        const int r = k - 2 * t;
        float4 f2old;
        const unsigned o1 = po[1 + t] + (A ? 1024u + 512u : 0u);            // line r: element A of its pair, columns A
        const unsigned o2 = A ? po[1 + t] + 512u : po[2 + t] + 1024u;        // line r - 1, columns A
#if RS_RSF
        const unsigned c1 = o1, c2 = o2;                                     // 1/s sits 2048 bytes behind -D0 (ring_c)
#else
        const unsigned c1 = po[1 + t] + (A ? 256u + 128u : 0u);
        const unsigned c2 = A ? po[1 + t] + 128u : po[2 + t] + 256u;
#endif
        const float4 snow = rs_step<A, STATS>(K, P, t, qo, up, P2[tl], P1[tl], F2[tl], F1[tl], K.ring_nd + o1, K.ring_c + c1, K.ring_nd + o2,
                                              K.ring_c + c2, f2old, r, mymax[tl]);
        qo = f2old;
        up = snow;
    }
    if (ROLE == 0) {
        // ---------------- hand iteration 3's output to warp 1 ----------------
        if (ti >= RS_QD) rs_bar_sync(1 + RS_QD + qs);        // warp 1 has read what tick ti - RS_QD left in this slot
        *reinterpret_cast<float4 *>(qp) = qo;
        *reinterpret_cast<float4 *>(qp + 512) = up;
        rs_bar_arrive(1 + qs);
        // ---------------- loader: staged line LL = k + 1 -> -D0, neighbour counts ----------------
        const int LL = k + 1;
        if (LL >= 0 && LL < K.nproc) {      // the loop's last (odd) tick may lie one past the last line
            const int st0 = LL & (RS_LST - 1), st1 = (LL + 1) & (RS_LST - 1);
            rs_wait(Q.b_full + 8u * (unsigned)st0, (unsigned)(LL / RS_LST) & 1u, Q.debug, (30 << 20) | LL);
            rs_wait(Q.b_full + 8u * (unsigned)st1, (unsigned)((LL + 1) / RS_LST) & 1u, Q.debug, (31 << 20) | LL);
            const unsigned char *s0 = Q.lstg + st0 * RS_LSTB, *s1 = Q.lstg + st1 * RS_LSTB;
            const float4 ua = *reinterpret_cast<const float4 *>(s0 + 32 * lane), ub = *reinterpret_cast<const float4 *>(s0 + 32 * lane + 16);
            const float4 na = *reinterpret_cast<const float4 *>(s1 + 32 * lane), nb = *reinterpret_cast<const float4 *>(s1 + 32 * lane + 16);
            const float4 va = *reinterpret_cast<const float4 *>(s0 + 1024 + 32 * lane), vb = *reinterpret_cast<const float4 *>(s0 + 1024 + 32 * lane + 16);
            const uint2 m8 = *reinterpret_cast<const uint2 *>(s0 + 2080 + 8 * lane);
            float vn = __shfl_down_sync(0xffffffffu, va.x, 1);
            if (lane == 31) vn = *reinterpret_cast<const float *>(s0 + 1024 + 1024);      // column jw0 + 256
            const int j0 = Q.jw0 + 8 * lane;
            const bool on = LL >= Q.live_lo && LL <= Q.live_hi && j0 >= 0 && j0 < Q.PIT;
            const unsigned clo = on ? (m8.x >> MK_CNT_SHIFT) & 0x07070707u : 0u, chi = on ? (m8.y >> MK_CNT_SHIFT) & 0x07070707u : 0u;
            const float dv0 = ((na.x - ua.x) + va.y) - va.x, dv1 = ((na.y - ua.y) + va.z) - va.y;
            const float dv2 = ((na.z - ua.z) + va.w) - va.z, dv3 = ((na.w - ua.w) + vb.x) - va.w;
            const float dv4 = ((nb.x - ub.x) + vb.y) - vb.x, dv5 = ((nb.y - ub.y) + vb.z) - vb.y;
            const float dv6 = ((nb.z - ub.z) + vb.w) - vb.z, dv7 = ((nb.w - ub.w) + vn) - vb.w;
            float4 dE, dO;
            dE.x = (clo & 0x000000ffu) ? -dv0 : 0.0f; dO.x = (clo & 0x0000ff00u) ? -dv1 : 0.0f;
            dE.y = (clo & 0x00ff0000u) ? -dv2 : 0.0f; dO.y = (clo & 0xff000000u) ? -dv3 : 0.0f;
            dE.z = (chi & 0x000000ffu) ? -dv4 : 0.0f; dO.z = (chi & 0x0000ff00u) ? -dv5 : 0.0f;
            dE.w = (chi & 0x00ff0000u) ? -dv6 : 0.0f; dO.w = (chi & 0xff000000u) ? -dv7 : 0.0f;
            // line LL = k + 1: odd tick parity -> element 1 of pair d = 1, even -> element 0 of pair d = 0
            unsigned char *pr = Q.ring + (A ? po[0] : po[1] + 1024u);
            *reinterpret_cast<float4 *>(pr + 16 * lane) = dE;
            *reinterpret_cast<float4 *>(pr + 512 + 16 * lane) = dO;
#if RS_RSF
            const float *lut = reinterpret_cast<const float *>(Q.ring + RS_OFF_RSLUT);     // count -> 1/s
            float4 rE, rO;
            rE.x = lut[clo & 7u]; rO.x = lut[(clo >> 8) & 7u]; rE.y = lut[(clo >> 16) & 7u]; rO.y = lut[clo >> 24];
            rE.z = lut[chi & 7u]; rO.z = lut[(chi >> 8) & 7u]; rE.w = lut[(chi >> 16) & 7u]; rO.w = lut[chi >> 24];
            *reinterpret_cast<float4 *>(pr + 2048 + 16 * lane) = rE;
            *reinterpret_cast<float4 *>(pr + 2048 + 512 + 16 * lane) = rO;
#else
            unsigned char *pc = Q.ring + 2048u + (A ? po[0] : po[1] + 256u);
            *reinterpret_cast<unsigned *>(pc + 4 * lane) = __byte_perm(clo, chi, 0x6420);
            *reinterpret_cast<unsigned *>(pc + 128 + 4 * lane) = __byte_perm(clo, chi, 0x7531);
#endif
            __syncwarp();                                    // every lane is done with staging slot st0
            if (lane == 0 && LL + RS_LST <= K.nproc) rs_stage_line(Q, LL + RS_LST);
        }
    } else {
        // ---------------- writer: line w of the last iteration -> U, V, p ----------------
        const int w = k - 15;
        // line w: columns A = second (up), columns 1-A = F1; line w-1: columns A = F2 before the step (qo), 1-A = last tick's second
        const float4 qE = A ? f1last : up, qO = A ? up : f1last;
        const float4 xE = A ? wsn : qo, xO = A ? qo : wsn;
        const float ql = __shfl_up_sync(0xffffffffu, qO.w, 1);
        if (w >= K.own0 && w <= K.last_owned) {
            const int n = w - K.own0, r = Q.i0c + n;
            const int ws = n & (RS_WST - 1);
            rs_wait(Q.b_wfull + 8u * (unsigned)ws, (unsigned)(n / RS_WST) & 1u, Q.debug, (32 << 20) | w);
            const int lo = lane - 2, j0 = Q.jw0 + 8 * lane;
            if (lo >= 0 && lo < (K.TJ >> 3) && j0 < Q.NY) {
                const unsigned char *sb = Q.wstg + ws * RS_WSTB;
                const float4 u0 = *reinterpret_cast<const float4 *>(sb + 32 * lo), u1 = *reinterpret_cast<const float4 *>(sb + 32 * lo + 16);
                const float4 v0 = *reinterpret_cast<const float4 *>(sb + 4 * K.TJ + 32 * lo);
                const float4 v1 = *reinterpret_cast<const float4 *>(sb + 4 * K.TJ + 32 * lo + 16);
                const uint2 m8 = *reinterpret_cast<const uint2 *>(sb + 8 * K.TJ + 8 * lo);
                const size_t o = (size_t)(r - Q.i_alloc0) * Q.PIT + j0;
                float pin[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (P.Pin) { unpack(ld4(P.Pin + o), pin); unpack(ld4(P.Pin + o + 4), pin + 4); }
                const float qc[8] = { qE.x, qO.x, qE.y, qO.y, qE.z, qO.z, qE.w, qO.w };
                const float qx[8] = { xE.x, xO.x, xE.y, xO.y, xE.z, xO.z, xE.w, xO.w };
                const float uu[8] = { u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w };
                const float vv[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
                float pu[8], pv[8], pp[8];
                const bool line_first = (r == 0);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const unsigned m = (c < 4 ? m8.x : m8.y) >> (8 * (c & 3));
                    const float qym = (c == 0) ? ql : qc[c > 0 ? c - 1 : 0];
                    const float a = (m & MK_XM) ? qc[c] : 0.0f;
                    const float b = ((m & MK_C) && !line_first) ? qx[c] : 0.0f;
                    const float t1 = uu[c] - a;
                    pu[c] = t1 + b;
                    const float a2 = (m & MK_YM) ? qc[c] : 0.0f;
                    const float b2 = ((m & MK_C) && (j0 + c) > 0) ? qym : 0.0f;
                    const float t2 = vv[c] - a2;
                    pv[c] = t2 + b2;
                    pp[c] = __fmaf_rn(P.cp, qc[c], pin[c]);
                }
                if (P.turb > 0.0f && r >= 1 && r <= Q.NX - 2) {          // fused addTurbulence (fluid.go:496-526)
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const unsigned m = (c < 4 ? m8.x : m8.y) >> (8 * (c & 3));
                        const int jj = j0 + c;
                        if ((m & MK_C) && jj >= 1 && jj <= Q.NY - 2) {
                            const float u2 = pu[c] * pu[c], v2 = pv[c] * pv[c];
                            const float localVel = sqrtf(u2 + v2);
                            if (localVel > 0.1f) {
                                const float nu = __ldg(P.noiseU + o + c) * P.turb;
                                const float nv = __ldg(P.noiseV + o + c) * P.turb;
                                const float factor = fminf(localVel * 0.5f, 1.0f);
                                const float du = nu * factor, dv = nv * factor;
                                pu[c] = pu[c] + du;
                                pv[c] = pv[c] + dv;
                            }
                        }
                    }
                }
                if (j0 + 7 < Q.NY) {
                    *reinterpret_cast<float4 *>(P.Uo + o) = make_float4(pu[0], pu[1], pu[2], pu[3]);
                    *reinterpret_cast<float4 *>(P.Uo + o + 4) = make_float4(pu[4], pu[5], pu[6], pu[7]);
                    *reinterpret_cast<float4 *>(P.Vo + o) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                    *reinterpret_cast<float4 *>(P.Vo + o + 4) = make_float4(pv[4], pv[5], pv[6], pv[7]);
                    *reinterpret_cast<float4 *>(P.Po + o) = make_float4(pp[0], pp[1], pp[2], pp[3]);
                    *reinterpret_cast<float4 *>(P.Po + o + 4) = make_float4(pp[4], pp[5], pp[6], pp[7]);
                } else {
                    for (int c = 0; c < 8 && j0 + c < Q.NY; c++) { P.Uo[o + c] = pu[c]; P.Vo[o + c] = pv[c]; P.Po[o + c] = pp[c]; }
                }
            }
            __syncwarp();
            if (lane == 0 && n + RS_WST < Q.nown) rs_stage_wline(Q, K.TJ, n + RS_WST);
        }
        wsn = up;
    }
}

template <int ROLE, bool STATS>
__device__ __forceinline__ void rs_run(const RSK &K, const RSIO &Q, const RBQ &P)
{
    float4 pa[RS_HS], pb[RS_HS], fa[RS_HS], fb[RS_HS];
    float mymax[RS_HS];
#pragma unroll
    for (int t = 0; t < RS_HS; t++) { pa[t] = pb[t] = fa[t] = fb[t] = make_float4(0.f, 0.f, 0.f, 0.f); mymax[t] = 0.0f; }
    float4 wsn = make_float4(0.f, 0.f, 0.f, 0.f);
    const int kend = K.last_owned + 15;                       // the tick that writes the last owned line
    const int klast = kend | 1;                               // the last tick of the loop (odd)
    int q1 = RS_NP - 1;                                       // ring pair of lines (k0, k0 + 1) at k0 = -2
#pragma unroll 1
    for (int k0 = -2; k0 <= kend; k0 += 2) {
        unsigned po[10];
#pragma unroll
        for (int d = 0; d < 10; d++) {
            int x = q1 + 1 - d;
            if (x < 0) x += RS_NP;
            if (x >= RS_NP) x -= RS_NP;
            po[d] = (unsigned)x * RS_PAIRB;
        }
        rs_tick<0, ROLE, STATS>(K, Q, P, k0, po, pa, pb, fa, fb, wsn, mymax, klast);
        rs_tick<1, ROLE, STATS>(K, Q, P, k0 + 1, po, pb, pa, fb, fa, wsn, mymax, klast);
        if (++q1 == RS_NP) q1 = 0;
    }
    if (STATS) {
        const int nit = P.nstages >> 1;
#pragma unroll
        for (int tl = 0; tl < RS_HS; tl++) {
            const int t = tl + ROLE * RS_HS;
            const float m = warp_max(mymax[tl]);
            if (K.lane == 0 && t < nit && m > 0.0f) atomicMax(P.stats + ((P.stage0 >> 1) + t), __float_as_uint(m));
        }
    }
}

template <bool STATS>
__global__ void __launch_bounds__(64, RS_CPS) k_rbq_stream(const RBQ P)
{
    extern __shared__ __align__(128) unsigned char rs_smem[];
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const Grid g = P.g;
    const int NX = g.NX, PIT = g.pitch;
    const int strip = blockIdx.x;
    const int i0c = P.ib + blockIdx.y * P.chunk;
    const int i1c = min(i0c + P.chunk, P.ie);
    if (i0c >= i1c) return;
    const int jw0 = strip * P.TJ - RS_H;
    const int e0 = (i0c - RS_H) & ~1, e1 = i1c + RS_H;        // even first line: tick parity == line parity == active column parity
    float *tw = reinterpret_cast<float *>(rs_smem + RS_OFF_TW);

    RSK K;
    K.ring_nd = rs_smem + 16 * lane; K.ring_c = rs_smem + 2048 + RS_CSTRIDE * lane; K.tw = tw;
    K.lane = lane; K.TJ = P.TJ;
    K.own0 = i0c - e0; K.last_owned = i1c - 1 - e0; K.nproc = e1 - e0;

    RSIO Q;
    Q.ring = rs_smem; Q.lstg = rs_smem + RS_OFF_LSTG; Q.wstg = rs_smem + RS_OFF_WSTG; Q.queue = rs_smem + RS_OFF_QUEUE;
    Q.b_full = rq_s32(rs_smem + RS_OFF_BARS); Q.b_wfull = Q.b_full + 8u * RS_LST;
    Q.U = P.U; Q.V = P.V; Q.mask = P.mask;
    Q.jw0 = jw0; Q.PIT = PIT;
    {
        const int cj0 = jw0 < 0 ? 0 : jw0;
        const int cjU = min(jw0 + RS_W, PIT), cjV = min(jw0 + RS_W + 4, PIT);
        Q.off = cj0 - jw0;
        Q.bU = (unsigned)(cjU - cj0) * 4u; Q.bV = (unsigned)(cjV - cj0) * 4u; Q.bM = (unsigned)(cjU - cj0);
        Q.relA = max(0, g.i_alloc0) - e0;
        Q.relB = cjU > cj0 ? min(NX, g.i_alloc0 + g.lines_alloc) - e0 : -1;
        Q.o0 = (long long)(e0 - g.i_alloc0) * PIT + cj0;
        // only interior lines inside this rank's slab hold updatable cells; line e1 is loaded but never swept
        Q.live_lo = max(1, g.i_alloc0) - e0;
        Q.live_hi = min(min(NX - 2, g.i_alloc0 + g.lines_alloc - 2), e1 - 1) - e0;
        const int c0 = strip * P.TJ, nc = max(0, min(P.TJ, PIT - c0));
        Q.bF = (unsigned)nc * 4u; Q.bMk = (unsigned)nc;
        Q.ow0 = (long long)(i0c - g.i_alloc0) * PIT + c0;
    }
    Q.i0c = i0c; Q.nown = i1c - i0c; Q.NX = NX; Q.NY = g.NY; Q.i_alloc0 = g.i_alloc0;
    Q.debug = P.debug;

    // ring = zeros (lines before the first read as count 0 / -D0 0), wd / s table, staging barriers
    for (int o = 16 * (int)threadIdx.x; o < RS_NP * RS_PAIRB; o += 1024) *reinterpret_cast<float4 *>(rs_smem + o) = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = threadIdx.x; k < 128; k += 64) {
        const int ns = k & 7;
        const float rs = ns == 0 ? 0.0f : (ns == 1 ? 1.0f : (ns == 2 ? 0.5f : (ns == 3 ? (1.0f / 3.0f) : 0.25f)));
        tw[k] = P.wd[k >> 3] * rs;
    }
    if (threadIdx.x < 8) {
        const int ns = threadIdx.x;
        reinterpret_cast<float *>(rs_smem + RS_OFF_RSLUT)[ns] = ns == 1 ? 1.0f : (ns == 2 ? 0.5f : (ns == 3 ? (1.0f / 3.0f) : (ns == 4 ? 0.25f : 0.0f)));
    }
    if (threadIdx.x < RS_LST + RS_WST) rq_mbar_init(reinterpret_cast<unsigned long long *>(rs_smem + RS_OFF_BARS) + threadIdx.x, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (role == 0) {
        if (lane == 0)
            for (int line = 0; line < RS_LST && line <= K.nproc; line++) rs_stage_line(Q, line);
        rs_run<0, STATS>(K, Q, P);
    } else {
        if (lane == 0)
            for (int n = 0; n < RS_WST && n < Q.nown; n++) rs_stage_wline(Q, P.TJ, n);
        rs_run<1, STATS>(K, Q, P);
    }
}
