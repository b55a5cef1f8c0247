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
//   k_rbq_stream  a strip of 256 columns is carried through all eight iterations by FOUR warps (a CTA), two
//                 iterations each: warp 0 also loads, warp 3 also writes.  A lane owns 8 consecutive columns
//                 (even | odd = two float4) and keeps, per iteration, the four vectors the two half sweeps of
//                 a step need (P2, P1, F2, F1 below): 2 x 16 registers.  Iteration t+1 runs two lines behind
//                 iteration t IN THE SAME THREAD, so a value handed from one iteration to the next never leaves
//                 the register file; left / right neighbours beyond a lane's columns come from the next lane by
//                 shuffle.  Between warps the hand-off is two float4 per lane and tick through a 2-deep
//                 shared-memory queue guarded by named barriers (bar.arrive / bar.sync: the waiting warp sleeps
//                 in hardware, nothing polls shared memory).  Shared memory otherwise holds only what every
//                 iteration re-reads, -D0 and 1/s (8 B per cell, a 24-line ring; a lane reads what the lane of
//                 the same index wrote), and two small TMA staging rings (U, V, mask in; U0, V0, mask again for
//                 the write-out).  CTAs never talk to each other: a strip overlaps its neighbours by the
//                 16-column dependency cone of 16 half sweeps on either side, a chunk by 16 lines.
//   Round-2 history (4098^2, ncu): two warps x four iterations, 199 registers, 6 warps per SM: 0.248 ms -- 5.3 warp
//   instructions per cell (the loader's and the writer's per-cell selects were two thirds of them), one instruction
//   issued per warp every 4.07 cycles (fixed-latency dependencies) and only 1.5 warps per scheduler to cover them.
//   Hence: four warps x two iterations (12 warps per SM), straight-line loader / writer bodies for lanes whose eight
//   cells are all interior fluid, suspend-time hints on the TMA waits.
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
#define RS_NW 4                      // warps per CTA (roles 0 .. 3)
#define RS_HS 2                      // iterations per warp
#define RS_QD 2                      // depth of a warp -> warp queue (ticks); power of two
#define RS_NP 12                     // ring of line PAIRS: 18 lines are alive between the loader and iteration 7, + (RS_NW - 1) * RS_QD of slack
#define RS_PAIRB 4096                // bytes per pair: -D0 [line parity][column parity][128] floats, then 1/s alike
#ifndef RS_LST
#define RS_LST 4                     // loader staging ring (lines): two in use, two in flight; power of two
#endif
#ifndef RS_WST
#define RS_WST 4                     // writer staging ring; power of two
#endif
#define RS_LSTB 2336                 // U 1024 | V 1056 (260 floats: one column beyond) | mask 256
#define RS_WSTB (9 * RS_TJ_MAX)      // U0 | V0 | mask of the owned columns
#ifndef RS_CPS
#define RS_CPS 3                     // CTAs per SM the kernel is compiled for
#endif
#define RS_THREADS (32 * RS_NW)
static_assert((RS_LST & (RS_LST - 1)) == 0 && (RS_WST & (RS_WST - 1)) == 0 && RS_LST >= 2 && (RS_QD & (RS_QD - 1)) == 0, "ring depths");
static_assert(2 * RS_NP > 16 + (RS_NW - 1) * RS_QD + 1, "a line's ring slot must outlive its last reader in the last warp");
static_assert(RS_NW * RS_HS == RQ_NIT, "the warps share the iterations of a pass");
static_assert(1 + 4 * (RS_NW - 1) <= 16 && RS_QD == 2, "named barriers: full / empty x RS_QD per hand-off");

#define RS_OFF_LSTG (RS_NP * RS_PAIRB)
#define RS_OFF_WSTG (RS_OFF_LSTG + RS_LST * RS_LSTB)
#define RS_OFF_QUEUE (RS_OFF_WSTG + RS_WST * RS_WSTB)
#define RS_OFF_BARS (RS_OFF_QUEUE + (RS_NW - 1) * RS_QD * 1024)
#define RS_OFF_TW (RS_OFF_BARS + 8 * (RS_LST + RS_WST))
#define RS_OFF_RSLUT (RS_OFF_TW + 16 * 8 * 4)
#define RS_SMEM (RS_OFF_RSLUT + 32)

__device__ __forceinline__ void rs_wait(unsigned bar, unsigned parity, int *debug, int tag)
{
#pragma unroll 1
    for (int k = 0; k < (1 << 17); k++)      // try_wait sleeps up to its suspend-time hint per poll: seconds in all at most
        if (rq_mbar_try_a(bar, parity)) return;
    // a TMA that never lands must not hang the GPU: latch a record, fall through (the host returns FB_ERR_CUDA)
    if (debug && atomicCAS(debug, 0, 1) == 0) {
        debug[1] = tag; debug[2] = (int)threadIdx.x; debug[3] = (int)blockIdx.x; debug[4] = (int)blockIdx.y; debug[5] = (int)parity;
        __threadfence();
    }
}
__device__ __forceinline__ void rs_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void rs_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
// hand-off r (warp r -> warp r + 1), queue slot s: ids 1 .. 12
__device__ __forceinline__ int rs_bar_full(int r, int s) { return 1 + 4 * r + s; }
__device__ __forceinline__ int rs_bar_empty(int r, int s) { return 3 + 4 * r + s; }

struct RSK {                         // per-CTA constants of the tick
    const unsigned char *ring_nd, *ring_c;   // ring + 16 * lane, ring + 2048 + 16 * lane
    int lane, TJ;
    int own0, last_owned, nproc;
};

__device__ __forceinline__ float4 rs_lds128(const unsigned char *a) { return *reinterpret_cast<const float4 *>(a); }

// One cell update, four same-colour cells at a time: rq_update's operations on the same values.  The neighbour vector
// that is shifted by one cell against the register pairs (left for even columns, right for odd ones) is added with
// scalar FADDs -- building misaligned pairs for FADD2 costs two moves per pair.  rs = 1/s of the four cells (0: never
// updated); wd * rs is the product the wd / s table of k_rbq_fused holds.
template <int A>
__device__ __forceinline__ float4 rs_update(const float4 qo, const float4 up, const float4 dn, const float4 ot, const float ox, const float4 nd,
                                            const float4 rs, const float wd, const float nwd, float4 &t_out)
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
    const float2 nw = make_float2(nwd, nwd), w2 = make_float2(wd, wd);
    const float2 q01 = make_float2(qo.x, qo.y), q23 = make_float2(qo.z, qo.w);
    const float2 b01 = __ffma2_rn(nw, q01, q01), b23 = __ffma2_rn(nw, q23, q23);
    const float2 n01 = __ffma2_rn(__fmul2_rn(w2, make_float2(rs.x, rs.y)), t01, b01);
    const float2 n23 = __ffma2_rn(__fmul2_rn(w2, make_float2(rs.z, rs.w)), t23, b23);
    t_out = make_float4(t01.x, t01.y, t23.x, t23.y);
    return make_float4(n01.x, n01.y, n23.x, n23.y);
}

__device__ __forceinline__ unsigned rs_counts(const float4 rs)       // 1/s -> s, one per byte (statistics only)
{
    const float v[4] = { rs.x, rs.y, rs.z, rs.w };
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) c |= (v[k] == 0.0f ? 0u : (unsigned)__float2int_rn(1.0f / v[k])) << (8 * k);
    return c;
}

// One step of one iteration (see the header).  qo / up come from the previous iteration (zeros for iteration 0);
// returns the second half sweep's result; F2 on return is first(r), `f2old` the F2 the step started with.
template <int A, bool STATS>
__device__ __forceinline__ float4 rs_step(const RSK &K, const RBQ &P, const int t, const float4 qo, const float4 up, float4 &P2, const float4 P1,
                                          float4 &F2, const float4 F1, const unsigned o1, const unsigned o2, float4 &f2old, const int r, float &mymax)
{
    const float4 nd = rs_lds128(K.ring_nd + o1), ndb = rs_lds128(K.ring_nd + o2);
    const float4 rs1 = rs_lds128(K.ring_c + o1), rs2 = rs_lds128(K.ring_c + o2);
    const float ox = A ? __shfl_down_sync(0xffffffffu, P1.x, 1) : __shfl_up_sync(0xffffffffu, P1.w, 1);
    const float ox2 = A ? __shfl_down_sync(0xffffffffu, F1.x, 1) : __shfl_up_sync(0xffffffffu, F1.w, 1);
    float4 tt;
    const float4 fnow = rs_update<A>(qo, up, P2, P1, ox, nd, rs1, P.wd[2 * t], P.nwd[2 * t], tt);
    if (STATS) rq_stat<true>(qo, tt, rs_counts(rs1), 8 * K.lane + A, r >= K.own0 && r <= K.last_owned, K.TJ, mymax);
    const float4 snow = rs_update<A>(P2, fnow, F2, F1, ox2, ndb, rs2, P.wd[2 * t + 1], P.nwd[2 * t + 1], tt);
    if (STATS) rq_stat<true>(P2, tt, rs_counts(rs2), 8 * K.lane + A, r - 1 >= K.own0 && r - 1 <= K.last_owned, K.TJ, mymax);
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

// ---------------- loader: staged line LL -> -D0 and 1/s in the ring slot at `pr` ----------------
__device__ __forceinline__ void rs_load_line(const RSK &K, const RSIO &Q, const int LL, unsigned char *pr)
{
    const int lane = K.lane;
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
    // -div = -(((U[i+1,j] - U[i,j]) + V[i,j+1]) - V[i,j]), the reference's operations (fluid.go:207) and a sign flip
    const float2 x01 = __fadd2_rn(make_float2(na.x, na.y), make_float2(-ua.x, -ua.y)), x23 = __fadd2_rn(make_float2(na.z, na.w), make_float2(-ua.z, -ua.w));
    const float2 x45 = __fadd2_rn(make_float2(nb.x, nb.y), make_float2(-ub.x, -ub.y)), x67 = __fadd2_rn(make_float2(nb.z, nb.w), make_float2(-ub.z, -ub.w));
    const float dv0 = (x01.x + va.y) - va.x, dv1 = (x01.y + va.z) - va.y, dv2 = (x23.x + va.w) - va.z, dv3 = (x23.y + vb.x) - va.w;
    const float dv4 = (x45.x + vb.y) - vb.x, dv5 = (x45.y + vb.z) - vb.y, dv6 = (x67.x + vb.w) - vb.z, dv7 = (x67.y + vn) - vb.w;
    float4 dE, dO, rE, rO;
    if (clo == 0x04040404u && chi == 0x04040404u) {          // eight interior fluid cells: nothing to select
        dE = make_float4(-dv0, -dv2, -dv4, -dv6); dO = make_float4(-dv1, -dv3, -dv5, -dv7);
        rE = rO = make_float4(0.25f, 0.25f, 0.25f, 0.25f);
    } else {
        dE.x = (clo & 0x000000ffu) ? -dv0 : 0.0f; dO.x = (clo & 0x0000ff00u) ? -dv1 : 0.0f;
        dE.y = (clo & 0x00ff0000u) ? -dv2 : 0.0f; dO.y = (clo & 0xff000000u) ? -dv3 : 0.0f;
        dE.z = (chi & 0x000000ffu) ? -dv4 : 0.0f; dO.z = (chi & 0x0000ff00u) ? -dv5 : 0.0f;
        dE.w = (chi & 0x00ff0000u) ? -dv6 : 0.0f; dO.w = (chi & 0xff000000u) ? -dv7 : 0.0f;
        const float *lut = reinterpret_cast<const float *>(Q.ring + RS_OFF_RSLUT);     // count -> 1/s
        rE.x = lut[clo & 7u]; rO.x = lut[(clo >> 8) & 7u]; rE.y = lut[(clo >> 16) & 7u]; rO.y = lut[clo >> 24];
        rE.z = lut[chi & 7u]; rO.z = lut[(chi >> 8) & 7u]; rE.w = lut[(chi >> 16) & 7u]; rO.w = lut[chi >> 24];
    }
    *reinterpret_cast<float4 *>(pr + 16 * lane) = dE;
    *reinterpret_cast<float4 *>(pr + 512 + 16 * lane) = dO;
    *reinterpret_cast<float4 *>(pr + 2048 + 16 * lane) = rE;
    *reinterpret_cast<float4 *>(pr + 2048 + 512 + 16 * lane) = rO;
    __syncwarp();                                    // every lane is done with staging slot st0
    if (lane == 0 && LL + RS_LST <= K.nproc) rs_stage_line(Q, LL + RS_LST);
}

// ---------------- writer: final q of line w (qE | qO) and of line w - 1 (xE | xO) -> U, V, p ----------------
__device__ __forceinline__ void rs_write_line(const RSK &K, const RSIO &Q, const RBQ &P, const int w, const float4 qE, const float4 qO,
                                              const float4 xE, const float4 xO, const float ql)
{
    const int lane = K.lane;
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
        const unsigned all = MK_C | MK_XM | MK_YM;
        if ((m8.x & (all * 0x01010101u)) == all * 0x01010101u && (m8.y & (all * 0x01010101u)) == all * 0x01010101u && r != 0 && j0 > 0) {
            // eight fluid cells with fluid on their -x and -y side: every select of the general form is taken
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float qym = (c == 0) ? ql : qc[c > 0 ? c - 1 : 0];
                const float t1 = uu[c] - qc[c];
                pu[c] = t1 + qx[c];
                const float t2 = vv[c] - qc[c];
                pv[c] = t2 + qym;
            }
        } else {
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
            }
        }
#pragma unroll
        for (int c = 0; c < 8; c++) pp[c] = __fmaf_rn(P.cp, qc[c], pin[c]);
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

// Tick k of warp ROLE (iterations 2 ROLE and 2 ROLE + 1; role 0 also loads, role RS_NW - 1 also writes).  Iteration t runs
// step k - 2t, the writer takes line k - 15 from iteration 7, the loader turns the staged line k + 1 into -D0 and 1/s.
// pb = byte offset of the ring pair that holds lines (k0, k0 + 1), k0 the even tick of the body; older pairs lie below
// it (modulo the ring).  Every tick runs every iteration: steps r < 0 read the zeroed ring (1/s = 0, -D0 = 0) and leave
// q = 0, and no tick of the loop reaches a step r >= nproc (the last tick is the write-out of the last owned line).
template <int A, int ROLE, bool STATS>
__device__ __forceinline__ void rs_tick(const RSK &K, const RSIO &Q, const RBQ &P, const int k, const unsigned (&po)[RS_HS + 3], float4 (&P2)[RS_HS],
                                        float4 (&P1)[RS_HS], float4 (&F2)[RS_HS], float4 (&F1)[RS_HS], float4 &wsn, float (&mymax)[RS_HS],
                                        const int klast)
{
    const int lane = K.lane;
    const int ti = k + 2;                                    // tick index from 0
    const int qs = ti & (RS_QD - 1);                         // queue slot of this tick
    float4 qo = make_float4(0.f, 0.f, 0.f, 0.f), up = qo;
    if (ROLE > 0) {                                          // what the previous warp handed over in ITS tick k
        const unsigned char *qp = Q.queue + ((ROLE - 1) * RS_QD + qs) * 1024 + 16 * lane;
        rs_bar_sync(rs_bar_full(ROLE - 1, qs));
        qo = *reinterpret_cast<const float4 *>(qp);
        up = *reinterpret_cast<const float4 *>(qp + 512);
        if (k + RS_QD <= klast) rs_bar_arrive(rs_bar_empty(ROLE - 1, qs));
    }
    // ---------------- two iterations ----------------
    // po[0]: pair of lines (k0 + 2, k0 + 3) (the loader's, role 0 only); po[1 + tl]: pair of lines (k0 - 2t, k0 - 2t + 1)
    const float4 f1last = F1[RS_HS - 1];
#pragma unroll
    for (int tl = 0; tl < RS_HS; tl++) {
        const int t = tl + ROLE * RS_HS;
        const int r = k - 2 * t;
        float4 f2old;
        const unsigned o1 = po[1 + tl] + (A ? 1024u + 512u : 0u);           // line r: element A of its pair, columns A
        const unsigned o2 = A ? po[1 + tl] + 512u : po[2 + tl] + 1024u;      // line r - 1, columns A
        const float4 snow = rs_step<A, STATS>(K, P, t, qo, up, P2[tl], P1[tl], F2[tl], F1[tl], o1, o2, f2old, r, mymax[tl]);
        qo = f2old;
        up = snow;
    }
    if (ROLE < RS_NW - 1) {
        // ---------------- hand the last iteration's output to the next warp ----------------
        unsigned char *qp = Q.queue + (ROLE * RS_QD + qs) * 1024 + 16 * lane;
        if (ti >= RS_QD) rs_bar_sync(rs_bar_empty(ROLE, qs));   // the next warp has read what tick ti - RS_QD left in this slot
        *reinterpret_cast<float4 *>(qp) = qo;
        *reinterpret_cast<float4 *>(qp + 512) = up;
        rs_bar_arrive(rs_bar_full(ROLE, qs));
    }
    if (ROLE == 0) {
        const int LL = k + 1;      // odd tick parity -> element 1 of the pair of (k0, k0 + 1), even -> element 0 of the pair of (k0 + 2, k0 + 3)
        if (LL >= 0 && LL < K.nproc)      // the loop's last (odd) tick may lie one past the last line
            rs_load_line(K, Q, LL, Q.ring + (A ? po[0] : po[1] + 1024u));
    }
    if (ROLE == RS_NW - 1) {
        const int w = k - 15;
        // line w: columns A = second (up), columns 1-A = F1; line w-1: columns A = F2 before the step (qo), 1-A = last tick's second
        const float4 qE = A ? f1last : up, qO = A ? up : f1last;
        const float4 xE = A ? wsn : qo, xO = A ? qo : wsn;
        const float ql = __shfl_up_sync(0xffffffffu, qO.w, 1);
        if (w >= K.own0 && w <= K.last_owned) rs_write_line(K, Q, P, w, qE, qO, xE, xO, ql);
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
        // pairs this warp touches: the loader's (q1 + 1) and those of lines k0 - 2t, t = 2 ROLE .. 2 ROLE + 2
        unsigned po[RS_HS + 3];
        {
            int x = q1 + 1; if (x >= RS_NP) x -= RS_NP;
            po[0] = (unsigned)x * RS_PAIRB;
        }
#pragma unroll
        for (int d = 0; d < RS_HS + 2; d++) {
            int x = q1 - ROLE * RS_HS - d;
            if (x < 0) x += RS_NP;
            po[1 + d] = (unsigned)x * RS_PAIRB;
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
__global__ void __launch_bounds__(RS_THREADS, RS_CPS) k_rbq_stream(const RBQ P)
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

    RSK K;
    K.ring_nd = rs_smem + 16 * lane; K.ring_c = rs_smem + 2048 + 16 * lane;
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

    // ring = zeros (lines before the first read as 1/s = 0, -D0 = 0), count -> 1/s table, staging barriers
    for (int o = 16 * (int)threadIdx.x; o < RS_NP * RS_PAIRB; o += 16 * RS_THREADS) *reinterpret_cast<float4 *>(rs_smem + o) = make_float4(0.f, 0.f, 0.f, 0.f);
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
    } else if (role == 1) {
        rs_run<1, STATS>(K, Q, P);
    } else if (role == 2) {
        rs_run<2, STATS>(K, Q, P);
    } else {
        if (lane == 0)
            for (int n = 0; n < RS_WST && n < Q.nown; n++) rs_stage_wline(Q, P.TJ, n);
        rs_run<3, STATS>(K, Q, P);
    }
}
