// rbq_stream.cuh -- the pressure solve of rbq_fused.cuh (red-black SOR in pressure form, fluid.go:188-234
// in red-black order, all <= 8 iterations of a pass in ONE trip over HBM) as a REGISTER pipeline.
//
// Same arithmetic, cell for cell, as k_rbq_fused and its CPU restatement fo_project_redblack_q
// (the test oracle under oracle/): q' = fma(wd/s, nb - D0, fma(-wd, q, q)); U, V, p materialised once at the end.
// What changed is where q lives between the half sweeps and who synchronises with whom:
//
//   k_rbq_fused   q circulates through a 28-line ring in shared memory; one pair of warps per iteration,
//                 hand-offs through per-line mbarriers: 6.5 warp instructions and ~720 shared-memory
//                 wavefronts per line of 512 cells; issue slots 66 %, shared-memory pipe 66 %, DRAM 30 %.
//   k_rbq_stream  ONE WARP (= one CTA) carries a strip of 128 columns through ALL eight iterations and never talks to
//                 another warp.  A lane owns 4 consecutive columns (even | odd = two float2) and keeps, per
//                 iteration, the four vectors the two half sweeps of a step need (P2, P1, F2, F1 below) plus the two
//                 vectors the previous iteration handed over: 8 x 12 registers.  Iteration t runs THREE lines behind
//                 iteration t-1 and consumes what that iteration produced in the PREVIOUS tick, so the eight
//                 iterations of a tick are independent instruction streams inside one thread: the fixed-latency
//                 dependencies that throttle a warp (one instruction every 4-9 cycles in the earlier forms of this
//                 kernel) are covered by instruction-level parallelism instead of by warps.  Left / right neighbours
//                 beyond a lane's columns come from the next lane by shuffle.  Shared memory holds only what every
//                 iteration re-reads, -D0 and 1/s (8 B per cell, a 26-line ring; a lane reads what it wrote itself, so
//                 not even __syncwarp is needed), and two small TMA staging rings (U, V, mask in; U0, V0, mask again
//                 for the write-out).  A strip overlaps its neighbours by the 16-column dependency cone of 16 half
//                 sweeps on either side, a chunk by 16 lines.
//   Round-2 history (4098^2, ncu; time of the solve, k_rbq_fused 0.182 ms):
//     two warps x four iterations, 8 columns per lane, 199 registers, 6 warps per SM                      0.248 ms
//       5.3 warp instructions per cell, one instruction per warp every 4.07 cycles, 1.5 warps per scheduler
//     four warps x two iterations, named-barrier queues between them, 12 warps per SM                     0.257 ms
//       39 % of the stall samples at the queue barriers (the pipeline runs at the pace of its slowest role and
//       the warps of one role share a scheduler), 13 % instruction-cache misses (four code paths)
//     one warp x eight iterations, lag 3 (this form)
//
// Step r of iteration t (A = parity of the active columns = parity of line r):
//   first  = colour-0 half sweep on line r     : own qo = old[r][A], up = old[r+1][A], dn = P2 = old[r-1][A],
//                                                left / right from P1 = old[r][1-A]
//   second = colour-1 half sweep on line r-1   : own P2, up = first(r), dn = F2 = first(r-2), left / right from
//                                                F1 = first(r-1)
//   then line r-1 is final for this iteration: columns A = second, columns 1-A = F1.
// "old" is the previous iteration's output.  Iteration t is at step k - 3t in tick k; what it needs from iteration
// t-1 was produced by THAT iteration's step of tick k-1: up = its second half sweep (X), qo = the F2 it started that
// step with (Y).  Names rotate instead of values moving (P2 <- up, F2 <- first; the roles of the a / b registers swap
// every tick; the loop body is two ticks); the iterations of a tick run in descending order so that one X / Y pair
// per iteration suffices.
#pragma once
#include "rbq_fused.cuh"

#define RS_W 128                     // columns per warp
#define RS_H 16                      // halo = dependency cone of 16 half sweeps
#define RS_TJ_MAX (RS_W - 2 * RS_H)  // 96 owned columns, 24 lanes
#define RS_LAG 3                     // lines between consecutive iterations
#define RS_NP 13                     // ring of line PAIRS: lines k + 1 (loader) .. k - 3 * 7 - 1 (iteration 7) are alive in tick k
#define RS_PAIRB 2048                // bytes per pair: -D0 [line parity][column parity][64] floats, then 1/s alike
#ifndef RS_LST
#define RS_LST 4                     // loader staging ring (lines): two in use, two in flight; power of two
#endif
#ifndef RS_WST
#define RS_WST 4                     // writer staging ring; power of two
#endif
#define RS_LSTB 1184                 // U 512 | V 544 (132 floats: one column beyond, padded) | mask 128
#define RS_WSTB (9 * RS_TJ_MAX)      // U0 | V0 | mask of the owned columns
#ifndef RS_CPS
#define RS_CPS 6                     // warps (= CTAs) per SM the kernel is compiled for
#endif
#define RS_THREADS 32
#define RS_WLAG (RS_LAG * (RQ_NIT - 1) + 2)      // the writer takes line k - 23 in tick k
static_assert((RS_LST & (RS_LST - 1)) == 0 && (RS_WST & (RS_WST - 1)) == 0 && RS_LST >= 2, "ring depths");
static_assert(2 * RS_NP >= RS_LAG * (RQ_NIT - 1) + 4, "a line's ring slot must outlive its last reader");

#define RS_OFF_LSTG (RS_NP * RS_PAIRB)
#define RS_OFF_WSTG (RS_OFF_LSTG + RS_LST * RS_LSTB)
#define RS_OFF_BARS (RS_OFF_WSTG + RS_WST * RS_WSTB)
#define RS_OFF_RSLUT (RS_OFF_BARS + 8 * (RS_LST + RS_WST))
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

struct RSK {                         // per-warp constants of the tick
    const unsigned char *ring_nd, *ring_c;   // ring + 8 * lane, ring + 1024 + 8 * lane
    int lane, TJ;
    int own0, last_owned, nproc;
};

__device__ __forceinline__ float2 rs_lds64(const unsigned char *a) { return *reinterpret_cast<const float2 *>(a); }

// One cell update, two same-colour cells at a time: rq_update's operations on the same values.  The neighbour that is
// shifted by one cell against the register pair (left for even columns, right for odd ones) is added with scalar
// FADDs.  rs = 1/s of the two cells (0: never updated); wd * rs is the product the wd / s table of k_rbq_fused holds.
template <int A>
__device__ __forceinline__ float2 rs_update(const float2 qo, const float2 up, const float2 dn, const float2 ot, const float ox, const float2 nd,
                                            const float2 rs, const float wd, const float nwd, float2 &t_out)
{
    // nb = ((q[i-1,j] + q[i+1,j]) + q[i,j-1]) + q[i,j+1];  t = nb - D0
    float2 s = __fadd2_rn(dn, up);
    if (A == 0) {          // even columns c0, c2: left = (ox, o0), right = (o0, o1)
        s.x = s.x + ox; s.y = s.y + ot.x;
        s = __fadd2_rn(s, ot);
    } else {               // odd columns c1, c3: left = (o0, o1), right = (o1, ox)
        s = __fadd2_rn(s, ot);
        s.x = s.x + ot.y; s.y = s.y + ox;
    }
    const float2 t = __fadd2_rn(s, nd);
    // q' = fma(wd*rs, t, fma(-wd, q, q))
    const float2 b = __ffma2_rn(make_float2(nwd, nwd), qo, qo);
    const float2 n = __ffma2_rn(__fmul2_rn(make_float2(wd, wd), rs), t, b);
    t_out = t;
    return n;
}

// per-iteration max |div| before the update (statistics only): cells of the owned rows / columns with s != 0
__device__ __forceinline__ void rs_stat(const float2 qo, const float2 t, const float2 rs, const int lj0, const bool row_owned, const int TJ, float &mymax)
{
    const float qv[2] = { qo.x, qo.y }, tv[2] = { t.x, t.y }, rv[2] = { rs.x, rs.y };
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int lj = lj0 + 2 * k;
        if (rv[k] != 0.0f && row_owned && lj >= RS_H && lj < RS_H + TJ) {
            const float ns = (float)__float2int_rn(1.0f / rv[k]);
            const float ad = fabsf(__fmaf_rn(ns, qv[k], -tv[k]));
            if (ad > mymax) mymax = ad;
        }
    }
}

// One step of one iteration (see the header).  qo / up come from the previous iteration (zeros for iteration 0);
// returns the second half sweep's result; F2 on return is first(r), `f2old` the F2 the step started with.
template <int A, bool STATS>
__device__ __forceinline__ float2 rs_step(const RSK &K, const RBQ &P, const int t, const float2 qo, const float2 up, float2 &P2, const float2 P1,
                                          float2 &F2, const float2 F1, const unsigned o1, const unsigned o2, float2 &f2old, const int r, float &mymax)
{
    const float2 nd = rs_lds64(K.ring_nd + o1), ndb = rs_lds64(K.ring_nd + o2);
    const float2 rs1 = rs_lds64(K.ring_c + o1), rs2 = rs_lds64(K.ring_c + o2);
    const float ox = A ? __shfl_down_sync(0xffffffffu, P1.x, 1) : __shfl_up_sync(0xffffffffu, P1.y, 1);
    const float ox2 = A ? __shfl_down_sync(0xffffffffu, F1.x, 1) : __shfl_up_sync(0xffffffffu, F1.y, 1);
    float2 tt;
    const float2 fnow = rs_update<A>(qo, up, P2, P1, ox, nd, rs1, P.wd[2 * t], P.nwd[2 * t], tt);
    if (STATS) rs_stat(qo, tt, rs1, 4 * K.lane + A, r >= K.own0 && r <= K.last_owned, K.TJ, mymax);
    const float2 snow = rs_update<A>(P2, fnow, F2, F1, ox2, ndb, rs2, P.wd[2 * t + 1], P.nwd[2 * t + 1], tt);
    if (STATS) rs_stat(P2, tt, rs2, 4 * K.lane + A, r - 1 >= K.own0 && r - 1 <= K.last_owned, K.TJ, mymax);
    f2old = F2;
    P2 = up;
    F2 = fnow;
    return snow;
}

struct RSIO {                        // loader / writer constants of a warp
    unsigned char *ring, *lstg, *wstg;
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
        rq_tma_load(sk + 512u + 4u * (unsigned)Q.off, Q.V + o, Q.bV, fb);
        rq_tma_load(sk + 1056u + (unsigned)Q.off, Q.mask + o, Q.bM, fb);
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
    const float4 ua = *reinterpret_cast<const float4 *>(s0 + 16 * lane);
    const float4 na = *reinterpret_cast<const float4 *>(s1 + 16 * lane);
    const float4 va = *reinterpret_cast<const float4 *>(s0 + 512 + 16 * lane);
    const unsigned m4 = *reinterpret_cast<const unsigned *>(s0 + 1056 + 4 * lane);
    float vn = __shfl_down_sync(0xffffffffu, va.x, 1);
    if (lane == 31) vn = *reinterpret_cast<const float *>(s0 + 512 + 512);        // column jw0 + 128
    const int j0 = Q.jw0 + 4 * lane;
    const bool on = LL >= Q.live_lo && LL <= Q.live_hi && j0 >= 0 && j0 < Q.PIT;
    const unsigned cnt = on ? (m4 >> MK_CNT_SHIFT) & 0x07070707u : 0u;
    // -div = -(((U[i+1,j] - U[i,j]) + V[i,j+1]) - V[i,j]), the reference's operations (fluid.go:207) and a sign flip
    const float2 x01 = __fadd2_rn(make_float2(na.x, na.y), make_float2(-ua.x, -ua.y)), x23 = __fadd2_rn(make_float2(na.z, na.w), make_float2(-ua.z, -ua.w));
    const float dv0 = (x01.x + va.y) - va.x, dv1 = (x01.y + va.z) - va.y, dv2 = (x23.x + va.w) - va.z, dv3 = (x23.y + vn) - va.w;
    float2 dE, dO, rE, rO;
    if (cnt == 0x04040404u) {                                // four interior fluid cells: nothing to select
        dE = make_float2(-dv0, -dv2); dO = make_float2(-dv1, -dv3);
        rE = rO = make_float2(0.25f, 0.25f);
    } else {
        dE.x = (cnt & 0x000000ffu) ? -dv0 : 0.0f; dO.x = (cnt & 0x0000ff00u) ? -dv1 : 0.0f;
        dE.y = (cnt & 0x00ff0000u) ? -dv2 : 0.0f; dO.y = (cnt & 0xff000000u) ? -dv3 : 0.0f;
        const float *lut = reinterpret_cast<const float *>(Q.ring + RS_OFF_RSLUT);     // count -> 1/s
        rE.x = lut[cnt & 7u]; rO.x = lut[(cnt >> 8) & 7u]; rE.y = lut[(cnt >> 16) & 7u]; rO.y = lut[cnt >> 24];
    }
    *reinterpret_cast<float2 *>(pr + 8 * lane) = dE;
    *reinterpret_cast<float2 *>(pr + 256 + 8 * lane) = dO;
    *reinterpret_cast<float2 *>(pr + 1024 + 8 * lane) = rE;
    *reinterpret_cast<float2 *>(pr + 1024 + 256 + 8 * lane) = rO;
    __syncwarp();                                    // every lane is done with staging slot st0
    if (lane == 0 && LL + RS_LST <= K.nproc) rs_stage_line(Q, LL + RS_LST);
}

// ---------------- writer: final q of line w (qE | qO) and of line w - 1 (xE | xO) -> U, V, p ----------------
__device__ __forceinline__ void rs_write_line(const RSK &K, const RSIO &Q, const RBQ &P, const int w, const float2 qE, const float2 qO,
                                              const float2 xE, const float2 xO, const float ql)
{
    const int lane = K.lane;
    const int n = w - K.own0, r = Q.i0c + n;
    const int ws = n & (RS_WST - 1);
    rs_wait(Q.b_wfull + 8u * (unsigned)ws, (unsigned)(n / RS_WST) & 1u, Q.debug, (32 << 20) | w);
    const int lo = lane - RS_H / 4, j0 = Q.jw0 + 4 * lane;
    if (lo >= 0 && lo < (K.TJ >> 2) && j0 < Q.NY) {
        const unsigned char *sb = Q.wstg + ws * RS_WSTB;
        const float4 u0 = *reinterpret_cast<const float4 *>(sb + 16 * lo);
        const float4 v0 = *reinterpret_cast<const float4 *>(sb + 4 * K.TJ + 16 * lo);
        const unsigned m4 = *reinterpret_cast<const unsigned *>(sb + 8 * K.TJ + 4 * lo);
        const size_t o = (size_t)(r - Q.i_alloc0) * Q.PIT + j0;
        float pin[4] = {0.f, 0.f, 0.f, 0.f};
        if (P.Pin) unpack(ld4(P.Pin + o), pin);
        const float qc[4] = { qE.x, qO.x, qE.y, qO.y };
        const float qx[4] = { xE.x, xO.x, xE.y, xO.y };
        const float uu[4] = { u0.x, u0.y, u0.z, u0.w };
        const float vv[4] = { v0.x, v0.y, v0.z, v0.w };
        float pu[4], pv[4], pp[4];
        const unsigned all = (MK_C | MK_XM | MK_YM) * 0x01010101u;
        if ((m4 & all) == all && r != 0 && j0 > 0) {
            // four fluid cells with fluid on their -x and -y side: every select of the general form is taken
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const float qym = (c == 0) ? ql : qc[c > 0 ? c - 1 : 0];
                const float t1 = uu[c] - qc[c];
                pu[c] = t1 + qx[c];
                const float t2 = vv[c] - qc[c];
                pv[c] = t2 + qym;
            }
        } else {
            const bool line_first = (r == 0);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const unsigned m = m4 >> (8 * c);
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
        for (int c = 0; c < 4; c++) pp[c] = __fmaf_rn(P.cp, qc[c], pin[c]);
        if (P.turb > 0.0f && r >= 1 && r <= Q.NX - 2) {          // fused addTurbulence (fluid.go:496-526)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const unsigned m = m4 >> (8 * c);
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
        if (j0 + 3 < Q.NY) {
            *reinterpret_cast<float4 *>(P.Uo + o) = make_float4(pu[0], pu[1], pu[2], pu[3]);
            *reinterpret_cast<float4 *>(P.Vo + o) = make_float4(pv[0], pv[1], pv[2], pv[3]);
            *reinterpret_cast<float4 *>(P.Po + o) = make_float4(pp[0], pp[1], pp[2], pp[3]);
        } else {
            for (int c = 0; c < 4 && j0 + c < Q.NY; c++) { P.Uo[o + c] = pu[c]; P.Vo[o + c] = pv[c]; P.Po[o + c] = pp[c]; }
        }
    }
    __syncwarp();
    if (lane == 0 && n + RS_WST < Q.nown) rs_stage_wline(Q, K.TJ, n + RS_WST);
}

// Byte offset (within a pair) of line element `e`, column parity `a`
#define RS_EO(e, a) ((unsigned)(e) * 512u + (unsigned)(a) * 256u)

// Tick k, KP = k & 1.  Iteration t runs step k - 3t (active column parity KP ^ (t & 1)), the writer takes line k - 23 from
// what iteration 7 produced in tick k - 1, the loader turns the staged line k + 1 into -D0 and 1/s.  po[d] = byte offset
// of the ring pair that holds lines (k0 + 2 - 2d, k0 + 3 - 2d), k0 the even tick of the body.  Every tick runs every
// iteration: steps r < 0 read the zeroed ring (1/s = 0, -D0 = 0) and leave q = 0; no step reaches line nproc.
template <int KP, bool STATS>
__device__ __forceinline__ void rs_tick(const RSK &K, const RSIO &Q, const RBQ &P, const int k, const unsigned (&po)[RS_NP], float2 (&P2)[RQ_NIT],
                                        float2 (&P1)[RQ_NIT], float2 (&F2)[RQ_NIT], float2 (&F1)[RQ_NIT], float2 (&X)[RQ_NIT + 1],
                                        float2 (&Y)[RQ_NIT + 1], float2 &wsn, float (&mymax)[RQ_NIT])
{
    // ---------------- writer: line w = k - 23, final since iteration 7's step of the PREVIOUS tick ----------------
    {
        const int w = k - RS_WLAG;
        // iteration 7 worked on column parity KP in tick k - 1.  Line w: columns KP = its second half sweep (X), columns
        // 1 - KP = its first(w), which is its F2 now; line w - 1: columns KP = the F2 it started with (Y), 1 - KP = the
        // second half sweep of the tick before (wsn)
        const float2 f2last = F2[RQ_NIT - 1];
        const float2 qE = KP ? f2last : X[RQ_NIT], qO = KP ? X[RQ_NIT] : f2last;
        const float2 xE = KP ? wsn : Y[RQ_NIT], xO = KP ? Y[RQ_NIT] : wsn;
        const float ql = __shfl_up_sync(0xffffffffu, qO.y, 1);
        if (w >= K.own0 && w <= K.last_owned) rs_write_line(K, Q, P, w, qE, qO, xE, xO, ql);
        wsn = X[RQ_NIT];
    }
    // ---------------- eight iterations, last first: each reads what its predecessor left in X / Y one tick ago ----------------
#pragma unroll
    for (int t = RQ_NIT - 1; t >= 0; t--) {
        const int r = k - RS_LAG * t;
        // pair index distance d of a line L from the pair of (k0 + 2, k0 + 3): d = (k0 + 2 - (L & ~1)) / 2
        // line r = k0 + KP - 3t, line r - 1
        constexpr int dummy = 0; (void)dummy;
        const int A = KP ^ (t & 1);
        const int off1 = KP - 3 * t;                 // r - k0
        const int e1 = off1 & 1, d1 = (2 - (off1 - e1)) / 2;
        const int off2 = off1 - 1;
        const int e2 = off2 & 1, d2 = (2 - (off2 - e2)) / 2;
        const unsigned o1 = po[d1] + RS_EO(e1, A), o2 = po[d2] + RS_EO(e2, A);
        float2 f2old, snow;
        const float2 qo = t == 0 ? make_float2(0.f, 0.f) : Y[t], up = t == 0 ? make_float2(0.f, 0.f) : X[t];
        if (A) snow = rs_step<1, STATS>(K, P, t, qo, up, P2[t], P1[t], F2[t], F1[t], o1, o2, f2old, r, mymax[t]);
        else snow = rs_step<0, STATS>(K, P, t, qo, up, P2[t], P1[t], F2[t], F1[t], o1, o2, f2old, r, mymax[t]);
        X[t + 1] = snow;
        Y[t + 1] = f2old;
    }
    // ---------------- loader: staged line LL = k + 1 ----------------
    {
        const int LL = k + 1;      // odd tick: element 1 of the pair of (k0, k0 + 1) (d = 1); even tick: element 0 of the pair of (k0 + 2, k0 + 3) (d = 0)
        if (LL >= 0 && LL < K.nproc)      // the loop's last (odd) tick may lie one past the last line
            rs_load_line(K, Q, LL, Q.ring + (KP ? po[0] : po[1] + 512u));
    }
}

template <bool STATS>
__global__ void __launch_bounds__(RS_THREADS, RS_CPS) k_rbq_stream(const RBQ P)
{
    extern __shared__ __align__(128) unsigned char rs_smem[];
    const int lane = threadIdx.x;
    const Grid g = P.g;
    const int NX = g.NX, PIT = g.pitch;
    const int strip = blockIdx.x;
    const int i0c = P.ib + blockIdx.y * P.chunk;
    const int i1c = min(i0c + P.chunk, P.ie);
    if (i0c >= i1c) return;
    const int jw0 = strip * P.TJ - RS_H;
    const int e0 = (i0c - RS_H) & ~1, e1 = i1c + RS_H;        // even first line: tick parity == line parity of iteration 0

    RSK K;
    K.ring_nd = rs_smem + 8 * lane; K.ring_c = rs_smem + 1024 + 8 * lane;
    K.lane = lane; K.TJ = P.TJ;
    K.own0 = i0c - e0; K.last_owned = i1c - 1 - e0; K.nproc = e1 - e0;

    RSIO Q;
    Q.ring = rs_smem; Q.lstg = rs_smem + RS_OFF_LSTG; Q.wstg = rs_smem + RS_OFF_WSTG;
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
    for (int o = 16 * lane; o < RS_NP * RS_PAIRB; o += 16 * RS_THREADS) *reinterpret_cast<float4 *>(rs_smem + o) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < 8) reinterpret_cast<float *>(rs_smem + RS_OFF_RSLUT)[lane] = lane == 1 ? 1.0f : (lane == 2 ? 0.5f : (lane == 3 ? (1.0f / 3.0f) : (lane == 4 ? 0.25f : 0.0f)));
    if (lane < RS_LST + RS_WST) rq_mbar_init(reinterpret_cast<unsigned long long *>(rs_smem + RS_OFF_BARS) + lane, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        for (int line = 0; line < RS_LST && line <= K.nproc; line++) rs_stage_line(Q, line);
        for (int n = 0; n < RS_WST && n < Q.nown; n++) rs_stage_wline(Q, P.TJ, n);
    }

    float2 pa[RQ_NIT], pb[RQ_NIT], fa[RQ_NIT], fb[RQ_NIT], X[RQ_NIT + 1], Y[RQ_NIT + 1];
    float mymax[RQ_NIT];
#pragma unroll
    for (int t = 0; t < RQ_NIT; t++) { pa[t] = pb[t] = fa[t] = fb[t] = make_float2(0.f, 0.f); mymax[t] = 0.0f; }
#pragma unroll
    for (int t = 0; t <= RQ_NIT; t++) X[t] = Y[t] = make_float2(0.f, 0.f);
    float2 wsn = make_float2(0.f, 0.f);

    const int kend = K.last_owned + RS_WLAG;                  // the tick that writes the last owned line
    int q1 = RS_NP - 1;                                       // ring pair of lines (k0, k0 + 1) at k0 = -2
#pragma unroll 1
    for (int k0 = -2; k0 <= kend; k0 += 2) {
        unsigned po[RS_NP];                                   // po[d]: pair of lines (k0 + 2 - 2d, k0 + 3 - 2d)
#pragma unroll
        for (int d = 0; d < RS_NP; d++) {
            int x = q1 + 1 - d;
            if (x < 0) x += RS_NP;
            if (x >= RS_NP) x -= RS_NP;
            po[d] = (unsigned)x * RS_PAIRB;
        }
        rs_tick<0, STATS>(K, Q, P, k0, po, pa, pb, fa, fb, X, Y, wsn, mymax);
        rs_tick<1, STATS>(K, Q, P, k0 + 1, po, pb, pa, fb, fa, X, Y, wsn, mymax);
        if (++q1 == RS_NP) q1 = 0;
    }
    if (STATS) {
        const int nit = P.nstages >> 1;
#pragma unroll
        for (int t = 0; t < RQ_NIT; t++) {
            const float m = warp_max(mymax[t]);
            if (lane == 0 && t < nit && m > 0.0f) atomicMax(P.stats + ((P.stage0 >> 1) + t), __float_as_uint(m));
        }
    }
}
