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
// (constant i, contiguous in j) stream through a ring of 35 slots in 223 KB of shared
// memory.  26 warps form a software pipeline WITHOUT block-wide barriers:
//   thread 768   producer 1: TMA bulk copies (cp.async.bulk + mbarrier complete_tx) of the U, V,
//                mask segments of a line into a 4-deep staging ring, 4 lines ahead of the loader
//   warps 16-19  loader: staging -> -D0, neighbour count (bits 5-7 of the mask byte), q = 0 in slot(L)
//   warp  s<16   half sweep s (colour s&1): may process line r once its predecessor
//                (loader for s=0, warp s-1 otherwise) has finished line r+1
//   thread 800   producer 2: TMA bulk copies of U0, V0, mask of the OWNED columns of the owned
//                lines into an 8-deep ring for the writer
//   warps 20-23  writer: slot(r), slot(r-1) + U0, V0, mask from its ring -> U, V, p in global
// Hand-offs are per-line mbarriers (64 per role; every lane of the role arrives, every lane of
// the successor polls with try_wait); the loader reuses a slot once the writer (or, for halo
// lines, the last half sweep) is past it.  Waits are bounded: a pipeline that stops latches a
// debug record and the host returns FB_ERR_CUDA instead of hanging the GPU.
// Even and odd columns live in separate arrays so one colour is contiguous: a lane
// updates 2 x 4 consecutive same-colour cells with LDS.128 / STS.128 and packed
// FADD2 / FFMA2 (sm_100a fp32x2, bit-identical to the scalar operations).
//
// Why this shape (ncu, 4098^2, 8 iterations; time of the solve):
//   face form (rb_fused.cuh), ~110 instructions per cell update, issue-bound          1.11 ms
//   pressure form, one __syncthreads per line (barrier = 3.4 of 9 stall cycles)       0.47
//   warp-specialised roles with mbarrier hand-offs, strips x chunks = one wave        0.31
//   writer inputs through TMA: its re-read of U0, V0 missed L2 two times in three and
//   the register prefetch ring did not survive code generation (74 % of the writer's
//   stall samples sat on the first use of those loads)                                0.226
//   neighbour counts baked into the mask, lean loader / writer loops                  0.211
// Now issue slots are 73 % and shared-memory wavefronts 74 % busy, DRAM 25 %.  Measured and NOT
// adopted: one arrive per warp instead of 32 (no change); a 2-instruction poll loop (0.224:
// try_wait is a shared-memory operation, faster polling takes wavefronts from the sweeps);
// nanosleep after a failed poll (no change); 37 / 40 ring slots, 8-deep loader staging (no
// change); re-reading the neighbour vectors instead of carrying them (0.231: LSU pipe 79 %);
// turbulence fused into the writer (solve + turbulence 3.52 -> 4.20 ms at 16386^2); four
// lines in flight per loader / writer warp (0.47).
#pragma once
#include "kernels.cuh"
#include "advect_fused.cuh"

#ifndef RQ_PAIR
#define RQ_PAIR 1         // 1: one warp per ITERATION (red + black half sweep fused), 0: one warp per half sweep
#endif
#ifndef RQ_NL
#define RQ_NL (RQ_PAIR ? 28 : 35)   // line slots
#endif
#define RQ_H 16
#define RQ_SW (RQ_PAIR ? 8 : 16)    // sweep warps
#define RQ_LD0 (32 * RQ_SW)         // first loader thread (4 warps)
#define RQ_WR0 (RQ_LD0 + 128)       // first writer thread (4 warps)
#define RQ_P1 (RQ_WR0 + 128)        // producer 1 (TMA for the loader)
#define RQ_P2 (RQ_P1 + 32)          // producer 2 (TMA for the writer)
#define RQ_THREADS (RQ_P2 + 32)
#define RQ_WROLE (RQ_SW + 1)        // hand-off role of the writer
#define RQ_TJ_MAX 464     // multiple of 16; WL = TJ + 48 <= 512 (a lane owns 2 groups of 4 cells per line;
                          // 16-byte granules for the TMA copies of the mask)
#ifndef RQ_STG
#define RQ_STG (RQ_PAIR ? 8 : 4)    // staging ring depth (lines in flight through TMA)
#endif
// shared memory at WL = 512, TJ = 464: 35 slots * WL * 9 B (q, -D0, neighbour count) = 157.5 KB, loader
// staging RQ_STG * (WL*9 + 16) B = 18.1 KB, writer staging RQ_WSTG * TJ * 9 B = 32.6 KB, hand-off
// mbarriers 9.1 KB, wd/s table 0.5 KB: 217.8 KB of the 227 KB a CTA may have
__host__ __device__ __forceinline__ size_t rq_stage_bytes(int WL) { return (size_t)WL * 9 + 16; }
#ifndef RQ_WSTG
#define RQ_WSTG 8         // writer staging ring depth
#endif
__host__ __device__ __forceinline__ size_t rq_wstage_bytes(int TJ) { return (size_t)TJ * 9; }   // U0, V0, mask of TJ columns
__host__ __device__ __forceinline__ size_t rq_smem_bytes(int WL, int TJ)
{
    return (size_t)RQ_NL * WL * 9 + RQ_STG * rq_stage_bytes(WL) + RQ_WSTG * rq_wstage_bytes(TJ) + 8 * (RQ_STG + RQ_WSTG) +
           8 * (RQ_SW + 2) * 64 + 16 * 8 * 4 + 64;
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

__device__ __forceinline__ unsigned rq_s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rq_mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rq_s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rq_mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rq_s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rq_arrive_a(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void rq_mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(rq_s32(bar)) : "memory");
}
// no ordering of the thread's other memory operations: for roles that only READ the slots they hand back
__device__ __forceinline__ void rq_mbar_arrive_relaxed(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(rq_s32(bar)) : "memory");
}
#ifndef RQ_WAIT_HINT
#define RQ_WAIT_HINT 0      // ns the hardware may suspend a failed try_wait (0: its default)
#endif
__device__ __forceinline__ bool rq_mbar_try_a(unsigned bar, unsigned parity)
{
    unsigned ok;
#if RQ_WAIT_HINT
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity), "r"(RQ_WAIT_HINT) : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
    return ok != 0;
}
// Bounded wait: a pipeline that stops making progress must never hang the GPU.  After
// ~2^21 failed polls (hundreds of milliseconds) the first waiter records who waited for
// what in P.debug and every waiter falls through; the host turns the flag into an error.
// The polling loop is kept to try_wait + branch (4 polls per trip): the kernel is
// issue-bound and spinning warps share the schedulers with the warps they wait for.
__device__ int *rq_debug;   // set per launch (RBQ::debug)
__device__ __noinline__ void rq_wait_slow(unsigned bar, unsigned parity, int tag)
{
    for (unsigned trip = 0;; trip++) {
#pragma unroll 1
        for (int k = 0; k < 256; k++) {
            if (rq_mbar_try_a(bar, parity)) return;
            if (rq_mbar_try_a(bar, parity)) return;
            if (rq_mbar_try_a(bar, parity)) return;
            if (rq_mbar_try_a(bar, parity)) return;
        }
        int *d = rq_debug;
        if (d && *reinterpret_cast<volatile int *>(d) != 0) return;         // someone already gave up: drain
        if (trip > (1u << 11)) {
            if (d && atomicCAS(d, 0, 1) == 0) {
                d[1] = tag; d[2] = (int)threadIdx.x; d[3] = (int)blockIdx.x; d[4] = (int)blockIdx.y; d[5] = (int)parity;
                __threadfence();
            }
            return;
        }
    }
}
// Polling loop.  A waiting warp shares its scheduler AND the shared-memory pipe (try_wait is a
// shared-memory operation) with the warps it waits for; measured at 4098^2, 8 iterations:
// 7-instruction poll 211 us, 2-instruction poll 224 us (more polls per microsecond, not fewer).
// RQ_POLL_SLEEP > 0 inserts a nanosleep after every failed poll.
#ifndef RQ_POLL_SLEEP
#define RQ_POLL_SLEEP 0
#endif
#ifndef RQ_SLEEP_IO
#define RQ_SLEEP_IO 0      // ns the loader / writer / producer roles sleep after a failed poll
#endif
template <int SLEEP = RQ_POLL_SLEEP>
__device__ __forceinline__ void rq_mbar_wait_a(unsigned bar, unsigned parity, int tag)
{
#pragma unroll 1
    for (int k = 0; k < 4096; k++) {
        if (rq_mbar_try_a(bar, parity)) return;
        if (SLEEP > 0) __nanosleep(SLEEP);
    }
    rq_wait_slow(bar, parity, tag);
}
// the roles around the sweeps (loader, writer, the two TMA producers)
__device__ __forceinline__ void rq_mbar_wait_io(unsigned bar, unsigned parity, int tag) { rq_mbar_wait_a<RQ_SLEEP_IO>(bar, parity, tag); }
__device__ __forceinline__ void rq_mbar_wait(unsigned long long *bar, unsigned parity, int tag = 0)
{
    rq_mbar_wait_a(rq_s32(bar), parity, tag);
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void rq_tma_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rq_s32(dst)), "l"(src), "r"(bytes), "r"(rq_s32(bar)) : "memory");
}

// One line of one half sweep: the active cells have column parity A.  e_own / e_up / e_dn are
// the element offsets of the three slots, NDO the distance from q to -D0 of the same cell.
// Carried across consecutive lines of one half sweep (registers): the `up` vector of line r is
// the other-parity vector of line r+1, and the other-parity vector of line r is the `down`
// vector of line r+1 (nobody writes them in between).  Shared-memory wavefronts are as scarce
// as issue slots here (ncu: 79 % of the LSU data pipe without the carry, 38 % with it).
struct RQCarry { float4 up[2], ot[2]; };
template <int A, bool STATS>
__device__ __forceinline__ void rq_line(float *__restrict__ sQ, const unsigned char *__restrict__ sC, int e_own, int e_up,
                                        int e_dn, int WQ, int NDO, int lane4, float wd, float c4, bool row_owned, int TJ,
                                        float &mymax, const float *__restrict__ tw, RQCarry &cy, bool have)
{
    const float2 nwd2 = make_float2(-wd, -wd);
    const float2 c44 = make_float2(c4, c4);               // wd * (1/s) for a cell with four fluid neighbours
    const int own = e_own + A * WQ + lane4, oth = e_own + (1 - A) * WQ + lane4;
    const int upo = e_up + A * WQ + lane4, dno = e_dn + A * WQ + lane4;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int h0 = 128 * half;
        if (lane4 + h0 >= WQ) continue;
        float *qown = sQ + own + h0;
        const float4 qo = *reinterpret_cast<const float4 *>(qown);
        const float4 up = *reinterpret_cast<const float4 *>(sQ + upo + h0);
#ifdef RQ_NO_CARRY
        const float4 dn = *reinterpret_cast<const float4 *>(sQ + dno + h0);
        const float4 ot = *reinterpret_cast<const float4 *>(sQ + oth + h0);
#else
        const float4 dn = have ? cy.ot[half] : *reinterpret_cast<const float4 *>(sQ + dno + h0);
        const float4 ot = have ? cy.up[half] : *reinterpret_cast<const float4 *>(sQ + oth + h0);
        cy.up[half] = up;
        cy.ot[half] = ot;
#endif
        const float ox = sQ[oth + h0 + (A ? 4 : -1)];
        const float4 nd = *reinterpret_cast<const float4 *>(qown + NDO);           // -D0
        const unsigned code = *reinterpret_cast<const unsigned *>(sC + own + h0);  // fluid-neighbour counts, 0 = skip
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
        const float2 b01 = __ffma2_rn(nwd2, q01, q01), b23 = __ffma2_rn(nwd2, q23, q23);
        float2 n01, n23;
        if (code == 0x04040404u) {                       // the common case: four interior cells
            n01 = __ffma2_rn(c44, t01, b01);
            n23 = __ffma2_rn(c44, t23, b23);
        } else {                                         // walls, obstacles, domain edge: wd / s from this half sweep's table
            n01 = __ffma2_rn(make_float2(tw[code & 7u], tw[(code >> 8) & 7u]), t01, b01);
            n23 = __ffma2_rn(make_float2(tw[(code >> 16) & 7u], tw[(code >> 24) & 7u]), t23, b23);
        }
        *reinterpret_cast<float4 *>(qown) = make_float4(n01.x, n01.y, n23.x, n23.y);
        if (STATS) {
            const float qv[4] = { qo.x, qo.y, qo.z, qo.w }, tv[4] = { t01.x, t01.y, t23.x, t23.y };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned ns = (code >> (8 * k)) & 0xffu;
                const int lj = 2 * (lane4 + h0 + k) + A;
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
#define RQ_RING 64
#define RQ_ROLES (RQ_SW + 2)
#define RQ_LAST_ROLE(nst) (RQ_PAIR ? ((nst) >> 1) : (nst))   // role whose arrivals mean "all half sweeps are past this line"

// ---------------------------------------------------------------------------------------------
// RQ_PAIR: one warp per ITERATION.  Step `rel` of the warp is the red half sweep on line rel
// ("first") followed by the black half sweep on line rel-1 ("second"); both touch the columns of
// the same parity A = (colour + line) & 1.  Everything `second` needs is already in registers:
//   own  = q_old[rel-1][A]   the `dn` of first(rel), which was the `up` loaded two steps ago
//   up   = first(rel)         computed a few instructions earlier
//   left / right = first(rel-1), dn = first(rel-2): results of the two previous steps
// and `first` reads only its own vector and the line above from shared memory (its `dn` and
// left / right are the `up` vectors of the two previous steps).  Per four cell updates this is
// 2.5 LDS.128 instead of 5, half the hand-offs per line, and 8 pipeline stages instead of 16.
// The rotation of the carried vectors is done by swapping argument names in a loop unrolled by
// two, so it costs no MOVs.
struct RQPair { float4 pa[2], pb[2], fa[2], fb[2]; };

template <int A>
__device__ __forceinline__ float4 rq_update(const float4 qo, const float4 up, const float4 dn, const float4 ot, const float ox,
                                            const float4 nd, const unsigned code, const float2 nwd2, const float2 c44,
                                            const float *__restrict__ tw, float4 &t_out)
{
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
    const float2 b01 = __ffma2_rn(nwd2, q01, q01), b23 = __ffma2_rn(nwd2, q23, q23);
    float2 n01, n23;
    if (code == 0x04040404u) {                       // the common case: four interior cells
        n01 = __ffma2_rn(c44, t01, b01);
        n23 = __ffma2_rn(c44, t23, b23);
    } else {                                         // walls, obstacles, domain edge: wd / s from this half sweep's table
        n01 = __ffma2_rn(make_float2(tw[code & 7u], tw[(code >> 8) & 7u]), t01, b01);
        n23 = __ffma2_rn(make_float2(tw[(code >> 16) & 7u], tw[(code >> 24) & 7u]), t23, b23);
    }
    t_out = make_float4(t01.x, t01.y, t23.x, t23.y);
    return make_float4(n01.x, n01.y, n23.x, n23.y);
}

template <bool STATS>
__device__ __forceinline__ void rq_stat(const float4 qo, const float4 t, const unsigned code, int lj0, bool row_owned, int TJ, float &mymax)
{
    if (!STATS) return;
    const float qv[4] = { qo.x, qo.y, qo.z, qo.w }, tv[4] = { t.x, t.y, t.z, t.w };
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned ns = (code >> (8 * k)) & 0xffu;
        const int lj = lj0 + 2 * k;
        if (ns && row_owned && lj >= RQ_H && lj < RQ_H + TJ) {
            const float ad = fabsf(__fmaf_rn((float)ns, qv[k], -tv[k]));     // |div| before the update
            if (ad > mymax) mymax = ad;
        }
    }
}

struct RQStage {
    float *sQ; const unsigned char *sC;
    int WQ, NDO, lane4, TJ;
    float2 nwd1, c41, nwd2, c42;        // (-wd, -wd) and (wd/4, wd/4) of the two half sweeps
    const float *tw1, *tw2;
    // hand-offs: wait on the predecessor's barriers, arrive on ours; slots of lines rel-1, rel, rel+1
    unsigned wbase, abase;
    int lag, nproc, tag;
    int e_dn, e_own, e_up, sl_up, ROW;
    int i0r, i1r;                       // owned lines, relative to e0 (STATS only)
    bool skip;
    float mymax;
};

// One step.  P2 (in: q_old[rel-1][A], out: the `up` loaded now), P1 (q_old[rel][1-A]), F2 (in: first(rel-2),
// out: first(rel)), F1 (first(rel-1)).  FIRST is false only for the step past the last line (line e1 is never
// swept and reads as q = 0), SECOND only for step 0.  In pair mode a slot is always 512 columns wide, so every
// lane owns two groups of four same-parity cells and nothing in the step is predicated.
template <int A, bool FIRST, bool SECOND, bool STATS>
__device__ __forceinline__ void rq_pair_step(RQStage &S, const int rel, float4 (&P2)[2], float4 (&P1)[2], float4 (&F2)[2],
                                             const float4 (&F1)[2])
{
    {
        const int w = min(rel + S.lag, S.nproc);
        rq_mbar_wait_a(S.wbase + 8u * (unsigned)(w & (RQ_RING - 1)), (unsigned)(w >> 6) & 1u, S.tag | w);
    }
    __syncwarp();          // the other lanes' stores of the previous step (left / right neighbours of `second`)
    float *const sQ = S.sQ;
    const int own = S.e_own + A * S.WQ + S.lane4, oth = S.e_own + (1 - A) * S.WQ + S.lane4;
    const int upo = S.e_up + A * S.WQ + S.lane4;
    const int own2 = S.e_dn + A * S.WQ + S.lane4, oth2 = S.e_dn + (1 - A) * S.WQ + S.lane4 + (A ? 4 : -1);
    if (!SECOND) {         // step 0: q_old[0][other parity] is the one carried vector that was never an `up`
        P1[0] = *reinterpret_cast<const float4 *>(sQ + oth);
        P1[1] = *reinterpret_cast<const float4 *>(sQ + oth + 128);
    }
    if (!S.skip) {
        float4 qo[2], up[2], nd[2], nd2[2];
        float ox[2], ox2[2];
        unsigned code[2], code2[2];
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int h0 = 128 * half;
            if (FIRST) {
                qo[half] = *reinterpret_cast<const float4 *>(sQ + own + h0);
                up[half] = *reinterpret_cast<const float4 *>(sQ + upo + h0);
                nd[half] = *reinterpret_cast<const float4 *>(sQ + own + h0 + S.NDO);
                ox[half] = sQ[oth + h0 + (A ? 4 : -1)];
                code[half] = *reinterpret_cast<const unsigned *>(S.sC + own + h0);
            }
            if (SECOND) {
                nd2[half] = *reinterpret_cast<const float4 *>(sQ + own2 + h0 + S.NDO);
                ox2[half] = sQ[oth2 + h0];
                code2[half] = *reinterpret_cast<const unsigned *>(S.sC + own2 + h0);
            }
        }
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int h0 = 128 * half;
            float4 t, fnow = make_float4(0.f, 0.f, 0.f, 0.f);
            if (FIRST) {
                fnow = rq_update<A>(qo[half], up[half], P2[half], P1[half], ox[half], nd[half], code[half], S.nwd1, S.c41, S.tw1, t);
                *reinterpret_cast<float4 *>(sQ + own + h0) = fnow;
                rq_stat<STATS>(qo[half], t, code[half], 2 * (S.lane4 + h0) + A, rel >= S.i0r && rel < S.i1r, S.TJ, S.mymax);
            }
            if (SECOND) {
                const float4 snow = rq_update<A>(P2[half], fnow, F2[half], F1[half], ox2[half], nd2[half], code2[half], S.nwd2, S.c42, S.tw2, t);
                *reinterpret_cast<float4 *>(sQ + own2 + h0) = snow;
                rq_stat<STATS>(P2[half], t, code2[half], 2 * (S.lane4 + h0) + A, rel - 1 >= S.i0r && rel - 1 < S.i1r, S.TJ, S.mymax);
            }
            if (FIRST) P2[half] = up[half];
            F2[half] = fnow;
        }
    }
    rq_arrive_a(S.abase + 8u * (unsigned)(rel & (RQ_RING - 1)));
    S.e_dn = S.e_own; S.e_own = S.e_up;
    if (++S.sl_up == RQ_NL) { S.sl_up = 0; S.e_up = 0; } else S.e_up += S.ROW;
}

// All steps 0 .. nproc of one iteration.  A0 = column parity of the active cells at step 0; it alternates
// from step to step, and the carried vectors swap names instead of moving (loop unrolled by two).
template <int A0, bool STATS>
__device__ __forceinline__ void rq_pair_stage(RQStage &S)
{
    float4 pa[2], pb[2], fa[2], fb[2];
#pragma unroll
    for (int half = 0; half < 2; half++) pa[half] = pb[half] = fa[half] = fb[half] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nproc = S.nproc;                                       // >= 2 * RQ_H + 1
    rq_pair_step<A0, true, false, STATS>(S, 0, pa, pb, fa, fb);
    int rel = 1;
    for (; rel + 1 < nproc; rel += 2) {
        rq_pair_step<1 - A0, true, true, STATS>(S, rel, pb, pa, fb, fa);
        rq_pair_step<A0, true, true, STATS>(S, rel + 1, pa, pb, fa, fb);
    }
    if (rel < nproc) {
        rq_pair_step<1 - A0, true, true, STATS>(S, rel, pb, pa, fb, fa);
        rq_pair_step<A0, false, true, STATS>(S, rel + 1, pa, pb, fa, fb);
    } else {
        rq_pair_step<1 - A0, false, true, STATS>(S, rel, pb, pa, fb, fa);
    }
}

// Every lane arrives and every lane polls: measured faster than one arrive / one poller per
// warp (lane-0 polling adds a divergent branch + __syncwarp to every hand-off: 0.35 -> 0.59 ms).
__device__ __forceinline__ void rq_done(unsigned long long *bars, int role, int line, int lane)
{
    rq_mbar_arrive(bars + role * RQ_RING + (line & (RQ_RING - 1)));
}
__device__ __forceinline__ void rq_wait_line(unsigned long long *bars, int role, int line)
{
    rq_mbar_wait_io(rq_s32(bars + role * RQ_RING + (line & (RQ_RING - 1))), (unsigned)(line / RQ_RING) & 1u, (role << 20) | line);
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
    float *sQ = reinterpret_cast<float *>(smem_raw);        // [slot][parity][q]
    float *sND = sQ + RQ_NL * WL;                            // -D0
    unsigned char *sC = reinterpret_cast<unsigned char *>(sND + RQ_NL * WL);   // fluid-neighbour count, 0 = never updated
    // staging ring: per slot WL floats of U, WL+4 floats of V, WL mask bytes (raw global data)
    unsigned char *stg = sC + RQ_NL * WL;                    // 16-byte aligned: WL is a multiple of 16
    const int STG = (int)rq_stage_bytes(WL);
    unsigned char *wstg = stg + RQ_STG * STG;                // writer staging ring: U0 | V0 | mask of the TJ owned columns
    const int WSTGB = (int)rq_wstage_bytes(P.TJ);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(wstg + RQ_WSTG * WSTGB);   // RQ_STG mbarriers
    unsigned long long *wfull = full + RQ_STG;               // RQ_WSTG mbarriers
    unsigned long long *bars = wfull + RQ_WSTG;              // RQ_ROLES * RQ_RING hand-off mbarriers
    float *tblw = reinterpret_cast<float *>(bars + RQ_ROLES * RQ_RING);   // [half sweep][fluid neighbours] -> wd / s

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

    if (tid == 0) rq_debug = P.debug;
    for (int k = tid; k < RQ_ROLES * RQ_RING; k += RQ_THREADS) {
        const int role = k / RQ_RING;
        rq_mbar_init(bars + k, (role == 0 || role == RQ_WROLE) ? 128 : 32);  // arrivals per phase = threads of the role
    }
    if (tid < RQ_STG) rq_mbar_init(full + tid, 1);
    if (tid >= 32 && tid < 32 + RQ_WSTG) rq_mbar_init(wfull + tid - 32, 1);
    if (tid < 128) {
        const int ns = tid & 7;
        const float rs = ns == 0 ? 0.0f : (ns == 1 ? 1.0f : (ns == 2 ? 0.5f : (ns == 3 ? (1.0f / 3.0f) : 0.25f)));
        tblw[tid] = P.wd[tid >> 3] * rs;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

#if RQ_PAIR
    if (warp < RQ_SW) {
        // ================= iteration `warp`: half sweeps 2*warp (first) and 2*warp + 1 (second) =================
        const int t = warp;
        if (2 * t >= nst) return;
        const int colour = P.stage0 & 1;                     // nst and stage0 are even: a pass is whole iterations
        RQStage S;
        S.sQ = sQ; S.sC = sC; S.WQ = WQ; S.NDO = RQ_NL * WL; S.lane4 = 4 * lane; S.TJ = P.TJ;
        S.nwd1 = make_float2(-P.wd[2 * t], -P.wd[2 * t]);
        S.nwd2 = make_float2(-P.wd[2 * t + 1], -P.wd[2 * t + 1]);
        { const float c1 = P.wd[2 * t] * 0.25f, c2 = P.wd[2 * t + 1] * 0.25f; S.c41 = make_float2(c1, c1); S.c42 = make_float2(c2, c2); }
        S.tw1 = tblw + 8 * (2 * t); S.tw2 = tblw + 8 * (2 * t + 1);
        // wait on the predecessor (loader: line rel+1 is loaded; iteration t-1: its step rel+2 is done, i.e. its
        // second half sweep is past line rel+1), arrive on role 1+t for every step 0 .. nproc
        S.wbase = rq_s32(bars + t * RQ_RING); S.abase = rq_s32(bars + (1 + t) * RQ_RING);
        S.lag = t == 0 ? 1 : 2; S.nproc = nproc; S.tag = t << 20;
        S.e_dn = (RQ_NL - 1) * ROW; S.e_own = 0; S.e_up = ROW; S.sl_up = 1; S.ROW = ROW;
        S.i0r = i0c - e0; S.i1r = i1c - e0;
        S.skip = (P.xflags & 1) != 0;
        S.mymax = 0.0f;
        if ((colour + e0) & 1) rq_pair_stage<1, STATS>(S); else rq_pair_stage<0, STATS>(S);
        float mymax = S.mymax;
        if (STATS) {
            mymax = warp_max(mymax);
            if (lane == 0 && mymax > 0.0f) atomicMax(P.stats + ((P.stage0 >> 1) + t), __float_as_uint(mymax));
        }
    } else if (warp < RQ_SW + 4) {
#else
    if (warp < 16) {
        // ================= half sweep `warp` =================
        const int s = warp;
        if (s >= nst) return;
        const int colour = (P.stage0 + s) & 1;
        const float wd = P.wd[s];
        const float c4 = wd * 0.25f;
        const float *tw = tblw + 8 * s;
        const int NDO = RQ_NL * WL, lane4 = 4 * lane;
        float mymax = 0.0f;
        int a = (colour + e0) & 1;
        // slots of lines rel-1, rel, rel+1 (element offsets), rotated line by line
        int e_dn = (RQ_NL - 1) * ROW, e_own = 0, e_up = ROW, sl_up = 1;
        // hand-off barriers: wait on the predecessor's (role s) barrier of line rel+1, arrive on ours (role 1+s) of line rel
        const unsigned wbase = rq_s32(bars + s * RQ_RING), abase = rq_s32(bars + (1 + s) * RQ_RING);
        const int rlo = 1 - e0, rhi = NX - 2 - e0;   // lines with updatable cells: rel in [rlo, rhi]
        const bool skip = (P.xflags & 1) != 0;
        RQCarry cy;
        bool have = false;
        for (int rel = 0; rel < nproc; rel++) {
            const int w = rel + 1;
            rq_mbar_wait_a(wbase + 8u * (unsigned)(w & (RQ_RING - 1)), (unsigned)(w >> 6) & 1u, (s << 20) | w);
            if (rel >= rlo && rel <= rhi && !skip) {
                const int r = e0 + rel;
                const bool row_owned = (r >= i0c) && (r < i1c);
                if (a) rq_line<1, STATS>(sQ, sC, e_own, e_up, e_dn, WQ, NDO, lane4, wd, c4, row_owned, P.TJ, mymax, tw, cy, have);
                else   rq_line<0, STATS>(sQ, sC, e_own, e_up, e_dn, WQ, NDO, lane4, wd, c4, row_owned, P.TJ, mymax, tw, cy, have);
                have = true;
            } else {
                have = false;
            }
            rq_arrive_a(abase + 8u * (unsigned)(rel & (RQ_RING - 1)));
            e_dn = e_own; e_own = e_up;
            if (++sl_up == RQ_NL) { sl_up = 0; e_up = 0; } else e_up += ROW;
            a ^= 1;
        }
        rq_arrive_a(abase + 8u * (unsigned)(nproc & (RQ_RING - 1)));   // line e1 is never swept: lets the next half sweep finish its last line
        if (STATS) {
            mymax = warp_max(mymax);
            if (lane == 0 && mymax > 0.0f) atomicMax(P.stats + ((P.stage0 + s) >> 1), __float_as_uint(mymax));
        }
    } else if (warp < 20) {
#endif
        // ================= loader: staging -> slot, lines e0 .. e1 =================
        const int ld = tid - RQ_LD0;
        const bool active = ld < (WL >> 2);
        const int j = jr0 + 4 * ld;
        const bool col_in = active && j >= 0 && j < PIT;
        const bool v4_in = j + 4 < PIT;
        if (active) {   // the slot "below" line e0 (relative -1) must read as q = 0
            const int q = 2 * ld, b0 = (RQ_NL - 1) * ROW + q;
            *reinterpret_cast<float2 *>(sQ + b0) = make_float2(0.f, 0.f);
            *reinterpret_cast<float2 *>(sQ + b0 + WQ) = make_float2(0.f, 0.f);
        }
        const int last_owned = (i1c - 1) - e0;
        // only interior lines inside this rank's slab hold updatable cells (the mask's count bits are
        // zero on the ring and in solids); line e1 is loaded but never swept
        const int live_lo = max(1, g.i_alloc0) - e0;
        const int live_hi = min(min(NX - 2, g.i_alloc0 + g.lines_alloc - 2), e1 - 1) - e0;
        const unsigned b_self = rq_s32(bars), b_writer = rq_s32(bars + RQ_WROLE * RQ_RING), b_last = rq_s32(bars + RQ_LAST_ROLE(nst) * RQ_RING);
        const unsigned b_full = rq_s32(full);
        // element offset of this thread's cells in slot 0; staging offsets of lines rel and rel+1
        int e_row = 2 * ld, sl = 0;
        int st0 = 0, st1 = (RQ_STG > 1) ? 1 : 0;
        unsigned par1 = 0;                                   // phase parity of staging slot st1's current use
        rq_mbar_wait_io(b_full, 0u, (30 << 20));              // line 0 has landed
        for (int rel = 0; rel <= nproc; rel++) {
            // slot(rel) last held line y-1 with y = rel-NL+1; its last readers work on line y: the writer
            // if y is an owned line, otherwise the last half sweep
            const int y = rel - RQ_NL + 1;
            if (y >= 0) {
                const unsigned bb = (y >= RQ_H && y <= last_owned) ? b_writer : b_last;
                rq_mbar_wait_io(bb + 8u * (unsigned)(y & (RQ_RING - 1)), (unsigned)(y >> 6) & 1u, (RQ_WROLE << 20) | y);
            }
            float d[4] = {0.f, 0.f, 0.f, 0.f};
            unsigned code = 0;
            // line rel+1 must have landed too (the producer stages every line 0 .. nproc); line rel was
            // waited for one iteration ago
            if (rel < nproc) rq_mbar_wait_io(b_full + 8u * (unsigned)st1, par1, (31 << 20) | rel);
            if (rel >= live_lo && rel <= live_hi && col_in) {
                const unsigned char *s0 = stg + st0 * STG, *s1 = stg + st1 * STG;
                const float *stU = reinterpret_cast<const float *>(s0) + 4 * ld, *stV = stU + WL;
                const float4 u0 = *reinterpret_cast<const float4 *>(stU);
                const float4 u1 = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(s1) + 4 * ld);
                const float4 v = *reinterpret_cast<const float4 *>(stV);
                const float v4 = v4_in ? stV[4] : 0.0f;
                const unsigned mk = *reinterpret_cast<const unsigned *>(s0 + (size_t)(2 * WL + 4) * 4 + 4 * ld);
                code = (mk >> MK_CNT_SHIFT) & 0x07070707u;   // fluid neighbours of updatable cells, else 0
                const float dv0 = ((u1.x - u0.x) + v.y) - v.x, dv1 = ((u1.y - u0.y) + v.z) - v.y;
                const float dv2 = ((u1.z - u0.z) + v.w) - v.z, dv3 = ((u1.w - u0.w) + v4) - v.w;
                d[0] = (code & 0x000000ffu) ? -dv0 : 0.0f;
                d[1] = (code & 0x0000ff00u) ? -dv1 : 0.0f;
                d[2] = (code & 0x00ff0000u) ? -dv2 : 0.0f;
                d[3] = (code & 0xff000000u) ? -dv3 : 0.0f;
            }
            if (active) {
                const int b0 = e_row, b1 = e_row + WQ;
                *reinterpret_cast<float2 *>(sND + b0) = make_float2(d[0], d[2]);
                *reinterpret_cast<float2 *>(sND + b1) = make_float2(d[1], d[3]);
                *reinterpret_cast<float2 *>(sQ + b0) = make_float2(0.f, 0.f);
                *reinterpret_cast<float2 *>(sQ + b1) = make_float2(0.f, 0.f);
                *reinterpret_cast<unsigned short *>(sC + b0) = (unsigned short)__byte_perm(code, 0u, 0x4420);
                *reinterpret_cast<unsigned short *>(sC + b1) = (unsigned short)__byte_perm(code, 0u, 0x4431);
            }
            rq_arrive_a(b_self + 8u * (unsigned)(rel & (RQ_RING - 1)));
            if (++sl == RQ_NL) { sl = 0; e_row = 2 * ld; } else e_row += ROW;
            st0 = st1;
            if (++st1 == RQ_STG) { st1 = 0; par1 ^= 1u; }
        }
    } else if (warp < RQ_SW + 8) {
        // ================= writer: owned lines -> U, V, p =================
        // U0, V0 and the mask of the line come from the writer's own TMA staging ring (second
        // producer below).  They were fetched through registers before: the re-read misses L2 more
        // often than not (ncu: 35 % read hit rate) and the register ring did not survive code
        // generation, so every line paid a DRAM round trip (74 % of the writer's stall samples).
        const int st = tid - RQ_WR0;
        const bool active = st < (P.TJ >> 2);
        const int w_lj = RQ_H + 4 * st, w_j = jr0 + w_lj;
        const bool col_ok = active && w_j < NY && !(P.xflags & 2);
        const bool full4 = w_j + 3 < NY;
        const unsigned b_self = rq_s32(bars + RQ_WROLE * RQ_RING), b_last = rq_s32(bars + RQ_LAST_ROLE(nst) * RQ_RING), b_wfull = rq_s32(wfull);
        // only owned lines are written, but the hand-off phases count every line: arrive for the halo lines first
        for (int rel = 0; rel < i0c - e0; rel++) rq_arrive_a(b_self + 8u * (unsigned)(rel & (RQ_RING - 1)));
        int sl = (i0c - e0) % RQ_NL;
        // this thread's cells: q index in a slot, staging offsets, global offset of line i0c
        const int q = w_lj >> 1;
        int e_row = sl * ROW + q, e_rowm = (sl == 0 ? RQ_NL - 1 : sl - 1) * ROW + q;
        int ws = 0;
        unsigned wpar = 0;
        const unsigned char *sb = wstg + 16 * st;
        const int oV = 4 * P.TJ, oM = 8 * P.TJ - 12 * st;    // byte offsets of V and the mask word from sb
        size_t o = (size_t)(i0c - g.i_alloc0) * PIT + w_j;
        const bool turb = P.turb > 0.0f;
        const float cp = P.cp;
        for (int r = i0c; r < i1c; r++) {
            const int rel = r - e0;
            {   // the last half sweep is past line r (RQ_PAIR: the last iteration has finished its step rel+1)
                const int wl = rel + RQ_PAIR;
                rq_mbar_wait_io(b_last + 8u * (unsigned)(wl & (RQ_RING - 1)), (unsigned)(wl >> 6) & 1u, (nst << 20) | rel);
            }
            rq_mbar_wait_io(b_wfull + 8u * (unsigned)ws, wpar, (32 << 20) | rel);     // U0, V0, mask of line r have landed
            if (col_ok) {
                const float4 u = *reinterpret_cast<const float4 *>(sb);
                const float4 v = *reinterpret_cast<const float4 *>(sb + oV);
                const unsigned m4 = *reinterpret_cast<const unsigned *>(sb + oM);
                float pin[4] = {0.f, 0.f, 0.f, 0.f};
                if (P.Pin) unpack(ld4(P.Pin + o), pin);
                const float2 ev = *reinterpret_cast<const float2 *>(sQ + e_row);
                const float2 od = *reinterpret_cast<const float2 *>(sQ + e_row + WQ);
                const float2 evm = *reinterpret_cast<const float2 *>(sQ + e_rowm);
                const float2 odm = *reinterpret_cast<const float2 *>(sQ + e_rowm + WQ);
                const float ql = sQ[e_row + WQ - 1];                   // column lj-1 (odd parity, index q-1)
                const float qc[4] = { ev.x, od.x, ev.y, od.y }, qx[4] = { evm.x, odm.x, evm.y, odm.y };
                const float uu[4] = { u.x, u.y, u.z, u.w }, vv[4] = { v.x, v.y, v.z, v.w };
                const bool line_first = (r == 0);
                float pu[4], pv[4], pp[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned m = m4 >> (8 * k);
                    const float qym = (k == 0) ? ql : qc[k - 1 < 0 ? 0 : k - 1];
                    const float a = (m & MK_XM) ? qc[k] : 0.0f;
                    const float b = ((m & MK_C) && !line_first) ? qx[k] : 0.0f;
                    const float t1 = uu[k] - a;
                    pu[k] = t1 + b;
                    const float a2 = (m & MK_YM) ? qc[k] : 0.0f;
                    const float b2 = ((m & MK_C) && (w_j + k) > 0) ? qym : 0.0f;
                    const float t2 = vv[k] - a2;
                    pv[k] = t2 + b2;
                    pp[k] = __fmaf_rn(cp, qc[k], pin[k]);
                }
                if (turb && r >= 1 && r <= NX - 2) {                 // fused addTurbulence (fluid.go:496-526)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const unsigned m = m4 >> (8 * k);
                        const int jj = w_j + k;
                        if ((m & MK_C) && jj >= 1 && jj <= NY - 2) {
                            const float u2 = pu[k] * pu[k], v2 = pv[k] * pv[k];
                            const float localVel = sqrtf(u2 + v2);
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
                if (full4) {
                    *reinterpret_cast<float4 *>(P.Uo + o) = make_float4(pu[0], pu[1], pu[2], pu[3]);
                    *reinterpret_cast<float4 *>(P.Vo + o) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                    *reinterpret_cast<float4 *>(P.Po + o) = make_float4(pp[0], pp[1], pp[2], pp[3]);
                } else {
                    for (int k = 0; k < 4 && w_j + k < NY; k++) { P.Uo[o + k] = pu[k]; P.Vo[o + k] = pv[k]; P.Po[o + k] = pp[k]; }
                }
            }
            rq_arrive_a(b_self + 8u * (unsigned)(rel & (RQ_RING - 1)));       // slots and staging of line r are free
            e_rowm = e_row;
            if (++sl == RQ_NL) { sl = 0; e_row = q; } else e_row += ROW;
            if (++ws == RQ_WSTG) { ws = 0; wpar ^= 1u; sb = wstg + 16 * st; } else sb += WSTGB;
            o += PIT;
        }
    } else if (tid == RQ_P2) {
        // ================= second producer: U0, V0, mask of the owned lines for the writer =================
        const int c0 = strip * P.TJ;                                          // first column the writer owns
        const int nc = min(P.TJ, PIT - c0);                                   // multiple of 16 (pitch % 32 == 0)
        const unsigned bF = (unsigned)nc * 4, bM = (unsigned)nc, bytes = 2 * bF + bM;
        const bool skip = (P.xflags & 4) || nc <= 0;
        const long long o0 = (long long)(i0c - g.i_alloc0) * PIT + c0;
        const float *gU = P.U + o0, *gV = P.V + o0;
        const unsigned char *gM = P.mask + o0;
        int ws = 0;
        for (int r = i0c; r < i1c; r++) {
            if (r - i0c >= RQ_WSTG) rq_wait_line(bars, RQ_WROLE, r - RQ_WSTG - e0);   // the writer is done with this stage
            unsigned char *sb = wstg + ws * WSTGB;
            if (!skip) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                rq_mbar_expect_tx(wfull + ws, bytes);
                rq_tma_load(sb, gU, bF, wfull + ws);
                rq_tma_load(sb + 4 * P.TJ, gV, bF, wfull + ws);
                rq_tma_load(sb + 8 * P.TJ, gM, bM, wfull + ws);
            } else {
                rq_mbar_arrive(wfull + ws);
            }
            gU += PIT; gV += PIT; gM += PIT;
            if (++ws == RQ_WSTG) ws = 0;
        }
    } else if (tid == RQ_P1) {
        // ================= producer: TMA bulk copies into the staging ring =================
        // line rel goes to staging slot rel % RQ_STG once the loader is past line rel - RQ_STG
        // (the loader reads staging slot(rel) for lines rel-1 and rel).  One thread: everything
        // that does not change from line to line is hoisted, the loop body is ~30 instructions.
        const int cj0 = jr0 < 0 ? 0 : jr0;                                   // first global column copied
        const int cjU = min(jr0 + WL, PIT), cjV = min(jr0 + WL + 4, PIT);    // one past the last column (U, mask / V)
        const int off = cj0 - jr0;                                            // staging column of global column cj0
        const unsigned bU = (unsigned)(cjU - cj0) * 4, bV = (unsigned)(cjV - cj0) * 4, bM = (unsigned)(cjU - cj0);
        const unsigned bytes = bU + bV + bM;
        // lines that exist in this rank's planes: relative [relA, relB)
        const int lineA = max(0, g.i_alloc0), lineB = min(NX, g.i_alloc0 + g.lines_alloc);
        const int relA = lineA - e0, relB = (cjU > cj0 && !(P.xflags & 4)) ? lineB - e0 : -1;
        unsigned dU[RQ_STG], dV[RQ_STG], dM[RQ_STG], fb[RQ_STG];
#pragma unroll
        for (int k = 0; k < RQ_STG; k++) {
            unsigned char *s0 = stg + k * STG;
            dU[k] = rq_s32(reinterpret_cast<float *>(s0) + off);
            dV[k] = rq_s32(reinterpret_cast<float *>(s0) + WL + off);
            dM[k] = rq_s32(s0 + (size_t)(2 * WL + 4) * 4 + off);
            fb[k] = rq_s32(full + k);
        }
        const long long o0 = (long long)(e0 - g.i_alloc0) * PIT + cj0;       // offset of relative line 0 (may be negative)
        const float *gU = P.U + o0, *gV = P.V + o0;
        const unsigned char *gM = P.mask + o0;
        const unsigned lbase = rq_s32(bars);                                  // loader hand-off barriers (role 0)
        for (int rel0 = 0; rel0 <= nproc; rel0 += RQ_STG) {
#pragma unroll
            for (int k = 0; k < RQ_STG; k++) {
                const int rel = rel0 + k;
                if (rel > nproc) break;
                if (rel >= RQ_STG) {
                    const int w = rel - RQ_STG;
                    rq_mbar_wait_io(lbase + 8u * (unsigned)(w & (RQ_RING - 1)), (unsigned)(w / RQ_RING) & 1u, w);
                }
                if (rel >= relA && rel < relB) {
                    // order prior generic-proxy reads of this staging slot before the async-proxy writes
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb[k]), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dU[k]), "l"(gU), "r"(bU), "r"(fb[k]) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dV[k]), "l"(gV), "r"(bV), "r"(fb[k]) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dM[k]), "l"(gM), "r"(bM), "r"(fb[k]) : "memory");
                } else {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb[k]) : "memory");
                }
                gU += PIT; gV += PIT; gM += PIT;
            }
        }
    }
}
