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
// (constant i, contiguous in j) stream through a ring of RQ_NL slots in shared memory
// (q, -D0, neighbour count: 9 B per cell, 512 columns per slot).  24 warps form a software
// pipeline WITHOUT block-wide barriers:
//   loader       4 warps, line-interleaved (warp k: lines k, k+4, ...): TMA staging -> -D0, neighbour
//                count (bits 5-7 of the mask byte), q = 0 in slot(line).  Lane 0 of the warp that
//                has consumed a staging slot refills it with the line RQ_STG further on: bulk copies
//                (cp.async.bulk + mbarrier complete_tx) of the U, V, mask segments of that line
//   stage t < 8  ITERATION t, RQ_SPLIT = 2 warps side by side (128 same-parity cells each, 4 per lane).
//                Step `rel` is the red half sweep on line rel ("first") followed by the black half
//                sweep on line rel-1 ("second"); both touch the columns of one parity.  Everything
//                `second` needs is in registers: its own vector is the `dn` of `first` (the `up` loaded
//                two steps ago), its `up` is the result of `first`, left / right and `dn` are the
//                results of the two previous steps; `first` reads only its own vector and the line
//                above.  The neighbour beyond a lane's four cells comes from the next lane by shuffle.
//                A stage may run step rel once its predecessor has finished step rel+2; the two warps
//                of a stage meet at a named barrier (bar.sync) once per step
//   writer       4 warps, line-interleaved: slot(r), slot(r-1) + U0, V0, mask (its own TMA staging
//                ring, refilled by lane 0 of the warp that owns the slot) -> U, V, p in global
// Hand-offs are per-line mbarriers (a ring of 64 per role; every lane of the warp(s) that own the
// line arrives, every lane of a consumer polls with try_wait); the loader reuses a slot once the
// writer (or, for halo lines, the last iteration) is past it.  Waits are bounded: a pipeline that
// stops latches a debug record and the host returns FB_ERR_CUDA instead of hanging the GPU.
// Even and odd columns live in separate arrays so one colour is contiguous: a lane updates 4
// consecutive same-colour cells with LDS.128 / STS.128 and packed FADD2 / FFMA2 (sm_100a fp32x2,
// bit-identical to the scalar operations).
//
// Why this shape (ncu, 4098^2, 8 iterations; time of the solve):
//   face form (rb_fused.cuh), ~110 instructions per cell update, issue-bound          1.11 ms
//   pressure form, one __syncthreads per line (barrier = 3.4 of 9 stall cycles)       0.47
//   one warp per half sweep, per-line mbarrier hand-offs, strips x chunks = one wave  0.31
//   writer inputs through TMA (its re-read of U0, V0 missed L2 two times in three)    0.226
//   neighbour counts baked into the mask, lean loader / writer loops                  0.211
//   one warp per iteration (two half sweeps fused, neighbours carried in registers):
//   2.5 LDS.128 per four updates instead of 5, 8 pipeline stages instead of 16        0.208
//   line-interleaved loader / writer warps (four line periods per line) + two warps
//   per stage: the time with the sweeps switched off fell 0.148 -> 0.120              0.191
//   edge neighbours by shuffle (the scalar LDS at a 16-byte stride was a 4-way bank
//   conflict: 8 of a step's 34 shared-memory wavefronts)                              0.188
//   TMA issued by the loader / writer warps (a single producer thread spent ~700
//   cycles per line on wait + proxy fence + three bulk copies)                        0.182
// Where the time goes now (profiles/): shared-memory wavefronts 68 % of peak (721 per line: 448 the
// sweeps, the rest TMA, loader, writer), issue slots 65 %; with ONE iteration the kernel takes
// 0.120 ms (writer warps 90 % busy), every further iteration adds ~0.01 ms.
// Measured and NOT adopted (FLUIDB200_RBQ_X switches roles off for such experiments):
//   progress counters in shared memory polled with LDS instead of mbarriers: 0.224 (a poll every
//   ~40 cycles per waiting warp is shared-memory traffic; try_wait suspends the warp ~100 cycles);
//   sleeping after a failed poll, a non-blocking test_wait first, one arrive per warp: no change;
//   issuing the loads of lines rel, rel-1 before the hand-off wait: 0.195; 24 / 32 ring slots,
//   deeper staging: no change; re-reading the neighbour vectors instead of carrying them (LSU
//   pipe 79 %); turbulence fused into the writer (solve + turbulence 3.52 -> 4.20 ms at 16386^2);
//   8 writer warps instead of 4 (-DRQ_WW=8, 69-72 registers, no spills): 0.184 vs 0.181, with a 16-deep
//   writer staging ring and 24 slots 0.216; 24 slots alone 0.212 (28 it stays).  RQ_LW = 8 needs a
//   staging ring deeper than RQ_STG = 8 (a loader warp refills the slot of its next-but-one line) and
//   is NOT a valid build as is: its result differs.
#pragma once
#include "kernels.cuh"
#include "advect_fused.cuh"

#ifndef RQ_NL
#define RQ_NL 28          // line slots
#endif
#define RQ_H 16
#define RQ_NIT 8          // iterations per pass = pipeline stages
#ifndef RQ_SPLIT
#define RQ_SPLIT 2        // warps per stage (1 or 2): a line is 2 groups of 128 same-parity cells, 4 per lane
#endif
#ifndef RQ_WL
#define RQ_WL 512         // columns per slot, whatever TJ is: every lane of a sweep always owns its groups of four cells
#endif
#define RQ_G (RQ_WL / 256)          // groups of 128 same-parity cells in a line (4 per lane)
#define RQ_Q (RQ_G / RQ_SPLIT)      // groups per sweep warp
#ifndef RQ_MINB
#define RQ_MINB 1         // CTAs per SM the kernel is compiled for (experiment: RQ_WL = 256, RQ_SPLIT = 1, RQ_MINB = 2)
#endif
static_assert(RQ_WL % 256 == 0 && RQ_G >= RQ_SPLIT && RQ_G % RQ_SPLIT == 0, "a sweep warp owns whole groups of 128 same-parity cells");
#define RQ_SW (RQ_NIT * RQ_SPLIT)   // sweep warps
#ifndef RQ_LW
#define RQ_LW 4           // loader warps (line-interleaved)
#endif
#ifndef RQ_WW
#define RQ_WW 4           // writer warps (line-interleaved)
#endif
#define RQ_THREADS (32 * (RQ_SW + RQ_LW + RQ_WW))
#ifndef RQ_RING_LOG2
#define RQ_RING_LOG2 (RQ_NL < 32 ? 6 : 7)
#endif
#define RQ_RING (1 << RQ_RING_LOG2)       // hand-off mbarriers per role; > 2 * RQ_NL (see "hand-offs" below)
// Experiments (tools/r02_rbq_ring.sh): RQ_LSPLIT / RQ_WSPLIT warps of the loader / writer share ONE line (each takes
// its part of the column groups) instead of taking whole lines in turn; a line then stays RQ_LW / RQ_LSPLIT line
// periods in the role instead of RQ_LW.  1 = line-interleaved (the measured default).
#ifndef RQ_LSPLIT
#define RQ_LSPLIT 1
#endif
#ifndef RQ_WSPLIT
#define RQ_WSPLIT 1
#endif
#define RQ_TJ_MAX (RQ_WL - 48)   // 464: multiple of 16; TJ + 2 * RQ_H + 16 <= RQ_WL (16-byte granules for the TMA copies of the mask)
#ifndef RQ_STG
#define RQ_STG 8          // staging ring depth (lines in flight through TMA); power of two
#endif
#ifndef RQ_WSTG
#define RQ_WSTG 8         // writer staging ring depth; power of two
#endif
static_assert((RQ_STG & (RQ_STG - 1)) == 0 && (RQ_WSTG & (RQ_WSTG - 1)) == 0, "staging ring depths are powers of two");
static_assert(RQ_STG >= 2 * RQ_LW && RQ_STG % RQ_LW == 0, "a loader warp keeps two of its own lines in flight; shallower rings were measured to give wrong results");
static_assert(RQ_WSTG % RQ_WW == 0, "a writer warp refills the staging slot of its own next lines");
static_assert((RQ_RING & (RQ_RING - 1)) == 0 && RQ_RING > 2 * RQ_NL, "a parity wait must refer to the current or the preceding phase");
static_assert(RQ_LW % RQ_LSPLIT == 0 && (RQ_WL / 128) % RQ_LSPLIT == 0 && RQ_STG % (RQ_LW / RQ_LSPLIT) == 0, "loader split");
static_assert(RQ_WW % RQ_WSPLIT == 0 && (RQ_WL / 128) % RQ_WSPLIT == 0 && RQ_WSTG % (RQ_WW / RQ_WSPLIT) == 0, "writer split");
// shared memory at TJ = 464: 28 slots * 512 * 9 B (q, -D0, neighbour count) = 126 KB, loader staging
// 8 * (512*9 + 16) B = 36.1 KB, writer staging 8 * 464 * 9 B = 32.6 KB, hand-off mbarriers 5 KB, wd/s table: 200 KB
__host__ __device__ __forceinline__ size_t rq_stage_bytes(int WL) { return (size_t)WL * 9 + 16; }
__host__ __device__ __forceinline__ size_t rq_wstage_bytes(int TJ) { return (size_t)TJ * 9; }   // U0, V0, mask of TJ columns
__host__ __device__ __forceinline__ size_t rq_smem_bytes(int WL, int TJ)
{
    return (size_t)RQ_NL * WL * 9 + RQ_STG * rq_stage_bytes(WL) + RQ_WSTG * rq_wstage_bytes(TJ) + 8 * (RQ_STG + RQ_WSTG) +
           8 * (RQ_NIT + 2) * RQ_RING + 16 * 8 * 4 + 64;
}

struct RBQ {
    Grid g;
    const float *U, *V;          // field the pass starts from
    const float *Pin;            // nullptr: pressure known to be zero
    const unsigned char *mask;
    float *Uo, *Vo, *Po;
    float wd[16];                // omega*damping per half sweep
    float nwd[16], c4[16];       // -wd and wd/4 (k_rbq_stream takes them from the constant bank)
    float cp;
    int nstages, stage0;
    int TJ, WL, chunk, ib, ie;
    unsigned *stats;             // per-iteration max |div| (only when STATS)
    int *debug;                  // [0] != 0: a pipeline wait timed out, [1..5] say which
    int xflags;                  // experiments (FLUIDB200_RBQ_X): 1 skip sweeps, 2 skip writer I/O, 4 skip TMA, 8 / 16 skip the loader's reads / stores
    const float *noiseU, *noiseV;
    float turb;
#ifdef RQ_TRACE
    long long *trace;            // [role 0 .. 9][line 0 .. 511][4] clock64 stamps of CTA (0, 0)
#endif
};
#ifdef RQ_TRACE
#define RQ_T(role, line, k) do { if (P_trace && (threadIdx.x & 31) == 0 && (line) >= 0 && (line) < 512) P_trace[(((role) * 512) + (line)) * 4 + (k)] = clock64(); } while (0)
#else
#define RQ_T(role, line, k) do { } while (0)
#endif

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
__device__ __forceinline__ void rq_arrive_a(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
#ifndef RQ_TRYWAIT_HINT
#define RQ_TRYWAIT_HINT 0           // ns a waiting warp may sleep per poll (0: no hint, the hardware default: measured best, 0.172 against 0.177 ms)
#endif
__device__ __forceinline__ bool rq_mbar_try_a(unsigned bar, unsigned parity)
{
    unsigned ok;
#if RQ_TRYWAIT_HINT
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(bar), "r"(parity), "r"((unsigned)RQ_TRYWAIT_HINT) : "memory");
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
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void rq_tma_load(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- hand-offs -----------------------------------------------------------------------------------
// Roles: 0 = loader, 1 + t = stage t, RQ_WROLE = writer.  Every role has a ring of RQ_RING mbarriers, one per
// line / step (line r uses barrier r % RQ_RING, phase parity (r / RQ_RING) & 1); every lane of the warp(s)
// that own the line arrives, every lane of a consumer polls with try_wait.  Each role arrives for EVERY line
// 0 .. nproc, and no role can be more than RQ_NL lines ahead of another (the loader waits for the slot), so
// with RQ_RING > 2 * RQ_NL a parity wait always refers to the current or the immediately preceding phase.
#define RQ_ROLES (RQ_NIT + 2)
#define RQ_WROLE (RQ_NIT + 1)
// Bounded waits: a pipeline that stops making progress must never hang the GPU.  After ~2^21 failed polls
// (hundreds of milliseconds) the first waiter records who waited for what in P.debug and every waiter
// falls through; the host turns the flag into an error.
#ifdef RQ_DEBUG_GLOBAL
__device__ int *rq_debug;   // round-1 form (one module-global for all handles); kept for A/B
#else
__shared__ int *rq_debug;   // per CTA, set by thread 0 at kernel entry (RBQ::debug): handles on one device never share it
#endif
__device__ __noinline__ void rq_wait_slow(unsigned bar, unsigned parity, int tag)
{
    for (unsigned trip = 0;; trip++) {
#pragma unroll 1
        for (int k = 0; k < 1024; k++)
            if (rq_mbar_try_a(bar, parity)) return;
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
// try_wait suspends the warp in hardware for a while (~100 cycles per poll), which is what keeps the pollers
// off the shared-memory pipe: polling a counter with LDS instead (one wavefront every ~40 cycles per waiting
// warp) was measured SLOWER (0.224 against 0.192 ms), and sleeping between polls changes nothing.
__device__ __forceinline__ void rq_wait_a(unsigned bar, unsigned parity, int tag)
{
#pragma unroll 1
    for (int k = 0; k < 4096; k++)
        if (rq_mbar_try_a(bar, parity)) return;
    rq_wait_slow(bar, parity, tag);
}
// line / step `line` of the role whose ring starts at `ring`.  RQ_ELECT (experiment): ONE lane per warp arrives (after
// __syncwarp, which orders the other lanes' stores before its release) and one lane polls (the others take the acquire
// through the __syncwarp behind it): 32x fewer operations on the hand-off word.  Call sites are warp-uniform.
#ifndef RQ_ELECT
#define RQ_ELECT 0
#endif
#define RQ_ARRIVALS (RQ_ELECT ? 1 : 32)      // arrivals per warp and hand-off
__device__ __forceinline__ void rq_wait_line(unsigned ring, int line, int tag)
{
#if RQ_ELECT
    if ((threadIdx.x & 31) == 0)
        rq_wait_a(ring + 8u * (unsigned)(line & (RQ_RING - 1)), (unsigned)(line >> RQ_RING_LOG2) & 1u, tag | line);
    __syncwarp();
#else
    rq_wait_a(ring + 8u * (unsigned)(line & (RQ_RING - 1)), (unsigned)(line >> RQ_RING_LOG2) & 1u, tag | line);
#endif
}
// the same wait for a call site that only ONE lane reaches (the TMA issuers)
__device__ __forceinline__ void rq_wait_line_one(unsigned ring, int line, int tag)
{
    rq_wait_a(ring + 8u * (unsigned)(line & (RQ_RING - 1)), (unsigned)(line >> RQ_RING_LOG2) & 1u, tag | line);
}
__device__ __forceinline__ void rq_done_line(unsigned ring, int line)
{
#if RQ_ELECT
    __syncwarp();
    if ((threadIdx.x & 31) == 0) rq_arrive_a(ring + 8u * (unsigned)(line & (RQ_RING - 1)));
#else
    rq_arrive_a(ring + 8u * (unsigned)(line & (RQ_RING - 1)));
#endif
}

// ---- one cell update, four cells at a time -------------------------------------------------------
template <int A>
__device__ __forceinline__ float4 rq_update(const float4 qo, const float4 up, const float4 dn, const float4 ot, const float ox,
                                            const float4 nd, const unsigned code, const float2 nwd2, const float2 c44,
                                            const float *__restrict__ tw, float4 &t_out)
{
    // nb = ((q[i-1,j] + q[i+1,j]) + q[i,j-1]) + q[i,j+1];  t = nb - D0.  The neighbour vector that is shifted by one cell
    // against the register pairs (left for even columns, right for odd ones) is added with scalar FADDs: assembling the
    // misaligned pairs a packed add needs cost two moves per pair (~25 moves per sweep step, ncu round 1)
    float2 nb01 = __fadd2_rn(make_float2(dn.x, dn.y), make_float2(up.x, up.y));
    float2 nb23 = __fadd2_rn(make_float2(dn.z, dn.w), make_float2(up.z, up.w));
    if (A) {               // odd columns: left = (o0, o1, o2, o3), right = (o1, o2, o3, ox)
        nb01 = __fadd2_rn(nb01, make_float2(ot.x, ot.y));
        nb23 = __fadd2_rn(nb23, make_float2(ot.z, ot.w));
        nb01.x = nb01.x + ot.y; nb01.y = nb01.y + ot.z; nb23.x = nb23.x + ot.w; nb23.y = nb23.y + ox;
    } else {               // even columns: left = (ox, o0, o1, o2), right = (o0, o1, o2, o3)
        nb01.x = nb01.x + ox; nb01.y = nb01.y + ot.x; nb23.x = nb23.x + ot.y; nb23.y = nb23.y + ot.z;
        nb01 = __fadd2_rn(nb01, make_float2(ot.x, ot.y));
        nb23 = __fadd2_rn(nb23, make_float2(ot.z, ot.w));
    }
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
#ifdef RQ_TRACE
    long long *trace; int role;
#endif
    float *sQ; const unsigned char *sC;
    int WQ, NDO, lane4, lane, TJ;
    float2 nwd1, c41, nwd2, c42;        // (-wd, -wd) and (wd/4, wd/4) of the two half sweeps
    const float *tw1, *tw2;
    // hand-offs: the predecessor's ring to wait on, our own to arrive on, the sibling barrier
    unsigned pred, mine;
    int lag, nproc, tag, bar_id;
    int e_dn, e_own, e_up, sl_up, ROW;  // slots of lines rel-1, rel, rel+1 (element offsets)
    int i0r, i1r;                       // owned lines, relative to e0 (STATS only)
    bool skip;
    float mymax;
};

// One step.  P2 (in: q_old[rel-1][A], out: the `up` loaded now), P1 (q_old[rel][1-A]), F2 (in: first(rel-2),
// out: first(rel)), F1 (first(rel-1)).  FIRST is false only for the step past the last line (line e1 is never
// swept and reads as q = 0), SECOND only for step 0.
// RQ_PAIRWAIT (experiment): iterations 1 .. 7 wait for their predecessor once per TWO steps (for its step rel+3, which
// implies rel+2: a stage finishes its steps in order) instead of once per step; needs RQ_NL >= 32.  Iteration 0 keeps one
// wait per step: the loader's warps finish their lines in any order.
#ifndef RQ_WEARLY
#define RQ_WEARLY 0        // experiment: the writer frees a line's ring slots before it computes and stores (see the writer)
#endif
#ifndef RQ_WUNROLL
#define RQ_WUNROLL 2
#endif
#define RQ_PRAGMA_(x) _Pragma(#x)
#define RQ_PRAGMA(x) RQ_PRAGMA_(x)
#ifndef RQ_PAIRWAIT
#define RQ_PAIRWAIT 0
#endif
template <int A, bool FIRST, bool SECOND, bool STATS, bool WAIT = true>
__device__ __forceinline__ void rq_pair_step(RQStage &S, const int rel, float4 (&P2)[RQ_Q], float4 (&P1)[RQ_Q], float4 (&F2)[RQ_Q],
                                             const float4 (&F1)[RQ_Q])
{
#ifdef RQ_TRACE
    long long *const P_trace = S.trace;
    RQ_T(S.role, rel, 0);
#endif
    if (WAIT || S.lag == 1) {
        // the loader has finished line rel+1 / the previous iteration has finished its step rel+2
        // (the loader's warps take the lines in turn, so "line 1 is loaded" says nothing about line 0: the
        // first step of the first iteration waits for both; every later line was waited for one step earlier)
        if (!SECOND && S.lag == 1) rq_wait_line(S.pred, 0, S.tag);
        rq_wait_line(S.pred, min(rel + S.lag, S.nproc), S.tag);
    }
#ifdef RQ_TRACE
    RQ_T(S.role, rel, 1);
#endif
    // the stores of the previous step by the other lanes of the stage (left / right neighbours of `second`)
    if (RQ_SPLIT > 1) asm volatile("bar.sync %0, %1;" ::"r"(S.bar_id), "n"(32 * RQ_SPLIT) : "memory");
    else __syncwarp();
#ifdef RQ_TRACE
    RQ_T(S.role, rel, 2);
#endif
    float *const sQ = S.sQ;
    const int own = S.e_own + A * S.WQ + S.lane4, oth = S.e_own + (1 - A) * S.WQ + S.lane4;
    const int upo = S.e_up + A * S.WQ + S.lane4;
    const int own2 = S.e_dn + A * S.WQ + S.lane4, oth2 = S.e_dn + (1 - A) * S.WQ + S.lane4 + (A ? 4 : -1);
    if (!SECOND) {         // step 0: q_old[0][other parity] is the one carried vector that was never an `up`
#pragma unroll
        for (int half = 0; half < RQ_Q; half++) P1[half] = *reinterpret_cast<const float4 *>(sQ + oth + 128 * half);
    }
    if (!S.skip) {
        float4 qo[RQ_Q], up[RQ_Q], nd[RQ_Q], nd2[RQ_Q];
        float ox[RQ_Q], ox2[RQ_Q];
        unsigned code[RQ_Q], code2[RQ_Q];
#pragma unroll
        for (int half = 0; half < RQ_Q; half++) {
            const int h0 = 128 * half;
            if (FIRST) {
                qo[half] = *reinterpret_cast<const float4 *>(sQ + own + h0);
                up[half] = *reinterpret_cast<const float4 *>(sQ + upo + h0);
                nd[half] = *reinterpret_cast<const float4 *>(sQ + own + h0 + S.NDO);
                // left / right neighbour beyond this lane's four cells: the next / previous lane holds it in a register
                // (a scalar LDS at a 16-byte stride is a 4-way bank conflict); only the lane at the end of the
                // warp's group reads shared memory (the other warp's cell, or the halo column)
                ox[half] = A ? __shfl_down_sync(0xffffffffu, P1[half].x, 1) : __shfl_up_sync(0xffffffffu, P1[half].w, 1);
                if (S.lane == (A ? 31 : 0)) ox[half] = sQ[oth + h0 + (A ? 4 : -1)];
                code[half] = *reinterpret_cast<const unsigned *>(S.sC + own + h0);
            }
            if (SECOND) {
                nd2[half] = *reinterpret_cast<const float4 *>(sQ + own2 + h0 + S.NDO);
                ox2[half] = A ? __shfl_down_sync(0xffffffffu, F1[half].x, 1) : __shfl_up_sync(0xffffffffu, F1[half].w, 1);
                if (S.lane == (A ? 31 : 0)) ox2[half] = sQ[oth2 + h0];
                code2[half] = *reinterpret_cast<const unsigned *>(S.sC + own2 + h0);
            }
        }
#pragma unroll
        for (int half = 0; half < RQ_Q; half++) {
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
    rq_done_line(S.mine, rel);
#ifdef RQ_TRACE
    RQ_T(S.role, rel, 3);
#endif
    S.e_dn = S.e_own; S.e_own = S.e_up;
    if (++S.sl_up == RQ_NL) { S.sl_up = 0; S.e_up = 0; } else S.e_up += S.ROW;
}

// All steps 0 .. nproc of one iteration.  A0 = column parity of the active cells at step 0; it alternates
// from step to step, and the carried vectors swap names instead of moving (loop unrolled by two).
template <int A0, bool STATS>
__device__ __forceinline__ void rq_pair_stage(RQStage &S)
{
    float4 pa[RQ_Q], pb[RQ_Q], fa[RQ_Q], fb[RQ_Q];
#pragma unroll
    for (int half = 0; half < RQ_Q; half++) pa[half] = pb[half] = fa[half] = fb[half] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nproc = S.nproc;                                       // >= 2 * RQ_H + 1
    rq_pair_step<A0, true, false, STATS>(S, 0, pa, pb, fa, fb);
    int rel = 1;
    for (; rel + 1 < nproc; rel += 2) {
        rq_pair_step<1 - A0, true, true, STATS>(S, rel, pb, pa, fb, fa);
        rq_pair_step<A0, true, true, STATS, !RQ_PAIRWAIT>(S, rel + 1, pa, pb, fa, fb);
    }
    if (rel < nproc) {
        rq_pair_step<1 - A0, true, true, STATS>(S, rel, pb, pa, fb, fa);
        rq_pair_step<A0, false, true, STATS>(S, rel + 1, pa, pb, fa, fb);
    } else {
        rq_pair_step<1 - A0, false, true, STATS>(S, rel, pb, pa, fb, fa);
    }
}

template <bool STATS>
__global__ void __launch_bounds__(RQ_THREADS, RQ_MINB) k_rbq_fused(const RBQ P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int WL = RQ_WL, WQ = WL >> 1, ROW = WL;            // elements per slot in one plane
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
    const unsigned ring_ld = rq_s32(bars), ring_wr = rq_s32(bars + RQ_WROLE * RQ_RING);

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
    const int nit = nst >> 1;                                 // iterations of this pass (nst and stage0 are even)
    const int nproc = e1 - e0;                                // lines each half sweep passes over
    const int own0 = i0c - e0, last_owned = (i1c - 1) - e0;   // owned lines, relative

    if (tid == 0) rq_debug = P.debug;
    // arrivals per phase: every lane of the warp(s) that own the line / step
    for (int k = tid; k < RQ_ROLES * RQ_RING; k += RQ_THREADS) {
        const int role = k / RQ_RING;
        rq_mbar_init(bars + k, RQ_ARRIVALS * (role == 0 ? RQ_LSPLIT : (role == RQ_WROLE ? RQ_WSPLIT : RQ_SPLIT)));
    }
    if (tid >= 32 && tid < 32 + RQ_STG) rq_mbar_init(full + tid - 32, 1);
    if (tid >= 64 && tid < 64 + RQ_WSTG) rq_mbar_init(wfull + tid - 64, 1);
    if (tid >= 128 && tid < 256) {
        const int k = tid - 128, ns = k & 7;
        const float rs = ns == 0 ? 0.0f : (ns == 1 ? 1.0f : (ns == 2 ? 0.5f : (ns == 3 ? (1.0f / 3.0f) : 0.25f)));
        tblw[k] = P.wd[k >> 3] * rs;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    if (warp < RQ_SW) {
        // ================= stage t = iteration t: half sweeps 2t (first) and 2t + 1 (second) =================
        const int t = warp / RQ_SPLIT, hw = warp % RQ_SPLIT;  // stage, and which part of the line this warp owns
        if (t >= nit) return;
        const int colour = P.stage0 & 1;
        RQStage S;
        S.sQ = sQ; S.sC = sC; S.WQ = WQ; S.NDO = RQ_NL * WL; S.lane4 = 4 * lane + 128 * RQ_Q * hw; S.lane = lane; S.TJ = P.TJ;
        S.nwd1 = make_float2(-P.wd[2 * t], -P.wd[2 * t]);
        S.nwd2 = make_float2(-P.wd[2 * t + 1], -P.wd[2 * t + 1]);
        { const float c1 = P.wd[2 * t] * 0.25f, c2 = P.wd[2 * t + 1] * 0.25f; S.c41 = make_float2(c1, c1); S.c42 = make_float2(c2, c2); }
        S.tw1 = tblw + 8 * (2 * t); S.tw2 = tblw + 8 * (2 * t + 1);
        S.pred = rq_s32(bars + t * RQ_RING); S.mine = rq_s32(bars + (1 + t) * RQ_RING);
        S.lag = t == 0 ? 1 : (RQ_PAIRWAIT ? 3 : 2); S.nproc = nproc; S.tag = t << 20; S.bar_id = 1 + t;
        S.e_dn = (RQ_NL - 1) * ROW; S.e_own = 0; S.e_up = ROW; S.sl_up = 1; S.ROW = ROW;
        S.i0r = own0; S.i1r = i1c - e0;
        S.skip = (P.xflags & 1) != 0;
        S.mymax = 0.0f;
#ifdef RQ_TRACE
        S.trace = (blockIdx.x == 0 && blockIdx.y == 0 && hw == 0) ? P.trace : nullptr; S.role = 1 + t;
#endif
        if ((colour + e0) & 1) rq_pair_stage<1, STATS>(S); else rq_pair_stage<0, STATS>(S);
        if (STATS) {
            const float mymax = warp_max(S.mymax);
            if (lane == 0 && mymax > 0.0f) atomicMax(P.stats + ((P.stage0 >> 1) + t), __float_as_uint(mymax));
        }
    } else if (warp < RQ_SW + RQ_LW) {
        // ================= loader: staging -> slot, lines e0 .. e1 =================
        // Loader warp k takes lines k, k + RQ_LW, ...: a lane owns 4 columns in each of the 4 groups of 128, so a
        // warp has RQ_LW line periods for its line and the groups are independent work.
        const int lw = warp - RQ_SW;
        constexpr int LGRP = RQ_LW / RQ_LSPLIT;               // lines in flight in the loader
        constexpr int LCG = (RQ_WL / 128) / RQ_LSPLIT;        // column groups of a line per warp
#if RQ_LSPLIT > 1
        const int lgrp = lw / RQ_LSPLIT, lpart = lw % RQ_LSPLIT;
        const bool l_issuer = lane == 0 && lpart == 0;        // the one lane that feeds this line's TMA slot
#else
        const int lgrp = lw;
        constexpr int lpart = 0;
#define l_issuer (lane == 0)
#endif
        // only interior lines inside this rank's slab hold updatable cells (the mask's count bits are
        // zero on the ring and in solids); line e1 is loaded but never swept
        const int live_lo = max(1, g.i_alloc0) - e0;
        const int live_hi = min(min(NX - 2, g.i_alloc0 + g.lines_alloc - 2), e1 - 1) - e0;
        const unsigned b_full = rq_s32(full), ring_last = rq_s32(bars + nit * RQ_RING);
        // TMA: lane 0 of the warp that has just consumed staging slot k refills it with the line RQ_STG further on (its
        // own next-but-one line: RQ_STG is a multiple of RQ_LW).  A single producer thread for all lines was the
        // bottleneck of the whole pipeline: ~700 cycles per line for wait + proxy fence + three bulk copies.
        const int cj0 = jr0 < 0 ? 0 : jr0;                                   // first global column copied
        const int cjU = min(jr0 + WL, PIT), cjV = min(jr0 + WL + 4, PIT);    // one past the last column (U, mask / V)
        const int off = cj0 - jr0;                                            // staging column of global column cj0
        const unsigned bU = (unsigned)(cjU - cj0) * 4, bV = (unsigned)(cjV - cj0) * 4, bM = (unsigned)(cjU - cj0);
        // lines that exist in this rank's planes: relative [relA, relB)
        const int relA = max(0, g.i_alloc0) - e0, relB = (cjU > cj0 && !(P.xflags & 4)) ? min(NX, g.i_alloc0 + g.lines_alloc) - e0 : -1;
        const long long o0 = (long long)(e0 - g.i_alloc0) * PIT + cj0;       // offset of relative line 0 (may be negative)
        auto stage_line = [&](int line) {                                    // one lane: line -> staging slot line % RQ_STG
            const int k = line & (RQ_STG - 1);
            unsigned char *sk = stg + k * STG;
            const unsigned fb = rq_s32(full + k);
            if (line >= relA && line < relB) {
                const long long o = o0 + (long long)line * PIT;
                // order the generic-proxy reads of this staging slot before the async-proxy writes
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                rq_mbar_expect_tx(full + k, bU + bV + bM);
                rq_tma_load(rq_s32(reinterpret_cast<float *>(sk) + off), P.U + o, bU, fb);
                rq_tma_load(rq_s32(reinterpret_cast<float *>(sk) + WL + off), P.V + o, bV, fb);
                rq_tma_load(rq_s32(sk + (size_t)(2 * WL + 4) * 4 + off), P.mask + o, bM, fb);
            } else {
                rq_mbar_arrive(full + k);
            }
        };
        if (l_issuer)
            for (int line = lgrp; line < RQ_STG && line <= nproc; line += LGRP) stage_line(line);
        int sl = lgrp;                                       // slot of line rel
        for (int rel = lgrp; rel <= nproc; rel += LGRP) {
            // slot(rel) last held line y-1 with y = rel - NL + 1.  Its readers: the stages up to the last one's step y
            // (which writes line y-1 for the last time) and, for owned lines, the writer at lines y-1 and y.
#ifdef RQ_TRACE
            long long *const P_trace = (blockIdx.x == 0 && blockIdx.y == 0) ? P.trace : nullptr;
            RQ_T(0, rel, 0);
#endif
            const int y = rel - RQ_NL + 1;
            if (y >= 0) {
                if (y - 1 >= own0 && y - 1 <= last_owned) rq_wait_line(ring_wr, y - 1, 40 << 20);
                if (y >= own0 && y <= last_owned) rq_wait_line(ring_wr, y, 41 << 20);
                else rq_wait_line(ring_last, y, 42 << 20);
            }
            // lines rel and rel+1 have landed (the producer stages every line 0 .. nproc)
            const int st0 = rel & (RQ_STG - 1), st1 = (rel + 1) & (RQ_STG - 1);
            rq_wait_a(b_full + 8u * (unsigned)st0, (unsigned)(rel / RQ_STG) & 1u, (30 << 20) | rel);
            if (rel < nproc) rq_wait_a(b_full + 8u * (unsigned)st1, (unsigned)((rel + 1) / RQ_STG) & 1u, (31 << 20) | rel);
#ifdef RQ_TRACE
            RQ_T(0, rel, 1);
#endif
            const bool live = rel >= live_lo && rel <= live_hi;
            const unsigned char *s0 = stg + st0 * STG, *s1 = stg + st1 * STG;
            const int e_slot = sl * ROW;
#pragma unroll
            for (int cgi = 0; cgi < LCG; cgi++) {
                const int cg = lpart * LCG + cgi;
                const int ld = lane + 32 * cg;
                const int j = jr0 + 4 * ld;
                float d[4] = {0.f, 0.f, 0.f, 0.f};
                unsigned code = 0;
                if (live && j >= 0 && j < PIT && !(P.xflags & 8)) {
                    const float *stU = reinterpret_cast<const float *>(s0) + 4 * ld, *stV = stU + WL;
                    const float4 u0 = *reinterpret_cast<const float4 *>(stU);
                    const float4 u1 = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(s1) + 4 * ld);
                    const float4 v = *reinterpret_cast<const float4 *>(stV);
                    const float v4 = (j + 4 < PIT) ? stV[4] : 0.0f;
                    const unsigned mk = *reinterpret_cast<const unsigned *>(s0 + (size_t)(2 * WL + 4) * 4 + 4 * ld);
                    code = (mk >> MK_CNT_SHIFT) & 0x07070707u;   // fluid neighbours of updatable cells, else 0
                    const float dv0 = ((u1.x - u0.x) + v.y) - v.x, dv1 = ((u1.y - u0.y) + v.z) - v.y;
                    const float dv2 = ((u1.z - u0.z) + v.w) - v.z, dv3 = ((u1.w - u0.w) + v4) - v.w;
                    d[0] = (code & 0x000000ffu) ? -dv0 : 0.0f;
                    d[1] = (code & 0x0000ff00u) ? -dv1 : 0.0f;
                    d[2] = (code & 0x00ff0000u) ? -dv2 : 0.0f;
                    d[3] = (code & 0xff000000u) ? -dv3 : 0.0f;
                }
                const int b0 = e_slot + 2 * ld, b1 = b0 + WQ;
                if (P.xflags & 16) continue;
                *reinterpret_cast<float2 *>(sND + b0) = make_float2(d[0], d[2]);
                *reinterpret_cast<float2 *>(sND + b1) = make_float2(d[1], d[3]);
                *reinterpret_cast<float2 *>(sQ + b0) = make_float2(0.f, 0.f);
                *reinterpret_cast<float2 *>(sQ + b1) = make_float2(0.f, 0.f);
                *reinterpret_cast<unsigned short *>(sC + b0) = (unsigned short)__byte_perm(code, 0u, 0x4420);
                *reinterpret_cast<unsigned short *>(sC + b1) = (unsigned short)__byte_perm(code, 0u, 0x4431);
            }
            rq_done_line(ring_ld, rel);
#ifdef RQ_TRACE
            RQ_T(0, rel, 2);
#endif
            __syncwarp();                                    // every lane is done with staging slot st0
            if (l_issuer && rel + RQ_STG <= nproc) {
#if RQ_LSPLIT > 1
                rq_wait_line_one(ring_ld, rel, 36 << 20);                         // the other warps of this line
#endif
                // the slot also was the "line above" of line rel-1, which another loader warp handles
                if (rel >= 1) rq_wait_line_one(ring_ld, rel - 1, 35 << 20);
                stage_line(rel + RQ_STG);
            }
#ifdef RQ_TRACE
            RQ_T(0, rel, 3);
#endif
            sl += LGRP; if (sl >= RQ_NL) sl -= RQ_NL;
        }
    } else if (warp < RQ_SW + RQ_LW + RQ_WW) {
        // ================= writer: owned lines -> U, V, p =================
        // Writer warp k takes the owned lines i0c + k, i0c + k + RQ_WW, ... (all their columns).  U0, V0 and the
        // mask of a line come from the writer's own TMA staging ring (second producer below): re-read through
        // registers they missed L2 two times in three and every line paid a DRAM round trip.
        const int ww = warp - RQ_SW - RQ_LW;
        constexpr int WGRP = RQ_WW / RQ_WSPLIT;               // lines in flight in the writer
        constexpr int WCG = (RQ_WL / 128) / RQ_WSPLIT;
#if RQ_WSPLIT > 1
        const int wgrp = ww / RQ_WSPLIT, wpart = ww % RQ_WSPLIT;
        const bool w_issuer = lane == 0 && wpart == 0;
#else
        const int wgrp = ww;
        constexpr int wpart = 0;
#define w_issuer (lane == 0)
#endif
        const unsigned b_wfull = rq_s32(wfull), ring_last = rq_s32(bars + nit * RQ_RING);
        // only owned lines are written, but the hand-off phases count every line: arrive for the halo lines first
        for (int rel = wgrp; rel < own0; rel += WGRP) rq_done_line(ring_wr, rel);
        int sl = (own0 + wgrp) % RQ_NL;
        const bool turb = P.turb > 0.0f;
        const float cp = P.cp;
        const int oV = 4 * P.TJ;
        // TMA of U0, V0, mask of the OWNED columns of an owned line into this warp's own staging slots (lane 0)
        const int c0 = strip * P.TJ;                                          // first column the writer owns
        const int nc = min(P.TJ, PIT - c0);                                   // multiple of 16 (pitch % 32 == 0)
        const unsigned bF = (unsigned)nc * 4, bMk = (unsigned)nc;
        const bool no_tma = (P.xflags & 4) || nc <= 0;
        const long long ow0 = (long long)(i0c - g.i_alloc0) * PIT + c0;
        auto stage_line = [&](int n) {
            const int k = n & (RQ_WSTG - 1);
            const unsigned sb = rq_s32(wstg + k * WSTGB), fb = rq_s32(wfull + k);
            if (!no_tma) {
                const long long o = ow0 + (long long)n * PIT;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                rq_mbar_expect_tx(wfull + k, 2 * bF + bMk);
                rq_tma_load(sb, P.U + o, bF, fb);
                rq_tma_load(sb + 4 * P.TJ, P.V + o, bF, fb);
                rq_tma_load(sb + 8 * P.TJ, P.mask + o, bMk, fb);
            } else {
                rq_mbar_arrive(wfull + k);
            }
        };
        if (w_issuer)
            for (int n = wgrp; n < RQ_WSTG && n < i1c - i0c; n += WGRP) stage_line(n);
        for (int n = wgrp; n < i1c - i0c; n += WGRP) {
            const int r = i0c + n, rel = r - e0;
#ifdef RQ_TRACE
            long long *const P_trace = (blockIdx.x == 0 && blockIdx.y == 0) ? P.trace : nullptr;
            RQ_T(9, rel, 0);
#endif
            rq_wait_line(ring_last, rel + 1, nst << 20);                // the last iteration has finished its step rel+1: line r is final
            const int ws = n & (RQ_WSTG - 1);
            rq_wait_a(b_wfull + 8u * (unsigned)ws, (unsigned)(n / RQ_WSTG) & 1u, (32 << 20) | rel);   // U0, V0, mask of line r have landed
#ifdef RQ_TRACE
            RQ_T(9, rel, 1);
#endif
            const int e_line = sl * ROW, e_linem = (sl == 0 ? RQ_NL - 1 : sl - 1) * ROW;
            const bool line_first = (r == 0);
#if RQ_WEARLY
            // RQ_WEARLY: everything the writer reads of the RING (q of lines r and r-1) goes into registers first and the
            // slots are handed back at once; U0, V0, mask come from the writer's own staging ring.  The time a line occupies
            // its slot, not the writer's throughput, is what the closed loader -> sweeps -> writer loop is short of.
            float2 r_ev[WCG], r_od[WCG], r_evm[WCG], r_odm[WCG];
            float r_ql[WCG];
#pragma unroll
            for (int cgi = 0; cgi < WCG; cgi++) {
                const int st = lane + 32 * (wpart * WCG + cgi);
                const int q = (RQ_H + 4 * st) >> 1, e_row = e_line + q, e_rowm = e_linem + q;
                r_ev[cgi] = *reinterpret_cast<const float2 *>(sQ + e_row);
                r_od[cgi] = *reinterpret_cast<const float2 *>(sQ + e_row + WQ);
                r_evm[cgi] = *reinterpret_cast<const float2 *>(sQ + e_rowm);
                r_odm[cgi] = *reinterpret_cast<const float2 *>(sQ + e_rowm + WQ);
                r_ql[cgi] = sQ[e_row + WQ - 1];
            }
            rq_done_line(ring_wr, rel);                               // the slots of line r are free
#endif
RQ_PRAGMA(unroll RQ_WUNROLL)
            for (int cgi = 0; cgi < WCG; cgi++) {
                const int cg = wpart * WCG + cgi;
                const int st = lane + 32 * cg;
                const int w_lj = RQ_H + 4 * st, w_j = jr0 + w_lj;
                if (st >= (P.TJ >> 2) || w_j >= NY || (P.xflags & 2)) continue;
                const int q = w_lj >> 1, e_row = e_line + q, e_rowm = e_linem + q;
                const unsigned char *sb = wstg + ws * WSTGB + 16 * st;
                const size_t o = (size_t)(r - g.i_alloc0) * PIT + w_j;
                const float4 u = *reinterpret_cast<const float4 *>(sb);
                const float4 v = *reinterpret_cast<const float4 *>(sb + oV);
                const unsigned m4 = *reinterpret_cast<const unsigned *>(sb + 2 * oV - 12 * st);
                float pin[4] = {0.f, 0.f, 0.f, 0.f};
                if (P.Pin) unpack(ld4(P.Pin + o), pin);
#if RQ_WEARLY
                const float2 ev = r_ev[cgi], od = r_od[cgi], evm = r_evm[cgi], odm = r_odm[cgi];
                const float ql = r_ql[cgi];
                (void)e_row; (void)e_rowm;
#else
                const float2 ev = *reinterpret_cast<const float2 *>(sQ + e_row);
                const float2 od = *reinterpret_cast<const float2 *>(sQ + e_row + WQ);
                const float2 evm = *reinterpret_cast<const float2 *>(sQ + e_rowm);
                const float2 odm = *reinterpret_cast<const float2 *>(sQ + e_rowm + WQ);
                const float ql = sQ[e_row + WQ - 1];                   // column lj-1 (odd parity, index q-1)
#endif
                const float qc[4] = { ev.x, od.x, ev.y, od.y }, qx[4] = { evm.x, odm.x, evm.y, odm.y };
                const float uu[4] = { u.x, u.y, u.z, u.w }, vv[4] = { v.x, v.y, v.z, v.w };
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
                if (w_j + 3 < NY) {
                    *reinterpret_cast<float4 *>(P.Uo + o) = make_float4(pu[0], pu[1], pu[2], pu[3]);
                    *reinterpret_cast<float4 *>(P.Vo + o) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                    *reinterpret_cast<float4 *>(P.Po + o) = make_float4(pp[0], pp[1], pp[2], pp[3]);
                } else {
                    for (int k = 0; k < 4 && w_j + k < NY; k++) { P.Uo[o + k] = pu[k]; P.Vo[o + k] = pv[k]; P.Po[o + k] = pp[k]; }
                }
            }
#if !RQ_WEARLY
            rq_done_line(ring_wr, rel);                               // the slots of line r are free
#endif
#ifdef RQ_TRACE
            RQ_T(9, rel, 2);
#endif
            __syncwarp();
#if RQ_WSPLIT > 1
            if (w_issuer && n + RQ_WSTG < i1c - i0c) {
                rq_wait_line_one(ring_wr, rel, 37 << 20);                         // the other warps of this line
                stage_line(n + RQ_WSTG);
            }
#else
            if (lane == 0 && n + RQ_WSTG < i1c - i0c) stage_line(n + RQ_WSTG);
#endif
            sl += WGRP; if (sl >= RQ_NL) sl -= RQ_NL;
        }
    }
}
#undef l_issuer
#undef w_issuer
