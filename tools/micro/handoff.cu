// Hand-off latency between two warps of one CTA (round 2, profiles/r02_rbq_stream.md): warp 0 and warp 1 pass a token back and forth
// N times; cycles per ONE-WAY hand-off = elapsed / (2 N).
//   mode 0: mbarrier, every lane arrives (count 32), every lane polls try_wait.parity       (what k_rbq_fused does)
//   mode 1: mbarrier, lane 0 arrives after __syncwarp (count 1), lane 0 polls, __syncwarp    (-DRQ_ELECT)
//   mode 2: a flag in shared memory: st.release / ld.acquire by lane 0, __syncwarp
//   mode 3: mode 0 with a second, already completed try_wait per step (the cost of a wait that does not have to wait)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o handoff tools/micro/handoff.cu ; run: ./handoff
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(unsigned long long *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mb_arrive(unsigned b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ bool mb_try(unsigned b, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mb_wait(unsigned b, unsigned parity) { while (!mb_try(b, parity)) { } }

template <int MODE>
__global__ void k_pingpong(int n, long long *out)
{
    __shared__ unsigned long long bars[2];
    __shared__ unsigned long long done_bar;
    __shared__ volatile unsigned flags[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mb_init(&bars[0], MODE == 1 ? 1 : 32); mb_init(&bars[1], MODE == 1 ? 1 : 32); mb_init(&done_bar, 32);
        flags[0] = flags[1] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    if (MODE == 3 && warp == 0) mb_arrive(s32(&done_bar));      // phase 0 of done_bar completes at once: waits on it never wait
    __syncthreads();
    const unsigned mine = s32(&bars[warp]), other = s32(&bars[1 - warp]);
    const long long t0 = clock64();
    for (int k = 0; k < n; k++) {
        const unsigned parity = (unsigned)k & 1u;
        if (warp == 0) {
            // send token k, then wait for its echo
            if (MODE == 0 || MODE == 3) { mb_arrive(other); mb_wait(mine, parity); }
            else if (MODE == 1) { __syncwarp(); if (lane == 0) { mb_arrive(other); mb_wait(mine, parity); } __syncwarp(); }
            else {
                __syncwarp();
                if (lane == 0) {
                    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(s32((const void *)&flags[1])), "r"((unsigned)k + 1) : "memory");
                    unsigned v;
                    do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(s32((const void *)&flags[0])) : "memory"); } while (v != (unsigned)k + 1);
                }
                __syncwarp();
            }
        } else {
            if (MODE == 0 || MODE == 3) { mb_wait(mine, parity); mb_arrive(other); }
            else if (MODE == 1) { if (lane == 0) mb_wait(mine, parity); __syncwarp(); if (lane == 0) mb_arrive(other); }
            else {
                if (lane == 0) {
                    unsigned v;
                    do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(s32((const void *)&flags[1])) : "memory"); } while (v != (unsigned)k + 1);
                }
                __syncwarp();
                if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(s32((const void *)&flags[0])), "r"((unsigned)k + 1) : "memory");
            }
        }
        if (MODE == 3) mb_wait(s32(&done_bar), 0);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main()
{
    long long *d, h;
    cudaMalloc(&d, sizeof(long long));
    const int n = 20000;
    const char *names[4] = { "mbarrier, 32 arrivals, 32 pollers", "mbarrier, elected lane + __syncwarp", "shared-memory flag (st.release / ld.acquire), elected lane",
                             "mode 0 + one satisfied try_wait per step" };
    for (int mode = 0; mode < 4; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            if (mode == 0) k_pingpong<0><<<1, 64>>>(n, d);
            if (mode == 1) k_pingpong<1><<<1, 64>>>(n, d);
            if (mode == 2) k_pingpong<2><<<1, 64>>>(n, d);
            if (mode == 3) k_pingpong<3><<<1, 64>>>(n, d);
            cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
        }
        printf("mode %d (%s): %.1f cycles per one-way hand-off%s\n", mode, names[mode], (double)h / (2.0 * n),
               mode == 3 ? " (difference to mode 0 x 2 = one satisfied wait)" : "");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
