// mg_emul.cpp -- TEST INFRASTRUCTURE.  Compiles the CUDA kernel header
// fluid_b200/csrc/multigrid.cuh with g++ (-ffp-contract=off = nvcc --fmad=false) and runs its
// kernels thread by thread on the CPU with the launch geometry of solve_multigrid_vcycle
// (fluidb200.cu), so that the indexing and the float32 operation order of the very source the
// GPU runs can be checked against the oracle without a GPU.  Threads of a launch run in REVERSE
// order: a kernel whose threads are not independent would not reproduce the sequential result.
// Not part of the product; nothing outside tests/ builds or loads it.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static dim3 blockIdx, blockDim, threadIdx;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define FB_HOST_EMULATION
#include "../../fluid_b200/csrc/multigrid.cuh"

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

template <typename F>
static void launch(dim3 grid, dim3 block, F body)
{
    blockDim = block;
    for (int by = (int)grid.y - 1; by >= 0; by--)
        for (int bx = (int)grid.x - 1; bx >= 0; bx--)
            for (int ty = (int)block.y - 1; ty >= 0; ty--)
                for (int tx = (int)block.x - 1; tx >= 0; tx--) {
                    blockIdx = dim3(bx, by, 0);
                    threadIdx = dim3(tx, ty, 0);
                    body();
                }
}

// The part of one V-cycle between the smoothing sweeps (fluid.go:580-592) on dense [NX][NY]
// host arrays: restrict -> coarse solve (exact != 0: lexicographic by anti-diagonals, else
// red-black) -> prolongate + apply.  Planes are repacked with the device pitch (multiple of 32).
extern "C" void mg_emul_correct(float *U, float *V, const float *S, float *P, int NX, int NY, float cp, int exact)
{
    Grid g;
    g.NX = NX; g.NY = NY; g.pitch = cdiv(NY, 32) * 32; g.i_alloc0 = 0; g.lines_alloc = NX; g.i_lo = 0; g.i_hi = NX;
    const size_t pf = (size_t)NX * g.pitch;
    std::vector<float> dU(pf, 0.f), dV(pf, 0.f), dS(pf, 0.f), dP(pf, 0.f);
    for (int i = 0; i < NX; i++) {
        memcpy(&dU[g.at(i, 0)], U + (size_t)i * NY, NY * 4); memcpy(&dV[g.at(i, 0)], V + (size_t)i * NY, NY * 4);
        memcpy(&dS[g.at(i, 0)], S + (size_t)i * NY, NY * 4); memcpy(&dP[g.at(i, 0)], P + (size_t)i * NY, NY * 4);
    }
    CoarseGrid c;
    c.NX = (NX + 1) / 2; c.NY = (NY + 1) / 2; c.pitch = cdiv(c.NY, 32) * 32;
    const size_t cf = (size_t)c.NX * c.pitch;
    std::vector<float> rhs(cf, 7.f), cS(cf, 7.f), cP(cf, 7.f);     // garbage: the kernels must initialise what they read
    const int coarse_iters = 40;
    const float coarse_relaxation = 1.6f;
    const dim3 blk(128, 2, 1);
    launch(dim3(cdiv(c.NY, blk.x), cdiv(c.NX, blk.y)), blk,
           [&] { k_mg_restrict(g, c, dU.data(), dV.data(), dS.data(), dP.data(), rhs.data(), cS.data(), cP.data()); });
    if (exact) {
        const int tau_last = (c.NX - 2) + (c.NY - 2) + 2 * (coarse_iters - 1);
        const dim3 dgrid(cdiv(c.NX - 2 > 0 ? c.NX - 2 : 1, 128), coarse_iters);
        if (c.NX > 2 && c.NY > 2)
            for (int tau = 2; tau <= tau_last; tau++)
                launch(dgrid, dim3(128), [&] { k_mg_coarse_diag(c, cP.data(), cS.data(), rhs.data(), tau, coarse_iters, coarse_relaxation); });
    } else {
        const dim3 cgrid(cdiv(c.NY / 2 + 1, blk.x), cdiv(c.NX - 2 > 0 ? c.NX - 2 : 1, blk.y));
        for (int it = 0; it < coarse_iters; it++)
            for (int colour = 0; colour < 2; colour++)
                launch(cgrid, blk, [&] { k_mg_coarse_redblack(c, cP.data(), cS.data(), rhs.data(), colour, coarse_relaxation); });
    }
    launch(dim3(cdiv(g.NY, blk.x), cdiv(g.NX, blk.y)), blk,
           [&] { k_mg_apply(g, c, dU.data(), dV.data(), dS.data(), dP.data(), cP.data(), cp); });
    for (int i = 0; i < NX; i++) {
        memcpy(U + (size_t)i * NY, &dU[g.at(i, 0)], NY * 4); memcpy(V + (size_t)i * NY, &dV[g.at(i, 0)], NY * 4);
        memcpy(P + (size_t)i * NY, &dP[g.at(i, 0)], NY * 4);
    }
}
