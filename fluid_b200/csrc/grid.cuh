// grid.cuh -- plane geometry shared by every kernel file.  Free of CUDA runtime includes so that
// the host-emulation harness of the tests (tests/emul/) can compile kernel headers with g++.
#pragma once
#include <stddef.h>
#include <stdint.h>

struct Grid {
    int NX, NY;        // global NumX, NumY (fluid.go:48-49)
    int pitch;         // floats per allocated line
    int i_alloc0;      // global i of allocated line 0
    int lines_alloc;   // allocated lines
    int i_lo, i_hi;    // owned global lines [i_lo, i_hi)
    __host__ __device__ __forceinline__ size_t at(int i, int j) const {
        return (size_t)(i - i_alloc0) * (size_t)pitch + (size_t)j;
    }
};
