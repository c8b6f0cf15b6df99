/* Host shim standing in for <cuda_runtime.h> when the DEVICE engine source
 * (pacbot-rl_b200/csrc/pb_engine.cuh) is compiled with g++ for tests/test_host_engine.py.
 * TEST INFRASTRUCTURE ONLY: nothing in the product links against this. */
#pragma once
#include <cstdint>
#include <cstdlib>

#define __host__
#define __device__
#define __forceinline__ inline
#define __constant__
#define __restrict__

struct uint4 {
    uint32_t x, y, z, w;
};

static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
