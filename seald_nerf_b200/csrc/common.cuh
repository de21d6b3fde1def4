// common.cuh — shared helpers for the sm_100a kernels of libseald_b200.so.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seald_b200.h"

#define SEALD_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

namespace seald {

template <typename T>
__host__ __device__ __forceinline__ T div_up(T a, T b) { return (a + b - 1) / b; }

inline cudaStream_t to_stream(seald_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int launch_status() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
    return 0;
}

__device__ __forceinline__ float clampf(const float x, const float lo, const float hi) {
    return fminf(hi, fmaxf(lo, x));
}

// warp-level inclusive scan (sum) of an unsigned value
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, const uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= (uint32_t)o) v += n;
    }
    return v;
}

// streaming (read-once) loads / write-once stores: keep L1 for the gathers
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

}  // namespace seald
