// Shared device/host helpers for libradarslam_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/radar_slam_b200.h"

void rs_set_error(const char* fmt, ...);

#define RS_CHECK_ARG(cond, ...)                      \
    do {                                             \
        if (!(cond)) {                               \
            rs_set_error(__VA_ARGS__);               \
            return RS_EINVAL;                        \
        }                                            \
    } while (0)

#define RS_CHECK_LAUNCH(name)                                                        \
    do {                                                                             \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            rs_set_error("%s: %s", name, cudaGetErrorString(e__));                   \
            return RS_ECUDA;                                                         \
        }                                                                            \
    } while (0)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

__device__ __forceinline__ uint32_t rs_make_key(int a, int r, int d) {
    return ((uint32_t)a << 24) | ((uint32_t)r << 12) | (uint32_t)d;
}
__device__ __forceinline__ void rs_split_key(uint32_t k, int& a, int& r, int& d) {
    a = (int)(k >> 24);
    r = (int)((k >> 12) & 0xFFFu);
    d = (int)(k & 0xFFFu);
}

// RDS device layout: rds[F][R][A][D] ("range-major planes"): for every range bin the A antenna rows of D Doppler
// cells follow each other.  Element (f, r, d, a); the A-channel snapshot of a cell is A elements at stride D.
__device__ __forceinline__ size_t rs_rds_index(int f, int r, int d, int a, int R, int D, int A) {
    return (((size_t)f * R + r) * A + a) * D + d;
}

static inline int rs_smem_optin_limit() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return v;
}
static inline int rs_sm_count() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
}
