// common.cuh -- shared helpers for libb2unet (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/b2unet.h"

namespace b2 {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0) {
    char buf[512];
    snprintf(buf, sizeof(buf), fmt, a, b);
    g_last_error = buf;
    return code;
}

#define B2_CHECK_ARG(cond)                                                                  \
    do {                                                                                    \
        if (!(cond)) return ::b2::fail(B2_EINVAL, "invalid argument: %s (line %lld)", #cond, __LINE__); \
    } while (0)

#define B2_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::b2::fail(B2_ECUDA, "CUDA error: %s (line %lld)", cudaGetErrorString(_e), __LINE__); \
    } while (0)

// Programmatic dependent launch (PDL): every kernel is launched with programmaticStreamSerialization allowed and begins
// with pdl_grid_sync(): `griddepcontrol.wait` blocks until the preceding grid in the stream has completed and its memory
// is visible (so the semantics stay plain stream order), `griddepcontrol.launch_dependents` lets the NEXT grid's CTAs be
// scheduled onto SMs as they drain instead of after the last CTA retires.  A step is ~250 dependent launches, many of
// them 5-20 us long, so launch latency and tails are a measurable share of it.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_grid_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif
extern int g_pdl;

// every kernel launch goes through this so gpu_launches is an honest count
#define B2_LAUNCH(kernel, grid_, block_, smem_, stream_, ...)                                   \
    do {                                                                                    \
        cudaLaunchConfig_t _cfg;                                                            \
        memset(&_cfg, 0, sizeof(_cfg));                                                     \
        _cfg.gridDim = dim3(grid_);                                                         \
        _cfg.blockDim = dim3(block_);                                                        \
        _cfg.dynamicSmemBytes = (smem_);                                                     \
        _cfg.stream = (cudaStream_t)(stream_);                                               \
        cudaLaunchAttribute _attr[1];                                                       \
        _attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                   \
        _attr[0].val.programmaticStreamSerializationAllowed = ::b2::g_pdl ? 1 : 0;          \
        _cfg.attrs = _attr;                                                                 \
        _cfg.numAttrs = 1;                                                                  \
        cudaError_t _e = cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__);                    \
        ::b2::g_launches.fetch_add(1, std::memory_order_relaxed);                           \
        if (_e == cudaSuccess) _e = cudaGetLastError();                                     \
        if (_e != cudaSuccess)                                                              \
            return ::b2::fail(B2_ECUDA, "launch failed: %s (line %lld)", cudaGetErrorString(_e), __LINE__); \
    } while (0)

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- dtype helpers ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// load 8 consecutive elements as floats (16B-aligned for bf16, 32B for fp32 via two float4)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 f = __bfloat1622float2(h[i]);
        v[2 * i] = f.x; v[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
}
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
    float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
    h[0] = __floats2bfloat162_rn(v[0], v[1]);
    h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = r;
}

// ---- deterministic reductions -------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// block-wide sum, result valid in thread 0 (fixed tree => bit-reproducible). `red` holds >= 32 elements.
template <typename V>
__device__ __forceinline__ V block_sum(V v, V* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    V r = 0;
    if (warp == 0) {
        r = lane < nw ? red[lane] : (V)0;
        r = warp_sum(r);
    }
    return r;
}

int num_sms();

}  // namespace b2
