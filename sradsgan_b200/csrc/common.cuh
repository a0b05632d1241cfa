// Shared helpers for the sradsgan_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sradsgan_b200.h"

namespace sr {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);   // cudaGetLastError -> SR_OK / SR_ERR_CUDA
// Tuning / experiment options set EXPLICITLY by the caller through sr_set_option() (api.cu); the library never reads the
// environment.  Unknown names return `dflt`.
int option(const char* name, int dflt);
// Side streams + events handed in by the caller through sr_set_aux_streams() (the library creates none): used to run
// independent launches of ONE call next to each other (fork / join with events: stream-ordered for the caller, capturable).
struct AuxStreams { cudaStream_t stream[3]; cudaEvent_t fork; cudaEvent_t join[3]; int n; };
const AuxStreams& aux_streams();

#define SR_REQUIRE(cond, ...)                \
    do {                                     \
        if (!(cond)) {                       \
            sr::set_error(__VA_ARGS__);      \
            return SR_ERR_ARG;               \
        }                                    \
    } while (0)

__host__ __device__ inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements -> float[4] (requires 4-element alignment of p)
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&o)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&o)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
    uint2 v = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&v.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&v.y);
    o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}

template <typename T> __device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <> __device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&lo); v.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = v;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    switch (act) {
        case SR_ACT_LRELU: return v > 0.f ? v : v * slope;
        case SR_ACT_RELU: return v > 0.f ? v : 0.f;
        case SR_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
        default: return v;
    }
}

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl(..., pdl = true) may be scheduled while its predecessor
// in the stream is still running; it must call pdl_wait() before it touches global memory (returns when the predecessor has
// completed and flushed).  A predecessor that calls pdl_trigger() lets the dependent's blocks take over its SMs as its own
// blocks exit — launch latency and the dependent's prologue (barrier init, TMEM allocation, descriptor prefetch) then overlap
// the predecessor's tail.  Both are no-ops for launches without the attribute.  Captured by CUDA graphs as programmatic edges.
// Option "SR_PDL" (default 0): a captured chain of 24 dependent RAB convolutions replays at 22.75 us per launch without and
// 22.83 us with the attribute (scripts/graph_gap_probe.py, profiles/r02_graph_gap_probe.txt) — graph replay already hides the
// launch latency, and a dependent's blocks cannot become resident before the predecessor's blocks (227 KB of shared memory
// each) have left.  The whole step: 22.98 ms with vs 23.06 ms without (within run-to-run noise; the kernel, fused-op, model and
// full-size parity tests all pass with SR_PDL=1, gpurun r2c23) — so the attribute stays off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (pdl) { at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[n].val.programmaticStreamSerializationAllowed = 1; ++n; }
    if (cluster_x > 1) { at[n].id = cudaLaunchAttributeClusterDimension; at[n].val.clusterDim.x = cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1; ++n; }
    cfg.attrs = at; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
    return launch_ex(kernel, grid, block, smem, st, pdl, 1, static_cast<Args&&>(args)...);
}
// thread-block clusters of `cluster_x` CTAs of this kernel that can be resident at once (0 on error)
template <typename... KArgs>
inline int max_active_clusters(void (*kernel)(KArgs...), int block, size_t smem, int cluster_x, int grid) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace sr
