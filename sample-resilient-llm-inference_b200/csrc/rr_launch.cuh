// rr_launch.cuh — programmatic dependent launch (PDL) plumbing.
//
// Every kernel of the library calls griddep_launch() at its top (lets the next kernel in the stream
// start its prologue / constant-data prefetch early) and griddep_wait() before it touches anything
// the preceding kernel produced.  Launched through launch_pdl() the stream edge becomes a programmatic
// dependency (also inside CUDA-graph capture); launched normally both instructions are no-ops.
#pragma once
#include <cuda_runtime.h>
#include <utility>

namespace rr {

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_use_pdl;   // rr_api.cu; env RR_NO_PDL=1 or rr_set_pdl(0) turns it off

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

}  // namespace rr
