// Programmatic dependent launch (griddepcontrol, sm_90+): the kernels of a decode step are short and strictly
// serial -- ~230 per step at batch 20 -- so the launch latency between them is a visible share of the step.  A kernel
// launched with the programmatic-stream-serialization attribute may be scheduled while its predecessor is still
// running; it must execute pdl_wait() before it touches anything the predecessor produces (or still reads), which
// returns once the predecessor grid has completed and its memory is visible.  pdl_launch_dependents() lets the NEXT
// kernel of the stream be scheduled early in turn.  Both are no-ops for a kernel launched the ordinary way.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

#include "common.cuh"

namespace sealb200 {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// first statement of a PDL-aware kernel whose prologue has nothing worth overlapping
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_launch_dependents(); }

inline bool pdl_enabled() {
    static const bool on = [] { const char* e = std::getenv("SEALB200_PDL"); return !e || std::atoi(e) != 0; }();
    return on;
}

// kern<<<grid, block, smem, s>>>(args...) with the programmatic-serialization attribute (SEALB200_PDL=0: without).
// Only for kernels that call pdl_wait() / pdl_enter().
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...));
}

}  // namespace sealb200
