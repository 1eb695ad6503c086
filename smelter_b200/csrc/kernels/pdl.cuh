// Programmatic dependent launch for the non-conv kernels of a plan: a kernel may be scheduled while its predecessor drains and tells
// its successor to do the same; it touches no memory before griddepcontrol.wait (reads of what the predecessor wrote, and writes into
// buffers the predecessor may still read, both come after it).  Saves ~2.5 us per kernel boundary against a plain stream-ordered
// launch -- TransformerNet has 68 such boundaries (pads, instance-norm passes, adds, upsamples), MobileNetV2 17 depthwise ones.
#pragma once
#include <cuda_runtime.h>

namespace smelter {
namespace k {

__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_smem(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args... args) {
    return launch_pdl_smem(kernel, grid, block, 0, s, args...);
}

}  // namespace k
}  // namespace smelter
