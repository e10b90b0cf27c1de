// Kernel launch helper: every kernel of the training chain goes through launch_k(), which sets the
// programmatic-stream-serialization attribute (programmatic dependent launch).  Each of those
// kernels executes `griddepcontrol.launch_dependents` first thing and `griddepcontrol.wait` before
// its first global-memory access, so the NEXT kernel of the stream is scheduled (its CTAs become
// resident and run their memory-free prologue: barrier init, TMEM allocation, descriptor prefetch)
// while the current one is still running, and only its first load/store waits for the
// predecessor to complete and flush.  Semantics are exactly stream order; what is removed is the
// launch latency + prologue of ~60 dependent kernels per step.  SRB_PDL=0 (or set_pdl(false),
// used around side-stream launches) turns the attribute off; the device instructions are no-ops
// for a kernel launched without it.
#pragma once
#include <cuda_runtime.h>

namespace srb {

extern int g_pdl;   // defined in elementwise_kernels.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace srb
