// tcgen05 GEMM family: op registration hooks (implemented in gemm_binding.cpp).
#pragma once
#include <torch/library.h>
namespace srb {
void register_gemm_ops(torch::Library& m);
void register_gemm_impls(torch::Library& m);
}  // namespace srb
