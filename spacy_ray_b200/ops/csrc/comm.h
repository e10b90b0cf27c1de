// Peer-memory collectives: op registration hooks (implemented in comm_binding.cpp).
#pragma once
#include <torch/library.h>
namespace srb {
void register_comm_ops(torch::Library& m);
void register_comm_impls(torch::Library& m);
}  // namespace srb
