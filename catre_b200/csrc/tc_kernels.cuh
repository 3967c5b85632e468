// placeholder: tensor-core path (filled in below)
#pragma once
#include <cuda_runtime.h>
#include <map>
#include <string>
#include <vector>
struct catre_engine;
namespace catre {
struct TcWeights {};
struct TcWorkspace {};
inline int tc_unsupported() { return -6; }
inline void tc_debug_buffers(TcWorkspace&, std::map<std::string, const void*>&) {}
inline int tc_workspace_alloc(TcWorkspace&, size_t, std::vector<void*>&, size_t*, catre_engine*) { return 0; }
inline int tc_pack_weights(TcWeights&, const std::map<std::string, std::vector<float>>&, const std::vector<float>&, bool,
                           std::vector<void*>&, catre_engine*) { return tc_unsupported(); }
inline int tc_tnet_trunk(TcWeights&, TcWorkspace&, cudaStream_t, const float*, bool, long long, int, int*, catre_engine*) { return tc_unsupported(); }
inline int tc_trunk(TcWeights&, TcWorkspace&, cudaStream_t, const float*, long long, int, int*, catre_engine*) { return tc_unsupported(); }
inline int tc_rot_layers(TcWeights&, TcWorkspace&, cudaStream_t, const float*, const float*, long long, int, float*, float*,
                         float*, float*, float*, const float*, const float*, int, catre_engine*) { return tc_unsupported(); }
}  // namespace catre
