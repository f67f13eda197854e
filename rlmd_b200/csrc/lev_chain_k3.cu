// Discrete CHAIN kernels for K = 3 outcomes (one translation unit per K so the
// template instantiations compile in parallel).
#include "lev_kernels.cuh"

namespace b200 {
template <>
int chain_discrete_launch<3>(const ChainLaunch& a) { return chain_launch_all<3>(a); }
}  // namespace b200
