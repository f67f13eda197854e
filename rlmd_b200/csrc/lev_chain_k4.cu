// Discrete CHAIN kernels for K = 4 outcomes (one translation unit per K so the
// template instantiations compile in parallel).
#include "lev_kernels.cuh"

namespace b200 {
template <>
int chain_discrete_launch<4>(const ChainLaunch& a) { return chain_launch_all<4>(a); }
}  // namespace b200
