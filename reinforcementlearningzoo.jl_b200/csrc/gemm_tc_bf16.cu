// tcgen05 GEMM / implicit-conv kernels, bf16 operand mode (see gemm_tc.cuh); split out so the three modes compile in parallel.
#include "gemm_tc.cuh"

namespace zoo {
zoo_status tc_launch_bf16(int kind, int BN, bool ak, bool bk, const CUtensorMap &tmA, const CUtensorMap &tmB, const TcParams &P, dim3 grid,
                        cudaStream_t st) {
    return tc_launch_mode<2, false>(kind, BN, ak, bk, tmA, tmB, P, grid, st);
}
}  // namespace zoo
