// tcgen05 (5th-gen tensor core) implementation of the matrix-free equivariant contraction.
// Declarations only; the kernels live in peg_tc.cu.
#pragma once
#include "peg_common.cuh"

namespace peg {

// operand buffers the tensor-core path needs in the caller's workspace
struct TcWs {
  float* Vt_hi;  // [B][dmax][npad]  tf32-rounded V^T (K-major B operand)
  float* Vt_lo;  // [B][dmax][npad]  residual V - tf32(V)
  int npad;
};

void tc_carve(Bump& bp, const PegDims& d, int dmax, TcWs& w);
bool tc_supported(const PegDims& d, int dcols);
int tc_contract(cudaStream_t st, const PegDims& d, const TcWs& w, const ContractArgs& a, bool bwd);
int tc_launches_per_contract(bool bwd);
void set_last_cuda(int err);   // records a cudaError_t for pegncde_last_cuda_error() (defined in pegncde.cu)

}  // namespace peg
