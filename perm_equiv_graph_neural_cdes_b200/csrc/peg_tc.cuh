// tcgen05 (5th-gen tensor core) implementation of the matrix-free equivariant contraction.
// Declarations only; the kernels live in peg_tc.cu.
#pragma once
#include "peg_common.cuh"

namespace peg {

// operand buffers the tensor-core path needs in the caller's workspace
struct TcWs {
  float* Vt_hi;  // [B][dmax][npad]  V^T, high part (K-major B operand): tf32-rounded fp32 words, or packed bf16 (tc_fmt16)
  float* Vt_lo;  // [B][dmax][npad]  residual V - high part, same format
  float* partial;  // [B][4][n][dmax] split-K accumulators (grids far smaller than the GPU)
  int npad;
  int fmt;          // operand format of this call (PEG_FMT_*), set by make_ctx from the flags and the control
  int* vexp;        // [B][vexp_stride] block exponents of V^T (fp16x2 format)
  unsigned int* vmax;   // [2][B][vexp_stride] scratch of k_block_exponent (maxima, arrival counters), zero between launches
  int vexp_stride;  // = ceil(ldk / 128)
  int ldk;          // row pitch of V^T in elements = padded GLOBAL node count (== npad unless row-sharded)
  int col0, blk0;   // row-sharded mode: this rank's rows are columns [col0, col0 + npad) of V^T (else 0)
};

// per-layer tf32 hi/lo copies of the Linear weights (built once per API call by tc_prep_weights):
//   Wn = W diag(norm.weight)  [dout][din]  (B operand of RMSNorm -> Linear; the norm scale rides in the epilogue)
//   cvec[o] = sum_k norm.bias[k] W[o][k] + bias[o]
struct TcLinear {
  float* Wn_hi[PEG_MAX_LAYERS];
  float* Wn_lo[PEG_MAX_LAYERS];
  float* cvec[PEG_MAX_LAYERS];
  float* Wt_hi[PEG_MAX_LAYERS];   // W^T [din][dout] (B operand of the Linear backward: Nbar = Mbar W)
  float* Wt_lo[PEG_MAX_LAYERS];
  bool ready;   // carved (tensor-core flag set)
};

void tc_refresh_env();   // re-reads the PEG_TC_* environment knobs (once per API call)
void tc_carve(Bump& bp, const PegDims& d, int dmax, TcWs& w);
void tc_carve_linear(Bump& bp, const PegDims& d, const Model& m, TcLinear& w);
bool tc_linear_supported(int din, int dout);
// splits the weights of every layer (enqueue only); call once per API entry before the first evaluation
int tc_prep_weights(cudaStream_t st, const Model& m, const float* params, const TcLinear& w);
// M = rmsnorm(Z) W^T + b on tcgen05 (3xTF32), fused producer outputs as k_norm_linear (peg_kernels.cuh)
bool tc_norm_linear_full_columns(int dout);   // one CTA holds every output column of its 128 nodes (it can emit fp16x2 V^T)
// backward of Linear + RMSNorm wrt the layer input on tcgen05, epilogue as k_linear_bwd (peg_kernels.cuh)
int tc_linear_bwd(cudaStream_t st, const PegDims& d, const TcLinear& w, int layer, const float* Mbar, const float* Z,
                  const float* nw, int din, int dout, int relu_mask, float* Zbar, float* g_nw, float* g_nb, const ProducerOut& po);
bool tc_linear_bwd_supported(int din, int dout);
// Wbar += Mbar^T N, bbar += 1^T Mbar over all rows, on tcgen05 (both operands transposed in the loaders)
bool tc_weight_grad_supported(int din, int dout);
int tc_weight_grad(cudaStream_t st, const PegDims& d, const float* Mbar, const float* N, size_t rows, int din, int dout, float* gW,
                   float* gb);
int tc_norm_linear(cudaStream_t st, const PegDims& d, const TcLinear& w, int layer, const float* Z, int din, int dout,
                   const float* nw, const float* nb, float* M, float* Nout, const ProducerOut& po);
bool tc_supported(const PegDims& d, int dcols);
int tc_fmt(const PegDims& d, bool have_absmax);   // operand format of the contraction (PEG_FMT_*)
int tc_convert_v(cudaStream_t st, const PegDims& d, const TcWs& w, const ContractArgs& a);
int tc_contract(cudaStream_t st, const PegDims& d, const TcWs& w, const ContractArgs& a, bool bwd);
int tc_launches_per_contract(bool bwd);
void set_last_cuda(int err);   // records a cudaError_t for pegncde_last_cuda_error() (defined in pegncde.cu)

}  // namespace peg
