// tcgen05 contraction -- placeholder translation unit (kernels land here next).
#include "peg_tc.cuh"

namespace peg {

void tc_carve(Bump& bp, const PegDims& d, int dmax, TcWs& w) {
  (void)bp; (void)d; (void)dmax;
  w.Vt_hi = w.Vt_lo = nullptr;
  w.npad = 0;
}
bool tc_supported(const PegDims& d, int dcols) { (void)d; (void)dcols; return false; }
int tc_contract(cudaStream_t, const PegDims&, const TcWs&, const ContractArgs&, bool) { return PEG_ERR_UNSUPPORTED; }
int tc_launches_per_contract(bool) { return 0; }

}  // namespace peg
