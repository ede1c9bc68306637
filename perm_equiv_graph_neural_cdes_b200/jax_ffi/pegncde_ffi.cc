// XLA FFI shim over the C-ABI (include/pegncde.h): what a JAX host registers with jax.ffi so that
// `fused_diffeqsolve` is an XLA custom call on the CUDA platform.
//
// STATUS: NOT BUILT OR RUN IN THIS REPOSITORY'S IMAGE.  jax / jaxlib (and therefore xla/ffi/api/ffi.h) are not
// installed and cannot be installed here (no network); this file documents the binding exactly as a maintainer
// with a JAX environment would compile it:
//   g++ -O2 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())") -I../../include -I$CUDA_HOME/include
//       pegncde_ffi.cc -L.. -lpegncde -L$CUDA_HOME/lib64 -lcudart -o libpegncde_ffi.so        (one command line)
// What IS checked here (tests/test_host.py): the file compiles against a MOCK of the FFI API (tests/mock_xla_ffi) that
// type-checks every handler against its binding and against the prototypes of pegncde.h.
// Everything it calls is exercised on B200 through the same C-ABI from the torch/ctypes host (tests/).
#include <cstdint>

#include <cuda_runtime_api.h>   // cudaStream_t, cudaMemsetAsync

#include "pegncde.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static PegDims dims_of(int32_t B, int32_t n, int32_t h, int32_t e, int32_t L, int32_t T, int32_t flags) {
  PegDims d{B, n, (n + 31) / 32 * 32, h, e, L, T, flags};
  return d;
}

// Control-path pre-pass, once per batch (src/engine/trainer_pgt.py:201-207 hands the model the reference-layout arrays):
// operands: d, c, b, a  each [B, T-1, n, n, 2]  ->  results: adj_coef, adj_rowsum, adj_diag, adj_total, tch_coef (PegControl fields)
static ffi::Error PackAdjImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> cd, ffi::Buffer<ffi::F32> cc, ffi::Buffer<ffi::F32> cb,
                              ffi::Buffer<ffi::F32> ca, ffi::ResultBuffer<ffi::F32> adj_coef,
                              ffi::ResultBuffer<ffi::F32> adj_rowsum, ffi::ResultBuffer<ffi::F32> adj_diag,
                              ffi::ResultBuffer<ffi::F32> adj_total, ffi::ResultBuffer<ffi::F32> tch_coef, int32_t B, int32_t n,
                              int32_t h, int32_t e, int32_t L, int32_t T, int32_t flags) {
  PegDims d = dims_of(B, n, h, e, L, T, flags);
  // (the pack kernel writes the zero padding of the ldn x ldn tiles itself: uninitialised result buffers are fine)
  const int rc = pegncde_pack_adj(stream, &d, cd.typed_data(), cc.typed_data(), cb.typed_data(), ca.typed_data(),
                                  adj_coef->typed_data(), adj_rowsum->typed_data(), adj_diag->typed_data(),
                                  adj_total->typed_data(), tch_coef->typed_data());
  if (rc != PEG_OK) return ffi::Error(ffi::ErrorCode::kInternal, pegncde_strerror(rc));
  return ffi::Error::Success();
}

// operands: d, c, b, a  each [B, T-1, n, e, 2]  ->  result: x_coef [B, T-1, 3, n, 2e]
static ffi::Error PackXImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> cd, ffi::Buffer<ffi::F32> cc, ffi::Buffer<ffi::F32> cb,
                            ffi::Buffer<ffi::F32> ca, ffi::ResultBuffer<ffi::F32> x_coef, int32_t B, int32_t n, int32_t h,
                            int32_t e, int32_t L, int32_t T, int32_t flags) {
  PegDims d = dims_of(B, n, h, e, L, T, flags);
  const int rc = pegncde_pack_x(stream, &d, cd.typed_data(), cc.typed_data(), cb.typed_data(), ca.typed_data(), x_coef->typed_data());
  if (rc != PEG_OK) return ffi::Error(ffi::ErrorCode::kInternal, pegncde_strerror(rc));
  return ffi::Error::Success();
}

// operands: params, ts, adj_coef, adj_rowsum, adj_diag, adj_total, tch_coef, x_coef, step_ts(host attr), y0
// results:  y_ckpt [S+1,B,n,h], stage_store, workspace (XLA-allocated scratch results)
static ffi::Error SolveFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> ts,
                               ffi::Buffer<ffi::F32> adj_coef, ffi::Buffer<ffi::F32> adj_rowsum,
                               ffi::Buffer<ffi::F32> adj_diag, ffi::Buffer<ffi::F32> adj_total,
                               ffi::Buffer<ffi::F32> tch_coef, ffi::Buffer<ffi::F32> x_coef, ffi::Buffer<ffi::F32> y0,
                               ffi::ResultBuffer<ffi::F32> y_ckpt, ffi::ResultBuffer<ffi::F32> stage_store,
                               ffi::ResultBuffer<ffi::U8> workspace, ffi::Span<const float> step_ts, int32_t B, int32_t n,
                               int32_t h, int32_t e, int32_t L, int32_t T, int32_t flags) {
  PegDims d = dims_of(B, n, h, e, L, T, flags);
  PegControl c{ts.typed_data(),       adj_coef.typed_data(), adj_rowsum.typed_data(), adj_diag.typed_data(),
               adj_total.typed_data(), tch_coef.typed_data(), e > 0 ? x_coef.typed_data() : nullptr};
  const int32_t steps = (int32_t)step_ts.size() - 1;
  const int rc = pegncde_solve_fwd(stream, &d, &c, params.typed_data(), step_ts.begin(), steps, y0.typed_data(), nullptr,
                                   y_ckpt->typed_data(), stage_store->element_count() ? stage_store->typed_data() : nullptr,
                                   workspace->typed_data(), workspace->element_count());
  if (rc != PEG_OK) return ffi::Error(ffi::ErrorCode::kInternal, pegncde_strerror(rc));
  return ffi::Error::Success();
}

static ffi::Error SolveBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> ts,
                               ffi::Buffer<ffi::F32> adj_coef, ffi::Buffer<ffi::F32> adj_rowsum,
                               ffi::Buffer<ffi::F32> adj_diag, ffi::Buffer<ffi::F32> adj_total,
                               ffi::Buffer<ffi::F32> tch_coef, ffi::Buffer<ffi::F32> x_coef, ffi::Buffer<ffi::F32> y_ckpt,
                               ffi::Buffer<ffi::F32> stage_store, ffi::Buffer<ffi::F32> g_ckpt,
                               ffi::ResultBuffer<ffi::F32> g_y0, ffi::ResultBuffer<ffi::F32> g_params,
                               ffi::ResultBuffer<ffi::U8> workspace, ffi::Span<const float> step_ts, int32_t B, int32_t n,
                               int32_t h, int32_t e, int32_t L, int32_t T, int32_t flags) {
  PegDims d = dims_of(B, n, h, e, L, T, flags);
  PegControl c{ts.typed_data(),       adj_coef.typed_data(), adj_rowsum.typed_data(), adj_diag.typed_data(),
               adj_total.typed_data(), tch_coef.typed_data(), e > 0 ? x_coef.typed_data() : nullptr};
  const int32_t steps = (int32_t)step_ts.size() - 1;
  cudaMemsetAsync(g_params->typed_data(), 0, g_params->element_count() * sizeof(float), stream);
  // the cotangent of the saved trajectory arrives as g_ckpt [S+1,B,n,h] (its last slab is the cotangent of y(T))
  const int rc = pegncde_solve_bwd(stream, &d, &c, params.typed_data(), step_ts.begin(), steps, y_ckpt.typed_data(),
                                   stage_store.element_count() ? stage_store.typed_data() : nullptr, nullptr,
                                   g_ckpt.typed_data(), /*g_stage=*/nullptr, g_y0->typed_data(), g_params->typed_data(), /*g_xcoef=*/nullptr,
                                   workspace->typed_data(), workspace->element_count());
  if (rc != PEG_OK) return ffi::Error(ffi::ErrorCode::kInternal, pegncde_strerror(rc));
  return ffi::Error::Success();
}

#define PEG_BIND_COMMON()                                                                                         \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define PEG_DIM_ATTRS()                                                                                           \
  Attr<int32_t>("B").Attr<int32_t>("n").Attr<int32_t>("h").Attr<int32_t>("e").Attr<int32_t>("L").Attr<int32_t>("T").Attr<int32_t>("flags")
XLA_FFI_DEFINE_HANDLER_SYMBOL(PegPackAdj, PackAdjImpl,
                              PEG_BIND_COMMON()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .PEG_DIM_ATTRS());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PegPackX, PackXImpl,
                              PEG_BIND_COMMON()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .PEG_DIM_ATTRS());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PegSolveFwd, SolveFwdImpl,
                              PEG_BIND_COMMON()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<ffi::Span<const float>>("step_ts")
                                  .Attr<int32_t>("B").Attr<int32_t>("n").Attr<int32_t>("h").Attr<int32_t>("e")
                                  .Attr<int32_t>("L").Attr<int32_t>("T").Attr<int32_t>("flags"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(PegSolveBwd, SolveBwdImpl,
                              PEG_BIND_COMMON()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<ffi::Span<const float>>("step_ts")
                                  .Attr<int32_t>("B").Attr<int32_t>("n").Attr<int32_t>("h").Attr<int32_t>("e")
                                  .Attr<int32_t>("L").Attr<int32_t>("T").Attr<int32_t>("flags"));
