// MOCK of the part of XLA's FFI C++ API (xla/ffi/api/ffi.h, jax >= 0.4.31) that jax_ffi/pegncde_ffi.cc uses -- TEST INFRASTRUCTURE ONLY.
// The real header is not available in this image (no jax / jaxlib, no network).  This mock keeps the unbuilt shim honest: it
// type-checks every handler against its binding (argument order, buffer element types, attribute types) and against the
// prototypes of include/pegncde.h, so a change of the C-ABI that the shim does not follow fails `pytest -m "not gpu"`.
// It says nothing about the real XLA runtime behaviour.
#ifndef MOCK_XLA_FFI_API_FFI_H_
#define MOCK_XLA_FFI_API_FFI_H_
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>

namespace xla {
namespace ffi {

enum DataType { F32, U8 };
template <DataType> struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<U8> { using type = uint8_t; };

template <DataType dt>
struct Buffer {
  using T = typename NativeType<dt>::type;
  T* typed_data() const { return nullptr; }
  size_t element_count() const { return 0; }
};
template <typename B>
struct Result {
  B* operator->() const { return nullptr; }
  B& operator*() const { static B b; return b; }
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

template <typename T>
struct Span {
  const T* begin() const { return nullptr; }
  const T* end() const { return nullptr; }
  size_t size() const { return 0; }
};

enum class ErrorCode { kInternal, kInvalidArgument };
struct Error {
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

template <typename T> struct PlatformStream {};
template <typename C> struct CtxDecode;
template <typename T> struct CtxDecode<PlatformStream<T>> { using type = T; };

template <typename F> struct HandlerArgs;
template <typename... As> struct HandlerArgs<Error (*)(As...)> { using type = std::tuple<As...>; };

template <typename... Ts>
struct Binding {
  template <typename C> Binding<Ts..., typename CtxDecode<C>::type> Ctx() const { return {}; }
  template <typename A> Binding<Ts..., A> Arg() const { return {}; }
  template <typename R> Binding<Ts..., Result<R>> Ret() const { return {}; }
  template <typename A> Binding<Ts..., A> Attr(const char*) const { return {}; }
  // exact match, no implicit conversions: the handler's parameter list IS the decoded binding
  template <typename F> static constexpr bool matches() { return std::is_same<typename HandlerArgs<F>::type, std::tuple<Ts...>>::value; }
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                        \
  static_assert(decltype(binding)::template matches<decltype(&impl)>(), "handler signature does not match its binding"); \
  extern "C" const void* const name = reinterpret_cast<const void*>(&impl)

#endif  // MOCK_XLA_FFI_API_FFI_H_
