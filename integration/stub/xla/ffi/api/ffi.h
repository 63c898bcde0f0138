// MINIMAL STAND-IN for jaxlib's xla/ffi/api/ffi.h — NOT the real header, test infrastructure only.
//
// jaxlib is not installable in this image, so integration/zb_xla_ffi.cc cannot be built against the real XLA FFI API here.
// This stub declares just enough of the API surface that file uses (Buffer / ResultBuffer, Error, the Bind() builder and
// XLA_FFI_DEFINE_HANDLER_SYMBOL) for `g++ -fsyntax-only` to type-check it ONCE (tests/test_integration_cpu.py), including the one
// property that matters for a binding: the handler function must be callable with exactly the context / attribute / operand /
// result types its Bind() chain declares, in that order (static_assert in the macro).  Semantics of the real header that the
// stub mirrors, as recalled [R] from jaxlib 0.6.x: Bind().Ctx<PlatformStream<T>>() passes a T; Attr<T>(name) passes a T;
// Arg<Buffer<DT>>() passes a Buffer<DT>; Ret<Buffer<DT>>() passes a Result<Buffer<DT>> (= ResultBuffer<DT>, used with ->).
#ifndef ZB_STUB_XLA_FFI_API_FFI_H_
#define ZB_STUB_XLA_FFI_API_FFI_H_
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum DataType { F32, S32, U8 };
template <DataType> struct NativeOf;
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U8> { using type = uint8_t; };

struct Span {
  const int64_t* p = nullptr;
  size_t n = 0;
  int64_t operator[](size_t i) const { return p[i]; }
  size_t size() const { return n; }
};

template <DataType DT>
class Buffer {
 public:
  using T = typename NativeOf<DT>::type;
  T* typed_data() const { return data_; }
  Span dimensions() const { return dims_; }
  size_t size_bytes() const { return bytes_; }
  size_t element_count() const { return bytes_ / sizeof(T); }

 private:
  T* data_ = nullptr;
  Span dims_;
  size_t bytes_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType DT> using ResultBuffer = Result<Buffer<DT>>;

enum class ErrorCode { kOk, kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T> struct PlatformStream { using type = T; };

template <typename... Ts>
struct Binding {
  template <typename C> Binding<Ts..., typename C::type> Ctx() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename B> Binding<Ts..., B> Arg() const { return {}; }
  template <typename B> Binding<Ts..., Result<B>> Ret() const { return {}; }
  template <typename F> static constexpr bool Accepts() { return std::is_invocable_r_v<Error, F, Ts...>; }
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                                                              \
  static_assert(decltype(binding)::template Accepts<decltype(&impl)>(),                                                \
                #impl " is not callable with the context / attribute / operand / result types of its Bind() chain");  \
  extern "C" XLA_FFI_Error* sym(XLA_FFI_CallFrame*) { return nullptr; }
#endif
