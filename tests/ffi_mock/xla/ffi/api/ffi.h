// TEST-ONLY stand-in for jaxlib's xla/ffi/api/ffi.h: just enough of the typed-FFI binding surface
// (Buffer / ResultBuffer / Error / Ffi::Bind().Ctx().Arg().Ret().Attr() / XLA_FFI_DEFINE_HANDLER_SYMBOL) to
// type-check vae_training_b200/csrc/xla_ffi_shim.cc where jaxlib is not installed (tests/test_abi.py).  It checks
// that every handler's C++ signature matches its binding, argument for argument, as the real header does; it
// does not implement the XLA call frame.  Never shipped, never on the product include path.
#ifndef VAE_TEST_XLA_FFI_MOCK_H_
#define VAE_TEST_XLA_FFI_MOCK_H_
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <string_view>
#include <type_traits>
#include <vector>

namespace xla { namespace ffi {
enum DataType { F32, S32, U32, U8 };
template <DataType> struct NativeOf;
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U32> { using type = uint32_t; };
template <> struct NativeOf<U8> { using type = uint8_t; };
template <DataType dt> class Buffer {
 public:
  using T = typename NativeOf<dt>::type;
  T* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  size_t size_bytes() const { return n_ * sizeof(T); }
  size_t element_count() const { return n_; }
  const std::vector<int64_t>& dimensions() const { return dims_; }
 private:
  T* data_ = nullptr; size_t n_ = 0; std::vector<int64_t> dims_;
};
template <typename B> class Result {
 public:
  B* operator->() { return &b_; }
 private:
  B b_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;
enum class ErrorCode { kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode, std::string m) : ok_(false), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool success() const { return ok_; }
 private:
  bool ok_ = true; std::string msg_;
};
template <typename T> struct PlatformStream {};
template <typename T> struct CtxArg { using type = T; };
template <typename T> struct CtxArg<PlatformStream<T>> { using type = T; };
template <typename T> struct RetArg { using type = Result<T>; };
// the binding accumulates the handler's expected parameter list as a type pack
template <typename... Ps> struct Binding {
  template <typename T> Binding<Ps..., typename CtxArg<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Ps..., T> Arg() const { return {}; }
  template <typename T> Binding<Ps..., typename RetArg<T>::type> Ret() const { return {}; }
  template <typename T> Binding<Ps..., T> Attr(const char*) const { return {}; }
  static constexpr bool Matches(Error (*)(Ps...)) { return true; }
  template <typename F> static constexpr bool Matches(F) { return false; }
};
struct Ffi { static Binding<> Bind() { return {}; } };
}}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                        \
  static_assert(decltype(binding)::Matches(&impl), #name ": handler signature does not match its binding");       \
  extern "C" void* name(void*) { return nullptr; }
#endif
