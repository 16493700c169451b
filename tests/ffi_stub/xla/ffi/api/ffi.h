// tests/ffi_stub/xla/ffi/api/ffi.h - COMPILE-CHECK STUB of the XLA FFI C++ API.  TEST INFRASTRUCTURE ONLY.
//
// The real header ships with jaxlib (`jax.ffi.include_dir()` -> xla/ffi/api/ffi.h) and is not available in the authoring
// container (no jax wheel, no network).  This file declares the subset of that public API that
// diffrax_b200/csrc/ffi/xla_ffi_shim.cc uses - same namespaces, class and member names, and the same binding grammar
// (Ffi::Bind().Ctx<>().Arg<>().Ret<>().Attr<>() ... XLA_FFI_DEFINE_HANDLER_SYMBOL) - so that
// tests/test_ffi_shim.py can compile and link the shim on every run and it cannot rot unnoticed:
//   * `Binding::To(fn)` static_asserts that `fn` is invocable with exactly the decoded types, in order
//     (context, then arguments, then results as Result<Buffer>, then attributes) and returns ffi::Error;
//   * the generated extern "C" symbol has the real signature `XLA_FFI_Error* (XLA_FFI_CallFrame*)`.
// It does NOT decode call frames: calling the handler through this stub returns an error object.  When jax is present the
// shim is built against the real header instead (diffrax_b200/build.py:build_ffi).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>

extern "C" {
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_CallFrame XLA_FFI_CallFrame;
typedef XLA_FFI_Error *XLA_FFI_Handler(XLA_FFI_CallFrame *);
}

namespace xla::ffi {

enum class DataType : uint8_t { INVALID = 0, PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
inline constexpr DataType PRED = DataType::PRED, S8 = DataType::S8, S16 = DataType::S16, S32 = DataType::S32,
                          S64 = DataType::S64, U8 = DataType::U8, U16 = DataType::U16, U32 = DataType::U32,
                          U64 = DataType::U64, F16 = DataType::F16, F32 = DataType::F32, F64 = DataType::F64,
                          BF16 = DataType::BF16;

namespace internal {
inline constexpr size_t kDynamicRank = static_cast<size_t>(-1);
template <DataType> struct NativeTypeOf { using type = void; };
template <> struct NativeTypeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeTypeOf<DataType::S64> { using type = int64_t; };
template <> struct NativeTypeOf<DataType::U32> { using type = uint32_t; };
template <> struct NativeTypeOf<DataType::U64> { using type = uint64_t; };
template <> struct NativeTypeOf<DataType::F32> { using type = float; };
template <> struct NativeTypeOf<DataType::F64> { using type = double; };
}  // namespace internal
template <DataType dtype> using NativeType = typename internal::NativeTypeOf<dtype>::type;

template <typename T> class Span {
 public:
  constexpr Span() = default;
  constexpr Span(T *data, size_t size) : data_(data), size_(size) {}
  constexpr T *begin() const { return data_; }
  constexpr T *end() const { return data_ + size_; }
  constexpr T *data() const { return data_; }
  constexpr size_t size() const { return size_; }
  constexpr T &operator[](size_t i) const { return data_[i]; }
  constexpr T &front() const { return data_[0]; }
  constexpr T &back() const { return data_[size_ - 1]; }
 private:
  T *data_ = nullptr;
  size_t size_ = 0;
};

enum class ErrorCode : uint8_t { kOk = 0, kCancelled, kUnknown, kInvalidArgument, kDeadlineExceeded, kNotFound,
                                 kAlreadyExists, kPermissionDenied, kResourceExhausted, kFailedPrecondition, kAborted,
                                 kOutOfRange, kUnimplemented, kInternal, kUnavailable, kDataLoss, kUnauthenticated };
class Error {
 public:
  Error() = default;
  Error(ErrorCode errc, std::string message) : errc_(errc), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
  static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
  bool success() const { return errc_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }
  ErrorCode errc() const { return errc_; }
  const std::string &message() const { return message_; }
 private:
  ErrorCode errc_ = ErrorCode::kOk;
  std::string message_;
};

template <DataType dtype, size_t rank = internal::kDynamicRank> class Buffer {
 public:
  using Dimensions = Span<const int64_t>;
  void *untyped_data() const { return data_; }
  NativeType<dtype> *typed_data() const { return reinterpret_cast<NativeType<dtype> *>(data_); }
  Dimensions dimensions() const { return Dimensions(dims_, rank_); }
  size_t element_count() const { size_t n = 1; for (size_t i = 0; i < rank_; ++i) n *= static_cast<size_t>(dims_[i]); return n; }
  size_t size_bytes() const { return element_count() * sizeof(std::conditional_t<std::is_void_v<NativeType<dtype>>, char, NativeType<dtype>>); }
  constexpr DataType element_type() const { return dtype; }
 private:
  void *data_ = nullptr;
  const int64_t *dims_ = nullptr;
  size_t rank_ = 0;
};
using AnyBuffer = Buffer<DataType::INVALID>;

template <typename T> class Result {
 public:
  Result() = default;
  T &operator*() { return value_; }
  T *operator->() { return &value_; }
 private:
  T value_;
};
template <DataType dtype, size_t rank = internal::kDynamicRank> using ResultBuffer = Result<Buffer<dtype, rank>>;

template <typename T> struct PlatformStream {};

namespace internal {
template <typename T> struct CtxDecoded { using type = T; };
template <typename T> struct CtxDecoded<PlatformStream<T>> { using type = T; };
template <typename... Ts> struct TypeList {};
template <typename L, typename T> struct Append;
template <typename... Ts, typename T> struct Append<TypeList<Ts...>, T> { using type = TypeList<Ts..., T>; };
template <typename Fn, typename L> struct Invocable;
template <typename Fn, typename... Ts> struct Invocable<Fn, TypeList<Ts...>> {
  static constexpr bool value = std::is_invocable_r_v<Error, Fn, Ts...>;
};
}  // namespace internal

class HandlerBase {
 public:
  virtual ~HandlerBase() = default;
  virtual XLA_FFI_Error *Call(XLA_FFI_CallFrame *) const { return reinterpret_cast<XLA_FFI_Error *>(const_cast<HandlerBase *>(this)); }
};
template <typename Fn> class Handler : public HandlerBase {
 public:
  explicit Handler(Fn fn) : fn_(std::move(fn)) {}
 private:
  Fn fn_;
};

// Ctx... then Arg... then Ret... then Attr...: the order the decoded values reach the implementation in.
template <typename CtxL, typename Args, typename Rets, typename Attrs> class Binding {
 public:
  template <typename T> auto Ctx() const {
    return Binding<typename internal::Append<CtxL, typename internal::CtxDecoded<T>::type>::type, Args, Rets, Attrs>();
  }
  template <typename T> auto Arg() const { return Binding<CtxL, typename internal::Append<Args, T>::type, Rets, Attrs>(); }
  template <typename T> auto Ret() const { return Binding<CtxL, Args, typename internal::Append<Rets, Result<T>>::type, Attrs>(); }
  template <typename T> auto Attr(std::string) const { return Binding<CtxL, Args, Rets, typename internal::Append<Attrs, T>::type>(); }
  template <typename Fn> auto To(Fn fn) const {
    using All = typename Concat<CtxL, Args, Rets, Attrs>::type;
    static_assert(internal::Invocable<Fn, All>::value,
                  "XLA FFI binding does not match the handler implementation: the implementation must accept, in order, the "
                  "context values, the argument buffers, the results (ffi::Result<ffi::Buffer<..>>) and the attributes, and "
                  "return ffi::Error");
    return new Handler<Fn>(std::move(fn));
  }
 private:
  template <typename... Ls> struct Concat;
  template <typename... As, typename... Bs, typename... Cs, typename... Ds>
  struct Concat<internal::TypeList<As...>, internal::TypeList<Bs...>, internal::TypeList<Cs...>, internal::TypeList<Ds...>> {
    using type = internal::TypeList<As..., Bs..., Cs..., Ds...>;
  };
};

class Ffi {
 public:
  static auto Bind() { return Binding<internal::TypeList<>, internal::TypeList<>, internal::TypeList<>, internal::TypeList<>>(); }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding, ...)                         \
  extern "C" XLA_FFI_Error *fn(XLA_FFI_CallFrame *call_frame) {                      \
    static auto *handler = (binding).To(impl);                                       \
    return handler->Call(call_frame);                                                \
  }
