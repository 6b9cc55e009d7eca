// cubecl_b200.hpp -- header-only C++ host layer above the C ABI (include/cubecl_b200.h).
//
// The reference's host side is Rust; this image has no Rust toolchain, so this is the compiled-language mirror of the
// reference's launch surface for the dense-LA path (the Python mirror in cubecl_b200/ serves the tests and the bench):
//
//   cubecl::ComputeClient   crates/cubecl-runtime/src/client.rs:44-48  (create_from_slice:452, empty:654, read_one:256,
//                           sync:1013, memory_usage:1048); one client per device (R::client(device))
//   cubecl::Handle          crates/cubecl-runtime/src/server/handle.rs:10-21  (ref-counted pool slice)
//   cubecl::TensorHandle    crates/cubecl-std/src/tensor/handle.rs:13-150     (handle + shape + strides in elements + dtype)
//   cubecl::matmul::launch  cubek matmul::launch; shape rule crates/cubecl-zspace/src/shape.rs:489-517
//   cubecl::reduce::launch  cubek reduce::launch; semantics examples/sum_things/src/lib.rs:6-33
//
// Error behaviour mirrors the reference: launch never throws; failures are queued on the client and surface as
// ServerError at the next sync()/read_one() (crates/cubecl-cuda/src/compute/server.rs:269-284,981-1002).
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "cubecl_b200.h"

namespace cubecl {

struct ServerError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct MatmulShapeError : std::invalid_argument {
  using std::invalid_argument::invalid_argument;
};

enum class DType : int { F32 = B200_F32, F16 = B200_F16, BF16 = B200_BF16, U32 = B200_U32, U8 = B200_U8, I8 = B200_I8,
                         F8E4M3 = B200_F8E4M3, F8E5M2 = B200_F8E5M2,
                         F4E2M1X2 = B200_F4E2M1X2,  // one element = one byte holding two e2m1 values
                         UE8M0 = B200_UE8M0 };
inline size_t dtype_size(DType d) {
  switch (d) {
    case DType::F16: case DType::BF16: return 2;
    case DType::F32: case DType::U32: return 4;
    default: return 1;
  }
}

using Shape = std::vector<uint64_t>;
using Strides = std::vector<uint64_t>;

inline Strides contiguous_strides(const Shape& shape) {
  Strides s(shape.size());
  uint64_t acc = 1;
  for (size_t i = shape.size(); i-- > 0;) { s[i] = acc; acc *= shape[i]; }
  return s;
}

/// calculate_matmul_output (crates/cubecl-zspace/src/shape.rs:489-517)
inline Shape calculate_matmul_output(const Shape& lhs, const Shape& rhs) {
  const size_t rank = lhs.size();
  if (rank != rhs.size()) throw MatmulShapeError("RankMismatch");
  if (rank < 2) throw MatmulShapeError("matmul needs rank >= 2");
  if (lhs[rank - 1] != rhs[rank - 2]) throw MatmulShapeError("IncompatibleShapes");
  Shape out;
  for (size_t i = 0; i + 2 < rank; ++i) {
    if (lhs[i] == rhs[i] || rhs[i] == 1) out.push_back(lhs[i]);
    else if (lhs[i] == 1) out.push_back(rhs[i]);
    else throw MatmulShapeError("IncompatibleDims");
  }
  out.push_back(lhs[rank - 2]);
  out.push_back(rhs[rank - 1]);
  return out;
}

class ComputeClient;

/// Pooled device buffer; the last copy returns it to the pool.
class Handle {
 public:
  Handle() = default;
  uint64_t ptr() const { return rec_ ? rec_->ptr : 0; }
  size_t size() const { return rec_ ? rec_->size : 0; }

 private:
  friend class ComputeClient;
  struct Rec {
    b200_ctx* ctx;
    uint64_t ptr;
    size_t size;
    Rec(b200_ctx* c, uint64_t p, size_t s) : ctx(c), ptr(p), size(s) {}
    Rec(const Rec&) = delete;
    Rec& operator=(const Rec&) = delete;
    ~Rec() { if (ctx && ptr) b200_free(ctx, ptr); }
  };
  std::shared_ptr<Rec> rec_;
};

class ComputeClient {
 public:
  explicit ComputeClient(int device = 0) : device_(device) {
    if (b200_init(device, &ctx_) != B200_OK) throw ServerError(std::string("b200_init: ") + b200_last_error());
  }
  ~ComputeClient() { if (ctx_) b200_destroy(ctx_); }
  ComputeClient(const ComputeClient&) = delete;
  ComputeClient& operator=(const ComputeClient&) = delete;

  b200_ctx* raw() const { return ctx_; }
  int device() const { return device_; }

  Handle empty(size_t size) {
    uint64_t p = 0;
    check(b200_alloc(ctx_, size, &p));
    Handle h;
    h.rec_ = std::shared_ptr<Handle::Rec>(new Handle::Rec{ctx_, p, size});  // in place: a temporary Rec would free the block
    return h;
  }
  Handle create_from_slice(const void* data, size_t bytes) {
    Handle h = empty(bytes);
    if (bytes) { check(b200_write(ctx_, nullptr, h.ptr(), data, bytes)); check(b200_sync(ctx_, nullptr)); }
    return h;
  }
  /// Blocking read of the whole buffer; surfaces deferred errors first (Result<Bytes, ServerError>).
  std::vector<uint8_t> read_one(const Handle& h) {
    std::vector<uint8_t> out(h.size());
    if (h.size()) check(b200_read(ctx_, nullptr, out.data(), h.ptr(), h.size()));
    sync();
    return out;
  }
  void sync() {
    if (b200_sync(ctx_, nullptr) != B200_OK) errors_.push_back(b200_last_error());
    flush();
  }
  void flush() {
    if (errors_.empty()) return;
    std::string msg = "ServerUnhealthy:";
    for (auto& e : errors_) msg += " " + e + ";";
    errors_.clear();
    throw ServerError(msg);
  }
  void defer(std::string e) { errors_.push_back(std::move(e)); }
  std::pair<uint64_t, uint64_t> memory_usage() {
    uint64_t a = 0, b = 0;
    check(b200_memory_usage(ctx_, &a, &b));
    return {a, b};
  }
  void set_option(const char* key, const char* value) { check(b200_set_option(ctx_, key, value)); }

 private:
  void check(int rc) {
    if (rc != B200_OK) throw ServerError(b200_last_error());
  }
  b200_ctx* ctx_ = nullptr;
  int device_;
  std::vector<std::string> errors_;
};

struct TensorHandle {
  Handle handle;
  Shape shape;
  Strides strides;  // elements
  DType dtype;

  static TensorHandle new_contiguous(Shape shape, Handle handle, DType dtype) {
    Strides st = contiguous_strides(shape);
    return TensorHandle{std::move(handle), std::move(shape), std::move(st), dtype};
  }
  static TensorHandle empty(ComputeClient& client, Shape shape, DType dtype) {
    uint64_t n = 1;
    for (auto s : shape) n *= s;
    return new_contiguous(std::move(shape), client.empty(n ? n * dtype_size(dtype) : 1), dtype);
  }
  /// Swap the last two dims without moving data (MatrixBatchLayout::MildlyPermuted{transposed}).
  TensorHandle transposed() const {
    TensorHandle t = *this;
    const size_t r = shape.size();
    std::swap(t.shape[r - 1], t.shape[r - 2]);
    std::swap(t.strides[r - 1], t.strides[r - 2]);
    return t;
  }
};

namespace matmul {
/// out = lhs @ rhs, f32 accumulation, batch dims broadcast.  Errors are deferred to client.sync().
inline void launch(ComputeClient& client, const TensorHandle& lhs, const TensorHandle& rhs, const TensorHandle& out) {
  const int rank = static_cast<int>(lhs.shape.size());
  if (rhs.shape.size() != lhs.shape.size() || out.shape.size() != lhs.shape.size() || lhs.dtype != rhs.dtype) {
    client.defer("InvalidArgument: matmul operands must have equal rank and dtype");
    return;
  }
  const int rc = b200_matmul(client.raw(), nullptr, static_cast<b200_dtype>(lhs.dtype), static_cast<b200_dtype>(out.dtype),
                             lhs.handle.ptr(), rhs.handle.ptr(), out.handle.ptr(), rank, lhs.shape.data(), lhs.strides.data(),
                             rhs.shape.data(), rhs.strides.data(), out.shape.data(), out.strides.data());
  if (rc != B200_OK) client.defer(b200_last_error());
}
/// Block-scaled (MX) matmul, the GEMM-level form of MmaDefinition::new_scaled / execute_scaled
/// (crates/cubecl-core/src/frontend/cmma.rs:438-460, 798-840): lhs [.., M, K] and rhs [.., N, K] K-contiguous (fp8, or both
/// F4E2M1X2 with K/2 bytes per row), scales [.., rows, K/scale_block] (UE8M0 for block 32; e4m3 bytes for NVFP4, block 16).
inline void launch_scaled(ComputeClient& client, const TensorHandle& lhs, const TensorHandle& rhs, const TensorHandle& lhs_scales,
                          const TensorHandle& rhs_scales, const TensorHandle& out, int scale_block = 32, bool scales_packed = false) {
  const size_t r = lhs.shape.size();
  if (r < 2 || rhs.shape.size() != r || out.shape.size() != r || lhs.shape[r - 1] != rhs.shape[r - 1]) {
    client.defer("InvalidArgument: matmul_scaled needs lhs [..,M,K] and rhs [..,N,K] of equal rank and K");
    return;
  }
  uint64_t batch = 1;
  for (size_t i = 0; i + 2 < r; ++i) batch *= lhs.shape[i];
  const uint64_t k = lhs.shape[r - 1] * (lhs.dtype == DType::F4E2M1X2 ? 2 : 1);
  const int rc = b200_matmul_scaled(client.raw(), nullptr, static_cast<b200_dtype>(lhs.dtype), static_cast<b200_dtype>(rhs.dtype),
                                    static_cast<b200_dtype>(out.dtype), lhs.handle.ptr(), rhs.handle.ptr(), lhs_scales.handle.ptr(),
                                    rhs_scales.handle.ptr(), out.handle.ptr(), batch, lhs.shape[r - 2], rhs.shape[r - 2], k, scale_block,
                                    scales_packed ? 1 : 0);
  if (rc != B200_OK) client.defer(b200_last_error());
}
}  // namespace matmul

namespace reduce {
enum class Op : int { Sum = B200_REDUCE_SUM, Prod = B200_REDUCE_PROD, Max = B200_REDUCE_MAX, Min = B200_REDUCE_MIN,
                      ArgMax = B200_REDUCE_ARGMAX, ArgMin = B200_REDUCE_ARGMIN, Mean = B200_REDUCE_MEAN };
/// Reduce `axis` of a contiguous input (-1 = every element); output f32 (u32 indices for arg ops). Errors deferred.
inline void launch(ComputeClient& client, const TensorHandle& input, const TensorHandle& output, int axis, Op op) {
  const int rc = b200_reduce(client.raw(), nullptr, static_cast<b200_reduce_op>(op), static_cast<b200_dtype>(input.dtype),
                             input.handle.ptr(), output.handle.ptr(), static_cast<int>(input.shape.size()), input.shape.data(), axis);
  if (rc != B200_OK) client.defer(b200_last_error());
}
}  // namespace reduce

}  // namespace cubecl
