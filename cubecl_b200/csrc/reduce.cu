// Hand-written sm_100a reductions: sum / mean / prod / max / min / argmax / argmin over all elements, over the
// innermost axis ("rows") or over an outer/middle axis ("columns").  HBM-bound: coalesced 128-bit (optionally 256-bit)
// streaming loads, many independent accumulators per thread, __shfl_down warp stage, smem block stage, and a
// last-block-done grid stage inside the same launch (no second kernel, no host sync).
//
// Replaces: the (out-of-tree, cubek) `reduce::launch` kernel bodies and the in-tree reduction-shaped kernels
//   examples/sum_things/src/lib.rs:6-33            (sum_basic / sum_subgroup -> plane_sum)
//   crates/cubecl-std/src/throughput/runners/memory_read.rs:68-154   (vec4 streaming read-accumulate)
//   cubecl-book/src/getting-started/src/bin/v5-gpu.rs:50-57          (row-sum, one unit per row)
// plane_sum in the reference is an xor butterfly (crates/cubecl-cpp/src/shared/plane.rs:61-70); a shfl_down tree
// produces the same value in lane 0 for commutative ops and needs no broadcast.
//
// Arg-reductions: ties -> lowest index; NaN compares as the extreme value (first NaN wins), i.e. numpy semantics.
// Compiled to a cubin: nvcc -cubin -gencode arch=compute_100a,code=sm_100a
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>

struct ReduceParams {
  uint64_t in;       // input, contiguous [outer, len, inner]
  uint64_t out;      // output [outer, inner] (f32 values, or u32 indices for arg ops)
  uint64_t ws;       // workspace: partial values f32[grid] | partial indices u64[grid] | u32 ticket
  uint64_t outer, len, inner;
  float scale;       // applied to the final value (mean = 1/len, sum = 1)
  uint32_t pad;
};

// Cross-GPU exchange fused into the grid stage (one kernel = local reduce + all-reduce of the scalar over NVLink peer
// memory).  Every rank owns a mailbox `uint64 slots[2][8]` (epoch parity x source rank) that its peers can write; an entry
// is (epoch << 32) | f32 bits, stored with ONE 64-bit system-scope store so value and flag arrive together.
struct XgpuParams {
  uint64_t mailbox[8];   // device pointers of every rank's mailbox (own included), indexed by rank
  uint32_t rank, nranks, epoch, pad;
  uint64_t index_offset; // arg ops: global index of this rank's element 0 (outer-axis shard offset)
};
// mailbox layout: [0,128) value slots[2][8]; [128,256) index slots[2][8] (arg ops: second word, same epoch tag)
constexpr uint32_t kMailboxIndexOffset = 128;

enum : int { OP_SUM = 0, OP_PROD = 1, OP_MAX = 2, OP_MIN = 3, OP_ARGMAX = 4, OP_ARGMIN = 5 };
enum : int { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };

constexpr int kMaxWarps = 32;
// workspace layout (host mirrors this): [0, 16 KiB) f32 partials, [16 KiB, 48 KiB) u64 partial indices, then the ticket
constexpr uint32_t kWsMaxBlocks = 4096;
constexpr uint32_t kWsIdxOffset = kWsMaxBlocks * 4;
constexpr uint32_t kWsTicketOffset = kWsIdxOffset + kWsMaxBlocks * 8;

// ------------------------------------------------------------------------------------------------ value ops
template <int OP>
struct ValOp;
template <>
struct ValOp<OP_SUM> {
  static __device__ __forceinline__ float identity() { return 0.f; }
  static __device__ __forceinline__ float apply(float a, float b) { return a + b; }
};
template <>
struct ValOp<OP_PROD> {
  static __device__ __forceinline__ float identity() { return 1.f; }
  static __device__ __forceinline__ float apply(float a, float b) { return a * b; }
};
// max/min propagate NaN (like the reference's `max`/`min` on floats lowered to fmaxf would NOT; we choose the
// numpy/IEEE-754-2019 "maximum" semantics and the oracle states the same rule).
template <>
struct ValOp<OP_MAX> {
  static __device__ __forceinline__ float identity() { return -INFINITY; }
  static __device__ __forceinline__ float apply(float a, float b) { return (a != a || b != b) ? NAN : fmaxf(a, b); }
};
template <>
struct ValOp<OP_MIN> {
  static __device__ __forceinline__ float identity() { return INFINITY; }
  static __device__ __forceinline__ float apply(float a, float b) { return (a != a || b != b) ? NAN : fminf(a, b); }
};

// ------------------------------------------------------------------------------------------------ loads
__device__ __forceinline__ float4 ldg_stream_v4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

struct float8 {
  float4 lo, hi;
};
__device__ __forceinline__ float8 ldg_stream_v8(const float* p) {
  float8 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v.lo.x), "=f"(v.lo.y), "=f"(v.lo.z), "=f"(v.lo.w), "=f"(v.hi.x), "=f"(v.hi.y), "=f"(v.hi.z),
                 "=f"(v.hi.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

template <int DT>
struct Elem;
template <>
struct Elem<DT_F32> {
  using T = float;
  static constexpr int VEC = 4;  // elements per 128-bit load
  static __device__ __forceinline__ float get(const void* base, uint64_t i) { return reinterpret_cast<const float*>(base)[i]; }
  static __device__ __forceinline__ void unpack(uint4 r, float (&f)[4]) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  }
};
template <>
struct Elem<DT_F16> {
  using T = __half;
  static constexpr int VEC = 8;
  static __device__ __forceinline__ float get(const void* base, uint64_t i) { return __half2float(reinterpret_cast<const __half*>(base)[i]); }
  static __device__ __forceinline__ void unpack(uint4 r, float (&f)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      f[2 * j] = t.x; f[2 * j + 1] = t.y;
    }
  }
};
template <>
struct Elem<DT_BF16> {
  using T = __nv_bfloat16;
  static constexpr int VEC = 8;
  static __device__ __forceinline__ float get(const void* base, uint64_t i) { return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]); }
  static __device__ __forceinline__ void unpack(uint4 r, float (&f)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f[2 * j] = __uint_as_float(w[j] << 16);
      f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
    }
  }
};

// ------------------------------------------------------------------------------------------------ block stages
template <int OP>
__device__ __forceinline__ float warp_reduce(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = ValOp<OP>::apply(v, __shfl_down_sync(0xffffffffu, v, off));
  return v;
}

// Result valid in thread 0.
template <int OP>
__device__ __forceinline__ float block_reduce(float v, float* smem /* kMaxWarps */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  v = warp_reduce<OP>(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = (lane < nwarps) ? smem[lane] : ValOp<OP>::identity();
    v = warp_reduce<OP>(v);
  }
  __syncthreads();
  return v;
}

// ------------------------------------------------------------------------------------------------ arg ops
// `better(a, ia, b, ib)`: should (b, ib) replace (a, ia)?  NaN is the extreme; ties keep the lower index.
template <int OP>
__device__ __forceinline__ bool arg_better(float a, uint64_t ia, float b, uint64_t ib) {
  const bool a_nan = a != a, b_nan = b != b;
  if (a_nan || b_nan) {
    if (a_nan && b_nan) return ib < ia;
    return b_nan;
  }
  if constexpr (OP == OP_ARGMAX) {
    return (b > a) || (b == a && ib < ia);
  } else {
    return (b < a) || (b == a && ib < ia);
  }
}

template <int OP>
__device__ __forceinline__ void warp_arg_reduce(float& v, uint64_t& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_down_sync(0xffffffffu, v, off);
    const uint64_t oi = __shfl_down_sync(0xffffffffu, i, off);
    if (arg_better<OP>(v, i, ov, oi)) { v = ov; i = oi; }
  }
}

template <int OP>
__device__ __forceinline__ void block_arg_reduce(float& v, uint64_t& i, float* sv, uint64_t* si) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  warp_arg_reduce<OP>(v, i);
  if (lane == 0) { sv[warp] = v; si[warp] = i; }
  __syncthreads();
  if (warp == 0) {
    if (lane < nwarps) { v = sv[lane]; i = si[lane]; }
    else { v = (OP == OP_ARGMAX) ? -INFINITY : INFINITY; i = ~0ull; }
    warp_arg_reduce<OP>(v, i);
  }
  __syncthreads();
}

__device__ __forceinline__ float arg_identity(int op) { return op == OP_ARGMAX ? -INFINITY : INFINITY; }

// ================================================================================================ reduce over ALL elements
// Grid-stride over 128-bit (VEC elements) vectors, UNROLL independent loads in flight per thread, one accumulator per
// load slot and vector lane.  Per-block partial -> workspace; the last block to finish (ticket) reduces the partials in
// block order (deterministic for a fixed grid) and writes out[0] * scale.
__device__ __forceinline__ void st_sys_u64(uint64_t addr, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_sys_u64(uint64_t addr) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int OP, int DT, int UNROLL, bool WIDE /* 256-bit loads, f32 only */, bool XGPU = false, bool BLOCKED = false>
__device__ __forceinline__ void reduce_all_body(const ReduceParams& p, const XgpuParams* xg = nullptr) {
  using E = Elem<DT>;
  constexpr int VEC = WIDE ? 8 : E::VEC;
  __shared__ float s_red[kMaxWarps];
  __shared__ bool s_last;

  const uint64_t n = p.len;
  const uint64_t nvec = n / VEC;
  const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t nthreads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const char* base = reinterpret_cast<const char*>(p.in);

  float acc[UNROLL][VEC];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u)
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[u][j] = ValOp<OP>::identity();

  uint64_t v = tid;
  if constexpr (BLOCKED) {
    // tile = blockDim * UNROLL consecutive vectors (64 KB for 512 threads x 8 x 16 B), tiles dealt round-robin to blocks:
    // at any instant the whole grid reads ONE contiguous window instead of UNROLL windows nthreads apart
    const uint64_t tile = static_cast<uint64_t>(blockDim.x) * UNROLL;
    const uint64_t ntiles = nvec / tile;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const uint64_t base_v = t * tile + threadIdx.x;
      uint4 r[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) r[u] = ldg_stream_u4(base + (base_v + static_cast<uint64_t>(u) * blockDim.x) * 16);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        float f[VEC];
        E::unpack(r[u], f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[u][j] = ValOp<OP>::apply(acc[u][j], f[j]);
      }
    }
    v = ntiles * tile + tid;  // leftover vectors: grid-stride below
  } else
  // main: UNROLL vectors per thread per trip, all loads issued before any use
  for (; v + static_cast<uint64_t>(UNROLL - 1) * nthreads < nvec; v += static_cast<uint64_t>(UNROLL) * nthreads) {
    if constexpr (WIDE) {
      float8 r[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) r[u] = ldg_stream_v8(reinterpret_cast<const float*>(base) + (v + u * nthreads) * 8);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        acc[u][0] = ValOp<OP>::apply(acc[u][0], r[u].lo.x); acc[u][1] = ValOp<OP>::apply(acc[u][1], r[u].lo.y);
        acc[u][2] = ValOp<OP>::apply(acc[u][2], r[u].lo.z); acc[u][3] = ValOp<OP>::apply(acc[u][3], r[u].lo.w);
        acc[u][4] = ValOp<OP>::apply(acc[u][4], r[u].hi.x); acc[u][5] = ValOp<OP>::apply(acc[u][5], r[u].hi.y);
        acc[u][6] = ValOp<OP>::apply(acc[u][6], r[u].hi.z); acc[u][7] = ValOp<OP>::apply(acc[u][7], r[u].hi.w);
      }
    } else {
      uint4 r[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) r[u] = ldg_stream_u4(base + (v + u * nthreads) * 16);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        float f[VEC];
        E::unpack(r[u], f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[u][j] = ValOp<OP>::apply(acc[u][j], f[j]);
      }
    }
  }
  // remaining whole vectors
  for (; v < nvec; v += nthreads) {
    if constexpr (WIDE) {
      float8 r = ldg_stream_v8(reinterpret_cast<const float*>(base) + v * 8);
      acc[0][0] = ValOp<OP>::apply(acc[0][0], r.lo.x); acc[0][1] = ValOp<OP>::apply(acc[0][1], r.lo.y);
      acc[0][2] = ValOp<OP>::apply(acc[0][2], r.lo.z); acc[0][3] = ValOp<OP>::apply(acc[0][3], r.lo.w);
      acc[0][4] = ValOp<OP>::apply(acc[0][4], r.hi.x); acc[0][5] = ValOp<OP>::apply(acc[0][5], r.hi.y);
      acc[0][6] = ValOp<OP>::apply(acc[0][6], r.hi.z); acc[0][7] = ValOp<OP>::apply(acc[0][7], r.hi.w);
    } else {
      float f[VEC];
      E::unpack(ldg_stream_u4(base + v * 16), f);
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[0][j] = ValOp<OP>::apply(acc[0][j], f[j]);
    }
  }
  // scalar tail (n % VEC elements)
  float local = ValOp<OP>::identity();
  for (uint64_t i = nvec * VEC + tid; i < n; i += nthreads) local = ValOp<OP>::apply(local, E::get(base, i));
#pragma unroll
  for (int u = 0; u < UNROLL; ++u)
#pragma unroll
    for (int j = 0; j < VEC; ++j) local = ValOp<OP>::apply(local, acc[u][j]);

  const float block_val = block_reduce<OP>(local, s_red);

  float* partials = reinterpret_cast<float*>(p.ws);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(p.ws + kWsTicketOffset);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_val;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    if constexpr (OP == OP_SUM) {
      // grid stage of a sum in f64: <= 4096 partials, so the only f32 roundings are inside the blocks and the last one
      __shared__ double s_dred[kMaxWarps];
      double d = 0.0;
      for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) d += static_cast<double>(__ldcg(partials + i));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) d += __shfl_down_sync(0xffffffffu, d, off);
      if ((threadIdx.x & 31) == 0) s_dred[threadIdx.x >> 5] = d;
      __syncthreads();
      if (threadIdx.x < 32) {
        const int nwarps = (blockDim.x + 31) >> 5;
        d = (static_cast<int>(threadIdx.x) < nwarps) ? s_dred[threadIdx.x] : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) d += __shfl_down_sync(0xffffffffu, d, off);
        if constexpr (!XGPU) {
          if (threadIdx.x == 0) {
            reinterpret_cast<float*>(p.out)[0] = static_cast<float>(d * static_cast<double>(p.scale));
            *ticket = 0;  // ready for the next launch on this stream
          }
        } else {
          // ---- fused all-reduce: publish this rank's scalar into every peer's mailbox, gather the others, sum in rank order
          __shared__ float s_peer[8];
          d = __shfl_sync(0xffffffffu, d, 0);
          const float mine = static_cast<float>(d * static_cast<double>(p.scale));
          const uint32_t slot_base = (xg->epoch & 1u) * 8u;
          if (threadIdx.x < xg->nranks) {
            const uint32_t peer = threadIdx.x;
            st_sys_u64(xg->mailbox[peer] + (slot_base + xg->rank) * 8ull,
                       (static_cast<uint64_t>(xg->epoch) << 32) | __float_as_uint(mine));
            const uint64_t src = xg->mailbox[xg->rank] + (slot_base + peer) * 8ull;
            const uint64_t t0 = globaltimer_ns();
            uint64_t w = ld_sys_u64(src);
            while (static_cast<uint32_t>(w >> 32) != xg->epoch) {
              if (globaltimer_ns() - t0 > 4000000000ull) asm volatile("trap;");  // a peer never arrived: fail loudly
              w = ld_sys_u64(src);
            }
            s_peer[peer] = __uint_as_float(static_cast<uint32_t>(w));
          }
          __syncwarp();
          if (threadIdx.x == 0) {
            double total = 0.0;
            for (uint32_t r = 0; r < xg->nranks; ++r) total += static_cast<double>(s_peer[r]);  // same order on every rank
            reinterpret_cast<float*>(p.out)[0] = static_cast<float>(total);
            *ticket = 0;
          }
        }
      }
    } else {
      float f = ValOp<OP>::identity();
      for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) f = ValOp<OP>::apply(f, __ldcg(partials + i));
      f = block_reduce<OP>(f, s_red);
      if (threadIdx.x == 0) {
        reinterpret_cast<float*>(p.out)[0] = f * p.scale;
        *ticket = 0;
      }
    }
  }
}

template <int OP, int DT, bool XGPU = false>
__device__ __forceinline__ void argreduce_all_body(const ReduceParams& p, const XgpuParams* xg = nullptr) {
  using E = Elem<DT>;
  constexpr int VEC = E::VEC;
  __shared__ float s_v[kMaxWarps];
  __shared__ uint64_t s_i[kMaxWarps];
  __shared__ bool s_last;
  const uint64_t n = p.len, nvec = n / VEC;
  const uint64_t tid = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const uint64_t nthreads = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  const char* base = reinterpret_cast<const char*>(p.in);

  float bv = arg_identity(OP);
  uint64_t bi = ~0ull;
  for (uint64_t v = tid; v < nvec; v += nthreads) {
    float f[VEC];
    E::unpack(ldg_stream_u4(base + v * 16), f);
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      if (arg_better<OP>(bv, bi, f[j], v * VEC + j)) { bv = f[j]; bi = v * VEC + j; }
  }
  for (uint64_t i = nvec * VEC + tid; i < n; i += nthreads) {
    const float f = E::get(base, i);
    if (arg_better<OP>(bv, bi, f, i)) { bv = f; bi = i; }
  }
  block_arg_reduce<OP>(bv, bi, s_v, s_i);

  float* pv = reinterpret_cast<float*>(p.ws);
  uint64_t* pi = reinterpret_cast<uint64_t*>(p.ws + kWsIdxOffset);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(p.ws + kWsTicketOffset);
  if (threadIdx.x == 0) {
    pv[blockIdx.x] = bv;
    pi[blockIdx.x] = bi;
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    float v = arg_identity(OP);
    uint64_t i = ~0ull;
    for (uint32_t k = threadIdx.x; k < gridDim.x; k += blockDim.x) {
      const float ov = __ldcg(pv + k);
      const uint64_t oi = __ldcg(pi + k);
      if (arg_better<OP>(v, i, ov, oi)) { v = ov; i = oi; }
    }
    block_arg_reduce<OP>(v, i, s_v, s_i);
    if constexpr (!XGPU) {
      if (threadIdx.x == 0) {
        reinterpret_cast<uint32_t*>(p.out)[0] = static_cast<uint32_t>(i);
        *ticket = 0;
      }
    } else if (threadIdx.x < 32) {
      // ---- fused (value, index) exchange: NCCL has no arg-reduce; every rank publishes its pair, then selects in rank
      // order with the same tie rule (lowest GLOBAL index), so all ranks agree.  Global indices must fit 32 bits.
      __shared__ float s_pv[8];
      __shared__ uint32_t s_pi[8];
      v = __shfl_sync(0xffffffffu, v, 0);
      i = __shfl_sync(0xffffffffu, i, 0);
      const uint32_t gi = static_cast<uint32_t>(i + xg->index_offset);
      const uint32_t slot_base = (xg->epoch & 1u) * 8u;
      const uint64_t tag = static_cast<uint64_t>(xg->epoch) << 32;
      if (threadIdx.x < xg->nranks) {
        const uint32_t peer = threadIdx.x;
        st_sys_u64(xg->mailbox[peer] + (slot_base + xg->rank) * 8ull, tag | __float_as_uint(v));
        st_sys_u64(xg->mailbox[peer] + kMailboxIndexOffset + (slot_base + xg->rank) * 8ull, tag | gi);
        const uint64_t src = xg->mailbox[xg->rank] + (slot_base + peer) * 8ull;
        const uint64_t t0 = globaltimer_ns();
        uint64_t w0 = ld_sys_u64(src), w1 = ld_sys_u64(src + kMailboxIndexOffset);
        while (static_cast<uint32_t>(w0 >> 32) != xg->epoch || static_cast<uint32_t>(w1 >> 32) != xg->epoch) {
          if (globaltimer_ns() - t0 > 4000000000ull) asm volatile("trap;");
          w0 = ld_sys_u64(src);
          w1 = ld_sys_u64(src + kMailboxIndexOffset);
        }
        s_pv[peer] = __uint_as_float(static_cast<uint32_t>(w0));
        s_pi[peer] = static_cast<uint32_t>(w1);
      }
      __syncwarp();
      if (threadIdx.x == 0) {
        float bv = s_pv[0];
        uint64_t bi = s_pi[0];
        for (uint32_t r = 1; r < xg->nranks; ++r)
          if (arg_better<OP>(bv, bi, s_pv[r], s_pi[r])) { bv = s_pv[r]; bi = s_pi[r]; }
        reinterpret_cast<uint32_t*>(p.out)[0] = static_cast<uint32_t>(bi);
        *ticket = 0;
      }
    }
  }
}

// ================================================================================================ rows: [outer, len], inner == 1
// gridDim.x blocks walk rows; `blockDim.x / TPR` rows per block pass, TPR threads per row (TPR = 32: a warp per row,
// TPR = blockDim: a block per row).  128-bit loads when the row length allows it.
template <int OP, int DT>
__device__ __forceinline__ void reduce_rows_body(const ReduceParams& p, int tpr_log2) {
  using E = Elem<DT>;
  constexpr int VEC = E::VEC;
  __shared__ float s_red[kMaxWarps];
  const uint32_t tpr = 1u << tpr_log2;
  const uint32_t rows_per_block = blockDim.x >> tpr_log2;
  const uint32_t sub = threadIdx.x >> tpr_log2;  // which row of this block pass
  const uint32_t t = threadIdx.x & (tpr - 1);
  const char* base = reinterpret_cast<const char*>(p.in);
  const bool vec_ok = (p.len % VEC) == 0 && (p.in % 16) == 0;
  const uint64_t nvec = vec_ok ? p.len / VEC : 0;

  if (tpr <= 32 && vec_ok && nvec <= tpr) {
    // short rows (at most one 128-bit vector per thread): 4 independent rows in flight per thread group, so the loads of
    // consecutive rows overlap instead of serialising behind each row's shuffle tree
    const uint64_t rstride = static_cast<uint64_t>(gridDim.x) * rows_per_block;
    for (uint64_t row0 = static_cast<uint64_t>(blockIdx.x) * rows_per_block; row0 < p.outer; row0 += 4 * rstride) {
      uint4 q[4];
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t row = row0 + j * rstride + sub;
        ok[j] = row < p.outer && t < nvec;
        if (ok[j]) q[j] = ldg_stream_u4(base + (row * p.len + static_cast<uint64_t>(t) * VEC) * sizeof(typename E::T));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float r = ValOp<OP>::identity();
        if (ok[j]) {
          float f[VEC];
          E::unpack(q[j], f);
#pragma unroll
          for (int e = 0; e < VEC; ++e) r = ValOp<OP>::apply(r, f[e]);
        }
        for (uint32_t o = tpr >> 1; o > 0; o >>= 1) r = ValOp<OP>::apply(r, __shfl_down_sync(0xffffffffu, r, o));
        const uint64_t row = row0 + j * rstride + sub;
        if (t == 0 && row < p.outer) reinterpret_cast<float*>(p.out)[row] = r * p.scale;
      }
    }
    return;
  }

  for (uint64_t row0 = static_cast<uint64_t>(blockIdx.x) * rows_per_block; row0 < p.outer;
       row0 += static_cast<uint64_t>(gridDim.x) * rows_per_block) {
    const uint64_t row = row0 + sub;
    float a0 = ValOp<OP>::identity(), a1 = a0, a2 = a0, a3 = a0;
    if (row < p.outer) {
      const uint64_t off = row * p.len;
      uint64_t v = t;
      for (; v + 3ull * tpr < nvec; v += 4ull * tpr) {
        uint4 r0 = ldg_stream_u4(base + (off + (v)*VEC) * sizeof(typename E::T));
        uint4 r1 = ldg_stream_u4(base + (off + (v + tpr) * VEC) * sizeof(typename E::T));
        uint4 r2 = ldg_stream_u4(base + (off + (v + 2ull * tpr) * VEC) * sizeof(typename E::T));
        uint4 r3 = ldg_stream_u4(base + (off + (v + 3ull * tpr) * VEC) * sizeof(typename E::T));
        float f0[VEC], f1[VEC], f2[VEC], f3[VEC];
        E::unpack(r0, f0); E::unpack(r1, f1); E::unpack(r2, f2); E::unpack(r3, f3);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          a0 = ValOp<OP>::apply(a0, f0[j]); a1 = ValOp<OP>::apply(a1, f1[j]);
          a2 = ValOp<OP>::apply(a2, f2[j]); a3 = ValOp<OP>::apply(a3, f3[j]);
        }
      }
      for (; v < nvec; v += tpr) {
        float f[VEC];
        E::unpack(ldg_stream_u4(base + (off + v * VEC) * sizeof(typename E::T)), f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) a0 = ValOp<OP>::apply(a0, f[j]);
      }
      for (uint64_t i = nvec * VEC + t; i < p.len; i += tpr) a1 = ValOp<OP>::apply(a1, E::get(base, off + i));
    }
    float r = ValOp<OP>::apply(ValOp<OP>::apply(a0, a1), ValOp<OP>::apply(a2, a3));
    if (tpr <= 32) {
      // sub-warp tree: rows never straddle a warp because tpr divides 32
      for (uint32_t o = tpr >> 1; o > 0; o >>= 1) r = ValOp<OP>::apply(r, __shfl_down_sync(0xffffffffu, r, o));
      if (t == 0 && row < p.outer) reinterpret_cast<float*>(p.out)[row] = r * p.scale;
    } else {
      r = block_reduce<OP>(r, s_red);  // tpr == blockDim.x: one row per block
      if (threadIdx.x == 0 && row < p.outer) reinterpret_cast<float*>(p.out)[row] = r * p.scale;
    }
  }
}

template <int OP, int DT>
__device__ __forceinline__ void argreduce_rows_body(const ReduceParams& p, int tpr_log2) {
  using E = Elem<DT>;
  __shared__ float s_v[kMaxWarps];
  __shared__ uint64_t s_i[kMaxWarps];
  const uint32_t tpr = 1u << tpr_log2;
  const uint32_t rows_per_block = blockDim.x >> tpr_log2;
  const uint32_t sub = threadIdx.x >> tpr_log2, t = threadIdx.x & (tpr - 1);
  const char* base = reinterpret_cast<const char*>(p.in);
  for (uint64_t row0 = static_cast<uint64_t>(blockIdx.x) * rows_per_block; row0 < p.outer;
       row0 += static_cast<uint64_t>(gridDim.x) * rows_per_block) {
    const uint64_t row = row0 + sub;
    float bv = arg_identity(OP);
    uint64_t bi = ~0ull;
    if (row < p.outer) {
      for (uint64_t i = t; i < p.len; i += tpr) {
        const float f = E::get(base, row * p.len + i);
        if (arg_better<OP>(bv, bi, f, i)) { bv = f; bi = i; }
      }
    }
    if (tpr <= 32) {
      for (uint32_t o = tpr >> 1; o > 0; o >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, bv, o);
        const uint64_t oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (arg_better<OP>(bv, bi, ov, oi)) { bv = ov; bi = oi; }
      }
      if (t == 0 && row < p.outer) reinterpret_cast<uint32_t*>(p.out)[row] = static_cast<uint32_t>(bi);
    } else {
      block_arg_reduce<OP>(bv, bi, s_v, s_i);
      if (threadIdx.x == 0 && row < p.outer) reinterpret_cast<uint32_t*>(p.out)[row] = static_cast<uint32_t>(bi);
    }
  }
}

// ================================================================================================ columns: [outer, len, inner], inner > 1
// One thread per output element (o, i); consecutive threads walk consecutive `inner` -> coalesced; 4 rows in flight.
template <int OP, int DT>
__device__ __forceinline__ void reduce_cols_body(const ReduceParams& p) {
  using E = Elem<DT>;
  constexpr int VEC = E::VEC;
  const uint64_t total = p.outer * p.inner;
  const char* base = reinterpret_cast<const char*>(p.in);
  if (p.inner % VEC == 0 && p.in % 16 == 0 && p.out % 16 == 0) {
    // vector path: each thread owns VEC consecutive columns -> one 128-bit load per row, 4 rows in flight
    const uint64_t inner_v = p.inner / VEC, total_v = p.outer * inner_v;
    for (uint64_t idx = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total_v;
         idx += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
      const uint64_t o = idx / inner_v, iv = idx - o * inner_v;
      const uint64_t off = o * p.len * p.inner + iv * VEC;  // elements
      float a[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < VEC; ++j) a[u][j] = ValOp<OP>::identity();
      uint64_t l = 0;
      for (; l + 3 < p.len; l += 4) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = ldg_stream_u4(base + (off + (l + u) * p.inner) * sizeof(typename E::T));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[VEC];
          E::unpack(r[u], f);
#pragma unroll
          for (int j = 0; j < VEC; ++j) a[u][j] = ValOp<OP>::apply(a[u][j], f[j]);
        }
      }
      for (; l < p.len; ++l) {
        float f[VEC];
        E::unpack(ldg_stream_u4(base + (off + l * p.inner) * sizeof(typename E::T)), f);
#pragma unroll
        for (int j = 0; j < VEC; ++j) a[0][j] = ValOp<OP>::apply(a[0][j], f[j]);
      }
      float* dst = reinterpret_cast<float*>(p.out) + o * p.inner + iv * VEC;
#pragma unroll
      for (int j = 0; j < VEC; j += 4) {
        float4 v;
        v.x = ValOp<OP>::apply(ValOp<OP>::apply(a[0][j], a[1][j]), ValOp<OP>::apply(a[2][j], a[3][j])) * p.scale;
        v.y = ValOp<OP>::apply(ValOp<OP>::apply(a[0][j + 1], a[1][j + 1]), ValOp<OP>::apply(a[2][j + 1], a[3][j + 1])) * p.scale;
        v.z = ValOp<OP>::apply(ValOp<OP>::apply(a[0][j + 2], a[1][j + 2]), ValOp<OP>::apply(a[2][j + 2], a[3][j + 2])) * p.scale;
        v.w = ValOp<OP>::apply(ValOp<OP>::apply(a[0][j + 3], a[1][j + 3]), ValOp<OP>::apply(a[2][j + 3], a[3][j + 3])) * p.scale;
        reinterpret_cast<float4*>(dst)[j / 4] = v;
      }
    }
    return;
  }
  for (uint64_t idx = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t o = idx / p.inner, i = idx - o * p.inner;
    const uint64_t off = o * p.len * p.inner + i;
    float a0 = ValOp<OP>::identity(), a1 = a0, a2 = a0, a3 = a0;
    uint64_t l = 0;
    for (; l + 3 < p.len; l += 4) {
      const float f0 = E::get(base, off + (l)*p.inner), f1 = E::get(base, off + (l + 1) * p.inner);
      const float f2 = E::get(base, off + (l + 2) * p.inner), f3 = E::get(base, off + (l + 3) * p.inner);
      a0 = ValOp<OP>::apply(a0, f0); a1 = ValOp<OP>::apply(a1, f1);
      a2 = ValOp<OP>::apply(a2, f2); a3 = ValOp<OP>::apply(a3, f3);
    }
    for (; l < p.len; ++l) a0 = ValOp<OP>::apply(a0, E::get(base, off + l * p.inner));
    reinterpret_cast<float*>(p.out)[idx] = ValOp<OP>::apply(ValOp<OP>::apply(a0, a1), ValOp<OP>::apply(a2, a3)) * p.scale;
  }
}

template <int OP, int DT>
__device__ __forceinline__ void argreduce_cols_body(const ReduceParams& p) {
  using E = Elem<DT>;
  const uint64_t total = p.outer * p.inner;
  const char* base = reinterpret_cast<const char*>(p.in);
  for (uint64_t idx = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint64_t o = idx / p.inner, i = idx - o * p.inner;
    const uint64_t off = o * p.len * p.inner + i;
    float bv = arg_identity(OP);
    uint64_t bi = ~0ull;
    for (uint64_t l = 0; l < p.len; ++l) {
      const float f = E::get(base, off + l * p.inner);
      if (arg_better<OP>(bv, bi, f, l)) { bv = f; bi = l; }
    }
    reinterpret_cast<uint32_t*>(p.out)[idx] = static_cast<uint32_t>(bi);
  }
}

// ================================================================================================ entry points
#define REDUCE_ALL(NAME, OP, DT, UNROLL, WIDE)                                                     \
  extern "C" __global__ void __launch_bounds__(512) NAME(const __grid_constant__ ReduceParams p) {  \
    reduce_all_body<OP, DT, UNROLL, WIDE>(p);                                                      \
  }
#define ARGREDUCE_ALL(NAME, OP, DT)                                                                \
  extern "C" __global__ void __launch_bounds__(1024) NAME(const __grid_constant__ ReduceParams p) { \
    argreduce_all_body<OP, DT>(p);                                                                 \
  }
#define REDUCE_ROWS(NAME, OP, DT)                                                                                   \
  extern "C" __global__ void __launch_bounds__(512) NAME(const __grid_constant__ ReduceParams p, int tpr_log2) {    \
    reduce_rows_body<OP, DT>(p, tpr_log2);                                                                          \
  }
#define ARGREDUCE_ROWS(NAME, OP, DT)                                                                                \
  extern "C" __global__ void __launch_bounds__(1024) NAME(const __grid_constant__ ReduceParams p, int tpr_log2) {   \
    argreduce_rows_body<OP, DT>(p, tpr_log2);                                                                       \
  }
#define REDUCE_COLS(NAME, OP, DT)                                                                  \
  extern "C" __global__ void __launch_bounds__(256) NAME(const __grid_constant__ ReduceParams p) {  \
    reduce_cols_body<OP, DT>(p);                                                                   \
  }
#define ARGREDUCE_COLS(NAME, OP, DT)                                                               \
  extern "C" __global__ void __launch_bounds__(1024) NAME(const __grid_constant__ ReduceParams p) { \
    argreduce_cols_body<OP, DT>(p);                                                                \
  }

#define ALL_SHAPES(OPN, OP, DTN, DT)                                                \
  REDUCE_ALL(reduce_all_##OPN##_##DTN, OP, DT, (DT == DT_F32 ? 8 : 4), false)       \
  REDUCE_ROWS(reduce_rows_##OPN##_##DTN, OP, DT)           \
  REDUCE_COLS(reduce_cols_##OPN##_##DTN, OP, DT)
#define ALL_ARG_SHAPES(OPN, OP, DTN, DT)               \
  ARGREDUCE_ALL(reduce_all_##OPN##_##DTN, OP, DT)      \
  ARGREDUCE_ROWS(reduce_rows_##OPN##_##DTN, OP, DT)    \
  ARGREDUCE_COLS(reduce_cols_##OPN##_##DTN, OP, DT)
#define ALL_DTYPES(M, OPN, OP) M(OPN, OP, f32, DT_F32) M(OPN, OP, f16, DT_F16) M(OPN, OP, bf16, DT_BF16)

ALL_DTYPES(ALL_SHAPES, sum, OP_SUM)
ALL_DTYPES(ALL_SHAPES, prod, OP_PROD)
ALL_DTYPES(ALL_SHAPES, max, OP_MAX)
ALL_DTYPES(ALL_SHAPES, min, OP_MIN)
ALL_DTYPES(ALL_ARG_SHAPES, argmax, OP_ARGMAX)
ALL_DTYPES(ALL_ARG_SHAPES, argmin, OP_ARGMIN)

// local sum + cross-GPU all-reduce of the scalar in one launch (see XgpuParams)
extern "C" __global__ void __launch_bounds__(512) reduce_all_sum_f32_xgpu(const __grid_constant__ ReduceParams p,
                                                                          const __grid_constant__ XgpuParams xg) {
  reduce_all_body<OP_SUM, DT_F32, 8, false, true>(p, &xg);
}

#define REDUCE_ALL_BLOCKED(NAME, UNROLL)                                                           \
  extern "C" __global__ void __launch_bounds__(512) NAME(const __grid_constant__ ReduceParams p) {  \
    reduce_all_body<OP_SUM, DT_F32, UNROLL, false, false, true>(p);                                \
  }
REDUCE_ALL_BLOCKED(reduce_all_sum_f32_b4, 4)
REDUCE_ALL_BLOCKED(reduce_all_sum_f32_b8, 8)

extern "C" __global__ void __launch_bounds__(1024) reduce_all_argmax_f32_xgpu(const __grid_constant__ ReduceParams p,
                                                                              const __grid_constant__ XgpuParams xg) {
  argreduce_all_body<OP_ARGMAX, DT_F32, true>(p, &xg);
}
extern "C" __global__ void __launch_bounds__(1024) reduce_all_argmin_f32_xgpu(const __grid_constant__ ReduceParams p,
                                                                              const __grid_constant__ XgpuParams xg) {
  argreduce_all_body<OP_ARGMIN, DT_F32, true>(p, &xg);
}

// tuning variants of the headline kernel (f32 sum over all elements); the host picks one by name.
REDUCE_ALL(reduce_all_sum_f32_u4, OP_SUM, DT_F32, 4, false)
REDUCE_ALL(reduce_all_sum_f32_u16, OP_SUM, DT_F32, 16, false)
REDUCE_ALL(reduce_all_sum_f32_u2, OP_SUM, DT_F32, 2, false)
REDUCE_ALL(reduce_all_sum_f32_w2, OP_SUM, DT_F32, 2, true)
REDUCE_ALL(reduce_all_sum_f32_w4, OP_SUM, DT_F32, 4, true)
