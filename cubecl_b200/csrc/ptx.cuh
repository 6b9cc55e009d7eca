// Inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM / commit),
// cluster primitives.  Everything here is device-only and header-only; compiled into the prebuilt cubins.
//
// These replace what the reference would have obtained from NVRTC-compiled generated C++:
//   mbarrier  -> crates/cubecl-cpp/src/cuda/barrier.rs (reference emits cuda::barrier / mbarrier PTX)
//   TMA       -> crates/cubecl-cpp/src/cuda/tma.rs:11-45
//   MMA       -> crates/cubecl-cpp/src/shared/mma.rs:48-174 (wmma) -- here tcgen05, which the reference lacks
#pragma once
#include <cstdint>
#include <cuda.h>

namespace b200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}

__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// Map a shared::cta address of *this* CTA to the shared::cluster address of the same offset in CTA `rank`.
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// ---------------------------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Arrive on a barrier that lives in another CTA of the cluster (address from mapa_shared).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}

#ifndef B200_MBAR_TIMEOUT_NS
#define B200_MBAR_TIMEOUT_NS 4000000000ull  // a pipeline wait longer than 4 s is a deadlock: trap loudly, never hang the GPU
#endif

// try_wait with a suspend-time hint: the thread may be parked by the hardware for up to `ns` before the instruction
// returns false, instead of re-issuing the poll (fewer executed instructions per waiting warp, less issue-slot and power
// pressure next to the MMA / TMA warps).
__device__ __forceinline__ uint32_t mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return done;
}

#ifndef B200_MBAR_SUSPEND_NS
#define B200_MBAR_SUSPEND_NS 20000u
#endif

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  // slow path: parked polls; the deadline is only looked at every 64 polls (the clock read is not free either)
  uint64_t t0 = 0;
  uint32_t polls = 0;
  while (!mbar_try_wait_hint(bar, parity, B200_MBAR_SUSPEND_NS)) {
    if ((++polls & 63u) == 0) {
      const uint64_t now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > B200_MBAR_TIMEOUT_NS) asm volatile("trap;");
    }
  }
}

// ---------------------------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

// 3-D tiled load, signalling an mbarrier in this CTA.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 3-D tiled load issued by one CTA of a cta_group::2 pair; `cluster_bar` may live in the peer (leader) CTA.
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t cluster_bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Multicast 3-D load: the box lands at the same smem offset in every CTA of `mask`, each CTA's own barrier
// (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, uint16_t mask, int c0,
                                               int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6}], [%2], %3;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "h"(mask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 3-D tiled store smem -> global (bulk async-group completion); the box is clipped at the tensor's edges.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// 1-D bulk copy global -> shared (UBLKCP in SASS): `bytes` and both addresses multiples of 16; completes on `bar`.
// `policy` is an L2 cache policy (createpolicy): streaming reads use evict_first so they do not displace resident data.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      :
      : "r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}

// ---------------------------------------------------------------------------------------------- tcgen05
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
}

template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
  if constexpr (CG == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}

template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem].  KIND: 0/1 = kind::f16 (f16 / bf16 inputs), 2 = kind::tf32, 3/4 = kind::f8f6f4
// (e4m3 / e5m2 inputs), 5/6 = kind::i8 (u8 / s8 inputs, s32 accumulate).  The operand formats themselves are encoded in
// the instruction descriptor.
#define B200_UMMA_ASM(CGS, KINDS)                                                                             \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                            \
               "tcgen05.mma.cta_group::" CGS ".kind::" KINDS " [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),       \
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)                                          \
               : "memory")
template <int CG, int KIND>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (KIND <= 1) {
    if constexpr (CG == 1) B200_UMMA_ASM("1", "f16"); else B200_UMMA_ASM("2", "f16");
  } else if constexpr (KIND == 2) {
    if constexpr (CG == 1) B200_UMMA_ASM("1", "tf32"); else B200_UMMA_ASM("2", "tf32");
  } else if constexpr (KIND <= 4) {
    if constexpr (CG == 1) B200_UMMA_ASM("1", "f8f6f4"); else B200_UMMA_ASM("2", "f8f6f4");
  } else {
    if constexpr (CG == 1) B200_UMMA_ASM("1", "i8"); else B200_UMMA_ASM("2", "i8");
  }
}

// Block-scaled forms (MX formats): D[tmem] (+)= (A * SFA) * (B * SFB) with one scale per row per 32 K elements, the scale
// factors read from TMEM.  MXKIND 0 = kind::mxf8f6f4 (K = 32 per instruction, one scale per row: byte `sf_id` of the
// 32-bit TMEM word, selected in the instruction descriptor), 1 = kind::mxf4 with scale_vec::2X (packed e2m1, K = 64 per
// instruction, two scales per row: bytes sf_id, sf_id + 1), 2 = kind::mxf4nvf4 with scale_vec::4X (NVFP4: packed e2m1, K = 64,
// four ue4m3 scales per row -- one per 16 elements -- the whole 32-bit word).
#define B200_UMMA_SCALED_ASM(CGS, KINDS)                                                                          \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                \
               "tcgen05.mma.cta_group::" CGS ".kind::" KINDS " [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(d_tmem), \
               "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)                \
               : "memory")
template <int CG, int MXKIND>
__device__ __forceinline__ void umma_ss_scaled(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t sfa_tmem, uint32_t sfb_tmem, uint32_t accumulate) {
  if constexpr (MXKIND == 0) {
    if constexpr (CG == 1) B200_UMMA_SCALED_ASM("1", "mxf8f6f4.block_scale"); else B200_UMMA_SCALED_ASM("2", "mxf8f6f4.block_scale");
  } else if constexpr (MXKIND == 1) {
    if constexpr (CG == 1) B200_UMMA_SCALED_ASM("1", "mxf4.block_scale.scale_vec::2X");
    else B200_UMMA_SCALED_ASM("2", "mxf4.block_scale.scale_vec::2X");
  } else {
    if constexpr (CG == 1) B200_UMMA_SCALED_ASM("1", "mxf4nvf4.block_scale.scale_vec::4X");
    else B200_UMMA_SCALED_ASM("2", "mxf4nvf4.block_scale.scale_vec::4X");
  }
}

// smem -> TMEM copy of one scale-factor chunk: 32 rows x 128 bits, broadcast to the four 32-lane sub-partitions
// (columns [taddr, taddr + 4)).  Ordered with the tcgen05.mma instructions issued by the same thread.
template <int CG>
__device__ __forceinline__ void tmem_cp_32x128b_warpx4(uint32_t taddr, uint64_t smem_desc) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(smem_desc) : "memory");
  else
    asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(smem_desc) : "memory");
}

// Instruction descriptor of the block-scaled kinds (f32 accumulate, K-major operands); the scale-factor byte ids (bits
// 29-30 for A, 4-5 for B) are OR-ed in per instruction.
//   a_fmt/b_fmt: kind::mxf8f6f4 -> 0 = e4m3, 1 = e5m2; kind::mxf4 / mxf4nvf4 -> 1 = e2m1.  ue8m0: scale format bit (23).
__host__ __device__ constexpr uint32_t make_idesc_scaled(uint32_t a_fmt, uint32_t b_fmt, uint32_t umma_m, uint32_t umma_n,
                                                         uint32_t ue8m0 = 1) {
  return (a_fmt << 7) | (b_fmt << 10) | ((umma_n >> 3) << 17) | (ue8m0 << 23) | ((umma_m >> 4) << 24);
}

// All previously issued tcgen05.mma of this thread arrive (once) on `bar` when they retire.
// CG==2: the arrive is multicast to the barrier at the same smem offset in both CTAs of the pair.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns (thread t <- lane base+t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.  Offsets are byte values, multiples of 16.
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset >> 4   bits [46,48) version = 1   bits [61,64) layout type
//   layout: 2 = SWIZZLE_128B (16-byte swizzle atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms; the only layout the
//   hardware accepts for MN-major 32-bit (tf32) operands -- pairs with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_smem_desc(saddr, lbo_bytes, sbo_bytes, 2);
}

// Instruction descriptor for kind::f16 / kind::tf32, f32 accumulate.
//   fmt: kind::f16 -> 0 = f16, 1 = bf16; kind::tf32 -> 2; kind::f8f6f4 -> 0 = e4m3, 1 = e5m2; kind::i8 -> 0 = u8, 1 = s8.
//   *_mn: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t a_mn, uint32_t b_mn, uint32_t umma_m,
                                                  uint32_t umma_n, uint32_t c_fmt = 1 /* 1 = f32, 2 = s32 */) {
  return (c_fmt << 4)         // accumulator format
         | (fmt << 7)         // A format
         | (fmt << 10)        // B format
         | (a_mn << 15)       // A major
         | (b_mn << 16)       // B major
         | ((umma_n >> 3) << 17) | ((umma_m >> 4) << 24);
}

__host__ __device__ constexpr uint32_t make_idesc_ab(uint32_t fmt_a, uint32_t fmt_b, uint32_t a_mn, uint32_t b_mn, uint32_t umma_m,
                                                     uint32_t umma_n, uint32_t c_fmt) {
  return (c_fmt << 4) | (fmt_a << 7) | (fmt_b << 10) | (a_mn << 15) | (b_mn << 16) | ((umma_n >> 3) << 17) | ((umma_m >> 4) << 24);
}

}  // namespace b200
