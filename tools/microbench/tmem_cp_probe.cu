// Microbenchmark (diagnostic, not product): what a tcgen05.cp smem -> TMEM copy costs the tensor pipe, per copy shape, alone and
// interleaved with block-scaled MMAs -- the question behind the block-scaled GEMM's scale-factor copies (gemm_tcgen05.cu, SCALED).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cubecl_b200/csrc tools/microbench/tmem_cp_probe.cu -o tools/microbench/bin/tmem_cp_probe
// Output: one line per (shape, copies per group): cycles per copy alone; extra cycles per copy when each group of copies is
// followed by four kind::mxf8f6f4 MMAs (256 x 256 x 32, cta_group::2) that read the copied columns.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "ptx.cuh"
using namespace b200;

__device__ __forceinline__ void cp_shape(int shape, uint32_t taddr, uint64_t d32, uint64_t d128w, uint64_t d128n, uint64_t d64, uint64_t d4) {
  switch (shape) {
    case 0: asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(d32) : "memory"); break;
    case 1: asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(d128w) : "memory"); break;
    case 2: asm volatile("tcgen05.cp.cta_group::2.128x128b [%0], %1;" ::"r"(taddr), "l"(d128n) : "memory"); break;
    case 3: asm volatile("tcgen05.cp.cta_group::2.64x128b.warpx2::02_13 [%0], %1;" ::"r"(taddr), "l"(d64) : "memory"); break;
    default: asm volatile("tcgen05.cp.cta_group::2.4x256b [%0], %1;" ::"r"(taddr), "l"(d4) : "memory"); break;
  }
}

// mode bit 0: MMAs after each group; bit 1: copies alternate between two TMEM scale buffers (the MMAs read the fresh one)
extern "C" __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
    cp_probe(unsigned long long* out, int shape, int cps, int mode, int groups) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = smem_base, sb = smem_base + 16384, sf = smem_base + 32768;  // sf: 16 KB of scale images
  const uint32_t done_bar = sf + 16384, tmem_slot = done_bar + 8;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = cluster_ctarank() == 0;
  for (uint32_t i = threadIdx.x; i < 32768 / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_base + 4 * i), "r"(0x38383838u) : "memory");
  for (uint32_t i = threadIdx.x; i < 16384 / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sf + 4 * i), "r"(0x7F7F7F7Fu) : "memory");
  fence_proxy_async_smem();
  if (warp == 1 && lane == 0) {
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (warp == 1 && leader && lane == 0) {
    const uint64_t a_desc = make_smem_desc_sw128(sa, 16, 1024), b_desc = make_smem_desc_sw128(sb, 16, 1024);
    const uint64_t d32 = make_smem_desc(sf, 0, 128, 0);       // 32 rows x 16 B, 8-row groups 128 B apart
    const uint64_t d128w = make_smem_desc(sf, 128, 256, 0);   // 128 rows x 32 B: [16 groups][2 chunks][8 rows][16 B]
    const uint64_t d128n = make_smem_desc(sf, 0, 128, 0);     // 128 rows x 16 B
    const uint64_t d64 = make_smem_desc(sf, 0, 128, 0);       // 64 rows x 16 B
    const uint64_t d4 = make_smem_desc(sf, 128, 256, 0);      // 4 rows x 32 B
    constexpr uint32_t idesc = make_idesc_scaled(0, 0, 256, 256);
    const uint32_t cols_per_cp = (shape == 1 || shape == 4) ? 8u : 4u;
    // warm the pipe
    for (int k = 0; k < 4; ++k) umma_ss_scaled<2, 0>(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, tmem_base + 256, tmem_base + 260, k ? 1u : 0u);
    umma_commit<2>(done_bar);
    mbar_wait(done_bar, 0);
    tcgen05_fence_after();
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint32_t sfbuf = tmem_base + 256 + ((mode & 2) ? (g & 1) * 128u : 0u);
      for (int c = 0; c < cps; ++c) cp_shape(shape, sfbuf + c * cols_per_cp, d32, d128w, d128n, d64, d4);
      if (mode & 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss_scaled<2, 0>(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc | (uint32_t(k) << 29) | (uint32_t(k) << 4), sfbuf, sfbuf + 4, 1u);
      }
    }
    umma_commit<2>(done_bar);
    mbar_wait(done_bar, 1);
    tcgen05_fence_after();
    const long long t1 = clock64();
    out[0] = static_cast<unsigned long long>(t1 - t0);
  }
  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static unsigned long long run(unsigned long long* d_out, int shape, int cps, int mode, int groups) {
  const int smem = 32768 + 16384 + 1024 + 1024;
  CK(cudaFuncSetAttribute(cp_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  unsigned long long best = ~0ull;
  for (int rep = 0; rep < 3; ++rep) {
    cp_probe<<<2, 256, smem>>>(d_out, shape, cps, mode, groups);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    unsigned long long h = 0;
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    if (h < best) best = h;
  }
  return best;
}


// Second experiment: copies as mixes of shapes (n_x4 x 32x128b.warpx4, n_wide x 128x256b, n_128 x 128x128b per group), issued
//   mode 0: by the MMA thread, all copies first, then the four MMAs (what the GEMM does today)
//   mode 1: by the MMA thread, copies spread between the MMAs of the group (copy targets the OTHER scale buffer)
//   mode 2: by a thread of ANOTHER warp (warp 3), unsynchronised with the MMA thread (is the tcgen05 front end shared?)
extern "C" __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
    cp_mix_probe(unsigned long long* out, int n_x4, int n_wide, int n_128, int mode, int groups) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = smem_base, sb = smem_base + 16384, sf = smem_base + 32768;
  const uint32_t done_bar = sf + 16384, done2_bar = done_bar + 8, done3_bar = done_bar + 16, tmem_slot = done_bar + 24;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = cluster_ctarank() == 0;
  for (uint32_t i = threadIdx.x; i < 32768 / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_base + 4 * i), "r"(0x38383838u) : "memory");
  for (uint32_t i = threadIdx.x; i < 16384 / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sf + 4 * i), "r"(0x7F7F7F7Fu) : "memory");
  fence_proxy_async_smem();
  if (warp == 1 && lane == 0) {
    mbar_init(done_bar, 1);
    mbar_init(done2_bar, 1);
    mbar_init(done3_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint64_t d32 = make_smem_desc(sf, 0, 128, 0), d128w = make_smem_desc(sf, 128, 256, 0), d128n = make_smem_desc(sf, 0, 128, 0);
  const int n_cp = n_x4 + n_wide + n_128;
  auto one_cp = [&](int i, uint32_t buf) {   // copy i of the group: wide ones first
    const uint32_t col = buf + 8u * static_cast<uint32_t>(i);
    if (i < n_wide) cp_shape(1, col, d32, d128w, d128n, d32, d128w);
    else if (i < n_wide + n_x4) cp_shape(0, col, d32, d128w, d128n, d32, d128w);
    else cp_shape(2, col, d32, d128w, d128n, d32, d128w);
  };
  if (warp == 1 && leader && lane == 0) {
    const uint64_t a_desc = make_smem_desc_sw128(sa, 16, 1024), b_desc = make_smem_desc_sw128(sb, 16, 1024);
    constexpr uint32_t idesc = make_idesc_scaled(0, 0, 256, 256);
    for (int k = 0; k < 4; ++k) umma_ss_scaled<2, 0>(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, tmem_base + 256, tmem_base + 260, k ? 1u : 0u);
    umma_commit<2>(done_bar);
    mbar_wait(done_bar, 0);
    tcgen05_fence_after();
    const long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      const uint32_t cur = tmem_base + 256 + (g & 1) * 128u, nxt = tmem_base + 256 + ((g + 1) & 1) * 128u;
      if (mode == 0)
        for (int i = 0; i < n_cp; ++i) one_cp(i, cur);
      int issued = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_ss_scaled<2, 0>(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc | (uint32_t(k) << 29) | (uint32_t(k) << 4), cur, cur + 4, 1u);
        if (mode == 1) {   // after MMA k: this MMA's share of the NEXT group's copies
          const int upto = (n_cp * (k + 1) + 3) / 4;
          for (; issued < upto; ++issued) one_cp(issued, nxt);
        }
      }
    }
    umma_commit<2>(done_bar);
    mbar_wait(done_bar, 1);
    if (mode >= 2) mbar_wait(done2_bar, 0);
    if (mode == 3) mbar_wait(done3_bar, 0);
    tcgen05_fence_after();
    const long long t1 = clock64();
    out[0] = static_cast<unsigned long long>(t1 - t0);
  } else if ((warp == 3 || warp == 5) && leader && lane == 0) {
    // mode 2: every copy from warp 3; mode 3: even copies from warp 3, odd copies from warp 5 (two issuing warps)
    const int me = (warp == 3) ? 0 : 1;
    if ((mode == 2 && me == 0) || mode == 3) {
      for (int g = 0; g < groups; ++g) {
        const uint32_t nxt = tmem_base + 256 + ((g + 1) & 1) * 128u;
        for (int i = 0; i < n_cp; ++i)
          if (mode == 2 || (i & 1) == me) one_cp(i, nxt);
      }
      umma_commit<2>(me == 0 ? done2_bar : done3_bar);
    }
  }
  __syncwarp();
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc<2>(tmem_base, 512);
}

static unsigned long long run_mix(unsigned long long* d_out, int n_x4, int n_wide, int n_128, int mode, int groups) {
  const int smem = 32768 + 16384 + 1024 + 1024;
  CK(cudaFuncSetAttribute(cp_mix_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  unsigned long long best = ~0ull;
  for (int rep = 0; rep < 3; ++rep) {
    cp_mix_probe<<<2, 256, smem>>>(d_out, n_x4, n_wide, n_128, mode, groups);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    unsigned long long h = 0;
    CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
    if (h < best) best = h;
  }
  return best;
}

int main() {
  unsigned long long* d_out;
  CK(cudaMalloc(&d_out, 8));
  const char* names[5] = {"32x128b.warpx4", "128x256b", "128x128b", "64x128b.warpx2", "4x256b"};
  const int groups = 2000;
  const unsigned long long mma_only = run(d_out, 0, 0, 1, groups);
  printf("4 x mxf8 MMA (256x256x32, cta_group::2) per group, no copies: %.1f cycles per group\n", double(mma_only) / groups);
  for (int shape = 0; shape < 5; ++shape)
    for (int cps : {1, 3, 6, 12}) {
      if ((shape == 1 || shape == 4) && cps > 6) continue;  // 8 columns per copy: stay inside the 128-column buffers
      const unsigned long long alone = run(d_out, shape, cps, 0, groups);
      const unsigned long long with = run(d_out, shape, cps, 1, groups);
      const unsigned long long with2 = run(d_out, shape, cps, 3, groups);
      printf("%-16s %2d copies/group: alone %6.1f cyc/copy | with 4 MMAs: +%6.1f cyc/copy (group %7.1f) | two TMEM buffers: +%6.1f cyc/copy\n", names[shape], cps,
             double(alone) / (double(groups) * cps), (double(with) - double(mma_only)) / (double(groups) * cps), double(with) / groups,
             (double(with2) - double(mma_only)) / (double(groups) * cps));
    }
  {
    // can a tensor map replicate a chunk through a zero stride?  (16 B, 32 rows, 4 replicas) over one 512-byte chunk
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    void* buf;
    CK(cudaMalloc(&buf, 1 << 16));
    for (unsigned long long s2 : {0ull, 16ull, 512ull}) {
      CUtensorMap tm;
      cuuint64_t dims[3] = {16, 32, 4}, strides[2] = {16, s2};
      cuuint32_t box[3] = {16, 32, 4}, es[3] = {1, 1, 1};
      CUresult r = reinterpret_cast<encode_fn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("cuTensorMapEncodeTiled dims (16,32,4) strides (16,%llu): CUresult %d\n", s2, (int)r);
    }
  }
  printf("\nmixes per group of four MMAs (cycles per group; 512 = copies fully hidden):\n");
  const int mixes[][3] = {{3, 0, 0}, {1, 1, 0}, {0, 2, 0}, {0, 1, 1}, {6, 0, 0}, {2, 2, 0}, {0, 3, 0}, {12, 0, 0}, {0, 6, 0}, {0, 0, 3}, {0, 0, 2}, {2, 0, 0}};
  for (const auto& m : mixes) {
    printf("  %2d x warpx4 + %d x 128x256b + %d x 128x128b: burst-then-MMAs %7.1f | spread between MMAs %7.1f | from another warp %7.1f | from two other warps %7.1f\n", m[0], m[1], m[2],
           double(run_mix(d_out, m[0], m[1], m[2], 0, groups)) / groups, double(run_mix(d_out, m[0], m[1], m[2], 1, groups)) / groups,
           double(run_mix(d_out, m[0], m[1], m[2], 2, groups)) / groups, double(run_mix(d_out, m[0], m[1], m[2], 3, groups)) / groups);
  }
  return 0;
}
