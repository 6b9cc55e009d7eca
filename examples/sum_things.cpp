// examples/sum_things (reference: examples/sum_things/src/lib.rs:178-227) through the C++ host layer: the same
// [-1, 10, 1, 5] demo plus the cmma.rs:552-576 golden matmul, every number computed on the GPU through the C ABI.
// Build: g++ -std=c++17 -Iinclude examples/sum_things.cpp -Lcubecl_b200/lib -lcubecl_b200 -Wl,-rpath,$PWD/cubecl_b200/lib
// Exit code 0 = all checks passed (used by tests/test_cpp_host_gpu.py on the GPU box).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cubecl_b200.hpp"

using namespace cubecl;

static uint16_t f32_to_f16_small_int(float v) {  // exact for the small non-negative integers used below
  if (v == 0.f) return 0;
  int e = 0;
  float m = v;
  while (m >= 2.f) { m /= 2.f; ++e; }
  return static_cast<uint16_t>(((e + 15) << 10) | static_cast<int>((m - 1.f) * 1024.f));
}

int main() {
  try {
    ComputeClient client(0);
    const float input[4] = {-1.f, 10.f, 1.f, 5.f};
    TensorHandle in = TensorHandle::new_contiguous({4}, client.create_from_slice(input, sizeof(input)), DType::F32);
    TensorHandle out = TensorHandle::empty(client, {1}, DType::F32);
    reduce::launch(client, in, out, -1, reduce::Op::Sum);
    float sum = 0;
    auto bytes = client.read_one(out.handle);
    std::memcpy(&sum, bytes.data(), 4);
    std::printf("[cuda-b200 - Basic]\n [%g, %g, %g, %g]\n", sum, sum, sum, sum);
    if (sum != 15.f) return 2;

    // Out = Lhs @ Rhs.T with lhs[i] = i, rhs[i] = i % 8 (f16, 16x16x16): row r must be 504 + 896 r (cmma.rs:552-576)
    std::vector<uint16_t> lhs(256), rhs(256);
    for (int i = 0; i < 256; ++i) { lhs[i] = f32_to_f16_small_int(static_cast<float>(i)); rhs[i] = f32_to_f16_small_int(static_cast<float>(i % 8)); }
    TensorHandle a = TensorHandle::new_contiguous({16, 16}, client.create_from_slice(lhs.data(), 512), DType::F16);
    TensorHandle b = TensorHandle::new_contiguous({16, 16}, client.create_from_slice(rhs.data(), 512), DType::F16).transposed();
    TensorHandle c = TensorHandle::empty(client, calculate_matmul_output(a.shape, b.shape), DType::F32);
    matmul::launch(client, a, b, c);
    auto cb = client.read_one(c.handle);
    const float* cf = reinterpret_cast<const float*>(cb.data());
    for (int r = 0; r < 16; ++r)
      for (int n = 0; n < 16; ++n)
        if (cf[r * 16 + n] != 504.f + 896.f * r) { std::printf("golden mismatch at %d,%d: %g\n", r, n, cf[r * 16 + n]); return 3; }
    std::printf("cmma golden ok (row r = 504 + 896 r)\n");

    // block-scaled matmul: e4m3 ones [32, 64] x [16, 64], row scales 2^(i % 4), column scales 2^-(j % 3), one per 32 K
    // -> out[i, j] = 64 * 2^(i % 4 - j % 3), exact (MmaDefinition::execute_scaled semantics, frontend/cmma.rs:798-840)
    std::vector<uint8_t> ones(32 * 64, 0x38), sa(32 * 2), sb(16 * 2);
    for (int i = 0; i < 32; ++i) sa[2 * i] = sa[2 * i + 1] = static_cast<uint8_t>(127 + i % 4);
    for (int j = 0; j < 16; ++j) sb[2 * j] = sb[2 * j + 1] = static_cast<uint8_t>(127 - j % 3);
    TensorHandle ml = TensorHandle::new_contiguous({32, 64}, client.create_from_slice(ones.data(), 32 * 64), DType::F8E4M3);
    TensorHandle mr = TensorHandle::new_contiguous({16, 64}, client.create_from_slice(ones.data(), 16 * 64), DType::F8E4M3);
    TensorHandle ms_a = TensorHandle::new_contiguous({32, 2}, client.create_from_slice(sa.data(), sa.size()), DType::UE8M0);
    TensorHandle ms_b = TensorHandle::new_contiguous({16, 2}, client.create_from_slice(sb.data(), sb.size()), DType::UE8M0);
    TensorHandle mo = TensorHandle::empty(client, {32, 16}, DType::F32);
    matmul::launch_scaled(client, ml, mr, ms_a, ms_b, mo);
    auto mb = client.read_one(mo.handle);
    const float* mf = reinterpret_cast<const float*>(mb.data());
    for (int i = 0; i < 32; ++i)
      for (int j = 0; j < 16; ++j)
        if (mf[i * 16 + j] != std::ldexp(64.f, i % 4 - j % 3)) { std::printf("scaled mismatch at %d,%d: %g\n", i, j, mf[i * 16 + j]); return 5; }
    std::printf("block-scaled matmul ok (out[i,j] = 64 * 2^(i%%4 - j%%3))\n");

    // deferred errors: inner dims differ -> nothing thrown at launch, ServerError at sync
    TensorHandle bad = TensorHandle::empty(client, {24, 16}, DType::F16);
    matmul::launch(client, a, bad, c);
    bool raised = false;
    try { client.sync(); } catch (const ServerError&) { raised = true; }
    if (!raised) return 4;
    std::printf("deferred error surfaced at sync\n");
    return 0;
  } catch (const std::exception& e) {
    std::printf("error: %s\n", e.what());
    return 1;
  }
}
