/* oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's semantics for the dense-LA hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file's
 * library.  The product (cubecl_b200/) never imports, links or calls it; there is no CPU fallback in the product.
 *
 * Why a restatement and not the reference itself: the reference's CPU runtime (crates/cubecl-cpu, an LLVM JIT) needs a
 * Rust toolchain + an LLVM 22 bundle fetched from git (Cargo.toml:177); neither exists here, it has no
 * bf16/CMMA/plane ops, and the tree ships no matmul / N-element reduce kernel to run on it.  So `oracle/_ref` does not
 * exist for this project ("unbuildable") and this file is the oracle.
 *
 * PINNING: checked in tests/test_oracle.py against every golden vector the reference's own tests hold for this path
 * (tests/golden/reference_golden.json, extracted from the reference sources by tests/golden/make_golden.py):
 *   cmma.rs:552-576 (simple_1 expected), cmma.rs:834-891 (tf32), cmma.rs:976-1002 (strided), cmma.rs:695-721 (formula),
 *   cmma.rs:1099-1196 (manual mma generator), plane.rs:154-189 (plane_sum), all_reduce.rs:52-59, sum_things lib.rs:180,
 *   shape.rs:1022-1063 (shape rule, in cubecl_b200/matmul.py).
 * NOT pinned by any reference vector (the reference has none): argmax/argmin tie + NaN rule, max/min NaN rule, and
 * NCCL's floating-point summation order.  Those rules are stated here and in DESIGN.md ("parity unpinned" items).
 *
 * Arithmetic notes: compiled with -ffp-contract=off so `sum += a * b` is a separate f32 multiply and f32 add, as the
 * Rust reference computes it (Rust never contracts to FMA).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ reductions */

/* examples/sum_things/src/lib.rs:11-18 (sum_basic): `sum += input[i]` for i in 0..end, f32, left to right. */
float oracle_sum_serial_f32(const float* x, size_t n) {
  float sum = 0.0f;
  for (size_t i = 0; i < n; ++i) sum += x[i];
  return sum;
}

/* Ground truth for tolerance checks (|gpu - f64| <= tol * sum|x|). */
double oracle_sum_f64(const float* x, size_t n) {
  double sum = 0.0;
  for (size_t i = 0; i < n; ++i) sum += (double)x[i];
  return sum;
}

double oracle_sum_abs_f64(const float* x, size_t n) {
  double sum = 0.0;
  for (size_t i = 0; i < n; ++i) sum += fabs((double)x[i]);
  return sum;
}

/* examples/sum_things/src/lib.rs:96-99 (SumThenMul / series): out[u] = sum * input[u]. */
void oracle_sum_then_mul_f32(const float* x, size_t n, float* out) {
  const float s = oracle_sum_serial_f32(x, n);
  for (size_t i = 0; i < n; ++i) out[i] = s * x[i];
}

/* crates/cubecl-cpp/src/shared/plane.rs:61-70 + cuda/plane.rs:31-35: plane_sum = xor butterfly over 32 lanes,
 * `acc = acc + shfl_xor(acc, off)` for off = 1,2,4,8,16; every lane ends with the same total.  vals: [32][vec]. */
void oracle_plane_sum_f32(const float* vals, int vec, float* out /* [32][vec] */) {
  float cur[32 * 8], nxt[32 * 8];
  memcpy(cur, vals, sizeof(float) * 32 * (size_t)vec);
  for (int off = 1; off < 32; off <<= 1) {
    for (int lane = 0; lane < 32; ++lane)
      for (int v = 0; v < vec; ++v) nxt[lane * vec + v] = cur[lane * vec + v] + cur[(lane ^ off) * vec + v];
    memcpy(cur, nxt, sizeof(float) * 32 * (size_t)vec);
  }
  memcpy(out, cur, sizeof(float) * 32 * (size_t)vec);
}

/* ops: 0 sum, 1 prod, 2 max, 3 min, 4 argmax, 5 argmin, 6 mean -- numbering of b200_reduce_op.
 * Layout: contiguous [outer, len, inner]; output [outer, inner].
 * Value ops accumulate serially in f32 along the axis in increasing index (the reference's only CPU reduce,
 * cubecl-book/src/getting-started/src/bin/v1-cpu.rs:7-15, is exactly this loop for sum); *_f64 gives ground truth.
 * max/min: NaN propagates.  arg ops: lowest index among equal extrema; NaN is the extreme, first NaN wins. */
static int arg_better(int op, float a, uint64_t ia, float b, uint64_t ib) {
  const int an = a != a, bn = b != b;
  if (an || bn) {
    if (an && bn) return ib < ia;
    return bn;
  }
  if (op == 4) return (b > a) || (b == a && ib < ia);
  return (b < a) || (b == a && ib < ia);
}

void oracle_reduce_axis_f32(int op, const float* x, uint64_t outer, uint64_t len, uint64_t inner, void* out) {
  for (uint64_t o = 0; o < outer; ++o) {
    for (uint64_t i = 0; i < inner; ++i) {
      const float* p = x + o * len * inner + i;
      if (op == 4 || op == 5) {
        float bv = (op == 4) ? -INFINITY : INFINITY;
        uint64_t bi = ~(uint64_t)0;
        for (uint64_t l = 0; l < len; ++l)
          if (arg_better(op, bv, bi, p[l * inner], l)) { bv = p[l * inner]; bi = l; }
        ((uint32_t*)out)[o * inner + i] = (uint32_t)bi;
      } else {
        float acc = (op == 1) ? 1.0f : (op == 2) ? -INFINITY : (op == 3) ? INFINITY : 0.0f;
        for (uint64_t l = 0; l < len; ++l) {
          const float v = p[l * inner];
          if (op == 0 || op == 6) acc += v;
          else if (op == 1) acc *= v;
          else if (op == 2) acc = (acc != acc || v != v) ? NAN : (v > acc ? v : acc);
          else acc = (acc != acc || v != v) ? NAN : (v < acc ? v : acc);
        }
        if (op == 6) acc = acc * (float)(1.0 / (double)len);
        ((float*)out)[o * inner + i] = acc;
      }
    }
  }
}

void oracle_reduce_axis_f64(int op, const float* x, uint64_t outer, uint64_t len, uint64_t inner, double* out) {
  for (uint64_t o = 0; o < outer; ++o)
    for (uint64_t i = 0; i < inner; ++i) {
      const float* p = x + o * len * inner + i;
      double acc = (op == 1) ? 1.0 : 0.0;
      for (uint64_t l = 0; l < len; ++l) {
        if (op == 1) acc *= (double)p[l * inner]; else acc += (double)p[l * inner];
      }
      if (op == 6) acc /= (double)len;
      out[o * inner + i] = acc;
    }
}

/* ------------------------------------------------------------------------------------------------ matmul */

/* crates/cubecl-core/src/runtime_tests/cmma.rs:695-721 (test_simple_cube_expected), generalised with strides:
 * inputs already widened to f32; for each (m, n): sum = 0f32; for k ascending: sum += lhs[m,k] * rhs[k,n].
 * Strides in elements.  The reference indexes rhs as [n*k + k_idx], i.e. rhs_sk = 1, rhs_sn = K. */
void oracle_matmul_f32(const float* lhs, const float* rhs, float* out, uint64_t M, uint64_t N, uint64_t K,
                       uint64_t lhs_sm, uint64_t lhs_sk, uint64_t rhs_sk, uint64_t rhs_sn, uint64_t out_sm, uint64_t out_sn) {
  for (uint64_t m = 0; m < M; ++m)
    for (uint64_t n = 0; n < N; ++n) {
      float sum = 0.0f;
      for (uint64_t k = 0; k < K; ++k) sum += lhs[m * lhs_sm + k * lhs_sk] * rhs[k * rhs_sk + n * rhs_sn];
      out[m * out_sm + n * out_sn] = sum;
    }
}

/* f64 ground truth + sum |a||b| (the scale of the tolerance), same indexing. */
void oracle_matmul_f64(const float* lhs, const float* rhs, double* out, double* out_abs, uint64_t M, uint64_t N, uint64_t K,
                       uint64_t lhs_sm, uint64_t lhs_sk, uint64_t rhs_sk, uint64_t rhs_sn) {
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < (int64_t)M; ++m)
    for (uint64_t n = 0; n < N; ++n) {
      double sum = 0.0, sa = 0.0;
      for (uint64_t k = 0; k < K; ++k) {
        const double p = (double)lhs[m * lhs_sm + k * lhs_sk] * (double)rhs[k * rhs_sk + n * rhs_sn];
        sum += p;
        sa += fabs(p);
      }
      out[m * N + n] = sum;
      if (out_abs) out_abs[m * N + n] = sa;
    }
}

/* Selected output elements only (full-size checks: 8192^3 is 1.1 TFLOP, so tests sample rows/cols). */
void oracle_matmul_points_f64(const float* lhs, const float* rhs, const uint64_t* ms, const uint64_t* ns, uint64_t count,
                              double* out, double* out_abs, uint64_t K, uint64_t lhs_sm, uint64_t lhs_sk, uint64_t rhs_sk,
                              uint64_t rhs_sn) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)count; ++i) {
    double sum = 0.0, sa = 0.0;
    for (uint64_t k = 0; k < K; ++k) {
      const double p = (double)lhs[ms[i] * lhs_sm + k * lhs_sk] * (double)rhs[k * rhs_sk + ns[i] * rhs_sn];
      sum += p;
      sa += fabs(p);
    }
    out[i] = sum;
    out_abs[i] = sa;
  }
}

/* Block-scaled matmul, reference semantics: the expected-value loop of test_cmma_scaled / test_cmma_scaled_fp4
 * (crates/cubecl-core/src/runtime_tests/cmma.rs:1572-1590, 1688-1706):
 *   sum += lhs_val * lhs_scale * rhs_val * rhs_scale, l increasing, l_scales = l / (k / scales_factor), f32 throughout
 * (Rust evaluates a * b * c * d left to right).  lhs [M,K], rhs [N,K] (the test's col-major rhs), scales already widened
 * to f32: lhs_scales [M, K/block], rhs_scales [N, K/block].  out_f32 = reference order; out_f64 / out_abs = ground truth
 * and its tolerance scale sum |terms| (either may be NULL). */
void oracle_matmul_scaled(const float* lhs, const float* rhs, const float* lhs_scales, const float* rhs_scales, float* out_f32,
                          double* out_f64, double* out_abs, uint64_t M, uint64_t N, uint64_t K, uint64_t block) {
  const uint64_t ns = K / block;
#pragma omp parallel for schedule(static)
  for (int64_t mi = 0; mi < (int64_t)M; ++mi) {
    const uint64_t m = (uint64_t)mi;
    for (uint64_t n = 0; n < N; ++n) {
      float sum = 0.0f;
      double d = 0.0, dabs = 0.0;
      for (uint64_t l = 0; l < K; ++l) {
        const uint64_t ls = l / block;
        const float lv = lhs[m * K + l], lsc = lhs_scales[m * ns + ls], rv = rhs[n * K + l], rsc = rhs_scales[n * ns + ls];
        float t = lv * lsc;
        t = t * rv;
        t = t * rsc;
        sum += t;
        const double td = (double)lv * (double)lsc * (double)rv * (double)rsc;
        d += td;
        dabs += td < 0 ? -td : td;
      }
      if (out_f32) out_f32[m * N + n] = sum;
      if (out_f64) out_f64[m * N + n] = d;
      if (out_abs) out_abs[m * N + n] = dabs;
    }
  }
}

/* ------------------------------------------------------------------------------------------------ CPU baseline legs
 * "cubecl-cpu execution model": one worker per core, each owning a contiguous slice (crates/cubecl-cpu/src/runtime.rs:
 * 95-121, compute/threadpool/mod.rs:86-107).  Used by bench.py's cpu_baseline / --impl reference; reference-order
 * arithmetic inside each slice. */
int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

float oracle_sum_blocked_f32(const float* x, size_t n, int threads) {
  if (threads < 1) threads = 1;
  float partial[256];
  if (threads > 256) threads = 256;
  const size_t chunk = (n + (size_t)threads - 1) / (size_t)threads;
#pragma omp parallel for num_threads(threads) schedule(static, 1)
  for (int t = 0; t < threads; ++t) {
    const size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    float s = 0.0f;
    for (size_t i = lo; i < hi; ++i) s += x[i];
    partial[t] = s;
  }
  float total = 0.0f;
  for (int t = 0; t < threads; ++t) total += partial[t];
  return total;
}

/* Row-parallel reference-order matmul on bf16/f16 data already widened to f32: rhs given [N,K] (the reference layout)
 * so the inner loop streams both operands.  Every output is still ONE serial f32 sum over increasing k of unfused products
 * (test_simple_cube_expected, cmma.rs:695-721) -- bit-identical to the plain triple loop -- but 4 x 4 outputs are carried at once
 * (sixteen independent serial sums hide the add latency a single dependent chain is bound by) and the rhs is walked in panels
 * of 64 columns that stay in a core's L2 across all rows of the thread's slab (the plain loop re-streamed the whole rhs for
 * every row: memory-bound and, on a two-socket host, noisy -- 0.05 to 0.24 TFLOP/s on the same 128 cores). */
static void oracle_dot_block(const float* lhs, const float* rhs_nk, float* out, uint64_t N, uint64_t K, uint64_t m0, uint64_t mh,
                             uint64_t n0, uint64_t nh) {
  if (mh == 4 && nh == 4) {
    const float *a0 = lhs + (m0 + 0) * K, *a1 = lhs + (m0 + 1) * K, *a2 = lhs + (m0 + 2) * K, *a3 = lhs + (m0 + 3) * K;
    const float *b0 = rhs_nk + (n0 + 0) * K, *b1 = rhs_nk + (n0 + 1) * K, *b2 = rhs_nk + (n0 + 2) * K, *b3 = rhs_nk + (n0 + 3) * K;
    float c00 = 0.f, c01 = 0.f, c02 = 0.f, c03 = 0.f, c10 = 0.f, c11 = 0.f, c12 = 0.f, c13 = 0.f;
    float c20 = 0.f, c21 = 0.f, c22 = 0.f, c23 = 0.f, c30 = 0.f, c31 = 0.f, c32 = 0.f, c33 = 0.f;
    for (uint64_t k = 0; k < K; ++k) {
      const float x0 = a0[k], x1 = a1[k], x2 = a2[k], x3 = a3[k];
      const float y0 = b0[k], y1 = b1[k], y2 = b2[k], y3 = b3[k];
      c00 += x0 * y0; c01 += x0 * y1; c02 += x0 * y2; c03 += x0 * y3;
      c10 += x1 * y0; c11 += x1 * y1; c12 += x1 * y2; c13 += x1 * y3;
      c20 += x2 * y0; c21 += x2 * y1; c22 += x2 * y2; c23 += x2 * y3;
      c30 += x3 * y0; c31 += x3 * y1; c32 += x3 * y2; c33 += x3 * y3;
    }
    float* o = out + m0 * N + n0;
    o[0] = c00; o[1] = c01; o[2] = c02; o[3] = c03;
    o += N; o[0] = c10; o[1] = c11; o[2] = c12; o[3] = c13;
    o += N; o[0] = c20; o[1] = c21; o[2] = c22; o[3] = c23;
    o += N; o[0] = c30; o[1] = c31; o[2] = c32; o[3] = c33;
    return;
  }
  for (uint64_t m = m0; m < m0 + mh; ++m)
    for (uint64_t n = n0; n < n0 + nh; ++n) {
      float sum = 0.0f;
      const float* a = lhs + m * K;
      const float* b = rhs_nk + n * K;
      for (uint64_t k = 0; k < K; ++k) sum += a[k] * b[k];
      out[m * N + n] = sum;
    }
}

void oracle_matmul_blocked_f32(const float* lhs, const float* rhs_nk, float* out, uint64_t M, uint64_t N, uint64_t K, int threads) {
  if (threads < 1) threads = 1;
  const uint64_t m_blocks = (M + 3) / 4, panel = 64;
#pragma omp parallel num_threads(threads)
  {
    /* contiguous slab of 4-row blocks per thread; rhs panels outermost so a panel is read from memory once per thread */
    const uint64_t t = (uint64_t)omp_get_thread_num(), T = (uint64_t)omp_get_num_threads();
    const uint64_t lo = m_blocks * t / T, hi = m_blocks * (t + 1) / T;
    for (uint64_t p0 = 0; p0 < N; p0 += panel) {
      const uint64_t p1 = (p0 + panel < N) ? p0 + panel : N;
      for (uint64_t mb = lo; mb < hi; ++mb) {
        const uint64_t m0 = mb * 4, mh = (m0 + 4 <= M) ? 4 : M - m0;
        for (uint64_t n0 = p0; n0 < p1; n0 += 4) oracle_dot_block(lhs, rhs_nk, out, N, K, m0, mh, n0, (n0 + 4 <= p1) ? 4 : p1 - n0);
      }
    }
  }
}
