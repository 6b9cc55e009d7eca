"""CPU: pin the oracle against every golden vector the reference's tests hold for this path."""
import numpy as np
import pytest

import oracle
from cubecl_b200 import synth
from cubecl_b200.matmul import MatmulShapeError, calculate_matmul_output


def test_cmma_simple_1_golden(golden):
    # cmma.rs:386-519, expected cmma.rs:552-576: f16 lhs[i]=i, rhs[i]=i%8 stored [N,K], Out = Lhs @ Rhs.T, exact
    g = golden["cmma_simple_1"]
    lhs = np.arange(256, dtype=np.float32).astype(np.float16).astype(np.float32).reshape(16, 16)
    rhs_nk = (np.arange(256) % 8).astype(np.float16).astype(np.float32).reshape(16, 16)
    out = oracle.matmul_f32(lhs, rhs_nk.T)
    assert out.ravel().tolist() == g["expected"]
    # the closed form the survey quotes: row r = 504 + 896 r
    assert all(out[r, 0] == 504 + 896 * r for r in range(16))


def test_cmma_tf32_golden(golden):
    # cmma.rs:834-891: 16x16x8, rhs ROW-major [8,16]
    g = golden["cmma_tf32"]
    lhs = np.arange(128, dtype=np.float32).reshape(16, 8)
    rhs = (np.arange(128) % 8).astype(np.float32).reshape(8, 16)
    assert oracle.matmul_f32(lhs, rhs).ravel().tolist() == g["expected"]


def test_cmma_strided_golden(golden):
    # cmma.rs:932-1005: lhs buffer [16,32] with the right k-tile zero; only the left 16x16 tile (row stride 32) is used
    g = golden["cmma_strided"]
    m = n = 16
    k = 32
    i = np.arange(m * k)
    lhs_buf = np.where((i % k) < 16, i - (i // k) * 16, 0).astype(np.float16).astype(np.float32).reshape(m, k)
    rhs_buf = (np.arange(n * k) % 8).astype(np.float16).astype(np.float32)
    rhs_nk = rhs_buf[:256].reshape(16, 16)  # col-major, stride 16
    out = oracle.matmul_f32(lhs_buf[:, :16], rhs_nk.T)
    assert out.ravel().tolist() == g["expected"]
    assert g["expected"] == golden["cmma_simple_1"]["expected"]


@pytest.mark.parametrize("m,n,k", [(16, 16, 16), (32, 8, 16), (8, 32, 16), (16, 16, 32)])
def test_simple_cube_formula(m, n, k):
    # cmma.rs:695-721 restated with plain loops; the oracle must agree bit for bit
    lhs = np.arange(m * k, dtype=np.float32).astype(np.float16).astype(np.float32)
    rhs = (np.arange(k * n) % 8).astype(np.float16).astype(np.float32)
    exp = np.zeros(m * n, dtype=np.float32)
    for mi in range(m):
        for ni in range(n):
            s = np.float32(0)
            for ki in range(k):
                s = np.float32(s + np.float32(lhs[mi * k + ki] * rhs[ni * k + ki]))
            exp[mi * n + ni] = s
    out = oracle.matmul_f32(lhs.reshape(m, k), rhs.reshape(n, k).T)
    assert np.array_equal(out.ravel(), exp)


@pytest.mark.parametrize("m,n,k", [(16, 8, 16), (16, 8, 8)])
def test_cmma_manual_generator(m, n, k):
    # cmma.rs:1099-1196: lhs[i,j]=2i+j, rhs[i,j]=3i+j, integer dot products (3% tolerance in the reference; exact here)
    lhs = np.array([[2 * i + j for j in range(k)] for i in range(m)], dtype=np.float32)
    rhs = np.array([[3 * i + j for j in range(n)] for i in range(k)], dtype=np.float32)
    exp = np.array([[sum((2 * i + l) * (3 * l + j) for l in range(k)) for j in range(n)] for i in range(m)], dtype=np.float64)
    assert np.array_equal(oracle.matmul_f32(lhs, rhs).astype(np.float64), exp)


@pytest.mark.parametrize("vec", [1, 2, 4])
def test_plane_sum_golden(vec):
    # plane.rs:154-189: 32 lanes x vec, value = flat index, expected[v] = sum_k input[v + k*vec]
    inp = np.arange(32 * vec, dtype=np.float32)
    exp = inp[:vec].copy()
    for k in range(1, 32):
        exp += inp[k * vec:(k + 1) * vec]
    out = oracle.plane_sum(inp.reshape(32, vec))
    assert np.array_equal(out[0], exp)
    assert np.array_equal(out, np.broadcast_to(exp, (32, vec)))  # butterfly: every lane holds the total
    if vec == 1:
        assert exp[0] == 496


def test_sum_things_golden(golden):
    g = golden["sum_things"]
    assert g["input"] == [-1.0, 10.0, 1.0, 5.0]
    assert float(oracle.sum_serial_f32(g["input"])) == g["expected_sum"]
    assert oracle.sum_then_mul(g["input"]).tolist() == g["expected_series"]


def test_all_reduce_formula(golden):
    # all_reduce.rs:52-59: value = dev + j per device; result = sum(dev) + j * ndev, identical everywhere
    g = golden["all_reduce"]
    for ndev in (2, 4, 8):
        for j in range(g["num_handles"]):
            bufs = [np.full(g["size"], d + j, dtype=np.float32) for d in range(ndev)]
            total = np.sum(bufs, axis=0, dtype=np.float32)
            assert np.all(total == sum(range(ndev)) + j * ndev)


def test_config1_sum_things_2pow20():
    # BASELINE config 1: f32 reduce-sum N=2^20 on the CPU path. (ii) exact-integer pattern, (iii) U[0,1) seed 0
    n = 1 << 20
    x = (np.arange(n) % 8).astype(np.float32)
    assert float(oracle.sum_serial_f32(x)) == 3670016.0
    assert oracle.sum_f64(x) == 3670016.0
    u = synth.uniform_f32(0, n, 0.0, 1.0)
    ref = oracle.sum_f64(u)
    assert abs(float(oracle.sum_serial_f32(u)) - ref) <= 1e-3 * oracle.sum_abs_f64(u)
    assert abs(float(oracle.sum_blocked_f32(u, 8)) - ref) <= 1e-5 * oracle.sum_abs_f64(u)


def test_reduce_semantics_unpinned_rules():
    # not pinned by the reference (it has no argmax): lowest index on ties, first NaN wins, max/min propagate NaN
    x = np.array([[1, 5, 5, 2], [np.nan, 9, np.nan, 0], [3, 3, 3, 3]], dtype=np.float32)
    assert oracle.reduce(x, 1, "argmax").tolist() == [1, 0, 0]
    assert oracle.reduce(x, 1, "argmin").tolist() == [0, 0, 0]
    assert np.isnan(oracle.reduce(x, 1, "max")[1]) and oracle.reduce(x, 1, "max")[0] == 5
    assert oracle.reduce(x, 0, "argmax").tolist() == [1, 1, 1, 2]
    assert np.allclose(oracle.reduce(x[[0, 2]], None, "sum"), [25.0])
    assert np.allclose(oracle.reduce(x[[0, 2]], 0, "mean"), [2, 4, 4, 2.5])


def test_reduce_axis_layouts():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 5, 7)).astype(np.float32)
    for axis in (0, 1, 2):
        assert np.allclose(oracle.reduce(x, axis, "sum"), x.sum(axis=axis), rtol=1e-5, atol=1e-5)
        assert np.array_equal(oracle.reduce(x, axis, "argmax"), x.argmax(axis=axis).astype(np.uint32))
        assert np.array_equal(oracle.reduce(x, axis, "min"), x.min(axis=axis))
        assert np.allclose(oracle.reduce_f64(x, axis, "sum"), x.astype(np.float64).sum(axis=axis))


def test_matmul_shape_rule():
    # crates/cubecl-zspace/src/shape.rs:1022-1063
    assert calculate_matmul_output([2, 4], [4, 2]) == [2, 2]
    assert calculate_matmul_output([1, 3, 2, 4], [2, 1, 4, 2]) == [2, 3, 2, 2]
    with pytest.raises(MatmulShapeError):
        calculate_matmul_output([3, 2, 4], [2, 1, 4, 2])   # RankMismatch
    with pytest.raises(MatmulShapeError):
        calculate_matmul_output([1, 3, 2, 4], [2, 1, 3, 2])  # IncompatibleShapes
    with pytest.raises(MatmulShapeError):
        calculate_matmul_output([1, 3, 2, 4], [2, 2, 4, 2])  # IncompatibleDims


def test_cast_identity_on_small_ints():
    # cmma.rs:766-832 (test_cmma_cast_f16 / _bf16): f32 0..255 -> f16 / bf16 is the identity on representable ints
    v = np.arange(256, dtype=np.float32)
    assert np.array_equal(v.astype(np.float16).astype(np.float32), v)
    assert np.array_equal(synth.bf16_bits_to_f32(synth.f32_to_bf16_bits(v)), v)


def test_bf16_rounding_matches_torch():
    import torch
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 1e3, np.array([0.0, -0.0, np.inf, -np.inf, 1.0039062, 3.3895314e38], dtype=np.float32)])
    ours = synth.f32_to_bf16_bits(x)
    theirs = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(ours, theirs)


def test_matmul_batch_broadcast_and_f64():
    rng = np.random.default_rng(2)
    a = rng.standard_normal((1, 3, 4, 5)).astype(np.float32)
    b = rng.standard_normal((2, 1, 5, 6)).astype(np.float32)
    out = oracle.matmul_f32(a, b)
    assert out.shape == (2, 3, 4, 6)
    assert np.allclose(out, np.matmul(a.astype(np.float64), b.astype(np.float64)), rtol=1e-5, atol=1e-5)
    f64, fabs = oracle.matmul_f64(a[0, 0], b[1, 0])
    assert np.allclose(f64, a[0, 0].astype(np.float64) @ b[1, 0].astype(np.float64))
    ms, ns = np.array([0, 3, 2]), np.array([5, 0, 1])
    pts, _ = oracle.matmul_points_f64(a[0, 0], b[1, 0], ms, ns)
    assert np.allclose(pts, f64[ms, ns])
    assert np.all(fabs >= np.abs(f64) - 1e-12)


def test_fp8_codecs_match_torch():
    # the numpy fp8 encode/decode used to build operands and expectations, pinned against torch's CPU conversion
    import torch
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.standard_normal(5000).astype(np.float32) * s for s in (0.01, 1, 30, 500)]
                       + [np.array([0, -0.0, 448, 449, 1e9, -1e9, 2 ** -9, 2 ** -10, 57344, 60000, 2 ** -16, 2 ** -17], dtype=np.float32)])
    for kind, tdt in (("f8e4m3", torch.float8_e4m3fn), ("f8e5m2", torch.float8_e5m2)):
        t = torch.from_numpy(x).clamp(-torch.finfo(tdt).max, torch.finfo(tdt).max).to(tdt)
        assert np.array_equal(synth.f32_to_fp8_bits(x, kind), t.view(torch.uint8).numpy())
        assert np.array_equal(synth.fp8_bits_to_f32(t.view(torch.uint8).numpy(), kind), t.float().numpy())


# ------------------------------------------------------------------------------------------------ block-scaled matmul
def _reference_scaled_expected(m, n, k, factor, lhs_val, rhs_val, lhs_scale_bits, rhs_scale_bits):
    """Literal restatement (pure Python, f32 at every step) of the expected-value loop of test_cmma_scaled /
    test_cmma_scaled_fp4: crates/cubecl-core/src/runtime_tests/cmma.rs:1572-1590, 1688-1706."""
    f = np.float32
    out = np.zeros((m, n), dtype=np.float32)
    for i in range(m):
        for j in range(n):
            acc = f(0.0)
            for l in range(k):
                ls = l // (k // factor)
                sl = f(synth.ue8m0_to_f32(np.uint8(lhs_scale_bits(i, ls))))
                sr = f(synth.ue8m0_to_f32(np.uint8(rhs_scale_bits(ls, j))))
                acc = f(acc + f(f(f(f(lhs_val(i, l)) * sl) * f(rhs_val(l, j))) * sr))
            out[i, j] = acc
    return out


def test_scaled_oracle_matches_the_reference_expected_loop_fp8():
    # generators of test_cmma_scaled (cmma.rs:1518-1533): lhs = 2i + j, rhs = 3i + j (col-major), scales 120 + ...
    m, n, k, factor = 16, 8, 32, 1
    exp = _reference_scaled_expected(m, n, k, factor, lambda i, l: i * 2 + l, lambda l, j: l * 3 + j,
                                     lambda i, s: i * 2 + s + 120, lambda s, j: s * 3 + j + 120)
    lhs = np.array([[i * 2 + l for l in range(k)] for i in range(m)], dtype=np.float32)
    rhs_nk = np.array([[l * 3 + j for l in range(k)] for j in range(n)], dtype=np.float32)
    sa = synth.ue8m0_to_f32(np.array([[i * 2 + s + 120 for s in range(factor)] for i in range(m)], dtype=np.uint8))
    sb = synth.ue8m0_to_f32(np.array([[s * 3 + j + 120 for s in range(factor)] for j in range(n)], dtype=np.uint8))
    o32, o64, oabs = oracle.matmul_scaled(lhs, rhs_nk, sa, sb, 32)
    assert np.array_equal(o32, exp)
    assert np.max(np.abs(o32 - o64) / oabs) < 1e-6


def test_scaled_oracle_matches_the_reference_expected_loop_fp4():
    # test_cmma_scaled_fp4 (cmma.rs:1621-1640): e2m1 codes ((i + j) % 15) + 1, two scales per row (k = 64, factor 2)
    m, n, k, factor = 16, 8, 64, 2
    val = lambda a, b: float(synth.e2m1_codes_to_f32(np.uint8(((a + b) % 15) + 1)))
    exp = _reference_scaled_expected(m, n, k, factor, val, val, lambda i, s: i * 2 + s + 120, lambda s, j: s * 3 + j + 120)
    lhs = np.array([[val(i, l) for l in range(k)] for i in range(m)], dtype=np.float32)
    rhs_nk = np.array([[val(l, j) for l in range(k)] for j in range(n)], dtype=np.float32)
    sa = synth.ue8m0_to_f32(np.array([[i * 2 + s + 120 for s in range(factor)] for i in range(m)], dtype=np.uint8))
    sb = synth.ue8m0_to_f32(np.array([[s * 3 + j + 120 for s in range(factor)] for j in range(n)], dtype=np.uint8))
    o32, _, _ = oracle.matmul_scaled(lhs, rhs_nk, sa, sb, 32)
    assert np.array_equal(o32, exp)


def test_mx_codecs():
    codes = np.arange(16, dtype=np.uint8)
    vals = synth.e2m1_codes_to_f32(codes)
    assert list(vals[:8]) == [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0] and np.array_equal(vals[8:], -vals[:8])
    assert np.array_equal(synth.f32_to_e2m1_codes(vals)[1:8], codes[1:8]) and np.array_equal(synth.f32_to_e2m1_codes(vals)[9:], codes[9:])
    assert list(synth.f32_to_e2m1_codes(np.array([0.25, 0.75, 1.25, 2.5, 5.0, 7.0, -100.0], dtype=np.float32))) == [0, 2, 2, 4, 6, 7, 15]
    packed = synth.pack_e2m1x2(np.array([[1, 2, 3, 15]], dtype=np.uint8))
    assert packed.tolist() == [[0x21, 0xF3]] and synth.unpack_e2m1x2(packed).tolist() == [[1, 2, 3, 15]]   # even element -> low nibble
    assert synth.ue8m0_to_f32(np.array([127, 128, 120, 0], dtype=np.uint8)).tolist() == [1.0, 2.0, 2.0 ** -7, 2.0 ** -127]
    assert np.isnan(synth.ue8m0_to_f32(np.array([255], dtype=np.uint8))[0])
    sc = (np.arange(300 * 9) % 251).astype(np.uint8).reshape(300, 9)
    pk = synth.pack_scale_chunks(sc)
    assert pk.shape == (3, 3, 512)
    for r, s_ in ((0, 0), (31, 3), (32, 4), (200, 6), (299, 8)):
        assert pk[r // 128, s_ // 4, (r % 32) * 16 + ((r % 128) // 32) * 4 + (s_ % 4)] == sc[r, s_]
    assert pk[2, 2, (299 % 32) * 16 + ((299 % 128) // 32) * 4 + 1] == 127 and pk[2, 0, 31 * 16 + 3 * 4] == 127   # padding = 1.0
