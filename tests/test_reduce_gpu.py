"""GPU parity: reductions through the C ABI vs the oracle."""
import numpy as np
import pytest

import oracle
from cubecl_b200 import ServerError, TensorHandle, reduce, synth

pytestmark = pytest.mark.gpu


def _run(client, x_f32, axis, op, dtype="f32"):
    dev = synth.to_device_dtype(x_f32, dtype)
    vals = synth.from_device_dtype(dev, dtype).reshape(x_f32.shape)
    t = TensorHandle.from_numpy(client, dev, dtype)
    out = reduce.launch_alloc(client, t, axis, op)
    return out.to_numpy(client), vals


def test_sum_things_kat(client, golden):
    # examples/sum_things/src/lib.rs:180,222-225: [-1,10,1,5] -> 15
    got, _ = _run(client, np.array(golden["sum_things"]["input"], dtype=np.float32), None, "sum")
    assert got.tolist() == [golden["sum_things"]["expected_sum"]]


@pytest.mark.parametrize("variant", ["auto", "u2", "u4", "u16", "b4", "b8", "w2", "w4"])
@pytest.mark.parametrize("n", [1, 3, 4, 1000, (1 << 20) + 5])
def test_sum_all_integer_pattern_exact(client, variant, n):
    # BASELINE config 1 pattern x[i] = i % 8: every partial is an exact integer, so the result must be exact
    client.set_option("reduce.variant", variant)
    try:
        x = (np.arange(n) % 8).astype(np.float32)
        got, _ = _run(client, x, None, "sum")
        assert float(got[0]) == float(x.astype(np.float64).sum())
    finally:
        client.set_option("reduce.variant", "auto")


@pytest.mark.parametrize("dtype", ["f32", "f16", "bf16"])
def test_sum_all_random_vs_f64(client, dtype):
    n = (1 << 22) + 13
    x = synth.uniform_f32(5, n, 0.0, 1.0)
    got, vals = _run(client, x, None, "sum", dtype)
    ref = oracle.sum_f64(vals)
    serial_err = abs(float(oracle.sum_serial_f32(vals)) - ref)
    gpu_err = abs(float(got[0]) - ref)
    assert gpu_err <= 1e-3 * oracle.sum_abs_f64(vals)      # north-star tolerance
    assert gpu_err <= 1e-6 * oracle.sum_abs_f64(vals)      # what a blocked f32 sum should achieve
    assert gpu_err <= serial_err + 1e-6 * abs(ref)          # never worse than the reference's own serial order


def test_sum_2pow28_exact_and_tolerance(client):
    # BASELINE config 4 at full size, data generated in HBM: (i) i%8 pattern -> 939,524,096 exactly
    n = 1 << 28
    t = TensorHandle.empty_contiguous(client, [n], "f32")
    client.fill_modulo(t.handle, "f32", n, 8)
    out = reduce.launch_alloc(client, t, None, "sum")
    assert float(out.to_numpy(client)[0]) == 939524096.0
    # (ii) U[0,1) seed 5: compare with the f64 sum of the host-regenerated stream (chunked)
    client.fill_uniform(t.handle, "f32", n, 5, 0.0, 1.0)
    out = reduce.launch_alloc(client, t, None, "sum")
    got = float(out.to_numpy(client)[0])
    ref = 0.0
    step = 1 << 24
    for s in range(0, n, step):
        ref += float(synth.uniform_f32(5, step, 0.0, 1.0, start=s).astype(np.float64).sum())
    assert abs(got - ref) <= 1e-6 * ref
    # idempotence: same input, same grid -> bitwise same answer
    out2 = reduce.launch_alloc(client, t, None, "sum")
    assert out2.to_numpy(client)[0] == np.float32(got)


@pytest.mark.parametrize("shape,axis", [([512, 8192], 1), ([128, 32768], 1), ([64, 256, 1024], 2), ([64, 64, 4096], 2),
                                        ([7, 33], 1), ([1000, 3], 1), ([4, 1 << 18], 1), ([5, 1 << 17], 0), ([3, 70, 11], 1),
                                        ([4096, 64], 0), ([2, 3, 4, 5], 2), ([9], 0)])
@pytest.mark.parametrize("op", ["sum", "mean", "max", "min", "prod", "argmax", "argmin"])
def test_axis_reductions(client, shape, axis, op):
    # row-sum shapes are the book's benchmarks (cubecl-book/.../benchmark.md, parallel_reduction*.md)
    n = int(np.prod(shape))
    if op == "prod":
        x = synth.uniform_f32(8, n, 0.9, 1.1).reshape(shape)
    else:
        x = synth.uniform_f32(8, n, -1.0, 1.0).reshape(shape)
    got, vals = _run(client, x, axis, op)
    exp = oracle.reduce(vals, axis, op)
    if op in ("argmax", "argmin", "max", "min"):
        assert np.array_equal(got, exp)
    elif op == "prod":
        assert np.allclose(got, exp, rtol=1e-3)
    else:
        ref = oracle.reduce_f64(vals, axis, op)
        scale = oracle.reduce_f64(np.abs(vals), axis, op)
        assert np.all(np.abs(got - ref) <= 1e-5 * scale + 1e-30)


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_axis_reductions_16bit(client, dtype):
    x = synth.uniform_f32(9, 64 * 1000, -1.0, 1.0).reshape(64, 1000)
    for axis in (0, 1):
        got, vals = _run(client, x, axis, "sum", dtype)
        assert np.allclose(got, oracle.reduce_f64(vals, axis, "sum"), rtol=0, atol=1e-3)
        got, vals = _run(client, x, axis, "argmax", dtype)
        assert np.array_equal(got, oracle.reduce(vals, axis, "argmax"))


def test_argmax_ties_and_nan_rules(client):
    # unpinned by the reference: lowest index among equal maxima; NaN is the extreme and the first NaN wins
    x = np.zeros(100003, dtype=np.float32)
    x[[777, 50000, 99999]] = 7.0
    got, _ = _run(client, x, None, "argmax")
    assert got.tolist() == [777]
    x[60000] = np.nan
    x[90000] = np.nan
    got, _ = _run(client, x, None, "argmax")
    assert got.tolist() == [60000]
    got, _ = _run(client, x, None, "argmin")
    assert got.tolist() == [60000]
    got, _ = _run(client, x, None, "max")
    assert np.isnan(got[0])
    rows = np.array([[1, 5, 5, 2], [np.nan, 9, np.nan, 0], [3, 3, 3, 3]], dtype=np.float32)
    got, vals = _run(client, rows, 1, "argmax")
    assert got.tolist() == oracle.reduce(vals, 1, "argmax").tolist() == [1, 0, 0]
    got, _ = _run(client, rows, 0, "argmin")
    assert got.tolist() == oracle.reduce(rows, 0, "argmin").tolist()


def test_bad_axis_is_deferred(client):
    t = TensorHandle.from_numpy(client, np.ones((4, 4), dtype=np.float32), "f32")
    out = TensorHandle.empty_contiguous(client, [4], "u32")  # wrong output dtype for a sum
    reduce.launch(client, t, out, 1, "sum")
    with pytest.raises(ServerError):
        client.sync()


def test_pitched_and_permuted_inputs(client):
    # TensorHandle::empty pitches rows (allocator.rs:21-72): [100, 72] f32 rows of 288 B pitch to 512 B; the reduce must
    # see the logical tensor.  Also a transposed (stride-swapped) view.
    rows, cols = 100, 72
    x = synth.uniform_f32(12, rows * cols, -1.0, 1.0).reshape(rows, cols)
    t = TensorHandle.empty(client, [rows, cols], "f32")
    assert t.strides[0] > cols
    host = np.zeros((rows, t.strides[0]), dtype=np.float32)
    host[:, :cols] = x
    host[:, cols:] = 1e30  # padding must never be read
    client.write(t.handle, host)
    for axis in (0, 1, None):
        got = reduce.launch_alloc(client, t, axis, "sum").to_numpy(client)
        assert np.allclose(got, oracle.reduce_f64(x, axis, "sum"), rtol=0, atol=1e-4)
        got = reduce.launch_alloc(client, t, axis, "argmax").to_numpy(client)
        assert np.array_equal(got, oracle.reduce(x, axis, "argmax"))
    tt = TensorHandle.from_numpy(client, x, "f32").transposed()          # logical [cols, rows]
    got = reduce.launch_alloc(client, tt, 1, "max").to_numpy(client)
    assert np.array_equal(got, x.T.max(axis=1))


def test_fuzz_shapes_axes_ops_dtypes(client):
    # 80 seeded random reductions: rank 1-4, any axis or all, every op and input dtype, odd extents (vector tails, short
    # and long rows, split paths) -- value ops vs the f64 oracle, index/extremum ops exactly
    rng = np.random.default_rng(77)
    ops = ["sum", "mean", "max", "min", "prod", "argmax", "argmin"]
    for case in range(80):
        rank = int(rng.integers(1, 5))
        budget = int(rng.choice([300, 5000, 200000, 3000000]))
        shape = []
        for d in range(rank):
            hi = max(2, int(round(budget ** (1.0 / (rank - d)))) * 2)
            ext = int(rng.integers(1, hi))
            shape.append(ext)
            budget = max(1, budget // ext)
        axis = None if rng.random() < 0.2 else int(rng.integers(0, rank))
        op = ops[case % len(ops)]
        dtype = ["f32", "f16", "bf16"][case % 3]
        n = int(np.prod(shape))
        lo, hi = (0.97, 1.03) if op == "prod" else (-1.0, 1.0)
        x = synth.uniform_f32(500 + case, n, lo, hi).reshape(shape)
        if op in ("argmax", "argmin", "max", "min") and n > 8:      # plant ties so the lowest-index rule is exercised
            flat = x.reshape(-1)
            flat[rng.integers(0, n, 4)] = 2.0 if op in ("argmax", "max") else -2.0
        got, vals = _run(client, x, axis, op, dtype)
        exp = oracle.reduce(vals, axis, op)
        tag = (case, shape, axis, op, dtype)
        if op in ("argmax", "argmin", "max", "min"):
            assert np.array_equal(got, exp), tag
        elif op == "prod":
            assert np.allclose(got, oracle.reduce_f64(vals, axis, "prod"), rtol=2e-3), tag
        else:
            ref = oracle.reduce_f64(vals, axis, op)
            scale = oracle.reduce_f64(np.abs(vals), axis, op)
            assert np.all(np.abs(got - ref) <= 2e-5 * scale + 1e-30), tag
