"""CPU: host-side logic that needs no device (layouts, shape rules, sharding, synthetic data mirrors)."""
import numpy as np
import pytest

from cubecl_b200 import reduce as b200_reduce
from cubecl_b200 import synth
from cubecl_b200.client import contiguous_strides, optimal_align, pitched_layout
from cubecl_b200.distributed import shard_range


def test_contiguous_strides():
    assert contiguous_strides([2, 3, 4]) == [12, 4, 1]
    assert contiguous_strides([5]) == [1]
    assert contiguous_strides([]) == []


def test_reduce_output_shape():
    assert b200_reduce.output_shape([4, 5, 6], 1) == [4, 6]
    assert b200_reduce.output_shape([4, 5, 6], -1) == [4, 5]
    assert b200_reduce.output_shape([7], 0) == [1]
    assert b200_reduce.output_shape([4, 5], None) == [1]
    with pytest.raises(ValueError):
        b200_reduce.output_shape([4, 5], 2)
    assert b200_reduce.output_dtype("argmax") == "u32" and b200_reduce.output_dtype("sum") == "f32"


@pytest.mark.parametrize("total,world", [(64, 8), (10, 4), (3, 8), (0, 2), (1 << 28, 8)])
def test_shard_range_partitions(total, world):
    spans = [shard_range(total, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_hash_generator_is_deterministic_and_uniform():
    a = synth.uniform_f32(5, 1 << 16, 0.0, 1.0)
    b = synth.uniform_f32(5, 1 << 16, 0.0, 1.0)
    assert np.array_equal(a, b)
    assert a.min() >= 0.0 and a.max() < 1.0
    assert abs(a.mean() - 0.5) < 0.01
    c = synth.uniform_f32(5, 100, -1.0, 1.0, start=1000)
    d = synth.uniform_at(5, np.arange(1000, 1100), -1.0, 1.0)
    assert np.array_equal(c, d)
    assert not np.array_equal(synth.uniform_f32(6, 100, 0, 1), synth.uniform_f32(7, 100, 0, 1))


def test_device_dtype_round_trip():
    x = np.array([0.0, 1.5, -2.25, 1e-3, 300.0], dtype=np.float32)
    for dt in ("f32", "f16", "bf16"):
        back = synth.from_device_dtype(synth.to_device_dtype(x, dt), dt)
        assert np.allclose(back, x, rtol=1e-2)


def test_pitched_layout_restates_the_reference_policy():
    # PitchedMemoryLayoutPolicy::apply (allocator.rs:21-72) + optimal_align (memory_pool/handle.rs:255-263), mem_alignment 512
    assert optimal_align(1, 4) == 4                      # unit rows stay contiguous
    assert optimal_align(3, 4) == 16                     # 12 B -> 16 (minimum)
    assert optimal_align(72, 2) == 256                   # 144 B -> next pow2
    assert optimal_align(8192, 2) == 512                 # clamped to the buffer alignment
    assert pitched_layout([100, 72], 2) == ([128, 1], 100 * 256)          # 144-byte rows pitch to 256 B = 128 bf16
    assert pitched_layout([100, 72], 4) == ([128, 1], 100 * 512)          # 288-byte rows pitch to 512 B = 128 f32
    assert pitched_layout([8192, 8192], 2) == ([8192, 1], 8192 * 16384)   # already 512-aligned: compact
    assert pitched_layout([2, 3, 5], 4) == ([24, 8, 1], 6 * 32)           # 20-byte rows -> 32 B; outer strides compact over the pitch
    assert pitched_layout([7, 1], 4) == ([1, 1], 7 * 4)                   # last dim 1: no padding
    assert pitched_layout([9], 4) == ([1], 64)                            # rank 1: stride 1, size padded to the row alignment
    assert pitched_layout([], 4) == ([], 4)
