"""CPU model of the GEMM's tail split (deterministic split-K): a Python transcription of `unit_decode` and of the slab /
ticket exchange in cubecl_b200/csrc/gemm_tcgen05.cu, checked for the invariants the kernel relies on -- every (tile, k-block)
is computed exactly once, every slice is non-empty, and the reduced tile does not depend on which slice arrives last."""
import itertools
import random

import numpy as np


def unit_decode(u, full_tiles, split_s, num_kb):
    """-> (tile, kb0, kb1, slice, partial); same integer arithmetic as the device function."""
    if split_s <= 1 or u < full_tiles:
        return u, 0, num_kb, 0, False
    v = u - full_tiles
    tile, sl = full_tiles + v // split_s, v % split_s
    return tile, (num_kb * sl) // split_s, (num_kb * (sl + 1)) // split_s, sl, True


def test_units_cover_every_tile_and_k_block_once():
    rng = random.Random(7)
    for _ in range(300):
        clusters = rng.choice([37, 74, 148])
        total_tiles = rng.randint(1, 600)
        rem = total_tiles % clusters
        num_kb = rng.randint(1, 400)
        split_s = rng.randint(1, min(8, num_kb))
        full_tiles, split_tiles = (total_tiles - rem, rem) if split_s > 1 and rem else (total_tiles, 0)
        units = full_tiles + split_tiles * split_s if split_tiles else total_tiles
        seen = {}
        for u in range(units):
            tile, kb0, kb1, sl, partial = unit_decode(u, full_tiles, split_s if split_tiles else 1, num_kb)
            assert 0 <= kb0 < kb1 <= num_kb                      # never an empty slice (S <= num_kb is enforced on the host)
            assert partial == (tile >= full_tiles and split_tiles > 0)
            for kb in range(kb0, kb1):
                assert (tile, kb) not in seen
                seen[(tile, kb)] = u
        assert len(seen) == total_tiles * num_kb                 # nothing skipped, nothing doubled
        # work units are dealt round-robin to CTA pairs: whole tiles first, so the sliced units form the last round(s)
        first_partial = next((u for u in range(units) if unit_decode(u, full_tiles, split_s if split_tiles else 1, num_kb)[4]), units)
        assert first_partial == full_tiles


def test_slab_exchange_is_order_independent_and_resets_its_ticket():
    # S slices publish f32 partials and take a ticket; whoever draws S - 1 adds the slabs in SLICE order and zeroes the ticket
    rng = np.random.default_rng(3)
    order_matters = False
    for S in (2, 3, 4):
        partials = [rng.standard_normal(64).astype(np.float32) * (10.0 ** rng.integers(-3, 4)) for _ in range(S)]
        results = set()
        for order in itertools.permutations(range(S)):
            ticket, slabs, out = 0, {}, None
            for sl in order:                                     # arrival order of the slices
                slabs[sl] = partials[sl]                         # publish (threadfence) ...
                old, ticket = ticket, ticket + 1                 # ... then atomicAdd
                if old == S - 1:                                 # last arriver: ordered reduction, ticket left ready for the next launch
                    ticket = 0
                    acc = np.zeros(64, dtype=np.float32)
                    for i in range(S):
                        acc = (acc + slabs[i]).astype(np.float32)
                    out = acc
            assert ticket == 0 and out is not None
            results.add(out.tobytes())
        assert len(results) == 1                                 # bit-identical for every arrival order
        # an arrival-order sum would NOT be (f32 addition is not associative): that is what the ordered reduction buys
        naive = set()
        for order in itertools.permutations(range(S)):
            acc = np.zeros(64, dtype=np.float32)
            for i in order:
                acc = (acc + partials[i]).astype(np.float32)
            naive.add(acc.tobytes())
        order_matters |= len(naive) > 1
    assert order_matters


def test_host_policy_formula_matches_measured_cases():
    # time(S) = full_waves + ceil(rem * S / C) / S + (14 + 8 S) / k_blocks, split for a >= 8 % gain (capi.cpp launch_tcgen05)
    def choose(total_tiles, clusters, num_kb):
        rem, full = total_tiles % clusters, total_tiles // clusters
        if rem == 0:
            return 1
        base = best = full + 1.0
        pick = 1
        for s in range(2, 5):
            if num_kb // s < 8:
                break
            t = full + -(-rem * s // clusters) / s + (14.0 + 8.0 * s) / num_kb
            if t < best - 1e-9:
                best, pick = t, s
        return pick if base - best >= 0.08 * base else 1
    assert choose(1024, 74, 128) == 1      # bf16 8192^3: tail 84 % full
    assert choose(256, 74, 64) == 1        # bf16 4096^3: the slab exchange eats the gain
    assert choose(256, 74, 384) == 2       # 3xTF32 4096^3: 384 k-blocks per tile
    assert choose(8, 74, 256) == 4         # 512^2 x 16384 on 256x128 tiles
    assert choose(32, 74, 128) == 2        # 1024^2 x 8192
