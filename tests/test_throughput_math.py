"""CPU: the reference's roofline arithmetic and sampling protocol, restated -- its own unit tests re-run here
(crates/cubecl-runtime/src/throughput/roofline.rs:109-213) plus a deterministic check of the benchmarker loops."""
import math

from cubecl_b200.throughput import (AchievedThroughput, ResourceBound, ThroughputBenchmarker, binding_achieved,
                                    binding_resource, score_resources)


def test_time_at_peak_is_amount_over_peak():
    assert ResourceBound(8, 4.0).time_at_peak() == 2.0


def test_time_at_peak_is_none_for_a_non_normal_peak():
    for p in (0.0, float("nan"), float("inf")):
        assert ResourceBound(8, p).time_at_peak() is None


def test_binding_resource_is_the_one_needing_the_most_time_at_peak():
    slower, faster = ResourceBound(8, 4.0), ResourceBound(8, 8.0)
    assert binding_resource([slower, faster]) == slower


def test_binding_resource_skips_non_normal_peaks_and_is_none_if_all_are():
    unusable, usable = ResourceBound(8, 0.0), ResourceBound(8, 4.0)
    assert binding_resource([unusable, usable]) == usable
    assert binding_resource([unusable]) is None
    assert binding_resource([]) is None


def test_score_resources_reports_achieved_rate_and_fraction_of_peak():
    scores = score_resources(1.0, [ResourceBound(100, 200.0), ResourceBound(400, 800.0)])
    assert (scores[0].achieved_per_s, scores[0].fraction_of_peak) == (100.0, 0.5)
    assert (scores[1].achieved_per_s, scores[1].fraction_of_peak) == (400.0, 0.5)


def test_a_zero_duration_reports_nan_instead_of_dividing_by_zero():
    s = score_resources(0.0, [ResourceBound(100, 200.0)])[0]
    assert math.isnan(s.achieved_per_s) and math.isnan(s.fraction_of_peak)


def test_resources_with_different_peaks_score_independently_and_pick_the_slower_one():
    read, write = ResourceBound(900_000, 1_000_000.0), ResourceBound(100_000, 200_000.0)
    assert binding_resource([read, write]) == read
    scores = score_resources(1.0, [read, write])
    assert (scores[0].achieved_per_s, scores[0].fraction_of_peak) == (900_000.0, 0.9)
    assert (scores[1].achieved_per_s, scores[1].fraction_of_peak) == (100_000.0, 0.5)
    assert binding_achieved(scores).fraction_of_peak == 0.9


def test_binding_achieved_skips_non_finite_entries_and_is_none_if_all_are():
    finite, bad = AchievedThroughput(10.0, 0.4), AchievedThroughput(float("nan"), float("nan"))
    assert binding_achieved([bad, finite]).fraction_of_peak == 0.4
    assert binding_achieved([bad]) is None and binding_achieved([]) is None


def test_fraction_is_not_clamped():
    assert score_resources(1.0, [ResourceBound(300, 200.0)])[0].fraction_of_peak == 1.5


def test_benchmarker_protocol_on_a_synthetic_kernel():
    # 50 us per iteration + 100 us fixed overhead, first calls 3x slower (cold): the protocol must grow iterations until
    # a sample lasts >= 20 ms and return the MIN per-iteration time
    calls = []

    def sample(iters):
        calls.append(iters)
        cold = 3.0 if len(calls) <= 2 else 1.0
        return (100e-6 + iters * 50e-6) * cold

    b = ThroughputBenchmarker()
    v = b.measure(sample, ops_count=1000)
    iters = calls[-1]
    assert (100e-6 + iters * 50e-6) * 1e3 >= b.TARGET_DURATION_MS
    assert abs(v.duration_s - (100e-6 + iters * 50e-6) / iters) < 1e-12
    assert abs(v.ops_per_s() - 1000 / v.duration_s) < 1e-6
    assert b.MIN_SAMPLES < sum(1 for c in calls if c == iters) <= b.MAX_SAMPLES + b.MAX_WARMUP
