"""Measurement methodology of the reference, restated (host-only arithmetic + the sampling protocol).

  ResourceBound / binding_resource / score_resources / binding_achieved
        crates/cubecl-runtime/src/throughput/roofline.rs:14-107 (fractions are NOT clamped to 1)
  ThroughputBenchmarker.{warmup, sample_peak_duration}
        crates/cubecl-runtime/src/throughput/benchmarker.rs:40-143: grow `iterations` until one sample lasts >= 20 ms, stop
        warming after 3 plateaus within 3 %, then take the MIN over 20..200 samples (stop after 12 stale ones)
  ThroughputValue.{ops_per_s, bytes_per_s}    crates/cubecl-runtime/src/throughput/base.rs:176-190

The reference clocks every sample with the HOST wall clock around launch + sync (compute_cmma.rs:20-39).  `device_sampler`
below offers both that protocol and CUDA events on the launching stream (what bench.py reports).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass
from typing import Callable, Optional, Sequence


@dataclass(frozen=True)
class ResourceBound:
    amount: int          # bytes or operations the run must move
    peak_per_s: float    # peak rate of that resource, same unit per second

    def time_at_peak(self) -> Optional[float]:
        """Seconds at peak; None for a zero / negative-subnormal / NaN / infinite peak (f64::is_normal)."""
        p = self.peak_per_s
        if p != p or math.isinf(p) or p == 0.0 or abs(p) < 2.2250738585072014e-308:
            return None
        t = self.amount / p
        return t if t >= 0 else None


@dataclass(frozen=True)
class AchievedThroughput:
    achieved_per_s: float
    fraction_of_peak: float  # not clamped


def binding_resource(bounds: Sequence[ResourceBound]) -> Optional[ResourceBound]:
    usable = [b for b in bounds if b.time_at_peak() is not None]
    return max(usable, key=lambda b: b.time_at_peak()) if usable else None


def score_resources(duration_s: float, bounds: Sequence[ResourceBound]) -> list[AchievedThroughput]:
    out = []
    for b in bounds:
        achieved = float("nan") if duration_s == 0 else b.amount / duration_s
        frac = achieved / b.peak_per_s if b.peak_per_s != 0 else (float("nan") if achieved != achieved or achieved == 0 else math.copysign(float("inf"), achieved))
        out.append(AchievedThroughput(achieved, frac))
    return out


def binding_achieved(scores: Sequence[AchievedThroughput]) -> Optional[AchievedThroughput]:
    finite = [s for s in scores if math.isfinite(s.fraction_of_peak)]
    return max(finite, key=lambda s: s.fraction_of_peak) if finite else None


@dataclass(frozen=True)
class ThroughputValue:
    ops_count: int
    duration_s: float

    def ops_per_s(self) -> float:
        return self.ops_count / self.duration_s if self.duration_s > 0 else 0.0

    bytes_per_s = ops_per_s


class ThroughputBenchmarker:
    MAX_WARMUP, MAX_ITERATIONS, PLATEAU_TOL, WARM_PATIENCE, TARGET_DURATION_MS = 50, 1000, 0.03, 3, 20.0
    MIN_SAMPLES, MAX_SAMPLES, REL_TOL, PATIENCE = 20, 200, 0.01, 12

    def measure(self, sample: Callable[[int], float], ops_count: int) -> ThroughputValue:
        """`sample(iterations)` runs the kernel `iterations` times and returns the elapsed SECONDS."""
        iterations = self.warmup(sample)
        return ThroughputValue(ops_count, self.sample_peak_duration(iterations, sample))

    def warmup(self, sample: Callable[[int], float]) -> int:
        best, stable, iterations = math.inf, 0, 1
        for _ in range(self.MAX_WARMUP):
            duration = sample(iterations) * 1000.0
            if duration < self.TARGET_DURATION_MS:
                if duration > 1e-6:
                    extra = math.ceil((self.TARGET_DURATION_MS - duration) / (duration / iterations))
                else:
                    extra = iterations
                iterations = min(iterations + max(extra, 1), self.MAX_ITERATIONS)
                best, stable = math.inf, 0
                continue
            per_iter = duration / iterations
            if per_iter < best * (1.0 - self.PLATEAU_TOL):
                best, stable = per_iter, 0
            else:
                best = min(best, per_iter)
                stable += 1
                if stable >= self.WARM_PATIENCE:
                    break
        return iterations

    def sample_peak_duration(self, iterations: int, sample: Callable[[int], float]) -> float:
        assert iterations > 0
        best, stale = math.inf, 0
        for i in range(self.MAX_SAMPLES):
            s = sample(iterations)
            if s < best * (1.0 - self.REL_TOL):
                best, stale = s, 0
            else:
                best = min(best, s)
                stale += 1
            if i > self.MIN_SAMPLES and stale >= self.PATIENCE:
                break
        return best / iterations


def device_sampler(client, launch: Callable[[], None], clock: str = "events") -> Callable[[int], float]:
    """sample(iterations) for a kernel launched through `client`. clock = "events" (CUDA events on the launching stream)
    or "host" (the reference's protocol: Instant::now() .. launch .. sync .. elapsed, compute_cmma.rs:20-39)."""
    if clock == "host":
        def sample(iterations: int) -> float:
            t0 = time.perf_counter()
            for _ in range(iterations):
                launch()
            client.sync()
            return time.perf_counter() - t0
        return sample
    e0, e1 = client.event(), client.event()

    def sample(iterations: int) -> float:
        client.record(e0)
        for _ in range(iterations):
            launch()
        client.record(e1)
        ms = client.elapsed_ms(e0, e1)
        client.sync()
        return ms * 1e-3
    return sample
