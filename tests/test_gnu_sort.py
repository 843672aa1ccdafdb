"""csrc/gnu_sort.h restates libstdc++'s std::sort (introsort) so that device code leaves tied entries in the order
the reference's std::sort calls do (src/zoic.cpp:317, :381).  Checked here against the toolchain's
own std::sort, through the host-only test hook of the C ABI, on random, tie-heavy and adversarial inputs."""
import numpy as np
import pytest

from zoic_b200.camera import debug_sort_orders
from zoic_b200.synth import hex_bokeh_image


def _same(values):
    ours, lib = debug_sort_orders(values)
    v = np.asarray(values, np.float32)
    assert np.array_equal(ours, lib)
    assert sorted(ours.tolist()) == list(range(len(v)))            # a permutation
    assert np.all(np.diff(v[ours]) <= 0) or np.isnan(v).any()      # descending


@pytest.mark.parametrize("n", [0, 1, 2, 3, 15, 16, 17, 31, 32, 33, 100, 255, 256, 1000, 5120, 65535])
def test_random_and_tied_values(n):
    rng = np.random.default_rng(n)
    _same(rng.random(n))
    _same(rng.integers(0, 4, n).astype(np.float32))        # heavy ties
    _same(np.zeros(n, np.float32))                         # all tied
    _same(np.arange(n, dtype=np.float32))                  # ascending = worst order for a descending sort
    _same(np.arange(n, dtype=np.float32)[::-1].copy())
    half = rng.random(n).astype(np.float32)
    half[rng.random(n) < 0.6] = 0.0                        # an aperture image row: zeros outside the shape
    _same(half)


def test_rows_of_the_benchmark_image():
    lum = hex_bokeh_image(255) @ np.array([0.3, 0.59, 0.11], np.float32)
    for r in range(0, 255, 7):
        _same(lum[r])
    _same(lum.sum(1))


def _median_of_three_killer(n):
    """Musser's sequence: drives median-of-three quicksort quadratic, so introsort falls back to heapsort."""
    assert n % 2 == 0
    k = n // 2
    a = np.zeros(n, np.float32)
    for i in range(1, k + 1):
        if i % 2 == 1:
            a[i - 1] = i
            a[i] = k + i
        a[k + i - 1] = 2 * i
    return a


@pytest.mark.parametrize("n", [64, 1024, 4096, 20000])
def test_adversarial_inputs_reach_the_heapsort_fallback(n):
    a = _median_of_three_killer(n)
    _same(a)
    _same(-a)
    _same(np.concatenate([a, a]))
    # organ pipe and sawtooth patterns with ties
    x = np.arange(n, dtype=np.float32)
    _same(np.minimum(x, n - 1 - x))
    _same(x % 17)
