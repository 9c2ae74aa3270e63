"""Host model of csr_midrow_exact_kernel's summation scheme (spmv_openmp_cuda_b200/csrc/kernels.cuh): a lane owns 8 consecutive
entries of every 256-entry chunk that starts at a 4-aligned index, entries outside the row are replaced by +0.0 and ADDED like the
rest, and the running sum goes through the lanes in order.  The kernel's claim is that this reproduces sgemvSerial
(src/SpMV_CSR_OMP.c:229-250) bit for bit because a running sum that starts at +0.0 is never -0.0, so "+ 0.0" is an exact no-op.
This test pins that argument on the CPU (the GPU test of the kernel itself is tests/test_gpu_parity.py::
test_exact_kind_sell_hybrid_on_skewed_rows)."""
import numpy as np
import pytest

C, CH = 8, 256


def serial(vals, xs):
    acc = np.float64(0.0)
    for v, x in zip(vals, xs):
        acc = np.float64(acc + np.float64(v * x))
    return acc


def lane_chain(vals, xs, s):
    """vals/xs: the row's entries; s: the row's start index inside the big array (decides the chunk alignment)."""
    e = s + len(vals)
    acc = np.float64(0.0)
    base = s & ~3
    while base < e:
        nl = (min(e, base + CH) - base + C - 1) // C
        for lane in range(nl):
            t = acc
            for j in range(C):
                i = base + C * lane + j
                p = np.float64(vals[i - s] * xs[i - s]) if s <= i < e else np.float64(0.0)
                t = np.float64(t + p)
            acc = t
        base += CH
    return acc


def bits(a):
    return np.float64(a).view(np.uint64)


@pytest.mark.parametrize("n", [1, 3, 7, 8, 9, 255, 256, 257, 300, 777, 2048])
@pytest.mark.parametrize("s", [0, 1, 2, 3, 5, 1022])
def test_lane_chain_equals_serial_order(n, s):
    rng = np.random.default_rng(1000 * n + s)
    vals = rng.uniform(-1, 1, n) * 10.0 ** rng.integers(-8, 8, n)
    xs = rng.uniform(-1, 1, n)
    with np.errstate(all="ignore"):
        assert bits(lane_chain(vals, xs, s)) == bits(serial(vals, xs))


@pytest.mark.parametrize("vals,xs", [
    ([-0.0, -0.0, -0.0], [1.0, 1.0, 1.0]),             # every product is -0.0: the serial sum is +0.0 (0.0 + -0.0), never -0.0
    ([1.0, -1.0, -0.0], [3.0, 3.0, 5.0]),              # exact cancellation, then a negative zero
    ([1e308, 1e308, -1e308], [10.0, 10.0, 10.0]),      # overflow to inf, inf - inf = nan
    ([5e-324, -5e-324, 5e-324], [1.0, 1.0, 0.5]),      # subnormals, a product that underflows to zero
    ([np.nan, 1.0], [1.0, 1.0]),
])
@pytest.mark.parametrize("s", [0, 1, 2, 3])
def test_lane_chain_special_values(vals, xs, s):
    with np.errstate(all="ignore"):
        a, b = lane_chain(np.array(vals), np.array(xs), s), serial(np.array(vals), np.array(xs))
    assert bits(a) == bits(b) or (np.isnan(a) and np.isnan(b))
    assert not (a == 0 and np.signbit(a))  # the running sum is never -0.0
