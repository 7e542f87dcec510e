"""Device primitives of the broad phase: radix sort (K2) and exclusive scan, through the C ABI test hooks."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lib():
    from gears_b200 import capi
    return capi.lib()


@pytest.mark.parametrize("n", [0, 1, 5, 255, 4096, 4097, 100_003, 3_000_001])
@pytest.mark.parametrize("bits", [1, 7, 8, 13, 26, 32])
def test_radix_sort_u32_stable(n, bits):
    rng = np.random.default_rng(n * 131 + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, hi + 1, size=n, dtype=np.uint64).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    k, v = keys.copy(), vals.copy()
    assert _lib().gp_test_radix_sort(0, n, bits, k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)) == 0
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])  # stability: equal keys keep input order


def test_radix_sort_u32_iota_values():
    n = 777_777
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 1 << 20, size=n).astype(np.uint32)
    k = keys.copy()
    assert _lib().gp_test_radix_sort(0, n, 20, k.ctypes.data_as(C.c_void_p), None) == 0
    assert np.array_equal(k, np.sort(keys))


@pytest.mark.parametrize("n", [1, 4095, 250_000])
def test_radix_sort_u64(n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
    vals = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    k, v = keys.copy(), vals.copy()
    assert _lib().gp_test_radix_sort64(0, n, 63, k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p)) == 0
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])


@pytest.mark.parametrize("n", [0, 1, 4096, 4097, 1_000_003, 16_777_217])
def test_exclusive_scan(n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 50, size=n).astype(np.uint32)
    d = a.copy()
    tot = C.c_uint32()
    assert _lib().gp_test_exclusive_scan(0, n, d.ctypes.data_as(C.c_void_p), C.byref(tot)) == 0
    ref = np.concatenate([[0], np.cumsum(a, dtype=np.uint64)]).astype(np.uint32)
    assert np.array_equal(d, ref[:-1])
    assert tot.value == int(ref[-1])
