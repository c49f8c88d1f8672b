"""CPU test: the device's sinf / cosf / powf (rayito_b200/csrc/rt_libm.cuh, compiled
for the host from the same source) must equal the C library's -- the reference's
third-party arithmetic (glibc libm) -- bit for bit on the ranges the render path uses."""
import ctypes as C
import ctypes.util

import numpy as np
import pytest

from tests.raybatches import bits


@pytest.fixture(scope="module")
def libm():
    lib = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    for name in ("sinf", "cosf"):
        getattr(lib, name).restype = C.c_float
        getattr(lib, name).argtypes = [C.c_float]
    lib.powf.restype = C.c_float
    lib.powf.argtypes = [C.c_float, C.c_float]
    return lib


def _libm_map(fn, *arrays):
    return np.array([fn(*[float(v) for v in vals]) for vals in zip(*arrays)], np.float32)


def test_sinf_cosf_bit_exact(capi, libm):
    rng = np.random.RandomState(1)
    # phi = 2*pi*u with u in [0,1): every angle the samplers produce, plus small and larger arguments
    u = rng.uniform(0, 1, 400000).astype(np.float32)
    phi = (np.float64(2 * np.pi) * u.astype(np.float64)).astype(np.float32)
    extra = np.concatenate([rng.uniform(-119, 119, 100000), rng.uniform(-1e-3, 1e-3, 20000),
                            [0.0, -0.0, 1e-5, 0.78539816, 0.7853982, 1.5707964, 3.1415927, 6.2831855, 100.0]]).astype(np.float32)
    x = np.concatenate([phi, extra])
    for kind, fn in ((0, libm.sinf), (1, libm.cosf), (3, libm.sinf), (4, libm.cosf)):
        mine = capi.libm_eval(kind, x)
        want = _libm_map(fn, x)
        bad = np.flatnonzero(bits(mine) != bits(want))
        assert bad.size == 0, (kind, bad.size, x[bad[:5]], mine[bad[:5]], want[bad[:5]])


def test_powf_bit_exact(capi, libm):
    rng = np.random.RandomState(2)
    n = 300000
    # Glossy: pow(|n.h|, e) and pow(1 - u, 1/(e+1)) with e = 1/roughness^2; displayImage: pow(c, 1/2.2)
    base = np.concatenate([rng.uniform(0, 1, n), rng.uniform(0, 4, n // 3), [1.0, 0.5, 1e-30, 0.99999994]]).astype(np.float32)
    base = base[base > 0]
    exps = rng.choice(np.array([100.0, 11.111111, 1 / 101.0, 1 / 12.111111, 1 / 2.2, 2.0, 0.3, 16.0], np.float32),
                      size=base.size).astype(np.float32)
    mine = capi.libm_eval(2, base, exps)
    want = _libm_map(libm.powf, base, exps)
    bad = np.flatnonzero(bits(mine) != bits(want))
    assert bad.size == 0, (bad.size, base[bad[:5]], exps[bad[:5]], mine[bad[:5]], want[bad[:5]])
    # special cases fall back to double pow and still agree on the easy ones
    assert capi.libm_eval(2, [0.0, 1.0, 2.0], [2.0, 0.0, 0.0]).tolist() == [0.0, 1.0, 1.0]
