"""CPU tests of the counter-based sample stream (host build of the same
rt_sampling.cuh the device compiles) against the reference's own Rng and
CorrelatedMultiJitterSampler (oracle/_ref)."""
import numpy as np
import pytest

from tests.raybatches import bits


def _chunks(width, height):
    cw = width // 4 if width >= 4 else 1
    ch = height // 4 if height >= 4 else 1
    nx = width // cw if width > 4 else 1
    ny = height // ch if height > 4 else 1
    if nx * cw < width:
        nx += 1
    if ny * ch < height:
        ny += 1
    return cw, ch, nx, ny


def _reference_permutations(ref, width, height, depth, x, y):
    """Step the reference Rng literally through every earlier pixel of the chunk
    (RaytraceMain.cpp:69-70, 82-108, 159-169)."""
    cw, ch, nx, ny = _chunks(width, height)
    cx, cy = x // cw, y // ch
    xs, ys = cx * cw, cy * ch
    xe, ye = min(xs + cw, width), min(ys + ch, height)
    z = (((xs << 16) | xe) ^ xs) & 0xffffffff
    w = (((ys << 16) | ye) ^ ys) & 0xffffffff
    k = (y - ys) * (xe - xs) + (x - xs)
    per = 5 * depth + 3
    seq = ref.rng_sequence(z, w, (k + 1) * per)[k * per:]
    out = seq.copy()
    if k == 0:      # construction order: time, lens, subpixel
        pass
    else:           # refill order: lens, time, subpixel
        out[5 * depth + 0], out[5 * depth + 1] = seq[5 * depth + 1], seq[5 * depth + 0]
    return out


@pytest.mark.parametrize("width,height,depth", [(64, 48, 3), (50, 31, 1), (3, 2, 2), (257, 129, 4), (3840, 2160, 3)])
def test_jump_ahead_matches_literal_stepping(capi, ref, width, height, depth):
    rng = np.random.RandomState(width * 7 + height)
    cw, ch, _nx, _ny = _chunks(width, height)
    pixels = {(0, 0), (width - 1, height - 1), (min(cw, width - 1), min(ch, height - 1)), (width - 1, 0)}
    if width > 1:
        pixels.add((1, 0))
    for _ in range(12):
        # keep literal stepping affordable: stay near the start of a chunk's rows
        x = int(rng.randint(0, width))
        y = int(rng.randint(0, height))
        y = (y // ch) * ch + min(y % ch, 3)
        pixels.add((x, min(y, height - 1)))
    _cw, _ch, nx, ny = _chunks(width, height)
    for (x, y) in sorted(pixels):
        if x >= nx * cw or y >= ny * ch:
            # quirk: images narrower than 4 pixels leave pixels outside every chunk unrendered
            with pytest.raises(capi.RtError, match="never rendered"):
                capi.sample_permutations(width, height, depth, x, y)
            continue
        mine = capi.sample_permutations(width, height, depth, x, y)
        want = _reference_permutations(ref, width, height, depth, x, y)
        assert np.array_equal(mine, want), (x, y)


def test_cmj_matches_reference(capi, ref):
    import ctypes as C
    lib = capi.core()
    rng = np.random.RandomState(3)
    for samples in (1, 2, 3, 16, 255, 256, 1024, 65536):
        for perm in rng.randint(0, 2 ** 32, size=3, dtype=np.uint64):
            perm = int(perm)
            count = min(samples, 300)
            want = ref.cmj_1d(samples, perm, count)
            mine = np.array([lib.rt_cmj_sample1d(i, samples, perm) for i in range(count)], np.float32)
            assert np.array_equal(bits(mine), bits(want)), (samples, perm)
    for xs, ys in ((1, 1), (2, 2), (3, 5), (16, 16), (32, 32), (7, 1)):
        for perm in rng.randint(0, 2 ** 32, size=3, dtype=np.uint64):
            perm = int(perm)
            count = min(xs * ys, 300)
            want = ref.cmj_2d(xs, ys, perm, count)
            u, v = C.c_float(), C.c_float()
            mine = np.zeros((count, 2), np.float32)
            for i in range(count):
                lib.rt_cmj_sample2d(i, xs, ys, perm, C.byref(u), C.byref(v))
                mine[i] = (u.value, v.value)
            assert np.array_equal(bits(mine), bits(want)), (xs, ys, perm)
            assert (mine >= 0).all() and (mine < 1).all()
