"""Host logic of the chunked host<->device pipeline (libecc_b200/csrc/eccb200.cu: chunk_bounds) without a GPU: the
boundaries cover the batch exactly, never exceed the stage capacity, ramp up from one wave and down to one wave when
shaping applies, and fall back to equal chunks otherwise."""
import ctypes

import numpy as np

import libecc_b200


def bounds(n, wave, eq, cap, shaped):
    lib = libecc_b200.load_library()
    out = np.zeros(1 << 16, dtype=np.uint32)
    k = lib.eccb200_pipeline_chunk_bounds(n, wave, eq, cap, int(shaped), out.ctypes.data_as(ctypes.c_void_p), out.size)
    assert 1 <= k <= out.size
    return [int(x) for x in out[:k]]


def test_bounds_cover_the_batch_and_respect_capacity():
    wave = 148 * 5 * 128
    eq, cap = 4 * wave, 8 * wave
    for shaped in (False, True):
        for n in (0, 1, 127, wave - 1, wave, wave + 1, 2 * wave, 2 * wave + 1, 3 * wave, 5 * wave + 17, 1 << 20, (1 << 20) + 1,
                  1 << 22, (1 << 24) + 12345, 64 * wave, 64 * wave - 1, 200 * wave + 3):
            b = bounds(n, wave, eq, cap, shaped)
            assert b[0] == 0 and b[-1] == n
            sizes = np.diff(b)
            assert (sizes > 0).all() if n else len(b) == 1
            assert (sizes <= cap).all()
            if not shaped:
                assert (sizes[:-1] == eq).all() and (n == 0 or sizes[-1] <= eq)


def test_shaped_batches_ramp_up_and_down():
    wave = 148 * 4 * 128
    eq, cap = 4 * wave, 8 * wave
    b = np.diff(bounds(1 << 20, wave, eq, cap, True))          # 13.8 waves
    assert b[0] == wave and b[1] == 2 * wave                    # the first kernel starts after a one-wave copy
    assert b[-1] <= wave and b[-2] <= 2 * wave                  # the last device->host copy is at most one wave
    assert max(b) <= 4 * wave                                   # four-wave steady state below 64 waves
    big = np.diff(bounds(1 << 24, wave, eq, cap, True))         # 221 waves: eight-wave steady state
    assert list(big[:4]) == [wave, 2 * wave, 4 * wave, 8 * wave] and max(big) == 8 * wave
    assert big[-1] <= wave and big[-2] <= 2 * wave
    assert (big[4:-3] == 8 * wave).all()
    tiny = np.diff(bounds(2 * wave, wave, eq, cap, True))       # two waves or less: no shaping
    assert list(tiny) == [2 * wave]
