"""K5 — zoom-video in-betweening (reference video.cpp:14-34) against its CPU restatement, bit for bit.
The resampling primitives (ByteImage::scaled / blit / blend) live in libbyteimage, which the reference
does not vendor, so the conventions are this repo's and parity with the reference is unpinned
(DESIGN.md); the test pins kernel == restatement and the structural properties of nextFrame."""
import ctypes as C

import numpy as np
import pytest

import newman_b200
import oracles


def oracle_inbetween(prev, nxt, nr, nc, rate):
    P = oracles.oraclep()
    P.oraclep_video_inbetween.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                          C.c_void_p]
    out = np.zeros((rate, nr, nc, 3), dtype=np.uint8)
    v = np.float32(1.5 ** (1.0 / rate))
    P.oraclep_video_inbetween(oracles.vp(prev), oracles.vp(nxt), prev.shape[0], prev.shape[1], nr, nc, rate, v, oracles.vp(out))
    return out


def test_oracle_inbetween_structure():
    """Frame 0 shows the previous key frame at scale 1 (centre crop, no blend: t = 0); a constant image
    stays constant; frames move monotonically towards the new key frame."""
    rng = np.random.default_rng(3)
    nr, nc, rate = 40, 64, 9
    H, W = nr * 3 // 2, nc * 3 // 2
    prev = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    nxt = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    out = oracle_inbetween(prev, nxt, nr, nc, rate)
    r0, c0 = (H - nr) // 2, (W - nc) // 2
    assert np.array_equal(out[0], prev[r0:r0 + nr, c0:c0 + nc])
    flat = oracle_inbetween(np.full((H, W, 3), 77, np.uint8), np.full((H, W, 3), 77, np.uint8), nr, nc, rate)
    assert (np.abs(flat.astype(int) - 77) <= 1).all()
    a = oracle_inbetween(np.zeros((H, W, 3), np.uint8), np.full((H, W, 3), 200, np.uint8), nr, nc, rate)
    means = a.reshape(rate, -1).mean(axis=1)
    assert means[0] == 0 and (np.diff(means) > 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("nr,nc,rate", [(40, 64, 9), (90, 160, 45), (33, 37, 5)])
def test_gpu_inbetween_equals_restatement(dev, nr, nc, rate):
    rng = np.random.default_rng(nr)
    H, W = nr * 3 // 2, nc * 3 // 2
    prev = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    nxt = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    got = dev.video_inbetween(prev, nxt, nr, nc, rate)
    assert np.array_equal(got, oracle_inbetween(prev, nxt, nr, nc, rate))


@pytest.mark.gpu
def test_gpu_video_key_frames_to_inbetweens(dev):
    """Two consecutive key frames of a zoom (1.5x apart, viewer.cpp:283) rendered, coloured at 1.5x the
    video size (viewer.cpp:271-274) and in-betweened: the path video.cpp drives, end to end on the GPU."""
    nr, nc = 48, 64
    H, W = nr * 3 // 2, nc * 3 // 2
    pal = newman_b200.MultiWaveGenerator(oracles.ROOT + "/tests/golden/default.pal").cache(256)
    keys = []
    m = newman_b200.Mandelbrot(H * 2, W * 2, N=256)        # sc = 2 for the recolour, like the reference
    for _ in range(2):
        m.precompute()
        keys.append(m.resolve(pal, sc=2, smooth=True))
        m.zoom(1.5)
    frames = dev.video_inbetween(keys[0], keys[1], nr, nc, 45)
    assert frames.shape == (45, nr, nc, 3)
    assert np.array_equal(frames, oracle_inbetween(keys[0], keys[1], nr, nc, 45))
    assert frames.std() > 5
