"""bench.py synthesises its frames with a faster copy of oracle.synth.image_model: same pixels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import synth


def test_frame_image_equals_the_oracle_model():
    import bench
    for w, h, seed in ((300, 50, 12345), (1024, 77, 12399), (8256, 16, 12345)):
        assert np.array_equal(bench.frame_image(w, h, seed), synth.image_model(w, h, seed))


def test_checksum_weights_are_odd_and_16_bit():
    import bench
    w = bench._weights(64, 32)
    assert w.shape == (32, 64) and w.dtype == np.uint64 and int(w.max()) <= 0xFFFF and bool((w & 1).all())


def test_c_generator_equals_the_oracle_model_and_its_checksums():
    import bench
    for w, h, seed in ((300, 50, 12345), (1024, 77, 12399)):
        img, s0, s1 = synth.image_model_c(w, h, seed)
        want = synth.image_model(w, h, seed)
        assert np.array_equal(img, want)
        v = want.astype(np.uint64)
        assert s0 == int(v.sum(dtype=np.uint64))
        with np.errstate(over="ignore"):
            assert s1 == int((v * bench._weights(w, h)).sum(dtype=np.uint64))
