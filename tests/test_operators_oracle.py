"""Pins oracle/operators_oracle.py to golden vectors made by the REAL reference
(tests/golden/make_golden_operators.py) and, where /root/reference exists, to the reference classes on fresh random inputs."""
import io
import os
import warnings
from contextlib import redirect_stderr, redirect_stdout

import numpy as np
import pytest

from oracle import operators_oracle as OP
from oracle import reference_harness as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "operators.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN))


NC_CASES = {"nc_default": dict(), "nc_custom": dict(lower_clip_fraction=0.05, upper_clip_fraction=0.02, minval=0, maxval=200),
            "nc_zero_clip": dict(lower_clip_fraction=0.0, upper_clip_fraction=0.0), "nc_not_per_section": dict(per_section=False)}


@pytest.mark.parametrize("tag", sorted(NC_CASES))
def test_normalize_contrast_golden(gold, tag):
    out = OP.normalize_contrast(gold["image"], **NC_CASES[tag])
    assert out.dtype == np.uint8
    np.testing.assert_array_equal(out, gold[tag])


def test_normalize_contrast_quirks(gold):
    # per_section=False is a no-op in the reference (the whole-array branch is the for loop's else clause)
    np.testing.assert_array_equal(gold["nc_not_per_section"], gold["image"])
    # black and constant sections are left alone by the per-section pass, but the trailing whole-array pass still applies
    assert (gold["nc_default"] != gold["image"]).any()


def test_quantize_maskout_crop_golden(gold):
    np.testing.assert_array_equal(OP.quantize(gold["aff"], "xy"), gold["quant_xy"])
    np.testing.assert_array_equal(OP.quantize(gold["aff"], "z"), gold["quant_z"])
    with pytest.raises(ValueError):
        OP.quantize(gold["aff"], "yz")
    np.testing.assert_array_equal(OP.maskout(gold["mask"], (2, 4, 4), gold["aff"], (1, 1, 1)), gold["maskout_aff"])
    np.testing.assert_array_equal(OP.maskout(gold["mask"], (8, 16, 16), gold["image2"], (4, 4, 4)), gold["maskout_img"])
    a3, o3 = OP.crop_margin(gold["aff"], (5, 6, 7), (1, 2, 3))
    a6, o6 = OP.crop_margin(gold["aff"], (5, 6, 7), (1, 0, 3, 2, 4, 0))
    np.testing.assert_array_equal(a3, gold["crop3"]); assert tuple(o3) == tuple(gold["crop3_offset"])
    np.testing.assert_array_equal(a6, gold["crop6"]); assert tuple(o6) == tuple(gold["crop6_offset"])
    with pytest.raises(ValueError):
        OP.crop_margin(gold["aff"], (0, 0, 0), (1, 2))


@pytest.mark.skipif(not H.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_against_real_reference(seed):
    _, Chunk, _ = H.import_reference()
    warnings.simplefilter("ignore")
    from chunkflow.chunk.image.base import Image
    from chunkflow.chunk.affinity_map import AffinityMap
    rng = np.random.default_rng(seed)
    z, y, x = 5, 33, 47
    img = (rng.random((z, y, x)) ** (1 + seed) * rng.integers(60, 256)).astype(np.uint8)
    img[seed % z] //= 8
    lo, hi = [(0.01, 0.01), (0.1, 0.0), (0.0, 0.2)][seed]
    im = Image(img.copy())
    with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
        im.normalize_contrast(lo, hi, 1 + seed, 255 - 10 * seed, True)
    np.testing.assert_array_equal(OP.normalize_contrast(img, lo, hi, 1 + seed, 255 - 10 * seed, True), np.asarray(im.array))
    aff = rng.random((3, 4, 12, 10), dtype=np.float32)
    for mode in ("xy", "z"):
        np.testing.assert_array_equal(OP.quantize(aff, mode), np.asarray(AffinityMap(aff.copy()).quantize(mode).array))
    mask = rng.integers(0, 2, size=(2, 3, 5), dtype=np.uint8)
    c = Chunk(aff.copy(), voxel_size=(4, 4, 4))
    Chunk(mask, voxel_size=(8, 16, 8)).maskout(c)
    np.testing.assert_array_equal(OP.maskout(mask, (8, 16, 8), aff, (4, 4, 4)), np.asarray(c.array))
    for margin in ((1, 2, 3), (0, 1, 2, 1, 0, 3)):
        r = Chunk(aff.copy(), voxel_offset=(3, 2, 1)).crop_margin(margin)
        a, o = OP.crop_margin(aff, (3, 2, 1), margin)
        np.testing.assert_array_equal(a, np.asarray(r.array))
        assert tuple(o) == tuple(r.voxel_offset)


@pytest.mark.skipif(not H.available(), reason="/root/reference not present (GPU box)")
def test_normalize_contrast_property_against_real_reference():
    """Randomised shapes, clip fractions and output ranges (hypothesis, bounded): oracle == real reference, bit for bit,
    including degenerate histograms (empty, single value, value 255 present / absent)."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    H.import_reference()
    warnings.simplefilter("ignore")
    from chunkflow.chunk.image.base import Image

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(st.integers(1, 4), st.integers(1, 24), st.integers(1, 24), st.integers(0, 2 ** 31 - 1),
           st.sampled_from([0.0, 0.01, 0.1, 0.5, 0.9, 1.0]), st.sampled_from([0.0, 0.01, 0.1, 0.5, 1.0]),
           st.integers(0, 40), st.integers(41, 255), st.sampled_from(["uniform", "dark", "top", "const", "zero"]))
    def check(z, y, x, seed, lo, hi, mn, mx, kind):
        rng = np.random.default_rng(seed)
        if kind == "uniform":
            img = rng.integers(0, 256, (z, y, x), dtype=np.uint8)
        elif kind == "dark":
            img = rng.integers(0, 12, (z, y, x), dtype=np.uint8)
        elif kind == "top":
            img = rng.integers(250, 256, (z, y, x), dtype=np.uint8)
        elif kind == "const":
            img = np.full((z, y, x), rng.integers(0, 256), np.uint8)
        else:
            img = np.zeros((z, y, x), np.uint8)
        im = Image(img.copy())
        with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
            im.normalize_contrast(lo, hi, mn, mx, True)
        np.testing.assert_array_equal(OP.normalize_contrast(img, lo, hi, mn, mx, True), np.asarray(im.array))

    check()
