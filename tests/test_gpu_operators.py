"""GPU parity of the operators either side of `inference`: DeviceChunk methods, through the
C ABI, against the numpy oracle and the golden vectors of the REAL reference -- bit-exact (integer / float32 elementwise)."""
import os

import numpy as np
import pytest

from chunkflow_b200 import Chunk
from oracle import operators_oracle as OP

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "operators.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLDEN))


def _dev(arr, **kw):
    from chunkflow_b200.chunk.device import DeviceChunk
    return DeviceChunk.from_chunk(Chunk(arr, **kw))


NC_CASES = {"nc_default": dict(), "nc_custom": dict(lower_clip_fraction=0.05, upper_clip_fraction=0.02, minval=0, maxval=200),
            "nc_zero_clip": dict(lower_clip_fraction=0.0, upper_clip_fraction=0.0), "nc_not_per_section": dict(per_section=False)}


@pytest.mark.parametrize("tag", sorted(NC_CASES))
def test_normalize_contrast_golden(gold, tag):
    d = _dev(gold["image"])
    d.normalize_contrast(**NC_CASES[tag])
    np.testing.assert_array_equal(d.to_chunk().array, gold[tag])


@pytest.mark.parametrize("shape,seed", [((3, 50, 70), 0), ((7, 129, 255), 1), ((2, 16, 16), 2), ((5, 3, 5), 3), ((1, 300, 301), 4)])
def test_normalize_contrast_random_vs_oracle(shape, seed):
    rng = np.random.default_rng(seed)
    img = (rng.random(shape) ** (1 + seed % 3) * rng.integers(40, 256)).astype(np.uint8)
    img[0] //= 4
    if shape[0] > 2:
        img[1] = 0       # black section
        img[2] = 200     # constant section
    lo, hi, mn, mx = [(0.01, 0.01, 1, 255), (0.2, 0.1, 0, 255), (0.0, 0.0, 1, 255), (0.01, 0.3, 10, 100), (0.5, 0.6, 1, 255)][seed]
    d = _dev(img)
    d.normalize_contrast(lo, hi, mn, mx, True)
    np.testing.assert_array_equal(d.to_chunk().array, OP.normalize_contrast(img, lo, hi, mn, mx, True))


def test_normalize_contrast_large_properties():
    """64 x 1024 x 1024: too big for the loop-based oracle in a test; check it against a vectorised restatement of
    the same tables (per-section LUTs composed with the whole-array LUT) built from numpy histograms."""
    import torch
    from chunkflow_b200.chunk.device import DeviceChunk
    g = torch.Generator(device="cuda").manual_seed(5)
    z, y, x = 64, 1024, 1024
    t = (torch.rand((z, y, x), device="cuda", generator=g) ** 2 * 230).to(torch.uint8)
    t[3] = 0
    before = t.clone()
    d = DeviceChunk(t)
    d.normalize_contrast()
    hists = torch.stack([torch.bincount(before[k].flatten().long(), minlength=256) for k in range(z)]).cpu().numpy()
    luts = []
    for k in range(z):
        h = hists[k][:256 if hists[k][255] else 255]
        lut = OP.hist_to_lookup_table(h.copy(), 0.01, 0.01, 1, 255)
        luts.append(np.arange(256, dtype=np.uint8) if lut is None else lut)
    gh = np.zeros(256, np.int64)
    for k in range(z):
        np.add.at(gh, luts[k], hists[k])
    glut = OP.hist_to_lookup_table(gh[:256 if gh[255] else 255].copy(), 0.01, 0.01, 1, 255)
    assert glut is not None
    for k in (0, 3, 17, z - 1):
        expect = torch.from_numpy(glut[luts[k]]).cuda()[before[k].long()]
        assert torch.equal(d.tensor[k], expect), k


def test_quantize_maskout_crop_golden(gold):
    a = _dev(gold["aff"])
    np.testing.assert_array_equal(a.quantize("xy").to_chunk().array, gold["quant_xy"])
    np.testing.assert_array_equal(a.quantize("z").to_chunk().array, gold["quant_z"])
    with pytest.raises(ValueError):
        a.quantize("yz")
    c = _dev(gold["aff"], voxel_size=(1, 1, 1))
    _dev(gold["mask"], voxel_size=(2, 4, 4)).maskout(c)
    np.testing.assert_array_equal(c.to_chunk().array, gold["maskout_aff"])
    c2 = _dev(gold["image2"], voxel_size=(4, 4, 4))
    _dev(gold["mask"], voxel_size=(8, 16, 16)).maskout(c2)
    np.testing.assert_array_equal(c2.to_chunk().array, gold["maskout_img"])
    r3 = _dev(gold["aff"], voxel_offset=(5, 6, 7)).crop_margin((1, 2, 3))
    r6 = _dev(gold["aff"], voxel_offset=(5, 6, 7)).crop_margin((1, 0, 3, 2, 4, 0))
    np.testing.assert_array_equal(r3.to_chunk().array, gold["crop3"]); assert tuple(r3.voxel_offset) == tuple(gold["crop3_offset"])
    np.testing.assert_array_equal(r6.to_chunk().array, gold["crop6"]); assert tuple(r6.voxel_offset) == tuple(gold["crop6_offset"])
    with pytest.raises(ValueError):
        a.crop_margin((1, 2))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_operators_random_vs_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    factor = [(1, 1, 1), (2, 4, 4), (3, 2, 5)][seed]
    msz = [(4, 9, 13), (3, 5, 7), (2, 8, 5)][seed]
    csz = tuple(m * f for m, f in zip(msz, factor))
    aff = (rng.random((3,) + csz, dtype=np.float32) * 1.2 - 0.05).astype(np.float32)
    mask_u8 = rng.integers(0, 3, size=msz, dtype=np.uint8)
    mask_f = rng.random(msz, dtype=np.float32)
    img = rng.integers(0, 256, size=csz, dtype=np.uint8)
    vs = (4, 4, 4)
    mvs = tuple(v * f for v, f in zip(vs, factor))
    for chunk, mask in ((aff, mask_u8), (aff, mask_f), (img, mask_u8), (img, mask_u8.astype(bool))):
        c = _dev(chunk, voxel_size=vs)
        _dev(mask, voxel_size=mvs).maskout(c)
        np.testing.assert_array_equal(c.to_chunk().array, OP.maskout(mask, mvs, chunk, vs))
    # 16-byte aligned uint8 rows: the vector path, with the packed four-voxels-per-multiply form (x factor a multiple of 4, also
    # one that is no power of two) and the per-voxel form (x factor 2, 3), full-range mask values (products wrap modulo 256)
    for fx, mx in ((4, 12), (12, 4), (8, 2), (2, 24), (3, 16)):
        big = rng.integers(0, 256, size=(3, 10, fx * mx), dtype=np.uint8)
        mk = rng.integers(0, 256, size=(3, 5, mx), dtype=np.uint8)
        c = _dev(big, voxel_size=(4, 4, 4))
        _dev(mk, voxel_size=(4, 8, 4 * fx)).maskout(c)
        np.testing.assert_array_equal(c.to_chunk().array, OP.maskout(mk, (4, 8, 4 * fx), big, (4, 4, 4)))
    for mode in ("xy", "z"):
        good = np.clip(aff, 0, 1)  # out-of-range casts are platform defined in numpy: compare where the reference is defined
        np.testing.assert_array_equal(_dev(good).quantize(mode).to_chunk().array, OP.quantize(good, mode))
    for margin in ((0, 0, 0), (1, 2, 1), (0, 1, 2, 1, 0, 3)):
        for arr in (aff, img):
            r = _dev(arr, voxel_offset=(3, 2, 1)).crop_margin(margin)
            a, o = OP.crop_margin(arr, (3, 2, 1), margin)
            np.testing.assert_array_equal(r.to_chunk().array, a)
            assert tuple(r.voxel_offset) == tuple(o)


def test_errors_are_loud():
    from chunkflow_b200 import _native
    img = np.zeros((2, 8, 8), np.uint8)
    with pytest.raises(_native.NativeError):
        _dev(img).normalize_contrast(minval=-1)
    c = _dev(img, voxel_size=(1, 1, 1))
    with pytest.raises(_native.NativeError):  # numpy refuses uint8 *= float32 as well
        _dev(np.ones((2, 8, 8), np.float32), voxel_size=(1, 1, 1)).maskout(c)
    with pytest.raises(_native.NativeError):
        _dev(img).crop_margin((1, 4, 0))


def test_device_resident_pipeline_matches_host_path():
    """normalize-contrast | inference | crop-margin | quantize with the chunk resident in HBM == the same operators run
    one by one on the host path (oracle operators around the product's host inference)."""
    from chunkflow_b200 import Inferencer
    from chunkflow_b200.chunk.device import DeviceChunk
    from conftest import MODEL_FILE
    rng = np.random.default_rng(9)
    img = (rng.random((12, 40, 48)) ** 2 * 200).astype(np.uint8)
    kw = dict(input_patch_size=(8, 32, 32), output_patch_overlap=(2, 8, 8), num_output_channels=3, framework="b200",
              mask_output_chunk=True)
    inf = Inferencer(MODEL_FILE, None, kw.pop("input_patch_size"), **kw)
    d = DeviceChunk.from_chunk(Chunk(img, voxel_offset=(10, 20, 30), voxel_size=(40, 4, 4)))
    d.normalize_contrast()
    aff = inf.infer_device(d)
    thumb = aff.crop_margin((1, 4, 4)).quantize("xy")
    host_img = OP.normalize_contrast(img)
    host_aff = inf(Chunk(host_img, voxel_offset=(10, 20, 30), voxel_size=(40, 4, 4)))
    np.testing.assert_array_equal(d.to_chunk().array, host_img)
    np.testing.assert_allclose(aff.to_chunk().array, host_aff.array, atol=2e-6)
    cropped, off = OP.crop_margin(host_aff.array, host_aff.voxel_offset, (1, 4, 4))
    ref_thumb = OP.quantize(cropped, "xy")
    got = thumb.to_chunk().array
    assert tuple(thumb.voxel_offset) == tuple(off) == (11, 24, 34)
    assert np.abs(got.astype(int) - ref_thumb.astype(int)).max() <= 1  # atomics order: 1 ulp of the sums can flip a truncation


def test_cli_device_resident_chain_equals_host_chain():
    """create-chunk | to-device | normalize-contrast | inference | crop-margin | quantize | to-host  ==  the same chain on
    host chunks (every operator then moves its chunk to the GPU and back)."""
    from click.testing import CliRunner
    from chunkflow_b200.flow import cli
    ops = ["normalize-contrast", "-l", "0.02", "-u", "0.02",
           "inference", "--input-patch-size", "8", "32", "32", "--output-patch-overlap", "2", "8", "8",
           "--num-output-channels", "3", "--framework", "b200", "--batch-size", "4", "--mask-output-chunk",
           "crop-margin", "-m", "1", "4", "4", "1", "4", "4", "quantize", "--mode", "xy"]
    create = ["create-chunk", "--size", "12", "40", "48"]
    r_dev = CliRunner().invoke(cli.main, ["--quiet"] + create + ["to-device"] + ops + ["to-host"], standalone_mode=False)
    r_host = CliRunner().invoke(cli.main, ["--quiet"] + create + ops, standalone_mode=False)
    assert r_dev.exception is None, r_dev.output
    assert r_host.exception is None, r_host.output
    a, b = r_dev.return_value[0]["chunk"], r_host.return_value[0]["chunk"]
    assert a.shape == b.shape == (10, 32, 40) and a.array.dtype == np.uint8
    assert tuple(a.voxel_offset) == tuple(b.voxel_offset) == (1, 4, 4)
    assert np.abs(a.array.astype(int) - b.array.astype(int)).max() <= 1
    assert "normalize-contrast-nkem" in r_dev.return_value[0]["log"]["timer"]


def test_reference_test_maskout_uint32_mask():
    """Port of the reference's tests/chunk/test_chunk.py:65-76 (a uint32 mask at voxel size (2,4,8) on a uint8 image and a
    float32 affinity map at (1,1,1)); the mask is converted on the host by DeviceChunk.mask_from_chunk."""
    from chunkflow_b200.chunk.device import DeviceChunk
    mask = np.ones((8, 16, 4), dtype=np.uint32)
    mask[:4, :8, :2] = 0
    mask = Chunk(mask, voxel_offset=(-2, -3, -4), voxel_size=(2, 4, 8))
    rng = np.random.default_rng(0)
    image = _dev(rng.integers(1, 256, size=(16, 64, 32), dtype=np.uint8), voxel_size=(1, 1, 1))
    before = image.to_chunk().array
    DeviceChunk.mask_from_chunk(mask, like=image).maskout(image)
    got = image.to_chunk().array
    np.testing.assert_array_equal(got[:8, :32, :16], 0)
    np.testing.assert_array_equal(got, OP.maskout(mask.array, (2, 4, 8), before, (1, 1, 1)))
    affs = _dev(rng.random((3, 16, 64, 32), dtype=np.float32) + 0.5, voxel_size=(1, 1, 1))
    before = affs.to_chunk().array
    DeviceChunk.mask_from_chunk(mask, like=affs).maskout(affs)
    got = affs.to_chunk().array
    np.testing.assert_array_equal(got[:, :8, :32, :16], 0)
    np.testing.assert_array_equal(got, OP.maskout(mask.array, (2, 4, 8), before, (1, 1, 1)))
