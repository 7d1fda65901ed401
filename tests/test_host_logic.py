"""Host logic that needs no GPU: boundary types, the native library's exports, the
device-free patch mask, and loud failure without a CUDA device."""
import hashlib
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from chunkflow_b200 import Cartesian, Chunk, to_cartesian
from chunkflow_b200 import _native
from oracle import inferencer_oracle as O


def test_cartesian_arithmetic_and_partial_order():
    a = Cartesian(4, 64, 64)
    assert a // 2 == Cartesian(2, 32, 32) and a * 2 == (8, 128, 128) and a - 1 == (3, 63, 63)
    assert (Cartesian(20, 256, 256) - Cartesian(16, 192, 192)) // 2 == (2, 32, 32)
    assert Cartesian(4, 64, 64) >= Cartesian(4, 64, 64)
    assert not (Cartesian(4, 64, 1) >= Cartesian(4, 64, 64))      # every axis must hold
    assert not (Cartesian(4, 64, 1) < Cartesian(4, 64, 64))       # ... so neither >= nor <
    assert to_cartesian(None) is None and to_cartesian([1, 2, 3]) == Cartesian(1, 2, 3)
    with pytest.raises((AssertionError, ValueError)):
        to_cartesian((1, 2))


def test_chunk_blend_clips_like_reference():
    buf = Chunk(np.zeros((2, 6, 8, 8), np.float32), voxel_offset=(10, 20, 30))
    patch = Chunk(np.ones((2, 4, 4, 4), np.float32), voxel_offset=(14, 26, 28))
    buf.blend(patch)
    ref = np.zeros((2, 6, 8, 8), np.float32)
    O._blend(ref, (10, 20, 30), patch.array, (14, 26, 28))
    assert np.array_equal(buf.array, ref) and buf.array.sum() == 2 * 2 * 2 * 2


def test_chunk_cutout_ufunc_and_myelin():
    c = Chunk(np.arange(4 * 5 * 6, dtype=np.float32).reshape(4, 5, 6), voxel_offset=(1, 2, 3), voxel_size=(40, 4, 4))
    sub = c.cutout((slice(2, 4), slice(3, 5), slice(4, 8)))
    assert sub.shape == (2, 2, 4) and tuple(sub.voxel_offset) == (2, 3, 4) and sub.array[0, 0, 0] == c.array[1, 1, 1]
    with pytest.raises(IndexError):
        c.cutout((slice(0, 2), slice(3, 5), slice(4, 8)))
    out = Chunk(np.ones((3, 4, 5, 6), np.float32), voxel_offset=(1, 2, 3), voxel_size=(40, 4, 4))
    out *= Chunk(np.full((4, 5, 6), 0.5, np.float32), voxel_offset=(1, 2, 3))
    assert isinstance(out, Chunk) and tuple(out.voxel_size) == (40, 4, 4) and np.all(out.array == 0.5)
    aff = Chunk(np.stack([np.ones((2, 2, 2), np.float32)] * 3 + [np.array([[[0.1, 0.5]] * 2] * 2, np.float32)]))
    masked = aff.mask_using_last_channel(threshold=0.3)
    assert masked.shape == (3, 2, 2, 2) and np.array_equal(masked.array[0, 0, 0], [1, 0])


def test_chunk_create_sin_matches_reference_formula():
    c = Chunk.create(size=(6, 10, 12), dtype=np.uint8, pattern="sin")
    iz, iy, ix = np.meshgrid(*[np.linspace(0, 1, n) for n in (6, 10, 12)], indexing="ij")
    assert np.array_equal(c.array, (np.abs(np.sin(4 * (iz + iy + ix))) * 255).astype(np.uint8))
    assert c.layer_type == "image"


def test_native_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "chunkflow_b200.h")).read()
    declared = set(re.findall(r"\b(cfb_[a-z0-9_]+)\s*\(", header)) - {"cfb_engine"}
    lib = _native.load()
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/chunkflow_b200.h but not exported"
    assert declared == set(_native.EXPORTS)
    assert lib.cfb_version() >= 100


@pytest.mark.parametrize("ps,ov", [((20, 256, 256), (4, 64, 64)), ((32, 256, 256), (8, 64, 64)), ((8, 32, 32), (2, 8, 8))])
def test_native_patch_mask_is_bit_identical_to_reference(geometry, ps, ov):
    m = _native.make_patch_mask(ps, ov)
    g = geometry["patch_masks"]["x".join(map(str, ps)) + "_" + "x".join(map(str, ov))]
    assert hashlib.sha256(m.tobytes()).hexdigest() == g["sha256"]


def test_patch_mask_partition_of_unity():
    ps, ov = (8, 32, 32), (2, 8, 8)
    m = _native.make_patch_mask(ps, ov).astype(np.float64)
    st = tuple(p - o for p, o in zip(ps, ov))
    acc = np.zeros(tuple(p + 2 * s for p, s in zip(ps, st)))
    for a in range(3):
        for b in range(3):
            for c in range(3):
                acc[a * st[0]:a * st[0] + ps[0], b * st[1]:b * st[1] + ps[1], c * st[2]:c * st[2] + ps[2]] += m
    centre = acc[st[0]:st[0] + ps[0], st[1]:st[1] + ps[1], st[2]:st[2] + ps[2]]
    np.testing.assert_allclose(centre, 1.0, atol=1e-6)


def test_product_tta_modes_match_the_oracle():
    """Host plug-in augmentation: 'reference' (default) is the reference's literal arithmetic -- flips on the channel /
    batch axes, inverse steps in forward order (transform.py:30-52,147-156), pinned to the real reference through the
    oracle in tests/test_oracle_vs_reference.py; 'spatial' is the explicit opt-in."""
    from chunkflow_b200.flow.divid_conquer.transform import TransformSequences
    from oracle import inferencer_oracle as O
    rng = np.random.default_rng(3)
    a = rng.random((2, 3, 3, 8, 8)).astype(np.float32)
    lit = TransformSequences()
    assert lit.mode == "reference"
    fw = lit.forward(a)
    assert len(fw) == 8 and all(np.array_equal(x, y) for x, y in zip(fw, O.tta_forward(a)))
    assert all(np.array_equal(x, y) for x, y in zip(lit.backward(fw), O.tta_backward(O.tta_forward(a))))
    assert all(np.array_equal(x, a) for x in lit.backward(fw))
    sp = TransformSequences("spatial")
    fw = sp.forward(a)
    assert len({x.tobytes() for x in fw}) == 8 and all(np.array_equal(x, a) for x in sp.backward(fw))
    with pytest.raises(ValueError):
        TransformSequences("rotate")
    assert _native.augment_code(True) == _native.AUGMENT_REFERENCE and _native.augment_code("spatial") == _native.AUGMENT_SPATIAL
    assert _native.augment_code(False) == _native.AUGMENT_NONE


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu():
    from chunkflow_b200 import Inferencer
    with pytest.raises(_native.NativeError) as ei:
        Inferencer(None, None, (8, 32, 32), output_patch_overlap=(2, 8, 8), framework="identity")
    assert ei.value.code == _native.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "chunkflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_cli_accepts_every_reference_inference_flag():
    """The `inference` command keeps the reference's option surface (flow/flow.py:1853-1893)."""
    from chunkflow_b200.flow.cli import inference, create_chunk
    opts = {o for p in inference.params for o in p.opts + p.secondary_opts}
    for flag in ["--name", "--convnet-model", "-m", "--convnet-weight-path", "-w", "--input-patch-size", "-s",
                 "--output-patch-size", "-z", "--output-patch-overlap", "-v", "--output-crop-margin", "--patch-num", "-n",
                 "--num-input-channels", "--num-output-channels", "-c", "--dtype", "-d", "--framework", "-f",
                 "--batch-size", "-b", "--bump", "--mask-output-chunk", "--no-mask-output-chunk",
                 "--mask-myelin-threshold", "-y", "--augment", "--no-augment", "--input-chunk-name", "-i",
                 "--output-chunk-name", "-o"]:
        assert flag in opts, flag
    fw = next(p for p in inference.params if p.name == "framework")
    assert {"universal", "identity", "pytorch", "b200"} <= set(fw.type.choices)
    assert next(p for p in inference.params if p.name == "mask_output_chunk").default is False   # CLI default (ctor: True)
    assert next(p for p in inference.params if p.name == "output_patch_overlap").default == (4, 64, 64)
    assert "--size" in {o for p in create_chunk.params for o in p.opts}


def test_cli_neighbour_operators_keep_the_reference_flags():
    """normalize-contrast / crop-margin / quantize keep the reference's option surface
    (flow/flow.py:1672-1687, 2053-2065, 2250-2256); to-device / to-host are the only additions."""
    from chunkflow_b200.flow import cli

    def opts(cmd):
        return {o for p in cmd.params for o in p.opts + p.secondary_opts}

    nc = opts(cli.normalize_contrast)
    for flag in ["--name", "--input-chunk-name", "-i", "--output-chunk-name", "-o", "--lower-clip-fraction", "-l",
                 "--upper-clip-fraction", "-u", "--minval", "--maxval", "--per-section", "--whole"]:
        assert flag in nc, flag
    defaults = {p.name: p.default for p in cli.normalize_contrast.params}
    assert defaults["lower_clip_fraction"] == 0.01 and defaults["upper_clip_fraction"] == 0.01
    assert defaults["minval"] == 1 and defaults["maxval"] == 255 and defaults["per_section"] is True
    assert defaults["name"] == "normalize-contrast-nkem"
    for flag in ["--name", "--margin-size", "-m", "--input-chunk-name", "-i", "--output-chunk-name", "-o"]:
        assert flag in opts(cli.crop_margin), flag
    assert next(p for p in cli.crop_margin.params if p.name == "margin_size").nargs == 6
    q = opts(cli.quantize)
    for flag in ["--input-chunk-name", "-i", "--output-chunk-name", "-o", "--mode"]:
        assert flag in q, flag
    assert set(next(p for p in cli.quantize.params if p.name == "mode").type.choices) == {"xy", "z"}
    assert {"to-device", "to-host", "normalize-contrast", "crop-margin", "quantize", "inference", "create-chunk"} <= set(cli.main.commands)


def test_device_chunk_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without CUDA the device-resident operators refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only box")
    from chunkflow_b200 import Chunk
    from chunkflow_b200.chunk.device import DeviceChunk
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceChunk.from_chunk(Chunk(np.zeros((2, 4, 4), np.uint8)))


def test_mask_conversion_preserves_numpy_semantics():
    """chunkflow_b200.chunk.device.mask_array_for: masks of any dtype numpy accepts in `chunk *= mask` map onto the two
    kernel mask dtypes without changing the result (reference tests/chunk/test_chunk.py:65-76 uses a uint32 mask)."""
    from chunkflow_b200.chunk.device import mask_array_for
    from oracle import operators_oracle as OP
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(4, 8, 12), dtype=np.uint8)
    aff = rng.standard_normal((3, 4, 8, 12)).astype(np.float32)
    for mdt, hi in ((np.uint32, 1000), (np.uint16, 700), (np.uint64, 5), (bool, 2), (np.uint8, 256)):
        mask = rng.integers(0, hi, size=(2, 4, 3)).astype(mdt)
        conv = mask_array_for(np.uint8, mask)
        assert conv.dtype in (np.uint8, np.bool_)
        np.testing.assert_array_equal(OP.maskout(conv, (2, 2, 4), img, (1, 1, 1)), OP.maskout(mask, (2, 2, 4), img, (1, 1, 1)))
    for mdt, lo, hi in ((np.uint32, 0, 2 ** 24), (np.int32, -1000, 1000), (np.int64, -3, 3), (np.uint16, 0, 65536)):
        mask = rng.integers(lo, hi, size=(2, 4, 3)).astype(mdt)
        conv = mask_array_for(np.float32, mask)
        assert conv.dtype == np.float32
        np.testing.assert_array_equal(OP.maskout(conv, (2, 2, 4), aff, (1, 1, 1)), OP.maskout(mask, (2, 2, 4), aff, (1, 1, 1)))
    for bad_chunk, bad_mask in ((np.uint8, np.ones((1, 1, 1), np.int32)), (np.uint8, np.ones((1, 1, 1), np.float32)),
                                (np.float32, np.ones((1, 1, 1), np.float64)), (np.float32, np.full((1, 1, 1), 2 ** 24, np.int64))):
        with pytest.raises(TypeError):
            mask_array_for(bad_chunk, bad_mask)
    # numpy itself refuses the first two of those
    with pytest.raises(TypeError):
        a = np.ones(3, np.uint8); a *= np.ones(3, np.int32)


def test_result_array_is_recycled_only_when_the_caller_dropped_it():
    """Inferencer._result_array: the previous result array is handed out again only when nothing references it any more
    (a view of it counts); see the method's docstring for why (12.9 GB results, page faults and munmap inside a VM)."""
    from chunkflow_b200.flow.divid_conquer.inferencer import Inferencer

    class Holder:
        pass
    h = Holder()
    shape = (3, 64, 512, 512)     # 201 MB: above the 64 MB threshold
    a = Inferencer._result_array(h, shape)
    addr = a.ctypes.data
    view = a[:-1]
    del a
    b = Inferencer._result_array(h, shape)
    assert b.ctypes.data != addr, "a result that is still referenced (through a view) must not be recycled"
    addr_b = b.ctypes.data
    del view, b
    c = Inferencer._result_array(h, shape)
    assert c.ctypes.data == addr_b, "a dropped result array is handed out again"
    d = Inferencer._result_array(h, (3, 8, 8, 8))
    assert d.shape == (3, 8, 8, 8) and getattr(h, "_last_result", None) is None or h._last_result is not d   # small results are not kept
