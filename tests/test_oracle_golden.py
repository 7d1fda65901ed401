"""The CPU oracle against the golden vectors the REAL reference produced
(tests/golden/make_golden.py) and the reference's own known-answer tests."""
import hashlib

import numpy as np
import pytest

from oracle import inferencer_oracle as O


def _key(*triples):
    return "_".join("x".join(map(str, t)) for t in triples)


@pytest.mark.parametrize("ps,ov", [((20, 256, 256), (4, 64, 64)), ((10, 128, 128), (2, 32, 32)), ((8, 32, 32), (2, 8, 8))])
def test_patch_mask_matches_reference(geometry, ps, ov):
    m = O.make_patch_mask(ps, ov)
    g = geometry["patch_masks"][_key(ps, ov)]
    assert hashlib.sha256(m.tobytes()).hexdigest() == g["sha256"]
    assert m.dtype == np.float32 and float(m.max()) == 1.0
    # interior is exactly 1 (reference patch_mask.py:43-46)
    assert np.all(m[ov[0]:-ov[0], ov[1]:-ov[1], ov[2]:-ov[2]] == 1)


def test_patch_mask_survey_known_answers():
    # SURVEY.md section 8c, derived from the reference code
    m = O.make_patch_mask((20, 256, 256), (4, 64, 64))
    assert hashlib.sha256(m.tobytes()).hexdigest()[:16] == "1bb1ccc24aace4c2"
    np.testing.assert_allclose(m[0, 0, 0], 3.56938381e-06, rtol=1e-7)
    np.testing.assert_allclose(m[0, 128, 128], 0.0200143699, rtol=1e-7)
    np.testing.assert_allclose(m.sum(dtype=np.float64), 589824, rtol=1e-7)


def test_patch_grid_matches_reference(geometry):
    for key, starts in geometry["patch_grids"].items():
        size, patch, ov = (tuple(map(int, t.split("x"))) for t in key.split("_"))
        geom = O.Geometry(patch, None, ov, None, True)
        got = [[s.start for s in pair[0]] for pair in O.patch_slices_list(geom, size)]
        assert got == starts, key


def test_identity_nonaligned_bit_exact(golden):
    g = golden("identity_nonaligned.npz")
    out, off = O.infer_chunk(g["input"], tuple(g["voxel_offset"]), input_patch_size=(8, 32, 32),
                             output_patch_overlap=(2, 8, 8), num_output_channels=2, framework="identity")
    assert np.array_equal(out, g["output"])
    # the reference's own assertion (tests/flow/divid_conquer/test_inferencer.py:141-169)
    np.testing.assert_allclose(g["input"].astype(np.float32) / 255, out[0], rtol=1e-5, atol=1e-5)


def test_identity_aligned_bit_exact(golden):
    g = golden("identity_aligned.npz")
    out, off = O.infer_chunk(g["input"], input_patch_size=(8, 32, 32), output_patch_overlap=(2, 8, 8),
                             num_output_channels=2, framework="identity", mask_output_chunk=False)
    assert off == tuple(g["voxel_offset"]) == (2, 8, 8)
    assert np.array_equal(out, g["output"])
    crop = g["input"][2:-2, 8:-8, 8:-8].astype(np.float32) / 255
    np.testing.assert_allclose(crop, out[0], rtol=1e-3, atol=1e-3)


def test_unet3l_matches_reference_pytorch_path(golden, geometry, unet_model):
    import torch
    g = golden("unet3l_small.npz")
    out, _ = O.infer_chunk(g["input"], input_patch_size=(8, 32, 32), output_patch_overlap=(2, 8, 8),
                           num_output_channels=3, framework="pytorch", model=unet_model)
    if torch.__version__ == geometry["torch"]:
        np.testing.assert_allclose(out, g["output"], rtol=0, atol=1e-6)
    else:  # another torch build may order the fp32 sums differently
        np.testing.assert_allclose(out, g["output"], rtol=0, atol=1e-4)
    assert out.min() > 0 and out.max() < 1 and out.std() > 0.1  # outputs span (0,1): a real discriminator


def test_seeded_weights_are_reproducible(geometry, unet_model):
    import torch
    if torch.__version__ != geometry["torch"]:
        pytest.skip("different torch build")
    h = hashlib.sha256()
    sd = unet_model.state_dict()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].numpy().astype(np.float32)).tobytes())
    assert h.hexdigest() == geometry["unet3l_state_sha256"]


def test_all_zero_shortcut():
    out, _ = O.infer_chunk(np.zeros((10, 40, 40), np.uint8), input_patch_size=(8, 32, 32), output_patch_overlap=(2, 8, 8),
                           num_output_channels=3, framework="identity")
    assert out.shape == (3, 10, 40, 40) and not out.any()


def test_tta_roundtrip_is_identity():
    rng = np.random.default_rng(0)
    a = rng.random((1, 1, 4, 16, 16)).astype(np.float32)
    back = O.tta_backward(O.tta_forward(a))
    for b in back:
        assert np.array_equal(a, b)
