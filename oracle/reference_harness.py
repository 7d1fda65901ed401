"""Import the REAL reference hot path from /root/reference (TEST INFRASTRUCTURE).

The reference's ``chunkflow.chunk`` imports file-format packages that are not installed
offline (h5py, tifffile, cc3d, cloudvolume, skimage); none of them is touched by the
inference path, so empty stub modules are registered first.
Only works where /root/reference exists (this container, not the GPU box).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CHUNKFLOW_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "chunkflow"))


def import_reference():
    """Returns (Inferencer, Chunk, PatchMask) of the real reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
        return sys.modules[name]

    dummy = type("Dummy", (), {})
    stub("h5py"); stub("tifffile"); stub("cc3d")
    cv = stub("cloudvolume", CloudVolume=dummy)
    cv.lib = stub("cloudvolume.lib", yellow=lambda s: s, Bbox=dummy, Vec=dummy)
    sk = stub("skimage")
    sk.feature = stub("skimage.feature", match_template=lambda *a, **k: None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from chunkflow.flow.divid_conquer.inferencer import Inferencer
        from chunkflow.chunk import Chunk
        from chunkflow.flow.divid_conquer.patch.patch_mask import PatchMask
    return Inferencer, Chunk, PatchMask
