"""`connected-components`: oracle self-consistency on CPU, CUDA kernels against the oracle on the GPU.
cc3d is absent from the reference tree and from this image: the oracle restates its published behaviour ("parity unpinned",
oracle/segmentation_oracle.py)."""
import numpy as np
import pytest

from oracle import segmentation_oracle as S


def _case(seed, shape, density=0.55, nvalues=1):
    rng = np.random.default_rng(seed)
    a = (rng.random(shape) > density).astype(np.uint8)
    if nvalues > 1:
        a = a * rng.integers(1, nvalues + 1, size=shape).astype(np.uint8)
    return a


@pytest.mark.parametrize("conn", [6, 18, 26])
def test_oracle_scipy_matches_pure_python_union_find(conn):
    for seed, nv in ((0, 1), (1, 3)):
        a = _case(seed, (5, 7, 9), nvalues=nv)
        assert np.array_equal(S.connected_components(a, conn), S.connected_components_slow(a, conn))


def test_oracle_known_answers():
    a = np.zeros((1, 3, 5), np.uint8)
    a[0, 0, :2] = 1; a[0, 2, 3:] = 1; a[0, 1, 2] = 1      # diagonal bridge: joined only with 18 / 26 connectivity
    assert S.connected_components(a, 6).max() == 3 and S.connected_components(a, 18).max() == 1
    lab = S.connected_components(a, 6)
    assert lab[0, 0, 0] == 1 and lab[0, 1, 2] == 2 and lab[0, 2, 4] == 3     # numbered in raster order of the first voxel
    b = np.array([[[1, 2, 2, 0, 2]]], np.uint8)
    assert S.connected_components(b, 6).tolist() == [[[1, 2, 2, 0, 3]]]       # equal values connect, different values do not
    assert S.threshold(np.array([[[[0.2, 0.7]]]], np.float32), 0.5).tolist() == [[[0, 1]]]


@pytest.mark.gpu
@pytest.mark.parametrize("conn", [6, 18, 26])
def test_device_connected_components_against_oracle(conn):
    import torch
    from chunkflow_b200 import Chunk
    from chunkflow_b200.chunk.device import DeviceChunk
    for seed, shape, dens, nv in ((3, (17, 40, 53), 0.55, 1), (4, (9, 33, 64), 0.35, 4), (5, (1, 1, 7), 0.5, 1), (6, (40, 70, 90), 0.75, 2)):
        a = _case(seed, shape, dens, nv)
        ref = S.connected_components(a, conn)
        dev = DeviceChunk(torch.from_numpy(a).cuda(), voxel_offset=(1, 2, 3), voxel_size=(4, 4, 4))
        out = dev.connected_component(connectivity=conn)
        got = out.tensor.cpu().numpy().view(np.uint32)
        assert out.num_components == ref.max() and tuple(out.voxel_offset) == (1, 2, 3)
        assert np.array_equal(got, ref), (seed, conn)
    # a float32 affinity-like map with a threshold, through the host Chunk API (reference chunk/base.py:128-137)
    rng = np.random.default_rng(9)
    m = rng.random((1, 12, 30, 31)).astype(np.float32)
    got = Chunk(m).connected_component(threshold=0.6, connectivity=conn)
    assert np.array_equal(np.asarray(got.array).view(np.uint32), S.chunk_connected_component(m, 0.6, conn))
    # all background / one solid block
    z = DeviceChunk(torch.zeros((4, 8, 8), dtype=torch.uint8, device="cuda")).connected_component(connectivity=conn)
    assert z.num_components == 0 and not z.tensor.cpu().numpy().any()
    o = DeviceChunk(torch.ones((4, 8, 8), dtype=torch.uint8, device="cuda")).connected_component(connectivity=conn)
    assert o.num_components == 1 and (o.tensor.cpu().numpy().view(np.uint32) == 1).all()


@pytest.mark.gpu
def test_connected_components_cli_and_large_volume_properties():
    import torch
    from click.testing import CliRunner
    from chunkflow_b200.chunk.device import DeviceChunk
    from chunkflow_b200.flow import cli
    res = CliRunner().invoke(cli.main, ["create-chunk", "--size", "16", "64", "64", "--dtype", "float32", "--pattern", "sin",
                                        "connected-components", "--threshold", "0.5", "--connectivity", "26"], standalone_mode=False)
    assert res.exception is None, res.output
    seg = res.return_value[0]["chunk"]
    assert seg.shape == (16, 64, 64) and seg.array.max() >= 1
    # size-independent properties on a volume the scipy oracle would take long for: labels are 1..N without gaps, every
    # component is value-uniform, relabelling the result is idempotent
    rng = np.random.default_rng(11)
    a = (rng.random((128, 256, 256)) > 0.6).astype(np.uint8)
    dev = DeviceChunk(torch.from_numpy(a).cuda())
    out = dev.connected_component(connectivity=6)
    lab = out.tensor.cpu().numpy().view(np.uint32)
    n = out.num_components
    assert lab.max() == n and np.array_equal(np.unique(lab), np.arange(0, n + 1))
    assert np.array_equal(lab > 0, a > 0)
    again = DeviceChunk(out.tensor.view(torch.int32)).connected_component(connectivity=6)
    assert again.num_components == n and np.array_equal(again.tensor.cpu().numpy().view(np.uint32), lab)
