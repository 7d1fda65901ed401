"""CPU oracle for the `connected-components` operator (TEST INFRASTRUCTURE).

Reference path: ``Chunk.connected_component`` (chunkflow/chunk/base.py:128-137) = ``Chunk.threshold`` (``array > threshold`` ->
uint8, base.py:728-737) for non-segmentation chunks, then ``cc3d.connected_components(seg, connectivity=connectivity)``.

PARITY UNPINNED: cc3d (PyPI ``connected-components-3d``, requirements.txt of the reference, no version pin) is a third-party
package that is neither vendored in /root/reference nor installed in this image, and the reference has no test with expected
label values for this operator (tests/flow/test_flow.py only runs it).  This file restates cc3d's published behaviour:
  * two voxels belong to one component when they are 6- / 18- / 26-neighbours and carry the SAME non-zero value,
  * 0 is background,
  * the output is renumbered 1..N in the order in which components are first met in a raster scan of the array's memory
    (cc3d's final relabel pass; for a C-order (z, y, x) array: x fastest).
``connected_components`` uses scipy.ndimage.label per distinct value (a C implementation of the same definition);
``connected_components_slow`` is a pure-Python union-find raster scan that pins it on small volumes.
"""
import numpy as np


def threshold(array: np.ndarray, thr: float) -> np.ndarray:
    """``Chunk.threshold`` (reference chunk/base.py:728-737)."""
    out = array > thr
    if out.ndim == 4:
        assert out.shape[0] == 1
        out = out[0]
    return out.astype(np.uint8)


def _renumber_by_first_occurrence(labels: np.ndarray) -> np.ndarray:
    flat = labels.ravel()
    vals, first = np.unique(flat, return_index=True)
    keep = vals != 0
    vals, first = vals[keep], first[keep]
    order = np.argsort(first, kind="stable")
    lut = np.zeros(int(flat.max()) + 1 if flat.size else 1, np.uint32)
    lut[vals[order]] = np.arange(1, len(order) + 1, dtype=np.uint32)
    return lut[labels].astype(np.uint32)


def connected_components(seg: np.ndarray, connectivity: int = 6) -> np.ndarray:
    from scipy import ndimage
    assert seg.ndim == 3 and connectivity in (6, 18, 26)
    structure = ndimage.generate_binary_structure(3, {6: 1, 18: 2, 26: 3}[connectivity])
    out = np.zeros(seg.shape, np.int64)
    offset = 0
    for v in np.unique(seg):
        if v == 0:
            continue
        lab, n = ndimage.label(seg == v, structure=structure)
        out[lab > 0] = lab[lab > 0] + offset
        offset += n
    return _renumber_by_first_occurrence(out)


def connected_components_slow(seg: np.ndarray, connectivity: int = 6) -> np.ndarray:
    """Union-find over a raster scan, pure Python: small volumes only."""
    assert seg.ndim == 3
    Z, Y, X = seg.shape
    nb = [(dz, dy, dx) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)
          if (dz, dy, dx) < (0, 0, 0) and abs(dz) + abs(dy) + abs(dx) <= {6: 1, 18: 2, 26: 3}[connectivity]]
    parent = list(range(Z * Y * X))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for z in range(Z):
        for y in range(Y):
            for x in range(X):
                v = seg[z, y, x]
                if not v:
                    continue
                i = (z * Y + y) * X + x
                for dz, dy, dx in nb:
                    zz, yy, xx = z + dz, y + dy, x + dx
                    if 0 <= zz < Z and 0 <= yy < Y and 0 <= xx < X and seg[zz, yy, xx] == v:
                        a, b = find(i), find((zz * Y + yy) * X + xx)
                        if a != b:
                            parent[max(a, b)] = min(a, b)
    out = np.zeros(Z * Y * X, np.uint32)
    nxt, seen = 1, {}
    flat = seg.ravel()
    for i in range(Z * Y * X):
        if flat[i]:
            r = find(i)
            if r not in seen:
                seen[r] = nxt
                nxt += 1
            out[i] = seen[r]
    return out.reshape(seg.shape)


def chunk_connected_component(array: np.ndarray, thr=None, connectivity: int = 6, is_segmentation: bool = False) -> np.ndarray:
    """``Chunk.connected_component`` (reference chunk/base.py:128-137)."""
    seg = threshold(array, thr) if (not is_segmentation and thr is not None) else array
    return connected_components(seg, connectivity)
