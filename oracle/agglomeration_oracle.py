"""CPU oracle for the `agglomerate` operator (TEST INFRASTRUCTURE).

Reference path: ``chunkflow/plugins/agglomerate.py:8-48`` -- flip the channel order (chunkflow stores x, y, z; waterz wants
z, y, x), make the map contiguous float32, call ``waterz.agglomerate(affs, [threshold], fragments=fragments,
aff_threshold_low=, aff_threshold_high=, scoring_function='OneMinus<MeanAffinity<RegionGraphType, ScoreValue>>')`` and take
the one segmentation it yields.

PARITY UNPINNED: waterz (github.com/funkey/waterz, a Cython/C++ package; the reference does not pin a version and does not
list it in requirements.txt) is neither vendored in /root/reference nor installed in this image, and the reference holds no
test or golden vector for this plugin.  This file restates waterz's published algorithm:

  1. fragments  = the steepest-ascent watershed of Zlateski & Seung (arXiv:1505.00249; waterz ``backend/watershed.hpp``):
     every voxel keeps the edges of its 6-neighbourhood whose affinity equals the voxel's maximum (if that exceeds
     ``aff_threshold_low``) or reaches ``aff_threshold_high``; plateaus are divided from their corners; the weakly
     connected components of the remaining directed graph are the basins, numbered 1..N in raster order of their first
     voxel, voxels without an edge are background 0.
       * ``watershed_literal``  -- the sequential algorithm, statement by statement (queue order and all), pure Python;
       * ``watershed``          -- the order-independent form the CUDA kernels implement: identical for plateau corners
         (their choice only depends on local information, see ``_final_direction``), and for plateau interiors the
         voxel drains towards its highest-numbered direction among the neighbours that are one breadth-first step
         closer to a corner -- the sequential code may also pick a neighbour of the SAME distance that happens to sit
         earlier in its queue.  The two agree wherever no plateau has interior voxels with such a choice: always for
         maps without exact ties (float32 network outputs), and for saturated regions (>= aff_threshold_high), which have
         no corners and stay whole.  tests/test_segmentation_agglomerate.py checks both statements.
  2. region graph = one edge per pair of touching fragments with the statistics of the affinities between them (waterz
     ``backend/region_graph.hpp`` + ``MeanAffinityProvider``): sum and count.  Sums are accumulated in 2^-30 fixed point
     (order independent, so that the device's atomics and this scan agree bit for bit); waterz adds float32 values in
     scan order -- the means differ by rounding only.
  3. agglomeration = waterz ``IterativeRegionMerging``: merge the edge of the lowest score ``1 - mean affinity`` until
     the lowest score reaches the threshold; merged regions pool the statistics of their edges.  Equal scores are taken
     in the order of the smallest ORIGINAL fragment pair pooled into the edge (waterz: the order its heap happens to hold,
     not part of its published interface); the merged region keeps the smaller id (waterz keeps one of the two ids as
     well).  The result is NOT renumbered (neither is waterz's).
"""
import heapq

import numpy as np

FIXED_ONE = float(1 << 30)
_INF = np.uint32(0xFFFFFFFF)
# direction d: 0..2 = towards the lower neighbour along axis d (z, y, x), 3..5 = towards the upper neighbour
_OPP = (3, 4, 5, 0, 1, 2)


def _edge_weights(affs: np.ndarray, low: float) -> np.ndarray:
    """(6, z, y, x): the affinity of each voxel's edge in direction d; outside the volume = ``low`` (watershed.hpp)."""
    affs = np.asarray(affs, np.float32)
    assert affs.ndim == 4 and affs.shape[0] == 3
    w = np.full((6,) + affs.shape[1:], np.float32(low), np.float32)
    w[0, 1:] = affs[0, 1:]
    w[1, :, 1:] = affs[1, :, 1:]
    w[2, :, :, 1:] = affs[2, :, :, 1:]
    w[3, :-1] = affs[0, 1:]
    w[4, :, :-1] = affs[1, :, 1:]
    w[5, :, :, :-1] = affs[2, :, :, 1:]
    return w


def steepest_ascent_mask(affs: np.ndarray, low: float, high: float) -> np.ndarray:
    """Step 1 of the watershed: bit d of a voxel = it keeps its edge in direction d."""
    w = _edge_weights(affs, low)
    m = w.max(axis=0)
    bits = np.zeros(w.shape[1:], np.uint8)
    for d in range(6):
        bits |= (((w[d] == m) | (w[d] >= np.float32(high))) & (m > np.float32(low))).astype(np.uint8) << d
    return bits


def _shift(a: np.ndarray, d: int, fill):
    """value of the neighbour in direction d at every voxel (``fill`` outside the volume)."""
    out = np.full_like(a, fill)
    ax = d % 3
    src = [slice(None)] * 3
    dst = [slice(None)] * 3
    if d < 3:
        dst[ax], src[ax] = slice(1, None), slice(None, -1)
    else:
        dst[ax], src[ax] = slice(None, -1), slice(1, None)
    out[tuple(dst)] = a[tuple(src)]
    return out


def plateau_distance(bits: np.ndarray):
    """(dist, recip): breadth-first distance over two-way edges from the plateau corners (voxels with an edge that the
    neighbour does not return); 0xFFFFFFFF where no corner is reachable (and on the background)."""
    recip = np.zeros((6,) + bits.shape, bool)
    for d in range(6):
        recip[d] = ((bits >> d) & 1).astype(bool) & ((_shift(bits, d, 0) >> _OPP[d]) & 1).astype(bool)
    has = [((bits >> d) & 1).astype(bool) for d in range(6)]
    corner = np.zeros(bits.shape, bool)
    for d in range(6):
        corner |= has[d] & ~recip[d]
    dist = np.full(bits.shape, _INF, np.uint32)
    dist[corner] = 0
    level = 0
    while True:
        level += 1
        reach = np.zeros(bits.shape, bool)
        for d in range(6):
            reach |= has[d] & (_shift(dist, d, _INF) == level - 1)
        new = reach & (dist == _INF) & (bits != 0)
        if not new.any():
            break
        dist[new] = level
    return dist, recip


def _final_direction(bits, dist, recip):
    """The ONE edge every voxel reached from a corner keeps (-1: background; -2: plateau without a corner, keeps all).
    Corners: the sequential code assigns ``to_set`` for every kept edge d, in the order d = 0..5, whose target does not point
    back AT THAT MOMENT -- targets that never did, and corners processed earlier (they sit at a lower raster index, i.e. in
    directions 0..2, and have been reduced to an edge that cannot lead to a voxel still holding its two-way edge) -- and the
    last assignment wins."""
    final = np.full(bits.shape, -1, np.int8)
    final[(bits != 0) & (dist == _INF)] = -2
    for d in range(6):
        has = ((bits >> d) & 1).astype(bool)
        nb_dist = _shift(dist, d, _INF)
        has = has & (_shift(bits, d, 0) != 0)   # (a voxel beside NaN affinities has no edges and is never joined to a basin)
        cand0 = has & (~recip[d] | ((nb_dist == 0) if d < 3 else False))
        cand = np.where(dist == 0, cand0, has & (dist != _INF) & (dist > 0) & (nb_dist == dist - 1))
        final[cand & (dist != _INF)] = d
    return final


def _label_components(n_vox, edges_a, edges_b, foreground):
    """Weakly connected components -> labels 1..N in raster order of each component's first voxel."""
    parent = np.arange(n_vox, dtype=np.int64)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in zip(edges_a.tolist(), edges_b.tolist()):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    out = np.zeros(n_vox, np.uint32)
    seen, nxt = {}, 1
    for i in np.flatnonzero(foreground).tolist():
        r = find(i)
        if r not in seen:
            seen[r] = nxt
            nxt += 1
        out[i] = seen[r]
    return out


def watershed(affs: np.ndarray, low: float = 0.001, high: float = 0.9999) -> np.ndarray:
    """Fragments (z, y, x) uint32, the order-independent form (see the module docstring)."""
    bits = steepest_ascent_mask(affs, low, high)
    dist, recip = plateau_distance(bits)
    final = _final_direction(bits, dist, recip)
    shape = bits.shape
    strides = (shape[1] * shape[2], shape[2], 1)
    idx = np.arange(bits.size, dtype=np.int64).reshape(shape)
    ea, eb = [], []
    for d in range(6):
        step = -strides[d] if d < 3 else strides[d - 3]
        sel = (final == d) | ((final == -2) & (((bits >> d) & 1) == 1))
        ea.append(idx[sel])
        eb.append(idx[sel] + step)
    return _label_components(bits.size, np.concatenate(ea), np.concatenate(eb), (bits != 0).ravel()).reshape(shape)


def watershed_literal(affs: np.ndarray, low: float = 0.001, high: float = 0.9999) -> np.ndarray:
    """waterz ``backend/watershed.hpp``, statement by statement (small volumes only)."""
    bits = steepest_ascent_mask(affs, low, high)
    shape = bits.shape
    seg = bits.astype(np.int64).ravel().tolist()
    size = len(seg)
    strides = (shape[1] * shape[2], shape[2], 1)
    step = [-strides[0], -strides[1], -strides[2], strides[0], strides[1], strides[2]]
    dirmask = [1, 2, 4, 8, 16, 32]
    idirmask = [8, 16, 32, 1, 2, 4]
    visited, high_bit = 0x40, 1 << 62
    # 2. plateau corners
    bfs = []
    for idx in range(size):
        for d in range(6):
            if seg[idx] & dirmask[d] and not (seg[idx + step[d]] & idirmask[d]):
                seg[idx] |= visited
                bfs.append(idx)
                break
    # 3. divide the plateaus
    k = 0
    while k < len(bfs):
        idx = bfs[k]
        to_set = 0
        for d in range(6):
            if seg[idx] & dirmask[d]:
                him = idx + step[d]
                if seg[him] & idirmask[d]:
                    if not (seg[him] & visited):
                        bfs.append(him)
                        seg[him] |= visited
                else:
                    to_set = dirmask[d]
        seg[idx] = to_set
        k += 1
    # 4. basins
    next_id = 1
    for idx in range(size):
        if seg[idx] == 0:
            seg[idx] = high_bit
        if not (seg[idx] & high_bit) and seg[idx]:
            bfs = [idx]
            seg[idx] |= visited
            k = 0
            while k < len(bfs):
                me = bfs[k]
                for d in range(6):
                    if seg[me] & dirmask[d]:
                        him = me + step[d]
                        if seg[him] & high_bit:
                            for it in bfs:
                                seg[it] = seg[him]
                            bfs = []
                            break
                        elif not (seg[him] & visited):
                            seg[him] |= visited
                            bfs.append(him)
                k += 1
            if bfs:
                for it in bfs:
                    seg[it] = high_bit | next_id
                next_id += 1
    return (np.array(seg, np.int64) & (high_bit - 1)).astype(np.uint32).reshape(shape)


def quantize_affinity(a: np.ndarray) -> np.ndarray:
    """2^-30 fixed point of an affinity clamped to [0, 1] (NaN counts as 0), round half to even."""
    a = np.nan_to_num(np.asarray(a, np.float32).astype(np.float64), nan=0.0)
    return np.rint(np.clip(a, 0.0, 1.0) * FIXED_ONE).astype(np.int64)


def region_graph(affs: np.ndarray, fragments: np.ndarray):
    """(u, v, sum_fixed, count) sorted by (u, v), u < v: waterz ``get_region_graph`` with the MeanAffinity statistics."""
    affs = np.asarray(affs, np.float32)
    frag = np.asarray(fragments).astype(np.int64)
    keys, vals = [], []
    for c in range(3):
        hi = [slice(None)] * 3
        lo = [slice(None)] * 3
        hi[c], lo[c] = slice(1, None), slice(None, -1)
        a, b = frag[tuple(hi)], frag[tuple(lo)]
        sel = (a != 0) & (b != 0) & (a != b)
        u, v = np.minimum(a[sel], b[sel]), np.maximum(a[sel], b[sel])
        keys.append((u << 32) | v)
        vals.append(quantize_affinity(affs[c][tuple(hi)][sel]))
    keys, vals = np.concatenate(keys), np.concatenate(vals)
    uniq, inv, counts = np.unique(keys, return_inverse=True, return_counts=True)
    sums = np.zeros(len(uniq), np.int64)
    np.add.at(sums, inv, vals)
    return ((uniq >> 32).astype(np.uint32), (uniq & 0xFFFFFFFF).astype(np.uint32), sums.astype(np.uint64),
            counts.astype(np.uint32))


def score(sum_fixed: int, count: int) -> float:
    """OneMinus<MeanAffinity>: 1 - mean (double arithmetic: one division, one subtraction)."""
    return 1.0 - float(sum_fixed) / (float(count) * FIXED_ONE)


def agglomerate_edges(num_nodes: int, u, v, sum_fixed, count, threshold: float) -> np.ndarray:
    """root_of (num_nodes,) uint32: the id every node ends up with (0 stays 0).

    Every edge carries an ANCHOR: the smallest (u, v) pair of ORIGINAL fragments among the faces pooled into it.  The queue is
    ordered by (score, anchor): among equal scores the edge holding the smallest original fragment pair goes first -- a rule
    that does not depend on how an implementation stores clusters or which side of a merge it relabels (the native library
    moves the SHORTER adjacency list; this restatement always moves the larger id's)."""
    adj = [dict() for _ in range(num_nodes)]
    heap = []
    for a, b, s, c in zip(np.asarray(u).tolist(), np.asarray(v).tolist(), np.asarray(sum_fixed).tolist(), np.asarray(count).tolist()):
        a, b = min(a, b), max(a, b)
        anchor = (a << 32) | b
        adj[a][b] = (s, c, anchor)
        adj[b][a] = (s, c, anchor)
        heap.append((score(s, c), anchor, a, b, s, c))
    heapq.heapify(heap)
    alive = [True] * num_nodes
    parent = list(range(num_nodes))
    thr = float(np.float32(threshold))
    while heap:
        sc, anchor, a, b, s, c = heapq.heappop(heap)
        if not (alive[a] and alive[b]) or adj[a].get(b) != (s, c, anchor):
            continue   # stale
        if not sc < thr:
            break
        # b (the larger id) is merged into a: the cluster keeps its smallest id
        alive[b] = False
        parent[b] = a
        del adj[a][b]
        for n, (sn, cn, an) in adj[b].items():
            if n == a:
                continue
            del adj[n][b]
            if n in adj[a]:
                so, co, ao = adj[a][n]
                sn, cn, an = sn + so, cn + co, min(an, ao)
            adj[a][n] = (sn, cn, an)
            adj[n][a] = (sn, cn, an)
            heapq.heappush(heap, (score(sn, cn), an, min(a, n), max(a, n), sn, cn))
        adj[b] = {}
    out = np.zeros(num_nodes, np.uint32)
    for i in range(num_nodes):
        r = i
        while parent[r] != r:
            r = parent[r]
        out[i] = r
    return out


def agglomerate(affs: np.ndarray, threshold: float = 0.7, fragments=None, aff_threshold_low: float = 0.001,
                aff_threshold_high: float = 0.9999, flip_channel: bool = True) -> np.ndarray:
    """``plugins/agglomerate.py: execute`` -> (z, y, x) uint64 segmentation."""
    affs = np.asarray(affs)
    if flip_channel:
        affs = np.flip(affs, axis=0)                      # reference agglomerate.py:28-29
    affs = np.ascontiguousarray(affs, dtype=np.float32)   # :33
    if fragments is None:
        fragments = watershed(affs, aff_threshold_low, aff_threshold_high)
    fragments = np.asarray(fragments)
    num_nodes = int(fragments.max()) + 1 if fragments.size else 1
    u, v, s, c = region_graph(affs, fragments)
    root = agglomerate_edges(num_nodes, u, v, s, c, threshold)
    return root[fragments].astype(np.uint64)
