"""CPU: the super-tile counting binning (csrc/binning.cuh) restated in numpy, step for step, against
the reference's definition of the binned state (rasterizer_impl.cu:70-138, :309-325: instances sorted by the 64-bit key
tile << 32 | depth bits, ties in emission = Gaussian index order). The CUDA kernels are checked bit for bit against the
reference rasterizer on the GPU; this test pins the ALGORITHM - chunking, the per-chunk / per-tile counts, the two
scans, the fill order - on cases a GPU test cannot enumerate as freely (degenerate rects, ties, huge splats, empty
super-tiles, chunk boundaries)."""
import numpy as np
import pytest

ST = 8           # super-tile side in tiles (ST_SHIFT = 3)
CHUNK = 1024     # copies per chunk (BIN_CHUNK)


def reference_order(rects, depth, gx, gy):
    """point_list and ranges as the reference builds them: one instance per (Gaussian, tile), row-major over the rect
    in Gaussian order, stable sort by (tile, depth bits)."""
    tiles, ids = [], []
    for g, (x0, y0, x1, y1) in enumerate(rects):
        for y in range(y0, y1):
            for x in range(x0, x1):
                tiles.append(y * gx + x)
                ids.append(g)
    tiles, ids = np.array(tiles, np.int64), np.array(ids, np.int64)
    key = (tiles << 32) | depth.view(np.uint32)[ids].astype(np.int64)
    order = np.argsort(key, kind="stable")
    tiles, ids = tiles[order], ids[order]
    ranges = np.zeros((gx * gy, 2), np.int64)
    for t in np.unique(tiles):
        w = np.nonzero(tiles == t)[0]
        ranges[t] = (w[0], w[-1] + 1)
    return ids, ranges


def supertile_binning(rects, depth, gx, gy):
    sgx, sgy = (gx + ST - 1) // ST, (gy + ST - 1) // ST
    # 1. the Gaussians in (depth bits, index) order; one copy per overlapped super-tile, row-major over the
    #    super-tile rect, in that order; stable sort of the copies by super-tile index
    perm = np.argsort(depth.view(np.uint32), kind="stable")
    cs, cid = [], []
    for g in perm:
        x0, y0, x1, y1 = rects[g]
        if x1 <= x0 or y1 <= y0:
            continue
        for sy in range(y0 // ST, (y1 + ST - 1) // ST):
            for sx in range(x0 // ST, (x1 + ST - 1) // ST):
                cs.append(sy * sgx + sx)
                cid.append(g)
    cs, cid = np.array(cs, np.int64), np.array(cid, np.int64)
    order = np.argsort(cs, kind="stable")
    cs, cid = cs[order], cid[order]
    # 2. chunks of CHUNK copies that never straddle a super-tile; per chunk and tile of the super-tile: copies covering it
    chunks = []   # (super-tile, first, n)
    for s in range(sgx * sgy):
        w = np.nonzero(cs == s)[0]
        for a in range(0, len(w), CHUNK):
            chunks.append((s, w[0] + a, min(CHUNK, len(w) - a)))

    def mask(g, s):   # 8 x 8 coverage of super-tile s by rect g (clipped)
        ox, oy = (s % sgx) * ST, (s // sgx) * ST
        x0, y0, x1, y1 = rects[g]
        m = np.zeros((ST, ST), bool)
        m[max(y0, oy) - oy:max(min(y1, oy + ST) - oy, 0), max(x0, ox) - ox:max(min(x1, ox + ST) - ox, 0)] = True
        return m.reshape(-1)

    ccnt = np.zeros((len(chunks), ST * ST), np.int64)
    for c, (s, first, n) in enumerate(chunks):
        for j in range(first, first + n):
            ccnt[c] += mask(cid[j], s)
    # 3. per super-tile and tile: exclusive scan over its chunks; tile totals; exclusive scan over the tiles in tile order
    cbase = np.zeros_like(ccnt)
    tile_cnt = np.zeros(gx * gy, np.int64)
    run = {}
    for c, (s, first, n) in enumerate(chunks):
        r = run.setdefault(s, np.zeros(ST * ST, np.int64))
        cbase[c] = r
        run[s] = r + ccnt[c]
    for s, r in run.items():
        for t in range(ST * ST):
            tx, ty = (s % sgx) * ST + t % ST, (s // sgx) * ST + t // ST
            if tx < gx and ty < gy:
                tile_cnt[ty * gx + tx] = r[t]
            else:
                assert r[t] == 0
    start = np.concatenate([[0], np.cumsum(tile_cnt)[:-1]])
    ranges = np.zeros((gx * gy, 2), np.int64)
    nz = tile_cnt > 0
    ranges[nz, 0], ranges[nz, 1] = start[nz], (start + tile_cnt)[nz]
    # 4. fill: chunk after chunk, copy after copy, every covered tile appends the Gaussian index
    point_list = np.full(int(tile_cnt.sum()), -1, np.int64)
    for c, (s, first, n) in enumerate(chunks):
        pos = cbase[c].copy()
        for j in range(first, first + n):
            for t in np.nonzero(mask(cid[j], s))[0]:
                tx, ty = (s % sgx) * ST + t % ST, (s // sgx) * ST + t // ST
                point_list[start[ty * gx + tx] + pos[t]] = cid[j]
                pos[t] += 1
    return point_list, ranges


def random_case(seed, P, gx, gy, big=0, ties=False):
    g = np.random.default_rng(seed)
    x0 = g.integers(0, gx, P)
    y0 = g.integers(0, gy, P)
    w = g.integers(0, 4, P)            # 0 = culled (empty rect)
    h = g.integers(0, 4, P)
    rects = np.stack([x0, y0, np.minimum(x0 + w, gx), np.minimum(y0 + h, gy)], 1)
    if big:
        rects[:big] = (0, 0, gx, gy)   # screen-filling
    depth = g.random(P).astype(np.float32) + 0.2
    if ties:
        depth[: P // 2] = depth[0]     # equal keys: the order inside a tile falls back to the Gaussian index
    return [tuple(int(v) for v in r) for r in rects], depth


@pytest.mark.parametrize("seed,P,gx,gy,big,ties", [
    (0, 300, 13, 7, 0, False),         # image not a multiple of the super-tile
    (1, 500, 16, 16, 3, False),        # screen-filling splats
    (2, 400, 9, 20, 0, True),          # ties broken by index
    (3, 2600, 8, 8, 2, True),          # one super-tile, several 1024-copy chunks
    (4, 50, 40, 3, 0, False),          # mostly empty super-tiles
])
def test_supertile_binning_equals_the_sorted_order(seed, P, gx, gy, big, ties):
    rects, depth = random_case(seed, P, gx, gy, big, ties)
    ref_list, ref_ranges = reference_order(rects, depth, gx, gy)
    got_list, got_ranges = supertile_binning(rects, depth, gx, gy)
    assert np.array_equal(got_ranges, ref_ranges)
    assert np.array_equal(got_list, ref_list)
