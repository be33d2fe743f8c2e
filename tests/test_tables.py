"""Hand-written device tables checked on the CPU (no GPU, no library call)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rank_update_units_cover_every_upper_tile_once():
    """c_units9 (uvs_build3.cu): the eight {first row tile, first column tile, 3x3 mask} units of k_window_system must
    cover the 45 upper 8x8 tiles of the 9 x 9 tile grid exactly once, with at most six tiles per unit, and every mask
    must be one of those rank_chunk<> is compiled for."""
    src = open(os.path.join(ROOT, "uv-slam_b200", "csrc", "uvs_build3.cu")).read()
    body = src[src.index("c_units9[8][3] = {"):]
    body = body[:body.index("};")]
    units = [tuple(int(x, 0) for x in m) for m in re.findall(r"\{(\d+),\s*(\d+),\s*(0x[0-9a-fA-F]+)\}", body)]
    assert len(units) == 8
    compiled = {int(m, 16) for m in re.findall(r"case 0x([0-9a-fA-F]+)u: rank_chunk<", src)}
    seen = {}
    for ta, tb, mask in units:
        assert mask in compiled, hex(mask)
        tiles = [(ta + u, tb + v) for u in range(3) for v in range(3) if mask >> (3 * u + v) & 1]
        assert 1 <= len(tiles) <= 6
        for t in tiles:
            assert 0 <= t[0] <= t[1] < 9, t          # upper triangle of the 9 x 9 grid
            assert t not in seen, (t, seen[t])
            seen[t] = (ta, tb, mask)
    assert len(seen) == 45
    # warps w and w + 4 share a scheduler: 12, 12, 12, 9 tiles
    load = [bin(units[w][2]).count("1") + bin(units[w + 4][2]).count("1") for w in range(4)]
    assert max(load) == 12 and sum(load) == 45
