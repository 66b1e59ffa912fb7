"""Cross-stage dependency rule of the trunk launch (flope_b200/csrc/trunk_chain.cuh, TrunkStage::producer): the tile range
of the previous stage that the producer of a stride-2 conv tile waits for must contain every full-resolution
position the tile's valid outputs read (3x3 stride-2 window, and the folded 1x1 stride-2 projection).

The device rule is restated here line by line and checked by brute force over the geometries the engine builds
(make_geom in engine.cu: pitch = size + 1, one shared zero row / column) - a missed tile would be a race that the GPU
parity tests might never see.
"""
import itertools

import pytest


def waited_range(tile_start, TM, halo_before, halo_after, n_positions, Hp2, Wp2, prevH, prevHp, prevWp, prev_tile_pos,
                 prev_n_m_tiles):
    """trunk_chain.cuh: `else if (prev != nullptr)` branch of TrunkStage::producer."""
    img = Hp2 * Wp2
    a = max(tile_start - halo_before, 0)
    b = min(tile_start + TM - 1 + halo_after, n_positions - 1)
    if a > b:
        return None
    na, ha = a // img, (a % img) // Wp2
    nb, hb = b // img, (b % img) // Wp2
    ra = 2 * ha if 2 * ha < prevH else prevH - 1
    rb = 2 * hb + 1 if 2 * hb + 1 < prevH else prevH - 1
    lo = (na * prevHp + ra) * prevWp
    hi = (nb * prevHp + rb) * prevWp + prevWp - 1
    return lo // prev_tile_pos, min(hi // prev_tile_pos, prev_n_m_tiles - 1)


def needed_tiles(tile_start, TM, n_positions, H2, W2, Hp2, Wp2, prevH, prevW, prevHp, prevWp, prev_tile_pos):
    """Brute force: previous-stage tiles holding a full-res pixel that a valid output of this CTA tile reads."""
    need = set()
    img = Hp2 * Wp2
    for p in range(tile_start, min(tile_start + TM, n_positions)):
        n, r = divmod(p, img)
        i, j = divmod(r, Wp2)
        if i >= H2 or j >= W2:
            continue                                           # padding position: its accumulator is discarded
        for h in range(2 * i - 1, 2 * i + 2):
            for w in range(2 * j - 1, 2 * j + 2):
                if 0 <= h < prevH and 0 <= w < prevW:
                    need.add(((n * prevHp + h) * prevWp + w) // prev_tile_pos)
    return need


# (previous stage side, previous pair-tile positions, this stage's positions per CTA)
CASES = [(56, 1024, 256), (28, 512, 128), (14, 256, 128),          # throughput shapes: 64x4 -> 128x2 -> 256x1 -> 256x1
         (56, 256, 128), (28, 256, 128), (56, 1024, 128),          # latency tiles in some / all stages
         (128, 1024, 256), (64, 512, 128), (32, 256, 128)]         # 512-pixel crops


@pytest.mark.parametrize("side,prev_tile_pos,TM", CASES)
@pytest.mark.parametrize("n_crops", [1, 2, 3, 7])
def test_waited_tiles_cover_every_input(side, prev_tile_pos, TM, n_crops):
    prevH = prevW = side
    prevHp, prevWp = side + 1, side + 1
    H2 = W2 = side // 2
    Hp2, Wp2 = H2 + 1, W2 + 1
    n_positions = n_crops * Hp2 * Wp2
    prev_positions = n_crops * prevHp * prevWp
    prev_n_m_tiles = -(-prev_positions // prev_tile_pos)
    halo = Wp2 + 1
    n_pair_tiles = -(-n_positions // (2 * TM))
    for m, rank in itertools.product(range(n_pair_tiles), (0, 1)):
        tile_start = m * 2 * TM + rank * TM
        need = needed_tiles(tile_start, TM, n_positions, H2, W2, Hp2, Wp2, prevH, prevW, prevHp, prevWp, prev_tile_pos)
        rng = waited_range(tile_start, TM, halo, halo, n_positions, Hp2, Wp2, prevH, prevHp, prevWp, prev_tile_pos,
                           prev_n_m_tiles)
        if not need:
            continue
        assert rng is not None, (m, rank)
        lo, hi = rng
        assert lo <= min(need) and max(need) <= hi, (m, rank, sorted(need), rng)
        assert max(need) < prev_n_m_tiles
        # the rule stays local: at most the needed span plus the rows of halo it rounds out to
        assert hi - lo <= (max(need) - min(need)) + 2 + (4 * halo * 2) // prev_tile_pos + 2


def test_same_stage_rule_is_the_three_neighbouring_tiles():
    """Inside a stage (3x3, stride 1) tile m reads positions [start - Wp - 1, start + TILE + Wp + 1): tiles m-1 .. m+1 as
    long as the halo is shorter than a tile - the condition plan_conv's shapes satisfy for every geometry."""
    for side, tile_pos in ((56, 1024), (28, 512), (14, 256), (7, 256), (56, 256), (128, 1024), (16, 256)):
        assert side + 2 <= tile_pos, (side, tile_pos)
