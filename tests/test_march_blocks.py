"""Host half of the march kernel (build_march_blocks in csrc/march_kernels.cuh, shared by celeste_abi.cu and the
emulation driver): the block list of a plan -- image groups, column segments, background masks, partial slots --
against a brute-force reading of the patch geometry.  CPU only."""
import numpy as np
import pytest

import cases
import emul_lib

MARCH_NIMG, MAXSEG = 5, 51


def _box(p):
    H2, W2 = p.active_pixel_bitmap.shape
    return int(p.bitmap_offset[0]), int(p.bitmap_offset[1]), H2, W2


@pytest.mark.parametrize("name", ["two_body", "clipped_and_empty", "crowded", "small_field", "wide_patch", "seven_images"])
@pytest.mark.parametrize("split", [10 ** 9, 3000, 1])
def test_march_blocks_cover_every_source_once(name, split):
    images, patches, tasks = cases.get(name)
    N = patches.shape[1]
    blocks, part = emul_lib.march_blocks(patches, tasks, split)
    assert part[0] == 0 and part[-1] == len(blocks)
    assert sorted(b["pidx"] for b in blocks) == list(range(len(blocks)))          # one partial vector per block
    cost = []
    for t, (rows, act, _vp) in enumerate(tasks):
        mine = sorted((b for b in blocks if b["task"] == t), key=lambda b: b["pidx"])
        assert [b["pidx"] for b in mine] == list(range(part[t], part[t + 1]))     # the epilogue sums exactly these, in order
        # image groups tile 0..N without gaps, each within MARCH_NIMG images
        assert mine[0]["n0"] == 0 and mine[-1]["n1"] == N
        for a, b in zip(mine, mine[1:]):
            assert a["n1"] == b["n0"]
        a_row = rows[act[0] - 1] - 1
        tot = sum(_box(patches[a_row, n])[2] * _box(patches[a_row, n])[3] for n in range(N))
        for b in mine:
            assert 1 <= b["n1"] - b["n0"] <= MARCH_NIMG
            if tot <= split:
                assert b["n1"] - b["n0"] == min(N - b["n0"], MARCH_NIMG)            # light sources are not cut
            boxes = [_box(patches[a_row, n]) for n in range(b["n0"], b["n1"])]
            live = [bx for bx in boxes if bx[2] > 0 and bx[3] > 0]
            # walks = rows x segments of the non-empty patches; segments never longer than the restart cap
            assert b["walks"] == sum(bx[2] for bx in live) * b["nseg"]
            if live:
                maxw = max(bx[3] for bx in live)
                assert b["nseg"] >= 1 and -(-maxw // b["nseg"]) <= MAXSEG
            # background mask == some other source of the task overlaps the active patch (neighbour's last column excluded)
            for k, n in enumerate(range(b["n0"], b["n1"])):
                oh, ow, H2, W2 = _box(patches[a_row, n])
                want = False
                if H2 > 0 and W2 > 0:
                    for j, r in enumerate(rows):
                        if j == act[0] - 1:
                            continue
                        ph, pw, pH2, pW2 = _box(patches[r - 1, n])
                        rows_meet = max(oh, ph) + 1 <= min(oh + H2, ph + pH2)
                        cols_meet = max(ow, pw) + 1 <= min(ow + W2, pw + pW2 - 1)
                        want = want or (rows_meet and cols_meet)
                assert bool((b["hasbg"] >> k) & 1) == want, (name, t, n)
            cost.append(tot)
    # heaviest first (stable): launch order is by decreasing cost
    order_cost = []
    for b in blocks:
        rows, act, _ = tasks[b["task"]]
        a_row = rows[act[0] - 1] - 1
        px = sum(_box(patches[a_row, n])[2] * _box(patches[a_row, n])[3] for n in range(b["n0"], b["n1"]))
        order_cost.append(px * (1 + (len(rows) - 1) // 4))
    assert order_cost == sorted(order_cost, reverse=True)
