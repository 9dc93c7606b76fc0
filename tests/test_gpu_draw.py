"""GPU parity of the target rasterisation (lv_draw_boxes, generating_train_bev.py:127-139) - bit-exact
against the golden targets the reference's own draw_boxes painted with cv2
(tests/golden/ref_draw_boxes.npz) and against oracle/draw_oracle.py on seeded polygons."""
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle.gen_golden_draw import SCENES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bev():
    import torch
    assert torch.cuda.is_available()
    from lyft3d_b200 import bev
    return bev


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_draw_boxes.npz"))


class _Box:
    def __init__(self, c, name):
        self._c, self.name = c, name

    def bottom_corners(self):
        return self._c


@pytest.mark.parametrize("i", range(len(SCENES)))
def test_scenes_match_reference_targets(bev, gold, i):
    seed, n, shape, vs, zo, ext = SCENES[i]
    corners, cls = synth.box_scene(seed, n, ext)
    im = np.zeros(shape, dtype=np.float32)
    bev.draw_boxes(im, vs, [_Box(c, synth.BOX_CLASSES[k]) for c, k in zip(corners, cls)], synth.BOX_CLASSES, zo)
    assert np.array_equal(im[:, :, 0].astype(np.uint8), gold["scene%d" % i])
    assert np.array_equal(im[:, :, 0], im[:, :, 1]) and np.array_equal(im[:, :, 0], im[:, :, 2])


def test_single_polygons_match_cv2_golden(bev, gold):
    polys = gold["polys"]
    masks = np.unpackbits(gold["poly_masks"])[: polys.shape[0] * 48 * 40].reshape(-1, 48, 40)
    shape, vs = (48, 40, 3), (0.5, 0.5, 1.0)
    # vertices with negative coordinates: trunc(-0.75) = 0, so place them at v - 0.25 instead
    c = np.zeros((polys.shape[0], 3, 4), dtype=np.float64)
    frac = np.where(polys >= 0, 0.25, -0.25)
    c[:, 0, :] = (polys[:, :, 0] + frac[:, :, 0] - shape[0] / 2) * vs[0]
    c[:, 1, :] = (polys[:, :, 1] + frac[:, :, 1] - shape[1] / 2) * vs[1]
    offs = np.arange(polys.shape[0] + 1, dtype=np.int64)          # one polygon per frame
    out = bev.rasterize_targets(c, np.ones(polys.shape[0], np.int32), offs, shape, vs, 0.0)
    assert out.shape == masks.shape
    bad = [k for k in range(polys.shape[0]) if not np.array_equal(out[k], masks[k])]
    assert not bad, (bad[:5], polys[bad[0]].tolist())


def test_random_scenes_against_oracle_on_device(bev):
    import torch
    from oracle import draw_oracle
    rng = np.random.default_rng(31)
    frames, cols = [], []
    for f in range(12):
        n = int(rng.integers(0, 90))
        c, k = synth.box_scene(500 + f, n, 80.0) if n else (np.zeros((0, 3, 4)), np.zeros(0, np.int32))
        if f % 3 == 0 and n:                      # thin slivers and huge boxes
            c[: n // 2, :2, :] *= rng.uniform(0.2, 3.0)
        frames.append(c)
        cols.append(k + 1)
    offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in frames])]).astype(np.int64)
    d_c = torch.from_numpy(np.concatenate(frames)).cuda()
    d_k = torch.from_numpy(np.concatenate(cols).astype(np.int32)).cuda()
    shape, vs = (200, 176, 3), (0.6, 0.7, 1.5)
    out = bev.rasterize_targets(d_c, d_k, offs, shape, vs, -2.0).cpu().numpy()
    for f, (c, k) in enumerate(zip(frames, cols)):
        ref = np.zeros(shape, dtype=np.float32)
        draw_oracle.draw_boxes(ref, vs, list(c), list(k), -2.0)
        assert np.array_equal(out[f], ref[:, :, 0].astype(np.uint8)), f


def test_painters_order_and_errors(bev):
    big = np.array([[-10, 10, 10, -10], [-10, -10, 10, 10], [0, 0, 0, 0]], dtype=np.float64)
    small = big * 0.3
    offs = np.array([0, 2], dtype=np.int64)
    a = bev.rasterize_targets(np.stack([big, small]), np.array([3, 7], np.int32), offs, synth.BEV_SHAPE,
                              synth.BEV_VOXEL_SIZE)[0]
    b = bev.rasterize_targets(np.stack([small, big]), np.array([7, 3], np.int32), offs, synth.BEV_SHAPE,
                              synth.BEV_VOXEL_SIZE)[0]
    assert set(np.unique(a)) == {0, 3, 7} and set(np.unique(b)) == {0, 3}
    with pytest.raises(ValueError):
        bev.draw_boxes(np.zeros(synth.BEV_SHAPE, np.float32), synth.BEV_VOXEL_SIZE, [_Box(big, "tram")],
                       synth.BOX_CLASSES)


def test_more_candidates_than_the_shared_list(bev):
    """> 1024 boxes meeting one tile: the kernel falls back to walking every box of the frame."""
    from oracle import draw_oracle
    rng = np.random.default_rng(5)
    n = 1300
    c, k = synth.box_scene(900, n, 3.0)            # all centres within 3 m of the origin
    shape, vs = (96, 96, 3), (0.4, 0.4, 1.5)
    out = bev.rasterize_targets(c, k + 1, np.array([0, n], dtype=np.int64), shape, vs, 0.0)[0]
    ref = np.zeros(shape, dtype=np.float32)
    draw_oracle.draw_boxes(ref, vs, list(c), list(k + 1), 0.0)
    assert np.array_equal(out, ref[:, :, 0].astype(np.uint8))
    assert rng is not None
