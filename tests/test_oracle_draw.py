"""Pins oracle/draw_oracle.py (target rasterisation, generating_train_bev.py:127-139) against the
reference: golden targets painted by the reference's own draw_boxes with cv2
(tests/golden/ref_draw_boxes.npz, made by oracle/gen_golden_draw.py), and - where cv2 is importable -
cv2.drawContours itself on seeded polygons, and the reference code re-run live."""
import os

import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import draw_oracle, ref_loader
from oracle.gen_golden_draw import SCENES, polygons


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_draw_boxes.npz"))


@pytest.mark.parametrize("i", range(len(SCENES)))
def test_scenes_match_reference_targets(gold, i):
    seed, n, shape, vs, zo, ext = SCENES[i]
    corners, cls = synth.box_scene(seed, n, ext)
    im = np.zeros(shape, dtype=np.float32)
    draw_oracle.draw_boxes(im, vs, list(corners), list(cls + 1), zo)
    assert np.array_equal(im[:, :, 0].astype(np.uint8), gold["scene%d" % i])
    assert np.array_equal(im[:, :, 0], im[:, :, 2])


def test_single_polygons_match_cv2_golden(gold):
    polys = gold["polys"]
    assert np.array_equal(polys, polygons())
    masks = np.unpackbits(gold["poly_masks"])[: polys.shape[0] * 48 * 40].reshape(-1, 48, 40)
    for p, m in zip(polys, masks):
        a = np.zeros((48, 40), dtype=np.uint8)
        draw_oracle.fill_polygon(a, p, 1)
        assert np.array_equal(a, m), p.tolist()


def test_live_against_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(123)
    for it in range(1500):
        h, w = (40, 36) if it % 2 else (33, 57)
        if it % 3 == 0:
            pts = rng.integers(-12, 70, (4, 2))
        else:                                   # rotated rectangles, like box footprints
            c, wl, th = rng.uniform(-10, 60, 2), rng.uniform(0.2, 30, 2), rng.uniform(0, 2 * np.pi)
            rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
            pts = np.intp((np.array([[1, 1], [1, -1], [-1, -1], [-1, 1]]) * wl / 2) @ rot.T + c)
        a = np.zeros((h, w), dtype=np.uint8)
        draw_oracle.fill_polygon(a, pts, 5)
        b = np.zeros((h, w, 3), dtype=np.float32)
        cv2.drawContours(b, np.intp([pts]), 0, (5, 5, 5), -1)
        assert np.array_equal(a, b[:, :, 0].astype(np.uint8)), pts.tolist()


def test_live_against_reference_code():
    pytest.importorskip("cv2")
    if not ref_loader.available():
        pytest.skip("needs /root/reference")
    corners, cls = synth.box_scene(77, 120)
    ref = np.zeros(synth.BEV_SHAPE, dtype=np.float32)
    ref_loader.run_bev_draw_boxes(ref, synth.BEV_VOXEL_SIZE, corners, cls, synth.BOX_CLASSES, synth.BEV_Z_OFFSET)
    got = np.zeros(synth.BEV_SHAPE, dtype=np.float32)
    draw_oracle.draw_boxes(got, synth.BEV_VOXEL_SIZE, list(corners), list(cls + 1), synth.BEV_Z_OFFSET)
    assert np.array_equal(got, ref)
