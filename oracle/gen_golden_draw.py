"""Generates tests/golden/ref_draw_boxes.npz by EXECUTING THE REFERENCE in the build container.

Run:  python -m oracle.gen_golden_draw        (needs /root/reference and cv2; CPU only)

TEST INFRASTRUCTURE.  The reference's own ``draw_boxes`` (generating-dataset/generating_train_bev.py:127-139,
read from the file and executed unchanged by oracle.ref_loader.run_bev_draw_boxes, with the installed
cv2 doing the fill) paints lyft3d_b200.synth.box_scene(seed) for a few seeds and grids; the first
channel of the painted target (``target[:,:,0]``, what :221-223 writes to the PNG) is stored as uint8
together with a set of single polygons painted by cv2.drawContours (in, across and outside the image).
"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)

from lyft3d_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

GOLD = os.path.join(_ROOT, "tests", "golden")

SCENES = [  # (seed, n_boxes, shape, voxel_size, z_offset, extent)
    (9000, 60, (336, 336, 3), (0.4, 0.4, 1.5), -2.0, 75.0),
    (9001, 200, (336, 336, 3), (0.4, 0.4, 1.5), -2.0, 70.0),
    (9002, 40, (1024, 1024, 3), (0.2, 0.2, 1.5), -2.0, 110.0),
    (9003, 80, (128, 128, 3), (0.8, 0.8, 1.5), 0.0, 60.0),
]


def polygons(seed=9100, n=400, size=48):
    rng = np.random.default_rng(seed)
    pts = rng.integers(-size // 2, size + size // 2, (n, 4, 2))
    return pts.astype(np.int32)


def main():
    import cv2
    assert ref_loader.available(), "needs /root/reference"
    out = {}
    for i, (seed, n, shape, vs, zo, ext) in enumerate(SCENES):
        corners, cls = synth.box_scene(seed, n, ext)
        im = np.zeros(shape, dtype=np.float32)
        ref_loader.run_bev_draw_boxes(im, vs, corners, cls, synth.BOX_CLASSES, zo)
        assert np.array_equal(im[:, :, 0], im[:, :, 1]) and np.array_equal(im[:, :, 0], im[:, :, 2])
        out["scene%d" % i] = im[:, :, 0].astype(np.uint8)
        print("scene", i, shape, "painted", int((im[:, :, 0] > 0).sum()))
    polys = polygons()
    masks = np.zeros((polys.shape[0], 48, 40), dtype=np.uint8)
    for k, p in enumerate(polys):
        im = np.zeros((48, 40, 3), dtype=np.float32)
        cv2.drawContours(im, np.intp([p]), 0, (1, 1, 1), -1)
        masks[k] = im[:, :, 0]
    out["polys"] = polys
    out["poly_masks"] = np.packbits(masks, axis=None)
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(GOLD, "ref_draw_boxes.npz"), **out)


if __name__ == "__main__":
    main()
