"""PNG oracle (oracle/png_oracle.py): the stream it states is a valid PNG that decodes to the image under
cv2 (what the reference reads its files with, dataset.py:83-90) and under an independent zlib decoder,
exactly like the file cv2.imwrite itself produces (generating_train_bev.py:215,224)."""
import numpy as np
import pytest

from lyft3d_b200 import synth
from oracle import bev_oracle as bo, png_oracle as po

cv2 = pytest.importorskip("cv2")


def _bev_u8(f):
    pts = np.ascontiguousarray(synth.c5_frame(f).T)
    raw = bo.create_voxel_pointcloud(pts, synth.BEV_SHAPE, synth.BEV_VOXEL_SIZE, synth.BEV_Z_OFFSET)
    return bo.quantize_u8(bo.normalize_voxel_intensities(raw))


def _images():
    rng = np.random.default_rng(11)
    yield "bev", _bev_u8(2)
    yield "zeros", np.zeros((17, 23, 3), np.uint8)
    yield "dense", rng.integers(0, 256, (40, 31, 3), dtype=np.uint8)
    yield "grey_target", (rng.random((64, 50)) > 0.8).astype(np.uint8) * 7
    yield "grey_runs", np.repeat(rng.integers(0, 256, (9, 13), dtype=np.uint8), 40, axis=1)
    yield "one_pixel", np.array([[200]], np.uint8)
    yield "long_run_lengths", np.concatenate([np.pad(np.full((1, k), 5, np.uint8), ((0, 0), (0, 600 - k)))
                                              for k in (1, 2, 3, 4, 10, 11, 12, 257, 258, 259, 260, 261, 262, 516, 517, 519, 600)])


@pytest.mark.parametrize("name,img", list(_images()))
def test_oracle_png_decodes_to_the_image(name, img):
    png = po.encode_png(img)
    dec = cv2.imdecode(np.frombuffer(png, np.uint8), cv2.IMREAD_UNCHANGED)
    assert dec.shape == img.shape and np.array_equal(dec, img)
    z = po.decode_png_zlib(png)                      # checks every chunk CRC and the Adler-32 on the way
    z = z[:, :, ::-1] if img.ndim == 3 else z[:, :, 0]
    assert np.array_equal(z, img)
    # the reference's own writer: same decoded content, same IHDR
    ok, ref = cv2.imencode(".png", img)
    assert ok and np.array_equal(cv2.imdecode(ref, cv2.IMREAD_UNCHANGED), dec)
    assert bytes(ref[:8 + 8 + 10]) == png[:8 + 8 + 10]   # signature, IHDR length/type, width, height, depth, colour type


def test_sparse_bev_file_is_an_order_of_magnitude_smaller_than_the_array():
    img = _bev_u8(5)
    assert len(po.encode_png(img)) * 10 < img.size


def test_token_tables():
    assert po.match_token(3) == (po._rev(1, 7), 12) and po.match_token(258)[1] == 13
    for length in range(3, 259):
        v, n = po.match_token(length)
        assert 12 <= n <= 18 and v < (1 << n)
    assert po.literal_token(0) == (0x0C, 8) and po.literal_token(255)[1] == 9
