"""Target-raster oracle (CPU).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates ``draw_boxes`` of generating-dataset/generating_train_bev.py:127-139:

    corners = box.bottom_corners()                                    (3, 4) float64, car space
    corners_voxel = car_to_voxel_coords(corners, im.shape, voxel_size, z_offset).transpose(1, 0)[:, :2]
    cv2.drawContours(im, np.int0([corners_voxel]), 0, (c, c, c), -1)  c = classes.index(box.name) + 1

i.e. boxes are painted in list order (later boxes overwrite earlier ones) as FILLED polygons of four
integer vertices (np.int0 = C truncation).  The fill itself lives in OpenCV (cv2.drawContours with
thickness < 0 -> CollectPolyEdges + FillEdgeCollection, modules/imgproc/src/drawing.cpp), a
dependency of the reference that IS installed in this image (opencv-python-headless 4.13.0), so the
restatement below is PINNED: tests/test_oracle_draw.py compares it with cv2 itself on thousands of
seeded polygons (rectangles in and out of the image, thin boxes, arbitrary quads) and on the golden
vectors oracle/gen_golden_draw.py made with cv2 (tests/golden/ref_draw_boxes.npz).

The rule (8-connected, non-antialiased, shift 0), as pinned:
  * every edge v[i-1] -> v[i] is drawn with ``Line``: the segment is first clipped to the image by
    ``clipLine`` (truncating fp64 interpolation, one coordinate at a time), then walked left to right
    by the integer Bresenham of ``LineIterator``: err0 = D - 2d, a minor step whenever err < 0;
  * a non-horizontal edge becomes a scanline edge [y0, y1) with a 16.16 fixed-point x and
    dx = trunc(((x1 - x0) << 16) / (y1 - y0)); when an endpoint lies outside the image the x's (and,
    unless the clipped segment is horizontal, the y's) of the CLIPPED segment define dx, and the start
    is extrapolated back to the unclipped row: x = xc + (y0 - yc) * dx;
  * rows y in [max(min y0, 0), min(max y1, H)): active edges sorted by x, paired (0,1), (2,3); the
    span [ceil(xa), floor(xb)] is filled.
"""
import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT


def _clip_line(w, h, p1, p2):
    """cv::clipLine(Size, Point2l&, Point2l&): returns (inside, p1, p2); the points may be modified
    even when the result is False."""
    x1, y1 = p1
    x2, y2 = p2
    right, bottom = w - 1, h - 1
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, (x1, y1), (x2, y2)


def _line(img, p1, p2, color):
    """cv::Line, connectivity 8: clip, then LineIterator(leftToRight=true)."""
    h, w = img.shape[:2]
    (x1, y1), (x2, y2) = p1, p2
    if not (0 <= x1 < w and 0 <= x2 < w and 0 <= y1 < h and 0 <= y2 < h):
        ok, (x1, y1), (x2, y2) = _clip_line(w, h, (x1, y1), (x2, y2))
        if not ok:
            return
    dx, dy = x2 - x1, y2 - y1
    sx = sy = 1
    if dx < 0:
        dx, dy = -dx, -dy
        x1, y1 = x2, y2
    if dy < 0:
        dy, sy = -dy, -1
    vert = dy > dx
    if vert:
        dx, dy = dy, dx
    err = dx - 2 * dy
    x, y = x1, y1
    for _ in range(dx + 1):
        img[y, x] = color
        minor = err < 0
        err += -2 * dy + (2 * dx if minor else 0)
        if vert:
            y += sy
            x += sx if minor else 0
        else:
            x += sx
            y += sy if minor else 0


def _cdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a < 0) == (b < 0) else -q


def fill_polygon(img, pts, color):
    """cv2.drawContours(img, [pts], 0, color, -1) for one polygon of integer (x, y) vertices."""
    h, w = img.shape[:2]
    pts = [(int(p[0]), int(p[1])) for p in pts]
    edges = []
    p0 = pts[-1]
    for p1 in pts:
        _line(img, p0, p1, color)
        c0, c1 = p0, p1
        if not (0 <= p0[0] < w and 0 <= p1[0] < w and 0 <= p0[1] < h and 0 <= p1[1] < h):
            _, q0, q1 = _clip_line(w, h, p0, p1)
            if q0[1] != q1[1]:
                c0, c1 = q0, q1
            else:
                c0, c1 = (q0[0], p0[1]), (q1[0], p1[1])
        if p0[1] != p1[1]:
            dxf = _cdiv((c1[0] - c0[0]) << XY_SHIFT, c1[1] - c0[1])
            if p0[1] < p1[1]:
                edges.append((p0[1], p1[1], (c0[0] << XY_SHIFT) + (p0[1] - c0[1]) * dxf, dxf))
            else:
                edges.append((p1[1], p0[1], (c1[0] << XY_SHIFT) + (p1[1] - c1[1]) * dxf, dxf))
        p0 = p1
    if len(edges) < 2:
        return
    ymin = min(e[0] for e in edges)
    ymax = min(max(e[1] for e in edges), h)
    for y in range(max(ymin, 0), ymax):
        act = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
        for k in range(0, len(act) - 1, 2):
            xa = (act[k] + XY_ONE - 1) >> XY_SHIFT
            xb = act[k + 1] >> XY_SHIFT
            if xa < w and xb >= 0:
                img[y, max(xa, 0):min(xb, w - 1) + 1] = color


def corners_to_voxel(corners, shape, voxel_size, z_offset=0.0):
    """generating_train_bev.py:130-132: (3, 4) float64 car-space corners -> (4, 2) integer (x, y)."""
    from . import bev_oracle
    cv = bev_oracle.car_to_voxel_coords(np.asarray(corners, dtype=np.float64), shape, voxel_size, z_offset)
    return cv.transpose(1, 0)[:, :2].astype(np.intp)      # np.int0: C truncation


def draw_boxes(im, voxel_size, corners_list, class_colors, z_offset=0.0):
    """generating_train_bev.py:127-139 with the devkit objects unpacked: corners_list[i] is
    box.bottom_corners() (3, 4), class_colors[i] = classes.index(box.name) + 1.  Mutates im (H, W, 3)."""
    for corners, color in zip(corners_list, class_colors):
        if color == 0:
            raise Exception("Unknown class")
        pts = corners_to_voxel(corners, im.shape, voxel_size, z_offset)
        mask = np.zeros(im.shape[:2], dtype=np.uint8)
        fill_polygon(mask, pts, 1)
        im[mask == 1] = color
    return im
