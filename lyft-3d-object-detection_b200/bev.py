"""BEV rasteriser drop-ins (reference: generating-dataset/generating_train_bev.py).

Same names, arguments and error behaviour as the closures nested in the
reference's ``generating_train_bev()`` (:47-104; copies in
generating_test_bev.py:155-212 and unet_baseline/unet-inference-with-map.py:368-425):

    create_transformation_matrix_to_voxel_space(shape, voxel_size, offset)
    transform_points(points, transf_matrix)
    car_to_voxel_coords(points, shape, voxel_size, z_offset=0)
    create_voxel_pointcloud(points, shape, voxel_size=(0.5,0.5,1), z_offset=0)
    normalize_voxel_intensities(bev, max_intensity=16)

``points`` is (3,N) or (4,N) float32 as produced by ``LidarPointCloud.points``.
numpy in -> numpy out (host buffers cross PCIe inside the call); a CUDA torch
tensor in -> CUDA tensor out on the current stream, no synchronisation.  All
arithmetic runs in the sm_100a kernels of liblyftvoxel_b200.so; there is no CPU
path.  ``rasterize_frames`` is the batched form used for throughput (many frames,
optional per-sweep sensor->car transforms, optional u8 / CHW+map outputs).
"""
import ctypes

import numpy as np

from . import _native as nat


def create_transformation_matrix_to_voxel_space(shape, voxel_size, offset):
    """generating_train_bev.py:47-62.  Host-side 4x4 (float64: the reference's
    float32 eye is promoted by the float64 scale row)."""
    shape = np.asarray(shape)
    voxel_size = np.asarray(voxel_size)
    offset = np.asarray(offset)
    scale = np.hstack((1 / voxel_size, [1]))
    tm = np.diag(scale).astype(np.result_type(np.float32, scale.dtype))
    tm[:3, 3] = shape / 2 + offset / voxel_size
    return tm


def _is_cuda_tensor(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def _rows_from_points(points):
    """(3|4, N) points -> (N, C) float32 C-contiguous rows (numpy or CUDA tensor)."""
    if _is_cuda_tensor(points):
        import torch
        if points.dim() != 2 or points.shape[0] not in (3, 4):
            raise Exception("Input points should be (3,N) or (4,N) in shape, found {}".format(tuple(points.shape)))
        if points.dtype != torch.float32:
            raise Exception("points must be float32, found {}".format(points.dtype))
        return points.t().contiguous()
    points = np.asarray(points)
    if points.ndim != 2 or points.shape[0] not in (3, 4):
        raise Exception("Input points should be (3,N) or (4,N) in shape, found {}".format(points.shape))
    if points.dtype != np.float32:
        raise Exception("points must be float32 (LidarPointCloud.points), found {}".format(points.dtype))
    return np.ascontiguousarray(points.T)


def _shape3(shape):
    if len(shape) != 3:
        raise Exception("Voxel volume shape should be 3 dimensions (x,y,z)")
    return (ctypes.c_int32 * 3)(*[int(s) for s in shape])


def _vs3(voxel_size):
    vs = np.asarray(voxel_size, dtype=np.float64)
    if vs.shape != (3,):
        raise Exception("voxel_size should have 3 entries (x,y,z)")
    return (ctypes.c_double * 3)(*vs.tolist())


def rasterize_frames(points_rows, frame_offsets, shape, voxel_size, z_offset=0.0, max_intensity=16.0,
                     seg_frame=None, seg_tm=None, n_frames=None, want=("raw",), map_u8=None, out=None,
                     handle=None):
    """Batched BEV rasterisation.

    points_rows   (N_total, C>=3) float32 rows, numpy (host) or CUDA tensor.
    frame_offsets int64 (S+1) segment offsets into the rows (a segment is one
                  sweep; by default segment s is frame s).
    seg_frame     optional int32 (S) frame id of each segment (non-decreasing).
    seg_tm        optional float64 (S,4,4) sensor->car transform per segment.
    want          subset of {"raw", "norm", "u8", "chw"}; "chw" needs map_u8
                  (n_frames, S0, S1, 3) uint8.
    out           optional dict of preallocated outputs (same keys as `want`).
    Returns a dict keyed like `want`; arrays are (n_frames, S0, S1, S2) or
    (n_frames, S2+3, S0, S1) for "chw".
    """
    lib = nat.load()
    offs = np.ascontiguousarray(frame_offsets, dtype=np.int64)
    n_seg = offs.shape[0] - 1
    if n_frames is None:
        n_frames = n_seg if seg_frame is None else (int(np.max(seg_frame)) + 1 if n_seg else 0)
    shp = _shape3(shape)
    vs = _vs3(voxel_size)
    S0, S1, S2 = int(shape[0]), int(shape[1]), int(shape[2])
    sf = None if seg_frame is None else np.ascontiguousarray(seg_frame, dtype=np.int32)
    tm = None if seg_tm is None else np.ascontiguousarray(seg_tm, dtype=np.float64).reshape(n_seg, 16)
    out = {} if out is None else out
    cuda = _is_cuda_tensor(points_rows)
    res = {}
    if cuda:
        import torch
        dev = points_rows.device
        h = handle or nat.get_handle(dev.index)
        if points_rows.dtype != torch.float32 or not points_rows.is_contiguous():
            raise Exception("points rows must be a contiguous float32 tensor")
        stride = points_rows.shape[1]

        def alloc(key, shp_, dt):
            if key in out:
                return out[key]
            return torch.empty(shp_, dtype=dt, device=dev)
        if "raw" in want:
            res["raw"] = alloc("raw", (n_frames, S0, S1, S2), torch.float32)
        if "norm" in want:
            res["norm"] = alloc("norm", (n_frames, S0, S1, S2), torch.float32)
        if "u8" in want:
            res["u8"] = alloc("u8", (n_frames, S0, S1, S2), torch.uint8)
        if "chw" in want:
            if map_u8 is None:
                raise Exception("chw output needs map_u8")
            res["chw"] = alloc("chw", (n_frames, S2 + 3, S0, S1), torch.float32)
        with torch.cuda.device(dev):
            nat.check(lib.lv_bev_rasterize(
                h.ptr, points_rows.data_ptr(), stride, n_seg, offs.ctypes.data, nat.dptr(sf), nat.dptr(tm),
                n_frames, shp, vs, float(z_offset), float(max_intensity), nat.dptr(res.get("raw")),
                nat.dptr(res.get("norm")), nat.dptr(res.get("u8")), nat.dptr(map_u8), nat.dptr(res.get("chw")),
                nat.current_stream_ptr(dev)))
        return res
    rows = np.ascontiguousarray(points_rows, dtype=np.float32)
    h = handle or nat.get_handle()
    stride = rows.shape[1] if rows.ndim == 2 else 4

    def alloc(key, shp_, dt):
        if key in out:
            return out[key]
        # results land in pinned host memory (blocks of torch's caching host allocator): one DMA instead of
        # the driver's pageable staging
        import torch
        n = int(np.prod(shp_))
        if n == 0:
            return np.empty(shp_, dtype=dt)
        tdt = {np.float32: torch.float32, np.uint8: torch.uint8}[dt]
        return torch.empty((n,), dtype=tdt, pin_memory=True).numpy().reshape(shp_)
    if "raw" in want:
        res["raw"] = alloc("raw", (n_frames, S0, S1, S2), np.float32)
    if "norm" in want:
        res["norm"] = alloc("norm", (n_frames, S0, S1, S2), np.float32)
    if "u8" in want:
        res["u8"] = alloc("u8", (n_frames, S0, S1, S2), np.uint8)
    if "chw" in want:
        if map_u8 is None:
            raise Exception("chw output needs map_u8")
        map_u8 = np.ascontiguousarray(map_u8, dtype=np.uint8)
        res["chw"] = alloc("chw", (n_frames, S2 + 3, S0, S1), np.float32)
    nat.check(lib.lv_bev_rasterize_host(
        h.ptr, rows.ctypes.data, stride, n_seg, offs.ctypes.data, nat.dptr(sf), nat.dptr(tm), n_frames, shp, vs,
        float(z_offset), float(max_intensity), nat.dptr(res.get("raw")), nat.dptr(res.get("norm")),
        nat.dptr(res.get("u8")), nat.dptr(map_u8), nat.dptr(res.get("chw"))))
    return res


def transform_points(points, transf_matrix):
    """generating_train_bev.py:64-70: (3,N)|(4,N) float32 points -> (3,N) float64."""
    shp = tuple(points.shape)
    if len(shp) != 2 or shp[0] not in [3, 4]:
        raise Exception("Points input should be (3,N) or (4,N) shape, received {}".format(shp))
    import torch
    lib = nat.load()
    rows = _rows_from_points(points)
    tm = np.ascontiguousarray(np.asarray(transf_matrix, dtype=np.float64).reshape(4, 4))
    cuda = _is_cuda_tensor(rows)
    d_rows = rows if cuda else torch.from_numpy(rows).cuda()
    n = d_rows.shape[0]
    out = torch.empty((3, n), dtype=torch.float64, device=d_rows.device)
    h = nat.get_handle(d_rows.device.index)
    with torch.cuda.device(d_rows.device):
        nat.check(lib.lv_transform_points(h.ptr, d_rows.data_ptr(), d_rows.shape[1], n, tm.ctypes.data,
                                          out.data_ptr(), nat.current_stream_ptr(d_rows.device)))
    return out if cuda else out.cpu().numpy()


def car_to_voxel_coords(points, shape, voxel_size, z_offset=0):
    """generating_train_bev.py:73-82."""
    if len(shape) != 3:
        raise Exception("Voxel volume shape should be 3 dimensions (x,y,z)")
    if len(points.shape) != 2 or points.shape[0] not in [3, 4]:
        raise Exception("Input points should be (3,N) or (4,N) in shape, found {}".format(tuple(points.shape)))
    tm = create_transformation_matrix_to_voxel_space(shape, voxel_size, (0, 0, z_offset))
    return transform_points(points, tm)


def create_voxel_pointcloud(points, shape, voxel_size=(0.5, 0.5, 1), z_offset=0):
    """generating_train_bev.py:84-101 -> (shape) float32 raw counts, bev[y, x, z]."""
    if len(shape) != 3:
        raise Exception("Voxel volume shape should be 3 dimensions (x,y,z)")
    rows = _rows_from_points(points)
    n = rows.shape[0]
    res = rasterize_frames(rows, np.array([0, n], dtype=np.int64), shape, voxel_size, z_offset, want=("raw",))
    return res["raw"][0]


def normalize_voxel_intensities(bev, max_intensity=16):
    """generating_train_bev.py:103-104: (bev / max_intensity).clip(0, 1)."""
    lib = nat.load()
    if _is_cuda_tensor(bev):
        import torch
        x = bev.contiguous()
        if x.dtype != torch.float32:
            raise Exception("bev must be float32")
        y = torch.empty_like(x)
        h = nat.get_handle(x.device.index)
        with torch.cuda.device(x.device):
            nat.check(lib.lv_bev_normalize(h.ptr, x.data_ptr(), x.numel(), float(max_intensity), y.data_ptr(),
                                           nat.current_stream_ptr(x.device)))
        return y
    import torch
    x = np.ascontiguousarray(bev, dtype=np.float32)
    h = nat.get_handle()
    dx = torch.from_numpy(x).cuda()
    dy = torch.empty_like(dx)
    nat.check(lib.lv_bev_normalize(h.ptr, dx.data_ptr(), dx.numel(), float(max_intensity), dy.data_ptr(),
                                   nat.current_stream_ptr()))
    return dy.cpu().numpy().reshape(np.shape(bev))


def quantize_u8(points, shape, voxel_size=(0.5, 0.5, 1), z_offset=0, max_intensity=16):
    """create_voxel_pointcloud -> normalize -> np.round(bev*255).astype(uint8)
    (generating_train_bev.py:210-213) fused in one call."""
    rows = _rows_from_points(points)
    res = rasterize_frames(rows, np.array([0, rows.shape[0]], dtype=np.int64), shape, voxel_size, z_offset,
                           max_intensity, want=("u8",))
    return res["u8"][0]


def png_slot_bytes(height, width, channels):
    """Slot size that always holds the PNG of a (height, width, channels) uint8 image."""
    n = int(nat.load().lv_png_max_bytes(int(height), int(width), int(channels)))
    if n <= 0:
        raise Exception("PNG images are (H,W) or (H,W,3) uint8")
    return (n + 255) // 256 * 256


def encode_png_frames(images, out=None, sizes=None, stride=None, swap_rb=True, handle=None):
    """Device-side PNG encoding of a batch: images (F,H,W) or (F,H,W,3) uint8 CUDA tensor ->
    (out, sizes): frame f's file is out[f, :sizes[f]].  `out` (F, stride) uint8 and `sizes` (F) int32 may be
    CUDA tensors or tensors over mapped pinned host memory (engine.MappedBuffer): the kernels then write the
    files straight into host memory.  Asynchronous on torch's current stream.
    cv2.imwrite of generating_train_bev.py:215,224 (B,G,R arrays are written R,G,B: swap_rb)."""
    import torch
    lib = nat.load()
    if not (_is_cuda_tensor(images) and images.dtype == torch.uint8 and images.dim() in (3, 4)):
        raise Exception("images must be a (F,H,W) or (F,H,W,3) uint8 CUDA tensor")
    images = images.contiguous()
    F, H, W = int(images.shape[0]), int(images.shape[1]), int(images.shape[2])
    ch = 1 if images.dim() == 3 else int(images.shape[3])
    if ch not in (1, 3):
        raise Exception("PNG images have 1 or 3 channels, got %d" % ch)
    if stride is None:
        stride = int(out.shape[1]) if out is not None else png_slot_bytes(H, W, ch)
    if out is None:
        out = torch.empty((F, stride), dtype=torch.uint8, device=images.device)
    if sizes is None:
        sizes = torch.empty((F,), dtype=torch.int32, device=images.device)
    h = handle or nat.get_handle(images.device.index)
    with torch.cuda.device(images.device):
        nat.check(lib.lv_png_encode(h.ptr, images.data_ptr(), F, H, W, ch, int(bool(swap_rb)), out.data_ptr(), int(stride),
                                    sizes.data_ptr(), nat.current_stream_ptr(images.device)))
    return out, sizes


def imencode_png(image):
    """cv2.imencode('.png', image)[1].tobytes() for a uint8 (H,W) / (H,W,3) array (numpy or CUDA tensor), or a
    list of files for a batch (F,H,W,3) / a CUDA (F,H,W) tensor given as `image[None]`.  Lossless: decoding with
    cv2.imread / cv2.imdecode returns the array."""
    import torch
    single = False
    if _is_cuda_tensor(image):
        t = image
        if t.dim() == 2 or (t.dim() == 3 and t.shape[2] == 3):
            t, single = t[None], True
        out, sizes = encode_png_frames(t)
        sizes = sizes.cpu().numpy()
        out = out.cpu().numpy()
    else:
        a = np.ascontiguousarray(image)
        if a.dtype != np.uint8:
            raise Exception("PNG images must be uint8, got %s" % a.dtype)
        if a.ndim == 2 or (a.ndim == 3 and a.shape[2] in (1, 3)):
            a, single = a[None], True
        if a.ndim == 4 and a.shape[3] == 1:
            a = a[..., 0]
        if a.ndim not in (3, 4) or (a.ndim == 4 and a.shape[3] != 3):
            raise Exception("PNG images are (H,W), (H,W,3) or a batch of them")
        F, H, W = a.shape[:3]
        ch = 1 if a.ndim == 3 else 3
        lib = nat.load()
        stride = png_slot_bytes(H, W, ch)
        out = np.empty((F, stride), np.uint8)
        sizes = np.empty((F,), np.int32)
        nat.check(lib.lv_png_encode_host(nat.get_handle().ptr, a.ctypes.data, F, H, W, ch, 1, out.ctypes.data, stride,
                                         sizes.ctypes.data))
    files = [out[f, :int(sizes[f])].tobytes() for f in range(len(sizes))]
    return files[0] if single else files


def imwrite(path, image):
    """cv2.imwrite(path, image) for 8-bit PNG files (generating_train_bev.py:215,224,229)."""
    if not str(path).lower().endswith(".png"):
        raise Exception("imwrite: only .png is produced, got %r" % (path,))
    data = imencode_png(image)
    with open(path, "wb") as f:
        f.write(data)
    return True


def write_training_pngs(sweeps, car_from_sensor, corners, colors, box_offsets, tokens, output_folder,
                        bev_shape=(336, 336, 3), voxel_size=(0.4, 0.4, 1.5), z_offset=-2.0, max_intensity=16,
                        device=0):
    """The per-sample body of ``prepare_training_data_for_scene`` (generating_train_bev.py:208-224) for a whole
    list of samples in one pass on the GPU: for sample i

        lidar_pointcloud = LidarPointCloud.from_file(...)            sweeps[i]: the (N_i, 5) float32 rows of the .bin
        lidar_pointcloud.transform(car_from_sensor)                  car_from_sensor[i]: (4, 4) float64
        bev = normalize_voxel_intensities(create_voxel_pointcloud(...));  bev_im = np.round(bev*255).astype(np.uint8)
        cv2.imwrite("{token}_input.png", bev_im)
        draw_boxes(target, ...); cv2.imwrite("{token}_target.png", target[:, :, 0])

    `corners` (n_boxes, 3, 4) float64 are the ``bottom_corners()`` of the boxes after ``move_boxes_to_car_space`` and
    ``scale_boxes``, `colors` (n_boxes,) = ``classes.index(box.name) + 1``, `box_offsets` (len(sweeps) + 1) says which
    boxes belong to which sample.  Points, boxes and the encoded FILES are all that crosses PCIe.  Returns the list
    of (input_path, target_path).  The semantic-map image (:226-229) is not produced here."""
    import os
    import torch
    n = len(sweeps)
    if not (len(tokens) == n and len(car_from_sensor) == n and len(box_offsets) == n + 1):
        raise Exception("sweeps, car_from_sensor, tokens and box_offsets disagree")
    if n == 0:
        return []
    dev = torch.device("cuda", device)
    rows = np.concatenate([np.ascontiguousarray(s, dtype=np.float32).reshape(-1, 5) for s in sweeps], axis=0)
    offs = np.concatenate([[0], np.cumsum([np.asarray(s).reshape(-1, 5).shape[0] for s in sweeps])]).astype(np.int64)
    tm = np.ascontiguousarray(np.stack([np.asarray(m, dtype=np.float64).reshape(4, 4) for m in car_from_sensor]))
    with torch.cuda.device(dev):
        d_rows = torch.from_numpy(rows).to(dev)
        res = rasterize_frames(d_rows, offs, bev_shape, voxel_size, z_offset, max_intensity, seg_tm=tm, want=("u8",))
        in_files, in_sizes = encode_png_frames(res["u8"])
        d_c = torch.from_numpy(np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 3, 4)).to(dev)
        d_k = torch.from_numpy(np.ascontiguousarray(colors, dtype=np.int32)).to(dev)
        tgt = rasterize_targets(d_c, d_k, box_offsets, bev_shape, voxel_size, z_offset)
        tg_files, tg_sizes = encode_png_frames(tgt)
        in_sizes, tg_sizes = in_sizes.cpu().numpy(), tg_sizes.cpu().numpy()
        in_max, tg_max = int(in_sizes.max()), int(tg_sizes.max())
        in_files = in_files[:, :in_max].cpu().numpy()        # only the used part of the slots comes back
        tg_files = tg_files[:, :tg_max].cpu().numpy()
    out = []
    for i, tok in enumerate(tokens):
        a = os.path.join(output_folder, "{}_input.png".format(tok))
        b = os.path.join(output_folder, "{}_target.png".format(tok))
        with open(a, "wb") as f:
            f.write(in_files[i, :in_sizes[i]].tobytes())
        with open(b, "wb") as f:
            f.write(tg_files[i, :tg_sizes[i]].tobytes())
        out.append((a, b))
    return out


def rasterize_targets(corners, colors, box_offsets, shape, voxel_size, z_offset=0.0, handle=None):
    """Batched target rasterisation (lv_draw_boxes): frame f paints boxes
    [box_offsets[f], box_offsets[f+1]) in order, later over earlier.

    corners (n_boxes, 3, 4) float64 car-space Box.bottom_corners(), colors (n_boxes,) int32
    (classes.index(name) + 1); numpy or CUDA tensors.  Returns uint8 (F, shape[0], shape[1]) =
    ``target[:, :, 0]`` of generating_train_bev.py:214-221 per frame (same kind as the input)."""
    import ctypes
    lib = nat.load()
    offs = np.ascontiguousarray(box_offsets, dtype=np.int64)
    F = offs.shape[0] - 1
    shp = (ctypes.c_int32 * 3)(*_shape3(shape))
    vs = (ctypes.c_double * 3)(*_vs3(voxel_size))
    H, W = int(shape[0]), int(shape[1])
    if _is_cuda_tensor(corners):
        import torch
        c = corners.contiguous()
        k = colors.contiguous()
        if c.dtype != torch.float64 or k.dtype != torch.int32 or tuple(c.shape[1:]) != (3, 4):
            raise Exception("corners must be (n, 3, 4) float64 and colors (n,) int32")
        out = torch.empty((F, H, W), dtype=torch.uint8, device=c.device)
        h = handle or nat.get_handle(c.device.index)
        with torch.cuda.device(c.device):
            nat.check(lib.lv_draw_boxes(h.ptr, c.data_ptr(), k.data_ptr(), F, offs.ctypes.data, shp, vs,
                                        float(z_offset), out.data_ptr(), nat.current_stream_ptr(c.device)))
        return out
    c = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 3, 4)
    k = np.ascontiguousarray(colors, dtype=np.int32)
    if c.shape[0] != k.shape[0] or (F > 0 and offs[-1] != c.shape[0]):
        raise Exception("corners, colors and box_offsets disagree")
    out = np.empty((F, H, W), dtype=np.uint8)
    h = handle or nat.get_handle()
    nat.check(lib.lv_draw_boxes_host(h.ptr, c.ctypes.data, k.ctypes.data, F, offs.ctypes.data, shp, vs,
                                     float(z_offset), out.ctypes.data))
    return out


def draw_boxes(im, voxel_size, boxes, classes, z_offset=0.0):
    """generating_train_bev.py:127-139, same signature and side effect: paints every box's bottom
    footprint into `im` (H, W, 3) in place with the value classes.index(box.name) + 1.  `boxes` are
    devkit Box objects (anything with ``bottom_corners()`` -> (3, 4) and ``name``).  An unknown class
    name raises, as ``classes.index`` does in the reference."""
    boxes = list(boxes)
    if not boxes:
        return
    colors = np.array([classes.index(b.name) + 1 for b in boxes], dtype=np.int32)
    corners = np.stack([np.asarray(b.bottom_corners(), dtype=np.float64) for b in boxes])
    target = rasterize_targets(corners, colors, np.array([0, len(boxes)], dtype=np.int64), im.shape, voxel_size,
                               z_offset)[0]
    painted = target > 0
    im[painted] = target[painted].astype(im.dtype)[:, None]
