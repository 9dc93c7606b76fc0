"""BEV rasteriser oracle (numpy, CPU).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates, function by function, the closures nested in
``generating_train_bev()`` (reference: generating-dataset/generating_train_bev.py:47-104;
identical copies: generating-dataset/generating_test_bev.py:155-212 and
unet_baseline/unet-inference-with-map.py:368-425).  They cannot be imported
(nested closures) so they are restated with the one change the
container forces: ``np.int0`` (removed in NumPy 2) -> ``np.intp`` - the same C
cast (truncation toward zero).

Semantics that matter:
  * the 4x4 is float64 (float32 eye * float64 row -> float64);
  * the voxel coordinate is TRUNCATED (C cast), not floored;
  * bounds: x against shape[0], y against shape[1], z against shape[2];
    the write is bev[y, x, z] (square grids only, F8).
"""
import numpy as np


def create_transformation_matrix_to_voxel_space(shape, voxel_size, offset):
    """generating_train_bev.py:47-62."""
    shape, voxel_size, offset = np.array(shape), np.array(voxel_size), np.array(offset)
    tm = np.eye(4, dtype=np.float32)
    translation = shape / 2 + offset / voxel_size
    tm = tm * np.array(np.hstack((1 / voxel_size, [1])))
    tm[:3, 3] = np.transpose(translation)
    return tm


def transform_points(points, transf_matrix):
    """generating_train_bev.py:64-70."""
    if points.shape[0] not in [3, 4]:
        raise Exception("Points input should be (3,N) or (4,N) shape, received {}".format(points.shape))
    return transf_matrix.dot(np.vstack((points[:3, :], np.ones(points.shape[1]))))[:3, :]


def car_to_voxel_coords(points, shape, voxel_size, z_offset=0):
    """generating_train_bev.py:73-82."""
    if len(shape) != 3:
        raise Exception("Voxel volume shape should be 3 dimensions (x,y,z)")
    if len(points.shape) != 2 or points.shape[0] not in [3, 4]:
        raise Exception("Input points should be (3,N) or (4,N) in shape, found {}".format(points.shape))
    tm = create_transformation_matrix_to_voxel_space(shape, voxel_size, (0, 0, z_offset))
    p = transform_points(points, tm)
    return p


def create_voxel_pointcloud(points, shape, voxel_size=(0.5, 0.5, 1), z_offset=0):
    """generating_train_bev.py:84-101 (np.int0 -> np.intp)."""
    points_voxel_coords = car_to_voxel_coords(points.copy(), shape, voxel_size, z_offset)
    points_voxel_coords = points_voxel_coords[:3].transpose(1, 0)
    with np.errstate(invalid="ignore"):
        points_voxel_coords = np.intp(points_voxel_coords)

    bev = np.zeros(shape, dtype=np.float32)
    bev_shape = np.array(shape)

    within_bounds = (np.all(points_voxel_coords >= 0, axis=1) * np.all(points_voxel_coords < bev_shape, axis=1))

    points_voxel_coords = points_voxel_coords[within_bounds]
    coord, count = np.unique(points_voxel_coords, axis=0, return_counts=True)

    # Note X and Y are flipped:
    bev[coord[:, 1], coord[:, 0], coord[:, 2]] = count
    return bev


def normalize_voxel_intensities(bev, max_intensity=16):
    """generating_train_bev.py:103-104."""
    return (bev / max_intensity).clip(0, 1)


def quantize_u8(bev_norm):
    """generating_train_bev.py:213 - np.round(bev*255).astype(np.uint8)."""
    return np.round(bev_norm * 255).astype(np.uint8)


def sensor_to_car(points, transf_matrix):
    """PointCloud.transform, nuscenes-devkit/lyft_dataset_sdk/utils/data_classes.py:188-195.

    ``points`` is the (4,N) float32 array of LidarPointCloud; rows 0..2 are
    overwritten with the float64 product rounded to float32 on store.  Returns
    a new array (the reference mutates in place).
    """
    out = np.array(points, dtype=np.float32, copy=True)
    tm = np.asarray(transf_matrix)
    out[:3, :] = tm.dot(np.vstack((out[:3, :], np.ones(out.shape[1]))))[:3, :]
    return out


def lidar_from_file(path):
    """LidarPointCloud.from_file, data_classes.py:269-284 -> (4,N) float32 view."""
    scan = np.fromfile(str(path), dtype=np.float32)
    return scan.reshape((-1, 5))[:, :4].T


def car_to_voxel_coords_elementwise(points, shape, voxel_size, z_offset=0):
    """the order-pinned definition of the voxel-space
    affine.  u_k = fl(fl(m_k * p_k) + t_k) in float64 with m_k = 1/vs_k computed
    first - an un-fused multiply then add.  tests assert this equals the BLAS
    ``np.dot`` of ``car_to_voxel_coords`` bit for bit; the CUDA kernel implements
    this form (``__dmul_rn`` / ``__dadd_rn``).
    """
    shape_a, vs = np.array(shape), np.array(voxel_size)
    tm = create_transformation_matrix_to_voxel_space(shape_a, vs, (0, 0, z_offset))
    p = np.asarray(points[:3, :], dtype=np.float64)
    out = np.empty_like(p)
    for k in range(3):
        out[k] = tm[k, k] * p[k] + tm[k, 3]
    return out


def bev_concat_map_chw(bev_u8, map_u8):
    """BEVImageDataset.__getitem__, deeplab_v3_baseline/dataset/dataset.py:83-106
    (the arithmetic only): concat on channel axis, /255 in float32, HWC->CHW."""
    im = np.concatenate((bev_u8, map_u8), axis=2)
    im = im.astype(np.float32) / 255
    return np.ascontiguousarray(im.transpose(2, 0, 1))
