"""PointPillars decoration / scatter oracle (numpy, CPU).  TEST INFRASTRUCTURE.

Restates the arithmetic of second/second/pytorch/models/pointpillars.py and
voxel_encoder.py in numpy float32.  PINNED: oracle/gen_golden.py executes the
reference's own classes (loaded by oracle/ref_loader.py) on seeded inputs and
stores their outputs under tests/golden/; tests/test_oracle_pillar.py checks
this restatement against them.

Channel layouts produced (T = max points per pillar, all multiplied by the
padding mask ``t < num[p]`` - voxel_encoder.py:27-48, pointpillars.py:226-231):

  variant "pfn"            pointpillars.py:203-231   [x y z f.. | cx cy cz | px py | (dist)]
  variant "old"            pointpillars.py:117-145   as "pfn" but channels 0,1 are
                           overwritten with px,py (the aliasing bug)
  variant "radius"         pointpillars.py:290-319   [r z f.. | cx cy cz | px py | (dist)]
  variant "radius_height"  pointpillars.py:378-411   [r z f.. | cx cy cz | px py | h | (dist)]
"""
import numpy as np

VARIANTS = ("pfn", "old", "radius", "radius_height")


def paddings_indicator(actual_num, max_num):
    """voxel_encoder.py:27-48 with axis=0: mask[p, t] = actual_num[p] > t."""
    return np.asarray(actual_num).astype(np.int32)[:, None] > np.arange(max_num, dtype=np.int32)[None, :]


def decorate(voxels, num_points, coors, voxel_size, pc_range, variant="pfn", with_distance=False):
    """Decoration of (P,T,C) pillars -> (P,T,C') float32.

    ``coors`` is (P,4) [batch, z, y, x] (preprocess.py:44-50).  The offsets are
    Python floats exactly as the module computes them (pointpillars.py:198-201):
    vx, vy and x_offset = vx/2 + pc_range[0] are *Python doubles*, multiplied into
    float32 tensors -> torch rounds the scalar to float32 first.
    """
    f32 = np.float32
    v = np.asarray(voxels, dtype=f32)
    P, T, C = v.shape
    num = np.asarray(num_points)
    vx, vy = float(voxel_size[0]), float(voxel_size[1])
    x_off = vx / 2 + float(pc_range[0])
    y_off = vy / 2 + float(pc_range[1])
    # pointpillars.py:208-210
    mean = v[:, :, :3].sum(axis=1, keepdims=True, dtype=f32) / num.astype(f32).reshape(-1, 1, 1)
    f_cluster = v[:, :, :3] - mean
    # pointpillars.py:213-217
    cx = coors[:, 3].astype(f32)[:, None] * f32(vx) + f32(x_off)
    cy = coors[:, 2].astype(f32)[:, None] * f32(vy) + f32(y_off)
    f_center = np.stack([v[:, :, 0] - cx, v[:, :, 1] - cy], axis=-1).astype(f32)
    if variant == "pfn":
        head = v
    elif variant == "old":
        head = v.copy()
        head[:, :, :2] = f_center  # pointpillars.py:127-134: f_center is a view of features
    elif variant in ("radius", "radius_height"):
        r = np.sqrt(v[:, :, 0] * v[:, :, 0] + v[:, :, 1] * v[:, :, 1], dtype=f32)[..., None]
        head = np.concatenate([r, v[:, :, 2:]], axis=-1)  # pointpillars.py:305-306
    else:
        raise ValueError(variant)
    parts = [head, f_cluster, f_center]
    if variant == "radius_height":
        # pointpillars.py:387-393 (min/max over all T slots, padding zeros included)
        h = v[:, :, 2:3].max(axis=1, keepdims=True) - v[:, :, 2:3].min(axis=1, keepdims=True)
        parts.append(np.broadcast_to(h, (P, T, 1)).astype(f32))
    if with_distance:
        # "old": the norm is taken AFTER x,y were overwritten in place (pointpillars.py:127-139)
        src = head if variant == "old" else v
        d = np.sqrt((src[:, :, :3] * src[:, :, :3]).sum(axis=2, dtype=f32), dtype=f32)[..., None]
        parts.append(d)
    feats = np.concatenate(parts, axis=-1).astype(f32)
    mask = paddings_indicator(num, T).astype(f32)[..., None]
    return feats * mask


def decorate_mismatch(out, ref, voxels, num_points, variant="pfn", with_distance=False):
    """Compares a decoration `out` with the reference `ref`; returns (ok, worst, max_abs):
      * channels that are copies or ONE subtraction (x, y, z, r, f_center, height): must be equal;
      * norms (radius, distance): 1e-6 relative (BASELINE.json north_star);
      * f_cluster = x - mean: the float32 sum over T is the one freedom of the op (torch's own CPU and CUDA sums differ
        the same way).  ANY two correctly rounded summation orders of n values differ by at most 2 (n-1) u sum|x|
        (u = 2^-24; Higham, Accuracy and Stability of Numerical Algorithms, (4.4)); the mean divides that by n, and the
        division and the subtraction add one ulp each.  ok requires every value inside this bound
            B = 2 u sum|x| (n-1)/n + ulp(sum|x| / n) + ulp(max|x|)
        `worst` is the largest error as a fraction of B, `max_abs` the largest absolute f_cluster error.  Measured
        against this numpy oracle over frames 0..1279 of the C5 set (tools/verify_frames.py; every other output bit-exact): max_abs
        1.14e-5 (e.g. a pillar of 14 points 46.9 m out; a fixed atol of 1e-5 fails on such pillars), worst 0.70 of B."""
    out, ref = np.asarray(out), np.asarray(ref)
    if out.shape != ref.shape:
        return False, float("inf"), float("inf")
    k = 4 if variant in ("pfn", "old") else 3
    cl = [k, k + 1, k + 2]
    norms = ([0] if variant in ("radius", "radius_height") else []) + ([ref.shape[2] - 1] if with_distance else [])
    rest = [c for c in range(ref.shape[2]) if c not in cl and c not in norms]
    ok = bool(np.array_equal(out[..., rest], ref[..., rest]))
    if norms:
        ok = ok and bool(np.allclose(out[..., norms], ref[..., norms], rtol=1e-6, atol=0))
    num = np.maximum(np.asarray(num_points), 1).astype(np.float64)[:, None]
    mask = np.arange(ref.shape[1])[None, :] < np.asarray(num_points)[:, None]
    v3 = np.abs(np.asarray(voxels, dtype=np.float32)[..., :3])
    ssum = v3.sum(axis=1, dtype=np.float64)
    bound = (2.0 ** -23 * ssum * (num - 1) / num + np.spacing((ssum / num).astype(np.float32)) + np.spacing(v3.max(axis=1)))[:, None, :]
    err = np.abs(out[..., cl].astype(np.float64) - ref[..., cl].astype(np.float64))
    worst = float((err / bound)[mask].max()) if mask.any() else 0.0
    max_abs = float(err[mask].max()) if mask.any() else 0.0
    pad_ok = bool((out[..., cl][~mask] == ref[..., cl][~mask]).all())
    return ok and pad_ok and worst <= 1.0, worst, max_abs


def scatter(voxel_features, coords, batch_size, ny, nx):
    """PointPillarsScatter.forward, pointpillars.py:444-476 -> (B,C,ny,nx)."""
    feats = np.asarray(voxel_features)
    C = feats.shape[1]
    out = np.zeros((batch_size, C, ny * nx), dtype=feats.dtype)
    for b in range(batch_size):
        sel = coords[:, 0] == b
        idx = coords[sel, 2].astype(np.int64) * nx + coords[sel, 3].astype(np.int64)
        out[b][:, idx] = feats[sel].T
    return out.reshape(batch_size, C, ny, nx)


def simple_voxel_mean(voxels, num_points, num_input_features=4):
    """SimpleVoxel.forward, voxel_encoder.py:219-225."""
    v = np.asarray(voxels, dtype=np.float32)
    return v[:, :, :num_input_features].sum(axis=1, dtype=np.float32) / \
        np.asarray(num_points).astype(np.float32).reshape(-1, 1)


def pfn_layer_eval(features, weight, bn_gamma, bn_beta, bn_mean, bn_var, eps=1e-3):
    """PFNLayer.forward (last layer, use_norm, eval mode), pointpillars.py:51-65:
    Linear(no bias) -> BatchNorm1d(eps=1e-3, running stats) -> ReLU -> max over T."""
    x = features.astype(np.float32) @ weight.T.astype(np.float32)
    x = (x - bn_mean) / np.sqrt(bn_var + np.float32(eps)) * bn_gamma + bn_beta
    x = np.maximum(x, 0)
    return x.max(axis=1).astype(np.float32)


def merge_batch_coords(coords_list):
    """merge_second_batch coordinate padding, second/second/data/preprocess.py:44-50:
    prepend the batch index -> (sum V, 4) [b, z, y, x]."""
    out = []
    for i, c in enumerate(coords_list):
        out.append(np.pad(c, ((0, 0), (1, 0)), mode="constant", constant_values=i))
    return np.concatenate(out, axis=0)


def decorate_half(voxels_h, num_points, coors, voxel_size, pc_range, variant="pfn", with_distance=False):
    """The same decoration on float16 pillars (apex O2: second/second/pytorch/train.py:34-47 casts "voxels" to
    half, configs/nuscenes/all.pp.mida.config:348): every torch op on half tensors computes in float32 and rounds
    its RESULT to half, so each statement of pointpillars.py:203-231 contributes one rounding:
        mean   = rh(rh(sum_T x) / num)              (:208-209, the sum accumulates in float32)
        f_clus = rh(x - mean)                       (:210)
        centre = rh(rh(coor * vx) + x_offset)       (:213-217; coors.to(half) is exact below 2048)
        f_cent = rh(x - centre)
        radius = rh(sqrt(x^2 + y^2)), dist = rh(sqrt(x^2 + y^2 + z^2))   (torch.norm accumulates in float32)
    Returns float16 (P,T,C')."""
    f32, f16 = np.float32, np.float16

    def rh(a):
        return np.asarray(a, dtype=f32).astype(f16).astype(f32)

    v = np.asarray(voxels_h, dtype=f16).astype(f32)
    P, T, C = v.shape
    num = np.asarray(num_points)
    vx, vy = float(voxel_size[0]), float(voxel_size[1])
    x_off = vx / 2 + float(pc_range[0])
    y_off = vy / 2 + float(pc_range[1])
    mean = rh(rh(v[:, :, :3].sum(axis=1, keepdims=True, dtype=f32)) / rh(num.astype(f32)).reshape(-1, 1, 1))
    f_cluster = rh(v[:, :, :3] - mean)
    cx = rh(rh(coors[:, 3].astype(f32)[:, None] * f32(vx)) + f32(x_off))
    cy = rh(rh(coors[:, 2].astype(f32)[:, None] * f32(vy)) + f32(y_off))
    f_center = rh(np.stack([v[:, :, 0] - cx, v[:, :, 1] - cy], axis=-1))
    if variant == "pfn":
        head = v
    elif variant == "old":
        head = v.copy()
        head[:, :, :2] = f_center
    elif variant in ("radius", "radius_height"):
        r = rh(np.sqrt(v[:, :, 0] * v[:, :, 0] + v[:, :, 1] * v[:, :, 1], dtype=f32))[..., None]
        head = np.concatenate([r, v[:, :, 2:]], axis=-1)
    else:
        raise ValueError(variant)
    parts = [head, f_cluster, f_center]
    if variant == "radius_height":
        h = rh(v[:, :, 2:3].max(axis=1, keepdims=True) - v[:, :, 2:3].min(axis=1, keepdims=True))
        parts.append(np.broadcast_to(h, (P, T, 1)).astype(f32))
    if with_distance:
        src = head if variant == "old" else v
        parts.append(rh(np.sqrt((src[:, :, :3] * src[:, :, :3]).sum(axis=2, dtype=f32), dtype=f32))[..., None])
    feats = np.concatenate(parts, axis=-1).astype(f32)
    mask = paddings_indicator(num, T).astype(f32)[..., None]
    return (feats * mask).astype(f16)
