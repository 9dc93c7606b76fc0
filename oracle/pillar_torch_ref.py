"""The reference's GPU path for decoration and scatter, restated op for op in eager PyTorch.
TEST INFRASTRUCTURE / bench.py baseline leg only (see oracle/__init__.py).

In the reference these two steps run ON THE GPU as chains of eager torch ops, called from
VoxelNet.network_forward (second/second/pytorch/models/voxelnet.py:327-333):
  decorate(...)  = PillarFeatureNet.forward up to (not including) the PFN layers, pointpillars.py:203-231
  scatter(...)   = PointPillarsScatter.forward, pointpillars.py:444-476
This file is what a maintainer of the reference sees today on the same B200; bench.py times it beside
lv_pillar_decorate / lv_pillar_scatter.  tests/test_oracle_pillar.py pins it against the outputs of the
reference's own classes (tests/golden/ref_pillar_decorate.npz, ref_scatter.npz)."""
import torch


def get_paddings_indicator(actual_num, max_num, axis=0):
    """voxel_encoder.py:27-48."""
    actual_num = torch.unsqueeze(actual_num, axis + 1)
    max_num_shape = [1] * len(actual_num.shape)
    max_num_shape[axis + 1] = -1
    max_num = torch.arange(max_num, dtype=torch.int, device=actual_num.device).view(max_num_shape)
    paddings_indicator = actual_num.int() > max_num
    return paddings_indicator


def decorate(features, num_voxels, coors, vx, vy, x_offset, y_offset, with_distance=False):
    """pointpillars.py:203-231 (PillarFeatureNet) without the PFN layers."""
    dtype = features.dtype
    points_mean = features[:, :, :3].sum(dim=1, keepdim=True) / num_voxels.type_as(features).view(-1, 1, 1)
    f_cluster = features[:, :, :3] - points_mean
    f_center = torch.zeros_like(features[:, :, :2])
    f_center[:, :, 0] = features[:, :, 0] - (coors[:, 3].to(dtype).unsqueeze(1) * vx + x_offset)
    f_center[:, :, 1] = features[:, :, 1] - (coors[:, 2].to(dtype).unsqueeze(1) * vy + y_offset)
    features_ls = [features, f_cluster, f_center]
    if with_distance:
        points_dist = torch.norm(features[:, :, :3], 2, 2, keepdim=True)
        features_ls.append(points_dist)
    features = torch.cat(features_ls, dim=-1)
    voxel_count = features.shape[1]
    mask = get_paddings_indicator(num_voxels, voxel_count, axis=0)
    mask = torch.unsqueeze(mask, -1).type_as(features)
    features *= mask
    return features


def scatter(voxel_features, coords, batch_size, ny, nx):
    """pointpillars.py:444-476."""
    nchannels = voxel_features.shape[1]
    batch_canvas = []
    for batch_itt in range(batch_size):
        canvas = torch.zeros(nchannels, nx * ny, dtype=voxel_features.dtype, device=voxel_features.device)
        batch_mask = coords[:, 0] == batch_itt
        this_coords = coords[batch_mask, :]
        indices = this_coords[:, 2] * nx + this_coords[:, 3]
        indices = indices.type(torch.long)
        voxels = voxel_features[batch_mask, :]
        voxels = voxels.t()
        canvas[:, indices] = voxels
        batch_canvas.append(canvas)
    batch_canvas = torch.stack(batch_canvas, 0)
    batch_canvas = batch_canvas.view(batch_size, nchannels, ny, nx)
    return batch_canvas
