"""PointPillars drop-ins: decoration + scatter on the B200 kernels.

Mirrors second/second/pytorch/models/pointpillars.py and the registry of
voxel_encoder.py / middle.py: same class names, constructor
arguments, ``forward`` signatures and ``state_dict`` keys
(``pfn_layers.{i}.linear.weight``, ``pfn_layers.{i}.norm.*``) so reference
checkpoints load and ``module_class_name`` in a config resolves unchanged.

  PillarFeatureNet / ...Old / ...Radius / ...RadiusHeight
      decoration (mean offset, pillar-centre offset, padding mask, variants) is ONE
      fused CUDA kernel (lv_pillar_decorate).  In training the PFNLayer (Linear 9->64 +
      BN + ReLU + max) stays in PyTorch exactly as in the reference (batch statistics,
      autograd); in eval mode a single-layer net runs decoration + PFNLayer as ONE kernel
      (lv_pillar_pfn) and the (P,T,9) tensor is never written.
  PointPillarsScatter
      scatter-as-gather CUDA kernel (lv_pillar_scatter), differentiable w.r.t.
      ``voxel_features`` (the backward is a torch gather of the canvas gradient).
  SimpleVoxel / SimpleVoxelRadius  (voxel_encoder.py:207-255) via lv_voxel_mean.

CUDA float32 tensors, and float16 ones for the decoration and the scatter (apex O2,
second/second/pytorch/train.py:34-47: "voxels" arrive as half); anything else raises - there is no
CPU path.
"""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import _native as nat

REGISTERED_VFE_CLASSES = {}
REGISTERED_MIDDLE_CLASSES = {}


def register_vfe(cls, name=None):
    """voxel_encoder.py:13-19."""
    if name is None:
        name = cls.__name__
    assert name not in REGISTERED_VFE_CLASSES, f"exist class: {REGISTERED_VFE_CLASSES}"
    REGISTERED_VFE_CLASSES[name] = cls
    return cls


def get_vfe_class(name):
    """voxel_encoder.py:21-24."""
    assert name in REGISTERED_VFE_CLASSES, f"available class: {REGISTERED_VFE_CLASSES}"
    return REGISTERED_VFE_CLASSES[name]


def register_middle(cls, name=None):
    """middle.py:17-23."""
    if name is None:
        name = cls.__name__
    assert name not in REGISTERED_MIDDLE_CLASSES, f"exist class: {REGISTERED_MIDDLE_CLASSES}"
    REGISTERED_MIDDLE_CLASSES[name] = cls
    return cls


def get_middle_class(name):
    """middle.py:25-28."""
    assert name in REGISTERED_MIDDLE_CLASSES, f"available class: {REGISTERED_MIDDLE_CLASSES}"
    return REGISTERED_MIDDLE_CLASSES[name]


def get_paddings_indicator(actual_num, max_num, axis=0):
    """voxel_encoder.py:27-48 (kept for callers; the kernels apply the mask themselves)."""
    actual_num = torch.unsqueeze(actual_num, axis + 1)
    shape = [1] * len(actual_num.shape)
    shape[axis + 1] = -1
    steps = torch.arange(max_num, dtype=torch.int, device=actual_num.device).view(shape)
    return actual_num.int() > steps


def _require_cuda_f32(t, name, half_ok=False):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise nat.LyftVoxelError(nat.LV_E_NODEVICE, "%s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != torch.float32 and not (half_ok and t.dtype == torch.float16):
        raise nat.LyftVoxelError(nat.LV_E_INVALID, "%s must be float32%s, got %s" %
                                 (name, " or float16" if half_ok else "", t.dtype))


def _f32(x):
    """Round a Python scalar the way torch does when it meets a float32 tensor."""
    return float(np.float32(x))


def decorate_pillars(features, num_voxels, coors, vx, vy, x_offset, y_offset, variant="pfn",
                     with_distance=False):
    """Fused decoration: (P,T,C) float32 -> (P,T,C_out) float32, padding rows zero.
    pointpillars.py:203-231 (+ :117-145, :290-319, :378-411).  float16 features (apex O2) give a float16
    result with torch's half roundings (lv_pillar_decorate_half)."""
    _require_cuda_f32(features, "features", half_ok=True)
    lib = nat.load()
    features = features.contiguous()
    P, T, C = features.shape
    num = num_voxels.to(torch.int32).contiguous()
    co = coors.to(torch.int32).contiguous()
    v = nat.PILLAR_VARIANTS[variant]
    c_out = lib.lv_pillar_out_channels(C, v, int(bool(with_distance)))
    if c_out <= 0:
        raise nat.LyftVoxelError(nat.LV_E_INVALID, "bad num_features %d" % C)
    out = torch.empty((P, T, c_out), dtype=features.dtype, device=features.device)
    h = nat.get_handle(features.device.index)
    if features.dtype == torch.float16:
        with torch.cuda.device(features.device):
            nat.check(lib.lv_pillar_decorate_half(h.ptr, features.data_ptr(), num.data_ptr(), co.data_ptr(), P, T, C,
                                                  _f32(vx), _f32(vy), _f32(x_offset), _f32(y_offset), v,
                                                  int(bool(with_distance)), out.data_ptr(),
                                                  nat.current_stream_ptr(features.device)))
        return out
    with torch.cuda.device(features.device):
        nat.check(lib.lv_pillar_decorate(h.ptr, features.data_ptr(), num.data_ptr(), co.data_ptr(), P, T, C,
                                         _f32(vx), _f32(vy), _f32(x_offset), _f32(y_offset), v,
                                         int(bool(with_distance)), out.data_ptr(),
                                         nat.current_stream_ptr(features.device)))
    return out


def fold_pfn_layer(layer):
    """(weight, scale, shift) float32 CUDA tensors of a last PFNLayer in eval mode:
    relu(BN(Wx)) == relu((Wx) * scale + shift) (pointpillars.py:51-62)."""
    w = layer.linear.weight.detach().float().contiguous()
    if isinstance(layer.norm, nn.BatchNorm1d):
        bn = layer.norm
        inv = torch.rsqrt(bn.running_var.detach().double() + bn.eps)
        g = bn.weight.detach().double() if bn.affine else torch.ones_like(inv)
        b = bn.bias.detach().double() if bn.affine else torch.zeros_like(inv)
        scale = g * inv
        shift = b - bn.running_mean.detach().double() * scale
        if layer.linear.bias is not None:
            shift = shift + layer.linear.bias.detach().double() * scale
    else:
        scale = torch.ones(w.shape[0], dtype=torch.float64, device=w.device)
        shift = (layer.linear.bias.detach().double() if layer.linear.bias is not None
                 else torch.zeros(w.shape[0], dtype=torch.float64, device=w.device))
    return w, scale.float().contiguous(), shift.float().contiguous()


def pillar_pfn(features, num_voxels, coors, vx, vy, x_offset, y_offset, weight, scale, shift, variant="pfn",
               with_distance=False):
    """Decoration + last PFNLayer (eval) fused: (P,T,4) float32 -> (P,units) float32."""
    _require_cuda_f32(features, "features")
    lib = nat.load()
    features = features.contiguous()
    P, T, C = features.shape
    num = num_voxels.to(torch.int32).contiguous()
    co = coors.to(torch.int32).contiguous()
    units = weight.shape[0]
    out = torch.empty((P, units), dtype=torch.float32, device=features.device)
    h = nat.get_handle(features.device.index)
    with torch.cuda.device(features.device):
        nat.check(lib.lv_pillar_pfn(h.ptr, features.data_ptr(), num.data_ptr(), co.data_ptr(), P, T, C,
                                    _f32(vx), _f32(vy), _f32(x_offset), _f32(y_offset),
                                    nat.PILLAR_VARIANTS[variant], int(bool(with_distance)), weight.data_ptr(),
                                    scale.data_ptr(), shift.data_ptr(), units, out.data_ptr(),
                                    nat.current_stream_ptr(features.device)))
    return out


class _FusedPfnTrain(torch.autograd.Function):
    """relu(BatchNorm1d_train(Linear(decorate(voxels)))) max-pooled over the points of a pillar
    (pointpillars.py:51-65 behind :203-231) as ONE differentiable op on the CUDA library
    (lv_pillar_pfn_train_forward / lv_pillar_pfn_train_backward): neither the (P,T,C_in) decorated tensor nor the
    (P,T,units) activations exist in HBM, in the forward or in the backward, and the batch statistics and the
    BatchNorm backward are computed on the device from float64 moments.
    Returns (out (P, units), batch mean (units), biased batch variance (units))."""

    @staticmethod
    def forward(ctx, features, num, coors, weight, gamma, beta, eps, geom):
        vx, vy, xo, yo, variant, with_distance = geom
        lib = nat.load()
        units, c_in = weight.shape
        P, T, C = features.shape
        dev = features.device
        w_f = weight.detach().float().contiguous()
        g_f, b_f = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        out = torch.empty((P, units), dtype=torch.float32, device=dev)
        stats = torch.empty((5, units), dtype=torch.float32, device=dev)
        mom = torch.empty((c_in + c_in * (c_in + 1) // 2,), dtype=torch.float64, device=dev)
        h = nat.get_handle(dev.index)
        with torch.cuda.device(dev):
            nat.check(lib.lv_pillar_pfn_train_forward(
                h.ptr, features.data_ptr(), num.data_ptr(), coors.data_ptr(), P, T, C, _f32(vx), _f32(vy), _f32(xo), _f32(yo),
                nat.PILLAR_VARIANTS[variant], int(bool(with_distance)), w_f.data_ptr(), g_f.data_ptr(), b_f.data_ptr(),
                float(eps), units, out.data_ptr(), stats.data_ptr(), mom.data_ptr(), nat.current_stream_ptr(dev)))
        ctx.geom, ctx.eps = geom, float(eps)
        ctx.save_for_backward(features, num, coors, w_f, g_f, stats, mom)
        mean, var = stats[2], stats[4]
        ctx.mark_non_differentiable(mean, var)
        return out, mean, var

    @staticmethod
    def backward(ctx, grad_out, _gm, _gv):
        features, num, coors, w_f, g_f, stats, mom = ctx.saved_tensors
        vx, vy, xo, yo, variant, with_distance = ctx.geom
        lib = nat.load()
        units, c_in = w_f.shape
        P, T, C = features.shape
        dev = features.device
        g = grad_out.float().contiguous()
        d_w = torch.empty((units, c_in), dtype=torch.float32, device=dev)
        d_gamma = torch.empty((units,), dtype=torch.float32, device=dev)
        d_beta = torch.empty((units,), dtype=torch.float32, device=dev)
        h = nat.get_handle(dev.index)
        with torch.cuda.device(dev):
            nat.check(lib.lv_pillar_pfn_train_backward(
                h.ptr, features.data_ptr(), num.data_ptr(), coors.data_ptr(), P, T, C, _f32(vx), _f32(vy), _f32(xo), _f32(yo),
                nat.PILLAR_VARIANTS[variant], int(bool(with_distance)), w_f.data_ptr(), g_f.data_ptr(), ctx.eps,
                stats.data_ptr(), mom.data_ptr(), units, g.data_ptr(), d_w.data_ptr(), d_gamma.data_ptr(), d_beta.data_ptr(),
                nat.current_stream_ptr(dev)))
        return None, None, None, d_w, d_gamma, d_beta, None, None


class PFNLayer(nn.Module):
    """pointpillars.py:17-65 - unchanged PyTorch layer (Linear, BatchNorm1d eps=1e-3
    momentum=0.01, ReLU, max over the points of a pillar)."""

    def __init__(self, in_channels, out_channels, use_norm=True, last_layer=False):
        super().__init__()
        self.name = 'PFNLayer'
        self.last_vfe = last_layer
        if not self.last_vfe:
            out_channels = out_channels // 2
        self.units = out_channels
        self.linear = nn.Linear(in_channels, self.units, bias=not use_norm)
        self.norm = nn.BatchNorm1d(self.units, eps=1e-3, momentum=0.01) if use_norm else nn.Identity()

    def forward(self, inputs):
        x = self.linear(inputs)
        x = self.norm(x.permute(0, 2, 1).contiguous()).permute(0, 2, 1).contiguous()
        x = F.relu(x)
        x_max = torch.max(x, dim=1, keepdim=True)[0]
        if self.last_vfe:
            return x_max
        return torch.cat([x, x_max.repeat(1, inputs.shape[1], 1)], dim=2)


class _PillarFeatureNetBase(nn.Module):
    _variant = "pfn"
    _extra = 5          # channels added by the decoration (before with_distance)
    _name = "PillarFeatureNet"

    def __init__(self, num_input_features=4, use_norm=True, num_filters=(64,), with_distance=False,
                 voxel_size=(0.2, 0.2, 4), pc_range=(0, -40, -3, 70.4, 40, 1)):
        super().__init__()
        self.name = self._name
        assert len(num_filters) > 0
        num_input_features += self._extra
        if with_distance:
            num_input_features += 1
        self._with_distance = with_distance
        num_filters = [num_input_features] + list(num_filters)
        layers = []
        for i in range(len(num_filters) - 1):
            layers.append(PFNLayer(num_filters[i], num_filters[i + 1], use_norm,
                                   last_layer=(i >= len(num_filters) - 2)))
        self.pfn_layers = nn.ModuleList(layers)
        # pointpillars.py:198-201
        self.vx = voxel_size[0]
        self.vy = voxel_size[1]
        self.x_offset = self.vx / 2 + pc_range[0]
        self.y_offset = self.vy / 2 + pc_range[1]

    def decorate(self, features, num_voxels, coors):
        return decorate_pillars(features, num_voxels, coors, self.vx, self.vy, self.x_offset, self.y_offset,
                                variant=self._variant, with_distance=self._with_distance)

    def can_fuse(self, features):
        """The fused inference kernel: eval mode, one PFNLayer, 4 input features,
        <= 64 points per pillar, 32/64/128 units, and nothing that needs a gradient."""
        if self.training or len(self.pfn_layers) != 1 or features.dtype != torch.float32:
            return False
        layer = self.pfn_layers[0]
        if features.dim() != 3 or features.shape[2] != 4 or features.shape[1] > 64 or layer.units not in (32, 64, 128):
            return False
        needs_grad = features.requires_grad or any(q.requires_grad for q in layer.parameters())
        return not (torch.is_grad_enabled() and needs_grad)

    fuse_training = True    # class-wide switch of the fused training path (set False to run the PyTorch PFNLayer)

    def can_fuse_training(self, features):
        """The fused training path: train mode, one PFNLayer with BatchNorm1d (no Linear bias), 4 input
        features, <= 64 points per pillar, 32/64/128 units, float32 voxels that need no gradient."""
        if not (self.training and self.fuse_training) or len(self.pfn_layers) != 1 or features.dtype != torch.float32:
            return False
        layer = self.pfn_layers[0]
        if not isinstance(layer.norm, nn.BatchNorm1d) or not layer.norm.affine or layer.linear.bias is not None:
            return False
        if features.dim() != 3 or features.shape[2] != 4 or features.shape[1] > 64 or layer.units not in (32, 64, 128):
            return False
        if layer.linear.weight.dtype != torch.float32 or features.shape[0] * features.shape[1] < 2:
            return False
        return features.is_cuda and not features.requires_grad

    def _forward_training_fused(self, features, num_voxels, coors):
        layer = self.pfn_layers[0]
        bn = layer.norm
        feats = features.contiguous()
        num = num_voxels.to(torch.int32).contiguous()
        co = coors.to(torch.int32).contiguous()
        geom = (self.vx, self.vy, self.x_offset, self.y_offset, self._variant, self._with_distance)
        out, mean, var = _FusedPfnTrain.apply(feats, num, co, layer.linear.weight, bn.weight, bn.bias, bn.eps, geom)
        if bn.track_running_stats:     # what nn.BatchNorm1d does in train mode (momentum 0.01, unbiased running_var)
            with torch.no_grad():
                n = float(feats.shape[0] * feats.shape[1])
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1.0 - m).add_(mean.to(bn.running_mean.dtype), alpha=m)
                bn.running_var.mul_(1.0 - m).add_((var * (n / (n - 1.0))).to(bn.running_var.dtype), alpha=m)
        return out.squeeze()

    def forward(self, features, num_voxels, coors):
        if self.can_fuse_training(features):
            return self._forward_training_fused(features, num_voxels, coors)
        if self.can_fuse(features):
            w, scale, shift = fold_pfn_layer(self.pfn_layers[0])
            out = pillar_pfn(features, num_voxels, coors, self.vx, self.vy, self.x_offset, self.y_offset, w, scale,
                             shift, variant=self._variant, with_distance=self._with_distance)
            return out.squeeze()
        features = self.decorate(features, num_voxels, coors)
        for pfn in self.pfn_layers:
            features = pfn(features)
        return features.squeeze()


@register_vfe
class PillarFeatureNetOld(_PillarFeatureNetBase):
    """pointpillars.py:68-151 (keeps the x,y aliasing of :127-134)."""
    _variant = "old"
    _name = "PillarFeatureNetOld"


@register_vfe
class PillarFeatureNet(_PillarFeatureNetBase):
    """pointpillars.py:154-237.  (The reference defines this class without
    @register_vfe and sets name 'PillarFeatureNetOld'; the Lyft configs select it
    by module_class_name, so it is registered here.)"""
    _variant = "pfn"
    _name = "PillarFeatureNet"


@register_vfe
class PillarFeatureNetRadius(_PillarFeatureNetBase):
    """pointpillars.py:240-325."""
    _variant = "radius"
    _extra = 4
    _name = "PillarFeatureNetRadius"


@register_vfe
class PillarFeatureNetRadiusHeight(_PillarFeatureNetBase):
    """pointpillars.py:328-417."""
    _variant = "radius_height"
    _extra = 5
    _name = "PillarFeatureNetRadiusHeight"


class _ScatterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, voxel_features, coords, batch_size, ny, nx):
        lib = nat.load()
        feats = voxel_features.contiguous()
        co = coords.to(torch.int32).contiguous()
        P, C = feats.shape
        canvas = torch.empty((batch_size, C, ny, nx), dtype=feats.dtype, device=feats.device)
        h = nat.get_handle(feats.device.index)
        fn = lib.lv_pillar_scatter_half if feats.dtype == torch.float16 else lib.lv_pillar_scatter
        with torch.cuda.device(feats.device):
            nat.check(fn(h.ptr, feats.data_ptr(), co.data_ptr(), P, C, batch_size, ny, nx,
                         canvas.data_ptr(), nat.current_stream_ptr(feats.device)))
        ctx.save_for_backward(co)
        ctx.dims = (batch_size, C, ny, nx)
        return canvas

    @staticmethod
    def backward(ctx, grad):
        """d feats[p] = d canvas[b, :, y, x] (lv_pillar_scatter_backward)."""
        (co,) = ctx.saved_tensors
        B, C, ny, nx = ctx.dims
        lib = nat.load()
        g = grad.contiguous()
        if g.dtype not in (torch.float32, torch.float16):
            g = g.float()
        out = torch.empty((co.shape[0], C), dtype=g.dtype, device=g.device)
        h = nat.get_handle(g.device.index)
        with torch.cuda.device(g.device):
            nat.check(lib.lv_pillar_scatter_backward(h.ptr, g.data_ptr(), co.data_ptr(), co.shape[0], C, B, ny, nx,
                                                     int(g.dtype == torch.float16), out.data_ptr(),
                                                     nat.current_stream_ptr(g.device)))
        return out.to(grad.dtype), None, None, None, None


def scatter_pillars(voxel_features, coords, batch_size, ny, nx):
    """canvas[b,:,y,x] = voxel_features[p] for coords[p] = (b,.,y,x); zeros elsewhere.  The canvas has the
    dtype of voxel_features (pointpillars.py:449-452): float32, or float16 under apex O2 - written in one pass."""
    _require_cuda_f32(voxel_features, "voxel_features", half_ok=True)
    return _ScatterFn.apply(voxel_features, coords, int(batch_size), int(ny), int(nx))


@register_middle
class PointPillarsScatter(nn.Module):
    """pointpillars.py:420-476."""

    def __init__(self, output_shape, use_norm=True, num_input_features=64, num_filters_down1=[64],
                 num_filters_down2=[64, 64], name='SpMiddle2K'):
        super().__init__()
        self.name = 'PointPillarsScatter'
        self.output_shape = output_shape
        self.ny = output_shape[2]
        self.nx = output_shape[3]
        self.nchannels = num_input_features

    def forward(self, voxel_features, coords, batch_size):
        return scatter_pillars(voxel_features, coords, batch_size, self.ny, self.nx)


def voxel_mean(features, num_voxels, num_input_features):
    _require_cuda_f32(features, "features")
    lib = nat.load()
    features = features.contiguous()
    P, T, C = features.shape
    num = num_voxels.to(torch.int32).contiguous()
    out = torch.empty((P, num_input_features), dtype=torch.float32, device=features.device)
    h = nat.get_handle(features.device.index)
    with torch.cuda.device(features.device):
        nat.check(lib.lv_voxel_mean(h.ptr, features.data_ptr(), num.data_ptr(), P, T, C, num_input_features,
                                    out.data_ptr(), nat.current_stream_ptr(features.device)))
    return out


@register_vfe
class SimpleVoxel(nn.Module):
    """voxel_encoder.py:207-225: per-voxel mean of the first num_input_features channels."""

    def __init__(self, num_input_features=4, use_norm=True, num_filters=[32, 128], with_distance=False,
                 voxel_size=(0.2, 0.2, 4), pc_range=(0, -40, -3, 70.4, 40, 1), name='VoxelFeatureExtractor'):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features

    def forward(self, features, num_voxels, coors):
        return voxel_mean(features, num_voxels, self.num_input_features)


@register_vfe
class SimpleVoxelRadius(nn.Module):
    """voxel_encoder.py:228-255: mean, then (|xy|, z, features...)."""

    def __init__(self, num_input_features=4, use_norm=True, num_filters=(32, 128), with_distance=False,
                 voxel_size=(0.2, 0.2, 4), pc_range=(0, -40, -3, 70.4, 40, 1), name='SimpleVoxelRadius'):
        super().__init__()
        self.num_input_features = num_input_features
        self.name = name

    def forward(self, features, num_voxels, coors):
        mean = voxel_mean(features, num_voxels, self.num_input_features)
        radius = torch.norm(mean[:, :2], p=2, dim=1, keepdim=True)
        return torch.cat([radius, mean[:, 2:self.num_input_features]], dim=1)
