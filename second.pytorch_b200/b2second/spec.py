"""Host-only description of a SECOND inference network, read off a module tree by duck typing.

The fused engine (``b2second.engine``) must sit behind the *reference's own* ``VoxelNet`` -- the object
``second/pytorch/builder/second_builder.py:22-133`` returns -- as well as behind the mirror classes in
``b2second.models``.  So nothing here looks at classes or at a ``ModelConfig``: only at the attribute names
the reference defines:

  VoxelNet            second/pytorch/models/voxelnet.py:104-158 (``_use_rotate_nms``, ``_multiclass_nms``,
                      ``_nms_*``, ``_post_center_range``, ``_dir_offset`` ..., ``voxel_generator``, ``target_assigner``)
  SimpleVoxel(/Radius) second/pytorch/models/voxel_encoder.py:206-255 (``num_input_features``)
  PillarFeatureNet    second/pytorch/models/pointpillars.py:153-201 (``pfn_layers[i].linear/.norm``, ``vx vy x_offset y_offset``)
  SpMiddleFHD*        second/pytorch/models/middle.py:110-210 (``middle_conv`` = SparseSequential, ``sparse_shape``)
  PointPillarsScatter pointpillars.py:420-442 (``ny nx nchannels``)
  RPNV2               second/pytorch/models/rpn.py:202-420 (``blocks deblocks conv_box conv_cls conv_dir_cls _upsample_start_idx``)

``spec_from_module`` needs no CUDA and no ``spconv`` import: sparse conv layers are recognised by their
attributes (``subm kernel_size stride padding dilation indice_key weight``), so a network built on ANY spconv
package (the CUDA drop-in, the CPU oracle) yields the same spec -- tests/test_reference_dropin.py checks that the
spec of the unmodified reference network equals the spec of the mirror for all five BASELINE configs.
"""
import hashlib

import numpy as np
import torch
from torch import nn


class UnsupportedNetwork(ValueError):
    """the module tree has a layer pattern the fused engine does not cover (callers fall back to the
    module-by-module path on the CUDA spconv drop-in)."""


def fold_bn(bn):
    """eval-mode BatchNorm -> per-channel (scale, shift), fp32."""
    w = bn.weight if bn.weight is not None else torch.ones_like(bn.running_mean)
    b = bn.bias if bn.bias is not None else torch.zeros_like(bn.running_mean)
    scale = (w / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
    shift = (b - bn.running_mean * scale).detach().float().contiguous()
    return scale, shift


def _is_sparse_conv(m):
    return all(hasattr(m, a) for a in ("subm", "kernel_size", "stride", "padding", "dilation", "indice_key", "weight",
                                       "in_channels", "out_channels"))


def _triple(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return [int(x) for x in v]
    return [int(v)] * 3


class NetSpec:
    """plain attribute bag; see spec_from_module."""

    def signature(self):
        """hashable summary (structure + sha1 of every weight) used to compare two specs."""
        def h(t):
            if t is None:
                return None
            return hashlib.sha1(np.ascontiguousarray(t.detach().cpu().float().numpy()).tobytes()).hexdigest()

        sig = {k: getattr(self, k) for k in (
            "voxel_size", "point_cloud_range", "grid_size", "max_points_per_voxel", "max_voxels", "num_point_features",
            "vfe_kind", "vfe_num_features", "is_pillars", "sparse_shape", "bev_channels", "num_class", "box_code_size",
            "num_anchors_per_loc", "use_direction_classifier", "num_direction_bins", "direction_offset",
            "direction_limit_offset", "use_rotate_nms", "multiclass_nms", "nms_class_agnostic", "nms_score_thresholds",
            "nms_pre_max_sizes", "nms_post_max_sizes", "nms_iou_thresholds", "post_center_range")}
        sig["layers"] = [{k: (h(v) if isinstance(v, torch.Tensor) else v) for k, v in lyr.items()}
                         for lyr in self.layers]
        if self.pfn is not None:
            sig["pfn"] = {k: (h(v) if isinstance(v, torch.Tensor) else v) for k, v in self.pfn.items()}
        sig["rpn"] = {k: h(v) for k, v in self.rpn.state_dict().items() if v.dtype.is_floating_point}
        return sig


def _f(v):
    """list of floats at float32 precision: protobuf configs hold float32 (0.30000001...), hand-written ones Python
    floats (0.3); every consumer is a C `float` parameter, so float32 is the value that counts."""
    return [float(np.float32(x)) for x in v]


def spec_from_module(net, max_voxels=None):
    """VoxelNet-like module tree (reference-built or mirror) -> NetSpec.  Raises UnsupportedNetwork."""
    s = NetSpec()
    s.device = next(net.parameters()).device
    # ---- voxel generator (spconv.utils.VoxelGeneratorV2 attributes, voxel_builder.py:23-32)
    vg = net.voxel_generator
    s.voxel_size = _f(vg.voxel_size)
    s.point_cloud_range = _f(vg.point_cloud_range)
    s.grid_size = [int(x) for x in vg.grid_size]                  # xyz
    T = getattr(vg, "max_num_points_per_voxel", None)
    if T is None:
        T = getattr(vg, "_max_num_points")
    s.max_points_per_voxel = int(T)
    s.max_voxels = int(max_voxels or getattr(vg, "_max_voxels", 20000))
    # ---- VFE
    vfe = net.voxel_feature_extractor
    s.pfn = None
    if hasattr(vfe, "pfn_layers"):
        s.vfe_kind = "pfn"
        if len(vfe.pfn_layers) != 1:
            raise UnsupportedNetwork("multi-layer PillarFeatureNet (every BASELINE config has num_filters=[64])")
        lyr = vfe.pfn_layers[0]
        if not isinstance(lyr.norm, nn.BatchNorm1d) or lyr.linear.bias is not None:
            raise UnsupportedNetwork("PFNLayer without BatchNorm (use_norm=False)")
        sc, sh = fold_bn(lyr.norm)
        s.pfn = {"w": lyr.linear.weight.detach().float().contiguous(), "scale": sc, "shift": sh,
                 "cout": int(lyr.linear.weight.shape[0]), "vx": float(vfe.vx), "vy": float(vfe.vy),
                 "x_offset": float(vfe.x_offset), "y_offset": float(vfe.y_offset),
                 "with_distance": bool(getattr(vfe, "_with_distance", False))}
        s.num_point_features = int(lyr.linear.weight.shape[1]) - 5 - (1 if s.pfn["with_distance"] else 0)
        s.vfe_num_features = s.num_point_features
    elif hasattr(vfe, "num_input_features") and not any(True for _ in vfe.parameters()):
        name = type(vfe).__name__
        if name not in ("SimpleVoxel", "SimpleVoxelRadius"):
            raise UnsupportedNetwork("voxel feature extractor %s" % name)
        s.vfe_kind = "mean" if name == "SimpleVoxel" else "mean_radius"
        s.vfe_num_features = int(vfe.num_input_features)
        s.num_point_features = int(getattr(net, "_num_input_features", s.vfe_num_features))
    else:
        raise UnsupportedNetwork("voxel feature extractor %s" % type(vfe).__name__)
    # ---- middle
    mid = net.middle_feature_extractor
    s.layers = []
    if hasattr(mid, "middle_conv") and hasattr(mid, "sparse_shape"):
        s.is_pillars = False
        s.sparse_shape = [int(x) for x in mid.sparse_shape]
        mods = list(mid.middle_conv._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if not _is_sparse_conv(m):
                raise UnsupportedNetwork("middle_conv[%d] is %s, expected a sparse conv" % (i, type(m).__name__))
            bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) else None
            relu = i + 1 + (bn is not None) < len(mods) and isinstance(mods[i + 1 + (bn is not None)], nn.ReLU)
            ks = _triple(m.kernel_size)
            K = int(np.prod(ks))
            if K > 27 or getattr(m, "transposed", False) or getattr(m, "inverse", False) or \
                    getattr(m, "groups", 1) != 1:
                raise UnsupportedNetwork("sparse conv with K > 27 / transposed / groups")
            lyr = {"K": K, "kernel_size": ks, "stride": _triple(m.stride), "padding": _triple(m.padding),
                   "dilation": _triple(m.dilation), "subm": bool(m.subm), "indice_key": m.indice_key,
                   "cin": int(m.in_channels), "cout": int(m.out_channels), "relu": bool(relu),
                   "w": m.weight.detach().float().contiguous().view(K, m.in_channels, m.out_channels)}
            bias = m.bias.detach().float() if getattr(m, "bias", None) is not None else None
            if bn is not None:
                lyr["scale"], lyr["shift"] = fold_bn(bn)
                if bias is not None:                       # conv bias ahead of a BN: BN(x + b) = scale*x + (shift + scale*b)
                    lyr["shift"] = (lyr["shift"] + lyr["scale"] * bias).contiguous()
            else:
                lyr["scale"], lyr["shift"] = None, (bias.contiguous() if bias is not None else None)
            s.layers.append(lyr)
            i += 1 + (bn is not None) + (1 if relu else 0)
        if not s.layers:
            raise UnsupportedNetwork("empty middle_conv")
        s.bev_channels = None                                  # = last cout * final depth, known after planning
    elif all(hasattr(mid, a) for a in ("nx", "ny", "nchannels")):
        s.is_pillars = True
        s.sparse_shape = [1, int(mid.ny), int(mid.nx)]
        s.bev_channels = int(mid.nchannels)
        if s.vfe_kind != "pfn":
            raise UnsupportedNetwork("PointPillarsScatter without a PillarFeatureNet")
    else:
        raise UnsupportedNetwork("middle feature extractor %s" % type(mid).__name__)
    if not s.is_pillars and s.vfe_kind == "pfn":
        raise UnsupportedNetwork("PillarFeatureNet in front of a sparse middle extractor")
    # ---- RPN (layer plan is made by b2second.tc.plan_rpn from the module itself)
    rpn = net.rpn
    for a in ("blocks", "deblocks", "conv_box", "conv_cls", "_upsample_start_idx", "_num_anchor_per_loc",
              "_box_code_size", "_num_class"):
        if not hasattr(rpn, a):
            raise UnsupportedNetwork("RPN %s lacks attribute %s" % (type(rpn).__name__, a))
    s.rpn = rpn
    s.num_class = int(rpn._num_class)
    s.box_code_size = int(rpn._box_code_size)
    s.num_anchors_per_loc = int(rpn._num_anchor_per_loc)
    s.use_direction_classifier = bool(getattr(rpn, "_use_direction_classifier", False))
    s.num_direction_bins = int(getattr(rpn, "_num_direction_bins", 2))
    # ---- predict() parameters (voxelnet.py:104-143)
    if not getattr(net, "_encode_background_as_zeros", True) or not getattr(net, "_use_sigmoid_score", True):
        raise UnsupportedNetwork("softmax scores / background class (every BASELINE config: sigmoid, background as zeros)")
    s.use_rotate_nms = bool(net._use_rotate_nms)
    s.multiclass_nms = bool(net._multiclass_nms)
    s.nms_class_agnostic = bool(getattr(net, "_nms_class_agnostic", False))
    s.nms_score_thresholds = _f(net._nms_score_thresholds)
    s.nms_pre_max_sizes = [int(x) for x in net._nms_pre_max_sizes]
    s.nms_post_max_sizes = [int(x) for x in net._nms_post_max_sizes]
    s.nms_iou_thresholds = _f(net._nms_iou_thresholds)
    s.post_center_range = _f(getattr(net, "_post_center_range", None) or [])
    s.direction_offset = float(np.float32(getattr(net, "_dir_offset", 0.0)))
    s.direction_limit_offset = float(np.float32(getattr(net, "_dir_limit_offset", 0.0)))
    # per-class anchor index ranges (target_assigner.anchors_range, used by the per-class NMS branch)
    s.class_anchor_counts = None
    ta = getattr(net, "target_assigner", None)
    if ta is not None and hasattr(ta, "_anchor_generators"):
        s.class_anchor_counts = [int(g.num_anchors_per_localization) for g in ta._anchor_generators]
    s.anchor_source = net
    return s


def anchors_for(spec, feature_hw):
    """[A, code] float32 anchors of the network for an H x W RPN output map.

    mirror: ``net.anchors()``; reference: ``net.target_assigner.generate_anchors([1, H, W])``
    (second/core/target_assigner.py:169-207; feature-map rule second/builder/dataset_builder.py:58,66-67)."""
    net = spec.anchor_source
    H, W = feature_hw
    if hasattr(net, "anchors") and callable(net.anchors):
        a = net.anchors()
    else:
        ret = net.target_assigner.generate_anchors([1, int(H), int(W)])
        a = ret["anchors"]
    a = np.asarray(a, dtype=np.float32).reshape(-1, spec.box_code_size)
    return a
