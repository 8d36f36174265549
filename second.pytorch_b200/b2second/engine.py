"""Fused device pipeline for ``VoxelNet`` inference: points in HBM -> detections in HBM, no host sync.

This is the fast path behind the ``VoxelNet.forward(example)`` contract ("a fast path may
additionally accept example['points']").  Where the module-by-module path (``models.VoxelNet`` on the
``spconv`` drop-in) mirrors the reference call by call -- one rulebook sync per strided conv, a
D2H -> CPU NMS -> H2D round trip per frame (second/pytorch/core/box_torch_ops.py:503,512) -- the engine
keeps every data-dependent count in device memory and runs a fixed launch sequence over
capacity-sized buffers, so the whole frame batch is ONE CUDA graph:

  b2s_voxelize (+ fused SimpleVoxel mean)                     voxelnet.py:325-328, preprocess.py:303-315
  per sparse layer: b2s_rulebook_{subm,conv} (cached per indice_key) + b2s_sparse_conv_tc (tcgen05, hi/lo
      planes flow from layer to layer) with BatchNorm1d/ReLU folded into the epilogue      middle.py:145-192
  b2s_pfn (PointPillars)                                       pointpillars.py:203-237
  b2s_to_bev_tc (NHWC + halo, hi/lo)                           middle.py:206-209 / pointpillars.py:444-476
  RPN as a program of b2s_conv2d_tc_ex launches (tc.plan_rpn)  rpn.py:314-331,393-420,467-497
      (rpn_impl="cudnn" keeps the torch modules as a cross-check: fp32, TF32 off)
  b2s_decode_filter_strided + b2s_nms (+ direction/range epilogue)   voxelnet.py:377-645

Output per batch: ``det [B, post_max, code+2]`` (box, score, label) + ``det_count [B]`` -- the fixed-stride
record the multi-GPU path all-gathers.
"""
import ctypes
import os

import numpy as np
import torch
from torch import nn

from . import models


def _pow2_at_least(n):
    c = 1024
    while c < n:
        c <<= 1
    return c


def _fold_bn(bn):
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
    shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
    return scale, shift


def ctypes_ptr(addr):
    return ctypes.c_void_p(int(addr))


class _Level:
    """one active-site set: coordinates, device count, coordinate->row hash, row capacity."""
    __slots__ = ("coors", "n_dev", "keys", "vals", "hcap", "cap", "shape")


class InferenceEngine:
    def __init__(self, net, batch_size=1, max_points=None, max_voxels=None, row_cap_factor=2.0,
                 cand_cap=None, use_cuda_graph=True, rpn_impl="auto", sparse_impl="tc"):
        import spconv as sp                      # the CUDA drop-in: fails loudly if the library is missing
        assert not getattr(sp, "__oracle__", False), "the engine is the product path; it never runs on the oracle"
        self.sp = sp
        self.lib = sp._lib.load()
        self._L = sp._lib
        self.net = net.eval()
        self.cfg = cfg = net.cfg
        self.B = int(batch_size)
        dev = next(net.parameters()).device
        assert dev.type == "cuda", "InferenceEngine needs the network on a CUDA device"
        self.dev = dev
        # parity bar is fp32: keep cuDNN/cuBLAS off TF32 (torch allows TF32 convs by default)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True
        self.F = cfg.num_point_features
        self.T = cfg.max_points_per_voxel
        self.max_voxels = int(max_voxels or cfg.max_voxels)
        self.P_cap = int(max_points or 40000) * self.B
        self.grid = cfg.grid_size
        self.code = cfg.box_code_size
        self.use_graph = use_cuda_graph
        self._graph = None
        # RPN: "tc" = hand-written tcgen05 implicit GEMM (csrc/conv_tc.cu, 3xTF32), "cudnn" = torch/cuDNN fp32.
        from . import tc as _tc
        if rpn_impl == "auto":
            rpn_impl = "tc" if _tc.supported(net.rpn) else "cudnn"
        assert rpn_impl in ("tc", "cudnn")
        if rpn_impl == "tc" and not _tc.supported(net.rpn):
            raise ValueError("rpn_impl='tc': this RPN has a layer b2s_conv2d_tc_ex does not cover (channels % 32, kernel > 4)")
        self.rpn_impl = rpn_impl
        assert sparse_impl in ("tc", "fma")       # sparse-conv inner product: tcgen05 3xTF32 | fp32 FMA tiles
        self.sparse_impl = sparse_impl
        self._plan_middle(row_cap_factor)
        self._alloc_voxel_buffers()
        if rpn_impl == "tc":
            self._alloc_tc_rpn(_tc)
        self._alloc_detect_buffers(cand_cap)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)

    # ---------------------------------------------------------------- planning / allocation
    def _new_level(self, cap, shape, with_storage=True):
        lv = _Level()
        lv.cap = int(cap)
        lv.shape = [int(s) for s in shape]
        lv.hcap = _pow2_at_least(2 * lv.cap)
        if with_storage:
            lv.coors = torch.zeros(lv.cap, 4, dtype=torch.int32, device=self.dev)
            lv.n_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
            lv.keys = torch.empty(lv.hcap, dtype=torch.int64, device=self.dev)
            lv.vals = torch.empty(lv.hcap, dtype=torch.int32, device=self.dev)
        return lv

    def _plan_middle(self, row_cap_factor):
        cfg, sp = self.cfg, self.sp
        mid = self.net.middle_feature_extractor
        self.is_pillars = isinstance(mid, models.PointPillarsScatter)
        self.layers = []
        self.rb_ws = None
        cap0 = self.B * self.max_voxels
        if self.is_pillars:
            self.level0 = self._new_level(cap0, [1, int(self.grid[1]), int(self.grid[0])], with_storage=False)
            pfn = self.net.voxel_feature_extractor
            assert len(pfn.pfn_layers) == 1, "engine: single-layer PillarFeatureNet only (all BASELINE configs)"
            lyr = pfn.pfn_layers[0]
            self.pfn_w = lyr.linear.weight.detach().float().contiguous()
            self.pfn_scale, self.pfn_shift = _fold_bn(lyr.norm)
            self.pfn_cout = lyr.units
            self.pfn_geom = (float(pfn.vx), float(pfn.vy), float(pfn.x_offset), float(pfn.y_offset))
            self.feat_final_c = self.pfn_cout
            self.final_level = self.level0
            self.pfn_out = torch.zeros(cap0, self.pfn_cout, dtype=torch.float32, device=self.dev)
            return
        shape0 = [int(s) for s in mid.sparse_shape]
        self.level0 = self._new_level(cap0, shape0, with_storage=False)   # storage = voxelizer outputs
        mods = list(mid.middle_conv._modules.values())
        level = self.level0
        rulebooks = {}
        i = 0
        max_ws = 0
        while i < len(mods):
            m = mods[i]
            assert isinstance(m, sp.SparseConvolution), "engine: expected (conv, BN, ReLU) triples"
            bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) else None
            relu = bn is not None and i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
            K = int(np.prod(m.kernel_size))
            lyr = {"index": len(self.layers), "conv": m, "K": K, "cin": m.in_channels, "cout": m.out_channels, "relu": relu,
                   "w": m.weight.detach().float().contiguous().view(K, m.in_channels, m.out_channels)}
            if bn is not None:
                lyr["scale"], lyr["shift"] = _fold_bn(bn)
            else:
                lyr["scale"], lyr["shift"] = None, (m.bias.detach().float().contiguous() if m.bias is not None else None)
            if m.subm:
                key = ("subm", m.indice_key, id(level)) if m.indice_key is not None else ("subm", id(m))
                if key not in rulebooks:
                    rulebooks[key] = {"nbr": torch.empty(level.cap, K, dtype=torch.int32, device=self.dev),
                                      "build": ("subm", level, m)}
                    lyr["build_rb"] = True
                else:
                    lyr["build_rb"] = False
                lyr["rb"] = rulebooks[key]
                lyr["in_level"], lyr["out_level"] = level, level
            else:
                out_shape = sp.ops.get_conv_output_size(level.shape, m.kernel_size, m.stride, m.padding, m.dilation)
                cells = int(np.prod(out_shape))
                fan = int(np.prod([-(-k // s) for k, s in zip(m.kernel_size, m.stride)]))
                per_frame = min(cells, int(self.max_voxels * row_cap_factor), level.cap // self.B * fan)
                new = self._new_level(self.B * per_frame, out_shape)
                rb = {"nbr": torch.empty(new.cap, K, dtype=torch.int32, device=self.dev), "build": ("conv", level, m)}
                lyr["rb"], lyr["build_rb"] = rb, True
                lyr["in_level"], lyr["out_level"] = level, new
                max_ws = max(max_ws, self.lib.b2s_rulebook_conv_workspace_bytes(self.B, self._L.i3(out_shape)))
                level = new
            lyr["out"] = torch.zeros(lyr["out_level"].cap, m.out_channels, dtype=torch.float32, device=self.dev)
            # tensor-pipe core (csrc/sparse_conv_tc.cu); thin layers (Cin 4/16) pack 8/2 kernel offsets per K block.
            # Other widths (e.g. 3 input features) stay on the fp32 FMA core.
            thin_ok = os.environ.get("B2S_THIN_TC", "1") != "0"      # A/B switch: thin layers on the FMA core
            lyr["tc"] = (self.sparse_impl == "tc" and m.in_channels in ((4, 16, 32, 64) if thin_ok else (32, 64))
                         and m.out_channels in (16, 32, 64) and K <= 27)
            self.layers.append(lyr)
            i += 1 + (1 if bn is not None else 0) + (1 if relu else 0)
        # once a layer runs on the tensor pipe all later ones must too (hi/lo planes flow forward)
        seen_tc = False
        for lyr in self.layers:
            if seen_tc and not lyr["tc"]:
                for l2 in self.layers:
                    l2["tc"] = False
                break
            seen_tc = seen_tc or lyr["tc"]
        from . import tc as _tc
        for j, lyr in enumerate(self.layers):
            if lyr["tc"]:
                lyr["w_hi"], lyr["w_lo"] = _tc.split_tf32(_tc.pack_sparse_weights(lyr["w"]))
                lyr["out_lo"] = torch.zeros_like(lyr["out"])
                if j == 0 or not self.layers[j - 1]["tc"]:
                    # first tensor-pipe layer: its fp32 input rows are split into hi/lo planes first
                    lyr["in_split"] = (torch.zeros(lyr["in_level"].cap, lyr["cin"], dtype=torch.float32, device=self.dev),
                                       torch.zeros(lyr["in_level"].cap, lyr["cin"], dtype=torch.float32, device=self.dev))
        self.any_sparse_tc = any(l["tc"] for l in self.layers)
        if self.any_sparse_tc:
            last = self.layers[-1]
            self.merged_out = torch.zeros_like(last["out"])
        self.rb_ws = torch.empty(max(max_ws, 1), dtype=torch.uint8, device=self.dev)
        self.rb_ws_bytes = max_ws
        self.final_level = level
        self.feat_final_c = self.layers[-1]["cout"]

    def _alloc_voxel_buffers(self):
        lib, dev = self.lib, self.dev
        cap = self.B * self.max_voxels
        self.points = torch.zeros(self.P_cap, self.F, dtype=torch.float32, device=dev)
        self.offsets = torch.zeros(self.B + 1, dtype=torch.int32, device=dev)
        self.vox_coors = torch.zeros(cap, 4, dtype=torch.int32, device=dev)
        self.vox_num = torch.zeros(cap, dtype=torch.int32, device=dev)
        self.vox_slots = torch.zeros(cap, self.T, dtype=torch.int32, device=dev)
        self.num_voxels = torch.zeros(1 + self.B, dtype=torch.int32, device=dev)
        self.vox_hcap = lib.b2s_voxelize_hash_capacity(self.P_cap)
        self.vox_keys = torch.empty(self.vox_hcap, dtype=torch.int64, device=dev)
        self.vox_vals = torch.empty(self.vox_hcap, dtype=torch.int32, device=dev)
        self.vox_ws_bytes = lib.b2s_voxelize_workspace_bytes(self.P_cap, self.B, self.max_voxels, self.T)
        self.vox_ws = torch.empty(max(self.vox_ws_bytes, 1), dtype=torch.uint8, device=dev)
        vfe = self.net.voxel_feature_extractor
        if isinstance(vfe, models.SimpleVoxel):
            self.vfe_mode, self.vfe_nf, c = 1, vfe.num_input_features, vfe.num_input_features
        elif isinstance(vfe, models.SimpleVoxelRadius):
            self.vfe_mode, self.vfe_nf, c = 2, vfe.num_input_features, vfe.num_input_features - 1
        else:
            self.vfe_mode, self.vfe_nf, c = 0, self.F, 0
        self.vfe_out = torch.zeros(cap, c, dtype=torch.float32, device=dev) if c else None
        # level 0 aliases the voxelizer outputs
        lv = self.level0
        lv.coors, lv.n_dev, lv.keys, lv.vals, lv.hcap = (self.vox_coors, self.num_voxels, self.vox_keys,
                                                         self.vox_vals, self.vox_hcap)
        C = self.feat_final_c
        D, H, W = self.final_level.shape
        self.bev = torch.zeros(self.B, C * D, H, W, dtype=torch.float32, device=dev)

    def _alloc_tc_rpn(self, _tc):
        """buffers of the tensor-core RPN: NHWC halo-padded hi/lo planes (halo zero-filled once, never written)."""
        dev, B = self.dev, self.B
        D, H, W = self.final_level.shape
        C = self.feat_final_c * D
        plan = _tc.plan_rpn(self.net.rpn, H, W)
        self.tc_prog = plan
        self.tc_plan = plan["ops"]
        assert plan["in_channels"] == C, "BEV channels %d != RPN input %d" % (C, plan["in_channels"])

        def plane(h, w, c):
            return (torch.zeros(B, h + 2, w + 2, c, dtype=torch.float32, device=dev),
                    torch.zeros(B, h + 2, w + 2, c, dtype=torch.float32, device=dev))
        self.tc_bev = plane(H, W, C)
        self.tc_bufs = {"in": self.tc_bev}
        for name, (h, w, c) in plan["buffers"].items():
            self.tc_bufs[name] = plane(h, w, c)
        hd = plan["heads"]
        self.tc_head_stride = hd["stride"]
        _, fH, fW = self.cfg.feature_map_size
        assert (hd["H"], hd["W"]) == (fH, fW), "RPN output %dx%d != anchor grid %dx%d" % (hd["H"], hd["W"], fH, fW)
        self.tc_heads = torch.zeros(B, hd["H"], hd["W"], self.tc_head_stride, dtype=torch.float32, device=dev)
        self.tc_bufs["heads"] = (self.tc_heads, None)
        self.tc_hw = (H, W)

    def _alloc_detect_buffers(self, cand_cap):
        cfg, dev = self.cfg, self.dev
        anchors = torch.from_numpy(self.net.anchors()).to(dev)
        self.anchors = anchors.contiguous()
        self.A = anchors.shape[0]
        self.a_loc = cfg.num_anchors_per_loc
        _, self.fH, self.fW = cfg.feature_map_size
        assert self.a_loc * self.fH * self.fW == self.A
        self.cand_cap = int(cand_cap or min(self.A, 32768))
        B, cc, code = self.B, self.cand_cap, self.code
        self.cand_box = torch.zeros(B, cc, code, dtype=torch.float32, device=dev)
        self.cand_score = torch.zeros(B, cc, dtype=torch.float32, device=dev)
        self.cand_label = torch.zeros(B, cc, dtype=torch.int32, device=dev)
        self.cand_dir = torch.zeros(B, cc, dtype=torch.int32, device=dev)
        self.cand_anchor = torch.zeros(B, cc, dtype=torch.int32, device=dev)
        self.cand_count = torch.zeros(B, dtype=torch.int32, device=dev)
        self.pre_max, self.post_max = cfg.nms_pre_max_size, cfg.nms_post_max_size
        self.nms_ws_bytes = self.lib.b2s_nms_workspace_bytes(B, cc, self.pre_max)
        self.nms_ws = torch.empty(max(self.nms_ws_bytes, 1), dtype=torch.uint8, device=dev)
        self.det = torch.zeros(B, self.post_max, code + 2, dtype=torch.float32, device=dev)
        self.det_count = torch.zeros(B, dtype=torch.int32, device=dev)
        rng = cfg.post_center_limit_range
        self.range_host = self._L.f6(rng) if len(rng) == 6 else None

    # ---------------------------------------------------------------- the launch sequence
    def _launch(self):
        L, lib, cfg = self._L, self.lib, self.cfg
        st = L.stream()
        self.status.zero_()
        self._mark("voxelize")
        L.check(lib.b2s_voxelize(
            L.ptr(self.points), L.ptr(self.offsets), self.P_cap_used, self.F, self.B,
            L.f3(cfg.point_cloud_range[:3]), L.f3(cfg.voxel_size), L.i3(self.grid), self.T, self.max_voxels,
            L.ptr(self.vox_coors), L.ptr(self.vox_num), L.ptr(self.vox_slots), None, self.vfe_mode, self.vfe_nf,
            L.ptr(self.vfe_out), L.ptr(self.num_voxels), L.ptr(self.vox_keys), L.ptr(self.vox_vals), self.vox_hcap,
            int(self.level0.shape[0]),       # hash keys in the middle encoder's shape (grid_z + 1, middle.py:139)
            L.ptr(self.vox_ws), self.vox_ws_bytes, L.ptr(self.status), st), "b2s_voxelize")
        if self.is_pillars:
            vx, vy, xo, yo = self.pfn_geom
            self._mark("pfn")
            L.check(lib.b2s_pfn(L.ptr(self.points), self.F, L.ptr(self.vox_slots), L.ptr(self.vox_num),
                                L.ptr(self.vox_coors), L.ptr(self.num_voxels), self.level0.cap, self.T,
                                L.ptr(self.pfn_w), L.ptr(self.pfn_scale), L.ptr(self.pfn_shift), self.pfn_cout,
                                vx, vy, xo, yo, L.ptr(self.pfn_out), st), "b2s_pfn")
            feats = self.pfn_out
        else:
            feats, feats_lo = self.vfe_out, None
            for lyr in self.layers:
                m, lin, lout = lyr["conv"], lyr["in_level"], lyr["out_level"]
                if lyr["build_rb"]:
                    self._mark("rulebook%d" % lyr["index"])
                    if m.subm:
                        L.check(lib.b2s_rulebook_subm(L.ptr(lin.coors), L.ptr(lin.n_dev), lin.cap, L.i3(lin.shape),
                                                      L.i3(m.kernel_size), L.i3(m.dilation), L.ptr(lin.keys),
                                                      L.ptr(lin.vals), lin.hcap, L.ptr(lyr["rb"]["nbr"]), st),
                                "b2s_rulebook_subm")
                    else:
                        L.check(lib.b2s_rulebook_conv(
                            L.ptr(lin.coors), L.ptr(lin.n_dev), lin.cap, self.B, L.i3(lin.shape), L.i3(lout.shape),
                            L.i3(m.kernel_size), L.i3(m.stride), L.i3(m.padding), L.i3(m.dilation), L.ptr(lin.keys),
                            L.ptr(lin.vals), lin.hcap, L.ptr(lout.coors), L.ptr(lout.n_dev), lout.cap,
                            L.ptr(lyr["rb"]["nbr"]), L.ptr(lout.keys), L.ptr(lout.vals), lout.hcap,
                            L.ptr(self.rb_ws), self.rb_ws_bytes, L.ptr(self.status), st), "b2s_rulebook_conv")
                self._mark("sparse_conv%d" % lyr["index"])
                if lyr["tc"]:
                    if "in_split" in lyr:
                        hi, lo = lyr["in_split"]
                        L.check(lib.b2s_split_tf32(L.ptr(feats), L.ptr(hi), L.ptr(lo), L.ptr(lin.n_dev), lin.cap,
                                                   lyr["cin"], st), "b2s_split_tf32")
                        feats, feats_lo = hi, lo
                    L.check(lib.b2s_sparse_conv_tc(L.ptr(feats), L.ptr(feats_lo), lin.cap, lyr["cin"], L.ptr(lyr["w_hi"]),
                                                   L.ptr(lyr["w_lo"]), L.ptr(lyr["rb"]["nbr"]), lyr["K"],
                                                   L.ptr(lout.n_dev), lout.cap, L.ptr(lyr["scale"]),
                                                   L.ptr(lyr["shift"]), 1 if lyr["relu"] else 0, L.ptr(lyr["out"]),
                                                   L.ptr(lyr["out_lo"]), lyr["cout"], st), "b2s_sparse_conv_tc")
                    feats, feats_lo = lyr["out"], lyr["out_lo"]
                else:
                    L.check(lib.b2s_sparse_conv(L.ptr(feats), lyr["cin"], L.ptr(lyr["w"]), L.ptr(lyr["rb"]["nbr"]),
                                                lyr["K"], L.ptr(lout.n_dev), lout.cap, L.ptr(lyr["scale"]),
                                                L.ptr(lyr["shift"]), 1 if lyr["relu"] else 0, L.ptr(lyr["out"]),
                                                lyr["cout"], st), "b2s_sparse_conv")
                    feats, feats_lo = lyr["out"], None
            if feats_lo is not None:      # back to plain fp32 rows for the BEV scatter (hi + lo is exact)
                lout = self.layers[-1]["out_level"]
                L.check(lib.b2s_merge_hilo(L.ptr(feats), L.ptr(feats_lo), L.ptr(self.merged_out), L.ptr(lout.n_dev),
                                           lout.cap, self.layers[-1]["cout"], st), "b2s_merge_hilo")
                feats = self.merged_out
        fl = self.final_level
        D, H, W = fl.shape
        if self.rpn_impl == "tc":
            self._launch_tc_tail(feats)
            return
        self._mark("to_bev")
        L.check(lib.b2s_to_bev(L.ptr(feats), L.ptr(fl.coors), L.ptr(fl.n_dev), fl.cap, self.feat_final_c, self.B,
                               D, H, W, L.ptr(self.bev), 0, st), "b2s_to_bev")
        self._mark("rpn")
        rpn = self.net.rpn
        x = rpn.backbone(self.bev)
        box = rpn.conv_box(x).contiguous()
        cls = rpn.conv_cls(x).contiguous()
        dirp = rpn.conv_dir_cls(x).contiguous() if cfg.use_direction_classifier else None
        self._keep = (x, box, cls, dirp)
        self._mark("decode_filter")
        L.check(lib.b2s_decode_filter(
            L.ptr(box), L.ptr(cls), L.ptr(dirp), L.ptr(self.anchors), None, self.B, self.a_loc, self.fH, self.fW,
            self.code, cfg.num_class, cfg.num_direction_bins, float(cfg.nms_score_threshold), L.ptr(self.cand_box),
            L.ptr(self.cand_score), L.ptr(self.cand_label), L.ptr(self.cand_dir), L.ptr(self.cand_anchor),
            L.ptr(self.cand_count), self.cand_cap, L.ptr(self.status), st), "b2s_decode_filter")
        self._mark("nms")
        L.check(lib.b2s_nms(
            L.ptr(self.cand_box), L.ptr(self.cand_score), L.ptr(self.cand_label), L.ptr(self.cand_dir),
            L.ptr(self.cand_anchor), L.ptr(self.cand_count), self.B, self.cand_cap, self.code,
            1 if cfg.use_rotate_nms else 0, self.pre_max, self.post_max, float(cfg.nms_iou_threshold),
            1 if cfg.use_direction_classifier else 0, float(cfg.direction_offset),
            float(cfg.direction_limit_offset), cfg.num_direction_bins, self.range_host, L.ptr(self.det),
            L.ptr(self.det_count), L.ptr(self.nms_ws), self.nms_ws_bytes, st), "b2s_nms")
        self._mark("end")

    def _launch_tc_tail(self, feats):
        """BEV (NHWC, halo, hi/lo) -> tcgen05 RPN layers -> packed heads -> decode/filter -> NMS."""
        L, lib, cfg = self._L, self.lib, self.cfg
        st = L.stream()
        fl = self.final_level
        D, H, W = fl.shape
        self._mark("to_bev")
        L.check(lib.b2s_to_bev_tc(L.ptr(feats), L.ptr(fl.coors), L.ptr(fl.n_dev), fl.cap, self.feat_final_c, self.B,
                                  D, H, W, L.ptr(self.tc_bev[0]), L.ptr(self.tc_bev[1]), st), "b2s_to_bev_tc")
        self._mark("rpn")
        marked_tail = False
        for op in self.tc_plan:
            if not op["v2"] and op["kind"] != "block" and not marked_tail:
                self._mark("rpn_1x1")          # the 3x3 stack (k_conv3x3_tc2) is timed apart from the deblock/heads tail
                marked_tail = True
            src, dst = self.tc_bufs[op["src"]], self.tc_bufs[op["dst"]]
            cdst = dst[0].shape[-1]                                   # channels per pixel of the destination map
            o_hi = ctypes_ptr(dst[0].data_ptr() + 4 * op["dst_coff"])
            o_lo = ctypes_ptr(dst[1].data_ptr() + 4 * op["dst_coff"]) if op["planes"] == 2 else None
            L.check(lib.b2s_conv2d_tc_ex(
                L.ptr(src[0]), L.ptr(src[1]), self.B, op["Hin"], op["Win"], op["cin"], L.ptr(op["w_hi"]),
                L.ptr(op["w_lo"]), op["kh"], op["kw"], op["stride"], op["pad"], op["cout"], op["n_pad"],
                L.ptr(op["scale"]) if op["scale"] is not None else None,
                L.ptr(op["shift"]) if op["shift"] is not None else None, 1 if op["relu"] else 0, op["Hg"], op["Wg"],
                o_hi, o_lo, op["Hout"], op["Wout"], 1 if op["padded"] else 0, cdst, op["out_mul"], op["off_h"],
                op["off_w"], st), "b2s_conv2d_tc_ex(%s)" % op["kind"])
        if not marked_tail:
            self._mark("rpn_1x1")
        S = self.tc_head_stride
        self._mark("decode_filter")
        offs = self.tc_prog["heads"]["offsets"]
        heads = self.tc_heads
        esz = heads.element_size()
        box_p = ctypes_ptr(heads.data_ptr() + offs[0] * esz)
        cls_p = ctypes_ptr(heads.data_ptr() + offs[1] * esz)
        dir_p = ctypes_ptr(heads.data_ptr() + offs[2] * esz) if cfg.use_direction_classifier else None
        bs = self.fH * self.fW * S
        L.check(lib.b2s_decode_filter_strided(
            box_p, cls_p, dir_p, bs, bs, bs, 1, S, L.ptr(self.anchors), None, self.B, self.a_loc, self.fH, self.fW,
            self.code, cfg.num_class, cfg.num_direction_bins, float(cfg.nms_score_threshold), L.ptr(self.cand_box),
            L.ptr(self.cand_score), L.ptr(self.cand_label), L.ptr(self.cand_dir), L.ptr(self.cand_anchor),
            L.ptr(self.cand_count), self.cand_cap, L.ptr(self.status), st), "b2s_decode_filter_strided")
        self._launch_nms(st)

    def _launch_nms(self, st):
        L, lib, cfg = self._L, self.lib, self.cfg
        self._mark("nms")
        L.check(lib.b2s_nms(
            L.ptr(self.cand_box), L.ptr(self.cand_score), L.ptr(self.cand_label), L.ptr(self.cand_dir),
            L.ptr(self.cand_anchor), L.ptr(self.cand_count), self.B, self.cand_cap, self.code,
            1 if cfg.use_rotate_nms else 0, self.pre_max, self.post_max, float(cfg.nms_iou_threshold),
            1 if cfg.use_direction_classifier else 0, float(cfg.direction_offset),
            float(cfg.direction_limit_offset), cfg.num_direction_bins, self.range_host, L.ptr(self.det),
            L.ptr(self.det_count), L.ptr(self.nms_ws), self.nms_ws_bytes, st), "b2s_nms")
        self._mark("end")

    # ---------------------------------------------------------------- instrumentation
    _marks = None

    def _mark(self, name):
        if self._marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    def run_timed(self, iters=5):
        """eager (no graph) replay with a CUDA event before every stage; returns {stage: mean ms}."""
        self.P_cap_used = self.P_cap
        acc = {}
        with torch.no_grad():
            self._launch()     # warm
            for _ in range(iters):
                self._marks = []
                self._launch()
                torch.cuda.synchronize()
                for (name, ev), (_, nxt) in zip(self._marks[:-1], self._marks[1:]):
                    acc[name] = acc.get(name, 0.0) + ev.elapsed_time(nxt)
                self._marks = None
        return {k: v / iters for k, v in acc.items()}

    def kernel_launches_per_step(self):
        """how many of OUR kernels one pipeline pass launches (memsets and cuDNN kernels not counted)."""
        n = 7                                    # b2s_voxelize: insert, flag_scan, scan_sums, assign, fill, gather, finish
        if self.is_pillars:
            n += 1                               # b2s_pfn
        for lyr in self.layers:
            if lyr["build_rb"]:
                n += 1 if lyr["conv"].subm else 6   # subm_nbr | mark, popc_scan, scan_sums, emit, hash_build, conv_nbr
            n += 1                               # b2s_sparse_conv / b2s_sparse_conv_tc
            if lyr.get("in_split") is not None:
                n += 1                           # b2s_split_tf32
        if getattr(self, "any_sparse_tc", False):
            n += 1                               # b2s_merge_hilo
        if self.rpn_impl == "tc":
            n += len(self.tc_plan)               # one tcgen05 conv kernel per RPN layer (heads = 1 launch)
        return n + 1 + 1 + 3                     # to_bev, decode_filter, nms: select_sort + iou_mask + reduce

    def sparse_layer_stats(self):
        """sync; per sparse layer: rows in/out, pairs, algorithmic bytes and flops (SURVEY.md §8d formulas)."""
        out = []
        for lyr in self.layers:
            n_out = int(lyr["out_level"].n_dev[0].item())
            n_in = int(lyr["in_level"].n_dev[0].item())
            pairs = int((lyr["rb"]["nbr"][:n_out] >= 0).sum().item())
            cin, cout, K = lyr["cin"], lyr["cout"], lyr["K"]
            out.append({"index": lyr["index"], "subm": bool(lyr["conv"].subm), "cin": cin, "cout": cout, "K": K,
                        "n_in": n_in, "n_out": n_out, "pairs": pairs,
                        "bytes": 4 * (n_in * cin + n_out * cout) + 8 * pairs + 4 * K * cin * cout,
                        "flops": 2 * pairs * cin * cout})
        return out

    def rpn_layer_stats(self):
        """per tcgen05 RPN layer: algorithmic fp32 flops (2*pixels*taps*cin*cout; the 3xTF32 split issues 3x that on
        the tensor pipe) and algorithmic bytes (hi/lo planes in and out + weights)."""
        if self.rpn_impl != "tc":
            return []
        out = []
        for i, op in enumerate(self.tc_plan):
            cin, cout, taps = op["cin"], op["cout"], op["taps"]
            px = self.B * op["Hg"] * op["Wg"]
            px_in = self.B * op["Hin"] * op["Win"]
            out.append({"index": i, "kind": op["kind"], "v2": op["v2"], "cin": cin, "cout": cout, "taps": taps,
                        "pixels": px, "flops": 2 * px * taps * cin * cout,
                        "bytes": 4 * (2 * px_in * cin + op["planes"] * px * cout + 2 * taps * cin * op["n_pad"])})
        return out

    # ---------------------------------------------------------------- public API
    def load_points(self, frames):
        """frames: list of B CUDA (or pinned/CPU) float32 [P_i, F] tensors -> static input buffers."""
        assert len(frames) == self.B
        sizes = [int(f.shape[0]) for f in frames]
        total = sum(sizes)
        assert total <= self.P_cap, "more points (%d) than the engine's capacity (%d)" % (total, self.P_cap)
        off = 0
        for f, n in zip(frames, sizes):
            if n:
                self.points[off:off + n].copy_(f, non_blocking=True)
            off += n
        offs = torch.tensor(np.cumsum([0] + sizes), dtype=torch.int32)
        self.offsets.copy_(offs.pin_memory() if offs.device.type == "cpu" else offs, non_blocking=True)
        return total

    def run(self):
        """launch the pipeline on the points currently in the static buffers (async)."""
        # the grid sizes of the per-point kernels are fixed at capacity so one graph serves any frame mix
        self.P_cap_used = self.P_cap
        if not self.use_graph:
            with torch.no_grad():
                self._launch()
            return self.det, self.det_count
        if self._graph is None:
            with torch.no_grad():
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for _ in range(3):          # warm-up: cuDNN autotune, lazy module init, func attributes
                        self._launch()
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch()
                self._graph = g
        self._graph.replay()
        return self.det, self.det_count

    def infer(self, frames):
        self.load_points(frames)
        return self.run()

    def check_status(self):
        """sync + raise on data-dependent overflow (call when results are read back)."""
        st = int(self.status.item())
        if st & ~1:   # bit 1 (voxel overflow) is the reference's own drop-extra-voxels behaviour
            raise RuntimeError("b2second engine: " + self._L.status_message(st))
        return st

    def detections(self):
        """sync and convert to the reference's list-of-dicts (voxelnet.py:616-645)."""
        det = self.det.cpu()
        cnt = self.det_count.cpu().tolist()
        self.check_status()
        out = []
        for b in range(self.B):
            d = det[b, :cnt[b]]
            out.append({"box3d_lidar": d[:, :self.code].clone(), "scores": d[:, self.code].clone(),
                        "label_preds": d[:, self.code + 1].long(), "metadata": None})
        return out
