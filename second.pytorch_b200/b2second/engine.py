"""Fused device pipeline for ``VoxelNet`` inference: points in HBM -> detections in HBM, no host sync.

This is the fast path behind the ``VoxelNet.forward(example)`` contract (reached through
``b2second.fastpath.accelerate(net)`` for a reference-built network and through ``models.VoxelNet.forward`` for
the mirror).  Where the module-by-module path mirrors the reference call by call -- one rulebook sync per
strided conv, a D2H -> CPU NMS -> H2D round trip per frame (second/pytorch/core/box_torch_ops.py:503,512) -- the
engine keeps every data-dependent count in device memory and runs a fixed launch sequence over capacity-sized
buffers, so the whole frame batch is ONE CUDA graph:

  b2s_voxelize (+ fused SimpleVoxel mean)                     voxelnet.py:325-328, preprocess.py:303-315
     or, for an example that arrives already voxelised (the reference's own dict):
     b2s_hash_build + b2s_vfe_mean over the caller's voxels   voxel_encoder.py:220-225,246-255
  per sparse layer: b2s_rulebook_{subm,conv} (cached per indice_key) + b2s_sparse_conv_tc (tcgen05, fp16 hi/lo
      planes flow from layer to layer) with BatchNorm1d/ReLU folded into the epilogue      middle.py:145-192
  b2s_pfn (PointPillars)                                       pointpillars.py:203-237
  b2s_to_bev_tc (NHWC + halo, fp16 hi/lo)                      middle.py:206-209 / pointpillars.py:444-476
  RPN as a program of b2s_conv2d_tc_ex launches (tc.plan_rpn)  rpn.py:314-331,393-420,467-497
      (rpn_impl="cudnn" keeps the torch modules as a cross-check: fp32, TF32 off)
  b2s_decode_filter_strided + b2s_nms (+ direction/range epilogue)   voxelnet.py:377-645

The network is read through ``b2second.spec.spec_from_module`` (attribute names of the reference), never through
``b2second.models`` classes or a ``ModelConfig``.

Output per batch: ``det [B, post_max, code+2]`` (box, score, label) + ``det_count [B]`` -- the fixed-stride
record the multi-GPU path all-gathers; ``det`` and ``det_count`` live in ONE buffer
(``det_record [B, post_max*(code+2) + 1]``) so the all-gather needs no packing step.
"""
import ctypes
import os

import numpy as np
import torch

from . import spec as _spec
from . import tc as _tc


def _pow2_at_least(n):
    c = 1024
    while c < n:
        c <<= 1
    return c


def ctypes_ptr(addr):
    return ctypes.c_void_p(int(addr))


class _Level:
    """one active-site set: coordinates, device count, coordinate->row hash, row capacity."""
    __slots__ = ("coors", "n_dev", "keys", "vals", "hcap", "cap", "shape")


class InferenceEngine:
    def __init__(self, net, batch_size=1, max_points=None, max_voxels=None, row_cap_factor=2.0,
                 cand_cap=None, use_cuda_graph=True, rpn_impl="auto", sparse_impl="tc", device=None):
        import spconv as sp                      # the CUDA drop-in: fails loudly if the library is missing
        assert not getattr(sp, "__oracle__", False), "the engine is the product path; it never runs on the oracle"
        self.sp = sp
        self.lib = sp._lib.load()
        self._L = sp._lib
        self.spec = s = net if isinstance(net, _spec.NetSpec) else _spec.spec_from_module(net, max_voxels)
        if max_voxels:
            s.max_voxels = int(max_voxels)
        self.B = int(batch_size)
        dev = torch.device(device) if device is not None else s.device
        assert dev.type == "cuda", "InferenceEngine needs the network on a CUDA device"
        self.dev = dev
        # parity bar is fp32: keep cuDNN/cuBLAS off TF32 (torch allows TF32 convs by default)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True
        self.F = s.num_point_features
        self.T = s.max_points_per_voxel
        self.max_voxels = s.max_voxels
        self.P_cap = int(max_points or 40000) * self.B
        self.grid = s.grid_size
        self.code = s.box_code_size
        self.use_graph = use_cuda_graph
        self._graphs = {}
        # RPN: "tc" = hand-written tcgen05 implicit GEMM (csrc/conv_tc.cu, 3xF16), "cudnn" = torch/cuDNN fp32.
        if rpn_impl == "auto":
            rpn_impl = "tc" if _tc.supported(s.rpn) else "cudnn"
        assert rpn_impl in ("tc", "cudnn")
        if rpn_impl == "tc" and not _tc.supported(s.rpn):
            raise ValueError("rpn_impl='tc': this RPN has a layer b2s_conv2d_tc_ex does not cover (channels % 64, kernel > 4)")
        self.rpn_impl = rpn_impl
        assert sparse_impl in ("tc", "fma")       # sparse-conv inner product: tcgen05 3xF16 | fp32 FMA tiles
        self.sparse_impl = sparse_impl
        with torch.cuda.device(dev):
            self.status = torch.zeros(1, dtype=torch.int32, device=dev)
            self._plan_middle(row_cap_factor)
            self._alloc_voxel_buffers()
            if rpn_impl == "tc":
                self._alloc_tc_rpn()
            self._alloc_detect_buffers(cand_cap)

    # ---------------------------------------------------------------- planning / allocation
    def _new_level(self, cap, shape, with_storage=True):
        lv = _Level()
        lv.cap = int(cap)
        lv.shape = [int(x) for x in shape]
        lv.hcap = _pow2_at_least(2 * lv.cap)
        if with_storage:
            lv.coors = torch.zeros(lv.cap, 4, dtype=torch.int32, device=self.dev)
            lv.n_dev = torch.zeros(1, dtype=torch.int32, device=self.dev)
            lv.keys = torch.empty(lv.hcap, dtype=torch.int64, device=self.dev)
            lv.vals = torch.empty(lv.hcap, dtype=torch.int32, device=self.dev)
        return lv

    def _hilo_rows(self, cap, c):
        """interleaved fp16 rows [cap][hi c | lo c] -> (buffer, hi view, lo view, row stride in halves)."""
        buf = torch.zeros(cap, 2, c, dtype=torch.float16, device=self.dev)
        return buf, buf[:, 0], buf[:, 1], 2 * c

    def _plan_middle(self, row_cap_factor):
        s, sp, dev = self.spec, self.sp, self.dev
        self.is_pillars = s.is_pillars
        self.layers = []
        self.rb_ws = None
        cap0 = self.B * self.max_voxels
        if self.is_pillars:
            self.level0 = self._new_level(cap0, s.sparse_shape, with_storage=False)
            p = s.pfn
            self.pfn_w = p["w"].to(dev)
            self.pfn_scale, self.pfn_shift = p["scale"].to(dev), p["shift"].to(dev)
            self.pfn_cout = p["cout"]
            if p["with_distance"]:
                raise _spec.UnsupportedNetwork("PillarFeatureNet with_distance (off in every BASELINE config)")
            self.pfn_geom = (p["vx"], p["vy"], p["x_offset"], p["y_offset"])
            self.feat_final_c = self.pfn_cout
            self.final_level = self.level0
            self.pfn_out = torch.zeros(cap0, self.pfn_cout, dtype=torch.float32, device=dev)
            return
        self.level0 = self._new_level(cap0, s.sparse_shape, with_storage=False)   # storage = voxelizer outputs
        level = self.level0
        rulebooks = {}
        max_ws = 0
        thin_ok = os.environ.get("B2S_THIN_TC", "1") != "0"      # A/B switch: thin layers on the FMA core
        plan_mode = int(os.environ.get("B2S_SP_PLAN", "4"))
        for j, ls in enumerate(s.layers):
            K = ls["K"]
            lyr = dict(ls)
            lyr["index"] = j
            lyr["w"] = ls["w"].to(dev)
            lyr["scale"] = None if ls["scale"] is None else ls["scale"].to(dev)
            lyr["shift"] = None if ls["shift"] is None else ls["shift"].to(dev)
            if ls["subm"]:
                key = ("subm", ls["indice_key"], id(level)) if ls["indice_key"] is not None else ("subm", j)
                if key not in rulebooks:
                    rulebooks[key] = {"nbr": torch.empty(level.cap, K, dtype=torch.int32, device=dev)}
                    lyr["build_rb"] = True
                    # the strided conv right before this layer produced `level`: its occupancy bitmap is still in the
                    # rulebook workspace, so this SubM table is built by rank lookups (b2s_rulebook_subm_ranked) and
                    # the level needs no hash table
                    prev = self.layers[-1] if self.layers else None
                    lyr["ranked"] = bool(prev is not None and not prev["subm"] and prev["out_level"] is level
                                         and os.environ.get("B2S_RB_RANKED", "1") != "0")
                else:
                    lyr["build_rb"] = False
                lyr["rb"] = rulebooks[key]
                lyr["in_level"], lyr["out_level"] = level, level
            else:
                out_shape = sp.ops.get_conv_output_size(level.shape, ls["kernel_size"], ls["stride"], ls["padding"],
                                                        ls["dilation"])
                cells = int(np.prod(out_shape))
                fan = int(np.prod([-(-k // st) for k, st in zip(ls["kernel_size"], ls["stride"])]))
                per_frame = min(cells, int(self.max_voxels * row_cap_factor), level.cap // self.B * fan)
                new = self._new_level(self.B * per_frame, out_shape)
                lyr["rb"] = {"nbr": torch.empty(new.cap, K, dtype=torch.int32, device=dev)}
                lyr["build_rb"] = True
                lyr["in_level"], lyr["out_level"] = level, new
                max_ws = max(max_ws, self.lib.b2s_rulebook_conv_workspace_bytes(self.B, self._L.i3(out_shape)))
                level = new
            # tile plan of the rulebook (csrc/sparse_plan.cu): which kernel offsets each 128-row tile needs, and an order
            # of the rows that makes that set small.  B2S_SP_PLAN: 0 off, 1 masks only, 2 rows grouped for SubM
            # rulebooks (shared by 2-3 layers) + masks for the strided ones, 3 rows grouped for every 27-offset rulebook,
            # 4 (default) SubM rulebooks only (a strided conv uses its table once: the plan costs more than it saves)
            rb = lyr["rb"]
            if "tile_mask" not in rb and plan_mode > 0 and (plan_mode != 4 or ls["subm"]):
                cap_rb = lyr["out_level"].cap
                rb["tile_mask"] = torch.zeros((cap_rb + 127) // 128, dtype=torch.int32, device=dev)
                rb["row_mask"] = torch.zeros(cap_rb, dtype=torch.int32, device=dev)    # written by the rulebook builder
                rb["sort"] = bool(K > 3 and (plan_mode == 3 or (plan_mode in (2, 4) and ls["subm"])))
                rb["perm"] = torch.zeros(cap_rb, dtype=torch.int32, device=dev) if rb["sort"] else None
            # tensor-pipe core (csrc/sparse_conv_tc.cu): the library says which (Cin, Cout) it was built for;
            # 3-/4-feature input layers are zero-padded to 8 channels.  Everything else: fp32 FMA core.
            cin_tc = _tc.sparse_tc_cin(ls["cin"])
            lyr["cin_tc"] = cin_tc
            lyr["tc"] = bool(self.sparse_impl == "tc" and cin_tc is not None and (thin_ok or cin_tc >= 32) and
                             self.lib.b2s_sparse_conv_tc_supported(cin_tc, ls["cout"]))
            self.layers.append(lyr)
        # a level's coordinate hash is only built when some SubM rulebook still looks rows up by hash
        for lyr in self.layers:
            if not lyr["subm"]:
                lyr["want_hash"] = any(l["subm"] and l.get("build_rb") and l["in_level"] is lyr["out_level"]
                                       and not l.get("ranked") for l in self.layers)
        # once a layer runs on the tensor pipe all later ones must too (hi/lo planes flow forward)
        seen_tc = False
        for lyr in self.layers:
            if seen_tc and not lyr["tc"]:
                for l2 in self.layers:
                    l2["tc"] = False
                break
            seen_tc = seen_tc or lyr["tc"]
        for j, lyr in enumerate(self.layers):
            cap = lyr["out_level"].cap
            if lyr["tc"]:
                wp = _tc.pack_sparse_weights(lyr["w"])
                ws = _tc.pow2_scale(wp)
                lyr["w_hi"], lyr["w_lo"] = _tc.split_f16(wp, ws)
                base = lyr["scale"] if lyr["scale"] is not None else torch.ones(lyr["cout"], device=dev)
                lyr["scale_tc"] = (base.float() / ws).contiguous()
                lyr["out_buf"], lyr["out_hi"], lyr["out_lo"], lyr["out_stride"] = self._hilo_rows(cap, lyr["cout"])
                if j == 0 or not self.layers[j - 1]["tc"]:
                    # first tensor-pipe layer: its fp32 input rows are split into fp16 hi/lo planes first
                    lyr["in_split"] = self._hilo_rows(lyr["in_level"].cap, lyr["cin_tc"])
            else:
                lyr["out"] = torch.zeros(cap, lyr["cout"], dtype=torch.float32, device=dev)
        self.any_sparse_tc = any(l["tc"] for l in self.layers)
        self.rb_ws = torch.empty(max(max_ws, 1), dtype=torch.uint8, device=dev)
        self.rb_ws_bytes = max_ws
        self.final_level = level
        self.feat_final_c = self.layers[-1]["cout"]

    def _alloc_voxel_buffers(self):
        lib, dev, s = self.lib, self.dev, self.spec
        cap = self.B * self.max_voxels
        self.points = torch.zeros(self.P_cap, self.F, dtype=torch.float32, device=dev)
        self.offsets = torch.zeros(self.B + 1, dtype=torch.int32, device=dev)
        self.vox_coors = torch.zeros(cap, 4, dtype=torch.int32, device=dev)
        self.vox_num = torch.zeros(cap, dtype=torch.int32, device=dev)
        self.vox_slots = torch.zeros(cap, self.T, dtype=torch.int32, device=dev)
        self.num_voxels = torch.zeros(1 + self.B, dtype=torch.int32, device=dev)
        # the voxelizer's hash doubles as level 0's coordinate->row locator: sized for the points it must absorb AND
        # (pre-voxelised entry, b2s_hash_build) for 2x the voxel rows
        self.vox_hcap = max(lib.b2s_voxelize_hash_capacity(self.P_cap), _pow2_at_least(2 * cap))
        self.vox_keys = torch.empty(self.vox_hcap, dtype=torch.int64, device=dev)
        self.vox_vals = torch.empty(self.vox_hcap, dtype=torch.int32, device=dev)
        # the workspace holds one int per hash slot: query it for a point count whose own hash capacity is vox_hcap
        self.vox_ws_bytes = lib.b2s_voxelize_workspace_bytes(max(self.P_cap, self.vox_hcap // 2), self.B,
                                                             self.max_voxels, self.T)
        self.vox_ws = torch.empty(max(self.vox_ws_bytes, 1), dtype=torch.uint8, device=dev)
        if s.vfe_kind == "mean":
            self.vfe_mode, self.vfe_nf, c = 1, s.vfe_num_features, s.vfe_num_features
        elif s.vfe_kind == "mean_radius":
            self.vfe_mode, self.vfe_nf, c = 2, s.vfe_num_features, s.vfe_num_features - 1
        else:
            self.vfe_mode, self.vfe_nf, c = 0, self.F, 0
        self.vfe_out = torch.zeros(cap, c, dtype=torch.float32, device=dev) if c else None
        # level 0 aliases the voxelizer outputs
        lv = self.level0
        lv.coors, lv.n_dev, lv.keys, lv.vals, lv.hcap = (self.vox_coors, self.num_voxels, self.vox_keys,
                                                         self.vox_vals, self.vox_hcap)
        C = self.feat_final_c
        D, H, W = self.final_level.shape
        self.bev_cd = C * D
        if self.spec.bev_channels is None:
            self.spec.bev_channels = C * D
        if self.rpn_impl == "cudnn":
            self.bev = torch.zeros(self.B, C * D, H, W, dtype=torch.float32, device=dev)
        # pre-voxelised entry (the reference's own example dict): the caller's voxels land here
        self.in_voxels = None

    def _alloc_tc_rpn(self):
        """buffers of the tensor-core RPN: NHWC halo-padded fp16 hi/lo planes (halo zero-filled once, never written)."""
        dev, B = self.dev, self.B
        D, H, W = self.final_level.shape
        C = self.feat_final_c * D
        rpn = self.spec.rpn
        if next(rpn.parameters()).device != dev:
            raise ValueError("the RPN's parameters must live on the engine's device")
        plan = _tc.plan_rpn(rpn, H, W)
        self.tc_prog = plan
        self.tc_plan = plan["ops"]
        # single-scale tail (one k = s = 1 deblock 128 -> 128 feeding the heads): one fused kernel, the deblock output
        # never goes to HBM (csrc/rpn_tail.cu; bit-identical to the two launches).  B2S_RPN_TAIL=0: two launches.
        self.tc_tail_fused = bool(os.environ.get("B2S_RPN_TAIL", "1") != "0" and _tc.fusable_tail(plan))
        assert plan["in_channels"] == C, "BEV channels %d != RPN input %d" % (C, plan["in_channels"])

        def plane(h, w, c):
            return (torch.zeros(B, h + 2, w + 2, c, dtype=torch.float16, device=dev),
                    torch.zeros(B, h + 2, w + 2, c, dtype=torch.float16, device=dev))
        self.tc_bev = plane(H, W, C)
        self.tc_bufs = {"in": self.tc_bev}
        for name, (h, w, c) in plan["buffers"].items():
            self.tc_bufs[name] = plane(h, w, c)
        hd = plan["heads"]
        self.tc_head_stride = hd["stride"]
        self.tc_heads = torch.zeros(B, hd["H"], hd["W"], self.tc_head_stride, dtype=torch.float32, device=dev)
        self.tc_bufs["heads"] = (self.tc_heads, None)
        self.tc_hw = (H, W)
        # background tiles (csrc/rpn_bg.cu): through the leading chain of 3x3 stride-1 layers, output tiles whose whole
        # receptive field is empty BEV are not computed but copied from the layer's response to an empty frame
        self.bg_idx = _tc.background_layers(plan) if os.environ.get("B2S_RPN_BG", "1") != "0" else []
        self.bg_fused_fill = os.environ.get("B2S_RPN_BG_FILL", "fused") == "fused"
        if self.bg_idx:
            nl = len(self.bg_idx)
            self.bg_tiles = B * (-(-H // 16)) * (-(-W // 16))
            self.bg_occ = torch.zeros(B, H, W, dtype=torch.uint8, device=dev)
            self.bg_scratch = torch.zeros(2, B, H, W, dtype=torch.uint8, device=dev)
            self.bg_flags = torch.zeros(nl, self.bg_tiles, dtype=torch.int32, device=dev)
            self.bg_work = torch.zeros(nl, self.bg_tiles, dtype=torch.int32, device=dev)
            self.bg_list = torch.zeros(nl, self.bg_tiles, dtype=torch.int32, device=dev)
            self.bg_counts = torch.zeros(nl, 2, dtype=torch.int32, device=dev)
            self.bg_field = self._empty_frame_response()

    def _empty_frame_response(self):
        """per background-tracked layer: what the layer's own kernel produces for an empty frame, as one halo-padded
        frame [H+2, W+2, C] of fp16 hi/lo planes.  Run once at plan time on frame 0 of the (still all-zero) buffers."""
        L, lib = self._L, self.lib
        st = L.stream()
        fields = []
        for oi in self.bg_idx:
            op = self.tc_plan[oi]
            src, dst = self.tc_bufs[op["src"]], self.tc_bufs[op["dst"]]
            cdst = dst[0].shape[-1]
            L.check(lib.b2s_conv2d_tc_ex(
                L.ptr(src[0]), L.ptr(src[1]), 1, op["Hin"], op["Win"], op["cin"], L.ptr(op["w_hi"]),
                L.ptr(op["w_lo"]), op["kh"], op["kw"], op["stride"], op["pad"], op["cout"], op["n_pad"],
                L.ptr(op["scale"]), L.ptr(op["shift"]) if op["shift"] is not None else None,
                1 if op["relu"] else 0, op["Hg"], op["Wg"], L.ptr(dst[0]), L.ptr(dst[1]), op["Hout"], op["Wout"],
                1, cdst, op["out_mul"], op["off_h"], op["off_w"], None, None, None, None, None, None,
                L.ptr(self.status), st), "b2s_conv2d_tc_ex(empty frame)")
            fields.append((dst[0][0, :, :, :op["cout"]].clone().contiguous(),
                           dst[1][0, :, :, :op["cout"]].clone().contiguous()))
        torch.cuda.synchronize(self.dev)
        for oi in self.bg_idx:                    # leave the activation buffers as they were (all zero)
            for t in self.tc_bufs[self.tc_plan[oi]["dst"]]:
                t[0].zero_()
        return fields

    def _feature_hw(self):
        if self.rpn_impl == "tc":
            return self.tc_prog["heads"]["H"], self.tc_prog["heads"]["W"]
        D, H, W = self.final_level.shape
        with torch.no_grad():
            y = self.spec.rpn.conv_box(self._rpn_backbone(torch.zeros(1, self.bev_cd, H, W, device=self.dev)))
        return int(y.shape[2]), int(y.shape[3])

    def _rpn_backbone(self, x):
        """RPNNoHeadBase.forward (rpn.py:314-331) on the torch modules: cuDNN cross-check path only."""
        rpn = self.spec.rpn
        ups = []
        for i in range(len(rpn.blocks)):
            x = rpn.blocks[i](x)
            if i - rpn._upsample_start_idx >= 0:
                ups.append(rpn.deblocks[i - rpn._upsample_start_idx](x))
        if len(ups) > 0:
            x = torch.cat(ups, dim=1)
        return x

    def _alloc_detect_buffers(self, cand_cap):
        s, dev = self.spec, self.dev
        self.fH, self.fW = self._feature_hw()
        self.a_loc = s.num_anchors_per_loc
        self.A = self.a_loc * self.fH * self.fW
        self.anchors = torch.zeros(self.A, self.code, dtype=torch.float32, device=dev)
        self._anchors_key = None
        try:
            a = _spec.anchors_for(s, (self.fH, self.fW))
        except Exception:
            a = None                                  # anchors then come with the first example (set_anchors)
        if a is not None:
            assert a.shape[0] == self.A, "anchor count %d != RPN output %d" % (a.shape[0], self.A)
            self.anchors.copy_(torch.from_numpy(a))
            self._anchors_key = "generated"
        self.anchors_mask = None                      # [B, A] uint8, allocated on first use
        self.use_mask = False
        self.cand_cap = int(cand_cap or min(self.A, 32768))
        B, cc, code = self.B, self.cand_cap, self.code
        # per-class NMS branch (voxelnet.py:458-547): one candidate list and one NMS per (class, frame) -- "virtual
        # frames" v = c*B + b -- then the per-class results are concatenated in class order
        self.mc = bool(s.multiclass_nms)
        self.ncls = s.num_class
        V = self.ncls * B if self.mc else B
        self.cand_box = torch.zeros(V, cc, code, dtype=torch.float32, device=dev)
        self.cand_score = torch.zeros(V, cc, dtype=torch.float32, device=dev)
        self.cand_label = torch.zeros(V, cc, dtype=torch.int32, device=dev)
        self.cand_dir = torch.zeros(V, cc, dtype=torch.int32, device=dev)
        self.cand_anchor = torch.zeros(V, cc, dtype=torch.int32, device=dev)
        self.cand_count = torch.zeros(V, dtype=torch.int32, device=dev)
        self.score_thresh = float(s.nms_score_thresholds[0])
        self.pre_max, self.post_max = int(s.nms_pre_max_sizes[0]), int(s.nms_post_max_sizes[0])
        self.iou_thresh = float(s.nms_iou_thresholds[0])
        if self.mc:
            n = self.ncls
            assert n <= 16 and all(len(x) >= n for x in (s.nms_score_thresholds, s.nms_pre_max_sizes,
                                                         s.nms_post_max_sizes, s.nms_iou_thresholds))
            self.mc_thresh = (ctypes.c_float * n)(*s.nms_score_thresholds[:n])
            self.mc_pre = [int(x) for x in s.nms_pre_max_sizes[:n]]
            self.mc_post = [int(x) for x in s.nms_post_max_sizes[:n]]
            self.mc_iou = [float(x) for x in s.nms_iou_thresholds[:n]]
            self.mc_uniform = len(set(self.mc_pre)) == 1 and len(set(self.mc_post)) == 1 and len(set(self.mc_iou)) == 1
            self.pre_max, self.post_max_class = max(self.mc_pre), max(self.mc_post)
            self.post_max = n * self.post_max_class               # rows of a frame's record
            if s.nms_class_agnostic:
                self.mc_lo = self.mc_hi = None
            else:
                counts = s.class_anchor_counts
                if counts is None or len(counts) != n or sum(counts) != self.a_loc:
                    raise _spec.UnsupportedNetwork("per-class NMS needs target_assigner anchor counts per class")
                lo = np.cumsum([0] + counts[:-1]).tolist()
                self.mc_lo = (ctypes.c_int * n)(*lo)
                self.mc_hi = (ctypes.c_int * n)(*[a + c for a, c in zip(lo, counts)])
            self.det_mc = torch.zeros(V, self.post_max_class, code + 2, dtype=torch.float32, device=dev)
            self.cnt_mc = torch.zeros(V, dtype=torch.int32, device=dev)
        self.nms_ws_bytes = self.lib.b2s_nms_workspace_bytes(V, cc, self.pre_max)
        self.nms_ws = torch.empty(max(self.nms_ws_bytes, 1), dtype=torch.uint8, device=dev)
        # one record per frame: post_max*(code+2) detection floats followed by the count (as a float: < 2^24) --
        # exactly the all-gather payload (b2second/dist.py), written in place by the NMS epilogue
        self.rec_width = self.post_max * (code + 2) + 1
        self.det_record = torch.zeros(B, self.rec_width, dtype=torch.float32, device=dev)
        self.det_count = torch.zeros(B, dtype=torch.int32, device=dev)
        rng = s.post_center_range
        self.range_host = self._L.f6(rng) if len(rng) == 6 else None

    @property
    def det(self):
        """[B, post_max, code+2] view of the detection records."""
        return self.det_record[:, :self.rec_width - 1].view(self.B, self.post_max, self.code + 2)

    # ---------------------------------------------------------------- the launch sequence
    def _launch(self, mode="points"):
        L, lib, s = self._L, self.lib, self.spec
        st = L.stream()
        self.status.zero_()
        if mode == "points":
            self._mark("voxelize")
            L.check(lib.b2s_voxelize(
                L.ptr(self.points), L.ptr(self.offsets), self.P_cap, self.F, self.B,
                L.f3(s.point_cloud_range[:3]), L.f3(s.voxel_size), L.i3(self.grid), self.T, self.max_voxels,
                L.ptr(self.vox_coors), L.ptr(self.vox_num), L.ptr(self.vox_slots), None, self.vfe_mode, self.vfe_nf,
                L.ptr(self.vfe_out), L.ptr(self.num_voxels), L.ptr(self.vox_keys), L.ptr(self.vox_vals), self.vox_hcap,
                int(self.level0.shape[0]),       # hash keys in the middle encoder's shape (grid_z + 1, middle.py:139)
                L.ptr(self.vox_ws), self.vox_ws_bytes, L.ptr(self.status), st), "b2s_voxelize")
            pfn_points, pfn_slots = self.points, self.vox_slots
        else:
            # the caller's voxels / num_points / coordinates are already in in_voxels / vox_num / vox_coors and
            # num_voxels[0] holds their count: build level 0's locator and the voxel features from them
            self._mark("voxelize")
            lv = self.level0
            if not self.is_pillars:
                L.check(lib.b2s_hash_build(L.ptr(lv.coors), L.ptr(lv.n_dev), lv.cap, L.i3(lv.shape), L.ptr(lv.keys),
                                           L.ptr(lv.vals), lv.hcap, L.ptr(self.status), st), "b2s_hash_build")
                L.check(lib.b2s_vfe_mean(L.ptr(self.in_voxels), L.ptr(self.vox_num), L.ptr(lv.n_dev), lv.cap, self.T,
                                         self.F, self.vfe_mode, self.vfe_nf, L.ptr(self.vfe_out), st), "b2s_vfe_mean")
            pfn_points, pfn_slots = self.in_voxels, self.in_slots
        if self.is_pillars:
            vx, vy, xo, yo = self.pfn_geom
            self._mark("pfn")
            L.check(lib.b2s_pfn(L.ptr(pfn_points), self.F, L.ptr(pfn_slots), L.ptr(self.vox_num),
                                L.ptr(self.vox_coors), L.ptr(self.num_voxels), self.level0.cap, self.T,
                                L.ptr(self.pfn_w), L.ptr(self.pfn_scale), L.ptr(self.pfn_shift), self.pfn_cout,
                                vx, vy, xo, yo, L.ptr(self.pfn_out), st), "b2s_pfn")
            feats, hilo = self.pfn_out, None
        else:
            feats, hilo = self.vfe_out, None          # hilo = (hi, lo, stride) once on the tensor pipe
            for lyr in self.layers:
                lin, lout = lyr["in_level"], lyr["out_level"]
                if lyr["build_rb"]:
                    self._mark("rulebook%d" % lyr["index"])
                    if lyr["subm"] and lyr.get("ranked"):
                        L.check(lib.b2s_rulebook_subm_ranked(
                            L.ptr(lin.coors), L.ptr(lin.n_dev), lin.cap, self.B, L.i3(lin.shape),
                            L.i3(lyr["kernel_size"]), L.i3(lyr["dilation"]), L.ptr(self.rb_ws), self.rb_ws_bytes,
                            L.ptr(lyr["rb"]["nbr"]), L.ptr(lyr["rb"].get("row_mask")), st), "b2s_rulebook_subm_ranked")
                    elif lyr["subm"]:
                        L.check(lib.b2s_rulebook_subm(L.ptr(lin.coors), L.ptr(lin.n_dev), lin.cap, L.i3(lin.shape),
                                                      L.i3(lyr["kernel_size"]), L.i3(lyr["dilation"]), L.ptr(lin.keys),
                                                      L.ptr(lin.vals), lin.hcap, L.ptr(lyr["rb"]["nbr"]),
                                                      L.ptr(lyr["rb"].get("row_mask")), st), "b2s_rulebook_subm")
                    else:
                        L.check(lib.b2s_rulebook_conv(
                            L.ptr(lin.coors), L.ptr(lin.n_dev), lin.cap, self.B, L.i3(lin.shape), L.i3(lout.shape),
                            L.i3(lyr["kernel_size"]), L.i3(lyr["stride"]), L.i3(lyr["padding"]), L.i3(lyr["dilation"]),
                            L.ptr(lin.keys), L.ptr(lin.vals), lin.hcap, L.ptr(lout.coors), L.ptr(lout.n_dev), lout.cap,
                            L.ptr(lyr["rb"]["nbr"]), L.ptr(lout.keys) if lyr["want_hash"] else None,
                            L.ptr(lout.vals) if lyr["want_hash"] else None, lout.hcap,
                            L.ptr(self.rb_ws), self.rb_ws_bytes, L.ptr(lyr["rb"].get("row_mask")), L.ptr(self.status), st),
                            "b2s_rulebook_conv")
                    if lyr["tc"] and "tile_mask" in lyr["rb"]:
                        rb = lyr["rb"]
                        L.check(lib.b2s_sparse_tile_plan(
                            L.ptr(rb["nbr"]), L.ptr(rb["row_mask"]), lyr["K"], L.i3(lyr["kernel_size"]), L.ptr(lout.n_dev), lout.cap,
                            1 if rb["sort"] else 0, L.ptr(rb["perm"]), L.ptr(rb["tile_mask"]), st),
                            "b2s_sparse_tile_plan")
                self._mark("sparse_conv%d" % lyr["index"])
                if lyr["tc"]:
                    if "in_split" in lyr:
                        _, hi, lo, stride = lyr["in_split"]
                        L.check(lib.b2s_split_f16(L.ptr(feats), L.ptr(hi), L.ptr(lo), L.ptr(lin.n_dev), lin.cap,
                                                  lyr["cin"], lyr["cin_tc"], stride, st), "b2s_split_f16")
                        hilo = (hi, lo, stride)
                    L.check(lib.b2s_sparse_conv_tc_plan(
                        L.ptr(hilo[0]), L.ptr(hilo[1]), hilo[2], lin.cap, lyr["cin_tc"], L.ptr(lyr["w_hi"]),
                        L.ptr(lyr["w_lo"]), L.ptr(lyr["rb"]["nbr"]), lyr["K"], L.ptr(lout.n_dev), lout.cap,
                        L.ptr(lyr["rb"].get("perm")), L.ptr(lyr["rb"].get("tile_mask")),
                        L.ptr(lyr["scale_tc"]), L.ptr(lyr["shift"]), 1 if lyr["relu"] else 0, L.ptr(lyr["out_hi"]),
                        L.ptr(lyr["out_lo"]), lyr["out_stride"], lyr["cout"], L.ptr(self.status), st),
                        "b2s_sparse_conv_tc_plan")
                    feats, hilo = None, (lyr["out_hi"], lyr["out_lo"], lyr["out_stride"])
                else:
                    L.check(lib.b2s_sparse_conv(L.ptr(feats), lyr["cin"], L.ptr(lyr["w"]), L.ptr(lyr["rb"]["nbr"]),
                                                lyr["K"], L.ptr(lout.n_dev), lout.cap, L.ptr(lyr["scale"]),
                                                L.ptr(lyr["shift"]), 1 if lyr["relu"] else 0, L.ptr(lyr["out"]),
                                                lyr["cout"], st), "b2s_sparse_conv")
                    feats, hilo = lyr["out"], None
        if self.rpn_impl == "tc":
            self._launch_tc_tail(feats, hilo)
            return
        fl = self.final_level
        D, H, W = fl.shape
        if hilo is not None:          # back to plain fp32 rows for the NCHW scatter (hi + lo is exact)
            if getattr(self, "merged_out", None) is None:
                self.merged_out = torch.zeros(fl.cap, self.feat_final_c, dtype=torch.float32, device=self.dev)
            L.check(lib.b2s_merge_f16(L.ptr(hilo[0]), L.ptr(hilo[1]), L.ptr(self.merged_out), L.ptr(fl.n_dev),
                                      fl.cap, self.feat_final_c, hilo[2], st), "b2s_merge_f16")
            feats = self.merged_out
        self._mark("to_bev")
        L.check(lib.b2s_to_bev(L.ptr(feats), L.ptr(fl.coors), L.ptr(fl.n_dev), fl.cap, self.feat_final_c, self.B,
                               D, H, W, L.ptr(self.bev), 0, st), "b2s_to_bev")
        self._mark("rpn")
        rpn = s.rpn
        x = self._rpn_backbone(self.bev)
        box = rpn.conv_box(x).contiguous()
        cls = rpn.conv_cls(x).contiguous()
        dirp = rpn.conv_dir_cls(x).contiguous() if s.use_direction_classifier else None
        self._keep = (x, box, cls, dirp)
        self._mark("decode_filter")
        HW = self.fH * self.fW
        self._launch_decode(L.ptr(box), L.ptr(cls), L.ptr(dirp), self.a_loc * self.code * HW, self.a_loc * s.num_class * HW,
                            self.a_loc * s.num_direction_bins * HW, HW, 1, st)      # NCHW conv outputs
        self._launch_nms(st)

    def _launch_tc_tail(self, feats, hilo):
        """BEV (NHWC, halo, fp16 hi/lo) -> tcgen05 RPN layers -> packed heads -> decode/filter -> NMS."""
        L, lib, s = self._L, self.lib, self.spec
        st = L.stream()
        fl = self.final_level
        D, H, W = fl.shape
        self._mark("to_bev")
        if hilo is not None:
            L.check(lib.b2s_to_bev_tc(None, L.ptr(hilo[0]), L.ptr(hilo[1]), hilo[2], L.ptr(fl.coors), L.ptr(fl.n_dev),
                                      fl.cap, self.feat_final_c, self.B, D, H, W, L.ptr(self.tc_bev[0]),
                                      L.ptr(self.tc_bev[1]), L.ptr(self.bg_occ) if self.bg_idx else None, st),
                    "b2s_to_bev_tc")
        else:
            L.check(lib.b2s_to_bev_tc(L.ptr(feats), None, None, 0, L.ptr(fl.coors), L.ptr(fl.n_dev), fl.cap,
                                      self.feat_final_c, self.B, D, H, W, L.ptr(self.tc_bev[0]), L.ptr(self.tc_bev[1]),
                                      L.ptr(self.bg_occ) if self.bg_idx else None, st), "b2s_to_bev_tc")
        self._mark("rpn")
        if self.bg_idx:
            L.check(lib.b2s_rpn_bg_plan(L.ptr(self.bg_occ), self.B, H, W, len(self.bg_idx), L.ptr(self.bg_scratch),
                                        L.ptr(self.bg_flags), L.ptr(self.bg_work), L.ptr(self.bg_list),
                                        L.ptr(self.bg_counts), st), "b2s_rpn_bg_plan")
        marked_tail = False
        n_ops = len(self.tc_plan) - (2 if self.tc_tail_fused else 0)
        for oi, op in enumerate(self.tc_plan[:n_ops]):
            if not op["v2"] and op["kind"] != "block" and not marked_tail:
                self._mark("rpn_1x1")          # the 3x3 stack (k_conv3x3_tc2) is timed apart from the deblock/heads tail
                marked_tail = True
            src, dst = self.tc_bufs[op["src"]], self.tc_bufs[op["dst"]]
            cdst = dst[0].shape[-1]                                   # channels per pixel of the destination map
            esz = dst[0].element_size()
            o_hi = ctypes_ptr(dst[0].data_ptr() + esz * op["dst_coff"])
            o_lo = ctypes_ptr(dst[1].data_ptr() + esz * op["dst_coff"]) if op["planes"] == 2 else None
            work = work_n = bgl = bgn = bhi = blo = None
            if oi in self.bg_idx:
                bl = self.bg_idx.index(oi)
                chi, clo = self.bg_field[bl]
                work, work_n = L.ptr(self.bg_work[bl]), L.ptr(self.bg_counts[bl, 0:])
                if self.bg_fused_fill:       # the conv kernel's epilogue warps store the constant into the background tiles
                    bgl, bgn, bhi, blo = L.ptr(self.bg_list[bl]), L.ptr(self.bg_counts[bl, 1:]), L.ptr(chi), L.ptr(clo)
                else:
                    L.check(lib.b2s_rpn_bg_fill(L.ptr(self.bg_list[bl]), L.ptr(self.bg_counts[bl, 1:]), self.B,
                                                op["Hout"], op["Wout"], op["cout"], L.ptr(chi), L.ptr(clo), o_hi, o_lo,
                                                cdst, st), "b2s_rpn_bg_fill")
            L.check(lib.b2s_conv2d_tc_ex(
                L.ptr(src[0]), L.ptr(src[1]), self.B, op["Hin"], op["Win"], op["cin"], L.ptr(op["w_hi"]),
                L.ptr(op["w_lo"]), op["kh"], op["kw"], op["stride"], op["pad"], op["cout"], op["n_pad"],
                L.ptr(op["scale"]), L.ptr(op["shift"]) if op["shift"] is not None else None,
                1 if op["relu"] else 0, op["Hg"], op["Wg"], o_hi, o_lo, op["Hout"], op["Wout"],
                1 if op["padded"] else 0, cdst, op["out_mul"], op["off_h"], op["off_w"], work, work_n, bgl, bgn, bhi, blo,
                L.ptr(self.status), st), "b2s_conv2d_tc_ex(%s)" % op["kind"])
        if not marked_tail:
            self._mark("rpn_1x1")
        if self.tc_tail_fused:
            d, h = self.tc_plan[-2], self.tc_plan[-1]
            src = self.tc_bufs[d["src"]]
            L.check(lib.b2s_rpn_tail_tc(
                L.ptr(src[0]), L.ptr(src[1]), self.B, d["Hin"], d["Win"], d["cin"], L.ptr(d["w_hi"]), L.ptr(d["w_lo"]),
                d["cout"], L.ptr(d["scale"]), L.ptr(d["shift"]), L.ptr(h["w_hi"]), L.ptr(h["w_lo"]), h["cout"], h["n_pad"],
                L.ptr(h["scale"]), L.ptr(h["shift"]), L.ptr(self.tc_heads), self.tc_head_stride, L.ptr(self.status), st),
                "b2s_rpn_tail_tc")
        S = self.tc_head_stride
        self._mark("decode_filter")
        offs = self.tc_prog["heads"]["offsets"]
        heads = self.tc_heads
        esz = heads.element_size()
        box_p = ctypes_ptr(heads.data_ptr() + offs[0] * esz)
        cls_p = ctypes_ptr(heads.data_ptr() + offs[1] * esz)
        dir_p = ctypes_ptr(heads.data_ptr() + offs[2] * esz) if s.use_direction_classifier else None
        bs = self.fH * self.fW * S
        self._launch_decode(box_p, cls_p, dir_p, bs, bs, bs, 1, S, st)               # packed NHWC head records
        self._launch_nms(st)

    def _launch_decode(self, box_p, cls_p, dir_p, box_bs, cls_bs, dir_bs, ch_stride, pix_stride, st):
        """sigmoid + score threshold + box decode over all anchors -> candidate lists (per frame, or per (class, frame))"""
        L, lib, s = self._L, self.lib, self.spec
        if self.mc:
            if self.use_mask:
                raise RuntimeError("anchors_mask with the per-class NMS branch is ill-defined upstream (masking shifts the "
                                   "anchors_range indices, voxelnet.py:432-439,492-501)")
            L.check(lib.b2s_decode_filter_multiclass(
                box_p, cls_p, dir_p, box_bs, cls_bs, dir_bs, ch_stride, pix_stride, L.ptr(self.anchors), self.B,
                self.a_loc, self.fH, self.fW, self.code, s.num_class, s.num_direction_bins, self.mc_lo, self.mc_hi,
                self.mc_thresh, L.ptr(self.cand_box), L.ptr(self.cand_score), L.ptr(self.cand_label),
                L.ptr(self.cand_dir), L.ptr(self.cand_anchor), L.ptr(self.cand_count), self.cand_cap,
                L.ptr(self.status), st), "b2s_decode_filter_multiclass")
            return
        L.check(lib.b2s_decode_filter_strided(
            box_p, cls_p, dir_p, box_bs, cls_bs, dir_bs, ch_stride, pix_stride, L.ptr(self.anchors),
            L.ptr(self.anchors_mask) if self.use_mask else None, self.B, self.a_loc, self.fH, self.fW,
            self.code, s.num_class, s.num_direction_bins, self.score_thresh, L.ptr(self.cand_box),
            L.ptr(self.cand_score), L.ptr(self.cand_label), L.ptr(self.cand_dir), L.ptr(self.cand_anchor),
            L.ptr(self.cand_count), self.cand_cap, L.ptr(self.status), st), "b2s_decode_filter_strided")

    def _launch_nms(self, st):
        L, lib, s = self._L, self.lib, self.spec
        self._mark("nms")
        if self.mc:
            B, cc, code, n = self.B, self.cand_cap, self.code, self.ncls
            groups = [(0, n * B, self.mc_pre[0], self.mc_post[0], self.mc_iou[0])] if self.mc_uniform else \
                [(c * B, B, self.mc_pre[c], self.mc_post[c], self.mc_iou[c]) for c in range(n)]
            for v0, nv, pre, post, iou in groups:         # class-major virtual frames: one class = B contiguous lists
                L.check(lib.b2s_nms(
                    L.ptr(self.cand_box[v0:]), L.ptr(self.cand_score[v0:]), L.ptr(self.cand_label[v0:]),
                    L.ptr(self.cand_dir[v0:]), L.ptr(self.cand_anchor[v0:]), L.ptr(self.cand_count[v0:]), nv, cc, code,
                    1 if s.use_rotate_nms else 0, pre, post, iou, 1 if s.use_direction_classifier else 0,
                    float(s.direction_offset), float(s.direction_limit_offset), s.num_direction_bins, self.range_host,
                    L.ptr(self.det_mc[v0:]), self.post_max_class * (code + 2), L.ptr(self.cnt_mc[v0:]),
                    L.ptr(self.nms_ws), self.nms_ws_bytes, st), "b2s_nms")
            L.check(lib.b2s_concat_class_detections(L.ptr(self.det_mc), L.ptr(self.cnt_mc), B, n, self.post_max_class,
                                                    code, L.ptr(self.det_record), self.rec_width, L.ptr(self.det_count),
                                                    st), "b2s_concat_class_detections")
            self._mark("end")
            return
        L.check(lib.b2s_nms(
            L.ptr(self.cand_box), L.ptr(self.cand_score), L.ptr(self.cand_label), L.ptr(self.cand_dir),
            L.ptr(self.cand_anchor), L.ptr(self.cand_count), self.B, self.cand_cap, self.code,
            1 if s.use_rotate_nms else 0, self.pre_max, self.post_max, self.iou_thresh,
            1 if s.use_direction_classifier else 0, float(s.direction_offset),
            float(s.direction_limit_offset), s.num_direction_bins, self.range_host, L.ptr(self.det_record),
            self.rec_width, L.ptr(self.det_count), L.ptr(self.nms_ws), self.nms_ws_bytes, st), "b2s_nms")
        self._mark("end")

    # ---------------------------------------------------------------- instrumentation
    _marks = None

    def _mark(self, name):
        if self._marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    def run_timed(self, iters=5, mode="points"):
        """eager (no graph) replay with a CUDA event before every stage; returns {stage: mean ms}."""
        acc = {}
        with torch.no_grad(), torch.cuda.device(self.dev):
            self._launch(mode)     # warm
            for _ in range(iters):
                self._marks = []
                self._launch(mode)
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
                # subm_nbr / subm_ranked | mark, summary_scan, scan_sums, compact_words, nz_scan, scan_sums, emit,
                # (hash_build,) scatter_nbr
                n += 1 if lyr["subm"] else (8 + (1 if lyr.get("want_hash") else 0))
                if lyr["tc"] and "tile_mask" in lyr["rb"]:
                    n += 1                       # b2s_sparse_tile_plan
            n += 1                               # b2s_sparse_conv / b2s_sparse_conv_tc_plan
            if lyr.get("in_split") is not None:
                n += 1                           # b2s_split_f16
        if self.rpn_impl == "tc":
            n += len(self.tc_plan) - (1 if self.tc_tail_fused else 0)   # one tcgen05 kernel per RPN layer (fused tail: 1 for 2)
            if self.bg_idx:
                # k_bg_layer per layer, k_bg_compact (+ k_bg_fill per layer unless the conv epilogue fills)
                n += len(self.bg_idx) + 1 + (0 if self.bg_fused_fill else len(self.bg_idx))
        elif getattr(self, "any_sparse_tc", False):
            n += 1                               # b2s_merge_f16
        if self.mc:
            return n + 1 + 1 + 3 * (1 if self.mc_uniform else self.ncls) + 1      # ... + k_concat_classes
        return n + 1 + 1 + 3                     # to_bev, decode_filter, nms: select_sort + iou_mask + reduce

    def sparse_layer_stats(self):
        """sync; per sparse layer: rows in/out, pairs, algorithmic bytes and flops (SURVEY.md §8d formulas)."""
        out = []
        for lyr in self.layers:
            n_out = int(lyr["out_level"].n_dev[0].item())
            n_in = int(lyr["in_level"].n_dev[0].item())
            pairs = int((lyr["rb"]["nbr"][:n_out] >= 0).sum().item())
            cin, cout, K = lyr["cin"], lyr["cout"], lyr["K"]
            out.append({"index": lyr["index"], "subm": bool(lyr["subm"]), "cin": cin, "cout": cout, "K": K,
                        "n_in": n_in, "n_out": n_out, "pairs": pairs,
                        "useful_mma_frac": pairs / max(1, K * n_out),
                        "bytes": 4 * (n_in * cin + n_out * cout) + 8 * pairs + 4 * K * cin * cout,
                        "flops": 2 * pairs * cin * cout})
        return out

    def rpn_layer_stats(self):
        """per tcgen05 RPN layer: algorithmic fp32 flops (2*pixels*taps*cin*cout; the 3xF16 split issues 3x that on
        the tensor pipe) and bytes moved (fp16 hi/lo planes in and out + weights)."""
        if self.rpn_impl != "tc":
            return []
        out = []
        for i, op in enumerate(self.tc_plan):
            cin, cout, taps = op["cin"], op["cout"], op["taps"]
            px = self.B * op["Hg"] * op["Wg"]
            px_in = self.B * op["Hin"] * op["Win"]
            out_b = 2 * 2 * px * cout if op["planes"] == 2 else 4 * px * cout
            frac = 1.0                        # share of the 16x16 output tiles actually computed (the rest: background)
            if getattr(self, "bg_idx", None) and i in self.bg_idx:
                frac = float(self.bg_counts[self.bg_idx.index(i), 0].item()) / self.bg_tiles
            out.append({"index": i, "kind": op["kind"], "v2": op["v2"], "cin": cin, "cout": cout, "taps": taps,
                        "tiles_computed_frac": frac,
                        "pixels": px, "flops": 2 * px * taps * cin * cout,
                        "bytes": 2 * 2 * px_in * cin + out_b + 2 * 2 * taps * cin * op["n_pad"],
                        "bytes_fp32_algorithmic": 4 * (px_in * cin + px * cout + taps * cin * cout)})
        return out

    # ---------------------------------------------------------------- public API
    def set_anchors(self, anchors):
        """anchors [A, code] or [B', A, code] (the example's ``anchors``; every frame shares one grid)."""
        a = anchors[0] if anchors.dim() == 3 else anchors
        key = (a.data_ptr(), tuple(a.shape), a._version, str(a.device))
        if key == self._anchors_key:
            return
        a = a.reshape(-1, a.shape[-1])
        assert a.shape[0] == self.A and a.shape[1] == self.code, \
            "num_anchors=%d, but num_output=%d. please check size" % (a.shape[0], self.A)
        self.anchors.copy_(a, non_blocking=True)
        self._anchors_key = key

    def set_anchors_mask(self, mask):
        """anchors_mask [B, A] (bool/uint8) or None (voxelnet.py:397-400,432-439)."""
        if mask is None:
            if self.use_mask:
                self.use_mask = False
                self._graphs.clear()
            return
        if self.anchors_mask is None:
            self.anchors_mask = torch.zeros(self.B, self.A, dtype=torch.uint8, device=self.dev)
        if not self.use_mask:
            self.use_mask = True
            self._graphs.clear()               # the launch arguments change (NULL -> buffer)
        self.anchors_mask.copy_(mask.reshape(self.B, self.A).to(torch.uint8), non_blocking=True)

    def load_points(self, frames):
        """frames: list of B CUDA (or pinned/CPU) float32 [P_i, F] tensors -> static input buffers."""
        assert len(frames) == self.B
        sizes = [int(f.shape[0]) for f in frames]
        total = sum(sizes)
        assert total <= self.P_cap, "more points (%d) than the engine's capacity (%d)" % (total, self.P_cap)
        # frames that lie back to back in memory (views of one pinned slab, or of one device tensor) go over in ONE copy:
        # a serving loop that reads clouds into a pinned ring buffer pays one cudaMemcpyAsync per batch instead of B.
        # The run structure only depends on where the frames lie: a serving loop cycles through a few slabs, so it is
        # remembered per (pointers, sizes, strides) -- finding it anew is ~0.25 ms of Python per call for 32 frames.
        key = (tuple(f.data_ptr() for f in frames), tuple(sizes), tuple(f.stride(0) for f in frames),
               frames[0].dtype, frames[0].device) if frames else None
        cache = getattr(self, "_run_cache", None)
        if cache is None:
            cache = self._run_cache = {}
        runs = cache.get(key)
        if runs is not None:
            # (a remembered multi-frame run must still lie inside ONE storage: same addresses, different allocations)
            for i, off, run_rows, whole in runs:
                if whole and frames[i].untyped_storage().nbytes() < 4 * (frames[i].storage_offset() + run_rows * self.F):
                    runs = None
                    del cache[key]
                    break
        if runs is not None:
            for i, off, run_rows, whole in runs:
                if whole:
                    src = torch.as_strided(frames[i], (run_rows, self.F), (self.F, 1))
                    self.points[off:off + run_rows].copy_(src, non_blocking=True)
                else:
                    self.points[off:off + run_rows].copy_(frames[i], non_blocking=True)
            return self._load_offsets(sizes, total)
        runs = []
        off, i = 0, 0
        while i < len(frames):
            f = frames[i]
            run_rows, j = sizes[i], i + 1
            if f.is_contiguous() and f.dtype == torch.float32:
                end = f.data_ptr() + 4 * f.numel()
                # (same STORAGE, not just adjacent addresses: two separate allocations can end up back to back in the
                # caching allocator, and a strided view must stay inside its own storage)
                base = f.untyped_storage().data_ptr()
                while j < len(frames) and frames[j].dtype == torch.float32 and frames[j].device == f.device and \
                        frames[j].is_contiguous() and frames[j].data_ptr() == end and \
                        frames[j].untyped_storage().data_ptr() == base:
                    end += 4 * frames[j].numel()
                    run_rows += sizes[j]
                    j += 1
            if run_rows:
                if j - i > 1:
                    src = torch.as_strided(f, (run_rows, self.F), (self.F, 1))      # the whole run through frame i's storage
                    self.points[off:off + run_rows].copy_(src, non_blocking=True)
                else:
                    self.points[off:off + run_rows].copy_(f, non_blocking=True)
                runs.append((i, off, run_rows, j - i > 1))
            off += run_rows
            i = j
        if len(cache) < 16 and all(f.dtype == torch.float32 for f in frames):
            cache[key] = runs
        return self._load_offsets(sizes, total)

    def _load_offsets(self, sizes, total):
        # frame offsets through a pinned staging buffer that is allocated once (a fresh pin_memory() per call is a
        # cudaHostAlloc on the serving path); two slots, so the copy of the previous call is never overwritten in flight
        if getattr(self, "_offs_host", None) is None:
            self._offs_host = [torch.zeros(self.B + 1, dtype=torch.int32).pin_memory() for _ in range(2)]
            self._offs_evt = [torch.cuda.Event(), torch.cuda.Event()]
            self._offs_slot = 0
        slot = self._offs_slot
        self._offs_slot ^= 1
        self._offs_evt[slot].synchronize()                 # (no-op unless two calls ago is still copying)
        self._offs_host[slot].copy_(torch.from_numpy(np.cumsum([0] + sizes).astype(np.int32)))
        self.offsets.copy_(self._offs_host[slot], non_blocking=True)
        self._offs_evt[slot].record()
        return total

    def load_points_cropped(self, frames, planes):
        """frames: list of B CUDA float32 [P_i, F] raw clouds; planes: one [n,4] float64 array per frame (or one for
        all): the points strictly inside the convex region (``b2second.inputs.frustum_planes``: the KITTI camera-FOV
        crop, box_np_ops.py:682-693) are compacted on the device straight into the static point buffer, in input
        order, and ``offsets`` is filled on the device -- no host sync, no host copy of the cropped cloud."""
        from . import inputs
        assert len(frames) == self.B
        if not isinstance(planes, (list, tuple)):
            planes = [planes] * self.B
        if getattr(self, "_crop", None) is None:
            self._crop = inputs.ConvexCrop(self.P_cap, self.dev)
            self.in_status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        with torch.cuda.device(self.dev):
            self.offsets.zero_()
            self._crop_keep = [self._crop.crop_into(f if f.is_cuda else f.to(self.dev, non_blocking=True), pl,
                                                    self.points, self.offsets, b, self.in_status)
                               for b, (f, pl) in enumerate(zip(frames, planes))]

    def load_sweeps(self, frames):
        """frames: list of B dicts {"sweeps": [CUDA float32 [P_i, F>=3], key frame first], "rotations": [..[3,3]..],
        "translations": [..[3]..], "time_lags": [..]} (the ``sweep2lidar_*`` fields of a NuScenes info record).  The
        merged [x, y, z, dt] cloud of every frame (nuscenes_dataset.py:166-185) is written straight into the static
        point buffer."""
        from . import inputs
        assert len(frames) == self.B and self.F == 4, "sweep merging produces [x, y, z, dt] rows"
        sizes = [sum(int(s.shape[0]) for s in fr["sweeps"]) for fr in frames]
        assert sum(sizes) <= self.P_cap, "more points (%d) than the engine's capacity (%d)" % (sum(sizes), self.P_cap)
        off = 0
        with torch.cuda.device(self.dev):
            for fr, n in zip(frames, sizes):
                sw = [s if s.is_cuda else s.to(self.dev, non_blocking=True) for s in fr["sweeps"]]
                inputs.merge_sweeps(sw, fr["rotations"], fr["translations"], fr["time_lags"], out=self.points[off:off + n])
                off += n
            offs = torch.tensor(np.cumsum([0] + sizes), dtype=torch.int32)
            self.offsets.copy_(offs.pin_memory(), non_blocking=True)
        return off

    def load_voxels(self, voxels, num_points, coordinates):
        """the reference's example tensors (voxels [N,T,F], num_points [N], coordinates [N,4] (b,z,y,x)), on any
        device, -> static buffers.  N is known on the host (a tensor shape): no sync."""
        n = int(voxels.shape[0])
        cap = self.level0.cap
        assert n <= cap, "more voxels (%d) than the engine's capacity (%d)" % (n, cap)
        assert tuple(voxels.shape[1:]) == (self.T, self.F), "voxels must be [N, %d, %d]" % (self.T, self.F)
        if self.in_voxels is None:
            self.in_voxels = torch.zeros(cap, self.T, self.F, dtype=torch.float32, device=self.dev)
            # PFN reads points through per-voxel slot indices: voxel v, slot t -> row v*T + t of in_voxels
            self.in_slots = torch.arange(cap * self.T, dtype=torch.int32, device=self.dev).view(cap, self.T)
            self._nvox_host = torch.zeros(1 + self.B, dtype=torch.int32).pin_memory()
        self.in_voxels[:n].copy_(voxels, non_blocking=True)
        self.vox_num[:n].copy_(num_points.to(torch.int32) if num_points.dtype != torch.int32 else num_points,
                               non_blocking=True)
        self.vox_coors[:n].copy_(coordinates.to(torch.int32) if coordinates.dtype != torch.int32 else coordinates,
                                 non_blocking=True)
        self._nvox_host[0] = n
        self.num_voxels.copy_(self._nvox_host, non_blocking=True)
        return n

    def run(self, mode="points"):
        """launch the pipeline on the data currently in the static buffers (async)."""
        # the grid sizes of the per-point kernels are fixed at capacity so one graph serves any frame mix
        with torch.cuda.device(self.dev):
            if not self.use_graph:
                with torch.no_grad():
                    self._launch(mode)
                return self.det, self.det_count
            g = self._graphs.get(mode)
            if g is None:
                with torch.no_grad():
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        for _ in range(3):          # warm-up: cuDNN autotune, lazy module init, func attributes
                            self._launch(mode)
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._launch(mode)
                    self._graphs[mode] = g
            g.replay()
        return self.det, self.det_count

    def infer(self, frames):
        self.load_points(frames)
        return self.run("points")

    def infer_voxels(self, voxels, num_points, coordinates):
        self.load_voxels(voxels, num_points, coordinates)
        return self.run("voxels")

    def check_status(self):
        """sync + raise on data-dependent overflow (call when results are read back)."""
        st = int(self.status.item())
        if getattr(self, "in_status", None) is not None:
            if int(self.in_status.item()):
                raise RuntimeError("b2second engine: cropped clouds exceed the point buffer (max_points)")
        if st & ~1:   # bit 1 (voxel overflow) is the reference's own drop-extra-voxels behaviour
            raise RuntimeError("b2second engine: " + self._L.status_message(st))
        return st

    def detections(self, metadata=None, output="host"):
        """sync and convert to the reference's list-of-dicts (voxelnet.py:616-645).

        output="host": ONE pinned D2H copy of the detection records (counts ride in the records' last element) +
        one stream sync, sliced on the host into CPU tensors.  output="device": the tensors stay on the GPU like the
        reference's (one snapshot clone, per-frame views)."""
        code, B = self.code, self.B
        meta = metadata if metadata is not None and len(metadata) > 0 else [None] * B
        if getattr(self, "_rec_host", None) is None:
            self._rec_host = torch.empty(B, self.rec_width, dtype=torch.float32).pin_memory()
            self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        with torch.cuda.device(self.dev):
            self._rec_host.copy_(self.det_record, non_blocking=True)
            self._status_host.copy_(self.status, non_blocking=True)
            if output == "device":
                snap = self.det.clone()
                labels = snap[..., code + 1].long()
            torch.cuda.current_stream().synchronize()          # the one sync of the step
        st = int(self._status_host[0])
        if st & ~1:   # bit 1 (voxel overflow) is the reference's own drop-extra-voxels behaviour
            raise RuntimeError("b2second engine: " + self._L.status_message(st))
        cnt = self._rec_host[:, -1].round().to(torch.int64).tolist()
        out = []
        if output == "device":
            for b in range(B):
                n = cnt[b]
                out.append({"box3d_lidar": snap[b, :n, :code], "scores": snap[b, :n, code],
                            "label_preds": labels[b, :n], "metadata": meta[b]})
            return out
        # per-frame views through numpy: 3 B tensor views cost ~0.1 ms this way and ~0.8 ms as torch indexing ops (the
        # Python between two steps is GPU idle time on the serving path)
        host = self._rec_host.numpy()[:, :-1].reshape(B, self.post_max, code + 2).copy()
        labels = host[..., code + 1].astype(np.int64)
        fn = torch.from_numpy
        for b in range(B):
            n = cnt[b]
            out.append({"box3d_lidar": fn(host[b, :n, :code]), "scores": fn(host[b, :n, code]),
                        "label_preds": fn(labels[b, :n]), "metadata": meta[b]})
        return out
