"""``spconv.utils`` (numpy-facing helpers second.pytorch imports) on libb2second.so.

  VoxelGeneratorV2                  second/builder/voxel_builder.py:23-32, second/data/preprocess.py:303-315,
                                    second/pytorch/train.py:60 (attributes voxel_size / point_cloud_range / grid_size)
  non_max_suppression               second/core/non_max_suppression/nms_gpu.py:10-19
  non_max_suppression_cpu           second/core/non_max_suppression/nms_cpu.py:14-17
  rotate_non_max_suppression_cpu    nms_cpu.py:20-31
  rbbox_iou / rbbox_intersection    second/core/box_np_ops.py:10-34 (target assignment / eval)
  rotate_iou_eval                   device-resident counterpart of nms_gpu.py:569-607 ``rotate_iou_gpu_eval``

The numpy signatures are kept (host arrays in, host arrays out); the work runs on the GPU.  The
device-resident fast path (points already on the GPU, no host round trip) is ``generate_device``.
"""
import ctypes

import numpy as np
import torch

from .. import _lib


def _as_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class VoxelGeneratorV2:
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, full_mean=False,
                 block_filtering=False, block_factor=8, block_size=3, height_threshold=0.1,
                 height_high_threshold=2.0):
        assert full_mean is False
        if block_filtering:
            raise NotImplementedError("block_filtering is off in every BASELINE config (SURVEY.md App. A)")
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = np.round((point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size).astype(np.int64)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = int(max_num_points)
        self._max_voxels = int(max_voxels)
        self._grid_size = grid_size
        self._device = None

    # -- device-resident API (no host round trip) -------------------------------------------------
    def generate_device(self, points, max_voxels=None, frame_offsets=None, batch=1, vfe_mode=0,
                        vfe_num_features=4, want_voxels=True, hash_key_depth=0):
        """points: CUDA float32 [P,F] (``batch`` frames back to back, ``frame_offsets`` int32 [batch+1]).

        Returns a dict of CUDA tensors with capacity rows ``batch*max_voxels`` (NOT sliced):
        ``coordinates`` [cap,4] (b,z,y,x), ``num_points_per_voxel`` [cap], ``point_slots`` [cap,T],
        ``voxels`` [cap,T,F] (if want_voxels), ``vfe`` [cap,C] (if vfe_mode), ``num_voxels`` int32 [1+batch]
        (device), ``hash`` (keys, vals, cap), ``status`` int32 [1].  Nothing synchronises."""
        _lib.require_cuda(points, "points")
        lib = _lib.load()
        pts = points.contiguous().float()
        P, F = pts.shape
        dev = pts.device
        mv = int(max_voxels or self._max_voxels)
        T = self._max_num_points
        cap = batch * mv
        hcap = lib.b2s_voxelize_hash_capacity(P)
        ws_bytes = lib.b2s_voxelize_workspace_bytes(P, batch, mv, T)
        out = {
            "coordinates": torch.empty(cap, 4, dtype=torch.int32, device=dev),
            "num_points_per_voxel": torch.empty(cap, dtype=torch.int32, device=dev),
            "point_slots": torch.empty(cap, T, dtype=torch.int32, device=dev),
            "voxels": torch.empty(cap, T, F, dtype=torch.float32, device=dev) if want_voxels else None,
            "vfe": None,
            "num_voxels": torch.zeros(1 + batch, dtype=torch.int32, device=dev),
            "status": torch.zeros(1, dtype=torch.int32, device=dev),
        }
        if vfe_mode:
            c = vfe_num_features if vfe_mode == 1 else vfe_num_features - 1
            out["vfe"] = torch.empty(cap, c, dtype=torch.float32, device=dev)
        keys = torch.empty(hcap, dtype=torch.int64, device=dev)
        vals = torch.empty(hcap, dtype=torch.int32, device=dev)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        _lib.check(lib.b2s_voxelize(
            _lib.ptr(pts), _lib.ptr(frame_offsets), P, F, batch, _lib.f3(self._point_cloud_range[:3]),
            _lib.f3(self._voxel_size), _lib.i3(self._grid_size), T, mv, _lib.ptr(out["coordinates"]),
            _lib.ptr(out["num_points_per_voxel"]), _lib.ptr(out["point_slots"]), _lib.ptr(out["voxels"]),
            int(vfe_mode), int(vfe_num_features), _lib.ptr(out["vfe"]), _lib.ptr(out["num_voxels"]),
            _lib.ptr(keys), _lib.ptr(vals), hcap, int(hash_key_depth), _lib.ptr(ws), ws_bytes,
            _lib.ptr(out["status"]), _lib.stream()),
            "b2s_voxelize")
        out["hash"] = (keys, vals, hcap)
        out["_keepalive"] = (pts, ws)
        return out

    # -- upstream numpy API ------------------------------------------------------------------------
    def _run_numpy(self, points, max_voxels):
        if not torch.cuda.is_available():
            raise RuntimeError("spconv.utils.VoxelGeneratorV2 (b2second) needs a CUDA device: no CPU fallback")
        pts = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).cuda(non_blocking=False)
        res = self.generate_device(pts, max_voxels)
        n = int(res["num_voxels"][0].item())
        return res, n

    def generate(self, points, max_voxels=None):
        res, n = self._run_numpy(points, max_voxels)
        num = res["num_points_per_voxel"][:n].cpu().numpy()
        T = self._max_num_points
        return {
            "voxels": res["voxels"][:n].cpu().numpy(),
            "coordinates": res["coordinates"][:n, 1:].contiguous().cpu().numpy(),
            "num_points_per_voxel": num,
            "voxel_point_mask": (np.arange(T)[None, :] < num[:, None])[..., None].astype(np.float32),
            "voxel_num": n,
        }

    def generate_multi_gpu(self, points, max_voxels=None):
        res, n = self._run_numpy(points, max_voxels)
        voxels = res["voxels"].cpu().numpy()
        coors = res["coordinates"][:, 1:].contiguous().cpu().numpy()
        num = res["num_points_per_voxel"].cpu().numpy()
        voxels[n:] = 0
        coors[n:] = 0
        num[n:] = 0
        T = self._max_num_points
        return {
            "voxels": voxels, "coordinates": coors, "num_points_per_voxel": num,
            "voxel_point_mask": (np.arange(T)[None, :] < num[:, None])[..., None].astype(np.float32),
            "voxel_num": n,
        }

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_num_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size


def non_max_suppression(boxes, keep_out, nms_overlap_thresh, device_id=0):
    """sorted_dets [N,5] f32 (descending score), keep_out [N] i32 (filled) -> num_out ("+1" IoU, '>')."""
    lib = _lib.load()
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return 0
    keep = np.zeros(n, dtype=np.int32)
    k = _lib.check(lib.b2s_nms_aligned_host(_as_ptr(boxes), n, float(nms_overlap_thresh), 1.0, 0, _as_ptr(keep),
                                            int(device_id)), "b2s_nms_aligned_host")
    keep_out[:k] = keep[:k]
    return k


def non_max_suppression_cpu(boxes, order, thresh, eps=0.0):
    """dets [N,5], order [N] i32 -> list of kept box ids (eps IoU, '>=')."""
    lib = _lib.load()
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    order = np.ascontiguousarray(order, dtype=np.int32)
    n = order.shape[0]
    if n == 0:
        return []
    sorted_dets = np.ascontiguousarray(boxes[order])
    keep = np.zeros(n, dtype=np.int32)
    # device -1 = the caller's current CUDA device (upstream's *_cpu functions take no device argument)
    k = _lib.check(lib.b2s_nms_aligned_host(_as_ptr(sorted_dets), n, float(thresh), float(eps), 1, _as_ptr(keep), -1),
                   "b2s_nms_aligned_host")
    return order[keep[:k]].tolist()


def rotate_non_max_suppression_cpu(box_corners, order, standup_iou, thresh):
    lib = _lib.load()
    box_corners = np.ascontiguousarray(box_corners, dtype=np.float32)
    order = np.ascontiguousarray(order, dtype=np.int32)
    n = order.shape[0]
    if n == 0:
        return []
    keep = np.zeros(n, dtype=np.int32)
    sorted_corners = np.ascontiguousarray(box_corners[order])      # row i = i-th box in score order
    ident = np.arange(n, dtype=np.int32)
    siou = np.ascontiguousarray(standup_iou, dtype=np.float32)     # gate is recomputed on the device
    k = _lib.check(lib.b2s_nms_rotated_host(_as_ptr(sorted_corners), _as_ptr(ident), _as_ptr(siou), n, float(thresh),
                                            _as_ptr(keep), -1), "b2s_nms_rotated_host")
    return order[keep[:k]].tolist()


def _rbbox(box_corners, qbox_corners, standup_iou, standup_thresh, criterion):
    lib = _lib.load()
    box_corners = np.ascontiguousarray(box_corners, dtype=np.float32).reshape(-1, 4, 2)
    qbox_corners = np.ascontiguousarray(qbox_corners, dtype=np.float32).reshape(-1, 4, 2)
    n, k = box_corners.shape[0], qbox_corners.shape[0]
    standup_iou = np.ascontiguousarray(standup_iou, dtype=np.float32).reshape(n, k)
    out = np.zeros((n, k), dtype=np.float32)
    if n and k:
        _lib.check(lib.b2s_rbbox_overlap_host(_as_ptr(box_corners), _as_ptr(qbox_corners), _as_ptr(standup_iou), n, k,
                                              float(standup_thresh), int(criterion), _as_ptr(out), -1),
                   "b2s_rbbox_overlap_host")
    return out


def rbbox_iou(box_corners, qbox_corners, standup_iou, standup_thresh):
    """rotated IoU of every (box, query) pair whose stand-up IoU exceeds ``standup_thresh`` (0 elsewhere):
    second/core/box_np_ops.py:10-20 ``riou_cc`` -> region_similarity.py:70."""
    return _rbbox(box_corners, qbox_corners, standup_iou, standup_thresh, -1)


def rbbox_intersection(box_corners, qbox_corners, standup_iou, standup_thresh):
    """rotated intersection AREA (box_np_ops.py:23-34 ``rinter_cc``; second/utils/eval.py:174-175 uses it
    interchangeably with ``rotate_iou_gpu_eval(..., criterion=2)`` = area_inter, nms_gpu.py:553-566)."""
    return _rbbox(box_corners, qbox_corners, standup_iou, standup_thresh, 2)


def rotate_iou_eval(boxes, query_boxes, criterion=-1):
    """device-resident ``rotate_iou_gpu_eval`` (nms_gpu.py:569-607): CUDA tensors [N,5] / [K,5] (x, y, w, l, r)
    -> CUDA tensor [N,K]; criterion -1 IoU, 0 inter/area(box), 1 inter/area(query), 2 intersection area."""
    _lib.require_cuda(boxes, "boxes")
    _lib.require_cuda(query_boxes, "query_boxes")
    lib = _lib.load()
    b = boxes.contiguous().float()
    q = query_boxes.contiguous().float()
    n, k = b.shape[0], q.shape[0]
    out = torch.zeros(n, k, dtype=torch.float32, device=b.device)
    if n and k:
        ws = torch.empty((n + k) * 8, dtype=torch.float32, device=b.device)
        _lib.check(lib.b2s_rotate_iou_eval(_lib.ptr(b), n, _lib.ptr(q), k, int(criterion), _lib.ptr(out), _lib.ptr(ws),
                                           _lib.stream()), "b2s_rotate_iou_eval")
    return out
