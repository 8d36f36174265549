"""CPU ORACLE of ``spconv.utils`` (spconv 1.x API as used by second.pytorch).

TEST INFRASTRUCTURE ONLY -- see oracle/README.md.  PARITY UNPINNED: spconv is not in
/root/reference; semantics follow SURVEY.md App. A and the reference call sites:

* ``VoxelGeneratorV2``           second/builder/voxel_builder.py:23-32,
                                 second/data/preprocess.py:303-315, second/pytorch/train.py:60
* ``non_max_suppression``        second/core/non_max_suppression/nms_gpu.py:10-19
* ``non_max_suppression_cpu``    second/core/non_max_suppression/nms_cpu.py:14-17
* ``rotate_non_max_suppression_cpu``  nms_cpu.py:20-31
* ``rbbox_iou`` / ``rbbox_intersection``  second/core/box_np_ops.py:10-34

The arithmetic lives in oracle/c/oracle.c (plain C, built by oracle/Makefile).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.abspath(os.path.join(_HERE, "..", "..", ".."))
_LIB_PATH = os.path.join(_ORACLE_DIR, "_build", "liboracle.so")
_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    lib = ctypes.CDLL(_LIB_PATH)
    f32p = ctypes.POINTER(ctypes.c_float)
    i32p = ctypes.POINTER(ctypes.c_int)
    f64p = ctypes.POINTER(ctypes.c_double)
    lib.orc_points_to_voxel.restype = ctypes.c_int
    lib.orc_points_to_voxel.argtypes = [f32p, ctypes.c_int, ctypes.c_int, f32p, f32p, i32p,
                                        ctypes.c_int, ctypes.c_int, f32p, i32p, i32p, i32p]
    lib.orc_nms_aligned_sorted.restype = ctypes.c_int
    lib.orc_nms_aligned_sorted.argtypes = [f32p, ctypes.c_int, ctypes.c_float, i32p]
    lib.orc_nms_cpu.restype = ctypes.c_int
    lib.orc_nms_cpu.argtypes = [f32p, i32p, ctypes.c_int, ctypes.c_float, ctypes.c_float, i32p]
    lib.orc_rotate_nms.restype = ctypes.c_int
    lib.orc_rotate_nms.argtypes = [f32p, i32p, f32p, ctypes.c_int, ctypes.c_float, i32p, f64p]
    lib.orc_rotate_nms_f32.restype = ctypes.c_int
    lib.orc_rotate_nms_f32.argtypes = [f32p, i32p, f32p, ctypes.c_int, ctypes.c_float, i32p, f64p]
    lib.orc_rbbox_iou.restype = None
    lib.orc_rbbox_iou.argtypes = [f32p, f32p, f32p, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                  ctypes.c_int, f32p]
    _lib = lib
    return lib


def _f32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


class VoxelGeneratorV2:
    """First-come voxeliser (spconv ``points_to_voxel_3d_np`` semantics)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000,
                 full_mean=False, block_filtering=False, block_factor=8, block_size=3,
                 height_threshold=0.1, height_high_threshold=2.0):
        assert full_mean is False, "full_mean is asserted False upstream"
        assert not block_filtering, "block_filtering is off in every BASELINE config (out of scope)"
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        grid_size = np.round(grid_size).astype(np.int64)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = int(max_num_points)
        self._max_voxels = int(max_voxels)
        self._grid_size = grid_size
        self._scratch = None  # persistent coor_to_voxelidx grid (D*H*W int32, -1)

    def _points_to_voxel(self, points, max_voxels):
        lib = _load()
        points = np.ascontiguousarray(points, dtype=np.float32)
        P, F = points.shape
        T = self._max_num_points
        grid = self._grid_size.astype(np.int32)
        if self._scratch is None:
            self._scratch = np.full(int(np.prod(grid)), -1, dtype=np.int32)
        voxels = np.zeros((max_voxels, T, F), dtype=np.float32)
        coors = np.zeros((max_voxels, 3), dtype=np.int32)
        num = np.zeros((max_voxels,), dtype=np.int32)
        lo = np.ascontiguousarray(self._point_cloud_range[:3])
        vs = np.ascontiguousarray(self._voxel_size)
        n = lib.orc_points_to_voxel(_f32(points), P, F, _f32(lo), _f32(vs), _i32(grid), T,
                                    int(max_voxels), _f32(voxels), _i32(coors), _i32(num),
                                    _i32(self._scratch))
        return voxels, coors, num, n

    def generate(self, points, max_voxels=None):
        mv = int(max_voxels or self._max_voxels)
        voxels, coors, num, n = self._points_to_voxel(points, mv)
        return {
            "voxels": voxels[:n],
            "coordinates": coors[:n],
            "num_points_per_voxel": num[:n],
            "voxel_point_mask": (np.arange(self._max_num_points)[None, :] < num[:n, None])[..., None]
            .astype(np.float32),
            "voxel_num": n,
        }

    def generate_multi_gpu(self, points, max_voxels=None):
        mv = int(max_voxels or self._max_voxels)
        voxels, coors, num, n = self._points_to_voxel(points, mv)
        return {
            "voxels": voxels,
            "coordinates": coors,
            "num_points_per_voxel": num,
            "voxel_point_mask": (np.arange(self._max_num_points)[None, :] < num[:, None])[..., None]
            .astype(np.float32),
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
    """sorted_dets [N,5] f32 (descending score), keep_out [N] i32 (filled) -> num_out."""
    lib = _load()
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return 0
    keep = np.zeros(n, dtype=np.int32)
    k = lib.orc_nms_aligned_sorted(_f32(boxes), n, float(nms_overlap_thresh), _i32(keep))
    keep_out[:k] = keep[:k]
    return k


def non_max_suppression_cpu(boxes, order, thresh, eps=0.0):
    lib = _load()
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    order = np.ascontiguousarray(order, dtype=np.int32)
    n = boxes.shape[0]
    if n == 0:
        return []
    keep = np.zeros(n, dtype=np.int32)
    k = lib.orc_nms_cpu(_f32(boxes), _i32(order), n, float(thresh), float(eps), _i32(keep))
    return keep[:k].tolist()


def rotate_non_max_suppression_cpu(box_corners, order, standup_iou, thresh, return_iou=False, precision="f64"):
    """precision="f64": polygon clip in double (default).  "f32": the same loop with the clip in float -- the
    precision the float corners arrive in and the CUDA kernel computes in; tests/test_oracle_nms.py counts how many
    keep lists depend on that choice."""
    lib = _load()
    box_corners = np.ascontiguousarray(box_corners, dtype=np.float32)
    order = np.ascontiguousarray(order, dtype=np.int32)
    standup_iou = np.ascontiguousarray(standup_iou, dtype=np.float32)
    n = box_corners.shape[0]
    if n == 0:
        return ([], np.zeros((0, 0))) if return_iou else []
    keep = np.zeros(n, dtype=np.int32)
    iou = np.full((n, n), -1.0, dtype=np.float64) if return_iou else None
    fn = lib.orc_rotate_nms if precision == "f64" else lib.orc_rotate_nms_f32
    k = fn(_f32(box_corners), _i32(order), _f32(standup_iou), n, float(thresh),
                           _i32(keep),
                           iou.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if return_iou else None)
    if return_iou:
        return keep[:k].tolist(), iou
    return keep[:k].tolist()


def _rbbox(box_corners, qbox_corners, standup_iou, standup_thresh, mode):
    lib = _load()
    box_corners = np.ascontiguousarray(box_corners, dtype=np.float32)
    qbox_corners = np.ascontiguousarray(qbox_corners, dtype=np.float32)
    standup_iou = np.ascontiguousarray(standup_iou, dtype=np.float32)
    n, k = box_corners.shape[0], qbox_corners.shape[0]
    out = np.zeros((n, k), dtype=np.float32)
    if n and k:
        lib.orc_rbbox_iou(_f32(box_corners), _f32(qbox_corners), _f32(standup_iou), n, k,
                          float(standup_thresh), mode, _f32(out))
    return out


def rbbox_iou(box_corners, qbox_corners, standup_iou, standup_thresh):
    return _rbbox(box_corners, qbox_corners, standup_iou, standup_thresh, 0)


def rbbox_intersection(box_corners, qbox_corners, standup_iou, standup_thresh):
    return _rbbox(box_corners, qbox_corners, standup_iou, standup_thresh, 1)


def quad_intersection_area(a, b):
    """Helper for tests: area of intersection of two 4-corner convex polygons (f64)."""
    lib = _load()
    lib.orc_quad_intersection.restype = ctypes.c_double
    lib.orc_quad_intersection.argtypes = [ctypes.POINTER(ctypes.c_float)] * 2 + \
        [ctypes.POINTER(ctypes.c_double)] * 2
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return lib.orc_quad_intersection(_f32(a), _f32(b), None, None)
