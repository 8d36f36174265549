"""ctypes binding of libb2second.so (include/b2second.h).

The product path has NO fallback: if the CUDA library is missing or a call fails, this raises.
Memory is owned by torch tensors; raw device pointers and the current CUDA stream are passed down.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2S_LIB", os.path.join(_HERE, "..", "csrc", "libb2second.so"))

_lib = None

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
_I3 = ctypes.c_int * 3
_F3 = ctypes.c_float * 3
_F6 = ctypes.c_float * 6

# name -> (restype, argtypes); mirrors include/b2second.h exactly (checked by tests/test_abi.py)
SIGNATURES = {
    "b2s_version": (c_int, []),
    "b2s_last_error": (ctypes.c_char_p, []),
    "b2s_transform_sweep": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "b2s_crop_workspace_bytes": (c_size_t, [c_int]),
    "b2s_crop_convex": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                                c_void_p, c_void_p]),
    "b2s_voxelize_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "b2s_voxelize_hash_capacity": (c_int, [c_int]),
    "b2s_voxelize": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "b2s_hash_build": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b2s_rulebook_subm": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_void_p]),
    "b2s_rulebook_conv_workspace_bytes": (c_size_t, [c_int, c_void_p]),
    "b2s_rulebook_conv": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                  c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "b2s_rulebook_subm_ranked": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                         c_void_p, c_void_p, c_void_p]),
    "b2s_rulebook_pairs": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b2s_sparse_conv": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                c_int, c_void_p, c_int, c_void_p]),
    "b2s_to_bev": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                           c_void_p]),
    "b2s_pfn": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                        c_void_p, c_int, c_float, c_float, c_float, c_float, c_void_p, c_void_p]),
    "b2s_decode_filter": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_void_p, c_void_p]),
    "b2s_sparse_conv_tc_supported": (c_int, [c_int, c_int]),
    "b2s_sparse_conv_tc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                   c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "b2s_sparse_tile_plan": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p]),
    "b2s_sparse_conv_tc_plan": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                        c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                        c_int, c_int, c_void_p, c_void_p]),
    "b2s_split_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "b2s_merge_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b2s_to_bev_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                              c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2s_rpn_bg_plan": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "b2s_rpn_bg_fill": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_void_p]),
    "b2s_conv2d_tc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "b2s_conv2d_tc_ex": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "b2s_rpn_tail_tc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b2s_decode_filter_strided": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_longlong,
                                          ctypes.c_longlong, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                          c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b2s_decode_filter_multiclass": (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, ctypes.c_longlong,
                                             ctypes.c_longlong, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                             c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "b2s_concat_class_detections": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                            c_void_p]),
    "b2s_nms_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "b2s_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                        c_int, c_float, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                        c_size_t, c_void_p]),
    "b2s_vfe_mean": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "b2s_rotate_iou_eval": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "b2s_rbbox_overlap_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_int]),
    "b2s_nms_aligned_host": (c_int, [c_void_p, c_int, c_float, c_float, c_int, c_void_p, c_int]),
    "b2s_nms_rotated_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_int]),
}


def load():
    """dlopen libb2second.so and declare every entry point.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.abspath(LIB_PATH)
    if not os.path.exists(path):
        raise ImportError(
            "libb2second.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C second.pytorch_b200/csrc`).  There is no CPU fallback on the product path." % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.b2s_version() < 200:
        raise ImportError("libb2second.so is too old")
    _lib = lib
    return lib


def check(rc, what):
    if rc < 0:
        raise RuntimeError("%s failed (%d): %s" % (what, rc, load().b2s_last_error().decode()))
    return rc


def ptr(t):
    """device (or host) pointer of a tensor / None."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def i3(v):
    return _I3(*[int(x) for x in v])


def f3(v):
    return _F3(*[float(x) for x in v])


def f6(v):
    return _F6(*[float(x) for x in v])


def require_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("%s must be a CUDA tensor: the b2second spconv backend has no CPU path" % name)


STATUS_BITS = {1: "voxel overflow (more voxels than max_voxels; extra voxels dropped as upstream does)",
               2: "row overflow (strided conv produced more rows than the buffer capacity)",
               4: "hash table full", 8: "candidate overflow (more score survivors than cand_cap)",
               16: "an activation left the fp16 range (|x| > 65504) on the tensor-core path"}


def status_message(word):
    return "; ".join(msg for bit, msg in STATUS_BITS.items() if word & bit)
