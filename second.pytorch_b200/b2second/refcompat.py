"""Import the UNMODIFIED reference (``/root/reference``) on a modern toolchain.

Only used in the build container (golden generation and drop-in tests); the GPU box has no
``/root/reference`` and nothing on the product path imports this module.

The shims are external and never touch the reference tree:
  1. sys.path: a ``spconv`` package (ours or the oracle's) + the reference root
  2. ``collections.Iterable`` alias (torchplus/train/optim.py:1)
  3. PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION=python (protoc-3 era *_pb2.py)
  4. ``np.meshgrid`` returning a list (second/core/box_np_ops.py:586-592,624-630 item-assign)
  5. config loading: only the ``model`` / ``eval_input_reader`` blocks (map<> fields elsewhere
     break protobuf>=4 pure-python)
  6. ``build_network`` restated from second/pytorch/train.py:58-68 (train.py itself needs fire etc.)
  7. NUMBA_ENABLE_CUDASIM=1 on GPU-less hosts (nms_gpu.py:478,549,564 eager @cuda.jit)
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("B2S_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "second"))


_installed = False


def install(spconv_path):
    """Put ``spconv_path`` (dir containing a ``spconv`` package) + the reference on sys.path."""
    global _installed
    import collections
    import collections.abc

    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "spconv" in sys.modules:
        mod = sys.modules["spconv"]
        have = os.path.dirname(os.path.dirname(os.path.abspath(mod.__file__)))
        if os.path.abspath(spconv_path) != have:
            raise RuntimeError("a different spconv (%s) is already imported" % have)
    os.environ.setdefault("PROTOCOL_BUFFERS_PYTHON_IMPLEMENTATION", "python")
    try:
        import torch
        if not torch.cuda.is_available():
            os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
    except Exception:
        os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    for p in (REFERENCE_ROOT, spconv_path):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not _installed:
        import numpy as np
        _orig = np.meshgrid

        def _meshgrid_list(*a, **k):
            return list(_orig(*a, **k))

        np.meshgrid = _meshgrid_list
        _installed = True


def _slice_block(text, name):
    """brace-match the top-level ``name: { ... }`` block of a prototxt."""
    import re
    m = re.search(r"^\s*%s\s*:?\s*\{" % re.escape(name), text, flags=re.M)
    if m is None:
        raise KeyError(name)
    i = m.end() - 1
    depth = 0
    for j in range(i, len(text)):
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                return text[m.start():j + 1]
    raise ValueError("unbalanced braces in block %s" % name)


def load_config(rel_path):
    """-> pipeline_pb2.TrainEvalPipelineConfig holding only model + eval_input_reader."""
    from google.protobuf import text_format
    from second.protos import pipeline_pb2
    path = rel_path if os.path.isabs(rel_path) else os.path.join(REFERENCE_ROOT, "second", "configs", rel_path)
    with open(path) as f:
        lines = [ln.split("#", 1)[0] for ln in f.read().splitlines()]
    text = "\n".join(lines)
    sliced = _slice_block(text, "model") + "\n" + _slice_block(text, "eval_input_reader")
    cfg = pipeline_pb2.TrainEvalPipelineConfig()
    text_format.Merge(sliced, cfg)
    return cfg


def build_network(model_cfg):
    """the reference's own builders, chained as second/pytorch/train.py:58-68 does."""
    from second.builder import target_assigner_builder, voxel_builder
    from second.pytorch.builder import box_coder_builder, second_builder
    voxel_generator = voxel_builder.build(model_cfg.voxel_generator)
    bv_range = voxel_generator.point_cloud_range[[0, 1, 3, 4]]
    box_coder = box_coder_builder.build(model_cfg.box_coder)
    target_assigner = target_assigner_builder.build(model_cfg.target_assigner, bv_range, box_coder)
    box_coder.custom_ndim = target_assigner._anchor_generators[0].custom_ndim
    net = second_builder.build(model_cfg, voxel_generator, target_assigner, measure_time=False)
    return net


def generate_anchors(net, model_cfg):
    """feature-map rule of second/builder/dataset_builder.py:58,66-67 + notebook cell 7."""
    import numpy as np
    from second.utils.config_tool import get_downsample_factor
    grid_size = net.voxel_generator.grid_size
    factor = get_downsample_factor(model_cfg)
    feature_map_size = grid_size[:2] // factor
    feature_map_size = [*feature_map_size, 1][::-1]
    ret = net.target_assigner.generate_anchors(feature_map_size)
    return ret["anchors"].reshape(-1, ret["anchors"].shape[-1]).astype(np.float32)
