"""The fused engine behind the reference's own call: ``net(example)``.

``accelerate(net)`` takes ANY VoxelNet-like module -- the object the reference's
``second.pytorch.builder.second_builder.build`` returns (running on this repository's ``spconv`` package), or the
mirror ``b2second.models.VoxelNet`` -- reads it through ``b2second.spec`` and re-binds ``net.forward`` so that an
eval-mode call

    preds = net(example)          # second/pytorch/models/voxelnet.py:339-375 contract: dict in, list of dicts out

runs as one CUDA graph per frame batch (``b2second.engine.InferenceEngine``).  Accepted examples:

* the reference's own dict (``voxels [N,T,F]``, ``num_points [N]``, ``coordinates [N,4] (b,z,y,x)``, ``anchors
  [B,A,7]``, optional ``anchors_mask``, ``metadata``) -- tensors on any device; the voxels are copied into the
  engine's static buffers and everything after the voxelizer is fused;
* the same dict with ``points`` (a list of B float32 ``[P_i, F]`` tensors: CUDA, pinned or plain host memory)
  instead of the three voxel keys -- then the voxelizer runs on the GPU as well (SURVEY.md §8b "a fast path may
  additionally accept example['points']"); with ``crop_planes`` (``b2second.inputs.frustum_planes``) the raw clouds
  are first cropped to the camera field of view on the GPU (the ``velodyne_reduced`` step, box_np_ops.py:682-693);
* ``sweeps``: per frame the NuScenes key-frame cloud + its sweeps with their ``sweep2lidar`` calibration, merged on
  the GPU (nuscenes_dataset.py:166-185) -- see ``InferenceEngine.load_sweeps``.

The list of result dicts holds GPU tensors like the reference's (``output="device"``); ``accelerate(net,
output="host")`` returns CPU tensors through ONE pinned device-to-host copy of the detection records instead of one
small copy per frame and field.

Anything the engine does not cover (training mode, the DataParallel padded layout, a layer pattern
``spec_from_module`` rejects) falls through to the module's original ``forward`` -- which still computes on the CUDA
``spconv`` drop-in; there is no CPU path anywhere.
"""
import types

import torch

from . import spec as _spec


class FastPath:
    def __init__(self, net, max_points=40000, max_voxels=None, use_cuda_graph=True, output="device", **engine_kw):
        self.net = net
        assert output in ("device", "host")
        self.output = output         # "device": tensors stay on the GPU like the reference's; "host": CPU tensors
                                     # through one pinned D2H of the whole detection record buffer
        self.post_run = None         # optional hook(engine) between the launch and the read-back (e.g. the
                                     # multi-GPU all-gather of engine.det_record, b2second/dist.py)
        self.max_points = int(max_points)
        self.use_cuda_graph = use_cuda_graph
        self.engine_kw = engine_kw
        self.spec = _spec.spec_from_module(net, max_voxels)      # raises UnsupportedNetwork for foreign layer patterns
        self.engines = {}

    def engine(self, batch_size):
        eng = self.engines.get(batch_size)
        if eng is None:
            from .engine import InferenceEngine
            eng = InferenceEngine(self.spec, batch_size=batch_size, max_points=self.max_points,
                                  use_cuda_graph=self.use_cuda_graph, **self.engine_kw)
            self.engines[batch_size] = eng
        return eng

    def applicable(self, example):
        if self.net.training:
            return False
        if "points" in example or "sweeps" in example:
            return True
        if not all(k in example for k in ("voxels", "num_points", "coordinates")):
            return False
        return example["num_points"].dim() == 1          # 2-D = DataParallel padded layout (voxelnet.py:345-357)

    def __call__(self, example):
        if "points" in example or "sweeps" in example:
            frames = example["points"] if "points" in example else example["sweeps"]
            B = len(frames)
        else:
            B = int(example["anchors"].shape[0])
        eng = self.engine(B)
        if "anchors" in example and example["anchors"] is not None:
            anchors = example["anchors"]
            eng.set_anchors(anchors.reshape(anchors.shape[0], -1, anchors.shape[-1]))   # asserts the anchor count
        elif eng._anchors_key is None:
            raise KeyError("example has no 'anchors' and the network cannot generate them")
        eng.set_anchors_mask(example.get("anchors_mask"))
        if "sweeps" in example:
            eng.load_sweeps(list(frames))             # NuScenes: key frame + sweeps merged on the GPU
            eng.run("points")
        elif "points" in example and example.get("crop_planes") is not None:
            eng.load_points_cropped(list(frames), example["crop_planes"])    # KITTI: camera-FOV crop on the GPU
            eng.run("points")
        elif "points" in example:
            eng.infer(list(frames))
        else:
            eng.infer_voxels(example["voxels"], example["num_points"], example["coordinates"])
        if self.post_run is not None:
            self.post_run(eng)
        return eng.detections(metadata=example.get("metadata"), output=self.output)


def accelerate(net, **kw):
    """re-bind ``net.forward`` to the fused engine (see module docstring); returns ``net``.
    ``net.b2s_fastpath`` is the FastPath object, ``net.b2s_reference_forward`` the original bound method."""
    if getattr(net, "b2s_fastpath", None) is not None:
        return net
    fast = FastPath(net, **kw)
    original = net.forward

    def forward(self, example):
        if fast.applicable(example):
            with torch.no_grad():
                return fast(example)
        return original(example)

    net.b2s_fastpath = fast
    net.b2s_reference_forward = original
    net.forward = types.MethodType(forward, net)
    return net
