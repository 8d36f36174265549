"""Model configuration for the SECOND inference hot path.

The reference describes a model with protobuf text configs (second/protos/second.proto,
second/configs/*.config) turned into objects by second/pytorch/builder/second_builder.py:22-133 and
second/pytorch/train.py:58-68.  Here the same information is a plain dataclass so that the host
side needs neither protobuf nor the reference tree:

* ``parse_prototxt`` / ``ModelConfig.from_prototxt`` read a reference-format ``.config`` file
  (only the ``model.second`` and ``eval_input_reader`` blocks are interpreted);
* ``BUILTIN`` holds the five BASELINE.json configurations, written out by hand from the values in
  SURVEY.md App. B (car.fhd, car.lite, all.fhd, pointpillars/car/xyres_16, nuscenes/all.pp.largea).
"""
import re
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class AnchorClass:
    class_name: str
    sizes: List[float] = field(default_factory=list)          # flat [w,l,h]*n ; empty = no_anchor
    rotations: List[float] = field(default_factory=lambda: [0.0, 1.57])
    anchor_ranges: Optional[List[float]] = None               # range generator (x0,y0,z0,x1,y1,z1)
    strides: Optional[List[float]] = None                     # stride generator
    offsets: Optional[List[float]] = None
    custom_values: List[float] = field(default_factory=list)

    @property
    def num_anchors_per_loc(self):
        if not self.sizes:
            return 0
        return (len(self.sizes) // 3) * len(self.rotations)


@dataclass
class ModelConfig:
    name: str
    point_cloud_range: List[float]
    voxel_size: List[float]
    max_points_per_voxel: int
    max_voxels: int                         # eval_input_reader.preprocess.max_number_of_voxels
    num_point_features: int
    vfe_class: str
    vfe_num_filters: List[int]
    vfe_with_distance: bool
    middle_class: str
    middle_num_input_features: int
    middle_downsample_factor: int
    rpn_class: str
    rpn_layer_nums: List[int]
    rpn_layer_strides: List[int]
    rpn_num_filters: List[int]
    rpn_upsample_strides: List[float]
    rpn_num_upsample_filters: List[int]
    rpn_num_input_features: int
    classes: List[AnchorClass]
    use_rotate_nms: bool
    use_multi_class_nms: bool
    nms_pre_max_size: int
    nms_post_max_size: int
    nms_score_threshold: float
    nms_iou_threshold: float
    post_center_limit_range: List[float]
    use_sigmoid_score: bool = True
    encode_background_as_zeros: bool = True
    use_direction_classifier: bool = True
    num_direction_bins: int = 2
    direction_limit_offset: float = 0.0
    direction_offset: float = 0.0
    nms_class_agnostic: bool = False
    eval_batch_size: int = 1
    anchor_area_threshold: float = -1.0
    box_code_size: int = 7

    # ---- derived quantities (mirrors of reference helpers) -------------------------------
    @property
    def grid_size(self):
        """xyz int64, as VoxelGeneratorV2.grid_size (round((hi-lo)/vs))."""
        r = np.array(self.point_cloud_range, dtype=np.float32)
        vs = np.array(self.voxel_size, dtype=np.float32)
        return np.round((r[3:] - r[:3]) / vs).astype(np.int64)

    @property
    def num_class(self):
        return len(self.classes)

    @property
    def num_anchors_per_loc(self):
        return sum(c.num_anchors_per_loc for c in self.classes)

    @property
    def downsample_factor(self):
        """second/utils/config_tool/__init__.py:45-52."""
        f = float(np.prod(self.rpn_layer_strides))
        if len(self.rpn_upsample_strides) > 0:
            f /= self.rpn_upsample_strides[-1]
        f *= self.middle_downsample_factor
        f = int(np.round(f))
        assert f > 0
        return f

    @property
    def feature_map_size(self):
        """[1, H, W] -- second/builder/dataset_builder.py:58,66-67."""
        fm = self.grid_size[:2] // self.downsample_factor
        return [1, int(fm[1]), int(fm[0])]

    @property
    def dense_shape(self):
        """[1, D, H, W, C] as second_builder.py:31."""
        g = self.grid_size
        return [1, int(g[2]), int(g[1]), int(g[0]), self.vfe_num_filters[-1]]

    # ---- prototxt -----------------------------------------------------------------------
    @staticmethod
    def from_prototxt(text, name="custom"):
        tree = parse_prototxt(text)
        m = _one(_one(tree["model"])["second"])
        ev = _one(tree["eval_input_reader"]) if "eval_input_reader" in tree else {}
        prep = _one(ev["preprocess"]) if "preprocess" in ev else {}
        vg = _one(m["voxel_generator"])
        vfe = _one(m["voxel_feature_extractor"])
        mid = _one(m["middle_feature_extractor"])
        rpn = _one(m["rpn"])
        ta = _one(m["target_assigner"])
        classes = []
        for cs in ta.get("class_settings", []):
            ac = AnchorClass(class_name=_one(cs.get("class_name", [""])))
            if "anchor_generator_range" in cs:
                g = _one(cs["anchor_generator_range"])
                ac.sizes = [float(v) for v in g["sizes"]]
                ac.rotations = [float(v) for v in g["rotations"]]
                ac.anchor_ranges = [float(v) for v in g["anchor_ranges"]]
                ac.custom_values = [float(v) for v in g.get("custom_values", [])]
            elif "anchor_generator_stride" in cs:
                g = _one(cs["anchor_generator_stride"])
                ac.sizes = [float(v) for v in g["sizes"]]
                ac.rotations = [float(v) for v in g["rotations"]]
                ac.strides = [float(v) for v in g["strides"]]
                ac.offsets = [float(v) for v in g["offsets"]]
                ac.custom_values = [float(v) for v in g.get("custom_values", [])]
            classes.append((ac, cs))
        first = classes[0][1]

        def cls_val(key, default):
            vals = [c[1].get(key, [default])[-1] for c in classes]
            assert all(v == vals[0] for v in vals), "per-class %s must agree (second_builder.py:55-59)" % key
            return vals[0]

        return ModelConfig(
            name=name,
            point_cloud_range=[float(v) for v in vg["point_cloud_range"]],
            voxel_size=[float(v) for v in vg["voxel_size"]],
            max_points_per_voxel=int(_one(vg["max_number_of_points_per_voxel"])),
            max_voxels=int(_one(prep.get("max_number_of_voxels", [20000]))),
            num_point_features=int(_one(m.get("num_point_features", [4]))),
            vfe_class=_one(vfe["module_class_name"]),
            vfe_num_filters=[int(v) for v in vfe.get("num_filters", [])],
            vfe_with_distance=bool(_one(vfe.get("with_distance", [False]))),
            middle_class=_one(mid["module_class_name"]),
            middle_num_input_features=int(_one(mid.get("num_input_features", [-1]))),
            middle_downsample_factor=int(_one(mid.get("downsample_factor", [1]))),
            rpn_class=_one(rpn["module_class_name"]),
            rpn_layer_nums=[int(v) for v in rpn["layer_nums"]],
            rpn_layer_strides=[int(v) for v in rpn["layer_strides"]],
            rpn_num_filters=[int(v) for v in rpn["num_filters"]],
            rpn_upsample_strides=[float(v) for v in rpn["upsample_strides"]],
            rpn_num_upsample_filters=[int(v) for v in rpn["num_upsample_filters"]],
            rpn_num_input_features=int(_one(rpn["num_input_features"])),
            classes=[c[0] for c in classes],
            use_rotate_nms=bool(cls_val("use_rotate_nms", False)),
            use_multi_class_nms=bool(cls_val("use_multi_class_nms", False)),
            nms_pre_max_size=int(cls_val("nms_pre_max_size", 1000)),
            nms_post_max_size=int(cls_val("nms_post_max_size", 100)),
            nms_score_threshold=float(cls_val("nms_score_threshold", 0.0)),
            nms_iou_threshold=float(cls_val("nms_iou_threshold", 0.5)),
            post_center_limit_range=[float(v) for v in m.get("post_center_limit_range", [])],
            use_sigmoid_score=bool(_one(m.get("use_sigmoid_score", [False]))),
            encode_background_as_zeros=bool(_one(m.get("encode_background_as_zeros", [False]))),
            use_direction_classifier=bool(_one(m.get("use_direction_classifier", [False]))),
            num_direction_bins=int(_one(m.get("num_direction_bins", [2]))),
            direction_limit_offset=float(_one(m.get("direction_limit_offset", [0.0]))),
            direction_offset=float(_one(m.get("direction_offset", [0.0]))),
            nms_class_agnostic=bool(_one(m.get("nms_class_agnostic", [False]))),
            eval_batch_size=int(_one(ev.get("batch_size", [1]))),
            anchor_area_threshold=float(_one(prep.get("anchor_area_threshold", [-1]))),
            box_code_size=7 + len(classes[0][0].custom_values),
        )

    @staticmethod
    def from_file(path, name=None):
        with open(path) as f:
            return ModelConfig.from_prototxt(f.read(), name or path)


def _one(v):
    assert isinstance(v, list) and len(v) >= 1
    return v[-1]


_TOKEN = re.compile(r'\s*(?:(#[^\n]*)|([{}\[\]:,])|"((?:[^"\\]|\\.)*)"|([^\s{}\[\]:,#"]+))')


def parse_prototxt(text):
    """Minimal protobuf text-format reader: -> dict name -> list of values (scalars or dicts)."""
    toks = []
    pos = 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if m is None:
            if text[pos:].strip() == "":
                break
            raise ValueError("prototxt: cannot tokenise at %d: %r" % (pos, text[pos:pos + 20]))
        pos = m.end()
        if m.group(1) is not None:
            continue
        if m.group(2) is not None:
            toks.append(("p", m.group(2)))
        elif m.group(3) is not None:
            toks.append(("s", m.group(3)))
        else:
            toks.append(("w", m.group(4)))

    def scalar(kind, tok):
        if kind == "s":
            return tok
        if tok in ("true", "True"):
            return True
        if tok in ("false", "False"):
            return False
        try:
            if re.fullmatch(r"[-+]?\d+", tok):
                return int(tok)
            return float(tok)
        except ValueError:
            return tok  # enum identifier

    def block(i, closing):
        out = {}
        while i < len(toks):
            kind, tok = toks[i]
            if kind == "p" and tok == closing:
                return out, i + 1
            if kind != "w":
                raise ValueError("prototxt: expected field name, got %r" % (tok,))
            name = tok
            i += 1
            if i < len(toks) and toks[i] == ("p", ":"):
                i += 1
            kind, tok = toks[i]
            if kind == "p" and tok == "{":
                val, i = block(i + 1, "}")
                out.setdefault(name, []).append(val)
            elif kind == "p" and tok == "[":
                i += 1
                while toks[i] != ("p", "]"):
                    if toks[i] == ("p", ","):
                        i += 1
                        continue
                    out.setdefault(name, []).append(scalar(*toks[i]))
                    i += 1
                out.setdefault(name, [])
                i += 1
            else:
                out.setdefault(name, []).append(scalar(kind, tok))
                i += 1
        if closing is not None:
            raise ValueError("prototxt: unbalanced braces")
        return out, i

    tree, _ = block(0, None)
    return tree


# ------------------------------------------------------------------------------------------
# The five BASELINE.json configurations (values: SURVEY.md App. B and the cited config lines)
# ------------------------------------------------------------------------------------------
def _car_fhd():
    # second/configs/car.fhd.config:1-110,205-217
    return ModelConfig(
        name="car.fhd", point_cloud_range=[0, -40, -3, 70.4, 40, 1], voxel_size=[0.05, 0.05, 0.1],
        max_points_per_voxel=5, max_voxels=40000, num_point_features=4,
        vfe_class="SimpleVoxel", vfe_num_filters=[16], vfe_with_distance=False,
        middle_class="SpMiddleFHD", middle_num_input_features=4, middle_downsample_factor=8,
        rpn_class="RPNV2", rpn_layer_nums=[5], rpn_layer_strides=[1], rpn_num_filters=[128],
        rpn_upsample_strides=[1], rpn_num_upsample_filters=[128], rpn_num_input_features=128,
        classes=[AnchorClass("Car", [1.6, 3.9, 1.56], [0, 1.57], anchor_ranges=[0, -40.0, -1.0, 70.4, 40.0, -1.0])],
        use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=100,
        nms_score_threshold=0.3, nms_iou_threshold=0.01,
        post_center_limit_range=[0, -40, -2.2, 70.4, 40, 0.8], direction_limit_offset=1.0,
        eval_batch_size=8)


def _car_lite():
    # second/configs/car.lite.config
    return ModelConfig(
        name="car.lite", point_cloud_range=[0, -32.0, -3, 52.8, 32.0, 1], voxel_size=[0.05, 0.05, 0.1],
        max_points_per_voxel=1, max_voxels=30000, num_point_features=4,
        vfe_class="SimpleVoxelRadius", vfe_num_filters=[16], vfe_with_distance=False,
        middle_class="SpMiddleFHDLite", middle_num_input_features=3, middle_downsample_factor=8,
        rpn_class="RPNV2", rpn_layer_nums=[5], rpn_layer_strides=[1], rpn_num_filters=[128],
        rpn_upsample_strides=[1], rpn_num_upsample_filters=[128], rpn_num_input_features=128,
        classes=[AnchorClass("Car", [1.6, 3.9, 1.56], [0, 1.57], anchor_ranges=[0, -32.0, -1.0, 52.8, 32.0, -1.0])],
        use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=100,
        nms_score_threshold=0.3, nms_iou_threshold=0.01,
        post_center_limit_range=[0, -40, -2.2, 70.4, 40, 0.8], direction_limit_offset=1.0,
        eval_batch_size=12)


def _all_fhd():
    # second/configs/all.fhd.config
    def rng(z):
        return [0, -32.0, z, 52.8, 32.0, z]
    return ModelConfig(
        name="all.fhd", point_cloud_range=[0, -32.0, -3, 52.8, 32.0, 1], voxel_size=[0.05, 0.05, 0.1],
        max_points_per_voxel=5, max_voxels=60000, num_point_features=4,
        vfe_class="SimpleVoxelRadius", vfe_num_filters=[16], vfe_with_distance=False,
        middle_class="SpMiddleFHD", middle_num_input_features=3, middle_downsample_factor=8,
        rpn_class="RPNV2", rpn_layer_nums=[5, 5], rpn_layer_strides=[1, 2], rpn_num_filters=[64, 128],
        rpn_upsample_strides=[1, 2], rpn_num_upsample_filters=[128, 128], rpn_num_input_features=128,
        classes=[
            AnchorClass("Car", [1.6, 3.9, 1.56], [0, 1.57], anchor_ranges=rng(-1.0)),
            AnchorClass("Cyclist", [0.6, 1.76, 1.73], [0, 1.57], anchor_ranges=rng(-0.6)),
            AnchorClass("Pedestrian", [0.6, 0.8, 1.73], [0, 1.57], anchor_ranges=rng(-0.6)),
            AnchorClass("Van", [1.87103749, 5.02808195, 2.20964255], [0, 1.57], anchor_ranges=rng(-1.41)),
        ],
        use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=100,
        nms_score_threshold=0.3, nms_iou_threshold=0.1,
        post_center_limit_range=[0, -40, -2.2, 70.4, 40, 0.8], direction_limit_offset=1.0,
        eval_batch_size=3)


def _pp_xyres16():
    # second/configs/pointpillars/car/xyres_16.config
    return ModelConfig(
        name="pointpillars.car.xyres_16", point_cloud_range=[0, -39.68, -3, 69.12, 39.68, 1],
        voxel_size=[0.16, 0.16, 4], max_points_per_voxel=100, max_voxels=12000, num_point_features=4,
        vfe_class="PillarFeatureNet", vfe_num_filters=[64], vfe_with_distance=False,
        middle_class="PointPillarsScatter", middle_num_input_features=64, middle_downsample_factor=1,
        rpn_class="RPNV2", rpn_layer_nums=[3, 5, 5], rpn_layer_strides=[2, 2, 2],
        rpn_num_filters=[64, 128, 256], rpn_upsample_strides=[1, 2, 4],
        rpn_num_upsample_filters=[128, 128, 128], rpn_num_input_features=64,
        classes=[AnchorClass("Car", [1.6, 3.9, 1.56], [0, 1.57], strides=[0.32, 0.32, 0.0],
                             offsets=[0.16, -39.52, -1.78])],
        use_rotate_nms=False, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=300,
        nms_score_threshold=0.05, nms_iou_threshold=0.5,
        post_center_limit_range=[0, -39.68, -5, 69.12, 39.68, 5], direction_limit_offset=1.0,
        eval_batch_size=2, anchor_area_threshold=1.0)


def _nusc_largea():
    # second/configs/nuscenes/all.pp.largea.config
    def rng(z):
        return [-50, -50, z, 50, 50, z]
    return ModelConfig(
        name="nuscenes.all.pp.largea", point_cloud_range=[-50, -50, -10, 50, 50, 10],
        voxel_size=[0.25, 0.25, 20], max_points_per_voxel=60, max_voxels=30000, num_point_features=4,
        vfe_class="PillarFeatureNet", vfe_num_filters=[64], vfe_with_distance=False,
        middle_class="PointPillarsScatter", middle_num_input_features=64, middle_downsample_factor=1,
        rpn_class="RPNV2", rpn_layer_nums=[3, 5, 5], rpn_layer_strides=[2, 2, 2],
        rpn_num_filters=[64, 128, 256], rpn_upsample_strides=[0.25, 0.5, 1],
        rpn_num_upsample_filters=[128, 128, 128], rpn_num_input_features=64,
        classes=[
            AnchorClass("car", [1.95017717, 4.60718145, 1.72270761], [0, 1.57], anchor_ranges=rng(-0.93897414)),
            AnchorClass("bus", [2.94046906, 11.1885991, 3.47030982], [0, 1.57], anchor_ranges=rng(-0.0715754)),
            AnchorClass("construction_vehicle", [2.73050468, 6.38352896, 3.13312415], [0, 1.57],
                        anchor_ranges=rng(-0.08168083)),
            AnchorClass("trailer", [3, 15, 3.8, 2, 3, 3.8], [0, 1.57], anchor_ranges=rng(0.22228277)),
            AnchorClass("truck", [2.4560939, 6.73778078, 2.73004906], [0, 1.57], anchor_ranges=rng(-0.37937912)),
            AnchorClass("bicycle"), AnchorClass("motorcycle"), AnchorClass("pedestrian"),
            AnchorClass("traffic_cone"), AnchorClass("barrier"),
        ],
        use_rotate_nms=False, use_multi_class_nms=False, nms_pre_max_size=1000, nms_post_max_size=300,
        nms_score_threshold=0.05, nms_iou_threshold=0.5,
        post_center_limit_range=[-59.6, -59.6, -10, 59.6, 59.6, 10], direction_limit_offset=0.0,
        direction_offset=0.78, nms_class_agnostic=True, eval_batch_size=1)


BUILTIN = {
    "car.fhd": _car_fhd,
    "car.lite": _car_lite,
    "all.fhd": _all_fhd,
    "pointpillars.car.xyres_16": _pp_xyres16,
    "nuscenes.all.pp.largea": _nusc_largea,
}

# reference config file (relative to second/configs) for each builtin -- used by the container-only
# drop-in tests to check the hand-written dataclasses against the parsed reference files.
REFERENCE_FILES = {
    "car.fhd": "car.fhd.config",
    "car.lite": "car.lite.config",
    "all.fhd": "all.fhd.config",
    "pointpillars.car.xyres_16": "pointpillars/car/xyres_16.config",
    "nuscenes.all.pp.largea": "nuscenes/all.pp.largea.config",
}


def get_config(name):
    if name not in BUILTIN:
        raise KeyError("unknown config %r (have %s)" % (name, sorted(BUILTIN)))
    return BUILTIN[name]()
