"""Host-side logic that needs neither GPU nor reference: config dataclasses / prototxt reader, anchors, model
structure (state-dict key convention second.pytorch checkpoints use), synthetic data."""
import numpy as np
import pytest
import torch

from b2second import anchors, config, models, synth

PROTO = """
model: { second: {
  voxel_generator { point_cloud_range : [0, -40, -3, 70.4, 40, 1]  voxel_size : [0.05, 0.05, 0.1]
                    max_number_of_points_per_voxel : 5 }   # a comment
  voxel_feature_extractor: { module_class_name: "SimpleVoxel" num_filters: [16] with_distance: false num_input_features: 4 }
  middle_feature_extractor: { module_class_name: "SpMiddleFHD" downsample_factor: 8 num_input_features: 4 }
  rpn: { module_class_name: "RPNV2" layer_nums: [5] layer_strides: [1] num_filters: [128] upsample_strides: [1]
         num_upsample_filters: [128] use_groupnorm: false num_groups: 32 num_input_features: 128 }
  num_point_features: 4 use_sigmoid_score: true encode_background_as_zeros: true use_direction_classifier: true
  num_direction_bins: 2 direction_limit_offset: 1 post_center_limit_range: [0, -40, -2.2, 70.4, 40, 0.8]
  target_assigner: { class_settings: {
      anchor_generator_range: { sizes: [1.6, 3.9, 1.56] anchor_ranges: [0, -40.0, -1.00, 70.4, 40.0, -1.00] rotations: [0, 1.57] }
      class_name: "Car" use_rotate_nms: true use_multi_class_nms: false nms_pre_max_size: 1000 nms_post_max_size: 100
      nms_score_threshold: 0.3 nms_iou_threshold: 0.01 } }
}}
eval_input_reader: { batch_size: 8 preprocess: { max_number_of_voxels: 40000 anchor_area_threshold: -1 } }
"""


def test_prototxt_reader_matches_builtin():
    parsed = config.ModelConfig.from_prototxt(PROTO, "car.fhd")
    builtin = config.get_config("car.fhd")
    for f in builtin.__dataclass_fields__:
        assert getattr(parsed, f) == getattr(builtin, f), f


def test_derived_sizes():
    sizes = {"car.fhd": ([1408, 1600, 40], [1, 200, 176], 70400, 2),
             "car.lite": ([1056, 1280, 40], [1, 160, 132], 42240, 2),
             "all.fhd": ([1056, 1280, 40], [1, 160, 132], 168960, 8),
             "pointpillars.car.xyres_16": ([432, 496, 1], [1, 248, 216], 107136, 2),
             "nuscenes.all.pp.largea": ([400, 400, 1], [1, 50, 50], 30000, 12)}
    for name, (grid, fmap, A, aloc) in sizes.items():
        c = config.get_config(name)
        assert c.grid_size.tolist() == grid and c.feature_map_size == fmap
        assert c.num_anchors_per_loc == aloc
        a = anchors.generate_anchors(c)
        assert a.shape == (A, 7) and a.dtype == np.float32


def test_anchor_layout_matches_head_layout():
    c = config.get_config("car.fhd")
    a = anchors.generate_anchors(c).reshape(2, 200, 176, 7)          # (rot, y, x)
    assert np.all(a[0, :, :, 6] == 0) and np.allclose(a[1, :, :, 6], 1.57)
    assert a[0, 0, 0, 0] == 0 and np.isclose(a[0, 0, -1, 0], 70.4) and np.isclose(a[0, -1, 0, 1], 40.0)
    assert np.all(np.diff(a[0, 0, :, 0]) > 0) and np.all(np.diff(a[0, :, 0, 1]) > 0)
    assert np.allclose(a[..., 3:6], [1.6, 3.9, 1.56]) and np.allclose(a[..., 2], -1.0)


@pytest.mark.parametrize("name,nkeys", [("car.fhd", 133), ("pointpillars.car.xyres_16", 127)])
def test_state_dict_keys_follow_reference_convention(oracle, name, nkeys):
    net = models.build_network(name, oracle)
    sd = net.state_dict()
    assert len(sd) == nkeys                      # reference has +16 training-metric buffers
    assert "global_step" in sd and "rpn.conv_cls.bias" in sd and "rpn.blocks.0.1.weight" in sd
    if name == "car.fhd":
        assert sd["middle_feature_extractor.middle_conv.0.weight"].shape == (3, 3, 3, 4, 16)
        assert sd["middle_feature_extractor.middle_conv.39.weight"].shape == (3, 1, 1, 64, 64)
        assert sd["rpn.deblocks.0.0.weight"].shape == (128, 128, 1, 1)
    else:
        assert sd["voxel_feature_extractor.pfn_layers.0.linear.weight"].shape == (64, 9)
        assert sd["rpn.deblocks.2.0.weight"].shape == (256, 128, 4, 4)


def test_synthetic_weights_are_reproducible(oracle):
    a = models.synthetic_weights_(models.build_network("car.lite", oracle), "car.lite")
    b = models.synthetic_weights_(models.build_network("car.lite", oracle), "car.lite")
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k


def test_synthetic_clouds():
    p = synth.kitti_cloud(0, 20000)
    assert p.shape == (20000, 4) and p.dtype == np.float32
    assert p[:, 0].min() >= 0 and p[:, 0].max() < 70.4 and np.abs(p[:, 1]).max() <= 40
    assert np.array_equal(p, synth.kitti_cloud(0, 20000)) and not np.array_equal(p, synth.kitti_cloud(1, 20000))
    q = synth.nuscenes_cloud(0, 50000)
    assert q.shape == (50000, 4) and np.allclose(np.unique(q[:, 3]), 0.05 * np.arange(10), atol=1e-6)


def test_forward_contract_on_oracle(oracle):
    """dict in, list of dicts out; empty cloud -> zero-length tensors (voxelnet.py:629-643)."""
    cfg = config.get_config("car.lite")
    net = models.synthetic_weights_(models.build_network(cfg, oracle).eval(), "car.lite")
    pts = synth.kitti_cloud(2, 3000, cfg.point_cloud_range)
    res = net.voxel_generator.generate(pts, cfg.max_voxels)
    coords = np.pad(res["coordinates"], ((0, 0), (1, 0)))
    ex = {"anchors": torch.from_numpy(net.anchors()[None]), "voxels": torch.from_numpy(res["voxels"]),
          "num_points": torch.from_numpy(res["num_points_per_voxel"]), "coordinates": torch.from_numpy(coords),
          "metadata": [{"image_idx": 7}]}
    with torch.no_grad():
        out = net(ex)
    assert isinstance(out, list) and len(out) == 1
    assert set(out[0]) == {"box3d_lidar", "scores", "label_preds", "metadata"}
    assert out[0]["metadata"] == {"image_idx": 7} and out[0]["box3d_lidar"].shape[1] == 7
    assert out[0]["label_preds"].dtype == torch.int64
    with pytest.raises(AssertionError, match="num_anchors"):
        bad = dict(ex)
        bad["anchors"] = ex["anchors"][:, :100]
        net(bad)


def test_pack_sparse_weights_layout():
    """b2second.tc.pack_sparse_weights: the packed K-block GEMM (64 fp16 columns = PACK offsets x Cin channels, Cin 3/4
    zero-padded to 8) must equal the per-offset contraction sum_k in[nbr[o,k]] @ W[k] that b2s_sparse_conv_tc's gather
    assembles for Cin < 64."""
    import torch
    from b2second import tc
    g = torch.Generator().manual_seed(0)
    for cin, cout in ((3, 16), (4, 16), (16, 32), (32, 32), (64, 64)):
        K, n = 27, 50
        w = torch.randn(K, cin, cout, generator=g)
        x = torch.randn(n, cin, generator=g)
        nbr = torch.randint(-1, n, (n, K), generator=g)
        ref = torch.zeros(n, cout)
        for k in range(K):
            ok = nbr[:, k] >= 0
            ref[ok] += x[nbr[ok, k]] @ w[k]
        p = tc.pack_sparse_weights(w)
        cin_tc = tc.sparse_tc_cin(cin)
        if cin_tc >= 64:
            assert p.shape == (K, cout, cin) and torch.equal(p, w.transpose(1, 2))
            continue
        pack = 64 // cin_tc
        nkb = (K + pack - 1) // pack
        assert p.shape == (nkb, cout, 64)
        xp = torch.zeros(n, cin_tc)
        xp[:, :cin] = x                                 # the split kernel zero-pads the rows to cin_tc channels
        got = torch.zeros(n, cout)
        for kb in range(nkb):
            a = torch.zeros(n, 64)                      # the gathered A tile row: PACK neighbours side by side
            for ko in range(pack):
                k = kb * pack + ko
                if k >= K:
                    continue
                ok = nbr[:, k] >= 0
                a[ok, ko * cin_tc:(ko + 1) * cin_tc] = xp[nbr[ok, k]]
            got += a @ p[kb].t()
        assert float((got - ref).abs().max()) < 1e-4


def test_split_f16_is_fp32_grade():
    """the 3xF16 operand split (csrc/tc_common.cuh): hi + lo reproduces an fp32 value to 2^-22 relative where lo is a
    normal fp16, and to 2^-25 absolute below; the power-of-two weight pre-scale is exact."""
    import torch
    from b2second import tc
    g = torch.Generator().manual_seed(0)
    x = torch.randn(100000, generator=g) * 3
    hi, lo = tc.split_f16(x)
    err = (tc.merge_f16(hi, lo).double() - x.double()).abs()
    assert float((err / x.abs().clamp(min=2.0 ** -3).double()).max()) <= 2.0 ** -21
    assert float(err.max()) <= 2.0 ** -21 * float(x.abs().max())
    small = torch.randn(10000, generator=g) * 1e-3
    hs, ls = tc.split_f16(small)
    assert float((tc.merge_f16(hs, ls).double() - small.double()).abs().max()) <= 2.0 ** -25 * 1.0001
    w = torch.randn(9, 128, 128, generator=g) * 0.03
    s = tc.pow2_scale(w)
    assert 2.0 ** 12 < float(w.abs().max()) * s <= 2.0 ** 13 and float(np.log2(s)).is_integer()
    wh, wl = tc.split_f16(w, s)
    rel = ((tc.merge_f16(wh, wl).double() / s - w.double()).abs() / w.abs().double().clamp(min=float(w.abs().max()) * 2.0 ** -16))
    assert float(rel.max()) <= 2.0 ** -21
    big = torch.tensor([1e6, -1e6, 65504.0])
    hb, lb = tc.split_f16(big)
    assert bool(torch.isfinite(hb.float()).all()) and float(hb[0]) == 65504.0      # saturating, never inf
