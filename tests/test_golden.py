"""The host-side mirror (b2second.models.VoxelNet) on the CPU oracle reproduces the golden fixtures that were
generated from the UNMODIFIED reference network (tests/golden/make_golden.py).  CPU only: pins the oracle +
mirror pair that every GPU parity test is then compared against."""
import numpy as np
import pytest
import torch

import golden_util as gu
from b2second import config, models

CASES = gu.cases()


def test_fixtures_present():
    names = {c[0] for c in CASES}
    assert {"car.fhd", "car.lite", "all.fhd", "pointpillars.car.xyres_16", "nuscenes.all.pp.largea"} <= names


@pytest.mark.parametrize("name,seed,n,path", CASES, ids=[f"{c[0]}-s{c[1]}-n{c[2]}" for c in CASES])
def test_mirror_on_oracle_reproduces_reference_golden(oracle, name, seed, n, path):
    fix = np.load(path)
    cfg = config.get_config(name)
    pts = gu.make_cloud(name, seed, n)
    assert gu.sha(pts) == str(fix["points_sha1"]), "synthetic cloud generator drifted"
    net = models.build_network(cfg, oracle).eval()
    models.synthetic_weights_(net, name, seed=0)
    anchors = net.anchors()
    assert gu.sha(anchors) == str(fix["anchors_sha1"]) and anchors.shape[0] == int(fix["num_anchors"])
    res = net.voxel_generator.generate(pts, gu.max_voxels_of(fix, name))
    assert res["voxel_num"] == int(fix["voxel_num"])
    assert gu.sha(res["coordinates"]) == str(fix["coords_sha1"])
    assert gu.sha(res["voxels"]) == str(fix["voxels_sha1"])
    coords = np.pad(res["coordinates"], ((0, 0), (1, 0)))
    ex = {"anchors": torch.from_numpy(anchors[None]), "voxels": torch.from_numpy(res["voxels"]),
          "num_points": torch.from_numpy(res["num_points_per_voxel"]), "coordinates": torch.from_numpy(coords)}
    with torch.no_grad():
        vf = net.voxel_feature_extractor(ex["voxels"], ex["num_points"], ex["coordinates"])
        sf = net.middle_feature_extractor(vf, ex["coordinates"], 1)
        out = net(ex)[0]
    assert list(sf.shape) == fix["bev_shape"].tolist()
    assert int((sf != 0).sum()) == int(fix["bev_nonzero"])
    np.testing.assert_allclose(sf.flatten()[fix["bev_sel_idx"]].numpy(), fix["bev_sel_val"], rtol=1e-5, atol=1e-5)
    gu.assert_detections_close({k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in out.items()}, fix)
    assert out["box3d_lidar"].shape[0] > 0
