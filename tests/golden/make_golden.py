"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Runs only in the build container (needs /root/reference): the reference's own ``VoxelNet`` (built by its own
``second_builder`` from its own config files) executes on the CPU oracle ``spconv`` package, with the seeded
synthetic weights of ``b2second.models.synthetic_weights_`` and the seeded synthetic clouds of
``b2second.synth``.  What is stored (small): the detections, plus checksums / samples of every stage so a
mismatch can be localised.  Usage:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REPO, "second.pytorch_b200"))

from b2second import config, models, refcompat, synth  # noqa: E402

CASES = [  # (config name, cloud kind, seed, num_points[, max_voxels override])
    ("car.fhd", "kitti", 0, 20000),
    ("car.fhd", "kitti", 1, 29000),
    ("car.fhd", "kitti", 2, 24000),
    ("car.lite", "kitti", 0, 20000),
    ("car.lite", "kitti", 1, 29000),
    ("car.lite", "kitti", 2, 12000),
    ("all.fhd", "kitti", 0, 20000),
    ("all.fhd", "kitti", 1, 29000),
    ("all.fhd", "kitti", 2, 12000),
    ("pointpillars.car.xyres_16", "kitti", 0, 20000),
    ("pointpillars.car.xyres_16", "kitti", 1, 29000),
    ("pointpillars.car.xyres_16", "kitti", 2, 12000),
    ("nuscenes.all.pp.largea", "nuscenes", 0, 60000),
    ("nuscenes.all.pp.largea", "nuscenes", 1, 120000),
    # BASELINE config 5 size: 10 merged sweeps, ~300 k points per frame
    ("nuscenes.all.pp.largea", "nuscenes", 2, 300000),
    # more voxels than max_voxels: the extra voxels are dropped in first-come order (spconv semantics)
    ("car.fhd", "kitti", 3, 29000, 9000),
    ("pointpillars.car.xyres_16", "kitti", 3, 20000, 3000),
]


def make_cloud(kind, seed, n, pc_range):
    if kind == "nuscenes":
        return synth.nuscenes_cloud(seed, n)
    return synth.kitti_cloud(seed, n, pc_range)


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import torch
    refcompat.install(os.path.join(REPO, "oracle", "spconv_cpu"))
    torch.set_num_threads(8)
    force = "--force" in sys.argv
    for case in CASES:
        name, kind, seed, npts = case[:4]
        max_voxels = case[4] if len(case) > 4 else None
        path = os.path.join(HERE, "%s.seed%d.n%d%s.npz" % (name, seed, npts, ".mv%d" % max_voxels if max_voxels else ""))
        if os.path.exists(path) and not force:
            continue
        cfgp = refcompat.load_config(config.REFERENCE_FILES[name])
        mcfg = cfgp.model.second
        net = refcompat.build_network(mcfg).eval()
        models.synthetic_weights_(net, name, seed=0)
        anchors = refcompat.generate_anchors(net, mcfg)
        b = config.get_config(name)
        pts = make_cloud(kind, seed, npts, b.point_cloud_range)
        res = net.voxel_generator.generate(pts, max_voxels or b.max_voxels)
        coords = np.pad(res["coordinates"], ((0, 0), (1, 0)))
        ex = {"anchors": torch.from_numpy(anchors[None]), "voxels": torch.from_numpy(res["voxels"]),
              "num_points": torch.from_numpy(res["num_points_per_voxel"]), "coordinates": torch.from_numpy(coords)}
        with torch.no_grad():
            vf = net.voxel_feature_extractor(ex["voxels"], ex["num_points"], ex["coordinates"])
            sf = net.middle_feature_extractor(vf, ex["coordinates"], 1)
            pd = net.rpn(sf)
            out = net(ex)[0]
        # per detection: margin between the best and second-best class logit of its anchor (label ties)
        cl = pd["cls_preds"].reshape(-1, b.num_class)
        top = torch.sigmoid(cl).max(1)[0]
        margins = []
        for sc in out["scores"]:
            a = int(torch.argmin((top - sc).abs()))
            v = torch.sort(cl[a], descending=True)[0]
            margins.append(float(v[0] - v[1]) if v.numel() > 1 else 1e9)
        sel = np.random.default_rng(0).choice(sf.numel(), 256, replace=False)
        nz = torch.nonzero(sf.flatten()).flatten().numpy()
        sel_nz = nz[np.random.default_rng(1).choice(nz.size, min(256, nz.size), replace=False)]
        fix = {
            "points_sha1": sha(pts), "num_points": pts.shape[0], "voxel_num": res["voxel_num"],
            "max_voxels": int(max_voxels or b.max_voxels),
            "coords_sha1": sha(res["coordinates"]), "num_points_per_voxel_sha1": sha(res["num_points_per_voxel"]),
            "voxels_sha1": sha(res["voxels"]), "coords_head": res["coordinates"][:64],
            "vfe_sum": float(vf.double().sum()), "vfe_abs": float(vf.double().abs().sum()),
            "bev_shape": np.array(sf.shape), "bev_nonzero": int((sf != 0).sum()),
            "bev_sum": float(sf.double().sum()), "bev_abs": float(sf.double().abs().sum()),
            "bev_sel_idx": sel_nz, "bev_sel_val": sf.flatten()[sel_nz].numpy(),
            "box_abs": float(pd["box_preds"].double().abs().sum()), "cls_abs": float(pd["cls_preds"].double().abs().sum()),
            "box_sel_idx": sel % pd["box_preds"].numel(), "box_sel_val": pd["box_preds"].flatten()[sel % pd["box_preds"].numel()].numpy(),
            "cls_sel_idx": sel % pd["cls_preds"].numel(), "cls_sel_val": pd["cls_preds"].flatten()[sel % pd["cls_preds"].numel()].numpy(),
            "anchors_sha1": sha(anchors), "num_anchors": anchors.shape[0],
            "box3d_lidar": out["box3d_lidar"].numpy(), "scores": out["scores"].numpy(),
            "label_preds": out["label_preds"].numpy(), "label_margin": np.array(margins, np.float32),
            "num_pass_threshold": int((torch.sigmoid(pd["cls_preds"]).reshape(-1, b.num_class).max(1)[0]
                                       >= b.nms_score_threshold).sum()),
        }
        np.savez_compressed(path, **fix)
        print(name, seed, npts, "voxels", res["voxel_num"], "pass", fix["num_pass_threshold"], "dets",
              out["box3d_lidar"].shape[0], "->", os.path.basename(path), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
