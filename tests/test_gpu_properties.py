"""Size-independent properties of the CUDA path at BASELINE.json's full car.fhd size (29 k points, 41x1600x1408 grid),
where the CPU oracle would take too long per case: rulebook symmetry / sortedness / pair conservation, sparse-conv
linearity, run-to-run bit identity of the engine, NMS idempotence, voxelizer conservation laws."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def _full_size_tensor(product, frames=2, points=29000, cin=16, seed=0):
    from b2second import config, synth
    cfg = config.get_config("car.fhd")
    gen = product.utils.VoxelGeneratorV2(cfg.voxel_size, cfg.point_cloud_range, cfg.max_points_per_voxel, cfg.max_voxels)
    idx = []
    for b in range(frames):
        res = gen.generate(synth.kitti_cloud(seed + b, points, cfg.point_cloud_range), cfg.max_voxels)
        c = res["coordinates"]
        idx.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], 1))
    idx = torch.from_numpy(np.concatenate(idx, 0)).cuda()
    feats = torch.randn(idx.shape[0], cin, device="cuda")
    shape = [41, 1600, 1408]
    return product.SparseConvTensor(feats, idx, shape, frames), cfg


@pytest.mark.timeout(300)
def test_voxelizer_conservation_full_size(product):
    from b2second import config, synth
    cfg = config.get_config("car.fhd")
    gen = product.utils.VoxelGeneratorV2(cfg.voxel_size, cfg.point_cloud_range, cfg.max_points_per_voxel, cfg.max_voxels)
    pts = synth.kitti_cloud(3, 29000, cfg.point_cloud_range)
    res = gen.generate(pts, cfg.max_voxels)
    coords, num, vox = res["coordinates"], res["num_points_per_voxel"], res["voxels"]
    # every voxel coordinate is distinct and inside the grid (z, y, x order)
    keys = (coords[:, 0].astype(np.int64) * 1600 + coords[:, 1]) * 1408 + coords[:, 2]
    assert len(np.unique(keys)) == len(keys)
    assert (coords >= 0).all() and (coords[:, 0] < 40).all() and (coords[:, 1] < 1600).all() and (coords[:, 2] < 1408).all()
    # the voxel set is exactly the set of cells hit by in-range points; per-voxel counts are min(hits, T)
    lo = np.asarray(cfg.point_cloud_range[:3], np.float32)
    vs = np.asarray(cfg.voxel_size, np.float32)
    c = np.floor((pts[:, :3] - lo) / vs).astype(np.int64)
    ok = ((c >= 0) & (c < np.array([1408, 1600, 40]))).all(1)
    pk = (c[ok, 2] * 1600 + c[ok, 1]) * 1408 + c[ok, 0]
    uk, cnt = np.unique(pk, return_counts=True)
    assert len(uk) == len(keys) and np.array_equal(np.sort(keys), uk)
    order = np.argsort(keys)
    assert np.array_equal(num[order], np.minimum(cnt, cfg.max_points_per_voxel))
    # slots beyond num_points are zero padding
    T = vox.shape[1]
    mask = np.arange(T)[None, :] >= num[:, None]
    assert float(np.abs(vox[mask]).max(initial=0.0)) == 0.0


@pytest.mark.timeout(300)
def test_rulebook_properties_full_size(product):
    x, _ = _full_size_tensor(product)
    n = x.indices.shape[0]
    # SubM: symmetric relation, centre tap is the identity, pair count is even off-centre
    rb = product.ops.build_rulebook(x, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], True)
    nbr = rb.nbr.long()
    rows = torch.arange(n, device="cuda")
    assert torch.equal(nbr[:, 13], rows)
    for k in (0, 5, 12):
        j = nbr[:, k]
        has = j >= 0
        assert torch.equal(nbr[j[has], 26 - k], rows[has])
    # strided conv: outputs strictly ascending in flat (b,z,y,x) order, every input row appears exactly
    # ceil(3/2)^3-ish times as predicted by its own coordinates, and nothing else
    rb2 = product.ops.build_rulebook(x, [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1], False)
    oi = rb2.out_indices[:rb2.num_out].long()
    D, H, W = rb2.out_shape
    flat = ((oi[:, 0] * D + oi[:, 1]) * H + oi[:, 2]) * W + oi[:, 3]
    assert bool((flat[1:] > flat[:-1]).all())
    nb = rb2.nbr[:rb2.num_out].long()
    used = nb[nb >= 0]
    counts = torch.bincount(used, minlength=n)
    ci = x.indices.long()
    expect = torch.ones(n, dtype=torch.long, device="cuda")
    for d, size_out in zip((1, 2, 3), (D, H, W)):
        c = ci[:, d]
        m = torch.zeros_like(c)
        for kk in range(3):
            num = c + 1 - kk
            m += ((num >= 0) & (num % 2 == 0) & (num // 2 < size_out)).long()
        expect *= m
    assert torch.equal(counts, expect)
    # each (output, k) entry points at the input the geometry says
    o_sel = torch.randint(0, rb2.num_out, (4096,), device="cuda")
    for k in (0, 13, 26):
        kk = (k // 9, (k // 3) % 3, k % 3)
        src = nb[o_sel, k]
        has = src >= 0
        want = oi[o_sel][has][:, 1:] * 2 - 1 + torch.tensor(kk, device="cuda")
        assert torch.equal(ci[src[has]][:, 1:], want) and torch.equal(ci[src[has]][:, 0], oi[o_sel][has][:, 0])


@pytest.mark.timeout(300)
@pytest.mark.parametrize("cin,cout", [(16, 32), (64, 64)])
def test_sparse_conv_tc_linearity_full_size(product, cin, cout):
    from b2second import tc
    L = product._lib
    lib = L.load()
    x, _ = _full_size_tensor(product, cin=cin)
    rb = product.ops.build_rulebook(x, [3, 3, 3], [1, 1, 1], [1, 1, 1], [1, 1, 1], True)
    n = rb.num_out
    w = torch.randn(27, cin, cout, device="cuda") * 0.1
    wp = tc.pack_sparse_weights(w)
    ws = tc.pow2_scale(wp)
    w_hi, w_lo = tc.split_f16(wp, ws)
    inv = torch.full((cout,), 1.0 / ws, device="cuda")
    nbr = rb.nbr.contiguous()
    status = torch.zeros(1, dtype=torch.int32, device="cuda")

    def conv(f):
        buf = torch.empty(n, 2, cin, dtype=torch.float16, device="cuda")        # interleaved rows [hi | lo]
        buf[:, 0], buf[:, 1] = tc.split_f16(f)
        f_hi, f_lo = buf[:, 0], buf[:, 1]
        o = torch.zeros(n, cout, device="cuda")            # fp32 output variant (out_lo NULL)
        L.check(lib.b2s_sparse_conv_tc(L.ptr(f_hi), L.ptr(f_lo), 2 * cin, n, cin, L.ptr(w_hi), L.ptr(w_lo), L.ptr(nbr), 27,
                                       L.ptr(rb.num_out_dev), n, L.ptr(inv), None, 0, L.ptr(o), None, 0, cout,
                                       L.ptr(status), L.stream()), "b2s_sparse_conv_tc")
        torch.cuda.synchronize()
        return o

    a, b = x.features, torch.randn_like(x.features)
    ya, yb, yab = conv(a), conv(b), conv(2.5 * a - b)
    ref = 2.5 * ya - yb
    assert float((yab - ref).abs().max()) <= 3e-5 * max(float(ref.abs().max()), 1.0)
    # an isolated impulse reproduces the filter: row i with only channel c set -> output at neighbour j, offset k
    # equals W[k][c, :]
    imp = torch.zeros_like(a)
    i = n // 2
    imp[i, 3] = 1.0
    y = conv(imp)
    nb = rb.nbr.long()
    for k in (13, 0, 26, 4):
        outs = (nb[:, k] == i).nonzero().flatten()
        for o in outs.tolist():
            assert float((y[o] - w[k, 3]).abs().max()) <= 2e-6


@pytest.mark.timeout(300)
def test_engine_bit_identical_across_runs_and_batch_position(product):
    sys.path.insert(0, os.path.join(REPO, "second.pytorch_b200"))
    import bench
    from b2second import config, models
    from b2second.engine import InferenceEngine
    cfg = config.get_config("car.fhd")
    net = models.build_network(cfg, product).eval()
    models.synthetic_weights_(net, "car.fhd", seed=0)
    net = net.cuda()
    B = 4
    eng = InferenceEngine(net, batch_size=B, max_points=30000, use_cuda_graph=True)
    clouds = [torch.from_numpy(c).cuda() for c in bench.make_clouds("car.fhd", B, 29000, seed0=77)]
    eng.infer(clouds)
    torch.cuda.synchronize()
    d0, c0 = eng.det.clone(), eng.det_count.clone()
    eng.infer(clouds)
    torch.cuda.synchronize()
    assert torch.equal(eng.det_count, c0) and torch.equal(eng.det, d0)        # idempotent, no run-to-run noise
    assert int(c0.min()) > 0
    # a frame's detections do not depend on its slot in the batch or on its neighbours
    perm = [2, 0, 3, 1]
    eng.infer([clouds[p] for p in perm])
    torch.cuda.synchronize()
    for slot, p in enumerate(perm):
        n = int(c0[p])
        assert int(eng.det_count[slot]) == n
        assert torch.equal(eng.det[slot, :n], d0[p, :n])
    eng.check_status()


@pytest.mark.timeout(120)
def test_rotated_nms_idempotent(product):
    rng = np.random.default_rng(5)
    n = 1000
    ctr = rng.uniform(0, 60, (n, 2)).astype(np.float32)
    dims = rng.uniform(1.5, 4.5, (n, 2)).astype(np.float32)
    ang = rng.uniform(-np.pi, np.pi, (n, 1)).astype(np.float32)
    boxes = np.concatenate([ctr, dims, ang], 1)
    from b2second import box_ops
    corners = box_ops.corners_2d_np(boxes[:, :2], boxes[:, 2:4], boxes[:, 4]).astype(np.float32)
    standup = box_ops.standup_np(corners).astype(np.float32)

    def run(cor, st):
        iou = box_ops.standup_iou_np(st, eps=0.0)
        return product.utils.rotate_non_max_suppression_cpu(cor, np.arange(len(cor), dtype=np.int32), iou, 0.1)

    keep = np.asarray(run(corners, standup), np.int64)
    assert 0 < len(keep) < n and np.all(np.diff(keep) > 0)
    keep2 = np.asarray(run(corners[keep], standup[keep]), np.int64)
    assert np.array_equal(keep2, np.arange(len(keep)))                           # survivors do not suppress each other
