"""Known-answer tests that pin the oracle's sparse convolution (none exist upstream in-tree):
sparse conv == dense F.conv3d (weight permuted (4,3,0,1,2)) sampled at the active output sites, and the
output-site set == {conv3d(occupancy, ones) > 0} (strided) / the input set (SubM).  SURVEY.md §8(c)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def random_sparse(rng, shape, batch, n, cin):
    D, H, W = shape
    flat = rng.choice(batch * D * H * W, size=n, replace=False)
    b, r = np.divmod(flat, D * H * W)
    z, r = np.divmod(r, H * W)
    y, x = np.divmod(r, W)
    idx = np.stack([b, z, y, x], 1).astype(np.int32)
    feats = rng.standard_normal((n, cin)).astype(np.float32)
    return torch.from_numpy(feats), torch.from_numpy(idx)


def densify(feats, idx, shape, batch):
    d = torch.zeros(batch, feats.shape[1], *shape)
    i = idx.long()
    d[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]] = feats
    return d


CASES = [
    # (shape, k, s, p, subm)
    ((9, 12, 10), 3, 1, 0, True),
    ((9, 12, 10), 3, 2, 1, False),
    ((11, 12, 10), 3, 2, [0, 1, 1], False),      # middle.py:175-176
    ((5, 12, 10), (3, 1, 1), (2, 1, 1), 0, False),  # middle.py:188-189
    ((7, 9, 8), 3, 1, 1, False),
    ((7, 9, 8), (1, 3, 3), 1, 0, True),
]


@pytest.mark.parametrize("shape,k,s,p,subm", CASES)
def test_sparse_conv_equals_dense_conv(oracle, shape, k, s, p, subm):
    rng = np.random.default_rng(hash((shape, subm)) % 2**31)
    batch, cin, cout = 2, 5, 7
    n = int(0.15 * batch * np.prod(shape))
    feats, idx = random_sparse(rng, shape, batch, n, cin)
    cls = oracle.SubMConv3d if subm else oracle.SparseConv3d
    conv = cls(cin, cout, k, s, padding=p, bias=True)
    x = oracle.SparseConvTensor(feats, idx, shape, batch)
    with torch.no_grad():
        y = conv(x)
        w = conv.weight.permute(4, 3, 0, 1, 2).contiguous()
        dense_in = densify(feats, idx, shape, batch)
        if subm:
            ks = conv.kernel_size
            ref = F.conv3d(dense_in, w, None, stride=1, padding=[kk // 2 for kk in ks])
            occ_out = densify(torch.ones(n, 1), idx, shape, batch)[:, 0] > 0
        else:
            ref = F.conv3d(dense_in, w, None, stride=conv.stride, padding=conv.padding)
            occ = densify(torch.ones(n, 1), idx, shape, batch)
            ones = torch.ones(1, 1, *conv.kernel_size)
            occ_out = F.conv3d(occ, ones, None, stride=conv.stride, padding=conv.padding)[:, 0] > 0
    assert list(ref.shape[2:]) == list(y.spatial_shape)
    # output site set
    got = torch.zeros_like(occ_out)
    oi = y.indices.long()
    got[oi[:, 0], oi[:, 1], oi[:, 2], oi[:, 3]] = True
    assert torch.equal(got, occ_out)
    assert oi.shape[0] == int(occ_out.sum())
    # values (bias is added at active output sites only)
    samp = ref[oi[:, 0], :, oi[:, 1], oi[:, 2], oi[:, 3]] + conv.bias
    torch.testing.assert_close(y.features, samp, rtol=1e-4, atol=1e-5)
    if not subm:
        # sorted ascending by flat (b,z,y,x) key
        D, H, W = y.spatial_shape
        keys = ((oi[:, 0] * D + oi[:, 1]) * H + oi[:, 2]) * W + oi[:, 3]
        assert torch.all(keys[1:] > keys[:-1])
    else:
        assert torch.equal(y.indices, x.indices)


def test_output_shape_rule_matches_reference_comments(oracle):
    # middle.py:152-189: 41->21->11->5->2, 1600->800->400->200, 1408->704->352->176
    g = oracle.ops.get_conv_output_size
    s = g([41, 1600, 1408], [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1])
    assert s == [21, 800, 704]
    s = g(s, [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1])
    assert s == [11, 400, 352]
    s = g(s, [3, 3, 3], [2, 2, 2], [0, 1, 1], [1, 1, 1])
    assert s == [5, 200, 176]
    s = g(s, [3, 1, 1], [2, 1, 1], [0, 0, 0], [1, 1, 1])
    assert s == [2, 200, 176]


def test_indice_key_cache_and_sequential(oracle):
    rng = np.random.default_rng(3)
    feats, idx = random_sparse(rng, (6, 8, 8), 1, 60, 4)
    net = oracle.SparseSequential(
        oracle.SubMConv3d(4, 8, 3, bias=False, indice_key="a"), torch.nn.BatchNorm1d(8), torch.nn.ReLU(),
        oracle.SubMConv3d(8, 8, 3, bias=False, indice_key="a"), torch.nn.BatchNorm1d(8), torch.nn.ReLU(),
        oracle.SparseConv3d(8, 8, 3, 2, padding=1, bias=False)).eval()
    assert [k for k, _ in net.named_children()] == [str(i) for i in range(7)]
    x = oracle.SparseConvTensor(feats, idx, (6, 8, 8), 1)
    with torch.no_grad():
        y = net(x)
    assert set(y.indice_dict.keys()) == {"a", None}
    assert y.dense().shape == (1, 8, 3, 4, 4)
    sd = net.state_dict()
    assert sd["0.weight"].shape == (3, 3, 3, 4, 8)


def test_dense_layout(oracle):
    feats = torch.arange(6, dtype=torch.float32).view(2, 3)
    idx = torch.tensor([[0, 1, 2, 3], [1, 0, 0, 1]], dtype=torch.int32)
    d = oracle.SparseConvTensor(feats, idx, [2, 3, 4], 2).dense()
    assert d.shape == (2, 3, 2, 3, 4)
    assert d[0, :, 1, 2, 3].tolist() == [0, 1, 2] and d[1, :, 0, 0, 1].tolist() == [3, 4, 5]
    assert float(d.abs().sum()) == 15.0
