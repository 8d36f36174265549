"""GPU parity: rulebooks (indices bit-exact) and sparse convolution (fp32, tol 1e-5 rel / 1e-5 abs per layer)
through the spconv drop-in modules, against the CPU oracle modules with identical weights."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def random_sparse(rng, shape, batch, n, cin):
    D, H, W = shape
    flat = rng.choice(batch * D * H * W, size=n, replace=False)
    b, r = np.divmod(flat, D * H * W)
    z, r = np.divmod(r, H * W)
    y, x = np.divmod(r, W)
    idx = np.stack([b, z, y, x], 1).astype(np.int32)
    feats = rng.standard_normal((n, cin)).astype(np.float32)
    return torch.from_numpy(feats), torch.from_numpy(idx)


def pair_set(pairs, pair_num):
    pairs = pairs.cpu().numpy()
    pn = pair_num.cpu().numpy()
    out = []
    for k in range(pairs.shape[0]):
        p = pairs[k, :, :pn[k]]
        out.append(sorted(zip(p[0].tolist(), p[1].tolist())))
    return out


RB_CASES = [
    ((9, 12, 10), 2, 3, 1, 0, True, 0.15),
    ((9, 12, 10), 2, 3, 2, 1, False, 0.15),
    ((11, 12, 10), 3, 3, 2, [0, 1, 1], False, 0.2),
    ((5, 12, 10), 1, (3, 1, 1), (2, 1, 1), 0, False, 0.3),
    ((41, 160, 140), 2, 3, 2, 1, False, 0.01),
    ((41, 160, 140), 2, 3, 1, 0, True, 0.01),
    ((7, 9, 8), 1, 3, 1, 1, False, 0.2),
    ((6, 6, 6), 1, 3, 2, 1, False, 0.0),     # empty input
]


@pytest.mark.parametrize("shape,batch,k,s,p,subm,dens", RB_CASES)
def test_rulebook_matches_oracle(product, oracle, shape, batch, k, s, p, subm, dens):
    rng = np.random.default_rng(abs(hash((shape, subm, batch))) % 2**31)
    n = int(dens * batch * np.prod(shape))
    _, idx = random_sparse(rng, shape, batch, n, 1)
    o_out, o_pairs, o_num = oracle.ops.get_indice_pairs(idx, batch, shape, k, s, p, 1, 0, subm)
    g_out, g_pairs, g_num = product.ops.get_indice_pairs(idx.cuda(), batch, shape, k, s, p, 1, 0, subm)
    assert g_out.dtype == torch.int32
    np.testing.assert_array_equal(g_out.cpu().numpy(), o_out.numpy())         # same rows, same (sorted) order
    np.testing.assert_array_equal(g_num.cpu().numpy(), o_num.numpy())
    assert pair_set(g_pairs, g_num) == pair_set(o_pairs, o_num)


CONV_CASES = [(4, 16), (3, 16), (16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (5, 7)]


@pytest.mark.parametrize("cin,cout", CONV_CASES)
@pytest.mark.parametrize("subm", [True, False])
def test_sparse_conv_matches_oracle(product, oracle, cin, cout, subm):
    rng = np.random.default_rng(cin * 100 + cout + subm)
    shape, batch = (11, 40, 36), 2
    n = 3000
    feats, idx = random_sparse(rng, shape, batch, n, cin)
    if subm:
        oc = oracle.SubMConv3d(cin, cout, 3, bias=True, indice_key="k")
        gc = product.SubMConv3d(cin, cout, 3, bias=True, indice_key="k")
    else:
        oc = oracle.SparseConv3d(cin, cout, 3, 2, padding=[0, 1, 1], bias=True)
        gc = product.SparseConv3d(cin, cout, 3, 2, padding=[0, 1, 1], bias=True)
    gc.load_state_dict(oc.state_dict())
    gc = gc.cuda()
    with torch.no_grad():
        yo = oc(oracle.SparseConvTensor(feats, idx, shape, batch))
        yg = gc(product.SparseConvTensor(feats.cuda(), idx.cuda(), shape, batch))
    assert yg.spatial_shape == yo.spatial_shape
    np.testing.assert_array_equal(yg.indices.cpu().numpy(), yo.indices.numpy())
    torch.testing.assert_close(yg.features.cpu(), yo.features, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(yg.dense().cpu(), yo.dense(), rtol=1e-5, atol=1e-5)


def test_sequential_fused_bn_relu_matches_oracle(product, oracle):
    """the SpMiddleFHD layer pattern: (conv, BN1d, ReLU)* with shared indice_key, eval mode, fused epilogue."""
    rng = np.random.default_rng(42)
    shape, batch = (21, 80, 72), 2
    feats, idx = random_sparse(rng, shape, batch, 6000, 4)

    def build(sp):
        bn = lambda c: torch.nn.BatchNorm1d(c, eps=1e-3, momentum=0.01)
        return sp.SparseSequential(
            sp.SubMConv3d(4, 16, 3, bias=False, indice_key="subm0"), bn(16), torch.nn.ReLU(),
            sp.SubMConv3d(16, 16, 3, bias=False, indice_key="subm0"), bn(16), torch.nn.ReLU(),
            sp.SparseConv3d(16, 32, 3, 2, padding=1, bias=False), bn(32), torch.nn.ReLU(),
            sp.SubMConv3d(32, 32, 3, bias=False, indice_key="subm1"), bn(32), torch.nn.ReLU(),
            sp.SparseConv3d(32, 64, 3, 2, padding=[0, 1, 1], bias=False), bn(64), torch.nn.ReLU(),
            sp.SubMConv3d(64, 64, 3, bias=False, indice_key="subm2"), bn(64), torch.nn.ReLU(),
            sp.SparseConv3d(64, 64, (3, 1, 1), (2, 1, 1), bias=False), bn(64), torch.nn.ReLU())

    torch.manual_seed(0)
    on = build(oracle)
    for m in on.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.1)
    on.eval()
    gn = build(product)
    gn.load_state_dict(on.state_dict())
    gn = gn.cuda().eval()
    with torch.no_grad():
        yo = on(oracle.SparseConvTensor(feats, idx, shape, batch))
        yg = gn(product.SparseConvTensor(feats.cuda(), idx.cuda(), shape, batch))
        gn.fuse_bn_relu = False
        yu = gn(product.SparseConvTensor(feats.cuda(), idx.cuda(), shape, batch))
    np.testing.assert_array_equal(yg.indices.cpu().numpy(), yo.indices.numpy())
    torch.testing.assert_close(yg.features.cpu(), yo.features, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(yu.features.cpu(), yo.features, rtol=1e-4, atol=1e-5)
    d = yg.dense()
    assert d.shape == yo.dense().shape
    torch.testing.assert_close(d.cpu(), yo.dense(), rtol=1e-4, atol=1e-5)
    assert set(yg.indice_dict.keys()) == set(yo.indice_dict.keys())


def test_empty_tensor(product):
    x = product.SparseConvTensor(torch.zeros(0, 4).cuda(), torch.zeros(0, 4, dtype=torch.int32).cuda(), [8, 8, 8], 1)
    with torch.no_grad():
        y = product.SubMConv3d(4, 16, 3, bias=False).cuda()(x)
        z = product.SparseConv3d(16, 16, 3, 2, padding=1, bias=False).cuda()(y)
    assert y.features.shape == (0, 16) and z.features.shape == (0, 16)
    assert float(z.dense().abs().sum()) == 0.0
