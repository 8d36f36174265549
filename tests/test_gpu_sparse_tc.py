"""GPU parity of the tensor-pipe sparse convolution (b2s_sparse_conv_tc, tcgen05 + 3xTF32) against the CPU oracle's
sparse conv (fp32).  Same rulebook (b2s_rulebook_*), same BN/ReLU epilogue; bar 2e-5 * max|ref| per layer."""
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


@pytest.mark.timeout(180)
@pytest.mark.parametrize("cin,cout", [(32, 32), (32, 64), (64, 64), (64, 32), (4, 16), (16, 16), (16, 32)])
@pytest.mark.parametrize("subm", [True, False])
@pytest.mark.parametrize("n", [5000, 100, 0])
def test_sparse_conv_tc_matches_oracle(product, oracle, cin, cout, subm, n):
    from b2second import tc
    L = product._lib
    lib = L.load()
    rng = np.random.default_rng(cin * 7 + cout + subm + n)
    shape, batch = (11, 40, 36), 2
    feats, idx = random_sparse(rng, shape, batch, n, cin)
    if subm:
        oc = oracle.SubMConv3d(cin, cout, 3, bias=False, indice_key="k")
        k, s, p = [3, 3, 3], [1, 1, 1], [1, 1, 1]
    else:
        oc = oracle.SparseConv3d(cin, cout, 3, 2, padding=[0, 1, 1], bias=False)
        k, s, p = [3, 3, 3], [2, 2, 2], [0, 1, 1]
    scale = torch.rand(cout) + 0.5
    shift = torch.randn(cout) * 0.1
    with torch.no_grad():
        yo = oc(oracle.SparseConvTensor(feats, idx, shape, batch))
        ref = torch.relu(yo.features * scale + shift)
    x = product.SparseConvTensor(feats.cuda(), idx.cuda(), shape, batch)
    rb = product.ops.build_rulebook(x, k, s, p, [1, 1, 1], subm)
    assert rb.num_out == yo.features.shape[0]
    if rb.num_out == 0:
        return
    w = oc.weight.detach().view(27, cin, cout).cuda()
    w_hi, w_lo = tc.split_tf32(tc.pack_sparse_weights(w))   # [K,Cout,Cin], or packed K blocks for Cin < 32
    f_hi, f_lo = tc.split_tf32(feats.cuda())
    o_hi = torch.zeros(rb.num_out, cout, device="cuda")
    o_lo = torch.zeros_like(o_hi)
    scale_d, shift_d, nbr = scale.cuda(), shift.cuda(), rb.nbr.contiguous()   # keep alive across the async launch
    L.check(lib.b2s_sparse_conv_tc(L.ptr(f_hi), L.ptr(f_lo), f_hi.shape[0], cin, L.ptr(w_hi), L.ptr(w_lo), L.ptr(nbr),
                                   27, L.ptr(rb.num_out_dev), rb.num_out, L.ptr(scale_d), L.ptr(shift_d), 1,
                                   L.ptr(o_hi), L.ptr(o_lo), cout, L.stream()), "b2s_sparse_conv_tc")
    torch.cuda.synchronize()
    got = (o_hi + o_lo).cpu()
    err = (got - ref).abs().max().item()
    assert err <= 2e-5 * max(ref.abs().max().item(), 1.0), "max err %g (ref max %g)" % (err, ref.abs().max().item())
    # split / merge helpers round trip
    m = torch.zeros_like(o_hi)
    L.check(lib.b2s_merge_hilo(L.ptr(o_hi), L.ptr(o_lo), L.ptr(m), L.ptr(rb.num_out_dev), rb.num_out, cout, L.stream()),
            "b2s_merge_hilo")
    h2, l2 = torch.zeros_like(o_hi), torch.zeros_like(o_hi)
    L.check(lib.b2s_split_tf32(L.ptr(m), L.ptr(h2), L.ptr(l2), L.ptr(rb.num_out_dev), rb.num_out, cout, L.stream()),
            "b2s_split_tf32")
    torch.cuda.synchronize()
    # merge is exact; re-splitting may pick the neighbouring tf32 value for hi at rounding ties, but the pair
    # still sums to the same fp32 value within the dropped 2^-23 tail
    assert torch.equal(m, o_hi + o_lo)
    assert float((h2 + l2 - m).abs().max()) <= 2e-7 * float(m.abs().max())
    assert int((h2.view(torch.int32) & 0x1FFF).abs().sum()) == 0 and int((l2.view(torch.int32) & 0x1FFF).abs().sum()) == 0
