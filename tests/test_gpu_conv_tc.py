"""GPU parity of the tensor-core RPN convolution (b2s_conv2d_tc, tcgen05 + 3xTF32 split) against torch fp32
conv2d with TF32 disabled.  Bar: fp32-grade accuracy, |err| <= 2e-5 * max|ref| (plain TF32 would be ~1e-3)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def pad_nhwc(x):
    """[B,C,H,W] -> NHWC with one-pixel zero halo [B,H+2,W+2,C]."""
    return F.pad(x.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1)).contiguous()


def run_tc(product, x, w_tco_ci, taps, cout, n_pad, scale, shift, relu, out_padded, out_stride, want_lo=True):
    from b2second import tc
    L = product._lib
    lib = L.load()
    B, C, H, W = x.shape
    hi, lo = tc.split_tf32(pad_nhwc(x))
    w_hi, w_lo = tc.split_tf32(tc._pad_rows(w_tco_ci, n_pad))
    shape = (B, H + 2, W + 2, out_stride) if out_padded else (B, H, W, out_stride)
    o_hi = torch.zeros(shape, device="cuda")
    o_lo = torch.zeros(shape, device="cuda") if want_lo else None
    L.check(lib.b2s_conv2d_tc(L.ptr(hi), L.ptr(lo), B, H, W, C, L.ptr(w_hi), L.ptr(w_lo), taps, cout, n_pad,
                              L.ptr(scale), L.ptr(shift), 1 if relu else 0, L.ptr(o_hi), L.ptr(o_lo),
                              1 if out_padded else 0, out_stride, L.stream()), "b2s_conv2d_tc")
    torch.cuda.synchronize()
    return o_hi, o_lo


@pytest.fixture(autouse=True)
def _fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


@pytest.mark.timeout(120)
@pytest.mark.parametrize("B,H,W,cin,cout", [(1, 8, 16, 32, 128), (2, 24, 40, 128, 128), (1, 200, 176, 128, 128),
                                            (1, 16, 32, 64, 64)])
def test_conv3x3_bn_relu(product, B, H, W, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H)
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * (2.0 / (9 * cin)) ** 0.5
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    shift = torch.randn(cout, device="cuda", generator=g) * 0.1
    ref = torch.relu(F.conv2d(x, w, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    wt = w.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous()
    n_pad = 128 if cout > 64 else 64
    o_hi, o_lo = run_tc(product, x, wt, 9, cout, n_pad, scale, shift, True, True, cout)
    got = (o_hi + o_lo)[:, 1:-1, 1:-1, :].permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), "max err %g (ref max %g)" % (err, ref.abs().max().item())
    # halo untouched (stays zero) and hi is exactly tf32-representable
    assert float(o_hi[:, 0].abs().sum() + o_hi[:, -1].abs().sum() + o_hi[:, :, 0].abs().sum()
                 + o_hi[:, :, -1].abs().sum()) == 0.0
    assert int((o_hi.view(torch.int32) & 0x1FFF).abs().sum()) == 0


@pytest.mark.timeout(120)
def test_conv1x1_heads_packed(product):
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W, cin, cout = 2, 20, 33, 128, 20
    x = torch.randn(B, cin, H, W, device="cuda", generator=g)
    w = torch.randn(cout, cin, 1, 1, device="cuda", generator=g) * 0.1
    bias = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv2d(x, w, bias)
    wt = w[:, :, 0, 0].unsqueeze(0).contiguous()
    o, _ = run_tc(product, x, wt, 1, cout, 32, None, bias, False, False, 32, want_lo=False)
    got = o[..., :cout].permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), "max err %g" % err
    assert float(o[..., cout:].abs().sum()) == 0.0


@pytest.mark.timeout(180)
def test_rpn_stack_matches_torch(product):
    """whole car.fhd RPN (6 blocks + deblock + heads) chained through hi/lo planes vs the torch modules."""
    from b2second import config, models, tc
    L = product._lib
    lib = L.load()
    net = models.build_network(config.get_config("car.fhd"), product).eval()
    models.synthetic_weights_(net, "car.fhd", seed=0)
    rpn = net.rpn.cuda()
    assert tc.supported(rpn)
    B, H, W = 2, 40, 48
    x = torch.relu(torch.randn(B, 128, H, W, device="cuda"))
    x = x * (torch.rand(B, 1, H, W, device="cuda") < 0.15)        # sparse BEV-like input
    with torch.no_grad():
        feat = rpn.backbone(x)
        ref = torch.cat([rpn.conv_box(feat), rpn.conv_cls(feat), rpn.conv_dir_cls(feat)], 1)
    plan = tc.plan_rpn(rpn)
    hi, lo = tc.split_tf32(pad_nhwc(x))
    for lyr in plan[:-1]:
        o_hi = torch.zeros(B, H + 2, W + 2, lyr["cout"], device="cuda")
        o_lo = torch.zeros_like(o_hi)
        L.check(lib.b2s_conv2d_tc(L.ptr(hi), L.ptr(lo), B, H, W, lyr["cin"], L.ptr(lyr["w_hi"]), L.ptr(lyr["w_lo"]),
                                  lyr["taps"], lyr["cout"], lyr["n_pad"], L.ptr(lyr["scale"]), L.ptr(lyr["shift"]),
                                  1, L.ptr(o_hi), L.ptr(o_lo), 1, lyr["cout"], L.stream()), "b2s_conv2d_tc")
        hi, lo = o_hi, o_lo
    lyr = plan[-1]
    out = torch.zeros(B, H, W, 32, device="cuda")
    L.check(lib.b2s_conv2d_tc(L.ptr(hi), L.ptr(lo), B, H, W, lyr["cin"], L.ptr(lyr["w_hi"]), L.ptr(lyr["w_lo"]), 1,
                              lyr["cout"], lyr["n_pad"], None, L.ptr(lyr["shift"]), 0, L.ptr(out), None, 0, 32,
                              L.stream()), "b2s_conv2d_tc")
    torch.cuda.synchronize()
    got = out[..., :20].permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    assert err <= 1e-4, "head tensors differ by %g (bar 1e-4 on the regression outputs)" % err
    assert lyr["head_offsets"] == [0, 14, 16, 20]
