"""Host side of the tensor-core RPN (csrc/conv_tc.cu, b2s_conv2d_tc): weight preparation and the layer plan.

RPNV2 for the sparse-conv configs (second/pytorch/models/rpn.py:467-497 with layer_strides [1], upsample
strides [1]) is: 6 x [Conv3x3 pad1 + BN + ReLU] -> deblock ConvTranspose2d(k=1,s=1) + BN + ReLU -> three 1x1
heads.  Each becomes one b2s_conv2d_tc launch on NHWC halo-padded hi/lo planes; the three heads are one launch
writing a packed 32-float record per pixel (box | cls | dir | pad) that b2s_decode_filter_strided reads.
"""
import numpy as np
import torch
from torch import nn


def split_tf32(t):
    """fp32 tensor -> (hi, lo): hi = value rounded to tf32 (10-bit mantissa, round half away), lo = t - hi (exact)."""
    t = t.detach().float().contiguous()

    def rn(x):
        r = ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        return torch.where(torch.isfinite(r), r, x)

    hi = rn(t)
    # lo is ALSO rounded to tf32: the tensor core then sees exactly representable operands, so no hardware
    # truncation (a biased error that accumulates ~K instead of ~sqrt(K)) can occur; |t - hi - lo| <= 2^-23 |t|.
    lo = rn(t - hi)
    return hi.contiguous(), lo.contiguous()


SPARSE_TC_CIN = (4, 16, 32, 64)
SPARSE_TC_COUT = (16, 32, 64)


def pack_sparse_weights(w):
    """spconv weight [K, Cin, Cout] -> the K-major B operand b2s_sparse_conv_tc expects.
    Cin >= 32: [K, Cout, Cin].  Cin < 32 (4 or 16): PACK = 32/Cin kernel offsets share one 128-byte K block, so the
    rows are packed [ceil(K/PACK), Cout, 32] with column (offset-in-pack * Cin + cin) and zero columns past K."""
    K, cin, cout = w.shape
    wt = w.detach().float().transpose(1, 2).contiguous()            # [K, Cout, Cin]
    if cin >= 32:
        return wt
    pack = 32 // cin
    nkb = (K + pack - 1) // pack
    out = torch.zeros(nkb * pack, cout, cin, dtype=wt.dtype, device=wt.device)
    out[:K] = wt
    return out.view(nkb, pack, cout, cin).permute(0, 2, 1, 3).reshape(nkb, cout, 32).contiguous()


def _fold_bn2d(bn):
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
    shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
    return scale, shift


def _pad_rows(w, n_pad):
    """[taps, Cout, Cin] -> [taps, n_pad, Cin] (zero rows)."""
    taps, cout, cin = w.shape
    if cout == n_pad:
        return w.contiguous()
    out = torch.zeros(taps, n_pad, cin, dtype=w.dtype, device=w.device)
    out[:, :cout] = w
    return out.contiguous()


def _n_pad(cout):
    for n in (32, 64, 128):
        if cout <= n:
            return n
    raise ValueError("Cout %d > 128 not supported by the tensor-core RPN" % cout)


def supported(rpn):
    """True when every layer of this RPNV2 maps onto b2s_conv2d_tc (stride-1 3x3 / 1x1, channels % 32 == 0)."""
    try:
        plan_rpn(rpn, dry=True)
        return True
    except (ValueError, AssertionError):
        return False


def plan_rpn(rpn, dry=False):
    """-> list of layer dicts {taps, cin, cout, n_pad, w_hi, w_lo, scale, shift, relu, kind}."""
    if len(rpn.blocks) != 1 or len(rpn.deblocks) != 1:
        raise ValueError("multi-stage RPN (strided blocks / upsampling deblocks) stays on cuDNN this round")
    layers = []

    def add(w_tco_ci, scale, shift, relu, kind, taps):
        taps_, cout, cin = w_tco_ci.shape
        assert taps_ == taps and cin % 32 == 0, "channels must be multiples of 32"
        n_pad = _n_pad(cout)
        d = {"taps": taps, "cin": cin, "cout": cout, "n_pad": n_pad, "relu": relu, "kind": kind,
             "scale": scale, "shift": shift}
        if not dry:
            d["w_hi"], d["w_lo"] = split_tf32(_pad_rows(w_tco_ci.float(), n_pad))
        layers.append(d)

    mods = list(rpn.blocks[0])
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.ZeroPad2d):
            assert tuple(m.padding) == (1, 1, 1, 1)
            conv, bn, relu = mods[i + 1], mods[i + 2], mods[i + 3]
            assert conv.padding == (0, 0)
            i += 4
        else:
            conv, bn, relu = mods[i], mods[i + 1], mods[i + 2]
            assert conv.padding == (1, 1)
            i += 3
        assert isinstance(conv, nn.Conv2d) and conv.kernel_size == (3, 3) and conv.stride == (1, 1) and conv.bias is None
        assert isinstance(bn, nn.BatchNorm2d) and isinstance(relu, nn.ReLU)
        w = conv.weight.detach()                                   # [Cout, Cin, 3, 3]
        w = w.permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1])   # [tap=ky*3+kx, Cout, Cin]
        s, b = _fold_bn2d(bn)
        add(w, s, b, True, "block", 9)
    up, bn, relu = list(rpn.deblocks[0])
    if isinstance(up, nn.ConvTranspose2d):
        assert up.kernel_size == (1, 1) and up.stride == (1, 1) and up.bias is None
        w = up.weight.detach()[:, :, 0, 0].t().unsqueeze(0)        # [Cin, Cout,1,1] -> [1, Cout, Cin]
    else:
        assert up.kernel_size == (1, 1) and up.stride == (1, 1) and up.bias is None
        w = up.weight.detach()[:, :, 0, 0].unsqueeze(0)
    s, b = _fold_bn2d(bn)
    add(w.contiguous(), s, b, True, "deblock", 1)
    # heads: box | cls | dir packed into one [1, n, Cin] matrix, bias as shift
    heads = [rpn.conv_box, rpn.conv_cls] + ([rpn.conv_dir_cls] if rpn._use_direction_classifier else [])
    w = torch.cat([h.weight.detach()[:, :, 0, 0] for h in heads], 0).unsqueeze(0)     # [1, sum Cout, Cin]
    bias = torch.cat([h.bias.detach() for h in heads], 0).float().contiguous()
    pad = (-w.shape[1]) % 4
    if pad:
        w = torch.cat([w, torch.zeros(1, pad, w.shape[2], device=w.device, dtype=w.dtype)], 1)
        bias = torch.cat([bias, torch.zeros(pad, device=bias.device)])
    add(w.contiguous(), None, bias, False, "heads", 1)
    offs = np.cumsum([0] + [h.weight.shape[0] for h in heads]).tolist()
    layers[-1]["head_offsets"] = offs                       # box at offs[0], cls at offs[1], dir at offs[2]
    return layers
