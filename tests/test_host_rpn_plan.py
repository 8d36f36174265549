"""Host logic of the tensor-core RPN: b2second.tc.plan_rpn turns every BASELINE config's RPNV2 (stride-1/2 blocks,
k = 1/2/4 ConvTranspose2d or strided-Conv2d deblocks, channel concat, 1x1 heads; rpn.py:264-299,469-497) into a
program of b2s_conv2d_tc_ex launches.  Here the program is INTERPRETED on the CPU with torch.conv2d using exactly
the op fields the CUDA entry point receives (kernel, stride, pad, GEMM grid, output sub-grid, channel offset), and
compared with the torch modules -- so the planner and the op semantics are pinned without a GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from b2second import config, loader, models, tc


def interpret(plan, x):
    """x [B, C, H, W] fp32 -> packed heads [B, S, Hh, Wh], executing the ops with plain torch."""
    B = x.shape[0]
    bufs = {"in": x}
    for name, (h, w, c) in plan["buffers"].items():
        bufs[name] = torch.zeros(B, c, h, w)
    hd = plan["heads"]
    bufs["heads"] = torch.zeros(B, hd["stride"], hd["H"], hd["W"])
    for op in plan["ops"]:
        src = bufs[op["src"]]
        assert src.shape[1] == op["cin"] and tuple(src.shape[2:]) == (op["Hin"], op["Win"])
        # fp16 hi/lo planes of the power-of-two pre-scaled weights; op["scale"] carries the inverse scale
        w = (op["w_hi"].float() + op["w_lo"].float())[:, :op["cout"]]      # [taps, cout, cin]
        w = w.reshape(op["kh"], op["kw"], op["cout"], op["cin"]).permute(2, 3, 0, 1)
        y = F.conv2d(src, w, stride=op["stride"], padding=op["pad"])
        assert tuple(y.shape[2:]) == (op["Hg"], op["Wg"])
        if op["scale"] is not None:
            y = y * op["scale"].view(1, -1, 1, 1)
        if op["shift"] is not None:
            y = y + op["shift"].view(1, -1, 1, 1)
        if op["relu"]:
            y = torch.relu(y)
        dst = bufs[op["dst"]]
        assert tuple(dst.shape[2:]) == (op["Hout"], op["Wout"])
        c0 = op["dst_coff"]
        dst[:, c0:c0 + op["cout"], op["off_h"]::op["out_mul"], op["off_w"]::op["out_mul"]] = y
    return bufs["heads"]


@pytest.mark.parametrize("name", list(config.BUILTIN))
def test_rpn_program_semantics_on_cpu(name):
    sp = loader.oracle_spconv()
    cfg = config.get_config(name)
    net = models.build_network(cfg, sp).eval()
    models.synthetic_weights_(net, name, seed=0)
    rpn = net.rpn
    assert tc.supported(rpn)
    H, W = 32, 48
    cin = rpn.blocks[0][1].in_channels
    torch.manual_seed(1)
    x = torch.relu(torch.randn(2, cin, H, W)) * (torch.rand(2, 1, H, W) < 0.2)
    with torch.no_grad():
        feat = rpn.backbone(x)
        heads = [rpn.conv_box(feat), rpn.conv_cls(feat)] + ([rpn.conv_dir_cls(feat)] if rpn._use_direction_classifier else [])
        ref = torch.cat(heads, 1)
        plan = tc.plan_rpn(rpn, H, W)
        got = interpret(plan, x)
    n = ref.shape[1]
    assert tuple(got.shape[2:]) == tuple(ref.shape[2:])
    err = float((got[:, :n] - ref).abs().max())
    assert err <= 2e-5 * max(1.0, float(ref.abs().max())), err
    assert plan["heads"]["offsets"][-1] == n


@pytest.mark.parametrize("name", list(config.BUILTIN))
def test_rpn_program_structure(name):
    sp = loader.oracle_spconv()
    cfg = config.get_config(name)
    net = models.build_network(cfg, sp).eval()
    _, fH, fW = cfg.feature_map_size
    # the BEV map the RPN sees: feature map x the config's downsample factor
    f = cfg.rpn_downsample_factor if hasattr(cfg, "rpn_downsample_factor") else None
    H = int(round(fH * (f if f else np.prod(cfg.rpn_layer_strides) / (cfg.rpn_upsample_strides[-1] if cfg.rpn_upsample_strides else 1))))
    W = int(round(fW * (f if f else np.prod(cfg.rpn_layer_strides) / (cfg.rpn_upsample_strides[-1] if cfg.rpn_upsample_strides else 1))))
    plan = tc.plan_rpn(net.rpn, H, W, dry=True)
    assert (plan["heads"]["H"], plan["heads"]["W"]) == (fH, fW)
    seen = {}
    for op in plan["ops"]:
        assert op["cin"] % 32 == 0 and op["n_pad"] in (32, 64, 128) and op["cout"] <= op["n_pad"] and op["cout"] % 4 == 0
        assert op["kh"] * op["kw"] == op["taps"] <= 16 and op["stride"] in (1, 2, 4) and op["pad"] in (0, 1)
        assert (op["Hg"] - 1) * op["out_mul"] + op["off_h"] < op["Hout"]
        assert (op["Wg"] - 1) * op["out_mul"] + op["off_w"] < op["Wout"]
        if op["kind"] == "deblock" and op["out_mul"] > 1:
            seen.setdefault((op["dst_coff"], op["out_mul"]), set()).add((op["off_h"], op["off_w"]))
    for (coff, s), offs in seen.items():
        assert offs == {(a, c) for a in range(s) for c in range(s)}, "ConvTranspose2d sub-grids must tile the output"
    # every channel of the concat buffer is written exactly once per sub-grid position
    if "cat" in plan["buffers"]:
        ccat = plan["buffers"]["cat"][2]
        cover = np.zeros(ccat, np.int64)
        for op in plan["ops"]:
            if op["dst"] == "cat" and (op["off_h"], op["off_w"]) == (0, 0):
                cover[op["dst_coff"]:op["dst_coff"] + op["cout"]] += 1
        assert (cover == 1).all()
