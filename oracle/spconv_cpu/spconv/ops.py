"""CPU ORACLE of ``spconv.ops`` (rulebook + native sparse convolution).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (spconv 1.x is external to the reference);
pinned by our own known-answer tests (dense ``F.conv3d`` equivalence, tests/test_oracle_conv.py).

Semantics restated ([EXT-RECALL]):
* output size (non-SubM) ``(in + 2p - d(k-1) - 1)//s + 1``; SubM: shape kept, stride 1, pad k//2
  -- anchored in-tree by the shape comments second/pytorch/models/middle.py:152-189
* cross-correlation: ``out[o] = sum_k W[k]^T in[o*s - p + k*d]``, kernel offset index row-major (kz,ky,kx)
* output rows of a strided conv sorted ascending by flat (b,z,y,x) key (upstream GPU path)
* conv (ConvAlgo.Native): SubM centre offset first, then k ascending: gather -> mm -> scatter-add
"""
import numpy as np
import torch


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    out = []
    for i in range(len(input_size)):
        size = (input_size[i] + 2 * padding[i] - dilation[i] * (kernel_size[i] - 1) - 1) // stride[i] + 1
        out.append(int(size))
    return out


def _flat(idx, shape):
    """idx [N,4] (b,z,y,x) int64 -> flat key."""
    D, H, W = (int(s) for s in shape)
    return ((idx[:, 0] * D + idx[:, 1]) * H + idx[:, 2]) * W + idx[:, 3]


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1,
                     out_padding=0, subm=False, transpose=False, grid=None, use_hash=False):
    """Returns (out_indices [M,4] i32, indice_pairs [K,2,L] i32 (-1 filled), indice_pair_num [K] i32).

    L = number of input rows (upstream allocates [K,2,N_in])."""
    assert not transpose, "transposed/inverse sparse conv is not on the SECOND inference path"
    ndim = 3
    if not isinstance(ksize, (list, tuple)):
        ksize = [ksize] * ndim
    if not isinstance(stride, (list, tuple)):
        stride = [stride] * ndim
    if not isinstance(padding, (list, tuple)):
        padding = [padding] * ndim
    if not isinstance(dilation, (list, tuple)):
        dilation = [dilation] * ndim
    idx = indices.detach().cpu().numpy().astype(np.int64)
    n_in = idx.shape[0]
    spatial_shape = [int(s) for s in spatial_shape]
    if subm:
        out_shape = spatial_shape
        stride = [1] * ndim
        padding = [(k // 2) * d for k, d in zip(ksize, dilation)]
    else:
        out_shape = get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    K = int(np.prod(ksize))
    pairs = np.full((K, 2, n_in), -1, dtype=np.int32)
    pair_num = np.zeros((K,), dtype=np.int32)

    # candidate outputs per kernel offset: o = (i + p - k*d) / s
    cand = []  # (k, in_rows, out_coords[?,4])
    kk = 0
    for kz in range(ksize[0]):
        for ky in range(ksize[1]):
            for kx in range(ksize[2]):
                koff = np.array([kz * dilation[0], ky * dilation[1], kx * dilation[2]], dtype=np.int64)
                num = idx[:, 1:] + np.array(padding, dtype=np.int64) - koff
                s = np.array(stride, dtype=np.int64)
                ok = (num % s == 0).all(axis=1)
                o = num // s
                ok &= (o >= 0).all(axis=1) & (o < np.array(out_shape, dtype=np.int64)).all(axis=1)
                rows = np.nonzero(ok)[0]
                oc = np.concatenate([idx[rows, :1], o[rows]], axis=1)
                cand.append((kk, rows, oc))
                kk += 1

    if subm:
        out_idx = idx
        in_keys = _flat(idx, out_shape)
        order = np.argsort(in_keys, kind="stable")
        sorted_keys = in_keys[order]
        for k, rows, oc in cand:
            if rows.size == 0:
                continue
            okeys = _flat(oc, out_shape)
            pos = np.searchsorted(sorted_keys, okeys)
            pos = np.minimum(pos, sorted_keys.size - 1)
            hit = sorted_keys[pos] == okeys
            r_in = rows[hit]
            r_out = order[pos[hit]]
            n = r_in.size
            pairs[k, 0, :n] = r_in
            pairs[k, 1, :n] = r_out
            pair_num[k] = n
    else:
        all_keys = np.concatenate([_flat(oc, out_shape) for _, _, oc in cand]) if cand else np.zeros(0, np.int64)
        uniq = np.unique(all_keys)  # sorted ascending: upstream GPU path order
        D, H, W = out_shape
        out_idx = np.stack([uniq // (D * H * W), (uniq // (H * W)) % D, (uniq // W) % H, uniq % W], axis=1)
        for k, rows, oc in cand:
            if rows.size == 0:
                continue
            okeys = _flat(oc, out_shape)
            pos = np.searchsorted(uniq, okeys)
            n = rows.size
            pairs[k, 0, :n] = rows
            pairs[k, 1, :n] = pos
            pair_num[k] = n
    return (torch.from_numpy(out_idx.astype(np.int32)), torch.from_numpy(pairs), torch.from_numpy(pair_num))


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, inverse=False, subm=False):
    """features [N,Cin]; filters [kD,kH,kW,Cin,Cout]; returns [num_activate_out, Cout]."""
    assert not inverse
    cin, cout = filters.shape[-2], filters.shape[-1]
    w = filters.reshape(-1, cin, cout)
    K = w.shape[0]
    out = torch.zeros(num_activate_out, cout, dtype=features.dtype)
    if K == 1 and subm:
        return torch.mm(features, w[0])
    centre = K // 2
    if subm:
        out = torch.mm(features, w[centre])
    pn = indice_pair_num.tolist()
    for k in range(K):
        n = pn[k]
        if n <= 0 or (subm and k == centre):
            continue
        i_in = indice_pairs[k, 0, :n].long()
        i_out = indice_pairs[k, 1, :n].long()
        out.index_add_(0, i_out, torch.mm(features.index_select(0, i_in), w[k]))
    return out


def nms(boxes, scores, pre_max_size, post_max_size, thresh, eps):
    """spconv.ops.nms (used by box_torch_ops.nms_v2 only; not on the default path)."""
    from .utils import non_max_suppression_cpu
    scores_np = scores.detach().cpu().numpy()
    boxes_np = boxes.detach().cpu().numpy()
    order = np.argsort(-scores_np, kind="stable").astype(np.int32)
    if pre_max_size > 0:
        order = order[:pre_max_size]
    dets = np.concatenate([boxes_np, scores_np[:, None]], axis=1).astype(np.float32)
    keep = non_max_suppression_cpu(dets, order, thresh, eps)
    if post_max_size > 0:
        keep = keep[:post_max_size]
    return torch.tensor(keep, dtype=torch.long)
