"""Box decode + NMS call-through points of the SECOND predict step (host side, torch/numpy).

These are the *module-by-module* (reference-faithful) versions: they do what
second/pytorch/core/box_torch_ops.py and second/core/non_max_suppression/{nms_cpu,nms_gpu}.py do,
going through the ``spconv.utils`` boundary functions of whichever ``spconv`` backend is passed in.
The fused device path (decode+filter+top-k+NMS in one stream of kernels, no host sync) lives in
``engine.py``.

Reference anchors:
  second_box_decode        second/pytorch/core/box_torch_ops.py:56-102
  limit_period             box_torch_ops.py:370-371
  rotate_nms / nms         box_torch_ops.py:454-477,492-515
  rotate_nms_cc            second/core/non_max_suppression/nms_cpu.py:20-31
  nms_gpu_cc               second/core/non_max_suppression/nms_gpu.py:10-19
  center_to_corner_box2d   second/core/box_np_ops.py:405-425 (corner order :195-199, rotation :344-357)
  corner_to_standup_nd     box_np_ops.py:278-283 ; iou_jit box_np_ops.py:696-725
"""
import numpy as np
import torch


def second_box_decode(box_encodings, anchors):
    """[..., 7(+c)] residuals + anchors -> boxes (x, y, z, w, l, h, r, *custom)."""
    xa, ya, za, wa, la, ha, ra = (anchors[..., i:i + 1] for i in range(7))
    xt, yt, zt, wt, lt, ht, rt = (box_encodings[..., i:i + 1] for i in range(7))
    diagonal = torch.sqrt(la ** 2 + wa ** 2)
    xg = xt * diagonal + xa
    yg = yt * diagonal + ya
    zg = zt * ha + za
    lg = torch.exp(lt) * la
    wg = torch.exp(wt) * wa
    hg = torch.exp(ht) * ha
    rg = rt + ra
    extra = [box_encodings[..., i:i + 1] + anchors[..., i:i + 1] for i in range(7, anchors.shape[-1])]
    return torch.cat([xg, yg, zg, wg, lg, hg, rg, *extra], dim=-1)


def limit_period(val, offset=0.5, period=np.pi):
    return val - torch.floor(val / period + offset) * period


_CORNER_SIGNS = np.array([[-0.5, -0.5], [-0.5, 0.5], [0.5, 0.5], [0.5, -0.5]], dtype=np.float32)


def corners_2d_np(centers, dims, angles):
    """numpy [N,2],[N,2],[N] -> [N,4,2]; clockwise corners from the min corner, rotated clockwise
    for positive angle (``p @ [[c,-s],[s,c]]``)."""
    dt = dims.dtype
    corners = dims.reshape(-1, 1, 2) * _CORNER_SIGNS.astype(dt).reshape(1, 4, 2)
    s, c = np.sin(angles), np.cos(angles)
    x, y = corners[..., 0], corners[..., 1]
    out = np.stack([x * c[:, None] + y * s[:, None], -x * s[:, None] + y * c[:, None]], axis=-1)
    return (out + centers.reshape(-1, 1, 2)).astype(dt)


def corners_2d_torch(centers, dims, angles):
    signs = torch.from_numpy(_CORNER_SIGNS).to(dims)
    corners = dims.view(-1, 1, 2) * signs.view(1, 4, 2)
    s, c = torch.sin(angles), torch.cos(angles)
    x, y = corners[..., 0], corners[..., 1]
    out = torch.stack([x * c[:, None] + y * s[:, None], -x * s[:, None] + y * c[:, None]], dim=-1)
    return out + centers.view(-1, 1, 2)


def standup_np(corners):
    return np.concatenate([corners.min(axis=1), corners.max(axis=1)], axis=-1)


def standup_torch(corners):
    return torch.cat([corners.min(dim=1)[0], corners.max(dim=1)[0]], dim=1)


def standup_iou_np(boxes, eps=0.0):
    """pairwise IoU of axis-aligned boxes [N,4] (iou_jit semantics: 0 unless iw>0 and ih>0)."""
    b = boxes.astype(np.float32)
    area = (b[:, 2] - b[:, 0] + eps) * (b[:, 3] - b[:, 1] + eps)
    iw = np.minimum(b[:, None, 2], b[None, :, 2]) - np.maximum(b[:, None, 0], b[None, :, 0]) + eps
    ih = np.minimum(b[:, None, 3], b[None, :, 3]) - np.maximum(b[:, None, 1], b[None, :, 1]) + eps
    ok = (iw > 0) & (ih > 0)
    inter = iw * ih
    ua = area[:, None] + area[None, :] - inter
    out = np.zeros_like(inter, dtype=np.float32)
    np.divide(inter, ua, out=out, where=ok)
    out[~ok] = 0
    return out


def _topk_desc(scores, k):
    """deterministic top-k: descending score, ties -> lower index first (stable)."""
    order = torch.sort(scores, descending=True, stable=True)[1]
    return order[:k]


def rotate_nms(backend, rbboxes, scores, pre_max_size, post_max_size, iou_threshold):
    """rbboxes [N,5] (x,y,w,l,r) torch; returns LongTensor of kept indices into the input."""
    indices = _topk_desc(scores, min(scores.shape[0], pre_max_size))
    rb = rbboxes[indices]
    sc = scores[indices]
    dets = torch.cat([rb, sc.unsqueeze(-1)], dim=1).detach().cpu().numpy()
    if dets.shape[0] == 0:
        return torch.zeros([0], dtype=torch.long, device=rbboxes.device)
    order = np.arange(dets.shape[0], dtype=np.int32)  # already sorted by score
    corners = corners_2d_np(dets[:, :2], dets[:, 2:4], dets[:, 4])
    st = standup_np(corners)
    siou = standup_iou_np(st, eps=0.0)
    keep = backend.utils.rotate_non_max_suppression_cpu(corners, order, siou, iou_threshold)
    keep = np.array(keep, dtype=np.int64)[:post_max_size]
    if keep.shape[0] == 0:
        return torch.zeros([0], dtype=torch.long, device=rbboxes.device)
    return indices[torch.from_numpy(keep).to(rbboxes.device)]


def aligned_nms(backend, bboxes, scores, pre_max_size, post_max_size, iou_threshold):
    """bboxes [N,4] standup boxes torch; '+1' IoU, '>' test (nms_gpu_cc)."""
    indices = _topk_desc(scores, min(scores.shape[0], pre_max_size))
    bb = bboxes[indices]
    sc = scores[indices]
    dets = torch.cat([bb, sc.unsqueeze(-1)], dim=1).detach().cpu().numpy().astype(np.float32)
    if dets.shape[0] == 0:
        return torch.zeros([0], dtype=torch.long, device=bboxes.device)
    keep = np.zeros(dets.shape[0], dtype=np.int32)
    num = backend.utils.non_max_suppression(np.ascontiguousarray(dets), keep, iou_threshold, 0)
    keep = keep[:num].astype(np.int64)[:post_max_size]
    if keep.shape[0] == 0:
        return torch.zeros([0], dtype=torch.long, device=bboxes.device)
    return indices[torch.from_numpy(keep).to(bboxes.device)]
