"""Frame sharding over GPUs: one process per GPU, no data-path collective inside a frame, ONE all-gather
of the fixed-stride detection records per batch.

The reference only has single-process ``torch.nn.DataParallel`` for training
(second/pytorch/train.py:203-206) and evaluates on one GPU (train.py:271-277); sharding inference
frames is new functionality, so there is no reference call site to mirror beyond the per-frame
contract of ``VoxelNet.forward``.

Layout: global frame ``g`` of a batch of ``W*B`` frames runs on rank ``g // B`` as local frame
``g % B`` (contiguous blocks, so rank-major gather order == global frame order).
Record per frame: ``det [post_max, code+2]`` (box, score, label) + ``count`` -> packed as one float32
tensor ``[B, post_max*(code+2) + 1]`` so a single collective moves everything.
"""
import torch
import torch.distributed as dist


def frames_for_rank(num_frames, rank, world):
    """contiguous block partition; returns (start, stop) global frame ids for ``rank``."""
    per = (num_frames + world - 1) // world
    start = min(rank * per, num_frames)
    return start, min(start + per, num_frames)


def pack_records(det, count):
    """det [B, post_max, S] f32, count [B] i32 -> [B, post_max*S + 1] f32 (count stored exactly: < 2^24)."""
    B = det.shape[0]
    return torch.cat([det.reshape(B, -1), count.to(det.dtype).view(B, 1)], dim=1).contiguous()


def unpack_records(rec, post_max, stride):
    det = rec[:, :post_max * stride].reshape(rec.shape[0], post_max, stride)
    count = rec[:, post_max * stride].round().to(torch.int32)
    return det, count


class DetectionGatherer:
    """pre-allocated send/recv buffers + one ``all_gather_into_tensor`` per batch (NCCL on GPUs, gloo in tests)."""

    def __init__(self, batch_per_rank, post_max, stride, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.B, self.post_max, self.stride = batch_per_rank, post_max, stride
        width = post_max * stride + 1
        self.send = torch.zeros(batch_per_rank, width, dtype=torch.float32, device=device)
        self.recv = torch.zeros(self.world * batch_per_rank, width, dtype=torch.float32, device=device)

    def gather_records(self, records):
        """records [B, post_max*S + 1]: the engine's ``det_record`` buffer (the NMS epilogue writes the detections AND
        the count straight into it, so there is no packing step) -> same return value as ``gather``.  The buffer is
        the collective's send buffer itself."""
        assert records.shape == self.send.shape and records.is_contiguous()
        if self.world == 1:
            self.recv.copy_(records)
        elif hasattr(dist, "all_gather_into_tensor") and records.is_cuda:
            dist.all_gather_into_tensor(self.recv, records, group=self.group)
        else:
            parts = list(self.recv.view(self.world, self.B, -1).unbind(0))
            dist.all_gather(parts, records, group=self.group)
        return unpack_records(self.recv, self.post_max, self.stride)

    def gather(self, det, count):
        """-> (det_all [W*B, post_max, S], count_all [W*B]) in global frame order, on every rank."""
        self.send.copy_(pack_records(det, count))
        if self.world == 1:
            self.recv.copy_(self.send)
        elif hasattr(dist, "all_gather_into_tensor") and self.send.is_cuda:
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        else:
            parts = list(self.recv.view(self.world, self.B, -1).unbind(0))
            dist.all_gather(parts, self.send, group=self.group)
        return unpack_records(self.recv, self.post_max, self.stride)
