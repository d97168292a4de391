"""Multi-GPU plumbing: images are independent on this path, so work is sharded by image and
there is NO collective on the hot path.  The only exchange is one all-gather of the final,
fixed-size detections buffer (NCCL over NVLink on the GPU box, gloo in the CPU tests).

One process per GPU (torchrun); rank r owns the contiguous image block `shard_range(...)`.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialises torch.distributed from RANK / WORLD_SIZE / MASTER_* when WORLD_SIZE > 1.
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def bind_to_gpu_numa(gpu_index):
    """Pins this process to the CPU cores NVML reports as local to physical GPU `gpu_index`, BEFORE any pinned
    staging buffer is allocated, so that first-touch places the buffers on the GPU's own NUMA node (with one
    process per GPU, eight concurrent 655 MB uploads otherwise cross the socket interconnect).  Best effort:
    returns the sorted core list it bound to, or None when NVML / the affinity call is unavailable or the
    intersection with the allowed cores is empty (containers with a restricted cpuset)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cores = local & os.sched_getaffinity(0)
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return sorted(cores)
    except Exception:
        return None


def shard_range(n_items, rank, world):
    """Contiguous block [start, stop) of `n_items` owned by `rank`; blocks differ by at most one."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def pack_detections(boxes, probs, cls, count):
    """(B,M,4) i32, (B,M) f32, (B,M) i32, (B,) i32 -> dets (B,M,6) f32 rows
    [x1,y1,x2,y2,prob,class] (rows >= count zeroed) + counts (B,) i32."""
    m = boxes.shape[1]
    live = (torch.arange(m, device=boxes.device)[None, :] < count[:, None]).unsqueeze(-1)
    dets = torch.cat([boxes.to(torch.float32), probs.unsqueeze(-1), cls.to(torch.float32).unsqueeze(-1)], dim=-1)
    return torch.where(live, dets, torch.zeros((), dtype=dets.dtype, device=dets.device)), count.to(torch.int32)


def all_gather_detections(dets, counts, group=None):
    """Every rank contributes dets (b,M,6) f32 + counts (b,) i32 with the SAME b and M; returns the
    image-major concatenation over ranks ((world*b, M, 6), (world*b,)).  Single-process: identity."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dets, counts
    world = dist.get_world_size(group)
    out_d = torch.empty((world * dets.shape[0],) + tuple(dets.shape[1:]), dtype=dets.dtype, device=dets.device)
    out_c = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out_d, dets.contiguous(), group=group)
    dist.all_gather_into_tensor(out_c, counts.contiguous(), group=group)
    return out_d, out_c


class _Pending:
    """Handles of collectives issued with async_op=True; `wait()` makes the current stream wait for them."""

    def __init__(self, works):
        self.works = works

    def wait(self):
        for w in self.works:
            w.wait()


def all_gather_rois(rois, count, out_rois=None, out_count=None, group=None, async_op=False):
    """All-gather of the proposal stage's result: rois (b,max_boxes,4) int16 + count (b,) int32 ->
    (world*b, max_boxes, 4) int16, (world*b,) int32.  NCCL has no int16 type, so the RoI rows travel
    as int32 pairs (a reinterpreting view, no copy).  `out_*` may be preallocated buffers.
    `async_op=True` returns (out_rois, out_count, pending): the exchange runs on NCCL's stream next to whatever
    is enqueued afterwards (the RoI layer needs only the LOCAL RoIs); call `pending.wait()` before the gathered
    buffers are read."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return (rois, count, _Pending([])) if async_op else (rois, count)
    world = dist.get_world_size(group)
    if out_rois is None:
        out_rois = torch.empty((world * rois.shape[0],) + tuple(rois.shape[1:]), dtype=rois.dtype, device=rois.device)
    if out_count is None:
        out_count = torch.empty((world * count.shape[0],), dtype=count.dtype, device=count.device)
    w1 = dist.all_gather_into_tensor(out_rois.view(torch.int32), rois.contiguous().view(torch.int32), group=group,
                                     async_op=async_op)
    w2 = dist.all_gather_into_tensor(out_count, count.contiguous(), group=group, async_op=async_op)
    return (out_rois, out_count, _Pending([w1, w2])) if async_op else (out_rois, out_count)
