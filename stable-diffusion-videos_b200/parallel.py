"""Frame-level data parallelism: one process per GPU, frames of a clip split into contiguous
per-rank blocks, weights broadcast once, decoded uint8 frames gathered to rank 0 over NCCL (NVLink / NVSwitch).

The reference's torch path is single-device; its only multi-device precedent is the Flax twin's `pmap` over the
frame axis with padding (flax_stable_diffusion_pipeline.py:546, 568-578, 594-597, 898-902) — a pure map with no
collective inside, which is what this is.  Works with the `gloo` backend on CPU tensors (tests) and `nccl` on GPU.
"""
import ctypes as C
import math
import os

import torch
import torch.distributed as dist

_COMM = None  # (handle, lib) of the library's own NCCL communicator (sdw_nccl_*), created on first use


def native_comm():
    """libsdwalk's NCCL communicator for CUDA payloads (include/sdwalk.h: sdw_nccl_*).  The torch process group is the
    plumbing: rank 0's ncclUniqueId travels by object broadcast, then every rank joins with its current CUDA device.
    None when single-process, on CPU (gloo tests) or with SDW_NATIVE_NCCL=0 — torch.distributed carries the bytes then."""
    global _COMM
    if _COMM is not None:
        return _COMM
    if world_size() == 1 or not torch.cuda.is_available() or os.environ.get("SDW_NATIVE_NCCL", "1") == "0":
        return None
    from . import _native as N

    lib = N.lib()
    idb = C.create_string_buffer(128)
    err = None
    if rank() == 0:
        try:
            N.check(lib.sdw_nccl_unique_id(idb))
        except N.SdwError as e:  # libnccl not loadable: every rank must take the same decision
            err = str(e)
    raw, err = broadcast_object((bytes(idb.raw), err), src=0)
    if err is not None:
        if rank() == 0:
            import sys
            print(f"[sdwalk] native NCCL unavailable ({err}); frames and weights travel over torch.distributed", file=sys.stderr)
        os.environ["SDW_NATIVE_NCCL"] = "0"
        return None
    h = C.c_void_p()
    N.check(lib.sdw_nccl_init(C.create_string_buffer(raw, 128), C.c_int(rank()), C.c_int(world_size()), C.byref(h)))
    _COMM = (h, lib)
    return _COMM


def destroy_native_comm():
    global _COMM
    if _COMM is not None:
        h, lib = _COMM
        lib.sdw_nccl_destroy.restype = None
        lib.sdw_nccl_destroy(h)
        _COMM = None


def frame_block(n, world, rank):
    """contiguous block [lo, hi) of rank `rank` out of n frames (ceil split; trailing ranks may get fewer / none)."""
    per = math.ceil(n / world) if n > 0 else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def barrier():
    if world_size() > 1:
        dist.barrier()


def broadcast_object(obj, src=0):
    if world_size() == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def init_distributed():
    """(rank, world, local_rank) from the torchrun environment; initialises the process group when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_state_dict(sd, src=0):
    """broadcast every tensor of a state dict from `src` as ONE flat fp16 buffer (1.8 GB for SD-1.4: one NCCL call)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sd
    names = sorted(sd)
    dev = sd[names[0]].device
    flat = torch.cat([sd[k].reshape(-1).to(torch.float16) for k in names])
    comm = native_comm() if flat.is_cuda else None
    if comm is not None:
        from . import _native as N

        N.check(comm[1].sdw_nccl_broadcast_weights(comm[0], N.ptr(flat), C.c_uint64(flat.numel() * 2), C.c_int(src),
                                                   N.stream_ptr()))
    else:
        dist.broadcast(flat, src=src)
    out, off = {}, 0
    for k in names:
        n = sd[k].numel()
        out[k] = flat[off:off + n].view(sd[k].shape).to(dev)
        off += n
    return out


def gather_frames(local_frames, n_total, dst=0):
    """gather per-rank uint8 frame blocks [k_r, H, W, 3] (contiguous frame_block split of n_total) to rank `dst`.

    Returns the [n_total, H, W, 3] tensor on `dst`, None elsewhere.  Every rank pads its block to ceil(n_total / world)
    frames and sends it ONCE: `sdw_nccl_gather_frames` (grouped ncclSend / ncclRecv over NVLink) for CUDA tensors, `dist.gather` on CPU (gloo) — only
    rank `dst` receives, so the bytes on the wire are the frames themselves (an all-gather would move world x that)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_frames
    world, rk = dist.get_world_size(), dist.get_rank()
    per = math.ceil(n_total / world)
    shape = tuple(local_frames.shape[1:])
    buf = torch.zeros((per,) + shape, dtype=local_frames.dtype, device=local_frames.device)
    buf[: local_frames.shape[0]] = local_frames
    comm = native_comm() if buf.is_cuda else None
    if comm is not None:
        from . import _native as N

        recv = torch.empty((world,) + tuple(buf.shape), dtype=buf.dtype, device=buf.device) if rk == dst else None
        N.check(comm[1].sdw_nccl_gather_frames(comm[0], N.ptr(buf), N.ptr(recv), C.c_uint64(buf.numel() * buf.element_size()),
                                               C.c_int(dst), N.stream_ptr()))
        parts = recv
    else:
        parts = [torch.empty_like(buf) for _ in range(world)] if rk == dst else None
        dist.gather(buf, gather_list=parts, dst=dst)
    if rk != dst:
        return None
    keep = []
    for r in range(world):
        lo, hi = frame_block(n_total, world, r)
        keep.append(parts[r][: hi - lo])
    return torch.cat(keep)
