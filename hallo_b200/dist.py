"""Frame-window sharding of one denoising window over the GPUs of a box.

Everything on the path except temporal attention is independent per (CFG half, frame), so a rank owns one CFG
half and a contiguous frame group; ranks [0, R/2) hold the uncond half, ranks [R/2, R) the cond half.
  R = 1 : both halves, all frames (no exchange)
  R = 2 : one half each (no temporal exchange; only the tiny CFG-combine all-gather)
  R = 4, 8 : 2 x R/2 frame groups; the temporal-attention K/V of the local frames are all-gathered inside the
             CFG group (NCCL over NVLink), the nm motion frames are replicated on every rank.
"""
from __future__ import annotations

from typing import List, Tuple

from .engine import Shard


def frame_groups(n_frames: int, groups: int) -> List[Tuple[int, ...]]:
    if n_frames % groups != 0:
        raise ValueError(f"{n_frames} frames do not split into {groups} equal groups")
    per = n_frames // groups
    return [tuple(range(g * per, (g + 1) * per)) for g in range(groups)]


def shard_layout(world: int, n_frames: int):
    """[(halves, frames)] per rank."""
    if world == 1:
        return [((0, 1), tuple(range(n_frames)))]
    if world % 2 != 0:
        raise ValueError("world size must be 1 or even (two CFG halves)")
    gs = world // 2
    fg = frame_groups(n_frames, gs)
    return [((r // gs,), fg[r % gs]) for r in range(world)]


def plan_shard(rank: int, world: int, n_frames: int) -> Shard:
    halves, frames = shard_layout(world, n_frames)[rank]
    if world == 1:
        return Shard(halves=halves, frames=frames)
    import torch.distributed as dist
    gs = world // 2
    groups = [dist.new_group(list(range(h * gs, (h + 1) * gs))) for h in (0, 1)]   # every rank creates both groups
    return Shard(halves=halves, frames=frames, group=groups[rank // gs], group_size=gs, world=dist.group.WORLD,
                 world_size=world, rank_in_group=rank % gs)


def gather_temporal_kv(local_kv, motion_kv, group, group_size: int):
    """[fl, L, 2C] local-frame K/V + [nm, L, 2C] replicated motion-frame K/V -> [nm + fl*G, L, 2C] in global frame
    order.  Same code path as DenoiseEngine._gather_kv; kept free of CUDA specifics so the gloo tests exercise it."""
    import torch
    import torch.distributed as dist
    fl, L, C2 = local_kv.shape
    nm = motion_kv.shape[0]
    full = torch.empty(nm + fl * group_size, L, C2, dtype=local_kv.dtype, device=local_kv.device)
    full[:nm].copy_(motion_kv)
    dist.all_gather_into_tensor(full[nm:].reshape(-1), local_kv.contiguous().reshape(-1), group=group)
    return full


def frames_to_pixels(gn_local, send_buf, out_rows, fl: int, group_size: int, group):
    """[fl, L, C] rows of this rank's frames -> out_rows [G*fl, L/G, C]: the pixel slice `rank_in_group` of ALL frames of
    the CFG group in global frame order (all-to-all; chunk r of the result comes from rank r)."""
    import torch.distributed as dist
    C = gn_local.shape[-1]
    Lg = gn_local.numel() // (fl * group_size * C)
    send_buf.view(group_size, fl, Lg, C).copy_(gn_local.view(fl, group_size, Lg, C).permute(1, 0, 2, 3))
    dist.all_to_all_single(out_rows, send_buf, group=group)
    return out_rows


def pixels_to_frames(y, recv_buf, residual, out, fl: int, group_size: int, group):
    """y [G*fl, L/G, C] (all frames, my pixel slice) -> out [fl, L, C] (my frames, all pixels) + residual."""
    import torch
    import torch.distributed as dist
    C = y.shape[-1]
    Lg = y.numel() // (fl * group_size * C)
    dist.all_to_all_single(recv_buf, y, group=group)             # chunk g = my frames at pixel slice g
    torch.add(recv_buf.view(group_size, fl, Lg, C).permute(1, 0, 2, 3), residual.view(fl, group_size, Lg, C),
              out=out.view(fl, group_size, Lg, C))
    return out
