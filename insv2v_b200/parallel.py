"""Clip-parallel execution: independent clips (video x prompt x cfg combinations — the reference's outer
`product(...)` loop, insv2v_run_loveu_tgve.py:83,101) are sharded across ranks, one process per GPU, full weight replica
per rank, no data-path collective; the only exchange is ONE all-gather of the decoded frames per batch of clips
(NCCL over NVLink on the GPU box, gloo in the CPU tests). Chained clips of one long video stay on one rank
(clip k+1 needs clip k's final latents, insv2v_run_loveu_tgve.py:138-161)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def clips_for_rank(n_clips, rank, world):
    """Round-robin: clip i -> rank i mod world. Returns the list of global clip indices this rank owns."""
    if n_clips < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad sharding request n_clips={n_clips} rank={rank} world={world}")
    return list(range(rank, n_clips, world))


def gather_frames(local_frames, n_clips, rank, world):
    """local_frames: [n_local, F, 3, H, W] decoded clips of this rank (n_local may differ by one between ranks).
    Returns [n_clips, F, 3, H, W] in global clip order on every rank, using a single all_gather_into_tensor."""
    if world == 1:
        return local_frames
    per_rank = (n_clips + world - 1) // world
    shape = local_frames.shape[1:]
    pad = torch.zeros((per_rank,) + tuple(shape), dtype=local_frames.dtype, device=local_frames.device)
    pad[:local_frames.shape[0]] = local_frames
    out = torch.empty((world * per_rank,) + tuple(shape), dtype=local_frames.dtype, device=local_frames.device)
    dist.all_gather_into_tensor(out, pad.contiguous())
    out = out.reshape(world, per_rank, *shape)
    # rank r holds clips r, r+world, ...: slot j of rank r is global clip j*world + r
    ordered = out.transpose(0, 1).reshape(per_rank * world, *shape)
    return ordered[:n_clips].contiguous()


def run_clips(edit_fn, clip_inputs, rank, world):
    """edit_fn(clip_input) -> [F, 3, H, W]; clip_inputs: list of per-clip inputs (same on every rank).
    Each rank edits its own clips; all ranks return all decoded clips. Needs at least one clip per rank: the check is
    the same deterministic condition on EVERY rank and runs before any work, so a short job fails everywhere instead
    of leaving the ranks that own clips waiting in the all-gather."""
    if len(clip_inputs) < world:
        raise ValueError(f"run_clips: {len(clip_inputs)} clips for {world} ranks; every rank needs at least one clip "
                         "(launch with fewer ranks, or batch more clips)")
    mine = clips_for_rank(len(clip_inputs), rank, world)
    local = torch.stack([edit_fn(clip_inputs[i]) for i in mine], dim=0)
    return gather_frames(local, len(clip_inputs), rank, world)
