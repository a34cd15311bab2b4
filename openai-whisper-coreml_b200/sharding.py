"""Chunk-level data parallelism: independent 30 s chunks are partitioned contiguously over the ranks of
one box; the hot path has no collective. The only exchange is the load-time weight broadcast and an optional gather of
the token IDs on rank 0."""
from __future__ import annotations

from typing import List, Tuple


def partition(n_chunks: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [start, end) block per rank; the first n_chunks % world_size ranks get one extra chunk
    (60 windows over 8 GPUs -> 8,8,8,8,7,7,7,7)."""
    if world_size < 1 or n_chunks < 0:
        raise ValueError("bad partition arguments")
    base, extra = divmod(n_chunks, world_size)
    out, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def gather_tokens(tokens, lens, n_total: int, world_size: int, rank: int, group=None):
    """All-gather the per-rank [n_local, L] int32 token matrices (ragged over ranks) into [n_total, L] on every rank.
    Works with any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist

    parts = partition(n_total, world_size)
    width = max(e - s for s, e in parts)
    L = tokens.shape[1]
    pad_t = torch.zeros((width, L), dtype=torch.int32, device=tokens.device)
    pad_l = torch.zeros((width,), dtype=torch.int32, device=tokens.device)
    n = tokens.shape[0]
    pad_t[:n] = tokens
    pad_l[:n] = lens
    all_t = [torch.empty_like(pad_t) for _ in range(world_size)]
    all_l = [torch.empty_like(pad_l) for _ in range(world_size)]
    dist.all_gather(all_t, pad_t, group=group)
    dist.all_gather(all_l, pad_l, group=group)
    toks = torch.cat([all_t[r][: e - s] for r, (s, e) in enumerate(parts)], dim=0)
    ls = torch.cat([all_l[r][: e - s] for r, (s, e) in enumerate(parts)], dim=0)
    return toks, ls


def broadcast_weights(whisper, device, src: int = 0, group=None) -> int:
    """Load-time NCCL broadcast of the packed weight arena from `src` (north_star: NCCL over NVLink only here).
    Returns the number of bytes broadcast."""
    import torch
    import torch.distributed as dist

    ptr, nbytes = whisper.weight_arena()

    class _Arena:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    t = torch.as_tensor(_Arena(), device=device)
    dist.broadcast(t, src=src, group=group)
    torch.cuda.synchronize(device)
    whisper.mark_weights_loaded()
    return nbytes


def verify_ranks(whisper, device, rank: int, world_size: int, group=None) -> int:
    """After the broadcast: all-gathers the 64-bit checksum of every rank's weight arena and raises unless they agree.
    Returns the number of ranks verified. (A broken broadcast would otherwise still decode — to different tokens.)"""
    import torch
    import torch.distributed as dist

    mine = whisper.weights_checksum()
    t = torch.tensor([mine & 0x7fffffff, (mine >> 31) & 0x7fffffff, mine >> 62], dtype=torch.int64, device=device)
    every = [torch.empty_like(t) for _ in range(world_size)]
    dist.all_gather(every, t, group=group)
    vals = [int(e[0]) | (int(e[1]) << 31) | (int(e[2]) << 62) for e in every]
    if any(v != vals[0] for v in vals):
        raise RuntimeError(f"weight arenas differ across ranks after the broadcast: {[hex(v) for v in vals]}")
    return world_size


def transcribe_windows_sharded(whisper, windows, opts, rank: int, world_size: int, device=None, group=None):
    """BASELINE config 5 as a job: `windows` [n, 480000] (every rank holds, or can produce, the same array) are partitioned
    contiguously over the ranks (60 windows on 8 GPUs -> 8,8,8,8,7,7,7,7), each rank transcribes its block `max_batch` at
    a time, and the token rows are all-gathered (the only exchange, a few KB). Returns (tokens [n, L], lens [n]) on every
    rank. With world_size == 1 no process group is needed."""
    import numpy as np
    import torch

    n = int(windows.shape[0])
    s, e = partition(n, world_size)[rank]
    L = len(opts.initial_tokens) + opts.sample_len
    toks = np.zeros((e - s, L), dtype=np.int32)
    lens = np.zeros((e - s,), dtype=np.int32)
    for i in range(s, e, whisper.max_batch):
        j = min(e, i + whisper.max_batch)
        t, l, _ = whisper.transcribe(windows[i:j], opts)
        toks[i - s:j - s] = t
        lens[i - s:j - s] = l
    if world_size == 1:
        return toks, lens
    dev = device if device is not None else torch.device("cpu")
    all_t, all_l = gather_tokens(torch.from_numpy(toks).to(dev), torch.from_numpy(lens).to(dev), n, world_size, rank, group=group)
    return all_t.cpu().numpy(), all_l.cpu().numpy()
