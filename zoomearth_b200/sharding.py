"""Crop-sharded multi-GPU execution (one process per GPU, torch.distributed).

Crops are independent sequences (attention never crosses cu_seqlens, HF modeling_qwen2_5_vl.py:488-502) and
the reference itself only ever runs data-parallel (``src/eval/infer.py:158-171``), so the path shards by crop
with no data-path collective.  The only exchange is the gather of the output embeddings (C1), a ragged
all-gather: counts first, then one ``all_gather`` of max-padded (T_r, out_hidden) blocks, then a permutation
back to the caller's crop order.  Works with NCCL (GPU) and gloo (CPU tests of the host logic).
"""
import numpy as np
import torch
import torch.distributed as dist


def crop_cost(grid_thw):
    """FLOP estimate per crop: linear + full attention + window attention."""
    g = np.asarray(grid_thw, dtype=np.float64).reshape(-1, 3)
    S = g[:, 0] * g[:, 1] * g[:, 2]
    T = S / 4
    return 5.125e9 * T + 4 * 4 * S * S * 1280 + 28 * 4 * 64 * S * 1280


def partition(costs, world_size):
    """Longest-processing-time-first greedy assignment; deterministic, identical on every rank.
    Returns a list (per rank) of crop indices, each in ascending order."""
    costs = np.asarray(costs, dtype=np.float64)
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world_size
    parts = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += costs[i]
    return [sorted(p) for p in parts]


def gather_embeddings(local_emb, local_tokens, parts, group=None):
    """local_emb: (T_r, D) embeddings of this rank's crops (in the order of parts[rank]); local_tokens: tokens of
    each of those crops.  Returns ((T_total, D) embeddings in global crop order, per-crop token counts)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local_emb.device
    n_total = sum(len(p) for p in parts)
    counts_local = torch.zeros(n_total, dtype=torch.int64, device=dev)
    if len(parts[rank]):
        counts_local[torch.as_tensor(parts[rank], device=dev)] = torch.as_tensor(local_tokens, dtype=torch.int64, device=dev)
    dist.all_reduce(counts_local, group=group)                       # every crop is owned by exactly one rank
    counts = counts_local.cpu().numpy()
    per_rank = [int(counts[p].sum()) if len(p) else 0 for p in parts]
    tmax = max(per_rank + [1])
    D = local_emb.shape[1]
    padded = torch.zeros((tmax, D), dtype=local_emb.dtype, device=dev)
    padded[: local_emb.shape[0]] = local_emb
    gathered = torch.empty((world, tmax, D), dtype=local_emb.dtype, device=dev)
    dist.all_gather_into_tensor(gathered.view(world * tmax, D), padded, group=group)
    # permutation back to global crop order
    starts = np.concatenate([[0], np.cumsum(counts)])
    out = torch.empty((int(starts[-1]), D), dtype=local_emb.dtype, device=dev)
    for r, p in enumerate(parts):
        off = 0
        for i in p:
            n = int(counts[i])
            out[int(starts[i]): int(starts[i]) + n] = gathered[r, off: off + n]
            off += n
    return out, counts


class PeerGather:
    """Gather buffer in symmetric memory for the fused GEMM -> gather path (C1 without a collective).

    Every rank owns one (rows_total, dim) buffer; all of them are mapped into every process
    (torch.distributed._symmetric_memory: CUDA VMM handles exchanged once at construction).  The tower's last GEMM
    (merger.mlp.2 + un-reorder) then stores each embedding row into its own buffer and, over NVLink, into every
    peer's buffer at the same global row, so after one cross-rank barrier each rank holds all embeddings - the
    transfer overlaps the GEMM tile by tile and no NCCL collective runs on the data path.
    """

    def __init__(self, rows_total, dim, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.buffer = symm_mem.empty((rows_total, dim), dtype=dtype, device=device)
        self.handle = symm_mem.rendezvous(self.buffer, self.group)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.local_ptr = ptrs[self.rank]
        self.peer_ptrs = [p for r, p in enumerate(ptrs) if r != self.rank]
        if len(self.peer_ptrs) > 8:
            raise ValueError("the fused gather supports at most 9 ranks per node")

    def barrier(self):
        """All ranks' stores are visible after this (device-side barrier over the signal pads, on the current stream)."""
        self.handle.barrier()
