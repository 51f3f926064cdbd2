"""Multi-GPU data parallelism over image pairs.

Pairs are independent units: rank r of `world` takes pairs r, r+world, ... ; every rank holds a
full replica of the weights and produces its own inputs.  No tensor crosses GPUs inside a pair, so
there is no data-path collective: torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU
tests) only (1) broadcasts the pair-index list from rank 0 and (2) gathers the fixed-shape match
tensors [steps, patches, 5] = (x1, y1, x2, y2, confidence) back.
The reference has no distributed code at all (single process, train_patch2pix.py:227).
"""
import torch
import torch.distributed as dist


class PairSharder:
    def __init__(self, rank, world, device):
        self.rank, self.world, self.device = int(rank), int(world), torch.device(device)
        if self.world > 1 and not dist.is_initialized():
            raise RuntimeError('torch.distributed must be initialised before sharding over more than one rank')

    def scatter_pair_indices(self, all_indices):
        """Rank 0's int64 index list is broadcast; returns this rank's shard (indices[rank::world])."""
        idx = all_indices.to(self.device, torch.int64).contiguous()
        if self.world > 1:
            n = torch.tensor([idx.numel()], dtype=torch.int64, device=self.device)
            dist.broadcast(n, src=0)
            if self.rank != 0:
                idx = torch.empty(int(n.item()), dtype=torch.int64, device=self.device)
            dist.broadcast(idx, src=0)
        return idx[self.rank::self.world].cpu()

    def gather_results(self, local):
        """all_gather of equally-shaped per-rank results -> [world, ...] (every rank gets the stack)."""
        if self.world == 1:
            return local.unsqueeze(0)
        parts = [torch.empty_like(local) for _ in range(self.world)]
        dist.all_gather(parts, local.contiguous())
        return torch.stack(parts)

    @staticmethod
    def interleave(stacked):
        """[world, steps, ...] gathered shards -> [steps*world, ...] in global pair order
        (pair p = step*world + rank)."""
        w, s = stacked.shape[:2]
        return stacked.transpose(0, 1).reshape(w * s, *stacked.shape[2:])
