"""One-process-per-GPU plumbing for the sharded path: shuffle blocks are independent, block i belongs
to rank i mod N, there is NO collective on the data path.  torch.distributed is used only for the barrier around the
timed region, the max-over-ranks of the timing, and (on the host side) gathering the per-rank partitionLengths /
checksums so rank 0 can assemble .index/.checksum — the same role Spark's MapOutputTracker plays for the reference.
Backend: nccl on GPUs (bench.py), gloo in the CPU tests.
"""
import os

import numpy as np


class Ranks:
    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.device = device
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            kw = {}
            if backend == "nccl" and device is not None:
                kw["device_id"] = device
            dist.init_process_group(backend or "gloo", **kw)
            self.dist = dist

    # ---- sharding: block i -> rank i mod N (north-star: round-robin)
    def my_blocks(self, n_blocks):
        return np.arange(self.rank, n_blocks, self.world, dtype=np.int64)

    @staticmethod
    def owner(block, world):
        return block % world

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        if self.device is not None:
            import torch

            torch.cuda.synchronize()

    def _tensor(self, values, dtype):
        import torch

        t = torch.tensor(values, dtype=dtype)
        return t.to(self.device) if self.device is not None else t

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        import torch

        t = self._tensor([x], torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        import torch

        t = self._tensor([x], torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_block_results(self, n_blocks, mine):
        """mine: int64 array [len(my_blocks), k] of per-block results (compressed length, checksum, ...).
        Returns the [n_blocks, k] array in block order on every rank (metadata only — never block bytes)."""
        mine = np.ascontiguousarray(mine, dtype=np.int64)
        k = mine.shape[1] if mine.ndim == 2 else 1
        mine = mine.reshape(-1, k)
        if self.dist is None:
            return mine
        import torch

        per = (n_blocks + self.world - 1) // self.world
        pad = np.zeros((per, k), dtype=np.int64)
        pad[: mine.shape[0]] = mine
        t = self._tensor(pad, torch.int64)
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        full = np.zeros((n_blocks, k), dtype=np.int64)
        for r, o in enumerate(outs):
            idx = np.arange(r, n_blocks, self.world)
            full[idx] = o.cpu().numpy()[: idx.size]
        return full

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None
