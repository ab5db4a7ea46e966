"""Multi-rank plumbing for the frame-sharded path: which frames a rank owns, a barrier,
and the max-over-ranks reduction of device times.  There is no data-path collective --
ranks never exchange activations; torch.distributed is used for measurement only
(backend nccl on the GPU box, gloo in the CPU tests)."""
import os


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def frames_of_rank(n_frames, rank, world):
    """round-robin ownership: frame f -> rank f mod world"""
    return list(range(rank, n_frames, world))


class Group:
    def __init__(self, backend=None, device=None):
        self.rank, self.world, self.local_rank = env_rank()
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group(backend or "nccl")
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, values):
        """elementwise max of a list of floats over all ranks"""
        if self.dist is None:
            return [float(v) for v in values]
        import torch
        t = torch.tensor(list(values), dtype=torch.float64, device=self.device or "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_over_ranks(self, values):
        if self.dist is None:
            return [float(v) for v in values]
        import torch
        t = torch.tensor(list(values), dtype=torch.float64, device=self.device or "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()
