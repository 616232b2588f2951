"""Episode / lock metric aggregation across the GPUs of one box.

The env kernels accumulate, per env, the episode-end sums of the metric set the reference's RLlib
callbacks log (src/trainers/callbacks.py:152,173,335-345: success_rate, goals_reached,
blocking_count, deadlock_count, livelock_count, deadlock_steps, livelock_steps, throughput,
completion_ratio).  ``mapf_metrics_reduce`` folds them over the envs of a shard; this module does the
only collective of the whole path -- ONE all-reduce(sum) of that 16-double vector, issued off the
step path -- and forms the means the callbacks would have reported (``reduce="mean"``).
"""
from __future__ import annotations

import torch

from ._native import METRIC_NAMES

_MEANS = {
    "return_mean": "return_sum", "length_mean": "length_sum", "success_rate": "success_sum",
    "goals_reached_mean": "goals_reached_sum", "blocking_count_mean": "blocking_count_sum",
    "deadlock_count_mean": "deadlock_count_sum", "livelock_count_mean": "livelock_count_sum",
    "deadlock_steps_mean": "deadlock_steps_sum", "livelock_steps_mean": "livelock_steps_sum",
    "throughput_mean": "throughput_sum", "completion_ratio_mean": "completion_ratio_sum",
    "wfg_cycle_steps_mean": "wfg_cycle_steps_sum",
}


def summarize(vec) -> dict:
    """float64[16] of sums -> {sums..., means...}; means are per finished episode."""
    vals = [float(x) for x in (vec.tolist() if hasattr(vec, "tolist") else vec)]
    out = dict(zip(METRIC_NAMES, vals))
    n = out["episodes"]
    for mean_key, sum_key in _MEANS.items():
        out[mean_key] = out[sum_key] / n if n > 0 else 0.0
    return {k: v for k, v in out.items() if not k.startswith("reserved")}


def allreduce_metrics(vec: torch.Tensor, world_size: int | None = None, group=None, stream=None) -> dict:
    """All-reduce(sum) the per-shard metric vector over the ranks (NCCL on GPUs, gloo in the CPU
    tests) and return sums + means.  ``stream``: optional side CUDA stream so the collective never
    serialises with step launches on the current stream."""
    import torch.distributed as dist

    if world_size is None:
        world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    v = vec.detach().clone()
    if world_size > 1:
        if stream is not None and v.is_cuda:
            stream.wait_stream(torch.cuda.current_stream(v.device))
            with torch.cuda.stream(stream):
                dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
            torch.cuda.current_stream(v.device).wait_stream(stream)
        else:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    return summarize(v)


class AsyncMetrics:
    """A metric all-reduce in flight on a side stream: the step stream never waits for it.

    ``start`` snapshots the shard's metric vector (the snapshot kernel is ordered on the CURRENT stream, behind the
    steps already queued), hands the collective to ``stream`` and returns at once; later step launches on the
    current stream overlap with the collective.  ``result()`` blocks the HOST on the collective's event and returns
    sums + means (``summarize``).  This is the reporting path BASELINE config 4 names: "NCCL metric all-reduce,
    kept off the step path"."""

    def __init__(self, vec: torch.Tensor, world_size: int | None = None, group=None, stream=None):
        import torch.distributed as dist

        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.local = vec.detach().clone()          # on the current stream: behind the steps queued so far
        self.total = self.local.clone()
        self._event = None
        if world_size > 1:
            if stream is not None and self.total.is_cuda:
                stream.wait_stream(torch.cuda.current_stream(self.total.device))
                with torch.cuda.stream(stream):
                    dist.all_reduce(self.total, op=dist.ReduceOp.SUM, group=group)
                    self._event = torch.cuda.Event()
                    self._event.record(stream)
                self.total.record_stream(stream)
            else:
                dist.all_reduce(self.total, op=dist.ReduceOp.SUM, group=group)

    def result(self) -> dict:
        if self._event is not None:
            self._event.synchronize()
        return summarize(self.total)


def shard_range(num_envs_total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block of global env ids owned by ``rank`` (first ranks get the remainder)."""
    base, rem = divmod(int(num_envs_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
