"""Multi-GPU plumbing.  Env batches shard by env index, one process per GPU, and no env
reads another's state (penv.py:52-59 only zips results), so the data path has NO
collective.  The one exchange is the report-time reduction of the four episode counters
(sum of returns, episodes, successes, sum of lengths) -- an all-reduce(sum) of 32 bytes
over NCCL (gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def shard(num_envs_per_gpu, rank=None):
    """Global env-index range owned by `rank`: Philox counters use the GLOBAL index, so
    results do not depend on how many GPUs the batch is spread over."""
    rank = int(os.environ.get('RANK', '0')) if rank is None else rank
    return rank * num_envs_per_gpu, (rank + 1) * num_envs_per_gpu


def reduce_counters(counters):
    """all-reduce(sum) of the float64 counter tensor (first four entries); returns a dict of floats plus
    the derived mean return / success rate / mean length."""
    c = counters.detach().clone()[:4]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ret, n, succ, length = (float(x) for x in c.cpu())
    return {'return_sum': ret, 'episodes': n, 'successes': succ, 'length_sum': length,
            'mean_return': ret / n if n else float('nan'),
            'success_rate': succ / n if n else float('nan'),
            'mean_length': length / n if n else float('nan')}
