"""world_size-2 gloo test of the path's only collective: the counter reduction."""
import os
import sys

import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from combinatorial_rl_tasks_b200.dist import reduce_counters, shard
    c = torch.tensor([10.0 * (rank + 1), 4.0, 1.0 + rank, 800.0], dtype=torch.float64)
    out = reduce_counters(c)
    q.put((rank, out, shard(1 << 20)))
    dist.destroy_process_group()


def test_counter_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, sh in res:
        assert out['return_sum'] == 30.0 and out['episodes'] == 8.0 and out['successes'] == 3.0
        assert out['mean_return'] == 30.0 / 8 and out['success_rate'] == 3.0 / 8 and out['mean_length'] == 200.0
        assert sh == (rank << 20, (rank + 1) << 20)
