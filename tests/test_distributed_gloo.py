"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: game sharding, per-rank seeds, weight-image
broadcast and replay gather.  The same functions run over NCCL on the GPU box (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import yy_b200  # noqa: F401
    from yinyang_game_alphazero_b200 import distributed as yyd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = yyd.shard_games(32768 + 1, world, rank)
        img = torch.arange(4096, dtype=torch.int32).view(torch.uint8).clone() if rank == 0 else torch.zeros(16384, dtype=torch.uint8)
        yyd.broadcast_image(img, 0)
        n = 5 + 3 * rank                                 # ragged record counts per rank
        rec = {"counts": torch.full((n, 4), rank + 1, dtype=torch.int16), "ply": torch.arange(n, dtype=torch.int16) + 100 * rank}
        out = yyd.gather_records(rec, dst=0)
        q.put((rank, lo, hi, int(img.view(torch.int32)[1234].item()), None if out is None else
               (out["counts"].numpy().copy(), out["ply"].numpy().copy()), yyd.rank_seed(7, rank)))
    finally:
        dist.destroy_process_group()


def test_sharding_broadcast_gather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, v0, out0, s0), (r1, lo1, hi1, v1, out1, s1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 16385, 16385, 32769)          # contiguous, covers everything, sizes differ by <= 1
    assert v0 == v1 == 1234                                           # rank 1 received rank 0's image
    assert out1 is None and out0 is not None
    counts, ply = out0
    assert counts.shape == (5 + 8, 4) and np.all(counts[:5] == 1) and np.all(counts[5:] == 2)
    assert np.array_equal(ply, np.concatenate([np.arange(5), 100 + np.arange(8)]))
    assert s0 != s1


def test_shard_games_properties():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import yy_b200  # noqa: F401
    from yinyang_game_alphazero_b200.distributed import shard_games
    for total in (0, 1, 7, 4096, 32768):
        for world in (1, 2, 4, 8):
            spans = [shard_games(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _learner_worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
    import yy_b200  # noqa: F401
    from yinyang_game_alphazero_b200 import learner
    from emu_learner_ops import TorchEmuOps
    from test_learner import _reference_net, _batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _reference_net(4, 4, 8, 1, seed=5 + 10 * rank)      # every rank starts from ITS OWN random initialisation ...
        L = learner.Learner(4, 4, 8, 1, batch_size=4, state_dict=net.state_dict(), _ops=TorchEmuOps())   # ... and adopts rank 0's
        planes, pi, z = _batch(net, 4, 4, 8, seed=6)
        sl = slice(4 * rank, 4 * rank + 4)                  # every rank trains on its own half of the global batch
        for _ in range(2):
            L.step(planes[sl], pi[sl], z[sl])
        q.put((rank, {k: v.numpy().copy() for k, v in L.state_dict().items() if "running" not in k and "tracked" not in k}))
    finally:
        dist.destroy_process_group()


def test_data_parallel_learner_world2():
    """Two ranks, one half batch each, DIFFERENT random initialisations per rank: the learner broadcasts rank 0's parameters,
    moments and running statistics when it is built, so both ranks hold identical weights after every step, equal to a
    single process that starts from rank 0's weights and averages the two half-batch gradients before Adam
    (DistributedDataParallel semantics)."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here)); sys.path.insert(0, here)
    import yy_b200  # noqa: F401
    from yinyang_game_alphazero_b200 import learner
    from emu_learner_ops import TorchEmuOps
    from test_learner import _reference_net, _batch
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_learner_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in res[0]:
        assert np.array_equal(res[0][k], res[1][k]), k
    # single-process restatement: two learners sharing weights, gradients averaged by hand
    net = _reference_net(4, 4, 8, 1, seed=5)
    planes, pi, z = _batch(net, 4, 4, 8, seed=6)
    A = learner.Learner(4, 4, 8, 1, batch_size=4, state_dict=net.state_dict(), _ops=TorchEmuOps())
    B = learner.Learner(4, 4, 8, 1, batch_size=4, state_dict=net.state_dict(), _ops=TorchEmuOps())
    for _ in range(2):
        for L, sl in ((A, slice(0, 4)), (B, slice(4, 8))):
            L.b, L.p = 4, 4 * L.A
            L.planes_in.copy_(planes[sl]); L.pi_in.copy_(pi[sl]); L.z_in.copy_(z[sl])
            L._run()
        mean = (A.grads + B.grads) / 2
        A.grads.copy_(mean); B.grads.copy_(mean)
        A._run_adam(); B._run_adam()
    ref = A.state_dict()
    for k in res[0]:
        assert np.allclose(res[0][k], ref[k].numpy(), rtol=1e-6, atol=1e-7), k
