"""World-size-2 gloo tests (CPU) of the data-parallel host logic: sharding helpers and the
sum-all-reduce + 1/world convention, checked against an N-virtual-replica oracle (the reference module
run on each batch shard, gradients averaged)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import crnn_ref as R
from sed_crnn_b200 import parallel


def test_shard_helpers_partition_exactly():
    for n in (0, 1, 7, 128, 10000):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in parallel.shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(parallel.shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
            rr = sorted(i for r in range(world) for i in parallel.clip_ids_for_rank(n, r, world))
            assert rr == list(range(n))
    assert parallel.batch_slice(1024, 3, 8) == slice(384, 512)
    with pytest.raises(ValueError):
        parallel.batch_slice(10, 0, 4)
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _flat_grads(model):
    return torch.cat([p.grad.reshape(-1) for _, p in model.canonical_named_params()])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = dict(R.PRESETS["fork"])
    torch.manual_seed(0)
    model = R.RefCRNN(**cfg)                                  # identical weights on every rank
    flat_w = torch.cat([p.detach().reshape(-1) for _, p in model.canonical_named_params()])
    if rank != 0:
        flat_w.add_(1.0)                                       # pretend this replica drifted ...
    parallel.broadcast_(flat_w, 0)                             # ... and is re-synchronised from rank 0
    x, y = R.synth_batch(cfg, 8, seed=5)
    sl = parallel.batch_slice(8, rank, world)
    model.train()
    R.focal_bce(model(x[sl]), y[sl]).backward()
    g = _flat_grads(model)
    scale = parallel.allreduce_sum_(g)
    g.mul_(scale)
    np.save(os.path.join(out_dir, f"g{rank}.npy"), g.numpy())
    np.save(os.path.join(out_dir, f"w{rank}.npy"), flat_w.numpy())
    assert parallel.world_info() == (rank, world)
    dist.destroy_process_group()


def test_two_rank_gradient_average_matches_virtual_replicas(tmp_path):
    world, port = 2, 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0, g1 = np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy")
    assert np.array_equal(g0, g1)                              # every rank ends with the same gradient
    assert np.array_equal(np.load(tmp_path / "w0.npy"), np.load(tmp_path / "w1.npy"))
    # oracle: run the reference module on each shard in ONE process and average
    cfg = dict(R.PRESETS["fork"])
    x, y = R.synth_batch(cfg, 8, seed=5)
    acc = None
    for r in range(world):
        torch.manual_seed(0)
        m = R.RefCRNN(**cfg)
        m.train()
        sl = parallel.batch_slice(8, r, world)
        R.focal_bce(m(x[sl]), y[sl]).backward()
        g = _flat_grads(m)
        acc = g if acc is None else acc + g
    np.testing.assert_allclose(g0, (acc / world).numpy(), rtol=1e-5, atol=2e-7)


def test_single_process_is_identity():
    t = torch.arange(4.0)
    assert parallel.allreduce_sum_(t) == 1.0 and t.tolist() == [0, 1, 2, 3]
    assert parallel.world_info() == (0, 1)
