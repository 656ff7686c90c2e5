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


# ----------------------------------------------------------------------------------------------- metrics across ranks
def test_metric_row_shard_cuts_on_block_boundaries():
    for n_rows in (0, 1, 4, 5, 43, 1000, 1024 * 8 + 3):
        for block in (1, 5, 43):
            for world in (1, 2, 3, 8):
                shards = [parallel.metric_row_shard(n_rows, block, r, world) for r in range(world)]
                assert [i for s in shards for i in s] == list(range(n_rows))
                for s in shards:
                    assert len(s) == 0 or (s.start % block == 0 and (s.stop % block == 0 or s.stop == n_rows))


def _epoch(world, steps=5, per=6, t_out=8, n_cls=2, seed=0):
    """Per-rank, per-step decisions / targets [B_local, T', C]; the last step is a partial global batch."""
    rng = np.random.default_rng(seed)
    data = []
    for s in range(steps):
        sizes = [per] * world if s < steps - 1 else [per - 2] + [per - 3] * (world - 1)
        data.append([(rng.random((b, t_out, n_cls)) < 0.3, rng.random((b, t_out, n_cls)) < 0.25) for b in sizes])
    return data


def _metrics_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import metrics_ref as M
    data = _epoch(world)
    O = np.concatenate([step[rank][0].reshape(-1, 2) for step in data])
    T = np.concatenate([step[rank][1].reshape(-1, 2) for step in data])
    step_rows = [step[rank][0].shape[0] * 8 for step in data]
    for block in (5, 43):
        c = parallel.sharded_metric_counts(torch.from_numpy(O), torch.from_numpy(T), step_rows, block,
                                           count_fn=lambda o, t, b: M.counts13(o.numpy(), t.numpy(), b))
        np.save(os.path.join(out_dir, f"c{block}_{rank}.npy"), c)
    dist.destroy_process_group()


def test_two_rank_metrics_equal_single_process_bit_for_bit(tmp_path):
    """integer counts all-reduced over block-aligned shards == metrics.py in one process
    (metrics.py:46-68; blocks straddle sample boundaries because T' = 8 is not a multiple of 5)."""
    from oracle import metrics_ref as M
    from sed_crnn_b200 import metrics as PM
    world, port = 2, 29300 + os.getpid() % 300
    mp.spawn(_metrics_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    data = _epoch(world)
    # what ONE process would have collected: every global batch = rank 0's slice, then rank 1's (crnn_lightning.py:104-105)
    O = np.concatenate([np.concatenate([step[r][0] for r in range(world)]) for step in data])
    T = np.concatenate([np.concatenate([step[r][1] for r in range(world)]) for step in data])
    for block in (5, 43):
        c0, c1 = np.load(tmp_path / f"c{block}_0.npy"), np.load(tmp_path / f"c{block}_1.npy")
        assert np.array_equal(c0, c1) and np.array_equal(c0, M.counts13(O, T, block))
        got = PM.scores_from_counts(c0)
        want = (M.f1_overall_framewise(O, T), M.er_overall_framewise(O, T), M.f1_overall_1sec(O, T, block),
                M.er_overall_1sec(O, T, block))
        assert got == want
