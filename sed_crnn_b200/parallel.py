"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in
the CPU tests).  The reference has no distributed code at all (train_lightning.py:46 pins devices=1); what
is added is exactly what the path needs:

  * log-mel extraction shards by clip with NO collective (clip i -> rank i mod world);
  * CRNN training is pure data parallel: every rank holds the full (0.4-2.3 M parameter) model, runs its
    slice of the global batch, and the flat fp32 gradient buffer is sum-all-reduced ONCE per step; the
    1/world scaling is folded into the fused clip+Adam kernel.  BatchNorm statistics stay per replica
    (DDP semantics; the reference has no SyncBN).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of range(n) for `rank` (first n % world ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def clip_ids_for_rank(n_clips: int, rank: int, world: int) -> range:
    """Round-robin clip assignment for bulk feature extraction (no collective needed)."""
    return range(rank, n_clips, world)


def batch_slice(global_batch: int, rank: int, world: int) -> slice:
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return slice(rank * per, (rank + 1) * per)


def allreduce_sum_(flat: torch.Tensor, group=None) -> float:
    """In-place sum all-reduce of a flat gradient buffer; returns the factor (1/world) the optimizer must
    apply.  One message per step: the buffer is 1.5-9 MB, i.e. latency-bound on NVLink 5, so there is no
    bucketing."""
    rank, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def broadcast_(flat: torch.Tensor, src: int = 0, group=None) -> None:
    """Make every replica start from rank `src`'s parameters."""
    _, world = world_info(group)
    if world > 1:
        dist.broadcast(flat, src=src, group=group)


# ----------------------------------------------------------------------------------------------- metrics across ranks
def metric_row_shard(n_rows: int, block: int, rank: int, world: int) -> range:
    """Rows of the flattened (N*T) decision matrix that `rank` counts: a contiguous run of WHOLE blocks of `block`
    rows (metrics.py:46-68 takes block maxima over consecutive rows of the flattened axis, utils.py:11-12), so that
    no block straddles two ranks and the 13 integer counts are additive.  Only the last non-empty shard can end in
    the global trailing partial block -- which f1_overall_1sec counts (ceil, metrics.py:50) and er_overall_1sec drops
    (int(), metrics.py:62), exactly as in one process."""
    if block < 1:
        raise ValueError("block must be >= 1")
    n_blocks = -(-n_rows // block)
    b = shard_range(n_blocks, rank, world)
    return range(min(b.start * block, n_rows), min(b.stop * block, n_rows))


def allreduce_counts_(counts: torch.Tensor, group=None) -> torch.Tensor:
    """In-place integer SUM all-reduce of the 13 metric counts (exact: int64)."""
    if counts.dtype != torch.int64:
        raise TypeError("metric counts are int64")
    _, world = world_info(group)
    if world > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def gather_rows_in_global_order(local: torch.Tensor, step_rows, group=None) -> torch.Tensor:
    """local: [rows_local, C] = this rank's rows of every step, concatenated in step order; step_rows: rows per step on
    this rank.  Returns the rows of ALL ranks in the order a single process would have seen them: for each step, rank
    0's slice of the global batch, then rank 1's, ... (the contiguous per-rank slices of `DeviceWindowLoader` /
    `batch_slice`).  Ranks may hold different row counts per step (a last partial batch)."""
    rank, world = world_info(group)
    if world == 1:
        return local
    step_rows = [int(r) for r in step_rows]
    if sum(step_rows) != local.shape[0]:
        raise ValueError("step_rows does not add up to the local row count")
    all_steps = [None] * world
    dist.all_gather_object(all_steps, step_rows, group=group)
    if len({len(s) for s in all_steps}) != 1:
        raise RuntimeError(f"ranks ran different numbers of steps: {[len(s) for s in all_steps]}")
    totals = [sum(s) for s in all_steps]
    pad = max(totals)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    offs = [0] * world
    pieces = []
    for s in range(len(step_rows)):
        for r in range(world):
            n = all_steps[r][s]
            pieces.append(gathered[r][offs[r]:offs[r] + n])
            offs[r] += n
    return torch.cat(pieces)


def sharded_metric_counts(decisions_local: torch.Tensor, targets_local: torch.Tensor, step_rows, block: int,
                          group=None, count_fn=None):
    """The 13 integer counts behind metrics.py for an epoch whose predictions are spread over data-parallel ranks
    (SURVEY 8e row 4): decisions / targets ([rows_local, C], uint8 or float) are exchanged once as bytes, every rank
    counts its block-aligned shard of the GLOBAL flattened axis on its own GPU (`sedb200_threshold_counts`), and the
    counts are summed with one integer all-reduce.  Equal, bit for bit, to counting everything in one process.
    `count_fn(O, T, block) -> 13 ints` defaults to the device counter (`metrics._counts`)."""
    if count_fn is None:
        from . import metrics
        count_fn = metrics._counts
    rank, world = world_info(group)
    O = gather_rows_in_global_order(decisions_local.to(torch.uint8), step_rows, group)
    T = gather_rows_in_global_order(targets_local.to(torch.uint8), step_rows, group)
    rows = metric_row_shard(O.shape[0], block, rank, world)
    import numpy as np
    if len(rows):
        c = np.asarray(count_fn(O[rows.start:rows.stop], T[rows.start:rows.stop], block), dtype=np.int64)
    else:
        c = np.zeros(13, dtype=np.int64)
    counts = torch.from_numpy(c.copy()).to(O.device)
    allreduce_counts_(counts, group)
    return counts.cpu().numpy()


class _RawCudaArray:
    """`__cuda_array_interface__` view of library-owned device memory, so torch can wrap it without a copy."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class P2PGradExchange:
    """The data-parallel exchange step over NVLink peer memory (single node): every rank's backward pass writes
    its flat gradients into a CUDA-IPC-shared region; ONE kernel (`sedb200_p2p_allreduce_clip_adam`) then sums the
    buffers of all ranks with peer loads, clips and applies Adam -- no NCCL call on the data path, sums bit-identical
    on every rank.  `torch.distributed` is used once, at construction, to swap the 64-byte IPC handles."""

    def __init__(self, n_floats: int, device, group=None):
        import ctypes as C
        from . import _lib
        self.L, self.C = _lib.lib(), C
        self.check = _lib.check
        self.device = torch.device(device)
        self.rank, self.world = world_info(group)
        self.n = int(n_floats)
        if self.n % 4:
            raise ValueError("flat parameter count must be a multiple of 4 floats")
        if self.world > 16:
            raise ValueError("P2PGradExchange supports up to 16 ranks on one node")
        nbytes = int(self.L.sedb200_p2p_region_bytes(self.n))
        # Set-up must fail on ALL ranks or on none (a rank that raised between two collectives would hang the
        # others), so errors are collected and agreed on with one MIN all-reduce at the end.
        err = None
        self.own, self.regions = None, []
        with torch.cuda.device(self.device):
            ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
            try:
                self.check(self.L.sedb200_p2p_region_alloc(nbytes, C.byref(ptr), handle))
                self.own = ptr.value
            except Exception as e:                                    # noqa: BLE001
                err = e
            handles = [bytes(handle) if err is None else None]
            if self.world > 1:
                gathered = [None] * self.world
                dist.all_gather_object(gathered, handles[0], group=group)
                handles = gathered
            for r, h in enumerate(handles):
                if r == self.rank or err is not None or h is None:
                    self.regions.append(self.own if r == self.rank else None)
                    continue
                q = C.c_void_p()
                try:
                    self.check(self.L.sedb200_p2p_region_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(q)))
                    self.regions.append(q.value)
                except Exception as e:                                # noqa: BLE001
                    err = e
                    self.regions.append(None)
            torch.cuda.synchronize()
            if self.world > 1:
                ok = torch.tensor([0 if (err is not None or any(q is None for q in self.regions)) else 1],
                                  device=self.device)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # also: every region zeroed + mapped
                if ok.item() == 0 and err is None:
                    err = RuntimeError("a peer rank could not set up its NVLink exchange region")
        if err is not None:
            self.close()
            raise RuntimeError(f"P2PGradExchange set-up failed on rank {self.rank}: {err}") from err
        self.table = (C.c_void_p * self.world)(*self.regions)
        self.grad_bufs = [
            torch.as_tensor(_RawCudaArray(self.own + int(self.L.sedb200_p2p_grad_offset_bytes(self.n, par)), self.n),
                            device=self.device) for par in (0, 1)]
        self.reduced = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self.scratch = torch.zeros(int(self.L.sedb200_p2p_scratch_bytes()) // 4, dtype=torch.float32,
                                   device=self.device)
        self.seq = 0
        # health: the region's sticky status word is copied to pinned host memory after every exchange (4 bytes, same
        # stream, no synchronisation) and looked at before the next one -- a peer that never published makes the
        # kernel abort (parameters untouched, gnorm = NaN) and the NEXT call raise, instead of training on garbage
        self._status_dev = torch.as_tensor(
            _RawCudaArray(self.own + int(self.L.sedb200_p2p_status_offset_bytes()), 1, "<i4"), device=self.device)
        self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._status_evt = torch.cuda.Event()
        self._status_pending = False

    def raise_if_failed(self, wait: bool = False) -> None:
        """Raise if an earlier exchange aborted.  wait=False looks only at copies that have already completed."""
        if self._status_pending and (wait or self._status_evt.query()):
            if wait:
                self._status_evt.synchronize()
            self._status_pending = False
        if not self._status_pending and int(self._status_host[0]) != 0:
            raise RuntimeError(f"P2PGradExchange (rank {self.rank}): a peer did not publish its gradients within the wait "
                               "bound; the optimizer step was skipped (parameters untouched) -- the job must stop")

    def next_grad_buffer(self) -> torch.Tensor:
        """The buffer the NEXT exchange will read on every rank: backward must write its gradients here."""
        return self.grad_bufs[(self.seq + 1) & 1]

    def allreduce_clip_adam(self, params, exp_avg, exp_avg_sq, *, step, lr, betas, eps, weight_decay, clip, gnorm_out):
        self.raise_if_failed()
        self.seq += 1
        with torch.cuda.device(self.device):
            self.check(self.L.sedb200_p2p_allreduce_clip_adam(
                self.table, self.world, self.rank, self.n, self.seq, int(step), params.data_ptr(), exp_avg.data_ptr(),
                exp_avg_sq.data_ptr(), self.reduced.data_ptr(), float(lr), float(betas[0]), float(betas[1]),
                float(eps), float(weight_decay), float(clip), 1.0 / self.world, gnorm_out.data_ptr(),
                self.scratch.data_ptr(), self.scratch.numel() * 4, torch.cuda.current_stream().cuda_stream))
            self._status_host.copy_(self._status_dev, non_blocking=True)
            self._status_evt.record()
            self._status_pending = True
        return self.reduced

    def after_graph_exchange(self) -> None:
        """Bookkeeping after a CUDA-graph replay that contained the exchange kernel (engine.CRNNEngine, cuda_graph=True):
        the exchange number advances and the health word is copied out, exactly as allreduce_clip_adam does."""
        self.seq += 1
        with torch.cuda.device(self.device):
            self._status_host.copy_(self._status_dev, non_blocking=True)
            self._status_evt.record()
            self._status_pending = True

    def status(self) -> int:
        """0 = healthy; non-zero = some peer failed to publish its gradients within the kernel's spin bound."""
        v = self.C.c_uint(0)
        with torch.cuda.device(self.device):
            self.check(self.L.sedb200_p2p_status(self.own, self.C.byref(v)))
        return int(v.value)

    def close(self) -> None:
        if getattr(self, "own", None) is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for r, q in enumerate(self.regions):
                if r != self.rank and q is not None:
                    self.L.sedb200_p2p_region_close(q)
            self.grad_bufs = []
            self._status_dev = None
            self.L.sedb200_p2p_region_free(self.own)
        self.own = None


class DevicePrefetcher:
    """Iterate over (x, y) HOST batches (pinned, as the reference's DataLoader(pin_memory=True) yields them,
    decorte_datamodule.py:130-137) and hand out DEVICE batches, copying batch i+1 on a side stream while batch i
    is being trained on.  Two device buffer pairs are rotated; the consumer stream waits on the copy event, and
    the copy stream waits until the consumer has finished with the buffer it is about to overwrite."""

    def __init__(self, batches, device="cuda"):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.freed = [torch.cuda.Event(), torch.cuda.Event()]
        self.slot = 0
        self.pending = None
        self._deferred = False
        self._released = [False, False]
        self._enqueue()

    def _enqueue(self):
        try:
            x, y = next(self.it)
        except StopIteration:
            self.pending = None
            return
        k = self.slot
        cur = torch.cuda.current_stream(self.device)
        if self.bufs[k] is None or self.bufs[k][0].shape != x.shape or self.bufs[k][1].shape != y.shape:
            # a new buffer pair (first use, or the batch shape changed: a last partial batch).  The caching allocator
            # may hand back a block with work still pending on the current stream, and whatever still reads the pair
            # being replaced runs there too: order the copy after everything enqueued so far, and tell the allocator
            # that the copy stream uses the new blocks
            self.bufs[k] = (torch.empty(x.shape, dtype=x.dtype, device=self.device),
                            torch.empty(y.shape, dtype=y.dtype, device=self.device))
            for b in self.bufs[k]:
                b.record_stream(self.copy_stream)
            self.copy_stream.wait_stream(cur)
        elif self._released[k]:
            self.copy_stream.wait_event(self.freed[k])
        else:
            # the consumer never called release(k): the only safe order is "after everything it has enqueued so far"
            self.copy_stream.wait_stream(cur)
        self._released[k] = False
        with torch.cuda.stream(self.copy_stream):
            self.bufs[k][0].copy_(x, non_blocking=True)
            self.bufs[k][1].copy_(y, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.pending = k
        self.slot ^= 1

    def __iter__(self):
        return self

    def __next__(self):
        if self._deferred:                    # the consumer never called release(): start the copy now
            self._deferred = False
            self._enqueue()
        if self.pending is None:
            raise StopIteration
        k = self.pending
        torch.cuda.current_stream(self.device).wait_event(self.ready[k])
        out = self.bufs[k]
        # the copy of the following batch (into the OTHER buffer pair) is enqueued by release(), i.e. after the consumer
        # has enqueued its own work: a consumer that synchronises every step (sed.py:138 reads the loss) then finds the
        # GPU busy ~25 us earlier
        self._deferred = True
        return out[0], out[1], k

    def release(self, k: int) -> None:
        """call after the work that reads buffer pair k has been enqueued on the current stream"""
        self.freed[k].record(torch.cuda.current_stream(self.device))
        self._released[k] = True
        if self._deferred:
            self._deferred = False
            self._enqueue()
