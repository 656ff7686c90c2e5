#!/usr/bin/env python3
"""bench.py -- CRNN train frames/sec (BASELINE.json metric) on N B200s, plus the log-mel leg.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one full training step of the hot path on one synthetic batch: forward + loss + backward
+ [NCCL all-reduce of the flat gradient] + global-norm clip + Adam.  Workload at N=1 is BASELINE.json
configs[1] (binaural SEDnet, seq_len 256, batch 128); at N>1 the per-GPU batch stays 128 (weak
scaling; N=8 is configs[3]'s global batch 1024).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CRNN train frames/sec"
UNIT = "frames/s"
PER_GPU_BATCH = 128


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]),
                    tf_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------ FLOP model
def crnn_flops(cfg, batch):
    """Algorithmic FLOPs of one training step, SURVEY.md 8(d) convention: 2*M*K*N per contraction,
    backward = dgrad + wgrad, no dgrad for conv1, elementwise work excluded."""
    H, W = cfg.H, cfg.W
    fwd = 0.0
    per = {}
    cin, w = cfg.in_ch, W
    for i, p in enumerate(cfg.pool):
        f = 2.0 * batch * H * w * 9 * cin * cfg.conv_ch
        per[f"conv{i}"] = f
        fwd += f
        w //= p
        cin = cfg.conv_ch
    T, gin = cfg.seq_len_out, cfg.flat
    for i, h in enumerate(cfg.gru_units):
        per[f"gru{i}.proj"] = 2.0 * batch * T * gin * 6 * h
        per[f"gru{i}.rec"] = 2.0 * batch * T * h * 6 * h
        fwd += per[f"gru{i}.proj"] + per[f"gru{i}.rec"]
        gin = 2 * h
    for i, u in enumerate(list(cfg.dense_units) + [cfg.n_classes]):
        per[f"dense{i}"] = 2.0 * batch * T * gin * u
        fwd += per[f"dense{i}"]
        gin = u
    total = 3.0 * fwd - per["conv0"]
    return total, per


def phase_model(cfg, B, per):
    """phase name -> (bound, algorithmic FLOPs or bytes per launch, kernel).  Bytes: each tensor the phase must
    read or write once, fp32."""
    H, C = cfg.H, cfg.conv_ch
    m = {}
    w, cin = cfg.W, cfg.in_ch
    tc_ok = cfg.tensor_cores and C % 128 == 0
    for i, p in enumerate(cfg.pool):
        y = 4.0 * B * H * w * C
        a = 4.0 * B * H * (w // p) * C
        xin = 4.0 * B * H * w * cin
        if i == 0:
            m["conv0.fwd"] = ("hbm", y + xin, "conv0_fwd_stats_kernel (direct fp32 conv + BN statistics)")
            m["conv0.bwd_fused"] = ("hbm", y + a + xin, "conv0_bwd_fused_t_kernel (BN/ReLU/pool backward + wgrad, dy never written)")
            # lean block 0 (conv output never stored): input + pooled output + one winner byte per output element
            m["conv0.stats"] = ("hbm", xin, "conv0_gram_kernel (patch moments of the input -> BatchNorm statistics)")
            m["conv0.fwd_fused"] = ("hbm", xin + a + a / 4,
                                    "conv0_tc_fwd_kernel (tcgen05 conv + BN + ReLU + max-pool + dropout epilogue, conv output never stored)"
                                    if cfg.tensor_cores else
                                    "conv0_lean_fwd_kernel (conv + BN + ReLU + max-pool + dropout in registers)")
            m["conv0.bwd_lean"] = ("hbm", xin + a + a / 4, "conv0_lean_bwd_kernel (winner contributions to dW / d gamma / d beta)")
        elif tc_ok:
            for ph, kn in (("fwd", "conv_tc_kernel"), ("dgrad", "conv_tc_kernel"), ("wgrad", "wgrad_tc_kernel")):
                m[f"conv{i}.{ph}"] = ("tensor", per[f"conv{i}"], f"{kn} (tcgen05 implicit GEMM, 3-term bf16 split) + bf16 plane split")
        m[f"pool{i}.fwd"] = ("hbm", y + a, "bn_relu_pool_fwd_t_kernel (BN + ReLU + max-pool + dropout: conv output in, pooled planes out)")
        m[f"pool{i}.bwd_sums"] = ("hbm", 2 * a, "bn_bwd_sums_act_kernel (BatchNorm backward sums from the saved block output and dA)")
        if i > 0:
            m[f"pool{i}.bwd_dy"] = ("hbm", 2 * y + a, "bn_pool_bwd_dy_t_kernel (dy of every conv output element as bf16 planes)")
        w //= p
        cin = C
    return m


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0, period_ms=50):
        self.index, self.rows, self.proc, self.period_ms = index, [], None, int(period_ms)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", str(self.period_ms)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first(self, timeout=5.0):
        """block until nvidia-smi has delivered its first sample (its start-up takes ~1 s)"""
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout and self.proc is not None:
            time.sleep(0.02)

    def mark(self):
        return time.perf_counter()

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        rows = [r for (t, r) in self.rows if (t_begin is None or t >= t_begin) and (t_end is None or t <= t_end)]
        window = f"timed + end-to-end regions, {self.period_ms} ms period"
        if len(rows) < 2:                       # region shorter than two sampling periods: add the warm-up steps
            rows = [r for (_, r) in self.rows]  # (same kernels, same load; the sampler starts just before them)
            window = f"warm-up + timed + end-to-end regions, {self.period_ms} ms period"
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on this box's host cores: the CRNN oracle is the
    same stock torch.nn modules the reference instantiates (crnn_lightning.py:41-73), the trainer step is
    forward + loss + backward + clip + Adam (oracle/crnn_ref.py).  /root/reference does not exist on the
    GPU box, so this is cpu_baseline.kind = "port"."""
    if rank != 0:
        return
    import torch
    from oracle import crnn_ref as R
    preset = dict(R.PRESETS[args.config])
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_b = args.ref_batch
    model = R.RefCRNN(**preset, dropout=0.5, dropout_each_block=True) if args.config != "fork" else R.RefCRNN(**preset, dropout=0.4)
    opt = R.make_adam(model, 1e-3, 1e-4)
    x, y = R.synth_batch(preset, sample_b, seed=0)
    for _ in range(max(1, min(args.warmup, 2))):
        R.train_step(model, opt, x, y, "bce", 1.0)
    steps = max(1, min(args.steps, 8))
    t0 = time.perf_counter()
    for _ in range(steps):
        R.train_step(model, opt, x, y, "bce", 1.0)
    dt = (time.perf_counter() - t0) / steps
    T = preset["seq_len"]
    val = sample_b * T / dt
    sample = (f"{steps} timed steps (of --steps {args.steps}) at batch {sample_b} of the {args.config} config "
              f"(full batch {PER_GPU_BATCH}), torch-CPU fp32, {torch.get_num_threads()} threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.gpus), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_name(config, n):
    base = {"c2": "binaural (2-ch, 80-band) SEDnet DCASE2017 task3 CRNN train step, seq_len 256, 128 filters, "
                  "pool [5,2,2] over mel, 2xBiGRU(32), dense 16, 6 classes",
            "c1": "mono SEDnet CRNN train step, seq_len 256, 128 filters, pool [5,2,2], 2xBiGRU(32), 6 classes",
            "c5": "long-context CRNN train step, seq_len 2048, 256 filters, 3xBiGRU(128), 16 classes",
            "fork": "fork-default CRNN (train_constants.py) train step", "sedpy": "sed.py CRNN train step"}[config]
    return f"{base}; batch {PER_GPU_BATCH}/GPU x {n} GPU = global batch {PER_GPU_BATCH * n}"


# ------------------------------------------------------------------------------------------ our arm
def parse_prof(L):
    import ctypes as C
    buf = C.create_string_buffer(1 << 16)
    L.sedb200_prof_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, cnt = line.split()
        out[name] = (float(ms), int(cnt))
    return out


def logmel_leg(torch, feature, L, pk, rank, world=1, dist=None):
    """BASELINE configs[2] leg: bulk log-mel of 3-min stereo clips.  Clips shard by index across ranks with NO
    collective (parallel.clip_ids_for_rank); every rank keeps `n_clips` synthetic clips resident in HBM
    (kernel-only, CUDA events, max over ranks) and rank 0 also measures the host-buffer path."""
    n_clips, S = 32, 180 * 44100
    g = torch.Generator(device="cuda").manual_seed(1000 + rank)
    x = torch.empty(n_clips, 2, S, device="cuda").normal_(0, 0.1, generator=g)            # 2.03 GB > L2
    out = torch.empty(n_clips, feature.n_frames(S), 80, device="cuda")
    for _ in range(3):
        feature.mbe_device(x, out=out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        feature.mbe_device(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    alg_bytes = x.numel() * 4 + out.numel() * 4
    gbs = alg_bytes / ms / 1e6
    res = {
        "workload": f"{n_clips} synthetic 3-min stereo clips per GPU resident in HBM (2.03 GB in, > L2), kernel-only, "
                    f"{world} GPU(s), clips sharded by index, no collective",
        "audio_s_per_s": world * n_clips * 180.0 / (ms * 1e-3),
        "frames_per_s": world * n_clips * 2 * feature.n_frames(S) / (ms * 1e-3),
        "ms_per_launch": ms, "ten_k_clips_s_est": 10000 / (n_clips * world) * ms * 1e-3,
        "roofline": {"bound": "hbm", "kernel": "logmel_kernel", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                     "frac": gbs / pk["hbm"], "traffic": 528320256 * n_clips / 8, "per_gpu": True,
                     "peak_source": pk["src"],
                     "traffic_source": "profiles/r01_logmel_full.ncu-rep (8 clips: 508.5 MB read + 19.8 MB written), scaled"},
    }
    if rank == 0:
        # e2e: pinned host PCM -> device -> kernel -> host log-mel, 4 clips per call; float32 as feature.py decodes it
        # and 16-bit PCM as a WAV / s16le decoder delivers it (sedb200_logmel_i16: half the H2D bytes).  Median of 7
        # calls after 3 warm-up calls (the first transfers on a fresh box run well below the link rate).
        hb = torch.empty(4, 2, S).normal_(0, 0.1).pin_memory()
        hb16 = (hb.clamp(-1, 1) * 32767).to(torch.int16).pin_memory()
        ho = torch.empty(4, feature.n_frames(S), 80).pin_memory()

        def e2e_call(src):
            t0 = time.perf_counter()
            d = src.cuda(non_blocking=True)
            feature.mbe_device(d, out=out[:4])
            ho.copy_(out[:4], non_blocking=True)
            torch.cuda.synchronize()
            return time.perf_counter() - t0

        for key, src, nbytes, note in (("e2e", hb, hb.numel() * 4, "float32 PCM, one rank, PCIe-bound"),
                                       ("e2e_int16", hb16, hb16.numel() * 2, "int16 PCM ingest, one rank, PCIe-bound")):
            for _ in range(3):
                e2e_call(src)
            ts = sorted(e2e_call(src) for _ in range(7))
            res[key] = {"audio_s_per_s": 4 * 180.0 / ts[3], "h2d_bytes_per_call": nbytes,
                        "d2h_bytes_per_call": ho.numel() * 4, "h2d_gb_per_s": nbytes / ts[3] / 1e9, "note": note}
    del x
    return res


def other_configs_leg(torch, config, engine, main_config):
    """The BASELINE configs that are not the bench line (parity-test cases), timed for reference on one GPU:
    device-resident batches, CUDA events, 3 warm-up + 5 timed steps each.  Informational only."""
    out = {}
    for name, B in (("c1", 128), ("c5", 128), ("fork", 128)):
        if name == main_config:
            continue
        try:
            cfg = config.PRESETS[name]
            eng = engine.CRNNEngine(cfg, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0, seed=99)
            eng.init_default(seed=0)
            g = torch.Generator(device="cuda").manual_seed(5)
            x = torch.randn(cfg.input_shape(B), device="cuda", generator=g)
            y = (torch.rand(cfg.target_shape(B), device="cuda", generator=g) < 0.2).float()
            for _ in range(3):
                eng.train_step(x, y)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                eng.train_step(x, y)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            out[name] = {"workload": workload_name(name, 1), "per_gpu_batch": B, "seq_len": cfg.seq_len, "ms_per_step": ms,
                         "frames_per_s": B * cfg.seq_len / (ms * 1e-3)}
            del eng, x, y
            torch.cuda.empty_cache()
        except Exception as e:                                  # noqa: BLE001 -- informational leg must not sink the run
            out[name] = {"error": str(e)[:200]}
    return out


def cpu_baselines(args, torch):
    """Reference CPU path beside the GPU number (rank 0, N=1 only; bounded samples)."""
    from oracle import crnn_ref as R, logmel_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    preset = dict(R.PRESETS[args.config])
    b = args.ref_batch
    model = R.RefCRNN(**preset, dropout=0.5, dropout_each_block=True)
    opt = R.make_adam(model, 1e-3, 1e-4)
    x, y = R.synth_batch(preset, b, seed=0)
    R.train_step(model, opt, x, y, "bce", 1.0)
    t0 = time.perf_counter()
    n = 2
    for _ in range(n):
        R.train_step(model, opt, x, y, "bce", 1.0)
    dt = (time.perf_counter() - t0) / n
    val = b * preset["seq_len"] / dt
    clip = logmel_ref.synth_clip(0, 180 * 44100, 1, "noise")[0]
    t0 = time.perf_counter()
    logmel_ref.mbe(clip)
    lm = 180.0 / (time.perf_counter() - t0)
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/crnn_ref.py (stock torch.nn, the modules the reference instantiates) on torch-CPU fp32, "
                      f"{torch.get_num_threads()} threads: 1 warm-up + {n} timed steps at batch {b} of the same config "
                      f"(full batch {PER_GPU_BATCH}); log-mel oracle (numpy/scipy float64 FFT, librosa semantics), "
                      f"one 3-min mono clip, single process",
            "logmel_audio_s_per_s": lm}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: libsedb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from sed_crnn_b200 import _lib, config, engine, feature
    L = _lib.lib()
    _lib.check(L.sedb200_device_check(-1))
    pk = peaks()
    cfg = config.PRESETS[args.config]
    B = PER_GPU_BATCH
    # gradient exchange: the fused NVLink peer-memory kernel when there is more than one rank ("auto"); if its
    # set-up fails (agreed on by all ranks) the run uses NCCL and SAYS so in config.grad_exchange
    gx_mode, gx_note = ("nccl" if world == 1 else "p2p") if args.grad_exchange == "auto" else args.grad_exchange, None
    try:
        eng = engine.CRNNEngine(cfg, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0, seed=1234 + rank,
                                grad_exchange=gx_mode)
    except RuntimeError as e:
        if gx_mode != "p2p" or args.grad_exchange == "p2p":
            raise
        gx_mode, gx_note = "nccl", f"p2p set-up failed: {e}"
        sys.stderr.write(f"[bench] rank {rank}: {gx_note}; using NCCL\n")
        eng = engine.CRNNEngine(cfg, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0, seed=1234 + rank)
    eng.init_default(seed=0)                         # identical weights on every rank

    gx = torch.Generator().manual_seed(100 + rank)
    n_in = 4                                         # rotate a few distinct batches
    xs_h = [torch.randn(cfg.input_shape(B), generator=gx).pin_memory() for _ in range(n_in)]
    ys_h = [(torch.rand(cfg.target_shape(B), generator=gx) < 0.2).float().pin_memory() for _ in range(n_in)]
    xs = [t.cuda() for t in xs_h]
    ys = [t.cuda() for t in ys_h]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (the clock sampler starts here so that it is running well before the timed region)
    sampler = ClockSampler(local_rank, args.clock_period_ms)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for i in range(max(args.warmup, 3)):
        eng.train_step(xs[i % n_in], ys[i % n_in])
    barrier()

    # ---- timed region: device-resident inputs
    l0 = L.sedb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_clk0 = sampler.mark()
    e0.record()
    for i in range(args.steps):
        loss, _ = eng.train_step(xs[i % n_in], ys[i % n_in])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = L.sedb200_launch_count() - l0
    t = torch.tensor([ms_total], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    frames_per_step = B * cfg.seq_len * world
    value = frames_per_step / (ms_step * 1e-3)
    final_loss = loss.item()

    # ---- end-to-end through the public API with HOST batches: every step copies its pinned host batch to the
    #      device (parallel.DevicePrefetcher: batch i+1 is copied on a side stream while batch i trains, the way
    #      DataLoader(pin_memory=True) is meant to be used) and reads the loss back to the host (sed.py:138)
    from sed_crnn_b200.parallel import DevicePrefetcher
    loss_h = torch.empty(1).pin_memory()

    def host_batches(n):
        for i in range(n):
            yield xs_h[i % n_in], ys_h[i % n_in]

    # one pipeline for warm-up and timed steps: 3 untimed steps bring the copy stream, the pinned buffers and the
    # caching allocator to steady state, then exactly args.steps steps are timed
    e2e_warm = 3
    pf = DevicePrefetcher(host_batches(e2e_warm + args.steps))
    for n_done, (xd, yd, k) in enumerate(pf):
        if n_done == e2e_warm:
            barrier()
            e0.record()
        l, _ = eng.train_step(xd, yd)
        pf.release(k)
        loss_h.copy_(l.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the caller reads the loss every step
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = frames_per_step / (t2.item() / args.steps * 1e-3)
    clocks = sampler.stop(t_clk0, sampler.mark()) if rank == 0 else None

    # ---- phase breakdown (extra instrumented steps, CUDA events on the launching stream)
    L.sedb200_prof_enable(1)
    prof_steps = 3
    for i in range(prof_steps):
        eng.train_step(xs[i % n_in], ys[i % n_in])
    torch.cuda.synchronize()
    prof = parse_prof(L)
    L.sedb200_prof_enable(0)

    lm = None
    if not args.no_logmel:
        lm = logmel_leg(torch, feature, L, pk, rank, world, dist if world > 1 else None)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_flops, per = crnn_flops(cfg, B)
    phases = {k: v[0] / prof_steps for k, v in prof.items()}
    model = phase_model(cfg, B, per)
    # dominant kernel = the KERNEL (not the launch) with the largest share of the step: phases that are launches of
    # the same kernel are pooled; achieved = algorithmic work of its launches / their device time
    groups = {}
    for k in phases:
        if k in model:
            gname = model[k][2].split(" ")[0]
            groups.setdefault(gname, []).append(k)
    dom_kernel = max(groups, key=lambda gname: sum(phases[k] for k in groups[gname]))
    dks = groups[dom_kernel]
    kind, kname = model[dks[0]][0], model[dks[0]][2]
    amount_step = sum(model[k][1] for k in dks)                       # algorithmic bytes / FLOPs of all its launches
    ms_kernel_step = sum(phases[k] for k in dks)
    n_launch = sum(prof[k][1] for k in dks) / prof_steps
    dom_ms = ms_kernel_step / n_launch
    amount = amount_step / n_launch
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get(args.config, {})
        have = [k for k in dks if k in tr]
        if have:                                   # per launch, averaged over the launches ncu captured
            traffic = sum(tr[k] for k in have) / len(have)
    except Exception:
        pass
    if kind == "tensor":
        ach = amount_step / (ms_kernel_step * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": kname, "phases": dks, "achieved": ach, "peak": pk["tf_sustained"],
                "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "traffic": traffic,
                "traffic_note": "dram bytes per launch, mean of this kernel's launches in one step as captured by ncu "
                "--set full (profiles/ncu_traffic.json, profiles/r01_conv_tc_final_metrics.csv)",
                "peak_source": pk["src"] + " (sustained bf16)", "algorithmic_flops_per_launch": amount,
                "mma_tflops_issued": 3 * ach, "frac_issued": 3 * ach / pk["tf_sustained"],
                "note": "fp32-grade 3-term bf16 split: the tensor pipe executes 3x the algorithmic FLOPs; frac is "
                "algorithmic FLOPs / bf16 peak, frac_issued the tensor-pipe work actually issued / bf16 peak"}
    else:
        ach = amount_step / (ms_kernel_step * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": kname, "phases": dks, "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                "frac": ach / pk["hbm"], "traffic": traffic, "peak_source": pk["src"],
                "algorithmic_bytes_per_launch": amount}
    roof.update({"launches_per_step": n_launch, "ms_per_launch": dom_ms,
                 "share_of_step": ms_kernel_step / sum(phases.values()),
                 "whole_step": {"algorithmic_tflops": total_flops / (ms_step * 1e-3) / 1e12,
                                "frac_of_bf16_peak": total_flops / (ms_step * 1e-3) / 1e12 / pk["tf_sustained"]}})
    # the largest CUDA-core (non-tensor) kernel is reported next to it: its bytes against the HBM roofline
    nt = [gname for gname in groups if model[groups[gname][0]][0] == "hbm"]
    if nt and kind == "tensor":
        gname = max(nt, key=lambda gname: sum(phases[k] for k in groups[gname]))
        ks = groups[gname]
        t_ms = sum(phases[k] for k in ks)
        by = sum(model[k][1] for k in ks)
        roof["cuda_core_kernel"] = {"kernel": model[ks[0]][2], "phases": ks, "ms_per_step": t_ms,
                                    "algorithmic_bytes_per_step": by, "achieved": by / (t_ms * 1e-3) / 1e9,
                                    "unit": "GB/s", "frac_of_hbm": by / (t_ms * 1e-3) / 1e9 / pk["hbm"]}
    tc = [k for k in phases if k in model and model[k][0] == "tensor"]
    if tc and kind != "tensor":
        k = max(tc, key=lambda k: phases[k])
        t_ms = phases[k] / (prof[k][1] / prof_steps)
        a = model[k][1] / (t_ms * 1e-3) / 1e12
        roof["tensor_kernel"] = {"phase": k, "kernel": model[k][2], "achieved": a, "unit": "TFLOP/s",
                                 "peak": pk["tf_sustained"], "frac": a / pk["tf_sustained"], "ms_per_launch": t_ms,
                                 "mma_tflops_issued": 3 * a}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, world), "per_gpu_batch": B, "global_batch": B * world,
                   "seq_len": cfg.seq_len, "parallelism": f"dp{world}", "l2": "per-step working set (>1 GB of "
                   "activations) exceeds the 126 MB L2; 4 input batches rotated", "loss": "bce", "optimizer":
                   "clip 1.0 + Adam(1e-3, wd 1e-4)", "dropout": cfg.dropout, "final_loss": final_loss,
                   "grad_exchange": {"nccl": "none (1 GPU)" if world == 1 else "NCCL all-reduce + clip/Adam kernels",
                                     "p2p": "one fused kernel: NVLink peer-memory all-reduce + clip + Adam"}[gx_mode],
                   **({"grad_exchange_note": gx_note} if gx_note else {})},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": (xs_h[0].numel() + ys_h[0].numel()) * 4 * world, "d2h_bytes_per_step": 4 * world},
        "gpu_launches": int(launches),
        "roofline": roof,
        "phases_ms": {k: round(v, 4) for k, v in phases.items()},
    }
    if world == 1:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baselines(args, torch)
        if not args.no_other_configs:
            line["other_configs"] = other_configs_leg(torch, config, engine, args.config)
    if lm is not None:
        line["logmel"] = lm
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c5", "fork", "sedpy"])
    ap.add_argument("--ref-batch", type=int, default=16, help="CPU-baseline sample batch")
    ap.add_argument("--grad-exchange", default="auto", choices=["auto", "nccl", "p2p"])
    ap.add_argument("--clock-period-ms", type=int, default=20, help="nvidia-smi sampling period during the run")
    ap.add_argument("--no-logmel", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C1 / C5 reference timings (N=1 only)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU-baseline leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl != "reference" and world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    # stdout carries exactly ONE line, the JSON result: anything a library prints there (NCCL's version banner, a
    # stray warning) is sent to stderr by pointing fd 1 at fd 2 for the whole run; the result goes to the saved fd
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global print
    _print = print

    def print(*a, **k):                                    # noqa: A001 -- every result line of this file
        if k.get("file") not in (None, sys.stdout):
            return _print(*a, **k)
        sys.stdout.flush()
        os.write(real_stdout, (" ".join(str(x) for x in a) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
