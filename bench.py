#!/usr/bin/env python3
"""bench.py -- CRNN train frames/sec (BASELINE.json metric) on N B200s, plus the log-mel leg.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one full training step of the hot path on one synthetic batch: forward + loss + backward
+ [gradient exchange] + global-norm clip + Adam.  Workload at N=1 is BASELINE.json configs[1] (binaural
SEDnet, seq_len 256, batch 128); at N>1 the per-GPU batch stays 128 (weak scaling).  `--config c4` is
configs[3] as written: C1 shapes, FIXED global batch 1024 split over the ranks (strong scaling); the
default line also carries that point for its N as `fixed_global_batch`.  Prints ONE JSON line on rank 0.

Legs of the default line (rank 0, besides the timed region): `e2e` (host batches, public API),
`dropin_e2e` (the reference's own interface: TimePooledCRNN / loss.backward() / clip_grad_norm_ /
torch.optim.Adam, and FusedClipAdam), `library_baseline` (stock PyTorch eager -- cuDNN / cuBLAS /
torch.stft -- on the SAME B200: the bar that says whether the hand-written kernels are any good),
`cpu_baseline` (the reference's CPU path, N=1 only), `logmel` (BASELINE configs[2], incl. a real
10 k-clip run with on-device synthesis), `other_configs`.
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
                                    "conv0_win_fwd_kernel (tcgen05 conv with a pooling window per MMA row: BN folded into the "
                                    "weights, max-pool / ReLU / dropout in registers, conv output never stored)"
                                    if cfg.tensor_cores else
                                    "conv0_lean_fwd_kernel (conv + BN + ReLU + max-pool + dropout in registers)")
            m["conv0.bwd_lean"] = ("hbm", xin + a + a / 4, "conv0_lean_bwd_kernel (winner contributions to dW / d gamma / d beta)")
        elif tc_ok:
            # tensor-pipe cost per k-step in fp16-pass units: the forward runs ONE fp16 pass + ONE e4m3 correction pass over
            # twice the K at twice the rate (= 2 units, fp32-grade to ~2^-15), the data / weight gradients ONE fp16 pass
            # (dy scaled by a per-tensor power of two) -- DESIGN.md section 3
            for ph, kn, terms in (("fwd", "conv_tc_kernel", 2), ("dgrad", "conv_tc_kernel", 1), ("wgrad", "wgrad_tc_kernel", 1)):
                m[f"conv{i}.{ph}"] = ("tensor", per[f"conv{i}"], f"{kn} (tcgen05 implicit GEMM: fp16 pass + e4m3 correction "
                                      f"pass forward, single-pass fp16 gradients)", terms)
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
DTYPE = "f32 (fp16 + e4m3-correction tensor-core operands in the conv forward, fp16 single-pass gradients, bf16x3 GRU GEMMs; fp32 accumulate)"


def _ref_model(R, config):
    preset = dict(R.PRESETS["c1" if config == "c4" else config])
    if config == "fork":
        return preset, R.RefCRNN(**preset, dropout=0.4)
    return preset, R.RefCRNN(**{**preset, "dropout": 0.5, "dropout_each_block": True})


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on this box's host cores: the CRNN oracle is the
    same stock torch.nn modules the reference instantiates (crnn_lightning.py:41-73), the trainer step is
    forward + loss + backward + clip + Adam (oracle/crnn_ref.py), at the SAME batch (128) as our arm.
    /root/reference does not exist on the GPU box, so this is cpu_baseline.kind = "port"."""
    if rank != 0:
        return
    import torch
    from oracle import crnn_ref as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    preset, model = _ref_model(R, args.config)
    B = args.ref_batch
    opt = R.make_adam(model, 1e-3, 1e-4)
    x, y = R.synth_batch(preset, B, seed=0)
    t0 = time.perf_counter()
    R.train_step(model, opt, x, y, "bce", 1.0)
    first = time.perf_counter() - t0
    # W warm-up + K timed steps as asked, bounded so that the whole run stays within ~150 s of CPU work
    budget = max(2, int(150.0 / max(first, 1e-3)))
    warm = max(0, min(args.warmup - 1, budget // 4))
    for _ in range(warm):
        R.train_step(model, opt, x, y, "bce", 1.0)
    steps = max(1, min(args.steps, budget - warm))
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        R.train_step(model, opt, x, y, "bce", 1.0)
        ts.append(time.perf_counter() - t0)
    dt = sum(ts) / steps
    T = preset["seq_len"]
    val = B * T / dt
    sample = (f"{1 + warm} warm-up + {steps} timed steps (asked: --warmup {args.warmup} --steps {args.steps}) at batch "
              f"{B} of the {args.config} config -- the same batch as the GPU arm's per-GPU batch; oracle/crnn_ref.py "
              f"(stock torch.nn, the modules the reference instantiates) on torch-CPU fp32, dropout 0.5, "
              f"{torch.get_num_threads()} threads; median step {sorted(ts)[len(ts) // 2] * 1e3:.0f} ms")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1 + warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.gpus), "per_gpu_batch": B, "sample": sample,
                   "note": "one CPU process with all host threads, whatever N is (the reference has no multi-device path)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_name(config, n):
    base = {"c2": "binaural (2-ch, 80-band) SEDnet DCASE2017 task3 CRNN train step, seq_len 256, 128 filters, "
                  "pool [5,2,2] over mel, 2xBiGRU(32), dense 16, 6 classes",
            "c1": "mono SEDnet CRNN train step, seq_len 256, 128 filters, pool [5,2,2], 2xBiGRU(32), 6 classes",
            "c5": "long-context CRNN train step, seq_len 2048, 256 filters, 3xBiGRU(128), 16 classes",
            "c4": "data-parallel mono SEDnet CRNN train step (C1 shapes [1024,1,256,40]), FIXED global batch 1024",
            "fork": "fork-default CRNN (train_constants.py) train step", "sedpy": "sed.py CRNN train step"}[config]
    if config == "c4":
        return f"{base}; batch {1024 // n}/GPU x {n} GPU"
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
        "ms_per_launch": ms,
        "roofline": {"bound": "hbm", "kernel": "logmel_kernel", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                     "frac": gbs / pk["hbm"], "traffic": 528320256 * n_clips / 8, "per_gpu": True,
                     "peak_source": pk["src"],
                     "traffic_source": "profiles/r01_logmel_full.ncu-rep (8 clips: 508.5 MB read + 19.8 MB written), scaled"},
    }
    # the two kernels side by side on the same resident clips (every rank times both; max over ranks): "fp32" = warp per
    # frame, 32 x 32 register FFT on the CUDA cores; "tc" = the DFT as two batched tcgen05 GEMMs on fp16 hi / lo planes.
    # The default (what the lines above ran) is the faster one; profiles/README.md holds the ncu comparison.
    kern = {}
    for name in ("fp32", "tc"):
        for _ in range(2):
            feature.mbe_device(x, out=out, kernel=name)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            feature.mbe_device(x, out=out, kernel=name)
        e1.record()
        torch.cuda.synchronize()
        tk_ = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
        if world > 1:
            dist.all_reduce(tk_, op=dist.ReduceOp.MAX)
        kern[name] = {"ms_per_launch": tk_.item(), "gb_per_s": alg_bytes / tk_.item() / 1e6,
                      "frac_of_hbm": alg_bytes / tk_.item() / 1e6 / pk["hbm"]}
    from sed_crnn_b200 import _lib as _l
    res["kernels"] = kern
    res["default_kernel"] = {1: "fp32", 2: "tc"}[_l.lib().sedb200_logmel_default_kernel()]
    res["roofline"]["kernel"] = "logmel_kernel" if res["default_kernel"] == "fp32" else "logmel_tc_kernel"
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
    # ---- BASELINE configs[2] as written: 10,000 synthetic 3-min stereo clips, sharded by clip index over the ranks
    #      (parallel.clip_ids_for_rank, no collective).  635 GB of PCM does not fit anywhere, so every chunk of 32
    #      clips is SYNTHESISED ON THE DEVICE (seeded per chunk and rank), converted, and reduced to a checksum that
    #      is read back at the end; reported: the sum of the kernel times (CUDA events around each call) and the wall
    #      time of the whole loop including synthesis.
    if not os.environ.get("SEDB200_BENCH_SKIP_10K"):
        from sed_crnn_b200.parallel import clip_ids_for_rank
        mine = len(clip_ids_for_rank(10000, rank, world))
        chunks = [n_clips] * (mine // n_clips) + ([mine % n_clips] if mine % n_clips else [])
        evs = []
        chk = torch.zeros((), device="cuda", dtype=torch.float64)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for ci, nc in enumerate(chunks):
            g.manual_seed(7_000_000 + 10_000 * rank + ci)
            x[:nc].normal_(0, 0.1, generator=g)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            feature.mbe_device(x[:nc], out=out[:nc])
            b.record()
            evs.append((a, b))
            chk += out[:nc, ::97, ::7].sum(dtype=torch.float64)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        tk = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) * 1e-3, wall], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        res["ten_k_clips"] = {
            "clips_total": 10000, "clips_this_rank": mine, "chunks_of": n_clips, "kernel_s": tk[0].item(),
            "wall_s_incl_on_device_synthesis": tk[1].item(), "audio_s_per_s_kernel": 10000 * 180.0 / tk[0].item(),
            "gb_per_s_kernel_per_gpu": mine * (2 * S * 4 + feature.n_frames(S) * 80 * 4) / tk[0].item() / 1e9,
            "checksum_finite": bool(torch.isfinite(chk).item()),
            "note": "max over ranks; synthesis = torch normal_ into the resident 32-clip buffer, seeded per (rank, chunk)"}
    if rank == 0:
        # ---- library bar on the same GPU: torch.stft (cuFFT) + |X|^2 + mel matmul (cuBLAS) + log, stock PyTorch eager
        try:
            fb = torch.from_numpy(feature.mel_filterbank()).cuda()                      # [40, 1025]
            win = torch.hann_window(2048, periodic=True, device="cuda")
            xs = x[:8]                                                                  # 8 clips: the 4 GB complex STFT of 32 would only measure the allocator

            def lib_call():
                sp = torch.stft(xs.reshape(-1, S), 2048, hop_length=1024, window=win, center=True, pad_mode="constant",
                                return_complex=True)                                    # [16, 1025, frames]
                pw = sp.real * sp.real + sp.imag * sp.imag
                return torch.log(torch.matmul(fb, pw)).transpose(1, 2)

            for _ in range(2):
                lib_call()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                lib_call()
            e1.record()
            torch.cuda.synchronize()
            lms = e0.elapsed_time(e1) / 5
            e0.record()
            for _ in range(5):
                feature.mbe_device(xs, out=out[:8])
            e1.record()
            torch.cuda.synchronize()
            oms = e0.elapsed_time(e1) / 5
            res["library_baseline"] = {
                "what": "stock PyTorch eager on the same B200: torch.stft (cuFFT, fp32) + power + mel matmul (cuBLAS) + "
                        "log, 8 resident 3-min stereo clips", "ms": lms, "audio_s_per_s": 8 * 180.0 / (lms * 1e-3),
                "ours_same_input_ms": oms, "vs_library": lms / oms}
        except Exception as e:                                   # noqa: BLE001 -- informational leg
            res["library_baseline"] = {"error": str(e)[:200]}
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
            eng = engine.CRNNEngine(cfg, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0, seed=99, cuda_graph=True)
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


def _pool_logmel(n):
    """worker of the process-pool log-mel baseline: one 3-min mono clip per task"""
    import numpy as np
    from oracle import logmel_ref
    clip = logmel_ref.synth_clip(n, 180 * 44100, 1, "noise")[0]
    t0 = time.perf_counter()
    out = logmel_ref.mbe(clip)
    return time.perf_counter() - t0, float(np.isfinite(out).mean())


def cpu_baselines(args, torch):
    """Reference CPU path beside the GPU number (rank 0, N=1 only; bounded samples): the CRNN training step at the
    SAME batch as the GPU arm, and the log-mel path both serial (the reference loop is serial, feature.py:70) and as a
    process pool over all host cores."""
    from oracle import crnn_ref as R, logmel_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    preset, model = _ref_model(R, args.config)
    b = args.ref_batch
    opt = R.make_adam(model, 1e-3, 1e-4)
    x, y = R.synth_batch(preset, b, seed=0)
    R.train_step(model, opt, x, y, "bce", 1.0)
    ts = []
    while len(ts) < 5 and (sum(ts) < 20.0 or len(ts) < 3):
        t0 = time.perf_counter()
        R.train_step(model, opt, x, y, "bce", 1.0)
        ts.append(time.perf_counter() - t0)
    dt = sorted(ts)[len(ts) // 2]
    val = b * preset["seq_len"] / dt
    clip = logmel_ref.synth_clip(0, 180 * 44100, 1, "noise")[0]
    logmel_ref.mbe(clip[:44100])
    t0 = time.perf_counter()
    logmel_ref.mbe(clip)
    lm = 180.0 / (time.perf_counter() - t0)
    pool = None
    try:
        import multiprocessing as mp
        n_tasks = 2 * cores
        with mp.get_context("spawn").Pool(cores) as pl:
            pl.map(_pool_logmel, range(cores))                       # warm the workers (imports, FFT plans)
            t0 = time.perf_counter()
            r = pl.map(_pool_logmel, range(n_tasks))
            wall = time.perf_counter() - t0
        pool = {"audio_s_per_s": n_tasks * 180.0 / wall, "processes": cores, "clips": n_tasks,
                "mean_clip_s": sum(t for t, _ in r) / n_tasks}
    except Exception as e:                                           # noqa: BLE001
        pool = {"error": str(e)[:200]}
    return {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/crnn_ref.py (stock torch.nn, the modules the reference instantiates) on torch-CPU fp32, "
                      f"{torch.get_num_threads()} threads: 1 warm-up + {len(ts)} timed steps (median) at batch {b} of the "
                      f"same config = the GPU arm's batch; log-mel oracle (numpy/scipy float64 FFT, librosa semantics), "
                      f"3-min mono clips: one process, and a pool of {cores} processes x {2 * cores} clips",
            "ms_per_step": dt * 1e3, "logmel_audio_s_per_s": lm, "logmel_process_pool": pool}


# ------------------------------------------------------------------------------------------ library bar (same GPU)
def _eager_crnn(torch, cfg):
    """The reference's network from stock torch.nn layers (what crnn_lightning.py:41-73 / sed.py:82-112 instantiate,
    generalised over the constants like config.CRNNConfig), to be run by PyTorch eager on the GPU: cuDNN convolutions /
    BatchNorm / GRU, cuBLAS dense layers.  Defined here (not imported from oracle/) because it is a measured bar."""
    nn = torch.nn

    class Eager(nn.Module):
        def __init__(self):
            super().__init__()
            self.convs, self.bns, self.pools = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
            c = cfg.in_ch
            for p in cfg.pool:
                self.convs.append(nn.Conv2d(c, cfg.conv_ch, 3, padding=1))
                self.bns.append(nn.BatchNorm2d(cfg.conv_ch))
                self.pools.append(nn.MaxPool2d((1, p)))
                c = cfg.conv_ch
            self.drop = nn.Dropout(cfg.dropout)
            self.grus = nn.ModuleList()
            d = cfg.flat
            for h in cfg.gru_units:
                self.grus.append(nn.GRU(d, h, bidirectional=True, batch_first=True))
                d = 2 * h
            self.denses = nn.ModuleList()
            for u in list(cfg.dense_units) + [cfg.n_classes]:
                self.denses.append(nn.Linear(d, u))
                d = u

        def forward(self, x):
            for i, (c, b, p) in enumerate(zip(self.convs, self.bns, self.pools)):
                x = p(torch.relu(b(c(x))))
                if cfg.dropout_each_block or i == len(self.convs) - 1:
                    x = self.drop(x)
            x = x.permute(0, 3, 1, 2) if cfg.mode == "fork" else x.permute(0, 2, 1, 3)
            b_, t_, c_, f_ = x.shape
            x = x.reshape(b_, t_, c_ * f_)
            for g in self.grus:
                x, _ = g(x)
            for i, d in enumerate(self.denses):
                x = d(x)
                if i < len(self.denses) - 1 and cfg.dense_relu:
                    x = torch.relu(x)
            return x

    return Eager()


def library_baseline_leg(torch, config, names, ours_ms):
    """Stock PyTorch eager on the same B200 (the reference's only GPU path: sed.py:42,133,
    train_lightning.py:15,45): forward + BCE-with-logits + backward + clip_grad_norm_(1.0) + torch.optim.Adam(fused)
    on device-resident batches, CUDA events, 3 warm-up + 10 timed steps.  Two precision settings: the reference's own
    (`set_float32_matmul_precision('medium')`, cuDNN TF32 convolutions allowed -- LOWER precision than this repo's
    path) and strict fp32 (`'highest'`, TF32 off -- the same numerical class as the 3-term split)."""
    out = {}
    for name in names:
        cfg = config.PRESETS[name]
        B = PER_GPU_BATCH
        res = {"workload": workload_name(name, 1)}
        for tag, prec, tf32 in (("reference_setting_tf32_medium", "medium", True), ("strict_fp32", "highest", False)):
            try:
                torch.set_float32_matmul_precision(prec)
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.benchmark = True
                torch.manual_seed(0)
                m = _eager_crnn(torch, cfg).cuda().train()
                opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-4, fused=True)
                g = torch.Generator(device="cuda").manual_seed(5)
                x = torch.randn(cfg.input_shape(B), device="cuda", generator=g)
                y = (torch.rand(cfg.target_shape(B), device="cuda", generator=g) < 0.2).float()
                lossf = torch.nn.BCEWithLogitsLoss()

                def step():
                    opt.zero_grad(set_to_none=True)
                    loss = lossf(m(x), y)
                    loss.backward()
                    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
                    opt.step()

                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                n = 10 if name != "c5" else 4
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                res[tag] = {"ms_per_step": ms, "frames_per_s": B * cfg.seq_len / (ms * 1e-3)}
                if ours_ms.get(name):
                    res[tag]["vs_library"] = ms / ours_ms[name]
                del m, opt, x, y
                torch.cuda.empty_cache()
            except Exception as e:                              # noqa: BLE001 -- informational leg
                res[tag] = {"error": str(e)[:200]}
        out[name] = res
    torch.set_float32_matmul_precision("highest")
    return out


def dropin_e2e_leg(torch, steps=20):
    """End to end through the REFERENCE'S OWN interface: the model is called as `model(x)`, the loss as
    `loss_fn(logits, y)`, then `loss.backward()`, `clip_grad_norm_`, `optimizer.step()` -- what Lightning's fit loop
    does (crnn_lightning.py:157-163, train_lightning.py:50, crnn_lightning.py:195-197) and sed.run_epoch does
    (sed.py:134-138).  Host batches, pinned, one H2D per step, loss read every step (sed.py:138)."""
    from sed_crnn_b200 import config, crnn_lightning, modules
    out = {}
    for name, make, loss_cls, shape_cfg in (
            ("fork_lightning_model", lambda: crnn_lightning.TimePooledCRNN(0.4), crnn_lightning.FocalBCELoss, config.FORK),
            ("c2_model", lambda: modules.CRNN(config.C2), modules.BCEWithLogitsLoss, config.C2)):
        B = PER_GPU_BATCH
        gx = torch.Generator().manual_seed(3)
        xs = [torch.randn(shape_cfg.input_shape(B), generator=gx).pin_memory() for _ in range(4)]
        ys = [(torch.rand(shape_cfg.target_shape(B), generator=gx) < 0.2).float().pin_memory() for _ in range(4)]
        res = {}
        for opt_name in ("torch_adam_clip_grad_norm", "fused_clip_adam"):
            try:
                torch.manual_seed(0)
                m = make().cuda().train()
                lossf = loss_cls()
                if opt_name == "fused_clip_adam":
                    opt = modules.FusedClipAdam(m, lr=1e-3, weight_decay=1e-4, max_norm=1.0)
                else:
                    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-4)

                def step(i):
                    x = xs[i % 4].cuda(non_blocking=True)
                    y = ys[i % 4].cuda(non_blocking=True)
                    opt.zero_grad()
                    loss = lossf(m(x), y)
                    loss.backward()
                    if opt_name != "fused_clip_adam":
                        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
                    opt.step()
                    return loss.item()

                for i in range(3):
                    step(i)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for i in range(steps):
                    step(i)
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) / steps * 1e3
                res[opt_name] = {"ms_per_step": ms, "frames_per_s": B * shape_cfg.seq_len / (ms * 1e-3)}
                del m, opt
            except Exception as e:                              # noqa: BLE001
                res[opt_name] = {"error": str(e)[:200]}
        res["h2d_bytes_per_step"] = (xs[0].numel() + ys[0].numel()) * 4
        res["d2h_bytes_per_step"] = 4
        out[name] = res
    torch.cuda.empty_cache()
    return out


def pipe_floor(issued_flops, ms, sm_mhz=1965.0, n_sm=148):
    """Tensor-pipe floor of a tcgen05 launch: issued fp16-pass FLOPs at 8192 dense FLOP / clk / SM."""
    floor_ms = issued_flops / (n_sm * 8192.0 * sm_mhz * 1e6) * 1e3
    return {"pipe_floor_ms": floor_ms, "frac_of_pipe_floor": floor_ms / ms if ms > 0 else None}


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
    strong = args.config == "c4"
    cfg = config.PRESETS["c1" if strong else args.config]
    if strong and 1024 % world:
        raise SystemExit("--config c4 splits a global batch of 1024: N must divide it")
    B = 1024 // world if strong else PER_GPU_BATCH
    # gradient exchange: the fused NVLink peer-memory kernel when there is more than one rank ("auto"); if its
    # set-up fails (agreed on by all ranks) the run uses NCCL and SAYS so in config.grad_exchange
    gx_mode, gx_note = ("nccl" if world == 1 else "p2p") if args.grad_exchange == "auto" else args.grad_exchange, None
    use_graph = not args.no_cuda_graph and (world == 1 or gx_mode == "p2p")
    try:
        eng = engine.CRNNEngine(cfg, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0, seed=1234 + rank,
                                grad_exchange=gx_mode, cuda_graph=use_graph)
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
    n_warm = max(args.warmup, 3)
    for i in range(n_warm):
        eng.train_step(xs[i % n_in], ys[i % n_in])
    if use_graph:                                    # untimed: one more pass so that every rotated batch has its graph
        for i in range(n_warm, n_warm + n_in + 1):
            eng.train_step(xs[i % n_in], ys[i % n_in])
    barrier()

    # ---- timed region: device-resident inputs.  A block = EXACTLY args.steps steps between barriers, CUDA events,
    #      max over ranks; blocks are repeated until at least 0.5 s has been timed (a 36 ms region is valid but thin) and
    #      the MEDIAN block is the reported one -- every rank runs the same number of blocks (the count is agreed on)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_block():
        l0 = L.sedb200_launch_count() + eng.graph_replays * eng.launches_per_graph_step
        barrier()
        e0.record()
        for i in range(args.steps):
            loss_, _ = eng.train_step(xs[i % n_in], ys[i % n_in])
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), L.sedb200_launch_count() + eng.graph_replays * eng.launches_per_graph_step - l0, loss_

    t_clk0 = sampler.mark()
    first_ms, launches, loss = timed_block()
    n_blocks = max(1, min(200, int(-(-args.min_timed_s * 1e3 // max(first_ms, 1e-3)))))
    block_ms = [first_ms]
    for _ in range(n_blocks - 1):
        bm, launches, loss = timed_block()
        block_ms.append(bm)
    ms_step = sorted(block_ms)[len(block_ms) // 2] / args.steps
    frames_per_step = B * cfg.seq_len * world
    value = frames_per_step / (ms_step * 1e-3)
    final_loss = loss.item()
    if eng.xch is not None:
        eng.xch.raise_if_failed(wait=True)                             # a skipped exchange step must not produce a number

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
    e2e_steps = args.steps * max(1, min(n_blocks, 10))        # same >= 0.5 s idea, one pipeline

    def e2e_run(sync_every_step):
        pf = DevicePrefetcher(host_batches(e2e_warm + e2e_steps))
        for n_done, (xd, yd, k) in enumerate(pf):
            if n_done == e2e_warm:
                barrier()
                e0.record()
            l, _ = eng.train_step(xd, yd)
            pf.release(k)
            loss_h.copy_(l.reshape(1), non_blocking=True)
            if sync_every_step:
                torch.cuda.current_stream().synchronize()      # the caller reads the loss every step (sed.py:138)
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        return frames_per_step / (t2.item() / e2e_steps * 1e-3)

    e2e_value = e2e_run(True)
    # Lightning's loop does NOT read the loss every step (crnn_lightning.py:157-163 logs a device tensor): same
    # pipeline, the loss copy stays asynchronous and is read once at the end
    e2e_nosync = e2e_run(False)
    clocks = sampler.stop(t_clk0, sampler.mark()) if rank == 0 else None
    if eng.xch is not None:
        eng.xch.raise_if_failed(wait=True)

    # ---- BASELINE configs[3] as written, for this N: C1 shapes, FIXED global batch 1024 (strong scaling)
    fixed = None
    if not strong and not args.no_fixed_global and 1024 % world == 0:
        try:
            Bf = 1024 // world
            cfg4 = config.PRESETS["c1"]
            eng4 = engine.CRNNEngine(cfg4, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0, seed=4321 + rank,
                                     grad_exchange=gx_mode, cuda_graph=use_graph)
            eng4.init_default(seed=0)
            g4 = torch.Generator(device="cuda").manual_seed(200 + rank)
            x4 = [torch.randn(cfg4.input_shape(Bf), device="cuda", generator=g4) for _ in range(2)]
            y4 = [(torch.rand(cfg4.target_shape(Bf), device="cuda", generator=g4) < 0.2).float() for _ in range(2)]
            for i in range(7):                          # eager first call + one graph per (batch, parity)
                eng4.train_step(x4[i % 2], y4[i % 2])
            n4 = 10
            blocks4 = []
            for _ in range(3):                          # median of three timed blocks (max over ranks each)
                barrier()
                e0.record()
                for i in range(n4):
                    eng4.train_step(x4[i % 2], y4[i % 2])
                e1.record()
                barrier()
                t4 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
                if world > 1:
                    dist.all_reduce(t4, op=dist.ReduceOp.MAX)
                blocks4.append(t4.item())
            t4 = torch.tensor([sorted(blocks4)[1]], device="cuda")
            if eng4.xch is not None:
                eng4.xch.raise_if_failed(wait=True)
                eng4.xch.close()
            ms4 = t4.item() / n4
            fixed = {"workload": workload_name("c4", world), "global_batch": 1024, "per_gpu_batch": Bf,
                     "scaling": "strong", "ms_per_step": ms4, "frames_per_s": 1024 * cfg4.seq_len / (ms4 * 1e-3),
                     "steps": n4}
            del eng4, x4, y4
            torch.cuda.empty_cache()
        except Exception as e:                                  # noqa: BLE001 -- informational leg
            fixed = {"error": str(e)[:200]}

    # ---- BASELINE configs[4] at N > 1: the long-context model data-parallel, batch 128 per GPU (weak scaling; the
    #      N = 1 point is `other_configs.c5` of the single-GPU line)
    long_ctx = None
    if world > 1 and not strong and args.config != "c5" and not args.no_other_configs:
        try:
            cfg5 = config.PRESETS["c5"]
            eng5 = engine.CRNNEngine(cfg5, loss="bce", lr=1e-3, weight_decay=1e-4, clip=1.0, seed=777 + rank,
                                     grad_exchange=gx_mode, cuda_graph=use_graph)
            eng5.init_default(seed=0)
            g5 = torch.Generator(device="cuda").manual_seed(300 + rank)
            x5 = [torch.randn(cfg5.input_shape(PER_GPU_BATCH), device="cuda", generator=g5) for _ in range(2)]
            y5 = [(torch.rand(cfg5.target_shape(PER_GPU_BATCH), device="cuda", generator=g5) < 0.2).float() for _ in range(2)]
            for i in range(5):
                eng5.train_step(x5[i % 2], y5[i % 2])
            n5 = 5
            barrier()
            e0.record()
            for i in range(n5):
                eng5.train_step(x5[i % 2], y5[i % 2])
            e1.record()
            barrier()
            t5 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
            if eng5.xch is not None:
                eng5.xch.raise_if_failed(wait=True)
                eng5.xch.close()
            ms5 = t5.item() / n5
            long_ctx = {"workload": workload_name("c5", world), "per_gpu_batch": PER_GPU_BATCH, "scaling": "weak",
                        "ms_per_step": ms5, "frames_per_s": world * PER_GPU_BATCH * cfg5.seq_len / (ms5 * 1e-3), "steps": n5}
            del eng5, x5, y5
            torch.cuda.empty_cache()
        except Exception as e:                                  # noqa: BLE001 -- informational leg
            long_ctx = {"error": str(e)[:200]}

    # ---- phase breakdown (extra instrumented steps, CUDA events on the launching stream)
    eng.cuda_graph = False                              # the phase profiler brackets individual launches: eager steps
    for i in range(5):                                  # eager warm-up (the timed region above ran graph replays)
        eng.train_step(xs[i % n_in], ys[i % n_in])
    torch.cuda.synchronize()
    prof_steps, prof_rounds, rounds = 4, 5, []
    for r in range(prof_rounds):                        # 5 rounds of 4 instrumented steps; per phase the MEDIAN round
        L.sedb200_prof_enable(1)
        for i in range(prof_steps):
            eng.train_step(xs[i % n_in], ys[i % n_in])
        torch.cuda.synchronize()
        rounds.append(parse_prof(L))
        L.sedb200_prof_enable(0)
    prof = {k: (sorted(rd[k][0] for rd in rounds)[prof_rounds // 2], rounds[0][k][1]) for k in rounds[0]}

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
        issued = sum(model[k][1] * (model[k][3] if len(model[k]) > 3 else 1) for k in dks) / (ms_kernel_step * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": kname, "phases": dks, "achieved": ach, "peak": pk["tf_sustained"],
                "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"], "traffic": traffic,
                "traffic_note": "dram bytes per launch, mean of this kernel's launches in one step as captured by ncu "
                "--set full (profiles/ncu_traffic.json)",
                "peak_source": pk["src"] + " (sustained bf16)", "algorithmic_flops_per_launch": amount,
                "mma_tflops_issued": issued, "frac_issued": issued / pk["tf_sustained"],
                "per_phase": {k: {"ms": phases[k], "algorithmic_tflops": model[k][1] / (phases[k] * 1e-3) / 1e12,
                                  "mma_passes": model[k][3] if len(model[k]) > 3 else 1,
                                  **pipe_floor(model[k][1] * (model[k][3] if len(model[k]) > 3 else 1), phases[k])}
                              for k in dks},
                "tensor_pipe": {**pipe_floor(sum(model[k][1] * (model[k][3] if len(model[k]) > 3 else 1) for k in dks),
                                             ms_kernel_step),
                                "note": "a 128x128x16 fp16 (or 128x128x32 e4m3) tcgen05.mma occupies an SM's tensor pipe for 64 "
                                "cycles (8192 dense FLOP per cycle and SM; ncu's sm__pipe_tensor_subpipe_hmma_cycles_active "
                                "counts exactly that: profiles/README.md); floor_ms = issued pass-FLOPs / (148 SMs x 8192 x "
                                "sm_max_mhz), frac_of_floor = floor_ms / measured ms = tensor-pipe busy fraction if the SM "
                                "clock held its maximum inside the launch"},
                "note": "frac = algorithmic FLOPs (2*M*K*N, SURVEY 8d) / measured sustained fp16/bf16 peak, pooled over this "
                "kernel's launches in a step; the forward launches run one fp16 pass plus one e4m3 correction pass (twice the K "
                "at twice the rate: 2 fp16-pass units per k-step), the data-gradient launches one pass; frac_issued counts "
                "the tensor-pipe work actually issued in fp16-pass units"}
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
                                 "mma_tflops_issued": (model[k][3] if len(model[k]) > 3 else 1) * a}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": workload_name(args.config, world), "per_gpu_batch": B, "global_batch": B * world,
                   "seq_len": cfg.seq_len, "parallelism": f"dp{world}", "l2": "per-step working set (>1 GB of "
                   "activations) exceeds the 126 MB L2; 4 input batches rotated",
                   "cuda_graph": (f"the step is ONE CUDA-graph launch ({eng.launches_per_graph_step} kernels per replay, "
                                  f"{eng.graph_replays} replays so far); seed / Adam step live in a device-side step state"
                                  if eng.graph_replays else "off (eager launches)"),
                   "timing": f"{len(block_ms)} blocks of {args.steps} steps (>= {args.min_timed_s} s timed), median block; "
                             f"block ms min/median/max = {min(block_ms):.3f}/{sorted(block_ms)[len(block_ms) // 2]:.3f}/"
                             f"{max(block_ms):.3f}", "loss": "bce", "optimizer":
                   "clip 1.0 + Adam(1e-3, wd 1e-4)", "dropout": cfg.dropout, "final_loss": final_loss,
                   "grad_exchange": {"nccl": "none (1 GPU)" if world == 1 else "NCCL all-reduce + clip/Adam kernels",
                                     "p2p": "one fused kernel: NVLink peer-memory all-reduce + clip + Adam"}[gx_mode],
                   **({"grad_exchange_note": gx_note} if gx_note else {})},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": (xs_h[0].numel() + ys_h[0].numel()) * 4 * world, "d2h_bytes_per_step": 4 * world,
                "steps": e2e_steps, "loop": "sed.run_epoch style: pinned host batch -> device every step, loss read on the "
                "host every step (sed.py:133-138)",
                "lightning_style_value": e2e_nosync, "lightning_style_loop": "same copies, loss left on the device and "
                "read once at the end (crnn_lightning.py:157-163 never reads it per step)"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "phases_ms": {k: round(v, 4) for k, v in phases.items()},
    }
    if fixed is not None:
        line["fixed_global_batch"] = fixed
    if long_ctx is not None:
        line["long_context"] = long_ctx
    if world == 1:
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baselines(args, torch)
        ours_ms = {args.config if not strong else "c1": ms_step}
        if not args.no_other_configs:
            line["other_configs"] = other_configs_leg(torch, config, engine, args.config)
            ours_ms.update({k: v["ms_per_step"] for k, v in line["other_configs"].items() if "ms_per_step" in v})
        if not args.no_library_baseline:
            names = [n for n in ("c2", "c1", "c5", "fork") if n in ours_ms]
            line["library_baseline"] = library_baseline_leg(torch, config, names, ours_ms)
            main_name = "c1" if strong else args.config
            lb = line["library_baseline"].get(main_name, {}).get("strict_fp32", {})
            line["vs_library"] = lb.get("vs_library")
        if not args.no_dropin:
            line["dropin_e2e"] = dropin_e2e_leg(torch)
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
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c4", "c5", "fork", "sedpy"])
    ap.add_argument("--ref-batch", type=int, default=PER_GPU_BATCH, help="CPU-baseline batch (default: the GPU arm's)")
    ap.add_argument("--min-timed-s", type=float, default=0.5, help="repeat the --steps block until this much is timed")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip the PyTorch-eager-on-GPU bar")
    ap.add_argument("--no-dropin", action="store_true", help="skip the drop-in-interface end-to-end leg")
    ap.add_argument("--no-cuda-graph", action="store_true", help="launch every kernel of the step eagerly (N = 1)")
    ap.add_argument("--no-fixed-global", action="store_true", help="skip the fixed-global-batch-1024 (configs[3]) point")
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
