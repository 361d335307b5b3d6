#!/usr/bin/env python3
"""bench.py -- headline benchmark of the streaming hot path (BASELINE.json metric: RTFx = audio seconds
processed per wall second, summed over streams; per-chunk latency alongside).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N = 1 workload = BASELINE.json configs[1]: nemotron-speech-streaming-en-0.6b (24 layers, random-init synthetic
weights of that architecture), bf16 compute, 64 concurrent streams, 160 ms chunks (att_right_context = 1).
A "step" = one batched engine step = one 160 ms chunk for every stream (64 x 2 encoder frames).
N > 1 (torchrun, one rank per GPU): streams are independent, so every rank runs its own 64 streams on its own
engine with no data-path collective ("weak" scaling); torch.distributed is used only for the barrier and the
max-over-ranks of the timed region.

  value : device-resident throughput -- PCM of the step already in HBM, the K steps enqueued back to back on the engine's
          stream, CUDA events around them (p50 / p99 chunk latency: single host-synchronised steps, outside the timed region)
  e2e   : the same K steps through the public C ABI with HOST buffers (two steps in flight): nsb_push_pcm_batch, nsb_engine_step_begin
          (pinned staging + H2D + kernels + D2H of the token ids enqueued), the next chunk pushed meanwhile,
          nsb_engine_step_end, nsb_pop_tokens_batch; wall clock

--impl reference times the CPU oracle port of the reference path (oracle/liboracle.so, all host threads) on a
bounded sample of the same workload; the reference's own ggml build cannot be produced offline.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

N_LAYERS = int(os.environ.get("NSB_BENCH_LAYERS", 24))
STREAMS = int(os.environ.get("NSB_BENCH_STREAMS", 64))
RIGHT_CONTEXT = int(os.environ.get("NSB_BENCH_R", 1))
PROFILE = os.environ.get("NSB_BENCH_PROFILE", "speech")   # synthetic-model calibration: "speech" = ~4.5 tokens per audio second, "parity" = the tests' dense emission
WARM_CHUNKS = 40            # > 70/T: the 70-frame attention cache is full (steady state) before timing
T = 1 + RIGHT_CONTEXT
CHUNK_S = 0.08 * T


# BASELINE.json configs by number (configs[0] is the CPU case = --impl reference). The driver's default run is config 2;
# the others are selected with --config N for the per-config evidence under profiles/ (same timing legs, same JSON line).
#        streams/GPU, att_right_context, GGUF type, compute, K/V ring
CONFIGS = {2: (64, 1, "f16", "bf16", "bf16"),
           3: (256, 6, "q8_0", "q8_0", "f16"),
           4: (128, 0, "f16", "bf16", "bf16"),
           5: (64, 13, "f16", "bf16", "bf16")}
WEIGHTS, COMPUTE, KV = "f16", "bf16", "bf16"


def select_config(n: int):
    global STREAMS, RIGHT_CONTEXT, T, CHUNK_S, WEIGHTS, COMPUTE, KV, WARM_CHUNKS
    STREAMS, RIGHT_CONTEXT, WEIGHTS, COMPUTE, KV = CONFIGS[n]
    T = 1 + RIGHT_CONTEXT
    CHUNK_S = 0.08 * T
    WARM_CHUNKS = max(8, 70 // T + 3)                  # > 70 / T: the attention cache is full before timing


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        # the GEMM roofline times the kernel ALONE (back-to-back launches): the burst bf16 figure is its denominator, not the sustained one
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", d.get("bf16_tflops_sustained", 1400.0))), "measured"
    return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe). NVML polled every 5 ms from a thread
    (the timed region is ~0.1 s: `nvidia-smi -lms 100` would give one sample); nvidia-smi is the fallback."""

    def __init__(self, device: int):
        self.rows, self.proc, self.device = [], None, device
        self.sm, self.mx, self.reason_bits, self.stop_flag, self.thread, self.nvml = [], None, 0, False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml = (pynvml, h)

            def poll():
                while not self.stop_flag:
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        self.reason_bits |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                    except Exception:
                        pass
                    time.sleep(0.005)
            self.thread = threading.Thread(target=poll, daemon=True); self.thread.start()
            return
        except Exception:
            self.nvml = None
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            pynvml = self.nvml[0]
            bits = {"hw_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(pynvml, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(pynvml, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            reasons = sorted(k for k, b in bits.items() if self.reason_bits & b)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                    "samples": len(self.sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvidia-smi"}


def gemm_algorithmic_bytes(rows: int, elem: int = 2):
    """Per engine step: bytes every layer GEMM must move once (weights + A operand + C result), and launch count."""
    shapes = [(4096, 1024, elem), (1024, 4096, 4), (3072, 1024, 4), (1024, 1024, 4), (2048, 1024, 4), (1024, 1024, 4),
              (4096, 1024, elem), (1024, 4096, 4)]                    # (N, K, bytes per C element); RESID epilogues read+write f32
    total = 0
    for n, k, cb in shapes:
        c_bytes = rows * n * cb * (2 if (n == 1024) else 1)
        total += n * k * elem + rows * k * elem + c_bytes
    return total * N_LAYERS, len(shapes) * N_LAYERS


def gemm_roofline(eng, rows: int, elem: int = 2, iters: int = 10, w_elem: float | None = None):
    """Dominant kernel = the tcgen05 layer GEMM (8 launches per conformer layer). Timed live, alone: for each of the six
    distinct shapes, 24 layers x `iters` launches back to back on the engine's stream (every launch on a different layer's
    weights: 24 x 2-8 MB > L2, so weights stream from HBM), CUDA events around the loop; same tile / split-K choice as the
    step. achieved = algorithmic bytes (weights + A operand + C result, once each) / mean launch duration, weighted by how
    often each shape occurs in a layer."""
    # kind: (N, K, bytes per C element incl. read-modify-write, occurrences per layer)
    kinds = {0: (4096, 1024, elem, 2), 1: (1024, 4096, 8, 2), 2: (3072, 1024, 4, 1), 3: (1024, 1024, 8, 1), 4: (2048, 1024, 4, 1), 5: (1024, 1024, 8, 1)}
    w_elem = elem if w_elem is None else w_elem       # HBM bytes per weight (Q8_0 blocks: 34 / 32)
    tot_us = tot_bytes = tot_flops = 0.0; n = 0; per = {}
    for kind, (N, K, cb, occ) in kinds.items():
        splits = 1
        if N == 1024 and rows <= 1024:                        # Engine::gemm_residual's split-K rule
            tiles = ((rows + 127) // 128) * (N // (64 if K >= 4096 else 32)); nk = K // 64
            while splits < 8 and tiles * splits < 120 and nk % (splits * 2) == 0 and nk // (splits * 2) >= 2:
                splits *= 2
        elif N == 1024 and K >= 4096 and rows > 1024 and w_elem == elem:
            splits = 4                                            # Engine::gemm_residual: FFN down-projection on 4 K slices of 256 x 256 pair tiles
        us = eng.bench_gemm(kind, rows, 0, 0, splits, 0, iters)
        b = N * K * w_elem + rows * K * elem + rows * N * cb
        per[("ff_up", "ff_down", "qkv", "attn_out", "pw1", "pw2")[kind]] = {"us": round(us, 2), "GBps": round(b / us / 1e3, 1),
                                                                           "TFLOPs": round(2.0 * rows * N * K / us / 1e6, 1), "splits": splits}
        tot_us += us * occ; tot_bytes += b * occ; tot_flops += 2.0 * rows * N * K * occ; n += occ
    return tot_bytes / n, tot_us / n, per, tot_flops / n


def cpu_baseline(threads: int | None, seconds: float = 1.6, streams: int = 2):
    """Oracle port (reference arithmetic restated, OpenMP over output rows) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    import synth
    if not threads:                                   # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1: ignore it)
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    O.lib().orc_set_threads(threads)
    cores = threads
    path = synth.cached_model("f32", N_LAYERS, R=RIGHT_CONTEXT, profile=PROFILE)
    m = O.Model(path, O.MM_REF)
    pcm = [synth.synth_pcm(s, seconds) for s in range(streams)]
    t0 = time.perf_counter()
    chunks = 0
    for s in range(streams):
        st = O.Stream(m, RIGHT_CONTEXT)
        st.push(pcm[s])
        chunks += st.chunks
    dt = time.perf_counter() - t0
    audio = chunks * CHUNK_S
    return {"value": audio / dt, "unit": "audio_s/s", "cores": cores, "kind": "port",
            "sample": f"{streams} streams x {seconds:.1f} s synthetic PCM, f32 weights, {N_LAYERS} layers, R={RIGHT_CONTEXT} "
                      f"({chunks} chunks, {dt:.1f} s wall); oracle/liboracle.so = CPU restatement of the reference streaming path "
                      f"(ggml build not reproducible offline)"}, dt


def run_reference(args, rank: int):
    if rank != 0:
        return
    base, dt = cpu_baseline(None, seconds=2.4, streams=2)
    line = {"impl": "reference", "metric": "rtfx", "value": base["value"], "unit": "audio_s/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * STREAMS * CHUNK_S / base["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": f"0.6B streaming FastConformer RNN-T, {N_LAYERS} layers, 160 ms chunks (R={RIGHT_CONTEXT}), "
                                                        f"bounded CPU sample of the {STREAMS}-stream workload"},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": "audio_s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json config number (default 2 = the headline)")
    args = ap.parse_args()
    if args.config != 2:
        select_config(args.config)
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import nsb200
    import synth
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    hbm_peak, tf_peak, peak_kind = peaks()

    # synthetic model (f16 GGUF of the 24-layer architecture, cast to bf16 at load) + per-stream synthetic PCM
    if rank == 0 or world == 1:
        path = synth.cached_model(WEIGHTS, N_LAYERS, R=RIGHT_CONTEXT, profile=PROFILE)
    if world > 1:
        dist.barrier()
        path = synth.cached_model(WEIGHTS, N_LAYERS, R=RIGHT_CONTEXT, profile=PROFILE)
    compute = {"f16": nsb200.COMPUTE_F16, "bf16": nsb200.COMPUTE_BF16, "q8_0": nsb200.COMPUTE_Q8_0}[COMPUTE]
    kv = {"f16": nsb200.KV_F16, "bf16": nsb200.KV_BF16}[KV]
    eng = nsb200.Engine(path, right_context=RIGHT_CONTEXT, max_streams=STREAMS, compute=compute, kv_dtype=kv, device=local,
                        cuda_graph=os.environ.get("NSB_BENCH_GRAPH", "1") != "0")
    shift = eng.shift_samples
    BENCH_CHUNKS = 8                                       # distinct chunks staged in HBM; the timed steps cycle through them
    need = 160 * (8 * T * (WARM_CHUNKS + BENCH_CHUNKS) - 1) + 256
    base = [synth.synth_pcm(1000 * rank + s, need / 16000.0 + 0.01)[:need] for s in range(8)]
    pcm = np.stack([np.roll(base[s % 8], 977 * (s // 8)) for s in range(STREAMS)])      # 64 distinct streams from 8 seeds
    eng.bench_prepare(pcm, WARM_CHUNKS)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing: K steps, CUDA events inside the engine per step (engine stream) ----
    for _ in range(max(args.warmup, BENCH_CHUNKS)):       # >= W untimed steps; at least one pass over every staged chunk (one CUDA graph each)
        eng.bench_step()
    st0 = eng.stats()
    sampler = ClockSampler(local); sampler.start()
    sync_all()
    t0 = time.perf_counter()
    dev_ms, per_step = eng.bench_steps(args.steps)        # K steps enqueued back to back, CUDA events on the engine stream
    sync_all()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    # per-chunk latency: one step at a time, host-synchronised (what a caller waiting for this chunk's tokens sees on the device)
    st1 = eng.stats()
    launches = int(st1.kernel_launches - st0.kernel_launches)
    lat = [eng.bench_step() for _ in range(min(args.steps, 30))]
    # device time of the K steps = first event -> last event on the engine stream; max over ranks
    t_dev = torch.tensor([dev_ms / 1e3, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    t_dev_s, wall_s = float(t_dev[0]), float(t_dev[1])
    audio_s = world * STREAMS * CHUNK_S * args.steps
    value = audio_s / t_dev_s

    # ---- per-kernel-class breakdown + roofline of the dominant kernel (layer GEMMs) ----
    prof, prof_total = eng.bench_profile()
    g_ms, g_n = prof["layer_gemm"]
    rows = STREAMS * T
    b_launch, us_launch, per_shape, f_launch = gemm_roofline(eng, rows, w_elem=34.0 / 32.0 if COMPUTE == "q8_0" else None)
    achieved = b_launch / us_launch / 1e3
    roofline = {"bound": "hbm", "kernel": f"gemm_tc_kernel (tcgen05 layer GEMMs, 8 launches per conformer layer, M = {rows} token rows)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch, occurrence-weighted over the 8 launches of a layer, from the
                # ncu --set full capture profiles/r01_ncu_full_gemm_summary.txt (A / C stay in L2, so it sits below the algorithmic bytes)
                "traffic": 6.53e6, "traffic_source": "profiles/r01_ncu_full_gemm_summary.txt",
                "peak_kind": peak_kind, "share_of_step": g_ms / prof_total if prof_total else None,
                "algorithmic_bytes_per_launch": b_launch, "avg_launch_us": us_launch, "per_shape": per_shape,
                "how": "CUDA events on the engine stream around 24 layers x 10 back-to-back launches per shape (kernel timed alone, weights from HBM)",
                "note": "at 128 token rows every CTA re-reads the whole activation tile: the kernel is bound by per-SM L2->SM ingest (~75 GB/s per SM measured), not by HBM or the tensor pipe; DESIGN.md section 5"}
    if rows >= 512:
        # configs 3 and 5: past the ridge (~208 rows at 2-byte weights) the layer GEMMs are bound by the tensor pipe, not by HBM
        tf = f_launch / us_launch / 1e6
        roofline.update({"bound": "tensor", "achieved": tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": tf / tf_peak, "traffic": None,
                         "traffic_source": None, "algorithmic_flops_per_launch": f_launch, "peak_kind": peak_kind + " (burst bf16: kernel timed alone)",
                         "note": "mean over the 8 layer GEMMs weighted by occurrence; per-shape TFLOP/s in per_shape (DESIGN.md section 5)"})
    breakdown = {k: {"ms": round(v[0], 4), "launches": v[1]} for k, v in prof.items()}

    # ---- end to end through the public C ABI with host buffers ----
    eng2_streams = list(range(STREAMS))
    for s in eng2_streams:
        eng.reset_stream(s)
    e2e_chunks = args.steps + args.warmup
    need2 = 160 * (8 * T * e2e_chunks - 1) + 256
    pcm2 = np.stack([np.resize(pcm[s], need2) for s in range(STREAMS)])
    first = 160 * (8 * T - 1) + 256
    ids = np.arange(STREAMS, dtype=np.int32)
    pos = 0
    def feed(n):
        nonlocal pos
        eng.push_batch(ids, pcm2[:, pos:pos + n]); pos += n
    feed(first)
    for _ in range(args.warmup):
        assert eng.step() == STREAMS
        feed(shift)
    eng.pop_tokens_batch(ids, 32 * T)
    sync_all()
    t0 = time.perf_counter()
    ntok = 0
    # two steps in flight: while the device runs step i the host hands over the next 160 ms of every stream and stages + enqueues
    # step i+1 (pinned staging, H2D, all kernels, D2H of the token ids), then waits for step i and pops its tokens
    assert eng.step_begin() == STREAMS
    for i in range(args.steps):
        if i + 1 < args.steps:
            feed(shift)
            assert eng.step_begin() == STREAMS
        assert eng.step_end() == STREAMS                  # wait for the OLDEST step in flight, queue its tokens
        _, cnt = eng.pop_tokens_batch(ids, 32 * T)
        ntok += int(cnt.sum())
    sync_all()
    e2e_wall = time.perf_counter() - t0
    t_e2e = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = audio_s / float(t_e2e[0])
    rl = 1280 * T + 353
    e2e = {"value": e2e_value, "unit": "audio_s/s", "h2d_bytes_per_step": STREAMS * rl * 2 + STREAMS * 4,
           "d2h_bytes_per_step": STREAMS * (10 * T + 1) * 4, "ms_per_step": 1e3 * float(t_e2e[0]) / args.steps, "tokens": ntok,
           "tokens_per_audio_s": ntok / (STREAMS * CHUNK_S * args.steps)}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = cpu_baseline(None)
        lat_sorted = sorted(lat)
        line = {"metric": "rtfx", "value": value, "unit": "audio_s/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * t_dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": COMPUTE, "data": "synthetic",
                "config": {"workload": f"nemotron-speech-streaming-en-0.6b architecture ({N_LAYERS} conformer layers, random-init synthetic weights), "
                                       f"{COMPUTE} tcgen05 GEMMs, {KV} K/V ring, {STREAMS} concurrent streams per GPU, {int(CHUNK_S * 1000)} ms chunks (att_right_context={RIGHT_CONTEXT}), "
                                       f"steady state after {WARM_CHUNKS} warm chunks; joint blank bias calibrated to a speech-like token rate "
                                       f"({e2e['tokens_per_audio_s']:.1f} tokens per audio second measured in the e2e leg)" if PROFILE == "speech" else
                                       f"nemotron-speech-streaming-en-0.6b architecture ({N_LAYERS} layers, synthetic weights, parity-test calibration = dense emission), {STREAMS} streams, R={RIGHT_CONTEXT}",
                           "baseline_config": args.config, "streams_per_gpu": STREAMS, "chunk_ms": int(CHUNK_S * 1000),
                           "l2_policy": f"per-step working set (weights {0.62 if COMPUTE == 'q8_0' else 1.16} GB + K/V ring {STREAMS * 6.88e-3:.2f} GB) > 126 MB L2", "token_profile": PROFILE,
                           "device_resident_input": f"{BENCH_CHUNKS} consecutive chunks per stream staged in HBM, cycled"},
                "p50_chunk_latency_ms": lat_sorted[len(lat) // 2], "p99_chunk_latency_ms": lat_sorted[min(len(lat) - 1, int(0.99 * len(lat)))],
                "wall_ms_per_step": 1e3 * wall_s / args.steps,
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "breakdown": breakdown}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
