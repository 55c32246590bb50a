#!/usr/bin/env python
"""Headline benchmark: audio-seconds/sec (1/RTF) of the offline hot path, paraformer-large, batch 32 x 10 s per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libpfasr.so via the C-ABI)
    python bench.py --impl reference ...                      # the reference's CPU path (oracle port), host cores

One "step" = one pass of the hot path over one batch: PCM -> fbank/LFR/CMVN -> SAN-M encoder -> CIF -> decoder ->
log-softmax -> greedy ids (what OfflineRecognizer.GetResults does for 32 streams, OfflineRecognizer.cs:110-198).

* value  : whole-job audio-s/s with the PCM already resident in HBM (pf_offline_run_staged), device-timed with CUDA
           events on the engine's streams, max over ranks.  --lanes L (default 3) keeps L batches in flight per GPU,
           one host thread and one execution lane each (pf_offline_create_mt); the K steps are split over the lanes
           and timed from the first lane's start to the last lane's end.  --lanes 1: one batch at a time, L2 flushed
           between steps.
* e2e    : the same metric through the public call a user makes (pf_offline_run_pcm) with pinned HOST buffers, from
           the same L host threads: H2D of the PCM and D2H of the token ids are inside the timed region.
* roofline: dominant kernel = the tcgen05 GEMM; achieved = algorithmic GEMM FLOPs of one step / the CUDA-event time of
           that step's GEMM launches re-issued back to back on the engine's stream (average launch duration x launches);
           the per-launch event pairs of the profiled step are reported beside it (they serialise the launch chain);
           peak = MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a step).
* cpu_baseline: the oracle (a port of the reference's CPU path; OnnxRuntime/dotnet are not available) on the SAME
           32-utterance batch, all host cores, N = 1 only (1 warm-up + 2 timed passes).
* latency_single_lane: one batch at a time on a one-lane handle, L2 flushed between steps (the latency a lone caller sees).
* strong_scaling: global batch 32 through ONE multi-device handle over the N GPUs of the job (pf_offline_create with N
           devices, rank 0 drives it while the other ranks sit in a CPU-side barrier).
* shard_check (N > 1): rank 0 recomputes rank 1's lane-0 batch on its own GPU and compares with the ids gathered over NCCL.
Every timed loop runs for at least ~1 s: the K steps are repeated for `passes` passes and the figures are per step.
The SURVEY 8(d) metric (host PCM -> host ids) is `e2e`; `value` is the same loop with the PCM already resident in HBM.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio-seconds/sec (RTF^-1) paraformer-large offline @ batch 32"
UNIT = "audio-s/s"
BATCH = 32
SECONDS = 10.0
WORKLOAD = "paraformer-large-zh-en offline, batch=32x10 s synthetic 16 kHz per GPU (BASELINE configs[1])"


def workload_config(world: int):
    """Identical in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * world, "audio_seconds_per_utt": SECONDS, "T_lfr": 166,
            "weights": "random-init paraformer-large 50+16 (seed 20260917)",
            "l2": "every step streams 0.43 GB of weights plus activations, far more than the 126 MB L2; the single-lane latency leg "
                  "additionally writes a 256 MB buffer between steps",
            "parallelism": f"dp{world} (utterances sharded, weights replicated, no data-path collective)"}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1403.0))), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (every 25 ms)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.rows.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": reasons}


def _gemm_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu capture
    (profiles/gemm_traffic_r02b.json: mean over the four GEMMs of an encoder layer; older captures as fallbacks)."""
    for name in ("gemm_traffic_r02b.json", "gemm_traffic_r02.json", "gemm_traffic_r01_v11.json"):
        try:
            return int(json.load(open(os.path.join(ROOT, "profiles", name)))["avg_dram_bytes_per_launch"])
        except Exception:
            continue
    return None


def _oracle_dims(cfg):
    from oracle import sanm
    return sanm.ModelDims(**{k: v for k, v in cfg.as_dict().items() if k in sanm.ModelDims.__dataclass_fields__})


def cpu_reference_step(pcm, weights, cfg):
    """One pass of the reference's CPU path (oracle port): PCM in host RAM -> token ids in host RAM."""
    from oracle import frontend as F, sanm
    from aliparaformerasr_b200 import synth
    shift, scale = synth.make_cmvn()
    speech = F.pad_sequence([F.extract_features(p, shift, scale, snip_edges=cfg.snip_edges) for p in pcm])
    return sanm.paraformer_forward(speech, weights, _oracle_dims(cfg))["tokens"]


def time_cpu(pcm_sample, weights, cfg, steps, warmup, budget_s=None):
    """Times the CPU path with a pinned thread count (all host cores).  budget_s bounds the whole call: the number of
    timed steps is cut (never below 2) so that warm-up + steps fit."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    cpu_reference_step(pcm_sample, weights, cfg)               # first warm-up doubles as the calibration
    first = time.perf_counter() - t0
    if budget_s is not None:
        warmup = min(warmup, max(1, int(0.2 * budget_s / max(first, 1e-3))))
        steps = max(2, min(steps, int((budget_s - warmup * first) / max(first, 1e-3))))
    for _ in range(max(0, warmup - 1)):
        cpu_reference_step(pcm_sample, weights, cfg)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_reference_step(pcm_sample, weights, cfg)
        ts.append(time.perf_counter() - t0)
    return ts, cores, warmup


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=BATCH, help="utterances per CPU pass (default: the full batch, same config as the GPU arm)")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds the --impl reference run may take (steps are cut to fit)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="every timed loop repeats its K steps until it lasted this long")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=3,
                    help="batches in flight per GPU (pf_offline_create_mt): one host thread + one execution lane each")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from aliparaformerasr_b200 import synth
    cfg = synth.paraformer_large()

    # ------------------------------------------------------------------ reference arm: CPU, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        weights = synth.make_weights(cfg)
        nb = max(1, min(args.cpu_sample, BATCH))
        pcm = [synth.make_pcm(i, SECONDS) for i in range(nb)]
        ts, cores, warm = time_cpu(pcm, weights, cfg, max(1, args.steps), args.warmup, budget_s=args.ref_budget)
        sec = statistics.median(ts)
        val = nb * SECONDS / sec
        sample = (f"the full batch: {nb} utterances x 10 s per step, {len(ts)} timed steps after {warm} warm-up (cut from --steps {args.steps} to fit "
                  f"{args.ref_budget:.0f} s); torch threads pinned to {cores}; oracle port of the reference CPU path (OnnxRuntime/dotnet unavailable); "
                  f"CPU: {cpu_model_name()}")
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": len(ts),
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(max(1, args.gpus)),
            "ms_per_step_min_median_max": [min(ts) * 1e3, sec * 1e3, max(ts) * 1e3],
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return

    # ------------------------------------------------------------------ B200 arm
    # stdout carries exactly one JSON line: anything libraries print to fd 1 meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import torch.distributed as dist
    from aliparaformerasr_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")       # host-side waits that must not spin on a GPU

    weights = synth.make_weights(cfg)
    L = max(1, min(args.lanes, 8))
    eng = Engine(cfg, weights, devices=[local_rank], lanes=L)
    eng.set_cmvn(*synth.make_cmvn())
    nsamp = int(SECONDS * cfg.fs)
    # pinned host PCM: every lane works on its own batch (global utterance index = (rank * L + lane) * BATCH + i)
    host = torch.empty((L, BATCH, nsamp), dtype=torch.float32, pin_memory=True)
    for l in range(L):
        for i in range(BATCH):
            host[l, i].copy_(torch.from_numpy(synth.make_pcm((rank * L + l) * BATCH + i, SECONDS)))
    pcm = [[host[l, i].numpy() for i in range(BATCH)] for l in range(L)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # 2x the 126 MB L2
    steps_of = [args.steps // L + (1 if l < args.steps % L else 0) for l in range(L)]

    # one long-lived host thread per lane: libpfasr binds a thread to lane (thread ordinal mod lanes), the ordinal being
    # taken at the thread's first call into the library - the L workers make that call one after the other, so they own
    # L consecutive ordinals = L distinct lanes
    import concurrent.futures as cf
    pools = [cf.ThreadPoolExecutor(max_workers=1) for _ in range(L)]
    lane_ids = [pools[l].submit(lambda: (eng._lib.pf_offline_lane_acquire(eng._handle()), eng._lib.pf_offline_lane_release(eng._handle()))[0]).result()
                for l in range(L)]
    assert sorted(lane_ids) == list(range(L)), lane_ids

    def on_lanes(fn):
        futs = [pools[l].submit(fn, l) for l in range(L)]
        return [f.result() for f in futs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -------- value: inputs resident in HBM
    def lane_setup(l):
        torch.cuda.set_device(local_rank)
        eng.stage_pcm(pcm[l])
        o = None
        for _ in range(args.warmup):
            o = eng.run_staged()
        return o, eng.launch_count(), torch.cuda.ExternalStream(eng.stream_ptr(0), device=local_rank)

    setup = on_lanes(lane_setup)
    out = setup[0][0]
    launches_per_step = setup[0][1]
    streams = [x[2] for x in setup]

    def lane_resident(l, passes):
        st = streams[l]
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps_of[l] * passes):
            eng.run_staged()
        e1.record(st)
        return e0, e1

    def resident_pass(passes):
        evs = on_lanes(lambda l: lane_resident(l, passes))
        torch.cuda.synchronize()
        # device time from the earliest lane start to the latest lane end (events of different streams share the device clock)
        first = min(range(L), key=lambda l: evs[0][0].elapsed_time(evs[l][0]))
        return max(evs[first][0].elapsed_time(evs[l][1]) for l in range(L))

    def agree_passes(ms_one_pass):
        """Passes needed for the loop to last --min-seconds, the same on every rank."""
        p = max(1, int(np.ceil(args.min_seconds * 1e3 / max(ms_one_pass, 1e-3))))
        if world > 1:
            t = torch.tensor([p], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            p = int(t[0])
        return min(p, 200)

    passes = agree_passes(resident_pass(1))               # calibration pass (also more warm-up)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    total_ms = resident_pass(passes)
    barrier()
    wall_resident = time.perf_counter() - t_wall0
    stage_ms = pools[0].submit(eng.timings).result()

    # -------- e2e: host PCM -> host token ids through the public call, every lane fed by its own host thread
    tok_host = [np.zeros((max(1, steps_of[l]), BATCH, 256), np.int32) for l in range(L)]

    def lane_e2e(l, n):
        for k in range(n):
            o = eng.run_pcm(pcm[l])
            tok_host[l][k % tok_host[l].shape[0], :, : o.tokens.shape[1]] = o.tokens
        return None

    gathered = []

    def gather_ids():
        if world > 1:   # C1: the ids of the timed steps reach rank 0 over NCCL (NVLink), one padded gather
            t = torch.from_numpy(np.concatenate(tok_host, axis=0)).cuda(non_blocking=True)
            gl = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
            dist.gather(t, gl, dst=0)
            gathered[:] = gl or []

    on_lanes(lambda l: lane_e2e(l, 3))
    gather_ids()            # includes the first gather: NCCL builds its communicator lazily
    barrier()
    t0 = time.perf_counter()
    on_lanes(lambda l: lane_e2e(l, steps_of[l]))
    e2e_passes = agree_passes((time.perf_counter() - t0) * 1e3)
    barrier()
    t0 = time.perf_counter()
    on_lanes(lambda l: lane_e2e(l, steps_of[l] * e2e_passes))
    gather_ids()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = BATCH * nsamp * 4 + BATCH * 28
    d2h = int(out.tokens.size * 4 + BATCH * 4 + 16)
    # the same call fed with the file's own 16-bit samples (pf_offline_run_audio: conversion on the device, half the H2D)
    from aliparaformerasr_b200 import _lib, audio as pf_audio
    host16 = torch.empty((L, BATCH, nsamp), dtype=torch.int16).pin_memory()
    host16.copy_((host * 32768.0).round().clamp(-32768, 32767).to(torch.int16))
    clips = [[pf_audio.Audio(host16[l, i].numpy(), _lib.PF_AUDIO_S16, 1, 16000) for i in range(BATCH)] for l in range(L)]

    def lane_s16(l, n):
        for _ in range(n):
            eng.run_audio(clips[l])

    on_lanes(lambda l: lane_s16(l, 3))
    barrier()
    t0 = time.perf_counter()
    on_lanes(lambda l: lane_s16(l, steps_of[l] * e2e_passes))
    barrier()
    e2e16_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None       # sampled across the three timed loops (resident, e2e, e2e 16-bit)

    # -------- shard equivalence (N > 1): rank 0 recomputes rank 1's lane-0 batch on its own GPU; make_pcm is keyed by the
    # global utterance index, so the ids must equal what rank 1 computed and sent over NCCL
    shard_check = None
    if world > 1 and rank == 0:
        other = [synth.make_pcm((1 * L + 0) * BATCH + i, SECONDS) for i in range(BATCH)]
        mine = pools[0].submit(lambda: eng.run_pcm(other)).result()
        theirs = gathered[1][0].cpu().numpy()[:, : mine.tokens.shape[1]]
        shard_check = {"rank": 1, "lane": 0, "utterances": BATCH, "ids_identical": bool(np.array_equal(mine.tokens, theirs)),
                       "mismatching_ids": int((mine.tokens != theirs).sum())}

    # -------- roofline leg: one profiled step on lane 0 (per-launch CUDA events on the GEMM kernel)
    def lane_profile():
        eng.set_profile(True)
        eng.run_staged()
        g_ms, g_fl = eng.gemm_ms(), eng.gemm_flops()
        pr = [p for p in eng.profile() if p.get("name", "gemm") == "gemm"]
        # the same launches again, back to back and PDL-chained as inside the step, between two CUDA events
        rp_ms = eng.replay_gemms(5)
        # in-situ (warm, stream-ordered, no PDL overlap) time of every kernel family of the layers
        eng.set_profile(2)
        eng.run_staged()
        k_ms = {}
        for p in eng.profile():
            k_ms[p.get("name", "gemm")] = k_ms.get(p.get("name", "gemm"), 0.0) + p["ms"]
        eng.set_profile(0)
        return g_ms, g_fl, pr, k_ms, rp_ms

    # all lanes replay their step's GEMMs at the same time (the way the timed loops run them): aggregate tensor throughput
    conc = None
    if L > 1:
        eng.set_profile(True)
        fl = on_lanes(lambda l: (eng.run_staged(), eng.gemm_flops())[1])
        gate = threading.Barrier(L)
        ms = on_lanes(lambda l: (gate.wait(), eng.replay_gemms(5))[1])
        eng.set_profile(0)
        if min(ms) > 0:
            conc = {"tflops": sum(fl) / (max(ms) * 1e-3) / 1e12, "ms_per_pass_per_lane": ms,
                    "method": f"{L} lanes re-launch the GEMMs of their own step concurrently (5 passes each, started together); "
                              "sum of the lanes' algorithmic FLOPs / the slowest lane's time per pass"}
    gemm_ms_events, gemm_flops, prof, kernel_ms, gemm_ms = pools[0].submit(lane_profile).result()
    if gemm_ms <= 0:
        gemm_ms = gemm_ms_events
    n_gemm = sum(p["launches"] for p in prof) or 1
    for pl in pools:
        pl.shutdown()
    eng.close()

    # -------- latency of a lone caller: one-lane handle, one batch at a time, L2 flushed between steps
    def single_lane_leg(devices, tag):
        e1 = Engine(cfg, weights, devices=devices, lanes=1)
        e1.set_cmvn(*synth.make_cmvn())
        st = torch.cuda.ExternalStream(e1.stream_ptr(0), device=devices[0])
        for _ in range(args.warmup):
            e1.run_pcm(pcm[0])
        res, e2e = [], []
        e1.stage_pcm(pcm[0])
        for _ in range(args.steps):
            for d in devices:                              # flush every device's L2, outside the step's events
                with torch.cuda.device(d):
                    torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{d}").zero_()
                    torch.cuda.synchronize(d)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            e1.run_staged()
            b.record(st)
            b.synchronize()
            res.append(a.elapsed_time(b))
        for _ in range(args.steps):
            t0 = time.perf_counter()
            e1.run_pcm(pcm[0])
            e2e.append((time.perf_counter() - t0) * 1e3)
        e1.close()
        torch.cuda.set_device(local_rank)
        audio = BATCH * SECONDS
        return {"handle": tag, "global_batch": BATCH, "n_gpus": len(devices), "steps": args.steps,
                "resident_ms_min_median_max": [min(res), statistics.median(res), max(res)],
                "e2e_ms_min_median_max": [min(e2e), statistics.median(e2e), max(e2e)],
                "resident_value": audio / (statistics.median(res) * 1e-3), "e2e_value": audio / (statistics.median(e2e) * 1e-3), "unit": UNIT}

    latency = strong = None
    if rank == 0:
        latency = single_lane_leg([local_rank], "one lane on one GPU, L2 flushed between steps (device events around pf_offline_run_staged; "
                                                "host clock around pf_offline_run_pcm)")
    if world > 1:
        dist.barrier(group=cpu_group)                      # the other ranks wait on the CPU: their GPUs are idle from here on
        if rank == 0:
            strong = single_lane_leg(list(range(world)), f"ONE handle over {world} GPUs (pf_offline_create, devices 0..{world - 1}): the batch of 32 is "
                                                         "split contiguously, one host thread per device, Lmax exchanged on the host")
        dist.barrier(group=cpu_group)
    elif rank == 0:
        strong = dict(latency, handle="1 GPU: identical to latency_single_lane")

    # -------- reduce over ranks (max time).  (The multi-device handle of the strong-scaling leg leaves the calling thread on
    # its last device: go back to this rank's own.)
    torch.cuda.set_device(local_rank)
    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=torch.device("cuda", local_rank))
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        audio_per_step = BATCH * SECONDS * world
        nsteps, nsteps_e2e = args.steps * passes, args.steps * e2e_passes
        value = audio_per_step * nsteps / (total_ms / 1e3)
        e2e_val = audio_per_step * nsteps_e2e / (e2e_ms / 1e3)
        peak, peak_src = _peaks()
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        config = workload_config(world)
        single = {"achieved": achieved, "frac": achieved / peak if peak else None, "gemm_ms_per_step": gemm_ms, "avg_launch_us": gemm_ms * 1e3 / n_gemm,
                  "method": "all GEMM launches of one step re-launched back to back on ONE lane's stream (programmatic-launch chained as inside the "
                            "step, weights streamed from HBM: 0.43 GB per pass), 5 passes between two CUDA events; with several lanes the handle "
                            "prepares 256-wide 'throughput' tiles, which are not the fastest for a lone stream",
                  "achieved_per_launch_events": gemm_flops / (gemm_ms_events * 1e-3) / 1e12 if gemm_ms_events > 0 else None,
                  "per_launch_events_note": "an event pair around every launch breaks the launch chain and adds an event round trip per launch; by_shape uses these"}
        if conc:
            # the regime `value` is measured in: L lanes in flight.  achieved = the lanes' GEMM launches replayed concurrently
            head = {"achieved": conc["tflops"], "frac": conc["tflops"] / peak if peak else None, "method": conc["method"],
                    "gemm_ms_per_step": max(conc["ms_per_pass_per_lane"]) / L, "avg_launch_us": max(conc["ms_per_pass_per_lane"]) * 1e3 / n_gemm,
                    "avg_launch_us_note": f"per lane; {L} launches overlap at any time"}
        else:
            head = dict(single)
        roofline = {"bound": "tensor", "achieved": head["achieved"], "peak": peak, "unit": "TFLOP/s", "frac": head["frac"],
                    "traffic": _gemm_traffic(), "traffic_unit": "bytes per launch (ncu dram read+write, profiles/gemm_traffic_r02b.json)",
                    "kernel": "pf_gemm_f16_tn_tcgen05", "peak_source": peak_src, "method": head["method"],
                    "gemm_flops_per_step": gemm_flops, "gemm_launches_per_step": n_gemm, "gemm_ms_per_step": head["gemm_ms_per_step"],
                    "gemm_share_of_step": head["gemm_ms_per_step"] / (total_ms / nsteps) if total_ms else None,
                    "avg_launch_us": head["avg_launch_us"], "single_stream": single, "by_shape": prof}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / nsteps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": "synthetic", "config": config,
            "headline_note": "e2e is the SURVEY 8(d) metric (host PCM -> host token ids through pf_offline_run_pcm); value is the same loop with "
                             "the PCM already resident in HBM (pf_offline_run_staged)",
            "run": {"lanes": L, "passes": passes, "timed_steps": nsteps, "timed_seconds": total_ms / 1e3, "Lmax": int(out.tokens.shape[1]),
                    "T_lfr": int(out.feat_frames), "operands": "fp16 operands / fp32 accumulate",
                    "in_flight": f"{L} batches on {L} lanes (one host thread each)"},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / nsteps_e2e, "timed_steps": nsteps_e2e, "timed_seconds": e2e_ms / 1e3,
                    "api": f"pf_offline_run_pcm (C-ABI, pinned host PCM -> host token ids), {L} host threads on {L} lanes",
                    "s16_input": {"api": "pf_offline_run_audio (16-bit file samples, converted on the device; rank 0, no gather)",
                                  "ms_per_step": e2e16_s * 1e3 / nsteps_e2e, "value": BATCH * SECONDS * nsteps_e2e / e2e16_s,
                                  "h2d_bytes_per_step": BATCH * nsamp * 2 + BATCH * 52}},
            "latency_single_lane": latency,
            "strong_scaling": strong,
            "shard_check": shard_check,
            "gpu_launches": int(launches_per_step * nsteps * world),
            "launches_per_step": int(launches_per_step),
            "rtf": (total_ms / 1e3) / (audio_per_step * nsteps),
            "stage_ms": stage_ms,
            "kernel_ms_profiled_step": kernel_ms,
            "roofline": roofline,
            "clocks": clocks,
            "wall_s_resident_loop": wall_resident,
        }
        if not args.no_cpu_baseline and world == 1:
            nb = max(1, min(args.cpu_sample, BATCH))
            ts, cores, warm = time_cpu(pcm[0][:nb], weights, cfg, 2, 1, budget_s=60.0)
            sec = statistics.median(ts)
            line["cpu_baseline"] = {"value": nb * SECONDS / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                    "ms_per_pass_min_median_max": [min(ts) * 1e3, sec * 1e3, max(ts) * 1e3],
                                    "sample": f"the same {nb}-utterance batch, {warm} warm-up + {len(ts)} timed passes, torch threads pinned to {cores}; "
                                              f"oracle port of the reference CPU path (OnnxRuntime/dotnet unavailable); CPU: {cpu_model_name()}"}
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = {"skipped": "reported at N = 1 only (and by --impl reference): at N > 1 the other ranks would have to wait for it"}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
