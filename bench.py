#!/usr/bin/env python
"""Headline benchmark: log-mel audio-seconds per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--recipe K|W]

One "step" = one pass of the hot path over one batch of the cfg2 workload (64 synthetic 30 s clips per
GPU, recipe K by default = what the reference's processor runs).  Under torchrun every rank holds its
own clips (weak scaling, no data-path collective); the timed region is bracketed by barrier +
synchronize, timed with CUDA events on the launching stream, max over ranks.

JSON line keys (see the driver contract): value = device-resident throughput; e2e = the same metric
through the extractor's public call with HOST buffers (pinned), H2D + kernels + D2H inside the timed
region; roofline = algorithmic bytes / dominant kernel's average duration vs the measured HBM peak;
cpu_baseline = the reference's CPU implementation timed on this box's host cores (rank 0, N = 1).

--impl reference times the reference's own CPU implementation of the path (the transformers extractor
that R/processor.py:101-105 calls, one clip per call, a process pool over all host cores like the
reference's DataLoader workers, R/training/trainer_unfreeze.py:1429).
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
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

CLIP_SECONDS = 30.0
CLIPS_PER_BATCH = 64
POOL_BATCHES = 4                       # distinct batches rotated so that inputs + outputs >> 126 MB of L2
K_BYTES_PER_CLIP = 4 * 480000 + 1499 * 160 * 4 + 1499 * 4       # 2 885 356 B per 30 s clip
W_BYTES_PER_CLIP = 4 * 480000 + 80 * 3000 * 4                   # 2 880 000 B
K_F64_FLOP_PER_FRAME = 7600.0          # FP64 instructions (counted as flop) per frame, frame chain + FFT (lazy-scale codelets)
FALLBACK_HBM_GBS = 6650.0


# --------------------------------------------------------------------------------------------
# reference CPU implementation (also the cpu_baseline leg)
# --------------------------------------------------------------------------------------------
_REF = {}


def _ref_parent_init(recipe, clips):
    """Import the reference implementation and stage the clips BEFORE forking, so the workers inherit both."""
    try:
        import transformers
        _REF["fe"] = (transformers.SeamlessM4TFeatureExtractor() if recipe == "K"
                      else transformers.WhisperFeatureExtractor())
        _REF["kind"] = "reference"
    except Exception:                  # transformers absent: the NumPy port of the same arithmetic
        from oracle import fbank_k, logmel_w
        _REF["fe"] = None
        _REF["port"] = fbank_k if recipe == "K" else logmel_w
        _REF["kind"] = "port"
    _REF["clips"] = clips


def _ref_child_init():
    import torch
    torch.set_num_threads(1)


def _ref_one(idx):
    clip = _REF["clips"][idx]
    if _REF["fe"] is not None:
        out = _REF["fe"](clip, sampling_rate=16000, return_tensors="pt")     # R/processor.py:101-105
        return tuple(out["input_features"].shape)
    return _REF["port"].extract([clip])[0].shape


class ReferencePool:
    """The reference's CPU feature extraction, one clip per call, over a pool of worker processes
    (the reference's own concurrency model: DataLoader(num_workers=12), R/training/trainer_unfreeze.py:1429)."""

    def __init__(self, recipe, procs, clips):
        import multiprocessing as mp
        self.procs = procs
        _ref_parent_init(recipe, clips)
        self.kind = _REF["kind"]
        self.n_clips = len(clips)
        self.pool = mp.get_context("fork").Pool(procs, initializer=_ref_child_init)
        self.pool.map(_ref_one, [i % self.n_clips for i in range(procs * 2)], chunksize=1)   # warm every worker

    def run(self, count=None):
        """Wall seconds for one pass over the first `count` staged clips."""
        idx = list(range(self.n_clips if count is None else count))
        t0 = time.perf_counter()
        self.pool.map(_ref_one, idx, chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def _cfg2_clips():
    from speech_transcript_embeddings_b200 import synth
    return synth.batch_fixed(CLIPS_PER_BATCH, CLIP_SECONDS, "G", 0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    procs = os.cpu_count() or 1
    pool = ReferencePool(args.recipe, procs, _cfg2_clips())
    for _ in range(args.warmup):
        pool.run(min(procs, CLIPS_PER_BATCH))
    times = [pool.run() for _ in range(args.steps)]
    pool.close()
    total = sum(times)
    value = args.steps * CLIPS_PER_BATCH * CLIP_SECONDS / total
    line = {
        "impl": "reference", "metric": "log-mel audio-seconds per second", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": procs, "kind": pool.kind,
                         "sample": f"{args.steps} x {CLIPS_PER_BATCH} clips of {CLIP_SECONDS:.0f} s, one clip per call, "
                                   f"pool of {procs} processes"},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def workload_config(args, world):
    return {
        "workload": f"cfg2: batch of {CLIPS_PER_BATCH} synthetic {CLIP_SECONDS:.0f} s 16 kHz clips per GPU, "
                    f"recipe {args.recipe} ({'SeamlessM4T/Kaldi fbank, what the reference runs' if args.recipe == 'K' else 'Whisper log-mel'})",
        "clips_per_step_per_gpu": CLIPS_PER_BATCH, "clip_seconds": CLIP_SECONDS, "recipe": args.recipe,
        "parallelism": f"clips sharded over {world} GPU(s), no data-path collective",
        "l2": f"{POOL_BATCHES} distinct batches rotated: {POOL_BATCHES * CLIPS_PER_BATCH * 480000 * 4 / 1e6:.0f} MB of PCM "
              f"+ {POOL_BATCHES} output buffers, larger than the 126 MB L2",
    }


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.samples.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy benchmark)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def ncu_traffic(recipe):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(f"recipe_{recipe}")
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from speech_transcript_embeddings_b200 import _lib, ops, synth
    from speech_transcript_embeddings_b200.feature_extraction import (B200SeamlessM4TFeatureExtractor,
                                                                      B200WhisperFeatureExtractor, PackedClips)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise _lib.StxError("bench.py needs a CUDA device: there is no CPU fallback for the B200 arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    recipe = args.recipe
    fe = B200SeamlessM4TFeatureExtractor(device=dev) if recipe == "K" else B200WhisperFeatureExtractor(device=dev)
    n = int(CLIP_SECONDS * 16000)
    B = CLIPS_PER_BATCH
    audio_s_per_step = B * CLIP_SECONDS

    # ---- synthetic pool: POOL_BATCHES batches of B clips, packed in pinned host memory and resident on the device
    host_batches, dev_batches = [], []
    for pb in range(POOL_BATCHES):
        lengths = np.full(B, n, np.int32)
        offsets = np.arange(B, dtype=np.int64) * n
        pinned = torch.empty(B * n, dtype=torch.float32, pin_memory=True)
        g = torch.Generator().manual_seed(100000 * rank + 1000 * pb + 7)
        torch.randn(B * n, generator=g, out=pinned)
        pinned.mul_(0.1)                                   # class G: 0.1 * N(0, 1)
        host_batches.append(PackedClips(pinned, offsets, lengths))
        pcm_d = pinned.to(dev, non_blocking=True)
        off_d = torch.from_numpy(offsets).to(dev)
        len_d = torch.from_numpy(lengths).to(dev)
        dev_batches.append((pcm_d, off_d, len_d))
    T_pad = 2 * ((ops.k_num_frames(n) + 1) // 2)
    if recipe == "K":
        outs = [torch.empty((B, T_pad // 2, 160), dtype=torch.float32, device=dev) for _ in range(POOL_BATCHES)]
        masks = [torch.empty((B, T_pad // 2), dtype=torch.int32, device=dev) for _ in range(POOL_BATCHES)]

        def step_device(i):
            pcm_d, off_d, len_d = dev_batches[i % POOL_BATCHES]
            ops.fbank_k(pcm_d, off_d, len_d, n, T_pad, out=outs[i % POOL_BATCHES], mask=masks[i % POOL_BATCHES],
                        uniform=True)      # every clip of the cfg2 batch has n samples
        bytes_per_clip = K_BYTES_PER_CLIP
        dominant = "k_frames<false>"
    else:
        outs = [torch.empty((B, 80, n // 160), dtype=torch.float32, device=dev) for _ in range(POOL_BATCHES)]

        def step_device(i):
            pcm_d, off_d, len_d = dev_batches[i % POOL_BATCHES]
            ops.logmel_w(pcm_d, off_d, len_d, n, out=outs[i % POOL_BATCHES])
        bytes_per_clip = W_BYTES_PER_CLIP
        dominant = "w_frames"
    torch.cuda.synchronize(dev)

    # ---- device-resident throughput (value) -------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(torch.cuda.current_stream(dev))
        for i in range(args.steps):
            step_device(i)
        ev1.record(torch.cuda.current_stream(dev))
        launches_timed = _lib.launch_count() - launches0
        barrier()
        # same work for two more seconds so that nvidia-smi (100 ms cadence, but 0.2-1 s per call on some boxes) samples the clocks under this load
        t_end = time.perf_counter() + 2.0
        extra = 0
        while time.perf_counter() < t_end:
            step_device(extra)
            extra += 1
            if extra % 8 == 0:
                torch.cuda.synchronize(dev)
        torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * args.steps * audio_s_per_step / (ms_max * 1e-3)

    # ---- per-kernel durations (roofline leg): CUDA events around every launch, on the launching stream
    _lib.profile(True)
    for i in range(args.steps):
        step_device(i)
    torch.cuda.synchronize(dev)
    recs = _lib.profile_collect()
    _lib.profile(False)
    per_kernel = {}
    for name, kms in recs:
        per_kernel.setdefault(name, []).append(kms)
    if dominant not in per_kernel:                         # recipe K: k_frames_duo<false> (default) or k_frames<false> (STX_K_SINGLE=1)
        dominant = max((k for k in per_kernel if k.startswith(dominant.split("<")[0])), key=lambda k: sum(per_kernel[k]))
    dom_ms = statistics.mean(per_kernel[dominant])
    step_ms_prof = sum(sum(v) for v in per_kernel.values()) / args.steps
    peak, peak_src = measured_hbm_peak()
    achieved = B * bytes_per_clip / (dom_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": ncu_traffic(recipe), "kernel": dominant, "kernel_ms": dom_ms,
        "kernel_share_of_step": dom_ms / step_ms_prof if step_ms_prof else None,
        "algorithmic_bytes_per_launch": B * bytes_per_clip, "peak_source": peak_src,
        "kernels_ms": {k: statistics.mean(v) for k, v in per_kernel.items()},
    }
    if recipe == "K":
        frames = B * ops.k_num_frames(n)
        f64 = frames * K_F64_FLOP_PER_FRAME / (dom_ms * 1e-3) / 1e12
        roofline["fp64_pipe"] = {"achieved_tflops": f64, "peak_tflops": 148 * 64 * 1.965e9 / 1e12,
                                 "frac": f64 / (148 * 64 * 1.965e9 / 1e12),
                                 "note": "binding roof of recipe K: FP64 instruction issue (64 lanes/clk/SM), see DESIGN.md"}

    # ---- end to end through the public call: pinned host PCM -> H2D -> kernels -> D2H (pinned) ----
    last = {}

    def step_e2e(i):
        # output="host": pinned CPU tensors, like the reference's own CPU tensors; the call returns after the last
        # D2H copy (chunked H2D | kernels | D2H pipeline inside the extractor)
        r = fe(host_batches[i % POOL_BATCHES], sampling_rate=16000, return_tensors="pt", output="host")
        last.clear()
        last.update(r)

    for i in range(max(args.warmup, 3)):
        step_e2e(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(torch.cuda.current_stream(dev))
    for i in range(args.steps):
        step_e2e(i)
    e1.record(torch.cuda.current_stream(dev))
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * audio_s_per_step / (float(e2e_ms.item()) * 1e-3)
    assert not last["input_features"].is_cuda and last["input_features"].is_pinned()
    h2d = B * n * 4 + 2 * B * 8
    d2h = sum(int(v.numel()) * v.element_size() for v in last.values())

    # ---- the same call from a list of pageable NumPy arrays (includes packing into pinned memory) ----
    clips_np = [host_batches[0].pcm[j * n:(j + 1) * n].numpy().copy() for j in range(B)]
    for _ in range(2):
        fe(clips_np, sampling_rate=16000, return_tensors="np")
    t0 = time.perf_counter()
    reps = max(2, min(args.steps, 5))
    for _ in range(reps):
        fe(clips_np, sampling_rate=16000, return_tensors="np")
    list_s = (time.perf_counter() - t0) / reps

    # ---- CPU baseline: the reference's implementation on this box's host cores (rank 0, N = 1) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        procs = os.cpu_count() or 1
        pool = ReferencePool(recipe, procs, _cfg2_clips())
        pool.run(min(procs, CLIPS_PER_BATCH))
        secs = pool.run()
        pool.close()
        cpu_baseline = {"value": CLIPS_PER_BATCH * CLIP_SECONDS / secs, "unit": "audio-s/s", "cores": procs,
                        "kind": pool.kind,
                        "sample": f"one pass over the {CLIPS_PER_BATCH} x {CLIP_SECONDS:.0f} s cfg2 batch, one clip per call "
                                  f"(R/processor.py:101-105), pool of {procs} processes; {secs:.2f} s"}

    line = None
    if rank == 0:
        line = {
            "metric": "log-mel audio-seconds per second", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if recipe == "K" else "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": float(e2e_ms.item()) / args.steps, "wall_ms_per_step": wall_ms / args.steps,
                    "call": "extractor(PackedClips in pinned host memory, sampling_rate=16000, return_tensors='pt', "
                            "output='host') -> input_features and attention_mask as pinned CPU tensors "
                            "(chunked H2D | kernels | D2H pipeline on three streams)",
                    "from_numpy_list": {"value": audio_s_per_step / list_s, "unit": "audio-s/s",
                                        "note": "extractor(list of 64 pageable NumPy arrays, return_tensors='np'): "
                                                "adds host packing into pinned memory; 1 rank, wall clock"}},
            "gpu_launches": launches_timed, "clocks": clocks.summary(),
        }
    if world > 1:
        dist.destroy_process_group()
    return line


class _QuietStdout:
    """Everything except the final JSON line goes to stderr, whatever libraries print (NCCL writes its
    version banner to stdout at communicator creation)."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--recipe", choices=["K", "W"], default="K")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    with _QuietStdout():
        line = run_reference(args) if args.impl == "reference" else run_b200(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
