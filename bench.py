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


_REF_LIMITER = []


def _ref_child_init(cap_threads=True):
    """One thread per worker process, as BASELINE.md section 4 prescribes (OMP_NUM_THREADS=1): the pool already uses every
    host core, and NumPy's BLAS (mel_filters.T @ P, TF/audio_utils.py:813) would otherwise start one thread per core in
    EACH worker (round 1: 8-10x slower at N = 1 than under torchrun, which exports OMP_NUM_THREADS=1 itself)."""
    if not cap_threads:
        return
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMEXPR_NUM_THREADS"):
        os.environ[var] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _REF_LIMITER.append(threadpool_limits(limits=1))      # the already-loaded BLAS / OpenMP runtimes of this process
    except Exception:
        pass
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

    def __init__(self, recipe, procs, clips, cap_threads=True):
        import multiprocessing as mp
        self.procs = procs
        _ref_parent_init(recipe, clips)
        self.kind = _REF["kind"]
        self.n_clips = len(clips)
        self.pool = mp.get_context("fork").Pool(procs, initializer=_ref_child_init, initargs=(cap_threads,))
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
    procs = host_cores()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    clips = _cfg2_clips()
    pool = ReferencePool(args.recipe, procs, clips)
    for _ in range(args.warmup):
        pool.run(min(procs, CLIPS_PER_BATCH))
    times = [pool.run() for _ in range(args.steps)]
    pool.close()
    total = sum(times)
    value = args.steps * CLIPS_PER_BATCH * CLIP_SECONDS / total
    # secondary figure: the same pool with the BLAS / OpenMP thread counts left at their defaults (what a user who does not
    # set OMP_NUM_THREADS gets; oversubscribed, and what round 1 reported by mistake)
    uncapped = None
    if not args.no_uncapped:
        pool_u = ReferencePool(args.recipe, procs, clips, cap_threads=False)
        secs_u = pool_u.run()
        pool_u.close()
        uncapped = {"value": CLIPS_PER_BATCH * CLIP_SECONDS / secs_u, "unit": "audio-s/s",
                    "note": "same pool, BLAS/OpenMP threads uncapped (oversubscribed); one pass over the batch"}
    line = {
        "impl": "reference", "metric": "log-mel audio-seconds per second", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.recipe == "K" else "f32",
        "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": procs, "kind": pool.kind,
                         "sample": f"{args.steps} x {CLIPS_PER_BATCH} clips of {CLIP_SECONDS:.0f} s, one clip per call, "
                                   f"pool of {procs} single-threaded processes (OMP/BLAS threads = 1 per worker)",
                         "uncapped_threads": uncapped},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def host_cores() -> int:
    """Host cores this process may use (cgroup / affinity aware), the size of the reference's worker pool."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


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
class Ctx:
    """Process-group plumbing shared by the measurement legs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
        from speech_transcript_embeddings_b200 import _lib
        if not torch.cuda.is_available():
            raise _lib.StxError("bench.py needs a CUDA device: there is no CPU fallback for the B200 arm")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())

    def timed(self, fn, steps, warmup=3):
        """ms for `steps` calls of fn(i): CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        for i in range(warmup):
            fn(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(self.dev))
        for i in range(steps):
            fn(i)
        e1.record(torch.cuda.current_stream(self.dev))
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))


def make_pool(ctx, n, B, batches, seed_base):
    """`batches` batches of B class-G clips of n samples: pinned host copy (PackedClips) + device-resident copy."""
    torch = ctx.torch
    from speech_transcript_embeddings_b200.feature_extraction import PackedClips
    host, devb = [], []
    for pb in range(batches):
        lengths = np.full(B, n, np.int32)
        offsets = np.arange(B, dtype=np.int64) * n
        pinned = torch.empty(B * n, dtype=torch.float32, pin_memory=True)
        g = torch.Generator().manual_seed(100000 * ctx.rank + 1000 * pb + seed_base)
        torch.randn(B * n, generator=g, out=pinned)
        pinned.mul_(0.1)                                   # class G: 0.1 * N(0, 1)
        host.append(PackedClips(pinned, offsets, lengths))
        devb.append((pinned.to(ctx.dev, non_blocking=True), torch.from_numpy(offsets).to(ctx.dev),
                     torch.from_numpy(lengths).to(ctx.dev)))
    torch.cuda.synchronize(ctx.dev)
    return host, devb


def measure_recipe(ctx, args, recipe, host_batches, dev_batches, with_clocks):
    """Device-resident throughput, per-kernel roofline and the end-to-end legs of one recipe on the cfg2 batch."""
    torch = ctx.torch
    from speech_transcript_embeddings_b200 import _lib, ops
    from speech_transcript_embeddings_b200.feature_extraction import (B200SeamlessM4TFeatureExtractor,
                                                                      B200WhisperFeatureExtractor)
    dev, world = ctx.dev, ctx.world
    fe = B200SeamlessM4TFeatureExtractor(device=dev) if recipe == "K" else B200WhisperFeatureExtractor(device=dev)
    n, B = int(CLIP_SECONDS * 16000), CLIPS_PER_BATCH
    audio_s_per_step = B * CLIP_SECONDS
    T_pad = 2 * ((ops.k_num_frames(n) + 1) // 2)
    if recipe == "K":
        outs = [torch.empty((B, T_pad // 2, 160), dtype=torch.float32, device=dev) for _ in range(POOL_BATCHES)]
        masks = [torch.empty((B, T_pad // 2), dtype=torch.int32, device=dev) for _ in range(POOL_BATCHES)]

        def step_device(i):
            pcm_d, off_d, len_d = dev_batches[i % POOL_BATCHES]
            ops.fbank_k(pcm_d, off_d, len_d, n, T_pad, out=outs[i % POOL_BATCHES], mask=masks[i % POOL_BATCHES],
                        uniform=True)      # every clip of the cfg2 batch has n samples
        bytes_per_clip, dominant = K_BYTES_PER_CLIP, "k_frames"
    else:
        outs = [torch.empty((B, 80, n // 160), dtype=torch.float32, device=dev) for _ in range(POOL_BATCHES)]

        def step_device(i):
            pcm_d, off_d, len_d = dev_batches[i % POOL_BATCHES]
            ops.logmel_w(pcm_d, off_d, len_d, n, out=outs[i % POOL_BATCHES])
        bytes_per_clip, dominant = W_BYTES_PER_CLIP, "w_frames"

    # ---- device-resident throughput (value) -------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    ctx.barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(ctx.local_rank) if with_clocks else None
    long_run = None
    if clocks:
        clocks.__enter__()
    ev0.record(torch.cuda.current_stream(dev))
    for i in range(args.steps):
        step_device(i)
    ev1.record(torch.cuda.current_stream(dev))
    launches_timed = _lib.launch_count() - launches0
    ctx.barrier()
    if clocks:
        # the same work for two more seconds: nvidia-smi (0.2-1 s per call on some boxes) samples the clocks under this load,
        # and the long region is timed as a self-check of the short one (a 2 % regression shows at thousands of steps)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_end = time.perf_counter() + 2.0
        extra = 0
        l0.record(torch.cuda.current_stream(dev))
        while time.perf_counter() < t_end:
            step_device(extra)
            extra += 1
            if extra % 8 == 0:
                torch.cuda.synchronize(dev)
        l1.record(torch.cuda.current_stream(dev))
        torch.cuda.synchronize(dev)
        clocks.__exit__()
        long_run = {"steps": extra, "ms_per_step": l0.elapsed_time(l1) / max(extra, 1),
                    "note": "same step repeated for 2 s (host synchronises every 8 steps); this rank only"}
    ms_max = ctx.max_over_ranks(ev0.elapsed_time(ev1))
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
    dominant = max((k for k in per_kernel if k.startswith(dominant)), key=lambda k: sum(per_kernel[k]))
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

    # ---- end to end through the reference-signature call: a list of pageable NumPy arrays in, CPU tensors out --------
    # (R/processor.py:88-105 hands the extractor pageable np.ndarrays; the reference's own result is a CPU tensor)
    clips_np = [[hb.pcm[j * n:(j + 1) * n].numpy().copy() for j in range(B)] for hb in host_batches]
    last = {}

    def step_list(i):
        r = fe(clips_np[i % POOL_BATCHES], sampling_rate=16000, return_tensors="pt", output="host")
        last.clear()
        last.update(r)

    list_ms = ctx.timed(step_list, args.steps, max(args.warmup, 3))
    assert not last["input_features"].is_cuda
    h2d = B * n * 4 + 2 * B * 8
    d2h = sum(int(v.numel()) * v.element_size() for v in last.values())

    # ---- the same call on clips already packed in pinned host memory (no host packing) ----
    def step_pinned(i):
        r = fe(host_batches[i % POOL_BATCHES], sampling_rate=16000, return_tensors="pt", output="host")
        last.clear()
        last.update(r)

    pinned_ms = ctx.timed(step_pinned, args.steps, max(args.warmup, 3))
    e2e = {"value": world * args.steps * audio_s_per_step / (list_ms * 1e-3), "unit": "audio-s/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": list_ms / args.steps,
           "call": "extractor(list of 64 pageable float32 np.ndarrays, sampling_rate=16000, return_tensors='pt', output='host') "
                   "-> input_features and attention_mask as CPU tensors: the reference's own call signature and input type "
                   "(R/processor.py:101-105); host packing + chunked H2D | kernels | D2H pipeline inside the timed region",
           "pinned_input": {"value": world * args.steps * audio_s_per_step / (pinned_ms * 1e-3), "unit": "audio-s/s",
                            "ms_per_step": pinned_ms / args.steps,
                            "note": "same call on PackedClips already in pinned host memory (no host packing)"}}
    return {"value": value, "ms_per_step": ms_max / args.steps, "roofline": roofline, "e2e": e2e,
            "gpu_launches": launches_timed, "clocks": clocks.summary() if clocks else None, "long_run": long_run}


def measure_cfg1(ctx, with_reference):
    """cfg1: one 30 s clip, batch 1, through process_audio_array (R/inference.py:106 -> R/processor.py:79-126):
    wall-clock latency, host array in, device tensors out, synchronised."""
    torch = ctx.torch
    from speech_transcript_embeddings_b200 import synth
    from speech_transcript_embeddings_b200.processor import AudioTextProcessor
    clip = synth.clip("G", 480000, 0)
    res = {"workload": "cfg1: one 30 s clip, batch 1, AudioTextProcessor.process_audio_array, wall clock (median of 20)"}
    for key, name in (("K", "facebook/w2v-bert-2.0"), ("W", "openai/whisper-small")):
        proc = AudioTextProcessor(audio_model_name=name, device=ctx.dev)
        for _ in range(5):
            out = proc.process_audio_array(clip, 16000)
        torch.cuda.synchronize(ctx.dev)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            out = proc.process_audio_array(clip, 16000)
            torch.cuda.synchronize(ctx.dev)
            ts.append(time.perf_counter() - t0)
        res[key] = {"ms": 1e3 * statistics.median(ts), "audio_s_per_s": 30.0 / statistics.median(ts),
                    "shape": list(out["input_features"].shape)}
    if with_reference:
        try:
            import transformers
            torch.set_num_threads(1)
            for key, fe in (("K", transformers.SeamlessM4TFeatureExtractor()), ("W", transformers.WhisperFeatureExtractor())):
                ts = []
                for it in range(4):
                    t0 = time.perf_counter()
                    fe(clip, sampling_rate=16000, return_tensors="pt")
                    if it:
                        ts.append(time.perf_counter() - t0)
                res[key]["reference_cpu_ms_one_thread"] = 1e3 * statistics.median(ts)
        except ImportError:
            pass
    return res


def measure_cfg3(ctx):
    """cfg3: 512 clips of 1..30 s (seed 1234), one batched call, device-resident; masks checked exactly."""
    torch = ctx.torch
    from speech_transcript_embeddings_b200 import ops, synth
    from speech_transcript_embeddings_b200.feature_extraction import _layout
    dev = ctx.dev
    lens = synth.variable_lengths(512, 1234, True).astype(np.int32)
    offsets, total = _layout(lens)
    g = torch.Generator(device=dev).manual_seed(77 + ctx.rank)
    pcm = 0.1 * torch.randn(total, generator=g, device=dev)
    off_d, len_d = torch.from_numpy(offsets).to(dev), torch.from_numpy(lens).to(dev)
    audio_s = float(lens.sum()) / 16000.0
    frames = np.array([ops.k_num_frames(int(x)) for x in lens])
    T_pad = int(frames.max() + (frames.max() & 1))
    out_k = torch.empty((512, T_pad // 2, 160), dtype=torch.float32, device=dev)
    mask_k = torch.empty((512, T_pad // 2), dtype=torch.int32, device=dev)
    out_w = torch.empty((512, 80, 3000), dtype=torch.float32, device=dev)
    ms_k = ctx.timed(lambda i: ops.fbank_k(pcm, off_d, len_d, int(lens.max()), T_pad, out=out_k, mask=mask_k), 10) / 10
    ms_w = ctx.timed(lambda i: ops.logmel_w(pcm, off_d, len_d, 480000, out=out_w), 10) / 10
    want = (2 * np.arange(T_pad // 2)[None, :] + 1 < frames[:, None]).astype(np.int32)       # mask[j] = (2 j + 1 < T)
    masks_ok = bool(np.array_equal(mask_k.cpu().numpy(), want))
    # per-clip CMVN property on every clip: each normalised mel bin has mean 0 over the clip's own frames
    x = out_k.view(512, T_pad, 80)
    valid = (torch.arange(T_pad, device=dev)[None, :] < torch.from_numpy(frames).to(dev)[:, None]).unsqueeze(-1)
    mean = (x.double() * valid).sum(1) / torch.from_numpy(frames).to(dev)[:, None]
    cmvn_ok = bool(mean.abs().max().item() < 1e-4)
    return {"workload": "cfg3: 512 clips of 1-30 s (seed 1234), one batched call, device-resident", "audio_s": audio_s,
            "K": {"ms": ms_k, "audio_s_per_s": ctx.world * audio_s / ms_k * 1e3},
            "W": {"ms": ms_w, "audio_s_per_s": ctx.world * audio_s / ms_w * 1e3,
                  "note": "every clip is padded to 30 s by the recipe; real audio seconds counted"},
            "masks_exact": masks_ok, "cmvn_zero_mean_all_clips": cmvn_ok}


def measure_cfg4(ctx, seconds=1.0, pool=512, batch=128):
    """cfg4: 10 000 clip-hours sharded over the GPUs, front end only, no collective.  3.6e7 audio-seconds cannot be stored:
    every rank cycles a device-resident pool (>> L2, generated on the device) in calls of `batch` clips for >= `seconds`
    and the whole job is extrapolated from the measured rate (tools/bench_cfg4.py runs it in full)."""
    torch = ctx.torch
    from speech_transcript_embeddings_b200 import ops
    dev, n = ctx.dev, 480000
    g = torch.Generator(device=dev).manual_seed(4000 + ctx.rank)
    pcm = torch.empty(pool * n, dtype=torch.float32, device=dev)
    for i in range(0, pool, 64):
        pcm[i * n:(i + 64) * n] = 0.1 * torch.randn(64 * n, generator=g, device=dev)
    off = torch.arange(batch, dtype=torch.int64, device=dev) * n
    lens = torch.full((batch,), n, dtype=torch.int32, device=dev)
    T = ops.k_num_frames(n)
    T_pad = T + (T & 1)
    ring = [torch.empty((batch, T_pad // 2, 160), dtype=torch.float32, device=dev) for _ in range(2)]
    masks = [torch.empty((batch, T_pad // 2), dtype=torch.int32, device=dev) for _ in range(2)]
    calls = pool // batch

    def call(i):
        j = i % calls
        ops.fbank_k(pcm[j * batch * n:(j + 1) * batch * n], off, lens, n, T_pad, out=ring[i & 1], mask=masks[i & 1], uniform=True)

    probe = ctx.timed(call, calls, calls) / calls                       # ms per call, also the warm-up pass
    steps = max(calls, int(np.ceil(seconds * 1e3 / probe / calls)) * calls)
    ms = ctx.timed(call, steps, 0)
    rate = ctx.world * steps * batch * 30.0 / (ms * 1e-3)
    del pcm, ring
    return {"workload": f"cfg4: 10 000 clip-hours over {ctx.world} GPU(s), recipe K, pool of {pool} x 30 s per GPU "
                        f"({pool * n * 4 / 1e9:.2f} GB > L2), calls of {batch} clips",
            "timed_seconds": ms * 1e-3, "audio_s_processed": ctx.world * steps * batch * 30.0, "audio_s_per_s": rate,
            "extrapolated_seconds_for_10k_clip_hours": 3.6e7 / rate}


def f64_stripe_err(torch, S, a_rows, b_all, rows):
    """max |S[:rows] - normalize(a)[:rows] @ normalize(b).T| with the product in float64 on the device."""
    an = torch.nn.functional.normalize(a_rows[:rows].double(), dim=1)
    bn = torch.nn.functional.normalize(b_all.double(), dim=1)
    return float((S[:rows].double() - an @ bn.T).abs().max().item())


def measure_scoring(ctx, iters=20):
    """cfg5 and its multi-GPU form: every rank holds [N/W, D] audio and text embedding shards; the text shards are
    all-gathered (the only exchange on the whole path) and each rank computes its [N/W, M] stripe of the cosine matrix.
    Fused (P2P / multicast push + flag-acquiring tcgen05 GEMM over NVLink peer memory) against NCCL all-gather + GEMM."""
    torch, dist = ctx.torch, ctx.dist
    from speech_transcript_embeddings_b200 import scoring
    dev, world, rank = ctx.dev, ctx.world, ctx.rank
    res, worst = {}, 0.0
    # (name, a rows per rank, b rows per rank, D, rows of the stripe checked against float64)
    cases = [("cfg5_d768", None, None, 768, None), ("cfg5_d1024", None, None, 1024, None)]
    if world > 1:
        # retrieval-shaped: few queries per rank against a large sharded corpus, so that the exchange (NVLink) takes about as
        # long as the contraction and the overlap of the fused kernel has something to hide
        cases.append(("wide_512_queries_x_32768_corpus_rows_per_rank_d768", 512, 32768, 768, 64))
    for name, n_rank, m_rank, D, check_rows in cases:
        if n_rank is None:
            N = M = 4096
            lo, hi = scoring.shard_rows(N, world, rank)
            n_loc = m_loc = hi - lo
            counts = [scoring.shard_rows(M, world, r)[1] - scoring.shard_rows(M, world, r)[0] for r in range(world)]
        else:
            n_loc, m_loc = n_rank, m_rank
            N, M = n_rank * world, m_rank * world
            counts = [m_rank] * world
        g = torch.Generator(device=dev).manual_seed(1000 + rank)
        b_loc = torch.nn.functional.normalize(torch.randn(m_loc, D, generator=g, device=dev), dim=1)
        a_loc = torch.nn.functional.normalize(b_loc[:n_loc] + 0.5 * torch.randn(n_loc, D, generator=g, device=dev), dim=1)
        out = torch.empty((n_loc, M), dtype=torch.float32, device=dev)
        entry = {"N": N, "M": M, "D": D, "a_rows_per_rank": n_loc, "b_rows_per_rank": m_loc}
        if world > 1:
            scorer = scoring.GatheredScorer(max(counts), D, device=dev)
            fused = lambda i: scorer(a_loc, b_loc, counts=counts, out=out)                      # noqa: E731
            nccl = lambda i: scoring.sharded_cosine_matrix(a_loc, b_loc, counts=counts)         # noqa: E731
            S = fused(0)
            S_nccl = nccl(0)
            b_all = scoring.all_gather_rows(b_loc, counts=counts)
            rows = min(check_rows or n_loc, n_loc)
            err = f64_stripe_err(torch, S, a_loc, b_all, rows)
            diff = float((S - S_nccl).abs().max().item())
            del S_nccl, b_all
            entry.update({"ms_fused": ctx.timed(fused, iters, 5) / iters, "ms_nccl_allgather_then_gemm": ctx.timed(nccl, iters, 5) / iters,
                          "max_abs_err_vs_f64": ctx.max_over_ranks(err), "err_rows_checked_per_rank": rows,
                          "max_abs_diff_vs_nccl_path": ctx.max_over_ranks(diff),
                          "push": "multicast" if scorer.multicast_ptr is not None else "unicast",
                          "exchange_bytes_received_per_rank": (M - m_loc) * ((D + 63) // 64 * 64) * 4})
            entry["fused_speedup"] = entry["ms_nccl_allgather_then_gemm"] / entry["ms_fused"]
            worst = max(worst, entry["max_abs_err_vs_f64"], entry["max_abs_diff_vs_nccl_path"])
            del scorer
        else:
            local = lambda i: scoring.cosine_matrix(a_loc, b_loc)                                # noqa: E731
            S = local(0)
            err = f64_stripe_err(torch, S, a_loc, b_loc, n_loc)
            entry.update({"ms": ctx.timed(local, iters, 5) / iters, "max_abs_err_vs_f64": err})
            entry["tflops_algorithmic"] = 2.0 * N * M * D / entry["ms"] / 1e9
            worst = max(worst, err)
        entry["scores_per_s"] = float(N) * M / ((entry.get("ms_fused") or entry.get("ms")) * 1e-3)
        res[name] = entry
        del out, S
        torch.cuda.empty_cache()
    res["worst_abs_err"] = worst
    res["tolerance"] = 1e-5
    res["ok"] = bool(worst <= 1e-5)
    return res


def run_b200(args):
    ctx = Ctx(args)
    torch, dist = ctx.torch, ctx.dist
    recipe = args.recipe
    n, B = int(CLIP_SECONDS * 16000), CLIPS_PER_BATCH
    host_batches, dev_batches = make_pool(ctx, n, B, POOL_BATCHES, 7)
    main = measure_recipe(ctx, args, recipe, host_batches, dev_batches, with_clocks=True)

    secondary, scoring_block = {}, None
    if not args.no_secondary:
        other = "W" if recipe == "K" else "K"
        sec = measure_recipe(ctx, args, other, host_batches, dev_batches, with_clocks=False)
        secondary[f"recipe_{other}"] = {
            "value": sec["value"], "unit": "audio-s/s", "ms_per_step": sec["ms_per_step"],
            "dtype": "f32" if other == "W" else "f64",
            "roofline": {k: sec["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "kernel_ms", "kernels_ms")},
            "e2e": {"value": sec["e2e"]["value"], "pinned_input": sec["e2e"]["pinned_input"]["value"], "unit": "audio-s/s"},
            "note": ("Whisper log-mel (north_star's literal kernel list, TF/models/whisper/feature_extraction_whisper.py:135-164)"
                     if other == "W" else "SeamlessM4T/Kaldi fbank") + ", same cfg2 batch"}
        del host_batches, dev_batches
        torch.cuda.empty_cache()
        secondary["cfg3"] = measure_cfg3(ctx)
        secondary["cfg4"] = measure_cfg4(ctx)
        if ctx.rank == 0:
            secondary["cfg1"] = measure_cfg1(ctx, with_reference=ctx.world == 1)
        ctx.barrier()
        scoring_block = measure_scoring(ctx)

    # ---- CPU baseline: the reference's implementation on this box's host cores (rank 0, N = 1) ----
    cpu_baseline = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        procs = host_cores()
        pool = ReferencePool(recipe, procs, _cfg2_clips())
        pool.run(min(procs, CLIPS_PER_BATCH))
        secs = pool.run()
        pool.close()
        cpu_baseline = {"value": CLIPS_PER_BATCH * CLIP_SECONDS / secs, "unit": "audio-s/s", "cores": procs,
                        "kind": pool.kind,
                        "sample": f"one pass over the {CLIPS_PER_BATCH} x {CLIP_SECONDS:.0f} s cfg2 batch, one clip per call "
                                  f"(R/processor.py:101-105), pool of {procs} single-threaded processes; {secs:.2f} s"}

    line = None
    if ctx.rank == 0:
        line = {
            "metric": "log-mel audio-seconds per second", "value": main["value"], "unit": "audio-s/s", "n_gpus": ctx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if recipe == "K" else "f32", "data": "synthetic",
            "config": workload_config(args, ctx.world),
            "roofline": main["roofline"], "cpu_baseline": cpu_baseline, "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"], "clocks": main["clocks"], "long_run": main["long_run"],
            "secondary": secondary or None, "scoring": scoring_block,
        }
    if ctx.world > 1:
        dist.destroy_process_group()
    rc = 0
    if scoring_block is not None and not scoring_block["ok"]:
        rc = 1
    return line, rc


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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--recipe", choices=["K", "W"], default="K")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the secondary legs (the other recipe, cfg1, cfg3, cfg4) and the scoring block")
    ap.add_argument("--no-uncapped", action="store_true", help="reference arm: skip the secondary uncapped-threads pass")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rc = 0
    with _QuietStdout():
        if args.impl == "reference":
            line = run_reference(args)
        else:
            line, rc = run_b200(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return rc


if __name__ == "__main__":
    sys.exit(main())
