#!/usr/bin/env python
"""bench.py -- headline benchmark of the dense feature-detection hot path on B200.

Metric (BASELINE.json): Mpixel/s (and frames/s) of FAST + greedy NMS selection + BRIEF-256 on a batch of
1024 synthetic 752x480 grayscale frames per GPU (configs[1]; option values are the reference demo's:
threshold 10, min distance 20, 200 features, BRIEF length 256 / half patch 8).  A "step" is one pass of
that path over the whole batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun, one rank per GPU; frames are independent, so every rank runs its own batch
with no data-path collective (weak scaling) and the timing is the max over ranks.  One JSON line on rank 0:
the headline (value / roofline / cpu_baseline / e2e), `parity_checks` (the GPU totals against the reference's
own run on the same frames; at N > 1 also the row-tiled 4K frame over all GPUs against the untiled run and the
gathered batch against the local results -- a mismatch exits non-zero), and `extras` with one object per other
BASELINE.json config (2: Shi-Tomasi 1280x720, 3: Harris 3840x2160 incl. the tiled form, 4: LSD field 1920x1080),
each with its own roofline / cpu_baseline / e2e in SURVEY.md 8(d)'s accounting.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 752, 480
THR, DIST, NEEDED = 10.0, 20, 200
BRIEF_LEN, BRIEF_HALF = 256, 8
CAND_CAPACITY = 65536  # per-frame candidate slots (kN=9 at the demo threshold yields up to a few 10^4 on busy frames)


# ---------------------------------------------------------------------------------------------------
def make_frames(n: int, start: int, w: int = W, h: int = H) -> np.ndarray:
    """frames start .. start+n-1 of the SURVEY.md 8d generator at w x h, cached under /tmp, built in worker processes
    (must run before CUDA is initialised in this process)."""
    from feature_detector_b200.synth import SEED, synth
    cache = f"/tmp/fd_b200_synth_{w}x{h}_{SEED}_{start}_{n}.npy"
    if os.path.exists(cache):
        try:
            a = np.load(cache)
            if a.shape == (n, h, w):
                return a
        except Exception:
            pass
    import multiprocessing as mp
    workers = max(1, min(32, (os.cpu_count() or 1)))
    with mp.get_context("fork").Pool(workers) as pool:
        frames = pool.starmap(synth, [(w, h, start + i) for i in range(n)], chunksize=max(1, min(8, n // workers or 1)))
    a = np.stack(frames)
    try:
        np.save(cache + ".tmp.npy", a)
        os.replace(cache + ".tmp.npy", cache)
    except Exception:
        pass
    return a


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, reasons, mx = [], set(), 0
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_traffic(key: str):
    """dram__bytes_read + dram__bytes_write per launch of the named kernel, from the committed ncu summary (profiles/kernel_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def bind_near_gpu(local_rank: int):
    """Multi-GPU runs: keep this rank's threads -- and so the pinned host buffers it is about to allocate (first touch) -- on the
    CPUs the driver reports as local to its GPU.  Best effort: any failure leaves the default placement."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].strip().isdigit() else local_rank
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpus = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpus + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def _checker():
    from oracle.bindings import Port, Ref, have_ref
    return (Ref(), "reference") if have_ref() else (Port(), "port")


def cpu_points(kind_name: str, frames: np.ndarray, thr: float, dist: int, needed: int, fast_n: int, brief_len: int, threads: int):
    """The reference's own CPU path (oracle/_ref when present, else the C port) on `frames`: DetectGoodFeatures (+ Compute when
    brief_len > 0), one detector object per thread.  Returns (seconds, (keypoints, candidates) totals, kind)."""
    from oracle import bindings as ob
    chk, kind = _checker()
    k = {"fast": ob.FAST, "harris": ob.HARRIS, "shi": ob.SHI_TOMAS}[kind_name]
    sec, totals = chk.bench_points(k, frames, thr, dist, needed, fast_n=fast_n, brief_length=brief_len, brief_half_patch=BRIEF_HALF, n_threads=threads)
    return sec, (int(totals[0]), int(totals[1])), kind


def cpu_reference_mpx(frames: np.ndarray, fast_n: int, threads: int):
    sec, totals, kind = cpu_points("fast", frames, THR, DIST, NEEDED, fast_n, BRIEF_LEN, threads)
    return frames.shape[0] * H * W / sec / 1e6, kind, sec, totals


# ---------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation on the host cores, same workload definition; each step is a
    bounded sample of the batch so the run ends within minutes."""
    if rank != 0:
        return
    threads = host_threads()
    # A step is a bounded sample of the batch, sized from a warm-up measurement so that the K timed steps take about a minute.
    probe = make_frames(min(args.frames, max(threads * 4, 64)), 0)
    t0 = time.perf_counter()
    for _ in range(max(args.warmup, 1)):
        cpu_reference_mpx(probe, args.fast_n, threads)
    fps = max(args.warmup, 1) * len(probe) / max(time.perf_counter() - t0, 1e-6)
    sample = int(min(args.frames, max(threads, 60.0 * fps / max(args.steps, 1))))
    frames = make_frames(args.frames, 0)[:sample]
    t0 = time.perf_counter()
    kind = "port"
    for _ in range(args.steps):
        _, kind, _, _ = cpu_reference_mpx(frames, args.fast_n, threads)
    dt = time.perf_counter() - t0
    mpx = args.steps * sample * H * W / dt / 1e6
    line = {
        "impl": "reference", "metric": "Mpixel/s (FAST+NMS+BRIEF, 752x480)", "value": round(mpx, 3), "unit": "Mpixel/s",
        "frames_per_s": round(args.steps * sample / dt, 2), "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(args, sample),
        "cpu_baseline": {"value": round(mpx, 3), "unit": "Mpixel/s", "cores": threads, "kind": kind,
                         "sample": f"{sample} frames per step, all {threads} host threads, one detector object per thread"},
        "e2e": {"value": round(mpx, 3), "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, frames_per_step):
    return {"workload": f"FAST(kN={args.fast_n}, diff 15; pre-check {'on' if args.fast_n >= 12 else 'off'}) + greedy min-distance selection + "
                        f"BRIEF-256 on {frames_per_step} synthetic {W}x{H} u8 frames per GPU (BASELINE.json configs[1])",
            "frames_per_gpu": frames_per_step, "rows": H, "cols": W, "min_valid_response": THR, "min_feature_distance": DIST,
            "needed_feature_num": NEEDED, "brief_length": BRIEF_LEN, "brief_half_patch": BRIEF_HALF, "fast_n": args.fast_n,
            "generator": "feature_detector_b200.synth.synth (SURVEY.md 8d), seed 20261018",
            "l2_policy": "input batch (370 MB) exceeds the 126 MB L2, no flush needed between steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500, help="timed steps (0.6 ms each at N=1: the default keeps the timed region long enough to sample clocks)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--fast-n", dest="fast_n", type=int, default=9, help="9 = full segment test on every pixel (configs[1] 'FAST-9'); 12 = reference default")
    ap.add_argument("--chunk", type=int, default=128, help="frames per chunk of the host pipeline (e2e leg)")
    ap.add_argument("--no-extras", action="store_true", help="skip the other configs and the secondary measurements")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # ---- inputs first (worker processes), then CUDA --------------------------------------------------
    near = bind_near_gpu(local_rank) if world > 1 else None
    frames = make_frames(args.frames, rank * args.frames)
    threads = host_threads()
    cpu = None
    cpu_check = None   # (frames checked, (keypoints, candidates)) of the reference on a prefix of this rank's batch
    if rank == 0 and world == 1:
        sample = args.frames                      # the whole batch, three times: ~3 s wall, ~50 core-seconds on 16 threads
        cpu_reference_mpx(frames[:max(threads * 2, 16)], args.fast_n, threads)   # warm the thread pool / page the frames in
        best, kind, totals, secs = 0.0, "port", (0, 0), []
        for _ in range(3):
            mpx, kind, sec, totals = cpu_reference_mpx(frames[:sample], args.fast_n, threads)
            best = max(best, mpx)
            secs.append(sec)
        cpu = {"value": round(best, 3), "unit": "Mpixel/s", "cores": threads, "kind": kind,
               "sample": f"all {sample} frames of the batch, best of 3 passes ({', '.join(f'{x:.2f}' for x in secs)} s), {threads} host threads "
                         f"(one detector object per thread: the reference itself is single-threaded); {int(totals[0])} keypoints, "
                         f"{int(totals[1])} candidates"}
        cpu_check = (sample, totals)
        n1 = min(sample, 128)
        mpx1, _, sec1, _ = cpu_reference_mpx(frames[:n1], args.fast_n, 1)
        cpu["single_thread_mpixel_s"] = round(mpx1, 3)
        cpu["single_thread_sample"] = f"first {n1} frames, {sec1:.2f} s"
    elif rank == 0:
        n_chk = min(args.frames, 128)             # N > 1: a bounded prefix is enough for the totals check
        _, _, _, totals = cpu_reference_mpx(frames[:n_chk], args.fast_n, threads)
        cpu_check = (n_chk, totals)

    # the other configs' frames: distinct within a batch; at N > 1 every rank runs the same set (rank 0 builds the cache first)
    extra_frames = {}
    if not args.no_extras:
        shapes = {"c2": (1280, 720, 512), "c3": (3840, 2160, 64), "c4": (1920, 1080, 256)}
        if rank == 0:
            for key, (w2, h2, n2) in shapes.items():
                extra_frames[key] = make_frames(n2, 0, w2, h2)

    import torch
    import torch.distributed as dist
    import feature_detector_b200 as fd

    torch.cuda.set_device(local_rank)
    cpu_group = None
    os.environ.setdefault("NCCL_DEBUG", "WARN")   # NCCL's default prints its version on stdout, next to the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")     # host-side barriers while one rank drives all GPUs (tiled path)
    dev = torch.device("cuda", local_rank)
    if not args.no_extras and rank != 0:
        dist.barrier(group=cpu_group)                  # rank 0 has written the cache by the time it gets here
        for key, (w2, h2, n2) in shapes.items():
            extra_frames[key] = make_frames(n2, 0, w2, h2)
    elif not args.no_extras and world > 1:
        dist.barrier(group=cpu_group)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        if world > 1:
            dist.barrier(group=cpu_group)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(flag: bool) -> bool:
        if world == 1:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    host = torch.from_numpy(frames).pin_memory()
    d_frames = host.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    n, px = args.frames, H * W

    ctx = fd.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)  # the kernels AND the timing events go on this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    det = fd.DetectParams(fd.FAST, THR, DIST, NEEDED, fast_n=args.fast_n)
    brief = fd.BriefParams(BRIEF_LEN, BRIEF_HALF)
    peak, peak_src = measured_peak_gbs()

    def device_step():
        ctx.detect(det, CAND_CAPACITY)
        ctx.describe_selected(brief)

    timed_launches = [0]

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        timed_launches[0] = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        timed_launches[0] = ctx.launch_count - timed_launches[0]   # kernels launched inside the timed region
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / 1e3)

    def wall_timed(fn, steps, warmup):
        """Host-clock timing of host-to-host paths (copies included), barrier + synchronize on both sides, max over ranks."""
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        return max_over_ranks(time.perf_counter() - t0)

    # ---- value: device-resident input -> device-resident keypoints + descriptors ----------------------
    ctx.bind_device(d_frames.data_ptr(), H, W, n)
    sampler = ClockSampler(local_rank)
    sampler.start()
    sec = timed(device_step, args.steps, args.warmup)
    launches = timed_launches[0]
    clocks = sampler.stop()
    ctx.sync()  # raises if a candidate slot overflowed
    kp_counts = ctx.keypoint_counts()
    cand_counts = ctx.candidate_counts()
    value = world * n * px * args.steps / sec / 1e6

    # ---- roofline of the dominant kernel (FAST candidates), timed alone ------------------------------
    def fast_only():
        ctx.compute_candidates(det, CAND_CAPACITY)

    sec_k = timed(fast_only, args.steps, args.warmup)
    per_launch = sec_k / args.steps
    # per frame, the u8 frame in + (8 + 32) bytes per keypoint out; x the frames one launch processes
    algo_bytes = n * px + int(kp_counts.sum()) * 40
    achieved = algo_bytes / per_launch / 1e9
    roofline = {"bound": "hbm", "kernel": "fdb::fast_sparse_kernel (fd_fast_sparse.cu; thr 10 leaves s_min >= 7, so the sparse form runs)", "achieved": round(achieved, 2),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": kernel_traffic(f"fast_sparse_kernel_kN{args.fast_n}"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "us_per_launch": round(per_launch * 1e6, 2),
                "step_frac": round(algo_bytes / (sec / args.steps) / 1e9 / peak, 4),
                "note": "SURVEY.md 8(d) numerator: frame bytes + 40 B per keypoint (the intermediate candidate keys are not counted). frac = the FAST kernel "
                        "alone; step_frac = the same bytes over the whole step (FAST + selection + BRIEF). 1 B/px kernels are instruction-issue bound, see "
                        "DESIGN.md section 4 and profiles/; traffic = dram bytes per launch from the ncu capture summarised in profiles/kernel_traffic.json"}

    # ---- e2e: pinned host frames in, keypoints + descriptors back on the host, every step -------------
    # The public host-batch call (feature_detector_b200.pipeline.HostPipeline.run): chunks of frames alternate between
    # two contexts so that the PCIe copies of one chunk overlap the kernels of the other.
    from feature_detector_b200.pipeline import HostPipeline
    pinned_kp = torch.zeros((n, NEEDED, 16), dtype=torch.uint8).pin_memory()
    pinned_cnt = torch.zeros((n,), dtype=torch.int32).pin_memory()
    pinned_desc = torch.zeros((n, NEEDED, 32), dtype=torch.uint8).pin_memory()
    kp_host = pinned_kp.numpy().view(fd.KEYPOINT_DTYPE).reshape(n, NEEDED)
    cnt_host = pinned_cnt.numpy()
    desc_host = pinned_desc.numpy()
    h2d = n * px
    d2h = n * NEEDED * 16 + n * 4 + n * NEEDED * 32
    pipe = HostPipeline(local_rank, chunk_frames=args.chunk)

    def e2e_step():
        pipe.run(host.data_ptr(), H, W, n, det, brief, kp_host, cnt_host, desc_host, CAND_CAPACITY)

    e2e_steps = max(3, min(args.steps, 10))
    e2e_sec = wall_timed(e2e_step, e2e_steps, 3)
    e2e_value = world * n * px * e2e_steps / e2e_sec / 1e6
    e2e_match = bool(np.array_equal(cnt_host, kp_counts))  # the pipelined result equals the device-resident run
    pipe.close()

    # the ceiling of that leg: the same pinned bytes, plain cudaMemcpyAsync host to device on every rank at once, no kernels
    d_sink = torch.empty_like(d_frames)

    def h2d_only():
        d_sink.copy_(host, non_blocking=True)

    h2d_sec = wall_timed(h2d_only, e2e_steps, 2)
    h2d_ceiling = {"gbytes_s_all_ranks": round(world * h2d * e2e_steps / h2d_sec / 1e9, 2), "gbytes_s_per_rank": round(h2d * e2e_steps / h2d_sec / 1e9, 2),
                   "e2e_fraction_of_ceiling": round((e2e_steps / e2e_sec) / (e2e_steps / h2d_sec), 3),
                   "what": f"{world} rank(s) concurrently copying their {h2d / 1e6:.0f} MB pinned batch host to device, no kernels, same barrier + wall clock as e2e"}
    del d_sink

    # ---- parity: the GPU totals against the reference's own run on the same frames ---------------------
    parity = {}
    ok = True
    if cpu_check is not None:
        n_chk, (cpu_kp, cpu_cand) = cpu_check
        parity["cpu_totals_match"] = bool(int(kp_counts[:n_chk].sum()) == cpu_kp and int(cand_counts[:n_chk].sum()) == cpu_cand)
        parity["cpu_totals"] = {"frames": n_chk, "keypoints": {"reference": cpu_kp, "gpu": int(kp_counts[:n_chk].sum())},
                                "candidates": {"reference": cpu_cand, "gpu": int(cand_counts[:n_chk].sum())}}
        ok &= parity["cpu_totals_match"]
    parity["e2e_equals_device_resident_run"] = all_true(e2e_match)
    ok &= parity["e2e_equals_device_resident_run"]

    if world > 1:
        # (a) the batch sharded over the ranks and gathered on rank 0 == every rank's local results
        from feature_detector_b200 import sharding
        nf_g = 16                                      # per rank; the gather moves fixed-capacity slots, a few frames show it all
        got = sharding.detect_sharded(ctx, d_frames[:nf_g], world * nf_g, rank, world, det, brief, gather=True, dst=0, cand_capacity=CAND_CAPACITY)
        ctx.bind_device(d_frames.data_ptr(), H, W, nf_g)
        ctx.detect(det, CAND_CAPACITY)
        ctx.describe_selected(brief)
        kp_l, cnt_l = ctx.keypoints(NEEDED)
        desc_l = ctx.descriptors(NEEDED)
        mine = torch.zeros((nf_g, NEEDED, 13), dtype=torch.float32, device=dev)   # x, y, response + 8 descriptor words, count in [.., 0, 12]
        packed = np.zeros((nf_g, NEEDED, 13), np.float32)
        packed[..., 0], packed[..., 1], packed[..., 2] = kp_l["x"], kp_l["y"], kp_l["response"]
        packed[..., 3:11] = desc_l.reshape(nf_g, NEEDED, 8, 4).view(np.float32)[..., 0] if False else desc_l.view(np.float32).reshape(nf_g, NEEDED, 8)
        for f in range(nf_g):
            packed[f, cnt_l[f]:, :] = 0
        packed[:, 0, 12] = cnt_l
        mine.copy_(torch.from_numpy(packed))
        everyone = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(everyone, mine)
        same = True
        if rank == 0:
            for r in range(world):
                e = everyone[r].cpu().numpy()
                blk = slice(r * nf_g, (r + 1) * nf_g)
                cnt_r = e[:, 0, 12].astype(np.int32)
                same &= bool(np.array_equal(got["counts"][blk], cnt_r))
                for f in range(nf_g):
                    k = cnt_r[f]
                    g_kp = got["keypoints"][blk][f, :k]
                    same &= bool(np.array_equal(g_kp["x"], e[f, :k, 0]) and np.array_equal(g_kp["y"], e[f, :k, 1]) and
                                 np.array_equal(g_kp["response"].view(np.uint32), e[f, :k, 2].view(np.uint32)))
                    same &= bool(np.array_equal(got["descriptors"][blk][f, :k].view(np.uint32).reshape(k, 8), e[f, :k, 3:11].view(np.uint32)))
        parity["gather_equals_local"] = all_true(same)
        ok &= parity["gather_equals_local"]
        ctx.bind_device(d_frames.data_ptr(), H, W, n)

        # (b) one 3840x2160 frame row-tiled over every GPU of the box (fd_tiled_*, one process: rank 0) == the untiled run
        tiled_same = True
        torch.cuda.synchronize()
        host_barrier()
        if rank == 0:
            from feature_detector_b200.synth import synth
            big = synth(3840, 2160, 0)[None]
            har = fd.DetectParams(fd.HARRIS, 30.0, 20, 200)
            # FAST: the running offset reaches 80 on a frame this size, so threshold 90 keeps the candidates to the last rows'
            # corners -- and only absolute pixel indices make the tiles agree with the whole frame
            fst = fd.DetectParams(fd.FAST, 90.0, 20, 200, fast_n=9)
            with fd.TiledDetector(list(range(world))) as td, fd.Context(0) as c1:
                td.upload(big)
                c1.upload(big)
                for prm in (har, fst):
                    td.detect(prm, 0)
                    c1.detect(prm, 0)
                    kp_t, cnt_t = td.keypoints(200)
                    kp_u, cnt_u = c1.keypoints(200)
                    tiled_same &= bool(cnt_t[0] == cnt_u[0] and np.array_equal(kp_t[0, :cnt_t[0]], kp_u[0, :cnt_u[0]]))
                    tiled_same &= bool(np.array_equal(td.candidates(0), c1.candidates(0)))
            torch.cuda.set_device(local_rank)
        host_barrier()
        parity["tiled_equals_untiled"] = all_true(tiled_same)
        ok &= parity["tiled_equals_untiled"]

    extras = {}
    if not args.no_extras:
        ctx.bind_device(d_frames.data_ptr(), H, W, n)
        steps2 = max(3, min(args.steps // 2, 100))

        def measure(name, params, cap, with_brief):
            def step():
                ctx.detect(params, cap)
                if with_brief:
                    ctx.describe_selected(brief)
            s = timed(step, steps2, 2)
            ctx.sync()
            extras[name] = {"mpixel_s": round(world * n * px * steps2 / s / 1e6, 1), "frames_s": round(world * n * steps2 / s, 1),
                            "hbm_frac_1Bpx": round(n * px * steps2 / s / 1e9 / peak, 4), "mean_keypoints": float(ctx.keypoint_counts().mean()),
                            "mean_candidates": float(ctx.candidate_counts().mean())}

        other_n = 12 if args.fast_n < 12 else 9
        measure(f"fast_kN{other_n}+select+brief", fd.DetectParams(fd.FAST, THR, DIST, NEEDED, fast_n=other_n), CAND_CAPACITY, True)
        measure("harris+select (thr 30, d 20, N 200)", fd.DetectParams(fd.HARRIS, 30.0, 20, 200), 65536, False)
        measure("shi_tomas+select (thr 40, d 20, N 200)", fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 200), 65536, False)

        # per-stage times of the headline step (each stage timed alone on the same stream)
        t_c = timed(lambda: ctx.compute_candidates(det, CAND_CAPACITY), steps2, 2)
        t_d = timed(lambda: ctx.detect(det, CAND_CAPACITY), steps2, 2)
        ctx.detect(det, CAND_CAPACITY)
        t_b = timed(lambda: ctx.describe_selected(brief), steps2, 2)
        extras["headline step by stage (us per 1024 frames)"] = {"fast_candidates": round(t_c / steps2 * 1e6, 1), "selection": round((t_d - t_c) / steps2 * 1e6, 1),
                                                               "brief": round(t_b / steps2 * 1e6, 1)}

        # ---- one frame at a time, host to host: what a caller of the drop-in classes sees per DetectGoodFeatures + Compute ----
        one = frames[0]
        lat = []
        for i in range(220):
            t0 = time.perf_counter()
            ctx.upload(one)
            ctx.detect(det, CAND_CAPACITY)
            ctx.describe_selected(brief)
            ctx.keypoints(NEEDED)
            ctx.descriptors(NEEDED)
            if i >= 20:
                lat.append(time.perf_counter() - t0)
        extras["single frame, host to host (upload, FAST, selection, BRIEF, download; pageable buffers)"] = {
            "latency_us_median": round(float(np.median(lat)) * 1e6, 1), "latency_us_p90": round(float(np.percentile(lat, 90)) * 1e6, 1),
            "mpixel_s": round(px / float(np.median(lat)) / 1e6, 1)}
        lat = []
        for i in range(220):
            t0 = time.perf_counter()
            ctx.detect_describe_host(one, det, brief, NEEDED, CAND_CAPACITY)
            if i >= 20:
                lat.append(time.perf_counter() - t0)
        extras["single frame, host to host, one call (fd_detect_describe_host: one synchronisation)"] = {
            "latency_us_median": round(float(np.median(lat)) * 1e6, 1), "latency_us_p90": round(float(np.percentile(lat, 90)) * 1e6, 1),
            "mpixel_s": round(px / float(np.median(lat)) / 1e6, 1)}
        with fd.Context(local_rank) as own:      # the drop-in classes' situation: an object with its own context and stream
            lat = []
            for i in range(220):
                t0 = time.perf_counter()
                own.detect_describe_host(one, det, brief, NEEDED, CAND_CAPACITY)
                if i >= 20:
                    lat.append(time.perf_counter() - t0)
        extras["single frame, host to host, one call, on a context of its own"] = {
            "latency_us_median": round(float(np.median(lat)) * 1e6, 1), "latency_us_p90": round(float(np.percentile(lat, 90)) * 1e6, 1),
            "mpixel_s": round(px / float(np.median(lat)) / 1e6, 1)}
        ctx.bind_device(d_frames.data_ptr(), H, W, n)

        # ---- dense-map output modes (reported separately; 1 B/px in + the map out) ----
        resp_map = torch.empty((n, H, W), dtype=torch.float32, device=dev)
        score_map = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
        for name, prm, resp_ptr, score_ptr, bpp in (("harris candidates + dense response map (5 B/px)", fd.DetectParams(fd.HARRIS, 30.0, 20, 200), resp_map.data_ptr(), 0, 5.0),
                                                   ("fast candidates + dense score map (2 B/px)", fd.DetectParams(fd.FAST, THR, DIST, NEEDED, fast_n=args.fast_n), 0, score_map.data_ptr(), 2.0)):
            ctx.set_dense_outputs(resp_ptr, score_ptr)
            s_d = timed(lambda: ctx.compute_candidates(prm, 65536), steps2, 2)
            ctx.sync()
            extras[name] = {"mpixel_s": round(world * n * px * steps2 / s_d / 1e6, 1), "ms_per_step": round(s_d / steps2 * 1e3, 3),
                            "hbm_frac": round(n * px * bpp * steps2 / s_d / 1e9 / peak, 4), "bytes_per_px": bpp}
        ctx.set_dense_outputs(0, 0)
        del resp_map, score_map

        # ---- the other BASELINE.json configs at their named batch sizes, distinct frames, SURVEY.md 8(d) accounting ---------
        def host_pipeline_e2e(host_t, h2, w2, count, prm, kp_cap, cap, chunk):
            """detect through HostPipeline: pinned frames in, keypoints back.  Returns (seconds per step, h2d bytes, d2h bytes, counts)."""
            o_kp = torch.zeros((count, kp_cap, 16), dtype=torch.uint8).pin_memory()
            o_cnt = torch.zeros((count,), dtype=torch.int32).pin_memory()
            kp_v = o_kp.numpy().view(fd.KEYPOINT_DTYPE).reshape(count, kp_cap)
            p2 = HostPipeline(local_rank, chunk_frames=chunk)
            reps = 3
            s = wall_timed(lambda: p2.run(host_t.data_ptr(), h2, w2, count, prm, None, kp_v, o_cnt.numpy(), None, cap), reps, 1)
            p2.close()
            return s / reps, count * h2 * w2, count * kp_cap * 16 + count * 4, o_cnt.numpy().copy()

        def config_corner(key, title, w2, h2, count, prm, kind_name, cap, cpu_sample, per_frame_out_bytes, kernel_name, chunk):
            fr = extra_frames[key]
            host_t = torch.from_numpy(fr).pin_memory()
            dev_fr = host_t.to(dev, non_blocking=True)
            torch.cuda.synchronize()
            ctx.bind_device(dev_fr.data_ptr(), h2, w2, count)
            s_step = timed(lambda: ctx.detect(prm, cap), steps2, 2)
            ctx.sync()
            kpc, cdc = ctx.keypoint_counts(), ctx.candidate_counts()
            s_kernel = timed(lambda: ctx.compute_candidates(prm, cap), steps2, 2)
            out_bytes = per_frame_out_bytes(kpc, cdc)
            algo = count * w2 * h2 + out_bytes
            e_sec, e_h2d, e_d2h, e_cnt = host_pipeline_e2e(host_t, h2, w2, count, prm, int(prm.needed_feature_num), cap, chunk)
            obj = {"workload": title, "frames_per_gpu": count, "ms_per_step": round(s_step / steps2 * 1e3, 4),
                   "mpixel_s": round(world * count * w2 * h2 * steps2 / s_step / 1e6, 1), "frames_s": round(world * count * steps2 / s_step, 1),
                   "mean_keypoints": float(kpc.mean()), "mean_candidates": float(cdc.mean()),
                   "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": round(algo / (s_kernel / steps2) / 1e9, 2), "peak": peak, "unit": "GB/s",
                                "frac": round(algo / (s_kernel / steps2) / 1e9 / peak, 4), "step_frac": round(algo / (s_step / steps2) / 1e9 / peak, 4),
                                "algorithmic_bytes_per_launch": int(algo), "us_per_launch": round(s_kernel / steps2 * 1e6, 2),
                                "traffic": kernel_traffic(f"{kernel_name.split(' ')[0]}_{key}")},
                   "e2e": {"value": round(world * count * w2 * h2 / e_sec / 1e6, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": e_h2d, "d2h_bytes_per_step": e_d2h,
                           "ms_per_step": round(e_sec * 1e3, 3), "path": f"HostPipeline.run, chunks of {chunk} frames: fd_upload_frames (pinned) -> fd_detect -> fd_download_keypoints",
                           "equals_device_resident_run": bool(np.array_equal(e_cnt, kpc))}}
            ok_local = obj["e2e"]["equals_device_resident_run"]
            if rank == 0:
                sub = fr[:cpu_sample]
                c_sec, (c_kp, c_cand), c_kind = cpu_points(kind_name, sub, float(prm.min_valid_response), int(prm.min_feature_distance), int(prm.needed_feature_num), 12, 0, threads)
                obj["cpu_baseline"] = {"value": round(cpu_sample * w2 * h2 / c_sec / 1e6, 3), "unit": "Mpixel/s", "cores": threads, "kind": c_kind,
                                       "sample": f"first {cpu_sample} frames of the batch, one pass ({c_sec:.2f} s), {threads} host threads, DetectGoodFeatures only"}
                # ties between equal responses may swap which of two keypoints is kept, never how many candidates there are
                obj["cpu_totals"] = {"frames": cpu_sample, "candidates": {"reference": c_cand, "gpu": int(cdc[:cpu_sample].sum())},
                                     "keypoints": {"reference": c_kp, "gpu": int(kpc[:cpu_sample].sum())}}
                ok_local &= c_cand == int(cdc[:cpu_sample].sum()) and c_kp == int(kpc[:cpu_sample].sum())
            extras[title] = obj
            return dev_fr, host_t, ok_local

        shi = fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 1000)
        keep = config_corner("c2", "configs[2] shi_tomas top-1000, 1280x720 x 512 per GPU", 1280, 720, 512, shi, "shi", 131072, 128,
                             lambda kpc, cdc: int(kpc.sum()) * 8, "corner_tma_kernel<1> (fd_corner_tma.cu)", 64)
        ok &= keep[2]
        del keep
        har = fd.DetectParams(fd.HARRIS, 30.0, 20, 200)
        keep = config_corner("c3", "configs[3] harris candidates + greedy d 20 N 200, 3840x2160 x 64, untiled on one GPU", 3840, 2160, 64, har, "harris", 1 << 20, 16,
                             lambda kpc, cdc: int(cdc.sum()) * 12, "corner_tma_kernel<0> (fd_corner_tma.cu)", 8)
        ok &= keep[2]
        c3_dev, c3_host = keep[0], keep[1]
        c3 = extras["configs[3] harris candidates + greedy d 20 N 200, 3840x2160 x 64, untiled on one GPU"]
        c3["candidates_only_ms_per_step"] = round(c3["roofline"]["us_per_launch"] / 1e3, 4)

        # the same batch row-tiled over all GPUs of the box in one process (rank 0 drives every device; the other ranks wait on the host)
        torch.cuda.synchronize()
        host_barrier()
        if rank == 0:
            try:
                tiles = list(range(world)) if world > 1 else [0, 0]
                with fd.TiledDetector(tiles) as td:
                    td.scatter_device(c3_dev.data_ptr(), 2160, 3840, 64)
                    td.detect(har, (1 << 21) // len(tiles))
                    td.sync()
                    kp_t, cnt_t = td.keypoints(200)
                    ctx.bind_device(c3_dev.data_ptr(), 2160, 3840, 64)
                    ctx.detect(har, 1 << 20)
                    kp_u, cnt_u = ctx.keypoints(200)
                    same_t = bool(np.array_equal(cnt_t, cnt_u) and all(np.array_equal(kp_t[f, :cnt_t[f]], kp_u[f, :cnt_u[f]]) for f in range(64)))

                    def tiled_step():
                        td.exchange_halos()
                        td.detect(har, (1 << 21) // len(tiles))
                    for _ in range(2):
                        tiled_step()
                    td.sync()
                    reps = 20
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        tiled_step()
                    td.sync()
                    s_t = (time.perf_counter() - t0) / reps

                    def cand_step():
                        td.exchange_halos()
                        td.compute_candidates(har, (1 << 21) // len(tiles))
                    cand_step()
                    td.sync()
                    t0 = time.perf_counter()
                    for _ in range(reps):
                        cand_step()
                    td.sync()
                    s_c = (time.perf_counter() - t0) / reps
                    c3["row_tiled"] = {"tiles": len(tiles), "devices": sorted(set(tiles)), "ms_per_step_detect": round(s_t * 1e3, 4), "ms_per_step_candidates_and_gather": round(s_c * 1e3, 4),
                                       "untiled_one_gpu_ms_per_step": c3["ms_per_step"], "halo_bytes_per_step": td.halo_bytes,
                                       "equals_untiled": same_t, "mpixel_s": round(64 * 3840 * 2160 / s_t / 1e6, 1),
                                       "what": "per step (fd_tiled_exchange_halos + fd_tiled_detect): halo exchange (device-to-device peer copies of 3 rows per seam side and frame) -> "
                                               "fd_compute_candidates + rank histogram per tile -> first rank limits on device 0 -> first ranges compacted per tile and "
                                               "gathered on device 0 by peer reads -> one global selection (frames that need more are finished from a conditional full "
                                               "gather); ms_per_step_candidates_and_gather = fd_tiled_compute_candidates, which gathers every key; frames' own rows resident "
                                               f"on their GPUs; host clock around {reps} steps with a sync of every tile stream on both sides"}
                    ok &= same_t
            except Exception as e:   # reported and fatal for the parity flag only if it was a mismatch
                c3["row_tiled"] = {"error": repr(e)[:400]}
            torch.cuda.set_device(local_rank)
        host_barrier()
        del c3_dev, c3_host, keep

        # ---- configs[4]: LSD level-line field, 1920x1080 x 256 ----
        fr4 = extra_frames["c4"]
        n4, w4, h4 = len(fr4), 1920, 1080
        host4 = torch.from_numpy(fr4).pin_memory()
        dev4 = host4.to(dev, non_blocking=True)
        torch.cuda.synchronize()
        ctx.bind_device(dev4.data_ptr(), h4, w4, n4)
        lsd_plain, lsd_sorted = fd.LsdParams(20.0, 0), fd.LsdParams(20.0, 1)
        s_field = timed(lambda: ctx.lsd_field(lsd_plain), steps2, 2)
        s_sorted = timed(lambda: ctx.lsd_field(lsd_sorted), steps2, 2)
        ctx.sync()
        n_valid = int(sum(ctx.lsd_download(f, True)["n_valid"] for f in range(0, n4, max(1, n4 // 8)))) * max(1, n4 // 8)   # sampled estimate of the seed count
        algo4 = n4 * w4 * h4 * 9 + 4 * n_valid       # u8 in, norm + angle f32 out, 4 B per sorted seed index
        title4 = "configs[4] lsd level-line field, 1920x1080 x 256 per GPU"
        o4 = {"workload": title4, "frames_per_gpu": n4, "ms_per_step": round(s_sorted / steps2 * 1e3, 4), "field_only_ms_per_step": round(s_field / steps2 * 1e3, 4),
              "mpixel_s": round(world * n4 * w4 * h4 * steps2 / s_sorted / 1e6, 1), "frames_s": round(world * n4 * steps2 / s_sorted, 1),
              "valid_fraction": round(n_valid / (n4 * w4 * h4), 4),
              "roofline": {"bound": "hbm", "kernel": "lsd_kernel (fd_lsd.cu), field only", "achieved": round(algo4 / (s_field / steps2) / 1e9, 2), "peak": peak, "unit": "GB/s",
                           "frac": round(algo4 / (s_field / steps2) / 1e9 / peak, 4), "step_frac": round(algo4 / (s_sorted / steps2) / 1e9 / peak, 4),
                           "algorithmic_bytes_per_launch": int(algo4), "us_per_launch": round(s_field / steps2 * 1e6, 2), "traffic": kernel_traffic("lsd_kernel_c4"),
                           "note": "step = field + seed order (the reference's std::sort, feature_line_detector.cpp:92-94); frac = the field kernel alone"}}
        # e2e at the C ABI: pinned frames in; norm, angle and seed order of every frame back in pinned memory
        chunk4 = 32
        o_norm = torch.zeros((chunk4, h4, w4), dtype=torch.float32).pin_memory()
        o_ang = torch.zeros((chunk4, h4, w4), dtype=torch.float32).pin_memory()
        o_idx = torch.zeros((chunk4, h4 * w4), dtype=torch.int32).pin_memory()
        import ctypes as C
        nv = C.c_int32(0)

        def lsd_e2e():
            for s in range(0, n4, chunk4):
                m = min(chunk4, n4 - s)
                ctx.upload_ptr(host4.data_ptr() + s * h4 * w4, h4, w4, m)
                ctx.lsd_field(lsd_sorted)
                for f in range(m):
                    ctx._ck(ctx._lib.fd_lsd_download(ctx._h, f, C.c_void_p(o_norm[f].data_ptr()), C.c_void_p(o_ang[f].data_ptr()), C.c_void_p(o_idx[f].data_ptr()),
                                                     h4 * w4, C.byref(nv)))
        s_e4 = wall_timed(lsd_e2e, 2, 1) / 2
        o4["e2e"] = {"value": round(world * n4 * w4 * h4 / s_e4 / 1e6, 1), "unit": "Mpixel/s", "h2d_bytes_per_step": n4 * w4 * h4, "d2h_bytes_per_step": n4 * w4 * h4 * 8 + 4 * n_valid,
                     "ms_per_step": round(s_e4 * 1e3, 2), "path": f"chunks of {chunk4} frames: fd_upload_frames (pinned) -> fd_lsd_field (sorted) -> fd_lsd_download per frame (pinned); "
                     "8 B/px come back, so this leg is the PCIe link"}
        ctx.bind_device(d_frames.data_ptr(), H, W, n)
        if rank == 0:
            chk, c_kind = _checker()
            if hasattr(chk, "bench_lsd"):
                c_n = 64
                c_sec, c_tot = chk.bench_lsd(fr4[:c_n], 20.0, False, threads)
                o4["cpu_baseline"] = {"value": round(c_n * w4 * h4 / c_sec / 1e6, 3), "unit": "Mpixel/s", "cores": threads, "kind": c_kind,
                                      "sample": f"first {c_n} frames, ComputeLineLevelAngleMap only (feature_line_detector.cpp:56-97, sort included), one pass ({c_sec:.2f} s), {threads} host threads"}
            b3 = os.path.join(ROOT, "oracle", "_ref", "fd_route_b3")
            if os.path.exists(b3):
                # the reference's own FeatureLineDetector (compiled in place) with ComputeLineLevelAngleMap replaced as INTEGRATION.md B.3 says
                tmp = "/tmp/fd_b200_route_b3_frames.u8"
                fr4[:8].tofile(tmp)
                try:
                    r = subprocess.run([b3, tmp, str(h4), str(w4), "8", "200"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
                    o4["dropin_route_b3"] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-300:]}
                except Exception as e:
                    o4["dropin_route_b3"] = {"error": repr(e)[:300]}
        extras[title4] = o4
        del dev4, host4, o_norm, o_ang, o_idx

        # ---- NN detector post-processing: heat maps -> keypoints -> sampled 256-channel descriptors ----
        from feature_detector_b200.synth import synth_descriptor_volume, synth_heatmap
        n_nn = min(n, 1024)
        heat = torch.from_numpy(np.stack([synth_heatmap(W, H, rank * 8 + i) for i in range(8)])).to(dev)
        heat = heat.repeat((n_nn + 7) // 8, 1, 1)[:n_nn].contiguous()
        vol = torch.from_numpy(np.stack([synth_descriptor_volume(256, H // 8, W // 8, i) for i in range(2)])).to(dev)
        vol = vol.repeat((n_nn + 1) // 2, 1, 1, 1)[:n_nn].contiguous()
        nn_prm = fd.NnParams(0.1, 3, 15, 240)

        def nn_step():
            ctx.nn_select(heat.data_ptr(), H, W, n_nn, nn_prm, 65536)
            ctx.nn_sample_descriptors(vol.data_ptr(), 256, H // 8, W // 8)
        s_nn = timed(nn_step, steps2, 2)
        ctx.sync()
        extras["nn post-processing (heat map -> 240 keypoints -> 256-channel descriptors), 752x480 x %d" % n_nn] = {
            "mpixel_s": round(world * n_nn * px * steps2 / s_nn / 1e6, 1), "frames_s": round(world * n_nn * steps2 / s_nn, 1),
            "ms_per_step": round(s_nn / steps2 * 1e3, 3), "mean_keypoints": float(ctx.keypoint_counts().mean()),
            "hbm_frac": round(n_nn * px * 4 * steps2 / s_nn / 1e9 / peak, 4), "bytes_per_px": 4.0}
        del heat, vol

        # ---- Hamming matching of the packed descriptors (no reference counterpart) ----
        if hasattr(ctx, "match_selected"):
            try:
                ctx.bind_device(d_frames.data_ptr(), H, W, n)
                device_step()
                s_m = timed(lambda: ctx.match_selected(), steps2, 2)
                extras["hamming matching of consecutive frames' BRIEF-256 descriptors (no reference counterpart)"] = {
                    "ms_per_step": round(s_m / steps2 * 1e3, 4), "frame_pairs": n - 1, "pairs_s": round(world * (n - 1) * steps2 / s_m, 1)}
            except Exception as e:
                extras["hamming matching"] = {"error": repr(e)[:300]}

        # ---- the one collective of the batch path: the optional gather of the result slots ----
        if world > 1 or os.environ.get("FD_BENCH_GATHER") == "1":
            try:
                from feature_detector_b200 import sharding
                ctx.bind_device(d_frames.data_ptr(), H, W, n)
                device_step()
                ctx.sync()
                slots = sharding.device_results(ctx, n, dev, True)     # torch views of the context's device buffers

                def gather_step():
                    for t in slots:
                        sharding.gather_frame_slots(t, world * n, rank, world)

                def gather_packed_step():
                    sharding.gather_frame_records(list(slots), world * n, rank, world)
                per_rank = n * NEEDED * (16 + 32) + n * 4
                name = "keypoint gather (all_gather of keypoint, descriptor and count slots over NCCL)"
                s_g = timed(gather_step, 10, 2)
                extras[name] = {"ms_per_gather": round(s_g / 10 * 1e3, 3), "bytes_per_rank": per_rank,
                                "gbytes_s_received_per_rank": round(per_rank * (world - 1) * 10 / s_g / 1e9, 2)}
                s_p = timed(gather_packed_step, 10, 2)
                extras[name]["packed_one_all_gather"] = {"ms_per_gather": round(s_p / 10 * 1e3, 3),
                                                         "gbytes_s_received_per_rank": round(per_rank * (world - 1) * 10 / s_p / 1e9, 2)}
            except Exception as e:   # reported, never fatal: the gather is not part of the metric
                extras.setdefault("keypoint gather (all_gather of keypoint, descriptor and count slots over NCCL)", {})["error"] = repr(e)[:300]

    parity["all_ok"] = all_true(bool(ok))
    if rank == 0:
        line = {
            "metric": "Mpixel/s (FAST+NMS+BRIEF, 752x480)", "value": round(value, 2), "unit": "Mpixel/s",
            "frames_per_s": round(world * n * args.steps / sec, 1), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(sec / args.steps * 1e3, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(args, n),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": round(e2e_value, 2), "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_sec / e2e_steps * 1e3, 3), "steps": e2e_steps,
                    "path": f"HostPipeline.run: chunks of {args.chunk} frames on two contexts; per chunk fd_upload_frames (pinned host) -> fd_detect -> "
                            "fd_describe_selected -> fd_download_keypoints + fd_download_descriptors (pinned host)",
                    "matches_device_resident_run": e2e_match, "h2d_ceiling": h2d_ceiling},
            "gpu_launches": int(launches), "clocks": clocks, "parity_checks": parity,
            "host_binding": (f"rank threads and pinned buffers on the {len(near)} CPUs local to the GPU (nvmlDeviceGetCpuAffinity)" if near else "default"),
            "mean_keypoints_per_frame": float(kp_counts.mean()), "mean_candidates_per_frame": float(cand_counts.mean()),
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if not parity["all_ok"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
