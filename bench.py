#!/usr/bin/env python
"""bench.py -- headline benchmark of the dense feature-detection hot path on B200.

Metric (BASELINE.json): Mpixel/s (and frames/s) of FAST + greedy NMS selection + BRIEF-256 on a batch of
1024 synthetic 752x480 grayscale frames per GPU (configs[1]; option values are the reference demo's:
threshold 10, min distance 20, 200 features, BRIEF length 256 / half patch 8).  A "step" is one pass of
that path over the whole batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun, one rank per GPU; frames are independent, so every rank runs its own batch
with no data-path collective (weak scaling) and the timing is the max over ranks.  One JSON line on rank 0.
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
def make_frames(n: int, start: int) -> np.ndarray:
    """frames start .. start+n-1 of the SURVEY.md 8d generator, cached under /tmp, built in worker processes
    (must run before CUDA is initialised in this process)."""
    from feature_detector_b200.synth import SEED, synth
    cache = f"/tmp/fd_b200_synth_{W}x{H}_{SEED}_{start}_{n}.npy"
    if os.path.exists(cache):
        try:
            a = np.load(cache)
            if a.shape == (n, H, W):
                return a
        except Exception:
            pass
    import multiprocessing as mp
    workers = max(1, min(32, (os.cpu_count() or 1)))
    with mp.get_context("fork").Pool(workers) as pool:
        frames = pool.starmap(synth, [(W, H, start + i) for i in range(n)], chunksize=8)
    a = np.stack(frames)
    try:
        np.save(cache, a)
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


def bind_near_gpu(local_rank: int):
    """Multi-GPU runs: keep this rank's threads -- and so the pinned host buffers it is about to allocate (first touch) -- on the
    CPUs the driver reports as local to its GPU.  Eight ranks pulling frames across sockets is what the end-to-end leg is bound
    by on an 8-GPU box.  Best effort: any failure leaves the default placement."""
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


def cpu_reference_mpx(frames: np.ndarray, fast_n: int, threads: int):
    """The reference's own CPU path (oracle/_ref when present, else the C port) on `frames`: DetectGoodFeatures + Compute."""
    from oracle.bindings import FAST, Port, Ref, have_ref
    if have_ref():
        chk, kind = Ref(), "reference"
    else:
        chk, kind = Port(), "port"
    sec, totals = chk.bench_points(FAST, frames, THR, DIST, NEEDED, fast_n=fast_n, brief_length=BRIEF_LEN, brief_half_patch=BRIEF_HALF, n_threads=threads)
    mpx = frames.shape[0] * H * W / sec / 1e6
    return mpx, kind, sec, totals


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
    ap.add_argument("--steps", type=int, default=500, help="timed steps (0.8 ms each at N=1: the default keeps the timed region long enough to sample clocks)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--fast-n", dest="fast_n", type=int, default=9, help="9 = full segment test on every pixel (configs[1] 'FAST-9'); 12 = reference default")
    ap.add_argument("--chunk", type=int, default=128, help="frames per chunk of the host pipeline (e2e leg)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (kN=12, Harris, Shi-Tomasi, LSD)")
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
    cpu = None
    if rank == 0 and world == 1:
        threads = host_threads()
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
        n1 = min(sample, 128)
        mpx1, _, sec1, _ = cpu_reference_mpx(frames[:n1], args.fast_n, 1)
        cpu["single_thread_mpixel_s"] = round(mpx1, 3)
        cpu["single_thread_sample"] = f"first {n1} frames, {sec1:.2f} s"

    import torch
    import torch.distributed as dist
    import feature_detector_b200 as fd

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
    algo_bytes = n * px + int(cand_counts.sum()) * 8 + n * 4  # u8 frame in, 8-byte candidate keys + counters out
    peak, peak_src = measured_peak_gbs()
    achieved = algo_bytes / per_launch / 1e9
    roofline = {"bound": "hbm", "kernel": "fdb::fast_sparse_kernel (fd_fast_sparse.cu; thr 10 leaves s_min >= 7, so the sparse form runs)", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes, "us_per_launch": round(per_launch * 1e6, 2),
                "note": "1 B/px kernel: instruction-issue bound (80 % issue slots busy, ALU pipe the fullest), see DESIGN.md section 4 and "
                        "profiles/r1_headline_step_ncu_summary.txt; traffic = dram bytes per launch from that ncu capture"}
    prof = os.path.join(ROOT, "profiles", "fast_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                roofline["traffic"] = json.load(f).get(f"kN{args.fast_n}")
        except Exception:
            pass

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
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_sec = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n * px * e2e_steps / e2e_sec / 1e6
    e2e_match = bool(np.array_equal(cnt_host, kp_counts))  # the pipelined result equals the device-resident run
    pipe.close()

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

        def lsd_step():
            ctx.lsd_field(fd.LsdParams(20.0, 0))
        nl = min(n, 256)
        ctx.bind_device(d_frames.data_ptr(), H, W, nl)
        s = timed(lsd_step, steps2, 2)
        extras["lsd_field (norm+angle maps, no seed sort)"] = {"mpixel_s": round(world * nl * px * steps2 / s / 1e6, 1),
                                                              "hbm_frac_9Bpx": round(nl * px * 9 * steps2 / s / 1e9 / peak, 4)}

        # ---- the other BASELINE.json configs at their named shapes (small batches; parity is in tests/) -------------
        from feature_detector_b200.synth import synth

        def named(name, w2, h2, count, fn, bytes_per_px):
            fr = np.stack([synth(w2, h2, rank * count + i) for i in range(min(count, 4))])
            fr = np.concatenate([fr] * ((count + len(fr) - 1) // len(fr)))[:count]   # cyclic copies of 4 generated frames
            dev_fr = torch.from_numpy(fr).to(dev)
            ctx.bind_device(dev_fr.data_ptr(), h2, w2, count)
            s2 = timed(fn, steps2, 2)
            ctx.sync()
            extras[name] = {"mpixel_s": round(world * count * w2 * h2 * steps2 / s2 / 1e6, 1), "frames_s": round(world * count * steps2 / s2, 1),
                            "ms_per_step": round(s2 / steps2 * 1e3, 3), "frames": count,
                            "hbm_frac": round(count * w2 * h2 * bytes_per_px * steps2 / s2 / 1e9 / peak, 4), "bytes_per_px": bytes_per_px}
            return dev_fr

        shi = fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 1000)
        keep = named("configs[2] shi_tomas top-1000, 1280x720 x 256", 1280, 720, 256, lambda: ctx.detect(shi, 131072), 1.0)
        extras["configs[2] shi_tomas top-1000, 1280x720 x 256"]["mean_keypoints"] = float(ctx.keypoint_counts().mean())
        har = fd.DetectParams(fd.HARRIS, 30.0, 20, 200)
        keep = named("configs[3] harris candidates, 3840x2160 x 64 (untiled, one GPU)", 3840, 2160, 64, lambda: ctx.compute_candidates(har, 1 << 20), 1.8)
        extras["configs[3] harris candidates, 3840x2160 x 64 (untiled, one GPU)"]["mean_candidates"] = float(ctx.candidate_counts().mean())
        keep = named("configs[3] harris + select, 3840x2160 x 64", 3840, 2160, 64, lambda: ctx.detect(har, 1 << 20), 1.8)
        keep = named("configs[4] lsd field, 1920x1080 x 64", 1920, 1080, 64, lsd_step, 9.0)
        lsd_sorted = fd.LsdParams(20.0, 1)
        keep = named("configs[4] lsd field + seed order, 1920x1080 x 64", 1920, 1080, 64, lambda: ctx.lsd_field(lsd_sorted), 9.2)
        del keep

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

    if rank == 0:
        line = {
            "metric": "Mpixel/s (FAST+NMS+BRIEF, 752x480)", "value": round(value, 2), "unit": "Mpixel/s",
            "frames_per_s": round(world * n * args.steps / sec, 1), "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(sec / args.steps * 1e3, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(args, n),
            "hbm_frac_1Bpx": round(n * px * args.steps / sec / 1e9 / peak, 4),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": round(e2e_value, 2), "unit": "Mpixel/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e2e_sec / e2e_steps * 1e3, 3), "steps": e2e_steps,
                    "path": f"HostPipeline.run: chunks of {args.chunk} frames on two contexts; per chunk fd_upload_frames (pinned host) -> fd_detect -> "
                            "fd_describe_selected -> fd_download_keypoints + fd_download_descriptors (pinned host)",
                    "matches_device_resident_run": e2e_match},
            "gpu_launches": int(launches), "clocks": clocks,
            "host_binding": (f"rank threads and pinned buffers on the {len(near)} CPUs local to the GPU (nvmlDeviceGetCpuAffinity)" if near else "default"),
            "mean_keypoints_per_frame": float(kp_counts.mean()), "mean_candidates_per_frame": float(cand_counts.mean()),
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
