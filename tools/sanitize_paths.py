"""Drive every kernel of libfd_b200.so once on small inputs, for compute-sanitizer (the reference has no
race detection or sanitizers; here `compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck` on this script is that row).

No oracle and no assertions on values -- parity is the test-suite's job; this only has to reach every kernel instantiation:
sparse FAST (plain, pre-check, masked), dense FAST (score map, masks), both corner kernels (Harris, Shi-Tomasi, masks, response
map), both selection forms, BRIEF (windowed / frame fetch, integral / fractional, both sampling modes, float form), the LSD field
with its seed order (ordinary frames and a ramp, which takes the in-place bucket sort), the NN post-processing, row tiles.

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_paths.py
    SANITIZE_QUICK=1 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_paths.py

Device buffers that are not context-owned come from torch with its caching allocator switched off, so every buffer is its own
cudaMalloc and an out-of-bounds access cannot hide inside a pool.
"""
import os
import sys

os.environ.setdefault("PYTORCH_NO_CUDA_MEMORY_CACHING", "1")
os.environ.setdefault("FD_B200_GUARD", "1")   # red zones round every context-owned buffer (include/fd_b200.h, fd_debug_check_guards)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    import torch
    import feature_detector_b200 as fd
    from feature_detector_b200 import tiling
    from feature_detector_b200.synth import synth, synth_descriptor_volume, synth_heatmap

    dev = torch.device("cuda", 0)
    quick = os.environ.get("SANITIZE_QUICK") == "1"   # racecheck is two orders of magnitude slower than memcheck: small frames only
    shapes = [(333, 217, 3), (752, 480, 2), (64, 40, 2), (131, 9, 1), (65535, 7, 1), (7, 65535, 1)]   # (cols, rows, frames); odd widths, a sliver, the coordinate limit both ways
    if quick:
        shapes = [(333, 217, 2), (64, 40, 2), (131, 9, 1)]
    done = []

    canaries = []

    def fenced(shape, dtype):
        """A caller-owned device buffer between two 256-byte fences of 0xA5 (its own cudaMalloc; verified at the end)."""
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        raw = torch.full((nbytes + 512,), 0xA5, dtype=torch.uint8, device=dev)
        canaries.append((raw, nbytes))
        return raw[256:256 + nbytes].view(dtype).view(shape)

    def existing(w, h, n, seed):
        rng = np.random.default_rng(seed)
        return np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1).astype(np.float32)

    def detectors(ctx, tag):
        for w, h, n in shapes:
            frames = np.stack([synth(w, h, i) for i in range(n)])
            ctx.upload(frames)
            for with_pre in (False, True):
                ctx.set_existing_features([existing(w, h, 7 + f, f) for f in range(n)] if with_pre else [])
                for prm in (fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9), fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=12),
                            fd.DetectParams(fd.FAST, 0.1, 15, 200, fast_n=12), fd.DetectParams(fd.FAST, 5.0, 9, 50, fast_n=9, fast_min_pixel_diff=7),
                            fd.DetectParams(fd.HARRIS, 30.0, 20, 200), fd.DetectParams(fd.HARRIS, 0.1, 15, 200),
                            fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 1000)):
                    ctx.detect(prm)
                    ctx.keypoints(max(int(prm.needed_feature_num), 1))
                    ctx.candidates(0)
                    ctx.describe_selected(fd.BriefParams(256, 8))
                    ctx.descriptors(max(int(prm.needed_feature_num), 1))
                    ctx.match_selected()                                   # Hamming matching of consecutive frames (fd_match.cu)
                    ctx.matches(max(int(prm.needed_feature_num), 1))
                    if not with_pre:                                       # the one-call host-to-host form shares every buffer with the calls above
                        ctx.detect_describe_host(frames, prm, fd.BriefParams(256, 8), max(int(prm.needed_feature_num), 1))
                # dense outputs
                resp = fenced((n, h, w), torch.float32)
                score = fenced((n, h, w), torch.uint8)
                ctx.set_dense_outputs(resp.data_ptr(), score.data_ptr())
                ctx.compute_candidates(fd.DetectParams(fd.HARRIS, 0.1, 15, 200))
                ctx.compute_candidates(fd.DetectParams(fd.SHI_TOMAS, 0.1, 15, 200))
                ctx.compute_candidates(fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9))
                ctx.sync()
                ctx.set_dense_outputs(0, 0)
            ctx.set_existing_features([])
            # a candidate slot that overflows must be reported, not written past
            try:
                ctx.detect(fd.DetectParams(fd.FAST, 0.1, 15, 10), 64)
                ctx.keypoints(10)
            except fd.FdError as e:
                assert e.status == 5, e
            # pitched device frames
            pitch = (w + 31) // 32 * 32 + 32
            d = torch.zeros((n, h, pitch), dtype=torch.uint8, device=dev)
            d[:, :, :w] = torch.from_numpy(frames).to(dev)
            ctx.bind_device(d.data_ptr(), h, w, n, pitch, pitch * h)
            ctx.detect(fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9))
            ctx.detect(fd.DetectParams(fd.HARRIS, 30.0, 20, 200))
            ctx.keypoints(200)
            # unaligned bound frames (the context re-pitches)
            d1 = torch.from_numpy(frames).to(dev)
            ctx.bind_device(d1.data_ptr(), h, w, n)
            ctx.detect(fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 100))
            ctx.keypoints(100)
        done.append(f"detectors[{tag}]")

    def brief(ctx):
        w, h = 333, 217
        frames = np.stack([synth(w, h, i) for i in range(2)])
        ctx.upload(frames)
        rng = np.random.default_rng(5)
        pts = [np.concatenate([np.stack([rng.uniform(0, w, 40), rng.uniform(0, h, 40)], 1),                       # fractional, some in the border
                               np.array([[19, 19], [w - 19, h - 19], [w - 20, h - 20], [19.5, 30.25], [w / 2, h / 2]]),  # last admissible rows / columns
                               np.stack([rng.integers(19, w - 19, 20), rng.integers(19, h - 19, 20)], 1)]).astype(np.float32) for _ in range(2)]
        for prm in (fd.BriefParams(256, 8), fd.BriefParams(128, 8), fd.BriefParams(256, 12), fd.BriefParams(77, 5, fd.SAMPLE_TRUNCATE)):
            cap = ctx.describe_points(prm, pts)
            ctx.descriptors(cap)
            ctx.descriptors_float(cap, prm.length)
        ints = [np.round(p) for p in pts]
        cap = ctx.describe_points(fd.BriefParams(256, 8), ints)
        ctx.descriptors(cap)
        done.append("brief")

    def lsd(ctx):
        for w, h, n in shapes + ([] if quick else [(1920, 1080, 1)]):
            if h < 3:
                continue
            ctx.upload(np.stack([synth(w, h, i) for i in range(n)]))
            for want in (0, 1):
                ctx.lsd_field(fd.LsdParams(20.0, want))
                ctx.lsd_download(0, bool(want))
            # caller-owned outputs
            norm, angle = fenced((n, h, w), torch.float32), fenced((n, h, w), torch.float32)
            order, n_valid = fenced((n, h, w), torch.int32), fenced((n,), torch.int32)
            ctx.lsd_field(fd.LsdParams(20.0, 1), norm.data_ptr(), angle.data_ptr(), order.data_ptr(), n_valid.data_ptr())
            ctx.sync()
        r, c = np.mgrid[0:200, 0:300]
        ramp = ((5 * r + 3 * c) & 0xFF).astype(np.uint8)   # one magnitude bucket holds nearly every seed: the in-place bucket sort
        ctx.upload(ramp)
        ctx.lsd_field(fd.LsdParams(2.0, 1))
        ctx.lsd_download(0)
        done.append("lsd")

    def nn(ctx):
        for w, h, n in [(333, 217, 2), (160, 120, 3)]:
            maps = np.stack([synth_heatmap(w, h, 40 + f, 1.0 / 32) for f in range(n)])
            d_maps = torch.from_numpy(maps).to(dev)
            for with_pre, prm in ((False, fd.NnParams(0.1, 3, 15, 240)), (True, fd.NnParams(0.02, 9, 40, 5)), (False, fd.NnParams(0.3, 0, 2, 1000))):
                ctx.set_existing_features([existing(w, h, 5 + f, f) for f in range(n)] if with_pre else [])
                ctx.nn_select(d_maps.data_ptr(), h, w, n, prm, 0)
                ctx.keypoints(max(int(prm.max_features), 1))
                for ch in (256, 128):
                    vol = np.stack([synth_descriptor_volume(ch, h // 8, w // 8, f) for f in range(n)])
                    d_vol = torch.from_numpy(vol).to(dev)
                    ctx.nn_sample_descriptors(d_vol.data_ptr(), ch, h // 8, w // 8)
                    ctx.nn_descriptors(max(int(prm.max_features), 1))
                    ctx.nn_descriptors_at(d_vol.data_ptr(), ch, h // 8, w // 8,
                                          [np.array([[0, 0], [w - 1, h - 1], [3.5, 7.25], [w - 9, 2]], np.float32) for _ in range(n)])
            ctx.set_existing_features([])
        done.append("nn")

    def tiles(ctx):
        for (w, h), n_tiles in (((333, 217), 3), ((160, 40), 7)):
            frame = torch.from_numpy(synth(w, h, 1)).to(dev)
            for prm in (fd.DetectParams(fd.HARRIS, 30.0, 20, 200), fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9),
                        fd.DetectParams(fd.FAST, 0.1, 15, 200, fast_n=12)):
                tiling.detect_tiled_local(ctx, frame, n_tiles, prm)
        done.append("tiles")

    checked = 0
    with fd.Context(0) as ctx:
        for part in (lambda: detectors(ctx, "default kernels"), lambda: brief(ctx), lambda: lsd(ctx), lambda: nn(ctx), lambda: tiles(ctx)):
            part()
            checked += ctx.check_guards()
    # the alternative instantiations behind the testing knobs
    os.environ.update(FD_B200_FAST_DENSE="1", FD_B200_CORNER_STREAM="1", FD_B200_SELECT_CELLS_MIN="0")
    with fd.Context(0) as ctx:
        detectors(ctx, "dense FAST, streaming corner, per-cell selection")
        checked += ctx.check_guards()
    os.environ.update(FD_B200_SELECT_CELLS_MIN=str(1 << 30))
    with fd.Context(0) as ctx:
        detectors(ctx, "dense FAST, streaming corner, per-candidate selection")
        checked += ctx.check_guards()
    # the check has to be able to fail: one byte written right behind the candidate keys must be reported, with the buffer's name
    with fd.Context(0) as ctx:
        from cuda.bindings import runtime as cudart
        ctx.upload(synth(64, 40, 0))
        ctx.compute_candidates(fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9), 100)
        ctx.sync()
        assert ctx.check_guards() > 0
        keys_ptr, _, cap = ctx.device_candidates()
        err, = cudart.cudaMemset(keys_ptr + cap * 8, 0, 1)
        assert int(err) == 0, err
        try:
            ctx.check_guards()
            raise AssertionError("an overwrite behind ctx->keys went unnoticed")
        except fd.FdError as e:
            assert "keys" in str(e) and "after" in str(e), e
            done.append(f"self-test ({e})")
    torch.cuda.synchronize()
    for raw, nbytes in canaries:
        fence = torch.cat([raw[:256], raw[256 + nbytes:]])
        assert bool((fence == 0xA5).all()), f"a kernel wrote outside a caller-owned buffer of {nbytes} bytes"
    print(f"sanitize_paths ok ({checked} red-zone checks of context buffers, {len(canaries)} of caller buffers):", ", ".join(done), flush=True)


if __name__ == "__main__":
    main()
