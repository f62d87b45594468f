"""The reference's UNMODIFIED OpenCL kernels on the GPU of this box (NVIDIA's OpenCL runtime) beside the CUDA path:
parity metrics on a small scene, then the bench scene (512^3, 1920x1080) timed both ways.

    python tests/probes/ref_opencl_bench.py [n_big=512] [frames=16]     -> JSON lines on stdout

TEST INFRASTRUCTURE (uses oracle/_ref/libref_ocl.so and the CPU oracle); not part of the product or of bench.py's arms."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle_lib as o  # noqa: E402
import ref_ocl_lib as R  # noqa: E402
from cl_volume_renderer_b200 import api, synth  # noqa: E402


def cache_metrics(a, b):
    neq = a != b
    nd = int(neq.sum())
    touched = int(((a != 0) | (b != 0)).sum())
    maxd = 0
    if nd:
        idx = np.flatnonzero(neq)
        maxd = int(np.abs(a[idx].astype(np.int32) - b[idx].astype(np.int32)).max())
    return {"lanes_identical": 1.0 - nd / a.size, "touched_lanes": touched, "touched_identical": 1.0 - nd / max(touched, 1),
            "max_abs_diff": maxd, "tokens_identical": bool(np.array_equal(a[3::4], b[3::4]))}


def frame_metrics(a, b):
    mse = float(np.mean((a[..., :3].astype(np.float64) - b[..., :3].astype(np.float64)) ** 2))
    return {"alpha_identical": bool(np.array_equal(a[..., 3], b[..., 3])), "psnr_db": 99.0 if mse == 0 else float(10 * np.log10(255.0 ** 2 / mse)),
            "max_abs_diff": int(np.abs(a.astype(int) - b.astype(int)).max()), "identical_pixels": float((a == b).all(axis=-1).mean())}


def main():
    n_big = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    if not R.available():
        print(json.dumps({"opencl": "unavailable", "why": R.error()}))
        return
    print(json.dumps({"opencl": R.info()}), flush=True)
    tf = synth.default_tf()
    tf_src = api.tf_format(tf)
    ctx = api.Context(0)

    # ---- parity on a small scene: reference-on-GPU vs CPU oracle vs CUDA -------------------------------------------------
    for (n, W, H, nf, cam, nearest) in [(64, 160, 120, 6, "default", True), (96, 200, 136, 3, "closeup", True), (128, 320, 240, 16, "default", True),
                                        (64, 160, 120, 6, "default", False), (96, 200, 136, 3, "closeup", False), (128, 320, 240, 16, "default", False)]:
        R.set_nearest(nearest)
        v, envimg = synth.synth_ct(n), synth.synth_env(128, 64)
        pos, d = synth.default_camera(n) if cam == "default" else synth.closeup_camera(n)
        seeds = synth.glibc_rand(nf)
        sc = R.Scene(v, envimg, tf_src, W, H)
        st, _ = R.fetch_stats(v)
        want_sdf = o.sdf_build(v, tf)[0]
        ocl_frame, _ = sc.render(pos, d, seeds)
        ocl_cache = sc.cache()
        orc = o.Renderer(v, envimg, tf, W, H)
        for s in seeds:
            orc_frame = orc.render_frame(pos, d, s)
        vol, env = api.Volume(ctx, v), api.EnvMap(ctx, envimg)
        r = api.Renderer(ctx, W, H)
        if not nearest:
            r.set_sampling(api.VR_SAMPLING_HW_LINEAR)   # the CUDA path sampling through the texture unit, like the kernels as shipped
        r.image_set(vol, env); r.next_event_code_set(tf_src); r.flush_changes()
        for s in seeds:
            cu_frame = r.render_frame(pos, d, s)
        cu_cache = r.cache_download()
        print(json.dumps({"parity": f"{n}^3 {W}x{H} {nf} frames {cam}", "sampler": "CLK_FILTER_NEAREST (one-token substitution)" if nearest else "as shipped (CLK_FILTER_LINEAR on integer images)",
                          "cuda_sampling": "VR_SAMPLING_NEAREST" if nearest else "VR_SAMPLING_HW_LINEAR",
                          "sdf_opencl_eq_oracle": bool(np.array_equal(sc.sdf(), want_sdf)), "sdf_opencl_eq_cuda": bool(np.array_equal(sc.sdf(), r.sdf_download())),
                          "sdf_iterations": sc.sdf_iterations, "stats_opencl": st, "stats_oracle": o.fetch_stats(v), "stats_cuda": vol.stats(),
                          "cache_opencl_vs_oracle": cache_metrics(ocl_cache, orc.cache), "cache_opencl_vs_cuda": cache_metrics(ocl_cache, cu_cache),
                          "cache_cuda_vs_oracle": cache_metrics(cu_cache, orc.cache),
                          "frame_opencl_vs_oracle": frame_metrics(ocl_frame, orc_frame), "frame_opencl_vs_cuda": frame_metrics(ocl_frame, cu_frame)}), flush=True)
        r.close(); env.close(); vol.close(); sc.close()

    # ---- volume kernels: tf_sort_values, bilateral_filter, apply_clip ----------------------------------------------------------
    for nearest in (True, False):
        R.set_nearest(nearest)
        for dims in ((45, 37, 29), (128, 128, 128), (n_big, n_big, n_big)):
            v = synth.synth_ct(0, dims=dims) if dims[0] != dims[1] else synth.synth_ct(dims[0])
            vol = api.Volume(ctx, v)
            st = vol.stats()
            rng = [float(x) for x in st]
            small = v.size <= 128 ** 3
            hb, h_ms = R.histogram(v, 500, 500, rng)
            cu_h = vol.histogram(500, 500, rng)
            fb, f_ms = R.bilateral(v)
            cs = tuple(max(d // 8, 1) for d in dims)
            size = tuple(dims[k] - 2 * cs[k] for k in range(3))
            cb, c_ms = R.clip(v, cs, size)
            vol.clip(cs, tuple(cs[k] + size[k] for k in range(3)))
            cu_c = vol.download()
            vol2 = api.Volume(ctx, v)
            vol2.filter()
            cu_f = vol2.download()
            row = {"volume_ops": "x".join(map(str, dims)), "sampler": "CLK_FILTER_NEAREST (one-token substitution)" if nearest else "as shipped",
                   "reference_opencl_ms": {"tf_sort_values": h_ms, "bilateral_filter": f_ms, "apply_clip": c_ms},
                   "histogram_opencl_vs_cuda": {"bins_differing": int((hb != cu_h).sum()), "voxels_moved": int(np.abs(hb.astype(np.int64) - cu_h.astype(np.int64)).sum() // 2),
                                                "total_opencl": int(hb.sum()), "total_cuda": int(cu_h.sum())},
                   "bilateral_opencl_vs_cuda": {"identical": float((fb == cu_f).mean()), "max_abs_diff": int(np.abs(fb.astype(np.int32) - cu_f.astype(np.int32)).max())},
                   "clip_opencl_eq_cuda": bool(np.array_equal(cb, cu_c))}
            if small:
                ob = o.histogram(v, 500, 500, rng)
                of = o.bilateral(v)
                row["histogram_opencl_vs_oracle"] = {"bins_differing": int((hb != ob).sum()), "voxels_moved": int(np.abs(hb.astype(np.int64) - ob.astype(np.int64)).sum() // 2)}
                row["bilateral_opencl_vs_oracle"] = {"identical": float((fb == of).mean()), "max_abs_diff": int(np.abs(fb.astype(np.int32) - of.astype(np.int32)).max())}
            print(json.dumps(row), flush=True)
            vol2.close(); vol.close()

    # ---- the bench scene, reference-on-GPU vs CUDA ---------------------------------------------------------------------------
    for nearest in (False, True):
        bench_scene(ctx, n_big, frames, tf_src, nearest)
    ctx.close()


def bench_scene(ctx, n_big, frames, tf_src, nearest):
    R.set_nearest(nearest)
    n, W, H = n_big, 1920, 1080
    v, envimg = synth.synth_ct(n), synth.synth_env(2048, 1024)
    pos, d = synth.default_camera(n)
    seeds = synth.glibc_rand(frames)
    t0 = time.perf_counter()
    sc = R.Scene(v, envimg, tf_src, W, H)
    create_s = time.perf_counter() - t0
    sc.render(pos, d, seeds[:2], readback=False)  # warm-up
    sc.reset()
    _, ms_pull = sc.render(pos, d, seeds, pull_every_frame=True)      # renderer.cpp:131-158: blocking pull per frame
    sc.reset()
    _, ms_nopull = sc.render(pos, d, seeds, pull_every_frame=False, readback=False)
    ocl_cache = sc.cache()
    _, stats_ms = R.fetch_stats(v)
    vol, env = api.Volume(ctx, v), api.EnvMap(ctx, envimg)
    r = api.Renderer(ctx, W, H)
    if not nearest:
        r.set_sampling(api.VR_SAMPLING_HW_LINEAR)   # like for like: the CUDA path sampling through the texture unit
    r.image_set(vol, env); r.next_event_code_set(tf_src)
    ctx.synchronize(); t0 = time.perf_counter()
    r.flush_changes(); ctx.synchronize()
    cu_flush_ms = 1e3 * (time.perf_counter() - t0)
    r.render_frames(pos, d, seeds[:2], readback=False); r.reset_cache(); ctx.synchronize()
    t0 = time.perf_counter()
    r.render_frames(pos, d, seeds, readback=False); ctx.synchronize()
    cu_ms = 1e3 * (time.perf_counter() - t0)
    cu_cache = r.cache_download()
    r.reset_cache(); ctx.synchronize()
    hf = r.host_frame()
    t0 = time.perf_counter()
    for s in seeds:
        r.render_frame(pos, d, s, out=hf)
    cu_pull_ms = 1e3 * (time.perf_counter() - t0)
    samples = W * H * frames
    print(json.dumps({"bench_scene": f"{n}^3 {W}x{H} {frames} frames from a reset cache, default camera",
                      "sampler": "CLK_FILTER_NEAREST (one-token substitution)" if nearest else "as shipped (CLK_FILTER_LINEAR on integer images)",
                      "reference_opencl_on_this_gpu": {
                          "sdf_build_ms": sc.sdf_ms, "sdf_iterations": sc.sdf_iterations, "sdf_jit_ms": sc.sdf_jit_ms, "render_jit_ms": sc.render_jit_ms,
                          "scene_create_s": create_s, "fetch_stats_ms": stats_ms,
                          "render_ms_per_frame_with_pull": ms_pull / frames, "msamples_per_s_with_pull": samples / ms_pull / 1e3,
                          "render_ms_per_frame_kernel_only": ms_nopull / frames, "msamples_per_s_kernel_only": samples / ms_nopull / 1e3},
                      "cuda": {"sampling": "VR_SAMPLING_NEAREST" if nearest else "VR_SAMPLING_HW_LINEAR (k_trace<LINEAR>, thread per pixel and frame)",
                               "flush_ms_cache_reset_plus_sdf": cu_flush_ms, "render_ms_per_frame_batched": cu_ms / frames,
                               "msamples_per_s_batched": samples / cu_ms / 1e3, "render_ms_per_frame_with_pull": cu_pull_ms / frames,
                               "msamples_per_s_with_pull": samples / cu_pull_ms / 1e3},
                      "sdf_opencl_eq_cuda": bool(np.array_equal(sc.sdf(), r.sdf_download())),
                      "cache_opencl_vs_cuda": cache_metrics(ocl_cache, cu_cache)}), flush=True)
    r.close(); env.close(); vol.close(); sc.close()


if __name__ == "__main__":
    main()
