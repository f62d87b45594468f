"""Runs every kernel family once on the bench-size volume; with `ncu --metrics gpu__time_duration.sum --csv` around it
(tools/kernel_table.py run [n]) and then `tools/kernel_table.py report launches.csv [n]` to print algorithmic GB/s."""
import sys, os, csv, collections, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run(n):
    import numpy as np
    from cl_volume_renderer_b200 import api, synth
    W, H = 1920, 1080
    ctx = api.Context(0)
    vol = api.Volume(ctx, synth.synth_ct(n))                       # k_fetch_stats
    env = api.EnvMap(ctx, synth.synth_env(2048, 1024))
    vol.set_value_clip(-2000, 3000); vol.set_gradient_clip(0, 4000)
    r = api.Renderer(ctx, W, H); r.image_set(vol, env); r.set_tf(synth.default_tf())
    r.flush_changes()                                              # k_cache_reset, k_sdf_base, k_sdf_level_warp x levels
    r.render_tf(500, 500)                                          # k_histogram, k_tf_color_frame
    pos, d = synth.default_camera(n)
    r.render_frames(pos, d, synth.glibc_rand(16), readback=False)  # k_trace, k_trace_pt, k_resolve
    r.filter_frame(2, 1.5, api.VR_FILTER2D_REFERENCE, readback=False)   # k_filter2d_reference (2d_image_filter.cl as written)
    r.filter_frame(4, 12.0, api.VR_FILTER2D_BILATERAL, readback=False)  # k_filter2d_bilateral
    r.sdf_download()                                               # k_sdf_unbrick
    r.close()
    vol.clip((8, 8, 8), (n - 8, n - 8, n - 8))                     # k_clip
    vol2 = api.Volume(ctx, synth.synth_ct(n))
    vol2.filter()                                                  # k_bilateral
    vol2.set_sampling(api.VR_SAMPLING_HW_LINEAR)                   # k_boxavg, k_fetch_stats_v8<LINEAR>
    vol2.set_value_clip(-2000, 3000); vol2.set_gradient_clip(0, 4000)
    r2 = api.Renderer(ctx, W, H); r2.set_sampling(api.VR_SAMPLING_HW_LINEAR); r2.image_set(vol2, env); r2.set_tf(synth.default_tf())
    r2.flush_changes()                                             # k_lin_field
    r2.render_tf(500, 500)                                         # k_histogram_v8<LINEAR>
    r2.render_frames(pos, d, synth.glibc_rand(16), readback=False) # LINEAR k_primary / k_trace_pt
    r2.close()
    vol2.filter()                                                  # k_bilateral on the box-averaged volume
    ctx.synchronize()
    print("done")

def report(path, n):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]; ki = h.index('Kernel Name'); vi = h.index('Metric Value'); ui = h.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi: continue
        k = r[ki].split('(')[0].replace('void ', '').split('<')[0]; v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    N = n ** 3; Nc = (n - 16) ** 3
    alg = {'k_fetch_stats_v8': 2 * N, 'k_fetch_stats': 2 * N, 'k_cache_reset': 8 * N, 'k_sdf_base': 3 * N, 'k_histogram_v8': 2 * N + 4 * 250000,
           'k_histogram': 2 * N + 4 * 250000, 'k_sdf_unbrick': 2 * N, 'k_clip': 4 * Nc, 'k_bilateral': 4 * N, 'k_tf_color_frame_ranked': 8 * 250000,
           'k_histogram_lut': 2 * N + 4 * 250000, 'k_sdf_count': N // 8 * 126, 'k_sdf_assemble8': N + N, 'k_sdf_band_bits9': 3 * N // 8,
           'k_sdf_wave9': 2 * N // 8, 'k_fetch_stats_v8i': 2 * N, 'k_boxavg': 4 * N, 'k_lin_field': 5 * N, 'k_sdf_events_v8': 2 * N + N // 8, 'k_sdf_assemble': N + N, 'k_sdf_band_bits': 3 * N // 8}
    out = {}
    for k, (cnt, us) in agg.items():
        b = alg.get(k)
        if k.startswith('k_sdf_level'): b = None
        if k.startswith('k_filter2d'): b = 8 * 1920 * 1080
        gbs = b * cnt / (us * 1e-6) / 1e9 if b else None  # alg bytes per launch x launches / total time
        out[k] = {"launches": cnt, "total_us": round(us, 1), "alg_bytes": b, "alg_GBps": round(gbs, 1) if gbs else None,
                  "frac_of_6461.5": round(gbs / 6461.5, 3) if gbs else None}
        print(f"{k:28s} n={cnt:4d} total={us:10.1f} us  alg={'%.0f MB' % (b/1e6) if b else '-':>10s}  {('%.0f GB/s' % gbs) if gbs else '':>10s} {('%.1f%%' % (100*gbs/6461.5)) if gbs else ''}")
    sdf = sum(v["total_us"] for k, v in out.items() if k.startswith("k_sdf_") and not k.startswith("k_sdf_unbrick"))
    if sdf: print(f"SDF build (base + levels): {sdf:.0f} us -> 3N/t = {3*N/(sdf*1e-6)/1e9:.0f} GB/s")
    return out

if __name__ == "__main__":
    n = int(sys.argv[3]) if len(sys.argv) > 3 else (int(sys.argv[2]) if sys.argv[1] == "run" and len(sys.argv) > 2 else 512)
    if sys.argv[1] == "run": run(n)
    else:
        o = report(sys.argv[2], n)
        json.dump(o, open(sys.argv[2].replace('.csv', '.json'), 'w'), indent=1)
