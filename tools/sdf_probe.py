"""SDF build wall time on the bench volume: python tools/sdf_probe.py [n]   (VR_SDF_MODE / VR_SDF_TILE_XW select variants)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cl_volume_renderer_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = api.Context(0)
vol = api.Volume(ctx, synth.synth_ct(n))
ts = []
for _ in range(5):
    ctx.synchronize(); t0 = time.perf_counter()
    s = api.Sdf(ctx, vol, synth.default_tf())
    ts.append(1e3 * (time.perf_counter() - t0)); lv = s.levels; s.close()
print(f"n={n} mode={os.environ.get('VR_SDF_MODE','default')} xw={os.environ.get('VR_SDF_TILE_XW','4')} sdf_build_ms min {min(ts):.3f} median {np.median(ts):.3f} levels {lv}")
