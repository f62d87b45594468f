"""SDF build wall time on the bench volume: python tools/sdf_probe.py [n | nx,ny,nz]   (with VR_LIB=tools/ab/libvr_ab.so: VR_SDF_MODE /
VR_SDF_WAVE / VR_SDF_W6 / VR_SDF_TZ select variants); prints the device-side checksum of the field so variants can be compared"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cl_volume_renderer_b200 import api, synth
a = sys.argv[1] if len(sys.argv) > 1 else "512"
dims = tuple(int(t) for t in a.split(",")) if "," in a else (int(a),) * 3
ctx = api.Context(0)
v = synth.synth_ct(dims[0]) if len(set(dims)) == 1 else synth.synth_ct(0, dims=dims)
vol = api.Volume(ctx, v)
ts = []
for _ in range(5):
    ctx.synchronize(); t0 = time.perf_counter()
    s = api.Sdf(ctx, vol, synth.default_tf())
    ts.append(1e3 * (time.perf_counter() - t0)); lv = s.levels; ck = s.checksum(); s.close()
env = {k: os.environ[k] for k in ("VR_SDF_MODE", "VR_SDF_WAVE", "VR_SDF_VARIANT", "VR_SDF_TZ", "VR_SDF_FLOW", "VR_SDF_PDL") if k in os.environ}
print(f"dims={dims} {env} sdf_build_ms min {min(ts):.3f} median {np.median(ts):.3f} levels {lv} checksum {ck:#x}")
