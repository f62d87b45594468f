"""Compares the CUDA texture unit's linear filtering of an int16 volume (tools/probes/tex_linear_probe) with what NVIDIA's OpenCL
returned for read_imagei + CLK_FILTER_LINEAR on the same volume and coordinates (tests/golden/opencl_linear_probe.npz, recorded by
tests/probes/ocl_linear_probe2.py)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = np.load(os.path.join(ROOT, "tests", "golden", "opencl_linear_probe.npz"))
vol = d["vol"]
sets = {"random": (d["random_coords"], d["random_out"])}
for k in ("z", "diag", "xy", "xz"):
    sets["sweep_" + k] = (d[f"sweep_{k}_coords"], d[f"sweep_{k}_out"])
tmp = tempfile.mkdtemp()
vol.tofile(os.path.join(tmp, "vol.i16"))
for name, (c, want) in sets.items():
    c = np.ascontiguousarray(c, dtype=np.float32)
    c.tofile(os.path.join(tmp, "c.f32"))
    subprocess.check_call([os.path.join(ROOT, "tools", "probes", "tex_linear_probe"), os.path.join(tmp, "vol.i16"), "16", "16", "16",
                           os.path.join(tmp, "c.f32"), str(len(c)), os.path.join(tmp, "o.f32")])
    t = np.fromfile(os.path.join(tmp, "o.f32"), dtype=np.float32).astype(np.float64)
    res = {}
    for scale in (32767.0, 32768.0):
        for rn, fn in (("rint", np.rint), ("half_up", lambda x: np.floor(x + 0.5)), ("trunc", np.trunc), ("floor", np.floor)):
            got = fn(t * scale).astype(np.int64)
            res[f"{int(scale)}/{rn}"] = (int((got != want).sum()), int(np.abs(got - want).max()))
    best = min(res.items(), key=lambda kv: kv[1])
    print(name, len(c), "best:", best, "all:", res)
