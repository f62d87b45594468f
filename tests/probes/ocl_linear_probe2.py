"""Second, larger probe of read_imagei + CLK_FILTER_LINEAR on a SIGNED_INT16 3-D image under NVIDIA's OpenCL (see ocl_linear_probe.py):
impulse responses (effective trilinear weights), fine 1-D / diagonal sweeps, random samples of a random volume.
    python tests/probes/ocl_linear_probe2.py gpurun_out/ocl_linear_probe2.npz"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import ref_ocl_lib as R  # noqa: E402

if not R.available():
    print("opencl unavailable:", R.error())
    sys.exit(0)
out = {}
for V in (16384, -16384, 1000):                                # impulse at texel (4,4,4) of a 9^3 volume
    imp = np.zeros((9, 9, 9), dtype=np.int16)
    imp[4, 4, 4] = V
    g = 3.5 + np.arange(65) / 32.0
    zz, yy, xx = np.meshgrid(g, g, g, indexing="ij")
    coords = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1).astype(np.float32)
    out[f"impulse_{V}_coords"] = coords
    out[f"impulse_{V}_out"] = R.probe_sample(imp, coords)[:, 0].astype(np.int32)
rs = np.random.default_rng(11)
vol = rs.integers(-3000, 3000, (16, 16, 16), dtype=np.int16)
out["vol"] = vol
t = np.arange(0, 4 * 512 + 1) / 512.0
sweeps = {"z": np.stack([np.full_like(t, 6.5), np.full_like(t, 7.5), 5.5 + t], 1),
          "diag": np.stack([5.5 + t, 6.5 + t, 4.5 + t], 1),
          "xy": np.stack([5.5 + t, 6.5 + 0.37 * t, np.full_like(t, 7.5)], 1),
          "xz": np.stack([5.5 + t, np.full_like(t, 7.5), 6.5 + 0.61 * t], 1)}
for k, cc in sweeps.items():
    out[f"sweep_{k}_coords"] = cc.astype(np.float32)
    out[f"sweep_{k}_out"] = R.probe_sample(vol, cc.astype(np.float32))
rc = rs.uniform(1.0, 15.0, (40000, 3)).astype(np.float32)
out["random_coords"] = rc
out["random_out"] = R.probe_sample(vol, rc)
np.savez_compressed(sys.argv[1], **out)
print("saved", sys.argv[1], R.info())
