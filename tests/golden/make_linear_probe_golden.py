"""Assembles tests/golden/opencl_linear_probe.npz from the two probe outputs recorded on a GPU box with the driver's OpenCL runtime:

    python tests/probes/ocl_linear_probe.py  > gpurun_out/ocl_linear_probe.json     # sweeps + out-of-range coordinates, 8x7x6 volume
    python tests/probes/ocl_linear_probe2.py gpurun_out/ocl_linear_probe2.npz       # impulses, sweeps, 40 000 random samples, 16^3 volume
    python tests/golden/make_linear_probe_golden.py gpurun_out/ocl_linear_probe.json gpurun_out/ocl_linear_probe2.npz

Only the value read with the linear sampler and float coordinates is kept (column 0 of the probe kernel's output)."""
import json
import os
import sys

import numpy as np

p1 = json.load(open(sys.argv[1]))
d = np.load(sys.argv[2])
out = {"vol": d["vol"], "info": np.array(p1["opencl"])}
for name in ("random", "sweep_z", "sweep_diag", "sweep_xy", "sweep_xz"):
    out[name + "_coords"] = d[name + "_coords"]
    out[name + "_out"] = d[name + "_out"][:, 0].astype(np.int32)
out["border_vol"] = np.array(p1["vol"], np.int16)
out["border_coords"] = np.array(p1["coords"], np.float32)
out["border_out"] = np.array(p1["out"])[:, 0].astype(np.int32)
out["impulse_coords"] = d["impulse_16384_coords"]
for V in (16384, -16384, 1000):
    out[f"impulse_{V}_out"] = d[f"impulse_{V}_out"].astype(np.int16)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "opencl_linear_probe.npz"), **out)
