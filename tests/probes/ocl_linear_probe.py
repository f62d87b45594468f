"""What NVIDIA's OpenCL returns for read_imagei + CLK_FILTER_LINEAR on a SIGNED_INT16 3-D image (undefined by OpenCL 1.2, requested by
every sampler of the reference): dumps sampled values for a coordinate sweep so that a model can be fitted offline.
    python tests/probes/ocl_linear_probe.py > gpurun_out/ocl_linear_probe.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import ref_ocl_lib as R  # noqa: E402

if not R.available():
    print(json.dumps({"opencl": "unavailable", "why": R.error()}))
    sys.exit(0)
rs = np.random.default_rng(7)
vol = rs.integers(-2000, 3000, (6, 7, 8), dtype=np.int16)   # [nz, ny, nx]
vol[2, 3, 2:5] = (0, 1000, -1000)                             # a clean row for the 1-D sweep
coords = []
for k in range(0, 3 * 512 + 1):                               # x sweep in steps of 1/512 at a texel centre in y and z
    coords.append((1.5 + k / 512.0, 3.5, 2.5))
for k in range(0, 257):                                       # same sweep in y
    coords.append((3.5, 2.5 + k / 256.0, 2.5))
for x in range(-1, 9):                                        # integer and half-integer coordinates incl. the borders
    for off in (0.0, 0.25, 0.5, 0.75):
        coords.append((x + off, 3.0, 2.0))
        coords.append((x + off, 3.5, 2.5))
coords += [tuple(c) for c in rs.uniform(-1.0, 9.0, (600, 3))]
out = R.probe_sample(vol, coords)
print(json.dumps({"opencl": R.info(), "vol": vol.tolist(), "coords": [[float(np.float32(v)) for v in c] for c in coords], "out": out.tolist()}))
