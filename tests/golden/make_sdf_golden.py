"""Regenerates tests/golden/sdf_ref.npz from the reference's own SDF test fixture.

Source (read-only, only present in the build container):
  /root/reference/tests/sdf/testdata.nrrd   38x35x38 `short`, gzip   (tests/sdf/sdf_test.cpp:12-20)
  /root/reference/tests/sdf/values.x        50540 expected int8 SDF values (tests/sdf/sdf_test.cpp:28-31)
TF used by that test: `return (value > 800);` (tests/sdf/sdf_test.cpp:22).

The NRRD payload is decoded the way app/nrrd_loader.cpp:126-151,164-198 does (header ends at the
first empty line; zlib inflate with gzip auto-detect; little-endian int16, x fastest).
"""
import sys, zlib
import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
raw = open(f"{ref}/tests/sdf/testdata.nrrd", "rb").read()
end = raw.index(b"\n\n") + 2
hdr = raw[:end].decode("ascii", "replace")
sizes = [int(t) for l in hdr.splitlines() if l.startswith("sizes:") for t in l.split(":")[1].split()]
data = zlib.decompress(raw[end:], 15 + 32)
nx, ny, nz = sizes
vol = np.frombuffer(data[: nx * ny * nz * 2], dtype="<i2").reshape(nz, ny, nx)
vals = np.array([int(l.strip().rstrip(",")) for l in open(f"{ref}/tests/sdf/values.x") if l.strip()], dtype=np.int8)
assert vals.size == nx * ny * nz == 50540
np.savez_compressed(__file__.replace("make_sdf_golden.py", "sdf_ref.npz"), volume=vol, sdf=vals.reshape(nz, ny, nx),
                    threshold=np.int32(800))
print("volume", vol.shape, vol.min(), vol.max(), "sum", int(vol.astype(np.int64).sum()), "sdf range", vals.min(), vals.max())
