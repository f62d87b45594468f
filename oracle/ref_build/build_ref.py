#!/usr/bin/env python
"""Builds oracle/_ref/libref.so: the reference's own OpenCL kernel sources compiled for the host CPU.

    python oracle/ref_build/build_ref.py /root/reference oracle/_ref

The kernels are read where they lie under <reference>/opencl_kernels; nothing is copied into the repository.  The only
text processing is what the reference's own loader does before handing the source to the OpenCL compiler —
recursive expansion of `#clw_include_once "file.cl"` with include-once semantics
(opencl_wrapper/include/clw_function.hpp:23-72) — plus ONE syntactic substitution:
    volume_filter's bilateral_kernel writes the vector literal `(float4)(x,y,z,0)` (utility_filter.cl:54); in C++ that
    parses as a cast of a comma expression, so it is rewritten to the constructor call `float4(x,y,z,0)`.
The expanded files live only for the duration of the compile (oracle/_ref/gen_*.inc, deleted afterwards).
The transfer function `is_event_gen`, which the reference generates at run time and prepends to the source
(app/ui.cpp:160-168), is provided by cl_emu.hpp as a table-driven function with the generated code's semantics.
"""
import os
import re
import subprocess
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "_ref")
here = os.path.dirname(os.path.abspath(__file__))
kdir = os.path.join(ref, "opencl_kernels")
os.makedirs(out, exist_ok=True)

KERNEL_FILES = ["ray_marching.cl", "signed_distance_field.cl", "histogram.cl", "volume_filter.cl", "buffer_reset.cl",
                "reference_volume_figures.cl", "reference_volume_clip.cl", "2d_image_filter.cl"]
INC = re.compile(r'^\s*#clw_include_once\s+"([^"]+)"\s*$')


def expand(name, seen):
    if name in seen:
        return ""
    seen.add(name)
    lines = []
    for line in open(os.path.join(kdir, name), encoding="utf-8", errors="replace").read().splitlines():
        m = INC.match(line)
        lines.append(expand(m.group(1), seen) if m else line)
    return "\n".join(lines) + "\n"


gen = []
for f in KERNEL_FILES:
    text = expand(f, set())
    if f == "volume_filter.cl":
        n = text.count("(float4)(x,y,z,0)")
        assert n == 1, f"expected exactly one `(float4)(x,y,z,0)` in bilateral_kernel, found {n}"
        text = text.replace("(float4)(x,y,z,0)", "float4(x,y,z,0)")
    p = os.path.join(out, "gen_" + f.replace(".cl", ".inc"))
    open(p, "w").write(text)
    gen.append(p)

cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-w", "-fpermissive", "-Wno-narrowing",
       "-I", here, "-I", out, os.path.join(here, "ref_driver.cpp"), "-o", os.path.join(out, "libref.so")]
try:
    subprocess.check_call(cmd)
finally:
    if not os.environ.get("KEEP_GEN"):
        for p in gen:
            os.remove(p)
print("built", os.path.join(out, "libref.so"))

# The same expanded kernel text, embedded as byte arrays, for the mini OpenCL host (ocl_host.cpp): on a GPU box whose driver
# ships an OpenCL runtime it runs the reference's unmodified kernels on the GPU itself.  /root/reference does not exist there,
# so the text travels inside the (git-ignored) binary; the generated include is deleted after the compile.
OCL_FILES = ["ray_marching.cl", "signed_distance_field.cl", "buffer_reset.cl", "reference_volume_figures.cl", "histogram.cl",
             "volume_filter.cl", "reference_volume_clip.cl"]
inc = os.path.join(out, "gen_ocl_sources.inc")
with open(inc, "w") as f:
    for name in OCL_FILES:
        data = expand(name, set()).encode("utf-8") + b"\0"
        f.write("static const unsigned char ocl_src_%s[] = {%s};\n" % (name.replace(".cl", ""), ",".join(str(b) for b in data)))
cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", out, os.path.join(here, "ocl_host.cpp"), "-ldl", "-o",
       os.path.join(out, "libref_ocl.so")]
try:
    subprocess.check_call(cmd)
finally:
    if not os.environ.get("KEEP_GEN"):
        os.remove(inc)
print("built", os.path.join(out, "libref_ocl.so"))

# The reference's own loaders (NRRD, env map through the vendored stb_image) compile as they are: a small main() around
# them gives tests/test_io_cpu.py the reference's behaviour for the ingest row.
app = os.path.join(ref, "app")
cmd = [cxx, "-std=c++17", "-O2", "-w", "-I", app, "-I", os.path.join(ref, "subprojects", "stb"),
       os.path.join(here, "ref_loaders.cpp")] + [os.path.join(app, f) for f in
                                                 ("nrrd_loader.cpp", "volume_block.cpp", "hdre_loader.cpp", "image.cpp")] + \
      ["-lz", "-o", os.path.join(out, "ref_loaders")]
subprocess.check_call(cmd)
print("built", os.path.join(out, "ref_loaders"))
