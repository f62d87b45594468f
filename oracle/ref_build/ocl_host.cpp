// ocl_host.cpp — runs the reference's UNMODIFIED OpenCL kernels on a real OpenCL device (the B200 through NVIDIA's
// OpenCL driver, when the GPU box has one), with the reference's own host sequences restated around them.
//
// TEST INFRASTRUCTURE (oracle/_ref/libref_ocl.so): used by tests/test_ref_opencl_gpu.py to check the CUDA path against the
// reference itself running on the same GPU, and by tests/probes/ref_opencl_bench.py for the "reference on this GPU" timing.  Never
// linked into the product.
//
// The image has no OpenCL headers and the box has no /etc/OpenCL/vendors: the few OpenCL 1.2 types, constants and entry
// points used here are declared by hand (values from the Khronos cl.h), the ICD loader of the CUDA toolkit
// (libOpenCL.so.1) is dlopen'ed, and it is pointed at the driver's libnvidia-opencl.so.1 with OCL_ICD_FILENAMES.
//
// Kernel sources: build_ref.py embeds the include-expanded text of /root/reference/opencl_kernels/*.cl (the expansion the
// reference's own loader performs, clw_function.hpp:23-72) as byte arrays in gen_ocl_sources.inc — generated at build
// time, deleted afterwards, never committed.  Programs are built exactly like clw_function.hpp:74-111: the generated
// is_event_gen text prepended + "\n", options "-cl-mad-enable -cl-std=CL1.2".
//
// Host sequences restated (they need <CL/opencl.h> and the wrapper, so they cannot be compiled as they are):
//   reference_volume ctor   app/reference_volume.cpp:11-44        signed_distance_field ctor  app/signed_distance_field.cpp:7-35
//   renderer::flush_changes app/renderer.cpp:25-43                renderer::render_frame      app/renderer.cpp:131-158
//   image creation          opencl_wrapper/include/clw_image.hpp:18-146 (CL_R / CL_RGBA, CL_MEM_READ_WRITE, blocking push/pull)
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "gen_ocl_sources.inc"  // static const unsigned char ocl_src_<name>[] (NUL-terminated)

// ---- hand-declared OpenCL 1.2 subset -----------------------------------------------------------------------------------
typedef int32_t cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_ulong cl_bitfield;
typedef cl_uint cl_bool;
typedef struct _cl_platform_id* cl_platform_id;
typedef struct _cl_device_id* cl_device_id;
typedef struct _cl_context* cl_context;
typedef struct _cl_command_queue* cl_command_queue;
typedef struct _cl_mem* cl_mem;
typedef struct _cl_program* cl_program;
typedef struct _cl_kernel* cl_kernel;
typedef struct _cl_event* cl_event;
typedef intptr_t cl_context_properties;
struct cl_image_format { cl_uint image_channel_order, image_channel_data_type; };
struct cl_image_desc {
  cl_uint image_type;
  size_t image_width, image_height, image_depth, image_array_size, image_row_pitch, image_slice_pitch;
  cl_uint num_mip_levels, num_samples;
  cl_mem buffer;
};
enum : cl_uint {
  CL_R = 0x10B0, CL_RGBA = 0x10B5,
  CL_SIGNED_INT8 = 0x10D7, CL_SIGNED_INT16 = 0x10D8, CL_UNSIGNED_INT8 = 0x10DA,
  CL_MEM_OBJECT_IMAGE2D = 0x10F1, CL_MEM_OBJECT_IMAGE3D = 0x10F2,
  CL_PROGRAM_BUILD_LOG = 0x1183, CL_PLATFORM_NAME = 0x0902, CL_DEVICE_NAME = 0x102B, CL_DRIVER_VERSION = 0x102D
};
static const cl_bitfield CL_MEM_READ_WRITE = 1u << 0, CL_DEVICE_TYPE_ALL = 0xFFFFFFFFu, CL_QUEUE_PROFILING_ENABLE = 1u << 1;
static const cl_bool CL_TRUE = 1;

#define OCL_FUNCS(X)                                                                                                        \
  X(cl_int, clGetPlatformIDs, (cl_uint, cl_platform_id*, cl_uint*))                                                         \
  X(cl_int, clGetPlatformInfo, (cl_platform_id, cl_uint, size_t, void*, size_t*))                                           \
  X(cl_int, clGetDeviceIDs, (cl_platform_id, cl_bitfield, cl_uint, cl_device_id*, cl_uint*))                                \
  X(cl_int, clGetDeviceInfo, (cl_device_id, cl_uint, size_t, void*, size_t*))                                               \
  X(cl_context, clCreateContext,                                                                                           \
    (const cl_context_properties*, cl_uint, const cl_device_id*, void (*)(const char*, const void*, size_t, void*), void*, cl_int*)) \
  X(cl_command_queue, clCreateCommandQueue, (cl_context, cl_device_id, cl_bitfield, cl_int*))                               \
  X(cl_mem, clCreateBuffer, (cl_context, cl_bitfield, size_t, void*, cl_int*))                                              \
  X(cl_mem, clCreateImage, (cl_context, cl_bitfield, const cl_image_format*, const cl_image_desc*, void*, cl_int*))         \
  X(cl_program, clCreateProgramWithSource, (cl_context, cl_uint, const char**, const size_t*, cl_int*))                     \
  X(cl_int, clBuildProgram, (cl_program, cl_uint, const cl_device_id*, const char*, void (*)(cl_program, void*), void*))    \
  X(cl_int, clGetProgramBuildInfo, (cl_program, cl_device_id, cl_uint, size_t, void*, size_t*))                             \
  X(cl_kernel, clCreateKernel, (cl_program, const char*, cl_int*))                                                          \
  X(cl_int, clSetKernelArg, (cl_kernel, cl_uint, size_t, const void*))                                                      \
  X(cl_int, clEnqueueNDRangeKernel,                                                                                        \
    (cl_command_queue, cl_kernel, cl_uint, const size_t*, const size_t*, const size_t*, cl_uint, const cl_event*, cl_event*)) \
  X(cl_int, clEnqueueReadBuffer, (cl_command_queue, cl_mem, cl_bool, size_t, size_t, void*, cl_uint, const cl_event*, cl_event*)) \
  X(cl_int, clEnqueueWriteBuffer, (cl_command_queue, cl_mem, cl_bool, size_t, size_t, const void*, cl_uint, const cl_event*, cl_event*)) \
  X(cl_int, clEnqueueReadImage,                                                                                            \
    (cl_command_queue, cl_mem, cl_bool, const size_t*, const size_t*, size_t, size_t, void*, cl_uint, const cl_event*, cl_event*)) \
  X(cl_int, clEnqueueWriteImage,                                                                                           \
    (cl_command_queue, cl_mem, cl_bool, const size_t*, const size_t*, size_t, size_t, const void*, cl_uint, const cl_event*, cl_event*)) \
  X(cl_int, clFinish, (cl_command_queue))                                                                                   \
  X(cl_int, clReleaseMemObject, (cl_mem))                                                                                   \
  X(cl_int, clReleaseKernel, (cl_kernel))                                                                                   \
  X(cl_int, clReleaseProgram, (cl_program))                                                                                 \
  X(cl_int, clReleaseCommandQueue, (cl_command_queue))                                                                      \
  X(cl_int, clReleaseContext, (cl_context))

#define X(ret, name, args) static ret(*p_##name) args = nullptr;
OCL_FUNCS(X)
#undef X

namespace {
char g_err[4096] = "";
void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
#define OCL_CHECK(call)                                                          \
  do {                                                                           \
    cl_int e__ = (call);                                                         \
    if (e__ != 0) { set_err("%s:%d: %s -> CL error %d", __FILE__, __LINE__, #call, (int)e__); return -1; } \
  } while (0)

void* g_lib = nullptr;
cl_platform_id g_platform = nullptr;
cl_device_id g_device = nullptr;
cl_context g_ctx = nullptr;
cl_command_queue g_q = nullptr;
std::string g_info;

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
unsigned evenness(unsigned g, unsigned l) { unsigned m = g % l; return m == 0 ? g : g + l - m; }  // app/common.hpp:59-66

// Every sampler in the reference's kernels asks for CLK_FILTER_LINEAR on INTEGER images (utility_ray.cl:130,149,
// utility_filter.cl:4, reference_volume_figures.cl:12, ...), for which OpenCL 1.2 defines no result (only CLK_FILTER_NEAREST
// is defined for read_imagei / read_imageui).  g_nearest = 1 rewrites that one token to CLK_FILTER_NEAREST before the
// build — the spec-defined reading, which is what oracle.cpp and the CUDA path implement; 0 runs the text as shipped.
int g_nearest = 0;

// clw_function ctor, clw_function.hpp:74-111
int build_kernel(const unsigned char* src, const std::string& prepend, const char* fn, cl_program* prog, cl_kernel* k) {
  std::string code = prepend + "\n" + reinterpret_cast<const char*>(src);
  if (g_nearest) {
    const std::string from = "CLK_FILTER_LINEAR", to = "CLK_FILTER_NEAREST";
    for (size_t at = code.find(from); at != std::string::npos; at = code.find(from, at + to.size())) code.replace(at, from.size(), to);
  }
  const char* ind[1] = {code.c_str()};
  cl_int err = 0;
  *prog = p_clCreateProgramWithSource(g_ctx, 1, ind, nullptr, &err);
  if (err != 0) { set_err("clCreateProgramWithSource(%s) -> %d", fn, (int)err); return -1; }
  err = p_clBuildProgram(*prog, 0, nullptr, "-cl-mad-enable -cl-std=CL1.2", nullptr, nullptr);
  if (err != 0) {
    static char log[3000];
    size_t len = 0;
    memset(log, 0, sizeof(log));
    p_clGetProgramBuildInfo(*prog, g_device, CL_PROGRAM_BUILD_LOG, sizeof(log) - 1, log, &len);
    if (log[0] == '\0' && len > 1) log[0] = ' ';
    set_err("clBuildProgram(%s) -> %d\n%s", fn, (int)err, log);
    return -1;
  }
  *k = p_clCreateKernel(*prog, fn, &err);
  if (err != 0) { set_err("clCreateKernel(%s) -> %d", fn, (int)err); return -1; }
  return 0;
}

// clw_image ctor, clw_image.hpp:18-146: CL_MEM_READ_WRITE, CL_R / CL_RGBA, pitches 0
cl_mem make_image(cl_uint order, cl_uint type, size_t w, size_t h, size_t d, cl_int* err) {
  cl_image_format f{order, type};
  cl_image_desc desc;
  memset(&desc, 0, sizeof(desc));
  desc.image_type = d > 1 ? CL_MEM_OBJECT_IMAGE3D : CL_MEM_OBJECT_IMAGE2D;
  desc.image_width = w; desc.image_height = h; desc.image_depth = d; desc.image_array_size = 1;
  return p_clCreateImage(g_ctx, CL_MEM_READ_WRITE, &f, &desc, nullptr, err);
}
int write_image(cl_mem img, size_t w, size_t h, size_t d, const void* host) {  // clw_image::push, blocking
  const size_t origin[3] = {0, 0, 0}, region[3] = {w, h, d};
  OCL_CHECK(p_clEnqueueWriteImage(g_q, img, CL_TRUE, origin, region, 0, 0, host, 0, nullptr, nullptr));
  return 0;
}
int read_image(cl_mem img, size_t w, size_t h, size_t d, void* host) {  // clw_image::pull, blocking
  const size_t origin[3] = {0, 0, 0}, region[3] = {w, h, d};
  OCL_CHECK(p_clEnqueueReadImage(g_q, img, CL_TRUE, origin, region, 0, 0, host, 0, nullptr, nullptr));
  return 0;
}
int launch(cl_kernel k, unsigned dim, const size_t* global, const size_t* local) {
  OCL_CHECK(p_clEnqueueNDRangeKernel(g_q, k, dim, nullptr, global, local, 0, nullptr, nullptr));
  return 0;
}
template <class T>
int set_arg(cl_kernel k, cl_uint pos, const T& v) {
  OCL_CHECK(p_clSetKernelArg(k, pos, sizeof(T), &v));
  return 0;
}
}  // namespace

struct ocl_scene {
  int nx = 0, ny = 0, nz = 0, ew = 0, eh = 0, W = 0, H = 0;
  cl_mem vol = nullptr, sdf = nullptr, sdf_pong = nullptr, env = nullptr, frame = nullptr, cache = nullptr, counter = nullptr;
  cl_program p_render = nullptr, p_reset = nullptr, p_base = nullptr, p_sdf = nullptr;
  cl_kernel k_render = nullptr, k_reset = nullptr, k_base = nullptr, k_sdf = nullptr;
  std::string tf_src;
  int sdf_iterations = 0;
  double sdf_ms = 0.0, sdf_jit_ms = 0.0, render_jit_ms = 0.0;
};

extern "C" {

const char* ocl_last_error(void) { return g_err; }
// 1: build every kernel from now on with CLK_FILTER_LINEAR replaced by CLK_FILTER_NEAREST (see g_nearest); 0: as shipped
void ocl_set_nearest(int on) { g_nearest = on ? 1 : 0; }
const char* ocl_info(void) { return g_info.c_str(); }

// 0: an OpenCL GPU device is usable; -1 otherwise (ocl_last_error says why)
int ocl_init(void) {
  if (g_q) return 0;
  if (!getenv("OCL_ICD_FILENAMES") && !getenv("OCL_ICD_VENDORS") && access("/etc/OpenCL/vendors", R_OK) != 0) {
    for (const char* cand : {"/usr/lib/libnvidia-opencl.so.1", "/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so.1",
                             "/usr/lib64/libnvidia-opencl.so.1"})
      if (access(cand, R_OK) == 0) { setenv("OCL_ICD_FILENAMES", cand, 1); break; }
  }
  for (const char* name : {"libOpenCL.so.1", "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1", "/usr/local/cuda/lib64/libOpenCL.so.1"}) {
    g_lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (g_lib) break;
  }
  if (!g_lib) { set_err("dlopen(libOpenCL.so.1): %s", dlerror()); return -1; }
#define X(ret, name, args)                                                        \
  p_##name = reinterpret_cast<ret(*) args>(dlsym(g_lib, #name));                  \
  if (!p_##name) { set_err("dlsym(%s) failed", #name); return -1; }
  OCL_FUNCS(X)
#undef X
  cl_uint np = 0;
  cl_int e = p_clGetPlatformIDs(0, nullptr, &np);
  if (e != 0 || np == 0) { set_err("clGetPlatformIDs -> %d, %u platforms (OCL_ICD_FILENAMES=%s)", (int)e, np, getenv("OCL_ICD_FILENAMES") ? getenv("OCL_ICD_FILENAMES") : ""); return -1; }
  std::vector<cl_platform_id> plats(np);
  OCL_CHECK(p_clGetPlatformIDs(np, plats.data(), nullptr));
  g_platform = plats[0];  // clw_context.cpp:38-41: first platform, first device
  cl_uint nd = 0;
  e = p_clGetDeviceIDs(g_platform, CL_DEVICE_TYPE_ALL, 0, nullptr, &nd);
  if (e != 0 || nd == 0) { set_err("clGetDeviceIDs -> %d, %u devices", (int)e, nd); return -1; }
  std::vector<cl_device_id> devs(nd);
  OCL_CHECK(p_clGetDeviceIDs(g_platform, CL_DEVICE_TYPE_ALL, nd, devs.data(), nullptr));
  g_device = devs[0];
  cl_int err = 0;
  g_ctx = p_clCreateContext(nullptr, 1, &g_device, nullptr, nullptr, &err);
  if (err != 0) { set_err("clCreateContext -> %d", (int)err); return -1; }
  g_q = p_clCreateCommandQueue(g_ctx, g_device, CL_QUEUE_PROFILING_ENABLE, &err);
  if (err != 0) { set_err("clCreateCommandQueue -> %d", (int)err); g_q = nullptr; return -1; }
  char a[256] = "", b[256] = "", c[256] = "";
  p_clGetPlatformInfo(g_platform, CL_PLATFORM_NAME, sizeof(a) - 1, a, nullptr);
  p_clGetDeviceInfo(g_device, CL_DEVICE_NAME, sizeof(b) - 1, b, nullptr);
  p_clGetDeviceInfo(g_device, CL_DRIVER_VERSION, sizeof(c) - 1, c, nullptr);
  g_info = std::string(a) + " / " + b + " / driver " + c;
  return 0;
}

void ocl_scene_destroy(ocl_scene* s) {
  if (!s) return;
  for (cl_kernel k : {s->k_render, s->k_reset, s->k_base, s->k_sdf}) if (k) p_clReleaseKernel(k);
  for (cl_program p : {s->p_render, s->p_reset, s->p_base, s->p_sdf}) if (p) p_clReleaseProgram(p);
  for (cl_mem m : {s->vol, s->sdf, s->sdf_pong, s->env, s->frame, s->cache, s->counter}) if (m) p_clReleaseMemObject(m);
  delete s;
}

// signed_distance_field ctor, app/signed_distance_field.cpp:7-35 (kernels JIT-compiled outside the timed part; the
// reference's TIME_PRINT includes them, sdf_jit_ms reports them separately)
static int scene_build_sdf(ocl_scene* s) {
  const size_t N = (size_t)s->nx * s->ny * s->nz;
  cl_int err = 0;
  if (!s->sdf) {
    s->sdf = make_image(CL_R, CL_SIGNED_INT8, s->nx, s->ny, s->nz, &err);
    if (err != 0) { set_err("clCreateImage(sdf) -> %d", (int)err); return -1; }
    s->sdf_pong = make_image(CL_R, CL_SIGNED_INT8, s->nx, s->ny, s->nz, &err);
    if (err != 0) { set_err("clCreateImage(sdf_pong) -> %d", (int)err); return -1; }
    s->counter = p_clCreateBuffer(g_ctx, CL_MEM_READ_WRITE, sizeof(int), nullptr, &err);
    if (err != 0) { set_err("clCreateBuffer(counter) -> %d", (int)err); return -1; }
  }
  {
    std::vector<char> zeros(N, 0);  // `sdf(c, std::vector<char>(len), size, true)`: pushed on construction
    if (write_image(s->sdf, s->nx, s->ny, s->nz, zeros.data())) return -1;
  }
  double t0 = now_ms();
  if (build_kernel(ocl_src_signed_distance_field, s->tf_src, "create_base_image", &s->p_base, &s->k_base)) return -1;
  if (build_kernel(ocl_src_signed_distance_field, s->tf_src, "create_signed_distance_field", &s->p_sdf, &s->k_sdf)) return -1;
  s->sdf_jit_ms = now_ms() - t0;
  OCL_CHECK(p_clFinish(g_q));
  t0 = now_ms();
  const size_t max_iterations = std::min((size_t)std::max(s->nx, std::max(s->ny, s->nz)) / 2, (size_t)127);
  const size_t global[3] = {evenness(s->nx, 8), evenness(s->ny, 8), evenness(s->nz, 8)}, local[3] = {4, 4, 4};
  const unsigned mi = (unsigned)max_iterations;
  if (set_arg(s->k_base, 0, s->vol) || set_arg(s->k_base, 1, s->sdf) || set_arg(s->k_base, 2, s->sdf_pong) || set_arg(s->k_base, 3, mi)) return -1;
  if (launch(s->k_base, 3, global, local)) return -1;
  cl_mem ping = s->sdf, pong = s->sdf_pong;
  unsigned i;
  for (i = 1; i <= max_iterations + (max_iterations % 2) + 1; ++i) {
    int zero = 0, count = 0;
    OCL_CHECK(p_clEnqueueWriteBuffer(g_q, s->counter, CL_TRUE, 0, sizeof(int), &zero, 0, nullptr, nullptr));  // push()
    if (set_arg(s->k_sdf, 0, ping) || set_arg(s->k_sdf, 1, pong) || set_arg(s->k_sdf, 2, i) || set_arg(s->k_sdf, 3, s->counter) || set_arg(s->k_sdf, 4, mi)) return -1;
    if (launch(s->k_sdf, 3, global, local)) return -1;
    std::swap(ping, pong);
    OCL_CHECK(p_clEnqueueReadBuffer(g_q, s->counter, CL_TRUE, 0, sizeof(int), &count, 0, nullptr, nullptr));  // pull()
    if (count == 0 && i % 2 == 1) break;
  }
  OCL_CHECK(p_clFinish(g_q));
  s->sdf_ms = now_ms() - t0;
  s->sdf_iterations = (int)i;
  return 0;
}

// reference_volume ctor + env_map ctor + renderer::image_set + next_event_code_set + flush_changes
// (reference_volume.cpp:11-21, env_map.hpp:10, renderer.cpp:25-43).  tf_src = the generated is_event_gen text.
int ocl_scene_create(const int16_t* vol, int nx, int ny, int nz, const uint8_t* env_rgba, int ew, int eh, const char* tf_src, int W,
                     int H, ocl_scene** out) {
  if (ocl_init()) return -1;
  if (W % 8 || H % 8) { set_err("frame size must be a multiple of the 8x8 work-group (renderer.cpp:145)"); return -1; }
  ocl_scene* s = new ocl_scene();
  s->nx = nx; s->ny = ny; s->nz = nz; s->ew = ew; s->eh = eh; s->W = W; s->H = H; s->tf_src = tf_src;
  cl_int err = 0;
  int rc = -1;
  do {
    s->vol = make_image(CL_R, CL_SIGNED_INT16, nx, ny, nz, &err);
    if (err != 0) { set_err("clCreateImage(volume) -> %d", (int)err); break; }
    if (write_image(s->vol, nx, ny, nz, vol)) break;
    s->env = make_image(CL_RGBA, CL_UNSIGNED_INT8, ew, eh, 1, &err);
    if (err != 0) { set_err("clCreateImage(env) -> %d", (int)err); break; }
    if (write_image(s->env, ew, eh, 1, env_rgba)) break;
    s->frame = make_image(CL_RGBA, CL_UNSIGNED_INT8, W, H, 1, &err);
    if (err != 0) { set_err("clCreateImage(frame) -> %d", (int)err); break; }
    s->cache = p_clCreateBuffer(g_ctx, CL_MEM_READ_WRITE, (size_t)nx * ny * nz * 4 * sizeof(unsigned short), nullptr, &err);
    if (err != 0) { set_err("clCreateBuffer(buffer_volume) -> %d", (int)err); break; }
    if (build_kernel(ocl_src_buffer_reset, "", "buffer_reset", &s->p_reset, &s->k_reset)) break;
    const double t0 = now_ms();
    if (build_kernel(ocl_src_ray_marching, s->tf_src, "render", &s->p_render, &s->k_render)) break;
    s->render_jit_ms = now_ms() - t0;
    if (scene_build_sdf(s)) break;
    rc = 0;
  } while (0);
  if (rc) { ocl_scene_destroy(s); return -1; }
  *out = s;
  return 0;
}

// buffer_reset, renderer.cpp:32-35
int ocl_scene_reset_cache(ocl_scene* s) {
  const size_t global[3] = {evenness(s->nx, 4), evenness(s->ny, 4), evenness(s->nz, 4)}, local[3] = {4, 4, 4};
  if (set_arg(s->k_reset, 0, s->vol) || set_arg(s->k_reset, 1, s->cache)) return -1;
  if (launch(s->k_reset, 3, global, local)) return -1;
  OCL_CHECK(p_clFinish(g_q));
  return 0;
}

// renderer::render_frame for n seeds, renderer.cpp:131-158.  pull_every_frame = the reference's behaviour (blocking
// frame.pull() after every launch); 0 = only the last frame is read back.  frame_out may be null (no readback at all).
// ms_out = host wall time of the whole loop including the final clFinish.
int ocl_scene_render(ocl_scene* s, const float pos[3], const float dir[3], const int32_t* seeds, int n, int pull_every_frame,
                     uint8_t* frame_out, double* ms_out) {
  const size_t global[2] = {(size_t)s->W, (size_t)s->H}, local[2] = {8, 8};
  OCL_CHECK(p_clFinish(g_q));
  const double t0 = now_ms();
  for (int k = 0; k < n; ++k) {
    if (set_arg(s->k_render, 0, s->frame) || set_arg(s->k_render, 1, s->vol) || set_arg(s->k_render, 2, s->sdf) ||
        set_arg(s->k_render, 3, s->env) || set_arg(s->k_render, 4, s->cache))
      return -1;
    for (int a = 0; a < 3; ++a)
      if (set_arg(s->k_render, 5 + a, pos[a]) || set_arg(s->k_render, 8 + a, dir[a])) return -1;
    if (set_arg(s->k_render, 11, seeds[k])) return -1;
    if (launch(s->k_render, 2, global, local)) return -1;
    if (frame_out && (pull_every_frame || k == n - 1))
      if (read_image(s->frame, s->W, s->H, 1, frame_out)) return -1;
  }
  OCL_CHECK(p_clFinish(g_q));
  if (ms_out) *ms_out = now_ms() - t0;
  return 0;
}

int ocl_scene_cache_download(ocl_scene* s, uint16_t* out) {
  OCL_CHECK(p_clEnqueueReadBuffer(g_q, s->cache, CL_TRUE, 0, (size_t)s->nx * s->ny * s->nz * 8, out, 0, nullptr, nullptr));
  return 0;
}
int ocl_scene_sdf_download(ocl_scene* s, int8_t* out) { return read_image(s->sdf, s->nx, s->ny, s->nz, out); }
// {iterations run, -, -}, {sdf loop ms, sdf JIT ms, render JIT ms}
void ocl_scene_timings(const ocl_scene* s, int* iterations, double ms[3]) {
  *iterations = s->sdf_iterations;
  ms[0] = s->sdf_ms; ms[1] = s->sdf_jit_ms; ms[2] = s->render_jit_ms;
}

// reference_volume ctor's fetch_stats, app/reference_volume.cpp:22-41
int ocl_fetch_stats(const int16_t* vol, int nx, int ny, int nz, int32_t out[4], double* ms_out) {
  if (ocl_init()) return -1;
  cl_int err = 0;
  cl_mem img = make_image(CL_R, CL_SIGNED_INT16, nx, ny, nz, &err);
  if (err != 0) { set_err("clCreateImage(volume) -> %d", (int)err); return -1; }
  cl_mem st = p_clCreateBuffer(g_ctx, CL_MEM_READ_WRITE, 5 * sizeof(int), nullptr, &err);
  cl_program p = nullptr;
  cl_kernel k = nullptr;
  int rc = -1;
  do {
    if (err != 0) { set_err("clCreateBuffer(stats) -> %d", (int)err); break; }
    if (write_image(img, nx, ny, nz, vol)) break;
    if (build_kernel(ocl_src_reference_volume_figures, "", "fetch_stats", &p, &k)) break;
    int init[5] = {INT_MAX, INT_MIN, INT_MAX, INT_MIN, INT_MIN};
    if (p_clEnqueueWriteBuffer(g_q, st, CL_TRUE, 0, sizeof(init), init, 0, nullptr, nullptr) != 0) { set_err("write stats"); break; }
    const size_t global[3] = {evenness(nx, 8), evenness(ny, 8), evenness(nz, 8)}, local[3] = {4, 4, 4};
    if (set_arg(k, 0, img) || set_arg(k, 1, st)) break;
    p_clFinish(g_q);
    const double t0 = now_ms();
    if (launch(k, 3, global, local)) break;
    if (p_clEnqueueReadBuffer(g_q, st, CL_TRUE, 0, sizeof(init), init, 0, nullptr, nullptr) != 0) { set_err("read stats"); break; }
    if (ms_out) *ms_out = now_ms() - t0;
    memcpy(out, init, 4 * sizeof(int));
    rc = 0;
  } while (0);
  if (k) p_clReleaseKernel(k);
  if (p) p_clReleaseProgram(p);
  if (st) p_clReleaseMemObject(st);
  p_clReleaseMemObject(img);
  return rc;
}

// tf_sort_values as launched by renderer::render_tf, app/renderer.cpp:49-61.  The kernel indexes frame[x*height + y] with x up to
// `width` and y up to `height`, i.e. it writes past width*height; the device buffer is padded so those writes stay
// inside it, and only the first width*height counters are returned.  The caller must pass a range whose minima are the volume's
// own (no negative index), as render_tf does.
int ocl_histogram(const int16_t* vol, int nx, int ny, int nz, int width, int height, const float range[4], uint32_t* bins_out, double* ms_out) {
  if (ocl_init()) return -1;
  cl_int err = 0;
  const size_t nb = (size_t)width * height, padded = nb + (size_t)height + 64;
  cl_mem img = make_image(CL_R, CL_SIGNED_INT16, nx, ny, nz, &err);
  if (err != 0) { set_err("clCreateImage(volume) -> %d", (int)err); return -1; }
  cl_mem bins = p_clCreateBuffer(g_ctx, CL_MEM_READ_WRITE, padded * sizeof(uint32_t), nullptr, &err);
  cl_program p = nullptr;
  cl_kernel k = nullptr;
  int rc = -1;
  do {
    if (err != 0) { set_err("clCreateBuffer(bins) -> %d", (int)err); break; }
    if (write_image(img, nx, ny, nz, vol)) break;
    std::vector<uint32_t> zeros(padded, 0u);
    if (p_clEnqueueWriteBuffer(g_q, bins, CL_TRUE, 0, padded * sizeof(uint32_t), zeros.data(), 0, nullptr, nullptr) != 0) { set_err("write bins"); break; }
    if (build_kernel(ocl_src_histogram, "", "tf_sort_values", &p, &k)) break;
    if (set_arg(k, 0, img) || set_arg(k, 1, bins) || set_arg(k, 2, width) || set_arg(k, 3, height) || set_arg(k, 4, range[0]) ||
        set_arg(k, 5, range[1]) || set_arg(k, 6, range[2]) || set_arg(k, 7, range[3]))
      break;
    const size_t global[3] = {evenness(nx, 8), evenness(ny, 8), evenness(nz, 8)}, local[3] = {4, 4, 4};
    p_clFinish(g_q);
    const double t0 = now_ms();
    if (launch(k, 3, global, local)) break;
    if (p_clFinish(g_q) != 0) { set_err("clFinish(tf_sort_values)"); break; }
    if (ms_out) *ms_out = now_ms() - t0;
    if (p_clEnqueueReadBuffer(g_q, bins, CL_TRUE, 0, nb * sizeof(uint32_t), bins_out, 0, nullptr, nullptr) != 0) { set_err("read bins"); break; }
    rc = 0;
  } while (0);
  if (k) p_clReleaseKernel(k);
  if (p) p_clReleaseProgram(p);
  if (bins) p_clReleaseMemObject(bins);
  p_clReleaseMemObject(img);
  return rc;
}

// reference_volume::filter, app/reference_volume.cpp:70-80 (volume_filter.cl bilateral_filter)
int ocl_bilateral(const int16_t* vol, int nx, int ny, int nz, int16_t* out, double* ms_out) {
  if (ocl_init()) return -1;
  cl_int err = 0;
  cl_mem img = make_image(CL_R, CL_SIGNED_INT16, nx, ny, nz, &err);
  if (err != 0) { set_err("clCreateImage(volume) -> %d", (int)err); return -1; }
  cl_mem buf = make_image(CL_R, CL_SIGNED_INT16, nx, ny, nz, &err);
  cl_program p = nullptr;
  cl_kernel k = nullptr;
  int rc = -1;
  do {
    if (err != 0) { set_err("clCreateImage(buffer) -> %d", (int)err); break; }
    if (write_image(img, nx, ny, nz, vol)) break;
    if (build_kernel(ocl_src_volume_filter, "", "bilateral_filter", &p, &k)) break;
    if (set_arg(k, 0, img) || set_arg(k, 1, buf)) break;
    const size_t global[3] = {evenness(nx, 8), evenness(ny, 8), evenness(nz, 8)}, local[3] = {4, 4, 4};
    p_clFinish(g_q);
    const double t0 = now_ms();
    if (launch(k, 3, global, local)) break;
    if (p_clFinish(g_q) != 0) { set_err("clFinish(bilateral_filter)"); break; }
    if (ms_out) *ms_out = now_ms() - t0;
    if (read_image(buf, nx, ny, nz, out)) break;
    rc = 0;
  } while (0);
  if (k) p_clReleaseKernel(k);
  if (p) p_clReleaseProgram(p);
  if (buf) p_clReleaseMemObject(buf);
  p_clReleaseMemObject(img);
  return rc;
}

// reference_volume::set_clipping, app/reference_volume.cpp:54-68 (reference_volume_clip.cl apply_clip)
int ocl_clip(const int16_t* vol, int nx, int ny, int nz, const unsigned start[3], const unsigned size[3], int16_t* out, double* ms_out) {
  if (ocl_init()) return -1;
  cl_int err = 0;
  cl_mem img = make_image(CL_R, CL_SIGNED_INT16, nx, ny, nz, &err);
  if (err != 0) { set_err("clCreateImage(volume) -> %d", (int)err); return -1; }
  cl_mem dst = make_image(CL_R, CL_SIGNED_INT16, size[0], size[1], size[2], &err);
  cl_mem b0 = nullptr, b1 = nullptr;
  cl_program p = nullptr;
  cl_kernel k = nullptr;
  int rc = -1;
  do {
    if (err != 0) { set_err("clCreateImage(cropped) -> %d", (int)err); break; }
    const unsigned st[3] = {start[0], start[1], start[2]}, len[4] = {size[0], size[1], size[2], 4};
    b0 = p_clCreateBuffer(g_ctx, CL_MEM_READ_WRITE, sizeof(st), nullptr, &err);
    if (err == 0) b1 = p_clCreateBuffer(g_ctx, CL_MEM_READ_WRITE, sizeof(len), nullptr, &err);
    if (err != 0) { set_err("clCreateBuffer(clip) -> %d", (int)err); break; }
    if (p_clEnqueueWriteBuffer(g_q, b0, CL_TRUE, 0, sizeof(st), st, 0, nullptr, nullptr) != 0 ||
        p_clEnqueueWriteBuffer(g_q, b1, CL_TRUE, 0, sizeof(len), len, 0, nullptr, nullptr) != 0) { set_err("write clip args"); break; }
    if (write_image(img, nx, ny, nz, vol)) break;
    if (build_kernel(ocl_src_reference_volume_clip, "", "apply_clip", &p, &k)) break;
    if (set_arg(k, 0, img) || set_arg(k, 1, dst) || set_arg(k, 2, b0) || set_arg(k, 3, b1)) break;
    const size_t global[3] = {evenness(size[0], 4), evenness(size[1], 4), evenness(size[2], 4)}, local[3] = {4, 4, 4};
    p_clFinish(g_q);
    const double t0 = now_ms();
    if (launch(k, 3, global, local)) break;
    if (p_clFinish(g_q) != 0) { set_err("clFinish(apply_clip)"); break; }
    if (ms_out) *ms_out = now_ms() - t0;
    if (read_image(dst, size[0], size[1], size[2], out)) break;
    rc = 0;
  } while (0);
  if (k) p_clReleaseKernel(k);
  if (p) p_clReleaseProgram(p);
  if (b0) p_clReleaseMemObject(b0);
  if (b1) p_clReleaseMemObject(b1);
  if (dst) p_clReleaseMemObject(dst);
  p_clReleaseMemObject(img);
  return rc;
}

// What does this OpenCL implementation return for read_imagei with a CLK_FILTER_LINEAR sampler on a CL_SIGNED_INT16 image (undefined
// by the specification, requested by every sampler of the reference)?  OUR OWN probe kernel, not reference code: for n float4
// coordinates it records {linear sampler + float coords, nearest sampler + float coords, linear sampler + int coords}.
static const char* k_probe_src =
    "__kernel void probe(__read_only image3d_t v, __global const float4* c, __global int* out, int n) {\n"
    "  const sampler_t lin = CLK_FILTER_LINEAR | CLK_ADDRESS_CLAMP;\n"
    "  const sampler_t nea = CLK_FILTER_NEAREST | CLK_ADDRESS_CLAMP;\n"
    "  int i = get_global_id(0);\n"
    "  if (i >= n) return;\n"
    "  float4 p = c[i];\n"
    "  int4 pi = {(int)p.x, (int)p.y, (int)p.z, 0};\n"
    "  out[3 * i + 0] = read_imagei(v, lin, p).x;\n"
    "  out[3 * i + 1] = read_imagei(v, nea, p).x;\n"
    "  out[3 * i + 2] = read_imagei(v, lin, pi).x;\n"
    "}\n";
int ocl_probe_sample(const int16_t* vol, int nx, int ny, int nz, const float* coords4, int n, int32_t* out3) {
  if (ocl_init()) return -1;
  cl_int err = 0;
  cl_mem img = make_image(CL_R, CL_SIGNED_INT16, nx, ny, nz, &err);
  if (err != 0) { set_err("clCreateImage(probe) -> %d", (int)err); return -1; }
  cl_mem c = p_clCreateBuffer(g_ctx, CL_MEM_READ_WRITE, (size_t)n * 16, nullptr, &err);
  cl_mem o = err == 0 ? p_clCreateBuffer(g_ctx, CL_MEM_READ_WRITE, (size_t)n * 12, nullptr, &err) : nullptr;
  cl_program p = nullptr;
  cl_kernel k = nullptr;
  int rc = -1;
  do {
    if (err != 0) { set_err("clCreateBuffer(probe) -> %d", (int)err); break; }
    if (write_image(img, nx, ny, nz, vol)) break;
    if (p_clEnqueueWriteBuffer(g_q, c, CL_TRUE, 0, (size_t)n * 16, coords4, 0, nullptr, nullptr) != 0) { set_err("write coords"); break; }
    const int keep = g_nearest;
    g_nearest = 0;
    const int b = build_kernel(reinterpret_cast<const unsigned char*>(k_probe_src), "", "probe", &p, &k);
    g_nearest = keep;
    if (b) break;
    if (set_arg(k, 0, img) || set_arg(k, 1, c) || set_arg(k, 2, o) || set_arg(k, 3, n)) break;
    const size_t global[1] = {(size_t)((n + 63) / 64 * 64)}, local[1] = {64};
    if (launch(k, 1, global, local)) break;
    if (p_clEnqueueReadBuffer(g_q, o, CL_TRUE, 0, (size_t)n * 12, out3, 0, nullptr, nullptr) != 0) { set_err("read probe"); break; }
    rc = 0;
  } while (0);
  if (k) p_clReleaseKernel(k);
  if (p) p_clReleaseProgram(p);
  if (o) p_clReleaseMemObject(o);
  if (c) p_clReleaseMemObject(c);
  p_clReleaseMemObject(img);
  return rc;
}

}  // extern "C"
