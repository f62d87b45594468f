// vr_render.cu — the per-pixel path tracer (opencl_kernels/ray_marching.cl `render` :152-199 and everything it
// calls) as two sm_100a kernels per frame:
//
//   k_trace   : phase 1 — generate_ray, box cut, SDF-driven march to the first event, token admission,
//               2 x (bounce + <=3 march segments) with environment lookups, packed atomic add into the voxel
//               cache.  Environment pixels are finished here; for shaded pixels only the hit voxel is recorded.
//   k_resolve : phase 2 — every shaded pixel reads its voxel's cache entry and tone-maps it.
//
// The split implements the two-phase frame semantics the oracle defines: in the reference the
// resolve read (ray_marching.cl:82) races with other work-items' atomic adds to the same voxel.
//
// B200-first restructuring of the march loop:
//   The reference evaluates the transfer function at every step from 7 volume texels (value + 6 gradient taps)
//   plus 1 SDF texel.  But sign(sdf[v]) < 0  <=>  is_event_gen(v) is true (create_base_image writes -1/+1 by
//   event state and no later pass changes a sign), and the SDF voxel tested for the event at the new position is
//   the very voxel the next `march` reads its step length from.  So one int8 gather per step carries both the
//   event test and the next step size; the 7 volume texels are fetched only at the <=7 hits per sample, where
//   the gradient is needed for the shading normal anyway.  Same positions, same events, 1/15 of the bytes.
#include <cstring>

#include "vr_device.cuh"

struct RenderParams {
  VolView vol;
  SdfView sdf;
  cudaSurfaceObject_t sdf_surf;  // VR_SDF_SURF=1: the field behind a surface object (k_trace_pt<.., SURF>)
  // VR_SAMPLING_HW_LINEAR (k_trace<.., LINEAR>): the volume (int16, linear filter, border 0, unnormalised coordinates) and the
  // environment map (RGBA8, linear filter, clamp to edge, normalised coordinates) behind texture objects with normalised-float reads
  cudaTextureObject_t vol_tex, env_tex;
  // VR_SAMPLING_HW_LINEAR: the step field (vr_quiet.cu) — per voxel cell the SDF byte and one "quiet" bit per octant, 16 bits behind
  // a surface object
  cudaSurfaceObject_t lin_surf;
  int lin_sched;               // k_trace_pt<.., LINEAR>: 0 two loops with leave rules, 1 weighted choice
  int lin_wf, lin_ws, lin_we;  // its parameters: weights of quiet steps / event tests / event processing
  int spc;                     // k_trace_pt: steps per scheduling decision
  int sm_k, sm_leave;          // k_trace_sm: steps per visit of a marching batch; a batch stops early below this many marching lanes
  const uchar4* __restrict__ env;
  int env_w, env_h;
  uint32_t* __restrict__ cache;
  uint32_t* __restrict__ hit;
  uchar4* __restrict__ frame;
  float fnx, fny, fnz;  // the volume's dims as floats (exit test of the step loops)
  int W, H, row0, row1;
  int blk_rows, blk_rank, blk_n;  // image-tile split: row block b = y / blk_rows is traced by rank b % blk_n (blk_n <= 1: every row)
  f3 cam_pos, cam_dir;
  f3 cam_side, cam_up;  // camera basis of generate_ray, evaluated once on the host in the same fp32 op order
  int token_cap;
  int nframes;          // frames in this launch: blockIdx.z selects the seed
  int pixel_major;      // k_trace_pt<.., REUSE> item order (see there)
  int nframes_shift;    // log2(nframes) when pixel_major == 1 and nframes is a power of two, else -1
  int rule_a, rule_b;   // k_trace_pt leaves its step loop when marching lanes * rule_a < waiting lanes * rule_b (VR_PT_RULE=a,b)
  int seeds[VR_MAX_BATCH];
  unsigned long long* counters;
  int* bbox;  // k_primary: bounding box of the shaded pixels {min x, min y, max x, max y} (may be null)
  // hybrid schedule: k_primary<.., PERFRAME> appends admitted primary hits here, k_trace_pt runs their secondary paths.
  // primary-reuse schedule: k_primary appends ONE record per shaded pixel, k_trace_pt<.., true> runs token admission and the
  // secondary paths for every (record, frame) pair.
  uint4* queue;       // 3 x uint4 per record (HitRecord)
  unsigned* qcount;   // [0] records appended, [1] records consumed
  unsigned qcap;
  TfTable tf;
};

// An admitted primary hit, everything ray_marching.cl:42-76 needs: pixel (RNG), seed, cache voxel, the clause whose colour
// is current, hit_information.origin + hit_information.direction, and the shading normal.
struct HitRecord {
  int xy, seed;
  unsigned voxel;
  int clause;
  f3 base, normal;
};
__device__ __forceinline__ void store_record(uint4* q, unsigned slot, const HitRecord& h) {
  q[3 * (size_t)slot + 0] = make_uint4((unsigned)h.xy, (unsigned)h.seed, h.voxel, (unsigned)h.clause);
  q[3 * (size_t)slot + 1] = make_uint4(__float_as_uint(h.base.x), __float_as_uint(h.base.y), __float_as_uint(h.base.z),
                                       __float_as_uint(h.normal.x));
  q[3 * (size_t)slot + 2] = make_uint4(__float_as_uint(h.normal.y), __float_as_uint(h.normal.z), 0u, 0u);
}
__device__ __forceinline__ HitRecord load_record(const uint4* q, unsigned slot) {
  const uint4 a = q[3 * (size_t)slot + 0], b = q[3 * (size_t)slot + 1], c = q[3 * (size_t)slot + 2];
  HitRecord h;
  h.xy = (int)a.x; h.seed = (int)a.y; h.voxel = a.z; h.clause = (int)a.w;
  h.base = {__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z)};
  h.normal = {__uint_as_float(b.w), __uint_as_float(c.x), __uint_as_float(c.y)};
  return h;
}

struct Ray {
  f3 o, d;
};

enum { EV_NONE = 0, EV_HIT = 1, EV_EXIT = 2 };

// generate_ray, utility_ray.cl:69-89.  cam_side / cam_up (utility_ray.cl:70-76) depend only on the camera, so the
// host evaluates them once per frame (camera_basis below) instead of once per pixel.
__device__ __forceinline__ Ray generate_ray(f3 cam_pos, f3 cam_dir, f3 cam_side, f3 cam_up, int x, int y, int x_total,
                                            int y_total) {
  const float x_f = (float)(x - x_total / 2);
  const float y_f = (float)(y - y_total / 2);
  const float aspect_ratio = (float)x_total / (float)y_total;
  const float x_offset = x_f / (float)x_total * aspect_ratio;
  const float y_offset = y_f / (float)y_total;
  f3 point = (cam_dir + x_offset * cam_side) + y_offset * cam_up;
  return {cam_pos, normalize3_shared_rcp(point)};
}

__device__ __forceinline__ bool lim(float p, int dim) { return p <= (float)dim && p >= 0.0f; }
__device__ __forceinline__ bool row_owned(const RenderParams& p, int y) {
  return p.blk_n <= 1 || (y / p.blk_rows) % p.blk_n == p.blk_rank;
}

// cut_min_eval + cut, utility_ray.cl:19-31,37-66
__device__ __forceinline__ float cut_min_eval(float a, float b) {
  if (a <= 0 || b <= 0) return 0.0f;
  return min_cl(a, b);
}
__device__ __forceinline__ bool cut_box(const VolView& v, Ray shot, f3* cut_point) {
  bool res = false;
  f3 cp = {0.0f, 0.0f, 0.0f};
  float tx = cut_min_eval(((float)v.nx - shot.o.x) / shot.d.x, (-shot.o.x) / shot.d.x);
  f3 xc = shot.o + tx * shot.d;
  float ty = cut_min_eval(((float)v.ny - shot.o.y) / shot.d.y, (-shot.o.y) / shot.d.y);
  f3 yc = shot.o + ty * shot.d;
  float tz = cut_min_eval(((float)v.nz - shot.o.z) / shot.d.z, (-shot.o.z) / shot.d.z);
  f3 zc = shot.o + tz * shot.d;
  if (lim(xc.y, v.ny) && lim(xc.z, v.nz)) { res = true; cp = xc; }
  if (lim(yc.x, v.nx) && lim(yc.z, v.nz)) { res = true; cp = yc; }
  if (lim(zc.x, v.nx) && lim(zc.y, v.ny)) { res = true; cp = zc; }
  *cut_point = cp;
  return res;
}

// What NVIDIA's OpenCL returns for read_imagei with the reference's CLK_FILTER_LINEAR | CLK_ADDRESS_CLAMP sampler on the int16
// volume (undefined by OpenCL 1.2): the texture unit interpolates the texels (centres at +0.5, 8-bit weights,
// border 0) and the result is rounded to an integer.  The same unit through a normalised-float read gives value / 32767;
// rint(t * 32767) in double equals the OpenCL value on all 48 196 probe samples (profiles/r1b_cuda_texture_vs_opencl_linear.txt).
// The filter works with 8 fraction bits, so value * 256 is an integer k, and the hardware's result is floor(k / 256 + 1/2)
// (oracle.cpp hw_round).  For |value| < 8192 one fused multiply-add into the binade of 2^23 recovers k exactly from the float
// (t * 32767 * 256 = k (1 + e), |k e| <= 2^21 * 2^-24; the single rounding of the fma is to an integer) and the rounding is an
// add and a shift — no conversion instructions, no fp64.  Larger values take the double-precision route.
__device__ __forceinline__ int tex_value(float t) {
  if (fabsf(t) < 0.25f) {
    const int k = __float_as_int(__fmaf_rn(t, 8388352.0f, 12582912.0f)) - 0x4B400000;  // 32767 * 256; 1.5 * 2^23
    return (k + 128) >> 8;
  }
  return __double2int_rn((double)t * 32767.0);
}
__device__ __forceinline__ int vol_linear(const RenderParams& p, float x, float y, float z) {
  return tex_value(tex3D<float>(p.vol_tex, x, y, z));
}
// gradient_prewitt_nn at a float position with that sampler, utility_filter.cl:2-35: taps at p +- 1 on each axis
__device__ __forceinline__ f3 gradient_linear(const RenderParams& p, f3 o) {
  const int dx = vol_linear(p, o.x + 1.0f, o.y, o.z) - vol_linear(p, o.x - 1.0f, o.y, o.z);
  const int dy = vol_linear(p, o.x, o.y + 1.0f, o.z) - vol_linear(p, o.x, o.y - 1.0f, o.z);
  const int dz = vol_linear(p, o.x, o.y, o.z + 1.0f) - vol_linear(p, o.x, o.y, o.z - 1.0f);
  return {(float)dx, (float)dy, (float)dz};
}

// ---- the step field of VR_SAMPLING_HW_LINEAR (built by vr_quiet.cu at the flush) -------------------------------------------------
// Under the interpolating reading the event test of a step no longer coincides with the sign of the per-voxel SDF, so the
// reference's loop costs 7 filtered fetches + 1 SDF byte per step.  But the filtered value at p is a convex combination
// (non-negative weights that sum to 1, result rounded to an integer) of the 2x2x2 texels of the hardware cell c = floor(p - 0.5
// in 8-bit fixed point), which is floor(p) - 1 or floor(p) per axis.  If the value interval of those eight texels meets no clause
// of the transfer function, no event is possible at p whatever the gradient is.  The field holds, per voxel cell floor(p), the
// SDF byte march() reads next (low byte) and that verdict for each of the 8 octants of the cell (high byte, bit ux + 2 uy + 4 uz,
// u = 1 when the coordinate's fraction is >= 127.5/256: the rounding of the fixed-point conversion) — ONE 2-byte gather per step
// decides whether the seven fetches can be skipped.  Conservative, hence bit-identical; the oracle counts 89-92 % of all event
// tests as skippable (orc_quiet_cells).  Outside the array the surface returns 0: step 0.5, not quiet.
__device__ __forceinline__ unsigned lin_cell(const RenderParams& p, int x, int y, int z) {
  return surf3Dread<unsigned short>(p.lin_surf, x * 2, y, z, cudaBoundaryModeZero);
}
__device__ __forceinline__ int lin_sdf(unsigned cell) { return (int)(signed char)(cell & 0xFFu); }
// o - floor(o) is exact in fp32 for o >= 0 (Sterbenz), so the comparison equals the oracle's double evaluation of the fixed-point cell
__device__ __forceinline__ bool lin_quiet(unsigned cell, float frac_x, float frac_y, float frac_z) {
  const float h = 0.498046875f;  // 127.5 / 256
  const unsigned oct = (frac_x >= h ? 1u : 0u) | (frac_y >= h ? 2u : 0u) | (frac_z >= h ? 4u : 0u);
  return ((cell >> (8u + oct)) & 1u) != 0u;
}
__device__ __forceinline__ bool lin_quiet(unsigned cell, f3 o, int vx, int vy, int vz) {
  return lin_quiet(cell, o.x - (float)vx, o.y - (float)vy, o.z - (float)vz);
}
// The step loops of k_trace_pt without the conversion pipe (ncu: XU 42 % busy with float<->int conversions of the step):
//  * floor of a coordinate: for 0 <= x < 2^23, x + 2^23 rounded DOWN is exactly 2^23 + floor(x) (ulp 1 there), so the integer is
//    a subtraction on the bit pattern and the float floor a subtraction of 2^23 — FADD.RM + IADD + FADD.  A negative x yields a
//    negative integer that is not its floor: the step loops only test its sign and feed it to the surface read (0 out of range).
//  * the SDF byte as a float: 1.5 * 2^23 + d is exact in the mantissa for |d| < 2^22.
__device__ __forceinline__ int floor_pair(float x, float* fl) {
  const float t = __fadd_rd(x, 8388608.0f);
  *fl = __fsub_rn(t, 8388608.0f);
  return __float_as_int(t) - 0x4B000000;
}
__device__ __forceinline__ float small_int_to_float(int d) { return __fsub_rn(__int_as_float(0x4B400000 + d), 12582912.0f); }

// sample_environment_map, utility_environment_map.cl:3-13: normalised coords, clamp to edge; nearest texel, or (LINEAR) the
// texture unit's bilinear interpolation of the RGBA8 texels rounded to integers
template <bool LINEAR = false>
__device__ __forceinline__ uchar4 env_sample(const RenderParams& p, f3 d) {
  float u = atan2f(d.x, d.z);
  float v = asinf(-d.y);
  u = u * 0.1591549431f;
  v = v * 0.318309886f;
  u = u + 0.5f;
  v = v + 0.5f;
  if (LINEAR) {
    // bilinear with 8 fraction bits: channel * 256 is an integer k <= 65280, recovered exactly by one fma (see tex_value); the
    // hardware's integer result is floor(k / 256 + 1/2)
    const float4 t = tex2D<float4>(p.env_tex, u, v);
    const int kx = __float_as_int(__fmaf_rn(t.x, 65280.0f, 12582912.0f)) - 0x4B400000, ky = __float_as_int(__fmaf_rn(t.y, 65280.0f, 12582912.0f)) - 0x4B400000;
    const int kz = __float_as_int(__fmaf_rn(t.z, 65280.0f, 12582912.0f)) - 0x4B400000, kw = __float_as_int(__fmaf_rn(t.w, 65280.0f, 12582912.0f)) - 0x4B400000;
    return make_uchar4((unsigned char)((kx + 128) >> 8), (unsigned char)((ky + 128) >> 8), (unsigned char)((kz + 128) >> 8),
                       (unsigned char)((kw + 128) >> 8));
  }
  int ix = f2i(floorf(u * (float)p.env_w));
  int iy = f2i(floorf(v * (float)p.env_h));
  ix = min(max(ix, 0), p.env_w - 1);
  iy = min(max(iy, 0), p.env_h - 1);
  return __ldg(p.env + (size_t)iy * p.env_w + ix);
}

// get_hemisphere_direction_reflective, utility_sampling.cl:40-50; xyprod = (get_global_id(0) + 1) * (get_global_id(1) + 1)
// the integer part, utility_sampling.cl:41-45: ra = the three hashes, comp = (ra % 2048) - 1024 with C's signed remainder
__device__ __forceinline__ void rng_triple(int seed, unsigned xyprod, int ra[3], int comp[3]) {
  const uint32_t useed = (uint32_t)seed + xyprod;
  ra[0] = (int)hash_u32(useed * 0x182205bdu);
  ra[1] = (int)hash_u32(useed * 0xe8d052f3u);
  ra[2] = (int)hash_u32(useed * 0xf1981dcfu);
#pragma unroll
  for (int k = 0; k < 3; ++k) comp[k] = (ra[k] % 2048) - 1024;
}
__device__ __forceinline__ f3 hemisphere_reflective_p(f3 normal, int seed, float roughness, unsigned xyprod) {
  int ra[3], comp[3];
  rng_triple(seed, xyprod, ra, comp);
  f3 direction = {(float)comp[0], (float)comp[1], (float)comp[2]};
  const float decider = dot3(direction, normal);
  const f3 correct = normalize3_shared_rcp(direction * decider);
  return normalize3_shared_rcp(normal * (1.0f - roughness) + correct * roughness);
}
__device__ __forceinline__ f3 hemisphere_reflective(f3 normal, int seed, float roughness, unsigned gx, unsigned gy) {
  return hemisphere_reflective_p(normal, seed, roughness, (gx + 1u) * (gy + 1u));
}

// march_to_next_event, utility_ray.cl:157-168 with march (:148-154) and get_event_and_value (:126-138) fused.
//   r        in/out: the ray, advanced to the event position
//   grad     out: gradient at the hit voxel (valid when EV_HIT) — the shading normal needs it next
//   color    in/out: written only when a TF clause with a colour matched (like `*color = tmp_color`)
template <bool COUNT, bool LINEAR = false>
__device__ __forceinline__ int march_to_next_event(const RenderParams& p, Ray& r, f3& grad, int color[4],
                                                   int& colour_clause, unsigned& steps) {
  const int nx = p.vol.nx, ny = p.vol.ny, nz = p.vol.nz;
  if (LINEAR) {
    // The reference's loop with the sampler behaviour of NVIDIA hardware: the SDF is read at integer coordinates (a plain texel
    // read), value and gradient taps at the float position are interpolated — but only where the step field cannot rule an
    // event out (see lin_cell above); a quiet step is one 2-byte gather.
    unsigned cell = lin_cell(p, f2i(r.o.x), f2i(r.o.y), f2i(r.o.z));  // march(), utility_ray.cl:148-154
    for (int i = 0; i < 70; ++i) {
      const float step_size = max_cl((float)lin_sdf(cell), 0.5f);
      r.o = r.o + step_size * r.d;
      if (COUNT) steps++;
      const int x = ifloor(r.o.x), y = ifloor(r.o.y), z = ifloor(r.o.z);
      cell = lin_cell(p, x, y, z);  // inside [0,dim] floor == trunc: also the next march's SDF read
      if (lin_quiet(cell, r.o, x, y, z)) continue;
      const bool exited = ((x | y | z) < 0) | ((float)nx < r.o.x) | ((float)ny < r.o.y) | ((float)nz < r.o.z);
      if (exited) return EV_EXIT;
      // get_event_and_value, utility_ray.cl:126-138.  The value first: when no clause can match it, the gradient (six more
      // fetches) cannot change the verdict
      const int value = vol_linear(p, r.o.x, r.o.y, r.o.z);
      if (!tf_value_may_match(p.tf, (int)(short)value)) continue;
      grad = gradient_linear(p, r.o);
      const int clause = tf_match(p.tf, (int)(short)value, f2s(length3(grad)));
      if (clause == 0) continue;
      const vr_tf_rect& q = p.tf.r[clause - 1];
      if (!(q.flags & VR_TF_THRESHOLD)) {
        color[0] = q.rgba[0]; color[1] = q.rgba[1]; color[2] = q.rgba[2]; color[3] = q.rgba[3];
        colour_clause = clause;
      }
      return EV_HIT;
    }
    return EV_NONE;
  }
  // SDF value at trunc(origin); border (any coordinate outside the field) reads 0
  int d = p.sdf.at(f2i(r.o.x), f2i(r.o.y), f2i(r.o.z));
  for (int i = 0; i < 70; ++i) {
    const float step_size = max_cl((float)d, 0.5f);
    r.o = r.o + step_size * r.d;
    if (COUNT) steps++;
    // exited_volume, utility_ray.cl:112-117 (strict).  floor() of a coordinate in (-1,0) is -1, so "any coordinate
    // < 0" is one sign test on the OR of the three floored coordinates; inside [0,dim] floor == trunc.
    const int x = ifloor(r.o.x), y = ifloor(r.o.y), z = ifloor(r.o.z);
    const bool exited = ((x | y | z) < 0) | ((float)nx < r.o.x) | ((float)ny < r.o.y) | ((float)nz < r.o.z);
    if (exited) return EV_EXIT;
    // One gather: the field has an apron at x == nx / y == ny / z == nz (a coordinate exactly on the far face) that
    // holds 0 — the border colour the reference's SDF read returns there — and real voxels are never 0.
    d = __ldg(p.sdf.f + p.sdf.addr(x, y, z));
    if (d > 0) continue;  // sign(sdf) > 0  <=>  no event at this voxel
    int clause;
    grad = gradient_voxel(p.vol, x, y, z);
    if (d < 0) {
      clause = tf_match(p.tf, p.vol.at(x, y, z), f2s(length3(grad)));
    } else {
      // far-face position: the value reads the border (0), the gradient taps are read as the reference would
      clause = tf_match(p.tf, 0, f2s(length3(grad)));
      if (clause == 0) continue;
    }
    if (clause > 0) {
      const vr_tf_rect& q = p.tf.r[clause - 1];
      if (!(q.flags & VR_TF_THRESHOLD)) {
        color[0] = q.rgba[0]; color[1] = q.rgba[1]; color[2] = q.rgba[2]; color[3] = q.rgba[3];
        colour_clause = clause;
      }
    }
    return EV_HIT;
  }
  return EV_NONE;
}

template <bool COUNT, bool LINEAR = false>
__global__ void __launch_bounds__(128, LINEAR ? 6 : 12) k_trace(const RenderParams p) {
  // one warp = an 8x4 pixel tile: neighbouring primary rays walk neighbouring voxels
  const int x = blockIdx.x * 8 + (threadIdx.x & 7);
  const int y = p.row0 + blockIdx.y * 16 + (threadIdx.x >> 3);
  unsigned c_steps = 0, c_normals = 0, c_env = 0, c_hits = 0, c_adm = 0, c_samples = 0;
  if (x < p.W && y < p.row1 && row_owned(p, y)) {
    c_samples = 1;
    const size_t pix = (size_t)y * p.W + x;
    const int seed = p.seeds[blockIdx.z];
    Ray vray = generate_ray(p.cam_pos, p.cam_dir, p.cam_side, p.cam_up, x, y, p.W, p.H);
    // in_volume / cut, ray_marching.cl:165-170
    bool is_cut;
    f3 cut_point;
    if (!(lim(vray.o.x, p.vol.nx) && lim(vray.o.y, p.vol.ny) && lim(vray.o.z, p.vol.nz)))
      is_cut = cut_box(p.vol, vray, &cut_point);
    else { is_cut = true; cut_point = vray.o; }

    int ev = EV_NONE;
    Ray cur = {cut_point, vray.d};
    f3 grad = {0.0f, 0.0f, 0.0f};
    int color[4] = {0, 0, 0, 0};
    int colour_clause = 0;
    if (is_cut) ev = march_to_next_event<COUNT, LINEAR>(p, cur, grad, color, colour_clause, c_steps);

    if (ev != EV_HIT) {
      // ray_marching.cl:172-178,188-194: environment colour, alpha 200
      uchar4 e = env_sample<LINEAR>(p, vray.d);
      e.w = 200;
      p.frame[pix] = e;
      p.hit[pix] = VR_MISS;
      c_env++;
    } else {
      // cache voxel, utility.cl:21 (trunc, clamped into the field)
      const int vx = min(max(f2i(cur.o.x), 0), p.vol.nx - 1);
      const int vy = min(max(f2i(cur.o.y), 0), p.vol.ny - 1);
      const int vz = min(max(f2i(cur.o.z), 0), p.vol.nz - 1);
      const size_t voxel = (size_t)p.vol.nx * p.vol.nz * vy + (size_t)p.vol.nx * vz + vx;
      p.hit[pix] = (uint32_t)voxel;
      c_hits++;
      uint32_t* hi = p.cache + 2 * voxel + 1;
      // atomic_allow_write_max, utility.cl:20-31
      bool admitted = false;
      {
        const int w = (int)(short)(__ldcv(hi) >> 16);
        if (!((unsigned)w > (unsigned)p.token_cap)) {
          const int t = (int)atomicAdd(hi, 0x00010000u);
          if ((unsigned)(t >> 16) < (unsigned)p.token_cap) admitted = true;
          else atomicSub(hi, 0x00010000u);
        }
      }
      if (admitted) {
        c_adm++;
        c_normals++;
        const Ray hit_information = cur;
        const f3 normal = -normalize3_shared_rcp(grad);
        float r_energy = (float)color[0] / 255.0f;
        float g_energy = (float)color[1] / 255.0f;
        float b_energy = (float)color[2] / 255.0f;
        unsigned bv0 = 0, bv1 = 0, bv2 = 0;
        for (int o = 1; o <= 2; ++o) {
          // ray_bounce_fake_reflectance, utility_ray.cl:106-109; ray_marching.cl:48-50
          cur.o = hit_information.o + hit_information.d;
          cur.d = hemisphere_reflective(normal, seed + o, (float)color[3] / 255.0f, (unsigned)x, (unsigned)y);
          cur.o = cur.o + normal * 2.0f;
          float atten = fabsf(dot3(cur.d, normal));
          for (int i = 8; i <= 10; ++i) {
            ev = march_to_next_event<COUNT, LINEAR>(p, cur, grad, color, colour_clause, c_steps);
            if (ev == EV_EXIT) {
              const float factor = 8.0f / (float)i;
              const uchar4 lm = env_sample<LINEAR>(p, cur.d);
              c_env++;
              // uint += float (ray_marching.cl:59-61): to float, add, truncate back
              bv0 = f2u((float)bv0 + atten * r_energy * (float)lm.x * factor / 1.0f);
              bv1 = f2u((float)bv1 + atten * g_energy * (float)lm.y * factor / 1.0f);
              bv2 = f2u((float)bv2 + atten * b_energy * (float)lm.z * factor / 1.0f);
              break;
            } else if (ev == EV_HIT) {
              const f3 n2 = -normalize3_shared_rcp(grad);
              c_normals++;
              cur.o = cur.o + cur.d;
              cur.d = hemisphere_reflective(n2, seed + o + i, (float)color[3] / 255.0f, (unsigned)x, (unsigned)y);
              cur.o = cur.o + n2 * 2.0f;
              atten *= fabsf(dot3(cur.d, n2));
              r_energy *= (float)color[0] / 255.0f;
              g_energy *= (float)color[1] / 255.0f;
              b_energy *= (float)color[2] / 255.0f;
            }
          }
        }
        bv0 /= 2u; bv1 /= 2u; bv2 /= 2u;  // buffer_value / dist_count
        // atomic_buffer_volume_add4, utility.cl:39-54 — results unused: RED.ADD
        const uint32_t low = (bv0 & 0xFFFFu) + ((bv1 & 0xFFFFu) << 16);
        const uint32_t high = (bv2 & 0xFFFFu);
        if (low) atomicAdd(p.cache + 2 * voxel, low);
        if (high) atomicAdd(hi, high);
      }
    }
  }
  if (COUNT) {
    unsigned v[6] = {c_steps, c_normals, c_env, c_hits, c_adm, c_samples};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      unsigned s = v[k];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) == 0 && s) atomicAdd(p.counters + k, (unsigned long long)s);
    }
  }
}

// ---- k_primary: the seed-independent part of a sample, once per pixel and call ------------------------------------------------
// Everything ray_marching.cl does before the first use of random_seed — generate_ray, cut, the primary march_to_next_event
// (:162-170, :21-33), the environment colour of a pixel whose ray leaves the volume (:172-178,188-194), the hit voxel and
// the shading normal (:42) — depends on the camera only.  A progressive batch (vr_render_frames: n seeds, one camera)
// therefore evaluates it ONCE per pixel; the n samples of the pixel differ from the token admission (:39) onwards, which
// k_trace_pt<.., true> runs per (pixel, frame).  Same values as n executions of the reference kernel, 1/n of the work.
//
// PERFRAME (hybrid schedule, vr_renderer_set_trace_mode(r, 1)): the same kernel once per pixel AND frame (blockIdx.z = frame), as
// when the camera moves between frames — nothing is shared between the samples of a pixel.  It then also runs the token
// admission of its sample (ray_marching.cl:39) and queues admitted hits only; k_trace_pt<.., false> runs their secondary paths.
template <bool COUNT, bool LINEAR = false, bool PERFRAME = false>
__global__ void __launch_bounds__(128, 12) k_primary(const RenderParams p) {
  const int x = blockIdx.x * 8 + (threadIdx.x & 7);
  const int y = p.row0 + blockIdx.y * 16 + (threadIdx.x >> 3);
  unsigned c_steps = 0, c_env = 0, c_hits = 0, c_samples = 0, c_adm = 0;
  if (x < p.W && y < p.row1 && row_owned(p, y)) {
    c_samples = 1;
    const size_t pix = (size_t)y * p.W + x;
    Ray vray = generate_ray(p.cam_pos, p.cam_dir, p.cam_side, p.cam_up, x, y, p.W, p.H);
    bool is_cut;
    f3 cut_point;
    if (!(lim(vray.o.x, p.vol.nx) && lim(vray.o.y, p.vol.ny) && lim(vray.o.z, p.vol.nz)))
      is_cut = cut_box(p.vol, vray, &cut_point);
    else { is_cut = true; cut_point = vray.o; }
    int ev = EV_NONE;
    Ray cur = {cut_point, vray.d};
    f3 grad = {0.0f, 0.0f, 0.0f};
    int color[4] = {0, 0, 0, 0};
    int colour_clause = 0;
    if (is_cut) ev = march_to_next_event<COUNT, LINEAR>(p, cur, grad, color, colour_clause, c_steps);
    if (ev != EV_HIT) {
      uchar4 e = env_sample<LINEAR>(p, vray.d);
      e.w = 200;
      p.frame[pix] = e;
      p.hit[pix] = VR_MISS;
      c_env++;
    } else {
      const int vx = min(max(f2i(cur.o.x), 0), p.vol.nx - 1);
      const int vy = min(max(f2i(cur.o.y), 0), p.vol.ny - 1);
      const int vz = min(max(f2i(cur.o.z), 0), p.vol.nz - 1);
      const size_t voxel = (size_t)p.vol.nx * p.vol.nz * vy + (size_t)p.vol.nx * vz + vx;
      p.hit[pix] = (uint32_t)voxel;
      c_hits++;
      bool queue_it = true;
      if (PERFRAME) {  // atomic_allow_write_max, utility.cl:20-31
        if (blockIdx.z == 0) {  // hits of one frame: sizes the launches of the call's other frames (launch_trace)
          const unsigned mh = __activemask();
          if ((threadIdx.x & 31) == (unsigned)(__ffs(mh) - 1)) atomicAdd(p.qcount + 2, (unsigned)__popc(mh));
        }
        uint32_t* hi = p.cache + 2 * voxel + 1;
        queue_it = false;
        const int w = (int)(short)(__ldcv(hi) >> 16);
        if (!((unsigned)w > (unsigned)p.token_cap)) {
          const int t = (int)atomicAdd(hi, 0x00010000u);
          if ((unsigned)(t >> 16) < (unsigned)p.token_cap) queue_it = true;
          else atomicSub(hi, 0x00010000u);
        }
        if (queue_it) c_adm++;
      }
      if (queue_it) {
      const unsigned m = __activemask();
      unsigned first = 0;
      const unsigned lane = threadIdx.x & 31;
      if (lane == (unsigned)(__ffs(m) - 1)) first = atomicAdd(p.qcount, (unsigned)__popc(m));
      first = __shfl_sync(m, first, __ffs(m) - 1);
      const unsigned slot = first + (unsigned)__popc(m & ((1u << lane) - 1u));  // < W*rows(*frames of the launch) <= qcap
      HitRecord h;
      h.xy = x | (y << 16); h.seed = PERFRAME ? p.seeds[blockIdx.z] : 0; h.voxel = (unsigned)voxel; h.clause = colour_clause;
      h.base = cur.o + cur.d;
      h.normal = -normalize3_shared_rcp(grad);
      store_record(p.queue, slot, h);
      if (!PERFRAME && p.bbox) {  // the incremental frame pull copies this box only (vr_api.cu read_frame)
        const int x0 = __reduce_min_sync(m, x), y0 = __reduce_min_sync(m, y), x1 = __reduce_max_sync(m, x), y1 = __reduce_max_sync(m, y);
        if (lane == (unsigned)(__ffs(m) - 1)) {
          atomicMin(p.bbox + 0, x0); atomicMin(p.bbox + 1, y0); atomicMax(p.bbox + 2, x1); atomicMax(p.bbox + 3, y1);
        }
      }
      }
    }
  }
  if (COUNT) {  // per-sample counters: the n samples of the pixel each own this primary segment (PERFRAME: one sample per thread)
    const unsigned n = PERFRAME ? 1u : (unsigned)p.nframes;
    unsigned v[6] = {c_steps * n, c_adm, c_env * n, c_hits * n, c_adm, c_samples * n};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      unsigned s = v[k];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) == 0 && s) atomicAdd(p.counters + k, (unsigned long long)s);
    }
  }
}

// ---- k_trace_pt: the secondary paths (ray_marching.cl:47-76) of the queued primary hits, on persistent warps ------------
// k_trace gives every pixel a thread for its whole life; with the secondary paths inline a warp runs until its LAST lane
// is done and ncu shows 12-17 of 32 lanes active per instruction (only the lanes whose primary ray hit do secondary work,
// 1..3 segments of 1..70 steps each, twice).  In the hybrid schedule k_primary<.., PERFRAME> stops at the admitted primary hit and
// appends a HitRecord; here a lane is a SLOT: warps pull records from the queue, all lanes with a segment in flight step
// together, and finished segments are processed in batches.  Event processing costs more than stepping (normalisations,
// the RNG bounce, the env lookup), so (a) the warp leaves the step loop only when the marching lanes are outnumbered 5 : 1, and
// (b) every lane that needs a bounce — new record, secondary hit, start of o = 2 — goes through ONE bounce site.
// What is computed per sample, every fp32 operation and its order, is unchanged; only the schedule differs.
enum { M_IDLE = 0, M_SECOND = 2 };
enum { EVP_NONE = 0, EVP_HIT = 1, EVP_EXIT = 2, EVP_SDF_NEG = 3, EVP_FARFACE = 4 };

//
// LINEAR (VR_SAMPLING_HW_LINEAR): a marching lane is in one of two states.  FAST: its last gather from the step field said "quiet"
// — it steps again at the cost of one 2-byte gather.  PENDING: it stands at a position where an event is possible and needs the
// reference's event test there (value + six gradient taps through the texture unit, the transfer function); without an event it
// steps on and the next gather decides its state.  The warp runs the two kinds in separate loops so that the lanes of an
// iteration all do the same thing: the quiet-step loop until too few lanes are FAST, then the event-test loop until too few are
// PENDING, and back — until the lanes waiting for event processing or a refill outnumber the rest (the rule above).
template <bool COUNT, bool REUSE, int CTAS, bool SURF = false, bool LINEAR = false>
__global__ void __launch_bounds__(128, CTAS) k_trace_pt(const RenderParams p, unsigned* __restrict__ work_counter) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int nx = p.vol.nx, ny = p.vol.ny, nz = p.vol.nz;
  // REUSE: the queue holds one record per shaded pixel (k_primary); work item = frame * records + record
  const unsigned records = min(__ldcv(p.qcount), p.qcap);
  const unsigned total = REUSE ? (p.pixel_major ? (records + p.pixel_major - 1) / p.pixel_major * p.pixel_major : records) * (unsigned)p.nframes
                               : records;

  // slot state
  int mode = M_IDLE;
  bool marching = false;
  bool unclassified = false;          // the segment ended inside the step loop; `ev` is decided after the loop
  bool lastq = false;                 // LINEAR: verdict of the step field at the position where the lane stopped
  bool pending = false;               // LINEAR: the event test at the current position is still to be done
  f3 hgrad = {0.0f, 0.0f, 0.0f};      // LINEAR: gradient of that test when it found a hit
  int ev = EVP_NONE;
  // Register diet (ncu: at 40 registers the spills of this kernel made 59 M local loads/stores per launch, 213 M sectors of
  // L1<->L2 traffic beside 337 M sectors of gathers): the pixel enters the RNG only as (x+1)*(y+1); base point, shading normal
  // and cache voxel of the primary hit are re-read from the record (L2-resident) the two times they are needed instead of
  // living in registers; the three radiance accumulators (<= 510 each: two adds of <= 255) share one register.
  unsigned xyp = 0, rec = 0;
  int seed = 0;
  f3 o = {0, 0, 0}, dv = {0, 0, 0};  // current ray
  int d = 0, steps_left = 0;
  float atten = 0, er = 0, eg = 0, eb = 0;
  unsigned bvp = 0;  // bv0 | bv1 << 10 | bv2 << 20
  int clause_col = 0;  // 1-based index of the clause that last wrote the colour (0: colour still {0,0,0,0})
  int po = 0, pi = 0;  // the loop variables o (1..2) and i (8..10) of ray_marching.cl:47,52
  bool exhausted = false;
  unsigned c_steps = 0, c_normals = 0, c_env = 0, c_adm = 0;

  // colour(k) / 255.0f of the clause whose colour is current ({0,0,0,0} before any clause wrote it)
  auto energy = [&](int k) -> float { return clause_col ? p.tf.e[clause_col - 1][k] : 0.0f; };

  for (;;) {
    bool need_bounce = false, reset_atten = false, need_start = false;
    f3 bn = {0, 0, 0};
    int bseed = 0;

    // ---- events of the slots whose segment ended ---------------------------------------------------------------------------
    if (mode != M_IDLE && !marching && !(LINEAR && pending)) {
      f3 grad = {0.0f, 0.0f, 0.0f};
      if (LINEAR) grad = hgrad;
      if (!LINEAR && (ev == EVP_SDF_NEG || ev == EVP_FARFACE)) {
        const int vx = ifloor(o.x), vy = ifloor(o.y), vz = ifloor(o.z);
        grad = gradient_voxel(p.vol, vx, vy, vz);
        const int value = ev == EVP_SDF_NEG ? p.vol.at(vx, vy, vz) : 0;
        const int clause = tf_match(p.tf, value, f2s(length3(grad)));
        if (ev == EVP_FARFACE && clause == 0) {
          // no event on the far face: `continue` in march_to_next_event
          ev = EVP_NONE;
          if (steps_left > 0) marching = true;
        } else {
          if (clause > 0 && !(p.tf.r[clause - 1].flags & VR_TF_THRESHOLD)) clause_col = clause;
          ev = EVP_HIT;
        }
      }
      if (!marching) {  // body of the i-loop, ray_marching.cl:53-72
        bool next_o = false;
        if (ev == EVP_EXIT) {
          const float factor = pi == 8 ? 8.0f / 8.0f : (pi == 9 ? 8.0f / 9.0f : 8.0f / 10.0f);  // 8.0f / i, i in {8,9,10}: constants
          const uchar4 lm = env_sample<LINEAR>(p, dv);
          if (COUNT) c_env++;
          const unsigned bv0 = f2u((float)(bvp & 1023u) + atten * er * (float)lm.x * factor / 1.0f);
          const unsigned bv1 = f2u((float)((bvp >> 10) & 1023u) + atten * eg * (float)lm.y * factor / 1.0f);
          const unsigned bv2 = f2u((float)(bvp >> 20) + atten * eb * (float)lm.z * factor / 1.0f);
          bvp = bv0 | (bv1 << 10) | (bv2 << 20);
          next_o = true;  // break
        } else {
          const bool more = pi < 10;
          if (ev == EVP_HIT) {
            if (COUNT) c_normals++;
            er *= energy(0); eg *= energy(1); eb *= energy(2);
            if (more) {  // at i == 10 the new ray (ray_marching.cl:64-67) is never marched and atten is reset next
              bn = -normalize3_shared_rcp(grad);
              o = o + dv;
              bseed = seed + po + pi;
              need_bounce = true;
            }
          }
          ++pi;
          if (more) need_start = true;
          else next_o = true;
        }
        if (next_o) {
          if (po == 1) {
            po = 2; pi = 8;
            const HitRecord h2 = load_record(p.queue, rec);  // hit_information.origin + direction and the primary normal again
            o = h2.base;
            bn = h2.normal; bseed = seed + po;
            need_bounce = true; reset_atten = true; need_start = true;
          } else {
            const unsigned bv0 = (bvp & 1023u) / 2u, bv1 = ((bvp >> 10) & 1023u) / 2u, bv2 = (bvp >> 20) / 2u;  // buffer_value / dist_count
            const uint32_t low = bv0 + (bv1 << 16);
            const uint32_t high = bv2;
            const size_t voxel = p.queue[3 * (size_t)rec].z;
            if (low) atomicAdd(p.cache + 2 * voxel, low);
            if (high) atomicAdd(p.cache + 2 * voxel + 1, high);
            mode = M_IDLE;
          }
        }
      }
    }

    // ---- refill free slots from the queue --------------------------------------------------------------------------------------
    for (;;) {
      const unsigned idle = __ballot_sync(0xffffffffu, mode == M_IDLE);
      if (!idle || exhausted) break;
      {
        unsigned first = 0;
        if (lane == 0) first = atomicAdd(work_counter, (unsigned)__popc(idle));
        first = __shfl_sync(0xffffffffu, first, 0);
        exhausted = first + (unsigned)__popc(idle) >= total;
        const unsigned idx = first + (unsigned)__popc(idle & lt_mask);
        bool take = mode == M_IDLE && idx < total;
        HitRecord h;
        if (take) {
          if (REUSE) {
            // item order.  Default (pixel_major = 1): consecutive items are the frames of one pixel, so the samples that start
            // from the same voxel run close together in time and find its neighbourhood in L1/L2 (2.74 -> 2.51 ms per 64-frame
            // step on the bench scene; groups of 2..32 pixels measure the same, so it is temporal, not intra-warp, locality).
            // VR_PT_ORDER=0: frame-major, consecutive items are neighbouring pixels of one frame.
            unsigned f;
            if (p.nframes_shift >= 0) {  // the default order (pixel_major 1) with a power-of-two batch: no divisions
              rec = idx >> p.nframes_shift;
              f = idx & ((1u << p.nframes_shift) - 1u);
            } else if (p.pixel_major) {  // groups of pixel_major pixels x nframes frames, the pixels of a group fastest
              const unsigned pb = (unsigned)p.pixel_major, gsz = pb * (unsigned)p.nframes;
              const unsigned grp = idx / gsz, within = idx - grp * gsz;
              f = within / pb;
              rec = grp * pb + (within - f * pb);
            } else { f = idx / records; rec = idx - f * records; }
            if (rec >= records) take = false;
            else {
            h = load_record(p.queue, rec);
            h.seed = p.seeds[f];
            // atomic_allow_write_max, utility.cl:20-31 (ray_marching.cl:39)
            uint32_t* hi = p.cache + 2 * (size_t)h.voxel + 1;
            const int w = (int)(short)(__ldcv(hi) >> 16);
            take = false;
            if (!((unsigned)w > (unsigned)p.token_cap)) {
              const int t = (int)atomicAdd(hi, 0x00010000u);
              if ((unsigned)(t >> 16) < (unsigned)p.token_cap) take = true;
              else atomicSub(hi, 0x00010000u);
            }
            if (COUNT && take) { c_adm++; c_normals++; }
            }
          } else {
            rec = idx;
            h = load_record(p.queue, idx);
          }
        }
        if (take) {  // ray_marching.cl:42-50 for o = 1
          xyp = (unsigned)((h.xy & 0xFFFF) + 1) * (unsigned)((h.xy >> 16) + 1);
          seed = h.seed; clause_col = h.clause;
          er = energy(0); eg = energy(1); eb = energy(2);
          bvp = 0;
          po = 1; pi = 8;
          o = h.base;
          bn = h.normal; bseed = seed + po;
          need_bounce = true; reset_atten = true; need_start = true;
          mode = M_SECOND;
        }
      }
      // rejected samples (voxel at the token cap) leave their slot free: draw again while at least half of the warp is idle
      if (!REUSE || __popc(__ballot_sync(0xffffffffu, mode == M_IDLE)) < 16) break;
    }

    // ---- the one bounce site: ray_bounce_fake_reflectance + `origin += normal*2` + attenuation ------------------------------------
    if (need_bounce) {
      dv = hemisphere_reflective_p(bn, bseed, energy(3), xyp);
      o = o + bn * 2.0f;
      const float a = fabsf(dot3(dv, bn));
      atten = reset_atten ? a : atten * a;
    }
    if (need_start) {  // first half of march(), utility_ray.cl:148-150
      if (LINEAR) d = lin_sdf(lin_cell(p, f2i(o.x), f2i(o.y), f2i(o.z)));
      else if (SURF) d = surf3Dread<signed char>(p.sdf_surf, f2i(o.x), f2i(o.y), f2i(o.z), cudaBoundaryModeZero);
      else d = p.sdf.at(f2i(o.x), f2i(o.y), f2i(o.z));
      steps_left = 70;
      marching = true;
      pending = false;
      ev = EVP_NONE;
    }
    if (!__ballot_sync(0xffffffffu, mode != M_IDLE)) break;  // queue empty and every slot free

    if (LINEAR) {
      // one march() + the gather that classifies the new position (utility_ray.cl:148-154, :112-117)
      auto advance = [&]() {
        const float step_size = max_cl(small_int_to_float(d), 0.5f);
        o = o + step_size * dv;
        if (COUNT) c_steps++;
        steps_left--;
        float fx, fy, fz;
        const int vx = floor_pair(o.x, &fx), vy = floor_pair(o.y, &fy), vz = floor_pair(o.z, &fz);
        const unsigned cell = lin_cell(p, vx, vy, vz);
        d = lin_sdf(cell);
        if (lin_quiet(cell, o.x - fx, o.y - fy, o.z - fz)) {  // no event possible here: get_event_and_value returns None
          pending = false;
          marching = steps_left != 0;
          if (!marching) ev = EVP_NONE;
        } else {
          const bool exited = ((vx | vy | vz) < 0) | (p.fnx < o.x) | (p.fny < o.y) | (p.fnz < o.z);
          marching = false;
          pending = !exited;
          if (exited) ev = EVP_EXIT;
        }
      };
      // get_event_and_value (utility_ray.cl:126-138) where an event is possible.  The value first: when no clause can match
      // it, the gradient (six more fetches) cannot change the verdict.
      // the same step inside the quiet-step loops: only "keeps marching or not" is decided there; classify() sorts the lanes that
      // stopped into step budget used up / left the volume / event test pending, once per loop exit
      auto advance_quiet = [&]() {
        const float step_size = max_cl(small_int_to_float(d), 0.5f);
        o = o + step_size * dv;
        if (COUNT) c_steps++;
        steps_left--;
        float fx, fy, fz;
        const int vx = floor_pair(o.x, &fx), vy = floor_pair(o.y, &fy), vz = floor_pair(o.z, &fz);
        const unsigned cell = lin_cell(p, vx, vy, vz);
        d = lin_sdf(cell);
        lastq = lin_quiet(cell, o.x - fx, o.y - fy, o.z - fz);
        marching = lastq & (steps_left != 0);
        unclassified = true;
      };
      auto classify = [&]() {
        if (unclassified && !marching) {
          if (lastq) ev = EVP_NONE;
          else {
            const bool exited = (o.x < 0.0f) | (o.y < 0.0f) | (o.z < 0.0f) | (p.fnx < o.x) | (p.fny < o.y) | (p.fnz < o.z);
            pending = !exited;
            if (exited) ev = EVP_EXIT;
          }
        }
        unclassified = false;
      };
      auto event_test = [&]() {
        const int value = vol_linear(p, o.x, o.y, o.z);
        int clause = 0;
        if (tf_value_may_match(p.tf, (int)(short)value)) {
          const f3 grad = gradient_linear(p, o);
          clause = tf_match(p.tf, (int)(short)value, f2s(length3(grad)));
          if (clause != 0) hgrad = grad;
        }
        if (clause != 0) {
          if (!(p.tf.r[clause - 1].flags & VR_TF_THRESHOLD)) clause_col = clause;
          ev = EVP_HIT;
          pending = false;
        } else if (steps_left == 0) {
          ev = EVP_NONE;
          pending = false;
        } else {
          advance();
        }
      };
      if (p.lin_sched == 0) {
        // Two loops: quiet steps until the lanes that still step are outnumbered (lin_wf : 1), then event tests until the
        // pending lanes are (lin_ws : 1); the region is left when the lanes waiting outside outnumber the rest (rule_a : rule_b).
        for (;;) {
          for (;;) {
            for (int u = 0; u < p.spc; ++u)
              if (marching) advance_quiet();
            const unsigned act = __ballot_sync(0xffffffffu, marching);
            if (!act) break;
            const unsigned others = __ballot_sync(0xffffffffu, !marching && (mode != M_IDLE || !exhausted));
            if (__popc(act) * p.lin_wf < __popc(others)) break;
          }
          classify();
          for (;;) {
            if (!__ballot_sync(0xffffffffu, pending)) break;
            if (pending) event_test();
            const unsigned pend = __ballot_sync(0xffffffffu, pending);
            if (!pend) break;
            const unsigned others = __ballot_sync(0xffffffffu, !pending && (marching || mode != M_IDLE || !exhausted));
            if (__popc(pend) * p.lin_ws < __popc(others)) break;
          }
          const unsigned go = __ballot_sync(0xffffffffu, marching || pending);
          if (!go) break;
          const unsigned waiting = __ballot_sync(0xffffffffu, !marching && !pending && (mode != M_IDLE || !exhausted));
          if (__popc(go) * p.lin_we < __popc(waiting)) break;
        }
        continue;
      }
      // Weighted choice (lin_sched 1).  Three kinds of work wait in a warp: quiet steps (cheap), event tests (1 to 7 filtered
      // fetches + the transfer function) and event processing / refills outside this region (normalisations, RNG bounce,
      // environment lookup).  Every round the warp does the kind whose lane count times its weight is largest: cheap work may
      // run with few lanes, expensive work waits until enough lanes want it.
      for (;;) {
        const unsigned mF = __ballot_sync(0xffffffffu, marching), mS = __ballot_sync(0xffffffffu, pending);
        if (!(mF | mS)) break;
        const unsigned mE = __ballot_sync(0xffffffffu, !marching && !pending && (mode != M_IDLE || !exhausted));
        const int F = __popc(mF) * p.lin_wf, S = __popc(mS) * p.lin_ws, E = __popc(mE) * p.lin_we;
        if (F >= S && F >= E) {
          for (int u = 0; u < p.spc; ++u)
            if (marching) advance_quiet();
          classify();
        } else if (S >= E) {
          if (pending) event_test();
        } else {
          break;
        }
      }
      continue;
    }

    // ---- step loop: march (second half) + get_event_and_value with the SDF-sign shortcut -------------------------------------------
    for (;;) {
      for (int u = 0; u < p.spc; ++u) {
        if (!marching) continue;
        if (SURF) {
          // conversion-free step (floor_pair): the surface returns 0 outside the field (and real voxels are never 0), so a step
          // ends exactly when the gather is <= 0 or the 70 steps are used up.  WHY it ended (exit, event voxel, far face, step
          // budget) is decided once, after the loop (`unclassified` below): segments are only ~6 steps long, so with ~20 lanes
          // marching some lane ends in almost every iteration, and a divergent tail inside the loop ran nearly every time
          // (ncu: 10 % of the kernel's instructions at 2-9 active lanes).
          const float step_size = max_cl(small_int_to_float(d), 0.5f);
          o = o + step_size * dv;
          if (COUNT) c_steps++;
          steps_left--;
          float fx, fy, fz;
          const int vx = floor_pair(o.x, &fx), vy = floor_pair(o.y, &fy), vz = floor_pair(o.z, &fz);
          d = surf3Dread<signed char>(p.sdf_surf, vx, vy, vz, cudaBoundaryModeZero);
          marching = (d > 0) & (steps_left != 0);
          unclassified = true;
        } else {
          const float step_size = max_cl((float)d, 0.5f);
          o = o + step_size * dv;
          if (COUNT) c_steps++;
          steps_left--;
          const int vx = ifloor(o.x), vy = ifloor(o.y), vz = ifloor(o.z);
          const bool exited = ((vx | vy | vz) < 0) | ((float)nx < o.x) | ((float)ny < o.y) | ((float)nz < o.z);
          if (exited) { marching = false; ev = EVP_EXIT; }
          else {
            d = __ldg(p.sdf.f + p.sdf.addr(vx, vy, vz));
            if (d <= 0) { marching = false; ev = d < 0 ? EVP_SDF_NEG : EVP_FARFACE; }
            else if (steps_left == 0) { marching = false; ev = EVP_NONE; }
          }
        }
      }
      const unsigned act = __ballot_sync(0xffffffffu, marching);
      if (!act) break;
      // leave rule: event processing (normalisations, the RNG bounce, env lookups) costs far more than a step, so the warp keeps
      // stepping until the lanes that still march are outnumbered 5 : 1 by the lanes that wait for an event or a refill
      const unsigned waiting = __ballot_sync(0xffffffffu, !marching && (mode != M_IDLE || !exhausted));
      if (__popc(act) * p.rule_a < __popc(waiting) * p.rule_b) break;
    }
    if (SURF && unclassified && !marching) {  // exited_volume (utility_ray.cl:112-117) and the SDF-sign event test, once per segment
      const bool exited = (o.x < 0.0f) | (o.y < 0.0f) | (o.z < 0.0f) | (p.fnx < o.x) | (p.fny < o.y) | (p.fnz < o.z);
      ev = d > 0 ? EVP_NONE : (exited ? EVP_EXIT : (d < 0 ? EVP_SDF_NEG : EVP_FARFACE));
      unclassified = false;
    }
  }
  if (COUNT) {
    unsigned v[5] = {c_steps, c_normals, c_env, 0u, c_adm};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      unsigned s = v[k];
      for (int q = 16; q > 0; q >>= 1) s += __shfl_xor_sync(0xffffffffu, s, q);
      if (lane == 0 && s) atomicAdd(p.counters + k, (unsigned long long)s);
    }
  }
}

#ifdef VR_AB
// ---- k_trace_pt2: two sample slots per lane (A/B build; measured and NOT adopted) ------------------------------------------------------
// ncu on k_trace_pt (profiles/r2f_nearest_trace_phase_64frames_ncu_full_summary.txt): 69 % of the stall samples are the long
// scoreboard — a step is ~25 instructions followed by a dependent gather of ~500 cycles, each lane has ONE gather in flight, and
// with 12 warps per scheduler that is 12 x 25 instructions per 500 cycles = the 0.59 issue rate measured.  Registers, not warp
// slots, cap the warps (40 registers at 48 warps per SM; 32 spill).  Here a lane carries TWO samples: every phase of the slot
// machine runs over both (statically unrolled, the state stays in registers), and the step loop issues both slots' gathers back to
// back — two independent chains per lane.  Same samples, same arithmetic; the production schedule only (batched primary reuse,
// NEAREST, surface gather).
// MEASURED (tools/slots_probe.py, bench scene, ms per 64-frame step, default / close-up view): k_trace_pt 1.93 / 5.58; two slots at
// 10 / 8 / 7 / 6 CTAs per SM (48 / 64 / 72 / 79 registers) 2.34 / 2.41 / 2.50 / 2.79 and 6.57 / 6.83 / 7.10 / 7.89 — slower, and the
// more warps the better even with spills: what hides the latency is warps in EVERY phase (event processing, refills, bounces run
// per slot, one after the other, with the divergence they have), not gathers in flight in the step loop alone.
template <int CTAS>
__global__ void __launch_bounds__(128, CTAS) k_trace_pt2(const RenderParams p, unsigned* __restrict__ work_counter) {
  constexpr int NS = 2;
  const unsigned lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned records = min(__ldcv(p.qcount), p.qcap);
  const unsigned total = (p.pixel_major ? (records + p.pixel_major - 1) / p.pixel_major * p.pixel_major : records) * (unsigned)p.nframes;

  int mode[NS], ev[NS], seed[NS], d[NS], steps_left[NS], clause_col[NS], po[NS], pi[NS];
  bool marching[NS], unclassified[NS];
  unsigned xyp[NS], rec[NS], bvp[NS];
  f3 o[NS], dv[NS];
  float atten[NS], er[NS], eg[NS], eb[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    mode[s] = M_IDLE; ev[s] = EVP_NONE; seed[s] = 0; d[s] = 0; steps_left[s] = 0; clause_col[s] = 0; po[s] = 0; pi[s] = 0;
    marching[s] = false; unclassified[s] = false; xyp[s] = 0; rec[s] = 0; bvp[s] = 0;
    o[s] = {0, 0, 0}; dv[s] = {0, 0, 0}; atten[s] = 0; er[s] = 0; eg[s] = 0; eb[s] = 0;
  }
  bool exhausted = false;

  for (;;) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      bool need_bounce = false, reset_atten = false, need_start = false;
      f3 bn = {0, 0, 0};
      int bseed = 0;
      auto energy = [&](int k) -> float { return clause_col[s] ? p.tf.e[clause_col[s] - 1][k] : 0.0f; };

      // ---- events of the slots whose segment ended (k_trace_pt has the commentary) -----------------------------------------
      if (mode[s] != M_IDLE && !marching[s]) {
        f3 grad = {0.0f, 0.0f, 0.0f};
        if (ev[s] == EVP_SDF_NEG || ev[s] == EVP_FARFACE) {
          const int vx = ifloor(o[s].x), vy = ifloor(o[s].y), vz = ifloor(o[s].z);
          grad = gradient_voxel(p.vol, vx, vy, vz);
          const int value = ev[s] == EVP_SDF_NEG ? p.vol.at(vx, vy, vz) : 0;
          const int clause = tf_match(p.tf, value, f2s(length3(grad)));
          if (ev[s] == EVP_FARFACE && clause == 0) {
            ev[s] = EVP_NONE;
            if (steps_left[s] > 0) marching[s] = true;
          } else {
            if (clause > 0 && !(p.tf.r[clause - 1].flags & VR_TF_THRESHOLD)) clause_col[s] = clause;
            ev[s] = EVP_HIT;
          }
        }
        if (!marching[s]) {  // body of the i-loop, ray_marching.cl:53-72
          bool next_o = false;
          if (ev[s] == EVP_EXIT) {
            const float factor = pi[s] == 8 ? 8.0f / 8.0f : (pi[s] == 9 ? 8.0f / 9.0f : 8.0f / 10.0f);
            const uchar4 lm = env_sample<false>(p, dv[s]);
            const unsigned bv0 = f2u((float)(bvp[s] & 1023u) + atten[s] * er[s] * (float)lm.x * factor / 1.0f);
            const unsigned bv1 = f2u((float)((bvp[s] >> 10) & 1023u) + atten[s] * eg[s] * (float)lm.y * factor / 1.0f);
            const unsigned bv2 = f2u((float)(bvp[s] >> 20) + atten[s] * eb[s] * (float)lm.z * factor / 1.0f);
            bvp[s] = bv0 | (bv1 << 10) | (bv2 << 20);
            next_o = true;
          } else {
            const bool more = pi[s] < 10;
            if (ev[s] == EVP_HIT) {
              er[s] *= energy(0); eg[s] *= energy(1); eb[s] *= energy(2);
              if (more) {
                bn = -normalize3_shared_rcp(grad);
                o[s] = o[s] + dv[s];
                bseed = seed[s] + po[s] + pi[s];
                need_bounce = true;
              }
            }
            ++pi[s];
            if (more) need_start = true;
            else next_o = true;
          }
          if (next_o) {
            if (po[s] == 1) {
              po[s] = 2; pi[s] = 8;
              const HitRecord h2 = load_record(p.queue, rec[s]);
              o[s] = h2.base;
              bn = h2.normal; bseed = seed[s] + po[s];
              need_bounce = true; reset_atten = true; need_start = true;
            } else {
              const unsigned bv0 = (bvp[s] & 1023u) / 2u, bv1 = ((bvp[s] >> 10) & 1023u) / 2u, bv2 = (bvp[s] >> 20) / 2u;
              const uint32_t low = bv0 + (bv1 << 16);
              const uint32_t high = bv2;
              const size_t voxel = p.queue[3 * (size_t)rec[s]].z;
              if (low) atomicAdd(p.cache + 2 * voxel, low);
              if (high) atomicAdd(p.cache + 2 * voxel + 1, high);
              mode[s] = M_IDLE;
            }
          }
        }
      }

      // ---- refill free slots from the queue ------------------------------------------------------------------------------
      for (;;) {
        const unsigned idle = __ballot_sync(0xffffffffu, mode[s] == M_IDLE);
        if (!idle || exhausted) break;
        {
          unsigned first = 0;
          if (lane == 0) first = atomicAdd(work_counter, (unsigned)__popc(idle));
          first = __shfl_sync(0xffffffffu, first, 0);
          exhausted = first + (unsigned)__popc(idle) >= total;
          const unsigned idx = first + (unsigned)__popc(idle & lt_mask);
          bool take = mode[s] == M_IDLE && idx < total;
          HitRecord h;
          if (take) {
            unsigned f;
            if (p.nframes_shift >= 0) {
              rec[s] = idx >> p.nframes_shift;
              f = idx & ((1u << p.nframes_shift) - 1u);
            } else if (p.pixel_major) {
              const unsigned pb = (unsigned)p.pixel_major, gsz = pb * (unsigned)p.nframes;
              const unsigned grp = idx / gsz, within = idx - grp * gsz;
              f = within / pb;
              rec[s] = grp * pb + (within - f * pb);
            } else { f = idx / records; rec[s] = idx - f * records; }
            if (rec[s] >= records) take = false;
            else {
              h = load_record(p.queue, rec[s]);
              h.seed = p.seeds[f];
              uint32_t* hi = p.cache + 2 * (size_t)h.voxel + 1;
              const int w = (int)(short)(__ldcv(hi) >> 16);
              take = false;
              if (!((unsigned)w > (unsigned)p.token_cap)) {
                const int t = (int)atomicAdd(hi, 0x00010000u);
                if ((unsigned)(t >> 16) < (unsigned)p.token_cap) take = true;
                else atomicSub(hi, 0x00010000u);
              }
            }
          }
          if (take) {
            xyp[s] = (unsigned)((h.xy & 0xFFFF) + 1) * (unsigned)((h.xy >> 16) + 1);
            seed[s] = h.seed; clause_col[s] = h.clause;
            er[s] = energy(0); eg[s] = energy(1); eb[s] = energy(2);
            bvp[s] = 0;
            po[s] = 1; pi[s] = 8;
            o[s] = h.base;
            bn = h.normal; bseed = seed[s] + po[s];
            need_bounce = true; reset_atten = true; need_start = true;
            mode[s] = M_SECOND;
          }
        }
        if (__popc(__ballot_sync(0xffffffffu, mode[s] == M_IDLE)) < 16) break;
      }

      // ---- the one bounce site + first half of march() -------------------------------------------------------------------------
      if (need_bounce) {
        dv[s] = hemisphere_reflective_p(bn, bseed, energy(3), xyp[s]);
        o[s] = o[s] + bn * 2.0f;
        const float a = fabsf(dot3(dv[s], bn));
        atten[s] = reset_atten ? a : atten[s] * a;
      }
      if (need_start) {
        d[s] = surf3Dread<signed char>(p.sdf_surf, f2i(o[s].x), f2i(o[s].y), f2i(o[s].z), cudaBoundaryModeZero);
        steps_left[s] = 70;
        marching[s] = true;
        ev[s] = EVP_NONE;
      }
    }
    if (!__ballot_sync(0xffffffffu, mode[0] != M_IDLE || mode[1] != M_IDLE)) break;  // queue empty and every slot free

    // ---- step loop: both slots' gathers are issued back to back ------------------------------------------------------------------
    for (;;) {
      for (int u = 0; u < p.spc; ++u) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          if (!marching[s]) continue;
          const float step_size = max_cl(small_int_to_float(d[s]), 0.5f);
          o[s] = o[s] + step_size * dv[s];
          steps_left[s]--;
          float fx, fy, fz;
          const int vx = floor_pair(o[s].x, &fx), vy = floor_pair(o[s].y, &fy), vz = floor_pair(o[s].z, &fz);
          d[s] = surf3Dread<signed char>(p.sdf_surf, vx, vy, vz, cudaBoundaryModeZero);
          marching[s] = (d[s] > 0) & (steps_left[s] != 0);
          unclassified[s] = true;
        }
      }
      const int act = __popc(__ballot_sync(0xffffffffu, marching[0])) + __popc(__ballot_sync(0xffffffffu, marching[1]));
      if (!act) break;
      const int waiting = __popc(__ballot_sync(0xffffffffu, !marching[0] && (mode[0] != M_IDLE || !exhausted))) +
                          __popc(__ballot_sync(0xffffffffu, !marching[1] && (mode[1] != M_IDLE || !exhausted)));
      if (act * p.rule_a < waiting * p.rule_b) break;
    }
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (unclassified[s] && !marching[s]) {  // exited_volume and the SDF-sign event test, once per segment
        const bool exited = (o[s].x < 0.0f) | (o[s].y < 0.0f) | (o[s].z < 0.0f) | (p.fnx < o[s].x) | (p.fny < o[s].y) | (p.fnz < o[s].z);
        ev[s] = d[s] > 0 ? EVP_NONE : (exited ? EVP_EXIT : (d[s] < 0 ? EVP_SDF_NEG : EVP_FARFACE));
        unclassified[s] = false;
      }
  }
}
template <int CTAS>
static int launch_pt2(vr_ctx* ctx, const RenderParams& p, unsigned* work_counter) {
  static int per_sm = 0;
  if (!per_sm) VR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_pt2<CTAS>, 128, 0));
  const unsigned blocks = (unsigned)ctx->sm_count * (unsigned)std::max(1, per_sm);
  k_trace_pt2<CTAS><<<blocks, 128, 0, ctx->stream>>>(p, work_counter);
  ctx->launches++;
  return VR_OK;
}
#endif

#ifdef VR_AB
// ---- k_trace_sm: the same secondary paths with the slots in SHARED memory and every kind of work packed -------------------------
// k_trace_pt keeps a sample in one lane for its whole life, so a lane whose segment ended is dead weight in the step loop until
// the warp leaves it, and event processing runs with whatever lanes happen to wait (ncu: 15-20 of 32 lanes active in the step
// loop, the kernel issue-bound).  Here a CTA owns SM_SLOTS sample slots in shared memory and three lists of slot numbers —
// slots whose segment ended or that are free (E), slots that march (M), and, under hw-linear sampling, slots that stand at a
// position where the event test must be evaluated (P).  A round processes every list of the current generation in batches of 32
// CONSECUTIVE list entries per warp — a batch is all events, all marching or all tests, so its lanes do the same thing — and
// appends each slot to the list of the next generation it now belongs to: marching lanes that end their segment go to E, the
// others back to M after at most sm_k steps (a batch stops early when fewer than sm_leave of its lanes still march), events
// that started a segment go to M, free slots that found no admitted sample return to E.  Two __syncthreads per round; what
// a sample computes is unchanged.
// MEASURED AND NOT ADOPTED (profiles/r2e_trace_sm_probe.jsonl): parity-green in every test, but 4.1-4.5 ms per 64-frame step
// against k_trace_pt's 2.02 (hw-linear 5.4-5.7 against 3.3).  The premise was wrong: the SDF makes segments so short (about 6
// steps) that marching is ~10 % of k_trace_pt's instructions; the rest is event processing, which k_trace_pt already runs in
// batches of ~24 lanes, and here pays a round trip of the slot state through shared memory plus two barriers per round.
// Compiled into the A/B build only (-DVR_AB, vr_renderer_set_trace_mode(r, 3)).
#define SM_SLOTS 512
#define SM_THREADS 256
enum { SMW_OX = 0, SMW_OY, SMW_OZ, SMW_DX, SMW_DY, SMW_DZ, SMW_PK, SMW_ATT, SMW_ER, SMW_EG, SMW_EB, SMW_BVP, SMW_REC, SMW_SEED, SMW_XYP,
       SMW_WORDS_NEAREST, SMW_GX = SMW_WORDS_NEAREST, SMW_GY, SMW_GZ, SMW_WORDS_LINEAR };
enum { SML_E = 0, SML_M = 1, SML_P = 2 };
static size_t sm_trace_smem_bytes(bool linear) {
  return (size_t)(linear ? SMW_WORDS_LINEAR : SMW_WORDS_NEAREST) * SM_SLOTS * 4 + (size_t)2 * 3 * SM_SLOTS * 2;
}

template <bool COUNT, bool REUSE, bool LINEAR>
__global__ void __launch_bounds__(SM_THREADS, LINEAR ? 4 : 6) k_trace_sm(const RenderParams p, unsigned* __restrict__ work_counter) {
  extern __shared__ uint32_t sm_dyn[];
  constexpr int NW = LINEAR ? SMW_WORDS_LINEAR : SMW_WORDS_NEAREST;
  uint32_t* st = sm_dyn;                                                          // [NW][SM_SLOTS]
  unsigned short* lists = reinterpret_cast<unsigned short*>(sm_dyn + NW * SM_SLOTS);  // [generation][kind][SM_SLOTS]
  __shared__ unsigned cnt[2][3];
  __shared__ unsigned cursor;
  const unsigned lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int nx = p.vol.nx, ny = p.vol.ny, nz = p.vol.nz;
  const unsigned records = min(__ldcv(p.qcount), p.qcap);
  const unsigned total = REUSE ? (p.pixel_major ? (records + p.pixel_major - 1) / p.pixel_major * p.pixel_major : records) * (unsigned)p.nframes
                               : records;
  // every slot starts free: the first round's E list is all of them
  for (unsigned i = threadIdx.x; i < SM_SLOTS; i += SM_THREADS) {
    lists[(0 * 3 + SML_E) * SM_SLOTS + i] = (unsigned short)i;
    st[SMW_PK * SM_SLOTS + i] = 0u;
  }
  if (threadIdx.x == 0) {
    cnt[0][SML_E] = SM_SLOTS; cnt[0][SML_M] = 0; cnt[0][SML_P] = 0;
    cnt[1][SML_E] = 0; cnt[1][SML_M] = 0; cnt[1][SML_P] = 0;
    cursor = 0;
  }
  unsigned c_steps = 0, c_normals = 0, c_env = 0, c_adm = 0;
  bool exhausted = false;  // per warp: this warp has seen the end of the work queue
  int cur = 0;

  // append the slots of the lanes with `pred` to list `kind` of generation `gen` (all lanes of the warp call this together)
  auto push = [&](int gen, int kind, bool pred, unsigned id) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    if (!m) return;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(&cnt[gen][kind], (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pred) lists[(gen * 3 + kind) * SM_SLOTS + base + (unsigned)__popc(m & lt_mask)] = (unsigned short)id;
  };
  // pk: d (int8) | steps_left << 8 | po << 16 | pi << 18 | ev << 22 | clause_col << 25 | busy << 30
  auto pack = [](int d, int steps_left, int po, int pi, int ev, int clause_col, int busy) -> uint32_t {
    return ((unsigned)d & 0xFFu) | ((unsigned)steps_left << 8) | ((unsigned)po << 16) | ((unsigned)pi << 18) | ((unsigned)ev << 22) |
           ((unsigned)clause_col << 25) | ((unsigned)busy << 30);
  };

  for (;;) {
    __syncthreads();  // the lists of generation `cur` are complete, `cursor` is 0
    const unsigned nE = cnt[cur][SML_E], nM = cnt[cur][SML_M], nP = LINEAR ? cnt[cur][SML_P] : 0u;
    if (nE + nM + nP == 0) break;
    const unsigned bM = (nM + 31) >> 5, bP = (nP + 31) >> 5, bE = (nE + 31) >> 5;
    const int nxt = cur ^ 1;
    for (;;) {
      unsigned b = 0;
      if (lane == 0) b = atomicAdd(&cursor, 1u);
      b = __shfl_sync(0xffffffffu, b, 0);
      if (b >= bM + bP + bE) break;
      if (b < bM) {
        // ---- a batch of marching slots: up to sm_k steps ----------------------------------------------------------------------
        const unsigned e = b * 32 + lane;
        const bool valid = e < nM;
        const unsigned id = valid ? lists[(cur * 3 + SML_M) * SM_SLOTS + e] : 0u;
        f3 o = {0, 0, 0}, dv = {0, 0, 0};
        uint32_t pk = 0;
        if (valid) {
          o = {__uint_as_float(st[SMW_OX * SM_SLOTS + id]), __uint_as_float(st[SMW_OY * SM_SLOTS + id]), __uint_as_float(st[SMW_OZ * SM_SLOTS + id])};
          dv = {__uint_as_float(st[SMW_DX * SM_SLOTS + id]), __uint_as_float(st[SMW_DY * SM_SLOTS + id]), __uint_as_float(st[SMW_DZ * SM_SLOTS + id])};
          pk = st[SMW_PK * SM_SLOTS + id];
        }
        int d = (int)(signed char)(pk & 0xFFu), steps_left = (int)((pk >> 8) & 0xFFu);
        int ev = EVP_NONE;
        bool marching = valid, pending = false;
        for (int k = 0; k < p.sm_k; ++k) {
          if (marching) {
            const float step_size = max_cl(small_int_to_float(d), 0.5f);
            o = o + step_size * dv;
            if (COUNT) c_steps++;
            steps_left--;
            float fx, fy, fz;
            const int vx = floor_pair(o.x, &fx), vy = floor_pair(o.y, &fy), vz = floor_pair(o.z, &fz);
            if (LINEAR) {
              const unsigned cell = lin_cell(p, vx, vy, vz);
              d = lin_sdf(cell);
              if (lin_quiet(cell, o.x - fx, o.y - fy, o.z - fz)) {
                if (steps_left == 0) { marching = false; ev = EVP_NONE; }
              } else {
                const bool exited = ((vx | vy | vz) < 0) | ((float)nx < o.x) | ((float)ny < o.y) | ((float)nz < o.z);
                marching = false;
                pending = !exited;
                if (exited) ev = EVP_EXIT;
              }
            } else {
              d = surf3Dread<signed char>(p.sdf_surf, vx, vy, vz, cudaBoundaryModeZero);
              if (d <= 0) {
                const bool exited = ((vx | vy | vz) < 0) | ((float)nx < o.x) | ((float)ny < o.y) | ((float)nz < o.z);
                marching = false;
                ev = exited ? EVP_EXIT : (d < 0 ? EVP_SDF_NEG : EVP_FARFACE);
              } else if (steps_left == 0) { marching = false; ev = EVP_NONE; }
            }
          }
          if (__popc(__ballot_sync(0xffffffffu, marching)) < p.sm_leave) break;
        }
        if (valid) {
          st[SMW_OX * SM_SLOTS + id] = __float_as_uint(o.x); st[SMW_OY * SM_SLOTS + id] = __float_as_uint(o.y); st[SMW_OZ * SM_SLOTS + id] = __float_as_uint(o.z);
          st[SMW_PK * SM_SLOTS + id] = (pk & 0x7E3F0000u) | ((unsigned)d & 0xFFu) | ((unsigned)steps_left << 8) | ((unsigned)ev << 22);
        }
        push(nxt, SML_M, valid && marching, id);
        if (LINEAR) push(nxt, SML_P, valid && pending, id);
        push(nxt, SML_E, valid && !marching && !pending, id);
      } else if (LINEAR && b < bM + bP) {
        // ---- a batch of slots that need get_event_and_value at their position (utility_ray.cl:126-138) --------------------------
        const unsigned e = (b - bM) * 32 + lane;
        const bool valid = e < nP;
        const unsigned id = valid ? lists[(cur * 3 + SML_P) * SM_SLOTS + e] : 0u;
        bool marching = false, pending = false;
        if (valid) {
          f3 o = {__uint_as_float(st[SMW_OX * SM_SLOTS + id]), __uint_as_float(st[SMW_OY * SM_SLOTS + id]), __uint_as_float(st[SMW_OZ * SM_SLOTS + id])};
          uint32_t pk = st[SMW_PK * SM_SLOTS + id];
          int d = (int)(signed char)(pk & 0xFFu), steps_left = (int)((pk >> 8) & 0xFFu);
          int clause_col = (int)((pk >> 25) & 31u);
          int ev = EVP_NONE;
          const int value = vol_linear(p, o.x, o.y, o.z);
          int clause = 0;
          if (tf_value_may_match(p.tf, (int)(short)value)) {
            const f3 grad = gradient_linear(p, o);
            clause = tf_match(p.tf, (int)(short)value, f2s(length3(grad)));
            if (clause != 0) {
              st[SMW_GX * SM_SLOTS + id] = __float_as_uint(grad.x); st[SMW_GY * SM_SLOTS + id] = __float_as_uint(grad.y);
              st[SMW_GZ * SM_SLOTS + id] = __float_as_uint(grad.z);
            }
          }
          if (clause != 0) {
            if (!(p.tf.r[clause - 1].flags & VR_TF_THRESHOLD)) clause_col = clause;
            ev = EVP_HIT;
          } else if (steps_left == 0) {
            ev = EVP_NONE;
          } else {  // no event here: the next march() and the gather that classifies the new position
            const f3 dv = {__uint_as_float(st[SMW_DX * SM_SLOTS + id]), __uint_as_float(st[SMW_DY * SM_SLOTS + id]), __uint_as_float(st[SMW_DZ * SM_SLOTS + id])};
            const float step_size = max_cl(small_int_to_float(d), 0.5f);
            o = o + step_size * dv;
            if (COUNT) c_steps++;
            steps_left--;
            float fx, fy, fz;
            const int vx = floor_pair(o.x, &fx), vy = floor_pair(o.y, &fy), vz = floor_pair(o.z, &fz);
            const unsigned cell = lin_cell(p, vx, vy, vz);
            d = lin_sdf(cell);
            if (lin_quiet(cell, o.x - fx, o.y - fy, o.z - fz)) {
              marching = steps_left != 0;
              if (!marching) ev = EVP_NONE;
            } else {
              const bool exited = ((vx | vy | vz) < 0) | ((float)nx < o.x) | ((float)ny < o.y) | ((float)nz < o.z);
              pending = !exited;
              if (exited) ev = EVP_EXIT;
            }
            st[SMW_OX * SM_SLOTS + id] = __float_as_uint(o.x); st[SMW_OY * SM_SLOTS + id] = __float_as_uint(o.y); st[SMW_OZ * SM_SLOTS + id] = __float_as_uint(o.z);
          }
          st[SMW_PK * SM_SLOTS + id] = (pk & 0x403F0000u) | ((unsigned)d & 0xFFu) | ((unsigned)steps_left << 8) | ((unsigned)ev << 22) | ((unsigned)clause_col << 25);
        }
        push(nxt, SML_M, valid && marching, id);
        push(nxt, SML_P, valid && pending, id);
        push(nxt, SML_E, valid && !marching && !pending, id);
      } else {
        // ---- a batch of slots whose segment ended, or that are free: ray_marching.cl:53-76, refill, bounce, start --------------
        const unsigned e = (b - bM - bP) * 32 + lane;
        const bool valid = e < nE;
        const unsigned id = valid ? lists[(cur * 3 + SML_E) * SM_SLOTS + e] : 0u;
        uint32_t pk = valid ? st[SMW_PK * SM_SLOTS + id] : 0u;
        int d = (int)(signed char)(pk & 0xFFu), steps_left = (int)((pk >> 8) & 0xFFu);
        int po = (int)((pk >> 16) & 3u), pi = (int)((pk >> 18) & 15u), ev = (int)((pk >> 22) & 7u), clause_col = (int)((pk >> 25) & 31u);
        bool busy = ((pk >> 30) & 1u) != 0u;
        f3 o = {0, 0, 0}, dv = {0, 0, 0};
        float atten = 0, er = 0, eg = 0, eb = 0;
        unsigned bvp = 0, rec = 0, xyp = 0;
        int seed = 0;
        if (busy) {
          o = {__uint_as_float(st[SMW_OX * SM_SLOTS + id]), __uint_as_float(st[SMW_OY * SM_SLOTS + id]), __uint_as_float(st[SMW_OZ * SM_SLOTS + id])};
          dv = {__uint_as_float(st[SMW_DX * SM_SLOTS + id]), __uint_as_float(st[SMW_DY * SM_SLOTS + id]), __uint_as_float(st[SMW_DZ * SM_SLOTS + id])};
          atten = __uint_as_float(st[SMW_ATT * SM_SLOTS + id]);
          er = __uint_as_float(st[SMW_ER * SM_SLOTS + id]); eg = __uint_as_float(st[SMW_EG * SM_SLOTS + id]); eb = __uint_as_float(st[SMW_EB * SM_SLOTS + id]);
          bvp = st[SMW_BVP * SM_SLOTS + id]; rec = st[SMW_REC * SM_SLOTS + id]; seed = (int)st[SMW_SEED * SM_SLOTS + id]; xyp = st[SMW_XYP * SM_SLOTS + id];
        }
        auto energy = [&](int k) -> float { return clause_col ? p.tf.e[clause_col - 1][k] : 0.0f; };
        bool need_bounce = false, reset_atten = false, need_start = false, resume = false;
        f3 bn = {0, 0, 0};
        int bseed = 0;
        if (busy) {
          f3 grad = {0.0f, 0.0f, 0.0f};
          if (LINEAR) {
            if (ev == EVP_HIT)
              grad = {__uint_as_float(st[SMW_GX * SM_SLOTS + id]), __uint_as_float(st[SMW_GY * SM_SLOTS + id]), __uint_as_float(st[SMW_GZ * SM_SLOTS + id])};
          } else if (ev == EVP_SDF_NEG || ev == EVP_FARFACE) {
            const int vx = ifloor(o.x), vy = ifloor(o.y), vz = ifloor(o.z);
            grad = gradient_voxel(p.vol, vx, vy, vz);
            const int value = ev == EVP_SDF_NEG ? p.vol.at(vx, vy, vz) : 0;
            const int clause = tf_match(p.tf, value, f2s(length3(grad)));
            if (ev == EVP_FARFACE && clause == 0) {
              ev = EVP_NONE;                     // no event on the far face: `continue` in march_to_next_event
              if (steps_left > 0) resume = true;
            } else {
              if (clause > 0 && !(p.tf.r[clause - 1].flags & VR_TF_THRESHOLD)) clause_col = clause;
              ev = EVP_HIT;
            }
          }
          if (!resume) {  // body of the i-loop, ray_marching.cl:53-72
            bool next_o = false;
            if (ev == EVP_EXIT) {
              const float factor = pi == 8 ? 8.0f / 8.0f : (pi == 9 ? 8.0f / 9.0f : 8.0f / 10.0f);
              const uchar4 lm = env_sample<LINEAR>(p, dv);
              if (COUNT) c_env++;
              const unsigned bv0 = f2u((float)(bvp & 1023u) + atten * er * (float)lm.x * factor / 1.0f);
              const unsigned bv1 = f2u((float)((bvp >> 10) & 1023u) + atten * eg * (float)lm.y * factor / 1.0f);
              const unsigned bv2 = f2u((float)(bvp >> 20) + atten * eb * (float)lm.z * factor / 1.0f);
              bvp = bv0 | (bv1 << 10) | (bv2 << 20);
              next_o = true;
            } else {
              const bool more = pi < 10;
              if (ev == EVP_HIT) {
                if (COUNT) c_normals++;
                er *= energy(0); eg *= energy(1); eb *= energy(2);
                if (more) {
                  bn = -normalize3_shared_rcp(grad);
                  o = o + dv;
                  bseed = seed + po + pi;
                  need_bounce = true;
                }
              }
              ++pi;
              if (more) need_start = true;
              else next_o = true;
            }
            if (next_o) {
              if (po == 1) {
                po = 2; pi = 8;
                const HitRecord h2 = load_record(p.queue, rec);
                o = h2.base;
                bn = h2.normal; bseed = seed + po;
                need_bounce = true; reset_atten = true; need_start = true;
              } else {
                const unsigned bv0 = (bvp & 1023u) / 2u, bv1 = ((bvp >> 10) & 1023u) / 2u, bv2 = (bvp >> 20) / 2u;
                const uint32_t low = bv0 + (bv1 << 16);
                const uint32_t high = bv2;
                const size_t voxel = p.queue[3 * (size_t)rec].z;
                if (low) atomicAdd(p.cache + 2 * voxel, low);
                if (high) atomicAdd(p.cache + 2 * voxel + 1, high);
                busy = false;
              }
            }
          }
        }
        // refill: free slots draw the next work items
        bool retry = false;
        {
          const unsigned idle = __ballot_sync(0xffffffffu, valid && !busy);
          if (idle && !exhausted) {
            unsigned first = 0;
            if (lane == 0) first = atomicAdd(work_counter, (unsigned)__popc(idle));
            first = __shfl_sync(0xffffffffu, first, 0);
            exhausted = first + (unsigned)__popc(idle) >= total;
            const unsigned idx = first + (unsigned)__popc(idle & lt_mask);
            bool take = valid && !busy && idx < total;
            retry = take;  // a drawn item that is rejected (voxel at its token cap, padding item) frees the slot for the next draw
            HitRecord h;
            if (take) {
              if (REUSE) {
                unsigned f;
                if (p.pixel_major) {
                  const unsigned pb = (unsigned)p.pixel_major, gsz = pb * (unsigned)p.nframes;
                  const unsigned grp = idx / gsz, within = idx - grp * gsz;
                  f = within / pb;
                  rec = grp * pb + (within - f * pb);
                } else { f = idx / records; rec = idx - f * records; }
                if (rec >= records) take = false;
                else {
                  h = load_record(p.queue, rec);
                  h.seed = p.seeds[f];
                  uint32_t* hi = p.cache + 2 * (size_t)h.voxel + 1;  // atomic_allow_write_max, utility.cl:20-31
                  const int w = (int)(short)(__ldcv(hi) >> 16);
                  take = false;
                  if (!((unsigned)w > (unsigned)p.token_cap)) {
                    const int t = (int)atomicAdd(hi, 0x00010000u);
                    if ((unsigned)(t >> 16) < (unsigned)p.token_cap) take = true;
                    else atomicSub(hi, 0x00010000u);
                  }
                  if (COUNT && take) { c_adm++; c_normals++; }
                }
              } else {
                rec = idx;
                h = load_record(p.queue, idx);
              }
            }
            if (take) {  // ray_marching.cl:42-50 for o = 1
              xyp = (unsigned)((h.xy & 0xFFFF) + 1) * (unsigned)((h.xy >> 16) + 1);
              seed = h.seed; clause_col = h.clause;
              er = energy(0); eg = energy(1); eb = energy(2);
              bvp = 0;
              po = 1; pi = 8;
              o = h.base;
              bn = h.normal; bseed = seed + po;
              need_bounce = true; reset_atten = true; need_start = true;
              busy = true;
              retry = false;
            }
          }
        }
        if (need_bounce) {
          dv = hemisphere_reflective_p(bn, bseed, energy(3), xyp);
          o = o + bn * 2.0f;
          const float a = fabsf(dot3(dv, bn));
          atten = reset_atten ? a : atten * a;
        }
        if (need_start) {  // first half of march(), utility_ray.cl:148-150
          if (LINEAR) d = lin_sdf(lin_cell(p, f2i(o.x), f2i(o.y), f2i(o.z)));
          else d = surf3Dread<signed char>(p.sdf_surf, f2i(o.x), f2i(o.y), f2i(o.z), cudaBoundaryModeZero);
          steps_left = 70;
          ev = EVP_NONE;
        }
        if (valid) {
          if (busy) {
            st[SMW_OX * SM_SLOTS + id] = __float_as_uint(o.x); st[SMW_OY * SM_SLOTS + id] = __float_as_uint(o.y); st[SMW_OZ * SM_SLOTS + id] = __float_as_uint(o.z);
            st[SMW_DX * SM_SLOTS + id] = __float_as_uint(dv.x); st[SMW_DY * SM_SLOTS + id] = __float_as_uint(dv.y); st[SMW_DZ * SM_SLOTS + id] = __float_as_uint(dv.z);
            st[SMW_ATT * SM_SLOTS + id] = __float_as_uint(atten);
            st[SMW_ER * SM_SLOTS + id] = __float_as_uint(er); st[SMW_EG * SM_SLOTS + id] = __float_as_uint(eg); st[SMW_EB * SM_SLOTS + id] = __float_as_uint(eb);
            st[SMW_BVP * SM_SLOTS + id] = bvp; st[SMW_REC * SM_SLOTS + id] = rec; st[SMW_SEED * SM_SLOTS + id] = (uint32_t)seed; st[SMW_XYP * SM_SLOTS + id] = xyp;
          }
          st[SMW_PK * SM_SLOTS + id] = pack(d, steps_left, po, pi, ev, clause_col, busy ? 1 : 0);
        }
        push(nxt, SML_M, valid && busy, id);           // a started (or resumed) segment marches next round
        push(nxt, SML_E, valid && !busy && retry, id);  // rejected draw: try the next item; exhausted queue: the slot retires
      }
    }
    __syncthreads();  // every warp is done with generation `cur`
    if (threadIdx.x == 0) {
      cnt[cur][SML_E] = 0; cnt[cur][SML_M] = 0; cnt[cur][SML_P] = 0;
      cursor = 0;
    }
    cur = nxt;
  }
  if (COUNT) {
    unsigned v[5] = {c_steps, c_normals, c_env, 0u, c_adm};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      unsigned sum = v[k];
      for (int q = 16; q > 0; q >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, q);
      if (lane == 0 && sum) atomicAdd(p.counters + k, (unsigned long long)sum);
    }
  }
}

#endif  // VR_AB

// phase 2: ray_marching.cl:82-99
__device__ __forceinline__ uchar4 resolve_rgbw(uint32_t r, uint32_t g, uint32_t b, uint32_t w);
__device__ __forceinline__ uchar4 resolve_entry(const uint2 c) {
  return resolve_rgbw(c.x & 0xFFFFu, c.x >> 16, c.y & 0xFFFFu, c.y >> 16);
}
__device__ __forceinline__ uchar4 resolve_rgbw(uint32_t r, uint32_t g, uint32_t b, uint32_t w) {
  if (w != 0) { r /= w; g /= w; b /= w; } else { r = g = b = 0; }
  const float inv_gamma = 1.0f / 1.77777777f;
  const float brightness = 4.0f;
  float fr = (float)r / 255.0f, fg = (float)g / 255.0f, fb = (float)b / 255.0f;
  fr = powf(fr * brightness, inv_gamma) * 255.0f;
  fg = powf(fg * brightness, inv_gamma) * 255.0f;
  fb = powf(fb * brightness, inv_gamma) * 255.0f;
  return make_uchar4((unsigned char)min(f2u(fr), 255u), (unsigned char)min(f2u(fg), 255u), (unsigned char)min(f2u(fb), 255u), 1);
}
__global__ void __launch_bounds__(256) k_resolve(const uint32_t* __restrict__ hit, const uint2* __restrict__ cache,
                                                 uchar4* __restrict__ frame, int W, int row0, int row1) {
  const size_t n = (size_t)W * (row1 - row0);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t pix = (size_t)row0 * W + i;
  const uint32_t voxel = hit[pix];
  if (voxel == VR_MISS) return;
  frame[pix] = resolve_entry(cache[voxel]);
}

// Host evaluation of utility_ray.cl:70-76 in fp32, same operation order as the device helpers (dot3/length3/normalize3).
// volatile keeps every intermediate in fp32 and forbids contraction, so the result is bit-identical to the per-pixel
// evaluation the reference does.
namespace {
struct h3 { float x, y, z; };
inline h3 h_cross(h3 a, h3 b) {
  volatile float x0 = a.y * b.z, x1 = a.z * b.y, y0 = a.z * b.x, y1 = a.x * b.z, z0 = a.x * b.y, z1 = a.y * b.x;
  volatile float x = x0 - x1, y = y0 - y1, z = z0 - z1;
  return {x, y, z};
}
inline h3 h_normalize(h3 a) {
  volatile float xx = a.x * a.x, yy = a.y * a.y, zz = a.z * a.z;
  volatile float s0 = xx + yy;
  volatile float s1 = s0 + zz;
  volatile float l = sqrtf(s1);
  if (l == 0.0f) return {0.0f, 0.0f, 0.0f};
  volatile float x = a.x / l, y = a.y / l, z = a.z / l;
  return {x, y, z};
}
void camera_basis(const float dir[3], f3* side, f3* up) {
  const h3 upv = {0.0f, 1.0f, 0.0f};
  const h3 d = {dir[0], dir[1], dir[2]};
  h3 s = h_normalize(h_cross(upv, d));
  h3 u = h_normalize(h_cross(d, s));
  if (u.y < 0) { u.x = -u.x; u.y = -u.y; u.z = -u.z; }
  *side = {s.x, s.y, s.z};
  *up = {u.x, u.y, u.z};
}
}  // namespace

// launch of one k_trace_pt instantiation on persistent CTAs: as many as are resident at once
template <bool COUNT, bool REUSE, int CTAS, bool SURF, bool LINEAR>
static int launch_pt(vr_ctx* ctx, const RenderParams& p, unsigned* work_counter) {
  static int per_sm = 0;  // every device this library accepts is a B200: one answer per instantiation
  if (!per_sm) VR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_pt<COUNT, REUSE, CTAS, SURF, LINEAR>, 128, 0));
  const unsigned blocks = (unsigned)ctx->sm_count * (unsigned)std::max(1, per_sm);
  k_trace_pt<COUNT, REUSE, CTAS, SURF, LINEAR><<<blocks, 128, 0, ctx->stream>>>(p, work_counter);
  ctx->launches++;
  return VR_OK;
}
// register budget of the production variants: 12 CTAs/SM = 40 registers for both samplings (the hw-linear variant spills a little
// at 40 and still wins: the kernel lives on warps in flight).  A -DVR_AB build adds the other budgets (vr_renderer_set_tuning).
template <bool COUNT, bool REUSE, bool LINEAR>
static int launch_pt_select(vr_renderer* r, const RenderParams& p, unsigned* wc) {
  vr_ctx* ctx = r->ctx;
  if (LINEAR) {
#ifdef VR_AB
    if (!COUNT && r->tune.pt_ctas == 6) return launch_pt<COUNT, REUSE, 6, true, true>(ctx, p, wc);
    if (!COUNT && r->tune.pt_ctas == 8) return launch_pt<COUNT, REUSE, 8, true, true>(ctx, p, wc);
    if (!COUNT && r->tune.pt_ctas == 10) return launch_pt<COUNT, REUSE, 10, true, true>(ctx, p, wc);
#endif
    return launch_pt<COUNT, REUSE, 12, true, true>(ctx, p, wc);  // measured 8 / 10 / 12 CTAs per SM: 3.15 / 3.10 / 2.91 ms per step
  }
  const bool surf = REUSE && !COUNT && r->sdf->surf != 0 && r->tune.surf;  // the surface-object gather serves the production schedule
#ifdef VR_AB
  if (surf && r->tune.pt_slots == 2) {
    switch (r->tune.pt2_ctas) {
      case 6: return launch_pt2<6>(ctx, p, wc);
      case 7: return launch_pt2<7>(ctx, p, wc);
      case 10: return launch_pt2<10>(ctx, p, wc);
      default: return launch_pt2<8>(ctx, p, wc);
    }
  }
#endif
#ifdef VR_AB
  if (REUSE && !COUNT) {
    const int c = r->tune.pt_ctas;
    if (surf && c == 8) return launch_pt<false, true, 8, true, false>(ctx, p, wc);
    if (surf && c == 10) return launch_pt<false, true, 10, true, false>(ctx, p, wc);
    if (surf && c == 14) return launch_pt<false, true, 14, true, false>(ctx, p, wc);
    if (surf && c == 16) return launch_pt<false, true, 16, true, false>(ctx, p, wc);
    if (!surf && c == 8) return launch_pt<false, true, 8, false, false>(ctx, p, wc);
    if (!surf && c == 10) return launch_pt<false, true, 10, false, false>(ctx, p, wc);
  }
#endif
  if (surf) return launch_pt<COUNT, REUSE, 12, true, false>(ctx, p, wc);
  return launch_pt<COUNT, REUSE, 12, false, false>(ctx, p, wc);
}

template <bool COUNT, bool LINEAR>
static int launch_trace(vr_renderer* r, RenderParams& p, const float pos[3], const float dir[3], int rows, int nframes, bool first_of_call) {
  vr_ctx* ctx = r->ctx;
  const size_t pixels = (size_t)r->W * rows * nframes;
  dim3 grid(div_up(r->W, 8), div_up(rows, 16), nframes);
  if (r->trace_mode == 0) {  // one thread per pixel and frame for its whole life
    r->primary_valid = false;
    k_trace<COUNT, LINEAR><<<grid, 128, 0, ctx->stream>>>(p);
    ctx->launches++;
    return VR_OK;
  }
  // mode 1 (hybrid): k_primary per pixel and frame (primary march + token admission) queues the admitted hits, persistent warps
  //                  run their secondary paths
  // mode 2 (primary reuse, default): k_primary once per pixel, persistent warps run admission + secondary paths per (pixel, frame)
  const bool reuse = r->trace_mode >= 2;
  const size_t cap = reuse ? (size_t)r->W * rows : std::max<size_t>((size_t)r->W * rows, std::min<size_t>(pixels, (size_t)32 << 20));
  if (r->queue_cap < cap) {
    if (r->queue) VR_CUDA(cudaFreeAsync(r->queue, ctx->stream));
    r->queue = nullptr; r->queue_cap = 0;
    r->primary_valid = false;
    VR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&r->queue), cap * 3 * sizeof(uint4), ctx->stream));
    r->queue_cap = cap;
  }
  p.queue = r->queue;
  p.qcap = (unsigned)r->queue_cap;
  p.qcount = reinterpret_cast<unsigned*>(r->counters + 6);
  if (!reuse) {
    // The queue never overflows: the primary hits of a call's frames are the same pixels (one camera), so the first frame runs
    // alone, its hit count h is read back (one 4-byte transfer per call), and the other frames go h-sized per launch pair.
    r->primary_valid = false;
    RenderParams q = p;
    unsigned* hits = p.qcount + 2;  // word [2] of the renderer's spare counter pair
    int fps = 1;
    for (int f0 = 0, nb = 0; f0 < nframes; f0 += nb) {
      nb = std::min(fps, nframes - f0);
      q.nframes = nb;
      for (int k = 0; k < nb; ++k) q.seeds[k] = p.seeds[f0 + k];
      VR_CUDA(cudaMemsetAsync(q.qcount, 0, 3 * sizeof(unsigned), ctx->stream));
      dim3 gq(div_up(r->W, 8), div_up(rows, 16), nb);
      k_primary<COUNT, LINEAR, true><<<gq, 128, 0, ctx->stream>>>(q);
      ctx->launches++;
      if (f0 == 0 && nframes > 1) {
        unsigned* pin = reinterpret_cast<unsigned*>(ctx->scratch_host) + 16;
        VR_CUDA(cudaMemcpyAsync(pin, hits, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
        VR_CUDA(cudaStreamSynchronize(ctx->stream));
        fps = (int)std::min<size_t>(VR_MAX_BATCH, std::max<size_t>(1, r->queue_cap / std::max<size_t>(*pin, 1)));
      }
      VR_TRY((launch_pt_select<COUNT, false, LINEAR>(r, q, q.qcount + 1)));
    }
    return VR_OK;
  }
  // The records stay valid while camera, rows and scene are unchanged.  Within one call they are always reused; across
  // calls only on request (vr_renderer_set_primary_reuse(r, 2): the frame_emitter loop calls render_frame once per
  // sample).  With the per-sample counters on, every batch re-marches (k_primary adds its share for the batch).
  const bool same = r->primary_valid && !memcmp(r->primary_pos, pos, 12) && !memcmp(r->primary_dir, dir, 12) &&
                    r->primary_rows[0] == r->row0 && r->primary_rows[1] == r->row1;
  if (!same || COUNT || (first_of_call && !r->primary_across_calls)) {
    VR_CUDA(cudaMemsetAsync(p.qcount, 0, 2 * sizeof(unsigned), ctx->stream));
    VR_CUDA(cudaMemsetAsync(r->bbox_dev, 0x7F, 2 * sizeof(int), ctx->stream));      // min fields: large
    VR_CUDA(cudaMemsetAsync(r->bbox_dev + 2, 0xFF, 2 * sizeof(int), ctx->stream));  // max fields: -1
    p.bbox = r->bbox_dev;
    dim3 g1(div_up(r->W, 8), div_up(rows, 16), 1);
    k_primary<COUNT, LINEAR><<<g1, 128, 0, ctx->stream>>>(p);
    ctx->launches++;
    VR_CUDA(cudaMemcpyAsync(r->bbox_pin, r->bbox_dev, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    r->primary_epoch++;
    memcpy(r->primary_pos, pos, 12); memcpy(r->primary_dir, dir, 12);
    r->primary_rows[0] = r->row0; r->primary_rows[1] = r->row1;
    r->primary_valid = true;
  }
  VR_CUDA(cudaMemsetAsync(p.qcount + 1, 0, sizeof(unsigned), ctx->stream));
#ifdef VR_AB
  if (r->trace_mode == 3 && (LINEAR || r->sdf->surf)) {  // slots in shared memory, packed batches (k_trace_sm)
    static int per_sm = 0;
    const size_t smem = sm_trace_smem_bytes(LINEAR);
    if (!per_sm) {
      VR_CUDA(cudaFuncSetAttribute(k_trace_sm<COUNT, true, LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      VR_CUDA(cudaFuncSetAttribute(k_trace_sm<COUNT, true, LINEAR>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      VR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace_sm<COUNT, true, LINEAR>, SM_THREADS, smem));
    }
    const unsigned blocks = (unsigned)ctx->sm_count * (unsigned)std::max(1, per_sm);
    k_trace_sm<COUNT, true, LINEAR><<<blocks, SM_THREADS, smem, ctx->stream>>>(p, p.qcount + 1);
    ctx->launches++;
    return VR_OK;
  }
#endif
  return launch_pt_select<COUNT, true, LINEAR>(r, p, p.qcount + 1);
}

int vrk_render(vr_renderer* r, const float pos[3], const float dir[3], const int32_t* seeds, int nframes, bool trace,
               bool resolve, bool first_of_call) {
  vr_ctx* ctx = r->ctx;
  const int rows = r->row1 - r->row0;
  if (rows <= 0) return VR_OK;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  if (nframes < 1 || nframes > VR_MAX_BATCH) { vr_set_error("vrk_render: bad batch size"); return VR_ERR_INVALID; }
  const bool lin = r->sampling == VR_SAMPLING_HW_LINEAR;
  if (trace && lin && (!r->vol_tex || !r->env_tex || !r->lin_surf)) {  // textures and step field are built by the flush
    vr_set_error("vrk_render: hw-linear sampling needs a vr_renderer_flush after vr_renderer_set_sampling");
    return VR_ERR_INVALID;
  }
  if (r->timing && trace) {
    while (r->ev.size() < r->ev_used + 3) {
      cudaEvent_t e;
      VR_CUDA(cudaEventCreate(&e));
      r->ev.push_back(e);
    }
    e0 = r->ev[r->ev_used]; e1 = r->ev[r->ev_used + 1]; e2 = r->ev[r->ev_used + 2];
    r->ev_used += 3;
    r->ev_frames.push_back(nframes);
    VR_CUDA(cudaEventRecord(e0, ctx->stream));
  }
  if (trace) {
    // cache dirtiness for the sparse frame reset: one camera / row window since the last reset -> only cache[hit[]] is touched
    if (r->cache_dirty == 0) {
      r->cache_dirty = 1;
      memcpy(r->dirty_pos, pos, 12); memcpy(r->dirty_dir, dir, 12);
      r->dirty_rows[0] = r->row0; r->dirty_rows[1] = r->row1;
    } else if (r->cache_dirty == 1 && (memcmp(r->dirty_pos, pos, 12) || memcmp(r->dirty_dir, dir, 12) || r->dirty_rows[0] != r->row0 ||
                                       r->dirty_rows[1] != r->row1)) {
      r->cache_dirty = 2;
    }
    RenderParams p;
    p.vol = VolView{r->vol->current(), r->vol->nx, r->vol->ny, r->vol->nz};
    p.sdf = SdfView{r->sdf->field, r->sdf->nx, r->sdf->ny, r->sdf->nz, r->sdf->nx / 8 + 1, r->sdf->ny / 8 + 1};
    p.sdf_surf = r->sdf->surf;
    p.fnx = (float)r->vol->nx; p.fny = (float)r->vol->ny; p.fnz = (float)r->vol->nz;
    p.vol_tex = r->vol_tex; p.env_tex = r->env_tex;
    p.lin_surf = r->lin_surf;
    p.env = r->env->texels;
    p.env_w = r->env->w;
    p.env_h = r->env->h;
    p.cache = r->cache;
    p.hit = r->hit;
    p.frame = r->frame;
    p.W = r->W; p.H = r->H; p.row0 = r->row0; p.row1 = r->row1;
    p.blk_rows = std::max(r->blk_rows, 1); p.blk_rank = r->blk_rank; p.blk_n = r->blk_n;
    p.cam_pos = {pos[0], pos[1], pos[2]};
    p.cam_dir = {dir[0], dir[1], dir[2]};
    camera_basis(dir, &p.cam_side, &p.cam_up);
    p.nframes = nframes;
    p.pixel_major = r->tune.pixel_major;
    p.nframes_shift = -1;
    if (p.pixel_major == 1 && (nframes & (nframes - 1)) == 0)
      for (int b = 0; b < 8; ++b)
        if ((1 << b) == nframes) p.nframes_shift = b;
    p.rule_a = r->tune.rule[0]; p.rule_b = r->tune.rule[1];
    p.lin_sched = r->tune.lin_sched;
    p.lin_wf = r->tune.lin_w[0]; p.lin_ws = r->tune.lin_w[1]; p.lin_we = r->tune.lin_w[2];
    p.spc = r->tune.spc;
    p.sm_k = r->tune.sm_k; p.sm_leave = r->tune.sm_leave;
    for (int k = 0; k < nframes; ++k) p.seeds[k] = seeds[k];
    p.token_cap = r->token_cap;
    p.counters = r->counters;
    p.queue = nullptr; p.qcount = nullptr; p.qcap = 0;
    p.bbox = nullptr;
    p.tf = r->tf_active;
    int st;
    if (lin) st = r->count ? launch_trace<true, true>(r, p, pos, dir, rows, nframes, first_of_call)
                           : launch_trace<false, true>(r, p, pos, dir, rows, nframes, first_of_call);
    else st = r->count ? launch_trace<true, false>(r, p, pos, dir, rows, nframes, first_of_call)
                       : launch_trace<false, false>(r, p, pos, dir, rows, nframes, first_of_call);
    VR_TRY(st);
  }
  if (e1) VR_CUDA(cudaEventRecord(e1, ctx->stream));
  if (resolve) {
    const size_t n = (size_t)r->W * rows;
    k_resolve<<<div_up(n, 256), 256, 0, ctx->stream>>>(r->hit, reinterpret_cast<const uint2*>(r->cache), r->frame, r->W,
                                                       r->row0, r->row1);
    ctx->launches++;
  }
  if (e2) VR_CUDA(cudaEventRecord(e2, ctx->stream));
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// ---- compact cache exchange for the spp split (multi-GPU, no reference counterpart) -----------------------------------
// Every rank of an spp split traces the SAME camera, so the primary hit voxel of a pixel — and therefore the set of cache
// entries touched since the last reset — is identical on all ranks.  Instead of all-reducing the dense cache (8 bytes x voxels:
// 1 GiB at 512^3) the ranks exchange one 8-byte entry per SHADED pixel: the shaded pixels are numbered in pixel order (a
// prefix sum over the hit buffer, identical on all ranks because the hit buffer is), gathered into a dense array of that many
// entries (1.3 MB instead of 16.6 MB at the bench's 8 % shaded pixels), summed across ranks as uint32 words — 16-bit lanes
// cannot carry with a per-rank token cap of 256/N — and written back by the kernel that also resolves the frame.
//   k_xc_count / k_xc_scan / k_xc_gather : xchg[rank_of(pix)] = cache[hit[pix]]
//   (vr_comm.cu: ncclAllReduce over 2 * count uint32)
//   k_xc_scatter_resolve : cache[hit[pix]] = xchg[..] (pixels sharing a voxel write the same global sum) and frame[pix] resolved
__global__ void __launch_bounds__(256) k_xc_count(const uint32_t* __restrict__ hit, size_t n, unsigned* __restrict__ block_counts) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const int c = __syncthreads_count(i < n && hit[i] != VR_MISS);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (unsigned)c;
}
// one block: exclusive scan of nb counts in place, counts[nb] = total
__global__ void __launch_bounds__(1024) k_xc_scan(unsigned* __restrict__ counts, unsigned nb) {
  __shared__ unsigned s[1024];
  const unsigned t = threadIdx.x, chunk = (nb + 1023u) / 1024u;
  const unsigned lo = min(t * chunk, nb), hi = min(lo + chunk, nb);
  unsigned sum = 0;
  for (unsigned i = lo; i < hi; ++i) sum += counts[i];
  s[t] = sum;
  __syncthreads();
  for (unsigned o = 1; o < 1024; o <<= 1) {
    const unsigned v = t >= o ? s[t - o] : 0u;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  unsigned run = s[t] - sum;
  for (unsigned i = lo; i < hi; ++i) { const unsigned c = counts[i]; counts[i] = run; run += c; }
  if (t == 1023) counts[nb] = s[t];
}
// WIDE: the four 16-bit lanes of an entry go into four uint32 words, for ranks that each accumulate up to the full cap of 256
// tokens (a weak-scaling spp split: N x 64 spp with N x 256 tokens per voxel in total) — the packed lanes would overflow in the sum
template <bool WIDE>
__global__ void __launch_bounds__(256) k_xc_gather(const uint32_t* __restrict__ hit, const uint2* __restrict__ cache,
                                                   const unsigned* __restrict__ block_off, uint32_t* __restrict__ cidx,
                                                   uint2* __restrict__ xchg, size_t n) {
  __shared__ unsigned wsum[8];
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t voxel = i < n ? hit[i] : VR_MISS;
  const bool shaded = voxel != VR_MISS;
  const unsigned m = __ballot_sync(0xffffffffu, shaded);
  if (lane == 0) wsum[warp] = (unsigned)__popc(m);
  __syncthreads();
  unsigned off = block_off[blockIdx.x];
  for (unsigned w = 0; w < warp; ++w) off += wsum[w];
  if (shaded) {
    const unsigned idx = off + (unsigned)__popc(m & ((1u << lane) - 1u));
    cidx[i] = idx;
    const uint2 c = cache[voxel];
    if (WIDE) reinterpret_cast<uint4*>(xchg)[idx] = make_uint4(c.x & 0xFFFFu, c.x >> 16, c.y & 0xFFFFu, c.y >> 16);
    else xchg[idx] = c;
  }
}
__global__ void __launch_bounds__(256) k_xc_scatter_resolve(const uint32_t* __restrict__ hit, const uint32_t* __restrict__ cidx,
                                                            const uint2* __restrict__ xchg, uint2* __restrict__ cache,
                                                            uchar4* __restrict__ frame, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t voxel = hit[i];
  if (voxel == VR_MISS) return;
  const uint2 e = xchg[cidx[i]];
  cache[voxel] = e;
  frame[i] = resolve_entry(e);
}
// wide sums: resolved as they are (ray_marching.cl:82-99 on the global totals); the per-rank caches keep their partial sums
__global__ void __launch_bounds__(256) k_xc_resolve_wide(const uint32_t* __restrict__ hit, const uint32_t* __restrict__ cidx,
                                                         const uint4* __restrict__ xw, uchar4* __restrict__ frame, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (hit[i] == VR_MISS) return;
  const uint4 e = xw[cidx[i]];
  frame[i] = resolve_rgbw(e.x, e.y, e.z, e.w);
}

// gather: returns the number of entries through *count_dev (device, read by the caller); needs the full frame traced
int vrk_xc_gather(vr_renderer* r, unsigned** count_dev, bool wide) {
  vr_ctx* ctx = r->ctx;
  const size_t n = (size_t)r->W * r->H;
  const unsigned nb = div_up(n, 256);
  if (!r->xchg) VR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&r->xchg), n * sizeof(uint4), ctx->stream));  // room for the wide form
  if (!r->xc_idx) VR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&r->xc_idx), n * sizeof(uint32_t), ctx->stream));
  if (!r->xc_counts) VR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&r->xc_counts), ((size_t)nb + 1) * sizeof(unsigned), ctx->stream));
  k_xc_count<<<nb, 256, 0, ctx->stream>>>(r->hit, n, r->xc_counts);
  k_xc_scan<<<1, 1024, 0, ctx->stream>>>(r->xc_counts, nb);
  if (wide) k_xc_gather<true><<<nb, 256, 0, ctx->stream>>>(r->hit, reinterpret_cast<const uint2*>(r->cache), r->xc_counts, r->xc_idx, r->xchg, n);
  else k_xc_gather<false><<<nb, 256, 0, ctx->stream>>>(r->hit, reinterpret_cast<const uint2*>(r->cache), r->xc_counts, r->xc_idx, r->xchg, n);
  ctx->launches += 3;
  VR_CUDA(cudaGetLastError());
  *count_dev = r->xc_counts + nb;
  return VR_OK;
}
int vrk_xc_scatter_resolve(vr_renderer* r, bool wide) {
  const size_t n = (size_t)r->W * r->H;
  if (wide) k_xc_resolve_wide<<<div_up(n, 256), 256, 0, r->ctx->stream>>>(r->hit, r->xc_idx, reinterpret_cast<const uint4*>(r->xchg), r->frame, n);
  else k_xc_scatter_resolve<<<div_up(n, 256), 256, 0, r->ctx->stream>>>(r->hit, r->xc_idx, r->xchg, reinterpret_cast<uint2*>(r->cache),
                                                                  r->frame, n);
  r->ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// the per-pixel form (one entry per pixel, zeros for environment pixels): kept for callers that run the collective themselves
// on the raw device pointer (vr_renderer_xchg_device_ptr)
__global__ void __launch_bounds__(256) k_xchg_gather(const uint32_t* __restrict__ hit, const uint2* __restrict__ cache,
                                                     uint2* __restrict__ xchg, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t voxel = hit[i];
  xchg[i] = voxel == VR_MISS ? make_uint2(0u, 0u) : cache[voxel];
}
__global__ void __launch_bounds__(256) k_xchg_scatter(const uint32_t* __restrict__ hit, uint2* __restrict__ cache,
                                                      const uint2* __restrict__ xchg, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t voxel = hit[i];
  if (voxel != VR_MISS) cache[voxel] = xchg[i];
}

int vrk_xchg(vr_renderer* r, uint2* xchg, bool scatter) {
  const size_t n = (size_t)r->W * r->H;
  if (scatter)
    k_xchg_scatter<<<div_up(n, 256), 256, 0, r->ctx->stream>>>(r->hit, reinterpret_cast<uint2*>(r->cache), xchg, n);
  else
    k_xchg_gather<<<div_up(n, 256), 256, 0, r->ctx->stream>>>(r->hit, reinterpret_cast<const uint2*>(r->cache), xchg, n);
  r->ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

// ---- device RNG known-answer dump (tests/test_parity_gpu.py::test_device_rng_known_answers) ------------------------------------------
// Runs the device functions the trace kernels call (rng_triple, hemisphere_reflective_p) over a caller-supplied (seed, gid0, gid1)
// list and returns the integer triples, the components and the final directions for a fixed normal / roughness per item.
__global__ void __launch_bounds__(128) k_rng_dump(const int32_t* __restrict__ seeds, const uint32_t* __restrict__ gid, int n,
                                                  const float* __restrict__ normal_rough, int32_t* __restrict__ ra_out,
                                                  int32_t* __restrict__ comp_out, float* __restrict__ dir_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned xyp = (gid[2 * i] + 1u) * (gid[2 * i + 1] + 1u);
  int ra[3], comp[3];
  rng_triple(seeds[i], xyp, ra, comp);
  for (int k = 0; k < 3; ++k) { ra_out[3 * i + k] = ra[k]; comp_out[3 * i + k] = comp[k]; }
  const f3 nrm = {normal_rough[4 * i], normal_rough[4 * i + 1], normal_rough[4 * i + 2]};
  const f3 d = hemisphere_reflective_p(nrm, seeds[i], normal_rough[4 * i + 3], xyp);
  dir_out[3 * i] = d.x; dir_out[3 * i + 1] = d.y; dir_out[3 * i + 2] = d.z;
}

// hw-linear fetch known answers (tests): the value get_event_and_value sees at n float positions, through the renderer's texture
__global__ void __launch_bounds__(128) k_linear_fetch(cudaTextureObject_t tex, const float* __restrict__ xyz, int n, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tex_value(tex3D<float>(tex, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
}
int vrk_linear_fetch(vr_ctx* ctx, cudaTextureObject_t tex, const float* xyz_dev, int n, int32_t* out_dev) {
  k_linear_fetch<<<div_up((size_t)n, 128), 128, 0, ctx->stream>>>(tex, xyz_dev, n, out_dev);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}

int vrk_rng_dump(vr_ctx* ctx, const int32_t* seeds_dev, const uint32_t* gid_dev, int n, const float* normal_rough_dev, int32_t* ra_dev,
                 int32_t* comp_dev, float* dir_dev) {
  k_rng_dump<<<div_up((size_t)n, 128), 128, 0, ctx->stream>>>(seeds_dev, gid_dev, n, normal_rough_dev, ra_dev, comp_dev, dir_dev);
  ctx->launches++;
  VR_CUDA(cudaGetLastError());
  return VR_OK;
}
