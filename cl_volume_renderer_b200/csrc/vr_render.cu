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
#include "vr_device.cuh"

struct RenderParams {
  VolView vol;
  SdfView sdf;
  const uchar4* __restrict__ env;
  int env_w, env_h;
  uint32_t* __restrict__ cache;
  uint32_t* __restrict__ hit;
  uchar4* __restrict__ frame;
  int W, H, row0, row1;
  f3 cam_pos, cam_dir;
  f3 cam_side, cam_up;  // camera basis of generate_ray, evaluated once on the host in the same fp32 op order
  int token_cap;
  int nframes;          // frames in this launch: blockIdx.z selects the seed
  int seeds[VR_MAX_BATCH];
  unsigned long long* counters;
  TfTable tf;
};

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
  return {cam_pos, normalize3(point)};
}

__device__ __forceinline__ bool lim(float p, int dim) { return p <= (float)dim && p >= 0.0f; }

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

// sample_environment_map, utility_environment_map.cl:3-13: normalised coords, clamp to edge, nearest texel
__device__ __forceinline__ uchar4 env_sample(const RenderParams& p, f3 d) {
  float u = atan2f(d.x, d.z);
  float v = asinf(-d.y);
  u = u * 0.1591549431f;
  v = v * 0.318309886f;
  u = u + 0.5f;
  v = v + 0.5f;
  int ix = f2i(floorf(u * (float)p.env_w));
  int iy = f2i(floorf(v * (float)p.env_h));
  ix = min(max(ix, 0), p.env_w - 1);
  iy = min(max(iy, 0), p.env_h - 1);
  return __ldg(p.env + (size_t)iy * p.env_w + ix);
}

// get_hemisphere_direction_reflective, utility_sampling.cl:40-50
__device__ __forceinline__ f3 hemisphere_reflective(f3 normal, int seed, float roughness, unsigned gx, unsigned gy) {
  const uint32_t useed = (uint32_t)seed + (gx + 1u) * (gy + 1u);
  const int rx = (int)hash_u32(useed * 0x182205bdu);
  const int ry = (int)hash_u32(useed * 0xe8d052f3u);
  const int rz = (int)hash_u32(useed * 0xf1981dcfu);
  f3 direction = {(float)((rx % 2048) - 1024), (float)((ry % 2048) - 1024), (float)((rz % 2048) - 1024)};
  const float decider = dot3(direction, normal);
  const f3 correct = normalize3(direction * decider);
  return normalize3(normal * (1.0f - roughness) + correct * roughness);
}

// march_to_next_event, utility_ray.cl:157-168 with march (:148-154) and get_event_and_value (:126-138) fused.
//   r        in/out: the ray, advanced to the event position
//   grad     out: gradient at the hit voxel (valid when EV_HIT) — the shading normal needs it next
//   color    in/out: written only when a TF clause with a colour matched (like `*color = tmp_color`)
template <bool COUNT>
__device__ __forceinline__ int march_to_next_event(const RenderParams& p, Ray& r, f3& grad, int color[4],
                                                   unsigned& steps) {
  const int nx = p.vol.nx, ny = p.vol.ny, nz = p.vol.nz;
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
      }
    }
    return EV_HIT;
  }
  return EV_NONE;
}

template <bool COUNT>
__global__ void __launch_bounds__(128) k_trace(const RenderParams p) {
  // one warp = an 8x4 pixel tile: neighbouring primary rays walk neighbouring voxels
  const int x = blockIdx.x * 8 + (threadIdx.x & 7);
  const int y = p.row0 + blockIdx.y * 16 + (threadIdx.x >> 3);
  unsigned c_steps = 0, c_normals = 0, c_env = 0, c_hits = 0, c_adm = 0, c_samples = 0;
  if (x < p.W && y < p.row1) {
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
    if (is_cut) ev = march_to_next_event<COUNT>(p, cur, grad, color, c_steps);

    if (ev != EV_HIT) {
      // ray_marching.cl:172-178,188-194: environment colour, alpha 200
      uchar4 e = env_sample(p, vray.d);
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
        const f3 normal = -normalize3(grad);
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
            ev = march_to_next_event<COUNT>(p, cur, grad, color, c_steps);
            if (ev == EV_EXIT) {
              const float factor = 8.0f / (float)i;
              const uchar4 lm = env_sample(p, cur.d);
              c_env++;
              // uint += float (ray_marching.cl:59-61): to float, add, truncate back
              bv0 = f2u((float)bv0 + atten * r_energy * (float)lm.x * factor / 1.0f);
              bv1 = f2u((float)bv1 + atten * g_energy * (float)lm.y * factor / 1.0f);
              bv2 = f2u((float)bv2 + atten * b_energy * (float)lm.z * factor / 1.0f);
              break;
            } else if (ev == EV_HIT) {
              const f3 n2 = -normalize3(grad);
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

// phase 2: ray_marching.cl:82-99
__global__ void __launch_bounds__(256) k_resolve(const uint32_t* __restrict__ hit, const uint2* __restrict__ cache,
                                                 uchar4* __restrict__ frame, int W, int row0, int row1) {
  const size_t n = (size_t)W * (row1 - row0);
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t pix = (size_t)row0 * W + i;
  const uint32_t voxel = hit[pix];
  if (voxel == VR_MISS) return;
  const uint2 c = cache[voxel];
  uint32_t r = c.x & 0xFFFFu, g = c.x >> 16, b = c.y & 0xFFFFu, w = c.y >> 16;
  if (w != 0) { r /= w; g /= w; b /= w; } else { r = g = b = 0; }
  const float inv_gamma = 1.0f / 1.77777777f;
  const float brightness = 4.0f;
  float fr = (float)r / 255.0f, fg = (float)g / 255.0f, fb = (float)b / 255.0f;
  fr = powf(fr * brightness, inv_gamma) * 255.0f;
  fg = powf(fg * brightness, inv_gamma) * 255.0f;
  fb = powf(fb * brightness, inv_gamma) * 255.0f;
  frame[pix] = make_uchar4((unsigned char)min(f2u(fr), 255u), (unsigned char)min(f2u(fg), 255u),
                           (unsigned char)min(f2u(fb), 255u), 1);
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

int vrk_render(vr_renderer* r, const float pos[3], const float dir[3], const int32_t* seeds, int nframes, bool trace,
               bool resolve) {
  vr_ctx* ctx = r->ctx;
  const int rows = r->row1 - r->row0;
  if (rows <= 0) return VR_OK;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  if (nframes < 1 || nframes > VR_MAX_BATCH) { vr_set_error("vrk_render: bad batch size"); return VR_ERR_INVALID; }
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
    RenderParams p;
    p.vol = VolView{r->vol->current(), r->vol->nx, r->vol->ny, r->vol->nz};
    p.sdf = SdfView{r->sdf->field, r->sdf->nx, r->sdf->ny, r->sdf->nz, r->sdf->nx / 8 + 1, r->sdf->ny / 8 + 1};
    p.env = r->env->texels;
    p.env_w = r->env->w;
    p.env_h = r->env->h;
    p.cache = r->cache;
    p.hit = r->hit;
    p.frame = r->frame;
    p.W = r->W; p.H = r->H; p.row0 = r->row0; p.row1 = r->row1;
    p.cam_pos = {pos[0], pos[1], pos[2]};
    p.cam_dir = {dir[0], dir[1], dir[2]};
    camera_basis(dir, &p.cam_side, &p.cam_up);
    p.nframes = nframes;
    for (int k = 0; k < nframes; ++k) p.seeds[k] = seeds[k];
    p.token_cap = r->token_cap;
    p.counters = r->counters;
    p.tf = r->tf_active;
    dim3 grid(div_up(r->W, 8), div_up(rows, 16), nframes);
    if (r->count) k_trace<true><<<grid, 128, 0, ctx->stream>>>(p);
    else k_trace<false><<<grid, 128, 0, ctx->stream>>>(p);
    ctx->launches++;
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

// ---- compact cache exchange for the spp split (multi-GPU hook, no reference counterpart) -----------------------------
// Every rank of an spp split traces the SAME camera, so the primary hit voxel of a pixel — and therefore the set of
// cache entries touched since the last reset — is identical on all ranks.  Instead of all-reducing the dense cache
// (8 bytes x voxels: 1 GiB at 512^3) the ranks exchange one 8-byte entry per pixel (16 MB at 1080p):
//   gather : xchg[pix] = cache[hit[pix]]  (0 for environment pixels)
//   (caller: sum-all-reduce xchg as int32 words — 16-bit lanes cannot carry with a per-rank token cap of 256/N)
//   scatter: cache[hit[pix]] = xchg[pix]  (pixels sharing a voxel write the same global sum)
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
