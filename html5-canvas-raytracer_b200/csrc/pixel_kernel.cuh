// pixel_kernel.cuh — one thread per pixel, the whole Whitted tree of that pixel evaluated in
// genuine post-order with an explicit frame stack (the reference recurses: main.js:270, 276).
//
// T = double : VALIDATION kernel.  Same operators in the same association order as main.js, no FMA
//              contraction (TU compiled with -fmad=false), IEEE div/sqrt.  Target: +-1 LSB on
//              >= 99.99 % of pixels against the float64 reference.
// T = float  : the same control flow in fp32 (small scenes / cross-check of the wavefront path).
//
// Result-neutral departures from the reference's control flow:
//   * the hit record (p, n, l, u, v — main.js:440-449) is built for the closest sphere only, not for
//     every intersected sphere;
//   * sphere geometry is staged once per CTA in shared memory.
#pragma once
#include "rtb_device.cuh"

template <typename T>
struct PixelFrame {       // what intersectWorld keeps live across its two recursive calls
  T p[3], n[3], l[3];     // hit.p / hit.n / hit.l          main.js:442-445
  T dir[3];               // incoming direction (specular)  main.js:312
  T rdir[3];              // refract_dir, parked while the reflect subtree runs (main.js:241-266)
  T refl[3];              // reflect_color * albedo[3]      main.js:270-271
  T u, v;                 // hit.u / hit.v                  main.js:446-447
  int32_t hit_i;          // main.js:229
  int32_t segs;
  int32_t stage;          // 0 = before reflect child, 1 = before refract child, 2 = children done
  int32_t has_refract;    // refract_len != 0
  uint32_t node_id;       // path code inside the pixel: root 1, reflect child 2k, refract child 2k+1
                          // (RTB_STARS_HASH key; traversal-order independent)
};

struct PixelCounters { unsigned long long nodes, shadow_rays, sphere_tests; };

// intersectSphere main.js:420-439 (the t part).  g = {Cx,Cy,Cz,r2}.  Returns Infinity on a miss and
// leaves (t0,t1) for the `l` flip of main.js:445.
template <typename T>
RTB_DEVINL T intersect_sphere_t(const T* __restrict__ g, const T org[3], const T dir[3], T eps, T& t0, T& t1) {
  const T INF = Num<T>::inf();
  const T L0 = g[0] - org[0], L1 = g[1] - org[1], L2 = g[2] - org[2];   // between3D(org, obj.origin)
  const T tca = dir[0] * L0 + dir[1] * L1 + dir[2] * L2;
  const T d2 = (L0 * L0 + L1 * L1 + L2 * L2) - tca * tca;
  if (d2 > g[3]) return INF;
  const T thc = Num<T>::sqrt(g[3] - d2);
  t0 = tca - thc;
  t1 = tca + thc;
  T t;
  if (t0 < t1) {
    if (t0 < eps) { if (t1 < eps) return INF; else t = t1; } else t = t0;
  } else {
    if (t1 < eps) { if (t0 < eps) return INF; else t = t0; } else t = t1;
  }
  return t;
}

// Lighting block main.js:280-318 for one node.  Returns diffuse/specular intensities.
template <typename T>
RTB_DEVINL void shade_lights(const DevSceneT<T>& sc, const T* __restrict__ geom, const DevMaterial<T>& m,
                             const PixelFrame<T>& f, T& diffuse_intensity, T& specular_intensity,
                             PixelCounters& cnt) {
  diffuse_intensity = T(0);
  specular_intensity = T(0);
  if (!((m.albedo[1] > T(0)) || (m.albedo[2] > T(0)))) return;
  const int n = sc.n_spheres;
  T light_intensity = sc.light_intensity;              // ONCE, outside the loop (main.js:284)
  for (int k = 0; k < sc.n_lights; k++) {
    T sv[3] = {sc.lights[k][0] - f.p[0], sc.lights[k][1] - f.p[1], sc.lights[k][2] - f.p[2]};
    const T light_mag = sv[0] * sv[0] + sv[1] * sv[1] + sv[2] * sv[2];
    const T light_len = Num<T>::sqrt(light_mag);
    if (light_len != T(0)) { const T r = T(1) / light_len; sv[0] *= r; sv[1] *= r; sv[2] *= r; }
    const T shadow_dot = sv[0] * f.l[0] + sv[1] * f.l[1] + sv[2] * f.l[2];
    if (shadow_dot <= T(0)) continue;                  // main.js:292
    cnt.shadow_rays++;
    for (int j = 0; j < n; j++) {                      // main.js:293-304
      if (j == f.hit_i) continue;
      T t0, t1;
      const T t = intersect_sphere_t<T>(geom + 4 * j, f.p, sv, sc.eps, t0, t1);
      cnt.sphere_tests++;
      if (t < light_len) {
        const T a4 = sc.mtl[j].albedo[4];
        if (a4 != T(0)) light_intensity /= a4;
        else { light_intensity = T(0); break; }
      }
    }
    if (light_intensity == T(0)) continue;             // main.js:305
    diffuse_intensity += light_intensity * shadow_dot / light_mag;
    if (m.albedo[2] > T(0)) {                          // main.js:307-314
      const T ld[3] = {-sv[0], -sv[1], -sv[2]};
      const T tt = -(T(2) * (ld[0] * f.l[0] + ld[1] * f.l[1] + ld[2] * f.l[2]));
      T lr[3] = {ld[0] + f.l[0] * tt, ld[1] + f.l[1] * tt, ld[2] + f.l[2] * tt};
      const T ll = Num<T>::sqrt(lr[0] * lr[0] + lr[1] * lr[1] + lr[2] * lr[2]);
      if (ll != T(0)) { const T r = T(1) / ll; lr[0] *= r; lr[1] *= r; lr[2] *= r; }
      const T spec_dot = f.dir[0] * (-lr[0]) + f.dir[1] * (-lr[1]) + f.dir[2] * (-lr[2]);
      if (spec_dot > T(0)) specular_intensity += Num<T>::pow(spec_dot, m.specular_exponent);
    }
  }
  diffuse_intensity = js_min<T>(T(1), diffuse_intensity) * m.albedo[1];
  specular_intensity = js_min<T>(T(1), specular_intensity) * m.albedo[2];
}

// intersectWorld(segs, objs, org, dir) main.js:220-337, iteratively.
template <typename T>
__device__ void trace_tree(const DevSceneT<T>& sc, const T* __restrict__ geom, const DevFrameT<T>& fr,
                           const T org0[3], const T dir0[3], uint32_t px, uint32_t py, uint32_t sub,
                           T out[3], PixelCounters& cnt, unsigned long long star_base, uint32_t& star_k) {
  const T INF = Num<T>::inf();
  PixelFrame<T> stack[RTB_MAX_SEGS];
  int sp = -1;
  uint32_t path = 1;
  T org[3] = {org0[0], org0[1], org0[2]};
  T dir[3] = {dir0[0], dir0[1], dir0[2]};
  int segs = fr.segs;
  T ret[3] = {T(0), T(0), T(0)};
  enum { CALL, RETURN, RESUME };
  int mode = CALL;
  const int n = sc.n_spheres;

  for (;;) {
    if (mode == CALL) {
      if (segs == 0) {                                   // main.js:221
        ret[0] = ret[1] = ret[2] = T(0);
        mode = RETURN;
      } else {
        cnt.nodes++;
        const uint32_t node_id = path;
        int hit_i = -1;                                  // main.js:223-230
        T best = INF, bt0 = T(0), bt1 = T(0);
        // fp32 only: the sphere this ray starts on is evaluated analytically (roots {0, 2 tca}), as in
        // wavefront.cu's slow_general — the float64 kernel keeps the reference's arithmetic untouched
        const int self_i = (sizeof(T) == 4 && sp >= 0) ? stack[sp].hit_i : -1;
        for (int i = 0; i < n; i++) {
          T t0, t1;
          T t;
          if (i == self_i) {
            const T* g = geom + 4 * i;
            const T tca = dir[0] * (g[0] - org[0]) + dir[1] * (g[1] - org[1]) + dir[2] * (g[2] - org[2]);
            t0 = T(0); t1 = T(2) * tca;
            t = (t1 >= sc.eps) ? t1 : INF;
          } else {
            t = intersect_sphere_t<T>(geom + 4 * i, org, dir, sc.eps, t0, t1);
          }
          if (t < best) { best = t; hit_i = i; bt0 = t0; bt1 = t1; }
        }
        cnt.sphere_tests += (unsigned long long)n;
        if (hit_i < 0) {                                 // main.js:231
          ret[0] = T(1); ret[1] = T(0); ret[2] = T(0);
          mode = RETURN;
        } else {
          PixelFrame<T>& f = stack[++sp];
          const T* g = geom + 4 * hit_i;
          const DevMaterial<T>& m = sc.mtl[hit_i];
          f.hit_i = hit_i; f.segs = segs; f.stage = 0; f.node_id = node_id;
          f.dir[0] = dir[0]; f.dir[1] = dir[1]; f.dir[2] = dir[2];
          // hit record main.js:441-447
          f.p[0] = org[0] + dir[0] * best; f.p[1] = org[1] + dir[1] * best; f.p[2] = org[2] + dir[2] * best;
          T nn[3] = {f.p[0] - g[0], f.p[1] - g[1], f.p[2] - g[2]};
          const T nl = Num<T>::sqrt(nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2]);
          if (nl != T(0)) { const T r = T(1) / nl; nn[0] *= r; nn[1] *= r; nn[2] *= r; }
          f.n[0] = nn[0]; f.n[1] = nn[1]; f.n[2] = nn[2];
          const bool flip = (bt0 < sc.eps) || (bt1 < sc.eps);
          f.l[0] = flip ? -nn[0] : nn[0]; f.l[1] = flip ? -nn[1] : nn[1]; f.l[2] = flip ? -nn[2] : nn[2];
          f.u = Num<T>::atan2(-nn[2], -nn[0]) / Num<T>::pi() / T(2) + T(0.5);
          f.v = Num<T>::asin(-nn[1]) / (Num<T>::pi() / T(2)) / T(2) + T(0.5);
          f.refl[0] = f.refl[1] = f.refl[2] = T(0);

          // reflect direction main.js:233-239
          T rd[3] = {T(0), T(0), T(0)};
          T reflect_len = T(0);
          if (m.albedo[3] > T(0)) {
            const T tt = -(T(2) * (dir[0] * nn[0] + dir[1] * nn[1] + dir[2] * nn[2]));
            rd[0] = dir[0] + nn[0] * tt; rd[1] = dir[1] + nn[1] * tt; rd[2] = dir[2] + nn[2] * tt;
            reflect_len = Num<T>::sqrt(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
            if (reflect_len != T(0)) { const T r = T(1) / reflect_len; rd[0] *= r; rd[1] *= r; rd[2] *= r; }
          }
          // refract direction main.js:241-266
          f.has_refract = 0;
          f.rdir[0] = f.rdir[1] = f.rdir[2] = T(0);
          if (m.albedo[4] > T(0)) {
            T norm[3], eta;
            const T dot = dir[0] * nn[0] + dir[1] * nn[1] + dir[2] * nn[2];
            T cosi = -js_max<T>(T(-1), js_min<T>(T(1), dot));
            if (cosi < T(0)) {
              cosi = -cosi; norm[0] = -nn[0]; norm[1] = -nn[1]; norm[2] = -nn[2];
              eta = m.refract_index;
            } else {
              norm[0] = nn[0]; norm[1] = nn[1]; norm[2] = nn[2];
              eta = T(1) / m.refract_index;
            }
            const T k = T(1) - eta * eta * (T(1) - cosi * cosi);
            T fd[3];
            if (k > T(0)) {
              const T q = eta * cosi - Num<T>::sqrt(k);
              fd[0] = dir[0] * eta + norm[0] * q; fd[1] = dir[1] * eta + norm[1] * q; fd[2] = dir[2] * eta + norm[2] * q;
            } else {                                     // total internal reflection main.js:261-263
              const T tt = -(T(2) * (dir[0] * norm[0] + dir[1] * norm[1] + dir[2] * norm[2]));
              fd[0] = dir[0] + norm[0] * tt; fd[1] = dir[1] + norm[1] * tt; fd[2] = dir[2] + norm[2] * tt;
            }
            const T fl = Num<T>::sqrt(fd[0] * fd[0] + fd[1] * fd[1] + fd[2] * fd[2]);
            if (fl != T(0)) { const T r = T(1) / fl; fd[0] *= r; fd[1] *= r; fd[2] *= r; f.has_refract = 1; }
            f.rdir[0] = fd[0]; f.rdir[1] = fd[1]; f.rdir[2] = fd[2];
          }
          // first recursive call main.js:268-272
          f.stage = 1;
          if (reflect_len != T(0)) {
            org[0] = f.p[0]; org[1] = f.p[1]; org[2] = f.p[2];
            dir[0] = rd[0]; dir[1] = rd[1]; dir[2] = rd[2];
            segs = f.segs - 1;
            path = 2u * f.node_id;
            mode = CALL;
          } else {
            mode = RESUME;
          }
        }
      }
    }
    if (mode == RETURN) {
      if (sp < 0) break;
      PixelFrame<T>& f = stack[sp];
      const DevMaterial<T>& m = sc.mtl[f.hit_i];
      if (f.stage == 1) {                                // back from the reflect child
        f.refl[0] = ret[0] * m.albedo[3]; f.refl[1] = ret[1] * m.albedo[3]; f.refl[2] = ret[2] * m.albedo[3];
        mode = RESUME;
      } else {                                           // back from the refract child: finish node
        const T rc[3] = {ret[0] * m.albedo[4], ret[1] * m.albedo[4], ret[2] * m.albedo[4]};
        T di, si, rgb[3];
        shade_lights<T>(sc, geom, m, f, di, si, cnt);
        // the sampler runs AFTER both recursive calls (main.js:320): post-order, which is the order the
        // reference consumes Math.random() in — star_k numbers this pixel's draws
        unsigned long long star_idx = 0;
        if (m.sampler_kind == RTB_SAMPLER_STARS) star_idx = star_base + (star_k++);
        run_sampler<T>(sc, m, f.n, f.u, f.v, fr.star_mode, fr.star_seed, px, py, sub, f.node_id, rgb,
                       fr.star_stream, star_idx, fr.star_stream_len);
        for (int c = 0; c < 3; c++) {                    // main.js:322-336
          const T shader = rgb[c] * di + rgb[c] * si + f.refl[c] + rc[c];
          ret[c] = js_max<T>(rgb[c] * m.albedo[0], js_min<T>(T(1), shader));
        }
        sp--;
        mode = RETURN;
        continue;
      }
    }
    if (mode == RESUME) {                                // reflect subtree done (or absent)
      PixelFrame<T>& f = stack[sp];
      f.stage = 2;
      if (f.has_refract) {                               // second recursive call main.js:274-278
        org[0] = f.p[0]; org[1] = f.p[1]; org[2] = f.p[2];
        dir[0] = f.rdir[0]; dir[1] = f.rdir[1]; dir[2] = f.rdir[2];
        segs = f.segs - 1;
        path = 2u * f.node_id + 1u;
        mode = CALL;
      } else {
        ret[0] = ret[1] = ret[2] = T(0);                 // refract_color = [0,0,0]
        mode = RETURN;
      }
    }
  }
  out[0] = ret[0]; out[1] = ret[1]; out[2] = ret[2];
}

// Primary ray main.js:186-193 for sub-sample (i,j) of a g x g stratified grid (g = 1: reference).
template <typename T, typename F>   // F: DevFrameT<T> or the wavefront pipeline's per-frame camera
RTB_DEVINL void primary_ray(const F& fr, int x, int y, int i, int j, T ray[3]) {
  const int g = fr.spp_grid;
  const T ox = (T(i) + T(0.5)) / T(g), oy = (T(j) + T(0.5)) / T(g);
  const T dist[3] = {T(x) - fr.projW + ox, fr.projH - T(y) - oy, fr.projD};
  T target[3];
  if (fr.camera_mode == RTB_CAMERA_REFERENCE) {          // per-component dist
#pragma unroll
    for (int k = 0; k < 3; k++)
      target[k] = fr.origin[k] + fr.axisX[k] * dist[k] + fr.axisY[k] * dist[k] + fr.axisZ[k] * dist[k];
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++)
      target[k] = fr.origin[k] + fr.axisX[k] * dist[0] + fr.axisY[k] * dist[1] + fr.axisZ[k] * dist[2];
  }
  T r[3] = {target[0] - fr.origin[0], target[1] - fr.origin[1], target[2] - fr.origin[2]};  // main.js:192
  const T l = Num<T>::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (l != T(0)) { const T q = T(1) / l; r[0] *= q; r[1] *= q; r[2] *= q; }
  ray[0] = r[0]; ray[1] = r[1]; ray[2] = r[2];
}

// Kernel: thread -> pixel (x fastest).  Block = 8x4 pixel tiles x 8 would help coherence; rows of
// 32 consecutive x already share most of their tree in these scenes, and row-major keeps the
// framebuffer store a coalesced 128-byte line per warp.
template <typename T>
__global__ void __launch_bounds__(128)
pixel_kernel(DevSceneT<T> sc, DevFrameT<T> fr, RowMap rm, uint8_t* __restrict__ out,
             unsigned long long* __restrict__ counters, int geom_in_smem) {
  extern __shared__ unsigned char smem_raw[];
  T* sg = reinterpret_cast<T*>(smem_raw);
  const T* geom = sc.geom;
  if (geom_in_smem) {
    for (int i = threadIdx.x; i < sc.n_spheres * 4; i += blockDim.x) sg[i] = sc.geom[i];
    __syncthreads();
    geom = sg;
  }
  const int W = fr.width;
  const long long total = (long long)rm.n_local() * W;
  PixelCounters cnt = {0ull, 0ull, 0ull};
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int lrow = (int)(idx / W), x = (int)(idx % W);
    const int y = rm.local_to_y(lrow);
    const int g = fr.spp_grid;
    T acc[3] = {T(0), T(0), T(0)};
    uint32_t star_k = 0;
    const unsigned long long star_base = fr.star_base ? fr.star_base[(size_t)y * W + x] : 0ull;
    for (int j = 0; j < g; j++)
      for (int i = 0; i < g; i++) {
        T ray[3], rgb[3];
        primary_ray<T>(fr, x, y, i, j, ray);
        trace_tree<T>(sc, geom, fr, fr.origin, ray, (uint32_t)x, (uint32_t)y, (uint32_t)(j * g + i), rgb, cnt,
                      star_base, star_k);
        if (g == 1) { acc[0] = rgb[0]; acc[1] = rgb[1]; acc[2] = rgb[2]; }
        else { acc[0] += rgb[0]; acc[1] += rgb[1]; acc[2] += rgb[2]; }
      }
    if (fr.star_counts) { fr.star_counts[(size_t)y * W + x] = star_k; continue; }   // counting pass: no pixel
    if (g > 1) { const T q = T(g * g); acc[0] /= q; acc[1] /= q; acc[2] /= q; }
    const uint32_t r8 = clamped_u8(T(255) * acc[0]), g8 = clamped_u8(T(255) * acc[1]),
                   b8 = clamped_u8(T(255) * acc[2]);
    uint32_t* o32 = reinterpret_cast<uint32_t*>(out) + rm.out_row(lrow) * (size_t)W + x;
    *o32 = r8 | (g8 << 8) | (b8 << 16) | (255u << 24);   // main.js:195-198, one 32-bit store
  }
  // warp-aggregate the counters, one atomic per warp per counter
  for (int o = 16; o > 0; o >>= 1) {
    cnt.nodes += __shfl_down_sync(0xffffffffu, cnt.nodes, o);
    cnt.shadow_rays += __shfl_down_sync(0xffffffffu, cnt.shadow_rays, o);
    cnt.sphere_tests += __shfl_down_sync(0xffffffffu, cnt.sphere_tests, o);
  }
  if ((threadIdx.x & 31) == 0 && counters) {
    atomicAdd(&counters[0], cnt.nodes);
    atomicAdd(&counters[1], cnt.shadow_rays);
    atomicAdd(&counters[2], cnt.sphere_tests);
  }
}
