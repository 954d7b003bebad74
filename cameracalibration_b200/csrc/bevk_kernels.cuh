// bevk_kernels.cuh -- the sm_100a kernels behind libbevk.so.
//
//   k_undistort_map   K1  cv2.fisheye.initUndistortRectifyMap / cv2.initUndistortRectifyMap
//   k_gather          K3/K4  cv2.remap, fused undistort (no map), cv2.warpPerspective on images
//   k_warp_maps       K2  cv2.warpPerspective on the 16SC2/16UC1 map planes (BEV LUT build),
//                         optionally fused with K1 so the full-size undistort map never exists
//   k_blend_masks     K10 BlendMask.get_blend_mask
//   k_vsum / k_delta  K8  per-camera sum of V = max(B,G,R) and the luminance offsets
//   k_bev             K3+K5+K6+K7(+K8 per tap): the fused per-frame kernel
//   k_gain            K9  grey-world gains + car overlay
//   k_sat_sum         multi-GPU compose of per-camera partial canvases
#pragma once
#include "bevk_device.cuh"

namespace bevk {

constexpr int BEVK_MAX_CAMERAS_K = 8;   // == BEVK_MAX_CAMERAS in include/bevk.h

// ---------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_undistort_map(CamModel cm, short2* __restrict__ map1,
                                                       unsigned short* __restrict__ map2) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cm.w || y >= cm.h) return;
  double u, v;
  undistort_point(cm, x, y, u, v);
  short mx, my;
  unsigned short fr;
  quantise_uv(u, v, mx, my, fr, pack_saturates(cm.model, x, cm.w));
  const size_t i = (size_t)y * cm.w + x;
  map1[i] = make_short2(mx, my);
  map2[i] = fr;
}

// ---------------------------------------------------------------------------------
// K3 / K4: generic gather.  MODE 0: maps in HBM, 1: camera model evaluated in-kernel
// (fused undistort), 2: homography (warpPerspective).
// ---------------------------------------------------------------------------------
struct GatherArgs {
  const uint8_t* src; int sw, sh; long long spitch;
  uint8_t* dst; int dw, dh; long long dpitch;
  const short2* map1; const unsigned short* map2;
  CamModel cm;
  Homog hm;
};

template <int C>
__device__ __forceinline__ void load_px(const uint8_t* __restrict__ src, long long spitch, int sw, int sh,
                                        int x, int y, int (&p)[C]) {
  if ((unsigned)x < (unsigned)sw && (unsigned)y < (unsigned)sh) {
    const uint8_t* q = src + (long long)y * spitch + (long long)x * C;
#pragma unroll
    for (int c = 0; c < C; ++c) p[c] = __ldg(q + c);
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) p[c] = 0;
  }
}

template <int MODE, int C, int LINEAR>
__global__ void __launch_bounds__(256) k_gather(GatherArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.dw || y >= a.dh) return;
  int sx, sy, fx = 0, fy = 0;
  if (MODE == 2) {
    int X, Y;
    if (LINEAR) {
      warp_point(a.hm, x, y, (double)TAB, X, Y);
      sx = sat_i16(X >> INTER_BITS); sy = sat_i16(Y >> INTER_BITS);
      fx = X & (TAB - 1); fy = Y & (TAB - 1);
    } else {
      warp_point(a.hm, x, y, 1.0, X, Y);
      sx = sat_i16(X); sy = sat_i16(Y);
    }
  } else {
    short mx, my;
    unsigned short fr;
    bool have_frac = true;
    if (MODE == 0) {
      const size_t i = (size_t)y * a.dw + x;
      const short2 m = a.map1[i];
      mx = m.x; my = m.y;
      have_frac = (a.map2 != nullptr);
      fr = have_frac ? a.map2[i] : 0;
    } else {
      double u, v;
      undistort_point(a.cm, x, y, u, v);
      quantise_uv(u, v, mx, my, fr, pack_saturates(a.cm.model, x, a.cm.w));
    }
    sx = mx; sy = my;
    fx = fr & (TAB - 1); fy = (fr >> INTER_BITS) & (TAB - 1);
    if (!LINEAR && have_frac) {  // OpenCV's NNDeltaTab_i is inverted: frac < 16 picks the +1 neighbour
      sx += (fx < 16); sy += (fy < 16);
    }
  }
  uint8_t* o = a.dst + (long long)y * a.dpitch + (long long)x * C;
  if (LINEAR) {
    int p00[C], p01[C], p10[C], p11[C];
    load_px<C>(a.src, a.spitch, a.sw, a.sh, sx, sy, p00);
    load_px<C>(a.src, a.spitch, a.sw, a.sh, sx + 1, sy, p01);
    load_px<C>(a.src, a.spitch, a.sw, a.sh, sx, sy + 1, p10);
    load_px<C>(a.src, a.spitch, a.sw, a.sh, sx + 1, sy + 1, p11);
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = (uint8_t)bilerp_q10(p00[c], p01[c], p10[c], p11[c], fx, fy);
  } else {
    int p[C];
    load_px<C>(a.src, a.spitch, a.sw, a.sh, sx, sy, p);
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = (uint8_t)p[c];
  }
}

// ---------------------------------------------------------------------------------
// K2: warpPerspective of the map planes.  FROM_MODEL=1 evaluates the undistort map at
// the four taps instead of reading und_w x und_h planes from HBM (fused K1+K2).
// ---------------------------------------------------------------------------------
struct WarpMapsArgs {
  const short2* in1; const unsigned short* in2; int sw, sh;   // FROM_MODEL=0
  CamModel cm;                                                // FROM_MODEL=1 (sw,sh = cm.w,cm.h)
  Homog hm;
  short2* out1; unsigned short* out2; int dw, dh;
};

// One destination pixel of that warp (host-capable: tests/host/kernel_math.cu runs it on a CPU against cv2).
template <int FROM_MODEL>
__host__ __device__ __forceinline__ void warp_maps_pixel(const WarpMapsArgs& a, int x, int y, short& ox, short& oy,
                                                         unsigned short& of) {
  int X, Y;
  warp_point(a.hm, x, y, (double)TAB, X, Y);
  const int sx = sat_i16(X >> INTER_BITS), sy = sat_i16(Y >> INTER_BITS);
  const float ax = fmul((float)(X & (TAB - 1)), 1.0f / TAB);
  const float ay = fmul((float)(Y & (TAB - 1)), 1.0f / TAB);
  const float w[4] = {fmul(fsub(1.f, ay), fsub(1.f, ax)), fmul(fsub(1.f, ay), ax), fmul(ay, fsub(1.f, ax)), fmul(ay, ax)};
  float accx = 0.f, accy = 0.f, accf = 0.f;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int tx = sx + (t & 1), ty = sy + (t >> 1);
    float vx = 0.f, vy = 0.f, vf = 0.f;
    if ((unsigned)tx < (unsigned)a.sw && (unsigned)ty < (unsigned)a.sh) {
      short mx, my;
      unsigned short fr;
      if (FROM_MODEL) {
        double u, v;
        undistort_point(a.cm, tx, ty, u, v);
        quantise_uv(u, v, mx, my, fr, pack_saturates(a.cm.model, tx, a.cm.w));
      } else {
        const size_t i = (size_t)ty * a.sw + tx;
        const short2 m = a.in1[i];
        mx = m.x; my = m.y; fr = a.in2[i];
      }
      vx = (float)mx; vy = (float)my; vf = (float)fr;
    }
    // acc = ((t0*w0 + t1*w1) + t2*w2) + t3*w3, each product and sum rounded (no FMA)
    if (t == 0) { accx = fmul(vx, w[0]); accy = fmul(vy, w[0]); accf = fmul(vf, w[0]); }
    else {
      accx = fadd(accx, fmul(vx, w[t]));
      accy = fadd(accy, fmul(vy, w[t]));
      accf = fadd(accf, fmul(vf, w[t]));
    }
  }
  ox = (short)sat_i16(f2i_rn(accx));
  oy = (short)sat_i16(f2i_rn(accy));
  const int fi = f2i_rn(accf);
  of = (unsigned short)(fi < 0 ? 0 : (fi > 65535 ? 65535 : fi));
}

template <int FROM_MODEL>
__global__ void __launch_bounds__(256) k_warp_maps(WarpMapsArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.dw || y >= a.dh) return;
  short ox, oy;
  unsigned short of;
  warp_maps_pixel<FROM_MODEL>(a, x, y, ox, oy, of);
  const size_t o = (size_t)y * a.dw + x;
  a.out1[o] = make_short2(ox, oy);
  a.out2[o] = of;
}

// ---------------------------------------------------------------------------------
// K10: blend weights.  polys/out: uint8[4][h][w]; lines: int[8][4] = (ax,ay,bx,by) for
// FL,FR,BL,BR,LF,LB,RF,RB.
// ---------------------------------------------------------------------------------
struct BlendArgs { const uint8_t* polys; uint8_t* out; int w, h; int lines[8][4]; };

__host__ __device__ __forceinline__ double seg_dist(double px, double py, const int* L) {
  const double ax = L[0], ay = L[1], bx = L[2], by = L[3];
  const double dx = bx - ax, dy = by - ay;
  const double d1x = px - ax, d1y = py - ay, d2x = px - bx, d2y = py - by;
  double sq;
  if (dadd(dmul(d1x, dx), dmul(d1y, dy)) <= 0) sq = dadd(dmul(d1x, d1x), dmul(d1y, d1y));
  else if (dadd(dmul(d2x, dx), dmul(d2y, dy)) >= 0) sq = dadd(dmul(d2x, d2x), dmul(d2y, d2y));
  else {
    const double cr = dadd(dmul(d1y, dx), -dmul(d1x, dy));
    sq = ddiv(dmul(cr, cr), dadd(dmul(dx, dx), dmul(dy, dy)));
  }
  return dsqrt(sq);
}

// One canvas pixel of the four blend masks (host-capable for tests/host/kernel_math.cu).
__host__ __device__ __forceinline__ void blend_pixel(const BlendArgs& a, int x, int y) {
  const size_t plane = (size_t)a.w * a.h, p = (size_t)y * a.w + x;
  uint8_t m[4];
#pragma unroll
  for (int n = 0; n < 4; ++n) m[n] = a.polys[n * plane + p];
  // (other, lineA, lineB) per camera, in the reference's order (surroundBEV.py:171-186)
  const int other[4][2] = {{2, 3}, {2, 3}, {0, 1}, {0, 1}};
  const int la[4][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}};
  const int lb[4][2] = {{4, 6}, {5, 7}, {0, 2}, {1, 3}};
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    int v = m[n];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (v != 0 && m[other[n][k]] != 0) {
        const double dA = seg_dist((double)x, (double)y, a.lines[la[n][k]]);
        const double dB = seg_dist((double)x, (double)y, a.lines[lb[n][k]]);
        const double a2 = dmul(dA, dA), b2 = dmul(dB, dB);
        v = (int)dmul(ddiv(a2, dadd(dadd(a2, b2), 1e-6)), 255.0);   // C cast: truncation
      }
    }
    a.out[n * plane + p] = (uint8_t)v;
  }
}

__global__ void __launch_bounds__(256) k_blend_masks(BlendArgs a) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= a.w || y >= a.h) return;
  blend_pixel(a, x, y);
}

// ---------------------------------------------------------------------------------
// K8 part 1: exact integer sum of V = max(B,G,R) over every dense BGR frame.
// grid = (blocks, n_frames).  48-byte (16-pixel) steps with three 16-byte loads.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ unsigned vsum_16px(const uint4 a, const uint4 b, const uint4 c) {
  const unsigned w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
  unsigned s = 0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {   // 12 bytes = 4 pixels per 3 words
    const unsigned w0 = w[3 * q], w1 = w[3 * q + 1], w2 = w[3 * q + 2];
    // pixel bytes: (w0.0,w0.1,w0.2) (w0.3,w1.0,w1.1) (w1.2,w1.3,w2.0) (w2.1,w2.2,w2.3)
    s += max(max(w0 & 255u, (w0 >> 8) & 255u), (w0 >> 16) & 255u);
    s += max(max(w0 >> 24, w1 & 255u), (w1 >> 8) & 255u);
    s += max(max((w1 >> 16) & 255u, w1 >> 24), w2 & 255u);
    s += max(max((w2 >> 8) & 255u, (w2 >> 16) & 255u), w2 >> 24);
  }
  return s;
}

__global__ void __launch_bounds__(256) k_vsum(const uint8_t* const* __restrict__ frames, long long frame_bytes,
                                              unsigned long long* __restrict__ vsum) {
  const uint8_t* f = frames[blockIdx.y];
  const long long n48 = frame_bytes / 48;
  unsigned long long acc = 0;
  if ((reinterpret_cast<uintptr_t>(f) & 15) == 0) {
    const uint4* f4 = reinterpret_cast<const uint4*>(f);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n48; i += (long long)gridDim.x * blockDim.x) {
      const uint4 a = __ldg(f4 + 3 * i), b = __ldg(f4 + 3 * i + 1), c = __ldg(f4 + 3 * i + 2);
      acc += vsum_16px(a, b, c);
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n48 * 16; i += (long long)gridDim.x * blockDim.x) {
      const uint8_t* q = f + 3 * i;
      acc += max(max(q[0], q[1]), q[2]);
    }
  }
  if (blockIdx.x == 0)   // tail pixels (frame_bytes % 48)
    for (long long i = n48 * 16 + threadIdx.x; i * 3 < frame_bytes; i += blockDim.x) {
      const uint8_t* q = f + 3 * i;
      acc += max(max(q[0], q[1]), q[2]);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ unsigned long long part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int k = 0; k < 8; ++k) t += part[k];
    atomicAdd(vsum + blockIdx.y, t);
  }
}

// K8 part 2: delta_c = cvRound(V_mean - V_c) per frame-set (cv2.add rounds the scalar).
// luminance offsets of one frame-set from its exact V sums (surroundBEV.py:66-74); host-capable for the CPU tests
__host__ __device__ __forceinline__ void lum_deltas(const unsigned long long* vsum, int n_cam, double npix, int* delta) {
  double tot = 0.0;
  for (int c = 0; c < n_cam; ++c) tot = dadd(tot, ddiv((double)vsum[c], npix));
  const double vmean = ddiv(tot, (double)n_cam);
  for (int c = 0; c < n_cam; ++c) delta[c] = cv_round(dadd(vmean, -ddiv((double)vsum[c], npix)));
}

__global__ void k_delta(const unsigned long long* __restrict__ vsum, int n_cam, int batch, double npix,
                        int* __restrict__ delta) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  lum_deltas(vsum + b * n_cam, n_cam, npix, delta + b * n_cam);
}

// ---------------------------------------------------------------------------------
// K9: color_balance (surroundBEV.py:43-55) + car overlay.  Gains from the channel sums;
// out = sat(cvRound(double(px) * K_c)) through a 3x256 table built per CTA in shared memory.
// ---------------------------------------------------------------------------------
// grey-world gains of one canvas from its channel sums, and one entry of the 3 x 256 output table (host-capable)
__host__ __device__ __forceinline__ void gray_world_gains(const unsigned long long* csum, double npix, double* gain) {
  const double B = ddiv((double)csum[0], npix), G = ddiv((double)csum[1], npix), R = ddiv((double)csum[2], npix);
  const double K = ddiv(dadd(dadd(R, G), B), 3.0);
  gain[0] = ddiv(K, B); gain[1] = ddiv(K, G); gain[2] = ddiv(K, R);
}
__host__ __device__ __forceinline__ uint8_t gain_entry(double gain, int v) {
  const int r = cv_round(dmul((double)v, gain));   // non-finite -> INT_MIN -> saturates to 0, as on x86
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

__global__ void __launch_bounds__(256) k_gain(uint8_t* __restrict__ canvas, long long canvas_bytes, double npix,
                                              const unsigned long long* __restrict__ csum,
                                              const uint8_t* __restrict__ car) {
  __shared__ uint8_t tab[3][256];
  const int b = blockIdx.y;
  {
    double gain[3];
    gray_world_gains(csum + b * 3, npix, gain);
    for (int i = threadIdx.x; i < 768; i += blockDim.x) tab[i >> 8][i & 255] = gain_entry(gain[i >> 8], i & 255);
  }
  __syncthreads();
  uint8_t* cv = canvas + (size_t)b * canvas_bytes;
  // 12-byte (4-pixel) steps keep the channel phase fixed per byte lane
  const long long n12 = canvas_bytes / 12;
  const bool aligned = (canvas_bytes % 4 == 0) && ((reinterpret_cast<uintptr_t>(canvas) & 3) == 0) &&
                       (!car || (reinterpret_cast<uintptr_t>(car) & 3) == 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n12; i += (long long)gridDim.x * blockDim.x) {
    if (aligned) {
      unsigned* p = reinterpret_cast<unsigned*>(cv + i * 12);
      unsigned w[3] = {p[0], p[1], p[2]};
      unsigned o[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        unsigned r = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int byte_idx = k * 4 + j;
          r |= (unsigned)tab[byte_idx % 3][(w[k] >> (8 * j)) & 255u] << (8 * j);
        }
        o[k] = r;
      }
      if (car) {
        const unsigned* c = reinterpret_cast<const unsigned*>(car + i * 12);
        o[0] = __vaddus4(o[0], __ldg(c)); o[1] = __vaddus4(o[1], __ldg(c + 1)); o[2] = __vaddus4(o[2], __ldg(c + 2));
      }
      p[0] = o[0]; p[1] = o[1]; p[2] = o[2];
    } else {
      for (int j = 0; j < 12; ++j) {
        int v = tab[j % 3][cv[i * 12 + j]];
        if (car) v = min(255, v + car[i * 12 + j]);
        cv[i * 12 + j] = (uint8_t)v;
      }
    }
  }
  if (blockIdx.x == 0)
    for (long long i = n12 * 12 + threadIdx.x; i < canvas_bytes; i += blockDim.x) {
      int v = tab[i % 3][cv[i]];
      if (car) v = min(255, v + car[i]);
      cv[i] = (uint8_t)v;
    }
}

// ---------------------------------------------------------------------------------
// Multi-GPU compose: saturating sum of n partial canvases (+ car).  16-byte vectors.
// ---------------------------------------------------------------------------------
struct SatSumArgs { const uint8_t* parts[BEVK_MAX_CAMERAS_K]; int n; unsigned long long bytes; const uint8_t* car; uint8_t* out; };

__global__ void __launch_bounds__(256) k_sat_sum(SatSumArgs a) {
  const unsigned long long n16 = a.bytes / 16;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    uint4 s = __ldg(reinterpret_cast<const uint4*>(a.parts[0]) + i);
    for (int k = 1; k < a.n; ++k) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.parts[k]) + i);
      s.x = __vaddus4(s.x, v.x); s.y = __vaddus4(s.y, v.y); s.z = __vaddus4(s.z, v.z); s.w = __vaddus4(s.w, v.w);
    }
    if (a.car) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.car) + i);
      s.x = __vaddus4(s.x, v.x); s.y = __vaddus4(s.y, v.y); s.z = __vaddus4(s.z, v.z); s.w = __vaddus4(s.w, v.w);
    }
    reinterpret_cast<uint4*>(a.out)[i] = s;
  }
  if (blockIdx.x == 0)
    for (unsigned long long i = n16 * 16 + threadIdx.x; i < a.bytes; i += blockDim.x) {
      int s = 0;
      for (int k = 0; k < a.n; ++k) s = min(255, s + a.parts[k][i]);
      if (a.car) s = min(255, s + a.car[i]);
      a.out[i] = (uint8_t)s;
    }
}

// ---------------------------------------------------------------------------------
// Stand-alone forms of K5/K6 (Mask / BlendMask.__call__), K8 (luminance_balance on whole
// frames) and the channel sums of K9, for callers that use those reference functions
// outside BevGenerator.  Dense BGR images.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_apply_mask(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask,
                                                    uint8_t* __restrict__ out, long long npx, int blend) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
    const unsigned w = mask[i];
    int b = img[3 * i], g = img[3 * i + 1], r = img[3 * i + 2];
    if (!blend) { if (!w) b = g = r = 0; }
    else {
      const float wf = __double2float_rn(__ddiv_rn((double)w, 255.0));
      b = (int)__fmul_rn((float)b, wf); g = (int)__fmul_rn((float)g, wf); r = (int)__fmul_rn((float)r, wf);
    }
    out[3 * i] = (uint8_t)b; out[3 * i + 1] = (uint8_t)g; out[3 * i + 2] = (uint8_t)r;
  }
}

__global__ void __launch_bounds__(256) k_chan_sum(const uint8_t* __restrict__ img, long long npx,
                                                  unsigned long long* __restrict__ csum) {
  unsigned long long sb = 0, sg = 0, sr = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
    sb += img[3 * i]; sg += img[3 * i + 1]; sr += img[3 * i + 2];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(csum, sb); atomicAdd(csum + 1, sg); atomicAdd(csum + 2, sr); }
}

// grid.y = frame index; delta[frame] from k_delta
__global__ void __launch_bounds__(256) k_lum_apply(const uint8_t* const* __restrict__ frames, uint8_t* const* __restrict__ outs,
                                                   int w, int h, const int* __restrict__ delta,
                                                   const int* __restrict__ hsv_tab) {
  __shared__ int s_tab[512];
  for (int i = threadIdx.x; i < 512; i += 256) s_tab[i] = hsv_tab[i];
  __syncthreads();
  const uint8_t* f = frames[blockIdx.y];
  uint8_t* o = outs[blockIdx.y];
  const int d = delta[blockIdx.y], tail = w - (w % 32);
  const long long npx = (long long)w * h;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long long)gridDim.x * blockDim.x) {
    int b = f[3 * i], g = f[3 * i + 1], r = f[3 * i + 2];
    hsv_roundtrip(b, g, r, d, (int)(i % w) >= tail, s_tab, s_tab + 256);
    o[3 * i] = (uint8_t)b; o[3 * i + 1] = (uint8_t)g; o[3 * i + 2] = (uint8_t)r;
  }
}

// ---------------------------------------------------------------------------------
// K8 part 3 for the fused BEV path: luminance_balance applied ONCE to every source pixel the LUT
// can sample.  spans[cam * FH + y] = (first, last+1) source column of row y that any tap of camera
// `cam` touches (bevk_bev_finalize); everything else in the frame is never read by k_bev, so it
// is not converted (about 17 % of a frame at the fixture geometry, 4.6x fewer HSV round trips
// than converting the four taps of every output pixel).  grid = (FH, n_frames).
// ---------------------------------------------------------------------------------
constexpr int LUM_ROWS = 16;   // source rows per CTA

// One CTA converts the sampled spans of LUM_ROWS consecutive source rows of one frame.  The work items -- groups of 4
// pixels = 3 aligned words in, 3 out -- of all its rows form ONE flat list (prefix sums of the rows' group counts in shared
// memory) that the threads stride through, so that a row with a short span costs nothing and every thread has several
// independent groups in flight.  Measured (cfg3, 32 x 4 frames at 1080p, profiles/r02_launches_summary.txt): a CTA per row
// 0.235 ms, this form 0.235 ms, with the branch-free sector selection of hsv_roundtrip 0.211 ms; staging 32-group chunks
// through shared memory for coalesced loads and stores was slower again (0.232 ms).
__global__ void __launch_bounds__(128) k_lum_spans(const uint8_t* const* __restrict__ frames, uint8_t* const* __restrict__ outs,
                                                   const int2* __restrict__ spans, int n_cam, int w, int h,
                                                   const int* __restrict__ delta, const int* __restrict__ hsv_tab) {
  const int f = blockIdx.y, y0 = blockIdx.x * LUM_ROWS, y1 = min(h, y0 + LUM_ROWS), nrows = y1 - y0;
  const int2* sp_cam = spans + (size_t)(f % n_cam) * h;
  __shared__ int s_tab[512];
  __shared__ int s_pref[LUM_ROWS + 1], s_g0[LUM_ROWS];
  const bool words = (w & 3) == 0 && ((reinterpret_cast<uintptr_t>(frames[f]) | reinterpret_cast<uintptr_t>(outs[f])) & 3) == 0;
  if (threadIdx.x == 0) {
    // groups cover each span rounded out to multiples of 4 pixels: the extra pixels are converted too, which nobody samples
    int acc = 0;
    for (int r = 0; r < nrows; ++r) {
      const int2 sp = sp_cam[y0 + r];
      const int g0 = sp.x >> 2, g1 = sp.y > sp.x ? (sp.y + 3) >> 2 : g0;
      s_pref[r] = acc; s_g0[r] = g0;
      acc += g1 - g0;
    }
    for (int r = nrows; r <= LUM_ROWS; ++r) s_pref[r] = acc;
  }
  for (int i = threadIdx.x; i < 512; i += 128) s_tab[i] = hsv_tab[i];
  __syncthreads();
  const int total = s_pref[LUM_ROWS];
  if (total == 0) return;
  const int d = delta[f], tail = w - (w % 32);
  if (words) {
    const size_t row_words = (size_t)w * 3 / 4;
    const unsigned* src0 = reinterpret_cast<const unsigned*>(frames[f]) + (size_t)y0 * row_words;
    unsigned* dst0 = reinterpret_cast<unsigned*>(outs[f]) + (size_t)y0 * row_words;
    int r = 0;
#pragma unroll 2
    for (int i = threadIdx.x; i < total; i += 128) {
      while (i >= s_pref[r + 1]) ++r;                       // the thread's items are visited in increasing order
      const int gidx = s_g0[r] + (i - s_pref[r]);
      const unsigned* q = src0 + (size_t)r * row_words + 3 * gidx;
      const unsigned w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);       // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
      int c[12];
#pragma unroll
      for (int k = 0; k < 4; ++k) { c[k] = (w0 >> (8 * k)) & 255; c[4 + k] = (w1 >> (8 * k)) & 255; c[8 + k] = (w2 >> (8 * k)) & 255; }
      const bool rt = 4 * gidx >= tail;      // tail is a multiple of 4 here: a group is entirely body or entirely row tail
#pragma unroll
      for (int px = 0; px < 4; ++px) hsv_roundtrip(c[3 * px], c[3 * px + 1], c[3 * px + 2], d, rt, s_tab, s_tab + 256);
      unsigned* o = dst0 + (size_t)r * row_words + 3 * gidx;
      o[0] = (unsigned)c[0] | ((unsigned)c[1] << 8) | ((unsigned)c[2] << 16) | ((unsigned)c[3] << 24);
      o[1] = (unsigned)c[4] | ((unsigned)c[5] << 8) | ((unsigned)c[6] << 16) | ((unsigned)c[7] << 24);
      o[2] = (unsigned)c[8] | ((unsigned)c[9] << 8) | ((unsigned)c[10] << 16) | ((unsigned)c[11] << 24);
    }
  } else {
    for (int y = y0; y < y1; ++y) {
      const int2 sp = sp_cam[y];
      const uint8_t* src = frames[f] + (size_t)y * w * 3;
      uint8_t* dst = outs[f] + (size_t)y * w * 3;
      for (int x = sp.x + threadIdx.x; x < sp.y; x += 128) {
        int b = src[3 * x], g = src[3 * x + 1], r = src[3 * x + 2];
        hsv_roundtrip(b, g, r, d, x >= tail, s_tab, s_tab + 256);
        dst[3 * x] = (uint8_t)b; dst[3 * x + 1] = (uint8_t)g; dst[3 * x + 2] = (uint8_t)r;
      }
    }
  }
}

// ---------------------------------------------------------------------------------
// Host -> device ingest for page-locked (mapped) host frames: instead of DMA rectangles, the SMs
// read exactly the sampled row spans of every frame straight out of host memory (zero-copy, 16-byte
// vectors, coalesced) and write them into the device frame buffers.  Moves ~17 % of each frame
// over PCIe instead of the 23-34 % a band / bounding-box DMA needs.  grid = (FH, n_frames).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_fetch_spans(const uint8_t* const* __restrict__ host_frames, uint8_t* const* __restrict__ dev_frames,
                                                     const int2* __restrict__ spans, int n_cam, int h, long long host_stride,
                                                     int row_bytes) {
  const int y = blockIdx.x, f = blockIdx.y;
  const int2 sp = spans[(f % n_cam) * h + y];
  if (sp.y <= sp.x) return;
  // 16-byte window around the span (+4 px each side for the aligned word reads of the gather)
  const int b0 = max(0, 3 * sp.x - 12) & ~15, b1 = min(row_bytes, (3 * sp.y + 12 + 15) & ~15);
  const uint4* src = reinterpret_cast<const uint4*>(host_frames[f] + (size_t)y * host_stride + b0);
  uint4* dst = reinterpret_cast<uint4*>(dev_frames[f] + (size_t)y * row_bytes + b0);
  for (int i = threadIdx.x; i < (b1 - b0) >> 4; i += 128) dst[i] = src[i];
}

}  // namespace bevk
