// bevk_device.cuh -- device-side arithmetic shared by the bevk kernels (sm_100a).
//
// Everything here is a bit-exact restatement target: the coordinate math is IEEE
// FP64 with NO fused multiply-add (OpenCV's x86 baseline build has none), so every
// product/sum goes through __dmul_rn/__dadd_rn, which the compiler may not contract.
// Specs: SURVEY.md Appendix A1 (fisheye map), A11 (pinhole map), A2 (fixed-point
// bilinear), A3 (warpPerspective coordinates), A9 (8-bit HSV round trip).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bevk {

constexpr int INTER_BITS = 5;
constexpr int TAB = 32;

struct CamModel {      // cv2.fisheye / cv2 initUndistortRectifyMap inputs, pre-digested on the host
  double iR[9];        // inv(P * R), R = I
  double k[5];         // fisheye: k1..k4 ; pinhole: k1,k2,p1,p2,k3
  double fx, fy, cx, cy;
  int model;           // BEVK_MODEL_*
  int w, h;            // size of the undistorted (destination) frame
};

struct Homog { double M[9]; };   // inv(H), as cv2.warpPerspective computes it

// OpenCV's closed-form 3x3 inverse (cv::invert, DECOMP_LU, n == 3, CV_64F).
inline bool inv3(const double* S, double* T) {
#define M(r, c) S[(r) * 3 + (c)]
  double d = M(0, 0) * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) - M(0, 1) * (M(1, 0) * M(2, 2) - M(1, 2) * M(2, 0)) +
             M(0, 2) * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
  if (d == 0.) return false;
  d = 1. / d;
  T[0] = (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) * d;
  T[1] = (M(0, 2) * M(2, 1) - M(0, 1) * M(2, 2)) * d;
  T[2] = (M(0, 1) * M(1, 2) - M(0, 2) * M(1, 1)) * d;
  T[3] = (M(1, 2) * M(2, 0) - M(1, 0) * M(2, 2)) * d;
  T[4] = (M(0, 0) * M(2, 2) - M(0, 2) * M(2, 0)) * d;
  T[5] = (M(0, 2) * M(1, 0) - M(0, 0) * M(1, 2)) * d;
  T[6] = (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0)) * d;
  T[7] = (M(0, 1) * M(2, 0) - M(0, 0) * M(2, 1)) * d;
  T[8] = (M(0, 0) * M(1, 1) - M(0, 1) * M(1, 0)) * d;
#undef M
  return true;
}

// does undistorted pixel column j take the saturating (vector-body) pack?  (pinhole model only)
__host__ __device__ __forceinline__ bool pack_saturates(int model, int j, int w) { return model == 1 && j < w - (w % 8); }

// Separately rounded FP64 operations (no FMA contraction), as OpenCV's scalar C++ evaluates them.  The host
// forms let tests/host/kernel_math.cu run the very same coordinate code on a CPU (built without contraction).
#ifdef __CUDA_ARCH__
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ int d2i_rn(double v) { return __double2int_rn(v); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ int f2i_rn(float v) { return __float2int_rn(v); }
#else
inline double dmul(double a, double b) { volatile double r = a * b; return r; }
inline double dadd(double a, double b) { volatile double r = a + b; return r; }
inline double ddiv(double a, double b) { volatile double r = a / b; return r; }
inline double dsqrt(double a) { return sqrt(a); }
inline double dinf() { return HUGE_VAL; }
inline int d2i_rn(double v) { return (int)nearbyint(v); }   // default rounding mode: to nearest even
inline float fmul(float a, float b) { volatile float r = a * b; return r; }
inline float fsub(float a, float b) { volatile float r = a - b; return r; }
inline float fadd(float a, float b) { volatile float r = a + b; return r; }
inline float ffma(float a, float b, float c) { return fmaf(a, b, c); }
inline int f2i_rn(float v) { return (int)nearbyintf(v); }
#endif

// cvRound / saturate_cast<int>(double): x86 cvtsd2si returns INT_MIN ("integer
// indefinite") for NaN and out-of-range inputs.
__host__ __device__ __forceinline__ int cv_round(double v) {
  if (!(fabs(v) < 2147483648.0)) return INT_MIN;
  return d2i_rn(v);
}

// A1 / A11: source-image position (u,v) of undistorted pixel (j,i).
__host__ __device__ __forceinline__ void undistort_point(const CamModel& c, int j, int i, double& u, double& v) {
  const double dj = (double)j, di = (double)i;
  const double _x = dadd(dmul(dj, c.iR[0]), dadd(dmul(di, c.iR[1]), c.iR[2]));
  const double _y = dadd(dmul(dj, c.iR[3]), dadd(dmul(di, c.iR[4]), c.iR[5]));
  const double _w = dadd(dmul(dj, c.iR[6]), dadd(dmul(di, c.iR[7]), c.iR[8]));
  if (c.model == 0) {  // equidistant fisheye
    if (_w <= 0) {
      const double inf = dinf();
      u = (_x > 0) ? -inf : inf;
      v = (_y > 0) ? -inf : inf;
      return;
    }
    const double x = ddiv(_x, _w), y = ddiv(_y, _w);
    const double r = dsqrt(dadd(dmul(x, x), dmul(y, y)));
    const double th = atan(r);
    const double t2 = dmul(th, th), t4 = dmul(t2, t2), t6 = dmul(t4, t2), t8 = dmul(t4, t4);
    const double poly = dadd(dadd(dadd(dadd(1.0, dmul(c.k[0], t2)), dmul(c.k[1], t4)), dmul(c.k[2], t6)), dmul(c.k[3], t8));
    const double thd = dmul(th, poly);
    const double s = (r == 0) ? 1.0 : ddiv(thd, r);
    u = dadd(dmul(dmul(c.fx, x), s), c.cx);
    v = dadd(dmul(dmul(c.fy, y), s), c.cy);
  } else {             // pinhole, k1 k2 p1 p2 k3
    const double w = ddiv(1.0, _w), x = dmul(_x, w), y = dmul(_y, w);
    const double x2 = dmul(x, x), y2 = dmul(y, y);
    const double r2 = dadd(x2, y2), _2xy = dmul(dmul(2.0, x), y);
    const double k1 = c.k[0], k2 = c.k[1], p1 = c.k[2], p2 = c.k[3], k3 = c.k[4];
    const double kr = dadd(1.0, dmul(dadd(dmul(dadd(dmul(k3, r2), k2), r2), k1), r2));
    const double xd = dadd(dadd(dmul(x, kr), dmul(p1, _2xy)), dmul(p2, dadd(r2, dmul(2.0, x2))));
    const double yd = dadd(dadd(dmul(y, kr), dmul(p1, dadd(r2, dmul(2.0, y2)))), dmul(p2, _2xy));
    u = dadd(dmul(c.fx, xd), c.cx);
    v = dadd(dmul(c.fy, yd), c.cy);
  }
}

// CV_16SC2 + CV_16UC1 quantisation of (u,v): map1 = (iu>>5, iv>>5) as int16 (wrapping
// cast), map2 = (iv&31)*32 + (iu&31).
// `saturate`: cv2.initUndistortRectifyMap's (pinhole) vector body packs with signed
// saturation for columns j < W - W%8; everything else wraps like the C cast it is.
__host__ __device__ __forceinline__ void quantise_uv(double u, double v, short& mx, short& my, unsigned short& frac,
                                            bool saturate = false) {
  const int iu = cv_round(dmul(u, (double)TAB));
  const int iv = cv_round(dmul(v, (double)TAB));
  const int hx = iu >> INTER_BITS, hy = iv >> INTER_BITS;
  mx = saturate ? (short)max(-32768, min(32767, hx)) : (short)hx;
  my = saturate ? (short)max(-32768, min(32767, hy)) : (short)hy;
  frac = (unsigned short)((iv & (TAB - 1)) * TAB + (iu & (TAB - 1)));
}

// A3: fixed-point pre-image of destination pixel (x,y) under cv2.warpPerspective.
// unit = 32 for INTER_LINEAR, 1 for INTER_NEAREST.  OpenCV evaluates in 64-pixel blocks.
__host__ __device__ __forceinline__ void warp_point(const Homog& hm, int x, int y, double unit, int& X, int& Y) {
  const double* M = hm.M;
  const int bxi = (x >> 6) << 6;
  const double bx = (double)bxi, x1 = (double)(x - bxi), dy = (double)y;
  const double X0 = dadd(dadd(dmul(M[0], bx), dmul(M[1], dy)), M[2]);
  const double Y0 = dadd(dadd(dmul(M[3], bx), dmul(M[4], dy)), M[5]);
  const double W0 = dadd(dadd(dmul(M[6], bx), dmul(M[7], dy)), M[8]);
  double W = dadd(W0, dmul(M[6], x1));
  W = (W != 0.0) ? ddiv(unit, W) : 0.0;
  double fX = dmul(dadd(X0, dmul(M[0], x1)), W);
  double fY = dmul(dadd(Y0, dmul(M[3], x1)), W);
  fX = fmax(-2147483648.0, fmin(2147483647.0, fX));   // std::max(INT_MIN, std::min(INT_MAX, .))
  fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
  X = cv_round(fX);
  Y = cv_round(fY);
}

__host__ __device__ __forceinline__ int sat_i16(int v) { return max(-32768, min(32767, v)); }

// ---- byte-lane primitives with a host form ------------------------------------
// The packed integer arithmetic of the gathers (interp_fast, sat_add_bgr, the tile write-out) is built from
// these; on the device they are single SASS instructions (PRMT, SHF, IDP.2A, VADDUS4-style), on the host plain
// C, so tests/host/kernel_math.cu can check the packed forms against the scalar definitions without a GPU.
__host__ __device__ __forceinline__ unsigned lane_perm(unsigned a, unsigned b, unsigned sel) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, sel);
#else
  const unsigned long long v = ((unsigned long long)b << 32) | a;   // selector nibbles 0..7 only (no sign replication here)
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 255u) << (8 * i);
  return r;
#endif
}
__host__ __device__ __forceinline__ unsigned lane_funnel_r(unsigned lo, unsigned hi, unsigned shift) {
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, shift);
#else
  shift &= 31u;
  return shift ? (lo >> shift) | (hi << (32u - shift)) : lo;
#endif
}
// two unsigned 16-bit weights (a) times the low / high byte pair of b, plus c
__host__ __device__ __forceinline__ unsigned lane_dp2a_lo(unsigned a, unsigned b, unsigned c) {
#ifdef __CUDA_ARCH__
  return __dp2a_lo(a, b, c);
#else
  return c + (a & 0xffffu) * (b & 255u) + (a >> 16) * ((b >> 8) & 255u);
#endif
}
__host__ __device__ __forceinline__ unsigned lane_dp2a_hi(unsigned a, unsigned b, unsigned c) {
#ifdef __CUDA_ARCH__
  return __dp2a_hi(a, b, c);
#else
  return c + (a & 0xffffu) * ((b >> 16) & 255u) + (a >> 16) * (b >> 24);
#endif
}
__host__ __device__ __forceinline__ unsigned lane_addus4(unsigned a, unsigned b) {   // per-byte saturating add
#ifdef __CUDA_ARCH__
  return __vaddus4(a, b);
#else
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned t = ((a >> (8 * i)) & 255u) + ((b >> (8 * i)) & 255u);
    r |= (t > 255u ? 255u : t) << (8 * i);
  }
  return r;
#endif
}

// A2: (sum w*p + 512) >> 10 with integer weights.
__host__ __device__ __forceinline__ int bilerp_q10(int p00, int p01, int p10, int p11, int fx, int fy) {
  const int w11 = fx * fy, w01 = (fx << 5) - w11, w10 = (fy << 5) - w11;
  const int w00 = 1024 - (fx << 5) - (fy << 5) + w11;
  return (w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11 + 512) >> 10;
}

// ---- A9: OpenCV's 8-bit BGR -> HSV -> (V + delta) -> BGR round trip ------------
// sdiv[i] = cvRound((255<<12)/i), hdiv[i] = cvRound((180<<12)/(6 i)); filled on the host.
struct HsvTables { int sdiv[256]; int hdiv[256]; };

__host__ __device__ __forceinline__ void hsv_roundtrip(int& b, int& g, int& r, int delta, bool rounding_tail,
                                              const int* __restrict__ sdiv, const int* __restrict__ hdiv) {
  const int v = max(b, max(g, r)), mn = min(b, min(g, r));
  const int d = v - mn;
  const int s = (d * sdiv[v] + 2048) >> 12;
  int hh = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * d) : (r - g + 4 * d));
  hh = (hh * hdiv[d] + 2048) >> 12;
  if (hh < 0) hh += 180;
  const int v2 = max(0, min(255, v + delta));
  // HSV2BGR, OpenCV's vector body: fp32 with these exact FMA contractions, truncation.
  const float sf = fmul((float)s, 1.0f / 255.0f);
  const float vf = fmul((float)v2, 1.0f / 255.0f);
  const float hx = fmul((float)hh, 6.0f / 180.0f);
  const float secf = truncf(hx);
  const float f = fsub(hx, secf);
  int sec = (int)secf;
  sec = sec % 6;
  const float t0 = vf;
  const float t1 = fmul(vf, fsub(1.0f, sf));
  const float t2 = fmul(vf, ffma(-sf, f, 1.0f));
  const float t3 = fmul(vf, ffma(-sf, fsub(1.0f, f), 1.0f));
  // sector table {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0} (b, g, r pick t[.]), anything else like sector 5 --
  // as selects, not as a switch: neighbouring pixels lie in different sectors, and a six-way divergent branch per pixel was
  // what k_lum_spans spent its time on.  Two bits per sector and channel:
  const unsigned sc = (unsigned)sec > 5u ? 10u : 2u * (unsigned)sec;
  const unsigned ib = ((1u | 1u << 2 | 3u << 4 | 0u << 6 | 0u << 8 | 2u << 10) >> sc) & 3u;
  const unsigned ig = ((3u | 0u << 2 | 0u << 4 | 2u << 6 | 1u << 8 | 1u << 10) >> sc) & 3u;
  const unsigned ir = ((0u | 2u << 2 | 1u << 4 | 1u << 6 | 3u << 8 | 0u << 10) >> sc) & 3u;
  float fb = (ib & 2u) ? ((ib & 1u) ? t3 : t2) : ((ib & 1u) ? t1 : t0);
  float fg = (ig & 2u) ? ((ig & 1u) ? t3 : t2) : ((ig & 1u) ? t1 : t0);
  float fr = (ir & 2u) ? ((ir & 1u) ? t3 : t2) : ((ir & 1u) ? t1 : t0);
  fb = fmul(fb, 255.0f); fg = fmul(fg, 255.0f); fr = fmul(fr, 255.0f);
  if (rounding_tail) {   // the < 32-pixel row tail goes through OpenCV's scalar path, which rounds
    b = max(0, min(255, f2i_rn(fb)));
    g = max(0, min(255, f2i_rn(fg)));
    r = max(0, min(255, f2i_rn(fr)));
  } else {
    b = max(0, min(255, (int)fb));
    g = max(0, min(255, (int)fg));
    r = max(0, min(255, (int)fr));
  }
}

}  // namespace bevk
