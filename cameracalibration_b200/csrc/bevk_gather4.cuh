// bevk_gather4.cuh -- 4-pixels-per-thread form of the stand-alone gathers for 3-channel INTER_LINEAR:
// cv2.remap with resident maps (MODE 0; Camera.undistort / InCalibrator.undistort / Tools/undistort.py,
// surroundBEV.py:110-111, intrinsicCalib.py:193-195, undistort.py:66), the same with the camera model
// evaluated in-kernel (MODE 1) and cv2.warpPerspective (MODE 2; extrinsicCalib.py:166-169).
// Same tap machinery as the fused BEV kernel (aligned 32-bit words, funnel shift, PRMT, DP2A);
// each thread produces 12 output bytes and stores them as three 32-bit words.  The generic
// k_gather stays for 1/4-channel images, INTER_NEAREST and sizes that are not multiples of 4.
#pragma once
#include "bevk_bev.cuh"
#include "bevk_kernels.cuh"

namespace bevk {

// one output pixel: fixed-point source position -> packed B | G<<8 | R<<16
__host__ __device__ __forceinline__ unsigned gather_px(const uint8_t* __restrict__ src, unsigned spitch, int sw, int sh, int sx, int sy,
                                              unsigned fx, unsigned fy) {
  const bool inside = sx >= 0 && sy >= 0 && sx + 1 < sw && sy + 1 < sh;
  if (inside) {
    const unsigned off = (unsigned)sy * spitch + 3u * (unsigned)sx;
    const unsigned off_al = off & ~3u, sh8 = (off & 3u) * 8u;
    const bool third = (sh8 == 24u);
    const uint8_t* q0 = src + off_al;
    const uint8_t* q1 = q0 + spitch;
    const unsigned a0 = ldg32(q0), a1 = ldg32(q0 + 4), a2 = third ? ldg32(q0 + 8) : 0u;
    const unsigned b0 = ldg32(q1), b1 = ldg32(q1 + 4), b2 = third ? ldg32(q1 + 8) : 0u;
    const unsigned w11 = fx * fy, w01 = (fx << 5) - w11, w10 = (fy << 5) - w11, w00 = 1024u - (fx << 5) - (fy << 5) + w11;
    return interp_fast(sh8, w00 | (w01 << 16), w10 | (w11 << 16), 65536u, a0, a1, a2, b0, b1, b2);
  }
  int p[4][3];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int tx = sx + (t & 1), ty = sy + (t >> 1);
    if ((unsigned)tx < (unsigned)sw && (unsigned)ty < (unsigned)sh) {
      const uint8_t* q = src + (size_t)ty * spitch + 3 * tx;
      p[t][0] = ldg8(q); p[t][1] = ldg8(q + 1); p[t][2] = ldg8(q + 2);
    } else { p[t][0] = p[t][1] = p[t][2] = 0; }
  }
  const unsigned ob = (unsigned)bilerp_q10(p[0][0], p[1][0], p[2][0], p[3][0], (int)fx, (int)fy);
  const unsigned og = (unsigned)bilerp_q10(p[0][1], p[1][1], p[2][1], p[3][1], (int)fx, (int)fy);
  const unsigned orr = (unsigned)bilerp_q10(p[0][2], p[1][2], p[2][2], p[3][2], (int)fx, (int)fy);
  return ob | (og << 8) | (orr << 16);
}

// Requirements checked by the host: channels == 3, INTER_LINEAR, dw % 4 == 0, dense dst (pitch 3*dw),
// source pitch % 4 == 0 with 4 bytes of readable slack after the frame (library-owned buffers).
template <int MODE>
__global__ void __launch_bounds__(256) k_gather4(GatherArgs a) {
  const int x4 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x4 >= a.dw || y >= a.dh) return;
  unsigned px[4];
  short mx[4], my[4];
  unsigned short fr[4];
  if (MODE == 0) {
    const size_t i = (size_t)y * a.dw + x4;
    const int4 m = *reinterpret_cast<const int4*>(a.map1 + i);        // 4 x (short x, short y)
    const uint2 f = *reinterpret_cast<const uint2*>(a.map2 + i);      // 4 x uint16
    mx[0] = (short)(m.x & 0xffff); my[0] = (short)(m.x >> 16);
    mx[1] = (short)(m.y & 0xffff); my[1] = (short)(m.y >> 16);
    mx[2] = (short)(m.z & 0xffff); my[2] = (short)(m.z >> 16);
    mx[3] = (short)(m.w & 0xffff); my[3] = (short)(m.w >> 16);
    fr[0] = (unsigned short)(f.x & 0xffffu); fr[1] = (unsigned short)(f.x >> 16);
    fr[2] = (unsigned short)(f.y & 0xffffu); fr[3] = (unsigned short)(f.y >> 16);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int sx, sy;
    unsigned fx, fy;
    if (MODE == 2) {
      int X, Y;
      warp_point(a.hm, x4 + q, y, (double)TAB, X, Y);
      sx = sat_i16(X >> INTER_BITS); sy = sat_i16(Y >> INTER_BITS);
      fx = X & (TAB - 1); fy = Y & (TAB - 1);
    } else {
      if (MODE == 1) {
        double u, v;
        undistort_point(a.cm, x4 + q, y, u, v);
        quantise_uv(u, v, mx[q], my[q], fr[q], pack_saturates(a.cm.model, x4 + q, a.cm.w));
      }
      sx = mx[q]; sy = my[q];
      fx = fr[q] & (TAB - 1); fy = (fr[q] >> INTER_BITS) & (TAB - 1);
    }
    px[q] = gather_px(a.src, (unsigned)a.spitch, a.sw, a.sh, sx, sy, fx, fy);
  }
  unsigned* o = reinterpret_cast<unsigned*>(a.dst + (size_t)y * a.dpitch + (size_t)x4 * 3);
  o[0] = __byte_perm(px[0], px[1], 0x4210);   // B0 G0 R0 B1
  o[1] = __byte_perm(px[1], px[2], 0x5421);   // G1 R1 B2 G2
  o[2] = __byte_perm(px[2], px[3], 0x6542);   // R2 B3 G3 R3
}

}  // namespace bevk
