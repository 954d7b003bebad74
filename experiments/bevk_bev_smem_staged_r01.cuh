// bevk_bev.cuh -- the fused per-frame surround-BEV kernel (sm_100a).
//
// Reference path fused here (SurroundBirdEyeView/surroundBEV.py:312-325): for every
// camera  raw2bev = cv2.remap(img, bev_map1, bev_map2, INTER_LINEAR)  (:116-117), then
// Mask / BlendMask.__call__ (:161-162 / :279-280), then the saturating cv2.add chain
// (:318-320), the optional car overlay (:323-324), and -- in the BALANCE variant --
// luminance_balance applied lazily to the sampled taps (:57-79) plus the channel sums
// color_balance needs (:44-47).  No intermediate image is written to HBM.
//
// Work decomposition
//   * canvas tiles of 32x32 px; per tile a list of "items" = cameras whose mask touches it;
//   * per item a thread-ordered LUT block of 1024 x 16 B, fully decoded at plan time
//     (bevk_bev_finalize):  .x = byte offset of tap (sy,sx)  (in the staged source box, or in
//     the frame for unstaged items; border entries: sx | sy<<16), .y/.z = the four bilinear
//     weights as 16-bit pairs for DP2A, .w = blend multiplier | frac << 17 | flags << 28.
//     Lanes run along the canvas direction that walks source ROWS (orientation flag);
//   * persistent CTAs loop over (tile, group of NB frame-sets); a shared accumulator tile per
//     frame-set takes the cameras in reference order (saturating add from the 2nd on);
//   * SOURCE STAGING: for every (item, frame-set) the source bounding box of the tile (rows x
//     row_bytes, computed at plan time) is brought into shared memory by the TMA engine --
//     one cp.async.bulk per source row, issued by warp 0, completion on an mbarrier -- through
//     a 3-deep ring that runs two boxes ahead of the math.  The gather then reads shared
//     memory (two aligned 32-bit words per source row, +1 predicated), so the L1 tag/data
//     path only sees the LUT.  Items whose box exceeds the ring slot (far-field tiles with
//     strong minification), frames whose pitch is not a multiple of 16 B, and the BALANCE
//     variant gather straight from global memory instead;
//   * taps -> funnel shift -> PRMT gathers the four taps of one channel into one register ->
//     two DP2A (16-bit weights x 8-bit pixels) give  sum w*p + 512 ;
//   * the blend weight is an exact integer form of the reference's float expression:
//       uint8(float32(px) * float32(mask/255.0)) == (px * (257*mask + 1)) >> 16   for all px, mask
//     in 0..255 (mask 0 -> 0; mask 255 -> identity), checked exhaustively in tests/.
#pragma once
#include "bevk_device.cuh"

namespace bevk {

constexpr int TILE = 32;
constexpr int ACC_WPITCH = TILE + 1;               // 33 words per row: rows and columns are both conflict-free
constexpr int ACC_WORDS = TILE * ACC_WPITCH;       // 1056 words = 4224 B per frame-set
constexpr unsigned LUT_ACTIVE = 1u << 28, LUT_BORDER = 2u << 28;
constexpr int STAGE_SLOTS = 3;                     // ring depth: the producer runs two boxes ahead
#ifndef BEVK_STAGE_CAP
#define BEVK_STAGE_CAP 12288                       // bytes per ring slot (source box of one item, one frame)
#endif
constexpr int STAGE_CAP = BEVK_STAGE_CAP;
constexpr int STAGE_SLOT_BYTES = STAGE_CAP + 128;  // slack: the word loads may run 8 B past the box

struct BevItem {           // 32 B
  int cam, orient;         // orient 0: lanes along canvas x, 1: lanes along canvas y
  int staged;              // 1: LUT offsets are relative to the staged source box
  unsigned src_off;        // byte offset of the box origin in the frame (16-B aligned)
  int rows, row_bytes;     // box: rows x row_bytes (row_bytes % 16 == 0)
  int pad0, pad1;
};

struct BevParams {
  const uint8_t* const* srcs;   // device array [batch * n_cam] of dense BGR frames
  int n_cam, FW, FH;
  unsigned pitch;               // source row pitch in bytes (= 3*FW)
  const int4* tiles;            // x0, y0, first item, item count
  const BevItem* items;
  const uint4* lut;             // [item][4][256]
  int n_tiles, batch;
  uint8_t* out; int BW, BH; long long canvas_bytes;
  const uint8_t* car;
  const int* delta;             // [batch * n_cam] luminance offsets (BALANCE)
  unsigned long long* csum;     // [batch * 3] channel sums of the composed canvas (BALANCE)
  const int* hsv_tab;           // sdiv[256] ++ hdiv[256]
  int cam_lo, cam_hi;
  int tail_start;               // FW - FW % 32: first column of OpenCV's scalar HSV2BGR row tail
  int stage;                    // 0 disables source staging (tuning / A-B measurements)
};

// ---- mbarrier / bulk-copy (TMA) primitives -------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Ampere-style 16-byte async copy (SASS: LDGSTS) + "arrive on the mbarrier when my copies land"
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef BEVK_FILL
#define BEVK_FILL 1   // 1: TMA bulk copy per source row, 2: cp.async 16-B chunks
#endif

__device__ __forceinline__ unsigned ldg32(const uint8_t* p) { return __ldg(reinterpret_cast<const unsigned*>(p)); }

// Fast path, phase 2: the six words of one entry -> weighted pixel packed B | G<<8 | R<<16.
__device__ __forceinline__ unsigned interp_fast(unsigned sh, unsigned wpx, unsigned wpy, unsigned wm, unsigned a0, unsigned a1,
                                                unsigned a2, unsigned b0, unsigned b1, unsigned b2) {
  const unsigned A = __funnelshift_r(a0, a1, sh), A2 = __funnelshift_r(a1, a2, sh);   // B0 G0 R0 B1 | G1 R1 . .
  const unsigned B = __funnelshift_r(b0, b1, sh), B2 = __funnelshift_r(b1, b2, sh);
  // four taps of one channel per register: [p00 p01 p10 p11]
  const unsigned pb = __byte_perm(A, B, 0x7430);
  const unsigned pg = __byte_perm(__byte_perm(A, A2, 0x0041), __byte_perm(B, B2, 0x0041), 0x5410);
  const unsigned pr = __byte_perm(__byte_perm(A, A2, 0x0052), __byte_perm(B, B2, 0x0052), 0x5410);
  const unsigned ob = __dp2a_hi(wpy, pb, __dp2a_lo(wpx, pb, 512u)) >> 10;
  const unsigned og = __dp2a_hi(wpy, pg, __dp2a_lo(wpx, pg, 512u)) >> 10;
  const unsigned orr = __dp2a_hi(wpy, pr, __dp2a_lo(wpx, pr, 512u)) >> 10;
  // BlendMask.__call__ / Mask.__call__ in exact integer form: (v * wm) < 2^24 and the weighted
  // value is its byte 2 -- pack the three byte-2s with two PRMTs
  return __byte_perm(__byte_perm(ob * wm, og * wm, 0x0062), orr * wm, 0x7610);
}

// Slow path (kept out of line so the hot loop stays inside the instruction cache): entries with
// out-of-frame taps (BORDER_CONSTANT 0 per tap; also every entry when the pitch is not a multiple
// of 4) and the BALANCE variant, which runs OpenCV's 8-bit HSV round trip on each of the four taps.
// `src` is the FRAME base; `off` is a frame offset (never a staged-box offset).
struct SlowGeo { unsigned pitch; int FW, FH, tail_start; };
template <bool BAL>
__device__ __noinline__ unsigned sample_slow(const SlowGeo P, const uint8_t* __restrict__ src, unsigned ex, unsigned ew,
                                             int delta, const int* s_hsv) {
  int p[4][3];
  const unsigned wm = ew & 0x1ffffu;
  if (ew & LUT_BORDER) {
    const int sx = (short)(ex & 0xffffu), sy = (short)(ex >> 16);
#pragma unroll 1
    for (int t = 0; t < 4; ++t) {
      const int tx = sx + (t & 1), ty = sy + (t >> 1);
      const bool in = ((unsigned)tx < (unsigned)P.FW) && ((unsigned)ty < (unsigned)P.FH);
      int c0 = 0, c1 = 0, c2 = 0;
      if (in) {
        const uint8_t* q = src + (size_t)ty * P.pitch + 3 * tx;
        c0 = __ldg(q); c1 = __ldg(q + 1); c2 = __ldg(q + 2);
        if (BAL) hsv_roundtrip(c0, c1, c2, delta, tx >= P.tail_start, s_hsv, s_hsv + 256);
      }   // else: the border constant is not an image pixel, no balance
      p[t][0] = c0; p[t][1] = c1; p[t][2] = c2;
    }
  } else {
    const unsigned off_al = ex & ~3u, sh = (ex & 3u) * 8u;
    const uint8_t* q0 = src + off_al;
    const uint8_t* q1 = q0 + P.pitch;
    const bool third = (sh == 24u);
    const unsigned a0 = ldg32(q0), a1 = ldg32(q0 + 4), a2 = third ? ldg32(q0 + 8) : 0u;
    const unsigned b0 = ldg32(q1), b1 = ldg32(q1 + 4), b2 = third ? ldg32(q1 + 8) : 0u;
    const unsigned A = __funnelshift_r(a0, a1, sh), A2 = __funnelshift_r(a1, a2, sh);
    const unsigned B = __funnelshift_r(b0, b1, sh), B2 = __funnelshift_r(b1, b2, sh);
    p[0][0] = A & 255u; p[0][1] = (A >> 8) & 255u; p[0][2] = (A >> 16) & 255u;
    p[1][0] = A >> 24;  p[1][1] = A2 & 255u;       p[1][2] = (A2 >> 8) & 255u;
    p[2][0] = B & 255u; p[2][1] = (B >> 8) & 255u; p[2][2] = (B >> 16) & 255u;
    p[3][0] = B >> 24;  p[3][1] = B2 & 255u;       p[3][2] = (B2 >> 8) & 255u;
    if (BAL) {
      const int sx0 = (P.tail_start != P.FW) ? (int)((ex % P.pitch) / 3u) : 0;
#pragma unroll 1
      for (int t = 0; t < 4; ++t)
        hsv_roundtrip(p[t][0], p[t][1], p[t][2], delta, (sx0 + (t & 1)) >= P.tail_start, s_hsv, s_hsv + 256);
    }
  }
  const int fx = (ew >> 17) & 31, fy = (ew >> 22) & 31;
  unsigned ob = (unsigned)bilerp_q10(p[0][0], p[1][0], p[2][0], p[3][0], fx, fy);
  unsigned og = (unsigned)bilerp_q10(p[0][1], p[1][1], p[2][1], p[3][1], fx, fy);
  unsigned orr = (unsigned)bilerp_q10(p[0][2], p[1][2], p[2][2], p[3][2], fx, fy);
  ob = (ob * wm) >> 16; og = (og * wm) >> 16; orr = (orr * wm) >> 16;
  return ob | (og << 8) | (orr << 16);
}

// cv2.add of two packed BGR pixels: per-byte saturating add (bytes 0..2; byte 3 stays 0)
__device__ __forceinline__ unsigned sat_add_bgr(unsigned a, unsigned b) {
  const unsigned lo = (a & 0x00ff00ffu) + (b & 0x00ff00ffu);          // bytes 0 and 2 -> 9-bit sums in 16-bit lanes
  const unsigned hi = ((a >> 8) & 0xffu) + ((b >> 8) & 0xffu);        // byte 1
  const unsigned lo_s = (lo | (((lo >> 8) & 0x00010001u) * 0xffu)) & 0x00ff00ffu;
  const unsigned hi_s = min(hi, 255u);
  return lo_s | (hi_s << 8);
}

// The walk over this CTA's work: units (tile x frame-set group) -> items -> frame-sets of the
// group.  Consumer (all warps) and producer (warp 0, running ahead) each keep one cursor and
// advance it with the same rules, so they agree on which (item, frame-set) uses which ring slot.
struct Cursor {
  long long unit;     // current unit, or >= n_units when exhausted
  int it, it_end;     // current item index / end of the unit's item list
  int j;              // frame-set within the group
  int b0, nb;
};

template <int NB>
__device__ __forceinline__ void cursor_load_unit(const BevParams& P, Cursor& c, long long n_units) {
  // position on the first accepted item of unit c.unit (or of a later unit); empty tiles are
  // visited by the consumer separately (it must still write zeros), so stop on any unit here
  if (c.unit < n_units) {
    const int4 tile = P.tiles[(int)(c.unit % P.n_tiles)];
    c.b0 = (int)(c.unit / P.n_tiles) * NB;
    c.nb = min(NB, P.batch - c.b0);
    c.it = tile.z; c.it_end = tile.z + tile.w; c.j = 0;
  }
}

// NB = frame-sets per work unit (1 for single-frame latency, 4 for batches).
#ifndef BEVK_MIN_CTAS
#define BEVK_MIN_CTAS 4
#endif
template <bool BAL, int NB>
__global__ void __launch_bounds__(256, BEVK_MIN_CTAS) k_bev(BevParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;                                                  // [STAGE_SLOTS][STAGE_SLOT_BYTES] (not BAL)
  unsigned* acc = reinterpret_cast<unsigned*>(smem_raw + (BAL ? 0 : STAGE_SLOTS * STAGE_SLOT_BYTES));   // [NB][ACC_WORDS] BGRX
  int* s_hsv = reinterpret_cast<int*>(acc + NB * ACC_WORDS);                       // [512] (BALANCE)
  __shared__ unsigned long long s_sum[BAL ? 3 * NB : 1];
  __shared__ __align__(8) unsigned long long bar_full[STAGE_SLOTS], bar_empty[STAGE_SLOTS];
  const int t = threadIdx.x, lane = t & 31, wrp = t >> 5;
  const bool staging = !BAL && P.stage;
  if (BAL) {
    s_hsv[t] = P.hsv_tab[t]; s_hsv[t + 256] = P.hsv_tab[t + 256];
    if (t < 3 * NB) s_sum[t] = 0ull;
  }
  if (t == 0) {
#pragma unroll
    for (int s = 0; s < STAGE_SLOTS; ++s) { mbar_init(&bar_full[s], BEVK_FILL == 1 ? 1 : 32); mbar_init(&bar_empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int groups = (P.batch + NB - 1) / NB;
  const long long n_units = (long long)P.n_tiles * groups;
  // accumulator word of this thread's first pixel / step to the next one, per orientation
  const int posx = (wrp * 4) * ACC_WPITCH + lane, stepx = ACC_WPITCH;   // lanes along x, k walks rows
  const int posy = lane * ACC_WPITCH + wrp * 4, stepy = 1;              // lanes along y, k walks columns
  __syncthreads();

  // ---- producer state (warp 0): cursor two fills ahead, ring position -----------------
  Cursor pc;
  pc.unit = blockIdx.x;
  cursor_load_unit<NB>(P, pc, n_units);
  unsigned p_fill = 0;      // fills issued so far (slot = p_fill % STAGE_SLOTS)
  unsigned c_fill = 0;      // fills consumed so far

  // issue boxes until the ring is full or the work is exhausted (warp 0 only, warp-uniform)
  auto produce = [&]() {
    while (pc.unit < n_units && p_fill - c_fill < (unsigned)STAGE_SLOTS) {
      if (pc.it >= pc.it_end) {           // next unit
        pc.unit += gridDim.x;
        cursor_load_unit<NB>(P, pc, n_units);
        continue;
      }
      const BevItem item = P.items[pc.it];
      const bool take = item.cam >= P.cam_lo && item.cam < P.cam_hi && item.staged;
      if (take) {
        const int b = pc.b0 + pc.j;
        const uint8_t* src = P.srcs[b * P.n_cam + item.cam];
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {      // same test on the consumer side
          const unsigned slot = p_fill % STAGE_SLOTS;
          if (p_fill >= (unsigned)STAGE_SLOTS) mbar_wait(&bar_empty[slot], ((p_fill / STAGE_SLOTS) - 1) & 1);
          unsigned char* dst = ring + slot * STAGE_SLOT_BYTES;
          const uint8_t* s0 = src + item.src_off;
#if BEVK_FILL == 1
          if (lane == 0) mbar_expect_tx(&bar_full[slot], (unsigned)(item.rows * item.row_bytes));
          __syncwarp();
          for (int r = lane; r < item.rows; r += 32)
            tma_bulk_g2s(dst + r * item.row_bytes, s0 + (size_t)r * P.pitch, (unsigned)item.row_bytes, &bar_full[slot]);
#else
          const int cpr = item.row_bytes >> 4, n16 = item.rows * cpr;   // 16-byte chunks per row / in the box
          for (int q = lane; q < n16; q += 32) {
            const int r = q / cpr, cix = q - r * cpr;
            cp_async16(dst + r * item.row_bytes + cix * 16, s0 + (size_t)r * P.pitch + cix * 16);
          }
          cp_async_arrive(&bar_full[slot]);
#endif
          ++p_fill;
        }
      }
      if (++pc.j >= pc.nb) { pc.j = 0; ++pc.it; }
    }
  };

  for (long long unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int tile_id = (int)(unit % P.n_tiles);
    const int b0 = (int)(unit / P.n_tiles) * NB;
    const int nb = min(NB, P.batch - b0);
    const int4 tile = P.tiles[tile_id];
    __syncthreads();   // previous unit's write-out is done with the accumulator
    bool first = true;   // no camera has written this tile yet: the first one stores (zeros where masked out)
    for (int it = tile.z; it < tile.z + tile.w; ++it) {
      const BevItem item = P.items[it];
      if (item.cam < P.cam_lo || item.cam >= P.cam_hi) continue;
      const uint4* __restrict__ L = P.lut + (size_t)it * (TILE * TILE) + t;
      const int pos = item.orient ? posy : posx, step = item.orient ? stepy : stepx;
      const SlowGeo geo = {P.pitch, P.FW, P.FH, P.tail_start};
#pragma unroll 1
      for (int j = 0; j < nb; ++j) {
        const int b = b0 + j;
        const uint8_t* src = P.srcs[b * P.n_cam + item.cam];
        const int dl = BAL ? P.delta[b * P.n_cam + item.cam] : 0;
        const bool staged = staging && item.staged && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        const unsigned char* box = nullptr;
        unsigned slot = 0;
        if (staging && wrp == 0) produce();
        if (staged) {
          slot = c_fill % STAGE_SLOTS;
          mbar_wait(&bar_full[slot], (c_fill / STAGE_SLOTS) & 1);
          box = ring + slot * STAGE_SLOT_BYTES;
        }
        unsigned* accj = acc + j * ACC_WORDS + pos;
        uint4 nxt = __ldg(L);
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const uint4 e = nxt;
          if (k < 3) nxt = __ldg(L + (k + 1) * 256);   // prefetch the next entry under this one's work
          unsigned* a = accj + k * step;
          if (!(e.w & LUT_ACTIVE)) {
            if (first) *a = 0u;
            continue;
          }
          unsigned v;
          if (BAL || (e.w & LUT_BORDER)) {
            v = sample_slow<BAL>(geo, src, e.x, e.w, dl, s_hsv);
          } else {
            const unsigned off_al = e.x & ~3u, sh = (e.x & 3u) * 8u;
            const bool third = (sh == 24u);
            unsigned a0, a1, a2 = 0u, b0w, b1w, b2w = 0u;
            if (staged) {
              const unsigned* q0 = reinterpret_cast<const unsigned*>(box + off_al);
              const unsigned* q1 = reinterpret_cast<const unsigned*>(box + off_al + item.row_bytes);
              a0 = q0[0]; a1 = q0[1]; b0w = q1[0]; b1w = q1[1];
              if (third) { a2 = q0[2]; b2w = q1[2]; }
            } else {
              const uint8_t* q0 = src + off_al;
              const uint8_t* q1 = q0 + P.pitch;
              a0 = ldg32(q0); a1 = ldg32(q0 + 4); b0w = ldg32(q1); b1w = ldg32(q1 + 4);
              if (third) { a2 = ldg32(q0 + 8); b2w = ldg32(q1 + 8); }
            }
            v = interp_fast(sh, e.y, e.z, e.w & 0x1ffffu, a0, a1, a2, b0w, b1w, b2w);
          }
          if (!first) v = sat_add_bgr(v, *a);   // cv2.add, camera order front, back, left, right
          *a = v;
        }
        if (staged) {      // this warp is done with the slot: let the producer refill it
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_empty[slot]);
          ++c_fill;
        }
      }
      first = false;
      __syncthreads();   // the next camera of this tile may touch the same pixels from other threads
    }
    // ---- write the tile(s): thread t -> row t/8, 4 pixels (12 bytes) at pixel 4*(t%8) ----
    const int row = t >> 3, chunk = t & 7;
    const int gy = tile.y + row, gx = tile.x + chunk * 4;
    const bool inb = (gy < P.BH) && (gx < P.BW);
    const size_t pix_off = (size_t)gy * P.BW * 3 + (size_t)gx * 3;
    const bool full = inb && (gx + 4 <= P.BW) && ((P.BW * 3) % 4 == 0) && (P.canvas_bytes % 4 == 0);
    const int npx = inb ? min(4, P.BW - gx) : 0;
    unsigned c0 = 0, c1 = 0, c2 = 0;
    if (!BAL && P.car && full) {
      const unsigned* c = reinterpret_cast<const unsigned*>(P.car + pix_off);
      c0 = __ldg(c); c1 = __ldg(c + 1); c2 = __ldg(c + 2);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (j >= nb) break;
      const unsigned* a = acc + j * ACC_WORDS + row * ACC_WPITCH + chunk * 4;
      unsigned x0 = a[0], x1 = a[1], x2 = a[2], x3 = a[3];                 // BGRX BGRX BGRX BGRX
      if (first) x0 = x1 = x2 = x3 = 0u;                                   // tile without a camera (car hole)
      unsigned w0 = __byte_perm(x0, x1, 0x4210);                           // B0 G0 R0 B1
      unsigned w1 = __byte_perm(x1, x2, 0x5421);                           // G1 R1 B2 G2
      unsigned w2 = __byte_perm(x2, x3, 0x6542);                           // R2 B3 G3 R3
      if (BAL) {   // channel sums of the composed canvas, before gains and car (surroundBEV.py:44-47)
        const unsigned px[4] = {x0, x1, x2, x3};
        unsigned sb = 0, sg = 0, sr = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < npx) { sb += px[q] & 255u; sg += (px[q] >> 8) & 255u; sr += (px[q] >> 16) & 255u; }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {   // every lane takes part (out-of-canvas lanes add 0)
          sb += __shfl_xor_sync(0xffffffffu, sb, s);
          sg += __shfl_xor_sync(0xffffffffu, sg, s);
          sr += __shfl_xor_sync(0xffffffffu, sr, s);
        }
        if (lane == 0) {
          atomicAdd(&s_sum[3 * j + 0], (unsigned long long)sb);
          atomicAdd(&s_sum[3 * j + 1], (unsigned long long)sg);
          atomicAdd(&s_sum[3 * j + 2], (unsigned long long)sr);
        }
      }
      if (!inb) continue;
      uint8_t* o = P.out + (size_t)(b0 + j) * P.canvas_bytes + pix_off;
      if (full) {
        if (!BAL && P.car) { w0 = __vaddus4(w0, c0); w1 = __vaddus4(w1, c1); w2 = __vaddus4(w2, c2); }
        unsigned* g = reinterpret_cast<unsigned*>(o);
        g[0] = w0; g[1] = w1; g[2] = w2;
      } else {
        const unsigned wv[3] = {w0, w1, w2};
#pragma unroll 1
        for (int i = 0; i < npx * 3; ++i) {
          int v = (wv[i >> 2] >> (8 * (i & 3))) & 255u;
          if (!BAL && P.car) v = min(255, v + P.car[pix_off + i]);
          o[i] = (uint8_t)v;
        }
      }
    }
    if (BAL) {
      __syncthreads();
      if (t < 3 * nb) { atomicAdd(P.csum + (size_t)(b0 + t / 3) * 3 + (t % 3), s_sum[t]); s_sum[t] = 0ull; }
    }
  }
}

constexpr size_t bev_smem_bytes(bool bal, int nb) {
  return (bal ? 0 : (size_t)STAGE_SLOTS * STAGE_SLOT_BYTES) + (size_t)nb * ACC_WORDS * 4 + 512 * sizeof(int);
}

}  // namespace bevk
