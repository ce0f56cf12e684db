// rtclj_p3_swar.h -- byte-parallel decimal formatting used by the device P3 writer
// (rtclj_p3_kernels.cuh).  Plain integer code, compiled for the device and -- by the host tests,
// which check it exhaustively -- for the CPU.
//
// Four 8-bit values arrive packed in a 32-bit word (first value in the low byte).  Everything that
// can be done for several values at once is: digit counts for four values (carry-free byte lanes),
// digits for two values (16-bit lanes).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define RTCLJ_HD __host__ __device__ __forceinline__
#else
#define RTCLJ_HD inline
#endif

namespace rtclj {

RTCLJ_HD uint32_t p3_byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, sel);
#else
  const uint64_t ab = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) r |= (uint32_t)((ab >> (8 * ((sel >> (4 * k)) & 7u))) & 0xffu) << (8 * k);
  return r;
#endif
}

// Per byte lane: (number of decimal digits - 1) of that byte, i.e. 0, 1 or 2.
RTCLJ_HD uint32_t p3_extra_digits4(uint32_t w) {
  const uint32_t low7 = w & 0x7f7f7f7fu;
  const uint32_t ge10 = ((low7 + 0x76767676u) | w) & 0x80808080u;   // bit 7 of a lane: value >= 10
  const uint32_t ge100 = ((low7 + 0x1c1c1c1cu) | w) & 0x80808080u;  // bit 7 of a lane: value >= 100
  return (ge10 >> 7) + (ge100 >> 7);
}

// Text bytes of the four values of w, each followed by one separator byte.
RTCLJ_HD uint32_t p3_len4(uint32_t w) {
  const uint32_t e = p3_extra_digits4(w);        // lanes hold 0..2
  const uint32_t s = (e & 0x00ff00ffu) + ((e >> 8) & 0x00ff00ffu);
  return 8u + (s & 0xffu) + (s >> 16);
}

// Two values in the 16-bit lanes of x (each 0..255): hundreds, tens and ones, lane-wise.
struct P3Digits2 { uint32_t ht, o; };  // ht lanes: hundreds | tens << 8 ; o lanes: ones
RTCLJ_HD P3Digits2 p3_digits2(uint32_t x) {
  const uint32_t h = ((x * 41u) >> 12) & 0x000f000fu;   // floor(v/100) for v < 256 (255*41 < 2^16: no carry between lanes)
  const uint32_t r = x - h * 100u;
  const uint32_t t = ((r * 205u) >> 11) & 0x000f000fu;  // floor(r/10) for r < 100
  P3Digits2 d;
  d.o = r - t * 10u;
  d.ht = h + (t << 8);
  return d;
}

// The 4-byte field "h t o sep" (ASCII) of the value in 16-bit lane `lane` (0 or 1).
RTCLJ_HD uint32_t p3_field(const P3Digits2& d, int lane, uint32_t sep) {
  // bytes: [ht.b0, ht.b1, o.b0, o.b1 (= 0)] for lane 0; [ht.b2, ht.b3, o.b2, o.b3 (= 0)] for lane 1
  const uint32_t f = p3_byte_perm(d.ht, d.o, lane ? 0x7632u : 0x5410u);
  return f + (0x00303030u | (sep << 24));
}

// ---- shifts that are defined for the whole range 0..32 (C++ leaves x << 32 undefined; PTX does not)
RTCLJ_HD uint32_t p3_shl(uint32_t x, uint32_t s) {   // x << s, 0 for s = 32
#ifdef __CUDA_ARCH__
  return __funnelshift_lc(0u, x, s);
#else
  return s >= 32u ? 0u : x << s;
#endif
}
RTCLJ_HD uint32_t p3_shr(uint32_t x, uint32_t s) {   // x >> s, 0 for s = 32
#ifdef __CUDA_ARCH__
  return __funnelshift_rc(x, 0u, s);
#else
  return s >= 32u ? 0u : x >> s;
#endif
}
RTCLJ_HD uint32_t p3_funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {  // high word of (hi:lo) << s, s in 0..31
#ifdef __CUDA_ARCH__
  return __funnelshift_l(lo, hi, s);
#else
  return s ? (hi << s) | (lo >> (32u - s)) : hi;
#endif
}

// The text of one pixel, "r g b\n", as 6..12 little-endian bytes in three words.
// f0..f2: the 4-byte fields of its values (p3_field), d0..d2: 8 * (leading zero digits to drop).
struct P3Pixel { uint32_t w0, w1, w2, bits; };
RTCLJ_HD P3Pixel p3_pixel_text(uint32_t f0, uint32_t f1, uint32_t f2, uint32_t d0, uint32_t d1, uint32_t d2) {
  const uint32_t c0 = f0 >> d0, c1 = f1 >> d1, c2 = f2 >> d2;   // d <= 16
  const uint32_t n0 = 32u - d0, n01 = n0 + (32u - d1);          // bits of r; of r and g (32..64)
  P3Pixel p;
  p.w0 = c0 | p3_shl(c1, n0);
  p.w1 = (c1 >> (32u - n0)) | p3_shl(c2, n01 - 32u);
  p.w2 = p3_shr(c2, 64u - n01);
  p.bits = n01 + (32u - d2);
  return p;
}

// Appends a pixel to a byte stream whose pending word `lo` holds fill8 (0, 8, 16, 24) valid bits.
// A pixel is at least 48 bits, so out[0] is always a complete word; nfull (1..3) words are complete.
struct P3Append { uint32_t out0, out1, out2, nfull, lo, fill8; };
RTCLJ_HD P3Append p3_append_pixel(uint32_t lo, uint32_t fill8, const P3Pixel& p) {
  const uint32_t t0 = lo | (p.w0 << fill8);
  const uint32_t t1 = p3_funnel_l(p.w0, p.w1, fill8);
  const uint32_t t2 = p3_funnel_l(p.w1, p.w2, fill8);
  const uint32_t t3 = p3_funnel_l(p.w2, 0u, fill8);
  const uint32_t total = fill8 + p.bits;   // 48..120
  P3Append a;
  a.out0 = t0; a.out1 = t1; a.out2 = t2;
  a.nfull = total >> 5;
  a.lo = a.nfull == 1u ? t1 : (a.nfull == 2u ? t2 : t3);
  a.fill8 = total & 31u;
  return a;
}

}  // namespace rtclj
