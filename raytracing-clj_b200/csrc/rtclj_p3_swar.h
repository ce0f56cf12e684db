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

// Little-endian byte accumulator: appends the text of one value (its field with the leading
// zero digits dropped) and reports whether a whole 32-bit word is ready.
struct P3Acc {
  uint32_t lo = 0, hi = 0;
  uint32_t fill8 = 0;  // valid bits in lo (0, 8, 16 or 24 between appends)
};
// drop8 = 8 * (number of leading zero digits to drop) = 16 - 8 * (digits - 1); drop8 = 32 appends nothing
// (the padding values of a partial thread).
RTCLJ_HD void p3_acc_append(P3Acc& a, uint32_t field, uint32_t drop8) {
#ifdef __CUDA_ARCH__
  const uint32_t c = __funnelshift_rc(field, 0u, drop8);  // field >> drop8, 0 for drop8 = 32
  a.hi = __funnelshift_l(c, 0u, a.fill8);                 // the bits of c that do not fit in lo (0 when fill8 == 0)
#else
  const uint32_t c = drop8 >= 32u ? 0u : field >> drop8;
  a.hi = a.fill8 ? c >> (32u - a.fill8) : 0u;
#endif
  a.lo |= c << a.fill8;
  a.fill8 += 32u - drop8;
}
RTCLJ_HD bool p3_acc_full(const P3Acc& a) { return a.fill8 >= 32u; }
RTCLJ_HD uint32_t p3_acc_pop(P3Acc& a) {
  const uint32_t w = a.lo;
  a.lo = a.hi; a.hi = 0u; a.fill8 -= 32u;
  return w;
}
// Per byte lane: 16 where the value at that position of word `wi` lies beyond the thread's n pixels
// (added to the drop lanes, it turns their 16 into the "append nothing" 32).
RTCLJ_HD uint32_t p3_padding_lanes(int wi, int n) {
  uint32_t m = 0;
  for (int j = 0; j < 4; ++j)
    if (4 * wi + j >= 3 * n) m |= 0x10u << (8 * j);
  return m;
}

}  // namespace rtclj
