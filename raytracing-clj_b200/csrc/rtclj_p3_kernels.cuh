// rtclj_p3_kernels.cuh -- the P3 text writer on the device.
//
// Replaces the write-color! loop of the reference (src/raytracing.clj:172-175,
// realm/raytracing.clj:350-358): "P3\nW H\n255\n" and one "r g b\n" line per pixel, in image order.
// Byte work, HBM bound: 3 bytes read and 6..12 bytes written per pixel.
//
// One persistent launch (p3_encode_kernel), every CTA resident at once, CTA b owns a contiguous run of
// 1024-pixel tiles:
//   1. text bytes of the whole run (one pass over its pixels)           -> state[2 + b]
//   2. start of the run = header + the counts of all CTAs before it     (spin on their words; a CTA
//      only ever waits for CTAs that took an earlier ticket, so the wait cannot deadlock)
//   3. per tile: thread offsets by a CTA scan; digits for two values at a time in 16-bit lanes, a pixel's
//      three fields concatenated into its 6..12 bytes in registers, shifted to the stream's byte phase
//      and stored as whole words; staged in shared memory at the alignment of its destination, written
//      with 16-byte stores.
//      The pixels come from L2 the second time (an image is far smaller than the 126 MB L2).
// count_only stops after step 2 (sizing calls, and callers whose buffer is below the worst case).
#pragma once
#include <cstddef>
#include <cstdint>

#include "rtclj_p3_swar.h"

namespace rtclj {

constexpr int kP3Threads = 256;
constexpr int kP3PixPerThread = 4;                            // 12 bytes = three 32-bit loads
constexpr int kP3PixPerBlock = kP3Threads * kP3PixPerThread;  // 1024 pixels, at most 12 KiB of text
constexpr int kP3StageBytes = kP3PixPerBlock * 12 + 32;
constexpr unsigned long long kP3Ready = 1ull << 63;

struct P3Header { char s[60]; int n; };

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// The (up to) four pixels of this thread in tile `tile`, packed little-endian into w[3]; returns how many exist.
__device__ __forceinline__ int p3_load(const unsigned char* __restrict__ rgb8, size_t npix, bool aligned4, size_t tile,
                                       uint32_t w[3]) {
  const size_t p = (tile * kP3Threads + threadIdx.x) * kP3PixPerThread;
  w[0] = w[1] = w[2] = 0u;
  if (p >= npix) return 0;
  const size_t left = npix - p;
  const int n = left < (size_t)kP3PixPerThread ? (int)left : kP3PixPerThread;
  const unsigned char* src = rgb8 + 3 * p;
  if (aligned4 && n == kP3PixPerThread) {
    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
    w[0] = __ldg(s32); w[1] = __ldg(s32 + 1); w[2] = __ldg(s32 + 2);
  } else {
#pragma unroll
    for (int k = 0; k < 12; ++k)
      if (k < 3 * n) w[k >> 2] |= (uint32_t)__ldg(src + k) << (8 * (k & 3));
  }
  return n;
}

// Text bytes of the thread's n pixels (padding bytes of a partial thread are zero: one digit + separator each).
__device__ __forceinline__ unsigned p3_thread_len(const uint32_t w[3], int n) {
  return p3_len4(w[0]) + p3_len4(w[1]) + p3_len4(w[2]) - 2u * (unsigned)(12 - 3 * n);
}

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long x, unsigned long long* red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
  __syncthreads();  // red may still be read from an earlier call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  unsigned long long t = 0;
#pragma unroll
  for (int w = 0; w < kP3Threads / 32; ++w) t += red[w];
  return t;
}

// state: [0] ticket, [1] total text length (out), [2 + b] (kP3Ready | text bytes of CTA b's run); zeroed by the host.
__global__ void __launch_bounds__(kP3Threads) p3_encode_kernel(const unsigned char* __restrict__ rgb8, size_t npix, int aligned4_,
                                                               size_t ntiles, size_t tiles_per_cta,
                                                               unsigned long long* __restrict__ state,
                                                               unsigned char* __restrict__ out, P3Header hdr, int count_only) {
  __shared__ __align__(16) unsigned char stage[kP3StageBytes];
  __shared__ unsigned long long red[kP3Threads / 32];
  __shared__ unsigned wsum[kP3Threads / 32];
  __shared__ unsigned cta_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool aligned4 = aligned4_ != 0;
  if (threadIdx.x == 0) cta_s = (unsigned)atomicAdd(state, 1ull);
  {  // boundary words of neighbouring threads are merged with atomicOr: the stage starts (and is kept) zero
    uint4* z = reinterpret_cast<uint4*>(stage);
    for (unsigned i = threadIdx.x; i < kP3StageBytes / 16; i += kP3Threads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  const size_t cta = cta_s;
  const size_t t0 = cta * tiles_per_cta;
  const size_t t1 = t0 + tiles_per_cta < ntiles ? t0 + tiles_per_cta : ntiles;
  if (t0 >= ntiles) {  // nothing to encode (grid rounding): still publish, later CTAs sum every predecessor
    if (threadIdx.x == 0) st_state(state + 2 + cta, kP3Ready);
    return;
  }

  // 1. text bytes of the run; the loads of eight tiles are in flight together
  unsigned long long mine = 0;
  for (size_t tb = t0; tb < t1; tb += 8) {
    uint32_t wb[8][3];
    int nb_[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      nb_[u] = 0;
      if (tb + u < t1) nb_[u] = p3_load(rgb8, npix, aligned4, tb + u, wb[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (tb + u < t1) mine += p3_thread_len(wb[u], nb_[u]);
  }
  const unsigned long long run_bytes = block_sum_u64(mine, red);
  if (threadIdx.x == 0) st_state(state + 2 + cta, kP3Ready | run_bytes);

  // 2. where the run starts
  unsigned long long before = 0;
  for (size_t j = threadIdx.x; j < cta; j += kP3Threads) {
    unsigned long long s;
    do { s = ld_state(state + 2 + j); } while (!(s & kP3Ready));
    before += s & ~kP3Ready;
  }
  unsigned long long g = (unsigned long long)hdr.n + block_sum_u64(before, red);
  if (t1 == ntiles && threadIdx.x == 0) state[1] = g + run_bytes;
  if (count_only) return;
  if (cta == 0 && (int)threadIdx.x < hdr.n) out[threadIdx.x] = (unsigned char)hdr.s[threadIdx.x];

  // 3. the text, tile by tile; the next tile's pixels are requested before this tile is formatted
  uint32_t w[3], wn[3] = {0u, 0u, 0u};
  int n = p3_load(rgb8, npix, aligned4, t0, w), nn = 0;
  for (size_t t = t0; t < t1; ++t) {
    if (t + 1 < t1) nn = p3_load(rgb8, npix, aligned4, t + 1, wn);
    unsigned len = p3_thread_len(w, n);
    unsigned inc = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += y;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned offset = inc - len, total = 0;
#pragma unroll
    for (int k = 0; k < kP3Threads / 32; ++k) {
      const unsigned s = wsum[k];
      if (k < warp) offset += s;
      total += s;
    }
    unsigned char* dst = out + g;
    const unsigned shift = (unsigned)(reinterpret_cast<uintptr_t>(dst) & 15u);
    if (n > 0) {
      // Pixel by pixel: the three fields of a pixel are concatenated into its 6..12 bytes of text
      // (three words), which are shifted to the stream's byte phase and stored.  A pixel is at least
      // six bytes, so every pixel completes at least one word: the thread's first word (which may hold
      // bytes of the previous thread) is always the first word of its first pixel, and its last,
      // partial word may hold bytes of the next thread -- those two are merged with atomicOr into the
      // zeroed stage, everything in between is plain predicated stores.  No per-value bookkeeping.
      uint32_t* stage32 = reinterpret_cast<uint32_t*>(stage);
      const unsigned s0 = shift + offset;
      unsigned widx = s0 >> 2;
      uint32_t lo = 0u, fill8 = 8u * (s0 & 3u);
      uint32_t drop[3];
      P3Digits2 even[3], odd[3];
#pragma unroll
      for (int wi = 0; wi < 3; ++wi) {
        drop[wi] = 0x10101010u - (p3_extra_digits4(w[wi]) << 3);  // per value: 8 * leading zeros dropped
        even[wi] = p3_digits2(w[wi] & 0x00ff00ffu);
        odd[wi] = p3_digits2((w[wi] >> 8) & 0x00ff00ffu);
      }
#pragma unroll
      for (int px = 0; px < kP3PixPerThread; ++px) {
        if (px < n) {
          uint32_t f[3], d[3];
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            const int k = 3 * px + ch, wi = k >> 2, j = k & 3;  // value k sits in byte j of word wi
            f[ch] = p3_field((j & 1) ? odd[wi] : even[wi], j >> 1, ch == 2 ? 0x0au : 0x20u);
            d[ch] = p3_byte_perm(drop[wi], 0u, 0x4440u + (unsigned)j);
          }
          const P3Pixel pix = p3_pixel_text(f[0], f[1], f[2], d[0], d[1], d[2]);
          const P3Append ap = p3_append_pixel(lo, fill8, pix);
          if (px == 0) atomicOr(stage32 + widx, ap.out0); else stage32[widx] = ap.out0;
          if (ap.nfull >= 2u) stage32[widx + 1] = ap.out1;
          if (ap.nfull == 3u) stage32[widx + 2] = ap.out2;
          widx += ap.nfull;
          lo = ap.lo;
          fill8 = ap.fill8;
        }
      }
      if (fill8) atomicOr(stage32 + widx, lo);
    }
    __syncthreads();
    // stage[shift .. shift+total) -> dst[0 .. total): whole 16-byte chunks as vectors, the two ends
    // bytewise; every chunk is zeroed again by the thread that copied it
    const unsigned end = shift + total;
    unsigned char* abase = dst - shift;  // 16-byte aligned
    for (unsigned c = threadIdx.x; c * 16u < end; c += kP3Threads) {
      uint4* sp = reinterpret_cast<uint4*>(stage) + c;
      const uint4 q = *sp;
      const unsigned lo = c * 16u;
      if (lo >= shift && lo + 16u <= end) {
        reinterpret_cast<uint4*>(abase)[c] = q;
      } else {
        const unsigned b0 = lo > shift ? lo : shift, b1 = lo + 16u < end ? lo + 16u : end;
        for (unsigned i = b0; i < b1; ++i) abase[i] = stage[i];
      }
      *sp = make_uint4(0u, 0u, 0u, 0u);
    }
    g += total;
    w[0] = wn[0]; w[1] = wn[1]; w[2] = wn[2]; n = nn;
  }
}

}  // namespace rtclj
