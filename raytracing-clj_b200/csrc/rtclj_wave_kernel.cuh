// rtclj_wave_kernel.cuh -- the render loop of keychera/raytracing-clj (src/raytracing.clj:141-171,
// src/realm/raytracing.clj:325-346) as a WAVEFRONT path tracer inside each CTA, for scenes whose cull
// table fits the kernel-parameter constant bank (<= 512 spheres: BASELINE.json configs 1-4).
//
// Why: in the lane-owns-a-path kernel (render_kernel<true>) the cull runs at 32 active lanes, but the
// exact resolve, the three materials, the sky, Philox and the refill run at 8-15 (ncu, round 1:
// smsp__thread_inst_executed_per_inst_executed = 24.75, ~55 % of warp-time in divergent code).
// Here a path is not tied to a lane.  Path state lives in shared memory ("slots", SoA); slot indices
// travel through four queues, one per phase, and every warp runs a small scheduler loop:
//
//     Q_TRACE   closest hit: fp32 cull over all spheres (FFMA2, table in uniform registers) +
//               exact fp64 resolve (hit-anything, raytracing.clj:33-43; hittable.clj:9-31)
//     Q_UNITV   lambertian / metal scatter (material.clj:13-28): unit-vector rejection sampling
//     Q_DIEL    dielectric scatter with Schlick (material.clj:30-46)
//     Q_FIN     end of a sample: sky (raytracing.clj:55-58) x attenuation product, accumulate
//               (raytracing.clj:153), unit bookkeeping, the pixel's mean + write-color! quantisation
//               (raytracing.clj:19-26,155), a new work unit from the global ticket counter, and the
//               next camera ray (raytracing.clj:144-151)
//
// A warp claims up to 32 slots of ONE queue, so all its lanes execute the same material / phase; the
// warp-level compaction against path-length divergence (north_star) is the queue itself.  There is no
// CTA barrier after start-up: warps that cull (FMA pipe) and warps that shade (fp64 / ALU pipes)
// overlap freely.  Results do not depend on the schedule: every draw is Philox keyed by
// (pixel, sample, stage, block) and each unit sum has one owner.
//
// Arithmetic is the same as render_kernel's and the oracle's: fp64, the reference's evaluation order,
// no FMA contraction (--fmad=false); the fp32 cull only prunes.
#pragma once
#include "rtclj_kernels.cuh"

namespace rtclj {

#ifndef RTCLJ_WAVE_THREADS
#define RTCLJ_WAVE_THREADS 640
#endif
#ifndef RTCLJ_WAVE_SLOTS
#define RTCLJ_WAVE_SLOTS 1024
#endif
constexpr int kWT = RTCLJ_WAVE_THREADS;  // threads per CTA, one CTA per SM
constexpr int kWS = RTCLJ_WAVE_SLOTS;    // paths in flight per CTA
constexpr int kWCap = 2048;              // queue capacity (entries), a power of two >= kWS
static_assert(kWS <= kWCap && (kWCap & (kWCap - 1)) == 0 && kWS < 0xffff, "queue capacity");
static_assert(kWT % 32 == 0 && kWS % 32 == 0, "whole warps");

// queue index = scheduling priority: the shading queues feed Q_TRACE, so they go first
enum { Q_UNITV = 0, Q_FIN = 1, Q_DIEL = 2, Q_TRACE = 3, Q_COUNT = 4 };
constexpr int K_FRESH = -4;              // a slot that has no work unit yet
constexpr unsigned kQEmpty = 0xffffu;    // queue entry not written yet / already consumed

// dynamic shared memory layout (bytes)
struct WaveSmem {
  static constexpr size_t ctl = 0;                                          // head[4] tail[4] dead: 64 B
  static constexpr size_t queues = 64;                                      // Q_COUNT x kWCap x u16
  static constexpr size_t masks = queues + (size_t)Q_COUNT * kWCap * 2;     // 32 blocks x kWT x u16 cull masks
  static constexpr size_t f64 = (masks + (size_t)32 * kWT * 2 + 15) & ~(size_t)15;  // 9 arrays of kWS doubles
  static constexpr size_t u32 = f64 + (size_t)9 * kWS * 8;                  // 7 arrays of kWS words
  static constexpr size_t total = u32 + (size_t)7 * kWS * 4;
};

// release / acquire across CTAs without a full fence.acq_rel.gpu (which also invalidates L1, where
// the sphere and material records live): the unit sums travel through L2 (st/ld.relaxed.gpu), the
// arrival counter is bumped with a RELEASE atomic, and the last arriver's loads depend on its result.
__device__ __forceinline__ void st_gpu(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_gpu(const double* p) {
  double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned arrive_release(unsigned* p) {
  unsigned old; asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p) : "memory"); return old;
}

// pixel mean (raytracing.clj:155 sum / spp; realm/raytracing.clj:25,344 sum * (1/spp)) and
// write-color! (raytracing.clj:19-26); p_local = index of the pixel among this shard's pixels
__device__ __noinline__ void finish_pixel(const KParams& P, unsigned p_local, double r, double g, double b) {
  if (P.flags & F_MEAN_DIVIDE) {
    const double s = (double)P.spp;
    r = r / s; g = g / s; b = b / s;
  } else {
    const double s = 1.0 / (double)P.spp;
    r = r * s; g = g * s; b = b * s;
  }
  const int lr = (int)(p_local / (unsigned)P.W);
  const int i = (int)(p_local - (unsigned)lr * (unsigned)P.W);
  const int tile = lr / P.shard_rows;
  const int j = (tile * P.shard_count + P.shard_index) * P.shard_rows + (lr - tile * P.shard_rows);
  const size_t o = 3ull * ((size_t)j * (size_t)P.W + (size_t)i);
  if (P.out_linear) { P.out_linear[o] = r; P.out_linear[o + 1] = g; P.out_linear[o + 2] = b; }
  if (P.out_rgb8) {
    const bool lin = (P.flags & F_QUANT_LINEAR) != 0;
    P.out_rgb8[o] = quantise(r, lin); P.out_rgb8[o + 1] = quantise(g, lin); P.out_rgb8[o + 2] = quantise(b, lin);
  }
}


// sky x the attenuations of the path's scattering hits, in the reference's order:
// ((sky*att_n)*att_{n-1})...*att_1 (raytracing.clj:52-53) or ((1*att_1)*att_2)...*att_n * sky
// (realm/raytracing.clj:206,225,236)
__device__ __noinline__ d3 attenuate(const KParams& P, const unsigned short* col, unsigned nst, d3 sky) {
  if (P.flags & F_REVERSE_PRODUCT) {
    d3 color = sky;
#pragma unroll 1
    for (int t = (int)nst - 1; t >= 0; --t) color = mulv(color, ld3(P.mat[__ldcg(col + (size_t)t * P.stack_stride)].albedo));
    return color;
  }
  d3 T = mk(1.0, 1.0, 1.0);
#pragma unroll 1
  for (int t = 0; t < (int)nst; ++t) T = mulv(T, ld3(P.mat[__ldcg(col + (size_t)t * P.stack_stride)].albedo));
  return mulv(T, sky);
}
// the chunk that arrives last adds the pixel's unit sums in index order and finishes the pixel
__device__ __noinline__ void finish_chunked(const KParams& P, unsigned unit, double sr, double sg, double sb) {
  double* out = P.partial + (size_t)unit * 3u;
  st_gpu(out, sr); st_gpu(out + 1, sg); st_gpu(out + 2, sb);
  const unsigned p_local = unit / (unsigned)P.nchunks;
  if (arrive_release(P.arrive + p_local) == (unsigned)P.nchunks - 1u) {
    const double* src = P.partial + (size_t)p_local * (size_t)P.nchunks * 3u;
    double r = 0.0, g = 0.0, b = 0.0;
#pragma unroll 1
    for (int c = 0; c < P.nchunks; ++c) { r = r + ld_gpu(src + 3 * c); g = g + ld_gpu(src + 3 * c + 1); b = b + ld_gpu(src + 3 * c + 2); }
    finish_pixel(P, p_local, r, g, b);
  }
}

// Appends this lane's slot to queue `dest` (-1: nothing).  MATCH groups the lanes by destination, the
// first lane of each group reserves the group's entries with one shared-memory atomic; a producer waits
// until the entry it was given has been consumed (it always has: at most kWS <= kWCap entries are ever
// unclaimed).
__device__ __forceinline__ void wave_push(volatile unsigned* ctl, unsigned short* qbuf, int dest, int slot, int lane) {
  __threadfence_block();  // the slot's state before its index
  const unsigned peers = __match_any_sync(0xffffffffu, dest);
  const int leader = __ffs(peers) - 1;
  unsigned pos = 0;
  if (lane == leader && dest >= 0) pos = atomicAdd(const_cast<unsigned*>(ctl) + 4 + dest, (unsigned)__popc(peers));
  pos = __shfl_sync(0xffffffffu, pos, leader);
  if (dest >= 0) {
    volatile unsigned short* e = qbuf + dest * kWCap + ((pos + (unsigned)__popc(peers & ((1u << lane) - 1u))) & (unsigned)(kWCap - 1));
    while (*e != kQEmpty) {}
    *e = (unsigned short)slot;
  }
}

// out-of-line helpers: the kernel's instruction footprint must stay inside the instruction cache (the
// first version of this kernel, 59 KB of SASS, stalled 5.5 warps per issue on instruction fetch)
__device__ __noinline__ d3 divs_by_ni(d3 v, double d) { return divs_by(v, d); }
// outward unit normal of sphere `best` at Pt (hittable.clj:25), flipped against the ray (hit.clj:14-15)
struct FaceNormal { d3 n; bool front; };
__device__ __forceinline__ FaceNormal face_normal(const Geom64* __restrict__ geom64, int best, d3 Pt, d3 D) {
  const double2 g0 = __ldg(reinterpret_cast<const double2*>(geom64 + best));
  const double2 g1 = __ldg(reinterpret_cast<const double2*>(geom64 + best) + 1);
  const d3 outward = divs_by_ni(sub(Pt, mk(g0.x, g0.y, g1.x)), g1.y);
  FaceNormal f;
  f.front = dot(D, outward) < 0.0;
  f.n = f.front ? outward : neg(outward);
  return f;
}

__global__ void __launch_bounds__(kWT, 1) render_wave_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  volatile unsigned* ctl = reinterpret_cast<volatile unsigned*>(smem_raw + WaveSmem::ctl);
  unsigned short* qbuf = reinterpret_cast<unsigned short*>(smem_raw + WaveSmem::queues);
  const int tid = threadIdx.x, lane = tid & 31;
  unsigned short* my_mask = reinterpret_cast<unsigned short*>(smem_raw + WaveSmem::masks) + tid;  // [block * kWT]
  double* const sOx = reinterpret_cast<double*>(smem_raw + WaveSmem::f64);
  double* const sOy = sOx + kWS; double* const sOz = sOy + kWS;
  double* const sDx = sOz + kWS; double* const sDy = sDx + kWS; double* const sDz = sDy + kWS;
  double* const sSr = sDz + kWS; double* const sSg = sSr + kWS; double* const sSb = sSg + kWS;
  unsigned* const sPixel = reinterpret_cast<unsigned*>(smem_raw + WaveSmem::u32);
  unsigned* const sK = sPixel + kWS; unsigned* const sKend = sK + kWS; unsigned* const sUnit = sKend + kWS;
  unsigned* const sDepth = sUnit + kWS; unsigned* const sNstack = sDepth + kWS;
  unsigned* const sMeta = sNstack + kWS;  // (kind & 0xff) << 16 | index of the sphere hit
  const unsigned FULL = 0xffffffffu;
  const unsigned flags = P.flags;
  const size_t gslot0 = (size_t)blockIdx.x * kWS;  // this CTA's columns of the attenuation stack

  // ---- start-up: empty queues; every slot enters Q_FIN as FRESH (it will draw a ticket there)
  for (int i = tid; i < Q_COUNT * kWCap; i += kWT) qbuf[i] = (unsigned short)kQEmpty;
  if (tid < 16) ctl[tid] = 0u;
  __syncthreads();
  for (int i = tid; i < kWS; i += kWT) {
    qbuf[Q_FIN * kWCap + i] = (unsigned short)i;
    sMeta[i] = ((unsigned)K_FRESH & 0xffu) << 16;
  }
  if (tid == 0) ctl[4 + Q_FIN] = (unsigned)kWS;
  __syncthreads();

  unsigned n_samples = 0, n_seg = 0, n_exact = 0, n_pref = 0;

  for (;;) {
    // ---- scheduler: claim up to 32 entries of one queue.  Shading queues go first once they fill a
    // warp (they feed Q_TRACE); a trace needs a full warp too, because the cull costs the same for 1
    // or 32 rays.  Partial warps run only when nothing fills a warp and nothing more is on its way.
    // Lane q < 4 looks at queue q.  Every decision below is taken on ballots / REDUX results, so the
    // compiler KNOWS the branches are warp-uniform (otherwise the cull loses its uniform table loads).
    int q = 0;
    unsigned base, n = 0;
    {
      int waited = 0;
      bool finished = false;
      unsigned h = 0, avail = 0;
      for (;;) {
        h = 0; avail = 0;
        if (lane < Q_COUNT) { h = ctl[lane]; avail = ctl[4 + lane] - h; }  // head BEFORE tail: never underflows
        unsigned pickable = __ballot_sync(FULL, avail >= 32u);             // queues that fill a warp
        if (pickable == 0u) {
          pickable = __ballot_sync(FULL, avail != 0u);
          const unsigned dead = ctl[8];
          if (pickable == 0u) {
            finished = __any_sync(FULL, dead >= (unsigned)kWS);  // every path of this CTA has ended
            if (finished) break;
            __nanosleep(200);
            continue;
          }
          // slots inside some warp's handler right now: fuller warps are coming, wait a little
          const unsigned queued = __reduce_add_sync(FULL, avail);
          if (__any_sync(FULL, (unsigned)kWS - dead - queued > 0u && waited < 8)) { ++waited; __nanosleep(150); continue; }
        }
        q = __ffs(pickable) - 1;
        unsigned ok = 0;
        if (lane == q) {
          n = avail < 32u ? avail : 32u;
          ok = atomicCAS(const_cast<unsigned*>(ctl) + q, h, h + n) == h ? 1u : 0u;
        }
        if (__any_sync(FULL, ok != 0u)) break;
      }
      if (finished) break;
      base = __shfl_sync(FULL, h, q);
      n = __shfl_sync(FULL, n, q);
    }
    int slot = -1;
    if ((unsigned)lane < n) {
      volatile unsigned short* e = qbuf + q * kWCap + ((base + (unsigned)lane) & (unsigned)(kWCap - 1));
      unsigned v;
      do { v = *e; } while (v == kQEmpty);
      *e = (unsigned short)kQEmpty;
      slot = (int)v;
    }
    __threadfence_block();  // the index before the slot's state
    // After a spin-wait ptxas no longer assumes the warp converged, and would demote every uniform
    // instruction that follows (the cull's LDCU / UR operands) to per-lane code: reconverge explicitly.
    __syncwarp();
    const bool live = slot >= 0;
    const int s = live ? slot : 0;

    int dest = -1;
    if (q == Q_TRACE) {
      // =========================================================== closest hit
      d3 O = mk(0.0, 0.0, 0.0), D = mk(0.0, 0.0, 1.0);
      if (live) { O = mk(sOx[s], sOy[s], sOz[s]); D = mk(sDx[s], sDy[s], sDz[s]); }
      // ---- fp32 view of the ray for the cull (coordinates translated by -shift)
      bool scan_all = (flags & F_NO_CULL) != 0;
      const float ofx = (float)(O.x - P.shift[0]), ofy = (float)(O.y - P.shift[1]), ofz = (float)(O.z - P.shift[2]);
      float dhx, dhy, dhz, len32;
      {
        const float dfx = (float)D.x, dfy = (float)D.y, dfz = (float)D.z;
        const float l2 = dfx * dfx + dfy * dfy + dfz * dfz;
        const float inv = rsqrtf(l2);
        if (!(l2 > 1e-30f && l2 < 1e30f)) scan_all = true;  // degenerate direction: exact scan
        dhx = dfx * inv; dhy = dfy * inv; dhz = dfz * inv;
        len32 = l2 * inv;  // |d| to ~8 eps
      }
      const float mo = fmaxf(fabsf(ofx), fmaxf(fabsf(ofy), fabsf(ofz)));
      // Conservative discriminant in expanded form (8 packed ops per sphere pair):
      //   D' = b^2 + s,  b = c.dhat - o.dhat,  s = Ws + 2 c.o - |o|^2(1 - 96 eps)   (DESIGN.md "cull error bound")
      const float nbetaf = -fmaf(ofz, dhz, fmaf(ofy, dhy, ofx * dhx));
      const float kqf = fmaf(ofz, ofz, fmaf(ofy, ofy, ofx * ofx)) * -(1.0f - 96.0f * kEps32);
      unsigned blkany = 0;  // which blocks have a survivor (block j at bit 32 - nconst + j)
      if (!(flags & F_NO_CULL)) {
        const f32x2 nbeta = splat2(nbetaf), kq = splat2(kqf);
        const f32x2 o2x = splat2(2.0f * ofx), o2y = splat2(2.0f * ofy), o2z = splat2(2.0f * ofz);
        const f32x2 dx2 = splat2(dhx), dy2 = splat2(dhy), dz2 = splat2(dhz);
        const int nhb = P.nconst;
        // every block's 16 sign bits are stored unconditionally (warp-uniform index): that keeps the
        // table loads uniform (LDCU -> UR operands), and the loop free of branches and lists
#pragma unroll 1
        for (int ub = 0; ub < nhb; ++ub) {
          unsigned acc = 0xffffffffu;
#pragma unroll
          for (int p = 0; p < kCBP; ++p) {
            const uint4 u = P.ctab[2 * (ub * kCBP + p)], v = P.ctab[2 * (ub * kCBP + p) + 1];
            const f32x2 cx = ((f32x2)u.y << 32) | u.x, cy = ((f32x2)u.w << 32) | u.z;
            const f32x2 cz = ((f32x2)v.y << 32) | v.x, rs = ((f32x2)v.w << 32) | v.z;
            const f32x2 bb = fma2(cz, dz2, fma2(cy, dy2, fma2(cx, dx2, nbeta)));
            const f32x2 ss = fma2(cz, o2z, fma2(cy, o2y, fma2(cx, o2x, add2(rs, kq))));
            const f32x2 dd = fma2(bb, bb, ss);
            acc = __funnelshift_l((unsigned)dd, acc, 1);
            acc = __funnelshift_l((unsigned)(dd >> 32), acc, 1);
          }
          my_mask[ub * kWT] = (unsigned short)acc;  // sphere s of the block -> bit 15 - s, clear = survivor
          blkany = (blkany >> 1) | (acc != 0xffffffffu ? 0x80000000u : 0u);
        }
      }
      int best = -1;
      double closest = __longlong_as_double(0x7ff0000000000000LL);
      if (live) {
        // ---- exact closest hit over the survivors (hit-anything, raytracing.clj:33-43), in the
        // order-independent form: lexicographic minimum of (root, list index)
        const double a = lensq(D);
        if (scan_all) {
#pragma unroll 1
          for (int i = 0; i < P.n; ++i) { const HitPick hp = exact_test_ni(P.geom64, i, O, D, a, closest, best); closest = hp.closest; best = hp.best; }
          n_exact += (unsigned)P.n;
        } else {
          const float tmin_lo = 1e-3f * len32 * (1.0f - 16.0f * kEps32);
          const double ya = recip_refined(a);
          const bool a_ok = recip_safe(a);
          int c1 = -1, c2 = -1, c3 = -1;
          float lo1 = 3.0e38f, lo2 = 3.0e38f, lo3 = 3.0e38f;
          unsigned cur = 0, any = blkany;
          int bbase = 0;
          const int nb_shift = 32 - P.nconst;
#pragma unroll 1
          for (;;) {
            if (cur == 0) {
              if (any == 0) break;
              const int j = (__ffs(any) - 1) - nb_shift;
              any &= any - 1;
              cur = (unsigned)(unsigned short)~my_mask[j * kWT];
              n_pref += (unsigned)__popc(cur);
              bbase = j * 16 - 16;  // __clz counts the 16 leading zeros too
            }
            const int bit = __clz(cur);
            cur &= ~(0x80000000u >> bit);
            int i = bbase + bit;
            if (i >= P.n) continue;
            const float4 g4 = __ldg(P.geomA + i);
            const float cx = g4.x, cy = g4.y, cz = g4.z, ws = g4.w;
            const float bb = fmaf(cz, dhz, fmaf(cy, dhy, fmaf(cx, dhx, nbetaf)));
            const float ss = fmaf(cz, 2.0f * ofz, fmaf(cy, 2.0f * ofy, fmaf(cx, 2.0f * ofx, ws + kqf)));
            const float dd = fmaf(bb, bb, ss);                           // >= D_true (inflated)
            const float sq = sqrt_approx(fmaxf(dd, 0.0f)) * (1.0f + 16.0f * kEps32);
            const float eb = kEps32 * (24.0f * (fabsf(cx) + fabsf(cy) + fabsf(cz)) + 40.0f * mo);
            const float far_hi = bb + sq + eb;
            float lo = bb - sq - eb;                                     // <= every root of sphere i
            const float clo_hi = __double2float_ru(closest) * len32 * (1.0f + 16.0f * kEps32);
            if (far_hi < tmin_lo || lo > clo_hi) continue;
            // keep the three candidates with the smallest lower bounds, sorted; a fourth is tested on the spot
            if (lo < lo1) { const int ti = c1; const float tl = lo1; c1 = i; lo1 = lo; i = ti; lo = tl; }
            if (i >= 0 && lo < lo2) { const int ti = c2; const float tl = lo2; c2 = i; lo2 = lo; i = ti; lo = tl; }
            if (i >= 0 && lo < lo3) { const int ti = c3; const float tl = lo3; c3 = i; lo3 = lo; i = ti; lo = tl; }
            if (i >= 0) {  // (very rare)
              const HitPick hp = exact_test_lex_ni(P.geom64, i, O, D, a, ya, a_ok, closest, best);
              closest = hp.closest; best = hp.best; n_exact++;
            }
          }
          // ONE inlined exact test site: the likeliest winner first, the others only while their lower
          // bound still allows them to win
#pragma unroll 1
          for (int s2 = 0; s2 < 3; ++s2) {
            const int ci = s2 == 0 ? c1 : (s2 == 1 ? c2 : c3);
            const float lo_i = s2 == 0 ? lo1 : (s2 == 1 ? lo2 : lo3);
            if (ci < 0) break;
            if (s2 && !(lo_i <= __double2float_ru(closest) * len32 * (1.0f + 16.0f * kEps32))) break;
            exact_test_lex(P.geom64, ci, O, D, a, ya, a_ok, closest, best);
            n_exact++;
          }
        }
        n_seg++;
      }
      // ---- classify: which phase shades this hit
      if (live) {
        int kind = K_MISS;
        dest = Q_FIN;
        if (best >= 0) {
          const d3 Pt = add(O, muls(D, closest));  // ray/at, ray.clj:7-8
          if (flags & F_NORMAL_SHADING) {          // raytracing_i.clj:62-66: colour = (N + 1) / 2
            const FaceNormal fn = face_normal(P.geom64, best, Pt, D);
            const d3 col = muls(add(fn.n, mk(1.0, 1.0, 1.0)), 0.5);
            sDx[s] = col.x; sDy[s] = col.y; sDz[s] = col.z;  // a finished sample carries its colour in D
            kind = K_NORMAL;
          } else {
            kind = P.mat[best].kind;
            // a hit with one segment left ends black (raytracing.clj:46-47)
            if (sDepth[s] <= 1u) kind = K_END;
            else { sOx[s] = Pt.x; sOy[s] = Pt.y; sOz[s] = Pt.z; dest = kind == K_DIELECTRIC ? Q_DIEL : Q_UNITV; }
          }
        }
        sMeta[s] = (((unsigned)kind & 0xffu) << 16) | ((unsigned)best & 0xffffu);
      }
    } else if (q != Q_FIN) {
      // =========================================================== scatter: Q_UNITV lambertian / metal
      // (material.clj:13-28), Q_DIEL dielectric with Schlick (material.clj:30-46); one queue per warp,
      // so the material branch below is warp-uniform
      const bool wants_unit = q == Q_UNITV;
      if (live) {
        const d3 Pt = mk(sOx[s], sOy[s], sOz[s]), D = mk(sDx[s], sDy[s], sDz[s]);
        const unsigned meta = sMeta[s];
        const int best = (int)(meta & 0xffffu);
        const int kind = (int)(signed char)(meta >> 16);
        const unsigned pixel = sPixel[s], k = sK[s], depth = sDepth[s];
        const FaceNormal fn = face_normal(P.geom64, best, Pt, D);
        const d3 N = fn.n;
        const unsigned stage = (unsigned)P.max_depth - depth + 1u;  // this is the stage-th scatter of the path
        // block 0 of the stage: unit-vector candidates 0 and 1, or the Schlick uniform (a dielectric draws
        // it only when refraction is possible -- `or`, material.clj:42 -- but the stream is counter-based,
        // so drawing it anyway is unobservable)
        uint4 w = philox_ni(pixel, k, stage, 0u, P.k0, P.k1);
        double cx = D.x, cy = D.y, cz = D.z, l2;
        if (wants_unit) {
          // vec3a/random-unit-vec3 (vec3a.clj:74-79): rejection sampling; block n holds candidates
          // 2n (words 0,1) and 2n+1 (words 2,3), 3 x 21 bits each
          unsigned block = 0;
          int half = 0;
#pragma unroll 1
          for (;;) {
            const unsigned wa = half ? w.z : w.x, wb = half ? w.w : w.y;
            cx = sym21(wa & 0x1fffffu);
            cy = sym21((wa >> 21) | ((wb & 0x3ffu) << 11));
            cz = sym21((wb >> 10) & 0x1fffffu);
            l2 = cx * cx + cy * cy + cz * cz;
            if ((l2 > 1e-160 && l2 <= 1.0) || block == 0xffffffu) break;
            if (half == 0) { half = 1; continue; }
            half = 0;
            w = philox_ni(pixel, k, stage, ++block, P.k0, P.k1);
          }
        } else {
          l2 = lensq(D);
        }
        // one sqrt and one 3-way divide serve every material: unit candidate / |d| normalisation
        const d3 U = divs_by_ni(mk(cx, cy, cz), dsqrt(l2));
        const MatRec* m = P.mat + best;
        d3 Dn;
        bool absorbed = false;
        if (!wants_unit) {
          // albedo[] of a dielectric record holds host-precomputed 1/ior and the two Schlick
          // ratios (1-ri)/(1+ri) for ri = 1/ior and ri = ior (same IEEE operations, done once)
          const double ri = fn.front ? m->albedo[0] : m->param;
          const double cos_t = jmin1(dot(neg(U), N));
          const double sin_t = dsqrt(1.0 - cos_t * cos_t);
          bool do_reflect = ri * sin_t > 1.0;
          if (!do_reflect && (flags & F_SCHLICK)) {
            const double q2 = fn.front ? m->albedo[1] : m->albedo[2];  // material/reflectance, material.clj:30-32
            const double r0 = q2 * q2;
            const double mm = 1.0 - cos_t;
            const double m2 = mm * mm;
            const double m5 = m2 * m2 * mm;
            do_reflect = (r0 + (1.0 - r0) * m5) > u24(w.x);
          }
          if (do_reflect) {  // vec3a/reflect, vec3a.clj:94-95
            Dn = sub(U, muls(N, 2.0 * dot(U, N)));
          } else {           // vec3a/refract, vec3a.clj:97-101
            const d3 perp = muls(add(U, muls(N, cos_t)), ri);
            const d3 para = muls(N, -dsqrt(fabs(1.0 - lensq(perp))));
            Dn = add(perp, para);
          }
        } else if (kind == K_LAMBERTIAN) {  // material.clj:13-19, realm/raytracing.clj:138-145
          Dn = add(U, N);
          if ((flags & F_NEAR_ZERO_GUARD) && fabs(Dn.x) < 1e-8 && fabs(Dn.y) < 1e-8 && fabs(Dn.z) < 1e-8) Dn = N;
        } else {                            // material.clj:21-28, realm/raytracing.clj:147-158
          d3 refl = sub(D, muls(N, 2.0 * dot(D, N)));  // vec3a/reflect on the un-normalised direction
          refl = add(muls(U, m->param), refl);
          absorbed = !(dot(refl, N) > 0.0);            // absorbed -> black
          Dn = refl;
        }
        if (absorbed) {
          sMeta[s] = (((unsigned)K_END & 0xffu) << 16) | (unsigned)best;
          dest = Q_FIN;
        } else {
          if (wants_unit) {
            // both product orders (raytracing.clj:52-53 innermost-first, realm/raytracing.clj:225,236
            // forward) are formed at the end of the path from this list of attenuating hits
            const unsigned nst = sNstack[s];
            __stcg(P.stack + (size_t)nst * P.stack_stride + gslot0 + (size_t)s, (unsigned short)best);
            sNstack[s] = nst + 1u;
          }
          sDx[s] = Dn.x; sDy[s] = Dn.y; sDz[s] = Dn.z;
          sDepth[s] = depth - 1u;
          dest = Q_TRACE;
        }
      }
    } else {
      // =========================================================== end of a sample, next camera ray
      unsigned pixel = 0, k = 0, kend = 0, unit = 0;
      bool need_unit = false;
      if (live) {
        const int kind = (int)(signed char)(sMeta[s] >> 16);
        need_unit = kind == K_FRESH;
        if (!need_unit) {
          pixel = sPixel[s]; k = sK[s]; kend = sKend[s]; unit = sUnit[s];
          d3 color = mk(0.0, 0.0, 0.0);
          if (kind == K_MISS) {  // sky, raytracing.clj:55-58 / realm/raytracing.clj:229-236
            const d3 D = mk(sDx[s], sDy[s], sDz[s]);
            const double len = dsqrt(lensq(D));
            const double uy = div_by(D.y, len, recip_refined(len), recip_safe(len));  // (unit-vector d).y
            const double g = 0.5 * (uy + 1.0);
            const d3 sky = mk((1.0 - g) * 1.0 + g * 0.5, (1.0 - g) * 1.0 + g * 0.7, (1.0 - g) * 1.0 + g * 1.0);
            const unsigned nst = sNstack[s];
            const unsigned short* col = P.stack + gslot0 + (size_t)s;
            color = attenuate(P, col, nst, sky);
          } else if (kind == K_NORMAL) {
            color = mk(sDx[s], sDy[s], sDz[s]);
          }
          const double sr = sSr[s] + color.x, sg = sSg[s] + color.y, sb = sSb[s] + color.z;  // raytracing.clj:153
          if (++k == kend) {
            need_unit = true;
            if (P.out_linear || P.out_rgb8) {
              if (P.nchunks == 1) {
                finish_pixel(P, unit, 0.0 + sr, 0.0 + sg, 0.0 + sb);
              } else {
                finish_chunked(P, unit, sr, sg, sb);
              }
            }
          } else {
            sSr[s] = sr; sSg[s] = sg; sSb[s] = sb;
          }
        }
      }
      // ---- new work units: ballot-compacted tickets from the global queue (replaces the
      // reference's row-chunk pool, raytracing.clj:157-171)
      bool died = false;
      {
        const unsigned mask = __ballot_sync(FULL, need_unit);
        if (mask) {
          const int leader = __ffs(mask) - 1;
          unsigned long long tb = 0;
          if (lane == leader) tb = atomicAdd(P.queue, (unsigned long long)__popc(mask));
          tb = __shfl_sync(FULL, tb, leader);
          if (need_unit) {
            const unsigned long long ticket = tb + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
            if (ticket >= P.total_units) {
              died = true;
            } else {
              unit = (unsigned)ticket;
              const unsigned p_local = unit / (unsigned)P.nchunks;
              const int chunk = (int)(unit - p_local * (unsigned)P.nchunks);
              const int lr = (int)(p_local / (unsigned)P.W);
              const int pi = (int)(p_local - (unsigned)lr * (unsigned)P.W);
              const int tile = lr / P.shard_rows;
              const int pj = (tile * P.shard_count + P.shard_index) * P.shard_rows + (lr - tile * P.shard_rows);
              pixel = (unsigned)pj * (unsigned)P.W + (unsigned)pi;
              k = (unsigned)(chunk * P.spu);
              kend = (unsigned)min((int)k + P.spu, P.spp);
              sPixel[s] = pixel; sKend[s] = kend; sUnit[s] = unit;
              sSr[s] = 0.0; sSg[s] = 0.0; sSb[s] = 0.0;
            }
          }
        }
        const unsigned dm = __ballot_sync(FULL, died);
        if (dm && lane == 0) atomicAdd(const_cast<unsigned*>(ctl) + 8, (unsigned)__popc(dm));
      }
      // ---- camera ray: raytracing.clj:144-151, realm/raytracing.clj:332-339
      if (live && !died) {
        const unsigned pj = pixel / (unsigned)P.W, pi = pixel - pj * (unsigned)P.W;
        uint4 w = philox_ni(pixel, k, 0u, 0u, P.k0, P.k1);
        const double sx = (double)pi + (u24(w.x) - 0.5);
        const double sy = (double)pj + (u24(w.y) - 0.5);
        const d3 ps = add(add(ld3(P.p00), muls(ld3(P.du), sx)), muls(ld3(P.dv), sy));
        d3 O = ld3(P.center);
        if (P.use_defocus) {  // vec3a/random-in-unit-disk, vec3a.clj:81-86
          double px = sym24(w.z), py = sym24(w.w);
          unsigned block = 0;
          int half = 1;
          while (!(px * px + py * py < 1.0) && block < 0xffffffu) {
            if (half == 1) { w = philox_ni(pixel, k, 0u, ++block, P.k0, P.k1); half = 0; } else half = 1;
            px = sym24(half ? w.z : w.x);
            py = sym24(half ? w.w : w.y);
          }
          O = add(add(O, muls(ld3(P.ddu), px)), muls(ld3(P.ddv), py));  // raytracing.clj:89-93
        }
        const d3 D = sub(ps, O);
        sOx[s] = O.x; sOy[s] = O.y; sOz[s] = O.z;
        sDx[s] = D.x; sDy[s] = D.y; sDz[s] = D.z;
        sK[s] = k;
        sDepth[s] = (unsigned)P.max_depth;
        sNstack[s] = 0u;
        n_samples++;
        dest = Q_TRACE;
      }
    }
    wave_push(ctl, qbuf, dest, slot, lane);
  }


  // ---- counters: REDUX on 16-bit halves (each lane's count fits 32 bits), one atomic per warp
  {
    const unsigned v[5] = {n_samples, n_seg, n_exact, 0u, n_pref};
#pragma unroll 1
    for (int qq = 0; qq < 5; ++qq) {
      const unsigned lo = __reduce_add_sync(FULL, v[qq] & 0xffffu), hi = __reduce_add_sync(FULL, v[qq] >> 16);
      if (lane == 0) atomicAdd(P.stats + qq, (unsigned long long)lo + ((unsigned long long)hi << 16));
    }
  }
}

}  // namespace rtclj
