// rtclj_split_kernel.cuh -- the render loop (src/raytracing.clj:141-171) with the closest-hit CULL taken out
// of the path warps: per SM sub-partition ONE warp only culls (four rays per lane, so every uniform load of a
// sphere pair feeds four rays: 91 instead of 123 instructions per 16 spheres x 32 rays, and one warp alone
// keeps the FMA pipe ~80 % busy -- tools/microbench/cull_loop8), and four warps only resolve / shade / sample.
//
// Why: in the kernels where every warp does everything (rtclj_kernels.cuh, rtclj_lane2_kernel.cuh) the FMA pipe
// is 64 % busy and the issue slots 76 %: the warps of a sub-partition drift into the same phase -- all shading
// (the pipe idles) or several culling (they fight for it).  A dedicated cull warp should feed the pipe evenly.
//
// MEASURED: bit-exact, and SLOWER -- 627 ms per bench frame against 536 ms for
// render_lane2_kernel (FMA pipe 56 % busy, issue slots 72 %).  Sixteen path warps do not hide the latency of
// the fp64 resolve / shade chains; the four warps given to the cull are missed there.  Kept behind
// RTCLJ_F_SPLIT_KERNEL as the measured answer to "why not warp-specialise".
//
// A path warp owns two paths per lane ("sets" 0 and 1), like render_lane2_kernel.  While it resolves and
// shades one set, the other set's rays are with the cull warp: the worker writes the eight fp32 cull operands
// of each ray to a mailbox, flags the set SUBMITTED, and picks up the other set once its flag says MASKS.  The
// cull warp takes up to four submitted sets of its sub-partition's workers per pass.  Everything a path does
// between two culls is path_step() (rtclj_path_step.cuh), shared with render_lane2_kernel: same arithmetic,
// same draw order, same results.
#pragma once
#include "rtclj_path_step.cuh"

namespace rtclj {

constexpr int kSplitThreads = 640;                 // 4 cull warps + 16 path warps
constexpr int kSplitCullWarps = 4;
constexpr int kSplitW = kSplitThreads - 32 * kSplitCullWarps;  // path lanes per CTA (512)
enum { SS_MASKS = 0, SS_SUBMITTED = 1, SS_DEAD = 2 };          // state of a (path warp, set)

// dynamic shared memory layout (bytes)
struct SplitSmem {
  static constexpr size_t state = 0;                                          // 16 warps x 2 sets words (+ pad)
  static constexpr size_t masks = 256;                                        // [set][block 32][lane] u16
  static constexpr size_t views = masks + (size_t)2 * 32 * kSplitW * 2;       // [set][8][lane] float
  static constexpr size_t bany = views + (size_t)2 * 8 * kSplitW * 4;         // [set][lane] u32
  static constexpr size_t sums = bany + (size_t)2 * kSplitW * 4;              // [set][3][lane] double
  static constexpr size_t f64 = sums + (size_t)2 * 3 * kSplitW * 8;           // [6][lane] double: O, D of the waiting path
  static constexpr size_t u32 = f64 + (size_t)6 * kSplitW * 8;                // [7][lane] words: its bookkeeping
  static constexpr size_t total = u32 + (size_t)7 * kSplitW * 4;
};

template <bool kSampleBuf>
__global__ void __launch_bounds__(kSplitThreads, 1) render_split_kernel(const __grid_constant__ KParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  volatile unsigned* const state = reinterpret_cast<volatile unsigned*>(smem_raw + SplitSmem::state);
  unsigned short* const masks = reinterpret_cast<unsigned short*>(smem_raw + SplitSmem::masks);
  float* const views = reinterpret_cast<float*>(smem_raw + SplitSmem::views);
  unsigned* const banys = reinterpret_cast<unsigned*>(smem_raw + SplitSmem::bany);
  const unsigned FULL = 0xffffffffu;

  if (tid < 2 * (kSplitThreads / 32 - kSplitCullWarps)) state[tid] = SS_MASKS;  // fresh paths need no masks
  __syncthreads();

  if (warp < kSplitCullWarps) {
    // ================================================================ cull warp of sub-partition `warp`
    // serves path warps w = warp, warp + 4, ... (the same hardware scheduler), sets 0 and 1 of each: 8 pairs
    for (;;) {
      unsigned st = SS_DEAD;
      int pair_flag = 0;
      if (lane < 8) { pair_flag = 2 * (warp + kSplitCullWarps * (lane >> 1)) + (lane & 1); st = state[pair_flag]; }
      const unsigned ready = __ballot_sync(FULL, st == SS_SUBMITTED) & 0xffu;
      // (ptxas keeps the cull's table loads uniform only with the loop exit nested in this ONE branch: a second
      // `continue`, or the exit test at the top of the loop, silently turns them into per-lane LDC -- 798 instead
      // of 627 ms per bench frame; tests/test_build_artifacts.py checks the SASS)
      // whatever is there, up to four sets (waiting for FULL passes -- unless a path warp has both its sets here --
      // was measured too: 627.8 against 627.4 ms per bench frame, no difference)
      const bool go = ready != 0u;
      if (!go) {
        if (__all_sync(FULL, st == SS_DEAD)) break;
        __nanosleep(32);
        continue;
      }
      __syncwarp();
      __threadfence_block();  // the flags before the views
      // up to four submitted pairs; missing ones repeat the first (their results are not written)
      int fl[4];      // flag index of the pair = 2 * path warp + set
      bool ok[4];
      {
        unsigned m = ready;
        const int first = __ffs(m) - 1;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          ok[r] = m != 0u;
          const int j = ok[r] ? __ffs(m) - 1 : first;
          m &= m - 1u;
          fl[r] = __shfl_sync(FULL, pair_flag, j);
        }
      }
      f32x2 nb[4], kq[4], ox[4], oy[4], oz[4], dx[4], dy[4], dz[4];
      unsigned short* mrow[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int set = fl[r] & 1, wl = (fl[r] >> 1) * 32 + lane;
        const float* v = views + (size_t)set * 8 * kSplitW + wl;
        nb[r] = splat2(v[0]); kq[r] = splat2(v[kSplitW]);
        ox[r] = splat2(v[2 * kSplitW]); oy[r] = splat2(v[3 * kSplitW]); oz[r] = splat2(v[4 * kSplitW]);
        dx[r] = splat2(v[5 * kSplitW]); dy[r] = splat2(v[6 * kSplitW]); dz[r] = splat2(v[7 * kSplitW]);
        mrow[r] = masks + (size_t)set * 32 * kSplitW + wl;
      }
      unsigned ba[4] = {0u, 0u, 0u, 0u};
      const int nhb = P.nconst;
#pragma unroll 1
      for (int ub = 0; ub < nhb; ++ub) {
        unsigned acc[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll
        for (int p = 0; p < kCBP; ++p) {
          const uint4 u = P.ctab[2 * (ub * kCBP + p)], v = P.ctab[2 * (ub * kCBP + p) + 1];
          const f32x2 cx = ((f32x2)u.y << 32) | u.x, cy = ((f32x2)u.w << 32) | u.z;
          const f32x2 cz = ((f32x2)v.y << 32) | v.x, rs = ((f32x2)v.w << 32) | v.z;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const f32x2 bb = fma2(cz, dz[r], fma2(cy, dy[r], fma2(cx, dx[r], nb[r])));
            const f32x2 ss = fma2(cz, oz[r], fma2(cy, oy[r], fma2(cx, ox[r], add2(rs, kq[r]))));
            const f32x2 dd = fma2(bb, bb, ss);
            acc[r] = __funnelshift_l((unsigned)dd, acc[r], 1);
            acc[r] = __funnelshift_l((unsigned)(dd >> 32), acc[r], 1);
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          if (ok[r]) mrow[r][ub * kSplitW] = (unsigned short)acc[r];  // warp-uniform predicate and row
          ba[r] = (ba[r] >> 1) | (acc[r] != 0xffffffffu ? 0x80000000u : 0u);
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (ok[r]) banys[(size_t)(fl[r] & 1) * kSplitW + (fl[r] >> 1) * 32 + lane] = ba[r];
      __threadfence_block();  // the masks before the flags
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (ok[r] && lane == r) state[fl[r]] = SS_MASKS;
    }
    return;
  }

  // ================================================================== path warp
  const int ww = warp - kSplitCullWarps;      // 0..15
  const int wl = tid - 32 * kSplitCullWarps;  // 0..511
  const unsigned gl = blockIdx.x * kSplitW + wl;
  double* const my_sums = reinterpret_cast<double*>(smem_raw + SplitSmem::sums) + wl;   // [(set * 3 + c) * kSplitW]
  double* const wO = reinterpret_cast<double*>(smem_raw + SplitSmem::f64) + wl;         // [c * kSplitW]
  unsigned* const wU = reinterpret_cast<unsigned*>(smem_raw + SplitSmem::u32) + wl;     // [c * kSplitW]
  PathRegs pr;
  pr.O = mk(0.0, 0.0, 0.0); pr.D = mk(0.0, 0.0, 1.0);
  pr.pixel = 0; pr.unit = 0; pr.k = 0; pr.k_end = 0; pr.depth_left = 0; pr.nstack = 0; pr.status = PS_FRESH;
  PathCounters pc = {0u, 0u, 0u, 0u};
  wU[6 * kSplitW] = (unsigned)PS_FRESH;  // the waiting path starts FRESH as well
  bool dead0 = false, dead1 = false;     // warp-uniform: every lane's path of the set has run out of work
  int set = 0;
  for (;;) {
    const bool set_dead = set ? dead1 : dead0;
    if (!set_dead) {
      // ---- the cull warp's answer for this set
      while (__any_sync(FULL, state[2 * ww + set] != SS_MASKS)) __nanosleep(32);
      __syncwarp();
      __threadfence_block();  // the flag before the masks
      const unsigned short* mrow = masks + (size_t)set * 32 * kSplitW + wl;
      path_step<kSampleBuf, kSplitW>(P, pr, pc, banys[(size_t)set * kSplitW + wl],
                                     [&](int j) { return ~(unsigned)mrow[j * kSplitW] & 0xffffu; },
                                     my_sums + set * 3 * kSplitW, (size_t)gl * 2u + (size_t)set, lane);
      // ---- hand the new ray to the cull warp
      {
        RayView v = make_view(P, pr.O, pr.D);
        float* o = views + (size_t)set * 8 * kSplitW + wl;
        o[0] = v.nbetaf; o[kSplitW] = v.kqf;
        o[2 * kSplitW] = 2.0f * v.ofx; o[3 * kSplitW] = 2.0f * v.ofy; o[4 * kSplitW] = 2.0f * v.ofz;
        o[5 * kSplitW] = v.dhx; o[6 * kSplitW] = v.dhy; o[7 * kSplitW] = v.dhz;
      }
      const bool all_dead = !__any_sync(FULL, pr.status != PS_DEAD);
      __threadfence_block();  // the views before the flag
      __syncwarp();
      if (lane == 0) state[2 * ww + set] = all_dead ? SS_DEAD : SS_SUBMITTED;
      if (set) dead1 = all_dead; else dead0 = all_dead;
    }
    // ---- swap: this path waits in shared memory while the other set is processed
    {
      const d3 tO = mk(wO[0], wO[kSplitW], wO[2 * kSplitW]), tD = mk(wO[3 * kSplitW], wO[4 * kSplitW], wO[5 * kSplitW]);
      const unsigned t_pixel = wU[0], t_unit = wU[kSplitW], t_k = wU[2 * kSplitW], t_kend = wU[3 * kSplitW];
      const unsigned t_depth = wU[4 * kSplitW], t_nstack = wU[5 * kSplitW], t_status = wU[6 * kSplitW];
      wO[0] = pr.O.x; wO[kSplitW] = pr.O.y; wO[2 * kSplitW] = pr.O.z;
      wO[3 * kSplitW] = pr.D.x; wO[4 * kSplitW] = pr.D.y; wO[5 * kSplitW] = pr.D.z;
      wU[0] = pr.pixel; wU[kSplitW] = pr.unit; wU[2 * kSplitW] = (unsigned)pr.k; wU[3 * kSplitW] = (unsigned)pr.k_end;
      wU[4 * kSplitW] = (unsigned)pr.depth_left; wU[5 * kSplitW] = (unsigned)pr.nstack; wU[6 * kSplitW] = (unsigned)pr.status;
      pr.O = tO; pr.D = tD;
      pr.pixel = t_pixel; pr.unit = t_unit; pr.k = (int)t_k; pr.k_end = (int)t_kend;
      pr.depth_left = (int)t_depth; pr.nstack = (int)t_nstack; pr.status = (int)t_status;
    }
    set ^= 1;
    if (dead0 && dead1) break;
  }

  // ---- counters: REDUX on 16-bit halves (each lane's count fits 32 bits), one atomic per warp
  {
    const unsigned v[5] = {pc.samples, pc.seg, pc.exact, 0u, pc.pref};
#pragma unroll 1
    for (int q = 0; q < 5; ++q) {
      const unsigned lo = __reduce_add_sync(FULL, v[q] & 0xffffu), hi = __reduce_add_sync(FULL, v[q] >> 16);
      if (lane == 0) atomicAdd(P.stats + q, (unsigned long long)lo + ((unsigned long long)hi << 16));
    }
  }
}

}  // namespace rtclj
