// kernels_f64.cuh -- sm_100a kernel of the double-precision detector (JoinCascador::Detect, fddb.method = 1).
//
//   k4_cascade_f64   JoinCascador::Validate for one window per warp            src/jda/cascador.cpp:166-211
//                    (Cart::Forward cart.cpp:392-404, Feature::CalcFeatureValue data.cpp:18-58,
//                     BoostCart::GenDeltaShape btcart.cpp:407-424)
//
// Arithmetic contract: the reference's double operations in the reference's order -- (shape + offset) * width,
// round() (half away from zero), clamp; score += leaf; score = (score - mean) / std; delta accumulated from 0 over
// the K rows in cart order, then shape += delta.  No FMA contraction (explicit __dadd_rn / __dmul_rn / __ddiv_rn).
// Windows come either from k2_scan's survivor queue (a conservative float32 prefilter of stage 0, see
// host_model_f64.hpp: every queued window is re-evaluated here from cart 0) or from a dense enumeration.
#pragma once
#include "kernels.cuh"
#include "host_model_f64.hpp"

namespace jda {

struct Cascade64Params {
  const uint8_t *frames;
  size_t frame_stride;
  int pitch;
  const NodeRecD *nodes;
  const double *leaf;   // [T*K*8]
  const double *cart3;  // [T*K*3] th, mean, std
  const double *w;      // [T][8K][2L]
  const double *mean_shape;
  int T, K, L;
  int stage, cart_last;  // full stages [0, stage), then carts [0, cart_last] of stage `stage` (no regression after them)
  int n_levels;
  int lv_win[kMaxLevels], lv_step[kMaxLevels], lv_nx[kMaxLevels], lv_ny[kMaxLevels];
  long long lv_base[kMaxLevels];
  long long windows_per_frame;
  const uint4 *surv;
  const unsigned *surv_count;
  unsigned surv_cap;
  int dense;  // 1: every window of every frame; 0: the survivor queue
  long long dense_total;
  double *hits;  // records of rec_doubles doubles: {frame | key, x | y, win | carts, score, shape[2L]}
  unsigned *hit_count;
  unsigned hit_cap;
  int rec_doubles;
  unsigned *work_counter;
  int *trace_n;     // dense mode, optional: carts evaluated per window
  double *trace_s;  //                       exit score per window
};

constexpr int kHit64Header = 4;
constexpr int K4_WARPS = 4;
constexpr int K4_G = 2;  // chunks of 32 carts walked together per window

#ifdef __CUDACC__

__device__ __forceinline__ int coord_f64(double s, double o, double fwin, int win) {
  // data.cpp:38-51: (s + offset) * width, round(), clamp to the view
  const double v = __dmul_rn(__dadd_rn(s, o), fwin);
  int c = (int)round(v);
  return min(max(c, 0), win - 1);
}

__global__ void __launch_bounds__(K4_WARPS * 32) k4_cascade_f64(const __grid_constant__ Cascade64Params P) {
  extern __shared__ __align__(16) uint8_t smem4[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = 2 * P.L;
  const int per_warp = kMaxDim * 8 + ((P.K + 15) & ~15);
  double *shape = reinterpret_cast<double *>(smem4 + (size_t)warp * per_warp);
  uint8_t *leafs = reinterpret_cast<uint8_t *>(shape + kMaxDim);

  const long long total = P.dense ? P.dense_total : (long long)min(*P.surv_count, P.surv_cap);
  long long e = (long long)blockIdx.x * K4_WARPS + warp - (long long)gridDim.x * K4_WARPS;
  for (;;) {
    if (P.dense) {
      e += (long long)gridDim.x * K4_WARPS;
    } else {
      unsigned t = 0;
      if (lane == 0) t = atomicAdd(P.work_counter, 1u);
      e = (long long)__shfl_sync(0xffffffffu, t, 0);
    }
    if (e >= total) break;
    int frame, level, xi, yi;
    if (P.dense) {
      frame = (int)(e / P.windows_per_frame);
      long long r = e - (long long)frame * P.windows_per_frame;
      level = 0;
      while (level + 1 < P.n_levels && r >= P.lv_base[level + 1]) level++;
      r -= P.lv_base[level];
      yi = (int)(r / P.lv_nx[level]);
      xi = (int)(r - (long long)yi * P.lv_nx[level]);
    } else {
      const uint4 s = P.surv[e];
      frame = (int)s.x;
      level = (int)(s.y >> 26);
      yi = (int)((s.y >> 13) & 0x1fff);
      xi = (int)(s.y & 0x1fff);
    }
    const int win = P.lv_win[level], step = P.lv_step[level];
    const int x = xi * step, y = yi * step;
    const double fwin = (double)win;
    const uint8_t *po = P.frames + (size_t)frame * P.frame_stride + (size_t)y * P.pitch + x;

    __syncwarp();
    for (int i = lane; i < D; i += 32) shape[i] = __dadd_rn(P.mean_shape[i], 0.0);  // RandomShape with a zero shift
    __syncwarp();

    double score = 0.0;
    int n_eval = 0;
    bool rejected = false;
    // full stages, then the carts of an unfinished stage (training snapshots); a finished model has none
    const int n_pass = P.stage + (P.cart_last >= 0 ? 1 : 0);
    for (int t = 0; t < n_pass && !rejected; t++) {
      const bool full = t < P.stage;
      const int kend = full ? P.K : P.cart_last + 1;
      for (int kc = 0; kc < kend && !rejected; kc += 32 * K4_G) {
        double ls[K4_G], cth[K4_G], cmean[K4_G], cstd[K4_G];
        int idx[K4_G];
        bool ok[K4_G];
        const NodeRecD *nd[K4_G];
#pragma unroll
        for (int g = 0; g < K4_G; g++) {
          const int k = kc + 32 * g + lane;
          ok[g] = k < kend;
          nd[g] = P.nodes + ((size_t)t * P.K + (ok[g] ? k : 0)) * kNodes;
          idx[g] = 0;
          ls[g] = 0.0; cth[g] = 0.0; cmean[g] = 0.0; cstd[g] = 1.0;
        }
#pragma unroll
        for (int lvl = 0; lvl < kDepth - 1; lvl++) {
#pragma unroll
          for (int g = 0; g < K4_G; g++) {
            const NodeRecD *n = nd[g] + idx[g];
            const int4 a = __ldg(reinterpret_cast<const int4 *>(n));            // scale, lm1, lm2, th
            const double2 o1 = __ldg(reinterpret_cast<const double2 *>(n) + 1);  // o1x, o1y
            const double2 o2 = __ldg(reinterpret_cast<const double2 *>(n) + 2);  // o2x, o2y
            const int x1 = coord_f64(shape[a.y], o1.x, fwin, win), y1 = coord_f64(shape[a.y + 1], o1.y, fwin, win);
            const int x2 = coord_f64(shape[a.z], o2.x, fwin, win), y2 = coord_f64(shape[a.z + 1], o2.y, fwin, win);
            const int p1 = __ldg(po + (size_t)y1 * P.pitch + x1);
            const int p2 = __ldg(po + (size_t)y2 * P.pitch + x2);
            idx[g] = (p1 - p2 <= a.w) ? 2 * idx[g] + 1 : 2 * idx[g] + 2;  // cart.cpp:398-399 in 0-based heap indices
          }
        }
#pragma unroll
        for (int g = 0; g < K4_G; g++) {
          if (ok[g]) {
            const int k = kc + 32 * g + lane;
            const size_t c = (size_t)t * P.K + k;
            const int lf = idx[g] - kNodes;
            leafs[k] = (uint8_t)lf;
            ls[g] = __ldg(P.leaf + c * kLeaves + lf);
            cth[g] = __ldg(P.cart3 + c * 3);
            cmean[g] = __ldg(P.cart3 + c * 3 + 1);
            cstd[g] = __ldg(P.cart3 + c * 3 + 2);
          }
        }
        // replay the score in cart order (cascador.cpp:184-193)
#pragma unroll
        for (int g = 0; g < K4_G; g++) {
          const int cnt = min(32, kend - (kc + 32 * g));
          if (cnt <= 0 || rejected) continue;
          const unsigned normed = __ballot_sync(0xffffffffu, cmean[g] != 0.0 || cstd[g] != 1.0);
          for (int j = 0; j < cnt; j++) {
            score = __dadd_rn(score, __shfl_sync(0xffffffffu, ls[g], j));
            if ((normed >> j) & 1u)  // (score - 0) / 1 is score exactly
              score = __ddiv_rn(__dsub_rn(score, __shfl_sync(0xffffffffu, cmean[g], j)), __shfl_sync(0xffffffffu, cstd[g], j));
            n_eval++;
            if (score < __shfl_sync(0xffffffffu, cth[g], j)) { rejected = true; break; }
          }
        }
      }
      if (rejected || !full) break;
      __syncwarp();
      // btcart.cpp:407-424: delta = sum of the K selected rows, accumulated from zero in cart order; shape += delta
      const double *wt = P.w + (size_t)t * P.K * kLeaves * D;
      for (int i = lane; i < D; i += 32) {
        double delta = 0.0;
        for (int k0 = 0; k0 < P.K; k0 += 8) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; u++) {
            const int k = min(k0 + u, P.K - 1);
            v[u] = __ldg(wt + (size_t)(k * kLeaves + leafs[k]) * D + i);
          }
#pragma unroll
          for (int u = 0; u < 8; u++)
            if (k0 + u < P.K) delta = __dadd_rn(delta, v[u]);
        }
        shape[i] = __dadd_rn(shape[i], delta);
      }
      __syncwarp();
    }
    if (P.dense && lane == 0) {
      const long long gw = (long long)frame * P.windows_per_frame + P.lv_base[level] + (long long)yi * P.lv_nx[level] + xi;
      if (P.trace_n) P.trace_n[gw] = n_eval;
      if (P.trace_s) P.trace_s[gw] = score;
    }
    if (rejected) continue;
    unsigned slot = 0;
    if (lane == 0) slot = atomicAdd(P.hit_count, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot < P.hit_cap) {
      double *rec = P.hits + (size_t)slot * P.rec_doubles;
      if (lane == 0) {
        int *ri = reinterpret_cast<int *>(rec);
        ri[0] = frame; ri[1] = (int)pack_key(level, yi, xi);
        ri[2] = x; ri[3] = y;
        ri[4] = win; ri[5] = n_eval;
        rec[3] = score;
      }
      for (int i = lane; i < D; i += 32) rec[kHit64Header + i] = shape[i];
    }
  }
}

#endif  // __CUDACC__
}  // namespace jda
