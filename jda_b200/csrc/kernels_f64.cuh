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
  // face.similarity_transform (data.cpp:64-126) and the initial shift DataSet::RandomShape adds (data.cpp:225-236)
  int similarity;
  double shift_x, shift_y;
};

// STParameter (include/jda/data.hpp:18-50); the default is the identity
struct STP64 {
  double scale, r00, r01, r10, r11;
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

// STParameter::Apply (data.hpp:42-45)
__device__ __forceinline__ void stp_apply(const STP64 &p, double x1, double y1, double &x2, double &y2) {
  x2 = __dmul_rn(p.scale, __dadd_rn(__dmul_rn(p.r00, x1), __dmul_rn(p.r01, y1)));
  y2 = __dmul_rn(p.scale, __dadd_rn(__dmul_rn(p.r10, x1), __dmul_rn(p.r11, y1)));
}

// STParameter::Calc(shape, mean_shape) (data.cpp:64-114), every lane of the warp for itself (the sums run over the
// landmarks in index order: there is nothing to split).  The centred / normalised copies the reference keeps in two
// temporaries are recomputed where they are used -- the same operations on the same values.  cv::norm = the square
// root of the squares summed in index order (what the pinned reference build's stand-in does; OpenCV's own
// accumulation order is third-party code that is not under /root/reference).
__device__ __forceinline__ STP64 stp_calc(const double *s1, const double *s2, int L) {
  double x1c = 0., y1c = 0., x2c = 0., y2c = 0.;
  for (int i = 0; i < L; i++) {
    x1c = __dadd_rn(x1c, s1[2 * i]); y1c = __dadd_rn(y1c, s1[2 * i + 1]);
    x2c = __dadd_rn(x2c, __ldg(s2 + 2 * i)); y2c = __dadd_rn(y2c, __ldg(s2 + 2 * i + 1));
  }
  const double fl = (double)L;
  x1c = __ddiv_rn(x1c, fl); y1c = __ddiv_rn(y1c, fl); x2c = __ddiv_rn(x2c, fl); y2c = __ddiv_rn(y2c, fl);
  double q1 = 0., q2 = 0.;
  for (int i = 0; i < L; i++) {
    const double ax = __dsub_rn(s1[2 * i], x1c), ay = __dsub_rn(s1[2 * i + 1], y1c);
    const double bx = __dsub_rn(__ldg(s2 + 2 * i), x2c), by = __dsub_rn(__ldg(s2 + 2 * i + 1), y2c);
    q1 = __dadd_rn(q1, __dmul_rn(ax, ax)); q1 = __dadd_rn(q1, __dmul_rn(ay, ay));
    q2 = __dadd_rn(q2, __dmul_rn(bx, bx)); q2 = __dadd_rn(q2, __dmul_rn(by, by));
  }
  const double scale1 = __dsqrt_rn(q1), scale2 = __dsqrt_rn(q2);
  STP64 p;
  p.scale = __ddiv_rn(scale1, scale2);
  double num = 0., den = 0.;
  for (int i = 0; i < L; i++) {
    const double ax = __ddiv_rn(__dsub_rn(s1[2 * i], x1c), scale1), ay = __ddiv_rn(__dsub_rn(s1[2 * i + 1], y1c), scale1);
    const double bx = __ddiv_rn(__dsub_rn(__ldg(s2 + 2 * i), x2c), scale2), by = __ddiv_rn(__dsub_rn(__ldg(s2 + 2 * i + 1), y2c), scale2);
    num = __dadd_rn(num, __dsub_rn(__dmul_rn(ay, bx), __dmul_rn(ax, by)));
    den = __dadd_rn(den, __dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)));
  }
  const double norm = __dsqrt_rn(__dadd_rn(__dmul_rn(num, num), __dmul_rn(den, den)));
  const double sin_theta = __ddiv_rn(num, norm), cos_theta = __ddiv_rn(den, norm);
  p.r00 = cos_theta; p.r01 = -sin_theta; p.r10 = sin_theta; p.r11 = cos_theta;
  return p;
}

// SIM: face.similarity_transform (P.similarity); the shipped configuration runs the instantiation without it
template <bool SIM>
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
    // RandomShape: mean + (x, y); src/test.cpp:17,75 run with a zero shift
    for (int i = lane; i < D; i += 32) shape[i] = __dadd_rn(P.mean_shape[i], (i & 1) ? P.shift_y : P.shift_x);
    __syncwarp();
    [[maybe_unused]] STP64 stp;  // STParameter stp_mc; (cascador.cpp:176)
    stp.scale = 1.0; stp.r00 = 1.0; stp.r01 = 0.0; stp.r10 = 0.0; stp.r11 = 1.0;

    double score = 0.0;
    int n_eval = 0;
    bool rejected = false;
    // full stages, then the carts of an unfinished stage (training snapshots); a finished model has none
    const int n_pass = P.stage + (P.cart_last >= 0 ? 1 : 0);
    for (int t = 0; t < n_pass && !rejected; t++) {
      const bool full = t < P.stage;
      const int kend = full ? P.K : P.cart_last + 1;
      // cascador.cpp:180 -- the unfinished stage keeps the transform of the last finished one (cascador.cpp:199-202)
      if (SIM && full) stp = stp_calc(shape, P.mean_shape, P.L);
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
            double2 o1 = __ldg(reinterpret_cast<const double2 *>(n) + 1);  // o1x, o1y
            double2 o2 = __ldg(reinterpret_cast<const double2 *>(n) + 2);  // o2x, o2y
            if (SIM) {  // stp_mc.Apply(offset), data.cpp:43-44 (the identity returns its input exactly)
              stp_apply(stp, o1.x, o1.y, o1.x, o1.y);
              stp_apply(stp, o2.x, o2.y, o2.x, o2.y);
            }
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
      // btcart.cpp:407-424: delta = sum of the K selected rows, accumulated from zero in cart order; the transform is
      // applied to it landmark by landmark; shape += delta.  A lane owns a landmark (both coordinates).
      const double *wt = P.w + (size_t)t * P.K * kLeaves * D;
      for (int j = lane; j < P.L; j += 32) {
        double dx = 0.0, dy = 0.0;
        for (int k0 = 0; k0 < P.K; k0 += 4) {
          double2 v[4];  // (four 16-byte row loads in flight: the bytes of round 1's eight 8-byte ones)
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const int k = min(k0 + u, P.K - 1);
            v[u] = __ldg(reinterpret_cast<const double2 *>(wt + (size_t)(k * kLeaves + leafs[k]) * D) + j);
          }
#pragma unroll
          for (int u = 0; u < 4; u++)
            if (k0 + u < P.K) { dx = __dadd_rn(dx, v[u].x); dy = __dadd_rn(dy, v[u].y); }
        }
        if (SIM) stp_apply(stp, dx, dy, dx, dy);
        shape[2 * j] = __dadd_rn(shape[2 * j], dx);
        shape[2 * j + 1] = __dadd_rn(shape[2 * j + 1], dy);
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
