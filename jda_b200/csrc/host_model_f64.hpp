// host_model_f64.hpp -- host side of the double-precision detector, JoinCascador::Detect with fddb.method = 1
// (the reference's C++ path).  Pure C++, compiled with -ffp-contract=off.
//
// Reference behaviour restated here (file:line under /root/reference):
//   model layout (doubles kept as doubles)  src/jda/cascador.cpp:126-164, src/jda/cart.cpp:406-428
//   window ladder                           src/jda/cascador.cpp:310-376 (detectMultiScale1)
//   node address arithmetic                 src/jda/data.cpp:18-58 (double, round(), clamp) -> stage-0 tables
//   nms                                     src/jda/cascador.cpp:387-429 (multimap by score)
//   relocation                              src/jda/cascador.cpp:462-474
//
// Scope: every node at scale == 0 (the shipped model); other models sample cv::resize'd planes whose arithmetic is
// OpenCV's (third party, unpinned) and are refused.  face.similarity_transform and the initial shift live in the kernel
// (kernels_f64.cuh); the stage-0 tables below are built for the mean shape and the identity transform, which is what
// stage 0 sees whenever the shift is zero.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "host_model.hpp"

namespace jda {

// 48-byte node record as the f64 cascade kernel reads it (three 16-byte loads)
struct alignas(16) NodeRecD {
  int scale;
  int lm1, lm2;  // landmark x index (already * 2)
  int th;
  double o1x, o1y, o2x, o2y;
};
static_assert(sizeof(NodeRecD) == 48, "NodeRecD must be 48 bytes");

struct HostModelD {
  int T = 0, K = 0, L = 0;
  int stage = 0, cart = -1;        // Validate's loop limits: full stages [0, stage), then carts [0, cart] of `stage`
  std::vector<double> mean_shape;  // [2L]
  std::vector<NodeRecD> nodes;     // [T*K*7]
  std::vector<double> leaf;        // [T*K*8]
  std::vector<double> cart3;       // [T*K*3] = th, mean, std
  std::vector<double> w;           // [T][K*8][2L]
  bool any_scaled = false;
  bool loaded = false;
  int D() const { return 2 * L; }
};

// Either file flavour; a float file (c/jda.c:644-716) is widened exactly.
inline bool load_model_f64(const char *path, bool dbl, HostModelD &m, std::string &err) {
  FILE *f = fopen(path, "rb");
  if (!f) { err = std::string("cannot open ") + path; return false; }
  bool ok = true;
  auto i32 = [&]() { int v = 0; if (ok && fread(&v, 4, 1, f) != 1) ok = false; return v; };
  auto real = [&]() -> double {
    if (!ok) return 0.;
    if (dbl) { double d; if (fread(&d, 8, 1, f) != 1) { ok = false; return 0.; } return d; }
    float v; if (fread(&v, 4, 1, f) != 1) { ok = false; return 0.; } return (double)v;
  };
  int hdr[7];
  for (int i = 0; i < 7; i++) hdr[i] = i32();
  m.T = hdr[1]; m.K = hdr[2]; m.L = hdr[3];
  if (!ok || m.T <= 0 || m.T > 32 || m.K <= 0 || m.K > 4096 || m.L <= 0 || 2 * m.L > kMaxDim || hdr[4] != kDepth) {
    err = "bad model header"; fclose(f); return false;
  }
  // cascador.cpp:136-141; the float writer stores stage T+1 (c/jda.c:662): a finished model either way
  m.stage = hdr[5]; m.cart = hdr[6];
  if (m.stage > m.T) { m.stage = m.T; m.cart = -1; }
  if (m.stage < 0 || m.cart < -1 || m.cart >= m.K) { err = "bad stage / cart index in header"; fclose(f); return false; }
  if (m.stage == m.T) m.cart = -1;
  const int D = m.D();
  const size_t C = (size_t)m.T * m.K;
  m.mean_shape.resize(D);
  m.nodes.resize(C * kNodes);
  m.leaf.resize(C * kLeaves);
  m.cart3.resize(C * 3);
  m.w.resize((size_t)m.T * m.K * kLeaves * D);
  for (int i = 0; i < D; i++) m.mean_shape[i] = real();
  for (int t = 0; t < m.T && ok; t++) {
    for (int k = 0; k < m.K && ok; k++) {
      const size_t c = (size_t)t * m.K + k;
      for (int i = 0; i < kNodes; i++) {
        NodeRecD &n = m.nodes[c * kNodes + i];
        n.scale = i32();
        n.lm1 = i32() << 1;
        n.lm2 = i32() << 1;
        n.o1x = real(); n.o1y = real(); n.o2x = real(); n.o2y = real();
        n.th = i32();
      }
      for (int j = 0; j < kLeaves; j++) m.leaf[c * kLeaves + j] = real();
      for (int j = 0; j < 3; j++) m.cart3[c * 3 + j] = real();
    }
    double *wt = m.w.data() + (size_t)t * m.K * kLeaves * D;
    for (size_t i = 0; i < (size_t)m.K * kLeaves * D; i++) wt[i] = real();
  }
  i32();
  fclose(f);
  if (!ok) { err = "short read"; return false; }
  for (const NodeRecD &nd : m.nodes) {
    if (nd.lm1 < 0 || nd.lm1 + 1 >= D || nd.lm2 < 0 || nd.lm2 + 1 >= D) { err = "node field out of range"; return false; }
    if (nd.scale != 0) m.any_scaled = true;
  }
  m.loaded = true;
  return true;
}

// window sizes of detectMultiScale1 (cascador.cpp:335,372-373): win = int(win * factor) while it fits
inline int enumerate_levels_f64(int w, int h, int minimum_size, double factor, int *wins, int cap) {
  if (minimum_size <= 0 || !(factor > 1.)) return 0;  // the reference never terminates here
  int n = 0;
  for (int win = minimum_size; win <= w && win <= h;) {
    if (n < cap) wins[n] = win;
    n++;
    const int nw = (int)(win * factor);
    if (nw <= win) break;
    win = nw;
  }
  return n;
}

// data.cpp:38-51 at the mean shape: the pixel coordinates of a scale-0 node depend only on the window size
inline void node_coords_f64(const HostModelD &m, const NodeRecD &n, int win, int xy[4]) {
  const double v[4] = {(m.mean_shape[n.lm1] + n.o1x) * win, (m.mean_shape[n.lm1 + 1] + n.o1y) * win,
                       (m.mean_shape[n.lm2] + n.o2x) * win, (m.mean_shape[n.lm2 + 1] + n.o2y) * win};
  for (int i = 0; i < 4; i++) {
    int c = (int)std::round(v[i]);
    if (c < 0) c = 0;
    if (c >= win) c = win - 1;
    xy[i] = c;
  }
}

// Stage-0 prefilter thresholds.  k2_scan accumulates float32 scores; the reference accumulates doubles.  For every
// cart k this returns delta[k] >= |s32_k - s64_k| for ANY window (any sequence of leaves), so that a window whose
// float score is below th_k - delta[k] is certainly below th_k in double: the scan may drop it.  Survivors of the
// scan are then evaluated exactly, in double, from cart 0 -- the scan is a conservative filter, never the answer.
// Bound: A = running max |score| (sum of the largest |leaf| so far, pushed through the normalisations),
// e = running error; one float add costs ulp/2 <= (A + e) * 2^-24 plus the leaf's own narrowing maxleaf * 2^-24;
// a normalisation (s - mean) / std costs three more roundings and the narrowing of mean and std.  Everything is
// over-estimated by a factor of 4 for safety; delta stays around 1e-4 for the shipped model.
inline bool stage0_filter_margins(const HostModelD &m, std::vector<double> &delta) {
  const double u = std::ldexp(1.0, -24);
  delta.assign(m.K, 0.);
  double A = 0., e = 0.;
  for (int k = 0; k < m.K; k++) {
    double mx = 0.;
    for (int j = 0; j < kLeaves; j++) mx = std::max(mx, std::fabs(m.leaf[(size_t)k * kLeaves + j]));
    A += mx;
    e += mx * u + (A + e) * u;
    const double mean = m.cart3[(size_t)k * 3 + 1], sd = m.cart3[(size_t)k * 3 + 2];
    const double asd = std::fabs(sd);
    if (!(asd > 1e-12) || !std::isfinite(mean) || !std::isfinite(sd)) return false;  // no usable bound: no filter
    if ((float)mean != 0.f || (float)sd != 1.f) {
      // (s - mean): inputs off by e and |mean| u, result rounded; / std: std off by relative u, result rounded
      const double num = A + std::fabs(mean) + e;
      e = (e + std::fabs(mean) * u + num * u) / asd * (1. + 4. * u) + num / asd * 3. * u;
      A = (A + std::fabs(mean)) / asd;
    } else if (mean != 0. || sd != 1.) {
      // float sees (0, 1) and skips the step; double applies a near-identity normalisation
      e = e + A * std::fabs(1. - 1. / asd) + std::fabs(mean) / asd;
      A = (A + std::fabs(mean)) / asd;
    }
    delta[k] = 4. * e + 1e-9;
    if (!std::isfinite(delta[k]) || !std::isfinite(A)) return false;
  }
  return true;
}

// Stage-0 table of the scan kernel (same 104-byte cart records as build_stage0_table) for the double detector:
// offsets from double arithmetic with round(), float-narrowed leaf scores, thresholds lowered by the margin
// (rounded down), mean / std narrowed to float.
inline void build_stage0_table_f64(const HostModelD &m, const std::vector<double> &delta, int win, int tile_pitch,
                                   uint8_t *out, Stage0Norm *norm) {
  int nn = 0;
  for (int k = 0; k < m.K; k++) {
    uint8_t *rec = out + (size_t)k * kCartBytes;
    uint32_t *nd = reinterpret_cast<uint32_t *>(rec);
    for (int i = 0; i < kNodes; i++) {
      const NodeRecD &n = m.nodes[(size_t)k * kNodes + i];
      int xy[4];
      node_coords_f64(m, n, win, xy);
      const int th = std::max(-256, std::min(255, n.th));
      if (tile_pitch > 0) {
        nd[2 * i] = (uint32_t)(xy[1] * tile_pitch + xy[0]) | ((uint32_t)(xy[3] * tile_pitch + xy[2]) << 16);
        nd[2 * i + 1] = (uint32_t)th;
      } else {
        nd[2 * i] = (uint32_t)xy[0] | ((uint32_t)xy[1] << 11) | ((uint32_t)(th + 256) << 22);
        nd[2 * i + 1] = (uint32_t)xy[2] | ((uint32_t)xy[3] << 11);
      }
    }
    float lf[kLeaves];
    for (int j = 0; j < kLeaves; j++) lf[j] = (float)m.leaf[(size_t)k * kLeaves + j];
    memcpy(rec + 56, lf, 32);
    const double thd = m.cart3[(size_t)k * 3] - delta[k];
    float th32 = (float)thd;
    if ((double)th32 > thd) th32 = std::nextafterf(th32, -INFINITY);  // never above the lowered threshold
    memcpy(rec + 88, &th32, 4);
    uint32_t flag = 0;
    const float mean = (float)m.cart3[(size_t)k * 3 + 1], sd = (float)m.cart3[(size_t)k * 3 + 2];
    if (mean != 0.f || sd != 1.f) {
      if (nn < kMaxNorm) { norm[nn].mean = mean; norm[nn].std = sd; }
      flag = (uint32_t)(++nn);
    }
    memcpy(rec + 92, &flag, 4);
  }
}

inline int count_normed_stage0(const HostModelD &m) {
  int n = 0;
  for (int k = 0; k < m.K; k++)
    if ((float)m.cart3[(size_t)k * 3 + 1] != 0.f || (float)m.cart3[(size_t)k * 3 + 2] != 1.f) n++;
  return n;
}

// cascador.cpp:387-429.  A std::multimap keeps equal keys in insertion order, so walking it is walking the indices
// sorted by (score, index); rbegin() is the last of them.  rects = x, y, w, h.  Returns the picked indices in pick order.
inline std::vector<int> nms_f64(int n, const int *rects, const double *scores, double overlap) {
  std::vector<int> order(n), picked;
  for (int i = 0; i < n; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return scores[a] < scores[b]; });
  std::vector<uint8_t> in(n, 1);
  int left = n;
  while (left > 0) {
    int lp = n - 1;
    while (!in[lp]) lp--;
    const int last = order[lp];
    picked.push_back(last);
    const double area_last = (double)(rects[4 * last + 2] * rects[4 * last + 3]);
    for (int p = 0; p < n; p++) {
      if (!in[p]) continue;
      const int idx = order[p];
      const double x1 = std::max(rects[4 * idx], rects[4 * last]);
      const double y1 = std::max(rects[4 * idx + 1], rects[4 * last + 1]);
      const double x2 = std::min(rects[4 * idx] + rects[4 * idx + 2], rects[4 * last] + rects[4 * last + 2]);
      const double y2 = std::min(rects[4 * idx + 1] + rects[4 * idx + 3], rects[4 * last + 1] + rects[4 * last + 3]);
      const double w = std::max(0., x2 - x1);
      const double h = std::max(0., y2 - y1);
      const double area_idx = (double)(rects[4 * idx + 2] * rects[4 * idx + 3]);
      const double ov = w * h / (area_idx + area_last - w * h);
      if (ov > overlap) { in[p] = 0; left--; }
    }
  }
  return picked;
}

}  // namespace jda
