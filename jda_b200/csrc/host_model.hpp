// host_model.hpp -- host side of the JDA detect path: model files, scan geometry, stage-0
// look-up tables, NMS and relocation.  Pure C++ (no CUDA), compiled with -ffp-contract=off so
// that float expressions round exactly like the reference's (FMA contraction of
// shape*size+x changes landmark bits).
//
// Reference behaviour restated here (file:line under /root/reference):
//   model layouts / loaders      c/jda.c:486-561 (double), 563-638 (float), README.md:84-111
//   serialiser                   c/jda.c:644-716
//   level / window enumeration   c/jda.c:320-339
//   node address arithmetic      c/jda.c:371-389   (used to pre-compute the stage-0 tables)
//   nms                          c/jda.c:237-316
//   relocation                   c/jda.c:465-474
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace jda {

constexpr int kDepth = 4;       // the reference's depth (c/jda.c:28); the scan kernel and its tables are built for it
constexpr int kNodes = 7;       // internal nodes per depth-4 cart
constexpr int kLeaves = 8;      // leaves per depth-4 cart
constexpr int kMinDepth = 2, kMaxDepth = 6;  // depths the generic cascade kernel runs (header field tree_depth)
constexpr int kMaxLevels = 64;  // pyramid levels per geometry (the survivor key has 6 bits for the level)
constexpr int kMaxDim = 128;    // 2 * landmark_n upper bound
constexpr int kMaxNorm = 32;    // stage-0 carts with non-trivial (mean, std) the scan kernel can hold
constexpr int kCartBytes = 104;  // stage-0 table record of one cart (layout below, at build_stage0_table)
constexpr int kMaxStage0TableBytes = 88 * 1024;  // K <= 866: the table + 12 x 11 KB of warp scratch fit 227 KB

// 32-byte node record as the cascade kernel reads it (two 16-byte loads)
struct alignas(16) NodeRec {
  int scale;
  int lm1, lm2;  // landmark x index (already *2, c/jda.c:521-523)
  int th;
  float o1x, o1y, o2x, o2y;
};
static_assert(sizeof(NodeRec) == 32, "NodeRec must be 32 bytes");

struct HostModel {
  int hdr[7] = {0, 0, 0, 0, 0, 0, 0};
  int T = 0, K = 0, L = 0;
  int depth = kDepth;             // tree_depth from the header (README.md:84-111); nn / nl follow from it
  int nn = kNodes, nl = kLeaves;  // internal nodes 2^(depth-1) - 1 and leaves 2^(depth-1) per cart
  std::vector<float> mean_shape;  // [2L]
  std::vector<NodeRec> nodes;     // [T*K*nn]
  std::vector<float> leaf;        // [T*K*nl]
  std::vector<float> cart;        // [T*K*4] = th, mean, std, 0
  std::vector<float> w;           // [T][K*nl][2L]
  bool any_scaled = false;        // some node samples the h/q planes
  bool stage0_lut_ok = false;     // stage 0 can run from per-level integer look-up tables
  int D() const { return 2 * L; }
};

namespace detail {
struct Reader {
  FILE *f;
  bool dbl;
  bool ok = true;
  int i32() {
    int v = 0;
    if (ok && fread(&v, 4, 1, f) != 1) ok = false;
    return v;
  }
  float real() {
    if (!ok) return 0.f;
    if (dbl) {
      double d;
      if (fread(&d, 8, 1, f) != 1) { ok = false; return 0.f; }
      return (float)d;  // same narrowing as c/jda.c:509
    }
    float v;
    if (fread(&v, 4, 1, f) != 1) { ok = false; return 0.f; }
    return v;
  }
};
}  // namespace detail

// Loads either flavour.  Unlike the reference (which ignores the header, c/jda.c:499-505) the
// dimensions are taken from it and sanity-checked.
inline bool load_model(const char *path, bool dbl, HostModel &m, std::string &err) {
  FILE *f = fopen(path, "rb");
  if (!f) { err = std::string("cannot open ") + path; return false; }
  detail::Reader r{f, dbl};
  for (int i = 0; i < 7; i++) m.hdr[i] = r.i32();
  m.T = m.hdr[1]; m.K = m.hdr[2]; m.L = m.hdr[3];
  int depth = m.hdr[4];
  if (!r.ok || m.T <= 0 || m.T > 32 || m.K <= 0 || m.K > 4096 || m.L <= 0 || 2 * m.L > kMaxDim) {
    err = "bad model header"; fclose(f); return false;
  }
  // c/jda.c:24-32 fixes tree_depth = 4 at compile time; here it comes from the header like T / K / L.  Depth 4 runs
  // the stage-0 scan kernel; other depths run every stage through the generic cascade kernel.
  if (depth < kMinDepth || depth > kMaxDepth) { err = "tree_depth outside 2..6"; fclose(f); return false; }
  m.depth = depth; m.nl = 1 << (depth - 1); m.nn = m.nl - 1;
  const int kNodes = m.nn, kLeaves = m.nl;  // (shadow the depth-4 constants inside this function)
  const int D = m.D();
  const size_t C = (size_t)m.T * m.K;
  m.mean_shape.resize(D);
  m.nodes.resize(C * kNodes);
  m.leaf.resize(C * kLeaves);
  m.cart.assign(C * 4, 0.f);
  m.w.resize((size_t)m.T * m.K * kLeaves * D);
  for (int i = 0; i < D; i++) m.mean_shape[i] = r.real();
  for (int t = 0; t < m.T && r.ok; t++) {
    for (int k = 0; k < m.K && r.ok; k++) {
      size_t c = (size_t)t * m.K + k;
      for (int i = 0; i < kNodes; i++) {
        NodeRec &n = m.nodes[c * kNodes + i];
        n.scale = r.i32();
        n.lm1 = r.i32() << 1;
        n.lm2 = r.i32() << 1;
        n.o1x = r.real(); n.o1y = r.real(); n.o2x = r.real(); n.o2y = r.real();
        n.th = r.i32();
      }
      for (int j = 0; j < kLeaves; j++) m.leaf[c * kLeaves + j] = r.real();
      m.cart[c * 4 + 0] = r.real();
      m.cart[c * 4 + 1] = r.real();
      m.cart[c * 4 + 2] = r.real();
    }
    float *wt = m.w.data() + (size_t)t * m.K * kLeaves * D;
    for (size_t i = 0; i < (size_t)m.K * kLeaves * D; i++) wt[i] = r.real();
  }
  r.i32();  // trailing mask
  fclose(f);
  if (!r.ok) { err = "short read"; return false; }
  // validate indices so no kernel can index outside the shape / plane arrays
  int norm0 = 0;
  bool s0_scaled = false;
  for (size_t n = 0; n < m.nodes.size(); n++) {
    const NodeRec &nd = m.nodes[n];
    if (nd.scale < 0 || nd.scale > 2 || nd.lm1 < 0 || nd.lm1 + 1 >= D || nd.lm2 < 0 || nd.lm2 + 1 >= D) {
      err = "node field out of range"; return false;
    }
    if (nd.scale != 0) {
      m.any_scaled = true;
      if (n < (size_t)m.K * kNodes) s0_scaled = true;
    }
  }
  for (int k = 0; k < m.K; k++)
    if (m.cart[k * 4 + 1] != 0.f || m.cart[k * 4 + 2] != 1.f) norm0++;
  // the scan kernel keeps one level's table (K x kCartBytes) in shared memory next to its tile buffers
  m.stage0_lut_ok = depth == kDepth && !s0_scaled && norm0 <= kMaxNorm && m.K * kCartBytes <= kMaxStage0TableBytes;
  return true;
}

// Model file writer.  Default = the float32 flavour with the byte layout of c/jda.c:644-716, including its header
// quirk: stage field T + 1, cart -1 (c/jda.c:662-665), which the C++ loader rejects (cascador.cpp:138 wants T).
//   stage_T : write T in the stage field instead
//   dbl     : double flavour (README.md:84-111; f32 -> f64 widening is exact), what JoinCascador::SerializeFrom reads
inline bool save_model(const HostModel &m, const char *path, bool stage_T, bool dbl) {
  FILE *f = fopen(path, "wb");
  if (!f) return false;
  const int kNodes = m.nn, kLeaves = m.nl;
  int h[7] = {0, m.T, m.K, m.L, m.depth, stage_T ? m.T : m.T + 1, -1};
  fwrite(h, 4, 7, f);
  auto reals = [&](const float *p, size_t n) {
    if (!dbl) { fwrite(p, 4, n, f); return; }
    for (size_t i = 0; i < n; i++) { const double d = (double)p[i]; fwrite(&d, 8, 1, f); }
  };
  const int D = m.D();
  reals(m.mean_shape.data(), D);
  for (int t = 0; t < m.T; t++) {
    for (int k = 0; k < m.K; k++) {
      size_t c = (size_t)t * m.K + k;
      for (int i = 0; i < kNodes; i++) {
        const NodeRec &n = m.nodes[c * kNodes + i];
        int a = n.lm1 >> 1, b = n.lm2 >> 1;
        fwrite(&n.scale, 4, 1, f); fwrite(&a, 4, 1, f); fwrite(&b, 4, 1, f);
        reals(&n.o1x, 4);
        fwrite(&n.th, 4, 1, f);
      }
      reals(&m.leaf[c * kLeaves], kLeaves);
      reals(&m.cart[c * 4], 3);
    }
    reals(m.w.data() + (size_t)t * m.K * kLeaves * D, (size_t)m.K * kLeaves * D);
  }
  int z = 0;
  fwrite(&z, 4, 1, f);
  fclose(f);
  return true;
}
inline bool save_model_f32(const HostModel &m, const char *path) { return save_model(m, path, false, false); }

// ------------------------------------------------------------------------------- geometry

// window sizes visited by c/jda.c:320-332 after the clamps of c/jda.c:459-460
inline int enumerate_levels(int w, int h, float scale, int min_size, int max_size, int *wins, int cap) {
  if (min_size < 24) min_size = 24;
  if (max_size <= 0) max_size = std::min(w, h);
  max_size = std::min(max_size, std::min(w, h));
  if (!(scale > 1.f)) return 0;  // the reference never terminates here
  int win = 24, n = 0;
  while (win < min_size) {
    int nw = (int)(win * scale);
    if (nw <= win) return 0;
    win = nw;
  }
  while (win <= max_size) {
    if (n < cap) wins[n] = win;
    n++;
    int nw = (int)(win * scale);
    if (nw <= win) break;
    win = nw;
  }
  return n;
}

inline int level_step(int win) { return (int)(win * 0.1f); }  // c/jda.c:333

inline long long count_windows(int w, int h, float scale, int min_size, int max_size) {
  if (w < 24 || h < 24) return 0;
  int wins[256];
  int n = std::min(enumerate_levels(w, h, scale, min_size, max_size, wins, 256), 256);
  long long tot = 0;
  for (int i = 0; i < n; i++) {
    int s = level_step(wins[i]);
    tot += (long long)((h - wins[i]) / s + 1) * ((w - wins[i]) / s + 1);
  }
  return tot;
}

// ------------------------------------------------------------------- stage-0 look-up tables
//
// In stage 0 every window's shape is the constant mean shape, so the pixel coordinates of
// c/jda.c:373-389 depend only on the node and the window size: they are folded, bit-exactly
// (same float add, float mul, truncation, clamp), into integers per (level, node).
//
// Cart record, 104 bytes (26 words: consecutive carts land in different banks for the lane = cart
// reads of straggler mode):
//   [0..56)   7 nodes x {u32 a, i32 b}
//                tile format  : a = off1 | off2 << 16  (byte offsets inside the smem tile), b = th
//                packed format: a = x1 | y1 << 11 | (th + 256) << 22, b = x2 | y2 << 11  (win < 2048)
//   [56..88)  8 leaf scores (f32)
//   [88]      cart threshold (f32)
//   [92]      0, or 1 + index into the norm table when (mean, std) != (0, 1)


struct Stage0Norm { float mean, std; };

inline void node_coords(const HostModel &m, const NodeRec &n, int win, int xy[4]) {
  float x1 = m.mean_shape[n.lm1] + n.o1x;
  float y1 = m.mean_shape[n.lm1 + 1] + n.o1y;
  float x2 = m.mean_shape[n.lm2] + n.o2x;
  float y2 = m.mean_shape[n.lm2 + 1] + n.o2y;
  float v[4] = {x1, y1, x2, y2};
  for (int i = 0; i < 4; i++) {
    int c = (int)(v[i] * win);
    if (c < 0) c = 0; else if (c >= win) c = win - 1;
    xy[i] = c;
  }
}

// tile_pitch > 0: tile format with that row pitch; tile_pitch == 0: packed format
inline void build_stage0_table(const HostModel &m, int win, int tile_pitch, uint8_t *out /* K*96 */,
                               Stage0Norm *norm /* kMaxNorm */) {
  int nn = 0;
  for (int k = 0; k < m.K; k++) {
    uint8_t *rec = out + (size_t)k * kCartBytes;
    uint32_t *nd = reinterpret_cast<uint32_t *>(rec);
    for (int i = 0; i < kNodes; i++) {
      const NodeRec &n = m.nodes[(size_t)k * kNodes + i];
      int xy[4];
      node_coords(m, n, win, xy);
      // feature = p1 - p2 is in [-255, 255]; clamping th to [-256, 255] keeps `feature <= th`
      int th = std::max(-256, std::min(255, n.th));
      if (tile_pitch > 0) {
        uint32_t o1 = (uint32_t)(xy[1] * tile_pitch + xy[0]);
        uint32_t o2 = (uint32_t)(xy[3] * tile_pitch + xy[2]);
        nd[2 * i] = o1 | (o2 << 16);
        nd[2 * i + 1] = (uint32_t)th;
      } else {
        nd[2 * i] = (uint32_t)xy[0] | ((uint32_t)xy[1] << 11) | ((uint32_t)(th + 256) << 22);
        nd[2 * i + 1] = (uint32_t)xy[2] | ((uint32_t)xy[3] << 11);
      }
    }
    memcpy(rec + 56, &m.leaf[(size_t)k * kLeaves], 32);
    memcpy(rec + 88, &m.cart[(size_t)k * 4], 4);
    uint32_t flag = 0;
    float mean = m.cart[(size_t)k * 4 + 1], sd = m.cart[(size_t)k * 4 + 2];
    if (mean != 0.f || sd != 1.f) {
      if (nn < kMaxNorm) { norm[nn].mean = mean; norm[nn].std = sd; }
      flag = (uint32_t)(++nn);
    }
    memcpy(rec + 92, &flag, 4);
  }
}

// -------------------------------------------------------------------------- post-processing

// c/jda.c:237-316: exchange sort of indices by score (strict <), greedy suppression of IoU > 0.3,
// survivors keep their original order.
inline void nms(int n, const int *box, const float *score, uint8_t *keep) {
  const float overlap = 0.3f;
  std::vector<int> idx(n > 0 ? n : 0);
  for (int i = 0; i < n; i++) { idx[i] = i; keep[i] = 1; }
  for (int i = 0; i + 1 < n; i++)
    for (int j = i + 1; j < n; j++)
      if (score[idx[i]] < score[idx[j]]) std::swap(idx[i], idx[j]);
  for (int i = 0; i + 1 < n; i++) {
    const int a = idx[i];
    if (!keep[a]) continue;
    const int ax = box[3 * a], ay = box[3 * a + 1], as = box[3 * a + 2];
    for (int j = i + 1; j < n; j++) {
      const int b = idx[j];
      if (!keep[b]) continue;
      const int bx = box[3 * b], by = box[3 * b + 1], bs = box[3 * b + 2];
      const int x1 = std::max(ax, bx), y1 = std::max(ay, by);
      const int x2 = std::min(ax + as, bx + bs), y2 = std::min(ay + as, by + bs);
      const int iw = std::max(0, x2 - x1), ih = std::max(0, y2 - y1);
      const float ov = (float)(iw * ih) / (float)(as * as + bs * bs - iw * ih);
      if (ov > overlap) keep[b] = 0;
    }
  }
}

// c/jda.c:470-473 (two roundings: mul, then add; never fused)
inline void relocate(const float *src, float *dst, int L, int x, int y, int size) {
  for (int j = 0; j < L; j++) {
    dst[2 * j] = src[2 * j] * size + x;
    dst[2 * j + 1] = src[2 * j + 1] * size + y;
  }
}

}  // namespace jda
