/*
 * jda_oracle_cpp.c -- CPU restatement of the reference's double-precision C++ detector,
 * JoinCascador::Detect with fddb.method = 1 (detectMultiScale1).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under jda_b200/ may include, link or execute this file.
 *
 * PINNED (round 2).  The reference's own src/jda/cascador.cpp and cart.cpp (whole files) plus the
 * detect-path functions of data.cpp / btcart.cpp / common.cpp are compiled from where they lie
 * against oracle/cvshim/ (a small stand-in for the OpenCV core headers: the image has no OpenCV C++)
 * into oracle/_ref_cpp/libjda_ref_cpp.so (oracle/Makefile: ref_cpp), and tests/test_oracle_cpp.py
 * compares this file with that binary bit for bit: JoinCascador::Detect (faces, scores, landmarks,
 * patch / cart statistics, with and without NMS), Validate on every window (carts evaluated, exit
 * score), the shipped model, full-precision synthetic models and training snapshots.
 * Scope of the pin = scope of this file: fddb.method = 1, every node at scale == 0; face.similarity_transform on or
 * off and any initial shift (jcpp_set_options; the reference build's RandomShape seed can be fixed for the
 * comparison) -- with cv::norm inside STParameter::Calc being the stand-in's index-order sum of squares, OpenCV's own
 * accumulation order is not pinned.  Models with scale != 0 nodes sample cv::resize'd planes
 * (cascador.cpp:330-331) -- OpenCV's arithmetic, third party, not under /root/reference: this file
 * refuses them, and the stand-in's resize is not OpenCV's.
 *
 * What it restates (reference file:line, /root/reference):
 *   model layout (double flavour)          src/jda/cascador.cpp:126-164, src/jda/cart.cpp:406-428
 *   window ladder / scan loops             src/jda/cascador.cpp:310-376   (detectMultiScale1)
 *   per-window cascade                     src/jda/cascador.cpp:166-211   (JoinCascador::Validate)
 *   initial shape                          src/jda/data.cpp:225-236       (RandomShape: mean + (x, y); (0, 0) is what
 *                                                                          test.cpp:17,75 force)
 *   tree walk                              src/jda/cart.cpp:392-404       (Cart::Forward, 1-based heap)
 *   pixel-difference feature               src/jda/data.cpp:18-58         (round(), per-view width, clamp:
 *                                                                          include/jda/common.hpp:227-232)
 *   similarity transform                   src/jda/data.cpp:64-126, include/jda/data.hpp:42-45   (STParameter::Calc / Apply;
 *                                                                          config.json ships face.similarity_transform = false:
 *                                                                          the identity, whose Apply returns its input exactly)
 *   global regression                      src/jda/btcart.cpp:407-424     (delta accumulated from 0, then shape += delta)
 *   nms                                    src/jda/cascador.cpp:387-429   (multimap by score, erase IoU > overlap)
 *   top level + relocation                 src/jda/cascador.cpp:431-477
 *
 * Build: gcc -std=c99 -O2 -ffp-contract=off -fPIC -shared (see Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int hdr[7]; /* mask, T, K, L, depth, current_stage_idx, current_cart_idx */
  int T, K, L, depth, nn, nl;
  int stage, cart; /* Validate's loop limits (cascador.cpp:178,199) */
  double *mean_shape;
  int *nd_scale, *nd_lm1, *nd_lm2, *nd_th; /* node i of cart c at c*nn + i, i = 0 is the root (heap index 1) */
  double *nd_off;                          /* 4 per node: o1x o1y o2x o2y */
  double *leaf, *cth, *cmean, *cstd;
  double *w; /* [t][K*nl][2L] */
  int any_scaled;
} ModelD;

static int rd_i32(FILE *f, int *v) { return fread(v, 4, 1, f) == 1; }
static int rd_real(FILE *f, int dbl, double *v) {
  if (dbl) return fread(v, 8, 1, f) == 1;
  float x;
  if (fread(&x, 4, 1, f) != 1) return 0;
  *v = (double)x; /* a float-flavour file (c/jda.c:644-716) widened exactly */
  return 1;
}

void jcpp_free(void *m_) {
  ModelD *m = (ModelD *)m_;
  if (!m) return;
  free(m->mean_shape); free(m->nd_scale); free(m->nd_lm1); free(m->nd_lm2); free(m->nd_th); free(m->nd_off);
  free(m->leaf); free(m->cth); free(m->cmean); free(m->cstd); free(m->w);
  free(m);
}

void *jcpp_load(const char *path, int dbl) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  ModelD *m = (ModelD *)calloc(1, sizeof(ModelD));
  int ok = 1;
  for (int i = 0; i < 7 && ok; i++) ok = rd_i32(f, &m->hdr[i]);
  m->T = m->hdr[1]; m->K = m->hdr[2]; m->L = m->hdr[3]; m->depth = m->hdr[4];
  if (!ok || m->T <= 0 || m->T > 64 || m->K <= 0 || m->K > 65536 || m->L <= 0 || m->L > 1024 || m->depth < 2 ||
      m->depth > 8) { fclose(f); free(m); return NULL; }
  /* cascador.cpp:136-141: 0 <= stage <= T, -1 <= cart < K.  The float writer stores stage T+1 (c/jda.c:662):
   * a finished model either way. */
  m->stage = m->hdr[5]; m->cart = m->hdr[6];
  if (m->stage > m->T) { m->stage = m->T; m->cart = -1; }
  if (m->stage < 0 || m->cart < -1 || m->cart >= m->K) { fclose(f); free(m); return NULL; }
  if (m->stage == m->T) m->cart = -1;
  m->nl = 1 << (m->depth - 1); m->nn = m->nl - 1;
  size_t C = (size_t)m->T * m->K, N = C * m->nn, D = 2 * (size_t)m->L;
  m->mean_shape = (double *)malloc(D * 8);
  m->nd_scale = (int *)malloc(N * 4); m->nd_lm1 = (int *)malloc(N * 4); m->nd_lm2 = (int *)malloc(N * 4);
  m->nd_th = (int *)malloc(N * 4); m->nd_off = (double *)malloc(N * 4 * 8);
  m->leaf = (double *)malloc(C * m->nl * 8);
  m->cth = (double *)malloc(C * 8); m->cmean = (double *)malloc(C * 8); m->cstd = (double *)malloc(C * 8);
  m->w = (double *)malloc((size_t)m->T * m->K * m->nl * D * 8);
  for (size_t i = 0; i < D && ok; i++) ok = rd_real(f, dbl, &m->mean_shape[i]);
  for (int t = 0; t < m->T && ok; t++) {
    for (int k = 0; k < m->K && ok; k++) {
      size_t c = (size_t)t * m->K + k;
      for (int i = 0; i < m->nn && ok; i++) {
        size_t n = c * m->nn + i;
        ok = ok && rd_i32(f, &m->nd_scale[n]) && rd_i32(f, &m->nd_lm1[n]) && rd_i32(f, &m->nd_lm2[n]);
        for (int j = 0; j < 4 && ok; j++) ok = rd_real(f, dbl, &m->nd_off[n * 4 + j]);
        ok = ok && rd_i32(f, &m->nd_th[n]);
        if (ok && (m->nd_lm1[n] < 0 || m->nd_lm1[n] >= m->L || m->nd_lm2[n] < 0 || m->nd_lm2[n] >= m->L)) ok = 0;
        if (ok && m->nd_scale[n] != 0) m->any_scaled = 1;
      }
      for (int j = 0; j < m->nl && ok; j++) ok = rd_real(f, dbl, &m->leaf[c * m->nl + j]);
      ok = ok && rd_real(f, dbl, &m->cth[c]) && rd_real(f, dbl, &m->cmean[c]) && rd_real(f, dbl, &m->cstd[c]);
    }
    size_t rows = (size_t)m->K * m->nl;
    double *wt = m->w + (size_t)t * rows * D;
    for (size_t i = 0; i < rows * D && ok; i++) ok = rd_real(f, dbl, &wt[i]);
  }
  fclose(f);
  if (!ok) { jcpp_free(m); return NULL; }
  return m;
}

void jcpp_dims(void *m_, int *out) {
  ModelD *m = (ModelD *)m_;
  out[0] = m->T; out[1] = m->K; out[2] = m->L; out[3] = m->depth; out[4] = m->stage; out[5] = m->cart; out[6] = m->any_scaled;
}

/* ---------------------------------------------------------------- one window */

/* face.similarity_transform (config.json) and the initial shift of DataSet::RandomShape (data.cpp:225-236).  The
 * reference draws the shift from an RNG seeded with the tick count, per window: a caller that wants a non-zero shift
 * passes the (x, y) that RNG produced.  Process-wide, like the reference's Config singleton. */
static struct { int similarity; double shift_x, shift_y; } g_opt = {0, 0., 0.};
void jcpp_set_options(int similarity, double shift_x, double shift_y) {
  g_opt.similarity = similarity; g_opt.shift_x = shift_x; g_opt.shift_y = shift_y;
}

/* STParameter (include/jda/data.hpp:18-50): default = identity */
typedef struct { double scale, r00, r01, r10, r11; } STP;
static const STP STP_IDENTITY = {1., 1., 0., 0., 1.};

/* data.hpp:42-45 */
static void stp_apply(const STP *p, double x1, double y1, double *x2, double *y2) {
  *x2 = p->scale * (p->r00 * x1 + p->r01 * y1);
  *y2 = p->scale * (p->r10 * x1 + p->r11 * y1);
}

/* data.cpp:64-114, shape1 -> shape2.  cv::norm is OpenCV's; the reference binary this file is pinned against is built
 * with oracle/cvshim's stand-in: the square root of the squares summed in index order.  tmp: 2 * D doubles. */
static STP stp_calc(const double *shape1, const double *shape2, int L, double *tmp) {
  STP p = STP_IDENTITY;
  if (!g_opt.similarity) return p;
  double x1c = 0., y1c = 0., x2c = 0., y2c = 0.;
  for (int i = 0; i < L; i++) {
    x1c += shape1[2 * i]; y1c += shape1[2 * i + 1];
    x2c += shape2[2 * i]; y2c += shape2[2 * i + 1];
  }
  x1c /= L; y1c /= L; x2c /= L; y2c /= L;
  double *t1 = tmp, *t2 = tmp + 2 * L;
  for (int i = 0; i < L; i++) {
    t1[2 * i] = shape1[2 * i] - x1c; t1[2 * i + 1] = shape1[2 * i + 1] - y1c;
    t2[2 * i] = shape2[2 * i] - x2c; t2[2 * i + 1] = shape2[2 * i + 1] - y2c;
  }
  double s1 = 0., s2 = 0.;
  for (int j = 0; j < 2 * L; j++) s1 += t1[j] * t1[j];
  for (int j = 0; j < 2 * L; j++) s2 += t2[j] * t2[j];
  const double scale1 = sqrt(s1), scale2 = sqrt(s2);
  p.scale = scale1 / scale2;
  for (int j = 0; j < 2 * L; j++) t1[j] /= scale1;
  for (int j = 0; j < 2 * L; j++) t2[j] /= scale2;
  double num = 0., den = 0.;
  for (int i = 0; i < L; i++) {
    num += t1[2 * i + 1] * t2[2 * i] - t1[2 * i] * t2[2 * i + 1];
    den += t1[2 * i] * t2[2 * i] + t1[2 * i + 1] * t2[2 * i + 1];
  }
  const double norm = sqrt(num * num + den * den);
  const double sin_theta = num / norm, cos_theta = den / norm;
  p.r00 = cos_theta; p.r01 = -sin_theta; p.r10 = sin_theta; p.r11 = cos_theta;
  return p;
}

/* data.cpp:18-58 for a scale == 0 node: the view is the win x win patch of the frame at (x, y) */
static int feature_value(const ModelD *m, size_t n, const double *s, const STP *stp, const unsigned char *img, int stride,
                         int x, int y, int win) {
  const double *o = m->nd_off + n * 4;
  /* stp_mc.Apply(offset): with the identity transform scale*(1*ox + 0*oy) == ox for every finite value */
  double o1x, o1y, o2x, o2y;
  stp_apply(stp, o[0], o[1], &o1x, &o1y);
  stp_apply(stp, o[2], o[3], &o2x, &o2y);
  const double x1 = (s[2 * m->nd_lm1[n]] + o1x) * win;
  const double y1 = (s[2 * m->nd_lm1[n] + 1] + o1y) * win;
  const double x2 = (s[2 * m->nd_lm2[n]] + o2x) * win;
  const double y2 = (s[2 * m->nd_lm2[n] + 1] + o2y) * win;
  int x1_ = (int)round(x1), y1_ = (int)round(y1), x2_ = (int)round(x2), y2_ = (int)round(y2);
  if (x1_ < 0) x1_ = 0; if (y1_ < 0) y1_ = 0; if (x1_ >= win) x1_ = win - 1; if (y1_ >= win) y1_ = win - 1;
  if (x2_ < 0) x2_ = 0; if (y2_ < 0) y2_ = 0; if (x2_ >= win) x2_ = win - 1; if (y2_ >= win) y2_ = win - 1;
  return (int)img[(size_t)(y + y1_) * stride + x + x1_] - (int)img[(size_t)(y + y2_) * stride + x + x2_];
}

/* cart.cpp:392-404 (1-based heap there; node j of the heap is stored at j-1 here) */
static int forward(const ModelD *m, size_t c, const double *s, const STP *stp, const unsigned char *img, int stride, int x,
                   int y, int win) {
  int node_idx = 1;
  int len = m->depth - 1;
  while (len--) {
    const size_t n = c * m->nn + (node_idx - 1);
    const int val = feature_value(m, n, s, stp, img, stride, x, y, win);
    if (val <= m->nd_th[n]) node_idx = 2 * node_idx;
    else node_idx = 2 * node_idx + 1;
  }
  return node_idx - m->nl;
}

/* cascador.cpp:166-211.  shape: [2L] out; lbf, delta: scratch.  Returns is_face; *n_out = carts evaluated. */
static int validate(const ModelD *m, const unsigned char *img, int stride, int x, int y, int win, double *score_out,
                    double *shape, int *n_out, int *lbf, double *delta) {
  const int D = 2 * m->L;
  /* RandomShape: mean + (x, y); test.cpp:17,75 force the shift to 0 */
  for (int j = 0; j < D; j++) shape[j] = m->mean_shape[j] + ((j & 1) ? g_opt.shift_y : g_opt.shift_x);
  double score = 0;
  int n = 0;
  STP stp = STP_IDENTITY; /* STParameter stp_mc; (cascador.cpp:176) */
  double *stp_tmp = g_opt.similarity ? (double *)malloc((size_t)2 * D * sizeof(double)) : NULL;
  for (int t = 0; t < m->stage; t++) {
    stp = stp_calc(shape, m->mean_shape, m->L, stp_tmp); /* cascador.cpp:180 */
    int offset = 0;
    for (int k = 0; k < m->K; k++) {
      const size_t c = (size_t)t * m->K + k;
      const int idx = forward(m, c, shape, &stp, img, stride, x, y, win);
      score += m->leaf[c * m->nl + idx];
      score = (score - m->cmean[c]) / m->cstd[c];
      n++;
      if (score < m->cth[c]) { *score_out = score; *n_out = n; free(stp_tmp); return 0; }
      lbf[k] = offset + idx;
      offset += m->nl;
    }
    /* btcart.cpp:407-424 */
    const double *wt = m->w + (size_t)t * m->K * m->nl * D;
    for (int j = 0; j < D; j++) delta[j] = 0.;
    for (int i = 0; i < m->K; i++) {
      const double *w_ptr = wt + (size_t)lbf[i] * D;
      for (int j = 0; j < D; j++) delta[j] += w_ptr[j];
    }
    /* stp_mc.Apply(delta_shape, delta_shape), btcart.cpp:422 */
    for (int j = 0; j < m->L; j++) stp_apply(&stp, delta[2 * j], delta[2 * j + 1], &delta[2 * j], &delta[2 * j + 1]);
    for (int j = 0; j < D; j++) shape[j] += delta[j];
  }
  /* unfinished stage of a training snapshot (empty for a finished model): with the transform of the last finished
   * stage -- the reference does not recompute it here (cascador.cpp:199-202) */
  for (int k = 0; k <= m->cart; k++) {
    const size_t c = (size_t)m->stage * m->K + k;
    const int idx = forward(m, c, shape, &stp, img, stride, x, y, win);
    score += m->leaf[c * m->nl + idx];
    score = (score - m->cmean[c]) / m->cstd[c];
    n++;
    if (score < m->cth[c]) { *score_out = score; *n_out = n; free(stp_tmp); return 0; }
  }
  free(stp_tmp);
  *score_out = score; *n_out = n;
  return 1;
}

/* ---------------------------------------------------------------- scan (cascador.cpp:310-376) */

int jcpp_levels(int W, int H, int minimum_size, double factor, int *wins, int cap) {
  int n = 0;
  if (minimum_size <= 0 || !(factor > 1.)) return 0; /* the reference never terminates here */
  for (int win = minimum_size; win <= W && win <= H;) {
    if (n < cap) wins[n] = win;
    n++;
    const int nw = (int)(win * factor);
    if (nw <= win) break;
    win = nw;
  }
  return n;
}

long long jcpp_count_windows(int W, int H, int minimum_size, int step, double factor) {
  int wins[256];
  int n = jcpp_levels(W, H, minimum_size, factor, wins, 256);
  if (n > 256) n = 256;
  if (step <= 0) return 0;
  long long tot = 0;
  for (int i = 0; i < n; i++) tot += (long long)((W - wins[i]) / step + 1) * ((H - wins[i]) / step + 1);
  return tot;
}

typedef struct {
  int n, cap;
  int *rects;     /* x y w h */
  double *scores;
  double *shapes; /* window-normalised */
} Hits;

static void hits_push(Hits *h, int D, int x, int y, int win, double score, const double *shape) {
  if (h->n == h->cap) {
    h->cap = h->cap ? 2 * h->cap : 64;
    h->rects = (int *)realloc(h->rects, (size_t)h->cap * 4 * sizeof(int));
    h->scores = (double *)realloc(h->scores, (size_t)h->cap * sizeof(double));
    h->shapes = (double *)realloc(h->shapes, (size_t)h->cap * D * sizeof(double));
  }
  int *r = h->rects + 4 * (size_t)h->n;
  r[0] = x; r[1] = y; r[2] = win; r[3] = win;
  h->scores[h->n] = score;
  memcpy(h->shapes + (size_t)h->n * D, shape, D * sizeof(double));
  h->n++;
}

/* Scan of detectMultiScale1.  trace_n / trace_s (may be NULL): per window in scan order, carts evaluated and
 * exit score.  Returns hits (caller frees the three arrays) or n = -1 for a model this file refuses. */
static Hits scan(const ModelD *m, const unsigned char *img, int W, int H, int minimum_size, int step, double factor,
                 int *trace_n, double *trace_s, long long *carts_total) {
  Hits h;
  memset(&h, 0, sizeof h);
  if (m->any_scaled) { h.n = -1; return h; }
  const int D = 2 * m->L;
  double *shape = (double *)malloc(D * sizeof(double)), *delta = (double *)malloc(D * sizeof(double));
  int *lbf = (int *)malloc(m->K * sizeof(int));
  int wins[256];
  int nl = jcpp_levels(W, H, minimum_size, factor, wins, 256);
  if (nl > 256) nl = 256;
  if (step <= 0) nl = 0;
  long long wi = 0, carts = 0;
  for (int li = 0; li < nl; li++) {
    const int win = wins[li];
    for (int y = 0; y <= H - win; y += step) {
      for (int x = 0; x <= W - win; x += step) {
        double score;
        int n;
        const int is_face = validate(m, img, W, x, y, win, &score, shape, &n, lbf, delta);
        carts += n;
        if (trace_n) trace_n[wi] = n;
        if (trace_s) trace_s[wi] = score;
        wi++;
        if (is_face) hits_push(&h, D, x, y, win, score, shape);
      }
    }
  }
  if (carts_total) *carts_total = carts;
  free(shape); free(delta); free(lbf);
  return h;
}

/* cascador.cpp:387-429.  A std::multimap keeps equal keys in insertion order, so iterating it is iterating the
 * indices sorted by (score, index); rbegin() is the last of them.  picked[] receives the result, returns its size. */
int jcpp_nms(int n, const int *rects, const double *scores, double overlap, int *picked) {
  if (!(overlap < 1.)) return -1; /* a box would not suppress itself: endless loop in the reference */
  int *order = (int *)malloc((n > 0 ? n : 1) * sizeof(int));
  unsigned char *in = (unsigned char *)malloc(n > 0 ? n : 1);
  for (int i = 0; i < n; i++) { order[i] = i; in[i] = 1; }
  /* stable insertion sort by score ascending == multimap order */
  for (int i = 1; i < n; i++) {
    const int v = order[i];
    int j = i - 1;
    while (j >= 0 && scores[order[j]] > scores[v]) { order[j + 1] = order[j]; j--; }
    order[j + 1] = v;
  }
  int picked_n = 0, left = n;
  while (left > 0) {
    int lp = n - 1;
    while (!in[lp]) lp--;
    const int last = order[lp];
    picked[picked_n++] = last;
    const double area_last = (double)(rects[4 * last + 2] * rects[4 * last + 3]);
    for (int p = 0; p < n; p++) {
      if (!in[p]) continue;
      const int idx = order[p];
      const double x1 = rects[4 * idx] > rects[4 * last] ? rects[4 * idx] : rects[4 * last];
      const double y1 = rects[4 * idx + 1] > rects[4 * last + 1] ? rects[4 * idx + 1] : rects[4 * last + 1];
      const int ax2 = rects[4 * idx] + rects[4 * idx + 2], bx2 = rects[4 * last] + rects[4 * last + 2];
      const int ay2 = rects[4 * idx + 1] + rects[4 * idx + 3], by2 = rects[4 * last + 1] + rects[4 * last + 3];
      const double x2 = ax2 < bx2 ? ax2 : bx2;
      const double y2 = ay2 < by2 ? ay2 : by2;
      const double w = 0. > x2 - x1 ? 0. : x2 - x1;
      const double hh = 0. > y2 - y1 ? 0. : y2 - y1;
      const double area_idx = (double)(rects[4 * idx + 2] * rects[4 * idx + 3]);
      const double ov = w * hh / (area_idx + area_last - w * hh);
      if (ov > overlap) { in[p] = 0; left--; }
    }
  }
  free(order); free(in);
  return picked_n;
}

/* cascador.cpp:431-477.  Outputs are malloc'd; free with jcpp_release.  Returns n, or -1 (refused model / bad arguments). */
int jcpp_detect(void *m_, const unsigned char *img, int W, int H, int minimum_size, int step, double factor,
                double overlap, int use_nms, int **rects_out, double **scores_out, double **shapes_out,
                long long *carts_total) {
  const ModelD *m = (const ModelD *)m_;
  const int D = 2 * m->L;
  Hits h = scan(m, img, W, H, minimum_size, step, factor, NULL, NULL, carts_total);
  if (h.n < 0) return -1;
  int *picked = (int *)malloc((h.n > 0 ? h.n : 1) * sizeof(int));
  int n;
  if (use_nms) {
    n = jcpp_nms(h.n, h.rects, h.scores, overlap, picked);
    if (n < 0) { free(picked); free(h.rects); free(h.scores); free(h.shapes); return -1; }
  } else {
    n = h.n;
    for (int i = 0; i < n; i++) picked[i] = i;
  }
  int *rects = (int *)malloc((n > 0 ? n : 1) * 4 * sizeof(int));
  double *scores = (double *)malloc((n > 0 ? n : 1) * sizeof(double));
  double *shapes = (double *)malloc((size_t)(n > 0 ? n : 1) * D * sizeof(double));
  for (int i = 0; i < n; i++) {
    const int index = picked[i];
    const int *r = h.rects + 4 * (size_t)index;
    const double *s = h.shapes + (size_t)index * D;
    for (int j = 0; j < m->L; j++) {
      shapes[(size_t)i * D + 2 * j] = r[0] + s[2 * j] * r[2];
      shapes[(size_t)i * D + 2 * j + 1] = r[1] + s[2 * j + 1] * r[3];
    }
    memcpy(rects + 4 * (size_t)i, r, 4 * sizeof(int));
    scores[i] = h.scores[index];
  }
  free(picked); free(h.rects); free(h.scores); free(h.shapes);
  *rects_out = rects; *scores_out = scores; *shapes_out = shapes;
  return n;
}

void jcpp_release(int *rects, double *scores, double *shapes) { free(rects); free(scores); free(shapes); }

/* per-window trace in scan order: carts evaluated (Validate's n) and exit score; returns the window count or -1 */
long long jcpp_trace(void *m_, const unsigned char *img, int W, int H, int minimum_size, int step, double factor,
                     int *trace_n, double *trace_s) {
  const ModelD *m = (const ModelD *)m_;
  Hits h = scan(m, img, W, H, minimum_size, step, factor, trace_n, trace_s, NULL);
  if (h.n < 0) return -1;
  free(h.rects); free(h.scores); free(h.shapes);
  return jcpp_count_windows(W, H, minimum_size, step, factor);
}
