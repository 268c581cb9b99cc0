/*
 * jda_oracle.c -- CPU restatement of the JDA float32 detect path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under jda_b200/ may include, link or
 * execute this file; it exists so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg can check the CUDA path.  The product path has
 * no CPU fallback.
 *
 * Parity pin: this restatement is checked, bit for bit, against the
 * reference's own c/jda.c compiled into oracle/_ref/libjda_ref.so (see
 * oracle/Makefile and tests/test_oracle_pin.py) on the shipped model, the
 * reference's face image and synthetic frames/models.  The reference has no
 * golden vectors of its own.
 *
 * What it restates (reference file:line, /root/reference):
 *   model load, double + float flavours   c/jda.c:486-561, 563-638
 *   float32 serialiser                    c/jda.c:644-716
 *   bilinear down-sample                  c/jda.c:203-230
 *   scan loops / level enumeration        c/jda.c:318-355
 *   node test, cart walk, early reject    c/jda.c:364-402
 *   global regression (leaf-row gather)   c/jda.c:403-411
 *   final threshold + emit                c/jda.c:413-427
 *   nms                                   c/jda.c:237-316
 *   top level + landmark relocation       c/jda.c:443-480
 *
 * Differences from the reference, all deliberate:
 *   - dimensions (T, K, L, depth) come from the file header instead of macros;
 *   - the model is held as flat arrays, not nested structs;
 *   - instrumented entry points expose pre-NMS hits, per-window trace and
 *     work counters (the reference can only be observed through jdaDetect);
 *   - a feature whose pyramid plane index would fall outside the plane buffer
 *     (undefined behaviour in the reference for scale 1/2 nodes, SURVEY.md
 *     section 8 a4) reads as 0; in-buffer indices that merely run past a row
 *     end behave exactly like the reference (linear addressing);
 *   - scale <= 1 (an endless loop in the reference) returns no detections.
 *
 * Build: gcc -std=c99 -O2 -ffp-contract=off -fPIC -shared (see Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int hdr[7];          /* mask, T, K, L, depth, stage, cart  (c/jda.c:499-505 reads and drops them) */
  int T, K, L, depth;
  int nn;              /* internal nodes per cart = 2^(depth-1) - 1 */
  int nl;              /* leaves per cart = 2^(depth-1) */
  float *mean_shape;   /* [2L] */
  /* per node, index (t*K + k)*nn + i */
  int *nd_scale, *nd_lm1, *nd_lm2, *nd_th;        /* lm* are already doubled (c/jda.c:521-523) */
  float *nd_o1x, *nd_o1y, *nd_o2x, *nd_o2y;
  /* per cart */
  float *leaf;         /* [(t*K+k)*nl + j] */
  float *cth, *cmean, *cstd;
  float *w;            /* [t][K*nl][2L] */
  int k_extra;         /* scan(): after the T full stages, carts [0, k_extra) of stage T without regression
                          (Validate's unfinished stage, src/jda/cascador.cpp:199-209); 0 in a loaded model */
} Model;

/* ---------------------------------------------------------------- model IO */

static int rd_i32(FILE *f, int *v) { return fread(v, 4, 1, f) == 1; }
static int rd_real(FILE *f, int dbl, float *v) {
  if (dbl) { double d; if (fread(&d, 8, 1, f) != 1) return 0; *v = (float)d; return 1; }
  return fread(v, 4, 1, f) == 1;
}

void jdo_free(void *m_) {
  Model *m = (Model *)m_;
  if (!m) return;
  free(m->mean_shape); free(m->nd_scale); free(m->nd_lm1); free(m->nd_lm2); free(m->nd_th);
  free(m->nd_o1x); free(m->nd_o1y); free(m->nd_o2x); free(m->nd_o2y);
  free(m->leaf); free(m->cth); free(m->cmean); free(m->cstd); free(m->w);
  free(m);
}

/* field order follows c/jda.c:499-556 (double) / :576-633 (float) */
void *jdo_load(const char *path, int dbl) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  Model *m = (Model *)calloc(1, sizeof(Model));
  int ok = 1;
  for (int i = 0; i < 7 && ok; i++) ok = rd_i32(f, &m->hdr[i]);
  m->T = m->hdr[1]; m->K = m->hdr[2]; m->L = m->hdr[3]; m->depth = m->hdr[4];
  if (!ok || m->T <= 0 || m->T > 64 || m->K <= 0 || m->K > 65536 || m->L <= 0 || m->L > 1024 ||
      m->depth < 2 || m->depth > 8) { fclose(f); free(m); return NULL; }
  m->nl = 1 << (m->depth - 1); m->nn = m->nl - 1;
  size_t C = (size_t)m->T * m->K, N = C * m->nn, D = 2 * (size_t)m->L;
  m->mean_shape = (float *)malloc(D * 4);
  m->nd_scale = (int *)malloc(N * 4); m->nd_lm1 = (int *)malloc(N * 4);
  m->nd_lm2 = (int *)malloc(N * 4);   m->nd_th = (int *)malloc(N * 4);
  m->nd_o1x = (float *)malloc(N * 4); m->nd_o1y = (float *)malloc(N * 4);
  m->nd_o2x = (float *)malloc(N * 4); m->nd_o2y = (float *)malloc(N * 4);
  m->leaf = (float *)malloc(C * m->nl * 4);
  m->cth = (float *)malloc(C * 4); m->cmean = (float *)malloc(C * 4); m->cstd = (float *)malloc(C * 4);
  m->w = (float *)malloc((size_t)m->T * m->K * m->nl * D * 4);
  for (size_t i = 0; i < D && ok; i++) ok = rd_real(f, dbl, &m->mean_shape[i]);
  for (int t = 0; t < m->T && ok; t++) {
    for (int k = 0; k < m->K && ok; k++) {
      size_t c = (size_t)t * m->K + k;
      for (int i = 0; i < m->nn && ok; i++) {
        size_t n = c * m->nn + i; int v;
        ok = ok && rd_i32(f, &m->nd_scale[n]);
        ok = ok && rd_i32(f, &v); m->nd_lm1[n] = v << 1;
        ok = ok && rd_i32(f, &v); m->nd_lm2[n] = v << 1;
        ok = ok && rd_real(f, dbl, &m->nd_o1x[n]) && rd_real(f, dbl, &m->nd_o1y[n]);
        ok = ok && rd_real(f, dbl, &m->nd_o2x[n]) && rd_real(f, dbl, &m->nd_o2y[n]);
        ok = ok && rd_i32(f, &m->nd_th[n]);
      }
      for (int j = 0; j < m->nl && ok; j++) ok = rd_real(f, dbl, &m->leaf[c * m->nl + j]);
      ok = ok && rd_real(f, dbl, &m->cth[c]) && rd_real(f, dbl, &m->cmean[c]) && rd_real(f, dbl, &m->cstd[c]);
    }
    size_t rows = (size_t)m->K * m->nl;
    float *wt = m->w + (size_t)t * rows * D;
    for (size_t i = 0; i < rows * D && ok; i++) ok = rd_real(f, dbl, &wt[i]);
  }
  fclose(f);
  if (!ok) { jdo_free(m); return NULL; }
  return m;
}

/* float32 flavour, byte layout of c/jda.c:644-716 (stage field written as T+1, cart as -1) */
int jdo_save_f32(void *m_, const char *path) {
  Model *m = (Model *)m_;
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  int h[7] = {0, m->T, m->K, m->L, m->depth, m->T + 1, -1};
  fwrite(h, 4, 7, f);
  size_t D = 2 * (size_t)m->L;
  fwrite(m->mean_shape, 4, D, f);
  for (int t = 0; t < m->T; t++) {
    for (int k = 0; k < m->K; k++) {
      size_t c = (size_t)t * m->K + k;
      for (int i = 0; i < m->nn; i++) {
        size_t n = c * m->nn + i;
        int a = m->nd_lm1[n] >> 1, b = m->nd_lm2[n] >> 1;
        fwrite(&m->nd_scale[n], 4, 1, f); fwrite(&a, 4, 1, f); fwrite(&b, 4, 1, f);
        fwrite(&m->nd_o1x[n], 4, 1, f); fwrite(&m->nd_o1y[n], 4, 1, f);
        fwrite(&m->nd_o2x[n], 4, 1, f); fwrite(&m->nd_o2y[n], 4, 1, f);
        fwrite(&m->nd_th[n], 4, 1, f);
      }
      fwrite(&m->leaf[c * m->nl], 4, m->nl, f);
      fwrite(&m->cth[c], 4, 1, f); fwrite(&m->cmean[c], 4, 1, f); fwrite(&m->cstd[c], 4, 1, f);
    }
    size_t rows = (size_t)m->K * m->nl;
    fwrite(m->w + (size_t)t * rows * D, 4, rows * D, f);
  }
  int z = 0; fwrite(&z, 4, 1, f);
  fclose(f);
  return 0;
}

void jdo_dims(void *m_, int *out4) {
  Model *m = (Model *)m_;
  out4[0] = m->T; out4[1] = m->K; out4[2] = m->L; out4[3] = m->depth;
}

/* ------------------------------------------------------------------ resize */

/* c/jda.c:203-230: ratio = (src-1)/dst, top-left tap by truncation, four
 * products summed left to right in float, cast to u8 by truncation. */
void jdo_resize(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh) {
  float xr = (float)(sw - 1) / dw;
  float yr = (float)(sh - 1) / dh;
  size_t o = 0;
  for (int i = 0; i < dh; i++) {
    for (int j = 0; j < dw; j++) {
      int x = (int)(xr * j);
      int y = (int)(yr * i);
      float xd = (xr * j) - x;
      float yd = (yr * i) - y;
      int idx = y * sw + x;
      int a = src[idx], b = src[idx + 1], c = src[idx + sw], d = src[idx + sw + 1];
      dst[o++] = (uint8_t)(a * (1.f - xd) * (1.f - yd) + b * (xd) * (1.f - yd) +
                           c * (1.f - xd) * (yd) + d * (xd) * (yd));
    }
  }
}

/* ------------------------------------------------------------------- levels */

/* window sizes visited by c/jda.c:320-332 (after the clamps of :459-460). */
int jdo_levels(int w, int h, float scale, int min_size, int max_size, int *wins, int cap) {
  if (min_size < 24) min_size = 24;
  if (max_size <= 0) max_size = (w < h) ? w : h;
  if (max_size > w) max_size = w;
  if (max_size > h) max_size = h;
  if (!(scale > 1.f)) return 0;
  int win = 24, n = 0;
  while (win < min_size) { int nw = (int)(win * scale); if (nw <= win) return 0; win = nw; }
  for (; win <= max_size;) {
    if (n < cap) wins[n] = win;
    n++;
    int nw = (int)(win * scale);
    if (nw <= win) break;
    win = nw;
  }
  return n;
}

/* ------------------------------------------------------------ cascade core */

typedef struct { const uint8_t *d; int w, h; } Plane;

typedef struct {
  long long windows, carts;      /* candidate windows; carts evaluated (incl. the rejecting one) */
  long long stage_survivors[64]; /* windows that completed stage t */
  long long ub_reads;            /* pixel reads that fell outside a plane buffer (defined as 0) */
} Stats;

/* one window, c/jda.c:356-414.  Returns 1 if it passes every cart threshold
 * (final threshold not applied here).  n_carts = carts evaluated.
 * leaves (optional) [T*K] gets the leaf index of every evaluated cart, 255 elsewhere. */
static int eval_window(const Model *m, const Plane pl[3], int x, int y, int win,
                       float *shape, float *score_out, int *n_carts, uint8_t *leaves, Stats *st) {
  const float r = 1.f / sqrtf(2.f);
  int ox[3], oy[3];
  ox[0] = x;            oy[0] = y;
  ox[1] = (int)(x * r); oy[1] = (int)(y * r);
  ox[2] = x / 2;        oy[2] = y / 2;
  const int D = 2 * m->L;
  memcpy(shape, m->mean_shape, D * sizeof(float));
  float score = 0.f;
  int evaluated = 0;
  int *lbf = (int *)malloc(m->K * sizeof(int));
  int pass = 1;
  const int t_end = m->T + (m->k_extra > 0 ? 1 : 0);
  for (int t = 0; t < t_end && pass; t++) {
    /* stage T of a truncated cascade stops after k_extra carts and has no regression
       (src/jda/cascador.cpp:199-209: carts [0, current_cart_idx] of the stage being trained) */
    const int Kt = t < m->T ? m->K : m->k_extra;
    for (int k = 0; k < Kt; k++) {
      size_t c = (size_t)t * m->K + k;
      int idx = 0;
      for (int lv = 0; lv < m->depth - 1; lv++) {
        size_t n = c * m->nn + idx;
        int l1 = m->nd_lm1[n], l2 = m->nd_lm2[n];
        float x1 = shape[l1] + m->nd_o1x[n];
        float y1 = shape[l1 + 1] + m->nd_o1y[n];
        float x2 = shape[l2] + m->nd_o2x[n];
        float y2 = shape[l2 + 1] + m->nd_o2y[n];
        int s = m->nd_scale[n];
        /* every view has w = win_size (c/jda.c:342,347,352) */
        int x1_ = (int)(x1 * win), y1_ = (int)(y1 * win);
        int x2_ = (int)(x2 * win), y2_ = (int)(y2 * win);
        if (x1_ < 0) x1_ = 0; else if (x1_ >= win) x1_ = win - 1;
        if (x2_ < 0) x2_ = 0; else if (x2_ >= win) x2_ = win - 1;
        if (y1_ < 0) y1_ = 0; else if (y1_ >= win) y1_ = win - 1;
        if (y2_ < 0) y2_ = 0; else if (y2_ >= win) y2_ = win - 1;
        const Plane *p = &pl[s];
        long long i1 = (long long)(oy[s] + y1_) * p->w + ox[s] + x1_;
        long long i2 = (long long)(oy[s] + y2_) * p->w + ox[s] + x2_;
        long long lim = (long long)p->w * p->h;
        int a = 0, b = 0;
        if (i1 < lim) a = p->d[i1]; else if (st) st->ub_reads++;
        if (i2 < lim) b = p->d[i2]; else if (st) st->ub_reads++;
        int feature = a - b;
        idx = (feature <= m->nd_th[n]) ? 2 * idx + 1 : 2 * idx + 2;
      }
      int leaf = idx - m->nn;
      evaluated++;
      if (leaves) leaves[c] = (uint8_t)leaf;
      score += m->leaf[c * m->nl + leaf];
      score = (score - m->cmean[c]) / m->cstd[c];
      if (score < m->cth[c]) { pass = 0; break; }
      lbf[k] = k * m->nl + leaf;
    }
    if (!pass || t >= m->T) break;
    if (st) st->stage_survivors[t]++;
    const float *wt = m->w + (size_t)t * m->K * m->nl * D;
    for (int k = 0; k < m->K; k++) {
      const float *row = wt + (size_t)lbf[k] * D;
      for (int i = 0; i < D; i++) shape[i] += row[i];
    }
  }
  free(lbf);
  *score_out = score;
  *n_carts = evaluated;
  if (st) { st->windows++; st->carts += evaluated; }
  return pass;
}

typedef struct { int n, cap; int *box; float *score; float *shape; } Hits;

static void hits_push(Hits *h, int D, int x, int y, int win, float score, const float *shape) {
  if (h->n == h->cap) {
    h->cap = h->cap ? 2 * h->cap : 256;
    h->box = (int *)realloc(h->box, (size_t)h->cap * 3 * sizeof(int));
    h->score = (float *)realloc(h->score, (size_t)h->cap * sizeof(float));
    h->shape = (float *)realloc(h->shape, (size_t)h->cap * D * sizeof(float));
  }
  int i = h->n++;
  h->box[3 * i] = x; h->box[3 * i + 1] = y; h->box[3 * i + 2] = win;
  h->score[i] = score;
  memcpy(h->shape + (size_t)i * D, shape, D * sizeof(float));
}

/*
 * Scan of c/jda.c:318-439 with instrumentation.
 *   t_limit  : stages to run (m->T for detection; fewer mirrors Validate's
 *              current_stage_idx loop for mining, src/jda/cascador.cpp:178-197)
 *   k_limit  : > 0: t_limit counts FULL stages (0 allowed) and carts [0, k_limit) of stage t_limit
 *              follow without regression (Validate's current_cart_idx + 1, cascador.cpp:199-209)
 *   use_th   : apply the final threshold (c/jda.c:414)
 *   trace_n  : optional [windows] carts evaluated per window, scan order
 *   trace_s  : optional [windows] score at exit per window
 *   trace_leaf, leaf_w0, leaf_w1 : optional leaf indices of windows [w0,w1), [T*K] each
 */
static void scan(const Model *m0, const uint8_t *img, int w, int h, float scale, int min_size,
                 int max_size, float th, int t_limit, int k_limit, int use_th, Hits *hits, Stats *st,
                 int *trace_n, float *trace_s, uint8_t *trace_leaf, long long leaf_w0, long long leaf_w1) {
  Model mm = *m0;
  mm.k_extra = 0;
  if (k_limit > 0 && t_limit >= 0 && t_limit < mm.T) {
    mm.T = t_limit;
    mm.k_extra = k_limit < mm.K ? k_limit : mm.K;
  } else if (t_limit > 0 && t_limit < mm.T) {
    mm.T = t_limit;
  }
  const Model *m = &mm;
  const float r = 1.f / sqrtf(2.f);
  int hw = (int)(w * r), hh = (int)(h * r), qw = w / 2, qh = h / 2;
  if (w < 24 || h < 24) return;
  uint8_t *hp = (uint8_t *)malloc((size_t)hw * hh), *qp = (uint8_t *)malloc((size_t)qw * qh);
  jdo_resize(img, w, h, hp, hw, hh);
  jdo_resize(img, w, h, qp, qw, qh);
  Plane pl[3] = {{img, w, h}, {hp, hw, hh}, {qp, qw, qh}};
  int wins[256];
  int nl = jdo_levels(w, h, scale, min_size, max_size, wins, 256);
  if (nl > 256) nl = 256;
  const int D = 2 * m->L;
  float *shape = (float *)malloc(D * sizeof(float));
  long long wi = 0;
  for (int li = 0; li < nl; li++) {
    int win = wins[li];
    int step = (int)(win * 0.1f);
    for (int y = 0; y <= h - win; y += step) {
      for (int x = 0; x <= w - win; x += step) {
        float score; int nc;
        uint8_t *lv = NULL;
        if (trace_leaf && wi >= leaf_w0 && wi < leaf_w1) {
          lv = trace_leaf + (size_t)(wi - leaf_w0) * m0->T * m0->K;
          memset(lv, 255, (size_t)m0->T * m0->K);
        }
        int pass = eval_window(m, pl, x, y, win, shape, &score, &nc, lv, st);
        if (trace_n) trace_n[wi] = nc;
        if (trace_s) trace_s[wi] = score;
        wi++;
        if (!pass) continue;
        if (use_th && score < th) continue;
        if (hits) hits_push(hits, D, x, y, win, score, shape);
      }
    }
  }
  free(shape); free(hp); free(qp);
}

/* c/jda.c:237-316: exchange sort by score (strict <), greedy IoU > 0.3 suppression,
 * survivors kept in their original (scan) order.  keep[] gets 0/1 per input. */
void jdo_nms(int n, const int *box, const float *score, uint8_t *keep) {
  const float overlap = 0.3f;
  int *idx = (int *)malloc((n > 0 ? n : 1) * sizeof(int));
  for (int i = 0; i < n; i++) { idx[i] = i; keep[i] = 1; }
  for (int i = 0; i < n - 1; i++)
    for (int j = i + 1; j < n; j++)
      if (score[idx[i]] < score[idx[j]]) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
  for (int i = 0; i < n - 1; i++) {
    int a = idx[i];
    if (!keep[a]) continue;
    int ax = box[3 * a], ay = box[3 * a + 1], as = box[3 * a + 2];
    for (int j = i + 1; j < n; j++) {
      int b = idx[j];
      if (!keep[b]) continue;
      int bx = box[3 * b], by = box[3 * b + 1], bs = box[3 * b + 2];
      int x1 = ax > bx ? ax : bx, y1 = ay > by ? ay : by;
      int x2 = (ax + as < bx + bs) ? ax + as : bx + bs;
      int y2 = (ay + as < by + bs) ? ay + as : by + bs;
      int iw = x2 - x1 > 0 ? x2 - x1 : 0, ih = y2 - y1 > 0 ? y2 - y1 : 0;
      float ov = (float)(iw * ih) / (float)(as * as + bs * bs - iw * ih);
      if (ov > overlap) keep[b] = 0;
    }
  }
  free(idx);
}

/* ----------------------------------------------------------- public surface */

typedef struct {
  int n, landmark_n;
  int *bboxes; float *shapes; float *scores;
} jdoResult;

void jdo_result_free(jdoResult r) { free(r.bboxes); free(r.shapes); free(r.scores); }

/* pre-NMS hits in scan order, shapes still window-normalised (c/jda.c:416-437) */
jdoResult jdo_detect_raw_k(void *m_, const uint8_t *img, int w, int h, float scale, int min_size,
                           int max_size, float th, int t_limit, int k_limit, int use_th, long long *stats_out) {
  Model *m = (Model *)m_;
  Hits hits = {0, 0, NULL, NULL, NULL};
  Stats st; memset(&st, 0, sizeof st);
  scan(m, img, w, h, scale, min_size, max_size, th, t_limit, k_limit, use_th, &hits, &st, NULL, NULL, NULL, 0, 0);
  if (stats_out) {
    stats_out[0] = st.windows; stats_out[1] = st.carts; stats_out[2] = st.ub_reads;
    for (int t = 0; t < 16; t++) stats_out[3 + t] = st.stage_survivors[t];
  }
  jdoResult r = {hits.n, m->L, hits.box, hits.shape, hits.score};
  return r;
}

jdoResult jdo_detect_raw(void *m_, const uint8_t *img, int w, int h, float scale, int min_size,
                         int max_size, float th, int t_limit, int use_th, long long *stats_out) {
  return jdo_detect_raw_k(m_, img, w, h, scale, min_size, max_size, th, t_limit, 0, use_th, stats_out);
}

/* full jdaDetect, c/jda.c:443-480 */
jdoResult jdo_detect(void *m_, const uint8_t *img, int w, int h, float scale, float step_ignored,
                     int min_size, int max_size, float th) {
  (void)step_ignored; /* shadowed at c/jda.c:333 */
  Model *m = (Model *)m_;
  jdoResult raw = jdo_detect_raw(m_, img, w, h, scale, min_size, max_size, th, 0, 1, NULL);
  const int D = 2 * m->L;
  uint8_t *keep = (uint8_t *)malloc(raw.n > 0 ? raw.n : 1);
  jdo_nms(raw.n, raw.bboxes, raw.scores, keep);
  jdoResult out;
  out.landmark_n = m->L;
  out.n = 0;
  out.bboxes = (int *)malloc((size_t)(raw.n > 0 ? raw.n : 1) * 3 * sizeof(int));
  out.scores = (float *)malloc((size_t)(raw.n > 0 ? raw.n : 1) * sizeof(float));
  out.shapes = (float *)malloc((size_t)(raw.n > 0 ? raw.n : 1) * D * sizeof(float));
  for (int i = 0; i < raw.n; i++) {
    if (!keep[i]) continue;
    int o = out.n++;
    memcpy(out.bboxes + 3 * o, raw.bboxes + 3 * i, 3 * sizeof(int));
    out.scores[o] = raw.scores[i];
    int x = raw.bboxes[3 * i], y = raw.bboxes[3 * i + 1], size = raw.bboxes[3 * i + 2];
    const float *s = raw.shapes + (size_t)i * D;
    float *d = out.shapes + (size_t)o * D;
    for (int j = 0; j < m->L; j++) {      /* c/jda.c:470-473, no FMA (-ffp-contract=off) */
      d[2 * j] = s[2 * j] * size + x;
      d[2 * j + 1] = s[2 * j + 1] * size + y;
    }
  }
  free(keep);
  jdo_result_free(raw);
  return out;
}

/* per-window trace: carts evaluated + exit score for every window (scan order),
 * leaf indices for windows [leaf_w0, leaf_w1).  Returns the window count. */
long long jdo_trace_k(void *m_, const uint8_t *img, int w, int h, float scale, int min_size, int max_size,
                      int t_limit, int k_limit, int *trace_n, float *trace_s, uint8_t *trace_leaf,
                      long long leaf_w0, long long leaf_w1) {
  Stats st; memset(&st, 0, sizeof st);
  scan((Model *)m_, img, w, h, scale, min_size, max_size, 0.f, t_limit, k_limit, 0, NULL, &st,
       trace_n, trace_s, trace_leaf, leaf_w0, leaf_w1);
  return st.windows;
}

long long jdo_trace(void *m_, const uint8_t *img, int w, int h, float scale, int min_size, int max_size,
                    int t_limit, int *trace_n, float *trace_s, uint8_t *trace_leaf,
                    long long leaf_w0, long long leaf_w1) {
  return jdo_trace_k(m_, img, w, h, scale, min_size, max_size, t_limit, 0, trace_n, trace_s, trace_leaf, leaf_w0, leaf_w1);
}

/* number of candidate windows enumerated by c/jda.c:332-339 */
long long jdo_count_windows(int w, int h, float scale, int min_size, int max_size) {
  int wins[256];
  int nl = jdo_levels(w, h, scale, min_size, max_size, wins, 256);
  if (w < 24 || h < 24) return 0;
  long long n = 0;
  for (int i = 0; i < nl && i < 256; i++) {
    int win = wins[i], step = (int)(win * 0.1f);
    n += (long long)((h - win) / step + 1) * ((w - win) / step + 1);
  }
  return n;
}
