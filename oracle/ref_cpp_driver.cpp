// C ABI around the reference's own C++ detector (JoinCascador, src/jda/cascador.cpp) as built by oracle/Makefile into
// oracle/_ref_cpp/libjda_ref_cpp.so.  TEST INFRASTRUCTURE ONLY: tests/test_oracle_cpp.py uses it to pin the
// restatement oracle/jda_oracle_cpp.c (and through it the CUDA double-precision path) against the real thing.
//
// Everything numerical below is the reference's code, called through its public class interface:
//   JoinCascador::SerializeFrom, ::Detect (detectMultiScale1, Validate, nms, relocation), ::Validate.
// This file only (a) points the Config singleton at a config.json and overrides the public fields a caller wants
// (dimensions from the model header -- the shipped model/config.json says landmark_n = 5 for a 27-landmark model --,
// shift_size = 0 exactly as src/test.cpp:17,75 do, the fddb.* keys), (b) wraps the caller's image in a cv::Mat header,
// (c) copies results out, and (d) for the per-window trace walks the windows in detectMultiScale1's own order
// (cascador.cpp:335-373) calling the reference's Validate on the same three views.
#include <stdio.h>
#include <unistd.h>

#include <cmath>
#include <string>
#include <vector>

#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include "jda/data.hpp"
#include "jda/cart.hpp"
#include "jda/common.hpp"
#include "jda/cascador.hpp"

using namespace jda;

// The two whole files reference five trainer-side functions whose definitions live in the parts of data.cpp /
// btcart.cpp that are NOT built here (image loaders, liblinear).  Nothing on the detect path calls them.
namespace jda {
static void not_built(const char *what) {
  fprintf(stderr, "jref: %s belongs to the trainer, which oracle/_ref_cpp does not build\n", what);
  abort();
}
void DataSet::CalcSTParameters(const cv::Mat_<double> &) { not_built("DataSet::CalcSTParameters"); }
void DataSet::Snapshot(const DataSet &, const DataSet &) { not_built("DataSet::Snapshot"); }
void BoostCart::Train(DataSet &, DataSet &) { not_built("BoostCart::Train"); }
cv::Mat_<int> DataSet::CalcFeatureValues(const std::vector<Feature> &, const std::vector<int> &) const {
  not_built("DataSet::CalcFeatureValues");
  return cv::Mat_<int>();
}
cv::Mat_<double> DataSet::CalcShapeResidual(const std::vector<int> &, int) const {
  not_built("DataSet::CalcShapeResidual");
  return cv::Mat_<double>();
}
}  // namespace jda

namespace {
Config &config_at(const char *run_dir) {
  // Config::Config() reads "../config.json" relative to the working directory (common.cpp:117)
  char cwd[4096];
  if (!getcwd(cwd, sizeof cwd)) cwd[0] = 0;
  if (chdir(run_dir) != 0) fprintf(stderr, "jref: cannot chdir to %s\n", run_dir);
  Config &c = Config::GetInstance();
  if (cwd[0] && chdir(cwd) != 0) fprintf(stderr, "jref: cannot restore the working directory\n");
  return c;
}
}  // namespace

extern "C" {

// model: double-flavour file (README.md:84-111).  run_dir: a directory whose parent holds config.json.
// Returns NULL when the file cannot be read or its header is not what SerializeFrom would accept (it would exit()).
void *jref_open(const char *model, const char *run_dir) {
  FILE *fd = fopen(model, "rb");
  if (!fd) return NULL;
  int hdr[7];
  if (fread(hdr, 4, 7, fd) != 7) { fclose(fd); return NULL; }
  rewind(fd);
  if (hdr[1] <= 0 || hdr[2] <= 0 || hdr[3] <= 0 || hdr[4] < 2 || hdr[5] < 0 || hdr[5] > hdr[1] || hdr[6] < -1 ||
      hdr[6] >= hdr[2]) { fclose(fd); return NULL; }
  Config &c = config_at(run_dir);
  c.T = hdr[1]; c.K = hdr[2]; c.landmark_n = hdr[3]; c.tree_depth = hdr[4];
  c.shift_size = 0.;  // src/test.cpp:17,75
  JoinCascador *jc = new JoinCascador();
  jc->SerializeFrom(fd);
  fclose(fd);
  return jc;
}

void jref_close(void *h) { delete (JoinCascador *)h; }

// Initial shift (data.cpp:225-236): shift_size as in config.json's face.random_shift, with the tick that seeds
// RandomShape's RNG fixed so that every window -- and the test -- draws the same (x, y).  tick = 0: real ticks.
// xy (may be NULL) receives the shift the reference's own RNG then produces.
void jref_set_shift(double shift_size, long long tick, double *xy) {
  Config::GetInstance().shift_size = shift_size;
  cv::shim_fixed_tick() = (cv::int64)tick;
  if (xy) {
    cv::RNG rng = cv::RNG(cv::getTickCount());
    xy[0] = rng.uniform(-shift_size, shift_size);
    xy[1] = rng.uniform(-shift_size, shift_size);
  }
}

void jref_dims(void *h, int *out6) {
  const JoinCascador *jc = (const JoinCascador *)h;
  out6[0] = jc->T; out6[1] = jc->K; out6[2] = jc->landmark_n; out6[3] = jc->tree_depth;
  out6[4] = jc->current_stage_idx; out6[5] = jc->current_cart_idx;
}

static void set_fddb(int minimum_size, int step, double scale, double overlap, int nms, int method, int similarity) {
  Config &c = Config::GetInstance();
  c.fddb_minimum_size = minimum_size; c.fddb_step = step; c.fddb_scale_factor = scale; c.fddb_overlap = overlap;
  c.fddb_nms = nms != 0; c.fddb_detect_method = method; c.with_similarity_transform = similarity != 0;
}

// JoinCascador::Detect.  rects: x y w h per face; shapes: 2L doubles per face (image pixels); stats4: patch_n,
// face_patch_n, nonface_patch_n, cart_gothrough_n.  Returns the face count; release with jref_release.
int jref_detect(void *h, const unsigned char *img, int w, int hh, int minimum_size, int step, double scale, double overlap,
                int nms, int similarity, int **rects, double **scores, double **shapes, double *stats4) {
  const JoinCascador *jc = (const JoinCascador *)h;
  set_fddb(minimum_size, step, scale, overlap, nms, 1, similarity);
  cv::Mat gray(hh, w, cv::CV_8UC1, (void *)img);
  std::vector<cv::Rect> r;
  std::vector<double> s;
  std::vector<cv::Mat_<double> > sh;
  DetectionStatisic st;
  const int n = jc->Detect(gray, r, s, sh, st);
  const int D = 2 * jc->landmark_n;
  *rects = (int *)malloc(sizeof(int) * 4 * (n > 0 ? n : 1));
  *scores = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
  *shapes = (double *)malloc(sizeof(double) * D * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    (*rects)[4 * i] = r[i].x; (*rects)[4 * i + 1] = r[i].y; (*rects)[4 * i + 2] = r[i].width; (*rects)[4 * i + 3] = r[i].height;
    (*scores)[i] = s[i];
    for (int j = 0; j < D; j++) (*shapes)[(size_t)i * D + j] = sh[i](0, j);
  }
  if (stats4) { stats4[0] = st.patch_n; stats4[1] = st.face_patch_n; stats4[2] = st.nonface_patch_n; stats4[3] = st.cart_gothrough_n; }
  return n;
}

void jref_release(int *rects, double *scores, double *shapes) { free(rects); free(scores); free(shapes); }

// Per-window trace: the reference's Validate on every window detectMultiScale1 visits, in its order (window size
// outer, y, then x); carts evaluated (Validate's n) and the score at exit.  Returns the window count.
long long jref_trace(void *h, const unsigned char *img, int w, int hh, int minimum_size, int step, double scale,
                     int similarity, int *carts, double *score_out, long long cap) {
  const JoinCascador *jc = (const JoinCascador *)h;
  set_fddb(minimum_size, step, scale, 0.3, 1, 1, similarity);
  cv::Mat gray(hh, w, cv::CV_8UC1, (void *)img);
  cv::Mat img_o = gray.clone(), img_h, img_q;
  cv::resize(gray, img_h, cv::Size(int(w / std::sqrt(2.)), int(hh / std::sqrt(2.))));
  cv::resize(gray, img_q, cv::Size(w / 2, hh / 2));
  long long n = 0;
  int win = minimum_size;
  while (win <= w && win <= hh) {
    for (int y = 0; y <= hh - win; y += step)
      for (int x = 0; x <= w - win; x += step) {
        const double r = std::sqrt(2.);
        cv::Mat po = img_o(cv::Rect(x, y, win, win));
        cv::Mat ph = img_h(cv::Rect(int(x / r), int(y / r), int(win / r), int(win / r)));
        cv::Mat pq = img_q(cv::Rect(x / 2, y / 2, win / 2, win / 2));
        double score;
        cv::Mat_<double> shape;
        int len = 0;
        jc->Validate(po, ph, pq, score, shape, len);
        if (n < cap) { if (carts) carts[n] = len; if (score_out) score_out[n] = score; }
        n++;
      }
    win = int(win * scale);
  }
  return n;
}

}  // extern "C"
