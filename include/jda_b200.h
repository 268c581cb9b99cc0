/*
 * jda_b200.h -- C ABI of libjda_b200.so, the B200 (sm_100a) implementation of JDA's
 * sliding-window detect-and-align path.
 *
 * Part 1 is the reference's own C API, symbol for symbol (reference c/jda.h:18-68,
 * implemented there by c/jda.c:443-727).  A program linked against the reference's
 * libjda can be relinked against libjda_b200.so unchanged.
 *
 * Part 2 is additive (reference has no equivalent): many frames per call, frames already
 * resident in HBM, truncated-cascade / raw-hit output for hard-negative mining
 * (the Validate loop of src/jda/cascador.cpp:166-211 as used by src/jda/data.cpp:971-1012),
 * device/stream selection, work counters and kernel timings.
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 */
#ifndef JDA_B200_H_
#define JDA_B200_H_

#include <stddef.h>

#if defined(_MSC_VER)
#define JDA_API __declspec(dllexport)
#else
#define JDA_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================== Part 1: reference API */

/* replaces c/jda.h:18-24.  bboxes = int[3n] (x, y, size); shapes = float[2*landmark_n*n] in
 * image pixels; scores = float[n].  The three arrays are malloc'd by the library and owned by
 * the caller until jdaResultRelease.  Extension: n < 0 signals a CUDA failure (the reference
 * has no error channel); the arrays are then NULL and jdaB200LastError() says why. */
typedef struct {
  int n;
  int landmark_n;
  int *bboxes;
  float *shapes;
  float *scores;
} jdaResult;

/* replaces c/jda.c:486-561 / :563-638.  Same file layouts (README.md:84-111), same (float)
 * narrowing of the double flavour.  Returns NULL on open/short-read failure like the reference,
 * and additionally on a header this build cannot run (tree_depth != 4).  No CUDA work happens
 * here: the device is touched on the first detect call. */
JDA_API void *jdaCascadorCreateDouble(const char *model);
JDA_API void *jdaCascadorCreateFloat(const char *model);

/* replaces c/jda.c:644-716.  Byte-identical output (float32 flavour, stage field T+1, cart -1). */
JDA_API void jdaCascadorSerializeTo(void *cascador, const char *model);

/* replaces c/jda.c:718-720.  NULL-safe; also frees device tables, streams and scratch. */
JDA_API void jdaCascadorRelease(void *cascador);

/* replaces c/jda.c:443-480.  `data`: 8-bit gray, row-major, stride == width, borrowed.
 * `step` is accepted and ignored exactly as in the reference (c/jda.c:333 shadows it with 0.1).
 * min_size is clamped up to 24, max_size <= 0 means min(width, height) (c/jda.c:459-460).
 * scale <= 1 (an endless loop in the reference) returns n = 0.
 * Re-entrant: calls on one handle from several threads are serialised internally. */
JDA_API jdaResult jdaDetect(void *cascador, unsigned char *data, int width, int height,
                            float scale, float step, int min_size, int max_size, float th);

/* replaces c/jda.c:722-727.  NULL-safe. */
JDA_API void jdaResultRelease(jdaResult result);

/* ========================================================================= Part 2: additive */

enum {
  JDA_B200_DEVICE_INPUT = 1, /* `frames` is a device pointer (frames resident in HBM)           */
  JDA_B200_RAW_HITS = 2,     /* no NMS, no relocation: every passing window, scan order,        */
                             /* shapes window-normalised (what c/jda.c:416-437 collects)        */
  JDA_B200_NO_FINAL_TH = 4,  /* skip the final score threshold of c/jda.c:414 (mining)          */
  JDA_B200_NO_TMA = 8,       /* debug: fill shared-memory tiles with plain loads, not TMA       */
  JDA_B200_NO_STAGE0_SCAN = 16 /* debug: skip the stage-0 scan kernel, run every window through */
                             /* the generic per-window kernel                                   */
};

/* One batch of equally sized frames.  Frame f starts at frames + f*frame_stride, rows are
 * `pitch` bytes apart.  With JDA_B200_DEVICE_INPUT the TMA tile path needs pitch % 16 == 0 and
 * a 16-byte aligned base (otherwise tiles are filled by plain loads). */
typedef struct {
  int n_frames;
  int width, height;
  int pitch;
  size_t frame_stride;
  float scale;
  int min_size, max_size;
  float th;
  int t_limit; /* 0 = all stages; 1..T = only the first t_limit stages (Validate's current_stage_idx) */
  int flags;
  int k_limit; /* 0 = stages are run whole.  > 0: a cascade that stops INSIDE a stage, as JoinCascador::Validate does
                * while that stage is being trained (src/jda/cascador.cpp:178-209, caller btcart.cpp:146-152): t_limit
                * then counts the FULL stages (0 .. T-1, 0 allowed) and carts [0, k_limit) of stage t_limit follow
                * (k_limit = current_cart_idx + 1), with no regression after them.  Mining passes this together with
                * JDA_B200_RAW_HITS | JDA_B200_NO_FINAL_TH. */
} jdaB200Batch;

/* work counters + device timings of the last batch call on this handle */
typedef struct {
  long long windows;          /* candidate windows enumerated (c/jda.c:332-339)                */
  long long stage0_survivors; /* windows that passed all K carts of stage 0                    */
  long long raw_hits;         /* windows that passed everything (pre-NMS)                      */
  long long detections;       /* after NMS                                                     */
  float ms_h2d, ms_resize, ms_scan, ms_cascade, ms_d2h, ms_host; /* CUDA-event / host timings  */
  int scan_launches, cascade_launches, resize_launches;
  int n_levels;
  int levels_smem;            /* levels served from TMA-filled shared-memory tiles             */
} jdaB200Stats;

/* Detect on n_frames frames.  results[n_frames] are filled like jdaDetect would fill them, one per
 * frame (release each with jdaResultRelease).  Returns 0, or a negative value on failure
 * (results then have n = -1). */
JDA_API int jdaB200DetectBatch(void *cascador, const unsigned char *frames, const jdaB200Batch *batch,
                               jdaResult *results, jdaB200Stats *stats /* may be NULL */);

/* jdaB200DetectBatch with ONE result for the whole batch: counts[n_frames] detections per frame, then the
 * detections of all frames back to back (frame order, scan order inside a frame) in three arrays -- four
 * allocations per call instead of three per frame, and a shape FFI hosts can wrap without a per-frame loop.
 * Same NMS / relocation / flags as jdaB200DetectBatch.  On failure n_frames = -1 and the arrays are NULL. */
typedef struct {
  int n_frames;
  int total;      /* sum of counts */
  int landmark_n;
  int *counts;    /* int[n_frames]            */
  int *bboxes;    /* int[3 * total]           */
  float *scores;  /* float[total]             */
  float *shapes;  /* float[2*landmark_n*total] */
} jdaB200FlatResult;

JDA_API int jdaB200DetectBatchFlat(void *cascador, const unsigned char *frames, const jdaB200Batch *batch,
                                   jdaB200FlatResult *result, jdaB200Stats *stats /* may be NULL */);
JDA_API void jdaB200FlatResultRelease(jdaB200FlatResult *result);

/* The same in two halves, so that two batches can be in flight on one handle (SURVEY.md 8(b) "additive exports"):
 * jdaB200Submit copies the batch in and launches its kernels without waiting for them and returns a ticket (>= 0);
 * jdaB200Collect waits for that ticket, runs NMS / relocation and fills `result` exactly as jdaB200DetectBatchFlat
 * would.  Submit batch i + 1 before collecting batch i: its host -> device copy then runs beside the scan of batch i,
 * and the host post-processing of batch i beside the scan of batch i + 1.  At most two tickets are outstanding
 * (a third jdaB200Submit returns -3); tickets are collected in the order they were issued; `frames` (host or device)
 * must stay valid and unchanged until its ticket is collected; the synchronous entry points refuse to run while a
 * ticket is outstanding.  Negative return values are failures (jdaB200LastError). */
JDA_API int jdaB200Submit(void *cascador, const unsigned char *frames, const jdaB200Batch *batch);
JDA_API int jdaB200Collect(void *cascador, int ticket, jdaB200FlatResult *result, jdaB200Stats *stats /* may be NULL */);

/* Frames of different sizes in one call (reference: test.cpp:73-235 feeds FDDB images of all shapes one by one to
 * the detector; here they share one launch).  Host memory only.  pitch = 0 means pitch == width. */
typedef struct {
  const unsigned char *data;
  int width, height;
  int pitch;
} jdaB200Frame;

/* results[f] is what jdaDetect(frames[f].data, width, height, scale, ., min_size, max_size, th) returns for that
 * frame alone (max_size <= 0: each frame's own min(width, height), c/jda.c:459-460).  The frames are laid out in
 * slots of a common canvas in HBM and scanned by one launch per kernel; every frame keeps exactly the windows
 * c/jda.c:320-339 enumerates for its own size.  t_limit / flags as in jdaB200Batch (JDA_B200_DEVICE_INPUT is
 * ignored).  Returns 0, or a negative value on failure (results then have n = -1). */
JDA_API int jdaB200DetectMixed(void *cascador, const jdaB200Frame *frames, int n_frames, float scale, int min_size,
                               int max_size, float th, int t_limit, int flags, jdaResult *results,
                               jdaB200Stats *stats /* may be NULL */);

/* ---- the reference's double-precision C++ detector ------------------------------------
 * JoinCascador::Detect with fddb.method = 1 (src/jda/cascador.cpp:310-376, 431-477): fixed pixel step, window
 * ladder win = int(win * scale) from fddb.minimum_size, JoinCascador::Validate per window in double precision with
 * round()ed pixel coordinates (data.cpp:18-58), no final score threshold, multimap NMS (cascador.cpp:387-429),
 * results in pick order.  Scope: models whose nodes are all at scale 0 (the shipped model); anything else is refused
 * with an error.  A handle created from a float-flavour file runs the exactly widened values.
 * Validate is DETERMINISTIC here: every window starts from mean_shape + (shift_x, shift_y) of the parameters below --
 * (0, 0) is DataSet::RandomShape with shift_size = 0 as src/test.cpp:17,75 (test, fddb) force it.  src/live.cpp leaves
 * config.json's random_shift = 0.02 in place, a tick-count-seeded shift per window that no two runs of the reference
 * share; a caller who wants that behaviour passes the shift of its choice, one per call. */
typedef struct {
  int n;
  int landmark_n;
  int *rects;     /* int[4n]: x, y, width, height (cv::Rect) */
  double *scores; /* double[n] */
  double *shapes; /* double[2*landmark_n*n], image pixels */
} jdaB200ResultF64;

/* the fddb.* keys of config.json that JoinCascador::Detect reads (src/jda/common.cpp:178-188) */
typedef struct {
  int minimum_size; /* fddb.minimum_size (model/config.json: 20) */
  int step;         /* fddb.step         (5)                     */
  double scale;     /* fddb.scale        (1.2)                   */
  double overlap;   /* fddb.overlap      (0.3)                   */
  int nms;          /* fddb.nms          (true)                  */
  int flags;        /* 0, or JDA_B200_NO_STAGE0_SCAN: every window through the double-precision kernel */
  int similarity_transform; /* face.similarity_transform (config.json: false): offsets and the regressed delta go through
                               STParameter::Calc(shape, mean_shape) / Apply (data.cpp:64-126), recomputed per stage.
                               Calc's cv::norm is taken as the square root of the squares summed in index order (the
                               reference build this is pinned against uses that stand-in; OpenCV is not in the image). */
  double shift_x, shift_y;  /* the shift DataSet::RandomShape adds to the mean shape (data.cpp:225-236).  The reference
                               draws it per window from an RNG seeded with the tick count (face.random_shift; forced to 0
                               by src/test.cpp:17,75): pass the values a run should use, the same for every window.
                               Non-zero: every window runs through the double-precision kernel (no stage-0 prefilter). */
} jdaB200CppParams;

/* n_frames equally sized host frames (8-bit gray, stride == width, frame f at frames + f*width*height).
 * results[f] is what JoinCascador::Detect returns for frame f (release with jdaB200ResultF64Release).
 * Returns 0, or a negative value on failure (results then have n = -1). */
JDA_API int jdaB200JoinCascadorDetect(void *cascador, const unsigned char *frames, int n_frames, int width, int height,
                                      const jdaB200CppParams *params, jdaB200ResultF64 *results,
                                      jdaB200Stats *stats /* may be NULL */);
JDA_API void jdaB200ResultF64Release(jdaB200ResultF64 *results, int n);

/* per-window trace of JoinCascador::Validate (scan order, one frame): carts evaluated (its `n`) and the score at
 * exit; every window runs through the double-precision kernel.  Returns the window count or a negative value. */
JDA_API long long jdaB200JoinCascadorTrace(void *cascador, const unsigned char *frame, int width, int height,
                                           const jdaB200CppParams *params, int *carts_evaluated, double *exit_score);

/* Stage 0 of the double-precision detector is prefiltered by the float32 scan kernel with every cart threshold
 * lowered by margins[k] >= |float32 running score - double running score| over all windows; survivors are then
 * evaluated exactly in double from cart 0, so the filter never decides a result.  Writes the K margins (host
 * only; for tests).  Returns K, 0 when this model runs without the prefilter, negative on failure. */
JDA_API int jdaB200JoinCascadorFilterMargins(void *cascador, double *margins, int cap);

/* window sizes detectMultiScale1 visits for a (w, h) frame (cascador.cpp:335,372-373); host only */
JDA_API int jdaB200JoinCascadorLevels(int width, int height, int minimum_size, double scale, int *wins, int cap);

/* jdaResultRelease for a whole array of results (one call instead of n). */
JDA_API void jdaB200ResultsRelease(jdaResult *results, int n);

/* Bind the handle to a CUDA device (default: device 0 / current at first use) and, optionally, to
 * a caller-owned cudaStream_t (NULL = the handle's own stream). */
JDA_API int jdaB200SetDevice(void *cascador, int device);
JDA_API int jdaB200SetStream(void *cascador, void *cuda_stream);

/* jdaCascadorSerializeTo with options.  The reference's writer stores the header's stage
 * field as T + 1 (c/jda.c:662-665), which the C++ loader refuses (cascador.cpp:138 wants T), and only knows the
 * float32 flavour.  flags = 0 writes exactly what jdaCascadorSerializeTo writes.  Returns 0, negative on failure. */
enum {
  JDA_B200_SAVE_STAGE_T = 1, /* header stage field = T                                                    */
  JDA_B200_SAVE_DOUBLE = 2   /* double flavour (README.md:84-111); with STAGE_T: loadable by the C++ tree */
};
JDA_API int jdaB200SerializeTo(void *cascador, const char *model, int flags);

/* model dimensions: out[0..3] = T, K, landmark_n, tree_depth (2..6 accepted; the reference fixes 4, c/jda.c:28) */
JDA_API void jdaB200ModelDims(void *cascador, int *out4);

/* jdaDetect (part 1) may be called from any number of host threads on one handle, like the reference's (c/jda.c:443-480
 * keeps no state in the cascador).  Calls that arrive while an earlier one is running are coalesced: the caller that finds
 * nobody serving runs every queued call with the same (scale, min_size, max_size, th) as ONE mixed-size batch (each
 * frame's result is bit for bit what it gives alone) and wakes the others; a lone caller is served at once.  Counters since
 * the handle was created: jdaDetect calls, device batches they were served in, frames of the largest batch. */
JDA_API void jdaB200CoalescingStats(void *cascador, long long *calls, long long *batches, int *largest);

/* last error text of the calling thread ("" if none) */
JDA_API const char *jdaB200LastError(void);

/* number of visible CUDA devices (0 on a CPU-only box; never fails) */
JDA_API int jdaB200DeviceCount(void);

/* ---- host-side helpers exported for tests and tools (no device needed) -------------------- */

/* window sizes visited for a (w, h) frame: the loop of c/jda.c:320-332.  Returns the count. */
JDA_API int jdaB200Levels(int width, int height, float scale, int min_size, int max_size,
                          int *wins, int cap);
/* candidate windows for a (w, h) frame (c/jda.c:332-339) */
JDA_API long long jdaB200CountWindows(int width, int height, float scale, int min_size, int max_size);
/* the greedy NMS of c/jda.c:237-316: keep[i] = 1 for survivors (scan order preserved) */
JDA_API void jdaB200Nms(int n, const int *bboxes, const float *scores, unsigned char *keep);

/* The scan kernel's tile plan for a (w, h) frame, one text line per pyramid level:
 * "win step nx ny tw th box_w box_h smem windows span".  Returns the number of levels (host only).
 * cap > 0: the throughput plan (batches of more than 4 frames); cap < 0: the latency plan, |cap| bytes. */
JDA_API int jdaB200DescribePlan(int width, int height, float scale, int min_size, int max_size,
                                char *buf, int cap);

/* ---- per-window trace (tests): scan order, one frame ------------------------------------- */
/* Runs the full device path on one host frame and records, for every candidate window, the
 * number of carts evaluated and the score at exit; for windows [leaf_w0, leaf_w1) also the leaf
 * index of every evaluated cart ([T*K] bytes each, 255 = not evaluated).  Returns the window
 * count or a negative value. */
JDA_API long long jdaB200Trace(void *cascador, const unsigned char *frame, int width, int height,
                               float scale, int min_size, int max_size, int t_limit, int flags,
                               int *carts_evaluated, float *exit_score, unsigned char *leaves,
                               long long leaf_w0, long long leaf_w1);
/* the same with jdaB200Batch's k_limit (a cascade that stops inside stage t_limit) */
JDA_API long long jdaB200TraceK(void *cascador, const unsigned char *frame, int width, int height,
                                float scale, int min_size, int max_size, int t_limit, int k_limit, int flags,
                                int *carts_evaluated, float *exit_score, unsigned char *leaves,
                                long long leaf_w0, long long leaf_w1);

/* Device bilinear down-sample (replaces jdaImageResize, c/jda.c:203-230); host in, host out. */
JDA_API int jdaB200Resize(void *cascador, const unsigned char *src, int sw, int sh,
                          unsigned char *dst, int dw, int dh);

#ifdef __cplusplus
}
#endif
#endif /* JDA_B200_H_ */
