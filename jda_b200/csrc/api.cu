// api.cu -- C ABI of libjda_b200.so (include/jda_b200.h): context, launches, post-processing.
//
// Call path of one detect (reference c/jda.c:443-480 -> here):
//   frames (host: H2D into a 16-byte-pitched device store | device: used in place)
//   [k1_resize  -> h/q planes, only if the model samples them]        c/jda.c:450-457
//   k2_scan     -> stage-0 survivors (queue in HBM)                    c/jda.c:332-402, t = 0
//   k3_cascade  -> raw hits (records in HBM)                           c/jda.c:356-427
//   D2H of the hit records, sort into scan order, host NMS + relocation c/jda.c:237-316, 465-474
#include <cuda.h>
#include <cuda_runtime.h>

#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/jda_b200.h"
#include "host_model.hpp"
#include "kernels.cuh"
#include "kernels_stages.cuh"
#include "kernels_regress.cuh"
#include "kernels_f64.cuh"

using namespace jda;

namespace {

thread_local std::string g_err;

void set_err(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  fprintf(stderr, "[jda_b200] error: %s\n", buf);
}

#define CU_OK(call)                                                                       \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return false;                                                                       \
    }                                                                                     \
  } while (0)

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  bool ensure(size_t n) {
    if (n <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 64;
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e != cudaSuccess) {
      set_err("cudaMalloc(%zu) failed: %s", want * sizeof(T), cudaGetErrorString(e));
      return false;
    }
    cap = want;
    return true;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct Geometry {
  int w = 0, h = 0, min_size = 0, max_size = 0;
  float scale = 0.f;
  int plan = 0;          // which tile plan (see plan_level): 0 throughput, 1 latency, 2 throughput with short global-memory tiles
  int step64 = 0;        // double-precision detector: fixed step and double scale factor (0 = the C path's ladder)
  double factor64 = 0.;
  int n_levels = 0;
  LevelInfo lv[kMaxLevels];
  long long windows_per_frame = 0;
  int table_bytes = 0;  // per level, padded to 128
  bool valid = false;
};

struct HitRec {
  int frame;
  uint32_t key;
  int x, y, win;
  float score;
  const float *shape;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

constexpr int kMaxChunks = 4;  // host batches are copied and scanned in up to 4 overlapping chunks

// A/B knobs and test hooks, read from the environment ONCE per handle (ctx_init) -- never on the call path.
// None of them changes results; the test hooks only force rarely taken host paths (queue growth, unchunked copies).
struct Tuning {
  int nw = 4;                     // JDA_B200_NW: windows per lane in k2_scan (1, 2, 4)
  int stragglers = 1;             // JDA_B200_STRAGGLERS
  int min_tile_windows = 128;     // JDA_B200_MIN_TILE_WINDOWS: a tile with fewer windows per warp pools buffers instead (r1n)
  int max_span = -1;              // JDA_B200_MAX_SPAN: -1 = by plan (throughput: up to 4 warps' buffers; latency: the whole block's)
  int latency_tile_windows = 256; // JDA_B200_LATENCY_TILE
  int global_tile_rows = 0;       // JDA_B200_GLOBAL_TILE_ROWS: window rows (of 32) per global-memory virtual tile; 0 = by plan
  int tune_pitch = 0;             // JDA_B200_TUNE_PITCH: also try wider pitches and pick by the bank-conflict model (no gain, r1)
  std::string pitch_extra;        // JDA_B200_PITCH_EXTRA="24:16,30:32": extra tile pitch per window size (A/B)
  bool even_chunks = false;       // JDA_B200_EVEN_CHUNKS
  bool no_chunks = false;         // JDA_B200_NO_CHUNKS (test hook)
  bool tiny_queues = false;       // JDA_B200_TINY_QUEUES (test hook: start with queues that overflow)
  bool old_regress = false;       // JDA_B200_OLD_REGRESS: the regression gather through k3_stage0 (round 1 kernel) instead of k3_regress (A/B)
  long long stage_min_windows = 60000000;  // JDA_B200_STAGE_MIN_WINDOWS: batches with fewer candidate windows take k3_cascade
                                  // for stages >= 1 (ten small launches are pure latency on a small survivor list:
                                  // 16 VGA frames 0.73 vs 0.37 ms, 256 frames 1.50 vs 1.41, 512 frames 2.26 vs 2.47;
                                  // profiles/r3d_small_batch.txt)
  bool no_stage_kernels = false;  // JDA_B200_NO_STAGE_KERNELS: stages >= 1 through k3_cascade (one warp per survivor) as in round 1 (A/B)
  double level_weight_exp = 1.0;  // JDA_B200_LEVEL_WEIGHT_EXP (r1q: scan -1.7 % against flat weights)
  std::string sched;              // JDA_B200_SCHED="4,8,16,...": phase ends of k2_scan
  int force_plan = 0;             // JDA_B200_FORCE_PLAN=latency|throughput (test hook: the per-window trace is a one-frame
                                  // call and would otherwise only ever see the latency tile plan)
};

Tuning read_tuning() {
  Tuning t;
  auto geti = [](const char *name, int &v) { if (const char *e = getenv(name)) v = atoi(e); };
  int nw = t.nw;
  geti("JDA_B200_NW", nw);
  if (nw == 1 || nw == 2 || nw == 4) t.nw = nw;
  geti("JDA_B200_STRAGGLERS", t.stragglers);
  t.stragglers = t.stragglers ? 1 : 0;
  geti("JDA_B200_MIN_TILE_WINDOWS", t.min_tile_windows);
  t.min_tile_windows = std::max(1, t.min_tile_windows);
  if (getenv("JDA_B200_MAX_SPAN")) { geti("JDA_B200_MAX_SPAN", t.max_span); t.max_span = std::max(1, t.max_span); }
  geti("JDA_B200_LATENCY_TILE", t.latency_tile_windows);
  geti("JDA_B200_GLOBAL_TILE_ROWS", t.global_tile_rows);
  t.global_tile_rows = std::max(0, std::min(K2_LIST_CAP / 32, t.global_tile_rows));
  t.latency_tile_windows = std::max(64, std::min(K2_LIST_CAP, t.latency_tile_windows));
  geti("JDA_B200_TUNE_PITCH", t.tune_pitch);
  if (const char *e = getenv("JDA_B200_PITCH_EXTRA")) t.pitch_extra = e;
  t.even_chunks = getenv("JDA_B200_EVEN_CHUNKS") != nullptr;
  t.no_stage_kernels = getenv("JDA_B200_NO_STAGE_KERNELS") != nullptr;
  if (const char *e = getenv("JDA_B200_STAGE_MIN_WINDOWS")) t.stage_min_windows = atoll(e);
  t.old_regress = getenv("JDA_B200_OLD_REGRESS") != nullptr;
  t.no_chunks = getenv("JDA_B200_NO_CHUNKS") != nullptr;
  if (const char *e = getenv("JDA_B200_TINY_QUEUES")) t.tiny_queues = atoi(e) != 0;
  if (const char *e = getenv("JDA_B200_LEVEL_WEIGHT_EXP")) t.level_weight_exp = atof(e);
  if (const char *e = getenv("JDA_B200_SCHED")) t.sched = e;
  if (const char *e = getenv("JDA_B200_FORCE_PLAN")) t.force_plan = e[0] == 'l' ? 1 : e[0] == 't' ? 2 : 0;
  return t;
}

constexpr int kSlots = 2;
constexpr int kCntSurv = kMaxChunks * kMaxLevels, kCntWork = kCntSurv + kMaxChunks, kCntHit = kCntWork + kMaxChunks,
              kCntStage = kCntHit + 1,                     // [kMaxStages + 1] windows that passed stage t (kernels_stages.cuh)
              kCntStageWork = kCntStage + kMaxStages + 1,  // [kMaxStages + 1] next list position k3_walk hands out at stage t
              kCntTotal = kCntStageWork + kMaxStages + 1;

struct Run;

// Device / pinned scratch of one batch in flight.
struct Scratch {
  DevBuf<uint8_t> d_frames, d_hq;
  DevBuf<int2> d_dims;             // mixed-size batches: (width, height) per frame
  DevBuf<uint8_t> d_packed;        // mixed-size batches: frames as they lie in host memory, back to back
  DevBuf<UnpackFrame> d_unpack;
  std::vector<UnpackFrame> h_unpack;
  DevBuf<uint4> d_surv;
  DevBuf<float> d_shape0;
  DevBuf<uint8_t> d_surv_leaves;
  DevBuf<uint2> d_stage_list[2];   // {queue entry, score} of the windows entering stage t (t & 1), kernels_stages.cuh
  DevBuf<float> d_hits;
  unsigned *d_counters = nullptr;  // [kMaxChunks][kMaxLevels] tile counters, surv_count, work, hit_count
  unsigned *h_counters = nullptr;  // pinned mirror
  std::vector<float> h_hits;
  float *h_eager = nullptr;        // pinned: the first kEagerHits hit records
  uint8_t *h_stage = nullptr;      // pinned staging for pageable inputs (8 MB to start with, grows to the largest batch seen)
  size_t h_stage_cap = 0;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_copy[kMaxChunks + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_done = nullptr;   // counters + first hit records of the batch are in pinned memory
  jdaB200Stats last;               // work counters + timings of the batch that used this set last
  bool ready = false;              // events and pinned buffers exist
  // a submitted batch waiting for jdaB200Collect
  bool busy = false;
  int ticket = -1;
  jdaB200Batch batch;
  const unsigned char *frames = nullptr;
  Run *run = nullptr;
};

// One jdaDetect call waiting to be served (see jdaDetect: concurrent callers are coalesced into batches).
struct DetectReq {
  const unsigned char *data;
  int width, height;
  float scale;
  int min_size, max_size;
  float th;
  jdaResult res;
  bool done;
};

struct Context {
  HostModel m;
  std::mutex mu;
  // jdaDetect's combiner: callers queue here; whoever finds no leader serves the oldest group of the queue
  std::mutex cq_mu;
  std::condition_variable cq_cv;
  std::vector<DetectReq *> cq;
  bool cq_leader = false;
  long long cq_calls = 0, cq_batches = 0;
  int cq_largest = 0;
  int device = -1;
  bool inited = false;
  cudaStream_t own_stream = nullptr, user_stream = nullptr, copy_stream = nullptr;
  int sm_count = 148;
  EncodeTiledFn encode = nullptr;
  // device model
  NodeRec *d_nodes = nullptr;
  float *d_leaf = nullptr;
  float4 *d_cart = nullptr;
  float *d_w = nullptr;
  float *d_wp = nullptr;  // w with every row padded to a multiple of four floats (16-byte rows for k3_regress)
  float *d_mean = nullptr;
  // geometry + stage-0 tables
  Geometry geo;
  DevBuf<uint8_t> d_tables;
  Stage0Norm *d_norms = nullptr;
  // scratch: two sets, so that a submitted batch (jdaB200Submit) can be copied in and scanned while the results of
  // the one before are still being collected; the synchronous entry points use set 0.  `sc` = the set in use.
  Scratch slot[kSlots];
  Scratch *sc = &slot[0];
  DevBuf<uint8_t> d_trace_leaf;    // (tracing is synchronous: one copy)
  DevBuf<int> d_trace_n;
  DevBuf<float> d_trace_s;
  size_t surv_cap = 0, hit_cap = 0;  // queue capacities (shared: a set's buffers are grown to them before use)
  cudaStream_t d2h_stream = nullptr; // read-back of hit records beyond the first few: must not queue behind the next batch
  int next_ticket = 0;
  // double-precision detector (JoinCascador::Detect, method 1): model kept as doubles, its own geometry and tables
  std::string path;
  bool path_dbl = false;
  HostModelD md;
  bool md_failed = false;
  NodeRecD *d_nodes64 = nullptr;
  double *d_leaf64 = nullptr, *d_cart64 = nullptr, *d_w64 = nullptr, *d_mean64 = nullptr;
  Geometry geo64;
  DevBuf<uint8_t> d_tables64;
  Stage0Norm *d_norms64 = nullptr;
  bool filter64_ok = false;        // stage 0 may be prefiltered by k2_scan (margins exist, <= kMaxNorm normalised carts)
  std::vector<double> margins64;
  DevBuf<double> d_hits64, d_trace_s64;
  DevBuf<int> d_trace_n64;
  size_t hit64_cap = 0;
  std::vector<double> h_hits64;
  Tuning tune;
  bool model64_ready = false;
  std::vector<short> sched;
  cudaStream_t stream() const { return user_stream ? user_stream : own_stream; }
};

constexpr size_t kEagerHits = 64;
constexpr int kLatencyFrames = 4;  // batches this small use the latency tile plan
// ... and so do batches of more, smaller frames up to this many candidate windows (12 VGA 3-octave frames): measured per
// call, latency / throughput plan -- 5 frames 0.85 / 1.35 ms, 8: 1.03 / 1.39, 12: 1.30 / 1.42, 16: 1.55 / 1.49
// (profiles/r3f_small_batch.txt)
constexpr long long kLatencyWindows = 2100000;

// ... and batches up to this many (47 VGA frames) keep the throughput plan but read the global-memory levels in
// 128-window virtual tiles instead of 512-window ones (plan 2): scan of 16 frames 1.03 -> 0.86 ms, 24: 1.39 -> 1.23,
// 32: 1.63 -> 1.59, 64: 2.99 -> 3.01 (profiles/r3h_small_batch_default.txt against r3g_)
constexpr long long kMediumWindows = 8000000;

// which tile plan a batch takes (plan_level): 0 throughput, 1 latency (also skips the cohort-staged stage 0), 2 medium
static int batch_plan(const Tuning &tn, const jdaB200Batch &b, const jdaB200Frame *mixed) {
  if (tn.force_plan) return tn.force_plan == 1 ? 1 : 0;
  if (b.n_frames <= kLatencyFrames) return 1;
  long long w = 0;
  if (mixed) {
    for (int f = 0; f < b.n_frames && w <= kMediumWindows; f++)
      w += count_windows(mixed[f].width, mixed[f].height, b.scale, b.min_size, b.max_size);
  } else {
    w = count_windows(b.width, b.height, b.scale, b.min_size, b.max_size) * b.n_frames;
  }
  return w <= kLatencyWindows ? 1 : w <= kMediumWindows ? 2 : 0;
}
// First frame of chunk `ch` when a host batch is copied and scanned in `nchunks` pieces.  The pieces grow
// (1/8, 2/8, 2/8, 3/8 of the batch): the first scan can only start when the first piece has landed, so it is small.
int chunk_begin(int n_frames, int ch, int nchunks, bool even = false) {
  if (nchunks != kMaxChunks) return (int)((long long)n_frames * ch / nchunks);
  static const int cum[kMaxChunks + 1] = {0, 1, 3, 5, 8};
  if (even) return (int)((long long)n_frames * ch / nchunks);
  return (int)((long long)n_frames * cum[ch] / 8);
}

size_t k2_smem_bytes(int table_bytes) { return (size_t)table_bytes + 256 + (size_t)K2_WARPS * K2_WARP_BYTES; }
size_t k3s_smem_bytes(int K, int D) {
  return (size_t)K3S_COHORT * leaf_bytes(K) + (size_t)2 * K3S_CHUNK * kLeaves * D * 4;
}
constexpr size_t kSmemOptIn = 227 * 1024;  // dynamic shared memory a block may opt in to on sm_100
// k3_walk: warps (= batches of survivors in flight) that fit next to the stage's tables; 0 = the tables do not fit
int k3w_warps(int K, bool trace) {
  const size_t tab = (k3w_table_bytes(K) + 15) & ~(size_t)15;
  if (tab + 8 * k3w_warp_bytes(trace) > kSmemOptIn) return 0;
  return (int)std::min<size_t>(K3W_MAX_WARPS, (kSmemOptIn - tab) / k3w_warp_bytes(trace));
}
size_t k3_smem_bytes(int K) { return (size_t)K3_WARPS * (kMaxDim * 4 + ((K + 15) & ~15)); }

void ctx_release_device(Context *c);

template <typename T>
void dev_free(T *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}
template <typename T>
void host_free(T *&p) {
  if (p) cudaFreeHost(p);
  p = nullptr;
}

// events, counters and pinned staging of one scratch set (set 0 with the handle, set 1 at the first jdaB200Submit)
bool slot_init(Scratch &sc) {
  if (sc.ready) return true;
  for (auto &e : sc.ev) CU_OK(cudaEventCreate(&e));
  for (auto &e : sc.ev_copy) CU_OK(cudaEventCreate(&e));
  CU_OK(cudaEventCreateWithFlags(&sc.ev_done, cudaEventDisableTiming));
  CU_OK(cudaMalloc(&sc.d_counters, kCntTotal * sizeof(unsigned)));
  CU_OK(cudaMallocHost(&sc.h_counters, kCntTotal * sizeof(unsigned)));
  CU_OK(cudaMallocHost(&sc.h_eager, kEagerHits * (kHitHeader + kMaxDim) * sizeof(float)));
  sc.h_stage_cap = (size_t)8 << 20;
  CU_OK(cudaMallocHost(&sc.h_stage, sc.h_stage_cap));
  sc.ready = true;
  return true;
}

void run_delete(Run *r);
void slot_free(Scratch &sc) {
  run_delete(sc.run);
  sc.run = nullptr;
  dev_free(sc.d_counters);
  host_free(sc.h_counters); host_free(sc.h_eager); host_free(sc.h_stage);
  sc.d_dims.release(); sc.d_packed.release(); sc.d_unpack.release(); sc.d_frames.release(); sc.d_hq.release();
  sc.d_surv.release(); sc.d_shape0.release(); sc.d_surv_leaves.release(); sc.d_hits.release();
  sc.d_stage_list[0].release(); sc.d_stage_list[1].release();
  for (auto &e : sc.ev) { if (e) cudaEventDestroy(e); e = nullptr; }
  for (auto &e : sc.ev_copy) { if (e) cudaEventDestroy(e); e = nullptr; }
  if (sc.ev_done) cudaEventDestroy(sc.ev_done);
  sc.ev_done = nullptr;
  sc.ready = false; sc.busy = false;
}

bool ctx_init_impl(Context *c) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_err("no CUDA device visible: libjda_b200 has no CPU path");
    return false;
  }
  if (c->device < 0) {
    int cur = 0;
    cudaGetDevice(&cur);
    c->device = cur;
  }
  CU_OK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CU_OK(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major < 10) {
    set_err("device %d is sm_%d%d; this library is built for sm_100a only", c->device, prop.major, prop.minor);
    return false;
  }
  c->sm_count = prop.multiProcessorCount;
  CU_OK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  CU_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CU_OK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  c->sc = &c->slot[0];
  if (!slot_init(c->slot[0])) return false;
  const HostModel &m = c->m;
  CU_OK(cudaMalloc(&c->d_nodes, m.nodes.size() * sizeof(NodeRec)));
  CU_OK(cudaMalloc(&c->d_leaf, m.leaf.size() * 4));
  CU_OK(cudaMalloc(&c->d_cart, m.cart.size() * 4));
  CU_OK(cudaMalloc(&c->d_w, m.w.size() * 4));
  CU_OK(cudaMalloc(&c->d_mean, m.mean_shape.size() * 4));
  CU_OK(cudaMalloc(&c->d_norms, kMaxNorm * sizeof(Stage0Norm)));
  CU_OK(cudaMemcpy(c->d_nodes, m.nodes.data(), m.nodes.size() * sizeof(NodeRec), cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(c->d_leaf, m.leaf.data(), m.leaf.size() * 4, cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(c->d_cart, m.cart.data(), m.cart.size() * 4, cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(c->d_w, m.w.data(), m.w.size() * 4, cudaMemcpyHostToDevice));
  {
    const int D = m.D(), Dp = k3r_dpad(D);
    const size_t rows = m.w.size() / D;
    std::vector<float> wp(rows * Dp, 0.f);
    for (size_t r = 0; r < rows; r++) memcpy(&wp[r * Dp], &m.w[r * D], (size_t)D * 4);
    CU_OK(cudaMalloc(&c->d_wp, wp.size() * 4));
    CU_OK(cudaMemcpy(c->d_wp, wp.data(), wp.size() * 4, cudaMemcpyHostToDevice));
  }
  CU_OK(cudaMemcpy(c->d_mean, m.mean_shape.data(), m.mean_shape.size() * 4, cudaMemcpyHostToDevice));
  {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      c->encode = (EncodeTiledFn)fn;
  }
  // Function attributes belong to the device context, not to a handle: every handle asks for the largest amount any
  // model may need, so that a handle with a small model cannot lower the limit under a handle with a large one
  // (r2: a K = 70 handle created between two calls of a K = 540 handle made the latter's scan launch fail).
  constexpr int kMaxK = 4096;  // load_model's bound
  const size_t s2 = k2_smem_bytes(kMaxStage0TableBytes);
  CU_OK(cudaFuncSetAttribute(k2_scan<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  CU_OK(cudaFuncSetAttribute(k2_scan<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  CU_OK(cudaFuncSetAttribute(k2_scan<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  CU_OK(cudaFuncSetAttribute(k2_scan<4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  CU_OK(cudaFuncSetAttribute(k2_scan<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  CU_OK(cudaFuncSetAttribute(k2_scan<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  CU_OK(cudaFuncSetAttribute(k2_scan<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2));
  CU_OK(cudaFuncSetAttribute(k3_stage0<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)k3s_smem_bytes(kMaxK, kMaxDim)));
  CU_OK(cudaFuncSetAttribute(k3_stage0<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)k3s_smem_bytes(kMaxK, kMaxDim)));
  const size_t s3 = k3_smem_bytes(kMaxK);
  CU_OK(cudaFuncSetAttribute(k3_cascade<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3));
  CU_OK(cudaFuncSetAttribute(k3_cascade<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3));
  CU_OK(cudaFuncSetAttribute(k3_cascade<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3));
  CU_OK(cudaFuncSetAttribute(k3_cascade<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s3));
  CU_OK(cudaFuncSetAttribute(k3_regress<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3r_smem_bytes(kMaxK, 64)));
  CU_OK(cudaFuncSetAttribute(k3_regress<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3r_smem_bytes(kMaxK, kMaxDim)));
  CU_OK(cudaFuncSetAttribute(k3_walk<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemOptIn));
  CU_OK(cudaFuncSetAttribute(k3_walk<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemOptIn));
  CU_OK(cudaFuncSetAttribute(k3_walk<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemOptIn));
  CU_OK(cudaFuncSetAttribute(k3_walk<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemOptIn));
  // tuning knobs (not behaviour; read from the environment when the handle was created): the phase schedule of k2_scan
  c->sched.clear();
  if (!c->tune.sched.empty()) {
    const char *p = c->tune.sched.c_str();
    while (*p) {
      int v = (int)strtol(p, (char **)&p, 10);
      if (v > 0 && v < m.K && (c->sched.empty() || v > c->sched.back()) && c->sched.size() < K2_MAX_SCHED - 1)
        c->sched.push_back((short)v);
      while (*p == ',' || *p == ' ') p++;
    }
  } else {
    // (round 2, with the straggler threshold at 32: a first phase of 8 carts instead of 4 + 4 is -0.6 %, a finer middle
    // -0.2 %, coarser late phases +1 %: profiles/r3k_sched_ab.txt)
    const int def[] = {8, 16, 24, 32, 40, 48, 64, 80, 96, 128, 160, 192, 256, 320, 384, 448};
    for (int v : def)
      if (v < m.K) c->sched.push_back((short)v);
  }
  c->sched.push_back((short)m.K);
  return true;
}

// Device state is all-or-nothing: a failure half way through releases what was created, so a retry starts clean
// and nothing leaks (`inited` is set only after every step succeeded).
bool ctx_init(Context *c) {
  if (c->inited) {
    CU_OK(cudaSetDevice(c->device));
    return true;
  }
  if (!ctx_init_impl(c)) {
    const std::string keep = g_err;
    ctx_release_device(c);
    g_err = keep;
    return false;
  }
  c->inited = true;
  return true;
}

void release_model64(Context *c) {
  dev_free(c->d_nodes64); dev_free(c->d_leaf64); dev_free(c->d_cart64); dev_free(c->d_w64); dev_free(c->d_mean64);
  dev_free(c->d_norms64);
  c->model64_ready = false;
}

// frees every device / pinned resource of the handle (safe on a partially initialised one)
void ctx_release_device(Context *c) {
  if (c->device >= 0) cudaSetDevice(c->device);
  dev_free(c->d_nodes); dev_free(c->d_leaf); dev_free(c->d_cart); dev_free(c->d_w); dev_free(c->d_wp); dev_free(c->d_mean);
  dev_free(c->d_norms);
  release_model64(c);
  c->d_tables64.release(); c->d_hits64.release(); c->d_trace_s64.release(); c->d_trace_n64.release();
  c->d_tables.release(); c->d_trace_leaf.release(); c->d_trace_n.release(); c->d_trace_s.release();
  for (auto &sl : c->slot) slot_free(sl);
  c->sc = &c->slot[0];
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  c->copy_stream = nullptr; c->d2h_stream = nullptr; c->own_stream = nullptr;
  c->geo.valid = false; c->geo64.valid = false;
  cudaGetLastError();
}

void ctx_free(Context *c) {
  ctx_release_device(c);
  delete c;
}

// Expected shared-memory wavefronts of one pixel read (1 = conflict free) for packets of 32 survivors
// drawn in row-major order from a tw x th tile at a few survival densities.  Bank = (byte address / 4)
// mod 32; lanes on the same word broadcast, lanes on different words of one bank serialise.  The tile
// pitch decides how the rows of a tile fold onto the 32 banks (tests/design_sims/sim_banks.py has the long version).
double estimate_wavefronts(int win, int step, int tw, int th, int pitch) {
  uint32_t rng = 12345u;
  auto next = [&]() { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
  const double dens[3] = {0.7, 0.4, 0.2};
  double tot = 0;
  int cnt = 0;
  for (double rho : dens) {
    for (int trial = 0; trial < 12; trial++) {
      int base[32], nb = 0;
      int start = (int)(next() % (uint32_t)(tw * th));
      for (int i = 0; i < tw * th && nb < 32; i++) {
        const int w = (start + i) % (tw * th);
        if ((next() & 0xffff) < (uint32_t)(rho * 65536)) base[nb++] = (w / tw) * step * pitch + (w % tw) * step;
      }
      if (nb < 8) continue;
      for (int o = 0; o < 4; o++) {
        const int off = (int)(next() % (uint32_t)win) * pitch + (int)(next() % (uint32_t)win);
        int worst = 0;
        for (int bnk = 0; bnk < 32; bnk++) {
          int words[32], nw = 0;
          for (int l = 0; l < nb; l++) {
            const int wd = (base[l] + off) >> 2;
            if ((wd & 31) != bnk) continue;
            bool seen = false;
            for (int q = 0; q < nw; q++) seen |= (words[q] == wd);
            if (!seen) words[nw++] = wd;
          }
          worst = std::max(worst, nw);
        }
        tot += worst;
        cnt++;
      }
    }
  }
  return cnt ? tot / cnt : 1.0;
}


// Per-level tile shapes.  A level runs from private shared-memory tiles when a tile of at least
// 64 windows (with its 16-byte-granular pixel box) fits the per-warp scratch; otherwise its windows
// read pixels from global memory in "virtual" tiles of 32 x 16 windows.  Among the shapes that fit,
// the one with the lowest modelled cost wins: shared-memory wavefronts per cart (4 table reads + 6
// pixel reads x conflict factor) x a tail factor that favours tiles with more windows.

// Extra tile pitch per window size.  The pitch decides how a tile's window rows fold onto the 32 banks, i.e. how
// often the pixel reads of a compacted packet (windows from several rows) collide; the exact model in
// tests/design_sims/sim_exact_banks.py ranks the candidates.  JDA_B200_PITCH_EXTRA="24:16,30:32" overrides (A/B).
int pitch_extra(const Tuning &tn, int win) {
  const char *p = tn.pitch_extra.c_str();
  while (*p) {
    char *q;
    const long w = strtol(p, &q, 10);
    if (*q != ':') break;
    const long e = strtol(q + 1, &q, 10);
    if (w == win) return (int)(e / 16 * 16);
    p = (*q == ',') ? q + 1 : q;
    if (*q != ',') break;
  }
  return 0;
}

// best shared-memory tile for a budget of `tile_bytes`; returns its window count (0 = none fits)
int plan_tile(const Tuning &tn, LevelInfo &L, int tile_bytes, int min_tl = 3, int max_windows = K2_LIST_CAP) {
  double best_cost = 1e30;
  int best_windows = 0;
  tile_bytes = std::min(tile_bytes, 65536);  // one TMA box (<= 256 x 256)
  for (int tl = 5; tl >= min_tl; tl--) {
    const int tw = 1 << tl;
    // TMA wants the box to start on a 16-byte boundary in x: tiles whose x origin (tx * tw * step) is not
    // a multiple of 16 start their box at the aligned address below it and carry up to 15 spare bytes
    const int slack = ((tw * L.step) % 16 == 0) ? 0 : 15;
    const int bw0 = ((((tw - 1) * L.step + L.win + slack) + 15) & ~15) + pitch_extra(tn, L.win);
    for (int bw = bw0; bw <= std::min(256, bw0 + (tn.tune_pitch ? 80 : 0)); bw += 16) {
      const int bh_max = std::min(256, tile_bytes / bw);
      if (bh_max < L.win) continue;
      int th = (bh_max - L.win) / L.step + 1;
      th = std::min(th, std::max(1, max_windows / tw));
      th = std::min(th, std::max(1, L.ny));
      const int bh = (th - 1) * L.step + L.win;
      if ((long long)L.win * bw + L.win >= 65536) continue;  // u16 tile offsets
      const int windows = std::min(tw, L.nx) * th;
      if (windows < (min_tl < 3 ? std::min(tn.min_tile_windows, 6) : tn.min_tile_windows)) continue;
      // default: most windows per tile, wider tiles on ties; tuned: modelled wavefronts x tail factor
      const double cost = tn.tune_pitch ? (4.0 + 6.0 * estimate_wavefronts(L.win, L.step, tw, th, bw)) * (1.0 + 40.0 / windows)
                                       : -(double)windows;
      if (cost < best_cost) {
        best_cost = cost;
        best_windows = windows;
        L.tw_log2 = tl; L.th = th; L.box_w = bw; L.box_h = bh;
      }
    }
  }
  return best_windows;
}

// Per-level tile shapes.  A level runs from shared-memory tiles when a tile of at least 128 windows
// (with its 16-byte-granular pixel box) fits a warp's 8 KB buffer -- or, for coarser levels, the
// buffers of 2 or 4 neighbouring warps (only every 2nd / 4th warp then works on that level).  Levels
// that fit neither read pixels from global memory in "virtual" tiles of 32 x 16 windows.
void plan_level(const Tuning &tn, LevelInfo &L, int plan) {
  const bool latency = plan == 1;
  // Two plans.  Throughput (many frames in flight): coarse levels pool at most 4 warps' buffers, what is
  // left reads global memory in 512-window virtual tiles -- measured fastest on 128+ frame batches.
  // Latency (a handful of frames): the coarsest levels pool the whole block's buffers and go straight to
  // cart-parallel straggler mode, global-memory tiles are small: a single VGA frame's scan drops from
  // 0.58 ms to 0.29 ms because no warp is left with a 0.5 ms dependent chain.
  const int max_span = tn.max_span > 0 ? tn.max_span : (latency ? K2_WARPS : 4);
  L.use_smem = 0;
  L.span = 1;
  // latency plan: 256-window tiles on the fine levels -- twice the tiles, shorter dependent chains per warp
  if (plan_tile(tn, L, K2_TILE_BYTES, 3, latency ? tn.latency_tile_windows : K2_LIST_CAP) > 0) {
    L.use_smem = 1;  // fits a single warp's buffer: every warp works
  } else {
    // pool buffers: more bytes per tile (more windows) against fewer independent groups.  With every
    // warp of the block on one tile (span == K2_WARPS) even the coarsest windows fit; their few windows
    // per warp go straight to cart-parallel straggler mode, which is also what keeps single-frame
    // latency short (a 512-window tile read from global memory is a ~0.5 ms dependent chain).
    double best = 0;
    LevelInfo pick = L;
    const int spans[3] = {2, 4, K2_WARPS};
    for (int span : spans) {
      if (span > max_span || K2_WARPS % span) continue;
      LevelInfo t = L;
      const int windows = plan_tile(tn, t, span * K2_TILE_BYTES, latency ? 1 : 3);
      const double score = windows * std::sqrt((double)K2_WARPS / span);
      if (windows > 0 && score > best) { best = score; pick = t; pick.use_smem = 1; pick.span = span; }
    }
    L = pick;
  }
  if (!L.use_smem) {  // global-memory virtual tiles
    // (medium batches, plan 2: 128-window virtual tiles -- a 512-window tile read from global memory is a ~1 ms chain of
    // dependent loads, the floor of every scan under ~100 frames; at 512 frames the short tiles cost 0.6 %)
    L.tw_log2 = 5; L.th = (latency || plan == 2) ? 4 : K2_LIST_CAP / 32; L.box_w = 0; L.box_h = 0;
    if (tn.global_tile_rows > 0) L.th = tn.global_tile_rows;
  }
  const int tw = 1 << L.tw_log2;
  L.ntx = (L.nx + tw - 1) / tw;
  L.nty = (L.ny + L.th - 1) / L.th;
}

bool ensure_geometry(Context *c, int w, int h, float scale, int min_size, int max_size, int plan) {
  Geometry &g = c->geo;
  if (g.valid && g.w == w && g.h == h && g.scale == scale && g.min_size == min_size && g.max_size == max_size &&
      g.plan == plan)
    return true;
  g.valid = false;
  g.w = w; g.h = h; g.scale = scale; g.min_size = min_size; g.max_size = max_size; g.plan = plan;
  int wins[kMaxLevels + 1];
  int n = (w < 24 || h < 24) ? 0 : enumerate_levels(w, h, scale, min_size, max_size, wins, kMaxLevels + 1);
  if (n > kMaxLevels) {
    set_err("more than %d pyramid levels (scale too close to 1)", kMaxLevels);
    return false;
  }
  g.n_levels = n;
  g.table_bytes = (c->m.K * kCartBytes + 127) & ~127;
  long long base = 0;
  for (int i = 0; i < n; i++) {
    LevelInfo &L = g.lv[i];
    memset(&L, 0, sizeof L);
    L.win = wins[i];
    L.step = level_step(L.win);
    if (L.win >= 2048) { set_err("windows of %d px exceed the 2047 px limit", L.win); return false; }
    L.nx = (w - L.win) / L.step + 1;
    L.ny = (h - L.win) / L.step + 1;
    if (L.nx > 8191 || L.ny > 8191) { set_err("frame too large for 13-bit window indices"); return false; }
    plan_level(c->tune, L, plan);
    L.table_off = i * g.table_bytes;
    L.win_base = base;
    base += (long long)L.nx * L.ny;
  }
  g.windows_per_frame = base;
  if (n > 0 && c->m.stage0_lut_ok) {
    std::vector<uint8_t> tab((size_t)n * g.table_bytes, 0);
    Stage0Norm norms[kMaxNorm];
    memset(norms, 0, sizeof norms);
    for (int i = 0; i < n; i++)
      build_stage0_table(c->m, g.lv[i].win, g.lv[i].use_smem ? g.lv[i].box_w : 0,
                         tab.data() + (size_t)i * g.table_bytes, norms);
    if (!c->d_tables.ensure(tab.size())) return false;
    CU_OK(cudaMemcpyAsync(c->d_tables.p, tab.data(), tab.size(), cudaMemcpyHostToDevice, c->stream()));
    CU_OK(cudaMemcpyAsync(c->d_norms, norms, sizeof norms, cudaMemcpyHostToDevice, c->stream()));
    CU_OK(cudaStreamSynchronize(c->stream()));
  }
  g.valid = true;
  return true;
}

// bytes of a host frame from its first to its last pixel (rows `pitch` apart)
inline size_t frame_bytes(const jdaB200Frame &f) {
  if (f.width <= 0 || f.height <= 0) return 0;
  return (size_t)(f.pitch > 0 ? f.pitch : f.width) * (f.height - 1) + f.width;
}

struct TraceOut {
  int *n = nullptr;
  float *s = nullptr;
  uint8_t *leaf = nullptr;
  long long w0 = 0, w1 = 0;
};

// ---------------------------------------------------------------------------------- one batch on the device
//
// run_device = stage_frames -> make_planes -> [launch_scan per chunk -> launch_cascade] -> collect,
// retried with larger queues when the device counters say a queue overflowed.

// State of one run_device call.
struct Run {
  Context *c;
  const jdaB200Batch *b;
  const jdaB200Frame *mixed;  // non-NULL: frames of different sizes, each in its own slot of a b->width x b->height canvas
  const TraceOut *trace;
  const Geometry *geo;        // the scan geometry of this call and its stage-0 tables
  const uint8_t *tables;
  const Stage0Norm *norms;
  bool timing, tracing;
  bool latency_plan;   // <= kLatencyFrames frames: latency tile plan, no cohort-staged stage 0
  bool use_scan;       // stage 0 from the LUT scan (k2_scan); false: every window through k3_cascade
  bool staged0;        // k3_stage0 computes the survivors' stage-0 shapes (batches)
  cudaStream_t s;
  const uint8_t *d_frames;
  int pitch;
  size_t fstride;
  int nchunks;         // scan launches (the host copy is split the same way)
  int chunks_copied;   // mixed-size batches: chunks whose host -> device copies have been issued
  bool host_chunks;
  int D, rec_words, t_run, k_extra, leaf_stride, leaf_pad;
  int scan_K;          // carts the scan walks: K, or k_extra when the truncated cascade ends inside stage 0
  long long total_windows;
  float r;             // 1.f / sqrtf(2.f) as the reference computes it (c/jda.c:341)
  int hw, hh, qw, qh;
  size_t hq_stride;
  size_t eager;        // hit records that travel with the counters (collect_start)
  bool async;          // submitted batch: its copies must not queue behind the batch before it
  bool mixed_staged;   // mixed-size batch from pageable host memory: chunks are repacked into the pinned staging area first
  bool one_chunk;      // submitted while the batch before it is still running: the whole copy hides under that batch, so
                       // the scan is ONE launch (every chunk's launch ends in a tail of half-idle SMs)
  size_t cap_surv, cap_hit;  // queue capacities the kernels of this attempt were launched with
  bool done;           // nothing to run (no frames, or no pyramid level fits)
};

void run_delete(Run *r) { delete r; }

// mixed-size batch: host -> canvas slots for the frames of chunk `ch` (each chunk once per call).  Every frame is
// copied as one contiguous blob into the packed staging area (frames that are neighbours in host memory share a
// copy), then k0_unpack spreads the chunk over its canvas slots; all on the copy stream.
bool copy_mixed_chunk(Run &R, int ch) {
  Context *c = R.c;
  const jdaB200Batch &b = *R.b;
  if (ch >= R.nchunks || ch < R.chunks_copied) return true;
  const int f0 = chunk_begin(b.n_frames, ch, R.nchunks, c->tune.even_chunks), f1 = chunk_begin(b.n_frames, ch + 1, R.nchunks, c->tune.even_chunks);
  const UnpackFrame *tab = c->sc->h_unpack.data();
  // runs of frames that are contiguous in host memory (and therefore in the staging area): {first frame, bytes}
  std::vector<std::pair<int, size_t>> runs;
  int run0 = f0;
  for (int f = f0; f <= f1; f++) {
    const bool extend = f < f1 && f > run0 && tab[f].src_off == tab[f - 1].src_off + frame_bytes(R.mixed[f - 1]) &&
                        R.mixed[f].data == R.mixed[f - 1].data + frame_bytes(R.mixed[f - 1]);
    if (extend) continue;
    if (f > run0) {
      const size_t bytes = tab[f - 1].src_off + frame_bytes(R.mixed[f - 1]) - tab[run0].src_off;
      if (bytes > 0) runs.push_back({run0, bytes});
    }
    run0 = f;
  }
  if (R.mixed_staged && !runs.empty()) {
    // pageable frames (separately malloc'd images, numpy arrays): a few host threads copy the runs into the pinned
    // staging area at their packed offsets, then the chunk goes across in ONE asynchronous copy -- the driver would
    // stage every frame through its bounce buffer on the calling thread
    uint8_t *hs = c->sc->h_stage;
    const jdaB200Frame *mixed = R.mixed;
    size_t total = 0;
    for (const auto &r : runs) total += r.second;
    auto work = [&runs, tab, hs, mixed](size_t i0, size_t i1) {
      for (size_t i = i0; i < i1; i++) memcpy(hs + tab[runs[i].first].src_off, mixed[runs[i].first].data, runs[i].second);
    };
    const unsigned hc = std::thread::hardware_concurrency();
    const size_t nt = std::min<size_t>(std::min<unsigned>(8u, hc ? hc / 2 : 1u), runs.size());
    if (total < ((size_t)2 << 20) || nt < 2) {
      work(0, runs.size());
    } else {
      std::vector<std::thread> th;
      for (size_t i = 1; i < nt; i++) th.emplace_back(work, runs.size() * i / nt, runs.size() * (i + 1) / nt);
      work(0, runs.size() / nt);
      for (auto &t : th) t.join();
    }
    const size_t o0 = tab[runs.front().first].src_off, o1 = tab[runs.back().first].src_off + runs.back().second;
    CU_OK(cudaMemcpyAsync(c->sc->d_packed.p + o0, hs + o0, o1 - o0, cudaMemcpyHostToDevice, c->copy_stream));
  } else {
    for (const auto &r : runs)
      CU_OK(cudaMemcpyAsync(c->sc->d_packed.p + tab[r.first].src_off, R.mixed[r.first].data, r.second, cudaMemcpyHostToDevice,
                            c->copy_stream));
  }
  if (f1 > f0) {
    dim3 grid((b.height + 7) / 8, f1 - f0);
    k0_unpack<<<grid, 128, 0, c->copy_stream>>>(c->sc->d_packed.p, c->sc->d_unpack.p, f0, c->sc->d_frames.p, R.fstride, R.pitch);
    CU_OK(cudaGetLastError());
  }
  CU_OK(cudaEventRecord(c->sc->ev_copy[ch], c->copy_stream));
  R.chunks_copied = ch + 1;
  return true;
}

// Frames [f0, f1) from the caller's (pageable) memory into the 16-byte-pitched pinned staging area, rows split over a
// few host threads: one memcpy stream runs at 6-10 GB/s, a 157 MB batch should not take longer than its H2D copy.
void repack_frames(uint8_t *dst, int dst_pitch, size_t dst_fstride, const unsigned char *src, int src_pitch, size_t src_fstride,
                   int width, int height, int f0, int f1) {
  const long long rows = (long long)(f1 - f0) * height;
  const size_t bytes = (size_t)rows * width;
  auto work = [=](long long r0, long long r1) {
    if (dst_pitch == src_pitch && dst_fstride == src_fstride && dst_fstride == (size_t)dst_pitch * height) {
      memcpy(dst + f0 * dst_fstride + r0 * dst_pitch, src + f0 * src_fstride + r0 * src_pitch, (size_t)(r1 - r0) * dst_pitch);
      return;
    }
    for (long long r = r0; r < r1; r++) {
      const long long f = f0 + r / height, y = r % height;
      memcpy(dst + f * dst_fstride + y * dst_pitch, src + f * src_fstride + y * src_pitch, width);
    }
  };
  unsigned hc = std::thread::hardware_concurrency();
  int nt = (int)std::min<unsigned>(8u, hc ? hc / 2 : 1u);
  if (bytes < ((size_t)4 << 20) || nt < 2) { work(0, rows); return; }
  std::vector<std::thread> th;
  for (int i = 1; i < nt; i++) th.emplace_back(work, rows * i / nt, rows * (i + 1) / nt);
  work(0, rows / nt);
  for (auto &t : th) t.join();
}

// Frames to HBM.  Device input is used in place; host input goes into a 16-byte-pitched store: large batches
// in kMaxChunks pieces on the copy stream (the scan of chunk i then overlaps the copy of chunk i+1), a small
// pageable input repacked through pinned staging (the driver's pageable path costs more than a one-frame detect).
bool stage_frames(Run &R, const unsigned char *frames) {
  Context *c = R.c;
  const jdaB200Batch &b = *R.b;
  const HostModel &m = c->m;
  R.nchunks = 1;
  R.host_chunks = false;
  if (b.flags & JDA_B200_DEVICE_INPUT) {
    R.d_frames = frames; R.pitch = b.pitch; R.fstride = b.frame_stride;
    return true;
  }
  R.pitch = (b.width + 15) & ~15;
  R.fstride = (size_t)R.pitch * b.height;
  if (!c->sc->d_frames.ensure(R.fstride * b.n_frames + 256)) return false;
  if (b.n_frames >= 128 && !m.any_scaled && R.use_scan && !R.tracing && !c->tune.no_chunks && !(R.one_chunk && !R.mixed)) {
    R.nchunks = kMaxChunks;
    R.host_chunks = true;
  }
  if (R.async) {
    // this scratch set's previous batch was collected (jdaB200Collect waited for it), so its frame store is free: the
    // copies start at once, next to the kernels of the batch submitted before
    CU_OK(cudaEventRecord(c->sc->ev_copy[kMaxChunks], c->copy_stream));
  } else {
    CU_OK(cudaEventRecord(c->sc->ev_copy[kMaxChunks], R.s));
    CU_OK(cudaStreamWaitEvent(c->copy_stream, c->sc->ev_copy[kMaxChunks], 0));  // scratch of the previous call is free
  }
  if (R.mixed) {
    // every frame goes to the top-left corner of its canvas slot; what lies outside a frame inside its slot is
    // never sampled (windows are enumerated from the frame's own width and height), so it is left as it is.
    // Only the first chunk's copies are issued here: launch_scan issues chunk i+1 right after the scan of
    // chunk i, so the host-side cost of thousands of small copy calls hides behind the running scan too.
    if (!c->sc->d_dims.ensure(b.n_frames)) return false;
    std::vector<int2> dims(b.n_frames);
    for (int f = 0; f < b.n_frames; f++) dims[f] = make_int2(std::max(R.mixed[f].width, 0), std::max(R.mixed[f].height, 0));
    CU_OK(cudaMemcpyAsync(c->sc->d_dims.p, dims.data(), dims.size() * sizeof(int2), cudaMemcpyHostToDevice, c->copy_stream));
    // staging layout: host neighbours stay neighbours (one copy per run), everything else starts 16-byte aligned
    c->sc->h_unpack.resize(b.n_frames);
    size_t off = 0;
    for (int f = 0; f < b.n_frames; f++) {
      const jdaB200Frame &fr = R.mixed[f];
      const bool adjacent = f > 0 && frame_bytes(R.mixed[f - 1]) > 0 && fr.data == R.mixed[f - 1].data + frame_bytes(R.mixed[f - 1]);
      if (!adjacent) off = (off + 15) & ~(size_t)15;
      c->sc->h_unpack[f] = UnpackFrame{(unsigned long long)off, std::max(fr.width, 0), std::max(fr.height, 0),
                                   fr.pitch > 0 ? fr.pitch : fr.width, 0};
      off += frame_bytes(fr);
    }
    if (!c->sc->d_packed.ensure(off + 16) || !c->sc->d_unpack.ensure(b.n_frames)) return false;
    {  // pageable sources (judged by the first frame that has data) go through the pinned staging area
      R.mixed_staged = false;
      const unsigned char *first = nullptr;
      for (int f = 0; f < b.n_frames && !first; f++)
        if (frame_bytes(R.mixed[f]) > 0) first = R.mixed[f].data;
      cudaPointerAttributes pa;
      const bool pageable = first && (cudaPointerGetAttributes(&pa, first) != cudaSuccess || pa.type == cudaMemoryTypeUnregistered);
      cudaGetLastError();
      if (pageable && off + 16 > c->sc->h_stage_cap) {
        uint8_t *bigger = nullptr;
        const size_t want = off + 16 + (off >> 3);
        if (cudaMallocHost(&bigger, want) == cudaSuccess) {
          host_free(c->sc->h_stage);
          c->sc->h_stage = bigger;
          c->sc->h_stage_cap = want;
        } else {
          cudaGetLastError();
        }
      }
      R.mixed_staged = pageable && off + 16 <= c->sc->h_stage_cap;
    }
    CU_OK(cudaMemcpyAsync(c->sc->d_unpack.p, c->sc->h_unpack.data(), (size_t)b.n_frames * sizeof(UnpackFrame), cudaMemcpyHostToDevice,
                          c->copy_stream));
    R.chunks_copied = 0;
    if (!copy_mixed_chunk(R, 0)) return false;
    if (!R.host_chunks) CU_OK(cudaStreamWaitEvent(R.s, c->sc->ev_copy[0], 0));
    R.d_frames = c->sc->d_frames.p;
    return true;
  }
  // Pageable input (plain malloc / numpy memory -- what a caller of the reference's API passes): the driver would stage
  // it through its own bounce buffers on one thread (~6-8 GB/s: 157 MB of VGA frames cost more than their scan).  It
  // is repacked into the scratch set's pinned staging by a few host threads instead and goes across with the same
  // asynchronous copies as pinned input, chunk by chunk.
  cudaPointerAttributes pa;
  const bool pageable = cudaPointerGetAttributes(&pa, frames) != cudaSuccess || pa.type == cudaMemoryTypeUnregistered;
  cudaGetLastError();
  const size_t total_bytes = R.fstride * b.n_frames;
  if (pageable && total_bytes > c->sc->h_stage_cap) {
    uint8_t *bigger = nullptr;
    if (cudaMallocHost(&bigger, total_bytes + (total_bytes >> 3)) == cudaSuccess) {
      host_free(c->sc->h_stage);
      c->sc->h_stage = bigger;
      c->sc->h_stage_cap = total_bytes + (total_bytes >> 3);
    } else {
      cudaGetLastError();  // no pinned memory to be had: the driver's pageable path below
    }
  }
  const bool staged = pageable && total_bytes <= c->sc->h_stage_cap;
  for (int ch = 0; ch < R.nchunks; ch++) {
    const int f0 = chunk_begin(b.n_frames, ch, R.nchunks, c->tune.even_chunks), f1 = chunk_begin(b.n_frames, ch + 1, R.nchunks, c->tune.even_chunks);
    if (staged) {
      repack_frames(c->sc->h_stage, R.pitch, R.fstride, frames, b.pitch, b.frame_stride, b.width, b.height, f0, f1);
      CU_OK(cudaMemcpyAsync(c->sc->d_frames.p + f0 * R.fstride, c->sc->h_stage + f0 * R.fstride, (size_t)(f1 - f0) * R.fstride,
                            cudaMemcpyHostToDevice, c->copy_stream));
    } else if (b.frame_stride == (size_t)b.pitch * b.height) {
      CU_OK(cudaMemcpy2DAsync(c->sc->d_frames.p + f0 * R.fstride, R.pitch, frames + f0 * b.frame_stride, b.pitch, b.width,
                              (size_t)b.height * (f1 - f0), cudaMemcpyHostToDevice, c->copy_stream));
    } else {
      for (int f = f0; f < f1; f++)
        CU_OK(cudaMemcpy2DAsync(c->sc->d_frames.p + f * R.fstride, R.pitch, frames + f * b.frame_stride, b.pitch, b.width,
                                b.height, cudaMemcpyHostToDevice, c->copy_stream));
    }
    CU_OK(cudaEventRecord(c->sc->ev_copy[ch], c->copy_stream));
  }
  if (!R.host_chunks) CU_OK(cudaStreamWaitEvent(R.s, c->sc->ev_copy[0], 0));
  R.d_frames = c->sc->d_frames.p;
  return true;
}

// h / q planes (c/jda.c:450-457), only when some node samples them
bool make_planes(Run &R) {
  Context *c = R.c;
  const jdaB200Batch &b = *R.b;
  R.r = 1.f / sqrtf(2.f);
  R.hw = (int)(b.width * R.r); R.hh = (int)(b.height * R.r); R.qw = b.width / 2; R.qh = b.height / 2;
  R.hq_stride = (size_t)R.hw * R.hh + (size_t)R.qw * R.qh;
  if (!c->m.any_scaled) return true;
  if (!c->sc->d_hq.ensure(R.hq_stride * b.n_frames)) return false;
  const int big = std::max(R.hw * R.hh, R.qw * R.qh);
  for (int f0 = 0; f0 < b.n_frames; f0 += 32768) {  // gridDim.z limit
    dim3 grid((big + 255) / 256, 2, std::min(32768, b.n_frames - f0));
    k1_resize<<<grid, 256, 0, R.s>>>(R.d_frames + (size_t)f0 * R.fstride, R.fstride, R.pitch, b.width, b.height,
                                    c->sc->d_hq.p + (size_t)f0 * R.hq_stride, R.hq_stride, R.hw, R.hh, R.qw, R.qh);
    CU_OK(cudaGetLastError());
  }
  c->sc->last.resize_launches = 1;
  return true;
}

bool prepare_trace(Run &R) {
  if (!R.tracing) return true;
  Context *c = R.c;
  if (!c->d_trace_n.ensure(R.total_windows) || !c->d_trace_s.ensure(R.total_windows)) return false;
  CU_OK(cudaMemsetAsync(c->d_trace_n.p, 0, R.total_windows * 4, R.s));
  CU_OK(cudaMemsetAsync(c->d_trace_s.p, 0, R.total_windows * 4, R.s));
  if (R.trace->leaf && R.trace->w1 > R.trace->w0) {
    const size_t nb = (size_t)(R.trace->w1 - R.trace->w0) * R.leaf_stride;
    if (!c->d_trace_leaf.ensure(nb)) return false;
    CU_OK(cudaMemsetAsync(c->d_trace_leaf.p, 255, nb, R.s));
  }
  return true;
}

// k2_scan over every chunk of frames; all chunks feed one survivor queue.
bool launch_scan(Run &R) {
  Context *c = R.c;
  const jdaB200Batch &b = *R.b;
  const Geometry &g = *R.geo;
  jdaB200Stats &st = c->sc->last;
  ScanParams P;
  memset(&P, 0, sizeof P);
  for (int i = 0; i < g.n_levels; i++) P.lv[i] = g.lv[i];
  P.frame_stride = R.fstride; P.pitch = R.pitch; P.W = b.width; P.H = b.height;
  P.n_levels = g.n_levels; P.K = R.scan_K; P.table_bytes = g.table_bytes;
  P.tables = R.tables; P.norms = R.norms;
  P.windows_per_frame = g.windows_per_frame;
  P.n_sched = 0;
  for (size_t i = 0; i < c->sched.size() && c->sched[i] < R.scan_K; i++) P.sched[P.n_sched++] = c->sched[i];
  P.sched[P.n_sched++] = (short)R.scan_K;  // the last phase ends at the last cart the scan walks
  P.surv = c->sc->d_surv.p; P.surv_count = c->sc->d_counters + kCntSurv; P.surv_cap = (unsigned)c->surv_cap;
  P.surv_leaves = R.t_run > 0 ? c->sc->d_surv_leaves.p : nullptr;  // the leaves feed the stage-0 regression only
  P.leaf_pad = R.leaf_pad;
  P.frame_dims = R.mixed ? c->sc->d_dims.p : nullptr;
  // TMA needs 16-byte aligned base and strides
  const bool tma_ok = c->encode && !(b.flags & JDA_B200_NO_TMA) && ((uintptr_t)R.d_frames % 16 == 0) &&
                      R.pitch % 16 == 0 && R.fstride % 16 == 0;
  int n_smem = 0;
  for (int i = 0; i < g.n_levels; i++) n_smem += g.lv[i].use_smem;
  st.levels_smem = n_smem;
  {  // share of the scan work per level, in processing order (coarse -> fine)
    double w[kMaxLevels], tot = 0;
    for (int i = 0; i < g.n_levels; i++) {
      const LevelInfo &L = g.lv[g.n_levels - 1 - i];
      // cost per window grows with the window size (coarse windows survive deeper and their tiles hold fewer windows
      // per warp): measured per level in profiles/r1m_level_probe.txt, ~ (win / 24)^1.6, global-memory levels x1.6
      const double e = c->tune.level_weight_exp;
      w[i] = (double)L.nx * L.ny * (e > 0 ? std::pow(L.win / 24.0, e) * (L.use_smem ? 1.0 : 1.6)
                                          : (L.use_smem ? (L.span == 1 ? 1.0 : 1.4) : 2.2));
      tot += w[i];
    }
    double acc = 0;
    for (int i = 0; i < g.n_levels; i++) { acc += w[i]; P.level_cum[i] = (float)(acc / tot); }
  }
  P.use_tma = tma_ok ? 1 : 0;
  P.stragglers = c->tune.stragglers;
  if (R.tracing) {
    P.trace_n = c->d_trace_n.p; P.trace_s = c->d_trace_s.p;
    P.trace_leaf = (R.trace->leaf && R.trace->w1 > R.trace->w0) ? c->d_trace_leaf.p : nullptr;
    P.leaf_w0 = R.trace->w0; P.leaf_w1 = R.trace->w1; P.leaf_stride = R.leaf_stride;
  }
  const size_t smem = k2_smem_bytes(g.table_bytes);
  const int grid = c->sm_count;
  for (int ch = 0; ch < R.nchunks; ch++) {
    const int f0 = chunk_begin(b.n_frames, ch, R.nchunks, c->tune.even_chunks), f1 = chunk_begin(b.n_frames, ch + 1, R.nchunks, c->tune.even_chunks);
    if (f1 <= f0) continue;
    P.frames = R.d_frames + (size_t)f0 * R.fstride;
    P.n_frames = f1 - f0;
    P.frame_base = f0;
    P.tile_counters = c->sc->d_counters + ch * kMaxLevels;
    for (int i = 0; i < g.n_levels && tma_ok; i++) {
      if (!g.lv[i].use_smem) continue;
      cuuint64_t dims[3] = {(cuuint64_t)b.width, (cuuint64_t)b.height, (cuuint64_t)(f1 - f0)};
      cuuint64_t strides[2] = {(cuuint64_t)R.pitch, (cuuint64_t)R.fstride};
      cuuint32_t box[3] = {(cuuint32_t)g.lv[i].box_w, (cuuint32_t)g.lv[i].box_h, 1};
      cuuint32_t es[3] = {1, 1, 1};
      CUresult cr = c->encode(&P.maps[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)P.frames, dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr != CUDA_SUCCESS) {
        set_err("cuTensorMapEncodeTiled failed (%d) for level %d box %dx%d", (int)cr, i, g.lv[i].box_w, g.lv[i].box_h);
        return false;
      }
    }
    if (R.mixed && !copy_mixed_chunk(R, ch)) return false;  // no-op unless an earlier chunk was empty
    if (R.host_chunks) CU_OK(cudaStreamWaitEvent(R.s, c->sc->ev_copy[ch], 0));
    const int nw = c->tune.nw;
    if (R.mixed) {  // per-frame window grids: always the MIXED instantiation (4 windows per lane whatever the A/B knob says)
      k2_scan<4, false, true><<<grid, K2_WARPS * 32, smem, R.s>>>(P);
    } else if (R.tracing) {
      if (nw == 1) k2_scan<1, true><<<grid, K2_WARPS * 32, smem, R.s>>>(P);
      else if (nw == 4) k2_scan<4, true><<<grid, K2_WARPS * 32, smem, R.s>>>(P);
      else k2_scan<2, true><<<grid, K2_WARPS * 32, smem, R.s>>>(P);
    } else if (nw == 1) {
      k2_scan<1, false><<<grid, K2_WARPS * 32, smem, R.s>>>(P);
    } else if (nw == 4) {
      k2_scan<4, false><<<grid, K2_WARPS * 32, smem, R.s>>>(P);
    } else {
      k2_scan<2, false><<<grid, K2_WARPS * 32, smem, R.s>>>(P);
    }
    CU_OK(cudaGetLastError());
    st.scan_launches++;
    if (R.mixed && R.host_chunks && !copy_mixed_chunk(R, ch + 1)) return false;
  }
  return true;
}

// k3_stage0 (batches) + k3_cascade, once, over the whole survivor queue (or over every window in dense mode).
// Running them per chunk on a second stream under the next chunk's scan was tried -- they fit on the SMs next
// to the scan blocks -- and slowed the scan by more than the cascade time it hid (r1: 31.9 vs 30.2 ms / step).
// The regression of stage t for the windows of `list` (NULL: every queue entry; then the shapes start from the mean
// shape), c/jda.c:403-411.  k3_regress; k3_stage0 is the round-1 kernel, kept for A/B (JDA_B200_OLD_REGRESS).
bool launch_regress(Run &R, int t, const uint2 *list, const unsigned *count) {
  Context *c = R.c;
  const HostModel &m = c->m;
  const float *in_shape = t > 0 ? c->sc->d_shape0.p : nullptr;
  if (c->tune.old_regress) {
    Stage0Params S;
    memset(&S, 0, sizeof S);
    S.surv_leaves = c->sc->d_surv_leaves.p;
    S.w0 = c->d_w + (size_t)t * m.K * kLeaves * R.D; S.mean_shape = c->d_mean; S.K = m.K; S.L = m.L;
    S.surv_count = count; S.surv_cap = (unsigned)c->surv_cap;
    S.list = list; S.in_shape = in_shape; S.out_shape = c->sc->d_shape0.p;
    if (R.D <= 64) k3_stage0<1><<<c->sm_count * 4, K3S_WARPS * 32, k3s_smem_bytes(m.K, R.D), R.s>>>(S);
    else k3_stage0<2><<<c->sm_count * 4, K3S_WARPS * 32, k3s_smem_bytes(m.K, R.D), R.s>>>(S);
  } else {
    RegressParams G;
    memset(&G, 0, sizeof G);
    G.leaves = c->sc->d_surv_leaves.p;
    G.wp = c->d_wp + (size_t)t * m.K * kLeaves * k3r_dpad(R.D); G.mean_shape = c->d_mean; G.K = m.K; G.L = m.L;
    G.count = count; G.cap = (unsigned)c->surv_cap;
    G.list = list; G.in_shape = in_shape; G.out_shape = c->sc->d_shape0.p;
    if (R.D <= 64) k3_regress<true><<<c->sm_count * 4, K3R_WARPS * 32, k3r_smem_bytes(m.K, R.D), R.s>>>(G);
    else k3_regress<false><<<c->sm_count * 3, K3R_WARPS * 32, k3r_smem_bytes(m.K, R.D), R.s>>>(G);
  }
  CU_OK(cudaGetLastError());
  c->sc->last.cascade_launches++;
  return true;
}

bool launch_cascade(Run &R) {
  Context *c = R.c;
  const jdaB200Batch &b = *R.b;
  const Geometry &g = *R.geo;
  const HostModel &m = c->m;
  jdaB200Stats &st = c->sc->last;
  if (R.staged0 && !launch_regress(R, 0, nullptr, c->sc->d_counters + kCntSurv)) return false;
  CascadeParams Q;
  memset(&Q, 0, sizeof Q);
  Q.frames = R.d_frames; Q.frame_stride = R.fstride; Q.pitch = R.pitch; Q.W = b.width; Q.H = b.height;
  Q.hq = m.any_scaled ? c->sc->d_hq.p : nullptr; Q.hq_stride = R.hq_stride; Q.hw = R.hw; Q.hh = R.hh; Q.qw = R.qw; Q.qh = R.qh;
  Q.nodes = c->d_nodes; Q.leaf = c->d_leaf; Q.cart = c->d_cart; Q.w = c->d_w; Q.mean_shape = c->d_mean;
  Q.depth = m.depth; Q.nn = m.nn; Q.nl = m.nl;
  Q.T = m.T; Q.K = m.K; Q.L = m.L; Q.t_run = R.t_run; Q.k_extra = R.k_extra; Q.r = R.r;
  Q.n_levels = g.n_levels;
  for (int i = 0; i < g.n_levels; i++) {
    Q.lv_win[i] = g.lv[i].win; Q.lv_step[i] = g.lv[i].step; Q.lv_nx[i] = g.lv[i].nx; Q.lv_ny[i] = g.lv[i].ny;
    Q.lv_base[i] = g.lv[i].win_base;
  }
  Q.windows_per_frame = g.windows_per_frame;
  Q.dense = R.use_scan ? 0 : 1; Q.dense_total = R.total_windows;
  // queue entries resume after stage 0 when its regression is already applied (k3_stage0) or does not exist (the
  // truncated cascade ended inside stage 0: the scan walked all of its k_extra carts)
  Q.t_start = (R.staged0 || (R.use_scan && R.t_run == 0)) ? 1 : 0;
  Q.n_eval0 = R.scan_K;
  Q.surv = c->sc->d_surv.p; Q.surv_count = c->sc->d_counters + kCntSurv; Q.surv_cap = (unsigned)c->surv_cap;
  Q.init_shape = R.staged0 ? c->sc->d_shape0.p : nullptr;
  Q.work_counter = c->sc->d_counters + kCntWork;
  Q.hits = c->sc->d_hits.p; Q.hit_count = c->sc->d_counters + kCntHit; Q.hit_cap = (unsigned)c->hit_cap;
  Q.rec_words = R.rec_words; Q.th = b.th; Q.use_th = (b.flags & JDA_B200_NO_FINAL_TH) ? 0 : 1;
  if (R.tracing) {
    Q.trace_n = c->d_trace_n.p; Q.trace_s = c->d_trace_s.p;
    Q.trace_leaf = (R.trace->leaf && R.trace->w1 > R.trace->w0) ? c->d_trace_leaf.p : nullptr;
    Q.leaf_w0 = R.trace->w0; Q.leaf_w1 = R.trace->w1; Q.leaf_stride = R.leaf_stride;
  }
  // Large batches: stages >= 1 one at a time (kernels_stages.cuh) -- the stage's tables in shared memory, the running
  // scores replayed with lane = survivor.  Small batches / one frame / dense mode / other tree depths: one warp per
  // window through k3_cascade.
  const int walk_warps = k3w_warps(m.K, R.tracing);
  if (R.staged0 && m.depth == kDepth && walk_warps > 0 && !c->tune.no_stage_kernels &&
      R.total_windows >= c->tune.stage_min_windows) {
    if (!c->sc->d_stage_list[0].ensure(c->surv_cap) || !c->sc->d_stage_list[1].ensure(c->surv_cap)) return false;
    WalkParams W;
    memset(&W, 0, sizeof W);
    W.C = Q;
    W.shape = c->sc->d_shape0.p;
    W.leaf_pad = R.leaf_pad;
    unsigned *cnt = c->sc->d_counters + kCntStage, *work = c->sc->d_counters + kCntStageWork;
    const size_t wsmem = ((k3w_table_bytes(m.K) + 15) & ~(size_t)15) + (size_t)walk_warps * k3w_warp_bytes(R.tracing);
    const uint2 *in = nullptr;
    const unsigned *in_cnt = Q.surv_count;
    int n_before = R.scan_K;
    const int t_end = R.t_run + (R.k_extra > 0 ? 1 : 0);
    for (int t = 1; t < t_end; t++) {
      const bool full = t < R.t_run;  // a whole stage, regression after it; else Validate's unfinished stage (cascador.cpp:199-209)
      uint2 *out = c->sc->d_stage_list[t & 1].p;
      W.t = t; W.Kt = full ? m.K : R.k_extra; W.n_eval_before = n_before;
      W.in_list = in; W.in_count = in_cnt; W.out_list = out; W.out_count = cnt + t; W.work = work + t;
      W.leaves = full ? c->sc->d_surv_leaves.p : nullptr;
      const dim3 wb(walk_warps * 32);
      if (R.tracing) {
        if (m.any_scaled) k3_walk<true, true><<<c->sm_count, wb, wsmem, R.s>>>(W);
        else k3_walk<true, false><<<c->sm_count, wb, wsmem, R.s>>>(W);
      } else {
        if (m.any_scaled) k3_walk<false, true><<<c->sm_count, wb, wsmem, R.s>>>(W);
        else k3_walk<false, false><<<c->sm_count, wb, wsmem, R.s>>>(W);
      }
      CU_OK(cudaGetLastError());
      st.cascade_launches++;
      // c/jda.c:403-411 for the windows that passed: shape += sum_k w[t][8k + leaf_k], in place
      if (full && !launch_regress(R, t, out, cnt + t)) return false;
      in = out; in_cnt = cnt + t; n_before += W.Kt;
    }
    W.in_list = in; W.in_count = in_cnt; W.n_eval_before = n_before;
    if (R.tracing) k3_emit<true><<<c->sm_count * 8, 128, 0, R.s>>>(W);
    else k3_emit<false><<<c->sm_count * 8, 128, 0, R.s>>>(W);
    CU_OK(cudaGetLastError());
    st.cascade_launches++;
    return true;
  }
  const int grid = c->sm_count * 16;  // survivors are handed out by an atomic counter: enough blocks to fill every SM
  const size_t smem = k3_smem_bytes(m.K);
  if (m.depth == kDepth) {
    if (R.tracing) k3_cascade<true><<<grid, K3_WARPS * 32, smem, R.s>>>(Q);
    else k3_cascade<false><<<grid, K3_WARPS * 32, smem, R.s>>>(Q);
  } else {  // tree_depth from the header (2..6): node / leaf counts are run-time values
    if (R.tracing) k3_cascade<true, false><<<grid, K3_WARPS * 32, smem, R.s>>>(Q);
    else k3_cascade<false, false><<<grid, K3_WARPS * 32, smem, R.s>>>(Q);
  }
  CU_OK(cudaGetLastError());
  st.cascade_launches++;
  return true;
}

// Counters + the first hit records start their way back to pinned memory right behind the kernels; ev_done marks
// their arrival.  (A submitted batch is collected later: the stream may by then hold the next batch's kernels, so
// nothing in collect_finish may wait for the stream itself.)
bool collect_start(Run &R) {
  Context *c = R.c;
  cudaStream_t s = R.s;
  CU_OK(cudaMemcpyAsync(c->sc->h_counters, c->sc->d_counters, kCntTotal * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
  // the first records ride along with the counters: a call with few hits needs a single round trip
  R.eager = std::min<size_t>(kEagerHits, c->hit_cap);
  CU_OK(cudaMemcpyAsync(c->sc->h_eager, c->sc->d_hits.p, R.eager * R.rec_words * 4, cudaMemcpyDeviceToHost, s));
  CU_OK(cudaEventRecord(c->sc->ev_done, s));
  return true;
}

// Waits for the batch, reads the rest of the hit records and sorts them into scan order.
// overflow = a queue was too small (capacities already grown): run the batch again.
bool collect_finish(Run &R, std::vector<HitRec> &hits, bool &overflow) {
  Context *c = R.c;
  jdaB200Stats &st = c->sc->last;
  overflow = false;
  CU_OK(cudaEventSynchronize(c->sc->ev_done));
  const size_t eager = R.eager;
  const size_t ns = c->sc->h_counters[kCntSurv], nh = c->sc->h_counters[kCntHit];
  if (ns > R.cap_surv || nh > R.cap_hit) {  // (the capacities this batch ran with: another batch may have grown them since)
    if (ns > c->surv_cap) c->surv_cap = ns + ns / 4;
    if (nh > c->hit_cap) c->hit_cap = nh + nh / 4;
    overflow = true;
    return true;
  }
  st.stage0_survivors = R.use_scan ? (long long)ns : 0;
  st.raw_hits = (long long)nh;
  cudaStream_t d = c->d2h_stream;  // the records are complete (ev_done): no need to queue behind the compute stream
  c->sc->h_hits.resize(std::max(nh, (size_t)1) * R.rec_words);
  memcpy(c->sc->h_hits.data(), c->sc->h_eager, std::min(nh, eager) * R.rec_words * 4);
  bool more = false;
  if (nh > eager) {
    CU_OK(cudaMemcpyAsync(c->sc->h_hits.data() + eager * R.rec_words, c->sc->d_hits.p + eager * R.rec_words,
                          (nh - eager) * R.rec_words * 4, cudaMemcpyDeviceToHost, d));
    more = true;
  }
  if (R.tracing) {
    const TraceOut *t = R.trace;
    if (t->n) CU_OK(cudaMemcpyAsync(t->n, c->d_trace_n.p, R.total_windows * 4, cudaMemcpyDeviceToHost, d));
    if (t->s) CU_OK(cudaMemcpyAsync(t->s, c->d_trace_s.p, R.total_windows * 4, cudaMemcpyDeviceToHost, d));
    if (t->leaf && t->w1 > t->w0)
      CU_OK(cudaMemcpyAsync(t->leaf, c->d_trace_leaf.p, (size_t)(t->w1 - t->w0) * R.leaf_stride, cudaMemcpyDeviceToHost, d));
  }
  if (R.timing) CU_OK(cudaEventRecord(c->sc->ev[5], d));
  if (more || R.tracing || R.timing) CU_OK(cudaStreamSynchronize(d));
  hits.resize(nh);
  for (size_t i = 0; i < nh; i++) {
    const float *rec = c->sc->h_hits.data() + i * R.rec_words;
    const int *ri = reinterpret_cast<const int *>(rec);
    hits[i] = HitRec{ri[0], (uint32_t)ri[1], ri[2], ri[3], ri[4], rec[5], rec + kHitHeader};
  }
  // the reference's scan order: level, then y, then x (c/jda.c:332-339) -- the key packs exactly that
  std::sort(hits.begin(), hits.end(), [](const HitRec &a, const HitRec &b2) {
    return a.frame != b2.frame ? a.frame < b2.frame : a.key < b2.key;
  });
  return true;
}

// Runs the device path for one batch; on success `hits` holds the raw hit records sorted into scan
// order (frame, level, y, x).  The record floats stay alive in the context until the next call.
// run_device = run_prepare (geometry, frames to HBM, planes) -> [run_enqueue (kernels + start of the read-back) ->
// run_finish (wait, hit records, scan order)] repeated with larger queues while one overflows.  jdaB200Submit stops
// after run_enqueue, jdaB200Collect picks up at run_finish.
// R.done = nothing to do (no frames / no levels): `hits` stays empty.
bool run_prepare(Context *c, Run &R, const unsigned char *frames, const jdaB200Batch &b, const TraceOut *trace,
                 bool timing, const jdaB200Frame *mixed, int async = 0 /* 1: submitted batch, 2: + the batch before it is still running */) {
  jdaB200Stats &st = c->sc->last;
  memset(&st, 0, sizeof st);
  memset(&R, 0, sizeof R);
  R.done = true;
  R.async = async != 0;
  R.one_chunk = async == 2;
  if (b.n_frames <= 0) return true;
  if (b.width <= 0 || b.height <= 0 ||
      (!mixed && (b.pitch < b.width || b.frame_stride < (size_t)b.pitch * (b.height - 1) + b.width))) {
    set_err("bad batch descriptor");
    return false;
  }
  R.c = c; R.b = &b; R.mixed = mixed; R.trace = trace; R.timing = timing; R.tracing = trace != nullptr;
  const int plan = batch_plan(c->tune, b, mixed);
  R.latency_plan = plan == 1;
  if (!ensure_geometry(c, b.width, b.height, b.scale, b.min_size, b.max_size, plan)) return false;
  const Geometry &g = c->geo;
  R.geo = &c->geo; R.tables = c->d_tables.p; R.norms = c->d_norms;
  st.n_levels = g.n_levels;
  st.windows = g.windows_per_frame * b.n_frames;
  if (mixed) {  // each frame has the windows c/jda.c:320-339 enumerates for its own size
    st.windows = 0;
    for (int f = 0; f < b.n_frames; f++)
      st.windows += count_windows(mixed[f].width, mixed[f].height, b.scale, b.min_size, b.max_size);
  }
  if (g.n_levels == 0) return true;
  R.done = false;
  const HostModel &m = c->m;
  R.s = c->stream();
  R.D = m.D();
  R.rec_words = kHitHeader + R.D;
  R.t_run = (b.t_limit > 0 && b.t_limit < m.T) ? b.t_limit : m.T;
  R.k_extra = 0;
  if (b.k_limit > 0 && b.t_limit >= 0 && b.t_limit < m.T) {  // Validate's unfinished stage (cascador.cpp:199-209)
    R.t_run = b.t_limit;
    R.k_extra = std::min(b.k_limit, m.K);
  }
  R.scan_K = R.t_run == 0 ? R.k_extra : m.K;
  R.leaf_stride = m.T * m.K;
  R.leaf_pad = leaf_bytes(m.K);
  R.total_windows = st.windows;
  R.use_scan = m.stage0_lut_ok && !(b.flags & JDA_B200_NO_STAGE0_SCAN);
  if (mixed && (!R.use_scan || m.any_scaled)) {
    set_err("internal: a mixed-size batch needs the stage-0 scan path");
    return false;
  }
  // a handful of frames: the cohort-staged k3_stage0 is a ~0.1 ms serial pipeline for a few survivors, so the
  // cascade kernel redoes stage 0 itself (same bits); batches take the staged path
  R.staged0 = R.use_scan && !R.latency_plan && R.t_run > 0;
  cudaStream_t s = R.s;

  if (timing) CU_OK(cudaEventRecord(c->sc->ev[0], s));
  if (!stage_frames(R, frames)) return false;
  if (timing) CU_OK(cudaEventRecord(c->sc->ev[1], s));
  if (!make_planes(R) || !prepare_trace(R)) return false;

  if (c->tune.tiny_queues) {  // test hook (read in ctx_init): start with queues that overflow, exercise grow-and-retry
    if (c->surv_cap == 0) { c->surv_cap = 8; c->hit_cap = 2; }
  } else {
    if (c->surv_cap == 0) c->surv_cap = 1 << 16;
    if (c->hit_cap == 0) c->hit_cap = 1 << 14;
    c->surv_cap = std::max(c->surv_cap, (size_t)b.n_frames * 1024);
    c->hit_cap = std::max(c->hit_cap, (size_t)b.n_frames * 128);
  }
  return true;
}

bool run_enqueue(Run &R) {
  Context *c = R.c;
  cudaStream_t s = R.s;
  if (!c->sc->d_surv.ensure(c->surv_cap) || !c->sc->d_hits.ensure(c->hit_cap * R.rec_words)) return false;
  if (R.use_scan && (!c->sc->d_shape0.ensure(c->surv_cap * R.D) || !c->sc->d_surv_leaves.ensure(c->surv_cap * R.leaf_pad))) return false;
  R.cap_surv = c->surv_cap; R.cap_hit = c->hit_cap;
  CU_OK(cudaMemsetAsync(c->sc->d_counters, 0, kCntTotal * sizeof(unsigned), s));
  if (R.timing) CU_OK(cudaEventRecord(c->sc->ev[2], s));
  if (R.use_scan && !launch_scan(R)) return false;
  if (R.timing) CU_OK(cudaEventRecord(c->sc->ev[3], s));
  if (!launch_cascade(R)) return false;
  if (R.timing) CU_OK(cudaEventRecord(c->sc->ev[4], s));
  return collect_start(R);
}

bool run_finish(Run &R, std::vector<HitRec> &hits, bool &overflow) {
  Context *c = R.c;
  jdaB200Stats &st = c->sc->last;
  if (!collect_finish(R, hits, overflow)) return false;
  if (overflow || !R.timing) return true;
  if (R.b->flags & JDA_B200_DEVICE_INPUT) st.ms_h2d = 0.f;
  else cudaEventElapsedTime(&st.ms_h2d, c->sc->ev_copy[kMaxChunks], c->sc->ev_copy[R.host_chunks ? R.nchunks - 1 : 0]);
  cudaEventElapsedTime(&st.ms_resize, c->sc->ev[1], c->sc->ev[2]);
  cudaEventElapsedTime(&st.ms_scan, c->sc->ev[2], c->sc->ev[3]);
  cudaEventElapsedTime(&st.ms_cascade, c->sc->ev[3], c->sc->ev[4]);
  cudaEventElapsedTime(&st.ms_d2h, c->sc->ev[4], c->sc->ev[5]);
  return true;
}

// finish + the rare re-runs with larger queues; on success `hits` holds the raw hit records sorted into scan order
// (frame, level, y, x).  The record floats stay alive in the scratch set until its next batch.
bool run_finish_retrying(Run &R, std::vector<HitRec> &hits) {
  for (int attempt = 0; attempt < 4; attempt++) {
    bool overflow = false;
    if (!run_finish(R, hits, overflow)) return false;
    if (!overflow) return true;
    if (attempt < 3 && !run_enqueue(R)) return false;  // queues were too small: they have been grown, run again
  }
  set_err("survivor / hit queues kept overflowing");
  return false;
}

bool run_device(Context *c, const unsigned char *frames, const jdaB200Batch &b, std::vector<HitRec> &hits,
                const TraceOut *trace, bool timing, const jdaB200Frame *mixed = nullptr) {
  hits.clear();
  if (b.n_frames > 0 && !ctx_init(c)) return false;
  for (const Scratch &sl : c->slot)
    if (sl.busy) { set_err("a submitted batch is waiting for jdaB200Collect: collect it before a synchronous call"); return false; }
  c->sc = &c->slot[0];
  Run R;
  if (!run_prepare(c, R, frames, b, trace, timing, mixed)) return false;
  if (R.done) return true;
  return run_enqueue(R) && run_finish_retrying(R, hits);
}

// =============================================================================== double-precision detector
//
// JoinCascador::Detect with fddb.method = 1 (src/jda/cascador.cpp:310-376, 431-477) on the device:
//   frames -> HBM (stage_frames) -> k2_scan with the f64-derived stage-0 tables as a conservative prefilter
//   (host_model_f64.hpp: stage0_filter_margins) -> k4_cascade_f64 re-evaluates every survivor exactly, in double,
//   from cart 0 -> D2H -> scan order -> multimap NMS + relocation on the host.
// Models / settings the prefilter cannot serve (no finished stage 0, > kMaxNorm normalised carts, unusable
// margins, JDA_B200_NO_STAGE0_SCAN) run every window through k4_cascade_f64 (dense mode).

struct Hit64 {
  int frame;
  uint32_t key;
  int x, y, win, carts;
  double score;
  const double *shape;
};

// host half: the model as doubles + the prefilter margins (no device needed)
bool ensure_model64_host(Context *c) {
  if (c->md.loaded) return true;
  if (c->md_failed) { set_err("the double-precision model could not be loaded earlier"); return false; }
  std::string err;
  if (!load_model_f64(c->path.c_str(), c->path_dbl, c->md, err)) {
    c->md_failed = true;
    set_err("double-precision detector: %s", err.c_str());
    return false;
  }
  if (c->md.any_scaled) {
    c->md.loaded = false; c->md_failed = true;
    set_err("double-precision detector: the model has scale != 0 nodes; they sample cv::resize'd planes "
            "(cascador.cpp:330-331), which this library does not reproduce");
    return false;
  }
  const HostModelD &m = c->md;
  c->filter64_ok = m.stage >= 1 && count_normed_stage0(m) <= kMaxNorm && m.K * kCartBytes <= kMaxStage0TableBytes &&
                   stage0_filter_margins(m, c->margins64);
  return true;
}

bool ensure_model64_impl(Context *c) {
  const HostModelD &m = c->md;
  CU_OK(cudaMalloc(&c->d_nodes64, m.nodes.size() * sizeof(NodeRecD)));
  CU_OK(cudaMalloc(&c->d_leaf64, m.leaf.size() * 8));
  CU_OK(cudaMalloc(&c->d_cart64, m.cart3.size() * 8));
  CU_OK(cudaMalloc(&c->d_w64, m.w.size() * 8));
  CU_OK(cudaMalloc(&c->d_mean64, m.mean_shape.size() * 8));
  CU_OK(cudaMalloc(&c->d_norms64, kMaxNorm * sizeof(Stage0Norm)));
  CU_OK(cudaMemcpy(c->d_nodes64, m.nodes.data(), m.nodes.size() * sizeof(NodeRecD), cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(c->d_leaf64, m.leaf.data(), m.leaf.size() * 8, cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(c->d_cart64, m.cart3.data(), m.cart3.size() * 8, cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(c->d_w64, m.w.data(), m.w.size() * 8, cudaMemcpyHostToDevice));
  CU_OK(cudaMemcpy(c->d_mean64, m.mean_shape.data(), m.mean_shape.size() * 8, cudaMemcpyHostToDevice));
  CU_OK(cudaFuncSetAttribute(k4_cascade_f64<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(K4_WARPS * (kMaxDim * 8 + 4096))));  // largest K (see ctx_init: the attribute is per device)
  CU_OK(cudaFuncSetAttribute(k4_cascade_f64<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(K4_WARPS * (kMaxDim * 8 + 4096))));
  return true;
}

// all-or-nothing like ctx_init: a failed upload frees its partial allocations, the next call starts over
bool ensure_model64(Context *c) {
  if (c->model64_ready) return true;
  if (!ensure_model64_host(c)) return false;
  if (!ensure_model64_impl(c)) {
    const std::string keep = g_err;
    release_model64(c);
    cudaGetLastError();
    g_err = keep;
    return false;
  }
  c->model64_ready = true;
  return true;
}

bool ensure_geometry64(Context *c, int w, int h, int minimum_size, int step, double factor, bool latency) {
  Geometry &g = c->geo64;
  if (g.valid && g.w == w && g.h == h && g.min_size == minimum_size && g.step64 == step && g.factor64 == factor &&
      g.plan == (latency ? 1 : 0))
    return true;
  g.valid = false;
  g.w = w; g.h = h; g.min_size = minimum_size; g.max_size = 0; g.scale = 0.f; g.step64 = step; g.factor64 = factor;
  g.plan = latency ? 1 : 0;
  int wins[kMaxLevels + 1];
  const int n = enumerate_levels_f64(w, h, minimum_size, factor, wins, kMaxLevels + 1);
  if (n > kMaxLevels) { set_err("more than %d pyramid levels (scale too close to 1)", kMaxLevels); return false; }
  g.n_levels = n;
  g.table_bytes = (c->md.K * kCartBytes + 127) & ~127;
  long long base = 0;
  for (int i = 0; i < n; i++) {
    LevelInfo &L = g.lv[i];
    memset(&L, 0, sizeof L);
    L.win = wins[i];
    L.step = step;
    if (L.win >= 2048) { set_err("windows of %d px exceed the 2047 px limit", L.win); return false; }
    L.nx = (w - L.win) / step + 1;
    L.ny = (h - L.win) / step + 1;
    if (L.nx > 8191 || L.ny > 8191) { set_err("frame too large for 13-bit window indices"); return false; }
    plan_level(c->tune, L, latency ? 1 : 0);
    L.table_off = i * g.table_bytes;
    L.win_base = base;
    base += (long long)L.nx * L.ny;
  }
  g.windows_per_frame = base;
  if (n > 0 && c->filter64_ok) {
    std::vector<uint8_t> tab((size_t)n * g.table_bytes, 0);
    Stage0Norm norms[kMaxNorm];
    memset(norms, 0, sizeof norms);
    for (int i = 0; i < n; i++)
      build_stage0_table_f64(c->md, c->margins64, g.lv[i].win, g.lv[i].use_smem ? g.lv[i].box_w : 0,
                             tab.data() + (size_t)i * g.table_bytes, norms);
    if (!c->d_tables64.ensure(tab.size())) return false;
    CU_OK(cudaMemcpyAsync(c->d_tables64.p, tab.data(), tab.size(), cudaMemcpyHostToDevice, c->stream()));
    CU_OK(cudaMemcpyAsync(c->d_norms64, norms, sizeof norms, cudaMemcpyHostToDevice, c->stream()));
    CU_OK(cudaStreamSynchronize(c->stream()));
  }
  g.valid = true;
  return true;
}

// One batch of equally sized host frames through the double-precision detector; `hits` comes back in scan order.
bool run_device64(Context *c, const unsigned char *frames, int n_frames, int width, int height,
                  const jdaB200CppParams &prm, std::vector<Hit64> &hits, int *trace_n, double *trace_s, bool timing) {
  hits.clear();
  jdaB200Stats &st = c->sc->last;
  memset(&st, 0, sizeof st);
  if (n_frames <= 0) return true;
  if (width <= 0 || height <= 0) { set_err("bad frame size"); return false; }
  if (prm.step <= 0 || prm.minimum_size <= 0) { set_err("fddb.step and fddb.minimum_size must be positive"); return false; }
  if (!ctx_init(c) || !ensure_model64(c)) return false;
  for (const Scratch &sl : c->slot)
    if (sl.busy) { set_err("a submitted batch is waiting for jdaB200Collect: collect it before a synchronous call"); return false; }
  c->sc = &c->slot[0];
  const HostModelD &m = c->md;
  jdaB200Batch b;
  memset(&b, 0, sizeof b);
  b.n_frames = n_frames; b.width = width; b.height = height; b.pitch = width; b.frame_stride = (size_t)width * height;
  Run R;
  memset(&R, 0, sizeof R);
  R.c = c; R.b = &b; R.timing = timing;
  R.latency_plan = n_frames <= kLatencyFrames;
  if (c->tune.force_plan) R.latency_plan = c->tune.force_plan == 1;
  if (!ensure_geometry64(c, width, height, prm.minimum_size, prm.step, prm.scale, R.latency_plan)) return false;
  const Geometry &g = c->geo64;
  R.geo = &g; R.tables = c->d_tables64.p; R.norms = c->d_norms64;
  st.n_levels = g.n_levels;
  st.windows = g.windows_per_frame * n_frames;
  if (g.n_levels == 0) return true;
  const bool tracing = trace_n || trace_s;
  R.s = c->stream();
  R.D = m.D();
  R.leaf_pad = leaf_bytes(m.K);
  R.scan_K = m.K;  // (R.t_run stays 0: the prefilter's survivors are re-evaluated from cart 0, no leaf records needed)
  R.total_windows = st.windows;
  // the prefilter's tables are built for the mean shape itself: a shifted initial shape goes through the double kernel
  // from cart 0 (the similarity transform of stage 0 is the identity for shape == mean shape, bit for bit)
  R.use_scan = c->filter64_ok && !(prm.flags & JDA_B200_NO_STAGE0_SCAN) && !tracing && prm.shift_x == 0.0 && prm.shift_y == 0.0;
  cudaStream_t s = R.s;
  if (timing) CU_OK(cudaEventRecord(c->sc->ev[0], s));
  if (!stage_frames(R, frames)) return false;
  if (timing) CU_OK(cudaEventRecord(c->sc->ev[1], s));
  if (tracing) {
    if (!c->d_trace_n64.ensure(R.total_windows) || !c->d_trace_s64.ensure(R.total_windows)) return false;
    CU_OK(cudaMemsetAsync(c->d_trace_n64.p, 0, R.total_windows * 4, s));
    CU_OK(cudaMemsetAsync(c->d_trace_s64.p, 0, R.total_windows * 8, s));
  }
  if (c->surv_cap == 0) c->surv_cap = 1 << 16;
  c->surv_cap = std::max(c->surv_cap, (size_t)n_frames * 1024);
  if (c->hit64_cap == 0) c->hit64_cap = 1 << 14;
  c->hit64_cap = std::max(c->hit64_cap, (size_t)n_frames * 512);
  const int rec = kHit64Header + R.D;
  for (int attempt = 0; attempt < 4; attempt++) {
    if (!c->sc->d_surv.ensure(c->surv_cap) || !c->d_hits64.ensure(c->hit64_cap * rec)) return false;
    if (R.use_scan && !c->sc->d_surv_leaves.ensure(c->surv_cap * R.leaf_pad)) return false;
    CU_OK(cudaMemsetAsync(c->sc->d_counters, 0, kCntTotal * sizeof(unsigned), s));
    if (timing) CU_OK(cudaEventRecord(c->sc->ev[2], s));
    if (R.use_scan && !launch_scan(R)) return false;
    if (!R.use_scan && R.host_chunks)  // dense mode reads every frame: wait for the whole copy
      CU_OK(cudaStreamWaitEvent(s, c->sc->ev_copy[R.nchunks - 1], 0));
    if (timing) CU_OK(cudaEventRecord(c->sc->ev[3], s));
    Cascade64Params Q;
    memset(&Q, 0, sizeof Q);
    Q.frames = R.d_frames; Q.frame_stride = R.fstride; Q.pitch = R.pitch;
    Q.nodes = c->d_nodes64; Q.leaf = c->d_leaf64; Q.cart3 = c->d_cart64; Q.w = c->d_w64; Q.mean_shape = c->d_mean64;
    Q.T = m.T; Q.K = m.K; Q.L = m.L; Q.stage = m.stage; Q.cart_last = m.cart;
    Q.n_levels = g.n_levels;
    for (int i = 0; i < g.n_levels; i++) {
      Q.lv_win[i] = g.lv[i].win; Q.lv_step[i] = g.lv[i].step; Q.lv_nx[i] = g.lv[i].nx; Q.lv_ny[i] = g.lv[i].ny;
      Q.lv_base[i] = g.lv[i].win_base;
    }
    Q.windows_per_frame = g.windows_per_frame;
    Q.surv = c->sc->d_surv.p; Q.surv_count = c->sc->d_counters + kCntSurv; Q.surv_cap = (unsigned)c->surv_cap;
    Q.dense = R.use_scan ? 0 : 1; Q.dense_total = R.total_windows;
    Q.hits = c->d_hits64.p; Q.hit_count = c->sc->d_counters + kCntHit; Q.hit_cap = (unsigned)c->hit64_cap;
    Q.rec_doubles = rec;
    Q.work_counter = c->sc->d_counters + kCntWork;
    Q.trace_n = tracing ? c->d_trace_n64.p : nullptr;
    Q.trace_s = tracing ? c->d_trace_s64.p : nullptr;
    Q.similarity = prm.similarity_transform ? 1 : 0; Q.shift_x = prm.shift_x; Q.shift_y = prm.shift_y;
    if (Q.similarity) k4_cascade_f64<true><<<c->sm_count * 8, K4_WARPS * 32, K4_WARPS * (kMaxDim * 8 + ((m.K + 15) & ~15)), s>>>(Q);
    else k4_cascade_f64<false><<<c->sm_count * 8, K4_WARPS * 32, K4_WARPS * (kMaxDim * 8 + ((m.K + 15) & ~15)), s>>>(Q);
    CU_OK(cudaGetLastError());
    st.cascade_launches++;
    if (timing) CU_OK(cudaEventRecord(c->sc->ev[4], s));
    CU_OK(cudaMemcpyAsync(c->sc->h_counters, c->sc->d_counters, kCntTotal * sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    CU_OK(cudaStreamSynchronize(s));
    const size_t ns = c->sc->h_counters[kCntSurv], nh = c->sc->h_counters[kCntHit];
    if ((R.use_scan && ns > c->surv_cap) || nh > c->hit64_cap) {
      if (ns > c->surv_cap) c->surv_cap = ns + ns / 4;
      if (nh > c->hit64_cap) c->hit64_cap = nh + nh / 4;
      continue;
    }
    st.stage0_survivors = R.use_scan ? (long long)ns : 0;
    st.raw_hits = (long long)nh;
    c->h_hits64.resize(std::max(nh, (size_t)1) * rec);
    if (nh) CU_OK(cudaMemcpyAsync(c->h_hits64.data(), c->d_hits64.p, nh * rec * 8, cudaMemcpyDeviceToHost, s));
    if (trace_n) CU_OK(cudaMemcpyAsync(trace_n, c->d_trace_n64.p, R.total_windows * 4, cudaMemcpyDeviceToHost, s));
    if (trace_s) CU_OK(cudaMemcpyAsync(trace_s, c->d_trace_s64.p, R.total_windows * 8, cudaMemcpyDeviceToHost, s));
    if (timing) CU_OK(cudaEventRecord(c->sc->ev[5], s));
    CU_OK(cudaStreamSynchronize(s));
    hits.resize(nh);
    for (size_t i = 0; i < nh; i++) {
      const double *r = c->h_hits64.data() + i * rec;
      const int *ri = reinterpret_cast<const int *>(r);
      hits[i] = Hit64{ri[0], (uint32_t)ri[1], ri[2], ri[3], ri[4], ri[5], r[3], r + kHit64Header};
    }
    std::sort(hits.begin(), hits.end(), [](const Hit64 &a, const Hit64 &b2) {
      return a.frame != b2.frame ? a.frame < b2.frame : a.key < b2.key;
    });
    if (timing) {
      cudaEventElapsedTime(&st.ms_h2d, c->sc->ev_copy[kMaxChunks], c->sc->ev_copy[R.host_chunks ? R.nchunks - 1 : 0]);
      cudaEventElapsedTime(&st.ms_scan, c->sc->ev[2], c->sc->ev[3]);
      cudaEventElapsedTime(&st.ms_cascade, c->sc->ev[3], c->sc->ev[4]);
      cudaEventElapsedTime(&st.ms_d2h, c->sc->ev[4], c->sc->ev[5]);
    }
    return true;
  }
  set_err("survivor / hit queues kept overflowing");
  return false;
}

jdaB200ResultF64 empty_result64(int L, int n) {
  jdaB200ResultF64 r;
  r.n = n; r.landmark_n = L; r.rects = nullptr; r.scores = nullptr; r.shapes = nullptr;
  return r;
}

// cascador.cpp:445-474: nms (or every hit), then shape = rect.xy + shape * rect.wh, output in pick order
jdaB200ResultF64 finish_frame64(const HostModelD &m, const Hit64 *h, int n, const jdaB200CppParams &prm) {
  const int D = m.D();
  std::vector<int> rects(4 * (size_t)n);
  std::vector<double> sc(n);
  for (int i = 0; i < n; i++) {
    rects[4 * i] = h[i].x; rects[4 * i + 1] = h[i].y; rects[4 * i + 2] = h[i].win; rects[4 * i + 3] = h[i].win;
    sc[i] = h[i].score;
  }
  std::vector<int> picked;
  if (prm.nms) picked = nms_f64(n, rects.data(), sc.data(), prm.overlap);
  else { picked.resize(n); for (int i = 0; i < n; i++) picked[i] = i; }
  const int k = (int)picked.size();
  jdaB200ResultF64 r = empty_result64(m.L, 0);
  r.rects = (int *)malloc(sizeof(int) * 4 * (k > 0 ? k : 1));
  r.scores = (double *)malloc(sizeof(double) * (k > 0 ? k : 1));
  r.shapes = (double *)malloc(sizeof(double) * D * (k > 0 ? k : 1));
  if (!r.rects || !r.scores || !r.shapes) {
    free(r.rects); free(r.scores); free(r.shapes);
    return empty_result64(m.L, -1);
  }
  for (int i = 0; i < k; i++) {
    const Hit64 &hi = h[picked[i]];
    r.rects[4 * i] = hi.x; r.rects[4 * i + 1] = hi.y; r.rects[4 * i + 2] = hi.win; r.rects[4 * i + 3] = hi.win;
    r.scores[i] = hi.score;
    for (int j = 0; j < m.L; j++) {
      r.shapes[(size_t)i * D + 2 * j] = hi.x + hi.shape[2 * j] * hi.win;
      r.shapes[(size_t)i * D + 2 * j + 1] = hi.y + hi.shape[2 * j + 1] * hi.win;
    }
  }
  r.n = k;
  return r;
}

jdaResult empty_result(int L, int n) {
  jdaResult r;
  r.n = n; r.landmark_n = L; r.bboxes = nullptr; r.shapes = nullptr; r.scores = nullptr;
  return r;
}

// hits of one frame (scan order) -> jdaResult, c/jda.c:463-474
jdaResult finish_frame(const HostModel &m, const HitRec *h, int n, bool raw) {
  const int D = m.D();
  jdaResult r = empty_result(m.L, 0);
  // like the reference, the arrays exist even when n == 0
  r.bboxes = (int *)malloc(sizeof(int) * 3 * (n > 0 ? n : 1));
  r.scores = (float *)malloc(sizeof(float) * (n > 0 ? n : 1));
  r.shapes = (float *)malloc(sizeof(float) * D * (n > 0 ? n : 1));
  if (!r.bboxes || !r.scores || !r.shapes) {
    free(r.bboxes); free(r.scores); free(r.shapes);
    return empty_result(m.L, -1);
  }
  std::vector<int> box(3 * (size_t)n);
  std::vector<float> sc(n);
  for (int i = 0; i < n; i++) {
    box[3 * i] = h[i].x; box[3 * i + 1] = h[i].y; box[3 * i + 2] = h[i].win;
    sc[i] = h[i].score;
  }
  std::vector<uint8_t> keep(n > 0 ? n : 1, 1);
  if (!raw) nms(n, box.data(), sc.data(), keep.data());
  int o = 0;
  for (int i = 0; i < n; i++) {
    if (!keep[i]) continue;
    r.bboxes[3 * o] = h[i].x; r.bboxes[3 * o + 1] = h[i].y; r.bboxes[3 * o + 2] = h[i].win;
    r.scores[o] = h[i].score;
    if (raw) memcpy(r.shapes + (size_t)o * D, h[i].shape, sizeof(float) * D);
    else relocate(h[i].shape, r.shapes + (size_t)o * D, m.L, h[i].x, h[i].y, h[i].win);
    o++;
  }
  r.n = o;
  return r;
}

// hits (sorted by frame, scan order inside a frame) -> one jdaResult per frame; returns the detection count
long long finish_frames(Context *c, const std::vector<HitRec> &hits, int n_frames, bool raw, jdaResult *results,
                        const int *frame_map /* local -> caller's index, or NULL */) {
  size_t i = 0;
  long long dets = 0;
  for (int f = 0; f < n_frames; f++) {
    size_t j = i;
    while (j < hits.size() && hits[j].frame == f) j++;
    jdaResult &r = results[frame_map ? frame_map[f] : f];
    r = finish_frame(c->m, hits.data() + i, (int)(j - i), raw);
    dets += r.n > 0 ? r.n : 0;
    i = j;
  }
  return dets;
}

// hits (sorted by frame, scan order inside a frame) -> one flat result for the whole batch: three arrays instead of
// three per frame.  NMS and relocation per frame exactly as finish_frame does them.
bool finish_flat(Context *c, const std::vector<HitRec> &hits, int n_frames, bool raw, jdaB200FlatResult *out) {
  const HostModel &m = c->m;
  const int D = m.D();
  out->n_frames = n_frames; out->landmark_n = m.L; out->total = 0;
  out->counts = (int *)calloc(n_frames > 0 ? n_frames : 1, sizeof(int));
  std::vector<uint8_t> keep(hits.size() > 0 ? hits.size() : 1, 1);
  std::vector<int> box;
  std::vector<float> sc;
  size_t i = 0, total = 0;
  for (int f = 0; f < n_frames && out->counts; f++) {
    size_t j = i;
    while (j < hits.size() && hits[j].frame == f) j++;
    const int n = (int)(j - i);
    if (n > 0 && !raw) {
      box.resize(3 * (size_t)n); sc.resize(n);
      for (int q = 0; q < n; q++) {
        box[3 * q] = hits[i + q].x; box[3 * q + 1] = hits[i + q].y; box[3 * q + 2] = hits[i + q].win;
        sc[q] = hits[i + q].score;
      }
      nms(n, box.data(), sc.data(), keep.data() + i);
    }
    int kept = 0;
    for (int q = 0; q < n; q++) kept += keep[i + q];
    out->counts[f] = kept;
    total += kept;
    i = j;
  }
  out->bboxes = (int *)malloc(sizeof(int) * 3 * (total > 0 ? total : 1));
  out->scores = (float *)malloc(sizeof(float) * (total > 0 ? total : 1));
  out->shapes = (float *)malloc(sizeof(float) * D * (total > 0 ? total : 1));
  if (!out->counts || !out->bboxes || !out->scores || !out->shapes) {
    free(out->counts); free(out->bboxes); free(out->scores); free(out->shapes);
    out->counts = nullptr; out->bboxes = nullptr; out->scores = nullptr; out->shapes = nullptr;
    out->n_frames = -1;
    set_err("out of host memory");
    return false;
  }
  size_t o = 0;
  for (size_t q = 0; q < hits.size(); q++) {
    if (!keep[q]) continue;
    const HitRec &h = hits[q];
    out->bboxes[3 * o] = h.x; out->bboxes[3 * o + 1] = h.y; out->bboxes[3 * o + 2] = h.win;
    out->scores[o] = h.score;
    if (raw) memcpy(out->shapes + o * D, h.shape, sizeof(float) * D);
    else relocate(h.shape, out->shapes + o * D, m.L, h.x, h.y, h.win);
    o++;
  }
  out->total = (int)total;
  return true;
}

int detect_batch(Context *c, const unsigned char *frames, const jdaB200Batch &b, jdaResult *results,
                 jdaB200Stats *stats, jdaB200FlatResult *flat = nullptr) {
  std::lock_guard<std::mutex> lock(c->mu);
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  std::vector<HitRec> hits;
  const bool ok = run_device(c, frames, b, hits, nullptr, stats != nullptr);
  if (!ok) {
    for (int f = 0; f < b.n_frames && results; f++) results[f] = empty_result(c->m.L, -1);
    if (flat) { memset(flat, 0, sizeof *flat); flat->n_frames = -1; flat->landmark_n = c->m.L; }
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return -1;
  }
  const auto t0 = std::chrono::steady_clock::now();
  bool fin = true;
  if (flat) {
    fin = finish_flat(c, hits, b.n_frames, (b.flags & JDA_B200_RAW_HITS) != 0, flat);
    c->sc->last.detections = fin ? flat->total : 0;
  } else {
    c->sc->last.detections = finish_frames(c, hits, b.n_frames, (b.flags & JDA_B200_RAW_HITS) != 0, results, nullptr);
  }
  c->sc->last.ms_host = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (stats) *stats = c->sc->last;
  if (prev_dev >= 0) cudaSetDevice(prev_dev);
  return fin ? 0 : -1;
}

void add_stats(jdaB200Stats &a, const jdaB200Stats &b) {
  a.windows += b.windows; a.stage0_survivors += b.stage0_survivors; a.raw_hits += b.raw_hits; a.detections += b.detections;
  a.ms_h2d += b.ms_h2d; a.ms_resize += b.ms_resize; a.ms_scan += b.ms_scan; a.ms_cascade += b.ms_cascade;
  a.ms_d2h += b.ms_d2h; a.ms_host += b.ms_host;
  a.scan_launches += b.scan_launches; a.cascade_launches += b.cascade_launches; a.resize_launches += b.resize_launches;
  a.n_levels = std::max(a.n_levels, b.n_levels); a.levels_smem = std::max(a.levels_smem, b.levels_smem);
}

// Frames of different sizes in one call (FDDB-shaped frames).  Result f is what jdaDetect
// would return for frame f alone.  One launch over a common canvas when stage 0 runs from the scan kernel; models
// that need the generic per-window kernel for stage 0, or sample the h / q planes, go shape group by shape group.
int detect_mixed(Context *c, const jdaB200Frame *frames, int n, float scale, int min_size, int max_size, float th,
                 int t_limit, int flags, jdaResult *results, jdaB200Stats *stats) {
  std::lock_guard<std::mutex> lock(c->mu);
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  const bool raw = (flags & JDA_B200_RAW_HITS) != 0;
  flags &= ~JDA_B200_DEVICE_INPUT;
  auto fail = [&]() {
    for (int f = 0; f < n; f++) results[f] = empty_result(c->m.L, -1);
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return -1;
  };
  for (int f = 0; f < n; f++) {
    results[f] = empty_result(c->m.L, 0);
    if (!frames[f].data && frames[f].width > 0 && frames[f].height > 0) { set_err("frame %d has no data", f); return fail(); }
    if (frames[f].pitch > 0 && frames[f].pitch < frames[f].width) { set_err("frame %d: pitch < width", f); return fail(); }
  }
  jdaB200Batch b;
  memset(&b, 0, sizeof b);
  b.scale = scale; b.min_size = min_size; b.max_size = max_size; b.th = th; b.t_limit = t_limit; b.flags = flags;
  std::vector<HitRec> hits;
  const bool canvas = c->m.stage0_lut_ok && !c->m.any_scaled && !(flags & JDA_B200_NO_STAGE0_SCAN);
  if (canvas) {
    int wc = 0, hc = 0, top = 0;
    for (int f = 0; f < n; f++) {
      wc = std::max(wc, frames[f].width); hc = std::max(hc, frames[f].height);
      top = std::max(top, std::min(frames[f].width, frames[f].height));
    }
    if (wc <= 0 || hc <= 0) {  // nothing but empty frames
      for (int f = 0; f < n; f++) results[f] = finish_frame(c->m, nullptr, 0, raw);
      memset(&c->sc->last, 0, sizeof c->sc->last);
      if (stats) *stats = c->sc->last;
      if (prev_dev >= 0) cudaSetDevice(prev_dev);
      return 0;
    }
    // the canvas visits the window sizes some frame has; every frame then keeps those that fit inside itself
    b.n_frames = n; b.width = wc; b.height = hc;
    const int user_max = max_size;
    b.max_size = user_max > 0 ? std::min(user_max, top) : top;
    if (!run_device(c, nullptr, b, hits, nullptr, stats != nullptr, frames)) return fail();
    const auto t0 = std::chrono::steady_clock::now();
    c->sc->last.detections = finish_frames(c, hits, n, raw, results, nullptr);
    c->sc->last.ms_host = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (stats) *stats = c->sc->last;
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return 0;
  }
  std::map<std::pair<int, int>, std::vector<int>> groups;
  for (int f = 0; f < n; f++) {
    if (frames[f].width <= 0 || frames[f].height <= 0) results[f] = finish_frame(c->m, nullptr, 0, raw);
    else groups[{frames[f].width, frames[f].height}].push_back(f);
  }
  jdaB200Stats total;
  memset(&total, 0, sizeof total);
  std::vector<unsigned char> pack;
  for (const auto &g : groups) {
    const int w = g.first.first, h = g.first.second, cnt = (int)g.second.size();
    pack.resize((size_t)w * h * cnt);
    for (int i = 0; i < cnt; i++) {
      const jdaB200Frame &fr = frames[g.second[i]];
      for (int y = 0; y < h; y++)
        memcpy(pack.data() + ((size_t)i * h + y) * w, fr.data + (size_t)y * (fr.pitch > 0 ? fr.pitch : w), w);
    }
    b.n_frames = cnt; b.width = w; b.height = h; b.pitch = w; b.frame_stride = (size_t)w * h;
    if (!run_device(c, pack.data(), b, hits, nullptr, stats != nullptr)) {
      for (int f = 0; f < n; f++) jdaResultRelease(results[f]);
      return fail();
    }
    const auto t0 = std::chrono::steady_clock::now();
    c->sc->last.detections = finish_frames(c, hits, cnt, raw, results, g.second.data());
    c->sc->last.ms_host = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    add_stats(total, c->sc->last);
  }
  c->sc->last = total;
  if (stats) *stats = total;
  if (prev_dev >= 0) cudaSetDevice(prev_dev);
  return 0;
}

void *create(const char *path, bool dbl) {
  if (!path) return nullptr;
  Context *c = new Context();
  std::string err;
  if (!load_model(path, dbl, c->m, err)) {
    g_err = err;
    delete c;
    return nullptr;
  }
  memset(&c->sc->last, 0, sizeof c->sc->last);
  c->tune = read_tuning();  // A/B knobs and test hooks: the environment is read once per handle, here
  c->path = path;
  c->path_dbl = dbl;
  return c;
}

}  // namespace

// =============================================================================== C ABI

extern "C" {

void *jdaCascadorCreateDouble(const char *model) { return create(model, true); }
void *jdaCascadorCreateFloat(const char *model) { return create(model, false); }

void jdaCascadorSerializeTo(void *cascador, const char *model) {
  if (!cascador || !model) return;
  save_model_f32(static_cast<Context *>(cascador)->m, model);
}

int jdaB200SerializeTo(void *cascador, const char *model, int flags) {
  if (!cascador || !model) return -2;
  return save_model(static_cast<Context *>(cascador)->m, model, (flags & JDA_B200_SAVE_STAGE_T) != 0,
                    (flags & JDA_B200_SAVE_DOUBLE) != 0) ? 0 : -1;
}

void jdaCascadorRelease(void *cascador) {
  if (cascador) ctx_free(static_cast<Context *>(cascador));
}

// c/jda.h:52-60.  The reference's jdaDetect is a pure function of a read-only cascador, so a host may call it from as
// many threads as it likes and they all make progress.  Here one device serves the handle: calls that arrive while an
// earlier one is running are COALESCED -- the caller that finds nobody serving takes every queued call with the same
// (scale, min_size, max_size, th) and runs them as one mixed-size batch (detect_mixed: result f is bit for bit what
// frame f alone gives), hands out the results and wakes the others.  A lone caller takes the one-frame path at once:
// nobody ever waits for a batch to fill.
constexpr int kMaxCoalesced = 256;

jdaResult jdaDetect(void *cascador, unsigned char *data, int width, int height, float scale, float step,
                    int min_size, int max_size, float th) {
  (void)step;  // ignored by the reference as well (c/jda.c:333)
  Context *c = static_cast<Context *>(cascador);
  if (!c || !data) return empty_result(c ? c->m.L : 0, -1);
  if (width <= 0 || height <= 0) return finish_frame(c->m, nullptr, 0, false);
  DetectReq rq;
  rq.data = data; rq.width = width; rq.height = height;
  rq.scale = scale; rq.min_size = min_size; rq.max_size = max_size; rq.th = th;
  rq.res = empty_result(c->m.L, -1);
  rq.done = false;
  std::unique_lock<std::mutex> lk(c->cq_mu);
  c->cq.push_back(&rq);
  c->cq_calls++;
  while (!rq.done) {
    if (c->cq_leader) {
      c->cq_cv.wait(lk);
      continue;
    }
    // serve the oldest call and everything queued with its parameters (mine may be in a later group)
    c->cq_leader = true;
    std::vector<DetectReq *> grp;
    const DetectReq &lead = *c->cq.front();
    const bool can_mix = c->m.stage0_lut_ok && !c->m.any_scaled;  // (else a mixed batch would run shape by shape anyway)
    for (size_t i = 0; i < c->cq.size();) {
      DetectReq *r = c->cq[i];
      const bool same = r->scale == lead.scale && r->min_size == lead.min_size && r->max_size == lead.max_size && r->th == lead.th;
      if (same && (grp.empty() || (can_mix && (int)grp.size() < kMaxCoalesced))) {
        grp.push_back(r);
        c->cq.erase(c->cq.begin() + i);
      } else {
        i++;
      }
    }
    c->cq_batches++;
    c->cq_largest = std::max(c->cq_largest, (int)grp.size());
    lk.unlock();
    if (grp.size() == 1) {
      DetectReq &r = *grp[0];
      jdaB200Batch b;
      memset(&b, 0, sizeof b);
      b.n_frames = 1; b.width = r.width; b.height = r.height; b.pitch = r.width;
      b.frame_stride = (size_t)r.width * r.height;
      b.scale = r.scale; b.min_size = r.min_size; b.max_size = r.max_size; b.th = r.th;
      detect_batch(c, r.data, b, &r.res, nullptr);
    } else {
      std::vector<jdaB200Frame> fr(grp.size());
      std::vector<jdaResult> out(grp.size());
      for (size_t i = 0; i < grp.size(); i++) {
        memset(&fr[i], 0, sizeof fr[i]);
        fr[i].data = grp[i]->data; fr[i].width = grp[i]->width; fr[i].height = grp[i]->height; fr[i].pitch = grp[i]->width;
      }
      const DetectReq &g0 = *grp[0];
      detect_mixed(c, fr.data(), (int)grp.size(), g0.scale, g0.min_size, g0.max_size, g0.th, 0, 0, out.data(), nullptr);
      for (size_t i = 0; i < grp.size(); i++) grp[i]->res = out[i];
    }
    lk.lock();
    for (DetectReq *r : grp) r->done = true;
    c->cq_leader = false;
    c->cq_cv.notify_all();
  }
  return rq.res;
}

void jdaResultRelease(jdaResult result) {
  free(result.bboxes);
  free(result.shapes);
  free(result.scores);
}

void jdaB200CoalescingStats(void *cascador, long long *calls, long long *batches, int *largest) {
  Context *c = static_cast<Context *>(cascador);
  if (!c) return;
  std::lock_guard<std::mutex> lk(c->cq_mu);
  if (calls) *calls = c->cq_calls;
  if (batches) *batches = c->cq_batches;
  if (largest) *largest = c->cq_largest;
}

int jdaB200DetectBatch(void *cascador, const unsigned char *frames, const jdaB200Batch *batch, jdaResult *results,
                       jdaB200Stats *stats) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || !batch || !results || (!frames && batch->n_frames > 0)) {
    set_err("null argument");
    return -2;
  }
  return detect_batch(c, frames, *batch, results, stats);
}

int jdaB200DetectMixed(void *cascador, const jdaB200Frame *frames, int n_frames, float scale, int min_size,
                       int max_size, float th, int t_limit, int flags, jdaResult *results, jdaB200Stats *stats) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || n_frames < 0 || (n_frames > 0 && (!frames || !results))) {
    set_err("null argument");
    return -2;
  }
  if (n_frames == 0) return 0;
  return detect_mixed(c, frames, n_frames, scale, min_size, max_size, th, t_limit, flags, results, stats);
}

int jdaB200JoinCascadorDetect(void *cascador, const unsigned char *frames, int n_frames, int width, int height,
                              const jdaB200CppParams *params, jdaB200ResultF64 *results, jdaB200Stats *stats) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || !params || n_frames < 0 || (n_frames > 0 && (!frames || !results))) {
    set_err("null argument");
    return -2;
  }
  if (n_frames == 0) return 0;
  std::lock_guard<std::mutex> lock(c->mu);
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  std::vector<Hit64> hits;
  bool ok = !(params->nms && !(params->overlap < 1.));  // a box would not suppress itself: endless loop in the reference
  if (!ok) set_err("fddb.overlap must be < 1 when nms is on");
  ok = ok && run_device64(c, frames, n_frames, width, height, *params, hits, nullptr, nullptr, stats != nullptr);
  if (!ok) {
    for (int f = 0; f < n_frames; f++) results[f] = empty_result64(c->md.L ? c->md.L : c->m.L, -1);
    if (prev_dev >= 0) cudaSetDevice(prev_dev);
    return -1;
  }
  const auto t0 = std::chrono::steady_clock::now();
  size_t i = 0;
  long long dets = 0;
  for (int f = 0; f < n_frames; f++) {
    size_t j = i;
    while (j < hits.size() && hits[j].frame == f) j++;
    results[f] = finish_frame64(c->md, hits.data() + i, (int)(j - i), *params);
    dets += results[f].n > 0 ? results[f].n : 0;
    i = j;
  }
  c->sc->last.detections = dets;
  c->sc->last.ms_host = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (stats) *stats = c->sc->last;
  if (prev_dev >= 0) cudaSetDevice(prev_dev);
  return 0;
}

void jdaB200ResultF64Release(jdaB200ResultF64 *results, int n) {
  if (!results) return;
  for (int i = 0; i < n; i++) {
    free(results[i].rects); free(results[i].scores); free(results[i].shapes);
    results[i].rects = nullptr; results[i].scores = nullptr; results[i].shapes = nullptr; results[i].n = 0;
  }
}

long long jdaB200JoinCascadorTrace(void *cascador, const unsigned char *frame, int width, int height,
                                   const jdaB200CppParams *params, int *carts_evaluated, double *exit_score) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || !frame || !params || (!carts_evaluated && !exit_score)) return -2;
  std::lock_guard<std::mutex> lock(c->mu);
  std::vector<Hit64> hits;
  // pass both pointers: run_device64 traces when either is set
  if (!run_device64(c, frame, 1, width, height, *params, hits, carts_evaluated, exit_score, false)) return -1;
  return c->sc->last.windows;
}

int jdaB200JoinCascadorFilterMargins(void *cascador, double *margins, int cap) {
  Context *c = static_cast<Context *>(cascador);
  if (!c) return -2;
  std::lock_guard<std::mutex> lock(c->mu);
  if (!ensure_model64_host(c)) return -1;
  if (!c->filter64_ok) return 0;
  for (int k = 0; k < c->md.K && k < cap && margins; k++) margins[k] = c->margins64[k];
  return c->md.K;
}

int jdaB200JoinCascadorLevels(int width, int height, int minimum_size, double scale, int *wins, int cap) {
  return enumerate_levels_f64(width, height, minimum_size, scale, wins, cap);
}

// ---- submit / collect: two batches in flight ----------------------------------------------------------------
// While batch i is scanned, batch i + 1 is already being copied in; while batch i + 1 is scanned, the host sorts,
// suppresses and relocates the hits of batch i.  Set t & 1 of the handle's two scratch sets serves ticket t.
static bool same_geometry(const Context *c, const jdaB200Batch &b, int plan) {
  const Geometry &g = c->geo;
  return g.valid && g.w == b.width && g.h == b.height && g.scale == b.scale && g.min_size == b.min_size &&
         g.max_size == b.max_size && g.plan == plan;
}

int jdaB200Submit(void *cascador, const unsigned char *frames, const jdaB200Batch *batch) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || !batch || (!frames && batch->n_frames > 0)) {
    set_err("null argument");
    return -2;
  }
  std::lock_guard<std::mutex> lock(c->mu);
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  auto done = [&](int rc) { if (prev_dev >= 0) cudaSetDevice(prev_dev); return rc; };
  if (!ctx_init(c)) return done(-1);
  const int ticket = c->next_ticket;
  Scratch &sl = c->slot[ticket & 1], &other = c->slot[(ticket & 1) ^ 1];
  if (sl.busy) {
    set_err("two batches are already in flight: jdaB200Collect ticket %d first", sl.ticket);
    return done(-3);
  }
  if (!slot_init(sl)) return done(-1);
  // the stage-0 tables belong to a geometry: a batch of another size or pyramid may only rebuild them once the
  // batch that is still using them has finished
  if (other.busy && !same_geometry(c, *batch, batch_plan(c->tune, *batch, nullptr))) cudaEventSynchronize(other.ev_done);
  c->sc = &sl;
  sl.batch = *batch;
  sl.frames = frames;
  if (!sl.run) sl.run = new Run();
  Run &R = *sl.run;
  const bool prev_running = other.busy && cudaEventQuery(other.ev_done) == cudaErrorNotReady;
  cudaGetLastError();
  if (!run_prepare(c, R, frames, sl.batch, nullptr, true, nullptr, prev_running ? 2 : 1) || (!R.done && !run_enqueue(R))) {
    c->sc = &c->slot[0];
    return done(-1);
  }
  sl.busy = true;
  sl.ticket = ticket;
  c->next_ticket++;
  c->sc = &c->slot[0];
  return done(ticket);
}

int jdaB200Collect(void *cascador, int ticket, jdaB200FlatResult *result, jdaB200Stats *stats) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || !result || ticket < 0) {
    set_err("null argument");
    return -2;
  }
  std::lock_guard<std::mutex> lock(c->mu);
  Scratch &sl = c->slot[ticket & 1], &other = c->slot[(ticket & 1) ^ 1];
  memset(result, 0, sizeof *result);
  result->n_frames = -1; result->landmark_n = c->m.L;
  if (!sl.busy || sl.ticket != ticket) {
    set_err("ticket %d is not in flight", ticket);
    return -3;
  }
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  cudaSetDevice(c->device);
  c->sc = &sl;
  Run &R = *sl.run;
  std::vector<HitRec> hits;
  bool ok = true;
  for (int attempt = 0; attempt < 4 && ok && !R.done; attempt++) {
    bool overflow = false;
    ok = run_finish(R, hits, overflow);
    if (!ok || !overflow) break;
    // a queue overflowed (the capacities have grown): the batch runs again from its frames -- which the caller keeps
    // valid until the ticket is collected -- after the other batch in flight has let go of the shared tables
    if (attempt == 3) { set_err("survivor / hit queues kept overflowing"); ok = false; break; }
    if (other.busy) cudaEventSynchronize(other.ev_done);
    ok = run_prepare(c, R, sl.frames, sl.batch, nullptr, true, nullptr, false) && (R.done || run_enqueue(R));
  }
  sl.busy = false;
  const auto t0 = std::chrono::steady_clock::now();
  if (ok) ok = finish_flat(c, hits, sl.batch.n_frames, (sl.batch.flags & JDA_B200_RAW_HITS) != 0, result);
  sl.last.detections = ok ? result->total : 0;
  sl.last.ms_host = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (stats) *stats = sl.last;
  c->sc = &c->slot[0];
  if (prev_dev >= 0) cudaSetDevice(prev_dev);
  return ok ? 0 : -1;
}

int jdaB200DetectBatchFlat(void *cascador, const unsigned char *frames, const jdaB200Batch *batch,
                           jdaB200FlatResult *result, jdaB200Stats *stats) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || !batch || !result || (!frames && batch->n_frames > 0)) {
    set_err("null argument");
    return -2;
  }
  return detect_batch(c, frames, *batch, nullptr, stats, result);
}

void jdaB200FlatResultRelease(jdaB200FlatResult *r) {
  if (!r) return;
  free(r->counts); free(r->bboxes); free(r->scores); free(r->shapes);
  r->counts = nullptr; r->bboxes = nullptr; r->scores = nullptr; r->shapes = nullptr;
  r->total = 0;
}

void jdaB200ResultsRelease(jdaResult *results, int n) {
  if (!results) return;
  for (int i = 0; i < n; i++) {
    jdaResultRelease(results[i]);
    results[i].bboxes = nullptr; results[i].shapes = nullptr; results[i].scores = nullptr; results[i].n = 0;
  }
}

int jdaB200SetDevice(void *cascador, int device) {
  Context *c = static_cast<Context *>(cascador);
  if (!c) return -2;
  std::lock_guard<std::mutex> lock(c->mu);
  if (c->inited && c->device != device) {
    set_err("device already bound to %d", c->device);
    return -1;
  }
  c->device = device;
  return 0;
}

int jdaB200SetStream(void *cascador, void *cuda_stream) {
  Context *c = static_cast<Context *>(cascador);
  if (!c) return -2;
  std::lock_guard<std::mutex> lock(c->mu);
  c->user_stream = static_cast<cudaStream_t>(cuda_stream);
  return 0;
}

void jdaB200ModelDims(void *cascador, int *out4) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || !out4) return;
  out4[0] = c->m.T; out4[1] = c->m.K; out4[2] = c->m.L; out4[3] = c->m.depth;
}

const char *jdaB200LastError(void) { return g_err.c_str(); }

int jdaB200DeviceCount(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int jdaB200Levels(int width, int height, float scale, int min_size, int max_size, int *wins, int cap) {
  if (width < 24 || height < 24) return 0;
  return enumerate_levels(width, height, scale, min_size, max_size, wins, cap);
}

long long jdaB200CountWindows(int width, int height, float scale, int min_size, int max_size) {
  return count_windows(width, height, scale, min_size, max_size);
}

int jdaB200DescribePlan(int width, int height, float scale, int min_size, int max_size, char *buf, int cap) {
  const bool latency = cap < 0;  // negative cap: describe the latency plan (batches of <= 4 frames)
  if (cap < 0) cap = -cap;
  const Tuning tn = read_tuning();
  int wins[kMaxLevels + 1];
  int n = (width < 24 || height < 24) ? 0 : enumerate_levels(width, height, scale, min_size, max_size, wins, kMaxLevels + 1);
  n = std::min(n, kMaxLevels);
  int o = 0;
  if (buf && cap > 0) buf[0] = 0;
  for (int i = 0; i < n; i++) {
    LevelInfo L;
    memset(&L, 0, sizeof L);
    L.win = wins[i]; L.step = level_step(L.win);
    L.nx = (width - L.win) / L.step + 1; L.ny = (height - L.win) / L.step + 1;
    plan_level(tn, L, latency ? 1 : 0);
    if (buf && o < cap)
      o += snprintf(buf + o, cap - o, "%d %d %d %d %d %d %d %d %d %d %d\n", L.win, L.step, L.nx, L.ny, 1 << L.tw_log2,
                    L.th, L.box_w, L.box_h, L.use_smem, (1 << L.tw_log2) * L.th, L.span);
  }
  return n;
}

void jdaB200Nms(int n, const int *bboxes, const float *scores, unsigned char *keep) {
  nms(n, bboxes, scores, keep);
}

long long jdaB200Trace(void *cascador, const unsigned char *frame, int width, int height, float scale,
                       int min_size, int max_size, int t_limit, int flags, int *carts_evaluated,
                       float *exit_score, unsigned char *leaves, long long leaf_w0, long long leaf_w1) {
  return jdaB200TraceK(cascador, frame, width, height, scale, min_size, max_size, t_limit, 0, flags, carts_evaluated,
                       exit_score, leaves, leaf_w0, leaf_w1);
}

long long jdaB200TraceK(void *cascador, const unsigned char *frame, int width, int height, float scale,
                        int min_size, int max_size, int t_limit, int k_limit, int flags, int *carts_evaluated,
                        float *exit_score, unsigned char *leaves, long long leaf_w0, long long leaf_w1) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || !frame) return -2;
  std::lock_guard<std::mutex> lock(c->mu);
  jdaB200Batch b;
  memset(&b, 0, sizeof b);
  b.n_frames = 1; b.width = width; b.height = height; b.pitch = width;
  b.frame_stride = (size_t)width * height;
  b.scale = scale; b.min_size = min_size; b.max_size = max_size; b.th = 0.f; b.t_limit = t_limit; b.k_limit = k_limit;
  b.flags = (flags & ~JDA_B200_DEVICE_INPUT) | JDA_B200_NO_FINAL_TH;
  TraceOut t;
  t.n = carts_evaluated; t.s = exit_score; t.leaf = leaves; t.w0 = leaf_w0; t.w1 = leaf_w1;
  std::vector<HitRec> hits;
  if (!run_device(c, frame, b, hits, &t, false)) return -1;
  return c->sc->last.windows;
}

int jdaB200Resize(void *cascador, const unsigned char *src, int sw, int sh, unsigned char *dst, int dw, int dh) {
  Context *c = static_cast<Context *>(cascador);
  if (!c || !src || !dst || sw < 2 || sh < 2 || dw <= 0 || dh <= 0) return -2;
  std::lock_guard<std::mutex> lock(c->mu);
  if (!ctx_init(c)) return -1;
  // k1 produces both planes; ask for the same size twice and read the first
  DevBuf<uint8_t> in, out;
  if (!in.ensure((size_t)sw * sh + 16) || !out.ensure((size_t)2 * dw * dh)) return -1;
  cudaStream_t s = c->stream();
  bool ok = cudaMemcpyAsync(in.p, src, (size_t)sw * sh, cudaMemcpyHostToDevice, s) == cudaSuccess;
  dim3 grid((dw * dh + 255) / 256, 1, 1);
  k1_resize<<<grid, 256, 0, s>>>(in.p, 0, sw, sw, sh, out.p, 0, dw, dh, dw, dh);
  ok = ok && cudaGetLastError() == cudaSuccess;
  ok = ok && cudaMemcpyAsync(dst, out.p, (size_t)dw * dh, cudaMemcpyDeviceToHost, s) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
  in.release();
  out.release();
  if (!ok) set_err("resize failed: %s", cudaGetErrorString(cudaGetLastError()));
  return ok ? 0 : -1;
}

}  // extern "C"
