// kernels.cuh -- sm_100a kernels of the JDA detect path.
//
//   k1_resize    bilinear down-sample to the h (1/sqrt2) and q (1/2) planes     c/jda.c:203-230
//   k2_scan      stage-0 scan of every candidate window from integer LUTs       c/jda.c:332-402 (t = 0)
//   k3_cascade   generic per-window cascade + regression gather + emit          c/jda.c:356-427
//
// Arithmetic contract: every float expression the reference evaluates is
// evaluated here with the same operations in the same order, round-to-nearest, no FMA
// contraction (explicit __fadd_rn/__fmul_rn/__fsub_rn/__fdiv_rn and -fmad=false), float->int by
// truncation.  Tree traversal and leaf indices are integer-exact; scores and shapes come out
// bit-identical to the reference's C path.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "host_model.hpp"

namespace jda {

// ------------------------------------------------------------------------------ parameters

struct LevelInfo {
  int win, step;
  int nx, ny;           // windows per row / column at this level
  int tw_log2, th;      // windows per tile: (1 << tw_log2) x th
  int ntx, nty;         // tiles per frame
  int box_w, box_h;     // shared-memory tile in pixels (box_w is the tile pitch); global path: 0
  int use_smem;
  int span;             // warps whose tile buffers one tile spans (1, 2 or 4): only every span-th warp works
  int table_off;        // byte offset of this level's stage-0 table
  long long win_base;   // scan-order index of the level's first window inside a frame
};

#ifndef JDA_K2_WARPS
#define JDA_K2_WARPS 12
#endif
#ifndef JDA_K2_GLOBAL_WIDE
#define JDA_K2_GLOBAL_WIDE 0  /* 8-wide groups on the global-memory levels: measured slower (r1) */
#endif
#ifndef JDA_K2_TILE_BYTES
#define JDA_K2_TILE_BYTES 8192
#endif
constexpr int K2_WARPS = JDA_K2_WARPS;            // warps per scan block (one block per SM)
constexpr int K2_TILE_BYTES = JDA_K2_TILE_BYTES;  // per-warp pixel tile
constexpr int K2_LIST_CAP = 512;     // windows per tile (fits u16 ids)
constexpr int K2_MAX_SCHED = 32;

// Shared memory of a scan block:
//   [ stage-0 table | norm table (256 B) | K2_WARPS pixel tiles, contiguous | K2_WARPS WarpLists ]
// The tiles are contiguous so that a coarse level can give one warp the buffers of 2 or 4
// neighbours (LevelInfo::span) while those neighbours sit the level out.
struct __align__(128) WarpLists {
  float lscore[K2_LIST_CAP];
  uint16_t lwid[K2_LIST_CAP];
  unsigned long long mbar;
  unsigned parity;  // phase of `mbar` to wait for next (kept here so a group of warps can share one barrier)
  unsigned char pad[116];
};
static_assert(sizeof(WarpLists) % 128 == 0, "WarpLists must keep 128-byte alignment");
constexpr size_t K2_WARP_BYTES = K2_TILE_BYTES + sizeof(WarpLists);

struct ScanParams {
  CUtensorMap maps[kMaxLevels];  // one 3-D u8 map per level (box = that level's tile)
  LevelInfo lv[kMaxLevels];
  const uint8_t *frames;
  size_t frame_stride;
  int pitch, W, H, n_frames;
  int frame_base;                // index of frames[0] inside the caller's batch (chunked launches)
  const int2 *frame_dims;        // mixed-size batches (k2_scan<.., MIXED = true>): (width, height) of every frame of
                                 // the caller's batch inside its W x H canvas slot
  int n_levels, K, table_bytes;
  const uint8_t *tables;         // n_levels x table_bytes (padded to 128)
  const Stage0Norm *norms;       // kMaxNorm entries
  unsigned *tile_counters;       // [n_levels] work-stealing counters
  uint4 *surv;                   // stage-0 survivors: {frame, level<<26 | yi<<13 | xi, score bits, 0}
  unsigned *surv_count;
  unsigned surv_cap;
  uint8_t *surv_leaves;          // [surv_cap][leaf_pad]: the K stage-0 leaf indices of every survivor, two per byte
  int leaf_pad;
  long long windows_per_frame;
  int n_sched;
  short sched[K2_MAX_SCHED];     // cart index at which each phase ends; last == K
  int use_tma;
  int stragglers;                // 1: finish nearly empty tiles in cart-parallel straggler mode
  float level_cum[kMaxLevels];   // cumulative share of the scan work in processing order (coarse -> fine)
  // trace (TRACE instantiation only)
  int *trace_n;
  float *trace_s;
  uint8_t *trace_leaf;
  long long leaf_w0, leaf_w1;
  int leaf_stride;
};

struct CascadeParams {
  const uint8_t *frames;
  size_t frame_stride;
  int pitch, W, H;
  const uint8_t *hq;  // per frame: h plane (hw*hh) then q plane (qw*qh); NULL when no node needs them
  size_t hq_stride;
  int hw, hh, qw, qh;
  const NodeRec *nodes;
  const float *leaf;
  const float4 *cart;  // th, mean, std, -
  const float *w;
  const float *mean_shape;
  int T, K, L, t_run;
  int k_extra;         // > 0: after the t_run full stages, carts [0, k_extra) of stage t_run, no regression after them
                       // (Validate's unfinished stage, src/jda/cascador.cpp:199-209)
  int n_eval0;         // carts already evaluated for a queue entry when the kernel resumes it (trace)
  int depth, nn, nl;   // tree depth (c/jda.c:28 fixes 4), internal nodes and leaves per cart: read by k3_cascade<.., false>
  float r;             // 1.f / sqrtf(2.f), computed on the host like c/jda.c:341
  int n_levels;
  int lv_win[kMaxLevels], lv_step[kMaxLevels], lv_nx[kMaxLevels], lv_ny[kMaxLevels];
  long long lv_base[kMaxLevels];
  long long windows_per_frame;
  const uint4 *surv;
  const unsigned *surv_count;
  unsigned surv_cap;
  const float *init_shape;  // [surv_cap][2L] shapes after stage 0 (from k3_stage0) when t_start == 1
  int t_start;              // 0: start from the mean shape; 1: stage 0 already done for every queue entry
  int dense;           // 1: enumerate every window of every frame instead of reading `surv`
  long long dense_total;
  float *hits;         // records of rec_words 4-byte words
  unsigned *hit_count;
  unsigned hit_cap;
  unsigned *work_counter;  // queue mode: next survivor to take (dynamic distribution evens out deep survivors)
  int rec_words;
  float th;
  int use_th;
  int *trace_n;
  float *trace_s;
  uint8_t *trace_leaf;
  long long leaf_w0, leaf_w1;
  int leaf_stride;
};

constexpr int kHitHeader = 6;  // frame, key, x, y, win, score  -- then 2L shape floats

__host__ __device__ inline uint32_t pack_key(int level, int yi, int xi) {
  return ((uint32_t)level << 26) | ((uint32_t)yi << 13) | (uint32_t)xi;
}

#ifdef __CUDACC__

// ------------------------------------------------------------------------------ PTX helpers

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a tile load that never lands (bad descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 22)) __trap();
}
// 3-D tiled TMA load (x, y, frame) -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int z,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}

// ------------------------------------------------------------------------------ k1: resize

// One thread per output pixel; blockIdx.z = frame, blockIdx.y selects the h (0) or q (1) plane.
// Expression order of c/jda.c:214-226.
__global__ void k1_resize(const uint8_t *__restrict__ frames, size_t frame_stride, int pitch, int W, int H,
                          uint8_t *__restrict__ hq, size_t hq_stride, int hw, int hh, int qw, int qh) {
  const int plane = blockIdx.y;
  const int dw = plane ? qw : hw, dh = plane ? qh : hh;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= dw * dh) return;
  const uint8_t *src = frames + (size_t)blockIdx.z * frame_stride;
  uint8_t *dst = hq + (size_t)blockIdx.z * hq_stride + (plane ? (size_t)hw * hh : 0);
  const int i = o / dw, j = o - i * dw;
  const float xr = __fdiv_rn((float)(W - 1), (float)dw);
  const float yr = __fdiv_rn((float)(H - 1), (float)dh);
  const float fx = __fmul_rn(xr, (float)j), fy = __fmul_rn(yr, (float)i);
  const int x = __float2int_rz(fx), y = __float2int_rz(fy);
  const float xd = __fsub_rn(fx, (float)x), yd = __fsub_rn(fy, (float)y);
  const uint8_t *p = src + (size_t)y * pitch + x;
  const float a = (float)p[0], b = (float)p[1], c = (float)p[pitch], d = (float)p[pitch + 1];
  const float ix = __fsub_rn(1.f, xd), iy = __fsub_rn(1.f, yd);
  float v = __fmul_rn(__fmul_rn(a, ix), iy);
  v = __fadd_rn(v, __fmul_rn(__fmul_rn(b, xd), iy));
  v = __fadd_rn(v, __fmul_rn(__fmul_rn(c, ix), yd));
  v = __fadd_rn(v, __fmul_rn(__fmul_rn(d, xd), yd));
  dst[o] = (uint8_t)__float2int_rz(v);
}

// ------------------------------------------------------------------------------ k0: unpack
//
// Mixed-size batches arrive as one contiguous blob per frame (rows of a few hundred bytes make a pitched
// host->device copy crawl: one DMA descriptor per row); this kernel moves every frame from the packed staging
// area into its 16-byte-pitched canvas slot.  blockIdx.y = frame (relative to f0), blockIdx.x = group of 8 rows.
struct UnpackFrame {
  unsigned long long src_off;  // byte offset of the frame inside the packed staging area
  int width, height, src_pitch, pad;
};

__global__ void k0_unpack(const uint8_t *__restrict__ packed, const UnpackFrame *__restrict__ tab, int f0,
                          uint8_t *__restrict__ frames, size_t frame_stride, int pitch) {
  const UnpackFrame fr = tab[f0 + blockIdx.y];
  const uint8_t *src = packed + fr.src_off;
  uint8_t *dst = frames + (size_t)(f0 + blockIdx.y) * frame_stride;
  const int r0 = blockIdx.x * 8, r1 = min(r0 + 8, fr.height);
  for (int r = r0; r < r1; r++)
    for (int x = threadIdx.x; x < fr.width; x += blockDim.x) dst[(size_t)r * pitch + x] = src[(size_t)r * fr.src_pitch + x];
}

// ------------------------------------------------------------------------------ k2: stage-0 scan
//
// One persistent block per SM.  The block keeps the current level's stage-0 table (K x 96 B) in
// shared memory; its warps are independent workers that pull tiles of that level from a global
// counter.  A warp owns a private pixel tile (TMA-filled, box chosen per level so that tile +
// window list fit its scratch) and walks the carts in phases: inside a phase all lanes evaluate
// the same cart (uniform table reads), dead lanes are squeezed out between phases by
// __ballot_sync compaction into an in-place (window id, score) list.  Windows that survive all K
// carts are appended to the global survivor queue for k3_cascade.
//
// Levels whose windows are too large for a private tile read pixels straight from global memory
// (L1/L2) with the packed-coordinate table format; everything else is identical.

template <bool SMEM>
struct PixBase;
template <>
struct PixBase<true> {
  uint32_t off;  // byte offset into the block's shared memory
};
template <>
struct PixBase<false> {
  const uint8_t *ptr;
};

template <bool SMEM>
__device__ __forceinline__ int node_test(const uint8_t *smem, uint2 n, const PixBase<SMEM> &b, int pitch) {
  if constexpr (SMEM) {
    const int p1 = smem[b.off + (n.x & 0xffffu)];
    const int p2 = smem[b.off + (n.x >> 16)];
    return (p1 - p2 <= (int)n.y) ? 1 : 2;
  } else {
    const int x1 = n.x & 0x7ff, y1 = (n.x >> 11) & 0x7ff, th = (int)(n.x >> 22) - 256;
    const int x2 = n.y & 0x7ff, y2 = n.y >> 11;
    const int p1 = __ldg(b.ptr + y1 * pitch + x1);
    const int p2 = __ldg(b.ptr + y2 * pitch + x2);
    return (p1 - p2 <= th) ? 1 : 2;
  }
}

// Everything scan_tile's helpers need about the current tile.
struct TileCtx {
  const ScanParams *P;
  const LevelInfo *lv;
  uint8_t *smem;
  const Stage0Norm *norms;
  float *lscore;
  uint16_t *lwid;
  const uint8_t *gbase;  // global address of the tile's first pixel
  long long gw0;         // scan-order index of the level's first window in this frame (trace)
  uint32_t tile_off;     // shared-memory byte offset of the tile
  int tw_log2, tw_mask, step, pitch;
  int x0w, y0w, cw;
  int lane;
};

template <bool SMEM>
__device__ __forceinline__ PixBase<SMEM> window_base(const TileCtx &c, int wid, bool alive) {
  // dead lanes point at the tile origin: valid memory, one broadcast word
  const int wx = alive ? (wid & c.tw_mask) : 0, wy = alive ? (wid >> c.tw_log2) : 0;
  PixBase<SMEM> b;
  if constexpr (SMEM) b.off = c.tile_off + (uint32_t)(wy * c.step * c.pitch + wx * c.step);
  else b.ptr = c.gbase + (size_t)(wy * c.step) * c.P->pitch + wx * c.step;
  return b;
}

__device__ __forceinline__ long long trace_index(const TileCtx &c, int wid) {
  return c.gw0 + (long long)(c.y0w + (wid >> c.tw_log2)) * c.lv->nx + c.x0w + (wid & c.tw_mask);
}

// One group of 32*NWG list entries through carts [cart, cend): every lane walks the same cart
// (uniform table reads), NWG windows per lane interleaved level by level for ILP.  Survivors are
// squeezed to the front of the in-place list at `out` (writes never pass the read cursor).
template <bool SMEM, int NWG, bool TRACE>
__device__ __forceinline__ void scan_group(const TileCtx &c, int ph, int base, int n, int cart, int cend,
                                           int &out) {
  const uint8_t *smem = c.smem;
  const int lane = c.lane;
  float score[NWG];
  int wid[NWG];
  bool alive[NWG];
  PixBase<SMEM> pb[NWG];
#pragma unroll
  for (int j = 0; j < NWG; j++) {
    const int e = base + j * 32 + lane;
    if (ph == 0) {
      wid[j] = e;
      alive[j] = (e < n) && ((e & c.tw_mask) < c.cw);
      score[j] = 0.f;
    } else {
      alive[j] = e < n;
      wid[j] = alive[j] ? (int)c.lwid[e] : 0;
      score[j] = alive[j] ? c.lscore[e] : 0.f;
    }
    pb[j] = window_base<SMEM>(c, wid[j], alive[j]);
  }
  for (int k = cart; k < cend; ++k) {
    bool any = false;
#pragma unroll
    for (int j = 0; j < NWG; j++) any |= alive[j];
    if (!__any_sync(0xffffffffu, any)) break;
    const uint32_t co = (uint32_t)k * kCartBytes;
    const uint2 n0 = *reinterpret_cast<const uint2 *>(smem + co);
    const uint2 tf = *reinterpret_cast<const uint2 *>(smem + co + 88);  // cart threshold | norm-table index + 1: one load
    const float cth = __uint_as_float(tf.x);
    const uint32_t nflag = tf.y;
    int idx[NWG];
    float s[NWG];
#pragma unroll
    for (int j = 0; j < NWG; j++) idx[j] = node_test<SMEM>(smem, n0, pb[j], c.pitch);
#pragma unroll
    for (int j = 0; j < NWG; j++) {
      const uint2 nd = *reinterpret_cast<const uint2 *>(smem + co + idx[j] * 8);
      idx[j] = 2 * idx[j] + node_test<SMEM>(smem, nd, pb[j], c.pitch);
    }
#pragma unroll
    for (int j = 0; j < NWG; j++) {
      const uint2 nd = *reinterpret_cast<const uint2 *>(smem + co + idx[j] * 8);
      idx[j] = 2 * idx[j] + node_test<SMEM>(smem, nd, pb[j], c.pitch);
    }
    // leaf = idx - 7; leaf scores start at byte 56 of the cart record
#pragma unroll
    for (int j = 0; j < NWG; j++)
      s[j] = __fadd_rn(score[j], *reinterpret_cast<const float *>(smem + co + (56 - 4 * kNodes) + 4 * idx[j]));
    if (nflag) {  // warp-uniform and rare: only carts with (mean, std) != (0, 1)
      const Stage0Norm nm = c.norms[nflag - 1];
#pragma unroll
      for (int j = 0; j < NWG; j++) s[j] = __fdiv_rn(__fsub_rn(s[j], nm.mean), nm.std);
    }
    if constexpr (TRACE) {
      const ScanParams &P = *c.P;
#pragma unroll
      for (int j = 0; j < NWG; j++) {
        if (alive[j]) {
          const long long gw = trace_index(c, wid[j]);
          if (P.trace_leaf && gw >= P.leaf_w0 && gw < P.leaf_w1)
            P.trace_leaf[(size_t)(gw - P.leaf_w0) * P.leaf_stride + k] = (uint8_t)(idx[j] - kNodes);
          if (s[j] < cth) {
            if (P.trace_n) P.trace_n[gw] = k + 1;
            if (P.trace_s) P.trace_s[gw] = s[j];
          }
        }
      }
    }
    // c/jda.c:399 -- branch-free: a rejected lane keeps walking its (valid) window until the phase ends,
    // its score is never read again
#pragma unroll
    for (int j = 0; j < NWG; j++) {
      score[j] = s[j];
      alive[j] = alive[j] && !(s[j] < cth);
    }
  }
  __syncwarp();  // every lane has read its list entries of this group before any lane overwrites one (in-place list)
#pragma unroll
  for (int j = 0; j < NWG; j++) {
    const unsigned m = __ballot_sync(0xffffffffu, alive[j]);
    if (alive[j]) {
      const int pos = out + __popc(m & ((1u << lane) - 1u));
      c.lwid[pos] = (uint16_t)wid[j];
      c.lscore[pos] = score[j];
    }
    out += __popc(m);
  }
}

// Straggler mode.  Once a tile is down to n <= K2_STRAG_MAX windows the parallel axis flips: for each remaining window
// the 32 lanes walk 32 CONSECUTIVE CARTS (the shape is the constant mean shape in stage 0, so the carts of a window are
// independent), park the 32 leaf scores in shared memory, and then the running scores are replayed cart by cart with
// lane = window -- the same sequential adds / compares as the reference, so reject cart and score bits are unchanged.
// A short list keeps a window-parallel warp waiting on one chain of dependent shared-memory reads per cart with most
// lanes idle; here every lane works and the chains of several windows overlap.  Round 2 measurements (512 mix frames):
// threshold 8 -> 24.6 ms, 15 -> 23.2, 23 -> 22.55, 32 -> 22.50, 48 -> 22.67, 64 -> 23.19 (profiles/r2p_ab.txt, r2q_ab.txt), hence batches: the list may be longer than the
// leaf-score scratch has rows and is walked chunk by chunk (32 carts), K2_STRAG_ROWS windows at a time, survivors
// compacted in place.  On return the list holds the windows that passed every cart; n is updated.
#ifndef JDA_K2_STRAG_MAX
#define JDA_K2_STRAG_MAX 32
#endif
constexpr int K2_STRAG_MAX = JDA_K2_STRAG_MAX;  // <= 64: the moved list (below) has 64 slots
constexpr int K2_STRAG_ROWS = 20;               // windows per batch = rows of leaf scores
constexpr int K2_LS_STRIDE = 33;                // conflict-free both for the lane = cart writes and lane = window reads
// scratch = the tile's own lists, WarpLists::lscore[512] followed by ::lwid[512], viewed as 768 floats:
//   [0, 660) leaf scores [row][33]   [672, 736) scores of the list   [736, 768) window ids of the list (64 x u16)
static_assert(K2_STRAG_ROWS * K2_LS_STRIDE <= 672 && K2_STRAG_MAX <= 64 && K2_LIST_CAP == 512, "straggler scratch layout");

template <bool SMEM, bool TRACE>
__device__ __forceinline__ void straggler_tail(const TileCtx &c, int &n, int cart) {
  const ScanParams &P = *c.P;
  const uint8_t *smem = c.smem;
  const int lane = c.lane, K = P.K;
  float *ls = c.lscore;
  float *sc = c.lscore + 672;
  uint16_t *wd = reinterpret_cast<uint16_t *>(c.lscore + 736);
  {  // move the list (<= 64 entries at the front of lscore / lwid) out of the leaf-score rows' way
    const bool v0 = lane < n, v1 = lane + 32 < n;
    const int a0 = v0 ? (int)c.lwid[lane] : 0, a1 = v1 ? (int)c.lwid[lane + 32] : 0;
    const float s0 = v0 ? c.lscore[lane] : 0.f, s1 = v1 ? c.lscore[lane + 32] : 0.f;
    __syncwarp();
    if (v0) { wd[lane] = (uint16_t)a0; sc[lane] = s0; }
    if (v1) { wd[lane + 32] = (uint16_t)a1; sc[lane + 32] = s1; }
    __syncwarp();
  }
  const int lrow = min(lane, K2_STRAG_ROWS - 1) * K2_LS_STRIDE;
  for (int k0 = cart; k0 < K && n > 0; k0 += 32) {
    const int k = min(k0 + lane, K - 1);
    const uint32_t co = (uint32_t)k * kCartBytes;
    const uint2 n0 = *reinterpret_cast<const uint2 *>(smem + co);
    const int cnt = min(32, K - k0);
    int out = 0;
    for (int b0 = 0; b0 < n; b0 += K2_STRAG_ROWS) {
      const int nb = min(K2_STRAG_ROWS, n - b0);
      bool alive = lane < nb;
      const int wid = alive ? (int)wd[b0 + lane] : 0;
      float score = alive ? sc[b0 + lane] : 0.f;
      [[maybe_unused]] uint8_t lf[K2_STRAG_ROWS];
      // two windows per step of the walk: the tree walk is a chain of dependent shared-memory reads
      for (int w = 0; w < nb; w += 2) {
        const int wb = min(w + 1, nb - 1);  // odd count: the last window walks twice (same values, same slots)
        const PixBase<SMEM> pa = window_base<SMEM>(c, __shfl_sync(0xffffffffu, wid, w), true);
        const PixBase<SMEM> pb = window_base<SMEM>(c, __shfl_sync(0xffffffffu, wid, wb), true);
        int ia = node_test<SMEM>(smem, n0, pa, c.pitch), ib = node_test<SMEM>(smem, n0, pb, c.pitch);
        uint2 na = *reinterpret_cast<const uint2 *>(smem + co + ia * 8), nb2 = *reinterpret_cast<const uint2 *>(smem + co + ib * 8);
        ia = 2 * ia + node_test<SMEM>(smem, na, pa, c.pitch);
        ib = 2 * ib + node_test<SMEM>(smem, nb2, pb, c.pitch);
        na = *reinterpret_cast<const uint2 *>(smem + co + ia * 8);
        nb2 = *reinterpret_cast<const uint2 *>(smem + co + ib * 8);
        ia = 2 * ia + node_test<SMEM>(smem, na, pa, c.pitch) - kNodes;
        ib = 2 * ib + node_test<SMEM>(smem, nb2, pb, c.pitch) - kNodes;
        ls[w * K2_LS_STRIDE + lane] = *reinterpret_cast<const float *>(smem + co + 56 + 4 * ia);
        ls[wb * K2_LS_STRIDE + lane] = *reinterpret_cast<const float *>(smem + co + 56 + 4 * ib);
        if constexpr (TRACE) {
#pragma unroll
          for (int q = 0; q < K2_STRAG_ROWS; q++) {
            if (q == w) lf[q] = (uint8_t)ia;
            if (q == wb) lf[q] = (uint8_t)ib;
          }
        }
      }
      __syncwarp();
      [[maybe_unused]] int died_at = alive ? cnt : -1;  // chunk-local cart after which this lane's window stopped
      // the score walk, eight carts per step of the loop: their leaf scores and thresholds are fetched together, only
      // the adds and compares form a chain (c/jda.c:395-401, same order)
      for (int j0 = 0; j0 < cnt; j0 += 8) {
        float v[8], th8[8];
        uint32_t nf = 0;
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const uint2 tf = *reinterpret_cast<const uint2 *>(smem + (uint32_t)min(k0 + j0 + u, K - 1) * kCartBytes + 88);
          th8[u] = __uint_as_float(tf.x);
          nf |= tf.y;
          v[u] = ls[lrow + j0 + u];  // lanes past the batch re-read its last row (in bounds)
        }
        if (nf == 0u) {  // (uniform) no cart of the eight has a real (mean, std)
#pragma unroll
          for (int u = 0; u < 8; u++) {
            if (j0 + u < cnt) {
              const float s = __fadd_rn(score, v[u]);
              if (alive) {
                score = s;
                if (s < th8[u]) {  // c/jda.c:399
                  alive = false;
                  if constexpr (TRACE) {
                    died_at = j0 + u;
                    const long long gw = trace_index(c, wid);
                    if (P.trace_n) P.trace_n[gw] = k0 + j0 + u + 1;
                    if (P.trace_s) P.trace_s[gw] = s;
                  }
                }
              }
            }
          }
        } else {
          for (int u = 0; u < 8 && j0 + u < cnt; u++) {
            const uint32_t nflag = *reinterpret_cast<const uint32_t *>(smem + (uint32_t)(k0 + j0 + u) * kCartBytes + 92);
            float s = __fadd_rn(score, v[u]);
            if (nflag) {
              const Stage0Norm nm = c.norms[nflag - 1];
              s = __fdiv_rn(__fsub_rn(s, nm.mean), nm.std);
            }
            if (alive) {
              score = s;
              if (s < th8[u]) {
                alive = false;
                if constexpr (TRACE) {
                  died_at = j0 + u;
                  const long long gw = trace_index(c, wid);
                  if (P.trace_n) P.trace_n[gw] = k0 + j0 + u + 1;
                  if (P.trace_s) P.trace_s[gw] = s;
                }
              }
            }
          }
        }
        if (!__any_sync(0xffffffffu, alive)) break;
      }
      if constexpr (TRACE) {
        // leaves of the carts the reference would have evaluated: up to and including the rejecting one
        if (P.trace_leaf) {
#pragma unroll
          for (int q = 0; q < K2_STRAG_ROWS; q++) {
            if (q < nb) {
              const int dw = __shfl_sync(0xffffffffu, died_at, q);
              const long long gw = trace_index(c, __shfl_sync(0xffffffffu, wid, q));
              if (gw >= P.leaf_w0 && gw < P.leaf_w1 && k0 + lane < K && lane <= min(dw, cnt - 1))
                P.trace_leaf[(size_t)(gw - P.leaf_w0) * P.leaf_stride + k0 + lane] = lf[q];
            }
          }
        }
      }
      __syncwarp();  // everyone is done with this batch's list entries and leaf scores
      const unsigned m = __ballot_sync(0xffffffffu, alive);
      if (alive) {
        const int pos = out + __popc(m & ((1u << lane) - 1u));
        wd[pos] = (uint16_t)wid;
        sc[pos] = score;
      }
      out += __popc(m);
      __syncwarp();
    }
    n = out;
  }
  {  // the windows that passed every cart, back at the front of the tile's lists
    const bool v0 = lane < n, v1 = lane + 32 < n;
    const int a0 = v0 ? (int)wd[lane] : 0, a1 = v1 ? (int)wd[lane + 32] : 0;
    const float s0 = v0 ? sc[lane] : 0.f, s1 = v1 ? sc[lane + 32] : 0.f;
    __syncwarp();
    if (v0) { c.lwid[lane] = (uint16_t)a0; c.lscore[lane] = s0; }
    if (v1) { c.lwid[lane + 32] = (uint16_t)a1; c.lscore[lane + 32] = s1; }
    __syncwarp();
  }
}

// Windows of tile (x0w, y0w) that exist in `frame`: the level's own nx x ny grid, or -- in a mixed-size batch,
// where tiles are enumerated over the canvas -- the grid of the frame's own width and height (c/jda.c:320-339 run
// on that frame alone).  false = no window of this tile exists in the frame.
template <bool MIXED>
__device__ __forceinline__ bool tile_extent(const ScanParams &P, const LevelInfo &lv, int frame, int x0w, int y0w,
                                            int &cw, int &ch) {
  int nx = lv.nx, ny = lv.ny;
  if constexpr (MIXED) {
    const int2 d = __ldg(P.frame_dims + frame + P.frame_base);
    if (d.x < lv.win || d.y < lv.win) return false;
    nx = (d.x - lv.win) / lv.step + 1;
    ny = (d.y - lv.win) / lv.step + 1;
  }
  cw = min(1 << lv.tw_log2, nx - x0w);
  ch = min(lv.th, ny - y0w);
  return cw > 0 && ch > 0;
}

template <bool SMEM, int NW, bool TRACE>
__device__ __forceinline__ void scan_tile(const ScanParams &P, const LevelInfo &lv, int li, uint8_t *smem,
                                          uint32_t norm_off, uint32_t tile_off, float *lscore,
                                          uint16_t *lwid, int frame, int x0w, int y0w, int cw, int ch,
                                          int lane) {
  TileCtx c;
  c.P = &P; c.lv = &lv; c.smem = smem;
  c.norms = reinterpret_cast<const Stage0Norm *>(smem + norm_off);
  c.lscore = lscore; c.lwid = lwid;
  c.tw_log2 = lv.tw_log2; c.tw_mask = (1 << lv.tw_log2) - 1; c.step = lv.step;
  c.pitch = SMEM ? lv.box_w : P.pitch;
  c.gbase = P.frames + (size_t)frame * P.frame_stride + (size_t)(y0w * lv.step) * P.pitch + (size_t)x0w * lv.step;
  c.gw0 = (long long)(frame + P.frame_base) * P.windows_per_frame + lv.win_base;
  c.tile_off = tile_off; c.x0w = x0w; c.y0w = y0w; c.cw = cw; c.lane = lane;

  int n = ch << c.tw_log2;  // dense enumeration; columns >= cw are masked off in phase 0
  int cart = 0;
  for (int ph = 0; ph < P.n_sched; ++ph) {
    const int cend = P.sched[ph];
    int out = 0, base = 0;
    // full groups of NWM packets, then the remainder with as few packets as it needs.  The global-memory
    // levels run twice as wide: their pixel reads are L2-latency bound, not shared-memory bound.
    constexpr int NWM = (SMEM || TRACE || NW > 4 || !JDA_K2_GLOBAL_WIDE) ? NW : 2 * NW;
    for (; n - base >= 32 * NWM; base += 32 * NWM) scan_group<SMEM, NWM, TRACE>(c, ph, base, n, cart, cend, out);
    if constexpr (NWM >= 8) {
      if (n - base > 128) { scan_group<SMEM, 8, TRACE>(c, ph, base, n, cart, cend, out); base = n; }
    }
    if constexpr (NWM >= 4) {
      if (n - base > 64) { scan_group<SMEM, 4, TRACE>(c, ph, base, n, cart, cend, out); base = n; }
    }
    if constexpr (NWM >= 2) {
      if (n - base > 32) { scan_group<SMEM, 2, TRACE>(c, ph, base, n, cart, cend, out); base = n; }
    }
    if (n - base > 0) scan_group<SMEM, 1, TRACE>(c, ph, base, n, cart, cend, out);
    __syncwarp();
    n = out;
    cart = cend;
    if (n == 0) break;
    if (n <= K2_STRAG_MAX && cart < P.K && P.stragglers) {
      straggler_tail<SMEM, TRACE>(c, n, cart);
      break;
    }
  }
  // windows that passed every cart of stage 0: queue them for the cascade kernels together with their K leaf
  // indices.  The phases did not keep the leaves, so each survivor (a handful per tile) is re-walked here, one
  // cart per lane, while its pixels and the table are still in shared memory -- ~17 dense steps per survivor,
  // instead of 9 x K scattered global loads later in k3_stage0.
  for (int base = 0; base < n; base += 32) {
    const int e = base + lane;
    const bool v = e < n;
    const unsigned m = __ballot_sync(0xffffffffu, v);
    unsigned slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(P.surv_count, (unsigned)__popc(m));
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    const unsigned slot = slot0 + __popc(m & ((1u << lane) - 1u));
    const int w = v ? (int)lwid[e] : 0;
    if (v && slot < P.surv_cap)
      P.surv[slot] = make_uint4((unsigned)(frame + P.frame_base), pack_key(li, y0w + (w >> c.tw_log2), x0w + (w & c.tw_mask)),
                                __float_as_uint(lscore[e]), 0u);
    for (unsigned rest = m; rest; rest &= rest - 1) {
      const int src = __ffs(rest) - 1;
      const unsigned sl = __shfl_sync(0xffffffffu, slot, src);
      const PixBase<SMEM> pb = window_base<SMEM>(c, __shfl_sync(0xffffffffu, w, src), true);
      if (sl >= P.surv_cap || !P.surv_leaves) continue;  // no leaf store: nothing regresses from this stage (truncated cascade)
      uint8_t *out = P.surv_leaves + (size_t)sl * P.leaf_pad;
      for (int k0 = 0; k0 < P.K; k0 += 32) {
        const int k = min(k0 + lane, P.K - 1);
        const uint32_t co = (uint32_t)k * kCartBytes;
        int idx = node_test<SMEM>(smem, *reinterpret_cast<const uint2 *>(smem + co), pb, c.pitch);
        idx = 2 * idx + node_test<SMEM>(smem, *reinterpret_cast<const uint2 *>(smem + co + idx * 8), pb, c.pitch);
        idx = 2 * idx + node_test<SMEM>(smem, *reinterpret_cast<const uint2 *>(smem + co + idx * 8), pb, c.pitch);
        // two 3-bit leaves per byte (cart k in the low nibble of byte k / 2): half the bytes of round 1's records
        const int lf = (k0 + lane < P.K) ? idx - kNodes : 0;
        const int hi = __shfl_down_sync(0xffffffffu, lf, 1);
        if (!(lane & 1) && k0 + lane < P.K) out[(k0 + lane) >> 1] = (uint8_t)(lf | (hi << 4));
      }
    }
  }
  __syncwarp();
}

template <int NW, bool TRACE, bool MIXED = false>
__global__ void __launch_bounds__(K2_WARPS * 32, 1) k2_scan(const __grid_constant__ ScanParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ int s_skip;
  __shared__ unsigned s_item[K2_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t table_sz = (uint32_t)(P.table_bytes + 127) & ~127u;
  const uint32_t norm_off = table_sz;
  const uint32_t tile_off = table_sz + 256u + (uint32_t)warp * K2_TILE_BYTES;
  WarpLists *wl = reinterpret_cast<WarpLists *>(smem + table_sz + 256u + (uint32_t)K2_WARPS * K2_TILE_BYTES) + warp;
  uint8_t *tile = smem + tile_off;
  const uint32_t bar = smem_u32(&wl->mbar);
  const uint32_t tile_s = smem_u32(tile);
  if (lane == 0) { mbar_init(bar, 1); wl->parity = 0; }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (int i = threadIdx.x; i < kMaxNorm * 2; i += blockDim.x)
    reinterpret_cast<float *>(smem + norm_off)[i] = reinterpret_cast<const float *>(P.norms)[i];

  // Levels are visited coarse -> fine, but every block starts at the level where its share of the total
  // work begins (and wraps around), so most blocks load one or two tables instead of all of them and the
  // block-wide barrier at a level switch is paid far less often.  Levels whose tiles are all taken are
  // skipped without loading their table.
  int start = 0;
  {
    const float f = ((float)blockIdx.x + 0.5f) / (float)gridDim.x;
    while (start + 1 < P.n_levels && P.level_cum[start] < f) start++;
  }
  for (int it = 0; it < P.n_levels; ++it) {
    int pos = start + it;
    if (pos >= P.n_levels) pos -= P.n_levels;
    const int li = P.n_levels - 1 - pos;
    const LevelInfo &lv = P.lv[li];
    const int tiles_per_frame = lv.ntx * lv.nty;
    const unsigned total = (unsigned)tiles_per_frame * (unsigned)P.n_frames;
    __syncthreads();  // everyone is done with the previous level's table
    if (threadIdx.x == 0) s_skip = *reinterpret_cast<volatile unsigned *>(&P.tile_counters[li]) >= total;
    __syncthreads();
    if (s_skip) continue;
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(P.tables + lv.table_off);
      uint4 *dst = reinterpret_cast<uint4 *>(smem);
      for (int i = threadIdx.x; i < (int)(table_sz / 16); i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int tw = 1 << lv.tw_log2;
    if (lv.span > 1) {
      // Coarse level: `span` neighbouring warps pool their tile buffers into one tile, load it once (the
      // group's first warp issues the TMA on its barrier) and split its window rows between them.
      const int gl = warp % lv.span, grp = warp / lv.span;
      WarpLists *lead = wl - gl;
      const uint32_t gbar = smem_u32(&lead->mbar);
      const uint32_t gtile_off = tile_off - (uint32_t)gl * K2_TILE_BYTES;
      const uint32_t gtile_s = tile_s - (uint32_t)gl * K2_TILE_BYTES;
      const int nthr = lv.span * 32;
      bool loaded = false;
      for (;;) {
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(nthr) : "memory");  // group is done with the previous tile
        int frame = 0, x0w = 0, y0w = 0, cw = 0, ch = 0;
        if (gl == 0) {
          if (lane == 0) {
            if (loaded && P.use_tma) lead->parity ^= 1u;
            const unsigned it2 = atomicAdd(&P.tile_counters[li], 1u);
            s_item[grp] = it2;
            if (it2 < total && P.use_tma) {
              const int f = it2 / tiles_per_frame, r = it2 - f * tiles_per_frame;
              const int ty = r / lv.ntx, tx = r - ty * lv.ntx;
              if (tile_extent<MIXED>(P, lv, f, tx * tw, ty * lv.th, cw, ch)) {
                mbar_expect_tx(gbar, (uint32_t)(lv.box_w * lv.box_h));
                tma_load_3d(gtile_s, &P.maps[li], (tx * tw * lv.step) & ~15, ty * lv.th * lv.step, f, gbar);
              }
            }
          }
          __syncwarp();
        }
        asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(nthr) : "memory");  // tile id published
        const unsigned item = s_item[grp];
        if (item >= total) break;
        frame = item / tiles_per_frame;
        {
          const int r = item - frame * tiles_per_frame;
          const int ty = r / lv.ntx, tx = r - ty * lv.ntx;
          x0w = tx * tw; y0w = ty * lv.th;
        }
        loaded = tile_extent<MIXED>(P, lv, frame, x0w, y0w, cw, ch);  // same answer in every thread of the group
        if (!loaded) continue;                                  // mixed-size batch: the tile lies outside this frame
        const int px0 = (x0w * lv.step) & ~15, py0 = y0w * lv.step;
        const int xs = x0w * lv.step - px0;
        if (P.use_tma) {
          mbar_wait(gbar, lead->parity);
        } else {
          uint8_t *gt = smem + gtile_off;
          const uint8_t *src = P.frames + (size_t)frame * P.frame_stride;
          for (int i = gl * 32 + lane; i < lv.box_w * lv.box_h; i += nthr) {
            const int yy = i / lv.box_w, xx = i - yy * lv.box_w;
            const int gx = px0 + xx, gy = py0 + yy;
            gt[i] = (gx < P.W && gy < P.H) ? src[(size_t)gy * P.pitch + gx] : (uint8_t)0;
          }
          asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(nthr) : "memory");
        }
        const int r0 = ch * gl / lv.span, r1 = ch * (gl + 1) / lv.span;
        if (r1 > r0)
          scan_tile<true, NW, TRACE>(P, lv, li, smem, norm_off, gtile_off + (uint32_t)(xs + r0 * lv.step * lv.box_w),
                                     wl->lscore, wl->lwid, frame, x0w, y0w + r0, cw, r1 - r0, lane);
      }
      continue;
    }
    for (;;) {
      unsigned item = 0;
      if (lane == 0) item = atomicAdd(&P.tile_counters[li], 1u);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= total) break;
      const int frame = item / tiles_per_frame;
      const int r = item - frame * tiles_per_frame;
      const int ty = r / lv.ntx, tx = r - ty * lv.ntx;
      const int x0w = tx * tw, y0w = ty * lv.th;
      int cw, ch;
      if (!tile_extent<MIXED>(P, lv, frame, x0w, y0w, cw, ch)) continue;  // mixed-size batch: outside this frame
      if (lv.use_smem) {
        // box origin: x rounded down to 16 bytes (TMA alignment), the windows sit `xs` bytes into the tile
        const int px0 = (x0w * lv.step) & ~15, py0 = y0w * lv.step;
        const int xs = x0w * lv.step - px0;
        if (P.use_tma) {
          if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t)(lv.box_w * lv.box_h));
            tma_load_3d(tile_s, &P.maps[li], px0, py0, frame, bar);
          }
          mbar_wait(bar, wl->parity);
          __syncwarp();
          if (lane == 0) wl->parity ^= 1u;
          __syncwarp();
        } else {
          const uint8_t *src = P.frames + (size_t)frame * P.frame_stride;
          for (int i = lane; i < lv.box_w * lv.box_h; i += 32) {
            const int yy = i / lv.box_w, xx = i - yy * lv.box_w;
            const int gx = px0 + xx, gy = py0 + yy;
            tile[i] = (gx < P.W && gy < P.H) ? src[(size_t)gy * P.pitch + gx] : (uint8_t)0;
          }
          __syncwarp();
        }
        scan_tile<true, NW, TRACE>(P, lv, li, smem, norm_off, tile_off + (uint32_t)xs, wl->lscore, wl->lwid, frame, x0w,
                                   y0w, cw, ch, lane);
      } else {
        scan_tile<false, NW, TRACE>(P, lv, li, smem, norm_off, tile_off, wl->lscore, wl->lwid, frame, x0w,
                                    y0w, cw, ch, lane);
      }
    }
  }
}

// ------------------------------------------------------------------------------ k3: generic cascade
//
// One warp per window.  Within a stage the shape is fixed, so the K tree walks are independent:
// lanes take 32 consecutive carts at a time (float address arithmetic in the reference's exact
// order), then the running score is replayed over those 32 carts in cart order (sequential adds
// and the (score - mean) / std normalisation, early exit at the first score < th), so the reject
// decision, the exit score and the cart count equal the reference's.  After a completed stage
// the regression is a leaf-index gather: lane i sums row (8k + leaf_k) of w[t] into shape[i]
// for k = 0..K-1 in ascending order -- 2L independent chains, each in the reference's order.

#ifndef JDA_K3_MIN_BLOCKS
#define JDA_K3_MIN_BLOCKS 8  /* blocks per SM the register allocation must allow: the kernel waits on L2, so warps in flight
                                count for more than registers (r2k A/B: 1 block / 4 chunks 3.55 ms, 8 / 2 chunks 2.93 ms) */
#endif
#ifndef JDA_K3_G
#define JDA_K3_G 2
#endif
constexpr int K3_WARPS = 4;
constexpr int K3_G = JDA_K3_G;  // chunks of 32 carts walked together per survivor

// D4 = true: depth-4 carts (7 nodes, 8 leaves) as compile-time constants -- the shipped model and the only depth the
// reference's C path can load (c/jda.c:24-32).  D4 = false: depth 2..6 from the model header.
template <bool TRACE, bool D4 = true>
__global__ void __launch_bounds__(K3_WARPS * 32, JDA_K3_MIN_BLOCKS) k3_cascade(const __grid_constant__ CascadeParams P) {
  extern __shared__ __align__(16) uint8_t smem3[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = 2 * P.L;
  const int NN = D4 ? kNodes : P.nn, NL = D4 ? kLeaves : P.nl, DL = D4 ? kDepth - 1 : P.depth - 1;
  const int per_warp = kMaxDim * 4 + ((P.K + 15) & ~15);
  float *shape = reinterpret_cast<float *>(smem3 + (size_t)warp * per_warp);
  uint8_t *leafs = reinterpret_cast<uint8_t *>(shape + kMaxDim);

  const long long total = P.dense ? P.dense_total : (long long)min(*P.surv_count, P.surv_cap);
  long long e = (long long)blockIdx.x * K3_WARPS + warp - (long long)gridDim.x * K3_WARPS;
  for (;;) {
    if (P.dense) {
      e += (long long)gridDim.x * K3_WARPS;
    } else {
      unsigned t = 0;
      if (lane == 0) t = atomicAdd(P.work_counter, 1u);
      e = (long long)__shfl_sync(0xffffffffu, t, 0);
    }
    if (e >= total) break;
    int frame, level, xi, yi;
    float score0 = 0.f;
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
      score0 = __uint_as_float(s.z);
      level = (int)(s.y >> 26);
      yi = (int)((s.y >> 13) & 0x1fff);
      xi = (int)(s.y & 0x1fff);
    }
    const int win = P.lv_win[level], step = P.lv_step[level];
    const int x = xi * step, y = yi * step;
    const float fwin = (float)win;
    // view origins, c/jda.c:344-354
    const int hx = __float2int_rz(__fmul_rn((float)x, P.r)), hy = __float2int_rz(__fmul_rn((float)y, P.r));
    const int qx = x / 2, qy = y / 2;
    const uint8_t *po = P.frames + (size_t)frame * P.frame_stride;
    const uint8_t *ph = P.hq ? P.hq + (size_t)frame * P.hq_stride : nullptr;
    const uint8_t *pq = P.hq ? ph + (size_t)P.hw * P.hh : nullptr;
    const long long gw = (long long)frame * P.windows_per_frame + P.lv_base[level] + (long long)yi * P.lv_nx[level] + xi;
    const bool trace_leaf = TRACE && P.trace_leaf && gw >= P.leaf_w0 && gw < P.leaf_w1;

    __syncwarp();
    const bool resumed = !P.dense && P.t_start > 0;
    for (int i = lane; i < D; i += 32)
      shape[i] = (resumed && P.init_shape) ? P.init_shape[(size_t)e * D + i] : P.mean_shape[i];
    __syncwarp();

    float score = resumed ? score0 : 0.f;
    int n_eval = resumed ? P.n_eval0 : 0;
    bool rejected = false;
    const int t_end = P.t_run + (P.k_extra > 0 ? 1 : 0);
    for (int t = resumed ? P.t_start : 0; t < t_end && !rejected; t++) {
      const int Kt = t < P.t_run ? P.K : P.k_extra;  // the unfinished stage of a truncated cascade stops after k_extra carts
      // K3_G chunks of 32 carts are walked together (independent load chains: the tree walk is a string of
      // dependent L2 accesses), then their scores are replayed chunk by chunk with early exit
      for (int kc = 0; kc < Kt && !rejected; kc += 32 * K3_G) {
        float ls[K3_G];
        float4 cp[K3_G];
        int idx[K3_G];
        const NodeRec *nd[K3_G];
        bool ok[K3_G];
#pragma unroll
        for (int g = 0; g < K3_G; g++) {
          const int k = kc + 32 * g + lane;
          ok[g] = k < Kt;
          nd[g] = P.nodes + ((size_t)t * P.K + (ok[g] ? k : 0)) * NN;
          idx[g] = 0;
          ls[g] = 0.f;
          cp[g] = make_float4(0.f, 0.f, 1.f, 0.f);
        }
#pragma unroll
        for (int lvl = 0; lvl < DL; lvl++) {
          int4 a[K3_G];
          float4 o[K3_G];
#pragma unroll
          for (int g = 0; g < K3_G; g++) {
            a[g] = __ldg(reinterpret_cast<const int4 *>(nd[g] + idx[g]));
            o[g] = __ldg(reinterpret_cast<const float4 *>(nd[g] + idx[g]) + 1);
          }
          int p1[K3_G], p2[K3_G];
#pragma unroll
          for (int g = 0; g < K3_G; g++) {
            // a = scale, lm1, lm2, th ; o = o1x, o1y, o2x, o2y   (c/jda.c:371-389)
            const float x1 = __fadd_rn(shape[a[g].y], o[g].x), y1 = __fadd_rn(shape[a[g].y + 1], o[g].y);
            const float x2 = __fadd_rn(shape[a[g].z], o[g].z), y2 = __fadd_rn(shape[a[g].z + 1], o[g].w);
            int x1_ = __float2int_rz(__fmul_rn(x1, fwin)), y1_ = __float2int_rz(__fmul_rn(y1, fwin));
            int x2_ = __float2int_rz(__fmul_rn(x2, fwin)), y2_ = __float2int_rz(__fmul_rn(y2, fwin));
            x1_ = min(max(x1_, 0), win - 1); y1_ = min(max(y1_, 0), win - 1);
            x2_ = min(max(x2_, 0), win - 1); y2_ = min(max(y2_, 0), win - 1);
            if (a[g].x == 0) {
              p1[g] = __ldg(po + (size_t)(y + y1_) * P.pitch + x + x1_);
              p2[g] = __ldg(po + (size_t)(y + y2_) * P.pitch + x + x2_);
            } else {
              // h / q views keep w = win (c/jda.c:347,352); linear index like the reference, reads
              // past the plane buffer (undefined there) are defined as 0 here
              const uint8_t *pp = (a[g].x == 1) ? ph : pq;
              const int pw = (a[g].x == 1) ? P.hw : P.qw, phh = (a[g].x == 1) ? P.hh : P.qh;
              const int bx = (a[g].x == 1) ? hx : qx, by = (a[g].x == 1) ? hy : qy;
              const long long lim = (long long)pw * phh;
              const long long i1 = (long long)(by + y1_) * pw + bx + x1_;
              const long long i2 = (long long)(by + y2_) * pw + bx + x2_;
              p1[g] = (i1 < lim) ? (int)__ldg(pp + i1) : 0;
              p2[g] = (i2 < lim) ? (int)__ldg(pp + i2) : 0;
            }
          }
#pragma unroll
          for (int g = 0; g < K3_G; g++) idx[g] = (p1[g] - p2[g] <= a[g].w) ? 2 * idx[g] + 1 : 2 * idx[g] + 2;
        }
#pragma unroll
        for (int g = 0; g < K3_G; g++) {
          if (ok[g]) {
            const int k = kc + 32 * g + lane;
            const size_t c = (size_t)t * P.K + k;
            const int lf = idx[g] - NN;
            leafs[k] = (uint8_t)lf;
            ls[g] = __ldg(P.leaf + c * NL + lf);
            cp[g] = __ldg(P.cart + c);
          }
        }
        // replay the score in cart order, one chunk of 32 at a time
#pragma unroll
        for (int g = 0; g < K3_G; g++) {
          const int cnt = min(32, Kt - (kc + 32 * g));
          if (cnt <= 0 || rejected) continue;
          const unsigned normed = __ballot_sync(0xffffffffu, cp[g].y != 0.f || cp[g].z != 1.f);  // carts with a real (mean, std)
          int stop = -1;
          for (int j = 0; j < cnt; j++) {
            const float sj = __shfl_sync(0xffffffffu, ls[g], j);
            const float thj = __shfl_sync(0xffffffffu, cp[g].x, j);
            score = __fadd_rn(score, sj);                      // c/jda.c:396
            if ((normed >> j) & 1u) {                          // c/jda.c:397; (score - 0) / 1 is score exactly
              const float mj = __shfl_sync(0xffffffffu, cp[g].y, j);
              const float dj = __shfl_sync(0xffffffffu, cp[g].z, j);
              score = __fdiv_rn(__fsub_rn(score, mj), dj);
            }
            n_eval++;
            if (score < thj) { stop = j; break; }              // c/jda.c:399
          }
          if (TRACE && trace_leaf && ok[g] && (stop < 0 || lane <= stop))
            P.trace_leaf[(size_t)(gw - P.leaf_w0) * P.leaf_stride + (size_t)t * P.K + kc + 32 * g + lane] = leafs[kc + 32 * g + lane];
          if (stop >= 0) rejected = true;
        }
      }
      if (rejected || t >= P.t_run) break;  // no regression after the carts of an unfinished stage (cascador.cpp:199-209)
      __syncwarp();
      // global regression, c/jda.c:403-411
      const float *wt = P.w + (size_t)t * P.K * NL * D;
      for (int i0 = 0; i0 < D; i0 += 64) {
        // two coordinates per lane per sweep over the K rows (independent chains, each k ascending).  The 16
        // row loads of a batch are issued together before their adds: the gather is a string of L2 reads.
        const bool on0 = i0 + lane < D, on1 = i0 + 32 + lane < D;
        float acc0 = on0 ? shape[i0 + lane] : 0.f, acc1 = on1 ? shape[i0 + 32 + lane] : 0.f;
        for (int k0 = 0; k0 < P.K; k0 += 16) {
          float v0[16], v1[16];
#pragma unroll
          for (int u = 0; u < 16; u++) {
            const int k = min(k0 + u, P.K - 1);
            const float *row = wt + (size_t)(k * NL + leafs[k]) * D + i0 + lane;
            v0[u] = on0 ? __ldg(row) : 0.f;
            v1[u] = on1 ? __ldg(row + 32) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 16; u++) {
            if (k0 + u < P.K) {
              acc0 = __fadd_rn(acc0, v0[u]);
              acc1 = __fadd_rn(acc1, v1[u]);
            }
          }
        }
        if (on0) shape[i0 + lane] = acc0;
        if (on1) shape[i0 + 32 + lane] = acc1;
      }
      __syncwarp();
    }
    if (TRACE) {
      if (lane == 0) {
        if (P.trace_n) P.trace_n[gw] = n_eval;
        if (P.trace_s) P.trace_s[gw] = score;
      }
    }
    if (rejected) continue;
    if (P.use_th && score < P.th) continue;  // c/jda.c:414
    unsigned slot = 0;
    if (lane == 0) slot = atomicAdd(P.hit_count, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot < P.hit_cap) {
      float *rec = P.hits + (size_t)slot * P.rec_words;
      if (lane == 0) {
        reinterpret_cast<int *>(rec)[0] = frame;
        reinterpret_cast<uint32_t *>(rec)[1] = pack_key(level, yi, xi);
        reinterpret_cast<int *>(rec)[2] = x;
        reinterpret_cast<int *>(rec)[3] = y;
        reinterpret_cast<int *>(rec)[4] = win;
        rec[5] = score;
      }
      for (int i = lane; i < D; i += 32) rec[kHitHeader + i] = shape[i];
    }
  }
}


// ------------------------------------------------------------------------------ k3_stage0
//
// Stage 0 for the survivors of k2_scan: they are known to pass every cart of the stage and k2 has
// their exit score, so only two things are left -- the K leaf indices and the regression gather
//   shape = mean_shape + sum_k w[0][8k + leaf_k]        (k ascending, c/jda.c:403-411).
// The leaves were recorded by k2_scan (a re-walk of each survivor while its tile was in shared memory).
// The gather is the expensive part: 8K rows x 2L floats stream from L2 per survivor if done naively
// (that made k3_cascade L2-bound in profiles/r1_v1).  Here a block takes a cohort of 32 survivors and
// stages w[0] through shared memory 8 carts (64 rows) at a time with cp.async double buffering, so
// each byte of w[0] is fetched once per cohort instead of once per survivor.
constexpr int K3S_WARPS = 8;
constexpr int K3S_PER_WARP = 4;
constexpr int K3S_COHORT = K3S_WARPS * K3S_PER_WARP;
constexpr int K3S_CHUNK = 4;  // carts staged per step (6.9 KB x 2 buffers: the block fits on an SM next to a scan block)
static_assert(K3S_CHUNK % 4 == 0, "leaf indices are fetched four at a time");
// bytes of one survivor's stage-0 leaf record: two leaves per byte, padded to 16
__host__ __device__ constexpr int leaf_bytes(int K) { return (((K + 1) >> 1) + 15) & ~15; }

struct Stage0Params {
  const uint8_t *surv_leaves;    // [surv_cap][leaf_bytes(K)] leaf indices written by k2_scan (k3_walk for t >= 1), two per byte
  const float *w0;               // w[t]: [8K][2L]
  const float *mean_shape;
  int K, L;
  const unsigned *surv_count;    // entries to process: of the survivor queue, or of `list`
  unsigned surv_cap;
  float *out_shape;              // [surv_cap][2L]
  // stages >= 1 (kernels_stages.cuh): the windows that passed stage t are named by a list of queue entries and the
  // regression continues from their shape after stage t - 1, in place (c/jda.c:403-411: shape[i] += w[...], k ascending)
  const uint2 *list;             // {queue entry, -}; NULL: entry i is queue entry i
  const float *in_shape;         // [surv_cap][2L]; NULL: the mean shape
};

__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// HALVES = ceil(2L / 64): coordinate pairs per lane.  L <= 32 (the shipped L = 27) needs one, so the second,
// fully predicated-off pass over every row is compiled out (it was 20 % of the kernel's instructions).
template <int HALVES>
__global__ void __launch_bounds__(K3S_WARPS * 32) k3_stage0(const __grid_constant__ Stage0Params P) {
  extern __shared__ __align__(16) uint8_t smem0[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = 2 * P.L, K = P.K;
  const int kpad = leaf_bytes(K);
  const int chunk_floats = K3S_CHUNK * kLeaves * D;
  uint8_t *leaves = smem0;                                                   // [cohort][kpad]
  float *rows = reinterpret_cast<float *>(smem0 + (size_t)K3S_COHORT * kpad);  // [2][chunk_floats]
  const int total = (int)min(*P.surv_count, P.surv_cap);
  const int n_chunks = (K + K3S_CHUNK - 1) / K3S_CHUNK;

  for (int c0 = blockIdx.x * K3S_COHORT; c0 < total; c0 += gridDim.x * K3S_COHORT) {
    // ---- leaves of my K3S_PER_WARP survivors: computed by k2_scan while the window was in shared memory
    unsigned ent[K3S_PER_WARP];  // queue entries of my seats
#pragma unroll
    for (int s = 0; s < K3S_PER_WARP; s++) {
      const int i = c0 + warp * K3S_PER_WARP + s;
      ent[s] = (i < total) ? (P.list ? P.list[i].x : (unsigned)i) : 0u;
    }
#pragma unroll
    for (int s = 0; s < K3S_PER_WARP; s++) {
      if (c0 + warp * K3S_PER_WARP + s >= total) continue;
      const uint32_t *src = reinterpret_cast<const uint32_t *>(P.surv_leaves + (size_t)ent[s] * kpad);
      uint32_t *dst = reinterpret_cast<uint32_t *>(leaves + (size_t)(warp * K3S_PER_WARP + s) * kpad);
      for (int i = lane; i < kpad / 4; i += 32) dst[i] = __ldg(src + i);
    }
    // ---- regression gather over staged chunks of w[0]
    float2 acc[K3S_PER_WARP][HALVES];
#pragma unroll
    for (int s = 0; s < K3S_PER_WARP; s++)
#pragma unroll
      for (int h = 0; h < HALVES; h++) {
        const int p = lane + 32 * h;
        const float *from = (P.in_shape && c0 + warp * K3S_PER_WARP + s < total) ? P.in_shape + (size_t)ent[s] * D : P.mean_shape;
        acc[s][h] = (2 * p + 1 < D) ? make_float2(from[2 * p], from[2 * p + 1]) : make_float2(0.f, 0.f);
      }
    auto stage = [&](int ci, int buf) {
      const int carts = min(K3S_CHUNK, K - ci * K3S_CHUNK);
      const int n16 = carts * kLeaves * D / 4;  // 16-byte pieces
      const float4 *src = reinterpret_cast<const float4 *>(P.w0 + (size_t)ci * chunk_floats);
      float4 *dst = reinterpret_cast<float4 *>(rows + (size_t)buf * chunk_floats);
      for (int i = threadIdx.x; i < n16; i += blockDim.x) cp_async16(dst + i, src + i);
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, 0);
    for (int ci = 0; ci < n_chunks; ci++) {
      if (ci + 1 < n_chunks) {
        stage(ci + 1, (ci + 1) & 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();  // chunk ci landed for everyone (and, first time round, the leaves are written)
      const float *rb = rows + (size_t)(ci & 1) * chunk_floats;
      const int carts = min(K3S_CHUNK, K - ci * K3S_CHUNK);
      const int k0 = ci * K3S_CHUNK;
#pragma unroll
      for (int s = 0; s < K3S_PER_WARP; s++) {
        if (c0 + warp * K3S_PER_WARP + s >= total) continue;  // no survivor in this seat: its leaves are stale
        // the chunk's leaf indices of this survivor, four per 16-bit load (two per byte; k0 is a multiple of 4)
        const uint16_t *lp = reinterpret_cast<const uint16_t *>(leaves + (warp * K3S_PER_WARP + s) * kpad + (k0 >> 1));
        uint32_t lq[K3S_CHUNK / 4];
#pragma unroll
        for (int q = 0; q < K3S_CHUNK / 4; q++) lq[q] = lp[q];
#pragma unroll
        for (int cl = 0; cl < K3S_CHUNK; cl++) {
          if (cl < carts) {
            const uint32_t leaf = (lq[cl >> 2] >> (4 * (cl & 3))) & 0xfu;
            const float2 *row = reinterpret_cast<const float2 *>(rb + (cl * kLeaves + leaf) * D);
#pragma unroll
            for (int h = 0; h < HALVES; h++) {
              const int p = lane + 32 * h;
              if (2 * p + 1 < D) {
                const float2 v = row[p];
                acc[s][h].x = __fadd_rn(acc[s][h].x, v.x);
                acc[s][h].y = __fadd_rn(acc[s][h].y, v.y);
              }
            }
          }
        }
      }
      __syncthreads();  // everyone is done with buffer ci & 1 before it is refilled
    }
#pragma unroll
    for (int s = 0; s < K3S_PER_WARP; s++) {
      if (c0 + warp * K3S_PER_WARP + s < total) {
#pragma unroll
        for (int h = 0; h < HALVES; h++) {
          const int p = lane + 32 * h;
          if (2 * p + 1 < D) reinterpret_cast<float2 *>(P.out_shape + (size_t)ent[s] * D)[p] = acc[s][h];
        }
      }
    }
  }
}

#endif  // __CUDACC__
}  // namespace jda
