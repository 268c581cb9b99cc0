// kernels_stages.cuh -- stages >= 1 for the scan's survivors, one stage at a time.
//
//   k3_walk      carts of ONE stage for a list of survivors (c/jda.c:364-402, t >= 1)
//   k3_emit      final threshold + hit records (c/jda.c:413-427)
// (the regression between two stages, c/jda.c:403-411, is k3_stage0 run on the list of the windows that passed)
//
// k3_cascade (kernels.cuh) gives every survivor a warp of its own for all of its stages: the node records of its carts
// come from L1 / L2 with one line per lane (lane = cart), and the running score is replayed by 32 lanes doing the same
// scalar work (47 % of that kernel's instructions, profiles/r2l_full_hotspots.txt).  Here the survivors of a batch are
// taken through the cascade stage by stage instead:
//   * the stage's node table sits in shared memory (one block per SM; K x 7 x 20 B, offsets and packed integers in
//     separate arrays: both are bank-conflict-free for lane = cart reads), so only the pixels and the survivor's shape
//     come from L1 / L2;
//   * a warp owns a batch of up to 32 survivors.  Per chunk of 32 carts it walks the batch's live survivors one pair
//     at a time with lane = cart (within a stage the shape is fixed, so the carts are independent), parks the leaf
//     scores in shared memory, and replays the running scores with lane = SURVIVOR: the same sequential adds,
//     normalisations and compares as the reference (c/jda.c:395-401), once per batch instead of once per survivor;
//   * windows that pass the stage are appended to the next stage's list together with their score; their leaf indices
//     go to the survivor leaf records (two per byte, k2_scan's format), from which k3_stage0 applies the regression.
// Reject cart, exit score, leaf indices and shapes are bit-identical to k3_cascade's (same float operations in the same
// order); the per-window trace is produced by both.
#pragma once
#include "kernels.cuh"

namespace jda {

constexpr int K3W_MAX_WARPS = 24;
constexpr int K3W_LS_STRIDE = 33;  // leaf scores [survivor][cart]: conflict-free for lane = cart writes and lane = survivor reads
constexpr int kMaxStages = 32;     // T <= 32 (load_model)

struct WalkParams {
  CascadeParams C;          // frames, planes, model, level geometry, survivor queue, hit queue, trace (as k3_cascade reads them)
  int t;                    // the stage
  int Kt;                   // carts of it to evaluate: K, or k_extra for the unfinished stage of a truncated cascade
  int n_eval_before;        // carts the reference has evaluated for a window when it enters this stage (trace)
  const uint2 *in_list;     // {queue entry, score bits} of the windows entering the stage; NULL: every entry of C.surv
  const unsigned *in_count;
  uint2 *out_list;          // windows that passed the Kt carts
  unsigned *out_count;
  unsigned *work;           // next list position to hand out
  const float *shape;       // [surv_cap][2L] shapes after stage t - 1
  uint8_t *leaves;          // [surv_cap][leaf_pad] leaf indices of this stage, two per byte; NULL: no regression follows
  int leaf_pad;
};

__host__ __device__ inline size_t k3w_table_bytes(int K) {
  // offs (float4) | packed (u32) | leaf scores | cart thresholds | normalisation bit mask
  return (size_t)K * kNodes * 16 + (size_t)K * kNodes * 4 + (size_t)K * kLeaves * 4 + (size_t)K * 4 + (size_t)((K + 31) / 32) * 4;
}
// per warp: leaf scores [32][33]; the traced instantiation also parks the chunk's leaf indices [32][32]
__host__ __device__ inline size_t k3w_warp_bytes(bool trace) { return (size_t)32 * K3W_LS_STRIDE * 4 + (trace ? 1024 : 0); }

#ifdef __CUDACC__

// scale | lm1 << 2 | lm2 << 10 | (th + 256) << 18: lm < 256 (2L <= kMaxDim = 128); the pixel difference lies in
// [-255, 255], so clamping th to [-256, 255] keeps every `feature <= th` (as in the stage-0 tables)
__device__ __forceinline__ uint32_t pack_node(const NodeRec &n) {
  const int th = min(max(n.th, -256), 255);
  return (uint32_t)(n.scale & 3) | ((uint32_t)n.lm1 << 2) | ((uint32_t)n.lm2 << 10) | ((uint32_t)(th + 256) << 18);
}

struct WalkWindow {
  const uint8_t *po;  // the window's first pixel in the frame
  const float *shape;
  int frame, x, y, win;
};

// One tree level of one cart for one window: the node test of c/jda.c:369-394 (float ops in the reference's order).
template <bool HQ>
__device__ __forceinline__ int walk_level(const WalkParams &P, const float4 *offs, const uint32_t *packed, int node,
                                          const WalkWindow &w) {
  const CascadeParams &C = P.C;
  const uint32_t pk = packed[node];
  const float4 o = offs[node];
  const int lm1 = (pk >> 2) & 0xff, lm2 = (pk >> 10) & 0xff, th = (int)(pk >> 18) - 256;
  const float2 s1 = __ldg(reinterpret_cast<const float2 *>(w.shape + lm1));
  const float2 s2 = __ldg(reinterpret_cast<const float2 *>(w.shape + lm2));
  const float fwin = (float)w.win;
  const float x1 = __fadd_rn(s1.x, o.x), y1 = __fadd_rn(s1.y, o.y);
  const float x2 = __fadd_rn(s2.x, o.z), y2 = __fadd_rn(s2.y, o.w);
  int x1_ = __float2int_rz(__fmul_rn(x1, fwin)), y1_ = __float2int_rz(__fmul_rn(y1, fwin));
  int x2_ = __float2int_rz(__fmul_rn(x2, fwin)), y2_ = __float2int_rz(__fmul_rn(y2, fwin));
  x1_ = min(max(x1_, 0), w.win - 1); y1_ = min(max(y1_, 0), w.win - 1);
  x2_ = min(max(x2_, 0), w.win - 1); y2_ = min(max(y2_, 0), w.win - 1);
  int p1, p2;
  const int scale = HQ ? (int)(pk & 3u) : 0;
  if (!HQ || scale == 0) {
    p1 = __ldg(w.po + (size_t)y1_ * C.pitch + x1_);
    p2 = __ldg(w.po + (size_t)y2_ * C.pitch + x2_);
  } else {
    // h / q views keep w = win (c/jda.c:347,352); linear index like the reference, reads past the plane buffer
    // (undefined there) are defined as 0 here -- as in k3_cascade
    const uint8_t *ph = C.hq + (size_t)w.frame * C.hq_stride;
    const uint8_t *pp = (scale == 1) ? ph : ph + (size_t)C.hw * C.hh;
    const int pw = (scale == 1) ? C.hw : C.qw, phh = (scale == 1) ? C.hh : C.qh;
    const int bx = (scale == 1) ? __float2int_rz(__fmul_rn((float)w.x, C.r)) : w.x / 2;
    const int by = (scale == 1) ? __float2int_rz(__fmul_rn((float)w.y, C.r)) : w.y / 2;
    const long long lim = (long long)pw * phh;
    const long long i1 = (long long)(by + y1_) * pw + bx + x1_;
    const long long i2 = (long long)(by + y2_) * pw + bx + x2_;
    p1 = (i1 < lim) ? (int)__ldg(pp + i1) : 0;
    p2 = (i2 < lim) ? (int)__ldg(pp + i2) : 0;
  }
  return (p1 - p2 <= th) ? 1 : 2;
}

template <bool TRACE, bool HQ>
__global__ void __launch_bounds__(K3W_MAX_WARPS * 32, 1) k3_walk(const __grid_constant__ WalkParams P) {
  extern __shared__ __align__(16) uint8_t smemw[];
  const CascadeParams &C = P.C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int K = C.K, Kt = P.Kt, D = 2 * C.L;
  float4 *offs = reinterpret_cast<float4 *>(smemw);
  uint32_t *packed = reinterpret_cast<uint32_t *>(offs + (size_t)K * kNodes);
  float *leaf = reinterpret_cast<float *>(packed + (size_t)K * kNodes);
  float *cth = leaf + (size_t)K * kLeaves;
  uint32_t *nmask = reinterpret_cast<uint32_t *>(cth + K);
  float *ls = reinterpret_cast<float *>(smemw + ((k3w_table_bytes(K) + 15) & ~(size_t)15) + (size_t)warp * k3w_warp_bytes(TRACE));
  [[maybe_unused]] uint8_t *lfs = reinterpret_cast<uint8_t *>(ls + 32 * K3W_LS_STRIDE);

  const unsigned total = P.in_list ? min(*P.in_count, C.surv_cap) : min(*C.surv_count, C.surv_cap);
  if (total == 0) return;
  {  // the stage's tables (only the carts this launch evaluates)
    const NodeRec *nodes = C.nodes + (size_t)P.t * K * kNodes;
    const float *lf = C.leaf + (size_t)P.t * K * kLeaves;
    const float4 *cart = C.cart + (size_t)P.t * K;
    for (int i = threadIdx.x; i < Kt * kNodes; i += blockDim.x) {
      const NodeRec n = nodes[i];
      offs[i] = make_float4(n.o1x, n.o1y, n.o2x, n.o2y);
      packed[i] = pack_node(n);
    }
    for (int i = threadIdx.x; i < Kt * kLeaves; i += blockDim.x) leaf[i] = lf[i];
    for (int i = threadIdx.x; i < (Kt + 31) / 32; i += blockDim.x) nmask[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < Kt; i += blockDim.x) {
      const float4 cp = cart[i];
      cth[i] = cp.x;
      if (cp.y != 0.f || cp.z != 1.f) atomicOr(&nmask[i >> 5], 1u << (i & 31));  // (score - 0) / 1 is score exactly
    }
    __syncthreads();
  }
  // batch size: up to 32 survivors per warp, but at least four batches per warp of the grid (batches are handed out
  // dynamically; with one or two per warp the slowest warp decides the kernel's time), and a short list is spread
  // over more warps
  const unsigned all_warps = gridDim.x * nwarps;
  const unsigned B = min(32u, max(4u, (total + 4 * all_warps - 1) / (4 * all_warps)));
  const float4 *cart_g = C.cart + (size_t)P.t * K;

  for (;;) {
    unsigned first = 0;
    if (lane == 0) first = atomicAdd(P.work, B);
    first = __shfl_sync(0xffffffffu, first, 0);
    if (first >= total) break;
    const int nb = (int)min(B, total - first);
    // lane = survivor of the batch
    bool alive = lane < nb;
    unsigned e = 0;
    float score = 0.f;
    if (alive) {
      if (P.in_list) {
        const uint2 q = P.in_list[first + lane];
        e = q.x;
        score = __uint_as_float(q.y);
      } else {
        e = first + lane;
      }
    }
    const uint4 sv = alive ? C.surv[e] : make_uint4(0u, 0u, 0u, 0u);
    if (alive && !P.in_list) score = __uint_as_float(sv.z);
    const int level = (int)(sv.y >> 26), yi = (int)((sv.y >> 13) & 0x1fff), xi = (int)(sv.y & 0x1fff);
    const int win = C.lv_win[level], step = C.lv_step[level];
    const int wx = xi * step, wy = yi * step;
    const unsigned long long po = (unsigned long long)(C.frames + (size_t)sv.x * C.frame_stride + (size_t)wy * C.pitch + wx);
    [[maybe_unused]] long long gw = 0;
    if constexpr (TRACE) gw = (long long)sv.x * C.windows_per_frame + C.lv_base[level] + (long long)yi * C.lv_nx[level] + xi;

    for (int kc = 0; kc < Kt; kc += 32) {
      const unsigned live = __ballot_sync(0xffffffffu, alive);
      if (!live) break;
      const int cnt = min(32, Kt - kc);
      const int k = min(kc + lane, Kt - 1);  // lanes past the stage's last cart walk it again, their result is not used
      // ---- walk: lane = cart, two live survivors of the batch at a time (independent chains of dependent loads)
      for (unsigned rest = live; rest;) {
        const int s0 = __ffs(rest) - 1;
        rest &= rest - 1;
        const int s1 = rest ? __ffs(rest) - 1 : s0;  // odd count: the last survivor walks twice (same values, same slots)
        rest &= rest - 1;
        WalkWindow w0, w1;
        const unsigned e0 = __shfl_sync(0xffffffffu, e, s0), e1 = __shfl_sync(0xffffffffu, e, s1);
        w0.po = reinterpret_cast<const uint8_t *>(__shfl_sync(0xffffffffu, po, s0));
        w1.po = reinterpret_cast<const uint8_t *>(__shfl_sync(0xffffffffu, po, s1));
        w0.win = __shfl_sync(0xffffffffu, win, s0); w1.win = __shfl_sync(0xffffffffu, win, s1);
        w0.shape = P.shape + (size_t)e0 * D; w1.shape = P.shape + (size_t)e1 * D;
        if constexpr (HQ) {
          w0.frame = (int)__shfl_sync(0xffffffffu, sv.x, s0); w1.frame = (int)__shfl_sync(0xffffffffu, sv.x, s1);
          w0.x = __shfl_sync(0xffffffffu, wx, s0); w1.x = __shfl_sync(0xffffffffu, wx, s1);
          w0.y = __shfl_sync(0xffffffffu, wy, s0); w1.y = __shfl_sync(0xffffffffu, wy, s1);
        } else {
          w0.frame = w1.frame = 0; w0.x = w1.x = 0; w0.y = w1.y = 0;
        }
        int i0 = 0, i1 = 0;
#pragma unroll
        for (int lvl = 0; lvl < kDepth - 1; lvl++) {
          const int a = walk_level<HQ>(P, offs, packed, k * kNodes + i0, w0);
          const int b = walk_level<HQ>(P, offs, packed, k * kNodes + i1, w1);
          i0 = 2 * i0 + a;
          i1 = 2 * i1 + b;
        }
        i0 -= kNodes; i1 -= kNodes;
        ls[s0 * K3W_LS_STRIDE + lane] = leaf[k * kLeaves + i0];
        ls[s1 * K3W_LS_STRIDE + lane] = leaf[k * kLeaves + i1];
        if (P.leaves) {  // two 3-bit leaves per byte, cart k in the low nibble of byte k / 2 (k2_scan's record format)
          const int h0 = __shfl_down_sync(0xffffffffu, i0, 1), h1 = __shfl_down_sync(0xffffffffu, i1, 1);
          if (!(lane & 1) && kc + lane < Kt) {
            const int hi0 = (kc + lane + 1 < Kt) ? h0 : 0, hi1 = (kc + lane + 1 < Kt) ? h1 : 0;
            P.leaves[(size_t)e0 * P.leaf_pad + ((kc + lane) >> 1)] = (uint8_t)(i0 | (hi0 << 4));
            P.leaves[(size_t)e1 * P.leaf_pad + ((kc + lane) >> 1)] = (uint8_t)(i1 | (hi1 << 4));
          }
        }
        if constexpr (TRACE) {
          lfs[s0 * 32 + lane] = (uint8_t)i0;
          lfs[s1 * 32 + lane] = (uint8_t)i1;
        }
      }
      __syncwarp();
      // ---- replay: lane = survivor, carts in order (c/jda.c:395-401), eight per step of the loop
      const float *row = ls + min(lane, nb - 1) * K3W_LS_STRIDE;
      [[maybe_unused]] int stop = alive ? cnt : -1;  // chunk-local cart at which this lane's window was rejected (cnt: none)
      const bool was_alive = alive;
      for (int j0 = 0; j0 < cnt; j0 += 8) {
        float v[8], th8[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          v[u] = row[min(j0 + u, 31)];
          th8[u] = cth[min(kc + j0 + u, Kt - 1)];
        }
        const int kb = kc + j0;  // multiple of 8: the eight mask bits lie in one word
        const uint32_t nm8 = (nmask[kb >> 5] >> (kb & 31)) & 0xffu;
#pragma unroll
        for (int u = 0; u < 8; u++) {
          if (j0 + u < cnt) {
            float s = __fadd_rn(score, v[u]);                      // c/jda.c:396
            if ((nm8 >> u) & 1u) {                                 // (uniform, rare) c/jda.c:397
              const float4 cp = __ldg(cart_g + kb + u);
              s = __fdiv_rn(__fsub_rn(s, cp.y), cp.z);
            }
            if (alive) {
              score = s;
              if (s < th8[u]) {                                    // c/jda.c:399
                alive = false;
                if constexpr (TRACE) stop = j0 + u;
              }
            }
          }
        }
        if (!__any_sync(0xffffffffu, alive)) break;
      }
      if constexpr (TRACE) {
        if (was_alive && !alive) {
          if (C.trace_n) C.trace_n[gw] = P.n_eval_before + kc + stop + 1;
          if (C.trace_s) C.trace_s[gw] = score;
        }
        if (C.trace_leaf) {  // leaves of the carts the reference evaluates: up to and including the rejecting one
          for (unsigned rest = live; rest; rest &= rest - 1) {
            const int sv_ = __ffs(rest) - 1;
            const int st = __shfl_sync(0xffffffffu, stop, sv_);
            const long long g = __shfl_sync(0xffffffffu, gw, sv_);
            if (g >= C.leaf_w0 && g < C.leaf_w1 && lane <= min(st, cnt - 1))
              C.trace_leaf[(size_t)(g - C.leaf_w0) * C.leaf_stride + (size_t)P.t * K + kc + lane] = lfs[sv_ * 32 + lane];
          }
        }
      }
      __syncwarp();  // everyone has read this chunk's leaf scores before the next chunk overwrites them
    }
    // ---- the windows that passed every cart of the stage
    const unsigned pm = __ballot_sync(0xffffffffu, alive);
    if (pm) {
      unsigned slot0 = 0;
      if (lane == 0) slot0 = atomicAdd(P.out_count, (unsigned)__popc(pm));
      slot0 = __shfl_sync(0xffffffffu, slot0, 0);
      if (alive) P.out_list[slot0 + __popc(pm & ((1u << lane) - 1u))] = make_uint2(e, __float_as_uint(score));
    }
  }
}

// Final threshold and hit records for the windows that passed every stage (c/jda.c:413-427).  One warp per window.
template <bool TRACE>
__global__ void __launch_bounds__(128) k3_emit(const __grid_constant__ WalkParams P) {
  const CascadeParams &C = P.C;
  const int lane = threadIdx.x & 31;
  const int D = 2 * C.L;
  const unsigned total = P.in_list ? min(*P.in_count, C.surv_cap) : min(*C.surv_count, C.surv_cap);
  const unsigned nw = gridDim.x * (blockDim.x >> 5);
  for (unsigned i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < total; i += nw) {
    unsigned e = i;
    float score;
    if (P.in_list) {
      const uint2 q = P.in_list[i];
      e = q.x;
      score = __uint_as_float(q.y);
    } else {
      score = __uint_as_float(C.surv[i].z);
    }
    const uint4 sv = C.surv[e];
    const int level = (int)(sv.y >> 26), yi = (int)((sv.y >> 13) & 0x1fff), xi = (int)(sv.y & 0x1fff);
    const int win = C.lv_win[level], step = C.lv_step[level];
    if constexpr (TRACE) {
      const long long gw = (long long)sv.x * C.windows_per_frame + C.lv_base[level] + (long long)yi * C.lv_nx[level] + xi;
      if (lane == 0) {
        if (C.trace_n) C.trace_n[gw] = P.n_eval_before;
        if (C.trace_s) C.trace_s[gw] = score;
      }
    }
    if (C.use_th && score < C.th) continue;  // c/jda.c:414
    unsigned slot = 0;
    if (lane == 0) slot = atomicAdd(C.hit_count, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot < C.hit_cap) {
      float *rec = C.hits + (size_t)slot * C.rec_words;
      if (lane == 0) {
        reinterpret_cast<int *>(rec)[0] = (int)sv.x;
        reinterpret_cast<uint32_t *>(rec)[1] = sv.y;
        reinterpret_cast<int *>(rec)[2] = xi * step;
        reinterpret_cast<int *>(rec)[3] = yi * step;
        reinterpret_cast<int *>(rec)[4] = win;
        rec[5] = score;
      }
      const float *shape = P.shape ? P.shape + (size_t)e * D : C.mean_shape;
      for (int j = lane; j < D; j += 32) rec[kHitHeader + j] = shape[j];
    }
  }
}

#endif  // __CUDACC__
}  // namespace jda
