// kernels_regress.cuh -- k3_regress: the global regression of one stage for a list of windows
//   shape[e] = shape_in[e] + sum_k w[t][8k + leaf_k(e)]        (k ascending, c/jda.c:403-411)
//
// Second version of k3_stage0 (kernels.cuh), same cohort staging -- a block takes a cohort of windows and streams
// w[t] through shared memory once per cohort, double-buffered cp.async -- with the inner loop rebuilt around its
// instruction count (k3_stage0 retired 15 warp instructions per window-cart, issue slots 83 % busy):
//   * rows are padded to a multiple of four floats in a device copy of w (16-byte aligned rows), a lane owns four
//     coordinates, and -- when 2L <= 64, i.e. 16 lanes cover a row -- the two half-warps serve two different windows
//     with one instruction stream: one 16-byte shared-memory load + four adds per window-cart and lane;
//   * eight carts per staged chunk (half the barriers and cp.async issue overhead per cart), their eight leaf
//     indices come out of one 32-bit load (two per byte, k2_scan's / k3_walk's record format);
//   * 64 windows per cohort (each byte of w[t] crosses L2 -> shared memory once per 64 windows), fewer when the list
//     is short so that it still spreads over the SMs.
// The adds of one coordinate are the same adds in the same order (k ascending, starting from the incoming shape), so
// the shapes are bit-identical to k3_stage0's and to the reference's.
#pragma once
#include "kernels.cuh"

namespace jda {

constexpr int K3R_WARPS = 8;
constexpr int K3R_CHUNK = 8;        // carts staged per step
constexpr int K3R_MAX_SEATS = 8;    // windows per warp (as 4 half-warp pairs when a row fits 16 lanes)
__host__ __device__ constexpr int k3r_dpad(int D) { return (D + 3) & ~3; }
__host__ __device__ inline size_t k3r_smem_bytes(int K, int D) {
  return (size_t)K3R_WARPS * K3R_MAX_SEATS * leaf_bytes(K) + (size_t)2 * K3R_CHUNK * kLeaves * k3r_dpad(D) * 4;
}

struct RegressParams {
  const uint8_t *leaves;         // [cap][leaf_bytes(K)] leaf indices of the stage, two per byte
  const float *wp;               // w[t] with rows padded to k3r_dpad(2L) floats: [8K][dpad]
  const float *mean_shape;
  int K, L;
  const unsigned *count;         // entries to process: of the survivor queue, or of `list`
  unsigned cap;
  const uint2 *list;             // {queue entry, -}; NULL: entry i is queue entry i
  const float *in_shape;         // [cap][2L]; NULL: the mean shape
  float *out_shape;              // [cap][2L] (may be in_shape)
};

#ifdef __CUDACC__

// PAIRED: 2L <= 64 -- lanes 0..15 and 16..31 serve two different windows (seat 2p and 2p + 1 of the warp).
template <bool PAIRED>
__global__ void __launch_bounds__(K3R_WARPS * 32, 4) k3_regress(const __grid_constant__ RegressParams P) {
  extern __shared__ __align__(16) uint8_t smemr[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int D = 2 * P.L, Dp = k3r_dpad(D), K = P.K;
  const int kpad = leaf_bytes(K);
  const int chunk_floats = K3R_CHUNK * kLeaves * Dp;
  uint8_t *leaves = smemr;                                                                     // [warps * seats][kpad]
  float *rows = reinterpret_cast<float *>(smemr + (size_t)K3R_WARPS * K3R_MAX_SEATS * kpad);   // [2][chunk_floats]
  const int total = (int)min(*P.count, P.cap);
  if (total == 0) return;
  const int n_chunks = (K + K3R_CHUNK - 1) / K3R_CHUNK;
  // windows per warp: 8 when the list keeps every block busy with full cohorts, fewer for short lists
  int seats = K3R_MAX_SEATS;
  while (seats > (PAIRED ? 2 : 1) && (long long)gridDim.x * K3R_WARPS * seats > (long long)total) seats >>= 1;
  const int cohort = K3R_WARPS * seats;
  constexpr int SLOTS = PAIRED ? K3R_MAX_SEATS / 2 : K3R_MAX_SEATS;  // accumulators per lane
  const int slots = PAIRED ? seats / 2 : seats;
  const int half = PAIRED ? (lane >> 4) : 0;
  const int q = PAIRED ? (lane & 15) : lane;  // my four coordinates: 4q .. 4q + 3
  const bool on = 4 * q < D;
  const int nq = min(4, D - 4 * q);           // how many of them exist (the last quad of a row may be short)

  for (int c0 = blockIdx.x * cohort; c0 < total; c0 += gridDim.x * cohort) {
    // my seat in slot p: window index c0 + warp * seats + (PAIRED ? 2p + half : p)
    unsigned ent[SLOTS];
    bool has[SLOTS];
#pragma unroll
    for (int p = 0; p < SLOTS; p++) {
      const int seat = PAIRED ? 2 * p + half : p;
      const int i = c0 + warp * seats + seat;
      has[p] = p < slots && i < total;
      ent[p] = has[p] ? (P.list ? P.list[i].x : (unsigned)i) : 0u;
    }
    __syncthreads();  // the previous cohort's leaf records are no longer read
    // ---- leaf records of the warp's windows -> shared memory (whole warp per record: coalesced)
    for (int s = 0; s < seats; s++) {
      const int i = c0 + warp * seats + s;
      if (i >= total) break;
      const unsigned e = P.list ? P.list[i].x : (unsigned)i;
      const uint32_t *src = reinterpret_cast<const uint32_t *>(P.leaves + (size_t)e * kpad);
      uint32_t *dst = reinterpret_cast<uint32_t *>(leaves + (size_t)(warp * K3R_MAX_SEATS + s) * kpad);
      for (int j = lane; j < kpad / 4; j += 32) dst[j] = __ldg(src + j);
    }
    float4 acc[SLOTS];
#pragma unroll
    for (int p = 0; p < SLOTS; p++) {
      acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has[p] && on) {
        const float *from = P.in_shape ? P.in_shape + (size_t)ent[p] * D : P.mean_shape;
        acc[p].x = from[4 * q];
        if (nq > 1) acc[p].y = from[4 * q + 1];
        if (nq > 2) acc[p].z = from[4 * q + 2];
        if (nq > 3) acc[p].w = from[4 * q + 3];
      }
    }
    auto stage = [&](int ci, int buf) {
      const int carts = min(K3R_CHUNK, K - ci * K3R_CHUNK);
      const int n16 = carts * kLeaves * Dp / 4;  // 16-byte pieces
      const float4 *src = reinterpret_cast<const float4 *>(P.wp + (size_t)ci * chunk_floats);
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
      __syncthreads();  // chunk ci landed for everyone (and, first time round, the leaf records are written)
      const float *rb = rows + (size_t)(ci & 1) * chunk_floats + 4 * q;
      const int carts = min(K3R_CHUNK, K - ci * K3R_CHUNK);
#pragma unroll
      for (int p = 0; p < SLOTS; p++) {
        if (p >= slots) break;                     // (uniform)
        const int seat = PAIRED ? 2 * p + half : p;
        // the chunk's eight leaf indices of my window: one 32-bit load (byte k / 2, low nibble first; 8 | chunk start)
        const uint32_t lq = *reinterpret_cast<const uint32_t *>(leaves + (size_t)(warp * K3R_MAX_SEATS + seat) * kpad + ci * (K3R_CHUNK / 2));
        if (has[p] && on) {
#pragma unroll
          for (int cl = 0; cl < K3R_CHUNK; cl++) {
            if (cl < carts) {
              const uint32_t leaf = (lq >> (4 * cl)) & 0xfu;
              const float4 v = *reinterpret_cast<const float4 *>(rb + (cl * kLeaves + leaf) * Dp);
              acc[p].x = __fadd_rn(acc[p].x, v.x);
              acc[p].y = __fadd_rn(acc[p].y, v.y);
              acc[p].z = __fadd_rn(acc[p].z, v.z);
              acc[p].w = __fadd_rn(acc[p].w, v.w);
            }
          }
        }
      }
      __syncthreads();  // everyone is done with buffer ci & 1 before it is refilled
    }
#pragma unroll
    for (int p = 0; p < SLOTS; p++) {
      if (has[p] && on) {
        float *to = P.out_shape + (size_t)ent[p] * D + 4 * q;
        to[0] = acc[p].x;
        if (nq > 1) to[1] = acc[p].y;
        if (nq > 2) to[2] = acc[p].z;
        if (nq > 3) to[3] = acc[p].w;
      }
    }
  }
}

#endif  // __CUDACC__
}  // namespace jda
