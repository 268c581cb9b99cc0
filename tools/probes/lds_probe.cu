// lds_probe.cu -- what does one shared-memory load cost the SM?  (design probe for k2_scan, not product code)
// 148 blocks x 384 threads (k2_scan's shape); every warp issues independent LDS of one kind in a loop.
// Prints SM cycles per warp-level LDS instruction for: 1-byte loads (32 distinct banks / 2-, 4-way conflicts /
// one word broadcast), 8-byte loads with 1, 2, 4 distinct addresses per warp, 16-byte loads with 4 distinct
// addresses, and a mix that mimics one packet-cart of the scan.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_probe lds_probe.cu && ./lds_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int WARPS = 12, ITERS = 4096, U = 8;

template <int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1) probe(unsigned *out, long long *cycles) {
  extern __shared__ __align__(16) uint8_t smem[];
  for (int i = threadIdx.x; i < 65536; i += blockDim.x) smem[i] = (uint8_t)(i * 7);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned acc = 0;
  unsigned base = warp * 4096;
  // per-lane offset pattern
  unsigned lo;
  if (MODE == 0) lo = lane * 4;            // u8, 32 distinct banks
  else if (MODE == 1) lo = lane * 8;       // u8, 2-way conflict (16 banks, 2 words each)
  else if (MODE == 2) lo = lane * 16;      // u8, 4-way conflict
  else if (MODE == 3) lo = lane & 3;       // u8, one word: broadcast
  else if (MODE == 4) lo = 0;              // u64, one address
  else if (MODE == 5) lo = (lane & 1) * 8; // u64, two addresses (interleaved lanes)
  else if (MODE == 6) lo = (lane & 3) * 8; // u64, four addresses
  else if (MODE == 7) lo = (lane & 3) * 16;// u128, four addresses
  else if (MODE == 8) lo = lane * 2;       // u8, x-adjacent windows at step 2 (dense packet)
  else lo = lane * 4;
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
    const unsigned a = base + ((it * 40) & 2047);
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned addr = a + u * 132 + lo;
      if (MODE <= 3 || MODE == 8) acc += smem[addr];
      else if (MODE <= 6) { const uint2 v = *reinterpret_cast<const uint2 *>(smem + ((addr) & ~7u)); acc += v.x ^ v.y; }
      else if (MODE == 7) { const uint4 v = *reinterpret_cast<const uint4 *>(smem + ((addr) & ~15u)); acc += v.x ^ v.y ^ v.z ^ v.w; }
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, unsigned *out, long long *cyc) {
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
  probe<MODE><<<148, WARPS * 32, 65536 + 1024>>>(out, cyc);
  probe<MODE><<<148, WARPS * 32, 65536 + 1024>>>(out, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; i++) avg += h[i];
  avg /= 148;
  printf("%-44s %.2f SM cycles per warp LDS (12 warps/SM)   err=%s\n", name, avg / ((double)ITERS * U * WARPS),
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  unsigned *out; long long *cyc;
  cudaMalloc(&out, 148 * WARPS * 32 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>("LDS.U8  32 distinct banks", out, cyc);
  run<8>("LDS.U8  lanes 2 bytes apart (dense step 2)", out, cyc);
  run<3>("LDS.U8  one word (broadcast)", out, cyc);
  run<1>("LDS.U8  2-way bank conflict", out, cyc);
  run<2>("LDS.U8  4-way bank conflict", out, cyc);
  run<4>("LDS.64  one address per warp", out, cyc);
  run<5>("LDS.64  two addresses per warp", out, cyc);
  run<6>("LDS.64  four addresses per warp", out, cyc);
  run<7>("LDS.128 four addresses per warp", out, cyc);
  return 0;
}
