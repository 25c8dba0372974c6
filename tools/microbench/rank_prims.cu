// Micro-benchmark of candidate digit-ranking primitives on B200 (sm_100a).
// Decides the design of the radix partition kernel: smem atomics vs match.any vs thread-private counters.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int BINS>
__global__ void k_atom_shared(int iters, uint32_t* out) {
  __shared__ uint32_t h[BINS];
  for (int i = threadIdx.x; i < BINS; i += blockDim.x) h[i] = 0;
  __syncthreads();
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    acc += atomicAdd(&h[(s >> 20) & (BINS - 1)], 1u);
  }
  __syncthreads();
  if (acc == 0xdeadbeef) out[0] = acc;
  if (threadIdx.x < BINS) atomicAdd(&out[1 + (threadIdx.x & 15)], h[threadIdx.x]);
}

template <int BINS>
__global__ void k_atom_shared_noret(int iters, uint32_t* out) {
  __shared__ uint32_t h[BINS];
  for (int i = threadIdx.x; i < BINS; i += blockDim.x) h[i] = 0;
  __syncthreads();
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    atomicAdd(&h[(s >> 20) & (BINS - 1)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < BINS) atomicAdd(&out[1 + (threadIdx.x & 15)], h[threadIdx.x]);
}

// warp-private histograms
template <int BINS, int WARPS>
__global__ void k_atom_warp_private(int iters, uint32_t* out) {
  __shared__ uint32_t h[WARPS][BINS];
  for (int i = threadIdx.x; i < BINS * WARPS; i += blockDim.x) (&h[0][0])[i] = 0;
  __syncthreads();
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  uint32_t acc = 0;
  int w = threadIdx.x >> 5;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    acc += atomicAdd(&h[w][(s >> 20) & (BINS - 1)], 1u);
  }
  __syncthreads();
  if (acc == 0xdeadbeef) out[0] = acc;
  if (threadIdx.x < BINS) atomicAdd(&out[1 + (threadIdx.x & 15)], h[0][threadIdx.x]);
}

template <int BITS>
__global__ void k_match(int iters, uint32_t* out) {
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  uint32_t acc = 0;
  uint32_t lt = (1u << (threadIdx.x & 31)) - 1;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t d = (s >> 20) & ((1u << BITS) - 1);
    uint32_t m = __match_any_sync(0xffffffffu, d);
    acc += __popc(m & lt);
  }
  if (acc == 0xdeadbeef) out[0] = acc;
}

// CUB-like: match + warp-private counter update by leader
template <int BITS, int WARPS>
__global__ void k_match_rank(int iters, uint32_t* out) {
  __shared__ uint32_t h[WARPS][1 << BITS];
  for (int i = threadIdx.x; i < (WARPS << BITS); i += blockDim.x) (&h[0][0])[i] = 0;
  __syncthreads();
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  uint32_t acc = 0;
  int w = threadIdx.x >> 5;
  uint32_t lane = threadIdx.x & 31;
  uint32_t lt = (1u << lane) - 1;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t d = (s >> 20) & ((1u << BITS) - 1);
    uint32_t m = __match_any_sync(0xffffffffu, d);
    uint32_t base = h[w][d];
    __syncwarp();
    uint32_t r = __popc(m & lt);
    if (r == 0) h[w][d] = base + __popc(m);
    __syncwarp();
    acc += base + r;
  }
  if (acc == 0xdeadbeef) out[0] = acc;
}

// thread-private packed byte counters: word [(d>>2)][tid], conflict-free LDS+STS
template <int BITS, int T>
__global__ void k_thread_private(int iters, uint32_t* out) {
  extern __shared__ uint32_t c[];
  constexpr int ROWS = (1 << BITS) / 4;
  for (int i = threadIdx.x; i < ROWS * T; i += blockDim.x) c[i] = 0;
  __syncthreads();
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t d = (s >> 20) & ((1u << BITS) - 1);
    uint32_t sh = (d & 3) * 8;
    uint32_t idx = (d >> 2) * T + threadIdx.x;
    uint32_t wv = c[idx];
    acc += (wv >> sh) & 255u;
    c[idx] = wv + (1u << sh);
    if ((i & 127) == 127) {  // avoid overflow of byte counters: reset own column (cost amortised)
      for (int r = 0; r < ROWS; ++r) c[r * T + threadIdx.x] = 0;
    }
  }
  if (acc == 0xdeadbeef) out[0] = acc;
}

// 16-bit packed counters (CUB classic layout): word [(d & (HALF-1))][tid], half = d / HALF
template <int BITS, int T>
__global__ void k_thread_private16(int iters, uint32_t* out) {
  extern __shared__ uint32_t c[];
  constexpr int ROWS = (1 << BITS) / 2;
  for (int i = threadIdx.x; i < ROWS * T; i += blockDim.x) c[i] = 0;
  __syncthreads();
  uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1);
  uint32_t acc = 0;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    uint32_t d = (s >> 20) & ((1u << BITS) - 1);
    uint32_t sh = (d / ROWS) * 16;
    uint32_t idx = (d & (ROWS - 1)) * T + threadIdx.x;
    uint32_t wv = c[idx];
    acc += (wv >> sh) & 0xffffu;
    c[idx] = wv + (1u << sh);
  }
  if (acc == 0xdeadbeef) out[0] = acc;
}

template <typename F>
float time_it(F launch, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
  uint32_t* out; CK(cudaMalloc(&out, 64 * 4)); CK(cudaMemset(out, 0, 64 * 4));
  const int iters = 2048, T = 256;
  auto report = [&](const char* name, float ms, int ctas, int threads) {
    double recs = (double)ctas * threads * iters;
    printf("%-34s ctas/SM %d thr %d : %8.3f ms  %8.2f Grec/s  %6.3f rec/clk/SM @1.9GHz\n", name, ctas / sms, threads, ms,
           recs / ms * 1e-6, recs / (ms * 1e-3) / sms / 1.9e9);
  };
  for (int occ : {1, 2, 4, 8}) {
    int g = sms * occ;
    report("atomS ret 256 bins", time_it([&] { k_atom_shared<256><<<g, T>>>(iters, out); }, 5), g, T);
    report("atomS noret 256 bins", time_it([&] { k_atom_shared_noret<256><<<g, T>>>(iters, out); }, 5), g, T);
    report("atomS ret 2048 bins", time_it([&] { k_atom_shared<2048><<<g, T>>>(iters, out); }, 5), g, T);
    report("atomS warp-private 256", time_it([&] { k_atom_warp_private<256, 8><<<g, T>>>(iters, out); }, 5), g, T);
    report("match.any 8b", time_it([&] { k_match<8><<<g, T>>>(iters, out); }, 5), g, T);
    report("match.any 4b", time_it([&] { k_match<4><<<g, T>>>(iters, out); }, 5), g, T);
    report("match+warp ctr 8b", time_it([&] { k_match_rank<8, 8><<<g, T>>>(iters, out); }, 5), g, T);
  }
  {
    CK(cudaFuncSetAttribute(k_thread_private<8, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_thread_private<8, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_thread_private16<8, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(k_thread_private16<7, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    for (int occ : {1, 2, 3}) {
      int g = sms * occ;
      report("thread-private u8 8b T256", time_it([&] { k_thread_private<8, 256><<<g, 256, 64 * 1024>>>(iters, out); }, 5), g, 256);
      report("thread-private u8 8b T128", time_it([&] { k_thread_private<8, 128><<<g, 128, 32 * 1024>>>(iters, out); }, 5), g, 128);
      report("thread-private u16 8b T128", time_it([&] { k_thread_private16<8, 128><<<g, 128, 64 * 1024>>>(iters, out); }, 5), g, 128);
      report("thread-private u16 7b T256", time_it([&] { k_thread_private16<7, 256><<<g, 256, 64 * 1024>>>(iters, out); }, 5), g, 256);
    }
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
