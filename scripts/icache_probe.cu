// Probe: does instruction supply collapse when two warps of one scheduler run different
// straight-line code?  (diagnostic for the no_instruction stalls of k_ihwbc_reg)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_probe icache_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int BODY, int COPY>
__device__ __noinline__ void body(float (&a)[8], float m, int reps) {
  for (int r = 0; r < reps; r++) {
#pragma unroll
    for (int i = 0; i < BODY / 8; i++) {
#pragma unroll
      for (int k = 0; k < 8; k++) a[k] = fmaf(a[k], m, float(COPY + 1) * 0.25f + float(i & 7));
    }
  }
}

template <int BODY>
__global__ void k_probe(float* out, float m, int reps, int mode, long long* cyc) {
  float a[8];
  for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 0.001f + k;
  const int w = threadIdx.x >> 5;
  const int copy = mode == 0 ? 0 : (mode == 1 ? (w >> 2) & 1 : w & 7);
  long long t0 = clock64();
  switch (copy) {
    case 0: body<BODY, 0>(a, m, reps); break;
    case 1: body<BODY, 1>(a, m, reps); break;
    case 2: body<BODY, 2>(a, m, reps); break;
    case 3: body<BODY, 3>(a, m, reps); break;
    case 4: body<BODY, 4>(a, m, reps); break;
    case 5: body<BODY, 5>(a, m, reps); break;
    case 6: body<BODY, 6>(a, m, reps); break;
    default: body<BODY, 7>(a, m, reps); break;
  }
  long long t1 = clock64();
  float s = 0;
  for (int k = 0; k < 8; k++) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int BODY>
void run(float* out, long long* cyc) {
  const int reps = (1 << 22) / BODY;
  for (int warps : {1, 4, 8, 16}) {
    for (int mode = 0; mode < 3; mode++) {
      if (warps <= 4 && mode == 1) continue;
      k_probe<BODY><<<148, warps * 32>>>(out, 0.999f, reps, mode, cyc);
      cudaDeviceSynchronize();
      k_probe<BODY><<<148, warps * 32>>>(out, 0.999f, reps, mode, cyc);
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double instr = double(reps) * BODY;
      printf("body %6d instr (%4d KB)  warps/SM %2d  mode %d (%s)  cycles/instr/warp %.3f  IPC/SM %.2f\n", BODY, BODY * 16 / 1024,
             warps, mode, mode == 0 ? "same code" : mode == 1 ? "2 copies, split by scheduler pair" : "copy = warp&7",
             c / instr, warps * instr / c);
    }
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  run<256>(out, cyc);
  run<1024>(out, cyc);
  run<4096>(out, cyc);
  run<16384>(out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
