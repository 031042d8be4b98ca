// read_bw.cu — how fast can a kernel of SpMV's size (352 MB, almost all reads) stream from HBM on this GPU?
// Reference point for the SpMV roofline fraction: the copy peak in MEASURED_PEAKS.json is a long
// read+write stream; a 60-90 us read-only kernel includes its ramp-up and tail.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o read_bw read_bw.cu ; run: ./read_bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_read(const uint4* __restrict__ p, size_t n, unsigned long long* out) {
  unsigned long long acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const uint4 a = p[i], b = p[i + stride], c = p[i + 2 * stride], d = p[i + 3 * stride];
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  for (; i < n; i += stride) acc += p[i].x;
  if (acc == 0x1234567887654321ull) *out = acc;
}
int main() {
  const size_t sizes[] = {352u << 20, 1200u << 20, 4000ull << 20};
  unsigned long long* out;
  cudaMalloc(&out, 8);
  for (size_t bytes : sizes) {
    uint4* p;
    cudaMalloc(&p, bytes);
    cudaMemset(p, 1, bytes);
    // flush buffer larger than L2 between repetitions
    uint4* fl;
    cudaMalloc(&fl, 512u << 20);
    for (int blocks_per_sm : {4, 8, 16}) {
      float best = 1e9f, sum = 0;
      const int reps = 10;
      for (int r = 0; r < reps + 2; ++r) {
        cudaMemset(fl, r, 512u << 20);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_read<<<148 * blocks_per_sm, 256>>>(p, bytes / 16, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 2) { sum += ms; if (ms < best) best = ms; }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
      }
      printf("{\"bytes\": %zu, \"blocks_per_sm\": %d, \"ms_mean\": %.4f, \"ms_best\": %.4f, \"gbs_mean\": %.0f}\n", bytes, blocks_per_sm,
             sum / reps, best, bytes / (sum / reps) / 1e6);
    }
    cudaFree(p); cudaFree(fl);
  }
  return 0;
}
