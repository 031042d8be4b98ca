// convert.cu — index-width conversion at the host boundary and small u64 utilities.
// The reference hands over usize (u64) indices (spam_csr/src/lib.rs:30-31); the device keeps
// u32 column indices.  Narrowing/widening happens on the device after/before a raw PCIe copy.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_narrow(const u64* __restrict__ in, u32* __restrict__ out, u64 n) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const u64 v = in[i];
    out[i] = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)v;  // saturate: the validation pass then reports it as >= cols
  }
}
__global__ void __launch_bounds__(256) k_widen(const u32* __restrict__ in, u64* __restrict__ out, u64 n) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}
__global__ void __launch_bounds__(256) k_add_offset(const u64* in, u64* out, u64 n, u64 off) {
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i] + off;
}

unsigned grid_for(const spam_handle* h, u64 n) {
  u64 g = (n + 255) / 256;
  const u64 cap = (u64)h->num_sms * 16;
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (unsigned)g;
}

}  // namespace

int narrow_u64_to_u32(spam_handle* h, const u64* in, u32* out, u64 n) {
  if (!n) return SPAM_OK;
  k_narrow<<<grid_for(h, n), 256, 0, h->stream>>>(in, out, n);
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}
int widen_u32_to_u64(spam_handle* h, const u32* in, u64* out, u64 n) {
  if (!n) return SPAM_OK;
  k_widen<<<grid_for(h, n), 256, 0, h->stream>>>(in, out, n);
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}
int add_offset_u64(spam_handle* h, u64* p, u64 n, u64 off, const u64* src) {
  if (!n) return SPAM_OK;
  k_add_offset<<<grid_for(h, n), 256, 0, h->stream>>>(src ? src : p, p, n, off);
  count_launch(h);
  CK(cudaGetLastError());
  return SPAM_OK;
}
