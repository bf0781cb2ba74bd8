// Microtest: K-major SWIZZLE_64B UMMA operands (rows of 64 bytes = 32 bf16 = two K=16 steps) with arbitrary row
// offsets — would let the K = 32-per-tap data-gradients of the RRDB trunk drop their zero-padded half chunk.
// Layout written by hand exactly as TMA SWIZZLE_64B would: 16-byte chunk index (0..3) XOR ((row >> 1) & 3).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sw64 sw64.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../gan_sr_wind_field_b200/csrc/ptx.cuh"
using namespace ws;

__host__ __device__ inline int aval(int r, int k) { return ((r * 7 + k * 3) % 13) - 6; }
__host__ __device__ inline int bval(int n, int k) { return ((n * 5 + k) % 7) - 3; }
constexpr int kRows = 320;

__global__ void __launch_bounds__(128, 1) sw64(int ro, int variant, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t base = ptx::smem_u32(smem);
  uint8_t* bsm = smem + kRows * 64 + 1024;
  bsm = (uint8_t*)(((uintptr_t)bsm + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  auto put = [&](uint8_t* tile, int row, int col, int v) {  // col: element 0..31 of the 64-byte row
    const uint32_t abs_row = (uint32_t)((ptx::smem_u32(tile) >> 6) + row);  // absolute 64-byte row index
    int chunk = col >> 3;
    if (variant == 0) chunk ^= (abs_row >> 1) & 3;   // address bits [4:5] ^= bits [7:8]
    else chunk ^= abs_row & 3;                       // alternative guess: bits [6:7]
    const int off = row * 64 + (chunk << 4) + (col & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(tile + off) = __float2bfloat16((float)v);
  };
  for (int i = threadIdx.x; i < kRows * 32; i += blockDim.x) put(smem, i / 32, i % 32, aval(i / 32, i % 32));
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) put(bsm, i / 32, i % 32, bval(i / 32, i % 32));
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(ptx::smem_u32(&tslot), 32); ptx::tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 0) {
    const uint32_t idesc = ptx::make_idesc(1u, 128u, 16u, 0u, 0u);
    // SWIZZLE_64B: layout_type 4, SBO = 8 rows x 64 B = 512
    uint64_t hi = 0;
    hi |= (uint64_t)((16u >> 4) & 0x3fffu) << 16;
    hi |= (uint64_t)((512u >> 4) & 0x3fffu) << 32;
    hi |= (uint64_t)1 << 46;
    hi |= (uint64_t)4 << 61;
    const uint32_t a_addr = base + (uint32_t)ro * 64u;
    const uint32_t b_addr = ptx::smem_u32(bsm);
    const uint64_t ad = hi | ((a_addr >> 4) & 0x3fff), bd = hi | ((b_addr >> 4) & 0x3fff);
    if (ptx::elect_one()) {
      for (int k = 0; k < 2; ++k) ptx::mma_f16_ss(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
      ptx::mma_commit(ptx::smem_u32(&bar));
    }
    __syncwarp();
  }
  ptx::mbar_wait(ptx::smem_u32(&bar), 0);
  ptx::tc_fence_after();
  uint32_t r[16];
  ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
  ptx::tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[threadIdx.x * 16 + j] = __uint_as_float(r[j]);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 32); }
}

int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  float h[128 * 16];
  cudaFuncSetAttribute(sw64, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int variant = 0; variant < 2; ++variant)
    for (int ro : {0, 8, 16, 1, 2, 3, 4, 5, 7, 12, 13}) {
      cudaMemset(d, 0, sizeof(h));
      sw64<<<1, 128, 48 * 1024>>>(ro, variant, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d ro %d: %s\n", variant, ro, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
          double ref = 0;
          for (int k = 0; k < 32; ++k) ref += (double)aval(ro + m, k) * bval(n, k);
          double err = fabs(ref - h[m * 16 + n]);
          if (err > worst) worst = err;
        }
      printf("SW64 pattern %s row_offset=%2d : max abs err %.1f %s\n", variant ? "bits[6:7]" : "bits[7:8]", ro, worst,
             worst == 0 ? "OK" : "WRONG");
    }
  return 0;
}
