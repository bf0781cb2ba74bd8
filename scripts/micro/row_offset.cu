// Microtest: can a SWIZZLE_128B UMMA operand start at an arbitrary 128-byte ROW of a 1024-byte-aligned tile
// (start address not a multiple of the 8-row swizzle atom)?  Variants: descriptor base_offset field 0 or
// (row & 7).  K-major A (row = M index) and MN-major A (row = K index).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o row_offset row_offset.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../gan_sr_wind_field_b200/csrc/ptx.cuh"
using namespace ws;

__host__ __device__ inline int aval(int r, int k) { return ((r * 7 + k * 3) % 13) - 6; }
__host__ __device__ inline int bval(int n, int k) { return ((n * 5 + k) % 7) - 3; }

constexpr int kRows = 320;  // rows of the A tile resident in smem (per 64-wide block)
constexpr int kBlk = kRows * 128;

// mode 0: A K-major  : A[m][k], m = row (0..kRows), k in 0..63.  D[m][n] = sum_k A[ro+m][k] B[n][k], K = 64
// mode 1: A MN-major : A[k][m], k = row, m in 0..127 (2 blocks). D[m][n] = sum_k A[ro+k][m] B[k][n], K = 64
__global__ void __launch_bounds__(128, 1) row_offset(int mode, int ro, int bo, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t base = ptx::smem_u32(smem);
  uint8_t* bsm = smem + 2 * kBlk;
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  auto put = [&](uint8_t* tile, int row, int col, int v) {  // col: element index within the 64-wide row
    const int off = row * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(tile + off) = __float2bfloat16((float)v);
  };
  if (mode == 0) {
    for (int i = threadIdx.x; i < kRows * 64; i += blockDim.x) put(smem, i / 64, i % 64, aval(i / 64, i % 64));
    for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) put(bsm, i / 64, i % 64, bval(i / 64, i % 64));
  } else {
    for (int i = threadIdx.x; i < kRows * 128; i += blockDim.x) {
      const int k = i / 128, m = i % 128;
      put(smem + (m / 64) * kBlk, k, m % 64, aval(m, k));
    }
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
      const int k = i / 64, n = i % 64;
      put(bsm, k, n, n < 16 ? bval(n, k) : 0);
    }
  }
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(ptx::smem_u32(&tslot), 32); ptx::tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 0) {
    const uint32_t idesc = ptx::make_idesc(1u, 128u, 16u, mode ? 1u : 0u, mode ? 1u : 0u);
    const uint64_t hi = ptx::make_smem_desc_sw128(0, mode ? (uint32_t)kBlk : 16u, 1024);
    const uint32_t a_addr = base + (uint32_t)ro * 128u;
    const uint32_t b_addr = base + 2u * kBlk;
    uint64_t ad = hi | ((a_addr >> 4) & 0x3fff);
    if (bo) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t bd = hi | ((b_addr >> 4) & 0x3fff);
    const int step = mode ? 128 : 2;  // per-k16 advance in the (addr >> 4) field
    if (ptx::elect_one()) {
      for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem, ad + (uint64_t)(k * step), bd + (uint64_t)(k * step), idesc, k ? 1u : 0u);
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
  cudaFuncSetAttribute(row_offset, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  for (int mode = 0; mode < 2; ++mode)
    for (int bo = 0; bo < 2; ++bo)
      for (int ro : {0, 8, 1, 2, 3, 5, 7, 9, 12, 13, 27}) {
        cudaMemset(d, 0, sizeof(h));
        row_offset<<<1, 128, 2 * kBlk + 8192 + 1024>>>(mode, ro, bo, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d bo %d ro %d: %s\n", mode, bo, ro, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double worst = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 16; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k)
              ref += mode == 0 ? (double)aval(ro + m, k) * bval(n, k) : (double)aval(m, ro + k) * bval(n, k);
            double err = fabs(ref - h[m * 16 + n]);
            if (err > worst) worst = err;
          }
        printf("%s base_offset=%s row_offset=%2d : max abs err %.1f %s\n", mode ? "MN-major" : "K-major ",
               bo ? "(row&7)" : "0      ", ro, worst, worst == 0 ? "OK" : "WRONG");
      }
  return 0;
}
