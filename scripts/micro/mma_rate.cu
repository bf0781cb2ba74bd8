// Microbenchmark: issue rate / throughput of tcgen05.mma cta_group::1 M=128 from resident smem (no loads).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gan_sr_wind_field_b200/csrc/ptx.cuh"
using namespace ws;

__global__ void __launch_bounds__(128, 1) mma_rate(int n_umma, int reps, int n_acc, int kmajor, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t base = ptx::smem_u32(smem);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(ptx::smem_u32(&tslot), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 0) {
    const uint32_t idesc = ptx::make_idesc(1u, 128u, (uint32_t)n_umma, kmajor ? 0u : 1u, kmajor ? 0u : 1u);
    const uint64_t hi = ptx::make_smem_desc_sw128(0, kmajor ? 16 : 16384, 1024);
    const uint64_t ad = hi | ((base >> 4) & 0x3fff), bd = hi | (((base + 16384) >> 4) & 0x3fff);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + (uint32_t)((r % n_acc) * n_umma);
      if (ptx::elect_one()) {
        ptx::mma_f16_ss(d, ad, bd, idesc, 1u);
        ptx::mma_f16_ss(d, ad + 2, bd + 2, idesc, 1u);
        ptx::mma_f16_ss(d, ad + 4, bd + 4, idesc, 1u);
        ptx::mma_f16_ss(d, ad + 6, bd + 6, idesc, 1u);
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (ptx::elect_one()) ptx::mma_commit(ptx::smem_u32(&bar));
    __syncwarp();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  int reps = 2000;
  for (int grid : {1, 148}) for (int kmajor : {1, 0}) for (int n : {16, 32, 64, 128, 144, 256}) for (int nacc : {1, 3}) {
    if (nacc * n > 512) continue;
    mma_rate<<<grid, 128, 64 * 1024>>>(n, reps, nacc, kmajor, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("grid %3d %s N=%3d acc=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d) %s\n", grid,
           kmajor ? "K-major " : "MN-major", n, nacc, h[0] / (4.0 * reps), h[1] / (4.0 * reps), n / 2,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
