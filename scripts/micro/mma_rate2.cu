// Microbenchmark 2: does the START ROW of a SWIZZLE_128B K-major A operand change the tcgen05.mma rate?
// conv_tc2.cu addresses the kx taps of a halo tile as descriptor row offsets (tap * slabrows rows, e.g. 40 rows:
// not a multiple of the 8-row swizzle atom).  Also: A operands walking through a large buffer instead of one
// resident 4 KB tile, and N = 128 / 144.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate2 mma_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gan_sr_wind_field_b200/csrc/ptx.cuh"
using namespace ws;

__global__ void __launch_bounds__(128, 1) mma_rate2(int n_umma, int reps, int n_acc, int a_off_rows, int walk_rows,
                                                    long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t base = ptx::smem_u32(smem);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); ptx::fence_mbar_init(); }
  if (warp == 1) { ptx::tmem_alloc(ptx::smem_u32(&tslot), 512); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 0) {
    const uint32_t idesc = ptx::make_idesc(1u, 128u, (uint32_t)n_umma, 0u, 0u);
    const uint64_t hi = ptx::make_smem_desc_sw128(0, 16, 1024);
    const uint32_t b_addr = base + 128 * 1024;  // B tile: 256 rows x 128 B at most
    const uint64_t bd = hi | ((b_addr >> 4) & 0x3fff);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long t0 = clock64();
    int w = 0;
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + (uint32_t)((r % n_acc) * n_umma);
      const uint32_t a_addr = base + (uint32_t)(a_off_rows + w) * 128u;
      const uint64_t ad = hi | ((a_addr >> 4) & 0x3fff);
      if (ptx::elect_one()) {
        ptx::mma_f16_ss(d, ad, bd, idesc, 1u);
        ptx::mma_f16_ss(d, ad + 2, bd + 2, idesc, 1u);
        ptx::mma_f16_ss(d, ad + 4, bd + 4, idesc, 1u);
        ptx::mma_f16_ss(d, ad + 6, bd + 6, idesc, 1u);
      }
      __syncwarp();
      w += walk_rows;
      if (w + a_off_rows + 128 > 768) w = 0;  // stay inside 96 KB
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
  cudaFuncSetAttribute(mma_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int reps = 2000;
  for (int n : {128, 144}) for (int off : {0, 8, 1, 4, 40}) for (int walk : {0, 128, 40}) {
    mma_rate2<<<148, 128, 170 * 1024>>>(n, reps, 3, off, walk, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d acc=3 A start row %2d walk %3d rows: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d) %s\n", n, off, walk,
           h[0] / (4.0 * reps), h[1] / (4.0 * reps), n / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
