// Microbenchmark 3: tcgen05.mma rate (a) as a CTA pair (cta_group::2, M = 256) and (b) while another warp streams
// bulk copies global -> shared memory (the TMA write traffic of a real kernel competing for the SMEM port).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate3 mma_rate3.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gan_sr_wind_field_b200/csrc/ptx.cuh"
using namespace ws;

__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <bool kPair>
__global__ void __launch_bounds__(128, 1) mma_rate3(int n_umma, int reps, int n_acc, int bg_chunk, const uint8_t* src,
                                                    long long* out, int random_data, int commit_every, int wait_every) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t base = ptx::smem_u32(smem);
  __shared__ uint64_t bar, bgbar[2];
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  __shared__ long long bg_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kPair ? ptx::cluster_ctarank() : 0u;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    // random bf16 pairs in [-2, 2): sign + exponent 0x3f/0x3e + random mantissa
    ((uint32_t*)smem)[i] = (random_data & 1) ? ((h & 0x80ff80ffu) | 0x3f003f00u) : 0x3c003c00u;
  }
  __shared__ uint64_t dummy_bar, ready_bar;
  if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&dummy_bar), 1); ptx::mbar_init(ptx::smem_u32(&ready_bar), 1); }
  if (threadIdx.x == 0) {
    ptx::mbar_init(ptx::smem_u32(&bar), 1);
    ptx::mbar_init(ptx::smem_u32(&bgbar[0]), 1);
    ptx::mbar_init(ptx::smem_u32(&bgbar[1]), 1);
    ptx::fence_mbar_init();
    stop = 0;
    bg_bytes = 0;
  }
  if (warp == 1) {
    if (kPair) { ptx::tmem_alloc2(ptx::smem_u32(&tslot), 512); ptx::tmem_relinquish2(); }
    else { ptx::tmem_alloc(ptx::smem_u32(&tslot), 512); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  if (kPair) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tslot;
  if (warp == 0) {
    long long t0 = clock64(), t1 = t0;
    if (rank == 0) {
      const uint32_t idesc = ptx::make_idesc(1u, kPair ? 256u : 128u, (uint32_t)n_umma, 0u, 0u);
      const uint64_t hi = ptx::make_smem_desc_sw128(0, 16, 1024);
      const uint64_t ad = hi | ((base >> 4) & 0x3fff), bd = hi | (((base + 32768) >> 4) & 0x3fff);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t d = tmem + (uint32_t)((r % n_acc) * n_umma);
        const uint64_t a = ad + (uint64_t)((r % 3) * 40 * 8);  // row offsets like the kx taps of a halo tile
        if (ptx::elect_one()) {
          if (kPair) {
            ptx::mma_f16_ss2(d, a, bd, idesc, 1u);
            ptx::mma_f16_ss2(d, a + 2, bd + 2, idesc, 1u);
            ptx::mma_f16_ss2(d, a + 4, bd + 4, idesc, 1u);
            ptx::mma_f16_ss2(d, a + 6, bd + 6, idesc, 1u);
          } else {
            ptx::mma_f16_ss(d, a, bd, idesc, 1u);
            ptx::mma_f16_ss(d, a + 2, bd + 2, idesc, 1u);
            ptx::mma_f16_ss(d, a + 4, bd + 4, idesc, 1u);
            ptx::mma_f16_ss(d, a + 6, bd + 6, idesc, 1u);
          }
        }
        __syncwarp();
        if (commit_every && (r + 1) % commit_every == 0) {
          if (ptx::elect_one()) { if (kPair) ptx::mma_commit2(ptx::smem_u32(&dummy_bar)); else ptx::mma_commit(ptx::smem_u32(&dummy_bar)); }
          __syncwarp();
        }
        if (wait_every && (r + 1) % wait_every == 0) {
          // an already-completed barrier wait + fence, as at every tap of the conv kernel (ready_bar never flips: parity 1 passes)
          if (random_data & 4) { while (!mbar_test_wait(ptx::smem_u32(&ready_bar), 1u)) {} }
          else ptx::mbar_wait(ptx::smem_u32(&ready_bar), 1u);
          if (wait_every > 0 && !(random_data & 2)) ptx::tc_fence_after();
        }
      }
      t1 = clock64();
      if (ptx::elect_one()) { if (kPair) ptx::mma_commit2(ptx::smem_u32(&bar)); else ptx::mma_commit(ptx::smem_u32(&bar)); }
      __syncwarp();
    }
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    long long t2 = clock64();
    stop = 1;
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 2 && bg_chunk > 0) {
    // background: bulk copies into a scratch region (64 KB .. 64 KB + 2 * bg_chunk), two in flight
    uint32_t ph[2] = {0, 0};
    long long n = 0;
    const uint8_t* s = src + (size_t)blockIdx.x * (1 << 20);
    int issued[2] = {0, 0};
    int i = 0;
    while (!stop) {
      const int b = i & 1;
      if (issued[b]) { ptx::mbar_wait(ptx::smem_u32(&bgbar[b]), ph[b]); ph[b] ^= 1u; n += bg_chunk; }
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(ptx::smem_u32(&bgbar[b]), (uint32_t)bg_chunk);
        bulk_g2s(base + 65536 + b * bg_chunk, s + (size_t)((i * bg_chunk) & ((1 << 20) - 1)), (uint32_t)bg_chunk,
                 ptx::smem_u32(&bgbar[b]));
      }
      __syncwarp();
      issued[b] = 1;
      ++i;
    }
    for (int b = 0; b < 2; ++b)
      if (issued[b]) ptx::mbar_wait(ptx::smem_u32(&bgbar[b]), ph[b]);
    if (lane == 0 && blockIdx.x == 0) out[2] = n;
  }
  if (kPair) ptx::cluster_sync(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if (kPair) ptx::tmem_dealloc2(tmem, 512); else ptx::tmem_dealloc(tmem, 512);
  }
}

template <bool kPair>
void run(int n, int bg, const uint8_t* src, long long* d, int rnd = 0, int ce = 0, int we = 0) {
  const int reps = 4000;
  cudaMemset(d, 0, 32);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kPair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, mma_rate3<kPair>, n, reps, 512 / n < 3 ? 512 / n : 3, bg, src, d, rnd, ce, we);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("%s N=%3d rnd=%d commit/%d wait/%d bg_chunk=%6d: issue %.1f cyc/MMA, complete %.1f cyc/MMA; background %.1f B/clk/SM %s\n",
         kPair ? "pair  " : "single", n, rnd, ce, we, bg, h[0] / (4.0 * reps), h[1] / (4.0 * reps), (double)h[2] / (double)h[1],
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 32);
  uint8_t* src; cudaMalloc(&src, (size_t)148 << 20); cudaMemset(src, 0, (size_t)148 << 20);
  cudaFuncSetAttribute(mma_rate3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(mma_rate3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int n : {128, 144}) {
    for (int rnd : {0, 1}) { run<false>(n, 0, src, d, rnd); run<true>(n, 0, src, d, rnd); }
    for (int ce : {1, 3}) { run<false>(n, 0, src, d, 1, ce, 0); run<true>(n, 0, src, d, 1, ce, 0); }
    for (int we : {1, 3}) { run<false>(n, 0, src, d, 1, we, we); run<true>(n, 0, src, d, 1, we, we); }
    // rnd & 2: the barrier wait WITHOUT the tcgen05 fence; commit 0: waits only
    run<false>(n, 0, src, d, 3, 1, 1); run<false>(n, 0, src, d, 1, 0, 1); run<false>(n, 0, src, d, 3, 0, 1);
    run<false>(n, 0, src, d, 7, 0, 1);  // rnd & 4: mbarrier.test_wait instead of try_wait
    run<true>(n, 16384, src, d, 1, 3, 3);
  }
  return 0;
}
