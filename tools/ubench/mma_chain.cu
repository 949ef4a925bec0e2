// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128) as a function of N, operand source (SS / TS) and the
// number of independent accumulators the stream of MMAs rotates over.  One CTA, one issuing lane.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_chain mma_chain.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t idesc(int N, int b_mn) {
  return (1u << 4) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// mode 0: TS (A from TMEM cols 256..), B MN-major;  mode 1: SS, both K-major
template <int mode, int N, int n_acc, int n_mma>
__global__ void __launch_bounds__(128, 1) chain_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 32) {
    const uint32_t tm = __shfl_sync(0xffffffffu, slot, 0);
    const uint32_t id = idesc(N, mode == 0 ? 1 : 0);
    const uint64_t a_desc = desc_sw64(smem_u32(smem)), b_desc = desc_sw64(smem_u32(smem) + 32768);
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = clock64();
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < n_mma; ++i) {
          const uint32_t d = tm + (uint32_t)(i % n_acc) * (uint32_t)N;
          if (mode == 0) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                         "r"(tm + 448 + (uint32_t)(i & 7) * 8), "l"(b_desc + (uint64_t)((i & 3) * 64)), "r"(id), "r"(1u) : "memory");
          } else {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a_desc + (uint64_t)((i & 1) * 2)), "l"(b_desc + (uint64_t)((i & 1) * 2)), "r"(id), "r"(1u) : "memory");
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      t1 = clock64();
      uint32_t done;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)(rep & 1)) : "memory");
      } while (!done);
      t2 = clock64();
    }
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
  }
}

template <int mode, int N, int n_acc>
void run(long long* d_out) {
  constexpr int n_mma = 48;
  long long h[2];
  cudaFuncSetAttribute(chain_kernel<mode, N, n_acc, n_mma>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  chain_kernel<mode, N, n_acc, n_mma><<<1, 128, 64 * 1024>>>(d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  printf("{\"mode\": \"%s\", \"N\": %d, \"accumulators\": %d, \"n_mma\": %d, \"issue_cyc_per_mma\": %.1f, \"total_cyc_per_mma\": %.1f, \"floor\": %d}\n",
         mode == 0 ? "TS" : "SS", N, n_acc, n_mma, (double)h[0] / n_mma, (double)h[1] / n_mma, 128 * N / 256);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  run<0, 32, 1>(d_out); run<0, 32, 2>(d_out); run<0, 32, 3>(d_out); run<0, 32, 6>(d_out);
  run<0, 64, 1>(d_out); run<0, 64, 2>(d_out); run<0, 64, 4>(d_out);
  run<0, 128, 1>(d_out); run<0, 128, 2>(d_out); run<0, 256, 1>(d_out);
  run<1, 32, 1>(d_out); run<1, 32, 4>(d_out); run<1, 64, 1>(d_out); run<1, 64, 2>(d_out); run<1, 64, 4>(d_out);
  run<1, 128, 1>(d_out); run<1, 128, 2>(d_out); run<1, 256, 1>(d_out);
  return 0;
}
