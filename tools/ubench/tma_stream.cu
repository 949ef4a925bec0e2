// tma_stream.cu — micro-benchmark: how fast can one producer thread per SM stream a [rows, 384] fp32 matrix into shared
// memory with cp.async.bulk.tensor boxes of [box_rows x 128 B] (SWIZZLE_128B), as a function of the box height and of the
// ring depth?  (Consumer = a second thread that only waits and releases the stage.)  Answers "is the wgrad / conv pipeline
// limited by the per-box cost of TMA?" — profiles/r02_ubench_tma_stream.jsonl.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_stream tools/ubench/tma_stream.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void tma2d(const CUtensorMap* m, uint64_t* bar, uint32_t dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"((uint64_t)m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// each CTA streams row blocks [blk*box_rows, +box_rows) for blk = blockIdx.x, +gridDim.x, ...; every row block = nbox boxes
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int n_blocks, int box_rows, int nbox, int stages, int lanes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t stage_bytes = (uint32_t)box_rows * 128 * nbox;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* empty = full + 16;
  const int tid = threadIdx.x;
  if (tid == 0) { for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  const uint32_t base = smem_u32(smem);
  if (tid < 32) {
    int it = 0;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      if (tid == 0) { mbar_wait(&empty[s], ph ^ 1); mbar_expect(&full[s], stage_bytes); }
      __syncwarp();
      if (lanes == 1) { if (tid == 0) for (int j = 0; j < nbox; ++j) tma2d(&tm, &full[s], base + s * stage_bytes + j * box_rows * 128, j * 32, blk * box_rows); }
      else if (tid < nbox) tma2d(&tm, &full[s], base + s * stage_bytes + tid * box_rows * 128, tid * 32, blk * box_rows);
      __syncwarp();
    }
  } else if (tid == 32) {
    int it = 0;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
      const int s = it % stages; const uint32_t ph = (it / stages) & 1;
      mbar_wait(&full[s], ph);
      mbar_arrive(&empty[s]);
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int64_t rows = 276480 * 2, cols = 384;
  float* d; cudaMalloc(&d, rows * cols * 4); cudaMemset(d, 0, rows * cols * 4);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int box_rows_list[] = {32, 64, 128, 256};
  for (int bi = 0; bi < 4; ++bi) for (int nbox = 1; nbox <= 4; nbox *= 2) for (int lanes = 1; lanes <= 2; ++lanes) for (int budget_kb = 64; budget_kb <= 192; budget_kb += 128) {
    const int box_rows = box_rows_list[bi];
    const uint32_t stage_bytes = box_rows * 128 * nbox;
    int stages = budget_kb * 1024 / stage_bytes; if (stages > 16) stages = 16; if (stages < 2) continue;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    const int n_blocks = (int)(rows / box_rows);
    const size_t smem = (size_t)stages * stage_bytes + 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) stream_kernel<<<148, 64, smem>>>(tm, n_blocks, box_rows, nbox, stages, lanes);
    cudaEventRecord(e0);
    for (int w = 0; w < 5; ++w) stream_kernel<<<148, 64, smem>>>(tm, n_blocks, box_rows, nbox, stages, lanes);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double bytes = (double)rows * 128.0 * nbox;
    printf("{\"box_rows\": %d, \"boxes_per_stage\": %d, \"issuing_lanes\": %d, \"stages\": %d, \"stage_KB\": %.0f, \"ms\": %.4f, \"GBps\": %.0f, \"cycles_per_box_at_1.9GHz\": %.0f, \"err\": \"%s\"}\n",
           box_rows, nbox, lanes, stages, stage_bytes / 1024.0, ms, bytes / ms / 1e6, ms * 1e-3 * 1.9e9 / ((double)n_blocks * nbox / 148.0), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
