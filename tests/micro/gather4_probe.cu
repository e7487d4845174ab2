// Probe of the sm_100 TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4): which tensor-map box shape it wants
// and how the four gathered rows land in SWIZZLE_128B shared memory.  Stand-alone; build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tests/micro/gather4_probe tests/micro/gather4_probe.cu -lcuda
// Prints one line per box variant: encode status, and whether smem row i == source row idx[i] (un-swizzled compare).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int r0, int r1, int r2, int r3, int col0, uint32_t tx_bytes,
                      uint4* out, int n_vec) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_vec * 16; ++i) smem[i] = 0xEE;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(tx_bytes));
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem)),
        "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(col0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
    uint32_t ok = 0;
    for (int spin = 0; spin < 2000000 && !ok; ++spin)
      asm volatile("{\n.reg .pred P;\nmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\nselp.u32 %0, 1, 0, P;\n}"
                   : "=r"(ok) : "r"(smem_u32(&bar)));
    out[n_vec] = make_uint4(ok, 0, 0, 0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_vec; i += blockDim.x) out[i] = reinterpret_cast<uint4*>(smem)[i];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 4096, cols = 128;
  __nv_bfloat16* h = (__nv_bfloat16*)malloc(sizeof(__nv_bfloat16) * rows * cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) h[r * cols + c] = __float2bfloat16((float)(r % 251) + (float)c / 256.f);
  __nv_bfloat16* d;
  CHECK(cudaMalloc(&d, sizeof(__nv_bfloat16) * rows * cols));
  CHECK(cudaMemcpy(d, h, sizeof(__nv_bfloat16) * rows * cols, cudaMemcpyHostToDevice));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  const int idx[4] = {5, 1700, 2, 901};
  const int n_vec = 4 * 128 / 16 * 2;   // room for 4 rows x 128 B, twice
  uint4* out;
  CHECK(cudaMalloc(&out, sizeof(uint4) * (n_vec + 1)));
  uint4* hout = (uint4*)malloc(sizeof(uint4) * (n_vec + 1));
  CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
  const int boxes[3][2] = {{64, 1}, {64, 4}, {64, 128}};
  const CUtensorMapSwizzle swz[2] = {CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_SWIZZLE_NONE};
  for (int s = 0; s < 2; ++s)
    for (int b = 0; b < 3; ++b) {
      CUtensorMap map;
      cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
      cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
      cuuint32_t box[2] = {(cuuint32_t)boxes[b][0], (cuuint32_t)boxes[b][1]};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz[s], CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("swizzle %s box (%d, %d): encode %d", s == 0 ? "128B" : "none", boxes[b][0], boxes[b][1], (int)r);
      if (r != CUDA_SUCCESS) { printf("\n"); continue; }
      CHECK(cudaMemset(out, 0, sizeof(uint4) * (n_vec + 1)));
      probe<<<1, 128, 8192>>>(map, idx[0], idx[1], idx[2], idx[3], 64, 4 * 64 * 2, out, n_vec);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" launch -> %s\n", cudaGetErrorString(e)); return 3; }
      CHECK(cudaMemcpy(hout, out, sizeof(uint4) * (n_vec + 1), cudaMemcpyDeviceToHost));
      printf(" barrier %u;", hout[n_vec].x);
      // row i of the destination: 128 bytes at i * 128; un-swizzle 16-byte chunk c -> c ^ (i & 7) when swizzled
      int good_plain = 0, good_swz = 0;
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat16* src = h + idx[i] * cols + 64;
        const uint8_t* dst = (const uint8_t*)hout + i * 128;
        int plain = 1, sw = 1;
        for (int c = 0; c < 8; ++c) {
          plain &= memcmp(dst + c * 16, src + c * 8, 16) == 0;
          sw &= memcmp(dst + ((c ^ (i & 7)) * 16), src + c * 8, 16) == 0;
        }
        good_plain += plain;
        good_swz += sw;
      }
      printf(" rows matching: linear %d/4, swizzled %d/4; first bf16 of rows:", good_plain, good_swz);
      for (int i = 0; i < 5; ++i) printf(" %.3f", __bfloat162float(((const __nv_bfloat16*)hout)[i * 64]));
      printf("\n");
    }
  return 0;
}
