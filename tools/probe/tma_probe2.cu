// Development probe: the CUDA programming guide's TMA example (libcu++ wrappers), 2-D.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out, int x, int y) {
  __shared__ alignas(128) float tile[8][64];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&tile, &tm, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = tile[i / 64][i % 64];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int W = 256, H = 64;
  std::vector<float> h(W * H);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 512 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap m;
  const cuuint64_t dims[2] = {W, H};
  const cuuint64_t str[1] = {W * 4};
  const cuuint32_t box[2] = {64, 8};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = ((Enc)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d\n", (int)r);
  k<<<1, 128>>>(m, o, 4, 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("run: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> r2(512);
    cudaMemcpy(r2.data(), o, 512 * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < 8; ++y) for (int x = 0; x < 64; ++x) if (r2[y * 64 + x] != (float)((2 + y) * W + 4 + x)) ++bad;
    printf("mismatches: %d\n", bad);
  }
  return 0;
}
