// standalone probe: 3-D u8 TMA box load like describe_tile_kernel
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
constexpr int BW = 256, BH = 90;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int z, uint8_t* out) {
  __shared__ __align__(128) uint8_t s_tile[BH][BW];
  __shared__ __align__(8) unsigned long long s_bar;
  if (threadIdx.x == 0) {
    const uint32_t bar = smem_u32(&s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BW * BH) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(&s_tile[0][0])), "l"(&map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
  }
  __syncthreads();
  const uint32_t bar = smem_u32(&s_bar);
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&s_tile[0][0])[i];
}
int main() {
  const int pitch = 1280, rows = 376, n = 2;
  std::vector<uint8_t> h((size_t)pitch * rows * n);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t *d, *o;
  cudaMalloc(&d, h.size()); cudaMalloc(&o, BW * BH);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  printf("entry point: %d %p %d\n", (int)e, fn, (int)qr);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap map;
  const cuuint64_t dims[3] = {pitch, rows, n}; const cuuint64_t strides[2] = {pitch, (cuuint64_t)pitch * rows};
  const cuuint32_t box[3] = {BW, BH, 1}; const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  const int cases[6][3] = {{0, 0, 0}, {16, 18, 1}, {1136, 338, 1}, {240, 17, 0}, {-16, -4, 0}, {-16, 360, 1}};
  for (auto& c : cases) {
    probe<<<1, 128>>>(map, c[0], c[1], c[2], o);
    e = cudaDeviceSynchronize();
    std::vector<uint8_t> got(BW * BH);
    cudaMemcpy(got.data(), o, got.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < BH; ++yy) for (int xx = 0; xx < BW; ++xx) {
      const int gx = c[0] + xx, gy = c[1] + yy;
      const uint8_t want = (gx >= 0 && gy >= 0 && gx < pitch && gy < rows) ? h[((size_t)c[2] * rows + gy) * pitch + gx] : 0;
      bad += got[yy * BW + xx] != want;
    }
    printf("case (%d,%d,%d): %s, mismatches %d\n", c[0], c[1], c[2], cudaGetErrorString(e), bad);
    if (e != cudaSuccess) break;
  }
  return 0;
}
