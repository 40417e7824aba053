// Is a row re-pitch that reads the images straight from page-locked HOST memory (SM loads over PCIe) faster than
// H2D copy + re-pitch from device memory?  One stereo pair, KITTI and 1920x1080 shapes, CUDA events, 200 repetitions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zc_probe zc_probe.cu && ./zc_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void repitch(const uint8_t* __restrict__ src, int cols, int rows, int pitch, uint8_t* __restrict__ dst) {
  const int row = blockIdx.y, img = blockIdx.z;
  const uint8_t* s = src + ((size_t)img * rows + row) * cols;
  uint32_t* d = reinterpret_cast<uint32_t*>(dst + ((size_t)img * rows + row) * pitch);
  const size_t a0 = reinterpret_cast<size_t>(s);
  for (int w = blockIdx.x * 256 + threadIdx.x; w < pitch / 4; w += gridDim.x * 256) {
    uint32_t v = 0;
    if (4 * w < cols) {
      const size_t a = a0 + 4 * (size_t)w;
      const uint32_t* p = reinterpret_cast<const uint32_t*>(a & ~(size_t)3);
      const unsigned sh = (unsigned)(a & 3) * 8;
      const uint32_t lo = p[0], hi = sh ? p[1] : 0u;
      v = __funnelshift_r(lo, hi, sh);
    }
    d[w] = v;
  }
}

int main() {
  const int shapes[2][2] = {{1241, 376}, {1920, 1080}};
  cudaStream_t st; cudaStreamCreate(&st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (auto& sh : shapes) {
    const int cols = sh[0], rows = sh[1], pitch = (cols + 127) / 128 * 128;
    const size_t bytes = (size_t)2 * cols * rows;
    uint8_t *h, *hd, *stage, *img;
    cudaHostAlloc(&h, bytes + 64, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&hd, h, 0);
    for (size_t i = 0; i < bytes; ++i) h[i] = (uint8_t)(i * 7);
    cudaMalloc(&stage, bytes + 64); cudaMalloc(&img, (size_t)2 * rows * pitch);
    dim3 grid((pitch / 4 + 255) / 256, rows, 2);
    float ms_copy = 0, ms_zc = 0;
    for (int mode = 0; mode < 2; ++mode) {
      for (int rep = 0; rep < 220; ++rep) {
        if (rep == 20) cudaEventRecord(e0, st);
        if (mode == 0) {
          cudaMemcpyAsync(stage, h, bytes, cudaMemcpyHostToDevice, st);
          repitch<<<grid, 256, 0, st>>>(stage, cols, rows, pitch, img);
        } else {
          repitch<<<grid, 256, 0, st>>>(hd, cols, rows, pitch, img);
        }
      }
      cudaEventRecord(e1, st); cudaStreamSynchronize(st);
      cudaEventElapsedTime(mode == 0 ? &ms_copy : &ms_zc, e0, e1);
    }
    printf("%dx%d pair (%.2f MB): copy + repitch %.1f us, zero-copy repitch %.1f us per pair (back to back in one stream)\n",
           cols, rows, bytes / 1e6, ms_copy * 1000 / 200, ms_zc * 1000 / 200);
    cudaFreeHost(h); cudaFree(stage); cudaFree(img);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
