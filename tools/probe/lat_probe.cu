// Dependent-issue latencies of the operations the 6x6 pivoted solve is made of, one warp, clock64 around 256 dependent ops.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>

#define CHAIN 256
__global__ void probe(double* out, long long* cycles, double seed, int iseed) {
  __shared__ double s[64];
  __shared__ int si[64];
  const int lane = threadIdx.x & 31;
  s[lane] = seed + lane;
  s[lane + 32] = seed;
  si[lane] = (lane + 1) & 31;
  si[lane + 32] = lane;
  __syncwarp();
  double x = seed, y = seed * 0.5 + 1.0;
  long long t0, t1;
  int k = 0;
  // DADD
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) x = __dadd_rn(x, y);
  t1 = clock64(); cycles[k++] = t1 - t0;
  // DMUL
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) x = __dmul_rn(x, 1.0000001);
  t1 = clock64(); cycles[k++] = t1 - t0;
  // DFMA
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) x = __fma_rn(x, 1.0000001, y);
  t1 = clock64(); cycles[k++] = t1 - t0;
  // DDIV
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) x = __ddiv_rn(y, x) + 3.0;
  t1 = clock64(); cycles[k++] = t1 - t0;   // (includes one DADD)
  // DSETP + select
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) x = (x > y) ? x - 1.0 : x + 2.0;
  t1 = clock64(); cycles[k++] = t1 - t0;   // (includes DADDs)
  // REDUX max (u32)
  unsigned u = (unsigned)iseed + lane;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) u = __reduce_max_sync(0xffffffffu, u + lane);
  t1 = clock64(); cycles[k++] = t1 - t0;
  // SHFL of a double (2 x 32-bit)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1);
  t1 = clock64(); cycles[k++] = t1 - t0;
  // LDS pointer chase (int)
  int p = lane;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) p = si[p];
  t1 = clock64(); cycles[k++] = t1 - t0;
  // LDS double + STS double + syncwarp round trip
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) {
    s[lane] = x;
    __syncwarp();
    x = s[(lane + 1) & 31];
    __syncwarp();
  }
  t1 = clock64(); cycles[k++] = t1 - t0;
  // match_any
  unsigned m = u;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) m = __match_any_sync(0xffffffffu, (m + lane) & 7);
  t1 = clock64(); cycles[k++] = t1 - t0;
  // sqrt
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) x = sqrt(x) + 2.0;
  t1 = clock64(); cycles[k++] = t1 - t0;
  // FSEL chain (32-bit select on a runtime predicate)
  float f = (float)seed;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < CHAIN; ++i) f = (iseed == i) ? f + 1.0f : f * 1.5f;
  t1 = clock64(); cycles[k++] = t1 - t0;
  if (lane == 0) out[0] = x + u + p + m + f;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8 * 16);
  for (int rep = 0; rep < 2; ++rep) probe<<<1, 32>>>(out, cyc, 1.25, 7);
  long long h[16];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"DADD", "DMUL", "DFMA", "DDIV(+DADD)", "DSETP+sel(+DADD)", "REDUX.MAX(+IADD)", "SHFL f64", "LDS chase",
                         "STS+sync+LDS+sync f64", "MATCH.ANY(+2)", "DSQRT(+DADD)", "FADD/FMUL+sel"};
  for (int i = 0; i < 12; ++i) printf("%-24s %7.1f cycles per dependent op\n", names[i], (double)h[i] / CHAIN);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
