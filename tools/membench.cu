// Microbenchmark: achievable read bandwidth of the converter access patterns (one CTA per SM-ish, 256 threads,
// 16 x LDG.128 per thread per item).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membench membench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// mode 0: direct pattern only      (128 rows x 128 B, row stride ldn*4, 4 planes)   row-major planes
// mode 1: direct + transposed, diagonal schedule (the real kernel's pattern)        row-major planes
// mode 2: contiguous 64 KB per item (tiled layout: [rowtile128][coltile32][4][128][32])
// mode 3: tiled, direct + transposed with the diagonal schedule: tile = [4 planes][32][32] = 16 KB, item = 4 tiles
template <int DEPTH>
__global__ void __launch_bounds__(256) k_read(const float* __restrict__ P, int n, int mode, float* out) {
  const int I = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ldn = n;
  const size_t pstride = (size_t)n * n;
  const float* G = P + (size_t)b * 4 * pstride;
  const int nkc = n / 32;
  const int items = (mode == 0 || mode == 2) ? nkc : 2 * nkc;
  float4 buf[DEPTH][16];
  float acc = 0.f;
  auto load = [&](int j, float4 (&bf)[16]) {
    if (mode == 0 || (mode == 1 && (j & 1) == 0)) {
      const int s = mode == 0 ? j : (j >> 1);
      const int kc = (4 * I + s) % nkc;
      const int gk = kc * 32 + 4 * (lane & 7);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int gi = I * 128 + 16 * warp + 4 * it + (lane >> 3);
          bf[q * 4 + it] = ldg_stream(G + q * pstride + (size_t)gi * ldn + gk);
        }
    } else if (mode == 1) {
      const int s = j >> 1;
      const int kc = ((4 * I - s) % nkc + nkc) % nkc;
      const int iq = 8 * (warp >> 1) + 2 * ((lane >> 3) & 3) + (lane & 1);
      const int kq = 4 * (warp & 1) + ((lane >> 1) & 3);
      const int gi = I * 128 + 4 * iq;
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int gk = kc * 32 + 4 * kq + kk;
          bf[q * 4 + kk] = ldg_stream(G + q * pstride + (size_t)gk * ldn + gi);
        }
    } else if (mode == 2) {
      const int kc = (4 * I + j) % nkc;
      const float* base = G + ((size_t)I * nkc + kc) * (4 * 128 * 32);
#pragma unroll
      for (int u = 0; u < 16; ++u) bf[u] = ldg_stream(base + (size_t)(u * 256 + tid) * 4);
    } else {
      // tiled 32x32: tile index (rt, ct) -> 16 KB block [4][32][32]; row tiles per dim = n/32
      const int s = j >> 1;
      const int nt = n / 32;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        int rt, ct;
        if ((j & 1) == 0) { rt = 4 * I + u; ct = (4 * I + s) % nkc; }
        else { rt = ((4 * I - s) % nkc + nkc) % nkc; ct = 4 * I + u; }
        const float* base = G + ((size_t)rt * nt + ct) * 4096;
#pragma unroll
        for (int v = 0; v < 4; ++v) bf[u * 4 + v] = ldg_stream(base + (size_t)(v * 256 + tid) * 4);
      }
    }
  };
#pragma unroll
  for (int dpt = 0; dpt < DEPTH; ++dpt) load(dpt, buf[dpt]);
  for (int j = 0; j < items; j += DEPTH) {
#pragma unroll
    for (int dpt = 0; dpt < DEPTH; ++dpt) {
#pragma unroll
      for (int u = 0; u < 16; ++u) acc += buf[dpt][u].x + buf[dpt][u].y + buf[dpt][u].z + buf[dpt][u].w;
      if (j + dpt + DEPTH < items) load(j + dpt + DEPTH, buf[dpt]);
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 4096;
  const int B = argc > 2 ? atoi(argv[2]) : 4;
  const size_t bytes = (size_t)B * 4 * n * n * 4;
  float *P, *out;
  cudaMalloc(&P, bytes);
  cudaMalloc(&out, 4);
  cudaMemset(P, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[4] = {"rowmajor direct-only", "rowmajor direct+transposed (diag)", "contiguous 64KB items", "tiled32 direct+transposed (diag)"};
  for (int depth = 1; depth <= 2; ++depth)
    for (int mode = 0; mode < 4; ++mode) {
      dim3 grid(n / 128, B);
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (depth == 1) k_read<1><<<grid, 256>>>(P, n, mode, out);
        else k_read<2><<<grid, 256>>>(P, n, mode, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double unique = (double)bytes;
      const double moved = (mode == 1 || mode == 3) ? 2.0 * unique : unique;
      printf("n=%d B=%d depth=%d %-36s %8.1f us  unique %6.0f GB/s  L2->SM %6.0f GB/s  (%s)\n", n, B, depth, names[mode], best * 1e3,
             unique / best / 1e6, moved / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
