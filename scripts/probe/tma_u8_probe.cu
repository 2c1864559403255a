// Probe (GPU box): TMA 2-D uint8 box loads with the descriptor in a kernel parameter vs in global memory,
// including negative / past-the-end coordinates (zero fill).  nvcc -arch=sm_100a tma_u8_probe.cu -o tma_u8_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../oadg_b200/csrc/oadg_tma.cuh"
using namespace oadg;

__global__ void probe(const __grid_constant__ CUtensorMap pmap, const void* gmap, int use_global, int x, int y, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    tma::mbar_init(&bar, 1);
    tma::mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const void* m = use_global ? gmap : (const void*)&pmap;
    if (use_global) tma::fence_tensormap_acquire(m);
    tma::mbar_expect_tx(&bar, 4096);
    tma::load_2d(sm, m, &bar, x, y);
  }
  tma::mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = sm[i];
}

int main() {
  const int Wb = 480, H = 96;
  std::vector<uint8_t> h(Wb * H);
  for (int i = 0; i < Wb * H; ++i) h[i] = (uint8_t)(i * 7 + (i / Wb));
  uint8_t *d, *out;
  cudaMalloc(&d, Wb * H);
  cudaMalloc(&out, 4096);
  cudaMemcpy(d, h.data(), Wb * H, cudaMemcpyHostToDevice);
  alignas(64) CUtensorMap m;
  int rc = tma::encode_u8_2d(&m, d, Wb, H, Wb, 256, 16);
  printf("encode rc %d\n", rc);
  void* gm;
  cudaMalloc(&gm, 128);
  cudaMemcpy(gm, &m, 128, cudaMemcpyHostToDevice);
  const int coords[][2] = {{0, 0}, {32, 5}, {-32, -4}, {400, 90}, {-16, 95}, {464, -15}, {33, 5}};
  for (int ug = 0; ug < 2; ++ug)
    for (auto& c : coords) {
      probe<<<1, 128, 4096>>>(m, gm, ug, c[0], c[1], out);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<uint8_t> o(4096);
      cudaMemcpy(o.data(), out, 4096, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int r = 0; r < 16; ++r)
        for (int k = 0; k < 256; ++k) {
          const int yy = c[1] + r, xx = c[0] + k;
          const uint8_t want = (yy >= 0 && yy < H && xx >= 0 && xx < Wb) ? h[yy * Wb + xx] : 0;
          bad += o[r * 256 + k] != want;
        }
      printf("desc %s coord (%d,%d): %s, %d mismatches\n", ug ? "global" : "param", c[0], c[1], cudaGetErrorString(e), bad);
      if (e != cudaSuccess) return 1;
    }
  return 0;
}
