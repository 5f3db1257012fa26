// Hardware probe (development tool, not part of the library): which shared-memory matrix descriptors does
// tcgen05.mma accept beyond the canonical ones?
//   P1  no-swizzle K-major A operand with OVERLAPPING rows: row stride 16 B, K-chunk stride (LBO) 16 B,
//       8-row-group stride (SBO) 128 B  ->  A[i][k] = X[8*i + k]   (sliding window over a flat array)
//   P2  128B-swizzled K-major A operand whose start address is an arbitrary multiple of 128 B (row shift r0),
//       with and without the descriptor's base_offset field
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o /tmp/probe tools/probe_umma_desc.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../implementation-phd-lab-vision_b200/csrc/ptx_sm100.cuh"

using namespace phdfxk;

constexpr int M = 128, N = 64;

struct Args {
  int mode;        // 1 = P1, 2 = P2
  int r0;          // P2: row shift
  int use_base_offset;
  int ksteps;      // number of K=16 MMAs
};

__global__ void probe_kernel(const __nv_bfloat16* __restrict__ gA, int a_bytes, const __nv_bfloat16* __restrict__ gB,
                             int b_bytes, float* __restrict__ out, Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // up to 64 KB
  uint8_t* sB = smem + 65536;      // 16 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(gA)[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(gB)[i];
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&tmem_ptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N);
    for (int k = 0; k < a.ksteps; ++k) {
      uint64_t adesc, bdesc;
      if (a.mode == 1) {
        // A: no swizzle, LBO = 16 B, SBO = 128 B, start advances 32 B per K=16 step
        const uint32_t addr = smem_u32(sA) + k * 32;
        adesc = (uint64_t((addr & 0x3FFFF) >> 4)) | (uint64_t(16 >> 4) << 16) | (uint64_t(128 >> 4) << 32) | (1ull << 46);
        // B: no swizzle canonical: [kchunk][n][8 elem]: LBO = N*16 B, SBO = 128 B
        const uint32_t baddr = smem_u32(sB) + k * 2 * (N * 16);
        bdesc = (uint64_t((baddr & 0x3FFFF) >> 4)) | (uint64_t((N * 16) >> 4) << 16) | (uint64_t(128 >> 4) << 32) |
                (1ull << 46);
      } else {
        const uint32_t addr = smem_u32(sA) + a.r0 * 128 + k * 32;
        adesc = make_kmajor_desc(addr, 128);
        if (a.use_base_offset) adesc |= (uint64_t(a.r0 & 7) << 49);
        bdesc = make_kmajor_desc(smem_u32(sB) + k * 32, 128);
      }
      umma_bf16(tmem, adesc, bdesc, idesc, k != 0);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < 4) {
    for (int c = 0; c < N / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + (uint32_t(warp * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c * 32 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  float* d_out;
  cudaMalloc(&d_out, M * N * 4);
  std::vector<float> out(M * N);
  srand(1);
  auto rnd = []() { return bf((rand() % 2001 - 1000) / 500.0f); };

  // ---------------- P1
  for (int ksteps : {1, 2}) {
    const int K = 16 * ksteps;
    std::vector<float> X(8 * M + 64);
    for (auto& v : X) v = rnd();
    std::vector<float> B(N * K);
    for (auto& v : B) v = rnd();
    std::vector<__nv_bfloat16> hX(X.size()), hB(N * K);
    for (size_t i = 0; i < X.size(); ++i) hX[i] = __float2bfloat16(X[i]);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) hB[(k / 8) * (N * 8) + n * 8 + (k % 8)] = __float2bfloat16(B[n * K + k]);
    __nv_bfloat16 *dA, *dB;
    const int a_bytes = ((int)hX.size() * 2 + 15) / 16 * 16, b_bytes = N * K * 2;
    cudaMalloc(&dA, a_bytes + 16);
    cudaMalloc(&dB, b_bytes);
    cudaMemcpy(dA, hX.data(), hX.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), b_bytes, cudaMemcpyHostToDevice);
    Args a{1, 0, 0, ksteps};
    probe_kernel<<<1, 128, 100 * 1024>>>(dA, a_bytes, dB, b_bytes, d_out, a);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), d_out, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int i = 0; i < M; ++i)
      for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)X[8 * i + k] * B[n * K + k];
        maxerr = fmax(maxerr, fabs(ref - out[i * N + n]));
      }
    printf("P1 overlapping no-swizzle rows, K=%d: %s  max_err %.4g  (%s)\n", K, maxerr < 1e-2 ? "PASS" : "FAIL", maxerr,
           cudaGetErrorString(e));
    cudaFree(dA);
    cudaFree(dB);
  }

  // ---------------- P2
  const int ROWS = 256, K = 64;
  std::vector<float> A(ROWS * K), B(N * K);
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  // swizzled images: physical chunk = logical chunk ^ (row & 7), tile base 1024-aligned
  std::vector<__nv_bfloat16> hA(ROWS * K), hB(N * K);
  for (int r = 0; r < ROWS; ++r)
    for (int k = 0; k < K; ++k) hA[r * 64 + (((k / 8) ^ (r & 7)) * 8) + (k % 8)] = __float2bfloat16(A[r * K + k]);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) hB[n * 64 + (((k / 8) ^ (n & 7)) * 8) + (k % 8)] = __float2bfloat16(B[n * K + k]);
  __nv_bfloat16 *dA, *dB;
  cudaMalloc(&dA, ROWS * K * 2);
  cudaMalloc(&dB, N * K * 2);
  cudaMemcpy(dA, hA.data(), ROWS * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  for (int ubo : {0, 1})
    for (int r0 : {0, 8, 1, 3, 5, 58, 59, 60, 117}) {
      Args a{2, r0, ubo, 4};
      probe_kernel<<<1, 128, 100 * 1024>>>(dA, ROWS * K * 2, dB, N * K * 2, d_out, a);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(out.data(), d_out, M * N * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int i = 0; i < M; ++i)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)A[(r0 + i) * K + k] * B[n * K + k];
          maxerr = fmax(maxerr, fabs(ref - out[i * N + n]));
        }
      printf("P2 sw128 row shift r0=%3d base_offset=%d: %s  max_err %.4g  (%s)\n", r0, ubo,
             maxerr < 2e-2 ? "PASS" : "FAIL", maxerr, cudaGetErrorString(e));
    }
  return 0;
}
