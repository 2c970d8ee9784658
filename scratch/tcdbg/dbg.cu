// standalone debug harness for the tcgen05 GEMM building blocks
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#define ONMF_TC_DEBUG 1
#include "../../onmf_ontf_ndl_b200/csrc/gemm_tc.cu"
namespace onmf { thread_local char g_err[512] = ""; }
using namespace onmf;

// minimal kernel: one CTA, load one K-major A tile (128x32) and one K-major B tile (64 x 32) by TMA, one MMA k-step set, read back
__global__ void __launch_bounds__(128, 1) mini_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                      float* out, float* dumpA, int b_mn) {
  extern __shared__ uint8_t smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  using namespace tc;
  if (warp == 0 && lane == 0) {
    mbar_init(smem_u32(&full_bar), 1); mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t a_s = smem_u32(smem), b_s = a_s + 16384;
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(smem_u32(&full_bar), 16384 + 8192);
    tma_load_2d(a_s, &tmA, smem_u32(&full_bar), 0, 0);
    if (b_mn) { tma_load_2d(b_s, &tmB, smem_u32(&full_bar), 0, 0); tma_load_2d(b_s + 4096, &tmB, smem_u32(&full_bar), 32, 0); }
    else tma_load_2d(b_s, &tmB, smem_u32(&full_bar), 0, 0);
    mbar_wait(smem_u32(&full_bar), 0);
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t da = make_desc<false>(a_s + ks * 32);
      uint64_t db = b_mn ? make_desc<true>(b_s + ks * 1024) : make_desc<false>(b_s + ks * 32);
      umma_tf32(tmem_base, da, db, idesc, ks > 0);
    }
    umma_commit(smem_u32(&done_bar));
  }
  __syncthreads();
  mbar_wait(smem_u32(&done_bar), 0);
  tc_fence_after();
  for (int i = threadIdx.x; i < 4096 + 2048; i += blockDim.x) dumpA[i] = reinterpret_cast<float*>(smem)[i];
  uint32_t v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
}

int main() {
  const int M = 128, N = 64, K = 32;
  std::vector<float> A(M * K), B(N * K), Bt(K * N);
  for (int i = 0; i < M * K; ++i) A[i] = (float)((i * 7) % 13) / 4.f;        // exactly representable in tf32
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { B[n * K + k] = (float)((n * 3 + k * 5) % 11) / 8.f; Bt[k * N + n] = B[n * K + k]; }
  float *dA, *dB, *dBt, *dOut, *dDump;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dBt, Bt.size() * 4); cudaMalloc(&dOut, M * N * 4); cudaMalloc(&dDump, 6144 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dBt, Bt.data(), Bt.size() * 4, cudaMemcpyHostToDevice);
  for (int b_mn = 0; b_mn < 2; ++b_mn) {
    CUtensorMap mA, mB;
    if (tc::make_map(&mA, dA, K, M, K, 32, 128)) { printf("mapA fail %s\n", g_err); return 1; }
    if (b_mn) { if (tc::make_map(&mB, dBt, N, K, N, 32, 32, true)) { printf("mapB fail %s\n", g_err); return 1; } }
    else { if (tc::make_map(&mB, dB, K, N, K, 32, 64)) { printf("mapB fail %s\n", g_err); return 1; } }
    cudaMemset(dOut, 0xff, M * N * 4);
    cudaFuncSetAttribute(mini_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    mini_kernel<<<1, 128, 65536>>>(mA, mB, dOut, dDump, b_mn);
    cudaError_t e = cudaDeviceSynchronize();
    printf("b_mn=%d kernel: %s\n", b_mn, cudaGetErrorString(e));
    std::vector<float> out(M * N), dump(6144);
    cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(dump.data(), dDump, dump.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int bad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double r = 0; for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
      double err = fabs(out[m * N + n] - r); if (err > maxerr) maxerr = err; if (r > maxref) maxref = r; if (err > 1e-3) ++bad;
    }
    printf("  max err %.4g (max ref %.4g) bad %d   out[0..3]= %g %g %g %g\n", maxerr, maxref, bad, out[0], out[1], out[2], out[3]);
    printf("  smem A row0: "); for (int i = 0; i < 8; ++i) printf("%g ", dump[i]); printf(" | expected A row0: "); for (int i = 0; i < 8; ++i) printf("%g ", A[i]); printf("\n");
    printf("  smem A row1: "); for (int i = 0; i < 8; ++i) printf("%g ", dump[32 + i]); printf(" | expected (swizzled by 16B chunks) A row1 k4..: "); for (int i = 4; i < 12; ++i) printf("%g ", A[32 + i]); printf("\n");
    printf("  smem B first: "); for (int i = 0; i < 8; ++i) printf("%g ", dump[4096 + i]); printf("\n");
  }
  return 0;
}
