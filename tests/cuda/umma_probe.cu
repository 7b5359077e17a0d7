// Standalone hardware probe for the tcgen05 building blocks used by snsde_tc.cu:
//   * K-major no-swizzle smem descriptors (LBO/SBO meaning), fp16 kind::f16 MMA, M=128, small N
//   * TMEM 32x32b load mapping (thread <-> lane, register <-> column)
//   * split-fp16 (hi + 2^-11 * lo') operands giving fp32-class products
//   * round-trip latency of the dependent chain MMA -> commit -> ld -> st.shared -> fence -> MMA
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../../stable-neural-sdes_b200/csrc/snsde_tc_ptx.cuh"
using namespace snsde::ptx;

constexpr int M = 128;

struct Args {
  const __half* A; const __half* B; float* D;
  int N, K;            // N multiple of 8, K multiple of 16
  int a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  int iters;           // >1: latency loop
  long long* cycles;
  const __half* A_rows; // row-major [128][K] copy of A (for the A-in-TMEM variant)
  int nacc;             // independent accumulators the K chunks are spread over (dependency chains)
  int a_tmem;           // 1: stage A into TMEM with tcgen05.st and use the TS form of the MMA
};

__global__ void __launch_bounds__(192) probe_kernel(Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.a_bytes + ((p.b_bytes + 15) & ~15));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < p.a_bytes / 4; i += blockDim.x) ((uint32_t*)sA)[i] = ((const uint32_t*)p.A)[i];
  for (int i = tid; i < p.b_bytes / 4; i += blockDim.x) ((uint32_t*)sB)[i] = ((const uint32_t*)p.B)[i];
  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 128); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t a_col0 = 128;
  if (p.a_tmem) {
    if (warp < 4) {
      for (int kc = 0; kc < p.K / 16; ++kc) {
        const uint4* src = reinterpret_cast<const uint4*>(p.A_rows + (size_t)tid * p.K + kc * 16);
        const uint4 lo = src[0], hi = src[1];
        const uint32_t r[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + a_col0 + kc * 8, r);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t idesc = umma_idesc_f16(M, p.N);
  const uint32_t bar_acc = smem_u32(&bars[0]), bar_in = smem_u32(&bars[1]);
  long long t0 = clock64();
  if (warp == 4) {
    const bool leader = elect_one();
    {
      for (int it = 0; it < p.iters; ++it) {
        if (it > 0) { mbar_wait(bar_in, (it - 1) & 1); tc_fence_after(); }
        {
          // descriptors built once; the loop body is two 64-bit adds + the MMA (what snsde_tc.cu does)
          uint64_t da = umma_smem_desc(smem_u32(sA), p.a_lbo, p.a_sbo);
          uint64_t db = umma_smem_desc(smem_u32(sB), p.b_lbo, p.b_sbo);
          const uint64_t da_step = (uint64_t)((2 * p.a_lbo) >> 4), db_step = (uint64_t)((2 * p.b_lbo) >> 4);
          uint32_t at = tmem + a_col0;
          const int nk = p.K / 16;
          if (p.nacc == 1) {
#pragma unroll 8
            for (int kb = 0; kb < nk; ++kb) {
              if (leader) {
                if (p.a_tmem) umma_f16_ts(tmem, at, db, idesc, kb > 0);
                else umma_f16(tmem, da, db, idesc, kb > 0);
              }
              da += da_step; db += db_step; at += 8;
            }
          } else {
#pragma unroll 4
            for (int kb = 0; kb < nk; kb += 2) {
              if (leader) {
                if (p.a_tmem) { umma_f16_ts(tmem, at, db, idesc, kb > 0); umma_f16_ts(tmem + p.N, at + 8, db + db_step, idesc, kb > 0); }
                else { umma_f16(tmem, da, db, idesc, kb > 0); umma_f16(tmem + p.N, da + da_step, db + db_step, idesc, kb > 0); }
              }
              da += 2 * da_step; db += 2 * db_step; at += 16;
            }
          }
        }
        if (leader) umma_commit(bar_acc);
        __syncwarp();
      }
    }
  } else if (warp < 4) {
    float v[8];
    for (int it = 0; it < p.iters; ++it) {
      mbar_wait(bar_acc, it & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < p.N; c0 += 8) {
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int a = 1; a < p.nacc && a < p.K / 16; ++a) {
          float w[8];
          tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + a * p.N + c0, w);
          tmem_ld_wait();
          for (int i = 0; i < 8; ++i) v[i] += w[i];
        }
        tmem_ld_wait();
        if (it == p.iters - 1)
          for (int i = 0; i < 8; ++i) p.D[(size_t)tid * p.N + c0 + i] = v[i];
      }
      if (it < p.iters - 1) {
        // emulate the epilogue's operand write-back: one fp16 store per row, then hand over
        ((__half*)sB)[(tid & 63)] = ((__half*)sB)[(tid & 63)];
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(bar_in);
      }
    }
  }
  long long t1 = clock64();
  if (tid == 0 && p.cycles) *p.cycles = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// canonical K-major no-swizzle image of X[rows][K] (fp16): 8x(16 byte) core matrices
static void pack(const std::vector<float>& X, int rows, int K, uint32_t lbo, uint32_t sbo, std::vector<__half>& img, size_t bytes) {
  img.assign(bytes / 2, __float2half(0.f));
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      size_t off = (size_t)(r / 8) * sbo + (size_t)(k / 8) * lbo + (r % 8) * 16 + (k % 8) * 2;
      img[off / 2] = __float2half(X[(size_t)r * K + k]);
    }
}

static int g_a_tmem = 0, g_nacc = 1;
static double run(int N, int K, bool swap_desc, uint32_t b_pad, int iters, double* cyc, const std::vector<float>& A, const std::vector<float>& B,
                  std::vector<float>* out = nullptr) {
  const uint32_t a_sbo = 128, a_lbo = 16 * 128;                 // row groups contiguous, then K chunks
  const uint32_t b_sbo = 128, b_lbo = (N / 8) * 128 + b_pad;
  const size_t a_bytes = (size_t)(K / 8) * a_lbo, b_bytes = (size_t)(K / 8) * b_lbo;
  std::vector<__half> ia, ib;
  pack(A, M, K, a_lbo, a_sbo, ia, a_bytes);
  pack(B, N, K, b_lbo, b_sbo, ib, b_bytes);
  std::vector<__half> arow((size_t)M * K);
  for (size_t i = 0; i < arow.size(); ++i) arow[i] = __float2half(A[i]);
  __half* dAr; cudaMalloc(&dAr, arow.size() * 2); cudaMemcpy(dAr, arow.data(), arow.size() * 2, cudaMemcpyHostToDevice);
  __half *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, a_bytes); cudaMalloc(&dB, b_bytes); cudaMalloc(&dD, sizeof(float) * M * N); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, ia.data(), a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, ib.data(), b_bytes, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, sizeof(float) * M * N);
  Args p{dA, dB, dD, N, K, (int)a_bytes, (int)b_bytes,
         swap_desc ? a_sbo : a_lbo, swap_desc ? a_lbo : a_sbo, swap_desc ? b_sbo : b_lbo, swap_desc ? b_lbo : b_sbo, iters, dC, dAr, g_nacc, g_a_tmem};
  const size_t smem = a_bytes + ((b_bytes + 15) & ~15) + 64;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 192, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<float> D((size_t)M * N);
  cudaMemcpy(D.data(), dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
  long long c; cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
  if (cyc) *cyc = (double)c / iters;
  double err = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)__half2float(__float2half(A[(size_t)m * K + k])) * (double)__half2float(__float2half(B[(size_t)n * K + k]));
      err = fmax(err, fabs(ref - D[(size_t)m * N + n]));
    }
  if (out) *out = D;
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return err;
}

int main() {
  srand(1);
  for (g_a_tmem = 0; g_a_tmem < 2; ++g_a_tmem) {
  printf("==== A operand from %s ====\n", g_a_tmem ? "TMEM (tcgen05.st + TS-form MMA)" : "shared memory");
  for (int N : {16, 32, 64}) {
    for (int K : {16, 64, 128}) {
      std::vector<float> A((size_t)M * K), B((size_t)N * K);
      for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
      for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.f;
      for (int sw = 0; sw < 1; ++sw)
        for (uint32_t pad : {16u}) {
          double e = run(N, K, sw, pad, 1, nullptr, A, B);
          printf("N=%3d K=%3d swap_lbo_sbo=%d b_pad=%2u  max_abs_err=%.3e %s\n", N, K, sw, pad, e, e < 1e-3 ? "OK" : "WRONG");
        }
    }
  }
  // latency of the dependent chain (K=128+48 like layer 1 of c2 with 3 passes ~ 33 MMAs -> use K=528)
  for (g_nacc = 1; g_nacc <= 2; g_nacc *= 2)
  for (int N : {16, 32}) {
    for (int K : {32, 128, 256}) {
      std::vector<float> A((size_t)M * K), B((size_t)N * K);
      for (auto& x : A) x = (rand() % 2001 - 1000) / 1000.f;
      for (auto& x : B) x = (rand() % 2001 - 1000) / 1000.f;
      double cyc;
      double e = run(N, K, 0, 16, 2000, &cyc, A, B);
      printf("chain nacc=%d N=%d K=%d (%d MMAs): %.0f cycles per round trip  (err %.1e)\n", g_nacc, N, K, K / 16, cyc, e);
    }
  }
  g_nacc = 1;
  }
  return 0;
}
