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
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

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


// Pure issue-rate probe: `n` MMAs (M=128, N, K=16 each) round-robin over NACC independent accumulators and 4 K chunks,
// one commit at the end; cycles measured by the issuing warp from the first issue to the commit's arrival.
// Everything is compile-time so the loop body is just the MMA instructions (warp-uniform code, one elected lane).
template <int N, int NACC, int ATMEM, int PADB = 0, int ALT = 0, int POLL = 0, int INIT = 0, int SAMEB = 0>
__global__ void __launch_bounds__(384) rate_kernel(int n, int two_issuers, long long* cycles, int reps = 1, int gap = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t lbo_b = (N / 8) * 128 + PADB;     // PADB: the kernel's de-conflicting pad between K groups
  constexpr int kA = 8 * 2048, kB = 8 * lbo_b;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kA + kB);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (kA + kB) / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (INIT) {                       // non-zero fp16 pairs in [-1, 1): is the MMA rate data dependent?
      const uint32_t hsh = (uint32_t)i * 2654435761u;
      const __half2 h2 = __floats2half2_rn(((int)(hsh >> 8 & 2047) - 1024) / 1024.f, ((int)(hsh >> 20 & 2047) - 1024) / 1024.f);
      v = *reinterpret_cast<const uint32_t*>(&h2);
    }
    ((uint32_t*)smem)[i] = v;
  }
  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (INIT && warp < 4) {             // A tiles in TMEM: same kind of data
    uint32_t r[8];
    for (int c = 0; c < 32; c += 8) {
      for (int i = 0; i < 8; ++i) r[i] = ((const uint32_t*)smem)[(tid * 32 + c + i) & 1023];
      tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 480 + c, r);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  if (INIT) { __syncthreads(); tc_fence_after(); }
  constexpr uint32_t idesc = umma_idesc_f16(M, N), idesc_h = umma_idesc_f16(M, N / 2);   // ALT: alternate N and N/2 (hi / lo products)
  if (warp == 0 || (two_issuers && warp == 1)) {
    const bool leader = elect_one();
    const uint64_t da0 = umma_smem_desc(smem_u32(smem), 2048, 128);
    const uint64_t db0 = umma_smem_desc(smem_u32(smem + kA), lbo_b, 128);
    const uint32_t acc0 = tmem + (warp == 1 ? 256u : 0u);
    const uint32_t at0 = tmem + 480;
    const long long t0 = clock64();
    long long t1 = t0, t2 = t0, issue = 0;
    for (int rep = 0; rep < reps; ++rep) {
      const long long ta = clock64();
      uint32_t accf = 0;
      for (int i = 0; i < n; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (leader) {
            if (ATMEM) umma_f16_ts(acc0 + (j % NACC) * N, at0 + (j & 3) * 8, db0 + (uint64_t)((((SAMEB ? j >> 1 : j) & 3) * 2 * lbo_b) >> 4), (ALT == 1 && (j & 1)) || (ALT == 2 && (j & 4)) ? idesc_h : idesc, accf);
            else umma_f16(acc0 + (j % NACC) * N, da0 + (uint64_t)(((j & 3) * 2 * 2048) >> 4), db0 + (uint64_t)((((SAMEB ? j >> 1 : j) & 3) * 2 * lbo_b) >> 4), (ALT == 1 && (j & 1)) || (ALT == 2 && (j & 4)) ? idesc_h : idesc, accf);
          }
        }
        accf = 1;
      }
      if (leader) umma_commit(smem_u32(&bars[warp]));
      __syncwarp();
      t1 = clock64();
      issue += t1 - ta;
      mbar_wait(smem_u32(&bars[warp]), rep & 1);
      t2 = clock64();
      if (gap) __nanosleep(gap);
    }
    if (reps > 1) { t1 = t0 + issue / reps; t2 = t0 + (t2 - t0) / reps; }
    if (lane_id() == 0 && warp == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
    if (POLL && lane_id() == 0 && warp == 0) mbar_arrive(smem_u32(&bars[1]));
  } else if (POLL && warp >= 4) {
    // bystanders polling an mbarrier in shared memory the way the kernel's waiting roles do:
    // POLL 1: every lane spins on try_wait; POLL 2: one lane spins, the rest park at __syncwarp; POLL 3: all lanes, nanosleep back-off
    const uint32_t bar = smem_u32(&bars[1]);
    if (POLL == 1) { while (!mbar_try_wait(bar, 0)) {} }
    else if (POLL == 2) { if (lane_id() == 0) { while (!mbar_try_wait(bar, 0)) {} } __syncwarp(); }
    else { while (!mbar_try_wait(bar, 0)) __nanosleep(64); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int NACC, int ATMEM, int PADB = 0, int ALT = 0, int POLL = 0, int INIT = 0, int SAMEB = 0>
static void rate(int n, int two, int reps = 1, int gap = 0) {
  long long* dC; cudaMalloc(&dC, 16);
  const size_t smem = 8 * 2048 + 8 * ((N / 8) * 128 + PADB) + 64;
  auto k = rate_kernel<N, NACC, ATMEM, PADB, ALT, POLL, INIT, SAMEB>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<1, POLL ? 384 : 128, smem>>>(n, two, dC, reps, gap);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); exit(2); }
  long long c[2]; cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
  printf("rate N=%3d nacc=%d A=%s padB=%d alt=%d poll=%d init=%d sameB=%d issuers=%d reps=%d gap=%d: %d MMAs  issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", N, NACC, ATMEM ? "tmem" : "smem", PADB, ALT, POLL, INIT, SAMEB,
         two ? 2 : 1, reps, gap, n, (double)c[0] / n, (double)c[1] / n);
  cudaFree(dC);
}


// Kernel-like burst: exactly the operand/accumulator geometry of snsde_tc_kernel<8,.,1> at c2 (N=16): per K chunk a
// hi MMA (N'=32, D at column DH) and a lo MMA (N=16, D at column DL), A-hi at AH + 8k, A-lo at AL + 8k, B chunk k.
__global__ void __launch_bounds__(512) klike_kernel(int nk, int DH, int DL, int AH, int AL, int reps, long long* cycles, int nwait, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t lbo_b = 4 * 128 + 16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16 * lbo_b);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16 * lbo_b / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003800u + i;
  if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), nwait > 0 ? nwait : 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp < 4) {
    uint32_t r[8];
    for (int c = 96; c < 512; c += 8) {
      for (int i = 0; i < 8; ++i) r[i] = 0x38003c00u + (uint32_t)(tid * 7 + c + i);
      tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  constexpr uint32_t idesc2 = umma_idesc_f16(M, 32), idesc1 = umma_idesc_f16(M, 16);
  if (warp == 0) {
    const bool leader = elect_one();
    const uint64_t db0 = umma_smem_desc(smem_u32(smem), lbo_b, 128);
    const uint64_t b_step = (uint64_t)((2 * lbo_b) >> 4);
    long long issue = 0, total = 0;
    for (int rep = 0; rep < reps; ++rep) {
      if (nwait > 0 && rep > 0) { mbar_wait(smem_u32(&bars[1]), (rep - 1) & 1); tc_fence_after(); }
      const long long ta = clock64();
      uint64_t db = db0;
      uint32_t ah = tmem + AH, al = tmem + AL, acc = 0;
#pragma unroll 4
      for (int k = 0; k < nk; ++k) {
        if (leader) {
          umma_f16_ts(tmem + DH, ah, db, idesc2, acc);
          umma_f16_ts(tmem + DL, al, db, idesc1, acc);
        }
        ah += 8; al += 8; db += b_step; acc = 1;
      }
      const long long tb = clock64();
      if (leader) umma_commit(smem_u32(&bars[0]));
      __syncwarp();
      mbar_wait(smem_u32(&bars[0]), rep & 1);
      const long long tc = clock64();
      issue += tb - ta; total += tc - ta;
    }
    if (lane_id() == 0 && blockIdx.x == 0) { cycles[0] = issue / reps; cycles[1] = total / reps; }
  } else if (warp >= 4 && warp < 4 + nwait) {
    // bystander "epilogue" warps: wait for the commit, (mode 1: read 8 accumulator columns, rewrite a B element), hand over
    for (int rep = 0; rep < reps; ++rep) {
      mbar_wait(smem_u32(&bars[0]), rep & 1);
      tc_fence_after();
      if (mode == 1) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + DH, v);
        tmem_ld_wait();
        ((__half*)smem)[tid] = __float2half(v[0] * 1e-9f + 1.f);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(smem_u32(&bars[1]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
static void klike(int nk, int DH, int DL, int AH, int AL, int nwait = 0, int mode = 0, int grid = 1) {
  long long* dC; cudaMalloc(&dC, 16);
  const size_t smem = 16 * (4 * 128 + 16) + 64;
  klike_kernel<<<grid, 128 + 32 * nwait, smem>>>(nk, DH, DL, AH, AL, 2000, dC, nwait, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); exit(2); }
  long long c[2]; cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
  printf("klike grid=%d nk=%d D(hi)=%d D(lo)=%d A(hi)=%d A(lo)=%d waiters=%d mode=%d: issue %lld, issue->complete %lld cycles per phase of %d MMAs\n", grid, nk, DH, DL, AH, AL, nwait, mode, c[0], c[1], 2 * nk);
  cudaFree(dC);
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

int main(int argc, char** argv) {
  srand(1);
  if (argc > 1) {
    // back-to-back issue rate: SS-form vs TS-form, small N (the numbers quoted in DESIGN.md / profiles/r1_umma_probe.txt)
    rate<16, 1, 0>(512, 0); rate<32, 1, 0>(512, 0); rate<64, 1, 0>(512, 0); rate<128, 2, 0>(512, 0); rate<256, 1, 0>(512, 0);
    rate<16, 1, 1>(512, 0); rate<16, 4, 1>(512, 0); rate<32, 4, 1>(512, 0); rate<64, 4, 1>(512, 0); rate<128, 2, 1>(512, 0);
    rate<32, 4, 1, 16, 1>(512, 0); rate<32, 4, 0, 16, 1>(512, 0);
    // bursts with a commit/wait round trip, as in a persistent per-layer pipeline
    for (int n : {8, 16, 32, 64}) rate<32, 4, 1, 16, 1>(n, 0, 200, 0);
    for (int n : {8, 16, 32, 64}) rate<32, 4, 0, 16, 1>(n, 0, 200, 0);
    // the solve kernel's operand geometry (c2: N = 16, 8 K chunks), alone / with waiting warps / on every SM
    klike(8, 48, 80, 96, 160);
    klike(8, 48, 80, 96, 160, 8, 1);
    for (int g : {1, 148}) klike(8, 48, 80, 96, 160, 8, 1, g);
    return 0;
  }
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
