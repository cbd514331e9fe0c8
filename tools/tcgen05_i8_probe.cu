// tools/tcgen05_i8_probe.cu — smallest possible check of the tcgen05 building blocks used by csrc/npw_ozaki_i8.cu, for
// the first GPU minutes of round 2 (never run in round 1).  One CTA:
//   1. the threads write an int8 A tile (128 x KB bytes) and B tile (64 x KB bytes) into shared memory in the K-major
//      SWIZZLE_128B layout BY HAND (no TMA): byte (row, k) lives at  row*128 + ((k/16) ^ (row%8))*16 + k%16  inside each
//      1024-byte group of 8 rows — the layout TMA produces with CU_TENSOR_MAP_SWIZZLE_128B for a 128-byte inner box;
//   2. one lane issues KB/32 tcgen05.mma kind::i8 (M=128, N=64) with the same descriptors as the product kernel;
//   3. tcgen05.commit -> mbarrier, four warps tcgen05.ld the 128 x 64 int32 accumulator and compare with a host product.
// Build + run (B200 only):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tcgen05_i8_probe tools/tcgen05_i8_probe.cu && timeout 60 tools/tcgen05_i8_probe
// Descriptor fields can be overridden from the command line to try alternative encodings in separate processes (a bad
// encoding may poison the context):   tcgen05_i8_probe [lbo=1] [sbo=64] [layout=2] [version=1]
// (defaults = what csrc/npw_ozaki_i8.cu uses; tools/round2_first_call.sh sweeps a few, each under its own timeout).
// Prints "PROBE OK" or the first mismatches (which tell descriptor / lane-mapping errors apart: a transposed or
// row-permuted result points at the descriptors, a quadrant shift at the tcgen05.ld lane field).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int M = 128, N = 64, KB = 128, UK = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

struct DescCfg {
  uint32_t lbo, sbo, layout, version;
};

__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr, DescCfg c) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(c.lbo & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(c.sbo & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(c.version & 3u) << 46;
  d |= static_cast<uint64_t>(c.layout & 7u) << 61;
  return d;
}

__global__ void __launch_bounds__(192) probe(const int8_t* A, const int8_t* B, int32_t* D, DescCfg cfg) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;                 // 128 rows x 128 B
  uint8_t* sb = smem + M * KB;        // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + N * KB);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < M * KB; e += blockDim.x) {
    const int r = e / KB, k = e % KB;
    sa[(r / 8) * 1024 + (r % 8) * 128 + (((k / 16) ^ (r % 8)) * 16) + k % 16] = static_cast<uint8_t>(A[e]);
  }
  for (int e = threadIdx.x; e < N * KB; e += blockDim.x) {
    const int r = e / KB, k = e % KB;
    sb[(r / 8) * 1024 + (r % 8) * 128 + (((k / 16) ^ (r % 8)) * 16) + k % 16] = static_cast<uint8_t>(B[e]);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // generic-proxy writes to shared memory must be made visible to the async proxy that tcgen05.mma reads through
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tptr;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
    for (int k4 = 0; k4 < KB / UK; ++k4) {
      const uint64_t ad = umma_desc_k128(smem_u32(sa) + k4 * UK, cfg), bd = umma_desc_k128(smem_u32(sb) + k4 * UK, cfg);
      const uint32_t acc = k4 > 0 ? 1u : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tbase), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  if (warp >= 2) {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      int32_t v[16];
      const uint32_t taddr = tbase + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64u) : "memory");
  }
}

int main(int argc, char** argv) {
  DescCfg cfg = {1u, 64u, 2u, 1u};
  if (argc > 1) cfg.lbo = static_cast<uint32_t>(atoi(argv[1]));
  if (argc > 2) cfg.sbo = static_cast<uint32_t>(atoi(argv[2]));
  if (argc > 3) cfg.layout = static_cast<uint32_t>(atoi(argv[3]));
  if (argc > 4) cfg.version = static_cast<uint32_t>(atoi(argv[4]));
  printf("descriptor: lbo=%u sbo=%u layout=%u version=%u\n", cfg.lbo, cfg.sbo, cfg.layout, cfg.version);
  int8_t *hA = (int8_t*)malloc(M * KB), *hB = (int8_t*)malloc(N * KB);
  int32_t* hD = (int32_t*)malloc(M * N * 4);
  srand(1);
  for (int i = 0; i < M * KB; ++i) hA[i] = (int8_t)(rand() % 129 - 64);
  for (int i = 0; i < N * KB; ++i) hB[i] = (int8_t)(rand() % 129 - 64);
  int8_t *dA, *dB;
  int32_t* dD;
  cudaMalloc(&dA, M * KB); cudaMalloc(&dB, N * KB); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * KB, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * KB, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, M * N * 4);
  const int smem = M * KB + N * KB + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 192, smem>>>(dA, dB, dD, cfg);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("PROBE CUDA ERROR: %s\n", cudaGetErrorString(e)); return 2; }
  cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < M && bad < 10; ++i)
    for (int j = 0; j < N && bad < 10; ++j) {
      int32_t ref = 0;
      for (int k = 0; k < KB; ++k) ref += (int32_t)hA[i * KB + k] * (int32_t)hB[j * KB + k];
      if (ref != hD[i * N + j]) { printf("mismatch D[%d][%d] = %d, expected %d\n", i, j, hD[i * N + j], ref); ++bad; }
    }
  printf(bad ? "PROBE FAILED\n" : "PROBE OK: 128x64x128 int8 tcgen05.mma + tcgen05.ld match the host product\n");
  return bad ? 1 : 0;
}
