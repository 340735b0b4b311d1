// Hardware probe for K1's tensor-core route (run on the GPU box; each mode in its own process because a faulting kernel
// poisons the context):
//   tc_probe mma <afmt> <bfmt> <N>     one tcgen05.mma kind::i8 (M=128, K=32) on 128B-swizzled K-major operands built by
//                                      the threads; checks D = A * B^T against the host (formats: 0 = u8, 1 = s8)
//   tc_probe tma <param|global> [fence] [cx] [cy] one TMA box (uint8, 128 B x 128 rows, 128B swizzle) from an image viewed as
//                                      super-rows of four rows, starting at an unaligned byte; tensor map passed as a
//                                      kernel parameter or read from global memory (optionally after fence.proxy.tensormap)
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tc_probe tools/micro/tc_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t ph) {
  for (uint32_t i = 0; i < 20000000u; ++i) if (mbar_try(b, ph)) return true;
  return false;
}
__device__ __forceinline__ uint64_t desc_k128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128) mma_probe(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, int n, uint32_t idesc,
                                                 int32_t* __restrict__ out, int* __restrict__ status) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                   // 128 rows x 128 B
  uint8_t* sB = smem + 16384;           // up to 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  // logical A[128][32], B[n][32] -> K positions 32..63 of the 128-byte rows (second K step: exercises the +32 B descriptor advance)
  for (int i = tid; i < 128 * 32; i += 128) {
    const int r = i >> 5, k = 32 + (i & 31);
    sA[(r >> 3) * 1024 + (r & 7) * 128 + (((k >> 4) ^ (r & 7)) << 4) + (k & 15)] = A[i];
  }
  for (int i = tid; i < n * 32; i += 128) {
    const int r = i >> 5, k = 32 + (i & 31);
    sB[(r >> 3) * 1024 + (r & 7) * 128 + (((k >> 4) ^ (r & 7)) << 4) + (k & 15)] = B[i];
  }
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (tid == 0) {
    const uint64_t da = desc_k128(smem_u32(sA)) + 2, db = desc_k128(smem_u32(sB)) + 2;      // K step 1 (+32 B)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem + 32), "l"(da), "l"(db), "r"(idesc), "r"(0) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  const bool ok = mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok) { if (tid == 0) *status = -1; }
  else {
    for (int c0 = 0; c0 < n; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 32 + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 16; ++i) out[tid * n + c0 + i] = (int32_t)r[i];
    }
    if (tid == 0) *status = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

template <bool PARAM>
__global__ void __launch_bounds__(128) tma_probe(const __grid_constant__ CUtensorMap tm_param, const CUtensorMap* tm_global, int fence,
                                                 int cx, int cy, uint8_t* __restrict__ out, int* __restrict__ status) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (tid == 0) {
    const void* tm = PARAM ? (const void*)&tm_param : (const void*)tm_global;
    if (!PARAM && fence) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tm) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(16384) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem)), "l"(tm), "r"(smem_u32(bar)), "r"(cx), "r"(cy) : "memory");
  }
  const bool ok = mbar_wait(bar, 0);
  if (!ok) { if (tid == 0) *status = -1; return; }
  for (int i = tid; i < 16384; i += 128) out[i] = smem[i];
  if (tid == 0) *status = 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage\n"); return 1; }
  int* d_status; CK(cudaMalloc(&d_status, 4)); CK(cudaMemset(d_status, 0, 4));
  if (!strcmp(argv[1], "mma")) {
    const int afmt = atoi(argv[2]), bfmt = atoi(argv[3]), n = atoi(argv[4]);
    std::vector<uint8_t> A(128 * 32), B(n * 32);
    srand(1);
    for (auto& v : A) v = (uint8_t)(rand() & 255);
    for (auto& v : B) v = (uint8_t)(rand() & 255);
    uint8_t *dA, *dB; int32_t* dO;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dO, 128 * n * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    const uint32_t idesc = (2u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    CK(cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024 + 64));
    mma_probe<<<1, 128, 16384 + 32768 + 1024 + 64>>>(dA, dB, n, idesc, dO, d_status);
    CK(cudaDeviceSynchronize());
    int st; CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    std::vector<int32_t> O(128 * n);
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int c = 0; c < n; ++c) {
        long s = 0;
        for (int k = 0; k < 32; ++k) {
          const int a = afmt ? (int)(int8_t)A[m * 32 + k] : (int)A[m * 32 + k];
          const int b = bfmt ? (int)(int8_t)B[c * 32 + k] : (int)B[c * 32 + k];
          s += a * b;
        }
        if (O[m * n + c] != (int32_t)s) { if (bad < 5) printf("  D[%d][%d] = %d, want %ld\n", m, c, O[m * n + c], s); ++bad; }
      }
    printf("mma afmt %d bfmt %d N %d: status %d, mismatches %ld of %d\n", afmt, bfmt, n, st, bad, 128 * n);
    return bad ? 1 : 0;
  }
  if (!strcmp(argv[1], "tma")) {
    const bool param = !strcmp(argv[2], "param");
    const int fence = argc > 3 ? atoi(argv[3]) : 0;
    const int H = 64, W = 68, pitch = 204;
    std::vector<uint8_t> img((size_t)H * pitch);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 7 + (i >> 8)) & 255);
    uint8_t *d_img, *d_out; CUtensorMap* d_tm;
    CK(cudaMalloc(&d_img, img.size())); CK(cudaMalloc(&d_out, 16384)); CK(cudaMalloc(&d_tm, sizeof(CUtensorMap)));
    CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(p);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)(3 * pitch + 3 * W), (cuuint64_t)(H / 4)};
    cuuint64_t strides[1] = {(cuuint64_t)(4 * pitch)};
    cuuint32_t box[2] = {128, 128}, estr[2] = {1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 3;
    CK(cudaMemcpy(d_tm, &tm, sizeof(tm), cudaMemcpyHostToDevice));
    const int cx = argc > 4 ? atoi(argv[4]) : 1 * pitch + 7, cy = argc > 5 ? atoi(argv[5]) : 2;
    printf("cx %d (mod 16 = %d) cy %d\n", cx, cx & 15, cy);
    if (param) { CK(cudaFuncSetAttribute(tma_probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024 + 64)); tma_probe<true><<<1, 128, 16384 + 1024 + 64>>>(tm, d_tm, fence, cx, cy, d_out, d_status); }
    else { CK(cudaFuncSetAttribute(tma_probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024 + 64)); tma_probe<false><<<1, 128, 16384 + 1024 + 64>>>(tm, d_tm, fence, cx, cy, d_out, d_status); }
    CK(cudaDeviceSynchronize());
    int st; CK(cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost));
    std::vector<uint8_t> out(16384);
    CK(cudaMemcpy(out.data(), d_out, 16384, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int rr = 0; rr < 128; ++rr)
      for (int k = 0; k < 128; ++k) {
        const long sr = cy + rr, x = cx + k;
        const uint8_t want = (sr < H / 4 && x < 3 * pitch + 3 * W) ? img[(size_t)sr * 4 * pitch + x] : 0;
        const uint8_t got = out[(rr >> 3) * 1024 + (rr & 7) * 128 + (((k >> 4) ^ (rr & 7)) << 4) + (k & 15)];
        if (got != want) { if (bad < 5) printf("  box[%d][%d] = %d, want %d\n", rr, k, got, want); ++bad; }
      }
    printf("tma %s fence %d: status %d, mismatches %ld of 16384\n", argv[2], fence, st, bad);
    return bad ? 1 : 0;
  }
  return 1;
}
