// membench.cu -- what does a B200 SM-side access pattern cost?  Diagnostic for the GroupNorm kernels, which sit
// at ~3.1 TB/s whatever the CTA shape while the driver's copy peak is 6.5 TB/s and K1's bulk stores reach 5.2.
// Each pattern moves the same buffers; prints GB/s (bytes read + written).  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/membench tools/membench.cu && gpurun_out/membench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- A: LDG read, U independent 16-byte loads in flight per thread ---------------------------------------
template <int U>
__global__ void read_ldg(const uint4* __restrict__ src, size_t n, uint4* sink) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) *sink = acc;
}
// ---- B: classic copy, U loads in flight ------------------------------------------------------------------
template <int U>
__global__ void copy_ldg(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) dst[i + u * stride] = v[u];
  }
}
// ---- C: slab pattern of the cluster GroupNorm kernels: a CTA bulk-loads ONE contiguous slab, waits for all of
//         it, then writes it back (mode 0: 16-byte st.global from ld.shared, mode 1: bulk stores), one slab per CTA
__global__ void slab_copy(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, uint32_t slab, int mode) {
  extern __shared__ __align__(128) unsigned char sm[];
  const uint32_t bar = smem_u32(sm), buf = smem_u32(sm + 128);
  const size_t off = (size_t)blockIdx.x * slab;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async();
    mbar_expect(bar, slab);
    for (uint32_t o = 0; o < slab; o += 32768) bulk_g2s(buf + o, src + off + o, slab - o < 32768 ? slab - o : 32768, bar);
  }
  __syncthreads();
  mbar_wait(bar, 0);
  if (mode == 0) {
    const uint4* s = reinterpret_cast<const uint4*>(sm + 128);
    uint4* d = reinterpret_cast<uint4*>(dst + off);
    for (uint32_t i = threadIdx.x; i < slab / 16; i += blockDim.x) d[i] = s[i];
  } else if (threadIdx.x == 0) {
    for (uint32_t o = 0; o < slab; o += 32768) bulk_s2g(dst + off + o, buf + o, slab - o < 32768 ? slab - o : 32768);
    bulk_commit();
    bulk_wait_read0();
  }
}
// ---- D: persistent ring: each CTA walks its slabs with a STAGES-deep ring of `chunk`-byte bulk loads and writes
//         each chunk back as soon as it lands (mode 0 st.global / mode 1 bulk store / mode 2 read only)
template <int STAGES>
__global__ void ring_copy(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, size_t total, uint32_t chunk, int mode) {
  extern __shared__ __align__(128) unsigned char sm[];
  const uint32_t bars = smem_u32(sm), buf = smem_u32(sm + 128);
  const size_t nchunks = total / chunk;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async();
  }
  __syncthreads();
  size_t issued = blockIdx.x, done = blockIdx.x;
  int is = 0, ds = 0;
  uint32_t dphase = 0;
  if (threadIdx.x == 0)
    for (; is < STAGES && issued < nchunks; ++is, issued += gridDim.x) {
      mbar_expect(bars + 8 * is, chunk);
      bulk_g2s(buf + is * chunk, src + issued * chunk, chunk, bars + 8 * is);
    }
  for (; done < nchunks; done += gridDim.x) {
    mbar_wait(bars + 8 * ds, dphase);
    if (mode == 0) {
      const uint4* s = reinterpret_cast<const uint4*>(sm + 128 + (size_t)ds * chunk);
      uint4* d = reinterpret_cast<uint4*>(dst + done * chunk);
      for (uint32_t i = threadIdx.x; i < chunk / 16; i += blockDim.x) d[i] = s[i];
    } else if (mode == 1 && threadIdx.x == 0) {
      bulk_s2g(dst + done * chunk, buf + ds * chunk, chunk);
      bulk_commit();
      bulk_wait_read0();
    }
    __syncthreads();                       // everyone is done with this stage
    if (threadIdx.x == 0) {
      const size_t nxt = done + (size_t)STAGES * gridDim.x;
      if (nxt < nchunks) {
        fence_async();
        mbar_expect(bars + 8 * ds, chunk);
        bulk_g2s(buf + ds * chunk, src + nxt * chunk, chunk, bars + 8 * ds);
      }
    }
    if (++ds == STAGES) { ds = 0; dphase ^= 1u; }
  }
}

template <typename F>
static float time_ms(F&& launch, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  CK(cudaGetLastError());
  float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)1 << 30;                 // 1 GiB in, 1 GiB out (>> L2)
  unsigned char *src, *dst; uint4* sink;
  CK(cudaMalloc(&src, bytes)); CK(cudaMalloc(&dst, bytes)); CK(cudaMalloc(&sink, 16));
  CK(cudaMemset(src, 1, bytes)); CK(cudaMemset(dst, 0, bytes));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t n16 = bytes / 16;
  auto gbs = [&](double moved, float ms) { return moved / ms / 1e6; };
  printf("SMs %d, buffer %zu MiB\n", sms, bytes >> 20);
  { float ms = time_ms([&] { CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice)); }); printf("cudaMemcpy D2D                         %8.0f GB/s\n", gbs(2.0 * bytes, ms)); }
  { float ms = time_ms([&] { CK(cudaMemsetAsync(dst, 0, bytes)); }); printf("cudaMemset                             %8.0f GB/s (write only)\n", gbs(1.0 * bytes, ms)); }
#define RD(U, CTAS) { float ms = time_ms([&] { read_ldg<U><<<sms * CTAS, 256>>>((const uint4*)src, n16, sink); }); printf("read  ldg  U=%d  %2d CTA/SM x256          %8.0f GB/s (read only)\n", U, CTAS, gbs(1.0 * bytes, ms)); }
  RD(1, 8) RD(4, 8) RD(8, 8) RD(8, 4) RD(8, 2)
#define CP(U, CTAS) { float ms = time_ms([&] { copy_ldg<U><<<sms * CTAS, 256>>>((const uint4*)src, (uint4*)dst, n16); }); printf("copy  ldg  U=%d  %2d CTA/SM x256          %8.0f GB/s\n", U, CTAS, gbs(2.0 * bytes, ms)); }
  CP(1, 8) CP(4, 8) CP(8, 8) CP(8, 4)
  for (int mode = 0; mode < 2; ++mode)
    for (uint32_t slab : {49152u, 98304u, 200704u}) {
      for (int threads : {256, 512}) {
        const size_t smem = slab + 128;
        CK(cudaFuncSetAttribute(slab_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = (int)(bytes / slab);
        float ms = time_ms([&] { slab_copy<<<grid, threads, smem>>>(src, dst, slab, mode); });
        printf("slab  %6u B x%3d thr  %s        %8.0f GB/s\n", slab, threads, mode ? "bulk store" : "st.global ", gbs(2.0 * (double)grid * slab, ms));
      }
    }
  for (int mode = 0; mode < 3; ++mode)
    for (uint32_t chunk : {16384u, 32768u}) {
      for (int ctas : {1, 2}) {
        const size_t smem = (size_t)4 * chunk + 128;
        CK(cudaFuncSetAttribute(ring_copy<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float ms = time_ms([&] { ring_copy<4><<<sms * ctas, 256, smem>>>(src, dst, bytes, chunk, mode); });
        const char* nm = mode == 0 ? "st.global " : (mode == 1 ? "bulk store" : "read only ");
        printf("ring4 %6u B  %d CTA/SM  %s        %8.0f GB/s\n", chunk, ctas, nm, gbs((mode == 2 ? 1.0 : 2.0) * bytes, ms));
      }
    }
  return 0;
}
