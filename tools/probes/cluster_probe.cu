// Probe: (1) can a cooperative launch carry cluster dimensions on this device, and how many clusters of size C are
// co-resident for a 512-thread / 170 KB CTA; (2) one-way latency of a flagged 64-bit word through DSMEM vs global (L2).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_rank(const void* p, uint32_t rank) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_remote(uint32_t addr, unsigned long long v) {
  asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_local(const void* p) {
  unsigned long long v; uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void st_g(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_g(const unsigned long long* p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

__global__ void __launch_bounds__(512, 1) probe(unsigned long long* g, long long* out, int iters) {
  __shared__ unsigned long long box[64];
  extern __shared__ unsigned char dyn[];
  if (threadIdx.x < 64) box[threadIdx.x] = 0;
  __syncthreads();
  cluster_sync();
  const uint32_t rank = cluster_rank(), csz = cluster_size();
  const int cluster_id = blockIdx.x / csz;
  if (cluster_id == 0 && threadIdx.x == 0 && csz >= 2) {
    // DSMEM ping-pong between rank 0 and the LAST rank of cluster 0
    const uint32_t peer = (rank == 0) ? csz - 1 : 0;
    if (rank == 0 || rank == csz - 1) {
      const uint32_t remote = map_rank(&box[0], peer);
      long long t0 = clock64();
      for (int i = 1; i <= iters; i++) {
        if (rank == 0) { st_remote(remote, (unsigned long long)i); while (ld_local(&box[0]) != (unsigned long long)i) {} }
        else { while (ld_local(&box[0]) != (unsigned long long)i) {} st_remote(remote, (unsigned long long)i); }
      }
      long long t1 = clock64();
      if (rank == 0) out[0] = (t1 - t0) / iters;  // round trip
    }
  }
  cluster_sync();
  // global ping-pong between block 0 and the last block of the grid
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && gridDim.x > 1) {
    const bool a = blockIdx.x == 0;
    long long t0 = clock64();
    for (int i = 1; i <= iters; i++) {
      if (a) { st_g(g, (unsigned long long)i); while (ld_g(g + 16) != (unsigned long long)i) {} }
      else { while (ld_g(g) != (unsigned long long)i) {} st_g(g + 16, (unsigned long long)i); }
    }
    long long t1 = clock64();
    if (a) out[1] = (t1 - t0) / iters;
  }
  // cluster barrier cost
  cluster_sync();
  if (cluster_id == 0) {
    long long t0 = clock64();
    for (int i = 0; i < 100; i++) cluster_sync();
    long long t1 = clock64();
    if (rank == 0 && threadIdx.x == 0) out[2] = (t1 - t0) / 100;
  }
  cluster_sync();
}

int main() {
  int dev = 0; cudaSetDevice(dev);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  printf("device %s SMs %d\n", p.name, p.multiProcessorCount);
  const size_t smem = 170 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  unsigned long long* g; long long* out;
  cudaMalloc(&g, 4096); cudaMalloc(&out, 64);
  for (int C : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    cfg.gridDim = dim3(C);
    int nClusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nClusters, probe, &cfg);
    printf("C=%2d maxActiveClusters=%d (%s) -> %d CTAs\n", C, nClusters, cudaGetErrorString(e), nClusters * C);
    if (nClusters < 1) continue;
    for (int coop = 1; coop >= 0; coop--) {
      cfg.gridDim = dim3(nClusters * C);
      cfg.numAttrs = coop ? 2 : 1;
      cudaMemset(g, 0, 4096); cudaMemset(out, 0, 64);
      int iters = 2000;
      e = cudaLaunchKernelEx(&cfg, probe, g, out, iters);
      cudaError_t e2 = cudaDeviceSynchronize();
      long long h[3] = {0, 0, 0};
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      printf("   coop=%d launch=%s sync=%s  dsmem round trip %lld cyc, global round trip %lld cyc, cluster barrier %lld cyc\n", coop,
             cudaGetErrorString(e), cudaGetErrorString(e2), h[0], h[1], h[2]);
      cudaGetLastError();
    }
  }
  return 0;
}
