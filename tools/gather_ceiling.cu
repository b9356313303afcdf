// gather_ceiling.cu — what can a B200 deliver for random gathers of rowB-byte rows?  (measurement tool, not product)
// Times three ways of fetching `nIdx` rows of `rowB` bytes out of a [R][rowB] table, with nothing else in the loop:
//   ldg   : LDG.128, rowB/16 lanes per row, U independent rows in flight per lane, full occupancy
//   tma4  : cp.async.bulk.tensor ... tile::gather4 into a per-warp shared-memory ring (depth D stages), elected-lane issue
//   bulk  : cp.async.bulk per row into the same ring
// for index streams with different L2 behaviour (table 32 MB: all L2 hits; 1 GB uniform: ~all DRAM; 50/50 hot/cold).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_ceiling tools/gather_ceiling.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void k_fill_idx(uint32_t* idx, uint64_t n, uint32_t R, uint32_t hotR, int mode, uint32_t seed) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t h = hash32((uint32_t)i * 2654435761u + seed);
    uint32_t h2 = hash32(h + 0x9e3779b9u);
    uint32_t v;
    if (mode == 0) v = h % R;                       // uniform over the table
    else v = (h2 % 100u < (uint32_t)mode) ? (h % hotR) : (h % R);   // mode % of the accesses go to a hot set of hotR rows
    idx[i] = v;
  }
}

template <int LANES, int U>
__global__ void __launch_bounds__(256) k_ldg(const float4* __restrict__ tab, const uint32_t* __restrict__ idx, uint64_t nIdx,
                                             uint32_t rowQ, float* sink) {
  const int lane = threadIdx.x % LANES;
  const uint64_t grp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) / LANES;
  const uint64_t ngrp = (uint64_t)gridDim.x * blockDim.x / LANES;
  const uint64_t per = (nIdx + ngrp - 1) / ngrp;
  uint64_t b = grp * per, e = b + per < nIdx ? b + per : nIdx;
  float4 acc = make_float4(0, 0, 0, 0);
  for (uint64_t i = b; i + U <= e; i += U) {
    uint32_t s[U];
#pragma unroll
    for (int u = 0; u < U; u++) s[u] = __ldg(idx + i + u);
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = __ldg(tab + (size_t)s[u] * rowQ + lane);
#pragma unroll
    for (int u = 0; u < U; u++) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t par) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar), "r"(par) : "memory");
}
__device__ __forceinline__ void gather4(uint32_t dst, const CUtensorMap* m, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"((uint64_t)m), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one ring per warp: D stages of G4 gather4s (4*G4 rows per stage); lane 0 issues, the warp waits, optionally reads
template <int MODE, int READ>
__global__ void __launch_bounds__(128) k_tma(const __grid_constant__ CUtensorMap tmap, const char* tab, const uint32_t* __restrict__ idx,
                                             uint64_t nIdx, uint32_t rowB, int D, int G4, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t stageB = (uint32_t)G4 * 4u * rowB;
  const uint32_t ringB = (uint32_t)D * stageB;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t ring = s0 + wi * ringB;
  const uint32_t bars = s0 + nw * ringB + wi * D * 8;
  const uint64_t gw = blockIdx.x * (uint64_t)nw + wi, ngw = (uint64_t)gridDim.x * nw;
  const uint64_t rowsPerStage = 4ull * G4;
  const uint64_t nStagesAll = nIdx / rowsPerStage;
  const uint64_t per = (nStagesAll + ngw - 1) / ngw;
  const uint64_t sb = gw * per, se = sb + per < nStagesAll ? sb + per : nStagesAll;
  if (sb >= se) return;
  const uint64_t n = se - sb;
  if (lane == 0) { for (int s = 0; s < D; s++) mbar_init(bars + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  auto issue = [&](uint64_t k, uint32_t slot) {
    if (k < n && lane == 0) {
      const uint32_t* ip = idx + (sb + k) * rowsPerStage;
      const uint32_t bar = bars + slot * 8, dst = ring + slot * stageB;
      mbar_expect(bar, stageB);
      for (int g = 0; g < G4; g++) {
        uint32_t i0 = __ldg(ip + 4 * g), i1 = __ldg(ip + 4 * g + 1), i2 = __ldg(ip + 4 * g + 2), i3 = __ldg(ip + 4 * g + 3);
        if (MODE == 0) gather4(dst + g * 4 * rowB, &tmap, i0, i1, i2, i3, bar);
        else {
          bulk(dst + (g * 4 + 0) * rowB, tab + (size_t)i0 * rowB, rowB, bar);
          bulk(dst + (g * 4 + 1) * rowB, tab + (size_t)i1 * rowB, rowB, bar);
          bulk(dst + (g * 4 + 2) * rowB, tab + (size_t)i2 * rowB, rowB, bar);
          bulk(dst + (g * 4 + 3) * rowB, tab + (size_t)i3 * rowB, rowB, bar);
        }
      }
    }
  };
  for (int k = 0; k < D; k++) issue(k, k);
  uint32_t slot = 0, par = 0;
  float4 acc = make_float4(0, 0, 0, 0);
  for (uint64_t k = 0; k < n; k++) {
    mbar_wait(bars + slot * 8, par);
    if (READ) {
      for (uint32_t o = lane * 16; o < stageB; o += 512) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ring + slot * stageB + o));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    __syncwarp();
    issue(k + D, slot);
    if (++slot == (uint32_t)D) { slot = 0; par ^= 1; }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}


// Producer / consumer ring: one producer warp streams through the CTA's contiguous index range (indices
// prefetched one 128-row block ahead with one coalesced LDG.128 per lane; lane l issues gather4 #l of the
// block from its own registers), chunk slots of 64 rows complete on a `full` mbarrier, W consumer warps
// take chunks round-robin, optionally read them, and release the slot through an `empty` mbarrier.
template <int READ>
__global__ void __launch_bounds__(288) k_ring(const __grid_constant__ CUtensorMap tmap, const uint32_t* __restrict__ idx, uint64_t nChunks,
                                              uint32_t rowB, int R, int W, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const uint32_t chunkB = 64u * rowB;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t fullB = s0 + (uint32_t)R * chunkB, emptyB = fullB + (uint32_t)R * 8;
  uint64_t per = (nChunks + gridDim.x - 1) / gridDim.x; per = (per + 1) & ~1ull;
  const uint64_t c0 = blockIdx.x * per, c1 = c0 + per < nChunks ? c0 + per : nChunks;
  if (threadIdx.x == 0) {
    for (int s = 0; s < R; s++) { mbar_init(fullB + 8 * s, 1); mbar_init(emptyB + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (c0 >= c1) return;
  if (wi == 0) {
    const uint4* ip = reinterpret_cast<const uint4*>(idx) + c0 * 16 + lane;
    uint4 nxt = __ldg(ip);
    for (uint64_t c = c0; c < c1; c += 2) {
      const uint4 cur = nxt;
      if (c + 2 < c1) nxt = __ldg(ip + (c + 2 - c0) * 16);
      const uint64_t mine = c + (lane >> 4);
      if (mine < c1) {
        const uint64_t k = mine - c0;
        const uint32_t slot = (uint32_t)(k % (uint64_t)R), par = (uint32_t)((k / (uint64_t)R) & 1);
        mbar_wait(emptyB + 8 * slot, par ^ 1u);
        if ((lane & 15) == 0) mbar_expect(fullB + 8 * slot, chunkB);
        gather4(s0 + slot * chunkB + (lane & 15) * 4 * rowB, &tmap, cur.x, cur.y, cur.z, cur.w, fullB + 8 * slot);
      }
    }
  } else {
    float4 acc = make_float4(0, 0, 0, 0);
    for (uint64_t c = c0 + (wi - 1); c < c1; c += W) {
      const uint64_t k = c - c0;
      const uint32_t slot = (uint32_t)(k % (uint64_t)R), par = (uint32_t)((k / (uint64_t)R) & 1);
      mbar_wait(fullB + 8 * slot, par);
      if (READ) {
        for (uint32_t o = lane * 16; o < chunkB; o += 512) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(s0 + slot * chunkB + o));
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(emptyB + 8 * slot) : "memory");
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const uint32_t rowB = argc > 1 ? atoi(argv[1]) : 256;
  const uint64_t nIdx = 64ull << 20;                  // 64 M row fetches per launch
  const uint32_t Rbig = (uint32_t)((1ull << 30) / rowB) * (argc > 2 ? atoi(argv[2]) : 1);   // 1 GB table by default
  const uint32_t Rhot = (uint32_t)((32ull << 20) / rowB);                                     // 32 MB hot set
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
  printf("# %s, %d SMs, rowB=%u, table=%.2f GB (%u rows), %llu fetches/launch\n", pr.name, pr.multiProcessorCount, rowB,
         (double)Rbig * rowB / 1e9, Rbig, (unsigned long long)nIdx);
  char* tab; CK(cudaMalloc(&tab, (size_t)Rbig * rowB)); CK(cudaMemset(tab, 0, (size_t)Rbig * rowB));
  uint32_t* idx; CK(cudaMalloc(&idx, nIdx * 4));
  float* sink; CK(cudaMalloc(&sink, 16));
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fnp;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  struct Dist { const char* name; int mode; uint32_t R; } dists[] = {{"L2-resident(32MB)", 0, Rhot}, {"uniform(table)", 0, Rbig}, {"50%hot/50%cold", 50, Rbig}, {"69%hot/31%cold", 69, Rbig}};
  for (auto& d : dists) {
    k_fill_idx<<<1184, 256>>>(idx, nIdx, d.R, Rhot, d.mode, 12345u);
    CK(cudaDeviceSynchronize());
    CUtensorMap tm;
    cuuint64_t dims[2] = {rowB / 4, Rbig}; cuuint64_t strides[1] = {rowB}; cuuint32_t box[2] = {rowB / 4, 1}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tab, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     rowB >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    auto report = [&](const char* what, float ms) {
      printf("%-20s %-34s %8.3f ms  %7.1f GB/s  %6.2f B/clk/SM@1.92GHz\n", d.name, what, ms, (double)nIdx * rowB / ms / 1e6,
             (double)nIdx * rowB / (ms * 1e-3) / pr.multiProcessorCount / 1.92e9);
      fflush(stdout);
    };
    auto timeit = [&](auto&& launch) {
      launch(); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); launch(); launch(); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaGetLastError()); return ms / 3;
    };
    const uint32_t rowQ = rowB / 16;
    char nm[96];
#define LDG_CASE(LANES, U, BPS)                                                                                   \
    if (rowQ == LANES) {                                                                                          \
      snprintf(nm, sizeof nm, "ldg U=%d blocks/SM=%d", U, BPS);                                                  \
      report(nm, timeit([&] { k_ldg<LANES, U><<<pr.multiProcessorCount * BPS, 256>>>((const float4*)tab, idx, nIdx, rowQ, sink); })); \
    }
    LDG_CASE(16, 8, 4) LDG_CASE(16, 8, 6) LDG_CASE(16, 8, 8) LDG_CASE(16, 16, 4) LDG_CASE(16, 16, 8)
    LDG_CASE(32, 4, 8) LDG_CASE(32, 8, 8) LDG_CASE(32, 16, 4)
    LDG_CASE(4, 8, 8) LDG_CASE(4, 16, 8)
    struct TC { int D, G4, bps; } tcs[] = {{4, 2, 4}, {2, 2, 8}};
    for (auto& t : tcs) {
      const size_t smem = (size_t)4 * t.D * t.G4 * 4 * rowB + 4 * t.D * 8;
      if (smem * t.bps > 220 * 1024 || smem > 200 * 1024) continue;
      CK(cudaFuncSetAttribute(k_tma<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaFuncSetAttribute(k_tma<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaFuncSetAttribute(k_tma<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int grid = pr.multiProcessorCount * t.bps;
      snprintf(nm, sizeof nm, "tma4 D=%d G4=%d CTAs/SM=%d (%zuKB/SM)", t.D, t.G4, t.bps, smem * t.bps / 1024);
      report(nm, timeit([&] { k_tma<0, 0><<<grid, 128, smem>>>(tm, tab, idx, nIdx, rowB, t.D, t.G4, sink); }));
      snprintf(nm, sizeof nm, "tma4+read D=%d G4=%d CTAs/SM=%d", t.D, t.G4, t.bps);
      report(nm, timeit([&] { k_tma<0, 1><<<grid, 128, smem>>>(tm, tab, idx, nIdx, rowB, t.D, t.G4, sink); }));
      snprintf(nm, sizeof nm, "bulk D=%d G4=%d CTAs/SM=%d", t.D, t.G4, t.bps);
      report(nm, timeit([&] { k_tma<1, 0><<<grid, 128, smem>>>(tm, tab, idx, nIdx, rowB, t.D, t.G4, sink); }));
    }

    struct RC { int R, W, bps; } rcs[] = {{6, 3, 2}, {4, 2, 3}, {2, 1, 6}, {2, 2, 6}};   // R % W == 0: a worker always reuses its own slots
    for (auto& t : rcs) {
      const size_t smem = (size_t)t.R * 64 * rowB + (size_t)t.R * 16;
      if (smem * t.bps > 222 * 1024 || smem > 220 * 1024) continue;
      CK(cudaFuncSetAttribute(k_ring<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaFuncSetAttribute(k_ring<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int grid = pr.multiProcessorCount * t.bps;
      snprintf(nm, sizeof nm, "ring R=%d W=%d CTAs/SM=%d (%zuKB/SM)", t.R, t.W, t.bps, smem * t.bps / 1024);
      report(nm, timeit([&] { k_ring<0><<<grid, 32 * (t.W + 1), smem>>>(tm, idx, nIdx / 64, rowB, t.R, t.W, sink); }));
      snprintf(nm, sizeof nm, "ring+read R=%d W=%d CTAs/SM=%d", t.R, t.W, t.bps);
      report(nm, timeit([&] { k_ring<1><<<grid, 32 * (t.W + 1), smem>>>(tm, idx, nIdx / 64, rowB, t.R, t.W, sink); }));
    }
  }
  return 0;
}
