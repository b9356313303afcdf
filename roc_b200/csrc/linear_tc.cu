// linear_tc.cu — tcgen05 / TMEM tensor-core GEMMs for the Linear op, fed by TMA.
//
// Replaces cublasSgemm of Linear::forward_task (linear_kernel.cu:76-80):
//     Y[v][o] = sum_i X[v][i] * W[o*in + i]        M = rows, N = outDim, K = inDim
// fp32 in, fp32 out, within 1e-4 of sgemm: the tensor cores have no fp32 MMA, so
// each operand is split hi + lo (hi = fp32 truncated to TF32's 10-bit mantissa,
// lo = exact remainder) and three kind::tf32 MMAs accumulate hi*hi + lo*hi + hi*lo
// in fp32 TMEM (3xTF32: error ~2^-20 relative per product).
//
// One persistent CTA per SM, 8 warps:
//   warp 0   TMA producer: X tile [128 rows][32 k] + W_hi / W_lo tiles, 128B swizzle,
//            4-stage mbarrier ring
//   warp 1   MMA issuer (one elected thread): 12 x tcgen05.mma (M128 x N x K8) per stage,
//            tcgen05.commit frees the stage / signals the epilogue
//   warp 2   TMEM allocator
//   warps 4-7 operand split (X tile -> hi in place, lo beside it; fence.proxy.async)
//            then, per tile, the epilogue: tcgen05.ld -> relu / row-norm -> global
// The kernel is HBM-bound on X (rows*inDim*4 bytes read once); W (tens of KB) stays
// in L2.  Roofline: DESIGN.md.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"

namespace roc {

using namespace tc;

constexpr int TC_BM = 128;       // rows per tile (UMMA M)
constexpr int TC_BK = 32;        // fp32 per k-block = one 128-byte swizzle row
constexpr int TC_UK = 8;         // UMMA K for tf32
constexpr int TC_THREADS = 256;
constexpr int TC_MAX_STAGES = 4;

struct TcFwdParams {
  float* Y; int64_t ldY;
  int64_t rows; int outDim; int BN; int numKb; int stages; uint32_t tmemCols;
  int relu;
  const uint64_t* rowEnd; uint64_t colLeft;
};

__global__ void k_split_w(int outDim, int inDim, int BN, int Kpad, const float* __restrict__ W,
                          float* __restrict__ Whi, float* __restrict__ Wlo) {
  int total = BN * Kpad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int o = i / Kpad, k = i - o * Kpad;
    float w = (o < outDim && k < inDim) ? W[(size_t)o * inDim + k] : 0.f;
    float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    Whi[i] = hi;
    Wlo[i] = w - hi;
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_linear_fwd(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWhi,
                const __grid_constant__ CUtensorMap mapWlo, const TcFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t aBytes = TC_BM * TC_BK * 4;                 // 16 KB
  const uint32_t bBytes = (uint32_t)p.BN * TC_BK * 4;
  const uint32_t stageBytes = 2 * aBytes + 2 * bBytes;
  uint8_t* barBase = smem + (size_t)p.stages * stageBytes;
  uint64_t* fullTma = reinterpret_cast<uint64_t*>(barBase);            // [stages] TMA landed
  uint64_t* fullSplit = fullTma + TC_MAX_STAGES;                       // [stages] hi/lo written
  uint64_t* empty = fullSplit + TC_MAX_STAGES;                         // [stages] MMAs done with the stage
  uint64_t* tmemFull = empty + TC_MAX_STAGES;                          // accumulator complete
  uint32_t* tmemAddr = reinterpret_cast<uint32_t*>(tmemFull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t numTiles = (p.rows + TC_BM - 1) / TC_BM;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapWhi); tma_prefetch_desc(&mapWlo);
    for (int s = 0; s < p.stages; s++) { mbar_init(&fullTma[s], 1); mbar_init(&fullSplit[s], 4); mbar_init(&empty[s], 1); }
    mbar_init(tmemFull, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmemAddr, p.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = *tmemAddr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.numKb; kb++) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = smem + (size_t)s * stageBytes;
          mbar_arrive_expect_tx(&fullTma[s], aBytes + 2 * bBytes);
          tma_load_2d(st, &mapX, kb * TC_BK, (int)(tile * TC_BM), &fullTma[s]);
          tma_load_2d(st + 2 * aBytes, &mapWhi, kb * TC_BK, 0, &fullTma[s]);
          tma_load_2d(st + 2 * aBytes + bBytes, &mapWlo, kb * TC_BK, 0, &fullTma[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(TC_BM, p.BN, 0, 0);
      int s = 0; uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.numKb; kb++) {
          mbar_wait(&fullSplit[s], ph);
          tc_fence_after();
          const uint32_t aHi = smem_u32(smem + (size_t)s * stageBytes);
          const uint32_t aLo = aHi + aBytes;
          const uint32_t bHi = aHi + 2 * aBytes;
          const uint32_t bLo = bHi + bBytes;
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UK; k++) {
            const uint32_t off = k * TC_UK * 4;   // bytes along K inside the swizzled row
            const uint64_t dAh = make_sdesc_sw128(aHi + off, 16, 1024), dAl = make_sdesc_sw128(aLo + off, 16, 1024);
            const uint64_t dBh = make_sdesc_sw128(bHi + off, 16, 1024), dBl = make_sdesc_sw128(bLo + off, 16, 1024);
            umma_tf32(tmemBase, dAl, dBh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmemBase, dAh, dBl, idesc, 1u);
            umma_tf32(tmemBase, dAh, dBh, idesc, 1u);
          }
          umma_commit(&empty[s]);                          // stage reusable once these MMAs retire
          if (kb == p.numKb - 1) umma_commit(tmemFull);   // accumulator of this tile complete
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== operand split, then the tile epilogue ==================
    const int t = threadIdx.x - 128;    // 0..127
    int s = 0; uint32_t ph = 0; uint32_t tilePh = 0;
    for (int64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.numKb; kb++) {
        mbar_wait(&fullTma[s], ph);
        float4* a = reinterpret_cast<float4*>(smem + (size_t)s * stageBytes);
        float4* al = reinterpret_cast<float4*>(smem + (size_t)s * stageBytes + aBytes);
#pragma unroll
        for (int j = 0; j < (TC_BM * TC_BK / 4) / 128; j++) {
          const int i = j * 128 + t;
          float4 v = a[i], h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
          a[i] = h; al[i] = l;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&fullSplit[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      // ---- epilogue: TMEM lane = row of the tile; this warp owns lanes 32*(warp%4)..+31
      mbar_wait(tmemFull, tilePh);
      tilePh ^= 1;
      tc_fence_after();
      const int64_t row = tile * TC_BM + (warp - 4) * 32 + lane;
      float d = 1.0f;
      if (p.rowEnd && row < p.rows) {
        uint64_t st = (row == 0) ? p.colLeft : p.rowEnd[row - 1];
        d = sqrtf((float)(uint32_t)(p.rowEnd[row] - st));
      }
      const RowDiv rd = rowdiv_make(d);
      const uint32_t taddr = tmemBase + ((uint32_t)((warp - 4) * 32) << 16);
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (row < p.rows) {
          float* y = p.Y + row * p.ldY + c0;
#pragma unroll
          for (int q = 0; q < 4; q++) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
              float x = __uint_as_float(r[q * 4 + k]);
              if (p.relu) x = relu_nanprop(x);
              if (p.rowEnd) x = rowdiv(x, rd);   // == x / d bit for bit (common.cuh)
              v[k] = x;
            }
            const int c = c0 + q * 4;
            if (c + 4 <= p.outDim) *reinterpret_cast<float4*>(y + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
            else
#pragma unroll
              for (int k = 0; k < 4; k++) if (c + k < p.outDim) y[q * 4 + k] = v[k];
          }
        }
      }
      tc_fence_before();   // TMEM reads done before the next tile's first MMA may overwrite it
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmemBase, p.tmemCols);
}

// ---------------------------------------------------------------------------
// TS variant: the X operand reaches the tensor core through TMEM instead of smem.
// The SS kernel above moves every X element through shared memory five times (TMA
// write, split read, hi+lo write, 3 MMA operand reads ~ 136 KB per k-block, ~1100
// smem cycles against ~700 cycles of HBM time for the same k-block); here the
// split warps read the TMA tile once and tcgen05.st hi / lo straight into TMEM
// (A operand, K-major: lane = row, column = k), so smem carries 16 KB in + 16 KB
// out of X and the W tiles only.  12 warps:
//   warp 0    TMA producer (X tile + W_hi + W_lo per stage)
//   warp 1    MMA issuer: 12 x tcgen05.mma [D], [A_tmem], B_smem per stage
//   warp 2    TMEM allocator (512 columns: 2 accumulators + `stages` A slots of 64)
//   warps 4-7 split: smem (128B-swizzled) -> regs -> hi/lo -> tcgen05.st
//   warps 8-11 epilogue of tile i overlaps the main loop of tile i+1 (two accumulators)
constexpr int TS_THREADS = 512;      // launched with 384 when splitGroups == 1
constexpr int TS_MAX_STAGES = 6;

struct TcTsParams {
  float* Y; int64_t ldY;
  int64_t rows; int outDim; int BN; int numKb; int stages;
  uint32_t tmemCols, dStride, aCol0;
  int splitGroups;      // 1 or 2 warpgroups of split warps (alternate k-blocks)
  const uint32_t* mask; int64_t ldm; float mscale;   // fused dropout of X (NULL = none)
  int relu;
  const uint64_t* rowEnd; uint64_t colLeft;
};

__device__ __forceinline__ void tc_epilogue_rows(uint32_t taddr, int BN, int outDim, int64_t row, int64_t rows,
                                                 float* Y, int64_t ldY, int relu, const uint64_t* rowEnd,
                                                 uint64_t colLeft) {
  float d = 1.0f;
  if (rowEnd && row < rows) {
    uint64_t st = (row == 0) ? colLeft : rowEnd[row - 1];
    d = sqrtf((float)(uint32_t)(rowEnd[row] - st));
  }
  const RowDiv rd = rowdiv_make(d);
  for (int c0 = 0; c0 < BN; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(taddr + (uint32_t)c0, r);
    tmem_ld_wait();
    if (row < rows) {
      float* y = Y + row * ldY + c0;
#pragma unroll
      for (int q = 0; q < 4; q++) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          float x = __uint_as_float(r[q * 4 + k]);
          if (relu) x = relu_nanprop(x);
          if (rowEnd) x = rowdiv(x, rd);   // == x / d bit for bit (common.cuh)
          v[k] = x;
        }
        const int c = c0 + q * 4;
        if (c + 4 <= outDim) *reinterpret_cast<float4*>(y + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
        else
#pragma unroll
          for (int k = 0; k < 4; k++) if (c + k < outDim) y[q * 4 + k] = v[k];
      }
    }
  }
}

__global__ void __launch_bounds__(TS_THREADS, 1)
k_tc_linear_fwd_ts(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWhi,
                   const __grid_constant__ CUtensorMap mapWlo, const TcTsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t aBytes = TC_BM * TC_BK * 4;                 // 16 KB of X per stage
  const uint32_t bBytes = (uint32_t)p.BN * TC_BK * 4;
  const uint32_t stageBytes = aBytes + 2 * bBytes;
  uint8_t* barBase = smem + (size_t)p.stages * stageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(barBase);    // [stages] TMA landed (X, W_hi, W_lo)
  uint64_t* aFull = full + TS_MAX_STAGES;                   // [stages] hi/lo of X are in TMEM slot s
  uint64_t* empty = aFull + TS_MAX_STAGES;                  // [stages] MMAs done with smem stage + TMEM slot
  uint64_t* dFull = empty + TS_MAX_STAGES;                  // [2] accumulator complete
  uint64_t* dEmpty = dFull + 2;                             // [2] accumulator drained by the epilogue
  uint32_t* tmemAddr = reinterpret_cast<uint32_t*>(dEmpty + 2);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int64_t numTiles = (p.rows + TC_BM - 1) / TC_BM;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapWhi); tma_prefetch_desc(&mapWlo);
    for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&aFull[s], 4); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&dFull[b], 1); mbar_init(&dEmpty[b], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmemAddr, p.tmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = *tmemAddr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    int s = 0; uint32_t ph = 0;
    for (int64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.numKb; kb++) {
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * stageBytes;
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], aBytes + 2 * bBytes);
          tma_load_2d(st, &mapX, kb * TC_BK, (int)(tile * TC_BM), &full[s]);
          tma_load_2d(st + aBytes, &mapWhi, kb * TC_BK, 0, &full[s]);
          tma_load_2d(st + aBytes + bBytes, &mapWlo, kb * TC_BK, 0, &full[s]);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================= MMA issuer =================================
    // the whole warp runs the loop converged; one elected lane issues (see elect_one)
    const uint32_t idesc = make_idesc_tf32(TC_BM, p.BN, 0, 0);
    int s = 0; uint32_t ph = 0; uint32_t tl = 0;
    for (int64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x, tl++) {
      const uint32_t b = tl & 1u;
      mbar_wait(&dEmpty[b], ((tl >> 1) & 1u) ^ 1u);   // the epilogue has drained accumulator b
      tc_fence_after();
      const uint32_t dAddr = tmemBase + b * p.dStride;
      for (int kb = 0; kb < p.numKb; kb++) {
        mbar_wait(&full[s], ph);                      // W tiles landed
        mbar_wait(&aFull[s], ph);                     // X hi/lo stored to TMEM slot s
        tc_fence_after();
        const uint32_t aHi = tmemBase + p.aCol0 + (uint32_t)s * 64u;
        const uint32_t aLo = aHi + 32u;
        const uint32_t bHi = smem_u32(smem + (size_t)s * stageBytes) + aBytes;
        const uint32_t bLo = bHi + bBytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < TC_BK / TC_UK; k++) {
            const uint32_t off = k * TC_UK * 4;   // bytes along K inside the swizzled row
            const uint64_t dBh = make_sdesc_sw128(bHi + off, 16, 1024), dBl = make_sdesc_sw128(bLo + off, 16, 1024);
            umma_tf32_ts(dAddr, aLo + k * TC_UK, dBh, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32_ts(dAddr, aHi + k * TC_UK, dBl, idesc, 1u);
            umma_tf32_ts(dAddr, aHi + k * TC_UK, dBh, idesc, 1u);
          }
          umma_commit(&empty[s]);                           // smem stage + TMEM slot reusable
          if (kb == p.numKb - 1) umma_commit(&dFull[b]);   // accumulator of this tile complete
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 4 + 4 * p.splitGroups) {
    // ============== operand split: smem X tile -> hi / lo in TMEM slot s ==============
    // group g of the split warps takes every splitGroups-th k-block, so one group's
    // tcgen05.st / wait::st latency overlaps the other group's loads
    const int g = (warp - 4) >> 2;
    const int t = (threadIdx.x - 128) & 127;    // row of the tile == TMEM lane
    const uint32_t laneBase = (uint32_t)((warp & 3) * 32) << 16;
    const int64_t myTiles = (numTiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int64_t iters = myTiles * p.numKb;
    const bool masked = p.mask != nullptr;
    int s = g % p.stages; uint32_t ph = (uint32_t)(g / p.stages) & 1u;
    int kb = g % p.numKb; int64_t tile = blockIdx.x + (int64_t)(g / p.numKb) * gridDim.x;
    for (int64_t it = g; it < iters; it += p.splitGroups) {
      // dropout fused into the operand load: bit c of mask[row][kb] keeps column 32 kb + c
      uint32_t mw = 0xFFFFFFFFu;
      if (masked) {
        const int64_t row = tile * TC_BM + t;
        mw = (row < p.rows) ? __ldg(p.mask + row * p.ldm + kb) : 0u;
      }
      mbar_wait(&full[s], ph);
      // 128B swizzle: 16-byte chunk j of row t sits at chunk j ^ (t & 7)
      const float4* xr = reinterpret_cast<const float4*>(smem + (size_t)s * stageBytes) + t * 8;
      const uint32_t taddr = tmemBase + laneBase + p.aCol0 + (uint32_t)s * 64u;
#pragma unroll
      for (int half = 0; half < 2; half++) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const float4 v = xr[(half * 4 + j) ^ (t & 7)];
          const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int k = 0; k < 4; k++) {
            float x = e[k];
            if (masked) x = ((mw >> (half * 16 + j * 4 + k)) & 1u) ? x * p.mscale : 0.f;   // == k_dropout
            const uint32_t h = __float_as_uint(x) & 0xFFFFE000u;
            hi[j * 4 + k] = h;
            lo[j * 4 + k] = __float_as_uint(x - __uint_as_float(h));
          }
        }
        tmem_st16(taddr + half * 16, hi);
        tmem_st16(taddr + 32 + half * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&aFull[s]);
      s += p.splitGroups; if (s >= p.stages) { s -= p.stages; ph ^= 1u; }
      kb += p.splitGroups; while (kb >= p.numKb) { kb -= p.numKb; tile += gridDim.x; }
    }
  } else if (warp >= 4 + 4 * p.splitGroups) {
    // ===== epilogue: TMEM lane = row of the tile; this warp owns lanes 32*(warp%4)..+31 =====
    uint32_t tl = 0;
    for (int64_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x, tl++) {
      const uint32_t b = tl & 1u;
      mbar_wait(&dFull[b], (tl >> 1) & 1u);
      tc_fence_after();
      const int64_t row = tile * TC_BM + (warp & 3) * 32 + lane;
      const uint32_t taddr = tmemBase + ((uint32_t)((warp & 3) * 32) << 16) + b * p.dStride;
      tc_epilogue_rows(taddr, p.BN, p.outDim, row, p.rows, p.Y, p.ldY, p.relu, p.rowEnd, p.colLeft);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dEmpty[b]);
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmemBase, p.tmemCols);
}

// scratch for the split weights (grow-only, per process; the host uses one stream)
static float* g_wsplit = nullptr;
static size_t g_wsplitFloats = 0;

int tc_linear_fwd(int64_t rows, int inDim, int outDim, const float* X, int64_t ldX, const float* W, float* Y,
                  int64_t ldY, int relu, const uint64_t* rowEnd, uint64_t colLeft, const DropMask* dm,
                  cudaStream_t st) {
  if (outDim > 256 || inDim < 8 || rows < 1) return ROC_ERR_UNSUPPORTED;
  if ((ldX % 4) || (ldY % 4) || !aligned16(X) || !aligned16(Y)) return ROC_ERR_UNSUPPORTED;
  if (rows > 0x7FFFFF00ll) return ROC_ERR_UNSUPPORTED;   // TMA coordinates are int32
  if (!encode_tiled_fn()) return ROC_ERR_UNSUPPORTED;
  const int BN = (outDim + 15) / 16 * 16;
  const int Kpad = (inDim + TC_BK - 1) / TC_BK * TC_BK;
  const size_t need = (size_t)2 * BN * Kpad;
  if (need > g_wsplitFloats) {
    if (g_wsplit) { ROC_CUDA(cudaDeviceSynchronize()); ROC_CUDA(cudaFree(g_wsplit)); g_wsplit = nullptr; g_wsplitFloats = 0; }
    ROC_CUDA(cudaMalloc(&g_wsplit, need * sizeof(float)));
    g_wsplitFloats = need;
  }
  float* Whi = g_wsplit;
  float* Wlo = g_wsplit + (size_t)BN * Kpad;
  k_split_w<<<(BN * Kpad + 255) / 256, 256, 0, st>>>(outDim, inDim, BN, Kpad, W, Whi, Wlo);
  ROC_LAUNCH_CHECK();

  CUtensorMap mapX, mapWhi, mapWlo;
  if (!make_tmap_f32_2d(&mapX, X, (uint64_t)rows, (uint64_t)inDim, (uint64_t)ldX, TC_BM, TC_BK)) return ROC_ERR_UNSUPPORTED;
  if (!make_tmap_f32_2d(&mapWhi, Whi, (uint64_t)BN, (uint64_t)Kpad, (uint64_t)Kpad, (uint32_t)BN, TC_BK)) return ROC_ERR_UNSUPPORTED;
  if (!make_tmap_f32_2d(&mapWlo, Wlo, (uint64_t)BN, (uint64_t)Kpad, (uint64_t)Kpad, (uint32_t)BN, TC_BK)) return ROC_ERR_UNSUPPORTED;

  const int64_t numTiles = (rows + TC_BM - 1) / TC_BM;
  int grid = sm_count();
  if (numTiles < grid) grid = (int)numTiles;
  {
    // TS path: 2 accumulators of dStride columns + `stages` A slots of 64 columns in 512 TMEM columns
    const char* g = getenv("ROC_B200_GEMM");
    const uint32_t dStride = (uint32_t)((BN + 31) / 32 * 32);
    const size_t stageBytesTs = (size_t)TC_BM * TC_BK * 4 + (size_t)2 * BN * TC_BK * 4;
    int stagesTs = (int)((512 - 2 * dStride) / 64);
    if (stagesTs > TS_MAX_STAGES) stagesTs = TS_MAX_STAGES;
    while (stagesTs > 0 && (size_t)stagesTs * stageBytesTs + 1024 + 256 > (size_t)220 * 1024) stagesTs--;
    if (stagesTs >= 3 && !(g && g[0] == 's' && g[1] == 's')) {
      TcTsParams q{};
      q.Y = Y; q.ldY = ldY; q.rows = rows; q.outDim = outDim; q.BN = BN; q.numKb = Kpad / TC_BK;
      q.stages = stagesTs; q.tmemCols = 512; q.dStride = dStride; q.aCol0 = 2 * dStride;
      q.relu = relu; q.rowEnd = rowEnd; q.colLeft = colLeft;
      { const char* e = getenv("ROC_TS_SPLIT"); q.splitGroups = (e && e[0] == '1') ? 1 : 2; }
      if (dm) { q.mask = dm->bits; q.ldm = dm->ld; q.mscale = dm->scale; }
      const size_t smemTs = (size_t)stagesTs * stageBytesTs + 1024 + 256;
      static size_t configuredTs = 0;
      if (smemTs > configuredTs) {
        ROC_CUDA(cudaFuncSetAttribute(k_tc_linear_fwd_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemTs));
        configuredTs = smemTs;
      }
      k_tc_linear_fwd_ts<<<grid, 256 + 128 * q.splitGroups, smemTs, st>>>(mapX, mapWhi, mapWlo, q);
      ROC_LAUNCH_CHECK();
      return ROC_OK;
    }
  }

  if (dm) return ROC_ERR_UNSUPPORTED;   // only the TS kernel fuses the dropout mask
  TcFwdParams p{};
  p.Y = Y; p.ldY = ldY; p.rows = rows; p.outDim = outDim; p.BN = BN; p.numKb = Kpad / TC_BK;
  p.relu = relu; p.rowEnd = rowEnd; p.colLeft = colLeft;
  uint32_t cols = 32;
  while ((int)cols < BN) cols <<= 1;
  p.tmemCols = cols;
  const size_t stageBytes = (size_t)2 * TC_BM * TC_BK * 4 + (size_t)2 * BN * TC_BK * 4;
  int stages = (int)((200 * 1024) / stageBytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) return ROC_ERR_UNSUPPORTED;
  p.stages = stages;
  const size_t smemBytes = (size_t)stages * stageBytes + 1024 /*align*/ + 256 /*barriers*/;
  static size_t configured = 0;
  if (smemBytes > configured) {
    ROC_CUDA(cudaFuncSetAttribute(k_tc_linear_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
    configured = smemBytes;
  }
  k_tc_linear_fwd<<<grid, TC_THREADS, smemBytes, st>>>(mapX, mapWhi, mapWlo, p);
  ROC_LAUNCH_CHECK();
  return ROC_OK;
}

}  // namespace roc
