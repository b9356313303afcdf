// linear_tc.cu — tcgen05 / TMEM tensor-core GEMMs for the Linear op, fed by TMA.
//
// Replaces cublasSgemm of Linear::forward_task (linear_kernel.cu:76-80) and the dX sgemm of
// Linear::backward_task (linear_kernel.cu:227-231):
//     fwd: Y[v][o]  = sum_i X[v][i]  * W[o*in + i]      M = rows, N = outDim, K = inDim
//     dX : dX[v][i] = sum_o dY[v][o] * W[o*in + i]      M = rows, N = inDim,  K = outDim
// fp32 in, fp32 out, within 1e-4 of sgemm: the tensor cores have no fp32 MMA, so
// each operand is split hi + lo (hi = fp32 truncated to TF32's 10-bit mantissa,
// lo = exact remainder) and three kind::tf32 MMAs accumulate hi*hi + lo*hi + hi*lo
// in fp32 TMEM (3xTF32: error ~2^-20 relative per product).
//
// One kernel serves both: a row-major [rows][K] streaming operand A (X or dY) times a
// small K-major weight operand prepared by k_split_w ([N][K] hi / lo; dX passes W
// transposed).  The streaming operand reaches the tensor core through TMEM: moving it
// through shared memory as hi + lo costs ~136 KB of smem traffic per k-block (~1100 smem
// cycles against ~700 cycles of HBM time); here the split warps read the TMA tile once
// and tcgen05.st hi / lo straight into TMEM (A operand, K-major: lane = row, column = k).
// One persistent CTA per SM (x N-tiles of <= 128 columns in grid.y), 12 or 16 warps:
//   warp 0     TMA producer (A tile [128 rows][32 k] + W_hi + W_lo tile per stage), elect.sync
//   warp 1     MMA issuer: 12 x tcgen05.mma [D], [A_tmem], B_smem per stage, elect.sync
//   warp 2     TMEM allocator (512 columns: 2 accumulators + `stages` A slots of 64)
//   warps 4-11 split (1 or 2 groups of 4 alternate k-blocks): smem (128B-swizzled) -> regs
//              [-> dropout mask] -> hi/lo -> tcgen05.st
//   last 4     epilogue of tile i (relu / dropout-backward mask / relu mask / row norm /
//              accumulate) overlaps the main loop of tile i+1 (two accumulators)
// HBM-bound on A (rows*K*4 bytes read once); W (tens of KB) stays in L2.  Roofline: DESIGN.md.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"

namespace roc {

using namespace tc;

constexpr int TC_BM = 128;       // rows per tile (UMMA M)
constexpr int TC_BK = 32;        // fp32 per k-block = one 128-byte swizzle row
constexpr int TC_UK = 8;         // UMMA K for tf32
constexpr int TS_THREADS = 512;  // 4 control warps + 4 x splitGroups split warps + epiWarps epilogue warps
constexpr int TS_MAX_STAGES = 6;

// Whi/Wlo[n][k] (n < Npad, k < Kpad, zero padded) = split of B(n, k), where
//   transposed == 0: B(n, k) = W[n * ldW + k]   (fwd: n = output o, k = input i, ldW = inDim)
//   transposed == 1: B(n, k) = W[k * ldW + n]   (dX : n = input i,  k = output o)
__global__ void k_split_w(int N, int K, int Npad, int Kpad, int ldW, int transposed, const float* __restrict__ W,
                          float* __restrict__ Whi, float* __restrict__ Wlo) {
  int total = Npad * Kpad;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int n = i / Kpad, k = i - n * Kpad;
    float w = 0.f;
    if (n < N && k < K) w = transposed ? W[(size_t)k * ldW + n] : W[(size_t)n * ldW + k];
    float hi = __uint_as_float(__float_as_uint(w) & 0xFFFFE000u);
    Whi[i] = hi;
    Wlo[i] = w - hi;
  }
}

struct TcTsParams {
  float* Y; int64_t ldY;
  int64_t rows; int outDim; int BN; int numKb; int stages;
  uint32_t tmemCols, dStride, aCol0;
  int splitGroups;      // 1 or 2 warpgroups of split warps (alternate k-blocks)
  int epiWarps;         // 4 or 8 epilogue warps (8: two warps per TMEM lane quarter alternate 32-column chunks)
  const uint32_t* mask; int64_t ldm; float mscale;      // dropout of A fused into the operand load (NULL = none)
  // epilogue, applied in this order to acc(row, col):
  int relu;                                             //   relu (Linear's activation, linear_kernel.cu:83-104)
  const uint32_t* omask; int64_t oldm; float oscale;    //   dropout backward: keep ? x * scale : 0
  const float* reluOf; int64_t ldR;                     //   relu backward: reluOf(row, col) > 0 ? x : 0
  const uint64_t* rowEnd; uint64_t colLeft;             //   indegree norm: x / sqrtf(deg(row))
  int accumulate;                                       //   Y += x instead of Y = x
  int epiStaged;        // transpose the accumulator through smem for row-coalesced global accesses
};

// Epilogue stage switches: compile-time when the kernel is instantiated for a fixed combination
// (EPI >= 0, a bit mask), read from the parameters for the generic instantiation (EPI < 0).  With all
// five stages dynamic the unrolled epilogue was ~2000 SASS instructions and the epilogue warps stalled
// on instruction fetch (r1 run 29).
enum { EPI_RELU = 1, EPI_OMASK = 2, EPI_RELUOF = 4, EPI_NORM = 8, EPI_ACC = 16 };
#define EPI_FLAGS(EPI, p)                                                        \
  const bool fRelu = (EPI) < 0 ? (p).relu != 0 : ((EPI) & EPI_RELU) != 0;         \
  const bool fOmask = (EPI) < 0 ? (p).omask != nullptr : ((EPI) & EPI_OMASK) != 0; \
  const bool fReluOf = (EPI) < 0 ? (p).reluOf != nullptr : ((EPI) & EPI_RELUOF) != 0; \
  const bool fNorm = (EPI) < 0 ? (p).rowEnd != nullptr : ((EPI) & EPI_NORM) != 0;  \
  const bool fAcc = (EPI) < 0 ? (p).accumulate != 0 : ((EPI) & EPI_ACC) != 0;

// Direct epilogue: each lane writes its own row (32 rows x 16 B per store instruction).
template <int EPI>
__device__ __forceinline__ void tc_epilogue_rows(uint32_t taddr, int n0, int64_t row, const TcTsParams& p, float d) {
  EPI_FLAGS(EPI, p)
  const RowDiv rd = rowdiv_make(d);
  for (int c0 = 0; c0 < p.BN; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(taddr + (uint32_t)c0, r);
    tmem_ld_wait();
    const int col0 = n0 + c0;
    if (row < p.rows && col0 < p.outDim) {
      float* y = p.Y + row * p.ldY + col0;
      uint32_t mw = 0xFFFFFFFFu;      // the 16 columns of this chunk lie in one mask word (col0 % 16 == 0)
      if (fOmask) mw = __ldg(p.omask + row * p.oldm + (col0 >> 5)) >> (col0 & 31);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int c = col0 + q * 4;
        const bool full = c + 4 <= p.outDim;
        float v[4], old[4] = {0.f, 0.f, 0.f, 0.f}, ro[4] = {1.f, 1.f, 1.f, 1.f};
        if (fAcc) {
          if (full) *reinterpret_cast<float4*>(old) = *reinterpret_cast<const float4*>(y + q * 4);
          else
#pragma unroll
            for (int k = 0; k < 4; k++) if (c + k < p.outDim) old[k] = y[q * 4 + k];
        }
        if (fReluOf) {
          const float* rp = p.reluOf + row * p.ldR + c;
          if (full) *reinterpret_cast<float4*>(ro) = __ldg(reinterpret_cast<const float4*>(rp));
          else
#pragma unroll
            for (int k = 0; k < 4; k++) if (c + k < p.outDim) ro[k] = rp[k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          float x = __uint_as_float(r[q * 4 + k]);
          if (fRelu) x = relu_nanprop(x);
          if (fOmask) x = ((mw >> (q * 4 + k)) & 1u) ? x * p.oscale : 0.f;
          if (fReluOf) x = (ro[k] > 0.f) ? x : 0.f;
          if (fNorm) x = rowdiv(x, rd);   // == x / d bit for bit (common.cuh)
          if (fAcc) x = old[k] + x;
          v[k] = x;
        }
        if (full) *reinterpret_cast<float4*>(y + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
        else
#pragma unroll
          for (int k = 0; k < 4; k++) if (c + k < p.outDim) y[q * 4 + k] = v[k];
      }
    }
  }
}

// Epilogue of one warp's 32 rows of a tile.  tcgen05.ld hands each lane ITS row (lane = TMEM
// lane); written straight to global that is 32 rows x 16 B per instruction.  So each 32-column
// chunk is transposed through a padded shared tile ([32][36] floats per warp) and then handled
// with 8 lanes per row: every global access (Y, the accumulate read, reluOf) is a 128-byte row
// segment, and the per-row scalars (norm divisor, dropout-mask word) are shared by the row's lanes.
constexpr int EPI_LD = 36;                        // padded row of the transposition tile: 16-byte aligned, conflict-free
constexpr int EPI_STG_FLOATS = 32 * EPI_LD + 32; // per warp: transposition tile + per-row divisors

template <int EPI>
__device__ __forceinline__ void tc_epilogue_warp(uint32_t taddr, int n0, int64_t row0, const TcTsParams& p,
                                                 float* stg, int lane, float myD, int chunk0, int chunkStep) {
  EPI_FLAGS(EPI, p)
  float* sd = stg + 32 * EPI_LD;
  sd[lane] = myD;     // this lane's row divisor (computed by the caller before the accumulator wait)
  const int rq = lane >> 3, cq = (lane & 7) * 4;
  for (int c0 = chunk0 * 32; c0 < p.BN; c0 += chunkStep * 32) {
    const int cw = min(32, p.BN - c0);     // 32, or 16 for the last chunk of BN % 32 == 16
    uint32_t r[32];
    tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
    if (cw == 32) tmem_ld16(taddr + (uint32_t)c0 + 16u, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
    tmem_ld_wait();
    __syncwarp();                           // the previous chunk's readers are done with the tile
#pragma unroll
    for (int k = 0; k < 32; k += 4)
      if (k < cw)
        *reinterpret_cast<uint4*>(stg + lane * EPI_LD + k) = make_uint4(r[k], r[k + 1], r[k + 2], r[k + 3]);
    __syncwarp();
    const int c = n0 + c0 + cq;             // global column of this lane's 4 values
    if (cq < cw && c < p.outDim) {
      const bool full = c + 4 <= p.outDim;
      // two batches of 4 rows: every load of a batch is issued before its arithmetic
#pragma unroll
      for (int hb = 0; hb < 2; hb++) {
        float acc[4][4], old[4][4], ro[4][4], dd[4];
        uint32_t mw[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int rl = (hb * 4 + i) * 4 + rq;
          const int64_t row = row0 + rl;
          const bool ok = row < p.rows;
          *reinterpret_cast<float4*>(acc[i]) = *reinterpret_cast<const float4*>(stg + rl * EPI_LD + cq);
          dd[i] = fNorm ? sd[rl] : 1.0f;
          mw[i] = 0xFu;                   // c % 4 == 0: the 4 mask bits sit in one word
          if (fOmask && ok) mw[i] = __ldg(p.omask + row * p.oldm + (c >> 5)) >> (c & 31);
#pragma unroll
          for (int k = 0; k < 4; k++) { old[i][k] = 0.f; ro[i][k] = 1.f; }
          if (fAcc && ok) {
            const float* y = p.Y + row * p.ldY + c;
            if (full) *reinterpret_cast<float4*>(old[i]) = *reinterpret_cast<const float4*>(y);
            else
#pragma unroll
              for (int k = 0; k < 4; k++) if (c + k < p.outDim) old[i][k] = y[k];
          }
          if (fReluOf && ok) {
            const float* rp = p.reluOf + row * p.ldR + c;
            if (full) *reinterpret_cast<float4*>(ro[i]) = __ldg(reinterpret_cast<const float4*>(rp));
            else
#pragma unroll
              for (int k = 0; k < 4; k++) if (c + k < p.outDim) ro[i][k] = rp[k];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int rl = (hb * 4 + i) * 4 + rq;
          const int64_t row = row0 + rl;
          if (row >= p.rows) continue;
          float v[4];
#pragma unroll
          for (int k = 0; k < 4; k++) {
            float x = acc[i][k];
            if (fRelu) x = relu_nanprop(x);
            if (fOmask) x = x * p.oscale;       // the kept value; dropped ones are zeroed after the divide
            v[k] = x;
          }
          // x / d first, masks after: (m ? x : 0) / d == m ? x / d : 0 bit for bit for a finite d > 0, and
          // the divide then never sees the masked-out zeros.  Degenerate divisors (deg 0 -> 0 / 0 = NaN must
          // survive) keep the literal order.
          const RowDiv rd = rowdiv_make(dd[i]);
          const bool maskFirst = fNorm && rd.plain;
#pragma unroll
          for (int pass = 0; pass < 2; pass++) {
            if (pass == (maskFirst ? 0 : 1)) {
#pragma unroll
              for (int k = 0; k < 4; k++) {
                float x = v[k];
                if (fOmask) x = ((mw[i] >> k) & 1u) ? x : 0.f;
                if (fReluOf) x = (ro[i][k] > 0.f) ? x : 0.f;
                v[k] = x;
              }
            } else if (fNorm) {
              rowdiv4(v, rd, full ? 4 : p.outDim - c);
            }
          }
          if (fAcc) {
#pragma unroll
            for (int k = 0; k < 4; k++) v[k] = old[i][k] + v[k];
          }
          float* y = p.Y + row * p.ldY + c;
          if (full) *reinterpret_cast<float4*>(y) = make_float4(v[0], v[1], v[2], v[3]);
          else
#pragma unroll
            for (int k = 0; k < 4; k++) if (c + k < p.outDim) y[k] = v[k];
        }
      }
    }
  }
  __syncwarp();
}

template <int EPI>
__global__ void __launch_bounds__(TS_THREADS, 1)
k_tc_linear_ts(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWhi,
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
  float* epiStg = reinterpret_cast<float*>(barBase + 256);   // 4 warps x EPI_STG_FLOATS

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int64_t numTiles = (p.rows + TC_BM - 1) / TC_BM;
  const int n0 = blockIdx.y * p.BN;          // this CTA's slice of the N (output column) dimension

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapWhi); tma_prefetch_desc(&mapWlo);
    for (int s = 0; s < p.stages; s++) { mbar_init(&full[s], 1); mbar_init(&aFull[s], 4); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&dFull[b], 1); mbar_init(&dEmpty[b], (uint32_t)p.epiWarps); }
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
          tma_load_2d(st + aBytes, &mapWhi, kb * TC_BK, n0, &full[s]);
          tma_load_2d(st + aBytes + bBytes, &mapWlo, kb * TC_BK, n0, &full[s]);
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
    // dropout fused into the operand load: bit c of mask[row][kb] keeps column 32 kb + c.  The word
    // for this group's NEXT k-block is fetched one iteration ahead so its latency hides behind the split.
    uint32_t mwNext = 0xFFFFFFFFu;
    if (masked && g < iters) {
      const int64_t row = tile * TC_BM + t;
      mwNext = (row < p.rows) ? __ldg(p.mask + row * p.ldm + kb) : 0u;
    }
    for (int64_t it = g; it < iters; it += p.splitGroups) {
      const uint32_t mw = mwNext;
      if (masked && it + p.splitGroups < iters) {
        int kbN = kb + p.splitGroups; int64_t tileN = tile;
        while (kbN >= p.numKb) { kbN -= p.numKb; tileN += gridDim.x; }
        const int64_t rowN = tileN * TC_BM + t;
        mwNext = (rowN < p.rows) ? __ldg(p.mask + rowN * p.ldm + kbN) : 0u;
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
      // the row's norm divisor is fetched BEFORE waiting for the accumulator: its DRAM latency
      // (rowEnd streams, always cold) then overlaps the main loop instead of adding to every tile
      const int64_t row0 = tile * TC_BM + (warp & 3) * 32;
      float myD = 1.0f;
      if (p.rowEnd && row0 + lane < p.rows) {
        const int64_t myRow = row0 + lane;
        const uint64_t st = (myRow == 0) ? p.colLeft : p.rowEnd[myRow - 1];
        myD = sqrtf((float)(uint32_t)(p.rowEnd[myRow] - st));
      }
      const int e = warp - 4 - 4 * p.splitGroups;      // epilogue warp index; warps e and e + 4 share a lane quarter
      if (p.epiStaged && p.rowEnd && (lane & 15) == 0) {   // next tile's degrees (32 x u64 per warp = 2 lines)
        const int64_t nrow = (tile + gridDim.x) * TC_BM + (warp & 3) * 32 + lane;
        if (nrow < p.rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.rowEnd + nrow));
      }
      if (p.epiStaged && (p.reluOf || p.accumulate || p.omask)) {
        // pull this warp's first chunk of the NEXT tile's epilogue operands into L2 (and, for the first
        // tile, this one's): they stream (always cold); fetched on demand each batch exposes a DRAM latency.
        // The epilogue paces short-K GEMMs, so the accumulator is usually ready: one tile ahead is needed.
        const int c = n0 + (e >> 2) * 32 + (lane & 7) * 4;
        if (c < p.outDim) {
#pragma unroll 1
          for (int64_t pt = (tl == 0) ? tile : tile + gridDim.x; pt <= tile + gridDim.x && pt < numTiles; pt += gridDim.x) {
            const int64_t prow0 = pt * TC_BM + (warp & 3) * 32;
#pragma unroll
            for (int i = 0; i < 8; i++) {
              const int64_t row = prow0 + i * 4 + (lane >> 3);
              if (row < p.rows) {
                if (p.reluOf) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.reluOf + row * p.ldR + c));
                if (p.accumulate) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.Y + row * p.ldY + c));
                if (p.omask && (lane & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.omask + row * p.oldm + (c >> 5)));
              }
            }
          }
        }
      }
      mbar_wait(&dFull[b], (tl >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmemBase + ((uint32_t)((warp & 3) * 32) << 16) + b * p.dStride;
      if (p.epiStaged) tc_epilogue_warp<EPI>(taddr, n0, row0, p, epiStg + e * EPI_STG_FLOATS, lane, myD, e >> 2, p.epiWarps >> 2);
      else if (e < 4) tc_epilogue_rows<EPI>(taddr, n0, row0 + lane, p, myD);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dEmpty[b]);
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmemBase, p.tmemCols);
}

// The split weights (W_hi / W_lo, padded) live in a stream-ordered allocation made for this call
// (cudaMallocAsync / cudaFreeAsync on the caller's stream): no process-wide scratch, so the C ABI can be driven
// from several threads, devices and streams at once (the reference's task bodies run one per GPU in one process).

struct TsEpilogue {
  int relu = 0;
  const DropMask* outMask = nullptr;
  const float* reluOf = nullptr; int64_t ldR = 0;
  const uint64_t* rowEnd = nullptr; uint64_t colLeft = 0;
  int accumulate = 0;
};

// Y[rows][N] (op)= A[rows][K] * B^T with B(n, k) taken from W as k_split_w describes
static int ts_gemm(int64_t rows, int K, int N, const float* A, int64_t ldA, const float* W, int ldW, int transposed,
                   float* Y, int64_t ldY, const DropMask* inMask, const TsEpilogue& e, cudaStream_t st) {
  if (K < 8 || N < 1 || rows < 1) return ROC_ERR_UNSUPPORTED;
  if ((ldA % 4) || (ldY % 4) || !aligned16(A) || !aligned16(Y)) return ROC_ERR_UNSUPPORTED;
  if (e.reluOf && ((e.ldR % 4) || !aligned16(e.reluOf))) return ROC_ERR_UNSUPPORTED;
  if (rows > 0x7FFFFF00ll) return ROC_ERR_UNSUPPORTED;   // TMA coordinates are int32
  if (!encode_tiled_fn()) return ROC_ERR_UNSUPPORTED;
  { const char* g = getenv("ROC_B200_GEMM"); if (g && g[0] == 'n' && g[1] == 'o') return ROC_ERR_UNSUPPORTED; }   // "notc"
  int BN = (N + 15) / 16 * 16;
  if (BN > 128) BN = 128;
  // (128-column tiles hung sporadically in r2 sessions 1 / 3 / 6 — every time on the 4.2 M-row 602 -> 128 product of
  // configs[3]: the ring had 3 stages and the two split groups met a stage's barrier only every other phase; fixed
  // below by keeping the stage count even.  Validated on configs 2-4 and the GEMM / model tests in sessions 8 / 9;
  // ROC_TS_BN128=0 goes back to 64-column tiles, which stream A once per 64 output columns.)
  { const char* e = getenv("ROC_TS_BN128"); if (BN > 64 && e && e[0] == '0') BN = 64; }
  const int nTiles = (N + BN - 1) / BN;
  const int Npad = nTiles * BN;
  const int Kpad = (K + TC_BK - 1) / TC_BK * TC_BK;
  const size_t need = (size_t)2 * Npad * Kpad;
  float* wsplit = nullptr;
  ROC_CUDA(cudaMallocAsync((void**)&wsplit, need * sizeof(float), st));
  struct FreeOnExit {   // returned to the pool in stream order, after the GEMM below (or at once on an early return)
    float* p; cudaStream_t s;
    ~FreeOnExit() { if (p) cudaFreeAsync(p, s); }
  } wsplitGuard{wsplit, st};
  float* Whi = wsplit;
  float* Wlo = wsplit + (size_t)Npad * Kpad;
  k_split_w<<<(Npad * Kpad + 255) / 256, 256, 0, st>>>(N, K, Npad, Kpad, ldW, transposed, W, Whi, Wlo);
  ROC_LAUNCH_CHECK();

  CUtensorMap mapA, mapWhi, mapWlo;
  if (!make_tmap_f32_2d(&mapA, A, (uint64_t)rows, (uint64_t)K, (uint64_t)ldA, TC_BM, TC_BK)) return ROC_ERR_UNSUPPORTED;
  if (!make_tmap_f32_2d(&mapWhi, Whi, (uint64_t)Npad, (uint64_t)Kpad, (uint64_t)Kpad, (uint32_t)BN, TC_BK)) return ROC_ERR_UNSUPPORTED;
  if (!make_tmap_f32_2d(&mapWlo, Wlo, (uint64_t)Npad, (uint64_t)Kpad, (uint64_t)Kpad, (uint32_t)BN, TC_BK)) return ROC_ERR_UNSUPPORTED;

  // 2 accumulators of dStride columns + `stages` A slots of 64 columns in the 512 TMEM columns
  const uint32_t dStride = (uint32_t)((BN + 31) / 32 * 32);
  const size_t stageBytes = (size_t)TC_BM * TC_BK * 4 + (size_t)2 * BN * TC_BK * 4;
  // Warp roles first (they size the epilogue staging), then the stage count.
  // Few k-blocks per tile (K <= 128): the epilogue, not the main loop, paces the kernel -> 8 epilogue warps and one
  // split group; otherwise two split groups (alternate k-blocks) and 4 epilogue warps.  512 threads either way.
  int epiStaged = 1;
  { const char* ee = getenv("ROC_TS_EPI"); if (ee) epiStaged = (ee[0] == 's'); }
  int splitGroups = (Kpad / TC_BK <= 4 && epiStaged) ? 1 : 2;
  { const char* s2 = getenv("ROC_TS_SPLIT"); if (s2) splitGroups = (s2[0] == '1') ? 1 : 2; }
  if (!epiStaged) splitGroups = 2;
  int stages = 0;
  size_t fixedBytes = 0;
  for (;;) {
    const int epiWarps = 12 - 4 * splitGroups;
    fixedBytes = 1024 /*align*/ + 256 /*barriers*/ + (size_t)epiWarps * EPI_STG_FLOATS * sizeof(float);
    stages = (int)((512 - 2 * dStride) / 64);
    if (stages > TS_MAX_STAGES) stages = TS_MAX_STAGES;
    while (stages > 0 && (size_t)stages * stageBytes + fixedBytes > (size_t)224 * 1024) stages--;
    // Two split groups take alternate k-blocks, i.e. alternate stages: with an EVEN stage count each group owns its
    // stages and meets every phase of their barriers.  With an odd count a group would meet a stage's `full` barrier
    // only every other phase, and an mbarrier parity wait that skipped a phase takes the stale completion for its
    // own (3 stages: sporadic hangs of the 128-column tiles, r2 sessions 1 / 3 / 6; 5 stages — what the 64-column
    // tiles ran with before the staging was sized by the warp roles — left two loads of slack and never tripped).
    if (splitGroups == 2 && (stages & 1)) stages--;
    if (stages >= 3 || splitGroups == 1) break;
    splitGroups = 1;                       // not enough room for an even ring of >= 4: one group, any stage count
  }
  if (stages < 3) return ROC_ERR_UNSUPPORTED;
  TcTsParams q{};
  q.Y = Y; q.ldY = ldY; q.rows = rows; q.outDim = N; q.BN = BN; q.numKb = Kpad / TC_BK;
  q.stages = stages; q.tmemCols = 512; q.dStride = dStride; q.aCol0 = 2 * dStride;
  q.accumulate = e.accumulate;
  q.epiStaged = epiStaged;
  q.splitGroups = splitGroups;
  q.epiWarps = 12 - 4 * q.splitGroups;
  if (inMask) { q.mask = inMask->bits; q.ldm = inMask->ld; q.mscale = inMask->scale; }
  q.relu = e.relu;
  if (e.outMask) { q.omask = e.outMask->bits; q.oldm = e.outMask->ld; q.oscale = e.outMask->scale; }
  q.reluOf = e.reluOf; q.ldR = e.ldR;
  q.rowEnd = e.rowEnd; q.colLeft = e.colLeft;
  const size_t smemBytes = (size_t)stages * stageBytes + fixedBytes;
  const int64_t numTiles = (rows + TC_BM - 1) / TC_BM;
  int gx = sm_count() / nTiles;
  if (gx < 1) gx = 1;
  if (numTiles < gx) gx = (int)numTiles;
  dim3 grid((unsigned)gx, (unsigned)nTiles, 1);
  const int epi = (q.relu ? EPI_RELU : 0) | (q.omask ? EPI_OMASK : 0) | (q.reluOf ? EPI_RELUOF : 0) |
                  (q.rowEnd ? EPI_NORM : 0) | (q.accumulate ? EPI_ACC : 0);
  const unsigned threads = TS_THREADS;
#define ROC_TS_LAUNCH(E)                                                                                          \
  do {                                                                                                            \
    static DynSmemCache configured;                                                                               \
    ROC_CUDA(ensure_dyn_smem(k_tc_linear_ts<E>, smemBytes, configured));                                          \
    k_tc_linear_ts<E><<<grid, threads, smemBytes, st>>>(mapA, mapWhi, mapWlo, q);                                 \
  } while (0)
  switch (epi) {   // the combinations the GCN path produces; anything else takes the generic kernel
    case 0: ROC_TS_LAUNCH(0); break;
    case EPI_NORM: ROC_TS_LAUNCH(EPI_NORM); break;
    case EPI_OMASK: ROC_TS_LAUNCH(EPI_OMASK); break;
    case EPI_OMASK | EPI_RELUOF | EPI_NORM: ROC_TS_LAUNCH(EPI_OMASK | EPI_RELUOF | EPI_NORM); break;
    default: ROC_TS_LAUNCH(-1); break;
  }
#undef ROC_TS_LAUNCH
  ROC_LAUNCH_CHECK();
  return ROC_OK;
}

int tc_linear_fwd(int64_t rows, int inDim, int outDim, const float* X, int64_t ldX, const float* W, float* Y,
                  int64_t ldY, int relu, const uint64_t* rowEnd, uint64_t colLeft, const DropMask* dm,
                  cudaStream_t st) {
  TsEpilogue e;
  e.relu = relu; e.rowEnd = rowEnd; e.colLeft = colLeft;
  return ts_gemm(rows, inDim, outDim, X, ldX, W, inDim, 0, Y, ldY, dm, e, st);
}

// dX (+)= dY W, optionally followed in the epilogue by the dropout backward of X's producer
int tc_linear_dx(int64_t rows, int inDim, int outDim, const float* dY, int64_t ldDY, const float* W, float* dX,
                 int64_t ldDX, int accumulate, const DropMask* dm, const float* reluOf, int64_t ldR,
                 const uint64_t* rowEnd, uint64_t colLeft, cudaStream_t st) {
  TsEpilogue e;
  e.outMask = dm; e.accumulate = accumulate;
  e.reluOf = reluOf; e.ldR = ldR; e.rowEnd = rowEnd; e.colLeft = colLeft;
  return ts_gemm(rows, outDim, inDim, dY, ldDY, W, inDim, 1, dX, ldDX, nullptr, e, st);
}

}  // namespace roc
