// halo.cu — boundary-node ("halo") structures for the vertex-range partitioned graph.
//
// The reference gives every partition the WHOLE feature matrix before ScatterGather
// (scattergather.cc:69-73 requests inputs[0].region, Legion copies all N rows through
// host memory).  Here a partition receives only the rows it actually reads: the sorted
// set of distinct source ids outside its own vertex range.  The canonical CSR (global
// ids) is untouched; a private remapped copy of col indexes [own rows | halo rows].
//   col_local[e] = colSrc[e] - rowLeft                      if rowLeft <= colSrc[e] <= rowRight
//                = Nloc + rank of colSrc[e] in haloIds      otherwise
// Because partitions are contiguous id ranges, the sorted halo is automatically grouped
// by owner, so each owner's rows land in one contiguous slab.
#include <cub/cub.cuh>
#include <new>
#include "common.cuh"

struct roc_halo {
  uint32_t nHalo = 0;
  uint64_t nEdges = 0;
  uint32_t* ids = nullptr;       // [nHalo] sorted distinct remote source ids (global)
  uint32_t* colLocal = nullptr;  // [nEdges]
};

namespace roc {

struct IsRemote {
  uint32_t lo, hi;
  __host__ __device__ bool operator()(const uint32_t& v) const { return v < lo || v > hi; }
};

__global__ void __launch_bounds__(256)
k_remap_col(uint64_t nEdges, uint32_t rowLeft, uint32_t rowRight, uint32_t nloc, uint32_t nHalo,
            const uint32_t* __restrict__ ids, const uint32_t* __restrict__ col, uint32_t* __restrict__ out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nEdges; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t s = col[i];
    uint32_t r;
    if (s >= rowLeft && s <= rowRight) {
      r = s - rowLeft;
    } else {
      uint32_t lo = 0, hi = nHalo;   // ids is sorted and contains s
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (ids[mid] < s) lo = mid + 1; else hi = mid;
      }
      r = nloc + lo;
    }
    out[i] = r;
  }
}

// dst[j][0:H] = src[rows[j]][0:H]  (whole float4s when the layout allows)
template <int VEC>
__global__ void __launch_bounds__(256)
k_pack_rows(int64_t nRows, int Wq, const uint32_t* __restrict__ rows, const float* __restrict__ src, int64_t ldSrc,
            float* __restrict__ dst, int64_t ldDst) {
  const int64_t total = nRows * (int64_t)Wq;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i / Wq;
    const int c = (int)(i - j * Wq);
    const int64_t r = rows[j];
    if (VEC == 4)
      *reinterpret_cast<float4*>(dst + j * ldDst + 4 * c) = *reinterpret_cast<const float4*>(src + r * ldSrc + 4 * c);
    else
      dst[j * ldDst + c] = src[r * ldSrc + c];
  }
}

// dst[sel[t]][0:H] = src[rows[sel[t]]][0:H]: packs a SUBSET of the send list (the rows of one row block) into their
// places in the send buffer, so that block can travel while the producer computes the next one.
template <int VEC>
__global__ void __launch_bounds__(256)
k_pack_rows_at(int64_t nSel, int Wq, const uint32_t* __restrict__ sel, const uint32_t* __restrict__ rows,
               const float* __restrict__ src, int64_t ldSrc, float* __restrict__ dst, int64_t ldDst) {
  const int64_t total = nSel * (int64_t)Wq;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / Wq;
    const int c = (int)(i - t * Wq);
    const int64_t j = sel[t];
    const int64_t r = rows[j];
    if (VEC == 4)
      *reinterpret_cast<float4*>(dst + j * ldDst + 4 * c) = __ldg(reinterpret_cast<const float4*>(src + r * ldSrc + 4 * c));
    else
      dst[j * ldDst + c] = src[r * ldSrc + c];
  }
}

// Fused pack + exchange: row srcRows[j] of this partition goes straight into the halo slab of the partition
// that reads it — peerBase[peer[j]] is that GPU's tensor (mapped through CUDA IPC; the stores travel over
// NVLink), dstRow[j] the row's place in its slab.  No staging buffer, no NCCL copy kernel: one pass over the
// rows at NVLink speed.  A warp writes whole 16-byte segments of consecutive columns (coalesced 256-byte rows).
struct PeerBases { float* p[ROC_MAX_PEERS]; };

// LR lanes share a row (one float4 column each; rows wider than LR float4 take several passes); every thread keeps
// U rows in flight — loads of the U row ids / destinations, then the U 16-byte loads, then the U peer stores — so
// that the few SMs a pipelined producer leaves free still fill NVLink: with one row per thread and iteration the
// kernel needs ~140 K threads in flight for 770 GB/s (3 us per dependent load -> load -> store chain).
template <int VEC, int LR, int U>
__global__ void __launch_bounds__(256)
k_push_rows(int64_t nRows, int Wq, const uint32_t* __restrict__ rows, const uint8_t* __restrict__ peer,
            const uint32_t* __restrict__ dstRow, const float* __restrict__ src, int64_t ldSrc, PeerBases bases,
            int64_t ldDst) {
  constexpr int RPB = 256 / LR;                 // rows per CTA and step
  const int c0 = threadIdx.x % LR;
  const int64_t slot = threadIdx.x / LR;
  const int64_t stride = (int64_t)gridDim.x * RPB;
  for (int64_t jb = blockIdx.x * (int64_t)RPB + slot; jb < nRows; jb += stride * U) {
    int64_t r[U];
    float* d[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t j = jb + (int64_t)u * stride;
      const bool ok = j < nRows;
      r[u] = ok ? (int64_t)rows[j] : -1;
      d[u] = ok ? bases.p[peer[j]] + (int64_t)dstRow[j] * ldDst : nullptr;
    }
    for (int c = c0; c < Wq; c += LR) {
      if (VEC == 4) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
          if (r[u] >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(src + r[u] * ldSrc + 4 * c));
#pragma unroll
        for (int u = 0; u < U; u++)
          if (r[u] >= 0) *reinterpret_cast<float4*>(d[u] + 4 * c) = v[u];
      } else {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; u++) if (r[u] >= 0) v[u] = __ldg(src + r[u] * ldSrc + c);
#pragma unroll
        for (int u = 0; u < U; u++) if (r[u] >= 0) d[u][c] = v[u];
      }
    }
  }
  __threadfence_system();   // the rows are in the peers' memory before this kernel counts as finished
}

}  // namespace roc

using namespace roc;

extern "C" int roc_halo_create(roc_vid_t rowLeft, roc_vid_t rowRight, uint64_t nEdges, const roc_vid_t* colSrc,
                               roc_stream_t stream, roc_halo** out) {
  if (!out || rowRight < rowLeft || (nEdges && !colSrc)) return ROC_ERR_INVALID;
  if (roc_device_count() <= 0) return ROC_ERR_NO_DEVICE;
  if (nEdges >= 0x7FFFFF00ull) return ROC_ERR_UNSUPPORTED;   // cub item counts are int
  cudaStream_t st = as_stream(stream);
  roc_halo* h = new (std::nothrow) roc_halo();
  if (!h) return ROC_ERR_NOMEM;
  h->nEdges = nEdges;
  uint32_t *remote = nullptr, *sorted = nullptr, *dcount = nullptr;
  void* tmp = nullptr;
  int rc = ROC_OK;
#define H_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = (int)e_; goto done; } } while (0)
  {
    const int n = (int)nEdges;
    uint32_t nRemote = 0;
    H_CUDA(cudaMalloc(&h->colLocal, sizeof(uint32_t) * (size_t)(nEdges ? nEdges : 1)));
    H_CUDA(cudaMalloc(&dcount, sizeof(uint32_t)));
    if (n > 0) {
      H_CUDA(cudaMalloc(&remote, sizeof(uint32_t) * (size_t)nEdges));
      H_CUDA(cudaMalloc(&sorted, sizeof(uint32_t) * (size_t)nEdges));
      size_t t1 = 0, t2 = 0, t3 = 0;
      IsRemote pred{rowLeft, rowRight};
      cub::DeviceSelect::If(nullptr, t1, colSrc, remote, dcount, n, pred, st);
      cub::DeviceRadixSort::SortKeys(nullptr, t2, remote, sorted, n, 0, 32, st);
      cub::DeviceSelect::Unique(nullptr, t3, sorted, remote, dcount, n, st);
      size_t tb = t1 > t2 ? t1 : t2;
      if (t3 > tb) tb = t3;
      H_CUDA(cudaMalloc(&tmp, tb ? tb : 16));
      H_CUDA(cub::DeviceSelect::If(tmp, t1, colSrc, remote, dcount, n, pred, st));
      H_CUDA(cudaMemcpyAsync(&nRemote, dcount, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      H_CUDA(cudaStreamSynchronize(st));
      if (nRemote) {
        H_CUDA(cub::DeviceRadixSort::SortKeys(tmp, t2, remote, sorted, (int)nRemote, 0, 32, st));
        H_CUDA(cub::DeviceSelect::Unique(tmp, t3, sorted, remote, dcount, (int)nRemote, st));
        H_CUDA(cudaMemcpyAsync(&h->nHalo, dcount, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        H_CUDA(cudaStreamSynchronize(st));
      }
      count_launch(6);
    }
    H_CUDA(cudaMalloc(&h->ids, sizeof(uint32_t) * (size_t)(h->nHalo ? h->nHalo : 1)));
    if (h->nHalo)
      H_CUDA(cudaMemcpyAsync(h->ids, remote, sizeof(uint32_t) * (size_t)h->nHalo, cudaMemcpyDeviceToDevice, st));
    if (n > 0) {
      int64_t blocks = ((int64_t)nEdges + 255) / 256;
      if (blocks > sm_count() * 16) blocks = sm_count() * 16;
      k_remap_col<<<(unsigned)blocks, 256, 0, st>>>(nEdges, rowLeft, rowRight, rowRight - rowLeft + 1, h->nHalo, h->ids,
                                                  colSrc, h->colLocal);
      count_launch();
      H_CUDA(cudaGetLastError());
    }
    H_CUDA(cudaStreamSynchronize(st));
  }
done:
#undef H_CUDA
  cudaFree(remote); cudaFree(sorted); cudaFree(dcount); cudaFree(tmp);
  if (rc != ROC_OK) { roc_halo_destroy(h); return rc; }
  *out = h;
  return ROC_OK;
}

extern "C" void roc_halo_destroy(roc_halo* h) {
  if (!h) return;
  cudaFree(h->ids); cudaFree(h->colLocal);
  delete h;
}

extern "C" uint32_t roc_halo_size(const roc_halo* h) { return h ? h->nHalo : 0; }
extern "C" const roc_vid_t* roc_halo_ids(const roc_halo* h) { return h ? h->ids : nullptr; }
extern "C" const roc_vid_t* roc_halo_col_local(const roc_halo* h) { return h ? h->colLocal : nullptr; }

extern "C" int roc_pack_rows(int64_t nRows, int H, const roc_vid_t* rows, const float* src, int64_t ldSrc, float* dst,
                             int64_t ldDst, roc_stream_t stream) {
  if (nRows < 0 || H <= 0 || ldSrc < H || ldDst < H) return ROC_ERR_INVALID;
  if (nRows == 0) return ROC_OK;
  if (!rows || !src || !dst) return ROC_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  const bool vec = (ldSrc % 4 == 0) && (ldDst % 4 == 0) && aligned16(src) && aligned16(dst);
  const int Wq = vec ? (H + 3) / 4 : H;
  int64_t blocks = (nRows * Wq + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  if (vec) k_pack_rows<4><<<(unsigned)blocks, 256, 0, st>>>(nRows, Wq, rows, src, ldSrc, dst, ldDst);
  else k_pack_rows<1><<<(unsigned)blocks, 256, 0, st>>>(nRows, Wq, rows, src, ldSrc, dst, ldDst);
  ROC_LAUNCH_CHECK();
  return ROC_OK;
}

extern "C" int roc_push_rows(int64_t nRows, int H, const roc_vid_t* srcRows, const uint8_t* peer, const roc_vid_t* dstRow,
                             const float* src, int64_t ldSrc, float* const* host_peerBase, int numPeers, int64_t ldDst,
                             int smLimit, roc_stream_t stream) {
  if (nRows < 0 || H <= 0 || ldSrc < H || ldDst < H || numPeers < 1 || numPeers > ROC_MAX_PEERS) return ROC_ERR_INVALID;
  if (nRows == 0) return ROC_OK;
  if (!srcRows || !peer || !dstRow || !src || !host_peerBase) return ROC_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  PeerBases b;
  bool vec = (ldSrc % 4 == 0) && (ldDst % 4 == 0) && aligned16(src);
  for (int q = 0; q < ROC_MAX_PEERS; q++) {
    b.p[q] = q < numPeers ? host_peerBase[q] : nullptr;
    if (b.p[q] && !aligned16(b.p[q])) vec = false;
  }
  const int Wq = vec ? (H + 3) / 4 : H;
  // NVLink-bound.  Beside a pipelined producer the grid is what fits on the SMs that producer left free (its CTAs
  // hold whole SMs; ours would only queue behind them, or, worse, take SMs first and keep its CTAs out).
  const int lr = Wq <= 16 ? 16 : 32;
  int64_t blocks = (nRows + (256 / lr) * 8 - 1) / ((256 / lr) * 8);
  const int64_t cap = (int64_t)(smLimit > 0 ? smLimit : sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (vec) {
    if (lr == 16) k_push_rows<4, 16, 8><<<(unsigned)blocks, 256, 0, st>>>(nRows, Wq, srcRows, peer, dstRow, src, ldSrc, b, ldDst);
    else k_push_rows<4, 32, 8><<<(unsigned)blocks, 256, 0, st>>>(nRows, Wq, srcRows, peer, dstRow, src, ldSrc, b, ldDst);
  } else {
    if (lr == 16) k_push_rows<1, 16, 8><<<(unsigned)blocks, 256, 0, st>>>(nRows, Wq, srcRows, peer, dstRow, src, ldSrc, b, ldDst);
    else k_push_rows<1, 32, 8><<<(unsigned)blocks, 256, 0, st>>>(nRows, Wq, srcRows, peer, dstRow, src, ldSrc, b, ldDst);
  }
  ROC_LAUNCH_CHECK();
  return ROC_OK;
}

extern "C" int roc_pack_rows_at(int64_t nSel, int H, const roc_vid_t* sel, const roc_vid_t* rows, const float* src,
                                int64_t ldSrc, float* dst, int64_t ldDst, roc_stream_t stream) {
  if (nSel < 0 || H <= 0 || ldSrc < H || ldDst < H) return ROC_ERR_INVALID;
  if (nSel == 0) return ROC_OK;
  if (!sel || !rows || !src || !dst) return ROC_ERR_INVALID;
  cudaStream_t st = as_stream(stream);
  const bool vec = (ldSrc % 4 == 0) && (ldDst % 4 == 0) && aligned16(src) && aligned16(dst);
  const int Wq = vec ? (H + 3) / 4 : H;
  int64_t blocks = (nSel * Wq + 255) / 256;
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  if (vec) k_pack_rows_at<4><<<(unsigned)blocks, 256, 0, st>>>(nSel, Wq, sel, rows, src, ldSrc, dst, ldDst);
  else k_pack_rows_at<1><<<(unsigned)blocks, 256, 0, st>>>(nSel, Wq, sel, rows, src, ldSrc, dst, ldDst);
  ROC_LAUNCH_CHECK();
  return ROC_OK;
}
