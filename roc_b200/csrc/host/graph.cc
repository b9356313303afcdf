// graph.cc — Graph: .lux header + row ends on the host, the reference's
// edge-balanced vertex-range partition, this process's CSR slice in HBM.
// Mirrors Graph::Graph (gnn.cc:751-872), load_graph_impl (load_task.cu:201-245)
// and init_task_impl's CSR build (load_task.cu:296-330) without Legion regions.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "host_internal.h"

using namespace roc::host;

Config::Config()
    : numGPUs(0), numMachines(1), totalGPUs(0), numEpochs(1), decay_steps(100), seed(1), verbose(false),
      learning_rate(0.01f), weight_decay(0.05f), dropout_rate(0.5f), decay_rate(1.0f), filename("") {}
// defaults: gnn.cc:31-41

void Graph::build(Context ctx, const E_ID* host_rowEnd, const V_ID* slice_colSrc) {
  RuntimeImpl* rt = ctx;
  // CSR sanity the reference asserts (gnn.cc:798-800)
  for (V_ID v = 1; v < numNodes; v++) ROC_ASSERT(host_rowEnd[v] >= host_rowEnd[v - 1]);
  ROC_ASSERT(host_rowEnd[numNodes - 1] == numEdges);
  vbounds.assign((size_t)numParts * 2, 0);
  ebounds.assign((size_t)numParts * 2, 0);
  int nranges = 0;
  int rc = roc_partition(numNodes, numEdges, numParts, host_rowEnd, vbounds.data(), ebounds.data(), &nranges);
  if (rc != ROC_OK) {
    // gnn.cc:829 assert(bounds.size() == numParts)
    char b[160];
    snprintf(b, sizeof(b), "partitioner produced %d ranges for %d parts (gnn.cc:829)", nranges, numParts);
    ROC_FATAL(b);
  }
  rowLeft = vbounds[2 * myPart]; rowRight = vbounds[2 * myPart + 1];
  colLeft = ebounds[2 * myPart]; colRight = ebounds[2 * myPart + 1];
  const size_t nloc = (size_t)rowRight - rowLeft + 1;
  const size_t eloc = (size_t)(colRight + 1 - colLeft);
  if (rt->numParts > 1)
    fprintf(stderr, "[roc_b200] part %d/%d: rows [%u, %u] edges [%zu, %zu]\n", myPart, numParts, rowLeft,
            rowRight, (size_t)colLeft, (size_t)colRight);
  // raw slices -> device, then the device CSR build (init_graph_kernel's job)
  E_ID* d_rawRows = (E_ID*)rt->dmalloc(nloc * sizeof(E_ID));
  V_ID* d_rawCols = (V_ID*)rt->dmalloc((eloc ? eloc : 1) * sizeof(V_ID));
  ROC_CHECK(cudaMemcpyAsync(d_rawRows, host_rowEnd + rowLeft, nloc * sizeof(E_ID), cudaMemcpyHostToDevice, rt->stream));
  if (eloc)
    ROC_CHECK(cudaMemcpyAsync(d_rawCols, slice_colSrc, eloc * sizeof(V_ID), cudaMemcpyHostToDevice, rt->stream));
  // The lean layout makes rowPtrs == rawRows and colSrc == rawCols, so the
  // uploaded slices ARE the device CSR; roc_build_csr is only needed to produce
  // the reference's EdgeStruct form, which nothing here consumes.
  d_rowEnd = d_rawRows;
  d_colSrc = d_rawCols;
  ROC_CHECK(cudaStreamSynchronize(rt->stream));
  plan = nullptr;
  halo = nullptr; numHalo = 0; d_sendRows = nullptr; numSendRows = 0;
  d_pushRows = nullptr; d_pushPeer = nullptr; d_pushDst = nullptr; d_packSel = nullptr;
  const char* he = getenv("ROC_B200_HALO");
  const bool useHalo = rt->numParts > 1 && rt->commReady && !(he && he[0] == '0');
  if (!useHalo) {
    ROC_CHECK(roc_sg_plan_create(rowLeft, rowRight, colLeft, d_rowEnd, d_colSrc, rt->stream, &plan));
    return;
  }
  // ---- halo: which remote rows do my edges read, and which of my rows do the others read
  const int P = rt->numParts, me = rt->myPart;
  ROC_CHECK(roc_halo_create(rowLeft, rowRight, (uint64_t)eloc, d_colSrc, rt->stream, &halo));
  numHalo = roc_halo_size(halo);
  std::vector<V_ID> hid(numHalo ? numHalo : 1);
  if (numHalo)
    ROC_CHECK(cudaMemcpyAsync(hid.data(), roc_halo_ids(halo), (size_t)numHalo * sizeof(V_ID), cudaMemcpyDeviceToHost, rt->stream));
  ROC_CHECK(cudaStreamSynchronize(rt->stream));
  recvCounts.assign((size_t)P, 0); recvOffs.assign((size_t)P, 0);
  sendCounts.assign((size_t)P, 0); sendOffs.assign((size_t)P, 0);
  {
    std::vector<uint64_t> rc((size_t)P), ro((size_t)P);
    ROC_CHECK(roc_halo_recv_layout(numHalo, hid.data(), P, me, vbounds.data(), rc.data(), ro.data()));
    for (int q = 0; q < P; q++) { recvCounts[(size_t)q] = (size_t)rc[(size_t)q]; recvOffs[(size_t)q] = (size_t)ro[(size_t)q]; }
  }
  // everyone learns everyone's request counts (P x P matrix), then the id lists travel to their owners
  int* d_cnt = (int*)rt->dmalloc(sizeof(int) * (size_t)P * (size_t)(P + 1));
  std::vector<int> mine((size_t)P), all((size_t)P * P);
  for (int q = 0; q < P; q++) mine[(size_t)q] = (int)recvCounts[(size_t)q];
  ROC_CHECK(cudaMemcpyAsync(d_cnt, mine.data(), sizeof(int) * P, cudaMemcpyHostToDevice, rt->stream));
  ROC_CHECK(rt->comm.allgather_i32(d_cnt, d_cnt + P, (size_t)P, rt->stream));
  ROC_CHECK(cudaMemcpyAsync(all.data(), d_cnt + P, sizeof(int) * (size_t)P * P, cudaMemcpyDeviceToHost, rt->stream));
  ROC_CHECK(cudaStreamSynchronize(rt->stream));
  {
    std::vector<uint64_t> sc((size_t)P), so((size_t)P);
    uint64_t total = 0;
    ROC_CHECK(roc_halo_send_layout(P, me, all.data(), sc.data(), so.data(), &total));
    for (int q = 0; q < P; q++) { sendCounts[(size_t)q] = (size_t)sc[(size_t)q]; sendOffs[(size_t)q] = (size_t)so[(size_t)q]; }
    numSendRows = (size_t)total;
  }
  d_sendRows = (V_ID*)rt->dmalloc(sizeof(V_ID) * (numSendRows ? numSendRows : 1));
  ROC_CHECK(rt->comm.alltoallv(roc_halo_ids(halo), recvCounts, recvOffs, d_sendRows, sendCounts, sendOffs,
                               /*isFloat=*/false, rt->stream));
  {
    std::vector<V_ID> rows(numSendRows ? numSendRows : 1);
    if (numSendRows)
      ROC_CHECK(cudaMemcpyAsync(rows.data(), d_sendRows, numSendRows * sizeof(V_ID), cudaMemcpyDeviceToHost, rt->stream));
    ROC_CHECK(cudaStreamSynchronize(rt->stream));
    for (size_t j = 0; j < numSendRows; j++) {
      ROC_ASSERT(rows[j] >= rowLeft && rows[j] <= rowRight);
      rows[j] -= rowLeft;                                       // global id -> my local row
    }
    if (numSendRows)
      ROC_CHECK(cudaMemcpyAsync(d_sendRows, rows.data(), numSendRows * sizeof(V_ID), cudaMemcpyHostToDevice, rt->stream));
    ROC_CHECK(cudaStreamSynchronize(rt->stream));
  }
  // ---- peer-write exchange: where does each send row land in its reader's slab?
  {
    // every rank's recvOffs (where owner r's rows start inside rank q's halo): [q][r]
    std::vector<int> myOffs((size_t)P), allOffs((size_t)P * P);
    for (int q = 0; q < P; q++) myOffs[(size_t)q] = (int)recvOffs[(size_t)q];
    ROC_CHECK(cudaMemcpyAsync(d_cnt, myOffs.data(), sizeof(int) * P, cudaMemcpyHostToDevice, rt->stream));
    ROC_CHECK(rt->comm.allgather_i32(d_cnt, d_cnt + P, (size_t)P, rt->stream));
    ROC_CHECK(cudaMemcpyAsync(allOffs.data(), d_cnt + P, sizeof(int) * (size_t)P * P, cudaMemcpyDeviceToHost, rt->stream));
    std::vector<V_ID> rows(numSendRows ? numSendRows : 1);
    if (numSendRows)
      ROC_CHECK(cudaMemcpyAsync(rows.data(), d_sendRows, numSendRows * sizeof(V_ID), cudaMemcpyDeviceToHost, rt->stream));
    ROC_CHECK(cudaStreamSynchronize(rt->stream));
    struct Ent { V_ID row; V_ID dst; unsigned char peer; };
    std::vector<Ent> ents(numSendRows);
    for (int q = 0; q < P; q++) {
      const V_ID nlocQ = vbounds[2 * q + 1] - vbounds[2 * q] + 1;
      const V_ID slab0 = nlocQ + (V_ID)allOffs[(size_t)q * P + me];       // my rows' place in q's [own | halo] slab
      for (size_t j = 0; j < sendCounts[(size_t)q]; j++) {
        Ent& e = ents[sendOffs[(size_t)q] + j];
        e.row = rows[sendOffs[(size_t)q] + j]; e.dst = slab0 + (V_ID)j; e.peer = (unsigned char)q;
      }
    }
    std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.row < b.row; });
    std::vector<V_ID> hr(numSendRows ? numSendRows : 1), hd(numSendRows ? numSendRows : 1);
    std::vector<unsigned char> hp(numSendRows ? numSendRows : 1);
    for (size_t j = 0; j < numSendRows; j++) { hr[j] = ents[j].row; hd[j] = ents[j].dst; hp[j] = ents[j].peer; }
    d_pushRows = (V_ID*)rt->dmalloc(hr.size() * sizeof(V_ID));
    d_pushDst = (V_ID*)rt->dmalloc(hd.size() * sizeof(V_ID));
    d_pushPeer = (unsigned char*)rt->dmalloc(hp.size());
    ROC_CHECK(cudaMemcpyAsync(d_pushRows, hr.data(), hr.size() * sizeof(V_ID), cudaMemcpyHostToDevice, rt->stream));
    ROC_CHECK(cudaMemcpyAsync(d_pushDst, hd.data(), hd.size() * sizeof(V_ID), cudaMemcpyHostToDevice, rt->stream));
    ROC_CHECK(cudaMemcpyAsync(d_pushPeer, hp.data(), hp.size(), cudaMemcpyHostToDevice, rt->stream));
    // row blocks (multiples of 128 rows: whole GEMM tiles) and the CSR offset in front of each
    int K = 4;
    if (const char* e = getenv("ROC_B200_PUSH_BLOCKS")) K = std::max(1, atoi(e));
    size_t per = (nloc + (size_t)K - 1) / (size_t)K;
    per = (per + 127) / 128 * 128;
    pushBlockRow.clear(); pushBlockOff.clear(); pushBlockColLeft.clear();
    for (size_t r = 0; r < nloc; r += per) pushBlockRow.push_back((V_ID)r);
    pushBlockRow.push_back((V_ID)nloc);
    const size_t nb = pushBlockRow.size() - 1;
    for (size_t k = 0; k <= nb; k++)
      pushBlockOff.push_back((size_t)(std::lower_bound(hr.begin(), hr.begin() + numSendRows, pushBlockRow[k]) - hr.begin()));
    for (size_t k = 0; k < nb; k++)
      pushBlockColLeft.push_back(pushBlockRow[k] == 0 ? colLeft : host_rowEnd[rowLeft + pushBlockRow[k] - 1]);
    // copy-engine exchange: the requester-grouped list, cut by row block
    {
      std::vector<V_ID> sel;
      sel.reserve(numSendRows ? numSendRows : 1);
      packBlockOff.assign(nb + 1, 0);
      packPeerOff.assign((size_t)P * (nb + 1), 0);
      peerSlab0.assign((size_t)P, 0);
      for (int q = 0; q < P; q++) {
        const V_ID nlocQ = vbounds[2 * q + 1] - vbounds[2 * q] + 1;
        peerSlab0[(size_t)q] = nlocQ + (V_ID)allOffs[(size_t)q * P + me];
        const V_ID* b = rows.data() + sendOffs[(size_t)q];
        const V_ID* e = b + sendCounts[(size_t)q];
        for (size_t k = 0; k <= nb; k++)
          packPeerOff[(size_t)q * (nb + 1) + k] = sendOffs[(size_t)q] + (size_t)(std::lower_bound(b, e, pushBlockRow[k]) - b);
      }
      for (size_t k = 0; k < nb; k++) {
        packBlockOff[k] = sel.size();
        for (int q = 0; q < P; q++)
          for (size_t j = packPeerOff[(size_t)q * (nb + 1) + k]; j < packPeerOff[(size_t)q * (nb + 1) + k + 1]; j++) sel.push_back((V_ID)j);
      }
      packBlockOff[nb] = sel.size();
      ROC_ASSERT(sel.size() == numSendRows);
      d_packSel = (V_ID*)rt->dmalloc((sel.size() ? sel.size() : 1) * sizeof(V_ID));
      if (!sel.empty())
        ROC_CHECK(cudaMemcpyAsync(d_packSel, sel.data(), sel.size() * sizeof(V_ID), cudaMemcpyHostToDevice, rt->stream));
      ROC_CHECK(cudaStreamSynchronize(rt->stream));
    }
    ROC_CHECK(cudaStreamSynchronize(rt->stream));
  }
  fprintf(stderr, "[roc_b200] part %d/%d: halo %u rows (%.1f%% of the %u remote vertices), sends %zu rows\n", me, P,
          numHalo, 100.0 * numHalo / std::max<double>(1.0, (double)numNodes - (double)nloc), (unsigned)(numNodes - nloc),
          numSendRows);
  // the kernel indexes [own rows | halo rows]; rowEnd offsets are unchanged
  ROC_CHECK(roc_sg_plan_create(rowLeft, rowRight, colLeft, d_rowEnd, roc_halo_col_local(halo), rt->stream, &plan));
}

Graph::Graph(Context ctx, Runtime* /*runtime*/, const Config& config)
    : numParts(config.totalGPUs), numMachines(config.numMachines), maxHidden(0) {
  RuntimeImpl* rt = ctx;
  if (numParts <= 0) numParts = rt->numParts;
  ROC_ASSERT(numParts == rt->numParts);
  myPart = rt->myPart;
  std::string luxfilename = config.filename + ".add_self_edge.lux";
  printf("Lux Filename: %s\n", luxfilename.c_str());
  FILE* fd = fopen(luxfilename.c_str(), "rb");
  if (!fd) ROC_FATAL(("cannot open " + luxfilename).c_str());
  ROC_ASSERT(fread(&numNodes, sizeof(V_ID), 1, fd) == 1);
  ROC_ASSERT(fread(&numEdges, sizeof(E_ID), 1, fd) == 1);
  fprintf(stderr, "[roc_b200] Load graph: numNodes(%u) numEdges(%zu)\n", numNodes, (size_t)numEdges);
  std::vector<E_ID> raw_rows(numNodes);
  ROC_ASSERT(fread(raw_rows.data(), sizeof(E_ID), (size_t)numNodes, fd) == (size_t)numNodes);
  // partition first (needs only the row ends), then read just this part's sources
  vbounds.assign((size_t)numParts * 2, 0);
  ebounds.assign((size_t)numParts * 2, 0);
  int nr = 0;
  if (roc_partition(numNodes, numEdges, numParts, raw_rows.data(), vbounds.data(), ebounds.data(), &nr) != ROC_OK)
    ROC_FATAL("partitioner did not produce numParts ranges (gnn.cc:829)");
  E_ID cl = ebounds[2 * myPart], cr = ebounds[2 * myPart + 1];
  size_t eloc = (size_t)(cr + 1 - cl);
  std::vector<V_ID> cols(eloc ? eloc : 1);
  if (eloc) {
    // load_task.cu:236-243
    ROC_ASSERT(fseeko(fd, (off_t)(FILE_HEADER_SIZE + sizeof(E_ID) * (size_t)numNodes + sizeof(V_ID) * (size_t)cl), SEEK_SET) == 0);
    ROC_ASSERT(fread(cols.data(), sizeof(V_ID), eloc, fd) == eloc);
  }
  fclose(fd);
  build(ctx, raw_rows.data(), cols.data());
}

Graph::Graph(Context ctx, Runtime* /*runtime*/, V_ID _numNodes, E_ID _numEdges, const E_ID* host_rowEnd,
             const V_ID* host_colSrc)
    : numNodes(_numNodes), numEdges(_numEdges), numMachines(1), maxHidden(0) {
  RuntimeImpl* rt = ctx;
  numParts = rt->numParts;
  myPart = rt->myPart;
  ROC_ASSERT(host_rowEnd != nullptr && numNodes > 0);
  // partition to find this part's slice of the full source array
  std::vector<V_ID> vb((size_t)numParts * 2);
  std::vector<E_ID> eb((size_t)numParts * 2);
  int nr = 0;
  if (roc_partition(numNodes, numEdges, numParts, host_rowEnd, vb.data(), eb.data(), &nr) != ROC_OK)
    ROC_FATAL("partitioner did not produce numParts ranges (gnn.cc:829)");
  build(ctx, host_rowEnd, host_colSrc ? host_colSrc + eb[2 * myPart] : nullptr);
}

// ---- host-side bookkeeping of the halo exchange (plain C ABI, host pointers; no device work) ----
// The sorted halo id list is grouped by owner because partitions are contiguous vertex ranges:
// recvOffs[q] = index of the first halo row owned by partition q, recvCounts[q] = how many.
extern "C" int roc_halo_recv_layout(uint32_t nHalo, const roc_vid_t* host_ids, int numParts, int myPart,
                                    const roc_vid_t* host_vbounds, uint64_t* recvCounts, uint64_t* recvOffs) {
  if (numParts <= 0 || myPart < 0 || myPart >= numParts || !host_vbounds || !recvCounts || !recvOffs) return ROC_ERR_INVALID;
  if (nHalo && !host_ids) return ROC_ERR_INVALID;
  uint64_t k = 0;
  for (int q = 0; q < numParts; q++) {
    recvOffs[q] = k;
    while (k < nHalo && host_ids[k] <= host_vbounds[2 * q + 1]) {
      if (host_ids[k] < host_vbounds[2 * q]) return ROC_ERR_INVALID;      // not sorted / not inside any range
      if (k > 0 && host_ids[k] <= host_ids[k - 1]) return ROC_ERR_INVALID; // must be strictly increasing
      k++;
    }
    recvCounts[q] = k - recvOffs[q];
  }
  if (k != nHalo || recvCounts[myPart] != 0) return ROC_ERR_INVALID;       // an own row is never a halo row
  return ROC_OK;
}

// allCounts[q * P + r] = rows partition q requests from owner r (every rank's recvCounts, all-gathered).
// This rank (me) packs, for q = 0..P-1 in order, the rows q asked of it: sendCounts[q] = allCounts[q][me].
extern "C" int roc_halo_send_layout(int numParts, int myPart, const int32_t* host_allCounts, uint64_t* sendCounts,
                                    uint64_t* sendOffs, uint64_t* numSendRows) {
  if (numParts <= 0 || myPart < 0 || myPart >= numParts || !host_allCounts || !sendCounts || !sendOffs) return ROC_ERR_INVALID;
  uint64_t total = 0;
  for (int q = 0; q < numParts; q++) {
    const int32_t c = host_allCounts[(size_t)q * numParts + myPart];
    if (c < 0 || (q == myPart && c != 0)) return ROC_ERR_INVALID;
    sendOffs[q] = total;
    sendCounts[q] = (uint64_t)c;
    total += (uint64_t)c;
  }
  if (numSendRows) *numSendRows = total;
  return ROC_OK;
}
