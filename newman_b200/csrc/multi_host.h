// One rank of a multi-GPU render group (C++11 host code): its GPU context, its NCCL communicator and the device
// staging the frame's exchanges go through. The path shards by bands of grid rows with NO collective on the per-pixel
// data path; what is exchanged per frame is
//   * the tables of each reference orbit (host GMP on rank 0)            -> ncclBroadcast over NVLink
//   * which glitched sample becomes the next secondary reference          -> one 8-byte MIN ncclAllReduce
//   * the finished bands (8 B/sample raster or 3 B/pixel RGB)             -> each rank's own D2H into the host raster
//     (ranks that share the caller's address space or a shared mapping), or ncclSend/ncclRecv to rank 0 + one D2H
// Ranks are threads of one process (class Mandelbrot with `devices` set: one host thread + one context per GPU) or
// processes (one per GPU, e.g. under torchrun: the 128-byte NCCL id travels through the caller). NCCL is bound at run
// time (dlopen "libnccl.so.2": the copy already in the process if there is one — e.g. PyTorch's — else the system's).
#ifndef NEWMAN_B200_MULTI_HOST_H
#define NEWMAN_B200_MULTI_HOST_H

#include <cstddef>
#include <cstdint>
#include <string>

#include "../../include/newman_b200.h"

namespace newman_b200 {

class RankLink {
public:
  int rank, world, device;
  nm_ctx* ctx;

  RankLink(int device, int rank, int world, const uint8_t* nccl_id);   // throws std::runtime_error
  ~RankLink();
  RankLink(const RankLink&) = delete;
  RankLink& operator=(const RankLink&) = delete;

  static void unique_id(uint8_t id[NMM_ID_BYTES]);

  // ---- collectives: every rank of the group calls them in the same order (world == 1: local) --------------------
  // rank 0's `bytes` at host_src (ignored elsewhere) land in a device buffer on every rank; the pointer stays valid
  // until the next bcast_device call with the same `slot` (0 or 1: two independent staging buffers)
  void* bcast_device(const void* host_src, size_t bytes, int slot = 0);
  // the same, delivered to a host buffer on every rank (small headers)
  void bcast_host(void* host_buf, size_t bytes);
  uint64_t allreduce_min(uint64_t v);
  uint64_t allreduce_max(uint64_t v);
  void allreduce_sum(uint64_t* v, int n);
  void barrier();

  // ---- bands: blocks of `band` grid rows, block b belongs to rank b % world ---------------------------------------
  static int blocks_of(int rank, int world, int n_blocks) { return n_blocks > rank ? (n_blocks - rank + world - 1) / world : 0; }
  // this rank's rows of a per-row device array (elem_bytes per row) as one contiguous device array (slot 0..3)
  void* gather_rows(const void* dev_full, size_t row_bytes, int band, int n_blocks, int slot);
  // a contiguous device buffer for this rank's finished band (raster or RGB)
  void* band_buffer(size_t bytes);
  // Return this rank's band (n_local blocks of block_bytes each, contiguous at band_dev) to the full host image whose
  // blocks are block_bytes apart in rank order. mode NMM_RETURN_LOCAL: this rank copies its blocks into `host_full`
  // itself (every rank must be able to address it: threads of one process, or a shared mapping); NMM_RETURN_ROOT:
  // ncclSend to rank 0, which writes all bands into ITS host_full (others may pass nullptr).
  void return_band(const void* band_dev, size_t block_bytes, int n_blocks, void* host_full, int mode);

  // grow-only pinned host scratch (rank 0 packs a reference's tables here before bcast_device)
  void* pinned(size_t bytes);
  void sync();
  void* stream() const { return stream_; }
  double exchange_ms() const { return exchange_ms_; }   // host time spent inside the collectives of this rank, cumulative

private:
  void* comm_;      // ncclComm_t
  void* stream_;    // the ctx's stream (cudaStream_t)
  struct Buf { void* p; size_t cap; };
  Buf stage_[2], rows_[4], band_, gather_, small_;
  void* hsmall_;    // pinned scratch (64 KB)
  void* hbig_;      // pinned scratch for table blobs
  size_t hbig_cap_;
  double exchange_ms_;
  void ensure(Buf& b, size_t bytes);
};

}  // namespace newman_b200
#endif
