// RankLink (multi_host.h): NCCL + CUDA-runtime plumbing of one rank of a multi-GPU render group. Plumbing only — no
// arithmetic on samples happens here.
#include "multi_host.h"

#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: the entry points are bound with dlsym below

#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>

namespace newman_b200 {
namespace {

struct NcclApi {
  void* lib = nullptr;
  std::string err;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;

  NcclApi() {
    // "libnccl.so.2" resolves to a copy already mapped into the process under that SONAME (PyTorch's bundled one when
    // the caller is a torchrun rank), else to the system library
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
#define NM_BIND(field, sym)                                                   \
  field = reinterpret_cast<decltype(field)>(dlsym(lib, sym));                 \
  if (!field) { err = std::string("libnccl lacks ") + sym; return; }
    NM_BIND(GetUniqueId, "ncclGetUniqueId") NM_BIND(CommInitRank, "ncclCommInitRank") NM_BIND(CommDestroy, "ncclCommDestroy")
    NM_BIND(GetErrorString, "ncclGetErrorString") NM_BIND(Broadcast, "ncclBroadcast") NM_BIND(AllReduce, "ncclAllReduce")
    NM_BIND(Send, "ncclSend") NM_BIND(Recv, "ncclRecv") NM_BIND(GroupStart, "ncclGroupStart") NM_BIND(GroupEnd, "ncclGroupEnd")
    NM_BIND(GetVersion, "ncclGetVersion")
#undef NM_BIND
  }
};

const NcclApi& nccl() {
  static NcclApi api;
  if (!api.err.empty()) throw std::runtime_error("newman_b200 multi-GPU: " + api.err);
  return api;
}

void cu(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("newman_b200 multi-GPU: ") + what + ": " + cudaGetErrorString(e));
}
void nc(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw std::runtime_error(std::string("newman_b200 multi-GPU: ") + what + ": " + nccl().GetErrorString(r));
}
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
struct Timed {
  double& acc; double t0;
  explicit Timed(double& a) : acc(a), t0(now_ms()) {}
  ~Timed() { acc += now_ms() - t0; }
};
}  // namespace

void RankLink::unique_id(uint8_t id[NMM_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) == NMM_ID_BYTES, "NMM_ID_BYTES must be sizeof(ncclUniqueId)");
  ncclUniqueId u;
  nc(nccl().GetUniqueId(&u), "ncclGetUniqueId");
  memcpy(id, &u, NMM_ID_BYTES);
}

RankLink::RankLink(int device_, int rank_, int world_, const uint8_t* nccl_id)
    : rank(rank_), world(world_), device(device_), ctx(nullptr), comm_(nullptr), stream_(nullptr), hsmall_(nullptr),
      hbig_(nullptr), hbig_cap_(0), exchange_ms_(0.0) {
  for (Buf& b : stage_) b = Buf{nullptr, 0};
  for (Buf& b : rows_) b = Buf{nullptr, 0};
  band_ = gather_ = small_ = Buf{nullptr, 0};
  if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("newman_b200 multi-GPU: bad rank / world");
  if (nm_create(device, &ctx) != NM_OK) throw std::runtime_error(std::string("newman_b200 multi-GPU: ") + nm_last_error(nullptr));
  void* s = nullptr;
  nm_get_stream(ctx, &s);
  stream_ = s;
  cu(cudaSetDevice(device), "cudaSetDevice");
  cu(cudaMallocHost(&hsmall_, 65536), "cudaMallocHost");
  ensure(small_, 65536);
  if (world > 1) {
    if (!nccl_id) throw std::runtime_error("newman_b200 multi-GPU: world > 1 needs the group's NCCL id");
    ncclUniqueId u;
    memcpy(&u, nccl_id, NMM_ID_BYTES);
    ncclComm_t c = nullptr;
    nc(nccl().CommInitRank(&c, world, u, rank), "ncclCommInitRank");
    comm_ = c;
  }
}

RankLink::~RankLink() {
  cudaSetDevice(device);
  if (stream_) cudaStreamSynchronize((cudaStream_t)stream_);
  if (comm_) nccl().CommDestroy((ncclComm_t)comm_);
  for (Buf& b : stage_) if (b.p) cudaFree(b.p);
  for (Buf& b : rows_) if (b.p) cudaFree(b.p);
  if (band_.p) cudaFree(band_.p);
  if (gather_.p) cudaFree(gather_.p);
  if (small_.p) cudaFree(small_.p);
  if (hsmall_) cudaFreeHost(hsmall_);
  if (hbig_) cudaFreeHost(hbig_);
  if (ctx) nm_destroy(ctx);
}

void RankLink::ensure(Buf& b, size_t bytes) {
  if (bytes <= b.cap) return;
  cu(cudaSetDevice(device), "cudaSetDevice");
  if (b.p) { cu(cudaStreamSynchronize((cudaStream_t)stream_), "sync"); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  const size_t cap = bytes + bytes / 8 + 256;
  cu(cudaMalloc(&b.p, cap), "cudaMalloc (exchange staging)");
  b.cap = cap;
}

void* RankLink::pinned(size_t bytes) {
  if (bytes <= hbig_cap_) return hbig_;
  cu(cudaSetDevice(device), "cudaSetDevice");
  if (hbig_) { cu(cudaStreamSynchronize((cudaStream_t)stream_), "sync"); cudaFreeHost(hbig_); hbig_ = nullptr; hbig_cap_ = 0; }
  const size_t cap = bytes + bytes / 8 + 4096;
  cu(cudaMallocHost(&hbig_, cap), "cudaMallocHost (table blob)");
  hbig_cap_ = cap;
  return hbig_;
}

void RankLink::sync() {
  cu(cudaSetDevice(device), "cudaSetDevice");
  cu(cudaStreamSynchronize((cudaStream_t)stream_), "cudaStreamSynchronize");
}

void* RankLink::bcast_device(const void* host_src, size_t bytes, int slot) {
  Timed t(exchange_ms_);
  cudaStream_t st = (cudaStream_t)stream_;
  cu(cudaSetDevice(device), "cudaSetDevice");
  Buf& b = stage_[slot & 1];
  ensure(b, bytes);
  if (rank == 0) cu(cudaMemcpyAsync(b.p, host_src, bytes, cudaMemcpyHostToDevice, st), "H2D (tables)");
  if (world > 1) nc(nccl().Broadcast(b.p, b.p, bytes, ncclChar, 0, (ncclComm_t)comm_, st), "ncclBroadcast");
  return b.p;
}

void RankLink::bcast_host(void* host_buf, size_t bytes) {
  if (bytes > 65536) throw std::runtime_error("newman_b200 multi-GPU: bcast_host is for small headers");
  if (world == 1) return;
  Timed t(exchange_ms_);
  cudaStream_t st = (cudaStream_t)stream_;
  cu(cudaSetDevice(device), "cudaSetDevice");
  if (rank == 0) {
    memcpy(hsmall_, host_buf, bytes);
    cu(cudaMemcpyAsync(small_.p, hsmall_, bytes, cudaMemcpyHostToDevice, st), "H2D (header)");
  }
  nc(nccl().Broadcast(small_.p, small_.p, bytes, ncclChar, 0, (ncclComm_t)comm_, st), "ncclBroadcast (header)");
  if (rank != 0) cu(cudaMemcpyAsync(hsmall_, small_.p, bytes, cudaMemcpyDeviceToHost, st), "D2H (header)");
  cu(cudaStreamSynchronize(st), "sync");   // rank 0 too: the pinned scratch is reused by the next exchange
  if (rank != 0) memcpy(host_buf, hsmall_, bytes);
}

static void reduce_u64(RankLink& l, void* comm, void* stream, void* dev, void* pinned, uint64_t* v, int n, ncclRedOp_t op,
                       int device) {
  cudaStream_t st = (cudaStream_t)stream;
  cu(cudaSetDevice(device), "cudaSetDevice");
  memcpy(pinned, v, sizeof(uint64_t) * (size_t)n);
  cu(cudaMemcpyAsync(dev, pinned, sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, st), "H2D (reduce)");
  nc(nccl().AllReduce(dev, dev, (size_t)n, ncclUint64, op, (ncclComm_t)comm, st), "ncclAllReduce");
  cu(cudaMemcpyAsync(pinned, dev, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, st), "D2H (reduce)");
  cu(cudaStreamSynchronize(st), "sync");
  memcpy(v, pinned, sizeof(uint64_t) * (size_t)n);
  (void)l;
}

uint64_t RankLink::allreduce_min(uint64_t v) {
  if (world == 1) return v;
  Timed t(exchange_ms_);
  reduce_u64(*this, comm_, stream_, small_.p, hsmall_, &v, 1, ncclMin, device);
  return v;
}
uint64_t RankLink::allreduce_max(uint64_t v) {
  if (world == 1) return v;
  Timed t(exchange_ms_);
  reduce_u64(*this, comm_, stream_, small_.p, hsmall_, &v, 1, ncclMax, device);
  return v;
}
void RankLink::allreduce_sum(uint64_t* v, int n) {
  if (world == 1) return;
  Timed t(exchange_ms_);
  reduce_u64(*this, comm_, stream_, small_.p, hsmall_, v, n, ncclSum, device);
}
void RankLink::barrier() {
  uint64_t one = 1;
  allreduce_sum(&one, 1);
}

void* RankLink::gather_rows(const void* dev_full, size_t row_bytes, int band, int n_blocks, int slot) {
  cudaStream_t st = (cudaStream_t)stream_;
  cu(cudaSetDevice(device), "cudaSetDevice");
  const int nb = blocks_of(rank, world, n_blocks);
  Buf& b = rows_[slot & 3];
  ensure(b, (size_t)(nb > 0 ? nb : 1) * band * row_bytes);
  if (nb > 0)
    cu(cudaMemcpy2DAsync(b.p, (size_t)band * row_bytes, (const char*)dev_full + (size_t)rank * band * row_bytes,
                         (size_t)world * band * row_bytes, (size_t)band * row_bytes, (size_t)nb, cudaMemcpyDeviceToDevice, st),
       "row gather");
  return b.p;
}

void* RankLink::band_buffer(size_t bytes) {
  ensure(band_, bytes);
  return band_.p;
}

void RankLink::return_band(const void* band_dev, size_t block_bytes, int n_blocks, void* host_full, int mode) {
  Timed t(exchange_ms_);
  cudaStream_t st = (cudaStream_t)stream_;
  cu(cudaSetDevice(device), "cudaSetDevice");
  const int nb = blocks_of(rank, world, n_blocks);
  if (mode == NMM_RETURN_LOCAL || world == 1) {
    if (!host_full) throw std::runtime_error("newman_b200 multi-GPU: NMM_RETURN_LOCAL needs the host image on every rank");
    if (nb > 0)
      cu(cudaMemcpy2DAsync((char*)host_full + (size_t)rank * block_bytes, (size_t)world * block_bytes, band_dev, block_bytes,
                           block_bytes, (size_t)nb, cudaMemcpyDeviceToHost, st),
         "D2H (band)");
    cu(cudaStreamSynchronize(st), "sync");
    return;
  }
  // funnel through rank 0: grouped send / recv over NVLink, then one interleaving D2H per source rank
  if (rank != 0) {
    if (nb > 0) nc(nccl().Send(band_dev, (size_t)nb * block_bytes, ncclChar, 0, (ncclComm_t)comm_, st), "ncclSend (band)");
    cu(cudaStreamSynchronize(st), "sync");
    return;
  }
  if (!host_full) throw std::runtime_error("newman_b200 multi-GPU: rank 0 needs the host image");
  size_t total = 0;
  for (int r = 1; r < world; r++) total += (size_t)blocks_of(r, world, n_blocks) * block_bytes;
  ensure(gather_, total ? total : 1);
  nc(nccl().GroupStart(), "ncclGroupStart");
  size_t off = 0;
  for (int r = 1; r < world; r++) {
    const size_t bytes = (size_t)blocks_of(r, world, n_blocks) * block_bytes;
    if (bytes) nc(nccl().Recv((char*)gather_.p + off, bytes, ncclChar, r, (ncclComm_t)comm_, st), "ncclRecv (band)");
    off += bytes;
  }
  nc(nccl().GroupEnd(), "ncclGroupEnd");
  off = 0;
  for (int r = 0; r < world; r++) {
    const int nbr = blocks_of(r, world, n_blocks);
    const char* src = r == 0 ? (const char*)band_dev : (const char*)gather_.p + off;
    if (nbr > 0)
      cu(cudaMemcpy2DAsync((char*)host_full + (size_t)r * block_bytes, (size_t)world * block_bytes, src, block_bytes, block_bytes,
                           (size_t)nbr, cudaMemcpyDeviceToHost, st),
         "D2H (gathered band)");
    if (r > 0) off += (size_t)nbr * block_bytes;
  }
  cu(cudaStreamSynchronize(st), "sync");
}

}  // namespace newman_b200
