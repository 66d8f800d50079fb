// TEST INFRASTRUCTURE. A file-based stand-in for the nine NCCL entry points the engine binds (spsph_engine.cu,
// spsph_dist_init), so that the multi-GPU layer -- halo exchange, migration, distributed list-growth search, slab
// re-planning -- runs on the host-emulated engine with one process per rank and no GPU
// (tests/test_dist_emulated_cpu.py; SPSPH_NCCL_SO points the engine at this library).
//
// Semantics kept from NCCL: point-to-point messages between two ranks are matched in issue order; collectives are
// issued by all ranks in the same order. Sends are buffered (a message is a file in a directory named by the unique
// id, moved into place atomically), receives poll for the file, so no grouping is needed to avoid deadlocks.
// Reductions combine the ranks' buffers in rank order on every rank: deterministic and identical everywhere.
#include <nccl.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <sys/stat.h>
#include <unistd.h>

struct ncclComm {
  int rank, nranks;
  std::string dir;
  std::vector<long long> sent, received;  // per peer: messages issued so far
};

namespace {
size_t dtype_size(ncclDataType_t t) {
  switch (t) {
    case ncclInt8: case ncclUint8: return 1;
    case ncclFloat16: return 2;
    case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
    case ncclInt64: case ncclUint64: case ncclFloat64: return 8;
    default: return 0;
  }
}
std::string msg_path(const ncclComm *c, int src, int dst, long long seq) {
  return c->dir + "/m_" + std::to_string(src) + "_" + std::to_string(dst) + "_" + std::to_string(seq);
}
ncclResult_t put(ncclComm *c, int peer, const void *buf, size_t bytes) {
  const std::string path = msg_path(c, c->rank, peer, c->sent[peer]++);
  const std::string tmp = path + ".tmp";
  FILE *f = std::fopen(tmp.c_str(), "wb");
  if (!f) return ncclSystemError;
  const size_t w = bytes ? std::fwrite(buf, 1, bytes, f) : 0;
  std::fclose(f);
  if (w != bytes || std::rename(tmp.c_str(), path.c_str()) != 0) return ncclSystemError;
  return ncclSuccess;
}
ncclResult_t get(ncclComm *c, int peer, void *buf, size_t bytes) {
  const std::string path = msg_path(c, peer, c->rank, c->received[peer]++);
  const auto t0 = std::chrono::steady_clock::now();
  struct stat sb;
  int spins = 0;
  while (stat(path.c_str(), &sb) != 0) {
    if (++spins > 200) std::this_thread::sleep_for(std::chrono::microseconds(200));
    if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(300)) return ncclSystemError;  // peer died
  }
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return ncclSystemError;
  const size_t want = bytes < (size_t)sb.st_size ? bytes : (size_t)sb.st_size;
  const size_t r = want ? std::fread(buf, 1, want, f) : 0;
  std::fclose(f);
  unlink(path.c_str());
  return r == want ? ncclSuccess : ncclSystemError;
}
template <class T>
void combine(T *acc, const T *in, size_t n, ncclRedOp_t op) {
  for (size_t i = 0; i < n; ++i) {
    if (op == ncclSum) acc[i] = acc[i] + in[i];
    else if (op == ncclMax) acc[i] = in[i] > acc[i] ? in[i] : acc[i];
    else if (op == ncclMin) acc[i] = in[i] < acc[i] ? in[i] : acc[i];
    else if (op == ncclProd) acc[i] = acc[i] * in[i];
  }
}
}  // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId *id) {
  std::memset(id->internal, 0, sizeof(id->internal));
  const auto now = std::chrono::high_resolution_clock::now().time_since_epoch().count();
  std::snprintf(id->internal, sizeof(id->internal), "spsph_fake_nccl_%d_%llx", (int)getpid(), (unsigned long long)now);
  return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
  id.internal[sizeof(id.internal) - 1] = 0;
  const char *base = std::getenv("SPSPH_FAKE_NCCL_DIR");
  ncclComm *c = new ncclComm;
  c->rank = rank;
  c->nranks = nranks;
  c->dir = std::string(base && *base ? base : "/tmp") + "/" + id.internal;
  c->sent.assign(nranks, 0);
  c->received.assign(nranks, 0);
  mkdir(c->dir.c_str(), 0700);  // every rank tries; EEXIST is fine
  struct stat sb;
  if (stat(c->dir.c_str(), &sb) != 0) {
    delete c;
    return ncclSystemError;
  }
  *comm = c;
  return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) {
  delete comm;
  return ncclSuccess;
}

ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  if (peer < 0 || peer >= c->nranks || peer == c->rank) return ncclInvalidArgument;
  return put(c, peer, buf, count * dtype_size(t));
}

ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, cudaStream_t) {
  if (peer < 0 || peer >= c->nranks || peer == c->rank) return ncclInvalidArgument;
  return get(c, peer, buf, count * dtype_size(t));
}

ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t c, cudaStream_t) {
  const size_t bytes = count * dtype_size(t);
  std::vector<char> mine((const char *)send, (const char *)send + bytes);  // send may alias recv
  for (int r = 0; r < c->nranks; ++r)
    if (r != c->rank && put(c, r, mine.data(), bytes) != ncclSuccess) return ncclSystemError;
  for (int r = 0; r < c->nranks; ++r) {
    char *dst = (char *)recv + (size_t)r * bytes;
    if (r == c->rank) std::memcpy(dst, mine.data(), bytes);
    else if (get(c, r, dst, bytes) != ncclSuccess) return ncclSystemError;
  }
  return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t c,
                           cudaStream_t s) {
  const size_t bytes = count * dtype_size(t);
  std::vector<char> all((size_t)c->nranks * bytes);
  const ncclResult_t rc = ncclAllGather(send, all.data(), count, t, c, s);
  if (rc != ncclSuccess) return rc;
  std::memcpy(recv, all.data(), bytes);
  for (int r = 1; r < c->nranks; ++r) {
    const char *in = all.data() + (size_t)r * bytes;
    switch (t) {
      case ncclFloat64: combine((double *)recv, (const double *)in, count, op); break;
      case ncclFloat32: combine((float *)recv, (const float *)in, count, op); break;
      case ncclInt64: combine((long long *)recv, (const long long *)in, count, op); break;
      case ncclUint64: combine((unsigned long long *)recv, (const unsigned long long *)in, count, op); break;
      case ncclInt32: combine((int *)recv, (const int *)in, count, op); break;
      case ncclUint32: combine((unsigned *)recv, (const unsigned *)in, count, op); break;
      default: return ncclInvalidArgument;
    }
  }
  return ncclSuccess;
}

ncclResult_t ncclGroupStart() { return ncclSuccess; }
ncclResult_t ncclGroupEnd() { return ncclSuccess; }
const char *ncclGetErrorString(ncclResult_t r) {
  return r == ncclSuccess ? "no error" : r == ncclInvalidArgument ? "fake nccl: invalid argument"
                                                                  : "fake nccl: file transport failed or a peer timed out";
}

}  // extern "C"
