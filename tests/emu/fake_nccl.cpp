// fake_nccl.cpp — TEST INFRASTRUCTURE: the four NCCL entry points the library dlopen()s, implemented over a POSIX
// shared-memory segment so that several HOST-EMULATION processes (tests/test_library_emulation.py, one per "rank") can
// run the sharded solve: ncclAllReduce = every rank copies its buffer into its slot, barrier, every rank reduces the
// slots in rank order (deterministic), barrier.  Built as tests/emu/_build/libnccl.so.2 and found through
// LD_LIBRARY_PATH; never part of the product.  The 128-byte unique id is the name of the segment.
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {
constexpr size_t kSlot = 64u << 20;   // bytes per rank and collective
struct Header {
  std::atomic<int> arrived;
  std::atomic<int> generation;
  int world;
};
struct Comm {
  Header* h;
  char* slots;
  int rank, world;
  size_t bytes;
};
void barrier(Comm* c) {
  const int gen = c->h->generation.load();
  if (c->h->arrived.fetch_add(1) + 1 == c->world) {
    c->h->arrived.store(0);
    c->h->generation.fetch_add(1);
  } else {
    while (c->h->generation.load() == gen) sched_yield();
  }
}
}  // namespace

struct ncclUniqueIdBlob { char internal[128]; };

extern "C" {
int ncclGetUniqueId(void* out) {
  char name[128] = {0};
  snprintf(name, sizeof name, "/b200emu_%d_%ld", (int)getpid(), (long)random());
  memcpy(out, name, 128);
  return 0;
}
int ncclCommInitRank(void** comm, int world, ncclUniqueIdBlob id, int rank) {
  const size_t bytes = sizeof(Header) + 4096 + kSlot * (size_t)world;
  int fd = -1;
  if (rank == 0) {
    fd = shm_open(id.internal, O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, (off_t)bytes) != 0) return 1;
  } else {
    for (int tries = 0; tries < 20000 && fd < 0; tries++) {
      fd = shm_open(id.internal, O_RDWR, 0600);
      struct stat st;
      if (fd >= 0 && (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes)) { close(fd); fd = -1; }
      if (fd < 0) usleep(1000);
    }
    if (fd < 0) return 1;
  }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return 1;
  Comm* c = new Comm{(Header*)p, (char*)p + 4096, rank, world, bytes};
  if (rank == 0) c->h->world = world;   // a fresh segment is zero-filled: arrived = generation = 0
  *comm = c;
  barrier(c);
  if (rank == 0) shm_unlink(id.internal);   // everybody has it mapped
  return 0;
}
int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void* /*stream*/) {
  Comm* c = (Comm*)comm;
  const size_t el = dtype == 8 ? 8 : 4;   // ncclFloat64 = 8, ncclInt32 = 2
  for (size_t done = 0; done < count;) {    // chunks of one slot
    const size_t n = (count - done) * el > kSlot ? kSlot / el : count - done;
    memcpy(c->slots + kSlot * (size_t)c->rank, (const char*)send + done * el, n * el);
    barrier(c);
    if (dtype == 8) {
      double* out = (double*)recv + done;
      for (size_t i = 0; i < n; i++) {
        double acc = ((const double*)c->slots)[i];
        for (int r = 1; r < c->world; r++) {
          const double v = ((const double*)(c->slots + kSlot * (size_t)r))[i];
          acc = op == 0 ? acc + v : (op == 2 ? (v > acc ? v : acc) : (v < acc ? v : acc));
        }
        out[i] = acc;
      }
    } else {
      int32_t* out = (int32_t*)recv + done;
      for (size_t i = 0; i < n; i++) {
        int32_t acc = ((const int32_t*)c->slots)[i];
        for (int r = 1; r < c->world; r++) {
          const int32_t v = ((const int32_t*)(c->slots + kSlot * (size_t)r))[i];
          acc = op == 0 ? acc + v : (op == 2 ? (v > acc ? v : acc) : (v < acc ? v : acc));
        }
        out[i] = acc;
      }
    }
    barrier(c);
    done += n;
  }
  return 0;
}
// ncclReduce: the sum lands on `root` only (the other ranks' receive buffers stay as they are) — through the all-reduce
// path into a scratch buffer, then copied on the root.  Group calls are no-ops: every rank issues the same sequence.
int ncclReduce(const void* send, void* recv, size_t count, int dtype, int op, int root, void* comm, void* stream) {
  Comm* c = (Comm*)comm;
  const size_t el = dtype == 8 ? 8 : 4;
  char* tmp = (char*)malloc(count * el + 8);
  const int rc = ncclAllReduce(send, tmp, count, dtype, op, comm, stream);
  if (rc == 0 && c->rank == root) memcpy(recv, tmp, count * el);
  free(tmp);
  return rc;
}
int ncclGroupStart() { return 0; }
int ncclGroupEnd() { return 0; }
int ncclCommDestroy(void* comm) {
  Comm* c = (Comm*)comm;
  munmap((void*)c->h, c->bytes);
  delete c;
  return 0;
}
const char* ncclGetErrorString(int) { return "fake NCCL (host emulation)"; }
}
