// TEST INFRASTRUCTURE ONLY -- a stand-in for libnccl.so.2 for the host simulation (fake_cudart.cpp): the five entry points
// libbm resolves with dlsym (bm_comm.cu), implemented across PROCESSES through POSIX shared memory.  "Device" buffers are host
// memory there, so an all-reduce is: copy my buffer to my slot, barrier, combine the slots in rank order, barrier.
// Selected with BM_NCCL_LIB=<path>; lets tools/dist_check.py run the data-parallel paths of the engines -- the sum-allreduce
// of the step's statistics, the max-allreduce of the mean-field test, the sharded AIS ladder -- on the CPU.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

namespace {
constexpr size_t SLOT_BYTES = 32u << 20;          // per rank
struct Header { std::atomic<int> count, gen, attached; };
struct Comm { Header* h; unsigned char* slots; int rank, nranks; size_t bytes; char name[64]; };
struct Id { char bytes[128]; };

void barrier(Comm* c) {
    const int gen = c->h->gen.load();
    if (c->h->count.fetch_add(1) + 1 == c->nranks) { c->h->count.store(0); c->h->gen.fetch_add(1); }
    else while (c->h->gen.load() == gen) sched_yield();
}
size_t dtype_size(int dt) { return dt == 8 || dt == 4 || dt == 5 ? 8 : (dt == 6 || dt == 9 ? 2 : (dt <= 1 ? 1 : 4)); }

template <typename T> void combine(Comm* c, void* recv, size_t count, int op) {
    T* out = static_cast<T*>(recv);
    for (size_t i = 0; i < count; ++i) {
        T acc = reinterpret_cast<const T*>(c->slots)[i];
        for (int r = 1; r < c->nranks; ++r) {
            const T v = reinterpret_cast<const T*>(c->slots + (size_t)r * SLOT_BYTES)[i];
            if (op == 0) acc = acc + v; else if (op == 2) acc = v > acc ? v : acc; else if (op == 3) acc = v < acc ? v : acc; else acc = acc * v;
        }
        out[i] = acc;
    }
}
}  // namespace

extern "C" {

int ncclGetUniqueId(void* id) {
    memset(id, 0, 128);
    struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts);
    snprintf(static_cast<char*>(id), 64, "/bm_fakenccl_%d_%ld", (int)getpid(), (long)(ts.tv_nsec % 1000000007L));
    return 0;
}
int ncclCommInitRank(void** comm, int nranks, Id id, int rank) {
    Comm* c = new Comm();
    c->rank = rank; c->nranks = nranks; c->bytes = sizeof(Header) + 4096 + (size_t)nranks * SLOT_BYTES;
    strncpy(c->name, id.bytes, sizeof(c->name) - 1);
    const int fd = shm_open(c->name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) { perror("fake nccl: shm_open"); return 1; }
    if (ftruncate(fd, (off_t)c->bytes) != 0) { perror("fake nccl: ftruncate"); return 1; }     // zero-filled on creation
    void* p = mmap(nullptr, c->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { perror("fake nccl: mmap"); return 1; }
    c->h = static_cast<Header*>(p);
    c->slots = static_cast<unsigned char*>(p) + 4096;
    c->h->attached.fetch_add(1);
    while (c->h->attached.load() < nranks) sched_yield();        // everybody is mapped before the first collective
    *comm = c;
    return 0;
}
int ncclAllReduce(const void* send, void* recv, size_t count, int dtype, int op, void* comm, void* /*stream*/) {
    Comm* c = static_cast<Comm*>(comm);
    const size_t bytes = count * dtype_size(dtype);
    if (bytes > SLOT_BYTES) { fprintf(stderr, "fake nccl: %zu bytes exceed the slot size\n", bytes); return 2; }
    memcpy(c->slots + (size_t)c->rank * SLOT_BYTES, send, bytes);
    barrier(c);
    if (dtype == 7) combine<float>(c, recv, count, op);
    else if (dtype == 8) combine<double>(c, recv, count, op);
    else if (dtype == 3) combine<uint32_t>(c, recv, count, op);
    else if (dtype == 2) combine<int32_t>(c, recv, count, op);
    else { fprintf(stderr, "fake nccl: data type %d not implemented\n", dtype); return 3; }
    barrier(c);                                                   // the slots may be overwritten again
    return 0;
}
int ncclCommDestroy(void* comm) {
    Comm* c = static_cast<Comm*>(comm);
    if (!c) return 0;
    const int rank = c->rank;
    char name[64]; strncpy(name, c->name, sizeof(name));
    munmap(c->h, c->bytes);
    if (rank == 0) shm_unlink(name);
    delete c;
    return 0;
}
const char* ncclGetErrorString(int rc) { return rc == 0 ? "no error" : "fake nccl error"; }

}  // extern "C"
