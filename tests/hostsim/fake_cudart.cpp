// TEST INFRASTRUCTURE ONLY -- a stand-in for the CUDA runtime that lets the HOST side of libbm.so run without a GPU.
//
// tests/hostsim/build.sh links the library's own object files (the ones build.sh compiles for sm_100a) against this file
// instead of libcudart: device memory becomes host memory, kernel launches are recorded and skipped, and everything the host
// code asks the runtime or the driver to check is checked here --
//   * every cudaMemcpy* / cudaMemset* range must lie inside one live allocation when it touches "device" memory;
//   * cuTensorMapEncodeTiled (reached through cudaGetDriverEntryPoint, as bm_tc.cu does) validates its arguments the way the
//     driver does (16-byte aligned base, strides multiples of 16 bytes, box dimensions 1..256, 128-byte swizzle span) AND that
//     the whole tensor view (rows, leading dimension) lies inside one live allocation -- a wrong TcMat is a failure here,
//     not a corrupted tile on the GPU;
//   * frees of unknown pointers and leaks at exit are reported.
// With fakecuda_set_execute(1) the recorded launches are interpreted on the CPU (kernels_cpu.cpp), which also checks the dataflow of
// every program launch for hazards no declared dependency covers (fakecuda_set_hazards: 0 off, 1 launches small enough to shadow
// -- the default --, 2 always; fakecuda_drop_dependency pretends one declaration away, for tests of the checker).
// With fakecuda_set_stream_order_check(1) (needs execute) the ORDER between streams is checked too: every copy, memset and
// interpreted launch runs with all device allocations protected, so that the pages it reads and writes are known exactly (page
// faults); each operation carries the vector clock its stream, the events it waited for and the host's synchronisations give it, and
// two operations that touch the same page, at least one writing, without one happening before the other are a violation ("stream
// race") -- the double-buffered epoch loops (copy stream against compute stream) are the customers.  fakecuda_ignore_event_waits(1)
// makes cudaStreamWaitEvent a no-op, for the tests of the checker itself.
// What the run proves: argument validation (every BM_REQUIRE), buffer sizing, pointer arithmetic, program construction and
// the control flow around the kernels.  What it cannot prove: anything a kernel computes (results are whatever the zeroed
// buffers hold).  It never ships: the product library links the real runtime and refuses to start without a B200.
#include <cuda_runtime_api.h>
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>
#include <array>
#include <signal.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>

namespace fakecuda {
bool execute(const std::string& name, dim3 grid, dim3 block, void** args);
void report_violation(const std::string& m);        // for kernels_cpu.cpp (dataflow hazards of a program launch)
int hazard_mode();                                   // 0 off, 1 on for launches small enough to shadow, 2 always
}

namespace {

// Construct-on-first-use: the nvcc-generated registration constructors of the library's objects call into this file
// while the shared object is still being initialised, possibly before this translation unit's own globals exist.
struct State {
    std::mutex mu;
    std::map<uintptr_t, size_t> alloc;               // live "device" allocations
    std::map<const void*, std::string> kernels;      // host stub -> device (mangled) name
    std::map<std::string, long> launches;            // per kernel name
    std::string violation;                           // first violation (sticky)
    long encodes = 0;
    long syncs = 0;                                  // host-blocking calls: stream / event synchronisation, synchronous copies
    long h2d_bytes = 0, d2h_bytes = 0;               // by the copy kind the caller states
    bool execute = false;                            // interpret launches on the CPU (kernels_cpu.cpp) instead of skipping them
    std::map<std::string, long> skipped;             // launches without a CPU restatement while `execute` was on
    int hazards = 1;                                 // check the dataflow of interpreted program launches (kernels_cpu.cpp)
    long hazard_launches = 0;                        // program launches that were checked
    long unhonoured = 0;                             // declared dependencies the kernel's wait loop skips (kernels_cpu.cpp, Hazards)
    int drop_op = -1, drop_dep = -1;                 // tests of the checker itself: pretend this dependency was not declared
};
State& st() { static State* s = new State(); return *s; }
#define g_mu (st().mu)
#define g_alloc (st().alloc)
#define g_kernels (st().kernels)
#define g_launches (st().launches)
#define g_violation (st().violation)
#define g_encodes (st().encodes)
struct CallCfg { dim3 grid, block; size_t smem; cudaStream_t stream; };
std::vector<CallCfg>& cfg_stack() { thread_local std::vector<CallCfg>* v = new std::vector<CallCfg>(); return *v; }
#define t_cfg (cfg_stack())

void violation(const std::string& m) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_violation.empty()) g_violation = m;
    fprintf(stderr, "[fake_cudart] VIOLATION: %s\n", m.c_str());
}
}  // namespace
namespace fakecuda {
void report_violation(const std::string& m) { violation(m); }
int hazard_mode() { return st().hazards; }
void count_hazard_launch() { st().hazard_launches++; }
void count_unhonoured(long n) { st().unhonoured += n; }
bool dropped_dependency(int op, int d) { return st().drop_op == op && st().drop_dep == d; }
}
namespace {
// the live allocation containing [p, p + n), or 0
bool inside_one_allocation(const void* p, size_t n) {
    std::lock_guard<std::mutex> lk(g_mu);
    const uintptr_t a = (uintptr_t)p;
    auto it = g_alloc.upper_bound(a);
    if (it == g_alloc.begin()) return false;
    --it;
    return a >= it->first && a + n <= it->first + it->second;
}
bool touches_device(const void* p) {
    std::lock_guard<std::mutex> lk(g_mu);
    const uintptr_t a = (uintptr_t)p;
    auto it = g_alloc.upper_bound(a);
    if (it == g_alloc.begin()) return false;
    --it;
    return a >= it->first && a < it->first + it->second;
}
void check_range(const void* p, size_t n, const char* what) {
    if (n == 0 || !touches_device(p)) return;          // host buffers of the caller are not tracked
    if (!inside_one_allocation(p, n)) {
        char buf[200];
        snprintf(buf, sizeof(buf), "%s: %zu bytes at %p run past the end of their device allocation", what, n, p);
        violation(buf);
    }
}

CUresult fake_encode_tiled(CUtensorMap* tm, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* gdim,
                           const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                           CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    ++g_encodes;
    char buf[256];
    auto bad = [&](const char* why) { snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled: %s", why); violation(buf); return CUDA_ERROR_INVALID_VALUE; };
    if (!tm || !base || !gdim || !box || !estr) return bad("null argument");
    if (rank < 1 || rank > 5) return bad("rank out of range");
    if (dt != CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) return bad("this library only maps bf16 tensors");
    const size_t esz = 2;
    if (((uintptr_t)base & 15) != 0) return bad("global address not 16-byte aligned");
    for (cuuint32_t i = 0; i < rank; ++i) {
        if (gdim[i] == 0 || gdim[i] > (1ull << 32)) return bad("global dimension out of range");
        if (box[i] == 0 || box[i] > 256) return bad("box dimension must be 1..256");
        if (estr[i] == 0 || estr[i] > 8) return bad("element stride must be 1..8");
    }
    for (cuuint32_t i = 0; i + 1 < rank; ++i) {
        if (gstride[i] % 16 != 0 || gstride[i] >= (1ull << 40)) return bad("global stride must be a multiple of 16 bytes below 2^40");
        if (i == 0 && gstride[0] < gdim[0] * esz) return bad("row stride smaller than a row");
    }
    if (il == CU_TENSOR_MAP_INTERLEAVE_NONE && sw == CU_TENSOR_MAP_SWIZZLE_128B && box[0] * esz > 128) return bad("inner box dimension exceeds the 128-byte swizzle span");
    if (il == CU_TENSOR_MAP_INTERLEAVE_NONE && (box[0] * esz) % 16 != 0) return bad("inner box dimension must be a multiple of 16 bytes");
    // the whole view must be backed by one allocation: last row start + one row
    size_t span = gdim[0] * esz;
    for (cuuint32_t i = 1; i < rank; ++i) span += (size_t)(gdim[i] - 1) * gstride[i - 1];
    if (!inside_one_allocation(base, span)) {
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled: tensor view of %zu bytes at %p (dims %llu x %llu, row stride %llu B) is not inside one device allocation",
                 span, base, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 1), (unsigned long long)(rank > 1 ? gstride[0] : 0));
        violation(buf);
        return CUDA_ERROR_INVALID_VALUE;
    }
    struct MapView { const void* base; unsigned long long cols, rows, ld_bytes; unsigned box0, box1; unsigned magic; } mv;   // read by kernels_cpu.cpp
    memset(&mv, 0, sizeof(mv));
    mv.base = base; mv.cols = gdim[0]; mv.rows = rank > 1 ? gdim[1] : 1; mv.ld_bytes = rank > 1 ? gstride[0] : gdim[0] * esz;
    mv.box0 = box[0]; mv.box1 = rank > 1 ? box[1] : 1; mv.magic = 0xB200C0DEu;
    memset(tm, 0, sizeof(*tm));
    memcpy(tm, &mv, sizeof(mv));
    return CUDA_SUCCESS;
}

std::string record_launch(const void* func) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_kernels.find(func);
    const std::string name = it == g_kernels.end() ? std::string("<unregistered>") : it->second;
    g_launches[name]++;
    return name;
}

}  // namespace

namespace {

// ---- stream-order check ---------------------------------------------------------------------------------------------------
constexpr int MAX_STREAMS = 16;
typedef std::array<unsigned, MAX_STREAMS> Clock;
constexpr size_t PAGE = 4096;
struct Access { uintptr_t lo, hi; bool write; int stream; Clock clock; std::string label; };   // pages [lo, hi)
struct Order {
    bool on = false, ignore_waits = false;
    std::map<cudaStream_t, int> ids;                  // 0: the legacy default stream
    std::map<int, bool> blocking;
    Clock host{};                                     // what the host has synchronised with
    Clock at[MAX_STREAMS]{};                          // per stream: everything that happens before its next operation
    std::map<cudaEvent_t, Clock> events;
    std::map<uintptr_t, std::vector<Access>> history; // per allocation
    long ops = 0, races = 0;
};
Order& ord() { static Order* o = new Order(); return *o; }
Clock merged(Clock a, const Clock& b) { for (int i = 0; i < MAX_STREAMS; ++i) a[i] = std::max(a[i], b[i]); return a; }
int stream_id(cudaStream_t s) {
    Order& o = ord();
    if (!s) return 0;
    auto it = o.ids.find(s);
    if (it != o.ids.end()) return it->second;
    const int id = (int)o.ids.size() + 1;
    o.ids[s] = id < MAX_STREAMS ? id : MAX_STREAMS - 1;
    return o.ids[s];
}
// page-fault recording: async-signal-safe, fixed storage
struct Region { uintptr_t base, size; };
Region g_regions[4096]; int g_n_regions = 0;
struct Fault { uintptr_t page; bool write; };
Fault* g_faults = nullptr; size_t g_n_faults = 0; constexpr size_t MAX_FAULTS = 1u << 22;
volatile bool g_recording = false;
struct sigaction g_prev_segv;
void segv_handler(int sig, siginfo_t* info, void* uc) {
    const uintptr_t a = (uintptr_t)info->si_addr;
    if (g_recording) {
        for (int i = 0; i < g_n_regions; ++i)
            if (a >= g_regions[i].base && a < g_regions[i].base + g_regions[i].size) {
                const bool wr = (((ucontext_t*)uc)->uc_mcontext.gregs[REG_ERR] & 2) != 0;
                const uintptr_t page = a & ~(uintptr_t)(PAGE - 1);
                if (g_n_faults < MAX_FAULTS) { g_faults[g_n_faults].page = page; g_faults[g_n_faults].write = wr; ++g_n_faults; }
                mprotect((void*)page, PAGE, wr ? (PROT_READ | PROT_WRITE) : PROT_READ);
                return;
            }
    }
    // not ours: hand over to whoever was installed before (Python's faulthandler, or the default action)
    if (g_prev_segv.sa_flags & SA_SIGINFO) { if (g_prev_segv.sa_sigaction) { g_prev_segv.sa_sigaction(sig, info, uc); return; } }
    else if (g_prev_segv.sa_handler != SIG_DFL && g_prev_segv.sa_handler != SIG_IGN) { g_prev_segv.sa_handler(sig); return; }
    signal(SIGSEGV, SIG_DFL);
}
void order_enable(bool on) {
    Order& o = ord();
    if (on && !g_faults) {
        g_faults = (Fault*)mmap(nullptr, MAX_FAULTS * sizeof(Fault), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        struct sigaction sa;
        memset(&sa, 0, sizeof(sa));
        sa.sa_sigaction = segv_handler; sa.sa_flags = SA_SIGINFO | SA_NODEFER;
        sigemptyset(&sa.sa_mask);
        sigaction(SIGSEGV, &sa, &g_prev_segv);
    }
    o.on = on;
    o.history.clear();
}
struct OpScope {                                       // one copy / memset / interpreted launch
    bool active = false; int sid = 0; Clock clock{}; std::string label;
    OpScope(cudaStream_t s, const std::string& what, bool host_synchronous = false) {
        Order& o = ord();
        if (!o.on) return;
        active = true; label = what; sid = stream_id(s);
        Clock c = merged(o.at[sid], o.host);
        if (host_synchronous || sid == 0)               // legacy default stream: after everything in the blocking streams
            for (auto& kv : o.blocking) if (kv.second) c = merged(c, o.at[kv.first]);
        c[sid] += 1;
        clock = c; o.at[sid] = c; ++o.ops;
        if (host_synchronous || sid == 0) {
            for (auto& kv : o.blocking) if (kv.second) o.at[kv.first] = merged(o.at[kv.first], c);
            if (host_synchronous) o.host = merged(o.host, c);
        }
        std::lock_guard<std::mutex> lk(g_mu);
        g_n_regions = 0;
        for (auto& kv : g_alloc) if (g_n_regions < 4096) { g_regions[g_n_regions].base = kv.first; g_regions[g_n_regions].size = (kv.second + PAGE - 1) & ~(PAGE - 1); ++g_n_regions; }
        for (int i = 0; i < g_n_regions; ++i) mprotect((void*)g_regions[i].base, g_regions[i].size, PROT_NONE);
        g_n_faults = 0;
        g_recording = true;
    }
    ~OpScope() {
        if (!active) return;
        g_recording = false;
        for (int i = 0; i < g_n_regions; ++i) mprotect((void*)g_regions[i].base, g_regions[i].size, PROT_READ | PROT_WRITE);
        Order& o = ord();
        // pages -> per allocation, reads and writes separately, consecutive pages coalesced
        std::vector<Fault> f(g_faults, g_faults + g_n_faults);
        std::sort(f.begin(), f.end(), [](const Fault& a, const Fault& b) { return a.write != b.write ? a.write < b.write : a.page < b.page; });
        std::vector<Access> mine;
        for (size_t i = 0; i < f.size();) {
            size_t j = i + 1;
            while (j < f.size() && f[j].write == f[i].write && f[j].page <= f[j - 1].page + PAGE) ++j;
            // (a coalesced run stays inside one allocation only if the allocations are not adjacent: split at region borders)
            uintptr_t lo = f[i].page;
            const uintptr_t end = f[j - 1].page + PAGE;
            while (lo < end) {
                uintptr_t base = 0, lim = 0;
                for (int r = 0; r < g_n_regions; ++r) if (lo >= g_regions[r].base && lo < g_regions[r].base + g_regions[r].size) { base = g_regions[r].base; lim = base + g_regions[r].size; }
                if (!base) break;
                const uintptr_t hi = std::min(end, lim);
                Access a; a.lo = lo; a.hi = hi; a.write = f[i].write; a.stream = sid; a.clock = clock; a.label = label;
                mine.push_back(a);
                (void)base;
                lo = hi;
            }
            i = j;
        }
        std::string report;
        for (const Access& a : mine) {
            uintptr_t base = 0;
            for (int r = 0; r < g_n_regions; ++r) if (a.lo >= g_regions[r].base && a.lo < g_regions[r].base + g_regions[r].size) base = g_regions[r].base;
            std::vector<Access>& h = o.history[base];
            for (const Access& p : h) {
                if (!(a.write || p.write) || p.hi <= a.lo || a.hi <= p.lo) continue;
                if (p.clock[p.stream] <= a.clock[p.stream]) continue;          // p happens before a
                if (report.empty()) {
                    char buf[512];
                    snprintf(buf, sizeof(buf), "stream race: '%s' (stream %d, %s) and the earlier '%s' (stream %d, %s) touch bytes [%zu, %zu) of the allocation at %p with no ordering between them",
                             a.label.c_str(), a.stream, a.write ? "writes" : "reads", p.label.c_str(), p.stream, p.write ? "writes" : "reads",
                             (size_t)(std::max(a.lo, p.lo) - base), (size_t)(std::min(a.hi, p.hi) - base), (void*)base);
                    report = buf;
                }
                ++o.races;
            }
        }
        for (const Access& a : mine) {
            uintptr_t base = 0;
            for (int r = 0; r < g_n_regions; ++r) if (a.lo >= g_regions[r].base && a.lo < g_regions[r].base + g_regions[r].size) base = g_regions[r].base;
            o.history[base].push_back(a);
        }
        if (!report.empty()) violation(report);
    }
};
void order_host_sync(const Clock& with) {               // the host has waited for `with`: older accesses are ordered before everything new
    Order& o = ord();
    o.host = merged(o.host, with);
    if (!o.on) return;
    for (auto& kv : o.history) {
        std::vector<Access>& h = kv.second;
        h.erase(std::remove_if(h.begin(), h.end(), [&](const Access& p) { return p.clock[p.stream] <= o.host[p.stream]; }), h.end());
    }
}

}  // namespace

extern "C" {

// ---- inspection hooks for the tests --------------------------------------------------------------------------------
const char* fakecuda_violation(void) { std::lock_guard<std::mutex> lk(g_mu); static std::string s; s = g_violation; return s.c_str(); }
void fakecuda_reset(void) { std::lock_guard<std::mutex> lk(g_mu); g_violation.clear(); g_launches.clear(); g_encodes = 0; st().skipped.clear(); st().syncs = 0; st().h2d_bytes = 0; st().d2h_bytes = 0; }
long fakecuda_launches(const char* substr) {
    std::lock_guard<std::mutex> lk(g_mu);
    long n = 0;
    for (auto& kv : g_launches) if (!substr || !*substr || kv.first.find(substr) != std::string::npos) n += kv.second;
    return n;
}
long fakecuda_tensor_maps(void) { return g_encodes; }
long fakecuda_syncs(void) { return st().syncs; }
long fakecuda_h2d_bytes(void) { return st().h2d_bytes; }
long fakecuda_d2h_bytes(void) { return st().d2h_bytes; }
void fakecuda_set_execute(int on) { st().execute = on != 0; }
void fakecuda_set_hazards(int mode) { st().hazards = mode; }
long fakecuda_hazard_launches(void) { return st().hazard_launches; }
long fakecuda_unhonoured_dependencies(void) { return st().unhonoured; }
void fakecuda_drop_dependency(int op, int d) { st().drop_op = op; st().drop_dep = d; }
void fakecuda_set_stream_order_check(int on) { order_enable(on != 0); }
void fakecuda_ignore_event_waits(int on) { ord().ignore_waits = on != 0; }
long fakecuda_stream_order_ops(void) { return ord().ops; }
long fakecuda_stream_races(void) { return ord().races; }
// kernels that were launched while executing but have no CPU restatement: "name xN; ..." ("" if none)
const char* fakecuda_skipped(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    static std::string s; s.clear();
    for (auto& kv : st().skipped) s += kv.first + " x" + std::to_string(kv.second) + "; ";
    return s.c_str();
}
long fakecuda_live_allocations(void) { std::lock_guard<std::mutex> lk(g_mu); return (long)g_alloc.size(); }

// ---- registration (called by the nvcc-generated host stubs at load time) ------------------------------------------------
void** __cudaRegisterFatBinary(void*) { static void* handle = nullptr; return &handle; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char* hostFun, char*, const char* deviceName, int, uint3*, uint3*, dim3*, dim3*, int*) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_kernels[(const void*)hostFun] = deviceName ? deviceName : "?";
}
void __cudaRegisterVar(void**, char*, char*, const char*, int, size_t, int, int) {}
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, struct CUstream_st* stream) {
    t_cfg.push_back(CallCfg{grid, block, smem, stream});
    return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* smem, void* stream) {
    if (t_cfg.empty()) return cudaErrorInvalidConfiguration;
    const CallCfg c = t_cfg.back(); t_cfg.pop_back();
    *grid = c.grid; *block = c.block; *smem = c.smem; *(cudaStream_t*)stream = c.stream;
    return cudaSuccess;
}

// ---- devices ----------------------------------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) { *n = 8; return cudaSuccess; }       // one host memory behind all of them
cudaError_t cudaSetDevice(int d) { return d >= 0 && d < 8 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "fake B200 (host simulation)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    p->sharedMemPerBlockOptin = 232448; p->totalGlobalMem = (size_t)180 << 30;
    return cudaSuccess;
}
static thread_local cudaError_t t_last_error = cudaSuccess;
cudaError_t cudaGetLastError(void) { const cudaError_t e = t_last_error; t_last_error = cudaSuccess; return e; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake_cudart error"; }
cudaError_t cudaFuncSetAttribute(const void*, enum cudaFuncAttribute, int) { return cudaSuccess; }
static int g_max_active_clusters = -1;        // -1: what 148 SMs with one CTA each hold
void fakecuda_set_max_active_clusters(int n) { g_max_active_clusters = n; }
cudaError_t cudaOccupancyMaxActiveClusters(int* n, const void*, const cudaLaunchConfig_t* c) {
    int cluster = 1;
    for (unsigned i = 0; i < c->numAttrs; ++i) if (c->attrs[i].id == cudaLaunchAttributeClusterDimension) cluster = (int)c->attrs[i].val.clusterDim.x;
    *n = g_max_active_clusters >= 0 ? g_max_active_clusters : 148 / cluster;
    return cudaSuccess;
}
cudaError_t cudaGetDriverEntryPoint(const char* symbol, void** fn, unsigned long long, enum cudaDriverEntryPointQueryResult* st) {
    if (symbol && !strcmp(symbol, "cuTensorMapEncodeTiled")) { *fn = (void*)&fake_encode_tiled; if (st) *st = cudaDriverEntryPointSuccess; return cudaSuccess; }
    *fn = nullptr; if (st) *st = cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}

// ---- memory -------------------------------------------------------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t n) {
    void* q = nullptr;
    if (n == 0) n = 1;
    // whole pages of its own (the stream-order check protects allocations page by page)
    if (posix_memalign(&q, PAGE, (n + PAGE - 1) & ~(PAGE - 1)) != 0) return cudaErrorMemoryAllocation;
    memset(q, 0xCD, n);                                  // device memory is NOT zero-initialised
    std::lock_guard<std::mutex> lk(g_mu);
    g_alloc[(uintptr_t)q] = n;
    *p = q;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_alloc.find((uintptr_t)p);
        if (it == g_alloc.end()) { if (g_violation.empty()) g_violation = "cudaFree of a pointer that is not a live allocation"; return cudaErrorInvalidValue; }
        g_alloc.erase(it);
    }
    ord().history.erase((uintptr_t)p);                   // (cudaFree waits for the device)
    free(p);
    return cudaSuccess;
}
// no inter-process memory on the stand-in: the engine keeps its ncclAllReduce path (bm_peer.cu)
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned int) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static void copy_checked(void* d, const void* s, size_t n, enum cudaMemcpyKind k) {
    check_range(d, n, "cudaMemcpy destination"); check_range(s, n, "cudaMemcpy source");
    if (k == cudaMemcpyHostToDevice) st().h2d_bytes += (long)n; else if (k == cudaMemcpyDeviceToHost) st().d2h_bytes += (long)n;
    memmove(d, s, n);
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, enum cudaMemcpyKind k) {
    st().syncs++;
    OpScope op(nullptr, "cudaMemcpy", true);
    copy_checked(d, s, n, k);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, enum cudaMemcpyKind k, cudaStream_t stream) {
    OpScope op(stream, "cudaMemcpyAsync");
    copy_checked(d, s, n, k);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, enum cudaMemcpyKind, cudaStream_t stream) {
    if (h) { check_range(d, (h - 1) * dp + w, "cudaMemcpy2D destination"); check_range(s, (h - 1) * sp + w, "cudaMemcpy2D source"); }
    OpScope op(stream, "cudaMemcpy2DAsync");
    for (size_t r = 0; r < h; ++r) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t stream) {
    check_range(d, n, "cudaMemset");
    OpScope op(stream, "cudaMemsetAsync");
    memset(d, v, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyToSymbolAsync(const void* sym, const void* s, size_t n, size_t off, enum cudaMemcpyKind, cudaStream_t) {
    memcpy((char*)sym + off, s, n);                       // the host shadow of the __constant__ array
    return cudaSuccess;
}

// ---- streams / events -------------------------------------------------------------------------------------------------------
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags) {
    *s = (cudaStream_t)malloc(8);
    ord().blocking[stream_id(*s)] = (flags & cudaStreamNonBlocking) == 0;
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) { ord().blocking.erase(stream_id(s)); ord().ids.erase(s); free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { st().syncs++; order_host_sync(ord().at[stream_id(s)]); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    Order& o = ord();
    auto it = o.events.find(e);
    if (!o.ignore_waits && it != o.events.end()) { const int id = stream_id(s); o.at[id] = merged(o.at[id], it->second); }
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { ord().events.erase(e); free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { Order& o = ord(); o.events[e] = merged(o.at[stream_id(s)], o.host); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) {
    st().syncs++;
    auto it = ord().events.find(e);
    if (it != ord().events.end()) order_host_sync(it->second);
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1e-3f; return cudaSuccess; }

// ---- launches: recorded, not executed --------------------------------------------------------------------------------------
cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t stream) {
    if (grid.x == 0 || grid.y == 0 || grid.z == 0 || block.x * block.y * block.z == 0 || block.x * block.y * block.z > 1024 ||
        grid.y > 65535 || grid.z > 65535 || smem > 232448) {
        char buf[160];
        snprintf(buf, sizeof(buf), "invalid launch configuration: grid (%u,%u,%u) block (%u,%u,%u) smem %zu", grid.x, grid.y, grid.z, block.x, block.y, block.z, smem);
        violation(buf);
        t_last_error = cudaErrorInvalidConfiguration;
        return cudaErrorInvalidConfiguration;
    }
    const std::string name = record_launch(func);
    bool done = true;
    if (st().execute) { OpScope op(stream, name); done = fakecuda::execute(name, grid, block, args); }
    if (!done) { std::lock_guard<std::mutex> lk(g_mu); st().skipped[name]++; }
    return cudaSuccess;
}
cudaError_t cudaLaunchKernelExC(const cudaLaunchConfig_t* c, const void* func, void** args) {
    for (unsigned i = 0; i < c->numAttrs; ++i)
        if (c->attrs[i].id == cudaLaunchAttributeClusterDimension) {
            const unsigned cx = c->attrs[i].val.clusterDim.x, cy = c->attrs[i].val.clusterDim.y, cz = c->attrs[i].val.clusterDim.z;
            if (cx == 0 || cy == 0 || cz == 0 || cx * cy * cz > 8 || c->gridDim.x % cx || c->gridDim.y % cy || c->gridDim.z % cz) {
                violation("cluster launch: the grid is not a whole number of (portable-size) clusters");
                return cudaErrorInvalidConfiguration;
            }
        }
    return cudaLaunchKernel(func, c->gridDim, c->blockDim, args, c->dynamicSmemBytes, c->stream);
}

}  // extern "C"
