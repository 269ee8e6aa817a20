// Data-parallel CD step over NVLink peer memory (one process per GPU): the reduce-scatter of the step's statistics, the
// sharded momentum update and the all-gather of the new weights (fp32 master, momentum, bf16 shadow) as TWO kernels
// whose transfers are plain stores to the peers' mapped buffers -- no staging copy, no library collective.
//
//   rank r owns rows [r * rows_per, (r + 1) * rows_per) of W.
//   push   (every rank): sums its split-K slices and STORES shard s of the result into rank s's inbox[r]; the three small
//          vectors (sum(h0 - hk), -sum(hk), sum(X - vk)) go to every rank's `small[r]`.  Last block: release ready[r] on all peers.
//   update (every rank): waits for ready[*], adds the inbox slices in rank order (one writer per row: bit-identical
//          weights everywhere), applies base_rbm.py:445-474 to its rows and STORES the new W / dW / bf16 W rows into every
//          rank's copy; biases and sparsity statistics are updated redundantly (identically) by every rank from `small`.
//          Last block: release done[r] on all peers.
//   wait   : one warp spins until done[*] of this step has arrived -- after it, every copy on this GPU is complete.
//
// Buffers are exchanged as CUDA IPC handles through the context's communicator at construction (an all-gather written as
// a sum-allreduce of a zero-padded table: no further NCCL symbol needed).  If any rank cannot export or map (same process,
// no peer access, the host simulation), every rank keeps the ncclAllReduce path.
#pragma once
#include "bm_internal.h"
#include "bm_tc_desc.h"

namespace bm {

constexpr int MAX_PEERS = 16;

struct PeerView {                 // one rank's exported buffers, as mapped in THIS process (own rank: the local pointers)
    float* inbox;                 // [nranks][shard_elems]
    float* small;                 // [nranks][small_len]
    int* flags;                   // ready[nranks] | done[nranks]
    float* W; float* dW; __nv_bfloat16* Wb;
};

struct DpStep {
    int rank, nranks, V, H, srow, rows_per, ldwb, small_len, step, replicate_fp32;
    unsigned long long shard_elems;
    PeerView peer[MAX_PEERS];
    unsigned int* counter;        // local scratch: blocks that have finished (two words: push, update)
    // this rank's statistics (see CdTail)
    const float* part;  unsigned long long stride;  int splits;
    const float* vpart; unsigned long long vstride; int vsplits;
    // update scalars
    float n_div, lr, mom, l2, damp, cost, target;
    float *vb, *hb, *dvb, *dhb;
    const float* q_old; float* q_new; float* pen;
};

struct PeerExchange {
    Ctx* ctx = nullptr;
    bool active = false;
    int V = 0, H = 0, rows_per = 0, small_len = 0, ldwb = 0;
    size_t shard_elems = 0;
    void* arena = nullptr;                    // local: inbox | small | flags | counter
    PeerView view[MAX_PEERS] = {};
    void* mapped[MAX_PEERS][4] = {};               // peers' bases opened with cudaIpcOpenMemHandle (arena, W, dW, Wb)
    int step = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};      // BM_PEER_PROFILE
    double prof_ms[3] = {0, 0, 0}; long prof_n = 0; bool prof_pending = false;

    // collective over the context's communicator; W / dW / Wb are whole cudaMalloc allocations
    void setup(Ctx* c, int V_, int H_, float* W, float* dW, __nv_bfloat16* Wb, int ldwb_);
    void fill(DpStep& s) const;
    void run(DpStep& s);                      // push, update, wait on the context's stream
    // after a step the fp32 master / momentum rows of the other ranks' shards are stale on this rank (only the bf16 shadow is
    // all-gathered): whoever reads them (get_param, the free-energy metrics, the L2 loss) pulls them from their owners first
    bool stale = false;
    void pull_replicas();
    void release();
    ~PeerExchange() { release(); }
};

}  // namespace bm
