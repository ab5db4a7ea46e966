// HRNet program representation shared by hrnet.cu (builder / executor / SIMT
// kernels) and conv_umma.cu (tcgen05 implicit-GEMM engine).
#pragma once
#include "hbp_internal.cuh"
#include <cuda.h>
#include <string>
#include <vector>

// One activation tensor, NHWC fp16: (P, h, w, c); batch dimension decided at run time.
struct HTensor {
    int c = 0, h = 0, w = 0;    // c: channels as stored (padded to the tensor engine's 64-channel rows where needed)
    int c_l = 0;                // logical channels (what the checkpoint has); channels c_l..c-1 are always zero
    int buf = -1;               // index into the per-batch buffer table
};

enum HOpKind { OP_STEM1 = 0, OP_CONV = 1, OP_HEAD = 2, OP_UPADD = 3,
               OP_GROUP = 4,        // one launch that runs `members` (independent OP_CONV ops of one fuse level)
               OP_UPADD_GROUP = 5,  // one launch that runs `members` (the OP_UPADD ops of one fuse stage)
               OP_CHAIN = 6         // one persistent launch that runs `members` IN ORDER: the 3x3 convs of one branch of a stage module
};

// out = act( conv_k,s(in) + bias [+ residual] ), optionally replicated `up` x `up`
// (nearest upsample fused into the store; residual is read at the upsampled position)
struct HOp {
    int kind = OP_CONV;
    std::string name;           // public HRNet state_dict prefix, e.g. "stage2.0.branches.1.0.conv1"
    int in = -1, out = -1, res = -1;   // tensor ids (res = -1: none; may equal out: in-place accumulate)
    int in2 = -1, in3 = -1, up2 = 1, up3 = 1;   // OP_UPADD: out = act(res + up(in) + up2(in2) + up3(in3))
    int cin = 0, cout = 0, k = 1, stride = 1, up = 1;
    int relu = 0;
    size_t w_off = 0;           // offset (in halfs) into the DEVICE weight blob: layout [tap][cout][cin], padded channel counts
    size_t b_off = 0;           // offset (in floats) into the device bias blob
    int cin_l = 0, cout_l = 0;  // logical channel counts and offsets: the layout of the blobs the caller passes (hbp_hrnet_describe)
    size_t w_off_l = 0, b_off_l = 0;
    int stream = 0;             // branch stream the op runs on
    int join_before = 0;        // all streams must have finished earlier ops before this op starts
    float sm_share = 0.f;       // > 0: fraction of the SMs this op's persistent launch may occupy (branches run side by side)
    std::vector<int> wait_ops;  // ops on OTHER streams that must have finished first (producers of this op's inputs since the last join)
    int signal = 0;             // some later op waits for this one: record an event after it
    int persist = 0;            // per-tap conv that owns the GPU while it runs: persistent tile walkers (conv_umma_pgroup_kernel)
    int grouped = 0;            // issued by the OP_GROUP / OP_UPADD_GROUP op that lists it in `members`
    std::vector<int> members;   // OP_GROUP / OP_UPADD_GROUP: indices of the member ops
    // L2 residency (tensors larger than a fraction of the 126 MB L2, i.e. layer1's 256-channel maps): reads of a tensor
    // whose last reader this op is carry the evict-first policy, and consecutive ops of a chain walk their tiles in
    // opposite directions so that an op starts on what its producer wrote last (conv_umma.cu: ConvParams::reverse)
    int reverse = 0, in_dead = 0, res_dead = 0, out_keep = 0;
};

struct UmmaPlan;                // conv_umma.cu: per-op tensor maps + tile shape (per batch size)
struct UmmaGroup;               // conv_umma.cu: device table of the member plans of one OP_GROUP launch
struct UmmaChain;               // conv_umma.cu: device table + dependency flags of one OP_CHAIN launch

struct HrnetModel {
    int width = 32, in_h = 256, in_w = 192;
    std::vector<HTensor> tensors;
    std::vector<HOp> ops;
    int n_bufs = 0;
    std::vector<size_t> buf_elems_per_image;   // halfs per image for each buffer
    size_t n_weights = 0, n_biases = 0;        // device blobs (padded channels)
    size_t n_weights_l = 0, n_biases_l = 0;    // caller's blobs (logical channels)
    __half* d_weights = nullptr;
    float* d_bias = nullptr;
    int engine = 0;             // 0 SIMT, 1 tcgen05 (falls back per-op where the shape is unsupported)
    // per-batch-size execution state
    int cap_P = 0;
    std::vector<__half*> bufs;
    std::vector<UmmaPlan*> umma;                // one per op (nullptr = SIMT)
    std::vector<UmmaGroup*> groups;             // one per op (OP_GROUP ops only)
    std::vector<UmmaChain*> chains;             // one per op (OP_CHAIN ops only)
    // CUDA graphs of the forward, keyed by (batch, input / output pointers, heatmap dtype, engine): a stream of frames
    // with a different person count each keeps one graph per count (least recently used of kMaxGraphs is dropped)
    struct HGraph {
        int P = 0, dtype = -1, engine = -1;
        const void* in = nullptr;
        void* out = nullptr;
        cudaGraphExec_t exec = nullptr;
        uint64_t nodes = 0, stamp = 0;
    };
    static constexpr int kMaxGraphs = 16;
    std::vector<HGraph> graphs;
    uint64_t graph_clock = 0;
    int graph_P = 0;                            // batch of the last forward (hrnet_debug_tensor sizes its copy by it)
    cudaStream_t side[7] = {};                  // streams 1.. (stream 0 is the context's): one per resolution branch
    cudaEvent_t ev_fork = nullptr, ev_join[7] = {};
    std::vector<cudaEvent_t> ev_pool;
    unsigned long long* d_timeline = nullptr;   // HBP_TIMELINE: [start, end] globaltimer ns per op, written by the kernels
    bool timeline_pending = false;
};

// builder (hrnet.cu)
void hrnet_build_program(HrnetModel& m);

// conv_umma.cu
bool umma_supported(const HrnetModel& m, const HOp& op);
int umma_plan_create(hbp_ctx* ctx, HrnetModel& m, int op_index, int P, UmmaPlan** out, bool for_group = false);
void umma_plan_destroy(UmmaPlan* p);
int umma_launch(hbp_ctx* ctx, HrnetModel& m, int op_index, UmmaPlan* plan, int P, cudaStream_t st);
// one launch for all member convs of the OP_GROUP op `group_index` (plans in m.umma, created with for_group)
int umma_group_launch(hbp_ctx* ctx, HrnetModel& m, int group_index, int P, cudaStream_t st);
void umma_group_destroy(UmmaGroup* g);
// one persistent launch for the member convs of the OP_CHAIN op `chain_index`, in order; returns 1 (nothing launched)
// when the members have to run as individual launches
int umma_chain_launch(hbp_ctx* ctx, HrnetModel& m, int chain_index, int P, cudaStream_t st);
void umma_chain_destroy(UmmaChain* c);
