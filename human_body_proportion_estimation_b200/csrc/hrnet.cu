// K5: HRNet-W32 / W48 pose network (public "pose_hrnet" definition, Sun et al.
// CVPR 2019) as a flat program of fused convolutions over NHWC fp16 tensors.
//
// Replaces the opaque network behind the reference's
// human_body_length_est/modules/pose_estimator.py:47-59 (onnxruntime session)
// and the Triton `hrnet` model of the ensemble
// (human_body_length_est/person_det_pose_edet4_trtserver.py:22-23).  The
// reference ships no network source; layer names follow the public HRNet
// state_dict so that real checkpoints map one to one.
//
// Every op is  out = act(conv(in) + bias [+ residual])  with BatchNorm folded
// into the weights at load time, optionally with the nearest-neighbour
// upsample of the fuse layers folded into the store.  Two engines execute the
// same program:
//   engine 0  SIMT implicit-GEMM tiles (this file)  -- any shape, the checker
//   engine 1  tcgen05/TMEM implicit GEMM fed by TMA (conv_umma.cu)
// The four resolution branches of a stage run on four streams; the whole
// forward is captured into one CUDA graph per (batch, buffers) key.
#include "hrnet.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

// ===========================================================================
// Program builder
// ===========================================================================
namespace {

constexpr int kStreams = 4;      // more streams were measured slower: every cross-stream edge of the graph costs ~10 us of start latency (profiles/r01_fuse_spread.log)

struct Builder {
    HrnetModel& m;
    std::vector<int> writer;        // tensor id -> index of the op that wrote it last
    int last_join_op = 0;           // ops before this index are ordered against everything by a global join
    // free buffers by size: usable by everyone / usable by one stream / waiting for the next join
    std::multimap<size_t, int> free_all;
    std::multimap<size_t, int> free_stream[kStreams];
    std::vector<std::pair<size_t, int>> pending;
    int cur_stream = 0;
    bool pending_join = false;

    explicit Builder(HrnetModel& mm) : m(mm) {}

    // Channel counts the tensor engine cannot take as they are (HRNet-W48's 48 and 96: its rows are 64 or 128 bytes
    // of K) are stored padded to the next multiple of 64; the extra channels carry zero weights and zero biases, so
    // they stay exactly zero through every conv / residual / ReLU and never change a logical channel.
    static int padc(int c) { return (c > 32 && c % 64 != 0) ? (c + 63) / 64 * 64 : c; }
    int new_tensor(int c_logical, int h, int w) {
        HTensor t;
        const int c = padc(c_logical);
        t.c = c; t.c_l = c_logical; t.h = h; t.w = w;
        const size_t sz = (size_t)c * h * w;
        auto take = [&](std::multimap<size_t, int>& pool) {
            auto it = pool.find(sz);
            if (it == pool.end()) return -1;
            int b = it->second;
            pool.erase(it);
            return b;
        };
        int b = take(free_stream[cur_stream]);
        if (b < 0) b = take(free_all);
        if (b < 0) {
            b = m.n_bufs++;
            m.buf_elems_per_image.push_back(sz);
        }
        t.buf = b;
        m.tensors.push_back(t);
        return (int)m.tensors.size() - 1;
    }
    // tensor no longer needed by ops issued so far.  same_stream_only: every reader
    // ran on cur_stream, so that stream may reuse the buffer at once.
    void release(int tid, bool same_stream_only) {
        const HTensor& t = m.tensors[tid];
        const size_t sz = (size_t)t.c * t.h * t.w;
        if (same_stream_only) free_stream[cur_stream].insert({sz, t.buf});
        else pending.push_back({sz, t.buf});
    }
    void join() {          // next op carries a global join
        pending_join = true;
        for (auto& p : pending) free_all.insert(p);
        pending.clear();
        for (int s = 0; s < kStreams; ++s) {
            for (auto& p : free_stream[s]) free_all.insert(p);
            free_stream[s].clear();
        }
        last_join_op = (int)m.ops.size();
    }
    // cross-stream read-after-write: the op about to be pushed (on `stream`) reads tensor `tid`
    void depend(HOp& op, int tid) {
        if (tid < 0 || tid >= (int)writer.size()) return;
        const int w = writer[tid];
        if (w < 0 || w < last_join_op || m.ops[w].stream == op.stream) return;
        op.wait_ops.push_back(w);
        m.ops[w].signal = 1;
    }
    void wrote(int tid, int op_index) {
        if ((int)writer.size() <= tid) writer.resize(tid + 1, -1);
        writer[tid] = op_index;
    }
    int conv(const std::string& name, int in, int cout, int k, int stride, int relu, int res = -1,
             int out = -1, int up = 1) {
        const HTensor ti = m.tensors[in];
        HOp op;
        op.kind = OP_CONV;
        op.name = name;
        op.in = in;
        const int cout_l = cout;
        cout = padc(cout_l);
        op.cin = ti.c; op.cout = cout; op.k = k; op.stride = stride; op.up = up; op.relu = relu;
        op.cin_l = ti.c_l; op.cout_l = cout_l;
        op.res = res;
        if (out < 0) out = new_tensor(cout_l, ti.h / stride * up, ti.w / stride * up);
        op.out = out;
        op.w_off = m.n_weights;
        op.b_off = m.n_biases;
        op.w_off_l = m.n_weights_l;
        op.b_off_l = m.n_biases_l;
        m.n_weights += (size_t)k * k * cout * ti.c;
        m.n_biases += cout;
        m.n_weights_l += (size_t)k * k * cout_l * ti.c_l;
        m.n_biases_l += cout_l;
        op.stream = cur_stream;
        op.join_before = pending_join ? 1 : 0;
        if (pending_join) last_join_op = (int)m.ops.size();
        pending_join = false;
        depend(op, in);
        depend(op, res);
        m.ops.push_back(op);
        wrote(out, (int)m.ops.size() - 1);
        return out;
    }
};

std::string S(const char* fmt, int a = 0, int b = 0, int c = 0, int d = 0) {
    char buf[160];
    snprintf(buf, sizeof(buf), fmt, a, b, c, d);
    return buf;
}

// one HighResolutionModule: `nb` branches x 4 BasicBlocks, then the fuse layers
void stage_module(Builder& B, const std::string& pre, std::vector<int>& x, const std::vector<int>& ch,
                  bool multi_scale_output) {
    HrnetModel& m = B.m;
    const int nb = (int)x.size();
    // The branches run side by side, one stream each, and every branch's persistent conv launches take a
    // fixed share of the SMs (~ its tensor-pipe + epilogue time per layer: equal FLOPs, but narrow N
    // and partly filled M-tiles cost more): the launches of a level are then resident at once, and
    // the launch gap / pipeline fill of one overlaps the steady state of the others
    // (profiles/r01_branch_share.log).  HBP_BRANCH_SHARE<nb>=a,b,.. overrides the table for nb branches.
    static const float kShare[5][4] = {{}, {1.f}, {0.64f, 0.36f}, {0.44f, 0.24f, 0.32f}, {0.38f, 0.20f, 0.24f, 0.18f}};     // (re-swept after the 32-byte stores / 216 KB plans: gpurun_out/r02y, r02z)
    float share[4] = {kShare[nb][0], kShare[nb][1], kShare[nb][2], kShare[nb][3]};
    {
        char key[32];
        snprintf(key, sizeof(key), "HBP_BRANCH_SHARE%d", nb);
        const char* e = getenv(key);
        if (!e) e = getenv("HBP_BRANCH_SHARE");
        if (e) {
            float v[4] = {0, 0, 0, 0};
            if (sscanf(e, "%f,%f,%f,%f", &v[0], &v[1], &v[2], &v[3]) >= nb) {
                float tot = 0;
                for (int i = 0; i < nb; ++i) tot += v[i];
                for (int i = 0; i < nb; ++i) share[i] = tot > 0 ? v[i] / tot : 0.f;
            }
        }
    }
    // The eight convolutions of a branch form an OP_CHAIN: one persistent launch whose CTAs walk the layers with per-tile
    // dependency flags instead of kernel boundaries (conv_umma_chain_kernel); the executor falls back to the members'
    // individual launches where the chain kernel does not apply (SIMT engine, the 256-channel per-tap branch).
    static const bool chained = getenv("HBP_CHAIN") ? atoi(getenv("HBP_CHAIN")) != 0 : false;
    for (int i = 0; i < nb; ++i) {
        B.cur_stream = i;
        std::vector<int> members;
        std::vector<int> waits;
        int join_flag = 0;
        auto took = [&]() {
            HOp& o = m.ops.back();
            o.sm_share = share[i];
            if (!chained) return;
            members.push_back((int)m.ops.size() - 1);
            join_flag |= o.join_before;
            for (int w : o.wait_ops) if (std::find(waits.begin(), waits.end(), w) == waits.end()) waits.push_back(w);
            o.join_before = 0; o.wait_ops.clear(); o.grouped = 1;
        };
        for (int blk = 0; blk < 4; ++blk) {
            const std::string p = pre + S(".branches.%d.%d", i, blk);
            const int t = B.conv(p + ".conv1", x[i], ch[i], 3, 1, 1);
            took();
            const int y = B.conv(p + ".conv2", t, ch[i], 3, 1, 1, /*res=*/x[i]);
            took();
            B.release(t, true);
            // the module input of blk 0 may have been produced on another stream /
            // is shared: only tensors created inside this loop are stream-local
            B.release(x[i], blk > 0);
            x[i] = y;
        }
        if (chained) {
            HOp g;
            g.kind = OP_CHAIN;
            g.name = pre + S(".branches.%d.chain", i);
            g.members = members;
            g.stream = i;
            g.join_before = join_flag;
            g.wait_ops = waits;
            g.sm_share = share[i];
            g.out = x[i];
            m.ops.push_back(g);
            B.wrote(x[i], (int)m.ops.size() - 1);
        }
    }
    // fuse: y_i = relu( sum_j f_ij(x_j) ), f_ii = identity.  Terms from higher-resolution
    // branches (j < i) are chains of i-j stride-2 convs whose last link accumulates into y_i; terms
    // from lower-resolution branches (j > i) are 1x1 convs at LOW resolution followed by ONE
    // coalesced upsample-add over all of them (a scatter inside the conv epilogue serialises
    // up*up dependent load/store pairs per thread).
    //
    // The whole stage is scheduled by LEVEL: link k of every chain and all the 1x1 convs (level 1)
    // are independent of each other, so each level is one grouped launch (OP_GROUP) and the
    // upsample-adds of all outputs are one launch (OP_UPADD_GROUP).  A dependent kernel level of the
    // captured graph costs ~10 us whatever its size, and the per-output chains used to put up to
    // six of them (plus cross-stream edges) on the critical path of each of the eight fuse stages.
    B.join();
    static const bool grouped = getenv("HBP_FUSE_GROUP") ? atoi(getenv("HBP_FUSE_GROUP")) != 0 : true;
    const int n_out = multi_scale_output ? nb : 1;
    std::vector<int> y(n_out);
    for (int i = 0; i < n_out; ++i) {
        const HTensor ti = m.tensors[x[i]];
        y[i] = B.new_tensor(ti.c_l, ti.h, ti.w);
    }
    const int n_levels = std::max(1, n_out - 1);
    std::vector<int> acc(n_out);                         // running sum of output i so far (starts at the identity term)
    for (int i = 0; i < n_out; ++i) acc[i] = x[i];
    std::vector<std::vector<int>> cur(n_out, std::vector<int>(nb, -1));   // chain (i, j): tensor its next link reads
    for (int i = 0; i < n_out; ++i) for (int j = 0; j < i; ++j) cur[i][j] = x[j];
    std::vector<std::vector<int>> lows(n_out);
    std::vector<std::vector<int>> ups(n_out);
    std::vector<int> temps;                              // intermediates, free after the closing join
    std::vector<int> group_of_level(n_levels + 1, -1);
    B.cur_stream = 0;
    for (int L = 1; L <= n_levels; ++L) {
        std::vector<int> members;
        int join_flag = 0;
        auto took = [&]() {
            members.push_back((int)m.ops.size() - 1);
            join_flag |= m.ops.back().join_before;
            if (grouped) { m.ops.back().join_before = 0; m.ops.back().grouped = 1; }
        };
        for (int i = L; i < n_out; ++i) {
            for (int j = 0; j + L <= i; ++j) {
                const int k = L - 1;                     // link index inside chain (i, j) of length i - j
                const std::string nm = pre + S(".fuse_layers.%d.%d.%d.0", i, j, k);
                if (L == i - j) {
                    // last link: accumulate into y_i; the deepest chain (j = 0) of the lowest-resolution
                    // output closes the sum (no upsample-add follows there) and applies the ReLU
                    const int last = (i == nb - 1) && (j == 0);
                    B.conv(nm, cur[i][j], ch[i], 3, 2, last, acc[i], y[i]);
                    acc[i] = y[i];
                } else {
                    const int nxt = B.conv(nm, cur[i][j], ch[j], 3, 2, 1);
                    temps.push_back(nxt);
                    cur[i][j] = nxt;
                }
                took();
            }
        }
        if (L == 1) {
            for (int i = 0; i < n_out && i < nb - 1; ++i)
                for (int j = i + 1; j < nb; ++j) {
                    const int t = B.conv(pre + S(".fuse_layers.%d.%d.0", i, j), x[j], ch[i], 1, 1, 0);
                    temps.push_back(t);
                    lows[i].push_back(t);
                    ups[i].push_back(1 << (j - i));
                    took();
                }
        }
        if (grouped && !members.empty()) {
            HOp g;
            g.kind = OP_GROUP;
            g.name = pre + S(".fuse_level%d", L);
            g.members = members;
            g.stream = 0;
            g.join_before = join_flag;
            m.ops.push_back(g);
            group_of_level[L] = (int)m.ops.size() - 1;
        }
    }
    // upsample-adds: output i (< nb-1) needs its 1x1 terms (level 1) and its chains (closed at level i).  Outputs
    // are grouped by the level they wait for: one launch per group, beside the deeper chain levels on stream 1,
    // so that only the smallest output (and the last chain link) is left after the last-but-one level.
    {
        static const bool split = getenv("HBP_UPADD_SPLIT") ? atoi(getenv("HBP_UPADD_SPLIT")) != 0 : true;
        const int n_up = std::min(n_out, nb - 1);
        const int max_need = std::max(1, n_up - 1);
        for (int need_level = 1; need_level <= max_need; ++need_level) {
            const bool side = grouped && need_level < n_levels;      // deeper chain levels still to run: beside them, on stream 1
            std::vector<int> members;
            int join_flag = 0;
            for (int i = 0; i < n_up; ++i) {
                const int need_i = split ? std::max(1, i) : max_need;
                if (need_i != need_level) continue;
                HOp op;
                op.kind = OP_UPADD;
                op.name = pre + S(".fuse_layers.%d.upadd", i);
                op.in = lows[i][0];
                op.in2 = lows[i].size() > 1 ? lows[i][1] : -1;
                op.in3 = lows[i].size() > 2 ? lows[i][2] : -1;
                op.up = ups[i][0];
                op.up2 = ups[i].size() > 1 ? ups[i][1] : 1;
                op.up3 = ups[i].size() > 2 ? ups[i][2] : 1;
                op.res = acc[i]; op.out = y[i];
                op.cin = op.cout = ch[i];
                op.relu = 1;
                op.stream = side ? 1 : 0;
                op.join_before = B.pending_join ? 1 : 0;
                if (B.pending_join) B.last_join_op = (int)m.ops.size();
                B.pending_join = false;
                join_flag |= op.join_before;
                if (grouped) { op.join_before = 0; op.grouped = 1; }
                m.ops.push_back(op);
                B.wrote(op.out, (int)m.ops.size() - 1);
                members.push_back((int)m.ops.size() - 1);
            }
            if (grouped && !members.empty()) {
                HOp g;
                g.kind = OP_UPADD_GROUP;
                g.name = pre + S(".fuse_upadd%d", need_level);
                g.members = members;
                g.stream = side ? 1 : 0;
                g.join_before = join_flag;
                if (side) {
                    g.wait_ops.push_back(group_of_level[need_level]);
                    m.ops[group_of_level[need_level]].signal = 1;
                }
                m.ops.push_back(g);
                // program order = issue order: the side-stream launch is recorded after the deeper levels
                // were pushed, but it only waits for `need_level` through its event
            }
        }
    }
    for (int t : temps) B.release(t, false);
    for (int j = 0; j < nb; ++j) B.release(x[j], false);
    B.join();
    x = y;
}

}  // namespace

void hrnet_build_program(HrnetModel& m) {
    m.tensors.clear(); m.ops.clear(); m.buf_elems_per_image.clear();
    m.n_bufs = 0; m.n_weights = 0; m.n_biases = 0; m.n_weights_l = 0; m.n_biases_l = 0;
    Builder B(m);
    const int C = m.width;
    const int H2 = m.in_h / 2, W2 = m.in_w / 2;
    // stem
    {
        HOp op;
        op.kind = OP_STEM1; op.name = "conv1";
        op.cin = 3; op.cout = 64; op.k = 3; op.stride = 2; op.relu = 1;
        op.out = B.new_tensor(64, H2, W2);
        op.cin_l = 3; op.cout_l = 64;
        op.w_off = m.n_weights; op.b_off = m.n_biases;
        op.w_off_l = m.n_weights_l; op.b_off_l = m.n_biases_l;
        m.n_weights += 9 * 64 * 3; m.n_biases += 64;
        m.n_weights_l += 9 * 64 * 3; m.n_biases_l += 64;
        m.ops.push_back(op);
    }
    int x = B.conv("conv2", m.ops[0].out, 64, 3, 2, 1);
    m.ops.back().persist = 1;
    B.release(m.ops[0].out, true);
    // layer1: 4 Bottlenecks (64 -> 256)
    for (int b = 0; b < 4; ++b) {
        const std::string p = S("layer1.%d", b);
        int res = x;
        if (b == 0) {
            // the projection shortcut is independent of the 1x1 -> 3x3 chain: it runs beside it on stream 1
            // and meets conv3 again through an event (HOp::wait_ops)
            B.cur_stream = 1;
            res = B.conv(p + ".downsample.0", x, 256, 1, 1, 0);
            B.cur_stream = 0;
        }
        // the 256-channel maps of this layer are 1.6 MB per image (100 MB at 64 crops, against 126 MB of L2): every conv
        // walks its tiles in the opposite direction of the one before it in the chain, so that it starts on what was
        // written last, and tensors are read with the evict-first policy by their last reader (HOp::reverse / in_dead)
        const int t1 = B.conv(p + ".conv1", x, 64, 1, 1, 1);
        m.ops.back().reverse = (b % 2 == 0);
        const int t2 = B.conv(p + ".conv2", t1, 64, 3, 1, 1);
        m.ops.back().reverse = (b % 2 == 1); m.ops.back().in_dead = 1;
        B.release(t1, true);
        const int y = B.conv(p + ".conv3", t2, 256, 1, 1, 1, res);
        m.ops.back().reverse = (b % 2 == 0); m.ops.back().in_dead = 1; m.ops.back().res_dead = 1; m.ops.back().out_keep = 1;
        B.release(t2, true);
        if (res != x) B.release(res, true);
        B.release(x, true);
        x = y;
    }
    // transition1
    std::vector<int> xs(2);
    xs[0] = B.conv("transition1.0.0", x, C, 3, 1, 1);
    m.ops.back().reverse = 1;          // layer1.3.conv3 wrote x front to back
    B.cur_stream = 1;                  // independent of transition1.0: side by side, on the stream of the branch it feeds
    xs[1] = B.conv("transition1.1.0.0", x, 2 * C, 3, 2, 1);
    m.ops.back().reverse = 1;

    B.cur_stream = 0;
    B.release(x, false);
    B.join();
    std::vector<int> ch = {C, 2 * C};
    stage_module(B, "stage2.0", xs, ch, true);
    // transition2: new branch from the last branch
    B.cur_stream = 2;
    xs.push_back(B.conv("transition2.2.0.0", xs[1], 4 * C, 3, 2, 1));

    ch.push_back(4 * C);
    for (int mod = 0; mod < 4; ++mod) stage_module(B, S("stage3.%d", mod), xs, ch, true);
    B.cur_stream = 3;
    xs.push_back(B.conv("transition3.3.0.0", xs[2], 8 * C, 3, 2, 1));

    ch.push_back(8 * C);
    for (int mod = 0; mod < 3; ++mod) stage_module(B, S("stage4.%d", mod), xs, ch, mod < 2);
    // head
    {
        B.cur_stream = 0;
        HOp op;
        op.kind = OP_HEAD; op.name = "final_layer";
        op.in = xs[0];
        op.cin = C; op.cout = 17; op.k = 1; op.stride = 1;
        op.cin = m.tensors[xs[0]].c; op.cin_l = C;
        op.cout_l = 17;
        op.w_off = m.n_weights; op.b_off = m.n_biases;
        op.w_off_l = m.n_weights_l; op.b_off_l = m.n_biases_l;
        m.n_weights += (size_t)17 * op.cin; m.n_biases += 17;
        m.n_weights_l += (size_t)17 * C; m.n_biases_l += 17;
        op.join_before = 1;
        m.ops.push_back(op);
    }
}

// ===========================================================================
// Engine 0 kernels
// ===========================================================================
namespace {

// conv1: NCHW fp16 (P,3,H,W) -> NHWC fp16 (P,H/2,W/2,64), 3x3 stride 2 pad 1, +bias, ReLU.
// K = 27 is far too thin for a TMA/tcgen05 pipeline and, as 1728 scalar FMAs per pixel, compute-bound on
// the fp32 pipes (110 us at batch 64, against an HBM floor of ~18 us for the 100 MB written).  Each warp
// therefore builds the im2col rows of its 32 pixels in shared memory (K padded to 32) and runs them
// through warp-level mma.sync m16n8k16 (fp16 x fp16 -> fp32): 32 MMAs per warp.  The warp's 32 pixels
// are 4 KB of contiguous output: staged in shared memory (XOR-swizzled 16-byte chunks) and written back
// as fully coalesced 512-byte rows.
constexpr int kStemLd = 40;      // halfs per im2col / weight row in shared memory (32 + 8: conflict-free fragment loads)

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128)
stem1_kernel(const __half* __restrict__ in, const __half* __restrict__ w, const float* __restrict__ bias,
             __half* __restrict__ out, int P, int H, int W) {
    __shared__ __align__(16) __half s_w[64 * kStemLd];          // [co][k], k = tap*3 + ci, zero for k >= 27
    __shared__ __align__(16) float s_b[64];
    __shared__ __align__(16) __half s_a[4][32 * kStemLd];       // per warp: [pixel][k]
    __shared__ __align__(16) uint4 s_out[4][32 * 8];
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
        const int co = i >> 5, k = i & 31;
        s_w[co * kStemLd + k] = k < 27 ? w[((size_t)(k / 3) * 64 + co) * 3 + (k % 3)] : __float2half(0.f);   // blob layout [tap][co][ci]
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_b[i] = bias[i];
    __syncthreads();                                   // weights + bias staged once per (persistent) block
    const int Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)P * Ho * Wo;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    // grid-stride over chunks of 128 pixels: the weight staging above (2048 scattered loads + a block barrier) used
    // to be paid by each of 6144 one-chunk blocks
    for (size_t chunk = blockIdx.x; chunk * blockDim.x < total; chunk += gridDim.x) {
    const size_t pix_raw = chunk * blockDim.x + threadIdx.x;
    const size_t pix = pix_raw < total ? pix_raw : total - 1;
    const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho), n = (int)(pix / ((size_t)Wo * Ho));
    // im2col row of this lane's pixel
    {
        __align__(16) __half v[32];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int hi = ho * 2 + dy - 1, wi = wo * 2 + dx - 1;
                const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci)
                    v[(dy * 3 + dx) * 3 + ci] = ok ? __ldg(in + (((size_t)n * 3 + ci) * H + hi) * W + wi) : __float2half(0.f);
            }
#pragma unroll
        for (int k = 27; k < 32; ++k) v[k] = __float2half(0.f);
        uint4* dst = reinterpret_cast<uint4*>(&s_a[wrp][lane * kStemLd]);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4*>(v)[q];
    }
    __syncwarp();                                      // im2col rows of this warp visible to its lanes
    const int g = lane >> 2, tq = lane & 3;            // fragment row group / thread-in-quad
    // A fragments: 2 M-tiles (16 pixels) x 2 K-steps
    uint32_t a[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const __half* r0 = &s_a[wrp][(mt * 16 + g) * kStemLd + ks * 16 + tq * 2];
            const __half* r1 = r0 + 8 * kStemLd;
            a[mt][ks][0] = *reinterpret_cast<const uint32_t*>(r0);
            a[mt][ks][1] = *reinterpret_cast<const uint32_t*>(r1);
            a[mt][ks][2] = *reinterpret_cast<const uint32_t*>(r0 + 8);
            a[mt][ks][3] = *reinterpret_cast<const uint32_t*>(r1 + 8);
        }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {                   // 8 output channels at a time
        float acc[2][4];
        const float2 bz = *reinterpret_cast<const float2*>(s_b + nt * 8 + tq * 2);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { acc[mt][0] = bz.x; acc[mt][1] = bz.y; acc[mt][2] = bz.x; acc[mt][3] = bz.y; }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const __half* br = &s_w[(nt * 8 + g) * kStemLd + ks * 16 + tq * 2];
            const uint32_t b0 = *reinterpret_cast<const uint32_t*>(br), b1 = *reinterpret_cast<const uint32_t*>(br + 8);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_16816(acc[mt], a[mt][ks], b0, b1);
        }
        // C fragment: rows g / g+8 of the M-tile, channels nt*8 + tq*2 + {0,1} -> 16-byte chunk nt of the pixel, word tq
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                const int pl = mt * 16 + hf * 8 + g;
                const __half2 r = __floats2half2_rn(fmaxf(acc[mt][2 * hf], 0.f), fmaxf(acc[mt][2 * hf + 1], 0.f));
                reinterpret_cast<__half2*>(&s_out[wrp][pl * 8 + (nt ^ (pl & 7))])[tq] = r;
            }
    }
    __syncwarp();
    const size_t warp_pix0 = chunk * blockDim.x + wrp * 32;
    uint4* o = reinterpret_cast<uint4*>(out + warp_pix0 * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int i = k * 32 + lane, pl = i >> 3, c = i & 7;
        if (warp_pix0 + pl < total) o[i] = s_out[wrp][pl * 8 + (c ^ (pl & 7))];
    }
    __syncwarp();                                      // s_a / s_out of this warp are rewritten by its next chunk
    }
}

// final_layer: NHWC fp16 (P,Hh,Wh,C) -> NCHW (P,17,Hh,Wh) fp16|fp32, 1x1 + bias
template <typename OutT>
__global__ void __launch_bounds__(128)
head_kernel(const __half* __restrict__ in, const __half* __restrict__ w, const float* __restrict__ bias,
            OutT* __restrict__ out, int P, int HW, int C) {
    extern __shared__ float s_hw[];          // [17][C] then bias[17]
    for (int i = threadIdx.x; i < 17 * C; i += blockDim.x) s_hw[i] = __half2float(w[i]);
    for (int i = threadIdx.x; i < 17; i += blockDim.x) s_hw[17 * C + i] = bias[i];
    __syncthreads();
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (size_t)P * HW) return;
    const int n = (int)(pix / HW), hw = (int)(pix % HW);
    float acc[17];
#pragma unroll
    for (int j = 0; j < 17; ++j) acc[j] = s_hw[17 * C + j];
    const uint4* src = reinterpret_cast<const uint4*>(in + pix * C);
    for (int c8 = 0; c8 < C / 8; ++c8) {
        const uint4 q = __ldg(src + c8);
        const __half2* h2 = reinterpret_cast<const __half2*>(&q);
        float f[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float2 t = __half22float2(h2[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
#pragma unroll
        for (int j = 0; j < 17; ++j) {
            const float* ww = s_hw + j * C + c8 * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[j] = fmaf(f[k], ww[k], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 17; ++j) {
        OutT* o = out + ((size_t)n * 17 + j) * HW + hw;
        if constexpr (sizeof(OutT) == 2) *o = __float2half_rn(acc[j]);
        else *o = acc[j];
    }
}

// out = act(res + up(a) [+ up2(b)] [+ up3(c)]), NHWC fp16, nearest upsample.  HBM-bound: every thread
// handles two 16-byte items half the tensor apart, all their loads issued before the first use.
__global__ void __launch_bounds__(256)
upsample_add_kernel(const __half* __restrict__ a, const __half* __restrict__ b, const __half* __restrict__ c,
                    const __half* res, __half* out, int P, int H, int W, int C, int fa, int fb, int fc, int relu) {
    const int c8n = C >> 3;
    const size_t total = (size_t)P * H * W * c8n;
    const size_t half_n = (total + 1) >> 1;
    const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= half_n) return;
    size_t idx[2] = {i0, i0 + half_n};
    uint4 q[2][4];
    bool on[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        on[u] = idx[u] < total;
        if (!on[u]) continue;
        const int c8 = (int)(idx[u] % c8n);
        const size_t pix = idx[u] / c8n;
        const int w = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
        auto src = [&](const __half* s, int f) {
            const int hs = H / f, ws = W / f;
            return __ldg(reinterpret_cast<const uint4*>(s + (((size_t)n * hs + h / f) * ws + w / f) * C + c8 * 8));
        };
        q[u][0] = *reinterpret_cast<const uint4*>(res + pix * C + c8 * 8);
        q[u][1] = src(a, fa);
        if (b) q[u][2] = src(b, fb);
        if (c) q[u][3] = src(c, fc);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        if (!on[u]) continue;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        auto add = [&](const uint4& v) {
            const __half2* hq = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 t = __half22float2(hq[k]); acc[2 * k] += t.x; acc[2 * k + 1] += t.y; }
        };
        add(q[u][0]); add(q[u][1]);
        if (b) add(q[u][2]);
        if (c) add(q[u][3]);
        __align__(16) __half2 pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float x0 = acc[2 * k], x1 = acc[2 * k + 1];
            if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
            pk[k] = __floats2half2_rn(x0, x1);
        }
        *reinterpret_cast<uint4*>(out + idx[u] * 8) = *reinterpret_cast<const uint4*>(pk);
    }
}

// All upsample-adds of one fuse stage as one launch: blockIdx.y picks the output.
struct UpaddProblem {
    const __half *a, *b, *c, *res;
    __half* out;
    int H, W, C, fa, fb, fc, relu;
    unsigned blocks;
};
struct UpaddGroup { UpaddProblem p[3]; };

__global__ void __launch_bounds__(256)
upsample_add_group_kernel(const __grid_constant__ UpaddGroup g, int P) {
    // grid-stride over the problem's blocks with a few CTAs per SM only: the launch then leaves room for the
    // persistent conv group (last links of the stride-2 chains) that runs beside it on the other stream.
    // 32-bit index arithmetic throughout (P*H*W*C/8 < 2^31): 64-bit div/mod dominated the old kernel.
    const UpaddProblem& q = g.p[blockIdx.y];
    const unsigned H = q.H, W = q.W, C = q.C;
    const unsigned c8n = C >> 3;
    const unsigned total = (unsigned)P * H * W * c8n;
    const unsigned half_n = (total + 1) >> 1;
    const bool has_b = q.b != nullptr, has_c = q.c != nullptr;
    for (unsigned blk = blockIdx.x; blk < q.blocks; blk += gridDim.x) {
        const unsigned i0 = blk * blockDim.x + threadIdx.x;
        if (i0 >= half_n) continue;
        const unsigned idx[2] = {i0, i0 + half_n};
        uint4 v[2][4];
        bool on[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            on[u] = idx[u] < total;
            if (!on[u]) continue;
            const unsigned c8 = idx[u] % c8n;
            const unsigned pix = idx[u] / c8n;
            const unsigned w = pix % W, hn = pix / W, h = hn % H, n = hn / H;
            auto src = [&](const __half* s, unsigned f) {
                const unsigned hs = H / f, ws = W / f;
                return __ldg(reinterpret_cast<const uint4*>(s + ((size_t)(n * hs + h / f) * ws + w / f) * C + c8 * 8));
            };
            v[u][0] = *reinterpret_cast<const uint4*>(q.res + (size_t)idx[u] * 8);
            v[u][1] = src(q.a, q.fa);
            if (has_b) v[u][2] = src(q.b, q.fb);
            if (has_c) v[u][3] = src(q.c, q.fc);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!on[u]) continue;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            auto add = [&](const uint4& t4) {
                const __half2* hq = reinterpret_cast<const __half2*>(&t4);
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float2 t = __half22float2(hq[k]); acc[2 * k] += t.x; acc[2 * k + 1] += t.y; }
            };
            add(v[u][0]); add(v[u][1]);
            if (has_b) add(v[u][2]);
            if (has_c) add(v[u][3]);
            __align__(16) __half2 pk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float x0 = acc[2 * k], x1 = acc[2 * k + 1];
                if (q.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                pk[k] = __floats2half2_rn(x0, x1);
            }
            *reinterpret_cast<uint4*>(q.out + (size_t)idx[u] * 8) = *reinterpret_cast<const uint4*>(pk);
        }
    }
}

// Generic fused conv, NHWC fp16, weights [tap][cout][cin], fp32 accumulate.
// CTA tile: 64 output pixels x 64 output channels, 256 threads, 4x4 per thread.
constexpr int kTP = 64, kTC = 64, kTK = 32;

__global__ void __launch_bounds__(256)
conv_simt_kernel(const __half* __restrict__ in, const __half* __restrict__ w, const float* __restrict__ bias,
                 const __half* res, __half* out, int P, int Hi, int Wi, int Cin, int Ho, int Wo,
                 int Cout, int k, int stride, int up, int relu) {
    __shared__ __align__(16) float As[kTK][kTP + 4];
    __shared__ __align__(16) float Bs[kTK][kTC + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const size_t total_px = (size_t)P * Ho * Wo;
    const size_t px0 = (size_t)blockIdx.x * kTP;
    const int co0 = blockIdx.y * kTC;
    const int pad = k / 2;
    // loader mapping: 4 threads per row, 8 channels (16 B) each
    const int lrow = tid >> 2, lci = (tid & 3) * 8;
    const size_t lpx = px0 + lrow;
    int ln = 0, lho = 0, lwo = 0;
    const bool lpx_ok = lpx < total_px;
    if (lpx_ok) { lwo = (int)(lpx % Wo); lho = (int)((lpx / Wo) % Ho); ln = (int)(lpx / ((size_t)Wo * Ho)); }
    const int lco = co0 + lrow;
    float acc[4][4] = {};
    for (int tap = 0; tap < k * k; ++tap) {
        const int dy = tap / k - pad, dx = tap % k - pad;
        const int hi = lho * stride + dy, wi = lwo * stride + dx;
        const bool a_ok = lpx_ok && hi >= 0 && hi < Hi && wi >= 0 && wi < Wi;
        const __half* a_src = in + (((size_t)ln * Hi + hi) * Wi + wi) * Cin;
        const __half* b_src = w + ((size_t)tap * Cout + lco) * Cin;
        for (int c0 = 0; c0 < Cin; c0 += kTK) {
            uint4 qa = make_uint4(0, 0, 0, 0), qb = make_uint4(0, 0, 0, 0);
            const bool c_ok = c0 + lci < Cin;
            if (a_ok && c_ok) qa = __ldg(reinterpret_cast<const uint4*>(a_src + c0 + lci));
            if (lco < Cout && c_ok) qb = __ldg(reinterpret_cast<const uint4*>(b_src + c0 + lci));
            const __half2* ha = reinterpret_cast<const __half2*>(&qa);
            const __half2* hb = reinterpret_cast<const __half2*>(&qb);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 fa = __half22float2(ha[q]), fb = __half22float2(hb[q]);
                As[lci + 2 * q][lrow] = fa.x; As[lci + 2 * q + 1][lrow] = fa.y;
                Bs[lci + 2 * q][lrow] = fb.x; Bs[lci + 2 * q + 1][lrow] = fb.y;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < kTK; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
    }
    // epilogue
    const int co = co0 + tx * 4;
    if (co >= Cout) return;
    float bz[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bz[j] = (co + j < Cout) ? bias[co + j] : 0.f;
    const int Hout = Ho * up, Wout = Wo * up;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const size_t px = px0 + ty * 4 + i;
        if (px >= total_px) continue;
        const int wo = (int)(px % Wo), ho = (int)((px / Wo) % Ho), n = (int)(px / ((size_t)Wo * Ho));
        for (int uy = 0; uy < up; ++uy)
            for (int ux = 0; ux < up; ++ux) {
                const size_t o = ((((size_t)n * Hout + ho * up + uy) * Wout) + wo * up + ux) * Cout + co;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bz[j];
                if (res) {
                    const uint2 q = *reinterpret_cast<const uint2*>(res + o);
                    const __half2* h = reinterpret_cast<const __half2*>(&q);
                    const float2 r0 = __half22float2(h[0]), r1 = __half22float2(h[1]);
                    v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y;
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                __align__(8) __half2 pk[2] = {__floats2half2_rn(v[0], v[1]), __floats2half2_rn(v[2], v[3])};
                *reinterpret_cast<uint2*>(out + o) = *reinterpret_cast<const uint2*>(pk);
            }
    }
}

}  // namespace

__global__ void timeline_stamp_kernel(unsigned long long* slot, int is_end) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (is_end) atomicMax(slot + 1, t); else atomicMin(slot, t);
}

// ===========================================================================
// Executor
// ===========================================================================
void hrnet_dims(hbp_ctx* ctx, int* h, int* w, int* width) {
    *h = ctx->hrnet ? ctx->hrnet->in_h : 0;
    *w = ctx->hrnet ? ctx->hrnet->in_w : 0;
    *width = ctx->hrnet ? ctx->hrnet->width : 0;
}

static void free_batch_state(hbp_ctx* ctx, HrnetModel* m) {
    cudaStreamSynchronize(ctx->stream);
    for (HrnetModel::HGraph& g : m->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    m->graphs.clear();
    for (UmmaPlan* p : m->umma) if (p) umma_plan_destroy(p);
    m->umma.clear();
    for (UmmaGroup* g : m->groups) if (g) umma_group_destroy(g);
    m->groups.clear();
    for (UmmaChain* c : m->chains) if (c) umma_chain_destroy(c);
    m->chains.clear();
    for (__half* b : m->bufs) if (b) cudaFree(b);
    m->bufs.clear();
    m->cap_P = 0;
    m->graph_P = 0;
}

void hrnet_free(hbp_ctx* ctx) {
    HrnetModel* m = ctx->hrnet;
    if (!m) return;
    free_batch_state(ctx, m);
    if (m->d_timeline) cudaFree(m->d_timeline);
    if (m->d_weights) cudaFree(m->d_weights);
    if (m->d_bias) cudaFree(m->d_bias);
    for (int i = 0; i < kStreams - 1; ++i) {
        if (m->side[i]) cudaStreamDestroy(m->side[i]);
        if (m->ev_join[i]) cudaEventDestroy(m->ev_join[i]);
    }
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    for (cudaEvent_t e : m->ev_pool) if (e) cudaEventDestroy(e);
    delete m;
    ctx->hrnet = nullptr;
}

int hrnet_load(hbp_ctx* ctx, int width, int in_h, int in_w, const void* w16, size_t nw,
               const float* bias, size_t nb) {
    hrnet_free(ctx);
    HrnetModel* m = new HrnetModel();
    m->width = width; m->in_h = in_h; m->in_w = in_w;
    hrnet_build_program(*m);
    if (nw != m->n_weights_l || nb != m->n_biases_l) {
        hbp_set_error("HRNet-W%d expects %zu weights and %zu biases, got %zu / %zu", width, m->n_weights_l,
                      m->n_biases_l, nw, nb);
        delete m;
        return HBP_ERR_INVALID;
    }
    ctx->hrnet = m;
    // caller's blobs (logical channels, hbp_hrnet_describe order) -> device blobs (channels padded with zeros)
    std::vector<__half> wp(m->n_weights, __float2half(0.f));
    std::vector<float> bp(m->n_biases, 0.f);
    const __half* wl = static_cast<const __half*>(w16);
    for (const HOp& op : m->ops) {
        if (op.kind != OP_CONV && op.kind != OP_STEM1 && op.kind != OP_HEAD) continue;
        const int taps = op.kind == OP_HEAD ? 1 : op.k * op.k;
        const int cout_p = op.kind == OP_HEAD ? 17 : op.cout, cout_l = op.cout_l, cin_p = op.cin, cin_l = op.cin_l;
        for (int t = 0; t < taps; ++t)
            for (int co = 0; co < cout_l; ++co)
                memcpy(&wp[op.w_off + ((size_t)t * cout_p + co) * cin_p], &wl[op.w_off_l + ((size_t)t * cout_l + co) * cin_l],
                       (size_t)cin_l * sizeof(__half));
        memcpy(&bp[op.b_off], &bias[op.b_off_l], (size_t)cout_l * sizeof(float));
    }
    HBP_CUDA(cudaMalloc(&m->d_weights, m->n_weights * sizeof(__half)));
    HBP_CUDA(cudaMalloc(&m->d_bias, m->n_biases * sizeof(float)));
    HBP_CUDA(cudaMemcpyAsync(m->d_weights, wp.data(), m->n_weights * sizeof(__half), cudaMemcpyHostToDevice, ctx->stream));
    HBP_CUDA(cudaMemcpyAsync(m->d_bias, bp.data(), m->n_biases * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    for (int i = 0; i < kStreams - 1; ++i) {
        HBP_CUDA(cudaStreamCreateWithFlags(&m->side[i], cudaStreamNonBlocking));
        HBP_CUDA(cudaEventCreateWithFlags(&m->ev_join[i], cudaEventDisableTiming));
    }
    HBP_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    HBP_CUDA(cudaStreamSynchronize(ctx->stream));
    m->engine = 1;
    return HBP_OK;
}

int hrnet_set_engine(hbp_ctx* ctx, int engine) {
    if (!ctx->hrnet) { hbp_set_error("no model loaded"); return HBP_ERR_STATE; }
    if (engine != 0 && engine != 1) { hbp_set_error("engine must be 0 (SIMT) or 1 (tcgen05)"); return HBP_ERR_INVALID; }
    ctx->hrnet->engine = engine;
    return HBP_OK;
}

static int ensure_batch(hbp_ctx* ctx, HrnetModel* m, int P) {
    if (P <= m->cap_P) return HBP_OK;
    free_batch_state(ctx, m);
    int cap = 8;
    while (cap < P) cap *= 2;
    m->bufs.assign(m->n_bufs, nullptr);
    for (int b = 0; b < m->n_bufs; ++b)
        HBP_CUDA(cudaMalloc(&m->bufs[b], m->buf_elems_per_image[b] * (size_t)cap * sizeof(__half)));
    m->cap_P = cap;
    if (!(ctx->attr_flags & ATTR_CONV)) {
        // The element-wise kernels run beside tensor-core CTAs that need ~200 KB of shared memory.  An SM
        // changes its L1 / shared-memory split only when it is idle, so a streaming kernel that keeps the
        // default (L1-heavy) split on every SM locks those CTAs out until it has finished (measured: the
        // last stride-2 links of a fuse stage started 15-24 us late, profiles/r01_timeline_fuse_levels.md).
        HBP_CUDA(cudaFuncSetAttribute(upsample_add_group_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        HBP_CUDA(cudaFuncSetAttribute(upsample_add_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        HBP_CUDA(cudaFuncSetAttribute(timeline_stamp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ctx->attr_flags |= ATTR_CONV;
    }
    m->umma.assign(m->ops.size(), nullptr);
    m->groups.assign(m->ops.size(), nullptr);
    m->chains.assign(m->ops.size(), nullptr);
    return HBP_OK;
}

static cudaStream_t stream_of(hbp_ctx* ctx, HrnetModel* m, int s) {
    static const bool one = getenv("HBP_ONE_STREAM") != nullptr;      // bring-up: serialise the branches on the origin stream
    return (s == 0 || one) ? ctx->stream : m->side[s - 1];
}

// all streams wait for each other.  Every stream records one event and waits for the events of the
// others directly: in the captured graph the first kernel after the join then depends on the last
// kernels before it through ONE edge each (gathering into the origin stream and forking out again put
// two empty nodes on every path, ~15 us per join on the device timeline).
static int join_all(hbp_ctx* ctx, HrnetModel* m) {
    static const bool one = getenv("HBP_ONE_STREAM") != nullptr;
    if (one) return HBP_OK;
    HBP_CUDA(cudaEventRecord(m->ev_fork, ctx->stream));
    for (int s = 1; s < kStreams; ++s) HBP_CUDA(cudaEventRecord(m->ev_join[s - 1], stream_of(ctx, m, s)));
    for (int s = 0; s < kStreams; ++s)
        for (int t = 0; t < kStreams; ++t) {
            if (t == s) continue;
            HBP_CUDA(cudaStreamWaitEvent(stream_of(ctx, m, s), t == 0 ? m->ev_fork : m->ev_join[t - 1], 0));
        }
    return HBP_OK;
}

static int issue_ops(hbp_ctx* ctx, HrnetModel* m, const __half* crops, int P, void* heatmaps, int out_dtype,
                     uint64_t* n_launch, int stop_after = -1) {
    // fork: side streams join the origin stream (required for capture)
    HBP_CUDA(cudaEventRecord(m->ev_fork, ctx->stream));
    for (int s = 1; s < kStreams; ++s) HBP_CUDA(cudaStreamWaitEvent(stream_of(ctx, m, s), m->ev_fork, 0));
    if (m->ev_pool.size() < m->ops.size()) m->ev_pool.resize(m->ops.size(), nullptr);
    uint64_t launches = 0;
    // bring-up: HBP_OP_TIMING=1 (with HBP_NO_GRAPH=1 HBP_ONE_STREAM=1) prints the device time of every op
    static const bool op_timing = getenv("HBP_OP_TIMING") != nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(ctx->stream, &cap);
    const bool timing = op_timing && cap == cudaStreamCaptureStatusNone;
    std::vector<cudaEvent_t> tev;
    if (timing) {
        tev.resize(2 * m->ops.size());
        for (auto& e : tev) cudaEventCreate(&e);
    }
    // one op (conv / upsample-add) on stream `st`; the groups below fall back to this per member
    auto issue_upadd = [&](const HOp& op, cudaStream_t st) {
        const HTensor& to = m->tensors[op.out];
        const size_t total = ((size_t)P * to.h * to.w * (to.c / 8) + 1) / 2;      // two items per thread
        upsample_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            m->bufs[m->tensors[op.in].buf], op.in2 >= 0 ? m->bufs[m->tensors[op.in2].buf] : nullptr,
            op.in3 >= 0 ? m->bufs[m->tensors[op.in3].buf] : nullptr, m->bufs[m->tensors[op.res].buf],
            m->bufs[to.buf], P, to.h, to.w, to.c, op.up, op.up2, op.up3, op.relu);
    };
    auto issue_conv = [&](size_t i, cudaStream_t st) -> int {
        const HOp& op = m->ops[i];
        if (m->engine == 1 && umma_supported(*m, op)) {
            if (!m->umma[i]) {
                int s = umma_plan_create(ctx, *m, (int)i, m->cap_P, &m->umma[i]);
                if (s) return s;
            }
            return umma_launch(ctx, *m, (int)i, m->umma[i], P, st);
        }
        const HTensor& ti = m->tensors[op.in];
        const int Ho = ti.h / op.stride, Wo = ti.w / op.stride;
        const size_t total = (size_t)P * Ho * Wo;
        dim3 grid((unsigned)((total + kTP - 1) / kTP), (op.cout + kTC - 1) / kTC);
        conv_simt_kernel<<<grid, 256, 0, st>>>(
            m->bufs[ti.buf], m->d_weights + op.w_off, m->d_bias + op.b_off,
            op.res >= 0 ? m->bufs[m->tensors[op.res].buf] : nullptr, m->bufs[m->tensors[op.out].buf],
            P, ti.h, ti.w, op.cin, Ho, Wo, op.cout, op.k, op.stride, op.up, op.relu);
        return HBP_OK;
    };
    for (size_t i = 0; i < m->ops.size(); ++i) {
        const HOp& op = m->ops[i];
        if (op.grouped) continue;            // issued by the group op that lists it
        if (op.join_before) { int s = join_all(ctx, m); if (s) return s; }
        cudaStream_t st = stream_of(ctx, m, op.stream);
        for (int w : op.wait_ops) HBP_CUDA(cudaStreamWaitEvent(st, m->ev_pool[w], 0));
        if (timing) cudaEventRecord(tev[2 * i], st);
        const bool stamp = m->d_timeline && op.kind != OP_CONV;      // conv kernels stamp themselves
        if (stamp) timeline_stamp_kernel<<<1, 1, 0, st>>>(m->d_timeline + 2 * i, 0);
        if (op.kind == OP_STEM1) {
            const HTensor& to = m->tensors[op.out];
            const size_t total = (size_t)P * to.h * to.w;
            stem1_kernel<<<(unsigned)std::min<size_t>((total + 127) / 128, (size_t)ctx->sm_count * 7), 128, 0, st>>>(
                crops, m->d_weights + op.w_off, m->d_bias + op.b_off, m->bufs[to.buf], P, m->in_h, m->in_w);
        } else if (op.kind == OP_HEAD) {
            const HTensor& ti = m->tensors[op.in];
            const size_t total = (size_t)P * ti.h * ti.w;
            const size_t sm = (size_t)(17 * ti.c + 17) * sizeof(float);
            if (out_dtype == HBP_F16)
                head_kernel<__half><<<(unsigned)((total + 127) / 128), 128, sm, st>>>(
                    m->bufs[ti.buf], m->d_weights + op.w_off, m->d_bias + op.b_off, (__half*)heatmaps, P, ti.h * ti.w, ti.c);
            else
                head_kernel<float><<<(unsigned)((total + 127) / 128), 128, sm, st>>>(
                    m->bufs[ti.buf], m->d_weights + op.w_off, m->d_bias + op.b_off, (float*)heatmaps, P, ti.h * ti.w, ti.c);
        } else if (op.kind == OP_UPADD) {
            issue_upadd(op, st);
        } else if (op.kind == OP_UPADD_GROUP) {
            if (op.members.size() > 3) { hbp_set_error("upsample-add group of %zu", op.members.size()); return HBP_ERR_INVALID; }
            UpaddGroup g = {};
            unsigned max_blocks = 0;
            for (size_t k = 0; k < op.members.size(); ++k) {
                const HOp& u = m->ops[op.members[k]];
                const HTensor& to = m->tensors[u.out];
                UpaddProblem& q = g.p[k];
                q.a = m->bufs[m->tensors[u.in].buf];
                q.b = u.in2 >= 0 ? m->bufs[m->tensors[u.in2].buf] : nullptr;
                q.c = u.in3 >= 0 ? m->bufs[m->tensors[u.in3].buf] : nullptr;
                q.res = m->bufs[m->tensors[u.res].buf];
                q.out = m->bufs[to.buf];
                q.H = to.h; q.W = to.w; q.C = to.c; q.fa = u.up; q.fb = u.up2; q.fc = u.up3; q.relu = u.relu;
                const size_t total = ((size_t)P * to.h * to.w * (to.c / 8) + 1) / 2;
                q.blocks = (unsigned)((total + 255) / 256);
                max_blocks = std::max(max_blocks, q.blocks);
            }
            static const int upadd_bpsm = getenv("HBP_UPADD_BPSM") ? atoi(getenv("HBP_UPADD_BPSM")) : 8;      // blocks per SM over all problems
            const unsigned per_problem = (unsigned)std::max<size_t>(1, (size_t)ctx->sm_count * upadd_bpsm / op.members.size());
            upsample_add_group_kernel<<<dim3(std::min(max_blocks, per_problem), (unsigned)op.members.size()), 256, 0, st>>>(g, P);
        } else if (op.kind == OP_CHAIN) {
            int s = m->engine == 1 ? umma_chain_launch(ctx, *m, (int)i, P, st) : 1;
            if (s < 0) return s;
            if (s == 1) {                    // member by member
                for (int k : op.members) { int s2 = issue_conv((size_t)k, st); if (s2) return s2; launches++; }
                --launches;
            }
        } else if (op.kind == OP_GROUP) {
            // grouped launch when every member runs on the tensor engine; member by member otherwise
            bool all = m->engine == 1;
            for (int k : op.members) all = all && umma_supported(*m, m->ops[k]);
            if (all) {
                for (int k : op.members)
                    if (!m->umma[k]) {
                        int s = umma_plan_create(ctx, *m, k, m->cap_P, &m->umma[k], /*for_group=*/true);
                        if (s) return s;
                    }
                int s = umma_group_launch(ctx, *m, (int)i, P, st);
                if (s) return s;
            } else {
                for (int k : op.members) { int s = issue_conv((size_t)k, st); if (s) return s; launches++; }
                --launches;
            }
        } else {
            int s = issue_conv(i, st);
            if (s) return s;
        }
        ++launches;
        if (op.signal) {
            if (!m->ev_pool[i]) HBP_CUDA(cudaEventCreateWithFlags(&m->ev_pool[i], cudaEventDisableTiming));
            HBP_CUDA(cudaEventRecord(m->ev_pool[i], st));
        }
        if (stamp) timeline_stamp_kernel<<<1, 1, 0, st>>>(m->d_timeline + 2 * i, 1);
        if (timing) cudaEventRecord(tev[2 * i + 1], st);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return hbp_cuda_fail(e, op.name.c_str(), __FILE__, __LINE__);
        if (stop_after >= 0 && (int)i >= stop_after) break;      // parity hook (hrnet_forward_until)
    }
    if (timing) {
        cudaDeviceSynchronize();
        for (size_t i = 0; i < m->ops.size(); ++i) {
            const HOp& op = m->ops[i];
            float ms = 0.f;
            cudaEventElapsedTime(&ms, tev[2 * i], tev[2 * i + 1]);
            fprintf(stderr, "[op] %-40s kind=%d cin=%d cout=%d k=%d s=%d stream=%d %8.2f us\n", op.name.c_str(), op.kind, op.cin, op.cout,
                    op.k, op.stride, op.stream, ms * 1e3f);
        }
        for (auto& e : tev) cudaEventDestroy(e);
    }
    // final join back into the origin stream
    for (int s = 1; s < kStreams; ++s) {
        HBP_CUDA(cudaEventRecord(m->ev_join[s - 1], stream_of(ctx, m, s)));
        HBP_CUDA(cudaStreamWaitEvent(ctx->stream, m->ev_join[s - 1], 0));
    }
    *n_launch = launches;
    return HBP_OK;
}

int hrnet_forward(hbp_ctx* ctx, const __half* crops, int P, void* heatmaps, int out_dtype) {
    HrnetModel* m = ctx->hrnet;
    if (!m) { hbp_set_error("no model loaded"); return HBP_ERR_STATE; }
    int s = ensure_batch(ctx, m, P);
    if (s) return s;
    static const bool want_tl = getenv("HBP_TIMELINE") != nullptr;
    if (want_tl) {
        const size_t n = m->ops.size();
        if (!m->d_timeline) HBP_CUDA(cudaMalloc(&m->d_timeline, 2 * n * sizeof(unsigned long long)));
        if (m->timeline_pending) {
            // dump the stamps of the previous forward (device ns, relative to its first op)
            HBP_CUDA(cudaStreamSynchronize(ctx->stream));
            std::vector<unsigned long long> h(2 * n);
            HBP_CUDA(cudaMemcpy(h.data(), m->d_timeline, 2 * n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            unsigned long long t0 = ~0ull;
            for (size_t i = 0; i < n; ++i) if (h[2 * i] < t0) t0 = h[2 * i];
            for (size_t i = 0; i < n; ++i) {
                const HOp& op = m->ops[i];
                if (h[2 * i] == ~0ull) continue;
                fprintf(stderr, "[tl] %3zu %-40s kind=%d cin=%d cout=%d k=%d s=%d stream=%d start %9.2f end %9.2f us\n", i, op.name.c_str(), op.kind,
                        op.cin, op.cout, op.k, op.stride, op.stream, (double)(h[2 * i] - t0) * 1e-3, (double)(h[2 * i + 1] - t0) * 1e-3);
            }
            fprintf(stderr, "[tl] ----\n");
        }
        std::vector<unsigned long long> init(2 * n);
        for (size_t i = 0; i < n; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0; }
        HBP_CUDA(cudaMemcpyAsync(m->d_timeline, init.data(), 2 * n * sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
        HBP_CUDA(cudaStreamSynchronize(ctx->stream));
        m->timeline_pending = true;
    }
    static const bool no_graph = getenv("HBP_NO_GRAPH") != nullptr;
    m->graph_P = P;
    uint64_t n = 0;
    if (no_graph) {
        s = issue_ops(ctx, m, crops, P, heatmaps, out_dtype, &n);
        if (s) return s;
        ctx->launches += n;
        return HBP_OK;
    }
    HrnetModel::HGraph* slot = nullptr;
    for (HrnetModel::HGraph& g : m->graphs)
        if (g.P == P && g.dtype == out_dtype && g.in == (const void*)crops && g.out == heatmaps && g.engine == m->engine) { slot = &g; break; }
    if (slot && slot->exec) {
        slot->stamp = ++m->graph_clock;
        HBP_CUDA(cudaGraphLaunch(slot->exec, ctx->stream));
        ctx->launches += slot->nodes;
        return HBP_OK;
    }
    if (slot) {
        // second call with this key: capture it (plans and group tables exist since the eager first call, so nothing is
        // allocated or encoded inside the capture)
        cudaGraph_t g = nullptr;
        // (defensive: the capture re-records the events of the fork / join edges the eager first call may still be waiting
        // on; once per key, so the wait costs nothing in steady state.  The intermittent launch failure first attributed to
        // this turned out to be the third epilogue team of the halo kernel: profiles/r02_epilogue_ablation.md)
        HBP_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < kStreams - 1; ++i) if (m->side[i]) HBP_CUDA(cudaStreamSynchronize(m->side[i]));
        HBP_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        s = issue_ops(ctx, m, crops, P, heatmaps, out_dtype, &n);
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        if (s) { if (g) cudaGraphDestroy(g); return s; }
        if (e != cudaSuccess) return hbp_cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
        e = cudaGraphInstantiate(&slot->exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { slot->exec = nullptr; return hbp_cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
        slot->nodes = n;
        slot->stamp = ++m->graph_clock;
        HBP_CUDA(cudaGraphLaunch(slot->exec, ctx->stream));
        ctx->launches += n;
        return HBP_OK;
    }
    // first call with this key: eager (creates the plans), and remember the key
    s = issue_ops(ctx, m, crops, P, heatmaps, out_dtype, &n);
    if (s) return s;
    ctx->launches += n;
    if ((int)m->graphs.size() >= HrnetModel::kMaxGraphs) {
        size_t lru = 0;
        for (size_t i = 1; i < m->graphs.size(); ++i) if (m->graphs[i].stamp < m->graphs[lru].stamp) lru = i;
        if (m->graphs[lru].exec) { HBP_CUDA(cudaStreamSynchronize(ctx->stream)); cudaGraphExecDestroy(m->graphs[lru].exec); }
        m->graphs.erase(m->graphs.begin() + lru);
    }
    HrnetModel::HGraph g;
    g.P = P; g.dtype = out_dtype; g.engine = m->engine; g.in = crops; g.out = heatmaps; g.stamp = ++m->graph_clock;
    m->graphs.push_back(g);
    return HBP_OK;
}

// Parity hook: run the program eagerly (no graph) up to and including op `stop_after`, so that the tensors
// live at that point (e.g. all outputs of a stage module when `stop_after` is the module's last op) can be read
// with hrnet_debug_tensor before later ops reuse their buffers.  Same plans and kernels as the graph path.
int hrnet_forward_until(hbp_ctx* ctx, const __half* crops, int P, int stop_after) {
    HrnetModel* m = ctx->hrnet;
    if (!m) { hbp_set_error("no model loaded"); return HBP_ERR_STATE; }
    if (stop_after < 0 || stop_after >= (int)m->ops.size()) { hbp_set_error("bad op index"); return HBP_ERR_INVALID; }
    if (m->ops[stop_after].kind == OP_HEAD) { hbp_set_error("the head writes the caller's heatmaps: use hbp_hrnet_forward"); return HBP_ERR_INVALID; }
    int s = ensure_batch(ctx, m, P);
    if (s) return s;
    uint64_t n = 0;
    s = issue_ops(ctx, m, crops, P, nullptr, HBP_F16, &n, stop_after);
    if (s) return s;
    ctx->launches += n;
    HBP_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < kStreams - 1; ++i) if (m->side[i]) HBP_CUDA(cudaStreamSynchronize(m->side[i]));
    m->graph_P = P;          // hrnet_debug_tensor sizes its copy by the batch of the last forward
    return HBP_OK;
}

int hrnet_op_count(hbp_ctx* ctx) { return ctx->hrnet ? (int)ctx->hrnet->ops.size() : 0; }
const char* hrnet_op_name(hbp_ctx* ctx, int id) {
    HrnetModel* m = ctx->hrnet;
    if (!m || id < 0 || id >= (int)m->ops.size()) return nullptr;
    return m->ops[id].name.c_str();
}

int hrnet_debug_tensor(hbp_ctx* ctx, int id, void* out_host, size_t max_bytes, int* n, int* h, int* w, int* c) {
    HrnetModel* m = ctx->hrnet;
    if (!m || m->cap_P == 0) { hbp_set_error("no forward has run"); return HBP_ERR_STATE; }
    if (id < 0 || id >= (int)m->ops.size() || m->ops[id].out < 0) { hbp_set_error("bad op index"); return HBP_ERR_INVALID; }
    const HTensor& t = m->tensors[m->ops[id].out];
    const int P = m->graph_P;
    const size_t bytes = (size_t)P * t.c * t.h * t.w * sizeof(__half);
    if (n) *n = P; if (h) *h = t.h; if (w) *w = t.w; if (c) *c = t.c;
    if (!out_host) return HBP_OK;
    if (bytes > max_bytes) { hbp_set_error("buffer too small"); return HBP_ERR_OVERFLOW; }
    HBP_CUDA(cudaStreamSynchronize(ctx->stream));
    HBP_CUDA(cudaMemcpy(out_host, m->bufs[t.buf], bytes, cudaMemcpyDeviceToHost));
    return HBP_OK;
}

// One fused convolution on caller-provided device buffers (NHWC fp16), through either
// engine: the unit-test / bring-up entry for the conv kernels.
int hrnet_single_conv(hbp_ctx* ctx, int engine, const __half* in, int P, int H, int W, int Cin,
                      const __half* w, const float* bias, const __half* res, int Cout, int k, int stride,
                      int up, int relu, __half* out, int* used_engine, int iters, float* avg_ms) {
    HrnetModel m;
    HTensor ti; ti.c = Cin; ti.h = H; ti.w = W; ti.buf = 0;
    HTensor to; to.c = Cout; to.h = H / stride * up; to.w = W / stride * up; to.buf = 1;
    HTensor tr = to; tr.buf = 2;
    m.tensors = {ti, to, tr};
    m.bufs = {const_cast<__half*>(in), out, const_cast<__half*>(res)};
    m.d_weights = const_cast<__half*>(w);
    m.d_bias = const_cast<float*>(bias);
    HOp op;
    op.kind = OP_CONV; op.name = "single"; op.in = 0; op.out = 1; op.res = res ? 2 : -1;
    op.cin = Cin; op.cout = Cout; op.k = k; op.stride = stride; op.up = up; op.relu = relu;
    m.ops = {op};
    int status = HBP_OK;
    if (engine == 1 && umma_supported(m, op)) {
        UmmaPlan* plan = nullptr;
        status = umma_plan_create(ctx, m, 0, P, &plan);
        if (status == HBP_OK) {
            m.umma = {plan};
            m.ops[0].persist = 1;
            status = umma_launch(ctx, m, 0, plan, P, ctx->stream);
            if (iters > 0 && avg_ms && status == HBP_OK) {
                // timing (bring-up / microbenchmark): the launches are captured into one CUDA graph so
                // that host launch overhead (large __grid_constant__ parameters) does not pace the GPU
                cudaGraph_t g = nullptr;
                cudaGraphExec_t ge = nullptr;
                cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
                for (int i = 0; i < iters && status == HBP_OK; ++i) status = umma_launch(ctx, m, 0, plan, P, ctx->stream);
                cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
                if (ce == cudaSuccess && status == HBP_OK) ce = cudaGraphInstantiate(&ge, g, 0);
                if (ce == cudaSuccess && status == HBP_OK) {
                    cudaGraphLaunch(ge, ctx->stream);              // warm
                    cudaEventRecord(ctx->ev_start[7], ctx->stream);
                    cudaGraphLaunch(ge, ctx->stream);
                    cudaEventRecord(ctx->ev_stop[7], ctx->stream);
                    cudaEventSynchronize(ctx->ev_stop[7]);
                    cudaEventElapsedTime(avg_ms, ctx->ev_start[7], ctx->ev_stop[7]);
                    *avg_ms /= iters;
                    ctx->launches += 2 * iters;
                } else if (status == HBP_OK) {
                    status = hbp_cuda_fail(ce, "graph capture of the timing loop", __FILE__, __LINE__);
                }
                if (ge) cudaGraphExecDestroy(ge);
                if (g) cudaGraphDestroy(g);
            }
            // the tensor maps are kernel parameters (copied at launch): the plan can go
            umma_plan_destroy(plan);
        }
        if (used_engine) *used_engine = 1;
    } else {
        const int Ho = H / stride, Wo = W / stride;
        const size_t total = (size_t)P * Ho * Wo;
        dim3 grid((unsigned)((total + kTP - 1) / kTP), (Cout + kTC - 1) / kTC);
        conv_simt_kernel<<<grid, 256, 0, ctx->stream>>>(in, w, bias, res, out, P, H, W, Cin, Ho, Wo, Cout, k,
                                                        stride, up, relu);
        if (iters > 0 && avg_ms) {
            cudaEventRecord(ctx->ev_start[7], ctx->stream);
            for (int i = 0; i < iters; ++i)
                conv_simt_kernel<<<grid, 256, 0, ctx->stream>>>(in, w, bias, res, out, P, H, W, Cin, Ho, Wo, Cout, k,
                                                                stride, up, relu);
            cudaEventRecord(ctx->ev_stop[7], ctx->stream);
            cudaEventSynchronize(ctx->ev_stop[7]);
            cudaEventElapsedTime(avg_ms, ctx->ev_start[7], ctx->ev_stop[7]);
            *avg_ms /= iters;
            ctx->launches += iters;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) status = hbp_cuda_fail(e, "conv_simt_kernel", __FILE__, __LINE__);
        if (used_engine) *used_engine = 0;
    }
    m.d_weights = nullptr; m.d_bias = nullptr; m.bufs.clear();
    if (!m.groups.empty()) {
        cudaStreamSynchronize(ctx->stream);
        for (UmmaGroup* g : m.groups) if (g) umma_group_destroy(g);
        m.groups.clear();
    }
    ctx->launches++;
    return status;
}

// host-only description of the program (no context needed): one line per conv
//   name cin cout k stride w_off b_off out_h out_w up   (out_h/out_w before the fused upsample)
extern "C" int hbp_hrnet_describe(int width, int in_h, int in_w, char* buf, size_t buf_bytes,
                                          size_t* n_weights, size_t* n_biases, size_t* needed) {
    if ((width != 32 && width != 48) || in_h % 32 || in_w % 32 || in_h <= 0 || in_w <= 0) {
        hbp_set_error("hbp_hrnet_describe: bad architecture");
        return HBP_ERR_INVALID;
    }
    HrnetModel m;
    m.width = width; m.in_h = in_h; m.in_w = in_w;
    hrnet_build_program(m);
    std::string s;
    char line[256];
    for (const HOp& op : m.ops) {
        if (op.kind == OP_UPADD || op.kind == OP_GROUP || op.kind == OP_UPADD_GROUP || op.kind == OP_CHAIN) continue;            // no parameters
        int ho = 0, wo = 0;
        if (op.kind == OP_STEM1) { ho = m.in_h / 2; wo = m.in_w / 2; }
        else { ho = m.tensors[op.in].h / op.stride; wo = m.tensors[op.in].w / op.stride; }
        snprintf(line, sizeof(line), "%s %d %d %d %d %zu %zu %d %d %d\n", op.name.c_str(), op.cin_l, op.cout_l, op.k,
                 op.stride, op.w_off_l, op.b_off_l, ho, wo, op.up);
        s += line;
    }
    if (n_weights) *n_weights = m.n_weights_l;
    if (n_biases) *n_biases = m.n_biases_l;
    if (needed) *needed = s.size() + 1;
    if (buf && buf_bytes > 0) {
        const size_t k = std::min(buf_bytes - 1, s.size());
        memcpy(buf, s.data(), k);
        buf[k] = 0;
        if (k < s.size()) return HBP_ERR_OVERFLOW;
    }
    return HBP_OK;
}
