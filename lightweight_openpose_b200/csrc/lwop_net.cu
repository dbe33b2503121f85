// Handle, weight store, whole-network executor and the extern "C" entry points
// declared in include/lwop.h.
//
// The layer table below restates the graph of the reference
// src/lightweight_openpose.py:8-48 (built from src/net_factory.py:4-37) in
// creation order; lightweight_openpose_b200/graph.py holds the same table on
// the Python side (variable names, BN folding) and the packed-weight blob must
// match lwop_weights_num_floats() exactly.
#include <cstring>
#include <map>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "lwop_kernels.cuh"

using namespace lwop;

namespace {

enum LayerKind { L_CONV = 0, L_SEP = 1 };

struct LayerSpec {
    int kind, cin, cout, k, stride, dil, act;
};

// clang-format off
const LayerSpec kLayers[] = {
    // BackBone (lightweight_openpose.py:9-21)
    {L_CONV,   3,  32, 3, 2, 1, LWOP_ACT_RELU},
    {L_SEP,   32,  64, 1, 1, 1, LWOP_ACT_RELU},
    {L_SEP,   64, 128, 1, 2, 1, LWOP_ACT_RELU},
    {L_SEP,  128, 128, 1, 1, 1, LWOP_ACT_RELU},
    {L_SEP,  128, 256, 1, 2, 1, LWOP_ACT_RELU},
    {L_SEP,  256, 256, 1, 1, 1, LWOP_ACT_RELU},
    {L_SEP,  256, 512, 1, 1, 1, LWOP_ACT_RELU},
    {L_SEP,  512, 512, 1, 1, 2, LWOP_ACT_RELU},
    {L_SEP,  512, 512, 1, 1, 1, LWOP_ACT_RELU},
    {L_SEP,  512, 512, 1, 1, 1, LWOP_ACT_RELU},
    {L_SEP,  512, 512, 1, 1, 1, LWOP_ACT_RELU},
    {L_SEP,  512, 512, 1, 1, 1, LWOP_ACT_RELU},
    // Cpm (:22-27)
    {L_CONV, 512, 128, 1, 1, 1, LWOP_ACT_RELU},
    {L_SEP,  128, 128, 1, 1, 1, LWOP_ACT_ELU},
    {L_SEP,  128, 128, 1, 1, 1, LWOP_ACT_ELU},
    {L_SEP,  128, 128, 1, 1, 1, LWOP_ACT_ELU},
    {L_CONV, 128, 128, 3, 1, 1, LWOP_ACT_RELU},
    // InitialStage (:28-36)
    {L_CONV, 128, 128, 3, 1, 1, LWOP_ACT_RELU},
    {L_CONV, 128, 128, 3, 1, 1, LWOP_ACT_RELU},
    {L_CONV, 128, 128, 3, 1, 1, LWOP_ACT_RELU},
    {L_CONV, 128, 512, 1, 1, 1, LWOP_ACT_RELU},
    {L_CONV, 512,  14, 1, 1, 1, LWOP_ACT_NONE},
    {L_CONV, 128, 512, 1, 1, 1, LWOP_ACT_RELU},
    {L_CONV, 512,  26, 1, 1, 1, LWOP_ACT_NONE},
    // RefinementStage (:37-46), 5 x RefinementStageBlock (net_factory.py:33-37)
    {L_CONV,  40, 128, 1, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 2, LWOP_ACT_RELU},
    {L_CONV, 128, 128, 1, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 2, LWOP_ACT_RELU},
    {L_CONV, 128, 128, 1, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 2, LWOP_ACT_RELU},
    {L_CONV, 128, 128, 1, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 2, LWOP_ACT_RELU},
    {L_CONV, 128, 128, 1, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 1, LWOP_ACT_RELU}, {L_CONV, 128, 128, 3, 1, 2, LWOP_ACT_RELU},
    {L_CONV, 128, 128, 1, 1, 1, LWOP_ACT_RELU},
    {L_CONV, 128,  14, 1, 1, 1, LWOP_ACT_NONE},
    {L_CONV, 128, 128, 1, 1, 1, LWOP_ACT_RELU},
    {L_CONV, 128,  26, 1, 1, 1, LWOP_ACT_NONE},
};
// clang-format on
constexpr int kNumLayers = sizeof(kLayers) / sizeof(kLayers[0]);
static_assert(kNumLayers == 43, "layer table out of sync with graph.py");

enum BufId { B_IN = 0, B_A, B_B, B_D, B_ALIGN, B_T0, B_T1, B_CAT, B_INIT, B_HEAT, B_PAF, B_H1, B_P1, NUM_BUF };

struct Step {
    int layer;
    bool is_dw;        // depthwise half of a separable layer
    int in, out, res;  // BufId (res = -1: none)
    int H, W;          // input spatial size
    int cin_pad;       // input channels / pitch
    int ldo, col0;     // output pitch / first column
    int out_f32;       // output is a caller-owned fp32 map (final heads)
    int out2;          // BufId of optional fp32 copy (stage-1 heads), -1 none
};

int pad_cout(int c) { return c <= 16 ? 16 : (c <= 32 ? 32 : (c <= 64 ? 64 : ((c + 127) / 128) * 128)); }
int pad_cin(int c) { return c == 40 ? 64 : c; }

struct PlanSet {
    std::vector<GemmPlan *> plans;   // per step (nullptr for non-GEMM steps)
    std::vector<DwPlan *> dw;        // per step (nullptr for non-depthwise steps)
    std::vector<SepPlan *> sep;      // per step: fused depthwise -> pointwise plan, stored at the depthwise step
    std::vector<HeadPairPlan *> heads;   // per step: fused head pair (this step + the next three)
    int launches = 0;                // kernels one forward launches with these plans
};

struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    const void *in = nullptr;
    void *heat = nullptr, *paf = nullptr, *h1 = nullptr, *p1 = nullptr;
    int in_u8 = 0;
    int launches = 0;
    // decode chain captured behind the forward (lwop_forward_decode): its tables and thresholds
    void *dec_counts = nullptr;
    float thre1 = 0.f;
    double thre2 = 0.0;
};

}  // namespace

struct lwop_handle {
    double host_us[6] = {0, 0, 0, 0, 0, 0};   // LWOP_TRACE_HOST=1: host time inside lwop_infer_host_begin by segment
    long long host_calls = 0;
    int device = 0, max_batch = 0, H = 0, W = 0;
    unsigned flags = 0;
    int h2 = 0, w2 = 0, h4 = 0, w4 = 0, h8 = 0, w8 = 0;
    std::string err;
    long long launches = 0;

    // weights
    bool has_weights = false;
    float *blob = nullptr;                       // fp32 folded weights, blob layout
    size_t off_dw[kNumLayers] = {}, off_w[kNumLayers] = {}, off_b[kNumLayers] = {};
    void *w_bf16[kNumLayers] = {};               // [Cout_pad][taps][Cin_pad] bf16
    float *b_pad[kNumLayers] = {};               // [Cout_pad] fp32
    float *w_f32pad[kNumLayers] = {};            // fp32 weights with padded Cin (only layer 24: 40 -> 64)
    Conv1Weights c1 = {};                        // first layer: folded taps + bias, a by-value kernel parameter

    // program + arenas
    std::vector<Step> steps;
    size_t buf_elems[NUM_BUF] = {};              // elements per image
    void *arena_bf16[NUM_BUF] = {};
    void *arena_f32[NUM_BUF] = {};
    std::map<int, PlanSet> plan_sets;            // by batch
    std::map<long long, std::vector<GraphEntry>> graphs;   // by batch * 8 + mode; entries differ in I/O pointers

    // decode scratch
    float *peaks = nullptr;
    int *peak_counts = nullptr;
    double *group_work = nullptr;
    int *pack_offsets = nullptr;
    int *draw_prio = nullptr;                    // canvas rasteriser: priority map [pixels], -1 between calls
    size_t draw_prio_cap = 0;
    double *pk_joints = nullptr, *pk_limbs = nullptr, *pk_persons = nullptr;
    float *pk_keypoints = nullptr;
    size_t pk_joints_cap = 0, pk_limbs_cap = 0, pk_persons_cap = 0;
    int *h_offsets = nullptr;                    // pinned
    // infer_host pipeline: two slots (input staging + maps), a copy stream and events
    int chunk = 0, pipe_chunk = 0;
    unsigned long long slot_seq = 0;
    bool pipe_ready = false;
    void *slot_in[2] = {};
    float *slot_heat[2] = {}, *slot_paf[2] = {};
    cudaEvent_t ev_h2d[2] = {}, ev_fwd[2] = {}, ev_done[2] = {}, ev_start = nullptr, ev_join = nullptr;
    cudaStream_t copy_stream = nullptr, decode_stream = nullptr;
    bool copy_stream_borrowed = false;           // set through the "copy_stream" option: several handles share one upload queue
    lwop_decode_tables own_tables = {};
    int own_tables_batch = 0;
    // batches in flight through lwop_infer_host_begin / _end: each lane owns its decode tables and packed buffers
    struct HostLane {
        lwop_decode_tables tables = {};
        int *pack_offsets = nullptr;
        double *pk_joints = nullptr, *pk_limbs = nullptr, *pk_persons = nullptr;
        float *pk_keypoints = nullptr;
        size_t cap_j = 0, cap_l = 0, cap_p = 0;
        cudaEvent_t ev_ready = nullptr;
        bool busy = false;
        int batch = 0;
        lwop_packed_tables out = {};
    } lanes[LWOP_MAX_INFLIGHT];
};

namespace {

std::string g_last_error;
std::mutex g_err_mutex;

int fail(lwop_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_last_error = msg;
    return code;
}

#define CU_OK(h, expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return fail(h, LWOP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
    } while (0)

size_t blob_floats() {
    size_t n = 0;
    for (int i = 0; i < kNumLayers; ++i) {
        const LayerSpec &l = kLayers[i];
        if (l.kind == L_SEP) n += 9 * (size_t)l.cin;
        n += (size_t)l.cout * l.k * l.k * l.cin + l.cout;
    }
    return n;
}

void build_program(lwop_handle *h) {
    auto &S = h->steps;
    S.clear();
    for (int i = 0; i < NUM_BUF; ++i) h->buf_elems[i] = 0;
    auto need = [&](int buf, size_t elems) {
        if (elems > h->buf_elems[buf]) h->buf_elems[buf] = elems;
    };
    int H = h->H, W = h->W;
    auto add = [&](int layer, bool is_dw, int in, int out, int res, int inH, int inW, int cin_pad, int ldo, int col0,
                   int out_f32 = 0, int out2 = -1) {
        Step s{layer, is_dw, in, out, res, inH, inW, cin_pad, ldo, col0, out_f32, out2};
        S.push_back(s);
    };
    // layer 0: conv1
    add(0, false, B_IN, B_A, -1, H, W, 3, 32, 0);
    H = h->h2; W = h->w2;
    need(B_A, (size_t)H * W * 32);
    int cur = B_A;
    for (int li = 1; li <= 11; ++li) {
        const LayerSpec &l = kLayers[li];
        const int oH = (H + l.stride - 1) / l.stride, oW = (W + l.stride - 1) / l.stride;
        add(li, true, cur, B_D, -1, H, W, l.cin, l.cin, 0);
        need(B_D, (size_t)oH * oW * l.cin);
        const int nxt = cur == B_A ? B_B : B_A;
        add(li, false, B_D, nxt, -1, oH, oW, l.cin, l.cout, 0);
        need(nxt, (size_t)oH * oW * l.cout);
        cur = nxt;
        H = oH; W = oW;
    }
    const size_t px = (size_t)H * W;      // stride-8 pixels
    for (int b : {B_ALIGN, B_T0, B_T1, B_INIT}) need(b, px * 128);
    need(B_CAT, px * 64);
    need(B_A, px * 512);
    need(B_B, px * 512);
    need(B_D, px * 128);
    // Cpm
    add(12, false, cur, B_ALIGN, -1, H, W, 512, 128, 0);
    add(13, true, B_ALIGN, B_D, -1, H, W, 128, 128, 0);
    add(13, false, B_D, B_T0, -1, H, W, 128, 128, 0);
    add(14, true, B_T0, B_D, -1, H, W, 128, 128, 0);
    add(14, false, B_D, B_T1, -1, H, W, 128, 128, 0);
    add(15, true, B_T1, B_D, -1, H, W, 128, 128, 0);
    add(15, false, B_D, B_T0, B_ALIGN, H, W, 128, 128, 0);          // tf.add(net_align, net_truck) (:27)
    add(16, false, B_T0, B_T1, -1, H, W, 128, 128, 0);
    // InitialStage
    add(17, false, B_T1, B_T0, -1, H, W, 128, 128, 0);
    add(18, false, B_T0, B_T1, -1, H, W, 128, 128, 0);
    add(19, false, B_T1, B_T0, -1, H, W, 128, 128, 0);
    add(20, false, B_T0, B_A, -1, H, W, 128, 512, 0);
    add(21, false, B_A, B_CAT, -1, H, W, 512, 64, 0, 0, B_H1);      // heatmaps1 -> concat cols 0..13 (:36)
    add(22, false, B_T0, B_A, -1, H, W, 128, 512, 0);
    add(23, false, B_A, B_CAT, -1, H, W, 512, 64, 14, 0, B_P1);     // pafs1 -> concat cols 14..39
    // RefinementStage
    int in = B_CAT;
    for (int b = 0; b < 5; ++b) {
        const int l0 = 24 + 3 * b;
        add(l0, false, in, B_INIT, -1, H, W, b == 0 ? 64 : 128, 128, 0);
        add(l0 + 1, false, B_INIT, B_T1, -1, H, W, 128, 128, 0);
        add(l0 + 2, false, B_T1, B_T0, B_INIT, H, W, 128, 128, 0);  // tf.add(net_initial, net_truck) (net_factory.py:37)
        in = B_T0;
    }
    add(39, false, B_T0, B_T1, -1, H, W, 128, 128, 0);
    add(40, false, B_T1, B_HEAT, -1, H, W, 128, 14, 0, 1);
    add(41, false, B_T0, B_T1, -1, H, W, 128, 128, 0);
    add(42, false, B_T1, B_PAF, -1, H, W, 128, 26, 0, 1);
}

int ensure_arena(lwop_handle *h, int mode) {
    void **arena = mode == LWOP_MODE_FP32_CHECK ? h->arena_f32 : h->arena_bf16;
    const size_t esz = mode == LWOP_MODE_FP32_CHECK ? 4 : 2;
    for (int b = B_A; b <= B_INIT; ++b) {
        if (arena[b] || h->buf_elems[b] == 0) continue;
        const size_t bytes = h->buf_elems[b] * esz * (size_t)h->max_batch + 4096;   // slack: TMA tail tiles never fault
        CU_OK(h, cudaMalloc(&arena[b], bytes));
        CU_OK(h, cudaMemset(arena[b], 0, bytes));
    }
    return LWOP_OK;
}

void destroy_plans(lwop_handle *h) {
    for (auto &kv : h->plan_sets) {
        for (GemmPlan *p : kv.second.plans)
            if (p) gemm_plan_destroy(p);
        for (DwPlan *p : kv.second.dw)
            if (p) dw_plan_destroy(p);
        for (SepPlan *p : kv.second.sep)
            if (p) sep_plan_destroy(p);
        for (HeadPairPlan *p : kv.second.heads)
            if (p) head_pair_plan_destroy(p);
    }
    h->plan_sets.clear();
    for (auto &kv : h->graphs)
        for (auto &g : kv.second)
            if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

void destroy_graphs_only(lwop_handle *h) {
    for (auto &kv : h->graphs)
        for (auto &g : kv.second)
            if (g.exec) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

bool step_uses_gemm(const Step &s) { return !s.is_dw && s.layer != 0; }

// A depthwise step and the pointwise step that follows it run as ONE fused kernel (lwop_sep.cu) when the
// shape is supported; LWOP_FLAG_NO_FUSE_SEP / LWOP_NO_FUSE_SEP=1 keep the two-kernel path (A/B timing).
bool step_fuses_with_next(const lwop_handle *h, size_t i) {
    static const bool env_off = getenv("LWOP_NO_FUSE_SEP") != nullptr;
    if (env_off || (h->flags & LWOP_FLAG_NO_FUSE_SEP)) return false;
    if (i + 1 >= h->steps.size()) return false;
    const Step &s = h->steps[i], &n = h->steps[i + 1];
    if (!s.is_dw || n.is_dw || n.layer != s.layer || n.out_f32 || n.out2 >= 0) return false;
    const LayerSpec &l = kLayers[s.layer];
    // LWOP_FUSE_ONLY=<layer>: debugging aid, fuse only this layer (bisecting a kernel problem by layer)
    static const char *only = getenv("LWOP_FUSE_ONLY");
    if (only && atoi(only) != s.layer) return false;
    return sep_plan_supported(l.cin, pad_cout(l.cout), l.stride, l.dil);
}

// Steps i..i+3 = conv(128->K2)+ReLU, conv(K2->14), conv(128->K2)+ReLU, conv(K2->26) on the same input run as ONE
// back-to-back GEMM kernel (lwop_heads.cu); LWOP_FLAG_NO_FUSE_HEADS / LWOP_NO_FUSE_HEADS=1 keep four launches.
bool step_starts_head_pair(const lwop_handle *h, size_t i) {
    static const bool env_off = getenv("LWOP_NO_FUSE_HEADS") != nullptr;
    if (env_off || (h->flags & LWOP_FLAG_NO_FUSE_HEADS)) return false;
    if (i + 3 >= h->steps.size()) return false;
    const Step *s = &h->steps[i];
    if (s[0].layer != 20 && s[0].layer != 39) return false;
    for (int k = 0; k < 4; ++k)
        if (s[k].is_dw || s[k].layer != s[0].layer + k) return false;
    const LayerSpec &l0 = kLayers[s[0].layer], &l1 = kLayers[s[1].layer], &l2 = kLayers[s[2].layer], &l3 = kLayers[s[3].layer];
    if (l0.k != 1 || l1.k != 1 || l2.k != 1 || l3.k != 1 || l0.cout != l2.cout || l1.cout != 14 || l3.cout != 26) return false;
    if (s[0].in != s[2].in) return false;
    return head_pair_supported(l0.cin, l0.cout);
}
int step_absorbed_by_head_pair(const lwop_handle *h, size_t i) {
    for (size_t k = 1; k <= 3 && k <= i; ++k)
        if (step_starts_head_pair(h, i - k)) return 1;
    return 0;
}

int get_plans(lwop_handle *h, int batch, PlanSet **out) {
    auto it = h->plan_sets.find(batch);
    if (it != h->plan_sets.end()) {
        *out = &it->second;
        return LWOP_OK;
    }
    PlanSet ps;
    ps.plans.assign(h->steps.size(), nullptr);
    ps.dw.assign(h->steps.size(), nullptr);
    ps.sep.assign(h->steps.size(), nullptr);
    ps.heads.assign(h->steps.size(), nullptr);
    char err[256];
    auto cleanup = [&]() {
        for (GemmPlan *p : ps.plans)
            if (p) gemm_plan_destroy(p);
        for (DwPlan *p : ps.dw)
            if (p) dw_plan_destroy(p);
        for (SepPlan *p : ps.sep)
            if (p) sep_plan_destroy(p);
        for (HeadPairPlan *p : ps.heads)
            if (p) head_pair_plan_destroy(p);
    };
    for (size_t i = 0; i < h->steps.size(); ++i) {
        const Step &s = h->steps[i];
        const LayerSpec &l = kLayers[s.layer];
        if (step_starts_head_pair(h, i)) {
            const Step &s1 = h->steps[i + 1], &s3 = h->steps[i + 3];
            HeadPairDesc d{};
            d.in = h->arena_bf16[s.in]; d.lda = s.cin_pad; d.Cin = l.cin;
            d.M = (long long)batch * s.H * s.W;
            d.hidden = l.cout;
            d.w1a = h->w_bf16[s.layer]; d.b1a = h->b_pad[s.layer];
            d.w2a = h->w_bf16[s.layer + 1]; d.b2a = h->b_pad[s.layer + 1];
            d.w1b = h->w_bf16[s.layer + 2]; d.b1b = h->b_pad[s.layer + 2];
            d.w2b = h->w_bf16[s.layer + 3]; d.b2b = h->b_pad[s.layer + 3];
            if (!s1.out_f32) {            // initial stage: heads land in the bf16 concat buffer
                d.cat = h->arena_bf16[s1.out]; d.ldc = s1.ldo; d.cat_col_a = s1.col0; d.cat_col_b = s3.col0;
            }
            ps.heads[i] = head_pair_plan_create(d, err, sizeof(err));
            if (!ps.heads[i]) {
                cleanup();
                return fail(h, LWOP_ERR_UNSUPPORTED, std::string("layer ") + std::to_string(s.layer) + " (fused heads): " + err);
            }
            i += 3;
            continue;
        }
        if (step_fuses_with_next(h, i)) {
            const Step &n = h->steps[i + 1];
            SepConvDesc d{};
            d.in = h->arena_bf16[s.in];
            d.dw_w = h->blob + h->off_dw[s.layer];
            d.w_packed = h->w_bf16[s.layer];
            d.bias = h->b_pad[s.layer];
            d.residual = n.res >= 0 ? h->arena_bf16[n.res] : nullptr;
            d.out = h->arena_bf16[n.out];
            d.B = batch; d.H = s.H; d.W = s.W; d.C = l.cin; d.Cout_pad = pad_cout(l.cout); d.dil = l.dil; d.act = l.act;
            d.ldo = n.ldo; d.out_col0 = n.col0; d.ldr = 128;
            ps.sep[i] = sep_plan_create(d, err, sizeof(err));
            if (!ps.sep[i]) {
                cleanup();
                return fail(h, LWOP_ERR_UNSUPPORTED, std::string("layer ") + std::to_string(s.layer) + " (fused separable): " + err);
            }
            ++i;           // the pointwise step is part of the fused kernel
            continue;
        }
        if (s.is_dw) {
            DwArgs a{};
            a.in = h->arena_bf16[s.in]; a.out = h->arena_bf16[s.out]; a.dtype = LWOP_DT_BF16;
            a.w = h->blob + h->off_dw[s.layer];
            a.B = batch; a.H = s.H; a.W = s.W; a.C = l.cin; a.stride = l.stride; a.dil = l.dil;
            ps.dw[i] = dw_plan_create(a, err, sizeof(err));
            if (!ps.dw[i]) {
                cleanup();
                return fail(h, LWOP_ERR_UNSUPPORTED, std::string("layer ") + std::to_string(s.layer) + " (dw): " + err);
            }
            continue;
        }
        if (!step_uses_gemm(s)) continue;
        GemmConvDesc d{};
        d.in = h->arena_bf16[s.in];
        d.w_packed = h->w_bf16[s.layer];
        d.bias = h->b_pad[s.layer];
        d.residual = s.res >= 0 ? h->arena_bf16[s.res] : nullptr;
        d.out = s.out_f32 ? nullptr : h->arena_bf16[s.out];
        d.out2 = nullptr;
        d.out_dtype = s.out_f32 ? LWOP_DT_F32 : LWOP_DT_BF16;
        d.lda = s.cin_pad;
        d.B = batch;
        d.H = s.H;
        d.W = s.W;
        d.Cin_pad = s.cin_pad;
        d.Cout = l.cout;
        d.Cout_pad = pad_cout(l.cout);
        d.ksize = l.k;
        d.dil = l.dil;
        d.act = l.act;
        d.ldo = s.ldo;
        d.out_col0 = s.col0;
        d.ldr = 128;
        ps.plans[i] = gemm_plan_create(d, err, sizeof(err));
        if (!ps.plans[i]) {
            cleanup();
            return fail(h, LWOP_ERR_UNSUPPORTED, std::string("layer ") + std::to_string(s.layer) + ": " + err);
        }
    }
    auto res = h->plan_sets.emplace(batch, std::move(ps));
    *out = &res.first->second;
    return LWOP_OK;
}

struct FwdIO {
    const void *in;
    int in_u8;
    float *heat, *paf, *h1, *p1;
    const DecodeArgs *dec = nullptr;     // decode chain to run behind the forward (same stream / same graph), or nullptr
};

// Issue every kernel of one forward pass on `s`.  With `ev` (2 events per step)
// every step is bracketed by event records for per-kernel timing.
int run_forward(lwop_handle *h, int batch, int mode, const FwdIO &io, cudaStream_t s, cudaEvent_t *ev = nullptr) {
    const bool f32 = mode == LWOP_MODE_FP32_CHECK;
    void **arena = f32 ? h->arena_f32 : h->arena_bf16;
    const int dt = f32 ? LWOP_DT_F32 : LWOP_DT_BF16;
    PlanSet *ps = nullptr;
    if (mode == LWOP_MODE_BF16) {
        int rc = get_plans(h, batch, &ps);
        if (rc != LWOP_OK) return rc;
    }
    const float *in_f32 = (const float *)io.in;
    if (io.in_u8 && mode != LWOP_MODE_BF16 && mode != LWOP_MODE_BF16_DIRECT) {
        // fp32 check path: normalise into a staging buffer first
        const long long n = (long long)batch * h->H * h->W * 3;
        if (!h->arena_f32[B_IN]) {
            CU_OK(h, cudaMalloc(&h->arena_f32[B_IN], (size_t)h->max_batch * h->H * h->W * 3 * 4));
        }
        CU_OK(h, launch_u8_to_f32((const uint8_t *)io.in, (float *)h->arena_f32[B_IN], n, s));
        h->launches++;
        in_f32 = (const float *)h->arena_f32[B_IN];
    }
    struct EvGuard {
        cudaEvent_t *ev; size_t i; cudaStream_t s;
        EvGuard(cudaEvent_t *e, size_t idx, cudaStream_t st) : ev(e), i(idx), s(st) { if (ev) cudaEventRecord(ev[2 * i], s); }
        ~EvGuard() { if (ev) cudaEventRecord(ev[2 * i + 1], s); }
    };
    for (size_t i = 0; i < h->steps.size(); ++i) {
        const Step &st = h->steps[i];
        const LayerSpec &l = kLayers[st.layer];
        EvGuard guard(ev, i, s);
        void *out = st.out == B_HEAT ? (void *)io.heat : st.out == B_PAF ? (void *)io.paf : arena[st.out];
        const void *in = st.in == B_IN ? (const void *)in_f32 : arena[st.in];
        float *out2 = st.out2 == B_H1 ? io.h1 : st.out2 == B_P1 ? io.p1 : nullptr;
        if (st.layer == 0) {
            if (!f32) {
                CU_OK(h, launch_conv1_bf16(h->c1, io.in_u8 ? io.in : (const void *)in_f32, io.in_u8, out, batch, st.H, st.W, s));
            } else {
                ConvArgs a{};
                a.in = in_f32; a.in_dtype = LWOP_DT_F32; a.w = h->blob + h->off_w[0]; a.bias = h->blob + h->off_b[0];
                a.out = out; a.out_dtype = LWOP_DT_F32;
                a.B = batch; a.H = st.H; a.W = st.W; a.Cin = 3; a.Cout = 32; a.ksize = 3; a.stride = 2; a.dil = 1;
                a.act = LWOP_ACT_RELU;
                CU_OK(h, launch_conv_direct(a, s));
            }
            h->launches++;
            continue;
        }
        if (mode == LWOP_MODE_BF16 && ps->heads[i]) {
            const Step &s1 = h->steps[i + 1];
            float *oa = s1.out_f32 ? io.heat : io.h1, *ob = s1.out_f32 ? io.paf : io.p1;
            CU_OK(h, head_pair_plan_launch(ps->heads[i], s, oa, ob));
            h->launches++;
            continue;
        }
        if (mode == LWOP_MODE_BF16 && step_absorbed_by_head_pair(h, i)) continue;
        if (mode == LWOP_MODE_BF16 && ps->sep[i]) {
            CU_OK(h, sep_plan_launch(ps->sep[i], s));
            h->launches++;
            continue;
        }
        if (mode == LWOP_MODE_BF16 && i > 0 && ps->sep[i - 1]) continue;   // pointwise half of the fused kernel above
        if (st.is_dw) {
            DwArgs a{};
            a.in = in; a.out = out; a.dtype = dt; a.w = h->blob + h->off_dw[st.layer];
            a.B = batch; a.H = st.H; a.W = st.W; a.C = l.cin; a.stride = l.stride; a.dil = l.dil;
            if (f32) CU_OK(h, launch_depthwise_simple(a, s));
            else if (mode == LWOP_MODE_BF16) CU_OK(h, dw_plan_launch(ps->dw[i], s));
            else CU_OK(h, launch_depthwise_bf16(a, s));
            h->launches++;
            continue;
        }
        if (mode == LWOP_MODE_BF16) {
            CU_OK(h, gemm_plan_launch(ps->plans[i], s, st.out_f32 ? out : nullptr, out2));
            h->launches++;
            continue;
        }
        // direct (CUDA-core) conv: fp32 check path or bf16-storage debugging path
        ConvArgs a{};
        a.in = in; a.in_dtype = dt;
        a.w = (st.cin_pad != l.cin && h->w_f32pad[st.layer]) ? h->w_f32pad[st.layer] : h->blob + h->off_w[st.layer];
        a.bias = h->blob + h->off_b[st.layer];
        a.residual = st.res >= 0 ? arena[st.res] : nullptr;
        a.out = out; a.out_dtype = st.out_f32 ? LWOP_DT_F32 : dt;
        a.B = batch; a.H = st.H; a.W = st.W; a.Cin = st.cin_pad; a.Cout = l.cout; a.ksize = l.k; a.stride = 1;
        a.dil = l.dil; a.act = l.act; a.ldi = st.cin_pad; a.ldo = st.ldo; a.out_col0 = st.col0; a.ldr = 128;
        CU_OK(h, launch_conv_direct(a, s));
        h->launches++;
        if (out2) {   // stage-1 heads also wanted as fp32 maps
            a.out = out2; a.out_dtype = LWOP_DT_F32; a.ldo = l.cout; a.out_col0 = 0;
            CU_OK(h, launch_conv_direct(a, s));
            h->launches++;
        }
    }
    if (io.dec) {       // the decode chain follows the last head kernel with programmatic dependent launches
        DecodeArgs a = *io.dec;
        a.heat = io.heat; a.paf = io.paf; a.B = batch; a.h = h->h8; a.w = h->w8; a.H = h->H; a.W = h->W;
        CU_OK(h, launch_decode(a, s, &h->launches));
    }
    return LWOP_OK;
}

int forward_impl(lwop_handle *h, const FwdIO &io, int batch, int mode, cudaStream_t s) {
    if (!h) return fail(nullptr, LWOP_ERR_INVALID, "null handle");
    if (!h->has_weights) return fail(h, LWOP_ERR_NO_WEIGHTS, "lwop_forward before lwop_load_weights");
    if (batch < 1 || batch > h->max_batch) return fail(h, LWOP_ERR_INVALID, "batch out of range");
    if (mode != LWOP_MODE_BF16 && mode != LWOP_MODE_FP32_CHECK && mode != LWOP_MODE_BF16_DIRECT)
        return fail(h, LWOP_ERR_INVALID, "unknown mode");
    if (!io.in || !io.heat || !io.paf) return fail(h, LWOP_ERR_INVALID, "null tensor pointer");
    CU_OK(h, cudaSetDevice(h->device));
    int rc = ensure_arena(h, mode);
    if (rc != LWOP_OK) return rc;
    const bool use_graph = !(h->flags & LWOP_FLAG_NO_GRAPH) && mode == LWOP_MODE_BF16;
    if (!use_graph) return run_forward(h, batch, mode, io, s);

    const long long key = (long long)batch * 8 + mode;
    std::vector<GraphEntry> &cache = h->graphs[key];
    GraphEntry *gp = nullptr;
    for (auto &c : cache)
        if (c.exec && c.in == io.in && c.heat == io.heat && c.paf == io.paf && c.h1 == io.h1 && c.p1 == io.p1 &&
            c.in_u8 == io.in_u8 && c.dec_counts == (io.dec ? (void *)io.dec->t.counts : nullptr) &&
            (!io.dec || (c.thre1 == io.dec->thre1 && c.thre2 == io.dec->thre2))) {
            gp = &c;
            break;
        }
    if (!gp) {
        if (cache.size() >= 32) {           // callers cycling through many buffers: drop the oldest graph
            if (cache.front().exec) cudaGraphExecDestroy(cache.front().exec);
            cache.erase(cache.begin());
        }
        // One eager pass first: creates the plans (tensor-map encoding calls the driver), sets the
        // per-kernel shared-memory attributes and surfaces launch errors outside of stream capture.
        const long long eager_before = h->launches;
        rc = run_forward(h, batch, mode, io, s);
        if (rc != LWOP_OK) return rc;
        const int launches_per_forward = (int)(h->launches - eager_before);
        CU_OK(h, cudaStreamSynchronize(s));
        cudaStream_t cs;
        CU_OK(h, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            const long long before = h->launches;
            rc = run_forward(h, batch, mode, io, cs);
            h->launches = before;                      // counted again at each replay
            e = cudaStreamEndCapture(cs, &graph);
        }
        if (e != cudaSuccess || rc != LWOP_OK || !graph) {
            cudaStreamDestroy(cs);
            if (graph) cudaGraphDestroy(graph);
            if (rc != LWOP_OK) return rc;
            return fail(h, LWOP_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
        }
        GraphEntry ne;
        e = cudaGraphInstantiate(&ne.exec, graph, 0);
        cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
        if (e != cudaSuccess) return fail(h, LWOP_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
        ne.in = io.in; ne.heat = io.heat; ne.paf = io.paf; ne.h1 = io.h1; ne.p1 = io.p1; ne.in_u8 = io.in_u8;
        if (io.dec) { ne.dec_counts = io.dec->t.counts; ne.thre1 = io.dec->thre1; ne.thre2 = io.dec->thre2; }
        ne.launches = launches_per_forward;
        cache.push_back(ne);
        gp = &cache.back();
    }
    GraphEntry &g = *gp;
    CU_OK(h, cudaGraphLaunch(g.exec, s));
    h->launches += g.launches;
    return LWOP_OK;
}

int ensure_decode_scratch(lwop_handle *h) {
    if (h->peaks) return LWOP_OK;
    CU_OK(h, cudaMalloc(&h->peaks, decode_scratch_peaks_bytes(h->max_batch)));
    CU_OK(h, cudaMalloc(&h->peak_counts, (size_t)h->max_batch * 16 * sizeof(int)));
    CU_OK(h, cudaMalloc(&h->group_work, decode_scratch_work_bytes(h->max_batch)));
    CU_OK(h, cudaMalloc(&h->pack_offsets, ((size_t)h->max_batch + 1) * 4 * sizeof(int)));
    CU_OK(h, cudaMallocHost(&h->h_offsets, ((size_t)h->max_batch + 1) * 4 * sizeof(int)));
    return LWOP_OK;
}

int ensure_pack_buffers(lwop_handle *h, size_t joints, size_t limbs, size_t persons) {
    if (joints > h->pk_joints_cap) {
        if (h->pk_joints) cudaFree(h->pk_joints);
        h->pk_joints_cap = joints + joints / 2 + 1024;
        CU_OK(h, cudaMalloc(&h->pk_joints, h->pk_joints_cap * 6 * sizeof(double)));
    }
    if (limbs > h->pk_limbs_cap) {
        if (h->pk_limbs) cudaFree(h->pk_limbs);
        h->pk_limbs_cap = limbs + limbs / 2 + 1024;
        CU_OK(h, cudaMalloc(&h->pk_limbs, h->pk_limbs_cap * 5 * sizeof(double)));
    }
    if (persons > h->pk_persons_cap) {
        if (h->pk_persons) cudaFree(h->pk_persons);
        if (h->pk_keypoints) cudaFree(h->pk_keypoints);
        h->pk_persons_cap = persons + persons / 2 + 256;
        CU_OK(h, cudaMalloc(&h->pk_persons, h->pk_persons_cap * 16 * sizeof(double)));
        CU_OK(h, cudaMalloc(&h->pk_keypoints, h->pk_persons_cap * 42 * sizeof(float)));
    }
    return LWOP_OK;
}

int ensure_own_tables(lwop_handle *h) {
    if (h->own_tables.counts) return LWOP_OK;
    const size_t B = h->max_batch;
    CU_OK(h, cudaMalloc(&h->own_tables.counts, B * LWOP_COUNTS * sizeof(int32_t)));
    CU_OK(h, cudaMalloc(&h->own_tables.joints, B * LWOP_MAX_JOINTS * 6 * sizeof(double)));
    CU_OK(h, cudaMalloc(&h->own_tables.limbs, B * LWOP_NUM_LIMBS * LWOP_MAX_PEAKS * 5 * sizeof(double)));
    CU_OK(h, cudaMalloc(&h->own_tables.persons, B * LWOP_MAX_PERSONS * 16 * sizeof(double)));
    CU_OK(h, cudaMalloc(&h->own_tables.keypoints, B * LWOP_MAX_PERSONS * 42 * sizeof(float)));
    return LWOP_OK;
}

uint32_t g_crc_table[8][256];
std::once_flag g_crc_once;
void crc_init() {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
        g_crc_table[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xff];
}

}  // namespace

extern "C" {

int lwop_abi_version(void) { return LWOP_ABI_VERSION; }

const char *lwop_error_string(int code) {
    switch (code) {
        case LWOP_OK: return "ok";
        case LWOP_ERR_INVALID: return "invalid argument";
        case LWOP_ERR_CUDA: return "CUDA failure";
        case LWOP_ERR_NO_WEIGHTS: return "weights not loaded";
        case LWOP_ERR_OVERFLOW: return "decode table overflow";
        case LWOP_ERR_UNSUPPORTED: return "unsupported shape or mode";
        case LWOP_ERR_NO_DEVICE: return "no usable CUDA device";
    }
    return "unknown error";
}

const char *lwop_last_error(const lwop_handle *h) {
    if (h) return h->err.c_str();
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_err_mutex);
    copy = g_last_error;
    return copy.c_str();
}

int lwop_same_out(int in, int stride) { return (in + stride - 1) / stride; }

size_t lwop_weights_num_floats(void) { return blob_floats(); }

int lwop_create(int device, int max_batch, int height, int width, unsigned flags, lwop_handle **out) {
    if (!out) return fail(nullptr, LWOP_ERR_INVALID, "out is null");
    *out = nullptr;
    if (max_batch < 1 || height < 16 || width < 16) return fail(nullptr, LWOP_ERR_INVALID, "bad geometry");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, LWOP_ERR_NO_DEVICE, "no CUDA device visible");
    if (device < 0 || device >= ndev) return fail(nullptr, LWOP_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
        return fail(nullptr, LWOP_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, LWOP_ERR_NO_DEVICE,
                    std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                        "; this library is built for sm_100a only");
    lwop_handle *h = new lwop_handle();
    h->device = device;
    h->max_batch = max_batch;
    h->H = height;
    h->W = width;
    h->flags = flags;
    h->h2 = (height + 1) / 2; h->w2 = (width + 1) / 2;
    h->h4 = (h->h2 + 1) / 2; h->w4 = (h->w2 + 1) / 2;
    h->h8 = (h->h4 + 1) / 2; h->w8 = (h->w4 + 1) / 2;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete h;
        return fail(nullptr, LWOP_ERR_CUDA, "cudaSetDevice failed");
    }
    build_program(h);
    *out = h;
    return LWOP_OK;
}

int lwop_destroy(lwop_handle *h) {
    if (h && h->host_calls)
        fprintf(stderr, "lwop host trace: %lld lwop_infer_host_begin calls, us per call: setup %.1f, upload+events %.1f, forward launch %.1f, "
                        "pack+download+events %.1f\n", h->host_calls, h->host_us[0] / h->host_calls, h->host_us[1] / h->host_calls,
                h->host_us[2] / h->host_calls, h->host_us[3] / h->host_calls);
    if (!h) return LWOP_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    destroy_plans(h);
    for (int i = 0; i < NUM_BUF; ++i) {
        if (h->arena_bf16[i]) cudaFree(h->arena_bf16[i]);
        if (h->arena_f32[i]) cudaFree(h->arena_f32[i]);
    }
    for (int i = 0; i < kNumLayers; ++i) {
        if (h->w_bf16[i]) cudaFree(h->w_bf16[i]);
        if (h->b_pad[i]) cudaFree(h->b_pad[i]);
        if (h->w_f32pad[i]) cudaFree(h->w_f32pad[i]);
    }
    if (h->blob) cudaFree(h->blob);
    if (h->peaks) cudaFree(h->peaks);
    if (h->peak_counts) cudaFree(h->peak_counts);
    if (h->group_work) cudaFree(h->group_work);
    if (h->pack_offsets) cudaFree(h->pack_offsets);
    if (h->draw_prio) cudaFree(h->draw_prio);
    if (h->h_offsets) cudaFreeHost(h->h_offsets);
    if (h->pk_joints) cudaFree(h->pk_joints);
    if (h->pk_limbs) cudaFree(h->pk_limbs);
    if (h->pk_persons) cudaFree(h->pk_persons);
    if (h->pk_keypoints) cudaFree(h->pk_keypoints);
    for (int k = 0; k < 2; ++k) {
        if (h->slot_in[k]) cudaFree(h->slot_in[k]);
        if (h->slot_heat[k]) cudaFree(h->slot_heat[k]);
        if (h->slot_paf[k]) cudaFree(h->slot_paf[k]);
        if (h->ev_h2d[k]) cudaEventDestroy(h->ev_h2d[k]);
        if (h->ev_done[k]) cudaEventDestroy(h->ev_done[k]);
        if (h->ev_fwd[k]) cudaEventDestroy(h->ev_fwd[k]);
    }
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->copy_stream && !h->copy_stream_borrowed) cudaStreamDestroy(h->copy_stream);
    if (h->decode_stream) cudaStreamDestroy(h->decode_stream);
    for (auto &ln : h->lanes) {
        if (ln.tables.counts) cudaFree(ln.tables.counts);
        if (ln.tables.joints) cudaFree(ln.tables.joints);
        if (ln.tables.limbs) cudaFree(ln.tables.limbs);
        if (ln.tables.persons) cudaFree(ln.tables.persons);
        if (ln.tables.keypoints) cudaFree(ln.tables.keypoints);
        if (ln.pack_offsets) cudaFree(ln.pack_offsets);
        if (ln.pk_joints) cudaFree(ln.pk_joints);
        if (ln.pk_limbs) cudaFree(ln.pk_limbs);
        if (ln.pk_persons) cudaFree(ln.pk_persons);
        if (ln.pk_keypoints) cudaFree(ln.pk_keypoints);
        if (ln.ev_ready) cudaEventDestroy(ln.ev_ready);
    }
    if (h->own_tables.counts) cudaFree(h->own_tables.counts);
    if (h->own_tables.joints) cudaFree(h->own_tables.joints);
    if (h->own_tables.limbs) cudaFree(h->own_tables.limbs);
    if (h->own_tables.persons) cudaFree(h->own_tables.persons);
    if (h->own_tables.keypoints) cudaFree(h->own_tables.keypoints);
    delete h;
    return LWOP_OK;
}

int lwop_load_weights(lwop_handle *h, const float *host_blob, size_t num_floats) {
    if (!h || !host_blob) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (num_floats != blob_floats())
        return fail(h, LWOP_ERR_INVALID,
                    "weight blob has " + std::to_string(num_floats) + " floats, expected " +
                        std::to_string(blob_floats()));
    CU_OK(h, cudaSetDevice(h->device));
    destroy_plans(h);
    if (!h->blob) CU_OK(h, cudaMalloc(&h->blob, num_floats * sizeof(float)));
    CU_OK(h, cudaMemcpy(h->blob, host_blob, num_floats * sizeof(float), cudaMemcpyHostToDevice));
    size_t off = 0;
    for (int i = 0; i < kNumLayers; ++i) {
        const LayerSpec &l = kLayers[i];
        if (l.kind == L_SEP) {
            h->off_dw[i] = off;
            off += 9 * (size_t)l.cin;
        }
        h->off_w[i] = off;
        off += (size_t)l.cout * l.k * l.k * l.cin;
        h->off_b[i] = off;
        off += l.cout;
    }
    // first conv: taps + bias packed for the by-value kernel parameter (per handle)
    CU_OK(h, conv1_pack_weights(host_blob + h->off_w[0], host_blob + h->off_b[0], &h->c1));
    // bf16 operand layouts for the tensor-core path: [Cout_pad][taps][Cin_pad], bias [Cout_pad]
    for (int i = 1; i < kNumLayers; ++i) {
        const LayerSpec &l = kLayers[i];
        const int taps = l.k * l.k, cin_pad = pad_cin(l.cin), cout_pad = pad_cout(l.cout);
        const size_t wbytes = (size_t)cout_pad * taps * cin_pad * 2;
        if (!h->w_bf16[i]) CU_OK(h, cudaMalloc(&h->w_bf16[i], wbytes));
        if (!h->b_pad[i]) CU_OK(h, cudaMalloc(&h->b_pad[i], (size_t)cout_pad * 4));
        CU_OK(h, cudaMemset(h->w_bf16[i], 0, wbytes));
        CU_OK(h, cudaMemset(h->b_pad[i], 0, (size_t)cout_pad * 4));
        // rows = Cout*taps (valid output channels only), cols = Cin -> Cin_pad
        CU_OK(h, launch_pack_bf16(h->blob + h->off_w[i], h->w_bf16[i], (long long)l.cout * taps, l.cin, cin_pad, 0));
        CU_OK(h, cudaMemcpy(h->b_pad[i], h->blob + h->off_b[i], (size_t)l.cout * 4, cudaMemcpyDeviceToDevice));
        if (cin_pad != l.cin) {   // fp32 copy with padded Cin for the CUDA-core paths (concat buffer pitch 64)
            if (!h->w_f32pad[i]) CU_OK(h, cudaMalloc(&h->w_f32pad[i], (size_t)l.cout * taps * cin_pad * 4));
            CU_OK(h, cudaMemset(h->w_f32pad[i], 0, (size_t)l.cout * taps * cin_pad * 4));
            CU_OK(h, cudaMemcpy2D(h->w_f32pad[i], (size_t)cin_pad * 4, h->blob + h->off_w[i], (size_t)l.cin * 4,
                                  (size_t)l.cin * 4, (size_t)l.cout * taps, cudaMemcpyDeviceToDevice));
        }
    }
    CU_OK(h, cudaDeviceSynchronize());
    h->has_weights = true;
    return LWOP_OK;
}

int lwop_forward(lwop_handle *h, const float *in_dev, int batch, int mode, float *heat_dev, float *paf_dev,
                 float *stage1_heat_dev, float *stage1_paf_dev, void *stream) {
    FwdIO io{in_dev, 0, heat_dev, paf_dev, stage1_heat_dev, stage1_paf_dev};
    return forward_impl(h, io, batch, mode, (cudaStream_t)stream);
}

int lwop_forward_u8(lwop_handle *h, const uint8_t *in_dev, int batch, int mode, float *heat_dev, float *paf_dev,
                    void *stream) {
    FwdIO io{in_dev, 1, heat_dev, paf_dev, nullptr, nullptr};
    return forward_impl(h, io, batch, mode, (cudaStream_t)stream);
}

int lwop_preprocess_bgr_u8(lwop_handle *h, const uint8_t *in_dev, int batch, int src_h, int src_w,
                           uint8_t *out_dev, int dst_h, int dst_w, void *stream) {
    if (!h || !in_dev || !out_dev) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (batch < 1 || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1) return fail(h, LWOP_ERR_INVALID, "bad geometry");
    CU_OK(h, cudaSetDevice(h->device));
    CU_OK(h, launch_preprocess_bgr(in_dev, batch, src_h, src_w, out_dev, dst_h, dst_w, (cudaStream_t)stream));
    h->launches++;
    return LWOP_OK;
}

int lwop_draw_poses(lwop_handle *h, const uint8_t *img_dev, int batch, int img_h, int img_w,
                    const lwop_decode_tables *t, uint8_t *canvas_dev, void *stream) {
    if (!h || !img_dev || !canvas_dev || !t || !t->counts || !t->joints || !t->persons)
        return fail(h, LWOP_ERR_INVALID, "null argument");
    if (batch < 1 || img_h < 1 || img_w < 1) return fail(h, LWOP_ERR_INVALID, "bad geometry");
    if (img_h > 16384 || img_w > 16384) return fail(h, LWOP_ERR_UNSUPPORTED, "frame larger than 16384 pixels per side");
    CU_OK(h, cudaSetDevice(h->device));
    const size_t pixels = (size_t)batch * img_h * img_w;
    if (pixels > h->draw_prio_cap) {
        // the map is all -1 between calls (the resolve pass resets what the mark pass touched); a new one starts so
        CU_OK(h, cudaStreamSynchronize((cudaStream_t)stream));
        if (h->draw_prio) cudaFree(h->draw_prio);
        h->draw_prio = nullptr;
        h->draw_prio_cap = 0;
        CU_OK(h, cudaMalloc(&h->draw_prio, pixels * sizeof(int)));
        h->draw_prio_cap = pixels;
        CU_OK(h, cudaMemsetAsync(h->draw_prio, 0xFF, pixels * sizeof(int), (cudaStream_t)stream));
    }
    CU_OK(h, launch_draw_poses(img_dev, batch, img_h, img_w, t->counts, t->joints, t->persons, h->draw_prio, canvas_dev,
                               (cudaStream_t)stream, &h->launches));
    return LWOP_OK;
}

int lwop_decode(lwop_handle *h, const float *heat_dev, const float *paf_dev, int batch, int map_h, int map_w,
                int img_h, int img_w, float thre1, double thre2, const lwop_decode_tables *t, void *stream) {
    if (!h || !heat_dev || !paf_dev || !t || !t->counts || !t->joints || !t->limbs || !t->persons || !t->keypoints)
        return fail(h, LWOP_ERR_INVALID, "null argument");
    if (batch < 1 || batch > h->max_batch) return fail(h, LWOP_ERR_INVALID, "batch out of range");
    if (map_h < 1 || map_w < 1 || img_h < map_h || img_w < map_w) return fail(h, LWOP_ERR_INVALID, "bad map size");
    if ((long long)map_h * map_w > 16384) return fail(h, LWOP_ERR_UNSUPPORTED, "heat map larger than 16384 pixels");
    CU_OK(h, cudaSetDevice(h->device));
    int rc = ensure_decode_scratch(h);
    if (rc != LWOP_OK) return rc;
    DecodeArgs a{};
    a.heat = heat_dev; a.paf = paf_dev; a.B = batch; a.h = map_h; a.w = map_w; a.H = img_h; a.W = img_w;
    a.thre1 = thre1; a.thre2 = thre2; a.t = *t; a.peaks = h->peaks; a.peak_counts = h->peak_counts; a.work = h->group_work;
    CU_OK(h, launch_decode(a, (cudaStream_t)stream, &h->launches));
    return LWOP_OK;
}

int lwop_forward_decode(lwop_handle *h, const void *in_dev, int is_u8, int batch, int mode, float thre1, double thre2,
                        float *heat_dev, float *paf_dev, const lwop_decode_tables *t, void *stream) {
    if (!h || !t || !t->counts || !t->joints || !t->limbs || !t->persons || !t->keypoints)
        return fail(h, LWOP_ERR_INVALID, "null argument");
    if ((long long)h->h8 * h->w8 > 16384) return fail(h, LWOP_ERR_UNSUPPORTED, "heat map larger than 16384 pixels");
    CU_OK(h, cudaSetDevice(h->device));
    int rc = ensure_decode_scratch(h);
    if (rc != LWOP_OK) return rc;
    DecodeArgs a{};
    a.thre1 = thre1; a.thre2 = thre2; a.t = *t; a.peaks = h->peaks; a.peak_counts = h->peak_counts; a.work = h->group_work;
    FwdIO io{in_dev, is_u8, heat_dev, paf_dev, nullptr, nullptr, &a};
    return forward_impl(h, io, batch, mode, (cudaStream_t)stream);
}

int lwop_nms(lwop_handle *h, const float *heat_dev, int batch, int map_h, int map_w, double factor, float thre1,
             int mode, float *peaks_dev, int32_t *peak_counts_dev, int32_t *counts_dev, void *stream) {
    if (!h || !heat_dev || !peaks_dev || !peak_counts_dev || !counts_dev) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (batch < 1 || map_h < 1 || map_w < 1 || !(factor > 0.0)) return fail(h, LWOP_ERR_INVALID, "bad geometry");
    if ((long long)map_h * map_w > 16384) return fail(h, LWOP_ERR_UNSUPPORTED, "heat map larger than 16384 pixels");
    if (mode & ~(LWOP_PEAKS_NO_REFINE | LWOP_PEAKS_CROSS | LWOP_PEAKS_GAUSS)) return fail(h, LWOP_ERR_INVALID, "unknown peak mode");
    if ((mode & LWOP_PEAKS_GAUSS) && (mode & LWOP_PEAKS_NO_REFINE))
        return fail(h, LWOP_ERR_INVALID, "the Gaussian filter applies to the refined patch only");
    if ((mode & LWOP_PEAKS_GAUSS) && 5.0 * factor > 48.4)
        return fail(h, LWOP_ERR_UNSUPPORTED, "Gaussian patch filter needs upsampFactor <= 9.6 (patch <= 48 x 48)");
    CU_OK(h, cudaSetDevice(h->device));
    CU_OK(h, launch_decode_peaks(heat_dev, batch, map_h, map_w, factor, thre1, mode, peaks_dev, peak_counts_dev,
                                 counts_dev, (cudaStream_t)stream));
    h->launches += 3;
    return LWOP_OK;
}

int lwop_find_connected_joints(lwop_handle *h, const float *paf_dev, int paf_upsampled, int batch, int map_h, int map_w,
                               int img_h, int img_w, double thre2, int num_intermed_pts, double max_dist_thresh,
                               const float *peaks_dev, const int32_t *peak_counts_dev, double *limbs_dev,
                               int32_t *counts_dev, void *stream) {
    if (!h || !paf_dev || !peaks_dev || !peak_counts_dev || !limbs_dev || !counts_dev)
        return fail(h, LWOP_ERR_INVALID, "null argument");
    if (batch < 1 || map_h < 1 || map_w < 1 || img_h < 1 || img_w < 1) return fail(h, LWOP_ERR_INVALID, "bad geometry");
    if (num_intermed_pts < 1 || num_intermed_pts > 64)
        return fail(h, LWOP_ERR_UNSUPPORTED, "num_intermed_pts must be in 1..64");
    CU_OK(h, cudaSetDevice(h->device));
    CU_OK(h, launch_decode_limbs(paf_dev, paf_upsampled, batch, map_h, map_w, img_h, img_w, thre2, num_intermed_pts,
                                 max_dist_thresh, peaks_dev,
                                 peak_counts_dev, limbs_dev, counts_dev, (cudaStream_t)stream));
    h->launches++;
    return LWOP_OK;
}

int lwop_group_limbs(lwop_handle *h, const float *peaks_dev, const int32_t *peak_counts_dev, int batch,
                     const lwop_decode_tables *t, void *stream) {
    if (!h || !peaks_dev || !peak_counts_dev || !t || !t->counts || !t->joints || !t->limbs || !t->persons || !t->keypoints)
        return fail(h, LWOP_ERR_INVALID, "null argument");
    if (batch < 1 || batch > h->max_batch) return fail(h, LWOP_ERR_INVALID, "batch out of range");
    CU_OK(h, cudaSetDevice(h->device));
    int rc = ensure_decode_scratch(h);
    if (rc != LWOP_OK) return rc;
    CU_OK(h, launch_decode_group(peaks_dev, peak_counts_dev, t->limbs, t->counts, t->joints, t->persons, t->keypoints,
                                 h->group_work, batch, (cudaStream_t)stream));
    h->launches += 2;
    return LWOP_OK;
}

int lwop_decode_fetch(lwop_handle *h, const lwop_decode_tables *t, int batch, const lwop_packed_tables *o,
                      void *stream) {
    if (!h || !t || !o || !o->counts) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (batch < 1 || batch > h->max_batch) return fail(h, LWOP_ERR_INVALID, "batch out of range");
    cudaStream_t s = (cudaStream_t)stream;
    CU_OK(h, cudaSetDevice(h->device));
    int rc = ensure_decode_scratch(h);
    if (rc != LWOP_OK) return rc;
    // worst case rows are bounded by the fixed capacities; size the packed device buffers for this batch
    // lazily from the actual totals (phase 1: counts + offsets)
    CU_OK(h, cudaMemcpyAsync(o->counts, t->counts, (size_t)batch * LWOP_COUNTS * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, s));
    CU_OK(h, cudaStreamSynchronize(s));
    size_t nj = 0, np = 0, nc = 0;
    bool overflow = false;
    for (int b = 0; b < batch; ++b) {
        const int32_t *c = o->counts + (size_t)b * LWOP_COUNTS;
        nj += c[0]; np += c[1]; nc += c[3];
        overflow |= c[2] != 0;
    }
    if (nj > o->joints_cap || nc > o->limbs_cap || np > o->persons_cap)
        return fail(h, LWOP_ERR_OVERFLOW, "host result buffers too small: need " + std::to_string(nj) + " joint, " +
                                              std::to_string(nc) + " limb, " + std::to_string(np) + " person rows");
    rc = ensure_pack_buffers(h, nj, nc, np);
    if (rc != LWOP_OK) return rc;
    CU_OK(h, launch_decode_pack(*t, batch, h->pack_offsets, h->pk_joints, h->pk_limbs, h->pk_persons,
                                h->pk_keypoints, s, &h->launches));
    if (nj) CU_OK(h, cudaMemcpyAsync(o->joints, h->pk_joints, nj * 6 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (nc) CU_OK(h, cudaMemcpyAsync(o->limbs, h->pk_limbs, nc * 5 * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (np) {
        CU_OK(h, cudaMemcpyAsync(o->persons, h->pk_persons, np * 16 * sizeof(double), cudaMemcpyDeviceToHost, s));
        CU_OK(h, cudaMemcpyAsync(o->keypoints, h->pk_keypoints, np * 42 * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    CU_OK(h, cudaStreamSynchronize(s));
    if (overflow) return fail(h, LWOP_ERR_OVERFLOW, "a per-image decode table capacity was exceeded (counts[b][2])");
    return LWOP_OK;
}

namespace {

int ensure_lane(lwop_handle *h, lwop_handle::HostLane &ln, const lwop_packed_tables *o) {
    const size_t B = h->max_batch;
    if (!ln.tables.counts) {
        CU_OK(h, cudaMalloc(&ln.tables.counts, B * LWOP_COUNTS * sizeof(int32_t)));
        CU_OK(h, cudaMalloc(&ln.tables.joints, B * LWOP_MAX_JOINTS * 6 * sizeof(double)));
        CU_OK(h, cudaMalloc(&ln.tables.limbs, B * LWOP_NUM_LIMBS * LWOP_MAX_PEAKS * 5 * sizeof(double)));
        CU_OK(h, cudaMalloc(&ln.tables.persons, B * LWOP_MAX_PERSONS * 16 * sizeof(double)));
        CU_OK(h, cudaMalloc(&ln.tables.keypoints, B * LWOP_MAX_PERSONS * 42 * sizeof(float)));
        CU_OK(h, cudaMalloc(&ln.pack_offsets, (B + 1) * 4 * sizeof(int)));
        CU_OK(h, cudaEventCreateWithFlags(&ln.ev_ready, cudaEventDisableTiming));
    }
    if (o->joints_cap > ln.cap_j) {
        if (ln.pk_joints) cudaFree(ln.pk_joints);
        ln.cap_j = o->joints_cap;
        CU_OK(h, cudaMalloc(&ln.pk_joints, ln.cap_j * 6 * sizeof(double)));
    }
    if (o->limbs_cap > ln.cap_l) {
        if (ln.pk_limbs) cudaFree(ln.pk_limbs);
        ln.cap_l = o->limbs_cap;
        CU_OK(h, cudaMalloc(&ln.pk_limbs, ln.cap_l * 5 * sizeof(double)));
    }
    if (o->persons_cap > ln.cap_p) {
        if (ln.pk_persons) cudaFree(ln.pk_persons);
        if (ln.pk_keypoints) cudaFree(ln.pk_keypoints);
        ln.cap_p = o->persons_cap;
        CU_OK(h, cudaMalloc(&ln.pk_persons, ln.cap_p * 16 * sizeof(double)));
        CU_OK(h, cudaMalloc(&ln.pk_keypoints, ln.cap_p * 42 * sizeof(float)));
    }
    return LWOP_OK;
}

}  // namespace

namespace {
int infer_host_begin_impl(lwop_handle *h, const void *images_host, int is_u8, int batch, int mode, float thre1,
                          double thre2, const lwop_packed_tables *o, float *heat_host, float *paf_host, void *stream,
                          int *ticket, bool lone_call);
}

int lwop_infer_host_begin(lwop_handle *h, const void *images_host, int is_u8, int batch, int mode, float thre1,
                          double thre2, const lwop_packed_tables *o, float *heat_host, float *paf_host, void *stream,
                          int *ticket) {
    return infer_host_begin_impl(h, images_host, is_u8, batch, mode, thre1, thre2, o, heat_host, paf_host, stream, ticket,
                                 false);
}

namespace {
int infer_host_begin_impl(lwop_handle *h, const void *images_host, int is_u8, int batch, int mode, float thre1,
                          double thre2, const lwop_packed_tables *o, float *heat_host, float *paf_host, void *stream,
                          int *ticket, bool lone_call) {
    if (!h || !images_host || !o || !o->counts || !ticket) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (batch < 1 || batch > h->max_batch) return fail(h, LWOP_ERR_INVALID, "batch out of range");
    if (!h->has_weights) return fail(h, LWOP_ERR_NO_WEIGHTS, "lwop_infer_host before lwop_load_weights");
    int lane = -1;
    for (int i = 0; i < LWOP_MAX_INFLIGHT; ++i)
        if (!h->lanes[i].busy) { lane = i; break; }
    if (lane < 0) return fail(h, LWOP_ERR_INVALID, "too many batches in flight: call lwop_infer_host_end first");
    static const bool trace_host = getenv("LWOP_TRACE_HOST") != nullptr;
    auto t_now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_seg = trace_host ? t_now() : 0.0;
    auto seg = [&](int k) { if (trace_host) { const double t = t_now(); h->host_us[k] += t - t_seg; t_seg = t; } };
    cudaStream_t s = (cudaStream_t)stream;
    CU_OK(h, cudaSetDevice(h->device));
    // The batch is cut into chunks that flow through a two-slot pipeline: the copy stream uploads chunk
    // k+1 while the compute stream runs the forward of chunk k and the decode stream decodes chunk k-1
    // (one CUDA graph per slot).  Slots are guarded by events across calls as well, so the upload of the NEXT
    // batch overlaps the kernels of this one when the caller keeps two batches in flight.
    // Default chunking: the synchronous lwop_infer_host hides its own upload by cutting the batch into chunks; a
    // batch submitted through the asynchronous entry hides its upload behind the kernels of the batch before it,
    // so it goes through whole (fewer, larger launches).
    int chunk = h->chunk > 0 ? h->chunk : (!lone_call ? batch : (is_u8 ? 128 : 64));
    if (chunk > batch) chunk = batch;
    if (chunk > h->max_batch) chunk = h->max_batch;
    const size_t img_elems = (size_t)h->H * h->W * 3;
    const size_t px = (size_t)h->h8 * h->w8;
    if (!h->pipe_ready) {
        const int alloc_chunk = h->max_batch;    // slots are sized for max_batch once: the chunk size may vary per call
        CU_OK(h, cudaDeviceSynchronize());       // nothing may still use the old slots
        for (int k = 0; k < 2; ++k) {
            if (h->slot_in[k]) cudaFree(h->slot_in[k]);
            if (h->slot_heat[k]) cudaFree(h->slot_heat[k]);
            if (h->slot_paf[k]) cudaFree(h->slot_paf[k]);
            CU_OK(h, cudaMalloc(&h->slot_in[k], (size_t)alloc_chunk * img_elems * 4));
            CU_OK(h, cudaMalloc(&h->slot_heat[k], (size_t)alloc_chunk * px * LWOP_NUM_JOINTS * 4));
            CU_OK(h, cudaMalloc(&h->slot_paf[k], (size_t)alloc_chunk * px * LWOP_NUM_PAFS * 4));
            if (!h->ev_h2d[k]) CU_OK(h, cudaEventCreateWithFlags(&h->ev_h2d[k], cudaEventDisableTiming));
            if (!h->ev_done[k]) CU_OK(h, cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming));
            if (!h->ev_fwd[k]) CU_OK(h, cudaEventCreateWithFlags(&h->ev_fwd[k], cudaEventDisableTiming));
        }
        if (!h->ev_join) CU_OK(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        if (!h->copy_stream) CU_OK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        if (!h->decode_stream) CU_OK(h, cudaStreamCreateWithFlags(&h->decode_stream, cudaStreamNonBlocking));
        h->pipe_chunk = alloc_chunk;
        h->pipe_ready = true;
        destroy_graphs_only(h);          // slot pointers changed
    }
    lwop_handle::HostLane &ln = h->lanes[lane];
    int rc = ensure_lane(h, ln, o);
    if (rc != LWOP_OK) return rc;
    rc = ensure_decode_scratch(h);
    if (rc != LWOP_OK) return rc;
    const size_t esz = is_u8 ? 1 : 4;
    cudaStream_t ds = h->decode_stream;
    seg(0);
    // chunk schedule: a half-size first chunk shortens the only upload that nothing can hide
    int k = 0;
    for (int b0 = 0, nb = 0; b0 < batch; b0 += nb, ++k) {
        nb = (k == 0 && batch > chunk && chunk >= 16) ? chunk / 2 : chunk;
        if (nb > batch - b0) nb = batch - b0;
        // The synchronous entry pipelines the chunks of ONE batch through the two slots (alternating across calls too);
        // a batch submitted through the asynchronous entry goes through whole and takes the slot of its result lane, so
        // that an engine only ever sees two (slot, lane) pairs -- each pair is a CUDA graph of its own, and with a free
        // running slot counter a drained pipeline could restart on the two pairs it had never captured.
        const int slot = lone_call ? (int)(h->slot_seq++ & 1) : lane;
        CU_OK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_fwd[slot], 0));   // previous frames in this slot consumed
        CU_OK(h, cudaStreamWaitEvent(s, h->ev_done[slot], 0));               // previous maps in this slot decoded
        CU_OK(h, cudaMemcpyAsync(h->slot_in[slot], (const char *)images_host + (size_t)b0 * img_elems * esz,
                                 (size_t)nb * img_elems * esz, cudaMemcpyHostToDevice, h->copy_stream));
        CU_OK(h, cudaEventRecord(h->ev_h2d[slot], h->copy_stream));
        CU_OK(h, cudaStreamWaitEvent(s, h->ev_h2d[slot], 0));
        seg(1);
        // forward + decode chain: ONE graph launch on `s` (the decode kernels follow the last head kernel with
        // programmatic dependent launches); compaction and the downloads run on the decode stream behind it
        lwop_decode_tables t = ln.tables;
        t.counts += (size_t)b0 * LWOP_COUNTS;
        t.joints += (size_t)b0 * LWOP_MAX_JOINTS * 6;
        t.limbs += (size_t)b0 * LWOP_NUM_LIMBS * LWOP_MAX_PEAKS * 5;
        t.persons += (size_t)b0 * LWOP_MAX_PERSONS * 16;
        t.keypoints += (size_t)b0 * LWOP_MAX_PERSONS * 42;
        DecodeArgs a{};
        a.thre1 = thre1; a.thre2 = thre2; a.t = t;
        a.peaks = h->peaks + (size_t)b0 * LWOP_NUM_JOINTS * LWOP_MAX_PEAKS * 4;
        a.peak_counts = h->peak_counts + (size_t)b0 * 16;
        a.work = h->group_work + (size_t)b0 * LWOP_NUM_LIMBS * LWOP_MAX_PEAKS * 16;
        FwdIO io{h->slot_in[slot], is_u8, h->slot_heat[slot], h->slot_paf[slot], nullptr, nullptr, &a};
        rc = forward_impl(h, io, nb, mode, s);
        if (rc != LWOP_OK) return rc;
        seg(2);
        CU_OK(h, cudaEventRecord(h->ev_fwd[slot], s));
        CU_OK(h, cudaStreamWaitEvent(ds, h->ev_fwd[slot], 0));
        if (heat_host)
            CU_OK(h, cudaMemcpyAsync(heat_host + (size_t)b0 * px * LWOP_NUM_JOINTS, h->slot_heat[slot],
                                     (size_t)nb * px * LWOP_NUM_JOINTS * 4, cudaMemcpyDeviceToHost, ds));
        if (paf_host)
            CU_OK(h, cudaMemcpyAsync(paf_host + (size_t)b0 * px * LWOP_NUM_PAFS, h->slot_paf[slot],
                                     (size_t)nb * px * LWOP_NUM_PAFS * 4, cudaMemcpyDeviceToHost, ds));
        CU_OK(h, cudaEventRecord(h->ev_done[slot], ds));
    }
    // compaction + download of the result tables, still asynchronous: the packed buffers are copied at the
    // capacities the caller provided (rows beyond a capacity are dropped and reported by lwop_infer_host_end)
    CU_OK(h, launch_decode_pack(ln.tables, batch, ln.pack_offsets, ln.pk_joints, ln.pk_limbs, ln.pk_persons,
                                ln.pk_keypoints, ds, &h->launches, (long long)o->joints_cap, (long long)o->limbs_cap,
                                (long long)o->persons_cap));
    CU_OK(h, cudaMemcpyAsync(o->counts, ln.tables.counts, (size_t)batch * LWOP_COUNTS * sizeof(int32_t),
                             cudaMemcpyDeviceToHost, ds));
    if (o->joints_cap) CU_OK(h, cudaMemcpyAsync(o->joints, ln.pk_joints, o->joints_cap * 6 * sizeof(double), cudaMemcpyDeviceToHost, ds));
    if (o->limbs_cap) CU_OK(h, cudaMemcpyAsync(o->limbs, ln.pk_limbs, o->limbs_cap * 5 * sizeof(double), cudaMemcpyDeviceToHost, ds));
    if (o->persons_cap) {
        CU_OK(h, cudaMemcpyAsync(o->persons, ln.pk_persons, o->persons_cap * 16 * sizeof(double), cudaMemcpyDeviceToHost, ds));
        CU_OK(h, cudaMemcpyAsync(o->keypoints, ln.pk_keypoints, o->persons_cap * 42 * sizeof(float), cudaMemcpyDeviceToHost, ds));
    }
    CU_OK(h, cudaEventRecord(ln.ev_ready, ds));
    seg(3);
    if (trace_host) ++h->host_calls;
    // `s` deliberately does NOT wait for ev_ready: the next batch's forward may start while this batch is still
    // being decoded / downloaded on the decode stream (lwop_infer_host_end synchronises on ev_ready)
    ln.busy = true;
    ln.batch = batch;
    ln.out = *o;
    *ticket = lane;
    return LWOP_OK;
}
}  // namespace

int lwop_infer_host_end(lwop_handle *h, int ticket) {
    if (!h || ticket < 0 || ticket >= LWOP_MAX_INFLIGHT || !h->lanes[ticket].busy)
        return fail(h, LWOP_ERR_INVALID, "bad ticket");
    lwop_handle::HostLane &ln = h->lanes[ticket];
    CU_OK(h, cudaSetDevice(h->device));
    cudaError_t e = cudaEventSynchronize(ln.ev_ready);
    ln.busy = false;
    if (e != cudaSuccess) return fail(h, LWOP_ERR_CUDA, std::string("infer_host: ") + cudaGetErrorString(e));
    size_t nj = 0, np = 0, nc = 0;
    bool overflow = false;
    for (int b = 0; b < ln.batch; ++b) {
        const int32_t *c = ln.out.counts + (size_t)b * LWOP_COUNTS;
        nj += c[0]; np += c[1]; nc += c[3];
        overflow |= c[2] != 0;
    }
    if (nj > ln.out.joints_cap || nc > ln.out.limbs_cap || np > ln.out.persons_cap)
        return fail(h, LWOP_ERR_OVERFLOW, "host result buffers too small: need " + std::to_string(nj) + " joint, " +
                                              std::to_string(nc) + " limb, " + std::to_string(np) + " person rows");
    if (overflow) return fail(h, LWOP_ERR_OVERFLOW, "a per-image decode table capacity was exceeded (counts[b][2])");
    return LWOP_OK;
}

int lwop_infer_host(lwop_handle *h, const void *images_host, int is_u8, int batch, int mode, float thre1,
                    double thre2, const lwop_packed_tables *o, float *heat_host, float *paf_host, void *stream) {
    int ticket = -1;
    int rc = infer_host_begin_impl(h, images_host, is_u8, batch, mode, thre1, thre2, o, heat_host, paf_host, stream,
                                   &ticket, true);
    if (rc != LWOP_OK) return rc;
    return lwop_infer_host_end(h, ticket);
}

int lwop_set_option(lwop_handle *h, const char *name, long long value) {
    if (!h || !name) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (strcmp(name, "chunk") == 0) {
        if (value < 0) return fail(h, LWOP_ERR_INVALID, "chunk must be >= 0 (0 = library default)");
        h->chunk = (int)value;
        return LWOP_OK;
    }
    if (strcmp(name, "copy_stream") == 0) {
        // uploads of lwop_infer_host* go through this caller-owned stream (a cudaStream_t passed as an integer) instead of
        // the handle's private one: handles that share it upload in submission order, one batch at a time at full PCIe
        // rate, instead of all at once at a fraction of it each (0 = back to a private stream)
        if (h->copy_stream && !h->copy_stream_borrowed) {
            cudaSetDevice(h->device);
            cudaStreamSynchronize(h->copy_stream);
            cudaStreamDestroy(h->copy_stream);
        }
        h->copy_stream = (cudaStream_t)(uintptr_t)value;
        h->copy_stream_borrowed = value != 0;
        return LWOP_OK;
    }
    return fail(h, LWOP_ERR_INVALID, std::string("unknown option ") + name);
}

int lwop_conv2d(lwop_handle *h, int mode, const void *in_dev, int in_dtype, int batch, int height, int width, int cin,
                const float *w_dev, const float *bias_dev, int cout, int ksize, int stride, int dilation, int act,
                const void *residual_dev, void *out_dev, int out_dtype, void *stream) {
    if (!h || !in_dev || !w_dev || !out_dev) return fail(h, LWOP_ERR_INVALID, "null argument");
    cudaStream_t s = (cudaStream_t)stream;
    CU_OK(h, cudaSetDevice(h->device));
    if (mode != LWOP_MODE_BF16) {
        ConvArgs a{};
        a.in = in_dev; a.in_dtype = in_dtype; a.w = w_dev; a.bias = bias_dev; a.residual = residual_dev;
        a.out = out_dev; a.out_dtype = out_dtype;
        a.B = batch; a.H = height; a.W = width; a.Cin = cin; a.Cout = cout; a.ksize = ksize; a.stride = stride;
        a.dil = dilation; a.act = act;
        CU_OK(h, launch_conv_direct(a, s));
        h->launches++;
        return LWOP_OK;
    }
    // tensor-core path: stride-1 1x1 / 3x3 with bf16 input; weights are packed per call (unit-test entry)
    if (in_dtype != LWOP_DT_BF16 || stride != 1 || (ksize != 1 && ksize != 3) || cin % 8 != 0)
        return fail(h, LWOP_ERR_UNSUPPORTED, "tcgen05 conv needs bf16 input, stride 1, k in {1,3}, Cin % 8 == 0");
    const int taps = ksize * ksize;
    const int cin_pad = cin % 32 == 0 ? cin : ((cin + 63) / 64) * 64;
    if (cin_pad != cin) return fail(h, LWOP_ERR_UNSUPPORTED, "tcgen05 conv entry needs Cin % 32 == 0 (pad the input)");
    const int cout_pad = pad_cout(cout);
    void *wp = nullptr;
    float *bp = nullptr;
    CU_OK(h, cudaMalloc(&wp, (size_t)cout_pad * taps * cin_pad * 2));
    CU_OK(h, cudaMalloc(&bp, (size_t)cout_pad * 4));
    CU_OK(h, cudaMemsetAsync(wp, 0, (size_t)cout_pad * taps * cin_pad * 2, s));
    CU_OK(h, cudaMemsetAsync(bp, 0, (size_t)cout_pad * 4, s));
    CU_OK(h, launch_pack_bf16(w_dev, wp, (long long)cout * taps, cin, cin_pad, s));
    if (bias_dev) CU_OK(h, cudaMemcpyAsync(bp, bias_dev, (size_t)cout * 4, cudaMemcpyDeviceToDevice, s));
    GemmConvDesc d{};
    d.in = in_dev; d.w_packed = wp; d.bias = bp; d.residual = residual_dev; d.out = out_dev; d.out2 = nullptr;
    d.out_dtype = out_dtype; d.lda = cin_pad; d.B = batch; d.H = height; d.W = width; d.Cin_pad = cin_pad;
    d.Cout = cout; d.Cout_pad = cout_pad; d.ksize = ksize; d.dil = dilation; d.act = act; d.ldo = cout; d.out_col0 = 0;
    d.ldr = cout;
    char err[256];
    GemmPlan *pl = gemm_plan_create(d, err, sizeof(err));
    int rc = LWOP_OK;
    if (!pl) {
        rc = fail(h, LWOP_ERR_UNSUPPORTED, err);
    } else {
        cudaError_t e = gemm_plan_launch(pl, s);
        h->launches += 2;
        if (e != cudaSuccess) rc = fail(h, LWOP_ERR_CUDA, std::string("gemm launch: ") + cudaGetErrorString(e));
        cudaError_t e2 = cudaStreamSynchronize(s);
        if (rc == LWOP_OK && e2 != cudaSuccess)
            rc = fail(h, LWOP_ERR_CUDA, std::string("gemm kernel: ") + cudaGetErrorString(e2));
        gemm_plan_destroy(pl);
    }
    cudaFree(wp);
    cudaFree(bp);
    return rc;
}

int lwop_first_conv(lwop_handle *h, const void *in_dev, int in_is_u8, int batch, int height, int width,
                    const float *w_host, const float *b_host, void *out_dev, void *stream) {
    if (!h || !in_dev || !w_host || !b_host || !out_dev) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (batch < 1 || height < 1 || width < 1) return fail(h, LWOP_ERR_INVALID, "bad geometry");
    CU_OK(h, cudaSetDevice(h->device));
    Conv1Weights cw;
    CU_OK(h, conv1_pack_weights(w_host, b_host, &cw));
    CU_OK(h, launch_conv1_bf16(cw, in_dev, in_is_u8, out_dev, batch, height, width, (cudaStream_t)stream));
    h->launches++;
    return LWOP_OK;
}

int lwop_depthwise3x3(lwop_handle *h, int mode, const void *in_dev, int dtype, int batch, int height, int width,
                      int channels, const float *w_dev, int stride, int dilation, void *out_dev, void *stream) {
    if (!h || !in_dev || !w_dev || !out_dev) return fail(h, LWOP_ERR_INVALID, "null argument");
    CU_OK(h, cudaSetDevice(h->device));
    DwArgs a{};
    a.in = in_dev; a.out = out_dev; a.dtype = dtype; a.w = w_dev;
    a.B = batch; a.H = height; a.W = width; a.C = channels; a.stride = stride; a.dil = dilation;
    if (mode == LWOP_MODE_BF16 && dtype == LWOP_DT_BF16 && dw_plan_supported(channels, stride, dilation)) {
        char err[256];
        DwPlan *pl = dw_plan_create(a, err, sizeof(err));
        if (!pl) return fail(h, LWOP_ERR_UNSUPPORTED, err);
        cudaError_t e = dw_plan_launch(pl, (cudaStream_t)stream);
        dw_plan_destroy(pl);      // the tensor map is a by-value kernel argument
        if (e != cudaSuccess) return fail(h, LWOP_ERR_CUDA, std::string("depthwise launch: ") + cudaGetErrorString(e));
    } else if ((mode == LWOP_MODE_BF16 || mode == LWOP_MODE_BF16_DIRECT) && dtype == LWOP_DT_BF16) {
        if (channels % 8 != 0 || !((stride == 1 && dilation <= 2) || (stride == 2 && dilation == 1)))
            return fail(h, LWOP_ERR_UNSUPPORTED, "vector depthwise needs C % 8 == 0 and (s1,d1|d2) or (s2,d1)");
        CU_OK(h, launch_depthwise_bf16(a, (cudaStream_t)stream));
    } else {
        CU_OK(h, launch_depthwise_simple(a, (cudaStream_t)stream));
    }
    h->launches++;
    return LWOP_OK;
}

int lwop_depthwise2d(lwop_handle *h, int mode, const void *in_dev, int dtype, int batch, int height, int width,
                     int channels, const float *w_dev, int ksize, int stride, int dilation, void *out_dev, void *stream) {
    if (ksize == 3)
        return lwop_depthwise3x3(h, mode, in_dev, dtype, batch, height, width, channels, w_dev, stride, dilation, out_dev, stream);
    if (!h || !in_dev || !w_dev || !out_dev) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (ksize < 1 || ksize > 15 || stride < 1 || dilation < 1 || batch < 1 || height < 1 || width < 1 || channels < 1)
        return fail(h, LWOP_ERR_INVALID, "bad depthwise geometry");
    if (dtype != LWOP_DT_F32 && dtype != LWOP_DT_BF16) return fail(h, LWOP_ERR_INVALID, "bad dtype");
    CU_OK(h, cudaSetDevice(h->device));
    DwArgs a{};
    a.in = in_dev; a.out = out_dev; a.dtype = dtype; a.w = w_dev;
    a.B = batch; a.H = height; a.W = width; a.C = channels; a.stride = stride; a.dil = dilation; a.ksize = ksize;
    CU_OK(h, launch_depthwise_simple(a, (cudaStream_t)stream));      // fp32 accumulation on CUDA cores, any mode
    h->launches++;
    return LWOP_OK;
}

int lwop_separable_conv2d(lwop_handle *h, int mode, const void *in_dev, int batch, int height, int width, int cin,
                          const float *dw_w_dev, const float *pw_w_dev, const float *bias_dev, int cout,
                          int dilation, int act, const void *residual_dev, void *out_dev, void *stream) {
    if (!h || !in_dev || !dw_w_dev || !pw_w_dev || !out_dev) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (mode != LWOP_MODE_BF16) return fail(h, LWOP_ERR_UNSUPPORTED, "fused separable conv exists in bf16 mode only");
    const int cout_pad = pad_cout(cout);
    if (cout != cout_pad || !sep_plan_supported(cin, cout_pad, 1, dilation))
        return fail(h, LWOP_ERR_UNSUPPORTED, "fused separable conv: unsupported (cin, cout, dilation)");
    cudaStream_t s = (cudaStream_t)stream;
    CU_OK(h, cudaSetDevice(h->device));
    void *wp = nullptr;
    float *bp = nullptr;
    CU_OK(h, cudaMalloc(&wp, (size_t)cout_pad * cin * 2));
    CU_OK(h, cudaMalloc(&bp, (size_t)cout_pad * 4));
    CU_OK(h, cudaMemsetAsync(bp, 0, (size_t)cout_pad * 4, s));
    CU_OK(h, launch_pack_bf16(pw_w_dev, wp, (long long)cout, cin, cin, s));
    if (bias_dev) CU_OK(h, cudaMemcpyAsync(bp, bias_dev, (size_t)cout * 4, cudaMemcpyDeviceToDevice, s));
    SepConvDesc d{};
    d.in = in_dev; d.dw_w = dw_w_dev; d.w_packed = wp; d.bias = bp; d.residual = residual_dev; d.out = out_dev;
    d.B = batch; d.H = height; d.W = width; d.C = cin; d.Cout_pad = cout_pad; d.dil = dilation; d.act = act;
    d.ldo = cout; d.out_col0 = 0; d.ldr = cout;
    char err[256];
    SepPlan *pl = sep_plan_create(d, err, sizeof(err));
    int rc = LWOP_OK;
    if (!pl) {
        rc = fail(h, LWOP_ERR_UNSUPPORTED, err);
    } else {
        cudaError_t e = sep_plan_launch(pl, s);
        h->launches += 2;
        if (e != cudaSuccess) rc = fail(h, LWOP_ERR_CUDA, std::string("fused separable launch: ") + cudaGetErrorString(e));
        cudaError_t e2 = cudaStreamSynchronize(s);
        if (rc == LWOP_OK && e2 != cudaSuccess)
            rc = fail(h, LWOP_ERR_CUDA, std::string("fused separable kernel: ") + cudaGetErrorString(e2));
        sep_plan_destroy(pl);
    }
    cudaFree(wp);
    cudaFree(bp);
    return rc;
}

int lwop_head_pair(lwop_handle *h, const void *x_dev, long long m, int hidden,
                   const float *w1a_dev, const float *b1a_dev, const float *w2a_dev, const float *b2a_dev,
                   const float *w1b_dev, const float *b1b_dev, const float *w2b_dev, const float *b2b_dev,
                   float *out_a_dev, float *out_b_dev, void *stream) {
    if (!h || !x_dev || !w1a_dev || !b1a_dev || !w2a_dev || !b2a_dev || !w1b_dev || !b1b_dev || !w2b_dev || !b2b_dev ||
        !out_a_dev || !out_b_dev)
        return fail(h, LWOP_ERR_INVALID, "null argument");
    if (m < 1 || !head_pair_supported(128, hidden)) return fail(h, LWOP_ERR_UNSUPPORTED, "head pair: hidden must be 128 or 512");
    cudaStream_t s = (cudaStream_t)stream;
    CU_OK(h, cudaSetDevice(h->device));
    unsigned char *buf = nullptr;
    const size_t w1b = (size_t)hidden * 128 * 2, w2ab = (size_t)16 * hidden * 2, w2bb = (size_t)32 * hidden * 2;
    const size_t total = 2 * w1b + w2ab + w2bb + 48 * 4;
    CU_OK(h, cudaMalloc(&buf, total));
    CU_OK(h, cudaMemsetAsync(buf, 0, total, s));
    void *w1a = buf, *w1bp = buf + w1b, *w2a = buf + 2 * w1b, *w2b = buf + 2 * w1b + w2ab;
    float *b2 = (float *)(buf + 2 * w1b + w2ab + w2bb);
    CU_OK(h, launch_pack_bf16(w1a_dev, w1a, hidden, 128, 128, s));
    CU_OK(h, launch_pack_bf16(w1b_dev, w1bp, hidden, 128, 128, s));
    CU_OK(h, launch_pack_bf16(w2a_dev, w2a, 14, hidden, hidden, s));
    CU_OK(h, launch_pack_bf16(w2b_dev, w2b, 26, hidden, hidden, s));
    CU_OK(h, cudaMemcpyAsync(b2, b2a_dev, 14 * 4, cudaMemcpyDeviceToDevice, s));
    CU_OK(h, cudaMemcpyAsync(b2 + 16, b2b_dev, 26 * 4, cudaMemcpyDeviceToDevice, s));
    HeadPairDesc d{};
    d.in = x_dev; d.lda = 128; d.Cin = 128; d.M = m; d.hidden = hidden;
    d.w1a = w1a; d.w1b = w1bp; d.b1a = b1a_dev; d.b1b = b1b_dev;
    d.w2a = w2a; d.w2b = w2b; d.b2a = b2; d.b2b = b2 + 16;
    d.out_a = out_a_dev; d.out_b = out_b_dev;
    char err[256];
    HeadPairPlan *pl = head_pair_plan_create(d, err, sizeof(err));
    int rc = LWOP_OK;
    if (!pl) {
        rc = fail(h, LWOP_ERR_UNSUPPORTED, err);
    } else {
        cudaError_t e = head_pair_plan_launch(pl, s);
        h->launches += 5;
        if (e != cudaSuccess) rc = fail(h, LWOP_ERR_CUDA, std::string("head pair launch: ") + cudaGetErrorString(e));
        cudaError_t e2 = cudaStreamSynchronize(s);
        if (rc == LWOP_OK && e2 != cudaSuccess)
            rc = fail(h, LWOP_ERR_CUDA, std::string("head pair kernel: ") + cudaGetErrorString(e2));
        head_pair_plan_destroy(pl);
    }
    cudaFree(buf);
    return rc;
}

uint32_t lwop_crc32c(const void *data, size_t n) {
    std::call_once(g_crc_once, crc_init);
    const unsigned char *p = (const unsigned char *)data;
    uint32_t crc = 0xFFFFFFFFu;
    while (n >= 8) {
        uint32_t lo, hi;
        memcpy(&lo, p, 4);
        memcpy(&hi, p + 4, 4);
        lo ^= crc;
        crc = g_crc_table[7][lo & 0xff] ^ g_crc_table[6][(lo >> 8) & 0xff] ^ g_crc_table[5][(lo >> 16) & 0xff] ^
              g_crc_table[4][lo >> 24] ^ g_crc_table[3][hi & 0xff] ^ g_crc_table[2][(hi >> 8) & 0xff] ^
              g_crc_table[1][(hi >> 16) & 0xff] ^ g_crc_table[0][hi >> 24];
        p += 8;
        n -= 8;
    }
    while (n--) crc = g_crc_table[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
    return crc ^ 0xFFFFFFFFu;
}

int lwop_num_steps(const lwop_handle *h) { return h ? (int)h->steps.size() : 0; }

int lwop_step_desc(const lwop_handle *h, int step, int *out10) {
    if (!h || !out10 || step < 0 || step >= (int)h->steps.size()) return LWOP_ERR_INVALID;
    const Step &st = h->steps[step];
    const LayerSpec &l = kLayers[st.layer];
    out10[0] = st.layer;
    out10[1] = st.is_dw ? 1 : 0;
    out10[2] = st.H;
    out10[3] = st.W;
    out10[4] = st.is_dw ? l.cin : (st.layer == 0 ? 3 : l.cin);
    out10[5] = st.is_dw ? l.cin : l.cout;
    out10[6] = st.is_dw ? 3 : l.k;
    out10[7] = (st.is_dw || st.layer == 0) ? l.stride : 1;
    out10[8] = (st.is_dw || l.kind == L_CONV) ? l.dil : 1;
    out10[9] = step_uses_gemm(st) ? 1 : 0;
    if (step_fuses_with_next(h, (size_t)step)) out10[9] = 2;                      // fused dw -> pw kernel runs here
    else if (step > 0 && step_fuses_with_next(h, (size_t)step - 1)) out10[9] = 3; // absorbed into the previous step
    if (step_starts_head_pair(h, (size_t)step)) out10[9] = 4;
    else if (step_absorbed_by_head_pair(h, (size_t)step)) out10[9] = 3;
    return LWOP_OK;
}

int lwop_profile_forward(lwop_handle *h, const float *in_dev, int batch, int mode, float *heat_dev, float *paf_dev,
                         float *step_ms, void *stream) {
    if (!h || !in_dev || !heat_dev || !paf_dev || !step_ms) return fail(h, LWOP_ERR_INVALID, "null argument");
    if (!h->has_weights) return fail(h, LWOP_ERR_NO_WEIGHTS, "lwop_profile_forward before lwop_load_weights");
    if (batch < 1 || batch > h->max_batch) return fail(h, LWOP_ERR_INVALID, "batch out of range");
    cudaStream_t s = (cudaStream_t)stream;
    CU_OK(h, cudaSetDevice(h->device));
    int rc = ensure_arena(h, mode);
    if (rc != LWOP_OK) return rc;
    const size_t n = h->steps.size();
    std::vector<cudaEvent_t> ev(2 * n);
    for (auto &e : ev) CU_OK(h, cudaEventCreate(&e));
    FwdIO io{in_dev, 0, heat_dev, paf_dev, nullptr, nullptr};
    rc = run_forward(h, batch, mode, io, s, ev.data());
    cudaError_t e = cudaStreamSynchronize(s);
    if (rc == LWOP_OK && e != cudaSuccess) rc = fail(h, LWOP_ERR_CUDA, std::string("profile forward: ") + cudaGetErrorString(e));
    if (rc == LWOP_OK)
        for (size_t i = 0; i < n; ++i) cudaEventElapsedTime(&step_ms[i], ev[2 * i], ev[2 * i + 1]);
    for (auto &e2 : ev) cudaEventDestroy(e2);
    return rc;
}

long long lwop_launch_count(const lwop_handle *h, int reset) {
    if (!h) return 0;
    long long v = h->launches;
    if (reset) const_cast<lwop_handle *>(h)->launches = 0;
    return v;
}

}  // extern "C"
