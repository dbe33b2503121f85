// Launch prototypes shared between the kernel translation units and the executor.
#pragma once
#include "lwop_common.cuh"

namespace lwop {

struct ConvArgs {
    const void *in;        // NHWC
    int in_dtype;          // LWOP_DT_*
    const float *w;        // [Cout][k*k][Cin] fp32
    const float *bias;     // [Cout] or nullptr
    const void *residual;  // same dtype/shape as out, added after the activation, or nullptr
    void *out;
    int out_dtype;
    int B, H, W, Cin, Cout, ksize, stride, dil, act;
    int ldi = 0;           // input pixel pitch in elements (0 = Cin)
    int ldo = 0;           // output row pitch in elements (0 = Cout)
    int out_col0 = 0;      // first output column
    int ldr = 0;           // residual row pitch (0 = Cout)
};

struct DwArgs {
    const void *in;
    void *out;
    int dtype;
    const float *w;        // [9][C] fp32
    int B, H, W, C, stride, dil;
    int ksize = 3;         // depthwise kernel edge (the tcgen05-side / vector kernels are 3x3 only; others take the generic kernel)
};

// lwop_direct.cu
cudaError_t launch_conv_direct(const ConvArgs &a, cudaStream_t s);
cudaError_t launch_depthwise_simple(const DwArgs &a, cudaStream_t s);
cudaError_t launch_depthwise_bf16(const DwArgs &a, cudaStream_t s);
// first conv (3 -> 32, stride 2): folded taps [27][32] + bias, passed to the kernel by value (per handle)
struct Conv1Weights {
    float w[27 * 32];
    float b[32];
};
cudaError_t conv1_pack_weights(const float *w_host_n27, const float *b_host, Conv1Weights *out);   // at weight load
cudaError_t launch_conv1_bf16(const Conv1Weights &cw, const void *in, int in_is_u8, void *out, int B, int H, int W,
                              cudaStream_t s);
cudaError_t launch_u8_to_f32(const uint8_t *in, float *out, long long n, cudaStream_t s);
cudaError_t launch_preprocess_bgr(const uint8_t *in, int B, int Hs, int Ws, uint8_t *out, int H, int W, cudaStream_t s);
cudaError_t launch_pack_bf16(const float *src, void *dst, long long rows, int cols, int cols_padded,
                             cudaStream_t s);

// lwop_gemm.cu -- tcgen05 implicit-GEMM conv (1x1 and 3x3, stride 1)
struct GemmPlan;   // opaque: tensor maps + launch geometry for one layer at one (B, H, W)

struct GemmConvDesc {
    const void *in;            // bf16 NHWC [B,H,W,Cin_pad]   (Cin_pad multiple of 32)
    const void *w_packed;      // bf16 [Cout_pad][taps][Cin_pad]
    const float *bias;         // fp32 [Cout_pad]
    const void *residual;      // bf16 [M][ldr] or nullptr
    void *out;                 // bf16 or fp32 [M][ldo] written at column offset out_col0
    float *out2;               // optional fp32 copy [M][Cout] (narrow heads only) or nullptr
    int out_dtype;
    int lda;                   // input pixel pitch in elements
    int B, H, W;
    int Cin_pad, Cout, Cout_pad, ksize, dil, act;
    int ldo, out_col0;         // output row pitch (elements) and first column
    int ldr;                   // residual row pitch (elements)
};

GemmPlan *gemm_plan_create(const GemmConvDesc &d, char *err, size_t errlen);
void gemm_plan_destroy(GemmPlan *p);
// out_override/out2_override (may be nullptr) replace the plan's output pointers for this launch.
cudaError_t gemm_plan_launch(const GemmPlan *p, cudaStream_t s, void *out_override = nullptr,
                             float *out2_override = nullptr);
bool gemm_driver_ready(char *err, size_t errlen);
// launch with the programmatic-dependent-launch attribute (unless LWOP_PDL=0) and an optional cluster width
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kfn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                                Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kfn, static_cast<KArgs>(args)...);
}
int gemm_num_sms();
// cuTensorMapEncodeTiled for a bf16 tensor (rank 2..4), no interleave, zero OOB fill
bool encode_tiled_bf16(CUtensorMap *map, int rank, const void *base, const cuuint64_t *dims,
                       const cuuint64_t *strides_bytes, const cuuint32_t *box, int swizzle_bytes, char *err,
                       size_t errlen);

bool encode_tiled_f32(CUtensorMap *map, int rank, const void *base, const cuuint64_t *dims,
                      const cuuint64_t *strides_bytes, const cuuint32_t *box, char *err, size_t errlen);

// lwop_dw.cu -- TMA-staged depthwise 3x3
struct DwPlan;
bool dw_plan_supported(int C, int stride, int dil);
DwPlan *dw_plan_create(const DwArgs &a, char *err, size_t errlen);
void dw_plan_destroy(DwPlan *p);
cudaError_t dw_plan_launch(const DwPlan *p, cudaStream_t s);

// lwop_sep.cu -- fused depthwise 3x3 -> pointwise 1x1 (tcgen05), stride 1
struct SepPlan;
struct SepConvDesc {
    const void *in;            // bf16 NHWC [B,H,W,C]
    const float *dw_w;         // fp32 [9][C] (device)
    const void *w_packed;      // bf16 [Cout_pad][C] pointwise weights (BN folded)
    const float *bias;         // fp32 [Cout_pad]
    const void *residual;      // bf16 NHWC [B,H,W,ldr] or nullptr (added after the activation)
    void *out;                 // bf16 NHWC [B,H,W,ldo], written at column offset out_col0
    int B, H, W, C, Cout_pad, dil, act;
    int ldo, out_col0, ldr;
};
bool sep_plan_supported(int C, int Cout_pad, int stride, int dil);
SepPlan *sep_plan_create(const SepConvDesc &d, char *err, size_t errlen);
void sep_plan_destroy(SepPlan *p);
cudaError_t sep_plan_launch(const SepPlan *p, cudaStream_t s);

// lwop_heads.cu -- fused pair of 2-layer 1x1 heads sharing one input (back-to-back tcgen05 GEMMs)
struct HeadPairPlan;
struct HeadPairDesc {
    const void *in;              // bf16 [M][lda], Cin = 128 channels used
    int lda, Cin;
    long long M;
    int hidden;                  // width of each head's hidden layer (512 or 128)
    const void *w1a, *w1b;       // bf16 [hidden][Cin]
    const float *b1a, *b1b;      // fp32 [hidden]
    const void *w2a, *w2b;       // bf16 [16][hidden], [32][hidden] (rows 14.., 26.. zero)
    const float *b2a, *b2b;      // fp32 [16], [32]
    void *cat;                   // optional bf16 [M][ldc]: head a at columns cat_col_a.., head b at cat_col_b..
    int ldc, cat_col_a, cat_col_b;
    float *out_a, *out_b;        // optional fp32 [M][14], [M][26]
};
bool head_pair_supported(int cin, int hidden);
HeadPairPlan *head_pair_plan_create(const HeadPairDesc &d, char *err, size_t errlen);
void head_pair_plan_destroy(HeadPairPlan *p);
cudaError_t head_pair_plan_launch(const HeadPairPlan *p, cudaStream_t s, float *out_a_override = nullptr,
                                  float *out_b_override = nullptr);

// lwop_draw.cu -- drawing half of plot_pose (circles + thick lines, OpenCV's pixel sets) through a priority map
// prio: int32 [B][H][W], all -1 on entry and again on exit
cudaError_t launch_draw_poses(const uint8_t *img, int B, int H, int W, const int *counts, const double *joints,
                              const double *persons, int *prio, uint8_t *canvas, cudaStream_t s, long long *launches);

// lwop_decode.cu
struct DecodeArgs {
    const float *heat;   // [B,h,w,14]
    const float *paf;    // [B,h,w,26]
    int B, h, w, H, W;
    float thre1;
    double thre2;
    lwop_decode_tables t;
    // scratch owned by the handle
    float *peaks;        // [B][14][LWOP_MAX_PEAKS][4]  (x, y, score, unused) in kept order
    int *peak_counts;    // [B][16]
    double *work;        // [B][13*LWOP_MAX_PEAKS][16] working list of persons (grouping)
};
cudaError_t launch_decode(const DecodeArgs &a, cudaStream_t s, long long *launches);
size_t decode_scratch_peaks_bytes(int B);
cudaError_t launch_decode_peaks(const float *heat, int B, int h, int w, double factor, float thre1, int mode,
                                float *peaks, int *peak_counts, int *counts, cudaStream_t s);
cudaError_t launch_decode_limbs(const float *paf, int upsampled, int B, int h, int w, int H, int W, double thre2,
                                int num_intermed_pts, double max_dist_thresh, const float *peaks, const int *peak_counts,
                                double *limbs, int *counts, cudaStream_t s);
cudaError_t launch_decode_group(const float *peaks, const int *peak_counts, const double *limbs, int *counts,
                                double *joints, double *persons, float *keypoints, double *work, int B, cudaStream_t s);
size_t decode_scratch_work_bytes(int B);
// compaction of the fixed-capacity tables: offsets[b][4] = exclusive prefix of (joints, persons, conns, 0)
cudaError_t launch_decode_pack(const lwop_decode_tables &t, int B, int *offsets, double *joints_packed,
                               double *limbs_packed, double *persons_packed, float *keypoints_packed,
                               cudaStream_t s, long long *launches, long long cap_joints = (1ll << 60),
                               long long cap_limbs = (1ll << 60), long long cap_persons = (1ll << 60));

}  // namespace lwop
