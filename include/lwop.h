/*
 * lwop.h -- C ABI of the B200-native lightweight_openpose inference hot path.
 *
 * The reference (murdockhou/lightweight_openpose) has no FFI: its boundary is
 * the Python function API
 *     lightweight_openpose(inputs, num_joints, num_pafs, is_training)   reference src/lightweight_openpose.py:8,48
 *     conv / conv_dw / conv_dw_no_bn / RefinementStageBlock             reference src/net_factory.py:4,14,23,33
 *     decode_pose(img_orig, param, heatmaps, pafs)                      reference src/pose_decode.py:460,517
 * driven by test&eval/model_test.py:42,68,71 (sess.run + decode per frame).
 * The Python package lightweight_openpose_b200.src mirrors those functions and
 * binds the entry points below with ctypes.
 *
 * Conventions
 *  - plain C types only; every pointer argument marked "dev" is a CUDA device
 *    pointer on the handle's device, "host" is ordinary (ideally pinned) memory;
 *  - tensors are NHWC, kernels arrive pre-folded as [Cout][kh*kw][Cin] float32;
 *  - every function returns 0 (LWOP_OK) or a negative error code and never
 *    throws; lwop_last_error() returns a human-readable message for the last
 *    failure on that handle (or the last global failure when handle is NULL);
 *  - `stream` is a cudaStream_t passed as void* (NULL = the legacy default
 *    stream); calls on one handle must be serialised by the caller;
 *  - the caller owns every input and output buffer; the handle owns weights,
 *    the activation arena, TMA descriptors and CUDA graphs.
 */
#ifndef LWOP_H_
#define LWOP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LWOP_ABI_VERSION 2

/* error codes */
#define LWOP_OK                 0
#define LWOP_ERR_INVALID       -1   /* bad argument */
#define LWOP_ERR_CUDA          -2   /* CUDA runtime / driver failure */
#define LWOP_ERR_NO_WEIGHTS    -3   /* forward called before lwop_load_weights */
#define LWOP_ERR_OVERFLOW      -4   /* a decode table capacity was exceeded */
#define LWOP_ERR_UNSUPPORTED   -5   /* shape / mode not supported by the kernels */
#define LWOP_ERR_NO_DEVICE     -6   /* no sm_100 device */

/* arithmetic mode of the forward: bf16 tensor-core path, or the fp32 check path
 * (CUDA-core fp32, used to bound the bf16 error: BASELINE.json tolerances) */
#define LWOP_MODE_BF16         0
#define LWOP_MODE_FP32_CHECK   1
#define LWOP_MODE_BF16_DIRECT  2   /* bf16 storage, CUDA-core kernels: debugging ladder for the tcgen05 path */

#define LWOP_ACT_NONE 0
#define LWOP_ACT_RELU 1            /* tf.nn.relu, net_factory.py:11,19 */
#define LWOP_ACT_ELU  2            /* tf.nn.elu,  net_factory.py:29   */

#define LWOP_DT_F32   0
#define LWOP_DT_BF16  1

/* lwop_create flags */
#define LWOP_FLAG_NO_GRAPH      1u /* launch kernels directly instead of replaying a CUDA graph */
#define LWOP_FLAG_NO_FUSE_SEP   2u /* separable layers as two kernels (depthwise, then 1x1 GEMM) instead of the fused one */
#define LWOP_FLAG_NO_FUSE_HEADS 4u /* the 2-layer head pairs as four GEMM launches instead of one back-to-back kernel */

/* network constants (reference src/train_config.py:25-26,31) */
#define LWOP_NUM_JOINTS   14
#define LWOP_NUM_PAFS     26
#define LWOP_NUM_LIMBS    13
#define LWOP_STRIDE        8

/* decode table capacities (per image) */
#define LWOP_MAX_PEAKS    256      /* peaks per joint channel: above the geometric maximum of the radius-5 search on
                                      maps up to ~4400 pixels (46 x 82: ~175), so BASELINE's frame sizes cannot overflow */
#define LWOP_MAX_JOINTS  3584      /* 14 * 256 rows of the joint list */
#define LWOP_MAX_PERSONS  256
#define LWOP_COUNTS        32      /* int32 slots per image, see below */
/* counts[b][0] = rows in the joint list, [1] = persons, [2] = overflow bitmask
 * (1: peaks/channel, 2: persons), [3] = total connections,
 * [4..17] = peaks per joint channel, [18..30] = connections per limb type,
 * [31] = internal (image handed from the register-resident grouping kernel to the general one) */

typedef struct lwop_handle lwop_handle;

int         lwop_abi_version(void);
const char *lwop_error_string(int code);
const char *lwop_last_error(const lwop_handle *h);

/* Replaces graph construction: tf.placeholder + lightweight_openpose(...) in
 * test&eval/model_test.py:39-42.  The handle is sized for images up to
 * [max_batch, height, width, 3]; smaller batches may be passed to lwop_forward. */
int lwop_create(int device, int max_batch, int height, int width, unsigned flags, lwop_handle **out);
int lwop_destroy(lwop_handle *h);

/* Replaces tf.train.Saver().restore (test&eval/model_test.py:43,49).  `blob` is
 * the float32 concatenation produced by lightweight_openpose_b200.graph.pack_blob
 * (per layer: depthwise [9][Cin] if separable, kernel [Cout][taps][Cin] with BN
 * folded, bias [Cout]); its length must equal lwop_weights_num_floats(). */
size_t lwop_weights_num_floats(void);
int    lwop_load_weights(lwop_handle *h, const float *host_blob, size_t num_floats);

/* Replaces sess.run([cpm, paf], feed_dict) (test&eval/model_test.py:68,92).
 * in_dev: [batch, H, W, 3] float32 RGB in [0,1];  heat_dev: [batch, H/8, W/8, 14]
 * float32;  paf_dev: [batch, H/8, W/8, 26] float32.  stage1_heat_dev/stage1_paf_dev
 * are optional (NULL) outputs of the initial stage (lightweight_openpose.py:32-35). */
int lwop_forward(lwop_handle *h, const float *in_dev, int batch, int mode,
                 float *heat_dev, float *paf_dev,
                 float *stage1_heat_dev, float *stage1_paf_dev, void *stream);

/* Same, from uint8 frames: does the reference's `img / 255.` (model_test.py:65)
 * on the device.  in_dev: [batch, H, W, 3] uint8 RGB. */
int lwop_forward_u8(lwop_handle *h, const uint8_t *in_dev, int batch, int mode,
                    float *heat_dev, float *paf_dev, void *stream);

/* Frame pre-processing of the reference drivers (test&eval/model_test.py:63-64, model_json.py:68-69):
 * cv2.cvtColor(BGR2RGB) + cv2.resize(frame, (dst_w, dst_h)) (8-bit INTER_LINEAR, OpenCV's fixed-point scheme) on
 * the device.  in_dev uint8 [batch][src_h][src_w][3] BGR -> out_dev uint8 [batch][dst_h][dst_w][3] RGB, which is what
 * lwop_forward_u8 (the `/ 255.` of :65) consumes. */
int lwop_preprocess_bgr_u8(lwop_handle *h, const uint8_t *in_dev, int batch, int src_h, int src_w,
                           uint8_t *out_dev, int dst_h, int dst_w, void *stream);

/* Decode tables, caller-allocated device memory (sizes for `batch` images):
 *   counts    int32  [batch][LWOP_COUNTS]
 *   joints    double [batch][LWOP_MAX_JOINTS][6]   (x, y, score, id, 0, joint_type)   pose_decode.py:481-482
 *   limbs     double [batch][13][LWOP_MAX_PEAKS][5] (src_id, dst_id, score, i, j)     pose_decode.py:287
 *   persons   double [batch][LWOP_MAX_PERSONS][16]                                    pose_decode.py:369-379
 *   keypoints float  [batch][LWOP_MAX_PERSONS][42]                                    pose_decode.py:403,516 */
typedef struct lwop_decode_tables {
    int32_t *counts;
    double  *joints;
    double  *limbs;
    double  *persons;
    float   *keypoints;
} lwop_decode_tables;

/* Replaces decode_pose (src/pose_decode.py:460-517) minus the canvas drawing,
 * batched.  heat_dev [batch,h,w,14], paf_dev [batch,h,w,26] float32; image size
 * (img_h, img_w) gives the up-sampling factor img_h / h exactly as :462 does.
 * Returns LWOP_ERR_OVERFLOW only after a later lwop_decode_status(). */
int lwop_decode(lwop_handle *h, const float *heat_dev, const float *paf_dev, int batch,
                int map_h, int map_w, int img_h, int img_w, float thre1, double thre2,
                const lwop_decode_tables *tables_dev, void *stream);

/* lwop_forward (or lwop_forward_u8 when is_u8) followed by lwop_decode of its maps as ONE stream operation: in bf16
 * mode the forward kernels and the decode chain are replayed from a single CUDA graph, the decode kernels chained to
 * the last head kernel with programmatic dependent launches (what the per-frame loop body of
 * test&eval/model_test.py:68,71 does, batched, without leaving the device).  Maps in heat_dev / paf_dev, results in
 * tables_dev; image size = the handle's (H, W). */
int lwop_forward_decode(lwop_handle *h, const void *in_dev, int is_u8, int batch, int mode, float thre1, double thre2,
                        float *heat_dev, float *paf_dev, const lwop_decode_tables *tables_dev, void *stream);

/* ---- the decode stages, separately callable (what src/pose_decode.py also exports) ----
 * Peak table layout shared by the three calls (device memory, caller-allocated):
 *   peaks        float [batch][14][LWOP_MAX_PEAKS][4]   (x, y, score, unused) per joint channel, in kept order
 *   peak_counts  int32 [batch][16]                      peaks per joint channel
 * Joint ids are the running counter over (channel, peak) that NMS assigns (pose_decode.py:135,182-183). */
#define LWOP_PEAKS_NO_REFINE 1   /* NMS(bool_refine_center=False): low-res peak snapped to the up-scaled grid (:176-182) */
#define LWOP_PEAKS_CROSS     2   /* peak search = find_peaks (:37-49, 3x3 cross maximum filter) instead of find_peaks_v2 */
#define LWOP_PEAKS_GAUSS     4   /* NMS(bool_gaussian_filt=True): scipy gaussian_filter(sigma=3) of the up-sampled patch before
                                    the argmax (:163-164); needs 5 * factor <= 48 */

/* NMS (pose_decode.py:107-185): heat_dev [batch][h][w][14] float32, `factor` = upsampFactor.  Also zeroes
 * counts_dev [batch][LWOP_COUNTS] and sets its overflow bit (counts[b][2] & 1) if a channel has more than
 * LWOP_MAX_PEAKS peaks. */
int lwop_nms(lwop_handle *h, const float *heat_dev, int batch, int map_h, int map_w, double factor, float thre1,
             int mode, float *peaks_dev, int32_t *peak_counts_dev, int32_t *counts_dev, void *stream);

/* find_connected_joints (pose_decode.py:188-301).  paf_dev is either the network's PAF map [batch][h][w][26]
 * (paf_upsampled = 0: the bicubic up-sampling of decode_pose :487 is evaluated on the fly) or an already
 * up-sampled [batch][H][W][26] map (paf_upsampled = 1, what the reference function is handed).
 * num_intermed_pts (1..64) and max_dist_thresh are the reference's keyword arguments of the same names
 * (defaults 10 and 1; lwop_decode uses the defaults, as decode_pose does).
 * Writes limbs_dev [batch][13][LWOP_MAX_PEAKS][5] and counts_dev[b][18 + limb_type]. */
int lwop_find_connected_joints(lwop_handle *h, const float *paf_dev, int paf_upsampled, int batch, int map_h, int map_w,
                               int img_h, int img_w, double thre2, int num_intermed_pts, double max_dist_thresh,
                               const float *peaks_dev, const int32_t *peak_counts_dev, double *limbs_dev,
                               int32_t *counts_dev, void *stream);

/* joint list + group_limbs_of_same_person + keypoint table (pose_decode.py:481-482, :304-396, :403-429):
 * fills joints / persons / keypoints of `tables_dev` and counts[b][0], [1], [3] from the peak and limb tables. */
int lwop_group_limbs(lwop_handle *h, const float *peaks_dev, const int32_t *peak_counts_dev, int batch,
                     const lwop_decode_tables *tables_dev, void *stream);

/* Drawing half of plot_pose (src/pose_decode.py:431-436) on the device: for every person of `tables_dev` (in table
 * order) and limb type (in order) whose two joints are set, two filled white circles of radius 3 and a line of
 * thickness 2 in colors[person % 16] (:26-31), painted in the reference's order onto a copy of the frame.  The pixel
 * sets are those of cv2.circle / cv2.line (OpenCV 4.x, LINE_8), so the canvas is byte-identical to the reference's.
 * img_dev, canvas_dev: uint8 [batch][img_h][img_w][3] (canvas_dev may be img_dev: draw in place); reads counts[b][1],
 * joints and persons of tables_dev (lwop_decode / lwop_forward_decode / lwop_group_limbs output). */
int lwop_draw_poses(lwop_handle *h, const uint8_t *img_dev, int batch, int img_h, int img_w,
                    const lwop_decode_tables *tables_dev, uint8_t *canvas_dev, void *stream);

/* Host-side result of lwop_infer_host / lwop_decode_fetch: rows of all images
 * concatenated in image order (per image: counts[b][0] joint rows, counts[b][1]
 * person / keypoint rows, counts[b][18+t] limb rows per limb type t in order). */
typedef struct lwop_packed_tables {
    int32_t *counts;        /* [batch][LWOP_COUNTS] */
    double  *joints;        /* [joints_cap][6]  */
    double  *limbs;         /* [limbs_cap][5]   */
    double  *persons;       /* [persons_cap][16] */
    float   *keypoints;     /* [persons_cap][42] */
    size_t   joints_cap, limbs_cap, persons_cap;   /* capacities in rows */
} lwop_packed_tables;

/* Compacts device tables produced by lwop_decode and copies them to the host
 * (two device->host phases: counts, then exactly the used rows).  Synchronises
 * `stream`.  Returns LWOP_ERR_OVERFLOW if a capacity (device table or host
 * buffer) was exceeded. */
int lwop_decode_fetch(lwop_handle *h, const lwop_decode_tables *tables_dev, int batch,
                      const lwop_packed_tables *out_host, void *stream);

/* Whole path with HOST buffers in one call (what model_test.py's loop body does
 * per frame, batched): H2D of the frames, forward, decode, D2H of the compacted
 * result tables.  images_host: [batch,H,W,3] float32 (or uint8 when is_u8).
 * heat_host / paf_host may be NULL when the maps themselves are not wanted.
 * The batch flows through a two-slot pipeline in chunks ("chunk" option): a
 * private copy stream uploads chunk k+1 while `stream` computes chunk k, so with
 * pinned host memory the PCIe transfer hides behind the kernels.
 * Synchronises `stream` before returning. */
int lwop_infer_host(lwop_handle *h, const void *images_host, int is_u8, int batch, int mode,
                    float thre1, double thre2, const lwop_packed_tables *out_host,
                    float *heat_host, float *paf_host, void *stream);

/* Asynchronous form of lwop_infer_host, for callers that keep batches in flight (a serving loop): _begin enqueues
 * upload, forward, decode, compaction and the download of the result tables into `out_host` and returns a ticket
 * without waiting; _end(ticket) waits for that batch and reports overflow.  With two batches in flight the upload
 * of batch k+1 overlaps the kernels of batch k.  The result tables are downloaded at the capacities given in
 * `out_host` (the used row counts are only known afterwards, from counts); LWOP_ERR_OVERFLOW from _end means the
 * capacities were too small for this batch: enlarge them and run the batch again.  `out_host` and `images_host`
 * must stay valid (and should be pinned) until _end returns.  lwop_infer_host == _begin + _end. */
#define LWOP_MAX_INFLIGHT 2
int lwop_infer_host_begin(lwop_handle *h, const void *images_host, int is_u8, int batch, int mode,
                          float thre1, double thre2, const lwop_packed_tables *out_host,
                          float *heat_host, float *paf_host, void *stream, int *ticket);
int lwop_infer_host_end(lwop_handle *h, int ticket);

/* Tunables: "chunk" = images per pipeline stage of lwop_infer_host (0 = default: 64 for float32 frames, 128 for
 * uint8; the asynchronous entry submits the whole batch as one chunk).  "copy_stream" = a caller-owned cudaStream_t
 * (passed as an integer) that carries this handle's uploads instead of its private copy stream: handles of one device
 * that share it upload in submission order, each batch at the full PCIe rate (0 = private stream again). */
int lwop_set_option(lwop_handle *h, const char *name, long long value);

/* ---- per-op entries mirroring src/net_factory.py (unit tests, eager mode) ---- */

/* conv(): dense conv2d 'same' + bias + activation (+ residual added after the
 * activation, tf.add in net_factory.py:37 / lightweight_openpose.py:27).
 * w_dev float32 [Cout][k*k][Cin], bias_dev float32 [Cout] or NULL. */
int lwop_conv2d(lwop_handle *h, int mode, const void *in_dev, int in_dtype,
                int batch, int height, int width, int cin,
                const float *w_dev, const float *bias_dev, int cout, int ksize, int stride, int dilation,
                int act, const void *residual_dev, void *out_dev, int out_dtype, void *stream);

/* First layer alone: conv(inputs, 32, stride=2, bias=False) + folded batch norm + ReLU (lightweight_openpose.py:10,
 * net_factory.py:4-12) in its production kernel (CUDA cores, packed fp32 FMA).  in_dev: NHWC frames, float32 in [0,1]
 * or uint8 (in_is_u8: the `/ 255.` of model_test.py:65 runs inside the kernel); w_host float32 [32][27] (tap index
 * (ky*3 + kx)*3 + c) and b_host float32 [32] in HOST memory (BN folded by the caller); out_dev bf16 NHWC
 * [batch][ceil(H/2)][ceil(W/2)][32]. */
int lwop_first_conv(lwop_handle *h, const void *in_dev, int in_is_u8, int batch, int height, int width,
                    const float *w_host, const float *b_host, void *out_dev, void *stream);

/* depthwise half of tf.layers.separable_conv2d (net_factory.py:15,25):
 * 3x3, channel multiplier 1, 'same', no bias, no activation.  w_dev float32 [9][C]. */
int lwop_depthwise3x3(lwop_handle *h, int mode, const void *in_dev, int dtype,
                      int batch, int height, int width, int channels,
                      const float *w_dev, int stride, int dilation, void *out_dev, void *stream);

/* The same for any kernel_size (conv_dw / conv_dw_no_bn take it as an argument, net_factory.py:14,23, although the
 * network only uses 3): w_dev float32 [ksize*ksize][C]; 'same' padding as TensorFlow splits it (smaller half first).
 * ksize == 3 is lwop_depthwise3x3; other sizes run a generic CUDA-core kernel (fp32 accumulation) in every mode. */
int lwop_depthwise2d(lwop_handle *h, int mode, const void *in_dev, int dtype,
                     int batch, int height, int width, int channels,
                     const float *w_dev, int ksize, int stride, int dilation, void *out_dev, void *stream);

/* conv_dw() / conv_dw_no_bn(): tf.layers.separable_conv2d (depthwise 3x3 then pointwise 1x1, nothing in
 * between) + folded batch norm / bias + activation (net_factory.py:15-19,25-29), optional residual added after
 * the activation (lightweight_openpose.py:27).  LWOP_MODE_BF16: ONE fused kernel (CUDA-core depthwise feeding
 * the tcgen05 GEMM through shared memory); needs bf16 tensors, stride 1, dilation 1 or 2 and
 * (cin, cout) in {(32, 64), (64k, 128), (64k, 256m)}, else LWOP_ERR_UNSUPPORTED.
 * dw_w_dev float32 [9][cin], pw_w_dev float32 [cout][cin], bias_dev float32 [cout] or NULL;
 * in_dev / residual_dev / out_dev bf16 NHWC. */
int lwop_separable_conv2d(lwop_handle *h, int mode, const void *in_dev, int batch, int height, int width, int cin,
                          const float *dw_w_dev, const float *pw_w_dev, const float *bias_dev, int cout,
                          int dilation, int act, const void *residual_dev, void *out_dev, void *stream);

/* Two 2-layer 1x1-conv heads on one input, as ONE kernel (lightweight_openpose.py:32-35 and :43-46):
 *   out_a = (relu(x W1a^T + b1a)) W2a^T + b2a   [M][14],   out_b = (relu(x W1b^T + b1b)) W2b^T + b2b   [M][26].
 * x_dev bf16 [M][128]; weights float32 on the device: w1 [hidden][128], b1 [hidden], w2a [14][hidden], b2a [14],
 * w2b [26][hidden], b2b [26]; hidden in {128, 512}; outputs float32. */
int lwop_head_pair(lwop_handle *h, const void *x_dev, long long m, int hidden,
                   const float *w1a_dev, const float *b1a_dev, const float *w2a_dev, const float *b2a_dev,
                   const float *w1b_dev, const float *b1b_dev, const float *w2b_dev, const float *b2b_dev,
                   float *out_a_dev, float *out_b_dev, void *stream);

/* output size of a 'same' conv: ceil(in / stride) */
int lwop_same_out(int in, int stride);

/* CRC32C (Castagnoli) for TensorBundle verification; host function. */
uint32_t lwop_crc32c(const void *data, size_t nbytes);

/* Layer program introspection and per-kernel timing (bench.py's live roofline).
 * lwop_step_desc fills out[10] = {layer, is_depthwise, in_h, in_w, cin, cout, ksize, stride, dilation, kernel} with
 * kernel 0 = CUDA-core kernel, 1 = tcgen05 GEMM, 2 = fused depthwise+pointwise tcgen05 kernel (this depthwise step
 * and the pointwise step after it), 3 = no launch (absorbed into an earlier fused step), 4 = fused head pair
 * (this step and the three after it: two 2-layer 1x1 heads as back-to-back GEMMs).
 * lwop_profile_forward runs one forward WITHOUT the CUDA graph, bracketing every
 * kernel with CUDA events on `stream`, synchronises and writes the per-step
 * durations in milliseconds to step_ms[lwop_num_steps()]. */
int lwop_num_steps(const lwop_handle *h);
int lwop_step_desc(const lwop_handle *h, int step, int *out10);
int lwop_profile_forward(lwop_handle *h, const float *in_dev, int batch, int mode,
                         float *heat_dev, float *paf_dev, float *step_ms, void *stream);

/* number of kernel launches issued by this handle since creation / last reset */
long long lwop_launch_count(const lwop_handle *h, int reset);

#ifdef __cplusplus
}
#endif
#endif /* LWOP_H_ */
