/* phdfx.h — C ABI of the B200-native ResNet-50 frame-feature extractor (libphdfx.so).
 *
 * The reference (ferreiraluisa/implementation-phd-lab-vision) has no FFI: its seam is one Python callable,
 *     feats = backbone(x).flatten(1).view(Bv, T, -1)          src/preprocess_resnet_features.py:296
 * fed by the CPU crop/resize/normalise in src/dataset.py:141-152,242-245.  These entry points are what a binding for
 * that seam calls (INTEGRATION.md shows the ctypes stub).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative phdfx_status otherwise; the message is phdfx_last_error().
 *   - all d_* pointers are DEVICE pointers owned by the caller and must stay alive until the stream work is done.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  No call synchronises the device;
 *     every call is CUDA-graph capturable.
 *   - a handle is bound to one device and is not thread-safe.  There is NO CPU fallback: creation fails on a device
 *     that is not compute capability 10.x.
 */
#ifndef PHDFX_H_
#define PHDFX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHDFX_VERSION 103 /* major*100 + minor */

typedef struct phdfx phdfx_t;

typedef enum phdfx_status {
  PHDFX_OK = 0,
  PHDFX_ERR_INVALID = -1,   /* bad argument / shape / n > max_frames */
  PHDFX_ERR_ARCH = -2,      /* device is not sm_100 */
  PHDFX_ERR_CUDA = -3,      /* CUDA runtime / driver error */
  PHDFX_ERR_STATE = -4,     /* weights not loaded */
} phdfx_status;

typedef enum phdfx_layer_kind {
  PHDFX_CONV = 0,    /* implicit-GEMM conv: 1x1 or 3x3, stride 1 or 2, Cin % 64 == 0, Cout % 64 == 0 */
  PHDFX_STEM = 1,    /* 7x7 stride-2 pad-3 conv, 3 -> 64 channels, on the NHWC4p input layout */
  PHDFX_MAXPOOL = 2, /* 3x3 stride-2 pad-1 max-pool */
  PHDFX_STEM_POOL = 3, /* stem conv + ReLU + 3x3/2 max-pool fused: NHWC4p in, [n][56][56][64] out; weights in the
                          no-swizzle smem image layout [7][4][64][8] + stacked row pairs [5][4][128][8]
                          (phdfx/weights.py: pack_stem_pool) */
} phdfx_layer_kind;

/* One entry of the execution list handed to phdfx_load_weights.  Mirrors one conv(+folded BN)(+ReLU)(+residual)
 * of torchvision's ResNet (models/resnet.py:143-163, :268-271).  Buffers are ids into the library-owned activation
 * arena; buffer 0 is the network input in NHWC4p layout. */
typedef struct phdfx_layer_desc {
  int32_t kind;            /* phdfx_layer_kind */
  int32_t cin, cout;
  int32_t r, s;            /* filter height / width */
  int32_t stride, pad;
  int32_t hin, win;        /* input spatial size */
  int32_t relu;            /* ReLU at the end of the epilogue */
  int32_t in_buf, out_buf; /* arena buffer ids */
  int32_t res_buf;         /* residual added before ReLU; -1 = none */
  int32_t gap;             /* 1: fuse AdaptiveAvgPool2d(1) (resnet.py:278): emit fp32 [n, cout] features */
  /* Optional second 1x1 source accumulated into the same output (K concatenated): the block's down-sample branch
   * (resnet.py:157-158, 239-243) fused into conv3:  out = relu(conv3(t2) + downsample(x) + b3 + bd).
   * in2_buf = -1: none.  The layer itself must then be a 1x1 stride-1 conv; weights are [cout][cin + cin2]. */
  int32_t in2_buf, cin2, stride2, hin2; /* second input buffer, its channels, its stride (1 or 2), its spatial size */
  int64_t w_off;           /* element offset of this layer's packed weights [cout][r][s][cin] (stem: [7][64][32]) */
  int64_t b_off;           /* element offset of this layer's folded-BN bias [cout] */
} phdfx_layer_desc;

/* Geometry of the NHWC4p network-input layout: bf16 [n][224][232][4]; pixel column = w + 4, channel 3 = 0. */
#define PHDFX_IMG 224
#define PHDFX_IN_WPAD 232
#define PHDFX_IN_LPAD 4
#define PHDFX_IN_CPAD 4
#define PHDFX_FEAT_DIM 2048

int phdfx_version(void);
const char* phdfx_last_error(const phdfx_t* h); /* h may be NULL: last error of the calling thread */

/* Replaces `backbone.to(device).eval()` (preprocess_resnet_features.py:207-209): binds a device, sizes the arena. */
int phdfx_create(phdfx_t** h, int device_ordinal, int max_frames);
int phdfx_destroy(phdfx_t* h);

/* Host pointers.  packed_bf16: n_weights bf16 values; bias_f32: n_bias floats; layers: execution list.
 * BN folding and packing are done by the host (phdfx/weights.py); the library copies to the device and owns them. */
int phdfx_load_weights(phdfx_t* h, const void* packed_bf16, int64_t n_weights, const float* bias_f32,
                       int64_t n_bias, const phdfx_layer_desc* layers, int n_layers);

/* K1 — replaces src/dataset.py:141-152 (_crop_and_resize_video_uint8) + :242-245 (Normalize).
 * d_frames_hwc: uint8 [n][H][W][3]; d_boxes: int32 [n][4] (top,left,h,w) or NULL (= whole frame);
 * flip_w != 0 mirrors the output horizontally (the reference's hflip variant, dataset.py:158-186);
 * d_out_nhwc4p: bf16 NHWC4p or NULL (= the arena's input buffer). */
int phdfx_preprocess_u8(phdfx_t* h, const uint8_t* d_frames_hwc, int n, int H, int W, const int32_t* d_boxes,
                        int flip_w, void* d_out_nhwc4p, void* stream);

/* K1 with the reference's colour-jitter augmentation in front of Normalize (src/dataset.py:188-198, the `cjitter`
 * variant of --augment: torchvision v2 ColorJitter(brightness=0.3, contrast=0.3, saturation=0.2, hue=0.05) on the
 * resized [0,1] clip, one parameter draw per clip).  d_jitter: device float [n][12], one row per FRAME (repeat a
 * clip's draw for its frames): [0..3] the four ops in application order (0 brightness, 1 contrast, 2 saturation,
 * 3 hue — ColorJitter.make_params' fn_idx), [4] brightness factor, [5] contrast factor, [6] (float)(1.0 - contrast),
 * [7] saturation factor, [8] (float)(1.0 - saturation), [9] hue factor, [10..11] unused.  Two launches (the contrast
 * op needs the frame's mean grey level first). */
int phdfx_preprocess_u8_jitter(phdfx_t* h, const uint8_t* d_frames_hwc, int n, int H, int W, const int32_t* d_boxes,
                               int flip_w, const float* d_jitter, void* d_out_nhwc4p, void* stream);

/* Seam A repack: fp32 NCHW [n,3,224,224], already normalised (the tensor the reference feeds to backbone(), :295)
 * -> NHWC4p bf16 (NULL = arena input buffer). */
int phdfx_nchw_f32_to_nhwc_bf16(phdfx_t* h, const float* d_x_nchw, int n, void* d_out_nhwc4p, void* stream);

/* The trunk — replaces backbone(x).flatten(1) (:296).  d_in_nhwc4p NULL = arena input buffer.
 * d_feats: fp32 [n][2048]. */
int phdfx_forward(phdfx_t* h, const void* d_in_nhwc4p, int n, float* d_feats, void* stream);

/* Profiling hook (BASELINE config 3): phdfx_forward with a CUDA event before every launch and after the last one.
 * Synchronises the stream (NOT graph-capturable), writes the elapsed milliseconds of each launch — in situ, i.e. with
 * the L2 contents the previous launch left — into the HOST array ms_per_launch[cap] and returns the number of
 * launches (> 0), or a negative status.  The events serialise the launches (no programmatic overlap). */
int phdfx_forward_timed(phdfx_t* h, const void* d_in_nhwc4p, int n, float* d_feats, void* stream,
                        float* ms_per_launch, int cap);

/* Seam B: preprocess + trunk. */
int phdfx_extract_u8(phdfx_t* h, const uint8_t* d_frames_hwc, int n, int H, int W, const int32_t* d_boxes,
                     int flip_w, float* d_feats, void* stream);

/* Seam B with colour jitter (phdfx_preprocess_u8_jitter + trunk). */
int phdfx_extract_u8_jitter(phdfx_t* h, const uint8_t* d_frames_hwc, int n, int H, int W, const int32_t* d_boxes,
                            int flip_w, const float* d_jitter, float* d_feats, void* stream);

/* Per-layer hook (parity tests, ncu, per-layer benchmark).  d_in / d_residual / d_out use the layer's own
 * layouts: NHWC bf16 activations (NHWC4p for the stem input); for a gap layer d_out is fp32 [n][cout]. */
int phdfx_run_layer(phdfx_t* h, int layer_id, const void* d_in, const void* d_residual, void* d_out, int n,
                    void* stream);
/* Same, for layers with a second input (in2_buf >= 0): d_in2 = that input in NHWC bf16. */
int phdfx_run_layer2(phdfx_t* h, int layer_id, const void* d_in, const void* d_in2, const void* d_residual,
                     void* d_out, int n, void* stream);

/* Fused spans.  phdfx_forward runs  conv2 (3x3) -> conv3 (+ identity | fused down-sample) [-> the next block's conv1]
 * (resnet.py:150-161 and the following :146-148) as ONE launch (bottleneck_chain_sm100.cuh) for the stride-1 blocks of
 * layer1 (56x56, width 64; the next conv1 rides along) and layer2 (28x28, width 128) whenever the execution list has
 * that pattern with distinct buffers; PHDFX_NO_CHAIN=1 in the environment at phdfx_create keeps the per-conv kernels.
 * phdfx_chain_span: number of list entries the launch starting at layer_id covers (0 = no fused launch starts there).
 * phdfx_run_chain: that launch on caller buffers (parity tests / per-kernel benchmark): d_t1 = conv2's input
 * [n][H][H][width]; d_x_or_res = the down-sample source [n][H][H][64] when conv3 carries a second input, else the
 * identity residual [n][H][H][4*width]; d_out = block output [n][H][H][4*width]; d_t1_next = the trailing conv1's
 * output [n][H][H][cout] (NULL when the span is 2).  All NHWC bf16. */
int phdfx_chain_span(const phdfx_t* h, int layer_id);
int phdfx_run_chain(phdfx_t* h, int first_layer_id, const void* d_t1, const void* d_x_or_res, void* d_out,
                    void* d_t1_next, int n, void* stream);

/* Execution schedule: frame waves.  The reference runs every layer over the whole batch before the next one starts
 * (torchvision models/resnet.py:266-279 under backbone(x), preprocess_resnet_features.py:296); at batch 256 the
 * activations of the early stages (1.6 MB per frame and layer in layer1) then make a full trip through HBM between any
 * two launches.  Frames are independent, so the list may instead be cut into consecutive STAGES, each run in waves of
 * `wave_frames` frames, depth first: stage s processes a wave as soon as stage s-1 has produced its frames, and what one
 * launch writes is what the next one reads while it is still in the 126 MB L2.  Results are bit-identical for every
 * schedule (each frame sees the same kernels, tiles and K order).
 *   first_layer[s]  first execution-list entry of stage s (first_layer[0] = 0, ascending; a stage may not start inside a
 *                   fused conv2 -> conv3 -> conv1 launch, see phdfx_chain_span); stage s ends where s+1 starts
 *   wave_frames[s]  frames per wave (0 = the whole call at once); a call's n frames are split into ceil(n / wave)
 *                   waves of nearly equal size
 *   flags           PHDFX_SCHED_REUSE: buffers that only hold a stage's intermediates are addressed wave-locally, i.e.
 *                   every wave rewrites the same wave_frames-sized region (lines are overwritten in L2 before they
 *                   are ever written back) instead of its own frame range of the arena.
 * Either way the tensors that enter or leave a wave stage must live in buffer ids that are not used for intermediates
 * (phdfx/weights.py: build_plan(stage_after_blocks=...)); a schedule that would overwrite live frames is refused.
 * With phdfx_extract_u8[_jitter], K1 runs per wave in front of stage 0, so its NHWC4p output is consumed from L2 too.
 * phdfx_load_weights resets the schedule to one stage without waves.  phdfx_get_schedule returns the stage count and
 * fills up to `cap` entries. */
#define PHDFX_SCHED_REUSE 1
int phdfx_set_schedule(phdfx_t* h, const int32_t* first_layer, const int32_t* wave_frames, int n_stages, int flags);
int phdfx_get_schedule(const phdfx_t* h, int32_t* first_layer, int32_t* wave_frames, int cap, int* flags);

/* Tile-granular dependencies between launches (experimental: a library built with PHDFX_EXPERIMENTAL=1, and
 * PHDFX_FLAGS=1 in the environment at phdfx_create).  A launch
 * starts (programmatic dependent launch) while its predecessor drains but touches no activation before that grid has
 * completed.  With the switch on, where it is safe — two consecutive plain conv launches on full grids, the second
 * reading only the first's output of at most PHDFX_FLAG_MAX_MB (default 64) MB — phdfx_forward instead lets the second
 * launch start each tile as soon as the first has published the frames that tile reads (per-frame progress counters,
 * csrc/conv_igemm_sm100.cuh).  Results are bit-identical.  Measured neutral on B200 (a CTA of the next
 * launch only becomes resident ~5 us after its predecessor's CTA exits, whatever it waits on afterwards), hence off by
 * default.  phdfx_linked_launches: how many launches of a pass over n frames start that way. */
int phdfx_linked_launches(const phdfx_t* h, int n);

int phdfx_layer_count(const phdfx_t* h);
int phdfx_layer_info(const phdfx_t* h, int layer_id, phdfx_layer_desc* out);
/* Number of kernels the last phdfx_forward / phdfx_extract_u8 / phdfx_preprocess_u8 call launched. */
int phdfx_last_launch_count(const phdfx_t* h);

#ifdef __cplusplus
}
#endif
#endif /* PHDFX_H_ */
