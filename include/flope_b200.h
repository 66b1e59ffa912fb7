/* flope_b200 - C ABI of the B200-native FloPE pose path.
 *
 * The reference (wvu-irl/flope) is pure Python and has no FFI layer; its boundary for this
 * path is the Python predictor API.  These entry points are what a binding for that API
 * calls (flope_b200/_lib.py is the ctypes binding; INTEGRATION.md shows the reference-side
 * stub).  Each declaration cites the reference interface it replaces, as file:line under
 * the reference root.
 *
 * Conventions: return 0 on success or a negative FLOPE_E* code, never throw or exit;
 * flope_last_error() gives the thread-local message.  Every pointer whose name starts with
 * d_ is a device pointer on the engine's device; everything else is host memory.  All work
 * is stream-ordered on the stream argument (a cudaStream_t passed as void*); there is no
 * hidden device synchronisation: the caller synchronises.  The caller owns every input and
 * output buffer; the engine owns weights and workspace sized by max_batch.  One engine per
 * (device, stream): calls on different engines may run concurrently (from different threads
 * and streams), calls on one engine may not.  Concurrency on one device is safe by
 * construction: the default schedule runs layer1..layer4 as persistent launches whose CTAs
 * wait for each other's tiles and therefore need the whole grid resident, so the library
 * passes every forward that contains such a launch through a per-device gate (an event
 * chain under a host mutex) - two engines on two streams run their backbones one after the
 * other and overlap everything else.  Engines that are meant to overlap their backbones
 * select a schedule that waits for nothing, flope_engine_set_schedule(e, FLOPE_SCHED_PER_LAYER).
 * A dependency wait that still times out (the SMs held for seconds by another context's
 * kernels) does not trap: the launch finishes, its results are invalid, and the next call on
 * any engine of the process returns FLOPE_ECUDA once.
 */
#ifndef FLOPE_B200_H
#define FLOPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLOPE_OK 0
#define FLOPE_EINVAL (-1)   /* bad argument                       */
#define FLOPE_ECUDA (-2)    /* CUDA runtime error                 */
#define FLOPE_ESTATE (-3)   /* weights not loaded / wrong engine  */
#define FLOPE_ENOMEM (-4)   /* allocation failed                  */

#define FLOPE_INTERP_LINEAR 0     /* cv2.INTER_LINEAR, bit-exact uint8 arithmetic (benchmark mode)  */
#define FLOPE_INTERP_LANCZOS4 1   /* cv2.INTER_LANCZOS4, bit-exact (the reference's mode)            */
#define FLOPE_OUT_F32_NCHW 0      /* (B,3,S,S) float32, the tensor the reference builds              */
#define FLOPE_OUT_ENGINE 1        /* write straight into the engine's bf16 stem input                */

#define FLOPE_SCHED_PERSISTENT 0  /* default: one persistent launch for layer1..layer4, serialised per device        */
#define FLOPE_SCHED_PER_LAYER 1   /* one launch per layer: nothing waits, engines on different streams overlap       */
#define FLOPE_SCHED_DYNAMIC 2     /* stage chains that claim work from a counter: safe under partial residency       */
#define FLOPE_SCHED_COOPERATIVE 3 /* stage chains launched cooperatively (gang-scheduled)                            */

typedef struct flope_engine flope_engine;

/* One entry of PoseResNet.state_dict() (sunflower/models/posenet.py:5-34; key schema in
 * SURVEY.md appendix E), float32 host data.  num_batches_tracked entries may be omitted. */
typedef struct flope_tensor_desc {
  const char* name;
  const float* data;
  int ndim;
  int64_t shape[4];
} flope_tensor_desc;

int flope_version(void);
const char* flope_last_error(void);

/* Replaces `PoseResNet().to(device)` (sunflower/predictor/fast_pose_predictor.py:31,
 * pose_predictor.py:51).  crop_hw is the side of the square crops the engine is sized for
 * (512 = reference, 224 = benchmark configuration); it must be a multiple of 32. */
int flope_engine_create(flope_engine** out, int device, int max_batch, int crop_hw);
void flope_engine_destroy(flope_engine* e);

/* Replaces `posenet.load_state_dict(torch.load(path))` (fast_pose_predictor.py:32,
 * pose_predictor.py:52).  Folds eval-mode BatchNorm into fp32 scale/bias, packs conv/fc
 * weights to bf16 tiles, uploads.  Synchronous. */
int flope_engine_load_weights(flope_engine* e, const flope_tensor_desc* tensors, int n);

/* Replaces squarify_bb + bb_in_frame (sunflower/utils/mvg.py:324-351) as looped at
 * fast_pose_predictor.py:69-82.  Host function, integer, bit-exact.  boxes_xyxy: (n,4)
 * int32.  out_sq: (n,4) int32 squarified boxes (all n rows are written); keep[i] = 1 when the
 * squarified box lies inside the H x W frame. */
int flope_squarify_filter(const int32_t* boxes_xyxy, int n, int H, int W, int32_t* out_sq, uint8_t* keep);

/* Replaces the crop-batch loop (pose_predictor.py:138-153, fast_pose_predictor.py:108-123,
 * scripts/test_posenet.py:124-140): slice, cv2.resize of image and mask to (S,S), background
 * removal, /255, float32, NHWC->NCHW.
 * d_frames: (n_frames,H,W,3) uint8, frame_stride bytes apart; d_masks: (n_frames,H,W) uint8 or
 * NULL (== all 255); d_boxes: (n,5) int32 rows [frame, xmin, ymin, xmax, ymax], already squarified
 * and in-frame.  out_fmt FLOPE_OUT_F32_NCHW writes d_out (n,3,S,S) float32; FLOPE_OUT_ENGINE
 * ignores d_out and fills the engine's stem input (n <= max_batch, S == crop_hw). */
int flope_roi_crop(flope_engine* e, const uint8_t* d_frames, int n_frames, int H, int W, int64_t frame_stride,
                   const uint8_t* d_masks, const int32_t* d_boxes, int n, int S, int interp, void* d_out,
                   int out_fmt, void* stream);

/* Replaces `self.posenet(image_batch)` in eval mode (posenet.py:31-34; call sites
 * fast_pose_predictor.py:126, pose_predictor.py:156, scripts/test_posenet.py:142).
 * d_in: (n,3,S,S) float32 in [0,1] NCHW, or NULL to run on the stem input already filled by
 * flope_roi_crop(..., FLOPE_OUT_ENGINE) (then n <= max_batch).  d_r9: (n,9) float32. */
int flope_posenet_forward(flope_engine* e, const float* d_in, int n, float* d_r9, void* stream);

/* The first half of flope_posenet_forward on its own: convert (n,3,S,S) float32 NCHW crops (n <= max_batch) into the
 * engine's stem input, so that flope_posenet_forward(e, NULL, n, ...) can follow later.  Lets a caller with two engines
 * stage step i+1 on one of them (this call, or flope_roi_crop(..., FLOPE_OUT_ENGINE)) while the other runs step i's
 * backbone - the staging kernels are small enough to share the SMs with the persistent conv kernels
 * (flope_b200.pipeline.EnginePool(serial_backbones=True)). */
int flope_ingest_crops(flope_engine* e, const float* d_in, int n, void* stream);

/* Replaces procrustes_to_rotmat (sunflower/utils/conversion.py:54-58 -> roma.special_procrustes)
 * and, when d_R_yaw != NULL, nullify_yaw_batch (sunflower/utils/mvg.py:240-251).
 * d_r9: (n,9) f32.  d_R: (n,9) f32 row-major rotations (nullable).  d_R_yaw: (n,9) f64 (nullable). */
int flope_pose_head(flope_engine* e, const float* d_r9, int n, float* d_R, double* d_R_yaw, void* stream);

/* nullify_yaw_batch alone (mvg.py:240-251): d_R_in (n,9) f32 rotations -> d_R_yaw (n,9) f64. */
int flope_nullify_yaw(flope_engine* e, const float* d_R_in, int n, double* d_R_yaw, void* stream);

/* The fused path of get_flower_poses after detection (fast_pose_predictor.py:108-131):
 * ROI crop -> PoseNet -> Procrustes -> yaw nullification, any n (chunked by max_batch).
 * Outputs are nullable. */
int flope_infer_frames(flope_engine* e, const uint8_t* d_frames, int n_frames, int H, int W, int64_t frame_stride,
                       const uint8_t* d_masks, const int32_t* d_boxes, int n, int interp, float* d_r9, float* d_R,
                       double* d_R_yaw, void* stream);

/* ---- "next" row N2 of SURVEY.md section 8(f): the depth branch of get_flower_poses ----
 * Replaces get_depth_value (sunflower/utils/image_manipulation.py:39-96; shrink_mask :21-36), called at
 * sunflower/predictor/pose_predictor.py:118-121 (depth/10000, near 0.1, far 2.5) and
 * fast_pose_predictor.py:90-93 (depth/1000).  Stateless (no engine): all buffers are the caller's, on `device`.
 *   d_depth      (H,W) float32 metres (depth_dtype 0) or uint16 sensor units (depth_dtype 1; metres = raw / depth_div,
 *                the reference's depth.astype(np.float32) / 10000)
 *   d_mask       (H,W) uint8 segmentation mask (set where > 128)
 *   d_boxes      (n,4) int32 xmin,ymin,xmax,ymax - the DETECTOR boxes that survived squarify + in-frame filtering
 *   erode_k      side of cv2's MORPH_ELLIPSE structuring element (the reference uses 10)
 *   d_scratch    (H,W) uint8 work buffer (receives the eroded validity mask)
 *   d_val        (n) float64: mean depth in metres over the box's valid pixels (0 when there is none)
 *   d_count      (n) int32: number of valid pixels; the reference calls a box reliable when count >= 50
 * Validity, erosion and counts are exact; the sum is accumulated exactly (64-bit fixed point, order-independent)
 * where numpy sums float32 pairwise: agreement ~1e-7 relative, results deterministic. */
int flope_depth_values(int device, const void* d_depth, int depth_dtype, float depth_div, const uint8_t* d_mask, int H, int W,
                       const int32_t* d_boxes, int n, float near_plane, float far_plane, int erode_k, uint8_t* d_scratch,
                       double* d_val, int32_t* d_count, void* stream);

/* ---- "next" row N1 of SURVEY.md section 8(f): YOLO-seg post-processing ----
 * Replaces the mask half of FastPosePredictor.get_bbox_mask (sunflower/predictor/fast_pose_predictor.py:44-57):
 *   mask = uint8(clip(sum(masks, 0), 0, 1) * 255);  mask = cv2.resize(mask, (W, H))          # INTER_LINEAR, bit-exact
 * d_masks (n,h,w) float32 instance masks (n may be 0: all-zero mask), d_small (h,w) uint8 scratch, d_out (H,W) uint8,
 * d_tables scratch of (W + H) * 8 bytes.  Stateless; everything stays on `device`. */
int flope_yolo_mask(int device, const float* d_masks, int n, int h, int w, uint8_t* d_small, uint8_t* d_out, int H, int W,
                    void* d_tables, void* stream);

/* Number of kernels the last call on this engine launched (bench.py reports it). */
int flope_engine_last_launches(const flope_engine* e);

/* CUDA-event timing for bench.py's roofline pass.  flope_engine_profile(e,1) clears and enables
 * recording of one start/stop event pair around every kernel launch of subsequent calls (isolated
 * launches: the events between kernels prevent programmatic-dependent-launch overlap);
 * flope_engine_profile(e,2) records ONE pair around the trunk's conv_igemm launches (stem .. layer4, the
 * launches back to back exactly as in production, entry name "trunk"); 0 disables.
 * flope_engine_profile_read synchronises the device and returns the number of recorded launches,
 * their names as a '\n'-joined string and their durations in milliseconds. */
int flope_engine_profile(flope_engine* e, int enable);
int flope_engine_profile_read(flope_engine* e, char* names, int names_len, float* ms, int max_entries);

/* ---- test / bring-up hooks (used only by tests/) ---- */
/* Copy a named intermediate activation of the last forward as (n,C,H,W) float32 into d_out.
 * Names: "stem" (only written by the two-kernel stem path), "maxpool", "layer1.0" ... "layer4.1".
 * Returns C*H*W, or a negative error. */
int64_t flope_debug_activation(flope_engine* e, const char* name, int n, float* d_out, void* stream);
/* Evaluate the device mask/normalise arithmetic for all (mask,img) uint8 pairs: d_out (256,256) f32. */
int flope_debug_normalise_lut(float* d_out, void* stream);
/* Set a named option (A/B switches for tests and tools; the defaults are the product configuration):
 *   "use_graph"   0/1  replay the backbone as a CUDA graph (default 1)
 *   "pdl"         0/1  programmatic dependent launch between the backbone kernels (default 1)
 *   "fuse_pool"   0/1  stem conv + max-pool as one kernel instead of two (default 1 when the crop side is <= 252)
 *   "roi_strip"   even 2..128  output rows per CTA of the bilinear ROI kernel (default 14)
 *   "chain"       0/1  one persistent launch per ResNet stage (four convs, per-tile dependencies) instead of one
 *                      launch per layer (default 1; see the concurrency note at the top)
 *   "trunk"       0/1  layer1..layer4 as ONE launch (trunk_chain_kernel) when the plan has its tile shapes (default 1;
 *                      needs "chain" 1, "pair" 1, no "chain_coop"/"chain_dynamic"; same concurrency note)
 *   "chain_coop"  0/1  launch the chains cooperatively (default 0; required when engines share a device)
 *   "chain_dynamic" 0/1  chains claim their work items in index order from an atomic counter instead of
 *                      round-robin by block index: safe under partial residency without a cooperative launch
 *                      (default 0: 3 % slower than the static deal on a single stream; 2 = the static deal routed
 *                      through the same item queue, for diagnosis)
 *   "timeline"    0/1  record per-CTA phase stamps in the conv kernels (default 0; read with flope_debug_timeline)
 *   "pair"        0/1  CTA-pair (tcgen05 cta_group::2) conv kernels instead of single-CTA ones (default 1)
 *   "small_tiles" 0/1  latency-oriented tiles when max_batch cannot fill the SMs (default 1)
 *   "fc_small"    0/1  fc GEMM on 128-crop x 64-channel single-CTA tiles instead of 256 x 128 (default 1)
 * "pair", "small_tiles" and "fc_small" change the packed-weight layout: call flope_engine_load_weights again afterwards.
 * Activation names for flope_debug_activation additionally include "x0" (the stem's space-to-depth input). */
/* Host helper of the predictors (no reference counterpart: the reference slices the numpy frame per box,
 * pose_predictor.py:141-143): copy the n box regions of a host (H,W,ch) uint8 image into n fixed-size slots of
 * slot_h x slot_w x ch bytes each, box i at the origin of slot i.  With a handful of flowers per frame the boxes are a
 * fraction of the frame, so the predictor uploads the packed slots (as n small "frames", boxes (i,0,0,s,s)) instead of the
 * whole frame.  Bytes of a slot outside its box are left as they are.  boxes: n x 4 xyxy, in-frame, sides <= slot_h, slot_w. */
int flope_pack_boxes(const uint8_t* img, int H, int W, int ch, const int32_t* boxes_xyxy, int n, int slot_h, int slot_w,
                     uint8_t* out);
int flope_debug_set(flope_engine* e, const char* key, int value);
/* How the backbone of this engine is launched (FLOPE_SCHED_*); see the concurrency paragraph at the top.  Product
 * option (flope_b200.pipeline.EnginePool uses it), cheap: it only drops the engine's captured CUDA graphs. */
int flope_engine_set_schedule(flope_engine* e, int schedule);
/* After flope_debug_set(e, "timeline", 1): synchronise and copy the phase stamps of the conv_igemm launches of the
 * last forward, [launch][148 CTAs][8] uint64 (%globaltimer ns at kernel entry, prologue done, dependency wait done, first
 * operands landed, last MMA issued, first accumulator ready, last store issued, exit;
 * zeros for CTAs a launch did not have).  Returns the number of launches copied (<= max_launches, <= 32). */
int flope_debug_timeline(flope_engine* e, unsigned long long* out, int max_launches);

#ifdef __cplusplus
}
#endif
#endif /* FLOPE_B200_H */
