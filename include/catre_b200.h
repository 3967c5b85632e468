/*
 * catre_b200.h -- C ABI of libcatre_b200.so: the B200-native (sm_100a) engine for the CATRE
 * iterative point-cloud pose-refinement forward.
 *
 * The reference (THU-DA-6D-Pose-Group/CATRE) is pure Python/PyTorch and has no FFI; this is the C
 * surface a binding for its model-plugin boundary calls (ctypes: catre_b200/engine.py, shown in
 * INTEGRATION.md).  Each entry point cites the reference interface it stands in for.
 *
 * Conventions
 *   - every function returns 0 on success or a negative catre_status; nothing throws across the ABI;
 *     catre_last_error() returns a human-readable message for the last failure on that engine.
 *   - all tensors are fp32, contiguous, caller-owned.  "dev" pointers are device pointers on the
 *     engine's device; "host" pointers are host memory (pinned for async copies).
 *   - calls are stream-ordered on the cudaStream_t passed (void* here so the header needs no CUDA
 *     include), never synchronise the host (except the *_host entry, which returns results), and are
 *     CUDA-graph capturable.  One engine per device; calls on one engine are not re-entrant.
 *   - objects are independent: B may be any value >= 0 (B == 0 is a no-op).  B larger than
 *     cfg.max_batch is processed in chunks of max_batch.
 */
#ifndef CATRE_B200_H_
#define CATRE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct catre_engine catre_engine;

typedef enum catre_status {
  CATRE_OK = 0,
  CATRE_ERR_INVALID_ARG = -1,   /* null pointer, negative size, unknown enum */
  CATRE_ERR_SHAPE = -2,         /* weight / tensor shape does not match the configuration */
  CATRE_ERR_UNKNOWN_WEIGHT = -3,/* name is not one of the 74 checkpoint tensors */
  CATRE_ERR_NOT_PACKED = -4,    /* forward/refine before catre_pack, or weights missing at pack */
  CATRE_ERR_CUDA = -5,          /* a CUDA runtime call failed (message has the cudaError string) */
  CATRE_ERR_UNSUPPORTED = -6,   /* configuration outside the shipped CATRE config */
  CATRE_ERR_NO_DEVICE = -7      /* no CUDA device / not an sm_100 device */
} catre_status;

/* arithmetic of the wide per-point contractions (everything else is always fp32 FMA) */
typedef enum catre_precision {
  CATRE_PREC_FP32_SIMT = 0,   /* fp32 FMA on CUDA cores everywhere (strict mode, validation) */
  CATRE_PREC_F16X3 = 1,       /* tcgen05 kind::f16, operands split x = hi + lo (fp16 each, 22 significand bits),
                                 3 products hi*hi + hi*lo + lo*hi, fp32 TMEM accumulate: the fp32-parity mode
                                 (<= 1e-4 on R,t,s vs the reference fp32 forward; measured ~1e-5) */
  CATRE_PREC_BF16 = 2         /* tcgen05 single-product bf16 (BASELINE.json config 3; looser tolerance) */
} catre_precision;

/* Replaces: the cfg.MODEL.CATRE / cfg.INPUT values the reference reads at model build time
 * (core/catre/models/CATRE_disR_shared.py:291-350, configs/catre/NOCS_REAL/aug05_..._120e.py:29-30,73-114). */
typedef struct catre_cfg {
  int32_t n_obs;        /* INPUT.NUM_PCL: observed points per object (multiple of 128; of 256 in the tensor-core modes) */
  int32_t n_prior;      /* INPUT.NUM_KPS: prior points per object (same multiples; may differ from n_obs: the reference
                           only ties conv_p to the sum, core/catre/models/heads/conv_out_per_rot_head.py:112.  The training
                           step, catre_train_step, needs n_prior == n_obs) */
  int32_t max_batch;    /* workspace is sized for this many objects per launch (>= 1) */
  int32_t precision;    /* catre_precision */
  int32_t device;       /* CUDA device ordinal */
  int32_t reserved[3];  /* must be 0 */
} catre_cfg;

/* Replaces: build_model_optimizer(cfg, is_test=True) -> model  (CATRE_disR_shared.py:291-350). */
int catre_create(catre_engine** out, const catre_cfg* cfg);

/* Replaces: MyCheckpointer(model).resume_or_load(cfg.MODEL.WEIGHTS) (core/catre/main_catre.py:151,
 * core/utils/my_checkpoint.py:48-84).  `name` is the checkpoint key (e.g. "pcl_net.stn.conv1.weight");
 * `data` may be a host or a device pointer (copied with cudaMemcpyDefault); shape is checked against
 * the configuration.  All 74 tensors must be set before catre_pack. */
int catre_set_weight(catre_engine* e, const char* name, const float* data, const int64_t* shape, int32_t ndim);

/* Number of checkpoint tensors the engine expects, and the i-th expected name (for bindings/tests). */
int32_t catre_num_weights(void);
const char* catre_weight_name(int32_t i);

/* Derive the engine's packed weights (stacked heads, folded identities, 16-bit hi/lo splits, FC-chain slices). Must
 * be called after the weights change and before forward/refine.  Ordering: a set-up call, not a hot one -- it (and
 * catre_set_weight) waits for the whole device (cudaDeviceSynchronize) before touching the weights, so work enqueued
 * earlier on ANY stream (catre_train_set_weight copies, kernels still reading the old weights) is complete, and it
 * returns with the packed weights resident.  `stream` is accepted for symmetry with the hot calls. */
int catre_pack(catre_engine* e, void* stream);

/* Bytes of device workspace the engine holds for a launch of B objects (0 <= B <= max_batch). */
size_t catre_workspace_bytes(const catre_engine* e, int32_t B);

/* Replaces: CATRE_disR_shared.forward(x, tfd_kps, init_pose, init_scale, K_zoom, ..., do_loss=False)
 * (CATRE_disR_shared.py:40-124) = ONE refinement iteration.
 *   x_pm      [B, n_obs, 3]   zero-centred observed points, point-major (the memory behind the
 *                             reference's permuted view batch["x"], core/catre/engine/batch_test.py:95)
 *   kps_pm    [B, n_prior, 3] transformed prior points R(s*kps), point-major (batch_test.py:85-92)
 *   pose      [B, 3, 4], scale [B, 3], K [B, 3, 3] (only K[0][0], K[1][1] are read)
 *   out_pose  [B, 3, 4], out_scale [B, 3]          (= out_dict["pose_i"], out_dict["scale_i"]) */
int catre_forward_once(catre_engine* e, const float* x_pm, const float* kps_pm, const float* pose,
                       const float* scale, const float* K, int32_t B, float* out_pose, float* out_scale,
                       void* stream);

/* Replaces: the evaluator's K-loop, batch_updater_test + model.forward per iteration
 * (core/catre/engine/catre_evaluator.py:292-311, core/catre/engine/batch_test.py:63-97).
 *   pcl       [B, n_obs, 3]   observed cloud, camera frame (batch["pcl"])
 *   prior     [B, n_prior, 3] normalised category prior per object (batch["obj_kps"])
 *   out_poses [n_iter+1, B, 3, 4], out_scales [n_iter+1, B, 3]; entry 0 = the initial estimate, entry i =
 *   out_dict["pose_i"/"scale_i"] (the evaluator consumes every iteration).  Points never leave the GPU. */
int catre_refine(catre_engine* e, const float* pcl, const float* prior, const float* init_pose,
                 const float* init_scale, const float* K, int32_t B, int32_t n_iter, float* out_poses,
                 float* out_scales, void* stream);

/* Same as catre_refine with HOST buffers (pinned or pageable): copies inputs host->device, refines,
 * copies the poses device->host and synchronises the stream before returning.  This is the call the
 * end-to-end (`e2e`) benchmark times. */
int catre_refine_host(catre_engine* e, const float* pcl, const float* prior, const float* init_pose,
                      const float* init_scale, const float* K, int32_t B, int32_t n_iter, float* out_poses,
                      float* out_scales, void* stream);

/* Multi-GPU result collection: the reference gathers pickled Python lists of per-object dicts
 * (core/catre/engine/catre_custom_evaluator.py:200-203, detectron2 all_gather).  Here a rank's final poses are packed on
 * the device as [B, 15] fp32 (R|t row-major 12, then scale 3) -- the one buffer a single ncclAllGather on the launching
 * stream moves (catre_b200/shard.py).
 *   catre_pack_poses:          poses [n_iter+1, B, 3, 4], scales [n_iter+1, B, 3] (device) -> packed [B, 15] of iteration `iter`.
 *   catre_refine_host_packed:  catre_refine_host that additionally leaves the LAST iteration's packed poses in the caller's
 *                              device buffer packed_dev [B, 15], so the end-to-end path needs no re-upload before the gather. */
int catre_pack_poses(const float* poses, const float* scales, int32_t B, int32_t iter, float* packed, void* stream);
int catre_refine_host_packed(catre_engine* e, const float* pcl, const float* prior, const float* init_pose,
                             const float* init_scale, const float* K, int32_t B, int32_t n_iter, float* out_poses,
                             float* out_scales, float* packed_dev, void* stream);

/* catre_refine with the priors given as a category table instead of one copy per object: object b uses
 * prior_table[prior_cls[b]].  Replaces the same K-loop; the table is what get_normed_kps builds once per
 * dataset (core/catre/engine/engine_utils.py:17-24: batch["obj_kps"] = the category's mean shape from
 * cr_normed_mean_model_points_spd.pkl, selected by batch["obj_cls"]), so a caller need not expand it to
 * [B, n_prior, 3] (halves the input bytes of the path).
 *   prior_table [n_cls, n_prior, 3] fp32, prior_cls [B] int32 in [0, n_cls).
 * Device entry: an out-of-range class id cannot be reported without a host sync; that object's pose comes
 * back NaN.  Host entry: class ids are validated (CATRE_ERR_INVALID_ARG) and n_cls <= max(max_batch, 16). */
int catre_refine_table(catre_engine* e, const float* pcl, const float* prior_table, const int32_t* prior_cls,
                       int32_t n_cls, const float* init_pose, const float* init_scale, const float* K, int32_t B,
                       int32_t n_iter, float* out_poses, float* out_scales, void* stream);
int catre_refine_table_host(catre_engine* e, const float* pcl, const float* prior_table, const int32_t* prior_cls,
                            int32_t n_cls, const float* init_pose, const float* init_scale, const float* K,
                            int32_t B, int32_t n_iter, float* out_poses, float* out_scales, void* stream);

/* ---- Observed-cloud producer (the step before the path) ------------------------------
 * Replaces, for all objects of one image, the per-object CPU loop of the reference's test data loader
 * (core/catre/datasets/data_loader.py:773-799): misc.backproject_th (lib/pysixd/misc.py:360-378) +
 * crop_ball_from_depth_image / sample_bp_depth / crop_ball_from_pts (core/utils/cat_data_utils.py:209-226,
 * 283-318, 380-400) with SAMPLE_DEPTH_FROM_BALL=True, FPS_SAMPLE=False.  Engine-independent; errors are reported
 * through catre_last_error(NULL).  All pointers are device pointers except `intr` (host: fx, fy, cx, cy).
 *
 * catre_cloud_select: candidates of object b = pixels with masks[b] != 0 and depth > 0, in row-major order; they are
 *   cropped to the first of the n_radii (<= 16, non-decreasing) balls |p - centers[b]| <= radii[b][i] that holds
 *   >= 10 of them, to the last ball otherwise, and to all candidates when that ball is empty.  Writes the selected
 *   pixel ids in order to sel_pix[b][0 .. n_sel[b]) (sel_pix is [B, H*W]) and their count to n_sel[b].
 *   depth [H*W] fp32 metres, masks [B, H*W] uint8, centers [B,3] (the initial translation), radii [B, n_radii]
 *   (the caller applies the reference's max(ratio*|R s|, 0.05) * 1.1^i rule); scratch >= catre_cloud_scratch_bytes.
 * catre_cloud_gather: pcl[b][j] = back-projection of selected pixel (sample_idx[b][j] mod n_sel[b]); sample_idx
 *   [B, n_pts] int64 is the caller's torch.randperm draw over the repeated index list (the random draw stays on
 *   the host generator so a seed reproduces the reference's cloud bit for bit).  An object with n_sel == 0 gets NaN. */
size_t catre_cloud_scratch_bytes(int32_t B, int32_t H, int32_t W);
int catre_cloud_select(const float* depth, const uint8_t* masks, const float* intr, const float* centers, const float* radii,
                       int32_t n_radii, int32_t B, int32_t H, int32_t W, int32_t* sel_pix, int32_t* n_sel, void* scratch,
                       void* stream);
int catre_cloud_gather(const float* depth, const float* intr, const int32_t* sel_pix, const int32_t* n_sel,
                       const int64_t* sample_idx, int32_t B, int32_t H, int32_t W, int32_t n_pts, float* pcl, void* stream);

/* ---- Pairwise NOCS pose metrics (the step after the path) ------------------------------
 * Replaces the two nested Python loops of compute_combination_3d_matches (core/catre/engine/test_utils.py:331-352):
 * for pair t = (pair_pred[t], pair_gt[t]) computes, in fp64 and stored as fp32 like the reference's arrays,
 *   iou[t]            = compute_3d_iou_new(pred_RT, gt_RT, pred_scale, gt_scale, gt_handle, class names)  (:140-205)
 *   deg_shift[t][0/1] = compute_combination_RT_degree_cm_symmetry(pred_RT, gt_RT, cbrt(det(gt_RT[:3,:3])), gt class,
 *                       gt_handle)                                                                       (:208-277)
 * for any number of images per call (the pair lists are flat).  RTs are row-major 4x4 fp64 [n,16], scales [n,3] fp64,
 * class ids / handle visibility int32.  Class rules are passed as bit masks over class ids: sym_class_mask = classes
 * symmetric about y (bottle, bowl, can), flip_class_mask = classes symmetric under a 180-degree y flip (phone,
 * eggbox, glue; empty for NOCS), mug_class = id of "mug" (symmetric when its handle is not visible), -1 for none.
 * All pointers are device pointers; engine-independent; errors through catre_last_error(NULL). */
int catre_pair_metrics(const double* pred_RT, const double* pred_scale, const int32_t* pred_cls, const double* gt_RT,
                       const double* gt_scale, const int32_t* gt_cls, const int32_t* gt_handle, const int32_t* pair_pred,
                       const int32_t* pair_gt, int32_t n_pairs, uint32_t sym_class_mask, uint32_t flip_class_mask,
                       int32_t mug_class, float* iou, float* deg_shift, void* stream);

/* The same pair kernel with the shift definition as an argument and an optional fp64 output:
 *   shift_mode 0: |T1 - T2| / cbrt(det(gt_RT[:3,:3]))  (compute_combination_RT_degree_cm_symmetry, test_utils.py:275)
 *   shift_mode 1: |T1 - T2| * 100 [cm]                 (compute_RT_degree_cm_symmetry, test_utils.py:619-690) -- with
 *     deg_shift64 [n_pairs, 2] this fills the overlaps array of compute_RT_overlaps (test_utils.py:692-712, fp64) for the
 *     metric the NOCS evaluator actually calls: compute_independent_mAP (core/catre/engine/catre_custom_evaluator.py:254).
 * deg_shift (fp32) or deg_shift64 may be NULL, not both. */
int catre_pair_metrics_ex(const double* pred_RT, const double* pred_scale, const int32_t* pred_cls, const double* gt_RT,
                          const double* gt_scale, const int32_t* gt_cls, const int32_t* gt_handle, const int32_t* pair_pred,
                          const int32_t* pair_gt, int32_t n_pairs, uint32_t sym_class_mask, uint32_t flip_class_mask,
                          int32_t mug_class, int32_t shift_mode, float* iou, float* deg_shift, double* deg_shift64, void* stream);

/* Greedy prediction <-> ground-truth matching for many (image, class) sub-problems and all thresholds in one launch
 * (one thread per sub-problem and threshold combination).
 *   mode 0 replaces the matching loops of compute_3d_matches           (test_utils.py:586-614): thr_a = IoU thresholds;
 *   mode 1 replaces the loops of compute_match_from_degree_cm          (test_utils.py:734-755): thr_a = degree, thr_b = shift.
 * Sub-problem k owns predictions [sub_pred_off[k], sub_pred_off[k+1]) (in score order), ground truths
 * [sub_gt_off[k], ..) and the row-major [P_k, G_k] pair table starting at sub_pair_off[k] in `iou` (mode 0, fp32) or
 * `deg_shift64` (mode 1, [.., 2] fp64).  order[pair] = the candidate order of each row (indices within the sub-problem;
 * numpy's argsort of the row, supplied by the caller because the reference's tie order is numpy's), n_cand[pred] = how
 * many of them are candidates.  Outputs: gt_match [n_a * n_b, n_gt] and pred_match [n_a * n_b, n_pred], the matched
 * index within the sub-problem or -1.  All pointers are device pointers. */
int catre_match_greedy(int32_t mode, const int32_t* sub_pred_off, const int32_t* sub_gt_off, const int32_t* sub_pair_off,
                       int32_t n_sub, int32_t n_pred, int32_t n_gt, const float* iou, const double* deg_shift64,
                       const int32_t* order, const int32_t* n_cand, const int32_t* pred_cls, const int32_t* gt_cls,
                       const double* thr_a, int32_t n_a, const double* thr_b, int32_t n_b, int32_t* gt_match,
                       int32_t* pred_match, void* stream);

/* ---- Training step --------------------------------------------------------------------
 * Replaces one refinement iteration of the reference's training loop (core/catre/engine/engine.py:293-352 without the
 * optimiser): CATRE_disR_shared.forward(..., do_loss=True) (CATRE_disR_shared.py:40-165), catre_loss with the shipped
 * LOSS_CFG (:168-288; core/catre/losses/pm_loss.py:110-130, rot_loss.py:45-58; closest symmetric ground truth of
 * core/utils/pose_utils.py:472-528) and `sum(loss_dict.values()).backward()`.  All fp32 on CUDA cores; the backward is
 * hand-derived (sparse max-pool backward through the arg-max points, GroupNorm backward from two group sums, the
 * rotation head's layer-0 split) -- about 1/6 of the multiply-adds autograd spends on the same step.
 *
 * catre_train_set_weight: device-to-device refresh of the engine's fp32 copy of checkpoint tensor `name` (the
 *   optimiser owns the parameters; call after every optimiser step for the tensors it changed).  Needs one earlier
 *   catre_pack (which allocates the copies).  Marks the packed inference weights stale: the inference entries return
 *   CATRE_ERR_NOT_PACKED until the next catre_pack, which first pulls the refreshed tensors back from the device.
 * catre_train_set_weights: the same refresh for every tensor in ONE launch: src_dev is a HOST array of
 *   catre_num_weights() device pointers in catre_weight_name() order (contiguous fp32, the checkpoint's element counts);
 *   a null entry leaves that tensor as it is.  What the drop-in calls after an optimiser step (74 copies -> 1 launch).
 * catre_train_step: forward + losses + backward for B objects, stream-ordered, no host sync.  From the second step of a
 *   (B, number of symmetric objects, n_sym_rots, loss weights) combination on, the ~240 launches of the chain are one
 *   CUDA graph replayed on engine-owned static copies of the inputs (captured on an internal stream, launched on
 *   `stream`; same kernels, same order, bit-identical results; CATRE_TRAIN_GRAPH=0 launches kernel by kernel).  Independent
 *   parts of the chain (the x / y rotation heads and the ts head; a layer's parameter and data gradients in the encoder's
 *   backward; index builds that depend on the forward only) are issued on two engine-owned side streams ordered by events --
 *   parallel branches of the graph -- with per-lane scratch; accumulations into shared gradients keep the one-lane order, so
 *   the results do not depend on it (CATRE_TRAIN_LANES=0: one lane).  The call is still stream-ordered on `stream`.
 *   x_pm / tfd_pm [B, n, 3]  the re-posed points the reference's forward receives (as in catre_forward_once)
 *   obj_kps [B, n, 3]        normalised category prior (batch["obj_kps"], the point-matching loss's points)
 *   pose [B,3,4], scale [B,3], K [B,3,3]; gt_pose [B,3,4] (batch["obj_pose"]), gt_scale [B,3]   -- all device
 *   is_sym_host [B]          host bytes: 1 = sym_info[b] is not None (symmetric about y)
 *   sym_rots_host [n_sym_rots, 3, 3] host fp32: the symmetry rotations the data loader attaches to such objects
 *                            (lib/pysixd/misc.py:220-231; at most 1024)
 *   out_pose [B,3,4], out_scale [B,3] device = out_dict["pose_i"/"scale_i"]
 *   out_losses [6] device = (loss_PM_R, loss_rot, loss_yaxis_rot, loss_trans_xy, loss_trans_z, loss_scale); the
 *                            reference omits loss_rot / loss_yaxis_rot from its dict when no object is asymmetric /
 *                            symmetric -- here they are 0.
 *   The workspace (about 50 MB per object at 1024 + 1024 points plus 50 MB of fixed scratch, i.e. 25 KB per point) is allocated on the first call and grown when B grows.
 * catre_train_grad: copy d(sum of losses)/d(tensor `name`) of the last catre_train_step to dst (device or host,
 *   stream-ordered); tensors the shipped config never uses (the heads' `norm`) have zero gradients.
 * catre_train_set_loss_weights: LOSS_CFG.PM_LW, ROT_LW (rotation and y-axis terms), TRANS_LW (xy and z terms), SCALE_LW;
 *   each scales its loss values and gradients (CATRE_disR_shared.py:217, 238, 250, 262-263, 286).  Default 1 (shipped
 *   config); must be > 0 (the reference drops a term whose weight is 0 from its dict -- not implemented). */
int catre_train_set_weight(catre_engine* e, const char* name, const float* src_dev, void* stream);
int catre_train_set_weights(catre_engine* e, const float* const* src_dev, void* stream);
int catre_train_set_loss_weights(catre_engine* e, float pm_lw, float rot_lw, float trans_lw, float scale_lw);
int catre_train_step(catre_engine* e, const float* x_pm, const float* tfd_pm, const float* obj_kps, const float* pose,
                     const float* scale, const float* K, const float* gt_pose, const float* gt_scale,
                     const uint8_t* is_sym_host, const float* sym_rots_host, int32_t n_sym_rots, int32_t B, float* out_pose,
                     float* out_scale, float* out_losses, void* stream);
int catre_train_grad(catre_engine* e, const char* name, float* dst, void* stream);
/* All gradients at once: the engine keeps them in one arena (checkpoint order, every tensor on a 256-byte boundary).
 * catre_train_grad_layout: offsets[i] = position (in floats) of checkpoint tensor i inside the arena, *total_floats = arena
 *   size; depends on the point count only, valid before the first step.
 * catre_train_grads_flat: dst[0 .. total_floats) = scale * arena (device pointer, stream-ordered): one copy (scale == 1) or
 *   one kernel instead of 68 per-tensor copies; a binding hands out views of dst as the parameters' gradients. */
int catre_train_grad_layout(catre_engine* e, int64_t* offsets, int64_t* total_floats);
int catre_train_grads_flat(catre_engine* e, float* dst, float scale, void* stream);

/* ---- Fused optimiser step (core/catre/engine/engine.py:349-352) -------------------------------
 * Replaces: `optimizer.step()` of the optimiser the shipped config trains with, Ranger = RAdam + Lookahead + gradient
 * centralisation (lib/torch_utils/solver/ranger.py:102-200: ~10 small launches per tensor, ~700 for the model), and
 * optionally the gradient NaN guard in front of it (engine.py:349-352), in two launches over all tensors.
 * Engine-independent and stateless: the caller owns parameters, gradients and optimiser state (exp_avg, exp_avg_sq,
 * slow_buffer -- the reference's state-dict entries) and passes their device addresses in a table.
 *   table        [n_tensors, 8] int64, device: per tensor (param, grad, exp_avg, exp_avg_sq, slow_buffer, numel, row_len, 0);
 *                row_len = numel / shape[0] for tensors whose gradient is centralised (more than 1 dimension, or more than
 *                3 with gc_conv_only), else 0
 *   lr_wd        [n_tensors, 2] fp32, device: the tensor's group learning rate and weight decay
 *   elem_start   [n_tensors + 1] int64, device: prefix sums of numel;  row_start [n_tensors + 1]: prefix sums of the
 *                centralised tensors' shape[0] (0 rows for the others)
 *   rowmean      [total_rows] fp32 device scratch
 *   a            host: this step's scalars; step_size / rectified are the RAdam quantities of ranger.py:160-178 for the
 *                step count, lookahead = (step % k == 0)
 * All tensors fp32 and contiguous.  Errors through catre_last_error(NULL). */
typedef struct catre_ranger_args {
  float beta1, beta2, eps;
  float one_minus_beta1, one_minus_beta2; /* evaluated in double, rounded once (torch passes them as Python scalars) */
  float step_size;
  int32_t rectified;
  float alpha;
  int32_t lookahead;
  int32_t nan_to_num;
} catre_ranger_args;
int catre_ranger_step(const int64_t* table, const float* lr_wd, const int64_t* elem_start, const int64_t* row_start,
                      int32_t n_tensors, int64_t total_elems, int64_t total_rows, float* rowmean, const catre_ranger_args* a,
                      void* stream);

/* Number of kernels the last forward/refine/train call launched (bench.py's `gpu_launches`). */
int64_t catre_last_launch_count(const catre_engine* e);

/* Debug tap (tests only): synchronise and copy `bytes` of the internal workspace buffer `name`
 * ("t3", "t64", "gmax_g", "cset", ...) of the last launch to host memory. */
int catre_debug_read(catre_engine* e, const char* name, void* dst_host, size_t bytes);

/* Debug entry (tests / tools only): one GEMM of the training chain, C(m,n,z) = act(sum_k A(m,k,z) B(k,n,z) + bias(n)) [+ C],
 * with the chain's own strided / batched parameters, on device buffers.  `strides` = {sam, sak, sab, sbk, sbn, sbb, scm, scn, scb}
 * (elements), `kernel`: 0 = CUDA-core tile kernel, 1 = tcgen05 kernel with fp16 hi/lo operands (the forward GEMMs of
 * catre_train_step, reference modules: pointnets/pointnet.py:24-41,97-116 and the heads' Conv1d / Linear layers),
 * 2 = tcgen05 kernel with bf16 hi/lo operands (the backward GEMMs).  `splits` > 1 (batch must be 1) sums k in `splits` slices
 * through `partial` (>= splits * M * N floats) in a fixed order.  With `vmax` != NULL (kernel 1 or 2 only) the output is not
 * stored: it is max-pooled over each group of `rows_per_set` consecutive rows (a multiple of 128 dividing M) inside the GEMM's
 * epilogue -> vmax [M / rows_per_set, N] and the row index inside the group arg [.., N] (first index wins ties, like torch.max;
 * pointnets/pointnet.py:32,65,115); `partial` then needs 2 * (M / rows_per_set) * N floats. */
int catre_debug_train_gemm(const float* A, const float* B, float* C, const float* bias, const int64_t* strides, int32_t M, int32_t N,
                           int32_t K, int32_t batch, int32_t relu, int32_t accumulate, int32_t splits, float* partial,
                           int32_t kernel, int32_t rows_per_set, float* vmax, int32_t* arg, void* stream);

/* Average device time (ms, CUDA events on the launching stream) of the kernel group `which` over the
 * calls since catre_profile_reset; `which` indexes catre_profile_name().  Profiling is off by default
 * (events cost launches); enable with catre_profile_enable(e, 1).  Used by bench.py's roofline leg. */
int catre_profile_enable(catre_engine* e, int32_t on);
int catre_profile_reset(catre_engine* e);
int32_t catre_profile_num(void);
const char* catre_profile_name(int32_t which);
int catre_profile_get(catre_engine* e, int32_t which, double* total_ms, int64_t* launches);

/* Message for the last error on this engine (or for a failed catre_create when e == NULL). */
const char* catre_last_error(const catre_engine* e);

void catre_destroy(catre_engine* e);

/* Library build info: "catre_b200 <version> sm_100a ..." */
const char* catre_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CATRE_B200_H_ */
