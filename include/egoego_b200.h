/*
 * egoego_b200.h -- C ABI of libegoego_b200.so: the B200-native (sm_100a) implementation of EgoEgo's
 * stage-2 conditional motion-diffusion sampling path.
 *
 * The reference (lijiaman/egoego_release) is pure Python/PyTorch and has no FFI layer; its call
 * boundary for this path is a set of methods on an nn.Module.  Each entry point below replaces one
 * of them (paths relative to the reference repo):
 *
 *   egoego_create / egoego_set_tensor / egoego_commit_weights
 *        CondGaussianDiffusion.__init__ + load_state_dict
 *        (egoego/model/transformer_cond_diffusion_model.py:144-214,
 *         trainer_amass_cond_motion_diffusion.py:108-122)
 *   egoego_denoiser_forward   TransformerDiffusionModel.forward          (..._diffusion_model.py:118-141)
 *   egoego_p_sample_step      CondGaussianDiffusion.p_sample (+ in-paint) (..._diffusion_model.py:248-256,392-397)
 *   egoego_sample             CondGaussianDiffusion.sample / p_sample_loop (..._diffusion_model.py:258-270,527-535)
 *   egoego_sample_host        same, host buffers in/out (H2D + loop + D2H inside the call)
 *   egoego_postprocess        convert_model_res_to_data + AMASSDataset.fk_smpl
 *                             (..._diffusion_model.py:469-525; egoego/data/amass_diffusion_dataset.py:265-293)
 *   egoego_canonicalize_head  rotate_at_frame_smplh + x_start build       (egoego/lafan1/utils.py:111-137;
 *                                                                           ..._diffusion_model.py:358-386)
 *   egoego_tail_condition     next-window conditioning from the tail      (..._diffusion_model.py:423-464)
 *
 * Conventions: every pointer is BORROWED for the duration of the call; the caller owns all buffers,
 * the library owns only its workspace and packed weights.  "dev" pointers are CUDA device pointers on
 * the handle's device; all tensors are dense row-major fp32 unless stated.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous on that stream
 * except the *_host entry points, which synchronise the stream before returning.  Every function
 * returns 0 on success, non-zero on error (message via egoego_last_error()).  There is no CPU
 * fallback: without a CUDA device of compute capability 10.x every call fails.
 * A handle is re-entrant across handles, not thread-safe per handle.
 */
#ifndef EGOEGO_B200_H
#define EGOEGO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct egoego_ctx* egoego_handle;

/* Mirrors the constructor arguments of CondGaussianDiffusion (..._diffusion_model.py:144-161). */
typedef struct egoego_cfg {
    int32_t d_feats;       /* 198 = 22*3 + 22*6                                        */
    int32_t d_model;       /* 512                                                      */
    int32_t n_head;        /* 4                                                        */
    int32_t n_dec_layers;  /* 4                                                        */
    int32_t d_k;           /* 256                                                      */
    int32_t d_v;           /* 256                                                      */
    int32_t max_timesteps; /* 121 = window + 1 (time token)                            */
    int32_t timesteps;     /* diffusion steps N (1000)                                 */
    int32_t objective;     /* 0 = pred_noise, 1 = pred_x0                              */
    int32_t max_batch;     /* workspace is sized for this many windows per call        */
    int32_t device;        /* CUDA ordinal                                             */
    int32_t engine;        /* EGOEGO_ENGINE_*                                          */
    int32_t precise_last_steps; /* tensor engine precision policy: the last K diffusion steps (t < K) read every weight as an
                              exact fp16 hi/lo pair -- the last 16 of them (env EGOEGO_SPLIT_STEPS) in the 3-term split with
                              hi/lo activations too (fp32-grade products), the others with fp16 activations (two passes over
                              K); earlier steps one fp16 pass over dithered weight copies (egoego_weight_sets), whose error
                              is damped by posterior_mean_coef1[t].
                              0 (a zero-initialised cfg) or -1 = default max(ceil(timesteps/16), 48); K >= timesteps = the
                              3-term split at EVERY step; EGOEGO_PRECISE_ALL_FP16 = every step single-pass (30 mm
                              worst-window error: measurements only).  The per-call entry points (denoiser_forward,
                              p_sample_step) always use the 3-term split. */
} egoego_cfg;

/* explicit opt-out of the precision policy (egoego_cfg.precise_last_steps): no split steps at all */
#define EGOEGO_PRECISE_ALL_FP16 (-2)

enum {
    EGOEGO_ENGINE_TCGEN05 = 0, /* tcgen05/TMA GEMMs, 3-term fp16 hi/lo split, fp32 accumulate (default) */
    EGOEGO_ENGINE_SIMT    = 1  /* fp32 CUDA-core GEMMs: validation / bisecting engine               */
};

/* Gaussian noise source of the sampler.
 * tape != NULL : draws are read from `tape`, draw k at tape + k*B*T*d_feats (device memory for the
 *                *device* entry points, host memory for egoego_sample_host).  Draw order is the
 *                reference's: 0 = x init, 1 = x_cond noise, 2+k = k-th executed step.
 * tape == NULL : counter-based Philox4x32-10 + Box-Muller keyed by (seed, window_offset + window,
 *                draw, element) -- independent of batch split and GPU count. */
typedef struct egoego_rng {
    const float* tape;
    uint64_t     seed;
    uint64_t     window_offset;
} egoego_rng;

const char* egoego_last_error(void);
int  egoego_version(void);

int  egoego_create(const egoego_cfg* cfg, egoego_handle* out);
int  egoego_destroy(egoego_handle h);

/* Weight upload by the reference's state_dict key, e.g.
 * "denoise_fn.motion_transformer.layer_stack.0.self_attn.w_q.weight" (an optional leading
 * "ema_model." / "model." prefix is ignored).  Also accepts the schedule buffers
 * "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped",
 * "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod" ([timesteps]).  Unknown keys return 0 and
 * are ignored (load_state_dict(strict=False) semantics); a wrong element count is an error. */
int  egoego_set_tensor(egoego_handle h, const char* name, const float* data, int64_t numel, int on_device);
/* Computes the schedule buffers from cfg.timesteps (cosine schedule, fp64 then cast), for callers
 * without the reference's registered buffers. */
int  egoego_make_cosine_schedule(egoego_handle h);
/* Packs weights for the engine (fused QKV, fp16 hi/lo planes + dithered fp16 copies, timestep-embedding table, TMA maps).
 * Must be called after the last set_tensor and before any compute call. */
int  egoego_commit_weights(egoego_handle h, void* stream);

/* out[B,T,d_feats] = denoise_fn(x_all[B,T,2*d_feats], t[B])   (t: int64 device array, 0 <= t < timesteps; values outside
 * that range are clamped on the device -- the timestep-embedding and schedule tables have `timesteps` rows)
 * padding_mask: NULL or float [B,T+1] (1 = keep, 0 = zero the row after every sub-layer). */
int  egoego_denoiser_forward(egoego_handle h, const float* x_all_dev, const int64_t* t_dev,
                             const float* padding_mask_dev, int B, int T, float* out_dev, void* stream);

/* x_out = p_sample(x, t, x_cond): denoiser -> (pred_x0|pred_noise) -> clamp(-1,1) if clip_denoised ->
 * posterior mean -> + 1[t>0] * exp(0.5*logvar[t]) * noise.  noise_dev NULL => Philox draw `draw_index`.
 * If inpaint_dev != NULL, frames [0, inpaint_len) of x_out are then overwritten with
 * inpaint_dev[B, inpaint_len, d_feats] (the sliding-window overlap conditioning).
 * x_out may alias x. */
int  egoego_p_sample_step(egoego_handle h, const float* x_dev, const int64_t* t_dev, const float* x_cond_dev,
                          const float* noise_dev, const egoego_rng* rng, uint64_t draw_index,
                          const float* padding_mask_dev, int clip_denoised,
                          const float* inpaint_dev, int inpaint_len,
                          int B, int T, float* x_out_dev, void* stream);

/* Whole sampling loop on the device: x ~ N(0,I); x_cond = x_start*(1-mask) + mask*N(0,I);
 * for t = N-1..0: x = p_sample(x, t, x_cond) [+ in-paint].  out_dev[B,T,d_feats].
 * inpaint_dev/inpaint_len as above (NULL/0 for plain sample()).  x_init_dev: NULL (draw 0 of the rng)
 * or an explicit initial x (the sliding-window path slices one long noise tensor). */
int  egoego_sample(egoego_handle h, const float* x_start_dev, const float* cond_mask_dev, int B, int T,
                   const egoego_rng* rng, const float* x_init_dev,
                   const float* inpaint_dev, int inpaint_len, float* out_dev, void* stream);

/* Same with HOST buffers (pinned or pageable): copies x_start/cond_mask (and the tape, if any) to the
 * device, runs the loop, copies the result back and synchronises.  Batches larger than
 * cfg.max_batch are processed in chunks. */
int  egoego_sample_host(egoego_handle h, const float* x_start_host, const float* cond_mask_host, int B, int T,
                        const egoego_rng* rng, float* out_host, void* stream);

/* Skeleton + normalisation statistics used by the post-processing kernels:
 * parents[22] (parents[0] = -1), rest_offsets[22*3], jpos_min[66], jpos_max[66] (host pointers). */
int  egoego_set_skeleton(egoego_handle h, const int32_t* parents, const float* rest_offsets,
                         const float* jpos_min, const float* jpos_max);

/* convert_model_res_to_data (+ optional FK): x[B,T,198] normalised sample, recover_quat[B,4] (wxyz,
 * NULL = identity).  Outputs (each nullable): local axis-angle aa[B,T,22,3], root[B,T,3], head[B,T,3],
 * FK global joint positions jpos[B,T,22,3] and global rotations gquat[B,T,22,4] (wxyz, w>=0). */
int  egoego_postprocess(egoego_handle h, const float* x_dev, const float* recover_quat_dev, int B, int T,
                        float* aa_dev, float* root_dev, float* head_dev, float* jpos_dev, float* gquat_dev,
                        void* stream);

/* AMASSDataset.fk_smpl: (root[N,3], aa[N,22,3]) -> gquat[N,22,4], jpos[N,22,3] (either nullable). */
int  egoego_fk_smpl(egoego_handle h, const float* root_dev, const float* aa_dev, int64_t N,
                    float* gquat_dev, float* jpos_dev, void* stream);

/* Head-frame canonicalisation of one window (rotate_at_frame_smplh with cano_t_idx=0, then move the
 * first frame's x,y to 0) and construction of the window's conditioning:
 *   head_pos[B,T,3], head_quat[B,T,4] (wxyz)  ->  x_start[B,T,198] (zeros except normalised head
 *   position at 45:48 and head rot6d at 156:162), recover_quat[B,4]. */
int  egoego_canonicalize_head(egoego_handle h, const float* head_pos_dev, const float* head_quat_dev,
                              int64_t pos_stride_frames, int B, int T,
                              float* x_start_dev, float* recover_quat_dev, void* stream);

/* Conditioning of the next sliding window (reference :423-464): FK result of the last n frames of a window,
 * gquat[B,n,22,4] (wxyz) / gjpos[B,n,22,3], re-canonicalised at the tail's first frame, normalised and converted to
 * rot6d -> inpaint_out[B,n,198] (the tensor egoego_sample's `inpaint_dev` expects). */
int  egoego_tail_condition(egoego_handle h, const float* gquat_dev, const float* gjpos_dev, int B, int n,
                           float* inpaint_out_dev, void* stream);

/* Post-sampling evaluation of B sequences of T frames (T >= 3) -- compute_metrics_for_smpl,
 * kinpoly/scripts/eval_metrics_imu_rec.py:264-342 (called from eval_stage2.py:192), with what it calls
 * (compute_accel :66-77, compute_error_accel :79-107, compute_foot_sliding_for_smpl :222-262, get_root_matrix /
 * get_frobenious_norm(_rot_only) kinpoly/relive/utils/metrics.py:15-24,64-82).  Inputs are device pointers:
 * global joint rotations quat[B,T,22,4] (wxyz, need not be unit) and positions jpos[B,T,22,3] of ground truth and
 * prediction, floor heights floor[B] each.  out_dev[B,35] = root_trans_dist, accel_pred, accel_gt, accel_err, pred_fs,
 * gt_fs, head_trans_dist, root_dist, root_rot_dist, mpjpe, mpjpe_wo_hand, head_dist, head_rot_dist, single_jpe[22]
 * (mm for distances / accelerations / sliding, Frobenius norms unitless).  No handle: needs only the device. */
#define EGOEGO_METRICS_OUT 35
int  egoego_eval_metrics(int device, const float* gt_quat_dev, const float* gt_jpos_dev, const float* gt_floor_dev,
                         const float* pred_quat_dev, const float* pred_jpos_dev, const float* pred_floor_dev,
                         int B, int T, float* out_dev, void* stream);

/* Floor height and contact labels of B sequences -- determine_floor_height_and_contacts,
 * utils/data_utils/process_amass_dataset.py:160-317 (+ detect_joint_contact :319-328), the host-side numpy + sklearn.DBSCAN step
 * eval_stage2.py:131,189 / eval_egoego.py:331,395 run before every metrics call.  jpos_dev[B,T,22,3] global joint positions
 * (z up), 2 <= T <= 2048.  floor_out_dev[B] = the function's first return value (smallest cluster median of the static toe
 * heights minus 0.01; 0 when no toe sample is static); contacts_out_dev[B,T,22] (nullable) its second (1.0 / 0.0 on feet, toes,
 * hands, knees); discard_out_dev[B] int32 (nullable) its third (terrain heuristic).  No handle: needs only the device. */
int  egoego_floor_contacts(int device, const float* jpos_dev, int B, int T, int fps, float* floor_out_dev, float* contacts_out_dev,
                           int32_t* discard_out_dev, void* stream);

/* ---- Stage-1 networks (shipped configuration --input_of_feats) ---------------------------------
 * A "sequence net" is the reference's Decoder used WITHOUT a leading token (egoego/model/transformer_module.py:172-226,
 * use_full_attention=True, row padding mask) followed by up to two MLP heads (egoego/model/mlp.py:4-27: ReLU after every
 * affine layer, then one Linear).  HeadFormer (egoego/model/head_estimation_transformer.py:50-100: d_feats 512, window 60,
 * heads va[1024,512,256]->3 and dist[1024,512,256]->1) and HeadNormalFormer (head_normal_estimation_transformer.py:64-106:
 * d_feats 18, window 120, head normal[512,256]->3 on token 0) are two instances.  fp32 CUDA-core kernels (batch 1, <= 120
 * tokens: latency-bound).  Tensor names (host fp32, reference layouts): "start_conv.weight" [d_model,d_feats],
 * "start_conv.bias", "position_vec.weight" [(window+1),d_model], "layer_stack.<l>.self_attn.{w_q,w_k,w_v,fc,layer_norm}.{weight,bias}",
 * "layer_stack.<l>.pos_ffn.{w_1,w_2,layer_norm}.{weight,bias}", "head<h>.affine_layers.<j>.{weight,bias}", "head<h>.fc.{weight,bias}". */
typedef struct egoego_seqnet_ctx* egoego_seqnet;
typedef struct egoego_seqnet_cfg {
    int32_t d_feats, d_model /* 256 */, n_head, n_layers, d_k /* 256 */, d_v /* 256 */, window /* <= 128 */, max_batch, device;
    int32_t n_heads;              /* 0..2 */
    int32_t head_n_hidden[2];     /* 1..3 hidden layers per head */
    int32_t head_hidden[2][3];    /* hidden sizes (multiples of 16) */
    int32_t head_out[2];          /* output size of the head's final Linear */
} egoego_seqnet_cfg;
int  egoego_seqnet_create(const egoego_seqnet_cfg* cfg, egoego_seqnet* out);
void egoego_seqnet_destroy(egoego_seqnet h);
int  egoego_seqnet_set_tensor(egoego_seqnet h, const char* name, const float* host_data, int64_t numel);
int  egoego_seqnet_commit(egoego_seqnet h);
/* feats_dev[B,T,d_feats] (T <= window; the window is zero-padded and every position attended to, exactly like the reference)
 * -> decoder output dec_out_dev[B,window,d_model] (nullable) and the heads' outputs head<h>_out_dev[B,T,out] (nullable), or
 * [B,out] of token 0 only when token0_only != 0 (HeadNormalFormer.forward :158). */
int  egoego_seqnet_forward(egoego_seqnet h, const float* feats_dev, int B, int T, float* dec_out_dev, float* head0_out_dev,
                           float* head1_out_dev, int token0_only, void* stream);
int64_t egoego_seqnet_launch_count(egoego_seqnet h);
/* HeadFormer.va2rot (:102-124): q0[B,4] (wxyz), angular velocities va[B,T,3] -> out[B,T+1,4]. */
int  egoego_va2rot(int device, const float* q0_dev, const float* va_dev, int B, int T, float dt, float* out_dev, void* stream);
/* HeadFormer.cal_scale_for_slam_w_pred_scale (:184-212) for one sequence: slam_trans[n_pose,3], predicted per-step
 * distances dist[n_dist] (divided by dist_scale) -> re-scaled trajectory trans_out[n_pose,3], scale_out[1]. */
int  egoego_rescale_slam(int device, const float* slam_trans_dev, int n_pose, const float* dist_dev, int n_dist, float dist_scale,
                         float* trans_out_dev, float* scale_out_dev, void* stream);
/* HeadNormalFormer input features (:128-137): rot_mat[B,n_pose_stride,3,3], trans[B,n_pose_stride,3], first n_pose poses
 * -> feats[B,n_pose-1,18]. */
int  egoego_slam_features(int device, const float* rot_mat_dev, const float* trans_dev, int B, int n_pose_stride, int n_pose,
                          float* feats_dev, void* stream);
/* HeadNormalFormer.forward_for_eval :219-250 (before the xy-plane alignment): rotation taking normal[B,3] to +z
 * (:47-62), applied with scale[B] to the SLAM translations (re-integrated from pose 0) and rotations.  Outputs nullable:
 * trans_out[B,n_pose,3], rot_out[B,n_pose,3,3], quat_out[B,n_pose,4] (wxyz), align_rot_out[B,3,3]. */
int  egoego_apply_floor_normal(int device, const float* normal_dev, const float* scale_dev, const float* rot_mat_dev,
                               const float* trans_dev, int B, int n_pose, float* trans_out_dev, float* rot_out_dev,
                               float* quat_out_dev, float* align_rot_out_dev, void* stream);

/* De-heading step of HeadNormalFormer.forward_for_eval (:268-276) with a given rotation rot3x3[B,3,3] (the xy-plane
 * alignment): rot_out = R rot_mat, trans_out = R (trans - trans[:,0]) + offset[B,3] (offset nullable); quat_out wxyz. */
int  egoego_rigid_apply(int device, const float* rot3x3_dev, const float* offset_dev, const float* rot_mat_dev, const float* trans_dev,
                        int B, int n_pose, float* trans_out_dev, float* rot_out_dev, float* quat_out_dev, void* stream);

/* ResNet-18 optical-flow encoder of HeadNet (HeadFormer with input_of_feats=False): egoego/model/resnet.py:5-23 (torchvision
 * resnet18 with fc -> out_dim) as called from head_estimation_transformer.py:216-224 on [T,224,224,2] flow + one zero channel.
 * eval() semantics (BatchNorm with running statistics, folded into the convolutions at commit).  Tensor names = the
 * state_dict keys of the wrapped torchvision model ("conv1.weight", "bn1.running_mean", "layer2.0.downsample.0.weight", ...,
 * "fc.weight"), host fp32.  forward: flow_dev[N,224,224,2] -> feats_dev[N,out_dim]. */
typedef struct egoego_resnet_ctx* egoego_resnet;
int  egoego_resnet18_create(int device, int out_dim, egoego_resnet* out);
void egoego_resnet18_destroy(egoego_resnet h);
int  egoego_resnet18_set_tensor(egoego_resnet h, const char* name, const float* host_data, int64_t numel);
int  egoego_resnet18_commit(egoego_resnet h);
int  egoego_resnet18_forward(egoego_resnet h, const float* flow_dev, int N, float* feats_dev, void* stream);
int64_t egoego_resnet18_launch_count(egoego_resnet h);

/* Training step of the denoiser: CondGaussianDiffusion.forward / p_losses
 * (egoego/model/transformer_cond_diffusion_model.py:557-625) -- q_sample, conditioning, denoiser forward with saved activations,
 * L1 (loss_l2 = 0) or L2 loss with the padding mask and per-sample weight, and the backward pass through every layer.
 * Handles created with EGOEGO_ENGINE_SIMT only (fp32); weights = the last egoego_commit_weights.  Dropout follows
 * egoego_train_set_dropout (default p = 0: the reference in eval() mode).  All pointers are device pointers:
 * x_start / cond_mask / noise / cond_noise [B,T,d_feats], padding_mask float [B,T+1] or NULL, t int64 [B],
 * sqrt_ac = sqrt_alphas_cumprod[t], sqrt_1mac = sqrt_one_minus_alphas_cumprod[t], weight = p2_loss_weight[t] (each float [B]).
 * loss_out_dev[1] receives the scalar loss; gradients stay in the handle until the next call and are read with
 * egoego_train_get_grad(name = reference state_dict key, dst_dev[numel] in the reference's layout).
 * Execution: the inputs are copied into the handle's staging buffers on `stream`; the ~340 launches of the step run on the handle's
 * own stream (ordered after / before `stream` by events) -- eagerly for the first step of a (B, T, loss, mask, dropout) shape, as ONE
 * captured CUDA graph from the second step on (a shape whose capture fails stays eager).  The input tensors may therefore be freed or
 * overwritten as soon as the call returns.  Environment: EGOEGO_TRAIN_GRAPH=0 (always eager), EGOEGO_TRAIN_SPLITK=0 (weight gradients
 * as single-pass products instead of split-K partial products added with atomics; the time-MLP gradients and the loss use atomics
 * either way: results vary by ~1e-7 relative run to run), EGOEGO_TRAIN_FUSE_EPI=0 (element-wise
 * epilogues as separate passes), EGOEGO_TRAIN_GEMM=simt (fp32 CUDA-core products). */
int  egoego_train_step(egoego_handle h, const float* x_start_dev, const float* cond_mask_dev, const float* padding_mask_dev,
                       const int64_t* t_dev, const float* noise_dev, const float* cond_noise_dev, const float* sqrt_ac_dev,
                       const float* sqrt_1mac_dev, const float* weight_dev, int loss_l2, int B, int T, float* loss_out_dev, void* stream);
int  egoego_train_get_grad(egoego_handle h, const char* name, float* dst_dev, int64_t numel, void* stream);
/* Dropout of the following egoego_train_step calls on this handle -- nn.Dropout(0.1) on the attention probabilities, the fc output
 * and the FFN output of every DecoderLayer while the reference module is in train() mode (egoego/model/transformer_module.py:53,59,
 * 84,92,105,113).  p = 0 (default) is the eval()-mode identity.  Masks are counter-based: element i of stream 4*layer + site keeps its
 * value (scaled by 1/(1-p)) iff word i%4 of Philox4x32-10(counter = (i/4, i>>34, stream, 0x44524f50), key = seed) < floor((1-p) 2^32);
 * site 0 index ((b H + h) 128 + query) 128 + key, sites 1 / 2 index (b 128 + token) 512 + channel.  The same masks are re-derived
 * in the backward pass, so no mask tensor is stored. */
int  egoego_train_set_dropout(egoego_handle h, double p, uint64_t seed);
/* Device-to-device refresh of one parameter tensor (reference state_dict key, reference layout) of a committed
 * EGOEGO_ENGINE_SIMT handle: what an optimizer step needs between two training steps (no host staging, no re-allocation;
 * the timestep-embedding table is rebuilt lazily when a time_mlp tensor changed). */
int  egoego_update_tensor_device(egoego_handle h, const char* name, const float* src_dev, int64_t numel, void* stream);
/* Batched forms of the two calls above (one FFI call per training step instead of one per tensor). */
int  egoego_train_get_grads(egoego_handle h, int n, const char* const* names, float* const* dst_dev, const int64_t* numels, void* stream);
int  egoego_update_tensors_device(egoego_handle h, int n, const char* const* names, const float* const* src_dev, const int64_t* numels, void* stream);

/* Introspection for tests/bench: number of kernels launched by this handle since creation, and the
 * cumulative count of denoiser steps executed. */
int64_t egoego_launch_count(egoego_handle h);

/* Content checksum of n device tensors of 32-bit elements (fp32 parameters): wrapping 64-bit sum of word * (2 * index + 1)
 * over the concatenation, one kernel launch + one 8-byte read-back (synchronises `stream`).  The host mirror uses it to
 * notice in-place parameter edits that bypass torch's version counters (p.data.copy_ / lerp_, as ema_pytorch does:
 * trainer_amass_cond_motion_diffusion.py:179-192) before it reuses the engine's packed weights.  No handle needed. */
int  egoego_tensors_checksum(int device, int n, const void* const* ptrs_dev, const int64_t* numels, uint64_t* out_host, void* stream);

/* Measurement hook (bench.py roofline / per-kernel table): average device time of one launch of ONE kernel of the
 * sampling step -- `which` = EGOEGO_KERNEL_* (layer 0's weights) -- over `iters` back-to-back launches on `stream`,
 * timed with CUDA events on that stream.  Operates in place on the handle's own workspace for B windows of T frames
 * (run a sampling call first so the workspace holds finite activations).  half_fmt selects the operand format
 * (0 = 3-term fp16 hi/lo split, 1 = single-pass fp16; for FC_LN / W2_LN the fp16 format is the fused GEMM+LayerNorm kernel,
 * the split format times the GEMM and the LayerNorm kernel together). */
enum {
    EGOEGO_KERNEL_START = 0,      /* x-half of start_conv + base/time token                        */
    EGOEGO_KERNEL_QKV = 1,        /* fused Q,K,V projection -> attention operand planes           */
    EGOEGO_KERNEL_ATTENTION = 2,  /* QK^T, softmax, PV for all (window, head)                     */
    EGOEGO_KERNEL_FC_LN = 3,      /* attention fc + bias + residual + LayerNorm                   */
    EGOEGO_KERNEL_W1 = 4,         /* FFN w_1 + bias + ReLU                                        */
    EGOEGO_KERNEL_W2_LN = 5,      /* FFN w_2 + bias + residual + LayerNorm                        */
    EGOEGO_KERNEL_OUT = 6,        /* linear_out                                                   */
    EGOEGO_KERNEL_DDPM_UPDATE = 7 /* clamp + posterior mean + noise + next-step operand staging   */
};
int  egoego_time_kernel(egoego_handle h, int B, int T, int which, int half_fmt, int iters, float* ms_per_launch, void* stream);
/* = egoego_time_kernel(h, B, max T, EGOEGO_KERNEL_QKV, ...) (kept for older callers). */
int  egoego_time_dominant_kernel(egoego_handle h, int B, int half_fmt, int iters, float* ms_per_launch, void* stream);
/* How many times kernel `which` is launched in one step of the sampling loop (EGOEGO_KERNEL_DDPM_UPDATE: 0 when the
 * tensor engine runs the DDPM update in the epilogue of linear_out -- opt-in, EGOEGO_FUSE_DDPM=1 -- and
 * EGOEGO_KERNEL_OUT then times that fused kernel). */
int  egoego_launches_per_step(egoego_handle h, int which);
/* Resolved precision policy: steps t < K read exact (hi/lo) weights (see egoego_cfg.precise_last_steps; how many of them run the
 * full 3-term split is the `split_steps=` field of egoego_engine_info). */
int  egoego_precise_last_steps(egoego_handle h);
/* Human-readable "key=value ..." line of the engine's resolved kernel choices (co-resident cluster counts of the cluster-of-4
 * GEMM+LayerNorm and the cluster-of-8 multicast GEMM, zig-zag tile order, weight sets, precise_last_steps), for logs and bench lines. */
int  egoego_engine_info(egoego_handle h, char* buf, int buf_len);
/* Debug hook for the streamed QKV-projection / attention pair of the fp16-format steps (tools/stream_timeline.py): enable = 1 / 0
 * switches per-CTA recording of %globaltimer on / off (-1 = leave), out (nullable) receives [2 kernels][160 CTAs][start, end] in
 * nanoseconds for the last launch of each kernel (n = number of uint64 to copy, <= 640).  Synchronises the device. */
int  egoego_debug_timeline(int device, int enable, uint64_t* out, int n);
/* Number R of dithered fp16 weight sets the single-pass steps cycle through (env EGOEGO_WEIGHT_SETS at weight commit,
 * default 8, 1 = plain round-to-nearest; always 1 for the fp32 engine).  Step i of the loop reads set i mod R, so the fp16
 * rounding of the weights -- the only rounding of those steps that survives to the final sample -- averages out over steps
 * (DESIGN.md 4).  A loop with fewer than 4 R single-pass steps reads the plain round-to-nearest copy instead (too few steps
 * to average over). */
int  egoego_weight_sets(egoego_handle h);
/* Host-only helper (no GPU needed): the fp16 bit patterns of copy `set` of `n_sets` for n fp32 weights, exactly as
 * egoego_commit_weights rounds them: RN_fp16(w + u ulp16(w)), u = (bitrev(set) + 1/2) / n_sets - 1/2 (u = 0 for n_sets = 1). */
int  egoego_dither_weights_f16(const float* w, int64_t n, int set, int n_sets, uint16_t* out);

/* Self test of the tensor-core GEMM primitive: C = A W^T (A[M,K], W[N,K] random fp32) computed by the
 * tcgen05 3-term fp16 hi/lo split kernel and by the fp32 CUDA-core kernel; reports max |difference|, max |reference|
 * and the average device time of one tcgen05 launch in milliseconds.  M%128 == N%256 == K%64 == 0.
 * two_cta != 0 selects the CTA-pair kernel (cta_group::2, 256x256 tiles; needs M%256 == 0); half_fmt != 0 the
 * single-pass fp16 operand format (errors are then fp16-grade, ~1e-3 relative). */
int  egoego_selftest_gemm(int device, int M, int N, int K, uint64_t seed, int two_cta, int half_fmt,
                          float* max_abs_err, float* max_abs_ref, float* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* EGOEGO_B200_H */
