/*
 * b2unet.h -- C ABI of libb2unet.so: the B200 (sm_100a) implementation of Lifelong-nnUNet's per-step training hot path.
 *
 * The reference (MECLabTUDA/Lifelong-nnUNet @ fb55c48) is pure Python/PyTorch and has NO FFI of its own; each entry
 * point below therefore cites the reference call it replaces (paths relative to the reference's nnunet_ext/).
 * Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller (PyTorch) owns all memory: parameters, workspaces, outputs; the library never allocates device
 *     memory behind the caller's back and never synchronises the host (all work is enqueued on `stream`);
 *   - return value: 0 = ok, negative = error (see b2_last_error()); no C++ exceptions cross the ABI;
 *   - activations live in HBM as NDHWC ("channels-last-3d"), fp32 (B2_F32 parity mode) or bf16 (B2_BF16);
 *     parameters, gradients, logits, losses, Fisher / importance maps are always fp32.
 */
#ifndef B2UNET_H_
#define B2UNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b2_stream_t; /* cudaStream_t */

enum { B2_F32 = 0, B2_BF16 = 1 };
enum { B2_OK = 0, B2_EINVAL = -1, B2_ECUDA = -2, B2_ENOMEM = -3, B2_EUNSUPPORTED = -4 };

/* version / diagnostics */
int b2_version(void);
const char* b2_last_error(void);
/* number of kernel launches issued by this library since process start (bench.py's gpu_launches) */
long long b2_launch_count(void);
/* runtime switches: "tensor_cores" (1 = tcgen05 path for bf16 where supported [default], 0 = SIMT only),
 * "tc_strided" (1 = strided convs also on the tensor-core path).  Must be set before plans are created. */
int b2_set_option(const char* name, int value);

/* ------------------------------------------------------------------------------------------------------------------
 * Network plan.  Replaces the module graph built by nnunet's nnUNetTrainerV2.initialize_network() ->
 * Generic_UNet(...) (ctor args restated at training/network_training/nnViTUNetTrainer.py:101-125) and the forward of
 * network_architecture/generic_ViT_UNet.py:222-230,261-286 (the "copied from original implementation" part).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct b2_unet_geometry {
    int32_t batch;                /* patches per step on this GPU */
    int32_t in_channels;
    int32_t num_classes;
    int32_t base_features;        /* 32 for plans v2.1 */
    int32_t max_features;         /* 320 (3D) */
    int32_t num_pool;             /* number of poolings == number of decoder levels (<= 7) */
    int32_t patch[3];             /* D, H, W */
    int32_t pool[7][3];           /* pool_op_kernel_sizes, stride per axis (1 or 2) */
    int32_t act_dtype;            /* B2_F32 | B2_BF16 : storage type of activations */
    float   lrelu_slope;          /* 1e-2 */
    float   norm_eps;             /* 1e-5 */
} b2_unet_geometry;

typedef struct b2_unet_plan b2_unet_plan;

int b2_unet_plan_create(const b2_unet_geometry* geom, b2_unet_plan** out);
void b2_unet_plan_destroy(b2_unet_plan* plan);

/* Parameter table.  Entry i describes the i-th parameter tensor in the plan's canonical order; `name` is the
 * dotted PyTorch state_dict key of the reference module tree (test/network_architecture/test_MultiHead_Module.py:
 * 281-432), e.g. "conv_blocks_context.0.blocks.0.conv.weight"; shape is the PyTorch shape (fp32, contiguous). */
typedef struct b2_param_info {
    char    name[96];
    int32_t ndim;
    int64_t shape[5];
    int64_t numel;
} b2_param_info;
int b2_unet_num_params(const b2_unet_plan* plan);
int b2_unet_param_info(const b2_unet_plan* plan, int idx, b2_param_info* out);

/* bytes of the activation workspace (holds every tensor kept for backward) and of the scratch workspace */
size_t b2_unet_workspace_bytes(const b2_unet_plan* plan);
/* shape of deep-supervision output `level` (0 = full resolution): writes {D,H,W}; logits are NCDHW fp32 */
int b2_unet_output_shape(const b2_unet_plan* plan, int level, int32_t dhw[3]);

/* forward:  replaces `output = self.network(data)` (training/network_training/multihead/nnUNetTrainerMultiHead.py:
 * 621/633).  params[i] = device pointer of parameter i (fp32).  input = NCDHW fp32 (as the trainer hands it over).
 * logits[l] = caller-allocated NCDHW fp32 output for level l (num_pool entries, highest resolution first, exactly the
 * tuple generic_ViT_UNet.py:282-284 returns).  With keep_for_backward == 0 the pass is inference-only (teacher /
 * validation forwards: plop:258, mib:141, lwf:315-346). */
int b2_unet_forward(b2_unet_plan* plan, const float* const* params, const float* input, void* workspace,
                    float* const* logits, int keep_for_backward, b2_stream_t stream);

/* backward: replaces `l.backward()` through the network (MultiHead:627/639).  dlogits[l] may be NULL (level without
 * loss contribution, e.g. weight 0).  grads[i] = fp32 gradient buffer of parameter i, OVERWRITTEN (not accumulated);
 * parameters that receive no gradient are zero-filled and flagged in has_grad_host[i] = 0 (caller may map that to
 * `param.grad is None`, ewc:300-301).  Bit-reproducible: no floating-point atomics anywhere. */
int b2_unet_backward(b2_unet_plan* plan, const float* const* params, const float* const* dlogits, void* workspace,
                     float* const* grads, int32_t* has_grad_host, b2_stream_t stream);

/* backward with gradient buckets (data-parallel training, SURVEY 8(e): the reference has no collective at all).  The
 * plan differentiates the layers in descending parameter order, so the suffix [first_param, num_params) of the
 * parameter list -- a contiguous tail of the caller's flat gradient arena -- is complete long before the first layers
 * are.  Bucket k's two events (cudaEvent_t, caller-created) are recorded on the caller's stream and on the plan's
 * weight-gradient stream when that suffix is complete; the caller lets its communication stream wait on both and
 * all-reduces the bucket while the remaining layers are still running.  Buckets in descending first_param order. */
typedef struct b2_grad_bucket {
    int32_t first_param;
    void*   event_main;   /* cudaEvent_t, nullable */
    void*   event_side;   /* cudaEvent_t, nullable */
} b2_grad_bucket;
int b2_unet_backward_buckets(b2_unet_plan* plan, const float* const* params, const float* const* dlogits, void* workspace,
                             float* const* grads, int32_t* has_grad_host, const b2_grad_bucket* buckets_host,
                             int n_buckets, b2_stream_t stream);

/* Partial passes for Generic_ViT_UNet (generic_ViT_UNet.py:217-287): the host runs the encoder stages, hands the first
 * skip to the ViT, writes the ViT result over the bottleneck activation (b2_unet_debug_view(2*num_pool+1, 1)) and
 * runs the decoder.  `parts` is a mask of B2_PART_*.  Forward: B2_PART_ENCODER converts the input and runs stages
 * 0..num_pool-1, B2_PART_BOTTLENECK the two bottleneck blocks (V1 discards their result, :230-253, so they may be
 * skipped), B2_PART_DECODER everything from tu[0] on (logits required).  Backward: the B2_PART_DECODER call opens the
 * pass (resets has_grad_host to 1, needs dlogits) and leaves d(bottleneck activation) in debug view (2*num_pool+1, 2);
 * a later call with B2_PART_ENCODER but without B2_PART_BOTTLENECK zero-fills the bottleneck parameters' gradients and
 * flags them has_grad = 0 (`param.grad is None`).  The host adds the ViT's input gradient into the first
 * skip's gradient (debug view (1, 2)) between the two calls.  grads[] of parts not named are left untouched. */
#define B2_PART_ENCODER 1
#define B2_PART_BOTTLENECK 2
#define B2_PART_DECODER 4
#define B2_PART_ALL 7
int b2_unet_forward_parts(b2_unet_plan* plan, const float* const* params, const float* input, void* workspace,
                          float* const* logits, int parts, b2_stream_t stream);
int b2_unet_backward_parts(b2_unet_plan* plan, const float* const* params, const float* const* dlogits, void* workspace,
                           float* const* grads, int32_t* has_grad_host, int parts, b2_stream_t stream);

/* LwF old-task heads (lwf:315-346, helpful_functions.py:226-259): logits of deep-supervision `level` for ANOTHER
 * `seg_outputs[num_pool-1-level]` weight ([num_classes][C] fp32) on the decoder activation the last forward of this plan
 * left in `workspace` -- the stored heads share the body with the running model, so one 1x1x1 launch per old head
 * replaces one full network forward per old head. */
int b2_unet_head_forward(b2_unet_plan* plan, void* workspace, int level, const float* weight, float* logits,
                         b2_stream_t stream);

/* Raw (pre-norm) output of conv module `conv_idx` kept by the last forward, as an NDHWC view into the workspace --
 * what the reference's forward hooks capture (plop:330-353: output.detach() of every conv.Conv* module).
 * conv_idx enumerates modules in named_modules() order of the reference tree; see b2_unet_num_convs/_conv_name. */
typedef struct b2_act_view {
    void*   ptr;
    int32_t n, d, h, w, c;
    int32_t pitch;     /* elements between consecutive voxels (>= c) */
    int32_t dtype;
} b2_act_view;
int b2_unet_num_convs(const b2_unet_plan* plan);
int b2_unet_conv_name(const b2_unet_plan* plan, int conv_idx, char name[96]);
int b2_unet_conv_output(const b2_unet_plan* plan, void* workspace, int conv_idx, b2_act_view* out);
/* activation views (used by the ViT hand-over above and by the debug tools): view of conv block `block_idx` (execution order of the 3x3x3 blocks): which = 0 raw output z,
 * 1 activated output y, 2 gradient wrt y, 3 block input, 4 gradient wrt the block input */
int b2_unet_debug_view(const b2_unet_plan* plan, void* workspace, int block_idx, int which, b2_act_view* out);

/* ------------------------------------------------------------------------------------------------------------------
 * Vision Transformer of Generic_ViT_UNet V1 (network_architecture/vision_transformer.py:218-458; hybrid forward
 * generic_ViT_UNet.py:217-287): 3D patch embedding (Conv3d k = s = patch, :43-50,70-78) of the first skip, class token +
 * position embedding (:423-428), `depth` pre-norm blocks (LayerNorm, qkv, softmax attention, proj, LayerNorm, fc1-GELU-fc2,
 * :186-198), final LayerNorm, class token -> Linear(embed -> out_features) (:436-458); the result IS the bottleneck
 * activation of the U-Net (`x.reshape(size)`, generic_ViT_UNet.py:253).  bf16 operands, fp32 accumulation / residual stream.
 * params / grads: device pointers in named_parameters() order of the reference module: cls_token, pos_embed_0,
 * blocks.layer.i.{norm1.weight, norm1.bias, attn.qkv.weight, attn.qkv.bias, attn.proj.weight, attn.proj.bias, norm2.weight,
 * norm2.bias, mlp.fc1.weight, mlp.fc1.bias, mlp.fc2.weight, mlp.fc2.bias} (i < depth), norm.weight, norm.bias,
 * patch_embeds.0.proj.weight, patch_embeds.0.proj.bias, heads.0.weight, heads.0.bias  (b2_vit_num_params entries).
 * forward: skip0 = NDHWC bf16 view of the ViT input (a plan activation, b2_unet_debug_view(1, 1)); the output goes to
 * `out_view` (NDHWC view of the bottleneck activation, debug view (2*num_pool+1, 1)) and / or `out_dense` ([B][F] fp32).
 * backward: d(output) from `dout_view` (debug view (2*num_pool+1, 2)) or `dout_dense`; parameter gradients OVERWRITTEN;
 * the input gradient is ADDED into `dskip0` (debug view (1, 2)) when given.  Fixed-order reductions: bit-reproducible.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct b2_vit_desc {
    int32_t batch, in_channels, D, H, W;     /* ViT input: first skip of the U-Net */
    int32_t patch;                            /* cubic patch edge (generic_ViT_UNet.py:148) */
    int32_t embed, heads, depth, mlp_ratio;   /* 768 / 12 / 12 / 4 for 'base' (generic_ViT_UNet.py:64-66) */
    int32_t out_features;                     /* == out_c * out_d * out_h * out_w */
    int32_t out_c, out_d, out_h, out_w;       /* bottleneck activation */
    float   ln_eps;                           /* 1e-6 */
    int32_t lsa;                              /* Locality Self-Attention (vision_transformer.py:90-135): qkv without bias (null
                                               * entries in the parameter table), diagonal of the scores masked, one learnable
                                               * temperature vector [heads] per block appended after the tail parameters */
    int32_t lsa_mask;                         /* number of leading tokens whose self-score is masked: the reference sizes its mask
                                               * from the 2-D patch count, (H / p) (W / p) + 1 (vision_transformer.py:289-294) */
} b2_vit_desc;
typedef struct b2_vit_plan b2_vit_plan;
int b2_vit_plan_create(const b2_vit_desc* desc, b2_vit_plan** out);
void b2_vit_plan_destroy(b2_vit_plan* plan);
size_t b2_vit_workspace_bytes(const b2_vit_plan* plan);
int b2_vit_num_params(const b2_vit_plan* plan);
/* byte offset inside the workspace of the patch-embedding output [B * tokens][embed] bf16 (what a forward hook on the
 * patch-embedding Conv3d observes, plop:330-353) */
size_t b2_vit_tokens_offset(const b2_vit_plan* plan);
int b2_vit_forward(b2_vit_plan* plan, const float* const* params, const b2_act_view* skip0, void* workspace,
                   const b2_act_view* out_view, float* out_dense, int keep_for_backward, b2_stream_t stream);
int b2_vit_backward(b2_vit_plan* plan, const float* const* params, const b2_act_view* dout_view, const float* dout_dense,
                    void* workspace, const b2_act_view* dskip0, float* const* grads, b2_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Deep-supervision Dice+CE loss, value and gradient in one sweep.  Replaces nnunet MultipleOutputLoss2(DC_and_CE_loss
 * ({'batch_dice':..,'smooth':1e-5,'do_bg':False},{})) built at MultiHead:1385-1386.
 * logits: (B,C,V) fp32; target: (B,1,V) fp32 class ids; weight = ds weight of this level; dlogits (nullable) gets
 * weight * d(CE + dice)/dlogits.  loss_out[0] += weight * loss (device scalar, caller zeroes it).  scratch >=
 * b2_dsloss_scratch_bytes(B,C,V).
 * ------------------------------------------------------------------------------------------------------------------ */
size_t b2_dsloss_scratch_bytes(int B, int C, int64_t V);
int b2_dsloss_fwd_bwd(const float* logits, const float* target, int B, int C, int64_t V, float weight, int batch_dice,
                      float smooth, int do_bg, int ignore_index /* <0: none */, int with_dice, float* dlogits,
                      float* loss_out, void* scratch, b2_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * EWC / RW quadratic penalty over a list of parameter tensors (multi-tensor, one launch).
 * Replaces the per-tensor Python loops of training/loss_functions/deep_supervision.py:65-80 (EWC:
 * coef = lambda/2, importance = NULL) and :115-132 (RW: coef = lambda, importance = S).
 * value: loss_out[0] += coef * sum_i (F_i [+ S_i]) (theta_i - theta*_i)^2 ; gradient: grads[t][i] += 2 coef (...)(...)
 * (grads may be NULL -> value only).  table_dev: device copy of `n_tensors` b2_pen_entry records.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct b2_pen_entry {
    const float* theta;
    const float* theta_star;
    const float* fisher;
    const float* importance;   /* NULL for EWC */
    float*       grad;         /* NULL -> no gradient written */
    int64_t      numel;
} b2_pen_entry;
size_t b2_quadpen_scratch_bytes(int n_tensors, int64_t total_numel);
int b2_quadpen_fwd_bwd(const b2_pen_entry* table_host, int n_tensors, float coef, float* loss_out, void* scratch,
                       b2_stream_t stream);

/* Fisher / Riemannian-walk updates (multi-tensor, elementwise => bit-stable given bit-stable grads).
 * b2_fisher_square: F = g^2                                   (ewc:303)
 * b2_rw_update:     S += relu(g (prev-theta) / (0.5 F (theta-prev)^2 + eps)) [if prev != NULL]; prev = theta;
 *                   F = alpha g^2 + (1-alpha) F               (rw:240-262) */
typedef struct b2_rw_entry {
    const float* theta;
    const float* grad;
    float*       prev;         /* theta_prev, updated in place */
    float*       fisher;
    float*       score;
    int64_t      numel;
} b2_rw_entry;
int b2_fisher_square(const float* const* grads_host, float* const* fisher_host, const int64_t* numel_host,
                     int n_tensors, void* scratch, b2_stream_t stream);
int b2_rw_update(const b2_rw_entry* table_host, int n_tensors, float alpha, float eps, int have_prev, void* scratch,
                 b2_stream_t stream);
size_t b2_multitensor_scratch_bytes(int n_tensors);

/* clip_grad_norm_(params, max_norm) + SGD(momentum, nesterov, weight_decay) in two launches.
 * Replaces MultiHead:629-630/640-641 with the optimizer of MultiHead:294-301.  norm_out (device float, nullable)
 * receives the total gradient L2 norm before clipping. */
typedef struct b2_sgd_entry {
    float*       theta;
    const float* grad;
    float*       momentum;
    int64_t      numel;
} b2_sgd_entry;
size_t b2_sgd_scratch_bytes(int n_tensors, int64_t total_numel);
int b2_sgd_clip_step(const b2_sgd_entry* table_host, int n_tensors, float lr, float momentum, float weight_decay,
                     int nesterov, float max_norm, int first_step, float* norm_out, void* scratch, b2_stream_t stream);

/* Persistent device tables: the *_host entry points above rebuild and upload their tables on every call (host work and
 * a pageable copy per step, not capturable in a CUDA graph).  A trainer builds each table once (b2_mt_blob_build into
 * b2_mt_blob_bytes of HOST memory), copies it to the device and calls the *_dev variants: no host work per step.
 * `part` = b2_mt_part_bytes(nblocks) of device scratch.  b2_sgd_clip_step_dev reads {lr, momentum, weight_decay,
 * max_norm} from DEVICE memory so that a captured step follows the poly learning-rate schedule (MultiHead:294-301). */
enum { B2_MT_PEN = 0, B2_MT_SGD = 1, B2_MT_RW = 2 };
size_t b2_mt_blob_bytes(int kind, int n_tensors);
int b2_mt_blob_build(int kind, const void* table_host, int n_tensors, void* blob_host, int32_t* nblocks_out);
size_t b2_mt_part_bytes(int nblocks);
int b2_quadpen_dev(const void* blob_dev, int n_tensors, int nblocks, float coef, float* loss_out, float* part,
                   b2_stream_t stream);
int b2_sgd_clip_step_dev(const void* blob_dev, int n_tensors, int nblocks, const float* hyper_dev, int nesterov,
                         float* norm_out, float* part, b2_stream_t stream);
int b2_rw_update_dev(const void* blob_dev, int n_tensors, int nblocks, float alpha, float eps, int have_prev,
                     b2_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Distillation terms (value only or value+gradient), one sweep over student+teacher logits each.
 * b2_kd_lwf : deep_supervision.py:185-199 -- KL(softmax(t/T) || softmax(p/T)), 'batchmean' (sum / B), value only
 *             (both operands are detached in the reference, lwf:343-349).
 * b2_kd_mib : knowledge_distillation.py:11-32 (equal class counts) -- -mean_v (1/C) sum_c softmax(alpha t)_c lsm(x)_c,
 *             value + gradient wrt x scaled by `scale` (= w_i * lkd, deep_supervision.py:411-413).
 * b2_plop_pseudo: deep_supervision.py:287-332 -- pseudo-label CE with entropy thresholds, value + gradient.
 * b2_pod_local:   embeddings.py:9-42 on an NDHWC activation pair, value only (both detached, plop:348-352).
 * ------------------------------------------------------------------------------------------------------------------ */
size_t b2_kd_scratch_bytes(int B, int C, int64_t V);
int b2_kd_lwf(const float* pred, const float* teacher, int B, int C, int64_t V, float temperature, float* loss_out,
              void* scratch, b2_stream_t stream);
int b2_kd_mib(const float* x, const float* teacher, int B, int C, int64_t V, float alpha, float scale, float* dlogits,
              float* loss_out, void* scratch, b2_stream_t stream);
int b2_plop_pseudo(const float* x, const float* x_old, const float* target, int B, int C, int D, int H, int W,
                   const float* thresholds, float max_entropy, float weight, float* dlogits, float* loss_out,
                   void* scratch, b2_stream_t stream);
/* PLOP threshold extraction (plop:113-182): hist[c][bin] += #background voxels (target == 0) with pseudo label c (argmax of
 * softmax(x_old)) whose entropy / max_entropy falls into bin (of nb_bins); integer counts, accumulated into the caller's table. */
int b2_plop_entropy_hist(const float* x_old, const float* target, int B, int C, int64_t V, float max_entropy, int nb_bins,
                         uint64_t* hist /* [C][nb_bins] */, b2_stream_t stream);
size_t b2_pod_scratch_bytes(const b2_act_view* a, int scales);
int b2_pod_local(const b2_act_view* a, const b2_act_view* a_old, int scales, float* value_out, void* scratch,
                 b2_stream_t stream);

/* hard tp/fp/fn per sample and foreground class (MultiHead:938-951): counts_out[(b*(C-1)+c-1)*3 + {0,1,2}] */
int b2_online_eval(const float* logits, const float* target, int B, int C, int64_t V, float* counts_out,
                   void* scratch /* >= b2_kd_scratch_bytes */, b2_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Sliding-window inference aggregation (validation / evaluation sweep; reference inference/predict.py:117-401 and
 * evaluation/evaluator.py drive nnunet's SegmentationNetwork.predict_3D).  One fused launch per predicted patch:
 * agg[c][z0+z][y0+y][x0+x] += scale * gauss[z][y][x] * softmax_c(logits[:, mirror(z, y, x)]);  wsum[..] += gauss (once per
 * patch position: add_weight).  flip_mask bit 0 / 1 / 2: the network saw the patch mirrored along d / h / w (test-time
 * mirroring).  b2_sliding_finalize: agg /= wsum in place (class probabilities), seg = argmax (nullable).
 * ------------------------------------------------------------------------------------------------------------------ */
int b2_sliding_accumulate(const float* logits, int C, int pd, int ph, int pw, const float* gauss, float* agg, float* wsum,
                          int D, int H, int W, int z0, int y0, int x0, int flip_mask, float scale, int add_weight,
                          b2_stream_t stream);
int b2_sliding_finalize(float* agg, const float* wsum, int C, int64_t V, int32_t* seg, b2_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * GPU-side patch pipeline.  Replaces nnunet's DataLoader3D + batchgenerators'
 * get_moreDA_augmentation (reference call sites training/network_training/multihead/nnUNetTrainerMultiHead.py:505-511 and
 * :904-922; both packages un-vendored) with the preprocessed cases resident in HBM.  The HOST draws the random parameters
 * (lifelong-nnunet_b200/b200unet/augment.py) and passes one small record per sample / per (sample, channel); the kernels
 * apply them.  Every `*_host` table is read during the call (kernel arguments), so the caller may reuse it at once.
 * Limits: B <= B2_AUG_MAX_SAMPLES, B * C <= B2_AUG_MAX_BC, <= B2_AUG_MAX_SCALES deep-supervision targets.
 * ------------------------------------------------------------------------------------------------------------------ */
#define B2_AUG_MAX_SAMPLES 32
#define B2_AUG_MAX_BC 64
#define B2_AUG_MAX_SCALES 6
#define B2_AUG_MAX_RADIUS 4
typedef struct {
    const float* volume;     /* device: preprocessed case [C + 1][D][H][W] fp32, LAST channel = segmentation (nnunet .npy layout) */
    int32_t dhw[3];
    int32_t lb[3];           /* crop origin inside the case; may reach outside: data padded with 0, segmentation with -1 */
    int32_t win_lo[3], win_hi[3];   /* only crop voxels inside this window are written (all of it: 0 .. gdhw; a sample without a
                              * spatial transform only needs its centre window) */
} b2_aug_case;
/* DataLoader3D.generate_train_batch: data [B][C][g], seg [B][1][g] = crops of extent gdhw */
int b2_aug_crop(const b2_aug_case* cases_host, int B, int C, const int32_t gdhw[3], float* data, float* seg, b2_stream_t stream);

typedef struct {
    float m[9];              /* source offset (d, h, w) = m (row-major 3x3) . (output voxel - (patch - 1) / 2) */
    float ctr[3];            /* source centre in the crop: extent / 2 - 0.5 (batchgenerators augment_spatial, random_crop=False) */
    int32_t lb[3];           /* modified == 0: plain centre crop at this origin (no interpolation) */
    int32_t modified;
} b2_aug_spatial_params;
/* SpatialTransform (rotation + scaling): data = cubic B-spline interpolation (scipy map_coordinates order 3, mode constant 0;
 * `crop_data` is overwritten by its B-spline coefficients for the modified samples), segmentation = per-label linear
 * interpolation thresholded at 0.5 (batchgenerators interpolate_img, is_seg, order 1, cval -1). */
int b2_aug_spatial(const b2_aug_spatial_params* tf_host, int B, int C, const int32_t gdhw[3], const int32_t pdhw[3], float* crop_data,
                   const float* crop_seg, float* data_out, float* seg_out, b2_stream_t stream);

/* per (sample, channel) statistics {mean, std (ddof 0), min, max} of data [BC][V] -> stats_out [BC][4] */
size_t b2_aug_stats_scratch_bytes(int BC, int64_t V);
int b2_aug_stats(const float* data, int BC, int64_t V, float* stats_out, void* scratch, b2_stream_t stream);

enum { B2_AUG_NONE = 0, B2_AUG_NOISE = 1, B2_AUG_MUL = 2, B2_AUG_CONTRAST = 3, B2_AUG_GAMMA_A = 4, B2_AUG_GAMMA_B = 5 };
typedef struct {
    int32_t op;              /* NOISE: x += p0 * N(0,1) (counter-based generator keyed by seed and element index);  MUL: x *= p0;
                              * CONTRAST: clip((x - mean) * p0 + mean, min, max) with stats_a;
                              * GAMMA_A: ((x - min) / (range + 1e-7)) ^ p0 * (range + 1e-7) + min with stats_a, on -x when p1 != 0;
                              * GAMMA_B (retain_stats): (x - mean_b) / (std_b + 1e-8) * std_a + mean_a, on -x when p1 != 0 */
    float p[3];
} b2_aug_op;
int b2_aug_pointwise(const b2_aug_op* ops_host, int BC, int64_t V, float* data, const float* stats_a, const float* stats_b,
                     uint64_t seed, b2_stream_t stream);

typedef struct {
    int32_t radius;          /* 0: channel left untouched */
    float w[B2_AUG_MAX_RADIUS + 1];   /* w[|t|], normalised (scipy gaussian_filter, truncate 4, reflect boundaries) */
} b2_aug_blur_taps;
int b2_aug_blur(const b2_aug_blur_taps* kernels_host, int BC, const int32_t pdhw[3], float* data, float* tmp, b2_stream_t stream);

/* SimulateLowResolutionTransform on ONE (sample, channel) volume [pd][ph][pw], in place: nearest-neighbour resize down to tdhw, cubic
 * B-spline resize back (both skimage.transform.resize(mode='edge', anti_aliasing=False) == scipy.ndimage.zoom(grid_mode=True,
 * mode='nearest')), result clipped to the value range of the down-sampled volume (skimage clip=True). */
size_t b2_aug_lowres_scratch_bytes(const int32_t pdhw[3]);
int b2_aug_lowres(float* vol, const int32_t pdhw[3], const int32_t tdhw[3], void* scratch, b2_stream_t stream);

/* MirrorTransform (flips bit 0 / 1 / 2 = axis d / h / w) + RemoveLabelTransform(-1, 0) + DownsampleSegForDSTransform2 (order 0:
 * target voxel q reads stride * q + stride / 2): data_out [B][C][p], targets_host[k] [B][1][p / stride_k] */
int b2_aug_finalize(const int32_t* flips_host, int B, int C, const int32_t pdhw[3], const float* data, const float* seg,
                    float* data_out, float* const* targets_host, const int32_t* strides_host /* [n_scales][3] */, int n_scales,
                    b2_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Building blocks, exported so tests can check every kernel against the oracle in isolation.
 * Weight layouts: `w_pt` = PyTorch Conv3d weight [Cout][Cin][3][3][3]; the plan keeps per-step shadows
 * w_f = [27][Cin][Cout] and w_b = [27][Cout][Cin] (activation dtype for tensor-core paths, fp32 otherwise).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct b2_conv_desc {
    int32_t n, d, h, w;       /* input spatial */
    int32_t cin, cout;
    int32_t stride[3];        /* 1 or 2 per axis; kernel 3, pad 1 */
    int32_t in_pitch, out_pitch;
    int32_t dtype;
} b2_conv_desc;
size_t b2_conv3d_scratch_bytes(const b2_conv_desc* d);
/* z = conv3d(x, w, bias); stats[n][c] = {mean, rstd} of z over the spatial axes (for the InstanceNorm that follows) */
int b2_conv3d_fwd(const b2_conv_desc* d, const void* x, const float* w_pt, const float* bias, void* z, float* stats,
                  float eps, void* scratch, b2_stream_t stream);
/* prepared-weights forward (bf16 tensor-core path only): build the [27][Cout][Cin] bf16 weight shadow once with
 * b2_conv3d_make_shadow, then b2_conv3d_fwd_shadow launches exactly one convolution kernel (no statistics). */
size_t b2_conv3d_shadow_bytes(const b2_conv_desc* d);
int b2_conv3d_make_shadow(const b2_conv_desc* d, const float* w_pt, void* shadow, b2_stream_t stream);
int b2_conv3d_fwd_shadow(const b2_conv_desc* d, const void* x, const void* shadow, const float* bias, void* z, void* scratch,
                         b2_stream_t stream);
/* the variant the training step launches: one convolution kernel whose epilogue also writes the InstanceNorm partial
 * sums into `scratch` (bench.py's roofline times this one) */
int b2_conv3d_fwd_shadow_stats(const b2_conv_desc* d, const void* x, const void* shadow, const float* bias, void* z,
                               void* scratch, b2_stream_t stream);
/* dx (nullable) = conv_transpose3d(dz, w); dw, dbias = parameter gradients (PyTorch layouts, overwritten) */
int b2_conv3d_bwd(const b2_conv_desc* d, const void* x, const void* dz, const float* w_pt, void* dx, int accumulate_dx,
                  float* dw, float* dbias, void* scratch, b2_stream_t stream);
/* kernel == stride transposed convolution (nn.ConvTranspose3d(k = s, bias=False), the `tu` modules) as a standalone op on NDHWC
 * buffers: what Generic_ViT_UNet V2 / V3 apply outside the decoder to fuse the bottleneck / the skips into the ViT input
 * (reference generic_ViT_UNet.py:299-338).  w_pt: PyTorch layout [Cin][Cout][kd][kh][kw].  scratch >= b2_tconv3d_scratch_bytes. */
typedef struct b2_tconv_desc {
    int32_t n, d, h, w;       /* input spatial */
    int32_t cin, cout;
    int32_t k[3];             /* kernel == stride, 1 or 2 per axis */
    int32_t in_pitch, out_pitch;
    int32_t dtype;
} b2_tconv_desc;
size_t b2_tconv3d_scratch_bytes(const b2_tconv_desc* d);
int b2_tconv3d_fwd(const b2_tconv_desc* d, const void* x, const float* w_pt, void* y, void* scratch, b2_stream_t stream);
/* dx (nullable, overwritten) = strided conv of dy with w; dw (nullable, overwritten) = weight gradient, fp32 PyTorch layout */
int b2_tconv3d_bwd(const b2_tconv_desc* d, const void* x, const void* dy, const float* w_pt, void* dx, float* dw, void* scratch,
                   b2_stream_t stream);
/* y = lrelu(gamma * (z - mean) * rstd + beta) */
int b2_norm_lrelu_fwd(const void* z, const float* stats, const float* gamma, const float* beta, void* y, int n,
                      int64_t vox, int c, int z_pitch, int y_pitch, int dtype, float slope, b2_stream_t stream);
int b2_norm_lrelu_bwd(const void* z, const void* y, const void* dy, const float* stats, const float* gamma, void* dz,
                      float* dgamma, float* dbeta, int n, int64_t vox, int c, int z_pitch, int y_pitch, int dy_pitch,
                      int dz_pitch, int dtype, float slope, void* scratch, b2_stream_t stream);
size_t b2_norm_scratch_bytes(int n, int64_t vox, int c);

#ifdef __cplusplus
}
#endif
#endif /* B2UNET_H_ */
