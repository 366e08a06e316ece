// plan.cu -- the network plan (layer graph, HBM layout of the workspace) and the forward / backward orchestration of
// the 3D Generic_UNet, plus the C ABI of include/b2unet.h.
//
// Graph semantics: nnunet@77bc485 Generic_UNet as constructed by nnUNetTrainerV2.initialize_network (SURVEY.md
// Appendix A; ctor args restated at reference nnunet_ext/training/network_training/nnViTUNetTrainer.py:101-125;
// forward restated at reference nnunet_ext/network_architecture/generic_ViT_UNet.py:222-230,261-286).
//
// HBM layout of the caller-provided workspace (one blob, offsets fixed at plan creation):
//   [activations: NDHWC, act dtype]   x_in, per conv block raw output z and activated output y; the encoder skips and
//                                      the transposed-conv outputs live side by side in per-level "concat" buffers
//                                      (pitch 2C) so torch.cat never happens
//   [activation gradients, act dtype]  one buffer per y tensor (+ concat buffers) and one reusable dz buffer
//   [fp32 statistics]                  per conv block (n,c) {mean, rstd}
//   [fp32 weight shadows]              per conv: Wf [27][Cin][Cout], Wb [27][Cout][Cin]; per tconv: Wq [K8][Cin][Cout]
//   [fp32 scratch]                     partial sums of the ordered reductions (max over layers, reused in stream order)
#include <string.h>

#include <array>
#include <type_traits>

#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace b2 {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
int g_use_tc = 1;
int g_tc_strided = 1;
int g_tc_wgrad = 1;
int g_pdl = 0;              // programmatic dependent launch (common.cuh); measured: no gain on the step (293 vs 296 patches/s), off by default
int g_norm_recompute = 1;   // norm backward recomputes the activation sign from z instead of reading y

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

struct Act {
    size_t off = 0;  // element offset inside the region (act or grad)
    int c = 0, pitch = 0;
    int n = 0, d = 0, h = 0, w = 0;
    long long vox() const { return (long long)d * h * w; }
    size_t elems() const { return (size_t)n * vox() * pitch; }
};

struct ConvBlock {
    std::string prefix;  // e.g. "conv_blocks_context.0.blocks.0"
    ConvShape shape;
    Act in, z, y;        // activations
    Act din, dy;         // gradients (din.c == 0 -> no dgrad)
    int din_accumulate = 0;
    int p_w = -1, p_b = -1, p_g = -1, p_be = -1;
    size_t stats_off = 0, wf_off = 0, wb_off = 0;  // fp32 offsets
    size_t wk_off = 0, wd_off = 0;                 // bf16 shadows for the tensor-core path (offsets in floats)
    bool tc_fwd = false, tc_dgrad = false, tc_wgrad = false, tc_dgrad_strided = false;
    int cat_level = -1;        // >= 0: the input is the concat buffer of that level (decoder): its data gradient can be split
    int skip_wait_level = -1;  // >= 0: din (accumulate) / dy is the skip half of that level's concat gradient
};

struct Tconv {
    std::string prefix;  // "tu.0"
    TconvShape shape;
    Act in, out, din, dout;
    int p_w = -1;
    size_t wq_off = 0, wqb_off = 0, wqd_off = 0;   // fp32 shadow; bf16 shadows for the tensor-core path
    bool tc = false;
};

struct Head {
    std::string prefix;  // "seg_outputs.0"
    Act in, din;
    int c = 0, level = 0;
    int p_w = -1;
};

struct ParamInfo { std::string name; std::vector<int64_t> shape; int64_t numel; };

}  // namespace b2

using namespace b2;

struct b2_unet_plan {
    b2_unet_geometry g;
    std::vector<int> feats;
    std::vector<std::array<int, 3>> sizes;
    std::vector<ConvBlock> convs;      // execution order of the 3x3x3 blocks
    std::vector<Tconv> tconvs;
    std::vector<Head> heads;
    std::vector<ParamInfo> params;
    // execution order of conv modules for hooks: entries (kind, index): 0 conv block, 1 tconv
    std::vector<std::pair<int, int>> conv_modules;
    Act x_in, dz_tmp, patch;          // patch: [voxels][32] im2col matrix of the first layer (bf16 tensor-core path)
    bool first_tc = false;
    size_t wp_off = 0;
    size_t act_elems = 0, grad_elems = 0, f32_floats = 0, scratch_floats = 0, wg_scratch_floats = 0;
    size_t off_act = 0, off_grad = 0, off_f32 = 0, off_scratch = 0, off_wg_scratch = 0, total_bytes = 0;
    int esz = 4;
    // backward overlap: the weight-gradient kernels of a layer run on a side stream next to its data-gradient kernels
    // (dz double-buffered, own scratch); created lazily, joined before b2_unet_backward returns to the caller's stream
    Act dz_tmp2;
    bool patch_valid = false;   // the first layer's patch matrix matches the current input (built by the GEMM forward or lazily in backward)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_dz[2] = {nullptr, nullptr}, ev_wg[2] = {nullptr, nullptr}, ev_misc = nullptr, ev_skip[8] = {};
    bool side_ok = false;
    ~b2_unet_plan() {
        for (int i = 0; i < 2; ++i) {
            if (ev_dz[i]) cudaEventDestroy(ev_dz[i]);
            if (ev_wg[i]) cudaEventDestroy(ev_wg[i]);
        }
        if (ev_misc) cudaEventDestroy(ev_misc);
        for (int i = 0; i < 8; ++i) if (ev_skip[i]) cudaEventDestroy(ev_skip[i]);
        if (side) cudaStreamDestroy(side);
    }
};

namespace b2 {

static int add_param(b2_unet_plan* p, const std::string& name, std::vector<int64_t> shape) {
    ParamInfo pi;
    pi.name = name;
    pi.shape = shape;
    pi.numel = 1;
    for (auto s : shape) pi.numel *= s;
    p->params.push_back(pi);
    return (int)p->params.size() - 1;
}

static Act alloc_act(size_t& cursor, int n, int d, int h, int w, int c, int pitch = 0) {
    Act a;
    a.n = n; a.d = d; a.h = h; a.w = w; a.c = c; a.pitch = pitch ? pitch : c;
    a.off = cursor;
    cursor += (a.elems() + 63) / 64 * 64;
    return a;
}

static Act slice(const Act& base, int c0, int c) {
    Act a = base;
    a.off = base.off + c0;
    a.c = c;
    return a;
}

template <typename T> static inline T* P(void* ws, const b2_unet_plan* p, const Act& a, bool grad) {
    return reinterpret_cast<T*>((char*)ws + (grad ? p->off_grad : p->off_act)) + a.off;
}
static inline float* F32(void* ws, const b2_unet_plan* p, size_t off) { return reinterpret_cast<float*>((char*)ws + p->off_f32) + off; }
static inline float* SCR(void* ws, const b2_unet_plan* p) { return reinterpret_cast<float*>((char*)ws + p->off_scratch); }
static inline float* SCR_WG(void* ws, const b2_unet_plan* p) { return reinterpret_cast<float*>((char*)ws + p->off_wg_scratch); }

int g_bwd_overlap = 1;
int g_dgrad_split = 0;      // decoder convs on a concat input: skip half of the data gradient on the side stream -- measured
                            // slightly slower (307 vs 313 patches/s: two N = 96 launches cost more than one N = 192), off by default

static bool ensure_side_stream(b2_unet_plan* p) {
    if (p->side_ok) return true;
    if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess) return false;
    for (int i = 0; i < 2; ++i) {
        if (cudaEventCreateWithFlags(&p->ev_dz[i], cudaEventDisableTiming) != cudaSuccess) return false;
        if (cudaEventCreateWithFlags(&p->ev_wg[i], cudaEventDisableTiming) != cudaSuccess) return false;
    }
    if (cudaEventCreateWithFlags(&p->ev_misc, cudaEventDisableTiming) != cudaSuccess) return false;
    for (int i = 0; i < 8; ++i)
        if (cudaEventCreateWithFlags(&p->ev_skip[i], cudaEventDisableTiming) != cudaSuccess) return false;
    p->side_ok = true;
    return true;
}

static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

static int build_plan(b2_unet_plan* p) {
    const b2_unet_geometry& g = p->g;
    const int P_ = g.num_pool, N = g.batch;
    p->esz = g.act_dtype == B2_BF16 ? 2 : 4;
    p->feats.clear();
    p->sizes.clear();
    int f = g.base_features;
    for (int d = 0; d <= P_; ++d) {
        p->feats.push_back(f < g.max_features ? f : g.max_features);
        f = (int)(f * 2);
    }
    {
        int d = g.patch[0], h = g.patch[1], w = g.patch[2];
        p->sizes.push_back({d, h, w});
        for (int i = 0; i < P_; ++i) {
            if (d % g.pool[i][0] || h % g.pool[i][1] || w % g.pool[i][2]) return fail(B2_EINVAL, "patch not divisible by pool strides%s", "");
            d /= g.pool[i][0]; h /= g.pool[i][1]; w /= g.pool[i][2];
            p->sizes.push_back({d, h, w});
        }
    }
    size_t ac = 0, gc = 0, fc = 0;
    p->x_in = alloc_act(ac, N, g.patch[0], g.patch[1], g.patch[2], g.in_channels);
    p->first_tc = g.act_dtype == B2_BF16 && g_use_tc && g_tc_wgrad && first_layer_tc_supported(g.in_channels, p->feats[0]);
    if (p->first_tc) {
        p->patch = alloc_act(ac, N, g.patch[0], g.patch[1], g.patch[2], 32);
        p->wp_off = fc; fc += ((size_t)p->feats[0] * 32 / 2 + 63) / 64 * 64;
        p->scratch_floats = max_sz(p->scratch_floats, first_layer_wgrad_part_floats(N, g.patch[0], g.patch[1], g.patch[2], p->feats[0]));
        p->wg_scratch_floats = max_sz(p->wg_scratch_floats, first_layer_wgrad_part_floats(N, g.patch[0], g.patch[1], g.patch[2], p->feats[0]));
        p->scratch_floats = max_sz(p->scratch_floats, instnorm_stats_scratch_floats(N, p->x_in.vox(), p->feats[0]));
    }
    // concat buffers per encoder level 0..P-1 (activation + gradient)
    std::vector<Act> cat(P_), dcat(P_);
    for (int l = 0; l < P_; ++l) {
        cat[l] = alloc_act(ac, N, p->sizes[l][0], p->sizes[l][1], p->sizes[l][2], 2 * p->feats[l]);
        dcat[l] = alloc_act(gc, N, p->sizes[l][0], p->sizes[l][1], p->sizes[l][2], 2 * p->feats[l]);
    }
    size_t max_z = 0;
    auto add_conv = [&](const std::string& prefix, const Act& in, const Act& din, int din_acc, int cout, const int stride[3],
                        const Act* y_fixed, const Act* dy_fixed) {
        ConvBlock cb;
        cb.prefix = prefix;
        cb.shape.n = N; cb.shape.d = in.d; cb.shape.h = in.h; cb.shape.w = in.w;
        cb.shape.cin = in.c; cb.shape.cout = cout;
        for (int i = 0; i < 3; ++i) cb.shape.stride[i] = stride[i];
        const int od = (in.d - 1) / stride[0] + 1, oh = (in.h - 1) / stride[1] + 1, ow = (in.w - 1) / stride[2] + 1;
        cb.in = in;
        cb.din = din;
        cb.din_accumulate = din_acc;
        cb.z = alloc_act(ac, N, od, oh, ow, cout);
        if (y_fixed) { cb.y = *y_fixed; cb.dy = *dy_fixed; }
        else { cb.y = alloc_act(ac, N, od, oh, ow, cout); cb.dy = alloc_act(gc, N, od, oh, ow, cout); }
        cb.shape.in_pitch = in.pitch;
        cb.shape.out_pitch = cb.z.pitch;
        cb.p_w = add_param(p, prefix + ".conv.weight", {cout, in.c, 3, 3, 3});
        cb.p_b = add_param(p, prefix + ".conv.bias", {cout});
        cb.p_g = add_param(p, prefix + ".instnorm.weight", {cout});
        cb.p_be = add_param(p, prefix + ".instnorm.bias", {cout});
        cb.stats_off = fc; fc += (size_t)N * cout * 2;
        cb.wf_off = fc; fc += (size_t)27 * in.c * cout;
        cb.wb_off = fc; fc += (size_t)27 * in.c * cout;
        fc = (fc + 63) / 64 * 64;
        const bool strided = stride[0] != 1 || stride[1] != 1 || stride[2] != 1;
        const bool tc_ok = g.act_dtype == B2_BF16 && g_use_tc && conv_tc_supported(in.c, cout) && in.pitch % 8 == 0;
        cb.tc_fwd = tc_ok && (!strided || g_tc_strided);
        cb.tc_dgrad = tc_ok && !strided && din.c > 0 && din.pitch % 8 == 0;
        cb.tc_dgrad_strided = tc_ok && strided && g_tc_strided && din.c > 0 && din.pitch % 8 == 0;
        cb.tc_wgrad = tc_ok && g_tc_wgrad && wgrad_tc_supported(in.c, cout) && (!strided || g_tc_strided);
        if (cb.tc_wgrad) {
            p->scratch_floats = max_sz(p->scratch_floats, wgrad_tc_part_floats(cb.shape));
            p->wg_scratch_floats = max_sz(p->wg_scratch_floats, wgrad_tc_part_floats(cb.shape));
        }
        if (tc_ok) {
            p->scratch_floats = max_sz(p->scratch_floats, conv_tc_splitk_scratch_floats(N, od, oh, ow, cout));
            p->scratch_floats = max_sz(p->scratch_floats, conv_tc_splitk_scratch_floats(N, in.d, in.h, in.w, in.c));
            cb.wk_off = fc; fc += ((size_t)27 * in.c * cout / 2 + 63) / 64 * 64;
            cb.wd_off = fc; fc += ((size_t)27 * in.c * cout / 2 + 63) / 64 * 64;
            p->scratch_floats = max_sz(p->scratch_floats, instnorm_stats_scratch_floats(N, (long long)od * oh * ow, cout));
            p->scratch_floats = max_sz(p->scratch_floats, (size_t)N * num_sms() * 8 * cout * 2);   // epilogue statistics partials (8 epilogue warps per CTA)
        }
        max_z = max_sz(max_z, cb.z.elems());
        p->scratch_floats = max_sz(p->scratch_floats, conv_stat_part_floats(cb.shape));
        p->scratch_floats = max_sz(p->scratch_floats, conv_wgrad_part_floats(cb.shape));
        p->wg_scratch_floats = max_sz(p->wg_scratch_floats, conv_wgrad_part_floats(cb.shape));
        p->scratch_floats = max_sz(p->scratch_floats, norm_bwd_scratch_floats(N, cb.z.vox(), cout));
        p->convs.push_back(cb);
        p->conv_modules.push_back({0, (int)p->convs.size() - 1});
        return p->convs.back();
    };
    const int one[3] = {1, 1, 1};
    Act cur = p->x_in, dcur;  // dcur.c == 0: no gradient wrt the network input
    char buf[128];
    for (int d = 0; d <= P_; ++d) {
        const int* st0 = d == 0 ? one : g.pool[d - 1];
        const int fo = p->feats[d];
        const bool skip_level = d < P_;
        Act ysk, dysk;
        if (skip_level) { ysk = slice(cat[d], fo, fo); dysk = slice(dcat[d], fo, fo); }
        if (d < P_) snprintf(buf, sizeof(buf), "conv_blocks_context.%d.blocks.0", d);
        else snprintf(buf, sizeof(buf), "conv_blocks_context.%d.0.blocks.0", d);
        // gradient wrt the stage input accumulates into the skip gradient of the previous level (already holding the
        // decoder's contribution)
        ConvBlock b0 = add_conv(buf, cur, dcur, d > 0 ? 1 : 0, fo, st0, nullptr, nullptr);
        if (d > 0) p->convs.back().skip_wait_level = d - 1;     // accumulates into the skip gradient of level d-1
        if (d < P_) snprintf(buf, sizeof(buf), "conv_blocks_context.%d.blocks.1", d);
        else snprintf(buf, sizeof(buf), "conv_blocks_context.%d.1.blocks.0", d);
        ConvBlock b1 = add_conv(buf, b0.y, b0.dy, 0, fo, one, skip_level ? &ysk : nullptr, skip_level ? &dysk : nullptr);
        if (skip_level) p->convs.back().skip_wait_level = d;    // its dy is the skip gradient of level d
        cur = b1.y;
        dcur = b1.dy;
    }
    // decoder
    for (int u = 0; u < P_; ++u) {
        const int lvl = P_ - 1 - u;
        const int fs = p->feats[lvl];
        Tconv t;
        snprintf(buf, sizeof(buf), "tu.%d", u);
        t.prefix = buf;
        t.shape.n = N; t.shape.d = cur.d; t.shape.h = cur.h; t.shape.w = cur.w;
        t.shape.cin = cur.c; t.shape.cout = fs;
        for (int i = 0; i < 3; ++i) t.shape.k[i] = g.pool[lvl][i];
        t.in = cur; t.din = dcur;
        t.out = slice(cat[lvl], 0, fs);
        t.dout = slice(dcat[lvl], 0, fs);
        t.shape.in_pitch = cur.pitch; t.shape.out_pitch = t.out.pitch;
        const int k8 = t.shape.k[0] * t.shape.k[1] * t.shape.k[2];
        t.p_w = add_param(p, t.prefix + ".weight", {cur.c, fs, t.shape.k[0], t.shape.k[1], t.shape.k[2]});
        t.wq_off = fc; fc += (size_t)k8 * cur.c * fs;
        fc = (fc + 63) / 64 * 64;
        t.tc = g.act_dtype == B2_BF16 && g_use_tc && cur.c % 32 == 0 && fs % 32 == 0 && cur.pitch % 8 == 0 && dcur.pitch % 8 == 0;
        if (t.tc && g_tc_wgrad && tconv_wgrad_tc_supported(cur.c, fs))
        {
            p->scratch_floats = max_sz(p->scratch_floats, tconv_wgrad_tc_part_floats(t.shape));
            p->wg_scratch_floats = max_sz(p->wg_scratch_floats, tconv_wgrad_tc_part_floats(t.shape));
        }
        if (t.tc) {
            t.wqb_off = fc; fc += ((size_t)k8 * cur.c * fs / 2 + 63) / 64 * 64;
            t.wqd_off = fc; fc += ((size_t)k8 * cur.c * fs / 2 + 63) / 64 * 64;
        }
        p->scratch_floats = max_sz(p->scratch_floats, tconv_bwd_scratch_floats(t.shape));
        p->tconvs.push_back(t);
        p->conv_modules.push_back({1, (int)p->tconvs.size() - 1});
        snprintf(buf, sizeof(buf), "conv_blocks_localization.%d.0.blocks.0", u);
        ConvBlock l0 = add_conv(buf, cat[lvl], dcat[lvl], 0, fs, one, nullptr, nullptr);
        p->convs.back().cat_level = lvl;
        p->wg_scratch_floats = max_sz(p->wg_scratch_floats, conv_tc_splitk_scratch_floats(N, cat[lvl].d, cat[lvl].h, cat[lvl].w, fs));
        snprintf(buf, sizeof(buf), "conv_blocks_localization.%d.1.blocks.0", u);
        ConvBlock l1 = add_conv(buf, l0.y, l0.dy, 0, fs, one, nullptr, nullptr);
        cur = l1.y;
        dcur = l1.dy;
        Head hd;
        snprintf(buf, sizeof(buf), "seg_outputs.%d", u);
        hd.prefix = buf;
        hd.in = cur; hd.din = dcur; hd.c = fs; hd.level = lvl;
        hd.p_w = add_param(p, hd.prefix + ".weight", {g.num_classes, fs, 1, 1, 1});
        p->scratch_floats = max_sz(p->scratch_floats, seghead_bwd_scratch_floats(N, cur.vox(), fs, g.num_classes));
        p->wg_scratch_floats = max_sz(p->wg_scratch_floats, seghead_bwd_scratch_floats(N, cur.vox(), fs, g.num_classes));
        p->heads.push_back(hd);
    }
    p->dz_tmp.off = gc;
    gc += (max_z + 63) / 64 * 64;
    p->dz_tmp2.off = gc;
    gc += (max_z + 63) / 64 * 64;
    p->act_elems = ac; p->grad_elems = gc; p->f32_floats = fc;
    p->off_act = 0;
    p->off_grad = align_up(ac * p->esz, 1024);
    p->off_f32 = p->off_grad + align_up(gc * p->esz, 1024);
    p->off_scratch = p->off_f32 + align_up(fc * 4, 1024);
    p->off_wg_scratch = p->off_scratch + align_up(p->scratch_floats * 4, 1024);
    p->total_bytes = p->off_wg_scratch + align_up(p->wg_scratch_floats * 4, 1024);
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
template <typename T>
static int forward_t(b2_unet_plan* p, const float* const* prm, const float* input, void* ws, float* const* logits,
                     int parts, cudaStream_t st) {
    const b2_unet_geometry& g = p->g;
    int rc;
    if (parts & B2_PART_ENCODER)
        if ((rc = nchw_to_ndhwc<T>(input, P<T>(ws, p, p->x_in, false), g.batch, g.in_channels, p->x_in.vox(), p->x_in.pitch, st))) return rc;
    size_t ci = 0, ti = 0;
    // bf16 weight shadows of every layer this call runs: ONE multi-tensor launch (tiled transposes) instead of one
    // scatter kernel per layer
    bool shadows_batched = false, fwd_side = false, shadows_pending = false, heads_on_side = false;
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
        std::vector<ShadowJob> jobs;
        bool all_ok = true;
        auto add_conv_job = [&](ConvBlock& cb) {
            if (p->first_tc && &cb == &p->convs[0]) return;
            if (!(cb.tc_fwd || cb.tc_dgrad || cb.tc_dgrad_strided)) return;
            if (!shadow_job_supported(cb.shape.cout, cb.shape.cin, 27)) { all_ok = false; return; }
            jobs.push_back(ShadowJob{prm[cb.p_w], (__nv_bfloat16*)F32(ws, p, cb.wk_off), (__nv_bfloat16*)F32(ws, p, cb.wd_off),
                                     cb.shape.cout, cb.shape.cin, 27, 1});
        };
        for (int d = 0; d <= g.num_pool; ++d) {
            const bool on = d < g.num_pool ? (parts & B2_PART_ENCODER) != 0 : (parts & B2_PART_BOTTLENECK) != 0;
            if (!on) continue;
            add_conv_job(p->convs[2 * d]);
            add_conv_job(p->convs[2 * d + 1]);
        }
        if (parts & B2_PART_DECODER)
            for (int u = 0; u < g.num_pool; ++u) {
                Tconv& t = p->tconvs[u];
                const int k8 = t.shape.k[0] * t.shape.k[1] * t.shape.k[2];
                if (t.tc) {
                    if (!shadow_job_supported(t.shape.cin, t.shape.cout, k8)) all_ok = false;
                    else jobs.push_back(ShadowJob{prm[t.p_w], (__nv_bfloat16*)F32(ws, p, t.wqd_off), (__nv_bfloat16*)F32(ws, p, t.wqb_off),
                                                  t.shape.cin, t.shape.cout, k8, 0});
                }
                add_conv_job(p->convs[2 * (g.num_pool + 1) + 2 * u]);
                add_conv_job(p->convs[2 * (g.num_pool + 1) + 2 * u + 1]);
            }
        if (all_ok && !jobs.empty()) {
            // on the side stream: the input layout conversion and the first layer do not need the shadows
            fwd_side = g_bwd_overlap && ensure_side_stream(p);
            if (fwd_side) {
                B2_CUDA(cudaEventRecord(p->ev_misc, st));          // (orders the shadow writes after everything already queued on st)
                B2_CUDA(cudaStreamWaitEvent(p->side, p->ev_misc, 0));
            }
            if ((rc = shadow_multi(jobs.data(), (int)jobs.size(), fwd_side ? p->side : st))) return rc;
            if (fwd_side) { B2_CUDA(cudaEventRecord(p->ev_dz[0], p->side)); shadows_pending = true; }
            shadows_batched = true;
        }
    }
    auto run_conv = [&](ConvBlock& cb) -> int {
        float* wf = F32(ws, p, cb.wf_off);
        float* wb = F32(ws, p, cb.wb_off);
        float* stats = F32(ws, p, cb.stats_off);
        bool need_wf = true, need_wb = cb.din.c > 0;
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
            need_wf = !cb.tc_fwd && !(p->first_tc && &cb == &p->convs[0]);
            need_wb = cb.din.c > 0 && !(cb.tc_dgrad || cb.tc_dgrad_strided);
        }
        int r = weight_shadow(prm[cb.p_w], cb.shape.cout, cb.shape.cin, need_wf ? wf : nullptr, need_wb ? wb : nullptr, st);
        if (r) return r;
        bool done = false;
        SplitKDefer defer;
        defer.deferred = 0;
        // deep stages: statistics + normalisation + activation in one launch (norm.cu, small tensors)
        const bool small = norm_small_supported(cb.z.vox(), cb.shape.cout, cb.z.pitch, cb.y.pitch, 8, 8);
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
            if (p->first_tc && &cb == &p->convs[0]) {
                p->patch_valid = true;
                __nv_bfloat16* P_ = P<T>(ws, p, p->patch, false);
                __nv_bfloat16* wp = (__nv_bfloat16*)F32(ws, p, p->wp_off);
                r = first_layer_patches(P<T>(ws, p, cb.in, false), g.batch, cb.in.d, cb.in.h, cb.in.w, cb.shape.cin, cb.in.pitch, P_,
                                        prm[cb.p_w], cb.shape.cout, wp, st);
                if (r) return r;
                TcGather tg;
                memset(&tg, 0, sizeof(tg));
                tg.src = P_; tg.N = g.batch; tg.Ds = cb.in.d; tg.Hs = cb.in.h; tg.Ws = cb.in.w; tg.K = 32; tg.src_pitch = 32;
                tg.wmat = wp; tg.w_rows = cb.shape.cout; tg.rows_per_tap = cb.shape.cout; tg.Nout = cb.shape.cout; tg.bias = prm[cb.p_b];
                tg.dst = P<T>(ws, p, cb.z, false); tg.Dd = cb.z.d; tg.Hd = cb.z.h; tg.Wd = cb.z.w; tg.dst_pitch = cb.z.pitch;
                tg.LD = cb.z.d; tg.LH = cb.z.h; tg.LW = cb.z.w;
                for (int a = 0; a < 3; ++a) { tg.stride[a] = 1; tg.os[a] = 1; tg.qk[a] = 1; }
                tg.ntaps = 1;
                int stat_slots = 0;
                if (!small) { tg.stat_part = SCR(ws, p); tg.stat_part_floats = p->scratch_floats; tg.stat_slots = &stat_slots; }
                r = conv_tc_gather(tg, st);
                if (r) return r;
                if (stat_slots > 0) r = stats_finalize(SCR(ws, p), stat_slots, g.batch, cb.z.vox(), cb.shape.cout, g.norm_eps, stats, st);
                else if (!small) r = instnorm_stats<T>(P<T>(ws, p, cb.z, false), g.batch, cb.z.vox(), cb.shape.cout, cb.z.pitch, SCR(ws, p), stats, g.norm_eps, st);
                if (r) return r;
                done = true;
            }
            if (!done && shadows_pending) {   // first consumer of the batched shadows: join the side stream
                B2_CUDA(cudaStreamWaitEvent(st, p->ev_dz[0], 0));
                shadows_pending = false;
            }
            if (!done && !shadows_batched && (cb.tc_fwd || cb.tc_dgrad || cb.tc_dgrad_strided)) {
                r = weight_shadow_bf16(prm[cb.p_w], cb.shape.cout, cb.shape.cin, (__nv_bfloat16*)F32(ws, p, cb.wk_off),
                                       (__nv_bfloat16*)F32(ws, p, cb.wd_off), st);
                if (r) return r;
            }
            if (cb.tc_fwd) {
                int stat_slots = 0;
                r = conv_tc_launch(P<T>(ws, p, cb.in, false), g.batch, cb.in.d, cb.in.h, cb.in.w, cb.shape.cin, cb.in.pitch,
                                   (const __nv_bfloat16*)F32(ws, p, cb.wk_off), cb.shape.cout, prm[cb.p_b], P<T>(ws, p, cb.z, false),
                                   cb.z.d, cb.z.h, cb.z.w, cb.z.pitch, cb.shape.stride, 0, st, SCR(ws, p), p->scratch_floats * sizeof(float),
                                   small ? nullptr : &stat_slots, 0, 0, (small && g_splitk_fuse) ? &defer : nullptr);
                if (r) return r;
                if (stat_slots > 0) r = stats_finalize(SCR(ws, p), stat_slots, g.batch, cb.z.vox(), cb.shape.cout, g.norm_eps, stats, st);
                else if (!small) r = instnorm_stats<T>(P<T>(ws, p, cb.z, false), g.batch, cb.z.vox(), cb.shape.cout, cb.z.pitch, SCR(ws, p), stats, g.norm_eps, st);
                if (r) return r;
                done = true;
            }
        }
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
            if (done && small && defer.deferred)     // the convolution left its split-K partials un-reduced: sum them here
                return splitk_norm_small_fwd(defer, prm[cb.p_b], P<T>(ws, p, cb.z, false), prm[cb.p_g], prm[cb.p_be], P<T>(ws, p, cb.y, false),
                                             stats, g.batch, cb.z.d, cb.z.h, cb.z.w, cb.shape.cout, cb.z.pitch, cb.y.pitch, g.lrelu_slope,
                                             g.norm_eps, st);
        }
        if (done && small)
            return norm_lrelu_fwd_small<T>(P<T>(ws, p, cb.z, false), prm[cb.p_g], prm[cb.p_be], P<T>(ws, p, cb.y, false), stats, g.batch,
                                           cb.z.vox(), cb.shape.cout, cb.z.pitch, cb.y.pitch, g.lrelu_slope, g.norm_eps, st);
        if (!done) {
            r = conv3d_fwd_simt<T>(cb.shape, P<T>(ws, p, cb.in, false), wf, prm[cb.p_b], P<T>(ws, p, cb.z, false), SCR(ws, p), stats, g.norm_eps, st);
            if (r) return r;
        }
        return norm_lrelu_fwd<T>(P<T>(ws, p, cb.z, false), stats, prm[cb.p_g], prm[cb.p_be], P<T>(ws, p, cb.y, false), g.batch,
                                 cb.z.vox(), cb.shape.cout, cb.z.pitch, cb.y.pitch, g.lrelu_slope, st);
    };
    for (int d = 0; d <= g.num_pool; ++d) {
        const bool on = d < g.num_pool ? (parts & B2_PART_ENCODER) != 0 : (parts & B2_PART_BOTTLENECK) != 0;
        if (!on) { ci += 2; continue; }
        if ((rc = run_conv(p->convs[ci++]))) return rc;
        if ((rc = run_conv(p->convs[ci++]))) return rc;
    }
    if (!(parts & B2_PART_DECODER)) return B2_OK;
    for (int u = 0; u < g.num_pool; ++u) {
        Tconv& t = p->tconvs[ti++];
        float* wq = F32(ws, p, t.wq_off);
        const int k8 = t.shape.k[0] * t.shape.k[1] * t.shape.k[2];
        bool tdone = false;
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
            if (t.tc) {
                if (shadows_pending) { B2_CUDA(cudaStreamWaitEvent(st, p->ev_dz[0], 0)); shadows_pending = false; }
                if (!shadows_batched)
                    if ((rc = tconv_shadow_bf16(prm[t.p_w], t.shape.cin, t.shape.cout, k8, (__nv_bfloat16*)F32(ws, p, t.wqb_off),
                                                (__nv_bfloat16*)F32(ws, p, t.wqd_off), st))) return rc;
                if ((rc = tconv_tc_fwd(P<T>(ws, p, t.in, false), g.batch, t.in.d, t.in.h, t.in.w, t.shape.cin, t.in.pitch,
                                       (const __nv_bfloat16*)F32(ws, p, t.wqb_off), t.shape.cout, t.shape.k, P<T>(ws, p, t.out, false),
                                       t.out.pitch, st))) return rc;
                tdone = true;
            }
        }
        if (!tdone) {
            if ((rc = tconv_shadow(prm[t.p_w], t.shape.cin, t.shape.cout, k8, wq, st))) return rc;
            if ((rc = tconv_fwd_q<T>(t.shape, P<T>(ws, p, t.in, false), wq, P<T>(ws, p, t.out, false), st))) return rc;
        }
        if ((rc = run_conv(p->convs[ci++]))) return rc;
        if ((rc = run_conv(p->convs[ci++]))) return rc;
        Head& h = p->heads[u];
        if (logits[h.level]) {
            // the heads of the lower-resolution levels run next to the rest of the decoder; the last one stays on st
            cudaStream_t hs = st;
            if (fwd_side && u + 1 < g.num_pool) {
                B2_CUDA(cudaEventRecord(p->ev_misc, st));
                B2_CUDA(cudaStreamWaitEvent(p->side, p->ev_misc, 0));
                hs = p->side;
                heads_on_side = true;
            }
            if ((rc = seghead_fwd<T>(P<T>(ws, p, h.in, false), prm[h.p_w], logits[h.level], g.batch, h.in.vox(), h.c, g.num_classes, h.in.pitch, hs))) return rc;
        }
    }
    if (shadows_pending) { B2_CUDA(cudaStreamWaitEvent(st, p->ev_dz[0], 0)); shadows_pending = false; }
    if (heads_on_side) {   // join: every logits tensor is complete in the caller's stream order
        B2_CUDA(cudaEventRecord(p->ev_misc, p->side));
        B2_CUDA(cudaStreamWaitEvent(st, p->ev_misc, 0));
    }
    return B2_OK;
}

template <typename T>
static int backward_t(b2_unet_plan* p, const float* const* prm, const float* const* dlogits, void* ws,
                      float* const* grads, int32_t* has_grad, int parts, cudaStream_t st,
                      const b2_grad_bucket* buckets = nullptr, int n_buckets = 0) {
    const b2_unet_geometry& g = p->g;
    const int P_ = g.num_pool;
    int rc;
    int next_bucket = 0;
    if (parts & B2_PART_DECODER)     // the decoder call opens a backward pass: every flag starts at 1
        for (size_t i = 0; i < p->params.size(); ++i) if (has_grad) has_grad[i] = 1;
    T* dzbuf[2] = {reinterpret_cast<T*>((char*)ws + p->off_grad) + p->dz_tmp.off,
                   reinterpret_cast<T*>((char*)ws + p->off_grad) + p->dz_tmp2.off};
    // weight-gradient kernels on the side stream (see b2_unet_plan): layer k uses dz buffer k & 1; the buffer is rewritten by
    // layer k + 2 only after the side stream has finished reading it
    const bool overlap = g_bwd_overlap && ensure_side_stream(p);
    cudaStream_t wst = overlap ? p->side : st;
    float* wscr = overlap ? SCR_WG(ws, p) : SCR(ws, p);
    // gradient buckets (data-parallel overlap): layers finish in descending parameter order, so the arena region
    // [first_param, end) of bucket k is complete once the layer owning `first_param` is enqueued -- record its events
    auto bucket_done = [&](int lowest_param) -> int {
        while (next_bucket < n_buckets && lowest_param <= buckets[next_bucket].first_param) {
            if (buckets[next_bucket].event_main) B2_CUDA(cudaEventRecord((cudaEvent_t)buckets[next_bucket].event_main, st));
            if (buckets[next_bucket].event_side) B2_CUDA(cudaEventRecord((cudaEvent_t)buckets[next_bucket].event_side, wst));
            ++next_bucket;
        }
        return B2_OK;
    };
    int layer_k = 0;
    bool wg_pending[2] = {false, false};
    bool skip_pending[8] = {false, false, false, false, false, false, false, false};
    auto conv_bwd = [&](ConvBlock& cb) -> int {
        const int buf = layer_k & 1;
        ++layer_k;
        T* dz = dzbuf[buf];
        if (overlap && wg_pending[buf]) B2_CUDA(cudaStreamWaitEvent(st, p->ev_wg[buf], 0));
        // the skip half of a concat gradient may still be in flight on the side stream: join before its first consumer
        // (conv_blocks_context.{l}.1 reads it as dy; conv_blocks_context.{l+1}.0 accumulates into it -- that one runs first)
        auto join_skip = [&](int lvl) -> int {
            if (lvl >= 0 && lvl < 8 && skip_pending[lvl]) {
                B2_CUDA(cudaStreamWaitEvent(st, p->ev_skip[lvl], 0));
                skip_pending[lvl] = false;
            }
            return B2_OK;
        };
        if (cb.skip_wait_level >= 0 && !cb.din_accumulate) { int rj = join_skip(cb.skip_wait_level); if (rj) return rj; }
        float* stats = F32(ws, p, cb.stats_off);
        int r = norm_lrelu_bwd<T>(P<T>(ws, p, cb.z, false), P<T>(ws, p, cb.y, false), P<T>(ws, p, cb.dy, true), stats, prm[cb.p_g],
                                  g_norm_recompute ? prm[cb.p_be] : nullptr, dz,
                                  grads[cb.p_g], grads[cb.p_be], g.batch, cb.z.vox(), cb.shape.cout, cb.z.pitch, cb.y.pitch,
                                  cb.dy.pitch, cb.shape.cout, g.lrelu_slope, SCR(ws, p), st);
        if (r) return r;
        if (overlap) {
            B2_CUDA(cudaEventRecord(p->ev_dz[buf], st));
            B2_CUDA(cudaStreamWaitEvent(wst, p->ev_dz[buf], 0));
        }
        ConvShape s = cb.shape;
        s.out_pitch = cb.shape.cout;  // dz is dense
        bool wdone = false;
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
            if (p->first_tc && &cb == &p->convs[0]) {
                if (!p->patch_valid) {   // forward ran the direct kernel: build the patch matrix now (side stream)
                    r = first_layer_patches(P<T>(ws, p, cb.in, false), g.batch, cb.in.d, cb.in.h, cb.in.w, cb.shape.cin, cb.in.pitch,
                                            P<T>(ws, p, p->patch, false), prm[cb.p_w], cb.shape.cout, nullptr, wst);
                    if (r) return r;
                    p->patch_valid = true;
                }
                r = first_layer_wgrad_tc(P<T>(ws, p, p->patch, false), dz, g.batch, cb.in.d, cb.in.h, cb.in.w, cb.shape.cin, cb.shape.cout,
                                         cb.shape.cout, wscr, grads[cb.p_w], grads[cb.p_b], wst);
                if (r) return r;
                wdone = true;
            }
            if (!wdone && cb.tc_wgrad) {
                r = conv3d_wgrad_tc(s, P<T>(ws, p, cb.in, false), dz, wscr, grads[cb.p_w], grads[cb.p_b], true, wst);
                if (r) return r;
                wdone = true;
            }
        }
        if (!wdone) {
            r = conv3d_wgrad_simt<T>(s, P<T>(ws, p, cb.in, false), dz, wscr, grads[cb.p_w], grads[cb.p_b], wst);
            if (r) return r;
        }
        if (cb.din.c > 0) {
            bool done = false;
            if (cb.din_accumulate && cb.skip_wait_level >= 0) { int rj = join_skip(cb.skip_wait_level); if (rj) return rj; }
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                if (cb.tc_dgrad && overlap && g_dgrad_split && cb.cat_level >= 0 && cb.cat_level < 8 && !cb.din_accumulate &&
                    cb.shape.cin % 64 == 0) {
                    // concat input [transposed-conv half | skip half]: the first half feeds the next link of the backward chain
                    // (main stream); the skip half is only read when the encoder backward reaches this level (side stream)
                    const int one[3] = {1, 1, 1};
                    const int half = cb.shape.cin / 2;
                    const __nv_bfloat16* wd = (const __nv_bfloat16*)F32(ws, p, cb.wd_off);
                    r = conv_tc_launch(dz, g.batch, cb.z.d, cb.z.h, cb.z.w, cb.shape.cout, cb.shape.cout, wd, half, nullptr,
                                       P<T>(ws, p, cb.din, true), cb.in.d, cb.in.h, cb.in.w, cb.din.pitch, one, 0, st, SCR(ws, p),
                                       p->scratch_floats * sizeof(float), nullptr, cb.shape.cin, 0);
                    if (r) return r;
                    r = conv_tc_launch(dz, g.batch, cb.z.d, cb.z.h, cb.z.w, cb.shape.cout, cb.shape.cout, wd, half, nullptr,
                                       P<T>(ws, p, cb.din, true) + half, cb.in.d, cb.in.h, cb.in.w, cb.din.pitch, one, 0, wst, wscr,
                                       p->wg_scratch_floats * sizeof(float), nullptr, cb.shape.cin, half);
                    if (r) return r;
                    B2_CUDA(cudaEventRecord(p->ev_skip[cb.cat_level], wst));
                    skip_pending[cb.cat_level] = true;
                    done = true;
                } else if (cb.tc_dgrad) {
                    const int one[3] = {1, 1, 1};
                    r = conv_tc_launch(dz, g.batch, cb.z.d, cb.z.h, cb.z.w, cb.shape.cout, cb.shape.cout,
                                       (const __nv_bfloat16*)F32(ws, p, cb.wd_off), cb.shape.cin, nullptr, P<T>(ws, p, cb.din, true),
                                       cb.in.d, cb.in.h, cb.in.w, cb.din.pitch, one, cb.din_accumulate, st, SCR(ws, p), p->scratch_floats * sizeof(float));
                    if (r) return r;
                    done = true;
                } else if (cb.tc_dgrad_strided) {
                    r = conv_tc_dgrad_strided(dz, g.batch, cb.z.d, cb.z.h, cb.z.w, cb.shape.cout, cb.shape.cout,
                                              (const __nv_bfloat16*)F32(ws, p, cb.wd_off), cb.shape.cin, P<T>(ws, p, cb.din, true),
                                              cb.in.d, cb.in.h, cb.in.w, cb.din.pitch, cb.shape.stride, cb.din_accumulate, st);
                    if (r) return r;
                    done = true;
                }
            }
            if (!done) {
                ConvShape sd = s;
                sd.in_pitch = cb.din.pitch;
                r = conv3d_dgrad_simt<T>(sd, dz, F32(ws, p, cb.wb_off), P<T>(ws, p, cb.din, true), cb.din_accumulate, st);
                if (r) return r;
            }
        }
        if (overlap) {   // everything the side stream reads from this layer's dz is enqueued
            B2_CUDA(cudaEventRecord(p->ev_wg[buf], wst));
            wg_pending[buf] = true;
        }
        return B2_OK;
    };
    // decoder, from full resolution (u = P-1) down to u = 0
    for (int u = (parts & B2_PART_DECODER) ? P_ - 1 : -1; u >= 0; --u) {
        Head& h = p->heads[u];
        const float* dl = dlogits[h.level];
        const bool first_writer = (u == P_ - 1);
        if (dl) {
            // data gradient on the main stream, weight gradient (+ its ordered reduce) on the side stream
            if ((rc = seghead_bwd<T>(P<T>(ws, p, h.in, false), prm[h.p_w], dl, P<T>(ws, p, h.din, true), first_writer ? 0 : 1, (float*)nullptr,
                                     g.batch, h.in.vox(), h.c, g.num_classes, h.in.pitch, h.din.pitch, SCR(ws, p), st))) return rc;
            if (overlap) {
                B2_CUDA(cudaEventRecord(p->ev_misc, st));       // dlogits were produced on st before this call
                B2_CUDA(cudaStreamWaitEvent(wst, p->ev_misc, 0));
            }
            if ((rc = seghead_bwd<T>(P<T>(ws, p, h.in, false), prm[h.p_w], dl, (T*)nullptr, 0, grads[h.p_w],
                                     g.batch, h.in.vox(), h.c, g.num_classes, h.in.pitch, h.din.pitch, wscr, wst))) return rc;
        } else {
            B2_CUDA(cudaMemsetAsync(grads[h.p_w], 0, p->params[h.p_w].numel * sizeof(float), st));
            if (has_grad) has_grad[h.p_w] = 0;
            if (first_writer) B2_CUDA(cudaMemsetAsync(P<T>(ws, p, h.din, true), 0, h.din.elems() * sizeof(T), st));
        }
        ConvBlock& l1 = p->convs[2 * (P_ + 1) + 2 * u + 1];
        ConvBlock& l0 = p->convs[2 * (P_ + 1) + 2 * u];
        if ((rc = bucket_done(h.p_w))) return rc;
        if ((rc = conv_bwd(l1))) return rc;
        if ((rc = bucket_done(l1.p_w))) return rc;
        if ((rc = conv_bwd(l0))) return rc;
        if ((rc = bucket_done(l0.p_w))) return rc;
        Tconv& t = p->tconvs[u];
        bool tdg = false;
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
            if (t.tc) {
                if ((rc = tconv_tc_dgrad(P<T>(ws, p, t.dout, true), g.batch, t.in.d, t.in.h, t.in.w, t.shape.cout, t.dout.pitch,
                                         (const __nv_bfloat16*)F32(ws, p, t.wqd_off), t.shape.cin, t.shape.k, P<T>(ws, p, t.din, true),
                                         t.din.pitch, st))) return rc;
                tdg = true;
            }
        }
        bool twg = false;
        if constexpr (std::is_same<T, __nv_bfloat16>::value) {
            if (t.tc && g_tc_wgrad && tconv_wgrad_tc_supported(t.shape.cin, t.shape.cout)) {
                if (overlap) {   // dout is complete (the localization convs' dgrad ran on st): fork
                    B2_CUDA(cudaEventRecord(p->ev_misc, st));
                    B2_CUDA(cudaStreamWaitEvent(wst, p->ev_misc, 0));
                }
                if ((rc = tconv_wgrad_tc(t.shape, P<T>(ws, p, t.in, false), P<T>(ws, p, t.dout, true), wscr, grads[t.p_w], wst))) return rc;
                twg = true;
            }
        }
        if (!(tdg && twg))
            if ((rc = tconv_bwd<T>(t.shape, P<T>(ws, p, t.in, false), P<T>(ws, p, t.dout, true), prm[t.p_w],
                                   tdg ? (T*)nullptr : P<T>(ws, p, t.din, true), twg ? (float*)nullptr : grads[t.p_w], SCR(ws, p), st))) return rc;
        if ((rc = bucket_done(t.p_w))) return rc;
    }
    for (int d = P_; d >= 0; --d) {
        const bool on = d < P_ ? (parts & B2_PART_ENCODER) != 0 : (parts & B2_PART_BOTTLENECK) != 0;
        if (!on) {
            // bottleneck bypassed (Generic_ViT_UNet V1 discards its output, generic_ViT_UNet.py:230-253): its parameters
            // receive no gradient (`param.grad is None`)
            if (d == P_ && (parts & B2_PART_ENCODER))
                for (int k = 0; k < 2; ++k) {
                    ConvBlock& cb = p->convs[2 * d + k];
                    for (int pi : {cb.p_w, cb.p_b, cb.p_g, cb.p_be}) {
                        B2_CUDA(cudaMemsetAsync(grads[pi], 0, p->params[pi].numel * sizeof(float), st));
                        if (has_grad) has_grad[pi] = 0;
                    }
                }
            if ((rc = bucket_done(p->convs[2 * d].p_w))) return rc;
            continue;
        }
        if ((rc = conv_bwd(p->convs[2 * d + 1]))) return rc;
        if ((rc = bucket_done(p->convs[2 * d + 1].p_w))) return rc;
        if ((rc = conv_bwd(p->convs[2 * d]))) return rc;
        if ((rc = bucket_done(p->convs[2 * d].p_w))) return rc;
    }
    if ((rc = bucket_done(0))) return rc;
    if (overlap) {   // join: every gradient is complete in the caller's stream order
        B2_CUDA(cudaEventRecord(p->ev_misc, wst));
        B2_CUDA(cudaStreamWaitEvent(st, p->ev_misc, 0));
    }
    return B2_OK;
}

}  // namespace b2

// ===============================================================================================================
// C ABI
// ===============================================================================================================
extern "C" int b2_version(void) { return 100; }
extern "C" int b2_set_option(const char* name, int value) {
    B2_CHECK_ARG(name);
    if (!strcmp(name, "tensor_cores")) { g_use_tc = value; return B2_OK; }
    if (!strcmp(name, "tc_strided")) { g_tc_strided = value; return B2_OK; }
    if (!strcmp(name, "tc_wgrad")) { g_tc_wgrad = value; return B2_OK; }
    if (!strcmp(name, "tc_halo")) { g_use_halo = value; return B2_OK; }
    if (!strcmp(name, "dgrad_mes")) { g_dgrad_mes = value; return B2_OK; }
    if (!strcmp(name, "dgrad_one_launch")) { g_dgrad_one_launch = value; return B2_OK; }
    if (!strcmp(name, "wgrad_halo")) { g_wgrad_halo = value; return B2_OK; }
    if (!strcmp(name, "wgrad_direct")) { g_wgrad_direct = value; return B2_OK; }
    if (!strcmp(name, "wgrad_dmerge")) { g_wgrad_dmerge = value; return B2_OK; }
    if (!strcmp(name, "halo_merge")) { g_halo_merge = value; return B2_OK; }
    if (!strcmp(name, "halo_dbg")) { g_halo_dbg = value; return B2_OK; }
    if (!strcmp(name, "halo_nsplit")) { g_halo_nsplit = value; return B2_OK; }
    if (!strcmp(name, "epi_stats")) { g_epi_stats = value; return B2_OK; }
    if (!strcmp(name, "pdl")) { g_pdl = value; return B2_OK; }
    if (!strcmp(name, "dgrad_split")) { g_dgrad_split = value; return B2_OK; }
    if (!strcmp(name, "bwd_overlap")) { g_bwd_overlap = value; return B2_OK; }
    if (!strcmp(name, "norm_cfg")) { g_norm_cfg = value; return B2_OK; }
    if (!strcmp(name, "norm_small")) { g_norm_small = value; return B2_OK; }
    if (!strcmp(name, "splitk_fuse")) { g_splitk_fuse = value; return B2_OK; }
    if (!strcmp(name, "norm_recompute")) { g_norm_recompute = value; return B2_OK; }
    if (!strcmp(name, "wgrad_desc_mode")) { g_wgrad_desc_mode = value; return B2_OK; }
    return fail(B2_EINVAL, "unknown option %s", name);
}
extern "C" const char* b2_last_error(void) { return g_last_error.c_str(); }
extern "C" long long b2_launch_count(void) { return g_launches.load(); }

extern "C" int b2_unet_plan_create(const b2_unet_geometry* geom, b2_unet_plan** out) {
    B2_CHECK_ARG(geom && out);
    B2_CHECK_ARG(geom->batch >= 1 && geom->in_channels >= 1 && geom->num_classes >= 2 && geom->num_classes <= 8);
    B2_CHECK_ARG(geom->num_pool >= 1 && geom->num_pool <= 7 && geom->base_features >= 1 && geom->max_features >= geom->base_features);
    B2_CHECK_ARG(geom->act_dtype == B2_F32 || geom->act_dtype == B2_BF16);
    for (int i = 0; i < geom->num_pool; ++i)
        for (int j = 0; j < 3; ++j) B2_CHECK_ARG(geom->pool[i][j] == 1 || geom->pool[i][j] == 2);
    b2_unet_plan* p = new (std::nothrow) b2_unet_plan();
    if (!p) return fail(B2_ENOMEM, "out of host memory%s", "");
    p->g = *geom;
    int rc = build_plan(p);
    if (rc) { delete p; return rc; }
    *out = p;
    return B2_OK;
}

extern "C" void b2_unet_plan_destroy(b2_unet_plan* plan) { delete plan; }

extern "C" int b2_unet_num_params(const b2_unet_plan* plan) { return plan ? (int)plan->params.size() : B2_EINVAL; }

extern "C" int b2_unet_param_info(const b2_unet_plan* plan, int idx, b2_param_info* out) {
    B2_CHECK_ARG(plan && out && idx >= 0 && idx < (int)plan->params.size());
    const ParamInfo& pi = plan->params[idx];
    memset(out, 0, sizeof(*out));
    strncpy(out->name, pi.name.c_str(), sizeof(out->name) - 1);
    out->ndim = (int)pi.shape.size();
    for (size_t i = 0; i < pi.shape.size(); ++i) out->shape[i] = pi.shape[i];
    out->numel = pi.numel;
    return B2_OK;
}

extern "C" size_t b2_unet_workspace_bytes(const b2_unet_plan* plan) { return plan ? plan->total_bytes : 0; }

extern "C" int b2_unet_output_shape(const b2_unet_plan* plan, int level, int32_t dhw[3]) {
    B2_CHECK_ARG(plan && dhw && level >= 0 && level < plan->g.num_pool);
    for (int i = 0; i < 3; ++i) dhw[i] = plan->sizes[level][i];
    return B2_OK;
}

extern "C" int b2_unet_forward(b2_unet_plan* plan, const float* const* params, const float* input, void* workspace,
                               float* const* logits, int keep_for_backward, b2_stream_t stream) {
    B2_CHECK_ARG(plan && params && input && workspace && logits);
    (void)keep_for_backward;
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->g.act_dtype == B2_F32) return forward_t<float>(plan, params, input, workspace, logits, B2_PART_ALL, st);
    return forward_t<__nv_bfloat16>(plan, params, input, workspace, logits, B2_PART_ALL, st);
}

extern "C" int b2_unet_forward_parts(b2_unet_plan* plan, const float* const* params, const float* input, void* workspace,
                                     float* const* logits, int parts, b2_stream_t stream) {
    B2_CHECK_ARG(plan && params && workspace && parts > 0 && parts <= B2_PART_ALL);
    B2_CHECK_ARG(!(parts & B2_PART_ENCODER) || input);
    B2_CHECK_ARG(!(parts & B2_PART_DECODER) || logits);
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->g.act_dtype == B2_F32) return forward_t<float>(plan, params, input, workspace, logits, parts, st);
    return forward_t<__nv_bfloat16>(plan, params, input, workspace, logits, parts, st);
}

extern "C" int b2_unet_backward(b2_unet_plan* plan, const float* const* params, const float* const* dlogits,
                                void* workspace, float* const* grads, int32_t* has_grad_host, b2_stream_t stream) {
    B2_CHECK_ARG(plan && params && dlogits && workspace && grads);
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->g.act_dtype == B2_F32) return backward_t<float>(plan, params, dlogits, workspace, grads, has_grad_host, B2_PART_ALL, st);
    return backward_t<__nv_bfloat16>(plan, params, dlogits, workspace, grads, has_grad_host, B2_PART_ALL, st);
}

// backward with gradient buckets: events are recorded as soon as the arena suffix [first_param, end) is complete, so the
// caller can all-reduce bucket k on a communication stream while the remaining layers are still being differentiated
extern "C" int b2_unet_backward_buckets(b2_unet_plan* plan, const float* const* params, const float* const* dlogits,
                                        void* workspace, float* const* grads, int32_t* has_grad_host,
                                        const b2_grad_bucket* buckets_host, int n_buckets, b2_stream_t stream) {
    B2_CHECK_ARG(plan && params && dlogits && workspace && grads && (n_buckets == 0 || buckets_host));
    for (int i = 1; i < n_buckets; ++i) B2_CHECK_ARG(buckets_host[i].first_param < buckets_host[i - 1].first_param);
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->g.act_dtype == B2_F32)
        return backward_t<float>(plan, params, dlogits, workspace, grads, has_grad_host, B2_PART_ALL, st, buckets_host, n_buckets);
    return backward_t<__nv_bfloat16>(plan, params, dlogits, workspace, grads, has_grad_host, B2_PART_ALL, st, buckets_host, n_buckets);
}

extern "C" int b2_unet_backward_parts(b2_unet_plan* plan, const float* const* params, const float* const* dlogits,
                                      void* workspace, float* const* grads, int32_t* has_grad_host, int parts,
                                      b2_stream_t stream) {
    B2_CHECK_ARG(plan && params && workspace && grads && parts > 0 && parts <= B2_PART_ALL);
    B2_CHECK_ARG(!(parts & B2_PART_DECODER) || dlogits);
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->g.act_dtype == B2_F32) return backward_t<float>(plan, params, dlogits, workspace, grads, has_grad_host, parts, st);
    return backward_t<__nv_bfloat16>(plan, params, dlogits, workspace, grads, has_grad_host, parts, st);
}

// LwF (reference lwf:315-346): every stored task head differs from the running model only in the 1x1x1 `seg_outputs`
// convolutions, so an old head's prediction is that head applied to the decoder activation the last forward left in the
// workspace -- one 1x1x1 launch instead of one full network forward per head.
extern "C" int b2_unet_head_forward(b2_unet_plan* plan, void* workspace, int level, const float* weight, float* logits,
                                    b2_stream_t stream) {
    B2_CHECK_ARG(plan && workspace && weight && logits && level >= 0 && level < plan->g.num_pool);
    cudaStream_t st = (cudaStream_t)stream;
    for (const Head& h : plan->heads) {
        if (h.level != level) continue;
        if (plan->g.act_dtype == B2_F32)
            return seghead_fwd<float>(P<float>(workspace, plan, h.in, false), weight, logits, plan->g.batch, h.in.vox(), h.c,
                                      plan->g.num_classes, h.in.pitch, st);
        return seghead_fwd<__nv_bfloat16>(P<__nv_bfloat16>(workspace, plan, h.in, false), weight, logits, plan->g.batch, h.in.vox(),
                                          h.c, plan->g.num_classes, h.in.pitch, st);
    }
    return fail(B2_EINVAL, "no head at level %lld%s", "", level);
}

extern "C" int b2_unet_num_convs(const b2_unet_plan* plan) { return plan ? (int)plan->conv_modules.size() : B2_EINVAL; }

extern "C" int b2_unet_conv_name(const b2_unet_plan* plan, int conv_idx, char name[96]) {
    B2_CHECK_ARG(plan && name && conv_idx >= 0 && conv_idx < (int)plan->conv_modules.size());
    auto km = plan->conv_modules[conv_idx];
    std::string s = km.first == 0 ? plan->convs[km.second].prefix + ".conv" : plan->tconvs[km.second].prefix;
    memset(name, 0, 96);
    strncpy(name, s.c_str(), 95);
    return B2_OK;
}

extern "C" int b2_unet_debug_view(const b2_unet_plan* plan, void* workspace, int block_idx, int which, b2_act_view* out) {
    B2_CHECK_ARG(plan && workspace && out && block_idx >= 0 && block_idx < (int)plan->convs.size() && which >= 0 && which <= 4);
    const ConvBlock& cb = plan->convs[block_idx];
    const Act& a = which == 0 ? cb.z : which == 1 ? cb.y : which == 2 ? cb.dy : which == 3 ? cb.in : cb.din;
    const bool grad = which == 2 || which == 4;
    out->ptr = (char*)workspace + (grad ? plan->off_grad : plan->off_act) + a.off * plan->esz;
    out->n = a.n; out->d = a.d; out->h = a.h; out->w = a.w; out->c = a.c; out->pitch = a.pitch;
    out->dtype = plan->g.act_dtype;
    return B2_OK;
}

extern "C" int b2_unet_conv_output(const b2_unet_plan* plan, void* workspace, int conv_idx, b2_act_view* out) {
    B2_CHECK_ARG(plan && workspace && out && conv_idx >= 0 && conv_idx < (int)plan->conv_modules.size());
    auto km = plan->conv_modules[conv_idx];
    const Act& a = km.first == 0 ? plan->convs[km.second].z : plan->tconvs[km.second].out;
    out->ptr = (char*)workspace + plan->off_act + a.off * plan->esz;
    out->n = a.n; out->d = a.d; out->h = a.h; out->w = a.w; out->c = a.c; out->pitch = a.pitch;
    out->dtype = plan->g.act_dtype;
    return B2_OK;
}

// ---- building blocks (tests) ------------------------------------------------------------------------------------
static ConvShape to_shape(const b2_conv_desc* d) {
    ConvShape s;
    s.n = d->n; s.d = d->d; s.h = d->h; s.w = d->w; s.cin = d->cin; s.cout = d->cout;
    for (int i = 0; i < 3; ++i) s.stride[i] = d->stride[i];
    s.in_pitch = d->in_pitch; s.out_pitch = d->out_pitch;
    return s;
}

extern "C" size_t b2_conv3d_scratch_bytes(const b2_conv_desc* d) {
    if (!d) return 0;
    ConvShape s = to_shape(d);
    size_t w = (size_t)27 * s.cin * s.cout;
    size_t part = conv_stat_part_floats(s);
    size_t wg = conv_wgrad_part_floats(s);
    if (wgrad_tc_supported(s.cin, s.cout)) { size_t t = wgrad_tc_part_floats(s); if (t > wg) wg = t; }
    size_t st2 = instnorm_stats_scratch_floats(s.n, (long long)s.d * s.h * s.w, s.cout);
    if (st2 > part) part = st2;
    size_t sk = conv_tc_splitk_scratch_floats(s.n, s.d, s.h, s.w, s.cout > s.cin ? s.cout : s.cin);
    if (sk > part) part = sk;
    size_t es = (size_t)s.n * num_sms() * 8 * s.cout * 2;   // InstanceNorm partials from the convolution epilogue
    if (es > part) part = es;
    return align_up((2 * w + (part > wg ? part : wg) + 64) * sizeof(float));
}

extern "C" int b2_conv3d_fwd(const b2_conv_desc* d, const void* x, const float* w_pt, const float* bias, void* z,
                             float* stats, float eps, void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(d && x && w_pt && z && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    ConvShape s = to_shape(d);
    float* wf = (float*)scratch;
    float* wb = wf + (size_t)27 * s.cin * s.cout;
    float* part = wb + (size_t)27 * s.cin * s.cout;
    int rc = weight_shadow(w_pt, s.cout, s.cin, wf, wb, st);
    if (rc) return rc;
    if (d->dtype == B2_F32) return conv3d_fwd_simt<float>(s, (const float*)x, wf, bias, (float*)z, part, stats, eps, st);
    const bool strided = s.stride[0] != 1 || s.stride[1] != 1 || s.stride[2] != 1;
    if (g_use_tc && conv_tc_supported(s.cin, s.cout) && (!strided || g_tc_strided) && s.in_pitch % 8 == 0 && s.out_pitch % 8 == 0) {
        // bf16 shadow aliases the (unused) fp32 forward shadow
        __nv_bfloat16* wk = (__nv_bfloat16*)wf;
        rc = weight_shadow_bf16(w_pt, s.cout, s.cin, wk, nullptr, st);
        if (rc) return rc;
        const int od = (s.d - 1) / s.stride[0] + 1, oh = (s.h - 1) / s.stride[1] + 1, ow = (s.w - 1) / s.stride[2] + 1;
        const size_t part_bytes = b2_conv3d_scratch_bytes(d) - 2 * (size_t)27 * s.cin * s.cout * sizeof(float) - 256;
        int stat_slots = 0;
        rc = conv_tc_launch((const __nv_bfloat16*)x, s.n, s.d, s.h, s.w, s.cin, s.in_pitch, wk, s.cout, bias, (__nv_bfloat16*)z, od, oh, ow,
                            s.out_pitch, s.stride, 0, st, part, part_bytes, stats ? &stat_slots : nullptr);
        if (rc) return rc;
        if (stats && stat_slots > 0) return stats_finalize(part, stat_slots, s.n, (long long)od * oh * ow, s.cout, eps, stats, st);
        if (stats) return instnorm_stats<__nv_bfloat16>((const __nv_bfloat16*)z, s.n, (long long)od * oh * ow, s.cout, s.out_pitch, part, stats, eps, st);
        return B2_OK;
    }
    return conv3d_fwd_simt<__nv_bfloat16>(s, (const __nv_bfloat16*)x, wf, bias, (__nv_bfloat16*)z, part, stats, eps, st);
}

// prepared-weights variant (bf16 tensor-core path): make the [27][Cout][Cin] bf16 shadow once, run the conv many times
extern "C" size_t b2_conv3d_shadow_bytes(const b2_conv_desc* d) { return d ? align_up((size_t)27 * d->cin * d->cout * 2) : 0; }

extern "C" int b2_conv3d_make_shadow(const b2_conv_desc* d, const float* w_pt, void* shadow, b2_stream_t stream) {
    B2_CHECK_ARG(d && w_pt && shadow && d->dtype == B2_BF16);
    return weight_shadow_bf16(w_pt, d->cout, d->cin, (__nv_bfloat16*)shadow, nullptr, (cudaStream_t)stream);
}

static int conv3d_fwd_shadow_impl(const b2_conv_desc* d, const void* x, const void* shadow, const float* bias, void* z, void* scratch,
                                  int epilogue_stats, cudaStream_t st) {
    ConvShape s = to_shape(d);
    if (!conv_tc_supported(s.cin, s.cout) || s.in_pitch % 8 != 0 || s.out_pitch % 8 != 0)
        return fail(B2_EUNSUPPORTED, "b2_conv3d_fwd_shadow: shape not covered by the tensor-core path%s", "");
    const int od = (s.d - 1) / s.stride[0] + 1, oh = (s.h - 1) / s.stride[1] + 1, ow = (s.w - 1) / s.stride[2] + 1;
    const size_t part_bytes = scratch ? b2_conv3d_scratch_bytes(d) - 2 * (size_t)27 * s.cin * s.cout * sizeof(float) - 256 : 0;
    float* part = scratch ? (float*)scratch + 2 * (size_t)27 * s.cin * s.cout : nullptr;
    int stat_slots = 0;
    return conv_tc_launch((const __nv_bfloat16*)x, s.n, s.d, s.h, s.w, s.cin, s.in_pitch, (const __nv_bfloat16*)shadow, s.cout, bias,
                          (__nv_bfloat16*)z, od, oh, ow, s.out_pitch, s.stride, 0, st, part, part_bytes,
                          epilogue_stats && part ? &stat_slots : nullptr);
}

extern "C" int b2_conv3d_fwd_shadow(const b2_conv_desc* d, const void* x, const void* shadow, const float* bias, void* z, void* scratch,
                                    b2_stream_t stream) {
    B2_CHECK_ARG(d && x && shadow && z && d->dtype == B2_BF16);
    return conv3d_fwd_shadow_impl(d, x, shadow, bias, z, scratch, 0, (cudaStream_t)stream);
}

// the variant the training step runs: the same single launch, with the InstanceNorm partial sums (sum z, sum z^2 per
// (n, c) and epilogue warp) produced by the convolution's epilogue into `scratch`
extern "C" int b2_conv3d_fwd_shadow_stats(const b2_conv_desc* d, const void* x, const void* shadow, const float* bias, void* z,
                                          void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(d && x && shadow && z && scratch && d->dtype == B2_BF16);
    return conv3d_fwd_shadow_impl(d, x, shadow, bias, z, scratch, 1, (cudaStream_t)stream);
}

static TconvShape to_tshape(const b2_tconv_desc* d) {
    TconvShape s;
    s.n = d->n; s.d = d->d; s.h = d->h; s.w = d->w; s.cin = d->cin; s.cout = d->cout;
    for (int a = 0; a < 3; ++a) s.k[a] = d->k[a];
    s.in_pitch = d->in_pitch; s.out_pitch = d->out_pitch;
    return s;
}
static bool tconv_desc_ok(const b2_tconv_desc* d) {
    if (!d || d->n < 1 || d->d < 1 || d->h < 1 || d->w < 1 || d->cin < 1 || d->cout < 1 || d->in_pitch < d->cin || d->out_pitch < d->cout)
        return false;
    for (int a = 0; a < 3; ++a)
        if (d->k[a] != 1 && d->k[a] != 2) return false;
    return d->dtype == B2_F32 || d->dtype == B2_BF16;
}

extern "C" size_t b2_tconv3d_scratch_bytes(const b2_tconv_desc* d) {
    if (!tconv_desc_ok(d)) return 0;
    const TconvShape s = to_tshape(d);
    const size_t k8 = (size_t)s.k[0] * s.k[1] * s.k[2];
    return align_up(k8 * s.cin * s.cout * sizeof(float)) + align_up(tconv_bwd_scratch_floats(s) * sizeof(float)) + 256;
}

extern "C" int b2_tconv3d_fwd(const b2_tconv_desc* d, const void* x, const float* w_pt, void* y, void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(tconv_desc_ok(d) && x && w_pt && y && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    const TconvShape s = to_tshape(d);
    float* wq = (float*)scratch;
    int rc = tconv_shadow(w_pt, s.cin, s.cout, s.k[0] * s.k[1] * s.k[2], wq, st);
    if (rc) return rc;
    if (d->dtype == B2_F32) return tconv_fwd_q<float>(s, (const float*)x, wq, (float*)y, st);
    return tconv_fwd_q<__nv_bfloat16>(s, (const __nv_bfloat16*)x, wq, (__nv_bfloat16*)y, st);
}

extern "C" int b2_tconv3d_bwd(const b2_tconv_desc* d, const void* x, const void* dy, const float* w_pt, void* dx, float* dw, void* scratch,
                              b2_stream_t stream) {
    B2_CHECK_ARG(tconv_desc_ok(d) && x && dy && w_pt && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    const TconvShape s = to_tshape(d);
    const size_t k8 = (size_t)s.k[0] * s.k[1] * s.k[2];
    float* part = (float*)((char*)scratch + align_up(k8 * s.cin * s.cout * sizeof(float)));
    if (d->dtype == B2_F32) return tconv_bwd<float>(s, (const float*)x, (const float*)dy, w_pt, (float*)dx, dw, part, st);
    return tconv_bwd<__nv_bfloat16>(s, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, w_pt, (__nv_bfloat16*)dx, dw, part, st);
}

extern "C" int b2_conv3d_bwd(const b2_conv_desc* d, const void* x, const void* dz, const float* w_pt, void* dx,
                             int accumulate_dx, float* dw, float* dbias, void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(d && x && dz && w_pt && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    ConvShape s = to_shape(d);
    float* wf = (float*)scratch;
    float* wb = wf + (size_t)27 * s.cin * s.cout;
    float* part = wb + (size_t)27 * s.cin * s.cout;
    int rc = weight_shadow(w_pt, s.cout, s.cin, wf, wb, st);
    if (rc) return rc;
    if (d->dtype == B2_F32) {
        if (dw && (rc = conv3d_wgrad_simt<float>(s, (const float*)x, (const float*)dz, part, dw, dbias, st))) return rc;
        if (dx && (rc = conv3d_dgrad_simt<float>(s, (const float*)dz, wb, (float*)dx, accumulate_dx, st))) return rc;
    } else {
        const bool strided_w = s.stride[0] != 1 || s.stride[1] != 1 || s.stride[2] != 1;
        if (dw && g_use_tc && g_tc_wgrad && wgrad_tc_supported(s.cin, s.cout) && (!strided_w || g_tc_strided) && s.in_pitch % 8 == 0 && s.out_pitch % 8 == 0) {
            if ((rc = conv3d_wgrad_tc(s, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, part, dw, dbias, false, st))) return rc;
        } else if (dw && (rc = conv3d_wgrad_simt<__nv_bfloat16>(s, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, part, dw, dbias, st))) return rc;
        const bool strided = s.stride[0] != 1 || s.stride[1] != 1 || s.stride[2] != 1;
        if (dx && g_use_tc && !strided && conv_tc_supported(s.cout, s.cin) && s.in_pitch % 8 == 0 && s.out_pitch % 8 == 0) {
            __nv_bfloat16* wd = (__nv_bfloat16*)wf;   // forward fp32 shadow is unused from here on
            if ((rc = weight_shadow_bf16(w_pt, s.cout, s.cin, nullptr, wd, st))) return rc;
            const int one[3] = {1, 1, 1};
            const size_t part_bytes = b2_conv3d_scratch_bytes(d) - 2 * (size_t)27 * s.cin * s.cout * sizeof(float) - 256;
            if ((rc = conv_tc_launch((const __nv_bfloat16*)dz, s.n, s.d, s.h, s.w, s.cout, s.out_pitch, wd, s.cin, nullptr,
                                     (__nv_bfloat16*)dx, s.d, s.h, s.w, s.in_pitch, one, accumulate_dx, st, part, part_bytes))) return rc;
        } else if (dx && g_use_tc && strided && g_tc_strided && conv_tc_supported(s.cout, s.cin) && s.in_pitch % 8 == 0 && s.out_pitch % 8 == 0) {
            __nv_bfloat16* wd = (__nv_bfloat16*)wf;
            if ((rc = weight_shadow_bf16(w_pt, s.cout, s.cin, nullptr, wd, st))) return rc;
            const int od = (s.d - 1) / s.stride[0] + 1, oh = (s.h - 1) / s.stride[1] + 1, ow = (s.w - 1) / s.stride[2] + 1;
            if ((rc = conv_tc_dgrad_strided((const __nv_bfloat16*)dz, s.n, od, oh, ow, s.cout, s.out_pitch, wd, s.cin, (__nv_bfloat16*)dx,
                                            s.d, s.h, s.w, s.in_pitch, s.stride, accumulate_dx, st))) return rc;
        } else if (dx && (rc = conv3d_dgrad_simt<__nv_bfloat16>(s, (const __nv_bfloat16*)dz, wb, (__nv_bfloat16*)dx, accumulate_dx, st))) return rc;
    }
    return B2_OK;
}

extern "C" size_t b2_norm_scratch_bytes(int n, int64_t vox, int c) { return align_up(norm_bwd_scratch_floats(n, vox, c) * sizeof(float)); }

extern "C" int b2_norm_lrelu_fwd(const void* z, const float* stats, const float* gamma, const float* beta, void* y, int n,
                                 int64_t vox, int c, int z_pitch, int y_pitch, int dtype, float slope, b2_stream_t stream) {
    B2_CHECK_ARG(z && stats && gamma && beta && y);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == B2_F32) return norm_lrelu_fwd<float>((const float*)z, stats, gamma, beta, (float*)y, n, vox, c, z_pitch, y_pitch, slope, st);
    return norm_lrelu_fwd<__nv_bfloat16>((const __nv_bfloat16*)z, stats, gamma, beta, (__nv_bfloat16*)y, n, vox, c, z_pitch, y_pitch, slope, st);
}

extern "C" int b2_norm_lrelu_bwd(const void* z, const void* y, const void* dy, const float* stats, const float* gamma,
                                 void* dz, float* dgamma, float* dbeta, int n, int64_t vox, int c, int z_pitch, int y_pitch,
                                 int dy_pitch, int dz_pitch, int dtype, float slope, void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(z && y && dy && stats && gamma && dz && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == B2_F32)
        return norm_lrelu_bwd<float>((const float*)z, (const float*)y, (const float*)dy, stats, gamma, nullptr, (float*)dz, dgamma, dbeta, n, vox, c,
                                     z_pitch, y_pitch, dy_pitch, dz_pitch, slope, (float*)scratch, st);
    return norm_lrelu_bwd<__nv_bfloat16>((const __nv_bfloat16*)z, (const __nv_bfloat16*)y, (const __nv_bfloat16*)dy, stats, gamma,
                                         nullptr, (__nv_bfloat16*)dz, dgamma, dbeta, n, vox, c, z_pitch, y_pitch, dy_pitch, dz_pitch, slope,
                                         (float*)scratch, st);
}
