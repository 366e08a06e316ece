// kernels.h -- internal (C++) launcher declarations shared by the translation units of libb2unet.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b2 {

struct TconvShape {
    int n, d, h, w;      // input spatial extent
    int cin, cout;
    int k[3];            // kernel == stride (1 or 2 per axis)
    int in_pitch, out_pitch;
};

struct ConvShape {
    int n, d, h, w;      // input spatial extent
    int cin, cout;
    int stride[3];
    int in_pitch, out_pitch;
};

// ---- conv3d_simt.cu -------------------------------------------------------------------------------------------
size_t conv_stat_part_floats(const ConvShape& s);
size_t conv_wgrad_part_floats(const ConvShape& s);
int weight_shadow(const float* w_pt, int cout, int cin, float* wf, float* wb, cudaStream_t st);
template <typename T>
int conv3d_fwd_simt(const ConvShape& s, const T* x, const float* wf, const float* bias, T* z, float* stat_part,
                    float* stats, float eps, cudaStream_t st);
template <typename T>
int conv3d_dgrad_simt(const ConvShape& s, const T* dz, const float* wb, T* dx, int accumulate, cudaStream_t st);
template <typename T>
int conv3d_wgrad_simt(const ConvShape& s, const T* x, const T* dz, float* part, float* dw, float* dbias,
                      cudaStream_t st);

// per-(n,c) {mean, rstd} of an NDHWC tensor (used when the conv epilogue does not produce the partial sums)
size_t instnorm_stats_scratch_floats(int n, long long vox, int c);
template <typename T>
int instnorm_stats(const T* z, int n, long long vox, int c, int pitch, float* part, float* stats, float eps, cudaStream_t st);

// ---- conv3d_tc.cu (tcgen05 / TMA path, bf16) ----------------------------------------------------------------------
bool conv_tc_supported(int K, int Nout);
int weight_shadow_bf16(const float* w_pt, int cout, int cin, __nv_bfloat16* wk, __nv_bfloat16* wd, cudaStream_t st);
// multi-tensor bf16 shadow: w[A][B][T] fp32 -> oab[t][A][B], oba[flip ? T-1-t : t][B][A] (either may be null)
struct ShadowJob { const float* w; __nv_bfloat16* oab; __nv_bfloat16* oba; int A, B, T, flip; };
bool shadow_job_supported(int A, int B, int T);
int shadow_multi(const ShadowJob* jobs, int n, cudaStream_t st);
// split-K second stage left to the consumer: in the deep stages the small-tensor norm kernel sums the fp32 partials itself
// (one launch and one round trip of z less per layer).  Filled by conv_tc_gather when it skipped splitk_reduce_kernel.
struct SplitKDefer {
    int deferred;
    int TN, TD, TH, TW, nt_d, nt_h, nt_w, nblk, BN, ksplit, otiles;
    const float* partial;        // [ksplit][otiles][128][BN]
};

struct TcGather {
    const __nv_bfloat16* src; int N, Ds, Hs, Ws, K, src_pitch;     // gathered tensor (NDHWC) and its channel count (GEMM K)
    const __nv_bfloat16* wmat; int w_rows, rows_per_tap, Nout;      // weight matrix [w_rows][K]; GEMM N
    int w_row0;                                                     // first row of the computed window inside a tap block
    const float* bias;
    __nv_bfloat16* dst; int Dd, Hd, Wd, dst_pitch, accumulate;      // produced tensor
    int LD, LH, LW;                                                 // logical grid tiled by the 128-voxel boxes
    int stride[3];                                                  // source = logical * stride + tap_off
    int os[3], oo[3];                                               // produced = logical * os + oo
    int ntaps; int tap_off[27][3]; int tap_w[27];
    int q_scatter, q_channels, qk[3];                               // transposed-conv forward column->voxel scatter
    float* splitk_scratch; size_t splitk_scratch_bytes;             // optional fp32 scratch enabling split-K for small volumes
    // tile classes (strided dgrad: one class per output-parity lattice, one launch): class c uses the taps
    // [cls_tap0[c], +cls_ntaps[c]) of the tables above, lattice offset cls_oo[c] and logical extent cls_L[c]
    int nclass; int cls_tap0[8], cls_ntaps[8], cls_oo[8][3], cls_L[8][3];
    // shared-A stages (merged-class strided dgrad, see TcConvParams in conv3d_tc.cu): stage s = shift tap_off[s] + blocks
    // [mes_blk0[s], +mes_nb[s]) of (weight row-block, class column block, first-contribution flag)
    int mes, mes_nst, mes_blk_rows, mes_nb[12], mes_blk0[13], mes_wrow[27], mes_cls[27], mes_first[27];
    // optional InstanceNorm partials from the epilogue: part[N][*stat_slots][Nout][2] (see EpiStats in tc_common.cuh)
    float* stat_part; size_t stat_part_floats; int* stat_slots;
    int out_f32;                                                    // produced tensor is fp32 (GEMM use)
    SplitKDefer* defer;                                             // optional: leave a split-K reduction to the consumer
};
int conv_tc_gather(const TcGather& g, cudaStream_t st);
int gemm_tn_bf16(const __nv_bfloat16* A, int M, int K, int lda, const __nv_bfloat16* W, int N, const float* bias, void* out, int ldo,
                 int out_f32, float* scratch, size_t scratch_bytes, cudaStream_t st);
int conv_tc_dgrad_strided(const __nv_bfloat16* dz, int N, int Do, int Ho, int Wo, int Cout, int dz_pitch, const __nv_bfloat16* wd,
                          int Cin, __nv_bfloat16* dx, int Di, int Hi, int Wi, int dx_pitch, const int stride[3], int accumulate,
                          cudaStream_t st);
int tconv_tc_fwd(const __nv_bfloat16* x, int N, int D, int H, int W, int Cin, int x_pitch, const __nv_bfloat16* wq, int Cout,
                 const int k[3], __nv_bfloat16* y, int y_pitch, cudaStream_t st);
int tconv_tc_dgrad(const __nv_bfloat16* dy, int N, int D, int H, int W, int Cout, int dy_pitch, const __nv_bfloat16* wqd, int Cin,
                   const int k[3], __nv_bfloat16* dx, int dx_pitch, cudaStream_t st);
int tconv_shadow_bf16(const float* w_pt, int cin, int cout, int k8, __nv_bfloat16* wq, __nv_bfloat16* wqd, cudaStream_t st);
int conv_tc_launch(const __nv_bfloat16* src, int N, int Ds, int Hs, int Ws, int K, int src_pitch, const __nv_bfloat16* wmat,
                   int Nout, const float* bias, __nv_bfloat16* dst, int Dd, int Hd, int Wd, int dst_pitch, const int stride[3],
                   int accumulate, cudaStream_t st, float* scratch = nullptr, size_t scratch_bytes = 0, int* stat_slots = nullptr,
                   int w_pitch = 0, int w_row0 = 0, SplitKDefer* defer = nullptr);
// mean / rstd from epilogue partials part[n][slots][c][2]
int stats_finalize(const float* part, int slots, int n, long long vox, int c, float eps, float* stats, cudaStream_t st);
size_t conv_tc_splitk_scratch_floats(int N, int D, int H, int W, int Nout);
bool wgrad_tc_supported(int cin, int cout);
size_t wgrad_tc_part_floats(const ConvShape& s);
int conv3d_wgrad_tc(const ConvShape& s, const __nv_bfloat16* x, const __nv_bfloat16* dz, float* part, float* dw, float* dbias,
                    bool bias_feeds_norm, cudaStream_t st);
bool tconv_wgrad_tc_supported(int cin, int cout);
size_t tconv_wgrad_tc_part_floats(const TconvShape& s);
int tconv_wgrad_tc(const TconvShape& s, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* part, float* dw, cudaStream_t st);
bool conv_tc_halo_supported(int K, int Nout, int N, int D, int H, int W);
int conv_tc_halo_launch(const __nv_bfloat16* src, int N, int D, int H, int W, int K, int src_pitch, const __nv_bfloat16* wmat, int Nout,
                        const float* bias, __nv_bfloat16* dst, int dst_pitch, int accumulate, cudaStream_t st, float* stat_part = nullptr,
                        size_t stat_part_floats = 0, int* stat_slots = nullptr, int w_pitch = 0, int w_row0 = 0);
bool first_layer_tc_supported(int cin, int cout);
int first_layer_patches(const __nv_bfloat16* x, int N, int D, int H, int W, int cin, int x_pitch, __nv_bfloat16* P, const float* w_pt,
                        int cout, __nv_bfloat16* wp, cudaStream_t st);
size_t first_layer_wgrad_part_floats(int N, int D, int H, int W, int cout);
int first_layer_wgrad_tc(const __nv_bfloat16* P, const __nv_bfloat16* dz, int N, int D, int H, int W, int cin, int cout, int dz_pitch,
                         float* part, float* dw, float* dbias, cudaStream_t st);
extern int g_halo_dbg;
extern int g_use_halo, g_halo_merge, g_halo_nsplit, g_dgrad_one_launch, g_dgrad_mes, g_epi_stats;
extern int g_wgrad_desc_mode, g_tc_wgrad, g_wgrad_dmerge, g_wgrad_direct, g_wgrad_halo;
extern int g_use_tc;   // 1: tensor-core path for bf16 plans where supported (default), 0: SIMT only

// ---- norm.cu ----------------------------------------------------------------------------------------------------
// y = lrelu(gamma*(z-mean)*rstd + beta)
template <typename T>
int norm_lrelu_fwd(const T* z, const float* stats, const float* gamma, const float* beta, T* y, int n, long long vox,
                   int c, int z_pitch, int y_pitch, float slope, cudaStream_t st);
// small tensors (deep stages): statistics + apply in ONE launch; also writes stats [n][c]{mean, rstd}
bool norm_small_supported(long long vox, int c, int p0, int p1, int p2, int p3);
template <typename T>
int norm_lrelu_fwd_small(const T* z, const float* gamma, const float* beta, T* y, float* stats, int n, long long vox, int c,
                         int z_pitch, int y_pitch, float slope, float eps, cudaStream_t st);
// the same on un-reduced split-K partials of the producing convolution: z = bf16(bias + sum_ks partial) is written on the way
int splitk_norm_small_fwd(const SplitKDefer& k, const float* bias, __nv_bfloat16* z, const float* gamma, const float* beta,
                          __nv_bfloat16* y, float* stats, int n, int D, int H, int W, int c, int z_pitch, int y_pitch, float slope,
                          float eps, cudaStream_t st);
extern int g_norm_small, g_norm_cfg, g_splitk_fuse;
size_t norm_bwd_scratch_floats(int n, long long vox, int c);
// dz = d(loss)/dz given dy; dgamma/dbeta overwritten. scratch: norm_bwd_scratch_floats floats.
template <typename T>
int norm_lrelu_bwd(const T* z, const T* y, const T* dy, const float* stats, const float* gamma, const float* beta, T* dz,
                   float* dgamma, float* dbeta, int n, long long vox, int c, int z_pitch, int y_pitch, int dy_pitch, int dz_pitch,
                   float slope, float* scratch, cudaStream_t st);

// ---- updown.cu --------------------------------------------------------------------------------------------------
// w_pt: PyTorch ConvTranspose3d weight [Cin][Cout][kd][kh][kw]; wq: shadow [K8][Cin][Cout]
int tconv_shadow(const float* w_pt, int cin, int cout, int k8, float* wq, cudaStream_t st);
template <typename T>
int tconv_fwd_q(const TconvShape& s, const T* x, const float* wq, T* y, cudaStream_t st);
size_t tconv_bwd_scratch_floats(const TconvShape& s);
template <typename T>
int tconv_bwd(const TconvShape& s, const T* x, const T* dy, const float* w_pt, T* dx, float* dw, float* scratch,
              cudaStream_t st);
// 1x1x1 head: logits NCDHW fp32 [n][ncls][vox] = y[n][vox][:] . w[ncls][c]
template <typename T>
int seghead_fwd(const T* y, const float* w, float* logits, int n, long long vox, int c, int ncls, int y_pitch,
                cudaStream_t st);
size_t seghead_bwd_scratch_floats(int n, long long vox, int c, int ncls);
// dy (+)= w^T dlogits ; dw overwritten
template <typename T>
int seghead_bwd(const T* y, const float* w, const float* dlogits, T* dy, int accumulate, float* dw, int n,
                long long vox, int c, int ncls, int y_pitch, int dy_pitch, float* scratch, cudaStream_t st);
// NCDHW fp32 -> NDHWC T
template <typename T>
int nchw_to_ndhwc(const float* src, T* dst, int n, int c, long long vox, int dst_pitch, cudaStream_t st);

int num_sms();

}  // namespace b2
