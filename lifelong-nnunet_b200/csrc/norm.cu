// norm.cu -- InstanceNorm3d(affine, eps) + LeakyReLU, forward apply and backward, on NDHWC tensors.
// Replaces nn.InstanceNorm3d + nn.LeakyReLU of nnunet's ConvDropoutNormNonlin (Appendix A).
// The statistics themselves come out of the convolution epilogue (conv3d_*.cu); here:
//   fwd : y  = lrelu(gamma * (z - mean) * rstd + beta)
//   bwd : du = dy * (y > 0 ? 1 : slope);  S1 = sum du, S2 = sum du*zhat  (per n,c; ordered two-stage reduction)
//         dz = gamma * rstd * (du - S1/V - zhat * S2/V);  dgamma = sum_n S2;  dbeta = sum_n S1
// All kernels are HBM-bound streaming kernels: 128-bit accesses, grid sized to a multiple of the SM count.
#include "common.cuh"
#include "kernels.h"

namespace b2 {

template <typename T, int VW> struct Vec;
template <typename T> struct Vec<T, 8> {
    static __device__ __forceinline__ void ld(const T* p, float (&v)[8]) { load8(p, v); }
    static __device__ __forceinline__ void st(T* p, const float (&v)[8]) { store8(p, v); }
};
template <typename T> struct Vec<T, 4> {
    static __device__ __forceinline__ void ld(const T* p, float (&v)[4]) { load4(p, v); }
    static __device__ __forceinline__ void st(T* p, const float (&v)[4]) { store4(p, v); }
};
template <typename T> struct Vec<T, 1> {
    static __device__ __forceinline__ void ld(const T* p, float (&v)[1]) { v[0] = to_f(*p); }
    static __device__ __forceinline__ void st(T* p, const float (&v)[1]) { *p = from_f<T>(v[0]); }
};

// Streaming layout shared by the apply kernels: grid = (row blocks, samples); a thread owns ONE channel group (VW channels)
// for its whole life, so mean / rstd / gamma / beta are loaded once into registers (the first version re-read 4 x VW
// scalars per 16 bytes of data and ran at ~55 % of the HBM peak); rows are walked two at a time so that every thread has
// two independent vector loads per input tensor in flight.
template <typename T, int VW, int ROWS>
__global__ void __launch_bounds__(256) norm_fwd_kernel(const T* __restrict__ z, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       T* __restrict__ y, int n, long long vox, int c, int z_pitch,
                                                       int y_pitch, float slope) {
    pdl_grid_sync();
    const int ncg = c / VW, R = 256 / ncg;
    const int cg = threadIdx.x % ncg, r = threadIdx.x / ncg;
    if (r >= R) return;
    const int nn = blockIdx.y, c0 = cg * VW;
    float mean[VW], rstd[VW], ga[VW], be[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        mean[j] = stats[((long long)nn * c + c0 + j) * 2];
        rstd[j] = stats[((long long)nn * c + c0 + j) * 2 + 1];
        ga[j] = gamma[c0 + j];
        be[j] = beta[c0 + j];
    }
    const T* zp = z + (long long)nn * vox * z_pitch + c0;
    T* yp = y + (long long)nn * vox * y_pitch + c0;
    const long long step = (long long)gridDim.x * R;
    long long v = (long long)blockIdx.x * R + r;
    for (; v + (ROWS - 1) * step < vox; v += ROWS * step) {
        float a[ROWS][VW];
#pragma unroll
        for (int i = 0; i < ROWS; ++i) Vec<T, VW>::ld(zp + (v + i * step) * z_pitch, a[i]);
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            float o[VW];
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const float u = ga[j] * ((a[i][j] - mean[j]) * rstd[j]) + be[j];
                o[j] = u > 0.f ? u : u * slope;
            }
            Vec<T, VW>::st(yp + (v + i * step) * y_pitch, o);
        }
    }
    for (; v < vox; v += step) {
        float a0[VW], o0[VW];
        Vec<T, VW>::ld(zp + v * z_pitch, a0);
#pragma unroll
        for (int j = 0; j < VW; ++j) {
            const float u0 = ga[j] * ((a0[j] - mean[j]) * rstd[j]) + be[j];
            o0[j] = u0 > 0.f ? u0 : u0 * slope;
        }
        Vec<T, VW>::st(yp + v * y_pitch, o0);
    }
}

// sums[n][c] = {S1, S2} from the per-slab partials; dgamma[c] = sum_n S2; dbeta[c] = sum_n S1.
// One BLOCK (8 warps) per channel: threads stride the slabs with independent loads in flight, double accumulation, fixed
// xor-shuffle tree per warp and a fixed-order sum over the warps => bit-reproducible.  (Round 1: one warp per channel, four
// blocks per launch at c = 32: 13 us of serialised L2 round trips per layer.)
__global__ void __launch_bounds__(256) norm_bwd_finalize_kernel(const float* __restrict__ part, int n, int slabs, int c,
                                                                float* __restrict__ sums, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta) {
    pdl_grid_sync();
    __shared__ double sh1[8], sh2[8];
    const int cc = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double g = 0.0, b = 0.0;
    for (int nn = 0; nn < n; ++nn) {
        double s1 = 0.0, s2 = 0.0;
        const float* p = part + ((long long)nn * slabs * c + cc) * 2;
        for (int t0 = threadIdx.x; t0 < slabs; t0 += 1024) {
            float2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = t0 + 256 * u;
                v[u] = t < slabs ? *reinterpret_cast<const float2*>(p + (long long)t * c * 2) : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { s1 += (double)v[u].x; s2 += (double)v[u].y; }
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        __syncthreads();
        if (lane == 0) { sh1[warp] = s1; sh2[warp] = s2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a1 = 0.0, a2 = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) { a1 += sh1[w]; a2 += sh2[w]; }
            sums[((long long)nn * c + cc) * 2] = (float)a1;
            sums[((long long)nn * c + cc) * 2 + 1] = (float)a2;
            b += a1;
            g += a2;
        }
    }
    if (threadIdx.x == 0) {
        if (dgamma) dgamma[cc] = (float)g;
        if (dbeta) dbeta[cc] = (float)b;
    }
}

// RECOMPUTE: the sign of the pre-activation u = gamma*zhat + beta is recomputed from z (same fp32 expression as the
// forward kernel) instead of reading y: one tensor read less in both backward sweeps.
template <typename T, int VW, bool RECOMPUTE, int ROWS>
__global__ void __launch_bounds__(256, 3) norm_bwd_reduce_kernel(const T* __restrict__ z, const T* __restrict__ y,
                                                              const T* __restrict__ dy, const float* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              int slabs, long long vox, int c, int z_pitch, int y_pitch,
                                                              int dy_pitch, float slope, float* __restrict__ part) {
    pdl_grid_sync();
    extern __shared__ float sh[];  // [R][c][2]
    const int ncg = c / VW;
    const int R = 256 / ncg;
    const int n = blockIdx.y, slab = blockIdx.x;
    const int cg = threadIdx.x % ncg, r = threadIdx.x / ncg;
    const long long per = (vox + slabs - 1) / slabs;
    const long long v0 = (long long)slab * per, v1 = (v0 + per < vox) ? v0 + per : vox;
    float s1[VW], s2[VW], mean[VW], rstd[VW], ga[VW], be[VW];
    const int c0 = cg * VW;
    if (r < R) {
#pragma unroll
        for (int j = 0; j < VW; ++j) {
            s1[j] = 0.f; s2[j] = 0.f;
            mean[j] = stats[((long long)n * c + c0 + j) * 2];
            rstd[j] = stats[((long long)n * c + c0 + j) * 2 + 1];
            ga[j] = RECOMPUTE ? gamma[c0 + j] : 0.f;
            be[j] = RECOMPUTE ? beta[c0 + j] : 0.f;
        }
        // ROWS rows per iteration: ROWS independent vector loads per tensor in flight (the summation order per thread stays
        // v0+r, v0+r+R, ... so the result does not depend on the unrolling)
        long long v = v0 + r;
        for (; v + (ROWS - 1) * R < v1; v += ROWS * R) {
            const long long row0 = (long long)n * vox + v;
            float a[ROWS][VW], b[ROWS][VW], g[ROWS][VW];
#pragma unroll
            for (int i = 0; i < ROWS; ++i) {
                Vec<T, VW>::ld(z + (row0 + i * R) * z_pitch + c0, a[i]);
                if (!RECOMPUTE) Vec<T, VW>::ld(y + (row0 + i * R) * y_pitch + c0, b[i]);
                Vec<T, VW>::ld(dy + (row0 + i * R) * dy_pitch + c0, g[i]);
            }
#pragma unroll
            for (int i = 0; i < ROWS; ++i)
#pragma unroll
                for (int j = 0; j < VW; ++j) {
                    const float zh = (a[i][j] - mean[j]) * rstd[j];
                    const float u = RECOMPUTE ? ga[j] * zh + be[j] : b[i][j];
                    const float du = u > 0.f ? g[i][j] : g[i][j] * slope;
                    s1[j] += du;
                    s2[j] += du * zh;
                }
        }
        for (; v < v1; v += R) {
            const long long row = (long long)n * vox + v;
            float a[VW], b[VW], g[VW];
            Vec<T, VW>::ld(z + row * z_pitch + c0, a);
            if (!RECOMPUTE) Vec<T, VW>::ld(y + row * y_pitch + c0, b);
            Vec<T, VW>::ld(dy + row * dy_pitch + c0, g);
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const float zh = (a[j] - mean[j]) * rstd[j];
                const float u = RECOMPUTE ? ga[j] * zh + be[j] : b[j];
                const float du = u > 0.f ? g[j] : g[j] * slope;
                s1[j] += du;
                s2[j] += du * zh;
            }
        }
#pragma unroll
        for (int j = 0; j < VW; ++j) {
            sh[((size_t)r * c + c0 + j) * 2] = s1[j];
            sh[((size_t)r * c + c0 + j) * 2 + 1] = s2[j];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < c * 2; e += 256) {
        float s = 0.f;
        for (int q = 0; q < R; ++q) s += sh[(size_t)q * c * 2 + e];
        part[(((long long)n * slabs + slab) * c) * 2 + e] = s;
    }
}

template <typename T, int VW, bool RECOMPUTE, int ROWS>
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const T* __restrict__ z, const T* __restrict__ y,
                                                             const T* __restrict__ dy, const float* __restrict__ stats,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ sums, T* __restrict__ dz, int n,
                                                             long long vox, int c, int z_pitch, int y_pitch, int dy_pitch,
                                                             int dz_pitch, float slope, float inv_v) {
    pdl_grid_sync();
    const int ncg = c / VW, R = 256 / ncg;
    const int cg = threadIdx.x % ncg, r = threadIdx.x / ncg;
    if (r >= R) return;
    const int nn = blockIdx.y, c0 = cg * VW;
    float mean[VW], rstd[VW], ga[VW], be[VW], k1[VW], k2[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) {
        const long long sc = (long long)nn * c + c0 + j;
        mean[j] = stats[sc * 2];
        rstd[j] = stats[sc * 2 + 1];
        ga[j] = gamma[c0 + j];
        be[j] = RECOMPUTE ? beta[c0 + j] : 0.f;
        k1[j] = sums[sc * 2] * inv_v;
        k2[j] = sums[sc * 2 + 1] * inv_v;
    }
    const long long base = (long long)nn * vox;
    const long long step = (long long)gridDim.x * R;
    long long v = (long long)blockIdx.x * R + r;
    for (; v + (ROWS - 1) * step < vox; v += ROWS * step) {
        // ROWS independent rows: all loads issue before the first is consumed
        float a[ROWS][VW], b[ROWS][VW], g[ROWS][VW];
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            const long long row = base + v + i * step;
            Vec<T, VW>::ld(z + row * z_pitch + c0, a[i]);
            if (!RECOMPUTE) Vec<T, VW>::ld(y + row * y_pitch + c0, b[i]);
            Vec<T, VW>::ld(dy + row * dy_pitch + c0, g[i]);
        }
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
            float o[VW];
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const float zh = (a[i][j] - mean[j]) * rstd[j];
                const float u = RECOMPUTE ? ga[j] * zh + be[j] : b[i][j];
                const float du = u > 0.f ? g[i][j] : g[i][j] * slope;
                o[j] = ga[j] * rstd[j] * (du - k1[j] - zh * k2[j]);
            }
            Vec<T, VW>::st(dz + (base + v + i * step) * dz_pitch + c0, o);
        }
    }
    for (; v < vox; v += step) {
        const long long row = base + v;
        float a[VW], b[VW], g[VW], o[VW];
        Vec<T, VW>::ld(z + row * z_pitch + c0, a);
        if (!RECOMPUTE) Vec<T, VW>::ld(y + row * y_pitch + c0, b);
        Vec<T, VW>::ld(dy + row * dy_pitch + c0, g);
#pragma unroll
        for (int j = 0; j < VW; ++j) {
            const float zh = (a[j] - mean[j]) * rstd[j];
            const float u = RECOMPUTE ? ga[j] * zh + be[j] : b[j];
            const float du = u > 0.f ? g[j] : g[j] * slope;
            o[j] = ga[j] * rstd[j] * (du - k1[j] - zh * k2[j]);
        }
        Vec<T, VW>::st(dz + row * dz_pitch + c0, o);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Small tensors (the deep, low-resolution stages: <= 4096 voxels per sample).  There the three-launch sequences
// (statistics partials -> finalize -> apply) are pure launch / latency cost -- the whole tensor is L2-resident -- so one
// block per (8 channels, sample) does everything in two sweeps over its [vox][8] column strip.
//   forward : sweep 1 sum z, sum z^2 -> mean, rstd (written to `stats` for the backward); sweep 2 y = lrelu(gamma*zhat+beta)
//   backward: block per 8 channels, loops the samples: sweep 1 S1 = sum du, S2 = sum du*zhat; sweep 2 dz; dgamma, dbeta
// All block reductions use a fixed tree (warp shuffles, then the 8 warps in order) => bit-reproducible.
// ---------------------------------------------------------------------------------------------------------------
constexpr int NS_THREADS = 512, NS_WARPS = NS_THREADS / 32;   // small-tensor kernels: 16 warps per block (rows in flight)
template <int NV>
__device__ __forceinline__ void block_sum_vec(float (&v)[NV], float (*sh)[NV] /*[NS_WARPS][NV]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[warp][i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < NS_WARPS; ++w) t += sh[w][i];
        v[i] = t;
    }
}

template <typename T>
__global__ void __launch_bounds__(NS_THREADS) norm_small_fwd_kernel(const T* __restrict__ z, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, T* __restrict__ y,
                                                             float* __restrict__ stats, int vox, int c, int z_pitch, int y_pitch,
                                                             float slope, float eps) {
    pdl_grid_sync();
    __shared__ float sh[NS_WARPS][16];
    const int c0 = blockIdx.x * 8, n = blockIdx.y;
    const T* zp = z + (long long)n * vox * z_pitch + c0;
    float s[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = 0.f;
    for (int v = threadIdx.x; v < vox; v += NS_THREADS) {
        float a[8];
        load8(zp + (long long)v * z_pitch, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j] += a[j]; s[8 + j] += a[j] * a[j]; }
    }
    block_sum_vec<16>(s, sh);
    float mean[8], rstd[8], ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double m = (double)s[j] * (1.0 / (double)vox);
        double var = (double)s[8 + j] * (1.0 / (double)vox) - m * m;
        if (var < 0.0) var = 0.0;
        mean[j] = (float)m;
        rstd[j] = (float)(1.0 / sqrt(var + (double)eps));
        ga[j] = gamma[c0 + j];
        be[j] = beta[c0 + j];
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            stats[((long long)n * c + c0 + j) * 2] = mean[j];
            stats[((long long)n * c + c0 + j) * 2 + 1] = rstd[j];
        }
    }
    T* yp = y + (long long)n * vox * y_pitch + c0;
    for (int v = threadIdx.x; v < vox; v += NS_THREADS) {
        float a[8], o[8];
        load8(zp + (long long)v * z_pitch, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float u = ga[j] * ((a[j] - mean[j]) * rstd[j]) + be[j];
            o[j] = u > 0.f ? u : u * slope;
        }
        store8(yp + (long long)v * y_pitch, o);
    }
}

// The same kernel fed with the UN-REDUCED split-K partials of the producing convolution (conv3d_tc.cu: fp32
// [ksplit][otile][128 rows][BN]): pass 1 forms z = bf16(bias + sum_ks partial) in the reduce kernel's order (bit-identical z),
// writes it (the backward pass reads it) and accumulates the statistics of the ROUNDED values, as the two-launch path does;
// pass 2 re-reads z (L2-hot) and writes y.  Saves splitk_reduce_kernel and one round trip of z per deep-stage layer.
__global__ void __launch_bounds__(NS_THREADS) splitk_norm_small_fwd_kernel(SplitKDefer k, const float* __restrict__ bias,
                                                                           __nv_bfloat16* __restrict__ z, const float* __restrict__ gamma,
                                                                           const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                                           float* __restrict__ stats, int D, int H, int W, int c,
                                                                           int z_pitch, int y_pitch, float slope, float eps) {
    pdl_grid_sync();
    __shared__ float sh[NS_WARPS][16];
    const int c0 = blockIdx.x * 8, n = blockIdx.y, vox = D * H * W;
    const int nb = c0 / k.BN, col = c0 - nb * k.BN, tn = n / k.TN, n_ = n - tn * k.TN;
    __nv_bfloat16* zp = z + (long long)n * vox * z_pitch + c0;
    float bs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bs[j] = bias ? bias[c0 + j] : 0.f;
    float s[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) s[j] = 0.f;
    for (int v = threadIdx.x; v < vox; v += NS_THREADS) {
        const int w = v % W, h = (v / W) % H, d = v / (W * H);
        const int td = d / k.TD, d_ = d - td * k.TD, th = h / k.TH, h_ = h - th * k.TH, tw = w / k.TW, w_ = w - tw * k.TW;
        const int r = ((n_ * k.TD + d_) * k.TH + h_) * k.TW + w_;
        const long long otile = ((((long long)tn * k.nt_d + td) * k.nt_h + th) * k.nt_w + tw) * k.nblk + nb;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = bs[j];
        for (int ks = 0; ks < k.ksplit; ++ks) {
            const float* pp = k.partial + (((size_t)ks * k.otiles + otile) * 128 + r) * k.BN + col;
            const float4 a = *reinterpret_cast<const float4*>(pp), b = *reinterpret_cast<const float4*>(pp + 4);
            f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
        }
        store8(zp + (long long)v * z_pitch, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = __bfloat162float(__float2bfloat16_rn(f[j]));
            s[j] += a; s[8 + j] += a * a;
        }
    }
    block_sum_vec<16>(s, sh);       // (ends with a block barrier: every z row of this block is visible to it below)
    float mean[8], rstd[8], ga[8], be[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double m = (double)s[j] * (1.0 / (double)vox);
        double var = (double)s[8 + j] * (1.0 / (double)vox) - m * m;
        if (var < 0.0) var = 0.0;
        mean[j] = (float)m;
        rstd[j] = (float)(1.0 / sqrt(var + (double)eps));
        ga[j] = gamma[c0 + j];
        be[j] = beta[c0 + j];
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            stats[((long long)n * c + c0 + j) * 2] = mean[j];
            stats[((long long)n * c + c0 + j) * 2 + 1] = rstd[j];
        }
    }
    __nv_bfloat16* yp = y + (long long)n * vox * y_pitch + c0;
    for (int v = threadIdx.x; v < vox; v += NS_THREADS) {
        float a[8], o[8];
        load8(zp + (long long)v * z_pitch, a);      // written by this same thread in pass 1
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float u = ga[j] * ((a[j] - mean[j]) * rstd[j]) + be[j];
            o[j] = u > 0.f ? u : u * slope;
        }
        store8(yp + (long long)v * y_pitch, o);
    }
}

int g_splitk_fuse = 0;   // measured neutral at cfg2 (6.238 vs 6.209 ms/step, 230 vs 240 launches): off by default

int splitk_norm_small_fwd(const SplitKDefer& k, const float* bias, __nv_bfloat16* z, const float* gamma, const float* beta,
                          __nv_bfloat16* y, float* stats, int n, int D, int H, int W, int c, int z_pitch, int y_pitch, float slope,
                          float eps, cudaStream_t st) {
    B2_CHECK_ARG(k.deferred && k.partial && c % 8 == 0 && k.BN % 8 == 0 && z_pitch % 8 == 0 && y_pitch % 8 == 0);
    dim3 grid(c / 8, n);
    B2_LAUNCH(splitk_norm_small_fwd_kernel, grid, NS_THREADS, 0, st, k, bias, z, gamma, beta, y, stats, D, H, W, c, z_pitch, y_pitch, slope, eps);
    return B2_OK;
}

template <typename T, bool RECOMPUTE>
__global__ void __launch_bounds__(NS_THREADS) norm_small_bwd_kernel(const T* __restrict__ z, const T* __restrict__ y, const T* __restrict__ dy,
                                                             const float* __restrict__ stats, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, T* __restrict__ dz,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta, int n, int vox,
                                                             int c, int z_pitch, int y_pitch, int dy_pitch, int dz_pitch, float slope) {
    pdl_grid_sync();
    __shared__ float sh[NS_WARPS][16];
    const int c0 = blockIdx.x * 8;
    float ga[8], be[8], dg[8], db[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { ga[j] = gamma[c0 + j]; be[j] = RECOMPUTE ? beta[c0 + j] : 0.f; dg[j] = 0.f; db[j] = 0.f; }
    const float inv_v = (float)(1.0 / (double)vox);
    for (int nn = 0; nn < n; ++nn) {
        float mean[8], rstd[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mean[j] = stats[((long long)nn * c + c0 + j) * 2];
            rstd[j] = stats[((long long)nn * c + c0 + j) * 2 + 1];
        }
        const T* zp = z + (long long)nn * vox * z_pitch + c0;
        const T* yp = y + (long long)nn * vox * y_pitch + c0;
        const T* gp = dy + (long long)nn * vox * dy_pitch + c0;
        T* op = dz + (long long)nn * vox * dz_pitch + c0;
        float s[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) s[j] = 0.f;
        for (int v = threadIdx.x; v < vox; v += NS_THREADS) {
            float a[8], b[8], g[8];
            load8(zp + (long long)v * z_pitch, a);
            if (!RECOMPUTE) load8(yp + (long long)v * y_pitch, b);
            load8(gp + (long long)v * dy_pitch, g);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float zh = (a[j] - mean[j]) * rstd[j];
                const float u = RECOMPUTE ? ga[j] * zh + be[j] : b[j];
                const float du = u > 0.f ? g[j] : g[j] * slope;
                s[j] += du;
                s[8 + j] += du * zh;
            }
        }
        block_sum_vec<16>(s, sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) { db[j] += s[j]; dg[j] += s[8 + j]; }
        for (int v = threadIdx.x; v < vox; v += NS_THREADS) {
            float a[8], b[8], g[8], o[8];
            load8(zp + (long long)v * z_pitch, a);
            if (!RECOMPUTE) load8(yp + (long long)v * y_pitch, b);
            load8(gp + (long long)v * dy_pitch, g);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float zh = (a[j] - mean[j]) * rstd[j];
                const float u = RECOMPUTE ? ga[j] * zh + be[j] : b[j];
                const float du = u > 0.f ? g[j] : g[j] * slope;
                o[j] = ga[j] * rstd[j] * (du - s[j] * inv_v - zh * s[8 + j] * inv_v);
            }
            store8(op + (long long)v * dz_pitch, o);
        }
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (dgamma) dgamma[c0 + j] = dg[j];
            if (dbeta) dbeta[c0 + j] = db[j];
        }
    }
}

int g_norm_small = 1;
constexpr long long NORM_SMALL_VOX = 4096;

bool norm_small_supported(long long vox, int c, int p0, int p1, int p2, int p3) {
    return g_norm_small && vox <= NORM_SMALL_VOX && c % 8 == 0 && p0 % 8 == 0 && p1 % 8 == 0 && p2 % 8 == 0 && p3 % 8 == 0;
}

// statistics + apply in one launch (small tensors); writes `stats` like instnorm_stats would
template <typename T>
int norm_lrelu_fwd_small(const T* z, const float* gamma, const float* beta, T* y, float* stats, int n, long long vox, int c,
                         int z_pitch, int y_pitch, float slope, float eps, cudaStream_t st) {
    B2_CHECK_ARG(norm_small_supported(vox, c, z_pitch, y_pitch, 8, 8));
    dim3 grid(c / 8, n);
    B2_LAUNCH((norm_small_fwd_kernel<T>), grid, NS_THREADS, 0, st, z, gamma, beta, y, stats, (int)vox, c, z_pitch, y_pitch, slope, eps);
    return B2_OK;
}
template int norm_lrelu_fwd_small<float>(const float*, const float*, const float*, float*, float*, int, long long, int, int, int, float, float, cudaStream_t);
template int norm_lrelu_fwd_small<__nv_bfloat16>(const __nv_bfloat16*, const float*, const float*, __nv_bfloat16*, float*, int, long long, int, int, int, float, float, cudaStream_t);

static int pick_vw(int c, int p0, int p1, int p2, int p3, size_t esz) {
    auto ok = [&](int vw) {
        return c % vw == 0 && p0 % vw == 0 && p1 % vw == 0 && p2 % vw == 0 && p3 % vw == 0;
    };
    (void)esz;
    if (ok(8)) return 8;
    if (ok(4)) return 4;
    return 1;
}

// grid of the apply kernels: (row blocks, samples); ~16 resident blocks per SM overall, every block >= 2 row sweeps
static dim3 apply_grid(int n, long long vox, int ncg) {
    const int R = 256 / ncg;
    long long blocks = (vox + R - 1) / R;
    long long cap = ((long long)num_sms() * 16 + n - 1) / n;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return dim3((unsigned)blocks, (unsigned)n);
}

// streaming configuration: bit 0 -> 4-wide vectors instead of 8-wide (half the per-thread state: more resident blocks),
// bit 1 -> four rows in flight per thread instead of two
int g_norm_cfg = 3;

template <typename T>
int norm_lrelu_fwd(const T* z, const float* stats, const float* gamma, const float* beta, T* y, int n, long long vox,
                   int c, int z_pitch, int y_pitch, float slope, cudaStream_t st) {
    int vw = pick_vw(c, z_pitch, y_pitch, 8, 8, sizeof(T));
    if (vw == 8 && (g_norm_cfg & 1) && c / 4 <= 256) vw = 4;
    B2_CHECK_ARG(c / vw <= 256);
    const dim3 grid = apply_grid(n, vox, c / vw);
    const bool r4 = (g_norm_cfg & 2) != 0;
#define B2_NF(VW_, R_) B2_LAUNCH((norm_fwd_kernel<T, VW_, R_>), grid, 256, 0, st, z, stats, gamma, beta, y, n, vox, c, z_pitch, y_pitch, slope)
    if (vw == 8) { if (r4) B2_NF(8, 4); else B2_NF(8, 2); }
    else if (vw == 4) { if (r4) B2_NF(4, 4); else B2_NF(4, 2); }
    else B2_NF(1, 2);
#undef B2_NF
    return B2_OK;
}

static int norm_slabs(int n, long long vox) {
    long long want = (4LL * num_sms() + n - 1) / n;
    long long maxs = (vox + 63) / 64;
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    return (int)want;
}

size_t norm_bwd_scratch_floats(int n, long long vox, int c) {
    return (size_t)n * norm_slabs(n, vox) * c * 2 + (size_t)n * c * 2;
}

template <typename T, int VW, bool RC, int ROWS>
static int norm_bwd_launch(const T* z, const T* y, const T* dy, const float* stats, const float* gamma, const float* beta, T* dz,
                           float* dgamma, float* dbeta, int n, long long vox, int c, int z_pitch, int y_pitch, int dy_pitch,
                           int dz_pitch, float slope, float* scratch, cudaStream_t st) {
    const int slabs = norm_slabs(n, vox);
    float* part = scratch;
    float* sums = scratch + (size_t)n * slabs * c * 2;
    const int ncg = c / VW, R = 256 / ncg;
    const size_t sh = (size_t)R * c * 2 * sizeof(float);
    dim3 grid(slabs, n);
    B2_LAUNCH((norm_bwd_reduce_kernel<T, VW, RC, ROWS>), grid, 256, sh, st, z, y, dy, stats, gamma, beta, slabs, vox, c, z_pitch, y_pitch,
              dy_pitch, slope, part);
    B2_LAUNCH(norm_bwd_finalize_kernel, c, 256, 0, st, part, n, slabs, c, sums, dgamma, dbeta);
    const dim3 g2 = apply_grid(n, vox, c / VW);
    const float inv_v = (float)(1.0 / (double)vox);
    B2_LAUNCH((norm_bwd_apply_kernel<T, VW, RC, ROWS>), g2, 256, 0, st, z, y, dy, stats, gamma, beta, sums, dz, n, vox, c, z_pitch, y_pitch,
              dy_pitch, dz_pitch, slope, inv_v);
    return B2_OK;
}

// beta != nullptr: the activation sign is recomputed from z (y is not read); beta == nullptr: read y.
template <typename T>
int norm_lrelu_bwd(const T* z, const T* y, const T* dy, const float* stats, const float* gamma, const float* beta, T* dz,
                   float* dgamma, float* dbeta, int n, long long vox, int c, int z_pitch, int y_pitch, int dy_pitch, int dz_pitch,
                   float slope, float* scratch, cudaStream_t st) {
    B2_CHECK_ARG(c <= 1024);
    if (norm_small_supported(vox, c, z_pitch, y_pitch, dy_pitch, dz_pitch)) {
        if (beta) B2_LAUNCH((norm_small_bwd_kernel<T, true>), c / 8, NS_THREADS, 0, st, z, y, dy, stats, gamma, beta, dz, dgamma, dbeta, n, (int)vox, c,
                            z_pitch, y_pitch, dy_pitch, dz_pitch, slope);
        else B2_LAUNCH((norm_small_bwd_kernel<T, false>), c / 8, NS_THREADS, 0, st, z, y, dy, stats, gamma, beta, dz, dgamma, dbeta, n, (int)vox, c,
                       z_pitch, y_pitch, dy_pitch, dz_pitch, slope);
        return B2_OK;
    }
    int vw = pick_vw(c, z_pitch, y_pitch, dy_pitch, dz_pitch, sizeof(T));
    if (vw == 8 && (g_norm_cfg & 1) && c / 4 <= 256) vw = 4;
    B2_CHECK_ARG(c / vw <= 256);
    const bool r4 = (g_norm_cfg & 2) != 0;
#define B2_NB(VW_, RC_, R_) norm_bwd_launch<T, VW_, RC_, R_>(z, y, dy, stats, gamma, beta, dz, dgamma, dbeta, n, vox, c, z_pitch, y_pitch, \
                                                         dy_pitch, dz_pitch, slope, scratch, st)
    if (beta) {
        if (vw == 8) return r4 ? B2_NB(8, true, 4) : B2_NB(8, true, 2);
        if (vw == 4) return r4 ? B2_NB(4, true, 4) : B2_NB(4, true, 2);
        return B2_NB(1, true, 2);
    }
    if (vw == 8) return r4 ? B2_NB(8, false, 4) : B2_NB(8, false, 2);
    if (vw == 4) return r4 ? B2_NB(4, false, 4) : B2_NB(4, false, 2);
    return B2_NB(1, false, 2);
#undef B2_NB
}

template int norm_lrelu_fwd<float>(const float*, const float*, const float*, const float*, float*, int, long long, int, int, int, float, cudaStream_t);
template int norm_lrelu_fwd<__nv_bfloat16>(const __nv_bfloat16*, const float*, const float*, const float*, __nv_bfloat16*, int, long long, int, int, int, float, cudaStream_t);
template int norm_lrelu_bwd<float>(const float*, const float*, const float*, const float*, const float*, const float*, float*, float*, float*, int, long long, int, int, int, int, int, float, float*, cudaStream_t);
template int norm_lrelu_bwd<__nv_bfloat16>(const __nv_bfloat16*, const __nv_bfloat16*, const __nv_bfloat16*, const float*, const float*, const float*, __nv_bfloat16*, float*, float*, int, long long, int, int, int, int, int, float, float*, cudaStream_t);

}  // namespace b2
