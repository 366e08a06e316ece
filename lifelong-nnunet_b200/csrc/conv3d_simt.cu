// conv3d_simt.cu -- fp32-accumulate SIMT implicit-GEMM 3x3x3 convolution (forward, dgrad, wgrad) on NDHWC tensors.
//
// This is the PARITY path (B2_F32 activations: exact fp32 FMA arithmetic, logits within 1e-3 of the oracle) and the
// fallback for shapes the tcgen05 path (conv3d_tc.cu) does not cover (Cin < 16, odd channel counts).
// Replaces cuDNN's conv fwd/dgrad/wgrad behind nn.Conv3d of nnunet's ConvDropoutNormNonlin.
//
//   fwd   : z[n,o,co]  = bias[co] + sum_{t,ci} x[n, o*s + t - 1, ci] * Wf[t][ci][co]
//           + per-CTA partial (sum z, sum z^2) per (n,co) for the InstanceNorm that follows
//   dgrad : dx[n,i,ci] = sum_{t,co} [ (i - t + 1) % s == 0 ] dz[n, (i - t + 1)/s, co] * Wb[t][co][ci]
//   wgrad : dW[t][ci][co] = sum_{n,o} x[n, o*s + t - 1, ci] * dz[n,o,co]   (deterministic split + ordered reduce)
#include "common.cuh"
#include "kernels.h"

namespace b2 {

struct GatherGeom {
    int N;
    int Ds, Hs, Ws;  // spatial dims of the tensor being gathered (x for fwd, dz for dgrad)
    int Dd, Hd, Wd;  // spatial dims of the tensor being produced
    int Cs, Cd;      // GEMM K-per-tap / GEMM N
    int sd, sh, sw;  // stride of the convolution
    int src_pitch, dst_pitch;
    int tiles_per_sample;
};

constexpr int BM = 128, BK = 16, NTHREADS = 256;

// source coordinate of tap `t` for destination coordinate `o` along one axis; returns -1 when out of range
template <int MODE>
__device__ __forceinline__ int src_coord(int o, int t, int s, int S) {
    if (MODE == 0) {
        int i = o * s + t - 1;
        return (i >= 0 && i < S) ? i : -1;
    } else {
        int num = o - t + 1;
        if (num < 0) return -1;
        int q = num / s;
        if (q * s != num || q >= S) return -1;
        return q;
    }
}

template <typename T, int BN, int MODE, bool VEC>
__global__ void __launch_bounds__(NTHREADS) conv_gemm_kernel(GatherGeom g, const T* __restrict__ src,
                                                             const float* __restrict__ Wm,
                                                             const float* __restrict__ bias, T* __restrict__ dst,
                                                             int accumulate, float* __restrict__ stat_part) {
    pdl_grid_sync();
    constexpr int TN = BN / 16;
    __shared__ __align__(16) float As[BK][BM];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ float red[16][BN];

    const int tid = threadIdx.x;
    const int n = blockIdx.x / g.tiles_per_sample;
    const int tile = blockIdx.x % g.tiles_per_sample;
    const int n0 = blockIdx.y * BN;
    const long long Vd = (long long)g.Dd * g.Hd * g.Wd;
    const int Ktot = 27 * g.Cs;
    const int nchunks = (Ktot + BK - 1) / BK;

    // A-load role: one voxel row, 8 consecutive k
    const int am = tid % BM, akh = tid / BM;
    const long long av = (long long)tile * BM + am;
    const bool avalid = av < Vd;
    int od = 0, oh = 0, ow = 0;
    if (avalid) {
        ow = (int)(av % g.Wd);
        long long r = av / g.Wd;
        oh = (int)(r % g.Hd);
        od = (int)(r / g.Hd);
    }
    const T* src_n = src + (long long)n * g.Ds * g.Hs * g.Ws * g.src_pitch;

    const int ty = tid / 16, tx = tid % 16;
    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int ch = 0; ch < nchunks; ++ch) {
        // ---- gather A ----
        float a8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a8[j] = 0.f;
        if (VEC) {
            const int cpt = g.Cs / BK;
            const int tap = ch / cpt, c0 = (ch % cpt) * BK + akh * 8;
            if (avalid) {
                int id = src_coord<MODE>(od, tap / 9, g.sd, g.Ds);
                int ih = src_coord<MODE>(oh, (tap / 3) % 3, g.sh, g.Hs);
                int iw = src_coord<MODE>(ow, tap % 3, g.sw, g.Ws);
                if ((id | ih | iw) >= 0) {
                    const T* p = src_n + (((long long)id * g.Hs + ih) * g.Ws + iw) * g.src_pitch + c0;
                    load8(p, a8);
                }
            }
        } else {
            if (avalid) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    int k = ch * BK + akh * 8 + j;
                    if (k < Ktot) {
                        int tap = k / g.Cs, c = k % g.Cs;
                        int id = src_coord<MODE>(od, tap / 9, g.sd, g.Ds);
                        int ih = src_coord<MODE>(oh, (tap / 3) % 3, g.sh, g.Hs);
                        int iw = src_coord<MODE>(ow, tap % 3, g.sw, g.Ws);
                        if ((id | ih | iw) >= 0)
                            a8[j] = to_f(src_n[(((long long)id * g.Hs + ih) * g.Ws + iw) * g.src_pitch + c]);
                    }
                }
            }
        }
        // ---- load B ----
        float b[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int idx = tid * TN + j;
            int kr = idx / BN, col = idx % BN;
            int kg = ch * BK + kr;
            b[j] = (kg < Ktot && n0 + col < g.Cd) ? Wm[(long long)kg * g.Cd + n0 + col] : 0.f;
        }
        __syncthreads();  // previous iteration's reads done
#pragma unroll
        for (int j = 0; j < 8; ++j) As[akh * 8 + j][am] = a8[j];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int idx = tid * TN + j;
            Bs[idx / BN][idx % BN] = b[j];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[8], bb[TN];
            float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
            for (int j = 0; j < TN; ++j) bb[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
    }

    // ---- epilogue ----
    float s1[TN], s2[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    T* dst_n = dst + (long long)n * Vd * g.dst_pitch;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        long long v = (long long)tile * BM + ty * 8 + i;
        if (v < Vd) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int col = n0 + tx * TN + j;
                if (col < g.Cd) {
                    float r = acc[i][j];
                    if (bias) r += bias[col];
                    T* p = dst_n + v * g.dst_pitch + col;
                    if (accumulate) r += to_f(*p);
                    *p = from_f<T>(r);
                    s1[j] += r;
                    s2[j] += r * r;
                }
            }
        }
    }
    if (stat_part) {
        // fixed-order reduction over the 16 row-groups of the tile => bit-reproducible statistics
        float* out = stat_part + ((long long)(n * g.tiles_per_sample + tile) * g.Cd) * 2;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < TN; ++j) red[ty][tx * TN + j] = pass == 0 ? s1[j] : s2[j];
            __syncthreads();
            if (tid < BN && n0 + tid < g.Cd) {
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < 16; ++r) s += red[r][tid];
                out[(n0 + tid) * 2 + pass] = s;
            }
        }
    }
}

// mean / rstd from per-tile partial sums.  grid = (n, C); one BLOCK (8 warps) per (sample, channel): every thread owns the
// tiles t = tid + 256 k with four independent loads in flight, double accumulation, fixed xor-shuffle tree per warp and a
// fixed-order sum over the 8 warps => bit-reproducible.  (Round 1 used one warp per channel and 8 blocks per launch: five
// to ten serialised L2 round trips, 10-17 us for a kernel that moves 300 KB.)
__global__ void __launch_bounds__(256) stats_finalize_kernel(const float* __restrict__ part, int tiles, int C,
                                                             double inv_count, float eps, float* __restrict__ stats) {
    pdl_grid_sync();
    __shared__ double sh1[8], sh2[8];
    const int n = blockIdx.x, c = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* p = part + ((long long)n * tiles * C + c) * 2;
    double s1 = 0.0, s2 = 0.0;
    for (int t0 = threadIdx.x; t0 < tiles; t0 += 1024) {
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + 256 * u;
            v[u] = t < tiles ? *reinterpret_cast<const float2*>(p + (long long)t * C * 2) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { s1 += (double)v[u].x; s2 += (double)v[u].y; }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { sh1[warp] = s1; sh2[warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { a1 += sh1[w]; a2 += sh2[w]; }
        double mean = a1 * inv_count;
        double var = a2 * inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[((long long)n * C + c) * 2] = (float)mean;
        stats[((long long)n * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// PyTorch [Cout][Cin][27] -> Wf [27][Cin][Cout] and Wb [27][Cout][Cin]
__global__ void weight_shadow_kernel(const float* __restrict__ w, int Cout, int Cin, float* __restrict__ wf,
                                     float* __restrict__ wb) {
    pdl_grid_sync();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)Cout * Cin * 27;
    if (i >= tot) return;
    int t = (int)(i % 27);
    long long r = i / 27;
    int ci = (int)(r % Cin), co = (int)(r / Cin);
    float v = w[i];
    if (wf) wf[((long long)t * Cin + ci) * Cout + co] = v;
    if (wb) wb[((long long)t * Cout + co) * Cin + ci] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad: each CTA owns a (ci-block, co-block) pair and a strided list of output-voxel chunks; the x halo of the chunk
// and the dz chunk are staged in shared memory as fp32; 27 x 4 accumulators per thread; partials written once.
// ---------------------------------------------------------------------------------------------------------------
struct WgradGeom {
    int N, Di, Hi, Wi, Do, Ho, Wo, Cin, Cout, sd, sh, sw, x_pitch, dz_pitch;
    int cd, chh, cw;        // chunk extent in output voxels
    int hd, hh, hw;         // halo extent in input voxels
    int nd, nh, nw;         // chunks per axis
    int nchunks, nsplit;    // chunks over all samples; CTAs along grid.x
};

template <typename T>
__global__ void __launch_bounds__(256) wgrad_kernel(WgradGeom g, const T* __restrict__ x, const T* __restrict__ dz,
                                                    float* __restrict__ part_w, float* __restrict__ part_b) {
    pdl_grid_sync();
    extern __shared__ __align__(16) float smem[];
    const int halo_vox = g.hd * g.hh * g.hw;
    const int chunk_vox = g.cd * g.chh * g.cw;
    float* xs = smem;                      // [halo_vox][32]
    float* zs = smem + (size_t)halo_vox * 32;  // [chunk_vox][32]

    const int tid = threadIdx.x;
    const int ci0 = blockIdx.y * 32, co0 = blockIdx.z * 32;
    const int ci = tid >> 3, cog = tid & 7;
    float acc[27][4];
#pragma unroll
    for (int t = 0; t < 27; ++t)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    const bool do_bias = (blockIdx.y == 0) && (ci == 0) && part_b;

    for (int chunk = blockIdx.x; chunk < g.nchunks; chunk += g.nsplit) {
        int c = chunk;
        const int kw = c % g.nw; c /= g.nw;
        const int kh = c % g.nh; c /= g.nh;
        const int kd = c % g.nd; c /= g.nd;
        const int n = c;
        const int od0 = kd * g.cd, oh0 = kh * g.chh, ow0 = kw * g.cw;
        const int id0 = od0 * g.sd - 1, ih0 = oh0 * g.sh - 1, iw0 = ow0 * g.sw - 1;
        __syncthreads();
        // stage x halo: 32 channels per voxel
        for (int e = tid; e < halo_vox * 32; e += 256) {
            int cc = e & 31, hv = e >> 5;
            int w_ = hv % g.hw, r = hv / g.hw;
            int h_ = r % g.hh, d_ = r / g.hh;
            int id = id0 + d_, ih = ih0 + h_, iw = iw0 + w_;
            float v = 0.f;
            if (id >= 0 && id < g.Di && ih >= 0 && ih < g.Hi && iw >= 0 && iw < g.Wi && ci0 + cc < g.Cin)
                v = to_f(x[((((long long)n * g.Di + id) * g.Hi + ih) * g.Wi + iw) * g.x_pitch + ci0 + cc]);
            xs[e] = v;
        }
        for (int e = tid; e < chunk_vox * 32; e += 256) {
            int cc = e & 31, cv = e >> 5;
            int w_ = cv % g.cw, r = cv / g.cw;
            int h_ = r % g.chh, d_ = r / g.chh;
            int od = od0 + d_, oh = oh0 + h_, ow = ow0 + w_;
            float v = 0.f;
            if (od < g.Do && oh < g.Ho && ow < g.Wo && co0 + cc < g.Cout)
                v = to_f(dz[((((long long)n * g.Do + od) * g.Ho + oh) * g.Wo + ow) * g.dz_pitch + co0 + cc]);
            zs[e] = v;
        }
        __syncthreads();
        for (int d_ = 0; d_ < g.cd; ++d_)
            for (int h_ = 0; h_ < g.chh; ++h_)
                for (int w_ = 0; w_ < g.cw; ++w_) {
                    const int cv = (d_ * g.chh + h_) * g.cw + w_;
                    const float4 z4 = *reinterpret_cast<const float4*>(&zs[cv * 32 + cog * 4]);
                    if (do_bias) { bsum[0] += z4.x; bsum[1] += z4.y; bsum[2] += z4.z; bsum[3] += z4.w; }
                    const int hb = ((d_ * g.sd) * g.hh + h_ * g.sh) * g.hw + w_ * g.sw;
#pragma unroll
                    for (int t = 0; t < 27; ++t) {
                        const int off = ((t / 9) * g.hh + (t / 3) % 3) * g.hw + (t % 3);
                        const float xv = xs[(hb + off) * 32 + ci];
                        acc[t][0] = fmaf(xv, z4.x, acc[t][0]);
                        acc[t][1] = fmaf(xv, z4.y, acc[t][1]);
                        acc[t][2] = fmaf(xv, z4.z, acc[t][2]);
                        acc[t][3] = fmaf(xv, z4.w, acc[t][3]);
                    }
                }
    }
    // partial layout: [split][27][Cin][Cout]
    if (ci0 + ci < g.Cin) {
        float* out = part_w + (long long)blockIdx.x * 27 * g.Cin * g.Cout;
#pragma unroll
        for (int t = 0; t < 27; ++t)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int co = co0 + cog * 4 + j;
                if (co < g.Cout) out[((long long)t * g.Cin + ci0 + ci) * g.Cout + co] = acc[t][j];
            }
    }
    if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = co0 + cog * 4 + j;
            if (co < g.Cout) part_b[(long long)blockIdx.x * g.Cout + co] = bsum[j];
        }
    }
}

// dW_pt[co][ci][t] = sum_split part[split][t][ci][co]  (ordered);  dbias likewise
__global__ void wgrad_reduce_kernel(const float* __restrict__ part_w, const float* __restrict__ part_b, int nsplit,
                                    int Cin, int Cout, float* __restrict__ dw, float* __restrict__ db) {
    pdl_grid_sync();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = 27LL * Cin * Cout;
    if (i < tot) {
        float s = 0.f;
        for (int k = 0; k < nsplit; ++k) s += part_w[(long long)k * tot + i];
        int co = (int)(i % Cout);
        long long r = i / Cout;
        int ci = (int)(r % Cin), t = (int)(r / Cin);
        dw[((long long)co * Cin + ci) * 27 + t] = s;
    } else if (db && i < tot + Cout) {
        int co = (int)(i - tot);
        float s = 0.f;
        for (int k = 0; k < nsplit; ++k) s += part_b[(long long)k * Cout + co];
        db[co] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-(n,c) statistics of a bf16/fp32 NDHWC tensor (InstanceNorm), two-stage ordered reduction
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int VW>
__global__ void __launch_bounds__(256) stats_reduce_kernel(const T* __restrict__ z, int slabs, long long vox, int c, int pitch,
                                                           float* __restrict__ part) {
    pdl_grid_sync();
    extern __shared__ float sh[];  // [R][c][2]
    const int ncg = c / VW;
    const int R = 256 / ncg;
    const int n = blockIdx.y, slab = blockIdx.x;
    const int cg = threadIdx.x % ncg, r = threadIdx.x / ncg;
    const long long per = (vox + slabs - 1) / slabs;
    const long long v0 = (long long)slab * per, v1 = (v0 + per < vox) ? v0 + per : vox;
    const int c0 = cg * VW;
    if (r < R) {
        float s1[VW], s2[VW];
#pragma unroll
        for (int j = 0; j < VW; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
        for (long long v = v0 + r; v < v1; v += R) {
            float a[VW];
            if (VW == 8) load8(z + ((long long)n * vox + v) * pitch + c0, *reinterpret_cast<float(*)[8]>(a));
            else a[0] = to_f(z[((long long)n * vox + v) * pitch + c0]);
#pragma unroll
            for (int j = 0; j < VW; ++j) { s1[j] += a[j]; s2[j] += a[j] * a[j]; }
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

static int stats_slabs(int n, long long vox) {
    long long want = (4LL * num_sms() + n - 1) / n;
    long long maxs = (vox + 63) / 64;   // small volumes: still several blocks (a single block is latency-bound)
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    return (int)want;
}


size_t instnorm_stats_scratch_floats(int n, long long vox, int c) { return (size_t)n * stats_slabs(n, vox) * c * 2; }

template <typename T>
int instnorm_stats(const T* z, int n, long long vox, int c, int pitch, float* part, float* stats, float eps, cudaStream_t st) {
    const bool v8 = (c % 8 == 0) && (pitch % 8 == 0);
    const int vw = v8 ? 8 : 1;
    B2_CHECK_ARG(c / vw <= 256);
    const int slabs = stats_slabs(n, vox);
    const int R = 256 / (c / vw);
    const size_t sh = (size_t)R * c * 2 * sizeof(float);
    dim3 grid(slabs, n);
    if (v8) B2_LAUNCH((stats_reduce_kernel<T, 8>), grid, 256, sh, st, z, slabs, vox, c, pitch, part);
    else B2_LAUNCH((stats_reduce_kernel<T, 1>), grid, 256, sh, st, z, slabs, vox, c, pitch, part);
    dim3 g2(n, c);
    B2_LAUNCH(stats_finalize_kernel, g2, 256, 0, st, part, slabs, c, 1.0 / (double)vox, eps, stats);
    return B2_OK;
}
int stats_finalize(const float* part, int slots, int n, long long vox, int c, float eps, float* stats, cudaStream_t st) {
    dim3 g2(n, c);
    B2_LAUNCH(stats_finalize_kernel, g2, 256, 0, st, part, slots, c, 1.0 / (double)vox, eps, stats);
    return B2_OK;
}
template int instnorm_stats<float>(const float*, int, long long, int, int, float*, float*, float, cudaStream_t);
template int instnorm_stats<__nv_bfloat16>(const __nv_bfloat16*, int, long long, int, int, float*, float*, float, cudaStream_t);

static inline int out_dim_(int i, int s) { return (i + 2 - 3) / s + 1; }

// ---------------------------------------------------------------------------------------------------------------
// wgrad for Cin <= 4 (the first layer): thread = (output channel, row lane); each CTA walks (n, d, h) rows of the output,
// streaming dz once (HBM-bound) and reading the 27 x-neighbours through L1; per-CTA partials, ordered reduce.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int CIN>
__global__ void __launch_bounds__(256) wgrad_smallcin_kernel(WgradGeom g, const T* __restrict__ x, const T* __restrict__ dz,
                                                             float* __restrict__ part_w, float* __restrict__ part_b, int lanes) {
    pdl_grid_sync();
    extern __shared__ float sh[];  // [lanes][27*CIN + 1][Cout]
    const int cout = g.Cout;
    const int co = threadIdx.x % cout, lane = threadIdx.x / cout;
    float acc[27 * CIN];
#pragma unroll
    for (int i = 0; i < 27 * CIN; ++i) acc[i] = 0.f;
    float bsum = 0.f;
    const long long rows = (long long)g.N * g.Do * g.Ho;
    if (lane < lanes) {
        for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
            const int oh = (int)(r % g.Ho);
            long long t = r / g.Ho;
            const int od = (int)(t % g.Do), n = (int)(t / g.Do);
            const T* zrow = dz + (((long long)n * g.Do + od) * g.Ho + oh) * g.Wo * g.dz_pitch;
            const T* xrow[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int id = od * g.sd + k / 3 - 1, ih = oh * g.sh + k % 3 - 1;
                xrow[k] = (id >= 0 && id < g.Di && ih >= 0 && ih < g.Hi)
                              ? x + (((long long)n * g.Di + id) * g.Hi + ih) * g.Wi * g.x_pitch : nullptr;
            }
            for (int ow = lane; ow < g.Wo; ow += lanes) {
                const float z = to_f(zrow[(long long)ow * g.dz_pitch + co]);
                bsum += z;
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    if (xrow[k] == nullptr) continue;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int iw = ow * g.sw + kw - 1;
                        if (iw < 0 || iw >= g.Wi) continue;
#pragma unroll
                        for (int ci = 0; ci < CIN; ++ci)
                            acc[(k * 3 + kw) * CIN + ci] = fmaf(to_f(xrow[k][(long long)iw * g.x_pitch + ci]), z, acc[(k * 3 + kw) * CIN + ci]);
                    }
                }
            }
        }
        float* s = sh + (size_t)lane * (27 * CIN + 1) * cout;
#pragma unroll
        for (int i = 0; i < 27 * CIN; ++i) s[i * cout + co] = acc[i];
        s[27 * CIN * cout + co] = bsum;
    }
    __syncthreads();
    // ordered sum over lanes; partial layout [split][27][Cin][Cout] (+ bias partial)
    for (int e = threadIdx.x; e < (27 * CIN + 1) * cout; e += 256) {
        float s = 0.f;
        for (int l = 0; l < lanes; ++l) s += sh[(size_t)l * (27 * CIN + 1) * cout + e];
        if (e < 27 * CIN * cout) part_w[(long long)blockIdx.x * 27 * CIN * cout + e] = s;
        else if (part_b) part_b[(long long)blockIdx.x * cout + (e - 27 * CIN * cout)] = s;
    }
}

static int smallcin_splits(const ConvShape& s) {
    long long rows = (long long)s.n * out_dim_(s.d, s.stride[0]) * out_dim_(s.h, s.stride[1]);
    long long want = 8LL * num_sms();
    return (int)(rows < want ? rows : want);
}

// ---------------------------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------------------------
static inline int out_dim(int i, int s) { return (i + 2 - 3) / s + 1; }

template <typename T, int MODE>
static int launch_gemm(const GatherGeom& g, const T* src, const float* Wm, const float* bias, T* dst, int accumulate,
                       float* stat_part, cudaStream_t st) {
    const bool vec = (g.Cs % 16 == 0) && (g.src_pitch % 8 == 0);
    const bool wide = g.Cd > 32;
    dim3 grid(g.N * g.tiles_per_sample, wide ? cdiv(g.Cd, 64) : 1);
    if (wide) {
        if (vec) B2_LAUNCH((conv_gemm_kernel<T, 64, MODE, true>), grid, NTHREADS, 0, st, g, src, Wm, bias, dst, accumulate, stat_part);
        else     B2_LAUNCH((conv_gemm_kernel<T, 64, MODE, false>), grid, NTHREADS, 0, st, g, src, Wm, bias, dst, accumulate, stat_part);
    } else {
        if (vec) B2_LAUNCH((conv_gemm_kernel<T, 32, MODE, true>), grid, NTHREADS, 0, st, g, src, Wm, bias, dst, accumulate, stat_part);
        else     B2_LAUNCH((conv_gemm_kernel<T, 32, MODE, false>), grid, NTHREADS, 0, st, g, src, Wm, bias, dst, accumulate, stat_part);
    }
    return B2_OK;
}

size_t conv_stat_part_floats(const ConvShape& s) {
    long long Vo = (long long)out_dim(s.d, s.stride[0]) * out_dim(s.h, s.stride[1]) * out_dim(s.w, s.stride[2]);
    return (size_t)s.n * cdiv(Vo, BM) * s.cout * 2;
}

template <typename T>
int conv3d_fwd_simt(const ConvShape& s, const T* x, const float* wf, const float* bias, T* z, float* stat_part,
                    float* stats, float eps, cudaStream_t st) {
    GatherGeom g;
    g.N = s.n; g.Ds = s.d; g.Hs = s.h; g.Ws = s.w;
    g.Dd = out_dim(s.d, s.stride[0]); g.Hd = out_dim(s.h, s.stride[1]); g.Wd = out_dim(s.w, s.stride[2]);
    g.Cs = s.cin; g.Cd = s.cout; g.sd = s.stride[0]; g.sh = s.stride[1]; g.sw = s.stride[2];
    g.src_pitch = s.in_pitch; g.dst_pitch = s.out_pitch;
    long long Vd = (long long)g.Dd * g.Hd * g.Wd;
    g.tiles_per_sample = cdiv(Vd, BM);
    int rc = launch_gemm<T, 0>(g, x, wf, bias, z, 0, stats ? stat_part : nullptr, st);
    if (rc) return rc;
    if (stats) {
        dim3 grid(s.n, s.cout);
        B2_LAUNCH(stats_finalize_kernel, grid, 256, 0, st, stat_part, g.tiles_per_sample, s.cout, 1.0 / (double)Vd, eps, stats);
    }
    return B2_OK;
}

template <typename T>
int conv3d_dgrad_simt(const ConvShape& s, const T* dz, const float* wb, T* dx, int accumulate, cudaStream_t st) {
    GatherGeom g;
    g.N = s.n;
    g.Ds = out_dim(s.d, s.stride[0]); g.Hs = out_dim(s.h, s.stride[1]); g.Ws = out_dim(s.w, s.stride[2]);
    g.Dd = s.d; g.Hd = s.h; g.Wd = s.w;
    g.Cs = s.cout; g.Cd = s.cin; g.sd = s.stride[0]; g.sh = s.stride[1]; g.sw = s.stride[2];
    g.src_pitch = s.out_pitch; g.dst_pitch = s.in_pitch;
    long long Vd = (long long)g.Dd * g.Hd * g.Wd;
    g.tiles_per_sample = cdiv(Vd, BM);
    return launch_gemm<T, 1>(g, dz, wb, nullptr, dx, accumulate, nullptr, st);
}

static void wgrad_plan(const ConvShape& s, WgradGeom& g) {
    g.N = s.n; g.Di = s.d; g.Hi = s.h; g.Wi = s.w;
    g.Do = out_dim(s.d, s.stride[0]); g.Ho = out_dim(s.h, s.stride[1]); g.Wo = out_dim(s.w, s.stride[2]);
    g.Cin = s.cin; g.Cout = s.cout; g.sd = s.stride[0]; g.sh = s.stride[1]; g.sw = s.stride[2];
    g.x_pitch = s.in_pitch; g.dz_pitch = s.out_pitch;
    // chunk: start from 4x8x8 outputs and shrink until the fp32 halo fits in ~100 KB of shared memory
    int cd = 4, ch = 8, cw = 8;
    auto smem = [&](int a, int b, int c) {
        long long halo = (long long)((a - 1) * g.sd + 3) * ((b - 1) * g.sh + 3) * ((c - 1) * g.sw + 3);
        return (halo + (long long)a * b * c) * 32 * 4;
    };
    while (smem(cd, ch, cw) > 100 * 1024) {
        if (cd > 1 && cd * g.sd >= ch * g.sh) cd /= 2;
        else if (ch > 1 && ch * g.sh >= cw * g.sw) ch /= 2;
        else if (cw > 1) cw /= 2;
        else if (cd > 1) cd /= 2;
        else ch /= 2;
    }
    if (cd > g.Do) cd = g.Do;
    if (ch > g.Ho) ch = g.Ho;
    if (cw > g.Wo) cw = g.Wo;
    g.cd = cd; g.chh = ch; g.cw = cw;
    g.hd = (cd - 1) * g.sd + 3; g.hh = (ch - 1) * g.sh + 3; g.hw = (cw - 1) * g.sw + 3;
    g.nd = cdiv(g.Do, cd); g.nh = cdiv(g.Ho, ch); g.nw = cdiv(g.Wo, cw);
    g.nchunks = g.N * g.nd * g.nh * g.nw;
    int blocks = cdiv(g.Cin, 32) * cdiv(g.Cout, 32);
    int target = 2 * num_sms();
    int ns = target / blocks;
    if (ns < 1) ns = 1;
    if (ns > g.nchunks) ns = g.nchunks;
    g.nsplit = ns;
}

size_t conv_wgrad_part_floats(const ConvShape& s) {
    if (s.cin <= 4 && s.cout <= 256) return (size_t)smallcin_splits(s) * (27ULL * s.cin * s.cout + s.cout);
    WgradGeom g;
    wgrad_plan(s, g);
    return (size_t)g.nsplit * (27ULL * s.cin * s.cout + s.cout);
}

template <typename T>
int conv3d_wgrad_simt(const ConvShape& s, const T* x, const T* dz, float* part, float* dw, float* dbias,
                      cudaStream_t st) {
    WgradGeom g;
    wgrad_plan(s, g);
    if (s.cin <= 4 && s.cout <= 256) {
        const int ns = smallcin_splits(s);
        float* pw = part;
        float* pb = part + (size_t)ns * 27 * s.cin * s.cout;
        int lanes = 256 / s.cout;
        const size_t per_lane = (size_t)(27 * s.cin + 1) * s.cout * sizeof(float);
        if ((size_t)lanes * per_lane > 48 * 1024) lanes = (int)(48 * 1024 / per_lane);
        B2_CHECK_ARG(lanes >= 1);
        const size_t sh = (size_t)lanes * per_lane;
        switch (s.cin) {
            case 1: B2_LAUNCH((wgrad_smallcin_kernel<T, 1>), ns, 256, sh, st, g, x, dz, pw, dbias ? pb : nullptr, lanes); break;
            case 2: B2_LAUNCH((wgrad_smallcin_kernel<T, 2>), ns, 256, sh, st, g, x, dz, pw, dbias ? pb : nullptr, lanes); break;
            case 3: B2_LAUNCH((wgrad_smallcin_kernel<T, 3>), ns, 256, sh, st, g, x, dz, pw, dbias ? pb : nullptr, lanes); break;
            default: B2_LAUNCH((wgrad_smallcin_kernel<T, 4>), ns, 256, sh, st, g, x, dz, pw, dbias ? pb : nullptr, lanes); break;
        }
        long long tot2 = 27LL * s.cin * s.cout + (dbias ? s.cout : 0);
        B2_LAUNCH(wgrad_reduce_kernel, cdiv(tot2, 256), 256, 0, st, pw, pb, ns, s.cin, s.cout, dw, dbias);
        return B2_OK;
    }
    float* part_w = part;
    float* part_b = part + (size_t)g.nsplit * 27 * s.cin * s.cout;
    size_t smem = ((size_t)g.hd * g.hh * g.hw + (size_t)g.cd * g.chh * g.cw) * 32 * sizeof(float);
    static bool attr_done[2] = {false, false};
    constexpr int ti = sizeof(T) == 4 ? 0 : 1;
    if (!attr_done[ti]) {
        B2_CUDA(cudaFuncSetAttribute(wgrad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        attr_done[ti] = true;
    }
    dim3 grid(g.nsplit, cdiv(s.cin, 32), cdiv(s.cout, 32));
    B2_LAUNCH(wgrad_kernel<T>, grid, 256, smem, st, g, x, dz, part_w, dbias ? part_b : nullptr);
    long long tot = 27LL * s.cin * s.cout + (dbias ? s.cout : 0);
    B2_LAUNCH(wgrad_reduce_kernel, cdiv(tot, 256), 256, 0, st, part_w, part_b, g.nsplit, s.cin, s.cout, dw, dbias);
    return B2_OK;
}

// Tiled ordered reduction for the tensor-core wgrad partials.  A block of 32 (co) x 8 (k lanes) threads owns one ci, 32
// consecutive co and up to NT consecutive taps: every load is a coalesced 128-byte row of part[k][t][ci][co0..co0+31], the
// tap loop is fully unrolled so a thread keeps NT (x4 over k) independent loads in flight; the 8 k-lanes are combined in a
// fixed order through shared memory, and the result is written as runs of consecutive taps of dw_pt[co][ci][t].
// (A thread-per-output kernel walks the partials serially and writes 4-byte elements with a 108-byte stride.)
// KB = partials per load batch (4: hundreds of partials of a thin layer; 1: the <= 8 partials of a deep layer, where the
// batch registers only cost occupancy -- the kernel is then bound by load latency x resident blocks)
template <int NT, int KB>
__global__ void __launch_bounds__(256, KB == 1 ? 3 : 1) wgrad_reduce_tiled_kernel(const float* __restrict__ part, int nsplit, int Cin, int Cout,
                                                                                  float* __restrict__ dw) {
    pdl_grid_sync();
    __shared__ float sh[8][NT][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int cob = Cout >> 5;
    const int ci = blockIdx.x / cob, co0 = (blockIdx.x % cob) << 5;
    const int t0 = blockIdx.y * NT;
    const long long tot = 27LL * Cin * Cout;
    const float* src = part + ((long long)t0 * Cin + ci) * Cout + co0 + tx;
    const long long tstride = (long long)Cin * Cout;
    float acc[NT];
    const int nt = 27 - t0 < NT ? 27 - t0 : NT;
    long long toff[NT];          // taps beyond the 27th re-read tap 0 of the block (branch-free loads; their sums are never stored)
#pragma unroll
    for (int t = 0; t < NT; ++t) { acc[t] = 0.f; toff[t] = (t < nt ? t : 0) * tstride; }
    // explicit load batches: all loads of a batch are issued before the first add (a predicated `acc += load` compiled to
    // one exposed L2 round trip per element: 80 us for 32 MB)
    int k = ty;
    if (KB == 4) {
        for (; k + 24 < nsplit; k += 32) {
            float tmp[4][NT];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int t = 0; t < NT; ++t) tmp[u][t] = __ldcs(src + (long long)(k + 8 * u) * tot + toff[t]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int t = 0; t < NT; ++t) acc[t] += tmp[u][t];
        }
    }
    for (; k < nsplit; k += 8) {
        float tmp[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) tmp[t] = __ldcs(src + (long long)k * tot + toff[t]);
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t] += tmp[t];
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) sh[ty][t][tx] = acc[t];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * nt; i += 256) {
        const int col = i / nt, t = i - col * nt;
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += sh[q][t][col];
        dw[((long long)(co0 + col) * Cin + ci) * 27 + t0 + t] = v;
    }
}

int wgrad_reduce(const float* part_w, const float* part_b, int nsplit, int cin, int cout, float* dw, float* db, cudaStream_t st) {
    if (!db && cout % 32 == 0) {
        const int xb = cin * (cout / 32);
        // few (ci, co-block) pairs (thin layers, hundreds of partials): split the taps over grid.y for parallelism
        if (xb >= 4 * num_sms()) {
            dim3 grid(xb, 1);
            if (nsplit > 8) B2_LAUNCH((wgrad_reduce_tiled_kernel<27, 4>), grid, 256, 0, st, part_w, nsplit, cin, cout, dw);
            else B2_LAUNCH((wgrad_reduce_tiled_kernel<27, 1>), grid, 256, 0, st, part_w, nsplit, cin, cout, dw);
        } else if (xb * 3 >= 4 * num_sms()) {
            dim3 grid(xb, 3);
            B2_LAUNCH((wgrad_reduce_tiled_kernel<9, 4>), grid, 256, 0, st, part_w, nsplit, cin, cout, dw);
        } else {
            dim3 grid(xb, 9);
            B2_LAUNCH((wgrad_reduce_tiled_kernel<3, 4>), grid, 256, 0, st, part_w, nsplit, cin, cout, dw);
        }
        return B2_OK;
    }
    long long tot = 27LL * cin * cout + (db ? cout : 0);
    B2_LAUNCH(wgrad_reduce_kernel, cdiv(tot, 256), 256, 0, st, part_w, part_b, nsplit, cin, cout, dw, db);
    return B2_OK;
}

int weight_shadow(const float* w, int cout, int cin, float* wf, float* wb, cudaStream_t st) {
    if (!wf && !wb) return B2_OK;
    long long tot = (long long)cout * cin * 27;
    B2_LAUNCH(weight_shadow_kernel, cdiv(tot, 256), 256, 0, st, w, cout, cin, wf, wb);
    return B2_OK;
}

template int conv3d_fwd_simt<float>(const ConvShape&, const float*, const float*, const float*, float*, float*, float*, float, cudaStream_t);
template int conv3d_fwd_simt<__nv_bfloat16>(const ConvShape&, const __nv_bfloat16*, const float*, const float*, __nv_bfloat16*, float*, float*, float, cudaStream_t);
template int conv3d_dgrad_simt<float>(const ConvShape&, const float*, const float*, float*, int, cudaStream_t);
template int conv3d_dgrad_simt<__nv_bfloat16>(const ConvShape&, const __nv_bfloat16*, const float*, __nv_bfloat16*, int, cudaStream_t);
template int conv3d_wgrad_simt<float>(const ConvShape&, const float*, const float*, float*, float*, float*, cudaStream_t);
template int conv3d_wgrad_simt<__nv_bfloat16>(const ConvShape&, const __nv_bfloat16*, const __nv_bfloat16*, float*, float*, float*, cudaStream_t);

}  // namespace b2
