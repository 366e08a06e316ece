// vit.cu -- the 3D Vision Transformer of Generic_ViT_UNet V1 (reference nnunet_ext/network_architecture/
// vision_transformer.py: PatchEmbed :16-79, Attention :120-151, Block :153-198, forward_features / forward :418-458; head
// output reshaped into the bottleneck, generic_ViT_UNet.py:253), forward and backward, bf16 operands / fp32 accumulation.
//
//   * every Linear (patch embedding = Conv3d k = s = patch as a GEMM over non-overlapping patches, qkv, proj, fc1, fc2)
//     runs on the tcgen05 gather-GEMM of conv3d_tc.cu (gemm_tn_bf16: one tap, TMA-staged operands, TMEM accumulators,
//     split-K with an ordered fp32 reduction when a launch has few output tiles).  Backward: dX = dY W uses the transposed
//     bf16 weight shadow, dW = dY^T X uses transposed bf16 activation copies and writes fp32 straight into the gradient.
//   * LayerNorm, GELU, residual adds, token assembly, bias gradients: HBM-bound elementwise / two-stage reductions.
//   * attention over T = 433 tokens x 64 dims per head: 1.2 GFLOP per pass -- flash-style warp-MMA kernels per (batch, head,
//     64-row block), online softmax in fp32, backward by recomputation from the saved log-sum-exp (no T x T tensor in HBM).
//   * head Linear(E -> prod(bottleneck)) acts on the B class tokens only: a streaming GEMV over the fp32 weight (212 MB at
//     cfg4), written straight into the plan's NDHWC bottleneck activation; backward = outer product + ordered GEMV.
// All reductions are fixed-order (bit-reproducible).  The residual stream is kept in fp32.
#include <string.h>

#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace b2 {

// ---------------------------------------------------------------------------------------------------------------
// elementwise / layout kernels
// ---------------------------------------------------------------------------------------------------------------
// fp32 [R][C] -> bf16 [R][ldo] (plain) ; optional transposed copy bf16 [C][ldt]
__global__ void __launch_bounds__(256) cast_transpose_kernel(const float* __restrict__ src, int R, int C, __nv_bfloat16* __restrict__ dst,
                                                             int ldo, __nv_bfloat16* __restrict__ dstT, int ldt) {
    pdl_grid_sync();
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        float v = 0.f;
        if (r < R && c < C) v = src[(long long)r * C + c];
        tile[i][tx] = v;
        if (dst && r < R && c < C) dst[(long long)r * ldo + c] = __float2bfloat16_rn(v);
    }
    if (!dstT) return;
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (r < R && c < C) dstT[(long long)c * ldt + r] = __float2bfloat16_rn(tile[tx][i]);
    }
}

// bf16 [R][lds] (first C columns) -> bf16 [C][ldt] (columns >= R of the destination are left untouched: zero padding)
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, int R, int C, int lds,
                                                             __nv_bfloat16* __restrict__ dstT, int ldt) {
    pdl_grid_sync();
    __shared__ __nv_bfloat16 tile[32][34];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < R && c < C) ? src[(long long)r * lds + c] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (r < R && c < C) dstT[(long long)c * ldt + r] = tile[tx][i];
    }
}

// fp32 [rows][C] -> bf16 [rows][C]
__global__ void __launch_bounds__(256) cast_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    pdl_grid_sync();
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * 1024) {
        float v[4];
        load4(src + i, v);
        store4(dst + i, v);
    }
}

// x (fp32) += y (bf16)
__global__ void __launch_bounds__(256) add_bf16_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ y, long long n) {
    pdl_grid_sync();
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * 1024) {
        float a[4], b[4];
        load4(x + i, a);
        load4(y + i, b);
#pragma unroll
        for (int e = 0; e < 4; ++e) a[e] += b[e];
        store4(x + i, a);
    }
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
// MODE 0: y = gelu(x);  MODE 1: y = y * gelu'(x)   (bf16 in / out)
template <int MODE>
__global__ void __launch_bounds__(256) gelu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
    pdl_grid_sync();
    for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 8; i < n; i += (long long)gridDim.x * 2048) {
        float a[8], b[8];
        load8(x + i, a);
        if (MODE == 1) load8(y + i, b);
#pragma unroll
        for (int e = 0; e < 8; ++e) b[e] = MODE == 0 ? gelu_f(a[e]) : b[e] * gelu_grad_f(a[e]);
        store8(y + i, b);
    }
}

// column sums of a [R][ld] matrix (first C columns), two stages, fixed order.  T = bf16 or float.
template <typename T>
__global__ void __launch_bounds__(256) vit_colsum_part_kernel(const T* __restrict__ src, int R, int C, int ld, int rows_per_chunk,
                                                          float* __restrict__ part) {
    pdl_grid_sync();
    __shared__ float sh[8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = r0 + rows_per_chunk < R ? r0 + rows_per_chunk : R;
    float s = 0.f;
    if (c < C)
        for (int r = r0 + w; r < r1; r += 8) s += to_f(src[(long long)r * ld + c]);
    sh[w][threadIdx.x & 31] = s;
    __syncthreads();
    if (w == 0 && c < C) {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) a += sh[k][threadIdx.x];
        part[(long long)blockIdx.y * C + c] = a;
    }
}
__global__ void __launch_bounds__(256) vit_colsum_final_kernel(const float* __restrict__ part, int chunks, int C, float* __restrict__ out) {
    pdl_grid_sync();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int k = 0; k < chunks; ++k) s += (double)part[(long long)k * C + c];
    out[c] = (float)s;
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm (warp per row; E % 32 == 0, E <= 2048)
// ---------------------------------------------------------------------------------------------------------------
constexpr int LN_MAXPER = 32;   // E / 32 (E <= 1024)

__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, long long row_stride, const float* __restrict__ g,
                                                     const float* __restrict__ b, __nv_bfloat16* __restrict__ y, long long y_stride,
                                                     float* __restrict__ yf, float* __restrict__ stats, int rows, int E, float eps) {
    pdl_grid_sync();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (long long)row * row_stride;
    float v[LN_MAXPER];
    const int per = E / 32;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXPER; ++i)
        if (i < per) { v[i] = xr[lane + 32 * i]; s += v[i]; }
    const float mean = warp_sum(s) / (float)E;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXPER; ++i)
        if (i < per) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / (float)E + eps);
    if (lane == 0 && stats) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
#pragma unroll
    for (int i = 0; i < LN_MAXPER; ++i) {
        if (i >= per) break;
        const int c = lane + 32 * i;
        const float o = (v[i] - mean) * rstd * g[c] + b[c];
        if (y) y[(long long)row * y_stride + c] = __float2bfloat16_rn(o);
        if (yf) yf[(long long)row * E + c] = o;
    }
}

// dx (+)= rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)); per-row-block partials of dgamma / dbeta.  DY = bf16 or float.
template <typename DY>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const DY* __restrict__ dy, long long dy_stride, const float* __restrict__ x,
                                                     long long row_stride, const float* __restrict__ stats, const float* __restrict__ g,
                                                     float* __restrict__ dx, long long dx_stride, int accumulate, int rows, int E,
                                                     float* __restrict__ part /* [blocks][2][E] */) {
    pdl_grid_sync();
    extern __shared__ float sh[];   // [8][2][E]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + w;
    const int per = E / 32;
    float* mine = sh + (size_t)w * 2 * E;
    if (row < rows) {
        const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
        const DY* dr = dy + (long long)row * dy_stride;
        const float* xr = x + (long long)row * row_stride;
        float dg[LN_MAXPER], xh[LN_MAXPER];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXPER; ++i) {
            if (i >= per) break;
            const int c = lane + 32 * i;
            const float d = to_f(dr[c]);
            xh[i] = (xr[c] - mean) * rstd;
            dg[i] = d * g[c];
            s1 += dg[i];
            s2 += dg[i] * xh[i];
            mine[c] = d * xh[i];          // dgamma contribution
            mine[E + c] = d;              // dbeta contribution
        }
        s1 = warp_sum(s1) / (float)E;
        s2 = warp_sum(s2) / (float)E;
        float* dxr = dx + (long long)row * dx_stride;
#pragma unroll
        for (int i = 0; i < LN_MAXPER; ++i) {
            if (i >= per) break;
            const int c = lane + 32 * i;
            const float o = rstd * (dg[i] - s1 - xh[i] * s2);
            dxr[c] = accumulate ? dxr[c] + o : o;
        }
    } else {
        for (int c = lane; c < 2 * E; c += 32) mine[c] = 0.f;
    }
    __syncthreads();
    if (part)
        for (int c = threadIdx.x; c < 2 * E; c += 256) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) a += sh[(size_t)k * 2 * E + c];
            part[(long long)blockIdx.x * 2 * E + c] = a;
        }
}
// dgamma[c] = sum_blocks part[b][0][c]; dbeta[c] = sum_blocks part[b][1][c]
__global__ void __launch_bounds__(256) ln_bwd_final_kernel(const float* __restrict__ part, int blocks, int E, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta) {
    pdl_grid_sync();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= 2 * E) return;
    double s = 0.0;
    for (int k = 0; k < blocks; ++k) s += (double)part[(long long)k * 2 * E + c];
    if (c < E) dgamma[c] = (float)s; else dbeta[c - E] = (float)s;
}

// ---------------------------------------------------------------------------------------------------------------
// patches <-> NDHWC volume
// ---------------------------------------------------------------------------------------------------------------
struct PatchGeom { int B, C, D, H, W, pitch, p, gd, gh, gw, Kp; };

// P[token][c * p^3 + (pd * p + ph) * p + pw] = x[b][gd_*p+pd][gh_*p+ph][gw_*p+pw][c]   (PyTorch Conv3d weight order)
// block = (token, pd): tile [p (ph)][p (pw)][C] through shared memory so that reads are 2C-byte runs and writes 2p-byte runs
__global__ void __launch_bounds__(256) patchify_kernel(PatchGeom g, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ P) {
    pdl_grid_sync();
    extern __shared__ __nv_bfloat16 tile[];   // [p*p][C + 2]
    const int token = blockIdx.x, pd = blockIdx.y;
    const int np = g.gd * g.gh * g.gw;
    const int b = token / np;
    int t = token % np;
    const int gw_ = t % g.gw; t /= g.gw;
    const int gh_ = t % g.gh; const int gd_ = t / g.gh;
    const int p = g.p, CP = g.C + 2;
    const __nv_bfloat16* base = x + ((((long long)b * g.D + gd_ * p + pd) * g.H + gh_ * p) * g.W + gw_ * p) * g.pitch;
    for (int i = threadIdx.x; i < p * p * g.C; i += 256) {
        const int c = i % g.C, v = i / g.C, pw = v % p, ph = v / p;
        tile[v * CP + c] = base[((long long)ph * g.W + pw) * g.pitch + c];
    }
    __syncthreads();
    __nv_bfloat16* out = P + (long long)token * g.Kp + (long long)pd * p * p;
    for (int i = threadIdx.x; i < p * p * g.C; i += 256) {
        const int v = i % (p * p), c = i / (p * p);
        out[(long long)c * p * p * p + v] = tile[v * CP + c];
    }
}
// dx[b][..][c] += dP[token][...]   (each voxel belongs to exactly one patch: no collisions)
__global__ void __launch_bounds__(256) unpatchify_add_kernel(PatchGeom g, const __nv_bfloat16* __restrict__ dP, __nv_bfloat16* __restrict__ dx) {
    pdl_grid_sync();
    extern __shared__ __nv_bfloat16 tile[];
    const int token = blockIdx.x, pd = blockIdx.y;
    const int np = g.gd * g.gh * g.gw;
    const int b = token / np;
    int t = token % np;
    const int gw_ = t % g.gw; t /= g.gw;
    const int gh_ = t % g.gh; const int gd_ = t / g.gh;
    const int p = g.p, CP = g.C + 2;
    const __nv_bfloat16* in = dP + (long long)token * g.Kp + (long long)pd * p * p;
    for (int i = threadIdx.x; i < p * p * g.C; i += 256) {
        const int v = i % (p * p), c = i / (p * p);
        tile[v * CP + c] = in[(long long)c * p * p * p + v];
    }
    __syncthreads();
    __nv_bfloat16* base = dx + ((((long long)b * g.D + gd_ * p + pd) * g.H + gh_ * p) * g.W + gw_ * p) * g.pitch;
    for (int i = threadIdx.x; i < p * p * g.C; i += 256) {
        const int c = i % g.C, v = i / g.C, pw = v % p, ph = v / p;
        __nv_bfloat16* q = base + ((long long)ph * g.W + pw) * g.pitch + c;
        *q = __float2bfloat16_rn(__bfloat162float(*q) + __bfloat162float(tile[v * CP + c]));
    }
}

// X[b][0] = cls + pos[0];  X[b][1+i] = tok[b*np + i] + pos[1+i]      (vision_transformer.py:423-428)
__global__ void __launch_bounds__(256) assemble_tokens_kernel(const __nv_bfloat16* __restrict__ tok, const float* __restrict__ cls,
                                                              const float* __restrict__ pos, float* __restrict__ X, int B, int np, int E) {
    pdl_grid_sync();
    const long long n = (long long)B * (np + 1) * E;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const int e = (int)(i % E);
        const long long r = i / E;
        const int t = (int)(r % (np + 1)), b = (int)(r / (np + 1));
        const float base = t == 0 ? cls[e] : __bfloat162float(tok[((long long)b * np + t - 1) * E + e]);
        X[i] = base + pos[(long long)t * E + e];
    }
}
// dcls[e] = sum_b dX[b][0][e]; dpos[t][e] = sum_b dX[b][t][e]; dtok[b*np+i][e] = bf16(dX[b][1+i][e])
__global__ void __launch_bounds__(256) assemble_tokens_bwd_kernel(const float* __restrict__ dX, float* __restrict__ dcls, float* __restrict__ dpos,
                                                                  __nv_bfloat16* __restrict__ dtok, int B, int np, int E) {
    pdl_grid_sync();
    const long long n = (long long)(np + 1) * E;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const int e = (int)(i % E), t = (int)(i / E);
        float s = 0.f;
        for (int b = 0; b < B; ++b) {
            const float v = dX[((long long)b * (np + 1) + t) * E + e];
            s += v;
            if (t > 0) dtok[((long long)b * np + t - 1) * E + e] = __float2bfloat16_rn(v);
        }
        dpos[i] = s;
        if (t == 0) dcls[e] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// attention (head dim 64), qkv: [B*T][3E] bf16 with column = which * E + h * 64 + d
//
// Flash-style warp-MMA kernels (mma.sync.m16n8k16 bf16 -> fp32; 1.2 GFLOP per pass at cfg4 is far too small for a tcgen05
// pipeline to pay for its set-up): a CTA of 4 warps owns 64 rows (queries, or keys in the dK / dV kernel), a warp 16 of them.
// The other side streams through shared memory in blocks of 64 rows, double-buffered with cp.async; row pitch 144 B makes
// every ldmatrix phase conflict-free.  Forward: online softmax in fp32 (exp2 with the scale folded in), P rounded to bf16
// for the P V product, log-sum-exp saved.  Backward: recomputation from the saved log-sum-exp; dQ in one kernel (per query
// block, over all keys) and dK / dV in another (per key block, over all queries), so that no gradient needs an atomic:
// every output element is produced by ONE thread in a fixed order (bit-reproducible).
// ---------------------------------------------------------------------------------------------------------------
constexpr int FA_THREADS = 128;  // 4 warps x 16 rows
constexpr int FA_ROWS = 64;
constexpr int FA_LD = 72;        // bf16 elements per shared-memory row (64 dims + 8 pad = 144 B)
constexpr int FA_TILE = FA_ROWS * FA_LD;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const __nv_bfloat16* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// stage rows [r0, r0 + 64) of a 64-column slice (row pitch ld elements) into a shared tile; rows >= T repeat row T - 1 (finite
// values; the consumers mask them)
__device__ __forceinline__ void fa_stage(const __nv_bfloat16* __restrict__ src, long long ld, int r0, int T, __nv_bfloat16* tile) {
    for (int i = threadIdx.x; i < FA_ROWS * 8; i += FA_THREADS) {
        const int r = i >> 3, c8 = (i & 7) * 8;
        const int gr = min(r0 + r, T - 1);
        cp_async16(tile + r * FA_LD + c8, src + (long long)gr * ld + c8);
    }
}
// A fragments (16 rows x 64 dims) of a warp's rows [w16, w16 + 16) of a tile
__device__ __forceinline__ void fa_load_a(uint32_t (&a)[4][4], const __nv_bfloat16* tile, int w16, int lane) {
    const __nv_bfloat16* p = tile + (w16 + (lane & 7) + ((lane >> 3) & 1) * 8) * FA_LD + (lane >> 4) * 8;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldsm4(a[kk], p + kk * 16);
}
// acc[16 x 64] = A (16 x 64 dims) . tile^T   (tile: 64 rows x 64 dims; acc column = tile row)
__device__ __forceinline__ void fa_mma_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], const __nv_bfloat16* tile, int lane) {
    const __nv_bfloat16* p = tile + ((lane & 7) + (lane >> 4) * 8) * FA_LD + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int nb2 = 0; nb2 < 4; ++nb2)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t b[4];
            ldsm4(b, p + nb2 * 16 * FA_LD + kk * 16);
            mma_bf16(acc[2 * nb2], a[kk], b[0], b[1]);
            mma_bf16(acc[2 * nb2 + 1], a[kk], b[2], b[3]);
        }
}
// acc[16 x 64 dims] += P (16 x 64 rows of the tile, A fragments) . tile   (tile: 64 rows x 64 dims)
__device__ __forceinline__ void fa_mma_nn(float (&acc)[8][4], const uint32_t (&pa)[4][4], const __nv_bfloat16* tile, int lane) {
    const __nv_bfloat16* p = tile + ((lane & 7) + ((lane >> 3) & 1) * 8) * FA_LD + (lane >> 4) * 8;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int nb2 = 0; nb2 < 4; ++nb2) {
            uint32_t b[4];
            ldsm4t(b, p + kk * 16 * FA_LD + nb2 * 16);
            mma_bf16(acc[2 * nb2], pa[kk], b[0], b[1]);
            mma_bf16(acc[2 * nb2 + 1], pa[kk], b[2], b[3]);
        }
}
// accumulator tile (16 x 64, fp32) -> A fragments (bf16)
__device__ __forceinline__ void fa_acc_to_a(uint32_t (&pa)[4][4], const float (&s)[8][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        pa[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pa[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pa[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    }
}
// write a warp's 16 x 64 accumulator (times `mul`) as bf16 rows (row pitch ld) for rows < T
__device__ __forceinline__ void fa_store(const float (&acc)[8][4], float mul0, float mul1, __nv_bfloat16* __restrict__ dst, long long ld,
                                         int row0, int T, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        if (row0 + g < T)
            *reinterpret_cast<uint32_t*>(dst + (long long)(row0 + g) * ld + nb * 8 + 2 * t) = pack_bf16(acc[nb][0] * mul0, acc[nb][1] * mul0);
        if (row0 + g + 8 < T)
            *reinterpret_cast<uint32_t*>(dst + (long long)(row0 + g + 8) * ld + nb * 8 + 2 * t) = pack_bf16(acc[nb][2] * mul1, acc[nb][3] * mul1);
    }
}

// out[b*T + i][h*64 + d] = sum_j softmax_j(scale q_i.k_j) v_j[d];  lse[(b*H + h)*T + i] = log sum_j exp(scale q_i.k_j)
// LSA (vision_transformer.py:90-135): `scale_h` (nullable) = learnable temperature per head replacing `scale`; `diag`: the score of
// token i with itself is masked out before the softmax for i < diag (the reference builds its mask from the 2-D patch count)
__global__ void __launch_bounds__(FA_THREADS) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                              float* __restrict__ lse, int B, int H, int T, float scale,
                                                              const float* __restrict__ scale_h, int diag) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t smraw[];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smraw);     // [64][FA_LD]
    __nv_bfloat16* Ks = Qs + FA_TILE;                                // [2][64][FA_LD]
    __nv_bfloat16* Vs = Ks + 2 * FA_TILE;                            // [2][64][FA_LD]
    const int E = H * 64, bh = blockIdx.x, b = bh / H, h = bh % H;
    const long long ld = 3LL * E;
    const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * 64;
    const int q0 = blockIdx.y * FA_ROWS, nblk = (T + FA_ROWS - 1) / FA_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    if (scale_h) scale = scale_h[h];
    const int qi0 = q0 + warp * 16 + g, qi1 = qi0 + 8;      // this thread's two query rows
    fa_stage(base, ld, q0, T, Qs);
    fa_stage(base + E, ld, 0, T, Ks);
    fa_stage(base + 2 * E, ld, 0, T, Vs);
    cp_async_commit();
    uint32_t aq[4][4];
    float o[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) o[nb][0] = o[nb][1] = o[nb][2] = o[nb][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const float sl2 = LOG2E;          // the scores are scaled right after the MMA (a learnable LSA temperature may be negative)
    for (int jb = 0; jb < nblk; ++jb) {
        const int cur = jb & 1;
        if (jb + 1 < nblk) {
            fa_stage(base + E, ld, (jb + 1) * FA_ROWS, T, Ks + (cur ^ 1) * FA_TILE);
            fa_stage(base + 2 * E, ld, (jb + 1) * FA_ROWS, T, Vs + (cur ^ 1) * FA_TILE);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (jb == 0) fa_load_a(aq, Qs, warp * 16, lane);
        float s[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
        fa_mma_nt(s, aq, Ks + cur * FA_TILE, lane);
        float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] *= scale; s[nb][1] *= scale; s[nb][2] *= scale; s[nb][3] *= scale;
            const int j = jb * FA_ROWS + nb * 8 + 2 * t;
            if (j >= T) s[nb][0] = s[nb][2] = -INFINITY;
            if (j + 1 >= T) s[nb][1] = s[nb][3] = -INFINITY;
            if (diag) {      // only the first `diag` tokens mask themselves (see b2_vit_desc.lsa_mask)
                if (j == qi0 && j < diag) s[nb][0] = -INFINITY;
                if (j + 1 == qi0 && j + 1 < diag) s[nb][1] = -INFINITY;
                if (j == qi1 && j < diag) s[nb][2] = -INFINITY;
                if (j + 1 == qi1 && j + 1 < diag) s[nb][3] = -INFINITY;
            }
            bm0 = fmaxf(bm0, fmaxf(s[nb][0], s[nb][1]));
            bm1 = fmaxf(bm1, fmaxf(s[nb][2], s[nb][3]));
        }
        bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
        // every key block holds at least one valid key; with the diagonal mask a block whose only valid key is the row's own
        // token (T % 64 == 1) is empty for that row: keep the running maximum finite so exp2(-inf - -inf) never appears
        float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
        if (n0 == -INFINITY) n0 = 0.f;
        if (n1 == -INFINITY) n1 = 0.f;
        const float al0 = exp2f((m0 - n0) * sl2), al1 = exp2f((m1 - n1) * sl2);
        m0 = n0; m1 = n1;
        float r0 = 0.f, r1 = 0.f;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = exp2f((s[nb][0] - n0) * sl2); s[nb][1] = exp2f((s[nb][1] - n0) * sl2);
            s[nb][2] = exp2f((s[nb][2] - n1) * sl2); s[nb][3] = exp2f((s[nb][3] - n1) * sl2);
            r0 += s[nb][0] + s[nb][1];
            r1 += s[nb][2] + s[nb][3];
            o[nb][0] *= al0; o[nb][1] *= al0; o[nb][2] *= al1; o[nb][3] *= al1;
        }
        l0 = l0 * al0 + r0;
        l1 = l1 * al1 + r1;
        uint32_t pa[4][4];
        fa_acc_to_a(pa, s);
        fa_mma_nn(o, pa, Vs + cur * FA_TILE, lane);
        __syncthreads();
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const int row0 = q0 + warp * 16;
    fa_store(o, 1.f / l0, 1.f / l1, out + (long long)b * T * E + h * 64, E, row0, T, lane);
    if (t == 0) {
        if (row0 + g < T) lse[(long long)bh * T + row0 + g] = m0 + logf(l0);
        if (row0 + g + 8 < T) lse[(long long)bh * T + row0 + g + 8] = m1 + logf(l1);
    }
}

// dQ (and the row terms D_i = dO_i . O_i):  dS_ij = p_ij (dO_i . v_j - D_i);  dq_i = scale sum_j dS_ij k_j
__global__ void __launch_bounds__(FA_THREADS) attn_bwd_q_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ O,
                                                                const __nv_bfloat16* __restrict__ dO, const float* __restrict__ lse,
                                                                __nv_bfloat16* __restrict__ dqkv, float* __restrict__ Drow, int B, int H,
                                                                int T, float scale, const float* __restrict__ scale_h, int diag,
                                                                float* __restrict__ dscale_part /* [B*H][gridDim.y], nullable */) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t smraw[];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smraw);     // [64][FA_LD]
    __nv_bfloat16* Gs = Qs + FA_TILE;                                // dO [64][FA_LD]
    __nv_bfloat16* Ks = Gs + FA_TILE;                                // [2][64][FA_LD]
    __nv_bfloat16* Vs = Ks + 2 * FA_TILE;                            // [2][64][FA_LD]
    float* Ds = reinterpret_cast<float*>(Vs + 2 * FA_TILE);          // [64]
    const int E = H * 64, bh = blockIdx.x, b = bh / H, h = bh % H;
    const long long ld = 3LL * E;
    const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * 64;
    const __nv_bfloat16* gbase = dO + (long long)b * T * E + h * 64;
    const __nv_bfloat16* obase = O + (long long)b * T * E + h * 64;
    const int q0 = blockIdx.y * FA_ROWS, nblk = (T + FA_ROWS - 1) / FA_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    if (scale_h) scale = scale_h[h];
    const int qi0 = q0 + warp * 16 + g, qi1 = qi0 + 8;
    float dsc = 0.f;                                                  // sum of dS_ij (q_i . k_j) over this thread's entries
    fa_stage(base, ld, q0, T, Qs);
    fa_stage(gbase, E, q0, T, Gs);
    fa_stage(base + E, ld, 0, T, Ks);
    fa_stage(base + 2 * E, ld, 0, T, Vs);
    cp_async_commit();
    {   // D_i = dO_i . O_i: two threads per row, 32 dims each, fixed order
        const int r = threadIdx.x >> 1, half = threadIdx.x & 1, gr = min(q0 + r, T - 1);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 a = *reinterpret_cast<const uint4*>(gbase + (long long)gr * E + half * 32 + c * 8);
            const uint4 o4 = *reinterpret_cast<const uint4*>(obase + (long long)gr * E + half * 32 + c * 8);
            const __nv_bfloat162* ap = reinterpret_cast<const __nv_bfloat162*>(&a);
            const __nv_bfloat162* op = reinterpret_cast<const __nv_bfloat162*>(&o4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 x = __bfloat1622float2(ap[e]), y = __bfloat1622float2(op[e]);
                acc = fmaf(x.x, y.x, acc);
                acc = fmaf(x.y, y.y, acc);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (half == 0) {
            Ds[r] = acc;
            if (q0 + r < T) Drow[(long long)bh * T + q0 + r] = acc;
        }
    }
    const int row0 = q0 + warp * 16;
    const float L0 = lse[(long long)bh * T + min(row0 + g, T - 1)] * LOG2E, L1 = lse[(long long)bh * T + min(row0 + g + 8, T - 1)] * LOG2E;
    uint32_t aq[4][4], ag[4][4];
    float dq[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) dq[nb][0] = dq[nb][1] = dq[nb][2] = dq[nb][3] = 0.f;
    const float sl2 = scale * LOG2E;
    float D0 = 0.f, D1 = 0.f;
    for (int jb = 0; jb < nblk; ++jb) {
        const int cur = jb & 1;
        if (jb + 1 < nblk) {
            fa_stage(base + E, ld, (jb + 1) * FA_ROWS, T, Ks + (cur ^ 1) * FA_TILE);
            fa_stage(base + 2 * E, ld, (jb + 1) * FA_ROWS, T, Vs + (cur ^ 1) * FA_TILE);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (jb == 0) {
            fa_load_a(aq, Qs, warp * 16, lane);
            fa_load_a(ag, Gs, warp * 16, lane);
            D0 = Ds[warp * 16 + g];
            D1 = Ds[warp * 16 + g + 8];
        }
        float s[8][4], dp[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
            dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f;
        }
        fa_mma_nt(s, aq, Ks + cur * FA_TILE, lane);
        fa_mma_nt(dp, ag, Vs + cur * FA_TILE, lane);
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int j = jb * FA_ROWS + nb * 8 + 2 * t;
            const bool v0 = j < T, v1 = j + 1 < T;
            const bool a0 = v0 && !(j < diag && j == qi0), a1 = v1 && !(j + 1 < diag && j + 1 == qi0);
            const bool a2 = v0 && !(j < diag && j == qi1), a3 = v1 && !(j + 1 < diag && j + 1 == qi1);
            const float r0 = s[nb][0], r1 = s[nb][1], r2 = s[nb][2], r3 = s[nb][3];       // raw q . k
            s[nb][0] = a0 ? exp2f(r0 * sl2 - L0) * (dp[nb][0] - D0) : 0.f;
            s[nb][1] = a1 ? exp2f(r1 * sl2 - L0) * (dp[nb][1] - D0) : 0.f;
            s[nb][2] = a2 ? exp2f(r2 * sl2 - L1) * (dp[nb][2] - D1) : 0.f;
            s[nb][3] = a3 ? exp2f(r3 * sl2 - L1) * (dp[nb][3] - D1) : 0.f;
            if (dscale_part) {      // rows >= T are clamped copies of row T - 1: they must not count
                if (qi0 < T) dsc += s[nb][0] * r0 + s[nb][1] * r1;
                if (qi1 < T) dsc += s[nb][2] * r2 + s[nb][3] * r3;
            }
        }
        uint32_t pa[4][4];
        fa_acc_to_a(pa, s);
        fa_mma_nn(dq, pa, Ks + cur * FA_TILE, lane);
        __syncthreads();
    }
    fa_store(dq, scale, scale, dqkv + (long long)b * T * ld + h * 64, ld, row0, T, lane);
    if (dscale_part) {              // fixed-order CTA reduction: lanes by shuffle tree, the four warps in order
        dsc = warp_sum(dsc);
        __syncthreads();
        if (lane == 0) Ds[warp] = dsc;
        __syncthreads();
        if (threadIdx.x == 0) dscale_part[(long long)bh * gridDim.y + blockIdx.y] = (Ds[0] + Ds[1]) + (Ds[2] + Ds[3]);
    }
}

// dscale[h] = sum over batch and query blocks of the per-CTA partials (fixed order)
__global__ void __launch_bounds__(32) attn_dscale_final_kernel(const float* __restrict__ part, int B, int H, int nblk, float* __restrict__ dscale) {
    pdl_grid_sync();
    const int h = blockIdx.x * 32 + threadIdx.x;
    if (h >= H) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < nblk; ++k) s += part[((long long)b * H + h) * nblk + k];
    dscale[h] = s;
}

// dK, dV:  dv_j = sum_i p_ij dO_i;  dk_j = scale sum_i dS_ij q_i      (a CTA owns 64 keys and streams the queries / dO)
__global__ void __launch_bounds__(FA_THREADS) attn_bwd_kv_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dO,
                                                                 const float* __restrict__ lse, const float* __restrict__ Drow,
                                                                 __nv_bfloat16* __restrict__ dqkv, int B, int H, int T, float scale,
                                                                 const float* __restrict__ scale_h, int diag) {
    pdl_grid_sync();
    extern __shared__ __align__(16) uint8_t smraw[];
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smraw);     // [64][FA_LD]
    __nv_bfloat16* Vs = Ks + FA_TILE;                                // [64][FA_LD]
    __nv_bfloat16* Qs = Vs + FA_TILE;                                // [2][64][FA_LD]
    __nv_bfloat16* Gs = Qs + 2 * FA_TILE;                            // dO [2][64][FA_LD]
    float* Ls = reinterpret_cast<float*>(Gs + 2 * FA_TILE);          // [2][64] lse * log2(e)
    float* Ds = Ls + 2 * FA_ROWS;                                    // [2][64]
    const int E = H * 64, bh = blockIdx.x, b = bh / H, h = bh % H;
    const long long ld = 3LL * E;
    const __nv_bfloat16* base = qkv + (long long)b * T * ld + h * 64;
    const __nv_bfloat16* gbase = dO + (long long)b * T * E + h * 64;
    const int k0 = blockIdx.y * FA_ROWS, nblk = (T + FA_ROWS - 1) / FA_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3;
    if (scale_h) scale = scale_h[h];
    const int kj0 = k0 + warp * 16 + (lane >> 2), kj1 = kj0 + 8;     // this thread's two key rows
    auto stage_rows = [&](int ib, int buf) {
        fa_stage(base, ld, ib * FA_ROWS, T, Qs + buf * FA_TILE);
        fa_stage(gbase, E, ib * FA_ROWS, T, Gs + buf * FA_TILE);
        if (threadIdx.x < FA_ROWS) {
            const int i = min(ib * FA_ROWS + (int)threadIdx.x, T - 1);
            Ls[buf * FA_ROWS + threadIdx.x] = lse[(long long)bh * T + i] * LOG2E;
            Ds[buf * FA_ROWS + threadIdx.x] = Drow[(long long)bh * T + i];
        }
    };
    fa_stage(base + E, ld, k0, T, Ks);
    fa_stage(base + 2 * E, ld, k0, T, Vs);
    stage_rows(0, 0);
    cp_async_commit();
    uint32_t ak[4][4], av[4][4];
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        dk[nb][0] = dk[nb][1] = dk[nb][2] = dk[nb][3] = 0.f;
        dv[nb][0] = dv[nb][1] = dv[nb][2] = dv[nb][3] = 0.f;
    }
    const float sl2 = scale * LOG2E;
    for (int ib = 0; ib < nblk; ++ib) {
        const int cur = ib & 1;
        if (ib + 1 < nblk) {
            stage_rows(ib + 1, cur ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (ib == 0) {
            fa_load_a(ak, Ks, warp * 16, lane);
            fa_load_a(av, Vs, warp * 16, lane);
        }
        // transposed scores: rows = my keys, columns = the block's queries
        float s[8][4], dp[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
            dp[nb][0] = dp[nb][1] = dp[nb][2] = dp[nb][3] = 0.f;
        }
        fa_mma_nt(s, ak, Qs + cur * FA_TILE, lane);
        fa_mma_nt(dp, av, Gs + cur * FA_TILE, lane);
        const float* Lc = Ls + cur * FA_ROWS;
        const float* Dc = Ds + cur * FA_ROWS;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const int c = nb * 8 + 2 * t, i = ib * FA_ROWS + c;
            const bool v0 = i < T, v1 = i + 1 < T;
            const float La = Lc[c], Lb = Lc[c + 1], Da = Dc[c], Db = Dc[c + 1];
            const bool a0 = v0 && !(i < diag && i == kj0), a1 = v1 && !(i + 1 < diag && i + 1 == kj0);
            const bool a2 = v0 && !(i < diag && i == kj1), a3 = v1 && !(i + 1 < diag && i + 1 == kj1);
            const float p0 = a0 ? exp2f(s[nb][0] * sl2 - La) : 0.f, p1 = a1 ? exp2f(s[nb][1] * sl2 - Lb) : 0.f;
            const float p2 = a2 ? exp2f(s[nb][2] * sl2 - La) : 0.f, p3 = a3 ? exp2f(s[nb][3] * sl2 - Lb) : 0.f;
            s[nb][0] = p0; s[nb][1] = p1; s[nb][2] = p2; s[nb][3] = p3;
            dp[nb][0] = p0 * (dp[nb][0] - Da); dp[nb][1] = p1 * (dp[nb][1] - Db);
            dp[nb][2] = p2 * (dp[nb][2] - Da); dp[nb][3] = p3 * (dp[nb][3] - Db);
        }
        uint32_t pa[4][4];
        fa_acc_to_a(pa, s);
        fa_mma_nn(dv, pa, Gs + cur * FA_TILE, lane);
        fa_acc_to_a(pa, dp);
        fa_mma_nn(dk, pa, Qs + cur * FA_TILE, lane);
        __syncthreads();
    }
    const int row0 = k0 + warp * 16;
    __nv_bfloat16* ob = dqkv + (long long)b * T * ld + h * 64;
    fa_store(dk, scale, scale, ob + E, ld, row0, T, lane);
    fa_store(dv, 1.f, 1.f, ob + 2 * E, ld, row0, T, lane);
}

// ---------------------------------------------------------------------------------------------------------------
// head: Linear(E -> F) on the B class tokens, output written into / gradient read from the NDHWC bottleneck buffers
// ---------------------------------------------------------------------------------------------------------------
constexpr int HEAD_MAXB = 4;
struct HeadGeom { int B, E, F, oc, od, oh, ow, pitch; };
__device__ __forceinline__ long long head_addr(const HeadGeom& g, int b, int j) {   // j = ((c*od + z)*oh + y)*ow + x  (x.reshape(size))
    const int x = j % g.ow; int r = j / g.ow;
    const int y = r % g.oh; r /= g.oh;
    const int z = r % g.od; const int c = r / g.od;
    return ((((long long)b * g.od + z) * g.oh + y) * g.ow + x) * g.pitch + c;
}
template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(HeadGeom g, const float* __restrict__ xn /* [B][E] */, const float* __restrict__ Wh,
                                                       const float* __restrict__ bh, T* __restrict__ out) {
    pdl_grid_sync();
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= g.F) return;
    const float* wr = Wh + (long long)j * g.E;
    float acc[HEAD_MAXB];
#pragma unroll
    for (int b = 0; b < HEAD_MAXB; ++b) acc[b] = 0.f;
    for (int k = lane; k < g.E; k += 32) {
        const float wv = wr[k];
#pragma unroll
        for (int b = 0; b < HEAD_MAXB; ++b)
            if (b < g.B) acc[b] = fmaf(wv, xn[b * g.E + k], acc[b]);
    }
#pragma unroll
    for (int b = 0; b < HEAD_MAXB; ++b)
        if (b < g.B) {
            const float s = warp_sum(acc[b]);
            if (lane == 0) out[head_addr(g, b, j)] = from_f<T>(s + bh[j]);
        }
}
// dW[j][k] = sum_b dout[b][j] xn[b][k]; db[j] = sum_b dout[b][j]; partial dxn[chunk][b][k] = sum_{j in chunk} dout[b][j] W[j][k]
template <typename T>
__global__ void __launch_bounds__(256) head_bwd_kernel(HeadGeom g, const T* __restrict__ dout, const float* __restrict__ xn,
                                                       const float* __restrict__ Wh, float* __restrict__ dW, float* __restrict__ db,
                                                       float* __restrict__ part, int rows_per_block) {
    pdl_grid_sync();
    __shared__ float dsh[HEAD_MAXB][64];
    const int j0 = blockIdx.x * rows_per_block;
    float acc[HEAD_MAXB][4], xr[HEAD_MAXB][4];
    const int kper = (g.E + 255) / 256;   // <= 4 (E <= 1024) -- host checks
#pragma unroll
    for (int b = 0; b < HEAD_MAXB; ++b)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            acc[b][u] = 0.f;
            const int k = threadIdx.x + 256 * u;
            xr[b][u] = (b < g.B && u < kper && k < g.E) ? xn[b * g.E + k] : 0.f;
        }
    for (int jb = 0; jb < rows_per_block; jb += 64) {
        __syncthreads();
        if (threadIdx.x < 64) {
            const int j = j0 + jb + threadIdx.x;
            float s = 0.f;
#pragma unroll
            for (int b = 0; b < HEAD_MAXB; ++b) {
                const float v = (b < g.B && j < g.F && jb + threadIdx.x < rows_per_block) ? to_f(dout[head_addr(g, b, j)]) : 0.f;
                dsh[b][threadIdx.x] = v;
                s += v;
            }
            if (j < g.F && jb + threadIdx.x < rows_per_block) db[j] = s;
        }
        __syncthreads();
        for (int r = 0; r < 64 && jb + r < rows_per_block && j0 + jb + r < g.F; ++r) {
            const long long j = j0 + jb + r;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = threadIdx.x + 256 * u;
                if (u < kper && k < g.E) {
                    const float wv = Wh[j * g.E + k];
                    float dwv = 0.f;
#pragma unroll
                    for (int b = 0; b < HEAD_MAXB; ++b) {
                        const float d = dsh[b][r];
                        acc[b][u] = fmaf(d, wv, acc[b][u]);
                        dwv = fmaf(d, xr[b][u], dwv);
                    }
                    dW[j * g.E + k] = dwv;
                }
            }
        }
    }
    for (int b = 0; b < g.B; ++b)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = threadIdx.x + 256 * u;
            if (u < kper && k < g.E) part[((long long)blockIdx.x * g.B + b) * g.E + k] = acc[b][u];
        }
}
__global__ void __launch_bounds__(256) head_bwd_final_kernel(const float* __restrict__ part, int blocks, int n, float* __restrict__ out) {
    pdl_grid_sync();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < blocks; ++k) s += (double)part[(long long)k * n + i];
    out[i] = (float)s;
}

}  // namespace b2

using namespace b2;

// ===============================================================================================================
// plan
// ===============================================================================================================
struct b2_vit_plan {
    b2_vit_desc d;
    int T, np, M, Mp, Mt, Mtp, Kp, Tp, E, H, F;
    // workspace offsets (bytes)
    size_t off_P, off_Pt, off_tok, off_X0, off_blocks, blk_bytes, off_clsn, off_lnf, off_scr, total;
    // per-block sub-offsets
    size_t b_Xin, b_st1, b_Xn, b_qkv, b_lse, b_O, b_Xmid, b_st2, b_Xn2, b_Hpre, b_Hact;
    // weight shadows (bf16): per block qkv, qkvT, proj, projT, fc1, fc1T, fc2, fc2T ; patch W, WT
    size_t off_w, w_blk_bytes, w_qkv, w_qkvT, w_proj, w_projT, w_fc1, w_fc1T, w_fc2, w_fc2T, off_wpe, off_wpeT;
    // scratch
    size_t s_tmp, s_tmp2, s_tA, s_tB, s_dX, s_part, s_gemm, gemm_scr_bytes, s_drow;
};

static size_t al(size_t v) { return (v + 255) / 256 * 256; }

extern "C" int b2_vit_plan_create(const b2_vit_desc* desc, b2_vit_plan** out) {
    B2_CHECK_ARG(desc && out);
    const b2_vit_desc& d = *desc;
    B2_CHECK_ARG(d.batch >= 1 && d.batch <= HEAD_MAXB && d.patch >= 1 && d.embed % 64 == 0 && d.embed <= 1024 && d.heads >= 1);
    B2_CHECK_ARG(d.embed == d.heads * 64 && d.depth >= 1 && d.mlp_ratio == 4 && d.in_channels % 8 == 0);
    B2_CHECK_ARG(d.D >= d.patch && d.H >= d.patch && d.W >= d.patch && d.out_features == d.out_c * d.out_d * d.out_h * d.out_w);
    b2_vit_plan* p = new (std::nothrow) b2_vit_plan();
    if (!p) return fail(B2_ENOMEM, "out of host memory%s", "");
    memset(p, 0, sizeof(*p));
    p->d = d;
    const int gd = d.D / d.patch, gh = d.H / d.patch, gw = d.W / d.patch;
    p->np = gd * gh * gw; p->T = p->np + 1;
    p->E = d.embed; p->H = d.heads; p->F = d.out_features;
    p->M = d.batch * p->T; p->Mp = (p->M + 127) / 128 * 128;
    p->Mt = d.batch * p->np; p->Mtp = (p->Mt + 127) / 128 * 128;
    p->Kp = d.in_channels * d.patch * d.patch * d.patch;
    p->Tp = (p->T + 31) / 32 * 32;
    if (p->Kp % 32 != 0) { delete p; return fail(B2_EUNSUPPORTED, "ViT: patch volume not a multiple of 32%s", ""); }
    const size_t E = p->E, Mp = p->Mp, M4 = 4 * E;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += al(bytes); return r; };
    p->off_P = take((size_t)p->Mtp * p->Kp * 2);
    p->off_Pt = take((size_t)p->Kp * p->Mtp * 2);
    p->off_tok = take((size_t)p->Mtp * E * 2);
    p->off_X0 = take(Mp * E * 4);
    {   // per block
        size_t b = 0;
        auto tk = [&](size_t bytes) { size_t r = b; b += al(bytes); return r; };
        p->b_Xin = tk(Mp * E * 4); p->b_st1 = tk(Mp * 2 * 4); p->b_Xn = tk(Mp * E * 2); p->b_qkv = tk(Mp * 3 * E * 2);
        p->b_lse = tk((size_t)d.batch * p->H * p->T * 4); p->b_O = tk(Mp * E * 2); p->b_Xmid = tk(Mp * E * 4); p->b_st2 = tk(Mp * 2 * 4);
        p->b_Xn2 = tk(Mp * E * 2); p->b_Hpre = tk(Mp * M4 * 2); p->b_Hact = tk(Mp * M4 * 2);
        p->blk_bytes = b;
    }
    p->off_blocks = take(p->blk_bytes * d.depth);
    p->off_clsn = take((size_t)HEAD_MAXB * E * 4);
    p->off_lnf = take((size_t)HEAD_MAXB * 2 * 4);
    {   // weight shadows
        size_t b = 0;
        auto tk = [&](size_t bytes) { size_t r = b; b += al(bytes); return r; };
        p->w_qkv = tk(3 * E * E * 2); p->w_qkvT = tk(3 * E * E * 2); p->w_proj = tk(E * E * 2); p->w_projT = tk(E * E * 2);
        p->w_fc1 = tk(M4 * E * 2); p->w_fc1T = tk(M4 * E * 2); p->w_fc2 = tk(M4 * E * 2); p->w_fc2T = tk(M4 * E * 2);
        p->w_blk_bytes = b;
    }
    p->off_w = take(p->w_blk_bytes * d.depth);
    p->off_wpe = take((size_t)E * p->Kp * 2);
    p->off_wpeT = take((size_t)p->Kp * E * 2);
    // scratch
    p->s_tmp = take(Mp * M4 * 2);                 // bf16 GEMM outputs / gradients wrt activations
    p->s_tmp2 = take(Mp * M4 * 2);
    p->s_tA = take(M4 * Mp * 2);                  // transposed bf16 operands of the weight-gradient GEMMs
    p->s_tB = take(M4 * Mp * 2);
    p->s_dX = take(Mp * E * 4);
    p->s_drow = take((size_t)d.batch * p->H * p->T * 4);
    const size_t part_floats = (size_t)((p->Mp + 7) / 8) * 2 * E + (size_t)64 * 3 * E * 4 + (size_t)((p->F + 255) / 256 + 1) * HEAD_MAXB * E + 1024;
    p->s_part = take(part_floats * 4);
    p->gemm_scr_bytes = (size_t)num_sms() * 2 * 128 * 256 * 4;
    p->s_gemm = take(p->gemm_scr_bytes);
    p->total = o;
    *out = p;
    return B2_OK;
}
extern "C" void b2_vit_plan_destroy(b2_vit_plan* p) { delete p; }
extern "C" size_t b2_vit_workspace_bytes(const b2_vit_plan* p) { return p ? p->total : 0; }
extern "C" size_t b2_vit_tokens_offset(const b2_vit_plan* p) { return p ? p->off_tok : 0; }
// LSA: one temperature vector [heads] per block appended after the 6 tail parameters; the qkv bias entries are null
extern "C" int b2_vit_num_params(const b2_vit_plan* p) { return p ? 2 + 12 * p->d.depth + 6 + (p->d.lsa ? p->d.depth : 0) : B2_EINVAL; }

namespace {
struct VP {   // parameter pointers in named_parameters() order
    const float* const* p; int depth;
    const float* cls() const { return p[0]; }
    const float* pos() const { return p[1]; }
    const float* lsa_scale(int l) const { return p[2 + 12 * depth + 6 + l]; }
    const float* blk(int l, int i) const { return p[2 + 12 * l + i]; }   // 0 n1w 1 n1b 2 qkvw 3 qkvb 4 projw 5 projb 6 n2w 7 n2b 8 fc1w 9 fc1b 10 fc2w 11 fc2b
    const float* normw() const { return p[2 + 12 * depth]; }
    const float* normb() const { return p[3 + 12 * depth]; }
    const float* pew() const { return p[4 + 12 * depth]; }
    const float* peb() const { return p[5 + 12 * depth]; }
    const float* headw() const { return p[6 + 12 * depth]; }
    const float* headb() const { return p[7 + 12 * depth]; }
};
template <typename T> T* at(void* ws, size_t off) { return reinterpret_cast<T*>((char*)ws + off); }

int cast_w(const float* w, int R, int C, __nv_bfloat16* dst, __nv_bfloat16* dstT, cudaStream_t st) {
    dim3 grid(cdiv(C, 32), cdiv(R, 32));
    B2_LAUNCH(cast_transpose_kernel, grid, 256, 0, st, w, R, C, dst, C, dstT, R);
    return B2_OK;
}
int transpose(const __nv_bfloat16* src, int R, int C, int lds, __nv_bfloat16* dstT, int ldt, cudaStream_t st) {
    dim3 grid(cdiv(C, 32), cdiv(R, 32));
    B2_LAUNCH(transpose_bf16_kernel, grid, 256, 0, st, src, R, C, lds, dstT, ldt);
    return B2_OK;
}
template <typename T>
int colsum(const T* src, int R, int C, int ld, float* part, float* out, cudaStream_t st) {
    const int chunks = R < 64 ? 1 : 64, rpc = cdiv(R, chunks);
    dim3 grid(cdiv(C, 32), chunks);
    B2_LAUNCH(vit_colsum_part_kernel<T>, grid, 256, 0, st, src, R, C, ld, rpc, part);
    B2_LAUNCH(vit_colsum_final_kernel, cdiv(C, 256), 256, 0, st, (const float*)part, chunks, C, out);
    return B2_OK;
}
int ln_fwd(const float* x, long long rs, const float* g, const float* b, __nv_bfloat16* y, long long ys, float* yf, float* stats, int rows,
           int E, float eps, cudaStream_t st) {
    B2_LAUNCH(ln_fwd_kernel, cdiv(rows, 8), 256, 0, st, x, rs, g, b, y, ys, yf, stats, rows, E, eps);
    return B2_OK;
}
template <typename DY>
int ln_bwd(const DY* dy, long long dys, const float* x, long long rs, const float* stats, const float* g, float* dx, long long dxs,
           int accumulate, int rows, int E, float* part, float* dgamma, float* dbeta, cudaStream_t st) {
    const int blocks = cdiv(rows, 8);
    static bool attr = false;
    if (!attr) {
        B2_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4));
        B2_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 2 * 1024 * 4));
        attr = true;
    }
    B2_LAUNCH(ln_bwd_kernel<DY>, blocks, 256, (size_t)8 * 2 * E * 4, st, dy, dys, x, rs, stats, g, dx, dxs, accumulate, rows, E, part);
    B2_LAUNCH(ln_bwd_final_kernel, cdiv(2 * E, 256), 256, 0, st, (const float*)part, blocks, E, dgamma, dbeta);
    return B2_OK;
}
long long grid_for(long long n, int per_thread) {
    long long g = (n / per_thread + 255) / 256, cap = (long long)num_sms() * 16;
    return g < 1 ? 1 : (g > cap ? cap : g);
}
}  // namespace

extern "C" int b2_vit_forward(b2_vit_plan* p, const float* const* params, const b2_act_view* skip0, void* ws, const b2_act_view* out_view,
                              float* out_dense, int keep_for_backward, b2_stream_t stream) {
    B2_CHECK_ARG(p && params && skip0 && ws && (out_view || out_dense));
    B2_CHECK_ARG(skip0->dtype == B2_BF16 && skip0->n == p->d.batch && skip0->c == p->d.in_channels && skip0->d == p->d.D && skip0->h == p->d.H &&
                 skip0->w == p->d.W);
    cudaStream_t st = (cudaStream_t)stream;
    const b2_vit_desc& d = p->d;
    VP P_{params, d.depth};
    const int E = p->E, M = p->M, Mp = p->Mp, T = p->T, H = p->H;
    int rc;
    __nv_bfloat16* Pm = at<__nv_bfloat16>(ws, p->off_P);
    __nv_bfloat16* tok = at<__nv_bfloat16>(ws, p->off_tok);
    float* gscr = at<float>(ws, p->s_gemm);
    // weight shadows (bf16; transposed copies feed the data gradients)
    for (int l = 0; l < d.depth; ++l) {
        char* wb = (char*)ws + p->off_w + (size_t)l * p->w_blk_bytes;
        const bool tr = keep_for_backward != 0;
        if ((rc = cast_w(P_.blk(l, 2), 3 * E, E, (__nv_bfloat16*)(wb + p->w_qkv), tr ? (__nv_bfloat16*)(wb + p->w_qkvT) : nullptr, st))) return rc;
        if ((rc = cast_w(P_.blk(l, 4), E, E, (__nv_bfloat16*)(wb + p->w_proj), tr ? (__nv_bfloat16*)(wb + p->w_projT) : nullptr, st))) return rc;
        if ((rc = cast_w(P_.blk(l, 8), 4 * E, E, (__nv_bfloat16*)(wb + p->w_fc1), tr ? (__nv_bfloat16*)(wb + p->w_fc1T) : nullptr, st))) return rc;
        if ((rc = cast_w(P_.blk(l, 10), E, 4 * E, (__nv_bfloat16*)(wb + p->w_fc2), tr ? (__nv_bfloat16*)(wb + p->w_fc2T) : nullptr, st))) return rc;
    }
    if ((rc = cast_w(P_.pew(), E, p->Kp, at<__nv_bfloat16>(ws, p->off_wpe), keep_for_backward ? at<__nv_bfloat16>(ws, p->off_wpeT) : nullptr, st))) return rc;
    // patch embedding (vision_transformer.py:70-78): tokens = patches . Wpe^T + b
    PatchGeom pg{d.batch, d.in_channels, d.D, d.H, d.W, skip0->pitch, d.patch, d.D / d.patch, d.H / d.patch, d.W / d.patch, p->Kp};
    const size_t psm = (size_t)d.patch * d.patch * (d.in_channels + 2) * 2;
    B2_CHECK_ARG(psm <= 48 * 1024);
    if (p->Mtp > p->Mt) B2_CUDA(cudaMemsetAsync(Pm + (size_t)p->Mt * p->Kp, 0, (size_t)(p->Mtp - p->Mt) * p->Kp * 2, st));
    B2_LAUNCH(patchify_kernel, dim3(p->Mt, d.patch), 256, psm, st, pg, (const __nv_bfloat16*)skip0->ptr, Pm);
    if ((rc = gemm_tn_bf16(Pm, p->Mtp, p->Kp, p->Kp, at<__nv_bfloat16>(ws, p->off_wpe), E, P_.peb(), tok, E, 0, gscr, p->gemm_scr_bytes, st))) return rc;
    float* X = at<float>(ws, p->off_X0);
    if (Mp > M) B2_CUDA(cudaMemsetAsync(X + (size_t)M * E, 0, (size_t)(Mp - M) * E * 4, st));
    B2_LAUNCH(assemble_tokens_kernel, (int)grid_for((long long)M * E, 1), 256, 0, st, (const __nv_bfloat16*)tok, P_.cls(), P_.pos(), X, d.batch, p->np, E);
    __nv_bfloat16* tmp = at<__nv_bfloat16>(ws, p->s_tmp);
    const size_t at_smem = (size_t)5 * FA_TILE * 2;
    const int at_blocks = cdiv(T, FA_ROWS);
    for (int l = 0; l < d.depth; ++l) {
        char* bb = (char*)ws + p->off_blocks + (size_t)l * p->blk_bytes;
        char* wb = (char*)ws + p->off_w + (size_t)l * p->w_blk_bytes;
        float* Xin = (float*)(bb + p->b_Xin);
        float* Xmid = (float*)(bb + p->b_Xmid);
        __nv_bfloat16* Xn = (__nv_bfloat16*)(bb + p->b_Xn);
        __nv_bfloat16* qkv = (__nv_bfloat16*)(bb + p->b_qkv);
        __nv_bfloat16* O = (__nv_bfloat16*)(bb + p->b_O);
        __nv_bfloat16* Xn2 = (__nv_bfloat16*)(bb + p->b_Xn2);
        __nv_bfloat16* Hpre = (__nv_bfloat16*)(bb + p->b_Hpre);
        __nv_bfloat16* Hact = (__nv_bfloat16*)(bb + p->b_Hact);
        B2_CUDA(cudaMemcpyAsync(Xin, X, (size_t)Mp * E * 4, cudaMemcpyDeviceToDevice, st));       // residual stream entering the block
        if ((rc = ln_fwd(X, E, P_.blk(l, 0), P_.blk(l, 1), Xn, E, nullptr, (float*)(bb + p->b_st1), M, E, d.ln_eps, st))) return rc;
        if (Mp > M) {
            B2_CUDA(cudaMemsetAsync(Xn + (size_t)M * E, 0, (size_t)(Mp - M) * E * 2, st));
            B2_CUDA(cudaMemsetAsync(O + (size_t)M * E, 0, (size_t)(Mp - M) * E * 2, st));
            B2_CUDA(cudaMemsetAsync(Xn2 + (size_t)M * E, 0, (size_t)(Mp - M) * E * 2, st));
        }
        if ((rc = gemm_tn_bf16(Xn, Mp, E, E, (__nv_bfloat16*)(wb + p->w_qkv), 3 * E, P_.blk(l, 3), qkv, 3 * E, 0, gscr, p->gemm_scr_bytes, st))) return rc;
        B2_LAUNCH(attn_fwd_kernel, dim3(d.batch * H, at_blocks), FA_THREADS, at_smem, st, (const __nv_bfloat16*)qkv, O, (float*)(bb + p->b_lse),
                  d.batch, H, T, 0.125f, d.lsa ? P_.lsa_scale(l) : (const float*)nullptr, d.lsa ? d.lsa_mask : 0);
        if ((rc = gemm_tn_bf16(O, Mp, E, E, (__nv_bfloat16*)(wb + p->w_proj), E, P_.blk(l, 5), tmp, E, 0, gscr, p->gemm_scr_bytes, st))) return rc;
        B2_LAUNCH(add_bf16_kernel, (int)grid_for((long long)M * E, 4), 256, 0, st, X, (const __nv_bfloat16*)tmp, (long long)M * E);
        B2_CUDA(cudaMemcpyAsync(Xmid, X, (size_t)Mp * E * 4, cudaMemcpyDeviceToDevice, st));
        if ((rc = ln_fwd(X, E, P_.blk(l, 6), P_.blk(l, 7), Xn2, E, nullptr, (float*)(bb + p->b_st2), M, E, d.ln_eps, st))) return rc;
        if ((rc = gemm_tn_bf16(Xn2, Mp, E, E, (__nv_bfloat16*)(wb + p->w_fc1), 4 * E, P_.blk(l, 9), Hpre, 4 * E, 0, gscr, p->gemm_scr_bytes, st))) return rc;
        B2_LAUNCH(gelu_kernel<0>, (int)grid_for((long long)Mp * 4 * E, 8), 256, 0, st, (const __nv_bfloat16*)Hpre, Hact, (long long)Mp * 4 * E);
        if ((rc = gemm_tn_bf16(Hact, Mp, 4 * E, 4 * E, (__nv_bfloat16*)(wb + p->w_fc2), E, P_.blk(l, 11), tmp, E, 0, gscr, p->gemm_scr_bytes, st))) return rc;
        B2_LAUNCH(add_bf16_kernel, (int)grid_for((long long)M * E, 4), 256, 0, st, X, (const __nv_bfloat16*)tmp, (long long)M * E);
    }
    // final LayerNorm on the class tokens + head (vision_transformer.py:436-440, :456-458)
    float* clsn = at<float>(ws, p->off_clsn);
    if ((rc = ln_fwd(X, (long long)T * E, P_.normw(), P_.normb(), nullptr, 0, clsn, at<float>(ws, p->off_lnf), d.batch, E, d.ln_eps, st))) return rc;
    if (out_view) {
        B2_CHECK_ARG(out_view->n == d.batch && out_view->c == d.out_c && out_view->d == d.out_d && out_view->h == d.out_h && out_view->w == d.out_w);
        HeadGeom hg{d.batch, E, p->F, d.out_c, d.out_d, d.out_h, d.out_w, out_view->pitch};
        if (out_view->dtype == B2_BF16) B2_LAUNCH(head_fwd_kernel<__nv_bfloat16>, cdiv(p->F, 8), 256, 0, st, hg, (const float*)clsn, P_.headw(), P_.headb(), (__nv_bfloat16*)out_view->ptr);
        else B2_LAUNCH(head_fwd_kernel<float>, cdiv(p->F, 8), 256, 0, st, hg, (const float*)clsn, P_.headw(), P_.headb(), (float*)out_view->ptr);
    }
    if (out_dense) {   // [B][F] fp32, row-major (what VisionTransformer.forward returns)
        HeadGeom hg{d.batch, E, p->F, p->F, 1, 1, 1, p->F};
        // dense layout: address(b, j) = b * F + j  <=>  NDHWC view with one voxel and F channels
        B2_LAUNCH(head_fwd_kernel<float>, cdiv(p->F, 8), 256, 0, st, hg, (const float*)clsn, P_.headw(), P_.headb(), out_dense);
    }
    return B2_OK;
}

extern "C" int b2_vit_backward(b2_vit_plan* p, const float* const* params, const b2_act_view* dout_view, const float* dout_dense, void* ws,
                               const b2_act_view* dskip0, float* const* grads, b2_stream_t stream) {
    B2_CHECK_ARG(p && params && ws && grads && (dout_view || dout_dense));
    cudaStream_t st = (cudaStream_t)stream;
    const b2_vit_desc& d = p->d;
    VP P_{params, d.depth};
    const int E = p->E, M = p->M, Mp = p->Mp, T = p->T, H = p->H, dep = d.depth;
    auto G = [&](int i) { return grads[i]; };
    auto GB = [&](int l, int i) { return grads[2 + 12 * l + i]; };
    int rc;
    float* part = at<float>(ws, p->s_part);
    float* gscr = at<float>(ws, p->s_gemm);
    float* dX = at<float>(ws, p->s_dX);
    float* X = at<float>(ws, p->off_X0);          // residual stream after the last block (forward left it there)
    float* clsn = at<float>(ws, p->off_clsn);
    // ---- head + final LayerNorm ------------------------------------------------------------------------------------------
    const int rpb = 256, hblocks = cdiv(p->F, rpb);
    float* dclsn = part + (size_t)hblocks * d.batch * E;     // [B][E] after the ordered reduce
    if (dout_view) {
        HeadGeom hg{d.batch, E, p->F, d.out_c, d.out_d, d.out_h, d.out_w, dout_view->pitch};
        if (dout_view->dtype == B2_BF16) B2_LAUNCH(head_bwd_kernel<__nv_bfloat16>, hblocks, 256, 0, st, hg, (const __nv_bfloat16*)dout_view->ptr, (const float*)clsn, P_.headw(), G(6 + 12 * dep), G(7 + 12 * dep), part, rpb);
        else B2_LAUNCH(head_bwd_kernel<float>, hblocks, 256, 0, st, hg, (const float*)dout_view->ptr, (const float*)clsn, P_.headw(), G(6 + 12 * dep), G(7 + 12 * dep), part, rpb);
    } else {
        HeadGeom hg{d.batch, E, p->F, p->F, 1, 1, 1, p->F};
        B2_LAUNCH(head_bwd_kernel<float>, hblocks, 256, 0, st, hg, dout_dense, (const float*)clsn, P_.headw(), G(6 + 12 * dep), G(7 + 12 * dep), part, rpb);
    }
    B2_LAUNCH(head_bwd_final_kernel, cdiv(d.batch * E, 256), 256, 0, st, (const float*)part, hblocks, d.batch * E, dclsn);
    B2_CUDA(cudaMemsetAsync(dX, 0, (size_t)Mp * E * 4, st));
    float* lnpart = dclsn + (size_t)HEAD_MAXB * E;
    if ((rc = ln_bwd<float>(dclsn, E, X, (long long)T * E, at<float>(ws, p->off_lnf), P_.normw(), dX, (long long)T * E, 0, d.batch, E, lnpart,
                            G(2 + 12 * dep), G(3 + 12 * dep), st))) return rc;
    __nv_bfloat16* tmp = at<__nv_bfloat16>(ws, p->s_tmp);
    __nv_bfloat16* tmp2 = at<__nv_bfloat16>(ws, p->s_tmp2);
    __nv_bfloat16* tA = at<__nv_bfloat16>(ws, p->s_tA);
    __nv_bfloat16* tB = at<__nv_bfloat16>(ws, p->s_tB);
    float* Drow = at<float>(ws, p->s_drow);
    // transposed operands: columns [M, Mp) must be zero (they are contraction padding)
    B2_CUDA(cudaMemsetAsync(tA, 0, (size_t)4 * E * Mp * 2, st));
    B2_CUDA(cudaMemsetAsync(tB, 0, (size_t)4 * E * Mp * 2, st));
    const size_t q_smem = (size_t)6 * FA_TILE * 2 + FA_ROWS * 4;
    const size_t kv_smem = (size_t)6 * FA_TILE * 2 + 4 * FA_ROWS * 4;
    const int at_blocks = cdiv(T, FA_ROWS);
    static bool at_attr = false;
    if (!at_attr) {
        B2_CUDA(cudaFuncSetAttribute(attn_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q_smem));
        B2_CUDA(cudaFuncSetAttribute(attn_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kv_smem));
        at_attr = true;
    }
    // dy = bf16(dX) (pad rows are zero because dX's are)
    auto cast_dx = [&](__nv_bfloat16* dst) -> int {
        B2_LAUNCH(cast_rows_kernel, (int)grid_for((long long)Mp * E, 4), 256, 0, st, (const float*)dX, dst, (long long)Mp * E);
        return B2_OK;
    };
    // one Linear backward: dIn[Mp][K] = dOut[Mp][N] . W  (WT shadow [K][N]);  dW[N][K] = dOut^T . In;  db = colsum(dOut)
    auto linear_bwd = [&](const __nv_bfloat16* dOut, int N, const __nv_bfloat16* In, int K, const __nv_bfloat16* WT, __nv_bfloat16* dIn,
                          float* dW, float* db) -> int {
        int r;
        if (dIn && (r = gemm_tn_bf16(dOut, Mp, N, N, WT, K, nullptr, dIn, K, 0, gscr, p->gemm_scr_bytes, st))) return r;
        if ((r = transpose(dOut, M, N, N, tA, Mp, st))) return r;
        if ((r = transpose(In, M, K, K, tB, Mp, st))) return r;
        if ((r = gemm_tn_bf16(tA, N, Mp, Mp, tB, K, nullptr, dW, K, 1, gscr, p->gemm_scr_bytes, st))) return r;
        return db ? colsum<__nv_bfloat16>(dOut, M, N, N, part, db, st) : B2_OK;     // (LSA: qkv has no bias)
    };
    for (int l = dep - 1; l >= 0; --l) {
        char* bb = (char*)ws + p->off_blocks + (size_t)l * p->blk_bytes;
        char* wb = (char*)ws + p->off_w + (size_t)l * p->w_blk_bytes;
        // ---- MLP branch ----
        if ((rc = cast_dx(tmp))) return rc;                                                             // d(fc2 out)
        if ((rc = linear_bwd(tmp, E, (const __nv_bfloat16*)(bb + p->b_Hact), 4 * E, (const __nv_bfloat16*)(wb + p->w_fc2T), tmp2, GB(l, 10), GB(l, 11)))) return rc;
        B2_LAUNCH(gelu_kernel<1>, (int)grid_for((long long)Mp * 4 * E, 8), 256, 0, st, (const __nv_bfloat16*)(bb + p->b_Hpre), tmp2, (long long)Mp * 4 * E);
        if ((rc = linear_bwd(tmp2, 4 * E, (const __nv_bfloat16*)(bb + p->b_Xn2), E, (const __nv_bfloat16*)(wb + p->w_fc1T), tmp, GB(l, 8), GB(l, 9)))) return rc;
        if ((rc = ln_bwd<__nv_bfloat16>(tmp, E, (const float*)(bb + p->b_Xmid), E, (const float*)(bb + p->b_st2), P_.blk(l, 6), dX, E, 1, M, E, part,
                                        GB(l, 6), GB(l, 7), st))) return rc;
        // ---- attention branch ----
        if ((rc = cast_dx(tmp))) return rc;                                                             // d(proj out)
        if ((rc = linear_bwd(tmp, E, (const __nv_bfloat16*)(bb + p->b_O), E, (const __nv_bfloat16*)(wb + p->w_projT), tmp2, GB(l, 4), GB(l, 5)))) return rc;
        __nv_bfloat16* dqkv = tmp;     // [Mp][3E]
        B2_CUDA(cudaMemsetAsync(dqkv + (size_t)M * 3 * E, 0, (size_t)(Mp - M) * 3 * E * 2, st));
        const float* sc_h = d.lsa ? P_.lsa_scale(l) : (const float*)nullptr;
        float* dsc_part = d.lsa ? part : (float*)nullptr;        // [B*H][at_blocks] (the colsum partials are consumed by now)
        B2_LAUNCH(attn_bwd_q_kernel, dim3(d.batch * H, at_blocks), FA_THREADS, q_smem, st, (const __nv_bfloat16*)(bb + p->b_qkv),
                  (const __nv_bfloat16*)(bb + p->b_O), (const __nv_bfloat16*)tmp2, (const float*)(bb + p->b_lse), dqkv, Drow, d.batch, H, T,
                  0.125f, sc_h, d.lsa ? d.lsa_mask : 0, dsc_part);
        if (d.lsa)
            B2_LAUNCH(attn_dscale_final_kernel, cdiv(H, 32), 32, 0, st, (const float*)dsc_part, d.batch, H, at_blocks, G(2 + 12 * dep + 6 + l));
        B2_LAUNCH(attn_bwd_kv_kernel, dim3(d.batch * H, at_blocks), FA_THREADS, kv_smem, st, (const __nv_bfloat16*)(bb + p->b_qkv),
                  (const __nv_bfloat16*)tmp2, (const float*)(bb + p->b_lse), (const float*)Drow, dqkv, d.batch, H, T, 0.125f, sc_h,
                  d.lsa ? d.lsa_mask : 0);
        if ((rc = linear_bwd(dqkv, 3 * E, (const __nv_bfloat16*)(bb + p->b_Xn), E, (const __nv_bfloat16*)(wb + p->w_qkvT), tmp2, GB(l, 2), GB(l, 3)))) return rc;
        if ((rc = ln_bwd<__nv_bfloat16>(tmp2, E, (const float*)(bb + p->b_Xin), E, (const float*)(bb + p->b_st1), P_.blk(l, 0), dX, E, 1, M, E, part,
                                        GB(l, 0), GB(l, 1), st))) return rc;
    }
    // ---- token assembly + patch embedding ---------------------------------------------------------------------------------------
    __nv_bfloat16* dtok = at<__nv_bfloat16>(ws, p->off_tok);
    if (p->Mtp > p->Mt) B2_CUDA(cudaMemsetAsync(dtok + (size_t)p->Mt * E, 0, (size_t)(p->Mtp - p->Mt) * E * 2, st));
    B2_LAUNCH(assemble_tokens_bwd_kernel, (int)grid_for((long long)T * E, 1), 256, 0, st, (const float*)dX, G(0), G(1), dtok, d.batch, p->np, E);
    if ((rc = colsum<__nv_bfloat16>(dtok, p->Mt, E, E, part, G(5 + 12 * dep), st))) return rc;
    __nv_bfloat16* Pm = at<__nv_bfloat16>(ws, p->off_P);
    __nv_bfloat16* Pt = at<__nv_bfloat16>(ws, p->off_Pt);
    B2_CUDA(cudaMemsetAsync(Pt, 0, (size_t)p->Kp * p->Mtp * 2, st));
    if ((rc = transpose(Pm, p->Mt, p->Kp, p->Kp, Pt, p->Mtp, st))) return rc;
    B2_CUDA(cudaMemsetAsync(tA, 0, (size_t)E * p->Mtp * 2, st));
    if ((rc = transpose(dtok, p->Mt, E, E, tA, p->Mtp, st))) return rc;
    if ((rc = gemm_tn_bf16(tA, E, p->Mtp, p->Mtp, Pt, p->Kp, nullptr, G(4 + 12 * dep), p->Kp, 1, gscr, p->gemm_scr_bytes, st))) return rc;
    if (dskip0) {
        B2_CHECK_ARG(dskip0->dtype == B2_BF16);
        if ((rc = gemm_tn_bf16(dtok, p->Mtp, E, E, at<__nv_bfloat16>(ws, p->off_wpeT), p->Kp, nullptr, Pm, p->Kp, 0, gscr, p->gemm_scr_bytes, st))) return rc;
        PatchGeom pg{d.batch, d.in_channels, d.D, d.H, d.W, dskip0->pitch, d.patch, d.D / d.patch, d.H / d.patch, d.W / d.patch, p->Kp};
        const size_t psm = (size_t)d.patch * d.patch * (d.in_channels + 2) * 2;
        B2_LAUNCH(unpatchify_add_kernel, dim3(p->Mt, d.patch), 256, psm, st, pg, (const __nv_bfloat16*)Pm, (__nv_bfloat16*)dskip0->ptr);
    }
    return B2_OK;
}
