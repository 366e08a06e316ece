// loss.cu -- deep-supervision Dice+CE loss (value + gradient in two sweeps), LwF / MiB distillation, PLOP pseudo-label
// loss, hard tp/fp/fn for the online evaluation.  Logits are NCDHW fp32 (B, C, V), exactly the tensors the reference's
// loss modules receive.  All reductions are two-stage with a fixed order (bit-reproducible); HBM-bound kernels.
//
//   dsloss : nnunet DC_and_CE_loss inside MultipleOutputLoss2 (reference MultiHead:1385-1386)
//   kd_lwf : reference loss_functions/deep_supervision.py:185-199
//   kd_mib : reference loss_functions/knowledge_distillation.py:11-32 (equal class counts)
//   plop   : reference loss_functions/deep_supervision.py:287-332, crossentropy.py:6-16
//   eval   : reference MultiHead:938-951
#include "common.cuh"
#include "kernels.h"

namespace b2 {

constexpr int MAXC = 8;
constexpr int MAXB = 32;     // samples per loss launch (PLOP trains with batch 25 from the second task on, plop:85)

static int loss_slabs(int B, long long V) {
    long long s = (4LL * num_sms() + B - 1) / B;
    long long maxs = (V + 1023) / 1024;
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    return (int)s;
}

// block-wide ordered sum of `nvals` per-thread values -> part[nvals] (thread 0..nvals-1 write)
template <int NV>
__device__ __forceinline__ void block_reduce_store(const float (&vals)[NV], int nvals, float* sh /*[8][NV]*/, float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i < nvals) {
            float v = warp_sum(vals[i]);
            if (lane == 0) sh[warp * NV + i] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < nvals) {
        float s = 0.f;
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; ++w) s += sh[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Dice + CE
// ---------------------------------------------------------------------------------------------------------------
// partial layout per (b, slab): [0..C) sum p, [C..2C) sum p*y, [2C..3C) sum y, [3C] ce sum, [3C+1] valid count
__global__ void __launch_bounds__(256) dsloss_reduce_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                            int C, long long V, int slabs, int ignore_index,
                                                            float* __restrict__ part) {
    pdl_grid_sync();
    __shared__ float sh[8 * (3 * MAXC + 2)];
    const int b = blockIdx.y, slab = blockIdx.x;
    const long long per = (V + slabs - 1) / slabs;
    const long long v0 = (long long)slab * per, v1 = v0 + per < V ? v0 + per : V;
    float acc[3 * MAXC + 2];
#pragma unroll
    for (int i = 0; i < 3 * MAXC + 2; ++i) acc[i] = 0.f;
    const float* lg = logits + (long long)b * C * V;
    const float* tg = target + (long long)b * V;
    for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        float x[MAXC];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { x[c] = lg[(long long)c * V + v]; m = fmaxf(m, x[c]); }
        float se = 0.f, xm[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { xm[c] = x[c] - m; x[c] = expf(xm[c]); se += x[c]; }
        const float inv = 1.f / se, lse = logf(se);
        const int y = (int)tg[v];
        const bool ign = (y == ignore_index);
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                const float p = x[c] * inv;
                acc[c] += p;
                if (c == y) {
                    acc[MAXC + c] += p;
                    acc[2 * MAXC + c] += 1.f;
                    if (!ign) { acc[3 * MAXC] += lse - xm[c]; }
                }
            }
        if (!ign) acc[3 * MAXC + 1] += 1.f;
    }
    block_reduce_store<3 * MAXC + 2>(acc, 3 * MAXC + 2, sh, part + ((long long)b * slabs + slab) * (3 * MAXC + 2));
}

// coef layout: [B][C][2] = {A, Bc}; then [0] = 1/valid_count
__global__ void dsloss_finalize_kernel(const float* __restrict__ part, int B, int C, int slabs, float weight,
                                       int batch_dice, float smooth, int do_bg, int with_dice, float* __restrict__ coef,
                                       float* __restrict__ loss_out) {
    pdl_grid_sync();
    constexpr int NV = 3 * MAXC + 2;
    // per sample: stage 1 (256 threads): lane = value index inside a partial row (coalesced 104-byte rows), warp = slab lane,
    // several independent loads in flight; stage 2: the 8 warps combined in order.  Thread 0 then evaluates the loss.
    __shared__ double wsum[8][NV];
    __shared__ double tot[MAXB][NV];
    const int Bc = B < MAXB ? B : MAXB;      // (the host entry point rejects B > MAXB)
    const int j = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b = 0; b < Bc; ++b) {
        if (j < NV) {
            double a = 0.0;
#pragma unroll 8
            for (int sl = w; sl < slabs; sl += 8) a += (double)part[((long long)b * slabs + sl) * NV + j];
            wsum[w][j] = a;
        }
        __syncthreads();
        if (threadIdx.x < NV) {
            double a = 0.0;
            for (int k = 0; k < 8; ++k) a += wsum[k][threadIdx.x];
            tot[b][threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double sp[MAXB][MAXC], spy[MAXB][MAXC], sy[MAXB][MAXC];
    double ce = 0.0, cnt = 0.0;
    for (int b = 0; b < Bc; ++b) {
        for (int c = 0; c < MAXC; ++c) {
            sp[b][c] = c < C ? tot[b][c] : 0.0;
            spy[b][c] = c < C ? tot[b][MAXC + c] : 0.0;
            sy[b][c] = c < C ? tot[b][2 * MAXC + c] : 0.0;
        }
        ce += tot[b][3 * MAXC];
        cnt += tot[b][3 * MAXC + 1];
    }
    double loss = ce / cnt;  // NaN when every voxel is ignored, like torch
    coef[(long long)B * C * 2] = (float)(1.0 / cnt);
    for (int i = 0; i < B * C * 2; ++i) coef[i] = 0.f;
    if (with_dice) {
        const int cstart = do_bg ? 0 : 1;
        if (batch_dice) {
            const double M = (double)(C - cstart);
            double dsum = 0.0;
            for (int c = cstart; c < C; ++c) {
                double tp = 0, P = 0, Y = 0;
                for (int b = 0; b < Bc; ++b) { tp += spy[b][c]; P += sp[b][c]; Y += sy[b][c]; }
                const double N = 2 * tp + smooth, D = P + Y + smooth + 1e-8;
                dsum += N / D;
                for (int b = 0; b < Bc; ++b) {
                    coef[((long long)b * C + c) * 2] = (float)(-2.0 / (M * D));
                    coef[((long long)b * C + c) * 2 + 1] = (float)(N / (M * D * D));
                }
            }
            loss += -dsum / M;
        } else {
            const double M = (double)Bc * (C - cstart);
            double dsum = 0.0;
            for (int b = 0; b < Bc; ++b)
                for (int c = cstart; c < C; ++c) {
                    const double N = 2 * spy[b][c] + smooth, D = sp[b][c] + sy[b][c] + smooth + 1e-8;
                    dsum += N / D;
                    coef[((long long)b * C + c) * 2] = (float)(-2.0 / (M * D));
                    coef[((long long)b * C + c) * 2 + 1] = (float)(N / (M * D * D));
                }
            loss += -dsum / M;
        }
    }
    loss_out[0] += (float)(weight * loss);
}

__global__ void __launch_bounds__(256) dsloss_grad_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                          int B, int C, long long V, int ignore_index, float weight,
                                                          const float* __restrict__ coef, float* __restrict__ dlogits) {
    pdl_grid_sync();
    const long long total = (long long)B * V;
    const float inv_cnt = coef[(long long)B * C * 2];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / V);
        const long long v = i % V;
        const float* lg = logits + (long long)b * C * V + v;
        float x[MAXC], g[MAXC];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { x[c] = lg[(long long)c * V]; m = fmaxf(m, x[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { x[c] = expf(x[c] - m); se += x[c]; }
        const float inv = 1.f / se;
        const int y = (int)target[i];
        const bool ign = (y == ignore_index);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                x[c] *= inv;
                const float* cf = coef + ((long long)b * C + c) * 2;
                g[c] = cf[0] * (c == y ? 1.f : 0.f) + cf[1];
                dot += x[c] * g[c];
            }
        float* dl = dlogits + (long long)b * C * V + v;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                float d = x[c] * (g[c] - dot);
                if (!ign) d += (x[c] - (c == y ? 1.f : 0.f)) * inv_cnt;
                dl[(long long)c * V] = weight * d;
            }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// distillation
// ---------------------------------------------------------------------------------------------------------------
// MODE 0: LwF KL (value only). MODE 1: MiB unbiased KD (value + gradient accumulate).
template <int MODE>
__global__ void __launch_bounds__(256) kd_kernel(const float* __restrict__ x, const float* __restrict__ t, int C,
                                                 long long V, int slabs, float a, float gscale,
                                                 float* __restrict__ dlogits, float* __restrict__ part) {
    pdl_grid_sync();
    __shared__ float sh[8];
    const int b = blockIdx.y, slab = blockIdx.x;
    const long long per = (V + slabs - 1) / slabs;
    const long long v0 = (long long)slab * per, v1 = v0 + per < V ? v0 + per : V;
    const float* xp = x + (long long)b * C * V;
    const float* tp = t + (long long)b * C * V;
    float acc = 0.f;
    for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        float xs[MAXC], ts[MAXC];
        float mx = -INFINITY, mt = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                xs[c] = xp[(long long)c * V + v] * (MODE == 0 ? a : 1.f);
                ts[c] = tp[(long long)c * V + v] * a;
                mx = fmaxf(mx, xs[c]);
                mt = fmaxf(mt, ts[c]);
            }
        float sx = 0.f, stt = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { sx += expf(xs[c] - mx); stt += expf(ts[c] - mt); }
        const float lsx = mx + logf(sx), lst = mt + logf(stt);
        float term = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                const float lq = ts[c] - lst, lp = xs[c] - lsx;
                const float q = expf(lq);
                if (MODE == 0) term += q * (lq - lp);
                else {
                    term += q * lp;
                    if (dlogits) {
                        float* d = dlogits + (long long)b * C * V + (long long)c * V + v;
                        *d += gscale * (q - expf(lp));
                    }
                }
            }
        acc += term;
    }
    float r = block_sum(acc, sh);
    if (threadIdx.x == 0) part[(long long)b * slabs + slab] = r;
}

__global__ void scalar_finalize_kernel(const float* __restrict__ part, int n, double scale, float* __restrict__ out) {
    pdl_grid_sync();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += part[i];
    out[0] += (float)(s * scale);
}

// ---------------------------------------------------------------------------------------------------------------
// PLOP pseudo-label loss (3D quirk Q10: adaptive factor per (b, d) row broadcast along the LAST axis index... see
// reference deep_supervision.py:306-311: num/den are summed over dims (1,2) of a (B,D,H,W) mask => shape (B,W);
// factor[:, None, None] has shape (B,1,1,W)?? no -- (B,W)[:,None,None] = (B,1,1,W); times a 0-dim CE => (B,1,1,W);
// .mean() => mean over b,w of factor[b,w] * (loss_pseudo + loss_not_pseudo).
// ---------------------------------------------------------------------------------------------------------------
// stage 1: per voxel classify {valid&bg -> pseudo label, else}; accumulate per (b,w): num, den; write label code
//   code[v] = pseudo label (0..C-1) if mask (bg & valid) else -1
__global__ void __launch_bounds__(256) plop_mask_kernel(const float* __restrict__ x_old, const float* __restrict__ target,
                                                        int C, int D, int H, int W, const float* __restrict__ thr,
                                                        float max_entropy, int8_t* __restrict__ code,
                                                        float* __restrict__ numden /*[Z][B][W][2]*/) {
    pdl_grid_sync();
    // one block per (b, w-chunk of 32 columns, row chunk z): threads x = w lane, y = row lanes; ordered reduce over (d,h).
    // (Round 1 had no row chunks: 12 blocks for a 48x192x192 level, 575 us per launch.)
    __shared__ float sh[8][32][2];
    const int b = blockIdx.y;
    const int w = blockIdx.x * 32 + (threadIdx.x & 31), lane = threadIdx.x >> 5;
    const long long V = (long long)D * H * W;
    const float factor = 1.f / logf((float)C + 1e-8f);
    float num = 0.f, den = 0.f;
    const int rows_per = (D * H + (int)gridDim.z - 1) / (int)gridDim.z;
    const int r_begin = (int)blockIdx.z * rows_per, r_end = r_begin + rows_per < D * H ? r_begin + rows_per : D * H;
    if (w < W) {
        for (int r = r_begin + lane; r < r_end; r += 8) {
            const long long v = (long long)r * W + w;
            float xs[MAXC];
            float m = -INFINITY;
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
                if (c < C) { xs[c] = x_old[((long long)b * C + c) * V + v]; m = fmaxf(m, xs[c]); }
            float se = 0.f;
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
                if (c < C) { xs[c] = expf(xs[c] - m); se += xs[c]; }
            float ent = 0.f, best = -1.f;
            int arg = 0;
#pragma unroll
            for (int c = 0; c < MAXC; ++c)
                if (c < C) {
                    const float p = xs[c] / se;
                    ent += p * logf(p + 1e-8f);
                    if (p > best) { best = p; arg = c; }
                }
            ent = -factor * (ent / (float)C);
            const bool valid = (ent / max_entropy) < thr[arg];
            const bool bg = target[(long long)b * V + v] == 0.f;
            code[(long long)b * V + v] = (valid && bg) ? (int8_t)arg : (int8_t)-1;
            if (bg) den += 1.f;
            if (bg && valid) num += 1.f;
        }
    }
    sh[lane][threadIdx.x & 31][0] = num;
    sh[lane][threadIdx.x & 31][1] = den;
    __syncthreads();
    if (lane == 0 && w < W) {
        float a = 0.f, d = 0.f;
        for (int l = 0; l < 8; ++l) { a += sh[l][threadIdx.x][0]; d += sh[l][threadIdx.x][1]; }
        float* nd = numden + (long long)blockIdx.z * gridDim.y * W * 2;
        nd[((long long)b * W + w) * 2] = a;
        nd[((long long)b * W + w) * 2 + 1] = d;
    }
}

// stage 2: two CE sums over all voxels: pseudo (label = code where code>=0, else ignored) and not-pseudo
// (label = y where code<0, else ignored).  partial per block: {ce_p, cnt_p, ce_n, cnt_n}
__global__ void __launch_bounds__(256) plop_ce_reduce_kernel(const float* __restrict__ x, const float* __restrict__ target,
                                                             const int8_t* __restrict__ code, int C, long long V,
                                                             int slabs, float* __restrict__ part) {
    pdl_grid_sync();
    __shared__ float sh[8 * 4];
    const int b = blockIdx.y, slab = blockIdx.x;
    const long long per = (V + slabs - 1) / slabs;
    const long long v0 = (long long)slab * per, v1 = v0 + per < V ? v0 + per : V;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        float xs[MAXC];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { xs[c] = x[((long long)b * C + c) * V + v]; m = fmaxf(m, xs[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) se += expf(xs[c] - m);
        const float lse = m + logf(se);
        const int cd = code[(long long)b * V + v];
        int lab = cd >= 0 ? cd : (int)target[(long long)b * V + v];
        float xl = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C && c == lab) xl = xs[c];
        if (lab >= 0 && lab < C) {
            if (cd >= 0) { acc[0] += lse - xl; acc[1] += 1.f; }
            else { acc[2] += lse - xl; acc[3] += 1.f; }
        }
    }
    block_reduce_store<4>(acc, 4, sh, part + ((long long)b * slabs + slab) * 4);
}

// finalize: value = weight * mean_{b,w}(num/den) * (ce_p/cnt_p + ce_n/cnt_n); coef = {fbar/cnt_p, fbar/cnt_n} * weight
__global__ void __launch_bounds__(256) plop_finalize_kernel(const float* __restrict__ part, int nparts, const float* __restrict__ numden,
                                                            int nz, int B, int W, float weight, float* __restrict__ coef,
                                                            float* __restrict__ loss_out) {
    pdl_grid_sync();
    // one block: strided partial sums per thread, shuffle tree per warp, the 8 warps combined in order (fixed order throughout;
    // the single-thread version of round 1 spent 195 us per launch on dependent loads)
    __shared__ double sh[8][5];
    double s[4] = {0, 0, 0, 0}, f = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256)
        for (int k = 0; k < 4; ++k) s[k] += part[(long long)i * 4 + k];
    for (int i = threadIdx.x; i < B * W; i += 256) {
        float num = 0.f, den = 0.f;     // counts: exact in fp32 in any order
        for (int z = 0; z < nz; ++z) { num += numden[((long long)z * B * W + i) * 2]; den += numden[((long long)z * B * W + i) * 2 + 1]; }
        f += (double)(num / den);  // fp32 division like torch
    }
    for (int k = 0; k < 4; ++k) s[k] = warp_sum(s[k]);
    f = warp_sum(f);
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 4; ++k) sh[threadIdx.x >> 5][k] = s[k];
        sh[threadIdx.x >> 5][4] = f;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int k = 0; k < 4; ++k) s[k] = 0.0;
    f = 0.0;
    for (int w = 0; w < 8; ++w) {
        for (int k = 0; k < 4; ++k) s[k] += sh[w][k];
        f += sh[w][4];
    }
    f /= (double)(B * W);
    const double lp = s[0] / s[1], ln = s[2] / s[3];
    loss_out[0] += (float)(weight * f * (lp + ln));
    coef[0] = (float)(weight * f / s[1]);
    coef[1] = (float)(weight * f / s[3]);
}

__global__ void __launch_bounds__(256) plop_grad_kernel(const float* __restrict__ x, const float* __restrict__ target,
                                                        const int8_t* __restrict__ code, int B, int C, long long V,
                                                        const float* __restrict__ coef, float* __restrict__ dlogits) {
    pdl_grid_sync();
    const long long total = (long long)B * V;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / V);
        const long long v = i % V;
        float xs[MAXC];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { xs[c] = x[((long long)b * C + c) * V + v]; m = fmaxf(m, xs[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { xs[c] = expf(xs[c] - m); se += xs[c]; }
        const int cd = code[i];
        const int lab = cd >= 0 ? cd : (int)target[i];
        const bool ok = lab >= 0 && lab < C;
        const float k = ok ? (cd >= 0 ? coef[0] : coef[1]) : 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) dlogits[((long long)b * C + c) * V + v] = k * (xs[c] / se - (c == lab ? 1.f : 0.f));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// online evaluation: per (b, foreground class): tp, fp, fn counts
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eval_reduce_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                          int C, long long V, int slabs, float* __restrict__ part) {
    pdl_grid_sync();
    __shared__ float sh[8 * 3 * MAXC];
    const int b = blockIdx.y, slab = blockIdx.x;
    const long long per = (V + slabs - 1) / slabs;
    const long long v0 = (long long)slab * per, v1 = v0 + per < V ? v0 + per : V;
    float acc[3 * MAXC];
#pragma unroll
    for (int i = 0; i < 3 * MAXC; ++i) acc[i] = 0.f;
    for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        float best = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                const float xv = logits[((long long)b * C + c) * V + v];
                if (xv > best) { best = xv; arg = c; }
            }
        const int y = (int)target[(long long)b * V + v];
#pragma unroll
        for (int c = 1; c < MAXC; ++c)
            if (c < C) {
                if (arg == c && y == c) acc[c * 3] += 1.f;
                if (arg == c && y != c) acc[c * 3 + 1] += 1.f;
                if (arg != c && y == c) acc[c * 3 + 2] += 1.f;
            }
    }
    block_reduce_store<3 * MAXC>(acc, 3 * MAXC, sh, part + ((long long)b * slabs + slab) * 3 * MAXC);
}

__global__ void eval_finalize_kernel(const float* __restrict__ part, int B, int C, int slabs, float* __restrict__ out) {
    pdl_grid_sync();
    const int i = threadIdx.x;
    if (i >= B * (C - 1) * 3) return;
    const int k = i % 3, c = (i / 3) % (C - 1) + 1, b = i / (3 * (C - 1));
    double s = 0.0;
    for (int sl = 0; sl < slabs; ++sl) s += part[((long long)b * slabs + sl) * 3 * MAXC + c * 3 + k];
    out[i] = (float)s;
}

}  // namespace b2

// ===============================================================================================================
// C ABI
// ===============================================================================================================
using namespace b2;

extern "C" size_t b2_dsloss_scratch_bytes(int B, int C, int64_t V) {
    return align_up(((size_t)B * loss_slabs(B, V) * (3 * MAXC + 2) + (size_t)B * C * 2 + 8) * sizeof(float));
}

extern "C" int b2_dsloss_fwd_bwd(const float* logits, const float* target, int B, int C, int64_t V, float weight,
                                 int batch_dice, float smooth, int do_bg, int ignore_index, int with_dice,
                                 float* dlogits, float* loss_out, void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(logits && target && loss_out && scratch);
    B2_CHECK_ARG(C >= 2 && C <= MAXC && B >= 1 && B <= MAXB && V > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const int slabs = loss_slabs(B, V);
    float* part = (float*)scratch;
    float* coef = part + (size_t)B * slabs * (3 * MAXC + 2);
    dim3 grid(slabs, B);
    B2_LAUNCH(dsloss_reduce_kernel, grid, 256, 0, st, logits, target, C, (long long)V, slabs, ignore_index, part);
    B2_LAUNCH(dsloss_finalize_kernel, 1, 256, 0, st, part, B, C, slabs, weight, batch_dice, smooth, do_bg, with_dice, coef, loss_out);
    if (dlogits) {
        long long total = (long long)B * V;
        long long g = (total + 255) / 256, cap = (long long)num_sms() * 16;
        if (g > cap) g = cap;
        B2_LAUNCH(dsloss_grad_kernel, (int)g, 256, 0, st, logits, target, B, C, (long long)V, ignore_index, weight, coef, dlogits);
    }
    return B2_OK;
}

extern "C" size_t b2_kd_scratch_bytes(int B, int C, int64_t V) {
    (void)C;
    size_t slabs = loss_slabs(B, V);
    // plop needs: code (B*V bytes) + numden + partials + coef
    return align_up((size_t)B * V + 256) + align_up(((size_t)B * slabs * 4 + (size_t)32 * B * 4096 * 2 + 16) * sizeof(float));
}

extern "C" int b2_kd_lwf(const float* pred, const float* teacher, int B, int C, int64_t V, float temperature,
                         float* loss_out, void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(pred && teacher && loss_out && scratch && C >= 2 && C <= MAXC && B >= 1);
    cudaStream_t st = (cudaStream_t)stream;
    const int slabs = loss_slabs(B, V);
    float* part = (float*)scratch;
    dim3 grid(slabs, B);
    B2_LAUNCH(kd_kernel<0>, grid, 256, 0, st, pred, teacher, C, (long long)V, slabs, 1.f / temperature, 0.f, (float*)nullptr, part);
    B2_LAUNCH(scalar_finalize_kernel, 1, 32, 0, st, part, B * slabs, 1.0 / (double)B, loss_out);
    return B2_OK;
}

extern "C" int b2_kd_mib(const float* x, const float* teacher, int B, int C, int64_t V, float alpha, float scale,
                         float* dlogits, float* loss_out, void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(x && teacher && loss_out && scratch && C >= 2 && C <= MAXC && B >= 1);
    cudaStream_t st = (cudaStream_t)stream;
    const int slabs = loss_slabs(B, V);
    float* part = (float*)scratch;
    dim3 grid(slabs, B);
    const double denom = (double)B * (double)V * (double)C;
    // L = -(1/(B V C)) sum q * lsm(x);  dL/dx_k = -(1/(BVC)) (q_k - p_k)
    B2_LAUNCH(kd_kernel<1>, grid, 256, 0, st, x, teacher, C, (long long)V, slabs, alpha, (float)(-(double)scale / denom), dlogits, part);
    B2_LAUNCH(scalar_finalize_kernel, 1, 32, 0, st, part, B * slabs, -(double)scale / denom, loss_out);
    return B2_OK;
}

// PLOP threshold extraction (reference plop:113-182; arthurdouillard/CVPR2021_PLOP train.py "find_median"): histogram over the
// background voxels (target == 0) of entropy(softmax(x_old)) / max_entropy in nb_bins bins, per pseudo label (argmax).  Integer
// counts: shared-memory histogram per block, 64-bit integer atomics into the caller's table (order-independent => reproducible).
__global__ void __launch_bounds__(256) plop_hist_kernel(const float* __restrict__ x_old, const float* __restrict__ target, int B, int C,
                                                        long long V, float max_entropy, int nb_bins, unsigned long long* __restrict__ hist) {
    pdl_grid_sync();
    extern __shared__ unsigned int sh_hist[];
    for (int i = threadIdx.x; i < C * nb_bins; i += 256) sh_hist[i] = 0u;
    __syncthreads();
    const float factor = 1.f / logf((float)C + 1e-8f);
    const long long total = (long long)B * V;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        if (target[i] != 0.f) continue;
        const long long b = i / V, v = i % V;
        float xs[MAXC];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { xs[c] = x_old[(b * C + c) * V + v]; m = fmaxf(m, xs[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) { xs[c] = expf(xs[c] - m); se += xs[c]; }
        float ent = 0.f, best = -1.f;
        int arg = 0;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < C) {
                const float p = xs[c] / se;
                ent += p * logf(p + 1e-8f);
                if (p > best) { best = p; arg = c; }
            }
        ent = -factor * (ent / (float)C);
        int bin = (int)((ent / max_entropy) * (float)nb_bins);      // .long(): truncation
        bin = bin > nb_bins - 1 ? nb_bins - 1 : (bin < 0 ? 0 : bin);
        atomicAdd(&sh_hist[arg * nb_bins + bin], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * nb_bins; i += 256)
        if (sh_hist[i]) atomicAdd(&hist[i], (unsigned long long)sh_hist[i]);
}

extern "C" int b2_plop_entropy_hist(const float* x_old, const float* target, int B, int C, int64_t V, float max_entropy, int nb_bins,
                                    uint64_t* hist, b2_stream_t stream) {
    B2_CHECK_ARG(x_old && target && hist && B >= 1 && C >= 2 && C <= MAXC && V >= 1 && nb_bins >= 1 && nb_bins <= 1024 && max_entropy > 0.f);
    long long g = ((long long)B * V + 255) / 256, cap = (long long)num_sms() * 8;
    if (g > cap) g = cap;
    B2_LAUNCH(plop_hist_kernel, (int)g, 256, (size_t)C * nb_bins * sizeof(unsigned int), (cudaStream_t)stream, x_old, target, B, C, (long long)V,
              max_entropy, nb_bins, (unsigned long long*)hist);
    return B2_OK;
}

extern "C" int b2_plop_pseudo(const float* x, const float* x_old, const float* target, int B, int C, int D, int H, int W,
                              const float* thresholds, float max_entropy, float weight, float* dlogits, float* loss_out,
                              void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(x && x_old && target && thresholds && loss_out && scratch && C >= 2 && C <= MAXC && B >= 1 && W <= 4096);
    cudaStream_t st = (cudaStream_t)stream;
    const long long V = (long long)D * H * W;
    const int slabs = loss_slabs(B, V);
    int8_t* code = (int8_t*)scratch;
    float* f = (float*)((char*)scratch + align_up((size_t)B * V + 256));
    float* numden = f;
    int nz = (4 * num_sms()) / (cdiv(W, 32) * B);      // row chunks: fill the GPU
    if (nz > 32) nz = 32;
    if (nz > cdiv(D * H, 8)) nz = cdiv(D * H, 8);
    if (nz < 1) nz = 1;
    float* part = numden + (size_t)nz * B * W * 2;
    float* coef = part + (size_t)B * slabs * 4;
    dim3 g1(cdiv(W, 32), B, nz);
    B2_LAUNCH(plop_mask_kernel, g1, 256, 0, st, x_old, target, C, D, H, W, thresholds, max_entropy, code, numden);
    dim3 g2(slabs, B);
    B2_LAUNCH(plop_ce_reduce_kernel, g2, 256, 0, st, x, target, code, C, V, slabs, part);
    B2_LAUNCH(plop_finalize_kernel, 1, 256, 0, st, part, B * slabs, numden, nz, B, W, weight, coef, loss_out);
    if (dlogits) {
        long long total = (long long)B * V;
        long long g = (total + 255) / 256, cap = (long long)num_sms() * 16;
        if (g > cap) g = cap;
        B2_LAUNCH(plop_grad_kernel, (int)g, 256, 0, st, x, target, code, B, C, V, coef, dlogits);
    }
    return B2_OK;
}

extern "C" int b2_online_eval(const float* logits, const float* target, int B, int C, int64_t V, float* counts_out,
                              void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(logits && target && counts_out && scratch && C >= 2 && C <= MAXC && B >= 1 && B * (C - 1) * 3 <= 1024);
    cudaStream_t st = (cudaStream_t)stream;
    const int slabs = loss_slabs(B, V);
    float* part = (float*)scratch;
    dim3 grid(slabs, B);
    B2_LAUNCH(eval_reduce_kernel, grid, 256, 0, st, logits, target, C, (long long)V, slabs, part);
    B2_LAUNCH(eval_finalize_kernel, 1, 1024, 0, st, part, B, C, slabs, counts_out);
    return B2_OK;
}
