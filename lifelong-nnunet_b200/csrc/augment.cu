// augment.cu -- GPU-side patch pipeline: the work nnunet's DataLoader3D + batchgenerators'
// get_moreDA_augmentation do on ~12 CPU worker processes per GPU (reference call sites
// training/network_training/multihead/nnUNetTrainerMultiHead.py:505-511, 904-922; both packages are un-vendored, the
// semantics restated here are those of nnunet@77bc485 / batchgenerators 0.21 calling scipy.ndimage), with the preprocessed cases
// resident in HBM.  Random numbers are drawn on the HOST (b200unet/augment.py: one small parameter record per sample), the
// kernels apply them:
//   aug_crop_kernel       DataLoader3D: crop of the generator patch with constant padding (data 0, segmentation -1)
//   aug_prefilter_kernel  cubic B-spline prefilter along one axis (scipy spline_filter1d, mirror boundaries), in place
//   aug_resample_kernel   SpatialTransform: rotation / scaling about the crop centre; data = cubic B-spline interpolation
//                         (map_coordinates order 3, mode constant), segmentation = per-label linear interpolation + 0.5
//                         threshold (interpolate_img is_seg, order 1); samples without a transform are centre-cropped
//   aug_stats_*           per (sample, channel) mean / std / min / max, two-stage, fixed order
//   aug_pointwise_kernel  Gaussian noise (counter-based generator), multiplicative brightness, contrast, gamma (with
//                         retain_stats, optionally on the inverted image)
//   aug_blur_kernel       Gaussian blur along one axis (scipy gaussian_filter: reflect boundaries, truncate 4)
//   aug_lowres_*          SimulateLowResolutionTransform: nearest-neighbour resize down into an edge-padded volume (min / max by
//                         integer atomics), the prefilter above, cubic B-spline resize back with clipping (scipy zoom pair)
//   aug_finalize_kernel   MirrorTransform + RemoveLabelTransform(-1, 0) + DownsampleSegForDSTransform2 (order 0): writes the
//                         network input and every deep-supervision target in one pass
// All of it is HBM-bound elementwise / gather work on a few MB per batch.
#include "common.cuh"
#include "kernels.h"

namespace b2 {

struct AugCrops { b2_aug_case c[B2_AUG_MAX_SAMPLES]; };
struct AugSpatials { b2_aug_spatial_params s[B2_AUG_MAX_SAMPLES]; };
struct AugOps { b2_aug_op o[B2_AUG_MAX_BC]; };
struct AugBlurs { b2_aug_blur_taps k[B2_AUG_MAX_BC]; };
struct AugFinal {
    int32_t flips[B2_AUG_MAX_SAMPLES];
    float* targets[B2_AUG_MAX_SCALES];
    int32_t stride[B2_AUG_MAX_SCALES][3];
    int32_t n_scales;
};

// data[b][c][z][y][x] / seg[b][0][z][y][x] = case volume at lb + (z, y, x), constants outside; only the window [win_lo, win_hi) of
// the crop is written (a sample without a spatial transform only ever reads the centre window).
// grid = (plane chunks, gd, B * (C + 1)): one integer division per element (a flat index costs five, which made this kernel
// instruction-bound at 185 us per batch).
__global__ void __launch_bounds__(256) aug_crop_kernel(AugCrops cr, int C, int gd, int gh, int gw, float* __restrict__ data,
                                                       float* __restrict__ seg) {
    pdl_grid_sync();
    const int z = blockIdx.y, bc = blockIdx.z, b = bc / (C + 1), c = bc - b * (C + 1);
    const b2_aug_case& k = cr.c[b];
    if (z < k.win_lo[0] || z >= k.win_hi[0]) return;
    const int pidx = blockIdx.x * 256 + threadIdx.x;
    if (pidx >= gh * gw) return;
    const int y = pidx / gw, x = pidx - y * gw;
    if (y < k.win_lo[1] || y >= k.win_hi[1] || x < k.win_lo[2] || x >= k.win_hi[2]) return;
    const int sz = k.lb[0] + z, sy = k.lb[1] + y, sx = k.lb[2] + x;
    const bool in = sz >= 0 && sz < k.dhw[0] && sy >= 0 && sy < k.dhw[1] && sx >= 0 && sx < k.dhw[2];
    float v = c == C ? -1.f : 0.f;
    if (in) v = k.volume[(((long long)c * k.dhw[0] + sz) * k.dhw[1] + sy) * k.dhw[2] + sx];
    const long long gv = (long long)gd * gh * gw, o = ((long long)z * gh + y) * gw + x;
    if (c == C) seg[(long long)b * gv + o] = v;
    else data[((long long)b * C + c) * gv + o] = v;
}

// scipy ni_splines.c, order 3 (one pole z = sqrt(3) - 2), mirror boundary: gain, exact causal initialisation over the whole line,
// causal and anticausal recursions.  One thread per line of a modified sample.
__global__ void __launch_bounds__(128) aug_prefilter_kernel(AugSpatials sp, int B, int C, int gd, int gh, int gw, int axis,
                                                            float* __restrict__ data) {
    pdl_grid_sync();
    const int dims[3] = {gd, gh, gw};
    const int n = dims[axis];
    const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;          // the two other axes
    const long long strides[3] = {(long long)gh * gw, gw, 1};
    const long long lines_per = (long long)dims[a1] * dims[a2], gv = (long long)gd * gh * gw;
    const long long total = (long long)B * C * lines_per;
    const float z = -0.26794919243112270647f;                            // sqrt(3) - 2
    const float gain = (1.f - z) * (1.f - 1.f / z);
    for (long long i = (long long)blockIdx.x * 128 + threadIdx.x; i < total; i += (long long)gridDim.x * 128) {
        const int b = (int)(i / (C * lines_per));
        if (!sp.s[b].modified || n < 2) continue;
        const long long r = i % lines_per;
        float* c = data + (i / lines_per) * gv + (r / dims[a2]) * strides[a1] + (r % dims[a2]) * strides[a2];
        const long long st = strides[axis];
        const float zn1 = powf(-z, (float)(n - 1)) * (((n - 1) & 1) ? -1.f : 1.f);
        const float last = c[(long long)(n - 1) * st] * gain;
        float c0 = c[0] * gain + zn1 * last, zi = z;
        // terms beyond ~28 samples are below fp32 resolution (|z|^28 = 1e-16)
        for (int k = 1; k < n - 1; ++k) {
            c0 += zi * (c[(long long)k * st] * gain + zn1 * c[(long long)(n - 1 - k) * st] * gain);
            zi *= z;
            if (fabsf(zi) < 1e-16f) break;
        }
        float prev = c0 / (1.f - zn1 * zn1);
        c[0] = prev;
        for (int k = 1; k < n; ++k) {
            prev = c[(long long)k * st] * gain + z * prev;
            c[(long long)k * st] = prev;
        }
        float nxt = (z * c[(long long)(n - 2) * st] + prev) * z / (z * z - 1.f);
        c[(long long)(n - 1) * st] = nxt;
        for (int k = n - 2; k >= 0; --k) {
            nxt = z * (nxt - c[(long long)k * st]);
            c[(long long)k * st] = nxt;
        }
    }
}

__device__ __forceinline__ int aug_mirror(int i, int n) {      // scipy NI_EXTEND_MIRROR index map (d c b | a b c d | c b a)
    if (n <= 1) return 0;
    const int s2 = 2 * n - 2;
    if (i < 0) {
        i = s2 * (-i / s2) + i;
        return i <= 1 - n ? i + s2 : -i;
    }
    if (i >= n) {
        i -= s2 * (i / s2);
        if (i >= n) i = s2 - i;
    }
    return i;
}
__device__ __forceinline__ void bspline3(float t, float (&w)[4]) {
    const float u = 1.f - t;
    w[0] = u * u * u * (1.f / 6.f);
    w[1] = (t * t * (3.f * t - 6.f) + 4.f) * (1.f / 6.f);
    w[2] = (u * u * (3.f * u - 6.f) + 4.f) * (1.f / 6.f);
    w[3] = t * t * t * (1.f / 6.f);
}

// out voxel o (zero-centred: o - (p - 1) / 2) -> source coordinate ctr + M (o - (p-1)/2) in the crop; data: cubic B-spline of the
// prefiltered crop (constant 0 outside [0, n-1]); seg: the largest label whose linear-interpolated indicator is >= 0.5, else 0
__global__ void __launch_bounds__(256) aug_resample_kernel(AugSpatials sp, int B, int C, int gd, int gh, int gw, int pd, int ph, int pw,
                                                           const float* __restrict__ coef, const float* __restrict__ seg,
                                                           float* __restrict__ data_out, float* __restrict__ seg_out) {
    pdl_grid_sync();
    const long long pv = (long long)pd * ph * pw, gv = (long long)gd * gh * gw, total = (long long)B * pv;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int x = (int)(i % pw), y = (int)((i / pw) % ph), z = (int)((i / ((long long)pw * ph)) % pd), b = (int)(i / pv);
        const b2_aug_spatial_params& s = sp.s[b];
        const long long o = ((long long)z * ph + y) * pw + x;
        if (!s.modified) {
            const long long src = ((long long)(z + s.lb[0]) * gh + (y + s.lb[1])) * gw + (x + s.lb[2]);
            for (int c = 0; c < C; ++c) data_out[((long long)b * C + c) * pv + o] = coef[((long long)b * C + c) * gv + src];
            seg_out[(long long)b * pv + o] = seg[(long long)b * gv + src];
            continue;
        }
        const float cz = z - 0.5f * (pd - 1), cy = y - 0.5f * (ph - 1), cx = x - 0.5f * (pw - 1);
        const float fz = s.ctr[0] + s.m[0] * cz + s.m[1] * cy + s.m[2] * cx;
        const float fy = s.ctr[1] + s.m[3] * cz + s.m[4] * cy + s.m[5] * cx;
        const float fx = s.ctr[2] + s.m[6] * cz + s.m[7] * cy + s.m[8] * cx;
        const bool inside = fz >= 0.f && fz <= gd - 1 && fy >= 0.f && fy <= gh - 1 && fx >= 0.f && fx <= gw - 1;
        if (!inside) {
            for (int c = 0; c < C; ++c) data_out[((long long)b * C + c) * pv + o] = 0.f;
            seg_out[(long long)b * pv + o] = 0.f;
            continue;
        }
        const int iz = (int)floorf(fz), iy = (int)floorf(fy), ix = (int)floorf(fx);
        const float tz = fz - iz, ty = fy - iy, tx = fx - ix;
        float wz[4], wy[4], wx[4];
        bspline3(tz, wz); bspline3(ty, wy); bspline3(tx, wx);
        int jz[4], jy[4], jx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            jz[k] = aug_mirror(iz - 1 + k, gd); jy[k] = aug_mirror(iy - 1 + k, gh); jx[k] = aug_mirror(ix - 1 + k, gw);
        }
        for (int c = 0; c < C; ++c) {
            const float* v = coef + ((long long)b * C + c) * gv;
            float acc = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float az = 0.f;
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    const float* row = v + ((long long)jz[a] * gh + jy[bb]) * gw;
                    az += wy[bb] * (wx[0] * row[jx[0]] + wx[1] * row[jx[1]] + wx[2] * row[jx[2]] + wx[3] * row[jx[3]]);
                }
                acc += wz[a] * az;
            }
            data_out[((long long)b * C + c) * pv + o] = acc;
        }
        // segmentation: trilinear indicator sums per label present among the 8 neighbours
        const int z1 = aug_mirror(iz + 1, gd), y1 = aug_mirror(iy + 1, gh), x1 = aug_mirror(ix + 1, gw);
        const float* sv = seg + (long long)b * gv;
        float lab[8], wt[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int zz = (k & 4) ? z1 : iz, yy = (k & 2) ? y1 : iy, xx = (k & 1) ? x1 : ix;
            lab[k] = sv[((long long)zz * gh + yy) * gw + xx];
            wt[k] = ((k & 4) ? tz : 1.f - tz) * ((k & 2) ? ty : 1.f - ty) * ((k & 1) ? tx : 1.f - tx);
        }
        float best = 0.f;
        bool found = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += lab[j] == lab[k] ? wt[j] : 0.f;
            if (sum >= 0.5f && (!found || lab[k] > best)) { best = lab[k]; found = true; }
        }
        seg_out[(long long)b * pv + o] = found ? best : 0.f;
    }
}

// ---- per (sample, channel) statistics ---------------------------------------------------------------------------------
constexpr int AUG_STAT_CHUNK = 16384;
__global__ void __launch_bounds__(256) aug_stats_part_kernel(const float* __restrict__ x, long long V, int chunks, double* __restrict__ part) {
    pdl_grid_sync();
    const int bc = blockIdx.y, ch = blockIdx.x;
    const float* p = x + (long long)bc * V;
    const long long lo = (long long)ch * AUG_STAT_CHUNK, hi = lo + AUG_STAT_CHUNK < V ? lo + AUG_STAT_CHUNK : V;
    double s = 0.0, ss = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    for (long long i = lo + threadIdx.x; i < hi; i += 256) {
        const float v = p[i];
        s += v; ss += (double)v * v;
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    __shared__ double sh[4][8];
    s = warp_sum(s); ss = warp_sum(ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[0][w] = s; sh[1][w] = ss; sh[2][w] = mn; sh[3][w] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, c = INFINITY, d = -INFINITY;
        for (int k = 0; k < 8; ++k) { a += sh[0][k]; b += sh[1][k]; c = fmin(c, sh[2][k]); d = fmax(d, sh[3][k]); }
        double* o = part + ((long long)bc * chunks + ch) * 4;
        o[0] = a; o[1] = b; o[2] = c; o[3] = d;
    }
}
// stats[bc] = {mean, std (ddof 0), min, max}
__global__ void __launch_bounds__(32) aug_stats_final_kernel(const double* __restrict__ part, int chunks, long long V, float* __restrict__ stats) {
    pdl_grid_sync();
    const int bc = blockIdx.x;
    double s = 0.0, ss = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int k = threadIdx.x; k < chunks; k += 32) {
        const double* p = part + ((long long)bc * chunks + k) * 4;
        s += p[0]; ss += p[1]; mn = fmin(mn, p[2]); mx = fmax(mx, p[3]);
    }
    s = warp_sum(s); ss = warp_sum(ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (threadIdx.x == 0) {
        const double mean = s / (double)V, var = ss / (double)V - mean * mean;
        stats[bc * 4 + 0] = (float)mean; stats[bc * 4 + 1] = (float)sqrt(var > 0.0 ? var : 0.0);
        stats[bc * 4 + 2] = (float)mn; stats[bc * 4 + 3] = (float)mx;
    }
}

// counter-based standard normal: splitmix64 of (seed, element index) -> two 24-bit uniforms -> Box-Muller
__device__ __forceinline__ float aug_normal(unsigned long long seed, unsigned long long idx) {
    unsigned long long h = seed + (idx + 1ULL) * 0x9E3779B97F4A7C15ULL;
    h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ULL;
    h = (h ^ (h >> 27)) * 0x94D049BB133111EBULL;
    h ^= h >> 31;
    const float u1 = ((float)(unsigned)(h >> 40) + 0.5f) * (1.f / 16777216.f);
    const float u2 = ((float)(unsigned)((h >> 16) & 0xFFFFFFULL) + 0.5f) * (1.f / 16777216.f);
    return sqrtf(-2.f * logf(u1)) * cosf(6.283185307179586f * u2);
}

__global__ void __launch_bounds__(256) aug_pointwise_kernel(AugOps ops, long long V, float* __restrict__ x, const float* __restrict__ sa,
                                                            const float* __restrict__ sb, unsigned long long seed) {
    pdl_grid_sync();
    const int bc = blockIdx.y;
    const b2_aug_op op = ops.o[bc];
    if (op.op == B2_AUG_NONE) return;
    float* p = x + (long long)bc * V;
    float mean = 0.f, sd = 0.f, mn = 0.f, mx = 0.f, mean2 = 0.f, sd2 = 0.f;
    if (sa) { mean = sa[bc * 4]; sd = sa[bc * 4 + 1]; mn = sa[bc * 4 + 2]; mx = sa[bc * 4 + 3]; }
    if (sb) { mean2 = sb[bc * 4]; sd2 = sb[bc * 4 + 1]; }
    const bool inv = op.p[1] != 0.f;
    if (inv && (op.op == B2_AUG_GAMMA_A || op.op == B2_AUG_GAMMA_B)) {     // statistics of the negated image
        const float t = mn;
        mean = -mean; mn = -mx; mx = -t; mean2 = -mean2;
    }
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < V; i += (long long)gridDim.x * 256) {
        float v = p[i];
        switch (op.op) {
            case B2_AUG_NOISE: v += op.p[0] * aug_normal(seed, (unsigned long long)bc * (unsigned long long)V + (unsigned long long)i); break;
            case B2_AUG_MUL: v *= op.p[0]; break;
            case B2_AUG_CONTRAST: v = fminf(fmaxf((v - mean) * op.p[0] + mean, mn), mx); break;
            case B2_AUG_GAMMA_A: {          // ((x - min) / (range + 1e-7)) ^ gamma * range + min, on (-x) when inverting
                if (inv) v = -v;
                const float rng = mx - mn;
                v = powf((v - mn) / (rng + 1e-7f), op.p[0]) * (rng + 1e-7f) + mn;
                if (inv) v = -v;
                break;
            }
            case B2_AUG_GAMMA_B: {          // retain_stats: back to the mean / std the channel had before the gamma curve
                if (inv) v = -v;
                v = (v - mean2) / (sd2 + 1e-8f) * sd + mean;
                if (inv) v = -v;
                break;
            }
            default: break;
        }
        p[i] = v;
    }
}

// correlate1d with a symmetric kernel of radius r along `axis`, reflect boundaries (d c b a | a b c d | d c b a)
__global__ void __launch_bounds__(256) aug_blur_kernel(AugBlurs bl, int pd, int ph, int pw, int axis, const float* __restrict__ src,
                                                       float* __restrict__ dst) {
    pdl_grid_sync();
    const int bc = blockIdx.y;
    const b2_aug_blur_taps k = bl.k[bc];
    const long long V = (long long)pd * ph * pw;
    const float* s = src + (long long)bc * V;
    float* d = dst + (long long)bc * V;
    const int dims[3] = {pd, ph, pw};
    const long long strides[3] = {(long long)ph * pw, pw, 1};
    const int n = dims[axis];
    const long long st = strides[axis];
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < V; i += (long long)gridDim.x * 256) {
        if (k.radius == 0) { d[i] = s[i]; continue; }
        const int pos = (int)((i / st) % n);
        const long long base = i - (long long)pos * st;
        float acc = 0.f;
        for (int t = -k.radius; t <= k.radius; ++t) {
            int j = pos + t;
            while (j < 0 || j >= n) j = j < 0 ? -j - 1 : 2 * n - 1 - j;
            acc += k.w[t < 0 ? -t : t] * s[base + (long long)j * st];
        }
        d[i] = acc;
    }
}

// ---- SimulateLowResolutionTransform (batchgenerators augment_linear_downsampling_scipy: skimage resize order 0 down, order 3 up,
// mode 'edge' == scipy.ndimage.zoom(grid_mode=True, mode='nearest')), one (sample, channel) volume per call -----------------------
constexpr int LR_PAD = 12;      // scipy pre-pads by 12 samples ('nearest') before the spline prefilter
__device__ __forceinline__ unsigned int lr_encode(float f) {     // order-preserving map float -> uint (min / max by integer atomics)
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float lr_decode(unsigned int u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// padded[i][j][k] = vol[nearest source of clamp(i - 12, 0, t - 1) ...]: output voxel q reads input floor((q + 0.5) * p / t)
__global__ void __launch_bounds__(256) aug_lowres_down_kernel(const float* __restrict__ vol, int pd, int ph, int pw, int td, int th, int tw,
                                                              float* __restrict__ padded, unsigned int* __restrict__ minmax) {
    pdl_grid_sync();
    const int Ph = th + 2 * LR_PAD, Pw = tw + 2 * LR_PAD;
    const int i = blockIdx.y, pidx = blockIdx.x * 256 + threadIdx.x;
    float v = 0.f;
    const bool on = pidx < Ph * Pw;
    if (on) {
        const int j = pidx / Pw, k = pidx - j * Pw;
        const int qz = min(max(i - LR_PAD, 0), td - 1), qy = min(max(j - LR_PAD, 0), th - 1), qx = min(max(k - LR_PAD, 0), tw - 1);
        const int sz = min((int)floor((qz + 0.5) * ((double)pd / td)), pd - 1), sy = min((int)floor((qy + 0.5) * ((double)ph / th)), ph - 1),
                  sx = min((int)floor((qx + 0.5) * ((double)pw / tw)), pw - 1);
        v = vol[((long long)sz * ph + sy) * pw + sx];
        padded[((long long)i * Ph + j) * Pw + k] = v;
    }
    float mn = on ? v : INFINITY, mx = on ? v : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0 && mn <= mx) {       // min / max are order-independent: integer atomics stay reproducible
        atomicMin(&minmax[0], lr_encode(mn));
        atomicMax(&minmax[1], lr_encode(mx));
    }
}

// out[z][y][x] = clip(cubic B-spline of the prefiltered padded volume at (o + 0.5) * t / p - 0.5 + 12, min, max)
__global__ void __launch_bounds__(256) aug_lowres_up_kernel(const float* __restrict__ coef, int td, int th, int tw, int pd, int ph, int pw,
                                                            const unsigned int* __restrict__ minmax, float* __restrict__ out) {
    pdl_grid_sync();
    const int Ph = th + 2 * LR_PAD, Pw = tw + 2 * LR_PAD;
    const int z = blockIdx.y, pidx = blockIdx.x * 256 + threadIdx.x;
    if (pidx >= ph * pw) return;
    const int y = pidx / pw, x = pidx - y * pw;
    const double cz = (z + 0.5) * ((double)td / pd) - 0.5 + LR_PAD, cy = (y + 0.5) * ((double)th / ph) - 0.5 + LR_PAD,
                 cx = (x + 0.5) * ((double)tw / pw) - 0.5 + LR_PAD;
    const int iz = (int)floor(cz), iy = (int)floor(cy), ix = (int)floor(cx);
    float wz[4], wy[4], wx[4];
    bspline3((float)(cz - iz), wz); bspline3((float)(cy - iy), wy); bspline3((float)(cx - ix), wx);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float az = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float* row = coef + ((long long)(iz - 1 + a) * Ph + (iy - 1 + b)) * Pw + ix - 1;
            az += wy[b] * (wx[0] * row[0] + wx[1] * row[1] + wx[2] * row[2] + wx[3] * row[3]);
        }
        acc += wz[a] * az;
    }
    const float lo = lr_decode(minmax[0]), hi = lr_decode(minmax[1]);
    out[((long long)z * ph + y) * pw + x] = fminf(fmaxf(acc, lo), hi);
}

// mirror + label clean-up + deep-supervision targets.  blockIdx.y selects the output: 0 = network input (one thread per voxel),
// 1 + k = deep-supervision target k (one thread per TARGET voxel q, which reads source voxel stride * q + stride / 2 -- the order-0
// resize rule -- so no thread ever tests divisibility).
__global__ void __launch_bounds__(256) aug_finalize_kernel(AugFinal f, int B, int C, int pd, int ph, int pw, const float* __restrict__ data,
                                                           const float* __restrict__ seg, float* __restrict__ data_out) {
    pdl_grid_sync();
    const long long pv = (long long)pd * ph * pw;
    const int which = blockIdx.y;
    int kz = 1, ky = 1, kx = 1;
    if (which > 0) { kz = f.stride[which - 1][0]; ky = f.stride[which - 1][1]; kx = f.stride[which - 1][2]; }
    const int td = pd / kz, th = ph / ky, tw = pw / kx;
    const long long tv = (long long)td * th * tw, total = (long long)B * tv;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int qx = (int)(i % tw), qy = (int)((i / tw) % th), qz = (int)((i / ((long long)tw * th)) % td), b = (int)(i / tv);
        const int z = qz * kz + kz / 2, y = qy * ky + ky / 2, x = qx * kx + kx / 2;
        const int fl = f.flips[b];
        const int sz = (fl & 1) ? pd - 1 - z : z, sy = (fl & 2) ? ph - 1 - y : y, sx = (fl & 4) ? pw - 1 - x : x;
        const long long src = ((long long)sz * ph + sy) * pw + sx;
        if (which == 0) {
            for (int c = 0; c < C; ++c) data_out[((long long)b * C + c) * pv + i - (long long)b * tv] = data[((long long)b * C + c) * pv + src];
        } else {
            float lab = seg[(long long)b * pv + src];
            if (lab == -1.f) lab = 0.f;
            f.targets[which - 1][i] = lab;
        }
    }
}
}  // namespace b2

using namespace b2;

static int grid1d(long long n, int block) {
    long long g = (n + block - 1) / block, cap = (long long)num_sms() * 16;
    return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

extern "C" int b2_aug_crop(const b2_aug_case* cases_host, int B, int C, const int32_t gdhw[3], float* data, float* seg, b2_stream_t stream) {
    B2_CHECK_ARG(cases_host && data && seg && B >= 1 && B <= B2_AUG_MAX_SAMPLES && C >= 1 && gdhw[0] >= 1 && gdhw[1] >= 1 && gdhw[2] >= 1);
    AugCrops cr;
    memset(&cr, 0, sizeof(cr));
    for (int b = 0; b < B; ++b) {
        B2_CHECK_ARG(cases_host[b].volume != nullptr);
        cr.c[b] = cases_host[b];
    }
    for (int b = 0; b < B; ++b)
        for (int a = 0; a < 3; ++a) B2_CHECK_ARG(cases_host[b].win_lo[a] >= 0 && cases_host[b].win_hi[a] <= gdhw[a]);
    B2_CHECK_ARG(gdhw[0] <= 65535 && B * (C + 1) <= 65535);
    B2_LAUNCH(aug_crop_kernel, dim3(cdiv((long long)gdhw[1] * gdhw[2], 256), gdhw[0], B * (C + 1)), 256, 0, stream, cr, C, gdhw[0], gdhw[1],
              gdhw[2], data, seg);
    return B2_OK;
}

extern "C" int b2_aug_spatial(const b2_aug_spatial_params* tf_host, int B, int C, const int32_t gdhw[3], const int32_t pdhw[3], float* crop_data,
                              const float* crop_seg, float* data_out, float* seg_out, b2_stream_t stream) {
    B2_CHECK_ARG(tf_host && crop_data && crop_seg && data_out && seg_out && B >= 1 && B <= B2_AUG_MAX_SAMPLES && C >= 1);
    AugSpatials sp;
    memset(&sp, 0, sizeof(sp));
    bool any = false;
    for (int b = 0; b < B; ++b) {
        sp.s[b] = tf_host[b];
        any = any || tf_host[b].modified;
        if (!tf_host[b].modified)
            for (int a = 0; a < 3; ++a) B2_CHECK_ARG(tf_host[b].lb[a] >= 0 && tf_host[b].lb[a] + pdhw[a] <= gdhw[a]);
    }
    if (any)
        for (int axis = 0; axis < 3; ++axis) {
            const long long lines = (long long)B * C * gdhw[0] * gdhw[1] * gdhw[2] / gdhw[axis];
            B2_LAUNCH(aug_prefilter_kernel, grid1d(lines, 128), 128, 0, stream, sp, B, C, gdhw[0], gdhw[1], gdhw[2], axis, crop_data);
        }
    const long long total = (long long)B * pdhw[0] * pdhw[1] * pdhw[2];
    B2_LAUNCH(aug_resample_kernel, grid1d(total, 256), 256, 0, stream, sp, B, C, gdhw[0], gdhw[1], gdhw[2], pdhw[0], pdhw[1], pdhw[2],
              (const float*)crop_data, crop_seg, data_out, seg_out);
    return B2_OK;
}

extern "C" size_t b2_aug_stats_scratch_bytes(int BC, int64_t V) {
    return (size_t)BC * (size_t)cdiv(V, AUG_STAT_CHUNK) * 4 * sizeof(double);
}

extern "C" int b2_aug_stats(const float* data, int BC, int64_t V, float* stats_out, void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(data && stats_out && scratch && BC >= 1 && BC <= B2_AUG_MAX_BC && V >= 1);
    const int chunks = cdiv(V, AUG_STAT_CHUNK);
    B2_LAUNCH(aug_stats_part_kernel, dim3(chunks, BC), 256, 0, stream, data, (long long)V, chunks, (double*)scratch);
    B2_LAUNCH(aug_stats_final_kernel, BC, 32, 0, stream, (const double*)scratch, chunks, (long long)V, stats_out);
    return B2_OK;
}

extern "C" int b2_aug_pointwise(const b2_aug_op* ops_host, int BC, int64_t V, float* data, const float* stats_a, const float* stats_b,
                                uint64_t seed, b2_stream_t stream) {
    B2_CHECK_ARG(ops_host && data && BC >= 1 && BC <= B2_AUG_MAX_BC && V >= 1);
    AugOps ops;
    memset(&ops, 0, sizeof(ops));
    for (int i = 0; i < BC; ++i) {
        ops.o[i] = ops_host[i];
        const int op = ops_host[i].op;
        B2_CHECK_ARG(op >= B2_AUG_NONE && op <= B2_AUG_GAMMA_B);
        if (op == B2_AUG_CONTRAST || op == B2_AUG_GAMMA_A || op == B2_AUG_GAMMA_B) B2_CHECK_ARG(stats_a != nullptr);
        if (op == B2_AUG_GAMMA_B) B2_CHECK_ARG(stats_b != nullptr);
    }
    int gx = grid1d(V, 256);
    if (gx > 1024) gx = 1024;
    B2_LAUNCH(aug_pointwise_kernel, dim3(gx, BC), 256, 0, stream, ops, (long long)V, data, stats_a, stats_b, (unsigned long long)seed);
    return B2_OK;
}

extern "C" int b2_aug_blur(const b2_aug_blur_taps* kernels_host, int BC, const int32_t pdhw[3], float* data, float* tmp, b2_stream_t stream) {
    B2_CHECK_ARG(kernels_host && data && tmp && BC >= 1 && BC <= B2_AUG_MAX_BC);
    AugBlurs bl;
    memset(&bl, 0, sizeof(bl));
    for (int i = 0; i < BC; ++i) {
        B2_CHECK_ARG(kernels_host[i].radius >= 0 && kernels_host[i].radius <= B2_AUG_MAX_RADIUS);
        bl.k[i] = kernels_host[i];
    }
    const long long V = (long long)pdhw[0] * pdhw[1] * pdhw[2];
    int gx = grid1d(V, 256);
    if (gx > 1024) gx = 1024;
    // three passes: data -> tmp -> data -> tmp, then one copy-shaped pass brings the result back into `data`
    B2_LAUNCH(aug_blur_kernel, dim3(gx, BC), 256, 0, stream, bl, pdhw[0], pdhw[1], pdhw[2], 0, (const float*)data, tmp);
    B2_LAUNCH(aug_blur_kernel, dim3(gx, BC), 256, 0, stream, bl, pdhw[0], pdhw[1], pdhw[2], 1, (const float*)tmp, data);
    B2_LAUNCH(aug_blur_kernel, dim3(gx, BC), 256, 0, stream, bl, pdhw[0], pdhw[1], pdhw[2], 2, (const float*)data, tmp);
    B2_CUDA(cudaMemcpyAsync(data, tmp, (size_t)BC * V * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return B2_OK;
}

extern "C" int b2_aug_finalize(const int32_t* flips_host, int B, int C, const int32_t pdhw[3], const float* data, const float* seg,
                               float* data_out, float* const* targets_host, const int32_t* strides_host, int n_scales, b2_stream_t stream) {
    B2_CHECK_ARG(flips_host && data && seg && data_out && targets_host && strides_host && B >= 1 && B <= B2_AUG_MAX_SAMPLES && C >= 1);
    B2_CHECK_ARG(n_scales >= 1 && n_scales <= B2_AUG_MAX_SCALES);
    AugFinal f;
    memset(&f, 0, sizeof(f));
    for (int b = 0; b < B; ++b) f.flips[b] = flips_host[b];
    f.n_scales = n_scales;
    for (int k = 0; k < n_scales; ++k) {
        B2_CHECK_ARG(targets_host[k] != nullptr);
        f.targets[k] = targets_host[k];
        for (int a = 0; a < 3; ++a) {
            const int s = strides_host[k * 3 + a];
            B2_CHECK_ARG(s >= 1 && pdhw[a] % s == 0);
            f.stride[k][a] = s;
        }
    }
    const long long total = (long long)B * pdhw[0] * pdhw[1] * pdhw[2];
    int gx = grid1d(total, 256);
    if (gx > 2048) gx = 2048;
    B2_LAUNCH(aug_finalize_kernel, dim3(gx, 1 + n_scales), 256, 0, stream, f, B, C, pdhw[0], pdhw[1], pdhw[2], data, seg, data_out);
    return B2_OK;
}

extern "C" size_t b2_aug_lowres_scratch_bytes(const int32_t pdhw[3]) {
    return ((size_t)(pdhw[0] + 2 * LR_PAD) * (pdhw[1] + 2 * LR_PAD) * (pdhw[2] + 2 * LR_PAD)) * sizeof(float) + 256;
}

extern "C" int b2_aug_lowres(float* vol, const int32_t pdhw[3], const int32_t tdhw[3], void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(vol && scratch);
    for (int a = 0; a < 3; ++a) B2_CHECK_ARG(tdhw[a] >= 2 && tdhw[a] <= pdhw[a] && pdhw[a] <= 65535);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int* minmax = (unsigned int*)scratch;
    float* padded = (float*)((char*)scratch + 256);
    const int Pd = tdhw[0] + 2 * LR_PAD, Ph = tdhw[1] + 2 * LR_PAD, Pw = tdhw[2] + 2 * LR_PAD;
    B2_CUDA(cudaMemsetAsync(minmax, 0xFF, 4, st));
    B2_CUDA(cudaMemsetAsync(minmax + 1, 0x00, 4, st));
    B2_LAUNCH(aug_lowres_down_kernel, dim3(cdiv((long long)Ph * Pw, 256), Pd), 256, 0, st, (const float*)vol, pdhw[0], pdhw[1], pdhw[2], tdhw[0],
              tdhw[1], tdhw[2], padded, minmax);
    AugSpatials sp;
    memset(&sp, 0, sizeof(sp));
    sp.s[0].modified = 1;
    for (int axis = 0; axis < 3; ++axis) {
        const long long lines = (long long)Pd * Ph * Pw / (axis == 0 ? Pd : axis == 1 ? Ph : Pw);
        B2_LAUNCH(aug_prefilter_kernel, grid1d(lines, 128), 128, 0, st, sp, 1, 1, Pd, Ph, Pw, axis, padded);
    }
    B2_LAUNCH(aug_lowres_up_kernel, dim3(cdiv((long long)pdhw[1] * pdhw[2], 256), pdhw[0]), 256, 0, st, (const float*)padded, tdhw[0], tdhw[1],
              tdhw[2], pdhw[0], pdhw[1], pdhw[2], (const unsigned int*)minmax, vol);
    return B2_OK;
}
