// infer.cu -- sliding-window inference aggregation.  Replaces the per-patch tensor arithmetic of nnunet's
// SegmentationNetwork._internal_predict_3D_3Dconv_tiled (un-vendored; reached from the reference's inference/predict.py:117-401
// and evaluation/evaluator.py) : softmax of the patch logits (inference_apply_nonlin = softmax_helper), optional un-mirroring of
// test-time-augmentation passes, multiplication with the Gaussian importance map, accumulation into the volume-sized class
// probability tensor and weight map -- one fused launch per predicted patch -- and the final normalisation + argmax.
#include "common.cuh"
#include "kernels.h"

namespace b2 {
constexpr int INF_MAXC = 8;

// agg[c][z0+z][y0+y][x0+x] += scale * gauss[z][y][x] * softmax_c(logits[:, fz(z), fy(y), fx(x)]);  wsum (+)= gauss when add_weight
__global__ void __launch_bounds__(256) sliding_accumulate_kernel(const float* __restrict__ logits, int C, int pd, int ph, int pw,
                                                                 const float* __restrict__ gauss, float* __restrict__ agg,
                                                                 float* __restrict__ wsum, int D, int H, int W, int z0, int y0, int x0,
                                                                 int flip, float scale, int add_weight) {
    pdl_grid_sync();
    const long long pv = (long long)pd * ph * pw, V = (long long)D * H * W;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < pv; i += (long long)gridDim.x * 256) {
        const int x = (int)(i % pw), y = (int)((i / pw) % ph), z = (int)(i / ((long long)pw * ph));
        // the network saw the patch flipped along the axes in `flip`: its output voxel (z, y, x) belongs to patch voxel (fz, fy, fx)
        const int sz = (flip & 1) ? pd - 1 - z : z, sy = (flip & 2) ? ph - 1 - y : y, sx = (flip & 4) ? pw - 1 - x : x;
        const long long src = ((long long)sz * ph + sy) * pw + sx;
        float v[INF_MAXC];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < INF_MAXC; ++c)
            if (c < C) { v[c] = logits[(long long)c * pv + src]; m = fmaxf(m, v[c]); }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < INF_MAXC; ++c)
            if (c < C) { v[c] = expf(v[c] - m); s += v[c]; }
        const float g = gauss ? gauss[i] : 1.f;
        const float f = scale * g / s;
        const long long dst = ((long long)(z0 + z) * H + (y0 + y)) * W + (x0 + x);
#pragma unroll
        for (int c = 0; c < INF_MAXC; ++c)
            if (c < C) agg[(long long)c * V + dst] += f * v[c];
        if (add_weight) wsum[dst] += g;
    }
}

// probs[c][v] = agg[c][v] / wsum[v] (in place); seg[v] = argmax_c
__global__ void __launch_bounds__(256) sliding_finalize_kernel(float* __restrict__ agg, const float* __restrict__ wsum, int C, long long V,
                                                               int32_t* __restrict__ seg) {
    pdl_grid_sync();
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < V; i += (long long)gridDim.x * 256) {
        const float inv = 1.f / wsum[i];
        float best = -INFINITY;
        int arg = 0;
        for (int c = 0; c < C; ++c) {
            const float p = agg[(long long)c * V + i] * inv;
            agg[(long long)c * V + i] = p;
            if (p > best) { best = p; arg = c; }
        }
        if (seg) seg[i] = arg;
    }
}
}  // namespace b2

using namespace b2;

extern "C" int b2_sliding_accumulate(const float* logits, int C, int pd, int ph, int pw, const float* gauss, float* agg, float* wsum,
                                     int D, int H, int W, int z0, int y0, int x0, int flip_mask, float scale, int add_weight,
                                     b2_stream_t stream) {
    B2_CHECK_ARG(logits && agg && wsum && C >= 1 && C <= INF_MAXC && pd >= 1 && ph >= 1 && pw >= 1);
    B2_CHECK_ARG(z0 >= 0 && y0 >= 0 && x0 >= 0 && z0 + pd <= D && y0 + ph <= H && x0 + pw <= W);
    const long long pv = (long long)pd * ph * pw;
    long long grid = (pv + 255) / 256, cap = (long long)num_sms() * 16;
    if (grid > cap) grid = cap;
    B2_LAUNCH(sliding_accumulate_kernel, (int)grid, 256, 0, (cudaStream_t)stream, logits, C, pd, ph, pw, gauss, agg, wsum, D, H, W, z0, y0, x0,
              flip_mask, scale, add_weight);
    return B2_OK;
}

extern "C" int b2_sliding_finalize(float* agg, const float* wsum, int C, int64_t V, int32_t* seg, b2_stream_t stream) {
    B2_CHECK_ARG(agg && wsum && C >= 1 && V >= 1);
    long long grid = (V + 255) / 256, cap = (long long)num_sms() * 16;
    if (grid > cap) grid = cap;
    B2_LAUNCH(sliding_finalize_kernel, (int)grid, 256, 0, (cudaStream_t)stream, agg, wsum, C, (long long)V, seg);
    return B2_OK;
}
