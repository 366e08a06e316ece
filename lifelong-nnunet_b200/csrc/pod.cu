// pod.cu -- local POD (Pooled Outputs Distillation) between a student and a teacher activation, on NDHWC tensors.
//
// Restates reference nnunet_ext/training/loss_functions/embeddings.py:3-42 (pod_embed / local_POD), including its
// quirks: for scale s the tile extent is e_s = W // 2^s and the tile origins are
// range(0, W - e_s, e_s) on BOTH axes (so scale 0 has no tile, scale 1 one tile, scale 2 3x3 tiles); the depth axis
// of a 5-D tensor is carried along un-pooled; tiles must be square (H == W).  With P = cat over tiles of
// [row-means ; col-means] (channel-concatenated), the layer value is mean_{b, c2 in 2C, d} || P - P_old ||_2 over the
// concatenated tile axis.  Means are linear, so everything is computed on e = h - h_old in one sweep per direction:
//   sq_row[b,d,c] = sum_s sum_{y < n_s e_s} sum_{k < n_s} ( mean_{x in seg k} e[y,x] )^2        (first C channels)
//   sq_col[b,d,c] = sum_s sum_{x < n_s e_s} sum_{k < n_s} ( mean_{y in seg k} e[y,x] )^2        (last C channels)
//   value = ( sum sqrt(sq_row) + sum sqrt(sq_col) ) / (B * 2C * D)
// Value only: both operands are detached in the reference (plop:348-352, Q7).  HBM/L2-bound; ordered reductions.
#include "common.cuh"
#include "kernels.h"

namespace b2 {

constexpr int MAXS = 6;
constexpr int POD_MAXZ = 16;
struct PodGeom {
    int B, D, H, W, C, pitch_a, pitch_b;
    int nscale;
    int e[MAXS], n[MAXS];
};

// block = (32 channels) x (8 lanes); grid = (B*D, ceil(C/32), Z): block z owns the rows / columns y, x = lane + 8 (z + Z k);
// out[((bd*C + c)*Z + z)*2 + {0,1}] = partial {sum of squared row means, sum of squared column means}; pod_sqrt_kernel adds the
// Z partials in order and takes the roots.  (Round 1 had no z split: 96 blocks for a 48x192x192 layer, 0.3-0.8 ms per layer.)
template <typename T>
__global__ void __launch_bounds__(256) pod_kernel(PodGeom g, const T* __restrict__ a, const T* __restrict__ b,
                                                  float* __restrict__ out) {
    pdl_grid_sync();
    __shared__ float sh[8][32][2];
    const int bd = blockIdx.x;
    const int c = blockIdx.y * 32 + (threadIdx.x & 31), lane = threadIdx.x >> 5;
    float srow = 0.f, scol = 0.f;
    if (c < g.C) {
        const T* pa = a + (long long)bd * g.H * g.W * g.pitch_a + c;
        const T* pb = b + (long long)bd * g.H * g.W * g.pitch_b + c;
        for (int s = 0; s < g.nscale; ++s) {
            const int e = g.e[s], n = g.n[s];
            const float inv = 1.f / (float)e;
            // rows: y in [0, n*e), segments k along x
            for (int y = lane + 8 * (int)blockIdx.z; y < n * e; y += 8 * (int)gridDim.z)
                for (int k = 0; k < n; ++k) {
                    float sum = 0.f;
                    for (int x = k * e; x < (k + 1) * e; ++x) {
                        const long long o = (long long)y * g.W + x;
                        sum += to_f(pa[o * g.pitch_a]) - to_f(pb[o * g.pitch_b]);
                    }
                    const float m = sum * inv;
                    srow += m * m;
                }
            // cols: x in [0, n*e), segments k along y
            for (int x = lane + 8 * (int)blockIdx.z; x < n * e; x += 8 * (int)gridDim.z)
                for (int k = 0; k < n; ++k) {
                    float sum = 0.f;
                    for (int y = k * e; y < (k + 1) * e; ++y) {
                        const long long o = (long long)y * g.W + x;
                        sum += to_f(pa[o * g.pitch_a]) - to_f(pb[o * g.pitch_b]);
                    }
                    const float m = sum * inv;
                    scol += m * m;
                }
        }
    }
    sh[lane][threadIdx.x & 31][0] = srow;
    sh[lane][threadIdx.x & 31][1] = scol;
    __syncthreads();
    if (lane == 0 && c < g.C) {
        float r = 0.f, q = 0.f;
        for (int l = 0; l < 8; ++l) { r += sh[l][threadIdx.x][0]; q += sh[l][threadIdx.x][1]; }
        out[(((long long)bd * g.C + c) * gridDim.z + blockIdx.z) * 2] = r;
        out[(((long long)bd * g.C + c) * gridDim.z + blockIdx.z) * 2 + 1] = q;
    }
}

// roots[i*2 + {0,1}] = sqrt(sum_z part[(i*Z + z)*2 + {0,1}])
__global__ void __launch_bounds__(256) pod_sqrt_kernel(const float* __restrict__ part, long long n, int Z, float* __restrict__ roots) {
    pdl_grid_sync();
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n * 2) return;
    const long long e = i >> 1;
    const int k = (int)(i & 1);
    float s = 0.f;
    for (int z = 0; z < Z; ++z) s += part[(e * Z + z) * 2 + k];
    roots[i] = sqrtf(s);
}

__global__ void __launch_bounds__(256) pod_finalize_kernel(const float* __restrict__ part, long long n, double scale,
                                                           float* __restrict__ out) {
    pdl_grid_sync();
    __shared__ double red[32];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += 256) s += (double)part[i];
    double r = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = (float)(r * scale);
}

}  // namespace b2

using namespace b2;

extern "C" size_t b2_pod_scratch_bytes(const b2_act_view* a, int scales) {
    (void)scales;
    if (!a) return 0;
    return align_up((size_t)a->n * a->d * a->c * 2 * (POD_MAXZ + 1) * sizeof(float) + 256);
}

extern "C" int b2_pod_local(const b2_act_view* a, const b2_act_view* a_old, int scales, float* value_out, void* scratch,
                            b2_stream_t stream) {
    B2_CHECK_ARG(a && a_old && value_out && scratch && scales >= 1 && scales <= MAXS);
    B2_CHECK_ARG(a->n == a_old->n && a->d == a_old->d && a->h == a_old->h && a->w == a_old->w && a->c == a_old->c && a->dtype == a_old->dtype);
    if (a->h != a->w) return fail(B2_EINVAL, "local_POD needs square tiles (H == W), like the reference (embeddings.py:5-7)%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    PodGeom g;
    g.B = a->n; g.D = a->d; g.H = a->h; g.W = a->w; g.C = a->c; g.pitch_a = a->pitch; g.pitch_b = a_old->pitch;
    g.nscale = 0;
    long long tiles = 0;
    for (int s = 0; s < scales; ++s) {
        const int e = a->w >> s;  // int(W / 2**s)
        if (e <= 0) return fail(B2_EINVAL, "too many POD scales for this extent (embeddings.py:26-27)%s", "");
        int n = 0;
        for (int i = 0; i < a->w - e; i += e) ++n;  // len(range(0, W - e, e))
        if (n > 0) { g.e[g.nscale] = e; g.n[g.nscale] = n; ++g.nscale; }
        tiles += (long long)n * n;
    }
    if (tiles == 0) return fail(B2_EINVAL, "local_POD produced no tile (reference would fail on zip(None))%s", "");
    const long long n = (long long)a->n * a->d * a->c * 2;
    int Z = (4 * num_sms()) / (a->n * a->d * cdiv(a->c, 32));
    if (Z > POD_MAXZ) Z = POD_MAXZ;
    if (Z > cdiv(a->w, 8)) Z = cdiv(a->w, 8);
    if (Z < 1) Z = 1;
    float* roots = (float*)scratch;
    float* part = roots + n;
    dim3 grid(a->n * a->d, cdiv(a->c, 32), Z);
    if (a->dtype == B2_F32) B2_LAUNCH(pod_kernel<float>, grid, 256, 0, st, g, (const float*)a->ptr, (const float*)a_old->ptr, part);
    else B2_LAUNCH(pod_kernel<__nv_bfloat16>, grid, 256, 0, st, g, (const __nv_bfloat16*)a->ptr, (const __nv_bfloat16*)a_old->ptr, part);
    B2_LAUNCH(pod_sqrt_kernel, cdiv(n, 256), 256, 0, st, (const float*)part, n / 2, Z, roots);
    B2_LAUNCH(pod_finalize_kernel, 1, 256, 0, st, (const float*)roots, n, 1.0 / (double)n, value_out);
    return B2_OK;
}
