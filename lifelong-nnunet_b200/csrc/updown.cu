// updown.cu -- kernel==stride transposed convolution (decoder up-sampling), 1x1x1 segmentation heads
// (K4) and the NCDHW->NDHWC input conversion.  All on NDHWC activations; parameter gradients use ordered two-stage
// reductions (no floating-point atomics) so they are bit-reproducible.
//
// Replaces nn.ConvTranspose3d(k=s, bias=False) (`tu[u]`) and nn.Conv3d(C, ncls, 1, bias=False) (`seg_outputs[u]`) of
// nnunet's Generic_UNet (Appendix A); forward order restated at reference generic_ViT_UNet.py:261-286.  The transposed
// convolution writes straight into the channel slice [0,Cout) of the concat buffer (pitch 2*Cout), which removes
// torch.cat (generic_ViT_UNet.py:263).
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace b2 {

// ---------------------------------------------------------------------------------------------------------------
// transposed conv, kernel == stride
// ---------------------------------------------------------------------------------------------------------------
// PyTorch [Cin][Cout][K8] -> Wq [K8][Cin][Cout]
__global__ void tconv_shadow_kernel(const float* __restrict__ w, int Cin, int Cout, int K8, float* __restrict__ wq) {
    pdl_grid_sync();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)Cin * Cout * K8;
    if (i >= tot) return;
    int q = (int)(i % K8);
    long long r = i / K8;
    int co = (int)(r % Cout), ci = (int)(r / Cout);
    wq[((long long)q * Cin + ci) * Cout + co] = w[i];
}

struct TG {
    int N, D, H, W, Cin, Cout, kd, kh, kw, K8, in_pitch, out_pitch;
};

constexpr int TI = 16;  // input voxels per CTA

template <typename T>
__global__ void __launch_bounds__(256) tconv_fwd_kernel(TG g, const T* __restrict__ x, const float* __restrict__ wq,
                                                        T* __restrict__ y) {
    pdl_grid_sync();
    extern __shared__ float xs[];  // [TI][Cin]
    const long long Vin = (long long)g.N * g.D * g.H * g.W;
    const long long v0 = (long long)blockIdx.x * TI;
    for (int e = threadIdx.x; e < TI * g.Cin; e += 256) {
        int vi = e / g.Cin, ci = e % g.Cin;
        long long v = v0 + vi;
        xs[e] = v < Vin ? to_f(x[v * g.in_pitch + ci]) : 0.f;
    }
    __syncthreads();
    const int Ho = g.H * g.kh, Wo = g.W * g.kw, Do = g.D * g.kd;
    const int items = TI * g.K8 * g.Cout;
    for (int e = threadIdx.x; e < items; e += 256) {
        const int co = e % g.Cout;
        int r = e / g.Cout;
        const int vi = r % TI, q = r / TI;
        const long long v = v0 + vi;
        if (v >= Vin) continue;
        const float* wp = wq + (long long)q * g.Cin * g.Cout + co;
        const float* xp = xs + vi * g.Cin;
        float acc = 0.f;
        for (int ci = 0; ci < g.Cin; ++ci) acc = fmaf(xp[ci], wp[(long long)ci * g.Cout], acc);
        long long t = v;
        const int iw = (int)(t % g.W); t /= g.W;
        const int ih = (int)(t % g.H); t /= g.H;
        const int id = (int)(t % g.D);
        const int n = (int)(t / g.D);
        const int qw = q % g.kw, qh = (q / g.kw) % g.kh, qd = q / (g.kw * g.kh);
        const long long o = (((long long)n * Do + id * g.kd + qd) * Ho + ih * g.kh + qh) * Wo + iw * g.kw + qw;
        y[o * g.out_pitch + co] = from_f<T>(acc);
    }
}

// dx[v][ci] = sum_q sum_co dy[out(v,q)][co] * W[ci][co][q]
template <typename T>
__global__ void __launch_bounds__(256) tconv_dgrad_kernel(TG g, const T* __restrict__ dy, const float* __restrict__ w_pt,
                                                          T* __restrict__ dx) {
    pdl_grid_sync();
    extern __shared__ float ds[];  // [TI][K8][Cout]
    const long long Vin = (long long)g.N * g.D * g.H * g.W;
    const long long v0 = (long long)blockIdx.x * TI;
    const int Ho = g.H * g.kh, Wo = g.W * g.kw, Do = g.D * g.kd;
    const int per = g.K8 * g.Cout;
    for (int e = threadIdx.x; e < TI * per; e += 256) {
        const int co = e % g.Cout;
        int r = e / g.Cout;
        const int q = r % g.K8, vi = r / g.K8;
        const long long v = v0 + vi;
        float val = 0.f;
        if (v < Vin) {
            long long t = v;
            const int iw = (int)(t % g.W); t /= g.W;
            const int ih = (int)(t % g.H); t /= g.H;
            const int id = (int)(t % g.D);
            const int n = (int)(t / g.D);
            const int qw = q % g.kw, qh = (q / g.kw) % g.kh, qd = q / (g.kw * g.kh);
            const long long o = (((long long)n * Do + id * g.kd + qd) * Ho + ih * g.kh + qh) * Wo + iw * g.kw + qw;
            val = to_f(dy[o * g.out_pitch + co]);
        }
        ds[e] = val;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TI * g.Cin; e += 256) {
        const int ci = e % g.Cin, vi = e / g.Cin;
        const long long v = v0 + vi;
        if (v >= Vin) continue;
        const float* dp = ds + vi * per;                       // [q][co]
        const float* wp = w_pt + (long long)ci * g.Cout * g.K8;  // [co][q]
        float acc = 0.f;
        for (int co = 0; co < g.Cout; ++co)
            for (int q = 0; q < g.K8; ++q) acc = fmaf(dp[q * g.Cout + co], wp[co * g.K8 + q], acc);
        dx[v * g.in_pitch + ci] = from_f<T>(acc);
    }
}

// dW[ci][co][q] partials: grid (split, ci-block, co-block*K8); thread (ci = tid/8, 4 co)
template <typename T>
__global__ void __launch_bounds__(256) tconv_wgrad_kernel(TG g, const T* __restrict__ x, const T* __restrict__ dy,
                                                          int nsplit, float* __restrict__ part) {
    pdl_grid_sync();
    __shared__ float xs[32][33];
    __shared__ __align__(16) float zs[32][32];
    const int cob = blockIdx.z / g.K8, q = blockIdx.z % g.K8;
    const int ci0 = blockIdx.y * 32, co0 = cob * 32;
    const int ci = threadIdx.x >> 3, cog = threadIdx.x & 7;
    const long long Vin = (long long)g.N * g.D * g.H * g.W;
    const int Ho = g.H * g.kh, Wo = g.W * g.kw, Do = g.D * g.kd;
    const int qw = q % g.kw, qh = (q / g.kw) % g.kh, qd = q / (g.kw * g.kh);
    const long long ntiles = (Vin + 31) / 32;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long tile = blockIdx.x; tile < ntiles; tile += nsplit) {
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            const int cc = e & 31, vi = e >> 5;
            const long long v = tile * 32 + vi;
            float xv = 0.f, zv = 0.f;
            if (v < Vin) {
                if (ci0 + cc < g.Cin) xv = to_f(x[v * g.in_pitch + ci0 + cc]);
                if (co0 + cc < g.Cout) {
                    long long t = v;
                    const int iw = (int)(t % g.W); t /= g.W;
                    const int ih = (int)(t % g.H); t /= g.H;
                    const int id = (int)(t % g.D);
                    const int n = (int)(t / g.D);
                    const long long o = (((long long)n * Do + id * g.kd + qd) * Ho + ih * g.kh + qh) * Wo + iw * g.kw + qw;
                    zv = to_f(dy[o * g.out_pitch + co0 + cc]);
                }
            }
            xs[vi][cc] = xv;
            zs[vi][cc] = zv;
        }
        __syncthreads();
#pragma unroll 8
        for (int vi = 0; vi < 32; ++vi) {
            const float xv = xs[vi][ci];
            const float4 z4 = *reinterpret_cast<const float4*>(&zs[vi][cog * 4]);
            acc[0] = fmaf(xv, z4.x, acc[0]);
            acc[1] = fmaf(xv, z4.y, acc[1]);
            acc[2] = fmaf(xv, z4.z, acc[2]);
            acc[3] = fmaf(xv, z4.w, acc[3]);
        }
    }
    if (ci0 + ci < g.Cin) {
        float* out = part + (long long)blockIdx.x * g.Cin * g.Cout * g.K8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + cog * 4 + j;
            if (co < g.Cout) out[((long long)(ci0 + ci) * g.Cout + co) * g.K8 + q] = acc[j];
        }
    }
}

__global__ void ordered_reduce_kernel(const float* __restrict__ part, int nsplit, long long tot, float* __restrict__ out) {
    pdl_grid_sync();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tot) return;
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += part[(long long)k * tot + i];
    out[i] = s;
}

static TG make_tg(const TconvShape& s) {
    TG g;
    g.N = s.n; g.D = s.d; g.H = s.h; g.W = s.w; g.Cin = s.cin; g.Cout = s.cout;
    g.kd = s.k[0]; g.kh = s.k[1]; g.kw = s.k[2]; g.K8 = s.k[0] * s.k[1] * s.k[2];
    g.in_pitch = s.in_pitch; g.out_pitch = s.out_pitch;
    return g;
}

static int tconv_splits(const TconvShape& s) {
    long long vin = (long long)s.n * s.d * s.h * s.w;
    long long tiles = (vin + 31) / 32;
    int blocks = cdiv(s.cin, 32) * cdiv(s.cout, 32) * s.k[0] * s.k[1] * s.k[2];
    long long ns = (4LL * num_sms() + blocks - 1) / blocks;
    if (ns > tiles) ns = tiles;
    if (ns < 1) ns = 1;
    return (int)ns;
}

size_t tconv_bwd_scratch_floats(const TconvShape& s) {
    size_t wsz = (size_t)s.cin * s.cout * s.k[0] * s.k[1] * s.k[2];
    return (size_t)tconv_splits(s) * wsz;
}

// scratch for fwd shadow: the plan keeps wq; this standalone launcher transforms on the fly into `wq`
int tconv_shadow(const float* w_pt, int cin, int cout, int k8, float* wq, cudaStream_t st) {
    long long tot = (long long)cin * cout * k8;
    B2_LAUNCH(tconv_shadow_kernel, cdiv(tot, 256), 256, 0, st, w_pt, cin, cout, k8, wq);
    return B2_OK;
}

template <typename T>
int tconv_fwd_q(const TconvShape& s, const T* x, const float* wq, T* y, cudaStream_t st) {
    TG g = make_tg(s);
    long long vin = (long long)s.n * s.d * s.h * s.w;
    size_t sh = (size_t)TI * s.cin * sizeof(float);
    B2_LAUNCH(tconv_fwd_kernel<T>, cdiv(vin, TI), 256, sh, st, g, x, wq, y);
    return B2_OK;
}

template <typename T>
int tconv_bwd(const TconvShape& s, const T* x, const T* dy, const float* w_pt, T* dx, float* dw, float* scratch,
              cudaStream_t st) {
    TG g = make_tg(s);
    long long vin = (long long)s.n * s.d * s.h * s.w;
    if (dx) {
        size_t sh = (size_t)TI * g.K8 * s.cout * sizeof(float);
        static bool done[2] = {false, false};
        constexpr int ti = sizeof(T) == 4 ? 0 : 1;
        if (!done[ti]) {
            B2_CUDA(cudaFuncSetAttribute(tconv_dgrad_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            done[ti] = true;
        }
        B2_CHECK_ARG(sh <= 200 * 1024);
        B2_LAUNCH(tconv_dgrad_kernel<T>, cdiv(vin, TI), 256, sh, st, g, dy, w_pt, dx);
    }
    if (dw) {
        int ns = tconv_splits(s);
        dim3 grid(ns, cdiv(s.cin, 32), cdiv(s.cout, 32) * g.K8);
        B2_LAUNCH(tconv_wgrad_kernel<T>, grid, 256, 0, st, g, x, dy, ns, scratch);
        long long tot = (long long)s.cin * s.cout * g.K8;
        B2_LAUNCH(ordered_reduce_kernel, cdiv(tot, 256), 256, 0, st, scratch, ns, tot, dw);
    }
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// 1x1x1 segmentation head
// ---------------------------------------------------------------------------------------------------------------
constexpr int MAXCLS = 8;

template <typename T, int VW>
__global__ void __launch_bounds__(256) seghead_fwd_kernel(const T* __restrict__ y, const float* __restrict__ w,
                                                          float* __restrict__ logits, int n, long long vox, int c,
                                                          int ncls, int y_pitch) {
    pdl_grid_sync();
    extern __shared__ float ws[];  // [ncls][c]
    for (int e = threadIdx.x; e < ncls * c; e += blockDim.x) ws[e] = w[e];
    __syncthreads();
    const long long total = (long long)n * vox;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
        float acc[MAXCLS];
#pragma unroll
        for (int k = 0; k < MAXCLS; ++k) acc[k] = 0.f;
        const T* p = y + v * y_pitch;
        for (int cc = 0; cc < c; cc += VW) {
            float a[VW];
            if (VW == 8) load8(p + cc, *reinterpret_cast<float(*)[8]>(a));
            else a[0] = to_f(p[cc]);
#pragma unroll
            for (int k = 0; k < MAXCLS; ++k)
                if (k < ncls) {
#pragma unroll
                    for (int j = 0; j < VW; ++j) acc[k] = fmaf(a[j], ws[k * c + cc + j], acc[k]);
                }
        }
        const int nn = (int)(v / vox);
        const long long vv = v % vox;
#pragma unroll
        for (int k = 0; k < MAXCLS; ++k)
            if (k < ncls) logits[((long long)nn * ncls + k) * vox + vv] = acc[k];
    }
}

template <typename T, int VW>
__global__ void __launch_bounds__(256) seghead_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ dl,
                                                            T* __restrict__ dy, int accumulate, int n, long long vox,
                                                            int c, int ncls, int dy_pitch) {
    pdl_grid_sync();
    extern __shared__ float ws[];
    for (int e = threadIdx.x; e < ncls * c; e += blockDim.x) ws[e] = w[e];
    __syncthreads();
    const int ncg = c / VW;
    const long long total = (long long)n * vox * ncg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % ncg);
        const long long v = i / ncg;
        const int nn = (int)(v / vox);
        const long long vv = v % vox;
        float acc[VW];
#pragma unroll
        for (int j = 0; j < VW; ++j) acc[j] = 0.f;
        for (int k = 0; k < ncls; ++k) {
            const float d = dl[((long long)nn * ncls + k) * vox + vv];
#pragma unroll
            for (int j = 0; j < VW; ++j) acc[j] = fmaf(d, ws[k * c + cg * VW + j], acc[j]);
        }
        T* p = dy + v * dy_pitch + cg * VW;
        if (VW == 8) {
            if (accumulate) {
                float o[8];
                load8(p, o);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += o[j];
            }
            store8(p, *reinterpret_cast<float(*)[8]>(acc));
        } else {
            if (accumulate) acc[0] += to_f(*p);
            *p = from_f<T>(acc[0]);
        }
    }
}

// dy = w^T dlogits for c % 16 == 0: one thread per (voxel, 16-channel block).  lane = voxel => the NCDHW dlogits are read
// with coalesced 128-byte loads (one per class), the 16 x ncls weights sit in registers, no 64-bit division anywhere
// (grid.y = sample, grid.z = channel block).  Round 1's kernel was instruction-bound: four threads per voxel, each with
// two 64-bit div / mod per element (71 us for the full-resolution head against a 25 us HBM floor).
template <typename T, int NCLS>
__global__ void __launch_bounds__(256, 3) seghead_dgrad32_kernel(const float* __restrict__ w, const float* __restrict__ dl,
                                                              T* __restrict__ dy, int accumulate, int vox, int c, int dy_pitch) {
    pdl_grid_sync();
    const int nn = blockIdx.y, cb = blockIdx.z;
    float wr[NCLS][16];
#pragma unroll
    for (int k = 0; k < NCLS; ++k)
#pragma unroll
        for (int j = 0; j < 16; ++j) wr[k][j] = __ldg(w + k * c + cb * 16 + j);
    const float* dln = dl + (long long)nn * NCLS * vox;
    T* dyn = dy + (long long)nn * vox * dy_pitch + cb * 16;
    for (int v = blockIdx.x * 256 + threadIdx.x; v < vox; v += gridDim.x * 256) {
        float d[NCLS];
#pragma unroll
        for (int k = 0; k < NCLS; ++k) d[k] = dln[(long long)k * vox + v];
        T* p = dyn + (long long)v * dy_pitch;
#pragma unroll
        for (int j0 = 0; j0 < 16; j0 += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < NCLS; ++k) a = fmaf(d[k], wr[k][j0 + j], a);
                acc[j] = a;
            }
            if (accumulate) {
                float o[8];
                load8(p + j0, o);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += o[j];
            }
            store8(p + j0, acc);
        }
    }
}

// dw[k][c] partial per (slab, sample): thread = (channel group of VW, row lane); VW-wide vector loads of y.
// grid = (slabs, n): a slab never straddles two samples, so the loop carries no 64-bit division.
template <typename T, int VW>
__global__ void __launch_bounds__(256) seghead_wgrad_kernel(const T* __restrict__ y, const float* __restrict__ dl,
                                                            int slabs, int n, long long vox, int c, int ncls,
                                                            int y_pitch, float* __restrict__ part) {
    pdl_grid_sync();
    extern __shared__ float sh[];  // [R][ncls][c]
    const int ncg = c / VW;
    const int R = 256 / ncg;
    const int cg = threadIdx.x % ncg, r = threadIdx.x / ncg;
    const int nn = blockIdx.y;
    const long long per = (vox + slabs - 1) / slabs;
    const long long v0 = (long long)blockIdx.x * per, v1 = v0 + per < vox ? v0 + per : vox;
    float acc[MAXCLS][VW];
#pragma unroll
    for (int k = 0; k < MAXCLS; ++k)
#pragma unroll
        for (int j = 0; j < VW; ++j) acc[k][j] = 0.f;
    if (r < R) {
        const T* yp = y + (long long)nn * vox * y_pitch + cg * VW;
        const float* dp = dl + (long long)nn * ncls * vox;
#pragma unroll 2
        for (long long v = v0 + r; v < v1; v += R) {
            float a[VW];
            if (VW == 8) load8(yp + v * y_pitch, *reinterpret_cast<float(*)[8]>(a));
            else a[0] = to_f(yp[v * y_pitch]);
#pragma unroll
            for (int k = 0; k < MAXCLS; ++k)
                if (k < ncls) {
                    const float d = dp[(long long)k * vox + v];
#pragma unroll
                    for (int j = 0; j < VW; ++j) acc[k][j] = fmaf(a[j], d, acc[k][j]);
                }
        }
#pragma unroll
        for (int k = 0; k < MAXCLS; ++k)   // fully unrolled: a runtime-indexed acc[k] would live in local memory
            if (k < ncls) {
#pragma unroll
                for (int j = 0; j < VW; ++j) sh[((size_t)r * ncls + k) * c + cg * VW + j] = acc[k][j];
            }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < ncls * c; e += 256) {
        float s = 0.f;
        for (int q = 0; q < R; ++q) s += sh[(size_t)q * ncls * c + e];
        part[((long long)nn * slabs + blockIdx.x) * ncls * c + e] = s;
    }
}

// Same partials for C % 32 == 0 and NC <= 4 classes: lane = voxel, a thread accumulates all NC x 32 products of its voxels in
// registers (one 64-byte row of y + NC logit gradients per voxel: fully coalesced, nothing shared between lanes), and the
// lanes are combined once at the end by a transpose-reduce (31 shuffles per 32 values, fixed tree => bit-reproducible).
// grid = (slabs, n, C / 32).  The channel-group kernel above reached 0.66 TB/s on the full-resolution head (241 us).
template <typename T, int NC>
__global__ void __launch_bounds__(256) seghead_wgrad32_kernel(const T* __restrict__ y, const float* __restrict__ dl, int slabs,
                                                              long long vox, int c, int ncls, int y_pitch, float* __restrict__ part) {
    pdl_grid_sync();
    __shared__ float sh[8][NC * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nn = blockIdx.y, c0 = blockIdx.z * 32;
    const long long per = (vox + slabs - 1) / slabs;
    const long long v0 = (long long)blockIdx.x * per, v1 = v0 + per < vox ? v0 + per : vox;
    float acc[NC][32];
#pragma unroll
    for (int k = 0; k < NC; ++k)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[k][j] = 0.f;
    const T* yp = y + (long long)nn * vox * y_pitch + c0;
    const float* dp = dl + (long long)nn * ncls * vox;
    for (long long v = v0 + threadIdx.x; v < v1; v += 256) {
        float a[32], d[NC];
#pragma unroll
        for (int j = 0; j < 32; j += 8) load8(yp + v * y_pitch + j, *reinterpret_cast<float(*)[8]>(&a[j]));
#pragma unroll
        for (int k = 0; k < NC; ++k) d[k] = k < ncls ? dp[(long long)k * vox + v] : 0.f;
#pragma unroll
        for (int k = 0; k < NC; ++k)
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[k][j] = fmaf(a[j], d[k], acc[k][j]);
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        // transpose-reduce: afterwards lane l holds the warp total of value l in acc[k][0]
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
            for (int i = 0; i < off; ++i) {
                const bool hi = (lane & off) != 0;
                const float send = hi ? acc[k][i] : acc[k][i + off];
                const float keep = hi ? acc[k][i + off] : acc[k][i];
                acc[k][i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
        sh[warp][k * 32 + lane] = acc[k][0];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NC * 32; e += 256) {
        const int k = e >> 5, j = e & 31;
        if (k < ncls) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += sh[w][e];
            part[(((long long)nn * slabs + blockIdx.x) * ncls + k) * c + c0 + j] = t;
        }
    }
}

// out[i] = sum_k part[k][i] for MANY partials of FEW outputs: one warp per output, lanes stride the partials, fixed
// shuffle tree (bit-reproducible).  (The thread-per-output kernel walks the partials serially: 76 us for 1184 x 96.)
__global__ void __launch_bounds__(256) ordered_reduce_wide_kernel(const float* __restrict__ part, int nsplit, long long tot,
                                                                  float* __restrict__ out) {
    pdl_grid_sync();
    const long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= tot) return;
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int k = lane; k < nsplit; k += 32) s += part[(long long)k * tot + i];
    s = warp_sum(s);
    if (lane == 0) out[i] = s;
}

template <typename T>
int seghead_fwd(const T* y, const float* w, float* logits, int n, long long vox, int c, int ncls, int y_pitch,
                cudaStream_t st) {
    B2_CHECK_ARG(ncls <= MAXCLS);
    long long total = (long long)n * vox;
    int grid = (int)((total + 255) / 256);
    int cap = num_sms() * 16;
    if (grid > cap) grid = cap;
    if (c % 8 == 0 && y_pitch % 8 == 0) B2_LAUNCH((seghead_fwd_kernel<T, 8>), grid, 256, (size_t)ncls * c * sizeof(float), st, y, w, logits, n, vox, c, ncls, y_pitch);
    else B2_LAUNCH((seghead_fwd_kernel<T, 1>), grid, 256, (size_t)ncls * c * sizeof(float), st, y, w, logits, n, vox, c, ncls, y_pitch);
    return B2_OK;
}

// slabs per sample
static int seghead_slabs(int n, long long vox) {
    long long s = (8LL * num_sms() + n - 1) / n;
    long long maxs = (vox + 63) / 64;
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    return (int)s;
}

size_t seghead_bwd_scratch_floats(int n, long long vox, int c, int ncls) {
    return (size_t)n * seghead_slabs(n, vox) * ncls * c;
}

template <typename T>
int seghead_bwd(const T* y, const float* w, const float* dlogits, T* dy, int accumulate, float* dw, int n,
                long long vox, int c, int ncls, int y_pitch, int dy_pitch, float* scratch, cudaStream_t st) {
    B2_CHECK_ARG(ncls <= MAXCLS && c <= 256);
    if (dy) {
        const bool v8 = (c % 8 == 0) && (dy_pitch % 8 == 0);
        long long total = (long long)n * vox * (v8 ? c / 8 : c);
        long long grid = (total + 255) / 256;
        long long cap = (long long)num_sms() * 16;
        if (grid > cap) grid = cap;
        if (v8 && c % 16 == 0 && ncls >= 2 && ncls <= 4 && vox < (1LL << 31) && !std::is_same<T, float>::value) {
            long long gx = (vox + 255) / 256, capx = (long long)num_sms() * 12 / (n * (c / 16)) + 1;
            if (gx > capx) gx = capx;
            dim3 g3((unsigned)gx, n, c / 16);
            if (ncls == 2) B2_LAUNCH((seghead_dgrad32_kernel<T, 2>), g3, 256, 0, st, w, dlogits, dy, accumulate, (int)vox, c, dy_pitch);
            else if (ncls == 3) B2_LAUNCH((seghead_dgrad32_kernel<T, 3>), g3, 256, 0, st, w, dlogits, dy, accumulate, (int)vox, c, dy_pitch);
            else B2_LAUNCH((seghead_dgrad32_kernel<T, 4>), g3, 256, 0, st, w, dlogits, dy, accumulate, (int)vox, c, dy_pitch);
        } else if (v8) B2_LAUNCH((seghead_dgrad_kernel<T, 8>), (int)grid, 256, (size_t)ncls * c * sizeof(float), st, w, dlogits, dy, accumulate, n, vox, c, ncls, dy_pitch);
        else B2_LAUNCH((seghead_dgrad_kernel<T, 1>), (int)grid, 256, (size_t)ncls * c * sizeof(float), st, w, dlogits, dy, accumulate, n, vox, c, ncls, dy_pitch);
    }
    if (dw) {
        int slabs = seghead_slabs(n, vox);
        const bool v8 = (c % 8 == 0) && (y_pitch % 8 == 0);
        const int ncg = v8 ? c / 8 : c;
        const int R = 256 / ncg;
        size_t sh = (size_t)R * ncls * c * sizeof(float);
        B2_CHECK_ARG(sh <= 48 * 1024);
        dim3 grid(slabs, n);
        if (v8 && c % 32 == 0 && ncls <= 4) {
            dim3 g3(slabs, n, c / 32);
            if (ncls <= 2) B2_LAUNCH((seghead_wgrad32_kernel<T, 2>), g3, 256, 0, st, y, dlogits, slabs, vox, c, ncls, y_pitch, scratch);
            else if (ncls == 3) B2_LAUNCH((seghead_wgrad32_kernel<T, 3>), g3, 256, 0, st, y, dlogits, slabs, vox, c, ncls, y_pitch, scratch);
            else B2_LAUNCH((seghead_wgrad32_kernel<T, 4>), g3, 256, 0, st, y, dlogits, slabs, vox, c, ncls, y_pitch, scratch);
        } else if (v8) B2_LAUNCH((seghead_wgrad_kernel<T, 8>), grid, 256, sh, st, y, dlogits, slabs, n, vox, c, ncls, y_pitch, scratch);
        else B2_LAUNCH((seghead_wgrad_kernel<T, 1>), grid, 256, sh, st, y, dlogits, slabs, n, vox, c, ncls, y_pitch, scratch);
        long long tot = (long long)ncls * c;
        if (slabs * n >= 64) B2_LAUNCH(ordered_reduce_wide_kernel, cdiv(tot, 8), 256, 0, st, scratch, slabs * n, tot, dw);
        else B2_LAUNCH(ordered_reduce_kernel, cdiv(tot, 256), 256, 0, st, scratch, slabs * n, tot, dw);
    }
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// layout conversion / fill
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_ndhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int n, int c, long long vox,
                                     int dst_pitch) {
    pdl_grid_sync();
    const long long total = (long long)n * c * vox;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long v = i % vox;
        const long long r = i / vox;
        const int cc = (int)(r % c), nn = (int)(r / c);
        dst[((long long)nn * vox + v) * dst_pitch + cc] = from_f<T>(src[i]);
    }
}

template <typename T>
int nchw_to_ndhwc(const float* src, T* dst, int n, int c, long long vox, int dst_pitch, cudaStream_t st) {
    long long total = (long long)n * c * vox;
    long long grid = (total + 255) / 256;
    long long cap = (long long)num_sms() * 16;
    if (grid > cap) grid = cap;
    B2_LAUNCH(nchw_to_ndhwc_kernel<T>, (int)grid, 256, 0, st, src, dst, n, c, vox, dst_pitch);
    return B2_OK;
}

template int tconv_fwd_q<float>(const TconvShape&, const float*, const float*, float*, cudaStream_t);
template int tconv_fwd_q<__nv_bfloat16>(const TconvShape&, const __nv_bfloat16*, const float*, __nv_bfloat16*, cudaStream_t);
template int tconv_bwd<float>(const TconvShape&, const float*, const float*, const float*, float*, float*, float*, cudaStream_t);
template int tconv_bwd<__nv_bfloat16>(const TconvShape&, const __nv_bfloat16*, const __nv_bfloat16*, const float*, __nv_bfloat16*, float*, float*, cudaStream_t);
template int seghead_fwd<float>(const float*, const float*, float*, int, long long, int, int, int, cudaStream_t);
template int seghead_fwd<__nv_bfloat16>(const __nv_bfloat16*, const float*, float*, int, long long, int, int, int, cudaStream_t);
template int seghead_bwd<float>(const float*, const float*, const float*, float*, int, float*, int, long long, int, int, int, int, float*, cudaStream_t);
template int seghead_bwd<__nv_bfloat16>(const __nv_bfloat16*, const float*, const float*, __nv_bfloat16*, int, float*, int, long long, int, int, int, int, float*, cudaStream_t);
template int nchw_to_ndhwc<float>(const float*, float*, int, int, long long, int, cudaStream_t);
template int nchw_to_ndhwc<__nv_bfloat16>(const float*, __nv_bfloat16*, int, int, long long, int, cudaStream_t);

}  // namespace b2
