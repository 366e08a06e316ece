// optim.cu -- multi-tensor kernels over the parameter list: EWC / RW quadratic penalty (+ analytic gradient), Fisher
// and Riemannian-walk importance updates, gradient-norm clipping + SGD-Nesterov step.
//
// The reference does each of these as a Python loop over ~98 tensors with ~6 tiny launches per tensor per stored task
// (reference loss_functions/deep_supervision.py:65-80, :115-132; ewc:298-304; rw:240-262; MultiHead:629-641).  Here each
// is ONE launch over all tensors (block -> (tensor, chunk) by binary search over a chunk-prefix table), HBM-bound,
// 128-bit accesses, ordered two-stage reductions (bit-reproducible values and gradients).
#include "common.cuh"
#include "kernels.h"

namespace b2 {

constexpr int CHUNK = 8192;  // elements per block
constexpr int MT_THREADS = 256;

struct MTLoc { int tensor; long long off; long long n; };

__device__ __forceinline__ MTLoc mt_locate(const int* __restrict__ prefix, int n_tensors, const long long* __restrict__ numel) {
    int lo = 0, hi = n_tensors;  // find largest t with prefix[t] <= blockIdx.x
    const int b = blockIdx.x;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid] <= b) lo = mid; else hi = mid;
    }
    MTLoc l;
    l.tensor = lo;
    l.off = (long long)(b - prefix[lo]) * CHUNK;
    long long rem = numel[lo] - l.off;
    l.n = rem < CHUNK ? rem : CHUNK;
    return l;
}

// device table header written by the host into scratch:
//   [int prefix[n+1]] [pad] [long long numel[n]] [entries...]
struct MTHost {
    std::string blob;
    int nblocks;
    size_t off_numel, off_entries;
};

template <typename E>
static MTHost mt_build(const E* table, int n) {
    MTHost h;
    size_t off_prefix = 0;
    h.off_numel = align_up((n + 1) * sizeof(int), 16);
    h.off_entries = h.off_numel + align_up(n * sizeof(long long), 16);
    h.blob.assign(h.off_entries + n * sizeof(E), '\0');
    int* prefix = (int*)(&h.blob[off_prefix]);
    long long* numel = (long long*)(&h.blob[h.off_numel]);
    int acc = 0;
    for (int i = 0; i < n; ++i) {
        prefix[i] = acc;
        numel[i] = table[i].numel;
        acc += (int)((table[i].numel + CHUNK - 1) / CHUNK);
    }
    prefix[n] = acc;
    h.nblocks = acc;
    memcpy(&h.blob[h.off_entries], table, n * sizeof(E));
    return h;
}

static inline size_t mt_table_bytes(int n, size_t esz) {
    return align_up(align_up((n + 1) * sizeof(int), 16) + align_up(n * sizeof(long long), 16) + n * esz);
}

__device__ __forceinline__ bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// ordered final sum of `n` partials by one block
__global__ void __launch_bounds__(256) ordered_sum_kernel(const float* __restrict__ part, int n, float scale,
                                                          float* __restrict__ out, int accumulate) {
    pdl_grid_sync();
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += (double)part[i];
    double r = block_sum(s, red);
    if (threadIdx.x == 0) {
        float v = (float)(r * (double)scale);
        out[0] = accumulate ? out[0] + v : v;
    }
}

// ---- quadratic penalty ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MT_THREADS) quadpen_kernel(const int* __restrict__ prefix, const long long* __restrict__ numel,
                                                             const b2_pen_entry* __restrict__ ent, int n_tensors,
                                                             float coef, float* __restrict__ part) {
    pdl_grid_sync();
    __shared__ float red[32];
    const MTLoc l = mt_locate(prefix, n_tensors, numel);
    const b2_pen_entry e = ent[l.tensor];
    const float* th = e.theta + l.off;
    const float* ts = e.theta_star + l.off;
    const float* fi = e.fisher + l.off;
    const float* im = e.importance ? e.importance + l.off : nullptr;
    float* gr = e.grad ? e.grad + l.off : nullptr;
    float acc = 0.f;
    const bool vec = aligned16(th) && aligned16(ts) && aligned16(fi) && (!im || aligned16(im)) && (!gr || aligned16(gr));
    const long long nv = vec ? (l.n / 4) * 4 : 0;
    for (long long i = (long long)threadIdx.x * 4; i < nv; i += MT_THREADS * 4) {
        const float4 a = *reinterpret_cast<const float4*>(th + i);
        const float4 b = *reinterpret_cast<const float4*>(ts + i);
        float4 f = *reinterpret_cast<const float4*>(fi + i);
        if (im) {
            const float4 s = *reinterpret_cast<const float4*>(im + i);
            f.x += s.x; f.y += s.y; f.z += s.z; f.w += s.w;
        }
        const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z, dw = a.w - b.w;
        acc += f.x * (dx * dx);
        acc += f.y * (dy * dy);
        acc += f.z * (dz * dz);
        acc += f.w * (dw * dw);
        if (gr) {
            float4 g = *reinterpret_cast<float4*>(gr + i);
            g.x += 2.f * coef * f.x * dx; g.y += 2.f * coef * f.y * dy;
            g.z += 2.f * coef * f.z * dz; g.w += 2.f * coef * f.w * dw;
            *reinterpret_cast<float4*>(gr + i) = g;
        }
    }
    for (long long i = nv + threadIdx.x; i < l.n; i += MT_THREADS) {
        float f = fi[i];
        if (im) f += im[i];
        const float d = th[i] - ts[i];
        acc += f * (d * d);
        if (gr) gr[i] += 2.f * coef * f * d;
    }
    const float r = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// ---- Fisher / RW ------------------------------------------------------------------------------------------------
struct SqEntry { const float* grad; float* fisher; long long numel; };

__global__ void __launch_bounds__(MT_THREADS) fisher_square_kernel(const int* __restrict__ prefix, const long long* __restrict__ numel,
                                                                   const SqEntry* __restrict__ ent, int n_tensors) {
    pdl_grid_sync();
    const MTLoc l = mt_locate(prefix, n_tensors, numel);
    const SqEntry e = ent[l.tensor];
    for (long long i = threadIdx.x; i < l.n; i += MT_THREADS) {
        const float g = e.grad[l.off + i];
        e.fisher[l.off + i] = g * g;
    }
}

__global__ void __launch_bounds__(MT_THREADS) rw_update_kernel(const int* __restrict__ prefix, const long long* __restrict__ numel,
                                                               const b2_rw_entry* __restrict__ ent, int n_tensors,
                                                               float alpha, float eps, int have_prev) {
    pdl_grid_sync();
    const MTLoc l = mt_locate(prefix, n_tensors, numel);
    const b2_rw_entry e = ent[l.tensor];
    for (long long i = threadIdx.x; i < l.n; i += MT_THREADS) {
        const long long k = l.off + i;
        const float th = e.theta[k], g = e.grad[k], F = e.fisher[k];
        if (have_prev) {
            // rw:243-251 -- evaluated in the reference's operation order
            const float pv = e.prev[k];
            const float delta = g * (pv - th);
            const float dd = th - pv;
            const float den = 0.5f * F * (dd * dd) + eps;
            float s = delta / den;
            if (s < 0.f) s = 0.f;
            e.score[k] += s;
        }
        e.prev[k] = th;                                       // rw:254
        e.fisher[k] = (alpha * (g * g)) + ((1.f - alpha) * F);  // rw:260-262
    }
}

// ---- clip + SGD -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MT_THREADS) gradnorm_kernel(const int* __restrict__ prefix, const long long* __restrict__ numel,
                                                              const b2_sgd_entry* __restrict__ ent, int n_tensors,
                                                              float* __restrict__ part) {
    pdl_grid_sync();
    __shared__ float red[32];
    const MTLoc l = mt_locate(prefix, n_tensors, numel);
    const float* g = ent[l.tensor].grad + l.off;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < l.n; i += MT_THREADS) { const float v = g[i]; acc += v * v; }
    const float r = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// norm = sqrt(sum); clip = min(1, max_norm/(norm+1e-6))  -> out[0] = norm, out[1] = clip
__global__ void __launch_bounds__(256) clipcoef_kernel(const float* __restrict__ part, int n, float max_norm,
                                                       float* __restrict__ out, const float* __restrict__ hyper) {
    pdl_grid_sync();
    if (hyper) max_norm = hyper[3];
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += (double)part[i];
    double r = block_sum(s, red);
    if (threadIdx.x == 0) {
        const float norm = (float)sqrt(r);
        float c = max_norm / (norm + 1e-6f);
        if (c > 1.f) c = 1.f;
        out[0] = norm;
        out[1] = c;
    }
}

__global__ void __launch_bounds__(MT_THREADS) sgd_kernel(const int* __restrict__ prefix, const long long* __restrict__ numel,
                                                         const b2_sgd_entry* __restrict__ ent, int n_tensors, float lr,
                                                         float momentum, float wd, int nesterov, int first_step,
                                                         const float* __restrict__ clip, const float* __restrict__ hyper) {
    pdl_grid_sync();
    if (hyper) { lr = hyper[0]; momentum = hyper[1]; wd = hyper[2]; }   // device-resident schedule: a captured graph follows lr changes
    const MTLoc l = mt_locate(prefix, n_tensors, numel);
    const b2_sgd_entry e = ent[l.tensor];
    const float c = clip[1];
    float* gw = const_cast<float*>(e.grad) + l.off;
    float* th = e.theta + l.off;
    float* mo = e.momentum + l.off;
    auto upd = [&](float g0, float p, float m, float& g_out, float& p_out, float& m_out) {
        const float g = g0 * c;   // clip_grad_norm_ scales .grad in place (RW reads it afterwards, rw:223-225)
        g_out = g;
        float d = g + wd * p;
        const float buf = first_step ? d : momentum * m + d;
        m_out = buf;
        d = nesterov ? d + momentum * buf : buf;
        p_out = p - lr * d;
    };
    // 16-byte accesses (3 reads + 3 writes per element: the step is pure HBM traffic)
    const bool vec = aligned16(gw) && aligned16(th) && aligned16(mo);
    const long long nv = vec ? (l.n / 4) * 4 : 0;
    for (long long i = (long long)threadIdx.x * 4; i < nv; i += MT_THREADS * 4) {
        float4 g = *reinterpret_cast<const float4*>(gw + i);
        float4 p = *reinterpret_cast<const float4*>(th + i);
        float4 m = first_step ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(mo + i);
        upd(g.x, p.x, m.x, g.x, p.x, m.x);
        upd(g.y, p.y, m.y, g.y, p.y, m.y);
        upd(g.z, p.z, m.z, g.z, p.z, m.z);
        upd(g.w, p.w, m.w, g.w, p.w, m.w);
        *reinterpret_cast<float4*>(gw + i) = g;
        *reinterpret_cast<float4*>(mo + i) = m;
        *reinterpret_cast<float4*>(th + i) = p;
    }
    for (long long i = nv + threadIdx.x; i < l.n; i += MT_THREADS) {
        float g, p, m;
        upd(gw[i], th[i], first_step ? 0.f : mo[i], g, p, m);
        gw[i] = g; mo[i] = m; th[i] = p;
    }
}

}  // namespace b2

using namespace b2;

extern "C" size_t b2_quadpen_scratch_bytes(int n_tensors, int64_t total_numel) {
    return mt_table_bytes(n_tensors, sizeof(b2_pen_entry)) + align_up(((size_t)total_numel / CHUNK + n_tensors + 8) * sizeof(float));
}

extern "C" int b2_quadpen_fwd_bwd(const b2_pen_entry* table_host, int n_tensors, float coef, float* loss_out,
                                  void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(table_host && n_tensors > 0 && loss_out && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    MTHost h = mt_build(table_host, n_tensors);
    char* dev = (char*)scratch;
    B2_CUDA(cudaMemcpyAsync(dev, h.blob.data(), h.blob.size(), cudaMemcpyHostToDevice, st));
    float* part = (float*)(dev + mt_table_bytes(n_tensors, sizeof(b2_pen_entry)));
    B2_LAUNCH(quadpen_kernel, h.nblocks, MT_THREADS, 0, st, (const int*)dev, (const long long*)(dev + h.off_numel),
              (const b2_pen_entry*)(dev + h.off_entries), n_tensors, coef, part);
    B2_LAUNCH(ordered_sum_kernel, 1, 256, 0, st, part, h.nblocks, coef, loss_out, 1);
    return B2_OK;
}

extern "C" size_t b2_multitensor_scratch_bytes(int n_tensors) { return mt_table_bytes(n_tensors, 64); }

extern "C" int b2_fisher_square(const float* const* grads_host, float* const* fisher_host, const int64_t* numel_host,
                                int n_tensors, void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(grads_host && fisher_host && numel_host && n_tensors > 0 && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    std::string tmp(n_tensors * sizeof(SqEntry), '\0');
    SqEntry* t = (SqEntry*)&tmp[0];
    for (int i = 0; i < n_tensors; ++i) { t[i].grad = grads_host[i]; t[i].fisher = fisher_host[i]; t[i].numel = numel_host[i]; }
    MTHost h = mt_build(t, n_tensors);
    char* dev = (char*)scratch;
    B2_CUDA(cudaMemcpyAsync(dev, h.blob.data(), h.blob.size(), cudaMemcpyHostToDevice, st));
    B2_LAUNCH(fisher_square_kernel, h.nblocks, MT_THREADS, 0, st, (const int*)dev, (const long long*)(dev + h.off_numel),
              (const SqEntry*)(dev + h.off_entries), n_tensors);
    return B2_OK;
}

extern "C" int b2_rw_update(const b2_rw_entry* table_host, int n_tensors, float alpha, float eps, int have_prev,
                            void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(table_host && n_tensors > 0 && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    MTHost h = mt_build(table_host, n_tensors);
    char* dev = (char*)scratch;
    B2_CUDA(cudaMemcpyAsync(dev, h.blob.data(), h.blob.size(), cudaMemcpyHostToDevice, st));
    B2_LAUNCH(rw_update_kernel, h.nblocks, MT_THREADS, 0, st, (const int*)dev, (const long long*)(dev + h.off_numel),
              (const b2_rw_entry*)(dev + h.off_entries), n_tensors, alpha, eps, have_prev);
    return B2_OK;
}

extern "C" size_t b2_sgd_scratch_bytes(int n_tensors, int64_t total_numel) {
    return mt_table_bytes(n_tensors, sizeof(b2_sgd_entry)) + align_up(((size_t)total_numel / CHUNK + n_tensors + 8) * sizeof(float));
}

extern "C" int b2_sgd_clip_step(const b2_sgd_entry* table_host, int n_tensors, float lr, float momentum,
                                float weight_decay, int nesterov, float max_norm, int first_step, float* norm_out,
                                void* scratch, b2_stream_t stream) {
    B2_CHECK_ARG(table_host && n_tensors > 0 && scratch);
    cudaStream_t st = (cudaStream_t)stream;
    MTHost h = mt_build(table_host, n_tensors);
    char* dev = (char*)scratch;
    B2_CUDA(cudaMemcpyAsync(dev, h.blob.data(), h.blob.size(), cudaMemcpyHostToDevice, st));
    float* part = (float*)(dev + mt_table_bytes(n_tensors, sizeof(b2_sgd_entry)));
    float* clip = part + h.nblocks;
    const int* prefix = (const int*)dev;
    const long long* numel = (const long long*)(dev + h.off_numel);
    const b2_sgd_entry* ent = (const b2_sgd_entry*)(dev + h.off_entries);
    B2_LAUNCH(gradnorm_kernel, h.nblocks, MT_THREADS, 0, st, prefix, numel, ent, n_tensors, part);
    B2_LAUNCH(clipcoef_kernel, 1, 256, 0, st, part, h.nblocks, max_norm, clip, (const float*)nullptr);
    B2_LAUNCH(sgd_kernel, h.nblocks, MT_THREADS, 0, st, prefix, numel, ent, n_tensors, lr, momentum, weight_decay, nesterov, first_step, clip, (const float*)nullptr);
    if (norm_out) B2_CUDA(cudaMemcpyAsync(norm_out, clip, sizeof(float), cudaMemcpyDeviceToDevice, st));
    return B2_OK;
}

// ---- persistent device tables ----------------------------------------------------------------------------------------
// The *_host entry points above rebuild and upload their (tensor, chunk) tables on every call (a pageable host copy: it
// serialises with the host and cannot be captured in a CUDA graph).  A trainer builds each table ONCE on the host
// (b2_mt_blob_build), keeps it in device memory and runs the *_dev variants: no host work per step, graph-capturable.
static size_t mt_entry_size(int kind) {
    return kind == B2_MT_PEN ? sizeof(b2_pen_entry) : kind == B2_MT_SGD ? sizeof(b2_sgd_entry) : kind == B2_MT_RW ? sizeof(b2_rw_entry) : 0;
}
extern "C" size_t b2_mt_blob_bytes(int kind, int n_tensors) {
    const size_t e = mt_entry_size(kind);
    return e && n_tensors > 0 ? mt_table_bytes(n_tensors, e) : 0;
}
extern "C" int b2_mt_blob_build(int kind, const void* table_host, int n_tensors, void* blob_host, int32_t* nblocks_out) {
    B2_CHECK_ARG(table_host && blob_host && nblocks_out && n_tensors > 0 && mt_entry_size(kind));
    MTHost h;
    if (kind == B2_MT_PEN) h = mt_build((const b2_pen_entry*)table_host, n_tensors);
    else if (kind == B2_MT_SGD) h = mt_build((const b2_sgd_entry*)table_host, n_tensors);
    else h = mt_build((const b2_rw_entry*)table_host, n_tensors);
    memset(blob_host, 0, b2_mt_blob_bytes(kind, n_tensors));
    memcpy(blob_host, h.blob.data(), h.blob.size());
    *nblocks_out = h.nblocks;
    return B2_OK;
}
static inline size_t blob_off_numel(int n) { return align_up((n + 1) * sizeof(int), 16); }
static inline size_t blob_off_entries(int n) { return blob_off_numel(n) + align_up(n * sizeof(long long), 16); }

extern "C" size_t b2_mt_part_bytes(int nblocks) { return align_up(((size_t)nblocks + 8) * sizeof(float)); }

extern "C" int b2_quadpen_dev(const void* blob_dev, int n_tensors, int nblocks, float coef, float* loss_out, float* part,
                              b2_stream_t stream) {
    B2_CHECK_ARG(blob_dev && n_tensors > 0 && nblocks > 0 && loss_out && part);
    cudaStream_t st = (cudaStream_t)stream;
    const char* dev = (const char*)blob_dev;
    B2_LAUNCH(quadpen_kernel, nblocks, MT_THREADS, 0, st, (const int*)dev, (const long long*)(dev + blob_off_numel(n_tensors)),
              (const b2_pen_entry*)(dev + blob_off_entries(n_tensors)), n_tensors, coef, part);
    B2_LAUNCH(ordered_sum_kernel, 1, 256, 0, st, (const float*)part, nblocks, coef, loss_out, 1);
    return B2_OK;
}

extern "C" int b2_sgd_clip_step_dev(const void* blob_dev, int n_tensors, int nblocks, const float* hyper_dev, int nesterov,
                                    float* norm_out, float* part, b2_stream_t stream) {
    B2_CHECK_ARG(blob_dev && n_tensors > 0 && nblocks > 0 && hyper_dev && part);
    cudaStream_t st = (cudaStream_t)stream;
    const char* dev = (const char*)blob_dev;
    const int* prefix = (const int*)dev;
    const long long* numel = (const long long*)(dev + blob_off_numel(n_tensors));
    const b2_sgd_entry* ent = (const b2_sgd_entry*)(dev + blob_off_entries(n_tensors));
    float* clip = part + nblocks;
    B2_LAUNCH(gradnorm_kernel, nblocks, MT_THREADS, 0, st, prefix, numel, ent, n_tensors, part);
    B2_LAUNCH(clipcoef_kernel, 1, 256, 0, st, (const float*)part, nblocks, 0.f, clip, hyper_dev);
    B2_LAUNCH(sgd_kernel, nblocks, MT_THREADS, 0, st, prefix, numel, ent, n_tensors, 0.f, 0.f, 0.f, nesterov, 0, (const float*)clip, hyper_dev);
    if (norm_out) B2_CUDA(cudaMemcpyAsync(norm_out, clip, sizeof(float), cudaMemcpyDeviceToDevice, st));
    return B2_OK;
}

extern "C" int b2_rw_update_dev(const void* blob_dev, int n_tensors, int nblocks, float alpha, float eps, int have_prev,
                                b2_stream_t stream) {
    B2_CHECK_ARG(blob_dev && n_tensors > 0 && nblocks > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const char* dev = (const char*)blob_dev;
    B2_LAUNCH(rw_update_kernel, nblocks, MT_THREADS, 0, st, (const int*)dev, (const long long*)(dev + blob_off_numel(n_tensors)),
              (const b2_rw_entry*)(dev + blob_off_entries(n_tensors)), n_tensors, alpha, eps, have_prev);
    return B2_OK;
}
