// conv3d_tc_halo.cu -- stride-1 3x3x3 bf16 convolution for the thin layers (K = 32 or 64 input channels), where the
// per-tap kernel of conv3d_tc.cu is bound by L2->SM traffic (every 128-voxel A tile is fetched 27 times: FLOP/byte of
// L2 traffic == Cout).  Here a CTA walks a strip of 16 (h) x 8 (w) output voxels along d and keeps a ring of three input
// d-slabs in shared memory; every slab is fetched ONCE as three w-shifted copies (w0-1, w0, w0+1) of an 18 (h) x 8 (w) box,
// so that the A operand of every tap (kd, kh, kw) is a plain, swizzle-atom-aligned sub-view:
//     slab (d + kd - 1) -> copy kw -> atom row kh .. kh + 15      (atom = 8 consecutive w voxels = 8 rows of KC*2 bytes)
// i.e. no per-tap loads at all: L2->SM traffic drops from 27 to 3.4 fetches per input voxel.  Weights stay resident in
// shared memory for the whole CTA when they fit (27 * N * K * 2 B <= 110 KB), otherwise they stream through a small ring.
// tcgen05.mma M = 128 (16 h x 8 w), N = Nout, K = 16; accumulators double-buffered in TMEM; same warp roles as
// conv3d_tc.cu plus an optional weight-producer warp.
#include <string.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace b2 {

struct HaloParams {
    int N, D, H, W;       // produced tensor == gathered tensor extent (stride 1, padding 1)
    int dst_pitch;
    int BN;               // output channels (<= 256)
    int HB, WB, DS, SD;   // strips per sample along h / w, segments along d and their length
    int num_items;
    int b_resident;       // weights resident in shared memory
    int b_stages;         // ring depth when streaming
    uint32_t idesc;
    uint32_t tmem_cols;
};

constexpr int HALO_THREADS = 224;  // warp 0: slab producer, 1: MMA, 2..5: epilogue, 6: weight producer
constexpr int HALO_MAX_BSTAGES = 8;

template <int KC>
__device__ __forceinline__ uint64_t halo_desc(uint32_t saddr) {
    constexpr uint32_t row_bytes = KC * 2;
    constexpr uint64_t layout = row_bytes == 128 ? 2 : 4;
    constexpr uint64_t sbo = (8 * row_bytes) >> 4;
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= sbo << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <int KC>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, HaloParams p,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst, int accumulate) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t sfull[3], sempty[3], bfull[HALO_MAX_BSTAGES], bempty[HALO_MAX_BSTAGES], wfull, tfull[2], tempty[2];
    __shared__ uint32_t tmem_base_smem;

    constexpr uint32_t ROW = KC * 2;
    constexpr uint32_t COPY_BYTES = 144 * ROW;       // 18 h x 8 w rows
    constexpr uint32_t SLAB_BYTES = 3 * COPY_BYTES;  // three w-shifted copies
    const uint32_t B_TILE = (uint32_t)p.BN * ROW;    // one tap of weights [BN][KC]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* slabs = smem;
    uint8_t* wsm = smem + 3 * SLAB_BYTES;            // resident: 27 tiles; streaming: b_stages tiles

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < 3; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 1); }
        for (int s = 0; s < HALO_MAX_BSTAGES; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); }
        mbar_init(&wfull, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    // item -> (n, hb, wb, ds)
    auto decode = [&](int item, int& n, int& h0, int& w0, int& d0, int& d1) {
        int t = item;
        const int ds = t % p.DS; t /= p.DS;
        const int wb = t % p.WB; t /= p.WB;
        const int hb = t % p.HB; t /= p.HB;
        n = t; h0 = hb * 16; w0 = wb * 8; d0 = ds * p.SD; d1 = d0 + p.SD < p.D ? d0 + p.SD : p.D;
    };

    if (warp == 0) {
        // ===================== slab producer =====================
        if (lane == 0) {
            uint32_t j = 0;  // running slab counter -> slot j % 3, phase (j / 3) & 1
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
                int n, h0, w0, d0, d1;
                decode(item, n, h0, w0, d0, d1);
                for (int s = d0 - 1; s <= d1; ++s, ++j) {
                    const uint32_t slot = j % 3, ph = (j / 3) & 1;
                    mbar_wait(&sempty[slot], ph ^ 1);
                    mbar_expect_tx(&sfull[slot], SLAB_BYTES);
                    uint8_t* base = slabs + slot * SLAB_BYTES;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw)
                        tma_load_5d(&tmA, &sfull[slot], base + kw * COPY_BYTES, 0, w0 + kw - 1, h0 - 1, s, n);
                }
            }
        }
    } else if (warp == 6) {
        // ===================== weight producer =====================
        if (lane == 0) {
            mbar_expect_tx(&wfull, 27 * B_TILE);
            for (int t = 0; t < 27; ++t) tma_load_2d(&tmB, &wfull, wsm + (size_t)t * B_TILE, 0, t * p.BN);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: ONE thread, straight-line code =====================
        // The issuing thread is blocked ~45 cycles per tcgen05.mma (M128 x N<=64 x K16, measured with tools/mma_probe), and
        // every scalar instruction between two MMAs adds to that: so descriptors are assembled from precomputed 32-bit
        // halves and the 27 taps are fully unrolled (all offsets are immediates).
        {
            constexpr uint32_t DESC_HI = (uint32_t)(((uint64_t)((8 * ROW) >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)(ROW == 128 ? 2 : 4) << 61) >> 32);
            auto mk = [](uint32_t lo) -> uint64_t { return ((uint64_t)DESC_HI << 32) | (uint64_t)lo; };
            const uint32_t slab_lo0 = ((smem_u32(slabs) & 0x3FFFF) >> 4) | 0x10000u;
            constexpr uint32_t SLAB16 = SLAB_BYTES >> 4;
            const uint32_t w_lo = ((smem_u32(wsm) & 0x3FFFF) >> 4) | 0x10000u;
            const uint32_t btile16 = B_TILE >> 4;
            uint32_t jrel = 0;   // slab counter (mod 3) of slab (d0 - 1) of the current item
            uint32_t jcnt = 0;   // absolute slab counter at the start of the item (for the barrier phases)
            int acc = 0;
            uint32_t acc_phase = 0;
            mbar_wait(&wfull, 0);
            for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
                int n, h0, w0, d0, d1;
                decode(item, n, h0, w0, d0, d1);
                uint32_t waited = 0;
                uint32_t s0 = jrel;   // slot of slab (od - 1)
                for (int od = d0; od < d1; ++od) {
                    const uint32_t need = (uint32_t)(od - d0) + 3;
                    while (waited < need) {
                        const uint32_t j = jcnt + waited;
                        mbar_wait(&sfull[j % 3], (j / 3) & 1);
                        ++waited;
                    }
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
                    const uint32_t s1 = s0 == 2 ? 0 : s0 + 1, s2 = s1 == 2 ? 0 : s1 + 1;
                    const uint32_t lo_kd0 = slab_lo0 + s0 * SLAB16, lo_kd1 = slab_lo0 + s1 * SLAB16, lo_kd2 = slab_lo0 + s2 * SLAB16;
                    uint32_t b_lo = w_lo;
                    if (elect_one()) {
#pragma unroll
                    for (int t = 0; t < 27; ++t) {
                        constexpr uint32_t C16 = COPY_BYTES >> 4, A16 = (8 * ROW) >> 4;
                        const uint32_t a_lo = (t / 9 == 0 ? lo_kd0 : t / 9 == 1 ? lo_kd1 : lo_kd2) + (uint32_t)(t % 3) * C16 + (uint32_t)((t / 3) % 3) * A16;
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k)
                            umma_bf16(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), p.idesc, (t | k) != 0 ? 1u : 0u);
                        b_lo += btile16;
                        if (t == 8) umma_commit(&sempty[s0]);          // slab od-1: last use issued
                    }
                    umma_commit(&tfull[acc]);
                    if (od == d1 - 1) { umma_commit(&sempty[s1]); umma_commit(&sempty[s2]); }
                    }
                    __syncwarp();
                    s0 = s1;
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                const uint32_t used = (uint32_t)(d1 - d0 + 2);
                jcnt += used;
                jrel = (jrel + used) % 3;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int w_ = r & 7, h_ = r >> 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
            int n, h0, w0, d0, d1;
            decode(item, n, h0, w0, d0, d1);
            const int oh = h0 + h_, ow = w0 + w_;
            const bool valid = oh < p.H && ow < p.W;
            for (int od = d0; od < d1; ++od) {
                __nv_bfloat16* row = dst + ((((long long)n * p.D + od) * p.H + oh) * p.W + ow) * p.dst_pitch;
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
                for (int c0 = 0; c0 < p.BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int jj = 0; jj < 32; jj += 8) {
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                f[e] = __uint_as_float(v[jj + e]);
                                if (bias) f[e] += bias[c0 + jj + e];
                            }
                            if (accumulate) {
                                float o[8];
                                load8(row + c0 + jj, o);
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] += o[e];
                            }
                            store8(row + c0 + jj, f);
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

int make_w_map(CUtensorMap* m, const void* ptr, int rows, int K, int KC, int BN);

static bool halo_plan(int K, int Nout, int D, int H, int W, int N, HaloParams& p, size_t& smem) {
    if (!(K == 32 || K == 64) || Nout % 32 != 0 || Nout > 256) return false;
    if (H < 16 || W < 8) return false;
    memset(&p, 0, sizeof(p));
    p.N = N; p.D = D; p.H = H; p.W = W; p.BN = Nout;
    p.HB = cdiv(H, 16); p.WB = cdiv(W, 8);
    const uint32_t row = K * 2, slab = 3 * 144 * row, btile = (uint32_t)Nout * row;
    const size_t budget = 210 * 1024;
    // weights must be resident (27 taps): streaming them costs one mbarrier round trip per tap on the MMA thread, which is
    // slower than the per-tap kernel of conv3d_tc.cu
    if (3ull * slab + 27ull * btile + 1024 > budget) return false;
    p.b_resident = 1; p.b_stages = 0; smem = 3ull * slab + 27ull * btile + 1024;
    // segment length: enough items for >= 4 waves when possible, at least 8 slabs per segment
    const int strips = N * p.HB * p.WB;
    int sd = D;
    while (sd > 8 && strips * cdiv(D, sd) < 4 * num_sms()) sd = (sd + 1) / 2;
    p.SD = sd; p.DS = cdiv(D, sd);
    p.num_items = strips * p.DS;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Nout >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * Nout)) cols *= 2;
    p.tmem_cols = cols;
    return true;
}

int g_use_halo = 1;

bool conv_tc_halo_supported(int K, int Nout, int N, int D, int H, int W) {
    HaloParams p;
    size_t smem;
    return g_use_halo && halo_plan(K, Nout, D, H, W, N, p, smem);
}

// src / dst: NDHWC bf16 of identical spatial extent; wmat: [27][Nout][K] bf16 (forward shadow, or the flipped shadow for dgrad)
int conv_tc_halo_launch(const __nv_bfloat16* src, int N, int D, int H, int W, int K, int src_pitch, const __nv_bfloat16* wmat, int Nout,
                        const float* bias, __nv_bfloat16* dst, int dst_pitch, int accumulate, cudaStream_t st) {
    HaloParams p;
    size_t smem;
    if (!halo_plan(K, Nout, D, H, W, N, p, smem)) return fail(B2_EUNSUPPORTED, "conv_tc_halo: unsupported shape%s", "");
    B2_CHECK_ARG(src_pitch % 8 == 0 && dst_pitch % 8 == 0);
    p.dst_pitch = dst_pitch;
    CUtensorMap tmA, tmB;
    int rc = make_act_map(&tmA, src, N, D, H, W, K, src_pitch, K, 1, 1, 18, 8, 1, 1, 1);
    if (rc) return rc;
    rc = make_w_map(&tmB, wmat, 27 * Nout, K, K, Nout);
    if (rc) return rc;
    const int grid = p.num_items < num_sms() ? p.num_items : num_sms();
    static bool a32 = false, a64 = false;
    if (K == 32) {
        if (!a32) { B2_CUDA(cudaFuncSetAttribute(conv_halo_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); a32 = true; }
        B2_LAUNCH(conv_halo_kernel<32>, grid, HALO_THREADS, smem, st, tmA, tmB, p, bias, dst, accumulate);
    } else {
        if (!a64) { B2_CUDA(cudaFuncSetAttribute(conv_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); a64 = true; }
        B2_LAUNCH(conv_halo_kernel<64>, grid, HALO_THREADS, smem, st, tmA, tmB, p, bias, dst, accumulate);
    }
    return B2_OK;
}

}  // namespace b2
