// conv3d_tc.cu -- bf16 3x3x3 convolution (forward and stride-1 dgrad) as an implicit GEMM on the 5th-generation tensor
// cores: TMA (cp.async.bulk.tensor) stages NDHWC activation boxes and K-major weight tiles into 128B/64B-swizzled shared
// memory, one elected thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) with the accumulator in TMEM, and four
// epilogue warps drain TMEM with tcgen05.ld, add the bias, convert to bf16 and store.  Persistent CTAs (one per SM),
// warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 = epilogue; the
// accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// GEMM view:  D[m, n] = sum_{tap, c} A_tap[m, c] * W[tap][n][c]
//   m   : 128 output voxels = one (TN,TD,TH,TW) box of the NDHWC tensor (rows in w-fastest order)
//   A   : the same box shifted by (kd-1, kh-1, kw-1) [times the conv stride: TMA elementStrides]; out-of-bounds voxels
//         are zero-filled by TMA, which IS the conv zero padding -- no im2col, no halo code
//   n   : output channels (BN <= 256 per CTA tile); K per pipeline stage: KC = 64 (SWIZZLE_128B) or 32 (SWIZZLE_64B)
// The same kernel computes the stride-1 dgrad (A = dz, W = flipped/transposed shadow).
//
// Replaces cuDNN's conv3d behind nn.Conv3d of nnunet's ConvDropoutNormNonlin (reference call sites: the
// `self.network(data)` of nnUNetTrainerMultiHead.py:621/633 and its backward :627/639).
#include "tc_common.cuh"
#include "kernels.h"

namespace b2 {

// K-major UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor, sm_100 version 1): rows of KC*2 bytes, 8-row
// swizzle atoms; SBO = stride between 8-row groups; LBO unused for swizzled K-major.
template <int KC>
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
    constexpr uint32_t row_bytes = KC * 2;               // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
    constexpr uint64_t layout = row_bytes == 128 ? 2 : 4;  // UMMA::LayoutType
    constexpr uint64_t sbo = (8 * row_bytes) >> 4;
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                   // leading byte offset (ignored for swizzled K-major), bits [16,30)
    d |= sbo << 32;                           // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
    d |= layout << 61;                        // layout type, bits [61,64)
    return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): bf16 x bf16 -> f32, A and B K-major
__host__ __device__ inline uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------
struct TcConvParams {
    int N, D, H, W;              // extent of the PRODUCED tensor
    int dst_pitch;
    int TN, TD, TH, TW;          // voxel box of a tile (product == 128)
    int nt_n, nt_d, nt_h, nt_w;  // tiles per axis
    int nblk, BN;                // output-channel blocks and their width
    int kchunks;                 // K / KC
    int rows_per_tap;            // rows of the weight matrix per tap (== total N)
    int sd, sh, sw;              // conv stride (source coordinate = out * s + tap - 1)
    int stages;
    int num_tiles;
    uint32_t idesc;
    uint32_t tmem_cols;
};

constexpr int TC_THREADS = 192;
constexpr int MAX_STAGES = 16;

template <int KC>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcConvParams p,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst, int accumulate) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t A_BYTES = 128 * KC * 2;
    const uint32_t B_BYTES = (uint32_t)p.BN * KC * 2;
    const uint32_t STAGE_BYTES = A_BYTES + ((B_BYTES + 1023) & ~1023u);
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    const int kiters = 27 * p.kchunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                int t = tile;
                const int nb = t % p.nblk; t /= p.nblk;
                const int tw = t % p.nt_w; t /= p.nt_w;
                const int th = t % p.nt_h; t /= p.nt_h;
                const int td = t % p.nt_d; t /= p.nt_d;
                const int tn = t;
                const int w0 = tw * p.TW * p.sw - 1, h0 = th * p.TH * p.sh - 1, d0 = td * p.TD * p.sd - 1, n0 = tn * p.TN;
                for (int it = 0; it < kiters; ++it) {
                    const int tap = it / p.kchunks, kc = it % p.kchunks;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
                    uint8_t* sb = sa + A_BYTES;
                    mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
                    tma_load_5d(&tmA, &full_bar[stage], sa, kc * KC, w0 + tap % 3, h0 + (tap / 3) % 3, d0 + tap / 9, n0);
                    tma_load_2d(&tmB, &full_bar[stage], sb, kc * KC, tap * p.rows_per_tap + nb * p.BN);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
            for (int it = 0; it < kiters; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    const uint32_t sb = sa + A_BYTES;
                    const uint64_t adesc = umma_desc_kmajor<KC>(sa);
                    const uint64_t bdesc = umma_desc_kmajor<KC>(sb);
#pragma unroll
                    for (int k = 0; k < KC / 16; ++k) {
                        // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
                        umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc, (it | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (it == kiters - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue: TMEM -> registers -> (+bias) -> bf16 -> global =====================
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access
        const int r = q * 32 + lane;            // accumulator row == voxel inside the box
        const int w_ = r % p.TW, h_ = (r / p.TW) % p.TH, d_ = (r / (p.TW * p.TH)) % p.TD, n_ = r / (p.TW * p.TH * p.TD);
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            int t = tile;
            const int nb = t % p.nblk; t /= p.nblk;
            const int tw = t % p.nt_w; t /= p.nt_w;
            const int th = t % p.nt_h; t /= p.nt_h;
            const int td = t % p.nt_d; t /= p.nt_d;
            const int tn = t;
            const int ow = tw * p.TW + w_, oh = th * p.TH + h_, od = td * p.TD + d_, on = tn * p.TN + n_;
            const bool valid = ow < p.W && oh < p.H && od < p.D && on < p.N;
            __nv_bfloat16* row = dst + ((((long long)on * p.D + od) * p.H + oh) * p.W + ow) * p.dst_pitch + nb * p.BN;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            f[e] = __uint_as_float(v[j + e]);
                            if (bias) f[e] += bias[nb * p.BN + c0 + j + e];
                        }
                        if (accumulate) {
                            float o[8];
                            load8(row + c0 + j, o);
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] += o[e];
                        }
                        store8(row + c0 + j, f);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------------------
EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 5D activation map over an NDHWC view: dims (C, W, H, D, N)
int make_act_map(CUtensorMap* m, const void* ptr, int N, int D, int H, int W, int C, int pitch, int KC, int TN, int TD,
                        int TH, int TW, int sd, int sh, int sw) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(B2_ECUDA, "cuTensorMapEncodeTiled entry point not available%s", "");
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2, (cuuint64_t)D * H * W * pitch * 2};
    // box = bounding box in (unstrided) elements; with elementStrides s the box holds ceil(box/s) elements per axis
    cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)((TW - 1) * sw + 1), (cuuint32_t)((TH - 1) * sh + 1), (cuuint32_t)((TD - 1) * sd + 1), (cuuint32_t)TN};
    cuuint32_t estr[5] = {1, (cuuint32_t)sw, (cuuint32_t)sh, (cuuint32_t)sd, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B2_ECUDA, "cuTensorMapEncodeTiled(activation) failed: code %s%lld", "", (long long)r);
    return B2_OK;
}

// 2D weight map: rows = taps * Ntotal, cols = K (K contiguous)
static int make_w_map(CUtensorMap* m, const void* ptr, int rows, int K, int KC, int BN) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(B2_ECUDA, "cuTensorMapEncodeTiled entry point not available%s", "");
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B2_ECUDA, "cuTensorMapEncodeTiled(weights) failed: code %s%lld", "", (long long)r);
    return B2_OK;
}

static int pow2_le(int v, int cap) {
    int p = 1;
    while (p * 2 <= v && p * 2 <= cap) p *= 2;
    return p;
}

bool conv_tc_supported(int K, int Nout) {
    if (K % 32 != 0 || Nout % 32 != 0) return false;
    return true;
}

// PyTorch [Cout][Cin][27] fp32 -> Wk [27][Cout][Cin] bf16 (forward) and Wd [27 flipped][Cin][Cout] bf16 (stride-1 dgrad)
__global__ void weight_shadow_bf16_kernel(const float* __restrict__ w, int Cout, int Cin, __nv_bfloat16* __restrict__ wk,
                                          __nv_bfloat16* __restrict__ wd) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)Cout * Cin * 27;
    if (i >= tot) return;
    int t = (int)(i % 27);
    long long r = i / 27;
    int ci = (int)(r % Cin), co = (int)(r / Cin);
    __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (wk) wk[((long long)t * Cout + co) * Cin + ci] = v;
    if (wd) wd[((long long)(26 - t) * Cin + ci) * Cout + co] = v;
}

int weight_shadow_bf16(const float* w, int cout, int cin, __nv_bfloat16* wk, __nv_bfloat16* wd, cudaStream_t st) {
    long long tot = (long long)cout * cin * 27;
    B2_LAUNCH(weight_shadow_bf16_kernel, cdiv(tot, 256), 256, 0, st, w, cout, cin, wk, wd);
    return B2_OK;
}

// Generic launcher.  src: NDHWC bf16 [N, Ds, Hs, Ws, K] (pitch src_pitch); wmat: [27][Nout][K] bf16; dst: NDHWC bf16
// [N, Dd, Hd, Wd, Nout] with source coordinate = dst * stride + tap - 1.
int conv_tc_launch(const __nv_bfloat16* src, int N, int Ds, int Hs, int Ws, int K, int src_pitch, const __nv_bfloat16* wmat,
                   int Nout, const float* bias, __nv_bfloat16* dst, int Dd, int Hd, int Wd, int dst_pitch, const int stride[3],
                   int accumulate, cudaStream_t st) {
    B2_CHECK_ARG(conv_tc_supported(K, Nout));
    B2_CHECK_ARG(src_pitch % 8 == 0 && dst_pitch % 8 == 0);
    const int KC = (K % 64 == 0) ? 64 : 32;
    int nblk = cdiv(Nout, 256);
    while (Nout % nblk != 0 || (Nout / nblk) % 32 != 0) ++nblk;
    const int BN = Nout / nblk;
    TcConvParams p;
    p.N = N; p.D = Dd; p.H = Hd; p.W = Wd; p.dst_pitch = dst_pitch;
    p.TW = pow2_le(Wd, 8);
    p.TH = pow2_le(Hd, 128 / p.TW > 8 ? 8 : 128 / p.TW);
    p.TD = pow2_le(Dd, 128 / (p.TW * p.TH));
    p.TN = 128 / (p.TW * p.TH * p.TD);
    p.nt_w = cdiv(Wd, p.TW); p.nt_h = cdiv(Hd, p.TH); p.nt_d = cdiv(Dd, p.TD); p.nt_n = cdiv(N, p.TN);
    p.nblk = nblk; p.BN = BN; p.kchunks = K / KC; p.rows_per_tap = Nout;
    p.sd = stride[0]; p.sh = stride[1]; p.sw = stride[2];
    p.num_tiles = p.nt_w * p.nt_h * p.nt_d * p.nt_n * nblk;
    p.idesc = umma_idesc_bf16(128, BN);
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * BN)) cols *= 2;
    p.tmem_cols = cols;
    const uint32_t a_bytes = 128 * KC * 2, b_bytes = ((uint32_t)BN * KC * 2 + 1023) & ~1023u;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return fail(B2_EUNSUPPORTED, "conv_tc: tile does not fit shared memory%s", "");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + 1024;

    CUtensorMap tmA, tmB;
    int rc = make_act_map(&tmA, src, N, Ds, Hs, Ws, K, src_pitch, KC, p.TN, p.TD, p.TH, p.TW, p.sd, p.sh, p.sw);
    if (rc) return rc;
    rc = make_w_map(&tmB, wmat, 27 * Nout, K, KC, BN);
    if (rc) return rc;

    static bool attr64 = false, attr32 = false;
    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    if (KC == 64) {
        if (!attr64) { B2_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr64 = true; }
        B2_LAUNCH(conv_tc_kernel<64>, grid, TC_THREADS, smem, st, tmA, tmB, p, bias, dst, accumulate);
    } else {
        if (!attr32) { B2_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr32 = true; }
        B2_LAUNCH(conv_tc_kernel<32>, grid, TC_THREADS, smem, st, tmA, tmB, p, bias, dst, accumulate);
    }
    return B2_OK;
}

}  // namespace b2
