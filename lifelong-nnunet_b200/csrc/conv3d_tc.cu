// conv3d_tc.cu -- bf16 3x3x3 convolution (forward and stride-1 dgrad) as an implicit GEMM on the 5th-generation tensor
// cores: TMA (cp.async.bulk.tensor) stages NDHWC activation boxes and K-major weight tiles into 128B/64B-swizzled shared
// memory, one elected thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) with the accumulator in TMEM, and the
// epilogue warps drain TMEM with tcgen05.ld, add the bias, convert to bf16 and store.  Persistent CTAs (one per SM),
// warp-specialised: warps 0..3 = TMA producers (one pipeline stage each, round robin), warp 4 = MMA issuer (+ TMEM
// allocation), warps 5..12 = two epilogue sets; the accumulator is double-buffered in TMEM (set s drains buffer s) so the
// epilogue of tile i overlaps the MMAs of tile i+1 and the scalar-heavy epilogues of two short-K tiles overlap each other.
// The thin stride-1 layers (K = 32 / 64) do not run here but in conv3d_tc_halo.cu (input-stationary, kd-merged).
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
#include <string.h>

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
    int N, D, H, W;              // extent of the PRODUCED tensor (bounds of the stores)
    int LD, LH, LW;              // logical grid the voxel boxes tile (== D,H,W unless the output is a strided sub-lattice)
    int dst_pitch;
    int TN, TD, TH, TW;          // voxel box of a tile (product == 128)
    int nt_n, nt_d, nt_h, nt_w;  // tiles per axis
    int nblk, BN;                // output-column blocks and their width
    int kchunks;                 // K / KC
    int rows_per_tap;            // rows of the weight matrix per tap block
    int w_row0;                  // first row of the computed output-channel window inside a tap block
    int sd, sh, sw;              // source coordinate = logical * s + tap_off
    int os_d, os_h, os_w;        // produced coordinate = logical * os + oo (+ q offset when q_scatter)
    int oo_d, oo_h, oo_w;
    int q_scatter, q_channels, qk_h, qk_w;   // transposed-conv forward: GEMM column -> (q = col / q_channels, channel = col % q_channels)
    int ntaps;
    int out_f32;                 // 1: the produced tensor is fp32 (plain GEMM use: weight gradients of the ViT linears)
    int stages;
    int num_tiles;               // output tiles x ksplit
    int nprod;                   // active TMA producer warps (<= stages, so that a producer can never lap the ring)
    int ksplit;                  // > 1: the (tap, k-chunk) iterations of a tile are split over ksplit CTAs writing fp32 partials
    uint32_t idesc;
    uint32_t tmem_cols;
    signed char tap_off[27][3];  // source offset (d, h, w) per tap, padding included
    unsigned char tap_w[27];     // weight row-block of the tap
    // tile classes (strided dgrad: one class per output-parity lattice, all in ONE launch).  nclass <= 1: a single class
    // described by the fields above.  Class c owns the tiles [cls_tile0[c], cls_tile0[c+1]), the taps
    // [cls_tap0[c], +cls_ntaps[c]) of the tables above, its own lattice offset and logical extent.
    int nclass;
    int cls_tile0[9];
    unsigned char cls_tap0[8], cls_ntaps[8];
    signed char cls_oo[8][3];
    short cls_L[8][3], cls_nt[8][3];
    // shared-A stages (merged-class strided dgrad): a pipeline stage is ONE source tile (a shift of the dz box) plus the weight
    // blocks of every (parity class, tap) pair that reads it; block j of stage s accumulates into the column block of its
    // class.  A tile of the coarse lattice then produces all parity classes at once: 8 source tiles per output tile instead
    // of 27 (the per-class launches were L2-bound on re-fetching dz), and the epilogue scatters column block -> class
    // lattice with the q_scatter mapping.
    int mes, mes_nst, mes_blk_rows;
    unsigned char mes_nb[12], mes_blk0[13];
    unsigned char mes_wrow[27], mes_cls[27], mes_first[27];
    // magic numbers for the index decode (FastDiv, tc_common.cuh)
    FastDiv fd_ksplit, fd_nblk, fd_ntw, fd_nth, fd_ntd, fd_qch, fd_qkw, fd_qkh;
    FastDiv cls_fd[8][3];        // per class: nt_d, nt_h, nt_w
};

struct TileInfo {
    int ks, nb, tw, th, td, tn, otile;
    int tap0, ntaps;
    int oo_d, oo_h, oo_w, LD, LH, LW;
};

// class of a tile (tile classes only) and its tap count -- all the MMA warp needs
__device__ __forceinline__ int tile_class(const TcConvParams& p, int tile) {
    int c = 0;
    while (c + 1 < p.nclass && tile >= p.cls_tile0[c + 1]) ++c;
    return c;
}

__device__ __forceinline__ void decode_tile(const TcConvParams& p, int tile, TileInfo& ti) {
    int t = tile;
    ti.tap0 = 0; ti.ntaps = p.ntaps;
    ti.oo_d = p.oo_d; ti.oo_h = p.oo_h; ti.oo_w = p.oo_w;
    ti.LD = p.LD; ti.LH = p.LH; ti.LW = p.LW;
    if (p.nclass > 1) {
        const int c = tile_class(p, tile);
        t = tile - p.cls_tile0[c];
        ti.tap0 = p.cls_tap0[c]; ti.ntaps = p.cls_ntaps[c];
        ti.oo_d = p.cls_oo[c][0]; ti.oo_h = p.cls_oo[c][1]; ti.oo_w = p.cls_oo[c][2];
        ti.LD = p.cls_L[c][0]; ti.LH = p.cls_L[c][1]; ti.LW = p.cls_L[c][2];
        ti.ks = 0;                                   // (no split-K with tile classes)
        ti.otile = t;
        t = fd_divmod(t, p.fd_nblk, ti.nb);
        t = fd_divmod(t, p.cls_fd[c][2], ti.tw);
        t = fd_divmod(t, p.cls_fd[c][1], ti.th);
        ti.tn = fd_divmod(t, p.cls_fd[c][0], ti.td);
        return;
    }
    t = fd_divmod(t, p.fd_ksplit, ti.ks);
    ti.otile = t;
    t = fd_divmod(t, p.fd_nblk, ti.nb);
    t = fd_divmod(t, p.fd_ntw, ti.tw);
    t = fd_divmod(t, p.fd_nth, ti.th);
    ti.tn = fd_divmod(t, p.fd_ntd, ti.td);
}

constexpr int TC_PRODUCERS = 4;                       // TMA producer warps (one stage each, round robin): the single-thread
                                                      // issue latency (~300 cycles per stage) was the bottleneck of round-1 v1
constexpr int TC_EPI_SETS = 2;                        // epilogue warp sets: set s drains TMEM accumulator s (tiles alternate), so
                                                      // the scalar-heavy epilogue of short-K tiles runs two tiles at a time
constexpr int TC_THREADS = 32 * (TC_PRODUCERS + 1 + 4 * TC_EPI_SETS);   // producers, 1 MMA warp, 2 x 4 epilogue warps
constexpr int MAX_STAGES = 16;

template <int KC, bool MES>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcConvParams p,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst, int accumulate, float* __restrict__ partial,
               EpiStats es) {
    pdl_grid_sync();
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
    if (warp == TC_PRODUCERS) tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < TC_PRODUCERS) {
        // (warps >= p.nprod idle)
        // ===================== TMA producers: warp w issues the pipeline iterations with git % TC_PRODUCERS == w =====================
        if (lane == 0 && warp < p.nprod) {
            // counters instead of divisions: gmod = git % nprod; (stage, phase) of this warp's next iteration
            int gmod = 0, stage = warp;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                TileInfo ti;
                decode_tile(p, tile, ti);
                const int nb = ti.nb;
                const int w0 = ti.tw * p.TW * p.sw, h0 = ti.th * p.TH * p.sh, d0 = ti.td * p.TD * p.sd, n0 = ti.tn * p.TN;
                const int kiters = (MES ? p.mes_nst : ti.ntaps) * p.kchunks;
                int it0 = 0, it1 = kiters, tap = ti.tap0, kc = 0;
                if (p.ksplit > 1) {
                    it0 = (int)((long long)kiters * ti.ks / p.ksplit); it1 = (int)((long long)kiters * (ti.ks + 1) / p.ksplit);
                    tap = ti.tap0 + it0 / p.kchunks; kc = it0 % p.kchunks;
                }
                for (int it = it0; it < it1; ++it) {
                    if (gmod == warp) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
                        if (MES) {
                            const int nbk = p.mes_nb[tap], b0 = p.mes_blk0[tap];
                            const uint32_t blk_bytes = (uint32_t)p.mes_blk_rows * KC * 2;
                            mbar_expect_tx(&full_bar[stage], A_BYTES + (uint32_t)nbk * blk_bytes);
                            tma_load_5d(&tmA, &full_bar[stage], sa, kc * KC, w0 + p.tap_off[tap][2], h0 + p.tap_off[tap][1],
                                        d0 + p.tap_off[tap][0], n0);
                            for (int j = 0; j < nbk; ++j)
                                tma_load_2d(&tmB, &full_bar[stage], sa + A_BYTES + (size_t)j * blk_bytes, kc * KC,
                                            (int)p.mes_wrow[b0 + j] * p.rows_per_tap);
                        } else {
                        mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
                        tma_load_5d(&tmA, &full_bar[stage], sa, kc * KC, w0 + p.tap_off[tap][2], h0 + p.tap_off[tap][1],
                                    d0 + p.tap_off[tap][0], n0);
                        tma_load_2d(&tmB, &full_bar[stage], sa + A_BYTES, kc * KC, (int)p.tap_w[tap] * p.rows_per_tap + p.w_row0 + nb * p.BN);
                        }
                        stage += p.nprod;
                        if (stage >= p.stages) { stage -= p.stages; phase ^= 1; }
                    }
                    if (++gmod == p.nprod) gmod = 0;
                    if (++kc == p.kchunks) { kc = 0; ++tap; }
                }
            }
        }
    } else if (warp == TC_PRODUCERS) {
        // ===================== MMA issuer (one thread) =====================
        {
            constexpr uint32_t ROWB = KC * 2;
            constexpr uint32_t DESC_HI = (uint32_t)(((uint64_t)((8 * ROWB) >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)(ROWB == 128 ? 2 : 4) << 61) >> 32);
            auto mk = [](uint32_t lo) -> uint64_t { return ((uint64_t)DESC_HI << 32) | (uint64_t)lo; };
            const uint32_t base_lo = ((smem_u32(smem) & 0x3FFFF) >> 4) | 0x10000u;
            const uint32_t stage16 = STAGE_BYTES >> 4, a16 = A_BYTES >> 4;
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0, a_lo = base_lo;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
                int n_it = p.ntaps * p.kchunks;            // (the MMA warp only needs the iteration count of the tile)
                if (MES) n_it = p.mes_nst * p.kchunks;
                else if (p.nclass > 1) n_it = p.cls_ntaps[tile_class(p, tile)] * p.kchunks;
                else if (p.ksplit > 1) {
                    int ks;
                    fd_divmod(tile, p.fd_ksplit, ks);
                    n_it = (int)((long long)n_it * (ks + 1) / p.ksplit) - (int)((long long)n_it * ks / p.ksplit);
                }
                int mes_s = 0, mes_kc = 0;
                for (int it = 0; it < n_it; ++it) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        if (MES) {
                            const int nbk = p.mes_nb[mes_s], b0 = p.mes_blk0[mes_s];
                            const uint32_t blk16 = ((uint32_t)p.mes_blk_rows * KC * 2) >> 4;
                            for (int j = 0; j < nbk; ++j) {
                                const uint32_t dj = d_tmem + (uint32_t)p.mes_cls[b0 + j] * (uint32_t)p.mes_blk_rows;
                                const uint32_t fresh = (p.mes_first[b0 + j] && mes_kc == 0) ? 1u : 0u;
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k)
                                    umma_bf16(dj, mk(a_lo + 2 * k), mk(a_lo + a16 + (uint32_t)j * blk16 + 2 * k), p.idesc,
                                              (fresh && k == 0) ? 0u : 1u);
                            }
                        } else {
#pragma unroll
                        for (int k = 0; k < KC / 16; ++k)   // +32 bytes along K inside the swizzle atom per K16 step
                            umma_bf16(d_tmem, mk(a_lo + 2 * k), mk(a_lo + a16 + 2 * k), p.idesc, (it | k) != 0 ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[stage]);
                        if (it == n_it - 1) umma_commit(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    a_lo += stage16;
                    if (++stage == p.stages) { stage = 0; phase ^= 1; a_lo = base_lo; }
                    if (MES) { if (++mes_kc == p.kchunks) { mes_kc = 0; ++mes_s; } }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: TMEM -> registers -> (+bias) -> bf16 -> global =====================
        const int q = warp & 3;                 // TMEM lane quadrant this warp may access (4 consecutive warps: 4 quadrants)
        const int ew = warp - (TC_PRODUCERS + 1), set = ew >> 2;
        const int r = q * 32 + lane;            // accumulator row == voxel inside the box
        const int w_ = r % p.TW, h_ = (r / p.TW) % p.TH, d_ = (r / (p.TW * p.TH)) % p.TD, n_ = r / (p.TW * p.TH * p.TD);
        const int acc = set;                    // tile i of this CTA uses accumulator i & 1 and is drained by set i & 1
        uint32_t acc_phase = 0;
        int tile_it = 0;
        // InstanceNorm statistics from the fp32 accumulators (host enables this only for TN == 1, one column block, no split-K):
        // lane = column of each 32-column chunk; flushed whenever the CTA moves on to the next sample
        float st1[8], st2[8];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) { st1[ch] = 0.f; st2[ch] = 0.f; }
        int n_cur = -1, n_done = 0;
        const int st_slot = (int)blockIdx.x * (4 * TC_EPI_SETS) + ew;
        auto st_write = [&](int nn, bool zero) {
            for (int ch = 0; ch < p.BN / 32; ++ch) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (k == ch && !zero) { a1 = st1[k]; a2 = st2[k]; }
                *reinterpret_cast<float2*>(es.part + (((long long)nn * es.slots + st_slot) * es.C + ch * 32 + lane) * 2) = make_float2(a1, a2);
            }
        };
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tile_it) {
            if ((tile_it & 1) != set) continue;
            TileInfo ti;
            decode_tile(p, tile, ti);
            const int ks = ti.ks, otile = ti.otile, nb = ti.nb;
            if (es.part && ti.tn != n_cur) {
                if (n_cur >= 0) { st_write(n_cur, false); n_done = n_cur + 1; }
                for (int nz = n_done; nz < ti.tn; ++nz) st_write(nz, true);
                n_done = ti.tn; n_cur = ti.tn;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) { st1[ch] = 0.f; st2[ch] = 0.f; }
            }
            const int lw = ti.tw * p.TW + w_, lh = ti.th * p.TH + h_, ld = ti.td * p.TD + d_, on = ti.tn * p.TN + n_;
            const int ow0 = lw * p.os_w + ti.oo_w, oh0 = lh * p.os_h + ti.oo_h, od0 = ld * p.os_d + ti.oo_d;
            const bool in_grid = lw < ti.LW && lh < ti.LH && ld < ti.LD && on < p.N;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
            for (int c0 = 0; c0 < p.BN; c0 += 32) {
                // column chunk -> produced voxel / channel (transposed conv: a tile spans several q, one per chunk group)
                int ow = ow0, oh = oh0, od = od0, chan0 = nb * p.BN;
                if (p.q_scatter) {
                    int chq, qw, qh;
                    const int col = nb * p.BN + c0, qq = fd_divmod(col, p.fd_qch, chq);
                    chan0 = chq - c0;
                    const int qd = fd_divmod(fd_divmod(qq, p.fd_qkw, qw), p.fd_qkh, qh);
                    ow += qw; oh += qh; od += qd;
                }
                const bool valid = in_grid && ow < p.W && oh < p.H && od < p.D;
                __nv_bfloat16* row = dst + ((((long long)on * p.D + od) * p.H + oh) * p.W + ow) * p.dst_pitch + chan0;
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                tmem_ld_wait();
                if (p.ksplit > 1) {
                    // fp32 partial [ks][output tile][row][BN]; reduced in fixed order by splitk_reduce_kernel
                    float* pp = partial + (((size_t)ks * (p.num_tiles / p.ksplit) + otile) * 128 + r) * p.BN + c0;
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        *reinterpret_cast<float4*>(pp + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                         __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
                } else if (es.part) {
                    // store + InstanceNorm partial sums of the fp32 values
                    float f[32], sq[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]);
                    if (bias) {     // 16-byte loads (chan0 + c0 is a multiple of 32)
                        const float4* b4 = reinterpret_cast<const float4*>(bias + chan0 + c0);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 b = __ldg(b4 + e);
                            f[4 * e] += b.x; f[4 * e + 1] += b.y; f[4 * e + 2] += b.z; f[4 * e + 3] += b.w;
                        }
                    }
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            float o8[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) o8[e] = f[j + e];
                            store8(row + c0 + j, o8);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        f[e] = valid ? f[e] : 0.f;
                        sq[e] = f[e] * f[e];
                    }
                    const float a1 = transpose_reduce32(f, lane), a2 = transpose_reduce32(sq, lane);
                    const int ch = c0 >> 5;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k == ch) { st1[k] += a1; st2[k] += a2; }
                } else if (valid && p.out_f32) {
                    float* rowf = reinterpret_cast<float*>(dst) + ((((long long)on * p.D + od) * p.H + oh) * p.W + ow) * p.dst_pitch + chan0 + c0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        if (bias) { const float4 b = __ldg(reinterpret_cast<const float4*>(bias + chan0 + c0 + j)); o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w; }
                        *reinterpret_cast<float4*>(rowf + j) = o;
                    }
                } else if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[j + e]);
                        if (bias) {
                            const float4* b4 = reinterpret_cast<const float4*>(bias + chan0 + c0 + j);
                            const float4 b0 = __ldg(b4), b1 = __ldg(b4 + 1);
                            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
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
            acc_phase ^= 1;
        }
        if (es.part) {
            if (n_cur >= 0) { st_write(n_cur, false); n_done = n_cur + 1; }
            for (int nz = n_done; nz < es.N; ++nz) st_write(nz, true);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_PRODUCERS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// split-K second stage: out = bf16( sum_ks partial[ks] + bias (+ out) ), same row -> voxel mapping as the epilogue
__global__ void __launch_bounds__(256) splitk_reduce_kernel(TcConvParams p, const float* __restrict__ partial, const float* __restrict__ bias,
                                                            __nv_bfloat16* __restrict__ dst, int accumulate) {
    pdl_grid_sync();
    const int otiles = p.num_tiles / p.ksplit;
    const long long total = (long long)otiles * 128 * (p.BN / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % (p.BN / 8));
        long long rr = i / (p.BN / 8);
        const int r = (int)(rr % 128);
        int t = (int)(rr / 128);
        const int otile = t;
        const int nb = t % p.nblk; t /= p.nblk;
        const int tw = t % p.nt_w; t /= p.nt_w;
        const int th = t % p.nt_h; t /= p.nt_h;
        const int td = t % p.nt_d; t /= p.nt_d;
        const int tn = t;
        const int w_ = r % p.TW, h_ = (r / p.TW) % p.TH, d_ = (r / (p.TW * p.TH)) % p.TD, n_ = r / (p.TW * p.TH * p.TD);
        const int lw = tw * p.TW + w_, lh = th * p.TH + h_, ld = td * p.TD + d_, on = tn * p.TN + n_;
        int ow = lw * p.os_w + p.oo_w, oh = lh * p.os_h + p.oo_h, od = ld * p.os_d + p.oo_d;
        int chan0 = nb * p.BN;
        if (p.q_scatter) {
            const int col = nb * p.BN + cg * 8, qq = col / p.q_channels;
            ow += qq % p.qk_w; oh += (qq / p.qk_w) % p.qk_h; od += qq / (p.qk_w * p.qk_h);
            chan0 = col - qq * p.q_channels - cg * 8;
        }
        if (!(lw < p.LW && lh < p.LH && ld < p.LD && on < p.N && ow < p.W && oh < p.H && od < p.D)) continue;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = bias ? bias[chan0 + cg * 8 + e] : 0.f;
        for (int ks = 0; ks < p.ksplit; ++ks) {
            const float* pp = partial + (((size_t)ks * otiles + otile) * 128 + r) * p.BN + cg * 8;
            const float4 a = *reinterpret_cast<const float4*>(pp), b = *reinterpret_cast<const float4*>(pp + 4);
            f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w; f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
        }
        if (p.out_f32) {
            float* rowf = reinterpret_cast<float*>(dst) + ((((long long)on * p.D + od) * p.H + oh) * p.W + ow) * p.dst_pitch + chan0 + cg * 8;
            *reinterpret_cast<float4*>(rowf) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(rowf + 4) = make_float4(f[4], f[5], f[6], f[7]);
            continue;
        }
        __nv_bfloat16* row = dst + ((((long long)on * p.D + od) * p.H + oh) * p.W + ow) * p.dst_pitch + chan0 + cg * 8;
        if (accumulate) {
            float o[8];
            load8(row, o);
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] += o[e];
        }
        store8(row, f);
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
int make_w_map(CUtensorMap* m, const void* ptr, int rows, int K, int KC, int BN) {
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

// PyTorch [Cout][Cin][27] fp32 -> Wk [27][Cout][Cin] bf16 (forward) and Wd [27 flipped][Cin][Cout] bf16 (stride-1 dgrad).
// Plain scatter: used only for shapes the tiled multi-tensor kernel below does not cover (channel counts % 32 != 0).
__global__ void weight_shadow_bf16_kernel(const float* __restrict__ w, int Cout, int Cin, __nv_bfloat16* __restrict__ wk,
                                          __nv_bfloat16* __restrict__ wd) {
    pdl_grid_sync();
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

// Multi-tensor tiled shadow kernel: ONE launch converts every weight tensor of a forward pass.  A job is a 3-D fp32 tensor
// w[A][B][T] (conv: A = Cout, B = Cin, T = 27; transposed conv: A = Cin, B = Cout, T = kd*kh*kw) and up to two bf16 outputs
//     oab[t][A][B]            (conv: forward shadow Wk;            tconv: dgrad shadow wqd)
//     oba[flip ? T-1-t : t][B][A]   (conv: flipped dgrad shadow Wd, flip = 1;  tconv: forward shadow wq)
// A CTA owns a 32 (A) x 32 (B) x T tile: coalesced fp32 reads (32 rows of 32*T contiguous floats), bf16 staging in shared
// memory (odd word strides: conflict-free for both read patterns), 16-byte stores of 64-byte row segments on both outputs.
// The per-layer scatter kernels this replaces ran at ~0.5 TB/s (2-byte stores with a 27-element stride): 0.57 ms per step.
constexpr int SH_ROW = 34;                          // bf16 elements per (t, a) row: 17 words
constexpr int SH_TSTRIDE = 32 * SH_ROW + 2;         // 545 words: odd, so consecutive taps hit consecutive banks
constexpr int SH_MAXT = 27;
constexpr int SHADOW_MAX_JOBS = 40;
struct ShadowJobsDev {
    const float* w[SHADOW_MAX_JOBS];
    __nv_bfloat16* oab[SHADOW_MAX_JOBS];
    __nv_bfloat16* oba[SHADOW_MAX_JOBS];
    int A[SHADOW_MAX_JOBS], B[SHADOW_MAX_JOBS];
    int tile0[SHADOW_MAX_JOBS + 1];                  // first tile of each job (prefix sum)
    unsigned char T[SHADOW_MAX_JOBS], flip[SHADOW_MAX_JOBS];
    int n;
};

__global__ void __launch_bounds__(256) shadow_multi_kernel(const __grid_constant__ ShadowJobsDev jobs) {
    pdl_grid_sync();
    extern __shared__ __align__(16) __nv_bfloat16 sh_w[];   // [T][32 a][SH_ROW] with tap stride SH_TSTRIDE
    int j = 0;
    while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.tile0[j + 1]) ++j;
    const int tile = (int)blockIdx.x - jobs.tile0[j];
    const int A = jobs.A[j], B = jobs.B[j], T = jobs.T[j];
    const int tb = B >> 5;
    const int a0 = (tile / tb) << 5, b0 = (tile % tb) << 5;
    const float* __restrict__ w = jobs.w[j];
    // load: 32 rows (a) of 32*T contiguous floats (16-byte aligned: b0 % 32 == 0), as float4 in batches of 4 independent loads
    const int row4 = 8 * T, n4 = 32 * row4;
    if (!(jobs.flip[j] & 2)) {
        // source not 16-byte aligned (a view at an odd offset): scalar loads
        const int rowlen = 32 * T;
        for (int a = 0; a < 32; ++a) {
            const float* src = w + ((long long)(a0 + a) * B + b0) * T;
            for (int r = threadIdx.x; r < rowlen; r += 256) {
                const int b = r / T, t = r - b * T;
                sh_w[t * SH_TSTRIDE + a * SH_ROW + b] = __float2bfloat16_rn(src[r]);
            }
        }
    } else
    for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * 256) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 256 * u;
            if (i < n4) {
                const int a = i / row4, r4 = i - a * row4;
                v[u] = *reinterpret_cast<const float4*>(w + ((long long)(a0 + a) * B + b0) * T + 4 * r4);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 256 * u;
            if (i < n4) {
                const int a = i / row4, r4 = i - a * row4;
                const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = 4 * r4 + q, b = r / T, t = r - b * T;
                    sh_w[t * SH_TSTRIDE + a * SH_ROW + b] = __float2bfloat16_rn(e[q]);
                }
            }
        }
    }
    __syncthreads();
    const int chunks = T * 128;                      // (t, row, 8-element chunk)
    __nv_bfloat16* __restrict__ oab = jobs.oab[j];
    __nv_bfloat16* __restrict__ oba = jobs.oba[j];
    if (oab) {
        for (int i = threadIdx.x; i < chunks; i += 256) {
            const int q = i & 3, a = (i >> 2) & 31, t = i >> 7;
            const uint32_t* s = reinterpret_cast<const uint32_t*>(sh_w + t * SH_TSTRIDE + a * SH_ROW + q * 8);
            uint4 v = make_uint4(s[0], s[1], s[2], s[3]);
            *reinterpret_cast<uint4*>(oab + ((long long)t * A + a0 + a) * B + b0 + q * 8) = v;
        }
    }
    if (oba) {
        const int flip = jobs.flip[j] & 1;
        for (int i = threadIdx.x; i < chunks; i += 256) {
            const int q = i & 3, b = (i >> 2) & 31, t = i >> 7;
            const unsigned short* s = reinterpret_cast<const unsigned short*>(sh_w + t * SH_TSTRIDE + (q * 8) * SH_ROW + b);
            uint32_t r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                r[e] = (uint32_t)s[(2 * e) * SH_ROW] | ((uint32_t)s[(2 * e + 1) * SH_ROW] << 16);
            const int tt = flip ? T - 1 - t : t;
            *reinterpret_cast<uint4*>(oba + ((long long)tt * B + b0 + b) * A + a0 + q * 8) = make_uint4(r[0], r[1], r[2], r[3]);
        }
    }
}

bool shadow_job_supported(int A, int B, int T) { return A % 32 == 0 && B % 32 == 0 && T >= 1 && T <= SH_MAXT; }

int shadow_multi(const ShadowJob* jobs, int n, cudaStream_t st) {
    static bool attr = false;
    const size_t smem = (size_t)SH_MAXT * SH_TSTRIDE * sizeof(__nv_bfloat16);
    if (!attr) { B2_CUDA(cudaFuncSetAttribute(shadow_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    for (int base = 0; base < n; base += SHADOW_MAX_JOBS) {
        ShadowJobsDev d;
        memset(&d, 0, sizeof(d));
        const int m = n - base < SHADOW_MAX_JOBS ? n - base : SHADOW_MAX_JOBS;
        int tiles = 0;
        for (int i = 0; i < m; ++i) {
            const ShadowJob& jb = jobs[base + i];
            B2_CHECK_ARG(shadow_job_supported(jb.A, jb.B, jb.T));
            d.w[i] = jb.w; d.oab[i] = jb.oab; d.oba[i] = jb.oba; d.A[i] = jb.A; d.B[i] = jb.B;
            d.T[i] = (unsigned char)jb.T;
            d.flip[i] = (unsigned char)((jb.flip ? 1 : 0) | (((uintptr_t)jb.w & 15) == 0 ? 2 : 0));   // bit 1: 16-byte aligned source
            d.tile0[i] = tiles;
            tiles += (jb.A / 32) * (jb.B / 32);
        }
        d.tile0[m] = tiles;
        d.n = m;
        if (tiles > 0) B2_LAUNCH(shadow_multi_kernel, tiles, 256, smem, st, d);
    }
    return B2_OK;
}

int weight_shadow_bf16(const float* w, int cout, int cin, __nv_bfloat16* wk, __nv_bfloat16* wd, cudaStream_t st) {
    if (!wk && !wd) return B2_OK;
    if (shadow_job_supported(cout, cin, 27)) {
        ShadowJob jb{w, wk, wd, cout, cin, 27, 1};
        return shadow_multi(&jb, 1, st);
    }
    long long tot = (long long)cout * cin * 27;
    B2_LAUNCH(weight_shadow_bf16_kernel, cdiv(tot, 256), 256, 0, st, w, cout, cin, wk, wd);
    return B2_OK;
}

// Generic launcher of the gather-GEMM.  src: NDHWC bf16 [N, Ds, Hs, Ws, K] (pitch src_pitch); wmat: [row blocks][Nout][K]
// bf16; dst: NDHWC bf16 [N, Dd, Hd, Wd, *] (pitch dst_pitch).  See TcGather for the tap / lattice description.
int conv_tc_gather(const TcGather& g, cudaStream_t st) {
    B2_CHECK_ARG(conv_tc_supported(g.K, g.Nout));
    B2_CHECK_ARG(g.src_pitch % 8 == 0 && g.dst_pitch % 8 == 0 && g.ntaps >= 1 && g.ntaps <= 27);
    const int KC = (g.K % 64 == 0) ? 64 : 32;
    int nblk, BN;
    if (g.mes) {
        BN = g.Nout; nblk = 1;          // all parity classes side by side: Nout = nclass * channels, <= 256
        B2_CHECK_ARG(BN <= 256 && g.q_scatter && g.q_channels == g.mes_blk_rows && g.mes_nst <= 12);
    } else if (g.q_scatter) {
        // 32-column chunks must not straddle q (q_channels % 32 == 0).  Narrow layers: one tile spans several q (BN = 256)
        // -- a tile per q made Cout = 32 layers run as 8x more (K-iteration-free, epilogue-bound) tiles; wide layers: block
        // the per-q channel count
        B2_CHECK_ARG(g.q_channels % 32 == 0);
        if (g.q_channels <= 256 && 256 % g.q_channels == 0) {
            BN = g.Nout < 256 ? g.Nout : 256;
            while (g.Nout % BN != 0) BN -= g.q_channels;
        } else {
            int per = cdiv(g.q_channels, 256);
            while (g.q_channels % per != 0 || (g.q_channels / per) % 32 != 0) ++per;
            BN = g.q_channels / per;
        }
        nblk = g.Nout / BN;
    } else {
        nblk = cdiv(g.Nout, 256);
        while (g.Nout % nblk != 0 || (g.Nout / nblk) % 32 != 0) ++nblk;
        BN = g.Nout / nblk;
    }
    TcConvParams p;
    memset(&p, 0, sizeof(p));
    p.N = g.N; p.D = g.Dd; p.H = g.Hd; p.W = g.Wd; p.dst_pitch = g.dst_pitch;
    p.LD = g.LD; p.LH = g.LH; p.LW = g.LW;
    p.TW = pow2_le(g.LW, 8);
    p.TH = pow2_le(g.LH, 128 / p.TW > 8 ? 8 : 128 / p.TW);
    p.TD = pow2_le(g.LD, 128 / (p.TW * p.TH));
    p.TN = 128 / (p.TW * p.TH * p.TD);
    p.nt_w = cdiv(g.LW, p.TW); p.nt_h = cdiv(g.LH, p.TH); p.nt_d = cdiv(g.LD, p.TD); p.nt_n = cdiv(g.N, p.TN);
    p.nblk = nblk; p.BN = BN; p.kchunks = g.K / KC; p.rows_per_tap = g.rows_per_tap; p.w_row0 = g.w_row0;
    p.sd = g.stride[0]; p.sh = g.stride[1]; p.sw = g.stride[2];
    p.os_d = g.os[0]; p.os_h = g.os[1]; p.os_w = g.os[2];
    p.oo_d = g.oo[0]; p.oo_h = g.oo[1]; p.oo_w = g.oo[2];
    p.q_scatter = g.q_scatter; p.q_channels = g.q_scatter ? g.q_channels : 1; p.qk_h = g.qk[1]; p.qk_w = g.qk[2];
    p.out_f32 = g.out_f32;
    p.fd_qch = make_fastdiv(p.q_channels); p.fd_qkw = make_fastdiv(p.qk_w); p.fd_qkh = make_fastdiv(p.qk_h);
    p.fd_nblk = make_fastdiv(nblk); p.fd_ntw = make_fastdiv(p.nt_w); p.fd_nth = make_fastdiv(p.nt_h); p.fd_ntd = make_fastdiv(p.nt_d);
    p.ntaps = g.ntaps;
    for (int t = 0; t < g.ntaps; ++t) {
        for (int a = 0; a < 3; ++a) p.tap_off[t][a] = (signed char)g.tap_off[t][a];
        p.tap_w[t] = (unsigned char)g.tap_w[t];
    }
    int otiles = p.nt_w * p.nt_h * p.nt_d * p.nt_n * nblk;
    if (g.nclass > 1) {
        // tile classes: same voxel box for all, tiles enumerated class after class
        B2_CHECK_ARG(g.nclass <= 8 && !g.q_scatter);
        p.nclass = g.nclass;
        int acc = 0;
        for (int c = 0; c < g.nclass; ++c) {
            p.cls_tile0[c] = acc;
            p.cls_tap0[c] = (unsigned char)g.cls_tap0[c]; p.cls_ntaps[c] = (unsigned char)g.cls_ntaps[c];
            for (int a = 0; a < 3; ++a) { p.cls_oo[c][a] = (signed char)g.cls_oo[c][a]; p.cls_L[c][a] = (short)g.cls_L[c][a]; }
            p.cls_nt[c][0] = (short)cdiv(g.cls_L[c][0], p.TD); p.cls_nt[c][1] = (short)cdiv(g.cls_L[c][1], p.TH);
            p.cls_nt[c][2] = (short)cdiv(g.cls_L[c][2], p.TW);
            for (int a = 0; a < 3; ++a) p.cls_fd[c][a] = make_fastdiv(p.cls_nt[c][a]);
            acc += p.cls_nt[c][0] * p.cls_nt[c][1] * p.cls_nt[c][2] * p.nt_n * nblk;
        }
        p.cls_tile0[g.nclass] = acc;
        otiles = acc;
    }
    // split-K when the output tiles cannot fill the GPU (deep, small-volume layers)
    int ksplit = 1;
    const int kiters_total = g.ntaps * p.kchunks;
    if (g.splitk_scratch && g.nclass <= 1 && !g.mes && otiles * 2 <= num_sms()) {
        ksplit = num_sms() / otiles;
        if (ksplit > kiters_total / 4) ksplit = kiters_total / 4;
        if (ksplit < 1) ksplit = 1;
        while (ksplit > 1 && (size_t)ksplit * otiles * 128 * BN * sizeof(float) > g.splitk_scratch_bytes) --ksplit;
    }
    p.ksplit = ksplit;
    p.fd_ksplit = make_fastdiv(ksplit);
    p.num_tiles = otiles * ksplit;
    p.idesc = umma_idesc_bf16(128, g.mes ? g.mes_blk_rows : BN);
    if (g.mes) {
        p.mes = 1; p.mes_nst = g.mes_nst; p.mes_blk_rows = g.mes_blk_rows;
        for (int i = 0; i < 12; ++i) p.mes_nb[i] = (unsigned char)g.mes_nb[i];
        for (int i = 0; i < 13; ++i) p.mes_blk0[i] = (unsigned char)g.mes_blk0[i];
        for (int i = 0; i < 27; ++i) {
            p.mes_wrow[i] = (unsigned char)g.mes_wrow[i]; p.mes_cls[i] = (unsigned char)g.mes_cls[i]; p.mes_first[i] = (unsigned char)g.mes_first[i];
        }
    }
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * BN)) cols *= 2;
    p.tmem_cols = cols;
    const uint32_t a_bytes = 128 * KC * 2, b_bytes = ((uint32_t)BN * KC * 2 + 1023) & ~1023u;
    const uint32_t stage_bytes = a_bytes + b_bytes;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return fail(B2_EUNSUPPORTED, "conv_tc: tile does not fit shared memory%s", "");
    // ring depth a multiple of the producer count: every stage is then always filled by the same producer warp, which
    // keeps the 1-bit mbarrier phase parity unambiguous
    if (stages >= TC_PRODUCERS) stages = stages / TC_PRODUCERS * TC_PRODUCERS;
    p.stages = stages;
    p.nprod = stages < TC_PRODUCERS ? stages : TC_PRODUCERS;
    const size_t smem = (size_t)stages * stage_bytes + 1024;

    CUtensorMap tmA, tmB;
    int rc = make_act_map(&tmA, g.src, g.N, g.Ds, g.Hs, g.Ws, g.K, g.src_pitch, KC, p.TN, p.TD, p.TH, p.TW, p.sd, p.sh, p.sw);
    if (rc) return rc;
    rc = make_w_map(&tmB, g.wmat, g.w_rows, g.K, KC, g.mes ? g.mes_blk_rows : BN);
    if (rc) return rc;

    int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    EpiStats es{nullptr, 0, g.Nout, g.N};
    if (g.stat_slots) *g.stat_slots = 0;
    // (few K iterations per tile = epilogue-bound kernel: a separate streaming pass over z is cheaper there)
    if (g.stat_part && g.stat_slots && g_epi_stats && !g.out_f32 && p.TN == 1 && nblk == 1 && ksplit == 1 && !g.q_scatter && g.nclass <= 1 && BN <= 256 &&
        !g.accumulate && kiters_total >= 8) {
        es.slots = grid * 4 * TC_EPI_SETS;
        if ((size_t)g.N * es.slots * g.Nout * 2 <= g.stat_part_floats) { es.part = g.stat_part; *g.stat_slots = es.slots; }
    }
    static bool attr_set[4] = {false, false, false, false};
#define B2_TC_LAUNCH(KC_, MES_, IDX_)                                                                                              \
    do {                                                                                                                         \
        if (!attr_set[IDX_]) {                                                                                                   \
            B2_CUDA(cudaFuncSetAttribute((conv_tc_kernel<KC_, MES_>), cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));   \
            attr_set[IDX_] = true;                                                                                               \
        }                                                                                                                        \
        B2_LAUNCH((conv_tc_kernel<KC_, MES_>), grid, TC_THREADS, smem, st, tmA, tmB, p, g.bias, g.dst, g.accumulate, g.splitk_scratch, es); \
    } while (0)
    if (KC == 64) { if (g.mes) B2_TC_LAUNCH(64, true, 0); else B2_TC_LAUNCH(64, false, 1); }
    else { if (g.mes) B2_TC_LAUNCH(32, true, 2); else B2_TC_LAUNCH(32, false, 3); }
#undef B2_TC_LAUNCH
    if (g.defer) g.defer->deferred = 0;
    if (p.ksplit > 1 && g.defer && !g.accumulate && !g.q_scatter && !g.out_f32 && g.nclass <= 1 && !g.mes && p.os_d == 1 && p.os_h == 1 &&
        p.os_w == 1 && p.oo_d == 0 && p.oo_h == 0 && p.oo_w == 0 && p.LD == p.D && p.LH == p.H && p.LW == p.W && BN % 8 == 0) {
        SplitKDefer& k = *g.defer;       // the consumer (small-tensor norm kernel) sums the partials: no reduce launch here
        k.deferred = 1;
        k.TN = p.TN; k.TD = p.TD; k.TH = p.TH; k.TW = p.TW; k.nt_d = p.nt_d; k.nt_h = p.nt_h; k.nt_w = p.nt_w;
        k.nblk = p.nblk; k.BN = p.BN; k.ksplit = p.ksplit; k.otiles = otiles; k.partial = g.splitk_scratch;
    } else if (p.ksplit > 1) {
        const long long total = (long long)otiles * 128 * (BN / 8);
        long long rg = (total + 255) / 256, cap = (long long)num_sms() * 8;
        if (rg > cap) rg = cap;
        B2_LAUNCH(splitk_reduce_kernel, (int)rg, 256, 0, st, p, g.splitk_scratch, g.bias, g.dst, g.accumulate);
    }
    return B2_OK;
}

static void fill_common(TcGather& g, const __nv_bfloat16* src, int N, int Ds, int Hs, int Ws, int K, int src_pitch,
                        const __nv_bfloat16* wmat, int Nout, const float* bias, __nv_bfloat16* dst, int Dd, int Hd, int Wd,
                        int dst_pitch, int accumulate) {
    memset(&g, 0, sizeof(g));
    g.src = src; g.N = N; g.Ds = Ds; g.Hs = Hs; g.Ws = Ws; g.K = K; g.src_pitch = src_pitch;
    g.wmat = wmat; g.Nout = Nout; g.rows_per_tap = Nout; g.bias = bias;
    g.dst = dst; g.Dd = Dd; g.Hd = Hd; g.Wd = Wd; g.dst_pitch = dst_pitch; g.accumulate = accumulate;
    g.LD = Dd; g.LH = Hd; g.LW = Wd;
    for (int a = 0; a < 3; ++a) { g.stride[a] = 1; g.os[a] = 1; g.oo[a] = 0; g.qk[a] = 1; }
}

// fp32 scratch that lets conv_tc_launch split K for a produced tensor of this size (0 when the tiles already fill the GPU)
size_t conv_tc_splitk_scratch_floats(int N, int D, int H, int W, int Nout) {
    const long long vox = (long long)N * D * H * W;
    const long long tiles = (vox + 127) / 128 + 8;
    int nblk = cdiv(Nout, 256);
    if (tiles * nblk * 2 > num_sms()) return 0;
    return (size_t)num_sms() * 2 * 128 * 256;   // ksplit * otiles <= num_sms, BN <= 256 (x2 slack for ragged boxes)
}

// Plain GEMM on the gather kernel (one tap, unit lattice):  out[m][n] = sum_k A[m][k] * W[n][k] (+ bias[n]),  A: [M][K] bf16 with
// row pitch lda, W: [N][K] bf16 dense, out: bf16 or fp32 with row pitch ldo.  M is tiled in boxes of 128 consecutive rows
// (rows past M are zero-filled by TMA and never stored), so the row block is described as an (M/64 x 8 x 8) lattice when M is
// a multiple of 64 and as a (1 x 1 x M) row otherwise.  K % 32 == 0, N % 32 == 0, lda % 8 == 0, ldo % 8 == 0.
// `scratch` (optional) enables split-K for launches with few output tiles (fp32 partials + ordered reduce: bit-reproducible).
int gemm_tn_bf16(const __nv_bfloat16* A, int M, int K, int lda, const __nv_bfloat16* W, int N, const float* bias, void* out, int ldo,
                 int out_f32, float* scratch, size_t scratch_bytes, cudaStream_t st) {
    B2_CHECK_ARG(A && W && out && M >= 1 && K % 32 == 0 && N % 32 == 0 && lda % 8 == 0 && ldo % 8 == 0 && lda >= K);
    TcGather g;
    const int D = M % 64 == 0 ? M / 64 : 1, H = M % 64 == 0 ? 8 : 1, Wd = M % 64 == 0 ? 8 : M;
    fill_common(g, A, 1, D, H, Wd, K, lda, W, N, bias, (__nv_bfloat16*)out, D, H, Wd, ldo, 0);
    g.w_rows = N; g.ntaps = 1; g.out_f32 = out_f32;
    g.tap_off[0][0] = g.tap_off[0][1] = g.tap_off[0][2] = 0; g.tap_w[0] = 0;
    g.splitk_scratch = scratch; g.splitk_scratch_bytes = scratch_bytes;
    return conv_tc_gather(g, st);
}

// 3x3x3, padding 1, any stride (forward) / stride 1 with the flipped shadow (dgrad)
int conv_tc_launch(const __nv_bfloat16* src, int N, int Ds, int Hs, int Ws, int K, int src_pitch, const __nv_bfloat16* wmat,
                   int Nout, const float* bias, __nv_bfloat16* dst, int Dd, int Hd, int Wd, int dst_pitch, const int stride[3],
                   int accumulate, cudaStream_t st, float* scratch, size_t scratch_bytes, int* stat_slots, int w_pitch, int w_row0,
                   SplitKDefer* defer) {
    // stat_slots != nullptr: the caller wants InstanceNorm partials in `scratch` ([N][*stat_slots][Nout][2], see EpiStats);
    // *stat_slots == 0 on return means the chosen kernel could not produce them (split-K, batch-spanning boxes, ...)
    if (stat_slots) *stat_slots = 0;
    if (defer) defer->deferred = 0;
    if (stride[0] == 1 && stride[1] == 1 && stride[2] == 1 && conv_tc_halo_supported(K, Nout, N, Dd, Hd, Wd))
        return conv_tc_halo_launch(src, N, Dd, Hd, Wd, K, src_pitch, wmat, Nout, bias, dst, dst_pitch, accumulate, st,
                                   stat_slots ? scratch : nullptr, scratch_bytes / sizeof(float), stat_slots, w_pitch, w_row0);
    TcGather g;
    fill_common(g, src, N, Ds, Hs, Ws, K, src_pitch, wmat, Nout, bias, dst, Dd, Hd, Wd, dst_pitch, accumulate);
    for (int a = 0; a < 3; ++a) g.stride[a] = stride[a];
    g.ntaps = 27; g.w_rows = 27 * Nout;
    if (w_pitch > 0) { g.w_rows = 27 * w_pitch; g.rows_per_tap = w_pitch; g.w_row0 = w_row0; }   // output-channel window
    g.splitk_scratch = scratch; g.splitk_scratch_bytes = scratch_bytes;
    g.defer = defer;
    g.stat_part = stat_slots ? scratch : nullptr; g.stat_part_floats = scratch_bytes / sizeof(float); g.stat_slots = stat_slots;
    for (int t = 0; t < 27; ++t) {
        g.tap_off[t][0] = t / 9 - 1; g.tap_off[t][1] = (t / 3) % 3 - 1; g.tap_off[t][2] = t % 3 - 1;
        g.tap_w[t] = t;
    }
    return conv_tc_gather(g, st);
}

// dgrad of a STRIDED 3x3x3 conv: one launch per output-parity class; class p receives the taps t with (p - t + 1) even
// and reads dz at j + (p - t + 1) / 2.  wd = flipped/transposed shadow [26 - t][ci][co] (weight_shadow_bf16).
int g_dgrad_one_launch = 1;
int g_dgrad_mes = 1;        // merged-class strided dgrad (shared-A stages) when all classes fit one accumulator (nclass * Cin <= 256)

// dx[j*s + p] for ALL parity classes p of a coarse-lattice tile j in one accumulator set (columns = (class, ci)).
static int conv_tc_dgrad_strided_mes(const __nv_bfloat16* dz, int N, int Do, int Ho, int Wo, int Cout, int dz_pitch, const __nv_bfloat16* wd,
                                     int Cin, __nv_bfloat16* dx, int Di, int Hi, int Wi, int dx_pitch, const int stride[3], int accumulate,
                                     cudaStream_t st) {
    const int nclass = stride[0] * stride[1] * stride[2];
    TcGather g;
    fill_common(g, dz, N, Do, Ho, Wo, Cout, dz_pitch, wd, nclass * Cin, nullptr, dx, Di, Hi, Wi, dx_pitch, accumulate);
    g.w_rows = 27 * Cin; g.rows_per_tap = Cin;
    // logical grid = coarse lattice of the largest class (class 0): ceil(dim / stride)
    g.LD = (Di + stride[0] - 1) / stride[0]; g.LH = (Hi + stride[1] - 1) / stride[1]; g.LW = (Wi + stride[2] - 1) / stride[2];
    for (int a = 0; a < 3; ++a) { g.os[a] = stride[a]; g.oo[a] = 0; g.qk[a] = stride[a]; }
    g.q_scatter = 1; g.q_channels = Cin;
    g.mes = 1; g.mes_blk_rows = Cin; g.mes_nst = 0;
    // enumerate (class, tap) pairs and group them by source shift
    struct Pair { int cls, t, off[3]; };
    Pair pairs[27];
    int np = 0;
    for (int pd = 0; pd < stride[0]; ++pd)
        for (int ph = 0; ph < stride[1]; ++ph)
            for (int pw = 0; pw < stride[2]; ++pw) {
                const int par[3] = {pd, ph, pw};
                const int cls = (pd * stride[1] + ph) * stride[2] + pw;     // == q of the scatter mapping (w fastest)
                int cnt[3], tt[3][3], off[3][3];
                for (int a = 0; a < 3; ++a) {
                    cnt[a] = 0;
                    for (int t = 0; t < 3; ++t) {
                        const int num = par[a] - t + 1;
                        if (stride[a] == 1) { tt[a][cnt[a]] = t; off[a][cnt[a]] = 1 - t; ++cnt[a]; }
                        else if ((num & 1) == 0) { tt[a][cnt[a]] = t; off[a][cnt[a]] = num / 2; ++cnt[a]; }
                    }
                }
                for (int i = 0; i < cnt[0]; ++i)
                    for (int j = 0; j < cnt[1]; ++j)
                        for (int k = 0; k < cnt[2]; ++k) {
                            if (np >= 27) return fail(B2_EINVAL, "dgrad_mes: more than 27 (class, tap) pairs%s", "");
                            pairs[np].cls = cls; pairs[np].t = tt[0][i] * 9 + tt[1][j] * 3 + tt[2][k];
                            pairs[np].off[0] = off[0][i]; pairs[np].off[1] = off[1][j]; pairs[np].off[2] = off[2][k];
                            ++np;
                        }
            }
    bool used[27] = {false}, seen_cls[64] = {false};
    int nblk_total = 0;
    for (int i = 0; i < np; ++i) {
        if (used[i]) continue;
        if (g.mes_nst >= 12) return fail(B2_EUNSUPPORTED, "dgrad_mes: more than 12 source shifts%s", "");
        const int s_ = g.mes_nst++;
        for (int a = 0; a < 3; ++a) g.tap_off[s_][a] = pairs[i].off[a];
        g.mes_blk0[s_] = nblk_total;
        int nb = 0;
        for (int j = i; j < np; ++j) {     // pairs are enumerated class-major: blocks of a stage come out class-ascending
            if (used[j] || pairs[j].off[0] != pairs[i].off[0] || pairs[j].off[1] != pairs[i].off[1] || pairs[j].off[2] != pairs[i].off[2]) continue;
            used[j] = true;
            g.mes_wrow[nblk_total] = 26 - pairs[j].t;
            g.mes_cls[nblk_total] = pairs[j].cls;
            g.mes_first[nblk_total] = seen_cls[pairs[j].cls] ? 0 : 1;
            seen_cls[pairs[j].cls] = true;
            ++nblk_total; ++nb;
        }
        g.mes_nb[s_] = nb;
        if (nb > nclass) return fail(B2_EINVAL, "dgrad_mes: stage with more blocks than classes%s", "");
    }
    g.mes_blk0[g.mes_nst] = nblk_total;
    g.ntaps = g.mes_nst;
    for (int t = 0; t < g.ntaps; ++t) g.tap_w[t] = 0;
    return conv_tc_gather(g, st);
}

int conv_tc_dgrad_strided(const __nv_bfloat16* dz, int N, int Do, int Ho, int Wo, int Cout, int dz_pitch, const __nv_bfloat16* wd,
                          int Cin, __nv_bfloat16* dx, int Di, int Hi, int Wi, int dx_pitch, const int stride[3], int accumulate,
                          cudaStream_t st) {
    const int dims_in[3] = {Di, Hi, Wi};
    if (g_dgrad_mes && g_dgrad_one_launch && Cin % 32 == 0 && stride[0] * stride[1] * stride[2] * Cin <= 256)
        return conv_tc_dgrad_strided_mes(dz, N, Do, Ho, Wo, Cout, dz_pitch, wd, Cin, dx, Di, Hi, Wi, dx_pitch, stride, accumulate, st);
    // all parity classes in ONE launch (tile classes of conv_tc_kernel): the per-class launches of the deep layers were
    // dominated by launch + pipeline fill (8 x 11..29 us for 0.3..4 GFLOP each)
    TcGather m;
    fill_common(m, dz, N, Do, Ho, Wo, Cout, dz_pitch, wd, Cin, nullptr, dx, Di, Hi, Wi, dx_pitch, accumulate);
    m.w_rows = 27 * Cin;
    m.ntaps = 0; m.nclass = 0;
    for (int a = 0; a < 3; ++a) m.os[a] = stride[a];
    for (int pd = 0; pd < stride[0]; ++pd)
        for (int ph = 0; ph < stride[1]; ++ph)
            for (int pw = 0; pw < stride[2]; ++pw) {
                const int par[3] = {pd, ph, pw};
                TcGather g;
                fill_common(g, dz, N, Do, Ho, Wo, Cout, dz_pitch, wd, Cin, nullptr, dx, Di, Hi, Wi, dx_pitch, accumulate);
                g.w_rows = 27 * Cin;
                int L[3];
                for (int a = 0; a < 3; ++a) {
                    g.os[a] = stride[a]; g.oo[a] = par[a];
                    L[a] = stride[a] == 1 ? dims_in[a] : (dims_in[a] - par[a] + 1) / 2;
                }
                g.LD = L[0]; g.LH = L[1]; g.LW = L[2];
                if (L[0] <= 0 || L[1] <= 0 || L[2] <= 0) continue;
                // taps per axis: (tap index, source offset)
                int cnt[3], tt[3][3], off[3][3];
                for (int a = 0; a < 3; ++a) {
                    cnt[a] = 0;
                    for (int t = 0; t < 3; ++t) {
                        const int num = par[a] - t + 1;
                        if (stride[a] == 1) { tt[a][cnt[a]] = t; off[a][cnt[a]] = 1 - t; ++cnt[a]; }
                        else if ((num & 1) == 0) { tt[a][cnt[a]] = t; off[a][cnt[a]] = num / 2; ++cnt[a]; }
                    }
                }
                g.ntaps = 0;
                const int c = m.nclass;
                m.cls_tap0[c] = m.ntaps;
                for (int i = 0; i < cnt[0]; ++i)
                    for (int j = 0; j < cnt[1]; ++j)
                        for (int k = 0; k < cnt[2]; ++k) {
                            const int t = tt[0][i] * 9 + tt[1][j] * 3 + tt[2][k];
                            g.tap_off[g.ntaps][0] = off[0][i]; g.tap_off[g.ntaps][1] = off[1][j]; g.tap_off[g.ntaps][2] = off[2][k];
                            g.tap_w[g.ntaps] = 26 - t;
                            ++g.ntaps;
                            if (m.ntaps < 27) {
                                for (int a = 0; a < 3; ++a) m.tap_off[m.ntaps][a] = g.tap_off[g.ntaps - 1][a];
                                m.tap_w[m.ntaps] = 26 - t;
                                ++m.ntaps;
                            }
                        }
                m.cls_ntaps[c] = g.ntaps;
                for (int a = 0; a < 3; ++a) { m.cls_oo[c][a] = par[a]; m.cls_L[c][a] = L[a]; }
                ++m.nclass;
                if (!g_dgrad_one_launch) {
                    int rc = conv_tc_gather(g, st);
                    if (rc) return rc;
                }
            }
    if (!g_dgrad_one_launch || m.nclass == 0) return B2_OK;
    // the voxel box is sized for the smallest class lattice (the class extents differ by at most one voxel per axis)
    m.LD = m.cls_L[m.nclass - 1][0]; m.LH = m.cls_L[m.nclass - 1][1]; m.LW = m.cls_L[m.nclass - 1][2];
    for (int c = 0; c < m.nclass; ++c)
        for (int a = 0; a < 3; ++a) {
            int& L = a == 0 ? m.LD : a == 1 ? m.LH : m.LW;
            if (m.cls_L[c][a] < L) L = m.cls_L[c][a];
        }
    return conv_tc_gather(m, st);
}

// transposed conv (kernel == stride) forward: out[(2v + q), co] = sum_ci x[v, ci] * W[ci][co][q]; wq: [(q, co)][ci] bf16
int tconv_tc_fwd(const __nv_bfloat16* x, int N, int D, int H, int W, int Cin, int x_pitch, const __nv_bfloat16* wq, int Cout,
                 const int k[3], __nv_bfloat16* y, int y_pitch, cudaStream_t st) {
    const int k8 = k[0] * k[1] * k[2];
    TcGather g;
    fill_common(g, x, N, D, H, W, Cin, x_pitch, wq, k8 * Cout, nullptr, y, D * k[0], H * k[1], W * k[2], y_pitch, 0);
    g.LD = D; g.LH = H; g.LW = W;
    g.ntaps = 1; g.w_rows = k8 * Cout; g.rows_per_tap = k8 * Cout;
    g.tap_off[0][0] = g.tap_off[0][1] = g.tap_off[0][2] = 0; g.tap_w[0] = 0;
    for (int a = 0; a < 3; ++a) { g.os[a] = k[a]; g.qk[a] = k[a]; }
    g.q_scatter = 1; g.q_channels = Cout;
    return conv_tc_gather(g, st);
}

// transposed conv dgrad: dx[v, ci] = sum_q sum_co dy[(2v + q), co] * W[ci][co][q]; wqd: [q][ci][co] bf16
int tconv_tc_dgrad(const __nv_bfloat16* dy, int N, int D, int H, int W, int Cout, int dy_pitch, const __nv_bfloat16* wqd, int Cin,
                   const int k[3], __nv_bfloat16* dx, int dx_pitch, cudaStream_t st) {
    const int k8 = k[0] * k[1] * k[2];
    TcGather g;
    fill_common(g, dy, N, D * k[0], H * k[1], W * k[2], Cout, dy_pitch, wqd, Cin, nullptr, dx, D, H, W, dx_pitch, 0);
    for (int a = 0; a < 3; ++a) g.stride[a] = k[a];
    g.ntaps = k8; g.w_rows = k8 * Cin;
    for (int q = 0; q < k8; ++q) {
        g.tap_off[q][2] = q % k[2]; g.tap_off[q][1] = (q / k[2]) % k[1]; g.tap_off[q][0] = q / (k[2] * k[1]);
        g.tap_w[q] = q;
    }
    return conv_tc_gather(g, st);
}

// PyTorch ConvTranspose3d weight [Cin][Cout][K8] fp32 -> wq [(q, co)][ci] and wqd [q][ci][co] (bf16)
__global__ void tconv_shadow_bf16_kernel(const float* __restrict__ w, int Cin, int Cout, int K8, __nv_bfloat16* __restrict__ wq,
                                         __nv_bfloat16* __restrict__ wqd) {
    pdl_grid_sync();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long tot = (long long)Cin * Cout * K8;
    if (i >= tot) return;
    int q = (int)(i % K8);
    long long r = i / K8;
    int co = (int)(r % Cout), ci = (int)(r / Cout);
    __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
    if (wq) wq[((long long)q * Cout + co) * Cin + ci] = v;
    if (wqd) wqd[((long long)q * Cin + ci) * Cout + co] = v;
}
int tconv_shadow_bf16(const float* w_pt, int cin, int cout, int k8, __nv_bfloat16* wq, __nv_bfloat16* wqd, cudaStream_t st) {
    if (shadow_job_supported(cin, cout, k8)) {
        ShadowJob jb{w_pt, wqd, wq, cin, cout, k8, 0};
        return shadow_multi(&jb, 1, st);
    }
    long long tot = (long long)cin * cout * k8;
    B2_LAUNCH(tconv_shadow_bf16_kernel, cdiv(tot, 256), 256, 0, st, w_pt, cin, cout, k8, wq, wqd);
    return B2_OK;
}

}  // namespace b2
