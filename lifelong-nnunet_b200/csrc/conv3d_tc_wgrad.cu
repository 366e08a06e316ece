// conv3d_tc_wgrad.cu -- weight gradient of the 3x3x3 convolution on the tensor cores (bf16 operands, fp32 accumulate).
//
//   dW[t][ci][co] = sum_v x[v*s + t - 1][ci] * dz[v][co]
//
// GEMM view: D[m, n] with m = (tap, ci) stacked to 128 rows, n = co (block of <= 256), K = voxels.  Both operands are
// channel-contiguous NDHWC boxes (the same TMA boxes the forward uses), i.e. MN-major UMMA operands: a box of 128 voxels
// x 64 channels lands in shared memory as 128 rows of 128 B and is consumed as a [64 (MN) x 128 (K)] tile, 16 voxels per
// tcgen05.mma.  One accumulator group (128 x co_blk fp32 in TMEM) per 128 stacked rows; a CTA keeps up to 512 TMEM
// columns of groups resident over its whole range of voxel tiles (split-K over voxels across CTAs), then writes one fp32
// partial; the ordered reduction over CTAs (wgrad_reduce_kernel) makes the result bit-reproducible -- no atomics.
//
// Stride-1 layers with Cin <= 128, Cout <= 64 use the d-merged formulation (TcWgradParams::dmerge); the thin layers
// (Cin, Cout in {32, 64}) use wgrad_halo_kernel further down, which additionally fetches every input voxel once.
// A layer that needs a single split writes dW directly from the epilogue (no partial tensor, no reduction launch).
//
// Warp roles: warps 0..3 shifted-operand TMA producers, 4 dz producer, 5 MMA issuer + TMEM owner, 6..9 epilogue.
#include <string.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace b2 {

struct TcWgradParams {
    int N, Do, Ho, Wo;           // dz extent
    int TN, TD, TH, TW;          // voxel box (128 voxels)
    int nt_n, nt_d, nt_h, nt_w;
    int sd, sh, sw;
    int Cin, Cout;
    int ci_sub;                  // channels per A chunk (32 -> SWIZZLE_64B, 64 -> SWIZZLE_128B)
    int a_chunks;                // chunks per group (128 / ci_sub)
    int stack_taps;              // 1: chunks of a group are consecutive taps (Cin <= 64); 0: consecutive ci sub-blocks
    int ci_items;                // work items along ci
    int co_blk, co_blks;         // output-channel block and count
    int co_sub, b_chunks;        // channels per B chunk (32/64), chunks per dz tile
    int groups;                  // accumulator groups per CTA (groups * co_blk <= 512)
    int tapsets;                 // work items along taps
    int taps_per_group;          // stack_taps ? a_chunks : 1
    int nsplit;                  // CTAs per work item (split over voxel tiles)
    int num_vtiles;
    int a_stages;
    int nprod;                   // active shifted-operand producer warps (<= a_stages)
    uint32_t idesc;
    uint32_t tmem_cols;
    // MN-major descriptor strides (bytes)
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    int ntaps;
    signed char tap_off[27][3];   // offset of the shifted operand per tap (padding included)
    // d-merged mode (stride-1 3x3x3, Cin <= 64): the shifted operand only carries the 9 (kh, kw) taps and the fixed operand
    // is THREE dz boxes shifted by +1, 0, -1 along d, side by side in N:  D[(kh,kw,ci), (kd,co)] = sum_u x[u+(0,kh-1,kw-1)][ci]
    // * dz[u-(kd-1,0,0)][co] == dW[(kd,kh,kw)][ci][co].  N grows from co_blk to 3*co_blk (an SS-mode MMA costs
    // max(~45, N/2) cycles, so N = 32 wastes 2/3 of the issue slots) and the TMA bytes per voxel tile drop from
    // 27 x-boxes + 1 dz-box to 9 + 3 -- the kernel is bound by the ~43 B/clk/SM L2->SM throughput.
    int dmerge;
    int nper;                     // accumulator columns per group (co_blk, or 3 * co_blk when d-merged)
    int out_taps;                 // taps of the partial tensor written by the epilogue (27 when d-merged, else ntaps)
};

// MN-major UMMA descriptor: `row_bytes` = bytes of one K row (one voxel's channel chunk: 64 or 128)
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t saddr, uint32_t row_bytes, uint32_t lbo, uint32_t sbo) {
    const uint64_t layout = row_bytes == 128 ? 2 : 4;
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

constexpr int WG_PRODUCERS = 4;                        // A-operand TMA producer warps (round robin over group stages)
constexpr int WG_THREADS = 32 * (WG_PRODUCERS + 6);    // + dz producer warp, MMA warp, 4 epilogue warps
constexpr int WG_MAX_ASTAGES = 6;

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmZ, TcWgradParams p,
                float* __restrict__ part, float* __restrict__ dw_direct) {
    pdl_grid_sync();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t fullA[WG_MAX_ASTAGES], emptyA[WG_MAX_ASTAGES], fullB[2], emptyB[2], done_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t a_chunk_bytes = 128u * p.ci_sub * 2;
    const uint32_t A_BYTES = a_chunk_bytes * p.a_chunks;          // 32 KB
    const uint32_t b_chunk_bytes = 128u * p.co_sub * 2;
    const int nb_boxes = p.dmerge ? 3 * p.b_chunks : p.b_chunks;
    const uint32_t B_BYTES = b_chunk_bytes * nb_boxes;
    uint8_t* smemB = smem + (size_t)p.a_stages * A_BYTES;

    // work item decode: blockIdx.x = ((item * nsplit) + split); item = (ci_item, co_blk, tapset)
    const int split = blockIdx.x % p.nsplit;
    int item = blockIdx.x / p.nsplit;
    const int tapset = item % p.tapsets; item /= p.tapsets;
    const int cob = item % p.co_blks; item /= p.co_blks;
    const int ci_item = item;
    const int tap0 = tapset * p.groups * p.taps_per_group;      // first tap of this CTA
    int my_groups = p.groups;                                    // groups that contain at least one valid tap
    {
        const int remaining = p.ntaps - tap0;
        const int need = (remaining + p.taps_per_group - 1) / p.taps_per_group;
        if (need < my_groups) my_groups = need;
    }

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmZ);
        for (int s = 0; s < p.a_stages; ++s) { mbar_init(&fullA[s], 1); mbar_init(&emptyA[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&fullB[s], 1); mbar_init(&emptyB[s], 1); }
        mbar_init(&done_bar, 1);
        fence_barrier_init();
    }
    if (warp == WG_PRODUCERS + 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < WG_PRODUCERS) {
        // ---- shifted-operand producers: warp w issues the group stages with gg % WG_PRODUCERS == w ----
        if (lane == 0 && warp < p.nprod) {
            int gmod = 0, as = warp;
            uint32_t aph = 0;
            for (int vt = split; vt < p.num_vtiles; vt += p.nsplit) {
                int t = vt;
                const int tw = t % p.nt_w; t /= p.nt_w;
                const int th = t % p.nt_h; t /= p.nt_h;
                const int td = t % p.nt_d; t /= p.nt_d;
                const int tn = t;
                const int w0 = tw * p.TW, h0 = th * p.TH, d0 = td * p.TD, n0 = tn * p.TN;
                for (int g = 0; g < my_groups; ++g) {
                    if (gmod == warp) {
                        mbar_wait(&emptyA[as], aph ^ 1);
                        mbar_expect_tx(&fullA[as], A_BYTES);
                        for (int c = 0; c < p.a_chunks; ++c) {
                            int tap, cch;
                            if (p.stack_taps) { tap = tap0 + g * p.a_chunks + c; cch = 0; }
                            else { tap = tap0 + g; cch = ci_item * 128 + c * p.ci_sub; }
                            if (tap > p.ntaps - 1) tap = p.ntaps - 1;  // rows of non-existent taps are ignored by the epilogue
                            tma_load_5d(&tmX, &fullA[as], smem + (size_t)as * A_BYTES + (size_t)c * a_chunk_bytes, cch,
                                        w0 * p.sw + p.tap_off[tap][2], h0 * p.sh + p.tap_off[tap][1], d0 * p.sd + p.tap_off[tap][0], n0);
                        }
                        as += p.nprod;
                        if (as >= p.a_stages) { as -= p.a_stages; aph ^= 1; }
                    }
                    if (++gmod == p.nprod) gmod = 0;
                }
            }
        }
    } else if (warp == WG_PRODUCERS) {
        // ---- fixed-operand (dz) producer ----
        if (lane == 0) {
            int bs = 0;
            uint32_t bph = 0;
            for (int vt = split; vt < p.num_vtiles; vt += p.nsplit) {
                int t = vt;
                const int tw = t % p.nt_w; t /= p.nt_w;
                const int th = t % p.nt_h; t /= p.nt_h;
                const int td = t % p.nt_d; t /= p.nt_d;
                const int tn = t;
                mbar_wait(&emptyB[bs], bph ^ 1);
                mbar_expect_tx(&fullB[bs], B_BYTES);
                for (int c = 0; c < nb_boxes; ++c) {
                    const int kd = c / p.b_chunks, cc = c - kd * p.b_chunks;
                    tma_load_5d(&tmZ, &fullB[bs], smemB + (size_t)bs * B_BYTES + (size_t)c * b_chunk_bytes,
                                cob * p.co_blk + cc * p.co_sub, tw * p.TW, th * p.TH, td * p.TD + (p.dmerge ? 1 - kd : 0), tn * p.TN);
                }
                if (++bs == 2) { bs = 0; bph ^= 1; }
            }
        }
    } else if (warp == WG_PRODUCERS + 1) {
        {
            // whole warp runs the control flow, one elected lane issues; counters instead of divisions, descriptors from precomputed halves (see conv3d_tc_halo.cu)
            const uint32_t a_row = p.ci_sub * 2, b_row = p.co_sub * 2;
            const uint64_t a_hi = umma_desc_mnmajor(0, a_row, p.a_lbo, p.a_sbo) & 0xFFFFFFFFFFFF0000ull;   // everything but the address
            const uint64_t b_hi = umma_desc_mnmajor(0, b_row, p.b_lbo, p.b_sbo) & 0xFFFFFFFFFFFF0000ull;
            const uint32_t a_base = (smem_u32(smem) & 0x3FFFF) >> 4, b_base = (smem_u32(smemB) & 0x3FFFF) >> 4;
            const uint32_t a_stage16 = A_BYTES >> 4, b_stage16 = B_BYTES >> 4, a_k16 = a_row, b_k16 = b_row;   // 16 rows * row_bytes / 16
            int as = 0, bs = 0;
            uint32_t aph = 0, bph = 0, a_lo = a_base;
            bool first = true;
            for (int vt = split; vt < p.num_vtiles; vt += p.nsplit) {
                mbar_wait(&fullB[bs], bph);
                const uint32_t b_lo = b_base + (uint32_t)bs * b_stage16;
                uint32_t d_tmem = tmem_base;
                for (int g = 0; g < my_groups; ++g) {
                    mbar_wait(&fullA[as], aph);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)   // 128 voxels = 8 x K16
                            umma_bf16(d_tmem, a_hi | (uint64_t)(a_lo + k * a_k16), b_hi | (uint64_t)(b_lo + k * b_k16), p.idesc,
                                      (!first || k != 0) ? 1u : 0u);
                        umma_commit(&emptyA[as]);
                        if (g == my_groups - 1) umma_commit(&emptyB[bs]);
                    }
                    __syncwarp();
                    d_tmem += (uint32_t)p.nper;
                    a_lo += a_stage16;
                    if (++as == p.a_stages) { as = 0; aph ^= 1; a_lo = a_base; }
                }
                first = false;
                if (++bs == 2) { bs = 0; bph ^= 1; }
            }
            if (elect_one()) umma_commit(&done_bar);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int chunk = r / p.ci_sub, cil = r % p.ci_sub;
        const bool has_work = split < p.num_vtiles;   // a CTA without voxel tiles writes zeros
        if (has_work) {
            mbar_wait(&done_bar, 0);
            tc_fence_after();
        }
        float* out = part + (size_t)split * p.out_taps * p.Cin * p.Cout;
        for (int g = 0; g < my_groups; ++g) {
            int tap, ci;
            if (p.stack_taps) { tap = tap0 + g * p.a_chunks + chunk; ci = cil; }
            else { tap = tap0 + g; ci = ci_item * 128 + chunk * p.ci_sub + cil; }
            const bool valid = tap < p.ntaps && ci < p.Cin;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * p.nper);
            for (int c0 = 0; c0 < p.nper; c0 += 32) {
                uint32_t v[32];
                if (has_work) {
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = 0u;
                }
                if (valid && dw_direct) {
                    // single split (deep layers: a handful of voxel tiles): this CTA holds the complete sums of its rows, so
                    // it writes dW in the PyTorch layout [co][ci][taps] itself -- no partial tensor, no reduction launch
                    const int kd = c0 / p.co_blk, cc0 = c0 - kd * p.co_blk;
                    const int tap_out = p.dmerge ? kd * 9 + tap : tap;
                    float* dstp = dw_direct + ((size_t)(cob * p.co_blk + cc0) * p.Cin + ci) * p.out_taps + tap_out;
                    const size_t cstride = (size_t)p.Cin * p.out_taps;
#pragma unroll
                    for (int e = 0; e < 32; ++e) dstp[e * cstride] = __uint_as_float(v[e]);
                } else if (valid) {
                    const int kd = c0 / p.co_blk, cc0 = c0 - kd * p.co_blk;    // d-merged: column block -> kd
                    const int tap_out = p.dmerge ? kd * 9 + tap : tap;
                    float* dstp = out + ((size_t)tap_out * p.Cin + ci) * p.Cout + cob * p.co_blk + cc0;
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        *reinterpret_cast<float4*>(dstp + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                           __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WG_PRODUCERS + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// dbias[co] = sum over all voxels of dz: ordered two-stage column sum
__global__ void __launch_bounds__(256) colsum_part_kernel(const __nv_bfloat16* __restrict__ dz, long long rows, int c, int pitch,
                                                          int slabs, float* __restrict__ part) {
    pdl_grid_sync();
    // thread = (column, lane); lanes stride rows
    const int lanes = 256 / c > 0 ? 256 / c : 1;
    const int col = threadIdx.x % c, ln = threadIdx.x / c;
    extern __shared__ float sh[];
    const long long per = (rows + slabs - 1) / slabs;
    const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
    float s = 0.f;
    if (ln < lanes)
        for (long long r = r0 + ln; r < r1; r += lanes) s += __bfloat162float(dz[r * pitch + col]);
    if (ln < lanes) sh[ln * c + col] = s;
    __syncthreads();
    if (threadIdx.x < c) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += sh[l * c + threadIdx.x];
        part[(long long)blockIdx.x * c + threadIdx.x] = t;
    }
}
// wide layers (c > 256): thread per column, rows sequential (their volumes are tiny)
__global__ void __launch_bounds__(256) colsum_wide_kernel(const __nv_bfloat16* __restrict__ dz, long long rows, int c, int pitch,
                                                          int slabs, float* __restrict__ part) {
    pdl_grid_sync();
    const long long per = (rows + slabs - 1) / slabs;
    const long long r0 = (long long)blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
    for (int col = threadIdx.x; col < c; col += 256) {
        float s = 0.f;
        for (long long r = r0; r < r1; ++r) s += __bfloat162float(dz[r * pitch + col]);
        part[(long long)blockIdx.x * c + col] = s;
    }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int slabs, int c, float* __restrict__ out) {
    pdl_grid_sync();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    float s = 0.f;
    for (int k = 0; k < slabs; ++k) s += part[(long long)k * c + i];
    out[i] = s;
}

// ---------------------------------------------------------------------------------------------------------------
int g_wgrad_desc_mode = 0;   // probe switch for the MN-major descriptor stride convention (tests/tools only)

static int pow2_le2(int v, int cap) {
    int p = 1;
    while (p * 2 <= v && p * 2 <= cap) p *= 2;
    return p;
}

bool wgrad_tc_supported(int cin, int cout) {
    if (cin % 32 != 0 || cout % 32 != 0) return false;
    if (cin > 64 && cin % 64 != 0) return false;
    if (cout > 32 && cout % 64 != 0) return false;
    return true;
}

int g_wgrad_dmerge = 1;
int g_wgrad_direct = 1;   // single-split layers write dW directly from the wgrad epilogue

static void wgrad_tc_plan(const ConvShape& s, TcWgradParams& p, int ntaps = 27) {
    memset(&p, 0, sizeof(p));
    const bool dmerge = g_wgrad_dmerge && ntaps == 27 && s.cin <= 128 && s.cout <= 64 && s.stride[0] == 1 && s.stride[1] == 1 &&
                        s.stride[2] == 1;
    p.dmerge = dmerge ? 1 : 0;
    p.out_taps = ntaps;
    if (dmerge) ntaps = 9;
    p.ntaps = ntaps;
    for (int t = 0; t < 27; ++t) { p.tap_off[t][0] = t / 9 - 1; p.tap_off[t][1] = (t / 3) % 3 - 1; p.tap_off[t][2] = t % 3 - 1; }
    if (dmerge) for (int t = 0; t < 9; ++t) { p.tap_off[t][0] = 0; p.tap_off[t][1] = t / 3 - 1; p.tap_off[t][2] = t % 3 - 1; }
    p.N = s.n;
    p.Do = (s.d - 1) / s.stride[0] + 1; p.Ho = (s.h - 1) / s.stride[1] + 1; p.Wo = (s.w - 1) / s.stride[2] + 1;
    p.sd = s.stride[0]; p.sh = s.stride[1]; p.sw = s.stride[2];
    p.TW = pow2_le2(p.Wo, 8);
    p.TH = pow2_le2(p.Ho, 128 / p.TW > 8 ? 8 : 128 / p.TW);
    p.TD = pow2_le2(p.Do, 128 / (p.TW * p.TH));
    p.TN = 128 / (p.TW * p.TH * p.TD);
    p.nt_w = cdiv(p.Wo, p.TW); p.nt_h = cdiv(p.Ho, p.TH); p.nt_d = cdiv(p.Do, p.TD); p.nt_n = cdiv(s.n, p.TN);
    p.num_vtiles = p.nt_w * p.nt_h * p.nt_d * p.nt_n;
    p.Cin = s.cin; p.Cout = s.cout;
    if (s.cin <= 64) { p.ci_sub = s.cin; p.stack_taps = 1; p.ci_items = 1; }
    else { p.ci_sub = 64; p.stack_taps = 0; p.ci_items = cdiv(s.cin, 128); }
    p.a_chunks = 128 / p.ci_sub;
    p.taps_per_group = p.stack_taps ? p.a_chunks : 1;
    p.co_blks = cdiv(s.cout, 256);
    while (s.cout % p.co_blks != 0 || (s.cout / p.co_blks) % 32 != 0) ++p.co_blks;
    p.co_blk = s.cout / p.co_blks;
    p.co_sub = p.co_blk % 64 == 0 ? 64 : 32;
    p.b_chunks = p.co_blk / p.co_sub;
    p.nper = dmerge ? 3 * p.co_blk : p.co_blk;
    const int total_groups = cdiv(ntaps, p.taps_per_group);
    int gmax = 512 / p.nper;
    if (gmax > total_groups) gmax = total_groups;
    p.groups = gmax;
    p.tapsets = cdiv(total_groups, gmax);
    const int items = p.ci_items * p.co_blks * p.tapsets;
    // one resident CTA per SM (200 KB of shared memory): more CTAs than SMs only add partials for the ordered reduction
    int ns = num_sms() / items;
    if (ns < 1) ns = 1;
    if (ns > p.num_vtiles) ns = p.num_vtiles;
    p.nsplit = ns;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.groups * p.nper)) cols *= 2;
    p.tmem_cols = cols;
    // kind::f16, bf16 x bf16 -> f32, A and B MN-major (bits 15, 16), M = 128, N = nper
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.nper >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_row = p.ci_sub * 2, b_row = p.co_sub * 2;
    // canonical MN-major layout (cute::UMMA, see tc_common.cuh): LBO = distance between MN chunks (one TMA box each),
    // SBO = distance between 8-row K groups
    p.a_lbo = 128u * a_row; p.a_sbo = 8u * a_row;
    p.b_lbo = 128u * b_row; p.b_sbo = 8u * b_row;
    if (g_wgrad_desc_mode == 1) { uint32_t t = p.a_lbo; p.a_lbo = p.a_sbo; p.a_sbo = t; t = p.b_lbo; p.b_lbo = p.b_sbo; p.b_sbo = t; }
    const uint32_t A_BYTES = 128u * p.ci_sub * 2 * p.a_chunks, B_BYTES = 128u * p.nper * 2;
    int st = (int)((200u * 1024 - 2 * B_BYTES) / A_BYTES);
    if (st > WG_MAX_ASTAGES) st = WG_MAX_ASTAGES;
    if (st < 2) st = 2;
    if (st >= WG_PRODUCERS) st = st / WG_PRODUCERS * WG_PRODUCERS;   // see conv3d_tc.cu: ring depth % producers == 0
    p.a_stages = st;
    p.nprod = st < WG_PRODUCERS ? st : WG_PRODUCERS;
}

size_t wgrad_tc_part_floats(const ConvShape& s) {
    TcWgradParams p;
    wgrad_tc_plan(s, p);
    size_t colsum = (size_t)(2 * num_sms()) * s.cout;
    int nsplit = p.nsplit;
    if (nsplit < num_sms() && (s.cin == 32 || s.cin == 64) && (s.cout == 32 || s.cout == 64)) nsplit = num_sms();   // halo-reuse plan
    return (size_t)nsplit * 27 * s.cin * s.cout + colsum + 64;
}

// wgrad_reduce_kernel lives in conv3d_simt.cu
int wgrad_reduce(const float* part_w, const float* part_b, int nsplit, int cin, int cout, float* dw, float* db, cudaStream_t st);

static int wgrad_bias(const ConvShape& s, const __nv_bfloat16* dz, float* part_b, float* dbias, bool bias_feeds_norm, cudaStream_t st);
static bool wgrad_halo_plan(const ConvShape& s, struct WgHaloParams& p, size_t& smem);
static int conv3d_wgrad_halo(const ConvShape& s, const __nv_bfloat16* x, const __nv_bfloat16* dz, float* part, float* dw, float* dbias,
                             bool bias_feeds_norm, cudaStream_t st);
static bool wgrad_halo_ok(const ConvShape& s);

int conv3d_wgrad_tc(const ConvShape& s, const __nv_bfloat16* x, const __nv_bfloat16* dz, float* part, float* dw, float* dbias,
                    bool bias_feeds_norm, cudaStream_t st) {
    B2_CHECK_ARG(wgrad_tc_supported(s.cin, s.cout) && s.in_pitch % 8 == 0 && s.out_pitch % 8 == 0);
    if (wgrad_halo_ok(s)) return conv3d_wgrad_halo(s, x, dz, part, dw, dbias, bias_feeds_norm, st);
    TcWgradParams p;
    wgrad_tc_plan(s, p);
    CUtensorMap tmX, tmZ;
    int rc = make_act_map(&tmX, x, s.n, s.d, s.h, s.w, s.cin, s.in_pitch, p.ci_sub, p.TN, p.TD, p.TH, p.TW, p.sd, p.sh, p.sw);
    if (rc) return rc;
    rc = make_act_map(&tmZ, dz, s.n, p.Do, p.Ho, p.Wo, s.cout, s.out_pitch, p.co_sub, p.TN, p.TD, p.TH, p.TW, 1, 1, 1);
    if (rc) return rc;
    const uint32_t A_BYTES = 128u * p.ci_sub * 2 * p.a_chunks, B_BYTES = 128u * p.nper * 2;
    const size_t smem = (size_t)p.a_stages * A_BYTES + 2 * (size_t)B_BYTES + 1024;
    if (smem > 220 * 1024) return fail(B2_EUNSUPPORTED, "wgrad_tc: tile does not fit shared memory%s", "");
    static bool attr = false;
    if (!attr) { B2_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    const int grid = p.ci_items * p.co_blks * p.tapsets * p.nsplit;
    const bool direct = g_wgrad_direct && p.nsplit == 1;
    B2_LAUNCH(wgrad_tc_kernel, grid, WG_THREADS, smem, st, tmX, tmZ, p, part, direct ? dw : (float*)nullptr);
    rc = wgrad_bias(s, dz, part + (size_t)p.nsplit * 27 * s.cin * s.cout, dbias, bias_feeds_norm, st);
    if (rc) return rc;
    // ordered reduction over the split CTAs, written in PyTorch layout [co][ci][27]
    if (!direct) {
        rc = wgrad_reduce(part, nullptr, p.nsplit, s.cin, s.cout, dw, nullptr, st);
        if (rc) return rc;
    }
    return B2_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// Halo-reuse weight gradient for the thin stride-1 layers (Cin, Cout in {32, 64}): the layers that carry half of the FLOPs.
// wgrad_tc_kernel re-fetches the 128-voxel x box once per tap (9 boxes per voxel tile even when d-merged) and is bound by the
// ~43 B/clk/SM L2->SM throughput.  Here a voxel tile (2 d x 8 h x 8 w) loads
//     A: three w-shifted copies (w0-1, w0, w0+1) of the x box [2 d][10 h][8 w]   (rows = voxels, Cin channels per row)
//     B: ONE dz box [4 d = d0-1 .. d0+2][8 h][8 w]
// and every operand of the d-merged formulation is a swizzle-atom-aligned sub-view of those:
//     tap (kh, kw)  ->  copy kw, rows d*80 + (kh + hh)*8 + ww          (kh shifts by whole 8-row atoms)
//     kd            ->  dz rows (dd + 2 - kd)*64 + hh*8 + ww           (three overlapping 128-row windows, 64 rows apart)
//     D[(kh, kw, ci), (kd, co)] += sum_u x[u + (0, kh-1, kw-1)][ci] * dz[u - (kd-1, 0, 0)][co]  ==  dW[(kd,kh,kw)][ci][co].
// Bytes per voxel tile: 3*160*Cin*2 + 256*Cout*2 = 46 KB at 32->32 (d-merged per-tap kernel: 120 KB; original: 232 KB).
// MN-major UMMA operands as in wgrad_tc_kernel: M = (tap chunk, ci) with the chunk stride (LBO) chosen per accumulator group
// (kh step = 8 rows, or kw step = one copy), N = (kd, co) with LBO = 64 rows; K = 16 voxels per instruction.
// Same partial layout / ordered reduction as wgrad_tc_kernel.
// ---------------------------------------------------------------------------------------------------------------
struct WgHaloParams {
    int N, D, H, W, Cin, Cout;
    int nt_n, nt_d, nt_h, nt_w, num_vtiles;
    int ngroups;                 // accumulator groups in total (3 for Cin = 32, 5 for Cin = 64)
    int groups;                  // groups per CTA (groups * 3 * Cout <= 512 TMEM columns)
    int tapsets, nsplit, stages;
    uint32_t a_off[5], a_lbo[5]; // per group: byte offset of chunk 0 inside the A stage, chunk stride
    signed char grp_tap[5][4];   // (kh*3 + kw) of each chunk of the group, -1 = unused rows
    uint32_t idesc, tmem_cols;
};

constexpr int WH_THREADS = 32 * 6;   // warp 0: TMA producer, 1: MMA issuer + TMEM owner, 2..5: epilogue
constexpr int WH_MAX_STAGES = 4;

__global__ void __launch_bounds__(WH_THREADS, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmZ, WgHaloParams p,
                  float* __restrict__ part, float* __restrict__ dw_direct) {
    pdl_grid_sync();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full[WH_MAX_STAGES], empty[WH_MAX_STAGES], done_bar;
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const uint32_t ROWA = (uint32_t)p.Cin * 2, ROWB = (uint32_t)p.Cout * 2;
    const uint32_t COPY = 160u * ROWA, A_BYTES = 3u * COPY, B_BYTES = 256u * ROWB, STAGE = A_BYTES + B_BYTES;
    const int nper = 3 * p.Cout;

    const int split = blockIdx.x % p.nsplit, tapset = blockIdx.x / p.nsplit;
    const int g0 = tapset * p.groups;
    const int my_groups = p.ngroups - g0 < p.groups ? p.ngroups - g0 : p.groups;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmZ);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int vt = split; vt < p.num_vtiles; vt += p.nsplit) {
                int t = vt;
                const int tw = t % p.nt_w; t /= p.nt_w;
                const int th = t % p.nt_h; t /= p.nt_h;
                const int td = t % p.nt_d; t /= p.nt_d;
                const int n = t, w0 = tw * 8, h0 = th * 8, d0 = td * 2;
                mbar_wait(&empty[st], ph ^ 1);
                mbar_expect_tx(&full[st], STAGE);
                uint8_t* base = smem + (size_t)st * STAGE;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) tma_load_5d(&tmX, &full[st], base + kw * COPY, 0, w0 + kw - 1, h0 - 1, d0, n);
                tma_load_5d(&tmZ, &full[st], base + A_BYTES, 0, w0, h0, d0 - 1, n);
                if (++st == p.stages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint64_t b_hi = umma_desc_mnmajor(0, ROWB, 64u * ROWB, 8u * ROWB) & 0xFFFFFFFFFFFF0000ull;
        uint64_t a_hi[5];
        uint32_t a_off16[5];
#pragma unroll
        for (int g = 0; g < 5; ++g) {
            const int gg = g0 + g < 5 ? g0 + g : 4;
            a_hi[g] = umma_desc_mnmajor(0, ROWA, p.a_lbo[gg], 8u * ROWA) & 0xFFFFFFFFFFFF0000ull;
            a_off16[g] = p.a_off[gg] >> 4;
        }
        const uint32_t base16 = (smem_u32(smem) & 0x3FFFF) >> 4, stage16 = STAGE >> 4, ab16 = A_BYTES >> 4;
        const uint32_t a_d16 = (80u * ROWA) >> 4, a_k16 = (16u * ROWA) >> 4, b_k16 = (16u * ROWB) >> 4;
        int st = 0;
        uint32_t ph = 0;
        bool first = true;
        for (int vt = split; vt < p.num_vtiles; vt += p.nsplit) {
            mbar_wait(&full[st], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t s16 = base16 + (uint32_t)st * stage16;
#pragma unroll
                for (int g = 0; g < 5; ++g) {
                    if (g < my_groups) {
                        const uint32_t d_tmem = tmem_base + (uint32_t)(g * nper);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {   // k = (d, ks): 16 voxels of slab d
                            const uint32_t a_lo = s16 + a_off16[g] + (uint32_t)(k >> 2) * a_d16 + (uint32_t)(k & 3) * a_k16;
                            const uint32_t b_lo = s16 + ab16 + (uint32_t)k * b_k16;
                            umma_bf16(d_tmem, a_hi[g] | (uint64_t)a_lo, b_hi | (uint64_t)b_lo, p.idesc, (!first || k != 0) ? 1u : 0u);
                        }
                    }
                }
                umma_commit(&empty[st]);
            }
            __syncwarp();
            first = false;
            if (++st == p.stages) { st = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(&done_bar);
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int chunk = r / p.Cin, cil = r % p.Cin;
        const bool has_work = split < p.num_vtiles;
        if (has_work) {
            mbar_wait(&done_bar, 0);
            tc_fence_after();
        }
        float* out = part + (size_t)split * 27 * p.Cin * p.Cout;
        for (int g = 0; g < my_groups; ++g) {
            const int tap9 = p.grp_tap[g0 + g][chunk];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * nper);
            for (int c0 = 0; c0 < nper; c0 += 32) {
                uint32_t v[32];
                if (has_work) {
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = 0u;
                }
                if (tap9 >= 0) {
                    const int kdi = c0 / p.Cout, cc0 = c0 - kdi * p.Cout;
                    const int tap = (2 - kdi) * 9 + tap9;       // column block kdi holds kd = 2 - kdi
                    if (dw_direct) {
                        float* dstp = dw_direct + ((size_t)cc0 * p.Cin + cil) * 27 + tap;
                        const size_t cstride = (size_t)p.Cin * 27;
#pragma unroll
                        for (int e = 0; e < 32; ++e) dstp[e * cstride] = __uint_as_float(v[e]);
                    } else {
                        float* dstp = out + ((size_t)tap * p.Cin + cil) * p.Cout + cc0;
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(dstp + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                               __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
                    }
                }
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

int g_wgrad_halo = 1;

static bool wgrad_halo_plan(const ConvShape& s, WgHaloParams& p, size_t& smem) {
    if (!g_wgrad_halo) return false;
    if (!(s.cin == 32 || s.cin == 64) || !(s.cout == 32 || s.cout == 64)) return false;
    if (s.stride[0] != 1 || s.stride[1] != 1 || s.stride[2] != 1) return false;
    if (s.h < 8 || s.w < 8 || s.d < 2) return false;
    memset(&p, 0, sizeof(p));
    p.N = s.n; p.D = s.d; p.H = s.h; p.W = s.w; p.Cin = s.cin; p.Cout = s.cout;
    p.nt_w = cdiv(s.w, 8); p.nt_h = cdiv(s.h, 8); p.nt_d = cdiv(s.d, 2); p.nt_n = s.n;
    p.num_vtiles = p.nt_w * p.nt_h * p.nt_d * p.nt_n;
    const uint32_t ROWA = s.cin * 2, COPY = 160u * ROWA;
    for (int g = 0; g < 5; ++g)
        for (int c = 0; c < 4; ++c) p.grp_tap[g][c] = -1;
    if (s.cin == 32) {
        // one group per kw: chunks kh = 0, 1, 2 (8 rows apart) + one unused chunk
        p.ngroups = 3;
        for (int kw = 0; kw < 3; ++kw) {
            p.a_off[kw] = kw * COPY; p.a_lbo[kw] = 8u * ROWA;
            for (int kh = 0; kh < 3; ++kh) p.grp_tap[kw][kh] = (signed char)(kh * 3 + kw);
        }
    } else {
        // two chunks per group: (kh 0, kh 1) of each kw; (kw 0, kw 1) of kh 2 (one copy apart); (kh 2, kw 2) + unused
        p.ngroups = 5;
        for (int kw = 0; kw < 3; ++kw) {
            p.a_off[kw] = kw * COPY; p.a_lbo[kw] = 8u * ROWA;
            p.grp_tap[kw][0] = (signed char)kw; p.grp_tap[kw][1] = (signed char)(3 + kw);
        }
        p.a_off[3] = 16u * ROWA; p.a_lbo[3] = COPY;
        p.grp_tap[3][0] = 6; p.grp_tap[3][1] = 7;
        p.a_off[4] = 2u * COPY + 16u * ROWA; p.a_lbo[4] = 8u * ROWA;
        p.grp_tap[4][0] = 8;
    }
    const int nper = 3 * s.cout;
    p.groups = 512 / nper < p.ngroups ? 512 / nper : p.ngroups;
    p.tapsets = cdiv(p.ngroups, p.groups);
    int ns = num_sms() / p.tapsets;
    if (ns < 1) ns = 1;
    if (ns > p.num_vtiles) ns = p.num_vtiles;
    p.nsplit = ns;
    const size_t stage = 3ull * COPY + 256ull * s.cout * 2;
    int st = (int)((200ull * 1024) / stage);
    if (st > WH_MAX_STAGES) st = WH_MAX_STAGES;
    if (st < 2) return false;
    p.stages = st;
    smem = (size_t)st * stage + 1024;
    uint32_t cols = 32;
    while (cols < (uint32_t)(p.groups * nper)) cols *= 2;
    p.tmem_cols = cols;
    p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(nper >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    return true;
}

// 5D activation map with an explicit box (channels, w, h, d, n)
static int make_box_map(CUtensorMap* m, const void* ptr, int N, int D, int H, int W, int C, int pitch, int bc, int bw, int bh, int bd) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(B2_ECUDA, "cuTensorMapEncodeTiled entry point not available%s", "");
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t gstr[4] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2, (cuuint64_t)D * H * W * pitch * 2};
    cuuint32_t box[5] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bd, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     bc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B2_ECUDA, "cuTensorMapEncodeTiled(box) failed: code %s%lld", "", (long long)r);
    return B2_OK;
}

// The bias of a conv that feeds InstanceNorm has an exactly-zero gradient in exact arithmetic (the norm removes the
// per-channel mean); PyTorch's value there is pure rounding noise, so the network plan asks for the exact value.
static int wgrad_bias(const ConvShape& s, const __nv_bfloat16* dz, float* part_b, float* dbias, bool bias_feeds_norm, cudaStream_t st) {
    if (!dbias) return B2_OK;
    if (bias_feeds_norm) {
        B2_CUDA(cudaMemsetAsync(dbias, 0, s.cout * sizeof(float), st));
        return B2_OK;
    }
    const int Do = (s.d - 1) / s.stride[0] + 1, Ho = (s.h - 1) / s.stride[1] + 1, Wo = (s.w - 1) / s.stride[2] + 1;
    const long long rows = (long long)s.n * Do * Ho * Wo;
    int slabs = 2 * num_sms();
    if (slabs > rows) slabs = (int)rows;
    if (s.cout <= 256) {
        const int lanes = 256 / s.cout;
        B2_LAUNCH(colsum_part_kernel, slabs, 256, (size_t)lanes * s.cout * sizeof(float), st, dz, rows, s.cout, s.out_pitch, slabs, part_b);
    } else {
        B2_LAUNCH(colsum_wide_kernel, slabs, 256, 0, st, dz, rows, s.cout, s.out_pitch, slabs, part_b);
    }
    B2_LAUNCH(colsum_final_kernel, cdiv(s.cout, 128), 128, 0, st, part_b, slabs, s.cout, dbias);
    return B2_OK;
}

static bool wgrad_halo_ok(const ConvShape& s) {
    WgHaloParams p;
    size_t smem;
    return wgrad_halo_plan(s, p, smem) && s.in_pitch % 8 == 0 && s.out_pitch % 8 == 0;
}

static int conv3d_wgrad_halo(const ConvShape& s, const __nv_bfloat16* x, const __nv_bfloat16* dz, float* part, float* dw, float* dbias,
                             bool bias_feeds_norm, cudaStream_t st) {
    WgHaloParams p;
    size_t smem;
    if (!wgrad_halo_plan(s, p, smem)) return fail(B2_EUNSUPPORTED, "wgrad_halo: unsupported shape%s", "");
    CUtensorMap tmX, tmZ;
    int rc = make_box_map(&tmX, x, s.n, s.d, s.h, s.w, s.cin, s.in_pitch, s.cin, 8, 10, 2);
    if (rc) return rc;
    rc = make_box_map(&tmZ, dz, s.n, s.d, s.h, s.w, s.cout, s.out_pitch, s.cout, 8, 8, 4);
    if (rc) return rc;
    static bool attr = false;
    if (!attr) { B2_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    const bool direct = g_wgrad_direct && p.nsplit == 1;
    B2_LAUNCH(wgrad_halo_kernel, p.tapsets * p.nsplit, WH_THREADS, smem, st, tmX, tmZ, p, part, direct ? dw : (float*)nullptr);
    rc = wgrad_bias(s, dz, part + (size_t)p.nsplit * 27 * s.cin * s.cout, dbias, bias_feeds_norm, st);
    if (rc) return rc;
    if (!direct) {
        rc = wgrad_reduce(part, nullptr, p.nsplit, s.cin, s.cout, dw, nullptr, st);
        if (rc) return rc;
    }
    return B2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// transposed-conv (kernel == stride) weight gradient on the same kernel: dW[ci][co][q] = sum_v x[v][ci] * dy[k*v + q][co].
// The shifted operand is dy ("taps" = q, traversal stride k), the fixed operand is x: D[(q, co)][ci].
// ---------------------------------------------------------------------------------------------------------------
__global__ void tconv_wgrad_reduce_kernel(const float* __restrict__ part, int nsplit, int k8, int Cin, int Cout, float* __restrict__ dw) {
    pdl_grid_sync();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over [q][co][ci]
    const long long tot = (long long)k8 * Cout * Cin;
    if (i >= tot) return;
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += part[(long long)k * tot + i];
    const int ci = (int)(i % Cin);
    long long r = i / Cin;
    const int co = (int)(r % Cout), q = (int)(r / Cout);
    dw[((long long)ci * Cout + co) * k8 + q] = s;
}

static void tconv_wgrad_tc_plan(const TconvShape& s, TcWgradParams& p) {
    const int k8 = s.k[0] * s.k[1] * s.k[2];
    // roles: "Cin" of the kernel = channels of the shifted operand (dy: Cout), "Cout" = channels of the fixed one (x: Cin)
    ConvShape c;
    c.n = s.n; c.d = s.d * s.k[0]; c.h = s.h * s.k[1]; c.w = s.w * s.k[2];
    c.cin = s.cout; c.cout = s.cin;
    c.stride[0] = s.k[0]; c.stride[1] = s.k[1]; c.stride[2] = s.k[2];
    c.in_pitch = s.out_pitch; c.out_pitch = s.in_pitch;
    wgrad_tc_plan(c, p, k8);
    // the voxel boxes tile the INPUT grid of the transposed conv (wgrad_tc_plan derives it as (dim-1)/k+1 == s.d ...)
    for (int q = 0; q < k8; ++q) {
        p.tap_off[q][2] = q % s.k[2]; p.tap_off[q][1] = (q / s.k[2]) % s.k[1]; p.tap_off[q][0] = q / (s.k[2] * s.k[1]);
    }
}

bool tconv_wgrad_tc_supported(int cin, int cout) { return wgrad_tc_supported(cout, cin); }

size_t tconv_wgrad_tc_part_floats(const TconvShape& s) {
    TcWgradParams p;
    tconv_wgrad_tc_plan(s, p);
    return (size_t)p.nsplit * p.ntaps * s.cin * s.cout + 64;
}

int tconv_wgrad_tc(const TconvShape& s, const __nv_bfloat16* x, const __nv_bfloat16* dy, float* part, float* dw, cudaStream_t st) {
    B2_CHECK_ARG(tconv_wgrad_tc_supported(s.cin, s.cout) && s.in_pitch % 8 == 0 && s.out_pitch % 8 == 0);
    TcWgradParams p;
    tconv_wgrad_tc_plan(s, p);
    CUtensorMap tmA, tmB;   // A: dy (shifted, strided traversal), B: x (fixed)
    int rc = make_act_map(&tmA, dy, s.n, s.d * s.k[0], s.h * s.k[1], s.w * s.k[2], s.cout, s.out_pitch, p.ci_sub, p.TN, p.TD, p.TH, p.TW,
                          s.k[0], s.k[1], s.k[2]);
    if (rc) return rc;
    rc = make_act_map(&tmB, x, s.n, s.d, s.h, s.w, s.cin, s.in_pitch, p.co_sub, p.TN, p.TD, p.TH, p.TW, 1, 1, 1);
    if (rc) return rc;
    const uint32_t A_BYTES = 128u * p.ci_sub * 2 * p.a_chunks, B_BYTES = 128u * p.nper * 2;
    const size_t smem = (size_t)p.a_stages * A_BYTES + 2 * (size_t)B_BYTES + 1024;
    if (smem > 220 * 1024) return fail(B2_EUNSUPPORTED, "tconv_wgrad_tc: tile does not fit shared memory%s", "");
    static bool attr = false;
    if (!attr) { B2_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    const int grid = p.ci_items * p.co_blks * p.tapsets * p.nsplit;
    B2_LAUNCH(wgrad_tc_kernel, grid, WG_THREADS, smem, st, tmA, tmB, p, part, (float*)nullptr);
    const long long tot = (long long)p.ntaps * s.cin * s.cout;
    B2_LAUNCH(tconv_wgrad_reduce_kernel, cdiv(tot, 256), 256, 0, st, part, p.nsplit, p.ntaps, s.cin, s.cout, dw);
    return B2_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// First layer (Cin <= 1..4, 27*Cin <= 32): im2col ONLY for this layer -- a [voxels][32] bf16 patch matrix P (27*Cin taps +
// zero padding, 64 B per voxel) turns the layer into a 1-tap GEMM that the tensor-core kernels above handle:
// forward z = P * Wp^T (conv_tc_gather, ntaps = 1), weight gradient dWp = P^T dz (wgrad_tc_kernel, ntaps = 1).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patch_matrix_kernel(const __nv_bfloat16* __restrict__ x, int N, int D, int H, int W, int cin,
                                                           int x_pitch, __nv_bfloat16* __restrict__ P) {
    pdl_grid_sync();
    // thread = voxel: gathers its 27*cin neighbours (L1-resident: adjacent threads share them) and writes the 64-byte row
    // with four 16-byte stores -- a warp writes 2 KB contiguous.  (The first version used a thread per 8 columns with a
    // div/mod chain per element: 286 us for 134 MB; this one is store-bound.)
    const long long total = (long long)N * D * H * W;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(v % W); long long r = v / W;
        const int h = (int)(r % H); r /= H;
        const int d = (int)(r % D);
        const int n = (int)(r / D);
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) packed[j] = 0u;
#pragma unroll
        for (int t = 0; t < 27; ++t) {   // cin == 1 (27 * cin <= 32): column k == tap t
            const int id = d + t / 9 - 1, ih = h + (t / 3) % 3 - 1, iw = w + t % 3 - 1;
            const bool in = id >= 0 && id < D && ih >= 0 && ih < H && iw >= 0 && iw < W;
            const unsigned short u = in ? __bfloat16_as_ushort(x[((((long long)n * D + id) * H + ih) * W + iw) * x_pitch]) : (unsigned short)0;
            packed[t >> 1] |= (uint32_t)u << ((t & 1) * 16);
        }
        uint4* dst = reinterpret_cast<uint4*>(P + v * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
    }
}

// PyTorch [Cout][Cin][27] -> Wp [Cout][32] bf16 (k = t*Cin + ci, zero padded)
__global__ void patch_weight_kernel(const float* __restrict__ w, int cout, int cin, __nv_bfloat16* __restrict__ wp) {
    pdl_grid_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cout * 32) return;
    const int k = i & 31, co = i >> 5;
    float v = 0.f;
    if (k < 27 * cin) { const int t = k / cin, ci = k % cin; v = w[((long long)co * cin + ci) * 27 + t]; }
    wp[i] = __float2bfloat16_rn(v);
}

__global__ void patch_wgrad_reduce_kernel(const float* __restrict__ part, int nsplit, int cin, int cout, float* __restrict__ dw) {
    pdl_grid_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over [k < 32][co]
    if (i >= 32 * cout) return;
    const int co = i % cout, k = i / cout;
    if (k >= 27 * cin) return;
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += part[(long long)sp * 32 * cout + i];
    const int t = k / cin, ci = k % cin;
    dw[((long long)co * cin + ci) * 27 + t] = s;
}

bool first_layer_tc_supported(int cin, int cout) { return 27 * cin <= 32 && cout % 32 == 0 && cout <= 256; }

int first_layer_patches(const __nv_bfloat16* x, int N, int D, int H, int W, int cin, int x_pitch, __nv_bfloat16* P, const float* w_pt,
                        int cout, __nv_bfloat16* wp, cudaStream_t st) {
    B2_CHECK_ARG(cin == 1);
    const long long total = (long long)N * D * H * W;
    long long grid = (total + 255) / 256, cap = (long long)num_sms() * 32;
    if (grid > cap) grid = cap;
    B2_LAUNCH(patch_matrix_kernel, (int)grid, 256, 0, st, x, N, D, H, W, cin, x_pitch, P);
    if (wp) B2_LAUNCH(patch_weight_kernel, cdiv(cout * 32, 256), 256, 0, st, w_pt, cout, cin, wp);
    return B2_OK;
}

static void first_wgrad_plan(int N, int D, int H, int W, int cout, int dz_pitch, TcWgradParams& p) {
    ConvShape c;
    c.n = N; c.d = D; c.h = H; c.w = W; c.cin = 32; c.cout = cout;
    c.stride[0] = c.stride[1] = c.stride[2] = 1;
    c.in_pitch = 32; c.out_pitch = dz_pitch;
    wgrad_tc_plan(c, p, 1);
    p.tap_off[0][0] = p.tap_off[0][1] = p.tap_off[0][2] = 0;
}

size_t first_layer_wgrad_part_floats(int N, int D, int H, int W, int cout) {
    TcWgradParams p;
    first_wgrad_plan(N, D, H, W, cout, cout, p);
    return (size_t)p.nsplit * 32 * cout + 64;
}

int first_layer_wgrad_tc(const __nv_bfloat16* P, const __nv_bfloat16* dz, int N, int D, int H, int W, int cin, int cout, int dz_pitch,
                         float* part, float* dw, float* dbias, cudaStream_t st) {
    TcWgradParams p;
    first_wgrad_plan(N, D, H, W, cout, dz_pitch, p);
    CUtensorMap tmA, tmB;
    int rc = make_act_map(&tmA, P, N, D, H, W, 32, 32, p.ci_sub, p.TN, p.TD, p.TH, p.TW, 1, 1, 1);
    if (rc) return rc;
    rc = make_act_map(&tmB, dz, N, D, H, W, cout, dz_pitch, p.co_sub, p.TN, p.TD, p.TH, p.TW, 1, 1, 1);
    if (rc) return rc;
    const uint32_t A_BYTES = 128u * p.ci_sub * 2 * p.a_chunks, B_BYTES = 128u * p.nper * 2;
    const size_t smem = (size_t)p.a_stages * A_BYTES + 2 * (size_t)B_BYTES + 1024;
    static bool attr = false;
    if (!attr) { B2_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
    const int grid = p.ci_items * p.co_blks * p.tapsets * p.nsplit;
    B2_LAUNCH(wgrad_tc_kernel, grid, WG_THREADS, smem, st, tmA, tmB, p, part, (float*)nullptr);
    B2_LAUNCH(patch_wgrad_reduce_kernel, cdiv(32 * cout, 256), 256, 0, st, part, p.nsplit, cin, cout, dw);
    if (dbias) B2_CUDA(cudaMemsetAsync(dbias, 0, cout * sizeof(float), st));   // bias feeds InstanceNorm: exactly zero
    return B2_OK;
}

}  // namespace b2
