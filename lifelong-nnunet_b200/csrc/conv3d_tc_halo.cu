// conv3d_tc_halo.cu -- stride-1 3x3x3 bf16 convolution for the thin layers (K = 32 or 64 input channels), where the
// per-tap kernel of conv3d_tc.cu is bound by L2->SM traffic (every 128-voxel A tile is fetched 27 times: FLOP/byte of
// L2 traffic == Cout).  Here a CTA walks a strip of 16 (h) x 8 (w) output voxels along d and keeps a ring of up to six input
// d-slabs in shared memory; every slab is fetched ONCE as three w-shifted copies (w0-1, w0, w0+1) of an 18 (h) x 8 (w) box,
// so that the A operand of every tap (kd, kh, kw) is a plain, swizzle-atom-aligned sub-view:
//     slab (d + kd - 1) -> copy kw -> atom row kh .. kh + 15      (atom = 8 consecutive w voxels = 8 rows of KC*2 bytes)
// i.e. no per-tap loads at all: L2->SM traffic drops from 27 to 3.4 fetches per input voxel.  The 27 weight tiles stay
// resident in shared memory; when [27][Nout][K] does not fit, the output channels are split over CTA classes (halo_plan).
// tcgen05.mma M = 128 (16 h x 8 w), K = 16, N = 3 * BN: the three kd taps of a (kh, kw) pair are ONE instruction that
// updates the accumulators of the output slabs d-1, d, d+1 (consecutive slots of a 16-slot TMEM ring) -- see the comment
// above conv_halo_kernel.  Warp roles: 0 slab producer, 1 MMA issuer, 2..5 epilogue (bias, bf16 store, InstanceNorm
// partial sums), 6 weight producer, 7..10 a second epilogue set.  ncu (profiles/r02_ncu_halo32_source.txt) showed the round-1
// kernel EPILOGUE-bound: the four epilogue warps were busy 85 % of the time (47 % of that in the per-slab 32 x 32
// transpose-reductions of the InstanceNorm sums, 28 % in the TMEM load + 32 scalar bias loads) while the MMA warp waited
// 45 % of its time.  Now (i) two epilogue sets drain alternate output slabs, (ii) for BN = 32 every thread keeps running
// per-column sums of its own voxel row in registers and the cross-lane reduction happens ONCE per (CTA, sample), (iii) the
// bias is loaded with 16-byte loads.
#include <string.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace b2 {

struct HaloParams {
    int N, D, H, W;       // produced tensor == gathered tensor extent (stride 1, padding 1)
    int dst_pitch;
    int BN;               // output channels per CTA (<= 128)
    int NS, Ntot;         // column splits (CTA c owns columns [(c % NS) * BN, +BN) for its whole life) and total channels
    int w_pitch, w_row0;  // rows per tap of the weight matrix and first row of the computed column window (w_pitch >= Ntot)
    int HB, WB, DS, SD;   // strips per sample along h / w, segments along d and their length
    int num_items;
    int nslot;            // slab ring depth (2..6)
    int nacc, nacc_log2;  // TMEM accumulator ring: nacc slots of BN columns
    int merge;            // 1: the three kd taps of a (kh, kw) pair are ONE tcgen05.mma of N = 3*BN (see below)
    uint32_t idesc, idesc2, idesc3;   // instruction descriptors for N = BN, 2*BN, 3*BN
    uint32_t tmem_cols;
    int dbg;              // probe switches (b2_set_option("halo_dbg")): 1 no global stores, 2 no MMA issue, 4 no TMA slab loads, 8 no TMEM loads, 16 empty epilogue, 32 arrive instead of commit, 64 no tempty wait
};

constexpr int HALO_THREADS = 352;  // warp 0: slab producer, 1: MMA, 2..5: epilogue set 0, 6: weight producer, 7..10: epilogue set 1
constexpr int HALO_MAX_SLOTS = 6;
constexpr int HALO_MAX_ACC = 16;

// kd-merged issue ("input-stationary along d").  Input slab s feeds the three output slabs s+1 (kd = 0), s (kd = 1) and
// s-1 (kd = 2) with the SAME A operand per (kh, kw).  Their accumulators sit in consecutive TMEM slots of the ring and the
// three weight tiles sit back to back in shared memory in the order kd = 2, 1, 0, so one instruction of N = 3*BN updates
// all three:   D[:, (s-1 | s | s+1) x BN] += A(slab s, kh, kw) * [W(2,kh,kw); W(1,kh,kw); W(0,kh,kw)]^T.
// An SS-mode M=128 tcgen05.mma costs max(~45, N/2) cycles (profiles/r01_mma_probe.txt): at BN = 32 three N=32 MMAs (135
// cycles) become one N=96 MMA (48 cycles).  The very first MMA of a slab is split in two (N = 2*BN accumulate + N = BN
// overwrite) because slot s+1 must be zero-initialised; slabs at a segment edge or where the three slots wrap around the
// ring take the per-kd path.
template <int KC>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, HaloParams p,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ dst, int accumulate, EpiStats es) {
    pdl_grid_sync();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t sfull[HALO_MAX_SLOTS], sempty[HALO_MAX_SLOTS], wfull, tfull[HALO_MAX_ACC], tempty[HALO_MAX_ACC];
    __shared__ uint32_t tmem_base_smem;

    constexpr uint32_t ROW = KC * 2;
    constexpr uint32_t COPY_BYTES = 144 * ROW;       // 18 h x 8 w rows
    constexpr uint32_t SLAB_BYTES = 3 * COPY_BYTES;  // three w-shifted copies
    const uint32_t B_TILE = (uint32_t)p.BN * ROW;    // one tap of weights [BN][KC]
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* slabs = smem;
    uint8_t* wsm = smem + (size_t)p.nslot * SLAB_BYTES;   // 27 resident weight tiles, order (kh*3+kw, 2-kd)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.nslot; ++s) { mbar_init(&sfull[s], 1); mbar_init(&sempty[s], 1); }
        mbar_init(&wfull, 1);
        for (int a = 0; a < p.nacc; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }   // one arrive per epilogue warp
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t amask = (uint32_t)p.nacc - 1;
    const int cs = (int)blockIdx.x % p.NS;                    // this CTA's column split
    const int item0 = (int)blockIdx.x / p.NS, item_step = (int)gridDim.x / p.NS;

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
            int slot = 0;
            uint32_t ph = 0;
            for (int item = item0; item < p.num_items; item += item_step) {
                int n, h0, w0, d0, d1;
                decode(item, n, h0, w0, d0, d1);
                for (int s = d0 - 1; s <= d1; ++s) {
                    mbar_wait(&sempty[slot], ph ^ 1);
                    if (p.dbg & 4) { mbar_arrive(&sfull[slot]); if (++slot == p.nslot) { slot = 0; ph ^= 1; } continue; }
                    mbar_expect_tx(&sfull[slot], SLAB_BYTES);
                    uint8_t* base = slabs + (size_t)slot * SLAB_BYTES;
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw)
                        tma_load_5d(&tmA, &sfull[slot], base + kw * COPY_BYTES, 0, w0 + kw - 1, h0 - 1, s, n);
                    if (++slot == p.nslot) { slot = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 6) {
        // ===================== weight producer =====================
        if (lane == 0) {
            mbar_expect_tx(&wfull, 27 * B_TILE);
            for (int t = 0; t < 27; ++t) {
                const int kd = t / 9, t9 = t % 9;
                tma_load_2d(&tmB, &wfull, wsm + (size_t)(t9 * 3 + 2 - kd) * B_TILE, 0, t * p.w_pitch + p.w_row0 + cs * p.BN);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one elected lane, straight-line code per slab) =====================
        // The issuing thread is blocked per tcgen05.mma and every scalar instruction between two MMAs adds to that:
        // descriptors are assembled from precomputed 32-bit halves and the taps are fully unrolled (offsets are immediates).
        {
            constexpr uint32_t DESC_HI = (uint32_t)(((uint64_t)((8 * ROW) >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)(ROW == 128 ? 2 : 4) << 61) >> 32);
            auto mk = [](uint32_t lo) -> uint64_t { return ((uint64_t)DESC_HI << 32) | (uint64_t)lo; };
            const uint32_t slab_lo0 = ((smem_u32(slabs) & 0x3FFFF) >> 4) | 0x10000u;
            constexpr uint32_t SLAB16 = SLAB_BYTES >> 4, C16 = COPY_BYTES >> 4, A16 = (8 * ROW) >> 4;
            const uint32_t w_lo = ((smem_u32(wsm) & 0x3FFFF) >> 4) | 0x10000u;
            const uint32_t btile16 = B_TILE >> 4;
            int slot = 0;
            uint32_t sph = 0;
            uint32_t ocb = 0;   // running output counter at d0 of the current item
            mbar_wait(&wfull, 0);
            for (int item = item0; item < p.num_items; item += item_step) {
                int n, h0, w0, d0, d1;
                decode(item, n, h0, w0, d0, d1);
                for (int s = d0 - 1; s <= d1; ++s) {
                    mbar_wait(&sfull[slot], sph);
                    // output s+1 starts with this slab (kd = 0): its accumulator must have been drained
                    const bool v0 = (s + 1 >= d0) && (s + 1 < d1), v1 = (s >= d0) && (s < d1), v2 = (s - 1 >= d0) && (s - 1 < d1);
                    const uint32_t oc0 = ocb + (uint32_t)(s + 1 - d0), oc1 = oc0 - 1, oc2 = oc0 - 2;
                    if (v0 && !(p.dbg & 64)) mbar_wait(&tempty[oc0 & amask], ((oc0 >> p.nacc_log2) & 1) ^ 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_base = slab_lo0 + (uint32_t)slot * SLAB16;
                        if (p.dbg & 2) {
                        } else if (p.merge && v0 && v1 && v2 && (oc0 & amask) >= 2) {
                            const uint32_t d2 = tmem_base + (oc2 & amask) * (uint32_t)p.BN;
                            const uint32_t d0t = d2 + 2u * (uint32_t)p.BN;
#pragma unroll
                            for (int t9 = 0; t9 < 9; ++t9) {
                                const uint32_t a_lo = a_base + (uint32_t)(t9 % 3) * C16 + (uint32_t)(t9 / 3) * A16;
                                const uint32_t b_lo = w_lo + (uint32_t)(t9 * 3) * btile16;
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k) {
                                    if (t9 == 0 && k == 0) {
                                        umma_bf16(d2, mk(a_lo), mk(b_lo), p.idesc2, 1u);                     // outputs s-1, s
                                        umma_bf16(d0t, mk(a_lo), mk(b_lo + 2 * btile16), p.idesc, 0u);       // output s+1: first write
                                    } else {
                                        umma_bf16(d2, mk(a_lo + 2 * k), mk(b_lo + 2 * k), p.idesc3, 1u);
                                    }
                                }
                            }
                        } else if (p.merge && v0 && v1 && !v2 && (oc0 & amask) >= 1) {
                            // second slab of a segment: outputs s (accumulate) and s+1 (first write) -- two-wide merge
                            const uint32_t d1t = tmem_base + (oc1 & amask) * (uint32_t)p.BN;
                            const uint32_t d0t = d1t + (uint32_t)p.BN;
#pragma unroll
                            for (int t9 = 0; t9 < 9; ++t9) {
                                const uint32_t a_lo = a_base + (uint32_t)(t9 % 3) * C16 + (uint32_t)(t9 / 3) * A16;
                                const uint32_t b_lo = w_lo + (uint32_t)(t9 * 3 + 1) * btile16;      // tiles kd = 1, kd = 0
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k) {
                                    if (t9 == 0 && k == 0) {
                                        umma_bf16(d1t, mk(a_lo), mk(b_lo), p.idesc, 1u);
                                        umma_bf16(d0t, mk(a_lo), mk(b_lo + btile16), p.idesc, 0u);
                                    } else {
                                        umma_bf16(d1t, mk(a_lo + 2 * k), mk(b_lo + 2 * k), p.idesc2, 1u);
                                    }
                                }
                            }
                        } else if (p.merge && !v0 && v1 && v2 && (oc1 & amask) >= 1) {
                            // last-but-one slab of a segment: outputs s-1 and s, both accumulate -- two-wide merge
                            const uint32_t d2 = tmem_base + (oc2 & amask) * (uint32_t)p.BN;
#pragma unroll
                            for (int t9 = 0; t9 < 9; ++t9) {
                                const uint32_t a_lo = a_base + (uint32_t)(t9 % 3) * C16 + (uint32_t)(t9 / 3) * A16;
                                const uint32_t b_lo = w_lo + (uint32_t)(t9 * 3) * btile16;          // tiles kd = 2, kd = 1
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k)
                                    umma_bf16(d2, mk(a_lo + 2 * k), mk(b_lo + 2 * k), p.idesc2, 1u);
                            }
                        } else {
#pragma unroll
                            for (int kd = 0; kd < 3; ++kd) {
                                const bool valid = kd == 0 ? v0 : kd == 1 ? v1 : v2;
                                if (valid) {
                                    const uint32_t oc = kd == 0 ? oc0 : kd == 1 ? oc1 : oc2;
                                    const uint32_t d_tmem = tmem_base + (oc & amask) * (uint32_t)p.BN;
#pragma unroll
                                    for (int t9 = 0; t9 < 9; ++t9) {
                                        const uint32_t a_lo = a_base + (uint32_t)(t9 % 3) * C16 + (uint32_t)(t9 / 3) * A16;
                                        const uint32_t b_lo = w_lo + (uint32_t)(t9 * 3 + 2 - kd) * btile16;
#pragma unroll
                                        for (int k = 0; k < KC / 16; ++k)
                                            umma_bf16(d_tmem, mk(a_lo + 2 * k), mk(b_lo + 2 * k), p.idesc, (kd | t9 | k) != 0 ? 1u : 0u);
                                    }
                                }
                            }
                        }
                        if (p.dbg & 32) {      // probe: plain arrives instead of tcgen05.commit (only meaningful with the MMAs off)
                            mbar_arrive(&sempty[slot]);
                            if (v2) mbar_arrive(&tfull[oc2 & amask]);
                        } else {
                            umma_commit(&sempty[slot]);                        // slab fully consumed
                            if (v2) umma_commit(&tfull[oc2 & amask]);          // output s-1 complete (its kd = 2 taps were the last)
                        }
                    }
                    __syncwarp();
                    if (++slot == p.nslot) { slot = 0; sph ^= 1; }
                }
                ocb += (uint32_t)(d1 - d0);
            }
        }
    } else {
        // ===================== epilogue (two sets of four warps; set e drains the output slabs with oc % 2 == e) ================
        const int eset = warp >= 7 ? 1 : 0;
        const int q = warp & 3;                 // TMEM lane quadrant this warp may read
        const int r = q * 32 + lane;
        const int w_ = r & 7, h_ = r >> 3;
        uint32_t oc = 0;   // running output counter: accumulator oc & amask, phase (oc >> nacc_log2) & 1
        // InstanceNorm statistics of the produced tensor (fp32 accumulator values, before the bf16 rounding)
        const bool fast = es.part != nullptr && p.BN == 32;   // per-thread running sums, one reduction per (CTA, sample)
        float s1[32], s2[32];                                 // fast: column sums of THIS thread's voxel row
        float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};   // generic: lane = column
#pragma unroll
        for (int e = 0; e < 32; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
        int n_cur = -1, n_done = 0;
        const int st_slot = ((int)blockIdx.x / p.NS) * 8 + eset * 4 + q;
        auto st_write = [&](int nn, bool zero) {
            if (fast) {
                float a1 = 0.f, a2 = 0.f;
                if (!zero) { a1 = transpose_reduce32(s1, lane); a2 = transpose_reduce32(s2, lane); }
                *reinterpret_cast<float2*>(es.part + (((long long)nn * es.slots + st_slot) * es.C + cs * p.BN + lane) * 2) = make_float2(a1, a2);
                return;
            }
            for (int ch = 0; ch < p.BN / 32; ++ch) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k)   // (select instead of a runtime index: keeps st1 / st2 in registers)
                    if (k == ch && !zero) { a1 = st1[k]; a2 = st2[k]; }
                *reinterpret_cast<float2*>(es.part + (((long long)nn * es.slots + st_slot) * es.C + cs * p.BN + ch * 32 + lane) * 2) = make_float2(a1, a2);
            }
        };
        for (int item = item0; item < p.num_items; item += item_step) {
            int n, h0, w0, d0, d1;
            decode(item, n, h0, w0, d0, d1);
            if (es.part && n != n_cur) {
                if (n_cur >= 0) { st_write(n_cur, false); n_done = n_cur + 1; }
                for (int nz = n_done; nz < n; ++nz) st_write(nz, true);
                n_done = n; n_cur = n;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) { st1[ch] = 0.f; st2[ch] = 0.f; }
#pragma unroll
                for (int e = 0; e < 32; ++e) { s1[e] = 0.f; s2[e] = 0.f; }
            }
            const int oh = h0 + h_, ow = w0 + w_;
            const bool valid = oh < p.H && ow < p.W;
            for (int od = d0; od < d1; ++od, ++oc) {
                if ((int)(oc & 1u) != eset) continue;
                const uint32_t acc = oc & amask;
                __nv_bfloat16* row = dst + ((((long long)n * p.D + od) * p.H + oh) * p.W + ow) * p.dst_pitch + cs * p.BN;
                mbar_wait(&tfull[acc], (oc >> p.nacc_log2) & 1);
                tc_fence_after();
                if (p.dbg & 16) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(&tempty[acc]); continue; }
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (uint32_t)p.BN;
#pragma unroll 1
                for (int c0 = 0; c0 < p.BN; c0 += 32) {
                    uint32_t v[32];
                    if (p.dbg & 8) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = 0u;
                    } else {
                        tmem_ld32(taddr + c0, v);
                    }
                    float f[32];
                    if (bias) {     // (issued under the TMEM load's latency; 16-byte loads, L1-resident)
                        const float4* b4 = reinterpret_cast<const float4*>(bias + cs * p.BN + c0);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 b = __ldg(b4 + e);
                            f[4 * e] = b.x; f[4 * e + 1] = b.y; f[4 * e + 2] = b.z; f[4 * e + 3] = b.w;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) f[e] = 0.f;
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 32; ++e) f[e] += __uint_as_float(v[e]);
                    if (valid && !((p.dbg & 1) && f[0] != 12345.f)) {
#pragma unroll
                        for (int jj = 0; jj < 32; jj += 8) {
                            float o8[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) o8[e] = f[jj + e];
                            if (accumulate) {
                                float o[8];
                                load8(row + c0 + jj, o);
#pragma unroll
                                for (int e = 0; e < 8; ++e) o8[e] += o[e];
                            }
                            store8(row + c0 + jj, o8);
                        }
                    }
                    if (fast) {
                        if (valid) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) { s1[e] += f[e]; s2[e] = fmaf(f[e], f[e], s2[e]); }
                        }
                    } else if (es.part) {
                        float sq[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            f[e] = valid ? f[e] : 0.f;
                            sq[e] = f[e] * f[e];
                        }
                        const float a1 = transpose_reduce32(f, lane), a2 = transpose_reduce32(sq, lane);
                        const int ch = c0 >> 5;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k == ch) { st1[k] += a1; st2[k] += a2; }
                    }
                }
                // one elected arrive per warp: 128 per-thread arrives on one mbarrier word are 128 serialised shared-memory
                // atomics per slab, on the port the MMA operand reads saturate
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
            }
        }
        if (es.part) {
            if (n_cur >= 0) { st_write(n_cur, false); n_done = n_cur + 1; }
            for (int nz = n_done; nz < es.N; ++nz) st_write(nz, true);
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

int g_halo_merge = 1;
int g_halo_dbg = 0;
int g_epi_stats = 1;      // InstanceNorm statistics from the convolution epilogues (no second pass over z)
int g_halo_nsplit = 1;   // allow splitting the output channels over CTA classes when the weights do not fit

static bool halo_plan(int K, int Nout, int D, int H, int W, int N, HaloParams& p, size_t& smem) {
    if (!(K == 32 || K == 64) || Nout % 32 != 0 || Nout > 256) return false;
    if (H < 16 || W < 8) return false;
    memset(&p, 0, sizeof(p));
    p.N = N; p.D = D; p.H = H; p.W = W; p.Ntot = Nout;
    p.HB = cdiv(H, 16); p.WB = cdiv(W, 8);
    const uint32_t row = K * 2, slab = 3 * 144 * row;
    const size_t budget = 223 * 1024;
    // The 27 weight tiles of a CTA must be resident (streaming them costs one mbarrier round trip per tap on the MMA thread
    // and 9x the slab bytes in L2 traffic).  When [27][Nout][K] does not fit next to >= 2 slabs the output channels are split
    // over NS CTAs-classes of BN = Nout / NS columns: every class re-reads the slabs, but at BN = 32 the kd-merged MMA
    // (N = 96) already runs at the full tensor rate and 2 x 55 KB per 3456 MMA cycles stays under the ~43 B/clk/SM L2 cap --
    // against 27 x 24 KB per 128 voxels in the per-tap kernel of conv3d_tc.cu.
    int ns = 1;
    while (ns <= 8 && (Nout % ns != 0 || (Nout / ns) % 32 != 0 || Nout / ns > 128 ||
                       2ull * slab + 27ull * (Nout / ns) * row + 1024 > budget)) ++ns;
    if (ns > 8) return false;
    if (ns > 1 && !g_halo_nsplit) return false;
    p.NS = ns; p.BN = Nout / ns;
    const uint32_t btile = (uint32_t)p.BN * row;
    int nslot = (int)((budget - 27ull * btile - 1024) / slab);
    if (nslot > HALO_MAX_SLOTS) nslot = HALO_MAX_SLOTS;
    p.nslot = nslot;
    smem = (size_t)nslot * slab + 27ull * btile + 1024;
    // segment length: enough items for >= 4 waves when possible, at least 8 slabs per segment
    const int strips = N * p.HB * p.WB;
    int sd = D;
    while (sd > 8 && strips * cdiv(D, sd) * ns < 4 * num_sms()) sd = (sd + 1) / 2;
    p.SD = sd; p.DS = cdiv(D, sd);
    p.num_items = strips * p.DS;
    auto idesc = [](int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
    p.idesc = idesc(p.BN); p.idesc2 = idesc(2 * p.BN); p.idesc3 = idesc(3 * p.BN);
    p.merge = g_halo_merge && 3 * p.BN <= 256;   // (the two-wide edge merges need 2 * BN <= 256 only, implied)
    int nacc = 4;
    while (nacc * 2 <= HALO_MAX_ACC && nacc * 2 * p.BN <= 512) nacc *= 2;
    p.nacc = nacc;
    p.nacc_log2 = 0;
    while ((1 << p.nacc_log2) < nacc) ++p.nacc_log2;
    uint32_t cols = 32;
    while (cols < (uint32_t)(nacc * p.BN)) cols *= 2;
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
                        const float* bias, __nv_bfloat16* dst, int dst_pitch, int accumulate, cudaStream_t st, float* stat_part,
                        size_t stat_part_floats, int* stat_slots, int w_pitch, int w_row0) {
    // (w_pitch, w_row0): compute only the output-channel window [w_row0, w_row0 + Nout) of a weight matrix with w_pitch rows per
    // tap -- the caller offsets dst / bias itself.  Used to split the data gradient of a concat input into its two halves.
    if (w_pitch <= 0) { w_pitch = Nout; w_row0 = 0; }
    HaloParams p;
    size_t smem;
    if (!halo_plan(K, Nout, D, H, W, N, p, smem)) return fail(B2_EUNSUPPORTED, "conv_tc_halo: unsupported shape%s", "");
    B2_CHECK_ARG(src_pitch % 8 == 0 && dst_pitch % 8 == 0);
    p.dst_pitch = dst_pitch;
    p.w_pitch = w_pitch; p.w_row0 = w_row0;
    p.dbg = g_halo_dbg;
    CUtensorMap tmA, tmB;
    int rc = make_act_map(&tmA, src, N, D, H, W, K, src_pitch, K, 1, 1, 18, 8, 1, 1, 1);
    if (rc) return rc;
    rc = make_w_map(&tmB, wmat, 27 * w_pitch, K, K, p.BN);
    if (rc) return rc;
    int grid = p.num_items * p.NS < num_sms() ? p.num_items * p.NS : num_sms();
    grid = grid / p.NS * p.NS;   // every column split gets the same number of CTAs
    EpiStats es{nullptr, 0, Nout, N};
    if (stat_slots) *stat_slots = 0;
    if (stat_part && stat_slots && g_epi_stats) {
        es.slots = grid / p.NS * 8;   // one slot per epilogue warp (two sets of four)
        if ((size_t)N * es.slots * Nout * 2 <= stat_part_floats) { es.part = stat_part; *stat_slots = es.slots; }
    }
    static bool a32 = false, a64 = false;
    if (K == 32) {
        if (!a32) { B2_CUDA(cudaFuncSetAttribute(conv_halo_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); a32 = true; }
        B2_LAUNCH(conv_halo_kernel<32>, grid, HALO_THREADS, smem, st, tmA, tmB, p, bias, dst, accumulate, es);
    } else {
        if (!a64) { B2_CUDA(cudaFuncSetAttribute(conv_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); a64 = true; }
        B2_LAUNCH(conv_halo_kernel<64>, grid, HALO_THREADS, smem, st, tmA, tmB, p, bias, dst, accumulate, es);
    }
    return B2_OK;
}

}  // namespace b2
