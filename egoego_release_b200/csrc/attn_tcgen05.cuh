// Fused self-attention of one (window, head) on the tensor cores: S = Q K^T, row softmax over the first L
// keys, O = P V -- S and P never leave the SM.
//
// Every product uses the same 3-term fp16 hi/lo split as the GEMMs (fp32-grade results):
//   S = Qh Kh^T + Qh Kl^T + Ql Kh^T           (UMMA 128x128x16, K = d_k = 256)
//   O = Ph Vh   + Ph Vl   + Pl Vh             (UMMA 128x256x16, K = 128 keys)
// Operand planes (written by the QKV projection epilogue, fp16 hi/lo, K-major):
//   Q, K, V : [(window*H + head)*128 + token, 256]   (Q pre-scaled by 1/sqrt(d_k))
// V keeps the projection's natural [key][dim] layout: in O = P V it is the MN-major B operand of the UMMA (N = dim is the
// contiguous index), so the QKV epilogue writes all three sections the same coalesced way and no transposed copy exists.
// Persistent CTAs (one per SM, 192 threads) loop over (window, head) items:
//   warp 0    TMA producer, 128 KB ring (2 x 64 KB stages in split, 4 x 32 KB in fp16 format): 4 Q/K k-blocks then
//             2 V k-blocks (64 keys x 256 dims as four [64 keys][64 dims] boxes) per item
//   warp 1    MMA issuer (S of item i+1 is issued right after P V of item i, overlapping its epilogue)
//   warps 2-9 softmax (TMEM -> registers -> P planes into swizzled smem) and O epilogue (TMEM -> operand planes);
//             two warps share each TMEM lane quarter and split the key / output columns, exchanging the row
//             max / sum partials through shared memory
// TMEM: S at columns [0,128), O at [128,384).
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "gemm_tcgen05.cuh"

namespace egoego {

constexpr int ATT_RING_BYTES = 131072;                   // operand ring: 2 x 64 KB stages (split) or 4 x 32 KB (fp16)
constexpr int ATT_P_BYTES = 2 * 2 * 128 * 128;          // hi/lo x 2 key blocks x [128 rows x 128 B]
constexpr int ATT_PART_BYTES = 2 * 2 * 128 * 4;          // row max / row sum partials of the two column halves
constexpr int ATT_SMEM_BYTES = ATT_RING_BYTES + ATT_P_BYTES + ATT_PART_BYTES + 1024 + 256;
constexpr int ATT_THREADS = 64 + 256;                    // TMA warp, MMA warp, 8 softmax/epilogue warps

// Split format: the probabilities are handed to the tensor core as P * 2^11 (exact scaling), so that the lo plane of a
// typical p ~ 1/L is a NORMAL fp16 (22 significand bits for p >= 2^-14 instead of an absolute 3e-8); the O epilogue
// multiplies the accumulator by 2^-11 (exact).  p <= 1 keeps P * 2^11 <= 2048, far inside the fp16 range.
constexpr float ATT_P_SCALE = 2048.0f;

template <int FMT>
struct OStore : EpiNoDirect, EpiNoPre {    // attention output rows -> operand planes of the fc GEMM
    __nv_bfloat16* hi; __nv_bfloat16* lo; long long base; int ld;
    __device__ __forceinline__ float4 bias4(int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4, float4) const {
        const long long o = base + (long long)row * ld + col;
        if (FMT == FMT_SPLIT) { constexpr float r = 1.0f / ATT_P_SCALE; a.x *= r; a.y *= r; a.z *= r; a.w *= r; }
        store_planes4<FMT>(hi + o, lo + o, a);
    }
};

template <int FMT>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap mQh, const __grid_constant__ CUtensorMap mQl,
                    const __grid_constant__ CUtensorMap mKh, const __grid_constant__ CUtensorMap mKl,
                    const __grid_constant__ CUtensorMap mVh, const __grid_constant__ CUtensorMap mVl,
                    __nv_bfloat16* __restrict__ Ohi, __nv_bfloat16* __restrict__ Olo, int ldo,
                    int n_items, int n_head, int L) {
    constexpr int NP = FmtTraits<FMT>::NP;            // FMT_HALF: single fp16 plane per operand, one MMA per k-step
    static_assert(FMT == FMT_SPLIT || FMT == FMT_HALF, "attention operands are fp16 planes (hi/lo pair or single)");
    constexpr uint32_t IDESC_S = ptx::make_idesc_f16(128, 128);
    constexpr uint32_t IDESC_O = ptx::make_idesc_f16(128, 256) | ptx::IDESC_B_MN_MAJOR;
    constexpr uint32_t QK_BYTES = NP * 2 * 16384, V_BYTES = NP * 32768;
    constexpr int ATT_STAGE_BYTES = NP * 32768;            // one k-block of Q+K planes, or of V planes
    constexpr int ATT_STAGES = ATT_RING_BYTES / ATT_STAGE_BYTES;
    constexpr int OFF_QL = 16384, OFF_KH = NP * 16384, OFF_KL = NP * 16384 + 16384, OFF_VL = 32768;
    constexpr uint32_t TM_S = 0, TM_O = 128;
    // split format: the cross terms of S (Qh Kl^T + Ql Kh^T, 2^-11 of the magnitude) accumulate in their own TMEM columns and are
    // added in the softmax -- the tensor core truncates its fp32 accumulator after every MMA, and the hi*hi accumulator then sees a
    // third of the accumulations (same reasoning as the dual-accumulator GEMM, gemm_tcgen05.cuh)
    constexpr uint32_t TM_S2 = 384;

    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzle tiles, computed as an OFFSET on the __shared__ array so the compiler keeps
    // the shared address space (integer round-trips turn every access into a generic LD/ST).
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* p_smem = smem + ATT_RING_BYTES;           // P_hi [2][128][128B], P_lo [2][128][128B]
    float* part = reinterpret_cast<float*>(p_smem + ATT_P_BYTES);     // [max|sum][half][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(p_smem + ATT_P_BYTES + ATT_PART_BYTES);
    uint64_t* full_bar = bars;          // [ATT_STAGES <= 4]
    uint64_t* empty_bar = bars + 4;     // [ATT_STAGES <= 4]
    uint64_t* s_full = bars + 8;        // S accumulator ready (MMA -> softmax)
    uint64_t* p_ready = bars + 9;       // P written to smem and S drained (softmax -> MMA), 256 arrivals
    uint64_t* o_full = bars + 10;       // O accumulator ready (MMA -> epilogue)
    uint64_t* o_free = bars + 11;       // O drained (epilogue -> MMA), 256 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    // Role index, not hardware warp id: the SM sub-partition arbiter favours HIGHER warp ids, so the TMA producer (role 0)
    // and the MMA issuer (role 1) live in the two highest hardware warps and are never starved of issue slots by the
    // epilogue warps (roles 2..9 = hardware warps 0..7, whose id % 4 selects their TMEM lane quarter).
    const int lane = threadIdx.x % 32;
    const int warp = (threadIdx.x / 32 + 2) % (ATT_THREADS / 32);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mQh); ptx::prefetch_tmap(&mQl); ptx::prefetch_tmap(&mKh);
        ptx::prefetch_tmap(&mKl); ptx::prefetch_tmap(&mVh); ptx::prefetch_tmap(&mVl);
        for (int s = 0; s < ATT_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        ptx::mbar_init(s_full, 1); ptx::mbar_init(p_ready, 256); ptx::mbar_init(o_full, 1); ptx::mbar_init(o_free, 256);
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();                                // prologue above overlaps the previous kernel's tail (PDL)

    if (warp == 0) {
        if (lane == 0) {                                   // ===== TMA producer =====
            int s = 0; uint32_t ph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                for (int kb = 0; kb < 4; ++kb) {           // Q/K k-blocks of 64 dims: Qh Ql Kh Kl (16 KB each)
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * ATT_STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&full_bar[s], QK_BYTES);
                    ptx::tma_load_2d(st, &mQh, &full_bar[s], kb * 64, item * 128);
                    if (NP == 2) ptx::tma_load_2d(st + OFF_QL, &mQl, &full_bar[s], kb * 64, item * 128);
                    ptx::tma_load_2d(st + OFF_KH, &mKh, &full_bar[s], kb * 64, item * 128);
                    if (NP == 2) ptx::tma_load_2d(st + OFF_KL, &mKl, &full_bar[s], kb * 64, item * 128);
                    if (++s == ATT_STAGES) { s = 0; ph ^= 1; }
                }
                for (int kb = 0; kb < 2; ++kb) {           // V k-blocks of 64 keys: Vh Vl (32 KB each = 4 dim blocks x [64 keys][128 B])
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * ATT_STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&full_bar[s], V_BYTES);
#pragma unroll
                    for (int db = 0; db < 4; ++db) {
                        ptx::tma_load_2d(st + db * 8192, &mVh, &full_bar[s], db * 64, item * 128 + kb * 64);
                        if (NP == 2) ptx::tma_load_2d(st + OFF_VL + db * 8192, &mVl, &full_bar[s], db * 64, item * 128 + kb * 64);
                    }
                    if (++s == ATT_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ===== MMA issuer =====
            int s = 0; uint32_t ph = 0; int it = 0;
            const uint32_t p_hi = ptx::smem_u32(p_smem), p_lo = p_hi + 2 * 16384;
            auto issue_S = [&]() {
                for (int kb = 0; kb < 4; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + s * ATT_STAGE_BYTES);
                    const uint64_t dQh = ptx::make_smem_desc_sw128(st), dQl = ptx::make_smem_desc_sw128(st + OFF_QL);
                    const uint64_t dKh = ptx::make_smem_desc_sw128(st + OFF_KH), dKl = ptx::make_smem_desc_sw128(st + OFF_KL);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t adv = (uint64_t)(kk * 2);
                        ptx::umma_f16(tmem_base + TM_S, dQh + adv, dKh + adv, IDESC_S, (kb | kk) != 0);
                        if (NP == 2) {
                            ptx::umma_f16(tmem_base + TM_S2, dQh + adv, dKl + adv, IDESC_S, (kb | kk) != 0);
                            ptx::umma_f16(tmem_base + TM_S2, dQl + adv, dKh + adv, IDESC_S, 1);
                        }
                    }
                    ptx::umma_commit(&empty_bar[s]);
                    if (++s == ATT_STAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit(s_full);
            };
            // Stream order of the ring is QK(i) V(i) QK(i+1) V(i+1) ..., so S(i+1) can only be issued after P V(i).
            if (blockIdx.x < n_items) issue_S();
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const uint32_t iph = it & 1;
                ptx::mbar_wait(p_ready, iph);              // P(i) in smem, S drained
                ptx::mbar_wait(o_free, iph ^ 1);           // O of the previous item drained
                ptx::tc_fence_after();
                for (int kb = 0; kb < 2; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + s * ATT_STAGE_BYTES);
                    // V: MN-major B, dim blocks 8 KB apart (LBO), 8-key groups 1 KB apart (SBO); 16 keys per UMMA = 2 KB
                    const uint64_t dVh = ptx::make_smem_desc_mn_sw128(st, 8192, 1024), dVl = ptx::make_smem_desc_mn_sw128(st + OFF_VL, 8192, 1024);
                    const uint64_t dPh = ptx::make_smem_desc_sw128(p_hi + kb * 16384), dPl = ptx::make_smem_desc_sw128(p_lo + kb * 16384);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t adv = (uint64_t)(kk * 2), advv = (uint64_t)(kk * 128);
                        ptx::umma_f16(tmem_base + TM_O, dPh + adv, dVh + advv, IDESC_O, (kb | kk) != 0);
                        if (NP == 2) {
                            ptx::umma_f16(tmem_base + TM_O, dPh + adv, dVl + advv, IDESC_O, 1);
                            ptx::umma_f16(tmem_base + TM_O, dPl + adv, dVh + advv, IDESC_O, 1);
                        }
                    }
                    ptx::umma_commit(&empty_bar[s]);
                    if (++s == ATT_STAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit(o_full);
                if (item + (int)gridDim.x < n_items) issue_S();
            }
        }
    } else {                                               // ===== softmax + epilogue warps 2..9 =====
        const int quarter = (warp - 2) & 3, hf = (warp - 2) >> 2;  // TMEM lane quarter, column half
        const int r = quarter * 32 + lane;                 // query row = TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            const uint32_t iph = it & 1;
            // ---- softmax over keys [0, L): this warp owns keys [64*hf, 64*hf + 64) of its 32 rows ----
            ptx::mbar_wait(s_full, iph);
            ptx::tc_fence_after();
            float v[64];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(lane_addr + TM_S + hf * 64 + c * 32, raw);
                if (NP == 2) {
                    uint32_t raw2[32];
                    ptx::tmem_ld_32x32(lane_addr + TM_S2 + hf * 64 + c * 32, raw2);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(raw[j]) + __uint_as_float(raw2[j]);
                } else {
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(raw[j]);
                }
            }
            const int k0 = hf * 64;
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 64; ++k) { if (k0 + k >= L) v[k] = -INFINITY; mx = fmaxf(mx, v[k]); }
            part[hf * 128 + r] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            mx = fmaxf(part[r], part[128 + r]);            // key 0 is always valid, so the row max is finite
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                const float e = (FMT == FMT_SPLIT) ? expf(v[k] - mx) : __expf(v[k] - mx);
                v[k] = (k0 + k < L) ? e : 0.f;
                sum += v[k];
            }
            part[256 + hf * 128 + r] = sum;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            const float inv = ((FMT == FMT_SPLIT) ? ATT_P_SCALE : 1.0f) / (part[256 + r] + part[256 + 128 + r]);
            // P planes -> K-major SW128 smem: key block hf, 16-byte chunk j (keys 8j..8j+7) of row r at chunk j^(r%8)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t ph4[4], pl4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (FMT == FMT_SPLIT) {
                        __nv_bfloat16 h0, l0, h1, l1;
                        split_f16(v[8 * j + 2 * q] * inv, h0, l0);
                        split_f16(v[8 * j + 2 * q + 1] * inv, h1, l1);
                        __nv_bfloat162 hh(h0, h1), ll(l0, l1);
                        ph4[q] = *reinterpret_cast<uint32_t*>(&hh); pl4[q] = *reinterpret_cast<uint32_t*>(&ll);
                    } else {
                        __half2 hh = __floats2half2_rn(v[8 * j + 2 * q] * inv, v[8 * j + 2 * q + 1] * inv);
                        ph4[q] = *reinterpret_cast<uint32_t*>(&hh); pl4[q] = 0u;
                    }
                }
                uint8_t* dst = p_smem + hf * 16384 + r * 128 + ((j ^ (r & 7)) * 16);
                *reinterpret_cast<uint4*>(dst) = make_uint4(ph4[0], ph4[1], ph4[2], ph4[3]);
                if (FMT == FMT_SPLIT) *reinterpret_cast<uint4*>(dst + 2 * 16384) = make_uint4(pl4[0], pl4[1], pl4[2], pl4[3]);
            }
            ptx::fence_proxy_async();                      // generic-proxy smem writes -> visible to the tensor core
            ptx::tc_fence_before();
            ptx::mbar_arrive(p_ready);
            // ---- O epilogue: this warp drains output columns [128*hf, 128*hf + 128) of its 32 rows ----
            ptx::mbar_wait(o_full, iph);
            ptx::tc_fence_after();
            // P(i) has been consumed once o_full fired.  Each warp reuses exactly ITS OWN 4 KB of the P buffer (the 32
            // rows x 64 keys it wrote) as the transpose tile of the coalesced O store, so no other warp's softmax of
            // item i+1 can overwrite a tile that is still in use.
            const int w = item / n_head, h = item % n_head;
            float4* tile = reinterpret_cast<float4*>(p_smem + hf * 16384 + quarter * 4096);
            OStore<FMT> ost{{}, {}, Ohi, Olo, ((long long)w * LP) * ldo + h * 256, ldo};
            epilogue_drain<128>(ost, tile, lane_addr + TM_O + hf * 128, lane, quarter * 32, hf * 128, []() {});
            ptx::tc_fence_before();
            ptx::mbar_arrive(o_free);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

// ---- fp16-format attention, software-pipelined -------------------------------------------------------------
// ncu on attention_tc_kernel<FMT_HALF> (8.4 us per item): softmax 48 %, transposing O epilogue 34 %, waiting for P V 16 %,
// waiting for S 12 % -- the eight softmax/epilogue warps are the critical path and sit idle while the tensor core runs.
// This kernel keeps the same per-item math but
//   * double-buffers S in TMEM ([0,128) / [384,512)) and P in shared memory, so softmax(i+1) runs while P V(i) executes:
//       MMA order  S(0) S(1) | PV(0) S(2) | PV(1) S(3) ...      (TMA ring order QK0 QK1 | V0 QK2 | V1 QK3 ...)
//       warps      softmax(0) | softmax(1) Oepi(0) | softmax(2) Oepi(1) ...
//   * softmax: invalid keys are masked per group of 8 (warp-uniform), exp via ex2.approx on packed fp32x2 arguments
//     (a masked -inf logit gives exp = 0 without a select), tree reductions;
//   * O epilogue in the accumulator's thread-per-row layout: fp16 rows into a 128B-swizzled staging box (one 16 KB box per
//     column half, used twice per item), TMA-stored by an I/O warp -- no smem transposes, no STG.
constexpr int ATT2_THREADS = ATT_THREADS + 32;           // + store I/O warp (role 10)
constexpr int ATT2_SMEM_BYTES = ATT_RING_BYTES + 2 * 32768 /*P x2*/ + 32768 /*O staging*/ + ATT_PART_BYTES + 256;

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(ATT2_THREADS, 1)
attention_half_kernel(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK,
                      const __grid_constant__ CUtensorMap mV /* 64-row boxes */, const __grid_constant__ CUtensorMap mO /* [M, H*256] fp16, 128-row x 64-col boxes */,
                      int n_items, int n_head, int L, int rev /* 1: walk the items from the last window to the first (L2 zig-zag) */,
                      const int* __restrict__ win_cnt /* nullptr, or: streamed mode -- this grid runs CONCURRENTLY with the QKV projection that
                                                         feeds it (gemm_half_tma_2cta_kernel with row_cnt = win_cnt) and loads window w once
                                                         win_cnt[w] >= tiles_per_window * (*d_step + 1) */,
                      const int* __restrict__ d_step, int tiles_per_window,
                      int early_ctas, int early_items /* streamed mode, two phases: CTAs [0, early_ctas) -- resident next to the producer grid --
                                                         share items [0, early_items) and follow the counters; the other CTAs become resident
                                                         as the producer's CTAs exit and share the remaining items (already produced, still in
                                                         the L2).  early_ctas = 0: every CTA strides over all items */) {
    constexpr uint32_t IDESC_S = ptx::make_idesc_f16(128, 128);
    constexpr uint32_t IDESC_O = ptx::make_idesc_f16(128, 256) | ptx::IDESC_B_MN_MAJOR;
    constexpr int STAGE = 32768, STAGES = ATT_RING_BYTES / STAGE;           // 4
    constexpr uint32_t TM_O = 128;
    extern __shared__ __align__(1024) uint8_t smem_att2[];
    uint8_t* smem = smem_att2;
    uint8_t* p_smem = smem + ATT_RING_BYTES;             // P[2]: [2 key blocks][128 rows][128 B] each
    uint8_t* o_smem = p_smem + 2 * 32768;                // O staging: one [128 rows][64 cols] fp16 box per column half
    float* part = reinterpret_cast<float*>(o_smem + 32768);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(part) + ATT_PART_BYTES);
    uint64_t* full_bar = bars;          // [4]
    uint64_t* empty_bar = bars + 4;     // [4]
    uint64_t* s_full = bars + 8;        // [2] S buffer ready (MMA -> softmax)
    uint64_t* p_ready = bars + 10;      // [2] P buffer written and S buffer drained (softmax -> MMA), 256 arrivals
    uint64_t* o_full = bars + 12;       // O accumulator ready (MMA -> epilogue)
    uint64_t* o_free = bars + 13;       // O drained (epilogue -> MMA), 8 arrivals
    uint64_t* out_ready = bars + 14;    // [2 column halves] staging box written by its 4 warps
    uint64_t* box_free = bars + 16;     // [2] store has finished reading the box
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int lane = threadIdx.x % 32;
    const int hw_warp = threadIdx.x / 32;               // hardware warps 0..7 = softmax/epilogue roles 2..9; 8, 9, 10 = roles 0, 1, 10
    const int warp = hw_warp < 8 ? hw_warp + 2 : (hw_warp == 10 ? 10 : hw_warp - 8);
    if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();   // swizzled tiles need the declared 1024-byte alignment

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mQ); ptx::prefetch_tmap(&mK); ptx::prefetch_tmap(&mV); ptx::prefetch_tmap(&mO);
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&s_full[b], 1); ptx::mbar_init(&p_ready[b], 256);
            ptx::mbar_init(&out_ready[b], 4); ptx::mbar_init(&box_free[b], 1);
        }
        ptx::mbar_init(o_full, 1); ptx::mbar_init(o_free, 8);
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch();
    // Streamed mode: no wait for the producer grid.  Its CTAs release this grid only after their own griddepcontrol.wait, so all
    // earlier kernels have completed (their reads of O, their write of *d_step) when the first thread gets here.
    if (!win_cnt) ptx::grid_dep_wait();                  // prologue above overlaps the previous kernel's tail (PDL)
    if (threadIdx.x == 0) timeline_mark(1, 0);
    // items of this CTA: first + j * stride < last
    int first = (int)blockIdx.x, stride = (int)gridDim.x, last = n_items;
    if (early_ctas > 0) {
        if ((int)blockIdx.x < early_ctas) { stride = early_ctas; last = early_items; }
        else { first = early_items + (int)blockIdx.x - early_ctas; stride = (int)gridDim.x - early_ctas; }
    }
    const int n_mine = first < last ? (last - 1 - first) / stride + 1 : 0;
    auto item_of = [&](int j) { const int i = first + j * stride; return rev ? n_items - 1 - i : i; };

    if (warp == 0) {
        if (lane == 0) {                                   // ===== TMA producer: QK(0) QK(1) | V(0) QK(2) | V(1) QK(3) ... =====
            int s = 0; uint32_t ph = 0;
            const int cnt_target = win_cnt ? tiles_per_window * ((d_step ? *d_step : 0) + 1) : 0;
            auto load_qk = [&](int item) {
                if (win_cnt) {                             // Q, K and V of this window are performed in global memory?
                    const int* c = win_cnt + item / n_head;
                    while (ptx::ld_acquire_gpu(c) < cnt_target) __nanosleep(64);
                }
                for (int kb = 0; kb < 4; ++kb) {
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * STAGE;
                    ptx::mbar_arrive_expect_tx(&full_bar[s], STAGE);
                    ptx::tma_load_2d(st, &mQ, &full_bar[s], kb * 64, item * 128);
                    ptx::tma_load_2d(st + 16384, &mK, &full_bar[s], kb * 64, item * 128);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            };
            auto load_v = [&](int item) {
                for (int kb = 0; kb < 2; ++kb) {
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * STAGE;
                    ptx::mbar_arrive_expect_tx(&full_bar[s], STAGE);
#pragma unroll
                    for (int db = 0; db < 4; ++db) ptx::tma_load_2d(st + db * 8192, &mV, &full_bar[s], db * 64, item * 128 + kb * 64);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            };
            if (n_mine > 0) load_qk(item_of(0));
            if (n_mine > 1) load_qk(item_of(1));
            for (int j = 0; j < n_mine; ++j) {
                load_v(item_of(j));
                if (j + 2 < n_mine) load_qk(item_of(j + 2));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ===== MMA issuer =====
            int s = 0; uint32_t ph = 0;
            const uint32_t p_addr = ptx::smem_u32(p_smem);
            auto issue_S = [&](int j) {
                const uint32_t tm_s = tmem_base + ((j & 1) ? 384u : 0u);
                for (int kb = 0; kb < 4; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + s * STAGE);
                    const uint64_t dQ = ptx::make_smem_desc_sw128(st), dK = ptx::make_smem_desc_sw128(st + 16384);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        ptx::umma_f16(tm_s, dQ + (uint64_t)(kk * 2), dK + (uint64_t)(kk * 2), IDESC_S, (kb | kk) != 0);
                    ptx::umma_commit(&empty_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit(&s_full[j & 1]);
            };
            if (n_mine > 0) issue_S(0);
            if (n_mine > 1) issue_S(1);
            for (int j = 0; j < n_mine; ++j) {
                ptx::mbar_wait(&p_ready[j & 1], (j >> 1) & 1);   // P(j) in smem, S buffer j&1 drained
                ptx::mbar_wait(o_free, (j & 1) ^ 1);             // O of item j-1 drained
                ptx::tc_fence_after();
                for (int kb = 0; kb < 2; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + s * STAGE);
                    const uint64_t dV = ptx::make_smem_desc_mn_sw128(st, 8192, 1024);
                    const uint64_t dP = ptx::make_smem_desc_sw128(p_addr + (j & 1) * 32768 + kb * 16384);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        ptx::umma_f16(tmem_base + TM_O, dP + (uint64_t)(kk * 2), dV + (uint64_t)(kk * 128), IDESC_O, (kb | kk) != 0);
                    ptx::umma_commit(&empty_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit(o_full);
                if (j + 2 < n_mine) issue_S(j + 2);
            }
        }
    } else if (warp == 10) {
        if (lane == 0) {                                   // ===== store I/O =====
            uint32_t u = 0;
            for (int j = 0; j < n_mine; ++j) {
                const int item = item_of(j), w = item / n_head, h = item % n_head;
#pragma unroll 1
                for (int half = 0; half < 2; ++half, ++u) {
#pragma unroll 1
                    for (int hf = 0; hf < 2; ++hf) {
                        ptx::mbar_wait(&out_ready[hf], u & 1);
                        ptx::tma_store_2d(&mO, o_smem + hf * 16384, h * 256 + hf * 128 + half * 64, w * LP);
                        ptx::tma_store_commit();
                        ptx::tma_store_wait_read();
                        ptx::mbar_arrive(&box_free[hf]);
                    }
                }
            }
            ptx::tma_store_wait_all();
            timeline_mark(1, 1);
        }
    } else {                                               // ===== softmax + epilogue warps 2..9 =====
        const int quarter = (warp - 2) & 3, hf = (warp - 2) >> 2;  // TMEM lane quarter, key / output-column half
        const int r = quarter * 32 + lane;                 // query row = TMEM lane
        const int sw = r & 7;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const int k0 = hf * 64;
        const int nvalid = L - k0 < 0 ? 0 : (L - k0 > 64 ? 64 : L - k0);     // valid keys of this warp's 64 (warp-uniform)
        const float kLog2e = 1.4426950408889634f;
        auto softmax = [&](int j) {
            const int b = j & 1;
            ptx::mbar_wait(&s_full[b], (j >> 1) & 1);
            ptx::tc_fence_after();
            float2 v[32];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(lane_addr + (b ? 384u : 0u) + hf * 64 + c * 32, raw);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; ++q) v[c * 16 + q] = make_float2(__uint_as_float(raw[2 * q]), __uint_as_float(raw[2 * q + 1]));
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {                  // mask invalid keys, one warp-uniform decision per group of 8
                if (g * 8 + 8 > nvalid) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (g * 8 + 2 * q >= nvalid) v[g * 4 + q].x = -INFINITY;
                        if (g * 8 + 2 * q + 1 >= nvalid) v[g * 4 + q].y = -INFINITY;
                    }
                }
            }
            float m8[8];
#pragma unroll
            for (int g = 0; g < 8; ++g)
                m8[g] = fmaxf(fmaxf(fmaxf(v[g * 4].x, v[g * 4].y), fmaxf(v[g * 4 + 1].x, v[g * 4 + 1].y)),
                              fmaxf(fmaxf(v[g * 4 + 2].x, v[g * 4 + 2].y), fmaxf(v[g * 4 + 3].x, v[g * 4 + 3].y)));
            float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
            part[hf * 128 + r] = mx;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            mx = fmaxf(part[r], part[128 + r]);            // key 0 is always valid, so the row max is finite
            const float2 l2 = make_float2(kLog2e, kLog2e), c2 = make_float2(-mx * kLog2e, -mx * kLog2e);
            float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const float2 a = __ffma2_rn(v[q], l2, c2);                  // (s - max) * log2(e); -inf stays -inf -> exp = 0
                v[q] = make_float2(ex2_approx(a.x), ex2_approx(a.y));
                acc[q & 3] = __fadd2_rn(acc[q & 3], v[q]);
            }
            const float2 t2 = __fadd2_rn(__fadd2_rn(acc[0], acc[1]), __fadd2_rn(acc[2], acc[3]));
            part[256 + hf * 128 + r] = t2.x + t2.y;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
            const float inv = 1.0f / (part[256 + r] + part[256 + 128 + r]);
            const float2 i2 = make_float2(inv, inv);
            uint8_t* prow = p_smem + b * 32768 + hf * 16384 + r * 128;      // key block hf of P[b], row r
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                uint32_t hw[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 pv = __fmul2_rn(v[jj * 4 + q], i2);
                    const __half2 hh = __floats2half2_rn(pv.x, pv.y);
                    hw[q] = *reinterpret_cast<const uint32_t*>(&hh);
                }
                *reinterpret_cast<uint4*>(prow + ((jj ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            }
            ptx::fence_proxy_async();                      // generic-proxy smem writes -> visible to the tensor core
            ptx::tc_fence_before();
            ptx::mbar_arrive(&p_ready[b]);
        };
        uint32_t u = 0;
        uint8_t* orow = o_smem + hf * 16384 + r * 128;
        if (n_mine > 0) softmax(0);
        for (int j = 0; j < n_mine; ++j) {
            if (j + 1 < n_mine) softmax(j + 1);            // overlaps P V(j) on the tensor core
            ptx::mbar_wait(o_full, j & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {                  // this warp's 128 output columns, 32 at a time
                uint32_t raw[32];
                ptx::tmem_ld_32x32(lane_addr + TM_O + hf * 128 + c * 32, raw);
                if ((c & 1) == 0) ptx::mbar_wait(&box_free[hf], (u & 1) ^ 1);
                ptx::tmem_ld_wait();
                if (c == 3) {                              // O drained: the next P V may overwrite it
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(o_free);
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    uint32_t hw[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const __half2 hh = __floats2half2_rn(__uint_as_float(raw[jj * 8 + 2 * q]), __uint_as_float(raw[jj * 8 + 2 * q + 1]));
                        hw[q] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                    *reinterpret_cast<uint4*>(orow + ((((c & 1) * 4 + jj) ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                }
                if (c & 1) {
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&out_ready[hf]);
                    ++u;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

// QKV projection epilogue writing the attention operand planes directly (one 256-wide tile = one (section, head));
// all three sections share the layout [(w*H+h)*128 + l][256].
template <int FMT>
struct TcEpiQKVPlanes : EpiNoDirect, EpiNoPre {
    __nv_bfloat16 *Qh, *Ql, *Kh, *Kl, *Vh, *Vl;
    const float* bias; int n_head; float q_scale;
    __device__ __forceinline__ float4 bias4(int col) const { return ld4(bias + col); }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4 b, float4) const {
        const int w = row / LP, l = row % LP;
        const int hw = n_head * 256;
        const int sec = col / hw, hc = col - sec * hw;
        const int h = hc >> 8, c = hc & 255;
        a = add4(a, b);
        if (sec == 0) a = make_float4(a.x * q_scale, a.y * q_scale, a.z * q_scale, a.w * q_scale);
        const long long o = ((long long)(w * n_head + h) * 128 + l) * 256 + c;
        store_planes4<FMT>((sec == 0 ? Qh : (sec == 1 ? Kh : Vh)) + o, (sec == 0 ? Ql : (sec == 1 ? Kl : Vl)) + o, a);
    }
};

}  // namespace egoego
