// Fused self-attention of one (window, head) on the tensor cores: S = Q K^T, row softmax over the first L
// keys, O = P V -- S and P never leave the SM.
//
// Every product uses the same 3-term bf16 hi/lo split as the GEMMs (fp32-grade results):
//   S = Qh Kh^T + Qh Kl^T + Ql Kh^T           (UMMA 128x128x16, K = d_k = 256)
//   O = Ph Vh   + Ph Vl   + Pl Vh             (UMMA 128x256x16, K = 128 keys)
// Operand planes (written by the QKV projection epilogue, bf16, K-major):
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

template <int FMT>
struct OStore : EpiNoDirect, EpiNoPre {    // attention output rows -> operand planes of the fc GEMM
    __nv_bfloat16* hi; __nv_bfloat16* lo; long long base; int ld;
    __device__ __forceinline__ float4 bias4(int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4, float4) const {
        const long long o = base + (long long)row * ld + col;
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
    constexpr uint32_t IDESC_S = (FMT == FMT_SPLIT) ? ptx::make_idesc_bf16(128, 128) : ptx::make_idesc_f16(128, 128);
    constexpr uint32_t IDESC_O = ((FMT == FMT_SPLIT) ? ptx::make_idesc_bf16(128, 256) : ptx::make_idesc_f16(128, 256)) | ptx::IDESC_B_MN_MAJOR;
    constexpr uint32_t QK_BYTES = NP * 2 * 16384, V_BYTES = NP * 32768;
    constexpr int ATT_STAGE_BYTES = NP * 32768;            // one k-block of Q+K planes, or of V planes
    constexpr int ATT_STAGES = ATT_RING_BYTES / ATT_STAGE_BYTES;
    constexpr int OFF_QL = 16384, OFF_KH = NP * 16384, OFF_KL = NP * 16384 + 16384, OFF_VL = 32768;
    constexpr uint32_t TM_S = 0, TM_O = 128;

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
                            ptx::umma_f16(tmem_base + TM_S, dQh + adv, dKl + adv, IDESC_S, 1);
                            ptx::umma_f16(tmem_base + TM_S, dQl + adv, dKh + adv, IDESC_S, 1);
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
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(raw[j]);
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
            const float inv = 1.0f / (part[256 + r] + part[256 + 128 + r]);
            // P planes -> K-major SW128 smem: key block hf, 16-byte chunk j (keys 8j..8j+7) of row r at chunk j^(r%8)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t ph4[4], pl4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (FMT == FMT_SPLIT) {
                        __nv_bfloat16 h0, l0, h1, l1;
                        split_bf16(v[8 * j + 2 * q] * inv, h0, l0);
                        split_bf16(v[8 * j + 2 * q + 1] * inv, h1, l1);
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
