// Split-precision GEMM on the 5th-generation tensor cores.
//
//   D[M,N] = A[M,K] * W[N,K]^T   with   A = A_hi + A_lo,  W = W_hi + W_lo  (fp16 planes in the sampler, bf16 in the training step; K-major)
//   D ~= A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T     (three tcgen05.mma per k-step, one fp32 accumulator)
//
// which reproduces fp32 GEMM results to ~3e-6 of max|D| with fp16 planes (bounded by the tensor core's own fp32
// accumulation; ~2^-16 relative with bf16 planes).  SURVEY.md 0.5 / 8d: single-pass bf16/fp16/tf32 on every step
// miss the 1e-3 m parity bar; the 3-term split meets it with 30x margin on the 1000-step golden.  The four operand planes of a
// k-block share one pipeline stage, so the split moves 4 tiles per 3 MMAs (better bytes/FLOP than a plain
// bf16 GEMM).
//
// Structure (persistent, one CTA per SM, 192 threads):
//   warp 0    TMA producer: cp.async.bulk.tensor 2D, 128B swizzle, mbarrier complete_tx
//   warp 1    TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma / tcgen05.commit)
//   warps 2-9 epilogue: tcgen05.ld 32x32b -> registers -> smem transpose -> fused, coalesced epilogue -> global
// Accumulators are double-buffered in TMEM (2 x BN columns of the 512) so the epilogue of tile i overlaps
// the main loop of tile i+1.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace egoego {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;            // 64 16-bit elements = 128 B = one swizzle row
constexpr int GEMM_EPI_WARPS = 8;                  // two warps per TMEM lane quarter, each takes half of the tile columns
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;

// Operand format of a GEMM / attention launch:
//   FMT_SPLIT       fp16 hi/lo planes (split_f16), three MMAs per k-step (fp32-grade; the default everywhere)
//   FMT_HALF        one fp16 plane, one MMA per k-step (used only for the early, high-noise diffusion steps whose
//                   error is damped by posterior_mean_coef1 -- see DESIGN.md "Precision policy")
//   FMT_SPLIT_BF16  bf16 hi/lo planes (split_bf16), three MMAs per k-step: the training step's products, whose
//                   gradients need the fp32 exponent range (tc_gemm_f32); same kernels, bf16 instruction descriptor
enum { FMT_SPLIT = 0, FMT_HALF = 1, FMT_SPLIT_BF16 = 2 };
__host__ __device__ constexpr bool fmt_is_split(int fmt) { return fmt != FMT_HALF; }
template <int FMT> struct FmtTraits {
    static constexpr int NP = fmt_is_split(FMT) ? 2 : 1;                        // operand planes per matrix
    static constexpr int MMAS = fmt_is_split(FMT) ? 3 : 1;                      // MMAs per k-step
};

template <int BN, int FMT> struct GemmCfg {
    static constexpr int NP = FmtTraits<FMT>::NP;
    static constexpr int STAGE_BYTES = NP * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2);   // A planes then W planes
    static constexpr int STAGES = fmt_is_split(FMT) ? 2 : 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_EPI_WARPS * 4096 /*epilogue tiles*/ + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ---- epilogues -----------------------------------------------------------------------------------------
// tcgen05.ld hands each thread one accumulator ROW (32 consecutive columns).  Writing global memory in that layout
// touches 32 cache lines per warp instruction (measured: 32 sectors/request, LSU-bound epilogue), so each epilogue
// warp first transposes its 32x32 chunk through a private 4 KB shared-memory tile (XOR-swizzled, conflict-free) and
// then runs the fused epilogue in a COALESCED layout: a lane owns 4 consecutive columns of one row, 8 lanes cover a
// 128-byte row segment, one warp instruction touches 4 full lines.  Functors implement
//     apply4(row, col, float4 acc)          coalesced element-wise epilogue (bias/residual/activation/split/store)
//     direct(col0) / apply_row(...)         optional thread-per-row path for TRANSPOSED destinations (V^T), where
//                                           the thread=row layout is already the coalesced one.
constexpr int EPI_TILE_BYTES = 32 * 32 * 4;              // per epilogue warp

// Per-warp prefetch of the functor's per-element read-only operand (residual / base) for one 32x32 chunk, in the
// coalesced layout; issued one chunk ahead (and before the accumulator-ready wait for chunk 0) to hide HBM latency.
template <class Epi>
__device__ __forceinline__ void epilogue_prefetch(const Epi& epi, float4 (&pre)[8], int lane, int row_base, int col0) {
#pragma unroll
    for (int it = 0; it < 8; ++it) pre[it] = epi.pre(row_base + it * 4 + (lane >> 3), col0 + 4 * (lane & 7));
}

// `bias_lane` holds the bias of the warp's whole 128-column range (lane l: columns 4l..4l+3); the values a lane needs
// for chunk `cidx` are fetched with warp shuffles, so the epilogue issues no per-chunk global loads for the bias.
template <class Epi>
__device__ __forceinline__ void epilogue_chunk(const Epi& epi, float4* tile, const uint32_t (&raw)[32], int lane,
                                               int row_base /* first row of this warp's 32 */, int col0, int cidx,
                                               float4 bias_lane, const float4 (&pre)[8]) {
    if (epi.direct(col0)) {                                // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float comp = (j & 3) == 0 ? bias_lane.x : ((j & 3) == 1 ? bias_lane.y : ((j & 3) == 2 ? bias_lane.z : bias_lane.w));
            v[j] = __uint_as_float(raw[j]) + __shfl_sync(0xffffffffu, comp, cidx * 8 + (j >> 2));
        }
        epi.apply_row(row_base + lane, col0, v);
        return;
    }
    const int src = cidx * 8 + (lane & 7);
    float4 b4;
    b4.x = __shfl_sync(0xffffffffu, bias_lane.x, src); b4.y = __shfl_sync(0xffffffffu, bias_lane.y, src);
    b4.z = __shfl_sync(0xffffffffu, bias_lane.z, src); b4.w = __shfl_sync(0xffffffffu, bias_lane.w, src);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        tile[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                                        __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
    __syncwarp();
    const int j = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + (lane >> 3);
        const float4 acc = tile[r * 8 + (j ^ (r & 7))];
        epi.apply4(row_base + r, col0 + 4 * j, acc, b4, pre[it]);
    }
    __syncwarp();
}

// Drain this warp's 32 rows x NCOLS (=128) columns of an accumulator: bias in registers, operand prefetch one chunk ahead.
// `wait_ready()` blocks until the accumulator may be read (called after the first prefetch has been issued).
// `taddr2` != 0: a second accumulator of the same shape is added to the first (fp32, round to nearest) before the epilogue runs --
// the cross-term accumulator of the dual-accumulator split GEMM.
template <int NCOLS, class Epi, class Wait>
__device__ __forceinline__ void epilogue_drain(const Epi& epi, float4* tile, uint32_t taddr, int lane, int row_base, int col_begin,
                                               Wait wait_ready, uint32_t taddr2 = 0) {
    static_assert(NCOLS == 128, "bias_lane covers exactly 32 lanes x 4 columns");
    const float4 bias_lane = epi.bias4(col_begin + 4 * lane);
    float4 pre[8];
    epilogue_prefetch(epi, pre, lane, row_base, col_begin);
    wait_ready();
#pragma unroll 1
    for (int c = 0; c < NCOLS / 32; ++c) {
        float4 cur[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) cur[it] = pre[it];
        if (c + 1 < NCOLS / 32) epilogue_prefetch(epi, pre, lane, row_base, col_begin + (c + 1) * 32);
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr + c * 32, r);
        if (taddr2) {
            uint32_t r2[32];
            ptx::tmem_ld_32x32(taddr2 + c * 32, r2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + __uint_as_float(r2[i]));
        } else {
            ptx::tmem_ld_wait();
        }
        epilogue_chunk(epi, tile, r, lane, row_base, col_begin + c * 32, c, bias_lane, cur);
    }
}

struct EpiNoDirect {
    __device__ __forceinline__ bool direct(int) const { return false; }
    __device__ __forceinline__ void apply_row(int, int, const float (&)[32]) const {}
};
struct EpiNoPre {
    __device__ __forceinline__ float4 pre(int, int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
};

__device__ __forceinline__ void store_half8(__nv_bfloat16* dst /* fp16 bits */, const float* v) {
    uint32_t p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __half2 h = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        p[q] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(p[0], p[1], p[2], p[3]);
}

__device__ __forceinline__ void store_split8(__nv_bfloat16* hi, __nv_bfloat16* lo, const float* v) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_f16(v[2 * q], h0, l0); split_f16(v[2 * q + 1], h1, l1);
        __nv_bfloat162 hh(h0, h1), ll(l0, l1);
        ph[q] = *reinterpret_cast<uint32_t*>(&hh); pl[q] = *reinterpret_cast<uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(hi) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(lo) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

template <int FMT>
__device__ __forceinline__ void store_planes8(__nv_bfloat16* hi, __nv_bfloat16* lo, const float* v) {
    static_assert(FMT != FMT_SPLIT_BF16, "operand planes written by epilogues are fp16 (sampling path)");
    if (FMT == FMT_SPLIT) store_split8(hi, lo, v); else store_half8(hi, v);
}

// 4 consecutive values -> operand plane(s): 8 bytes per plane
template <int FMT>
__device__ __forceinline__ void store_planes4(__nv_bfloat16* hi, __nv_bfloat16* lo, float4 v) {
    static_assert(FMT != FMT_SPLIT_BF16, "operand planes written by epilogues are fp16 (sampling path)");
    if (FMT == FMT_SPLIT) {
        __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
        split_f16(v.x, h0, l0); split_f16(v.y, h1, l1); split_f16(v.z, h2, l2); split_f16(v.w, h3, l3);
        __nv_bfloat162 a0(h0, h1), a1(h2, h3), b0(l0, l1), b1(l2, l3);
        *reinterpret_cast<uint2*>(hi) = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
        *reinterpret_cast<uint2*>(lo) = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
    } else {
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        *reinterpret_cast<uint2*>(hi) = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
}

// read-only inputs of an epilogue (bias, residual, tables) go through the non-coherent path so the compiler may
// hoist all loads of a chunk above its stores
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

struct TcEpiPlain : EpiNoDirect, EpiNoPre {   // C = acc (+ bias): self-test / generic
    float* C; int ldc; const float* bias; int n_valid;
    __device__ __forceinline__ float4 bias4(int col) const { return bias ? ld4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4 b, float4) const {
        if (col >= n_valid) return;
        *reinterpret_cast<float4*>(C + (long long)row * ldc + col) = add4(a, b);
    }
};

struct TcEpiPlainAcc : EpiNoDirect {         // C = acc (accumulate == 0), C += acc (1), or atomic C += acc (2: the partial products of a
    float* C; int ldc; int n_valid; int accumulate;      // split-K launch, whose tiles of one output block run on different CTA pairs)
    __device__ __forceinline__ float4 bias4(int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ float4 pre(int row, int col) const {
        return (accumulate == 1 && col < n_valid) ? *reinterpret_cast<const float4*>(C + (long long)row * ldc + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4, float4 old) const {
        if (col >= n_valid) return;
        float4* dst = reinterpret_cast<float4*>(C + (long long)row * ldc + col);
        if (accumulate == 2) atomicAdd(dst, a);
        else *dst = add4(a, old);
    }
};

struct TcEpiBase : EpiNoDirect, EpiNoPre {    // base = x_cond-half of start_conv + bias + positional row (constant per window)
    float* base; int ld; const float* bias; const float* pos; int T;
    __device__ __forceinline__ float4 bias4(int col) const { return ld4(bias + col); }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4 b, float4) const {
        const int l = row % LP;
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l >= 1 && l <= T) r = add4(add4(a, b), ld4(pos + (long long)(l + 1) * ld + col));
        *reinterpret_cast<float4*>(base + (long long)row * ld + col) = r;
    }
};

template <int FMT>
struct TcEpiStart : EpiNoDirect {         // H = x-half GEMM + base ; row 0 = time token ; writes fp32 + operand planes
    float* H; __nv_bfloat16* Hhi; __nv_bfloat16* Hlo; int ld;
    const float* base; const float* pos; const float* temb; TSrc ts; int T; int n_windows;
    __device__ __forceinline__ float4 bias4(int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
    // the per-element operand, fetched one chunk ahead: `base` for frame rows, the time token (timestep embedding + positional row 1)
    // for row 0 of a window -- its dependent loads (step counter -> table row) then sit in the prefetch, not in the drain
    __device__ __forceinline__ float4 pre(int row, int col) const {
        const int w = row / LP;
        if (row % LP == 0 && w < n_windows) return add4(ld4(temb + (long long)ts.get(w) * ld + col), ld4(pos + ld + col));
        return ld4(base + (long long)row * ld + col);
    }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4, float4 bv) const {
        const int w = row / LP;
        const int l = (w < n_windows) ? row % LP : LP;      // rows of the rounding-up window are padding
        float4 r;
        if (l == 0)      r = bv;
        else if (l <= T) r = add4(a, bv);
        else             r = make_float4(0.f, 0.f, 0.f, 0.f);
        const long long o = (long long)row * ld + col;
        if (H) *reinterpret_cast<float4*>(H + o) = r;       // the fp32 residual copy is not needed by the fused-LN (fp16) path
        store_planes4<FMT>(Hhi + o, Hlo + o, r);
    }
};

struct TcEpiBiasScaleF32 : EpiNoDirect, EpiNoPre {  // QKV projection -> fp32 [M, ldc] (q block pre-scaled by 1/sqrt(d_k))
    float* C; int ldc; const float* bias; int scale_cols; float scale;
    __device__ __forceinline__ float4 bias4(int col) const { return ld4(bias + col); }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4 b, float4) const {
        const float s = col < scale_cols ? scale : 1.0f;
        a = add4(a, b);
        *reinterpret_cast<float4*>(C + (long long)row * ldc + col) = make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
    }
};

struct TcEpiBiasResidF32 : EpiNoDirect {  // fc / w_2: acc + bias + residual -> fp32 (pre-LayerNorm)
    float* C; int ldc; const float* bias; const float* res;
    __device__ __forceinline__ float4 bias4(int col) const { return ld4(bias + col); }
    __device__ __forceinline__ float4 pre(int row, int col) const { return ld4(res + (long long)row * ldc + col); }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4 b, float4 rv) const {
        *reinterpret_cast<float4*>(C + (long long)row * ldc + col) = add4(add4(a, b), rv);
    }
};

template <int FMT>
struct TcEpiBiasReluSplit : EpiNoDirect, EpiNoPre { // w_1: relu(acc + bias) -> operand planes (A operand of w_2)
    __nv_bfloat16* hi; __nv_bfloat16* lo; int ld; const float* bias;
    __device__ __forceinline__ float4 bias4(int col) const { return ld4(bias + col); }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4 b, float4) const {
        a = add4(a, b);
        a = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
        const long long o = (long long)row * ld + col;
        store_planes4<FMT>(hi + o, lo + o, a);
    }
};

struct TcEpiOut : EpiNoDirect, EpiNoPre { // linear_out: tokens 1..T, first d_feats columns -> compact [B,T,d_feats]
    float* out; int d_feats; const float* bias; int T; int n_windows;
    __device__ __forceinline__ float4 bias4(int col) const {    // bias has d_feats entries: guard the padded tail
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < d_feats) { b.x = bias[col]; b.y = bias[col + 1]; }
        if (col + 2 < d_feats) { b.z = bias[col + 2]; b.w = bias[col + 3]; }
        return b;
    }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4 b, float4) const {
        const int w = row / LP, l = row % LP;
        if (l < 1 || l > T || w >= n_windows || col >= d_feats) return;
        float* o = out + ((long long)w * T + (l - 1)) * d_feats + col;       // 8-byte aligned (d_feats even, col % 4 == 0)
        *reinterpret_cast<float2*>(o) = make_float2(a.x + b.x, a.y + b.y);
        if (col + 2 < d_feats) *reinterpret_cast<float2*>(o + 2) = make_float2(a.z + b.z, a.w + b.w);
    }
};

// ---- the kernel ---------------------------------------------------------------------------------
template <int BN, int FMT, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_split3_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                   const __grid_constant__ CUtensorMap mWh, const __grid_constant__ CUtensorMap mWl,
                   int M, int N, int K, Epi epi) {
    using Cfg = GemmCfg<BN, FMT>;
    constexpr int NP = Cfg::NP;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    constexpr int W_BYTES = BN * GEMM_BK * 2;
    constexpr uint32_t IDESC = (FMT == FMT_SPLIT_BF16) ? ptx::make_idesc_bf16(GEMM_BM, BN) : ptx::make_idesc_f16(GEMM_BM, BN);
    constexpr int ACC_STAGES = 512 / BN >= 2 ? 2 : 1;

    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzle tiles, computed as an OFFSET on the __shared__ array so the compiler keeps
    // the shared address space (integer round-trips turn every access into a generic LD/ST).
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    float4* epi_tiles = reinterpret_cast<float4*>(smem + STAGES * Cfg::STAGE_BYTES);          // 4 x 4 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + GEMM_EPI_WARPS * 4096);
    uint64_t* full_bar = bars;                         // [STAGES]
    uint64_t* empty_bar = bars + STAGES;               // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;           // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    // Role index, not hardware warp id: the SM sub-partition arbiter favours HIGHER warp ids, so the TMA producer (role 0)
    // and the MMA issuer (role 1) live in the two highest hardware warps and are never starved of issue slots by the
    // epilogue warps (roles 2..9 = hardware warps 0..7, whose id % 4 selects their TMEM lane quarter).
    const int lane = threadIdx.x % 32;
    const int warp = (threadIdx.x / 32 + 2) % (GEMM_THREADS / 32);
    const int m_tiles = M / GEMM_BM, n_tiles = N / BN, k_blocks = K / GEMM_BK;
    const int total_tiles = m_tiles * n_tiles;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mAh); ptx::prefetch_tmap(&mAl); ptx::prefetch_tmap(&mWh); ptx::prefetch_tmap(&mWl);
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 32 * GEMM_EPI_WARPS); }
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
        if (lane == 0) {                                 // ===== TMA producer =====
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * GEMM_BM, n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                    ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                    ptx::tma_load_2d(st, &mAh, &full_bar[s], kb * GEMM_BK, m0);
                    if (NP == 2) ptx::tma_load_2d(st + A_BYTES, &mAl, &full_bar[s], kb * GEMM_BK, m0);
                    ptx::tma_load_2d(st + NP * A_BYTES, &mWh, &full_bar[s], kb * GEMM_BK, n0);
                    if (NP == 2) ptx::tma_load_2d(st + NP * A_BYTES + W_BYTES, &mWl, &full_bar[s], kb * GEMM_BK, n0);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                 // ===== MMA issuer =====
            int s = 0; uint32_t ph = 0; int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int a = (ACC_STAGES == 2) ? (it & 1) : 0;
                const uint32_t aph = (ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
                ptx::mbar_wait(&tempty_bar[a], aph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint64_t dAh = ptx::make_smem_desc_sw128(st);
                    const uint64_t dAl = ptx::make_smem_desc_sw128(st + A_BYTES);
                    const uint64_t dWh = ptx::make_smem_desc_sw128(st + NP * A_BYTES);
                    const uint64_t dWl = ptx::make_smem_desc_sw128(st + NP * A_BYTES + W_BYTES);
#pragma unroll
                    for (int kk = 0; kk < GEMM_BK / 16; ++kk) {
                        const uint64_t adv = (uint64_t)(kk * 32 >> 4);       // +32 B per UMMA_K inside the swizzle row
                        ptx::umma_f16(d_tmem, dAh + adv, dWh + adv, IDESC, (kb | kk) != 0);
                        if (NP == 2) {
                            ptx::umma_f16(d_tmem, dAh + adv, dWl + adv, IDESC, 1);
                            ptx::umma_f16(d_tmem, dAl + adv, dWh + adv, IDESC, 1);
                        }
                    }
                    ptx::umma_commit(&empty_bar[s]);                        // frees the smem stage when the MMAs retire
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit(&tfull_bar[a]);                            // accumulator ready for the epilogue
            }
        }
    } else {                                             // ===== epilogue warps 2..9 =====
        const int quarter = (warp - 2) & 3;                    // TMEM lane quarter this warp may access
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int a = (ACC_STAGES == 2) ? (it & 1) : 0;
            const uint32_t aph = (ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
            const int m0 = (tile / n_tiles) * GEMM_BM, n0 = (tile % n_tiles) * BN;
            const int chalf = (warp - 2) >> 2;                 // which half of the tile columns this warp drains
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * BN + chalf * (BN / 2);
            float4* etile = epi_tiles + (warp - 2) * 256;
            epilogue_drain<BN / 2>(epi, etile, taddr, lane, m0 + quarter * 32, n0 + chalf * (BN / 2),
                                   [&]() { ptx::mbar_wait(&tfull_bar[a], aph); ptx::tc_fence_after(); });
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty_bar[a]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

// ---- 2-CTA variant: CTA pairs (cluster of 2) share the W tile ---------------------------------------
// One pair computes a 256 x 256 output tile with tcgen05.mma.cta_group::2 (M = 256): each CTA stages its own
// 128 A rows (hi+lo) and HALF of the W tile (128 of the 256 N rows, hi+lo), so operand traffic per CTA drops
// from 96 KB to 64 KB per k-block and three stages fit in shared memory.  The leader (even) CTA issues the
// MMAs for the pair; TMA completions of both CTAs are credited to the leader's `full` barrier; `empty` and
// `tmem_full` are signalled to both CTAs by a multicast tcgen05.commit; the epilogues of both CTAs release the
// accumulator through remote arrives on the leader's `tmem_empty` barrier.
template <int FMT> struct Gemm2Cfg {
    static constexpr int NP = FmtTraits<FMT>::NP;
    static constexpr int STAGES = fmt_is_split(FMT) ? 3 : 6;
    static constexpr int STAGE_BYTES = 2 * NP * GEMM_BM * GEMM_BK * 2;     // A planes + W-half planes, 16 KB each
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_EPI_WARPS * 4096 /*epilogue tiles*/ + 1024 + 256;
};

// DUAL = true (split formats only): the hi*hi products accumulate in one TMEM accumulator and the two cross terms (hi*lo, lo*hi:
// 2^-11 of the magnitude) in a second one; the epilogue adds them in fp32.  The tensor core TRUNCATES its fp32 accumulator after
// every MMA (measured: ~28 ulp on a K = 512 three-term product, i.e. ~0.3 ulp per accumulation, all of one sign), and that bias is
// the same at every diffusion step -- the sampler integrates it like it integrated the fp16 weight rounding.  With
// two accumulators the significant one sees a third of the accumulations (K/16 instead of 3K/16) and the truncation of the cross
// accumulator is 2^-11 smaller.  Cost: both accumulators of a 256-column tile fill the 512 TMEM columns, so the epilogue of tile
// i no longer overlaps the main loop of tile i + 1 (used for the 3-term steps only, whose main loop is three times longer).
template <int FMT, class Epi, bool DUAL = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_split3_2cta_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
                        const __grid_constant__ CUtensorMap mWh, const __grid_constant__ CUtensorMap mWl,   // W maps: 128-row boxes
                        int M, int N, int K, Epi epi, int rev = 0 /* 1: walk the tiles from the last row block to the first (L2 zig-zag) */,
                        int wpasses = 1 /* FMT_HALF only: 2 = weights as an fp16 pair, pass 0 multiplies A by mWh (the LO plane goes
                                           first), pass 1 by mWl -- see gemm_half_tma_2cta_kernel */,
                        int ksplit = 1 /* split-K: every output tile is computed as `ksplit` partial products over K / ksplit each, on
                                          different CTA pairs; the functor must accumulate atomically (TcEpiPlainAcc mode 2).  For the
                                          weight gradients of the training step: 512 x 512 outputs with K = 4096 are 4 tiles */) {
    constexpr int BN = 256;
    constexpr int NP = Gemm2Cfg<FMT>::NP;
    static_assert(!DUAL || NP == 2, "the dual-accumulator variant is for the 3-term split formats");
    constexpr int GEMM2_STAGES = Gemm2Cfg<FMT>::STAGES;
    constexpr int GEMM2_STAGE_BYTES = Gemm2Cfg<FMT>::STAGE_BYTES;
    constexpr int T_BYTES = GEMM_BM * GEMM_BK * 2;                   // 16 KB tile
    constexpr uint32_t IDESC = (FMT == FMT_SPLIT_BF16) ? ptx::make_idesc_bf16(256, BN) : ptx::make_idesc_f16(256, BN);
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzle tiles, computed as an OFFSET on the __shared__ array so the compiler keeps
    // the shared address space (integer round-trips turn every access into a generic LD/ST).
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    float4* epi_tiles = reinterpret_cast<float4*>(smem + GEMM2_STAGES * GEMM2_STAGE_BYTES);   // 4 x 4 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM2_STAGES * GEMM2_STAGE_BYTES + GEMM_EPI_WARPS * 4096);
    uint64_t* full_bar = bars;                          // [S]  (used on the leader)
    uint64_t* empty_bar = bars + GEMM2_STAGES;          // [S]  (both CTAs)
    uint64_t* tfull_bar = bars + 2 * GEMM2_STAGES;      // [2]  (both CTAs)
    uint64_t* tempty_bar = bars + 2 * GEMM2_STAGES + 2; // [2]  (used on the leader)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GEMM2_STAGES + 4);

    // Role index, not hardware warp id: the SM sub-partition arbiter favours HIGHER warp ids, so the TMA producer (role 0)
    // and the MMA issuer (role 1) live in the two highest hardware warps and are never starved of issue slots by the
    // epilogue warps (roles 2..9 = hardware warps 0..7, whose id % 4 selects their TMEM lane quarter).
    const int lane = threadIdx.x % 32;
    const int warp = (threadIdx.x / 32 + 2) % (GEMM_THREADS / 32);
    const uint32_t rank = ptx::cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x / 2, n_pairs = gridDim.x / 2;
    const int m_tiles = M / 256, n_tiles = N / BN, k_real = K / GEMM_BK / ksplit;   // k-blocks of one (partial) product
    const int k_blocks = (NP == 1 ? wpasses : 1) * k_real;          // k-blocks streamed per tile
    const int out_tiles = m_tiles * n_tiles;
    const int total_tiles = out_tiles * ksplit;                      // tile index = split * out_tiles + output tile

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mAh); ptx::prefetch_tmap(&mAl); ptx::prefetch_tmap(&mWh); ptx::prefetch_tmap(&mWl);
        for (int s = 0; s < GEMM2_STAGES; ++s) { ptx::mbar_init(&full_bar[s], 2); ptx::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 2 * GEMM_EPI_WARPS); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc_2cta(tmem_slot, 512); ptx::tmem_relinquish_2cta(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                                 // peer's barriers are initialised before any remote arrive
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();                                // prologue above overlaps the previous kernel's tail (PDL)

    if (warp == 0) {
        if (lane == 0) {                                 // ===== TMA producer (both CTAs) =====
            int s = 0; uint32_t ph = 0;
            for (int tl = pair; tl < total_tiles; tl += n_pairs) {
                const int tile_s = rev ? total_tiles - 1 - tl : tl;
                const int tile = tile_s % out_tiles, k_first = (tile_s / out_tiles) * k_real;
                const int m0 = (tile / n_tiles) * 256 + (int)rank * 128;
                const int n0 = (tile % n_tiles) * BN + (int)rank * 128;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * GEMM2_STAGE_BYTES;
                    if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * GEMM2_STAGE_BYTES);
                    else        ptx::mbar_arrive_cluster(&full_bar[s], 0);
                    const int kx = k_first + ((NP == 1 && kb >= k_real) ? kb - k_real : kb);
                    ptx::tma_load_2d_2cta(st, &mAh, &full_bar[s], kx * GEMM_BK, m0);
                    if (NP == 2) ptx::tma_load_2d_2cta(st + T_BYTES, &mAl, &full_bar[s], kx * GEMM_BK, m0);
                    ptx::tma_load_2d_2cta(st + NP * T_BYTES, (NP == 1 && kb >= k_real) ? &mWl : &mWh, &full_bar[s], kx * GEMM_BK, n0);
                    if (NP == 2) ptx::tma_load_2d_2cta(st + 3 * T_BYTES, &mWl, &full_bar[s], kx * GEMM_BK, n0);
                    if (++s == GEMM2_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {                       // ===== MMA issuer (leader CTA only) =====
            int s = 0; uint32_t ph = 0; int it = 0;
            for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
                const int a = DUAL ? 0 : (it & 1);
                const uint32_t aph = DUAL ? (it & 1) : ((it >> 1) & 1);
                ptx::mbar_wait(&tempty_bar[a], aph ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * BN;
                const uint32_t d_cross = DUAL ? tmem_base + BN : d_tmem;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + s * GEMM2_STAGE_BYTES);
                    const uint64_t dAh = ptx::make_smem_desc_sw128(st), dAl = ptx::make_smem_desc_sw128(st + T_BYTES);
                    const uint64_t dWh = ptx::make_smem_desc_sw128(st + NP * T_BYTES), dWl = ptx::make_smem_desc_sw128(st + 3 * T_BYTES);
#pragma unroll
                    for (int kk = 0; kk < GEMM_BK / 16; ++kk) {
                        const uint64_t adv = (uint64_t)(kk * 2);
                        ptx::umma_f16_2cta(d_tmem, dAh + adv, dWh + adv, IDESC, (kb | kk) != 0);
                        if (NP == 2) {
                            ptx::umma_f16_2cta(d_cross, dAh + adv, dWl + adv, IDESC, DUAL ? (uint32_t)((kb | kk) != 0) : 1u);
                            ptx::umma_f16_2cta(d_cross, dAl + adv, dWh + adv, IDESC, 1);
                        }
                    }
                    ptx::umma_commit_2cta(&empty_bar[s]);
                    if (++s == GEMM2_STAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit_2cta(&tfull_bar[a]);
            }
        }
    } else {                                             // ===== epilogue warps 2..9 (both CTAs) =====
        const int quarter = (warp - 2) & 3;
        int it = 0;
        for (int tl = pair; tl < total_tiles; tl += n_pairs, ++it) {
            const int tile = (rev ? total_tiles - 1 - tl : tl) % out_tiles;
            const int a = DUAL ? 0 : (it & 1);
            const uint32_t aph = DUAL ? (it & 1) : ((it >> 1) & 1);
            const int m0 = (tile / n_tiles) * 256 + (int)rank * 128, n0 = (tile % n_tiles) * BN;
            const int chalf = (warp - 2) >> 2;                 // which half of the tile columns this warp drains
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * BN + chalf * (BN / 2);
            float4* etile = epi_tiles + (warp - 2) * 256;
            epilogue_drain<BN / 2>(epi, etile, taddr, lane, m0 + quarter * 32, n0 + chalf * (BN / 2),
                                   [&]() { ptx::mbar_wait(&tfull_bar[a], aph); ptx::tc_fence_after(); }, DUAL ? taddr + BN : 0u);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(&tempty_bar[a], 0);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                                 // nobody leaves while the peer may still touch its smem / TMEM
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc_2cta(tmem_base, 512); }
}

// ---- CTA-pair fp16 GEMM with a TMA-store epilogue ----------------------------------------------------------
// Same main loop as gemm_split3_2cta_kernel<FMT_HALF> (256 x 256 pair tiles, cta_group::2, double-buffered TMEM), but the
// epilogue never leaves the accumulator's thread-per-row layout: a thread adds the bias (broadcast from shared memory),
// applies the functor on packed fp32x2, converts to fp16 and writes its row into a 128B-swizzled staging block
// ([128 rows][256 cols] fp16 = four boxes of 64 columns) with conflict-free 16-byte stores; an I/O warp TMA-stores every box
// as soon as its four warps have published it.  No smem transposes, no per-element address arithmetic, no STG in the
// epilogue warps (the transposing epilogue issued ~1 instruction / 13 cycles / warp and bounded the QKV projection).
// Functor interface:
//   float  tile_scale(n0)                     per-tile scalar handed to apply()
//   float2 apply(float2 acc_plus_bias, s)     element-wise epilogue
//   void   box(row0, n0, b, map, x, y)        TMA-store destination of box b (64 columns) of the CTA's 128 x 256 block
constexpr int GEMM_TMAEPI_THREADS = GEMM_THREADS + 32;                 // + store I/O warp (role 10)
struct GemmTmaEpiCfg {
    static constexpr int STAGES = 4;
    static constexpr int T_BYTES = GEMM_BM * GEMM_BK * 2;              // 16 KB
    static constexpr int STAGE_BYTES = 2 * T_BYTES;
    static constexpr int OUT_BYTES = 4 * T_BYTES;                      // staging block
    static constexpr int MAX_N = 3072;                                 // bias vector kept in shared memory
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + MAX_N * 4 + 1024 + 256;
};

// CL8 = true: the same kernel in a cluster of EIGHT = 2 (M) x 2 (N) CTA pairs that walk 512 x 512 super-tiles in lock step.  The
// streaming kernel moves 640 KB through the L2 per 256 x 256 tile (512 KB of operands for 128 KB of output) -- 11.6 KB / clk chip-wide
// at the tensor rate against the ~6.3 KB / clk the L2 delivers (B300_MICROARCH.md), which is why it (and cuBLAS on this shape)
// stops near 55 % of the tensor peak.  Here the two pairs of a super-tile ROW need the same A rows and the two pairs of a super-tile
// COLUMN the same W rows: every CTA fetches HALF of its A tile and HALF of its W-half tile (64-row boxes) and TMA-multicasts each
// to the CTA of the same rank in the neighbouring pair, so operand requests per CTA drop from 32 KB to 16 KB per k-block (multicast
// pays from cluster size 8 up: the L2 already merges <= 4 concurrent unicast requests).  Barrier protocol: a stage of pair P is
// written by P itself, by its row neighbour (A) and by its column neighbour (W), so P's MMA commit releases the stage on the
// `empty` barriers of those three pairs (6 CTAs), and every producer waits for three releases -- of exactly the pairs it writes to.
template <class Epi, bool CL8 = false>
__global__ void __launch_bounds__(GEMM_TMAEPI_THREADS, 1)
gemm_half_tma_2cta_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mW /*128-row boxes; CL8: 64-row boxes*/,
                          int M, int N, int K, const float* __restrict__ bias, const __grid_constant__ Epi epi, int rev,
                          const __grid_constant__ CUtensorMap mW2, int wpasses,
                          int* __restrict__ row_cnt /* nullptr, or one counter per 128-row block (window): +1 for every tile of the block
                                                       whose output is PERFORMED in global memory -- a consumer kernel running
                                                       concurrently (attention_half_kernel) polls it instead of waiting for this grid */) {
    // wpasses = 2 ("fp16 activations x fp16-pair weights", the steps between the single-pass and the 3-term format):
    // the k loop runs twice over the same A k-blocks, first against mW (the caller passes the LO weight plane so the small
    // terms meet an empty accumulator), then against mW2 (the hi plane): D = A W_lo^T + A W_hi^T with A rounded to fp16 once.
    // The weight rounding -- the error the sampler integrates coherently over steps -- is gone; the activation rounding, which is
    // fresh at every step, stays.  Not combined with CL8.
    using Cfg = GemmTmaEpiCfg;
    constexpr int STAGES = Cfg::STAGES, T_BYTES = Cfg::T_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES, BN = 256;
    constexpr uint32_t IDESC = ptx::make_idesc_f16(256, BN);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* out_smem = smem + STAGES * STAGE_BYTES;
    float* vec = reinterpret_cast<float*>(out_smem + Cfg::OUT_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(vec + Cfg::MAX_N);
    uint64_t* full_bar = bars;                          // [S] leader
    uint64_t* empty_bar = bars + STAGES;                // [S] both
    uint64_t* tfull_bar = bars + 2 * STAGES;            // [2] both
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;       // [2] leader
    uint64_t* out_ready = bars + 2 * STAGES + 4;        // [4 boxes] written by the box's 4 warps
    uint64_t* box_free = bars + 2 * STAGES + 8;         // [4 boxes] previous store has finished reading the box
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 12);

    const int lane = threadIdx.x % 32;
    const int hw_warp = threadIdx.x / 32;               // hardware warps 0..7 = epilogue roles 2..9; 8, 9, 10 = roles 0, 1, 10
    const int warp = hw_warp < 8 ? hw_warp + 2 : (hw_warp == 10 ? 10 : hw_warp - 8);
    const uint32_t crank = ptx::cluster_ctarank();       // rank in the cluster (pair = crank / 2 when CL8)
    const uint32_t rank = crank & 1u;                    // rank in the CTA pair
    const bool leader = rank == 0;
    const uint32_t leader_rank = crank & ~1u;
    const int cmi = CL8 ? (int)(crank >> 2) : 0, cni = CL8 ? (int)((crank >> 1) & 1u) : 0;     // pair position in the 2 x 2 super-tile
    // CL8: `pair` / `n_pairs` count CLUSTERS and `total_tiles` super-tiles; tile_of() maps to this pair's 256 x 256 tile
    const int pair = CL8 ? blockIdx.x / 8 : blockIdx.x / 2, n_pairs = CL8 ? gridDim.x / 8 : gridDim.x / 2;
    const int m_tiles = M / 256, n_tiles = N / BN, k_real = K / GEMM_BK;
    const int k_blocks = (CL8 ? 1 : wpasses) * k_real;                // k-blocks streamed per tile
    const int total_tiles = CL8 ? (m_tiles / 2) * (n_tiles / 2) : m_tiles * n_tiles;
    auto tile_of = [&](int tl, int& tm, int& tn) {
        const int t = rev ? total_tiles - 1 - tl : tl;
        if (CL8) { const int sn = n_tiles / 2; tm = 2 * (t / sn) + cmi; tn = 2 * (t % sn) + cni; }
        else     { tm = t / n_tiles; tn = t % n_tiles; }
    };
    const uint16_t pair_mask = (uint16_t)(3u << leader_rank);
    // CL8: the three pairs whose producers write into this pair's stages: itself, its row neighbour (A), its column neighbour (W)
    const uint16_t release_mask = CL8 ? (uint16_t)((3u << (2 * (cmi * 2 + cni))) | (3u << (2 * (cmi * 2 + (1 - cni)))) | (3u << (2 * ((1 - cmi) * 2 + cni))))
                                      : pair_mask;
    const uint16_t mcast_a = (uint16_t)((1u << (2 * (cmi * 2 + 0) + rank)) | (1u << (2 * (cmi * 2 + 1) + rank)));
    const uint16_t mcast_w = (uint16_t)((1u << (2 * (0 * 2 + cni) + rank)) | (1u << (2 * (1 * 2 + cni) + rank)));

    for (int i = threadIdx.x; i < N; i += blockDim.x) vec[i] = bias[i];
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mA); ptx::prefetch_tmap(&mW); ptx::prefetch_tmap(&mW2);
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 2); ptx::mbar_init(&empty_bar[s], CL8 ? 3 : 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 2 * GEMM_EPI_WARPS); }
        for (int b = 0; b < 4; ++b) { ptx::mbar_init(&out_ready[b], 4); ptx::mbar_init(&box_free[b], 1); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc_2cta(tmem_slot, 512); ptx::tmem_relinquish_2cta(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (row_cnt) {
        // streamed consumer: the dependent grid starts as soon as every CTA here has passed its own wait (so everything before
        // this kernel has completed when the consumer runs) and then follows the counters, not this grid's completion
        ptx::grid_dep_wait();
        ptx::grid_dep_launch();
        if (threadIdx.x == 0) timeline_mark(0, 0);
    } else {
        ptx::grid_dep_launch();
        ptx::grid_dep_wait();                            // prologue above overlaps the previous kernel's tail (PDL)
    }

    if (warp == 0) {
        if (lane == 0) {                                 // ===== TMA producer (both CTAs) =====
            int s = 0; uint32_t ph = 0;
            for (int tl = pair; tl < total_tiles; tl += n_pairs) {
                int tm, tn;
                tile_of(tl, tm, tn);
                const int m0 = tm * 256 + (int)rank * 128;
                const int n0 = tn * BN + (int)rank * 128;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                    else        ptx::mbar_arrive_cluster(&full_bar[s], leader_rank);
                    if (CL8) {       // half of each tile, multicast to the same-rank CTA of the row (A) / column (W) neighbour pair
                        ptx::tma_load_2d_2cta_mc(st + cni * (T_BYTES / 2), &mA, &full_bar[s], kb * GEMM_BK, m0 + cni * 64, mcast_a);
                        ptx::tma_load_2d_2cta_mc(st + T_BYTES + cmi * (T_BYTES / 2), &mW, &full_bar[s], kb * GEMM_BK, n0 + cmi * 64, mcast_w);
                    } else {
                        const bool second = kb >= k_real;
                        const int kx = second ? kb - k_real : kb;
                        ptx::tma_load_2d_2cta(st, &mA, &full_bar[s], kx * GEMM_BK, m0);
                        ptx::tma_load_2d_2cta(st + T_BYTES, second ? &mW2 : &mW, &full_bar[s], kx * GEMM_BK, n0);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {                       // ===== MMA issuer (leader CTA only) =====
            int s = 0; uint32_t ph = 0; int it = 0;
            for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
                const int a = it & 1;
                ptx::mbar_wait(&tempty_bar[a], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t dA = ptx::make_smem_desc_sw128(st), dW = ptx::make_smem_desc_sw128(st + T_BYTES);
#pragma unroll
                    for (int kk = 0; kk < GEMM_BK / 16; ++kk)
                        ptx::umma_f16_2cta(d_tmem, dA + (uint64_t)(kk * 2), dW + (uint64_t)(kk * 2), IDESC, (kb | kk) != 0);
                    ptx::umma_commit_2cta_mask(&empty_bar[s], release_mask);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit_2cta_mask(&tfull_bar[a], pair_mask);
            }
        }
    } else if (warp == 10) {
        if (lane == 0) {                                 // ===== store I/O (both CTAs) =====
            int it = 0;
            int prev_blk = -1;                           // row block of the previous tile, published one tile late (no stall)
            auto publish = [&](int blk) {           // the block's bulk stores are PERFORMED (wait_group, non-.read) before this release
                __threadfence();
                ptx::red_release_gpu_add(row_cnt + blk, 1);
            };
            for (int tl = pair; tl < total_tiles; tl += n_pairs, ++it) {
                int tm, tn;
                tile_of(tl, tm, tn);
                const int row0 = tm * 256 + (int)rank * 128, n0 = tn * BN;
#pragma unroll 1
                for (int o = 0; o < 4; ++o) {
                    const int b = (o & 1) * 2 + (o >> 1);              // 0, 2, 1, 3: both column halves' first box first
                    const CUtensorMap* map; int x, y;
                    epi.box(row0, n0, b, map, x, y);
                    ptx::mbar_wait(&out_ready[b], it & 1);
                    ptx::tma_store_2d(map, out_smem + b * T_BYTES, x, y);
                    ptx::tma_store_commit();
                    ptx::tma_store_wait_read();
                    ptx::mbar_arrive(&box_free[b]);
                }
                if (row_cnt) {
                    if (prev_blk >= 0) { ptx::tma_store_wait_pending<4>(); publish(prev_blk); }   // all but this tile's 4 boxes are performed
                    prev_blk = row0 / GEMM_BM;
                }
            }
            ptx::tma_store_wait_all();
            if (row_cnt && prev_blk >= 0) publish(prev_blk);
            if (row_cnt) timeline_mark(0, 1);
        }
    } else {                                             // ===== epilogue warps 2..9 (both CTAs) =====
        const int quarter = (warp - 2) & 3, hf = (warp - 2) >> 2;
        const int r = quarter * 32 + lane;                // row within the CTA's 128 == TMEM lane
        const int sw = r & 7;
        int it = 0;
        for (int tl = pair; tl < total_tiles; tl += n_pairs, ++it) {
            int tm, tn;
            tile_of(tl, tm, tn);
            const int a = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int n0 = tn * BN;
            const float ts = epi.tile_scale(n0);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * BN + hf * 128;
            const float* bvec = vec + n0 + hf * 128;
            ptx::mbar_wait(&tfull_bar[a], aph);
            ptx::tc_fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(taddr + c * 32, raw);
                const int box = 2 * hf + (c >> 1);
                if ((c & 1) == 0) ptx::mbar_wait(&box_free[box], (it & 1) ^ 1);
                uint8_t* rrow = out_smem + box * T_BYTES + r * 128;
                ptx::tmem_ld_wait();
                if (c == 3) {                             // accumulator stage drained: hand it back to the MMA issuer
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(&tempty_bar[a], leader_rank);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t hw[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int e = j * 8 + q * 2;
                        const float2 bs = *reinterpret_cast<const float2*>(bvec + c * 32 + e);
                        float2 v = __fadd2_rn(make_float2(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1])), bs);
                        v = epi.apply(v, ts);
                        const __half2 h = __floats2half2_rn(v.x, v.y);
                        hw[q] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(rrow + ((((c & 1) * 4 + j) ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                }
                if (c & 1) {                              // box complete for this warp: publish to the async proxy
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&out_ready[box]);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc_2cta(tmem_base, 512); }
}

// QKV projection: tile (256 rows = 2 windows, 256 columns = one (section, head)) -> plane [(w*H + h)*128 + l][256]
struct TmaEpiQKV {
    CUtensorMap mQ, mK, mV;                               // fp16 planes, 128-row x 64-col boxes
    int n_head; float q_scale;
    __device__ __forceinline__ float tile_scale(int n0) const { return n0 < n_head * 256 ? q_scale : 1.0f; }
    __device__ __forceinline__ float2 apply(float2 v, float s) const { return make_float2(v.x * s, v.y * s); }
    __device__ __forceinline__ void box(int row0, int n0, int b, const CUtensorMap*& m, int& x, int& y) const {
        const int hw = n_head * 256, sec = n0 / hw, h = (n0 - sec * hw) >> 8;
        m = sec == 0 ? &mQ : (sec == 1 ? &mK : &mV);
        x = b * 64;
        y = ((row0 / LP) * n_head + h) * 128;
    }
};
// FFN w_1: relu(acc + bias) -> F plane [M, 512]
struct TmaEpiRelu {
    CUtensorMap mF;
    __device__ __forceinline__ float tile_scale(int) const { return 1.0f; }
    __device__ __forceinline__ float2 apply(float2 v, float) const { return make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)); }
    __device__ __forceinline__ void box(int row0, int n0, int b, const CUtensorMap*& m, int& x, int& y) const {
        m = &mF; x = n0 + b * 64; y = row0;
    }
};

// ---- column-split fused GEMM + LayerNorm: cluster of 4 = two CTA pairs ---------------------------------
// The post-LN sub-layers (attention fc, FFN w_2) need the statistics of whole 512-wide rows.  A full-row tile would fill all
// 512 TMEM columns, so MMA and LayerNorm epilogue of a CTA would serialise (round 1 measured such kernels at 6 % tensor-pipe
// activity; they are gone from the tree).  Here a cluster of four CTAs owns a 256-row block: pair p
// (cluster ranks 2p, 2p+1; cta_group::2, M = 256) computes columns [256p, 256p+256) -- an ordinary 256 x 256
// pair tile whose fp32 accumulator takes 256 TMEM columns, DOUBLE-BUFFERED, so the MMAs of block i+1 run under
// the epilogue of block i.  A CTA therefore holds 128 rows x 256 columns; the row statistics of LayerNorm
// (sum, sum of squares -- one pass) are exchanged with the CTA of the other pair that holds the same rows
// (rank ^ 2) through distributed shared memory: st.async writes the partial into the partner's stat buffer and
// completes transaction bytes on the partner's mbarrier (no fences).  A thread keeps its 128 x-values (acc + bias
// + residual) in registers between the statistics and the normalise pass: one TMEM read, no TMEM write-back,
// and the accumulator stage is released before the exchange.
//
// Epilogue I/O is all TMA and all in the accumulator's native thread-per-row layout (no smem transposes, no
// LDG/STG in the epilogue warps -- ncu: the transposing version issued one instruction per 13 cycles per warp):
// the fp16 residual block (128 rows x 256 columns = four 128B-swizzled boxes of 64 columns) is TMA-loaded into
// shared memory, a thread reads its own row with conflict-free 16-byte loads, later overwrites the same bytes
// with the normalised fp16 output, and a dedicated I/O warp TMA-stores each box as soon as its four warps are
// done with it, then refills it with the NEXT block's residual.  Math runs on packed fp32x2 (FADD2 / FFMA2).
constexpr int GEMM_LN4_THREADS = GEMM_THREADS + 32;                    // + epilogue I/O warp (role 10)
struct GemmLn4Cfg {
    static constexpr int STAGES = 4;
    static constexpr int T_BYTES = GEMM_BM * GEMM_BK * 2;              // 16 KB
    static constexpr int STAGE_BYTES = 2 * T_BYTES;                    // A rows of this CTA + its 128-row half of the pair's W tile
    static constexpr int RES_BYTES = 4 * T_BYTES;                      // residual / output block: 4 boxes [128 rows][64 cols] fp16
    static constexpr int STAT_BYTES = 2 * 4 * 128 * 8;                 // [tile parity][source][row] (sum, sumsq)
    static constexpr int VEC_BYTES = 3 * 256 * 4;                      // bias | gamma | beta of this pair's 256 columns
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RES_BYTES + STAT_BYTES + VEC_BYTES + 1024 + 256;
};

__device__ __forceinline__ float2 ld_shared_f2(const float* p) { return *reinterpret_cast<const float2*>(p); }

static __global__ void __launch_bounds__(GEMM_LN4_THREADS, 1)
gemm_ln_half_c4_kernel(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mW /*128-row boxes*/,
                       const __grid_constant__ CUtensorMap mH /* residual in / output out: [M,512] fp16, 128-row x 64-col boxes */,
                       int M, int K, const float* __restrict__ bias, const float* __restrict__ gamma,
                       const float* __restrict__ beta, int rev,
                       const __grid_constant__ CUtensorMap mW2, int wpasses /* 2: second pass over K against mW2, see gemm_half_tma_2cta_kernel */) {
    using Cfg = GemmLn4Cfg;
    constexpr int STAGES = Cfg::STAGES, T_BYTES = Cfg::T_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
    constexpr uint32_t IDESC = ptx::make_idesc_f16(256, 256);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* res_smem = smem + STAGES * STAGE_BYTES;                                   // 1024-aligned boxes
    float2* stat = reinterpret_cast<float2*>(res_smem + Cfg::RES_BYTES);
    float* vec = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stat) + Cfg::STAT_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(vec + 3 * 256);
    uint64_t* full_bar = bars;                         // [S] pair leader
    uint64_t* empty_bar = bars + STAGES;               // [S] both CTAs of the pair
    uint64_t* tfull_bar = bars + 2 * STAGES;           // [2] both
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;      // [2] pair leader
    uint64_t* stat_bar = bars + 2 * STAGES + 4;        // [2 parities][4 lane quarters]: 64 local arrivals + 512 remote tx bytes
    uint64_t* res_full = bars + 2 * STAGES + 12;       // [4 boxes] residual landed (I/O thread expect_tx)
    uint64_t* out_ready = bars + 2 * STAGES + 16;      // [4 boxes] output written by the box's 4 warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 20);

    const int lane = threadIdx.x % 32;
    // role index: hardware warps 0..7 = epilogue roles 2..9 (warp id % 4 selects the TMEM lane quarter), hardware warps
    // 8, 9, 10 = TMA producer (0), MMA issuer (1), epilogue I/O (10) -- see gemm_split3_kernel
    const int hw_warp = threadIdx.x / 32;
    const int warp = hw_warp < 8 ? hw_warp + 2 : (hw_warp == 10 ? 10 : hw_warp - 8);
    const uint32_t rank4 = ptx::cluster_ctarank();
    const uint32_t prank = rank4 & 1u, chalf = rank4 >> 1, leader_rank = rank4 & ~1u, partner = rank4 ^ 2u;
    const bool leader = prank == 0;
    const uint16_t pair_mask = (uint16_t)(3u << leader_rank);
    const int cl = blockIdx.x / 4, n_cl = gridDim.x / 4;
    const int m_tiles = M / 256, k_real = K / GEMM_BK, kb_total = wpasses * k_real;

    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        vec[i] = bias[chalf * 256 + i]; vec[256 + i] = gamma[chalf * 256 + i]; vec[512 + i] = beta[chalf * 256 + i];
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&mA); ptx::prefetch_tmap(&mW); ptx::prefetch_tmap(&mH); ptx::prefetch_tmap(&mW2);
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 2); ptx::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 2 * GEMM_EPI_WARPS); }
        for (int b = 0; b < 8; ++b) ptx::mbar_init(&stat_bar[b], 64);
        for (int b = 0; b < 4; ++b) { ptx::mbar_init(&res_full[b], 1); ptx::mbar_init(&out_ready[b], 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) { ptx::tmem_alloc_2cta(tmem_slot, 512); ptx::tmem_relinquish_2cta(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();                                // prologue above overlaps the previous kernel's tail (PDL)

    if (warp == 0) {
        if (lane == 0) {                                 // ===== TMA producer (every CTA) =====
            int s = 0; uint32_t ph = 0;
            for (int tl = cl; tl < m_tiles; tl += n_cl) {
                const int tile = rev ? m_tiles - 1 - tl : tl;
                const int m0 = tile * 256 + (int)prank * 128;
                const int n0 = (int)chalf * 256 + (int)prank * 128;
                for (int kb = 0; kb < kb_total; ++kb) {
                    ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    if (leader) ptx::mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
                    else        ptx::mbar_arrive_cluster(&full_bar[s], leader_rank);
                    const bool second = kb >= k_real;
                    const int kx = second ? kb - k_real : kb;
                    ptx::tma_load_2d_2cta(st, &mA, &full_bar[s], kx * GEMM_BK, m0);
                    ptx::tma_load_2d_2cta(st + T_BYTES, second ? &mW2 : &mW, &full_bar[s], kx * GEMM_BK, n0);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {                       // ===== MMA issuer (pair leaders) =====
            int s = 0; uint32_t ph = 0; int it = 0;
            for (int tile = cl; tile < m_tiles; tile += n_cl, ++it) {
                const int a = it & 1;
                ptx::mbar_wait(&tempty_bar[a], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * 256;
                for (int kb = 0; kb < kb_total; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + s * STAGE_BYTES);
                    const uint64_t dA = ptx::make_smem_desc_sw128(st), dW = ptx::make_smem_desc_sw128(st + T_BYTES);
#pragma unroll
                    for (int kk = 0; kk < GEMM_BK / 16; ++kk)
                        ptx::umma_f16_2cta(d_tmem, dA + (uint64_t)(kk * 2), dW + (uint64_t)(kk * 2), IDESC, (kb | kk) != 0);
                    ptx::umma_commit_2cta_mask(&empty_bar[s], pair_mask);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit_2cta_mask(&tfull_bar[a], pair_mask);
            }
        }
    } else if (warp == 10) {
        if (lane == 0) {                                 // ===== epilogue I/O: residual loads, output stores (every CTA) =====
            const int gx = (int)chalf * 256;
            auto load_box = [&](int b, int tile) {
                ptx::mbar_arrive_expect_tx(&res_full[b], T_BYTES);
                ptx::tma_load_2d(res_smem + b * T_BYTES, &mH, &res_full[b], gx + b * 64, tile * 256 + (int)prank * 128);
            };
            auto tmap = [&](int tl) { return rev ? m_tiles - 1 - tl : tl; };
            if (cl < m_tiles) { const int t0 = tmap(cl); load_box(0, t0); load_box(2, t0); load_box(1, t0); load_box(3, t0); }
            int it = 0;
            for (int tl = cl; tl < m_tiles; tl += n_cl, ++it) {
                const int tile = tmap(tl);
                const int nxt = tl + n_cl;
#pragma unroll 1
                for (int o = 0; o < 4; ++o) {
                    const int b = (o & 1) * 2 + (o >> 1);              // 0, 2, 1, 3: both column halves' first box first
                    ptx::mbar_wait(&out_ready[b], it & 1);
                    ptx::tma_store_2d(&mH, res_smem + b * T_BYTES, gx + b * 64, tile * 256 + (int)prank * 128);
                    ptx::tma_store_commit();
                    if (nxt < m_tiles) { ptx::tma_store_wait_read(); load_box(b, tmap(nxt)); }
                }
            }
            ptx::tma_store_wait_all();
        }
    } else {                                             // ===== LayerNorm epilogue warps 2..9 (every CTA) =====
        const int quarter = (warp - 2) & 3, hf = (warp - 2) >> 2;
        const int r = quarter * 32 + lane;                // row within the CTA's 128 == TMEM lane
        const int ccol = hf * 128;                        // this warp's first column inside the CTA's 256
        const int src = (int)chalf * 2 + hf;              // which of the row's four partial statistics this thread owns
        const int sw = r & 7;                             // 128B-swizzle phase of this row
        int it = 0;
        for (int tile = cl; tile < m_tiles; tile += n_cl, ++it) {
            const int a = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * 256 + ccol;
            ptx::mbar_wait(&tfull_bar[a], aph);
            ptx::tc_fence_after();
            float2 x[64];
            float2 sum2 = make_float2(0.f, 0.f), sq2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t raw[32];
                ptx::tmem_ld_32x32(taddr + c * 32, raw);
                const int box = 2 * hf + (c >> 1);
                if ((c & 1) == 0) ptx::mbar_wait(&res_full[box], it & 1);
                const uint8_t* rrow = res_smem + box * T_BYTES + r * 128;
                uint4 rv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) rv[j] = *reinterpret_cast<const uint4*>(rrow + ((((c & 1) * 4 + j) ^ sw) << 4));
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t hw[4] = {rv[j].x, rv[j].y, rv[j].z, rv[j].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int e = j * 8 + q * 2;          // column inside the chunk
                        const float2 rs = __half22float2(*reinterpret_cast<const __half2*>(&hw[q]));
                        const float2 bs = ld_shared_f2(vec + ccol + c * 32 + e);
                        float2 v = __fadd2_rn(make_float2(__uint_as_float(raw[e]), __uint_as_float(raw[e + 1])), rs);
                        v = __fadd2_rn(v, bs);
                        x[c * 16 + j * 4 + q] = v;
                        sum2 = __fadd2_rn(sum2, v);
                        sq2 = __ffma2_rn(v, v, sq2);
                    }
                }
            }
            ptx::tc_fence_before();                       // accumulator stage drained: hand it back to the MMA issuer
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(&tempty_bar[a], leader_rank);
            // ---- row statistics: four partials per row (2 pairs x 2 column halves), two of them remote ----
            const float sum = sum2.x + sum2.y, sq = sq2.x + sq2.y;
            const int buf = it & 1;
            float2* mine = stat + (buf * 4 + src) * 128 + r;
            uint64_t* sb = &stat_bar[buf * 4 + quarter];
            *mine = make_float2(sum, sq);
            ptx::st_async_f2(mine, sb, partner, sum, sq);
            ptx::mbar_arrive_expect_tx(sb, 8);            // this thread's arrival + the 8 bytes its remote counterpart sends
            ptx::mbar_wait(sb, aph);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 p = stat[(buf * 4 + k) * 128 + r]; s1 += p.x; s2 += p.y; }
            const float mean = s1 * (1.0f / 512.0f);
            const float rstd = rsqrtf(fmaxf(s2 * (1.0f / 512.0f) - mean * mean, 0.f) + 1e-5f);
            const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(-mean * rstd, -mean * rstd);
            // ---- normalise from registers, fp16 output over the residual bytes of the same row ----
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int box = 2 * hf + (c >> 1);
                uint8_t* rrow = res_smem + box * T_BYTES + r * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t hw[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int e = j * 8 + q * 2;
                        const float2 g2 = ld_shared_f2(vec + 256 + ccol + c * 32 + e), b2 = ld_shared_f2(vec + 512 + ccol + c * 32 + e);
                        float2 y = __ffma2_rn(x[c * 16 + j * 4 + q], rs2, nm2);
                        y = __ffma2_rn(y, g2, b2);
                        const __half2 h = __floats2half2_rn(y.x, y.y);
                        hw[q] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(rrow + ((((c & 1) * 4 + j) ^ sw) << 4)) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                }
                if (c & 1) {                              // box complete for this warp: publish to the async proxy
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&out_ready[box]);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();                                 // nobody leaves while a peer may still write its smem / TMEM
    if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc_2cta(tmem_base, 512); }
}

}  // namespace egoego
