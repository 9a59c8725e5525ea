// Implicit-GEMM convolution on the 5th-generation tensor cores (ResNet-18 flow encoder of HeadNet, egoego/model/resnet.py:9-18
// as called from head_estimation_transformer.py:216-224; SURVEY.md 8a row a22 / BASELINE configs[3]).
//
//   Y[m, co] = relu( sum_k A[m, k] Wf[co, k] + bias[co] (+ identity[m, co]) ),   m = (n, oy, ox),  k = (ky, kx, ci)
//
// with fp16 operands (one rounding of each operand to 11 significant bits -- the numerics of the reference's own GPU path, cuDNN
// with torch's default allow_tf32 = True, which rounds both conv operands to TF32's 11 bits), fp32 accumulation in TMEM and an
// fp32 epilogue; the residual stream between BasicBlocks stays fp32.
//
// The im2col matrix is never materialised for the 3x3 / 1x1 convolutions: activations are NHWC fp16 with Cin % 64 == 0, so one
// 64-wide k-block of an output pixel's row is ONE contiguous 128-byte run of the input (a single filter tap, 64 channels) or
// zeros (padding).  Four producer warps gather those runs with coalesced 16-byte loads (8 lanes per row) straight into the
// 128-byte-swizzled K-major tile layout tcgen05.mma reads; the weight tile [BN, 64] comes by TMA.  The 7x7 stem (3 input
// channels) goes through a small explicit im2col ([M, 256] fp16, k = tap * 4 + channel) and then the same kernel as a "1x1
// convolution" over that matrix.
//
// Structure (persistent, one CTA per SM, cta_group::1, 416 threads):
//   warps 0-3   A producers (gather -> swizzled smem, fence.proxy.async, mbarrier arrive); warp 0 lane 0 also issues the W TMA
//   warp 4      MMA issuer: tcgen05.mma 128 x BN x 16, four per k-block, tcgen05.commit frees the stage
//   warps 5-12  epilogue: tcgen05.ld -> smem transpose -> coalesced bias / residual / ReLU -> fp32 and / or fp16 NHWC stores
// Accumulators are double-buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the main loop of tile i + 1.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "gemm_tcgen05.cuh"

namespace egoego {

struct ConvTcDesc {
    int Cin, Cout, kh, kw, stride, pad, Hin, Win, Hout, Wout;     // Cin % 64 == 0 (dense mode: kh = kw = 1, Cin = K)
};

constexpr int CONV_TC_PRODUCERS = 128;
constexpr int CONV_TC_THREADS = CONV_TC_PRODUCERS + 32 + 32 * GEMM_EPI_WARPS;      // 416
template <int BN> struct ConvTcCfg {
    static constexpr int A_BYTES = 128 * 64 * 2;                    // 16 KB
    static constexpr int W_BYTES = BN * 64 * 2;
    static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
    static constexpr int STAGES = BN == 256 ? 3 : 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_EPI_WARPS * 4096 + 1024 + 256;
};

// epilogue functor: coalesced layout (a lane owns 4 consecutive channels of one output pixel)
struct ConvTcEpi : EpiNoDirect {
    const float* bias; const float* resid; float* y32; __half* y16; int Cout, M, relu;
    __device__ __forceinline__ float4 bias4(int col) const { return col < Cout ? ld4(bias + col) : make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ float4 pre(int row, int col) const {
        return (resid && row < M && col < Cout) ? ld4(resid + (long long)row * Cout + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4 b, float4 r) const {
        if (row >= M || col >= Cout) return;
        float4 v = make_float4(a.x + b.x + r.x, a.y + b.y + r.y, a.z + b.z + r.z, a.w + b.w + r.w);
        if (relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
        const long long o = (long long)row * Cout + col;
        if (y32) *reinterpret_cast<float4*>(y32 + o) = v;
        if (y16) {
            const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
            *reinterpret_cast<uint2*>(y16 + o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        }
    }
};

template <int BN>
__global__ void __launch_bounds__(CONV_TC_THREADS, 1)
conv_tc_kernel(const __half* __restrict__ X /* NHWC fp16 */, const __grid_constant__ CUtensorMap mW /* [Cout_pad, Kpad] fp16, BN-row boxes */,
               ConvTcDesc d, int M, int k_blocks, ConvTcEpi epi) {
    using Cfg = ConvTcCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr uint32_t IDESC = ptx::make_idesc_f16(128, BN);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    float4* epi_tiles = reinterpret_cast<float4*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + GEMM_EPI_WARPS * 4096);
    uint64_t* full_bar = bars;                         // [STAGES]: 128 producer arrivals + the TMA thread's expect_tx arrival
    uint64_t* empty_bar = bars + STAGES;               // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;           // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int m_tiles = (M + 127) / 128, n_tiles = (d.Cout + BN - 1) / BN;
    const int total_tiles = m_tiles * n_tiles;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&mW);
        for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], CONV_TC_PRODUCERS + 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tfull_bar[a], 1); ptx::mbar_init(&tempty_bar[a], 32 * GEMM_EPI_WARPS); }
        ptx::fence_barrier_init();
    }
    if (warp == 4) { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {                                      // ===== A producers (+ W TMA) =====
        const int p = threadIdx.x;                       // 0..127
        const int chunk = p & 7, rg = p >> 3;            // 16-byte chunk of the row's 128 bytes; rows rg + 16 i
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * BN;
            // decode this thread's 8 output pixels once per tile
            const __half* xb[8]; int iy0[8], ix0[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + rg + 16 * i;
                if (m < M) {
                    const int n = m / (d.Hout * d.Wout), rem = m - n * d.Hout * d.Wout, oy = rem / d.Wout, ox = rem - oy * d.Wout;
                    xb[i] = X + (long long)n * d.Hin * d.Win * d.Cin;
                    iy0[i] = oy * d.stride - d.pad; ix0[i] = ox * d.stride - d.pad;
                } else { xb[i] = nullptr; iy0[i] = 0; ix0[i] = 0; }
            }
            for (int kb = 0; kb < k_blocks; ++kb) {
                const int k = kb * 64;
                const int tap = k / d.Cin, ci = k - tap * d.Cin, ky = tap / d.kw, kx = tap - ky * d.kw;
                uint4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {            // all loads first (8 independent 16-byte requests per thread)
                    const int iy = iy0[i] + ky, ix = ix0[i] + kx;
                    v[i] = make_uint4(0u, 0u, 0u, 0u);
                    if (xb[i] && iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win)
                        v[i] = __ldg(reinterpret_cast<const uint4*>(xb[i] + ((long long)iy * d.Win + ix) * d.Cin + ci) + chunk);
                }
                ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                if (p == 0) {
                    ptx::mbar_arrive_expect_tx(&full_bar[s], Cfg::W_BYTES);
                    ptx::tma_load_2d(st + Cfg::A_BYTES, &mW, &full_bar[s], k, n0);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {            // K-major SW128: row r at r * 128 B, chunk j at (j ^ (r & 7)) * 16 B
                    const int r = rg + 16 * i;
                    *reinterpret_cast<uint4*>(st + r * 128 + ((chunk ^ (r & 7)) << 4)) = v[i];
                }
                ptx::fence_proxy_async();                // generic-proxy smem writes -> visible to the tensor core
                ptx::mbar_arrive(&full_bar[s]);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {                                 // ===== MMA issuer =====
            int s = 0; uint32_t ph = 0; int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int a = it & 1;
                ptx::mbar_wait(&tempty_bar[a], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + a * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph);
                    ptx::tc_fence_after();
                    const uint32_t st = ptx::smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint64_t dA = ptx::make_smem_desc_sw128(st), dW = ptx::make_smem_desc_sw128(st + Cfg::A_BYTES);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        ptx::umma_f16(d_tmem, dA + (uint64_t)(kk * 2), dW + (uint64_t)(kk * 2), IDESC, (kb | kk) != 0);
                    ptx::umma_commit(&empty_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                ptx::umma_commit(&tfull_bar[a]);
            }
        }
    } else {                                             // ===== epilogue warps 5..12 =====
        const int ew = warp - 5;                         // 0..7
        const int quarter = warp & 3;                    // TMEM lane quarter this HARDWARE warp may access (warp id % 4)
        // two warps share each lane quarter (hardware warps q, q + 4 of the eight): they split the tile's columns
        const int chalf = (ew >> 2);                     // warps 5..8 -> 0, 9..12 -> 1 (5%4=1,6%4=2,7%4=3,8%4=0 / 9..12 likewise)
        constexpr int CW = BN / 2;                       // columns per warp
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * BN;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + a * BN + chalf * CW;
            float4* etile = epi_tiles + ew * 256;
            ptx::mbar_wait(&tfull_bar[a], aph);
            ptx::tc_fence_after();
            const int row_base = m0 + quarter * 32, col_begin = n0 + chalf * CW;
#pragma unroll 1
            for (int c = 0; c < CW / 32; ++c) {
                const int col0 = col_begin + c * 32;
                float4 pre[8];
                epilogue_prefetch(epi, pre, lane, row_base, col0);
                const float4 b4 = epi.bias4(col0 + 4 * (lane & 7));
                uint32_t raw[32];
                ptx::tmem_ld_32x32(taddr + c * 32, raw);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    etile[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                                                                    __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
                __syncwarp();
                const int j = lane & 7;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int r = q * 4 + (lane >> 3);
                    epi.apply4(row_base + r, col0 + 4 * j, etile[r * 8 + (j ^ (r & 7))], b4, pre[q]);
                }
                __syncwarp();
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(&tempty_bar[a]);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem_base, 512); }
}

// stem: flow [N, 224, 224, 2] fp32 -> im2col rows [M = N * 112 * 112, 256] fp16 of the 7x7 stride-2 pad-3 convolution over the
// (flow, 0) 3-channel image: k = (ky * 7 + kx) * 4 + c (c = 0, 1 the flow, 2 the reference's zero channel, 3 alignment), k >= 196 zero.
// One thread per (row, tap): a 2-float load and one 8-byte store; taps 49..63 write the zero padding.
static __global__ void rn_stem_im2col_kernel(const float* __restrict__ flow, __half* __restrict__ A, long long M) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * 64) return;
    const long long m = i >> 6;
    const int tap = (int)(i & 63);
    uint2 out = make_uint2(0u, 0u);
    if (tap < 49) {
        const int n = (int)(m / (112 * 112)), rem = (int)(m - (long long)n * 112 * 112), oy = rem / 112, ox = rem - oy * 112;
        const int ky = tap / 7, kx = tap - ky * 7, iy = oy * 2 - 3 + ky, ix = ox * 2 - 3 + kx;
        if (iy >= 0 && iy < 224 && ix >= 0 && ix < 224) {
            const float2 f = __ldg(reinterpret_cast<const float2*>(flow + (((long long)n * 224 + iy) * 224 + ix) * 2));
            const __half2 h = __floats2half2_rn(f.x, f.y);
            out.x = *reinterpret_cast<const uint32_t*>(&h);
        }
    }
    *reinterpret_cast<uint2*>(A + m * 256 + tap * 4) = out;
}

// MaxPool2d(kernel 3, stride 2, padding 1) on NHWC fp16, 8 channels per thread
static __global__ void rn_maxpool16_kernel(const __half* __restrict__ X, __half* __restrict__ Y, int N, int Hin, int Win, int C, int Hout, int Wout) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c8n = C / 8;
    if (i >= (long long)N * Hout * Wout * c8n) return;
    const int c8 = (int)(i % c8n);
    long long p = i / c8n;
    const int ox = (int)(p % Wout); p /= Wout;
    const int oy = (int)(p % Hout); const int n = (int)(p / Hout);
    __half2 m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) m[q] = __float2half2_rn(-65504.f);
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= Hin) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            if (ix < 0 || ix >= Win) continue;
            const uint4 v = *reinterpret_cast<const uint4*>(X + (((long long)n * Hin + iy) * Win + ix) * C + c8 * 8);
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int q = 0; q < 4; ++q) m[q] = __hmax2(m[q], h[q]);
        }
    }
    *reinterpret_cast<uint4*>(Y + i * 8) = *reinterpret_cast<const uint4*>(m);
}

}  // namespace egoego
