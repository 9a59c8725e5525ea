// fp32 CUDA-core kernels: the validation engine (EGOEGO_ENGINE_SIMT) and the element-wise /
// reduction kernels shared by both engines.  Everything here is fp32 SIMT, vectorised where the
// layout allows; the tensor-core engine lives in gemm_tcgen05.cuh / attn_tcgen05.cuh.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace egoego {

// ---------------------------------------------------------------------------------------------
// Epilogues (shared vocabulary of both engines).  Rows are padded-token rows: row = w*LP + l,
// l = 0 time token, 1..T frames, > T padding.
// ---------------------------------------------------------------------------------------------
struct EpiStart {            // Decoder.forward: start_conv + time token + positional rows
    float* H; int ldh;       // [M, d_model]
    const float* bias;       // start_conv.bias (nullptr if already folded into `base`)
    const float* base;       // optional [M, d_model] pre-computed x_cond contribution (incl. bias+pos)
    const float* pos;        // position table [max_timesteps+1, d_model]
    const float* temb;       // [timesteps, d_model]
    TSrc ts; int T;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const {
        int w = row / LP, l = row % LP;
        float v;
        if (l == 0)      v = temb[(long long)ts.get(w) * ldh + col] + pos[ldh + col];
        else if (l <= T) v = base ? acc + base[(long long)row * ldh + col]
                                  : acc + bias[col] + pos[(long long)(l + 1) * ldh + col];
        else             v = 0.f;
        H[(long long)row * ldh + col] = v;
    }
};

struct EpiBiasScale {        // QKV projection: (acc + bias) * (col < scale_cols ? scale : 1)
    float* C; int ldc; const float* bias; int scale_cols; float scale;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const {
        float v = acc + bias[col];
        if (col < scale_cols) v *= scale;
        C[(long long)row * ldc + col] = v;
    }
};

struct EpiBiasRelu {
    float* C; int ldc; const float* bias;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const {
        C[(long long)row * ldc + col] = fmaxf(acc + bias[col], 0.f);
    }
};

struct EpiBiasResid {        // fc / w_2: acc + bias + residual (pre-LayerNorm)
    float* C; int ldc; const float* bias; const float* res;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const {
        C[(long long)row * ldc + col] = acc + bias[col] + res[(long long)row * ldc + col];
    }
};

struct EpiBiasDropResid {    // training: dropout(acc + bias) + residual (transformer_module.py:92,113)
    float* C; int ldc; const float* bias; const float* res; DropCfg drop; uint32_t stream;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const {
        const float v = (acc + bias[col]) * drop_factor(drop, stream, (unsigned long long)row * 512ull + (unsigned long long)col);
        C[(long long)row * ldc + col] = v + res[(long long)row * ldc + col];
    }
};

// training step (train.cuh); defined here because the tensor-core products instantiate their fused epilogues in engine_tc.cu
struct EpiPlainBias {        // C = acc (+ bias): all rows, cols < N
    float* C; int ldc; const float* bias;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const { C[(long long)row * ldc + col] = acc + (bias ? bias[col] : 0.f); }
};
struct EpiAccum {            // C += acc
    float* C; int ldc;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const { C[(long long)row * ldc + col] += acc; }
};

struct EpiOut {              // linear_out on tokens 1..T -> compact [B,T,d_feats]
    float* out; int d_feats; const float* bias; int T;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const {
        int w = row / LP, l = row % LP;
        if (l >= 1 && l <= T && col < d_feats)
            out[((long long)w * T + (l - 1)) * d_feats + col] = acc + bias[col];
    }
};

// ---------------------------------------------------------------------------------------------
// Small-M SGEMM: same contract as sgemm_tn_kernel below with 32 x 32 x 32 tiles (2 x 2 per thread, M % 32 == 0, K % 32 == 0): the
// stage-1 sequence nets run one 128-row window at a time, where 128 x 128 tiles leave all but 2-24 of the 148 SMs idle and
// each CTA walks the whole K alone (latency-bound: ~37 us for K = 1024); 16x more, 16x lighter CTAs finish in a few us.
// The fmaf chain over k is in the same order as in sgemm_tn_kernel, so results are bit-identical to it.
// ---------------------------------------------------------------------------------------------
template <class Epi>
__global__ void __launch_bounds__(256) sgemm_tn_small_kernel(const float* __restrict__ A, int lda,
                                                             const float* __restrict__ W, int ldw,
                                                             int N, int K, Epi epi) {
    __shared__ __align__(16) float As[32][32 + 4];
    __shared__ __align__(16) float Bs[32][32 + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    const int r = tid / 8, k4 = (tid % 8) * 4;            // one float4 of A and of W per thread and k-block
    for (int k0 = 0; k0 < K; k0 += 32) {
        const float4 a = *reinterpret_cast<const float4*>(A + (long long)(m0 + r) * lda + k0 + k4);
        As[k4 + 0][r] = a.x; As[k4 + 1][r] = a.y; As[k4 + 2][r] = a.z; As[k4 + 3][r] = a.w;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + r < N) b = *reinterpret_cast<const float4*>(W + (long long)(n0 + r) * ldw + k0 + k4);
        Bs[k4 + 0][r] = b.x; Bs[k4 + 1][r] = b.y; Bs[k4 + 2][r] = b.z; Bs[k4 + 3][r] = b.w;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1], b0 = Bs[k][tx * 2], b1 = Bs[k][tx * 2 + 1];
            acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
            acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = n0 + tx * 2 + j;
            if (col < N) epi(m0 + ty * 2 + i, col, acc[i][j]);
        }
}

// ---------------------------------------------------------------------------------------------
// SGEMM: C[M,N] = A[M,K] * W[N,K]^T, fp32, 128x128x16 tiles, 8x8 per thread.
// Requires M % 128 == 0, K % 16 == 0, lda/ldw % 4 == 0; N arbitrary (guarded).
// ---------------------------------------------------------------------------------------------
template <class Epi>
__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ A, int lda,
                                                       const float* __restrict__ W, int ldw,
                                                       int N, int K, Epi epi) {
    __shared__ __align__(16) float As[16][128 + 4];
    __shared__ __align__(16) float Bs[16][128 + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int idx = tid + it * 256;        // 0..511
            int r = idx / 4, k4 = (idx % 4) * 4;
            float4 a = *reinterpret_cast<const float4*>(A + (long long)(m0 + r) * lda + k0 + k4);
            As[k4 + 0][r] = a.x; As[k4 + 1][r] = a.y; As[k4 + 2][r] = a.z; As[k4 + 3][r] = a.w;
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + r < N) b = *reinterpret_cast<const float4*>(W + (long long)(n0 + r) * ldw + k0 + k4);
            Bs[k4 + 0][r] = b.x; Bs[k4 + 1][r] = b.y; Bs[k4 + 2][r] = b.z; Bs[k4 + 3][r] = b.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
            *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (col < N) epi(row, col, acc[i][j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// SIMT attention for one (window, head): S = Q K^T (Q pre-scaled by 1/sqrt(d_k)), softmax over the
// first L keys, O = P V.  QKV is [M, 3*H*DH] fp32 (q | k | v blocks), O is [M, H*DH].
// DH fixed at 256 (d_k = d_v = 256); one block of 256 threads per (window, head).
// ---------------------------------------------------------------------------------------------
constexpr int ATT_SIMT_SMEM = (128 * 129 + 16 * 132 * 2) * 4 > (128 * 129 + 16 * 256) * 4
                                  ? (128 * 129 + 16 * 132 * 2) * 4 : (128 * 129 + 16 * 256) * 4;

template <bool SPLIT>
__global__ void __launch_bounds__(256) attention_simt_kernel(const float* __restrict__ QKV, int ldq,
                                                             float* __restrict__ O, __nv_bfloat16* __restrict__ Ohi,
                                                             __nv_bfloat16* __restrict__ Olo, int ldo,
                                                             int n_head, int L, DropCfg drop = DropCfg{0, 0, 1.0f, 0},
                                                             uint32_t drop_stream = 0) {
    constexpr int DH = 256;
    extern __shared__ __align__(16) float sm[];
    float (*S)[129] = reinterpret_cast<float (*)[129]>(sm);
    float* tile = sm + 128 * 129;
    const int w = blockIdx.x / n_head, h = blockIdx.x % n_head;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const float* Q = QKV + (long long)w * LP * ldq + h * DH;
    const float* Kp = Q + n_head * DH;
    const float* V = Q + 2 * n_head * DH;

    {   // phase 1: S[128][128]
        float (*As)[132] = reinterpret_cast<float (*)[132]>(tile);
        float (*Bs)[132] = reinterpret_cast<float (*)[132]>(tile + 16 * 132);
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < DH; k0 += 16) {
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                int idx = tid + it * 256;
                int r = idx / 4, k4 = (idx % 4) * 4;
                float4 a = *reinterpret_cast<const float4*>(Q + (long long)r * ldq + k0 + k4);
                As[k4 + 0][r] = a.x; As[k4 + 1][r] = a.y; As[k4 + 2][r] = a.z; As[k4 + 3][r] = a.w;
                float4 b = *reinterpret_cast<const float4*>(Kp + (long long)r * ldq + k0 + k4);
                Bs[k4 + 0][r] = b.x; Bs[k4 + 1][r] = b.y; Bs[k4 + 2][r] = b.z; Bs[k4 + 3][r] = b.w;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float a[8], b[8];
                *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
                *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int c = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                S[r][c] = acc[i][j];
            }
        }
    }
    __syncthreads();
    {   // phase 2: row softmax over keys [0, L)
        const int warp = tid / 32, lane = tid % 32;
        for (int r = warp; r < 128; r += 8) {
            float v[4], mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int c = lane + 32 * j;
                v[j] = (c < L) ? S[r][c] : -INFINITY;
                mx = fmaxf(mx, v[j]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { v[j] = (lane + 32 * j < L) ? expf(v[j] - mx) : 0.f; sum += v[j]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            float inv = 1.0f / sum;
#pragma unroll
            for (int j = 0; j < 4; ++j) S[r][lane + 32 * j] = v[j] * inv;
        }
    }
    __syncthreads();
    if (drop.on) {   // training: dropout on the attention probabilities (transformer_module.py:84), one Philox call per 4 keys
        const unsigned long long base = (unsigned long long)blockIdx.x * 128ull * 128ull;     // blockIdx.x = b * H + h
        for (int q = tid; q < 128 * 32; q += 256) {
            const int r = q >> 5, c4 = (q & 31) * 4;
            const uint4 wd = drop_words(drop, drop_stream, (base + (unsigned long long)r * 128ull + c4) >> 2);
            S[r][c4 + 0] *= wd.x < drop.thresh ? drop.scale : 0.f; S[r][c4 + 1] *= wd.y < drop.thresh ? drop.scale : 0.f;
            S[r][c4 + 2] *= wd.z < drop.thresh ? drop.scale : 0.f; S[r][c4 + 3] *= wd.w < drop.thresh ? drop.scale : 0.f;
        }
        __syncthreads();
    }
    {   // phase 3: O[128][256] = P[128][128] V[128][256]; thread owns 8 rows x 16 cols
        float (*Vs)[256] = reinterpret_cast<float (*)[256]>(tile);
        float acc[8][16];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
        for (int k0 = 0; k0 < 128; k0 += 16) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                int idx = tid + it * 256;      // 0..1023 float4
                int r = idx / 64, c4 = (idx % 64) * 4;
                *reinterpret_cast<float4*>(&Vs[r][c4]) =
                    *reinterpret_cast<const float4*>(V + (long long)(k0 + r) * ldq + c4);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                float a[8], b[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = S[ty * 8 + i][k0 + k];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(&b[q * 4]) = *reinterpret_cast<const float4*>(&Vs[k][q * 64 + tx * 4]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
        const long long obase = (long long)w * LP * ldo + h * DH;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const long long o = obase + (long long)(ty * 8 + i) * ldo + q * 64 + tx * 4;
                if (SPLIT) {
                    __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
                    split_f16(acc[i][q * 4 + 0], h0, l0); split_f16(acc[i][q * 4 + 1], h1, l1);
                    split_f16(acc[i][q * 4 + 2], h2, l2); split_f16(acc[i][q * 4 + 3], h3, l3);
                    __nv_bfloat162 a0(h0, h1), a1(h2, h3), b0(l0, l1), b1(l2, l3);
                    *reinterpret_cast<uint2*>(Ohi + o) = make_uint2(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1));
                    *reinterpret_cast<uint2*>(Olo + o) = make_uint2(*reinterpret_cast<uint32_t*>(&b0), *reinterpret_cast<uint32_t*>(&b1));
                } else {
                    *reinterpret_cast<float4*>(O + o) =
                        make_float4(acc[i][q * 4 + 0], acc[i][q * 4 + 1], acc[i][q * 4 + 2], acc[i][q * 4 + 3]);
                }
            }
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over d_model = 512 (eps 1e-5, biased variance), one warp per row, float4 loads.
// Optional row mask (padding_mask, [B, T+1]) multiplies the normalised row (DecoderLayer :135,139).
// Optionally also emits the fp16 hi/lo planes consumed by the tensor-core engine.
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) layernorm512_kernel(const float* __restrict__ Y, float* __restrict__ H,
                                                           __nv_bfloat16* __restrict__ Hhi,
                                                           __nv_bfloat16* __restrict__ Hlo,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const float* __restrict__ row_mask, int T, int M,
                                                           int half_fmt /* 1: write one fp16 plane into Hhi */) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");          // PDL (no-op without the launch attribute)
    const int row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
    if (row >= M) return;
    const float4* y4 = reinterpret_cast<const float4*>(Y + (long long)row * 512);
    float4 v[4];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = y4[lane + 32 * j]; s += v[j].x + v[j].y + v[j].z + v[j].w; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / 512.0f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        q += a * a + b * b + c * c + d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / 512.0f) + 1e-5f);
    float mk = 1.f;
    if (row_mask) { int w = row / LP, l = row % LP; mk = (l <= T) ? row_mask[(long long)w * (T + 1) + l] : 0.f; }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 g = g4[lane + 32 * j], b = b4[lane + 32 * j], o;
        o.x = ((v[j].x - mean) * rstd * g.x + b.x) * mk;
        o.y = ((v[j].y - mean) * rstd * g.y + b.y) * mk;
        o.z = ((v[j].z - mean) * rstd * g.z + b.z) * mk;
        o.w = ((v[j].w - mean) * rstd * g.w + b.w) * mk;
        reinterpret_cast<float4*>(H + (long long)row * 512)[lane + 32 * j] = o;
        if (Hhi && half_fmt) {
            __half2 a = __floats2half2_rn(o.x, o.y), b2 = __floats2half2_rn(o.z, o.w);
            uint2 ph;
            ph.x = *reinterpret_cast<uint32_t*>(&a); ph.y = *reinterpret_cast<uint32_t*>(&b2);
            reinterpret_cast<uint2*>(Hhi + (long long)row * 512)[lane + 32 * j] = ph;
        } else if (Hhi) {
            __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
            split_f16(o.x, h0, l0); split_f16(o.y, h1, l1); split_f16(o.z, h2, l2); split_f16(o.w, h3, l3);
            __nv_bfloat162 hh0(h0, h1), hh1(h2, h3), ll0(l0, l1), ll1(l2, l3);
            uint2 ph, pl;
            ph.x = *reinterpret_cast<uint32_t*>(&hh0); ph.y = *reinterpret_cast<uint32_t*>(&hh1);
            pl.x = *reinterpret_cast<uint32_t*>(&ll0); pl.y = *reinterpret_cast<uint32_t*>(&ll1);
            reinterpret_cast<uint2*>(Hhi + (long long)row * 512)[lane + 32 * j] = ph;
            reinterpret_cast<uint2*>(Hlo + (long long)row * 512)[lane + 32 * j] = pl;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Timestep-embedding table: temb[t] = Linear(256->512)(GELU(Linear(64->256)([sin(t f), cos(t f)])))
// for every t in [0, timesteps)  (transformer_cond_diffusion_model.py:61-73,111-116).
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) time_table_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                                         const float* __restrict__ w2, const float* __restrict__ b2,
                                                         float* __restrict__ temb, int d_model) {
    __shared__ float emb[64];
    __shared__ float hid[256];
    const int t = blockIdx.x, tid = threadIdx.x;
    if (tid < 32) {
        float f = expf((float)tid * -(logf(10000.0f) / 31.0f));
        float a = (float)t * f;
        emb[tid] = sinf(a);
        emb[tid + 32] = cosf(a);
    }
    __syncthreads();
    {
        float s = b1[tid];
        for (int k = 0; k < 64; ++k) s = fmaf(emb[k], w1[tid * 64 + k], s);
        hid[tid] = 0.5f * s * (1.0f + erff(s * 0.70710678118654752440f));
    }
    __syncthreads();
    for (int o = tid; o < d_model; o += 256) {
        float s = b2[o];
        for (int k = 0; k < 256; ++k) s = fmaf(hid[k], w2[o * 256 + k], s);
        temb[(long long)t * d_model + o] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Sampler element-wise kernels.  One thread per quad (4 consecutive elements of a window) so one
// Philox call feeds four outputs; windows have T*D elements (D = 198).
// ---------------------------------------------------------------------------------------------
// x = noise(draw 0)   and/or   x_cond = x_start*(1-m) + m*noise(draw 1)
static __global__ void init_sample_kernel(float* __restrict__ x, float* __restrict__ x_cond,
                                   const float* __restrict__ x_init, const float* __restrict__ x_start,
                                   const float* __restrict__ cond_mask, NoiseSrc ns, int B, int T, int D) {
    const long long epw = (long long)T * D;
    const long long quads_pw = (epw + 3) / 4;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= quads_pw * B) return;
    int w = (int)(gid / quads_pw);
    int e0 = (int)(gid % quads_pw) * 4;
    float n0[4], n1[4];
    if (ns.tape) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            n0[r] = (e0 + r < epw && !x_init) ? ns.tape[(long long)w * epw + e0 + r] : 0.f;
            n1[r] = (e0 + r < epw) ? ns.tape[ns.draw_stride + (long long)w * epw + e0 + r] : 0.f;
        }
    } else {
        float4 a = philox_normal4(ns.seed, ns.window_offset + w, 0u, (uint32_t)(e0 >> 2));
        float4 b = philox_normal4(ns.seed, ns.window_offset + w, 1u, (uint32_t)(e0 >> 2));
        n0[0] = a.x; n0[1] = a.y; n0[2] = a.z; n0[3] = a.w;
        n1[0] = b.x; n1[1] = b.y; n1[2] = b.z; n1[3] = b.w;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (e0 + r >= epw) break;
        long long i = (long long)w * epw + e0 + r;
        x[i] = x_init ? x_init[i] : n0[r];
        float m = cond_mask[i];
        x_cond[i] = x_start[i] * (1.0f - m) + m * n1[r];
    }
}

// Scatter compact [B,T,*] rows into the padded fp32 A operand of the SIMT start GEMM:
// Ain[w*LP + 1 + f][col0 + c] = src[w][f][c]; other rows / pad columns are zeroed once at setup.
static __global__ void stage_rows_f32_kernel(float* __restrict__ Ain, int lda, int col0,
                                      const float* __restrict__ src, int src_ld, int src_col0, int ncols,
                                      int B, int T) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * T * ncols;
    if (gid >= total) return;
    int c = (int)(gid % ncols);
    long long ft = gid / ncols;
    int f = (int)(ft % T), w = (int)(ft / T);
    Ain[((long long)w * LP + 1 + f) * lda + col0 + c] = src[((long long)w * T + f) * src_ld + src_col0 + c];
}

// Same, into fp16 hi/lo planes (tensor-core engine A operand).
static __global__ void stage_rows_split_kernel(__nv_bfloat16* __restrict__ Ahi, __nv_bfloat16* __restrict__ Alo,
                                        __half* __restrict__ A16 /* nullable: extra fp16 plane */, int lda,
                                        const float* __restrict__ src, int src_ld, int src_col0, int ncols,
                                        int B, int T) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)B * T * ncols;
    if (gid >= total) return;
    int c = (int)(gid % ncols);
    long long ft = gid / ncols;
    int f = (int)(ft % T), w = (int)(ft / T);
    __nv_bfloat16 hi, lo;
    split_f16(src[((long long)w * T + f) * src_ld + src_col0 + c], hi, lo);
    long long o = ((long long)w * LP + 1 + f) * lda + c;
    Ahi[o] = hi; Alo[o] = lo;
    if (A16) A16[o] = __float2half_rn(src[((long long)w * T + f) * src_ld + src_col0 + c]);
}

// DDPM update (p_mean_variance tail + p_sample, transformer_cond_diffusion_model.py:235-256) fused with
// the in-paint overwrite (:395-397) and the staging of the next step's GEMM A operand.

__device__ __forceinline__ void ddpm_update_body(const DdpmArgs& a) {
    // 32-bit index arithmetic throughout (the kernel is issue-bound: 64-bit divisions cost more than the Philox rounds);
    // grid = (quads of one window, windows)
    const int epw = a.T * a.D;
    const int w = blockIdx.y;
    const int e0 = (int)(blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (e0 >= epw) return;
    const int t = a.ts.get(w);
    const float c1 = a.coef1[t], c2 = a.coef2[t];
    const float sigma = (t == 0) ? 0.f : expf(0.5f * a.logvar[t]);
    const float k_mo = (a.objective == 0) ? -a.sqrt_recipm1[t] : 1.0f, k_x = (a.objective == 0) ? a.sqrt_recip[t] : 0.0f;
    const long long i0 = (long long)w * epw + e0;
    // float4 path: no tail quad and every window base 16-byte aligned (warp-uniform)
    const bool vec = (epw & 3) == 0 && (((uintptr_t)a.x | (uintptr_t)a.x_out | (uintptr_t)a.model_out) & 15) == 0;
    float nz[4], xv[4], mo[4];
    // the two HBM reads first: the Philox rounds and the Box-Muller below (a ~400-cycle dependent chain) then run under their latency
    // (SASS of the previous order: both LDG.128 were issued after the last MUFU)
    if (vec) {
        const float4 xq = *reinterpret_cast<const float4*>(a.x + i0), mq = *reinterpret_cast<const float4*>(a.model_out + i0);
        xv[0] = xq.x; xv[1] = xq.y; xv[2] = xq.z; xv[3] = xq.w; mo[0] = mq.x; mo[1] = mq.y; mo[2] = mq.z; mo[3] = mq.w;
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) { const bool ok = e0 + r < epw; xv[r] = ok ? a.x[i0 + r] : 0.f; mo[r] = ok ? a.model_out[i0 + r] : 0.f; }
    }
    const int draw = a.ns.draw();
    if (a.ns.tape) {
        const float* tp = a.ns.tape + (long long)draw * a.ns.draw_stride + i0;
        if (vec && ((a.ns.draw_stride & 3) == 0) && ((uintptr_t)a.ns.tape & 15) == 0) { const float4 n = *reinterpret_cast<const float4*>(tp); nz[0] = n.x; nz[1] = n.y; nz[2] = n.z; nz[3] = n.w; }
        else {
#pragma unroll
            for (int r = 0; r < 4; ++r) nz[r] = (e0 + r < epw) ? tp[r] : 0.f;
        }
    } else {
        float4 n = philox_normal4(a.ns.seed, a.ns.window_offset + w, (uint32_t)draw, (uint32_t)(e0 >> 2));
        nz[0] = n.x; nz[1] = n.y; nz[2] = n.z; nz[3] = n.w;
    }
    float v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float x0 = (a.objective == 0) ? (k_x * xv[r] + k_mo * mo[r]) : mo[r];
        if (a.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
        v[r] = (c1 * x0 + c2 * xv[r]) + sigma * nz[r];
    }
    // d_feats is even and e0 is a multiple of 4: the pairs (e0, e0+1) and (e0+2, e0+3) never straddle a frame
    const int f0 = (int)((unsigned)e0 / (unsigned)a.D), c0 = e0 - f0 * a.D;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int e = e0 + 2 * h;
        if (e >= epw) break;
        const bool wrap = c0 + 2 * h >= a.D;             // the second pair may start the next frame (d_feats is even)
        const int f = f0 + (wrap ? 1 : 0), c = c0 + 2 * h - (wrap ? a.D : 0);
        if (a.inpaint && f < a.inpaint_len) {
            const float* ip = a.inpaint + ((long long)w * a.inpaint_len + f) * a.D + c;
            v[2 * h] = ip[0]; v[2 * h + 1] = ip[1];
        }
        if (!vec) { a.x_out[i0 + 2 * h] = v[2 * h]; a.x_out[i0 + 2 * h + 1] = v[2 * h + 1]; }
        if (a.stage_f32) *reinterpret_cast<float2*>(a.stage_f32 + ((long long)w * LP + 1 + f) * a.stage_ld + c) = make_float2(v[2 * h], v[2 * h + 1]);
        if (a.stage_hi) {
            const long long o = ((long long)w * LP + 1 + f) * a.stage_ld16 + c;
            if (a.stage_mode != 1) {
                __nv_bfloat16 h0, l0, h1, l1;
                split_f16(v[2 * h], h0, l0); split_f16(v[2 * h + 1], h1, l1);
                *reinterpret_cast<__nv_bfloat162*>(a.stage_hi + o) = __nv_bfloat162(h0, h1);
                *reinterpret_cast<__nv_bfloat162*>(a.stage_lo + o) = __nv_bfloat162(l0, l1);
            }
            if (a.stage_h16 && a.stage_mode != 0) *reinterpret_cast<__half2*>(a.stage_h16 + o) = __floats2half2_rn(v[2 * h], v[2 * h + 1]);
        }
    }
    if (vec) *reinterpret_cast<float4*>(a.x_out + i0) = make_float4(v[0], v[1], v[2], v[3]);
}

static __global__ void ddpm_update_kernel(DdpmArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");          // PDL: everything below reads the previous kernel's output
    ddpm_update_body(a);
    if (a.advance) {
        // Sampling loop: the LAST block to finish advances the device step counter (every block has read it by then), which saves
        // the separate one-thread advance_step_kernel node of every step.
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(a.done, 1u) == gridDim.x * gridDim.y - 1u) { *a.done = 0u; __threadfence(); (*a.advance)++; }
        }
    }
}

static __global__ void advance_step_kernel(int* d_step) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) (*d_step)++;
}

}  // namespace egoego
