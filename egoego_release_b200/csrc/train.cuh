// Training step of the stage-2 denoiser on the device (BASELINE config 5): forward with saved
// activations, L1/L2 loss with the padding mask, and the full backward pass -- gradients of all 72 parameter tensors.
//
// Follows CondGaussianDiffusion.forward / p_losses / q_sample (egoego/model/transformer_cond_diffusion_model.py:557-625) and the
// modules they run (transformer_module.py:61-142,188-226; time_mlp :105-116).  All 53 matrix products of a step run on the
// tensor cores through tc_gemm_f32 (3-term bf16 split, fp32-grade; engine_tc.cu) -- EGOEGO_TRAIN_GEMM=simt keeps the fp32
// CUDA-core sgemm for bisecting; operands of the weight-gradient products are transposed explicitly; attention forward /
// backward (the backward recomputes P) are fp32 CUDA-core kernels.
// Dropout is not applied (identity): the parity bar of this row is the reference with its modules in eval() mode
// (oracle/training.py) -- torch's dropout stream cannot be reproduced outside torch.
//
// Included at the end of egoego_b200.cu (uses egoego_ctx and its packed fp32 weights; SIMT-engine handles only).
#pragma once

namespace egoego {

// [R, C] (ld) -> [C, R] (ldo), 32 x 32 tiles
static __global__ void tr_transpose_kernel(const float* __restrict__ src, int R, int C, int ld, float* __restrict__ dst, int ldo) {
    __shared__ float t[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        t[i][threadIdx.x] = (r < R && c < C) ? src[(long long)r * ld + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < C && r < R) dst[(long long)c * ldo + r] = t[threadIdx.x][i];
    }
}

// q_sample + x_cond (:557-563,581-586) -> rows 1..T of the start_conv operand [M, kin_pad] (x | x_cond)
static __global__ void tr_prep_kernel(const float* __restrict__ x0, const float* __restrict__ mask, const float* __restrict__ noise,
                                      const float* __restrict__ cnoise, const float* __restrict__ sa, const float* __restrict__ sb,
                                      float* __restrict__ Ain, int lda, int B, int T, int D) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * T * D) return;
    const int c = (int)(i % D);
    const long long ft = i / D;
    const int f = (int)(ft % T), b = (int)(ft / T);
    const float x = sa[b] * x0[i] + sb[b] * noise[i];
    const float m = mask[i];
    const float xc = x0[i] * (1.0f - m) + m * cnoise[i];
    float* row = Ain + ((long long)b * LP + 1 + f) * lda;
    row[c] = x; row[D + c] = xc;
}

// LayerNorm(512) forward that keeps mean / rstd for the backward pass; rows multiplied by the padding mask
static __global__ void __launch_bounds__(256) tr_ln_fwd_kernel(const float* __restrict__ Y, float* __restrict__ H, float* __restrict__ stats /*[M,2]*/,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const float* __restrict__ row_mask, int T, int M) {
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
    for (int j = 0; j < 4; ++j) { const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean; q += a * a + b * b + c * c + d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / 512.0f) + 1e-5f);
    if (lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
    const int w = row / LP, l = row % LP;
    const float mk = (l <= T) ? (row_mask ? row_mask[(long long)w * (T + 1) + l] : 1.f) : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * j], b = reinterpret_cast<const float4*>(beta)[lane + 32 * j];
        float4 o;
        o.x = ((v[j].x - mean) * rstd * g.x + b.x) * mk; o.y = ((v[j].y - mean) * rstd * g.y + b.y) * mk;
        o.z = ((v[j].z - mean) * rstd * g.z + b.z) * mk; o.w = ((v[j].w - mean) * rstd * g.w + b.w) * mk;
        reinterpret_cast<float4*>(H + (long long)row * 512)[lane + 32 * j] = o;
    }
}

// LayerNorm backward for one row per warp: dY = rstd (g dH - mean(g dH) - xhat mean(g dH xhat)), with dH already masked here;
// also writes xhat * dH and dH (masked) into scratch for the parameter reductions
static __global__ void __launch_bounds__(256) tr_ln_bwd_kernel(const float* __restrict__ dH, const float* __restrict__ Y, const float* __restrict__ stats,
                                                               const float* __restrict__ gamma, const float* __restrict__ row_mask, int T, int M,
                                                               float* __restrict__ dY, float* __restrict__ dgam_part /*[M,512] = dH*xhat*/,
                                                               float* __restrict__ dbet_part /*[M,512] = dH masked*/) {
    const int row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
    if (row >= M) return;
    const int w = row / LP, l = row % LP;
    const float mk = (l <= T) ? (row_mask ? row_mask[(long long)w * (T + 1) + l] : 1.f) : 0.f;
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float4 xh[4], gd[4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 y = reinterpret_cast<const float4*>(Y + (long long)row * 512)[lane + 32 * j];
        float4 d = reinterpret_cast<const float4*>(dH + (long long)row * 512)[lane + 32 * j];
        const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * j];
        d.x *= mk; d.y *= mk; d.z *= mk; d.w *= mk;
        xh[j] = make_float4((y.x - mean) * rstd, (y.y - mean) * rstd, (y.z - mean) * rstd, (y.w - mean) * rstd);
        reinterpret_cast<float4*>(dbet_part + (long long)row * 512)[lane + 32 * j] = d;
        reinterpret_cast<float4*>(dgam_part + (long long)row * 512)[lane + 32 * j] = make_float4(d.x * xh[j].x, d.y * xh[j].y, d.z * xh[j].z, d.w * xh[j].w);
        gd[j] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        s1 += gd[j].x + gd[j].y + gd[j].z + gd[j].w;
        s2 += gd[j].x * xh[j].x + gd[j].y * xh[j].y + gd[j].z * xh[j].z + gd[j].w * xh[j].w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const float m1 = s1 * (1.0f / 512.0f), m2 = s2 * (1.0f / 512.0f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 o;
        o.x = rstd * (gd[j].x - m1 - xh[j].x * m2); o.y = rstd * (gd[j].y - m1 - xh[j].y * m2);
        o.z = rstd * (gd[j].z - m1 - xh[j].z * m2); o.w = rstd * (gd[j].w - m1 - xh[j].w * m2);
        reinterpret_cast<float4*>(dY + (long long)row * 512)[lane + 32 * j] = o;
    }
}

// column sums of X [M, C] (ld): two deterministic stages -- partial[rb][c] over TR_CS_RB row blocks (enough CTAs to fill the
// GPU: a bias gradient reduces 4096 rows x 512..3072 columns), then the sum over the row blocks
constexpr int TR_CS_RB = 64;
static __global__ void tr_colsum_part_kernel(const float* __restrict__ X, int M, int C, int ld, float* __restrict__ part /*[TR_CS_RB, C]*/) {
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x, rb = blockIdx.y;
    const int r0 = (int)((long long)M * rb / TR_CS_RB), r1 = (int)((long long)M * (rb + 1) / TR_CS_RB);
    float s = 0.f;
    if (c < C) for (int r = r0 + threadIdx.y; r < r1; r += 8) s += X[(long long)r * ld + c];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
        part[(long long)rb * C + c] = t;
    }
}
static __global__ void tr_colsum_final_kernel(const float* __restrict__ part, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float t = 0.f;
    for (int rb = 0; rb < TR_CS_RB; ++rb) t += part[(long long)rb * C + c];
    out[c] = t;
}

static __global__ void tr_relu_bwd_kernel(float* __restrict__ dF, const float* __restrict__ F, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && F[i] <= 0.f) dF[i] = 0.f;
}
static __global__ void tr_add_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}
// keep rows 1..T of every window, zero the time-token row and the padding rows (start_conv / linear_out see frames only)
static __global__ void tr_frame_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int T, int M, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)M * C) return;
    const int l = (int)((i / C) % LP);
    dst[i] = (l >= 1 && l <= T) ? src[i] : 0.f;
}

// loss (:588-605) and d loss / d model_out.  out [M, ldo] (rows = padded tokens), target x0 or noise [B,T,D].
// loss = mean_b( mean_{f,c}( err * pm[b,f+1] ) * w_b );  L1: err = |o - y|, d = sign;  L2: err = (o - y)^2, d = 2 (o - y)
static __global__ void tr_loss_kernel(const float* __restrict__ out, int ldo, const float* __restrict__ target, const float* __restrict__ row_mask,
                                      const float* __restrict__ wgt, int l2, int B, int T, int D, double* __restrict__ loss, float* __restrict__ dOut) {
    __shared__ double red[256];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double my = 0.0;
    if (i < (long long)B * T * D) {
        const int c = (int)(i % D);
        const long long ft = i / D;
        const int f = (int)(ft % T), b = (int)(ft / T);
        const long long row = (long long)b * LP + 1 + f;
        const float diff = out[row * ldo + c] - target[i];
        const float pm = row_mask ? row_mask[(long long)b * (T + 1) + 1 + f] : 1.f;
        const float k = pm * wgt[b] / ((float)T * (float)D * (float)B);
        my = (double)((l2 ? diff * diff : fabsf(diff)) * k);
        dOut[row * ldo + c] = (l2 ? 2.f * diff : (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f))) * k;
    }
    red[threadIdx.x] = my;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) atomicAdd(loss, red[0]);
}

// timestep embedding of the B timesteps of this batch only (the sampler's table has all `timesteps` rows; rebuilding it after every
// optimizer step cost 0.6 ms): temb_b[b] = Linear(256->512)(GELU(Linear(64->256)([sin(t f), cos(t f)])))  (:61-73,105-116)
static __global__ void __launch_bounds__(256) tr_time_fwd_kernel(const long long* __restrict__ t_arr, const float* __restrict__ w1, const float* __restrict__ b1,
                                                                 const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ temb_b, int d_model) {
    __shared__ float emb[64], hid[256];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float t = (float)t_arr[b];
    if (tid < 32) { const float f = expf((float)tid * -(logf(10000.0f) / 31.0f)); emb[tid] = sinf(t * f); emb[tid + 32] = cosf(t * f); }
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
        temb_b[(long long)b * d_model + o] = s;
    }
}
static __global__ void tr_iota_kernel(long long* __restrict__ p, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = i; }

static __global__ void tr_loss_finish_kernel(const double* __restrict__ acc, float* __restrict__ out) { *out = (float)*acc; }

// time MLP backward (:105-116,122-123): d temb[b] = dH0[row b*128]; recompute emb / pre / gelu from t; accumulate parameter grads
static __global__ void __launch_bounds__(256) tr_time_bwd_kernel(const float* __restrict__ dH0, const long long* __restrict__ t_arr,
                                                                 const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                                                 float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2, int d_model) {
    __shared__ float emb[64], pre[256], g[256], dg[256], dt[512];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float t = (float)t_arr[b];
    if (tid < 32) { const float f = expf((float)tid * -(logf(10000.0f) / 31.0f)); emb[tid] = sinf(t * f); emb[tid + 32] = cosf(t * f); }
    for (int o = tid; o < d_model; o += 256) dt[o] = dH0[((long long)b * LP) * d_model + o];
    __syncthreads();
    {
        float s = b1[tid];
        for (int k = 0; k < 64; ++k) s = fmaf(emb[k], w1[tid * 64 + k], s);
        pre[tid] = s;
        g[tid] = 0.5f * s * (1.0f + erff(s * 0.70710678118654752440f));
    }
    __syncthreads();
    for (int o = tid; o < d_model; o += 256) {          // second Linear: dW2[o][k] += dt[o] g[k], db2[o] += dt[o]
        atomicAdd(&db2[o], dt[o]);
        for (int k = 0; k < 256; ++k) atomicAdd(&dw2[o * 256 + k], dt[o] * g[k]);
    }
    {
        float s = 0.f;                                   // dg[tid] = sum_o w2[o][tid] dt[o]
        for (int o = 0; o < d_model; ++o) s = fmaf(w2[o * 256 + tid], dt[o], s);
        const float x = pre[tid];
        const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
        const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
        dg[tid] = s * (cdf + x * pdf);                   // d gelu(x) / dx (exact erf form)
    }
    __syncthreads();
    atomicAdd(&db1[tid], dg[tid]);
    for (int k = 0; k < 64; ++k) atomicAdd(&dw1[tid * 64 + k], dg[tid] * emb[k]);
}

// ---- attention backward for one (window, head), recomputing P from the saved (scaled) Q, K --------------------------------
// QKV [M, ldq] (q | k | v blocks, head slices of 256, q pre-scaled by `scale`), dO [M, ldo]; writes dQKV (same layout as QKV):
//   dV = P^T dO,  dP = dO V^T,  dS = P (dP - rowsum(dP P)),  d q_unscaled = (dS K) scale,  dK = dS^T Q_scaled
constexpr int ATT_BWD_SMEM = (2 * 128 * 129 + 2 * 16 * 132 > 2 * 128 * 129 + 16 * 256 ? 2 * 128 * 129 + 2 * 16 * 132 : 2 * 128 * 129 + 16 * 256) * 4;

__device__ __forceinline__ void tr_tile_qk(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, float* tile, float (&acc)[8][8], int tid, int tx, int ty) {
    // acc[i][j] = sum_k A[row_i][k] Bm[row_j][k], k over 256 dims, rows 0..127 of both
    constexpr int DH = 256;
    float (*As)[132] = reinterpret_cast<float (*)[132]>(tile);
    float (*Bs)[132] = reinterpret_cast<float (*)[132]>(tile + 16 * 132);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < DH; k0 += 16) {
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = tid + it * 256, r = idx / 4, k4 = (idx % 4) * 4;
            const float4 a = *reinterpret_cast<const float4*>(A + (long long)r * lda + k0 + k4);
            As[k4 + 0][r] = a.x; As[k4 + 1][r] = a.y; As[k4 + 2][r] = a.z; As[k4 + 3][r] = a.w;
            const float4 b = *reinterpret_cast<const float4*>(Bm + (long long)r * ldb + k0 + k4);
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
}

// out[r][c] (128 x 256) = sum_k W[k][r] (TRANS) or W[r][k] (!TRANS) times X[k][c], k over 128 rows; thread owns 8 rows x 16 cols
template <bool TRANS>
__device__ __forceinline__ void tr_tile_pv(const float (*W)[129], const float* __restrict__ X, int ld, float* tile, float (&acc)[8][16], int tid, int tx, int ty) {
    float (*Xs)[256] = reinterpret_cast<float (*)[256]>(tile);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < 128; k0 += 16) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int idx = tid + it * 256, r = idx / 64, c4 = (idx % 64) * 4;
            *reinterpret_cast<float4*>(&Xs[r][c4]) = *reinterpret_cast<const float4*>(X + (long long)(k0 + r) * ld + c4);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[8], b[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = TRANS ? W[k0 + k][ty * 8 + i] : W[ty * 8 + i][k0 + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(&b[q * 4]) = *reinterpret_cast<const float4*>(&Xs[k][q * 64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void tr_store_pv(float* __restrict__ dst, int ld, const float (&acc)[8][16], float s, int tx, int ty) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(dst + (long long)(ty * 8 + i) * ld + q * 64 + tx * 4) =
                make_float4(acc[i][q * 4] * s, acc[i][q * 4 + 1] * s, acc[i][q * 4 + 2] * s, acc[i][q * 4 + 3] * s);
}

static __global__ void __launch_bounds__(256) attention_bwd_simt_kernel(const float* __restrict__ QKV, int ldq, const float* __restrict__ dO, int ldo,
                                                                        float* __restrict__ dQKV, int n_head, int L, float scale,
                                                                        DropCfg drop, uint32_t drop_stream) {
    constexpr int DH = 256;
    extern __shared__ __align__(16) float sm[];
    float (*P)[129] = reinterpret_cast<float (*)[129]>(sm);
    float (*Dm)[129] = reinterpret_cast<float (*)[129]>(sm + 128 * 129);
    float* tile = sm + 2 * 128 * 129;
    const int w = blockIdx.x / n_head, h = blockIdx.x % n_head;
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const float* Q = QKV + (long long)w * LP * ldq + h * DH;
    const float* Kp = Q + n_head * DH;
    const float* V = Q + 2 * n_head * DH;
    const float* dOh = dO + (long long)w * LP * ldo + h * DH;
    float* dQ = dQKV + (long long)w * LP * ldq + h * DH;
    float* dK = dQ + n_head * DH;
    float* dV = dQ + 2 * n_head * DH;
    {   // S = Q K^T -> P (softmax over keys < L); rows >= L get zeros (their dO is zero anyway)
        float acc[8][8];
        tr_tile_qk(Q, ldq, Kp, ldq, tile, acc, tid, tx, ty);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) P[i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4)][j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)] = acc[i][j];
    }
    __syncthreads();
    {
        const int warp = tid / 32, lane = tid % 32;
        for (int r = warp; r < 128; r += 8) {
            float v[4], mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int c = lane + 32 * j; v[j] = (c < L) ? P[r][c] : -INFINITY; mx = fmaxf(mx, v[j]); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) { v[j] = (lane + 32 * j < L) ? expf(v[j] - mx) : 0.f; sum += v[j]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = (r < L) ? 1.0f / sum : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) P[r][lane + 32 * j] = v[j] * inv;
        }
    }
    __syncthreads();
    {   // dP = dO V^T -> Dm
        float acc[8][8];
        tr_tile_qk(dOh, ldo, V, ldq, tile, acc, tid, tx, ty);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) Dm[i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4)][j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)] = acc[i][j];
    }
    __syncthreads();
    // dropout on the probabilities (forward: O = (P o m) V): d(P o m) = dO V^T, so dP = (dO V^T) o m -- the same Philox words
    // as the forward kernel; the softmax backward below uses the UN-dropped P, dV further down the dropped one
    auto drop_pass = [&](float (*X)[129]) {
        const unsigned long long base = (unsigned long long)blockIdx.x * 128ull * 128ull;
        for (int q = tid; q < 128 * 32; q += 256) {
            const int r = q >> 5, c4 = (q & 31) * 4;
            const uint4 wd = drop_words(drop, drop_stream, (base + (unsigned long long)r * 128ull + c4) >> 2);
            X[r][c4 + 0] *= wd.x < drop.thresh ? drop.scale : 0.f; X[r][c4 + 1] *= wd.y < drop.thresh ? drop.scale : 0.f;
            X[r][c4 + 2] *= wd.z < drop.thresh ? drop.scale : 0.f; X[r][c4 + 3] *= wd.w < drop.thresh ? drop.scale : 0.f;
        }
        __syncthreads();
    };
    if (drop.on) drop_pass(Dm);
    {   // dS = P (dP - rowsum(dP P)) in place in Dm
        const int warp = tid / 32, lane = tid % 32;
        for (int r = warp; r < 128; r += 8) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) s += Dm[r][lane + 32 * j] * P[r][lane + 32 * j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
#pragma unroll
            for (int j = 0; j < 4; ++j) Dm[r][lane + 32 * j] = P[r][lane + 32 * j] * (Dm[r][lane + 32 * j] - s);
        }
    }
    __syncthreads();
    if (drop.on) drop_pass(P);                                                                            // P o m for dV
    float acc[8][16];
    tr_tile_pv<true>(P, dOh, ldo, tile, acc, tid, tx, ty);    tr_store_pv(dV, ldq, acc, 1.0f, tx, ty);     // dV = (P o m)^T dO
    tr_tile_pv<false>(Dm, Kp, ldq, tile, acc, tid, tx, ty);   tr_store_pv(dQ, ldq, acc, scale, tx, ty);    // d q_unscaled = (dS K) scale
    tr_tile_pv<true>(Dm, Q, ldq, tile, acc, tid, tx, ty);     tr_store_pv(dK, ldq, acc, 1.0f, tx, ty);     // dK = dS^T Q_scaled
}

// dZ = dY o m: gradient through dropout(acc + bias) of the fc / FFN output (the residual branch keeps dY)
static __global__ void tr_dropout_bwd_kernel(const float* __restrict__ dY, float* __restrict__ dZ, long long n, DropCfg drop, uint32_t stream) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // quads of 4 consecutive channels
    if (q * 4 >= n) return;
    const uint4 wd = drop_words(drop, stream, (unsigned long long)q);
    float4 v = *reinterpret_cast<const float4*>(dY + q * 4);
    v.x *= wd.x < drop.thresh ? drop.scale : 0.f; v.y *= wd.y < drop.thresh ? drop.scale : 0.f;
    v.z *= wd.z < drop.thresh ? drop.scale : 0.f; v.w *= wd.w < drop.thresh ? drop.scale : 0.f;
    *reinterpret_cast<float4*>(dZ + q * 4) = v;
}

}  // namespace egoego
