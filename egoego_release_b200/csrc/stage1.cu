// Stage-1 networks of EgoEgo on the device, shipped configuration
// (scripts/test_egoego_pipeline.sh: --input_of_feats): a "sequence net" = Decoder (no leading token, row padding mask,
// full attention) + up to two MLP heads, plus the small geometric kernels around it.
//
//   HeadFormer        egoego/model/head_estimation_transformer.py:214-308 (forward_for_eval), :102-124 (va2rot),
//                     :184-212 (cal_scale_for_slam_w_pred_scale)
//   HeadNormalFormer  egoego/model/head_normal_estimation_transformer.py:118-165 (forward), :47-62 (rotation from the floor
//                     normal), :219-250 (rotation + scale applied to the SLAM trajectory)
//   Decoder / MLP     egoego/model/transformer_module.py:119-142,172-226; egoego/model/mlp.py:4-27
//
// These workloads are tiny (batch 1, <= 120 tokens, d_model 256): latency-bound, so everything runs on the fp32 CUDA-core
// kernels of the validation engine (sgemm_tn_kernel / attention_simt_kernel) with a LayerNorm for d_model = 256 -- no
// tensor-core reshaping.  Rows are padded tokens: row = window_index * 128 + position.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

#include "common.cuh"
#include "kernels_simt.cuh"
#include "postprocess.cuh"

namespace egoego {

// ---- epilogues -----------------------------------------------------------------------------------
struct S1EpiStart {          // start_conv + positional rows 1..window (every position, padded or not); rows >= window: 0
    float* H; int ld; const float* bias; const float* pos; int window;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const {
        const int l = row % LP;
        H[(long long)row * ld + col] = l < window ? acc + bias[col] + pos[(long long)(l + 1) * ld + col] : 0.f;
    }
};
struct S1EpiHeadOut {        // head's final Linear -> compact [B, n_out_rows, out_dim] (n_out_rows = T, or 1 for token 0 only)
    float* out; int out_dim; const float* bias; int n_rows;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const {
        const int w = row / LP, l = row % LP;
        if (l < n_rows && col < out_dim) out[((long long)w * n_rows + l) * out_dim + col] = acc + bias[col];
    }
};

// feats[B,T,D] -> Ain[w*LP + l][0..D) (everything else of the B windows zeroed by the caller)
static __global__ void s1_stage_rows_kernel(float* __restrict__ Ain, int lda, const float* __restrict__ src, int D, int B, int T) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * T * D) return;
    const int c = (int)(gid % D);
    const long long ft = gid / D;
    const int f = (int)(ft % T), w = (int)(ft / T);
    Ain[((long long)w * LP + f) * lda + c] = src[gid];
}

// LayerNorm over DM (multiple of 128, <= 1024) columns, eps 1e-5, biased variance; one warp per row; rows at positions
// >= n_valid[w] (padding) are zeroed after normalisation (DecoderLayer :135,139)
template <int DM>
static __global__ void __launch_bounds__(256) s1_layernorm_kernel(const float* __restrict__ Y, float* __restrict__ H,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const int* __restrict__ n_valid, int M) {
    constexpr int V = DM / 128;
    const int row = blockIdx.x * 8 + threadIdx.x / 32, lane = threadIdx.x % 32;
    if (row >= M) return;
    const float4* y4 = reinterpret_cast<const float4*>(Y + (long long)row * DM);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) { v[j] = y4[lane + 32 * j]; s += v[j].x + v[j].y + v[j].z + v[j].w; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / DM);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        q += a * a + b * b + c * c + d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / DM) + 1e-5f);
    const float mk = (row % LP) < n_valid[row / LP] ? 1.f : 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const float4 g = reinterpret_cast<const float4*>(gamma)[lane + 32 * j], b = reinterpret_cast<const float4*>(beta)[lane + 32 * j];
        float4 o;
        o.x = ((v[j].x - mean) * rstd * g.x + b.x) * mk; o.y = ((v[j].y - mean) * rstd * g.y + b.y) * mk;
        o.z = ((v[j].z - mean) * rstd * g.z + b.z) * mk; o.w = ((v[j].w - mean) * rstd * g.w + b.w) * mk;
        reinterpret_cast<float4*>(H + (long long)row * DM)[lane + 32 * j] = o;
    }
}

// ---- geometry around the networks ------------------------------------------------------------------
// va2rot (head_estimation_transformer.py:102-124): q_{t+1} = normalise(quat((q_t w_t q_t^-1) dt) * q_t); one thread per sequence
static __global__ void s1_va2rot_kernel(const float* __restrict__ q0, const float* __restrict__ va, int B, int T, float dt,
                                        float* __restrict__ out /* [B, T+1, 4] */) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    Q4 q = {q0[b * 4], q0[b * 4 + 1], q0[b * 4 + 2], q0[b * 4 + 3]};
    float* o = out + (long long)b * (T + 1) * 4;
    o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z;
    for (int t = 0; t < T; ++t) {
        const float w3[3] = {va[((long long)b * T + t) * 3], va[((long long)b * T + t) * 3 + 1], va[((long long)b * T + t) * 3 + 2]};
        float angv[3];
        q_apply(q, w3, angv);
        const float aa[3] = {angv[0] * dt, angv[1] * dt, angv[2] * dt};
        const Q4 n = q_mul(aa_to_q(aa), q);                              // quaternion_multiply standardises to w >= 0
        const float nn = sqrtf(n.w * n.w + n.x * n.x + n.y * n.y + n.z * n.z);
        q.w = n.w / nn; q.x = n.x / nn; q.y = n.y / nn; q.z = n.z / nn;
        o[(t + 1) * 4] = q.w; o[(t + 1) * 4 + 1] = q.x; o[(t + 1) * 4 + 2] = q.y; o[(t + 1) * 4 + 3] = q.z;
    }
}

// cal_scale_for_slam_w_pred_scale (:184-212) for one sequence: scale = mean(dist[:n]) / mean(|slam step|[:n]),
// n = min(P - 1, n_dist); out[0] = slam[0], out[t+1] = out[t] + scale * (slam[t+1] - slam[t]) (sequential, as the reference)
static __global__ void __launch_bounds__(256) s1_rescale_kernel(const float* __restrict__ slam /* [P,3] */, int P,
                                                                const float* __restrict__ dist, int n_dist, float dist_scale,
                                                                float* __restrict__ out /* [P,3] */, float* __restrict__ scale_out) {
    __shared__ float red[2][256];
    const int n = min(P - 1, n_dist), tid = threadIdx.x;
    float sd = 0.f, sl = 0.f;
    for (int t = tid; t < n; t += 256) {
        sd += dist[t] / dist_scale;
        const float dx = slam[(t + 1) * 3] - slam[t * 3], dy = slam[(t + 1) * 3 + 1] - slam[t * 3 + 1], dz = slam[(t + 1) * 3 + 2] - slam[t * 3 + 2];
        sl += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    red[0][tid] = sd; red[1][tid] = sl;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (tid < o) { red[0][tid] += red[0][tid + o]; red[1][tid] += red[1][tid + o]; } __syncthreads(); }
    if (tid == 0) {
        const float scale = (red[0][0] / (float)n) / (red[1][0] / (float)n);
        *scale_out = scale;
        float c[3] = {slam[0], slam[1], slam[2]};
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
        for (int t = 0; t + 1 < P; ++t) {
#pragma unroll
            for (int k = 0; k < 3; ++k) { c[k] = c[k] + scale * (slam[(t + 1) * 3 + k] - slam[t * 3 + k]); out[(t + 1) * 3 + k] = c[k]; }
        }
    }
}

// SLAM features (head_normal_estimation_transformer.py:128-137): rot_mat[B,P,9], trans[B,P,3] (first n_pose poses used)
// -> feats[B, n_pose-1, 18] = rot6d(R_t) | trans_t | rot6d(R_{t+1} R_t^T) | trans_{t+1} - trans_t
static __global__ void s1_slam_features_kernel(const float* __restrict__ rot, const float* __restrict__ trans, int B, int P, int n_pose,
                                               float* __restrict__ feats) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, T = n_pose - 1;
    if (gid >= B * T) return;
    const int b = gid / T, t = gid % T;
    const float* R0 = rot + ((long long)b * P + t) * 9;
    const float* R1 = R0 + 9;
    const float* p0 = trans + ((long long)b * P + t) * 3;
    float* f = feats + (long long)gid * 18;
#pragma unroll
    for (int k = 0; k < 6; ++k) f[k] = R0[k];                           // matrix_to_rotation_6d = first two rows
#pragma unroll
    for (int k = 0; k < 3; ++k) { f[6 + k] = p0[k]; f[15 + k] = p0[3 + k] - p0[k]; }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) f[9 + i * 3 + j] = R1[i * 3] * R0[j * 3] + R1[i * 3 + 1] * R0[j * 3 + 1] + R1[i * 3 + 2] * R0[j * 3 + 2];
}

// rotation taking the predicted floor normal to +z (:47-62, fp64 Rodrigues form), applied with `scale` to the frame-to-frame
// SLAM translations (re-integrated from pose 0, sequentially) and to the SLAM rotations (:219-250); one block per sequence
static __global__ void __launch_bounds__(128) s1_apply_normal_kernel(const float* __restrict__ normal /* [B,3] */, const float* __restrict__ scale /* [B] */,
                                                                     const float* __restrict__ rot /* [B,P,9] */, const float* __restrict__ trans /* [B,P,3] */, int P,
                                                                     float* __restrict__ trans_out, float* __restrict__ rot_out, float* __restrict__ quat_out,
                                                                     float* __restrict__ align_out /* [B,9] nullable */) {
    __shared__ float Ra[9];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        double a[3] = {normal[b * 3], normal[b * 3 + 1], normal[b * 3 + 2]};
        const double na = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        a[0] /= na; a[1] /= na; a[2] /= na;
        const double v[3] = {a[1], -a[0], 0.0};                        // a x (0,0,1)
        const double c = a[2], s2 = v[0] * v[0] + v[1] * v[1];
        const double K[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
        const double f = (1.0 - c) / s2;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double kk = 0.0;
                for (int k = 0; k < 3; ++k) kk += K[i * 3 + k] * K[k * 3 + j];
                Ra[i * 3 + j] = (float)((i == j ? 1.0 : 0.0) + K[i * 3 + j] + kk * f);
            }
        if (align_out) for (int k = 0; k < 9; ++k) align_out[b * 9 + k] = Ra[k];
    }
    __syncthreads();
    const float* Rb = rot + (long long)b * P * 9;
    const float* tb = trans + (long long)b * P * 3;
    for (int t = tid; t < P; t += blockDim.x) {                        // rotations: independent per pose
        M3 m;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) m.m[i * 3 + j] = Ra[i * 3] * Rb[t * 9 + j] + Ra[i * 3 + 1] * Rb[t * 9 + 3 + j] + Ra[i * 3 + 2] * Rb[t * 9 + 6 + j];
        if (rot_out) for (int k = 0; k < 9; ++k) rot_out[((long long)b * P + t) * 9 + k] = m.m[k];
        if (quat_out) { const Q4 q = mat_to_q(m); float* o = quat_out + ((long long)b * P + t) * 4; o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z; }
    }
    if (tid == 0 && trans_out) {                                        // translations: sequential re-integration
        const float sc = scale[b];
        float c[3] = {tb[0], tb[1], tb[2]};
        float* o = trans_out + (long long)b * P * 3;
        o[0] = c[0]; o[1] = c[1]; o[2] = c[2];
        for (int t = 0; t + 1 < P; ++t) {
            const float d[3] = {tb[(t + 1) * 3] - tb[t * 3], tb[(t + 1) * 3 + 1] - tb[t * 3 + 1], tb[(t + 1) * 3 + 2] - tb[t * 3 + 2]};
#pragma unroll
            for (int k = 0; k < 3; ++k) { c[k] = c[k] + (Ra[k * 3] * d[0] + Ra[k * 3 + 1] * d[1] + Ra[k * 3 + 2] * d[2]) * sc; o[(t + 1) * 3 + k] = c[k]; }
        }
    }
}

// de-heading step of HeadNormalFormer.forward_for_eval (:268-276): out_rot = R rot, out_trans = R (trans - trans[0]) + offset
static __global__ void s1_rigid_apply_kernel(const float* __restrict__ Rm /* [B,9] */, const float* __restrict__ offset /* [B,3] nullable */,
                                             const float* __restrict__ rot, const float* __restrict__ trans, int B, int P,
                                             float* __restrict__ trans_out, float* __restrict__ rot_out, float* __restrict__ quat_out) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * P) return;
    const int b = gid / P;
    const float* R = Rm + b * 9;
    if (trans_out) {
        const float* t0 = trans + (long long)b * P * 3;
        const float* tt = trans + (long long)gid * 3;
        const float d[3] = {tt[0] - t0[0], tt[1] - t0[1], tt[2] - t0[2]};
#pragma unroll
        for (int k = 0; k < 3; ++k)
            trans_out[(long long)gid * 3 + k] = (R[k * 3] * d[0] + R[k * 3 + 1] * d[1] + R[k * 3 + 2] * d[2]) + (offset ? offset[b * 3 + k] : 0.f);
    }
    if (rot_out || quat_out) {
        const float* A = rot + (long long)gid * 9;
        M3 m;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) m.m[i * 3 + j] = R[i * 3] * A[j] + R[i * 3 + 1] * A[3 + j] + R[i * 3 + 2] * A[6 + j];
        if (rot_out) for (int k = 0; k < 9; ++k) rot_out[(long long)gid * 9 + k] = m.m[k];
        if (quat_out) { const Q4 q = mat_to_q(m); float* o = quat_out + (long long)gid * 4; o[0] = q.w; o[1] = q.x; o[2] = q.y; o[3] = q.z; }
    }
}

struct S1Buf {
    float* p = nullptr;
    int alloc(size_t n) { release(); EG_CUDA(cudaMalloc(&p, n * sizeof(float))); EG_CUDA(cudaMemset(p, 0, n * sizeof(float))); return 0; }
    int upload(const std::vector<float>& h) { if (alloc(h.size())) return 1; EG_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice)); return 0; }
    void release() { if (p) cudaFree(p); p = nullptr; }
    ~S1Buf() { release(); }
};

struct S1Layer { S1Buf wqkv, bqkv, fc_w, fc_b, ln1_g, ln1_b, w1, b1, w2, b2, ln2_g, ln2_b; };
struct S1Head { std::vector<S1Buf> w, b; S1Buf fc_w, fc_b; std::vector<int> dims; int out = 0; };

}  // namespace egoego

using namespace egoego;

struct egoego_seqnet_ctx {
    egoego_seqnet_cfg cfg{};
    int kpad = 0, max_hidden = 0;
    std::map<std::string, std::vector<float>> staged;
    bool committed = false;
    S1Buf start_w, start_b, pos;
    std::vector<std::unique_ptr<S1Layer>> layers;
    std::vector<std::unique_ptr<S1Head>> heads;
    S1Buf Ain, H, Y, QKV, O, F, T1, T2;
    int* n_valid = nullptr;
    int64_t launches = 0;
    ~egoego_seqnet_ctx() { if (n_valid) cudaFree(n_valid); }
};

static int s1_expected(const egoego_seqnet_ctx* c, std::map<std::string, int64_t>& e) {
    const egoego_seqnet_cfg& g = c->cfg;
    const int d = g.d_model, H = g.n_head, dk = g.d_k;
    e["start_conv.weight"] = (int64_t)d * g.d_feats; e["start_conv.bias"] = d;
    e["position_vec.weight"] = (int64_t)(g.window + 1) * d;
    for (int l = 0; l < g.n_layers; ++l) {
        const std::string a = "layer_stack." + std::to_string(l) + ".self_attn.", f = "layer_stack." + std::to_string(l) + ".pos_ffn.";
        for (const char* nm : {"w_q", "w_k", "w_v"}) { e[a + nm + ".weight"] = (int64_t)H * dk * d; e[a + nm + ".bias"] = H * dk; }
        e[a + "fc.weight"] = (int64_t)d * H * dk; e[a + "fc.bias"] = d; e[a + "layer_norm.weight"] = d; e[a + "layer_norm.bias"] = d;
        e[f + "w_1.weight"] = (int64_t)d * d; e[f + "w_1.bias"] = d; e[f + "w_2.weight"] = (int64_t)d * d; e[f + "w_2.bias"] = d;
        e[f + "layer_norm.weight"] = d; e[f + "layer_norm.bias"] = d;
    }
    for (int h = 0; h < g.n_heads; ++h) {
        int last = d;
        const std::string p = "head" + std::to_string(h) + ".";
        for (int j = 0; j < g.head_n_hidden[h]; ++j) {
            e[p + "affine_layers." + std::to_string(j) + ".weight"] = (int64_t)g.head_hidden[h][j] * last;
            e[p + "affine_layers." + std::to_string(j) + ".bias"] = g.head_hidden[h][j];
            last = g.head_hidden[h][j];
        }
        e[p + "fc.weight"] = (int64_t)g.head_out[h] * last; e[p + "fc.bias"] = g.head_out[h];
    }
    return 0;
}

// C = A W^T + epilogue: 32 x 32 tiles while the problem is a handful of windows (latency-bound), 128 x 128 tiles otherwise
template <class Epi>
static void s1_gemm(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const Epi& e, cudaStream_t s) {
    if (M <= 1024 && K % 32 == 0) sgemm_tn_small_kernel<<<dim3((N + 31) / 32, M / 32), 256, 0, s>>>(A, lda, W, ldw, N, K, e);
    else                          sgemm_tn_kernel<<<dim3((N + 127) / 128, M / 128), 256, 0, s>>>(A, lda, W, ldw, N, K, e);
}

extern "C" {

int egoego_seqnet_create(const egoego_seqnet_cfg* cfg, egoego_seqnet* out) {
    EG_CHECK(cfg && out, "null argument");
    EG_CHECK(cfg->d_model == 256, "sequence nets are specialised for d_model = 256 (LayerNorm / GEMM tiles)");
    EG_CHECK(cfg->d_k == 256 && cfg->d_v == 256, "only d_k = d_v = 256 is supported");
    EG_CHECK(cfg->n_head >= 1 && cfg->n_layers >= 1, "n_head and n_layers must be >= 1");
    EG_CHECK(cfg->window >= 1 && cfg->window <= LP, "window must be in [1,128]");
    EG_CHECK(cfg->d_feats >= 1 && cfg->d_feats <= 4096, "d_feats out of range");
    EG_CHECK(cfg->max_batch >= 1, "max_batch must be >= 1");
    EG_CHECK(cfg->n_heads >= 0 && cfg->n_heads <= 2, "at most two MLP heads");
    int mh = cfg->d_model;
    for (int h = 0; h < cfg->n_heads; ++h) {
        EG_CHECK(cfg->head_n_hidden[h] >= 1 && cfg->head_n_hidden[h] <= 3 && cfg->head_out[h] >= 1 && cfg->head_out[h] <= 128, "bad head shape");
        for (int j = 0; j < cfg->head_n_hidden[h]; ++j) {
            EG_CHECK(cfg->head_hidden[h][j] >= 16 && cfg->head_hidden[h][j] % 16 == 0, "head hidden sizes must be multiples of 16");
            mh = std::max(mh, cfg->head_hidden[h][j]);
        }
    }
    int ndev = 0;
    EG_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "no CUDA device: libegoego_b200 has no CPU fallback");
    EG_CHECK(cfg->device >= 0 && cfg->device < ndev, "bad device ordinal");
    cudaDeviceProp prop;
    EG_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    EG_CHECK(prop.major == 10, "libegoego_b200 is built for sm_100a (B200) only");
    egoego_seqnet_ctx* c = new egoego_seqnet_ctx();
    c->cfg = *cfg;
    c->kpad = ((cfg->d_feats + 15) / 16) * 16;
    c->max_hidden = mh;
    *out = c;
    return 0;
}

void egoego_seqnet_destroy(egoego_seqnet c) { delete c; }

int egoego_seqnet_set_tensor(egoego_seqnet c, const char* name, const float* host, int64_t numel) {
    EG_CHECK(c && name && host, "null argument");
    std::map<std::string, int64_t> e;
    s1_expected(c, e);
    auto it = e.find(name);
    if (it == e.end()) return 0;                        // strict=False: unknown keys are ignored
    EG_CHECK(it->second == numel, std::string("tensor '") + name + "': expected " + std::to_string(it->second) + " elements, got " + std::to_string(numel));
    c->staged[name].assign(host, host + numel);
    c->committed = false;
    return 0;
}

int egoego_seqnet_commit(egoego_seqnet c) {
    EG_CHECK(c, "null handle");
    std::map<std::string, int64_t> e;
    s1_expected(c, e);
    for (auto& kv : e) EG_CHECK(c->staged.count(kv.first), "missing tensor: " + kv.first);
    EG_CUDA(cudaSetDevice(c->cfg.device));
    const egoego_seqnet_cfg& g = c->cfg;
    const int d = g.d_model, H = g.n_head, dk = g.d_k;
    {   // start_conv weight [d, d_feats] -> zero-padded [d, kpad]
        std::vector<float> w((size_t)d * c->kpad, 0.f);
        const std::vector<float>& s = c->staged["start_conv.weight"];
        for (int r = 0; r < d; ++r) memcpy(&w[(size_t)r * c->kpad], &s[(size_t)r * g.d_feats], g.d_feats * sizeof(float));
        if (c->start_w.upload(w) || c->start_b.upload(c->staged["start_conv.bias"]) || c->pos.upload(c->staged["position_vec.weight"])) return 1;
    }
    c->layers.clear();
    for (int l = 0; l < g.n_layers; ++l) {
        std::unique_ptr<S1Layer> L(new S1Layer());
        const std::string a = "layer_stack." + std::to_string(l) + ".self_attn.", f = "layer_stack." + std::to_string(l) + ".pos_ffn.";
        std::vector<float> wqkv, bqkv;
        for (const char* nm : {"w_q", "w_k", "w_v"}) {
            const std::vector<float>& w = c->staged[a + nm + ".weight"]; wqkv.insert(wqkv.end(), w.begin(), w.end());
            const std::vector<float>& b = c->staged[a + nm + ".bias"]; bqkv.insert(bqkv.end(), b.begin(), b.end());
        }
        if (L->wqkv.upload(wqkv) || L->bqkv.upload(bqkv) || L->fc_w.upload(c->staged[a + "fc.weight"]) || L->fc_b.upload(c->staged[a + "fc.bias"]) ||
            L->ln1_g.upload(c->staged[a + "layer_norm.weight"]) || L->ln1_b.upload(c->staged[a + "layer_norm.bias"]) ||
            L->w1.upload(c->staged[f + "w_1.weight"]) || L->b1.upload(c->staged[f + "w_1.bias"]) ||
            L->w2.upload(c->staged[f + "w_2.weight"]) || L->b2.upload(c->staged[f + "w_2.bias"]) ||
            L->ln2_g.upload(c->staged[f + "layer_norm.weight"]) || L->ln2_b.upload(c->staged[f + "layer_norm.bias"])) return 1;
        c->layers.push_back(std::move(L));
    }
    c->heads.clear();
    for (int h = 0; h < g.n_heads; ++h) {
        std::unique_ptr<S1Head> Hd(new S1Head());
        const std::string p = "head" + std::to_string(h) + ".";
        Hd->w.resize(g.head_n_hidden[h]); Hd->b.resize(g.head_n_hidden[h]);
        for (int j = 0; j < g.head_n_hidden[h]; ++j) {
            if (Hd->w[j].upload(c->staged[p + "affine_layers." + std::to_string(j) + ".weight"]) ||
                Hd->b[j].upload(c->staged[p + "affine_layers." + std::to_string(j) + ".bias"])) return 1;
            Hd->dims.push_back(g.head_hidden[h][j]);
        }
        if (Hd->fc_w.upload(c->staged[p + "fc.weight"]) || Hd->fc_b.upload(c->staged[p + "fc.bias"])) return 1;
        Hd->out = g.head_out[h];
        c->heads.push_back(std::move(Hd));
    }
    const size_t M = (size_t)g.max_batch * LP;
    if (c->Ain.alloc(M * c->kpad) || c->H.alloc(M * d) || c->Y.alloc(M * d) || c->QKV.alloc(M * 3 * H * dk) || c->O.alloc(M * H * dk) ||
        c->F.alloc(M * d) || c->T1.alloc(M * c->max_hidden) || c->T2.alloc(M * c->max_hidden)) return 1;
    if (c->n_valid) cudaFree(c->n_valid);
    EG_CUDA(cudaMalloc(&c->n_valid, g.max_batch * sizeof(int)));
    EG_CUDA(cudaFuncSetAttribute(attention_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SIMT_SMEM));
    c->committed = true;
    return 0;
}

int egoego_seqnet_forward(egoego_seqnet c, const float* feats, int B, int T, float* dec_out, float* head0_out, float* head1_out,
                          int token0_only, void* stream_v) {
    EG_CHECK(c && feats, "null argument");
    EG_CHECK(c->committed, "egoego_seqnet_commit has not been called");
    const egoego_seqnet_cfg& g = c->cfg;
    EG_CHECK(B >= 1 && B <= g.max_batch, "B out of range (1..max_batch)");
    EG_CHECK(T >= 1 && T <= g.window, "T out of range (1..window)");
    EG_CHECK((head0_out == nullptr || g.n_heads >= 1) && (head1_out == nullptr || g.n_heads >= 2), "head output requested for an undefined head");
    EG_CUDA(cudaSetDevice(g.device));
    cudaStream_t s = (cudaStream_t)stream_v;
    const int M = B * LP, d = g.d_model, H = g.n_head, dk = g.d_k, nqkv = 3 * H * dk;
    std::vector<int> nv(B, T);
    EG_CUDA(cudaMemcpyAsync(c->n_valid, nv.data(), B * sizeof(int), cudaMemcpyHostToDevice, s));
    EG_CUDA(cudaMemsetAsync(c->Ain.p, 0, (size_t)M * c->kpad * sizeof(float), s));
    {
        const long long tot = (long long)B * T * g.d_feats;
        s1_stage_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(c->Ain.p, c->kpad, feats, g.d_feats, B, T);
        S1EpiStart e{c->H.p, d, c->start_b.p, c->pos.p, g.window};
        s1_gemm(c->Ain.p, c->kpad, c->start_w.p, c->kpad, M, d, c->kpad, e, s);
        c->launches += 2;
    }
    for (auto& Lp : c->layers) {
        S1Layer& w = *Lp;
        EpiBiasScale eq{c->QKV.p, nqkv, w.bqkv.p, H * dk, 1.0f / sqrtf((float)dk)};
        s1_gemm(c->H.p, d, w.wqkv.p, d, M, nqkv, d, eq, s);
        attention_simt_kernel<false><<<B * H, 256, ATT_SIMT_SMEM, s>>>(c->QKV.p, nqkv, c->O.p, nullptr, nullptr, H * dk, H, g.window);
        EpiBiasResid ef{c->Y.p, d, w.fc_b.p, c->H.p};
        s1_gemm(c->O.p, H * dk, w.fc_w.p, H * dk, M, d, H * dk, ef, s);
        s1_layernorm_kernel<256><<<M / 8, 256, 0, s>>>(c->Y.p, c->H.p, w.ln1_g.p, w.ln1_b.p, c->n_valid, M);
        EpiBiasRelu e1{c->F.p, d, w.b1.p};
        s1_gemm(c->H.p, d, w.w1.p, d, M, d, d, e1, s);
        EpiBiasResid e2{c->Y.p, d, w.b2.p, c->H.p};
        s1_gemm(c->F.p, d, w.w2.p, d, M, d, d, e2, s);
        s1_layernorm_kernel<256><<<M / 8, 256, 0, s>>>(c->Y.p, c->H.p, w.ln2_g.p, w.ln2_b.p, c->n_valid, M);
        c->launches += 7;
    }
    if (dec_out)      // compact copy [B, window, d] of the decoder output
        EG_CUDA(cudaMemcpy2DAsync(dec_out, (size_t)g.window * d * 4, c->H.p, (size_t)LP * d * 4, (size_t)g.window * d * 4, B, cudaMemcpyDeviceToDevice, s));
    float* outs[2] = {head0_out, head1_out};
    for (int h = 0; h < g.n_heads; ++h) {
        if (!outs[h]) continue;
        S1Head& hd = *c->heads[h];
        const float* x = c->H.p;
        int last = d;
        float* buf[2] = {c->T1.p, c->T2.p};
        for (size_t j = 0; j < hd.dims.size(); ++j) {
            EpiBiasRelu e{buf[j & 1], hd.dims[j], hd.b[j].p};
            s1_gemm(x, last, hd.w[j].p, last, M, hd.dims[j], last, e, s);
            x = buf[j & 1]; last = hd.dims[j];
            c->launches++;
        }
        S1EpiHeadOut eo{outs[h], hd.out, hd.fc_b.p, token0_only ? 1 : T};
        s1_gemm(x, last, hd.fc_w.p, last, M, hd.out, last, eo, s);
        c->launches++;
    }
    EG_CUDA(cudaGetLastError());
    return 0;
}

int64_t egoego_seqnet_launch_count(egoego_seqnet c) { return c ? c->launches : -1; }

int egoego_va2rot(int device, const float* q0, const float* va, int B, int T, float dt, float* out, void* stream_v) {
    EG_CHECK(q0 && va && out && B >= 1 && T >= 1, "bad argument");
    EG_CUDA(cudaSetDevice(device));
    s1_va2rot_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream_v>>>(q0, va, B, T, dt, out);
    EG_CUDA(cudaGetLastError());
    return 0;
}

int egoego_rescale_slam(int device, const float* slam_trans, int n_pose, const float* dist, int n_dist, float dist_scale,
                        float* trans_out, float* scale_out, void* stream_v) {
    EG_CHECK(slam_trans && dist && trans_out && scale_out && n_pose >= 2 && n_dist >= 1 && dist_scale != 0.f, "bad argument");
    EG_CUDA(cudaSetDevice(device));
    s1_rescale_kernel<<<1, 256, 0, (cudaStream_t)stream_v>>>(slam_trans, n_pose, dist, n_dist, dist_scale, trans_out, scale_out);
    EG_CUDA(cudaGetLastError());
    return 0;
}

int egoego_slam_features(int device, const float* rot_mat, const float* trans, int B, int n_pose_stride, int n_pose, float* feats, void* stream_v) {
    EG_CHECK(rot_mat && trans && feats && B >= 1 && n_pose >= 2 && n_pose_stride >= n_pose, "bad argument");
    EG_CUDA(cudaSetDevice(device));
    const int n = B * (n_pose - 1);
    s1_slam_features_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_v>>>(rot_mat, trans, B, n_pose_stride, n_pose, feats);
    EG_CUDA(cudaGetLastError());
    return 0;
}

int egoego_apply_floor_normal(int device, const float* normal, const float* scale, const float* rot_mat, const float* trans, int B, int n_pose,
                              float* trans_out, float* rot_out, float* quat_out, float* align_rot_out, void* stream_v) {
    EG_CHECK(normal && scale && rot_mat && trans && B >= 1 && n_pose >= 1, "bad argument");
    EG_CUDA(cudaSetDevice(device));
    s1_apply_normal_kernel<<<B, 128, 0, (cudaStream_t)stream_v>>>(normal, scale, rot_mat, trans, n_pose, trans_out, rot_out, quat_out, align_rot_out);
    EG_CUDA(cudaGetLastError());
    return 0;
}

int egoego_rigid_apply(int device, const float* rot3x3, const float* offset, const float* rot_mat, const float* trans, int B, int n_pose,
                       float* trans_out, float* rot_out, float* quat_out, void* stream_v) {
    EG_CHECK(rot3x3 && rot_mat && trans && B >= 1 && n_pose >= 1, "bad argument");
    EG_CUDA(cudaSetDevice(device));
    const int n = B * n_pose;
    s1_rigid_apply_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream_v>>>(rot3x3, offset, rot_mat, trans, B, n_pose, trans_out, rot_out, quat_out);
    EG_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
