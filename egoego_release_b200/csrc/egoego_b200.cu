// libegoego_b200: C-ABI + orchestration of the stage-2 sampling path (see include/egoego_b200.h).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <set>
#include <type_traits>

#include "common.cuh"
#include "kernels_simt.cuh"
#include "postprocess.cuh"
#include "metrics.cuh"
#include "floor.cuh"
#include "engine_tc.cuh"

namespace egoego {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
    int alloc(size_t n) {
        release();
        if (n == 0) return 0;
        EG_CUDA(cudaMalloc(&p, n));
        bytes = n;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct LayerW {
    DevBuf wqkv, bqkv;       // fused [3*H*dk, d], [3*H*dk]
    DevBuf fc_w, fc_b, ln1_g, ln1_b;
    DevBuf w1, b1, w2, b2, ln2_g, ln2_b;
};

}  // namespace egoego

using namespace egoego;

struct egoego_ctx {
    egoego_cfg cfg{};
    int D = 0, d = 0, H = 0, dk = 0, NL = 0, N = 0, Tmax = 0;
    int kin_pad = 0;                       // 2*D rounded up to 16 (SIMT start GEMM K)
    std::map<std::string, std::vector<float>> staged;   // host copies until commit
    bool committed = false;

    // packed fp32 weights (both engines)
    DevBuf start_w, start_b, pos, out_w, out_b, t_w1, t_b1, t_w2, t_b2, temb;
    std::vector<LayerW> layers;
    DevBuf coef1, coef2, logvar, sqrt_recip, sqrt_recipm1;
    bool have_sched = false;

    // workspace (sized for cfg.max_batch windows)
    DevBuf Ain, Hbuf, Ybuf, QKV, Obuf, Fbuf, model_out, x_cur, x_cond, d_step, t_tmp;
    DevBuf h_xstart, h_mask, h_out, h_tape;   // device staging of the *_host entry point

    Skeleton sk{}; bool have_sk = false;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev_in = nullptr, ev_out = nullptr;
    // one captured step per format: [0] = FMT_SPLIT with dual accumulators (the last `dual_last` steps), [1 + r] = FMT_HALF reading
    // dithered weight set r (engine_tc.cu, upload_weight), [1 + MAX_WEIGHT_SETS] = FMT_SPLIT with a single accumulator,
    // [2 + MAX_WEIGHT_SETS] = FMT_HALF kernels reading the weights as an fp16 pair in two passes ("pair" steps)
    static constexpr int MAX_WEIGHT_SETS = 16;
    cudaGraphExec_t step_graph[3 + MAX_WEIGHT_SETS] = {};
    int split_last = 16;                   // of the last `precise_last` steps, those with t < split_last run the 3-term split; the others
                                           // (split_last <= t < precise_last) fp16 activations x fp16-pair weights (EGOEGO_SPLIT_STEPS; >= N: all split)
    int dual_last = 16;                    // steps t < dual_last run the dual-accumulator split GEMMs (EGOEGO_DUAL_STEPS; >= N: every split step)
    int graph_B = -1, graph_T = -1; const void* graph_key[4] = {nullptr};
    int precise_last = 0;                  // steps t < precise_last use the 3-term split; earlier steps single-pass fp16
    bool use_graph = true;
    bool temb_dirty = false;               // time_mlp weights were updated in place (egoego_update_tensor_device): rebuild the table before use
    bool fuse_ddpm = false;                // EGOEGO_FUSE_DDPM=1: DDPM update in linear_out's epilogue.  Opt-in: measured 90 us vs 17.6 + 32.9 us for
                                           // linear_out + ddpm_update_kernel at B = 256 (8 epilogue warps per SM are too few for the Philox work)
    int64_t launches = 0;
    double drop_p = 0.0;                   // training step: dropout probability (0 = eval-mode semantics) and mask seed
    unsigned long long drop_seed = 0;      // (egoego_train_set_dropout)
    std::unique_ptr<TcEngine> tc;
};

namespace egoego {

void train_release(egoego_ctx* c);     // frees the training workspace of a handle (defined with the training step below)
void train_drop_graphs(egoego_ctx* c); // captured training steps hold the handle's buffer addresses: dropped whenever those may move

// largest opt-in dynamic shared-memory size requested so far per device (function attributes are per device)
struct PerDeviceMax {
    size_t v[64] = {};
    bool raise(int dev, size_t want) { dev &= 63; if (want <= v[dev]) return false; v[dev] = want; return true; }
};

static inline dim3 grid1d(long long n, int bs) { return dim3((unsigned)((n + bs - 1) / bs)); }
// ddpm_update_kernel: x = quads (4 consecutive elements) of one window in blocks of 256 threads, y = window
static inline dim3 ddpm_grid(int T, int D, int B) { return dim3((unsigned)(((T * D + 3) / 4 + 255) / 256), (unsigned)B); }

// ---- SIMT engine: one denoiser call.  Expects Ain staged; writes model_out[B,T,D]. ---------------
static int denoiser_simt(egoego_ctx* c, int B, int T, TSrc ts, const float* pmask, cudaStream_t s) {
    const int M = B * LP, d = c->d, H = c->H, dk = c->dk;
    const int L = T + 1;
    {
        EpiStart e{c->Hbuf.as<float>(), d, c->start_b.as<float>(), nullptr, c->pos.as<float>(), c->temb.as<float>(), ts, T};
        sgemm_tn_kernel<<<dim3(d / 128, M / 128), 256, 0, s>>>(c->Ain.as<float>(), c->kin_pad, c->start_w.as<float>(),
                                                              c->kin_pad, d, c->kin_pad, e);
        c->launches++;
    }
    for (int l = 0; l < c->NL; ++l) {
        LayerW& w = c->layers[l];
        const int nqkv = 3 * H * dk;
        EpiBiasScale eq{c->QKV.as<float>(), nqkv, w.bqkv.as<float>(), H * dk, 1.0f / sqrtf((float)dk)};
        sgemm_tn_kernel<<<dim3(nqkv / 128, M / 128), 256, 0, s>>>(c->Hbuf.as<float>(), d, w.wqkv.as<float>(), d, nqkv, d, eq);
        attention_simt_kernel<false><<<B * H, 256, ATT_SIMT_SMEM, s>>>(c->QKV.as<float>(), nqkv, c->Obuf.as<float>(), nullptr, nullptr, H * dk, H, L);
        EpiBiasResid ef{c->Ybuf.as<float>(), d, w.fc_b.as<float>(), c->Hbuf.as<float>()};
        sgemm_tn_kernel<<<dim3(d / 128, M / 128), 256, 0, s>>>(c->Obuf.as<float>(), H * dk, w.fc_w.as<float>(), H * dk, d, H * dk, ef);
        layernorm512_kernel<<<M / 8, 256, 0, s>>>(c->Ybuf.as<float>(), c->Hbuf.as<float>(), nullptr, nullptr,
                                                  w.ln1_g.as<float>(), w.ln1_b.as<float>(), pmask, T, M, 0);
        EpiBiasRelu e1{c->Fbuf.as<float>(), d, w.b1.as<float>()};
        sgemm_tn_kernel<<<dim3(d / 128, M / 128), 256, 0, s>>>(c->Hbuf.as<float>(), d, w.w1.as<float>(), d, d, d, e1);
        EpiBiasResid e2{c->Ybuf.as<float>(), d, w.b2.as<float>(), c->Hbuf.as<float>()};
        sgemm_tn_kernel<<<dim3(d / 128, M / 128), 256, 0, s>>>(c->Fbuf.as<float>(), d, w.w2.as<float>(), d, d, d, e2);
        layernorm512_kernel<<<M / 8, 256, 0, s>>>(c->Ybuf.as<float>(), c->Hbuf.as<float>(), nullptr, nullptr,
                                                  w.ln2_g.as<float>(), w.ln2_b.as<float>(), pmask, T, M, 0);
        c->launches += 7;
    }
    {
        EpiOut eo{c->model_out.as<float>(), c->D, c->out_b.as<float>(), T};
        sgemm_tn_kernel<<<dim3((c->D + 127) / 128, M / 128), 256, 0, s>>>(c->Hbuf.as<float>(), d, c->out_w.as<float>(), d, c->D, d, eo);
        c->launches++;
    }
    EG_CUDA(cudaGetLastError());
    return 0;
}

static int run_denoiser(egoego_ctx* c, int B, int T, TSrc ts, const float* pmask, cudaStream_t s, int fmt = 0,
                        const DdpmArgs* fuse = nullptr) {
    if (c->cfg.engine == EGOEGO_ENGINE_SIMT) return denoiser_simt(c, B, T, ts, pmask, s);
    int64_t n = 0;
    int rc = c->tc->denoiser(B, T, ts, pmask, c->model_out.as<float>(), s, &n, fmt, fuse);
    c->launches += n;
    return rc;
}

// Stage the GEMM input (x half and/or x_cond half) of the engine from compact [B,T,*] rows.
static int stage_input(egoego_ctx* c, const float* src, int src_ld, int src_col0, bool cond_half, int B, int T, cudaStream_t s) {
    const long long tot = (long long)B * T * c->D;
    if (c->cfg.engine == EGOEGO_ENGINE_SIMT) {
        stage_rows_f32_kernel<<<grid1d(tot, 256), 256, 0, s>>>(c->Ain.as<float>(), c->kin_pad, cond_half ? c->D : 0,
                                                              src, src_ld, src_col0, c->D, B, T);
        c->launches++;
        EG_CUDA(cudaGetLastError());
        return 0;
    }
    int64_t n = 0;
    int rc = c->tc->stage(src, src_ld, src_col0, cond_half, B, T, s, &n);
    c->launches += n;
    return rc;
}

// After the x_cond half has been staged: engine-specific per-window constant work (tensor engine
// folds the x_cond contribution of start_conv + bias + positional rows into a base tensor).
static int prepare_cond(egoego_ctx* c, int B, int T, cudaStream_t s) {
    if (c->cfg.engine == EGOEGO_ENGINE_SIMT) return 0;
    int64_t n = 0;
    int rc = c->tc->prepare_cond(B, T, s, &n);
    c->launches += n;
    return rc;
}

static int clear_staging(egoego_ctx* c, int B, cudaStream_t s) {
    if (c->cfg.engine == EGOEGO_ENGINE_SIMT) {
        EG_CUDA(cudaMemsetAsync(c->Ain.p, 0, (size_t)B * LP * c->kin_pad * sizeof(float), s));
        return 0;
    }
    return c->tc->clear_staging(B, s);
}

static void fill_ddpm(egoego_ctx* c, DdpmArgs& a, const float* x, float* x_out, TSrc ts, NoiseSrc ns, int clip,
                      const float* inpaint, int inpaint_len, int B, int T, bool stage_next) {
    a.model_out = c->model_out.as<float>(); a.x = x; a.x_out = x_out;
    a.coef1 = c->coef1.as<float>(); a.coef2 = c->coef2.as<float>(); a.logvar = c->logvar.as<float>();
    a.sqrt_recip = c->sqrt_recip.as<float>(); a.sqrt_recipm1 = c->sqrt_recipm1.as<float>();
    a.objective = c->cfg.objective; a.clip = clip; a.inpaint = inpaint; a.inpaint_len = inpaint_len;
    a.stage_f32 = nullptr; a.stage_ld = 0; a.stage_hi = nullptr; a.stage_lo = nullptr; a.stage_ld16 = 0; a.stage_h16 = nullptr;
    a.stage_mode = 2;
    if (stage_next) {
        if (c->cfg.engine == EGOEGO_ENGINE_SIMT) { a.stage_f32 = c->Ain.as<float>(); a.stage_ld = c->kin_pad; }
        else c->tc->stage_targets(&a.stage_hi, &a.stage_lo, &a.stage_h16, &a.stage_ld16);
    }
    a.ts = ts; a.ns = ns; a.B = B; a.T = T; a.D = c->D;
    a.advance = nullptr; a.done = nullptr;
}


// ---- content checksum of a set of device tensors (weight-staleness check of the host mirror) ----------------------
// One launch over all tensors: every 32-bit word contributes word * (2 * global_index + 1) to a wrapping 64-bit sum, so a
// change of any single bit of any element changes the result.  The table of (pointer, words) travels as a kernel argument.
constexpr int CHK_MAX_TENSORS = 128;
struct ChkTable { const uint32_t* p[CHK_MAX_TENSORS]; unsigned long long n[CHK_MAX_TENSORS]; unsigned long long start[CHK_MAX_TENSORS]; int count; };
__global__ void tensors_checksum_kernel(const __grid_constant__ ChkTable tab, unsigned long long* out) {
    unsigned long long acc = 0;
    for (int t = blockIdx.y; t < tab.count; t += gridDim.y) {
        const uint32_t* p = tab.p[t];
        const unsigned long long n = tab.n[t], base = tab.start[t];
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
            acc += (unsigned long long)p[i] * (2ull * (base + i) + 1ull);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

static int check_ready(egoego_ctx* c, int B, int T) {
    EG_CHECK(c != nullptr, "null handle");
    EG_CHECK(c->committed, "egoego_commit_weights has not been called");
    EG_CHECK(B >= 1, "B must be >= 1");
    EG_CHECK(T >= 1 && T <= c->Tmax, "T out of range (1..max_timesteps-1)");
    return 0;
}

}  // namespace egoego

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char* egoego_last_error(void) { return g_err.c_str(); }
int egoego_version(void) { return 100; }

int egoego_create(const egoego_cfg* cfg, egoego_handle* out) {
    EG_CHECK(cfg && out, "null argument");
    EG_CHECK(cfg->d_model == 512, "only d_model = 512 is supported (LayerNorm / tile shapes are specialised)");
    EG_CHECK(cfg->d_k == 256 && cfg->d_v == 256, "only d_k = d_v = 256 is supported");
    EG_CHECK(cfg->n_head >= 1 && cfg->n_head * cfg->d_k % 128 == 0, "n_head * d_k must be a multiple of 128");
    EG_CHECK(cfg->max_timesteps >= 2 && cfg->max_timesteps <= LP, "max_timesteps must be in [2,128]");
    EG_CHECK(cfg->d_feats >= 1 && cfg->d_feats <= 256 && cfg->d_feats % 2 == 0, "d_feats must be even and <= 256");
    EG_CHECK(cfg->timesteps >= 1, "timesteps must be >= 1");
    EG_CHECK(cfg->objective == 0 || cfg->objective == 1, "objective must be 0 (pred_noise) or 1 (pred_x0)");
    EG_CHECK(cfg->max_batch >= 1 && cfg->max_batch <= 65535, "max_batch must be in [1, 65535]");
    EG_CHECK(cfg->engine == EGOEGO_ENGINE_TCGEN05 || cfg->engine == EGOEGO_ENGINE_SIMT, "unknown engine");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    EG_CHECK(e == cudaSuccess && ndev > 0, "no CUDA device: libegoego_b200 has no CPU fallback");
    EG_CHECK(cfg->device >= 0 && cfg->device < ndev, "bad device ordinal");
    cudaDeviceProp prop;
    EG_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    EG_CHECK(prop.major == 10, "libegoego_b200 is built for sm_100a (B200) only; device is sm_" +
                                   std::to_string(prop.major) + std::to_string(prop.minor));
    EG_CUDA(cudaSetDevice(cfg->device));
    auto* c = new egoego_ctx();
    c->cfg = *cfg;
    c->D = cfg->d_feats; c->d = cfg->d_model; c->H = cfg->n_head; c->dk = cfg->d_k; c->NL = cfg->n_dec_layers;
    c->N = cfg->timesteps; c->Tmax = cfg->max_timesteps - 1;
    c->kin_pad = ((2 * c->D + 15) / 16) * 16;
    c->layers.resize(c->NL);
    { const char* fd = getenv("EGOEGO_FUSE_DDPM"); c->fuse_ddpm = fd && fd[0] == '1'; }
    const char* g = getenv("EGOEGO_GRAPH");
    c->use_graph = !(g && g[0] == '0');
    // precision policy: the last `precise_last` steps (t < precise_last) run the 3-term fp16 hi/lo split,
    // earlier steps a single fp16 pass.  cfg.precise_last_steps: 0 (the value of a zero-initialised egoego_cfg) and -1 select
    // the default max(ceil(N/16), 48); K > 0 = K split steps (K >= N: every step); EGOEGO_PRECISE_ALL_FP16 (-2) is the only
    // way to run every step single-pass -- out of tolerance, kept for measurements.
    {
        int pl = cfg->precise_last_steps;
        const char* e = getenv("EGOEGO_PRECISE_STEPS");
        if (e && e[0]) pl = atoi(e);
        EG_CHECK(pl >= EGOEGO_PRECISE_ALL_FP16, "precise_last_steps must be >= 0, -1 (default) or EGOEGO_PRECISE_ALL_FP16");
        if (pl == EGOEGO_PRECISE_ALL_FP16) pl = 0;
        else if (pl <= 0) { pl = (cfg->timesteps + 15) / 16; if (pl < 48) pl = 48; }
        if (pl > cfg->timesteps) pl = cfg->timesteps;
        c->precise_last = (cfg->engine == EGOEGO_ENGINE_TCGEN05) ? pl : cfg->timesteps;
    }
    { const char* e = getenv("EGOEGO_DUAL_STEPS"); if (e && e[0]) c->dual_last = atoi(e); if (c->dual_last < 0) c->dual_last = 0; }
    // precise_last_steps >= timesteps is the "3-term split at every step" engine (the in-library fp32-grade reference of tests and tools)
    if (c->precise_last >= cfg->timesteps) c->split_last = cfg->timesteps;
    { const char* e = getenv("EGOEGO_SPLIT_STEPS"); if (e && e[0]) c->split_last = atoi(e); if (c->split_last < 0) c->split_last = 0; }
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming) != cudaSuccess) {
        set_error("stream/event creation failed");
        delete c;
        return 1;
    }
    EG_CUDA(cudaFuncSetAttribute(attention_simt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SIMT_SMEM));
    *out = c;
    return 0;
}

int egoego_destroy(egoego_handle c) {
    if (!c) return 0;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    train_release(c);
    for (auto& g : c->step_graph) if (g) cudaGraphExecDestroy(g);
    DevBuf* bufs[] = {&c->start_w, &c->start_b, &c->pos, &c->out_w, &c->out_b, &c->t_w1, &c->t_b1, &c->t_w2, &c->t_b2,
                      &c->temb, &c->coef1, &c->coef2, &c->logvar, &c->sqrt_recip, &c->sqrt_recipm1, &c->Ain, &c->Hbuf,
                      &c->Ybuf, &c->QKV, &c->Obuf, &c->Fbuf, &c->model_out, &c->x_cur, &c->x_cond, &c->d_step, &c->t_tmp,
                      &c->h_xstart, &c->h_mask, &c->h_out, &c->h_tape};
    for (DevBuf* b : bufs) b->release();
    for (auto& l : c->layers) {
        DevBuf* lb[] = {&l.wqkv, &l.bqkv, &l.fc_w, &l.fc_b, &l.ln1_g, &l.ln1_b, &l.w1, &l.b1, &l.w2, &l.b2, &l.ln2_g, &l.ln2_b};
        for (DevBuf* b : lb) b->release();
    }
    c->tc.reset();
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->ev_in) cudaEventDestroy(c->ev_in);
    if (c->ev_out) cudaEventDestroy(c->ev_out);
    delete c;
    return 0;
}

int egoego_set_tensor(egoego_handle c, const char* name_in, const float* data, int64_t numel, int on_device) {
    EG_CHECK(c && name_in && data, "null argument");
    EG_CHECK(numel > 0, "numel must be > 0");
    std::string name(name_in);
    for (const char* pre : {"ema_model.", "model.", "module."})
        if (name.rfind(pre, 0) == 0) name = name.substr(strlen(pre));
    // expected sizes
    const int D = c->D, d = c->d, H = c->H, dk = c->dk;
    std::map<std::string, int64_t> expect = {
        {"denoise_fn.motion_transformer.start_conv.weight", (int64_t)d * 2 * D},
        {"denoise_fn.motion_transformer.start_conv.bias", d},
        {"denoise_fn.motion_transformer.position_vec.weight", (int64_t)(c->cfg.max_timesteps + 1) * d},
        {"denoise_fn.linear_out.weight", (int64_t)D * d}, {"denoise_fn.linear_out.bias", D},
        {"denoise_fn.time_mlp.1.weight", 256 * 64}, {"denoise_fn.time_mlp.1.bias", 256},
        {"denoise_fn.time_mlp.3.weight", (int64_t)d * 256}, {"denoise_fn.time_mlp.3.bias", d},
        {"posterior_mean_coef1", c->N}, {"posterior_mean_coef2", c->N}, {"posterior_log_variance_clipped", c->N},
        {"sqrt_recip_alphas_cumprod", c->N}, {"sqrt_recipm1_alphas_cumprod", c->N}};
    for (int l = 0; l < c->NL; ++l) {
        std::string a = "denoise_fn.motion_transformer.layer_stack." + std::to_string(l) + ".self_attn.";
        std::string f = "denoise_fn.motion_transformer.layer_stack." + std::to_string(l) + ".pos_ffn.";
        for (const char* nm : {"w_q", "w_k", "w_v"}) { expect[a + nm + ".weight"] = (int64_t)H * dk * d; expect[a + nm + ".bias"] = H * dk; }
        expect[a + "fc.weight"] = (int64_t)d * H * dk; expect[a + "fc.bias"] = d;
        expect[a + "layer_norm.weight"] = d; expect[a + "layer_norm.bias"] = d;
        expect[f + "w_1.weight"] = (int64_t)d * d; expect[f + "w_1.bias"] = d;
        expect[f + "w_2.weight"] = (int64_t)d * d; expect[f + "w_2.bias"] = d;
        expect[f + "layer_norm.weight"] = d; expect[f + "layer_norm.bias"] = d;
    }
    auto it = expect.find(name);
    if (it == expect.end()) return 0;   // strict=False: ignore unknown keys
    EG_CHECK(it->second == numel, "tensor '" + name + "': expected " + std::to_string(it->second) +
                                      " elements, got " + std::to_string(numel));
    std::vector<float>& v = c->staged[name];
    v.resize(numel);
    if (on_device) {
        EG_CUDA(cudaSetDevice(c->cfg.device));
        EG_CUDA(cudaMemcpy(v.data(), data, numel * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        memcpy(v.data(), data, numel * sizeof(float));
    }
    c->committed = false;
    return 0;
}

int egoego_make_cosine_schedule(egoego_handle c) {
    EG_CHECK(c, "null handle");
    // transformer_cond_diffusion_model.py:47-57,173-214 in double precision, cast to fp32 at the end
    const int N = c->N;
    std::vector<double> ac(N + 1), betas(N), acp(N), cum(N);
    const double s = 0.008;
    for (int i = 0; i <= N; ++i) {
        double x = (double)i * ((double)N / (double)N);   // linspace(0, N, N+1)
        double v = cos(((x / N) + s) / (1 + s) * M_PI * 0.5);
        ac[i] = v * v;
    }
    for (int i = N; i >= 0; --i) ac[i] /= ac[0];
    for (int i = 0; i < N; ++i) { double b = 1 - ac[i + 1] / ac[i]; betas[i] = b < 0 ? 0 : (b > 0.999 ? 0.999 : b); }
    double run = 1.0;
    for (int i = 0; i < N; ++i) { run *= (1.0 - betas[i]); cum[i] = run; acp[i] = i == 0 ? 1.0 : cum[i - 1]; }
    std::vector<float> c1(N), c2(N), lv(N), sr(N), srm(N);
    for (int i = 0; i < N; ++i) {
        double pv = betas[i] * (1.0 - acp[i]) / (1.0 - cum[i]);
        lv[i] = (float)log(pv < 1e-20 ? 1e-20 : pv);
        c1[i] = (float)(betas[i] * sqrt(acp[i]) / (1.0 - cum[i]));
        c2[i] = (float)((1.0 - acp[i]) * sqrt(1.0 - betas[i]) / (1.0 - cum[i]));
        sr[i] = (float)sqrt(1.0 / cum[i]);
        srm[i] = (float)sqrt(1.0 / cum[i] - 1);
    }
    if (egoego_set_tensor(c, "posterior_mean_coef1", c1.data(), N, 0)) return 1;
    if (egoego_set_tensor(c, "posterior_mean_coef2", c2.data(), N, 0)) return 1;
    if (egoego_set_tensor(c, "posterior_log_variance_clipped", lv.data(), N, 0)) return 1;
    if (egoego_set_tensor(c, "sqrt_recip_alphas_cumprod", sr.data(), N, 0)) return 1;
    if (egoego_set_tensor(c, "sqrt_recipm1_alphas_cumprod", srm.data(), N, 0)) return 1;
    return 0;
}

static int upload(DevBuf& b, const std::vector<float>& v) {
    if (b.alloc(v.size() * sizeof(float))) return 1;
    EG_CUDA(cudaMemcpy(b.p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

int egoego_commit_weights(egoego_handle c, void* stream_v) {
    EG_CHECK(c, "null handle");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    train_drop_graphs(c);
    cudaStream_t s = (cudaStream_t)stream_v;
    auto need = [&](const std::string& k) -> const std::vector<float>* {
        auto it = c->staged.find(k);
        return it == c->staged.end() ? nullptr : &it->second;
    };
#define NEED(var, key) const std::vector<float>* var = need(key); EG_CHECK(var, std::string("missing tensor: ") + (key))
    const int D = c->D, d = c->d, H = c->H, dk = c->dk;
    const std::string pre = "denoise_fn.motion_transformer.";
    NEED(sw, pre + "start_conv.weight"); NEED(sb, pre + "start_conv.bias"); NEED(pv, pre + "position_vec.weight");
    NEED(ow, "denoise_fn.linear_out.weight"); NEED(ob, "denoise_fn.linear_out.bias");
    NEED(tw1, "denoise_fn.time_mlp.1.weight"); NEED(tb1, "denoise_fn.time_mlp.1.bias");
    NEED(tw2, "denoise_fn.time_mlp.3.weight"); NEED(tb2, "denoise_fn.time_mlp.3.bias");
    NEED(c1, "posterior_mean_coef1"); NEED(c2, "posterior_mean_coef2"); NEED(lv, "posterior_log_variance_clipped");
    if (c->cfg.objective == 0) { EG_CHECK(need("sqrt_recip_alphas_cumprod") && need("sqrt_recipm1_alphas_cumprod"),
                                          "pred_noise objective needs sqrt_recip(m1)_alphas_cumprod"); }
    // start_conv weight padded to kin_pad columns
    std::vector<float> swp((size_t)d * c->kin_pad, 0.f);
    for (int o = 0; o < d; ++o) memcpy(&swp[(size_t)o * c->kin_pad], &(*sw)[(size_t)o * 2 * D], 2 * D * sizeof(float));
    if (upload(c->start_w, swp) || upload(c->start_b, *sb) || upload(c->pos, *pv) || upload(c->out_w, *ow) ||
        upload(c->out_b, *ob) || upload(c->t_w1, *tw1) || upload(c->t_b1, *tb1) || upload(c->t_w2, *tw2) ||
        upload(c->t_b2, *tb2) || upload(c->coef1, *c1) || upload(c->coef2, *c2) || upload(c->logvar, *lv)) return 1;
    {
        std::vector<float> ones(c->N, 1.f), zeros(c->N, 0.f);
        const std::vector<float>* sr = need("sqrt_recip_alphas_cumprod");
        const std::vector<float>* srm = need("sqrt_recipm1_alphas_cumprod");
        if (upload(c->sqrt_recip, sr ? *sr : ones) || upload(c->sqrt_recipm1, srm ? *srm : zeros)) return 1;
    }
    std::vector<std::vector<float>> host_layers;   // for the tensor engine packer
    for (int l = 0; l < c->NL; ++l) {
        std::string a = pre + "layer_stack." + std::to_string(l) + ".self_attn.";
        std::string f = pre + "layer_stack." + std::to_string(l) + ".pos_ffn.";
        NEED(wq, a + "w_q.weight"); NEED(wk, a + "w_k.weight"); NEED(wv, a + "w_v.weight");
        NEED(bq, a + "w_q.bias"); NEED(bk, a + "w_k.bias"); NEED(bv, a + "w_v.bias");
        NEED(fw, a + "fc.weight"); NEED(fb, a + "fc.bias"); NEED(g1, a + "layer_norm.weight"); NEED(be1, a + "layer_norm.bias");
        NEED(w1, f + "w_1.weight"); NEED(b1, f + "w_1.bias"); NEED(w2, f + "w_2.weight"); NEED(b2, f + "w_2.bias");
        NEED(g2, f + "layer_norm.weight"); NEED(be2, f + "layer_norm.bias");
        std::vector<float> wqkv; wqkv.reserve((size_t)3 * H * dk * d);
        wqkv.insert(wqkv.end(), wq->begin(), wq->end()); wqkv.insert(wqkv.end(), wk->begin(), wk->end());
        wqkv.insert(wqkv.end(), wv->begin(), wv->end());
        std::vector<float> bqkv; bqkv.insert(bqkv.end(), bq->begin(), bq->end());
        bqkv.insert(bqkv.end(), bk->begin(), bk->end()); bqkv.insert(bqkv.end(), bv->begin(), bv->end());
        LayerW& L = c->layers[l];
        if (upload(L.wqkv, wqkv) || upload(L.bqkv, bqkv) || upload(L.fc_w, *fw) || upload(L.fc_b, *fb) ||
            upload(L.ln1_g, *g1) || upload(L.ln1_b, *be1) || upload(L.w1, *w1) || upload(L.b1, *b1) ||
            upload(L.w2, *w2) || upload(L.b2, *b2) || upload(L.ln2_g, *g2) || upload(L.ln2_b, *be2)) return 1;
    }
#undef NEED
    // timestep-embedding table for every t in [0, N)
    if (c->temb.alloc((size_t)c->N * d * sizeof(float))) return 1;
    time_table_kernel<<<c->N, 256, 0, s>>>(c->t_w1.as<float>(), c->t_b1.as<float>(), c->t_w2.as<float>(),
                                           c->t_b2.as<float>(), c->temb.as<float>(), d);
    c->launches++;
    EG_CUDA(cudaGetLastError());
    // workspace
    const size_t MB = (size_t)c->cfg.max_batch, M = MB * LP, xe = MB * c->Tmax * D;
    if (c->model_out.alloc(xe * 4) || c->x_cur.alloc(xe * 4) || c->x_cond.alloc(xe * 4) || c->d_step.alloc(64) ||
        c->t_tmp.alloc(MB * sizeof(long long))) return 1;
    if (c->cfg.engine == EGOEGO_ENGINE_SIMT) {
        if (c->Ain.alloc(M * c->kin_pad * 4) || c->Hbuf.alloc(M * d * 4) || c->Ybuf.alloc(M * d * 4) ||
            c->QKV.alloc(M * 3 * H * dk * 4) || c->Obuf.alloc(M * H * dk * 4) || c->Fbuf.alloc(M * d * 4)) return 1;
    } else {
        c->tc.reset(new TcEngine());
        TcWeights tw;
        tw.D = D; tw.d = d; tw.H = H; tw.dk = dk; tw.NL = c->NL; tw.N = c->N; tw.Tmax = c->Tmax; tw.max_batch = c->cfg.max_batch;
        tw.start_w = sw->data(); tw.start_b = c->start_b.as<float>(); tw.pos = c->pos.as<float>(); tw.temb = c->temb.as<float>();
        tw.out_w = ow->data(); tw.out_b = c->out_b.as<float>();
        for (int l = 0; l < c->NL; ++l) {
            std::string a = pre + "layer_stack." + std::to_string(l) + ".self_attn.";
            std::string f = pre + "layer_stack." + std::to_string(l) + ".pos_ffn.";
            TcLayerW lw;
            lw.wq = c->staged[a + "w_q.weight"].data(); lw.wk = c->staged[a + "w_k.weight"].data();
            lw.wv = c->staged[a + "w_v.weight"].data(); lw.fc = c->staged[a + "fc.weight"].data();
            lw.w1 = c->staged[f + "w_1.weight"].data(); lw.w2 = c->staged[f + "w_2.weight"].data();
            LayerW& L = c->layers[l];
            lw.bqkv = L.bqkv.as<float>(); lw.fc_b = L.fc_b.as<float>(); lw.b1 = L.b1.as<float>(); lw.b2 = L.b2.as<float>();
            lw.ln1_g = L.ln1_g.as<float>(); lw.ln1_b = L.ln1_b.as<float>(); lw.ln2_g = L.ln2_g.as<float>(); lw.ln2_b = L.ln2_b.as<float>();
            tw.layers.push_back(lw);
        }
        if (c->tc->init(tw, s)) return 1;
    }
    EG_CUDA(cudaStreamSynchronize(s));
    for (auto& g : c->step_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    c->graph_B = -1;
    c->committed = true;
    return 0;
}

int egoego_denoiser_forward(egoego_handle c, const float* x_all, const int64_t* t_dev, const float* pmask,
                            int B, int T, float* out, void* stream_v) {
    if (check_ready(c, B, T)) return 1;
    EG_CHECK(x_all && t_dev && out, "null argument");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t s = (cudaStream_t)stream_v;
    for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch) {
        int Bc = std::min(c->cfg.max_batch, B - b0);
        const float* xa = x_all + (size_t)b0 * T * 2 * c->D;
        if (clear_staging(c, Bc, s)) return 1;
        if (stage_input(c, xa, 2 * c->D, c->D, true, Bc, T, s)) return 1;
        if (prepare_cond(c, Bc, T, s)) return 1;
        if (stage_input(c, xa, 2 * c->D, 0, false, Bc, T, s)) return 1;
        TSrc ts{reinterpret_cast<const long long*>(t_dev) + b0, nullptr, 0, c->N - 1};
        if (run_denoiser(c, Bc, T, ts, pmask ? pmask + (size_t)b0 * (T + 1) : nullptr, s)) return 1;
        EG_CUDA(cudaMemcpyAsync(out + (size_t)b0 * T * c->D, c->model_out.p, (size_t)Bc * T * c->D * 4, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

int egoego_p_sample_step(egoego_handle c, const float* x, const int64_t* t_dev, const float* x_cond,
                         const float* noise, const egoego_rng* rng, uint64_t draw_index, const float* pmask,
                         int clip, const float* inpaint, int inpaint_len, int B, int T, float* x_out, void* stream_v) {
    if (check_ready(c, B, T)) return 1;
    EG_CHECK(x && t_dev && x_cond && x_out, "null argument");
    EG_CHECK(noise || rng, "either an explicit noise tensor or an rng must be given");
    EG_CHECK(inpaint_len >= 0 && inpaint_len <= T, "inpaint_len out of range");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t s = (cudaStream_t)stream_v;
    const int D = c->D;
    for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch) {
        int Bc = std::min(c->cfg.max_batch, B - b0);
        size_t off = (size_t)b0 * T * D;
        if (clear_staging(c, Bc, s)) return 1;
        if (stage_input(c, x_cond + off, D, 0, true, Bc, T, s)) return 1;
        if (prepare_cond(c, Bc, T, s)) return 1;
        if (stage_input(c, x + off, D, 0, false, Bc, T, s)) return 1;
        TSrc ts{reinterpret_cast<const long long*>(t_dev) + b0, nullptr, 0, c->N - 1};
        if (run_denoiser(c, Bc, T, ts, pmask ? pmask + (size_t)b0 * (T + 1) : nullptr, s)) return 1;
        NoiseSrc ns{};
        if (noise) { ns.tape = noise + off; ns.draw_stride = 0; ns.draw_static = 0; }
        else { ns.tape = nullptr; ns.seed = rng->seed; ns.window_offset = rng->window_offset + b0; ns.draw_static = (int)draw_index; }
        DdpmArgs a;
        fill_ddpm(c, a, x + off, x_out + off, ts, ns, clip, inpaint ? inpaint + (size_t)b0 * inpaint_len * D : nullptr,
                  inpaint_len, Bc, T, false);
        ddpm_update_kernel<<<ddpm_grid(T, D, Bc), 256, 0, s>>>(a);
        c->launches++;
        EG_CUDA(cudaGetLastError());
    }
    return 0;
}

// One chunk (Bc <= max_batch) of the sampling loop on stream `s` (must be capturable, i.e. not legacy).
static int sample_chunk(egoego_ctx* c, const float* x_start, const float* cond_mask, int Bc, int T, NoiseSrc ns_base,
                        const float* x_init, const float* inpaint, int inpaint_len, float* out, cudaStream_t s) {
    const int D = c->D, N = c->N;
    const long long quads = ((long long)T * D + 3) / 4 * Bc;
    float* xc = c->x_cur.as<float>();
    float* xcond = c->x_cond.as<float>();
    int* d_step = c->d_step.as<int>();
    init_sample_kernel<<<grid1d(quads, 256), 256, 0, s>>>(xc, xcond, x_init, x_start, cond_mask, ns_base, Bc, T, D);
    c->launches++;
    EG_CUDA(cudaGetLastError());
    if (clear_staging(c, Bc, s)) return 1;
    if (stage_input(c, xcond, D, 0, true, Bc, T, s)) return 1;
    if (prepare_cond(c, Bc, T, s)) return 1;
    if (stage_input(c, xc, D, 0, false, Bc, T, s)) return 1;
    EG_CUDA(cudaMemsetAsync(d_step, 0, 64, s));                 // step counter + the block counter of ddpm_update_kernel (word 8)
    if (c->tc && c->tc->reset_stream_counters(s)) return 1;
    TSrc ts{nullptr, d_step, N - 1, N - 1};
    NoiseSrc ns = ns_base;
    ns.d_step = d_step; ns.draw_static = 2;
    DdpmArgs a;
    fill_ddpm(c, a, xc, xc, ts, ns, 1, inpaint, inpaint_len, Bc, T, true);

    // A step stages the next step's start_conv operand only in its OWN format; at the fp16 -> split switch the split
    // planes are produced once from x_cur (restage) before the first split step.
    const bool fused = c->fuse_ddpm && c->cfg.engine == EGOEGO_ENGINE_TCGEN05;   // DDPM update in the epilogue of linear_out
    auto one_step = [&](cudaStream_t st, int fmt) -> int {
        a.stage_mode = (c->cfg.engine == EGOEGO_ENGINE_SIMT) ? 2 : (fmt ? 1 : 0);
        if (run_denoiser(c, Bc, T, ts, nullptr, st, fmt, fused ? &a : nullptr)) return 1;
        LaunchCfg ld(1, 256, 0, st), la(1, 32, 0, st);
        ld.cfg.gridDim = ddpm_grid(T, D, Bc);
        if (!fused) {           // the update's last block advances the step counter
            a.advance = d_step; a.done = reinterpret_cast<unsigned*>(d_step + 8);
            EG_CUDA(cudaLaunchKernelEx(&ld.cfg, ddpm_update_kernel, a));
        } else {
            EG_CUDA(cudaLaunchKernelEx(&la.cfg, advance_step_kernel, d_step));
        }
        c->launches += 1;
        EG_CUDA(cudaGetLastError());
        return 0;
    };
    const bool tc_engine = c->cfg.engine == EGOEGO_ENGINE_TCGEN05 && c->tc;
    // The last `precise_last` steps come in two precisions: t < split_last in the 3-term split (fp32-grade products), the others with
    // fp16 activations against the fp16 hi/lo PAIR of every weight (FMT_HALF kernels, two passes over K).  What the sampler
    // integrates coherently over steps is the WEIGHT rounding; the activation rounding is fresh at every step and its
    // effect on the final sample is scaled by ~1/(t (t+1)).
    auto pair_of_step = [&](int i) -> bool { const int t = N - 1 - i; return tc_engine && t < c->precise_last && t >= c->split_last; };
    auto fmt_of_step = [&](int i) -> int { return ((N - 1 - i) >= c->precise_last || pair_of_step(i)) ? 1 : 0; };   // i-th executed step has t = N-1-i
    // The 3-term split steps come in two kinds: the last `dual_last` steps (t < dual_last) keep hi*hi and the cross terms in separate
    // TMEM accumulators (the tensor core's truncating fp32 accumulation is a bias that is the same at every step; an error made at
    // step t reaches the final sample scaled by ~1/(t (t+1)), so only the last steps need the tighter, 18 % slower kernels).
    auto dual_of_step = [&](int i) -> bool { return tc_engine && (N - 1 - i) < c->dual_last; };
    // single-pass fp16 steps cycle through the dithered fp16 weight sets so that the weight rounding averages out over steps
    // (slot 0 = dual split step, slot 1 + r = fp16 step reading copy r, last slot = single-accumulator split step).  Averaging needs
    // several full cycles: a run of fewer than 4 R fp16 steps reads the plain round-to-nearest copy instead (a single dithered copy
    // is up to one ulp off, plain RN half).
    int n_sets = tc_engine ? std::min(c->tc->n_weight_sets(), (int)egoego_ctx::MAX_WEIGHT_SETS) : 1;
    int n_half = 0;
    for (int i = 0; i < N; ++i) n_half += (fmt_of_step(i) && !pair_of_step(i)) ? 1 : 0;
    const bool dither = n_sets > 1 && n_half >= 4 * n_sets;
    if (!dither) n_sets = 1;
    constexpr int SLOT_SINGLE = 1 + egoego_ctx::MAX_WEIGHT_SETS, SLOT_PAIR = 2 + egoego_ctx::MAX_WEIGHT_SETS;
    auto slot_of_step = [&](int i) -> int {
        if (pair_of_step(i)) return SLOT_PAIR;
        return fmt_of_step(i) ? 1 + (i % n_sets) : ((tc_engine && !dual_of_step(i)) ? SLOT_SINGLE : 0);
    };
    auto select_set = [&](int slot) {
        if (!c->tc) return;
        c->tc->use_weight_set((dither && slot > 0 && slot < SLOT_SINGLE) ? slot - 1 : -1);
        c->tc->set_dual_acc(slot != SLOT_SINGLE);
        c->tc->set_weight_pair(slot == SLOT_PAIR);
    };
    const void* key[4] = {ns.tape, inpaint, (const void*)(uintptr_t)(ns.seed ^ (ns.window_offset * 0x9E3779B97F4A7C15ull)),
                          (const void*)(uintptr_t)(((uint64_t)inpaint_len << 32) ^ (uint64_t)ns.draw_stride)};
    if (c->use_graph) {
        bool reuse = c->graph_B == Bc && c->graph_T == T && !memcmp(key, c->graph_key, sizeof(key));
        if (!reuse) for (auto& g : c->step_graph) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
        for (int slot = 0; slot <= SLOT_PAIR; ++slot) {
            bool needed = false;
            for (int i = 0; i < N && !needed; ++i) needed = slot_of_step(i) == slot;
            if (!needed || c->step_graph[slot]) continue;
            cudaGraph_t g = nullptr;
            int64_t before = c->launches;
            select_set(slot);                           // tensor maps and the accumulator mode are launch-time choices: captured by value
            EG_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            int rc = one_step(s, ((slot >= 1 && slot < SLOT_SINGLE) || slot == SLOT_PAIR) ? 1 : 0);
            cudaError_t ce = cudaStreamEndCapture(s, &g);
            c->launches = before;                       // captured, not launched
            select_set(0);
            EG_CHECK(rc == 0, std::string("capture failed: ") + g_err);
            EG_CUDA(ce);
            EG_CUDA(cudaGraphInstantiate(&c->step_graph[slot], g, 0));
            cudaGraphDestroy(g);
        }
        c->graph_B = Bc; c->graph_T = T; memcpy(c->graph_key, key, sizeof(key));
        for (int i = 0; i < N; ++i) {
            const int fmt = fmt_of_step(i);
            if (i > 0 && fmt != fmt_of_step(i - 1) && stage_input(c, xc, D, 0, false, Bc, T, s)) return 1;
            EG_CUDA(cudaGraphLaunch(c->step_graph[slot_of_step(i)], s));
            c->launches += (c->cfg.engine == EGOEGO_ENGINE_SIMT ? (2 + 7 * c->NL) : c->tc->launches_per_denoiser(fmt)) + 1;
        }
    } else {
        for (int i = 0; i < N; ++i) {
            if (i > 0 && fmt_of_step(i) != fmt_of_step(i - 1) && stage_input(c, xc, D, 0, false, Bc, T, s)) return 1;
            select_set(slot_of_step(i));
            int rc = one_step(s, fmt_of_step(i));
            select_set(0);
            if (rc) return 1;
        }
    }
    EG_CUDA(cudaMemcpyAsync(out, xc, (size_t)Bc * T * D * 4, cudaMemcpyDeviceToDevice, s));
    return 0;
}

int egoego_sample(egoego_handle c, const float* x_start, const float* cond_mask, int B, int T, const egoego_rng* rng,
                  const float* x_init, const float* inpaint, int inpaint_len, float* out, void* stream_v) {
    if (check_ready(c, B, T)) return 1;
    EG_CHECK(x_start && cond_mask && rng && out, "null argument");
    EG_CHECK(inpaint_len >= 0 && inpaint_len <= T && (inpaint || inpaint_len == 0), "bad inpaint arguments");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t user = (cudaStream_t)stream_v;
    cudaStream_t s = c->own_stream;            // capturable stream; ordered after/before the caller's stream
    EG_CUDA(cudaEventRecord(c->ev_in, user));
    EG_CUDA(cudaStreamWaitEvent(s, c->ev_in, 0));
    const int D = c->D;
    for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch) {
        int Bc = std::min(c->cfg.max_batch, B - b0);
        size_t off = (size_t)b0 * T * D;
        NoiseSrc ns{};
        ns.tape = rng->tape ? rng->tape + off : nullptr;
        ns.draw_stride = (long long)B * T * D;
        ns.seed = rng->seed; ns.window_offset = rng->window_offset + b0;
        if (sample_chunk(c, x_start + off, cond_mask + off, Bc, T, ns, x_init ? x_init + off : nullptr,
                         inpaint ? inpaint + (size_t)b0 * inpaint_len * D : nullptr, inpaint_len, out + off, s)) return 1;
    }
    EG_CUDA(cudaEventRecord(c->ev_out, s));
    EG_CUDA(cudaStreamWaitEvent(user, c->ev_out, 0));
    return 0;
}

int egoego_sample_host(egoego_handle c, const float* x_start_h, const float* cond_mask_h, int B, int T,
                       const egoego_rng* rng, float* out_h, void* stream_v) {
    if (check_ready(c, B, T)) return 1;
    EG_CHECK(x_start_h && cond_mask_h && rng && out_h, "null argument");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t user = (cudaStream_t)stream_v;
    const int D = c->D;
    const size_t MB = (size_t)c->cfg.max_batch, xe = MB * T * D;
    if (c->h_xstart.bytes < xe * 4) { if (c->h_xstart.alloc(xe * 4) || c->h_mask.alloc(xe * 4) || c->h_out.alloc(xe * 4)) return 1; }
    for (int b0 = 0; b0 < B; b0 += c->cfg.max_batch) {
        int Bc = std::min(c->cfg.max_batch, B - b0);
        size_t off = (size_t)b0 * T * D, n = (size_t)Bc * T * D;
        EG_CUDA(cudaMemcpyAsync(c->h_xstart.p, x_start_h + off, n * 4, cudaMemcpyHostToDevice, user));
        EG_CUDA(cudaMemcpyAsync(c->h_mask.p, cond_mask_h + off, n * 4, cudaMemcpyHostToDevice, user));
        egoego_rng r = *rng;
        r.window_offset = rng->window_offset + b0;
        if (rng->tape) {   // host tape [N+2, B, T, D] -> device tape [N+2, Bc, T, D] of this chunk
            size_t need = (size_t)(c->N + 2) * n * 4;
            if (c->h_tape.bytes < need && c->h_tape.alloc(need)) return 1;
            EG_CUDA(cudaMemcpy2DAsync(c->h_tape.p, n * 4, rng->tape + off, (size_t)B * T * D * 4, n * 4, c->N + 2,
                                      cudaMemcpyHostToDevice, user));
            r.tape = c->h_tape.as<float>();
        }
        // the device entry point expects tape stride = its own B: call it per chunk
        if (egoego_sample(c, c->h_xstart.as<float>(), c->h_mask.as<float>(), Bc, T, &r, nullptr, nullptr, 0,
                          c->h_out.as<float>(), user)) return 1;
        EG_CUDA(cudaMemcpyAsync(out_h + off, c->h_out.p, n * 4, cudaMemcpyDeviceToHost, user));
    }
    EG_CUDA(cudaStreamSynchronize(user));
    return 0;
}

int egoego_set_skeleton(egoego_handle c, const int32_t* parents, const float* rest_offsets, const float* jmin, const float* jmax) {
    EG_CHECK(c && parents && rest_offsets && jmin && jmax, "null argument");
    EG_CHECK(parents[0] == -1, "parents[0] must be -1");
    Skeleton& sk = c->sk;
    sk.max_depth = 0;
    for (int j = 0; j < NJ; ++j) {
        EG_CHECK(j == 0 || (parents[j] >= 0 && parents[j] < j), "parents must be topologically ordered");
        sk.parents[j] = parents[j];
        sk.depth[j] = j == 0 ? 0 : sk.depth[parents[j]] + 1;
        sk.max_depth = std::max(sk.max_depth, sk.depth[j]);
        for (int k = 0; k < 3; ++k) sk.off[j][k] = rest_offsets[j * 3 + k];
    }
    memcpy(sk.jmin, jmin, sizeof(sk.jmin));
    memcpy(sk.jmax, jmax, sizeof(sk.jmax));
    c->have_sk = true;
    return 0;
}

int egoego_postprocess(egoego_handle c, const float* x, const float* recover_quat, int B, int T, float* aa, float* root,
                       float* head, float* jpos, float* gquat, void* stream_v) {
    EG_CHECK(c && x, "null argument");
    EG_CHECK(c->have_sk, "egoego_set_skeleton has not been called");
    EG_CHECK(c->D == 198, "post-processing is defined for d_feats = 198 (22*3 + 22*6)");
    EG_CHECK(B >= 1 && T >= 1, "bad shape");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    postprocess_kernel<<<(B * T + 7) / 8, 256, 0, (cudaStream_t)stream_v>>>(c->sk, x, recover_quat, B, T, aa, root, head, jpos, gquat);
    c->launches++;
    EG_CUDA(cudaGetLastError());
    return 0;
}

int egoego_fk_smpl(egoego_handle c, const float* root, const float* aa, int64_t N, float* gquat, float* jpos, void* stream_v) {
    EG_CHECK(c && root && aa, "null argument");
    EG_CHECK(c->have_sk, "egoego_set_skeleton has not been called");
    EG_CHECK(N >= 1, "bad shape");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    fk_smpl_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream_v>>>(c->sk, root, aa, N, gquat, jpos);
    c->launches++;
    EG_CUDA(cudaGetLastError());
    return 0;
}

int egoego_eval_metrics(int device, const float* gt_quat, const float* gt_jpos, const float* gt_floor,
                        const float* pred_quat, const float* pred_jpos, const float* pred_floor,
                        int B, int T, float* out, void* stream_v) {
    EG_CHECK(gt_quat && gt_jpos && gt_floor && pred_quat && pred_jpos && pred_floor && out, "null argument");
    EG_CHECK(B >= 1 && T >= 3, "eval metrics need B >= 1 sequences of T >= 3 frames (accelerations use 3 frames)");
    int ndev = 0;
    EG_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "no CUDA device: libegoego_b200 has no CPU fallback");
    EG_CHECK(device >= 0 && device < ndev, "bad device ordinal");
    EG_CUDA(cudaSetDevice(device));
    eval_metrics_kernel<<<B, MET_WARPS * 32, 0, (cudaStream_t)stream_v>>>(gt_quat, gt_jpos, gt_floor, pred_quat, pred_jpos, pred_floor, T, out);
    EG_CUDA(cudaGetLastError());
    return 0;
}


int egoego_floor_contacts(int device, const float* jpos, int B, int T, int fps, float* floor_out, float* contacts, int* discard, void* stream_v) {
    EG_CHECK(jpos && floor_out, "null argument");
    EG_CHECK(B >= 1 && T >= 2 && T <= FLOOR_MAX_T, "floor height needs B >= 1 sequences of 2 <= T <= 2048 frames");
    int ndev = 0;
    EG_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "no CUDA device: libegoego_b200 has no CPU fallback");
    EG_CHECK(device >= 0 && device < ndev, "bad device ordinal");
    EG_CUDA(cudaSetDevice(device));
    const size_t smem = floor_smem_bytes(T);
    static PerDeviceMax smem_set;
    if (smem > 48 * 1024 && smem_set.raise(device, smem))
        EG_CUDA(cudaFuncSetAttribute(floor_contacts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    floor_contacts_kernel<<<B, FLOOR_THREADS, smem, (cudaStream_t)stream_v>>>(jpos, T, fps, floor_out, contacts, discard);
    EG_CUDA(cudaGetLastError());
    return 0;
}

int egoego_canonicalize_head(egoego_handle c, const float* head_pos, const float* head_quat, int64_t stride_frames,
                             int B, int T, float* x_start, float* recover_quat, void* stream_v) {
    EG_CHECK(c && head_pos && head_quat && x_start, "null argument");
    EG_CHECK(c->have_sk, "egoego_set_skeleton has not been called");
    EG_CHECK(c->D == 198 && B >= 1 && T >= 1 && stride_frames >= T, "bad shape");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    canonicalize_head_kernel<<<(B * T + 127) / 128, 128, 0, (cudaStream_t)stream_v>>>(c->sk, head_pos, head_quat, stride_frames,
                                                                                      B, T, x_start, recover_quat);
    c->launches++;
    EG_CUDA(cudaGetLastError());
    return 0;
}

int egoego_tail_condition(egoego_handle c, const float* gquat, const float* gjpos, int B, int n, float* inpaint_out, void* stream_v) {
    EG_CHECK(c && gquat && gjpos && inpaint_out, "null argument");
    EG_CHECK(c->have_sk, "egoego_set_skeleton has not been called");
    EG_CHECK(c->D == 198 && B >= 1 && n >= 1, "bad shape");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    tail_condition_kernel<<<(B * n + 7) / 8, 256, 0, (cudaStream_t)stream_v>>>(c->sk, gquat, gjpos, B, n, inpaint_out);
    c->launches++;
    EG_CUDA(cudaGetLastError());
    return 0;
}

int64_t egoego_launch_count(egoego_handle c) { return c ? c->launches : -1; }

int egoego_tensors_checksum(int device, int n, const void* const* ptrs_dev, const int64_t* numels, uint64_t* out_host, void* stream_v) {
    EG_CHECK(ptrs_dev && numels && out_host && n >= 0, "null argument");
    int ndev = 0;
    EG_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && device >= 0 && device < ndev, "no such CUDA device");
    EG_CUDA(cudaSetDevice(device));
    cudaStream_t s = (cudaStream_t)stream_v;
    static thread_local unsigned long long* d_acc[64] = {};
    unsigned long long*& acc = d_acc[device & 63];
    if (!acc) EG_CUDA(cudaMalloc(&acc, 8));
    EG_CUDA(cudaMemsetAsync(acc, 0, 8, s));
    unsigned long long start = 0;
    for (int t0 = 0; t0 < n; t0 += CHK_MAX_TENSORS) {
        ChkTable tab{};
        tab.count = std::min(CHK_MAX_TENSORS, n - t0);
        for (int t = 0; t < tab.count; ++t) {
            EG_CHECK(numels[t0 + t] >= 0 && (ptrs_dev[t0 + t] || numels[t0 + t] == 0), "egoego_tensors_checksum: bad tensor");
            tab.p[t] = reinterpret_cast<const uint32_t*>(ptrs_dev[t0 + t]); tab.n[t] = (unsigned long long)numels[t0 + t]; tab.start[t] = start;
            start += tab.n[t];
        }
        tensors_checksum_kernel<<<dim3(64, (unsigned)std::min(tab.count, 32)), 256, 0, s>>>(tab, acc);
        EG_CUDA(cudaGetLastError());
    }
    unsigned long long h = 0;
    EG_CUDA(cudaMemcpyAsync(&h, acc, 8, cudaMemcpyDeviceToHost, s));
    EG_CUDA(cudaStreamSynchronize(s));
    *out_host = h;
    return 0;
}


int egoego_time_kernel(egoego_handle c, int B, int T, int which, int half_fmt, int iters, float* ms_per_launch, void* stream_v) {
    EG_CHECK(c && ms_per_launch, "null argument");
    EG_CHECK(c->committed && c->cfg.engine == EGOEGO_ENGINE_TCGEN05, "needs a committed tensor-core engine");
    EG_CHECK(B >= 1 && B <= c->cfg.max_batch && T >= 1 && T <= c->Tmax && iters >= 1, "bad arguments");
    EG_CHECK(which >= 0 && which <= EGOEGO_KERNEL_DDPM_UPDATE, "unknown kernel id");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t s = (cudaStream_t)stream_v;
    // DDPM update (clamp + posterior mean + Philox noise + staging of the next step's A operand), t fixed at N/2
    TSrc ts{nullptr, nullptr, c->N / 2, c->N - 1};
    NoiseSrc ns{};
    ns.tape = nullptr; ns.seed = 1; ns.window_offset = 0; ns.draw_static = 2;
    DdpmArgs a;
    fill_ddpm(c, a, c->x_cur.as<float>(), c->x_cur.as<float>(), ts, ns, 1, nullptr, 0, B, T, true);
    a.stage_mode = half_fmt ? 1 : 0;
    if (which < EGOEGO_KERNEL_DDPM_UPDATE)   // linear_out is timed as the sampling loop runs it: with the fused DDPM epilogue when that is on
        return c->tc->time_stage(B, T, which, half_fmt ? 1 : 0, iters, c->model_out.as<float>(), s, ms_per_launch,
                                 (which == EGOEGO_KERNEL_OUT && c->fuse_ddpm) ? &a : nullptr);
    cudaEvent_t e0, e1;
    EG_CUDA(cudaEventCreate(&e0)); EG_CUDA(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) ddpm_update_kernel<<<ddpm_grid(T, c->D, B), 256, 0, s>>>(a);
    EG_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i) ddpm_update_kernel<<<ddpm_grid(T, c->D, B), 256, 0, s>>>(a);
    EG_CUDA(cudaEventRecord(e1, s));
    EG_CUDA(cudaEventSynchronize(e1));
    EG_CUDA(cudaGetLastError());
    float t = 0.f;
    EG_CUDA(cudaEventElapsedTime(&t, e0, e1));
    *ms_per_launch = t / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

int egoego_time_dominant_kernel(egoego_handle c, int B, int half_fmt, int iters, float* ms_per_launch, void* stream_v) {
    EG_CHECK(c, "null argument");
    return egoego_time_kernel(c, B, c->Tmax, EGOEGO_KERNEL_QKV, half_fmt, iters, ms_per_launch, stream_v);
}

int egoego_precise_last_steps(egoego_handle c) { return c ? c->precise_last : -1; }
int egoego_engine_info(egoego_handle c, char* buf, int n) {
    EG_CHECK(c && buf && n > 0, "null argument");
    std::string s = (c->cfg.engine == EGOEGO_ENGINE_TCGEN05 && c->tc) ? c->tc->info() : std::string(c->cfg.engine == EGOEGO_ENGINE_SIMT ? "engine=simt" : "engine=tcgen05 (weights not committed)");
    s += " precise_last_steps=" + std::to_string(c->precise_last) + " split_steps=" + std::to_string(std::min(c->split_last, c->precise_last)) +
         " dual_accumulator_steps=" + std::to_string(std::min(c->dual_last, std::min(c->split_last, c->precise_last)));
    snprintf(buf, (size_t)n, "%s", s.c_str());
    return 0;
}
int egoego_debug_timeline(int device, int enable, uint64_t* out, int n) {
    EG_CUDA(cudaSetDevice(device));
    return tc_debug_timeline(enable, reinterpret_cast<unsigned long long*>(out), n);
}
int egoego_dither_weights_f16(const float* w, int64_t n, int set, int n_sets, uint16_t* out) {
    EG_CHECK(w && out && n >= 0 && n_sets >= 1 && n_sets <= 16 && set >= 0 && set < n_sets, "egoego_dither_weights_f16: bad arguments");
    dither_weights_host(w, (long long)n, set, n_sets, out);
    return 0;
}
int egoego_weight_sets(egoego_handle c) { return (c && c->cfg.engine == EGOEGO_ENGINE_TCGEN05 && c->tc) ? c->tc->n_weight_sets() : 1; }

int egoego_launches_per_step(egoego_handle c, int which) {
    if (!c || which < 0 || which > EGOEGO_KERNEL_DDPM_UPDATE) return -1;
    if (which == EGOEGO_KERNEL_START || which == EGOEGO_KERNEL_OUT) return 1;
    if (which == EGOEGO_KERNEL_DDPM_UPDATE) return (c->fuse_ddpm && c->cfg.engine == EGOEGO_ENGINE_TCGEN05) ? 0 : 1;
    return c->NL;
}

int egoego_selftest_gemm(int device, int M, int N, int K, uint64_t seed, int two_cta, int half_fmt, float* max_abs_err, float* max_abs_ref, float* ms) {
    EG_CHECK(max_abs_err && max_abs_ref && ms, "null argument");
    int ndev = 0;
    EG_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && device >= 0 && device < ndev, "no such CUDA device");
    EG_CUDA(cudaSetDevice(device));
    return selftest_gemm(M, N, K, seed, two_cta, half_fmt, max_abs_err, max_abs_ref, ms);
}

}  // extern "C"

// ================================================================================================
// Training step: see train.cuh
// ================================================================================================
#include "train.cuh"

namespace egoego {

struct TrainLayerBufs { DevBuf QKV, O, Y1, st1, H1, F, Y2, st2; };

struct TrainWs {
    int B = 0;                                   // windows the workspace is sized for
    std::vector<DevBuf> Hin;                     // NL + 1 residual streams [M,512]
    std::vector<TrainLayerBufs> L;
    DevBuf OUT, dOUT, dH, dH1, dY, dZ, dF, dO, dQKV, Ta, Tb, WT, gp, bp, loss, Tmp, temb_b, iota;
    std::map<std::string, DevBuf> grads;         // fused / padded gradient buffers
    // The step is captured into a CUDA graph per (B, T, loss, mask, dropout) and replayed: its inputs are first copied into these
    // buffers (the caller's tensors move between steps), the dropout seed and the loss live in device memory.
    DevBuf s_x, s_cm, s_noise, s_cn, s_pm, s_t, s_sa, s_sb, s_wt, s_seed, s_loss;
    struct GraphKey { int B, T, l2, pm, drop; uint32_t thresh; bool operator<(const GraphKey& o) const {
        return std::tie(B, T, l2, pm, drop, thresh) < std::tie(o.B, o.T, o.l2, o.pm, o.drop, o.thresh); } };
    std::map<GraphKey, cudaGraphExec_t> graphs;  // nullptr value: the key has run once eagerly (allocations, attributes) and is captured next time
    std::set<GraphKey> no_graph;                 // shapes whose capture failed: eager from then on
};

static std::map<egoego_ctx*, std::unique_ptr<TrainWs>> g_train;

void train_drop_graphs(egoego_ctx* c) {
    auto it = g_train.find(c);
    if (it == g_train.end() || !it->second) return;
    for (auto& kv : it->second->graphs) if (kv.second) cudaGraphExecDestroy(kv.second);
    it->second->graphs.clear();
}

void train_release(egoego_ctx* c) {
    auto it = g_train.find(c);
    if (it == g_train.end()) return;
    if (TrainWs* w = it->second.get()) {
        for (auto& h : w->Hin) h.release();
        for (auto& l : w->L) for (DevBuf* b : {&l.QKV, &l.O, &l.Y1, &l.st1, &l.H1, &l.F, &l.Y2, &l.st2}) b->release();
        for (DevBuf* b : {&w->OUT, &w->dOUT, &w->dH, &w->dH1, &w->dY, &w->dZ, &w->dF, &w->dO, &w->dQKV, &w->Ta, &w->Tb, &w->WT, &w->gp, &w->bp, &w->loss, &w->Tmp, &w->temb_b, &w->iota}) b->release();
        for (auto& kv : w->grads) kv.second.release();
        for (DevBuf* b : {&w->s_x, &w->s_cm, &w->s_noise, &w->s_cn, &w->s_pm, &w->s_t, &w->s_sa, &w->s_sb, &w->s_wt, &w->s_seed, &w->s_loss}) b->release();
        for (auto& kv : w->graphs) if (kv.second) cudaGraphExecDestroy(kv.second);
    }
    g_train.erase(it);
}

// products of the training step: tensor cores (tc_gemm_f32: 3-term bf16 split, fp32-grade) by default, the fp32 CUDA-core
// sgemm with EGOEGO_TRAIN_GEMM=simt (bisecting).  Epilogue functors of the validation engine are applied by an element-wise
// pass over the raw product in the tensor-core path.
static bool train_use_tc() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("EGOEGO_TRAIN_GEMM"); v = (e && strcmp(e, "simt") == 0) ? 0 : 1; }
    return v == 1;
}
template <class Epi>
static __global__ void tr_epi_kernel(const float* __restrict__ C, int ldc, int rows, int N, Epi e) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)rows * N) return;
    const int r = (int)(i / N), c = (int)(i % N);
    e(r, c, C[(long long)r * ldc + c]);
}
template <class Epi>
static int tr_gemm(TrainWs* w, const float* A, int lda, const float* W, int ldw, int rows, int N, int K, const Epi& e, cudaStream_t s) {
    if (!train_use_tc()) { sgemm_tn_kernel<<<dim3((N + 127) / 128, rows / 128), 256, 0, s>>>(A, lda, W, ldw, N, K, e); return 0; }
    // The functor runs in the GEMM's own epilogue -- except the dropout one while dropout is on: a Philox call per element on the 8
    // epilogue warps of an SM costs more than the separate full-occupancy pass saves (measured 5.37 vs 5.28 ms per step with
    // everything fused; the same lesson as the DDPM update in linear_out's epilogue).  EGOEGO_TRAIN_FUSE_EPI=0: never fuse.
    static const bool fuse = []() { const char* v = getenv("EGOEGO_TRAIN_FUSE_EPI"); return !(v && v[0] == '0'); }();
    bool heavy = false;
    if constexpr (std::is_same<Epi, EpiBiasDropResid>::value) heavy = e.drop.on != 0;
    if (fuse && !heavy) return tc_gemm_f32_epi(A, lda, W, ldw, rows, N, K, e, s);
    const int ldt = ((N + 3) / 4) * 4;
    if (tc_gemm_f32(A, lda, W, ldw, rows, N, K, w->Tmp.as<float>(), ldt, ldt, 0, s)) return 1;
    tr_epi_kernel<<<(unsigned)(((long long)rows * N + 255) / 256), 256, 0, s>>>(w->Tmp.as<float>(), ldt, rows, N, e);
    return 0;
}
// C (= or +=) A W^T without an epilogue (N % 4 == 0)
static int tr_gemm_plain(const float* A, int lda, const float* W, int ldw, int rows, int N, int K, float* C, int ldc, bool acc, cudaStream_t s) {
    if (train_use_tc()) return tc_gemm_f32(A, lda, W, ldw, rows, N, K, C, ldc, N, acc ? 1 : 0, s);
    const int rp = ((rows + 127) / 128) * 128;
    if (acc) sgemm_tn_kernel<<<dim3((N + 127) / 128, rp / 128), 256, 0, s>>>(A, lda, W, ldw, N, K, EpiAccum{C, ldc});
    else     sgemm_tn_kernel<<<dim3((N + 127) / 128, rp / 128), 256, 0, s>>>(A, lda, W, ldw, N, K, EpiPlainBias{C, ldc, nullptr});
    return 0;
}
static void tr_transpose(const float* src, int R, int C, int ld, float* dst, int ldo, cudaStream_t s) {
    tr_transpose_kernel<<<dim3((C + 31) / 32, (R + 31) / 32), dim3(32, 8), 0, s>>>(src, R, C, ld, dst, ldo);
}
static float* g_cs_part = nullptr;      // [TR_CS_RB, 3072] partial column sums (one training step at a time per process)
static void tr_colsum(const float* X, int M, int C, int ld, float* out, cudaStream_t s) {
    if (!g_cs_part) cudaMalloc(&g_cs_part, (size_t)TR_CS_RB * 3072 * sizeof(float));
    tr_colsum_part_kernel<<<dim3((C + 31) / 32, TR_CS_RB), dim3(32, 8), 0, s>>>(X, M, C, ld, g_cs_part);
    tr_colsum_final_kernel<<<(C + 127) / 128, 128, 0, s>>>(g_cs_part, C, out);
}
// g [rowsW, Kd] = dY^T [rowsW, M] X [M, Kd]  (dY [M, ldy] with rowsW valid columns, X [M, ldx] with Kd valid columns)
static int tr_weight_grad(TrainWs* w, const float* dY, int ldy, int rowsW, const float* X, int ldx, int Kd, int M, float* g, int ldg, cudaStream_t s) {
    // tensor cores: both operands are stored with the reduction index (the token row) as the ROW -- converted to K-major planes directly
    if (train_use_tc()) return tc_gemm_f32_ex(dY, ldy, 1, X, ldx, 1, rowsW, Kd, M, g, ldg, Kd, 0, s);
    tr_transpose(dY, M, rowsW, ldy, w->Ta.as<float>(), M, s);          // [rowsW, M] (SIMT path: rows up to the next multiple of 128 may hold
    tr_transpose(X, M, Kd, ldx, w->Tb.as<float>(), M, s);              // stale data; their products land in the padded rows of g)
    return tr_gemm_plain(w->Ta.as<float>(), M, w->Tb.as<float>(), M, rowsW, Kd, M, g, ldg, false, s);
}

// C [rows, N] (= or +=) A [rows, K] Wsrc [K, N]: the input gradient of a linear layer whose weight is stored [out = K, in = N]
static int tr_gemm_wt(TrainWs* w, const float* A, int lda, const float* Wsrc, int ldw, int rows, int N, int K, float* C, int ldc, bool acc, cudaStream_t s) {
    if (train_use_tc()) return tc_gemm_f32_ex(A, lda, 0, Wsrc, ldw, 1, rows, N, K, C, ldc, N, acc ? 1 : 0, s);
    const int Kp = ((K + 63) / 64) * 64;                 // fp32 CUDA-core path: explicit transpose into WT [N, Kp], zero padded
    float* WT = w->WT.as<float>();
    EG_CUDA(cudaMemsetAsync(WT, 0, (size_t)N * Kp * 4, s));
    tr_transpose(Wsrc, K, N, ldw, WT, Kp, s);
    return tr_gemm_plain(A, lda, WT, Kp, rows, N, Kp, C, ldc, acc, s);
}

static int train_alloc(egoego_ctx* c, TrainWs* w, int B) {
    if (w->B >= B) return 0;
    for (auto& kv : w->graphs) if (kv.second) cudaGraphExecDestroy(kv.second);      // captured pointers die with the old workspace
    w->graphs.clear();
    w->no_graph.clear();
    {
        const size_t xe = (size_t)B * (c->cfg.max_timesteps - 1) * c->D * 4;
        if (w->s_x.alloc(xe) || w->s_cm.alloc(xe) || w->s_noise.alloc(xe) || w->s_cn.alloc(xe) || w->s_pm.alloc((size_t)B * c->cfg.max_timesteps * 4) ||
            w->s_t.alloc((size_t)B * 8) || w->s_sa.alloc((size_t)B * 4) || w->s_sb.alloc((size_t)B * 4) || w->s_wt.alloc((size_t)B * 4) ||
            w->s_seed.alloc(8) || w->s_loss.alloc(4)) return 1;
    }
    const size_t M = (size_t)((B + 1) / 2) * 2 * LP, d = c->d, nq = 3 * c->H * c->dk, hd = c->H * c->dk;   // rows rounded to the 256-row tile grid
    w->Hin.resize(c->NL + 1); w->L.resize(c->NL);
    for (auto& h : w->Hin) if (h.alloc(M * d * 4)) return 1;
    for (auto& l : w->L)
        if (l.QKV.alloc(M * nq * 4) || l.O.alloc(M * hd * 4) || l.Y1.alloc(M * d * 4) || l.st1.alloc(M * 2 * 4) || l.H1.alloc(M * d * 4) ||
            l.F.alloc(M * d * 4) || l.Y2.alloc(M * d * 4) || l.st2.alloc(M * 2 * 4)) return 1;
    if (w->OUT.alloc(M * 256 * 4) || w->dOUT.alloc(M * 256 * 4) || w->dH.alloc(M * d * 4) || w->dH1.alloc(M * d * 4) || w->dY.alloc(M * d * 4) || w->dZ.alloc(M * d * 4) ||
        w->dF.alloc(M * d * 4) || w->dO.alloc(M * hd * 4) || w->dQKV.alloc(M * nq * 4) || w->Ta.alloc(nq * M * 4) || w->Tb.alloc(hd * M * 4) ||
        w->WT.alloc(nq * d * 4) || w->gp.alloc(M * d * 4) || w->bp.alloc(M * d * 4) || w->loss.alloc(8) || w->Tmp.alloc(M * nq * 4) || w->temb_b.alloc((size_t)B * d * 4) ||
        w->iota.alloc((size_t)B * 8)) return 1;
    tr_iota_kernel<<<(B + 255) / 256, 256>>>(w->iota.as<long long>(), B);
    if (c->Ain.bytes < M * c->kin_pad * 4 && c->Ain.alloc(M * c->kin_pad * 4)) return 1;
    auto G = [&](const std::string& k, size_t n) -> int { return w->grads[k].alloc(n * 4); };
    if (G("start_w", (size_t)d * c->kin_pad) || G("start_b", d) || G("out_w", (size_t)256 * d) || G("out_b", 256) ||
        G("t_w1", 256 * 64) || G("t_b1", 256) || G("t_w2", d * 256) || G("t_b2", d)) return 1;
    for (int l = 0; l < c->NL; ++l) {
        const std::string p = "L" + std::to_string(l) + ".";
        if (G(p + "wqkv", nq * d) || G(p + "bqkv", nq) || G(p + "fc_w", d * hd) || G(p + "fc_b", d) || G(p + "ln1_g", d) || G(p + "ln1_b", d) ||
            G(p + "w1", d * d) || G(p + "b1", d) || G(p + "w2", d * d) || G(p + "b2", d) || G(p + "ln2_g", d) || G(p + "ln2_b", d)) return 1;
    }
    w->B = B;
    return 0;
}

// Many small device-to-device copies in ONE launch (parameter refresh before a training step: 72 tensors; gradients after it: 72
// tensors -- 144 copy nodes of 2-3 us each otherwise).  Entries are 2-D fp32 blocks (rows x cols with leading dimensions).
constexpr int MCOPY_MAX = 96;
struct MCopyEntry { float* dst; const float* src; int rows, cols, dld, sld; };
struct MCopyTable { MCopyEntry e[MCOPY_MAX]; };
static __global__ void multi_copy_kernel(const __grid_constant__ MCopyTable tab) {
    const MCopyEntry& e = tab.e[blockIdx.y];
    const long long n = (long long)e.rows * e.cols;
    const bool v4 = (e.cols & 3) == 0 && (e.dld & 3) == 0 && (e.sld & 3) == 0 &&
                    (((uintptr_t)e.dst | (uintptr_t)e.src) & 15) == 0;
    if (v4) {
        const int c4 = e.cols >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(i / c4), c = (int)(i % c4) * 4;
            *reinterpret_cast<float4*>(e.dst + (long long)r * e.dld + c) = *reinterpret_cast<const float4*>(e.src + (long long)r * e.sld + c);
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(i / e.cols), c = (int)(i % e.cols);
            e.dst[(long long)r * e.dld + c] = e.src[(long long)r * e.sld + c];
        }
    }
}
// While `g_copy_batch` is set (the batched entry points below), copy_block() collects instead of issuing a memcpy.
static thread_local std::vector<MCopyEntry>* g_copy_batch = nullptr;
static int copy_block(float* dst, int dld, const float* src, int sld, int rows, int cols, cudaStream_t s) {
    if (g_copy_batch) { g_copy_batch->push_back(MCopyEntry{dst, src, rows, cols, dld, sld}); return 0; }
    if (rows == 1 || (dld == cols && sld == cols)) { EG_CUDA(cudaMemcpyAsync(dst, src, (size_t)rows * cols * 4, cudaMemcpyDeviceToDevice, s)); }
    else { EG_CUDA(cudaMemcpy2DAsync(dst, (size_t)dld * 4, src, (size_t)sld * 4, (size_t)cols * 4, (size_t)rows, cudaMemcpyDeviceToDevice, s)); }
    return 0;
}
static int flush_copy_batch(std::vector<MCopyEntry>& v, cudaStream_t s) {
    for (size_t i0 = 0; i0 < v.size(); i0 += MCOPY_MAX) {
        MCopyTable tab;
        const int n = (int)std::min(v.size() - i0, (size_t)MCOPY_MAX);
        for (int i = 0; i < n; ++i) tab.e[i] = v[i0 + i];
        multi_copy_kernel<<<dim3(32, n), 256, 0, s>>>(tab);
        EG_CUDA(cudaGetLastError());
    }
    v.clear();
    return 0;
}

static __global__ void tr_set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }

// The launches of one training step on stream `s` (capturable: no allocation, no host synchronisation; every pointer is either the
// handle's own or one of the workspace's staging buffers).
static int train_step_body(egoego_ctx* c, TrainWs* w, const float* x_start, const float* cond_mask, const float* pmask, const int64_t* t_dev,
                           const float* noise, const float* cond_noise, const float* sqrt_ac, const float* sqrt_1mac, const float* weight,
                           int loss_l2, int B, int T, float* loss_out, DropCfg drop, cudaStream_t s) {
    const int M = B * LP, d = c->d, H = c->H, dk = c->dk, nq = 3 * H * dk, hd = H * dk, D = c->D, L = T + 1, KP = c->kin_pad;
    const long long nel = (long long)B * T * D;
    const float qs = 1.0f / sqrtf((float)dk);
    auto nblk = [](long long n) { return (unsigned)((n + 255) / 256); };

    // the batch's own timestep embeddings from the CURRENT time_mlp weights; the start epilogue indexes them by window (the
    // sampler's full table stays stale until the next egoego_commit_weights -- training handles never sample)
    tr_time_fwd_kernel<<<B, 256, 0, s>>>(reinterpret_cast<const long long*>(t_dev), c->t_w1.as<float>(), c->t_b1.as<float>(), c->t_w2.as<float>(),
                                         c->t_b2.as<float>(), w->temb_b.as<float>(), d);
    TSrc ts_b{w->iota.as<long long>(), nullptr, 0, B - 1};
    // ---------------- forward ----------------
    EG_CUDA(cudaMemsetAsync(c->Ain.p, 0, (size_t)M * KP * 4, s));
    tr_prep_kernel<<<nblk(nel), 256, 0, s>>>(x_start, cond_mask, noise, cond_noise, sqrt_ac, sqrt_1mac, c->Ain.as<float>(), KP, B, T, D);
    if (tr_gemm(w, c->Ain.as<float>(), KP, c->start_w.as<float>(), KP, M, d, KP,
                EpiStart{w->Hin[0].as<float>(), d, c->start_b.as<float>(), nullptr, c->pos.as<float>(), w->temb_b.as<float>(), ts_b, T}, s)) return 1;
    for (int l = 0; l < c->NL; ++l) {
        LayerW& W = c->layers[l];
        TrainLayerBufs& b = w->L[l];
        float* Hin = w->Hin[l].as<float>();
        if (tr_gemm(w, Hin, d, W.wqkv.as<float>(), d, M, nq, d, EpiBiasScale{b.QKV.as<float>(), nq, W.bqkv.as<float>(), hd, qs}, s)) return 1;
        attention_simt_kernel<false><<<B * H, 256, ATT_SIMT_SMEM, s>>>(b.QKV.as<float>(), nq, b.O.as<float>(), nullptr, nullptr, hd, H, L, drop, 4u * l + 0u);
        if (tr_gemm(w, b.O.as<float>(), hd, W.fc_w.as<float>(), hd, M, d, hd, EpiBiasDropResid{b.Y1.as<float>(), d, W.fc_b.as<float>(), Hin, drop, 4u * l + 1u}, s)) return 1;
        tr_ln_fwd_kernel<<<M / 8, 256, 0, s>>>(b.Y1.as<float>(), b.H1.as<float>(), b.st1.as<float>(), W.ln1_g.as<float>(), W.ln1_b.as<float>(), pmask, T, M);
        if (tr_gemm(w, b.H1.as<float>(), d, W.w1.as<float>(), d, M, d, d, EpiBiasRelu{b.F.as<float>(), d, W.b1.as<float>()}, s)) return 1;
        if (tr_gemm(w, b.F.as<float>(), d, W.w2.as<float>(), d, M, d, d, EpiBiasDropResid{b.Y2.as<float>(), d, W.b2.as<float>(), b.H1.as<float>(), drop, 4u * l + 2u}, s)) return 1;
        tr_ln_fwd_kernel<<<M / 8, 256, 0, s>>>(b.Y2.as<float>(), w->Hin[l + 1].as<float>(), b.st2.as<float>(), W.ln2_g.as<float>(), W.ln2_b.as<float>(), pmask, T, M);
    }
    if (tr_gemm(w, w->Hin[c->NL].as<float>(), d, c->out_w.as<float>(), d, M, D, d, EpiPlainBias{w->OUT.as<float>(), 256, c->out_b.as<float>()}, s)) return 1;
    EG_CUDA(cudaMemsetAsync(w->dOUT.p, 0, (size_t)M * 256 * 4, s));
    EG_CUDA(cudaMemsetAsync(w->loss.p, 0, 8, s));
    tr_loss_kernel<<<nblk(nel), 256, 0, s>>>(w->OUT.as<float>(), 256, c->cfg.objective == 0 ? noise : x_start, pmask, weight, loss_l2, B, T, D,
                                             w->loss.as<double>(), w->dOUT.as<float>());

    // ---------------- backward ----------------
    auto G = [&](const std::string& k) { return w->grads[k].as<float>(); };
    float *dH = w->dH.as<float>(), *dH1 = w->dH1.as<float>(), *dY = w->dY.as<float>(), *dF = w->dF.as<float>();
    // linear_out (:102,139): out = H[:, 1:] Wout^T + b
    if (tr_weight_grad(w, w->dOUT.as<float>(), 256, D, w->Hin[c->NL].as<float>(), d, d, M, G("out_w"), d, s)) return 1;
    tr_colsum(w->dOUT.as<float>(), M, D, 256, G("out_b"), s);
    if (tr_gemm_wt(w, w->dOUT.as<float>(), 256, c->out_w.as<float>(), d, M, d, D, dH, d, false, s)) return 1;
    for (int l = c->NL - 1; l >= 0; --l) {
        LayerW& W = c->layers[l];
        TrainLayerBufs& b = w->L[l];
        const std::string p = "L" + std::to_string(l) + ".";
        // ---- FFN (transformer_module.py:98-116): H2 = LN2(F W2^T + b2 + H1) * pm, F = relu(H1 W1^T + b1)
        tr_ln_bwd_kernel<<<M / 8, 256, 0, s>>>(dH, b.Y2.as<float>(), b.st2.as<float>(), W.ln2_g.as<float>(), pmask, T, M, dY, w->gp.as<float>(), w->bp.as<float>());
        tr_colsum(w->gp.as<float>(), M, d, d, G(p + "ln2_g"), s);
        tr_colsum(w->bp.as<float>(), M, d, d, G(p + "ln2_b"), s);
        // through dropout(F W2^T + b2): dZ = dY o mask (the residual branch below keeps dY)
        const float* dZ2 = dY;
        if (drop.on) { tr_dropout_bwd_kernel<<<nblk((long long)M * d / 4), 256, 0, s>>>(dY, w->dZ.as<float>(), (long long)M * d, drop, 4u * l + 2u); dZ2 = w->dZ.as<float>(); }
        tr_colsum(dZ2, M, d, d, G(p + "b2"), s);
        if (tr_weight_grad(w, dZ2, d, d, b.F.as<float>(), d, d, M, G(p + "w2"), d, s)) return 1;
        if (tr_gemm_wt(w, dZ2, d, W.w2.as<float>(), d, M, d, d, dF, d, false, s)) return 1;
        tr_relu_bwd_kernel<<<nblk((long long)M * d), 256, 0, s>>>(dF, b.F.as<float>(), (long long)M * d);
        tr_colsum(dF, M, d, d, G(p + "b1"), s);
        if (tr_weight_grad(w, dF, d, d, b.H1.as<float>(), d, d, M, G(p + "w1"), d, s)) return 1;
        EG_CUDA(cudaMemcpyAsync(dH1, dY, (size_t)M * d * 4, cudaMemcpyDeviceToDevice, s));      // residual branch
        if (tr_gemm_wt(w, dF, d, W.w1.as<float>(), d, M, d, d, dH1, d, true, s)) return 1;
        // ---- attention block (:61-95): H1 = LN1(O Wfc^T + bfc + Hin) * pm
        tr_ln_bwd_kernel<<<M / 8, 256, 0, s>>>(dH1, b.Y1.as<float>(), b.st1.as<float>(), W.ln1_g.as<float>(), pmask, T, M, dY, w->gp.as<float>(), w->bp.as<float>());
        tr_colsum(w->gp.as<float>(), M, d, d, G(p + "ln1_g"), s);
        tr_colsum(w->bp.as<float>(), M, d, d, G(p + "ln1_b"), s);
        const float* dZ1 = dY;                                                              // through dropout(O Wfc^T + bfc)
        if (drop.on) { tr_dropout_bwd_kernel<<<nblk((long long)M * d / 4), 256, 0, s>>>(dY, w->dZ.as<float>(), (long long)M * d, drop, 4u * l + 1u); dZ1 = w->dZ.as<float>(); }
        tr_colsum(dZ1, M, d, d, G(p + "fc_b"), s);
        if (tr_weight_grad(w, dZ1, d, d, b.O.as<float>(), hd, hd, M, G(p + "fc_w"), hd, s)) return 1;
        if (tr_gemm_wt(w, dZ1, d, W.fc_w.as<float>(), hd, M, hd, d, w->dO.as<float>(), hd, false, s)) return 1;
        attention_bwd_simt_kernel<<<B * H, 256, ATT_BWD_SMEM, s>>>(b.QKV.as<float>(), nq, w->dO.as<float>(), hd, w->dQKV.as<float>(), H, L, qs, drop, 4u * l + 0u);
        tr_colsum(w->dQKV.as<float>(), M, nq, nq, G(p + "bqkv"), s);
        if (tr_weight_grad(w, w->dQKV.as<float>(), nq, nq, w->Hin[l].as<float>(), d, d, M, G(p + "wqkv"), d, s)) return 1;
        EG_CUDA(cudaMemcpyAsync(dH, dY, (size_t)M * d * 4, cudaMemcpyDeviceToDevice, s));       // residual branch
        if (tr_gemm_wt(w, w->dQKV.as<float>(), nq, W.wqkv.as<float>(), d, M, d, nq, dH, d, true, s)) return 1;
    }
    // ---- time token (:105-116,122-123) and start_conv (transformer_module.py:203)
    for (const char* k : {"t_w1", "t_b1", "t_w2", "t_b2"}) EG_CUDA(cudaMemsetAsync(w->grads[k].p, 0, w->grads[k].bytes, s));
    tr_time_bwd_kernel<<<B, 256, 0, s>>>(dH, reinterpret_cast<const long long*>(t_dev), c->t_w1.as<float>(), c->t_b1.as<float>(), c->t_w2.as<float>(),
                                         G("t_w1"), G("t_b1"), G("t_w2"), G("t_b2"), d);
    tr_frame_rows_kernel<<<nblk((long long)M * d), 256, 0, s>>>(dH, dY, T, M, d);
    tr_colsum(dY, M, d, d, G("start_b"), s);
    if (tr_weight_grad(w, dY, d, d, c->Ain.as<float>(), KP, KP, M, G("start_w"), KP, s)) return 1;
    tr_loss_finish_kernel<<<1, 1, 0, s>>>(w->loss.as<double>(), loss_out);      // fp64 accumulator -> float, no host round trip
    EG_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace egoego

extern "C" {

int egoego_train_step(egoego_handle c, const float* x_start, const float* cond_mask, const float* pmask, const int64_t* t_dev,
                      const float* noise, const float* cond_noise, const float* sqrt_ac, const float* sqrt_1mac, const float* weight,
                      int loss_l2, int B, int T, float* loss_out, void* stream_v) {
    if (check_ready(c, B, T)) return 1;
    EG_CHECK(x_start && cond_mask && t_dev && noise && cond_noise && sqrt_ac && sqrt_1mac && weight && loss_out, "null argument");
    EG_CHECK(c->cfg.engine == EGOEGO_ENGINE_SIMT, "the training step runs on the fp32 engine: create the handle with EGOEGO_ENGINE_SIMT");
    EG_CHECK(c->d == 512 && c->dk == 256, "training step is specialised for d_model = 512, d_k = 256");
    EG_CHECK(B <= c->cfg.max_batch, "training batch exceeds cfg.max_batch");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    {
        static bool attr = false;
        if (!attr) { EG_CUDA(cudaFuncSetAttribute(attention_bwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_BWD_SMEM)); attr = true; }
    }
    cudaStream_t user = (cudaStream_t)stream_v;
    std::unique_ptr<TrainWs>& wp = g_train[c];
    if (!wp) wp.reset(new TrainWs());
    TrainWs* w = wp.get();
    if (train_alloc(c, w, B)) return 1;
    // dropout of the three sites per layer (common.cuh: DropCfg); p = 0 keeps the eval-mode semantics
    DropCfg drop{c->drop_seed, 0u, 1.0f, 0, nullptr};
    if (c->drop_p > 0.0) { drop.on = 1; drop.thresh = (uint32_t)((1.0 - c->drop_p) * 4294967296.0); drop.scale = (float)(1.0 / (1.0 - c->drop_p)); }
    c->launches += 40 * c->NL + 20;
    static const bool use_graph = []() { const char* e = getenv("EGOEGO_TRAIN_GRAPH"); return !(e && e[0] == '0'); }();
    if (!use_graph)
        return train_step_body(c, w, x_start, cond_mask, pmask, t_dev, noise, cond_noise, sqrt_ac, sqrt_1mac, weight, loss_l2, B, T, loss_out, drop, user);

    // ---- graph path: stage the inputs (caller's stream), then run / capture / replay the body on the handle's own stream ----
    const size_t xe = (size_t)B * T * c->D * 4;
    EG_CUDA(cudaMemcpyAsync(w->s_x.p, x_start, xe, cudaMemcpyDeviceToDevice, user));
    EG_CUDA(cudaMemcpyAsync(w->s_cm.p, cond_mask, xe, cudaMemcpyDeviceToDevice, user));
    EG_CUDA(cudaMemcpyAsync(w->s_noise.p, noise, xe, cudaMemcpyDeviceToDevice, user));
    EG_CUDA(cudaMemcpyAsync(w->s_cn.p, cond_noise, xe, cudaMemcpyDeviceToDevice, user));
    if (pmask) EG_CUDA(cudaMemcpyAsync(w->s_pm.p, pmask, (size_t)B * (T + 1) * 4, cudaMemcpyDeviceToDevice, user));
    EG_CUDA(cudaMemcpyAsync(w->s_t.p, t_dev, (size_t)B * 8, cudaMemcpyDeviceToDevice, user));
    EG_CUDA(cudaMemcpyAsync(w->s_sa.p, sqrt_ac, (size_t)B * 4, cudaMemcpyDeviceToDevice, user));
    EG_CUDA(cudaMemcpyAsync(w->s_sb.p, sqrt_1mac, (size_t)B * 4, cudaMemcpyDeviceToDevice, user));
    EG_CUDA(cudaMemcpyAsync(w->s_wt.p, weight, (size_t)B * 4, cudaMemcpyDeviceToDevice, user));
    tr_set_u64_kernel<<<1, 1, 0, user>>>(w->s_seed.as<unsigned long long>(), c->drop_seed);
    drop.seed_dev = w->s_seed.as<unsigned long long>();
    cudaStream_t s = c->own_stream;
    EG_CUDA(cudaEventRecord(c->ev_in, user));
    EG_CUDA(cudaStreamWaitEvent(s, c->ev_in, 0));
    auto body = [&](cudaStream_t st) -> int {
        return train_step_body(c, w, w->s_x.as<float>(), w->s_cm.as<float>(), pmask ? w->s_pm.as<float>() : nullptr, w->s_t.as<int64_t>(),
                               w->s_noise.as<float>(), w->s_cn.as<float>(), w->s_sa.as<float>(), w->s_sb.as<float>(), w->s_wt.as<float>(),
                               loss_l2, B, T, w->s_loss.as<float>(), drop, st);
    };
    const TrainWs::GraphKey key{B, T, loss_l2, pmask ? 1 : 0, drop.on, drop.thresh};
    auto it = w->graphs.find(key);
    if (it == w->graphs.end() || w->no_graph.count(key)) {   // first step of this shape (plane caches, function attributes, scratch
        if (body(s)) return 1;                               // buffers), or a shape whose capture failed once: eager launches
        if (it == w->graphs.end()) w->graphs[key] = nullptr;
    } else {
        if (!it->second) {                         // second step: capture; if anything in it cannot be captured, stay eager for this shape
            cudaGraph_t g = nullptr;
            cudaGraphExec_t ex = nullptr;
            const cudaError_t cb = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
            const int rc = cb == cudaSuccess ? body(s) : 1;
            const cudaError_t ce = cb == cudaSuccess ? cudaStreamEndCapture(s, &g) : cb;
            if (rc == 0 && ce == cudaSuccess && g && cudaGraphInstantiate(&ex, g, 0) == cudaSuccess) it->second = ex;
            else { cudaGetLastError(); w->no_graph.insert(key); }
            if (g) cudaGraphDestroy(g);
        }
        if (it->second) { EG_CUDA(cudaGraphLaunch(it->second, s)); }
        else if (body(s)) return 1;
    }
    EG_CUDA(cudaEventRecord(c->ev_out, s));
    EG_CUDA(cudaStreamWaitEvent(user, c->ev_out, 0));
    EG_CUDA(cudaMemcpyAsync(loss_out, w->s_loss.p, 4, cudaMemcpyDeviceToDevice, user));
    return 0;
}

// Dropout of the following egoego_train_step calls: p = 0 -> identity (the reference's eval() mode), p = 0.1 -> nn.Dropout(0.1) at
// the three sites of every DecoderLayer with Philox masks keyed by `seed` (common.cuh: DropCfg; oracle/training.py restates them).
int egoego_train_set_dropout(egoego_handle c, double p, uint64_t seed) {
    EG_CHECK(c, "null handle");
    EG_CHECK(p >= 0.0 && p < 1.0, "dropout probability must be in [0, 1)");
    c->drop_p = p; c->drop_seed = seed;
    return 0;
}

// In-place device-to-device refresh of ONE parameter tensor of a committed fp32 (SIMT-engine) handle -- what an optimizer step
// needs between two training steps (a full egoego_set_tensor + egoego_commit_weights round trip goes through host staging and
// re-allocates the workspace).  src_dev: fp32 device pointer in the reference's layout.
int egoego_update_tensor_device(egoego_handle c, const char* name_in, const float* src, int64_t numel, void* stream_v) {
    EG_CHECK(c && name_in && src, "null argument");
    EG_CHECK(c->committed && c->cfg.engine == EGOEGO_ENGINE_SIMT, "needs a committed EGOEGO_ENGINE_SIMT handle");
    EG_CUDA(cudaSetDevice(c->cfg.device));
    cudaStream_t s = (cudaStream_t)stream_v;
    std::string name(name_in);
    for (const char* pre : {"ema_model.", "model.", "module."})
        if (name.rfind(pre, 0) == 0) name = name.substr(strlen(pre));
    const int d = c->d, hd = c->H * c->dk, D = c->D;
    float* dst = nullptr; int64_t expect = 0;
    const std::string pre = "denoise_fn.motion_transformer.";
    if (name == pre + "start_conv.weight") {
        EG_CHECK(numel == (int64_t)d * 2 * D, "tensor '" + name + "': wrong size");
        return copy_block(c->start_w.as<float>(), c->kin_pad, src, 2 * D, d, 2 * D, s);
    }
    if (name == pre + "start_conv.bias") { dst = c->start_b.as<float>(); expect = d; }
    else if (name == pre + "position_vec.weight") { dst = c->pos.as<float>(); expect = (int64_t)(c->cfg.max_timesteps + 1) * d; }
    else if (name == "denoise_fn.linear_out.weight") { dst = c->out_w.as<float>(); expect = (int64_t)D * d; }
    else if (name == "denoise_fn.linear_out.bias") { dst = c->out_b.as<float>(); expect = D; }
    else if (name == "denoise_fn.time_mlp.1.weight") { dst = c->t_w1.as<float>(); expect = 256 * 64; c->temb_dirty = true; }
    else if (name == "denoise_fn.time_mlp.1.bias") { dst = c->t_b1.as<float>(); expect = 256; c->temb_dirty = true; }
    else if (name == "denoise_fn.time_mlp.3.weight") { dst = c->t_w2.as<float>(); expect = (int64_t)d * 256; c->temb_dirty = true; }
    else if (name == "denoise_fn.time_mlp.3.bias") { dst = c->t_b2.as<float>(); expect = d; c->temb_dirty = true; }
    else if (name.rfind(pre + "layer_stack.", 0) == 0) {
        const std::string rest = name.substr((pre + "layer_stack.").size());
        const size_t dot = rest.find('.');
        EG_CHECK(dot != std::string::npos, "unknown tensor: " + name);
        const int l = atoi(rest.substr(0, dot).c_str());
        EG_CHECK(l >= 0 && l < c->NL, "layer index out of range: " + name);
        LayerW& L = c->layers[l];
        const std::string k = rest.substr(dot + 1);
        const char* secs[3] = {"w_q", "w_k", "w_v"};
        for (int sct = 0; sct < 3; ++sct) {
            if (k == std::string("self_attn.") + secs[sct] + ".weight") { dst = L.wqkv.as<float>() + (size_t)sct * hd * d; expect = (int64_t)hd * d; }
            if (k == std::string("self_attn.") + secs[sct] + ".bias") { dst = L.bqkv.as<float>() + (size_t)sct * hd; expect = hd; }
        }
        if (k == "self_attn.fc.weight") { dst = L.fc_w.as<float>(); expect = (int64_t)d * hd; }
        else if (k == "self_attn.fc.bias") { dst = L.fc_b.as<float>(); expect = d; }
        else if (k == "self_attn.layer_norm.weight") { dst = L.ln1_g.as<float>(); expect = d; }
        else if (k == "self_attn.layer_norm.bias") { dst = L.ln1_b.as<float>(); expect = d; }
        else if (k == "pos_ffn.w_1.weight") { dst = L.w1.as<float>(); expect = (int64_t)d * d; }
        else if (k == "pos_ffn.w_1.bias") { dst = L.b1.as<float>(); expect = d; }
        else if (k == "pos_ffn.w_2.weight") { dst = L.w2.as<float>(); expect = (int64_t)d * d; }
        else if (k == "pos_ffn.w_2.bias") { dst = L.b2.as<float>(); expect = d; }
        else if (k == "pos_ffn.layer_norm.weight") { dst = L.ln2_g.as<float>(); expect = d; }
        else if (k == "pos_ffn.layer_norm.bias") { dst = L.ln2_b.as<float>(); expect = d; }
    }
    if (!dst) return 0;                                  // schedule buffers etc.: not used by the training step
    EG_CHECK(expect == numel, "tensor '" + name + "': expected " + std::to_string(expect) + " elements, got " + std::to_string(numel));
    EG_CHECK(numel < (1ll << 31), "tensor too large");
    return copy_block(dst, (int)numel, src, (int)numel, 1, (int)numel, s);
}

// Gradient of the tensor `name` (reference state_dict key, e.g. "denoise_fn.motion_transformer.layer_stack.2.self_attn.w_k.weight")
// from the last egoego_train_step, copied to dst_dev[numel] in the reference's layout.
int egoego_train_get_grad(egoego_handle c, const char* name_in, float* dst, int64_t numel, void* stream_v) {
    EG_CHECK(c && name_in && dst, "null argument");
    auto it = g_train.find(c);
    EG_CHECK(it != g_train.end() && it->second && it->second->B > 0, "egoego_train_step has not been called on this handle");
    TrainWs* w = it->second.get();
    cudaStream_t s = (cudaStream_t)stream_v;
    std::string name(name_in);
    for (const char* pre : {"ema_model.", "model.", "module."})
        if (name.rfind(pre, 0) == 0) name = name.substr(strlen(pre));
    const int d = c->d, hd = c->H * c->dk, D = c->D;
    const float* src = nullptr; int64_t rows = 1, cols = 0, ld = 0;
    auto G = [&](const std::string& k) { return w->grads[k].as<float>(); };
    const std::string pre = "denoise_fn.motion_transformer.";
    if (name == pre + "start_conv.weight") { src = G("start_w"); rows = d; cols = 2 * D; ld = c->kin_pad; }
    else if (name == pre + "start_conv.bias") { src = G("start_b"); cols = d; }
    else if (name == "denoise_fn.linear_out.weight") { src = G("out_w"); rows = D; cols = d; ld = d; }
    else if (name == "denoise_fn.linear_out.bias") { src = G("out_b"); cols = D; }
    else if (name == "denoise_fn.time_mlp.1.weight") { src = G("t_w1"); cols = 256 * 64; }
    else if (name == "denoise_fn.time_mlp.1.bias") { src = G("t_b1"); cols = 256; }
    else if (name == "denoise_fn.time_mlp.3.weight") { src = G("t_w2"); cols = (int64_t)d * 256; }
    else if (name == "denoise_fn.time_mlp.3.bias") { src = G("t_b2"); cols = d; }
    else if (name.rfind(pre + "layer_stack.", 0) == 0) {
        const std::string rest = name.substr((pre + "layer_stack.").size());
        const size_t dot = rest.find('.');
        EG_CHECK(dot != std::string::npos, "unknown tensor: " + name);
        const int l = atoi(rest.substr(0, dot).c_str());
        EG_CHECK(l >= 0 && l < c->NL, "layer index out of range: " + name);
        const std::string k = rest.substr(dot + 1), p = "L" + std::to_string(l) + ".";
        const char* secs[3] = {"w_q", "w_k", "w_v"};
        for (int sct = 0; sct < 3; ++sct) {
            if (k == std::string("self_attn.") + secs[sct] + ".weight") { src = G(p + "wqkv") + (size_t)sct * hd * d; cols = (int64_t)hd * d; }
            if (k == std::string("self_attn.") + secs[sct] + ".bias") { src = G(p + "bqkv") + (size_t)sct * hd; cols = hd; }
        }
        if (k == "self_attn.fc.weight") { src = G(p + "fc_w"); cols = (int64_t)d * hd; }
        else if (k == "self_attn.fc.bias") { src = G(p + "fc_b"); cols = d; }
        else if (k == "self_attn.layer_norm.weight") { src = G(p + "ln1_g"); cols = d; }
        else if (k == "self_attn.layer_norm.bias") { src = G(p + "ln1_b"); cols = d; }
        else if (k == "pos_ffn.w_1.weight") { src = G(p + "w1"); cols = (int64_t)d * d; }
        else if (k == "pos_ffn.w_1.bias") { src = G(p + "b1"); cols = d; }
        else if (k == "pos_ffn.w_2.weight") { src = G(p + "w2"); cols = (int64_t)d * d; }
        else if (k == "pos_ffn.w_2.bias") { src = G(p + "b2"); cols = d; }
        else if (k == "pos_ffn.layer_norm.weight") { src = G(p + "ln2_g"); cols = d; }
        else if (k == "pos_ffn.layer_norm.bias") { src = G(p + "ln2_b"); cols = d; }
    }
    EG_CHECK(src != nullptr, "no gradient for tensor: " + name);
    EG_CHECK(rows * cols == numel, "tensor '" + name + "': expected " + std::to_string(rows * cols) + " elements, got " + std::to_string(numel));
    EG_CHECK(cols < (1ll << 31) && rows < (1ll << 31), "tensor too large");
    return copy_block(dst, (int)cols, src, rows == 1 ? (int)cols : (int)ld, (int)rows, (int)cols, s);
}

// batched forms (one FFI call per step instead of one per tensor)
int egoego_train_get_grads(egoego_handle c, int n, const char* const* names, float* const* dsts, const int64_t* numels, void* stream_v) {
    EG_CHECK(n >= 0 && (n == 0 || (names && dsts && numels)), "null argument");
    std::vector<MCopyEntry> batch;
    g_copy_batch = &batch;
    int rc = 0;
    for (int i = 0; i < n && !rc; ++i) rc = egoego_train_get_grad(c, names[i], dsts[i], numels[i], stream_v);
    g_copy_batch = nullptr;
    return rc ? 1 : flush_copy_batch(batch, (cudaStream_t)stream_v);
}
int egoego_update_tensors_device(egoego_handle c, int n, const char* const* names, const float* const* srcs, const int64_t* numels, void* stream_v) {
    EG_CHECK(n >= 0 && (n == 0 || (names && srcs && numels)), "null argument");
    std::vector<MCopyEntry> batch;
    g_copy_batch = &batch;
    int rc = 0;
    for (int i = 0; i < n && !rc; ++i) rc = egoego_update_tensor_device(c, names[i], srcs[i], numels[i], stream_v);
    g_copy_batch = nullptr;
    return rc ? 1 : flush_copy_batch(batch, (cudaStream_t)stream_v);
}

}  // extern "C"
