// Tensor-core engine (EGOEGO_ENGINE_TCGEN05): interface used by egoego_b200.cu.
// Implementation: engine_tc.cu (tcgen05 + TMA GEMMs with a 3-term fp16 hi/lo split).
#pragma once
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace egoego {

struct TcLayerW {
    const float *wq, *wk, *wv, *fc, *w1, *w2;                                  // host fp32 (reference layouts)
    const float *bqkv, *fc_b, *b1, *b2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;        // device fp32
};

struct TcWeights {
    int D, d, H, dk, NL, N, Tmax, max_batch;
    const float* start_w;   // host [d, 2D]
    const float* start_b;   // device [d]
    const float* pos;       // device [max_timesteps+1, d]
    const float* temb;      // device [N, d]
    const float* out_w;     // host [D, d]
    const float* out_b;     // device [D]
    std::vector<TcLayerW> layers;
};

struct TcImpl;

class TcEngine {
public:
    TcEngine();
    ~TcEngine();
    int init(const TcWeights& w, cudaStream_t s);
    // zero the padded A-operand planes for B windows
    int clear_staging(int B, cudaStream_t s);
    // scatter compact rows into the fp16 hi/lo A-operand planes (x half or x_cond half)
    int stage(const float* src, int src_ld, int src_col0, bool cond_half, int B, int T, cudaStream_t s, int64_t* n);
    // base[M, d] = x_cond-half of start_conv + bias + positional rows (constant over the loop)
    int prepare_cond(int B, int T, cudaStream_t s, int64_t* n);
    // one denoiser call; expects the x half staged; writes model_out[B, T, D]
    // `fuse` != nullptr: the DDPM update of the sampling loop runs in the epilogue of linear_out (x_out / next-step operand
    // planes are written instead of model_out; no ddpm_update_kernel launch is needed afterwards)
    int denoiser(int B, int T, TSrc ts, const float* pmask, float* model_out, cudaStream_t s, int64_t* n, int fmt /* FMT_SPLIT=0 | FMT_HALF=1 */,
                 const DdpmArgs* fuse = nullptr);
    void stage_targets(__nv_bfloat16** hi, __nv_bfloat16** lo, __half** h16, int* ld);
    int launches_per_denoiser(int fmt = 0) const;
    // dithered fp16 weight sets of the single-pass format: the sampling loop uses set (step index mod n) for step i;
    // r < 0 selects the plain round-to-nearest copy (the default outside the loop)
    int n_weight_sets() const;
    void use_weight_set(int r);
    // 3-term split launches: hi*hi and the cross terms in separate TMEM accumulators (true, the default) or in one (false)
    void set_dual_acc(bool on);
    // FMT_HALF launches: weights as an fp16 hi/lo pair in two passes over K (fp16 activations, exact weights) instead of one fp16 copy
    void set_weight_pair(bool on);
    // zero the per-window completion counters of the streamed attention; call wherever the loop's device step counter is reset
    int reset_stream_counters(cudaStream_t s);
    int time_stage(int B, int T, int stage, int fmt, int iters, float* model_out, cudaStream_t s, float* ms, const DdpmArgs* fuse = nullptr);
    // "key=value ..." description of the resolved kernel choices (cluster counts, zig-zag, fused LN)
    std::string info() const;
private:
    TcImpl* impl_;
};

// debug: per-CTA start / end timestamps of the streamed QKV projection and attention kernels (enable: 1 / 0, -1 = leave; out: nullable)
int tc_debug_timeline(int enable, unsigned long long* out, int n);

// C[M, N] (= or +=) A[M, K] W[N, K]^T, fp32 row-major device operands, on the tensor cores (3-term bf16 split, fp32 accumulate).
// C must have ceil(M / 256) * 256 rows of ldc floats; columns [0, n_valid) are written (n_valid % 4 == 0 <= ldc).
// tc_gemm_f32_ex: trans_a / trans_w = 1 means that operand is stored with the reduction index as the ROW ([K, M] / [K, N], leading
// dimension lda / ldw): C = A^T W for the weight gradients, C = A W' for a weight stored [out, in] -- no separate transposes.
int tc_gemm_f32_ex(const float* A, int lda, int trans_a, const float* W, int ldw, int trans_w, int M, int N, int K, float* C, int ldc, int n_valid,
                   int accumulate, cudaStream_t s);
// C = A W^T with the element-wise functor `e` (kernels_simt.cuh / train.cuh: operator()(row, col, acc), rows < M, cols < N) applied in
// the GEMM epilogue; instantiated in engine_tc.cu for the functors of the training step's forward pass.
template <class E>
int tc_gemm_f32_epi(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const E& e, cudaStream_t s);
int tc_gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, float* C, int ldc, int n_valid, int accumulate,
                cudaStream_t s);

// TMA map of a row-major 16-bit [rows, cols] tensor with [box_rows, 64]-element boxes and the 128-byte swizzle (engine_tc.cu)
int tc_make_map16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

// Host-side rounding of fp32 weights to the r-th of R dithered fp16 copies (bit patterns), exactly as the weight commit does.
void dither_weights_host(const float* w, long long n, int r, int R, unsigned short* out);

// tcgen05 split GEMM vs fp32 SIMT GEMM on random data (C = A W^T, no epilogue); returns max |err|, max |ref|, ms/launch.
int selftest_gemm(int M, int N, int K, unsigned long long seed, int two_cta, int half_fmt, float* max_abs_err, float* max_abs_ref, float* ms);

}  // namespace egoego
