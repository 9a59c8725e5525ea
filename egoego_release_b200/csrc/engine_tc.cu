// Tensor-core engine: tcgen05/TMA 3-term split GEMMs (fp16 hi/lo planes) for every projection of the denoiser, warp-level
// LayerNorm, attention; see gemm_tcgen05.cuh for the GEMM kernel.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <memory>

#include "engine_tc.cuh"
#include "gemm_tcgen05.cuh"
#include "attn_tcgen05.cuh"
#include "kernels_simt.cuh"

namespace egoego {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
    return fn;
}

// 2D 16-bit (fp16 / bf16: the copy engine does not care) row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128-byte swizzle.
static int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode();
    EG_CHECK(enc, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return 0;
}

int tc_make_map16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) { return make_map(m, base, rows, cols, box_rows); }

struct Plane {                 // a 16-bit hi/lo pair (fp16 in the sampler, bf16 in tc_gemm_f32) with its TMA maps
    __nv_bfloat16 *hi = nullptr, *lo = nullptr;
    CUtensorMap mhi, mlo;
    CUtensorMap mhi128, mlo128;     // 128-row boxes (per-CTA half of a W tile in the 2-CTA GEMM)
    __nv_bfloat16* h16 = nullptr;   // fp16 plane for FMT_HALF launches: a separate buffer (weights, X) or an alias of `hi`
    CUtensorMap m16, m16_128;
    CUtensorMap mhi64, m16_64;      // 64-row boxes: the multicast halves of the cluster-of-8 GEMM (gemm_half_tma_2cta_kernel<.., true>)
    bool own16 = false;
    size_t rows = 0, cols = 0;
    int alloc(size_t r, size_t c, uint32_t box_rows) {
        rows = r; cols = c;
        EG_CUDA(cudaMalloc(&hi, r * c * 2));
        EG_CUDA(cudaMalloc(&lo, r * c * 2));
        EG_CUDA(cudaMemset(hi, 0, r * c * 2));
        EG_CUDA(cudaMemset(lo, 0, r * c * 2));
        if (make_map(&mhi, hi, r, c, box_rows) || make_map(&mlo, lo, r, c, box_rows)) return 1;
        if (make_map(&mhi128, hi, r, c, 128) || make_map(&mlo128, lo, r, c, 128) || make_map(&mhi64, hi, r, c, 64)) return 1;
        h16 = hi; m16 = mhi; m16_128 = mhi128; m16_64 = mhi64; own16 = false;       // activations: the fp16 plane reuses the hi buffer
        box = box_rows;
        return 0;
    }
    int alloc16() {                                                  // separate fp16 buffer (weights, sampler input)
        EG_CUDA(cudaMalloc(&h16, rows * cols * 2));
        EG_CUDA(cudaMemset(h16, 0, rows * cols * 2));
        own16 = true;
        if (make_map(&m16, h16, rows, cols, box) || make_map(&m16_128, h16, rows, cols, 128) || make_map(&m16_64, h16, rows, cols, 64)) return 1;
        return 0;
    }
    uint32_t box = 128;
    // Weights only: R fp16 planes, each rounded with its own dither offset (see upload_weight); use_set(r) makes set r the
    // plane FMT_HALF launches read (the maps are passed to kernels by value, so this is a host-side switch).
    std::vector<__nv_bfloat16*> set16;
    std::vector<CUtensorMap> set_m16, set_m16_128, set_m16_64;
    void use_set(int r) {                              // r < 0: the plain round-to-nearest copy (= the hi plane of the pair)
        if (set16.empty()) return;
        if (r < 0) { h16 = hi; m16 = mhi; m16_128 = mhi128; m16_64 = mhi64; return; }
        r %= (int)set16.size();
        h16 = set16[r]; m16 = set_m16[r]; m16_128 = set_m16_128[r]; m16_64 = set_m16_64[r];
    }
    void release() {
        if (hi) cudaFree(hi); if (lo) cudaFree(lo); if (own16 && h16) cudaFree(h16);
        for (auto* p : set16) cudaFree(p);
        set16.clear(); set_m16.clear(); set_m16_128.clear(); set_m16_64.clear();
        hi = lo = h16 = nullptr;
    }
};

// Number of dithered fp16 weight sets of the single-pass format (EGOEGO_WEIGHT_SETS, default 8; 1 = plain round-to-nearest).
// Why: in the fp16 steps the ACTIVATION rounding is harmless (fresh, zero-mean every step) but the WEIGHT rounding is the
// same perturbation of the network at every step, and the sampler integrates it coherently over hundreds of steps
// (tools/rounding_study.py: weights-only rounding reproduces the whole fp16-step error, activations-only sits on the fp32
// floor).  Set r rounds W + u_r ulp16(W) to nearest with u_r = (bitrev(r) + 1/2) / R - 1/2; step i of the loop uses set
// i mod R, so over any R consecutive steps the mean rounded weight is within ulp16 / (2R) of W instead of ulp16 / 2.
static int weight_sets() {                          // read at every weight commit (not cached: tools compare settings in one process)
    const char* e = getenv("EGOEGO_WEIGHT_SETS");
    int v = e ? atoi(e) : 8;
    if (v < 1) v = 1;
    if (v > 16) v = 16;
    return v;
}
static float dither_offset(int r, int R) {
    if (R <= 1) return 0.f;
    int rev = r;
    if ((R & (R - 1)) == 0) {                       // power of two: bit-reversed order spreads consecutive steps
        rev = 0;
        for (int b = 1, x = r; b < R; b <<= 1, x >>= 1) rev = (rev << 1) | (x & 1);
    }
    return ((float)rev + 0.5f) / (float)R - 0.5f;
}
static __half dither_round(float v, float u) {
    if (u == 0.f || v == 0.f) return __float2half_rn(v);
    int e = 0;
    (void)frexpf(v, &e);                            // |v| = m 2^e, m in [0.5, 1): fp16 spacing 2^(e-11), 2^-24 once subnormal
    const int q = e - 11 < -24 ? -24 : e - 11;
    return __float2half_rn(v + u * ldexpf(1.0f, q));
}

void dither_weights_host(const float* w, long long n, int r, int R, unsigned short* out) {
    const float u = dither_offset(r, R);
    for (long long i = 0; i < n; ++i) {
        const __half h = dither_round(w[i], u);
        memcpy(&out[i], &h, 2);
    }
}

// Host fp32 [rows, src_ld] (columns [col0, col0+ncols)) -> zero-padded [rows_pad, cols_pad] hi/lo planes.
static int upload_weight(Plane& p, const float* w, int rows, int src_ld, int col0, int ncols, int rows_pad, int cols_pad, uint32_t box_rows) {
    if (p.alloc(rows_pad, cols_pad, box_rows)) return 1;
    std::vector<__half> hi((size_t)rows_pad * cols_pad, __float2half(0.f)), lo(hi);      // fp16 hi/lo pair (split_f16)
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < ncols; ++c) {
            float v = w[(size_t)r * src_ld + col0 + c];
            __half h = __float2half_rn(v);
            hi[(size_t)r * cols_pad + c] = h;
            lo[(size_t)r * cols_pad + c] = __float2half_rn(v - __half2float(h));
        }
    EG_CUDA(cudaMemcpy(p.hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice));
    EG_CUDA(cudaMemcpy(p.lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice));
    const int R = weight_sets();
    std::vector<__half> h16((size_t)rows_pad * cols_pad, __float2half(0.f));
    p.set16.assign(R, nullptr); p.set_m16.resize(R); p.set_m16_128.resize(R); p.set_m16_64.resize(R);
    for (int k = 0; k < R; ++k) {
        const float u = dither_offset(k, R);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < ncols; ++c) h16[(size_t)r * cols_pad + c] = dither_round(w[(size_t)r * src_ld + col0 + c], u);
        EG_CUDA(cudaMalloc(&p.set16[k], h16.size() * 2));
        EG_CUDA(cudaMemcpy(p.set16[k], h16.data(), h16.size() * 2, cudaMemcpyHostToDevice));
        if (make_map(&p.set_m16[k], p.set16[k], rows_pad, cols_pad, box_rows) || make_map(&p.set_m16_128[k], p.set16[k], rows_pad, cols_pad, 128) ||
            make_map(&p.set_m16_64[k], p.set16[k], rows_pad, cols_pad, 64)) return 1;
    }
    p.own16 = false;
    p.use_set(-1);                                       // outside the sampling loop FMT_HALF launches read the plain RN copy
    return 0;
}

struct TcLayer {
    Plane wqkv, fc, w1, w2;
    const float *bqkv, *fc_b, *b1, *b2, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

struct TcImpl {
    TcWeights w;
    int M = 0, kx = 0, nout = 0, sms = 148;
    Plane Wx, Wc, Wout;
    std::vector<TcLayer> layers;
    Plane X, C, Hs, O, F;                       // activation planes (A operands)
    Plane Qp, Kp, VT;                           // attention operand planes (see attn_tcgen05.cuh); VT = V in [key][dim] layout, 64-row TMA boxes
    bool fuse_ln = true;                        // EGOEGO_FUSE_LN=0 keeps GEMM + LayerNorm separate in the fp16 format too
    bool attn_tc = true;                        // EGOEGO_ATTN=simt selects the fp32 CUDA-core attention (bisecting)
    bool attn_v2 = true;                        // fp16 steps: software-pipelined attention_half_kernel (EGOEGO_ATTN=v1: attention_tc_kernel<FMT_HALF>)
    int ln4_clusters = 0;                       // co-resident clusters of 4 for gemm_ln_half_c4_kernel (0 = unfused GEMM + LayerNorm)
    // L2 zig-zag (EGOEGO_ZIGZAG, default on): consecutive kernels of a step walk the windows in OPPOSITE directions, so a
    // kernel starts with the rows its producer wrote last -- the part of its input that is still in the 126 MB L2 (the
    // per-kernel working set at 256 windows is 100-270 MB, so a same-direction walk misses everywhere).
    bool dual_acc = true;                       // sampler engines: dual-accumulator split GEMMs (the selftest / training instances set it per call)
    int c8_clusters = 0;                        // co-resident clusters of 8 for the multicast GEMM (EGOEGO_GEMM_C8=1; 0 = off)
    bool zigzag = true;
    // FMT_HALF launches read the weights as an fp16 PAIR in two passes over K (A W_lo^T first, then A W_hi^T): the format of the steps
    // between the single-pass and the 3-term ones (set per captured step by the sampler, like dual_acc)
    bool w_pair = false;
    // Streamed attention (EGOEGO_STREAM_ATT=1; off by default, see DESIGN.md 5.2: measured no faster): in the sampling loop the fp16-format QKV projection runs on `stream_pairs`
    // CTA pairs and attention_half_kernel CONCURRENTLY on the remaining SMs, following the projection window by window through
    // per-window completion counters (one row of `att_cnt` per layer; a window is complete at 12 tiles x (steps since reset)).
    // Q / K / V (201 MB per layer at 256 windows, more than the L2 holds) are then read back while they are still in the L2.
    bool stream_att = false;
    int stream_pairs = 62;
    double stream_ratio = 0.96;                 // attention item time / projection tile time (EGOEGO_STREAM_ATT_RATIO; measured 4.1 / 4.28 us)
    int* att_cnt = nullptr;                     // [NL][cnt_ld]
    int cnt_ld = 0;
    int dir = 0;                                // direction of the next kernel launched (0 = ascending windows)
    int next_dir() { const int d = zigzag ? dir : 0; dir ^= 1; return d; }
    float *base = nullptr, *H = nullptr, *Y = nullptr, *QKV = nullptr;
    ~TcImpl() {
        for (Plane* p : {&Wx, &Wc, &Wout, &X, &C, &Hs, &O, &F, &Qp, &Kp, &VT}) p->release();
        for (auto& l : layers) for (Plane* p : {&l.wqkv, &l.fc, &l.w1, &l.w2}) p->release();
        for (float* p : {base, H, Y, QKV}) if (p) cudaFree(p);
        if (att_cnt) cudaFree(att_cnt);
    }
};

// First use of a kernel on each device of the process (one process per GPU is the deployment, but egoego_cfg.device allows more).
struct PerDeviceOnce {
    bool done[64] = {};
    bool need() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess) return true;
        d &= 63;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

static inline int Mr(int B) { return ((B + 1) / 2) * 2 * LP; }

template <int BN, int FMT, class Epi>
static int launch_gemm(TcImpl* I, const Plane& A, const Plane& W, int M, int N, int K, const Epi& epi, cudaStream_t s) {
    using Cfg = GemmCfg<BN, FMT>;
    static PerDeviceOnce attr_once;          // opt-in shared-memory size is a per-device function attribute
    auto kern = gemm_split3_kernel<BN, FMT, Epi>;
    if (attr_once.need()) {
        EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    }
    EG_CHECK(M % GEMM_BM == 0 && N % BN == 0 && K % GEMM_BK == 0, "gemm shape not tile-aligned");
    const int tiles = (M / GEMM_BM) * (N / BN);
    const int grid = tiles < I->sms ? tiles : I->sms;
    LaunchCfg lc(grid, GEMM_THREADS, Cfg::SMEM_BYTES, s);
    if (fmt_is_split(FMT)) { EG_CUDA(cudaLaunchKernelEx(&lc.cfg, kern, A.mhi, A.mlo, W.mhi, W.mlo, M, N, K, epi)); }
    else                  { EG_CUDA(cudaLaunchKernelEx(&lc.cfg, kern, A.m16, A.m16, W.m16, W.m16, M, N, K, epi)); }
    return 0;
}

static bool use_2cta() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("EGOEGO_GEMM"); v = (e && strcmp(e, "1cta") == 0) ? 0 : 1; }   // default: CTA-pair tiles (10-12 % faster, profiles/)
    return v == 1;
}

// dual-accumulator split GEMM (hi*hi and the cross terms in separate TMEM accumulators, see gemm_split3_2cta_kernel): default for the
// 3-term steps of the sampler; EGOEGO_SPLIT_DUAL=0 restores the single accumulator
static bool use_dual_acc() {
    const char* e = getenv("EGOEGO_SPLIT_DUAL");          // read per launch: tools compare both settings in one process
    return !(e && e[0] == '0');
}
template <int FMT, class Epi>
static int launch_gemm_2cta_dual(TcImpl* I, const Plane& A, const Plane& W, int M, int N, int K, const Epi& epi, cudaStream_t s) {
    static PerDeviceOnce attr_once;
    auto kern = gemm_split3_2cta_kernel<FMT, Epi, true>;
    constexpr int GEMM2_SMEM_BYTES = Gemm2Cfg<FMT>::SMEM_BYTES;
    if (attr_once.need()) {
        EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM_BYTES));
    }
    const int tiles = (M / 256) * (N / 256);
    int pairs = I->sms / 2;
    if (tiles < pairs) pairs = tiles;
    LaunchCfg lc(2 * pairs, GEMM_THREADS, GEMM2_SMEM_BYTES, s, 2);
    EG_CUDA(cudaLaunchKernelEx(&lc.cfg, kern, A.mhi, A.mlo, W.mhi128, W.mlo128, M, N, K, epi, I->next_dir(), 1, 1));
    return 0;
}

// 2-CTA (cluster of 2, cta_group::2) launch: 256 x 256 tiles per CTA pair.
template <int FMT, class Epi>
static int launch_gemm_2cta(TcImpl* I, const Plane& A, const Plane& W, int M, int N, int K, const Epi& epi, cudaStream_t s, int ksplit = 1) {
    EG_CHECK(M % 256 == 0 && N % 256 == 0 && K % GEMM_BK == 0, "2-CTA gemm shape not tile-aligned");
    EG_CHECK(ksplit >= 1 && (K / GEMM_BK) % ksplit == 0 && (ksplit == 1 || fmt_is_split(FMT)), "bad split-K factor");
    if constexpr (FMT == FMT_SPLIT) {
        if (I->dual_acc && use_dual_acc()) return launch_gemm_2cta_dual<FMT>(I, A, W, M, N, K, epi, s);
    }
    static PerDeviceOnce attr_once;          // opt-in shared-memory size is a per-device function attribute
    auto kern = gemm_split3_2cta_kernel<FMT, Epi>;
    constexpr int GEMM2_SMEM_BYTES = Gemm2Cfg<FMT>::SMEM_BYTES;
    if (attr_once.need()) {
        EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM2_SMEM_BYTES));
    }
    EG_CHECK(M % 256 == 0 && N % 256 == 0 && K % GEMM_BK == 0, "2-CTA gemm shape not tile-aligned");
    const int tiles = (M / 256) * (N / 256) * ksplit;
    int pairs = I->sms / 2;
    if (tiles < pairs) pairs = tiles;
    LaunchCfg lc(2 * pairs, GEMM_THREADS, GEMM2_SMEM_BYTES, s, 2);
    const int rev = I->next_dir();
    if (fmt_is_split(FMT)) { EG_CUDA(cudaLaunchKernelEx(&lc.cfg, kern, A.mhi, A.mlo, W.mhi128, W.mlo128, M, N, K, epi, rev, 1, ksplit)); }
    else if (I->w_pair)   { EG_CUDA(cudaLaunchKernelEx(&lc.cfg, kern, A.m16, A.m16, W.mlo128, W.mhi128, M, N, K, epi, rev, 2, 1)); }
    else                  { EG_CUDA(cudaLaunchKernelEx(&lc.cfg, kern, A.m16, A.m16, W.m16_128, W.m16_128, M, N, K, epi, rev, 1, 1)); }
    return 0;
}

// CTA-pair fp16 GEMM with the TMA-store epilogue (default for the fp16-format QKV projection and FFN w_1;
// EGOEGO_TMA_EPI=0 keeps the transposing epilogue)
static bool use_tma_epi() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("EGOEGO_TMA_EPI"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1 && use_2cta();
}
// cluster-of-8 variant (2 x 2 CTA pairs, A and W multicast): co-resident clusters on this device, 0 if the launch is not possible
template <class Epi>
static int c8_query(TcImpl* I) {
    auto kern = gemm_half_tma_2cta_kernel<Epi, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmTmaEpiCfg::SMEM_BYTES) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((I->sms / 8) * 8); cfg.blockDim = dim3(GEMM_TMAEPI_THREADS); cfg.dynamicSmemBytes = GemmTmaEpiCfg::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return nc < I->sms / 8 ? nc : I->sms / 8;
}
template <class Epi>
static int launch_gemm_tma_c8(TcImpl* I, const Plane& A, const Plane& W, int M, int N, int K, const float* bias, const Epi& epi, cudaStream_t s) {
    auto kern = gemm_half_tma_2cta_kernel<Epi, true>;
    EG_CHECK(M % 512 == 0 && N % 512 == 0 && K % GEMM_BK == 0 && N <= GemmTmaEpiCfg::MAX_N && I->c8_clusters > 0, "cluster-of-8 gemm shape not supported");
    const int super_tiles = (M / 512) * (N / 512);
    const int clusters = super_tiles < I->c8_clusters ? super_tiles : I->c8_clusters;
    LaunchCfg lc(8 * clusters, GEMM_TMAEPI_THREADS, GemmTmaEpiCfg::SMEM_BYTES, s, 8);
    EG_CUDA(cudaLaunchKernelEx(&lc.cfg, kern, A.m16_64, W.m16_64, M, N, K, bias, epi, I->next_dir(), W.m16_64, 1, (int*)nullptr));
    return 0;
}

// `row_cnt` != nullptr: streamed consumer (see TcImpl::stream_att): at most `max_pairs` CTA pairs, per-window completion counters
// published by the store warp; *pairs_used / *rev_used report the grid and the tile direction to the caller (the consumer follows both).
template <class Epi>
static int launch_gemm_tma_epi(TcImpl* I, const Plane& A, const Plane& W, int M, int N, int K, const float* bias, const Epi& epi, cudaStream_t s,
                               int* row_cnt = nullptr, int max_pairs = 0, int* pairs_used = nullptr, int* rev_used = nullptr) {
    if (I->c8_clusters > 0 && M % 512 == 0 && N % 512 == 0 && !I->w_pair && !row_cnt) return launch_gemm_tma_c8(I, A, W, M, N, K, bias, epi, s);
    static PerDeviceOnce attr_once;          // opt-in shared-memory size is a per-device function attribute
    auto kern = gemm_half_tma_2cta_kernel<Epi>;
    if (attr_once.need()) {
        EG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmTmaEpiCfg::SMEM_BYTES));
    }
    EG_CHECK(M % 256 == 0 && N % 256 == 0 && K % GEMM_BK == 0 && N <= GemmTmaEpiCfg::MAX_N, "TMA-epilogue gemm shape not supported");
    const int tiles = (M / 256) * (N / 256);
    int pairs = I->sms / 2;
    if (row_cnt && max_pairs > 0 && max_pairs < pairs) pairs = max_pairs;
    if (tiles < pairs) pairs = tiles;
    LaunchCfg lc(2 * pairs, GEMM_TMAEPI_THREADS, GemmTmaEpiCfg::SMEM_BYTES, s, 2);
    const int rev = I->next_dir();
    if (pairs_used) *pairs_used = pairs;
    if (rev_used) *rev_used = rev;
    if (I->w_pair) { EG_CUDA(cudaLaunchKernelEx(&lc.cfg, kern, A.m16, W.mlo128, M, N, K, bias, epi, rev, W.mhi128, 2, row_cnt)); }
    else           { EG_CUDA(cudaLaunchKernelEx(&lc.cfg, kern, A.m16, W.m16_128, M, N, K, bias, epi, rev, W.m16_128, 1, row_cnt)); }
    return 0;
}

template <int FMT, class Epi>
static int gemm(TcImpl* I, const Plane& A, const Plane& W, int M, int N, int K, const Epi& epi, cudaStream_t s) {
    if (use_2cta() && M % 256 == 0) return launch_gemm_2cta<FMT>(I, A, W, M, N, K, epi, s);
    return launch_gemm<256, FMT>(I, A, W, M, N, K, epi, s);
}

TcEngine::TcEngine() : impl_(nullptr) {}
TcEngine::~TcEngine() { delete impl_; }

int TcEngine::init(const TcWeights& w, cudaStream_t) {
    delete impl_;
    impl_ = new TcImpl();
    TcImpl* I = impl_;
    I->w = w;
    const int D = w.D, d = w.d, H = w.H, dk = w.dk;
    EG_CHECK(d == 512 && dk == 256, "tensor engine is specialised for d_model = 512, d_k = d_v = 256");
    int dev = 0;
    EG_CUDA(cudaGetDevice(&dev));
    EG_CUDA(cudaDeviceGetAttribute(&I->sms, cudaDevAttrMultiProcessorCount, dev));
    I->M = ((w.max_batch + 1) / 2) * 2 * LP;          // even number of windows: 2-CTA tiles cover 256 rows
    I->kx = ((D + 63) / 64) * 64;                 // 198 -> 256
    I->nout = ((D + 255) / 256) * 256;            // 198 -> 256 (zero rows)
    const size_t M = I->M;
    // weights: [N, K] K-major planes, B-operand boxes of 256 rows
    if (upload_weight(I->Wx, w.start_w, d, 2 * D, 0, D, d, I->kx, 256)) return 1;
    if (upload_weight(I->Wc, w.start_w, d, 2 * D, D, D, d, I->kx, 256)) return 1;
    if (upload_weight(I->Wout, w.out_w, D, d, 0, d, I->nout, d, 256)) return 1;
    I->layers.resize(w.NL);
    for (int l = 0; l < w.NL; ++l) {
        const TcLayerW& s = w.layers[l];
        TcLayer& L = I->layers[l];
        std::vector<float> qkv((size_t)3 * H * dk * d);
        memcpy(&qkv[0], s.wq, (size_t)H * dk * d * 4);
        memcpy(&qkv[(size_t)H * dk * d], s.wk, (size_t)H * dk * d * 4);
        memcpy(&qkv[(size_t)2 * H * dk * d], s.wv, (size_t)H * dk * d * 4);
        if (upload_weight(L.wqkv, qkv.data(), 3 * H * dk, d, 0, d, 3 * H * dk, d, 256)) return 1;
        if (upload_weight(L.fc, s.fc, d, H * dk, 0, H * dk, d, H * dk, 256)) return 1;
        if (upload_weight(L.w1, s.w1, d, d, 0, d, d, d, 256)) return 1;
        if (upload_weight(L.w2, s.w2, d, d, 0, d, d, d, 256)) return 1;
        L.bqkv = s.bqkv; L.fc_b = s.fc_b; L.b1 = s.b1; L.b2 = s.b2;
        L.ln1_g = s.ln1_g; L.ln1_b = s.ln1_b; L.ln2_g = s.ln2_g; L.ln2_b = s.ln2_b;
    }
    // activation planes: A-operand boxes of 128 rows
    if (I->X.alloc(M, I->kx, 128) || I->X.alloc16() || I->C.alloc(M, I->kx, 128) || I->Hs.alloc(M, d, 128) ||
        I->O.alloc(M, (size_t)H * dk, 128) || I->F.alloc(M, d, 128)) return 1;
    EG_CUDA(cudaMalloc(&I->base, M * d * 4));
    EG_CUDA(cudaMalloc(&I->H, M * d * 4));
    EG_CUDA(cudaMalloc(&I->Y, M * d * 4));
    {
        const char* fl = getenv("EGOEGO_FUSE_LN");
        I->fuse_ln = !(fl && fl[0] == '0');
        EG_CUDA(cudaFuncSetAttribute(gemm_ln_half_c4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmLn4Cfg::SMEM_BYTES));
        // column-split cluster-of-4 kernel: launch exactly as many clusters as can be co-resident (GPC boundaries may leave a few
        // SMs without a complete cluster); if the query fails the fp16 steps fall back to the unfused GEMM + LayerNorm kernels
        if (use_2cta()) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((I->sms / 4) * 4); cfg.blockDim = dim3(GEMM_LN4_THREADS); cfg.dynamicSmemBytes = GemmLn4Cfg::SMEM_BYTES;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int nc = 0;
            if (cudaOccupancyMaxActiveClusters(&nc, gemm_ln_half_c4_kernel, &cfg) == cudaSuccess && nc > 0)
                I->ln4_clusters = nc < I->sms / 4 ? nc : I->sms / 4;
            else
                cudaGetLastError();
        }
        if (I->ln4_clusters == 0) I->fuse_ln = false;
    }
    { const char* zz = getenv("EGOEGO_ZIGZAG"); I->zigzag = !(zz && zz[0] == '0'); }
    {
        const char* sa = getenv("EGOEGO_STREAM_ATT");
        I->stream_att = sa && sa[0] == '1';
        const char* sp = getenv("EGOEGO_STREAM_ATT_PAIRS");
        if (sp && sp[0]) I->stream_pairs = atoi(sp);
        if (I->stream_pairs < 1) I->stream_pairs = 1;
        if (I->stream_pairs > I->sms / 2 - 2) I->stream_pairs = I->sms / 2 - 2;       // the consumer keeps at least four SMs
        const char* sr = getenv("EGOEGO_STREAM_ATT_RATIO");
        if (sr && sr[0]) I->stream_ratio = atof(sr);
        if (!(I->stream_ratio > 0.05 && I->stream_ratio < 20.0)) I->stream_ratio = 0.96;
        I->cnt_ld = ((w.max_batch + 1) / 2) * 2;
        EG_CUDA(cudaMalloc(&I->att_cnt, (size_t)w.NL * I->cnt_ld * sizeof(int)));
        EG_CUDA(cudaMemset(I->att_cnt, 0, (size_t)w.NL * I->cnt_ld * sizeof(int)));
    }
    {   // cluster-of-8 multicast GEMM for the QKV projection and w_1 (opt-in until measured: EGOEGO_GEMM_C8=1)
        const char* c8 = getenv("EGOEGO_GEMM_C8");
        I->c8_clusters = 0;
        if (c8 && c8[0] == '1' && use_2cta()) {
            const int a = c8_query<TmaEpiQKV>(I), b = c8_query<TmaEpiRelu>(I);
            I->c8_clusters = a < b ? a : b;
        }
    }
    const char* am = getenv("EGOEGO_ATTN");
    I->attn_tc = !(am && strcmp(am, "simt") == 0);
    if (I->attn_tc) {
        const size_t MBe = (size_t)((w.max_batch + 1) / 2) * 2;
        if (I->Qp.alloc(MBe * H * 128, 256, 128) || I->Kp.alloc(MBe * H * 128, 256, 128) ||
            I->VT.alloc(MBe * H * 128, 256, 64)) return 1;
        EG_CUDA(cudaFuncSetAttribute(attention_tc_kernel<FMT_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
        EG_CUDA(cudaFuncSetAttribute(attention_tc_kernel<FMT_HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
        EG_CUDA(cudaFuncSetAttribute(attention_half_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM_BYTES));
        I->attn_v2 = !(am && strcmp(am, "v1") == 0);
    } else {
        EG_CUDA(cudaMalloc(&I->QKV, M * 3 * H * dk * 4));
        EG_CUDA(cudaFuncSetAttribute(attention_simt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SIMT_SMEM));
    }
    return 0;
}

int TcEngine::clear_staging(int B, cudaStream_t s) {
    TcImpl* I = impl_;
    const size_t n = (size_t)B * LP * I->kx * 2;
    EG_CUDA(cudaMemsetAsync(I->X.hi, 0, n, s)); EG_CUDA(cudaMemsetAsync(I->X.lo, 0, n, s));
    EG_CUDA(cudaMemsetAsync(I->C.hi, 0, n, s)); EG_CUDA(cudaMemsetAsync(I->C.lo, 0, n, s));
    EG_CUDA(cudaMemsetAsync(I->X.h16, 0, n, s));
    return 0;
}

int TcEngine::stage(const float* src, int src_ld, int src_col0, bool cond_half, int B, int T, cudaStream_t s, int64_t* n) {
    TcImpl* I = impl_;
    Plane& P = cond_half ? I->C : I->X;
    const long long tot = (long long)B * T * I->w.D;
    stage_rows_split_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(P.hi, P.lo, cond_half ? nullptr : reinterpret_cast<__half*>(P.h16), I->kx, src, src_ld, src_col0, I->w.D, B, T);
    EG_CUDA(cudaGetLastError());
    *n += 1;
    return 0;
}

int TcEngine::prepare_cond(int B, int T, cudaStream_t s, int64_t* n) {
    TcImpl* I = impl_;
    TcEpiBase e{{}, {}, I->base, I->w.d, I->w.start_b, I->w.pos, T};
    if (gemm<FMT_SPLIT>(I, I->C, I->Wc, Mr(B), I->w.d, I->kx, e, s)) return 1;
    *n += 1;
    return 0;
}

std::string TcEngine::info() const {
    const TcImpl* I = impl_;
    if (!I) return "engine=tcgen05 (not initialised)";
    return "engine=tcgen05 sms=" + std::to_string(I->sms) + " ln4_clusters=" + std::to_string(I->ln4_clusters) + " c8_clusters=" + std::to_string(I->c8_clusters) +
           " zigzag=" + std::to_string(I->zigzag ? 1 : 0) + " fuse_ln=" + std::to_string(I->fuse_ln ? 1 : 0) + " stream_att=" + std::to_string(I->stream_att ? I->stream_pairs : 0) + " weight_sets=" + std::to_string(n_weight_sets());
}

int TcEngine::n_weight_sets() const { return impl_ && !impl_->Wx.set16.empty() ? (int)impl_->Wx.set16.size() : 1; }
void TcEngine::set_dual_acc(bool on) { if (impl_) impl_->dual_acc = on; }
void TcEngine::set_weight_pair(bool on) { if (impl_) impl_->w_pair = on; }
int TcEngine::reset_stream_counters(cudaStream_t s) {
    TcImpl* I = impl_;
    if (!I || !I->att_cnt) return 0;
    EG_CUDA(cudaMemsetAsync(I->att_cnt, 0, (size_t)I->w.NL * I->cnt_ld * sizeof(int), s));
    return 0;
}
void TcEngine::use_weight_set(int r) {
    TcImpl* I = impl_;
    if (!I) return;
    for (Plane* p : {&I->Wx, &I->Wc, &I->Wout}) p->use_set(r);
    for (auto& l : I->layers) for (Plane* p : {&l.wqkv, &l.fc, &l.w1, &l.w2}) p->use_set(r);
}

void TcEngine::stage_targets(__nv_bfloat16** hi, __nv_bfloat16** lo, __half** h16, int* ld) {
    *hi = impl_->X.hi; *lo = impl_->X.lo; *h16 = reinterpret_cast<__half*>(impl_->X.h16); *ld = impl_->kx;
}

int TcEngine::launches_per_denoiser(int fmt) const {
    const bool fused = fmt == FMT_HALF && impl_->fuse_ln && impl_->attn_tc;
    return 2 + (fused ? 5 : 7) * impl_->w.NL;
}

// fp16 GEMM + residual + bias + LayerNorm (gemm_ln_half_c4_kernel); residual in / output out = the Hs fp16 plane (in place)
static int launch_gemm_ln(TcImpl* I, const Plane& A, const Plane& W, int M, int K, const float* bias, const float* g,
                          const float* b, cudaStream_t s) {
    EG_CHECK(M % 256 == 0 && K % GEMM_BK == 0 && I->ln4_clusters > 0, "fused-LN gemm shape not tile-aligned");
    const int tiles = M / 256;
    const int clusters = tiles < I->ln4_clusters ? tiles : I->ln4_clusters;
    LaunchCfg lc(4 * clusters, GEMM_LN4_THREADS, GemmLn4Cfg::SMEM_BYTES, s, 4);
    if (I->w_pair) { EG_CUDA(cudaLaunchKernelEx(&lc.cfg, gemm_ln_half_c4_kernel, A.m16_128, W.mlo128, I->Hs.m16_128, M, K, bias, g, b, I->next_dir(), W.mhi128, 2)); }
    else           { EG_CUDA(cudaLaunchKernelEx(&lc.cfg, gemm_ln_half_c4_kernel, A.m16_128, W.m16_128, I->Hs.m16_128, M, K, bias, g, b, I->next_dir(), W.m16_128, 1)); }
    return 0;
}

// linear_out with the DDPM update of the sampling loop in its epilogue (p_mean_variance / q_posterior / p_sample,
// egoego/model/transformer_cond_diffusion_model.py:216-256): x0 = acc + bias (pred_x0) -> clamp -> posterior mean -> + sigma * noise
// -> in-paint -> x_out (fp32, compact) and the next step's start_conv operand plane(s).  Same arithmetic and the same
// Philox stream (window, draw, flat element quad) as ddpm_update_kernel; model_out is never written.
template <int FMT>
struct TcEpiOutDdpm : EpiNoDirect, EpiNoPre {
    DdpmArgs a; const float* bias; int n_windows;
    __device__ __forceinline__ float4 bias4(int col) const {
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < a.D) { b.x = bias[col]; b.y = bias[col + 1]; }
        if (col + 2 < a.D) { b.z = bias[col + 2]; b.w = bias[col + 3]; }
        return b;
    }
    __device__ __forceinline__ void apply4(int row, int col, float4 acc, float4 b, float4) const {
        const int w = row / LP, l = row % LP;
        if (l < 1 || l > a.T || w >= n_windows || col >= a.D) return;
        const int f = l - 1, e0 = f * a.D + col;                      // flat element inside the window (even)
        const bool four = col + 2 < a.D;
        const long long i0 = (long long)w * a.T * a.D + e0;
        const int t = a.ts.get(w);
        const float c1 = a.coef1[t], c2 = a.coef2[t];
        const float sigma = (t == 0) ? 0.f : expf(0.5f * a.logvar[t]);
        const float k_mo = (a.objective == 0) ? -a.sqrt_recipm1[t] : 1.0f, k_x = (a.objective == 0) ? a.sqrt_recip[t] : 0.0f;
        float mo[4] = {acc.x + b.x, acc.y + b.y, acc.z + b.z, acc.w + b.w};
        float xv[4] = {0.f, 0.f, 0.f, 0.f}, nz[4] = {0.f, 0.f, 0.f, 0.f};
        { const float2 q = *reinterpret_cast<const float2*>(a.x + i0); xv[0] = q.x; xv[1] = q.y; }
        if (four) { const float2 q = *reinterpret_cast<const float2*>(a.x + i0 + 2); xv[2] = q.x; xv[3] = q.y; }
        const int draw = a.ns.draw();
        if (a.ns.tape) {
            const float* tp = a.ns.tape + (long long)draw * a.ns.draw_stride + i0;
            nz[0] = tp[0]; nz[1] = tp[1];
            if (four) { nz[2] = tp[2]; nz[3] = tp[3]; }
        } else if (sigma != 0.f) {
            // elements are grouped in quads of the flat window index; rows with odd f start in the middle of a quad
            const float4 n0 = philox_normal4(a.ns.seed, a.ns.window_offset + w, (uint32_t)draw, (uint32_t)(e0 >> 2));
            if ((e0 & 2) == 0) { nz[0] = n0.x; nz[1] = n0.y; nz[2] = n0.z; nz[3] = n0.w; }
            else {
                nz[0] = n0.z; nz[1] = n0.w;
                if (four) { const float4 n1 = philox_normal4(a.ns.seed, a.ns.window_offset + w, (uint32_t)draw, (uint32_t)(e0 >> 2) + 1u); nz[2] = n1.x; nz[3] = n1.y; }
            }
        }
        float v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float x0 = (a.objective == 0) ? (k_x * xv[r] + k_mo * mo[r]) : mo[r];
            if (a.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
            v[r] = (c1 * x0 + c2 * xv[r]) + sigma * nz[r];
        }
        if (a.inpaint && f < a.inpaint_len) {
            const float* ip = a.inpaint + ((long long)w * a.inpaint_len + f) * a.D + col;
            v[0] = ip[0]; v[1] = ip[1];
            if (four) { v[2] = ip[2]; v[3] = ip[3]; }
        }
        *reinterpret_cast<float2*>(a.x_out + i0) = make_float2(v[0], v[1]);
        if (four) *reinterpret_cast<float2*>(a.x_out + i0 + 2) = make_float2(v[2], v[3]);
        const long long o = ((long long)w * LP + l) * a.stage_ld16 + col;      // plane columns [D, ld16) stay zero
        if (FMT == FMT_SPLIT) {
            __nv_bfloat16 h0, l0, h1, l1;
            split_f16(v[0], h0, l0); split_f16(v[1], h1, l1);
            *reinterpret_cast<__nv_bfloat162*>(a.stage_hi + o) = __nv_bfloat162(h0, h1);
            *reinterpret_cast<__nv_bfloat162*>(a.stage_lo + o) = __nv_bfloat162(l0, l1);
            if (four) {
                split_f16(v[2], h0, l0); split_f16(v[3], h1, l1);
                *reinterpret_cast<__nv_bfloat162*>(a.stage_hi + o + 2) = __nv_bfloat162(h0, h1);
                *reinterpret_cast<__nv_bfloat162*>(a.stage_lo + o + 2) = __nv_bfloat162(l0, l1);
            }
        } else {
            *reinterpret_cast<__half2*>(a.stage_h16 + o) = __floats2half2_rn(v[0], v[1]);
            if (four) *reinterpret_cast<__half2*>(a.stage_h16 + o + 2) = __floats2half2_rn(v[2], v[3]);
        }
    }
};

// Stages of one denoiser call (egoego_time_kernel's `which`): 0 start, 1 QKV projection, 2 attention, 3 fc (+LN),
// 4 w_1, 5 w_2 (+LN), 6 linear_out.  `only` < 0 runs everything; otherwise just that stage of layer 0 (timing hook).
template <int FMT>
static int denoiser_impl(TcImpl* I, int B, int T, TSrc ts, const float* pmask, float* model_out, cudaStream_t s, int only = -1,
                         const DdpmArgs* fuse = nullptr) {
    const int M = B * LP, d = I->w.d, H = I->w.H, dk = I->w.dk, L = T + 1;
    const int Mg = Mr(B);                          // GEMM rows: whole 256-row tiles (an odd window count is rounded up)
    const int nqkv = 3 * H * dk;
    const int half = (FMT == FMT_HALF) ? 1 : 0;
    const bool fused = (FMT == FMT_HALF) && I->fuse_ln && I->attn_tc && pmask == nullptr;
    auto on = [&](int stage, int l) { return only < 0 || (only == stage && l == 0); };
    I->dir = 0;                                    // a call always starts ascending: 22 kernels per call, so replayed graphs stay in phase
    if (on(0, 0)) {
        TcEpiStart<FMT> e{{}, fused ? nullptr : I->H, I->Hs.hi, I->Hs.lo, d, I->base, I->w.pos, I->w.temb, ts, T, B};
        if (gemm<FMT>(I, I->X, I->Wx, Mg, d, I->kx, e, s)) return 1;
    }
    for (int l = 0; l < I->w.NL; ++l) {
        TcLayer& W = I->layers[l];
        // streamed attention: loop mode only (the counters count steps since the last reset_stream_counters(), read through *d_step)
        const bool streamed = FMT == FMT_HALF && I->stream_att && I->attn_tc && I->attn_v2 && use_tma_epi() && Mg % 256 == 0 && only < 0 &&
                              ts.d_step != nullptr && ts.t_arr == nullptr && I->c8_clusters == 0 && I->att_cnt != nullptr;
        int* cnt_l = streamed ? I->att_cnt + (size_t)l * I->cnt_ld : nullptr;
        int q_pairs = 0, q_rev = 0;
        if (I->attn_tc) {
            if (on(1, l)) {
                TcEpiQKVPlanes<FMT> eq{{}, {}, I->Qp.hi, I->Qp.lo, I->Kp.hi, I->Kp.lo, I->VT.hi, I->VT.lo, W.bqkv, H, 1.0f / sqrtf((float)dk)};
                if (FMT == FMT_HALF && use_tma_epi() && Mg % 256 == 0) {
                    TmaEpiQKV te{I->Qp.m16_128, I->Kp.m16_128, I->VT.m16_128, H, 1.0f / sqrtf((float)dk)};
                    if (launch_gemm_tma_epi(I, I->Hs, W.wqkv, Mg, nqkv, d, W.bqkv, te, s, cnt_l, I->stream_pairs, &q_pairs, &q_rev)) return 1;
                } else {
                    if (gemm<FMT>(I, I->Hs, W.wqkv, Mg, nqkv, d, eq, s)) return 1;
                }
            }
            if (on(2, l)) {
                const int items = B * H;
                if (FMT == FMT_HALF && I->attn_v2) {
                    if (streamed) {
                        // Concurrently with the projection, in the SAME window order.  Two phases: G = sms - 2 q_pairs CTAs are resident next to
                        // the projection and take the first X items; the other 2 q_pairs CTAs become resident as the projection's CTAs exit and
                        // take the rest.  X balances the two finishing times for a per-item attention time of `stream_ratio` projection-tile times:
                        //   X r / G = tiles / q_pairs + (items - X) r / (2 q_pairs),   then (items - X) rounded to whole waves of the late CTAs
                        const int G = I->sms - 2 * q_pairs, late = 2 * q_pairs;
                        const int tiles = (Mg / 256) * (nqkv / 256);
                        int X = items, early = G < items ? G : items, total = early;
                        if (G < items) {
                            const double r = I->stream_ratio;
                            const double x = ((double)tiles / q_pairs + (double)items * r / late) / (r / G + r / late);
                            double restf = (double)items - x;
                            int rest = restf >= late ? (int)(restf / late + 0.5) * late : (int)(restf + 0.999);   // whole waves once there is one
                            if (rest < 0) rest = 0;
                            if (rest > items - 1) rest = items - 1;
                            X = items - rest;
                            if (early > X) early = X;
                            total = early + (rest < late ? rest : late);
                        }
                        LaunchCfg lc(total, ATT2_THREADS, ATT2_SMEM_BYTES, s);
                        EG_CUDA(cudaLaunchKernelEx(&lc.cfg, attention_half_kernel, I->Qp.m16, I->Kp.m16, I->VT.m16, I->O.m16_128, items, H, L, q_rev,
                                                   (const int*)cnt_l, ts.d_step, nqkv / 256, total > early ? early : 0, X));
                    } else {
                        LaunchCfg lc(items < I->sms ? items : I->sms, ATT2_THREADS, ATT2_SMEM_BYTES, s);
                        EG_CUDA(cudaLaunchKernelEx(&lc.cfg, attention_half_kernel, I->Qp.m16, I->Kp.m16, I->VT.m16, I->O.m16_128, items, H, L, I->next_dir(),
                                                   (const int*)nullptr, (const int*)nullptr, 0, 0, 0));
                    }
                } else {
                    LaunchCfg lc(items < I->sms ? items : I->sms, ATT_THREADS, ATT_SMEM_BYTES, s);
                    EG_CUDA(cudaLaunchKernelEx(&lc.cfg, attention_tc_kernel<FMT>, I->Qp.mhi, I->Qp.mlo, I->Kp.mhi, I->Kp.mlo, I->VT.mhi, I->VT.mlo,
                                               I->O.hi, I->O.lo, H * dk, items, H, L));
                }
            }
        } else {
            if (on(1, l)) {
                TcEpiBiasScaleF32 eq{{}, {}, I->QKV, nqkv, W.bqkv, H * dk, 1.0f / sqrtf((float)dk)};
                if (gemm<FMT_SPLIT>(I, I->Hs, W.wqkv, Mg, nqkv, d, eq, s)) return 1;
            }
            if (on(2, l))
                attention_simt_kernel<true><<<B * H, 256, ATT_SIMT_SMEM, s>>>(I->QKV, nqkv, nullptr, I->O.hi, I->O.lo, H * dk, H, L);
        }
        if (on(3, l)) {
            if (fused) {
                if (launch_gemm_ln(I, I->O, W.fc, Mg, H * dk, W.fc_b, W.ln1_g, W.ln1_b, s)) return 1;
            } else {
                TcEpiBiasResidF32 ef{{}, I->Y, d, W.fc_b, I->H};
                if (gemm<FMT>(I, I->O, W.fc, Mg, d, H * dk, ef, s)) return 1;
                LaunchCfg ll(M / 8, 256, 0, s);
                EG_CUDA(cudaLaunchKernelEx(&ll.cfg, layernorm512_kernel, (const float*)I->Y, I->H, I->Hs.hi, I->Hs.lo, W.ln1_g, W.ln1_b, pmask, T, M, half));
            }
        }
        if (on(4, l)) {
            if (FMT == FMT_HALF && use_tma_epi() && Mg % 256 == 0) {
                TmaEpiRelu te{I->F.m16_128};
                if (launch_gemm_tma_epi(I, I->Hs, W.w1, Mg, d, d, W.b1, te, s)) return 1;
            } else {
                TcEpiBiasReluSplit<FMT> e1{{}, {}, I->F.hi, I->F.lo, d, W.b1};
                if (gemm<FMT>(I, I->Hs, W.w1, Mg, d, d, e1, s)) return 1;
            }
        }
        if (on(5, l)) {
            if (fused) {
                if (launch_gemm_ln(I, I->F, W.w2, Mg, d, W.b2, W.ln2_g, W.ln2_b, s)) return 1;
            } else {
                TcEpiBiasResidF32 e2{{}, I->Y, d, W.b2, I->H};
                if (gemm<FMT>(I, I->F, W.w2, Mg, d, d, e2, s)) return 1;
                LaunchCfg ll(M / 8, 256, 0, s);
                EG_CUDA(cudaLaunchKernelEx(&ll.cfg, layernorm512_kernel, (const float*)I->Y, I->H, I->Hs.hi, I->Hs.lo, W.ln2_g, W.ln2_b, pmask, T, M, half));
            }
        }
    }
    if (on(6, 0)) {
        if (fuse) {
            TcEpiOutDdpm<FMT> eo{{}, {}, *fuse, I->w.out_b, B};
            if (gemm<FMT>(I, I->Hs, I->Wout, Mg, I->nout, d, eo, s)) return 1;
        } else {
            TcEpiOut eo{{}, {}, model_out, I->w.D, I->w.out_b, T, B};
            if (gemm<FMT>(I, I->Hs, I->Wout, Mg, I->nout, d, eo, s)) return 1;
        }
    }
    EG_CUDA(cudaGetLastError());
    return 0;
}

int TcEngine::denoiser(int B, int T, TSrc ts, const float* pmask, float* model_out, cudaStream_t s, int64_t* n, int fmt, const DdpmArgs* fuse) {
    TcImpl* I = impl_;
    EG_CHECK(fmt == FMT_SPLIT || (I->attn_tc), "fp16 single-pass steps need the tensor-core attention");
    EG_CHECK(!fuse || (fuse->stage_ld16 == I->kx && (fmt == FMT_HALF ? fuse->stage_h16 != nullptr : (fuse->stage_hi && fuse->stage_lo))),
             "fused DDPM update needs the engine's own staging planes");
    int rc = (fmt == FMT_HALF) ? denoiser_impl<FMT_HALF>(I, B, T, ts, pmask, model_out, s, -1, fuse)
                               : denoiser_impl<FMT_SPLIT>(I, B, T, ts, pmask, model_out, s, -1, fuse);
    *n += launches_per_denoiser(fmt);
    return rc;
}

// Time ONE stage of the denoiser (see denoiser_impl; layer 0's weights) in isolation on the engine's own buffers:
// `iters` back-to-back launches between two CUDA events on stream `s`.  The workspace keeps whatever the last
// sampling call left in it (finite activations), which the stage overwrites in place as it does inside a step.
int TcEngine::time_stage(int B, int T, int stage, int fmt, int iters, float* model_out, cudaStream_t s, float* ms, const DdpmArgs* fuse) {
    TcImpl* I = impl_;
    EG_CHECK(I && I->attn_tc, "time_stage needs the tensor-core attention layout");
    EG_CHECK(B >= 1 && B <= I->w.max_batch && iters >= 1 && stage >= 0 && stage <= 6, "bad arguments");
    cudaEvent_t e0, e1;
    EG_CUDA(cudaEventCreate(&e0)); EG_CUDA(cudaEventCreate(&e1));
    TSrc ts{nullptr, nullptr, 0, 0};
    auto run = [&]() -> int {
        return fmt == FMT_HALF ? denoiser_impl<FMT_HALF>(I, B, T, ts, nullptr, model_out, s, stage, fuse)
                               : denoiser_impl<FMT_SPLIT>(I, B, T, ts, nullptr, model_out, s, stage, fuse);
    };
    for (int i = 0; i < 3; ++i) if (run()) return 1;
    EG_CUDA(cudaEventRecord(e0, s));
    for (int i = 0; i < iters; ++i) if (run()) return 1;
    EG_CUDA(cudaEventRecord(e1, s));
    EG_CUDA(cudaEventSynchronize(e1));
    float t = 0.f;
    EG_CUDA(cudaEventElapsedTime(&t, e0, e1));
    *ms = t / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

// Debug timeline of the streamed QKV projection / attention pair (common.cuh: g_timeline): switch recording on / off, or copy the
// timestamps of the last launch ([kernel][cta][start, end], nanoseconds of %globaltimer) to the host.
int tc_debug_timeline(int enable, unsigned long long* out, int n) {
    if (out) {
        EG_CHECK(n <= 2 * TIMELINE_CTAS * 2, "timeline buffer too large");
        EG_CUDA(cudaDeviceSynchronize());
        EG_CUDA(cudaMemcpyFromSymbol(out, g_timeline, (size_t)n * sizeof(unsigned long long)));
    }
    if (enable >= 0) {
        const int v = enable ? 1 : 0;
        unsigned long long z[2 * TIMELINE_CTAS * 2] = {};
        EG_CUDA(cudaMemcpyToSymbol(g_timeline, z, sizeof(z)));
        EG_CUDA(cudaMemcpyToSymbol(g_timeline_on, &v, sizeof(int)));
    }
    return 0;
}

// ---- self test: split GEMM against the fp32 SIMT GEMM on random data ------------------------------
__global__ void fill_uniform_kernel(float* p, long long n, unsigned long long seed, float scale) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 7u, 9u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    p[i] = ((float)r.x * 2.3283064365386963e-10f * 2.0f - 1.0f) * scale;
}
__global__ void split_rows_kernel(const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    split_f16(src[i], hi[i], lo[i]);
}
__global__ void to_half_kernel(const float* src, __half* dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2half_rn(src[i]);
}
__global__ void maxdiff_kernel(const float* a, const float* b, long long n, float* out /*[2]: max|a-b|, max|b|*/) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float d = 0.f, m = 0.f;
    if (i < n) { d = fabsf(a[i] - b[i]); m = fabsf(b[i]); if (!(d == d)) d = INFINITY; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { d = fmaxf(d, __shfl_xor_sync(0xffffffffu, d, o)); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMax(reinterpret_cast<int*>(out), __float_as_int(d)); atomicMax(reinterpret_cast<int*>(out) + 1, __float_as_int(m)); }
}

// ---- generic fp32 product on the tensor cores (training step) ------------------------------------------------------
// C[M, N] (= or +=) A[M, K] W[N, K]^T for fp32 row-major operands in device memory: both are split into bf16 hi/lo planes
// (zero-padded to the tile grid: rows to 256, K to 64) and multiplied with the 3-term split CTA-pair kernel -- fp32-grade
// results at tensor-core speed.  Planes (with their TMA maps) are cached per padded shape and role.
__global__ void split_pad_kernel(const float* __restrict__ src, int rows, int cols, int ld, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, int rows_pad, int cols_pad) {
    // four consecutive columns per thread (cols_pad % 64 == 0); float4 loads when the source row allows it
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int qpr = cols_pad / 4;
    if (q >= (long long)rows_pad * qpr) return;
    const int r = (int)(q / qpr), c = (int)(q % qpr) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < rows) {
        const float* p = src + (long long)r * ld + c;
        if (c + 3 < cols && ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
            const float4 f = *reinterpret_cast<const float4*>(p);
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (c + j < cols) v[j] = p[j];
        }
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(v[j], h[j], l[j]);
    const long long o = (long long)r * cols_pad + c;
    __nv_bfloat162 h01(h[0], h[1]), h23(h[2], h[3]), l01(l[0], l[1]), l23(l[2], l[3]);
    *reinterpret_cast<uint2*>(hi + o) = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
    *reinterpret_cast<uint2*>(lo + o) = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
}

// The same conversion for an operand given TRANSPOSED: src is [cols, rows] row-major (leading dimension ld), the planes are
// [rows_pad, cols_pad] -- plane[r][k] = src[k][r].  32 x 32 tiles through shared memory, coalesced on both sides.  Lets the weight
// gradients (dY^T X: both operands are stored with the reduction index as the ROW) and the input gradients (dY W: the weight is
// stored [out, in]) feed the K-major GEMM without separate fp32 transposes.
__global__ void split_pad_t_kernel(const float* __restrict__ src, int rows, int cols, int ld, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ lo, int rows_pad, int cols_pad) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;          // plane rows r0.., plane columns k0..
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + threadIdx.y + 8 * i, r = r0 + threadIdx.x;
        tile[threadIdx.y + 8 * i][threadIdx.x] = (k < cols && r < rows) ? src[(long long)k * ld + r] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + threadIdx.y + 8 * i, k = k0 + threadIdx.x;
        if (r < rows_pad && k < cols_pad) {
            __nv_bfloat16 h, l;
            split_bf16(tile[threadIdx.x][threadIdx.y + 8 * i], h, l);
            hi[(long long)r * cols_pad + k] = h;
            lo[(long long)r * cols_pad + k] = l;
        }
    }
}

struct TcGemmCache {
    std::map<std::pair<long long, int>, std::unique_ptr<Plane>> planes;     // (rows_pad << 20 | cols_pad, role) -> plane
    TcImpl I;
    bool init = false;
    ~TcGemmCache() { for (auto& kv : planes) kv.second->release(); }
};
static TcGemmCache g_tcg;

int tc_gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, float* C, int ldc, int n_valid, int accumulate,
                cudaStream_t s) {
    return tc_gemm_f32_ex(A, lda, 0, W, ldw, 0, M, N, K, C, ldc, n_valid, accumulate, s);
}

// operand planes of one product (cached per padded shape, role and device), filled from the fp32 operands
static int tc_prep_planes(const float* A, int lda, int trans_a, const float* W, int ldw, int trans_w, int M, int N, int K,
                          Plane*& PA, Plane*& PW, int& Mp, int& Np, int& Kp, cudaStream_t s) {
    int dev = 0;
    EG_CUDA(cudaGetDevice(&dev));
    if (!g_tcg.init) {
        EG_CUDA(cudaDeviceGetAttribute(&g_tcg.I.sms, cudaDevAttrMultiProcessorCount, dev));
        g_tcg.init = true;
    }
    Mp = ((M + 255) / 256) * 256; Np = ((N + 255) / 256) * 256; Kp = ((K + 63) / 64) * 64;
    auto plane = [&](int rows, int cols, int role, uint32_t box) -> Plane* {
        auto key = std::make_pair(((long long)rows << 20) | cols, role + 2 * dev);      // planes live on the device that made them
        auto it = g_tcg.planes.find(key);
        if (it != g_tcg.planes.end()) return it->second.get();
        std::unique_ptr<Plane> p(new Plane());
        if (p->alloc(rows, cols, box)) return nullptr;
        Plane* raw = p.get();
        g_tcg.planes[key] = std::move(p);
        return raw;
    };
    PA = plane(Mp, Kp, 0, 128);
    PW = plane(Np, Kp, 1, 256);
    EG_CHECK(PA && PW, "tc_gemm_f32: plane allocation failed");
    if (trans_a) split_pad_t_kernel<<<dim3(Mp / 32, Kp / 32), dim3(32, 8), 0, s>>>(A, M, K, lda, PA->hi, PA->lo, Mp, Kp);
    else         split_pad_kernel<<<(unsigned)(((long long)Mp * Kp / 4 + 255) / 256), 256, 0, s>>>(A, M, K, lda, PA->hi, PA->lo, Mp, Kp);
    if (trans_w) split_pad_t_kernel<<<dim3(Np / 32, Kp / 32), dim3(32, 8), 0, s>>>(W, N, K, ldw, PW->hi, PW->lo, Np, Kp);
    else         split_pad_kernel<<<(unsigned)(((long long)Np * Kp / 4 + 255) / 256), 256, 0, s>>>(W, N, K, ldw, PW->hi, PW->lo, Np, Kp);
    return 0;
}

int tc_gemm_f32_ex(const float* A, int lda, int trans_a, const float* W, int ldw, int trans_w, int M, int N, int K, float* C, int ldc, int n_valid,
                   int accumulate, cudaStream_t s) {
    EG_CHECK(M >= 1 && N >= 1 && K >= 1 && n_valid % 4 == 0 && ldc % 4 == 0, "tc_gemm_f32: bad shape");
    Plane *PA = nullptr, *PW = nullptr;
    int Mp, Np, Kp;
    if (tc_prep_planes(A, lda, trans_a, W, ldw, trans_w, M, N, K, PA, PW, Mp, Np, Kp, s)) return 1;
    // Split-K for products with few output tiles and a long K (the weight gradients: 512 x 512 x 4096 is 4 tiles, i.e. 8 of 148 SMs for
    // 64 k-blocks): the largest power of two that keeps every CTA pair at one tile or less and a partial product at >= 8 k-blocks.
    // Partials are added with float4 atomics into a zeroed (or, when accumulating, the existing) C -- fp32 sums in arrival order.
    const int out_tiles = (Mp / 256) * (Np / 256), kb = Kp / 64, pairs = g_tcg.I.sms / 2;
    int ksplit = 1;
    static const bool allow_split = []() { const char* e = getenv("EGOEGO_TRAIN_SPLITK"); return !(e && e[0] == '0'); }();
    while (allow_split && ksplit < 16 && out_tiles * ksplit * 2 <= pairs && kb % (ksplit * 2) == 0 && kb / (ksplit * 2) >= 8) ksplit *= 2;
    if (ksplit > 1 && !accumulate) EG_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)n_valid * 4, (size_t)Mp, s));
    TcEpiPlainAcc e{{}, C, ldc, n_valid, ksplit > 1 ? 2 : accumulate};
    return launch_gemm_2cta<FMT_SPLIT_BF16>(&g_tcg.I, *PA, *PW, Mp, Np, Kp, e, s, ksplit);
}

// The forward products of the training step with their element-wise epilogue (bias, scale, ReLU, dropout, residual, time token) run
// in the GEMM's own epilogue: any functor of the fp32 engine (kernels_simt.cuh: operator()(row, col, acc)) is adapted to the
// coalesced 4-columns-per-lane layout of the tensor-core epilogue, instead of a separate pass over the raw product.
template <class E>
struct TcEpiScalar : EpiNoDirect, EpiNoPre {
    E e; int rows, n_valid;
    __device__ __forceinline__ float4 bias4(int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void apply4(int row, int col, float4 a, float4, float4) const {
        if (row >= rows) return;
        if (col < n_valid) e(row, col, a.x);
        if (col + 1 < n_valid) e(row, col + 1, a.y);
        if (col + 2 < n_valid) e(row, col + 2, a.z);
        if (col + 3 < n_valid) e(row, col + 3, a.w);
    }
};
template <class E>
int tc_gemm_f32_epi(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const E& e, cudaStream_t s) {
    EG_CHECK(M >= 1 && N >= 1 && K >= 1, "tc_gemm_f32_epi: bad shape");
    Plane *PA = nullptr, *PW = nullptr;
    int Mp, Np, Kp;
    if (tc_prep_planes(A, lda, 0, W, ldw, 0, M, N, K, PA, PW, Mp, Np, Kp, s)) return 1;
    TcEpiScalar<E> te{{}, {}, e, M, N};
    return launch_gemm_2cta<FMT_SPLIT_BF16>(&g_tcg.I, *PA, *PW, Mp, Np, Kp, te, s, 1);
}
template int tc_gemm_f32_epi<EpiStart>(const float*, int, const float*, int, int, int, int, const EpiStart&, cudaStream_t);
template int tc_gemm_f32_epi<EpiBiasScale>(const float*, int, const float*, int, int, int, int, const EpiBiasScale&, cudaStream_t);
template int tc_gemm_f32_epi<EpiBiasRelu>(const float*, int, const float*, int, int, int, int, const EpiBiasRelu&, cudaStream_t);
template int tc_gemm_f32_epi<EpiBiasDropResid>(const float*, int, const float*, int, int, int, int, const EpiBiasDropResid&, cudaStream_t);
template int tc_gemm_f32_epi<EpiPlainBias>(const float*, int, const float*, int, int, int, int, const EpiPlainBias&, cudaStream_t);

struct EpiStore { float* C; int ldc; __device__ void operator()(int r, int c, float a) const { C[(long long)r * ldc + c] = a; } };

int selftest_gemm(int M, int N, int K, unsigned long long seed, int two_cta, int half_fmt, float* max_abs_err, float* max_abs_ref, float* ms) {
    EG_CHECK(M % 128 == 0 && N % 256 == 0 && K % 64 == 0, "selftest_gemm: need M%128==0, N%256==0, K%64==0");
    TcImpl I;
    int dev = 0;
    EG_CUDA(cudaGetDevice(&dev));
    EG_CUDA(cudaDeviceGetAttribute(&I.sms, cudaDevAttrMultiProcessorCount, dev));
    float *A = nullptr, *W = nullptr, *C1 = nullptr, *C2 = nullptr, *res = nullptr;
    EG_CUDA(cudaMalloc(&A, (size_t)M * K * 4)); EG_CUDA(cudaMalloc(&W, (size_t)N * K * 4));
    EG_CUDA(cudaMalloc(&C1, (size_t)M * N * 4)); EG_CUDA(cudaMalloc(&C2, (size_t)M * N * 4)); EG_CUDA(cudaMalloc(&res, 8));
    EG_CUDA(cudaMemset(res, 0, 8)); EG_CUDA(cudaMemset(C1, 0xff, (size_t)M * N * 4));
    fill_uniform_kernel<<<(unsigned)(((long long)M * K + 255) / 256), 256>>>(A, (long long)M * K, seed, 1.0f);
    fill_uniform_kernel<<<(unsigned)(((long long)N * K + 255) / 256), 256>>>(W, (long long)N * K, seed + 1, 0.05f);
    Plane PA, PW;
    if (PA.alloc(M, K, 128) || PW.alloc(N, K, 256)) return 1;
    split_rows_kernel<<<(unsigned)(((long long)M * K + 255) / 256), 256>>>(A, PA.hi, PA.lo, (long long)M * K);
    split_rows_kernel<<<(unsigned)(((long long)N * K + 255) / 256), 256>>>(W, PW.hi, PW.lo, (long long)N * K);
    TcEpiPlain e{{}, {}, C1, N, nullptr, N};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    EG_CHECK(!two_cta || M % 256 == 0, "2-CTA self test needs M % 256 == 0");
    if (half_fmt) {
        if (PA.alloc16() || PW.alloc16()) return 1;
        to_half_kernel<<<(unsigned)(((long long)M * K + 255) / 256), 256>>>(A, reinterpret_cast<__half*>(PA.h16), (long long)M * K);
        to_half_kernel<<<(unsigned)(((long long)N * K + 255) / 256), 256>>>(W, reinterpret_cast<__half*>(PW.h16), (long long)N * K);
    }
    auto run = [&]() -> int {
        if (half_fmt) return two_cta ? launch_gemm_2cta<FMT_HALF>(&I, PA, PW, M, N, K, e, 0) : launch_gemm<256, FMT_HALF>(&I, PA, PW, M, N, K, e, 0);
        return two_cta ? launch_gemm_2cta<FMT_SPLIT>(&I, PA, PW, M, N, K, e, 0) : launch_gemm<256, FMT_SPLIT>(&I, PA, PW, M, N, K, e, 0);
    };
    if (run()) return 1;
    EG_CUDA(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    const int reps = 5;
    for (int r = 0; r < reps; ++r) if (run()) return 1;
    cudaEventRecord(e1);
    EG_CUDA(cudaDeviceSynchronize());
    float t = 0.f; cudaEventElapsedTime(&t, e0, e1); *ms = t / reps;
    EpiStore st{C2, N};
    sgemm_tn_kernel<<<dim3((N + 127) / 128, M / 128), 256>>>(A, K, W, K, N, K, st);
    maxdiff_kernel<<<(unsigned)(((long long)M * N + 255) / 256), 256>>>(C1, C2, (long long)M * N, res);
    float h[2];
    EG_CUDA(cudaMemcpy(h, res, 8, cudaMemcpyDeviceToHost));
    *max_abs_err = h[0]; *max_abs_ref = h[1];
    PA.release(); PW.release();
    cudaFree(A); cudaFree(W); cudaFree(C1); cudaFree(C2); cudaFree(res);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

}  // namespace egoego
