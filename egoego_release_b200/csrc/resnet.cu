// ResNet-18 optical-flow encoder of HeadNet (BASELINE config 4; HeadFormer with input_of_feats=False):
// egoego/model/resnet.py:5-23 wraps torchvision's resnet18 with fc -> 512; head_estimation_transformer.py:216-224 feeds it
// [T, 224, 224, 2] optical flow padded with a zero third channel.  Inference only (eval(): BatchNorm uses running statistics).
//
// One kernel per convolution: implicit GEMM over NHWC activations (C[M = N*Ho*Wo, Cout] = im2col(X) * W^T, gathered on the
// fly), BatchNorm folded into the weights at commit and bias + residual add + ReLU fused in the epilogue.  Two engines:
//   default            conv_tc_kernel (conv_tcgen05.cuh): tcgen05 tensor cores, fp16 operands (the operand rounding of the reference's
//                      own cuDNN-TF32 GPU path), fp32 accumulate, fp32 residual stream
//   EGOEGO_RESNET=simt conv_igemm_kernel below: fp32 CUDA cores, 128 x 128 x 16 tiles (validation / bisecting)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_simt.cuh"
#include "engine_tc.cuh"
#include "conv_tcgen05.cuh"

namespace egoego {

struct ConvDesc { int Cin, Cout, kh, kw, stride, pad, Hin, Win, Hout, Wout; };

// X NHWC [N,Hin,Win,Cin] (Cin % 4 == 0), Wf [Cout, Kpad] with k = (ky*kw + kx)*Cin + ci zero-padded to Kpad % 16 == 0,
// Y / resid [M, Cout] (= NHWC of the output), M = N*Hout*Wout
// BN = 128 or 64 output channels per tile (the 64-channel stem / layer1 convolutions are 64 % of the FLOPs: a 128-wide tile would
// waste half of its columns on them)
template <bool RELU, bool RESID, int BN>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const float* __restrict__ X, const float* __restrict__ Wf,
                                                         const float* __restrict__ bias, const float* __restrict__ resid,
                                                         float* __restrict__ Y, ConvDesc d, int M, int K, int Kpad) {
    constexpr int NG = BN / 64;                         // column groups of 4 per thread (tx*4 + 64*g)
    __shared__ __align__(16) float As[16][128 + 4];
    __shared__ __align__(16) float Bs[16][BN + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * BN;
    // the two tile rows this thread gathers: r = tid / 4 and 64 + tid / 4
    const float* xb[2]; int iy0[2], ix0[2]; bool rv[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int m = m0 + it * 64 + tid / 4;
        rv[it] = m < M;
        const int mm = rv[it] ? m : 0;
        const int n = mm / (d.Hout * d.Wout), rem = mm - n * d.Hout * d.Wout, oy = rem / d.Wout, ox = rem - oy * d.Wout;
        xb[it] = X + (long long)n * d.Hin * d.Win * d.Cin;
        iy0[it] = oy * d.stride - d.pad; ix0[it] = ox * d.stride - d.pad;
    }
    const int k4 = (tid % 4) * 4;
    float acc[8][4 * NG];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4 * NG; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < Kpad; k0 += 16) {
        const int k = k0 + k4;
        const int t = k / d.Cin, ci = k - t * d.Cin, ky = t / d.kw, kx = t - ky * d.kw;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int r = it * 64 + tid / 4;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            const int iy = iy0[it] + ky, ix = ix0[it] + kx;
            if (rv[it] && k < K && iy >= 0 && iy < d.Hin && ix >= 0 && ix < d.Win)
                a = *reinterpret_cast<const float4*>(xb[it] + ((long long)iy * d.Win + ix) * d.Cin + ci);
            As[k4 + 0][r] = a.x; As[k4 + 1][r] = a.y; As[k4 + 2][r] = a.z; As[k4 + 3][r] = a.w;
            if (it < NG) {
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + r < d.Cout) b = *reinterpret_cast<const float4*>(Wf + (long long)(n0 + r) * Kpad + k0 + k4);
                Bs[k4 + 0][r] = b.x; Bs[k4 + 1][r] = b.y; Bs[k4 + 2][r] = b.z; Bs[k4 + 3][r] = b.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[8], b[4 * NG];
            *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
#pragma unroll
            for (int g = 0; g < NG; ++g) *reinterpret_cast<float4*>(&b[4 * g]) = *reinterpret_cast<const float4*>(&Bs[kk][64 * g + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4 * NG; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    // fused epilogue: folded-BatchNorm bias, residual, ReLU; float4 stores (Cout % 4 == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= M) continue;
#pragma unroll
        for (int h = 0; h < NG; ++h) {
            const int c = n0 + h * 64 + tx * 4;
            if (c >= d.Cout) continue;
            const float4 bv = *reinterpret_cast<const float4*>(bias + c);
            float4 v = make_float4(acc[i][h * 4] + bv.x, acc[i][h * 4 + 1] + bv.y, acc[i][h * 4 + 2] + bv.z, acc[i][h * 4 + 3] + bv.w);
            if (RESID) {
                const float4 rr = *reinterpret_cast<const float4*>(resid + (long long)m * d.Cout + c);
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            }
            if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            *reinterpret_cast<float4*>(Y + (long long)m * d.Cout + c) = v;
        }
    }
}

// flow [N,H,W,2] -> NHWC4 [N,H,W,4] (channels 2, 3 zero: the reference appends ONE zero channel, the 4th only aligns loads)
static __global__ void rn_pad_flow_kernel(const float* __restrict__ flow, float* __restrict__ out, long long npix) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float2 f = *reinterpret_cast<const float2*>(flow + i * 2);
    *reinterpret_cast<float4*>(out + i * 4) = make_float4(f.x, f.y, 0.f, 0.f);
}

// MaxPool2d(kernel 3, stride 2, padding 1) on NHWC, 4 channels per thread
static __global__ void rn_maxpool_kernel(const float* __restrict__ X, float* __restrict__ Y, int N, int Hin, int Win, int C, int Hout, int Wout) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c4n = C / 4;
    if (i >= (long long)N * Hout * Wout * c4n) return;
    const int c4 = (int)(i % c4n);
    long long p = i / c4n;
    const int ox = (int)(p % Wout); p /= Wout;
    const int oy = (int)(p % Hout); const int n = (int)(p / Hout);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= Hin) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            if (ix < 0 || ix >= Win) continue;
            const float4 v = *reinterpret_cast<const float4*>(X + (((long long)n * Hin + iy) * Win + ix) * C + c4 * 4);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    *reinterpret_cast<float4*>(Y + i * 4) = m;
}

// AdaptiveAvgPool2d(1) on NHWC [N,HW,C] -> rows of the fc GEMM operand [rows_pad, C]
static __global__ void rn_avgpool_kernel(const float* __restrict__ X, float* __restrict__ Y, int N, int HW, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C) return;
    const int n = i / C, c = i - n * C;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += X[((long long)n * HW + p) * C + c];
    Y[i] = s / (float)HW;
}

static __global__ void rn_half_to_float_kernel(const __half* __restrict__ x, float* __restrict__ y, long long n) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const uint2 v = *reinterpret_cast<const uint2*>(x + i);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
    *reinterpret_cast<float4*>(y + i) = make_float4(a.x, a.y, b.x, b.y);
}

struct RnEpiFc { float* out; int ld; const float* bias; int n_rows;
    __device__ __forceinline__ void operator()(int row, int col, float acc) const { if (row < n_rows) out[(long long)row * ld + col] = acc + bias[col]; } };

struct RnBuf {
    float* p = nullptr; size_t n = 0;
    int alloc(size_t cnt) { release(); EG_CUDA(cudaMalloc(&p, cnt * sizeof(float))); EG_CUDA(cudaMemset(p, 0, cnt * sizeof(float))); n = cnt; return 0; }
    int upload(const std::vector<float>& h) { if (alloc(h.size())) return 1; EG_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice)); return 0; }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~RnBuf() { release(); }
};

struct RnConv {
    ConvDesc d; int K, Kpad; RnBuf w, b;
    // tensor-core path: fp16 weights [Cout, K64] (K padded to 64; the stem uses k = tap * 4 + channel, K64 = 256), their TMA map
    __half* w16 = nullptr; CUtensorMap map; int K64 = 0, BN = 64;
    ~RnConv() { if (w16) cudaFree(w16); }
};

}  // namespace egoego

using namespace egoego;

struct egoego_resnet_ctx {
    int device = 0, out_dim = 512, chunk = 64;      // frames per pass: 4 x 205 MB of activation buffers, >= 100 CTAs in the deepest layers
    std::map<std::string, std::vector<float>> staged;
    bool committed = false;
    std::vector<std::unique_ptr<RnConv>> convs;      // conv1, then per block: conv1, conv2, (downsample)
    RnBuf fc_w, fc_b, buf[4], in4, pooled;
    int64_t launches = 0;
    bool tc = true;                                  // tensor-core engine (EGOEGO_RESNET=simt: fp32 CUDA cores)
    __half* a16[3] = {nullptr, nullptr, nullptr};    // fp16 NHWC activations (conv inputs)
    __half* col16 = nullptr;                         // stem im2col [chunk * 112 * 112, 256]
    float* r32[3] = {nullptr, nullptr, nullptr};     // fp32 residual stream (block outputs / downsample outputs)
    ~egoego_resnet_ctx() { for (auto* p : a16) if (p) cudaFree(p); for (auto* p : r32) if (p) cudaFree(p); if (col16) cudaFree(col16); }
};

static const int RN_PLANES[4] = {64, 128, 256, 512};

// conv weight [Cout,Cin,kh,kw] + BatchNorm (gamma, beta, running mean / var, eps 1e-5) -> [Cout, Kpad] in (ky,kx,ci) order + bias
static int rn_fold(egoego_resnet_ctx* c, const std::string& conv, const std::string& bn, int Cin, int Cin_pad, int Cout, int kh, int kw,
                   int stride, int pad, int Hin, RnConv* out) {
    auto need = [&](const std::string& k) -> const std::vector<float>* { auto it = c->staged.find(k); return it == c->staged.end() ? nullptr : &it->second; };
    const std::vector<float>*w = need(conv + ".weight"), *g = need(bn + ".weight"), *be = need(bn + ".bias"), *mu = need(bn + ".running_mean"), *var = need(bn + ".running_var");
    EG_CHECK(w && g && be && mu && var, "missing tensor for " + conv + " / " + bn);
    EG_CHECK((int64_t)w->size() == (int64_t)Cout * Cin * kh * kw && (int)g->size() == Cout, "bad tensor size for " + conv);
    out->d = ConvDesc{Cin_pad, Cout, kh, kw, stride, pad, Hin, Hin, (Hin + 2 * pad - kh) / stride + 1, (Hin + 2 * pad - kh) / stride + 1};
    out->K = kh * kw * Cin_pad;
    out->Kpad = ((out->K + 15) / 16) * 16;
    std::vector<float> wf((size_t)Cout * out->Kpad, 0.f), bf(Cout);
    for (int o = 0; o < Cout; ++o) {
        const double s = (double)(*g)[o] / std::sqrt((double)(*var)[o] + 1e-5);
        bf[o] = (float)((double)(*be)[o] - (double)(*mu)[o] * s);
        for (int ci = 0; ci < Cin; ++ci)
            for (int ky = 0; ky < kh; ++ky)
                for (int kx = 0; kx < kw; ++kx)
                    wf[(size_t)o * out->Kpad + (ky * kw + kx) * Cin_pad + ci] = (float)((double)(*w)[(((size_t)o * Cin + ci) * kh + ky) * kw + kx] * s);
    }
    if (out->w.upload(wf) || out->b.upload(bf)) return 1;
    if (c->tc) {                                     // fp16 copy with K padded to 64 (same (ky, kx, ci) order) + TMA map of BN-row boxes
        out->K64 = ((out->K + 63) / 64) * 64;
        out->BN = Cout >= 256 ? 256 : Cout;
        std::vector<__half> wh((size_t)Cout * out->K64, __float2half(0.f));
        for (int o = 0; o < Cout; ++o)
            for (int k = 0; k < out->K; ++k) wh[(size_t)o * out->K64 + k] = __float2half_rn(wf[(size_t)o * out->Kpad + k]);
        if (out->w16) { cudaFree(out->w16); out->w16 = nullptr; }
        EG_CUDA(cudaMalloc(&out->w16, wh.size() * 2));
        EG_CUDA(cudaMemcpy(out->w16, wh.data(), wh.size() * 2, cudaMemcpyHostToDevice));
        if (tc_make_map16(&out->map, out->w16, Cout, out->K64, out->BN)) return 1;
    }
    return 0;
}

// tensor-core convolution launch: x16 NHWC fp16 (or the stem's im2col matrix with `dense` = its row length), optional fp32 identity,
// outputs fp32 and / or fp16
static int rn_conv_tc(egoego_resnet_ctx* c, const RnConv& cv, const __half* x16, int dense_k, const float* resid, float* y32, __half* y16,
                      bool relu, int N, cudaStream_t s) {
    const int M = N * cv.d.Hout * cv.d.Wout;
    ConvTcDesc d{cv.d.Cin, cv.d.Cout, cv.d.kh, cv.d.kw, cv.d.stride, cv.d.pad, cv.d.Hin, cv.d.Win, cv.d.Hout, cv.d.Wout};
    if (dense_k) d = ConvTcDesc{dense_k, cv.d.Cout, 1, 1, 1, 0, cv.d.Hout, cv.d.Wout, cv.d.Hout, cv.d.Wout};
    EG_CHECK(d.Cin % 64 == 0 && cv.K64 % 64 == 0, "tensor-core convolution needs Cin % 64 == 0");
    ConvTcEpi e{{}, cv.b.p, resid, y32, y16, cv.d.Cout, M, relu ? 1 : 0};
    const int tiles = ((M + 127) / 128) * ((cv.d.Cout + cv.BN - 1) / cv.BN);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    const int grid = tiles < sms ? tiles : sms;
    static bool attr[64] = {};
    if (!attr[c->device & 63]) {
        EG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<64>::SMEM_BYTES));
        EG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<128>::SMEM_BYTES));
        EG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvTcCfg<256>::SMEM_BYTES));
        attr[c->device & 63] = true;
    }
    const int kb = cv.K64 / 64;
    if (cv.BN == 64)       conv_tc_kernel<64><<<grid, CONV_TC_THREADS, ConvTcCfg<64>::SMEM_BYTES, s>>>(x16, cv.map, d, M, kb, e);
    else if (cv.BN == 128) conv_tc_kernel<128><<<grid, CONV_TC_THREADS, ConvTcCfg<128>::SMEM_BYTES, s>>>(x16, cv.map, d, M, kb, e);
    else                   conv_tc_kernel<256><<<grid, CONV_TC_THREADS, ConvTcCfg<256>::SMEM_BYTES, s>>>(x16, cv.map, d, M, kb, e);
    c->launches++;
    return 0;
}

template <bool RELU, bool RESID>
static int rn_conv(egoego_resnet_ctx* c, const RnConv& cv, const float* x, const float* resid, float* y, int N, cudaStream_t s) {
    const int M = N * cv.d.Hout * cv.d.Wout;
    if (cv.d.Cout <= 64) {
        dim3 grid((cv.d.Cout + 63) / 64, (M + 127) / 128);
        conv_igemm_kernel<RELU, RESID, 64><<<grid, 256, 0, s>>>(x, cv.w.p, cv.b.p, resid, y, cv.d, M, cv.K, cv.Kpad);
    } else {
        dim3 grid((cv.d.Cout + 127) / 128, (M + 127) / 128);
        conv_igemm_kernel<RELU, RESID, 128><<<grid, 256, 0, s>>>(x, cv.w.p, cv.b.p, resid, y, cv.d, M, cv.K, cv.Kpad);
    }
    c->launches++;
    return 0;
}

extern "C" {

int egoego_resnet18_create(int device, int out_dim, egoego_resnet* out) {
    EG_CHECK(out, "null argument");
    EG_CHECK(out_dim >= 4 && out_dim % 4 == 0 && out_dim <= 4096, "out_dim must be a multiple of 4");
    int ndev = 0;
    EG_CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0, "no CUDA device: libegoego_b200 has no CPU fallback");
    EG_CHECK(device >= 0 && device < ndev, "bad device ordinal");
    egoego_resnet_ctx* c = new egoego_resnet_ctx();
    c->device = device; c->out_dim = out_dim;
    { const char* e = getenv("EGOEGO_RESNET"); c->tc = !(e && strcmp(e, "simt") == 0); }
    if (c->tc) c->chunk = 160;                       // whole demo sequences (139 frames) in one pass: ~2.3 GB of activation buffers
    *out = c;
    return 0;
}

void egoego_resnet18_destroy(egoego_resnet c) { delete c; }

int egoego_resnet18_set_tensor(egoego_resnet c, const char* name, const float* host, int64_t numel) {
    EG_CHECK(c && name && host && numel > 0, "bad argument");
    c->staged[name].assign(host, host + numel);
    c->committed = false;
    return 0;
}

int egoego_resnet18_commit(egoego_resnet c) {
    EG_CHECK(c, "null handle");
    EG_CUDA(cudaSetDevice(c->device));
    c->convs.clear();
    {   // stem: 7x7 stride 2 pad 3, 3 input channels padded to 4, on 224 x 224
        std::unique_ptr<RnConv> cv(new RnConv());
        if (rn_fold(c, "conv1", "bn1", 3, 4, 64, 7, 7, 2, 3, 224, cv.get())) return 1;
        c->convs.push_back(std::move(cv));
    }
    int H = 56, inpl = 64;
    for (int L = 0; L < 4; ++L) {
        for (int blk = 0; blk < 2; ++blk) {
            const int planes = RN_PLANES[L], stride = (L > 0 && blk == 0) ? 2 : 1;
            const std::string p = "layer" + std::to_string(L + 1) + "." + std::to_string(blk) + ".";
            std::unique_ptr<RnConv> c1(new RnConv()), c2(new RnConv());
            if (rn_fold(c, p + "conv1", p + "bn1", inpl, inpl, planes, 3, 3, stride, 1, H, c1.get())) return 1;
            const int Ho = c1->d.Hout;
            if (rn_fold(c, p + "conv2", p + "bn2", planes, planes, planes, 3, 3, 1, 1, Ho, c2.get())) return 1;
            c->convs.push_back(std::move(c1)); c->convs.push_back(std::move(c2));
            if (stride != 1 || inpl != planes) {
                std::unique_ptr<RnConv> ds(new RnConv());
                if (rn_fold(c, p + "downsample.0", p + "downsample.1", inpl, inpl, planes, 1, 1, stride, 0, H, ds.get())) return 1;
                c->convs.push_back(std::move(ds));
            }
            H = Ho; inpl = planes;
        }
    }
    auto fw = c->staged.find("fc.weight"), fb = c->staged.find("fc.bias");
    EG_CHECK(fw != c->staged.end() && fb != c->staged.end(), "missing tensor: fc.weight / fc.bias");
    EG_CHECK((int64_t)fw->second.size() == (int64_t)c->out_dim * 512 && (int)fb->second.size() == c->out_dim, "bad fc size");
    if (c->fc_w.upload(fw->second) || c->fc_b.upload(fb->second)) return 1;
    const size_t act = (size_t)c->chunk * 112 * 112 * 64;                 // largest activation of a chunk (stem output)
    if (c->tc) {
        const size_t act56 = (size_t)c->chunk * 56 * 56 * 64;              // largest activation after the max-pool
        for (auto*& p : c->a16) { if (p) cudaFree(p); EG_CUDA(cudaMalloc(&p, act * 2)); }
        for (auto*& p : c->r32) { if (p) cudaFree(p); EG_CUDA(cudaMalloc(&p, act56 * 4)); }
        if (c->col16) cudaFree(c->col16);
        EG_CUDA(cudaMalloc(&c->col16, (size_t)c->chunk * 112 * 112 * 256 * 2));
        if (c->pooled.alloc((size_t)256 * 512)) return 1;
    } else {
        for (auto& b : c->buf) if (b.alloc(act)) return 1;
        if (c->in4.alloc((size_t)c->chunk * 224 * 224 * 4) || c->pooled.alloc((size_t)128 * 512)) return 1;
    }
    c->committed = true;
    return 0;
}

int egoego_resnet18_forward(egoego_resnet c, const float* flow, int N, float* feats, void* stream_v) {
    EG_CHECK(c && flow && feats && N >= 1, "bad argument");
    EG_CHECK(c->committed, "egoego_resnet18_commit has not been called");
    EG_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream_v;
    for (int f0 = 0; c->tc && f0 < N; f0 += c->chunk) {
        const int n = std::min(c->chunk, N - f0);
        size_t ci = 0;
        // stem: explicit im2col (3 input channels) -> tensor-core GEMM + BN + ReLU -> fp16 [n,112,112,64] -> max-pool -> [n,56,56,64]
        const long long M1 = (long long)n * 112 * 112;
        rn_stem_im2col_kernel<<<(unsigned)((M1 * 64 + 255) / 256), 256, 0, s>>>(flow + (long long)f0 * 224 * 224 * 2, c->col16, M1);
        if (rn_conv_tc(c, *c->convs[ci++], c->col16, 256, nullptr, nullptr, c->a16[1], true, n, s)) return 1;
        {
            const long long tot = (long long)n * 56 * 56 * 8;
            rn_maxpool16_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(c->a16[1], c->a16[0], n, 112, 112, 64, 56, 56);
        }
        // residual stream: x16 (conv input, fp16) + x32 (identity, fp32; the max-pool output is exactly representable in fp16, so the
        // first block's identity is read from x16 through a one-off conversion below)
        __half *x16 = c->a16[0], *y16 = c->a16[1], *t16 = c->a16[2];
        float *x32 = c->r32[0], *y32 = c->r32[1], *d32 = c->r32[2];
        {
            const long long tot = (long long)n * 56 * 56 * 64;
            rn_half_to_float_kernel<<<(unsigned)((tot / 4 + 255) / 256), 256, 0, s>>>(x16, x32, tot);
        }
        int inpl = 64;
        for (int L = 0; L < 4; ++L)
            for (int blk = 0; blk < 2; ++blk) {
                const int planes = RN_PLANES[L], stride = (L > 0 && blk == 0) ? 2 : 1;
                const RnConv& c1 = *c->convs[ci++];
                const RnConv& c2 = *c->convs[ci++];
                const float* identity = x32;
                if (rn_conv_tc(c, c1, x16, 0, nullptr, nullptr, t16, true, n, s)) return 1;              // conv1 + BN + ReLU -> fp16
                if (stride != 1 || inpl != planes) {
                    if (rn_conv_tc(c, *c->convs[ci++], x16, 0, nullptr, d32, nullptr, false, n, s)) return 1;   // downsample conv + BN -> fp32
                    identity = d32;
                }
                if (rn_conv_tc(c, c2, t16, 0, identity, y32, y16, true, n, s)) return 1;                 // conv2 + BN + identity + ReLU -> fp32 + fp16
                std::swap(x16, y16); std::swap(x32, y32);
                inpl = planes;
            }
        rn_avgpool_kernel<<<(n * 512 + 255) / 256, 256, 0, s>>>(x32, c->pooled.p, n, 49, 512);
        RnEpiFc e{feats + (long long)f0 * c->out_dim, c->out_dim, c->fc_b.p, n};
        sgemm_tn_kernel<<<dim3((c->out_dim + 127) / 128, (n + 127) / 128), 256, 0, s>>>(c->pooled.p, 512, c->fc_w.p, 512, c->out_dim, 512, e);
        c->launches += 5;
    }
    for (int f0 = 0; !c->tc && f0 < N; f0 += c->chunk) {
        const int n = std::min(c->chunk, N - f0);
        const long long npix = (long long)n * 224 * 224;
        rn_pad_flow_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, s>>>(flow + (long long)f0 * 224 * 224 * 2, c->in4.p, npix);
        float *x = c->buf[0].p, *y = c->buf[1].p, *t1 = c->buf[2].p, *t2 = c->buf[3].p;
        size_t ci = 0;
        if (rn_conv<true, false>(c, *c->convs[ci++], c->in4.p, nullptr, y, n, s)) return 1;          // stem conv + BN + ReLU -> [n,112,112,64]
        {
            const long long tot = (long long)n * 56 * 56 * 16;
            rn_maxpool_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(y, x, n, 112, 112, 64, 56, 56);
        }
        int inpl = 64;
        for (int L = 0; L < 4; ++L)
            for (int blk = 0; blk < 2; ++blk) {
                const int planes = RN_PLANES[L], stride = (L > 0 && blk == 0) ? 2 : 1;
                const RnConv& c1 = *c->convs[ci++];
                const RnConv& c2 = *c->convs[ci++];
                const float* identity = x;
                if (rn_conv<true, false>(c, c1, x, nullptr, t1, n, s)) return 1;                     // conv1 + BN + ReLU
                if (stride != 1 || inpl != planes) {
                    if (rn_conv<false, false>(c, *c->convs[ci++], x, nullptr, t2, n, s)) return 1;   // downsample conv + BN
                    identity = t2;
                }
                if (rn_conv<true, true>(c, c2, t1, identity, y, n, s)) return 1;                     // conv2 + BN + identity + ReLU
                std::swap(x, y);
                inpl = planes;
            }
        rn_avgpool_kernel<<<(n * 512 + 255) / 256, 256, 0, s>>>(x, c->pooled.p, n, 49, 512);
        RnEpiFc e{feats + (long long)f0 * c->out_dim, c->out_dim, c->fc_b.p, n};
        sgemm_tn_kernel<<<dim3((c->out_dim + 127) / 128, 1), 256, 0, s>>>(c->pooled.p, 512, c->fc_w.p, 512, c->out_dim, 512, e);
        c->launches += 4;
    }
    EG_CUDA(cudaGetLastError());
    return 0;
}

int64_t egoego_resnet18_launch_count(egoego_resnet c) { return c ? c->launches : -1; }

}  // extern "C"
