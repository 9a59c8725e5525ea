// Shared declarations for libegoego_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <cstdlib>
#include <string>
#include <vector>
#include <map>

#include "../../include/egoego_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libegoego_b200 is written for sm_100a (Blackwell B200) only"
#endif

namespace egoego {

constexpr int LP = 128;         // padded tokens per window: 1 time token + up to 120 frames + pad
constexpr int NJ = 22;          // SMPL body joints
constexpr int HEAD_IDX = 15;

void set_error(const std::string& msg);

#define EG_CHECK(cond, msg)                                                      \
    do { if (!(cond)) { ::egoego::set_error(std::string(msg)); return 1; } } while (0)

#define EG_CUDA(call)                                                            \
    do { cudaError_t _e = (call);                                                \
         if (_e != cudaSuccess) {                                                \
             ::egoego::set_error(std::string(#call) + ": " + cudaGetErrorString(_e) + \
                                 " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
             return 1; } } while (0)

// Launch configuration with optional thread-block cluster and programmatic dependent launch (PDL, see tc_ptx.cuh).
// PDL is on for every kernel of the sampling step unless EGOEGO_PDL=0.
inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("EGOEGO_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
struct LaunchCfg {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[2];
    LaunchCfg(unsigned grid, unsigned block, size_t smem, cudaStream_t s, int cluster = 1, bool pdl = true) {
        cfg = cudaLaunchConfig_t{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        int n = 0;
        if (cluster > 1) {
            at[n].id = cudaLaunchAttributeClusterDimension;
            at[n].val.clusterDim.x = cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
            ++n;
        }
        if (pdl && pdl_enabled()) {
            at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[n].val.programmaticStreamSerializationAllowed = 1;
            ++n;
        }
        cfg.attrs = at; cfg.numAttrs = n;
    }
    LaunchCfg(const LaunchCfg&) = delete;
};

// Debug timeline (tools/stream_timeline.py): when switched on, every CTA of the streamed QKV projection (kernel 0) and of the
// attention kernel (kernel 1) records %globaltimer after its start-up waits and before it exits; the last launch wins.  Per
// translation unit; the kernels and the accessor (tc_debug_timeline) live in engine_tc.cu.
constexpr int TIMELINE_CTAS = 160;
static __device__ unsigned long long g_timeline[2][TIMELINE_CTAS][2];
static __device__ int g_timeline_on = 0;
__device__ __forceinline__ void timeline_mark(int kernel, int which) {
    if (g_timeline_on && blockIdx.x < TIMELINE_CTAS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_timeline[kernel][blockIdx.x][which] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// Which diffusion step a window is at.  Loop mode: t = t_start - *d_step (device counter advanced
// once per step so one captured CUDA graph serves every step).  Step-API mode: explicit int64 array.
// ---------------------------------------------------------------------------------------------
struct TSrc {
    const long long* t_arr;   // [B] or nullptr
    const int*       d_step;  // device counter or nullptr
    int              t_start;
    int              t_max;   // timesteps - 1: a caller-supplied t outside [0, t_max] is clamped (tables have `timesteps` rows;
                              // the reference fails with a device-side assert there, this library must not fault the GPU)
    __device__ __forceinline__ int get(int w) const {
        const int t = t_arr ? (int)t_arr[w] : t_start - (d_step ? *d_step : 0);
        return min(max(t, 0), t_max);
    }
};

// Gaussian source: tape (explicit) or Philox4x32-10 + Box-Muller.
struct NoiseSrc {
    const float* tape;          // base of the tape (or of one explicit draw), nullptr => philox
    long long    draw_stride;   // elements between consecutive draws in the tape
    const int*   d_step;        // device step counter added to draw_static (nullable)
    int          draw_static;
    unsigned long long seed;
    unsigned long long window_offset;
    __device__ __forceinline__ int draw() const { return draw_static + (d_step ? *d_step : 0); }
};

// One DDPM update (p_sample after the denoiser): arguments of ddpm_update_kernel and of the linear_out epilogue that fuses it.
struct DdpmArgs {
    const float* model_out;  // [B,T,D]
    const float* x;          // [B,T,D]
    float* x_out;            // [B,T,D] (may alias x)
    const float* coef1; const float* coef2; const float* logvar;
    const float* sqrt_recip; const float* sqrt_recipm1;
    int objective;           // 0 pred_noise, 1 pred_x0
    int clip;
    const float* inpaint; int inpaint_len;
    float* stage_f32; int stage_ld;                                  // SIMT engine A operand (nullable)
    __nv_bfloat16* stage_hi; __nv_bfloat16* stage_lo; int stage_ld16; // tensor engine A operand (nullable)
    __half* stage_h16;                                                // fp16 plane for FMT_HALF steps (nullable)
    int stage_mode;          // tensor engine planes to write: 0 = fp16 hi/lo pair only, 1 = single fp16 plane only, 2 = all (next step's format unknown)
    TSrc ts; NoiseSrc ns;
    int B, T, D;
    int* advance; unsigned* done;   // ddpm_update_kernel in the sampling loop: step counter to advance once all blocks are done + its block counter (nullable)
};

__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
    unsigned long long p = (unsigned long long)a * b;
    *hi = (uint32_t)(p >> 32);
    return (uint32_t)p;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, hi1;
        uint32_t lo0 = mulhilo32(M0, ctr.x, &hi0);
        uint32_t lo1 = mulhilo32(M1, ctr.z, &hi1);
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

// 4 standard normals for (window, draw, quad index) -- quad = element_index / 4 within the window.
__device__ __forceinline__ float4 philox_normal4(unsigned long long seed, unsigned long long window,
                                                 uint32_t draw, uint32_t quad) {
    uint4 r = philox4x32_10(make_uint4(quad, draw, (uint32_t)window, (uint32_t)(window >> 32)),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float k = 2.3283064365386963e-10f;  // 2^-32
    float u0 = ((float)r.x + 0.5f) * k, u1 = ((float)r.y + 0.5f) * k;
    float u2 = ((float)r.z + 0.5f) * k, u3 = ((float)r.w + 0.5f) * k;
    u0 = fminf(u0, 0.99999994f); u2 = fminf(u2, 0.99999994f);
    // Box-Muller on the SFU: lg2.approx / sqrt.approx / sin.approx / cos.approx (abs error ~1e-6 on a unit normal -- the
    // generator defines the noise stream, there is no external bit pattern to match; angles kept in [-pi, pi))
    const float kNeg2Ln2 = -1.3862943611198906f, kTwoPi = 6.283185307179586f;
    float r0, r1;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(kNeg2Ln2 * __log2f(u0)));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(kNeg2Ln2 * __log2f(u2)));
    float s0, c0, s1, c1;
    __sincosf(kTwoPi * (u1 - 0.5f), &s0, &c0);
    __sincosf(kTwoPi * (u3 - 0.5f), &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// One normal for flat element e of window w (elements are grouped in quads).
__device__ __forceinline__ float noise_at(const NoiseSrc& ns, int draw, int w, long long elems_per_window, int e) {
    if (ns.tape) return ns.tape[(long long)draw * ns.draw_stride + (long long)w * elems_per_window + e];
    float4 n = philox_normal4(ns.seed, ns.window_offset + (unsigned long long)w, (uint32_t)draw, (uint32_t)(e >> 2));
    int r = e & 3;
    return r == 0 ? n.x : (r == 1 ? n.y : (r == 2 ? n.z : n.w));
}

// ---------------------------------------------------------------------------------------------
// Dropout of the training step (nn.Dropout(0.1) on the attention probabilities, the fc output and the FFN output of every
// DecoderLayer: egoego/model/transformer_module.py:53,59,84,92,105,113).  torch's own dropout stream cannot be reproduced, so
// the masks are a counter-based function of (seed, stream, element index) that the CPU oracle restates bit for bit
// (oracle/training.py: dropout_keep): Philox4x32-10 with counter (index / 4, index >> 34, stream, 'DROP') and key = seed; element
// i keeps its value iff component i % 4 of the output is below thresh = floor((1 - p) 2^32).  Kept values are scaled by 1 / (1 - p).
//   stream = 4 * layer + site;  site 0: attention probabilities, index ((b H + h) 128 + query) 128 + key
//                               site 1 / 2: fc / FFN output, index (b 128 + token) 512 + channel      (token 0 = time token)
// ---------------------------------------------------------------------------------------------
struct DropCfg {
    unsigned long long seed;
    uint32_t thresh;      // keep iff philox word < thresh
    float scale;          // 1 / (1 - p)
    int on;               // 0: identity (eval mode)
    const unsigned long long* seed_dev;   // non-null: the seed lives in device memory (a captured training step is replayed with a new seed)
};
__device__ __forceinline__ uint4 drop_words(const DropCfg& d, uint32_t stream, unsigned long long quad) {
    const unsigned long long sd = d.seed_dev ? *d.seed_dev : d.seed;
    return philox4x32_10(make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), stream, 0x44524f50u),
                         make_uint2((uint32_t)sd, (uint32_t)(sd >> 32)));
}
__device__ __forceinline__ float drop_factor(const DropCfg& d, uint32_t stream, unsigned long long idx) {
    if (!d.on) return 1.0f;
    const uint4 r = drop_words(d, stream, idx >> 2);
    const uint32_t w = (idx & 3) == 0 ? r.x : ((idx & 3) == 1 ? r.y : ((idx & 3) == 2 ? r.z : r.w));
    return w < d.thresh ? d.scale : 0.0f;
}

// bf16 hi/lo split of an fp32 value: v ~= hi + lo with |v - hi - lo| <= 2^-17 |v| over the whole fp32 exponent range
// (training-step products: gradients are far below the fp16 range).
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// fp16 hi/lo split of an fp32 value (sampling path, FMT_SPLIT): hi = fp16(v), lo = fp16(v - hi), both kept as raw 16-bit
// patterns in the plane storage type.  |v - hi - lo| <= max(2^-23 |v|, 2^-25): 22 significand bits where lo is a normal
// fp16 (|v| >= 2^-3), an absolute 3e-8 below (lo subnormal) -- 2^6 (activations) to 2^3 (weights ~ 0.03) tighter than the
// bf16 pair at the same tensor-core cost.  Operands of the sampler are O(1) (the single-plane fp16 format of the early
// steps relies on the same range), so the fp16 exponent range is not a constraint here.
__device__ __forceinline__ void split_f16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    hi = __ushort_as_bfloat16(__half_as_ushort(h));
    lo = __ushort_as_bfloat16(__half_as_ushort(l));
}

}  // namespace egoego
