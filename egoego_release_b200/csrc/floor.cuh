// Floor height and contact labels of a motion sequence on the device: what the evaluation scripts compute
// on the host with numpy + sklearn before every metrics call (eval_stage2.py:131,189, eval_egoego.py:331,395).
//
// Follows determine_floor_height_and_contacts / detect_joint_contact, utils/data_utils/process_amass_dataset.py:160-328
// (constants :52-61, joint ids body_model/utils.py:5-8):
//   1. toe speeds (frame differences, last value repeated) -> "static" toe samples (speed < 0.005), left foot then right foot;
//   2. sklearn DBSCAN(eps = 0.005, min_samples = 3) on their heights (1-D) -- restated from its published algorithm, see
//      oracle/floor.py: core = >= 3 samples within eps (itself included, distances in fp64), clusters grown from core samples
//      in index order, a border sample joins the first cluster that reaches it, the rest is noise (label -1);
//   3. median height of EVERY label (noise included), floor = smallest median, returned minus 0.01;
//   4. terrain heuristic (discard flag) from the clusters' median root heights over their unique frames;
//   5. contact flags of feet / toes / hands / knees from speed and height above the floor.
// One block per sequence.  The sample count is at most 2 T, so every set operation is a brute-force O(n^2) pass over shared
// memory (n = 240 at the reference's T = 120): ranks replace sorting, so the medians are exactly numpy's.
#pragma once
#include "common.cuh"

namespace egoego {

constexpr int FLOOR_MAX_T = 2048;
constexpr int FLOOR_THREADS = 256;
__host__ __device__ inline size_t floor_smem_bytes(int T) {
    const size_t n = 2 * (size_t)T;
    return (2 * (size_t)T + 4 * n + 8 * (n + 1) + 16) * 4;
}

__device__ __forceinline__ float speed3(const float* a, const float* b) {   // |b - a| exactly as numpy evaluates it in fp32
    const float dx = __fsub_rn(b[0], a[0]), dy = __fsub_rn(b[1], a[1]), dz = __fsub_rn(b[2], a[2]);
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}
__device__ __forceinline__ float joint_speed(const float* seq, int T, int t, int joint) {
    const int a = t < T - 1 ? t : T - 2;                     // np.append(v, v[-1])
    return speed3(seq + ((long long)a * NJ + joint) * 3, seq + ((long long)(a + 1) * NJ + joint) * 3);
}
__device__ __forceinline__ bool lex_less(float xa, int ia, float xb, int ib) { return xa < xb || (xa == xb && ia < ib); }

__global__ void __launch_bounds__(FLOOR_THREADS)
floor_contacts_kernel(const float* __restrict__ jpos /* [B,T,22,3] */, int T, int fps, float* __restrict__ floor_out /* [B] */,
                      float* __restrict__ contacts /* [B,T,22] or null */, int* __restrict__ discard_out /* [B] or null */) {
    constexpr int HIPS = 0, L_LEG = 4, R_LEG = 5, L_FOOT = 7, R_FOOT = 8, L_TOE = 10, R_TOE = 11, L_HAND = 20, R_HAND = 21;
    constexpr float VEL = 0.005f, OFFSET = 0.01f, TOE_H = 0.04f, ANKLE_H = 0.08f, TERRAIN_H = 0.04f, ROOT_H = 0.04f;
    constexpr double EPS = 0.005;
    extern __shared__ float fsm[];
    const int n_max = 2 * T;
    float* lv = fsm;                      // [T] left toe speed
    float* rv = lv + T;                   // [T] right toe speed
    float* x = rv + T;                    // [n] static heights
    int* idx = reinterpret_cast<int*>(x + n_max);      // [n] frame of the sample
    int* label = idx + n_max;             // [n]
    int* aux = label + n_max;             // [n] core flag, later component start flag
    float* lo_mid = reinterpret_cast<float*>(aux + n_max);   // [n+1] per label slot (slot 0 = noise, slot 1 + c = cluster c)
    float* hi_mid = lo_mid + n_max + 1;
    float* rlo = hi_mid + n_max + 1;      // root-height middles
    float* rhi = rlo + n_max + 1;
    int* cnt = reinterpret_cast<int*>(rhi + n_max + 1);
    int* rcnt = cnt + n_max + 1;
    int* first = rcnt + n_max + 1;        // smallest sample index among the cluster's core samples (discovery order)
    int* used = first + n_max + 1;        // label slot occurs
    int* misc = used + n_max + 1;         // [0] n_left, [1] n, [2] n_clusters
    const float* seq = jpos + (long long)blockIdx.x * T * NJ * 3;
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int t = tid; t < T; t += nt) { lv[t] = joint_speed(seq, T, t, L_TOE); rv[t] = joint_speed(seq, T, t, R_TOE); }
    __syncthreads();
    // ---- compaction in the reference's order: left statics by frame, then right statics by frame (one warp, ballot scan) ----
    if (tid < 32) {
        int base = 0;
        for (int side = 0; side < 2; ++side) {
            const float* v = side ? rv : lv;
            const int toe = side ? R_TOE : L_TOE;
            for (int t0 = 0; t0 < T; t0 += 32) {
                const int t = t0 + tid;
                const bool st = t < T && v[t] < VEL;
                const unsigned m = __ballot_sync(0xffffffffu, st);
                if (st) { const int k = base + __popc(m & ((1u << tid) - 1u)); x[k] = seq[((long long)t * NJ + toe) * 3 + 2]; idx[k] = t; }
                base += __popc(m);
            }
            if (side == 0 && tid == 0) misc[0] = base;
        }
        if (tid == 0) misc[1] = base;
    }
    __syncthreads();
    const int n = misc[1], n_left = misc[0];
    float floor_h = 0.f;
    int discard = 0;
    if (n > 0) {
        // ---- core samples ----
        for (int k = tid; k < n; k += nt) {
            const double xk = (double)x[k];
            int c = 0;
            for (int j = 0; j < n; ++j) c += (fabs((double)x[j] - xk) <= EPS) ? 1 : 0;
            aux[k] = c >= 3 ? 1 : 0;
        }
        for (int k = tid; k <= n; k += nt) { first[k] = 0x7fffffff; cnt[k] = 0; rcnt[k] = 0; used[k] = 0; }
        __syncthreads();
        // ---- components of the core samples: in 1-D, sorted core samples chain while consecutive gaps are <= eps.  A core sample
        // starts a component iff no core sample precedes it (in (value, index) order) within eps; its component number is the
        // count of starts at or before it. ----
        for (int k = tid; k < n; k += nt) {
            int start = 0;
            if (aux[k]) {
                start = 1;
                const double xk = (double)x[k];
                for (int j = 0; j < n; ++j)
                    if (aux[j] && j != k && lex_less(x[j], j, x[k], k) && xk - (double)x[j] <= EPS) { start = 0; break; }
            }
            label[k] = start;                   // temporarily: start flag
        }
        __syncthreads();
        for (int k = tid; k < n; k += nt) {
            int comp = -1;
            if (aux[k]) {
                comp = 0;
                for (int j = 0; j < n; ++j) comp += (label[j] && (j == k || lex_less(x[j], j, x[k], k))) ? 1 : 0;
                comp -= 1;
                atomicMin(&first[1 + comp], k);
            }
            lo_mid[k] = __int_as_float(comp);   // park the component id (label[] still holds the start flags being read)
        }
        __syncthreads();
        // ---- labels: core -> its component; non-core within eps of a core -> the component discovered first; else noise ----
        for (int k = tid; k < n; k += nt) {
            int lab = -1;
            if (aux[k]) lab = __float_as_int(lo_mid[k]);
            else {
                const double xk = (double)x[k];
                int best_first = 0x7fffffff;
                for (int j = 0; j < n; ++j)
                    if (aux[j] && fabs((double)x[j] - xk) <= EPS) {
                        const int c = __float_as_int(lo_mid[j]);
                        if (first[1 + c] < best_first) { best_first = first[1 + c]; lab = c; }
                    }
            }
            rhi[k] = __int_as_float(lab);       // park again: lo_mid is still being read by other threads
        }
        __syncthreads();
        for (int k = tid; k < n; k += nt) { label[k] = __float_as_int(rhi[k]); used[1 + label[k]] = 1; }
        __syncthreads();
        // a frame appears twice in a label when both feet are static there: the right-foot sample is then a duplicate for the
        // root-height statistics (np.unique of the frame indices); aux[] (core flags, no longer needed) takes that flag
        for (int k = tid; k < n; k += nt) {
            int dup = 0;
            if (k >= n_left)
                for (int j = 0; j < n_left; ++j) if (idx[j] == idx[k] && label[j] == label[k]) { dup = 1; break; }
            aux[k] = dup;
        }
        __syncthreads();
        // ---- medians per label: the samples of rank (m-1)/2 and m/2 inside the label; root heights over the label's UNIQUE frames ----
        for (int k = tid; k < n; k += nt) {
            const int lab = label[k], f = idx[k];
            int m = 0, r = 0;
            for (int j = 0; j < n; ++j)
                if (label[j] == lab) { ++m; r += lex_less(x[j], j, x[k], k) ? 1 : 0; }
            if (r == (m - 1) / 2) lo_mid[1 + lab] = x[k];
            if (r == m / 2) hi_mid[1 + lab] = x[k];
            cnt[1 + lab] = m;
            if (!aux[k]) {
                const float rk = seq[((long long)f * NJ + HIPS) * 3 + 2];
                int rm = 0, rr = 0;
                for (int j = 0; j < n; ++j) {
                    if (label[j] != lab || aux[j]) continue;
                    const int g = idx[j];
                    ++rm;
                    rr += lex_less(seq[((long long)g * NJ + HIPS) * 3 + 2], g, rk, f) ? 1 : 0;
                }
                if (rr == (rm - 1) / 2) rlo[1 + lab] = rk;
                if (rr == rm / 2) rhi[1 + lab] = rk;
                rcnt[1 + lab] = rm;
            }
        }
        __syncthreads();
        // ---- floor = smallest median over the labels in np.unique order (-1 first); terrain heuristic ----
        float min_med = INFINITY, min_root = INFINITY;
        for (int s = 0; s <= n; ++s) {
            if (!used[s]) continue;
            const float med = __fmul_rn(__fadd_rn(lo_mid[s], hi_mid[s]), 0.5f);
            if (med < min_med) { min_med = med; min_root = __fmul_rn(__fadd_rn(rlo[s], rhi[s]), 0.5f); }
        }
        floor_h = min_med;
        const int size_thresh = (int)(0.25 * (double)fps);
        for (int s = 0; s <= n && !discard; ++s) {
            if (!used[s]) continue;
            const float med = __fmul_rn(__fadd_rn(lo_mid[s], hi_mid[s]), 0.5f), rmed = __fmul_rn(__fadd_rn(rlo[s], rhi[s]), 0.5f);
            if (rmed > __fadd_rn(min_root, ROOT_H) && med > __fadd_rn(min_med, TERRAIN_H) && cnt[s] > size_thresh) discard = 1;
        }
    }
    if (tid == 0) {
        floor_out[blockIdx.x] = n > 0 ? __fsub_rn(floor_h, OFFSET) : 0.f;
        if (discard_out) discard_out[blockIdx.x] = discard;
    }
    if (contacts) {
        float* c = contacts + (long long)blockIdx.x * T * NJ;
        for (int i = tid; i < T * NJ; i += nt) {
            const int t = i / NJ, j = i % NJ;
            float v = 0.f;
            const bool foot = j == L_FOOT || j == R_FOOT, toe = j == L_TOE || j == R_TOE;
            if (foot || toe || j == L_HAND || j == R_HAND || j == L_LEG || j == R_LEG) {
                const float sp = toe ? (j == L_TOE ? lv[t] : rv[t]) : joint_speed(seq, T, t, j);
                const float h = __fsub_rn(seq[((long long)t * NJ + j) * 3 + 2], floor_h);
                v = (sp < VEL && h < (toe ? TOE_H : ANKLE_H)) ? 1.f : 0.f;
            }
            c[i] = v;
        }
    }
}

}  // namespace egoego
