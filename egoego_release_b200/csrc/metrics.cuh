// Post-sampling evaluation metrics on the device: one block per sequence, one warp per frame,
// lane = joint; per-term arithmetic in fp32 (the reference's arrays are float32), accumulation over frames in fp64.
//
// Follows compute_metrics_for_smpl (kinpoly/scripts/eval_metrics_imu_rec.py:264-342) and what it calls:
// compute_accel / compute_error_accel (:66-107), compute_foot_sliding_for_smpl (:222-262), get_root_matrix /
// get_frobenious_norm(_rot_only) (kinpoly/relive/utils/metrics.py:15-24,64-82), quaternion_matrix
// (kinpoly/relive/utils/transformation.py:1346-1370: wxyz, normalised, identity below _EPS).
// HBM-bound: reads the joint positions (2 x T x 22 x 3 floats) and the root / head quaternions (2 x T x 2 x 4 floats) of a
// sequence (~71 KB at T = 120), writes 35 floats.
#pragma once
#include "common.cuh"

namespace egoego {

constexpr int MET_SCALARS = 13;                 // order = oracle/metrics.py KEYS
constexpr int MET_OUT = MET_SCALARS + NJ;       // + single_jpe[22]
constexpr int MET_WARPS = 8;
constexpr int MET_CHUNK = 32;                   // frames per shared-memory chunk

// quaternion_matrix: rotation of a (not necessarily unit) wxyz quaternion, identity when |q|^2 < 4 eps(double).
// fp32: the inputs are fp32 and the distances below are O(1e-3 .. 1); fp64 issue rate would bound the whole kernel
// (measured: 209 us with fp64 pose algebra for 2048 sequences)
__device__ __forceinline__ void quat_matrix32(const float* q, float* R) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float n = w * w + x * x + y * y + z * z;
    if (n < 8.881784197001252e-16f) { R[0] = R[4] = R[8] = 1.f; R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0.f; return; }
    const float s = 2.0f / n;
    R[0] = 1.0f - s * (y * y + z * z); R[1] = s * (x * y - z * w);        R[2] = s * (x * z + y * w);
    R[3] = s * (x * y + z * w);        R[4] = 1.0f - s * (x * x + z * z); R[5] = s * (y * z - x * w);
    R[6] = s * (x * z - y * w);        R[7] = s * (y * z + x * w);        R[8] = 1.0f - s * (x * x + y * y);
}

// || I - X Y^-1 ||_F for rigid X = [Rx tx], Y = [Ry ty] (Y^-1 = [Ry^T, -Ry^T ty]); rot_only drops the translation column
__device__ __forceinline__ void frob_pair(const float* qx, const float* tx, const float* qy, const float* ty,
                                          float* full, float* rot) {
    float Rx[9], Ry[9], E[9];
    quat_matrix32(qx, Rx); quat_matrix32(qy, Ry);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float e = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) e = fmaf(Rx[i * 3 + k], Ry[j * 3 + k], e);       // Rx Ry^T
            E[i * 3 + j] = e;
            const float d = (i == j ? 1.0f : 0.0f) - e;
            acc = fmaf(d, d, acc);
        }
    *rot = sqrtf(acc);
    float tacc = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float d = tx[i] - (E[i * 3] * ty[0] + E[i * 3 + 1] * ty[1] + E[i * 3 + 2] * ty[2]);
        tacc = fmaf(d, d, tacc);
    }
    *full = sqrtf(acc + tacc);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// gt_quat/pred_quat [B,T,22,4], gt_jpos/pred_jpos [B,T,22,3], floors [B] (gt, pred); out [B, MET_OUT]
// ncu (round 1): the first version spent ~440 instructions per frame per warp (divergent two- and four-lane branches with
// sqrt / exp2 / divisions executed by whole warps, three warp reductions per frame) and was issue-bound at 190 us for 2048
// sequences.  Now every branch-specific quantity has its own lane mapping:
//   pass A  lane = joint            MPJPE terms, acceleration terms (per-lane sums over frames, reduced once at the end)
//   pass B  lane = (frame, root|head)        translation errors and pose-matrix distances
//   pass C  lane = (frame, foot joint)       foot sliding
static __global__ void __launch_bounds__(MET_WARPS * 32, 3)
eval_metrics_kernel(const float* __restrict__ gt_quat, const float* __restrict__ gt_jpos, const float* __restrict__ gt_floor,
                    const float* __restrict__ pred_quat, const float* __restrict__ pred_jpos, const float* __restrict__ pred_floor,
                    int T, float* __restrict__ out) {
    __shared__ double red[MET_WARPS][MET_OUT];
    __shared__ float sp[2][(MET_CHUNK + 2) * NJ * 3];     // [pred | gt] joint positions of the current chunk
    const int b = blockIdx.x, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const bool jl = lane < NJ;
    const float* gq = gt_quat + (long long)b * T * NJ * 4;
    const float* pq = pred_quat + (long long)b * T * NJ * 4;
    const float* gj = gt_jpos + (long long)b * T * NJ * 3;
    const float* pj = pred_jpos + (long long)b * T * NJ * 3;
    const float gfl = gt_floor[b], pfl = pred_floor[b];
    double jpe = 0.0, s_acc_p = 0.0, s_acc_g = 0.0, s_acc_e = 0.0, s_fs_p = 0.0, s_fs_g = 0.0;
    double s_root_t = 0.0, s_head_t = 0.0, s_root_d = 0.0, s_root_r = 0.0, s_head_d = 0.0, s_head_r = 0.0;
    const int lj = jl ? lane : 0;                        // lanes >= 22 shadow joint 0 (their results are discarded)

    for (int c0 = 0; c0 < T; c0 += MET_CHUNK) {
        const int nload = min(MET_CHUNK + 2, T - c0) * NJ * 3;
        __syncthreads();                                   // previous chunk fully consumed
        for (int i = threadIdx.x; i < nload; i += MET_WARPS * 32) {
            sp[0][i] = pj[(long long)c0 * NJ * 3 + i];
            sp[1][i] = gj[(long long)c0 * NJ * 3 + i];
        }
        __syncthreads();
        const int t_end = min(c0 + MET_CHUNK, T);
        // ---- pass A: lane = joint ----
        for (int t = c0 + warp; t < t_end; t += MET_WARPS) {
            const float* P = sp[0] + (t - c0) * NJ * 3;
            const float* G = sp[1] + (t - c0) * NJ * 3;
            float p0[3], g0[3], dn = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                p0[c] = P[lj * 3 + c]; g0[c] = G[lj * 3 + c];
                const float d = (p0[c] - P[c]) - (g0[c] - G[c]);      // root-relative, fp32 like the reference's tensors
                dn = fmaf(d, d, dn);
            }
            jpe += (double)sqrtf(dn);
            if (t + 2 < T) {                               // warp-uniform
                float ap = 0.f, ag = 0.f, ae = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float p1 = P[(NJ + lj) * 3 + c], p2 = P[(2 * NJ + lj) * 3 + c], g1 = G[(NJ + lj) * 3 + c], g2 = G[(2 * NJ + lj) * 3 + c];
                    const float a_p = (p2 - p1) - (p1 - p0[c]), a_g = (g2 - g1) - (g1 - g0[c]);       // compute_accel: diff of velocities
                    const float e_p = p0[c] - 2.f * p1 + p2, e_g = g0[c] - 2.f * g1 + g2;             // compute_error_accel form
                    ap = fmaf(a_p, a_p, ap); ag = fmaf(a_g, a_g, ag); ae = fmaf(e_p - e_g, e_p - e_g, ae);
                }
                s_acc_p += (double)sqrtf(ap); s_acc_g += (double)sqrtf(ag); s_acc_e += (double)sqrtf(ae);   // per joint; /22 at the end
            }
        }
        // ---- pass B: lane = (frame, root | head): translation errors + pose-matrix distances (quaternions from global) ----
        for (int t0 = c0 + warp * 16; t0 < t_end; t0 += MET_WARPS * 16) {
            const int t = t0 + (lane & 15), j = lane < 16 ? 0 : HEAD_IDX;
            if (t < t_end) {
                const float* P = sp[0] + ((t - c0) * NJ + j) * 3;
                const float* G = sp[1] + ((t - c0) * NJ + j) * 3;
                const float pp[3] = {P[0], P[1], P[2]}, gp[3] = {G[0], G[1], G[2]};
                const float dx = pp[0] - gp[0], dy = pp[1] - gp[1], dz = pp[2] - gp[2];
                const float te = sqrtf(dx * dx + dy * dy + dz * dz);
                float full, rot;
                const long long o = (long long)t * NJ + j;
                frob_pair(pq + o * 4, pp, gq + o * 4, gp, &full, &rot);
                if (lane < 16) { s_root_t += (double)te; s_root_d += (double)full; s_root_r += (double)rot; }
                else           { s_head_t += (double)te; s_head_d += (double)full; s_head_r += (double)rot; }
            }
        }
        // ---- pass C: lane = (frame, foot joint): |disp_xy| (2 - 2^(z/H)) where z < H, heights relative to the floor ----
        for (int t0 = c0 + warp * 8; t0 < t_end; t0 += MET_WARPS * 8) {
            const int t = t0 + (lane & 7), f = lane >> 3;                  // f: l ankle, l toe, r ankle, r toe
            if (t < t_end && t + 1 < T) {
                const int j = f == 0 ? 7 : (f == 1 ? 10 : (f == 2 ? 8 : 11));
                const float h = (f & 1) ? 0.04f : 0.08f;
                const float* P = sp[0] + ((t - c0) * NJ + j) * 3;
                const float* G = sp[1] + ((t - c0) * NJ + j) * 3;
                const float zp = P[2] - pfl, zg = G[2] - gfl;
                const float dpx = P[NJ * 3] - P[0], dpy = P[NJ * 3 + 1] - P[1], dgx = G[NJ * 3] - G[0], dgy = G[NJ * 3 + 1] - G[1];
                if (zp < h) s_fs_p += (double)fabsf(sqrtf(dpx * dpx + dpy * dpy) * (2.f - exp2f(zp / h)));
                if (zg < h) s_fs_g += (double)fabsf(sqrtf(dgx * dgx + dgy * dgy) * (2.f - exp2f(zg / h)));
            }
        }
    }
    if (!jl) { jpe = 0.0; s_acc_p = 0.0; s_acc_g = 0.0; s_acc_e = 0.0; }
    s_acc_p = warp_sum(s_acc_p) / NJ; s_acc_g = warp_sum(s_acc_g) / NJ; s_acc_e = warp_sum(s_acc_e) / NJ;
    s_root_t = warp_sum(s_root_t); s_head_t = warp_sum(s_head_t);
    s_root_d = warp_sum(s_root_d); s_root_r = warp_sum(s_root_r); s_head_d = warp_sum(s_head_d); s_head_r = warp_sum(s_head_r);
    // warp-level assembly: scalars into red[warp][0..12], per-joint jpe sums into red[warp][13 + joint]
    const double fs_p = warp_sum(s_fs_p), fs_g = warp_sum(s_fs_g);
    const double head_t = s_head_t, head_d = s_head_d, head_r = s_head_r;
    if (lane == 0) {
        double* r = red[warp];
        r[0] = s_root_t; r[1] = s_acc_p; r[2] = s_acc_g; r[3] = s_acc_e; r[4] = fs_p; r[5] = fs_g; r[6] = head_t;
        r[7] = s_root_d; r[8] = s_root_r; r[9] = 0.0; r[10] = 0.0; r[11] = head_d; r[12] = head_r;
    }
    if (jl) red[warp][MET_SCALARS + lane] = jpe;
    __syncthreads();
    if (threadIdx.x < MET_OUT) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < MET_WARPS; ++w) v += red[w][threadIdx.x];
        red[0][threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double* r = red[0];
        float* o = out + (long long)b * MET_OUT;
        const double nT = (double)T, nA = (double)(T - 2);
        double all = 0.0, wo_hand = 0.0;
        for (int j = 0; j < NJ; ++j) {
            const double sj = r[MET_SCALARS + j] / nT * 1000.0;
            o[MET_SCALARS + j] = (float)sj;
            all += sj;
            if (j < 18) wo_hand += sj;
        }
        o[0] = (float)(r[0] / nT * 1000.0);
        o[1] = (float)(r[1] / nA * 1000.0); o[2] = (float)(r[2] / nA * 1000.0); o[3] = (float)(r[3] / nA * 1000.0);
        o[4] = (float)(r[4] / nT * 1000.0 / 4.0); o[5] = (float)(r[5] / nT * 1000.0 / 4.0);
        o[6] = (float)(r[6] / nT * 1000.0);
        o[7] = (float)(r[7] / nT); o[8] = (float)(r[8] / nT);
        o[9] = (float)(all / NJ); o[10] = (float)(wo_hand / 18.0);
        o[11] = (float)(r[11] / nT); o[12] = (float)(r[12] / nT);
    }
}

}  // namespace egoego
