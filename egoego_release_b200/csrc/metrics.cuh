// Post-sampling evaluation metrics on the device: one block per sequence, one warp per frame,
// lane = joint; fp64 accumulation, one pass over gt/pred joint positions [T,22,3] and global quaternions [T,22,4].
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

// quaternion_matrix: rotation of a (not necessarily unit) wxyz quaternion, identity when |q|^2 < 4 eps
__device__ __forceinline__ void quat_matrix64(const float* q, double* R) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double n = w * w + x * x + y * y + z * z;
    if (n < 8.881784197001252e-16) { R[0] = R[4] = R[8] = 1.0; R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0.0; return; }
    const double s = 2.0 / n;
    R[0] = 1.0 - s * (y * y + z * z); R[1] = s * (x * y - z * w);       R[2] = s * (x * z + y * w);
    R[3] = s * (x * y + z * w);       R[4] = 1.0 - s * (x * x + z * z); R[5] = s * (y * z - x * w);
    R[6] = s * (x * z - y * w);       R[7] = s * (y * z + x * w);       R[8] = 1.0 - s * (x * x + y * y);
}

// || I - X Y^-1 ||_F for rigid X = [Rx tx], Y = [Ry ty] (Y^-1 = [Ry^T, -Ry^T ty]); rot_only drops the translation column
__device__ __forceinline__ void frob_pair(const float* qx, const float* tx, const float* qy, const float* ty,
                                          double* full, double* rot) {
    double Rx[9], Ry[9], E[9];
    quat_matrix64(qx, Rx); quat_matrix64(qy, Ry);
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double e = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) e += Rx[i * 3 + k] * Ry[j * 3 + k];       // Rx Ry^T
            E[i * 3 + j] = e;
            const double d = (i == j ? 1.0 : 0.0) - e;
            acc += d * d;
        }
    *rot = sqrt(acc);
    double tacc = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double d = (double)tx[i] - (E[i * 3] * (double)ty[0] + E[i * 3 + 1] * (double)ty[1] + E[i * 3 + 2] * (double)ty[2]);
        tacc += d * d;
    }
    *full = sqrt(acc + tacc);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// gt_quat/pred_quat [B,T,22,4], gt_jpos/pred_jpos [B,T,22,3], floors [B] (gt, pred); out [B, MET_OUT]
static __global__ void __launch_bounds__(MET_WARPS * 32, 3)
eval_metrics_kernel(const float* __restrict__ gt_quat, const float* __restrict__ gt_jpos, const float* __restrict__ gt_floor,
                    const float* __restrict__ pred_quat, const float* __restrict__ pred_jpos, const float* __restrict__ pred_floor,
                    int T, float* __restrict__ out) {
    __shared__ double red[MET_WARPS][MET_OUT];
    const int b = blockIdx.x, warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const bool jl = lane < NJ;
    const float* gq = gt_quat + (long long)b * T * NJ * 4;
    const float* pq = pred_quat + (long long)b * T * NJ * 4;
    const float* gj = gt_jpos + (long long)b * T * NJ * 3;
    const float* pj = pred_jpos + (long long)b * T * NJ * 3;
    const float gfl = gt_floor[b], pfl = pred_floor[b];
    // per-thread partials: lane-owned (jpe of joint `lane`, foot sliding of the lane's foot joint) and warp-level (lane 0)
    double jpe = 0.0, s_acc_p = 0.0, s_acc_g = 0.0, s_acc_e = 0.0, s_fs_p = 0.0, s_fs_g = 0.0;
    double s_root_t = 0.0, s_head_t = 0.0, s_root_d = 0.0, s_root_r = 0.0, s_head_d = 0.0, s_head_r = 0.0;
    const float hfoot = (lane == 7 || lane == 8) ? 0.08f : 0.04f;
    const bool foot = lane == 7 || lane == 8 || lane == 10 || lane == 11;

    for (int t = warp; t < T; t += MET_WARPS) {
        float p0[3] = {0, 0, 0}, g0[3] = {0, 0, 0};
        if (jl) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { p0[c] = pj[((long long)t * NJ + lane) * 3 + c]; g0[c] = gj[((long long)t * NJ + lane) * 3 + c]; }
        }
        // MPJPE: root-relative (fp32 subtraction like the reference's torch tensors), norm in fp32 -> fp64 sum
        float dn = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float pr = p0[c] - __shfl_sync(0xffffffffu, p0[c], 0), gr = g0[c] - __shfl_sync(0xffffffffu, g0[c], 0);
            const float d = pr - gr;
            dn += d * d;
        }
        if (jl) jpe += (double)sqrtf(dn);
        // translation errors of root (lane 0) and head (lane 15)
        if (lane == 0 || lane == HEAD_IDX) {
            float d2 = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float d = p0[c] - g0[c]; d2 += d * d; }
            if (lane == 0) s_root_t += (double)sqrtf(d2); else s_head_t += (double)sqrtf(d2);
        }
        // accelerations over frames t, t+1, t+2 (T - 2 terms), mean over the 22 joints
        if (t + 2 < T) {
            float ap = 0.f, ag = 0.f, ae = 0.f;
            if (jl) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float p1 = pj[((long long)(t + 1) * NJ + lane) * 3 + c], p2 = pj[((long long)(t + 2) * NJ + lane) * 3 + c];
                    const float g1 = gj[((long long)(t + 1) * NJ + lane) * 3 + c], g2 = gj[((long long)(t + 2) * NJ + lane) * 3 + c];
                    const float a_p = (p2 - p1) - (p1 - p0[c]), a_g = (g2 - g1) - (g1 - g0[c]);       // compute_accel: diff of velocities
                    const float e_p = p0[c] - 2.f * p1 + p2, e_g = g0[c] - 2.f * g1 + g2;             // compute_error_accel form
                    ap += a_p * a_p; ag += a_g * a_g; ae += (e_p - e_g) * (e_p - e_g);
                }
            }
            const double wp = warp_sum(jl ? (double)sqrtf(ap) : 0.0), wg = warp_sum(jl ? (double)sqrtf(ag) : 0.0),
                         we = warp_sum(jl ? (double)sqrtf(ae) : 0.0);
            if (lane == 0) { s_acc_p += wp / NJ; s_acc_g += wg / NJ; s_acc_e += we / NJ; }
        }
        // foot sliding (frames t, t+1): |disp_xy| * (2 - 2^(z/H)) where z < H, heights relative to the floor
        if (foot && t + 1 < T) {
            const float p1x = pj[((long long)(t + 1) * NJ + lane) * 3], p1y = pj[((long long)(t + 1) * NJ + lane) * 3 + 1];
            const float g1x = gj[((long long)(t + 1) * NJ + lane) * 3], g1y = gj[((long long)(t + 1) * NJ + lane) * 3 + 1];
            const float zp = p0[2] - pfl, zg = g0[2] - gfl;
            if (zp < hfoot) s_fs_p += (double)fabsf(sqrtf((p1x - p0[0]) * (p1x - p0[0]) + (p1y - p0[1]) * (p1y - p0[1])) * (2.f - exp2f(zp / hfoot)));
            if (zg < hfoot) s_fs_g += (double)fabsf(sqrtf((g1x - g0[0]) * (g1x - g0[0]) + (g1y - g0[1]) * (g1y - g0[1])) * (2.f - exp2f(zg / hfoot)));
        }
    }
    // pose-matrix distances (fp64, ~300 flops per pose pair): a second pass with lane = (frame, root | head), so all 32 lanes
    // carry fp64 work instead of two lanes per frame holding up their warp
    for (int t0 = warp * 16; t0 < T; t0 += MET_WARPS * 16) {
        const int t = t0 + (lane & 15), j = lane < 16 ? 0 : HEAD_IDX;
        if (t < T) {
            const long long o = (long long)t * NJ + j;
            const float pp[3] = {pj[o * 3], pj[o * 3 + 1], pj[o * 3 + 2]}, gp[3] = {gj[o * 3], gj[o * 3 + 1], gj[o * 3 + 2]};
            double full, rot;
            frob_pair(pq + o * 4, pp, gq + o * 4, gp, &full, &rot);
            if (lane < 16) { s_root_d += full; s_root_r += rot; } else { s_head_d += full; s_head_r += rot; }
        }
    }
    s_root_d = warp_sum(s_root_d); s_root_r = warp_sum(s_root_r); s_head_d = warp_sum(s_head_d); s_head_r = warp_sum(s_head_r);
    // warp-level assembly: scalars into red[warp][0..12], per-joint jpe sums into red[warp][13 + joint]
    const double fs_p = warp_sum(s_fs_p), fs_g = warp_sum(s_fs_g);
    const double head_t = __shfl_sync(0xffffffffu, s_head_t, HEAD_IDX), head_d = s_head_d, head_r = s_head_r;
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
