// Post-sampling geometry on the device: rot6d -> rotation, IK to local axis-angle, SMPL FK,
// (de)normalisation and head-frame canonicalisation.  One warp per frame, lane = joint; the kinematic
// tree is walked level-synchronously with warp shuffles (tree depth of the 22-joint SMPL body is 7).
//
// Follows convert_model_res_to_data (egoego/model/transformer_cond_diffusion_model.py:469-525),
// quat_ik_torch / fk_smpl / (de_)normalize_jpos_min_max (egoego/data/amass_diffusion_dataset.py:107-125,
// 265-293,379-392), rotate_at_frame_smplh (egoego/lafan1/utils.py:111-137) and the published
// definitions of the pytorch3d.transforms functions they call (wxyz quaternions).
#pragma once
#include "common.cuh"

namespace egoego {

struct Skeleton {
    int   parents[NJ];
    int   depth[NJ];
    float off[NJ][3];
    float jmin[NJ * 3];
    float jmax[NJ * 3];
    int   max_depth;
};

struct Q4 { float w, x, y, z; };
struct M3 { float m[9]; };

__device__ __forceinline__ Q4 q_raw_mul(Q4 a, Q4 b) {
    Q4 o;
    o.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    o.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    o.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
    o.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
    return o;
}
__device__ __forceinline__ Q4 q_std(Q4 q) { if (q.w < 0.f) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; } return q; }
__device__ __forceinline__ Q4 q_mul(Q4 a, Q4 b) { return q_std(q_raw_mul(a, b)); }
__device__ __forceinline__ Q4 q_inv(Q4 q) { Q4 o = {q.w, -q.x, -q.y, -q.z}; return o; }
__device__ __forceinline__ void q_apply(Q4 q, const float p[3], float out[3]) {
    Q4 p4 = {0.f, p[0], p[1], p[2]};
    Q4 r = q_raw_mul(q_raw_mul(q, p4), q_inv(q));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
__device__ __forceinline__ M3 q_to_mat(Q4 q) {
    float two_s = 2.0f / (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    float r = q.w, i = q.x, j = q.y, k = q.z;
    M3 o;
    o.m[0] = 1 - two_s * (j * j + k * k); o.m[1] = two_s * (i * j - k * r); o.m[2] = two_s * (i * k + j * r);
    o.m[3] = two_s * (i * j + k * r); o.m[4] = 1 - two_s * (i * i + k * k); o.m[5] = two_s * (j * k - i * r);
    o.m[6] = two_s * (i * k - j * r); o.m[7] = two_s * (j * k + i * r); o.m[8] = 1 - two_s * (i * i + j * j);
    return o;
}
__device__ __forceinline__ float sqrt_pos(float x) { return x > 0.f ? sqrtf(x) : 0.f; }
__device__ __forceinline__ Q4 mat_to_q(const M3& M) {
    const float m00 = M.m[0], m01 = M.m[1], m02 = M.m[2], m10 = M.m[3], m11 = M.m[4], m12 = M.m[5],
                m20 = M.m[6], m21 = M.m[7], m22 = M.m[8];
    float qa[4] = {sqrt_pos(1.0f + m00 + m11 + m22), sqrt_pos(1.0f + m00 - m11 - m22),
                   sqrt_pos(1.0f - m00 + m11 - m22), sqrt_pos(1.0f - m00 - m11 + m22)};
    int best = 0;
#pragma unroll
    for (int c = 1; c < 4; ++c) if (qa[c] > qa[best]) best = c;   // first max wins (argmax)
    float c0, c1, c2, c3;
    if (best == 0)      { c0 = qa[0] * qa[0]; c1 = m21 - m12; c2 = m02 - m20; c3 = m10 - m01; }
    else if (best == 1) { c0 = m21 - m12; c1 = qa[1] * qa[1]; c2 = m10 + m01; c3 = m02 + m20; }
    else if (best == 2) { c0 = m02 - m20; c1 = m10 + m01; c2 = qa[2] * qa[2]; c3 = m12 + m21; }
    else                { c0 = m10 - m01; c1 = m20 + m02; c2 = m21 + m12; c3 = qa[3] * qa[3]; }
    float d = 2.0f * fmaxf(qa[best], 0.1f);
    Q4 q = {c0 / d, c1 / d, c2 / d, c3 / d};
    return q;
}
__device__ __forceinline__ M3 rot6d_to_mat(const float d[6]) {
    float n1 = fmaxf(sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), 1e-12f);
    float b1[3] = {d[0] / n1, d[1] / n1, d[2] / n1};
    float dot = b1[0] * d[3] + b1[1] * d[4] + b1[2] * d[5];
    float b2[3] = {d[3] - dot * b1[0], d[4] - dot * b1[1], d[5] - dot * b1[2]};
    float n2 = fmaxf(sqrtf(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]), 1e-12f);
    b2[0] /= n2; b2[1] /= n2; b2[2] /= n2;
    M3 o;
    o.m[0] = b1[0]; o.m[1] = b1[1]; o.m[2] = b1[2];
    o.m[3] = b2[0]; o.m[4] = b2[1]; o.m[5] = b2[2];
    o.m[6] = b1[1] * b2[2] - b1[2] * b2[1];
    o.m[7] = b1[2] * b2[0] - b1[0] * b2[2];
    o.m[8] = b1[0] * b2[1] - b1[1] * b2[0];
    return o;
}
__device__ __forceinline__ void q_to_aa(Q4 q, float aa[3]) {
    float n = sqrtf(q.x * q.x + q.y * q.y + q.z * q.z);
    float half = atan2f(n, q.w);
    float ang = 2.0f * half;
    float s = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.0f) : (sinf(half) / ang);
    aa[0] = q.x / s; aa[1] = q.y / s; aa[2] = q.z / s;
}
__device__ __forceinline__ Q4 aa_to_q(const float aa[3]) {
    float ang = sqrtf(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]);
    float half = 0.5f * ang;
    float s = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.0f) : (sinf(half) / ang);
    Q4 q = {cosf(half), aa[0] * s, aa[1] * s, aa[2] * s};
    return q;
}
__device__ __forceinline__ Q4 shfl_q(Q4 q, int src) {
    Q4 o;
    o.w = __shfl_sync(0xffffffffu, q.w, src); o.x = __shfl_sync(0xffffffffu, q.x, src);
    o.y = __shfl_sync(0xffffffffu, q.y, src); o.z = __shfl_sync(0xffffffffu, q.z, src);
    return o;
}

// FK over the tree for one frame held by a warp (lane = joint).  lrot: local quaternion of this lane's
// joint; returns global quaternion + position (root translation already added).
__device__ __forceinline__ void warp_fk(const Skeleton& sk, int j, Q4 lrot, const float root[3], Q4& gr, float gp[3]) {
    const bool act = j < NJ;
    const int par = act ? sk.parents[j] : 0;
    const int dep = act ? sk.depth[j] : -1;
    gr = lrot;
    gp[0] = act ? sk.off[j][0] : 0.f; gp[1] = act ? sk.off[j][1] : 0.f; gp[2] = act ? sk.off[j][2] : 0.f;
    for (int lvl = 1; lvl <= sk.max_depth; ++lvl) {
        int src = par < 0 ? 0 : par;
        Q4 pr = shfl_q(gr, src);
        float pp0 = __shfl_sync(0xffffffffu, gp[0], src);
        float pp1 = __shfl_sync(0xffffffffu, gp[1], src);
        float pp2 = __shfl_sync(0xffffffffu, gp[2], src);
        if (dep == lvl) {
            float o[3] = {sk.off[j][0], sk.off[j][1], sk.off[j][2]}, r[3];
            q_apply(pr, o, r);
            gp[0] = r[0] + pp0; gp[1] = r[1] + pp1; gp[2] = r[2] + pp2;
            gr = q_mul(pr, lrot);
        }
    }
    gp[0] += root[0]; gp[1] += root[1]; gp[2] += root[2];
}

static __global__ void __launch_bounds__(256) postprocess_kernel(Skeleton sk, const float* __restrict__ x,
                                                          const float* __restrict__ recover_quat, int B, int T,
                                                          float* __restrict__ aa_out, float* __restrict__ root_out,
                                                          float* __restrict__ head_out, float* __restrict__ jpos_out,
                                                          float* __restrict__ gquat_out) {
    const int frame = blockIdx.x * 8 + threadIdx.x / 32;   // warp-uniform
    const int j = threadIdx.x % 32;
    if (frame >= B * T) return;
    const int w = frame / T;
    const float* xf = x + (long long)frame * 198;
    const bool act = j < NJ;
    Q4 rq = {1.f, 0.f, 0.f, 0.f};
    if (recover_quat) { rq.w = recover_quat[w * 4 + 0]; rq.x = recover_quat[w * 4 + 1];
                        rq.y = recover_quat[w * 4 + 2]; rq.z = recover_quat[w * 4 + 3]; }
    float d6[6] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f};
    if (act) {
#pragma unroll
        for (int c = 0; c < 6; ++c) d6[c] = xf[66 + j * 6 + c];
    }
    Q4 gq = mat_to_q(rot6d_to_mat(d6));
    Q4 ori = q_mul(rq, gq);
    Q4 grot = mat_to_q(q_to_mat(ori));                       // quat_ik_torch re-derives from matrices
    int par = act ? sk.parents[j] : -1;
    Q4 pq = shfl_q(grot, par < 0 ? 0 : par);
    Q4 local = (par < 0) ? grot : q_mul(q_inv(pq), grot);
    Q4 lq = mat_to_q(q_to_mat(local));                       // matrix_to_axis_angle(quaternion_to_matrix(res))
    float aa[3];
    q_to_aa(lq, aa);
    if (act && aa_out) {
        float* o = aa_out + ((long long)frame * NJ + j) * 3;
        o[0] = aa[0]; o[1] = aa[1]; o[2] = aa[2];
    }
    // de-normalised root (joint 0) / head (joint 15) positions rotated back by recover_quat
    float pos[3] = {0.f, 0.f, 0.f};
    if (act) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float n = (xf[j * 3 + c] + 1.0f) * 0.5f;
            pos[c] = n * (sk.jmax[j * 3 + c] - sk.jmin[j * 3 + c]) + sk.jmin[j * 3 + c];
        }
    }
    float rp[3];
    q_apply(rq, pos, rp);
    if (j == 0 && root_out) { float* o = root_out + (long long)frame * 3; o[0] = rp[0]; o[1] = rp[1]; o[2] = rp[2]; }
    if (j == HEAD_IDX && head_out) { float* o = head_out + (long long)frame * 3; o[0] = rp[0]; o[1] = rp[1]; o[2] = rp[2]; }
    if (jpos_out || gquat_out) {
        float root[3];
        root[0] = __shfl_sync(0xffffffffu, rp[0], 0);
        root[1] = __shfl_sync(0xffffffffu, rp[1], 0);
        root[2] = __shfl_sync(0xffffffffu, rp[2], 0);
        Q4 lrot = mat_to_q(q_to_mat(aa_to_q(aa)));           // fk_smpl: axis_angle_to_matrix -> matrix_to_quaternion
        Q4 gr; float gp[3];
        warp_fk(sk, j, lrot, root, gr, gp);
        if (act && jpos_out) { float* o = jpos_out + ((long long)frame * NJ + j) * 3; o[0] = gp[0]; o[1] = gp[1]; o[2] = gp[2]; }
        if (act && gquat_out) { float* o = gquat_out + ((long long)frame * NJ + j) * 4; o[0] = gr.w; o[1] = gr.x; o[2] = gr.y; o[3] = gr.z; }
    }
}

static __global__ void __launch_bounds__(256) fk_smpl_kernel(Skeleton sk, const float* __restrict__ root_in,
                                                      const float* __restrict__ aa_in, long long N,
                                                      float* __restrict__ gquat_out, float* __restrict__ jpos_out) {
    const long long frame = (long long)blockIdx.x * 8 + threadIdx.x / 32;
    const int j = threadIdx.x % 32;
    if (frame >= N) return;
    const bool act = j < NJ;
    float aa[3] = {0.f, 0.f, 0.f};
    if (act) { const float* a = aa_in + (frame * NJ + j) * 3; aa[0] = a[0]; aa[1] = a[1]; aa[2] = a[2]; }
    float root[3] = {root_in[frame * 3 + 0], root_in[frame * 3 + 1], root_in[frame * 3 + 2]};
    Q4 lrot = mat_to_q(q_to_mat(aa_to_q(aa)));
    Q4 gr; float gp[3];
    warp_fk(sk, j, lrot, root, gr, gp);
    if (act && jpos_out) { float* o = jpos_out + (frame * NJ + j) * 3; o[0] = gp[0]; o[1] = gp[1]; o[2] = gp[2]; }
    if (act && gquat_out) { float* o = gquat_out + (frame * NJ + j) * 4; o[0] = gr.w; o[1] = gr.x; o[2] = gr.y; o[3] = gr.z; }
}

// rotate_at_frame_smplh(cano_t_idx = 0) + "move first frame's x,y to 0" + x_start construction
// (transformer_cond_diffusion_model.py:358-386).  One thread per (window, frame).
__device__ __forceinline__ void qv_rot(Q4 q, const float x[3], float out[3]) {   // lafan1 quat_mul_vec
    float t[3] = {2.0f * (q.y * x[2] - q.z * x[1]), 2.0f * (q.z * x[0] - q.x * x[2]), 2.0f * (q.x * x[1] - q.y * x[0])};
    out[0] = x[0] + q.w * t[0] + (q.y * t[2] - q.z * t[1]);
    out[1] = x[1] + q.w * t[1] + (q.z * t[0] - q.x * t[2]);
    out[2] = x[2] + q.w * t[2] + (q.x * t[1] - q.y * t[0]);
}

static __global__ void canonicalize_head_kernel(Skeleton sk, const float* __restrict__ head_pos,
                                         const float* __restrict__ head_quat, long long stride_frames, int B, int T,
                                         float* __restrict__ x_start, float* __restrict__ recover_quat) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= B * T) return;
    const int w = gid / T, f = gid % T;
    const float* p0 = head_pos + (long long)w * stride_frames * 3;
    const float* q0 = head_quat + (long long)w * stride_frames * 4;
    Q4 key = {q0[0], q0[1], q0[2], q0[3]};
    const float ex[3] = {1.f, 0.f, 0.f};
    float fw[3];
    qv_rot(key, ex, fw);
    fw[2] = 0.f;                                                  // project to the xy plane
    float fn = sqrtf(fw[0] * fw[0] + fw[1] * fw[1]) + 1e-8f;
    fw[0] /= fn; fw[1] /= fn;
    // quat_between([1,0,0], fw) = [ |x||y| + x.y , x cross y ], then normalise
    float yw = sqrtf(1.0f * (fw[0] * fw[0] + fw[1] * fw[1] + fw[2] * fw[2])) + fw[0];
    float yx = 0.f, yy = -fw[2], yz = fw[1];
    float yn = sqrtf(yw * yw + yx * yx + yy * yy + yz * yz) + 1e-8f;
    Q4 yrot = {yw / yn, yx / yn, yy / yn, yz / yn};
    Q4 inv = q_inv(yrot);
    if (f == 0 && recover_quat) { float* o = recover_quat + w * 4; o[0] = yrot.w; o[1] = yrot.x; o[2] = yrot.y; o[3] = yrot.z; }
    float pf[3] = {p0[f * 3 + 0], p0[f * 3 + 1], p0[f * 3 + 2]}, pz[3] = {p0[0], p0[1], p0[2]};
    float af[3], a0[3];
    qv_rot(inv, pf, af);
    qv_rot(inv, pz, a0);
    af[0] -= a0[0]; af[1] -= a0[1];                               // z kept
    Q4 qf = {q0[f * 4 + 0], q0[f * 4 + 1], q0[f * 4 + 2], q0[f * 4 + 3]};
    // lafan1 quat_mul(x=inv, y=q): raw Hamilton product x*y (no sign standardisation)
    Q4 aq = q_raw_mul(inv, qf);
    M3 Rm = q_to_mat(aq);
    float* xs = x_start + (long long)gid * 198;
    for (int c = 0; c < 198; ++c) xs[c] = 0.f;
    // normalize_jpos_min_max applied to ALL 22 joints of a frame whose other joints are zero
    for (int jj = 0; jj < NJ; ++jj)
        for (int c = 0; c < 3; ++c) {
            float v = (jj == HEAD_IDX) ? af[c] : 0.f;
            float n = (v - sk.jmin[jj * 3 + c]) / (sk.jmax[jj * 3 + c] - sk.jmin[jj * 3 + c]);
            xs[jj * 3 + c] = n * 2.0f - 1.0f;
        }
#pragma unroll
    for (int c = 0; c < 6; ++c) xs[66 + HEAD_IDX * 6 + c] = Rm.m[c];
}

// Conditioning of the NEXT sliding window from the tail of the current one (reference :423-464): the last `n`
// frames' FK result (global quaternions / joint positions) is re-canonicalised at its first frame (yaw of the head
// joint, x/y of the head moved to 0), positions are min-max normalised and rotations converted to rot6d, giving the
// [B, n, 198] tensor that overwrites the first n frames after every step of the next window.
// One warp per frame, lane = joint.
static __global__ void __launch_bounds__(256) tail_condition_kernel(Skeleton sk, const float* __restrict__ gquat,
                                                                    const float* __restrict__ gjpos, int B, int n,
                                                                    float* __restrict__ out) {
    const int frame = blockIdx.x * 8 + threadIdx.x / 32;
    const int j = threadIdx.x % 32;
    if (frame >= B * n) return;
    const int w = frame / n;
    const float* q0 = gquat + ((long long)w * n * NJ + HEAD_IDX) * 4;      // head joint, first frame of the tail
    const float* p0 = gjpos + ((long long)w * n * NJ + HEAD_IDX) * 3;
    Q4 key = {q0[0], q0[1], q0[2], q0[3]};
    const float ex[3] = {1.f, 0.f, 0.f};
    float fw[3];
    qv_rot(key, ex, fw);
    fw[2] = 0.f;
    float fn = sqrtf(fw[0] * fw[0] + fw[1] * fw[1]) + 1e-8f;
    fw[0] /= fn; fw[1] /= fn;
    float yw = sqrtf(fw[0] * fw[0] + fw[1] * fw[1]) + fw[0];
    float yy = -fw[2], yz = fw[1];
    float yn = sqrtf(yw * yw + yy * yy + yz * yz) + 1e-8f;
    Q4 yrot = {yw / yn, 0.f, yy / yn, yz / yn};
    Q4 inv = q_inv(yrot);
    float hp[3] = {p0[0], p0[1], p0[2]}, a0[3];
    qv_rot(inv, hp, a0);                                   // rotate_at_frame_smplh on the head positions, frame 0
    if (j >= NJ) return;
    const float* pj = gjpos + ((long long)frame * NJ + j) * 3;
    const float* qj = gquat + ((long long)frame * NJ + j) * 4;
    float pin[3] = {pj[0], pj[1], pj[2]}, pr[3];
    q_apply(inv, pin, pr);
    pr[0] -= a0[0]; pr[1] -= a0[1];
    float* o = out + (long long)frame * 198;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        o[j * 3 + c] = (pr[c] - sk.jmin[j * 3 + c]) / (sk.jmax[j * 3 + c] - sk.jmin[j * 3 + c]) * 2.0f - 1.0f;
    Q4 qq = {qj[0], qj[1], qj[2], qj[3]};
    M3 Rm = q_to_mat(q_mul(inv, qq));
#pragma unroll
    for (int c = 0; c < 6; ++c) o[66 + j * 6 + c] = Rm.m[c];
}

}  // namespace egoego
