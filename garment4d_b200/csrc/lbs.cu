// SMPL linear-blend skinning for sm_100a -- replaces the ~45 small torch kernels (einsum/matmul/cat/pad,
// and a Python loop of 23 sequential 4x4 matmuls) that the reference's lbs() issues
// (smplx/smplx/lbs.py:152-248) with five launches, all fp32 (1e-5 abs budget):
//   lbs_pose_kernel        batch_rodrigues (lbs.py:312-346) + pose_feature = R[1:] - I (lbs.py:216-225)
//   lbs_shape_kernel       v_shaped = v_template + blend_shapes (lbs.py:205, 288-309)
//   lbs_joints_kernel      J = J_regressor . v_shaped (lbs.py:209, 251-268)
//   sgemm_acc_kernel       v_posed = pose_feature . posedirs + v_shaped (lbs.py:220-229), FFMA GEMM
//   lbs_rigid_kernel       batch_rigid_transform (lbs.py:362-419): one warp walks the kinematic tree
//   lbs_skin_kernel        T = W.A ; verts = T.[v;1] (lbs.py:233-246) without materialising T (441 KB/frame)
#include "common.cuh"

namespace g4d {

// ---- Rodrigues: R = I + sin(a) K + (1-cos(a)) K^2, a = ||v + 1e-8||, K = skew(v / a)  (lbs.py:330-345)
__device__ __forceinline__ void rodrigues(const float vx, const float vy, const float vz, float* R) {
    const float ax = vx + 1e-8f, ay = vy + 1e-8f, az = vz + 1e-8f;
    const float angle = sqrtf(ax * ax + ay * ay + az * az);
    const float rx = vx / angle, ry = vy / angle, rz = vz / angle;
    float s, c;
    sincosf(angle, &s, &c);
    const float omc = 1.f - c;
    // K = [[0,-rz,ry],[rz,0,-rx],[-ry,rx,0]];  K^2 written out
    const float k00 = -(rz * rz) - ry * ry, k01 = rx * ry, k02 = rx * rz;
    const float k11 = -(rz * rz) - rx * rx, k12 = ry * rz;
    const float k22 = -(ry * ry) - rx * rx;
    R[0] = 1.f + omc * k00;        R[1] = s * -rz + omc * k01;   R[2] = s * ry + omc * k02;
    R[3] = s * rz + omc * k01;     R[4] = 1.f + omc * k11;       R[5] = s * -rx + omc * k12;
    R[6] = s * -ry + omc * k02;    R[7] = s * rx + omc * k12;    R[8] = 1.f + omc * k22;
}

__global__ void batch_rodrigues_kernel(int n, const float* __restrict__ rot_vecs, float* __restrict__ R) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float r[9];
    rodrigues(rot_vecs[3 * i], rot_vecs[3 * i + 1], rot_vecs[3 * i + 2], r);
#pragma unroll
    for (int e = 0; e < 9; ++e) R[9 * (size_t)i + e] = r[e];
}

// one thread per (frame, joint): rot_mats (F,J,9) and pose_feature (F,(J-1)*9)
__global__ void lbs_pose_kernel(int F, int J, int pose2rot, const float* __restrict__ pose, float* __restrict__ rot_mats,
                                float* __restrict__ pose_feature) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * J) return;
    const int f = i / J, j = i - f * J;
    float r[9];
    if (pose2rot) rodrigues(pose[3 * (size_t)i], pose[3 * (size_t)i + 1], pose[3 * (size_t)i + 2], r);
    else {
#pragma unroll
        for (int e = 0; e < 9; ++e) r[e] = pose[9 * (size_t)i + e];
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) rot_mats[9 * (size_t)i + e] = r[e];
    if (j > 0) {
        float* pf = pose_feature + (size_t)f * (J - 1) * 9 + (size_t)(j - 1) * 9;
#pragma unroll
        for (int e = 0; e < 9; ++e) pf[e] = r[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
    }
}

// v_shaped[f, e] = v_template[e] + sum_l betas[f,l] * shapedirs[e, l],  e in [0, 3V)
__global__ void lbs_shape_kernel(int F, int V3, int NB, int betas_rows, const float* __restrict__ betas,
                                 const float* __restrict__ v_template, const float* __restrict__ shapedirs,
                                 float* __restrict__ v_shaped) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= V3) return;
    const float vt = v_template ? __ldg(v_template + e) : 0.f;      // null template: the displacement alone (blend_shapes)
    const float* sd = shapedirs + (size_t)e * NB;
    for (int f = blockIdx.y; f < F; f += gridDim.y) {
        const float* bt = betas + (size_t)(betas_rows == 1 ? 0 : f) * NB;
        float acc = 0.f;
        for (int l = 0; l < NB; ++l) acc = fmaf(__ldg(bt + l), __ldg(sd + l), acc);
        v_shaped[(size_t)f * V3 + e] = vt + acc;
    }
}

// J[f,j,k] = sum_v Jr[j,v] * verts[f,v,k]; one CTA per (frame, joint group of 8); regressor_stride = 0 for a shared
// (J,V) regressor, J*V for a per-frame (F,J,V) one (vertices2jointsB, lbs.py:270-286).
constexpr int JG = 8;
__global__ void __launch_bounds__(256)
lbs_joints_kernel(int V, int J, size_t regressor_stride, const float* __restrict__ Jr_all, const float* __restrict__ verts_all,
                  float* __restrict__ joints_all) {
    const int f = blockIdx.x, j0 = blockIdx.y * JG;
    const float* verts = verts_all + (size_t)f * V * 3;
    const float* Jr = Jr_all + (size_t)f * regressor_stride;
    float acc[JG][3];
#pragma unroll
    for (int j = 0; j < JG; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;
    for (int v = threadIdx.x; v < V; v += 256) {
        const float x = __ldg(verts + 3 * v), y = __ldg(verts + 3 * v + 1), z = __ldg(verts + 3 * v + 2);
#pragma unroll
        for (int j = 0; j < JG; ++j) {
            const float w = (j0 + j < J) ? __ldg(Jr + (size_t)(j0 + j) * V + v) : 0.f;
            acc[j][0] = fmaf(w, x, acc[j][0]); acc[j][1] = fmaf(w, y, acc[j][1]); acc[j][2] = fmaf(w, z, acc[j][2]);
        }
    }
    __shared__ float red[8][JG * 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < JG; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float v = acc[j][k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
            if (lane == 0) red[warp][j * 3 + k] = v;
        }
    __syncthreads();
    if (threadIdx.x < JG * 3) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
        const int j = threadIdx.x / 3, k = threadIdx.x - 3 * j;
        if (j0 + j < J) joints_all[((size_t)f * J + j0 + j) * 3 + k] = s;
    }
}

// C[M,N] (+)= A[M,K] . B[K,N], all row-major fp32.  64x64 tile, 16-deep, 4x4 register blocking.
// accumulate = 1: C already holds the addend (v_shaped) and receives += (v_posed = pose_offsets + v_shaped).
constexpr int GM = 64, GN = 64, GK = 16;
__global__ void __launch_bounds__(256)
sgemm_acc_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, int accumulate) {
    __shared__ float As[GK][GM + 4];
    __shared__ float Bs[GK][GN];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < K; k0 += GK) {
        // A tile: 64 rows x 16 k -> 1024 elements, 4 per thread (k fastest in global -> coalesced along k)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e = threadIdx.x + 256 * r;
            const int mm = e >> 4, kk = e & 15;
            const int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < K) ? __ldg(A + (size_t)gm * lda + gk) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e = threadIdx.x + 256 * r;
            const int kk = e >> 6, nn = e & 63;
            const int gk = k0 + kk, gn = n0 + nn;
            Bs[kk][nn] = (gk < K && gn < N) ? __ldg(B + (size_t)gk * ldb + gn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float* c = C + (size_t)gm * ldc + gn;
            *c = accumulate ? (acc[i][j] + *c) : acc[i][j];
        }
    }
}

// batch_rigid_transform: one warp per frame.  chain[i] = chain[parent[i]] . [R_i | J_i - J_parent(i)]  (lbs.py:386-407),
// posed = chain[:, :3, 3], A = chain with last column -= chain[:3,:3] . J_i  (lbs.py:412-417).
constexpr int MAXJ = 64;
__global__ void __launch_bounds__(128)
lbs_rigid_kernel(int F, int J, const float* __restrict__ rot_mats, const float* __restrict__ joints,
                 const int* __restrict__ parents, float* __restrict__ posed_joints, float* __restrict__ A_out) {
    __shared__ float chain[4][MAXJ][12];
    __shared__ float tm[4][MAXJ][12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * 4 + warp;
    if (f >= F) return;
    const float* R = rot_mats + (size_t)f * J * 9;
    const float* Jt = joints + (size_t)f * J * 3;
    for (int j = lane; j < J; j += 32) {
        const int p = j > 0 ? __ldg(parents + j) : 0;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) tm[warp][j][r * 4 + c] = __ldg(R + j * 9 + r * 3 + c);
            tm[warp][j][r * 4 + 3] = j > 0 ? __ldg(Jt + j * 3 + r) - __ldg(Jt + p * 3 + r) : __ldg(Jt + r);
        }
    }
    __syncwarp();
    if (lane < 12) chain[warp][0][lane] = tm[warp][0][lane];
    __syncwarp();
    for (int i = 1; i < J; ++i) {
        const int p = __ldg(parents + i);
        if (lane < 12) {
            const int r = lane >> 2, c = lane & 3;
            const float* P = chain[warp][p] + r * 4;
            float v = P[0] * tm[warp][i][c];
            v = fmaf(P[1], tm[warp][i][4 + c], v);
            v = fmaf(P[2], tm[warp][i][8 + c], v);
            if (c == 3) v += P[3];
            chain[warp][i][lane] = v;
        }
        __syncwarp();
    }
    for (int j = lane; j < J; j += 32) {
        const float* Cj = chain[warp][j];
        const float jx = __ldg(Jt + j * 3), jy = __ldg(Jt + j * 3 + 1), jz = __ldg(Jt + j * 3 + 2);
        float* Ao = A_out + ((size_t)f * J + j) * 16;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            Ao[r * 4 + 0] = Cj[r * 4 + 0]; Ao[r * 4 + 1] = Cj[r * 4 + 1]; Ao[r * 4 + 2] = Cj[r * 4 + 2];
            const float corr = fmaf(Cj[r * 4 + 2], jz, fmaf(Cj[r * 4 + 1], jy, Cj[r * 4 + 0] * jx));
            Ao[r * 4 + 3] = Cj[r * 4 + 3] - corr;
            posed_joints[((size_t)f * J + j) * 3 + r] = Cj[r * 4 + 3];
        }
        Ao[12] = 0.f; Ao[13] = 0.f; Ao[14] = 0.f; Ao[15] = 1.f;
    }
}

// Skinning: thread = vertex (its J weights live in registers), CTA loops over a chunk of frames whose
// A matrices (J x 12 floats each) are broadcast from shared memory.  weights_stride = 0 for shared (V,J) weights,
// V*J for per-frame (F,V,J) weights (the interpolated garment weights of mesh_encoder.py:347,393).
constexpr int SKIN_FPB = 8;
template <int JT>
__global__ void __launch_bounds__(128)
lbs_skin_kernel(int F, int V, int J, size_t weights_stride, const float* __restrict__ v_posed, const float* __restrict__ A,
                const float* __restrict__ W, float* __restrict__ verts) {
    __shared__ __align__(16) float As[SKIN_FPB][MAXJ * 12];
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int f0 = blockIdx.y * SKIN_FPB;
    const int nf = min(SKIN_FPB, F - f0);
    const int JJ = JT > 0 ? JT : J;
    for (int e = threadIdx.x; e < nf * JJ * 12; e += blockDim.x) {
        const int ff = e / (JJ * 12), r = e - ff * (JJ * 12);
        const int j = r / 12, q = r - j * 12;
        As[ff][j * 12 + q] = __ldg(A + ((size_t)(f0 + ff) * JJ + j) * 16 + q);
    }
    __syncthreads();
    if (v >= V) return;
    float w[JT > 0 ? JT : 1];
    if (JT > 0 && weights_stride == 0) {
#pragma unroll
        for (int j = 0; j < JT; ++j) w[j] = __ldg(W + (size_t)v * JT + j);
    }
    for (int ff = 0; ff < nf; ++ff) {
        const size_t f = f0 + ff;
        const float* Wrow = W + f * weights_stride + (size_t)v * JJ;
        float T[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) T[q] = 0.f;
        if (JT > 0) {
#pragma unroll
            for (int j = 0; j < JT; ++j) {
                const float wj = weights_stride == 0 ? w[j] : __ldg(Wrow + j);
                const float4 a0 = *reinterpret_cast<const float4*>(&As[ff][j * 12]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[ff][j * 12 + 4]);
                const float4 a2 = *reinterpret_cast<const float4*>(&As[ff][j * 12 + 8]);
                T[0] = fmaf(wj, a0.x, T[0]); T[1] = fmaf(wj, a0.y, T[1]); T[2] = fmaf(wj, a0.z, T[2]); T[3] = fmaf(wj, a0.w, T[3]);
                T[4] = fmaf(wj, a1.x, T[4]); T[5] = fmaf(wj, a1.y, T[5]); T[6] = fmaf(wj, a1.z, T[6]); T[7] = fmaf(wj, a1.w, T[7]);
                T[8] = fmaf(wj, a2.x, T[8]); T[9] = fmaf(wj, a2.y, T[9]); T[10] = fmaf(wj, a2.z, T[10]); T[11] = fmaf(wj, a2.w, T[11]);
            }
        } else {
            for (int j = 0; j < JJ; ++j) {
                const float wj = __ldg(Wrow + j);
#pragma unroll
                for (int q = 0; q < 12; ++q) T[q] = fmaf(wj, As[ff][j * 12 + q], T[q]);
            }
        }
        const float* vp = v_posed + (f * V + v) * 3;
        const float x = __ldg(vp), y = __ldg(vp + 1), z = __ldg(vp + 2);
        float* o = verts + (f * V + v) * 3;
#pragma unroll
        for (int r = 0; r < 3; ++r)
            o[r] = fmaf(T[r * 4 + 2], z, fmaf(T[r * 4 + 1], y, T[r * 4 + 0] * x)) + T[r * 4 + 3];
    }
}

static int launch_skin(int F, int V, int J, size_t wstride, const float* v_posed, const float* A, const float* W, float* verts, cudaStream_t s) {
    if (J > MAXJ) return bad_arg("lbs skin: more than 64 joints");
    dim3 grid((V + 127) / 128, (F + SKIN_FPB - 1) / SKIN_FPB);
    if (J == 24) lbs_skin_kernel<24><<<grid, 128, 0, s>>>(F, V, J, wstride, v_posed, A, W, verts);
    else lbs_skin_kernel<0><<<grid, 128, 0, s>>>(F, V, J, wstride, v_posed, A, W, verts);
    return finish_launch("g4d lbs skin");
}

}  // namespace g4d

using namespace g4d;

// batch_rodrigues (lbs.py:312-346; identical copy at smplx/transfer_model/utils/pose_utils.py:62-99): (n,3) -> (n,3,3)
G4D_API int g4d_batch_rodrigues(int n, const float* rot_vecs, float* rot_mats, void* stream) {
    if (n < 0) return bad_arg("batch_rodrigues: negative n");
    if (n == 0) return 0;
    batch_rodrigues_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, rot_vecs, rot_mats);
    return finish_launch("g4d batch_rodrigues");
}

// blend_shapes (lbs.py:288-309): out[f, v, k] = sum_l betas[f, l] * shape_disps[v, k, l];  betas (F, NB), shape_disps (V, 3, NB),
// out (F, V, 3).
G4D_API int g4d_blend_shapes(int F, int V, int NB, const float* betas, const float* shape_disps, float* out, void* stream) {
    if (F < 0 || V < 0 || NB < 0) return bad_arg("blend_shapes: negative size");
    if (F == 0 || V == 0) return 0;
    if (!betas || !shape_disps || !out) return bad_arg("blend_shapes: null pointer");
    const int V3 = V * 3;
    dim3 grid((V3 + 255) / 256, F < 64 ? F : 64);
    lbs_shape_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(F, V3, NB, F, betas, nullptr, shape_disps, out);
    return finish_launch("g4d blend_shapes");
}

// vertices2joints / vertices2jointsB (lbs.py:251-286).  per_frame_regressor = 0: J_regressor (J,V); 1: (F,J,V).
G4D_API int g4d_vertices2joints(int F, int V, int J, int per_frame_regressor, const float* J_regressor, const float* vertices,
                                float* joints, void* stream) {
    if (F < 0 || V < 0 || J < 0) return bad_arg("vertices2joints: negative size");
    if (F == 0 || J == 0) return 0;
    dim3 grid(F, (J + JG - 1) / JG);
    lbs_joints_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(V, J, per_frame_regressor ? (size_t)J * V : 0, J_regressor, vertices, joints);
    return finish_launch("g4d vertices2joints");
}

// batch_rigid_transform (lbs.py:362-419).  parents: device int32 (J), parents[0] ignored.
G4D_API int g4d_batch_rigid_transform(int F, int J, const float* rot_mats, const float* joints, const int* parents,
                                      float* posed_joints, float* rel_transforms, void* stream) {
    if (F < 0 || J < 0) return bad_arg("batch_rigid_transform: negative size");
    if (J > MAXJ) return bad_arg("batch_rigid_transform: more than 64 joints");
    if (F == 0 || J == 0) return 0;
    lbs_rigid_kernel<<<(F + 3) / 4, 128, 0, (cudaStream_t)stream>>>(F, J, rot_mats, joints, parents, posed_joints, rel_transforms);
    return finish_launch("g4d batch_rigid_transform");
}

// Skinning tail (lbs.py:233-246): verts = (W . A) [v_posed; 1].  per_frame_weights = 1: W is (F,V,J).
G4D_API int g4d_lbs_skin(int F, int V, int J, int per_frame_weights, const float* v_posed, const float* A, const float* W,
                         float* verts, void* stream) {
    if (F < 0 || V < 0 || J < 0) return bad_arg("lbs_skin: negative size");
    if (F == 0 || V == 0) return 0;
    return launch_skin(F, V, J, per_frame_weights ? (size_t)V * J : 0, v_posed, A, W, verts, (cudaStream_t)stream);
}

G4D_API size_t g4d_lbs_workspace_bytes(int F, int V, int J) {
    // rot_mats F*J*9 | pose_feature F*(J-1)*9 | joints F*J*3 | A F*J*16 | v_posed F*V*3
    return sizeof(float) * ((size_t)F * J * 9 + (size_t)F * (J - 1) * 9 + (size_t)F * J * 3 + (size_t)F * J * 16 + (size_t)F * V * 3);
}

// lbs() (lbs.py:152-248).  F = max(betas_rows, pose rows) frames; betas_rows is 1 (broadcast) or F.
// pose: (F,J,3) axis-angle when pose2rot, else (F,J,3,3).  shapedirs (V,3,NB), posedirs ((J-1)*9, V*3),
// J_regressor (J,V), parents device int32 (J), lbs_weights (V,J).  Outputs verts (F,V,3), joints (F,J,3).
// ws: device scratch of g4d_lbs_workspace_bytes(F,V,J).
G4D_API int g4d_lbs(int F, int V, int J, int NB, int betas_rows, int pose2rot, const float* betas, const float* pose,
                    const float* v_template, const float* shapedirs, const float* posedirs, const float* J_regressor,
                    const int* parents, const float* lbs_weights, float* verts, float* joints, void* ws, size_t ws_bytes,
                    void* stream) {
    if (F < 0 || V <= 0 || J <= 0 || NB < 0) return bad_arg("lbs: bad size");
    if (J > MAXJ) return bad_arg("lbs: more than 64 joints");
    if (betas_rows != 1 && betas_rows != F) return bad_arg("lbs: betas rows must be 1 or F");
    if (F == 0) return 0;
    if (ws_bytes < g4d_lbs_workspace_bytes(F, V, J) || !ws) return bad_arg("lbs: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    float* rot = (float*)ws;
    float* pf = rot + (size_t)F * J * 9;
    float* Jrest = pf + (size_t)F * (J - 1) * 9;
    float* A = Jrest + (size_t)F * J * 3;
    float* v_posed = A + (size_t)F * J * 16;
    const int P = (J - 1) * 9, V3 = V * 3;
    lbs_pose_kernel<<<(F * J + 127) / 128, 128, 0, s>>>(F, J, pose2rot, pose, rot, pf);
    {
        dim3 grid((V3 + 255) / 256, F < 64 ? F : 64);
        lbs_shape_kernel<<<grid, 256, 0, s>>>(F, V3, NB, betas_rows, betas, v_template, shapedirs, v_posed);   // v_shaped for now
    }
    {
        dim3 grid(F, (J + JG - 1) / JG);
        lbs_joints_kernel<<<grid, 256, 0, s>>>(V, J, 0, J_regressor, v_posed, Jrest);
    }
    if (P > 0) {
        dim3 grid((V3 + GN - 1) / GN, (F + GM - 1) / GM);
        sgemm_acc_kernel<<<grid, 256, 0, s>>>(F, V3, P, pf, P, posedirs, V3, v_posed, V3, 1);
    }
    lbs_rigid_kernel<<<(F + 3) / 4, 128, 0, s>>>(F, J, rot, Jrest, parents, joints, A);
    int rc = finish_launch("g4d lbs (pose/shape/joints/blend/rigid)", P > 0 ? 5 : 4);
    if (rc) return rc;
    return launch_skin(F, V, J, 0, v_posed, A, lbs_weights, verts, s);
}
