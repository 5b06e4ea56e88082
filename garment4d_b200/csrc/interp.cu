// three_nn / three_interpolate (feature propagation) for sm_100a.
//
// Replaces (reference, modules/pointnet2/pointnet2/src/interpolate_gpu.cu):
//   three_nn_kernel_fast               :9-52    (thread per unknown point, known cloud re-read from L1)
//   three_interpolate_kernel_fast      :77-97   (thread per (c, n): idx/weight re-read once per channel)
//   three_interpolate_grad_kernel_fast :120-142
// Here the known cloud is staged in shared memory tiles, and interpolation loads idx/weight once per
// point and loops over a channel slab.  Arithmetic order matches the reference build bit for bit:
// d = fma(dz,dz,fma(dx,dx,dy*dy)) with dx = u - k, strict '<' insertion (lowest index wins ties),
// out = fma(w2,p2, fma(w0,p0, w1*p1)).
#include "common.cuh"

namespace g4d {

constexpr int NN_TILE = 1024;

__global__ void __launch_bounds__(256)
three_nn_kernel(int n, int m, const float* __restrict__ unknown_all, const float* __restrict__ known_all,
                float* __restrict__ dist2_all, int* __restrict__ idx_all) {
    __shared__ float tile[NN_TILE * 3];
    const size_t bi = blockIdx.y;
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = pt < n;
    const float* u = unknown_all + (bi * n + (live ? pt : 0)) * 3;
    const float* known = known_all + bi * (size_t)m * 3;
    const float ux = __ldg(u), uy = __ldg(u + 1), uz = __ldg(u + 2);
    // the reference keeps running bests in double initialised to 1e40; a float distance compares
    // against them exactly like against +inf, and (float)1e40 == +inf is what it stores when m < 3.
    float best1 = INFINITY, best2 = INFINITY, best3 = INFINITY;
    int besti1 = 0, besti2 = 0, besti3 = 0;
    for (int base = 0; base < m; base += NN_TILE) {
        const int tn = min(NN_TILE, m - base);
        __syncthreads();
        for (int e = threadIdx.x; e < tn * 3; e += blockDim.x) tile[e] = __ldg(known + (size_t)base * 3 + e);
        __syncthreads();
        for (int k = 0; k < tn; ++k) {
            const float d = sqdist_ref(ux - tile[3 * k], uy - tile[3 * k + 1], uz - tile[3 * k + 2]);   // broadcast reads
            if (d < best1) {
                best3 = best2; besti3 = besti2;
                best2 = best1; besti2 = besti1;
                best1 = d; besti1 = base + k;
            } else if (d < best2) {
                best3 = best2; besti3 = besti2;
                best2 = d; besti2 = base + k;
            } else if (d < best3) {
                best3 = d; besti3 = base + k;
            }
        }
    }
    if (live) {
        float* od = dist2_all + (bi * n + pt) * 3;
        int* oi = idx_all + (bi * n + pt) * 3;
        od[0] = best1; od[1] = best2; od[2] = best3;
        oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
    }
}

__global__ void three_interpolate_kernel(int c, int m, int n, const float* __restrict__ points, const int* __restrict__ idx,
                                         const float* __restrict__ weight, float* __restrict__ out) {
    const size_t bi = blockIdx.z;
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= n) return;
    const int* id = idx + (bi * n + pt) * 3;
    const float* w = weight + (bi * n + pt) * 3;
    const int i0 = __ldg(id), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y) {
        const float* p = points + (bi * c + ci) * (size_t)m;
        out[(bi * c + ci) * (size_t)n + pt] = __fmaf_rn(w2, __ldg(p + i2), __fmaf_rn(w0, __ldg(p + i0), __fmul_rn(w1, __ldg(p + i1))));
    }
}

__global__ void three_interpolate_grad_kernel(int c, int n, int m, const float* __restrict__ grad_out, const int* __restrict__ idx,
                                              const float* __restrict__ weight, float* __restrict__ grad_points) {
    const size_t bi = blockIdx.z;
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= n) return;
    const int* id = idx + (bi * n + pt) * 3;
    const float* w = weight + (bi * n + pt) * 3;
    const int i0 = __ldg(id), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
    const float w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    for (int ci = blockIdx.y; ci < c; ci += gridDim.y) {
        const float g = __ldg(grad_out + (bi * c + ci) * (size_t)n + pt);
        float* gp = grad_points + (bi * c + ci) * (size_t)m;
        atomicAdd(gp + i0, g * w0);
        atomicAdd(gp + i1, g * w1);
        atomicAdd(gp + i2, g * w2);
    }
}

}  // namespace g4d

using namespace g4d;

// three_nn_kernel_launcher_fast (interpolate_gpu.h:16-17).  Writes SQUARED distances, like the reference kernel.
G4D_API int g4d_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2, int* idx, void* stream) {
    if (b < 0 || n < 0 || m < 0) return bad_arg("three_nn: negative size");
    if (b == 0 || n == 0) return 0;
    if (!unknown || !dist2 || !idx || (m > 0 && !known)) return bad_arg("three_nn: null pointer");
    dim3 grid((n + 255) / 256, b);
    three_nn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, unknown, known, dist2, idx);
    return finish_launch("g4d three_nn");
}

// three_interpolate_kernel_launcher_fast (interpolate_gpu.h:21-22)
G4D_API int g4d_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx, const float* weight, float* out, void* stream) {
    if (b < 0 || c < 0 || n < 0) return bad_arg("three_interpolate: negative size");
    if (b == 0 || c == 0 || n == 0) return 0;
    dim3 grid((n + 255) / 256, c < 16 ? c : 16, b);
    three_interpolate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, m, n, points, idx, weight, out);
    return finish_launch("g4d three_interpolate");
}

// three_interpolate_grad_kernel_launcher_fast (interpolate_gpu.h:27-28); grad_points pre-zeroed by the caller.
G4D_API int g4d_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx, const float* weight, float* grad_points, void* stream) {
    if (b < 0 || c < 0 || n < 0) return bad_arg("three_interpolate_grad: negative size");
    if (b == 0 || c == 0 || n == 0) return 0;
    dim3 grid((n + 255) / 256, c < 16 ? c : 16, b);
    three_interpolate_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, m, grad_out, idx, weight, grad_points);
    return finish_launch("g4d three_interpolate_grad");
}

// ---------------------------------------------------------------------------------------------------------
// y[b,c,:] = max(y[b,c,:] + bias[c], 0) in place, one pass (channel-major (b,c,n)).  Epilogue of the feature-propagation
// 1x1 convolutions that stay on the library GEMM (eval-mode BatchNorm folded into weight/bias): replaces the separate
// bias-add and ReLU passes torch issues after cudnn/cutlass convolutions.
namespace g4d {
__global__ void __launch_bounds__(256)
bias_relu_kernel(int c, long long n4, int relu, float4* __restrict__ y, const float* __restrict__ bias) {
    const long long row = blockIdx.y;                     // b * c + channel
    const float bv = __ldg(bias + (int)(row % c));
    float4* p = y + row * n4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = p[i];
        v.x += bv; v.y += bv; v.z += bv; v.w += bv;
        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        p[i] = v;
    }
}
__global__ void bias_relu_scalar_kernel(int c, long long n, int relu, float* __restrict__ y, const float* __restrict__ bias) {
    const long long row = blockIdx.y;
    const float bv = __ldg(bias + (int)(row % c));
    float* p = y + row * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = p[i] + bv;
        p[i] = relu ? fmaxf(v, 0.f) : v;
    }
}
}  // namespace g4d

G4D_API int g4d_bias_relu_inplace(int b, int c, long long n, float* y, const float* bias, int relu, void* stream) {
    if (b < 0 || c < 0 || n < 0) return bad_arg("bias_relu_inplace: negative size");
    if (b == 0 || c == 0 || n == 0) return 0;
    if (!y || !bias) return bad_arg("bias_relu_inplace: null pointer");
    if ((long long)b * c > 65535ll * 32768ll) return bad_arg("bias_relu_inplace: too many rows");
    const long long rows = (long long)b * c;
    if (rows > 2147483647ll) return bad_arg("bias_relu_inplace: too many rows");
    if (n % 4 == 0 && ((uintptr_t)y & 15) == 0) {
        const long long n4 = n / 4;
        dim3 grid((unsigned)((n4 + 255) / 256 > 64 ? 64 : (n4 + 255) / 256), (unsigned)rows);
        if (rows > 65535) { // fold rows into x when the y-dimension would overflow
            return bad_arg("bias_relu_inplace: b*c > 65535 rows not supported");
        }
        g4d::bias_relu_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n4, relu, (float4*)y, bias);
    } else {
        if (rows > 65535) return bad_arg("bias_relu_inplace: b*c > 65535 rows not supported");
        dim3 grid((unsigned)((n + 255) / 256 > 64 ? 64 : (n + 255) / 256), (unsigned)rows);
        g4d::bias_relu_scalar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, relu, y, bias);
    }
    return finish_launch("g4d bias_relu_inplace");
}

// ---------------------------------------------------------------------------------------------------------
// Front half of PointnetFPModule.forward in one pass (eval route; pointnet2_modules.py:138-152):
//     dist_recip = 1 / (dist + 1e-8) ; norm = sum(dist_recip) ; weight = dist_recip / norm        (5 torch kernels)
//     interpolated = three_interpolate(known_feats, idx, weight)                                  (interpolate_gpu.cu:77-97)
//     new_features = cat([interpolated, unknow_feats], dim=1)                                     (1 torch kernel)
// One thread per unknown point: the weights are computed once (correctly rounded sqrt / divide in torch's order, so they
// equal the torch tensors bit for bit), then the thread walks a slab of output channels; the skip channels are a
// coalesced copy.  Same FMUL/FFMA order as three_interpolate_kernel.
namespace g4d {
__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__half* p, float v) { *p = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)); }

// out element (bi, ci, pt) lives at ci * out_cs + bi * out_bs + pt: (b, c, n) fp32 for the reference layout, or (c, b, n) fp16 so
// that the following 1x1 convolutions are ONE (Cout x Cin) . (Cin x b*n) GEMM over the whole batch.
template <typename OutT>
__global__ void __launch_bounds__(256)
fp_interp_concat_kernel(int c2, int c1, int m, int n, const float* __restrict__ dist2, const int* __restrict__ idx,
                        const float* __restrict__ known_feats, const float* __restrict__ skip, OutT* __restrict__ out,
                        long long out_cs, long long out_bs) {
    const size_t bi = blockIdx.z;
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= n) return;
    const int ctot = c2 + c1;
    int i0 = 0, i1 = 0, i2 = 0;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
    if ((int)blockIdx.y < c2) {                     // this channel slab holds at least one interpolated channel
        const int* id = idx + (bi * n + pt) * 3;
        const float* d2 = dist2 + (bi * n + pt) * 3;
        i0 = __ldg(id); i1 = __ldg(id + 1); i2 = __ldg(id + 2);
        const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2)), 1e-8f));
        const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + 1)), 1e-8f));
        const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + 2)), 1e-8f));
        const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
        w0 = __fdiv_rn(r0, norm); w1 = __fdiv_rn(r1, norm); w2 = __fdiv_rn(r2, norm);
    }
    OutT* o = out + bi * out_bs + pt;
    for (int ci = blockIdx.y; ci < ctot; ci += gridDim.y) {
        float v;
        if (ci < c2) {
            const float* p = known_feats + (bi * c2 + ci) * (size_t)m;
            v = __fmaf_rn(w2, __ldg(p + i2), __fmaf_rn(w0, __ldg(p + i0), __fmul_rn(w1, __ldg(p + i1))));
        } else {
            v = __ldg(skip + (bi * c1 + (ci - c2)) * (size_t)n + pt);
        }
        store_out(o + ci * out_cs, v);
    }
}

// Same result as fp_interp_concat_kernel<__half> when the known features are given fp16 POINT-major (b, m, c2) -- the copy every
// fused level already emits for the next gather: the three taps are three contiguous rows, read 8 channels per 16-byte load
// (the channel-major form gathers 3 scattered words per channel and was instruction-bound: 64 % SM busy at 18 % of DRAM).
__global__ void __launch_bounds__(256)
fp_interp_concat_pm_kernel(int c2, int c1, int m, int n, const float* __restrict__ dist2, const int* __restrict__ idx,
                           const __half* __restrict__ known_pm, const float* __restrict__ skip, __half* __restrict__ out,
                           long long out_cs, long long out_bs) {
    const size_t bi = blockIdx.z;
    const int pt = blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= n) return;
    __half* o = out + bi * out_bs + pt;
    const int nch8 = c2 >> 3;
    if ((int)blockIdx.y < nch8) {
        const int* id = idx + (bi * n + pt) * 3;
        const float* d2 = dist2 + (bi * n + pt) * 3;
        const int i0 = __ldg(id), i1 = __ldg(id + 1), i2 = __ldg(id + 2);
        const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2)), 1e-8f));
        const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + 1)), 1e-8f));
        const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(d2 + 2)), 1e-8f));
        const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
        const float w0 = __fdiv_rn(r0, norm), w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm);
        const uint4* q0 = reinterpret_cast<const uint4*>(known_pm + (bi * m + i0) * (size_t)c2);
        const uint4* q1 = reinterpret_cast<const uint4*>(known_pm + (bi * m + i1) * (size_t)c2);
        const uint4* q2 = reinterpret_cast<const uint4*>(known_pm + (bi * m + i2) * (size_t)c2);
#pragma unroll 2
        for (int ch8 = blockIdx.y; ch8 < nch8; ch8 += gridDim.y) {
            const uint4 a = __ldg(q0 + ch8), b = __ldg(q1 + ch8), c = __ldg(q2 + ch8);
            const __half2* ha = reinterpret_cast<const __half2*>(&a);
            const __half2* hb = reinterpret_cast<const __half2*>(&b);
            const __half2* hc = reinterpret_cast<const __half2*>(&c);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f0 = __half22float2(ha[e]), f1 = __half22float2(hb[e]), f2 = __half22float2(hc[e]);
                store_out(o + (size_t)(ch8 * 8 + 2 * e) * out_cs, __fmaf_rn(w2, f2.x, __fmaf_rn(w0, f0.x, __fmul_rn(w1, f1.x))));
                store_out(o + (size_t)(ch8 * 8 + 2 * e + 1) * out_cs, __fmaf_rn(w2, f2.y, __fmaf_rn(w0, f0.y, __fmul_rn(w1, f1.y))));
            }
        }
    }
    for (int ci = blockIdx.y; ci < c1; ci += gridDim.y)
        store_out(o + (size_t)(c2 + ci) * out_cs, __ldg(skip + (bi * c1 + ci) * (size_t)n + pt));
}

// Point-major form of the FP-module front half: known_pm (b, m, c2) and skip_pm (b, n, c1) fp16 point-major (the copies the
// fused levels emit) -> out (b*n, c2 + c1) fp16 row-major, the activation operand of `x @ W^T`.  A warp handles 4 points x 8
// lanes; lane j of a point moves the 16-byte chunks j, j+8, ... of its row, so every load and store instruction of the warp
// covers full 128-byte segments, and a block writes one contiguous 32 x (c2+c1) x 2 byte region (the (C, b*n) layout wrote
// 64-byte pieces 480 KB apart: 0.9 TB/s).  Weights: lanes 0..2 of a point each take one tap (one sqrt, two divides, the same
// correctly rounded operations in the same order as everywhere else), results shared by shuffles.
__global__ void __launch_bounds__(256)
fp_interp_concat_rows_kernel(int c2, int c1, int m, int n, const float* __restrict__ dist2, const int* __restrict__ idx,
                             const __half* __restrict__ known_pm, const __half* __restrict__ skip_pm, __half* __restrict__ out) {
    const size_t bi = blockIdx.y;
    const int lane = threadIdx.x & 31, j = lane & 7, gbase = lane & ~7;
    const int pt = blockIdx.x * 32 + (threadIdx.x >> 3);
    const bool live = pt < n;
    float r = 0.f;
    int it = 0;
    if (live && j < 3) {
        r = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + (bi * n + pt) * 3 + j)), 1e-8f));
        it = __ldg(idx + (bi * n + pt) * 3 + j);
    }
    const float r0 = __shfl_sync(0xFFFFFFFFu, r, gbase), r1 = __shfl_sync(0xFFFFFFFFu, r, gbase + 1), r2 = __shfl_sync(0xFFFFFFFFu, r, gbase + 2);
    const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
    const float w0 = __fdiv_rn(r0, norm), w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm);
    const int i0 = __shfl_sync(0xFFFFFFFFu, it, gbase), i1 = __shfl_sync(0xFFFFFFFFu, it, gbase + 1), i2 = __shfl_sync(0xFFFFFFFFu, it, gbase + 2);
    if (!live) return;
    const int ctot = c2 + c1;
    uint4* orow = reinterpret_cast<uint4*>(out + (bi * n + pt) * (size_t)ctot);
    const uint4* q0 = reinterpret_cast<const uint4*>(known_pm + (bi * m + i0) * (size_t)c2);
    const uint4* q1 = reinterpret_cast<const uint4*>(known_pm + (bi * m + i1) * (size_t)c2);
    const uint4* q2 = reinterpret_cast<const uint4*>(known_pm + (bi * m + i2) * (size_t)c2);
    const int nch2 = c2 >> 3, nch1 = c1 >> 3;
#pragma unroll 2
    for (int ch = j; ch < nch2; ch += 8) {
        const uint4 a = __ldg(q0 + ch), b = __ldg(q1 + ch), c = __ldg(q2 + ch);
        const __half2* ha = reinterpret_cast<const __half2*>(&a);
        const __half2* hb = reinterpret_cast<const __half2*>(&b);
        const __half2* hc = reinterpret_cast<const __half2*>(&c);
        uint4 o;
        __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f0 = __half22float2(ha[e]), f1 = __half22float2(hb[e]), f2 = __half22float2(hc[e]);
            const float vx = __fmaf_rn(w2, f2.x, __fmaf_rn(w0, f0.x, __fmul_rn(w1, f1.x)));
            const float vy = __fmaf_rn(w2, f2.y, __fmaf_rn(w0, f0.y, __fmul_rn(w1, f1.y)));
            ho[e] = __floats2half2_rn(fminf(fmaxf(vx, -65504.f), 65504.f), fminf(fmaxf(vy, -65504.f), 65504.f));
        }
        orow[ch] = o;
    }
    if (nch1) {
        const uint4* srow = reinterpret_cast<const uint4*>(skip_pm + (bi * n + pt) * (size_t)c1);
        for (int ch = j; ch < nch1; ch += 8) orow[nch2 + ch] = __ldg(srow + ch);
    }
}

// y (rows, c) fp16 row-major, in place: y[r, ch] = act(y[r, ch] + bias[ch]); c % 8 == 0.  FIXED_CH: the grid stride is a
// multiple of c/8 (always the case when 256 % (c/8) == 0), so a thread sees the same 8 channels in every iteration and keeps
// their biases in registers (no modulo, no bias loads in the loop).
template <bool FIXED_CH>
__global__ void __launch_bounds__(256)
bias_relu_rows_h_kernel(long long total8, int c8, int relu, uint4* __restrict__ y, const float* __restrict__ bias) {
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    float bv[8];
    if (FIXED_CH) {
        const int ch = (int)(i0 % c8) * 8;
        *reinterpret_cast<float4*>(bv) = __ldg(reinterpret_cast<const float4*>(bias + ch));
        *reinterpret_cast<float4*>(bv + 4) = __ldg(reinterpret_cast<const float4*>(bias + ch + 4));
    }
    for (long long i = i0; i < total8; i += (long long)gridDim.x * blockDim.x) {
        if (!FIXED_CH) {
            const int ch = (int)(i % c8) * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) bv[k] = __ldg(bias + ch + k);
        }
        uint4 v = y[i];
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 f = __half22float2(h[k]);
            f.x += bv[2 * k]; f.y += bv[2 * k + 1];
            if (relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
            h[k] = __floats2half2_rn(fminf(f.x, 65504.f), fminf(f.y, 65504.f));
        }
        y[i] = v;
    }
}

// Last layer of the point-major route: yin (b, n, c) fp32 pre-activations -> out_cm (b, c, n) fp32 = act(yin + bias) (the
// reference layout) and, optionally, the same values fp16 point-major out_pm (b, n, c).  32 x 32 tiles through shared memory.
__global__ void __launch_bounds__(256)
bias_relu_rows_unpack_kernel(int c, int n, int relu, const float* __restrict__ yin, const float* __restrict__ bias,
                             float* __restrict__ out_cm, __half* __restrict__ out_pm) {
    __shared__ float tile[32][33];
    const size_t bi = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int pt = p0 + ty + 8 * r, ch = c0 + tx;
        float v = 0.f;
        if (ch < c && pt < n) {
            v = __ldg(yin + (bi * n + pt) * (size_t)c + ch) + __ldg(bias + ch);
            if (relu) v = fmaxf(v, 0.f);
            if (out_pm) out_pm[(bi * n + pt) * (size_t)c + ch] = __float2half_rn(fminf(v, 65504.f));
        }
        tile[ty + 8 * r][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int ch = c0 + ty + 8 * r, pt = p0 + tx;
        if (ch < c && pt < n) out_cm[(bi * c + ch) * (size_t)n + pt] = tile[tx][ty + 8 * r];
    }
}

// y (c, len) fp16: y[ch, :] = act(y[ch, :] + bias[ch]) in place, 8 halves per thread (len % 8 == 0)
__global__ void __launch_bounds__(256)
bias_relu_h_kernel(long long len8, int relu, uint4* __restrict__ y, const float* __restrict__ bias) {
    const float bv = __ldg(bias + blockIdx.y);
    uint4* p = y + (size_t)blockIdx.y * len8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len8; i += (long long)gridDim.x * blockDim.x) {
        uint4 v = p[i];
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 f = __half22float2(h[k]);
            f.x += bv; f.y += bv;
            if (relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); }
            h[k] = __floats2half2_rn(fminf(f.x, 65504.f), fminf(f.y, 65504.f));
        }
        p[i] = v;
    }
}

__device__ __forceinline__ float load_in(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_in(const __half* p) { return __half2float(*p); }

// Last layer of the batched-GEMM route: yin (c, b, n) (fp32 or fp16, pre-activation) -> out_cm (b, c, n) fp32 = act(yin + bias)
// and, optionally, the same values fp16 point-major out_pm (b, n, c).  32 x 32 (channel x point) tiles through shared memory.
template <typename InT>
__global__ void __launch_bounds__(256)
bias_relu_unpack_kernel(int b, int c, int n, int relu, const InT* __restrict__ yin, const float* __restrict__ bias,
                        float* __restrict__ out_cm, __half* __restrict__ out_pm) {
    __shared__ float tile[32][33];
    const size_t bi = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int ch = c0 + ty + 8 * r, pt = p0 + tx;
        float v = 0.f;
        if (ch < c && pt < n) {
            v = load_in(yin + ((size_t)ch * b + bi) * n + pt) + __ldg(bias + ch);
            if (relu) v = fmaxf(v, 0.f);
            out_cm[(bi * c + ch) * (size_t)n + pt] = v;
        }
        tile[ty + 8 * r][tx] = v;
    }
    if (!out_pm) return;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int pt = p0 + ty + 8 * r, ch = c0 + tx;
        if (ch < c && pt < n) out_pm[(bi * n + pt) * (size_t)c + ch] = __float2half_rn(fminf(tile[tx][ty + 8 * r], 65504.f));
    }
}

// y (b,c,n) fp32 channel-major: y = act(y + bias[c]) in place AND the same values as fp16 point-major (b,n,c) -- the gather
// layout of the fused tcgen05 kernels (replaces a bias/ReLU pass + transpose().to(half).contiguous() = 3 passes).
// 32 x 32 (channel x point) tiles through shared memory; both the fp32 read-modify-write and the fp16 write are coalesced.
__global__ void __launch_bounds__(256)
bias_relu_pm_kernel(int c, int n, int relu, float* __restrict__ y, const float* __restrict__ bias, __half* __restrict__ out_pm) {
    __shared__ float tile[32][33];
    const size_t bi = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int ch = c0 + ty + 8 * r, pt = p0 + tx;
        float v = 0.f;
        if (ch < c && pt < n) {
            float* q = y + (bi * c + ch) * (size_t)n + pt;
            v = *q + __ldg(bias + ch);
            if (relu) v = fmaxf(v, 0.f);
            *q = v;
        }
        tile[ty + 8 * r][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int pt = p0 + ty + 8 * r, ch = c0 + tx;
        if (ch < c && pt < n) out_pm[(bi * n + pt) * (size_t)c + ch] = __float2half_rn(fminf(tile[tx][ty + 8 * r], 65504.f));
    }
}
}  // namespace g4d

static int fp_interp_concat_check(int b, int c2, int c1, int m, int n, const float* dist2, const int* idx, const float* known_feats,
                                  const float* skip, const void* out) {
    if (b < 0 || c2 < 0 || c1 < 0 || n < 0 || m < 0) return bad_arg("fp_interp_concat: negative size");
    if (b == 0 || n == 0 || c2 + c1 == 0) return -1;
    if (!out || (c2 > 0 && (!dist2 || !idx || !known_feats || m == 0)) || (c1 > 0 && !skip)) return bad_arg("fp_interp_concat: null pointer");
    if (b > 65535) return bad_arg("fp_interp_concat: b > 65535");
    return 0;
}

G4D_API int g4d_fp_interp_concat(int b, int c2, int c1, int m, int n, const float* dist2, const int* idx, const float* known_feats,
                                 const float* skip, float* out, void* stream) {
    const int rc = fp_interp_concat_check(b, c2, c1, m, n, dist2, idx, known_feats, skip, out);
    if (rc) return rc < 0 ? 0 : rc;
    const int ctot = c2 + c1;
    dim3 grid((n + 255) / 256, ctot < 16 ? ctot : 16, b);
    fp_interp_concat_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(c2, c1, m, n, dist2, idx, known_feats, skip, out,
                                                                           (long long)n, (long long)ctot * n);
    return finish_launch("g4d fp_interp_concat");
}

G4D_API int g4d_fp_interp_concat_cbn_h(int b, int c2, int c1, int m, int n, const float* dist2, const int* idx, const float* known_feats,
                                       const float* skip, void* out_h, void* stream) {
    const int rc = fp_interp_concat_check(b, c2, c1, m, n, dist2, idx, known_feats, skip, out_h);
    if (rc) return rc < 0 ? 0 : rc;
    const int ctot = c2 + c1;
    dim3 grid((n + 255) / 256, ctot < 16 ? ctot : 16, b);
    fp_interp_concat_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(c2, c1, m, n, dist2, idx, known_feats, skip, (__half*)out_h,
                                                                            (long long)b * n, (long long)n);
    return finish_launch("g4d fp_interp_concat_cbn_h");
}

G4D_API int g4d_fp_interp_concat_pm_cbn_h(int b, int c2, int c1, int m, int n, const float* dist2, const int* idx, const void* known_pm_h,
                                          const float* skip, void* out_h, void* stream) {
    const int rc = fp_interp_concat_check(b, c2, c1, m, n, dist2, idx, (const float*)known_pm_h, skip, out_h);
    if (rc) return rc < 0 ? 0 : rc;
    if (c2 % 8 || ((uintptr_t)known_pm_h & 15)) return bad_arg("fp_interp_concat_pm: c2 must be a multiple of 8 and known_pm 16-byte aligned");
    const int work = (c2 / 8 > c1 ? c2 / 8 : c1);
    dim3 grid((n + 255) / 256, work < 4 ? (work < 1 ? 1 : work) : 4, b);
    fp_interp_concat_pm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c2, c1, m, n, dist2, idx, (const __half*)known_pm_h, skip, (__half*)out_h,
                                                                     (long long)b * n, (long long)n);
    return finish_launch("g4d fp_interp_concat_pm_cbn_h");
}

G4D_API int g4d_fp_interp_concat_rows_h(int b, int c2, int c1, int m, int n, const float* dist2, const int* idx, const void* known_pm_h,
                                        const void* skip_pm_h, void* out_rows_h, void* stream) {
    if (b < 0 || c2 < 0 || c1 < 0 || n < 0 || m < 0) return bad_arg("fp_interp_concat_rows: negative size");
    if (b == 0 || n == 0 || c2 + c1 == 0) return 0;
    if (!out_rows_h || (c2 > 0 && (!dist2 || !idx || !known_pm_h || m == 0)) || (c1 > 0 && !skip_pm_h)) return bad_arg("fp_interp_concat_rows: null pointer");
    if (c2 % 8 || c1 % 8) return bad_arg("fp_interp_concat_rows: c2 and c1 must be multiples of 8");
    if (((uintptr_t)known_pm_h | (uintptr_t)skip_pm_h | (uintptr_t)out_rows_h) & 15) return bad_arg("fp_interp_concat_rows: pointers must be 16-byte aligned");
    if (b > 65535) return bad_arg("fp_interp_concat_rows: b > 65535");
    dim3 grid((n + 31) / 32, b);
    fp_interp_concat_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c2, c1, m, n, dist2, idx, (const __half*)known_pm_h,
                                                                       (const __half*)skip_pm_h, (__half*)out_rows_h);
    return finish_launch("g4d fp_interp_concat_rows_h");
}

G4D_API int g4d_bias_relu_rows_h(long long rows, int c, void* y_h, const float* bias, int relu, void* stream) {
    if (rows < 0 || c < 0) return bad_arg("bias_relu_rows_h: negative size");
    if (rows == 0 || c == 0) return 0;
    if (!y_h || !bias) return bad_arg("bias_relu_rows_h: null pointer");
    if (c % 8 || ((uintptr_t)y_h & 15)) return bad_arg("bias_relu_rows_h: c must be a multiple of 8 and y 16-byte aligned");
    const long long total8 = rows * (c / 8);
    long long blocks = (total8 + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (256 % (c / 8) == 0 && ((uintptr_t)bias & 15) == 0)
        bias_relu_rows_h_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(total8, c / 8, relu, (uint4*)y_h, bias);
    else
        bias_relu_rows_h_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(total8, c / 8, relu, (uint4*)y_h, bias);
    return finish_launch("g4d bias_relu_rows_h");
}

G4D_API int g4d_bias_relu_rows_unpack(int b, int c, int n, const float* yin_rows, const float* bias, int relu, float* out_cm, void* out_pm,
                                      void* stream) {
    if (b < 0 || c < 0 || n < 0) return bad_arg("bias_relu_rows_unpack: negative size");
    if (b == 0 || c == 0 || n == 0) return 0;
    if (!yin_rows || !bias || !out_cm) return bad_arg("bias_relu_rows_unpack: null pointer");
    if (b > 65535 || (c + 31) / 32 > 65535) return bad_arg("bias_relu_rows_unpack: b or c/32 > 65535");
    dim3 grid((n + 31) / 32, (c + 31) / 32, b);
    bias_relu_rows_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, relu, yin_rows, bias, out_cm, (__half*)out_pm);
    return finish_launch("g4d bias_relu_rows_unpack");
}

G4D_API int g4d_bias_relu_h(int c, long long len, void* y_h, const float* bias, int relu, void* stream) {
    if (c < 0 || len < 0) return bad_arg("bias_relu_h: negative size");
    if (c == 0 || len == 0) return 0;
    if (!y_h || !bias) return bad_arg("bias_relu_h: null pointer");
    if (len % 8 || ((uintptr_t)y_h & 15)) return bad_arg("bias_relu_h: row length must be a multiple of 8 and y 16-byte aligned");
    if (c > 65535) return bad_arg("bias_relu_h: c > 65535");
    const long long len8 = len / 8;
    dim3 grid((unsigned)((len8 + 255) / 256 > 256 ? 256 : (len8 + 255) / 256), (unsigned)c);
    bias_relu_h_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(len8, relu, (uint4*)y_h, bias);
    return finish_launch("g4d bias_relu_h");
}

G4D_API int g4d_bias_relu_unpack(int b, int c, int n, const void* yin_cbn, int in_half, const float* bias, int relu, float* out_cm,
                                 void* out_pm, void* stream) {
    if (b < 0 || c < 0 || n < 0) return bad_arg("bias_relu_unpack: negative size");
    if (b == 0 || c == 0 || n == 0) return 0;
    if (!yin_cbn || !bias || !out_cm) return bad_arg("bias_relu_unpack: null pointer");
    if (b > 65535 || (c + 31) / 32 > 65535) return bad_arg("bias_relu_unpack: b or c/32 > 65535");
    dim3 grid((n + 31) / 32, (c + 31) / 32, b);
    if (in_half) bias_relu_unpack_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(b, c, n, relu, (const __half*)yin_cbn, bias, out_cm, (__half*)out_pm);
    else bias_relu_unpack_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(b, c, n, relu, (const float*)yin_cbn, bias, out_cm, (__half*)out_pm);
    return finish_launch("g4d bias_relu_unpack");
}

G4D_API int g4d_bias_relu_pm(int b, int c, int n, float* y, const float* bias, int relu, void* out_pm, void* stream) {
    if (b < 0 || c < 0 || n < 0) return bad_arg("bias_relu_pm: negative size");
    if (b == 0 || c == 0 || n == 0) return 0;
    if (!y || !bias || !out_pm) return bad_arg("bias_relu_pm: null pointer");
    if (b > 65535 || (c + 31) / 32 > 65535) return bad_arg("bias_relu_pm: b or c/32 > 65535");
    dim3 grid((n + 31) / 32, (c + 31) / 32, b);
    bias_relu_pm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, relu, y, bias, (__half*)out_pm);
    return finish_launch("g4d bias_relu_pm");
}
