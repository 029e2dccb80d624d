// (a10)+(a11) per-keypoint observation covariance: 31x31 Gaussian-weighted depth statistics and the
// closed-form 2D -> 3D covariance projection. One warp per keypoint.
//
// Replaces MatchCovariance.estimate (Module/Covariance/Project2to3.py:124-182), gaussain_full_kernels
// (Utility/Math.py:44-63), Covariance_2to3_full (Project2to3.py:377-424), create_3x3_matrix (:426-434,
// which assembles the result on the CPU through 9 implicit device->host copies) and pixel2point_NED
// (Utility/Point.py:15-17). Reference quirks kept: in-place clamp of the caller's
// flow_cov; the kernel axis weighted by sigma_uu runs along image ROWS of the depth patch; the
// depth_cov argument is ignored whenever flow_cov is given.
//
// L2-gather bound: K * (961*4 + 20) B in, K * 72 B out (2 MB at K=512, 16 MB at K=4096).
// fp32 arithmetic like the reference, result widened to fp64 at the end (`.double()`).
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int MAX_PER_LANE = 31;   // kernel_size <= 31 -> <= 961 taps -> <= 31 per lane

template <typename KP>
__global__ void __launch_bounds__(128)
match_cov_kernel(const KP* __restrict__ kp, int k, const float* __restrict__ depth, int h, int w,
                 float* __restrict__ flow_cov, float fx, float fy, float cx, float cy, int ksize, float min_flow_var,
                 float min_depth_cov, float match_cov_default, double* __restrict__ out_cov,
                 float* __restrict__ out_point, int* __restrict__ status) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= k) return;
    const KP ku = kp[2 * warp], kv = kp[2 * warp + 1];
    const float u = (float)ku, v = (float)kv;                   // value used in the closed form
    const long long ul = (long long)ku, vl = (long long)kv;     // .long(): truncation, used for indexing

    float suu, svv, suv;
    if (flow_cov) {
        const float a = flow_cov[3 * warp], b = flow_cov[3 * warp + 1];
        suu = (a != a) ? a : fmaxf(a, min_flow_var);             // clamp_(min=min_flow_cov**2); NaN stays NaN
        svv = (b != b) ? b : fmaxf(b, min_flow_var);
        suv = flow_cov[3 * warp + 2];
        __syncwarp();
        if (lane == 0) { flow_cov[3 * warp] = suu; flow_cov[3 * warp + 1] = svv; }
    } else {
        suu = svv = match_cov_default;
        suv = 0.f;
    }
    // 2x2 inverse (the reference uses pinverse: identical for the non-singular matrices of this path)
    const float det = __fsub_rn(__fmul_rn(suu, svv), __fmul_rn(suv, suv));
    const float idet = __frcp_rn(det);
    const float i00 = __fmul_rn(svv, idet), i11 = __fmul_rn(suu, idet), i01 = -__fmul_rn(suv, idet);
    const float norm_c = __fmul_rn(2.f * CUDART_PI_F, sqrtf(det));
    const int half = ksize / 2, taps = ksize * ksize;

    float z[MAX_PER_LANE], pv[MAX_PER_LANE];
    float zsum = 0.f;
    bool oob = false;
#pragma unroll
    for (int t = 0; t < MAX_PER_LANE; ++t) {
        const int e = lane + 32 * t;
        z[t] = 0.f; pv[t] = 0.f;
        if (e < taps) {
            const int a = e / ksize, b = e - a * ksize;            // a: kernel x-axis (sigma_uu) <-> image ROW offset
            const float xa = (float)(a - half), yb = (float)(b - half);
            // exp(-0.5 * [xa, yb] inv [xa, yb]^T)
            const float quad = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(xa, xa), i00),
                                                   __fmul_rn(__fmul_rn(2.f * xa, yb), i01)),
                                         __fmul_rn(__fmul_rn(yb, yb), i11));
            z[t] = __fdiv_rn(expf(-0.5f * quad), norm_c);
            long long yy = vl + (a - half), xx = ul + (b - half);
            if (yy < 0) yy += h;                                     // python-style negative index wrap
            if (xx < 0) xx += w;
            if (yy < 0 || yy >= h || xx < 0 || xx >= w) oob = true;
            else pv[t] = __ldg(depth + yy * w + xx);
            zsum += z[t];
        }
    }
    zsum = warp_sum(zsum);
    float wavg = 0.f;
#pragma unroll
    for (int t = 0; t < MAX_PER_LANE; ++t) {
        z[t] = __fdiv_rn(z[t], zsum);                               // normalised weights
        wavg = fmaf(z[t], pv[t], wavg);
    }
    wavg = warp_sum(wavg);
    float wvar = 0.f;
#pragma unroll
    for (int t = 0; t < MAX_PER_LANE; ++t) {
        const float dd = pv[t] - wavg;
        wvar = fmaf(z[t], dd * dd, wvar);
    }
    wvar = warp_sum(wvar);
    wvar = (wvar != wvar) ? wvar : fmaxf(wvar, min_depth_cov);        // clamp(min=...) keeps NaN
    if (__any_sync(0xffffffffu, oob) && lane == 0) atomicExch(status, 1);

    if (lane == 0) {
        const float du = __fsub_rn(u, cx), dv = __fsub_rn(v, cy);
        const float d2 = __fmul_rn(wavg, wavg);
        const float s_xx = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(du, du), wvar), __fmul_rn(d2, suu)),
                                              __fmul_rn(suu, wvar)), __fmul_rn(fx, fx));
        const float s_yy = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(dv, dv), wvar), __fmul_rn(d2, svv)),
                                              __fmul_rn(svv, wvar)), __fmul_rn(fy, fy));
        const float s_zz = wvar;
        const float s_xy = __fdiv_rn(__fadd_rn(__fmul_rn(__fmul_rn(du, dv), wvar),
                                              __fmul_rn(__fadd_rn(d2, wvar), suv)), __fmul_rn(fx, fy));
        const float s_xz = __fdiv_rn(__fmul_rn(wvar, du), fx);
        const float s_yz = __fdiv_rn(__fmul_rn(wvar, dv), fy);
        double* o = out_cov + 9LL * warp;
        o[0] = s_zz; o[1] = s_xz; o[2] = s_yz;
        o[3] = s_xz; o[4] = s_xx; o[5] = s_xy;
        o[6] = s_yz; o[7] = s_xy; o[8] = s_yy;
        if (out_point) {   // pixel2point_NED with the CENTRE pixel's depth (Odometry/MACVO.py:209,239)
            const long long yy = vl < 0 ? vl + h : vl, xx = ul < 0 ? ul + w : ul;
            const float d = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? depth[yy * w + xx] : CUDART_NAN_F;
            out_point[3 * warp] = d;
            out_point[3 * warp + 1] = __fmul_rn(__fdiv_rn(du, fx), d);
            out_point[3 * warp + 2] = __fmul_rn(__fdiv_rn(dv, fy), d);
        }
    }
}

}  // namespace

extern "C" int macvo_match_covariance(const void* kp, int kp_is_int64, int k, const float* depth, int h, int w,
                                      float* flow_cov, float fx, float fy, float cx, float cy, int kernel_size,
                                      float min_flow_cov, float min_depth_cov, float match_cov_default,
                                      double* out_cov, float* out_point, int* status, void* stream) {
    if (k < 0 || h <= 0 || w <= 0 || kernel_size < 1 || (kernel_size & 1) == 0 || kernel_size > 31) return MACVO_E_ARG;
    if (k == 0) return MACVO_OK;
    if (!kp || !depth || !out_cov || !status) return MACVO_E_ARG;
    const float min_flow_var = min_flow_cov * min_flow_cov;
    const int blocks = ceil_div(k * 32, 128);
    if (kp_is_int64)
        match_cov_kernel<int64_t><<<blocks, 128, 0, as_stream(stream)>>>(
            static_cast<const int64_t*>(kp), k, depth, h, w, flow_cov, fx, fy, cx, cy, kernel_size, min_flow_var,
            min_depth_cov, match_cov_default, out_cov, out_point, status);
    else
        match_cov_kernel<float><<<blocks, 128, 0, as_stream(stream)>>>(
            static_cast<const float*>(kp), k, depth, h, w, flow_cov, fx, fy, cx, cy, kernel_size, min_flow_var,
            min_depth_cov, match_cov_default, out_cov, out_point, status);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}
