// Warp-level observation-covariance routine shared by match_cov_kernel (cov2to3.cu) and the fused observation
// kernel (observe.cu). See cov2to3.cu for the reference lines it follows.
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace macvo {

constexpr int COV_MAX_PER_LANE = 31;   // kernel_size <= 31 -> <= 961 taps -> <= 31 per lane

struct CovParams {
    float fx, fy, cx, cy;
    int ksize;
    float min_depth_cov;
};

// One warp: Gaussian-weighted depth statistics around (ul, vl) + closed-form 2D -> 3D projection.
// (suu, svv, suv) already clamped by the caller; depth_var_override >= 0 selects `wvar_depth = depth_cov`.
// Lane 0 receives the 6 unique entries s = [zz, xz, yz, xx, xy, yy] (NED order) in fp32; returns the warp-uniform
// out-of-image flag.
__device__ __forceinline__ bool match_cov_warp(float u, float v, long long ul, long long vl,
                                               const float* __restrict__ depth, int h, int w, float suu, float svv,
                                               float suv, bool override_var, float depth_var, const CovParams& P,
                                               int lane, float s6[6]) {
    // 2x2 inverse (the reference uses pinverse: identical for the non-singular matrices of this path)
    const float det = __fsub_rn(__fmul_rn(suu, svv), __fmul_rn(suv, suv));
    const float idet = __frcp_rn(det);
    const float i00 = __fmul_rn(svv, idet), i11 = __fmul_rn(suu, idet), i01 = -__fmul_rn(suv, idet);
    const float norm_c = __fmul_rn(2.f * CUDART_PI_F, sqrtf(det));
    const int ksize = P.ksize, half = ksize / 2, taps = ksize * ksize;

    float z[COV_MAX_PER_LANE], pv[COV_MAX_PER_LANE];
    float zsum = 0.f;
    bool oob = false;
#pragma unroll
    for (int t = 0; t < COV_MAX_PER_LANE; ++t) {
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
    for (int t = 0; t < COV_MAX_PER_LANE; ++t) {
        z[t] = __fdiv_rn(z[t], zsum);                               // normalised weights
        wavg = fmaf(z[t], pv[t], wavg);
    }
    wavg = warp_sum(wavg);
    float wvar = 0.f;
#pragma unroll
    for (int t = 0; t < COV_MAX_PER_LANE; ++t) {
        const float dd = pv[t] - wavg;
        wvar = fmaf(z[t], dd * dd, wvar);
    }
    wvar = warp_sum(wvar);
    // `wvar_depth = depth_cov` when no flow covariance is given but a per-keypoint depth variance is (Project2to3.py:163-171)
    if (override_var) wvar = depth_var;
    wvar = (wvar != wvar) ? wvar : fmaxf(wvar, P.min_depth_cov);        // clamp(min=...) keeps NaN
    oob = __any_sync(0xffffffffu, oob);

    const float fx = P.fx, fy = P.fy;
    const float du = __fsub_rn(u, P.cx), dv = __fsub_rn(v, P.cy);
    const float d2 = __fmul_rn(wavg, wavg);
    s6[3] = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(du, du), wvar), __fmul_rn(d2, suu)),
                                __fmul_rn(suu, wvar)), __fmul_rn(fx, fx));                         // xx
    s6[5] = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(dv, dv), wvar), __fmul_rn(d2, svv)),
                                __fmul_rn(svv, wvar)), __fmul_rn(fy, fy));                         // yy
    s6[0] = wvar;                                                                                  // zz
    s6[4] = __fdiv_rn(__fadd_rn(__fmul_rn(__fmul_rn(du, dv), wvar),
                                __fmul_rn(__fadd_rn(d2, wvar), suv)), __fmul_rn(fx, fy));          // xy
    s6[1] = __fdiv_rn(__fmul_rn(wvar, du), fx);                                                    // xz
    s6[2] = __fdiv_rn(__fmul_rn(wvar, dv), fy);                                                    // yz
    return oob;
}

// (K,3,3) float64 NED layout [[zz, xz, yz], [xz, xx, xy], [yz, xy, yy]]
__device__ __forceinline__ void store_cov9(double* o, const float s6[6]) {
    o[0] = s6[0]; o[1] = s6[1]; o[2] = s6[2];
    o[3] = s6[1]; o[4] = s6[3]; o[5] = s6[4];
    o[6] = s6[2]; o[7] = s6[4]; o[8] = s6[5];
}

}  // namespace macvo
