// (f3) Observation building + outlier filter + MatchObs packing on the device.
//
// Replaces, for the two-frame pose graph, the per-frame host code between keypoint selection and the optimiser in
// Odometry/MACVO.py:198-270 (flow lookup, in-bound filter `filterPointsInRange`, nine `retrieve_pixels` gathers, two
// `ObsCovModel.estimate` calls, `pixel2point_NED`, `MatchObs.init`), the CovarianceSanityFilter
// (Module/OutlierFilter.py:91-100) and the point registration `pp.SE3.Act(prev_pose, pos0_Tc)` (:279-283), which the
// reference runs as ~14 device->host copies plus CPU tensor code per frame.
//
//   observe_kernel   one warp per selected keypoint: every gather, both 31x31 covariance estimates, the NaN / Inf
//                    sanity test and the world point, written to slot i of an uncompacted record table;
//   pack_kernel      one CTA: order-preserving compaction of the surviving records into (a) the float64
//                    structure-of-arrays the LM kernel reads (pgo.cu) and (b) the MatchObs columns the map wants,
//                    all inside ONE buffer so that a single asynchronous device->host copy ships a frame's
//                    observations; the survivor count stays on the device (pgo_lm_kernel reads it there).
//
// Arithmetic: fp32 in the reference's operation order (explicit round-to-nearest intrinsics, no contraction), widened
// to fp64 exactly where the reference calls `.double()`.
#include "cov2to3.cuh"

namespace {

struct ObserveArgs {
    const int64_t* kp0;            // (k,2) [u,v] selected keypoints on frame 0
    int k;
    const float* flow;             // (2,h,w) match flow frame0 -> frame1
    const float* match_cov;        // (3,h,w) [uu, vv, uv]
    const float* depth0;           // (h,w)
    const float* depth1;           // (h,w)
    const float* disparity1;       // (h,w)
    const float* disp_unc1;        // (h,w)
    int h, w, edge;
    macvo::CovParams P0, P1;       // intrinsics of frame 0 / frame 1 + covariance-model parameters
    float min_flow_var, match_cov_default;
    const double* prev_pose;       // (7) [t, q_xyzw] optimised pose of frame 0 (device, float64)
    double* next_pose;             // (7) receives the motion-model prediction for frame 1 (= prev pose, fp32-rounded)
};

// record slot (floats): 0 keep | 1,2 kp1 uv | 3 depth0 | 4 disp1 | 5 disp_unc1 | 6..8 uv cov (clamped) | 9..11 pos_Tw |
//                       12..17 cov0 (6 unique) | 18..23 cov1 (6 unique) | 24 inbound
constexpr int REC = 25;

__device__ __forceinline__ bool bad6(const float* s) {
    bool b = false;
#pragma unroll
    for (int i = 0; i < 6; ++i) b |= !isfinite(s[i]);
    return b;
}

__global__ void __launch_bounds__(128)
observe_kernel(ObserveArgs A, float* __restrict__ rec, int* __restrict__ status) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x < 7) {
        // StaticMotionModel.predict: est_pose = previous pose; the map stores poses in fp32 (Module/Map: FrameNode "pose")
        A.next_pose[threadIdx.x] = (double)(float)A.prev_pose[threadIdx.x];
    }
    if (i >= A.k) return;
    const long long u0 = A.kp0[2 * i], v0 = A.kp0[2 * i + 1];
    const int hw = A.h * A.w;
    float* r = rec + (long long)i * REC;
    const bool in0 = u0 >= 0 && u0 < A.w && v0 >= 0 && v0 < A.h;
    if (!in0) {                                  // cannot happen for selector output; keep the table defined
        if (lane == 0) { r[0] = 0.f; r[24] = 0.f; atomicExch(status, 2); }
        return;
    }
    const int p0 = (int)(v0 * A.w + u0);
    // kp1 = kp0 + flow[kp0]   (int64 + fp32 -> fp32)
    const float u1 = __fadd_rn((float)u0, A.flow[p0]);
    const float v1 = __fadd_rn((float)v0, A.flow[hw + p0]);
    // filterPointsInRange: strict inequalities against python ints
    const bool inb = (u1 < (float)(A.w - A.edge)) && (u1 > (float)A.edge) && (v1 < (float)(A.h - A.edge)) && (v1 > (float)A.edge);
    if (!inb) {
        if (lane == 0) { r[0] = 0.f; r[24] = 0.f; }
        return;
    }
    const long long ul1 = (long long)u1, vl1 = (long long)v1;         // .long(): truncation
    const int p1 = (int)(vl1 * A.w + ul1);                            // inside the image: edge > 0
    const float d0 = A.depth0[p0];
    const float disp1 = A.disparity1[p1], dunc1 = A.disp_unc1[p1];
    // frame-0 keypoints: constant quantisation covariance, clamped like any flow_cov (Project2to3.py:130-133)
    const float s0 = fmaxf(A.match_cov_default, A.min_flow_var);
    float c0[6], c1[6];
    bool oob = macvo::match_cov_warp((float)u0, (float)v0, u0, v0, A.depth0, A.h, A.w, s0, s0, 0.f, false, 0.f, A.P0, lane, c0);
    // frame-1 keypoints: the network's match covariance at the source pixel, clamped in place
    const float a = A.match_cov[p0], b = A.match_cov[hw + p0];
    const float suu = (a != a) ? a : fmaxf(a, A.min_flow_var);
    const float svv = (b != b) ? b : fmaxf(b, A.min_flow_var);
    const float suv = A.match_cov[2 * hw + p0];
    oob |= macvo::match_cov_warp(u1, v1, ul1, vl1, A.depth1, A.h, A.w, suu, svv, suv, false, 0.f, A.P1, lane, c1);
    if (lane != 0) return;
    if (oob) atomicExch(status, 1);
    // CovarianceSanityFilter: drop observations with NaN / Inf covariance on either frame
    const bool keep = !(bad6(c0) || bad6(c1));
    // pixel2point_NED (Utility/Point.py:15-17): [d, (u - cx) / fx * d, (v - cy) / fy * d]
    const float px = d0;
    const float py = __fmul_rn(__fdiv_rn(__fsub_rn((float)u0, A.P0.cx), A.P0.fx), d0);
    const float pz = __fmul_rn(__fdiv_rn(__fsub_rn((float)v0, A.P0.cy), A.P0.fy), d0);
    // SE3.Act in fp32: p + w * (2 q_v x p) + q_v x (2 q_v x p) + t
    const float tx = (float)A.prev_pose[0], ty = (float)A.prev_pose[1], tz = (float)A.prev_pose[2];
    const float qx = (float)A.prev_pose[3], qy = (float)A.prev_pose[4], qz = (float)A.prev_pose[5], qw = (float)A.prev_pose[6];
    const float ax = __fmul_rn(2.f, __fsub_rn(__fmul_rn(qy, pz), __fmul_rn(qz, py)));
    const float ay = __fmul_rn(2.f, __fsub_rn(__fmul_rn(qz, px), __fmul_rn(qx, pz)));
    const float az = __fmul_rn(2.f, __fsub_rn(__fmul_rn(qx, py), __fmul_rn(qy, px)));
    const float bx = __fsub_rn(__fmul_rn(qy, az), __fmul_rn(qz, ay));
    const float by = __fsub_rn(__fmul_rn(qz, ax), __fmul_rn(qx, az));
    const float bz = __fsub_rn(__fmul_rn(qx, ay), __fmul_rn(qy, ax));
    r[0] = keep ? 1.f : 0.f;
    r[1] = u1; r[2] = v1; r[3] = d0; r[4] = disp1; r[5] = dunc1;
    r[6] = suu; r[7] = svv; r[8] = suv;
    r[9] = __fadd_rn(__fadd_rn(__fadd_rn(px, __fmul_rn(qw, ax)), bx), tx);
    r[10] = __fadd_rn(__fadd_rn(__fadd_rn(py, __fmul_rn(qw, ay)), by), ty);
    r[11] = __fadd_rn(__fadd_rn(__fadd_rn(pz, __fmul_rn(qw, az)), bz), tz);
#pragma unroll
    for (int e = 0; e < 6; ++e) { r[12 + e] = c0[e]; r[18 + e] = c1[e]; }
    r[24] = 1.f;
}

// packed float64 buffer, sections sized by the CAPACITY cap (fixed pointers for the LM kernel):
//   [0,3c) pos_Tw | [3c,5c) kp2 uv | [5c,6c) kp2 disp | [6c,9c) uv cov | [9c,10c) disp cov          <- pgo.cu inputs
//   [10c,19c) obs1_covTc (c,3,3) | [19c,28c) obs2_covTc (c,3,3) | [28c,30c) pixel1_uv | [30c,31c) pixel1_d
//   [31c,31c+4) header: n_obs, n_inbound, k, status
__global__ void __launch_bounds__(1024)
pack_kernel(const float* __restrict__ rec, const int64_t* __restrict__ kp0, int k, int cap, double* __restrict__ out,
            int* __restrict__ n_obs, const int* __restrict__ status) {
    __shared__ int s_warp[32];
    __shared__ int s_base, s_inb;
    if (threadIdx.x == 0) { s_base = 0; s_inb = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long c = cap;
    for (int start = 0; start < k; start += 1024) {
        const int i = start + threadIdx.x;
        const float* r = rec + (long long)i * REC;
        const bool keep = i < k && r[0] != 0.f;
        const bool inb = i < k && r[24] != 0.f;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        const unsigned bal_in = __ballot_sync(0xffffffffu, inb);
        if (lane == 0) { s_warp[warp] = __popc(bal); if (bal_in) atomicAdd(&s_inb, __popc(bal_in)); }
        __syncthreads();
        int off = 0, tot = 0;
        for (int wv = 0; wv < 32; ++wv) { if (wv < warp) off += s_warp[wv]; tot += s_warp[wv]; }
        const int j = s_base + off + __popc(bal & ((1u << lane) - 1));
        if (keep && j < cap) {
            out[3 * j] = r[9]; out[3 * j + 1] = r[10]; out[3 * j + 2] = r[11];
            out[3 * c + 2 * j] = r[1]; out[3 * c + 2 * j + 1] = r[2];
            out[5 * c + j] = r[4];
            out[6 * c + 3 * j] = r[6]; out[6 * c + 3 * j + 1] = r[7]; out[6 * c + 3 * j + 2] = r[8];
            out[9 * c + j] = r[5];
            macvo::store_cov9(out + 10 * c + 9LL * j, r + 12);
            macvo::store_cov9(out + 19 * c + 9LL * j, r + 18);
            out[28 * c + 2 * j] = (double)kp0[2 * i]; out[28 * c + 2 * j + 1] = (double)kp0[2 * i + 1];
            out[30 * c + j] = r[3];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int n = min(s_base, cap);
        *n_obs = n;
        out[31 * c] = n; out[31 * c + 1] = s_inb; out[31 * c + 2] = k; out[31 * c + 3] = *status;
    }
}

// CovarianceSanityFilter.filter (Module/OutlierFilter.py:91-100) for device-resident covariances: one thread per observation
__global__ void cov_sanity_kernel(const double* __restrict__ c1, const double* __restrict__ c2, int k, uint8_t* __restrict__ good) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    bool ok = true;
#pragma unroll
    for (int e = 0; e < 9; ++e) ok &= isfinite(c1[9LL * i + e]) && isfinite(c2[9LL * i + e]);
    good[i] = ok ? 1 : 0;
}

}  // namespace

extern "C" int macvo_cov_sanity_filter(const double* obs1_cov, const double* obs2_cov, int k, uint8_t* good, void* stream) {
    if (k < 0) return MACVO_E_ARG;
    if (k == 0) return MACVO_OK;
    if (!obs1_cov || !obs2_cov || !good) return MACVO_E_ARG;
    cov_sanity_kernel<<<ceil_div(k, 256), 256, 0, as_stream(stream)>>>(obs1_cov, obs2_cov, k, good);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}

extern "C" size_t macvo_observe_workspace_bytes(int capacity) {
    return (size_t)capacity * REC * sizeof(float);
}

extern "C" size_t macvo_observe_packed_doubles(int capacity) { return (size_t)31 * capacity + 4; }

extern "C" int macvo_observe_pack(const int64_t* kp0_uv, int k, int capacity, const float* flow, const float* match_cov,
                                  const float* depth0, const float* depth1, const float* disparity1,
                                  const float* disp_unc1, int h, int w, int edge_width, const float* intr0,
                                  const float* intr1, int kernel_size, float min_flow_cov, float min_depth_cov,
                                  float match_cov_default, const double* prev_pose, double* next_pose, double* packed,
                                  int* n_obs, int* status, void* workspace, size_t workspace_bytes, void* stream) {
    if (k < 0 || capacity < 1 || k > capacity || h <= 0 || w <= 0 || edge_width <= 0 || kernel_size < 1 ||
        (kernel_size & 1) == 0 || kernel_size > 31)
        return MACVO_E_ARG;
    if (!flow || !match_cov || !depth0 || !depth1 || !disparity1 || !disp_unc1 || !intr0 || !intr1 || !prev_pose ||
        !next_pose || !packed || !n_obs || !status || !workspace || (k > 0 && !kp0_uv))
        return MACVO_E_ARG;
    if (workspace_bytes < macvo_observe_workspace_bytes(capacity)) return MACVO_E_WORKSPACE;
    ObserveArgs A;
    A.kp0 = kp0_uv; A.k = k; A.flow = flow; A.match_cov = match_cov; A.depth0 = depth0; A.depth1 = depth1;
    A.disparity1 = disparity1; A.disp_unc1 = disp_unc1; A.h = h; A.w = w; A.edge = edge_width;
    A.P0 = macvo::CovParams{intr0[0], intr0[1], intr0[2], intr0[3], kernel_size, min_depth_cov};   // HOST {fx, fy, cx, cy}
    A.P1 = macvo::CovParams{intr1[0], intr1[1], intr1[2], intr1[3], kernel_size, min_depth_cov};
    A.min_flow_var = min_flow_cov * min_flow_cov;
    A.match_cov_default = match_cov_default;
    A.prev_pose = prev_pose; A.next_pose = next_pose;
    cudaStream_t st = as_stream(stream);
    float* rec = static_cast<float*>(workspace);
    observe_kernel<<<max(1, ceil_div(k * 32, 128)), 128, 0, st>>>(A, rec, status);
    MACVO_LAUNCH_CHECK();
    pack_kernel<<<1, 1024, 0, st>>>(rec, kp0_uv, k, capacity, packed, n_obs, status);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}
