// (f4) Trajectory post-process at `terminate()`: MotionInterpolate.elaborate_map (Module/MapProcessor.py:52-79) with
// interpolate_pose (Utility/Math.py:96-122) and NormalizeQuat (:125-135), on the device, float64 like the reference.
//
//   motions_i = P_i^-1 P_{i+1};  motions flagged need_interp (except the first / last two) are replaced by the se3-linear
//   interpolation  Exp(t Log(M_g1 M_g0^-1)) M_g0  between the nearest unflagged motions g0 < i < g1;  the trajectory is
//   re-integrated as the inclusive left fold  C_0 = M_0, C_i = N(C_{i-1}) N(M_i)  (N = quaternion renormalisation; the
//   reference's `pp.cumops` with that lambda) and  P_{i+1} <- P_0 C_i  is written back in fp32.
//
// One CTA: the per-motion work is parallel; the fold is a blocked scan (each thread folds a contiguous chunk, thread 0 folds
// the chunk totals, each thread re-applies its prefix) — composition of rigid motions is associative, so this equals the
// sequential fold up to float64 rounding. Runs once per sequence on (F, 7) poses: latency, not bandwidth.
#include "common.cuh"

namespace {

constexpr int NT = 1024;

struct Se3 { double t[3], q[4]; };   // q = [x, y, z, w]

__device__ __forceinline__ void qrot(const double* q, const double* p, double* o) {
    const double ux = 2 * (q[1] * p[2] - q[2] * p[1]), uy = 2 * (q[2] * p[0] - q[0] * p[2]), uz = 2 * (q[0] * p[1] - q[1] * p[0]);
    o[0] = p[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    o[1] = p[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    o[2] = p[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
__device__ __forceinline__ Se3 mul(const Se3& a, const Se3& b) {
    Se3 o;
    qrot(a.q, b.t, o.t);
    o.t[0] += a.t[0]; o.t[1] += a.t[1]; o.t[2] += a.t[2];
    const double ax = a.q[0], ay = a.q[1], az = a.q[2], aw = a.q[3], bx = b.q[0], by = b.q[1], bz = b.q[2], bw = b.q[3];
    o.q[0] = aw * bx + ax * bw + ay * bz - az * by;
    o.q[1] = aw * by - ax * bz + ay * bw + az * bx;
    o.q[2] = aw * bz + ax * by - ay * bx + az * bw;
    o.q[3] = aw * bw - ax * bx - ay * by - az * bz;
    return o;
}
__device__ __forceinline__ Se3 inv(const Se3& a) {
    Se3 o;
    o.q[0] = -a.q[0]; o.q[1] = -a.q[1]; o.q[2] = -a.q[2]; o.q[3] = a.q[3];
    double r[3];
    qrot(o.q, a.t, r);
    o.t[0] = -r[0]; o.t[1] = -r[1]; o.t[2] = -r[2];
    return o;
}
__device__ __forceinline__ Se3 normq(Se3 a) {
    const double n = sqrt(a.q[0] * a.q[0] + a.q[1] * a.q[1] + a.q[2] * a.q[2] + a.q[3] * a.q[3]);
    a.q[0] /= n; a.q[1] /= n; a.q[2] /= n; a.q[3] /= n;
    return a;
}
__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
constexpr double EPS = 2.220446049250313e-16;

// SE3 Log -> [tau, phi]
__device__ void se3_log(const Se3& x, double* xi) {
    const double n = sqrt(x.q[0] * x.q[0] + x.q[1] * x.q[1] + x.q[2] * x.q[2]), w = x.q[3];
    const double f = n > EPS ? 2.0 * atan(n / w) / n : 2.0 / w - 2.0 * n * n / (3.0 * w * w * w);
    double phi[3] = {x.q[0] * f, x.q[1] * f, x.q[2] * f};
    const double t2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2], th = sqrt(t2);
    const double coef = th > EPS ? 1.0 / t2 - (1 + cos(th)) / (2 * th * sin(th)) : 1.0 / 12 + t2 / 720 + t2 * t2 / 30240;
    double a[3], b[3];
    cross3(phi, x.t, a);          // K t
    cross3(phi, a, b);            // K^2 t
    for (int i = 0; i < 3; ++i) { xi[i] = x.t[i] - 0.5 * a[i] + coef * b[i]; xi[3 + i] = phi[i]; }
}
// SE3 Exp of [tau, phi]
__device__ Se3 se3_exp(const double* xi) {
    const double* tau = xi; const double* phi = xi + 3;
    const double t2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2], th = sqrt(t2);
    double c1, c2, imag, real;
    if (th > EPS) { c1 = (1 - cos(th)) / t2; c2 = (th - sin(th)) / (t2 * th); imag = sin(0.5 * th) / th; real = cos(0.5 * th); }
    else { c1 = 0.5 - t2 / 24 + t2 * t2 / 720; c2 = 1.0 / 6 - t2 / 120 + t2 * t2 / 5040; imag = 0.5 - t2 / 48 + t2 * t2 / 3840; real = 1 - t2 / 8 + t2 * t2 / 384; }
    double a[3], b[3];
    cross3(phi, tau, a);
    cross3(phi, a, b);
    Se3 o;
    for (int i = 0; i < 3; ++i) { o.t[i] = tau[i] + c1 * a[i] + c2 * b[i]; o.q[i] = phi[i] * imag; }
    o.q[3] = real;
    return o;
}
__device__ __forceinline__ Se3 load(const double* p) { Se3 s; s.t[0] = p[0]; s.t[1] = p[1]; s.t[2] = p[2]; s.q[0] = p[3]; s.q[1] = p[4]; s.q[2] = p[5]; s.q[3] = p[6]; return s; }
__device__ __forceinline__ void store(double* p, const Se3& s) { p[0] = s.t[0]; p[1] = s.t[1]; p[2] = s.t[2]; p[3] = s.q[0]; p[4] = s.q[1]; p[5] = s.q[2]; p[6] = s.q[3]; }

// ws: motions (M,7) | chunk totals (NT,7) | bad flags (M bytes)
__global__ void __launch_bounds__(NT)
motion_interpolate_kernel(float* __restrict__ poses, const uint8_t* __restrict__ need_interp, int F, double* __restrict__ ws,
                          int* __restrict__ n_interp) {
    const int M = F - 1, tid = threadIdx.x;
    double* mot = ws;
    double* tot = ws + 7LL * M;
    uint8_t* bad = reinterpret_cast<uint8_t*>(tot + 7LL * NT);
    __shared__ int s_count;
    if (tid == 0) s_count = 0;
    // relative motions + flags
    for (int i = tid; i < M; i += NT) {
        Se3 a, b;
        for (int e = 0; e < 3; ++e) { a.t[e] = poses[7LL * i + e]; b.t[e] = poses[7LL * (i + 1) + e]; }
        for (int e = 0; e < 4; ++e) { a.q[e] = poses[7LL * i + 3 + e]; b.q[e] = poses[7LL * (i + 1) + 3 + e]; }
        store(mot + 7LL * i, mul(inv(a), b));
        bad[i] = (need_interp[i + 1] != 0 && i >= 2 && i < M - 2) ? 1 : 0;
    }
    __syncthreads();
    // interpolation of the flagged motions (only unflagged ones are read)
    for (int i = tid; i < M; i += NT) {
        if (!bad[i]) continue;
        int g0 = i - 1, g1 = i + 1;
        while (bad[g0]) --g0;                  // indices 0, 1 and M-2, M-1 are never flagged
        while (bad[g1]) ++g1;
        const double t = (double)(i - g0) / (double)(g1 - g0);
        const Se3 m0 = load(mot + 7LL * g0), m1 = load(mot + 7LL * g1);
        double xi[6];
        se3_log(mul(m1, inv(m0)), xi);
        for (int e = 0; e < 6; ++e) xi[e] *= t;
        store(mot + 7LL * i, mul(se3_exp(xi), m0));
        atomicAdd(&s_count, 1);
    }
    __syncthreads();
    // blocked inclusive fold  C_i = N(C_{i-1}) N(M_i)
    const int chunk = (M + NT - 1) / NT, lo = tid * chunk, hi = min(lo + chunk, M);
    if (lo < hi) {
        Se3 c = load(mot + 7LL * lo);
        if (lo > 0) c = normq(c);              // inside the fold every right operand is renormalised; M_0 itself is kept raw
        for (int i = lo + 1; i < hi; ++i) c = mul(normq(c), normq(load(mot + 7LL * i)));
        store(tot + 7LL * tid, c);
    }
    __syncthreads();
    if (tid == 0) {                            // exclusive prefix of the chunk totals
        const int nchunks = (M + chunk - 1) / chunk;
        Se3 run = load(tot);
        for (int c = 1; c < nchunks; ++c) {
            const Se3 mine = load(tot + 7LL * c);
            store(tot + 7LL * c, run);         // prefix = fold of all earlier chunks
            run = mul(normq(run), normq(mine));
        }
    }
    __syncthreads();
    // re-apply the prefix and write  P_{i+1} = P_0 C_i  (fp32, like `frames.data["pose"][1:] = (...).float()`)
    if (lo < hi) {
        Se3 p0;
        for (int e = 0; e < 3; ++e) p0.t[e] = poses[e];
        for (int e = 0; e < 4; ++e) p0.q[e] = poses[3 + e];
        Se3 c = load(mot + 7LL * lo);
        if (lo > 0) c = mul(normq(load(tot + 7LL * tid)), normq(c));
        for (int i = lo; i < hi; ++i) {
            if (i > lo) c = mul(normq(c), normq(load(mot + 7LL * i)));
            const Se3 o = mul(p0, c);
            float* dst = poses + 7LL * (i + 1);
            dst[0] = (float)o.t[0]; dst[1] = (float)o.t[1]; dst[2] = (float)o.t[2];
            dst[3] = (float)o.q[0]; dst[4] = (float)o.q[1]; dst[5] = (float)o.q[2]; dst[6] = (float)o.q[3];
        }
    }
    if (tid == 0 && n_interp) *n_interp = s_count;
}

}  // namespace

extern "C" size_t macvo_motion_interpolate_workspace_bytes(int num_frames) {
    if (num_frames < 2) return 0;
    const size_t m = (size_t)num_frames - 1;
    return (7 * m + 7 * (size_t)NT) * sizeof(double) + (m + 255) / 256 * 256;
}

extern "C" int macvo_motion_interpolate(float* poses, const uint8_t* need_interp, int num_frames, int* n_interp,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    if (num_frames < 0 || (num_frames > 0 && (!poses || !need_interp))) return MACVO_E_ARG;
    if (num_frames < 2) return MACVO_OK;                           // nothing to integrate
    if (!workspace || workspace_bytes < macvo_motion_interpolate_workspace_bytes(num_frames)) return MACVO_E_WORKSPACE;
    motion_interpolate_kernel<<<1, NT, 0, as_stream(stream)>>>(poses, need_interp, num_frames,
                                                                static_cast<double*>(workspace), n_interp);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}
