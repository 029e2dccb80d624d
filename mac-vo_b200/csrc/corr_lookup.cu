// (a5) 9x9 bilinear window lookup into the all-pairs correlation volume.
//
// Replaces MemoryDecoder.encode_flow_token (Module/Network/FlowFormer/core/decoder.py:141-153),
// bilinear_sampler (core/utils.py:26-34) and the `delta` buffer (decoder.py:124-129).
//
// Layout: one CTA serves 32 consecutive query pixels (a 128-byte output row per channel).
//   phase 1  each warp stages the <=12x12 footprint of 4 queries' own cost maps in shared memory
//            (each query reads ~0.5 KB out of its 19 KB map; rows are contiguous 48-byte segments);
//   phase 2  thread (lane = query, warp = channel group) evaluates the 81 taps with exactly ATen's
//            CPU grid_sampler arithmetic (align_corners=True, zeros padding) and writes out[b,c,q..q+31]
//            as full 128-byte lines.
// HBM-bound by the scattered footprint reads + the (B,81,H1,W1) write: ~7 MB per decoder iteration at
// 640x480 — a few microseconds; the kernel is launch/latency bound.
#include "common.cuh"

namespace {

constexpr int QPB = 32;        // queries per CTA
constexpr int FP = 12;         // staged footprint edge (9 window + 1 bilinear + 2 rounding slack)
constexpr int FPS = FP * FP + 1; // padded row: odd stride -> conflict-free lane-per-query reads
constexpr int THREADS = 256;

template <bool NHWC>
__global__ void __launch_bounds__(THREADS)
corr_lookup_kernel(const float* __restrict__ cost_maps, const float* __restrict__ coords, float* __restrict__ out,
                   int batch, int h1, int w1, int h2, int w2) {
    __shared__ float fp[QPB][FPS];
    __shared__ float s_cx[QPB], s_cy[QPB];
    __shared__ int s_ax[QPB], s_ay[QPB];

    const int n1 = h1 * w1;
    const long long total = (long long)batch * n1;
    const long long q0 = (long long)blockIdx.x * QPB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid < QPB) {
        long long q = q0 + tid;
        float cx = 0.f, cy = 0.f;
        if (q < total) {
            int b = (int)(q / n1), p = (int)(q % n1);
            cx = coords[((long long)b * 2 + 0) * n1 + p];
            cy = coords[((long long)b * 2 + 1) * n1 + p];
        }
        s_cx[tid] = cx;
        s_cy[tid] = cy;
        // anchor = floor(c) - 5 : window offsets -4..4, +1 for the bilinear neighbour, +-1 slack for the
        // normalise/unnormalise round trip. Clamp wild coordinates so the int conversion is defined.
        float fx = fminf(fmaxf(floorf(cx), -1.0e6f), 1.0e6f), fy = fminf(fmaxf(floorf(cy), -1.0e6f), 1.0e6f);
        s_ax[tid] = (int)fx - 5;
        s_ay[tid] = (int)fy - 5;
    }
    __syncthreads();

    // ---- phase 1: stage footprints (warp w handles queries w, w+8, ...) ---------------------------
    for (int ql = warp; ql < QPB; ql += THREADS / 32) {
        long long q = q0 + ql;
        if (q >= total) break;
        const float* map = cost_maps + q * (long long)h2 * w2;
        const int ax = s_ax[ql], ay = s_ay[ql];
        for (int e = lane; e < FP * FP; e += 32) {
            int ry = e / FP, rx = e - ry * FP;
            int y = ay + ry, x = ax + rx;
            float v = 0.f;
            if (x >= 0 && x < w2 && y >= 0 && y < h2) v = __ldg(map + (long long)y * w2 + x);
            fp[ql][e] = v;
        }
    }
    __syncthreads();

    // ---- phase 2: 81 taps per query --------------------------------------------------------------
    const long long q = q0 + lane;
    if (q >= total) return;
    const int b = (int)(q / n1), p = (int)(q % n1);
    const float cx = s_cx[lane], cy = s_cy[lane];
    const int ax = s_ax[lane], ay = s_ay[lane];
    const float wm1 = (float)(w2 - 1), hm1 = (float)(h2 - 1);
    const float sx = wm1 / 2.f, sy = hm1 / 2.f;             // ATen: scaling_factor = (size - 1) / 2
    const float* f = fp[lane];
    const float* map = cost_maps + q * (long long)h2 * w2;

    for (int c = warp; c < 81; c += THREADS / 32) {
        const int i = c / 9, j = c - i * 9;
        // reference: coords = centroid + delta; delta[i][j] = (i-4, j-4) added to (x, y)
        const float x = __fadd_rn(cx, (float)(i - 4));
        const float y = __fadd_rn(cy, (float)(j - 4));
        // bilinear_sampler: 2*x/(W-1) - 1 ; grid_sample unnormalise: (g + 1) * ((W-1)/2)
        const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, x), wm1), 1.f);
        const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, y), hm1), 1.f);
        const float ix = __fmul_rn(__fadd_rn(gx, 1.f), sx);
        const float iy = __fmul_rn(__fadd_rn(gy, 1.f), sy);
        const float x_w = floorf(ix), y_n = floorf(iy);
        const float w = __fsub_rn(ix, x_w), e = __fsub_rn(1.f, w);
        const float n = __fsub_rn(iy, y_n), s = __fsub_rn(1.f, n);
        const float nw = __fmul_rn(s, e), ne = __fmul_rn(s, w), sw = __fmul_rn(n, e), se = __fmul_rn(n, w);
        float acc = 0.f;
        if (x_w > -2.f && x_w < (float)w2 && y_n > -2.f && y_n < (float)h2) {   // else: all four taps outside
            const int x0 = (int)x_w, y0 = (int)y_n;
            const int rx = x0 - ax, ry = y0 - ay;
            float v00, v01, v10, v11;
            if (rx >= 0 && rx + 1 < FP && ry >= 0 && ry + 1 < FP) {
                v00 = f[ry * FP + rx];
                v01 = f[ry * FP + rx + 1];
                v10 = f[(ry + 1) * FP + rx];
                v11 = f[(ry + 1) * FP + rx + 1];
            } else {  // outside the staged footprint (cannot happen for finite coords; kept for safety)
                auto ld = [&](int yy, int xx) -> float {
                    return (xx >= 0 && xx < w2 && yy >= 0 && yy < h2) ? __ldg(map + (long long)yy * w2 + xx) : 0.f;
                };
                v00 = ld(y0, x0); v01 = ld(y0, x0 + 1); v10 = ld(y0 + 1, x0); v11 = ld(y0 + 1, x0 + 1);
            }
            acc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, nw), __fmul_rn(v01, ne)), __fmul_rn(v10, sw)),
                            __fmul_rn(v11, se));
        }
        if (NHWC) out[q * 81 + c] = acc;                       // pixels-major rows for the decoder's token MLP
        else out[((long long)b * 81 + c) * n1 + p] = acc;
    }
}

}  // namespace

extern "C" int macvo_corr_lookup(const float* cost_maps, const float* coords, float* out, int batch, int h1, int w1,
                                 int h2, int w2, void* stream) {
    if (!cost_maps || !coords || !out || batch <= 0 || h1 <= 0 || w1 <= 0 || h2 <= 1 || w2 <= 1) return MACVO_E_ARG;
    const long long total = (long long)batch * h1 * w1;
    const int grid = (int)((total + QPB - 1) / QPB);
    corr_lookup_kernel<false><<<grid, THREADS, 0, as_stream(stream)>>>(cost_maps, coords, out, batch, h1, w1, h2, w2);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}

// same lookup, output (batch*h1*w1, 81) pixels-major (the NHWC view of the reference's (batch, 81, h1, w1) map)
extern "C" int macvo_corr_lookup_rows(const float* cost_maps, const float* coords, float* out, int batch, int h1, int w1,
                                      int h2, int w2, void* stream) {
    if (!cost_maps || !coords || !out || batch <= 0 || h1 <= 0 || w1 <= 0 || h2 <= 1 || w2 <= 1) return MACVO_E_ARG;
    const long long total = (long long)batch * h1 * w1;
    const int grid = (int)((total + QPB - 1) / QPB);
    corr_lookup_kernel<true><<<grid, THREADS, 0, as_stream(stream)>>>(cost_maps, coords, out, batch, h1, w1, h2, w2);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}
