// Frontend "next" row: the elementwise glue of one MemoryDecoder refinement iteration
// (Module/Network/FlowFormer/core/gru.py:22-43 SepConvGRU, gma.py:84-130 Aggregate, covhead.py:95-131),
// fused so that the recurrent state lives in NHWC [h | x] buffers the cuDNN convolutions read in place:
//
//   gru_input   x = [inp | mf | mf + gamma * agg] written into the x-part (channels 128..511) of the 4 GRU input
//               buffers ([h|x] and [r*h|x] for the flow GRU and for the covariance GRU) — replaces 2 torch.cat
//               + mul + add per iteration and 8 torch.cat per GRU pass
//   gru_gates   z = sigmoid(zr[:, :128]);  [r*h | x] buffer <- sigmoid(zr[:, 128:]) * h
//   gru_blend   h <- (1 - z) * h + z * tanh(q), written to the [h|x] buffer and to a dense (P,128) copy
//
// All tensors fp32, pixels-major (P = B*H*W rows). Exact same arithmetic as the torch ops, one rounding each
// (sigmoid = 1 / (1 + exp(-x)) and tanhf at ATen's fp32 accuracy), parity in tests/test_gpu_nn_kernels.py.
#include "common.cuh"

namespace {

constexpr int HID = 128;          // hidden channels of SepConvGRU
constexpr int GRU_IN = 512;       // [h (128) | x (384)]

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256)
gru_input_kernel(const float* __restrict__ mf, const float* __restrict__ agg, const float* __restrict__ gamma,
                 float* __restrict__ b0, float* __restrict__ b1, float* __restrict__ b2, float* __restrict__ b3,
                 long long pixels) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // one float4 of one pixel's 128 channels
    if (e >= pixels * (HID / 4)) return;
    const long long p = e / (HID / 4);
    const int c = (int)(e % (HID / 4)) * 4;
    const float g = __ldg(gamma);
    const float4 m = *reinterpret_cast<const float4*>(mf + p * HID + c);
    const float4 a = *reinterpret_cast<const float4*>(agg + p * HID + c);
    const float4 s = make_float4(fmaf(g, a.x, m.x), fmaf(g, a.y, m.y), fmaf(g, a.z, m.z), fmaf(g, a.w, m.w));
    float* bufs[4] = {b0, b1, b2, b3};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!bufs[i]) continue;
        *reinterpret_cast<float4*>(bufs[i] + p * GRU_IN + 2 * HID + c) = m;
        *reinterpret_cast<float4*>(bufs[i] + p * GRU_IN + 3 * HID + c) = s;
    }
}

__global__ void __launch_bounds__(256)
gru_gates_kernel(const float* __restrict__ zr, const float* __restrict__ bias, const float* __restrict__ hx,
                 float* __restrict__ z_out, float* __restrict__ rhx, long long pixels) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= pixels * (HID / 4)) return;
    const long long p = e / (HID / 4);
    const int c = (int)(e % (HID / 4)) * 4;
    float4 zz = *reinterpret_cast<const float4*>(zr + p * 2 * HID + c);
    float4 rr = *reinterpret_cast<const float4*>(zr + p * 2 * HID + HID + c);
    if (bias) {                                                     // conv bias folded in (conv launched without one)
        const float4 bz = *reinterpret_cast<const float4*>(bias + c), br = *reinterpret_cast<const float4*>(bias + HID + c);
        zz = make_float4(zz.x + bz.x, zz.y + bz.y, zz.z + bz.z, zz.w + bz.w);
        rr = make_float4(rr.x + br.x, rr.y + br.y, rr.z + br.z, rr.w + br.w);
    }
    const float4 h = *reinterpret_cast<const float4*>(hx + p * GRU_IN + c);
    *reinterpret_cast<float4*>(z_out + p * HID + c) =
        make_float4(sigmoidf_(zz.x), sigmoidf_(zz.y), sigmoidf_(zz.z), sigmoidf_(zz.w));
    *reinterpret_cast<float4*>(rhx + p * GRU_IN + c) =
        make_float4(sigmoidf_(rr.x) * h.x, sigmoidf_(rr.y) * h.y, sigmoidf_(rr.z) * h.z, sigmoidf_(rr.w) * h.w);
}

__global__ void __launch_bounds__(256)
gru_blend_kernel(const float* __restrict__ q, const float* __restrict__ bias, const float* __restrict__ z,
                 float* __restrict__ hx, float* __restrict__ h_dense, long long pixels) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= pixels * (HID / 4)) return;
    const long long p = e / (HID / 4);
    const int c = (int)(e % (HID / 4)) * 4;
    float4 qq = *reinterpret_cast<const float4*>(q + p * HID + c);
    if (bias) {
        const float4 bq = *reinterpret_cast<const float4*>(bias + c);
        qq = make_float4(qq.x + bq.x, qq.y + bq.y, qq.z + bq.z, qq.w + bq.w);
    }
    const float4 zz = *reinterpret_cast<const float4*>(z + p * HID + c);
    const float4 h = *reinterpret_cast<const float4*>(hx + p * GRU_IN + c);
    float4 o;                                    // (1 - z) * h + z * tanh(q), same association as gru.py:33,41
    o.x = (1.f - zz.x) * h.x + zz.x * tanhf(qq.x);
    o.y = (1.f - zz.y) * h.y + zz.y * tanhf(qq.y);
    o.z = (1.f - zz.z) * h.z + zz.z * tanhf(qq.z);
    o.w = (1.f - zz.w) * h.w + zz.w * tanhf(qq.w);
    *reinterpret_cast<float4*>(hx + p * GRU_IN + c) = o;
    if (h_dense) *reinterpret_cast<float4*>(h_dense + p * HID + c) = o;
}

// q_in = LayerNorm(query) + sine_embed(coords1): the input of the cross-attention query projection
// (decoder.py:56-66, attention.py:71-101 LinearPositionEmbeddingSine with dim 64: [sin x f | cos x f | sin y f | cos y f],
// f = k * pi / 200, k = 0..15 — passed in as the fp32 table torch builds). One warp per pixel, two channels per lane.
__global__ void __launch_bounds__(256)
query_prep_kernel(const float* __restrict__ query, const float* __restrict__ w, const float* __restrict__ b,
                  const float* __restrict__ coords, const float* __restrict__ freq, float* __restrict__ out,
                  long long pixels, int n1, float eps) {
    const long long p = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= pixels) return;
    const float2 t = *reinterpret_cast<const float2*>(query + p * 64 + lane * 2);
    const float mean = warp_sum(t.x + t.y) * (1.f / 64);
    const float d0 = t.x - mean, d1 = t.y - mean;
    const float rstd = rsqrtf(warp_sum(fmaf(d0, d0, d1 * d1)) * (1.f / 64) + eps);
    const float2 ww = *reinterpret_cast<const float2*>(w + lane * 2), bb = *reinterpret_cast<const float2*>(b + lane * 2);
    const long long bi = p / n1, pi = p % n1;
    const int grp = lane >> 3;                                   // channels 2*lane, 2*lane+1 -> group of 16
    const float c = coords[(bi * 2 + (grp >> 1)) * n1 + pi];     // groups 0,1 use x; 2,3 use y
    const int k = (2 * lane) & 15;
    const float a0 = __fmul_rn(c, freq[k]), a1 = __fmul_rn(c, freq[k + 1]);
    const float e0 = (grp & 1) ? cosf(a0) : sinf(a0), e1 = (grp & 1) ? cosf(a1) : sinf(a1);
    *reinterpret_cast<float2*>(out + p * 64 + lane * 2) =
        make_float2(__fadd_rn(d0 * rstd * ww.x + bb.x, e0), __fadd_rn(d1 * rstd * ww.y + bb.y, e1));
}

inline unsigned grid_for(long long pixels) { return (unsigned)((pixels * (HID / 4) + 255) / 256); }

}  // namespace

extern "C" int macvo_gru_input(const float* mf, const float* agg, const float* gamma, float* buf0, float* buf1,
                               float* buf2, float* buf3, long long pixels, void* stream) {
    if (!mf || !agg || !gamma || !buf0 || pixels < 0) return MACVO_E_ARG;
    if (pixels == 0) return MACVO_OK;
    gru_input_kernel<<<grid_for(pixels), 256, 0, as_stream(stream)>>>(mf, agg, gamma, buf0, buf1, buf2, buf3, pixels);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}

extern "C" int macvo_gru_gates(const float* zr, const float* bias, const float* hx, float* z_out, float* rhx,
                               long long pixels, void* stream) {
    if (!zr || !hx || !z_out || !rhx || pixels < 0) return MACVO_E_ARG;
    if (pixels == 0) return MACVO_OK;
    gru_gates_kernel<<<grid_for(pixels), 256, 0, as_stream(stream)>>>(zr, bias, hx, z_out, rhx, pixels);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}

extern "C" int macvo_gru_blend(const float* q, const float* bias, const float* z, float* hx, float* h_dense,
                               long long pixels, void* stream) {
    if (!q || !z || !hx || pixels < 0) return MACVO_E_ARG;
    if (pixels == 0) return MACVO_OK;
    gru_blend_kernel<<<grid_for(pixels), 256, 0, as_stream(stream)>>>(q, bias, z, hx, h_dense, pixels);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}

extern "C" int macvo_query_prep(const float* query, const float* ln_weight, const float* ln_bias, const float* coords,
                                const float* freq, float* out, int batch, int n1, float eps, void* stream) {
    if (!query || !ln_weight || !ln_bias || !coords || !freq || !out || batch <= 0 || n1 <= 0) return MACVO_E_ARG;
    const long long pixels = (long long)batch * n1;
    query_prep_kernel<<<(unsigned)((pixels + 7) / 8), 256, 0, as_stream(stream)>>>(query, ln_weight, ln_bias, coords, freq,
                                                                                   out, pixels, n1, eps);
    MACVO_LAUNCH_CHECK();
    return MACVO_OK;
}
