// Hardware probe for the GRU convolution kernel's A-operand reuse (not product code):
//   (1) does cuTensorMapEncodeTiled accept a 3-D map whose dim-1 stride (16 rows) exceeds its dim-2 stride (1 row), and does the
//       box {64 ch, 8 segments, 20 steps} land in shared memory as row 8 j + i = global row 16 i + j (SWIZZLE_128B)?
//   (2) tcgen05.mma (SS, kind::f16) with the A descriptor start advanced by t * 1024 B over that tile -> tap t of a 1-D
//       convolution on 128 permuted pixels;
//   (3) natural row order, A descriptor start advanced by s * 128 B with base_offset 0 / s: which semantics shift by s rows?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/_probe_umma tools/probe_umma_shift.cu
#include "../mac-vo_b200/csrc/tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

namespace {
__device__ __forceinline__ void umma_f16_ss_1(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// mode 0: permuted 3-D A map, start += t * 1024      mode 1: natural 2-D A map (136 rows), start += s * 128, base_offset = bo
__global__ void __launch_bounds__(128)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out, uint8_t* dump,
             int mode, int shift, int bo) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                       // 160 rows x 128 B = 20 KB
    uint8_t* sb = smem + 20480;               // 128 rows x 128 B = 16 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 20480 + 16384);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    const uint32_t bar_full = smem_u32(bars), bar_done = bar_full + 8;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar_full, 1); mbar_init(bar_done, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc_1cta(smem_u32(tslot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tslot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar_full, (mode == 0 ? 20480 : 136 * 128) + 16384);
        if (mode == 0) tma_load_3d(smem_u32(sa), &map_a, bar_full, 0, 0, 0);
        else tma_load_2d(smem_u32(sa), &map_a, bar_full, 0, 0);
        tma_load_2d(smem_u32(sb), &map_b, bar_full, 0, 0);
        mbar_wait(bar_full, 0);
        tc_fence_after();
        uint64_t da = make_kmajor_sw128_desc(smem_u32(sa) + (mode == 0 ? shift * 1024 : shift * 128));
        if (mode == 1) da |= (uint64_t)(bo & 7) << 49;
        const uint64_t db = make_kmajor_sw128_desc(smem_u32(sb));
        const uint32_t idesc = make_idesc_f16(128, 128);
        for (int k = 0; k < 4; ++k) umma_f16_ss_1(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
        umma_commit(bar_done);
    }
    __syncthreads();
    mbar_wait(bar_done, 0);
    tc_fence_after();
    if (dump) for (int i = threadIdx.x; i < 20480; i += 128) dump[i] = sa[i];
    uint32_t r[32];
    for (int c = 0; c < 4; ++c) {
        tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[(size_t)threadIdx.x * 128 + c * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc_1cta(tmem, 128);
}
}  // namespace

int main() {
    const int ROWS = 512, C = 64;
    std::vector<__half> ha((size_t)ROWS * C), hb((size_t)128 * C);
    std::vector<float> fa(ha.size()), fb(hb.size());
    srand(1);
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (float)((rand() % 17) - 8) / 8.f; ha[i] = __float2half(fa[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)((rand() % 13) - 6) / 4.f; hb[i] = __float2half(fb[i]); }
    __half *da, *db; float* dout; uint8_t* ddump;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 128 * 4); cudaMalloc(&ddump, 20480);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap mb, ma3, ma2;
    if (!make_map_2d(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, db, C, 128, C * 2, 64, 128)) { printf("map_b failed\n"); return 1; }
    if (!make_map_2d(&ma2, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, da, C, ROWS, C * 2, 64, 136)) { printf("map_a2 failed\n"); return 1; }
    // permuted 3-D map: dims (C, segments, steps), strides (16 rows, 1 row)
    {
        PFN_encodeTiled enc = get_encode_fn();
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)(ROWS / 16 - 2), 20};
        cuuint64_t strides[2] = {(cuuint64_t)16 * C * 2, (cuuint64_t)C * 2};
        cuuint32_t box[3] = {64, 8, 20};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&ma3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, da, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("permuted 3-D map encode: %d\n", (int)r);
        if (r != CUDA_SUCCESS) ma3 = ma2;
    }
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    std::vector<float> out(128 * 128);
    std::vector<uint8_t> dump(20480);
    auto check = [&](auto rowmap, const char* name) {
        double worst = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                const int ar = rowmap(m);
                double ref = 0;
                for (int k = 0; k < C; ++k) ref += (double)fa[(size_t)ar * C + k] * fb[(size_t)n * C + k];
                worst = fmax(worst, fabs(ref - out[m * 128 + n]));
            }
        printf("%-44s max|err| = %.3g  %s\n", name, worst, worst < 1e-3 ? "OK" : "MISMATCH");
    };
    // (1) + (2)
    for (int t = 0; t < 5; ++t) {
        probe_kernel<<<1, 128, 40000>>>(ma3, mb, dout, ddump, 0, t, 0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode 0 t=%d: %s\n", t, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(dump.data(), ddump, dump.size(), cudaMemcpyDeviceToHost);
        if (t == 0) {
            // smem row r (de-swizzled) should hold global row 16 (r & 7) + (r >> 3)
            int bad = 0;
            for (int r = 0; r < 160; ++r)
                for (int ch = 0; ch < 8; ++ch) {
                    const __half* p = reinterpret_cast<const __half*>(dump.data() + r * 128 + ((ch ^ (r & 7)) << 4));
                    const int g = 16 * (r & 7) + (r >> 3);
                    for (int e2 = 0; e2 < 8; ++e2) bad += __half2float(p[e2]) != fa[(size_t)g * C + ch * 8 + e2];
                }
            printf("permuted box layout: %d mismatching elements\n", bad);
        }
        char nm[64]; snprintf(nm, 64, "permuted tile, tap %d (start += %d KB)", t, t);
        check([&](int m) { return 16 * (m & 7) + (m >> 3) + t; }, nm);
    }
    // (3)
    for (int s = 0; s < 5; ++s)
        for (int bo : {0, s}) {
            if (s == 0 && bo != 0) continue;
            probe_kernel<<<1, 128, 40000>>>(ma2, mb, dout, nullptr, 1, s, bo);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode 1 s=%d bo=%d: %s\n", s, bo, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
            char nm[64]; snprintf(nm, 64, "natural tile, start += %d rows, base_offset %d", s, bo);
            check([&](int m) { return m + s; }, nm);
        }
    return 0;
}
