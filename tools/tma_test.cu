// scratch: minimal TMA 3D tile load experiment (not part of the product)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
struct Maps { CUtensorMap m[3]; };
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ Maps maps, const CUtensorMap *gmap, int bw, int bh, int c0, int c1, int c2, uint8_t *out) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint32_t s0 = smem_u32(sm), bar = s0 + 8192;
    int lane = threadIdx.x;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bw * bh) : "memory");
        const CUtensorMap *tm = MODE == 0 ? &maps.m[1] : gmap;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s0), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(0) : "memory");
    for (int i = lane; i < bw * bh; i += 32) out[i] = sm[i];
}
#include <cstdlib>
template <int MODE>
__global__ void k2(const __grid_constant__ Maps maps, const CUtensorMap *gmap, int bw, int bh, int c0, int c1, uint8_t *out) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint32_t s0 = smem_u32(sm), bar = s0 + 8192;
    int lane = threadIdx.x;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bw * bh) : "memory");
        const CUtensorMap *tm = MODE == 0 ? &maps.m[1] : gmap;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s0), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(0) : "memory");
    for (int i = lane; i < bw * bh; i += 32) out[i] = sm[i];
}
int main(int argc, char **argv) {
    const int pitch = 512, rows = 1000, frames = 2;
    const int rank = argc > 1 ? atoi(argv[1]) : 3, bw = argc > 2 ? atoi(argv[2]) : 16, bh = argc > 3 ? atoi(argv[3]) : 16, onlymode = argc > 4 ? atoi(argv[4]) : -1;
    printf("rank %d box %dx%d mode %d\n", rank, bw, bh, onlymode);
    size_t fb = (size_t)pitch * rows;
    std::vector<uint8_t> h(fb * frames);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i % pitch) * 7 + (i / pitch) * 13);
    uint8_t *d, *out; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice); cudaMalloc(&out, 8192);
    void *fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    auto enc = (CUresult(*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fnp;
    Maps maps; memset(&maps, 0, sizeof(maps));
    cuuint64_t dims[3] = {pitch, rows, frames}; cuuint64_t strides[2] = {pitch, fb}; cuuint32_t box[3] = {bw, bh, 1}; cuuint32_t es[3] = {1, 1, 1};
    for (int i = 0; i < 3; i++) { CUresult r = enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); printf("encode %d -> %d\n", i, (int)r); }
    CUtensorMap *gmap; cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &maps.m[1], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; mode++) {
        if (onlymode >= 0 && mode != onlymode) continue;
        cudaMemset(out, 0, 8192);
        int c0 = 13, c1 = 37, c2 = 1;
        if (rank == 2) { c2 = 0; if (mode == 0) k2<0><<<1, 32, 8192 + 64>>>(maps, gmap, bw, bh, c0, c1, out); else k2<1><<<1, 32, 8192 + 64>>>(maps, gmap, bw, bh, c0, c1, out); }
        else if (mode == 0) k<0><<<1, 32, 8192 + 64>>>(maps, gmap, bw, bh, c0, c1, c2, out); else k<1><<<1, 32, 8192 + 64>>>(maps, gmap, bw, bh, c0, c1, c2, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> o(bw * bh); cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < bh; y++) for (int x = 0; x < bw; x++) if (o[y * bw + x] != h[fb * c2 + (size_t)(c1 + y) * pitch + c0 + x]) bad++;
        printf("mode %d: mismatches %d\n", mode, bad);
    }
    return 0;
}
