// scratch: guide-style TMA sample with runtime geometry
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int bytes, unsigned char *out) {
    extern __shared__ __align__(128) unsigned char smem_buffer[];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem_buffer[i];
}
int main(int argc, char **argv) {
    int GW = atoi(argv[1]), GH = atoi(argv[2]), SW = atoi(argv[3]), SH = atoi(argv[4]), x = atoi(argv[5]), y = atoi(argv[6]), nt = atoi(argv[7]);
    std::vector<unsigned char> h((size_t)GW * GH); for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)(i * 7 + 3 + i / GW);
    unsigned char *d, *o; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice); cudaMalloc(&o, SH * SW);
    void *fnp = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fnp;
    CUtensorMap tm{}; cuuint64_t size[2] = {(cuuint64_t)GW, (cuuint64_t)GH}; cuuint64_t stride[1] = {(cuuint64_t)GW}; cuuint32_t box[2] = {(cuuint32_t)SW, (cuuint32_t)SH}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("G %dx%d box %dx%d at (%d,%d) nt %d: encode %d; ", GW, GH, SW, SH, x, y, nt, (int)r);
    kernel<<<1, nt, SW * SH>>>(tm, x, y, SW * SH, o);
    cudaError_t e = cudaDeviceSynchronize(); printf("%s; ", cudaGetErrorString(e)); if (e) { printf("\n"); return 1; }
    std::vector<unsigned char> res(SH * SW); cudaMemcpy(res.data(), o, res.size(), cudaMemcpyDeviceToHost);
    int bad = 0; for (int yy = 0; yy < SH; yy++) for (int xx = 0; xx < SW; xx++) bad += res[yy * SW + xx] != h[(size_t)(y + yy) * GW + x + xx];
    printf("mismatches %d\n", bad); return 0;
}
