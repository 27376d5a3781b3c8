// scratch: the CUDA programming guide's TMA sample (2-D int32 tile), to isolate an environment problem
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 1024, GH = 1024, SW = 32, SH = 8;
template <typename T>
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, T *out) {
    __shared__ alignas(128) T smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) out[i] = smem_buffer[i / SW][i % SW];
}
template <typename T> int run(CUtensorMapDataType dt, const char *name) {
    std::vector<T> h((size_t)GW * GH); for (size_t i = 0; i < h.size(); i++) h[i] = (T)(i * 7 + 3);
    T *d, *o; cudaMalloc(&d, h.size() * sizeof(T)); cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice); cudaMalloc(&o, SH * SW * sizeof(T));
    void *fnp = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    auto enc = (PFN_cuTensorMapEncodeTiled_v12000)fnp;
    CUtensorMap tm{}; cuuint64_t size[2] = {GW, GH}; cuuint64_t stride[1] = {GW * sizeof(T)}; cuuint32_t box[2] = {SW, SH}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, dt, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("%s encode %d; ", name, (int)r);
    const unsigned long long *w = (const unsigned long long *)&tm; for (int i = 0; i < 16; i++) printf("%016llx ", w[i]); printf("\n");
    kernel<T><<<1, 128>>>(tm, 64, 16, o);
    cudaError_t e = cudaDeviceSynchronize(); printf("%s: %s\n", name, cudaGetErrorString(e)); if (e) return 1;
    std::vector<T> res(SH * SW); cudaMemcpy(res.data(), o, res.size() * sizeof(T), cudaMemcpyDeviceToHost);
    int bad = 0; for (int yy = 0; yy < SH; yy++) for (int xx = 0; xx < SW; xx++) bad += res[yy * SW + xx] != h[(size_t)(16 + yy) * GW + 64 + xx];
    printf("%s mismatches %d\n", name, bad); return 0;
}
#include <cudaTypedefs.h>
int main(int argc, char **argv) { int w = argc > 1 ? atoi(argv[1]) : 0; return w == 0 ? run<int>(CU_TENSOR_MAP_DATA_TYPE_INT32, "int32") : run<unsigned char>(CU_TENSOR_MAP_DATA_TYPE_UINT8, "uint8"); }
