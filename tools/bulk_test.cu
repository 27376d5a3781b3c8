// scratch: 1-D bulk copy (UBLKCP) sanity check
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const uint8_t *src, uint8_t *out) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint32_t s0 = smem_u32(sm), bar = s0 + 4096;
    int lane = threadIdx.x;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(1024) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s0), "l"(src), "r"(1024), "r"(bar) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(0) : "memory");
    for (int i = lane; i < 1024; i += 32) out[i] = sm[i];
}
int main() {
    uint8_t *d, *o; cudaMalloc(&d, 4096); cudaMalloc(&o, 4096);
    uint8_t h[1024]; for (int i = 0; i < 1024; i++) h[i] = (uint8_t)(i * 3);
    cudaMemcpy(d, h, 1024, cudaMemcpyHostToDevice);
    k<<<1, 32, 4096 + 64>>>(d, o);
    cudaError_t e = cudaDeviceSynchronize(); printf("bulk: %s\n", cudaGetErrorString(e));
    if (e) return 1;
    uint8_t r[1024]; cudaMemcpy(r, o, 1024, cudaMemcpyDeviceToHost); int bad = 0; for (int i = 0; i < 1024; i++) bad += r[i] != h[i]; printf("bulk mismatches %d\n", bad);
    return 0;
}
