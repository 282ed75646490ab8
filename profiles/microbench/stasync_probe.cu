#include <cstdint>
#include <cstdio>
__global__ void __cluster_dims__(2,1,1) k(float2* o) {
    __shared__ float2 s[64];
    __shared__ __align__(8) unsigned long long mbar;
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s);
    uint32_t rank; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(64 * 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const uint32_t peer = rank ^ 1;
    uint32_t rs, rm;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rs) : "r"(sa), "r"(peer));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rm) : "r"(mb), "r"(peer));
    float2 v = o[threadIdx.x];
    unsigned long long bits = ((unsigned long long)__float_as_uint(v.y) << 32) | __float_as_uint(v.x);
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" :: "r"(rs + threadIdx.x * 8), "l"(bits), "r"(rm) : "memory");
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb), "r"(0) : "memory");
    }
    o[64 + blockIdx.x * 64 + threadIdx.x] = s[63 - threadIdx.x];
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
int main() {
    float2* d; cudaMalloc(&d, sizeof(float2) * 64 * 3);
    float2 h[192]; for (int i = 0; i < 64; ++i) h[i] = make_float2(i, -i);
    cudaMemcpy(d, h, sizeof(float2) * 64, cudaMemcpyHostToDevice);
    k<<<2, 64>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(float2) * 192, cudaMemcpyDeviceToHost);
    int bad = 0; for (int b = 0; b < 2; ++b) for (int i = 0; i < 64; ++i) if (h[64 + b * 64 + i].x != 63 - i) ++bad;
    printf("%s bad=%d\n", cudaGetErrorString(e), bad);
    return 0;
}
