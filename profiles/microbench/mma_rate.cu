// Issue-rate probe for the warp-synchronous tensor-core path on sm_100a (mma.sync), used to decide whether the
// angle scan's grid contraction should move to tensor cores:  nvcc -arch=sm_100a -O3 mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int KIND>
__global__ void probe(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = 0x3f800000u, b1 = 0x3f000000u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            } else if (KIND == 1) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            } else {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(b0));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, int sms, double fma_per_mma) {
    float* out;
    const int threads = 256, blocks = sms * 4, iters = 4096;
    cudaMalloc(&out, sizeof(float) * threads * blocks);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<KIND><<<blocks, threads>>>(out, 16);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<KIND><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n_mma = (double)blocks * (threads / 32) * iters * 8;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double per_sm_clk = n_mma / (ms * 1e-3) / sms / (clk * 1e3);
    printf("%-22s %.3f ms  %.3f mma/clk/SM (at %d MHz nominal)  %.1f TFLOP/s\n", name, ms, per_sm_clk, clk / 1000,
           2 * fma_per_mma * n_mma / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0>("m16n8k8 tf32", sms, 16 * 8 * 8);
    run<1>("m16n8k16 bf16", sms, 16 * 8 * 16);
    run<2>("m16n8k4 tf32", sms, 16 * 8 * 4);
    return 0;
}
