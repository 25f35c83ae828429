// fma_probe.cu -- which FFMA operand patterns reach the FP32 pipe peak on sm_100a?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_probe fma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__global__ void __launch_bounds__(256) k(float* out, int iters, float seed) {
    float a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = seed + i; b[i] = 1.0f + 1e-7f * (threadIdx.x + i); }
    float x = seed * 0.5f, y = seed * 0.25f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (V == 0) a[i] = fmaf(x, b[i], a[i]);          // shared multiplicand, per-acc multiplier (hot loop shape)
                if (V == 1) a[i] = fmaf(a[i], x, y);             // one live register per FFMA
                if (V == 2) a[i] = fmaf(a[i], 1.0001f, a[i]);    // immediate form
                if (V == 3) a[i] = fmaf(a[i], x, a[i]);          // two reads of the same register
                if (V == 4) a[i] = fmaf(b[i], b[(i + 1) & 15], a[i]);  // three distinct registers
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] + b[i];
    if (s == 12345.678f) out[0] = s;
}

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}

// packed FP32x2 FMA (SASS FFMA2): 16 pair accumulators
template <int V>
__global__ void __launch_bounds__(256) k2(float* out, int iters, float seed) {
    unsigned long long a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = pack2(seed + i, seed - i); b[i] = pack2(1.0f + 1e-7f * (threadIdx.x + i), 1.0f - 1e-7f * i); }
    unsigned long long x = pack2(seed * 0.5f, seed * 0.5f), y = pack2(seed * 0.25f, seed);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (V == 0) a[i] = ffma2(x, b[i], a[i]);
                if (V == 1) a[i] = ffma2(a[i], x, y);
                if (V == 2) a[i] = ffma2(b[i], b[(i + 1) & 15], a[i]);
            }
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i] ^ b[i];
    if (s == 12345678ull) out[0] = 1.f;
}

template <int V>
void run2(const char* name) {
    float* d; cudaMalloc(&d, 16);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096;
    for (int occ : {1, 2, 4}) {
        const int blocks = sms * occ, threads = 256;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k2<V><<<blocks, threads>>>(d, iters, 1.0f);
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0); k2<V><<<blocks, threads>>>(d, iters, 1.0f); cudaEventRecord(e1);
            cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double fl = 2.0 * 2 * 16 * 4 * (double)iters * threads * blocks;
        printf("%-28s ctas/SM=%d  %.2f TFLOP/s  (%.1f lanes/clk/SM at 1965 MHz)\n", name, occ, fl / best / 1e9,
               fl / 2 / (best * 1e-3) / sms / 1.965e9);
    }
    cudaFree(d);
}

template <int V>
void run(const char* name) {
    float* d; cudaMalloc(&d, 16);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096;
    for (int occ : {1, 2, 4, 8}) {
        const int blocks = sms * occ, threads = 256;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<V><<<blocks, threads>>>(d, iters, 1.0f);
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0); k<V><<<blocks, threads>>>(d, iters, 1.0f); cudaEventRecord(e1);
            cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double fl = 2.0 * 16 * 8 * (double)iters * threads * blocks;
        printf("%-28s ctas/SM=%d  %.2f TFLOP/s  (%.1f lanes/clk/SM at 1965 MHz)\n", name, occ, fl / best / 1e9,
               fl / 2 / (best * 1e-3) / sms / 1.965e9);
    }
    cudaFree(d);
}

int main() {
    run<0>("x*b[i]+a[i]");
    run<1>("a[i]*x+y");
    run<2>("a[i]*imm+a[i]");
    run<3>("a[i]*x+a[i]");
    run<4>("b[i]*b[i+1]+a[i]");
    run2<0>("f32x2: x*b[i]+a[i]");
    run2<1>("f32x2: a[i]*x+y");
    run2<2>("f32x2: b[i]*b[i+1]+a[i]");
    return 0;
}
