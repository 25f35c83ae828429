// tc_probe.cu -- measurements behind the tcgen05 step kernel:
//   1. TMEM -> register read rate (tcgen05.ld.32x32b) per SM, 1 and 2 CTAs per SM
//   2. numerics + operand layout of tcgen05.mma kind::tf32, A (chains x coefficients) in TMEM,
//      B (observations x coefficients) K-major in shared memory, 3xTF32 split, against FP64
//   3. issue rate of that MMA shape (M=128, N=208, K=8)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc(unsigned* slot, unsigned ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(unsigned addr, unsigned ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void ldtm32(unsigned addr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
}
__device__ __forceinline__ void ldtm16(unsigned addr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
}
__device__ __forceinline__ void sttm8(unsigned addr, const unsigned (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned mb, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mb, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(mb),
        "r"(parity) : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned mb) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mb) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc], kind::tf32
__device__ __forceinline__ void mma_tf32_ts(unsigned d, unsigned a, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a),
        "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__host__ __device__ inline unsigned long long smem_desc(unsigned addr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__host__ __device__ inline unsigned idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// ------------------------------------------------------------------ 1. TMEM read rate
__global__ void __launch_bounds__(128) ldtm_kernel(int iters, unsigned* sink, long long* clk) {
    __shared__ unsigned base_s;
    if (threadIdx.x < 32) tmem_alloc(&base_s, 256);
    fence_before();
    __syncthreads();
    fence_after();
    const unsigned base = base_s + ((threadIdx.x >> 5) << 21);   // lane field = 32 * warp
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {      // 6 x 32 columns = 192 columns of this warp's 32 lanes
            unsigned v[32];
            ldtm32(base + 32 * j, v);
            wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += v[k];
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_free(base_s, 256);
}

// same with two loads in flight before the wait and packed FFMA2 consumption
__global__ void __launch_bounds__(128) ldtm_ffma2_kernel(int iters, float* sink, long long* clk) {
    __shared__ unsigned base_s;
    if (threadIdx.x < 32) tmem_alloc(&base_s, 256);
    fence_before();
    __syncthreads();
    fence_after();
    const unsigned base = base_s + ((threadIdx.x >> 5) << 21);
    unsigned long long s[4] = {0ull, 0ull, 0ull, 0ull};
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        unsigned v[2][32];
        ldtm32(base, v[0]);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            if (j + 1 < 6) ldtm32(base + 32 * (j + 1), v[(j + 1) & 1]);
            if (j + 1 < 6) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");   // waits for both; fine for a rate test
            else wait_ld();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                unsigned long long r;
                asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(v[j & 1][2 * k]), "r"(v[j & 1][2 * k + 1]));
                asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s[k & 3]) : "l"(r));
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    float lo, hi, t = 0.f;
    for (int k = 0; k < 4; ++k) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s[k])); t += lo + hi; }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = t;
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_free(base_s, 256);
}

// ------------------------------------------------------------------ 2./3. MMA numerics and rate
// B slab layout (no swizzle, K-major): element (n, k) of a [N x 8] slab at byte
//   (n / 8) * 256 + (k / 4) * 128 + (n % 8) * 16 + (k % 4) * 4
// i.e. 8-row x 16-byte core matrices, the two K halves 128 bytes apart, 8-row groups 256 bytes apart.
constexpr int NOBS = 208;
constexpr int SLAB = NOBS * 32;     // bytes
__host__ __device__ inline int slab_index(int n, int k) { return ((n >> 3) * 256 + (k >> 2) * 128 + (n & 7) * 16 + (k & 3) * 4) / 4; }

// mode 0: LBO = 128 (K halves), SBO = 256 (row groups); mode 1: swapped
__global__ void __launch_bounds__(128) mma_kernel(const float* slabs, const float* coef /*[128][8]*/, float* out /*[128][NOBS]*/,
                                                  int mode, int reps, long long* clk) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned base_s;
    __shared__ unsigned long long mbar_s;
    const unsigned mb = smem_u32(&mbar_s);
    float* sm = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 3 * SLAB / 4; i += blockDim.x) sm[i] = slabs[i];
    if (threadIdx.x == 0) mbar_init(mb, 1);
    if (threadIdx.x < 32) tmem_alloc(&base_s, 256);
    // generic-proxy writes to shared memory must be visible to the tensor core (async proxy)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const unsigned tbase = base_s;
    const unsigned lane_base = tbase + ((threadIdx.x >> 5) << 21);
    const unsigned colA = 224;
    // A: this thread's chain -> hi / lo split of its 8 coefficients, and the ones row for ne
    unsigned hi[8], lo[8], one[8];
    for (int k = 0; k < 8; ++k) {
        const float b = coef[threadIdx.x * 8 + k];
        const unsigned h = __float_as_uint(b) & 0xFFFFE000u;
        hi[k] = h;
        lo[k] = __float_as_uint(b - __uint_as_float(h));
        one[k] = k < 3 ? __float_as_uint(1.0f) : 0u;
    }
    sttm8(lane_base + colA, hi);
    sttm8(lane_base + colA + 8, lo);
    sttm8(lane_base + colA + 16, one);
    wait_st();
    fence_before();
    __syncthreads();
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        fence_after();
        const unsigned lbo = mode == 0 ? 128 : 256, sbo = mode == 0 ? 256 : 128;
        const unsigned s0 = smem_u32(smem);
        const unsigned long long d_hi = smem_desc(s0, lbo, sbo), d_lo = smem_desc(s0 + SLAB, lbo, sbo), d_ne = smem_desc(s0 + 2 * SLAB, lbo, sbo);
        const unsigned id = idesc_tf32(128, NOBS);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            mma_tf32_ts(tbase, tbase + colA, d_hi, id, 0);          // A_hi . X_hi
            mma_tf32_ts(tbase, tbase + colA + 8, d_hi, id, 1);      // A_lo . X_hi
            mma_tf32_ts(tbase, tbase + colA, d_lo, id, 1);          // A_hi . X_lo
            mma_tf32_ts(tbase, tbase + colA + 16, d_ne, id, 1);     // 1 . (ne_hi, ne_mid, ne_lo)
        }
        mma_commit(mb);
    }
    mbar_wait(mb, 0);
    if (threadIdx.x == 0) { t1 = clock64(); clk[blockIdx.x] = t1 - t0; }
    fence_after();
    for (int j = 0; j < NOBS / 16; ++j) {
        unsigned v[16];
        ldtm16(lane_base + 16 * j, v);
        wait_ld();
        if (blockIdx.x == 0)
            for (int k = 0; k < 16; ++k) out[threadIdx.x * NOBS + 16 * j + k] = __uint_as_float(v[k]);
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_free(tbase, 256);
}

// The step kernel's MMA stream: per "sweep" two chunks (N = n0, then n1) of 4 MMAs each, a commit per
// chunk, optionally waiting for each commit before issuing the next chunk (as the kernel must, its
// accumulator being single-buffered).  128 TMEM columns per CTA so that four CTAs share an SM.
__global__ void __launch_bounds__(128) mma_stream_kernel(int n0, int n1, int sweeps, int wait_each, long long* clk) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned base_s;
    __shared__ unsigned long long mbar_s;
    const unsigned mb = smem_u32(&mbar_s);
    float* sm = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 3 * SLAB / 4; i += blockDim.x) sm[i] = 0.f;
    if (threadIdx.x == 0) mbar_init(mb, 1);
    if (threadIdx.x < 32) tmem_alloc(&base_s, 128);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const unsigned tbase = base_s;
    unsigned z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    sttm8(tbase + ((threadIdx.x >> 5) << 21) + 112, z);
    sttm8(tbase + ((threadIdx.x >> 5) << 21) + 120, z);
    wait_st();
    fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_after();
        const unsigned s0 = smem_u32(smem);
        unsigned parity = 0;
        const long long t0 = clock64();
        for (int sw = 0; sw < sweeps; ++sw) {
            for (int c = 0; c < 2; ++c) {
                const int n = c ? n1 : n0;
                if (n == 0) continue;
                const unsigned id = idesc_tf32(128, n);
                const unsigned base = s0 + (c ? n0 * 32 : 0);
                mma_tf32_ts(tbase, tbase + 112, smem_desc(base, 128, 256), id, 0);
                mma_tf32_ts(tbase, tbase + 120, smem_desc(base, 128, 256), id, 1);
                mma_tf32_ts(tbase, tbase + 112, smem_desc(base + SLAB, 128, 256), id, 1);
                mma_tf32_ts(tbase, tbase + 120, smem_desc(base + 2 * SLAB, 128, 256), id, 1);
                if (wait_each) { mma_commit(mb); mbar_wait(mb, parity); parity ^= 1; }
            }
        }
        if (!wait_each) { mma_commit(mb); mbar_wait(mb, 0); }   // one commit covers everything issued before it
        clk[blockIdx.x] = clock64() - t0;
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_free(tbase, 128);
}

static float tf32_trunc(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }

int main() {
    int dev = 0, sms = 0, khz = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    printf("SMs %d, clock %.0f MHz\n", sms, khz / 1e3);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    long long* clk; unsigned* sink;
    CK(cudaMalloc(&clk, sizeof(long long) * sms * 4));
    CK(cudaMalloc(&sink, 4 * 128 * sms * 4));
    std::vector<long long> h(sms * 4);

    for (int variant = 0; variant < 2; ++variant)
        for (int per_sm = 1; per_sm <= 2; ++per_sm) {
            const int iters = 2000, grid = sms * per_sm;
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                if (variant == 0) ldtm_kernel<<<grid, 128>>>(iters, sink, clk);
                else ldtm_ffma2_kernel<<<grid, 128>>>(iters, (float*)sink, clk);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
            }
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(h.data(), clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = std::max(mx, h[i]);
            const double bytes_cta = (double)iters * 6 * 32 * 128 * 4;
            printf("ldtm %s, %d CTA/SM: %.1f B/clk/SM by in-kernel clocks (%lld clk), %.1f B/clk/SM by wall (%.3f ms)\n",
                   variant ? "x32 double-buffered + FFMA2" : "x32 + wait + IADD", per_sm, bytes_cta * per_sm / mx, mx,
                   bytes_cta * grid / (ms * 1e-3 * khz * 1e3) / sms, ms);
        }

    // MMA numerics
    const int K = 8;
    std::vector<float> X(NOBS * K), ne(NOBS), coef(128 * K), slabs(3 * SLAB / 4, 0.f);
    srand(1);
    auto rnd = []() { return (float)((rand() / (double)RAND_MAX) * 2.0 - 1.0); };
    for (auto& v : X) v = rnd() * 1.7f;
    for (auto& v : ne) v = rnd() * 3.1f;
    for (auto& v : coef) v = rnd() * 0.9f;
    for (int n = 0; n < NOBS; ++n) {
        for (int k = 0; k < K; ++k) {
            const float hi = tf32_trunc(X[n * K + k]);
            slabs[slab_index(n, k)] = hi;
            slabs[SLAB / 4 + slab_index(n, k)] = X[n * K + k] - hi;
        }
        const float h1 = tf32_trunc(ne[n]);
        const float m1 = tf32_trunc(ne[n] - h1);
        slabs[2 * SLAB / 4 + slab_index(n, 0)] = h1;
        slabs[2 * SLAB / 4 + slab_index(n, 1)] = m1;
        slabs[2 * SLAB / 4 + slab_index(n, 2)] = (ne[n] - h1) - m1;
    }
    float *d_slabs, *d_coef, *d_out;
    CK(cudaMalloc(&d_slabs, slabs.size() * 4)); CK(cudaMalloc(&d_coef, coef.size() * 4)); CK(cudaMalloc(&d_out, 128 * NOBS * 4));
    CK(cudaMemcpy(d_slabs, slabs.data(), slabs.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_coef, coef.data(), coef.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * SLAB));
    std::vector<float> out(128 * NOBS);
    for (int mode = 0; mode < 2; ++mode) {
        CK(cudaMemset(d_out, 0, 128 * NOBS * 4));
        mma_kernel<<<1, 128, 3 * SLAB>>>(d_slabs, d_coef, d_out, mode, 1, clk);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0, worst_rel = 0, scale = 0;
        for (int c = 0; c < 128; ++c)
            for (int n = 0; n < NOBS; ++n) {
                double r = ne[n];
                for (int k = 0; k < K; ++k) r += (double)X[n * K + k] * (double)coef[c * K + k];
                const double err = fabs(out[c * NOBS + n] - r);
                worst = std::max(worst, err);
                scale = std::max(scale, fabs(r));
            }
        worst_rel = worst / scale;
        printf("mma mode %d (LBO=%d SBO=%d): max |err| %.3e, relative to max |r| %.3e  [out[0][0..3] = %g %g %g %g]\n", mode,
               mode == 0 ? 128 : 256, mode == 0 ? 256 : 128, worst, worst_rel, out[0], out[1], out[2], out[3]);
    }
    // MMA rate: all SMs, 1 and 2 CTAs per SM
    for (int per_sm = 1; per_sm <= 2; ++per_sm) {
        const int reps = 2000, grid = sms * per_sm;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            mma_kernel<<<grid, 128, 3 * SLAB>>>(d_slabs, d_coef, d_out, 0, reps, clk);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
        }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(h.data(), clk, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = std::max(mx, h[i]);
        printf("mma rate, %d CTA/SM: %.1f clk per M128 N%d K8 tf32 MMA per CTA (in-kernel), %.1f clk per MMA per SM by wall; "
               "%.1f TFLOP/s tf32 chip-wide\n", per_sm, (double)mx / (reps * 4), NOBS,
               ms * 1e-3 * khz * 1e3 / (reps * 4.0 * per_sm), 2.0 * 128 * NOBS * 8 * reps * 4.0 * grid / (ms * 1e-3) / 1e12);
    }
    // the step kernel's MMA stream, 1 and 4 CTAs per SM
    CK(cudaFuncSetAttribute(mma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * SLAB));
    {
        const int shapes[4][2] = {{208, 0}, {112, 96}, {112, 0}, {96, 0}};
        for (int per_sm = 1; per_sm <= 4; per_sm *= 4)
            for (int wait_each = 0; wait_each <= 1; ++wait_each)
                for (auto& sh : shapes) {
                    if (per_sm > 2 && sh[0] > 112) continue;   // 128 TMEM columns per CTA
                    const int sweeps = 1000, grid = sms * per_sm;
                    for (int rep = 0; rep < 2; ++rep) {
                        CK(cudaEventRecord(e0));
                        mma_stream_kernel<<<grid, 128, 3 * SLAB>>>(sh[0], sh[1], sweeps, wait_each, clk);
                        CK(cudaEventRecord(e1));
                        CK(cudaDeviceSynchronize());
                    }
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    printf("mma stream N=%d+%d, %d CTA/SM, %s: %.0f clk per sweep (8 or 4 MMAs) per SM by wall\n", sh[0], sh[1], per_sm,
                           wait_each ? "wait after each chunk" : "back to back", ms * 1e-3 * khz * 1e3 / (sweeps * (double)per_sm));
                }
    }
    // MMA round-trip latency: first issue -> mbarrier completion observed, one CTA, 4 MMAs (one chunk)
    for (int reps = 1; reps <= 4; reps *= 2) {
        mma_kernel<<<1, 128, 3 * SLAB>>>(d_slabs, d_coef, d_out, 0, reps, clk);
        CK(cudaDeviceSynchronize());
        mma_kernel<<<1, 128, 3 * SLAB>>>(d_slabs, d_coef, d_out, 0, reps, clk);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), clk, sizeof(long long), cudaMemcpyDeviceToHost));
        printf("mma latency: %d x 4 MMAs (N=%d) issue -> mbarrier observed: %lld clk\n", reps, NOBS, h[0]);
    }
    return 0;
}
