// Microbenchmark: latency of a warp-wide float sum, shuffle butterfly vs fixed-point REDUX (dependent chain).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float bfly(float s)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// sum of 32 non-negative floats, each < 2^(E+8) where 2^E <= Sprev < 2^(E+1)
__device__ __forceinline__ float redux_sum(float sl, float k1, float k2)
{
    const float M0 = 8388608.0f, M1 = 12582912.0f; // 2^23, 1.5*2^23
    const float a = fmaf(sl, k1, M0);
    const float ah = a - M0;
    const float rem = fmaf(sl, k1, -ah);
    const float b = fmaf(rem, 8388608.0f, M1);
    const int sa = __reduce_add_sync(0xffffffffu, __float_as_int(a)) - 32 * 0x4B000000;
    const int sb = __reduce_add_sync(0xffffffffu, __float_as_int(b)) - 32 * 0x4B400000;
    const float hi = (float)sa, lo = (float)sb;
    return fmaf(lo, k2 * (1.0f / 8388608.0f), hi * k2);
}

__global__ void k_bfly(float *out, long long *cyc, int iters)
{
    float s = 1.0f + threadIdx.x * 1e-3f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) s = bfly(s) * (1.0f / 32.0f);
    const long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_redux(float *out, long long *cyc, int iters)
{
    float s = 1.0f + threadIdx.x * 1e-3f;
    float sprev = 33.0f;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        // scale from the previous sum's exponent: k1 = 2^(23-E-8), k2 = 1/k1
        const int e = (__float_as_int(sprev) >> 23) & 0xff; // biased exponent of Sprev
        const float k1 = __int_as_float((127 + 23 - 8 + 127 - e) << 23);
        const float k2 = __int_as_float((e - 23 + 8) << 23);
        const float S = redux_sum(s, k1, k2);
        sprev = S;
        s = S * (1.0f / 32.0f);
    }
    const long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_check(float *out)
{
    // accuracy: random-ish partials
    unsigned x = 1234567u + threadIdx.x * 7919u;
    double worst = 0;
    for (int it = 0; it < 2000; it++) {
        x = x * 1664525u + 1013904223u;
        float sl = ((x >> 8) * (1.0f / 16777216.0f)) * ((it % 7) ? 1.0f : 1e-4f) * 37.5f;
        float ref = bfly(sl);
        double dref = 0;
        for (int l = 0; l < 32; l++) dref += (double)__shfl_sync(0xffffffffu, sl, l);
        float sprev = ref * ((it & 1) ? 0.02f : 900.0f); // S in [Sprev/1000, 100 Sprev]
        sprev = ref / ((it % 3 == 0) ? 100.0f : ((it % 3 == 1) ? 1.0f : 0.0011f));
        const int e = (__float_as_int(sprev) >> 23) & 0xff;
        const float k1 = __int_as_float((127 + 23 - 8 + 127 - e) << 23);
        const float k2 = __int_as_float((e - 23 + 8) << 23);
        float S = redux_sum(sl, k1, k2);
        double rel = fabs((double)S - dref) / dref;
        if (rel > worst) worst = rel;
        double relb = fabs((double)ref - dref) / dref;
        if (threadIdx.x == 0 && it < 4) printf("it %d ref %.9g redux %.9g exact %.12g rel %.3g (bfly rel %.3g)\n", it, ref, S, dref, rel, relb);
    }
    if (threadIdx.x == 0) printf("worst rel err of redux sum: %.3g\n", worst);
    out[threadIdx.x] = (float)worst;
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
    const int iters = 20000;
    long long h;
    for (int rep = 0; rep < 2; rep++) {
        k_bfly<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("butterfly: %.1f cycles per dependent reduction\n", (double)h / iters);
        k_redux<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("redux    : %.1f cycles per dependent reduction\n", (double)h / iters);
    }
    k_check<<<1, 32>>>(out);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
