// Microbenchmark: issue/pipe cost of packed add.f32x2 (FADD2), predicated FMUL and ALU ops on one SM sub-partition.
// Inline PTX so the compiler cannot rewrite the arithmetic.  Timed with CUDA events over the whole grid
// (148 SMs x WPS warps per scheduler); result = SM cycles per loop body per scheduler, at the 1965 MHz boost clock.
#include <cstdio>
#include <cuda_runtime.h>

#define FADD2(a, r) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(r))
#define FADD(a, r) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(r))
#define FMUL(a, r) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(r))
#define LOP(a, b) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a) : "r"(b))

// MODE 0: 16 FADD2                 MODE 1: 32 FADD                MODE 2: 16 FADD2 + 16 LOP     MODE 3: 16 FADD2 + 32 LOP
// MODE 4: 32 FADD + 32 LOP         MODE 5: 16 FADD2 + 32 FMUL     MODE 6: 16 FADD2+32 FMUL+16 FADD2 (the paint mix)
// MODE 7: MODE 6 + 32 LOP          MODE 8: 32 FMUL + 32 LOP       MODE 9: 64 LOP
template <int MODE> __global__ void __launch_bounds__(128) mix(float *out, int iters)
{
    unsigned long long a[16], s[4];
    float f[32];
    unsigned q[32];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = ((unsigned long long)__float_as_uint(1.0f + i) << 32) | __float_as_uint(threadIdx.x * 1e-3f);
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) { f[i] = 1.0f + i * 1e-3f; q[i] = threadIdx.x + i; }
    unsigned long long r = ((unsigned long long)__float_as_uint(1e-9f) << 32) | __float_as_uint(1e-9f);
    float m = 1.0000001f;
    unsigned x = threadIdx.x * 2654435761u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 5 || MODE == 6 || MODE == 7) FADD2(a[i], r);
            if (MODE == 1 || MODE == 4) { FADD(f[2 * i], m); FADD(f[2 * i + 1], m); }
            if (MODE == 5 || MODE == 6 || MODE == 7 || MODE == 8) { FMUL(f[2 * i], m); FMUL(f[2 * i + 1], m); }
            if (MODE == 6 || MODE == 7) FADD2(s[i & 3], a[i]);
            if (MODE == 2) LOP(q[i], x);
            if (MODE == 3 || MODE == 4 || MODE == 7 || MODE == 8) { LOP(q[2 * i], x); LOP(q[2 * i + 1], x); }
            if (MODE == 9) { LOP(q[2 * i], x); LOP(q[2 * i + 1], x); LOP(q[(2 * i + 7) & 31], x); LOP(q[(2 * i + 12) & 31], x); }
        }
    }
    float acc = 0;
    unsigned u = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc += __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32));
#pragma unroll
    for (int i = 0; i < 4; i++) acc += __uint_as_float((unsigned)s[i]);
#pragma unroll
    for (int i = 0; i < 32; i++) { acc += f[i]; u ^= q[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (float)u;
}

template <int MODE> void run(float *out, const char *what, int ninstr)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int wps : {1, 2, 4, 8}) {
        mix<MODE><<<148 * wps, 128>>>(out, 100);
        cudaEventRecord(e0);
        mix<MODE><<<148 * wps, 128>>>(out, iters);   // 128 threads = 4 warps = 1 per scheduler per block
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double cyc = ms * 1e-3 * 1.965e9 / iters / wps;
        printf("%-44s %d warps/sched: %6.1f cycles per body (%d instr) -> IPC %.2f\n", what, wps, cyc, ninstr, ninstr / cyc);
    }
}

int main()
{
    float *out;
    cudaMalloc(&out, 148 * 8 * 128 * 4);
    run<0>(out, "16 FADD2", 16);
    run<1>(out, "32 FADD", 32);
    run<2>(out, "16 FADD2 + 16 LOP", 32);
    run<3>(out, "16 FADD2 + 32 LOP", 48);
    run<4>(out, "32 FADD + 32 LOP", 64);
    run<5>(out, "16 FADD2 + 32 FMUL", 48);
    run<6>(out, "16 FADD2 + 32 FMUL + 16 FADD2", 64);
    run<7>(out, "16 FADD2 + 32 FMUL + 16 FADD2 + 32 LOP", 96);
    run<8>(out, "32 FMUL + 32 LOP", 64);
    run<9>(out, "64 LOP", 64);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
