// pipes.cu -- issue / pipe throughput of the instructions the traversal loop is made of, on the GPU at hand.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
// Every kernel runs `iters` iterations of 8 independent dependency chains per thread with 1024 threads per SM
// (8 warps per scheduler) on every SM; reported: warp-instructions per clock per SM (4 = one per scheduler per clock).
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 16

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float b, float c, unsigned sel)
{
    float a[CHAINS];
    unsigned u[CHAINS];
    unsigned long long p[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) { a[j] = threadIdx.x * 1e-3f + j; u[j] = threadIdx.x * 2654435761u + j; p[j] = ((unsigned long long)u[j] << 32) | u[j]; }
    const unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    const unsigned long long cc = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < UNROLL; ++r) {
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) {
                if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b), "f"(c));
                if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(bb), "l"(cc));
                if (OP == 2) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u[j]) : "r"(0x4B000000u), "r"(sel));
                if (OP == 3) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(b));
                if (OP == 4) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b), "f"(c));
                if (OP == 5) {  // the node-step mix: PRMT -> FFMA -> FMNMX
                    unsigned t;
                    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(u[j]), "r"(0x4B000000u), "r"(sel));
                    float f = __uint_as_float(t);
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(b), "f"(c));
                    asm volatile("max.f32 %0, %0, %1;" : "+f"(a[j]) : "f"(f));
                }
                if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[j]) : "r"(sel), "r"(0x22u));
                if (OP == 7) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[j]) : "r"(sel), "r"(0x22u));
                if (OP == 8) {  // FFMA + PRMT interleaved: do the two pipes issue in the same clock budget?
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[j]) : "f"(b), "f"(c));
                    asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u[j]) : "r"(0x4B000000u), "r"(sel));
                }
                if (OP == 9) {  // FFMA2 + PRMT + PRMT
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[j]) : "l"(bb), "l"(cc));
                    asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(u[j]) : "r"(0x4B000000u), "r"(sel));
                }
                if (OP == 10) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(u[j]) : "r"(sel));
                if (OP == 11) {  // cvt f16 -> f32 (HADD2.F32 on the fma pipe?)
                    unsigned short h = (unsigned short)u[j];
                    float f;
                    asm volatile("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
                    a[j] += f;  // FADD
                }
                if (OP == 12) {  // I2F of a 16-bit field
                    float f;
                    asm volatile("cvt.rn.f32.u16 %0, %1;" : "=f"(f) : "h"((unsigned short)u[j]));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[j]) : "f"(f), "f"(b));
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) s += a[j] + __uint_as_float(u[j]) + __uint_as_float((unsigned)p[j]) + __uint_as_float((unsigned)(p[j] >> 32));
    if (s == 123.456f) out[0] = s;
}

// stack traffic: per-lane stack in local memory vs [slot][thread] in shared memory, divergent stack pointers
template <int MODE>
__global__ void __launch_bounds__(128) stack_k(float* out, int iters)
{
    __shared__ int sh[16 * 128];
    int loc[64];
    unsigned rng = threadIdx.x * 747796405u + blockIdx.x * 2891336453u + 1u;
    int sp = 0, acc = 0;
    for (int i = 0; i < iters; ++i) {
        rng = rng * 1664525u + 1013904223u;
        const bool push = ((rng >> 16) & 1u) || sp == 0;
        if (push && sp < 15) {
            if (MODE == 0) loc[sp] = (int)rng; else sh[sp * 128 + threadIdx.x] = (int)rng;
            ++sp;
        } else if (sp > 0) {
            --sp;
            acc += MODE == 0 ? loc[sp] : sh[sp * 128 + threadIdx.x];
        }
    }
    if (acc == 12345) out[0] = (float)acc;
}

template <typename F>
float time_ms(F f)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int sms = pr.multiProcessorCount;
    float* out;
    cudaMalloc(&out, 4);
    const int iters = 2000;
    const char* names[] = {"FFMA", "FFMA2 (fma.rn.f32x2)", "PRMT", "FMNMX", "FMNMX3", "PRMT->FFMA->FMNMX (3 instr)", "LOP3", "IMAD", "FFMA+PRMT (2 instr)",
                           "FFMA2+PRMT (2 instr)", "HADD2", "cvt.f32.f16 + FADD (2 instr)", "I2F.U16 + FFMA (2 instr)"};
    const int per[] = {1, 1, 1, 1, 1, 3, 1, 1, 2, 2, 1, 2, 2};
    printf("%s, %d SMs, %.0f MHz nominal\n", pr.name, sms, khz / 1e3);
#define RUN(OP)                                                                                                      \
    {                                                                                                                \
        float ms = time_ms([&] { k<OP><<<sms, 1024>>>(out, iters, 1.0001f, 0.5f, 0x7410u); });                       \
        double winst = (double)sms * 32 * iters * UNROLL * CHAINS * per[OP];                                         \
        printf("%-32s %8.3f ms  %6.2f warp-instr/clk/SM (at nominal clock)\n", names[OP], ms, winst / (ms * 1e-3) / (khz * 1e3) / sms); \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
    float m0 = time_ms([&] { stack_k<0><<<sms * 8, 128>>>(out, 200000); });
    float m1 = time_ms([&] { stack_k<1><<<sms * 8, 128>>>(out, 200000); });
    printf("divergent per-lane stack, 200k push/pop per thread, 8 blocks x 128 per SM: local %.3f ms, shared [slot][thread] %.3f ms\n", m0, m1);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
