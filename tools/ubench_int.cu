// Micro-benchmark: per-SM throughput of the integer instructions the ME kernel leans on (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int tools/ubench_int.cu && ./ubench_int
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

#define CHAINS 8
#define ITERS 4096

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(unsigned *out, unsigned seed) {
    unsigned a[CHAINS], b = seed + threadIdx.x, c = seed * 3 + 1;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) a[i] = __dp4a(b, c, a[i]);                       // IDP.4A.U8.U8
            if (OP == 1) a[i] = __vabsdiffu4(a[i], b);                     // VABSDIFF4
            if (OP == 2) a[i] = a[i] * b + c;                              // IMAD
            if (OP == 3) a[i] = __dp2a_lo(b, c, a[i]);                     // IDP.2A
            if (OP == 4) a[i] = (a[i] + b) ^ c;                            // IADD3/LOP3 pair -> 2 instr
            if (OP == 5) a[i] = __funnelshift_r(a[i], b, 8);               // SHF
            if (OP == 6) a[i] = __byte_perm(a[i], b, 0x5410 + (c & 1));    // PRMT
            if (OP == 7) a[i] = min(a[i] + 1u, b);                         // VIADDMNMX?
            if (OP == 8) { float f = __uint_as_float(a[i]); f = fmaf(f, 1.0001f, 0.5f); a[i] = __float_as_uint(f); }  // FFMA
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mma.sync m16n8k32 u8*u8 -> s32
__global__ void __launch_bounds__(1024, 1) k_mma(unsigned *out, unsigned seed) {
    int c[4][4];
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = seed, b1 = seed * 9;
#pragma unroll
    for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    int s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (unsigned)s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    unsigned *out; cudaMalloc(&out, sizeof(unsigned) * sms * 1024);
    const char *names[] = {"IDP.4A", "VABSDIFF4", "IMAD", "IDP.2A", "IADD3+LOP3 (2 instr)", "SHF", "PRMT", "VIADDMNMX", "FFMA"};
    printf("SMs %d, nominal clock %.0f MHz (rates below assume the nominal clock)\n", sms, clk_khz / 1e3);
#define RUN(OP) { float ms = timeit([&] { k<OP><<<sms, 1024>>>(out, 12345u); }); \
    double ops = (double)sms * 1024 * CHAINS * ITERS; \
    printf("%-22s %8.3f ms  %6.1f lane-ops/clk/SM\n", names[OP], ms, ops / (ms * 1e-3) / sms / (clk_khz * 1e3)); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8)
    {
        float ms = timeit([&] { k_mma<<<sms, 1024>>>(out, 12345u); });
        double mmas = (double)sms * 32 * 4 * ITERS;   // warp-level mma instructions
        printf("%-22s %8.3f ms  %6.1f MAC/clk/SM (m16n8k32 = 4096 MAC)\n", "mma.sync u8 m16n8k32", ms, mmas * 4096 / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    return 0;
}
