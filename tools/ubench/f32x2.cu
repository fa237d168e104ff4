// Microbenchmark: issue cost of packed FFMA2 (fma.rn.f32x2) vs scalar FMUL/FADD on sm_100a, alone and
// interleaved with integer ALU work (does an FFMA2 take one issue slot or two?).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int NINT>
__global__ void k_scalar(float* out, float a, float b, int c) {
    float x[CH * 2]; int y[8];
    for (int i = 0; i < CH * 2; ++i) x[i] = a + i + threadIdx.x;
    for (int i = 0; i < 8; ++i) y[i] = c + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH * 2; ++i) { x[i] = x[i] * a; x[i] = x[i] + b; }       // 32 scalar FP32 instr
#pragma unroll
        for (int i = 0; i < NINT; ++i) { y[i % 8] = (y[i % 8] ^ c) + (y[(i + 1) % 8] >> 3); }   // ~3 ALU instr each
    }
    float s = 0;
    for (int i = 0; i < CH * 2; ++i) s += x[i];
    for (int i = 0; i < 8; ++i) s += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NINT>
__global__ void k_packed(float* out, float a, float b, int c) {
    float2 x[CH]; int y[8];
    const float2 A = make_float2(a, a), B = make_float2(b, b), Z = make_float2(-0.0f, -0.0f), ONE = make_float2(1.0f, 1.0f);
    for (int i = 0; i < CH; ++i) x[i] = make_float2(a + 2 * i + threadIdx.x, a + 2 * i + 1 + threadIdx.x);
    for (int i = 0; i < 8; ++i) y[i] = c + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) { x[i] = fma2(x[i], A, Z); x[i] = fma2(x[i], ONE, B); }   // 16 FFMA2 = same 32 roundings
#pragma unroll
        for (int i = 0; i < NINT; ++i) { y[i % 8] = (y[i % 8] ^ c) + (y[(i + 1) % 8] >> 3); }
    }
    float s = 0;
    for (int i = 0; i < CH; ++i) s += x[i].x + x[i].y;
    for (int i = 0; i < 8; ++i) s += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NINT>
__global__ void k_packed_ma(float* out, float a, float b, int c) {
    float2 x[CH]; int y[8];
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int i = 0; i < CH; ++i) x[i] = make_float2(a + 2 * i + threadIdx.x, a + 2 * i + 1 + threadIdx.x);
    for (int i = 0; i < 8; ++i) y[i] = c + i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) { x[i] = __fmul2_rn(x[i], A); x[i] = __fadd2_rn(x[i], B); }   // 8 FMUL2 + 8 FADD2
#pragma unroll
        for (int i = 0; i < NINT; ++i) { y[i % 8] = (y[i % 8] ^ c) + (y[(i + 1) % 8] >> 3); }
    }
    float s = 0;
    for (int i = 0; i < CH; ++i) s += x[i].x + x[i].y;
    for (int i = 0; i < 8; ++i) s += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 4, threads = 512;
    float* out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double warps = (double)blocks * threads / 32, clk = khz * 1e3;
#define RUN(NAME, K, NINT)                                                                                    \
    { float t = timeit([&] { K<NINT><<<blocks, threads>>>(out, 1.0001f, 0.5f, 12345); });                      \
      printf("{\"kernel\": \"%s\", \"int_groups\": %d, \"ms\": %.4f, \"cycles_per_iter_per_smsp_warp\": %.2f}\n", NAME, NINT, t, \
             t * 1e-3 * clk / ITERS / (warps / (sms * 4))); }
    RUN("scalar 32 FMUL/FADD", k_scalar, 0) RUN("packed 16 FFMA2", k_packed, 0)
    RUN("scalar 32 FMUL/FADD", k_scalar, 8) RUN("packed 16 FFMA2", k_packed, 8)
    RUN("scalar 32 FMUL/FADD", k_scalar, 16) RUN("packed 16 FFMA2", k_packed, 16)
    RUN("packed 8 FMUL2 + 8 FADD2", k_packed_ma, 0) RUN("packed 8 FMUL2 + 8 FADD2", k_packed_ma, 8) RUN("packed 8 FMUL2 + 8 FADD2", k_packed_ma, 16)
    printf("{\"sms\": %d, \"clock_khz\": %d}\n", sms, khz);
    return 0;
}
