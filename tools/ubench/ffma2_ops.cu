// Microbenchmark: FFMA2 issue rate on sm_100a by operand shape (how many of the three sources are
// 64-bit register pairs vs broadcast scalars), and FRND / mixed streams.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o ffma2_ops ffma2_ops.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define CH 8
struct P { float* out; float s, u; float2 e, f; };

template <int MODE>
__global__ void k(const __grid_constant__ P p) {
    float2 x[CH];
    float2 e = p.e, f = p.f;
    e.x += threadIdx.x; f.y += threadIdx.x;
    for (int i = 0; i < CH; ++i) x[i] = make_float2(p.s + 2 * i + threadIdx.x, p.u + i);
    float y[CH];
    for (int i = 0; i < CH; ++i) y[i] = p.s * i + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) x[i] = __ffma2_rn(x[i], make_float2(p.s, p.s), make_float2(p.u, p.u));   // 1 pair
            if (MODE == 1) x[i] = __ffma2_rn(x[i], make_float2(p.s, p.s), e);                        // 2 pairs
            if (MODE == 2) x[i] = __ffma2_rn(x[i], e, f);                                            // 3 pairs
            if (MODE == 3) { x[i] = __ffma2_rn(x[i], e, f); y[i] = fmaf(y[i], p.s, p.u); }           // FFMA2 + FFMA
            if (MODE == 4) { y[i] = fmaf(y[i], p.s, p.u); }                                          // FFMA only
            if (MODE == 5) { y[i] = floorf(y[i] * p.s); }                                            // FMUL + FRND
            if (MODE == 6) { x[i] = __ffma2_rn(x[i], e, f); x[i].x = fmaf(x[i].x, p.s, p.u); }        // FFMA2 then scalar on a lane
            if (MODE == 7) { x[i] = __ffma2_rn(x[i], make_float2(y[i], y[i]), f); y[i] = y[i] * p.s; }  // R.F32 broadcast of a varying reg
        }
    }
    float s = 0;
    for (int i = 0; i < CH; ++i) s += x[i].x + x[i].y + y[i];
    p.out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount, blocks = sms * 4, threads = 512;
    P p; cudaMalloc(&p.out, sizeof(float) * blocks * threads);
    p.s = 1.0001f; p.u = 0.5f; p.e = make_float2(0.999f, 1.001f); p.f = make_float2(0.25f, 0.125f);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double warps = (double)blocks * threads / 32, clk = khz * 1e3;
    const char* names[] = {"8 FFMA2 (1 pair, 2 bcast)", "8 FFMA2 (2 pairs, 1 bcast)", "8 FFMA2 (3 pairs)", "8 FFMA2 (3 pairs) + 8 FFMA",
                           "8 FFMA", "8 FMUL + 8 FRND", "8 FFMA2 + 8 dependent FFMA on a lane", "8 FFMA2 (R.F32 bcast) + 8 FMUL"};
#define RUN(M) { float t = timeit([&] { k<M><<<blocks, threads>>>(p); }); \
      printf("{\"mode\": %d, \"what\": \"%s\", \"ms\": %.4f, \"cycles_per_iter_per_smsp_warp\": %.2f}\n", M, names[M], t, \
             t * 1e-3 * clk / ITERS / (warps / (sms * 4))); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    return 0;
}
