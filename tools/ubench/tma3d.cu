// tma3d.cu -- does a 3-D fp32 TMA box of BX x 4 x 4 texels land in shared memory as expected?  (the staging load of
// csrc/native/app_clouds_tex_native.h in isolation)   nvcc -arch=sm_100a tma3d.cu -lcuda -o tma3d && ./tma3d
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int BX> __global__ void k(const __grid_constant__ CUtensorMap map, const CUtensorMap* gmap, float* out, int x, int y, int z, int use_global) {
    __shared__ __align__(128) float tile[BX * 16];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "n"(BX * 16 * 4) : "memory");
        const void* m = use_global ? (const void*)gmap : (const void*)&map;
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(d), "l"(m), "r"(x), "r"(y), "r"(z), "r"(b) : "memory");
    }
    __syncthreads();
    unsigned done = 0;
    while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b) : "memory");
    for (int i = threadIdx.x; i < BX * 16; i += blockDim.x) out[i] = tile[i];
}

template <int BX> int run(int use_global) {
    const int n = 18, pitch = 20;
    std::vector<float> h((size_t)pitch * n * n);
    for (int z = 0; z < n; ++z) for (int y = 0; y < n; ++y) for (int x = 0; x < pitch; ++x) h[((size_t)z * n + y) * pitch + x] = x + 100 * y + 10000 * z;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, BX * 16 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)n}, strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * n * 4};
    cuuint32_t box[3] = {BX, 4, 4}, es[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap* gmap; cudaMalloc(&gmap, sizeof map); cudaMemcpy(gmap, &map, sizeof map, cudaMemcpyHostToDevice);
    k<BX><<<1, 32>>>(map, gmap, out, 3, 5, 7, use_global);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(BX * 16);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int z = 0; z < 4; ++z) for (int y = 0; y < 4; ++y) for (int x = 0; x < BX; ++x) bad += o[(z * 4 + y) * BX + x] != (3 + x) + 100 * (5 + y) + 10000 * (7 + z);
    printf("box %d x 4 x 4, descriptor in %s: encode %d, kernel %s, %d wrong texels (first %g)\n", BX, use_global ? "global" : "param", (int)r, cudaGetErrorString(e), bad, o[0]);
    return e != cudaSuccess;
}

int main(int argc, char** argv) {
    cuInit(0);
    cudaFree(0);
    const int bx = argc > 1 ? atoi(argv[1]) : 4, g = argc > 2 ? atoi(argv[2]) : 0;
    return bx == 4 ? run<4>(g) : bx == 8 ? run<8>(g) : bx == 16 ? run<16>(g) : run<32>(g);   // a dead context after a fault: one case per process
}
