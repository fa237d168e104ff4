// tma3d_b.cu -- 3-D fp32 TMA through libcu++'s wrappers; argv: n pitch bx   (one case per process)
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, int z, int floats, float* out) {
    __shared__ alignas(128) float tile[32 * 16];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_3d_global_to_shared(tile, &map, x, y, z, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, floats * 4);
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < floats; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
    cuInit(0); cudaFree(0);
    const int n = argc > 1 ? atoi(argv[1]) : 32, pitch = argc > 2 ? atoi(argv[2]) : 32, bx = argc > 3 ? atoi(argv[3]) : 8;
    const int X = argc > 4 ? atoi(argv[4]) : 3, Y = 5, Z = 7;
    std::vector<float> h((size_t)pitch * n * n);
    for (int z = 0; z < n; ++z) for (int y = 0; y < n; ++y) for (int x = 0; x < pitch; ++x) h[((size_t)z * n + y) * pitch + x] = x + 100 * y + 10000 * z;
    float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 32 * 16 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap map{};
    cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)n}, strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * n * 4};
    cuuint32_t box[3] = {(cuuint32_t)bx, 4, 4}, es[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    k<<<1, 64>>>(map, X, Y, Z, bx * 16, out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(bx * 16); cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int z = 0; z < 4; ++z) for (int y = 0; y < 4; ++y) for (int x = 0; x < bx; ++x) bad += o[(z * 4 + y) * bx + x] != (X + x) + 100 * (Y + y) + 10000 * (Z + z);
    printf("3-D fp32 n %d pitch %d box %dx4x4 at x=%d (libcu++): encode %d, kernel %s, %d wrong\n", n, pitch, bx, X, (int)r, cudaGetErrorString(e), bad);
    return 0;
}
