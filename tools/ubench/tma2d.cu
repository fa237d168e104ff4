// tma2d.cu -- the CUDA programming guide's 2-D TMA example (cuda::barrier + cde::cp_async_bulk_tensor_2d_global_to_shared), as a
// sanity check that tensor TMA works on the box at all.   nvcc -arch=sm_100a tma2d.cu -lcuda -o tma2d
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 256, GH = 256, SW = 32, SH = 32;

__global__ void k(const __grid_constant__ CUtensorMap map, int x, int y, int* out) {
    __shared__ alignas(128) int tile[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&tile, &map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(tile));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) out[i] = tile[i / SW][i % SW];
}

int main() {
    cuInit(0); cudaFree(0);
    std::vector<int> h(GW * GH);
    for (int i = 0; i < GW * GH; ++i) h[i] = i;
    int *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, SW * SH * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap map{};
    cuuint64_t size[2] = {GW, GH}, stride[1] = {GW * sizeof(int)};
    cuuint32_t box[2] = {SW, SH}, es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    k<<<1, 128>>>(map, 64, 32, out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int> o(SW * SH); cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < SH; ++y) for (int x = 0; x < SW; ++x) bad += o[y * SW + x] != (32 + y) * GW + 64 + x;
    printf("2-D int32 32x32 box: encode %d, kernel %s, %d wrong\n", (int)r, cudaGetErrorString(e), bad);
    return 0;
}
