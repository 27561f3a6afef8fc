// tma_probe2.cu -- second bring-up probe (diagnostic only).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int ROWS = 22;

// variant 4: 1-D bulk copy
__global__ void k_bulk1d(const uint8_t *src, uint8_t *out) {
    __shared__ alignas(128) uint8_t buf[4096];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(buf)), "l"(src), "r"(4096), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0, spins = 0;
    while (!ok && spins < (1u << 20)) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        spins++;
    }
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = buf[i];
}

// variant 5: the CUDA programming guide's libcu++ example
template <int BW>
__global__ void k_libcu(const __grid_constant__ CUtensorMap tensor_map, uint8_t *out, int x, int y) {
    __shared__ alignas(128) uint8_t smem_buffer[ROWS][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) {
        init(&bar, blockDim.x);
        cde::fence_proxy_async_shared_cta();
    }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else {
        token = bar.arrive();
    }
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < ROWS * BW; i += blockDim.x) out[i] = (&smem_buffer[0][0])[i];
}

// variant 6: descriptor in global memory
__global__ void k_gmem_desc(const CUtensorMap *tmap, uint8_t *out, int x, int y) {
    __shared__ alignas(128) uint8_t buf[ROWS * 256];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(ROWS * 256) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(buf)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0, spins = 0;
    while (!ok && spins < (1u << 20)) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        spins++;
    }
    for (int i = threadIdx.x; i < ROWS * 256; i += blockDim.x) out[i] = buf[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 4;
    const int W = 304, H = 200;
    std::vector<uint8_t> img((size_t)W * H);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d_img, *d_out;
    CK(cudaMalloc(&d_img, img.size()));
    CK(cudaMalloc(&d_out, ROWS * 256));
    CK(cudaMemset(d_out, 0xEE, ROWS * 256));
    CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    int bw = (variant == 8 || variant == 9) ? 64 : 256;
    int x = (variant == 7 || variant == 8) ? 16 : -8, y = (variant == 7 || variant == 8) ? 5 : -3;
    if (variant == 5 || variant == 9) { x = 16; y = 5; }
    if (variant == 10) { x = -8; y = -3; }
    CUtensorMap tmap;
    cuuint64_t dims[2] = {300, (cuuint64_t)H};
    cuuint64_t strides[1] = {(cuuint64_t)W};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)ROWS};
    cuuint32_t es[2] = {1, 1};
    CUresult cr = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_img, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode -> %d (box %d x %d at %d,%d)\n", variant, (int)cr, bw, ROWS, x, y);
    if (variant == 4) k_bulk1d<<<1, 128>>>(d_img, d_out);
    if (variant == 5 || variant == 10) k_libcu<256><<<1, 128>>>(tmap, d_out, x, y);
    if (variant == 9) k_libcu<64><<<1, 128>>>(tmap, d_out, x, y);
    if (variant == 6 || variant == 7 || variant == 8) {
        CUtensorMap *d_map;
        CK(cudaMalloc(&d_map, sizeof(CUtensorMap)));
        CK(cudaMemcpy(d_map, &tmap, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
        if (variant == 8) { printf("variant 8 not implemented for gmem kernel with bw 64\n"); return 0; }
        k_gmem_desc<<<1, 128>>>(d_map, d_out, x, y);
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> out(ROWS * 256);
    CK(cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost));
    int bad = 0;
    if (variant == 4) {
        for (int i = 0; i < 4096; i++) bad += out[i] != img[i];
    } else {
        for (int r = 0; r < ROWS; r++)
            for (int j = 0; j < bw; j++) {
                int yy = y + r, xx = x + j;
                uint8_t want = (yy >= 0 && yy < H && xx >= 0 && xx < 300) ? img[(size_t)yy * W + xx] : 0;
                bad += out[r * bw + j] != want;
            }
    }
    printf("variant %d: %d mismatching bytes\n", variant, bad);
    return 0;
}
