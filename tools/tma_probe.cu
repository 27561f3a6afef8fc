// tma_probe.cu -- bring-up probe for the TMA / mbarrier plumbing (diagnostic, not part of the product).
// usage: tma_probe <variant>   variants: 0 = mbarrier only, 1 = 2-D load, 2 = 3-D load, 3 = 3-D load + prefetch
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK, bool PREFETCH>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, uint8_t *out, int x, int y, int z, int rows, int do_tma) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16384);
    if (threadIdx.x == 0) {
        if (PREFETCH) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (do_tma) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(rows * 256) : "memory");
            if (RANK == 3)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
        }
        uint32_t ok = 0, spins = 0;
        while (!ok && spins < (1u << 20)) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
            spins++;
        }
        if (threadIdx.x == 0 && !ok) printf("probe: mbarrier wait timed out\n");
    }
    for (int i = threadIdx.x; i < rows * 256; i += blockDim.x) out[i] = smem[i];
    if (threadIdx.x == 0) printf("probe: smem base 0x%x\n", smem_u32(smem));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 304, H = 200, F = 2, ROWS = 22;
    std::vector<uint8_t> img((size_t)W * H * F);
    for (size_t i = 0; i < img.size(); i++) img[i] = (uint8_t)(i * 7 + (i >> 8));
    uint8_t *d_img, *d_out;
    CK(cudaMalloc(&d_img, img.size()));
    CK(cudaMalloc(&d_out, ROWS * 256));
    CK(cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    printf("entry point %p query %d\n", fn, (int)q);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const int rank = variant == 1 ? 2 : 3;
    cuuint64_t dims[3] = {300, (cuuint64_t)H, (cuuint64_t)F};
    cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {256, (cuuint32_t)ROWS, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult cr = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d_img, dims, strides, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rank %d -> %d\n", rank, (int)cr);
    const int x = -8, y = -3, z = 1;
    const size_t smem = 16384 + 64;
    if (variant == 0) { CK(cudaFuncSetAttribute(probe<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe<3, false><<<1, 256, smem>>>(tmap, d_out, x, y, z, ROWS, 0); }
    if (variant == 1) { CK(cudaFuncSetAttribute(probe<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe<2, false><<<1, 256, smem>>>(tmap, d_out, x, y, 0, ROWS, 1); }
    if (variant == 2) { CK(cudaFuncSetAttribute(probe<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe<3, false><<<1, 256, smem>>>(tmap, d_out, x, y, z, ROWS, 1); }
    if (variant == 3) { CK(cudaFuncSetAttribute(probe<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); probe<3, true><<<1, 256, smem>>>(tmap, d_out, x, y, z, ROWS, 1); }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<uint8_t> out(ROWS * 256);
    CK(cudaMemcpy(out.data(), d_out, out.size(), cudaMemcpyDeviceToHost));
    if (variant > 0) {
        int bad = 0;
        const int zz = variant == 1 ? 0 : z;
        for (int r = 0; r < ROWS; r++)
            for (int j = 0; j < 256; j++) {
                int yy = y + r, xx = x + j;
                uint8_t want = (yy >= 0 && yy < H && xx >= 0 && xx < 300) ? img[(size_t)zz * W * H + (size_t)yy * W + xx] : 0;
                if (out[r * 256 + j] != want) bad++;
            }
        printf("variant %d: %d mismatching bytes\n", variant, bad);
    } else {
        printf("variant 0 done\n");
    }
    return 0;
}
