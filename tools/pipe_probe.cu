// pipe_probe.cu -- measures issue rates (warp-instructions per cycle per SM) of the integer / SIMD-video
// instructions the detection kernel is built from, alone and in pairs, to find which ones share a pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITER 2048
#define CHAINS 8

template <int OP>
__device__ __forceinline__ void op(uint32_t &a, uint32_t b, uint32_t c) {
    if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 3) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(0u));
    if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x6543;" : "+r"(a) : "r"(b));
    if (OP == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(a) : "r"(b));
    if (OP == 6) asm volatile("{.reg .u32 t; min.s32 t, %0, %1; min.s32 %0, t, %2;}" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 7) asm volatile("popc.b32 %0, %0;" : "+r"(a));
    if (OP == 8) asm volatile("bfind.u32 %0, %0;" : "+r"(a));
    if (OP == 9) asm volatile("shl.b32 %0, %0, 3;" : "+r"(a));
    if (OP == 10) asm volatile("shr.u32 %0, %0, 3;" : "+r"(a));
    if (OP == 11) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 12) asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 13) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 14) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 15) asm volatile("add.u32 %0, %0, 12345;" : "+r"(a));
    if (OP == 16) asm volatile("mad.lo.u32 %0, %0, 8, %1;" : "+r"(a) : "r"(b));
    if (OP == 17) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 18) asm volatile("vmin2.s32.s32.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(0u));
    if (OP == 19) asm volatile("{.reg .pred p; setp.ne.u32 p, %0, %1; @p add.u32 %0, %0, %2;}" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 20) a = __vimin3_u16x2(a, b, c);               // VIMNMX3.U16x2
    if (OP == 21) a = __vminu2(a, b);                        // VIMNMX.U16x2
    if (OP == 22) a = __viaddmax_s16x2_relu(a, b, c);        // VIADDMNMX.S16x2.RELU
    if (OP == 23) a = __vimax3_u32(a, b, c);                 // VIMNMX3.U32
    if (OP == 24) a = __dp4a(a, b, c);                       // IDP.4A
    if (OP == 25) a = (uint32_t)__viaddmax_s32((int)a, (int)b, (int)c);  // VIADDMNMX
    // round 2: could the 16-bit-lane min/max of phase B run on the floating-point pipes instead of the logic pipe?
    if (OP == 26) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));                       // HMNMX2
    if (OP == 27) asm volatile("fma.rn.relu.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));   // HFMA2.RELU
    if (OP == 28) asm volatile("add.f16x2 %0, %0, %1;" : "+r"(a) : "r"(b));                       // HADD2
    if (OP == 29) asm volatile("min.bf16x2 %0, %0, %1;" : "+r"(a) : "r"(b));                      // HMNMX2.BF16
    if (OP == 30) asm volatile("{.reg .f32 t; min.f32 t, %0, %1; mov.b32 %0, t;}" : "+r"(a) : "r"(b));  // FMNMX
}

template <int OPA, int OPB>
__global__ void __launch_bounds__(1024) probe(uint32_t *out, uint32_t b, uint32_t c, long long *cycles) {
    uint32_t v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) v[i] = threadIdx.x * 7 + i;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            op<OPA>(v[i], b, c);
            if (OPB >= 0) op<OPB < 0 ? 0 : OPB>(v[(i + 4) % CHAINS], c, b);
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OPA, int OPB>
void run(const char *name, uint32_t *out, long long *cyc) {
    probe<OPA, OPB><<<148, 1024>>>(out, 0x01020304u, 0x7f7f7f7fu, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    const double winst = 32.0 * ITER * CHAINS * (OPB >= 0 ? 2 : 1);  // warp-instructions per SM (32 warps)
    printf("%-28s %7.3f warp-inst/clk/SM   (%.2f per SMSP)  err=%s\n", name, winst / avg, winst / avg / 4,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    run<0, -1>("LOP3", out, cyc);
    run<1, -1>("IADD (reg)", out, cyc);
    run<15, -1>("IADD (imm)", out, cyc);
    run<2, -1>("IMAD (reg)", out, cyc);
    run<16, -1>("IMAD (imm mul)", out, cyc);
    run<3, -1>("VABSDIFF4", out, cyc);
    run<14, -1>("VABSDIFF4.ACC", out, cyc);
    run<4, -1>("PRMT", out, cyc);
    run<5, -1>("SHF.L.W imm", out, cyc);
    run<17, -1>("SHF.R.W reg", out, cyc);
    run<9, -1>("SHL imm", out, cyc);
    run<10, -1>("SHR imm", out, cyc);
    run<6, -1>("VIMNMX3 (min3)", out, cyc);
    run<18, -1>("vmin2 (VIMNMX.S16x2)", out, cyc);
    run<20, -1>("VIMNMX3.U16x2", out, cyc);
    run<21, -1>("VIMNMX.U16x2", out, cyc);
    run<22, -1>("VIADDMNMX.S16x2.RELU", out, cyc);
    run<23, -1>("VIMNMX3.U32", out, cyc);
    run<24, -1>("IDP.4A", out, cyc);
    run<25, -1>("VIADDMNMX.S32", out, cyc);
    run<20, 2>("VIMNMX3.U16x2 + IMAD", out, cyc);
    run<24, 0>("IDP.4A + LOP3", out, cyc);
    run<24, 2>("IDP.4A + IMAD", out, cyc);
    run<26, -1>("HMNMX2 (min.f16x2)", out, cyc);
    run<29, -1>("HMNMX2.BF16", out, cyc);
    run<30, -1>("FMNMX", out, cyc);
    run<27, -1>("HFMA2.RELU", out, cyc);
    run<28, -1>("HADD2", out, cyc);
    run<26, 0>("HMNMX2 + LOP3", out, cyc);
    run<26, 2>("HMNMX2 + IMAD", out, cyc);
    run<26, 20>("HMNMX2 + VIMNMX3.U16x2", out, cyc);
    run<27, 0>("HFMA2.RELU + LOP3", out, cyc);
    run<27, 2>("HFMA2.RELU + IMAD", out, cyc);
    run<27, 13>("HFMA2.RELU + FFMA", out, cyc);
    run<7, -1>("POPC", out, cyc);
    run<8, -1>("FLO (bfind)", out, cyc);
    run<11, -1>("IMAD.HI", out, cyc);
    run<12, -1>("ISETP+SEL", out, cyc);
    run<19, -1>("ISETP+@p IADD", out, cyc);
    run<13, -1>("FFMA", out, cyc);
    run<0, 2>("LOP3 + IMAD", out, cyc);
    run<0, 3>("LOP3 + VABSDIFF4", out, cyc);
    run<0, 4>("LOP3 + PRMT", out, cyc);
    run<0, 5>("LOP3 + SHF", out, cyc);
    run<0, 1>("LOP3 + IADD", out, cyc);
    run<0, 13>("LOP3 + FFMA", out, cyc);
    run<2, 13>("IMAD + FFMA", out, cyc);
    run<3, 2>("VABSDIFF4 + IMAD", out, cyc);
    run<3, 4>("VABSDIFF4 + PRMT", out, cyc);
    run<0, 6>("LOP3 + VIMNMX3", out, cyc);
    run<0, 7>("LOP3 + POPC", out, cyc);
    run<2, 11>("IMAD + IMAD.HI", out, cyc);
    run<0, 11>("LOP3 + IMAD.HI", out, cyc);
    return 0;
}
