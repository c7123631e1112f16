// dmma_bench.cu -- is the FP64 tensor pipe (DMMA) a faster home for the dense k = 5 gate than the
// FP64 FMA pipe?  (VERDICT r1 item 8 / BASELINE north star: "tensor-core (DMMA/tcgen05) complex
// matvec only where the fused matrix is large enough to be a real dense contraction, justified by
// tensor-pipe counters".)
//
// A k = 5 complex128 gate is, per group of 32 amplitudes, a real 64x64 by 64-vector product: over
// all groups a GEMM C[64 x G] = A[64 x 64] B[64 x G] with 256 flops and 32 bytes per amplitude
// (intensity 8 flop/B: 52 TFLOP/s of FP64 at the measured 6.55 TB/s).  tcgen05 has no FP64 kind;
// the FP64 tensor path of sm_100a is mma.sync m8n8k4 (SASS DMMA).  This program measures, on
// registers only (no memory traffic: the upper bound of either pipe):
//   1. DMMA  m8n8k4 f64 issue rate with ILP 1..8 independent accumulators per warp;
//   2. DFMA  issue rate with the same structure;
// and prints TFLOP/s for the chip.  Run it under ncu with
//   --metrics sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,
//             smsp__inst_executed_pipe_tensor_op_dmma.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg
// for the tensor-pipe counters (tools/r2_dmma.sh).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) k_dmma(double *out, int iters) {
    double c[ILP][2];
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters) {
    double c[ILP];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = double(i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(c[i]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 4096;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    printf("{\"device\": \"%s\", \"sms\": %d, \"results\": [\n", prop.name, sms);
    const double warps = double(blocks) * threads / 32.0;
#define RUN_DMMA(ILP)                                                                                         \
    {                                                                                                         \
        float ms = time_ms([&] { k_dmma<ILP><<<blocks, threads>>>(out, iters); });                           \
        double flops = warps * double(iters) * ILP * (2.0 * 8 * 8 * 4);                                       \
        printf("  {\"pipe\": \"dmma m8n8k4 f64\", \"ilp\": %d, \"ms\": %.3f, \"tflops\": %.2f},\n", ILP, ms, flops / ms / 1e9); \
    }
#define RUN_DFMA(ILP)                                                                                         \
    {                                                                                                         \
        float ms = time_ms([&] { k_dfma<ILP><<<blocks, threads>>>(out, iters); });                           \
        double flops = warps * 32.0 * double(iters) * ILP * 2.0;                                              \
        printf("  {\"pipe\": \"dfma\", \"ilp\": %d, \"ms\": %.3f, \"tflops\": %.2f},\n", ILP, ms, flops / ms / 1e9); \
    }
    RUN_DMMA(1) RUN_DMMA(2) RUN_DMMA(4) RUN_DMMA(8)
    RUN_DFMA(1) RUN_DFMA(2) RUN_DFMA(4) RUN_DFMA(8)
    printf("  {\"note\": \"registers only, 8 CTAs x 256 threads per SM\"}\n]}\n");
    cudaFree(out);
    return 0;
}
