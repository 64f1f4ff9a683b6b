// FP64 roofline denominators for B200 (sm_100a): pure-DFMA and DMMA (mma.sync m8n8k4 f64)
// micro-kernels, timed with CUDA events.  MEASURED_PEAKS.json has no FP64 entry, so this
// tool supplies the denominator used by bench.py's roofline (see DESIGN.md "Measurement").
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
    double av = a + threadIdx.x * 1e-12, bv = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(av), "d"(bv));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int threads = warps * 32, blocks = sms * 2;
        float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 8 * iters * (double)threads * blocks;
        printf(", \"dfma_tflops_w%d\": %.3f", warps * 2, fl / ms * 1e-9);
        ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        fl = 2.0 * 256 * 8 * iters * (double)warps * blocks;
        printf(", \"dmma_tflops_w%d\": %.3f", warps * 2, fl / ms * 1e-9);
    }
    // sustained (~2 s) DMMA
    {
        int threads = 512, blocks = sms * 2;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        int n = 60;
        for (int r = 0; r < n; r++) dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 256 * 8 * iters * 16.0 * blocks * n;
        printf(", \"dmma_tflops_sustained\": %.3f, \"sustained_ms\": %.1f", fl / ms * 1e-9, ms);
    }
    printf("}\n");
    return 0;
}
