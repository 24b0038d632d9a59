// Measures the attainable FP32 FMA issue rate on this GPU (the secondary roofline for the
// Gaussian passes and the bit-exact flatten, which are FP32-pipe bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_peak ffma_peak.cu && ./ffma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, const float *wts, int iters, unsigned long long *cyc) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = threadIdx.x * 1e-3f + i;
    float x = out[threadIdx.x & 7], y = out[(threadIdx.x + 1) & 7];
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {  // 3-register FFMA
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = __fmaf_rn(acc[i], x, y);
        } else if (MODE == 1) {  // FFMA with a warp-uniform (constant-bank / uniform-register) operand
            float w = wts[it & 63];
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = __fmaf_rn(x, w, acc[i]);
        } else if (MODE == 3) {  // packed fma.rn.f32x2 (FFMA2): 16 instructions = 32 FMAs
            unsigned long long *a2 = reinterpret_cast<unsigned long long *>(acc);
            unsigned long long x2, y2;
            asm("mov.b64 %0, {%1, %1};" : "=l"(x2) : "f"(x));
            asm("mov.b64 %0, {%1, %1};" : "=l"(y2) : "f"(y));
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i]) : "l"(x2), "l"(y2));
        } else if (MODE == 4) {  // FFMA2 interleaved with one integer op per FFMA2 (do spare issue slots exist?)
            unsigned long long *a2 = reinterpret_cast<unsigned long long *>(acc);
            unsigned long long x2, y2;
            asm("mov.b64 %0, {%1, %1};" : "=l"(x2) : "f"(x));
            asm("mov.b64 %0, {%1, %1};" : "=l"(y2) : "f"(y));
            unsigned k0 = it, k1 = it + 1, k2 = it + 2, k3 = it + 3;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i]) : "l"(x2), "l"(y2));
                if ((i & 3) == 0) k0 = k0 * 3 + 1; else if ((i & 3) == 1) k1 = k1 * 5 + 1; else if ((i & 3) == 2) k2 ^= k0; else k3 += k1;
            }
            if ((k0 ^ k1 ^ k2 ^ k3) == 0x12345u) acc[0] += 1.0f;
        } else {  // separate FMUL + FADD (the bit-exact path)
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = __fadd_rn(acc[i], __fmul_rn(x, acc[(i + 7) & 31] ));
        }
    }
    unsigned long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char *name, int sms, float *out, float *wts, unsigned long long *cyc) {
    const int iters = 200000, blocks = sms * 8;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, 256>>>(out, wts, 1000, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(out, wts, iters, cyc);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    unsigned long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = (double)blocks * 256 * iters * 32 * (MODE == 2 ? 2 : 1);  // scalar FP32 operations
    double mhz = (double)c / (ms * 1e-3) / 1e6;
    printf("{\"kernel\": \"%s\", \"ms\": %.3f, \"fp32_instr_per_s\": %.4e, \"lanes_per_clk_per_sm\": %.2f, \"sm_mhz_est\": %.0f}\n",
           name, ms, ops / (ms * 1e-3), ops / (double)c / sms, mhz);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float *out, *wts; unsigned long long *cyc;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4 + 64);
    cudaMemset(out, 0, (size_t)p.multiProcessorCount * 8 * 256 * 4 + 64);
    cudaMalloc(&wts, 256); cudaMemset(wts, 0, 256);
    cudaMalloc(&cyc, 8);
    run<0>("ffma_3reg", p.multiProcessorCount, out, wts, cyc);
    run<1>("ffma_uniform_operand", p.multiProcessorCount, out, wts, cyc);
    run<2>("fmul_fadd", p.multiProcessorCount, out, wts, cyc);
    run<3>("ffma2_packed", p.multiProcessorCount, out, wts, cyc);
    run<4>("ffma2_plus_int", p.multiProcessorCount, out, wts, cyc);
    return 0;
}
