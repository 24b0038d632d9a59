// Isolates the Gaussian inner loop (N outputs x 4 channels per thread, rotating weight window,
// one LDS.128 input + one LDS.64 weight per step, FFMA2 math) from all staging, barriers and
// stores, at several occupancies, to find out what the FMA pipe can sustain with this operand
// pattern.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gauss_loop_bench gauss_loop_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int MODE>
__global__ void __launch_bounds__(256) k(float *out, int steps, int reps) {
    extern __shared__ __align__(16) unsigned char smem[];
    float2 *wsm = reinterpret_cast<float2 *>(smem);           // 512 weights
    float4 *tile = reinterpret_cast<float4 *>(smem + 4096);   // (steps + N) rows x 32 lanes per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) wsm[i] = make_float2(1e-3f * i, 1e-3f * i);
    float4 *mine = tile;  // all warps read one tile (smem budget)
    (void)warp;
    for (int i = threadIdx.x; i < (steps + N) * 32; i += blockDim.x) mine[i] = make_float4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    float2 lo[N], hi[N];
#pragma unroll
    for (int j = 0; j < N; j++) lo[j] = hi[j] = make_float2(0.f, 0.f);
    float2 R[N];
    for (int rep = 0; rep < reps; rep++) {
#pragma unroll
        for (int m = 0; m < N - 1; m++) R[m] = wsm[m];
        const float4 *col = mine + lane;
        for (int g = 0; g < steps; g += N) {
#pragma unroll
            for (int s = 0; s < N; s++) {
                R[(s + N - 1) % N] = wsm[g + s + N - 1];
                const float4 in = col[(size_t)(g + s) * 32];
#pragma unroll
                for (int j = 0; j < N; j++) {
                    const float2 w2 = R[(s - j - 1 + 2 * N) % N];
                    if (MODE == 0) {
                        lo[j] = __ffma2_rn(make_float2(in.x, in.y), w2, lo[j]);
                        hi[j] = __ffma2_rn(make_float2(in.z, in.w), w2, hi[j]);
                    } else {
                        lo[j].x = __fmaf_rn(in.x, w2.x, lo[j].x); lo[j].y = __fmaf_rn(in.y, w2.x, lo[j].y);
                        hi[j].x = __fmaf_rn(in.z, w2.x, hi[j].x); hi[j].y = __fmaf_rn(in.w, w2.x, hi[j].y);
                    }
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < N; j++) s += lo[j].x + lo[j].y + hi[j].x + hi[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N, int MODE>
void run(const char *name, int sms, int warps_per_block, int blocks_per_sm, float *out) {
    const int steps = 32, reps = 1600;
    size_t smem = 4096 + (size_t)(steps + N) * 32 * 16;
    cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<N, MODE>, warps_per_block * 32, smem);
    if (occ < blocks_per_sm) { printf("{\"kernel\": \"%s\", \"skipped\": \"occupancy %d < %d\"}\n", name, occ, blocks_per_sm); return; }
    // pad smem so exactly blocks_per_sm fit
    size_t pad = (224 * 1024) / blocks_per_sm;
    if (pad > smem) { smem = pad - 2048; cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }
    const int blocks = sms * blocks_per_sm;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<N, MODE><<<blocks, warps_per_block * 32, smem>>>(out, steps, 4);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<N, MODE><<<blocks, warps_per_block * 32, smem>>>(out, steps, reps);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    double fma = (double)blocks * warps_per_block * 32 * reps * steps * N * 4;
    printf("{\"kernel\": \"%s\", \"N\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"fma_per_s\": %.4e, \"err\": \"%s\"}\n", name, N,
           warps_per_block * blocks_per_sm, ms, fma / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float *out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    const int sms = p.multiProcessorCount;
    run<8, 0>("ffma2_N8", sms, 4, 1, out); run<8, 0>("ffma2_N8", sms, 4, 2, out); run<8, 0>("ffma2_N8", sms, 4, 4, out); run<8, 0>("ffma2_N8", sms, 4, 6, out);
    run<8, 0>("ffma2_N8", sms, 8, 2, out); run<8, 0>("ffma2_N8", sms, 8, 1, out);
    run<16, 0>("ffma2_N16", sms, 4, 2, out); run<16, 0>("ffma2_N16", sms, 4, 4, out); run<16, 0>("ffma2_N16", sms, 8, 2, out);
    run<4, 0>("ffma2_N4", sms, 4, 4, out); run<4, 0>("ffma2_N4", sms, 4, 8, out);
    run<8, 1>("ffma_scalar_N8", sms, 4, 2, out); run<8, 1>("ffma_scalar_N8", sms, 4, 4, out); run<8, 1>("ffma_scalar_N8", sms, 4, 6, out);
    run<16, 1>("ffma_scalar_N16", sms, 4, 4, out);
    return 0;
}
