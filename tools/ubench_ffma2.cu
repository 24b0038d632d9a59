// Micro-benchmark for DESIGN.md section 9, item 1: what paces a stream of packed FMAs shaped like the Gaussian inner
// loop - `acc[j] = fma2(in, w[j], acc[j])`, 8 accumulators per input pair - when the weight operand comes from
//   (a) ordinary registers                (what gauss_h_kernel / gauss_v_tile_kernel do on main),
//   (b) a __grid_constant__ table, i.e. uniform registers (what the UW = true kernels of this branch do).
// Build and run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/ubench_ffma2 tools/ubench_ffma2.cu && /tmp/ubench_ffma2
// Prints cycles per FFMA2 per SM sub-partition for 1, 2 and 4 resident warps per sub-partition (2.0 = the pipe's
// issue rate for a 64-lane operation; anything above is operand delivery or dependency stalls).
#include <cstdio>
#include <cuda_runtime.h>

struct Table {
    float2 w[64];
};

constexpr int kIters = 4096;

template <bool UNIFORM>
__global__ void __launch_bounds__(512) stream_kernel(const __grid_constant__ Table T, const float2 *wreg_src, float2 *out, long long *cycles) {
    float2 acc[8];
    float2 wr[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        acc[j] = make_float2(0.f, 0.f);
        wr[j] = wreg_src[j];
    }
    float2 in = make_float2(1.0f + threadIdx.x, 2.0f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int s = 0; s < 8; s++) {  // 8 steps: the weight window slides by one per step, as in PFE_GAUSS_GROUP
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float2 w = UNIFORM ? T.w[(it & 7) + ((s + j) & 7)] : wr[(s + j) & 7];
                acc[j] = __ffma2_rn(in, w, acc[j]);
            }
            in.x += 1.0f;  // a new input pair per step
        }
    }
    const long long t1 = clock64();
    float2 sum = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; j++) { sum.x += acc[j].x; sum.y += acc[j].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    Table T;
    for (int i = 0; i < 64; i++) T.w[i] = make_float2(1.0f / (i + 1), 1.0f / (i + 1));
    float2 *wsrc, *out;
    long long *cyc;
    cudaMalloc(&wsrc, sizeof(T.w));
    cudaMemcpy(wsrc, T.w, sizeof(T.w), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 512 * sizeof(float2));
    cudaMalloc(&cyc, sizeof(long long));
    for (int warps_per_smsp = 1; warps_per_smsp <= 4; warps_per_smsp *= 2) {
        const int threads = warps_per_smsp * 4 * 32;  // one block on one SM: warps spread over the 4 sub-partitions
        for (int uniform = 0; uniform < 2; uniform++) {
            long long c = 0;
            for (int rep = 0; rep < 3; rep++) {
                if (uniform) stream_kernel<true><<<1, threads>>>(T, wsrc, out, cyc);
                else stream_kernel<false><<<1, threads>>>(T, wsrc, out, cyc);
                cudaDeviceSynchronize();
                cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
            }
            const double per = (double)c / ((double)kIters * 64.0 * warps_per_smsp);
            printf("%d warp(s) per sub-partition, weights in %s registers: %.3f cycles per FFMA2\n", warps_per_smsp,
                   uniform ? "uniform" : "ordinary", per);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
