// Micro-benchmark: what paces a stream of packed FMAs shaped like the Gaussian inner loop,
//   acc[j] = fma2(in, w[s - j], acc[j])   (N accumulator pairs per input pair, weights sliding by one per step),
// as a function of where the weight operand comes from and how many warps share an SM sub-partition.
//   V0  weight pair (w, w) in ordinary registers           (gauss_*_kernel<UW = false>)
//   V1  weight pair (w, w) from a __grid_constant__ table   (uniform registers, LDCU.64; UW = true)
//   V2  scalar weight w from the table, broadcast           (make_float2(w, w) of ONE uniform 32-bit value)
//   V3  scalar weight in an ordinary register, broadcast
//   V4  scalar FFMA x2 with the weight in a uniform register (the unpacked equivalent, for reference)
//   V5  V2 + the input pair of every step comes from shared memory (one LDS.128 per step, as in the kernels)
//   V6  V2 + the weight of every step is fetched with a RUNTIME index (one LDCU per step: the sliding window)
//   V7  V5 + V6: the Gaussian inner loop's exact instruction mix
//   V8  V7 with the step's weight read from SHARED memory into an ordinary register (LDS.32 broadcast)
// Build and run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_ffma2 tools/ubench_ffma2.cu && /tmp/ubench_ffma2
// Prints FMA-pipe cycles per FFMA2 per SM sub-partition (2.0 = the pipe's rate for a 64-lane operation).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 8;          // accumulator pairs per input pair
constexpr int kSteps = 32;    // steps per loop iteration (fully unrolled: kSteps * N packed FMAs)
constexpr int kIters = 512;

struct Table {
    float2 w2[64];
    float w1[64];
};

struct BigTable {
    float w1[512];
};

// V5 - V8: per step one new input (LDS.128) and / or one new weight entering a rotating window of N weights
template <int V>
__global__ void __launch_bounds__(1024) loop_kernel(const __grid_constant__ BigTable T, int base, float2 *out, long long *cycles) {
    __shared__ float4 tile[64 * 33];
    __shared__ float wsh[512];
    for (int i = threadIdx.x; i < 64 * 33; i += blockDim.x) tile[i] = make_float4(i, i + 1, i + 2, i + 3);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) wsh[i] = T.w1[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float2 lo[N], hi[N];
    float R[N];
#pragma unroll
    for (int j = 0; j < N; j++) { lo[j] = hi[j] = make_float2(0.f, 0.f); R[j] = T.w1[j]; }
    const float4 fixed = tile[lane];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
        const int wb = base + (it & 7) * 16;  // runtime: the compiler cannot pre-load the weights
#pragma unroll 1
        for (int g = 0; g < 32; g += N) {
#pragma unroll
            for (int s = 0; s < N; s++) {
                if (V == 6 || V == 7) R[(s + N - 1) % N] = T.w1[wb + g + s];
                if (V == 8) R[(s + N - 1) % N] = wsh[wb + g + s];
                const float4 in = (V == 5 || V == 7 || V == 8) ? tile[(g + s) * 33 + lane] : fixed;
#pragma unroll
                for (int j = 0; j < N; j++) {
                    const float w = R[(s - j - 1 + 2 * N) % N];
                    lo[j] = __ffma2_rn(make_float2(in.x, in.y), make_float2(w, w), lo[j]);
                    hi[j] = __ffma2_rn(make_float2(in.z, in.w), make_float2(w, w), hi[j]);
                }
            }
        }
    }
    const long long t1 = clock64();
    float2 sum = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < N; j++) { sum.x += lo[j].x + hi[j].x; sum.y += lo[j].y + hi[j].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run_loop(const char *name, const BigTable &T, float2 *out, long long *cyc) {
    for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
        const int threads = warps_per_smsp * 4 * 32;
        long long c = 0;
        for (int rep = 0; rep < 3; rep++) {
            loop_kernel<V><<<1, threads>>>(T, 0, out, cyc);
            cudaDeviceSynchronize();
            cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        }
        const double per = (double)c / ((double)kIters * 32 * N * 2 * warps_per_smsp);
        printf("{\"variant\": \"%s\", \"warps_per_subpartition\": %d, \"cycles_per_ffma2\": %.3f}\n", name, warps_per_smsp, per);
    }
}

template <int V>
__global__ void __launch_bounds__(1024) stream_kernel(const __grid_constant__ Table T, const float2 *wsrc, float2 *out, long long *cycles) {
    float2 acc[N];
    float2 wr[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        acc[j] = make_float2(0.f, 0.f);
        wr[j] = wsrc[j];
    }
    float2 in[4];
#pragma unroll
    for (int i = 0; i < 4; i++) in[i] = make_float2(1.0f + threadIdx.x + i, 2.0f + i);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int s = 0; s < kSteps; s++) {
            const float2 x = in[s & 3];  // the input changes every step; no arithmetic on it inside the timed loop
#pragma unroll
            for (int j = 0; j < N; j++) {
                const int wi = (s - j) & 31;
                if (V == 0) acc[j] = __ffma2_rn(x, wr[(s - j) & (N - 1)], acc[j]);
                if (V == 1) acc[j] = __ffma2_rn(x, T.w2[wi], acc[j]);
                if (V == 2) acc[j] = __ffma2_rn(x, make_float2(T.w1[wi], T.w1[wi]), acc[j]);
                if (V == 3) acc[j] = __ffma2_rn(x, make_float2(wr[(s - j) & (N - 1)].x, wr[(s - j) & (N - 1)].x), acc[j]);
                if (V == 4) { acc[j].x = __fmaf_rn(x.x, T.w1[wi], acc[j].x); acc[j].y = __fmaf_rn(x.y, T.w1[wi], acc[j].y); }
            }
        }
    }
    const long long t1 = clock64();
    float2 sum = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < N; j++) { sum.x += acc[j].x; sum.y += acc[j].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int V>
void run(const char *name, const Table &T, const float2 *wsrc, float2 *out, long long *cyc) {
    for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
        const int threads = warps_per_smsp * 4 * 32;  // one block on one SM: warps spread over the 4 sub-partitions
        long long c = 0;
        for (int rep = 0; rep < 3; rep++) {
            stream_kernel<V><<<1, threads>>>(T, wsrc, out, cyc);
            cudaDeviceSynchronize();
            cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        }
        const double per = (double)c / ((double)kIters * kSteps * N * warps_per_smsp);
        printf("{\"variant\": \"%s\", \"warps_per_subpartition\": %d, \"cycles_per_ffma2\": %.3f}\n", name, warps_per_smsp, per);
    }
}

int main() {
    Table T;
    for (int i = 0; i < 64; i++) { T.w2[i] = make_float2(1.0f / (i + 1), 1.0f / (i + 1)); T.w1[i] = 1.0f / (i + 1); }
    float2 *wsrc, *out;
    long long *cyc;
    cudaMalloc(&wsrc, sizeof(T.w2));
    cudaMemcpy(wsrc, T.w2, sizeof(T.w2), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 1024 * sizeof(float2));
    cudaMalloc(&cyc, sizeof(long long));
    run<0>("V0 weight pair in registers", T, wsrc, out, cyc);
    run<1>("V1 weight pair in uniform registers (LDCU.64)", T, wsrc, out, cyc);
    run<2>("V2 scalar weight in a uniform register, broadcast", T, wsrc, out, cyc);
    run<3>("V3 scalar weight in a register, broadcast", T, wsrc, out, cyc);
    run<4>("V4 two scalar FFMA, weight in a uniform register (per pair)", T, wsrc, out, cyc);
    BigTable B;
    for (int i = 0; i < 512; i++) B.w1[i] = 1.0f / (i + 1);
    run_loop<5>("V5 broadcast uniform weight (fixed window) + LDS.128 input per step", B, out, cyc);
    run_loop<6>("V6 weight window slides: one LDCU (runtime index) per step, fixed input", B, out, cyc);
    run_loop<7>("V7 LDS.128 input + LDCU weight per step (the Gaussian inner loop)", B, out, cyc);
    run_loop<8>("V8 LDS.128 input + LDS.32 weight per step (weight in an ordinary register)", B, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
