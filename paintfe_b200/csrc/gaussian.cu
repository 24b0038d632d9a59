// Separable Gaussian blur (and the unsharp-mask epilogue that rides on its V pass).
//
// Replaces build_gaussian_kernel / parallel_gaussian_blur / blur_with_selection
// (src/ops/filters.rs:141-316), sharpen_core (src/ops/effects/stylize.rs:96-141) and
// GpuRenderer::blur_rgba (src/gpu/renderer.rs:915).  Same structure as the reference: H pass
// u8 -> f32 intermediate, V pass f32 -> u8 with round-half-away + clamp, straight alpha,
// clamp-to-edge, taps accumulated in ascending index order.
//
// Both passes are FP32-pipe bound (16*(2r+1) FLOP per pixel against 8 compulsory bytes), so the
// design goal is "every issue slot is an FFMA":
//   * register blocking along the filter axis: a thread owns N consecutive outputs (4N
//     accumulators) and streams N+2r inputs past them, so one input load feeds 4N FMAs;
//   * the N live weights sit in a rotating register window (one uniform smem read per step);
//   * H pass: a warp owns one row segment of 32*N pixels; the u8 pixels are converted to f32 once
//     while being staged into a skewed shared-memory tile (conflict-free LDS.128), and results go
//     back through the same tile so global stores are fully coalesced;
//   * V pass: lanes run along x (512 B coalesced f32x4 rows); re-reads of neighbouring row blocks
//     are served by L2.
// EXACT=true keeps the reference's separate multiply and add (bit-exact, 2x the FP32 work);
// EXACT=false uses FMA (one rounding per tap instead of two; results within +-1 level).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

struct GaussParams {
    const uint8_t *src;   // H: region origin inside the source image
    float *mid;           // rw*rh*4 f32 intermediate
    uint8_t *dst;         // V: region origin inside the destination image
    const float *wp;      // padded weights (scalars), device
    const uint8_t *orig;  // sharpen: original pixels (same geometry as dst), else null
    const uint8_t *mask;  // sharpen: selection mask plane (w*h) at region origin, or null
    float amount;         // sharpen amount / glow intensity
    int epilogue;         // 0 store the blur, 1 unsharp mask (sharpen_core), 2 screen glow (glow_core)
    uint32_t src_pitch;   // pixels per source row
    uint32_t dst_pitch;   // pixels per destination row
    uint32_t mask_pitch;
    uint32_t rw, rh;      // region size
    uint32_t v_y0, v_rows; // V pass: produce output rows [v_y0, v_y0 + v_rows) only (band pipelining)
    int radius;
    float sigma;          // host side only: key of the device-resident weight table
    int steps;            // T: padded step count, multiple of N
    int wp_len;           // steps + N - 1
    const uint32_t *bb;       // selection blur: device {min_x, min_y, max_x, max_y} of the mask, or null (see bb_skip)
    int dbg;                  // diagnosis only (PFE_GAUSS_DBG): 1 = H pass skips its staging loads, 2 = skips its stores
    int tri;                  // steps == N + taps - 1: triangular first / last groups (see PFE_GAUSS_GROUP)
    int vchunk;               // V tile kernel: N-row groups per ring chunk (launch_v)
    uint32_t *queue;          // persistent passes: {next task, finished grabbers} of this launch (pfe_queue_slot)
    int seg_rows, nseg, lag;  // fused kernel: rows per strip segment, segments per strip, V lag in batches
    // 1 and -0 for the EXACT path's packed arithmetic. They travel as parameters so that no compiler
    // stage can see their values (see tap<true> below and blend.cuh).
    float one, nzero;
};

// One RGBA accumulator as two packed f32x2 halves: sm_100's FFMA2 / FMUL2 / FADD2 retire two IEEE
// f32 operations per issue slot, which leaves the other slot free for the loads and keeps the FMA
// pipe (not the issue port) the limiter.  EXACT mode (separate multiply and add, as the reference
// does) must stay scalar: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even
// under --fmad=false, whereas scalar __fmul_rn/__fadd_rn are never fused.
struct Acc4 {
    float2 lo, hi;  // (r,g) (b,a)
};
// EXACT: the reference's separately rounded multiply and add, two lanes per instruction.  a*b == fma(a, b, -0) and
// t + c == fma(t, 1, c) exactly, and with the constants hidden in kernel parameters neither NVVM nor ptxas can turn
// the pair back into a contractable mul + add (ptxas 12.9 fuses mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even
// under --fmad=false).  Same FMA-pipe time as four scalar FMUL + four FADD, half the issue slots and half the code.
template <bool EXACT>
__device__ __forceinline__ void tap(Acc4 &acc, const float4 &in, const float w, const float one1, const float nzero1) {
    // the weight is ONE 32-bit operand broadcast to both lanes (SASS `FFMA2 R, R, UR.F32, R`): measured with
    // tools/ubench_ffma2.cu, that form issues at the FMA pipe's rate (2.02 cycles per FFMA2 per sub-partition),
    // a 64-bit (w, w) pair operand at 2.34 from uniform registers and 2.30 from ordinary ones
    const float2 ilo = make_float2(in.x, in.y), ihi = make_float2(in.z, in.w), w2 = make_float2(w, w);
    const float2 one = make_float2(one1, one1), nzero = make_float2(nzero1, nzero1);
    if (EXACT) {
        acc.lo = __ffma2_rn(__ffma2_rn(ilo, w2, nzero), one, acc.lo);
        acc.hi = __ffma2_rn(__ffma2_rn(ihi, w2, nzero), one, acc.hi);
    } else {
        acc.lo = __ffma2_rn(ilo, w2, acc.lo);
        acc.hi = __ffma2_rn(ihi, w2, acc.hi);
    }
}

// wp[m] = w[m-(N-1)] inside the kernel support, 0 outside: every output uses the same unrolled
// N-step body; zero-weight taps are exact no-ops in both modes (x*0 + acc == acc).  The N live
// weights sit in a rotating register window R[] (one uniform 8-byte read per step).
// LOAD(s) yields the input for step g+s.  Output j meets weight w[g+s-j] at step s of group g.
#define PFE_GAUSS_GROUP_C(LOAD, COND)                                                           \
    _Pragma("unroll") for (int s = 0; s < N; s++) {                                             \
        R[(s + N - 1) % N] = wsm[g + s + N - 1];                                                \
        const float4 in = LOAD(s);                                                              \
        _Pragma("unroll") for (int j = 0; j < N; j++)                                           \
            if (COND) tap<EXACT>(acc[j], in, R[(s - j - 1 + 2 * N) % N], P.one, P.nzero);       \
    }
// When the padded step count is exactly N + taps - 1 (P.tri: 2*radius divisible by N - sigma 20 with N = 8 is), the
// zero-weight taps are the upper triangle of the first group (w[s-j] with j > s) and the lower triangle of the last
// one (j < s): those groups run triangular bodies that skip them - N-1 of the N+taps-1 FMAs per output (5.5 % at
// sigma 20, 10 % in the fused small-radius kernel).  Skipping an exact no-op changes no result.
#define PFE_GAUSS_GROUP(LOAD)                                                                   \
    if (P.tri && g == 0) { PFE_GAUSS_GROUP_C(LOAD, j <= s) }                                    \
    else if (P.tri && g == P.steps - N) { PFE_GAUSS_GROUP_C(LOAD, j >= s) }                     \
    else { PFE_GAUSS_GROUP_C(LOAD, true) }

// The same with the FIRST input of each group carried across the loop in `pre_in`: the group loads the next group's
// first input (LOAD_NEXT) before its own FMAs, so the address arithmetic and shared-memory latency of that load hide
// under 128 FFMA2 instead of stalling the first FFMA2 of every group (10 % of all stall samples in the r02 ncu source
// view of the H pass).  NEXT_OK (warp-uniform) says whether a next group exists whose input may be read already.
#define PFE_GAUSS_GROUP_PC(LOAD, LOAD_NEXT, NEXT_OK, COND)                                      \
    {                                                                                           \
        const float4 first_in = pre_in;                                                         \
        if (NEXT_OK) pre_in = LOAD_NEXT;                                                        \
        _Pragma("unroll") for (int s = 0; s < N; s++) {                                         \
            R[(s + N - 1) % N] = wsm[g + s + N - 1];                                            \
            const float4 in = s == 0 ? first_in : LOAD(s);                                      \
            _Pragma("unroll") for (int j = 0; j < N; j++)                                       \
                if (COND) tap<EXACT>(acc[j], in, R[(s - j - 1 + 2 * N) % N], P.one, P.nzero);   \
        }                                                                                       \
    }
#define PFE_GAUSS_GROUP_P(LOAD, LOAD_NEXT, NEXT_OK)                                             \
    if (P.tri && g == 0) PFE_GAUSS_GROUP_PC(LOAD, LOAD_NEXT, NEXT_OK, j <= s)                   \
    else if (P.tri && g == P.steps - N) PFE_GAUSS_GROUP_PC(LOAD, LOAD_NEXT, NEXT_OK, j >= s)    \
    else PFE_GAUSS_GROUP_PC(LOAD, LOAD_NEXT, NEXT_OK, true)

__host__ __device__ __forceinline__ int skew(int p, int n) { return p + p / n; }

// blur_with_selection (filters.rs:141-207) crops to the mask's bounding box grown by ceil(3 sigma), blurs the crop and
// copies back where the mask is set.  A selected pixel is at least the radius away from every crop edge that is not an
// image edge, so its taps (and the taps of the H values its V pass reads) never meet the crop's clamp: blurring in the
// full image's geometry gives it the very same value.  The kernels therefore keep the whole-image geometry and only
// SKIP work that no selected pixel can see - decided on the device from the bounding box the mask_bbox kernel left
// there, so nothing is read back and the call stays asynchronous.  True when the pixel rectangle [x0, x1] x [y0, y1]
// (inclusive) lies outside the box grown by (gx, gy); an empty selection (min > max) skips everything.
__device__ __forceinline__ bool bb_skip(const uint32_t *bb, int x0, int x1, int y0, int y1, int gx, int gy) {
    const int bx0 = (int)__ldg(bb), by0 = (int)__ldg(bb + 1), bx1 = (int)__ldg(bb + 2), by1 = (int)__ldg(bb + 3);
    if ((uint32_t)bx0 > (uint32_t)bx1 || (uint32_t)by0 > (uint32_t)by1) return true;
    return x1 < bx0 - gx || x0 > bx1 + gx || y1 < by0 - gy || y0 > by1 + gy;
}

// u8 -> f32 uses PRMT + FADD against 2^23 (ALU/FMA pipes) instead of the quarter-rate I2F:
// 0x4B0000xx is 2^23 + xx as a float, so one PRMT per channel and an exact subtract.
__device__ __forceinline__ float4 to_f4(uint32_t v) {
    const float m = 8388608.0f;
    return make_float4(__uint_as_float(__byte_perm(v, 0x4B000000u, 0x7440)) - m, __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7441)) - m,
                       __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7442)) - m, __uint_as_float(__byte_perm(v, 0x4B000000u, 0x7443)) - m);
}

// ---- H pass: u8 -> f32 ----------------------------------------------------------------------
// A warp owns one row segment of 32*N pixels. The u8 pixels are converted to f32 once while being
// staged into a skewed shared-memory tile (one pad float4 every N, so lane stride N+1 keeps
// LDS.128 conflict free); results return through the same tile for fully coalesced stores.
// UW: the padded weight table travels in the kernel parameters.  The per-step weight is then an `LDCU` into a uniform
// register and every tap `FFMA2 acc, in, UR.F32, acc` (one scalar weight broadcast to both lanes - the operand form that
// issues at the pipe's own rate, tools/ubench_ffma2.cu): no weight LDS, no per-thread weight registers.  Same table, same
// arithmetic, same results; measured 1.47 -> 1.25 ms at sigma 20, 8K (profiles/r02_gauss_steps.txt).
constexpr int kWeightTableLen = 800;  // floats: 3.2 KB of the 4 KB parameter space; wp_len <= 800 covers sigma <= ~128
struct WeightTable {
    float wk[kWeightTableLen];
};

// BB: the selection-blur variant that skips work outside the mask's bounding box (see bb_skip).  A separate
// instantiation, because the test inside the task loop changed ptxas' allocation of the ordinary kernel from 71 to
// 96 registers (5 instead of 7 resident CTAs, 0.57 -> 0.69 ms at 8K); kernels without a BB instantiation simply
// blur everything, which gives the same selected pixels.
template <int N, bool EXACT, int WARPS, bool UW, bool BB = false>
__global__ void __launch_bounds__(WARPS * 32) gauss_h_kernel(const __grid_constant__ GaussParams P, const __grid_constant__ WeightTable W) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *wshared = reinterpret_cast<float *>(smem_raw);
    const float *wsm = UW ? W.wk : wshared;
    const int wp_pad = UW ? 0 : (P.wp_len + 3) & ~3;
    const int tile_px = 31 * N + P.steps;
    const int tile_len = skew(tile_px, N) + 1;
    float4 *tiles = reinterpret_cast<float4 *>(wshared + wp_pad);
    if (!UW) {
        for (int i = threadIdx.x; i < P.wp_len; i += blockDim.x) wshared[i] = P.wp[i];
        __syncthreads();
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 *tile = tiles + (size_t)warp * tile_len;
    constexpr int SEG = 32 * N;
    const uint32_t nseg = (P.rw + SEG - 1) / SEG;
    const uint64_t ntask = (uint64_t)nseg * P.rh;
    const int rw = (int)P.rw;

    // Tasks come from a queue, not from a fixed stride: the warp scheduler favours the older warps of an SM
    // sub-partition, so with equal shares they finished one after the other and the partition ran its last seventh on a
    // single warp (r02 ncu: 4.6 of 7 warps resident on average, FMA pipe 80 % active).  Every warp now holds one task
    // in hand (`next`) and takes another whenever it starts one, so all seven stay busy until the queue is empty.
    // The NEXT task's source pixels travel in registers while this one is filtered (PRE words per lane), so the global
    // load latency of the staging step hides under ~4000 cycles of FMAs instead of stalling the warp three dependent
    // round trips per task (measured at 8K sigma 20: 0.60 ms with the loads in line, 0.53 ms without any loads at all).
    // Tiles wider than 32 * PRE pixels (large sigma) keep the in-line staging.
    constexpr int PRE = 12;
    const bool use_pre = !BB && tile_px <= 32 * PRE && !(P.dbg & (1 | 4));  // dbg 4: in-line staging (A/B)
    uint32_t pre[PRE];
    auto fetch = [&](uint64_t t) {
        const uint32_t fy = (uint32_t)(t / nseg);
        const int fx0 = (int)(t % nseg) * SEG - P.radius;
        const uint32_t *frow = reinterpret_cast<const uint32_t *>(P.src) + (size_t)fy * P.src_pitch;
#pragma unroll
        for (int i = 0; i < PRE; i++) {
            const int p = lane + 32 * i;
            pre[i] = p < tile_px ? __ldg(frow + min(max(fx0 + p, 0), rw - 1)) : 0u;
        }
    };
    auto grab = [&]() -> uint64_t {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(P.queue, 1u);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    uint64_t task = grab();
    if (use_pre && task < ntask) fetch(task);

    while (task < ntask) {
        const uint64_t next = grab();
        const uint64_t cur = task;
        task = next;
        const uint32_t y = (uint32_t)(cur / nseg);
        const int x0 = (int)(cur % nseg) * SEG;
        const uint32_t *row = reinterpret_cast<const uint32_t *>(P.src) + (size_t)y * P.src_pitch;
        // selection blur: H values are read by the V pass of selected pixels in the same columns, up to radius rows away
        if (BB && bb_skip(P.bb, x0, x0 + SEG - 1, (int)y, (int)y, 0, P.radius)) continue;
        // stage + convert: tile[p] = pixel clamp(x0 - r + p)
        if (use_pre) {
#pragma unroll
            for (int i = 0; i < PRE; i++) {
                const int p = lane + 32 * i;
                if (p < tile_px) tile[skew(p, N)] = to_f4(pre[i]);
            }
            if (next < ntask) fetch(next);
        } else if (!(P.dbg & 1)) {
#pragma unroll 4
            for (int p = lane; p < tile_px; p += 32) tile[skew(p, N)] = to_f4(__ldg(row + min(max(x0 - P.radius + p, 0), rw - 1)));
        }
        __syncwarp();
        Acc4 acc[N];
        float R[N];
#pragma unroll
        for (int j = 0; j < N; j++) acc[j].lo = acc[j].hi = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < N - 1; m++) R[m] = wsm[m];
        // input i of this lane sits at skew(lane*N + i) = lane*(N+1) + i + i/N
        const float4 *tl = tile + lane * (N + 1);
        float4 pre_in = tl[0];
        for (int g = 0; g < P.steps; g += N) {
            const float4 *tg = tl + g + g / N;  // within a group i/N is constant: step s is tg[s]
#define H_LOAD(s) tg[s]
            // the next group's first input sits one skew slot further: tl[(g + N) + (g + N) / N] = tg[N + 1]
            PFE_GAUSS_GROUP_P(H_LOAD, tg[N + 1], g + N < P.steps)
#undef H_LOAD
        }
        __syncwarp();
        // results through the tile -> coalesced 512 B stores
        float4 *to = tile + lane * (N + 1);
#pragma unroll
        for (int j = 0; j < N; j++) to[j] = make_float4(acc[j].lo.x, acc[j].lo.y, acc[j].hi.x, acc[j].hi.y);
        __syncwarp();
        float4 *out = reinterpret_cast<float4 *>(P.mid) + (size_t)y * P.rw;
#pragma unroll 4
        for (int p = lane; p < SEG; p += 32)
            if (x0 + p < rw && !(P.dbg & 2)) out[x0 + p] = tile[skew(p, N)];
        __syncwarp();
    }
    // every warp makes exactly one grab past the end; the last one to have done so leaves the queue zeroed for the
    // launch that gets this slot next
    if (lane == 0 && atomicAdd(P.queue + 1, 1u) == gridDim.x * WARPS - 1) {
        P.queue[0] = 0u;
        P.queue[1] = 0u;
    }
}

// ---- V pass: f32 -> u8 ----------------------------------------------------------------------
// Round, clamp and store one output pixel; optionally the unsharp-mask epilogue.
__device__ __forceinline__ void v_store(const GaussParams &P, const Acc4 &acc, int x, int y) {
    uint32_t r8 = pfe_round_u8_nonneg(acc.lo.x), g8 = pfe_round_u8_nonneg(acc.lo.y), b8 = pfe_round_u8_nonneg(acc.hi.x), a8 = pfe_round_u8_nonneg(acc.hi.y);
    uint32_t outv;
    if (P.orig) {  // sharpen_core, stylize.rs:116-134
        uint32_t s = reinterpret_cast<const uint32_t *>(P.orig)[(size_t)y * P.dst_pitch + x];
        if (P.mask && P.mask[(size_t)y * P.mask_pitch + x] == 0) {
            outv = s;
        } else if (P.epilogue == 2) {  // glow_core, stylize.rs:62-70: screen of the source with blur*intensity
            uint32_t o[3];
            const uint32_t bl[3] = {r8, g8, b8};
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float sv = (float)((s >> (8 * c)) & 255u) / 255.0f, bv = (float)bl[c] / 255.0f;
                o[c] = pfe_round_u8((1.0f - (1.0f - sv) * (1.0f - bv * P.amount)) * 255.0f);
            }
            outv = pfe_pack(o[0], o[1], o[2], s >> 24);
        } else {
            const float4 sf = to_f4(s);
            outv = pfe_pack(pfe_round_u8(sf.x + P.amount * (sf.x - pfe_u8_to_f32(r8))), pfe_round_u8(sf.y + P.amount * (sf.y - pfe_u8_to_f32(g8))),
                            pfe_round_u8(sf.z + P.amount * (sf.z - pfe_u8_to_f32(b8))), s >> 24);
        }
    } else {
        outv = pfe_pack(r8, g8, b8, a8);
    }
    reinterpret_cast<uint32_t *>(P.dst)[(size_t)y * P.dst_pitch + x] = outv;
}

// Direct variant: inputs straight from global memory (L1/L2 serve the (N+2r)/N-fold re-reads).
// Used when the tile variant's shared-memory footprint does not fit (very large sigma).
template <int N, bool EXACT>
__global__ void __launch_bounds__(128) gauss_v_kernel(const __grid_constant__ GaussParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *wsm = reinterpret_cast<float *>(smem_raw);
    for (int i = threadIdx.x; i < P.wp_len; i += blockDim.x) wsm[i] = P.wp[i];
    __syncthreads();
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= (int)P.rw) return;
    const int rh = (int)P.rh;
    const float4 *mid = reinterpret_cast<const float4 *>(P.mid) + x;
    const int y_end = (int)(P.v_y0 + P.v_rows);
    for (int y0 = (int)P.v_y0 + blockIdx.y * N; y0 < y_end; y0 += gridDim.y * N) {
        if (P.bb && bb_skip(P.bb, x, x, y0, y0 + N - 1, 0, 0)) continue;
        Acc4 acc[N];
        float R[N];
#pragma unroll
        for (int j = 0; j < N; j++) acc[j].lo = acc[j].hi = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < N - 1; m++) R[m] = wsm[m];
        const int ybase = y0 - P.radius;
        for (int g = 0; g < P.steps; g += N) {
#define V_LOAD(s) __ldg(mid + (size_t)min(max(ybase + g + (s), 0), rh - 1) * P.rw)
            PFE_GAUSS_GROUP(V_LOAD)
#undef V_LOAD
        }
#pragma unroll
        for (int j = 0; j < N; j++)
            if (y0 + j < y_end) v_store(P, acc[j], x, y0 + j);
    }
}

// Tile variant: warp-specialised producer/consumer pipeline.
// A CTA owns 32-pixel-wide tiles of WARPS*N output rows. The tile's input rows (tile height + 2r)
// live in shared memory as a ring of kChunks chunks, each guarded by a full/empty mbarrier pair:
//   * one producer warp streams rows in with 512-byte cp.async.bulk copies (clamp-to-edge is a clamped
//     source row index; completion is counted in bytes on the chunk's `full` barrier) and refills a
//     chunk for the NEXT tile as soon as every consumer warp has released it, so loads run a whole
//     tile ahead of the math and there is no end-of-tile bubble;
//   * WARPS consumer warps each stream their N+2r rows out of the ring with conflict-free LDS.128,
//     waiting only on the chunk they are about to read and releasing chunks behind them.
// A row of the f32 intermediate crosses L2 ~(TH+2r)/TH times instead of (N+2r)/N times.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

constexpr int kChunks = 32;     // barrier pairs laid out; a tile uses ceil(groups / chunk_groups) of them
constexpr int kDefaultChunks = 3;  // measured at sigma 20, 8K (12 warps, 27 groups per tile): 9 groups per chunk 0.572 ms, 4: 0.578, 2: 0.599, 14: 0.597

// Blur outputs are non-negative (weights and inputs are >= 0): the clamp-free rounding applies directly.
__device__ __forceinline__ uint32_t round_u8_nonneg(float x) { return pfe_round_u8_nonneg(x); }

template <int N, bool EXACT, int WARPS, bool UW, bool BB = false>
__global__ void __launch_bounds__((WARPS + 1) * 32) gauss_v_tile_kernel(const __grid_constant__ GaussParams P, const __grid_constant__ WeightTable W) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TH = WARPS * N;
    const int rows = TH + P.steps - N;  // a multiple of N
    const int ring_rows = rows;  // region size the host laid out
    // Chunks hold a whole number of N-row groups: a consumer's group then never straddles two chunks, so it waits once
    // per chunk, runs that chunk's groups in a tight loop and hands the chunk back - no per-group bookkeeping.
    const int chunk_groups = P.vchunk;
    const int chunk_rows = chunk_groups * N;
    const int nchunks = (rows + chunk_rows - 1) / chunk_rows;  // <= kChunks
    float4 *tile = reinterpret_cast<float4 *>(smem_raw);                       // ring_rows x 32 float4
    float *wshared = reinterpret_cast<float *>(smem_raw + (size_t)ring_rows * 512);
    const float *wsm = UW ? W.wk : wshared;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)ring_rows * 512 + (size_t)((P.wp_len + 1) & ~1) * 8);
    if (!UW)
        for (int i = threadIdx.x; i < P.wp_len; i += blockDim.x) wshared[i] = P.wp[i];
    const uint32_t full0 = smem_addr(bars), empty0 = smem_addr(bars + kChunks);
    if (threadIdx.x == 0) {
        for (int c = 0; c < nchunks; c++) {
            // `empty` counts only the warps that READ the chunk (warp w reads ring rows [w*N, w*N + steps)):
            // a reader arrives for tile t after it has seen full[c] of tile t, which the producer only
            // signals after empty[c] of tile t-1 completed - so no warp, however far it runs ahead, can
            // contribute two arrivals to one phase. Warps that never touch a chunk take no part in it.
            int readers = 0;
            for (int w = 0; w < WARPS; w++)
                if (w * N < (c + 1) * chunk_rows && w * N + P.steps > c * chunk_rows) readers++;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8u * c), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8u * c), "r"(max(readers, 1)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();

    // the shuffle tells the compiler that `warp` is warp-uniform: the role branch below is then uniform control flow and
    // the consumer loop may use the uniform datapath (weights in uniform registers)
    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int rw = (int)P.rw, rh = (int)P.rh;
    const int y_end = (int)(P.v_y0 + P.v_rows);
    const int tx = (rw + 31) / 32, ty = ((int)P.v_rows + TH - 1) / TH;
    const int ntiles = tx * ty;

    // Tiles come from a queue (P.queue), not from a fixed stride: with equal shares the CTAs - and the warps the
    // scheduler favours inside them - ran out of work one after the other (r02 ncu: 20 of 26 warps resident on average).
    // The producer warp takes the next tile and publishes its index in tile_id[lap parity] before it arms the tile's
    // first chunk; a consumer reads it once the first chunk it waits for has landed.  Slot reuse is safe for the reason
    // the ring is: the producer reaches tile L only after every chunk of tile L-2 was handed back, so every consumer has
    // long read tile L-2's index.  When the queue is empty the producer sends a null tile (-1) through the same
    // barriers, so that consumers parked on `full` wake up and leave.
    volatile int *tile_id = reinterpret_cast<volatile int *>(bars + 2 * kChunks);
    if (warp == WARPS) {
        // ===== producer warp =====
        for (uint32_t it = 0;; it++) {
            int t;
            for (;;) {
                uint32_t q = 0;
                if (lane == 0) q = atomicAdd(P.queue, 1u);
                q = __shfl_sync(0xffffffffu, q, 0);
                t = q < (uint32_t)ntiles ? (int)q : -1;
                if (!BB || t < 0) break;
                // column-major tile order: consecutive tiles walk down a strip, so halo rows are L2-hot
                const int bx0 = (t / ty) * 32, by0 = (int)P.v_y0 + (t % ty) * TH;
                if (!bb_skip(P.bb, bx0, bx0 + 31, by0, by0 + TH - 1, 0, 0)) break;
            }
            const int x0 = (max(t, 0) / ty) * 32, y0 = (int)P.v_y0 + (max(t, 0) % ty) * TH;
            const uint32_t row_bytes = (uint32_t)min(32, rw - x0) * 16u;
            for (int c = 0; c < nchunks; c++) {
                const int r0 = c * chunk_rows, nrows = min(chunk_rows, rows - r0);
                if (it > 0) mbar_wait(empty0 + 8u * c, (it - 1) & 1u);  // its readers are done with the previous tile's chunk c
                if (lane == 0) {
                    if (c == 0) tile_id[it & 1u] = t;
                    if (t >= 0)
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full0 + 8u * c), "r"(row_bytes * (uint32_t)nrows) : "memory");
                    else
                        mbar_arrive(full0 + 8u * c);
                }
                __syncwarp();
                if (t < 0) continue;
                for (int rr = r0 + lane; rr < r0 + nrows; rr += 32) {
                    const int sy = min(max(y0 - P.radius + rr, 0), rh - 1);
                    const float4 *src = reinterpret_cast<const float4 *>(P.mid) + (size_t)sy * rw + x0;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                     smem_addr(tile + (size_t)rr * 32)),
                                 "l"(src), "r"(row_bytes), "r"(full0 + 8u * c)
                                 : "memory");
                }
            }
            if (t < 0) break;
        }
        // one grab past the end per CTA; the last CTA to make it leaves the queue zeroed for the slot's next launch
        if (lane == 0 && atomicAdd(P.queue + 1, 1u) == gridDim.x - 1) {
            P.queue[0] = 0u;
            P.queue[1] = 0u;
        }
        return;
    }

    // ===== consumer warps =====
    const int row_first = warp * N;  // first ring row this warp reads
    for (uint32_t it = 0;; it++) {
        const uint32_t parity = it & 1u;
        mbar_wait(full0 + 8u * (uint32_t)(warp / chunk_groups), parity);  // the first chunk this warp reads
        const int t = tile_id[parity];
        if (t < 0) break;
        const int x0 = (t / ty) * 32, y0 = (int)P.v_y0 + (t % ty) * TH;
        Acc4 acc[N];
        float R[N];
#pragma unroll
        for (int j = 0; j < N; j++) acc[j].lo = acc[j].hi = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < N - 1; m++) R[m] = wsm[m];
        const float4 *col = tile + (size_t)row_first * 32 + lane;
        // this warp reads ring groups [warp, warp + steps / N): it waits for, and hands back, exactly the chunks those
        // groups lie in (a warp may only watch barriers it also gates, or it could fall two phases behind one and
        // mis-read its parity)
        int q = warp, g = 0;  // ring group about to be read; step offset of that group in this warp's filter
        const int q_end = warp + P.steps / N;
        int c = q / chunk_groups;
        int q_stop = min(q_end, (c + 1) * chunk_groups);  // first group beyond chunk c
        float4 pre_in = col[0];  // the first chunk has landed (waited for above): its first group's first row
        for (; q < q_end; q++, g += N) {
            const bool chunk_ends = q + 1 == q_stop, more = q + 1 < q_end;
            // The next chunk is waited for BEFORE this chunk's last group, so that group can read the next group's
            // first row ahead like every other one: the FMA stream does not break at a chunk boundary.
            if (chunk_ends && more) mbar_wait(full0 + 8u * (uint32_t)(c + 1), parity);
            const float4 *cg = col + (size_t)g * 32;
#define VT_LOAD(s) cg[(s) * 32]
            PFE_GAUSS_GROUP_P(VT_LOAD, cg[N * 32], more)
#undef VT_LOAD
            if (chunk_ends) {
                __syncwarp();  // every lane has read the chunk's rows
                if (lane == 0) mbar_arrive(empty0 + 8u * (uint32_t)c);
                c++;
                q_stop = min(q_end, (c + 1) * chunk_groups);
            }
        }
        const int x = x0 + lane, yw = y0 + row_first;
        if (x < rw) {
#pragma unroll
            for (int j = 0; j < N; j++) {
                if (yw + j >= y_end) break;
                if (P.orig) {
                    v_store(P, acc[j], x, yw + j);
                } else {
                    reinterpret_cast<uint32_t *>(P.dst)[(size_t)(yw + j) * P.dst_pitch + x] =
                        pfe_pack(round_u8_nonneg(acc[j].lo.x), round_u8_nonneg(acc[j].lo.y), round_u8_nonneg(acc[j].hi.x),
                                 round_u8_nonneg(acc[j].hi.y));
                }
            }
        }
    }
}

// ---- Fused H+V for small radii ------------------------------------------------------------------
// For small radii (use_fused below) the f32 intermediate never leaves the SM.  A CTA walks down a 128-pixel-wide
// strip segment in batches of 8 rows, its warps split into two roles that meet in a shared-memory ring of
// lag+2 batches of the intermediate (lag = ceil(2r/8)), each batch guarded by a full/empty mbarrier pair:
//   * 4 H warps: warp w turns source rows w and w+4 of the batch into intermediate rows exactly as gauss_h_kernel does
//     (u8 -> f32 once into a skewed per-warp tile, a lane owns 4 consecutive outputs; the next row's pixels are
//     already in registers while this one is filtered), waits for the ring slot to be free, and arrives on `full`;
//   * 4 V warps: 128 lanes = 128 columns.  Once batch b is full, outputs of batch b-lag have all their 8+2r rows in
//     the ring: a lane streams them past 4 outputs at a time (conflict-free LDS.128, lanes run along x), stores u8
//     - with the sharpen / glow epilogue if set - and arrives on `empty` of the batch it no longer needs.
// Both roles carry the same FMA count per batch, one warp of each role lands on each SM sub-partition, and no
// CTA-wide barrier is left: the FMA pipe stays fed while either role waits on memory.
// Protocol safety (same argument as the V ring above): `full[s]` of lap L+1 needs `empty[s]` of lap L, which needs
// every V warp to have waited `full[s]` of lap L first, and vice versa - no waiter can fall two phases behind.
// Every input row is filtered horizontally once per segment (plus 2r warm-up rows at the head of a segment) and
// the image crosses HBM once in, once out: 8 B per pixel instead of the two-pass 40 B.  Tap order and arithmetic
// are those of the two-pass kernels, so the results are identical to theirs in both modes.
constexpr int kFusedMaxRadius = 16;
constexpr int kFusedTW = 128;
constexpr int kFusedMaxRing = 8;  // batches; lag + 2 <= 6 for radius <= 16
constexpr int kFusedMaxWp = 40;   // padded weights for N = 4: steps + 3 <= 36 + 3

// The padded weight table travels in the kernel parameters: the per-step weight is a uniform constant-bank read
// (no shared-memory wavefront, no staging), which matters because at N = 4 this kernel is bound by the
// shared-memory pipe, not by the FMA pipe.
struct FusedWeights {
    float wk[kFusedMaxWp];
};

// Ring column swizzle: pixel x of a ring row lives in 16-byte slot x ^ ((x >> 3) & 7).  The H warps write 4
// consecutive pixels per lane (a quarter-warp would hit only 2 of the 8 bank groups: 4-way conflicts), the V warps
// read consecutive pixels across lanes; with the swizzle both are conflict free.
__device__ __forceinline__ int fused_swz(int x) { return x ^ ((x >> 3) & 7); }

template <bool EXACT>
__global__ void __launch_bounds__(256, 2) gauss_fused_kernel(const __grid_constant__ GaussParams P, const __grid_constant__ FusedWeights W) {
    constexpr int N = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const float *wsm = W.wk;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);  // full[kFusedMaxRing], empty[kFusedMaxRing]
    const int tile_px = 31 * N + P.steps;
    const int tile_len = skew(tile_px, N) + 1;
    float4 *tiles = reinterpret_cast<float4 *>(bars + 2 * kFusedMaxRing);
    float4 *ring = tiles + (size_t)4 * tile_len;  // RB batches x 8 rows x 128 float4
    const int RB = P.lag + 2;
    const int ring_rows = RB * 8;
    const uint32_t full0 = smem_addr(bars), empty0 = smem_addr(bars + kFusedMaxRing);
    if (threadIdx.x == 0) {
        for (int c = 0; c < RB; c++) {
            // every lane of a role arrives for itself (release / acquire pairs lane to lane, nothing rides on a
            // warp-level barrier plus one elected arrival - which compute-sanitizer's racecheck does not follow)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8u * c), "r"(128));   // 4 H warps
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(empty0 + 8u * c), "r"(128));  // 4 V warps
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform role
    const int rw = (int)P.rw, rh = (int)P.rh, r = P.radius;
    const int nstrips = (rw + kFusedTW - 1) / kFusedTW;
    const int ntask = nstrips * P.nseg;

    if (warp < 4) {
        // ===== H warps =====
        float4 *tile = tiles + (size_t)warp * tile_len;
        int slot = 0;        // ring slot of the batch being produced
        uint32_t lap = 0;    // how many times the ring has wrapped
        for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
            // consecutive CTAs take neighbouring strips of one segment: their halo columns are L2-hot
            const int x0 = (task % nstrips) * kFusedTW;
            const int ys = (task / nstrips) * P.seg_rows, ye = min(ys + P.seg_rows, rh);
            const int nb = (ye - ys + 7) / 8 + P.lag;
            // this warp's k-th row is batch k/2, row warp + 4*(k&1); its source pixels travel in registers one row
            // ahead (tile_px <= 31*4 + 36 = 160 = 5 per lane)
            uint32_t pre[5];
            auto fetch = [&](int k) {
                const int yy = min(max(ys - r + 8 * (k >> 1) + warp + 4 * (k & 1), 0), rh - 1);  // clamp-to-edge
                const uint32_t *row = reinterpret_cast<const uint32_t *>(P.src) + (size_t)yy * P.src_pitch;
#pragma unroll
                for (int i = 0; i < 5; i++) {
                    const int p = lane + 32 * i;
                    pre[i] = p < tile_px ? __ldg(row + min(max(x0 - r + p, 0), rw - 1)) : 0u;
                }
            };
            fetch(0);
            for (int b = 0; b < nb; b++) {
                if (lap > 0) mbar_wait(empty0 + 8u * (uint32_t)slot, (lap - 1) & 1u);  // the V warps are done with this slot's previous batch
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
#pragma unroll
                    for (int i = 0; i < 5; i++) {
                        const int p = lane + 32 * i;
                        if (p < tile_px) tile[skew(p, N)] = to_f4(pre[i]);
                    }
                    if (2 * b + half + 1 < 2 * nb) fetch(2 * b + half + 1);
                    __syncwarp();
                    Acc4 acc[N];
                    float R[N];
#pragma unroll
                    for (int j = 0; j < N; j++) acc[j].lo = acc[j].hi = make_float2(0.f, 0.f);
#pragma unroll
                    for (int m = 0; m < N - 1; m++) R[m] = wsm[m];
                    const float4 *tl = tile + lane * (N + 1);
                    for (int g = 0; g < P.steps; g += N) {
                        const float4 *tg = tl + g + g / N;
#define F_HLOAD(s) tg[s]
                        PFE_GAUSS_GROUP(F_HLOAD)
#undef F_HLOAD
                    }
                    float4 *mrow = ring + (size_t)(slot * 8 + warp + 4 * half) * kFusedTW;
#pragma unroll
                    for (int j = 0; j < N; j++) mrow[fused_swz(lane * N + j)] = make_float4(acc[j].lo.x, acc[j].lo.y, acc[j].hi.x, acc[j].hi.y);
                    __syncwarp();  // every lane is done with the tile (and has written its part of the row)
                }
                mbar_arrive(full0 + 8u * (uint32_t)slot);
                if (++slot == RB) { slot = 0; lap++; }
            }
        }
        return;
    }

    // ===== V warps =====
    const int vc = (warp - 4) * 32 + lane;  // column inside the strip
    const int vcs = fused_swz(vc);          // ... and where it lives in a ring row
    int wslot = 0;       // ring slot of the next batch to wait for
    uint32_t wlap = 0;
    int rslot = 0;       // ring slot of the next batch to filter and release
    for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
        const int x0 = (task % nstrips) * kFusedTW;
        const int ys = (task / nstrips) * P.seg_rows, ye = min(ys + P.seg_rows, rh);
        const int nb = (ye - ys + 7) / 8 + P.lag;
        const int x = x0 + vc;
        for (int b = 0; b < nb; b++) {
            mbar_wait(full0 + 8u * (uint32_t)wslot, wlap & 1u);
            if (++wslot == RB) { wslot = 0; wlap++; }
            if (b < P.lag) continue;
            // output rows ys + 8(b-lag) + [0,8): their first ring row is row 0 of batch b-lag (= slot rslot)
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                const int yo = ys + 8 * (b - P.lag) + 4 * half;
                int row = rslot * 8 + 4 * half;  // a multiple of 4: a group of 4 rows never straddles the wrap
                Acc4 acc[N];
                float R[N];
#pragma unroll
                for (int j = 0; j < N; j++) acc[j].lo = acc[j].hi = make_float2(0.f, 0.f);
#pragma unroll
                for (int m = 0; m < N - 1; m++) R[m] = wsm[m];
                for (int g = 0; g < P.steps; g += N) {
                    const float4 *cg = ring + (size_t)row * kFusedTW + vcs;
#define F_VLOAD(s) cg[(s) * kFusedTW]
                    PFE_GAUSS_GROUP(F_VLOAD)
#undef F_VLOAD
                    row += N;
                    if (row >= ring_rows) row -= ring_rows;
                }
                if (x < rw) {
#pragma unroll
                    for (int j = 0; j < N; j++) {
                        if (yo + j >= ye) break;
                        if (P.orig) {
                            v_store(P, acc[j], x, yo + j);
                        } else {
                            reinterpret_cast<uint32_t *>(P.dst)[(size_t)(yo + j) * P.dst_pitch + x] =
                                pfe_pack(round_u8_nonneg(acc[j].lo.x), round_u8_nonneg(acc[j].lo.y), round_u8_nonneg(acc[j].hi.x),
                                         round_u8_nonneg(acc[j].hi.y));
                        }
                    }
                }
            }
            mbar_arrive(empty0 + 8u * (uint32_t)rslot);
            if (++rslot == RB) rslot = 0;
        }
        // the last `lag` batches of the segment were only ever read as the tail of earlier outputs
        for (int i = 0; i < P.lag; i++) {
            mbar_arrive(empty0 + 8u * (uint32_t)rslot);
            if (++rslot == RB) rslot = 0;
        }
    }
}

// blur_with_selection's copy-back (filters.rs:183-200): dst = mask > 0 ? blurred : src
__global__ void select_kernel(const uint32_t *src, const uint32_t *blur, const uint8_t *mask, uint32_t *dst,
                              size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = mask[i] > 0 ? blur[i] : src[i];
}

// Selection bounding box (filters.rs:146-162): bb = {min_x, min_y, max_x, max_y}
__global__ void bbox_kernel(const uint8_t *mask, uint32_t w, uint32_t h, uint32_t *bb) {
    uint32_t mnx = 0xFFFFFFFFu, mny = 0xFFFFFFFFu, mxx = 0, mxy = 0;
    bool any = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)w * h; i += (size_t)gridDim.x * blockDim.x) {
        if (mask[i] > 0) {
            uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i - (size_t)y * w);
            mnx = min(mnx, x); mny = min(mny, y); mxx = max(mxx, x); mxy = max(mxy, y);
            any = true;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    any = __any_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && any) {
        atomicMin(&bb[0], mnx); atomicMin(&bb[1], mny); atomicMax(&bb[2], mxx); atomicMax(&bb[3], mxy);
    }
}

// build_gaussian_kernel, filters.rs:214-234 — host arithmetic with libm expf, like the reference.
std::vector<float> build_kernel(float sigma, int *radius_out) {
    float c = ceilf(sigma * 3.0f);
    int radius = (c != c || c <= 0.0f) ? 0 : (c >= 1.0e9f ? 1000000000 : (int)c);
    *radius_out = radius;
    if (radius == 0) return std::vector<float>(1, 1.0f);
    std::vector<float> k((size_t)radius * 2 + 1);
    volatile float s2 = 2.0f * sigma * sigma;
    float sum = 0.0f;
    for (size_t i = 0; i < k.size(); i++) {
        float x = (float)i - (float)radius;
        volatile float q = -x * x;   // keep each op separately rounded
        volatile float e = q / s2;
        float v = expf(e);
        k[i] = v;
        volatile float t = sum + v;
        sum = t;
    }
    volatile float inv = 1.0f / sum;
    for (float &v : k) { volatile float t = v * inv; v = t; }
    return k;
}

// Padded, duplicated weights for register-block size N: wp[m] = (w, w)[m-(N-1)] inside the support.
template <int N>
void set_steps(GaussParams &P, const std::vector<float> &k) {
    const int taps = (int)k.size();
    P.steps = ((N + taps - 1 + N - 1) / N) * N;  // N + 2r rounded up to a multiple of N
    P.wp_len = P.steps + N - 1;
    P.tri = (N > 1 && P.steps == N + taps - 1 && getenv("PFE_GAUSS_NO_TRI") == nullptr) ? 1 : 0;
    P.dbg = getenv("PFE_GAUSS_DBG") ? atoi(getenv("PFE_GAUSS_DBG")) : 0;
}
template <int N>
int upload_weights(pfe_ctx *ctx, GaussParams &P, const std::vector<float> &k, float sigma) {
    const int taps = (int)k.size();
    set_steps<N>(P, k);
    const size_t bytes = (size_t)P.wp_len * sizeof(float);
    uint32_t sigma_bits;
    memcpy(&sigma_bits, &sigma, 4);
    const bool cacheable = bytes <= pfe_ctx::kGaussSlotBytes;
    if (cacheable) {
        if (!ctx->gauss_mem) PFE_CUDA(ctx, cudaMalloc(&ctx->gauss_mem, pfe_ctx::kGaussSlots * pfe_ctx::kGaussSlotBytes));
        for (int i = 0; i < pfe_ctx::kGaussSlots; i++) {
            pfe_ctx::GaussSlot &g = ctx->gauss_slots[i];
            if (g.valid && g.sigma_bits == sigma_bits && g.n == N) {  // the table is already on the device
                g.stamp = ++ctx->gauss_clock;
                P.wp = (const float *)((char *)ctx->gauss_mem + (size_t)i * pfe_ctx::kGaussSlotBytes);
                return PFE_OK;
            }
        }
    }
    std::vector<float> wp((size_t)P.wp_len, 0.0f);
    for (int t = 0; t < taps; t++) wp[(size_t)t + N - 1] = k[(size_t)t];
    void *wdev;
    PFE_TRY(pfe_small_upload(ctx, wp.data(), bytes, &wdev));
    P.wp = (const float *)wdev;
    if (cacheable) {  // keep a copy in the least recently used slot (stream ordered after its last reader)
        int lru = 0;
        for (int i = 0; i < pfe_ctx::kGaussSlots; i++) {
            if (!ctx->gauss_slots[i].valid) { lru = i; break; }
            if (ctx->gauss_slots[i].stamp < ctx->gauss_slots[lru].stamp) lru = i;
        }
        void *slot = (char *)ctx->gauss_mem + (size_t)lru * pfe_ctx::kGaussSlotBytes;
        PFE_CUDA(ctx, cudaMemcpyAsync(slot, wdev, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->gauss_slots[lru] = pfe_ctx::GaussSlot{sigma_bits, N, ++ctx->gauss_clock, true};
    }
    return PFE_OK;
}

// Uniform-weight variants apply when the padded table fits the parameter block; PFE_GAUSS_UW=0 selects the
// shared-memory table for A/B runs.
static bool use_uniform_weights(int wp_len) {
    if (wp_len > kWeightTableLen) return false;
    const char *e = getenv("PFE_GAUSS_UW");
    return !e || atoi(e) != 0;
}
template <int N>
static void fill_weight_table(WeightTable &W, const std::vector<float> &k) {
    memset(&W, 0, sizeof(W));
    for (size_t t = 0; t < k.size() && t + N - 1 < (size_t)kWeightTableLen; t++) W.wk[t + N - 1] = k[t];
}

template <int N, bool EXACT, bool UW>
int launch_h(pfe_ctx *ctx, const GaussParams &P0, const WeightTable &W) {
    GaussParams P = P0;
    P.queue = pfe_queue_slot(ctx);
    const int wp_pad = UW ? 0 : (P.wp_len + 3) & ~3;  // float entries, keeps the tiles 16-byte aligned
    const int tile_len = skew(31 * N + P.steps, N) + 1;
    int warps = 4;
    size_t smem = (size_t)wp_pad * 4 + (size_t)warps * tile_len * 16;
    if (smem > 200 * 1024) { warps = 1; smem = (size_t)wp_pad * 4 + (size_t)tile_len * 16; }
    if (smem > 220 * 1024) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "gaussian: sigma too large for the H-pass tile");
    const uint64_t ntask = (uint64_t)pfe_div_up(P.rw, 32 * N) * P.rh;
    // Exactly one resident wave of grid-stride CTAs: the task loop strides by the grid, so a CTA that does not fit
    // next to the others would run its share alone after they finish (the r01 grid of 8 per SM did exactly that
    // where 7 fit: 30 % warps-active average and a second wave at 1/7 occupancy).
    bool done = false;
    if constexpr (UW && N >= 4) {
        if (warps == 4 && P.bb) {  // selection blur
            PFE_CUDA(ctx, cudaFuncSetAttribute(gauss_h_kernel<N, EXACT, 4, UW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned blocks = pfe_persistent_grid(ctx, gauss_h_kernel<N, EXACT, 4, UW, true>, 128, smem, (ntask + 3) / 4);
            PFE_KERNEL(ctx, "gauss_h", gauss_h_kernel<N, EXACT, 4, UW, true><<<blocks, 128, smem, ctx->stream>>>(P, W));
            done = true;
        }
    }
    if (done) {
    } else if (warps == 4) {
        PFE_CUDA(ctx, cudaFuncSetAttribute(gauss_h_kernel<N, EXACT, 4, UW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned blocks = pfe_persistent_grid(ctx, gauss_h_kernel<N, EXACT, 4, UW>, 128, smem, (ntask + 3) / 4);
        PFE_KERNEL(ctx, "gauss_h", gauss_h_kernel<N, EXACT, 4, UW><<<blocks, 128, smem, ctx->stream>>>(P, W));
    } else {
        PFE_CUDA(ctx, cudaFuncSetAttribute(gauss_h_kernel<N, EXACT, 1, UW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned blocks = pfe_persistent_grid(ctx, gauss_h_kernel<N, EXACT, 1, UW>, 32, smem, ntask);
        PFE_KERNEL(ctx, "gauss_h", gauss_h_kernel<N, EXACT, 1, UW><<<blocks, 32, smem, ctx->stream>>>(P, W));
    }
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

template <int N, bool EXACT>
int run_h(pfe_ctx *ctx, GaussParams P, const std::vector<float> &k) {
    WeightTable W;
    fill_weight_table<N>(W, k);
    set_steps<N>(P, k);
    if (use_uniform_weights(P.wp_len)) return launch_h<N, EXACT, true>(ctx, P, W);  // the table travels in the parameters
    PFE_TRY(upload_weights<N>(ctx, P, k, P.sigma));
    return launch_h<N, EXACT, false>(ctx, P, W);
}

// Consumer warps of the V tile kernel for this radius: 12 when that ring fits (measured at sigma 20, 8K: N = 8 with
// 12 warps 0.64 ms, N = 8 or 16 with 8 warps 0.67 ms), else 8, else 0 = the direct-from-global kernel.
// PFE_GAUSS_V_WARPS=8|12 and PFE_GAUSS_V_DIRECT=1 force a variant.
template <int N>
int v_tile_warps(const GaussParams &P) {
    if (N < 4 || getenv("PFE_GAUSS_V_DIRECT") != nullptr) return 0;
    const size_t extra = (size_t)((P.wp_len + 1) & ~1) * 8 + 16 * kChunks + 64;
    auto fits = [&](int warps) {
        const int rows = warps * N + P.steps - N;
        return (size_t)rows * 512 + extra <= 225 * 1024;
    };
    const char *force = getenv("PFE_GAUSS_V_WARPS");
    const int want = force ? atoi(force) : 12;
    if (want == 12 && fits(12)) return 12;
    return fits(8) ? 8 : 0;
}

// V pass: tile variant when its shared-memory footprint fits, else the direct variant
template <int N, bool EXACT, bool UW>
int launch_v(pfe_ctx *ctx, const GaussParams &P0, const WeightTable &W) {
    const int wp_pad = (P0.wp_len + 1) & ~1;
    uint32_t *const queue = pfe_queue_slot(ctx);
    const size_t extra = (size_t)wp_pad * 8 + 16 * kChunks + 64;
    auto tile_smem = [&](int warps) {
        const int rows = warps * N + P0.steps - N;
        return (size_t)rows * 512 + extra;
    };
    const int vw = v_tile_warps<N>(P0);
    // Ring chunk size in N-row groups.  A chunk is refilled for the next tile once its last reader has released it, and
    // the rows in the middle of the tile are read by every warp: the last of them at the very end of its tile, the first
    // right at the start of the next.  Coarse chunks leave that refill little slack, fine chunks cost a wait, a warp
    // barrier and an arrive each; measured, a tile in kDefaultChunks chunks is the best trade (PFE_GAUSS_VCHUNK overrides).
    GaussParams P = P0;
    {
        const int groups = (std::max(vw, 1) * N + P.steps - N) / N;
        const int forced = getenv("PFE_GAUSS_VCHUNK") ? atoi(getenv("PFE_GAUSS_VCHUNK")) : 0;
        int cg = forced > 0 ? forced : (groups + kDefaultChunks - 1) / kDefaultChunks;
        if ((groups + cg - 1) / cg > kChunks) cg = (groups + kChunks - 1) / kChunks;
        P.vchunk = cg;
        P.queue = queue;
    }
    bool done = false;
    if constexpr (UW && N >= 4) {
        if (vw == 12 && P.bb) {  // selection blur
            const size_t smem = tile_smem(12);
            const unsigned tiles12 = pfe_div_up(P.rw, 32) * pfe_div_up(P.v_rows, 12 * N);
            PFE_CUDA(ctx, cudaFuncSetAttribute(gauss_v_tile_kernel<N, EXACT, 12, UW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const unsigned blocks = pfe_persistent_grid(ctx, gauss_v_tile_kernel<N, EXACT, 12, UW, true>, 416, smem, tiles12);
            PFE_KERNEL(ctx, "gauss_v", gauss_v_tile_kernel<N, EXACT, 12, UW, true><<<blocks, 416, smem, ctx->stream>>>(P, W));
            done = true;
        }
    }
    if (done) {
    } else if (vw == 12) {
        // 12 consumer warps: a taller tile (less halo per output row) and, with N = 8, two CTAs per SM
        const size_t smem = tile_smem(12);
        const unsigned tiles12 = pfe_div_up(P.rw, 32) * pfe_div_up(P.v_rows, 12 * N);
        PFE_CUDA(ctx, cudaFuncSetAttribute(gauss_v_tile_kernel<N, EXACT, 12, UW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned blocks = pfe_persistent_grid(ctx, gauss_v_tile_kernel<N, EXACT, 12, UW>, 416, smem, tiles12);
        PFE_KERNEL(ctx, "gauss_v", gauss_v_tile_kernel<N, EXACT, 12, UW><<<blocks, 416, smem, ctx->stream>>>(P, W));
    } else if (vw == 8) {
        const size_t smem = tile_smem(8);
        const unsigned tiles8 = pfe_div_up(P.rw, 32) * pfe_div_up(P.v_rows, 8 * N);
        PFE_CUDA(ctx, cudaFuncSetAttribute(gauss_v_tile_kernel<N, EXACT, 8, UW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned blocks = pfe_persistent_grid(ctx, gauss_v_tile_kernel<N, EXACT, 8, UW>, 288, smem, tiles8);
        PFE_KERNEL(ctx, "gauss_v", gauss_v_tile_kernel<N, EXACT, 8, UW><<<blocks, 288, smem, ctx->stream>>>(P, W));
    } else {
        size_t smem = (size_t)wp_pad * 4;  // the direct variant reads its weights from P.wp (run_v uploads them)
        if (smem > 200 * 1024) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "gaussian: sigma too large");
        if (smem > 48 * 1024)
            PFE_CUDA(ctx, cudaFuncSetAttribute(gauss_v_kernel<N, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(pfe_div_up(P.rw, 128), std::min<unsigned>(pfe_div_up(P.v_rows, N), 65535u));
        PFE_KERNEL(ctx, "gauss_v", gauss_v_kernel<N, EXACT><<<grid, 128, smem, ctx->stream>>>(P));
    }
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

// V pass: tile variant when its shared-memory footprint fits, else the direct variant
template <int N, bool EXACT>
int run_v(pfe_ctx *ctx, GaussParams P, const std::vector<float> &k) {
    WeightTable W;
    fill_weight_table<N>(W, k);
    set_steps<N>(P, k);
    if (use_uniform_weights(P.wp_len) && v_tile_warps<N>(P) != 0) return launch_v<N, EXACT, true>(ctx, P, W);
    PFE_TRY(upload_weights<N>(ctx, P, k, P.sigma));
    return launch_v<N, EXACT, false>(ctx, P, W);
}

// Register-block size per pass: the one that wastes the fewest FMA-pipe slots on zero-padded steps plus per-step
// overhead (input load, weight fetch, ring bookkeeping: worth about 6 FFMA2 slots per step with the weights in
// uniform registers, half that relative to the doubled arithmetic of the EXACT path).  Measured at sigma 20 on
// 8K (profiles/r02_ops_a.jsonl): N = 8 for both passes (128 steps for 121 taps; N = 16 pads to 144).
// n16_penalty: the N = 16 V tile kernel runs one 9-warp CTA per SM at 128 registers; measured on 8K (fast path) N = 8 is
// 6 % faster at sigma 35 (211 taps) and equal at sigma 50 where this model gave N = 16 a 1-5 % edge.
static int pick_n(int taps, double per_step_overhead, const char *pass_knob, double n16_penalty = 1.0) {
    auto cost = [&](int n) {
        int steps = ((n + taps - 1 + n - 1) / n) * n;
        return (double)steps * (4.0 * n + per_step_overhead) / n * (n == 16 ? n16_penalty : 1.0);
    };
    int best = 1;
    if (taps >= 3) {
        best = 4;
        if (taps >= 9 && cost(8) < cost(best)) best = 8;
        if (taps >= 17 && cost(16) < cost(best)) best = 16;
    }
    for (const char *knob : {"PFE_GAUSS_N", pass_knob}) {  // tuning aids: both passes / this pass only
        const char *force = getenv(knob);
        if (!force) continue;
        const int n = atoi(force);
        if ((n == 4 || n == 8 || n == 16) && taps >= n + 1) best = n;
    }
    return best;
}

template <bool EXACT>
int dispatch_h(pfe_ctx *ctx, const GaussParams &P, const std::vector<float> &k) {
    switch (pick_n((int)k.size(), EXACT ? 3.0 : 6.0, "PFE_GAUSS_NH")) {
        case 16: return run_h<16, EXACT>(ctx, P, k);
        case 8: return run_h<8, EXACT>(ctx, P, k);
        case 4: return run_h<4, EXACT>(ctx, P, k);
        default: return run_h<1, EXACT>(ctx, P, k);
    }
}

template <bool EXACT>
int dispatch_v(pfe_ctx *ctx, const GaussParams &P, const std::vector<float> &k) {
    const int taps = (int)k.size();
    switch (pick_n(taps, EXACT ? 3.0 : 6.0, "PFE_GAUSS_NV", EXACT ? 1.0 : 1.07)) {
        case 16: return run_v<16, EXACT>(ctx, P, k);
        case 8: return run_v<8, EXACT>(ctx, P, k);
        case 4: return run_v<4, EXACT>(ctx, P, k);
        default: return run_v<1, EXACT>(ctx, P, k);
    }
}

// Fused H+V launch: segments are sized so that one wave of resident CTAs covers the image, but never so short that
// the 2r warm-up rows at the head of a segment dominate.
template <bool EXACT>
int run_fused(pfe_ctx *ctx, GaussParams P, const std::vector<float> &k) {
    constexpr int N = 4;
    const int taps = (int)k.size();
    P.steps = ((N + taps - 1 + N - 1) / N) * N;
    P.wp_len = P.steps + N - 1;
    P.tri = (P.steps == N + taps - 1 && getenv("PFE_GAUSS_NO_TRI") == nullptr) ? 1 : 0;
    FusedWeights W;
    memset(&W, 0, sizeof(W));
    for (int t = 0; t < taps; t++) W.wk[t + N - 1] = k[(size_t)t];
    P.lag = (2 * P.radius + 7) / 8;
    const int tile_len = skew(31 * N + P.steps, N) + 1;
    const size_t smem = 2 * kFusedMaxRing * 8 + (size_t)4 * tile_len * 16 + (size_t)(P.lag + 2) * 8 * kFusedTW * 16;
    PFE_CUDA(ctx, cudaFuncSetAttribute(gauss_fused_kernel<EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned nstrips = pfe_div_up(P.rw, kFusedTW);
    const unsigned resident = pfe_persistent_grid(ctx, gauss_fused_kernel<EXACT>, 256, smem, 0xFFFFFFFFu);
    // segments per strip: fill whole waves of resident CTAs without letting the 2r warm-up rows of a segment dominate
    const unsigned max_seg = std::max(1u, P.rh / (unsigned)std::max(64, 48 * P.lag));
    unsigned nseg = 1;
    double best = 0.0;
    for (unsigned n = 1; n <= max_seg; n++) {
        const double tasks = (double)nstrips * n, seg = (double)P.rh / n;
        const double eff = tasks / (std::ceil(tasks / resident) * resident) * seg / (seg + P.radius);
        if (eff > best + 1e-9) { best = eff; nseg = n; }
    }
    if (const char *force = getenv("PFE_GAUSS_FUSED_SEGS")) nseg = std::max(1, atoi(force));  // tuning aid
    P.seg_rows = (int)(pfe_div_up(pfe_div_up(P.rh, nseg), 8) * 8);
    P.nseg = (int)pfe_div_up(P.rh, (unsigned)P.seg_rows);
    const unsigned blocks = std::min(resident, nstrips * (unsigned)P.nseg);
    PFE_KERNEL(ctx, "gauss_fused", gauss_fused_kernel<EXACT><<<blocks, 256, smem, ctx->stream>>>(P, W));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

// The fused kernel wins wherever it applies (radius <= 16) provided the image has enough strip segments to occupy
// every SM (measured on B200, profiles/r01_gauss_fused.jsonl: 1.5x at radius 2 down to 1.03x at radius 16 on 4K / 8K
// images, but 2-4x slower on a 0.7 Mpx image).  PFE_GAUSS_FUSED=1 / =0 force it.
static bool use_fused(const pfe_ctx *ctx, int radius, uint32_t rw, uint32_t rh) {
    if (radius < 1 || radius > kFusedMaxRadius) return false;
    if (const char *force = getenv("PFE_GAUSS_FUSED")) return atoi(force) != 0;
    const int lag = (2 * radius + 7) / 8;
    const uint64_t tasks = (uint64_t)pfe_div_up(rw, kFusedTW) * std::max(1u, rh / (unsigned)std::max(64, 48 * lag));
    return tasks >= (uint64_t)ctx->sm_count;
}

template <bool EXACT>
int dispatch_n(pfe_ctx *ctx, const GaussParams &P, const std::vector<float> &k) {
    if (use_fused(ctx, P.radius, P.rw, P.rh)) return run_fused<EXACT>(ctx, P, k);
    PFE_TRY(dispatch_h<EXACT>(ctx, P, k));
    return dispatch_v<EXACT>(ctx, P, k);
}

int gauss_common(pfe_ctx *ctx, const uint8_t *src, uint8_t *dst, uint32_t src_pitch, uint32_t dst_pitch,
                 uint32_t rw, uint32_t rh, float sigma, uint32_t flags, const uint8_t *orig, float amount,
                 const uint8_t *mask, uint32_t mask_pitch, int epilogue = 1) {
    int radius;
    std::vector<float> k = build_kernel(sigma, &radius);
    if (radius > 4000) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "gaussian: sigma too large");
    void *mid = nullptr;  // the fused kernel keeps the intermediate in shared memory
    if (!use_fused(ctx, radius, rw, rh)) PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_F32, (size_t)rw * rh * 16, &mid));
    GaussParams P;
    memset(&P, 0, sizeof(P));
    P.one = 1.0f; P.nzero = -0.0f;
    P.src = src; P.mid = (float *)mid; P.dst = dst;
    P.orig = orig; P.mask = mask; P.amount = amount; P.epilogue = orig ? epilogue : 0;
    P.src_pitch = src_pitch; P.dst_pitch = dst_pitch; P.mask_pitch = mask_pitch;
    P.rw = rw; P.rh = rh; P.radius = radius; P.sigma = sigma;
    P.v_y0 = 0; P.v_rows = rh;
    return (flags & PFE_GAUSS_EXACT) ? dispatch_n<true>(ctx, P, k) : dispatch_n<false>(ctx, P, k);
}

}  // namespace

// Band-pipelining hooks (csrc/ctx.cu): the H pass is row-local, so it can run on rows [y0, y0+rows)
// as soon as they exist; the V pass of rows [y0, y0+rows) needs the H rows within +-radius of them.
// `mid` is the caller-provided w*h*4 f32 intermediate of the whole image.
int pfe_gauss_h_rows(pfe_ctx *ctx, const uint8_t *src, float *mid, uint32_t w, uint32_t h, uint32_t y0, uint32_t rows,
                     float sigma, uint32_t flags) {
    int radius;
    std::vector<float> k = build_kernel(sigma, &radius);
    if (radius > 4000) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "gaussian: sigma too large");
    GaussParams P;
    memset(&P, 0, sizeof(P));
    P.one = 1.0f; P.nzero = -0.0f;
    P.src = src + (size_t)y0 * w * 4; P.mid = mid + (size_t)y0 * w * 4;
    P.src_pitch = w; P.dst_pitch = w; P.rw = w; P.rh = rows; P.radius = radius; P.sigma = sigma;
    (void)h;
    return (flags & PFE_GAUSS_EXACT) ? dispatch_h<true>(ctx, P, k) : dispatch_h<false>(ctx, P, k);
}
int pfe_gauss_v_rows(pfe_ctx *ctx, float *mid, uint8_t *dst, uint32_t w, uint32_t h, uint32_t y0, uint32_t rows,
                     float sigma, uint32_t flags) {
    int radius;
    std::vector<float> k = build_kernel(sigma, &radius);
    GaussParams P;
    memset(&P, 0, sizeof(P));
    P.one = 1.0f; P.nzero = -0.0f;
    P.mid = mid; P.dst = dst; P.src_pitch = w; P.dst_pitch = w; P.rw = w; P.rh = h; P.radius = radius; P.sigma = sigma;
    P.v_y0 = y0; P.v_rows = rows;
    return (flags & PFE_GAUSS_EXACT) ? dispatch_v<true>(ctx, P, k) : dispatch_v<false>(ctx, P, k);
}
int pfe_gauss_radius(float sigma) {
    int radius;
    build_kernel(sigma, &radius);
    return radius;
}

int pfe_gauss_region(pfe_ctx *ctx, const uint8_t *src, uint8_t *dst, uint32_t pitch_px, uint32_t x0, uint32_t y0,
                     uint32_t rw, uint32_t rh, float sigma, uint32_t flags) {
    const uint8_t *s = src + ((size_t)y0 * pitch_px + x0) * 4;
    uint8_t *d = dst + ((size_t)y0 * pitch_px + x0) * 4;
    return gauss_common(ctx, s, d, pitch_px, pitch_px, rw, rh, sigma, flags, nullptr, 0.0f, nullptr, 0);
}

extern "C" int pfe_dev_gaussian_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float sigma,
                                     const uint8_t *mask, uint8_t *dst, uint32_t flags) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !dst || !w || !h) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "gaussian: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!mask) {
        // in place would race: CTAs store dst rows while neighbouring strips / tiles still read them as source
        if (src == dst) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "gaussian: in-place not supported without a mask");
        return gauss_common(ctx, src, dst, w, w, w, h, sigma, flags, nullptr, 0.0f, nullptr, 0);
    }
    // blur_with_selection, filters.rs:141-207.  The bounding box stays on the device (see bb_skip): no read-back, no
    // stream synchronisation - the call is asynchronous like every other pfe_dev_* entry point.
    void *bbd;
    uint32_t init[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u};
    PFE_TRY(pfe_small_upload(ctx, init, sizeof(init), &bbd));
    PFE_KERNEL(ctx, "mask_bbox", bbox_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(mask, w, h, (uint32_t *)bbd));
    PFE_LAUNCHED(ctx);
    const size_t n = (size_t)w * h;
    int radius;
    std::vector<float> k = build_kernel(sigma, &radius);
    if (radius > 4000) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "gaussian: sigma too large");
    void *tmp, *mid;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_B, n * 4, &tmp));
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_F32, n * 16, &mid));
    GaussParams P;
    memset(&P, 0, sizeof(P));
    P.one = 1.0f; P.nzero = -0.0f;
    P.src = src; P.mid = (float *)mid; P.dst = (uint8_t *)tmp;
    P.src_pitch = w; P.dst_pitch = w; P.rw = w; P.rh = h; P.radius = radius; P.sigma = sigma;
    P.v_y0 = 0; P.v_rows = h;
    P.bb = (const uint32_t *)bbd;
    if (flags & PFE_GAUSS_EXACT) { PFE_TRY(dispatch_h<true>(ctx, P, k)); PFE_TRY(dispatch_v<true>(ctx, P, k)); }
    else { PFE_TRY(dispatch_h<false>(ctx, P, k)); PFE_TRY(dispatch_v<false>(ctx, P, k)); }
    // pixels the kernels skipped are never selected (they lie outside the mask's bounding box), so the stale parts of
    // tmp are masked out by mask == 0
    PFE_KERNEL(ctx, "select", select_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((const uint32_t *)src, (const uint32_t *)tmp, mask,
                                                              (uint32_t *)dst, n));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

// Row-band form for a canvas split across GPUs: the caller's extended band (own rows + halo rows
// received from the neighbours) is filtered with the ordinary kernels; the H pass can be issued per row range so
// that the band's own rows are filtered while the halo is still in flight.
extern "C" int pfe_dev_gaussian_band_h(pfe_ctx *ctx, const uint8_t *ext, uint32_t w, uint32_t ext_rows, uint32_t y0,
                                       uint32_t rows, float sigma, uint32_t flags) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!ext || !w || !ext_rows || (uint64_t)y0 + rows > ext_rows) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "gaussian_band_h: bad args");
    if (!rows) return PFE_OK;
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    void *mid;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_F32, (size_t)w * ext_rows * 16, &mid));
    return pfe_gauss_h_rows(ctx, ext, (float *)mid, w, ext_rows, y0, rows, sigma, flags);
}
extern "C" int pfe_dev_gaussian_band_v(pfe_ctx *ctx, uint32_t w, uint32_t ext_rows, uint32_t y0, uint32_t rows, float sigma,
                                       uint8_t *dst_rows, uint32_t flags) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!dst_rows || !w || !ext_rows || (uint64_t)y0 + rows > ext_rows) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "gaussian_band_v: bad args");
    if (!rows) return PFE_OK;
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    if ((size_t)w * ext_rows * 16 > ctx->scratch_bytes[PFE_SCRATCH_F32])
        return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "gaussian_band_v: no H pass of this band precedes it");
    // the V kernels address dst by absolute ext row: hand them the address row 0 would have
    uint8_t *base = dst_rows - (size_t)y0 * w * 4;
    return pfe_gauss_v_rows(ctx, (float *)ctx->scratch[PFE_SCRATCH_F32], base, w, ext_rows, y0, rows, sigma, flags);
}

extern "C" int pfe_dev_sharpen(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, float radius,
                               const uint8_t *mask, uint8_t *dst, uint32_t flags) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !dst || !w || !h || src == dst) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "sharpen: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    // blur(sigma = radius) with the unsharp epilogue fused into the V pass: the blurred image is
    // quantised to u8 in registers exactly as the reference stores it, never written to memory.
    return gauss_common(ctx, src, dst, w, w, w, h, radius, flags, src, amount, mask, w);
}

extern "C" int pfe_dev_glow(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius, float intensity,
                            const uint8_t *mask, uint8_t *dst, uint32_t flags) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !dst || !w || !h || src == dst) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "glow: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    // blur(sigma = radius) with the screen-glow epilogue fused into the V pass (stylize.rs:26-76)
    return gauss_common(ctx, src, dst, w, w, w, h, radius, flags, src, intensity, mask, w, 2);
}
