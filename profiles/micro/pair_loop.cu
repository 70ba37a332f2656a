// Microbenchmark of the NMS pair-mask inner loop (decode_nms.cuh, block_fast): which formulation issues the fewest
// slots per box pair on sm_100a?  Lanes = rows, columns broadcast from shared memory, 32 columns per mask word.
//   V0  the round-1 loop: sorted table {shared address of the box, t*area} -> LDS.64 + dependent LDS.128, scalar fp32
//   V1  sorted structure-of-arrays (box float4[k], ta float[k]): LDS.128 + LDS.32 at immediate offsets, scalar fp32
//   V2  V1 + FMNMX3 for the running "too close" minimum (one ALU-pipe instruction per two pairs)
//   V3  two columns per step with packed fp32x2 (FADD2 / FFMA2), FMNMX3, ta pairs by LDS.64
//   V4  the diagonal tile with cyclic pairing (lane l meets (l+k)&31, k = 1..16): per-lane addresses into the sorted
//       SoA (conflict-free: consecutive lanes, consecutive entries), wrapped bits moved with one ballot per step
// Every variant must produce the same mask words (checked on the host).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pair_loop pair_loop.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
typedef unsigned long long u64;
constexpr int NB = 96;            // boxes per CTA (3 tiles)
constexpr float kEps = 1e-5f;
constexpr float kScale = 1.220703125e-4f;  // 2^-13

__device__ __forceinline__ float4 lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

#include <cuda_fp16.h>
// fp16 column table of one 32-column tile: entry k packs columns k (low half) and k+16 (high half)
struct H16Tile {
    uint4 xy[16];       // x = X1 pair, y = Y1 pair, z = X2 pair, w = Y2 pair (half2 each)
    uint32_t ta[16];    // TA pair
};
constexpr float kSX = 0.5f, kSY = 32.0f;

struct Sm {
    H16Tile h16[NB / 32];
    float4 box[NB];     // sorted boxes
    float ta[NB + 8];   // t * area * 2^-13
    uint2 ord[NB + 8];  // {shared address of box, ta bits}
};

__device__ __forceinline__ void step_scalar(const float4 Cb, float cta, const float4 &R, float rta, uint32_t &bits, float &m) {
    const float w = __fsub_rn(fminf(R.z, Cb.z), fmaxf(R.x, Cb.x));
    const float h = __fsub_rn(fminf(R.w, Cb.w), fmaxf(R.y, Cb.y));
    const float ws = __saturatef(__fmul_rn(w, kScale));
    const float sum = __fadd_rn(rta, cta);
    const float d = __fmaf_rn(-ws, h, sum);
    m = fminf(m, __fmaf_rn(sum, -kEps, fabsf(d)));
    bits = __funnelshift_l(__float_as_uint(d), bits, 1);
}

template <int V>
__device__ __forceinline__ uint32_t block32(const Sm &s, int c0, const float4 R, float rta, float &mout) {
    uint32_t bits = 0u;
    float m = INFINITY;
    if (V == 0) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const uint2 e = s.ord[c0 + k];
            step_scalar(lds_f4(e.x), __uint_as_float(e.y), R, rta, bits, m);
        }
    } else if (V == 1) {
#pragma unroll
        for (int k = 0; k < 32; ++k) step_scalar(s.box[c0 + k], s.ta[c0 + k], R, rta, bits, m);
    } else if (V == 2) {
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
            float g[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const float4 Cb = s.box[c0 + k + u];
                const float cta = s.ta[c0 + k + u];
                const float w = __fsub_rn(fminf(R.z, Cb.z), fmaxf(R.x, Cb.x));
                const float h = __fsub_rn(fminf(R.w, Cb.w), fmaxf(R.y, Cb.y));
                const float ws = __saturatef(__fmul_rn(w, kScale));
                const float sum = __fadd_rn(rta, cta);
                const float d = __fmaf_rn(-ws, h, sum);
                g[u] = __fmaf_rn(sum, -kEps, fabsf(d));
                bits = __funnelshift_l(__float_as_uint(d), bits, 1);
            }
            m = fmin3(m, g[0], g[1]);
        }
    } else if (V == 3) {
        const u64 rta2 = pk(rta, rta);
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
            const float4 C0 = s.box[c0 + k], C1 = s.box[c0 + k + 1];
            const float2 ct = *reinterpret_cast<const float2 *>(&s.ta[c0 + k]);
            const u64 w2 = sub2(pk(fminf(R.z, C0.z), fminf(R.z, C1.z)), pk(fmaxf(R.x, C0.x), fmaxf(R.x, C1.x)));
            const u64 h2 = sub2(pk(fminf(R.w, C0.w), fminf(R.w, C1.w)), pk(fmaxf(R.y, C0.y), fmaxf(R.y, C1.y)));
            float w0, w1;
            upk(w2, w0, w1);
            const u64 nws2 = pk(-__saturatef(__fmul_rn(w0, kScale)), -__saturatef(__fmul_rn(w1, kScale)));
            const u64 sum2 = add2(rta2, pk(ct.x, ct.y));
            const u64 d2 = fma2(nws2, h2, sum2);
            float d0, d1, s0, s1;
            upk(d2, d0, d1);
            upk(sum2, s0, s1);
            m = fmin3(m, __fmaf_rn(s0, -kEps, fabsf(d0)), __fmaf_rn(s1, -kEps, fabsf(d1)));
            bits = __funnelshift_l(__float_as_uint(d0), bits, 1);
            bits = __funnelshift_l(__float_as_uint(d1), bits, 1);
        }
    }
    mout = m;
    return __brev(bits);
}

__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ __half2 u2h(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t hmin2(uint32_t a, uint32_t b) { return h2u(__hmin2(u2h(a), u2h(b))); }
__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b) { return h2u(__hmax2(u2h(a), u2h(b))); }
__device__ __forceinline__ uint32_t hsub2(uint32_t a, uint32_t b) { return h2u(__hsub2(u2h(a), u2h(b))); }
__device__ __forceinline__ uint32_t hsub2_sat(uint32_t a, uint32_t b) { return h2u(__hsub2_sat(u2h(a), u2h(b))); }
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { return h2u(__hadd2(u2h(a), u2h(b))); }
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) { return h2u(__hfma2(u2h(a), u2h(b), u2h(c))); }

// conservative fp16 prefilter of one 32-column tile: bit j set = column j MAY be suppressed by this lane's row
__device__ __forceinline__ uint32_t prefilter32(const H16Tile &t, uint32_t rx1, uint32_t ry1, uint32_t rx2, uint32_t ry2, uint32_t rnta) {
    uint32_t acc = 0u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint4 c = t.xy[k];
        const uint32_t w = hsub2_sat(hmin2(rx2, c.z), hmax2(rx1, c.x));
        const uint32_t h = hsub2(hmin2(ry2, c.w), hmax2(ry1, c.y));
        const uint32_t nsum = hsub2(rnta, t.ta[k]);              // -(TAr + TAc)
        const uint32_t d = hfma2(w, h, nsum);                    // >= 0: maybe
        acc = (acc >> 1) | (d & 0x80008000u);
    }
    return ~acc;   // bit k = column k, bit 16 + k = column 16 + k
}

// cyclic diagonal: rows/columns of ONE tile starting at c0; returns the word of row `lane` restricted to later columns
__device__ __forceinline__ uint32_t diag_cyclic(const Sm &s, int c0, const float4 R, float rta, float &mout) {
    const int lane = threadIdx.x & 31;
    uint32_t word = 0u;
    float m = INFINITY;
#pragma unroll
    for (int k = 1; k <= 16; ++k) {
        const int j = (lane + k) & 31;
        const float4 Cb = s.box[c0 + j];
        const float cta = s.ta[c0 + j];
        const float w = __fsub_rn(fminf(R.z, Cb.z), fmaxf(R.x, Cb.x));
        const float h = __fsub_rn(fminf(R.w, Cb.w), fmaxf(R.y, Cb.y));
        const float ws = __saturatef(__fmul_rn(w, kScale));
        const float sum = __fadd_rn(rta, cta);
        const float d = __fmaf_rn(-ws, h, sum);
        m = fminf(m, __fmaf_rn(sum, -kEps, fabsf(d)));
        const bool sup = d < 0.0f;
        const bool wrap = j < lane;                     // the pair belongs to row j, column lane
        const bool own = (k < 16) || (lane < 16);        // k == 16: every pair is met from both sides
        const uint32_t bal = __ballot_sync(0xffffffffu, sup && wrap && own);
        if (sup && !wrap && own) word |= 1u << j;
        // lane r receives the pairs (l, r) with l = (r - k) & 31 > r  <=> wrapped at lane l
        const int l = (lane - k) & 31;
        if (l > lane && ((bal >> l) & 1u)) word |= 1u << l;
    }
    mout = m;
    return word;
}

template <int V>
__global__ void __launch_bounds__(512, 2) bench_kernel(const float4 *boxes, int iters, uint32_t *out, float tsc) {
    __shared__ Sm s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < NB; i += blockDim.x) {
        const float4 b = boxes[(size_t)blockIdx.x * NB + i];
        s.box[i] = b;
        const float ta = (b.z - b.x) * (b.w - b.y) * tsc;
        s.ta[i] = ta;
        s.ord[i] = make_uint2((uint32_t)__cvta_generic_to_shared(&s.box[i]), __float_as_uint(ta));
    }
    __syncthreads();
    for (int i = tid; i < NB; i += blockDim.x) {
        const float4 b = s.box[i];
        H16Tile &t = s.h16[i >> 5];
        const int k = i & 15, hi = (i & 31) >> 4;
        __half *px = reinterpret_cast<__half *>(&t.xy[k]);
        px[0 + hi] = __float2half_rd(b.x * kSX);
        px[2 + hi] = __float2half_rd(b.y * kSY);
        px[4 + hi] = __float2half_ru(b.z * kSX);
        px[6 + hi] = __float2half_ru(b.w * kSY);
        const float area = (b.z - b.x) * (b.w - b.y);
        reinterpret_cast<__half *>(&t.ta[k])[hi] = __float2half_rd(area * (0.45f / 1.45f) * kSX * kSY * (1.0f - 0.00390625f));
    }
    __syncthreads();
    if (tid < 8) { s.ta[NB + tid] = __int_as_float(0x7fc00000); s.ord[NB + tid] = make_uint2((uint32_t)__cvta_generic_to_shared(&s.box[0]), 0x7fc00000u); }
    __syncthreads();
    uint32_t acc = 0u;
    float macc = INFINITY;
    for (int it = 0; it < iters; ++it) {
        const int rt = (warp + it) % 3, ct = (it >> 2) % 3;
        const float4 R = s.box[32 * rt + lane];
        const float rta = s.ta[32 * rt + lane];
        float m;
        uint32_t w;
        if (V == 5) {
            const int r = 32 * rt + lane;
            const H16Tile &rtile = s.h16[rt];
            const int k = r & 15, hi = (r & 31) >> 4;
            const uint32_t sel = hi ? 0x3232u : 0x1010u;
            const uint4 e = rtile.xy[k];
            w = prefilter32(s.h16[ct], __byte_perm(e.x, 0, sel), __byte_perm(e.y, 0, sel), __byte_perm(e.z, 0, sel), __byte_perm(e.w, 0, sel),
                            __byte_perm(rtile.ta[k], 0, sel) ^ 0x80008000u);
            m = 1.f;
        } else if (V == 4) w = diag_cyclic(s, 32 * ct, s.box[32 * ct + lane], s.ta[32 * ct + lane], m);
        else w = block32<V>(s, 32 * ct, R, rta, m);
        acc = acc * 31u + w;
        macc = fminf(macc, m);
    }
    out[(size_t)blockIdx.x * blockDim.x + tid] = acc ^ (macc <= 0.f ? 1u : 0u);
}

// check kernel: one block evaluation per variant on the same data
template <int V>
__global__ void check_kernel(const float4 *boxes, uint32_t *out, float tsc, int diag) {
    __shared__ Sm s;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < NB; i += blockDim.x) {
        const float4 b = boxes[i];
        s.box[i] = b;
        const float ta = (b.z - b.x) * (b.w - b.y) * tsc;
        s.ta[i] = ta;
        s.ord[i] = make_uint2((uint32_t)__cvta_generic_to_shared(&s.box[i]), __float_as_uint(ta));
    }
    __syncthreads();
    for (int i = tid; i < NB; i += blockDim.x) {
        const float4 b = s.box[i];
        H16Tile &t = s.h16[i >> 5];
        const int k = i & 15, hi = (i & 31) >> 4;
        __half *px = reinterpret_cast<__half *>(&t.xy[k]);
        px[0 + hi] = __float2half_rd(b.x * kSX);
        px[2 + hi] = __float2half_rd(b.y * kSY);
        px[4 + hi] = __float2half_ru(b.z * kSX);
        px[6 + hi] = __float2half_ru(b.w * kSY);
        const float area = (b.z - b.x) * (b.w - b.y);
        reinterpret_cast<__half *>(&t.ta[k])[hi] = __float2half_rd(area * (0.45f / 1.45f) * kSX * kSY * (1.0f - 0.00390625f));
    }
    __syncthreads();
    float m;
    uint32_t w;
    if (V == 5) {
        const int r = (diag ? 32 : 0) + lane;
        const H16Tile &rtile = s.h16[r >> 5];
        const int k = r & 15, hi = (r & 31) >> 4;
        const uint32_t sel = hi ? 0x3232u : 0x1010u;
        const uint4 e = rtile.xy[k];
        const uint32_t rx1 = __byte_perm(e.x, 0, sel), ry1 = __byte_perm(e.y, 0, sel), rx2 = __byte_perm(e.z, 0, sel), ry2 = __byte_perm(e.w, 0, sel);
        const uint32_t rnta = __byte_perm(rtile.ta[k], 0, sel) ^ 0x80008000u;
        w = prefilter32(s.h16[1], rx1, ry1, rx2, ry2, rnta);
        if (diag) w &= ~((2u << lane) - 1u);
        out[lane] = w;
        return;
    }
    if (V == 4) w = diag_cyclic(s, 32, s.box[32 + lane], s.ta[32 + lane], m);
    else {
        w = block32<V>(s, 32, s.box[(diag ? 32 : 0) + lane], s.ta[(diag ? 32 : 0) + lane], m);
        if (diag) w &= ~((2u << lane) - 1u);
    }
    out[lane] = w;
}

int main() {
    const int grid = 296, threads = 512, iters = 3000;
    std::vector<float4> hb((size_t)grid * NB);
    srand(1);
    for (auto &b : hb) {
        const float cx = rand() / (float)RAND_MAX, cy = rand() / (float)RAND_MAX;
        const float w = 0.05f + 0.5f * rand() / (float)RAND_MAX, h = 0.05f + 0.5f * rand() / (float)RAND_MAX;
        b = make_float4(cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2);
    }
    float4 *db;
    uint32_t *dout;
    CK(cudaMalloc(&db, hb.size() * sizeof(float4)));
    CK(cudaMalloc(&dout, (size_t)grid * threads * 4));
    CK(cudaMemcpy(db, hb.data(), hb.size() * sizeof(float4), cudaMemcpyHostToDevice));
    const float tsc = (float)(0.45 / 1.45) * kScale;
    // correctness: all variants agree on an off-diagonal block; V4 agrees with the masked diagonal block of V1
    uint32_t ref[32], got[32];
    check_kernel<0><<<1, 32>>>(db, dout, tsc, 0);
    CK(cudaMemcpy(ref, dout, 128, cudaMemcpyDeviceToHost));
    int nset = 0;
    for (int i = 0; i < 32; ++i) nset += __builtin_popcount(ref[i]);
#define CHECK(V, D, REF) do { check_kernel<V><<<1, 32>>>(db, dout, tsc, D); CK(cudaMemcpy(got, dout, 128, cudaMemcpyDeviceToHost)); \
        int bad = 0; for (int i = 0; i < 32; ++i) bad += got[i] != REF[i]; printf("check V%d diag=%d: %s\n", V, D, bad ? "MISMATCH" : "ok"); } while (0)
    CHECK(1, 0, ref); CHECK(2, 0, ref); CHECK(3, 0, ref);
    uint32_t refd[32];
    check_kernel<1><<<1, 32>>>(db, dout, tsc, 1);
    CK(cudaMemcpy(refd, dout, 128, cudaMemcpyDeviceToHost));
    CHECK(4, 1, refd);
    printf("suppress bits in the reference block: %d of 1024\n", nset);
    {   // the fp16 prefilter must cover every true suppress bit; report how many extra "maybe" bits it raises
        check_kernel<5><<<1, 32>>>(db, dout, tsc, 0); CK(cudaMemcpy(got, dout, 128, cudaMemcpyDeviceToHost));
        int miss = 0, extra = 0; for (int i = 0; i < 32; ++i) { miss += __builtin_popcount(ref[i] & ~got[i]); extra += __builtin_popcount(got[i] & ~ref[i]); }
        printf("prefilter V5 offdiag: missed %d (must be 0), extra maybe bits %d of 1024\n", miss, extra);
        check_kernel<5><<<1, 32>>>(db, dout, tsc, 1); CK(cudaMemcpy(got, dout, 128, cudaMemcpyDeviceToHost));
        miss = 0; extra = 0; for (int i = 0; i < 32; ++i) { miss += __builtin_popcount(refd[i] & ~got[i]); extra += __builtin_popcount(got[i] & ~refd[i]); }
        printf("prefilter V5 diag: missed %d (must be 0), extra maybe bits %d\n", miss, extra);
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
#define RUN(V, PAIRS) do { bench_kernel<V><<<grid, threads>>>(db, 10, dout, tsc); CK(cudaDeviceSynchronize()); \
        CK(cudaEventRecord(e0)); bench_kernel<V><<<grid, threads>>>(db, iters, dout, tsc); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); \
        const double blocks = (double)grid * (threads / 32) * iters; \
        const double clk_per_block_smsp = ms * 1e-3 * clk_khz * 1e3 / (blocks / (148.0 * 4)); \
        printf("V%d: %.3f ms, %.1f issue-clk per 32-lane block per SMSP (%.2f per lane-pair step; useful pairs per block %d -> %.2f clk per 32 useful pairs)\n", \
               V, ms, clk_per_block_smsp, clk_per_block_smsp / 32, PAIRS, clk_per_block_smsp * 32.0 / PAIRS); } while (0)
    RUN(0, 1024); RUN(1, 1024); RUN(2, 1024); RUN(3, 1024); RUN(4, 496); RUN(5, 1024);
    return 0;
}
