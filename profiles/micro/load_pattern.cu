// Microbenchmark: how fast can one CTA per image stream its (A*(5+C), H, W) slabs with the
// thread-per-cell access pattern of the decode phase?  Build: nvcc -arch=sm_100a -O3 -o lp load_pattern.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int POLICY> __device__ __forceinline__ float ld(const float *p) {
    if (POLICY == 0) return __ldcs(p);
    if (POLICY == 1) return __ldg(p);
    return *p;
}

// variant 0: thread per cell, 25 scalar loads per cell, all in flight, rounds of blockDim cells
template <int POLICY>
__global__ void __launch_bounds__(512, 2) k_cell(const float *h0, const float *h1, int HW0, int HW1, float *out) {
    extern __shared__ float pad[];
    const int b = blockIdx.x, tid = threadIdx.x;
    float acc = 0.f;
    for (int hh = 0; hh < 2; ++hh) {
        const int HW = hh ? HW1 : HW0;
        const float *hb = (hh ? h1 : h0) + (size_t)b * 75 * HW;
        for (int base = 0; base < 3 * HW; base += blockDim.x) {
            const int local = base + tid;
            if (local < 3 * HW) {
                const int a = local / HW, pos = local - a * HW;
                const float *q = hb + (size_t)a * 25 * HW + pos;
                float x[25];
#pragma unroll
                for (int u = 0; u < 25; ++u) x[u] = ld<POLICY>(q + (size_t)u * HW);
                float m = x[0];
#pragma unroll
                for (int u = 1; u < 25; ++u) m = fmaxf(m, x[u]);
                acc += m;
            }
        }
    }
    if (acc == 12345.678f) out[b * blockDim.x + tid] = acc;
    if (tid == 0) pad[0] = acc;
}

// variant 1: flat float4 streaming of the whole image slab (perfectly coalesced, no structure)
__global__ void __launch_bounds__(512, 2) k_flat(const float *h0, const float *h1, int HW0, int HW1, float *out) {
    extern __shared__ float pad[];
    const int b = blockIdx.x, tid = threadIdx.x;
    float acc = 0.f;
    {
        const float4 *p = reinterpret_cast<const float4 *>(h1 + (size_t)b * 75 * HW1);
        const int n4 = 75 * HW1 / 4;
        for (int i = tid; i < n4; i += blockDim.x) { float4 v = __ldcs(p + i); acc += v.x + v.y + v.z + v.w; }
        const float *p0 = h0 + (size_t)b * 75 * HW0;
        for (int i = tid; i < 75 * HW0; i += blockDim.x) acc += __ldcs(p0 + i);
    }
    if (acc == 12345.678f) out[b * blockDim.x + tid] = acc;
    if (tid == 0) pad[0] = acc;
}

// variant 2: flat float4 streaming, 8 independent loads in flight per thread
__global__ void __launch_bounds__(512, 2) k_flat8(const float *h0, const float *h1, int HW0, int HW1, float *out) {
    extern __shared__ float pad[];
    const int b = blockIdx.x, tid = threadIdx.x;
    float acc = 0.f;
    {
        const float4 *p = reinterpret_cast<const float4 *>(h1 + (size_t)b * 75 * HW1);
        const int n4 = 75 * HW1 / 4;
        int i = tid;
        for (; i + 7 * (int)blockDim.x < n4; i += 8 * blockDim.x) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcs(p + i + u * blockDim.x);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
        for (; i < n4; i += blockDim.x) { float4 v = __ldcs(p + i); acc += v.x + v.y + v.z + v.w; }
        const float *p0 = h0 + (size_t)b * 75 * HW0;
        for (int j = tid; j < 75 * HW0; j += blockDim.x) acc += __ldcs(p0 + j);
    }
    if (acc == 12345.678f) out[b * blockDim.x + tid] = acc;
    if (tid == 0) pad[0] = acc;
}

// ---- variant 3: bulk async copies (TMA engine, bypasses L1) into a small smem ring -----------
__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n .reg .pred P1;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n }" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// head 1 (HW % 4 == 0): stages of (512 cells x G planes); head 0: plain loads.  S-stage ring.
template <int G, int S>
__global__ void __launch_bounds__(512, 2) k_tma(const float *h0, const float *h1, int HW0, int HW1, float *out) {
    extern __shared__ __align__(128) unsigned char sm[];
    float *ring = reinterpret_cast<float *>(sm);                      // [S][G][512]
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + S * G * 2048);  // full[S], empty[S]
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(saddr(&bars[i]), 1); mbar_init(saddr(&bars[S + i]), 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const float *hb = h1 + (size_t)b * 75 * HW1;
    const int cells = 3 * HW1;
    const int rounds = (cells + 511) / 512, groups = 25 / G;  // 25 % G == 0 here
    const int nst = rounds * groups;
    auto issue = [&](int st) {  // warp 0: lane l -> plane l / 2, segment l % 2
        const int r = st / groups, g = st - r * groups, slot = st % S;
        const int c0 = r * 512, c1 = min(cells, c0 + 512);
        const uint32_t full = saddr(&bars[slot]);
        if (lane == 0) mbar_expect_tx(full, (uint32_t)(G * (c1 - c0) * 4));
        __syncwarp();
        for (int cp = lane; cp < 2 * G; cp += 32) {
            const int pl = cp >> 1, seg = cp & 1;
            const int a0 = c0 / HW1;
            const int a = a0 + seg;
            const int s0 = max(c0, a * HW1), s1 = min(c1, (a + 1) * HW1);
            if (s1 > s0) {
                const float *src = hb + ((size_t)a * 25 + g * G + pl) * HW1 + (s0 - a * HW1);
                bulk_g2s(saddr(ring + ((size_t)slot * G + pl) * 512 + (s0 - c0)), src, (uint32_t)(s1 - s0) * 4, full);
            }
        }
    };
    float acc = 0.f;
    if (warp == 0) for (int st = 0; st < min(S - 1, nst); ++st) issue(st);
    // head 0 through L1 while the ring fills
    {
        const float *p0 = h0 + (size_t)b * 75 * HW0;
        const int local = tid;
        if (local < 3 * HW0) {
            const int a = local / HW0, pos = local - a * HW0;
            const float *q = p0 + (size_t)a * 25 * HW0 + pos;
            float x[25];
#pragma unroll
            for (int u = 0; u < 25; ++u) x[u] = __ldcs(q + (size_t)u * HW0);
            float m = x[0];
#pragma unroll
            for (int u = 1; u < 25; ++u) m = fmaxf(m, x[u]);
            acc += m;
        }
    }
    for (int st = 0; st < nst; ++st) {
        const int slot = st % S, use = st / S;
        if (warp == 0) {
            const int nx = st + S - 1;
            if (nx < nst) {
                const int nslot = nx % S, nuse = nx / S;
                if (nuse > 0) mbar_wait(saddr(&bars[S + nslot]), (uint32_t)((nuse - 1) & 1));
                issue(nx);
            }
        }
        mbar_wait(saddr(&bars[slot]), (uint32_t)(use & 1));
        const float *stg = ring + (size_t)slot * G * 512;
        float m = stg[tid];
#pragma unroll
        for (int u = 1; u < G; ++u) m = fmaxf(m, stg[u * 512 + tid]);
        acc += m;
        __syncwarp();
        if (lane == 0) mbar_arrive(saddr(&bars[S + slot]));
    }
    if (acc == 12345.678f) out[b * blockDim.x + tid] = acc;
}


// ---- variant 4: plane-major stages.  One bulk copy per stage = a CONTIGUOUS run of whole planes of one
// (image, anchor) slab (16-byte-aligned superset of it), S-slot ring, thread per cell keeps a running value.
template <int S, int SLOT>
__global__ void __launch_bounds__(512, 2) k_planes(const float *h0, const float *h1, int HW0, int HW1, float *out) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + S * SLOT);  // full[S], empty[S]
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(saddr(&bars[i]), 1); mbar_init(saddr(&bars[S + i]), 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int G1 = max(1, min(25, (SLOT - 32) / (HW1 * 4)));   // planes per stage, head 1
    const int G0 = max(1, min(25, (SLOT - 32) / (HW0 * 4)));
    const int per0 = (25 + G0 - 1) / G0, per1 = (25 + G1 - 1) / G1;
    const int nst = 3 * per0 + 3 * per1;
    auto stage_desc = [&](int st, const float *&src, int &planes, int &HW) {
        if (st < 3 * per0) {
            const int a = st / per0, g = st - a * per0;
            HW = HW0; planes = min(G0, 25 - g * G0);
            src = h0 + ((size_t)(b * 3 + a) * 25 + g * G0) * HW0;
        } else {
            st -= 3 * per0;
            const int a = st / per1, g = st - a * per1;
            HW = HW1; planes = min(G1, 25 - g * G1);
            src = h1 + ((size_t)(b * 3 + a) * 25 + g * G1) * HW1;
        }
    };
    auto issue = [&](int st) {
        const float *src; int planes, HW;
        stage_desc(st, src, planes, HW);
        const uintptr_t lo = (uintptr_t)src & ~(uintptr_t)15, hi = ((uintptr_t)(src + (size_t)planes * HW) + 15) & ~(uintptr_t)15;
        const uint32_t full = saddr(&bars[st % S]);
        mbar_expect_tx(full, (uint32_t)(hi - lo));
        bulk_g2s(saddr(sm + (st % S) * SLOT), (const void *)lo, (uint32_t)(hi - lo), full);
    };
    float acc = 0.f, m = -1e30f;
    if (tid == 0) for (int st = 0; st < min(S - 1, nst); ++st) issue(st);
    for (int st = 0; st < nst; ++st) {
        const int slot = st % S, use = st / S;
        if (tid == 0) {
            const int nx = st + S - 1;
            if (nx < nst) {
                const int nslot = nx % S, nuse = nx / S;
                if (nuse > 0) mbar_wait(saddr(&bars[S + nslot]), (uint32_t)((nuse - 1) & 1));
                issue(nx);
            }
        }
        const float *src; int planes, HW;
        stage_desc(st, src, planes, HW);
        mbar_wait(saddr(&bars[slot]), (uint32_t)(use & 1));
        const float *stg = reinterpret_cast<const float *>(sm + slot * SLOT + ((uintptr_t)src & 15));
        if (tid < HW) {
            for (int u = 0; u < planes; ++u) m = fmaxf(m, stg[u * HW + tid]);
        }
        acc += m;
        __syncwarp();
        if (lane == 0) mbar_arrive(saddr(&bars[S + slot]));
    }
    if (acc == 12345.678f) out[b * blockDim.x + tid] = acc;
}

int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 256, HW0 = 121, HW1 = 484, R = 9;
    std::vector<float *> h0(R), h1(R);
    for (int r = 0; r < R; ++r) {
        CK(cudaMalloc(&h0[r], (size_t)N * 75 * HW0 * 4 + 64));
        CK(cudaMalloc(&h1[r], (size_t)N * 75 * HW1 * 4));
        CK(cudaMemset(h0[r], 0, (size_t)N * 75 * HW0 * 4));
        CK(cudaMemset(h1[r], 0, (size_t)N * 75 * HW1 * 4));
    }
    float *out;
    CK(cudaMalloc(&out, (size_t)N * 1024 * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)N * 75 * (HW0 + HW1) * 4;
    auto run = [&](const char *name, auto kern, int threads, int smem) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int i = 0; i < 5; ++i) kern<<<N, threads, smem>>>(h0[i % R], h1[i % R], HW0, HW1, out);
        CK(cudaDeviceSynchronize());
        const int steps = 90;
        cudaEventRecord(e0);
        for (int i = 0; i < steps; ++i) kern<<<N, threads, smem>>>(h0[i % R], h1[i % R], HW0, HW1, out);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-34s threads %4d smem %6d: %7.2f us/launch  %7.1f GB/s\n", name, threads, smem, 1e3 * ms / steps, bytes / (ms / steps * 1e-3) / 1e9);
    };
    const int half = 113 * 1024;
    run("cell .cs  (2 CTA/SM)", k_cell<0>, 512, half);
    run("cell .nc  (2 CTA/SM)", k_cell<1>, 512, half);
    run("cell plain(2 CTA/SM)", k_cell<2>, 512, half);
    run("cell .cs  (no smem limit)", k_cell<0>, 512, 1024);
    for (int kb : {16, 32, 48, 64, 72, 80, 88, 96, 104}) {
        char nm[64];
        snprintf(nm, sizeof nm, "cell .cs smem %d KB", kb);
        run(nm, k_cell<0>, 512, kb * 1024);
    }
    run("cell .cs  256 thr", k_cell<0>, 256, 1024);
    run("tma ring G=5 S=3 (2 CTA/SM)", k_tma<5, 3>, 512, half);
    run("tma ring G=5 S=4 (2 CTA/SM)", k_tma<5, 4>, 512, half);
    run("tma ring G=5 S=6 (2 CTA/SM)", k_tma<5, 6>, 512, half);
    run("tma ring G=25 S=2 (2 CTA/SM)", k_tma<25, 2>, 512, half);
    run("planes S=3 slot 12.2K (2 CTA/SM, 113K)", k_planes<3, 12544>, 512, half);
    run("planes S=4 slot 12.2K (2 CTA/SM, 113K)", k_planes<4, 12544>, 512, half);
    run("planes S=3 slot 12.2K (smem 40K)", k_planes<3, 12544>, 512, 40 * 1024);
    run("planes S=6 slot 12.2K (2 CTA/SM, 113K)", k_planes<6, 12544>, 512, half);
    run("planes S=2 slot 24.3K (2 CTA/SM, 113K)", k_planes<2, 24832>, 512, half);
    run("planes S=3 slot 24.3K (2 CTA/SM, 113K)", k_planes<3, 24832>, 512, half);
    run("planes S=8 slot 6.2K (2 CTA/SM, 113K)", k_planes<8, 6400>, 512, half);
    run("flat float4 (2 CTA/SM)", k_flat, 512, half);
    run("flat float4 x8 (2 CTA/SM)", k_flat8, 512, half);
    run("flat float4 x8 (no smem limit)", k_flat8, 512, 1024);
    return 0;
}
