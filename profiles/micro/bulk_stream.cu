// Microbenchmark: one CTA per "image" streams a contiguous 181,504-byte region from HBM (cold) with
//   (a) cp.async.bulk (TMA engine) into a shared-memory ring: NC copies of CB bytes per stage, S stages, or
//   (b) plain 16-byte loads with U loads in flight per thread.
// Prints us/launch, GB/s and, for CTA 0, the arrival time of every stage (SM clock, cycles after CTA start).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_stream.bin bulk_stream.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int kRegion = 181504;  // bytes per CTA (multiple of 128)

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

// mode 0: every stage is consumed (read) by all threads; producer = warp 0 (lane l issues copy l, l+32, ..)
template <int NC, int CB, int S>
__global__ void __launch_bounds__(512, 2) k_bulk(const unsigned char *src, float *out, long long *stamps) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int SB = NC * CB;  // stage bytes
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + S * SB);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long t0 = clock64();
    if (tid == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(saddr(&bars[i]), 1); mbar_init(saddr(&bars[S + i]), 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned char *base = src + (size_t)b * kRegion;
    const int nst = (kRegion + SB - 1) / SB;
    auto issue = [&](int st) {
        const int off = st * SB, bytes = min(SB, kRegion - off);
        const uint32_t full = saddr(&bars[st % S]);
        if (lane == 0) mbar_expect_tx(full, (uint32_t)bytes);
        __syncwarp();
        for (int c = lane; c * CB < bytes; c += 32) {
            const int n = min(CB, bytes - c * CB);
            bulk_g2s(saddr(sm + (st % S) * SB + c * CB), base + off + c * CB, (uint32_t)n, full);
        }
    };
    float acc = 0.f;
    if (warp == 0) for (int st = 0; st < min(S - 1, nst); ++st) issue(st);
    for (int st = 0; st < nst; ++st) {
        const int slot = st % S, use = st / S;
        if (warp == 0) {
            const int nx = st + S - 1;
            if (nx < nst) {
                if (nx / S > 0) mbar_wait(saddr(&bars[S + nx % S]), (uint32_t)((nx / S - 1) & 1));
                issue(nx);
            }
        }
        mbar_wait(saddr(&bars[slot]), (uint32_t)(use & 1));
        if (b == 0 && tid == 0 && stamps) stamps[st] = clock64() - t0;
        const float *stg = reinterpret_cast<const float *>(sm + slot * SB);
        for (int i = tid; i < SB / 4; i += 512) acc += stg[i];
        __syncwarp();
        if (lane == 0) mbar_arrive(saddr(&bars[S + slot]));
    }
    if (acc == 12345.678f) out[b * 512 + tid] = acc;
}

// mode 1: no ring reuse at all: the whole region in NC*... copies issued up front into 176 KB of smem (1 CTA/SM)
template <int CB>
__global__ void __launch_bounds__(512, 1) k_bulk_all(const unsigned char *src, float *out, long long *stamps) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm + kRegion);
    const int b = blockIdx.x, tid = threadIdx.x;
    const long long t0 = clock64();
    if (tid == 0) { mbar_init(saddr(&bars[0]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const unsigned char *base = src + (size_t)b * kRegion;
    if (tid == 0) mbar_expect_tx(saddr(&bars[0]), kRegion);
    __syncthreads();
    for (int c = tid; c * CB < kRegion; c += 512) bulk_g2s(saddr(sm + c * CB), base + c * CB, (uint32_t)min(CB, kRegion - c * CB), saddr(&bars[0]));
    mbar_wait(saddr(&bars[0]), 0);
    if (b == 0 && tid == 0 && stamps) stamps[0] = clock64() - t0;
    float acc = 0.f;
    const float *stg = reinterpret_cast<const float *>(sm);
    for (int i = tid; i < kRegion / 4; i += 512) acc += stg[i];
    if (acc == 12345.678f) out[b * 512 + tid] = acc;
}

// plain loads: U 16-byte loads in flight per thread
template <int U, int POLICY>
__global__ void __launch_bounds__(512, 2) k_ldg(const unsigned char *src, float *out, long long *stamps) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float4 *p = reinterpret_cast<const float4 *>(src + (size_t)b * kRegion);
    constexpr int n4 = kRegion / 16;  // 11344
    float acc = 0.f;
    for (int i0 = 0; i0 < n4; i0 += U * 512) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * 512 + tid;
            if (i < n4) {
                if (POLICY == 0) v[u] = __ldcs(p + i);
                else asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p + i));
            } else v[u] = make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 12345.678f) out[b * 512 + tid] = acc;
    if (tid == 0) sm[0] = (unsigned char)acc;
}

// LDGSTS (cp.async 16 B) into smem: all of a 45 KB chunk in flight, 4 chunks
__global__ void __launch_bounds__(512, 2) k_cpasync(const unsigned char *src, float *out, long long *stamps) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const unsigned char *base = src + (size_t)b * kRegion;
    constexpr int CH = kRegion / 4;  // 45376 B per chunk, 2 buffers
    float acc = 0.f;
    auto issue = [&](int ch) {
        for (int i = tid * 16; i < CH; i += 512 * 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr(sm + (ch & 1) * CH + i)), "l"(base + ch * CH + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0);
    for (int ch = 0; ch < 4; ++ch) {
        if (ch + 1 < 4) { issue(ch + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float *stg = reinterpret_cast<const float *>(sm + (ch & 1) * CH);
        for (int i = tid; i < CH / 4; i += 512) acc += stg[i];
        __syncthreads();
    }
    if (acc == 12345.678f) out[b * 512 + tid] = acc;
}

int main(int argc, char **argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 256, R = 9;
    std::vector<unsigned char *> buf(R);
    for (int r = 0; r < R; ++r) { CK(cudaMalloc(&buf[r], (size_t)N * kRegion + 256)); CK(cudaMemset(buf[r], 0, (size_t)N * kRegion + 256)); }
    float *out; CK(cudaMalloc(&out, (size_t)N * 512 * 4));
    long long *stamps; CK(cudaMalloc(&stamps, 4096 * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)N * kRegion;
    auto run = [&](const char *name, auto kern, int smem, int nstamps) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaMemset(stamps, 0, 4096 * 8));
        for (int i = 0; i < 5; ++i) kern<<<N, 512, smem>>>(buf[i % R], out, stamps);
        CK(cudaDeviceSynchronize());
        const int steps = 90;
        cudaEventRecord(e0);
        for (int i = 0; i < steps; ++i) kern<<<N, 512, smem>>>(buf[i % R], out, stamps);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s smem %6d: %7.2f us/launch %7.1f GB/s", name, smem, 1e3 * ms / steps, bytes / (ms / steps * 1e-3) / 1e9);
        if (nstamps) {
            std::vector<long long> h(nstamps);
            CK(cudaMemcpy(h.data(), stamps, nstamps * 8, cudaMemcpyDeviceToHost));
            printf("  | stage arrivals (kcyc):");
            for (int i = 0; i < nstamps && i < 24; ++i) printf(" %.1f", h[i] / 1e3);
        }
        printf("\n");
    };
    const int half = 113 * 1024;
#define BULK(NC, CB, S) run("bulk NC=" #NC " CB=" #CB " S=" #S, k_bulk<NC, CB, S>, half, (kRegion + NC * CB - 1) / (NC * CB))
    BULK(1, 12288, 3);
    BULK(1, 12288, 6);
    BULK(4, 3072, 3);
    BULK(4, 3072, 6);
    BULK(12, 1024, 3);
    BULK(12, 1024, 6);
    BULK(32, 384, 3);
    BULK(32, 1024, 3);
    BULK(32, 1024, 2);
    BULK(1, 32768, 3);
    BULK(2, 16384, 3);
    BULK(8, 4096, 3);
    BULK(16, 2048, 3);
    BULK(1, 4096, 16);
    BULK(1, 2048, 32);
    run("bulk all-up-front CB=4096 (1 CTA/SM)", k_bulk_all<4096>, kRegion + 64, 1);
    run("bulk all-up-front CB=1024 (1 CTA/SM)", k_bulk_all<1024>, kRegion + 64, 1);
    run("bulk all-up-front CB=16384 (1 CTA/SM)", k_bulk_all<16384>, kRegion + 64, 1);
    run("ldg.128 .cs U=4 (113 KB smem)", k_ldg<4, 0>, half, 0);
    run("ldg.128 .cs U=8 (113 KB smem)", k_ldg<8, 0>, half, 0);
    run("ldg.128 .cs U=8 (1 KB smem)", k_ldg<8, 0>, 1024, 0);
    run("ldg.128 .cs U=4 (1 KB smem)", k_ldg<4, 0>, 1024, 0);
    run("ldg.128 .cs U=12 (1 KB smem)", k_ldg<12, 0>, 1024, 0);
    run("ldg.128 nc.no_allocate U=8 (113 KB smem)", k_ldg<8, 1>, half, 0);
    run("ldg.128 nc.no_allocate U=8 (1 KB smem)", k_ldg<8, 1>, 1024, 0);
    run("cp.async 16B double buffer 2x45KB (113 KB)", k_cpasync, half, 0);
    return 0;
}
