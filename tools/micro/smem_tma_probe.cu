// Data-path probes behind the hit-count kernel's design (profiles/r2_datapath_probe.txt):
//   1. LDS.128 read rate of 16 warps                                   (the consumer side of a shared-memory staged design)
//   2. cp.async.bulk (TMA) of 512-byte chunks global -> shared memory  (the producer side: one chunk = one bit-row slice of a
//      4096-reference tile; rows are scattered, so every slice is its own bulk copy) issued by 1 lane or by 32 lanes, with and
//      without the LDS stream beside it
//   3. ld.global.nc.v2 of scattered, L2-resident 256-byte slices (one per warp and load, 32 in flight per warp, no arithmetic):
//      the ceiling of the path the kernel uses (L2 -> L1 fill -> register)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_tma_probe smem_tma_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 ldg64(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

constexpr int kConsumers = 16;
constexpr int kStages = 4;
constexpr int kStageBytes = 16384;  // 32 slices of 512 bytes

// mode bit 0: LDS consumers on; bit 1: TMA producer on; bit 2: all 32 producer lanes issue copies
__global__ void __launch_bounds__((kConsumers + 1) * 32, 1) probe(const unsigned char* __restrict__ src, size_t src_bytes, int iters, int mode, int chunk,
                                                                   unsigned long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned char* cons = sm;                      // kConsumers * 8 KB
    unsigned char* ring = sm + kConsumers * 8192;  // kStages * kStageBytes
    __shared__ uint64_t full[kStages];
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        stop = 0;
    }
    for (int i = tid; i < kConsumers * 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(cons)[i] = i * 2654435761u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const long long t0 = clock64();
    if (warp < kConsumers) {
        if (mode & 1) {
            const uint32_t base = smem_u32(cons + warp * 8192) + lane * 16;
            uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint4 v = lds128(base + k * 512);
                    acc0 ^= v.x, acc1 ^= v.y, acc2 ^= v.z, acc3 ^= v.w;
                }
            }
            const long long t1 = clock64();
            if (lane == 0) {
                out[(size_t)blockIdx.x * 64 + warp] = (unsigned long long)(t1 - t0);
                out[(size_t)blockIdx.x * 64 + 32 + warp] = acc0 ^ acc1 ^ acc2 ^ acc3;
            }
        }
        __syncwarp();
        if ((mode & 1) && warp == 0 && lane == 0) stop = 1;  // consumer 0 finished: the producer stops
    } else if (mode & 2) {
        const int n_issuers = (mode & 4) ? 32 : 1;
        const int per_stage = kStageBytes / chunk;
        unsigned long long bytes = 0;
        uint32_t phase[kStages] = {0, 0, 0, 0};
        bool armed[kStages] = {false, false, false, false};
        size_t off = (((size_t)blockIdx.x * 1315423911u) % (src_bytes / 2)) & ~(size_t)4095;
        int s = 0;
        const int max_stages = (mode & 1) ? (1 << 30) : iters;  // without consumers: a fixed number of stages
        for (int n = 0; n < max_stages && !stop; ++n) {
            if (armed[s]) {
                mbar_wait(&full[s], phase[s]);
                phase[s] ^= 1u;
            }
            if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)kStageBytes);
            __syncwarp();
            for (int c = lane; c < per_stage; c += 32) {
                if (lane < n_issuers || n_issuers == 32) {
                    for (int cc = c; cc < (n_issuers == 1 ? per_stage : c + 1); ++cc)
                        bulk_g2s(ring + (size_t)s * kStageBytes + (size_t)cc * chunk, src + ((off + (size_t)cc * (4096 + chunk)) % (src_bytes - chunk) & ~(size_t)15),
                                 (uint32_t)chunk, &full[s]);
                }
                if (n_issuers == 1) break;
            }
            off = (off + (size_t)per_stage * (4096 + chunk)) % (src_bytes / 2);
            armed[s] = true;
            bytes += kStageBytes;
            s = (s + 1) % kStages;
        }
        for (int q = 0; q < kStages; ++q)
            if (armed[q]) mbar_wait(&full[q], phase[q]);
        if (lane == 0) {
            out[(size_t)blockIdx.x * 64 + 63] = bytes;
            out[(size_t)blockIdx.x * 64 + 62] = (unsigned long long)(clock64() - t0);
        }
    }
}

// 16 warps per SM, every load of a warp = one scattered 256-byte slice (row stride 12 800 bytes as on the 100 k-reference index),
// 32 loads in flight per warp
__global__ void __launch_bounds__(512, 1) ldg_probe(const unsigned char* __restrict__ src, size_t src_bytes, int iters, unsigned long long* out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    (void)src_bytes;  // 4096 rows x 12 800 bytes = 52 MB used
    uint32_t acc = 0;
    uint32_t h = (blockIdx.x * 16 + warp) * 2654435761u + 12345u;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint2 v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            h = h * 1664525u + 1013904223u;
            const size_t row = (size_t)(h >> 8) & 4095u;
            v[k] = ldg64(src + row * 12800 + (size_t)((h >> 3) & 31) * 256 + lane * 8);
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) acc ^= v[k].x ^ v[k].y;
    }
    const long long t1 = clock64();
    if (lane == 0) {
        out[(size_t)blockIdx.x * 64 + warp] = (unsigned long long)(t1 - t0);
        out[(size_t)blockIdx.x * 64 + 32 + warp] = acc;
    }
}

int main() {
    const size_t src_bytes = 96ull << 20;  // L2-resident source (126 MB L2)
    unsigned char* src;
    cudaMalloc(&src, src_bytes);
    cudaMemset(src, 1, src_bytes);
    unsigned long long* out;
    cudaMalloc(&out, 148 * 64 * 8);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<unsigned long long> h(148 * 64);
    const size_t smem = kConsumers * 8192 + (size_t)kStages * kStageBytes;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    auto run = [&](int mode, int chunk, int iters) {
        cudaMemset(out, 0, 148 * 64 * 8);
        probe<<<sms, (kConsumers + 1) * 32, smem>>>(src, src_bytes, iters, mode, chunk, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(e));
            return;
        }
        cudaMemcpy(h.data(), out, 148 * 64 * 8, cudaMemcpyDeviceToHost);
        double cyc = 0, wb = 0, wc = 0;
        for (int b = 0; b < sms; ++b) {
            unsigned long long m = 0;
            for (int w = 0; w < kConsumers; ++w) m = m > h[b * 64 + w] ? m : h[b * 64 + w];
            cyc += (double)m;
            wb += (double)h[b * 64 + 63];
            wc += (double)h[b * 64 + 62];
        }
        cyc /= sms;
        const double read_bytes = (double)kConsumers * iters * 16 * 512;
        printf("LDS.128 consumers %s | TMA producer %s (%2d lanes issue, %4d-byte copies): LDS read %6.1f B/clk/SM   TMA write %6.1f B/clk/SM (%.0f clk per copy)\n",
               (mode & 1) ? "on " : "off", (mode & 2) ? "on " : "off", (mode & 4) ? 32 : 1, chunk, (mode & 1) ? read_bytes / cyc : 0.0, wc > 0 ? wb / wc : 0.0,
               wb > 0 ? wc / (wb / chunk) : 0.0);
    };
    for (int rep = 0; rep < 2; ++rep) {
        run(1, 512, 20000);
        run(2, 512, 4000);
        run(2 | 4, 512, 4000);
        run(2 | 4, 2048, 4000);
        run(1 | 2 | 4, 512, 20000);
        cudaMemset(out, 0, 148 * 64 * 8);
        const int iters = 2000;
        ldg_probe<<<sms, 512>>>(src, src_bytes, iters, out);
        cudaDeviceSynchronize();
        cudaMemcpy(h.data(), out, 148 * 64 * 8, cudaMemcpyDeviceToHost);
        double cyc = 0;
        for (int b = 0; b < sms; ++b) {
            unsigned long long m = 0;
            for (int w = 0; w < 16; ++w) m = m > h[b * 64 + w] ? m : h[b * 64 + w];
            cyc += (double)m;
        }
        cyc /= sms;
        printf("ld.global.nc.v2 of scattered L2-resident 256-byte slices, 16 warps x 32 loads in flight: %.1f B/clk/SM\n", 16.0 * iters * 32 * 256 / cyc);
    }
    return 0;
}
