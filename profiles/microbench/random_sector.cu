// Random-sector HBM microbenchmark: what DRAM rate does B200 sustain for scattered 32-byte (and 64/128-byte)
// accesses over a multi-GB arena?  That, not the streaming-copy peak, bounds the hypersparse pivot updates of path 3
// (each basis change touches ~300 scattered 32-byte items of a 1 MB tableau, 8.6 GB of tableaux per GPU).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o random_sector random_sector.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// each thread: `iters` rounds of U independent random accesses of GRAN bytes (read, optionally read+write)
template <int GRAN, int U, bool RW>
__global__ void rnd(double2 *buf, uint64_t nsect, int iters, double *sink)
{
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t s = mix(tid + 12345);
    double acc = 0;
    constexpr int V = GRAN / 16;
    for (int it = 0; it < iters; ++it) {
        double2 v[U][V];
        uint64_t off[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 6364136223846793005ULL + 1442695040888963407ULL; s ^= s >> 29;
            off[u] = (s & (nsect - 1)) * V;      // nsect is a power of two
#pragma unroll
            for (int x = 0; x < V; ++x) v[u][x] = __ldcg(buf + off[u] + x);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int x = 0; x < V; ++x) {
                acc += v[u][x].x;
                if (RW) { v[u][x].x += 1.0; __stcg(buf + off[u] + x, v[u][x]); }
            }
        }
    }
    if (acc == 1.2345) *sink = acc;
}

// the solver's pattern: a warp picks a random 1 MB tableau, 13 random rows (4 KB each) and 24 random 32-byte items
// per row (the same item set for every row), reads and rewrites them; lanes = items, U rows in flight per lane
template <int U>
__global__ void pivot_like(double2 *buf, uint64_t ntab, int iters, double *sink)
{
    const uint64_t wid = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    uint64_t s = mix(wid + 777);
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint64_t tab = ((s >> 20) % ntab) * (1u << 16);            // double2 index of the tableau (1 MB = 65536 double2)
        uint64_t sl = mix(s + lane);
        const unsigned item = (unsigned)(sl & 127);                      // my item of the row
        for (int r0 = 0; r0 < 16; r0 += U) {
            double2 v0[U], v1[U];
            uint64_t off[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned row = (unsigned)((s >> (8 + 3 * (r0 + u))) & 255);
                off[u] = tab + (uint64_t)row * 256 + 2 * item;
                if (lane < 24) { v0[u] = __ldcg(buf + off[u]); v1[u] = __ldcg(buf + off[u] + 1); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (lane < 24) {
                    v0[u].x += 1.0; v1[u].y += 1.0; acc += v0[u].y;
                    __stcg(buf + off[u], v0[u]); __stcg(buf + off[u] + 1, v1[u]);
                }
            }
        }
    }
    if (acc == 1.2345) *sink = acc;
}
template <int U>
void run_pivot_like(double2 *buf, size_t bytes, double *sink, int warps_per_sm, int sms = 148)
{
    const uint64_t ntab = bytes >> 20;
    const int threads = 128, blocks = sms * warps_per_sm / 4, iters = 4096 / warps_per_sm;
    if (sms != 148) printf("[%3d SMs] ", sms);
    pivot_like<U><<<blocks, threads>>>(buf, ntab, 2, sink);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    pivot_like<U><<<blocks, threads>>>(buf, ntab, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double piv = (double)blocks * (threads / 32) * iters;
    const double gb = piv * 16 * 24 * 32 * 2 / 1e9;
    printf("pivot-like RMW 16 rows x 24 items, %d rows in flight, %2d warps/SM: arena %.1f GB  %.3f ms  %.1f M pivots/s  %.1f GB/s (read+write bytes)\n",
           U, warps_per_sm, bytes / 1e9, ms, piv / ms / 1e3, gb / ms * 1e3);
}

template <int GRAN, int U, bool RW>
void run(double2 *buf, size_t bytes, double *sink, const char *name, int bps = 16)
{
    const uint64_t nsect = bytes / GRAN;
    const int blocks = 148 * bps, threads = 128, iters = 256 * 16 / bps;
    rnd<GRAN, U, RW><<<blocks, threads>>>(buf, nsect, 4, sink);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    rnd<GRAN, U, RW><<<blocks, threads>>>(buf, nsect, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double acc = (double)blocks * threads * iters * U;
    const double gb = acc * GRAN * (RW ? 2 : 1) / 1e9;
    printf("%-28s arena %.1f GB  %.3f ms  %.1f G accesses/s  %.1f GB/s (%s)\n", name, bytes / 1e9, ms, acc / ms / 1e6, gb / ms * 1e3,
           RW ? "read+write bytes" : "read bytes");
}

int main()
{
    {
        // Little's law check: an L2-resident arena (64 MiB) and 1 / 4 / 16 CTAs of 128 threads per SM
        const size_t bytes = size_t(64) << 20;
        double2 *buf; double *sink;
        cudaMalloc(&buf, bytes); cudaMalloc(&sink, 8);
        cudaMemset(buf, 0, bytes);
        run<32, 8, false>(buf, bytes, sink, "L2: 32 B read, 16 CTA/SM", 16);
        run<32, 8, false>(buf, bytes, sink, "L2: 32 B read, 4 CTA/SM", 4);
        run<32, 8, false>(buf, bytes, sink, "L2: 32 B read, 1 CTA/SM", 1);
        run<32, 8, true>(buf, bytes, sink, "L2: 32 B r+w, 16 CTA/SM", 16);
        cudaFree(buf); cudaFree(sink);
    }
    for (size_t gbs : {size_t(8)}) {
        const size_t bytes = gbs << 30;
        double2 *buf; double *sink;
        cudaMalloc(&buf, bytes); cudaMalloc(&sink, 8);
        cudaMemset(buf, 0, bytes);
        for (int w : {4, 8, 16, 32, 64}) run_pivot_like<8>(buf, bytes, sink, w);
        for (int sm : {18, 37, 74}) { run_pivot_like<8>(buf, bytes, sink, 4, sm); run_pivot_like<8>(buf, bytes, sink, 16, sm); }
        run<32, 8, false>(buf, bytes, sink, "32 B read, 8 in flight");
        run<32, 8, false>(buf, bytes, sink, "32 B read, 4 CTA/SM", 4);
        run<32, 8, false>(buf, bytes, sink, "32 B read, 1 CTA/SM", 1);
        run<32, 16, false>(buf, bytes, sink, "32 B read, 16 in flight");
        run<64, 8, false>(buf, bytes, sink, "64 B read, 8 in flight");
        run<128, 4, false>(buf, bytes, sink, "128 B read, 4 in flight");
        run<32, 8, true>(buf, bytes, sink, "32 B read+write, 8 in flight");
        run<64, 8, true>(buf, bytes, sink, "64 B read+write, 8 in flight");
        run<128, 4, true>(buf, bytes, sink, "128 B read+write, 4 in flight");
        cudaFree(buf); cudaFree(sink);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
