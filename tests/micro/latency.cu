// micro-benchmark (not a test): dependent global-load chain latency on B200, one warp, L2-resident data,
// (a) inside one allocation, (b) hopping across 96 separate cudaMalloc allocations (TLB pressure).
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void chase(const int *const *bufs, int nbuf, int hops, int *out, long long *cycles) {
    int idx = 0, b = 0;
    long long t0 = clock64();
    for (int i = 0; i < hops; i++) { idx = bufs[b][idx & 1023]; b = (b + 1 == nbuf) ? 0 : b + 1; }
    long long t1 = clock64();
    if (threadIdx.x == 0) { *out = idx; *cycles = t1 - t0; }
}
__global__ void chase1(const int *buf, int hops, int *out, long long *cycles) {
    int idx = 0;
    long long t0 = clock64();
    for (int i = 0; i < hops; i++) idx = buf[idx];
    long long t1 = clock64();
    if (threadIdx.x == 0) { *out = idx; *cycles = t1 - t0; }
}
int main() {
    const int NB = 96, N = 1024, HOPS = 2000;
    std::vector<int*> d(NB); std::vector<int> h(N);
    for (int i = 0; i < N; i++) h[i] = (i * 37 + 11) % N;
    for (int b = 0; b < NB; b++) { cudaMalloc(&d[b], (b % 3 == 0 ? 8 : 1) << 20); cudaMemcpy(d[b], h.data(), N * 4, cudaMemcpyHostToDevice); }
    const int **dp; cudaMalloc(&dp, NB * sizeof(int*)); cudaMemcpy(dp, d.data(), NB * sizeof(int*), cudaMemcpyHostToDevice);
    int *out; long long *cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    long long c;
    for (int rep = 0; rep < 3; rep++) {
        chase1<<<1, 32>>>(d[0], HOPS, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("one buffer      : %.1f cycles/hop\n", (double)c / HOPS);
        for (int nb : {1, 8, 32, 96}) {
            chase<<<1, 32>>>(dp, nb, HOPS, out, cyc); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("hopping %2d bufs : %.1f cycles/hop (2 dependent loads per hop: pointer + data)\n", nb, (double)c / HOPS);
        }
    }
    // launch + event overhead of an empty-ish kernel in a stream
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); for (int i = 0; i < 200; i++) chase1<<<1, 32>>>(d[0], 1, out, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); printf("back-to-back tiny kernels: %.2f us each\n", ms * 1000 / 200);
    return 0;
}
