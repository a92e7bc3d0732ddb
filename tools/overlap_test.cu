// dev probe: do host<->device copies overlap kernels on this box, also from cudaHostRegister'ed memory?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <chrono>
__global__ void spin(unsigned long long cycles, unsigned *out) { unsigned long long t0 = clock64(); while (clock64() - t0 < cycles) {} if (out) out[0] = 1; }
int main() {
    size_t n = 1ull << 30; void *h, *h2, *d; unsigned *o; void *small_h, *small_d;
    cudaMallocHost(&h, n); cudaMalloc(&d, n); cudaMalloc(&o, 4);
    h2 = aligned_alloc(4096, n); for (size_t i = 0; i < n; i += 4096) ((char *)h2)[i] = 1;
    printf("register %d\n", (int)cudaHostRegister(h2, n, cudaHostRegisterDefault));
    cudaMallocHost(&small_h, 1 << 20); cudaMalloc(&small_d, 1 << 20);
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    for (int rep = 0; rep < 2; rep++) {
        auto t0 = now(); cudaMemcpyAsync(d, h2, n, cudaMemcpyHostToDevice, s1); auto t0b = now(); cudaStreamSynchronize(s1); auto t1 = now();
        spin<<<148, 256, 0, s2>>>(40000000ull, o); cudaStreamSynchronize(s2); auto t2 = now();
        cudaMemcpyAsync(d, h2, n, cudaMemcpyHostToDevice, s1); spin<<<148, 256, 0, s2>>>(40000000ull, o); cudaDeviceSynchronize(); auto t3 = now();
        // pattern of the engine: small H2D on s2, kernel on s2, then big copy on s1 enqueued after
        cudaMemcpyAsync(small_d, small_h, 128 << 10, cudaMemcpyHostToDevice, s2); spin<<<148, 256, 0, s2>>>(40000000ull, o);
        cudaMemcpyAsync(small_h, small_d, 4096, cudaMemcpyDeviceToHost, s2);
        cudaMemcpyAsync(d, h2, n, cudaMemcpyHostToDevice, s1); cudaDeviceSynchronize(); auto t4 = now();
        printf("registered: copy %.2f ms (enqueue %.3f), kernel %.2f ms, both %.2f ms, engine pattern %.2f ms\n", ms(t0, t1), ms(t0, t0b), ms(t1, t2), ms(t2, t3), ms(t3, t4));
    }
    return 0;
}
