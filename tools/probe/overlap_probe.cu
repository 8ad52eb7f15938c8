// overlap_probe.cu -- do two kernels launched on two non-blocking streams share the machine?  (round 2: the engine's
// attempts at running kernels side by side all measured as if they ran one after the other)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void spinStatic(long long cycles, int *sink) {
    __shared__ int pad[4000];
    long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (threadIdx.x == 0 && pad[threadIdx.x & 1] == 12345) *sink = 1;
}
__global__ void __launch_bounds__(128, 5) spinDynamic(long long cycles, int *sink) {
    extern __shared__ int dyn[];
    long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (threadIdx.x == 0 && dyn[threadIdx.x & 1] == 12345) *sink = 1;
}

int main() {
    int *sink; CK(cudaMalloc(&sink, 4));
    cudaStream_t a, b; CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, f; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreateWithFlags(&f, cudaEventDisableTiming));
    CK(cudaFuncSetAttribute(spinDynamic, cudaFuncAttributeMaxDynamicSharedMemorySize, 26624));
    const long long cyc = 2000000;   // about 1 ms
    int sms = 148;
    auto run = [&](const char *what, int gridA, int gridB, bool dynA, bool dynB, bool oneStream) -> int {
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, a));
            CK(cudaEventRecord(f, a)); CK(cudaStreamWaitEvent(b, f, 0));
            cudaStream_t sb = oneStream ? a : b;
            if (dynA) spinDynamic<<<gridA, 128, 26624, a>>>(cyc, sink); else spinStatic<<<gridA, 256, 0, a>>>(cyc, sink);
            if (dynB) spinDynamic<<<gridB, 128, 26624, sb>>>(cyc, sink); else spinStatic<<<gridB, 256, 0, sb>>>(cyc, sink);
            if (!oneStream) { CK(cudaEventRecord(f, b)); CK(cudaStreamWaitEvent(a, f, 0)); }
            CK(cudaEventRecord(e1, a));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep == 2) printf("%-70s %.3f ms\n", what, ms);
        }
        return 0;
    };
    run("one stream, static + static, 1 CTA/SM each (serial reference)", sms, sms, false, false, true);
    run("two streams, static + static, 1 CTA/SM each", sms, sms, false, false, false);
    run("two streams, dynamic + dynamic, 2 CTAs/SM each", 2 * sms, 2 * sms, true, true, false);
    run("two streams, dynamic(2/SM) + static(1/SM)", 2 * sms, sms, true, false, false);
    run("two streams, static(1/SM) + dynamic(2/SM)", sms, 2 * sms, false, true, false);
    run("two streams, dynamic(5/SM = full) + dynamic(5/SM)", 5 * sms, 5 * sms, true, true, false);
    run("two streams, dynamic(3/SM) + dynamic(2/SM)", 3 * sms, 2 * sms, true, true, false);
    for (int c : {100, 50}) {
        CK(cudaFuncSetAttribute(spinDynamic, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        CK(cudaFuncSetAttribute(spinStatic, cudaFuncAttributePreferredSharedMemoryCarveout, c));
        printf("carveout %d for both:\n", c);
        run("  two streams, dynamic(2/SM) + static(1/SM)", 2 * sms, sms, true, false, false);
        run("  two streams, static(2/SM) + dynamic(3/SM)", 2 * sms, 3 * sms, false, true, false);
    }
    return 0;
}
