// cuda_utils.h -- the helpers of /root/reference/cuda_utils.h that main.cpp uses (CHECK, CheckMsg,
// initDevice, cpuTimer, GpuTimer, iAlignUp, iDivUp; main.cpp:16,93,135,208,210,212), written for
// this library. Same names and behaviour: errors print and exit(-1) (cuda_utils.h:18-37).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>

#define H_PI 1.5707963267948966f
#define CHECK(err) ::surf_compat::check_((err), __FILE__, __LINE__)
#define CheckMsg(msg) ::surf_compat::check_msg_((msg), __FILE__, __LINE__)

namespace surf_compat {
inline void check_(cudaError_t err, const char* file, int line) {
    if (err == cudaSuccess) return;
    std::fprintf(stderr, "CHECK() Runtime API error in file <%s>, line %i : %s.\n", file, line, cudaGetErrorString(err));
    std::exit(-1);
}
inline void check_msg_(const char* msg, const char* file, int line) {
    const cudaError_t err = cudaGetLastError();
    if (err == cudaSuccess) return;
    std::fprintf(stderr, "CheckMsg() CUDA error: %s in file <%s>, line %i : %s.\n", msg, file, line, cudaGetErrorString(err));
    std::exit(-1);
}
}  // namespace surf_compat

// Select device `dev` (clamped to the available range) and print what was chosen.
inline bool initDevice(int dev) {
    int count = 0;
    CHECK(cudaGetDeviceCount(&count));
    if (count == 0) { std::fprintf(stderr, "CUDA error: no devices supporting CUDA.\n"); return false; }
    dev = std::max(0, std::min(dev, count - 1));
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, dev));
    CHECK(cudaSetDevice(dev));
    int drv = 0, rt = 0;
    CHECK(cudaDriverGetVersion(&drv));
    CHECK(cudaRuntimeGetVersion(&rt));
    std::fprintf(stderr, "Using Device %d: %s, CUDA Driver Version: %d.%d, Runtime Version: %d.%d\n", dev, prop.name,
                 drv / 1000, drv % 1000, rt / 1000, rt % 1000);
    return true;
}

// wall clock in microseconds
inline long long cpuTimer() {
    return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}

// cudaEvent stopwatch started at construction; read() returns elapsed milliseconds
class GpuTimer {
public:
    explicit GpuTimer(cudaStream_t s = 0) : stream_(s) {
        cudaEventCreate(&t0_);
        cudaEventCreate(&t1_);
        cudaEventRecord(t0_, stream_);
    }
    ~GpuTimer() { cudaEventDestroy(t0_); cudaEventDestroy(t1_); }
    float read() {
        cudaEventRecord(t1_, stream_);
        cudaEventSynchronize(t1_);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0_, t1_);
        return ms;
    }
private:
    cudaEvent_t t0_, t1_;
    cudaStream_t stream_;
};

inline int iAlignUp(int a, int b) { return (a % b) ? a - a % b + b : a; }
inline int iDivUp(int a, int b) { return (a + b - 1) / b; }
