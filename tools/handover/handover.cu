// Micro-benchmark (tuning evidence, not part of the library): how long after an H2D copy ends does a
// consumer on another stream start, (a) through cudaEventRecord + cudaStreamWaitEvent, (b) through a
// stream memory operation (cuStreamWriteValue32 behind the copy) observed by a kernel that is already
// resident and polls the word? Build: nvcc -arch=sm_100a -o handover handover.cu -ldl ; run on a GPU box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));       \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// consumer: (poll the flag until it holds `want`, at most ~5 ms), then stamp
__global__ void k_consume(const volatile unsigned* flag, unsigned want, int poll, unsigned long long* stamp, const volatile float* data, float* sink) {
  if (poll) {
    const unsigned long long t0 = gtimer();
    while ((int) (*flag - want) < 0 && gtimer() - t0 < 5000000ull)
      __nanosleep(64);
  }
  stamp[0] = gtimer();
  sink[0]  = data[0]; // touch what was uploaded
}
__global__ void k_stamp(unsigned long long* stamp) {
  stamp[0] = gtimer();
}

typedef CUresult (*write32_t)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

int main() {
  CK(cudaSetDevice(0));
  void* h = dlopen("libcuda.so.1", RTLD_NOW);
  write32_t write32 = h ? (write32_t) dlsym(h, "cuStreamWriteValue32_v2") : nullptr;
  if (!write32 && h)
    write32 = (write32_t) dlsym(h, "cuStreamWriteValue32");
  printf("cuStreamWriteValue32: %s\n", write32 ? "found" : "missing");
  const size_t bytes = 1228800; // one 640x480 depth image
  float *hp, *dp, *sink;
  CK(cudaMallocHost(&hp, bytes));
  CK(cudaMalloc(&dp, bytes));
  CK(cudaMalloc(&sink, 64));
  unsigned* flag;
  CK(cudaMalloc(&flag, 64));
  CK(cudaMemset(flag, 0, 64));
  unsigned long long *st_copy_end, *st_cons;
  CK(cudaMalloc(&st_copy_end, 8));
  CK(cudaMalloc(&st_cons, 8));
  cudaStream_t sc, sk;
  CK(cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking));
  cudaEvent_t ev;
  CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  const int reps = 60;
  for (int mode = 0; mode < (write32 ? 2 : 1); ++mode) {
    double sum = 0, mn = 1e30, mx = 0;
    for (int r = 0; r < reps; ++r) {
      const unsigned want = (unsigned) (mode * 1000 + r + 1);
      if (mode == 1) // the consumer is launched first and waits on the device
        k_consume<<<1, 1, 0, sk>>>(flag, want, 1, st_cons, dp, sink);
      CK(cudaMemcpyAsync(dp, hp, bytes, cudaMemcpyHostToDevice, sc));
      k_stamp<<<1, 1, 0, sc>>>(st_copy_end); // (device clock right behind the copy; costs a launch on the copy stream)
      if (mode == 0) {
        CK(cudaEventRecord(ev, sc));
        CK(cudaStreamWaitEvent(sk, ev, 0));
        k_consume<<<1, 1, 0, sk>>>(flag, want, 0, st_cons, dp, sink);
      } else {
        if (write32(sc, (CUdeviceptr) flag, want, 0) != CUDA_SUCCESS) {
          fprintf(stderr, "cuStreamWriteValue32 failed\n");
          return 1;
        }
      }
      CK(cudaDeviceSynchronize());
      unsigned long long a, b;
      CK(cudaMemcpy(&a, st_copy_end, 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&b, st_cons, 8, cudaMemcpyDeviceToHost));
      const double us = ((double) b - (double) a) / 1e3;
      if (r >= 10) {
        sum += us;
        mn = us < mn ? us : mn, mx = us > mx ? us : mx;
      }
    }
    printf("%s: consumer starts %.1f us after the stamp kernel behind the copy (min %.1f, max %.1f)\n", mode == 0 ? "event + cudaStreamWaitEvent            " : "cuStreamWriteValue32 + resident consumer", sum / (reps - 10), mn, mx);
  }
  // reference: how long does the copy itself take, and a back-to-back kernel pair on one stream
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, sc));
  for (int r = 0; r < 50; ++r)
    CK(cudaMemcpyAsync(dp, hp, bytes, cudaMemcpyHostToDevice, sc));
  CK(cudaEventRecord(e1, sc));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("copy of %zu bytes: %.1f us\n", bytes, ms * 1e3 / 50);
  return 0;
}
