// Optional per-launch timing (CUDA events on the launching stream), aggregated by kernel tag.
// Off by default; bench.py turns it on for one extra step to attribute the step time to kernels
// (there is no nsys in this image).  Also counts launches for bench.py's `gpu_launches`.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

void dcpt_prof_begin(const char* tag, double flops, double bytes, cudaStream_t st);
void dcpt_prof_end(cudaStream_t st);
extern bool g_dcpt_prof_on;
extern bool g_dcpt_prof_shapes;  // dcpt_prof_enable(2): GEMM tags carry MxNxK
const char* dcpt_prof_intern(const char* s);
extern long long g_dcpt_launches;

struct DcptProfScope {
  cudaStream_t st;
  bool on;
  DcptProfScope(const char* tag, double flops, double bytes, cudaStream_t s) : st(s), on(g_dcpt_prof_on) {
    ++g_dcpt_launches;
    if (on) dcpt_prof_begin(tag, flops, bytes, st);
  }
  ~DcptProfScope() {
    if (on) dcpt_prof_end(st);
  }
};
#define DCPT_PROF(tag, flops, bytes, st) DcptProfScope _prof_scope((tag), (double)(flops), (double)(bytes), (st))
// tag + " <a>x<b>" when per-shape tags are requested (dcpt_prof_enable(2)); a, b = rows / channels of the launch
inline const char* dcpt_prof_tag2(const char* tag, long long a, long long b) {
  if (!(g_dcpt_prof_on && g_dcpt_prof_shapes)) return tag;
  char buf[96];
  snprintf(buf, sizeof(buf), "%s %lldx%lld", tag, a, b);
  return dcpt_prof_intern(buf);
}
