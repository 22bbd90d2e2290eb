// Optional per-kernel-class timing with CUDA events on the launch stream (used by bench.py for the live roofline
// numbers) and an always-on launch counter (bench.py's gpu_launches).  Profiling is OFF by default: no events, no syncs.
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace mmnn {

enum ProfClass {
  PC_PACK = 0, PC_S2D, PC_STEM_FPROP, PC_MAXPOOL, PC_CONV1_FPROP, PC_CONV2_FPROP, PC_TRANS_POOL, PC_TRANS_FPROP, PC_NORM5,
  PC_BN_RUNNING, PC_NORM5_BWD, PC_EXTRACT, PC_CONV2_WGRAD, PC_CONV2_DGRAD, PC_BN_APPLY, PC_CONV1_WGRAD, PC_CONV1_DGRAD,
  PC_TRANS_WGRAD, PC_TRANS_DGRAD, PC_AVGPOOL_BWD, PC_MAXPOOL_BWD, PC_STEM_WGRAD, PC_TAILS, PC_HEADS, PC_SGD, PC_PREPROC,
  PC_RN_FPROP, PC_RN_DGRAD, PC_RN_WGRAD, PC_RN_ELTWISE, PC_RN_HEAD, PC_COUNT
};

struct ProfRec { int cls; cudaEvent_t a, b; int side; };
struct ProfState {
  bool on = false;
  bool timeline = false;   // keep the real stream structure (side stream, early start, PDL) while recording: the records then give
                           // each launch's start / end on its own stream relative to a base event (mmnn_profile_timeline)
  long long launches = 0;
  std::vector<ProfRec> recs;
};
ProfState& prof_state();

struct ProfScope {
  int idx = -1;
  cudaStream_t st;
  ProfScope(int cls, cudaStream_t stream, int nlaunch = 1) : st(stream) {
    ProfState& s = prof_state();
    s.launches += nlaunch;
    if (s.on) {
      ProfRec r; r.cls = cls; r.side = 0;
      cudaEventCreate(&r.a); cudaEventCreate(&r.b);
      cudaEventRecord(r.a, st);
      idx = (int)s.recs.size();
      s.recs.push_back(r);
    }
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(prof_state().recs[idx].b, st);
  }
};

}  // namespace mmnn
