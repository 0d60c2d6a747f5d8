// shared between capi.cu and groth16_host.cu
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gnark_b200.h"
#include "internal.h"
#include "msm.cuh"

#define GB200_MAX_DEVICES 16

namespace gb200 {

struct DeviceCtx {
  // Host-side state of a device (current stream, fork/join events, pending tail) is shared by every entry point:
  // each extern "C" function holds this lock from its device_ctx() lookup to its return, so concurrent callers
  // (goroutines) on ONE device are serialised here instead of by a process-global mutex (the reference:
  // deviceProveMu, icicle.go:53-60); different devices do not contend.  Recursive: entry points call one another.
  std::recursive_mutex mu;
  bool ready = false;
  int dev = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;  // own_stream or the caller's (b200_set_stream)
  // pipelined MSMs (MsmPipe, internal.h): the accumulate kernel of MSM i runs on acc_stream and its reduction tail on
  // tail_stream while MSM i+1 decomposes and sorts on `stream`
  cudaStream_t acc_stream = nullptr;
  cudaStream_t tail_stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D of later-needed inputs overlaps compute on `stream`
  cudaEvent_t copy_ev = nullptr;
  cudaEvent_t front_ev = nullptr;  // front done (recorded on stream)
  cudaEvent_t acc_ev[2] = {nullptr, nullptr};   // accumulate of pipelined MSM k done (recorded on acc_stream), slot k & 1
  uint64_t pipe_seq = 0;           // pipelined MSMs enqueued since the last join
  cudaEvent_t tail_ev = nullptr;   // tail done (recorded on tail_stream)
  bool tail_pending = false;
  // multi-GPU (comm.cu): NCCL communicator of this device (ncclComm_t), nullptr = single device
  void* comm = nullptr;
  int comm_world = 1, comm_rank = 0;
};
void comm_teardown(DeviceCtx& c);   // comm.cu
int32_t points_allreduce_on_stream(DeviceCtx* ctx, const MsmOps* ops, const void* d_partials, size_t count, void* d_totals);

int32_t set_error(const std::string& msg);
int32_t cuda_fail(const char* what, cudaError_t e);
int32_t device_ctx(int dev, DeviceCtx** out);
int msm_window_for(size_t n);
void msm_tuning(size_t n, int nwin, int c, int precomp, int acc_blocks_per_sm, uint32_t* task_len, uint32_t* chunk);

}  // namespace gb200

struct b200_table_s {
  int dev, curve, group;
  size_t n;          // bases
  int c, nwin, precomp;
  size_t bytes;
  void* d_points = nullptr;
  const gb200::MsmOps* ops;
  // owns its device buffer: a table that fails half way through its construction does not leak it
  ~b200_table_s() { if (d_points) cudaFree(d_points); }
};

struct b200_domain_s {
  int dev, curve, logn;
  const gb200::NttOps* ops;
  void* impl;
};

namespace gb200 {
// stream-ordered scratch memory for the host orchestrations (plonk_host.cu): cudaMallocAsync / cudaFreeAsync on the
// device's stream, served from the never-shrinking pool - a proof allocates and releases GiBs of temporaries, and
// cudaMalloc / cudaFree (b200_alloc / b200_free: synchronous, unmap on free) were a visible part of its wall time
int32_t scratch_alloc(int dev, size_t bytes, void** out);
int32_t scratch_free(int dev, void* p);
int32_t msm_on_stream(DeviceCtx* ctx, b200_table_s* t, size_t off, size_t n, const void* d_scalars, void* d_out,
                      cudaEvent_t* stage_events = nullptr, bool pipelined = false);
int32_t msm_join(DeviceCtx* ctx);  // make ctx->stream wait for the pipelined tails
}

#define CK(x)                                                        \
  do {                                                               \
    cudaError_t e_ = (x);                                            \
    if (e_ != cudaSuccess) return gb200::cuda_fail(#x, e_);          \
  } while (0)

// device_ctx() + the per-device lock for the rest of the enclosing scope
#define GB_DEVICE(ctxvar, dev)                                        \
  gb200::DeviceCtx* ctxvar;                                           \
  {                                                                   \
    int32_t rc_dev_ = gb200::device_ctx((dev), &ctxvar);              \
    if (rc_dev_) return rc_dev_;                                      \
  }                                                                   \
  std::lock_guard<std::recursive_mutex> lock_##ctxvar(ctxvar->mu)

#define GUARD_BEGIN try {
#define GUARD_END                                                               \
  }                                                                             \
  catch (const std::exception& ex) { return gb200::set_error(std::string("exception: ") + ex.what()); } \
  catch (...) { return gb200::set_error("unknown exception"); }
