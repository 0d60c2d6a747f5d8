// C ABI of libgnark_b200.so (declared in include/gnark_b200.h): runtime, memory,
// MSM base tables, MSM, NTT, vector ops.  The Groth16 host layer lives in
// groth16_host.cu.  Requires a CUDA device: there is no CPU fallback.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "capi_common.h"
#include "file_stage.h"
#include "fixed_base.cuh"

namespace gb200 {

// ---- registries ------------------------------------------------------------
static const MsmOps* g_msm_ops[4][3];
static const NttOps* g_ntt_ops[4];
static const HostGroupOps* g_host_ops[4][3];
void register_msm_ops(int curve, int group, const MsmOps* ops) { g_msm_ops[curve][group] = ops; }
void register_ntt_ops(int curve, const NttOps* ops) { g_ntt_ops[curve] = ops; }
void register_host_group_ops(int curve, int group, const HostGroupOps* ops) { g_host_ops[curve][group] = ops; }
const MsmOps* get_msm_ops(int curve, int group) {
  if (curve < 0 || curve > 3 || group < 1 || group > 2) return nullptr;
  return g_msm_ops[curve][group];
}
const NttOps* get_ntt_ops(int curve) { return (curve < 0 || curve > 3) ? nullptr : g_ntt_ops[curve]; }
const HostGroupOps* get_host_group_ops(int curve, int group) {
  if (curve < 0 || curve > 3 || group < 1 || group > 2) return nullptr;
  return g_host_ops[curve][group];
}

// ---- error + device context --------------------------------------------------
thread_local std::string g_last_error;
int32_t set_error(const std::string& msg) { g_last_error = msg; return 1; }
int32_t cuda_fail(const char* what, cudaError_t e) {
  g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
  return 2;
}

// MSM tuning (mirrors the reference's env knob ICICLE_MSM_MAX_WINDOW, icicle.go:586-598)
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
static std::mutex g_ctx_mu;
static DeviceCtx g_ctx[GB200_MAX_DEVICES];

int32_t device_ctx(int dev, DeviceCtx** out) {
  if (dev < 0 || dev >= GB200_MAX_DEVICES) return set_error("invalid device id " + std::to_string(dev));
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  DeviceCtx& c = g_ctx[dev];
  if (!c.ready) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return cuda_fail("cudaGetDeviceCount (a CUDA device is required; no CPU fallback)", e);
    if (dev >= count) return set_error("device " + std::to_string(dev) + " not present (" + std::to_string(count) + " visible)");
    CK(cudaSetDevice(dev));
    // priorities: reduction tail (latency-bound, few blocks) > the library's main stream (MSM fronts, NTTs) > accumulate
    // (fills the machine for milliseconds; the next MSM's sort slips into the SM slots its blocks free)
    int lo_prio = 0, hi_prio = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    CK(cudaStreamCreateWithPriority(&c.own_stream, cudaStreamNonBlocking, (lo_prio + hi_prio) / 2));
    c.stream = c.own_stream;
    CK(cudaStreamCreateWithPriority(&c.tail_stream, cudaStreamNonBlocking, hi_prio));
    CK(cudaStreamCreateWithPriority(&c.acc_stream, cudaStreamNonBlocking, lo_prio));
    CK(cudaEventCreateWithFlags(&c.front_ev, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++) CK(cudaEventCreateWithFlags(&c.acc_ev[k], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c.tail_ev, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c.copy_ev, cudaEventDisableTiming));
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t thr = UINT64_MAX;  // keep freed workspace cached in the pool
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    // L2 fill granularity on a miss: 64 bytes.  The bucket accumulate gathers 64-byte (BN254 G1) / 96..192-byte points at
    // random from a table far larger than L2; with the driver's default (128) ncu shows 2.33 GB of DRAM traffic for
    // 1.14 GB of gathered points, with 64 it is 1.24 GB (profiles/r02_session4.md).  The kernels are multiplier-bound,
    // so the time does not move - this removes wasted HBM traffic, nothing else.  GB200_L2_FETCH=32|64|128 overrides, 0 leaves
    // the device limit alone.
    if (const int g = env_int("GB200_L2_FETCH", 64)) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g));
    c.dev = dev;
    c.ready = true;
  }
  cudaError_t e = cudaSetDevice(dev);
  if (e != cudaSuccess) return cuda_fail("cudaSetDevice", e);
  *out = &c;
  return 0;
}

int msm_window_for(size_t n) {
  int c = env_int("GB200_MSM_WINDOW", 0);
  if (c > 0) return c < 2 ? 2 : (c > 24 ? 24 : c);
  int lg = 0;
  while ((1ull << (lg + 1)) <= n) lg++;
  c = lg - 4;
  if (c < 4) c = 4;
  if (c > 16) c = 16;
  return c;
}
void msm_tuning(size_t n, int nwin, int c, int precomp, int acc_blocks_per_sm, uint32_t* task_len, uint32_t* chunk) {
  const size_t m = n * (size_t)nwin;
  const size_t buckets = (precomp ? (size_t)1 : (size_t)nwin) << (c - 1);
  // Task length.  One thread adds a task's entries; a bucket of k entries is cut into ceil(k / task_len) tasks.  Short
  // tasks fill the machine evenly (the accumulate grid is many waves deep, the short last task of every bucket idles its
  // lane for less) at the price of more partial sums for the combine kernel.  Measured on B200 (round 2, BN254 G1 2^20,
  // pipelined ms per MSM): 16: 3.50, 24: 3.42, 32: 3.43, 40: 3.45, 48: 3.49, 64: 3.54; BN254 G2 / BLS12-381 G1: 32 beats
  // 64 by 4-5 %.  Within [24, 40] ([48, 64] for grids of a dozen waves and more) the length is picked so that the grid is as close as possible to a whole number of
  // waves of resident blocks (148 SMs x blocks per SM from the occupancy API): the last, partly filled wave is the
  // other loss (BW6-761 2^18 with 64-entry tasks: 3.04 waves of 296 blocks, a quarter of the kernel's time on 4 % of it).
  int sms = 148;
  { int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const double resident = (double)sms * (acc_blocks_per_sm > 0 ? acc_blocks_per_sm : 1);
  // a grid a dozen waves deep loses little to its last wave: there the longer task wins again, because every partial
  // costs the combine kernel one full addition (2^24 points, 64-entry tasks: combine 3.2 of 50 ms; at 40 entries 5 ms)
  const double waves64 = ((double)m / 64.0 + 0.5 * (double)buckets) / 128.0 / resident;
  size_t lo = waves64 >= 12.0 ? 48 : 24, hi = waves64 >= 12.0 ? 64 : 40;
  const size_t per_bucket = m / (buckets ? buckets : 1);
  if (per_bucket < 2 * hi) { lo = 8; hi = per_bucket / 2 > 8 ? per_bucket / 2 : 8; }     // small problems: >= 2 tasks per bucket
  if (hi > 64) hi = 64;
  if (lo > hi) lo = hi;
  size_t best = hi;
  double best_waste = 2.0;
  for (size_t tl = hi; tl >= lo; tl--) {
    const double blocks = ((double)m / (double)tl + 0.5 * (double)buckets) / 128.0;    // expected tasks / threads per block
    const double waves = blocks / resident;
    const double full = waves <= 1.0 ? 1.0 : (double)(size_t)(waves + 0.999999);
    const double waste = (full - waves) / full;
    if (waste < best_waste - 0.005) { best_waste = waste; best = tl; }                 // ties: the longer task
    if (tl == lo) break;
  }
  *task_len = (uint32_t)env_int("GB200_MSM_TASK_LEN", (int)best);
  *chunk = (uint32_t)env_int("GB200_MSM_CHUNK", c >= 12 ? 8 : 4);
}

int32_t scratch_alloc(int dev, size_t bytes, void** out) {
  GB_DEVICE(ctx, dev);
  CK(cudaMallocAsync(out, bytes ? bytes : 1, ctx->stream));
  return 0;
}
int32_t scratch_free(int dev, void* p) {
  if (!p) return 0;
  GB_DEVICE(ctx, dev);
  int32_t rc = msm_join(ctx);          // pipelined MSM tails may still read the buffer on the tail stream
  if (rc) return rc;
  CK(cudaFreeAsync(p, ctx->stream));
  return 0;
}

int32_t msm_join(DeviceCtx* ctx) {
  if (ctx->tail_pending) {
    CK(cudaStreamWaitEvent(ctx->stream, ctx->tail_ev, 0));
    ctx->tail_pending = false;
    ctx->pipe_seq = 0;
  }
  return 0;
}

int32_t msm_on_stream(DeviceCtx* ctx, b200_table_s* t, size_t off, size_t n, const void* d_scalars, void* d_out,
                      cudaEvent_t* stage_events, bool pipelined) {
  if (off + n > t->n) return set_error("msm: range [off, off+n) exceeds the table");
  // the pipeline indexes its n * nwin (scalar, window) entries with 31-bit counts (radix sort, offsets, tasks)
  if ((size_t)n * (size_t)t->nwin >= ((size_t)1 << 31))
    return set_error("msm: n * windows exceeds 2^31 entries - split the call into point ranges (or shard the table) and add the results");
  if (stage_events || n == 0) pipelined = false;
  if (!pipelined) { int32_t rc = msm_join(ctx); if (rc) return rc; }
  uint32_t task_len, chunk;
  msm_tuning(n, t->nwin, t->c, t->precomp, t->ops->acc_blocks_per_sm(), &task_len, &chunk);
  size_t ws_bytes = 0;
  CK(t->ops->ws_bytes((uint32_t)n, (uint32_t)t->n, t->c, t->precomp, task_len, chunk, &ws_bytes));
  AsyncBuf ws_buf;
  CK(ws_buf.alloc(ws_bytes, ctx->stream));
  void* ws = ws_buf.p;
  MsmPipe pipe;
  if (pipelined) {
    // at most one front ahead of the accumulate stream (bounds the live workspaces): the front of MSM k waits for the
    // accumulate of MSM k - 2, whose event slot it then reuses
    const int slot = (int)(ctx->pipe_seq & 1);
    if (ctx->pipe_seq >= 2) CK(cudaStreamWaitEvent(ctx->stream, ctx->acc_ev[slot], 0));
    pipe.acc = env_int("GB200_MSM_ACC_STREAM", 1) ? ctx->acc_stream : nullptr;
    pipe.tail = ctx->tail_stream;
    pipe.front_ev = ctx->front_ev;
    pipe.acc_ev = ctx->acc_ev[slot];
    ctx->pipe_seq++;
  }
  cudaError_t e = t->ops->run(ctx->stream, (uint32_t)n, (uint32_t)t->n, (uint32_t)off, t->c, t->precomp, task_len, chunk,
                              t->d_points, d_scalars, d_out, ws, stage_events, pipelined ? &pipe : nullptr);
  // the workspace is last used by the tail kernels
  cudaError_t e2 = ws_buf.release_on(pipelined ? ctx->tail_stream : ctx->stream);
  if (pipelined && e == cudaSuccess) {
    cudaError_t e3 = cudaEventRecord(ctx->tail_ev, ctx->tail_stream);
    if (e3 != cudaSuccess) return cuda_fail("cudaEventRecord", e3);
    ctx->tail_pending = true;
  }
  if (e != cudaSuccess) return cuda_fail("msm enqueue", e);
  if (e2 != cudaSuccess) return cuda_fail("cudaFreeAsync", e2);
  return 0;
}

}  // namespace gb200

using namespace gb200;

extern "C" {

const char* b200_version(void) { return "gnark_b200 0.1.0 (sm_100a)"; }
const char* b200_last_error(void) { return g_last_error.c_str(); }

int32_t b200_device_count(int32_t* out) {
  GUARD_BEGIN
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return cuda_fail("cudaGetDeviceCount", e);
  *out = count;
  return 0;
  GUARD_END
}

int32_t b200_init(int32_t n_dev, const int32_t* dev_ids) {
  GUARD_BEGIN
  if (n_dev <= 0 || !dev_ids) { DeviceCtx* c; return device_ctx(0, &c); }
  for (int i = 0; i < n_dev; i++) { GB_DEVICE(c, dev_ids[i]); [[maybe_unused]] int32_t rc = 0; }
  return 0;
  GUARD_END
}

int32_t b200_shutdown(void) {
  GUARD_BEGIN
  std::lock_guard<std::mutex> lk(g_ctx_mu);
  for (int d = 0; d < GB200_MAX_DEVICES; d++) {
    DeviceCtx& c = g_ctx[d];
    if (!c.ready) continue;
    cudaSetDevice(d);
    cudaStreamSynchronize(c.stream);
    cudaStreamSynchronize(c.acc_stream);
    cudaStreamSynchronize(c.tail_stream);
    comm_teardown(c);
    cudaStreamDestroy(c.own_stream);
    cudaStreamDestroy(c.acc_stream);
    cudaStreamDestroy(c.tail_stream);
    cudaStreamDestroy(c.copy_stream);
    cudaEventDestroy(c.copy_ev);
    cudaEventDestroy(c.front_ev);
    cudaEventDestroy(c.acc_ev[0]);
    cudaEventDestroy(c.acc_ev[1]);
    cudaEventDestroy(c.tail_ev);
    // back to the initial state (the lock itself stays)
    c.ready = false; c.tail_pending = false; c.pipe_seq = 0;
    c.own_stream = c.stream = c.acc_stream = c.tail_stream = c.copy_stream = nullptr;
    c.copy_ev = c.front_ev = c.acc_ev[0] = c.acc_ev[1] = c.tail_ev = nullptr;
  }
  return 0;
  GUARD_END
}

int32_t b200_set_stream(int32_t dev, void* stream) {
  GUARD_BEGIN
  GB_DEVICE(c, dev); [[maybe_unused]] int32_t rc = 0;
  c->stream = stream ? reinterpret_cast<cudaStream_t>(stream) : c->own_stream;
  return 0;
  GUARD_END
}

int32_t b200_sync(int32_t dev) {
  GUARD_BEGIN
  GB_DEVICE(c, dev); [[maybe_unused]] int32_t rc = 0;
  rc = msm_join(c); if (rc) return rc;
  CK(cudaStreamSynchronize(c->stream));
  return 0;
  GUARD_END
}

int32_t b200_alloc(int32_t dev, size_t bytes, void** out) {
  GUARD_BEGIN
  GB_DEVICE(c, dev); [[maybe_unused]] int32_t rc = 0;
  CK(cudaMalloc(out, bytes ? bytes : 1));
  return 0;
  GUARD_END
}
int32_t b200_free(int32_t dev, void* p) {
  GUARD_BEGIN
  GB_DEVICE(c, dev); [[maybe_unused]] int32_t rc = 0;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaFree(p));
  return 0;
  GUARD_END
}
int32_t b200_h2d(int32_t dev, void* dst, const void* src, size_t bytes) {
  GUARD_BEGIN
  GB_DEVICE(c, dev); [[maybe_unused]] int32_t rc = 0;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));  // host pointer is only borrowed for the call
  return 0;
  GUARD_END
}
int32_t b200_d2h(int32_t dev, void* dst, const void* src, size_t bytes) {
  GUARD_BEGIN
  GB_DEVICE(c, dev); [[maybe_unused]] int32_t rc = 0;
  rc = msm_join(c); if (rc) return rc;
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
  GUARD_END
}
int32_t b200_host_alloc(size_t bytes, void** out) {
  GUARD_BEGIN
  CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
  GUARD_END
}
int32_t b200_host_free(void* p) {
  GUARD_BEGIN
  CK(cudaFreeHost(p));
  return 0;
  GUARD_END
}

// ---- tables ------------------------------------------------------------------
}  // extern "C"

namespace gb200 {
// device table of n points with room for the precomputed slabs; slab 0 (= the first n entries of d_points) is filled by
// the caller - from host / device memory, from a file region, or by the point decoder - and table_finish builds the rest
static int32_t table_new(int dev, int curve, int group, size_t n, int flags, std::unique_ptr<b200_table_s>& t) {
  const MsmOps* ops = get_msm_ops(curve, group);
  if (!ops) return set_error("table_upload: unsupported curve/group");
  t.reset(new b200_table_s());
  t->dev = dev; t->curve = curve; t->group = group; t->n = n; t->ops = ops;
  t->c = msm_window_for(n ? n : 1);
  while (msm_num_windows(ops->scalar_bits, t->c) > 64) t->c++;   // MSM_MAX_WINDOWS of the table precompute
  t->nwin = msm_num_windows(ops->scalar_bits, t->c);
  t->precomp = (flags & B200_TABLE_PRECOMP) ? 1 : 0;
  if (env_int("GB200_MSM_PRECOMP", -1) >= 0) t->precomp = env_int("GB200_MSM_PRECOMP", 0) ? 1 : 0;
  const size_t slabs = t->precomp ? (size_t)t->nwin : 1;
  if (slabs * n >= (1ull << 31)) return set_error("table_upload: table too large for 31-bit point indices; shard it");
  t->bytes = slabs * (n ? n : 1) * ops->affine_bytes;
  CK(cudaMalloc(&t->d_points, t->bytes));
  return 0;
}
static int32_t table_finish(DeviceCtx* ctx, b200_table_s* t) {
  if (t->precomp) CK(t->ops->precompute(ctx->stream, (uint32_t)t->n, t->nwin, t->c, t->d_points));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
}  // namespace gb200

extern "C" {

int32_t b200_table_upload(int32_t dev, int32_t curve, int32_t group, const void* points, size_t n, int32_t flags,
                          b200_table_t* out) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  if (!out) return set_error("table_upload: null argument");
  if (n && !points) return set_error("table_upload: null points");
  std::unique_ptr<b200_table_s> t;
  rc = table_new(dev, curve, group, n, flags, t); if (rc) return rc;
  const cudaMemcpyKind kind = (flags & B200_TABLE_SRC_ON_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  if (n) CK(cudaMemcpyAsync(t->d_points, points, n * t->ops->affine_bytes, kind, ctx->stream));
  rc = table_finish(ctx, t.get()); if (rc) return rc;
  *out = t.release();
  return 0;
  GUARD_END
}

// Table from a slice in gnark-crypto's SERIALISED point encoding (points_decode.cuh): what backend/plonk/<curve>/marshal.go
// writes for pk.Kzg / pk.KzgLagrange (compressed with WriteTo, uncompressed with WriteRawTo) after the slice's uint32
// length.  The bytes go up as they are; one thread per point decodes (a square root per compressed point) straight into
// the table.  Every point is checked (canonical coordinates, on the curve); subgroup membership is not, as with
// gnark's UnsafeReadFrom.  An invalid point fails the call with its index.
int32_t b200_table_upload_encoded(int32_t dev, int32_t curve, int32_t group, const void* bytes, size_t n, int32_t encoding,
                                  int32_t flags, b200_table_t* out) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  if (!out || (n && !bytes)) return set_error("table_upload_encoded: null argument");
  if (flags & B200_TABLE_SRC_ON_DEVICE) return set_error("table_upload_encoded: the encoded slice is read from host memory");
  if (curve < 0 || curve > 3) return set_error("table_upload_encoded: unsupported curve");
  std::unique_ptr<b200_table_s> t;
  rc = table_new(dev, curve, group, n, flags, t); if (rc) return rc;
  const size_t per = t->ops->encoded_bytes(encoding);
  if (per == 0) return set_error("table_upload_encoded: unknown encoding (1 = raw, 2 = compressed)");
  if (n) {
    AsyncBuf d_bytes, d_status;
    CK(d_bytes.alloc(n * per, ctx->stream));
    CK(d_status.alloc(2 * sizeof(uint32_t), ctx->stream));
    CK(cudaMemsetAsync(d_status.p, 0, 2 * sizeof(uint32_t), ctx->stream));
    CK(cudaMemcpyAsync(d_bytes.p, bytes, n * per, cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = t->ops->decode(ctx->stream, d_bytes.p, n, encoding, curve, group, t->d_points, (uint32_t*)d_status.p);
    if (e == cudaErrorNotSupported) return set_error("table_upload_encoded: encoding not supported for this group");
    if (e != cudaSuccess) return cuda_fail("points decode", e);
    uint32_t status[2] = {0, 0};
    CK(cudaMemcpyAsync(status, d_status.p, sizeof(status), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (status[0]) {
      static const char* why[] = {"", "invalid metadata bits", "coordinate not reduced modulo p", "point not on the curve"};
      return set_error(std::string("table_upload_encoded: point ") + std::to_string(status[1]) + ": " + why[status[0] < 4 ? status[0] : 0]);
    }
  }
  rc = table_finish(ctx, t.get()); if (rc) return rc;
  *out = t.release();
  return 0;
  GUARD_END
}

// Table straight from a file region: the point slices of gnark's ProvingKey dump
// (backend/groth16/bn254/marshal.go:375-539: G1.A, G1.B, G1.Z, G1.K, G2.B written by unsafe.WriteSlice as raw memory
// images, i.e. exactly the device table layout) are read chunk by chunk into two pinned slots and copied to the device
// while the next chunk is being read, then handed to b200_table_upload as a device-resident source.  The Go shim reads
// the dump's header with gnark-crypto's own decoder and passes the byte offset of each slice's payload; a shard passes
// the offset of ITS point range, so that no process ever holds the whole key in host memory.
int32_t b200_table_upload_file(int32_t dev, int32_t curve, int32_t group, const char* path, uint64_t byte_offset, size_t n,
                               int32_t flags, b200_table_t* out) {
  GUARD_BEGIN
  if (!path || !out) return set_error("table_upload_file: null argument");
  if (flags & B200_TABLE_SRC_ON_DEVICE) return set_error("table_upload_file: the source is a file");
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const MsmOps* ops = get_msm_ops(curve, group);
  if (!ops) return set_error("table_upload_file: unsupported curve/group");
  const size_t bytes = n * ops->affine_bytes;
  struct Fd { int fd = -1; ~Fd() { if (fd >= 0) close(fd); } } f;
  f.fd = open(path, O_RDONLY | O_CLOEXEC);
  if (f.fd < 0) return set_error(std::string("table_upload_file: cannot open ") + path + ": " + strerror(errno));
  struct stat st;
  if (fstat(f.fd, &st) != 0) return set_error(std::string("table_upload_file: fstat: ") + strerror(errno));
  if ((uint64_t)st.st_size < byte_offset || (uint64_t)st.st_size - byte_offset < bytes)
    return set_error("table_upload_file: [offset, offset + n * sizeof(point)) exceeds the file");
  std::unique_ptr<b200_table_s> t;
  rc = table_new(dev, curve, group, n, flags, t); if (rc) return rc;
  if (n) {
    // the file region goes through two pinned slots straight into slab 0 of the table: no second device copy of the
    // slice, peak HBM = the table itself
    const size_t slot_bytes = bytes < ((size_t)32 << 20) ? bytes : ((size_t)32 << 20);
    struct Pinned { void* p[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr};
                    ~Pinned() { for (int i = 0; i < 2; i++) { if (p[i]) cudaFreeHost(p[i]); if (ev[i]) cudaEventDestroy(ev[i]); } } } pin;
    for (int i = 0; i < 2; i++) {
      CK(cudaHostAlloc(&pin.p[i], slot_bytes, cudaHostAllocDefault));
      CK(cudaEventCreateWithFlags(&pin.ev[i], cudaEventDisableTiming));
    }
    bool used[2] = {false, false};
    cudaError_t ce = cudaSuccess;
    std::string err;
    char* dst = reinterpret_cast<char*>(t->d_points);
    const int r = stage_file_region(
        f.fd, byte_offset, bytes, pin.p, slot_bytes,
        [&](int slot, const void* data, size_t pos, size_t len) {
          ce = cudaMemcpyAsync(dst + pos, data, len, cudaMemcpyHostToDevice, ctx->copy_stream);
          if (ce == cudaSuccess) ce = cudaEventRecord(pin.ev[slot], ctx->copy_stream);
          used[slot] = true;
          return ce == cudaSuccess ? 0 : 1;
        },
        [&](int slot) {
          if (!used[slot]) return 0;
          ce = cudaEventSynchronize(pin.ev[slot]);
          return ce == cudaSuccess ? 0 : 1;
        },
        &err);
    cudaError_t ce2 = cudaStreamSynchronize(ctx->copy_stream);
    if (ce != cudaSuccess) return cuda_fail("table_upload_file", ce);
    if (r != 0) return set_error("table_upload_file: " + err);
    if (ce2 != cudaSuccess) return cuda_fail("table_upload_file", ce2);
  }
  rc = table_finish(ctx, t.get()); if (rc) return rc;
  *out = t.release();
  return 0;
  GUARD_END
}

int32_t b200_table_free(b200_table_t t) {
  GUARD_BEGIN
  if (!t) return 0;
  GB_DEVICE(ctx, t->dev); [[maybe_unused]] int32_t rc = 0;
  CK(cudaStreamSynchronize(ctx->stream));
  delete t;       // ~b200_table_s releases the device buffers
  return 0;
  GUARD_END
}

int32_t b200_table_info(b200_table_t t, size_t* n, int32_t* c, int32_t* nwin, int32_t* precomp, size_t* bytes) {
  GUARD_BEGIN
  if (!t) return set_error("table_info: null table");
  if (n) *n = t->n;
  if (c) *c = t->c;
  if (nwin) *nwin = t->nwin;
  if (precomp) *precomp = t->precomp;
  if (bytes) *bytes = t->bytes;
  return 0;
  GUARD_END
}

// ---- MSM -----------------------------------------------------------------------
int32_t b200_msm_async(b200_table_t t, size_t off, size_t n, const void* d_scalars, void* d_out) {
  GUARD_BEGIN
  if (!t) return set_error("msm: null table");
  GB_DEVICE(ctx, t->dev); [[maybe_unused]] int32_t rc = 0;
  return msm_on_stream(ctx, t, off, n, d_scalars, d_out);
  GUARD_END
}

int32_t b200_fixed_base_batch(int32_t dev, int32_t curve, int32_t group, const void* base_affine,
                              const void* scalars, int32_t scalars_on_device, size_t n, void* out_affine,
                              int32_t out_on_device) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const MsmOps* ops = get_msm_ops(curve, group);
  if (!ops) return set_error("fixed_base_batch: unsupported curve/group");
  if (!base_affine || (n && (!scalars || !out_affine))) return set_error("fixed_base_batch: null argument");
  if (n == 0) return 0;
  rc = msm_join(ctx); if (rc) return rc;
  void* d_sc = const_cast<void*>(scalars);
  void* d_out = out_affine;
  AsyncBuf sc_buf, out_buf;
  if (!scalars_on_device) {
    CK(sc_buf.alloc(n * ops->fr_bytes, ctx->stream));
    d_sc = sc_buf.p;
    CK(cudaMemcpyAsync(d_sc, scalars, n * ops->fr_bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (!out_on_device) { CK(out_buf.alloc(n * ops->affine_bytes, ctx->stream)); d_out = out_buf.p; }
  int c = env_int("GB200_FIXED_BASE_WINDOW", 0);
  if (c < 2 || c > 16) c = fixed_base_window_for(n);
  cudaError_t e = ops->fixed_base(ctx->stream, base_affine, d_sc, n, c, d_out);
  if (e == cudaSuccess && !out_on_device)
    e = cudaMemcpyAsync(out_affine, d_out, n * ops->affine_bytes, cudaMemcpyDeviceToHost, ctx->stream);
  sc_buf.release_on(ctx->stream);
  out_buf.release_on(ctx->stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);   // host pointers are only borrowed for the call
  if (e != cudaSuccess) return cuda_fail("fixed_base_batch", e);
  if (e2 != cudaSuccess) return cuda_fail("fixed_base_batch", e2);
  return 0;
  GUARD_END
}

// step profile: device time of each pipeline stage of one MSM (ms), stages =
// {decompose, sort, offsets+task scan, accumulate, combine, reduce chunks, set sum+finish}
int32_t b200_msm_profile(b200_table_t t, size_t off, size_t n, const void* d_scalars, void* d_out, float* stage_ms) {
  GUARD_BEGIN
  if (!t || !stage_ms) return set_error("msm_profile: null argument");
  GB_DEVICE(ctx, t->dev); [[maybe_unused]] int32_t rc = 0;
  cudaEvent_t ev[8];
  for (int k = 0; k < 8; k++) CK(cudaEventCreate(&ev[k]));
  rc = msm_on_stream(ctx, t, off, n, d_scalars, d_out, ev);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc == 0 && e == cudaSuccess && n > 0)
    for (int k = 0; k < 7; k++) cudaEventElapsedTime(&stage_ms[k], ev[k], ev[k + 1]);
  for (int k = 0; k < 8; k++) cudaEventDestroy(ev[k]);
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail("msm_profile", e);
  return 0;
  GUARD_END
}

int32_t b200_msm_pipelined(b200_table_t t, size_t off, size_t n, const void* d_scalars, void* d_out) {
  GUARD_BEGIN
  if (!t) return set_error("msm: null table");
  GB_DEVICE(ctx, t->dev); [[maybe_unused]] int32_t rc = 0;
  return msm_on_stream(ctx, t, off, n, d_scalars, d_out, nullptr, true);
  GUARD_END
}
int32_t b200_msm_join(int32_t dev) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  return msm_join(ctx);
  GUARD_END
}

int32_t b200_msm(b200_table_t t, size_t off, size_t n, const void* scalars, int32_t on_dev, void* out_host) {
  GUARD_BEGIN
  if (!t) return set_error("msm: null table");
  if (n && !scalars) return set_error("msm: null scalars");
  GB_DEVICE(ctx, t->dev); [[maybe_unused]] int32_t rc = 0;
  AsyncBuf d_sc, d_out;      // released on every path, early error returns included
  CK(d_out.alloc(t->ops->jac_bytes, ctx->stream));
  const void* sc = scalars;
  if (!on_dev && n) {
    CK(d_sc.alloc(n * t->ops->fr_bytes, ctx->stream));
    CK(cudaMemcpyAsync(d_sc.p, scalars, n * t->ops->fr_bytes, cudaMemcpyHostToDevice, ctx->stream));
    sc = d_sc.p;
  }
  rc = msm_on_stream(ctx, t, off, n, sc, d_out.p);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out_host, d_out.p, t->ops->jac_bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail("msm result copy", e);
  }
  d_sc.release_on(ctx->stream);
  d_out.release_on(ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail("msm", e);
  return 0;
  GUARD_END
}

// Asynchronous twin of b200_msm for a STREAM of MSMs from host buffers: the scalar upload of call i+1 (copy
// stream) and the reduction tail + result download of call i (tail stream) overlap the sort/accumulate of the
// calls around them.  Both host buffers must stay valid until b200_sync(dev) returns - use b200_host_alloc
// memory (pinned, C-owned: legal to hold across cgo calls), never Go memory.
int32_t b200_msm_submit(b200_table_t t, size_t off, size_t n, const void* scalars_host, void* out_host) {
  GUARD_BEGIN
  if (!t) return set_error("msm_submit: null table");
  if (!out_host || (n && !scalars_host)) return set_error("msm_submit: null argument");
  GB_DEVICE(ctx, t->dev); [[maybe_unused]] int32_t rc = 0;
  AsyncBuf d_sc, d_out;
  CK(d_out.alloc(t->ops->jac_bytes, ctx->stream));
  if (n) {
    CK(d_sc.alloc(n * t->ops->fr_bytes, ctx->copy_stream));
    CK(cudaMemcpyAsync(d_sc.p, scalars_host, n * t->ops->fr_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->copy_ev, ctx->copy_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev, 0));
  }
  rc = msm_on_stream(ctx, t, off, n, d_sc.p, d_out.p, nullptr, /*pipelined=*/true);
  // the result is produced on the tail stream when the call was pipelined (n > 0), else on the main stream
  cudaStream_t res_stream = (n > 0) ? ctx->tail_stream : ctx->stream;
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(out_host, d_out.p, t->ops->jac_bytes, cudaMemcpyDeviceToHost, res_stream);
    if (e != cudaSuccess) rc = cuda_fail("msm_submit result copy", e);
    if (rc == 0 && n > 0) {   // the join point must cover the download too
      e = cudaEventRecord(ctx->tail_ev, ctx->tail_stream);
      if (e != cudaSuccess) rc = cuda_fail("cudaEventRecord", e);
    }
  }
  d_sc.release_on(ctx->stream);       // last read by the decompose kernel on the main stream
  d_out.release_on(res_stream);
  return rc;
  GUARD_END
}

// b200_msm_submit with the result left on the DEVICE (d_out_jac, valid after b200_msm_join / b200_sync): the form a
// sharded MSM needs, whose partial results go through b200_points_allreduce before anything is downloaded.
int32_t b200_msm_submit_dev(b200_table_t t, size_t off, size_t n, const void* scalars_host, void* d_out_jac) {
  GUARD_BEGIN
  if (!t) return set_error("msm_submit_dev: null table");
  if (!d_out_jac || (n && !scalars_host)) return set_error("msm_submit_dev: null argument");
  GB_DEVICE(ctx, t->dev); [[maybe_unused]] int32_t rc = 0;
  AsyncBuf d_sc;
  if (n) {
    CK(d_sc.alloc(n * t->ops->fr_bytes, ctx->copy_stream));
    CK(cudaMemcpyAsync(d_sc.p, scalars_host, n * t->ops->fr_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    CK(cudaEventRecord(ctx->copy_ev, ctx->copy_stream));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->copy_ev, 0));
  }
  rc = msm_on_stream(ctx, t, off, n, d_sc.p, d_out_jac, nullptr, /*pipelined=*/true);
  d_sc.release_on(ctx->stream);       // last read by the decompose kernel on the main stream
  return rc;
  GUARD_END
}

int32_t b200_msm_g1(b200_table_t t, size_t off, size_t n, const void* s, int32_t on_dev, void* out) {
  if (t && t->group != 1) return set_error("msm_g1: table holds G2 points");
  return b200_msm(t, off, n, s, on_dev, out);
}
int32_t b200_msm_g2(b200_table_t t, size_t off, size_t n, const void* s, int32_t on_dev, void* out) {
  if (t && t->group != 2) return set_error("msm_g2: table holds G1 points");
  return b200_msm(t, off, n, s, on_dev, out);
}

// ---- NTT -----------------------------------------------------------------------
int32_t b200_ntt_domain_new(int32_t dev, int32_t curve, uint32_t log2n, const void* gen, const void* coset,
                            b200_domain_t* out) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const NttOps* ops = get_ntt_ops(curve);
  if (!ops) return set_error("ntt_domain_new: unsupported curve");
  if ((int)log2n > ops->two_adicity || log2n > 30)
    return set_error("ntt_domain_new: log2n exceeds the field's 2-adicity / supported size");
  cudaError_t e = cudaSuccess;
  void* impl = ops->domain_new(ctx->stream, (int)log2n, gen, coset, &e);
  if (!impl) return cuda_fail("ntt_domain_new", e);
  b200_domain_s* d = new b200_domain_s();
  d->dev = dev; d->curve = curve; d->logn = (int)log2n; d->ops = ops; d->impl = impl;
  *out = d;
  return 0;
  GUARD_END
}

int32_t b200_ntt_domain_free(b200_domain_t d) {
  GUARD_BEGIN
  if (!d) return 0;
  GB_DEVICE(ctx, d->dev); [[maybe_unused]] int32_t rc = 0;
  CK(cudaStreamSynchronize(ctx->stream));
  d->ops->domain_free(d->impl);
  delete d;
  return 0;
  GUARD_END
}

int32_t b200_ntt_async(b200_domain_t d, void* d_data, int32_t inverse, int32_t decimation, int32_t on_coset) {
  GUARD_BEGIN
  if (!d) return set_error("ntt: null domain");
  if (decimation != B200_DIF && decimation != B200_DIT) return set_error("ntt: decimation must be DIF or DIT");
  GB_DEVICE(ctx, d->dev); [[maybe_unused]] int32_t rc = 0;
  CK(d->ops->ntt(ctx->stream, d->impl, d_data, inverse, decimation, on_coset));
  return 0;
  GUARD_END
}

int32_t b200_ntt(b200_domain_t d, void* data, int32_t on_dev, int32_t inverse, int32_t decimation, int32_t on_coset) {
  GUARD_BEGIN
  if (!d) return set_error("ntt: null domain");
  if (decimation != B200_DIF && decimation != B200_DIT) return set_error("ntt: decimation must be DIF or DIT");
  GB_DEVICE(ctx, d->dev); [[maybe_unused]] int32_t rc = 0;
  const size_t bytes = ((size_t)1 << d->logn) * d->ops->fr_bytes;
  void* buf = data;
  AsyncBuf tmp;
  if (!on_dev) {
    CK(tmp.alloc(bytes, ctx->stream));
    buf = tmp.p;
    CK(cudaMemcpyAsync(buf, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  cudaError_t e = d->ops->ntt(ctx->stream, d->impl, buf, inverse, decimation, on_coset);
  if (!on_dev) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(data, buf, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    tmp.release_on(ctx->stream);
  }
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail("ntt", e);
  if (e2 != cudaSuccess) return cuda_fail("ntt", e2);
  return 0;
  GUARD_END
}

int32_t b200_groth16_compute_h(b200_domain_t d, const void* a, const void* b, const void* c, size_t len,
                               int32_t in_dev, void* h_out, int32_t out_dev) {
  GUARD_BEGIN
  if (!d) return set_error("compute_h: null domain");
  GB_DEVICE(ctx, d->dev); [[maybe_unused]] int32_t rc = 0;
  const size_t n = (size_t)1 << d->logn;
  if (len > n) return set_error("compute_h: len exceeds the domain");
  const size_t fb = d->ops->fr_bytes;
  AsyncBuf vb[3];
  void* v[3] = {nullptr, nullptr, nullptr};
  const void* src[3] = {a, b, c};
  const cudaMemcpyKind kind = in_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (int k = 0; k < 3; k++) {
    CK(vb[k].alloc(n * fb, ctx->stream));
    v[k] = vb[k].p;
    if (len) CK(cudaMemcpyAsync(v[k], src[k], len * fb, kind, ctx->stream));
    if (len < n) CK(cudaMemsetAsync((char*)v[k] + len * fb, 0, (n - len) * fb, ctx->stream));
  }
  cudaError_t e = d->ops->compute_h(ctx->stream, d->impl, v[0], v[1], v[2]);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(h_out, v[0], n * fb, out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream);
  for (int k = 0; k < 3; k++) vb[k].release_on(ctx->stream);
  cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cuda_fail("compute_h", e);
  if (e2 != cudaSuccess) return cuda_fail("compute_h", e2);
  return 0;
  GUARD_END
}

// ---- vector ops -----------------------------------------------------------------
int32_t b200_vec_op(int32_t dev, int32_t curve, int32_t op, void* out, const void* a, const void* b, size_t n) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const NttOps* ops = get_ntt_ops(curve);
  if (!ops) return set_error("vec_op: unsupported curve");
  if (op < 0 || op > 2) return set_error("vec_op: unknown op");
  CK(ops->vec_op(ctx->stream, op, out, a, b, n));
  return 0;
  GUARD_END
}
int32_t b200_vec_bit_reverse(int32_t dev, int32_t curve, void* data, uint32_t log2n) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const NttOps* ops = get_ntt_ops(curve);
  if (!ops) return set_error("vec_bit_reverse: unsupported curve");
  CK(ops->bit_reverse(ctx->stream, data, log2n));
  return 0;
  GUARD_END
}
int32_t b200_vec_scale_powers(int32_t dev, int32_t curve, void* data, size_t n, const void* s, const void* g) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const NttOps* ops = get_ntt_ops(curve);
  if (!ops) return set_error("vec_scale_powers: unsupported curve");
  CK(ops->scale_powers(ctx->stream, data, n, s, g));
  return 0;
  GUARD_END
}

int32_t b200_vec_batch_invert(int32_t dev, int32_t curve, void* data, size_t n) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const NttOps* ops = get_ntt_ops(curve);
  if (!ops) return set_error("vec_batch_invert: unsupported curve");
  CK(ops->batch_invert(ctx->stream, data, n));
  return 0;
  GUARD_END
}

int32_t b200_vec_axpy(int32_t dev, int32_t curve, void* y, const void* a, const void* x, size_t n) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const NttOps* ops = get_ntt_ops(curve);
  if (!ops) return set_error("vec_axpy: unsupported curve");
  if (!a || (n && (!y || !x))) return set_error("vec_axpy: null argument");
  CK(ops->axpy(ctx->stream, y, a, x, n));
  return 0;
  GUARD_END
}
int32_t b200_vec_scan(int32_t dev, int32_t curve, int32_t op, void* data, size_t n, int32_t exclusive) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const NttOps* ops = get_ntt_ops(curve);
  if (!ops) return set_error("vec_scan: unsupported curve");
  if (op != B200_SCAN_PRODUCT && op != B200_SCAN_SUM) return set_error("vec_scan: unknown op");
  CK(ops->scan(ctx->stream, op, data, n, exclusive));
  return 0;
  GUARD_END
}
int32_t b200_plonk_build_z(b200_domain_t d0, const void* l, const void* r, const void* o, const int64_t* perm,
                           const void* beta, const void* gamma, void* z) {
  GUARD_BEGIN
  if (!d0 || !l || !r || !o || !perm || !beta || !gamma || !z) return set_error("plonk_build_z: null argument");
  GB_DEVICE(ctx, d0->dev); [[maybe_unused]] int32_t rc = 0;
  CK(d0->ops->plonk_build_z(ctx->stream, d0->impl, l, r, o, perm, beta, gamma, z));
  return 0;
  GUARD_END
}
int32_t b200_poly_eval(int32_t dev, int32_t curve, const void* c, size_t n, const void* x, void* out) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const NttOps* ops = get_ntt_ops(curve);
  if (!ops) return set_error("poly_eval: unsupported curve");
  if (!x || !out || (n && !c)) return set_error("poly_eval: null argument");
  CK(ops->poly_eval(ctx->stream, c, n, x, out));
  return 0;
  GUARD_END
}
int32_t b200_poly_div_by_linear(int32_t dev, int32_t curve, void* c, size_t n, const void* z, void* rem) {
  GUARD_BEGIN
  GB_DEVICE(ctx, dev); [[maybe_unused]] int32_t rc = 0;
  const NttOps* ops = get_ntt_ops(curve);
  if (!ops) return set_error("poly_div_by_linear: unsupported curve");
  if (!z || !rem || (n && !c)) return set_error("poly_div_by_linear: null argument");
  CK(ops->poly_div_linear(ctx->stream, c, n, z, rem));
  return 0;
  GUARD_END
}

// ---- PLONK ------------------------------------------------------------------------
int32_t b200_plonk_constraints_coset(b200_domain_t d0, const void* big_coset_gen, const void* big_gen,
                                     const b200_plonk_coset_args* args) {
  GUARD_BEGIN
  if (!d0 || !big_coset_gen || !big_gen || !args) return set_error("plonk_constraints_coset: null argument");
  GB_DEVICE(ctx, d0->dev); [[maybe_unused]] int32_t rc = 0;
  cudaError_t e = d0->ops->plonk_coset(ctx->stream, d0->impl, big_coset_gen, big_gen, args);
  if (e == cudaErrorInvalidValue) return set_error("plonk_constraints_coset: invalid rho / coset index / blinding length");
  if (e != cudaSuccess) return cuda_fail("plonk_constraints_coset", e);
  return 0;
  GUARD_END
}
int32_t b200_plonk_bsb22_coset(b200_domain_t d0, const void* d_qcp, const void* d_pi2, uint32_t coset_index, uint32_t rho,
                               void* d_out) {
  GUARD_BEGIN
  if (!d0 || !d_qcp || !d_pi2 || !d_out) return set_error("plonk_bsb22_coset: null argument");
  GB_DEVICE(ctx, d0->dev); [[maybe_unused]] int32_t rc = 0;
  cudaError_t e = d0->ops->plonk_bsb22(ctx->stream, d0->impl, d_qcp, d_pi2, coset_index, rho, d_out);
  if (e == cudaErrorInvalidValue) return set_error("plonk_bsb22_coset: invalid rho / coset index");
  if (e != cudaSuccess) return cuda_fail("plonk_bsb22_coset", e);
  return 0;
  GUARD_END
}
int32_t b200_plonk_divide_by_zh(b200_domain_t d1, uint32_t log_n0, void* d_data) {
  GUARD_BEGIN
  if (!d1 || !d_data) return set_error("plonk_divide_by_zh: null argument");
  GB_DEVICE(ctx, d1->dev); [[maybe_unused]] int32_t rc = 0;
  cudaError_t e = d1->ops->plonk_divide_by_zh(ctx->stream, d1->impl, log_n0, d_data);
  if (e == cudaErrorInvalidValue) return set_error("plonk_divide_by_zh: invalid domain ratio");
  if (e != cudaSuccess) return cuda_fail("plonk_divide_by_zh", e);
  return 0;
  GUARD_END
}

}  // extern "C"
