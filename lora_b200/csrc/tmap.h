// Host-side TMA tensor-map encoding. cuTensorMapEncodeTiled is fetched through the runtime's
// driver-entry-point query so the library has no link-time dependency on libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace lb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cudaFuncSetAttribute is per DEVICE: remember which devices of this process already have the
// dynamic shared memory limit of `func` raised (one bit per device ordinal; a process that drives
// cuda:1 after cuda:0 -- single-process DP, tests on a second GPU -- must set it again).
inline bool ensure_dyn_smem(const void* func, int bytes, unsigned long long& done_mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (dev >= 0 && dev < 64 && ((done_mask >> dev) & 1ull)) return true;
  if (cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess)
    return false;
  if (dev >= 0 && dev < 64) done_mask |= 1ull << dev;
  return true;
}
inline int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  static int cached[64] = {0};
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 0;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

// Launch through cudaLaunchKernelEx so that launch attributes can be attached: a thread-block
// cluster along z (cluster_z > 1) and/or programmatic dependent launch (pdl; see ptx.cuh).
inline int& pdl_flag() {
  static int flag = -1;
  if (flag < 0) {
    const char* e = getenv("LB_PDL");
    flag = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return flag;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                             cudaStream_t stream, int cluster_z, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_z > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = 1;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = cluster_z;
    ++n;
  }
  if (pdl_flag()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Benchmark knob: LB_NO_SPLITK=1 in the environment disables the automatic split-K plans.
inline bool splitk_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LB_NO_SPLITK");
    v = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  return v != 0;
}

// Split-K workspace (fused_core.cuh, SPLITK): per-tile fp32 accumulation buffers, zero on entry
// and wiped by the kernel that used them. One allocation per device, handed out as a ring so
// that consecutive launches never share a region (a launch is stream-ordered after the previous
// user of its region as long as fewer than ~4 launches run concurrently). Allocated at the first
// split-K launch -- which must not happen inside a stream capture (cudaMalloc is illegal there);
// the training engine's warm-up steps precede its capture.
struct SplitWorkspace {
  float* buf = nullptr;
  unsigned int* counters = nullptr;
  size_t bytes = 0, off = 0;
  unsigned int n_counters = 0, counter_off = 0;
};
inline bool split_ws_reserve(cudaStream_t stream, size_t need_bytes, unsigned int need_counters,
                             float** ws, unsigned int** counters) {
  static SplitWorkspace per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  SplitWorkspace& w = per_dev[dev];
  if (w.buf == nullptr) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return false;
    const size_t bytes = size_t(96) << 20;
    const unsigned int n = 1u << 16;
    if (cudaMalloc(&w.buf, bytes) != cudaSuccess) return false;
    if (cudaMemset(w.buf, 0, bytes) != cudaSuccess) return false;   // accumulation buffers start at zero
    if (cudaMalloc(&w.counters, n * sizeof(unsigned int)) != cudaSuccess) return false;
    if (cudaMemset(w.counters, 0, n * sizeof(unsigned int)) != cudaSuccess) return false;
    w.bytes = bytes;
    w.n_counters = n;
  }
  need_bytes = (need_bytes + 255) & ~size_t(255);
  if (need_bytes > w.bytes / 2 || need_counters > w.n_counters / 2) return false;
  if (w.off + need_bytes > w.bytes) w.off = 0;
  if (w.counter_off + need_counters > w.n_counters) w.counter_off = 0;
  *ws = reinterpret_cast<float*>(reinterpret_cast<char*>(w.buf) + w.off);
  *counters = w.counters + w.counter_off;
  w.off += need_bytes;
  w.counter_off += need_counters;
  return true;
}

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Row-major 2-D tensor [rows, cols] (cols contiguous), box = [box_rows, box_cols].
// OOB reads are zero-filled, OOB writes clipped. swz128: the inner box extent must be 128 bytes.
inline bool tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes,
                    uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows,
                    bool swz128) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swz128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// NHWC activation [N, H, W, C] (C contiguous) viewed as a 4-D tensor {C, W, H, N};
// box = {box_c, box_w, box_h, 1}. Used by the implicit-GEMM conv path: a (box_h x box_w) pixel
// rectangle lands in shared memory as box_h*box_w rows of box_c channels (128 B each).
inline bool tmap_nhwc(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elem_bytes,
                      uint64_t C, uint64_t W, uint64_t H, uint64_t N, uint32_t box_c,
                      uint32_t box_w, uint32_t box_h) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {C, W, H, N};
  cuuint64_t strides[3] = {C * static_cast<uint64_t>(elem_bytes),
                           W * C * static_cast<uint64_t>(elem_bytes),
                           H * W * C * static_cast<uint64_t>(elem_bytes)};
  cuuint32_t box[4] = {box_c, box_w, box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, dt, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace lb
