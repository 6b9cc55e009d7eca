// Host side of the C ABI declared in include/cubecl_b200.h.
//
// What this file is: the B200-native stand-in for the pieces of cubecl-cuda that sit between `ComputeClient::launch` and
// `cuLaunchKernel` on the dense-LA hot path -- minus NVRTC.  It loads libcuda (and, lazily, libnccl) with dlopen exactly
// like cudarc's dynamic loading does (reference Cargo.toml:179-187), retains the device's primary context, loads the
// PREBUILT sm_100a cubins embedded in this library with cuModuleLoadData (the call the reference makes with NVRTC's PTX at
// crates/cubecl-cuda/src/compute/context.rs:293), keeps a small exclusive-page memory pool + pinned staging
// (semantics of crates/cubecl-runtime/src/memory_management, crates/cubecl-cuda/src/compute/storage/gpu.rs:174-207),
// encodes TMA descriptors (same cuTensorMapEncodeTiled call as crates/cubecl-cuda/src/compute/server.rs:1210-1224) and
// launches.  No kernels are generated, compiled or autotuned at run time and nothing here can run without a GPU.
#include "../../include/cubecl_b200.h"

#include <cuda.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

// ================================================================================================ embedded cubins
extern "C" {
extern const unsigned char b200_cubin_gemm[];
extern const unsigned char b200_cubin_gemm_end[];
extern const unsigned char b200_cubin_gemm_b[];
extern const unsigned char b200_cubin_gemm_b_end[];
extern const unsigned char b200_cubin_gemm_c[];
extern const unsigned char b200_cubin_gemm_c_end[];
extern const unsigned char b200_cubin_gemm_mx[];
extern const unsigned char b200_cubin_gemm_mx_end[];
extern const unsigned char b200_cubin_reduce[];
extern const unsigned char b200_cubin_reduce_end[];
extern const unsigned char b200_cubin_aux[];
extern const unsigned char b200_cubin_aux_end[];
}

// ================================================================================================ errors
static thread_local std::string g_last_error;

static int fail(int status, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return status;
}

extern "C" const char* b200_last_error(void) { return g_last_error.c_str(); }
extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }

// ================================================================================================ driver loading
#define DRV_FUNCTIONS(X)                                                                                              \
  X(cuInit) X(cuDeviceGetCount) X(cuDeviceGet) X(cuDeviceGetName) X(cuDeviceGetAttribute) X(cuDeviceTotalMem)        \
  X(cuDevicePrimaryCtxRetain) X(cuDevicePrimaryCtxRelease) X(cuCtxSetCurrent) X(cuCtxSynchronize)                    \
  X(cuModuleLoadData) X(cuModuleUnload) X(cuModuleGetFunction) X(cuFuncSetAttribute) X(cuFuncGetAttribute)           \
  X(cuLaunchKernelEx) X(cuLaunchKernel) X(cuOccupancyMaxActiveClusters)                                              \
  X(cuMemAlloc) X(cuMemFree) X(cuMemAllocHost) X(cuMemFreeHost) X(cuMemcpyHtoDAsync) X(cuMemcpyDtoHAsync)            \
  X(cuMemcpyDtoDAsync) X(cuMemsetD32Async) X(cuMemGetInfo)                                                           \
  X(cuStreamCreate) X(cuStreamDestroy) X(cuStreamSynchronize) X(cuStreamWaitEvent)                                   \
  X(cuEventCreate) X(cuEventRecord) X(cuEventElapsedTime) X(cuEventDestroy) X(cuEventSynchronize) X(cuEventQuery)                    \
  X(cuTensorMapEncodeTiled) X(cuGetErrorString) X(cuGetErrorName)                                                    \
  X(cuIpcGetMemHandle) X(cuIpcOpenMemHandle) X(cuIpcCloseMemHandle) X(cuCtxEnablePeerAccess) X(cuDeviceCanAccessPeer)

struct Driver {
  void* lib = nullptr;
  bool ok = false;
  std::string why;
#define X(n) decltype(&n) n##_p = nullptr;
  DRV_FUNCTIONS(X)
#undef X
};
static Driver g_drv;
static std::once_flag g_drv_once;

static void load_driver() {
  const char* names[] = {"libcuda.so.1", "libcuda.so"};
  for (const char* n : names) {
    g_drv.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_drv.lib) break;
  }
  if (!g_drv.lib) {
    g_drv.why = "cannot dlopen libcuda.so.1 (no NVIDIA driver): this library has no CPU fallback";
    return;
  }
  using GetProc = CUresult (*)(const char*, void**, int, cuuint64_t, CUdriverProcAddressQueryResult*);
  GetProc get_proc = reinterpret_cast<GetProc>(dlsym(g_drv.lib, "cuGetProcAddress_v2"));
  if (!get_proc) {
    g_drv.why = "libcuda has no cuGetProcAddress_v2 (driver older than CUDA 12)";
    return;
  }
#define X(n)                                                                                              \
  {                                                                                                       \
    void* fp = nullptr;                                                                                   \
    CUdriverProcAddressQueryResult q;                                                                     \
    CUresult r = get_proc(#n, &fp, 12080, CU_GET_PROC_ADDRESS_DEFAULT, &q);                               \
    if (r != CUDA_SUCCESS || !fp) {                                                                       \
      g_drv.why = std::string("driver symbol missing: ") + #n;                                            \
      return;                                                                                             \
    }                                                                                                     \
    g_drv.n##_p = reinterpret_cast<decltype(&n)>(fp);                                                     \
  }
  DRV_FUNCTIONS(X)
#undef X
  CUresult r = g_drv.cuInit_p(0);
  if (r != CUDA_SUCCESS) {
    g_drv.why = "cuInit failed (" + std::to_string(static_cast<int>(r)) + "): no usable GPU";
    return;
  }
  g_drv.ok = true;
}

static int ensure_driver() {
  std::call_once(g_drv_once, load_driver);
  if (!g_drv.ok) return fail(B200_ERR_NO_DEVICE, "%s", g_drv.why.c_str());
  return B200_OK;
}

static const char* cu_err(CUresult r) {
  const char* s = nullptr;
  if (g_drv.cuGetErrorString_p && g_drv.cuGetErrorString_p(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown CUDA error";
}

static int map_cu(CUresult r) {
  switch (r) {
    case CUDA_ERROR_OUT_OF_MEMORY: return B200_ERR_OUT_OF_MEMORY;
    case CUDA_ERROR_LAUNCH_OUT_OF_RESOURCES: return B200_ERR_TOO_MANY_RESOURCES;
    case CUDA_ERROR_INVALID_IMAGE:
    case CUDA_ERROR_NO_BINARY_FOR_GPU:
    case CUDA_ERROR_INVALID_SOURCE: return B200_ERR_COMPILATION;
    case CUDA_ERROR_ILLEGAL_ADDRESS:
    case CUDA_ERROR_LAUNCH_FAILED:
    case CUDA_ERROR_ILLEGAL_INSTRUCTION:
    case CUDA_ERROR_MISALIGNED_ADDRESS:
    case CUDA_ERROR_HARDWARE_STACK_ERROR:
    case CUDA_ERROR_LAUNCH_TIMEOUT: return B200_ERR_UNHEALTHY;
    default: return B200_ERR_UNKNOWN;
  }
}

#define CU_CHECK(call)                                                                               \
  do {                                                                                               \
    CUresult _r = (call);                                                                            \
    if (_r != CUDA_SUCCESS) return fail(map_cu(_r), "%s failed: %s (%d)", #call, cu_err(_r), (int)_r); \
  } while (0)

// ================================================================================================ NCCL loading (lazy)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
// ncclDataType_t / ncclRedOp_t numeric values are ABI-stable across NCCL 2.x
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6,
       ncclFloat32 = 7, ncclFloat64 = 8, ncclBfloat16 = 9 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3, ncclAvg = 4 };

struct Nccl {
  void* lib = nullptr;
  bool ok = false;
  std::string why;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, CUstream) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static Nccl g_nccl;
static std::once_flag g_nccl_once;

static void load_nccl() {
  // If torch already mapped its bundled libnccl.so.2 the soname lookup returns that copy; otherwise the system one.
  g_nccl.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!g_nccl.lib) {
    g_nccl.why = std::string("cannot dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
    return;
  }
#define L(field, sym)                                                        \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, sym)); \
  if (!g_nccl.field) { g_nccl.why = std::string("nccl symbol missing: ") + sym; return; }
  L(GetUniqueId, "ncclGetUniqueId")
  L(CommInitRank, "ncclCommInitRank")
  L(CommDestroy, "ncclCommDestroy")
  L(AllReduce, "ncclAllReduce")
  L(GetErrorString, "ncclGetErrorString")
#undef L
  g_nccl.ok = true;
}

static int ensure_nccl() {
  std::call_once(g_nccl_once, load_nccl);
  if (!g_nccl.ok) return fail(B200_ERR_COMM, "%s", g_nccl.why.c_str());
  return B200_OK;
}

#define NCCL_CHECK(call)                                                                                      \
  do {                                                                                                        \
    int _r = (call);                                                                                          \
    if (_r != ncclSuccess) return fail(B200_ERR_COMM, "%s failed: %s", #call, g_nccl.GetErrorString(_r));     \
  } while (0)

// ================================================================================================ kernel parameter blocks
// (layouts mirror the structs in gemm_tcgen05.cu / reduce.cu / aux_kernels.cu)
struct GemmParams {
  uint64_t out, out_row_stride, out_batch_stride;
  uint32_t M, N, K, batch;
  uint32_t tiles_m, tiles_n, group_m;
  uint32_t a_bmul, b_bmul, vec_store;
  uint32_t k_segments, epi_act;
  uint64_t bias;
  float alpha;
  uint32_t epi_on;
  uint32_t full_tiles, sk_tiles, sk_ranges, sk_umax;  // stream-K head, see gemm_tcgen05.cu
  uint64_t split_ws, split_tickets;
  uint32_t sf_fmt_a, sf_fmt_b, sf_tiles_a, sf_tiles_b;  // block-scaled kinds; operand formats of a mixed 8-bit pair
  uint32_t tma_store, fmt_mixed;
  uint32_t hyb, hyb_nba, hyb_nbb;  // hybrid f32 schedule: tf32 main product + two bf16 cross terms (gemm_tcgen05.cu)
  uint32_t sf_flags;               // block-scaled kinds: bit 0 = scale copies by the MMA thread (gemm.sf_copy=mma)
};
struct PackScalesParams {
  uint64_t in, out;
  uint32_t batch, rows, n_scales, tiles, atoms, pad_value;
  uint32_t tile_rows, chunks_per_tile;
};
struct ScaledSimtParams {
  uint64_t a, b, sa, sb, out;
  uint32_t batch, M, N, K;
  uint32_t a_dtype, b_dtype, out_dtype, scale_block;
  uint32_t a_bmul, b_bmul, scale_ue4m3, pad1;
};
struct ReduceParams {
  uint64_t in, out, out2, final_out, ws;
  uint64_t outer, len, inner;
  uint64_t s_outer, s_len;       // element strides of the outer and the reduced axis
  uint64_t row_len, row_pitch;   // inner offset i -> (i / row_len) * row_pitch + i % row_len (row_len == inner: no pitch)
  uint64_t seg_len;              // the reduced axis is cut into nseg segments (first pass of a two-pass reduction)
  uint32_t nseg, ctu;
  float scale;
  uint32_t flags;                // 1: record stage timings in the workspace debug words, 2: column kernels use vector units
};
struct ArgCombineParams {
  uint64_t keys, idx, out;
  uint64_t outer, nseg, inner;
};
struct FillParams {
  uint64_t out, n, seed;
  float lo, scale;
  uint32_t dtype, mode, modulus, pad;
};
struct SimtGemmParams {
  uint64_t a, b, out;
  uint64_t a_sb, a_sm, a_sk, b_sb, b_sk, b_sn, o_sb, o_sm, o_sn;
  uint32_t M, N, K, batch, in_dtype, out_dtype;
  uint64_t bias;
  float alpha;
  uint32_t epi_act, epi_on, b_dtype_p1;
};
struct SplitParams {
  uint64_t in, out, batch, rows, cols, in_bs, in_rs, out_rs;
};
struct XgpuParams {
  uint64_t mailbox[8];
  uint32_t rank, nranks, epoch, pad;
  uint64_t index_offset;
};
// One 256-byte slot block (value + index words, two epoch parities, eight source ranks) per DEVICE SET, addressed by the
// set's device bitmask: overlapping sets ({0,1} and {0,1,2,3}) never share slots, and every rank derives the same offset.
static constexpr size_t kMailboxSetBytes = 256;
static constexpr size_t kMailboxBytes = 256 * kMailboxSetBytes;
static constexpr uint32_t kWsMaxBlocks = 4096;
static constexpr uint32_t kWsTicketOffset = kWsMaxBlocks * 4 + kWsMaxBlocks * 8;
static constexpr uint32_t kWsDebugOffset = kWsTicketOffset + 64;         // four u64 words written by the reduce grid stage on request
static constexpr uint32_t kWsGemmTicketOffset = kWsTicketOffset + 256;  // u32 per (tail tile, CTA rank) of a split GEMM
static constexpr uint32_t kWsGemmTickets = 1024;                             // u32 entries
static constexpr uint32_t kWsColTicketOffset = kWsGemmTicketOffset + kWsGemmTickets * 4;  // u32[1024]: one per (outer, column tile) of a fused split column reduction
static constexpr uint32_t kWsColTickets = 1024;
static constexpr size_t kWsBytes = kWsColTicketOffset + kWsColTickets * 4;

// ================================================================================================ context
struct PoolBlock {
  size_t size;
  bool in_use;
  // Stream-ordered reuse (the reference's pools are per-stream and cross-stream hand-offs wait on events,
  // cubecl-runtime/src/stream/event.rs:50-57): a freed page remembers the stream that last used it and an event recorded
  // there; the same stream may take it back at once, any other requester only once the event has completed.
  CUstream last_stream = nullptr;
  CUevent done = nullptr;
  bool pending = false;
};

struct CommState {
  ncclComm_t comm = nullptr;
  int rank = -1, n = 0;
};

// Peer-memory exchange state for one device set (fused reduce + all-reduce over NVLink).
struct P2PState {
  uint64_t mailbox[8] = {0};
  int rank = -1, n = 0;
  uint32_t epoch = 0;
  std::vector<CUdeviceptr> opened;  // IPC mappings to close
};

struct b200_ctx {
  int device = -1;
  CUdevice dev{};
  CUcontext cuctx = nullptr;
  b200_props props{};
  CUstream stream = nullptr;       // default compute stream
  CUstream comm_stream = nullptr;  // dedicated NCCL stream (server.rs:944)
  CUevent comm_event = nullptr;
  std::vector<CUmodule> modules;
  std::unordered_map<std::string, CUfunction> funcs;
  // exclusive-page pool: exact-size free lists
  std::map<size_t, std::vector<CUdeviceptr>> free_lists;
  std::unordered_map<CUdeviceptr, PoolBlock> blocks;
  uint64_t bytes_in_use = 0, bytes_reserved = 0;
  std::unordered_map<void*, size_t> pinned;
  std::unordered_map<CUstream, CUdeviceptr> reduce_ws;
  std::vector<CUstream> dead_streams;   // streams destroyed through b200_stream_destroy (drained there): never record on them again
  std::map<std::string, CUtensorMap> tmap_cache;
  std::map<std::vector<int>, CommState> comms;
  std::map<std::vector<int>, P2PState> p2p;
  CUdeviceptr mailbox = 0;
  std::unordered_map<std::string, std::string> options;
  std::unordered_map<std::string, std::string> env_cache;   // B200_<KEY> environment defaults, read once
  uint64_t launches = 0;
  // Dry-run planning context (no driver, no device): every launch / descriptor / pool request is RECORDED instead of
  // executed, so the host logic (validation, batch collapse, variant choice, split plans) is testable on a CPU box.
  // Mirrors the reference's DryRun mode (crates/cubecl-runtime/src/dry_run.rs:45,88,121).
  bool dry = false;
  std::string plan;
  std::string pending_kernel;
  std::string last_kernel;         // name of the most recently launched kernel (b200_last_kernel)
  // Programmatic dependent launch of back-to-back all-element reductions on the context's own stream (every kernel there
  // is launched by this library, so the predecessor is known): output pointer of the reduce_all launch that is the LAST
  // kernel queued on c->stream, 0 if the last kernel was anything else.
  uint64_t pdl_prev_out = 0;
  uint64_t fake_next = 0x7000000000ull;
};

static inline CUstream resolve_stream(b200_ctx* c, b200_stream s) { return s ? static_cast<CUstream>(s) : c->stream; }

#define CTX_ENTER(c)                                                  \
  if (!(c)) return fail(B200_ERR_INVALID_ARG, "null context");        \
  if (!(c)->dry) CU_CHECK(g_drv.cuCtxSetCurrent_p((c)->cuctx));
// entry points that talk to the driver directly (copies, streams, events, collectives) do not exist on a planning context
#define CTX_ENTER_DEVICE(c)                                           \
  CTX_ENTER(c);                                                       \
  if ((c)->dry) return fail(B200_ERR_UNSUPPORTED, "%s is not available on a dry-run planning context", __func__);

static std::string opt(b200_ctx* c, const char* key, const char* dflt) {
  auto it = c->options.find(key);
  if (it != c->options.end()) return it->second;
  // environment fallback B200_<KEY>, looked up once per key (getenv on every launch is measurable next to a 20 us kernel)
  auto ev = c->env_cache.find(key);
  if (ev == c->env_cache.end()) {
    std::string env = std::string("B200_") + key;
    for (auto& ch : env) ch = (ch == '.') ? '_' : static_cast<char>(toupper(ch));
    const char* e = getenv(env.c_str());
    ev = c->env_cache.emplace(key, e ? std::string(e) : std::string("\x01unset")).first;
  }
  if (ev->second != "\x01unset") return ev->second;
  return dflt;
}

static int load_module(b200_ctx* c, const unsigned char* begin, const unsigned char* end, const char* what) {
  if (end <= begin) return fail(B200_ERR_COMPILATION, "embedded cubin '%s' is empty (library was built without kernels)", what);
  CUmodule m;
  CUresult r = g_drv.cuModuleLoadData_p(&m, begin);
  if (r != CUDA_SUCCESS)
    return fail(B200_ERR_COMPILATION, "cuModuleLoadData(%s) failed: %s -- the cubins are sm_100a only", what, cu_err(r));
  c->modules.push_back(m);
  return B200_OK;
}

static int get_func(b200_ctx* c, const std::string& name, CUfunction* out) {
  c->last_kernel = name;
  if (c->dry) { c->pending_kernel = name; *out = nullptr; return B200_OK; }
  auto it = c->funcs.find(name);
  if (it != c->funcs.end()) { *out = it->second; return B200_OK; }
  // modules are loaded in the order gemm, reduce, aux, gemm_mx, gemm_b, gemm_c; the kernel name says where a kernel lives (no
  // failing lookups, which API-level tools such as compute-sanitizer would report)
  auto starts = [&](const char* pfx) { return name.rfind(pfx, 0) == 0; };
  auto has = [&](const char* part) { return name.find(part) != std::string::npos; };
  const bool mx = starts("gemm_mxf8_") || starts("gemm_mxf4_") || starts("gemm_nvf4_") || starts("umma_probe_e4m3") ||
                  starts("umma_probe_mxf8") || starts("umma_probe_mxf4");
  const bool tc_gemm = starts("gemm_") && name != "gemm_simt_strided" && name != "gemm_scaled_simt";
  const size_t home = mx ? 3
                      : tc_gemm && has("_2sm_m512_") ? 5
                      : tc_gemm && (has("_2sm_n128_") || has("_1sm_n128_")) ? 4
                      : tc_gemm || starts("umma_") ? 0
                      : starts("reduce_") ? 1 : 2;
  if (home < c->modules.size()) {
    CUfunction f;
    if (g_drv.cuModuleGetFunction_p(&f, c->modules[home], name.c_str()) == CUDA_SUCCESS) {
      c->funcs[name] = f;
      *out = f;
      return B200_OK;
    }
  }
  return fail(B200_ERR_COMPILATION, "kernel '%s' not found in the prebuilt cubins", name.c_str());
}

extern "C" int b200_get_cubin(const char* name, const void** image, size_t* size) {
  if (!name || !image || !size) return fail(B200_ERR_INVALID_ARG, "get_cubin: null argument");
  const unsigned char *b = nullptr, *e = nullptr;
  if (!strcmp(name, "gemm")) { b = b200_cubin_gemm; e = b200_cubin_gemm_end; }
  else if (!strcmp(name, "reduce")) { b = b200_cubin_reduce; e = b200_cubin_reduce_end; }
  else if (!strcmp(name, "aux")) { b = b200_cubin_aux; e = b200_cubin_aux_end; }
  else if (!strcmp(name, "gemm_mx")) { b = b200_cubin_gemm_mx; e = b200_cubin_gemm_mx_end; }
  else if (!strcmp(name, "gemm_b")) { b = b200_cubin_gemm_b; e = b200_cubin_gemm_b_end; }
  else if (!strcmp(name, "gemm_c")) { b = b200_cubin_gemm_c; e = b200_cubin_gemm_c_end; }
  else return fail(B200_ERR_INVALID_ARG, "get_cubin: unknown image '%s' (gemm|gemm_b|gemm_c|gemm_mx|reduce|aux)", name);
  *image = b;
  *size = static_cast<size_t>(e - b);
  return B200_OK;
}

extern "C" int b200_device_count(int* count) {
  if (!count) return fail(B200_ERR_INVALID_ARG, "null count");
  int rc = ensure_driver();
  if (rc) return rc;
  CU_CHECK(g_drv.cuDeviceGetCount_p(count));
  return B200_OK;
}

extern "C" int b200_init(int device, b200_ctx** out) {
  if (!out) return fail(B200_ERR_INVALID_ARG, "null out");
  *out = nullptr;
  int rc = ensure_driver();
  if (rc) return rc;
  int n = 0;
  CU_CHECK(g_drv.cuDeviceGetCount_p(&n));
  if (device < 0 || device >= n) return fail(B200_ERR_NO_DEVICE, "device %d out of range (%d devices)", device, n);
  b200_ctx* c = new b200_ctx();
  c->device = device;
  auto bail = [&](int code) { delete c; return code; };
  CUresult r;
  if ((r = g_drv.cuDeviceGet_p(&c->dev, device)) != CUDA_SUCCESS) return bail(fail(map_cu(r), "cuDeviceGet: %s", cu_err(r)));
  if ((r = g_drv.cuDevicePrimaryCtxRetain_p(&c->cuctx, c->dev)) != CUDA_SUCCESS)
    return bail(fail(map_cu(r), "cuDevicePrimaryCtxRetain: %s", cu_err(r)));
  if ((r = g_drv.cuCtxSetCurrent_p(c->cuctx)) != CUDA_SUCCESS) return bail(fail(map_cu(r), "cuCtxSetCurrent: %s", cu_err(r)));
  auto attr = [&](CUdevice_attribute a) { int v = 0; g_drv.cuDeviceGetAttribute_p(&v, a, c->dev); return v; };
  c->props.device = device;
  c->props.cc_major = attr(CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR);
  c->props.cc_minor = attr(CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR);
  c->props.num_sms = attr(CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT);
  c->props.max_shared_per_block = attr(CU_DEVICE_ATTRIBUTE_MAX_SHARED_MEMORY_PER_BLOCK_OPTIN);
  c->props.clock_khz = attr(CU_DEVICE_ATTRIBUTE_CLOCK_RATE);
  c->props.mem_clock_khz = attr(CU_DEVICE_ATTRIBUTE_MEMORY_CLOCK_RATE);
  c->props.plane_size = 32;
  size_t total = 0;
  g_drv.cuDeviceTotalMem_p(&total, c->dev);
  c->props.total_mem = total;
  g_drv.cuDeviceGetName_p(c->props.name, sizeof(c->props.name), c->dev);
  if (c->props.cc_major != 10) {
    g_drv.cuDevicePrimaryCtxRelease_p(c->dev);
    return bail(fail(B200_ERR_NO_DEVICE, "device %d is sm_%d%d; this library ships sm_100a cubins only (B200)", device,
                     c->props.cc_major, c->props.cc_minor));
  }
  if ((rc = load_module(c, b200_cubin_gemm, b200_cubin_gemm_end, "gemm")) ||
      (rc = load_module(c, b200_cubin_reduce, b200_cubin_reduce_end, "reduce")) ||
      (rc = load_module(c, b200_cubin_aux, b200_cubin_aux_end, "aux")) ||
      (rc = load_module(c, b200_cubin_gemm_mx, b200_cubin_gemm_mx_end, "gemm_mx")) ||
      (rc = load_module(c, b200_cubin_gemm_b, b200_cubin_gemm_b_end, "gemm_b")) ||
      (rc = load_module(c, b200_cubin_gemm_c, b200_cubin_gemm_c_end, "gemm_c"))) {
    for (CUmodule m : c->modules) g_drv.cuModuleUnload_p(m);
    g_drv.cuDevicePrimaryCtxRelease_p(c->dev);
    return bail(rc);
  }
  if ((r = g_drv.cuStreamCreate_p(&c->stream, CU_STREAM_NON_BLOCKING)) != CUDA_SUCCESS ||
      (r = g_drv.cuStreamCreate_p(&c->comm_stream, CU_STREAM_NON_BLOCKING)) != CUDA_SUCCESS ||
      (r = g_drv.cuEventCreate_p(&c->comm_event, CU_EVENT_DISABLE_TIMING)) != CUDA_SUCCESS) {
    // release what exists: streams created so far, the four loaded modules, the retained primary context
    if (c->comm_stream) g_drv.cuStreamDestroy_p(c->comm_stream);
    if (c->stream) g_drv.cuStreamDestroy_p(c->stream);
    for (CUmodule m : c->modules) g_drv.cuModuleUnload_p(m);
    g_drv.cuDevicePrimaryCtxRelease_p(c->dev);
    return bail(fail(map_cu(r), "stream/event creation failed: %s", cu_err(r)));
  }
  *out = c;
  return B200_OK;
}

extern "C" int b200_destroy(b200_ctx* c) {
  if (!c) return B200_OK;
  if (c->dry) { delete c; return B200_OK; }
  if (g_drv.ok && g_drv.cuCtxSetCurrent_p(c->cuctx) == CUDA_SUCCESS) {
    g_drv.cuCtxSynchronize_p();
    for (auto& kv : c->comms)
      if (kv.second.comm && g_nccl.ok) g_nccl.CommDestroy(kv.second.comm);
    for (auto& kv : c->p2p)
      for (CUdeviceptr q : kv.second.opened) g_drv.cuIpcCloseMemHandle_p(q);
    if (c->mailbox) g_drv.cuMemFree_p(c->mailbox);
    for (auto& kv : c->blocks) {
      if (kv.second.done) g_drv.cuEventDestroy_p(kv.second.done);
      g_drv.cuMemFree_p(kv.first);
    }
    for (auto& kv : c->reduce_ws) g_drv.cuMemFree_p(kv.second);
    for (auto& kv : c->pinned) g_drv.cuMemFreeHost_p(kv.first);
    if (c->comm_event) g_drv.cuEventDestroy_p(c->comm_event);
    if (c->comm_stream) g_drv.cuStreamDestroy_p(c->comm_stream);
    if (c->stream) g_drv.cuStreamDestroy_p(c->stream);
    for (CUmodule m : c->modules) g_drv.cuModuleUnload_p(m);
    g_drv.cuDevicePrimaryCtxRelease_p(c->dev);
  }
  delete c;
  return B200_OK;
}

extern "C" int b200_plan_begin(int num_sms, b200_ctx** out) {
  if (!out || num_sms < 1) return fail(B200_ERR_INVALID_ARG, "plan_begin: bad arguments");
  b200_ctx* c = new b200_ctx();
  c->dry = true;
  c->device = 0;
  c->props.num_sms = num_sms;
  c->props.cc_major = 10;
  c->props.plane_size = 32;
  snprintf(c->props.name, sizeof(c->props.name), "dry-run (%d SMs)", num_sms);
  *out = c;
  return B200_OK;
}

extern "C" int b200_plan_text(b200_ctx* c, char* buf, size_t capacity, size_t* needed) {
  if (!c || !c->dry) return fail(B200_ERR_INVALID_ARG, "plan_text: not a planning context");
  if (needed) *needed = c->plan.size() + 1;
  if (buf && capacity) {
    const size_t n = std::min(capacity - 1, c->plan.size());
    memcpy(buf, c->plan.data(), n);
    buf[n] = 0;
    if (capacity > c->plan.size()) c->plan.clear();
  }
  return B200_OK;
}

extern "C" int b200_get_props(b200_ctx* c, b200_props* out) {
  if (!c || !out) return fail(B200_ERR_INVALID_ARG, "null argument");
  *out = c->props;
  return B200_OK;
}

extern "C" int b200_set_option(b200_ctx* c, const char* key, const char* value) {
  if (!c || !key || !value) return fail(B200_ERR_INVALID_ARG, "null argument");
  static const char* known[] = {"gemm.variant", "gemm.f32", "gemm.group_m", "gemm.split_k", "gemm.epilogue", "gemm.l2_promotion", "gemm.stage", "gemm.sf_copy", "reduce.variant", "reduce.threads",
                                "reduce.blocks_per_sm", "reduce.rows_vpt", "reduce.rows_blocks_per_sm", "reduce.cols_blocks_per_sm",
                                "reduce.debug", "reduce.pdl", "reduce.tma_stages", "reduce.tma_ctas_per_sm", "reduce.cols_split_target", "reduce.cols_loads", "reduce.cols_fused"};
  for (const char* k : known)
    if (!strcmp(k, key)) { c->options[key] = value; return B200_OK; }
  return fail(B200_ERR_INVALID_ARG, "unknown option '%s'", key);
}

extern "C" int b200_last_kernel(b200_ctx* c, char* buf, size_t capacity) {
  if (!c || !buf || capacity == 0) return fail(B200_ERR_INVALID_ARG, "last_kernel: bad arguments");
  snprintf(buf, capacity, "%s", c->last_kernel.c_str());
  return B200_OK;
}

extern "C" int b200_launch_count(b200_ctx* c, uint64_t* count) {
  if (!c || !count) return fail(B200_ERR_INVALID_ARG, "null argument");
  *count = c->launches;
  return B200_OK;
}

// ================================================================================================ memory
static size_t pool_round(size_t bytes) {
  if (bytes == 0) bytes = 1;
  const size_t align = bytes >= (64u << 20) ? (2u << 20) : bytes >= (1u << 20) ? (64u << 10) : 512;
  return (bytes + align - 1) / align * align;
}

static int pool_alloc(b200_ctx* c, size_t bytes, CUdeviceptr* out, CUstream for_stream = nullptr) {
  const size_t sz = pool_round(bytes);
  if (c->dry) {
    *out = c->fake_next;
    c->fake_next += (sz + 511) / 512 * 512;
    char line[96];
    snprintf(line, sizeof(line), "alloc %zu\n", sz);
    c->plan += line;
    return B200_OK;
  }
  auto it = c->free_lists.find(sz);
  if (it != c->free_lists.end()) {
    std::vector<CUdeviceptr>& fl = it->second;
    for (size_t i = fl.size(); i-- > 0;) {
      PoolBlock& b = c->blocks[fl[i]];
      bool usable = !b.pending || (for_stream != nullptr && b.last_stream == for_stream);
      if (!usable && b.done && g_drv.cuEventQuery_p(b.done) == CUDA_SUCCESS) { b.pending = false; usable = true; }
      if (!usable) continue;  // still in flight on another stream: leave it for later
      *out = fl[i];
      fl.erase(fl.begin() + static_cast<long>(i));
      b.in_use = true;
      c->bytes_in_use += sz;
      return B200_OK;
    }
  }
  CUdeviceptr p = 0;
  CUresult r = g_drv.cuMemAlloc_p(&p, sz);
  if (r == CUDA_ERROR_OUT_OF_MEMORY) {
    // release cached pages and retry once (memory_cleanup semantics); cuMemFree synchronises, so pending pages are safe
    for (auto& fl : c->free_lists) {
      for (CUdeviceptr q : fl.second) {
        PoolBlock& b = c->blocks[q];
        if (b.done) g_drv.cuEventDestroy_p(b.done);
        g_drv.cuMemFree_p(q);
        c->bytes_reserved -= fl.first;
        c->blocks.erase(q);
      }
      fl.second.clear();
    }
    r = g_drv.cuMemAlloc_p(&p, sz);
  }
  if (r != CUDA_SUCCESS) return fail(map_cu(r), "cuMemAlloc(%zu bytes) failed: %s", sz, cu_err(r));
  PoolBlock nb;
  nb.size = sz;
  nb.in_use = true;
  c->blocks[p] = nb;
  c->bytes_reserved += sz;
  c->bytes_in_use += sz;
  *out = p;
  return B200_OK;
}

// `used_on`: the stream whose queued work may still touch the page (nullptr = the context's compute stream).
static int pool_free(b200_ctx* c, CUdeviceptr p, CUstream used_on = nullptr) {
  if (c->dry) return B200_OK;
  auto it = c->blocks.find(p);
  if (it == c->blocks.end() || !it->second.in_use) return fail(B200_ERR_INVALID_ARG, "b200_free: pointer not owned by this context");
  PoolBlock& b = it->second;
  b.in_use = false;
  // a stream this context already destroyed was synchronised at that point: the page is idle as far as it is concerned
  if (used_on && std::find(c->dead_streams.begin(), c->dead_streams.end(), used_on) != c->dead_streams.end()) used_on = nullptr;
  b.last_stream = used_on ? used_on : c->stream;
  if (!b.done && g_drv.cuEventCreate_p(&b.done, CU_EVENT_DISABLE_TIMING) != CUDA_SUCCESS) b.done = nullptr;
  b.pending = (b.done != nullptr) && g_drv.cuEventRecord_p(b.done, b.last_stream) == CUDA_SUCCESS;
  c->bytes_in_use -= b.size;
  c->free_lists[b.size].push_back(p);
  return B200_OK;
}

extern "C" int b200_alloc(b200_ctx* c, size_t bytes, b200_dptr* out) {
  CTX_ENTER(c);
  if (!out) return fail(B200_ERR_INVALID_ARG, "null out");
  CUdeviceptr p;
  int rc = pool_alloc(c, bytes, &p);
  if (rc) return rc;
  *out = static_cast<b200_dptr>(p);
  return B200_OK;
}

extern "C" int b200_free(b200_ctx* c, b200_dptr ptr) {
  CTX_ENTER(c);
  return pool_free(c, static_cast<CUdeviceptr>(ptr));
}

// The stream-ordered form: `last_use` is the stream whose queued work may still touch the buffer (NULL = the context's
// stream).  The page is handed out again only once an event recorded there has completed (or to that same stream).
extern "C" int b200_free_async(b200_ctx* c, b200_dptr ptr, b200_stream last_use) {
  CTX_ENTER(c);
  return pool_free(c, static_cast<CUdeviceptr>(ptr), resolve_stream(c, last_use));
}

extern "C" int b200_memory_usage(b200_ctx* c, uint64_t* in_use, uint64_t* reserved) {
  if (!c) return fail(B200_ERR_INVALID_ARG, "null context");
  if (in_use) *in_use = c->bytes_in_use;
  if (reserved) *reserved = c->bytes_reserved;
  return B200_OK;
}

extern "C" int b200_memory_cleanup(b200_ctx* c) {
  CTX_ENTER_DEVICE(c);
  CU_CHECK(g_drv.cuCtxSynchronize_p());
  for (auto& fl : c->free_lists) {
    for (CUdeviceptr q : fl.second) {
      PoolBlock& b = c->blocks[q];
      if (b.done) g_drv.cuEventDestroy_p(b.done);
      g_drv.cuMemFree_p(q);
      c->bytes_reserved -= fl.first;
      c->blocks.erase(q);
    }
    fl.second.clear();
  }
  return B200_OK;
}

extern "C" int b200_host_alloc(b200_ctx* c, size_t bytes, void** out) {
  CTX_ENTER_DEVICE(c);
  if (!out) return fail(B200_ERR_INVALID_ARG, "null out");
  void* p = nullptr;
  CU_CHECK(g_drv.cuMemAllocHost_p(&p, bytes ? bytes : 1));
  c->pinned[p] = bytes;
  *out = p;
  return B200_OK;
}

extern "C" int b200_host_free(b200_ctx* c, void* ptr) {
  CTX_ENTER_DEVICE(c);
  auto it = c->pinned.find(ptr);
  if (it == c->pinned.end()) return fail(B200_ERR_INVALID_ARG, "b200_host_free: pointer not owned by this context");
  CU_CHECK(g_drv.cuMemFreeHost_p(ptr));
  c->pinned.erase(it);
  return B200_OK;
}

extern "C" int b200_write(b200_ctx* c, b200_stream s, b200_dptr dst, const void* src, size_t bytes) {
  CTX_ENTER_DEVICE(c);
  if (bytes == 0) return B200_OK;
  CU_CHECK(g_drv.cuMemcpyHtoDAsync_p(static_cast<CUdeviceptr>(dst), src, bytes, resolve_stream(c, s)));
  return B200_OK;
}

extern "C" int b200_read(b200_ctx* c, b200_stream s, void* dst, b200_dptr src, size_t bytes) {
  CTX_ENTER_DEVICE(c);
  if (bytes == 0) return B200_OK;
  CU_CHECK(g_drv.cuMemcpyDtoHAsync_p(dst, static_cast<CUdeviceptr>(src), bytes, resolve_stream(c, s)));
  return B200_OK;
}

extern "C" int b200_copy(b200_ctx* c, b200_stream s, b200_dptr dst, b200_dptr src, size_t bytes) {
  CTX_ENTER_DEVICE(c);
  if (bytes == 0) return B200_OK;
  CU_CHECK(g_drv.cuMemcpyDtoDAsync_p(static_cast<CUdeviceptr>(dst), static_cast<CUdeviceptr>(src), bytes, resolve_stream(c, s)));
  return B200_OK;
}

extern "C" int b200_memset32(b200_ctx* c, b200_stream s, b200_dptr dst, uint32_t value, size_t words) {
  CTX_ENTER_DEVICE(c);
  if (words == 0) return B200_OK;
  CU_CHECK(g_drv.cuMemsetD32Async_p(static_cast<CUdeviceptr>(dst), value, words, resolve_stream(c, s)));
  return B200_OK;
}

// ================================================================================================ streams / events
extern "C" int b200_stream_create(b200_ctx* c, b200_stream* out) {
  CTX_ENTER_DEVICE(c);
  if (!out) return fail(B200_ERR_INVALID_ARG, "null out");
  CUstream s;
  CU_CHECK(g_drv.cuStreamCreate_p(&s, CU_STREAM_NON_BLOCKING));
  c->dead_streams.erase(std::remove(c->dead_streams.begin(), c->dead_streams.end(), s), c->dead_streams.end());  // the handle value may be recycled
  *out = s;
  return B200_OK;
}

extern "C" int b200_stream_destroy(b200_ctx* c, b200_stream s) {
  CTX_ENTER_DEVICE(c);
  if (!s) return B200_OK;
  CU_CHECK(g_drv.cuStreamSynchronize_p(static_cast<CUstream>(s)));   // pages freed on this stream are idle from here on
  for (auto& kv : c->blocks)
    if (kv.second.last_stream == static_cast<CUstream>(s)) { kv.second.last_stream = nullptr; kv.second.pending = false; }
  auto it = c->reduce_ws.find(static_cast<CUstream>(s));
  if (it != c->reduce_ws.end()) { g_drv.cuMemFree_p(it->second); c->reduce_ws.erase(it); }
  CU_CHECK(g_drv.cuStreamDestroy_p(static_cast<CUstream>(s)));
  c->dead_streams.push_back(static_cast<CUstream>(s));
  return B200_OK;
}

extern "C" int b200_sync(b200_ctx* c, b200_stream s) {
  CTX_ENTER(c);
  if (c->dry) return B200_OK;
  CUresult r = g_drv.cuStreamSynchronize_p(resolve_stream(c, s));
  if (r != CUDA_SUCCESS) return fail(B200_ERR_UNHEALTHY, "stream sync surfaced a device fault: %s (%d)", cu_err(r), (int)r);
  return B200_OK;
}

extern "C" int b200_event_create(b200_ctx* c, b200_event* out) {
  CTX_ENTER_DEVICE(c);
  if (!out) return fail(B200_ERR_INVALID_ARG, "null out");
  CUevent e;
  CU_CHECK(g_drv.cuEventCreate_p(&e, CU_EVENT_DEFAULT));
  *out = e;
  return B200_OK;
}

extern "C" int b200_event_record(b200_ctx* c, b200_event e, b200_stream s) {
  CTX_ENTER_DEVICE(c);
  CU_CHECK(g_drv.cuEventRecord_p(static_cast<CUevent>(e), resolve_stream(c, s)));
  return B200_OK;
}

extern "C" int b200_stream_wait_event(b200_ctx* c, b200_stream s, b200_event e) {
  CTX_ENTER_DEVICE(c);
  CU_CHECK(g_drv.cuStreamWaitEvent_p(resolve_stream(c, s), static_cast<CUevent>(e), 0));
  return B200_OK;
}

extern "C" int b200_event_elapsed_ms(b200_ctx* c, b200_event a, b200_event b, float* ms) {
  CTX_ENTER_DEVICE(c);
  if (!ms) return fail(B200_ERR_INVALID_ARG, "null ms");
  CU_CHECK(g_drv.cuEventSynchronize_p(static_cast<CUevent>(b)));
  CU_CHECK(g_drv.cuEventElapsedTime_p(ms, static_cast<CUevent>(a), static_cast<CUevent>(b)));
  return B200_OK;
}

extern "C" int b200_event_destroy(b200_ctx* c, b200_event e) {
  CTX_ENTER_DEVICE(c);
  if (e) CU_CHECK(g_drv.cuEventDestroy_p(static_cast<CUevent>(e)));
  return B200_OK;
}

// ================================================================================================ launch helper
static int launch(b200_ctx* c, CUfunction f, unsigned grid_x, unsigned grid_y, unsigned grid_z, unsigned block,
                  unsigned smem, unsigned cluster_x, CUstream st, void** args, bool pdl = false) {
  CUlaunchConfig cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDimX = grid_x; cfg.gridDimY = grid_y; cfg.gridDimZ = grid_z;
  cfg.blockDimX = block; cfg.blockDimY = 1; cfg.blockDimZ = 1;
  cfg.sharedMemBytes = smem;
  cfg.hStream = st;
  CUlaunchAttribute at[2];
  unsigned nat = 0;
  if (cluster_x > 1) {
    at[nat].id = CU_LAUNCH_ATTRIBUTE_CLUSTER_DIMENSION;
    at[nat].value.clusterDim.x = cluster_x; at[nat].value.clusterDim.y = 1; at[nat].value.clusterDim.z = 1;
    ++nat;
  }
  if (pdl) {  // this kernel may begin once every block of the preceding kernel has triggered (griddepcontrol) or exited
    at[nat].id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
    at[nat].value.programmaticStreamSerializationAllowed = 1;
    ++nat;
  }
  if (nat) { cfg.attrs = at; cfg.numAttrs = nat; }
  if (st == c->stream) c->pdl_prev_out = 0;   // whoever launches a reduce_all sets it again after this call
  if (c->dry) {
    char line[256];
    snprintf(line, sizeof(line), "launch %s grid=(%u,%u,%u) block=%u smem=%u cluster=%u%s\n", c->pending_kernel.c_str(), grid_x,
             grid_y, grid_z, block, smem, cluster_x, pdl ? " pdl" : "");
    c->plan += line;
    c->launches++;
    return B200_OK;
  }
  CUresult r = g_drv.cuLaunchKernelEx_p(&cfg, f, args, nullptr);
  if (r != CUDA_SUCCESS) return fail(map_cu(r), "cuLaunchKernelEx failed: %s (%d)", cu_err(r), (int)r);
  c->launches++;
  return B200_OK;
}

static size_t dtype_size(int dt) {
  switch (dt) {
    case B200_F32: case B200_U32: case B200_I32: return 4;
    case B200_F16: case B200_BF16: return 2;
    case B200_F64: case B200_I64: case B200_U64: return 8;
    case B200_U8: case B200_I8: case B200_F8E4M3: case B200_F8E5M2: return 1;
    default: return 0;
  }
}

// ================================================================================================ matmul
struct GemmVariant {
  const char* tag;  // suffix in the kernel name
  int cg, block_n, stages;
  double eff;       // measured MMA-pipe efficiency relative to 2sm_n256 (8192^3, B200): the N=128 shapes need
                    // 128 B/cycle/SM of operand reads from shared memory and are smem-bandwidth bound.
                    // 0 = never chosen automatically (gemm.variant=<tag> only)
  int mt;           // 128-row sub-tiles of M per CTA: the pair tile is (128 * cg * mt) x block_n
};
static const GemmVariant kVariants[] = {{"2sm_n256", 2, 256, 6, 1.0, 1}, {"2sm_n128", 2, 128, 8, 0.66, 1}, {"1sm_n128", 1, 128, 6, 0.59, 1},
                                        // 512 x 256 pair tile, one accumulator stage, 384 threads (gemm_tcgen05.cu, MT = 2).  Measured
                                        // bf16 8192^3, same box, power state equalised (profiles/r01_pair_tile_ab.log): x1.059 of
                                        // 2sm_n256 in a 50-launch burst and x1.058 held for 1 s (25 % less L2->SM operand traffic
                                        // -> 1.53 instead of 1.45 GHz under the power cap), 0.98 of cuBLAS; fp8 e4m3 8192^3: x1.04
                                        // (3223 vs 3098 TFLOP/s, profiles/r01_bench_n1.json tile_variants_8192)
                                        {"2sm_m512", 2, 256, 4, 1.06, 2},
                                        // block-scaled kinds only: 256 x 224 tile -- two accumulator stages and the scale columns fit
                                        // the 512 TMEM columns, so the epilogue hides behind the next tile (the 256-wide scaled tiles
                                        // hold ONE stage).  Needs the B scales packed per 224-row tile, i.e. row-major scales.
                                        // MEASURED SLOWER (profiles/r02_block_scaled_sweep.log, 8192^3 -> bf16): mxfp8 2046 vs 2448 TFLOP/s,
                                        // mxfp4 3798 vs 3866 -- UMMA N = 224 issues at the N = 256 rate, which costs more than the hidden
                                        // epilogue returns.  eff 0: never chosen automatically (gemm.variant=2sm_n224 only).
                                        {"2sm_n224", 2, 224, 6, 0.0, 1},
                                        // diagnostic: 256 x 256 tile with ONE accumulator stage (bf16 -> bf16, K-major lhs only)
                                        {"2sm_n256a1", 2, 256, 6, 0.0, 1}};
static bool variant_has_dtype(const GemmVariant& v, int in_dtype) {
  if (!strcmp(v.tag, "2sm_n224")) return false;   // block-scaled kinds only
  const bool bits8 = (in_dtype == B200_F8E4M3 || in_dtype == B200_F8E5M2 || in_dtype == B200_U8 || in_dtype == B200_I8);
  if (!strcmp(v.tag, "2sm_m512")) return in_dtype == B200_BF16 || in_dtype == B200_F16 || in_dtype == B200_F8E4M3 || in_dtype == B200_F8E5M2;
  if (!strcmp(v.tag, "2sm_n256a1")) return in_dtype == B200_BF16 || in_dtype == B200_F8E4M3;
  if (bits8) return !strcmp(v.tag, "2sm_n256") || !strcmp(v.tag, "1sm_n128");
  return true;
}

// mx_kind: 0 = unscaled, 1 = mxf8 (one 512-byte scale chunk per 128 rows per k-block), 2 = mxf4 (two), 3 = nvfp4 (four)
static unsigned mx_atoms(int mx_kind) { return mx_kind == 3 ? 4u : (unsigned)mx_kind; }
// scale atoms of one stage (512 bytes each: A rows of the CTA, B rows of the whole tile), padded to the 1 KB stage alignment
static unsigned gemm_sf_stage_bytes(const GemmVariant& v, int mx_kind) {
  if (!mx_kind) return 0;
  return (512u * mx_atoms(mx_kind) * (1 + (v.block_n + 127) / 128) + 1023u) / 1024u * 1024u;
}
// block-scaled kinds: as many stages as fit 227 KB, eight at most (mirror of mx_stages() in gemm_tcgen05.cu)
static int gemm_stages(const GemmVariant& v, int mx_kind) {
  if (!mx_kind) return v.stages;
  const int stage = 16384 + (v.block_n / v.cg) * 128 + (int)gemm_sf_stage_bytes(v, mx_kind);
  return std::min(8, (232448 - 1024 - 1024 - 16384) / stage);
}
// alignment slack + operand ring (+ scale chunks) + barrier block + epilogue staging (4 * mt warps x [32 rows x 128 B])
static unsigned gemm_smem_bytes(const GemmVariant& v, int mx_kind = 0) {
  return 1024 + gemm_stages(v, mx_kind) * (16384 * v.mt + (v.block_n / v.cg) * 128 + gemm_sf_stage_bytes(v, mx_kind)) + 1024 + 16384 * v.mt;
}

static int encode_tmap(b200_ctx* c, CUtensorMap* out, CUtensorMapDataType dt, size_t esz, uint64_t base, uint64_t d0,
                       uint64_t d1, uint64_t d2, uint64_t s1_elems, uint64_t s2_elems, uint32_t b0, uint32_t b1,
                       CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, uint32_t b2 = 1) {
  // L2 promotion: how much of a line's neighbourhood a TMA load pulls into L2 (tuning knob; 256 B measured best so far)
  const int promo_bytes = atoi(opt(c, "gemm.l2_promotion", "256").c_str());
  const CUtensorMapL2promotion promo = promo_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                       : promo_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                       : promo_bytes >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  char key[256];
  snprintf(key, sizeof(key), "%d|%d|%llx|%llu|%llu|%llu|%llu|%llu|%u|%u|%u|%d", (int)dt, (int)swz, (unsigned long long)base,
           (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1_elems,
           (unsigned long long)s2_elems, b0, b1, b2, (int)promo);
  if (c->dry) {
    char line[256];
    if (b2 == 1)
      snprintf(line, sizeof(line), "tmap esz=%zu dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u) swizzle=%d\n", esz,
               (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)(s1_elems * esz),
               (unsigned long long)(s2_elems * esz), b0, b1, (int)swz);
    else
      snprintf(line, sizeof(line), "tmap esz=%zu dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u,%u) swizzle=%d\n", esz,
               (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)(s1_elems * esz),
               (unsigned long long)(s2_elems * esz), b0, b1, b2, (int)swz);
    c->plan += line;
    memset(out, 0, sizeof(*out));
    return B200_OK;
  }
  auto it = c->tmap_cache.find(key);
  if (it != c->tmap_cache.end()) { *out = it->second; return B200_OK; }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_elems * esz, s2_elems * esz};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_drv.cuTensorMapEncodeTiled_p(out, dt, 3, reinterpret_cast<void*>(base), dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(B200_ERR_INVALID_ARG, "cuTensorMapEncodeTiled failed: %s (dims %llu,%llu,%llu strides %llu,%llu box %u,%u)",
                cu_err(r), (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                (unsigned long long)strides[0], (unsigned long long)strides[1], b0, b1);
  if (c->tmap_cache.size() > 512) c->tmap_cache.clear();
  c->tmap_cache[key] = *out;
  return B200_OK;
}

// Scale-factor tensor map: the packed tensor [tiles][k atoms][512 B] viewed as (one 512-byte atom = 128 words, k atoms, tiles);
// a box of (128, atoms, tiles) lands in shared memory as [tile][atom][512 B].  Whole atoms are the box rows: round 2 used
// (16 B, 32 rows x atoms, tiles) boxes, whose 16-byte rows cost the TMA unit far more per byte.
static int encode_sf_tmap(b200_ctx* c, CUtensorMap* out, uint64_t base, uint64_t k_atoms, uint64_t tiles, uint32_t box_atoms, uint32_t box_tiles) {
  char key[256];
  snprintf(key, sizeof(key), "sf|%llx|%llu|%llu|%u|%u", (unsigned long long)base, (unsigned long long)k_atoms, (unsigned long long)tiles, box_atoms, box_tiles);
  if (c->dry) {
    char line[256];
    snprintf(line, sizeof(line), "tmap scales esz=4 dims=(128,%llu,%llu) strides=(512,%llu) box=(128,%u,%u)\n", (unsigned long long)k_atoms,
             (unsigned long long)tiles, (unsigned long long)(512 * k_atoms), box_atoms, box_tiles);
    c->plan += line;
    memset(out, 0, sizeof(*out));
    return B200_OK;
  }
  auto it = c->tmap_cache.find(key);
  if (it != c->tmap_cache.end()) { *out = it->second; return B200_OK; }
  cuuint64_t dims[3] = {128, k_atoms, tiles};
  cuuint64_t strides[2] = {512, 512 * k_atoms};
  cuuint32_t box[3] = {128, box_atoms, box_tiles};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_drv.cuTensorMapEncodeTiled_p(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, reinterpret_cast<void*>(base), dims, strides, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(B200_ERR_INVALID_ARG, "cuTensorMapEncodeTiled (scales) failed: %s", cu_err(r));
  if (c->tmap_cache.size() > 512) c->tmap_cache.clear();
  c->tmap_cache[key] = *out;
  return B200_OK;
}

// One batched problem with LINEAR batch strides (0 = broadcast).  Strides in elements.
struct GemmProblem {
  int in_dtype, out_dtype;
  int rhs_dtype = -1;           // >= 0: a mixed 8-bit pair (fp8 e4m3 x e5m2, u8 x i8); in_dtype is then the lhs format
  uint64_t a, b, out;
  uint64_t a_lo = 0, b_lo = 0;  // 3xTF32: compact low parts (same logical layout class as a / b), 0 otherwise
  bool hybrid = false;          // a_lo / b_lo are bf16 PAIR buffers [2][entries][rows][pitch]: bf16(x) planes, then bf16(x - trunc_tf32(x))
  uint64_t bias = 0;            // fused epilogue: out = act(alpha * acc + bias[n])
  float alpha = 1.0f;
  uint32_t act = 0;
  uint64_t M, N, K, batch;
  uint64_t a_sm, a_sk, a_sb;
  uint64_t b_sk, b_sn, b_sb;
  uint64_t o_sm, o_sn, o_sb;
  // block-scaled (MX) problems: in_dtype is a 1-byte marker, K / a_sm / b_sn count BYTES of K-major packed operands
  int mx_kind = 0;                 // 0 unscaled, 1 mxf8 (e4m3 / e5m2), 2 mxf4 (packed e2m1, ue8m0 / 32), 3 nvfp4 (packed e2m1, ue4m3 / 16)
  uint32_t fmt_a = 0, fmt_b = 0;   // instruction-descriptor operand formats
  uint64_t sfa = 0, sfb = 0;       // packed scale tensors [batch * tiles][atoms][512 B]
  uint64_t sf_atoms = 0;           // 4-scale atoms along K
  int sfb_tile_rows = 128;         // how the rhs scales are packed: per 128-row chunk, or per 224-row GEMM tile (2sm_n224 only)
  bool sfb_any_layout = false;     // planning pass: the caller packs the scales AFTER the variant is known
};

static int launch_simt(b200_ctx* c, CUstream st, const GemmProblem& g) {
  CUfunction f;
  int rc = get_func(c, "gemm_simt_strided", &f);
  if (rc) return rc;
  if (g.batch > 65535) return fail(B200_ERR_UNSUPPORTED, "simt matmul: batch %llu > 65535", (unsigned long long)g.batch);
  if (g.M >= (1ull << 32) || g.N >= (1ull << 32) || g.K >= (1ull << 32))
    return fail(B200_ERR_UNSUPPORTED, "simt matmul: extents must fit 32 bits (M=%llu N=%llu K=%llu)", (unsigned long long)g.M,
                (unsigned long long)g.N, (unsigned long long)g.K);
  const uint32_t epi_on = (g.alpha != 1.0f || g.bias != 0 || g.act != 0) ? 1u : 0u;
  SimtGemmParams p{g.a, g.b, g.out, g.a_sb, g.a_sm, g.a_sk, g.b_sb, g.b_sk, g.b_sn, g.o_sb, g.o_sm, g.o_sn,
                   (uint32_t)g.M, (uint32_t)g.N, (uint32_t)g.K, (uint32_t)g.batch, (uint32_t)g.in_dtype, (uint32_t)g.out_dtype,
                   g.bias, g.alpha, g.act, epi_on, g.rhs_dtype >= 0 ? (uint32_t)g.rhs_dtype + 1u : 0u};
  void* args[] = {&p};
  // gridDim.y is limited to 65535: taller problems (M > 1,048,560) walk their 16-row tiles with a stride of gridDim.y
  const uint64_t tiles_m = (g.M + 15) / 16;
  return launch(c, f, (unsigned)((g.N + 15) / 16), (unsigned)std::min<uint64_t>(tiles_m, 65535), (unsigned)g.batch, 256, 0, 1, st, args);
}

// Can TMA describe one operand in place?  `mn` x K elements, strides in elements.  K-major (rows of K) or MN-major (rows of
// the M / N extent): unit inner stride, 16-byte aligned base, row pitch and batch stride, pitch >= row, strides < 2^40 bytes.
static bool operand_tma_ok(uint64_t ptr, size_t esz, uint64_t mn, uint64_t K, uint64_t s_mn, uint64_t s_k, uint64_t s_b, bool* mn_major) {
  auto al16 = [&](uint64_t elems) { return (elems * esz) % 16 == 0; };
  const uint64_t lim = 1ull << 40;
  if (ptr % 16 || !al16(s_b)) return false;
  if (s_mn * esz >= lim || s_k * esz >= lim || s_b * esz >= lim) return false;
  if ((s_k == 1 || K == 1) && (mn == 1 || (al16(s_mn) && s_mn >= K))) { *mn_major = false; return true; }
  if ((s_mn == 1 || mn == 1) && (K == 1 || (al16(s_k) && s_k >= mn))) { *mn_major = true; return true; }
  return false;
}
static bool extents_tma_ok(const GemmProblem& g) {
  return g.M < (1ull << 31) && g.N < (1ull << 31) && g.K < (1ull << 31) && g.batch < (1ull << 31) && (g.o_sn == 1 || g.N == 1);
}
static bool tma_ok(const GemmProblem& g, bool* a_mn, bool* b_mn) {
  const size_t esz = dtype_size(g.in_dtype);
  return extents_tma_ok(g) && operand_tma_ok(g.a, esz, g.M, g.K, g.a_sm, g.a_sk, g.a_sb, a_mn) &&
         operand_tma_ok(g.b, esz, g.N, g.K, g.b_sn, g.b_sk, g.b_sb, b_mn);
}

static int reduce_workspace(b200_ctx* c, CUstream st, CUdeviceptr* out);

// Stream-K head plan for `tiles` tiles on `clusters` CTA pairs (gemm_tcgen05.cu, GemmParams): the rem = tiles % clusters
// tiles that would form a partial last wave are instead cut along K into `ranges` equal ranges processed FIRST, so all
// pairs stay busy and the slab exchange runs under the whole tiles that follow.
//   time (in tile-times) without: full_waves + 1.
//   with: full_waves + 1.4 * head + overhead / num_kb, head = ceil(ranges / clusters) * share / num_kb.
// Measured (profiles/r02_split_sweep.log): while the head runs, S pairs stream the operands of ONE tile, so the phase needs
// S times the operand bandwidth of a normal wave with less panel sharing in L2 -- it runs ~1.4x longer than its MMA time
// (bf16 4096^3: 92.0 us with S = 2 against 94.6 us for two waves of pair tiles; tf32 4096^3: 179 against 186 us).  Equal
// parts (ranges = rem * S) beat the even cut over all pairs whenever they fit (fp8 4096^3: 54.6 against 72.3 us), so S =
// floor(clusters / rem) when that is >= 2.  The exchange itself is hidden when whole tiles follow (~4 k-block times) and
// exposed, (14 + 8 parts) k-block times, when the head is the whole problem (round 1 measurement).
// gemm.split_k: auto (only when the model gains >= 4 %), off, on (whenever rem != 0), or N = 1..8 (N ranges per tile).
struct SkPlan {
  double time = 0;           // modelled time in tile-times
  uint64_t sk_tiles = 0, ranges = 0, umax = 0;
  bool bad_option = false;
};
static SkPlan sk_plan(uint64_t tiles, uint64_t clusters, uint64_t num_kb, const std::string& option, bool eligible) {
  SkPlan pl;
  const uint64_t full_waves = tiles / clusters, rem = tiles % clusters;
  pl.time = static_cast<double>(full_waves + (rem ? 1 : 0));
  if (option == "off" || !eligible || rem == 0 || num_kb < 2) return pl;
  const uint64_t total_kb = rem * num_kb;
  auto model = [&](uint64_t ranges, bool even_parts) {
    const uint64_t share = (total_kb + ranges - 1) / ranges;
    const double parts = std::max(1.0, static_cast<double>(ranges) / static_cast<double>(rem));
    const double head = static_cast<double>((ranges + clusters - 1) / clusters) * static_cast<double>(share) / static_cast<double>(num_kb);
    const double overhead = (full_waves >= 1 ? 4.0 : 14.0 + 8.0 * parts) / static_cast<double>(num_kb);
    return static_cast<double>(full_waves) + (even_parts ? 1.4 : 1.6) * head + overhead;
  };
  uint64_t ranges = 0;
  bool force = false, even = true;
  if (option == "auto" || option == "on") {
    force = (option == "on");
    const uint64_t s_fit = std::min<uint64_t>(clusters / rem, 8);
    if (s_fit >= 2) {
      // equal parts; when whole tiles follow, as many as fit; when the head is everything, the S the model likes best
      uint64_t best_s = s_fit;
      if (full_waves == 0)
        for (uint64_t s2 = 2; s2 <= s_fit; ++s2)
          if (num_kb / s2 >= 8 && model(rem * s2, true) < model(rem * best_s, true) - 1e-12) best_s = s2;
      ranges = rem * best_s;
    } else {
      ranges = clusters;       // more than half a wave of tiles: an even cut over all pairs, tiles in 1-2 uneven parts
      even = false;
    }
  } else {
    const int want = atoi(option.c_str());
    if (want < 1 || want > 8) { pl.bad_option = true; return pl; }
    if (want == 1) return pl;
    ranges = rem * static_cast<uint64_t>(want);
    force = true;
  }
  ranges = std::min(ranges, total_kb);                                    // every range owns at least one k-block
  if (ranges <= rem && !force) return pl;                                 // no tile would be cut
  const uint64_t share = (total_kb + ranges - 1) / ranges;
  const double t_sk = model(ranges, even);
  if (!force) {
    if (share < 8) return pl;                                             // slices too thin to amortise an exchange
    if (t_sk > 0.96 * pl.time) return pl;
  }
  pl.time = t_sk;
  pl.sk_tiles = rem;
  pl.ranges = ranges;
  pl.umax = (share + num_kb - 1) / num_kb + 1;                            // tiles one range can touch
  return pl;
}

// k-blocks (pipeline stages) one tile's K loop runs: 3xTF32 walks K three times; the hybrid f32 schedule once in tf32 stages
// (32 elements) and twice in bf16 stages (64 elements)
static uint64_t gemm_num_kb(const GemmProblem& g, uint32_t block_k) {
  const uint64_t seg = (g.K + block_k - 1) / block_k;
  if (!(g.a_lo != 0 && g.b_lo != 0)) return seg;
  return g.hybrid ? seg + 2 * ((g.K + 63) / 64) : 3 * seg;
}

// Tile variant by modelled time (ties -> larger tile, less L2 traffic): waves of tiles, with the last partial wave replaced by a
// stream-K head where that pays (sk_plan).  nullptr: gemm.variant names no variant this dtype / kind has.
static const GemmVariant* pick_variant(b200_ctx* c, const GemmProblem& g, SkPlan* sk_out) {
  const size_t esz = dtype_size(g.in_dtype);
  const uint32_t block_k = static_cast<uint32_t>(128 / esz);
  const std::string forced = opt(c, "gemm.variant", "auto");
  const std::string split_opt = opt(c, "gemm.split_k", "auto");
  const bool float_acc = !(g.in_dtype == B200_U8 || g.in_dtype == B200_I8);
  const uint64_t num_kb = gemm_num_kb(g, block_k);
  const GemmVariant* best = nullptr;
  double best_cost = 0;
  for (const GemmVariant& v : kVariants) {
    if (forced != "auto" && forced != v.tag) continue;
    if (forced == "auto" && v.eff <= 0.0) continue;
    if (!g.mx_kind && !variant_has_dtype(v, g.in_dtype)) continue;
    if (forced == "auto" && v.cg == 2 && g.M <= 128) continue;  // a CTA pair would idle its second half: one CTA per tile
    if (g.mx_kind && v.mt != 1) continue;                       // block-scaled kinds have no two-unit instantiation
    if (g.mx_kind == 3 && v.block_n == 224) continue;           // NVFP4: 448 accumulator columns leave room for one scale buffer only
    if (v.block_n == 224 && !(g.mx_kind && (g.sfb_any_layout || g.sfb_tile_rows == 224))) continue;
    if (v.block_n != 224 && g.mx_kind && !g.sfb_any_layout && g.sfb_tile_rows == 224) continue;   // scales already packed for 224-row tiles
    // the two-unit tile hides its epilogue only on the packed-register path (16-bit outputs); f32 outputs stay on 2sm_n256
    if (forced == "auto" && v.mt == 2 && !(g.out_dtype == B200_BF16 || g.out_dtype == B200_F16)) continue;
    const uint64_t tile_m = 128ull * v.cg * v.mt;
    const uint64_t tm = (g.M + tile_m - 1) / tile_m, tn = (g.N + v.block_n - 1) / v.block_n;
    const uint64_t tiles = tm * tn * g.batch;
    const uint64_t clusters = std::max(1, c->props.num_sms / v.cg);
    // the slab exchange is a per-128-row-CTA-tile protocol with f32 accumulators: not for the pair tile, not for integers
    const SkPlan sk = sk_plan(tiles, clusters, num_kb, split_opt, float_acc && v.mt == 1);
    // block-scaled kinds: measured 8192^3 ratios to the 256-wide tile are 0.63 (2sm_n128) and 0.61 (1sm_n128) for mxfp8,
    // 0.63 / 0.66 for mxfp4 -- the same ordering as the unscaled table, so it is reused
    const double eff = v.eff > 0 ? v.eff : 1.0;
    const double cost = sk.time * (128.0 * v.mt * v.block_n) / eff;  // per-SM MMA time
    if (!best || cost < best_cost * 0.999) { best = &v; best_cost = cost; *sk_out = sk; }
  }
  return best;
}

static int launch_tcgen05(b200_ctx* c, CUstream st, const GemmProblem& g, bool a_mn, bool b_mn) {
  const size_t esz = dtype_size(g.in_dtype), osz = dtype_size(g.out_dtype);
  const char* in_tag = g.mx_kind == 1 ? "mxf8" : g.mx_kind == 2 ? "mxf4" : g.mx_kind == 3 ? "nvf4"
                       : g.in_dtype == B200_BF16 ? "bf16" : g.in_dtype == B200_F16 ? "f16" : g.in_dtype == B200_F8E4M3 ? "e4m3"
                       : g.in_dtype == B200_F8E5M2 ? "e5m2" : g.in_dtype == B200_U8 ? "u8" : g.in_dtype == B200_I8 ? "s8" : "tf32";
  const char* out_tag = g.out_dtype == B200_BF16 ? "bf16" : g.out_dtype == B200_F16 ? "f16" : g.out_dtype == B200_I32 ? "i32" : "f32";
  const uint32_t block_k = static_cast<uint32_t>(128 / esz);
  const std::string forced = opt(c, "gemm.variant", "auto");
  const uint64_t k_segments = (g.a_lo != 0 && g.b_lo != 0) ? 3 : 1;
  SkPlan best_sk;
  const GemmVariant* best = pick_variant(c, g, &best_sk);
  if (!best) return fail(B200_ERR_INVALID_ARG, "gemm.variant '%s' is not a tcgen05 variant for this dtype", forced.c_str());
  if (best_sk.bad_option) return fail(B200_ERR_INVALID_ARG, "gemm.split_k must be auto, off, on or 1..8");
  const GemmVariant& v = *best;

  const std::string name = std::string("gemm_") + in_tag + "_" + out_tag + "_" + v.tag + (a_mn ? "_m" : "_k") + (b_mn ? "n" : "k");
  CUfunction f;
  int rc = get_func(c, name, &f);
  if (rc) return rc;
  const unsigned smem = gemm_smem_bytes(v, g.mx_kind);
  if (!c->dry) CU_CHECK(g_drv.cuFuncSetAttribute_p(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));

  const CUtensorMapDataType dt = g.in_dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : g.in_dtype == B200_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                 : esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                            : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const bool a_bcast = (g.a_sb == 0 || g.batch == 1), b_bcast = (g.b_sb == 0 || g.batch == 1);
  CUtensorMap ta, tb;
  auto pad16 = [&](uint64_t elems) { const uint64_t q = 16 / esz; return (elems + q - 1) / q * q; };
  const uint32_t chunk = static_cast<uint32_t>(128 / esz);  // MN-major operands: elements per 128-byte row
  const CUtensorMapSwizzle mn_swz = esz == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  if (!a_mn) {
    const uint64_t a_sm = g.M > 1 ? g.a_sm : pad16(g.K);
    rc = encode_tmap(c, &ta, dt, esz, g.a, g.K, g.M, a_bcast ? 1 : g.batch, a_sm, a_bcast ? a_sm * g.M : g.a_sb, block_k, 128);
  } else {
    const uint64_t a_sk = g.K > 1 ? g.a_sk : pad16(g.M);
    rc = encode_tmap(c, &ta, dt, esz, g.a, g.M, g.K, a_bcast ? 1 : g.batch, a_sk, a_bcast ? a_sk * g.K : g.a_sb, chunk, block_k, mn_swz);
  }
  if (rc) return rc;
  const uint32_t n_local = v.block_n / v.cg;
  if (!b_mn) {
    const uint64_t b_sn = g.N > 1 ? g.b_sn : pad16(g.K);
    rc = encode_tmap(c, &tb, dt, esz, g.b, g.K, g.N, b_bcast ? 1 : g.batch, b_sn, b_bcast ? b_sn * g.N : g.b_sb, block_k, n_local);
  } else {
    const uint64_t b_sk = g.K > 1 ? g.b_sk : pad16(g.N);
    // MN-major 32-bit operands: 32-byte swizzle atoms (matches the SWIZZLE_128B_BASE32B smem descriptor)
    rc = encode_tmap(c, &tb, dt, esz, g.b, g.N, g.K, b_bcast ? 1 : g.batch, b_sk, b_bcast ? b_sk * g.K : g.b_sb, chunk, block_k, mn_swz);
  }
  if (rc) return rc;
  // 3xTF32: compact low parts, same operand-major class as the originals (K-major: [rows, K]; MN-major: [K, cols])
  CUtensorMap ta_lo = ta, tb_lo = tb;
  const bool split = (g.a_lo != 0 && g.b_lo != 0);
  if (split && g.hybrid) {
    // bf16 pair buffers: planes [0, entries) = bf16(x), [entries, 2 entries) = bf16(lo); 64 elements of K per stage
    const uint64_t ab = a_bcast ? 1 : g.batch, bb = b_bcast ? 1 : g.batch;
    auto pad8 = [](uint64_t e) { return (e + 7) / 8 * 8; };
    const CUtensorMapDataType d16 = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    rc = !a_mn ? encode_tmap(c, &ta_lo, d16, 2, g.a_lo, g.K, g.M, 2 * ab, pad8(g.K), pad8(g.K) * g.M, 64, 128)
               : encode_tmap(c, &ta_lo, d16, 2, g.a_lo, g.M, g.K, 2 * ab, pad8(g.M), pad8(g.M) * g.K, 64, 64);
    if (rc) return rc;
    rc = !b_mn ? encode_tmap(c, &tb_lo, d16, 2, g.b_lo, g.K, g.N, 2 * bb, pad8(g.K), pad8(g.K) * g.N, 64, n_local)
               : encode_tmap(c, &tb_lo, d16, 2, g.b_lo, g.N, g.K, 2 * bb, pad8(g.N), pad8(g.N) * g.K, 64, 64);
    if (rc) return rc;
  } else if (split) {
    const uint64_t ab = a_bcast ? 1 : g.batch, bb = b_bcast ? 1 : g.batch;
    rc = !a_mn ? encode_tmap(c, &ta_lo, dt, esz, g.a_lo, g.K, g.M, ab, pad16(g.K), pad16(g.K) * g.M, block_k, 128)
               : encode_tmap(c, &ta_lo, dt, esz, g.a_lo, g.M, g.K, ab, pad16(g.M), pad16(g.M) * g.K, chunk, block_k, mn_swz);
    if (rc) return rc;
    rc = !b_mn ? encode_tmap(c, &tb_lo, dt, esz, g.b_lo, g.K, g.N, bb, pad16(g.K), pad16(g.K) * g.N, block_k, n_local)
               : encode_tmap(c, &tb_lo, dt, esz, g.b_lo, g.N, g.K, bb, pad16(g.N), pad16(g.N) * g.K, chunk, block_k, mn_swz);
    if (rc) return rc;
  }

  GemmParams p;
  memset(&p, 0, sizeof(p));
  if (g.mx_kind) {
    // packed scale tensors viewed as (16 B, 32 rows x atoms, 128-row tiles); one box = the chunks of one k-block
    // 128-row chunks per batch entry; the rhs scales packed per 224-row tile take two chunks per tile (see pack_scales)
    const uint64_t tiles_a = (g.M + 127) / 128;
    const uint64_t tiles_b = g.sfb_tile_rows == 224 ? 2 * ((g.N + 223) / 224) : (g.N + 127) / 128;
    const uint64_t ab = a_bcast ? 1 : g.batch, bb = b_bcast ? 1 : g.batch;
    rc = encode_sf_tmap(c, &ta_lo, g.sfa, g.sf_atoms, tiles_a * ab, mx_atoms(g.mx_kind), 1);
    if (rc) return rc;
    rc = encode_sf_tmap(c, &tb_lo, g.sfb, g.sf_atoms, tiles_b * bb, mx_atoms(g.mx_kind), (v.block_n + 127) / 128);
    if (rc) return rc;
    {  // who issues the scale copies: one copy thread (default), two copy threads in different warps, or the MMA thread (A/B reference)
      const std::string sfc = opt(c, "gemm.sf_copy", "thread");
      if (sfc != "thread" && sfc != "thread2" && sfc != "mma") return fail(B200_ERR_INVALID_ARG, "gemm.sf_copy must be thread, thread2 or mma");
      p.sf_flags = sfc == "mma" ? 1u : sfc == "thread2" ? 2u : 0u;
    }
    p.sf_fmt_a = g.fmt_a; p.sf_fmt_b = g.fmt_b;
    p.sf_tiles_a = (uint32_t)tiles_a; p.sf_tiles_b = (uint32_t)tiles_b;
  }
  if (!g.mx_kind && g.rhs_dtype >= 0 && g.rhs_dtype != g.in_dtype) {
    auto fmt8 = [](int dt) { return (dt == B200_F8E5M2 || dt == B200_I8) ? 1u : 0u; };   // kind::f8f6f4: e4m3 0 / e5m2 1; kind::i8: u8 0 / s8 1
    p.sf_fmt_a = fmt8(g.in_dtype); p.sf_fmt_b = fmt8(g.rhs_dtype); p.fmt_mixed = 1;
  }
  p.k_segments = (uint32_t)k_segments;
  if (split && g.hybrid) {
    p.hyb = 1;
    p.hyb_nba = a_bcast ? 1u : (uint32_t)g.batch;
    p.hyb_nbb = b_bcast ? 1u : (uint32_t)g.batch;
  }
  p.alpha = g.alpha; p.bias = g.bias; p.epi_act = g.act;
  p.epi_on = (g.alpha != 1.0f || g.bias != 0 || g.act != 0) ? 1u : 0u;
  p.out = g.out;
  p.out_row_stride = g.o_sm;
  p.out_batch_stride = g.o_sb;
  p.M = (uint32_t)g.M; p.N = (uint32_t)g.N; p.K = (uint32_t)g.K; p.batch = (uint32_t)g.batch;
  p.tiles_m = (uint32_t)((g.M + 128 * v.cg * v.mt - 1) / (128 * v.cg * v.mt));
  p.tiles_n = (uint32_t)((g.N + v.block_n - 1) / v.block_n);
  p.group_m = (uint32_t)std::max(1, atoi(opt(c, "gemm.group_m", "8").c_str()));
  p.a_bmul = a_bcast ? 0 : 1;
  p.b_bmul = b_bcast ? 0 : 1;
  p.vec_store = (g.out % 16 == 0 && (g.o_sm * osz) % 16 == 0 && (g.o_sb * osz) % 16 == 0) ? 1 : 0;
  // whole tiles leave through swizzled staging tiles and TMA stores when `out` is describable: (N, M, batch), 128-byte boxes
  CUtensorMap tout;
  memset(&tout, 0, sizeof(tout));
  const uint64_t lim40 = 1ull << 40;
  if (p.vec_store && opt(c, "gemm.epilogue", "tma") == "tma" && g.o_sm * osz < lim40 && g.o_sb * osz < lim40 && (g.M == 1 || g.o_sm >= g.N)) {
    const uint64_t o_sm = g.M > 1 ? g.o_sm : (g.N + 15) / 16 * 16;
    const uint64_t o_sb = g.batch > 1 ? g.o_sb : o_sm * g.M;
    rc = encode_tmap(c, &tout, osz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, osz, g.out, g.N, g.M, g.batch,
                     o_sm, o_sb, static_cast<uint32_t>(128 / osz), 32);
    if (rc) return rc;
    p.tma_store = 1;
  }

  const uint64_t total_tiles = static_cast<uint64_t>(p.tiles_m) * p.tiles_n * p.batch;
  if (total_tiles >= (1ull << 32)) return fail(B200_ERR_UNSUPPORTED, "too many tiles");
  const unsigned max_clusters = (unsigned)std::max(1, c->props.num_sms / v.cg);
  unsigned clusters = (unsigned)std::min<uint64_t>(total_tiles, max_clusters);

  // Stream-K head (plan chosen with the variant above)
  CUdeviceptr slabs = 0;
  p.full_tiles = (uint32_t)total_tiles;
  if (best_sk.sk_tiles) {
    if (best_sk.sk_tiles * v.cg > kWsGemmTickets) return fail(B200_ERR_UNSUPPORTED, "stream-K head: %llu tiles exceed the ticket area", (unsigned long long)best_sk.sk_tiles);
    CUdeviceptr ws = 0;
    rc = reduce_workspace(c, st, &ws);
    if (rc) return rc;
    const uint64_t slab_bytes = 128ull * v.cg * v.block_n * 4;
    rc = pool_alloc(c, best_sk.ranges * best_sk.umax * slab_bytes, &slabs, st);
    if (rc) return rc;
    p.full_tiles = (uint32_t)(total_tiles - best_sk.sk_tiles);
    p.sk_tiles = (uint32_t)best_sk.sk_tiles;
    p.sk_ranges = (uint32_t)best_sk.ranges;
    p.sk_umax = (uint32_t)best_sk.umax;
    p.split_ws = slabs;
    p.split_tickets = ws + kWsGemmTicketOffset;
    // the head's ranges are dealt to pairs 0 .. ranges-1 (mod the grid); whole tiles to every pair
    clusters = (unsigned)std::min<uint64_t>(std::max<uint64_t>(p.full_tiles, best_sk.ranges), max_clusters);
    if (c->dry) {
      char line[200];
      snprintf(line, sizeof(line), "gemm stream-k head: %u whole tiles + %u tiles in %u k-ranges (<= %u slabs per range)\n", p.full_tiles, p.sk_tiles,
               p.sk_ranges, p.sk_umax);
      c->plan += line;
    }
  }
  void* args[] = {&ta, &tb, &ta_lo, &tb_lo, &tout, &p};
  // block-scaled kernels carry one more warp (the optional second scale-copy thread)
  rc = launch(c, f, clusters * v.cg, 1, 1, 256 + 128 * (v.mt - 1) + (g.mx_kind ? 32 : 0), smem, v.cg, st, args);
  if (slabs) pool_free(c, slabs, st);  // stream-ordered: reusable by later work once this launch has drained
  return rc;
}

static inline uint64_t pad4(uint64_t elems) { return (elems + 3) / 4 * 4; }

// lo = x - trunc_tf32(x) of a logical [batch, rows, cols] view (cols innermost), written with row pitch pad4(cols)
static int launch_split(b200_ctx* c, CUstream st, uint64_t in, uint64_t out, uint64_t batch, uint64_t rows, uint64_t cols,
                        uint64_t in_bs, uint64_t in_rs) {
  CUfunction f;
  int rc = get_func(c, "split_tf32_lo", &f);
  if (rc) return rc;
  SplitParams p{in, out, batch, rows, cols, in_bs, in_rs, pad4(cols)};
  const uint64_t total = batch * rows * cols;
  const unsigned grid = (unsigned)std::min<uint64_t>((total / 4 + 255) / 256 + 1, (uint64_t)c->props.num_sms * 16);
  void* args[] = {&p};
  return launch(c, f, std::max(1u, grid), 1, 1, 256, 0, 1, st, args);
}

// bf16 pair of a logical [batch, rows, cols] f32 view: plane 0 = bf16(x), plane 1 = bf16(x - trunc_tf32(x)), rows pitched to 8 elements
static inline uint64_t pad8e(uint64_t elems) { return (elems + 7) / 8 * 8; }
static int launch_split_pair(b200_ctx* c, CUstream st, uint64_t in, uint64_t out, uint64_t batch, uint64_t rows, uint64_t cols,
                             uint64_t in_bs, uint64_t in_rs) {
  CUfunction f;
  int rc = get_func(c, "split_f32_bf16_pair", &f);
  if (rc) return rc;
  SplitParams p{in, out, batch, rows, cols, in_bs, in_rs, pad8e(cols)};
  const uint64_t total = batch * rows * cols;
  const unsigned grid = (unsigned)std::min<uint64_t>((total / 8 + 255) / 256 + 1, (uint64_t)c->props.num_sms * 16);
  void* args[] = {&p};
  return launch(c, f, std::max(1u, grid), 1, 1, 256, 0, 1, st, args);
}

struct RepitchParams {
  uint64_t in, out;
  uint64_t batch, rows, cols;        // logical [batch, rows, cols] of the copy, cols innermost in the OUTPUT
  uint64_t in_sb, in_sr, in_sc;      // input strides in elements
  uint64_t out_pitch;                // output row pitch in elements (16-byte multiple)
  uint32_t esz, pad;
};

// One pass that copies an operand TMA cannot describe (row pitch or base not 16-byte aligned, no unit stride) into a pooled
// buffer it can: [batch, rows, pitch] with the operand's own contiguous dimension innermost when it has one.
static int stage_operand(b200_ctx* c, CUstream st, size_t esz, uint64_t ptr, uint64_t batch, uint64_t mn, uint64_t K, uint64_t s_mn, uint64_t s_k,
                         uint64_t s_b, CUdeviceptr* out, uint64_t* o_smn, uint64_t* o_sk, uint64_t* o_sb) {
  const bool keep_mn_major = (s_mn == 1 && s_k != 1);          // rows of the M / N extent: keep them (coalesced both ways)
  const uint64_t rows = keep_mn_major ? K : mn, cols = keep_mn_major ? mn : K;
  const uint64_t q = 16 / esz, pitch = (cols + q - 1) / q * q;
  const uint64_t nb = (s_b == 0) ? 1 : batch;                  // a broadcast operand is staged once
  CUdeviceptr buf;
  int rc = pool_alloc(c, nb * rows * pitch * esz, &buf, st);
  if (rc) return rc;
  CUfunction f;
  rc = get_func(c, "repitch_rows", &f);
  if (rc) { pool_free(c, buf, st); return rc; }
  RepitchParams p{ptr, buf, nb, rows, cols, s_b, keep_mn_major ? s_k : s_mn, keep_mn_major ? s_mn : s_k, pitch, (uint32_t)esz, 0};
  const uint64_t vecs = nb * rows * (pitch / q);               // one 16-byte output vector per thread
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((vecs + 255) / 256, 0x7FFFFFFFull));
  void* args[] = {&p};
  rc = launch(c, f, grid, 1, 1, 256, 0, 1, st, args);
  if (rc) { pool_free(c, buf, st); return rc; }
  *out = buf;
  *o_smn = keep_mn_major ? 1 : pitch;
  *o_sk = keep_mn_major ? pitch : 1;
  *o_sb = (s_b == 0) ? 0 : rows * pitch;
  return B200_OK;
}

static int run_gemm_staged(b200_ctx* c, CUstream st, const GemmProblem& g);

static int run_gemm(b200_ctx* c, CUstream st, const GemmProblem& g) {
  if (g.M == 0 || g.N == 0 || g.batch == 0) return B200_OK;
  const std::string forced = opt(c, "gemm.variant", "auto");
  bool a_mn = false, b_mn = false;
  const bool tma = g.K > 0 && tma_ok(g, &a_mn, &b_mn);
  if (forced == "simt" || !tma) {
    if (forced != "simt" && forced != "auto")
      return fail(B200_ERR_UNSUPPORTED, "gemm.variant=%s forced but operands are not TMA-describable", forced.c_str());
    // Operands whose pitch / base TMA cannot describe (bf16 with K = 4097, an odd sub-view): one staging pass into an
    // aligned pooled copy, then the tensor-core kernel -- the strided SIMT kernel is kept for tiny problems and for
    // outputs without a unit inner stride.  gemm.stage=off keeps the SIMT path (reference-order arithmetic) for everything.
    const bool big = g.M * g.N * g.K * g.batch >= (1ull << 21);
    if (forced == "auto" && g.K > 0 && big && extents_tma_ok(g) && opt(c, "gemm.stage", "on") == "on") return run_gemm_staged(c, st, g);
    return launch_simt(c, st, g);
  }
  const std::string f32_mode = g.in_dtype == B200_F32 ? opt(c, "gemm.f32", "hybrid") : std::string();
  if (g.in_dtype == B200_F32 && f32_mode != "hybrid" && f32_mode != "3xtf32" && f32_mode != "tf32")
    return fail(B200_ERR_INVALID_ARG, "gemm.f32 must be hybrid, 3xtf32 or tf32");
  if (f32_mode == "hybrid") {
    // f32-grade product in TWO tensor passes' worth of time, one launch: x = hi + lo with hi = the top 19 bits (what the tf32
    // datapath reads from the original tensor).  hi*hi runs as kind::tf32 on the originals; the cross terms A*B_lo + A_lo*B run
    // as kind::f16 on bf16 copies (bf16(A), bf16(B_lo), bf16(A_lo), bf16(B)) at twice the tf32 rate, into the same f32
    // accumulators.  The cross terms are ~2^-11 of the product, so their bf16 rounding (2^-9) lands at ~2^-20 of it.
    const uint64_t ab = (g.a_sb == 0) ? 1 : g.batch, bb = (g.b_sb == 0) ? 1 : g.batch;
    const uint64_t a_elems = !a_mn ? g.M * pad8e(g.K) : g.K * pad8e(g.M), b_elems = !b_mn ? g.N * pad8e(g.K) : g.K * pad8e(g.N);
    CUdeviceptr a_p = 0, b_p = 0;
    int rc = pool_alloc(c, 2 * ab * a_elems * 2, &a_p, st);
    if (rc) return rc;
    rc = pool_alloc(c, 2 * bb * b_elems * 2, &b_p, st);
    if (rc) { pool_free(c, a_p, st); return rc; }
    rc = !a_mn ? launch_split_pair(c, st, g.a, a_p, ab, g.M, g.K, g.a_sb, g.a_sm) : launch_split_pair(c, st, g.a, a_p, ab, g.K, g.M, g.a_sb, g.a_sk);
    if (!rc) rc = !b_mn ? launch_split_pair(c, st, g.b, b_p, bb, g.N, g.K, g.b_sb, g.b_sn) : launch_split_pair(c, st, g.b, b_p, bb, g.K, g.N, g.b_sb, g.b_sk);
    if (!rc) {
      GemmProblem h = g;
      h.a_lo = a_p;
      h.b_lo = b_p;
      h.hybrid = true;
      rc = launch_tcgen05(c, st, h, a_mn, b_mn);
    }
    pool_free(c, a_p, st);
    pool_free(c, b_p, st);
    return rc;
  }
  if (f32_mode == "3xtf32") {
    // 3xTF32 in ONE GEMM launch: the tf32 datapath reads only the top 19 bits of an f32 operand, so the original tensors
    // are the "hi" parts; only lo = x - hi is materialised (compact), and the kernel runs K three times:
    // (A,B) + (A,B_lo) + (A_lo,B), f32 accumulation throughout.
    const uint64_t ab = (g.a_sb == 0) ? 1 : g.batch, bb = (g.b_sb == 0) ? 1 : g.batch;
    const uint64_t a_elems = !a_mn ? g.M * pad4(g.K) : g.K * pad4(g.M), b_elems = !b_mn ? g.N * pad4(g.K) : g.K * pad4(g.N);
    CUdeviceptr a_lo = 0, b_lo = 0;
    int rc = pool_alloc(c, ab * a_elems * 4, &a_lo, st);
    if (rc) return rc;
    rc = pool_alloc(c, bb * b_elems * 4, &b_lo, st);
    if (rc) { pool_free(c, a_lo, st); return rc; }
    rc = !a_mn ? launch_split(c, st, g.a, a_lo, ab, g.M, g.K, g.a_sb, g.a_sm) : launch_split(c, st, g.a, a_lo, ab, g.K, g.M, g.a_sb, g.a_sk);
    if (!rc) rc = !b_mn ? launch_split(c, st, g.b, b_lo, bb, g.N, g.K, g.b_sb, g.b_sn) : launch_split(c, st, g.b, b_lo, bb, g.K, g.N, g.b_sb, g.b_sk);
    if (!rc) {
      GemmProblem h = g;
      h.a_lo = a_lo;
      h.b_lo = b_lo;
      rc = launch_tcgen05(c, st, h, a_mn, b_mn);
    }
    // stream-ordered reuse: the pool hands these pages out again only to later work on this context
    pool_free(c, a_lo, st);
    pool_free(c, b_lo, st);
    return rc;
  }
  return launch_tcgen05(c, st, g, a_mn, b_mn);
}

static int run_gemm_staged(b200_ctx* c, CUstream st, const GemmProblem& g) {
  const size_t esz = dtype_size(g.in_dtype);
  GemmProblem h = g;
  CUdeviceptr sa = 0, sb = 0;
  bool mn;
  int rc = B200_OK;
  if (!operand_tma_ok(g.a, esz, g.M, g.K, g.a_sm, g.a_sk, g.a_sb, &mn)) {
    rc = stage_operand(c, st, esz, g.a, g.batch, g.M, g.K, g.a_sm, g.a_sk, g.a_sb, &sa, &h.a_sm, &h.a_sk, &h.a_sb);
    if (!rc) h.a = sa;
  }
  if (!rc && !operand_tma_ok(g.b, esz, g.N, g.K, g.b_sn, g.b_sk, g.b_sb, &mn)) {
    rc = stage_operand(c, st, esz, g.b, g.batch, g.N, g.K, g.b_sn, g.b_sk, g.b_sb, &sb, &h.b_sn, &h.b_sk, &h.b_sb);
    if (!rc) h.b = sb;
  }
  if (!rc) {
    bool a_mn = false, b_mn = false;
    rc = tma_ok(h, &a_mn, &b_mn) ? run_gemm(c, st, h) : launch_simt(c, st, g);
  }
  if (sa) pool_free(c, sa, st);
  if (sb) pool_free(c, sb, st);
  return rc;
}

// Collapse batch dims [0, nb) of one operand into a linear stride; false if the offsets are not linear in the flat index.
static bool linear_batch(int nb, const uint64_t* out_shape, const uint64_t* shape, const uint64_t* strides, uint64_t* flat) {
  uint64_t inner_stride = 0, inner_extent = 1;
  bool have = false;
  for (int i = nb - 1; i >= 0; --i) {
    if (out_shape[i] == 1) continue;
    const uint64_t st = (shape[i] == 1) ? 0 : strides[i];
    if (!have) { inner_stride = st; inner_extent = out_shape[i]; have = true; *flat = st; continue; }
    if (st != inner_stride * inner_extent) return false;
    inner_extent *= out_shape[i];
  }
  if (!have) *flat = 0;
  return true;
}

static int matmul_rec(b200_ctx* c, CUstream st, GemmProblem g, int nb, const uint64_t* ob, const uint64_t* ls,
                      const uint64_t* lst, const uint64_t* rs, const uint64_t* rst, const uint64_t* ost) {
  uint64_t fa = 0, fb = 0, fo = 0;
  if (linear_batch(nb, ob, ls, lst, &fa) && linear_batch(nb, ob, rs, rst, &fb) && linear_batch(nb, ob, ob, ost, &fo)) {
    uint64_t batch = 1;
    for (int i = 0; i < nb; ++i) batch *= ob[i];
    g.batch = batch; g.a_sb = fa; g.b_sb = fb; g.o_sb = fo;
    if (batch == 1) { g.a_sb = g.b_sb = 0; g.o_sb = 0; }
    return run_gemm(c, st, g);
  }
  // peel the outermost non-unit batch dim and recurse
  int d = 0;
  while (d < nb && ob[d] == 1) ++d;
  const size_t esz = dtype_size(g.in_dtype), osz = dtype_size(g.out_dtype);
  for (uint64_t i = 0; i < ob[d]; ++i) {
    GemmProblem h = g;
    h.a += (ls[d] == 1 ? 0 : i * lst[d]) * esz;
    h.b += (rs[d] == 1 ? 0 : i * rst[d]) * esz;
    h.out += i * ost[d] * osz;
    int rc = matmul_rec(c, st, h, nb - d - 1, ob + d + 1, ls + d + 1, lst + d + 1, rs + d + 1, rst + d + 1, ost + d + 1);
    if (rc) return rc;
  }
  return B200_OK;
}

static int matmul_impl(b200_ctx* c, b200_stream s, b200_dtype in_dtype, b200_dtype out_dtype, b200_dptr lhs,
                       b200_dptr rhs, b200_dptr out, int rank, const uint64_t* shape_lhs, const uint64_t* strides_lhs,
                       const uint64_t* shape_rhs, const uint64_t* strides_rhs, const uint64_t* shape_out,
                       const uint64_t* strides_out, const b200_epilogue* ep, int rhs_dtype = -1);

extern "C" int b200_matmul(b200_ctx* c, b200_stream s, b200_dtype in_dtype, b200_dtype out_dtype, b200_dptr lhs,
                           b200_dptr rhs, b200_dptr out, int rank, const uint64_t* shape_lhs, const uint64_t* strides_lhs,
                           const uint64_t* shape_rhs, const uint64_t* strides_rhs, const uint64_t* shape_out,
                           const uint64_t* strides_out) {
  return matmul_impl(c, s, in_dtype, out_dtype, lhs, rhs, out, rank, shape_lhs, strides_lhs, shape_rhs, strides_rhs, shape_out,
                     strides_out, nullptr);
}

extern "C" int b200_matmul_fused(b200_ctx* c, b200_stream s, b200_dtype in_dtype, b200_dtype out_dtype, b200_dptr lhs,
                                 b200_dptr rhs, b200_dptr out, int rank, const uint64_t* shape_lhs, const uint64_t* strides_lhs,
                                 const uint64_t* shape_rhs, const uint64_t* strides_rhs, const uint64_t* shape_out,
                                 const uint64_t* strides_out, const b200_epilogue* ep) {
  if (!ep) return fail(B200_ERR_INVALID_ARG, "matmul_fused: null epilogue");
  if (ep->activation < 0 || ep->activation > 2) return fail(B200_ERR_INVALID_ARG, "matmul_fused: unknown activation %d", ep->activation);
  if (in_dtype == B200_U8 || in_dtype == B200_I8) return fail(B200_ERR_UNSUPPORTED, "matmul_fused: integer accumulators have no float epilogue");
  return matmul_impl(c, s, in_dtype, out_dtype, lhs, rhs, out, rank, shape_lhs, strides_lhs, shape_rhs, strides_rhs, shape_out,
                     strides_out, ep);
}

// ------------------------------------------------------------------------------------------------ block-scaled matmul
static int launch_pack_scales(b200_ctx* c, CUstream st, uint64_t in, uint64_t out, uint64_t batch, uint64_t rows,
                              uint64_t n_scales, uint64_t tiles, uint64_t atoms, uint32_t pad_value, uint32_t tile_rows = 128) {
  CUfunction f;
  int rc = get_func(c, "pack_scales", &f);
  if (rc) return rc;
  PackScalesParams p{in, out, (uint32_t)batch, (uint32_t)rows, (uint32_t)n_scales, (uint32_t)tiles, (uint32_t)atoms, pad_value,
                     tile_rows, tile_rows == 224 ? 2u : 1u};
  const uint64_t words = batch * tiles * atoms * 128;
  const unsigned grid = (unsigned)std::min<uint64_t>((words + 255) / 256, (uint64_t)c->props.num_sms * 8);
  void* args[] = {&p};
  return launch(c, f, std::max(1u, grid), 1, 1, 256, 0, 1, st, args);
}

extern "C" int b200_matmul_scaled(b200_ctx* c, b200_stream s, b200_dtype lhs_dtype, b200_dtype rhs_dtype, b200_dtype out_dtype,
                                  b200_dptr lhs, b200_dptr rhs, b200_dptr lhs_scales, b200_dptr rhs_scales, b200_dptr out,
                                  uint64_t batch, uint64_t M, uint64_t N, uint64_t K, int scale_block, int scales_packed) {
  CTX_ENTER(c);
  const bool fp4 = (lhs_dtype == B200_F4E2M1X2);
  auto is_fp8 = [](int d) { return d == B200_F8E4M3 || d == B200_F8E5M2; };
  if (!(fp4 ? rhs_dtype == B200_F4E2M1X2 : (is_fp8(lhs_dtype) && is_fp8(rhs_dtype))))
    return fail(B200_ERR_UNSUPPORTED, "matmul_scaled: operands must be fp8 (e4m3/e5m2, mixable) or both packed e2m1");
  if (out_dtype != B200_F32 && out_dtype != B200_BF16 && out_dtype != B200_F16)
    return fail(B200_ERR_UNSUPPORTED, "matmul_scaled: output must be f32, bf16 or f16");
  // scale_block 32: ue8m0 scales (MXFP8 / MXFP4); scale_block 16: ue4m3 scales, packed e2m1 operands only (NVFP4)
  const bool nvf4 = (scale_block == 16);
  if (scale_block != 32 && !(nvf4 && fp4))
    return fail(B200_ERR_UNSUPPORTED, "matmul_scaled: scale block %d (32 = ue8m0 scales; 16 = ue4m3 scales, packed e2m1 operands only)", scale_block);
  if (K == 0 || K % 32) return fail(B200_ERR_INVALID_ARG, "matmul_scaled: K = %llu must be a positive multiple of 32", (unsigned long long)K);
  if (batch == 0 || M == 0 || N == 0) return B200_OK;
  if (!lhs || !rhs || !lhs_scales || !rhs_scales || !out) return fail(B200_ERR_INVALID_ARG, "matmul_scaled: null device pointer");
  if (M >= (1ull << 31) || N >= (1ull << 31) || K >= (1ull << 31) || batch >= (1ull << 20)) return fail(B200_ERR_UNSUPPORTED, "matmul_scaled: extent too large");
  CUstream st = resolve_stream(c, s);
  const uint64_t n_scales = K / scale_block, atoms = (n_scales + 3) / 4;
  const uint64_t k_bytes = fp4 ? K / 2 : K;
  const std::string forced = opt(c, "gemm.variant", "auto");
  const bool tma = (lhs % 16 == 0 && rhs % 16 == 0 && k_bytes % 16 == 0 && (!scales_packed || (lhs_scales % 16 == 0 && rhs_scales % 16 == 0)));
  if (forced == "simt" || !tma) {
    if (scales_packed) return fail(B200_ERR_UNSUPPORTED, "matmul_scaled: packed scales need 16-byte aligned operands and K rows");
    if (forced != "simt" && forced != "auto")
      return fail(B200_ERR_UNSUPPORTED, "gemm.variant=%s forced but operands are not TMA-describable", forced.c_str());
    CUfunction f;
    int rc = get_func(c, "gemm_scaled_simt", &f);
    if (rc) return rc;
    ScaledSimtParams p{lhs, rhs, lhs_scales, rhs_scales, out, (uint32_t)batch, (uint32_t)M, (uint32_t)N, (uint32_t)K,
                       (uint32_t)lhs_dtype, (uint32_t)rhs_dtype, (uint32_t)out_dtype, (uint32_t)scale_block, 1u, 1u, nvf4 ? 1u : 0u, 0u};
    const uint64_t total = batch * M * N;
    const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, (uint64_t)c->props.num_sms * 16);
    void* args[] = {&p};
    return launch(c, f, std::max(1u, grid), 1, 1, 256, 0, 1, st, args);
  }
  // the problem as the GEMM sees it (operands described to TMA as bytes); the tile variant is chosen BEFORE the scales are
  // packed, because the 256 x 224 variant wants the rhs scales packed per 224-row tile
  GemmProblem g{};
  g.in_dtype = B200_F8E4M3;  // 1-byte marker
  g.out_dtype = out_dtype;
  g.a = lhs; g.b = rhs; g.out = out;
  g.M = M; g.N = N; g.K = k_bytes; g.batch = batch;
  g.a_sm = k_bytes; g.a_sk = 1; g.a_sb = batch > 1 ? M * k_bytes : 0;
  g.b_sn = k_bytes; g.b_sk = 1; g.b_sb = batch > 1 ? N * k_bytes : 0;
  g.o_sm = N; g.o_sn = 1; g.o_sb = batch > 1 ? M * N : 0;
  g.mx_kind = nvf4 ? 3 : fp4 ? 2 : 1;
  g.fmt_a = fp4 ? 1u : (lhs_dtype == B200_F8E5M2 ? 1u : 0u);
  g.fmt_b = fp4 ? 1u : (rhs_dtype == B200_F8E5M2 ? 1u : 0u);
  g.sf_atoms = atoms;
  g.sfb_any_layout = !scales_packed;           // pre-packed scales are in the plain 128-row layout
  SkPlan sk_unused;
  const GemmVariant* v = pick_variant(c, g, &sk_unused);
  if (!v) return fail(B200_ERR_INVALID_ARG, "gemm.variant '%s' is not a tcgen05 variant for block-scaled operands%s", forced.c_str(),
                      scales_packed ? " with pre-packed scales" : "");
  g.sfb_any_layout = false;
  g.sfb_tile_rows = (v->block_n == 224) ? 224 : 128;
  // scales -> the tensor core's packed chunks (skipped when the caller already holds them in that form)
  const uint64_t tiles_a = (M + 127) / 128;
  const uint64_t tiles_b = g.sfb_tile_rows == 224 ? 2 * ((N + 223) / 224) : (N + 127) / 128;
  if (batch * tiles_a >= (1ull << 31) || batch * tiles_b >= (1ull << 31) || atoms * 32 >= (1ull << 31))
    return fail(B200_ERR_UNSUPPORTED, "matmul_scaled: scale tensor too large for 32-bit TMA coordinates");
  CUdeviceptr sfa = lhs_scales, sfb = rhs_scales;
  int rc = B200_OK;
  if (!scales_packed) {
    sfa = sfb = 0;
    rc = pool_alloc(c, batch * tiles_a * atoms * 512, &sfa, st);
    if (rc) return rc;
    rc = pool_alloc(c, batch * tiles_b * atoms * 512, &sfb, st);
    if (rc) { pool_free(c, sfa, st); return rc; }
    const uint32_t one = nvf4 ? 0x38u : 127u;
    rc = launch_pack_scales(c, st, lhs_scales, sfa, batch, M, n_scales, tiles_a, atoms, one);
    if (!rc) rc = launch_pack_scales(c, st, rhs_scales, sfb, batch, N, n_scales, tiles_b, atoms, one, (uint32_t)g.sfb_tile_rows);
  }
  if (!rc) {
    g.sfa = sfa; g.sfb = sfb;
    rc = launch_tcgen05(c, st, g, false, false);
  }
  if (!scales_packed) { pool_free(c, sfa, st); pool_free(c, sfb, st); }
  return rc;
}

// Mixed 8-bit operand formats (the cartesian products the reference instantiates for its manual MMA,
// crates/cubecl-cpp/src/cuda/mma/manual.rs:151-166 i8 x u8 / u8 x i8 and :170-186 fp8 pairs): same kernels, the two format
// fields of the tcgen05 instruction descriptor differ.
extern "C" int b200_matmul_mixed(b200_ctx* c, b200_stream s, b200_dtype lhs_dtype, b200_dtype rhs_dtype, b200_dtype out_dtype, b200_dptr lhs,
                                 b200_dptr rhs, b200_dptr out, int rank, const uint64_t* shape_lhs, const uint64_t* strides_lhs,
                                 const uint64_t* shape_rhs, const uint64_t* strides_rhs, const uint64_t* shape_out,
                                 const uint64_t* strides_out) {
  auto fp8 = [](int d) { return d == B200_F8E4M3 || d == B200_F8E5M2; };
  auto int8 = [](int d) { return d == B200_U8 || d == B200_I8; };
  if (lhs_dtype != rhs_dtype && !((fp8(lhs_dtype) && fp8(rhs_dtype)) || (int8(lhs_dtype) && int8(rhs_dtype))))
    return fail(B200_ERR_UNSUPPORTED, "matmul_mixed: operand formats %d x %d cannot be mixed (fp8 e4m3/e5m2 pairs, u8/i8 pairs)", (int)lhs_dtype, (int)rhs_dtype);
  return matmul_impl(c, s, lhs_dtype, out_dtype, lhs, rhs, out, rank, shape_lhs, strides_lhs, shape_rhs, strides_rhs, shape_out,
                     strides_out, nullptr, lhs_dtype == rhs_dtype ? -1 : (int)rhs_dtype);
}

static int matmul_impl(b200_ctx* c, b200_stream s, b200_dtype in_dtype, b200_dtype out_dtype, b200_dptr lhs,
                       b200_dptr rhs, b200_dptr out, int rank, const uint64_t* shape_lhs, const uint64_t* strides_lhs,
                       const uint64_t* shape_rhs, const uint64_t* strides_rhs, const uint64_t* shape_out,
                       const uint64_t* strides_out, const b200_epilogue* ep, int rhs_dtype) {
  CTX_ENTER(c);
  if (rank < 2 || rank > 8) return fail(B200_ERR_INVALID_ARG, "matmul: rank %d unsupported (need 2..8)", rank);
  if (!shape_lhs || !strides_lhs || !shape_rhs || !strides_rhs || !shape_out || !strides_out)
    return fail(B200_ERR_INVALID_ARG, "matmul: null shape/stride array");
  const bool fp8 = (in_dtype == B200_F8E4M3 || in_dtype == B200_F8E5M2);
  const bool int8 = (in_dtype == B200_U8 || in_dtype == B200_I8);
  if (in_dtype != B200_F32 && in_dtype != B200_F16 && in_dtype != B200_BF16 && !fp8 && !int8)
    return fail(B200_ERR_UNSUPPORTED, "matmul: input dtype %d unsupported (f32, f16, bf16, f8e4m3, f8e5m2, u8, i8)", (int)in_dtype);
  if (int8 ? (out_dtype != B200_I32)
      : fp8 ? (out_dtype != B200_F32 && out_dtype != B200_BF16 && out_dtype != B200_F16)
            : (out_dtype != in_dtype && out_dtype != B200_F32))
    return fail(B200_ERR_UNSUPPORTED, "matmul: output dtype must equal the input dtype or be f32 (fp8 inputs: bf16/f16/f32; u8/i8 inputs: i32)");
  const int nb = rank - 2;
  const uint64_t M = shape_lhs[rank - 2], K = shape_lhs[rank - 1], K2 = shape_rhs[rank - 2], N = shape_rhs[rank - 1];
  // shape.rs:489-517: inner dims must agree, batch dims broadcast 1 vs d
  if (K != K2) return fail(B200_ERR_INVALID_ARG, "matmul: inner dimensions differ (lhs k=%llu, rhs k=%llu)", (unsigned long long)K, (unsigned long long)K2);
  if (shape_out[rank - 2] != M || shape_out[rank - 1] != N)
    return fail(B200_ERR_INVALID_ARG, "matmul: output is [%llu,%llu], expected [%llu,%llu]", (unsigned long long)shape_out[rank - 2],
                (unsigned long long)shape_out[rank - 1], (unsigned long long)M, (unsigned long long)N);
  for (int i = 0; i < nb; ++i) {
    const uint64_t l = shape_lhs[i], r = shape_rhs[i], o = shape_out[i];
    const uint64_t expect = l == r ? l : (l == 1 ? r : (r == 1 ? l : 0));
    if (expect == 0 && !(l == 0 && r == 0)) return fail(B200_ERR_INVALID_ARG, "matmul: batch dim %d cannot broadcast (%llu vs %llu)", i, (unsigned long long)l, (unsigned long long)r);
    if (o != expect) return fail(B200_ERR_INVALID_ARG, "matmul: output batch dim %d is %llu, expected %llu", i, (unsigned long long)o, (unsigned long long)expect);
  }
  if (!lhs || !rhs || !out) {
    uint64_t n = M * N;
    for (int i = 0; i < nb; ++i) n *= shape_out[i];
    if (n == 0) return B200_OK;
    return fail(B200_ERR_INVALID_ARG, "matmul: null device pointer");
  }
  CUstream st = resolve_stream(c, s);
  GemmProblem g{};
  g.in_dtype = in_dtype; g.out_dtype = out_dtype; g.rhs_dtype = rhs_dtype;
  g.a = lhs; g.b = rhs; g.out = out;
  g.M = M; g.N = N; g.K = K; g.batch = 1;
  g.a_sm = strides_lhs[rank - 2]; g.a_sk = strides_lhs[rank - 1];
  g.b_sk = strides_rhs[rank - 2]; g.b_sn = strides_rhs[rank - 1];
  g.o_sm = strides_out[rank - 2]; g.o_sn = strides_out[rank - 1];
  if (ep) { g.alpha = ep->alpha; g.bias = ep->bias; g.act = (uint32_t)ep->activation; }
  for (int i = 0; i < nb; ++i)
    if (shape_out[i] == 0) return B200_OK;
  return matmul_rec(c, st, g, nb, shape_out, shape_lhs, strides_lhs, shape_rhs, strides_rhs, strides_out);
}

// ================================================================================================ reduce
static int reduce_workspace(b200_ctx* c, CUstream st, CUdeviceptr* out) {
  if (c->dry) { *out = 0x6000000000ull; return B200_OK; }
  auto it = c->reduce_ws.find(st);
  if (it != c->reduce_ws.end()) { *out = it->second; return B200_OK; }
  CUdeviceptr p;
  CUresult r = g_drv.cuMemAlloc_p(&p, kWsBytes);
  if (r != CUDA_SUCCESS) return fail(map_cu(r), "reduce workspace allocation failed: %s", cu_err(r));
  r = g_drv.cuMemsetD32Async_p(p, 0, kWsBytes / 4, st);  // ticket starts at 0; kernels reset it themselves
  if (r != CUDA_SUCCESS) { g_drv.cuMemFree_p(p); return fail(map_cu(r), "reduce workspace memset failed: %s", cu_err(r)); }
  c->reduce_ws[st] = p;
  *out = p;
  return B200_OK;
}

static const char* op_tag(int op) {
  switch (op) {
    case B200_REDUCE_SUM: case B200_REDUCE_MEAN: return "sum";
    case B200_REDUCE_PROD: return "prod";
    case B200_REDUCE_MAX: return "max";
    case B200_REDUCE_MIN: return "min";
    case B200_REDUCE_ARGMAX: return "argmax";
    case B200_REDUCE_ARGMIN: return "argmin";
    default: return nullptr;
  }
}
static const char* dt_tag(int dt) { return dt == B200_F32 ? "f32" : dt == B200_F16 ? "f16" : dt == B200_BF16 ? "bf16" : nullptr; }
static bool fill_dtype_ok(int dt) { return dt_tag(dt) || dt == B200_F8E4M3 || dt == B200_F8E5M2; }

extern "C" int b200_into_contiguous(b200_ctx* c, b200_stream s, b200_dtype dtype, b200_dptr in, b200_dptr out, int rank,
                                    const uint64_t* shape, const uint64_t* strides);

// A reducible VIEW of the input: logical [outer, len, inner] with explicit element strides and an optional row pitch.
struct RView {
  uint64_t in = 0;
  uint64_t outer = 1, len = 1, inner = 1;
  uint64_t s_outer = 0, s_len = 1;
  uint64_t row_len = 1, row_pitch = 1;
};

static unsigned opt_uint(b200_ctx* c, const char* key, unsigned dflt, unsigned lo, unsigned hi) {
  const std::string v = opt(c, key, "");
  if (v.empty()) return dflt;
  const long x = atol(v.c_str());
  return (unsigned)std::min<long>(hi, std::max<long>(lo, x));
}
static uint64_t pow2_ceil(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }
static uint64_t pow2_floor(uint64_t x) { uint64_t p = 1; while (p * 2 <= x) p <<= 1; return p; }
static uint64_t ceil_div(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// Reduce every element of the view `v` (v.len elements in logical rows of v.row_len, v.row_pitch apart) to out[0].
static int launch_reduce_all(b200_ctx* c, CUstream st, int op, int dt, const RView& v, uint64_t out, float scale) {
  const bool arg = (op == B200_REDUCE_ARGMAX || op == B200_REDUCE_ARGMIN);
  const bool pitched = v.row_len != v.len;
  const uint64_t n = v.len;
  const size_t esz = dtype_size(dt);
  std::string name = std::string(pitched ? "reduce_allp_" : "reduce_all_") + op_tag(op) + "_" + dt_tag(dt);
  unsigned threads = opt_uint(c, "reduce.threads", 512, 32, 512) / 32 * 32;
  unsigned bps = opt_uint(c, "reduce.blocks_per_sm", 4, 1, 64);
  unsigned smem = 0, stages = 0;
  bool bulk = false;
  if (!pitched && !arg) {
    // variants: the plain 128-bit streaming kernel, its tuning forms (f32 sum only) and the bulk-copy staged kernel
    // auto = the bulk-copy staged kernel once the input is big enough to fill a ring on every SM (measured, 1 GiB f32 sum:
    // 7.08 TB/s against 6.92 for the best plain-load form, profiles/r02_reduce_sweep.log); plain loads below that
    const std::string var = opt(c, "reduce.variant", "auto");
    if (var == "tma" || var == "auto") {
      bulk = n * esz >= (var == "tma" ? 64ull * 16384 : (uint64_t)c->props.num_sms * 8 * 16384);
    } else if (var != "u8") {
      if (std::string(op_tag(op)) == "sum" && dt == B200_F32) name += "_" + var;
    }
  }
  const uint64_t vec = 16 / esz;
  unsigned grid;
  if (bulk) {
    name += "_tma";
    threads = 256 + 32;                       // eight consumer warps + one producer warp
    stages = opt_uint(c, "reduce.tma_stages", 6, 2, 8);   // measured: 6 x 16 KB 145.4 us, 8 x 16 KB 147.3 us, 4 x 16 KB 150.0 us (1 GiB f32)
    smem = stages * 16384 + 128;
    const uint64_t tiles = n * esz / 16384;
    const unsigned per_sm = stages <= 6 ? opt_uint(c, "reduce.tma_ctas_per_sm", 1, 1, 2) : 1;
    grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)c->props.num_sms * per_sm));
  } else {
    const uint64_t want = (n / vec + threads - 1) / threads;  // blocks that still get >= 1 vector per thread
    grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, std::min<uint64_t>((uint64_t)c->props.num_sms * bps, kWsMaxBlocks)));
  }
  CUfunction f;
  int rc = get_func(c, name, &f);
  if (rc) return rc;
  if (smem && !c->dry) CU_CHECK(g_drv.cuFuncSetAttribute_p(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
  CUdeviceptr ws;
  rc = reduce_workspace(c, st, &ws);
  if (rc) return rc;
  ReduceParams p{};
  p.in = v.in; p.out = out; p.ws = ws;
  p.outer = 1; p.len = n; p.inner = 1;
  p.row_len = v.row_len; p.row_pitch = v.row_pitch;
  p.seg_len = n; p.nseg = 1; p.scale = scale;
  p.ctu = stages;                             // bulk-copy kernels: ring depth
  void* args[] = {&p};
  // Overlap with the preceding all-element reduction on the context's stream: safe because that kernel writes only its
  // 4-byte result and the workspace, this one reads neither before its own griddepcontrol.wait -- unless its input is the
  // predecessor's output.
  const uint64_t in_end = v.in + (pitched ? (n / v.row_len) * v.row_pitch : n) * esz;
  const bool pdl = st == c->stream && c->pdl_prev_out != 0 && opt(c, "reduce.pdl", "on") == "on" &&
                   !(c->pdl_prev_out + 4 > v.in && c->pdl_prev_out < in_end);
  rc = launch(c, f, grid, 1, 1, threads, smem, 1, st, args, pdl);
  if (!rc && st == c->stream) c->pdl_prev_out = out;
  return rc;
}

// One launch of the rows kernel: items = outer x nseg, item (o, s) covers elements [s * seg_len, ..) of row o.
static int launch_rows_kernel(b200_ctx* c, CUstream st, int op, int dt, const RView& v, uint64_t seg_len, uint64_t out, uint64_t out2, float scale,
                              bool pdl = false) {
  const std::string name = std::string("reduce_rows_") + op_tag(op) + "_" + dt_tag(dt);
  CUfunction f;
  int rc = get_func(c, name, &f);
  if (rc) return rc;
  const size_t esz = dtype_size(dt);
  const uint64_t vec = 16 / esz;
  const uint64_t nseg = ceil_div(v.len, seg_len), items = v.outer * nseg;
  // Threads per item (power of two): about `vpt` 128-bit vectors per thread, so a 32 KB row is one 256-thread block and the
  // grid has many more blocks than resident slots (the hardware scheduler balances the tail block by block -- a warp per
  // 32 KB row left the last, nearly empty wave running at a third of the bandwidth: ncu, round 1).
  // measured (profiles/r02_reduce_sweep.log): 16 vectors per thread for inputs that stream from HBM for a while ([9000,16384]
  // 6.38 vs 6.09 TB/s, argmax [8192,8192] 5.65 vs 4.90), 8 for small launch-bound inputs ([512,8192] 7.1 vs 8.6 us)
  const bool big = v.outer * v.len * esz >= (128ull << 20);
  const unsigned vpt = opt_uint(c, "reduce.rows_vpt", big ? 16 : 8, 1, 64);
  const uint64_t nv = ceil_div(std::min(seg_len, v.len), vec);
  uint64_t tpr = std::min<uint64_t>(512, pow2_ceil(ceil_div(nv, vpt)));
  int tpr_log2 = 0;
  while ((1ull << tpr_log2) < tpr) ++tpr_log2;
  const unsigned threads = tpr > 256 ? (unsigned)tpr : 256;
  const uint64_t items_per_block = threads >> tpr_log2;
  uint64_t blocks = ceil_div(items, items_per_block);
  const bool uniform = (v.in % 16) == 0 && ((v.s_outer * esz) % 16) == 0 && (v.len % vec) == 0;
  if (tpr <= 32 && uniform && nseg == 1 && v.len / vec <= tpr) blocks = ceil_div(blocks, 4);  // short rows: four rows in flight per thread group
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(blocks, 0x7FFFFFFFull));
  ReduceParams p{};
  p.in = v.in; p.out = out; p.out2 = out2;
  p.outer = v.outer; p.len = v.len; p.inner = 1;
  p.s_outer = v.s_outer; p.s_len = 1;
  p.row_len = 1; p.row_pitch = 1;
  p.seg_len = seg_len; p.nseg = (uint32_t)nseg; p.scale = scale;
  void* args[] = {&p, &tpr_log2};
  return launch(c, f, grid, 1, 1, threads, 0, 1, st, args, pdl);
}

// One launch of the column kernel over the view (items = outer x nseg x column tiles).
// `final_out` != 0 with a segmented axis: finish in the same launch when the (outer, column tile) tickets fit the workspace
// (*fused = true), else the caller runs the second pass.
static int launch_cols_kernel(b200_ctx* c, CUstream st, int op, int dt, const RView& v, uint64_t seg_len, uint64_t out, uint64_t out2, float scale,
                              uint64_t final_out = 0, bool* fused = nullptr, bool pdl = false) {
  const bool arg_op = (op == B200_REDUCE_ARGMAX || op == B200_REDUCE_ARGMIN);
  const std::string name = std::string("reduce_cols_") + op_tag(op) + "_" + dt_tag(dt) + (opt(c, "reduce.cols_loads", arg_op ? "4" : "8") == "8" ? "_n8" : "");
  CUfunction f;
  int rc = get_func(c, name, &f);
  if (rc) return rc;
  const size_t esz = dtype_size(dt);
  const uint64_t vec = 16 / esz;
  const bool vector = v.inner % vec == 0 && v.row_len % vec == 0 && v.in % 16 == 0 && (v.s_len * esz) % 16 == 0 &&
                      (v.s_outer * esz) % 16 == 0 && (v.row_pitch * esz) % 16 == 0;
  const uint64_t units = vector ? v.inner / vec : v.inner;
  const uint64_t nseg = ceil_div(v.len, seg_len);
  const uint64_t seg = std::min(seg_len, v.len);
  // tile shape: RL row lanes x ctu column units per 256-thread block.  A warp's worth of units (512 contiguous bytes per row
  // in vector mode) keeps the loads coalesced; the rest of the block goes to row lanes as long as every lane still has ~4 rows
  const uint64_t ctu_min = std::min<uint64_t>(units, 32);
  const uint64_t rl_max = 256 / ctu_min;
  const uint64_t rl = std::min<uint64_t>(rl_max, std::max<uint64_t>(1, pow2_floor(std::max<uint64_t>(1, seg / 4))));
  const uint64_t ctu = std::min<uint64_t>(units, 256 / rl);
  const uint64_t tiles = ceil_div(units, ctu);
  const uint64_t items = v.outer * nseg * tiles;
  const unsigned bps = 8;
  // short axis: a block's item is small, so blocks walk several items (persistent grid); long axis: one item per block
  const uint64_t cap = seg * ctu * (vector ? 16 : esz) >= (64u << 10) ? 0x7FFFFFFFull : (uint64_t)c->props.num_sms * bps;
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(items, cap));
  ReduceParams p{};
  p.in = v.in; p.out = out; p.out2 = out2;
  p.outer = v.outer; p.len = v.len; p.inner = v.inner;
  p.s_outer = v.s_outer; p.s_len = v.s_len;
  p.row_len = v.row_len; p.row_pitch = v.row_pitch;
  p.seg_len = seg_len; p.nseg = (uint32_t)nseg; p.ctu = (uint32_t)ctu; p.scale = scale;
  p.flags = vector ? 2u : 0u;
  if (fused) *fused = false;
  if (final_out && nseg > 1 && v.outer * tiles <= kWsColTickets && opt(c, "reduce.cols_fused", "off") == "on") {
    CUdeviceptr ws;
    rc = reduce_workspace(c, st, &ws);
    if (rc) return rc;
    p.ws = ws; p.final_out = final_out; p.flags |= 4u;
    if (fused) *fused = true;
  }
  if (nseg > 1 && !(p.flags & 4u)) p.scale = 1.0f;   // first pass of a two-launch reduction: the scale (mean) belongs to the second
  void* args[] = {&p};
  return launch(c, f, grid, 1, 1, 256, 0, 1, st, args, pdl);
}

static int launch_argcombine(b200_ctx* c, CUstream st, uint64_t keys, uint64_t idx, uint64_t out, uint64_t outer, uint64_t nseg, uint64_t inner) {
  CUfunction f;
  int rc = get_func(c, "reduce_argcombine", &f);
  if (rc) return rc;
  ArgCombineParams p{keys, idx, out, outer, nseg, inner};
  const uint64_t threads = nseg >= 8 ? outer * inner * 32 : outer * inner;   // a warp per output when there are many segments
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(ceil_div(threads, 256), (uint64_t)c->props.num_sms * 8));
  void* args[] = {&p};
  return launch(c, f, grid, 1, 1, 256, 0, 1, st, args);
}

// Reduce the `len` axis of the view.  Few outputs with a long axis are reduced in two passes (segments of the axis first,
// then the per-segment partials), both deterministic; everything else is one launch.
static int reduce_axis_view(b200_ctx* c, CUstream st, int op, int dt, const RView& v, uint64_t out, float scale) {
  const bool arg = (op == B200_REDUCE_ARGMAX || op == B200_REDUCE_ARGMIN);
  const uint64_t sms = c->props.num_sms;
  const size_t esz = dtype_size(dt);
  const uint64_t vec = 16 / esz;
  // Segment the axis when whole rows / columns cannot fill the machine: aim at ~16 blocks per SM (several waves of small
  // blocks, so the block scheduler evens out the tail) while every segment keeps a useful amount of work.
  uint64_t seg_len = v.len;
  if (v.inner == 1) {
    const unsigned bps = opt_uint(c, "reduce.rows_blocks_per_sm", 4, 1, 64);
    if (v.outer < sms * bps && v.len >= 16384 && v.outer * v.len * esz >= (32ull << 20)) {   // below ~32 MB one launch wins (launch-bound)
      const uint64_t nseg = std::min<uint64_t>(ceil_div(sms * 16, v.outer), v.len / 4096);
      if (nseg > 1) seg_len = ceil_div(ceil_div(v.len, nseg), 512) * 512;   // 512 elements: segments stay vector-aligned
    }
  } else {
    const uint64_t units = (v.inner % vec == 0) ? v.inner / vec : v.inner;
    const uint64_t tiles = ceil_div(units, std::min<uint64_t>(units, 32));
    const unsigned bps = opt_uint(c, "reduce.cols_blocks_per_sm", 4, 1, 64);
    if (v.outer * tiles < sms * bps && v.len >= 256) {
      // ~8 blocks per SM (two resident waves of big blocks) measured best on average for value ops (eight loads in flight per
      // thread) and arg ops (four) alike -- profiles/r02_cols_sweep.log.  Target 0: exactly one resident wave (4 blocks per SM).
      const unsigned tgt = opt_uint(c, "reduce.cols_split_target", 8, 0, 256);
      uint64_t nseg;
      if (tgt == 0) nseg = std::max<uint64_t>(1, (sms * 4) / (v.outer * tiles));
      else nseg = ceil_div(sms * tgt, v.outer * tiles);
      nseg = std::min<uint64_t>(nseg, v.len / 64);
      if (nseg > 1) seg_len = ceil_div(v.len, nseg);
    }
  }
  const uint64_t nseg = ceil_div(v.len, seg_len);
  if (arg && nseg > 1 && !c->dry && v.len >= (1ull << 32)) return fail(B200_ERR_UNSUPPORTED, "arg-reduce: axis extent does not fit u32 indices");
  if (nseg == 1) {
    return v.inner == 1 ? launch_rows_kernel(c, st, op, dt, v, v.len, out, 0, scale) : launch_cols_kernel(c, st, op, dt, v, v.len, out, 0, scale);
  }
  // two passes: partials [outer, nseg, inner] (f32 values, or u32 keys + u32 indices), then the partials
  const uint64_t count = v.outer * nseg * v.inner;
  CUdeviceptr tmp = 0, tmp2 = 0;
  int rc = pool_alloc(c, count * 4, &tmp, st);
  if (rc) return rc;
  if (arg) {
    rc = pool_alloc(c, count * 4, &tmp2, st);
    if (rc) { pool_free(c, tmp, st); return rc; }
  }
  bool fused = false;
  rc = v.inner == 1 ? launch_rows_kernel(c, st, op, dt, v, seg_len, tmp, tmp2, 1.0f)
                    : launch_cols_kernel(c, st, op, dt, v, seg_len, tmp, tmp2, scale, out, &fused);
  if (!rc && !fused) {
    if (arg) {
      rc = launch_argcombine(c, st, tmp, tmp2, out, v.outer, nseg, v.inner);
    } else {
      RView t;
      t.in = tmp; t.outer = v.outer; t.len = nseg; t.inner = v.inner;
      t.s_outer = nseg * v.inner; t.s_len = v.inner; t.row_len = v.inner; t.row_pitch = v.inner;
      // the second pass only depends on the first: launched with programmatic serialization, its blocks are resident (and
      // past their prologue) when the first pass drains; they wait in griddepcontrol.wait
      const bool pdl = opt(c, "reduce.pdl", "on") == "on";
      rc = v.inner == 1 ? launch_rows_kernel(c, st, op, B200_F32, t, nseg, out, 0, scale, pdl)
                        : launch_cols_kernel(c, st, op, B200_F32, t, nseg, out, 0, scale, 0, nullptr, pdl);
    }
  }
  pool_free(c, tmp, st);
  if (tmp2) pool_free(c, tmp2, st);
  return rc;
}

// Can the strided tensor be reduced IN PLACE?  Yes when, in memory order, it is a dense tensor whose innermost rows may be
// pitched (what PitchedMemoryLayoutPolicy produces, crates/cubecl-runtime/src/allocator.rs:21-72) -- in any axis
// permutation that keeps the kept axes in their logical order (a transposed view reduces the other physical axis).
// Anything else (broadcast strides, gaps elsewhere, permuted outputs) goes through into_contiguous.
static bool plan_view(int rank, const uint64_t* shape, const uint64_t* strides, int axis, bool arg, uint64_t in, RView* v) {
  struct Dim { uint64_t ext, st; int pos; bool ax; };
  std::vector<Dim> d;
  bool unit_axis = false;
  uint64_t expect = 1;
  std::vector<uint64_t> cst(rank);
  for (int i = rank - 1; i >= 0; --i) { cst[i] = expect; expect *= shape[i]; }
  for (int i = 0; i < rank; ++i) {
    if (shape[i] == 1) { if (i == axis) unit_axis = true; continue; }
    d.push_back(Dim{shape[i], strides ? strides[i] : cst[i], i, i == axis});
  }
  std::stable_sort(d.begin(), d.end(), [](const Dim& a, const Dim& b) { return a.st > b.st; });
  const int k = (int)d.size();
  bool pitched = false;
  if (k > 0) {
    if (d[k - 1].st != 1) return false;
    for (int j = k - 2; j >= 0; --j) {
      const uint64_t e = d[j + 1].st * d[j + 1].ext;
      if (d[j].st == e) continue;
      if (j == k - 2 && d[j].st > e) { pitched = true; continue; }
      return false;
    }
  }
  // kept axes must appear in memory order exactly as in logical order (otherwise the output would need a permuted store)
  int last = -1;
  for (int j = 0; j < k; ++j) {
    if (d[j].ax && !(axis < 0)) continue;
    if (axis < 0 && !arg) continue;     // a value reduction over everything is order-independent
    if (d[j].pos < last) return false;
    last = d[j].pos;
  }
  v->in = in;
  uint64_t total = 1;
  for (int j = 0; j < k; ++j) total *= d[j].ext;
  const uint64_t row = k > 0 ? d[k - 1].ext : 1, pitch = pitched ? d[k - 2].st : row;
  if (axis < 0) {
    v->outer = 1; v->len = total; v->inner = 1;
    v->row_len = pitched ? row : total; v->row_pitch = pitched ? pitch : total;
    return true;
  }
  if (unit_axis) {  // reducing an axis of extent 1: a (converting) copy of everything else
    v->outer = 1; v->len = 1; v->s_len = 0; v->inner = total;
    v->row_len = pitched ? row : total; v->row_pitch = pitched ? pitch : total;
    return true;
  }
  int pa = -1;
  for (int j = 0; j < k; ++j) if (d[j].ax) pa = j;
  if (pa < 0) return false;
  v->outer = 1; v->inner = 1;
  for (int j = 0; j < pa; ++j) v->outer *= d[j].ext;
  for (int j = pa + 1; j < k; ++j) v->inner *= d[j].ext;
  v->len = d[pa].ext;
  v->s_len = d[pa].st;
  v->s_outer = pa > 0 ? d[pa - 1].st : 0;
  if (pitched && pa <= k - 3) { v->row_len = row; v->row_pitch = pitch; }
  else { v->row_len = v->inner; v->row_pitch = v->inner; }
  return true;
}

static int reduce_impl(b200_ctx* c, b200_stream s, b200_reduce_op op, b200_dtype in_dtype, b200_dptr in, b200_dptr out,
                       int rank, const uint64_t* shape, const uint64_t* strides, int axis) {
  if (!op_tag(op)) return fail(B200_ERR_INVALID_ARG, "reduce: unknown op %d", (int)op);
  if (!dt_tag(in_dtype)) return fail(B200_ERR_UNSUPPORTED, "reduce: input dtype %d unsupported (f32, f16, bf16)", (int)in_dtype);
  if (rank < 1 || rank > 8 || !shape) return fail(B200_ERR_INVALID_ARG, "reduce: bad rank/shape");
  if (axis < -1 || axis >= rank) return fail(B200_ERR_INVALID_ARG, "reduce: axis %d out of range for rank %d", axis, rank);
  uint64_t n = 1, len = 1;
  for (int i = 0; i < rank; ++i) n *= shape[i];
  len = axis < 0 ? n : shape[axis];
  uint64_t outputs = 1;
  for (int i = 0; i < rank; ++i) if (axis >= 0 && i != axis) outputs *= shape[i];
  if (outputs == 0) return B200_OK;  // empty output
  if (len == 0) return fail(B200_ERR_INVALID_ARG, "reduce: reduced extent is 0 (identity-filled outputs are not defined by the reference)");
  if (!in || !out) return fail(B200_ERR_INVALID_ARG, "reduce: null device pointer");
  const bool arg = (op == B200_REDUCE_ARGMAX || op == B200_REDUCE_ARGMIN);
  if (arg && len >= (1ull << 32)) return fail(B200_ERR_UNSUPPORTED, "arg-reduce: axis extent %llu does not fit u32 indices", (unsigned long long)len);
  if (in % dtype_size(in_dtype)) return fail(B200_ERR_INVALID_ARG, "reduce: input pointer is not aligned to its element size");
  const float scale = (op == B200_REDUCE_MEAN) ? static_cast<float>(1.0 / static_cast<double>(len)) : 1.0f;
  CUstream st = resolve_stream(c, s);
  RView v;
  if (plan_view(rank, shape, strides, axis, arg, in, &v)) {
    if (axis < 0) return launch_reduce_all(c, st, op, in_dtype, v, out, scale);
    return reduce_axis_view(c, st, op, in_dtype, v, out, scale);
  }
  // not reducible in place: gather into a compact temporary first (into_contiguous), then reduce that
  CUdeviceptr tmp;
  int rc = pool_alloc(c, n * dtype_size(in_dtype), &tmp, st);
  if (rc) return rc;
  rc = b200_into_contiguous(c, s, in_dtype, in, tmp, rank, shape, strides);
  if (!rc) {
    RView w;
    const bool ok = plan_view(rank, shape, nullptr, axis, arg, tmp, &w);
    rc = !ok ? fail(B200_ERR_UNKNOWN, "reduce: contiguous plan failed")
             : axis < 0 ? launch_reduce_all(c, st, op, in_dtype, w, out, scale) : reduce_axis_view(c, st, op, in_dtype, w, out, scale);
  }
  pool_free(c, tmp, st);
  return rc;
}

extern "C" int b200_reduce(b200_ctx* c, b200_stream s, b200_reduce_op op, b200_dtype in_dtype, b200_dptr in, b200_dptr out,
                           int rank, const uint64_t* shape, int axis) {
  CTX_ENTER(c);
  return reduce_impl(c, s, op, in_dtype, in, out, rank, shape, nullptr, axis);
}

extern "C" int b200_reduce_strided(b200_ctx* c, b200_stream s, b200_reduce_op op, b200_dtype in_dtype, b200_dptr in, b200_dptr out,
                                   int rank, const uint64_t* shape, const uint64_t* strides, int axis) {
  CTX_ENTER(c);
  return reduce_impl(c, s, op, in_dtype, in, out, rank, shape, strides, axis);
}

// Stage timings (ns) of the most recent fused reduce + exchange launched with option reduce.debug=1 on stream `s`:
// words[0] = exchange (publish -> all peers seen), words[1] = partials + f64 tree of the last block.  Synchronises the stream.
extern "C" int b200_reduce_debug(b200_ctx* c, b200_stream s, uint64_t* words4) {
  CTX_ENTER_DEVICE(c);
  if (!words4) return fail(B200_ERR_INVALID_ARG, "reduce_debug: null output");
  CUstream st = resolve_stream(c, s);
  CUdeviceptr ws;
  int rc = reduce_workspace(c, st, &ws);
  if (rc) return rc;
  CU_CHECK(g_drv.cuMemcpyDtoHAsync_p(words4, ws + kWsDebugOffset, 32, st));
  CU_CHECK(g_drv.cuStreamSynchronize_p(st));
  return B200_OK;
}

struct GatherParams {
  uint64_t in, out, n;
  uint64_t shape[8], strides[8];
  uint32_t rank, esz;
};

extern "C" int b200_into_contiguous(b200_ctx* c, b200_stream s, b200_dtype dtype, b200_dptr in, b200_dptr out, int rank,
                                    const uint64_t* shape, const uint64_t* strides) {
  CTX_ENTER(c);
  const size_t esz = dtype_size(dtype);
  if (!esz) return fail(B200_ERR_INVALID_ARG, "into_contiguous: unknown dtype %d", (int)dtype);
  if (rank < 1 || rank > 8 || !shape || !strides) return fail(B200_ERR_INVALID_ARG, "into_contiguous: bad rank/shape/strides");
  GatherParams p;
  memset(&p, 0, sizeof(p));
  p.in = in; p.out = out; p.rank = (uint32_t)rank; p.esz = (uint32_t)esz; p.n = 1;
  for (int i = 0; i < rank; ++i) { p.shape[i] = shape[i]; p.strides[i] = strides[i]; p.n *= shape[i]; }
  if (p.n == 0) return B200_OK;
  if (!in || !out) return fail(B200_ERR_INVALID_ARG, "into_contiguous: null device pointer");
  CUfunction f;
  int rc = get_func(c, "gather_strided", &f);
  if (rc) return rc;
  const unsigned grid = (unsigned)std::min<uint64_t>((p.n + 255) / 256, (uint64_t)c->props.num_sms * 32);
  void* args[] = {&p};
  return launch(c, f, std::max(1u, grid), 1, 1, 256, 0, 1, resolve_stream(c, s), args);
}

// ================================================================================================ collectives
extern "C" int b200_comm_get_unique_id(b200_ctx* c, void* id128) {
  CTX_ENTER_DEVICE(c);
  if (!id128) return fail(B200_ERR_INVALID_ARG, "null id");
  int rc = ensure_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  NCCL_CHECK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return B200_OK;
}

static std::vector<int> sorted_ids(const int* ids, int n) {
  std::vector<int> v(ids, ids + n);
  std::sort(v.begin(), v.end());
  return v;
}

extern "C" int b200_comm_init(b200_ctx* c, const int* device_ids, int n, const void* id128) {
  CTX_ENTER_DEVICE(c);
  if (!device_ids || n < 1 || !id128) return fail(B200_ERR_INVALID_ARG, "comm_init: bad arguments");
  int rc = ensure_nccl();
  if (rc) return rc;
  std::vector<int> key = sorted_ids(device_ids, n);
  if (c->comms.count(key)) return B200_OK;  // idempotent, like ensure_init_collective (client.rs:755-767)
  auto it = std::find(key.begin(), key.end(), c->device);
  if (it == key.end()) return fail(B200_ERR_INVALID_ARG, "comm_init: device %d is not in the device set", c->device);
  CommState cs;
  cs.rank = static_cast<int>(it - key.begin());
  cs.n = n;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NCCL_CHECK(g_nccl.CommInitRank(&cs.comm, n, id, cs.rank));
  c->comms[key] = cs;
  return B200_OK;
}

extern "C" int b200_all_reduce(b200_ctx* c, b200_stream compute, b200_dptr src, b200_dptr dst, size_t bytes, b200_dtype dtype,
                               b200_comm_op op, const int* device_ids, int n) {
  CTX_ENTER_DEVICE(c);
  if (!device_ids || n < 1) return fail(B200_ERR_INVALID_ARG, "all_reduce: bad device set");
  int rc = ensure_nccl();
  if (rc) return rc;
  auto it = c->comms.find(sorted_ids(device_ids, n));
  if (it == c->comms.end()) return fail(B200_ERR_COMM, "all_reduce: no communicator for this device set (call b200_comm_init)");
  int nt;
  switch (dtype) {  // communication.rs:34-108
    case B200_F32: nt = ncclFloat32; break;
    case B200_F16: nt = ncclFloat16; break;
    case B200_BF16: nt = ncclBfloat16; break;
    case B200_F64: nt = ncclFloat64; break;
    case B200_I32: nt = ncclInt32; break;
    case B200_U32: nt = ncclUint32; break;
    case B200_I64: nt = ncclInt64; break;
    case B200_U64: nt = ncclUint64; break;
    case B200_I8: nt = ncclInt8; break;
    case B200_U8: nt = ncclUint8; break;
    default: return fail(B200_ERR_UNSUPPORTED, "all_reduce: dtype %d", (int)dtype);
  }
  const size_t esz = dtype_size(dtype);
  if (bytes % esz) return fail(B200_ERR_INVALID_ARG, "all_reduce: %zu bytes is not a multiple of the element size", bytes);
  CUstream cs = resolve_stream(c, compute);
  // compute -> comm dependency, then the collective on the comm stream (server.rs:749)
  CU_CHECK(g_drv.cuEventRecord_p(c->comm_event, cs));
  CU_CHECK(g_drv.cuStreamWaitEvent_p(c->comm_stream, c->comm_event, 0));
  NCCL_CHECK(g_nccl.AllReduce(reinterpret_cast<const void*>(src), reinterpret_cast<void*>(dst), bytes / esz, nt,
                              op == B200_COMM_MEAN ? ncclAvg : ncclSum, it->second.comm, c->comm_stream));
  return B200_OK;
}

extern "C" int b200_sync_collective(b200_ctx* c, b200_stream compute) {
  CTX_ENTER_DEVICE(c);
  CU_CHECK(g_drv.cuEventRecord_p(c->comm_event, c->comm_stream));
  CU_CHECK(g_drv.cuStreamWaitEvent_p(resolve_stream(c, compute), c->comm_event, 0));
  return B200_OK;
}

// ------------------------------------------------------------------------------------------------ peer-memory exchange

static int ensure_mailbox(b200_ctx* c) {
  if (c->mailbox) return B200_OK;
  CU_CHECK(g_drv.cuMemAlloc_p(&c->mailbox, kMailboxBytes));
  CU_CHECK(g_drv.cuMemsetD32Async_p(c->mailbox, 0, kMailboxBytes / 4, c->stream));
  CU_CHECK(g_drv.cuStreamSynchronize_p(c->stream));
  return B200_OK;
}

extern "C" int b200_p2p_export(b200_ctx* c, void* ipc_handle64, uint64_t* local_ptr, int64_t* pid) {
  CTX_ENTER_DEVICE(c);
  if (!ipc_handle64 || !local_ptr || !pid) return fail(B200_ERR_INVALID_ARG, "p2p_export: null argument");
  int rc = ensure_mailbox(c);
  if (rc) return rc;
  CUipcMemHandle h;
  static_assert(sizeof(CUipcMemHandle) == B200_IPC_HANDLE_BYTES, "IPC handle size");
  CU_CHECK(g_drv.cuIpcGetMemHandle_p(&h, c->mailbox));
  memcpy(ipc_handle64, &h, sizeof(h));
  *local_ptr = c->mailbox;
  *pid = static_cast<int64_t>(getpid());
  return B200_OK;
}

extern "C" int b200_p2p_connect(b200_ctx* c, const int* device_ids, int n, const void* ipc_handles, const uint64_t* local_ptrs,
                                const int64_t* pids) {
  CTX_ENTER_DEVICE(c);
  if (!device_ids || n < 1 || n > 8 || !ipc_handles || !local_ptrs || !pids)
    return fail(B200_ERR_INVALID_ARG, "p2p_connect: bad arguments (1..8 devices)");
  int rc = ensure_mailbox(c);
  if (rc) return rc;
  // the arrays are indexed like device_ids; ranks follow the SORTED device set (same keying as b200_comm_init)
  std::vector<int> key = sorted_ids(device_ids, n);
  if (c->p2p.count(key)) return B200_OK;
  P2PState st;
  st.n = n;
  unsigned mask = 0;
  for (int id : key) mask |= 1u << (static_cast<unsigned>(id) & 7u);
  const uint64_t set_off = static_cast<uint64_t>(mask & 0xFFu) * kMailboxSetBytes;
  for (int r = 0; r < n; ++r) {
    int src = -1;
    for (int j = 0; j < n; ++j)
      if (device_ids[j] == key[r]) src = j;
    if (key[r] == c->device) {
      st.rank = r;
      st.mailbox[r] = c->mailbox + set_off;
      continue;
    }
    if (pids[src] == static_cast<int64_t>(getpid())) {
      // same process (the reference's one-process model): enable peer access to that device's primary context
      CUdevice pd;
      CUcontext pctx;
      CU_CHECK(g_drv.cuDeviceGet_p(&pd, key[r]));
      int can = 0;
      CU_CHECK(g_drv.cuDeviceCanAccessPeer_p(&can, c->dev, pd));
      if (!can) return fail(B200_ERR_UNSUPPORTED, "p2p_connect: device %d cannot access device %d", c->device, key[r]);
      CU_CHECK(g_drv.cuDevicePrimaryCtxRetain_p(&pctx, pd));
      CUresult r2 = g_drv.cuCtxEnablePeerAccess_p(pctx, 0);
      g_drv.cuDevicePrimaryCtxRelease_p(pd);
      if (r2 != CUDA_SUCCESS && r2 != CUDA_ERROR_PEER_ACCESS_ALREADY_ENABLED)
        return fail(map_cu(r2), "cuCtxEnablePeerAccess(%d) failed: %s", key[r], cu_err(r2));
      st.mailbox[r] = local_ptrs[src] + set_off;
    } else {
      CUipcMemHandle h;
      memcpy(&h, static_cast<const char*>(ipc_handles) + static_cast<size_t>(src) * B200_IPC_HANDLE_BYTES, sizeof(h));
      CUdeviceptr mapped = 0;
      CUresult r2 = g_drv.cuIpcOpenMemHandle_p(&mapped, h, CU_IPC_MEM_LAZY_ENABLE_PEER_ACCESS);
      if (r2 != CUDA_SUCCESS) return fail(map_cu(r2), "cuIpcOpenMemHandle(device %d) failed: %s", key[r], cu_err(r2));
      st.opened.push_back(mapped);
      st.mailbox[r] = mapped + set_off;
    }
  }
  if (st.rank < 0) return fail(B200_ERR_INVALID_ARG, "p2p_connect: device %d is not in the device set", c->device);
  c->p2p[key] = st;
  return B200_OK;
}

static int reduce_all_reduce_impl(b200_ctx* c, b200_stream s, b200_reduce_op op, b200_dtype in_dtype, b200_dptr in, b200_dptr out,
                                  uint64_t n, uint64_t index_offset, const int* device_ids, int ndev) {
  const bool arg = (op == B200_REDUCE_ARGMAX || op == B200_REDUCE_ARGMIN);
  if (op != B200_REDUCE_SUM && !arg) return fail(B200_ERR_UNSUPPORTED, "reduce_all_reduce: SUM, ARGMAX and ARGMIN are fused");
  if (in_dtype != B200_F32) return fail(B200_ERR_UNSUPPORTED, "reduce_all_reduce: only f32 input is fused");
  if (!device_ids || ndev < 1) return fail(B200_ERR_INVALID_ARG, "reduce_all_reduce: bad device set");
  if (!in || !out || n == 0) return fail(B200_ERR_INVALID_ARG, "reduce_all_reduce: null pointer or empty input");
  if (arg && index_offset + n > (1ull << 32)) return fail(B200_ERR_UNSUPPORTED, "reduce_all_reduce: global indices must fit 32 bits");
  auto it = c->p2p.find(sorted_ids(device_ids, ndev));
  if (it == c->p2p.end()) return fail(B200_ERR_COMM, "reduce_all_reduce: device set not connected (call b200_p2p_connect)");
  P2PState& st = it->second;
  CUstream cs = resolve_stream(c, s);
  CUfunction f;
  int rc = get_func(c, op == B200_REDUCE_SUM ? "reduce_all_sum_f32_xgpu" : op == B200_REDUCE_ARGMAX ? "reduce_all_argmax_f32_xgpu" : "reduce_all_argmin_f32_xgpu", &f);
  if (rc) return rc;
  CUdeviceptr ws;
  rc = reduce_workspace(c, cs, &ws);
  if (rc) return rc;
  unsigned threads = (unsigned)std::min(512, std::max(32, atoi(opt(c, "reduce.threads", "512").c_str()))) / 32 * 32;
  unsigned bps = (unsigned)std::max(1, atoi(opt(c, "reduce.blocks_per_sm", "4").c_str()));
  const uint64_t want = (n / 4 + threads - 1) / threads;
  unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, std::min<uint64_t>((uint64_t)c->props.num_sms * bps, kWsMaxBlocks)));
  ReduceParams p{};
  p.in = in; p.out = out; p.ws = ws;
  p.outer = 1; p.len = n; p.inner = 1;
  p.row_len = n; p.row_pitch = n; p.seg_len = n; p.nseg = 1; p.scale = 1.0f;
  p.flags = opt(c, "reduce.debug", "0") == "1" ? 1u : 0u;
  XgpuParams xg;
  memset(&xg, 0, sizeof(xg));
  for (int r = 0; r < st.n; ++r) xg.mailbox[r] = st.mailbox[r];
  xg.rank = (uint32_t)st.rank;
  xg.nranks = (uint32_t)st.n;
  xg.index_offset = index_offset;
  xg.epoch = st.epoch + 1;  // every rank calls in the same order (collective semantics), so epochs agree
  void* args[] = {&p, &xg};
  const bool pdl = cs == c->stream && c->pdl_prev_out != 0 && opt(c, "reduce.pdl", "on") == "on" &&
                   !(c->pdl_prev_out + 4 > in && c->pdl_prev_out < in + n * 4);
  rc = launch(c, f, grid, 1, 1, threads, 0, 1, cs, args, pdl);
  if (!rc && cs == c->stream) c->pdl_prev_out = out;
  if (!rc) st.epoch += 1;   // only a launch that really went out consumes the epoch (a failed call must not desynchronise the ranks)
  return rc;
}

extern "C" int b200_reduce_all_reduce(b200_ctx* c, b200_stream s, b200_reduce_op op, b200_dtype in_dtype, b200_dptr in,
                                      b200_dptr out, uint64_t n, const int* device_ids, int ndev) {
  CTX_ENTER(c);
  if (op != B200_REDUCE_SUM) return fail(B200_ERR_UNSUPPORTED, "reduce_all_reduce: SUM only (use b200_argreduce_all_reduce for arg ops)");
  return reduce_all_reduce_impl(c, s, op, in_dtype, in, out, n, 0, device_ids, ndev);
}

extern "C" int b200_argreduce_all_reduce(b200_ctx* c, b200_stream s, b200_reduce_op op, b200_dtype in_dtype, b200_dptr in,
                                         b200_dptr out, uint64_t n, uint64_t index_offset, const int* device_ids, int ndev) {
  CTX_ENTER(c);
  if (op != B200_REDUCE_ARGMAX && op != B200_REDUCE_ARGMIN) return fail(B200_ERR_INVALID_ARG, "argreduce_all_reduce: ARGMAX or ARGMIN");
  return reduce_all_reduce_impl(c, s, op, in_dtype, in, out, n, index_offset, device_ids, ndev);
}

// ================================================================================================ generators / probes
static int launch_fill(b200_ctx* c, b200_stream s, int dtype, uint64_t out, uint64_t n, FillParams p) {
  if (!fill_dtype_ok(dtype)) return fail(B200_ERR_UNSUPPORTED, "fill: dtype %d unsupported", dtype);
  if (n == 0) return B200_OK;
  CUfunction f;
  int rc = get_func(c, "fill_kernel", &f);
  if (rc) return rc;
  p.out = out; p.n = n; p.dtype = (uint32_t)dtype;
  const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)c->props.num_sms * 32);
  void* args[] = {&p};
  return launch(c, f, grid, 1, 1, 256, 0, 1, resolve_stream(c, s), args);
}

extern "C" int b200_fill_uniform(b200_ctx* c, b200_stream s, b200_dtype dtype, b200_dptr out, uint64_t n, uint64_t seed, float lo, float hi) {
  CTX_ENTER(c);
  FillParams p{};
  p.seed = seed; p.lo = lo; p.scale = hi - lo; p.mode = 0; p.modulus = 1;
  return launch_fill(c, s, dtype, out, n, p);
}

extern "C" int b200_fill_modulo(b200_ctx* c, b200_stream s, b200_dtype dtype, b200_dptr out, uint64_t n, uint32_t modulus) {
  CTX_ENTER(c);
  if (modulus == 0) return fail(B200_ERR_INVALID_ARG, "fill_modulo: modulus 0");
  FillParams p{};
  p.mode = 1; p.modulus = modulus;
  return launch_fill(c, s, dtype, out, n, p);
}

extern "C" int b200_probe_wmma(b200_ctx* c, b200_stream s, b200_dtype dtype, uint32_t n_iter, b200_dptr scratch, double* ops) {
  CTX_ENTER(c);
  CUfunction f;
  int rc = get_func(c, dtype == B200_BF16 ? "wmma_probe_bf16" : "wmma_probe_f16", &f);
  if (rc) return rc;
  const unsigned grid = (unsigned)c->props.num_sms * 32, block = 256;
  uint64_t sp = scratch;
  void* args[] = {&sp, &n_iter};
  rc = launch(c, f, grid, 1, 1, block, 0, 1, resolve_stream(c, s), args);
  if (!rc && ops) *ops = static_cast<double>(grid) * (block / 32) * 2.0 * 16 * 16 * 16 * n_iter;
  return rc;
}

extern "C" int b200_probe_umma(b200_ctx* c, b200_stream s, uint32_t n_iter, b200_dptr scratch, double* ops) {
  CTX_ENTER_DEVICE(c);
  CUfunction f;
  int rc = get_func(c, "umma_probe_bf16_2sm", &f);
  if (rc) return rc;
  const unsigned smem = 32768 + 1024 + 64;
  CU_CHECK(g_drv.cuFuncSetAttribute_p(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
  const unsigned clusters = (unsigned)std::max(1, c->props.num_sms / 2);
  uint64_t sp = scratch;
  void* args[] = {&sp, &n_iter};
  rc = launch(c, f, clusters * 2, 1, 1, 256, smem, 2, resolve_stream(c, s), args);
  if (!rc && ops) *ops = static_cast<double>(clusters) * n_iter * 4.0 * 2.0 * 256 * 256 * 16;
  return rc;
}

extern "C" int b200_probe_umma_kind(b200_ctx* c, b200_stream s, b200_dtype dtype, int block_scaled, uint32_t n_iter, b200_dptr scratch,
                                    double* ops) {
  CTX_ENTER_DEVICE(c);
  if (dtype == B200_BF16 && !block_scaled) return b200_probe_umma(c, s, n_iter, scratch, ops);
  const char* name = (dtype == B200_F8E4M3 && !block_scaled) ? "umma_probe_e4m3_2sm"
                     : (dtype == B200_F8E4M3 && block_scaled) ? "umma_probe_mxf8_2sm"
                     : (dtype == B200_F4E2M1X2 && block_scaled) ? "umma_probe_mxf4_2sm" : nullptr;
  if (!name) return fail(B200_ERR_UNSUPPORTED, "probe_umma_kind: bf16, f8e4m3 (plain or block-scaled) or block-scaled f4e2m1x2");
  CUfunction f;
  int rc = get_func(c, name, &f);
  if (rc) return rc;
  const unsigned smem = 32768 + 2048 + 1024 + 64;
  CU_CHECK(g_drv.cuFuncSetAttribute_p(f, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem));
  const unsigned clusters = (unsigned)std::max(1, c->props.num_sms / 2);
  uint64_t sp = scratch;
  void* args[] = {&sp, &n_iter};
  rc = launch(c, f, clusters * 2, 1, 1, 256, smem, 2, resolve_stream(c, s), args);
  const double k_per_instr = dtype == B200_F4E2M1X2 ? 64.0 : 32.0;
  if (!rc && ops) *ops = static_cast<double>(clusters) * n_iter * 4.0 * 2.0 * 256 * 256 * k_per_instr;
  return rc;
}

extern "C" int b200_probe_memread(b200_ctx* c, b200_stream s, b200_dptr buf, uint64_t bytes, b200_dptr scratch) {
  CTX_ENTER(c);
  CUfunction f;
  int rc = get_func(c, "memread_probe_vec4", &f);
  if (rc) return rc;
  const unsigned grid = (unsigned)c->props.num_sms * 32, block = 256;
  uint64_t lines = bytes / 16;
  const uint64_t per_pass = static_cast<uint64_t>(grid) * block;
  uint32_t steps = (uint32_t)((lines + per_pass - 1) / per_pass);
  uint64_t in = buf, out = scratch;
  void* args[] = {&in, &out, &lines, &steps};
  return launch(c, f, grid, 1, 1, block, 0, 1, resolve_stream(c, s), args);
}

// mode 0: write-only (memory_write.rs), mode 1: copy (memory_direct.rs); `bytes` per buffer
extern "C" int b200_probe_memwrite(b200_ctx* c, b200_stream s, b200_dptr dst, uint64_t bytes) {
  CTX_ENTER(c);
  CUfunction f;
  int rc = get_func(c, "memwrite_probe_vec4", &f);
  if (rc) return rc;
  const unsigned grid = (unsigned)c->props.num_sms * 32, block = 256;
  uint64_t lines = bytes / 16, out = dst;
  const uint64_t per_pass = static_cast<uint64_t>(grid) * block;
  uint32_t steps = (uint32_t)((lines + per_pass - 1) / per_pass);
  void* args[] = {&out, &lines, &steps};
  return launch(c, f, grid, 1, 1, block, 0, 1, resolve_stream(c, s), args);
}

extern "C" int b200_probe_memcopy(b200_ctx* c, b200_stream s, b200_dptr dst, b200_dptr src, uint64_t bytes) {
  CTX_ENTER(c);
  CUfunction f;
  int rc = get_func(c, "memcopy_probe_vec4", &f);
  if (rc) return rc;
  const unsigned grid = (unsigned)c->props.num_sms * 32, block = 256;
  uint64_t lines = bytes / 16, in = src, out = dst;
  const uint64_t per_pass = static_cast<uint64_t>(grid) * block;
  uint32_t steps = (uint32_t)((lines + per_pass - 1) / per_pass);
  void* args[] = {&in, &out, &lines, &steps};
  return launch(c, f, grid, 1, 1, block, 0, 1, resolve_stream(c, s), args);
}
