/* cubecl_b200.h -- C ABI of the B200-native dense linear-algebra hot path (tcgen05 matmul + HBM-bound reduction).
 *
 * This is the drop-in boundary a CubeCL maintainer binds from Rust (see INTEGRATION.md for the `extern "C"` block and the
 * `CudaServer` hook).  Plain pointers and sizes only; no torch / C++ types.  All device pointers are CUdeviceptr values
 * carried as uint64_t, all streams are CUstream carried as void* (NULL = the context's own compute stream).
 *
 * Citations are into the reference tree (tracel-ai/cubecl @ 4057f39e), i.e. the interface each entry point replaces.
 *
 * Threading (crates/cubecl-common/src/device/handle/mod.rs:18-24): one thread per device at a time; distinct contexts
 * are independent.  The only process-global state is the dlopen'ed driver/NCCL symbol tables and the last-error string,
 * which is thread-local.
 *
 * Errors (crates/cubecl-runtime/src/server/base.rs:177-272): every call returns a b200_status; the message is available
 * from b200_last_error() on the calling thread.  Asynchronous device faults surface at the next b200_sync()/b200_read(),
 * like ServerError::ServerUnhealthy does at sync/read (crates/cubecl-cuda/src/compute/server.rs:981-1022).
 *
 * There is no CPU fallback anywhere behind this header: without a CUDA driver and an sm_100 device b200_init() fails.
 */
#ifndef CUBECL_B200_H
#define CUBECL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

typedef struct b200_ctx b200_ctx;
typedef uint64_t b200_dptr;   /* CUdeviceptr */
typedef void* b200_stream;    /* CUstream; NULL = context compute stream */
typedef void* b200_event;     /* CUevent */

/* Mirrors LaunchError / ServerError variants (server/base.rs:177-272) so a Rust shim can map 1:1. */
typedef enum b200_status {
  B200_OK = 0,
  B200_ERR_COMPILATION = 1,        /* LaunchError::CompilationError  -- here: cubin image failed to load */
  B200_ERR_OUT_OF_MEMORY = 2,      /* LaunchError::OutOfMemory */
  B200_ERR_TOO_MANY_RESOURCES = 3, /* LaunchError::TooManyResources (smem / units / cube dim) */
  B200_ERR_UNKNOWN = 4,            /* LaunchError::Unknown */
  B200_ERR_IO = 5,                 /* LaunchError::IoError / IoError::* */
  B200_ERR_INVALID_ARG = 6,        /* shape/stride/dtype validation failed before launch */
  B200_ERR_UNSUPPORTED = 7,        /* feature absent (dtype/layout not implemented) -- callers self-skip like runtime_tests do */
  B200_ERR_NO_DEVICE = 8,          /* no driver / no sm_100 device: fail loudly, never fall back */
  B200_ERR_COMM = 9,               /* NCCL failure -> ServerError::Generic (server.rs:773-776) */
  B200_ERR_UNHEALTHY = 10          /* deferred device fault surfaced at sync -> ServerError::ServerUnhealthy */
} b200_status;

/* Element types.  Values double as the NCCL dtype selector of communication.rs:34-108 for all_reduce. */
typedef enum b200_dtype {
  B200_F32 = 0, B200_F16 = 1, B200_BF16 = 2, B200_U32 = 3, B200_I32 = 4, B200_F64 = 5, B200_I64 = 6, B200_U64 = 7,
  B200_U8 = 8, B200_I8 = 9,
  B200_F8E4M3 = 10, B200_F8E5M2 = 11,  /* fp8 matmul inputs (FloatKind::E4M3 / E5M2; manual-MMA dtypes of cuda/mma/manual.rs:108-186) */
  B200_F4E2M1X2 = 12,                  /* two e2m1 per byte, element 2i in the low nibble (e2m1x2, cubecl-common/src/float/fp4.rs:28,204-216) */
  B200_UE8M0 = 13                      /* block scale 2^(bits-127) (FloatKind::UE8M0; scales_type of ScaledMmaConfig) */
} b200_dtype;

/* Reduction instructions of the `reduce::launch` surface (cubek); in-tree semantics: examples/sum_things/src/lib.rs:6-33,
 * cubecl-book/.../v1-cpu.rs:7-15.  Arg ops: ties -> lowest index, NaN is the extreme and the first NaN wins. */
typedef enum b200_reduce_op {
  B200_REDUCE_SUM = 0, B200_REDUCE_PROD = 1, B200_REDUCE_MAX = 2, B200_REDUCE_MIN = 3,
  B200_REDUCE_ARGMAX = 4, B200_REDUCE_ARGMIN = 5, B200_REDUCE_MEAN = 6
} b200_reduce_op;

/* ReduceOperation{Sum,Mean} -- crates/cubecl-runtime/src/server/base.rs:623-628 */
typedef enum b200_comm_op { B200_COMM_SUM = 0, B200_COMM_MEAN = 1 } b200_comm_op;

/* HardwareProperties subset -- crates/cubecl-ir/src/properties.rs:26-58, probed like cubecl-cuda/src/runtime.rs:52-350 */
typedef struct b200_props {
  int32_t device;
  int32_t cc_major, cc_minor;
  int32_t num_sms;                 /* num_streaming_multiprocessors */
  int32_t max_shared_per_block;    /* opt-in maximum (max_shared_memory_size) */
  int32_t clock_khz, mem_clock_khz;
  int32_t plane_size;              /* 32 */
  uint64_t total_mem;
  char name[128];
} b200_props;

/* ---- lifecycle: R::client(device) -> DeviceService::init (cubecl-cuda/src/runtime.rs:52-350) ------------------------ */
int b200_abi_version(void);
int b200_device_count(int* count);
/* The embedded prebuilt sm_100a images ("gemm" | "gemm_b" | "gemm_c" | "gemm_mx" | "reduce" | "aux"), for a host that prefers to cuModuleLoadData them into
 * its own module cache (CudaContext::modules, crates/cubecl-cuda/src/compute/context.rs:38-62,293). No GPU needed. */
int b200_get_cubin(const char* name, const void** image, size_t* size);
int b200_init(int device, b200_ctx** out);   /* cuInit, primary ctx retain, load the embedded sm_100a cubins (context.rs:293) */
int b200_destroy(b200_ctx* ctx);
int b200_get_props(b200_ctx* ctx, b200_props* out);
/* Dry-run planning context (DryRun, crates/cubecl-runtime/src/dry_run.rs:45,88,121): needs no driver and no device.  Ops
 * called on it (b200_matmul, b200_reduce*, b200_alloc, ...) validate and plan exactly like a real context but RECORD each
 * pooled allocation, TMA descriptor and kernel launch as one text line instead of executing it.  b200_plan_text copies the
 * log (and clears it when it fit); *needed = bytes required.  Device pointers passed to ops may be any non-zero values. */
int b200_plan_begin(int num_sms, b200_ctx** out);
int b200_plan_text(b200_ctx* ctx, char* buf, size_t capacity, size_t* needed);
/* Runtime knobs, string-typed like cubecl.toml keys (config/base.rs:18-120).  Keys: "gemm.variant"
 * (auto|2sm_m512|2sm_n256|2sm_n128|1sm_n128|simt; 2sm_n256a1 = single-accumulator diagnostic), "gemm.f32" (hybrid|3xtf32|tf32: f32 inputs as one
 * tf32 pass + two bf16 cross-term passes in ONE launch (default, ~2^-20 of the product), three tf32 passes, or one), "gemm.sf_copy" (thread|thread2|mma:
 * block-scaled kinds, who issues the scale-factor copies to TMEM -- the dedicated copy thread, two of them, or the MMA thread for an A/B), "gemm.group_m",
 * "gemm.l2_promotion" (256|128|64|0: TMA L2 promotion bytes of the operand tensor maps), "gemm.split_k" (auto|off|on|1..8:
 * deterministic stream-K head -- the tiles of a partial last wave are cut along K into equal ranges that run FIRST, slabs
 * added in k order; N = ranges per tile), "gemm.epilogue" (tma|direct), "gemm.stage" (on|off: operands TMA cannot describe --
 * unaligned row pitch / base -- are first copied into an aligned pooled buffer and run on the tensor cores; off = strided SIMT kernel), "reduce.variant"
 * (auto|u2|u4|u8|u16|b4|b8|w2|w4: load-unroll / blocked / 256-bit forms of the all-elements kernel; tma: 16 KB bulk copies
 * into a shared-memory ring), "reduce.threads", "reduce.blocks_per_sm" (all-elements kernel), "reduce.rows_vpt" (128-bit
 * vectors per thread that size the threads-per-row of the row kernel), "reduce.rows_blocks_per_sm" /
 * "reduce.cols_blocks_per_sm" (below this many blocks per SM a long reduced axis is cut into segments: two passes),
 * "reduce.debug" (1: the fused reduce + exchange records its stage timings, see b200_reduce_debug), "reduce.pdl" (on|off:
 * back-to-back all-element reductions on the context's own stream overlap through programmatic dependent launch -- the next
 * launch streams its input while the previous one's last block finishes; results are unchanged). */
int b200_set_option(b200_ctx* ctx, const char* key, const char* value);
/* Number of device kernels this context has launched so far (bench.py reports it as gpu_launches). */
int b200_launch_count(b200_ctx* ctx, uint64_t* count);
/* Name of the kernel (cubin entry point) the context launched most recently, NUL-terminated, truncated to `capacity`:
 * what a measurement harness reports as the kernel it timed (bench.py: roofline.kernel). */
int b200_last_kernel(b200_ctx* ctx, char* buf, size_t capacity);

/* ---- memory: ComputeClient::{empty,create_from_slice,read_one} (cubecl-runtime/src/client.rs:654,452,256) ----------- */
/* Pooled device allocation, 512-byte aligned (mem_alignment, cubecl-cuda/src/runtime.rs:81). */
int b200_alloc(b200_ctx* ctx, size_t bytes, b200_dptr* out);
/* b200_free returns the page to the pool as of the CONTEXT's stream: a buffer that work queued on another stream (the
 * `s` argument of the compute / copy entry points) may still touch must be freed with b200_free_async naming that stream --
 * the pool then re-issues the page only after an event recorded there has completed (the reference binds pool memory to
 * its stream and waits on cross-stream events, crates/cubecl-runtime/src/stream/event.rs:50-57). */
int b200_free(b200_ctx* ctx, b200_dptr ptr);
int b200_free_async(b200_ctx* ctx, b200_dptr ptr, b200_stream last_use);
int b200_memory_usage(b200_ctx* ctx, uint64_t* bytes_in_use, uint64_t* bytes_reserved);  /* MemoryUsage, memory_management/base.rs:7-28 */
int b200_memory_cleanup(b200_ctx* ctx);                                                  /* client.memory_cleanup */
/* Pinned host staging (compute/stream.rs:138-178 pinned pool). */
int b200_host_alloc(b200_ctx* ctx, size_t bytes, void** out);
int b200_host_free(b200_ctx* ctx, void* ptr);
/* Stream-ordered copies.  b200_write/read are asynchronous when `host` is pinned; call b200_sync before reusing `host`. */
int b200_write(b200_ctx* ctx, b200_stream s, b200_dptr dst, const void* host_src, size_t bytes);
int b200_read(b200_ctx* ctx, b200_stream s, void* host_dst, b200_dptr src, size_t bytes);
int b200_copy(b200_ctx* ctx, b200_stream s, b200_dptr dst, b200_dptr src, size_t bytes);
int b200_memset32(b200_ctx* ctx, b200_stream s, b200_dptr dst, uint32_t value, size_t words);

/* ---- streams / sync / timing: ComputeClient::sync (client.rs:1013), Fence = CUevent (compute/sync/fence.rs) ---------- */
int b200_stream_create(b200_ctx* ctx, b200_stream* out);
int b200_stream_destroy(b200_ctx* ctx, b200_stream s);
int b200_sync(b200_ctx* ctx, b200_stream s);
int b200_event_create(b200_ctx* ctx, b200_event* out);
int b200_event_record(b200_ctx* ctx, b200_event e, b200_stream s);
int b200_stream_wait_event(b200_ctx* ctx, b200_stream s, b200_event e);   /* cross-stream dependency (MultiStream::resolve, stream/event.rs:220) */
int b200_event_elapsed_ms(b200_ctx* ctx, b200_event start, b200_event end, float* ms);
int b200_event_destroy(b200_ctx* ctx, b200_event e);

/* ---- matmul::launch (cubek; shape rule crates/cubecl-zspace/src/shape.rs:489-517) ----------------------------------
 * out[..,m,n] = sum_k lhs[..,m,k] * rhs[..,k,n]; equal rank >= 2, leading dims broadcast (1 vs d); shapes/strides in
 * ELEMENTS (TensorHandle, cubecl-std/src/tensor/handle.rs:13-23).  f32 accumulation over k.  Inputs f16/bf16/f32 with
 * `out_dtype` = the input dtype or F32; fp8 inputs (F8E4M3 / F8E5M2, both operands the same format, kind::f8f6f4) with
 * `out_dtype` BF16, F16 or F32; U8 / I8 inputs (kind::i8) with exact I32 accumulation and `out_dtype` I32.  Row strides may be pitched (allocator.rs:21-72); rhs may be given transposed
 * (stride_k == 1, MatrixBatchLayout::MildlyPermuted{transposed}, matrix_batch_layout.rs:8-19).  f32 inputs run on the tf32
 * tensor pipe, by default with the hybrid split (tf32 main product + bf16 cross terms, f32-grade accuracy; see "gemm.f32").
 * Returns B200_ERR_INVALID_ARG on shape mismatch -- the MatmulShapeError of shape.rs:489-517. */
int b200_matmul(b200_ctx* ctx, b200_stream s, b200_dtype in_dtype, b200_dtype out_dtype,
                b200_dptr lhs, b200_dptr rhs, b200_dptr out, int rank,
                const uint64_t* shape_lhs, const uint64_t* strides_lhs,
                const uint64_t* shape_rhs, const uint64_t* strides_rhs,
                const uint64_t* shape_out, const uint64_t* strides_out);
/* The same product with DIFFERENT 8-bit formats for the two operands -- the pairs the reference instantiates for its manual
 * MMA: i8 x u8 / u8 x i8 -> i32 (crates/cubecl-cpp/src/cuda/mma/manual.rs:151-166) and fp8 e4m3 x e5m2 / e5m2 x e4m3
 * (:170-186).  Same tcgen05 kernels (kind::i8 / kind::f8f6f4 take one format field per operand in the instruction
 * descriptor); equal formats behave exactly like b200_matmul.  Other combinations: B200_ERR_UNSUPPORTED. */
int b200_matmul_mixed(b200_ctx* ctx, b200_stream s, b200_dtype lhs_dtype, b200_dtype rhs_dtype, b200_dtype out_dtype,
                      b200_dptr lhs, b200_dptr rhs, b200_dptr out, int rank,
                      const uint64_t* shape_lhs, const uint64_t* strides_lhs,
                      const uint64_t* shape_rhs, const uint64_t* strides_rhs,
                      const uint64_t* shape_out, const uint64_t* strides_out);

/* Fused epilogue: out = act(alpha * (lhs @ rhs) + bias[n]) applied to the f32 accumulators inside the GEMM
 * epilogue (TMEM -> registers -> here -> store), no extra pass over the output.  bias: f32[N] device pointer or 0.
 * activation: 0 none, 1 relu, 2 gelu (erf form).  Float inputs only. */
typedef struct b200_epilogue {
  float alpha;
  int32_t activation;
  b200_dptr bias;
} b200_epilogue;
int b200_matmul_fused(b200_ctx* ctx, b200_stream s, b200_dtype in_dtype, b200_dtype out_dtype,
                      b200_dptr lhs, b200_dptr rhs, b200_dptr out, int rank,
                      const uint64_t* shape_lhs, const uint64_t* strides_lhs,
                      const uint64_t* shape_rhs, const uint64_t* strides_rhs,
                      const uint64_t* shape_out, const uint64_t* strides_out, const b200_epilogue* epilogue);

/* Block-scaled (MX) matmul: out[b,m,n] = sum_k (lhs[b,m,k] * lhs_scales[b,m,k/32]) * (rhs[b,n,k] * rhs_scales[b,n,k/32]),
 * f32 accumulation.  Replaces MmaDefinition::new_scaled / execute_scaled (crates/cubecl-core/src/frontend/cmma.rs:438-460,
 * 798-840), the ScaledMmaConfig feature rows (crates/cubecl-ir/src/features.rs:190-211; on CUDA the reference offers them
 * for sm_120 only, cubecl-cpp/src/cuda/mma/manual.rs:201-255 -- sm_100 needs tcgen05 kind::mxf8f6f4 / kind::mxf4) and is
 * pinned by test_cmma_scaled / test_cmma_scaled_fp4 (crates/cubecl-core/src/runtime_tests/cmma.rs:1476-1700).
 * Layouts follow those tests: lhs [batch, m, k] and rhs [batch, n, k] K-contiguous ("col-major" rhs), dtypes B200_F8E4M3 /
 * B200_F8E5M2 (mixable) or both B200_F4E2M1X2 (k / 2 bytes per row); scales are B200_UE8M0 bytes [batch, rows, k / 32]
 * row-major (scales_packed = 0) or already in the tensor core's packed form [batch * ceil(rows/128)][ceil(k/128)][512 B],
 * byte (r % 32) * 16 + (r / 32) * 4 + s (scales_packed = 1).  out [batch, m, n] contiguous, f32 / bf16 / f16.
 * scale_block = 32: ue8m0 scales (MXFP8 / MXFP4).  scale_block = 16: NVFP4 -- packed e2m1 operands with e4m3 scale bytes
 * [batch, rows, k / 16] whose sign is ignored (the third ScaledMmaConfig row of manual.rs:241-250; kind::mxf4nvf4).
 * k must be a multiple of 32.  Operands that TMA cannot describe take the reference-order SIMT path. */
int b200_matmul_scaled(b200_ctx* ctx, b200_stream stream, b200_dtype lhs_dtype, b200_dtype rhs_dtype, b200_dtype out_dtype,
                       b200_dptr lhs, b200_dptr rhs, b200_dptr lhs_scales, b200_dptr rhs_scales, b200_dptr out,
                       uint64_t batch, uint64_t m, uint64_t n, uint64_t k, int scale_block, int scales_packed);

/* ---- reduce::launch (cubek) -----------------------------------------------------------------------------------------
 * Reduces `axis` (0..rank-1) of a CONTIGUOUS row-major input, or every element when axis == -1.  Output is contiguous
 * with the reduced axis removed (one element for axis == -1): F32 values, or U32 indices along the axis for arg ops.
 * Input dtype F32 / F16 / BF16, f32 accumulation; `in` needs element alignment only (a sub-slice view is fine: the kernels
 * peel scalar head / tail elements around the 128-bit body).  One launch, no host sync -- two launches when few outputs
 * meet a long axis (segments first, then the partials; deterministic); uses a per-stream workspace owned by ctx. */
int b200_reduce(b200_ctx* ctx, b200_stream s, b200_reduce_op op, b200_dtype in_dtype,
                b200_dptr in, b200_dptr out, int rank, const uint64_t* shape, int axis);

/* Same, for an input described by strides in elements.  Pitched rows (TensorHandle::empty -> PitchedMemoryLayoutPolicy,
 * crates/cubecl-runtime/src/allocator.rs:21-72) and axis permutations that keep the kept axes in order (a transposed view)
 * are reduced IN PLACE -- the kernels take the outer / axis strides and the row pitch, so the traffic is 1x the logical
 * bytes and the padding is never read.  Only views no such description fits (broadcast strides, gaps between outer
 * dimensions, permuted outputs) are first gathered into a pooled compact temporary (into_contiguous,
 * crates/cubecl-std/src/tensor/contiguous.rs).  strides == NULL means contiguous. */
int b200_reduce_strided(b200_ctx* ctx, b200_stream s, b200_reduce_op op, b200_dtype in_dtype,
                        b200_dptr in, b200_dptr out, int rank, const uint64_t* shape, const uint64_t* strides, int axis);
/* Stage timings of the most recent fused reduce + exchange launched on `s` with option "reduce.debug" = 1, in ns:
 * words4[0] = exchange (publish to the peers' mailboxes -> every peer's value seen), words4[1] = partials + f64 tree of the
 * last block; words4[2..3] reserved.  Synchronises the stream. */
int b200_reduce_debug(b200_ctx* ctx, b200_stream s, uint64_t* words4);
/* out (compact row-major) = gather of the strided rank<=8 tensor `in`. */
int b200_into_contiguous(b200_ctx* ctx, b200_stream s, b200_dtype dtype, b200_dptr in, b200_dptr out, int rank,
                         const uint64_t* shape, const uint64_t* strides);

/* ---- collectives: ServerCommunication (server/base.rs:632-739), CUDA impl cubecl-cuda/src/compute/server.rs:666-926 -- */
#define B200_UNIQUE_ID_BYTES 128
int b200_comm_get_unique_id(b200_ctx* ctx, void* id128);            /* ncclGetUniqueId (communication.rs:11-25 holds it per device set) */
/* Communicator for the device set `device_ids` (keyed by the sorted set, CommunicationId server/base.rs:605-620);
 * this context's rank is the position of its device in the sorted list (server.rs:669-703). */
int b200_comm_init(b200_ctx* ctx, const int* device_ids, int n, const void* id128);
/* src -> dst (may alias) of bytes/elem_size elements on the communicator's comm stream, after an event wait on
 * `compute` (server.rs:705-780). */
int b200_all_reduce(b200_ctx* ctx, b200_stream compute, b200_dptr src, b200_dptr dst, size_t bytes, b200_dtype dtype,
                    b200_comm_op op, const int* device_ids, int n);
/* Make `compute` wait for everything issued on the comm stream (server.rs:782-798). */
int b200_sync_collective(b200_ctx* ctx, b200_stream compute);

/* ---- fused reduce + all-reduce over NVLink peer memory (no NCCL on the data path) -----------------------------------
 * The reduce path's exchange step is one f32 per rank.  Instead of reduce kernel -> event -> ncclAllReduce -> event
 * (client.rs:790 + server.rs:705-798), the LAST block of the reduce kernel stores this rank's scalar (value+epoch in one
 * 64-bit system-scope store) into every peer's mailbox and gathers theirs, so local reduce + exchange are ONE launch.
 * Setup: every rank exports its mailbox, the handles are exchanged out of band, every rank connects.  Ranks = position in
 * the sorted device set (as for b200_comm_init).  Works across processes (CUDA IPC) and inside one process (peer access). */
#define B200_IPC_HANDLE_BYTES 64
int b200_p2p_export(b200_ctx* ctx, void* ipc_handle64, uint64_t* local_ptr, int64_t* pid);
/* arrays are indexed like device_ids: n x 64-byte handles, n pointers, n pids (as returned by b200_p2p_export on each rank) */
int b200_p2p_connect(b200_ctx* ctx, const int* device_ids, int n, const void* ipc_handles, const uint64_t* local_ptrs,
                     const int64_t* pids);
/* out[0] (on every rank) = sum over ranks of sum(in[0..n)).  Collective: every rank of the set must call it, in the same
 * order; a missing peer trips the kernel's 4 s deadline (device trap -> B200_ERR_UNHEALTHY at sync).  SUM of F32 only. */
int b200_reduce_all_reduce(b200_ctx* ctx, b200_stream s, b200_reduce_op op, b200_dtype in_dtype, b200_dptr in, b200_dptr out,
                           uint64_t n, const int* device_ids, int ndev);

/* Arg-reduce across ranks (NCCL has no arg-reduce: SURVEY 8e "all-gather 8 pairs + local select", here inside the kernel):
 * out[0] (u32, on every rank) = GLOBAL index of the extremum of the concatenation of all ranks' inputs, where this rank's
 * element 0 has global index `index_offset`.  Same tie / NaN rule as b200_reduce.  Global indices must fit 32 bits. */
int b200_argreduce_all_reduce(b200_ctx* ctx, b200_stream s, b200_reduce_op op, b200_dtype in_dtype, b200_dptr in, b200_dptr out,
                              uint64_t n, uint64_t index_offset, const int* device_ids, int ndev);

/* ---- synthetic operands + reference-equivalent probes (examples/throughput) ----------------------------------------- */
/* out[i] = lo + u(seed,i) * (hi - lo), u in [0,1) from a counter hash the host can reproduce (cubecl_b200/synth.py);
 * mode 1: out[i] = i % modulus. */
int b200_fill_uniform(b200_ctx* ctx, b200_stream s, b200_dtype dtype, b200_dptr out, uint64_t n, uint64_t seed, float lo, float hi);
int b200_fill_modulo(b200_ctx* ctx, b200_stream s, b200_dtype dtype, b200_dptr out, uint64_t n, uint32_t modulus);
/* compute_cmma_throughput as CubeCL would JIT it today (wmma 16x16x16): launches grid = SMs*32, block = 256, `n_iter`
 * dependent mma_sync per plane; *ops = cubes * planes * 2*m*n*k * n_iter (compute_cmma.rs:16,41-42). dtype F16 or BF16. */
int b200_probe_wmma(b200_ctx* ctx, b200_stream s, b200_dtype dtype, uint32_t n_iter, b200_dptr scratch_1k, double* ops);
/* The same accounting on the 5th-gen tensor cores: every CTA pair issues n_iter x 4 dependent UMMA 256x256x16 (bf16->f32,
 * TMEM accumulator) on smem-resident operands; *ops = pairs * n_iter * 4 * 2*256*256*16.  scratch: >= 4 * num_sms/2 bytes;
 * scratch[pair] = 64 * n_iter afterwards. */
int b200_probe_umma(b200_ctx* ctx, b200_stream s, uint32_t n_iter, b200_dptr scratch, double* ops);
/* The same probe for the other operand kinds: dtype B200_BF16 (as above), B200_F8E4M3 (kind::f8f6f4, or kind::mxf8f6f4
 * block-scaled when block_scaled != 0) or B200_F4E2M1X2 (kind::mxf4, block_scaled only).  UMMA 256x256xK with K = 16 / 32 /
 * 64 elements; *ops = pairs * n_iter * 4 * 2*256*256*K; scratch[pair] = 4 * K * n_iter afterwards. */
int b200_probe_umma_kind(b200_ctx* ctx, b200_stream s, b200_dtype dtype, int block_scaled, uint32_t n_iter, b200_dptr scratch,
                         double* ops);
/* memory_read_throughput with float_4 lines over `bytes` of `buf` (memory_read.rs:68-154): grid = SMs*32, block = 256. */
int b200_probe_memread(b200_ctx* ctx, b200_stream s, b200_dptr buf, uint64_t bytes, b200_dptr scratch_16);
/* memory_write_throughput (memory_write.rs: writes only) and memory_direct (memory_direct.rs: copy, both directions counted). */
int b200_probe_memwrite(b200_ctx* ctx, b200_stream s, b200_dptr dst, uint64_t bytes);
int b200_probe_memcopy(b200_ctx* ctx, b200_stream s, b200_dptr dst, b200_dptr src, uint64_t bytes);

/* Thread-local message of the last failing call on this thread ("" if none). */
const char* b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CUBECL_B200_H */
