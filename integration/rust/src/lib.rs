//! `cubecl-b200-sys`: raw bindings (`sys`, generated from `include/cubecl_b200.h`, ABI version 1) plus a thin safe layer.
//!
//! SOURCE ONLY -- never compiled in the authoring image (no Rust toolchain).  `sys.rs` is GENERATED from the header by
//! `tools/gen_rust_sys.py` and the CPU test `tests/test_rust_bindings.py` fails when header and bindings drift apart; the
//! safe layer below is hand-written against those signatures and the same test checks that every `sys::` function it
//! calls exists with the argument count used here.
//!
//! Intended use inside cubecl-cuda: `CudaServer` resolves `BufferBinding`s to `CUdeviceptr`s on the runner thread and
//! passes them here together with the `CUstream` of the current `StreamId` (crates/cubecl-cuda/src/compute/server.rs:
//! 1024-1144); errors are queued on the stream like any launch error (server.rs:269-284).  See INTEGRATION.md.

pub mod launch;
pub mod sys;

use core::ffi::{c_int, c_void, CStr};
use std::ffi::CString;

pub use sys::{b200_dptr, b200_event, b200_stream};

/// `b200_status`, mirroring LaunchError / ServerError (crates/cubecl-runtime/src/server/base.rs:177-272).
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Status {
    Ok = 0,
    Compilation = 1,
    OutOfMemory = 2,
    TooManyResources = 3,
    Unknown = 4,
    Io = 5,
    InvalidArg = 6,
    Unsupported = 7,
    NoDevice = 8,
    Comm = 9,
    Unhealthy = 10,
}

/// `b200_dtype`.
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum DType {
    F32 = 0,
    F16 = 1,
    BF16 = 2,
    U32 = 3,
    I32 = 4,
    F64 = 5,
    I64 = 6,
    U64 = 7,
    U8 = 8,
    I8 = 9,
    F8E4M3 = 10,
    F8E5M2 = 11,
    F4E2M1x2 = 12,
    UE8M0 = 13,
}

/// `b200_reduce_op`.
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum ReduceOp {
    Sum = 0,
    Prod = 1,
    Max = 2,
    Min = 3,
    ArgMax = 4,
    ArgMin = 5,
    Mean = 6,
}

/// `b200_comm_op` = ReduceOperation{Sum,Mean} (server/base.rs:623-628).
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum CommOp {
    Sum = 0,
    Mean = 1,
}

/// Activation of the fused GEMM epilogue (`b200_epilogue.activation`).
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Activation {
    None = 0,
    Relu = 1,
    Gelu = 2,
}

/// Error carrying the status and the thread-local message of the failing call.
#[derive(Debug)]
pub struct Error {
    pub status: i32,
    pub message: String,
}

fn check(rc: c_int) -> Result<(), Error> {
    if rc == 0 {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(sys::b200_last_error()) }.to_string_lossy().into_owned();
    Err(Error { status: rc, message })
}

/// One context per device; `Send` but not `Sync`, like `CudaServer` (one runner thread per device).
pub struct Context(*mut sys::b200_ctx);
unsafe impl Send for Context {}

/// Strided tensor view: device pointer + shape + strides in ELEMENTS (TensorHandle, cubecl-std/src/tensor/handle.rs:13-23).
pub struct TensorView<'a> {
    pub ptr: b200_dptr,
    pub shape: &'a [u64],
    pub strides: &'a [u64],
}

/// out = act(alpha * acc + bias[n]) inside the GEMM epilogue; `bias` is an f32[N] device pointer or 0.
pub struct Epilogue {
    pub alpha: f32,
    pub activation: Activation,
    pub bias: b200_dptr,
}

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut p = core::ptr::null_mut();
        check(unsafe { sys::b200_init(device, &mut p) })?;
        Ok(Self(p))
    }

    pub fn raw(&self) -> *mut sys::b200_ctx {
        self.0
    }

    pub fn properties(&self) -> Result<sys::b200_props, Error> {
        let mut props = core::mem::MaybeUninit::<sys::b200_props>::zeroed();
        check(unsafe { sys::b200_get_props(self.0, props.as_mut_ptr()) })?;
        Ok(unsafe { props.assume_init() })
    }

    /// String-typed runtime knob ("gemm.variant", "gemm.f32", "reduce.variant", ...; see the header).
    pub fn set_option(&mut self, key: &str, value: &str) -> Result<(), Error> {
        let k = CString::new(key).map_err(|_| Error { status: Status::InvalidArg as i32, message: "NUL in key".into() })?;
        let v = CString::new(value).map_err(|_| Error { status: Status::InvalidArg as i32, message: "NUL in value".into() })?;
        check(unsafe { sys::b200_set_option(self.0, k.as_ptr(), v.as_ptr()) })
    }

    /// Entry-point name of the kernel this context launched most recently (what a harness reports as the kernel it timed).
    pub fn last_kernel(&self) -> Result<String, Error> {
        let mut buf = [0u8; 256];
        check(unsafe { sys::b200_last_kernel(self.0, buf.as_mut_ptr() as *mut core::ffi::c_char, buf.len()) })?;
        let end = buf.iter().position(|&b| b == 0).unwrap_or(buf.len());
        Ok(String::from_utf8_lossy(&buf[..end]).into_owned())
    }

    pub fn launch_count(&self) -> Result<u64, Error> {
        let mut n = 0u64;
        check(unsafe { sys::b200_launch_count(self.0, &mut n) })?;
        Ok(n)
    }

    // ---- memory (standalone use; inside CubeCL the retained pool owns the buffers) --------------------------------
    pub fn alloc(&mut self, bytes: usize) -> Result<b200_dptr, Error> {
        let mut p: b200_dptr = 0;
        check(unsafe { sys::b200_alloc(self.0, bytes, &mut p) })?;
        Ok(p)
    }

    /// # Safety
    /// `ptr` must come from [`Context::alloc`] on this context and must not be used by work enqueued after this call.
    pub unsafe fn free(&mut self, ptr: b200_dptr) -> Result<(), Error> {
        check(sys::b200_free(self.0, ptr))
    }

    /// Frees in the order of `last_use`, the stream whose queued work may still touch the buffer (null = the context's
    /// stream): the pool re-issues the page only once an event recorded there has completed.
    ///
    /// # Safety
    /// Same contract as [`Context::free`].
    pub unsafe fn free_async(&mut self, ptr: b200_dptr, last_use: b200_stream) -> Result<(), Error> {
        check(sys::b200_free_async(self.0, ptr, last_use))
    }

    /// Stream-ordered host -> device copy; asynchronous when `src` is pinned (sync before reusing it).
    ///
    /// # Safety
    /// `dst` must be a live device allocation of at least `src.len()` bytes.
    pub unsafe fn write(&mut self, stream: b200_stream, dst: b200_dptr, src: &[u8]) -> Result<(), Error> {
        check(sys::b200_write(self.0, stream, dst, src.as_ptr() as *const c_void, src.len()))
    }

    /// Stream-ordered device -> host copy; call [`Context::sync`] before reading `dst`.
    ///
    /// # Safety
    /// `src` must be a live device allocation of at least `dst.len()` bytes.
    pub unsafe fn read(&mut self, stream: b200_stream, dst: &mut [u8], src: b200_dptr) -> Result<(), Error> {
        check(sys::b200_read(self.0, stream, dst.as_mut_ptr() as *mut c_void, src, dst.len()))
    }

    // ---- streams / events ------------------------------------------------------------------------------------------
    pub fn stream_create(&mut self) -> Result<b200_stream, Error> {
        let mut s: b200_stream = core::ptr::null_mut();
        check(unsafe { sys::b200_stream_create(self.0, &mut s) })?;
        Ok(s)
    }

    /// # Safety
    /// `stream` must come from [`Context::stream_create`] and must not be used afterwards.
    pub unsafe fn stream_destroy(&mut self, stream: b200_stream) -> Result<(), Error> {
        check(sys::b200_stream_destroy(self.0, stream))
    }

    /// Waits for `stream` (null = the context's compute stream); deferred device faults surface here as `Unhealthy`.
    pub fn sync(&mut self, stream: b200_stream) -> Result<(), Error> {
        check(unsafe { sys::b200_sync(self.0, stream) })
    }

    pub fn event_create(&mut self) -> Result<b200_event, Error> {
        let mut e: b200_event = core::ptr::null_mut();
        check(unsafe { sys::b200_event_create(self.0, &mut e) })?;
        Ok(e)
    }

    /// # Safety
    /// `event` / `stream` must be live handles of this context.
    pub unsafe fn event_record(&mut self, event: b200_event, stream: b200_stream) -> Result<(), Error> {
        check(sys::b200_event_record(self.0, event, stream))
    }

    /// # Safety
    /// `event` / `stream` must be live handles of this context.
    pub unsafe fn stream_wait_event(&mut self, stream: b200_stream, event: b200_event) -> Result<(), Error> {
        check(sys::b200_stream_wait_event(self.0, stream, event))
    }

    /// # Safety
    /// Both events must be live, recorded handles of this context.
    pub unsafe fn event_elapsed_ms(&mut self, start: b200_event, end: b200_event) -> Result<f32, Error> {
        let mut ms = 0f32;
        check(sys::b200_event_elapsed_ms(self.0, start, end, &mut ms))?;
        Ok(ms)
    }

    /// # Safety
    /// `event` must be a live handle of this context and must not be used afterwards.
    pub unsafe fn event_destroy(&mut self, event: b200_event) -> Result<(), Error> {
        check(sys::b200_event_destroy(self.0, event))
    }

    // ---- matmul::launch ---------------------------------------------------------------------------------------------
    /// out = lhs @ rhs, f32 accumulate, batch broadcast (shape.rs:489-517), enqueued on `stream`.
    ///
    /// # Safety
    /// The pointers must be live device allocations of this device large enough for the described views, and must stay
    /// alive until the stream has passed this launch (the pool's handle ref-count guarantees that inside CubeCL).
    pub unsafe fn matmul(
        &mut self, stream: b200_stream, in_dtype: DType, out_dtype: DType,
        lhs: &TensorView, rhs: &TensorView, out: &TensorView,
    ) -> Result<(), Error> {
        let rank = lhs.shape.len();
        assert!(rhs.shape.len() == rank && out.shape.len() == rank);
        assert!(lhs.strides.len() == rank && rhs.strides.len() == rank && out.strides.len() == rank);
        check(sys::b200_matmul(
            self.0, stream, in_dtype as c_int, out_dtype as c_int, lhs.ptr, rhs.ptr, out.ptr, rank as c_int,
            lhs.shape.as_ptr(), lhs.strides.as_ptr(), rhs.shape.as_ptr(), rhs.strides.as_ptr(),
            out.shape.as_ptr(), out.strides.as_ptr(),
        ))
    }

    /// The same product with different 8-bit formats per operand (i8 x u8, e4m3 x e5m2 ...: the reference's manual-MMA
    /// pairs, crates/cubecl-cpp/src/cuda/mma/manual.rs:151-186).
    ///
    /// # Safety
    /// Same contract as [`Context::matmul`].
    pub unsafe fn matmul_mixed(
        &mut self, stream: b200_stream, lhs_dtype: DType, rhs_dtype: DType, out_dtype: DType,
        lhs: &TensorView, rhs: &TensorView, out: &TensorView,
    ) -> Result<(), Error> {
        let rank = lhs.shape.len();
        assert!(rhs.shape.len() == rank && out.shape.len() == rank);
        assert!(lhs.strides.len() == rank && rhs.strides.len() == rank && out.strides.len() == rank);
        check(sys::b200_matmul_mixed(
            self.0, stream, lhs_dtype as c_int, rhs_dtype as c_int, out_dtype as c_int, lhs.ptr, rhs.ptr, out.ptr,
            rank as c_int, lhs.shape.as_ptr(), lhs.strides.as_ptr(), rhs.shape.as_ptr(), rhs.strides.as_ptr(),
            out.shape.as_ptr(), out.strides.as_ptr(),
        ))
    }

    /// out = act(alpha * (lhs @ rhs) + bias[n]) in one launch.
    ///
    /// # Safety
    /// Same contract as [`Context::matmul`]; `epilogue.bias` must be 0 or an f32[N] device allocation.
    pub unsafe fn matmul_fused(
        &mut self, stream: b200_stream, in_dtype: DType, out_dtype: DType,
        lhs: &TensorView, rhs: &TensorView, out: &TensorView, epilogue: &Epilogue,
    ) -> Result<(), Error> {
        let rank = lhs.shape.len();
        assert!(rhs.shape.len() == rank && out.shape.len() == rank);
        assert!(lhs.strides.len() == rank && rhs.strides.len() == rank && out.strides.len() == rank);
        let e = sys::b200_epilogue { alpha: epilogue.alpha, activation: epilogue.activation as i32, bias: epilogue.bias };
        check(sys::b200_matmul_fused(
            self.0, stream, in_dtype as c_int, out_dtype as c_int, lhs.ptr, rhs.ptr, out.ptr, rank as c_int,
            lhs.shape.as_ptr(), lhs.strides.as_ptr(), rhs.shape.as_ptr(), rhs.strides.as_ptr(),
            out.shape.as_ptr(), out.strides.as_ptr(), &e,
        ))
    }

    /// Block-scaled (MX / NVFP4) matmul: lhs [batch, m, k], rhs [batch, n, k] K-contiguous, scales per `scale_block`
    /// (32: ue8m0, 16: e4m3) elements of K; replaces `MmaDefinition::new_scaled` / `execute_scaled` tiles
    /// (crates/cubecl-core/src/frontend/cmma.rs:438-460, 798-840) at GEMM level.
    ///
    /// # Safety
    /// Same contract as [`Context::matmul`] for all five pointers.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn matmul_scaled(
        &mut self, stream: b200_stream, lhs_dtype: DType, rhs_dtype: DType, out_dtype: DType,
        lhs: b200_dptr, rhs: b200_dptr, lhs_scales: b200_dptr, rhs_scales: b200_dptr, out: b200_dptr,
        batch: u64, m: u64, n: u64, k: u64, scale_block: i32, scales_packed: bool,
    ) -> Result<(), Error> {
        check(sys::b200_matmul_scaled(
            self.0, stream, lhs_dtype as c_int, rhs_dtype as c_int, out_dtype as c_int, lhs, rhs, lhs_scales, rhs_scales,
            out, batch, m, n, k, scale_block as c_int, scales_packed as c_int,
        ))
    }

    // ---- reduce::launch ---------------------------------------------------------------------------------------------
    /// Reduce `axis` (None = every element); output contiguous f32 (u32 indices for arg ops).
    ///
    /// # Safety
    /// Same contract as [`Context::matmul`].
    pub unsafe fn reduce(
        &mut self, stream: b200_stream, op: ReduceOp, in_dtype: DType, input: &TensorView, out: b200_dptr,
        axis: Option<usize>,
    ) -> Result<(), Error> {
        assert!(input.strides.len() == input.shape.len());
        check(sys::b200_reduce_strided(
            self.0, stream, op as c_int, in_dtype as c_int, input.ptr, out, input.shape.len() as c_int,
            input.shape.as_ptr(), input.strides.as_ptr(), axis.map(|a| a as c_int).unwrap_or(-1),
        ))
    }

    /// out (compact row-major) = gather of the strided tensor `input` (into_contiguous, cubecl-std/src/tensor/contiguous.rs).
    ///
    /// # Safety
    /// Same contract as [`Context::matmul`].
    pub unsafe fn into_contiguous(
        &mut self, stream: b200_stream, dtype: DType, input: &TensorView, out: b200_dptr,
    ) -> Result<(), Error> {
        assert!(input.strides.len() == input.shape.len());
        check(sys::b200_into_contiguous(
            self.0, stream, dtype as c_int, input.ptr, out, input.shape.len() as c_int, input.shape.as_ptr(),
            input.strides.as_ptr(),
        ))
    }

    // ---- collectives (ServerCommunication, server/base.rs:632-739) ---------------------------------------------------
    pub fn comm_unique_id(&mut self) -> Result<[u8; sys::B200_UNIQUE_ID_BYTES], Error> {
        let mut id = [0u8; sys::B200_UNIQUE_ID_BYTES];
        check(unsafe { sys::b200_comm_get_unique_id(self.0, id.as_mut_ptr() as *mut c_void) })?;
        Ok(id)
    }

    pub fn comm_init(&mut self, device_ids: &[i32], id: &[u8; sys::B200_UNIQUE_ID_BYTES]) -> Result<(), Error> {
        check(unsafe {
            sys::b200_comm_init(self.0, device_ids.as_ptr(), device_ids.len() as c_int, id.as_ptr() as *const c_void)
        })
    }

    /// In-place or out-of-place all-reduce of a whole buffer on the comm stream, ordered after `compute`.
    ///
    /// # Safety
    /// `src` / `dst` must be live device allocations of at least `bytes` bytes.
    pub unsafe fn all_reduce(
        &mut self, compute: b200_stream, src: b200_dptr, dst: b200_dptr, bytes: usize, dtype: DType, op: CommOp,
        device_ids: &[i32],
    ) -> Result<(), Error> {
        check(sys::b200_all_reduce(
            self.0, compute, src, dst, bytes, dtype as c_int, op as c_int, device_ids.as_ptr(), device_ids.len() as c_int,
        ))
    }

    /// Make `compute` wait for everything issued on the comm stream (server.rs:782-798).
    pub fn sync_collective(&mut self, compute: b200_stream) -> Result<(), Error> {
        check(unsafe { sys::b200_sync_collective(self.0, compute) })
    }

    /// Local f32 sum + exchange of the scalar through NVLink peer mailboxes in ONE kernel (after `b200_p2p_connect`).
    ///
    /// # Safety
    /// Same contract as [`Context::matmul`]; every rank of `device_ids` must make the same call.
    pub unsafe fn reduce_all_reduce(
        &mut self, stream: b200_stream, input: b200_dptr, out: b200_dptr, n: u64, device_ids: &[i32],
    ) -> Result<(), Error> {
        check(sys::b200_reduce_all_reduce(
            self.0, stream, ReduceOp::Sum as c_int, DType::F32 as c_int, input, out, n, device_ids.as_ptr(),
            device_ids.len() as c_int,
        ))
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { sys::b200_destroy(self.0) };
    }
}

/// The embedded prebuilt sm_100a image `name` ("gemm" | "gemm_mx" | "reduce" | "aux") for a host that prefers to
/// `cuModuleLoadData` it into its own module cache (CudaContext::modules, cubecl-cuda/src/compute/context.rs:38-62,293).
pub fn cubin(name: &str) -> Result<&'static [u8], Error> {
    let n = CString::new(name).map_err(|_| Error { status: Status::InvalidArg as i32, message: "NUL in name".into() })?;
    let mut image: *const c_void = core::ptr::null();
    let mut size = 0usize;
    check(unsafe { sys::b200_get_cubin(n.as_ptr(), &mut image, &mut size) })?;
    Ok(unsafe { core::slice::from_raw_parts(image as *const u8, size) })
}
