//! Raw bindings + a thin safe layer over `include/cubecl_b200.h` (ABI version 1).
//!
//! SOURCE ONLY -- never compiled in the authoring image (no Rust toolchain).  Signatures are kept in lock-step with the
//! header by hand; `tests/test_abi.py` checks the header against the built library and the Python ctypes table.
//!
//! Intended use inside cubecl-cuda: `CudaServer` resolves `BufferBinding`s to `CUdeviceptr`s on the runner thread and
//! passes them here together with the `CUstream` of the current `StreamId` (crates/cubecl-cuda/src/compute/server.rs:
//! 1024-1144); errors are queued on the stream like any launch error (server.rs:269-284).
#![allow(non_camel_case_types)]

use core::ffi::{c_char, c_int, c_void, CStr};

#[repr(C)]
pub struct b200_ctx {
    _private: [u8; 0],
}
pub type b200_dptr = u64; // CUdeviceptr
pub type b200_stream = *mut c_void; // CUstream, null = the context's own compute stream

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

#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum DType {
    F32 = 0,
    F16 = 1,
    BF16 = 2,
    U32 = 3,
}

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

extern "C" {
    pub fn b200_abi_version() -> c_int;
    pub fn b200_init(device: c_int, out: *mut *mut b200_ctx) -> c_int;
    pub fn b200_destroy(ctx: *mut b200_ctx) -> c_int;
    pub fn b200_sync(ctx: *mut b200_ctx, s: b200_stream) -> c_int;
    pub fn b200_matmul(
        ctx: *mut b200_ctx, s: b200_stream, in_dtype: c_int, out_dtype: c_int,
        lhs: b200_dptr, rhs: b200_dptr, out: b200_dptr, rank: c_int,
        shape_lhs: *const u64, strides_lhs: *const u64,
        shape_rhs: *const u64, strides_rhs: *const u64,
        shape_out: *const u64, strides_out: *const u64,
    ) -> c_int;
    /// Block-scaled (MX) matmul: lhs [batch, m, k], rhs [batch, n, k] K-contiguous (e4m3 / e5m2, or both packed e2m1),
    /// ue8m0 scales [batch, rows, k / 32]; replaces `MmaDefinition::new_scaled` / `execute_scaled` tiles
    /// (crates/cubecl-core/src/frontend/cmma.rs:438-460, 798-840) at GEMM level.
    pub fn b200_matmul_scaled(
        ctx: *mut b200_ctx, s: b200_stream, lhs_dtype: c_int, rhs_dtype: c_int, out_dtype: c_int,
        lhs: b200_dptr, rhs: b200_dptr, lhs_scales: b200_dptr, rhs_scales: b200_dptr, out: b200_dptr,
        batch: u64, m: u64, n: u64, k: u64, scale_block: c_int, scales_packed: c_int,
    ) -> c_int;
    pub fn b200_reduce_strided(
        ctx: *mut b200_ctx, s: b200_stream, op: c_int, in_dtype: c_int, input: b200_dptr, out: b200_dptr,
        rank: c_int, shape: *const u64, strides: *const u64, axis: c_int,
    ) -> c_int;
    pub fn b200_reduce_all_reduce(
        ctx: *mut b200_ctx, s: b200_stream, op: c_int, in_dtype: c_int, input: b200_dptr, out: b200_dptr,
        n: u64, device_ids: *const c_int, ndev: c_int,
    ) -> c_int;
    pub fn b200_last_error() -> *const c_char;
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
    let message = unsafe { CStr::from_ptr(b200_last_error()) }.to_string_lossy().into_owned();
    Err(Error { status: rc, message })
}

/// One context per device; `Send` but not `Sync`, like `CudaServer` (one runner thread per device).
pub struct Context(*mut b200_ctx);
unsafe impl Send for Context {}

/// Strided tensor view: device pointer + shape + strides in ELEMENTS (TensorHandle, cubecl-std/src/tensor/handle.rs:13-23).
pub struct TensorView<'a> {
    pub ptr: b200_dptr,
    pub shape: &'a [u64],
    pub strides: &'a [u64],
}

impl Context {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut p = core::ptr::null_mut();
        check(unsafe { b200_init(device, &mut p) })?;
        Ok(Self(p))
    }

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
        check(b200_matmul(
            self.0, stream, in_dtype as c_int, out_dtype as c_int, lhs.ptr, rhs.ptr, out.ptr, rank as c_int,
            lhs.shape.as_ptr(), lhs.strides.as_ptr(), rhs.shape.as_ptr(), rhs.strides.as_ptr(),
            out.shape.as_ptr(), out.strides.as_ptr(),
        ))
    }

    /// Reduce `axis` (None = every element); output contiguous f32 (u32 indices for arg ops).
    ///
    /// # Safety
    /// Same contract as [`Context::matmul`].
    pub unsafe fn reduce(
        &mut self, stream: b200_stream, op: ReduceOp, in_dtype: DType, input: &TensorView, out: b200_dptr,
        axis: Option<usize>,
    ) -> Result<(), Error> {
        check(b200_reduce_strided(
            self.0, stream, op as c_int, in_dtype as c_int, input.ptr, out, input.shape.len() as c_int,
            input.shape.as_ptr(), input.strides.as_ptr(), axis.map(|a| a as c_int).unwrap_or(-1),
        ))
    }

    pub fn sync(&mut self, stream: b200_stream) -> Result<(), Error> {
        check(unsafe { b200_sync(self.0, stream) })
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { b200_destroy(self.0) };
    }
}
