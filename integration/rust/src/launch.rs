//! cubek-shaped launch surface over the safe layer: `matmul::launch` / `reduce::launch` taking tensor handles, the
//! std-lib op convention of the reference (`op::launch(client, &TensorHandle...)`, crates/cubecl-std/src/tensor/
//! identity.rs:39-83).  SOURCE ONLY (no Rust toolchain in the authoring image); the Python mirror of exactly this logic
//! (cubecl_b200/matmul.py, reduce.py) is what the GPU parity tests drive.
//!
//! Inside cubecl-cuda the `Context` is owned by `CudaServer` and `TensorHandle::ptr` is what `BufferBinding` resolves to
//! on the runner thread; standalone, it is a pointer from `Context::alloc`.

use crate::{b200_dptr, b200_stream, Context, DType, Error, Epilogue, ReduceOp, Status, TensorView};

/// Buffer + shape + strides (in ELEMENTS) + dtype: `TensorHandle<R>` of crates/cubecl-std/src/tensor/handle.rs:13-23
/// with the `Handle` already resolved to a device pointer.
#[derive(Clone, Debug)]
pub struct TensorHandle {
    pub ptr: b200_dptr,
    pub shape: Vec<u64>,
    pub strides: Vec<u64>,
    pub dtype: DType,
}

impl TensorHandle {
    /// Row-major contiguous strides for `shape` (`TensorHandle::new_contiguous`, handle.rs:62-75).
    pub fn new_contiguous(ptr: b200_dptr, shape: &[u64], dtype: DType) -> Self {
        let mut strides = vec![1u64; shape.len()];
        for i in (0..shape.len().saturating_sub(1)).rev() {
            strides[i] = strides[i + 1] * shape[i + 1];
        }
        Self { ptr, shape: shape.to_vec(), strides, dtype }
    }

    /// Swap the last two dims without moving data (`MatrixBatchLayout::MildlyPermuted { transposed: true, .. }`,
    /// crates/cubecl-std/src/tensor/matrix_batch_layout.rs:8-19): goes straight into a TMA descriptor, no copy.
    pub fn transposed(&self) -> Self {
        let r = self.shape.len();
        assert!(r >= 2);
        let mut t = self.clone();
        t.shape.swap(r - 2, r - 1);
        t.strides.swap(r - 2, r - 1);
        t
    }

    fn view(&self) -> TensorView<'_> {
        TensorView { ptr: self.ptr, shape: &self.shape, strides: &self.strides }
    }
}

fn invalid(message: String) -> Error {
    Error { status: Status::InvalidArg as i32, message }
}

pub mod matmul {
    use super::*;

    /// Batch-broadcast matmul shape rule, restated from crates/cubecl-zspace/src/shape.rs:489-517: equal rank >= 2;
    /// leading dims equal or one of them 1; inner dims agree.
    pub fn calculate_matmul_output(lhs: &[u64], rhs: &[u64]) -> Result<Vec<u64>, Error> {
        let rank = lhs.len();
        if rank != rhs.len() {
            return Err(invalid(format!("rank mismatch: lhs {}, rhs {}", rank, rhs.len())));
        }
        if rank < 2 {
            return Err(invalid("matmul needs rank >= 2".to_string()));
        }
        let mut out = Vec::with_capacity(rank);
        for i in 0..rank - 2 {
            let (l, r) = (lhs[i], rhs[i]);
            if l == r || r == 1 {
                out.push(l);
            } else if l == 1 {
                out.push(r);
            } else {
                return Err(invalid(format!("batch dims {} and {} cannot broadcast", l, r)));
            }
        }
        if lhs[rank - 1] != rhs[rank - 2] {
            return Err(invalid(format!("inner dims differ: lhs k={}, rhs k={}", lhs[rank - 1], rhs[rank - 2])));
        }
        out.push(lhs[rank - 2]);
        out.push(rhs[rank - 1]);
        Ok(out)
    }

    /// `matmul::launch(client, lhs, rhs, out)`: out[.., m, n] = sum_k lhs[.., m, k] * rhs[.., k, n], f32 accumulation.
    /// Shape errors are returned here; device faults surface at the next `Context::sync` like the reference's launches.
    ///
    /// # Safety
    /// The handles must describe live device allocations of `ctx`'s device that outlive the enqueued launch.
    pub unsafe fn launch(
        ctx: &mut Context, stream: b200_stream, lhs: &TensorHandle, rhs: &TensorHandle, out: &TensorHandle,
    ) -> Result<(), Error> {
        if lhs.dtype != rhs.dtype {
            return Err(invalid("lhs and rhs dtypes differ".to_string()));
        }
        let expect = calculate_matmul_output(&lhs.shape, &rhs.shape)?;
        if expect != out.shape {
            return Err(invalid(format!("out shape {:?} != {:?}", out.shape, expect)));
        }
        ctx.matmul(stream, lhs.dtype, out.dtype, &lhs.view(), &rhs.view(), &out.view())
    }

    /// out = act(alpha * (lhs @ rhs) + bias[n]) inside the GEMM epilogue.
    ///
    /// # Safety
    /// As [`launch`]; `epilogue.bias` must be 0 or an f32[N] device allocation.
    pub unsafe fn launch_fused(
        ctx: &mut Context, stream: b200_stream, lhs: &TensorHandle, rhs: &TensorHandle, out: &TensorHandle,
        epilogue: &Epilogue,
    ) -> Result<(), Error> {
        if lhs.dtype != rhs.dtype {
            return Err(invalid("lhs and rhs dtypes differ".to_string()));
        }
        let expect = calculate_matmul_output(&lhs.shape, &rhs.shape)?;
        if expect != out.shape {
            return Err(invalid(format!("out shape {:?} != {:?}", out.shape, expect)));
        }
        ctx.matmul_fused(stream, lhs.dtype, out.dtype, &lhs.view(), &rhs.view(), &out.view(), epilogue)
    }
}

pub mod reduce {
    use super::*;

    /// Shape of the result: the reduced axis removed, `[1]` for `axis == None` or a rank-1 input.
    pub fn output_shape(shape: &[u64], axis: Option<usize>) -> Result<Vec<u64>, Error> {
        match axis {
            None => Ok(vec![1]),
            Some(a) if a >= shape.len() => Err(invalid(format!("axis {} out of range for rank {}", a, shape.len()))),
            Some(a) => {
                let mut out: Vec<u64> = shape.iter().enumerate().filter(|(i, _)| *i != a).map(|(_, d)| *d).collect();
                if out.is_empty() {
                    out.push(1);
                }
                Ok(out)
            }
        }
    }

    /// U32 indices for the arg ops, F32 values otherwise.
    pub fn output_dtype(op: ReduceOp) -> DType {
        match op {
            ReduceOp::ArgMax | ReduceOp::ArgMin => DType::U32,
            _ => DType::F32,
        }
    }

    /// `reduce::launch(client, input, output, axis, op)`; `output` must be contiguous with [`output_shape`] and
    /// [`output_dtype`].  Arg ops: ties -> lowest index, the first NaN wins.
    ///
    /// # Safety
    /// As [`super::matmul::launch`].
    pub unsafe fn launch(
        ctx: &mut Context, stream: b200_stream, input: &TensorHandle, output: &TensorHandle, axis: Option<usize>,
        op: ReduceOp,
    ) -> Result<(), Error> {
        if output.dtype != output_dtype(op) {
            return Err(invalid("reduce: wrong output dtype for this op".to_string()));
        }
        if output.shape != output_shape(&input.shape, axis)? {
            return Err(invalid("reduce: wrong output shape".to_string()));
        }
        ctx.reduce(stream, op, input.dtype, &input.view(), output.ptr, axis)
    }
}
