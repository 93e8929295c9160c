//! `CudaTensor`: `impl Tensor` (crabml-core/src/tensor/api.rs:11-79) over the C ABI.
//!
//! A tensor is {refcounted device buffer, host-side `TensorStrider`, device Arc, optional name}.  Metadata-only methods
//! (`reshape`, `transpose`, `with_strider`, `resize`, `shape`, `strider`) never cross the ABI; every op passes a `cc_view`
//! = buffer handle + the strider's shape and strides.  Ops are enqueued on the device's stream; only `export` (and the
//! debug tap behind `with_name`) wait for the GPU -- the reference synchronises in the same places (llama2.rs:209).
use std::ffi::CString;
use std::ptr;
use std::ptr::NonNull;

use crabml::bail;
use crabml::error::ErrorKind;
use crabml::error::Result;
use crabml::gguf::GGMLType;
use crabml::tensor::RopeMode;
use crabml::tensor::Tensor;
use crabml::tensor::TensorStrider;

use crate::device::CudaTensorDeviceRef;
use crate::ffi;

pub struct CudaTensor {
    buf: NonNull<ffi::cc_buf>,
    dtype: GGMLType,
    strider: TensorStrider,
    device: CudaTensorDeviceRef,
    name: Option<String>,
}

// the buffer handle is an atomically refcounted object of the C library
unsafe impl Send for CudaTensor {}
unsafe impl Sync for CudaTensor {}

impl Clone for CudaTensor {
    /// shares the storage, like `Arc<wgpu::Buffer>` in the wgpu backend
    fn clone(&self) -> Self {
        unsafe { ffi::cc_tensor_retain(self.buf.as_ptr()) };
        Self {
            buf: self.buf,
            dtype: self.dtype,
            strider: self.strider.clone(),
            device: self.device.clone(),
            name: self.name.clone(),
        }
    }
}

impl Drop for CudaTensor {
    fn drop(&mut self) {
        unsafe { ffi::cc_tensor_release(self.buf.as_ptr()) };
    }
}

impl CudaTensor {
    /// adopts ONE reference of `raw`
    fn adopt(
        raw: *mut ffi::cc_buf,
        dtype: GGMLType,
        strider: TensorStrider,
        device: CudaTensorDeviceRef,
    ) -> Result<Self> {
        match NonNull::new(raw) {
            Some(buf) => Ok(Self {
                buf,
                dtype,
                strider,
                device,
                name: None,
            }),
            None => bail!(ErrorKind::TensorError, "crabml-cuda returned a null buffer"),
        }
    }

    /// host helper for tests and tools: an F32 tensor from a slice of floats
    pub fn new(src: &[f32], shape: &[usize], device: CudaTensorDeviceRef) -> Result<Self> {
        let bytes = unsafe { std::slice::from_raw_parts(src.as_ptr() as *const u8, std::mem::size_of_val(src)) };
        Self::from_cpu(bytes, shape, GGMLType::F32, device)
    }

    pub fn device(&self) -> &CudaTensorDeviceRef {
        &self.device
    }

    pub fn name(&self) -> Option<&str> {
        self.name.as_deref()
    }

    /// elements of the backing storage (the bound `resize` checks against, cpu_tensor.rs:180)
    pub fn capacity(&self) -> usize {
        unsafe { ffi::cc_tensor_capacity(self.buf.as_ptr()) as usize }
    }

    fn view(&self) -> Result<ffi::cc_view> {
        let dims = self.strider.dims();
        if dims == 0 || dims > ffi::CC_MAX_DIMS {
            bail!(
                ErrorKind::TensorError,
                "crabml-cuda supports 1..={} dimensions, got shape {:?}",
                ffi::CC_MAX_DIMS,
                self.strider.shape()
            );
        }
        let mut v = ffi::cc_view {
            buf: self.buf.as_ptr(),
            ndim: dims as i32,
            shape: [0; ffi::CC_MAX_DIMS],
            strides: [0; ffi::CC_MAX_DIMS],
        };
        for i in 0..dims {
            v.shape[i] = self.strider.shape()[i] as i64;
            v.strides[i] = self.strider.strides()[i] as i64;
        }
        Ok(v)
    }

    fn dims_i64(shape: &[usize]) -> Vec<i64> {
        shape.iter().map(|&d| d as i64).collect()
    }

    /// exchange step of the sharded decode path (not part of the reference's trait, which is single-device):
    /// self = sum over ranks of self, the same bits on every rank
    pub fn all_reduce_sum_inplace(self) -> Result<Self> {
        let v = self.view()?;
        self.device
            .check(unsafe { ffi::cc_all_reduce_sum_inplace(self.device.raw, &v) })?;
        Ok(self)
    }

    /// self[r * n .. (r + 1) * n] = `slice` of rank r
    pub fn all_gather_from(&mut self, slice: &Self) -> Result<()> {
        let (d, s) = (self.view()?, slice.view()?);
        self.device
            .check(unsafe { ffi::cc_all_gather(self.device.raw, &d, &s) })
    }
}

impl Tensor for CudaTensor {
    type DeviceRef = CudaTensorDeviceRef;

    /// copies GGUF-layout bytes (row-major rows of quant blocks, any GGMLType the CPU backend reads) to the device,
    /// where quantized blocks are repacked once into the coalesced plane layout.  The size is taken from
    /// shape x block size, not from `buf.len()`: `GGUFTensorInfo::data()` may run to the next tensor's offset
    /// (gguf.rs:742-747).
    fn from_cpu(buf: &[u8], shape: &[usize], dtype: GGMLType, device: Self::DeviceRef) -> Result<Self> {
        let dims = Self::dims_i64(shape);
        let mut out: *mut ffi::cc_buf = ptr::null_mut();
        device.check(unsafe {
            ffi::cc_tensor_from_cpu(
                device.raw,
                buf.as_ptr() as *const _,
                buf.len(),
                dims.as_ptr(),
                dims.len() as i32,
                dtype as i32,
                &mut out,
            )
        })?;
        Self::adopt(out, dtype, TensorStrider::new(shape.to_vec()), device)
    }

    /// F32 (zero-filled) or F16 storage for activations and kv caches (cpu_tensor.rs:138-165)
    fn alloc(shape: &[usize], dtype: GGMLType, device: Self::DeviceRef) -> Result<Self> {
        if dtype != GGMLType::F32 && dtype != GGMLType::F16 {
            bail!(ErrorKind::TensorError, "only f32/f16 is supported on alloc");
        }
        let dims = Self::dims_i64(shape);
        let mut out: *mut ffi::cc_buf = ptr::null_mut();
        device.check(unsafe {
            ffi::cc_tensor_alloc(device.raw, dims.as_ptr(), dims.len() as i32, dtype as i32, &mut out)
        })?;
        Self::adopt(out, dtype, TensorStrider::new(shape.to_vec()), device)
    }

    /// metadata only (cpu_tensor.rs:167-197): strides are kept, the new shape must fit the storage
    fn resize(mut self, axis: usize, n: usize) -> Result<Self> {
        if axis >= self.shape().len() {
            bail!(
                ErrorKind::TensorError,
                "resize: axis {} is larger than the current shape {:?}",
                axis,
                self.shape()
            );
        }
        let mut new_shape = self.shape().to_vec();
        new_shape[axis] = n;
        let new_len: usize = new_shape.iter().product();
        if new_len > self.capacity() {
            bail!(
                ErrorKind::TensorError,
                "resize: new shape {:?} is larger than the current shape {:?}",
                new_shape,
                self.shape()
            );
        }
        self.strider = self.strider.resize(&new_shape)?;
        self.name = None;
        Ok(self)
    }

    fn dtype(&self) -> GGMLType {
        self.dtype
    }

    fn with_strider(mut self, strider: TensorStrider) -> Result<Self> {
        self.strider = strider;
        Ok(self)
    }

    /// records the name; with `debug_named_tensors` the tensor is also snapshotted to the host
    /// (cpu_tensor.rs:232-241) -- in lazy mode this forces a flush, exactly what the cross-backend tap test
    /// (llama2.rs:768-784) needs
    fn with_name(mut self, name: String) -> Self {
        if self.device.opts.debug_named_tensors {
            if let (Ok(c_name), Ok(v)) = (CString::new(name.as_str()), self.view()) {
                // the trait gives with_name no way to fail: a failed tap leaves the name unset on the device side and
                // `dump_debug_tensor` returns None for it
                let _ = unsafe { ffi::cc_debug_tensor_tap(self.device.raw, c_name.as_ptr(), &v) };
            }
        }
        self.name = Some(name);
        self
    }

    fn reshape(mut self, shape: &[usize]) -> Result<Self> {
        self.strider = self.strider.reshape(shape.to_vec())?;
        Ok(self)
    }

    fn transpose(mut self, dims: &[usize]) -> Result<Self> {
        self.strider = self.strider.transpose(dims)?;
        Ok(self)
    }

    fn contiguous(self) -> Result<Self> {
        let v = self.view()?;
        let mut out: *mut ffi::cc_buf = ptr::null_mut();
        self.device
            .check(unsafe { ffi::cc_contiguous(self.device.raw, &v, &mut out) })?;
        // an already contiguous tensor comes back as the same (retained) storage
        let shape = self.shape().to_vec();
        Self::adopt(out, self.dtype, TensorStrider::new(shape), self.device.clone())
    }

    fn shape(&self) -> &[usize] {
        self.strider.shape()
    }

    fn strider(&self) -> &TensorStrider {
        &self.strider
    }

    /// writes `rhs` behind the current extent along `axis` using self's (pre-allocated) strides, then grows the
    /// shape (cpu_tensor.rs:251-292, concatenate.rs:12-77): the kv-cache append
    fn concatenate(&mut self, rhs: &Self, axis: usize) -> Result<()> {
        if axis >= self.shape().len() || rhs.shape().len() != self.shape().len() {
            bail!(
                ErrorKind::TensorError,
                "shape mismatch on concatenate, want {:?} but got {:?}",
                self.shape(),
                rhs.shape()
            );
        }
        let (a, r) = (self.view()?, rhs.view()?);
        self.device
            .check(unsafe { ffi::cc_concatenate(self.device.raw, &a, &r, axis as i32) })?;
        let mut new_shape = self.shape().to_vec();
        new_shape[axis] += rhs.shape()[axis];
        self.strider = self.strider.resize(&new_shape)?;
        Ok(())
    }

    /// embedding lookup: rows of a (possibly quantized) 2-d tensor are dequantized into self
    fn copy_rows_from(&mut self, rhs: &Self, rows: &[usize]) -> Result<()> {
        let (d, s) = (self.view()?, rhs.view()?);
        let rows_i64: Vec<i64> = rows.iter().map(|&r| r as i64).collect();
        self.device.check(unsafe {
            ffi::cc_copy_rows_from(self.device.raw, &d, &s, rows_i64.as_ptr(), rows_i64.len() as i32)
        })
    }

    /// the one synchronising call of a forward pass (llama2.rs:209)
    fn export(&self, buf: &mut [f32]) -> Result<()> {
        let v = self.view()?;
        self.device
            .check(unsafe { ffi::cc_tensor_export_f32(self.device.raw, &v, buf.as_mut_ptr(), buf.len()) })
    }

    fn dup(&self) -> Result<Self> {
        let v = self.view()?;
        let mut out: *mut ffi::cc_buf = ptr::null_mut();
        self.device
            .check(unsafe { ffi::cc_tensor_dup(self.device.raw, &v, &mut out) })?;
        Self::adopt(
            out,
            self.dtype,
            TensorStrider::new(self.shape().to_vec()),
            self.device.clone(),
        )
    }

    fn rope_inplace(self, mode: RopeMode, pos: usize, rope_dims: usize) -> Result<Self> {
        let c_mode = match mode {
            RopeMode::Llama => ffi::CC_ROPE_LLAMA,
            RopeMode::Neox => ffi::CC_ROPE_NEOX,
        };
        let v = self.view()?;
        self.device.check(unsafe {
            ffi::cc_rope_inplace(self.device.raw, &v, c_mode, pos as i64, rope_dims as i64)
        })?;
        Ok(self)
    }

    fn rms_norm_inplace(self, eps: f32) -> Result<Self> {
        let v = self.view()?;
        self.device
            .check(unsafe { ffi::cc_rms_norm_inplace(self.device.raw, &v, eps) })?;
        Ok(self)
    }

    fn softmax_inplace(self, axis: usize) -> Result<Self> {
        let v = self.view()?;
        self.device
            .check(unsafe { ffi::cc_softmax_inplace(self.device.raw, &v, axis as i32) })?;
        Ok(self)
    }

    fn silu_inplace(self) -> Result<Self> {
        let v = self.view()?;
        self.device
            .check(unsafe { ffi::cc_silu_inplace(self.device.raw, &v) })?;
        Ok(self)
    }

    fn gelu_inplace(self) -> Result<Self> {
        let v = self.view()?;
        self.device
            .check(unsafe { ffi::cc_gelu_inplace(self.device.raw, &v) })?;
        Ok(self)
    }

    fn mul_inplace(self, rhs: &Self) -> Result<Self> {
        let (v, r) = (self.view()?, rhs.view()?);
        self.device
            .check(unsafe { ffi::cc_mul_inplace(self.device.raw, &v, &r) })?;
        Ok(self)
    }

    fn add_inplace(self, rhs: &Self) -> Result<Self> {
        let (v, r) = (self.view()?, rhs.view()?);
        self.device
            .check(unsafe { ffi::cc_add_inplace(self.device.raw, &v, &r) })?;
        Ok(self)
    }

    fn scale_inplace(self, rhs: f32) -> Result<Self> {
        let v = self.view()?;
        self.device
            .check(unsafe { ffi::cc_scale_inplace(self.device.raw, &v, rhs) })?;
        Ok(self)
    }

    /// W (m, k) in any GGML block type times an F32 activation (k) or (b, k): the activation is quantized on the fly to
    /// W's partner type (buf/api.rs:142-159) and every output row is one vec_dot -- the decode hot path
    fn matmul_vec(&self, y: &Self) -> Result<Self> {
        let (w, x) = (self.view()?, y.view()?);
        let mut out: *mut ffi::cc_buf = ptr::null_mut();
        self.device
            .check(unsafe { ffi::cc_matmul_vec(self.device.raw, &w, &x, &mut out) })?;
        let shape = if y.shape().len() == 1 {
            vec![self.shape()[0]]
        } else {
            vec![y.shape()[0], self.shape()[0]]
        };
        Self::adopt(out, GGMLType::F32, TensorStrider::new(shape), self.device.clone())
    }

    /// (b, m, k) x (b', k, n) -> (b, m, n); the rhs may be a strided F32 or F16 kv cache with fewer (grouped) heads
    fn batch_matmul(&self, y: &Self) -> Result<Self> {
        let (a, b) = (self.view()?, y.view()?);
        let mut out: *mut ffi::cc_buf = ptr::null_mut();
        self.device
            .check(unsafe { ffi::cc_batch_matmul(self.device.raw, &a, &b, &mut out) })?;
        let shape = vec![self.shape()[0], self.shape()[1], y.shape()[2]];
        Self::adopt(out, GGMLType::F32, TensorStrider::new(shape), self.device.clone())
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use crate::device::CudaTensorDevice;
    use crate::device::CudaTensorDeviceOptions;

    fn device() -> CudaTensorDeviceRef {
        CudaTensorDevice::new(CudaTensorDeviceOptions::new().with_lazy(0)).unwrap()
    }

    // the known answers of the CPU and wgpu backends' own tests (cpu_tensor.rs:530-541, wgpu_tensor.rs:880-895)
    #[test]
    fn test_matmul_vec_f32() -> Result<()> {
        let d = device();
        let w = CudaTensor::new(&[1.0, 2.0, 3.0, 4.0, 5.0, 6.0], &[2, 3], d.clone())?;
        let b = CudaTensor::new(&[1.0, 2.0, 3.0], &[3], d.clone())?;
        let out = w.matmul_vec(&b)?;
        let mut dst = vec![0.0f32; 2];
        out.export(&mut dst)?;
        assert_eq!(dst, vec![14.0, 32.0]);
        Ok(())
    }

    #[test]
    fn test_resize_and_concatenate() -> Result<()> {
        let d = device();
        let mut cache = CudaTensor::alloc(&[2, 4, 3], GGMLType::F32, d.clone())?.resize(1, 0)?;
        let row = CudaTensor::new(&[1.0, 2.0, 3.0, 4.0, 5.0, 6.0], &[2, 1, 3], d.clone())?;
        cache.concatenate(&row, 1)?;
        cache.concatenate(&row, 1)?;
        assert_eq!(cache.shape(), &[2, 2, 3]);
        let dense = cache.contiguous()?;
        let mut dst = vec![0.0f32; 12];
        dense.export(&mut dst)?;
        assert_eq!(dst, vec![1.0, 2.0, 3.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 4.0, 5.0, 6.0]);
        Ok(())
    }

    #[test]
    fn test_softmax_and_silu_match_lut_numerics() -> Result<()> {
        let d = device();
        let t = CudaTensor::new(&[1.0, 2.0, 3.0, 4.0, 5.0, 6.0], &[2, 3], d.clone())?;
        let t = t.softmax_inplace(1)?;
        let mut dst = vec![0.0f32; 6];
        t.export(&mut dst)?;
        // cpu_tensor.rs:544-555
        approx::assert_relative_eq!(dst[0], 0.09003057, epsilon = 1e-3);
        approx::assert_relative_eq!(dst[2], 0.66524094, epsilon = 1e-3);
        Ok(())
    }
}
