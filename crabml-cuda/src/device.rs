//! `CudaTensorDevice`: the `DeviceRef` of the CUDA backend (cf. `CpuTensorDevice`, crabml-core/src/cpu/cpu_device.rs:50-107,
//! and `WgpuTensorDevice`).  Owns one `cc_device` = one GPU, one stream, one activation pool.
use std::ffi::CStr;
use std::ffi::CString;
use std::os::raw::c_int;
use std::ptr;
use std::sync::Arc;

use crabml::error::Error;
use crabml::error::ErrorKind;
use crabml::error::Result;
use crabml::tensor::TensorMetrics;

use crate::ffi;

#[derive(Debug, Clone)]
pub struct CudaTensorDeviceOptions {
    /// CUDA device index
    pub device_ordinal: i32,

    /// when enabled, whenever a tensor is called with `with_name`, the name and the tensor are snapshotted to the
    /// host and can be read back with `dump_debug_tensor`. only used in test (cpu_device.rs:14-16).
    pub debug_named_tensors: bool,

    /// 0: eager, one kernel launch per trait call.
    /// 1: trait calls are recorded, fused and replayed as a CUDA graph at the next `export`.
    /// 2: as 1, and a fused token runs as ONE persistent kernel (the default: fastest decode).
    pub lazy: i32,

    /// verification mode: every reduction in the scalar order of the CPU backend, bit-identical logits (slow).
    pub exact_order: bool,

    /// activation pool size hint in bytes, 0 = default
    pub pool_bytes: u64,

    pub metrics: TensorMetrics,
}

impl Default for CudaTensorDeviceOptions {
    fn default() -> Self {
        Self {
            device_ordinal: 0,
            debug_named_tensors: false,
            lazy: 2,
            exact_order: false,
            pool_bytes: 0,
            metrics: TensorMetrics::default(),
        }
    }
}

impl CudaTensorDeviceOptions {
    pub fn new() -> Self {
        Self::default()
    }

    pub fn with_device_ordinal(mut self, ordinal: i32) -> Self {
        self.device_ordinal = ordinal;
        self
    }

    pub fn with_debug_named_tensors(mut self, debug_named_tensors: bool) -> Self {
        self.debug_named_tensors = debug_named_tensors;
        self
    }

    pub fn with_lazy(mut self, lazy: i32) -> Self {
        self.lazy = lazy;
        self
    }

    pub fn with_exact_order(mut self, exact_order: bool) -> Self {
        self.exact_order = exact_order;
        self
    }

    pub fn with_pool_bytes(mut self, pool_bytes: u64) -> Self {
        self.pool_bytes = pool_bytes;
        self
    }

    pub fn with_metrics(mut self, metrics: TensorMetrics) -> Self {
        self.metrics = metrics;
        self
    }
}

pub struct CudaTensorDevice {
    pub(crate) raw: *mut ffi::cc_device,
    pub(crate) opts: CudaTensorDeviceOptions,
    pub(crate) metrics: TensorMetrics,
}

// The C library serialises access to a device's queues internally (one stream, a mutex around the activation pool) and
// makes the device current inside every entry point; the runner itself is single-threaded (llama2.rs).
unsafe impl Send for CudaTensorDevice {}
unsafe impl Sync for CudaTensorDevice {}

pub type CudaTensorDeviceRef = Arc<CudaTensorDevice>;

impl CudaTensorDevice {
    pub fn new(opts: CudaTensorDeviceOptions) -> Result<CudaTensorDeviceRef> {
        let c_opts = ffi::cc_device_options {
            device_ordinal: opts.device_ordinal,
            debug_named_tensors: opts.debug_named_tensors as i32,
            lazy: opts.lazy,
            exact_order: opts.exact_order as i32,
            pool_bytes: opts.pool_bytes,
        };
        let mut raw: *mut ffi::cc_device = ptr::null_mut();
        let rc = unsafe { ffi::cc_device_create(&c_opts, &mut raw) };
        if rc != ffi::CC_OK || raw.is_null() {
            // there is no CPU fallback: without a usable GPU the backend cannot be created
            return Err(Self::error_from(ptr::null_mut(), rc));
        }
        let metrics = opts.metrics.clone();
        Ok(Arc::new(Self { raw, opts, metrics }))
    }

    pub fn metrics(&self) -> &TensorMetrics {
        &self.metrics
    }

    pub fn options(&self) -> &CudaTensorDeviceOptions {
        &self.opts
    }

    /// wait until everything enqueued on the device's stream has finished (in lazy mode: flush first)
    pub fn synchronize(&self) -> Result<()> {
        self.check(unsafe { ffi::cc_device_synchronize(self.raw) })
    }

    /// lazy mode: run everything recorded so far, asynchronously; a no-op in eager mode
    pub fn flush(&self) -> Result<()> {
        self.check(unsafe { ffi::cc_device_flush(self.raw) })
    }

    /// kernels launched by the library since the device was created
    pub fn launch_count(&self) -> u64 {
        unsafe { ffi::cc_device_launch_count(self.raw) }
    }

    /// the snapshot `with_name(name)` took of a tensor (cpu_device.rs:96-98); None when the name was never tapped or
    /// `debug_named_tensors` is off
    pub fn dump_debug_tensor(&self, name: &str) -> Option<Vec<f32>> {
        let c_name = CString::new(name).ok()?;
        let mut n: usize = 0;
        let rc = unsafe { ffi::cc_dump_debug_tensor(self.raw, c_name.as_ptr(), ptr::null_mut(), &mut n) };
        if rc != ffi::CC_OK {
            return None;
        }
        let mut out = vec![0.0f32; n];
        let rc = unsafe { ffi::cc_dump_debug_tensor(self.raw, c_name.as_ptr(), out.as_mut_ptr(), &mut n) };
        if rc != ffi::CC_OK {
            return None;
        }
        out.truncate(n);
        Some(out)
    }

    /// status code of a C call -> `Result`: every failure is a `TensorError` carrying the library's message
    /// (error.rs:24-25,74-79: no numeric error code crosses the trait)
    pub(crate) fn check(&self, rc: c_int) -> Result<()> {
        if rc == ffi::CC_OK {
            return Ok(());
        }
        Err(Self::error_from(self.raw, rc))
    }

    fn error_from(raw: *mut ffi::cc_device, rc: c_int) -> Error {
        let msg_ptr = unsafe { ffi::cc_last_error(raw) };
        let text = if msg_ptr.is_null() {
            String::new()
        } else {
            unsafe { CStr::from_ptr(msg_ptr) }.to_string_lossy().into_owned()
        };
        let what = match rc {
            ffi::CC_ERR_TENSOR => "tensor error",
            ffi::CC_ERR_CUDA => "cuda error",
            ffi::CC_ERR_ARG => "bad argument",
            ffi::CC_ERR_UNSUPPORTED => "unsupported",
            _ => "error",
        };
        Error {
            kind: ErrorKind::TensorError,
            message: format!("crabml-cuda {} [{}]: {}", what, rc, text),
            cause: None,
        }
    }
}

impl Drop for CudaTensorDevice {
    fn drop(&mut self) {
        // every CudaTensor holds an Arc of its device, so all tensors are gone by now
        unsafe { ffi::cc_device_destroy(self.raw) };
    }
}
