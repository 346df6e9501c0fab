//! sol::Context for this path (src/context.rs:239-369 creates instance / device / queues; here: one CUDA device + stream).
use crate::ffi::*;
use std::ffi::CStr;
use std::os::raw::c_void;
use std::sync::Arc;

pub struct Context {
    pub(crate) raw: *mut solb_ctx,
}

// One ctx = one device = one stream; like the reference's Context it is shared through Arc but driven from one thread.
unsafe impl Send for Context {}
unsafe impl Sync for Context {}

impl Context {
    /// A context on CUDA device `device` with a private stream.
    pub fn new(device: i32) -> Arc<Context> {
        Self::with_stream(device, std::ptr::null_mut())
    }

    /// Adopt an existing cudaStream_t (`cudaStreamLegacy` = 0x1 for the legacy default stream; NULL creates a private one).
    pub fn with_stream(device: i32, stream: *mut c_void) -> Arc<Context> {
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { solb_ctx_create(device, stream, &mut raw) };
        if rc != 0 {
            panic!("solb_ctx_create: {}", last_error(std::ptr::null_mut()));
        }
        Arc::new(Context { raw })
    }

    pub(crate) fn check(&self, rc: i32) {
        if rc != 0 {
            panic!("libsolb: {}", last_error(self.raw));
        }
    }

    /// queue_wait_idle (src/context.rs:539-559)
    pub fn synchronize(&self) {
        self.check(unsafe { solb_synchronize(self.raw) });
    }

    /// Loads every kernel now (what shader compilation in Pipeline::new is to the reference). Idempotent.
    pub fn preload(&self) {
        self.check(unsafe { solb_ctx_preload(self.raw) });
    }

    /// Returns cached build scratch to the driver.
    pub fn trim(&self) {
        self.check(unsafe { solb_ctx_trim(self.raw) });
    }

    pub fn stats(&self) -> SolbStats {
        let mut s = SolbStats::default();
        self.check(unsafe { solb_stats_get(self.raw, &mut s) });
        s
    }

    pub fn reset_stats(&self) {
        self.check(unsafe { solb_stats_reset(self.raw) });
    }

    /// cudaEvent pairs around builds / traces (the reference's timestamp queries, src/renderer.rs:204-225).
    pub fn set_timing(&self, enabled: bool) {
        self.check(unsafe { solb_set_timing(self.raw, enabled as i32) });
    }

    /// Binding 2 of 4-ray-ao (examples/4-ray-ao.rs:282): rgba8 texels, rows already flipped like Texture2d::new.
    pub fn set_blue_noise(&self, rgba8: &[u8], width: u32, height: u32) {
        assert_eq!(rgba8.len(), (width as usize) * (height as usize) * 4);
        self.check(unsafe { solb_set_blue_noise(self.raw, rgba8.as_ptr(), width, height) });
    }

    pub fn version() -> u32 {
        unsafe { solb_version() }
    }
}

/// Per-frame fence of a render loop with frames in flight (AppFrameData::in_flight_fence, src/renderer.rs:8): created
/// signalled; `signal` after the frame's work is enqueued (queue_submit(.., fence), src/renderer.rs:310-317), `wait` before
/// the slot is reused (wait_for_and_reset_fence, src/renderer.rs:123-131).
pub struct Fence {
    context: Arc<Context>,
    raw: *mut solb_fence,
}

impl Fence {
    pub fn new(context: Arc<Context>) -> Fence {
        let mut raw = std::ptr::null_mut();
        context.check(unsafe { solb_fence_create(context.raw, &mut raw) });
        Fence { context, raw }
    }

    pub fn signal(&self) {
        self.context.check(unsafe { solb_fence_signal(self.raw) });
    }

    pub fn wait(&self) {
        self.context.check(unsafe { solb_fence_wait(self.raw) });
    }
}

impl Drop for Fence {
    fn drop(&mut self) {
        unsafe {
            solb_fence_destroy(self.raw);
        }
    }
}

/// Page-locked host bytes for uploads and read-backs (the reference's CpuToGpu / GpuToCpu buffers, src/buffer.rs:76-83).
pub struct HostBuffer {
    context: Arc<Context>,
    ptr: *mut u8,
    len: usize,
}

impl HostBuffer {
    pub fn new(context: Arc<Context>, len: usize) -> HostBuffer {
        let mut p: *mut c_void = std::ptr::null_mut();
        context.check(unsafe { solb_host_alloc(context.raw, len, &mut p) });
        HostBuffer { context, ptr: p as *mut u8, len }
    }

    pub fn as_slice(&self) -> &[u8] {
        unsafe { std::slice::from_raw_parts(self.ptr, self.len) }
    }

    pub fn as_mut_slice(&mut self) -> &mut [u8] {
        unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}

impl Drop for HostBuffer {
    fn drop(&mut self) {
        unsafe {
            solb_host_free(self.context.raw, self.ptr as *mut c_void);
        }
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe {
            solb_ctx_destroy(self.raw);
        }
    }
}

pub(crate) fn last_error(ctx: *mut solb_ctx) -> String {
    unsafe { CStr::from_ptr(solb_last_error(ctx)).to_string_lossy().into_owned() }
}
