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
