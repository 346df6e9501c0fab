//! sol::Image2d storage images for this path (src/texture.rs:36-96 + examples/5-pathtrace.rs:57-80 create_image_target).
use crate::ffi::*;
use crate::Context;
use std::os::raw::c_void;
use std::sync::Arc;

/// The three formats the path uses (vk::Format of the reference in the comments).
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum ImageFormat {
    /// R32G32B32A32_SFLOAT: accumulation image (examples/5-pathtrace.rs:222)
    Rgba32f,
    /// R8G8B8A8_UNORM: render image (examples/5-pathtrace.rs:229)
    Rgba8,
    /// (instance, primitive) ids of the primary hit; 0xffffffff = miss
    Rg32ui,
}

impl ImageFormat {
    fn raw(self) -> u32 {
        match self {
            ImageFormat::Rgba32f => SOLB_FORMAT_RGBA32F,
            ImageFormat::Rgba8 => SOLB_FORMAT_RGBA8,
            ImageFormat::Rg32ui => SOLB_FORMAT_RG32UI,
        }
    }
    pub fn bytes_per_pixel(self) -> usize {
        match self {
            ImageFormat::Rgba32f => 16,
            ImageFormat::Rgba8 => 4,
            ImageFormat::Rg32ui => 8,
        }
    }
}

/// src/texture.rs ImageInfo, reduced to what create_image_target sets.
#[derive(Clone, Copy, Debug)]
pub struct ImageInfo {
    pub width: u32,
    pub height: u32,
    pub format: ImageFormat,
}

pub struct Image2d {
    context: Arc<Context>,
    pub(crate) raw: *mut solb_target,
    info: ImageInfo,
}

impl Image2d {
    /// Zero-initialised device image.
    pub fn new(context: Arc<Context>, info: ImageInfo) -> Image2d {
        let mut raw = std::ptr::null_mut();
        context.check(unsafe { solb_target_create(context.raw, info.width, info.height, info.format.raw(), &mut raw) });
        Image2d { context, raw, info }
    }

    pub fn get_info(&self) -> ImageInfo {
        self.info
    }

    pub fn size_bytes(&self) -> usize {
        self.info.width as usize * self.info.height as usize * self.info.format.bytes_per_pixel()
    }

    pub fn clear(&self) {
        self.context.check(unsafe { solb_target_clear(self.raw) });
    }

    /// Replaces cmd_blit_to(present image) (examples/5-pathtrace.rs:360-361): the whole image to host memory.
    pub fn readback(&self, host: &mut [u8]) {
        assert_eq!(host.len(), self.size_bytes());
        self.context.check(unsafe { solb_target_readback(self.raw, host.as_mut_ptr() as *mut c_void, host.len()) });
    }

    /// The same copy enqueued without the wait: `host` is complete once a `Fence` signalled after this call has been waited
    /// for (frames in flight; the reference's blit is a queued command as well).
    pub fn readback_async(&self, host: &mut crate::HostBuffer) {
        let bytes = host.as_mut_slice();
        assert_eq!(bytes.len(), self.size_bytes());
        self.context.check(unsafe { solb_target_readback_async(self.raw, bytes.as_mut_ptr() as *mut c_void, bytes.len()) });
    }

    pub fn upload(&self, host: &[u8]) {
        assert_eq!(host.len(), self.size_bytes());
        self.context.check(unsafe { solb_target_upload(self.raw, host.as_ptr() as *const c_void, host.len()) });
    }

    /// Raw device pointer (CUDA interop).
    pub fn device_ptr(&self) -> *mut c_void {
        let mut p = std::ptr::null_mut();
        self.context.check(unsafe { solb_target_device_ptr(self.raw, &mut p) });
        p
    }
}

impl Drop for Image2d {
    fn drop(&mut self) {
        unsafe {
            solb_target_destroy(self.raw);
        }
    }
}
