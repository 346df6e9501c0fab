"""Offscreen outputs: the reference blits the render target to the swapchain and presents it
(examples/5-pathtrace.rs:360-368); headless, the frame is written to disk instead.  Also an accumulation
checkpoint — the reference keeps the accumulation image alive across frames and restarts it through
`accumulation_start_frame` (examples/5-pathtrace.rs:50,243,286); saving (accum, start frame, next frame) lets a
long convergence run (config 3: 512 frames) resume with bit-identical results."""
import struct
import zlib

import numpy as np

from . import _native as N


def write_png(path, rgba8):
    """rgba8: uint8 [h, w, 4] (or [h, w, 3]) -> 8-bit PNG, no external dependency."""
    a = np.ascontiguousarray(rgba8, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] in (3, 4)
    h, w, c = a.shape
    raw = np.concatenate([np.zeros((h, 1), dtype=np.uint8), a.reshape(h, w * c)], axis=1).tobytes()  # filter 0 per row

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6 if c == 4 else 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))


def write_ppm(path, rgba8):
    a = np.ascontiguousarray(rgba8, dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (a.shape[1], a.shape[0]))
        f.write(a[..., :3].tobytes())


def save_checkpoint(path, accum_target, accumulation_start_frame, next_frame):
    """accum_target: sol_rs_b200.Image2d (rgba32f)."""
    acc = accum_target.readback()
    np.savez(path, accum=acc, accumulation_start_frame=np.int64(accumulation_start_frame), next_frame=np.int64(next_frame),
             width=np.int64(accum_target.width), height=np.int64(accum_target.height))


def load_checkpoint(path, context):
    """-> (Image2d accum, accumulation_start_frame, next_frame)"""
    from .context import Image2d

    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    img = Image2d(context, int(z["width"]), int(z["height"]), N.FORMAT_RGBA32F)
    img.upload(z["accum"])
    return img, int(z["accumulation_start_frame"]), int(z["next_frame"])


def load_texture_rgba8(path, flipv=True):
    """Texture2d::new (src/texture.rs:488-494): image::open(path) -> flipv() -> to_rgba8().  Returns uint8 [h, w, 4].
    16-bit channels become 8-bit by the image 0.24 crate's rule (v + 128) / 257; grey / grey-alpha / rgb are expanded
    with alpha 255.  Decoding itself is delegated to OpenCV or Pillow (host-side asset loading, not the hot path)."""
    im = None
    try:
        import cv2

        im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if im is not None and im.ndim == 3 and im.shape[2] >= 3:
            im = im[:, :, [2, 1, 0] + ([3] if im.shape[2] == 4 else [])]  # BGR(A) -> RGB(A)
    except ImportError:
        pass
    if im is None:
        from PIL import Image

        im = np.asarray(Image.open(path))
    if im.ndim == 2:
        im = im[:, :, None]
    if im.dtype == np.uint16:
        im = ((im.astype(np.uint32) + 128) // 257).astype(np.uint8)
    elif im.dtype != np.uint8:
        raise ValueError("unsupported channel type %s in %s" % (im.dtype, path))
    h, w, c = im.shape
    out = np.full((h, w, 4), 255, dtype=np.uint8)
    if c == 1:
        out[..., :3] = im
    elif c == 2:
        out[..., :3] = im[..., :1]
        out[..., 3] = im[..., 1]
    else:
        out[..., :c] = im[..., :c]
    if flipv:
        out = out[::-1]
    return np.ascontiguousarray(out)
