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


def write_exr(path, rgba32f, channels="RGB"):
    """float32 [h, w, >= len(channels)] -> OpenEXR 2 scan-line file, uncompressed, 32-bit float channels (the accumulation
    target keeps linear radiance: SURVEY 8f item 2 names EXR for it).  No external dependency: magic, version 2, the eight
    required attributes, the scan-line offset table, then per line (y, byte count, channels in alphabetical order)."""
    a = np.ascontiguousarray(rgba32f, dtype=np.float32)
    assert a.ndim == 3 and a.shape[2] >= len(channels)
    h, w = a.shape[:2]
    names = sorted(channels)  # the file stores channels alphabetically
    src = {c: "RGBA".index(c) for c in names}

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data

    chlist = b"".join(c.encode() + b"\0" + struct.pack("<iBxxxii", 2, 0, 1, 1) for c in names) + b"\0"  # 2 = FLOAT
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = (struct.pack("<ii", 20000630, 2) + attr("channels", "chlist", chlist) + attr("compression", "compression", b"\0")
              + attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0")
              + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
              + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    line_bytes = len(names) * w * 4
    first = len(header) + 8 * h
    with open(path, "wb") as f:
        f.write(header)
        f.write(struct.pack("<%dq" % h, *[first + y * (8 + line_bytes) for y in range(h)]))
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            for c in names:
                f.write(a[y, :, src[c]].astype("<f4").tobytes())


def read_exr(path):
    """the files write_exr makes (uncompressed float scan lines) -> (float32 [h, w, n], channel names); for tests / resuming"""
    raw = open(path, "rb").read()
    assert struct.unpack_from("<i", raw, 0)[0] == 20000630
    off, attrs = 8, {}
    while raw[off] != 0:
        e = raw.index(b"\0", off)
        name = raw[off:e].decode()
        e2 = raw.index(b"\0", e + 1)
        n = struct.unpack_from("<i", raw, e2 + 1)[0]
        attrs[name] = raw[e2 + 5: e2 + 5 + n]
        off = e2 + 5 + n
    off += 1
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    assert attrs["compression"] == b"\0"
    names, p = [], 0
    ch = attrs["channels"]
    while ch[p] != 0:
        e = ch.index(b"\0", p)
        names.append(ch[p:e].decode())
        assert struct.unpack_from("<i", ch, e + 1)[0] == 2
        p = e + 17
    offs = struct.unpack_from("<%dq" % h, raw, off)
    out = np.zeros((h, w, len(names)), dtype=np.float32)
    for y in range(h):
        yy, nb = struct.unpack_from("<ii", raw, offs[y])
        out[yy - y0] = np.frombuffer(raw, dtype="<f4", count=len(names) * w, offset=offs[y] + 8).reshape(len(names), w).T
    return out, names


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
