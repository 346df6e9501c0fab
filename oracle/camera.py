"""Camera + SceneUniforms restatement.  ORACLE SIDE (test infrastructure only).

Follows /root/reference  src/scene/camera.rs:54-127,264-270  and  examples/5-pathtrace.rs:7-32.
The matrix helpers restate glam 0.20.2 (Cargo.lock:509-510), which is not vendored:
`Mat4::perspective_rh`, `Mat4::look_at_rh` (= look_to_lh(eye, eye - center, up)),
`Mat4::inverse` (see gltf_flatten.mat4_inverse), `f32::to_radians`.  PARITY UNPINNED.
Arrays are indexed [col][row] (column-major, like glam's cols array).
"""
import math

import numpy as np

from .gltf_flatten import F32, mat4_inverse, mat4_mul


def _sin_cos_f32(x):
    # glam calls f32::sin_cos; libm sinf/cosf are correctly rounded to <1 ulp, so evaluate in f64 and round
    return F32(math.sin(float(x))), F32(math.cos(float(x)))


def perspective_rh(fov_y_radians, aspect, z_near, z_far):
    """glam Mat4::perspective_rh: RH, depth 0..1, no Y flip."""
    fov_y_radians, aspect, z_near, z_far = F32(fov_y_radians), F32(aspect), F32(z_near), F32(z_far)
    s, c = _sin_cos_f32(F32(0.5) * fov_y_radians)
    h = c / s
    w = h / aspect
    r = z_far / (z_near - z_far)
    m = np.zeros((4, 4), dtype=F32)
    m[0] = [w, 0, 0, 0]
    m[1] = [0, h, 0, 0]
    m[2] = [0, 0, r, -1]
    m[3] = [0, 0, r * z_near, 0]
    return m


def _normalize(v):
    v = v.astype(F32)
    l = F32(np.sqrt(F32(F32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])))
    return (v * (F32(1.0) / l)).astype(F32)  # glam: self.mul(self.length_recip())


def _cross(a, b):
    return np.array([a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]], dtype=F32)


def _dot(a, b):
    return F32(F32(a[0] * b[0] + a[1] * b[1]) + a[2] * b[2])


def look_at_rh(eye, center, up):
    """glam Mat4::look_at_rh -> look_to_lh(eye, eye - center, up)."""
    eye = np.array(eye, dtype=F32)
    center = np.array(center, dtype=F32)
    up = np.array(up, dtype=F32)
    f = _normalize(eye - center)
    s = _normalize(_cross(up, f))
    u = _cross(f, s)
    m = np.zeros((4, 4), dtype=F32)
    m[0] = [s[0], u[0], f[0], 0]
    m[1] = [s[1], u[1], f[1], 0]
    m[2] = [s[2], u[2], f[2], 0]
    m[3] = [-_dot(s, eye), -_dot(u, eye), -_dot(f, eye), 1]
    return m


def to_radians(deg):
    return F32(F32(deg) * F32(F32(math.pi) / F32(180.0)))  # Rust: self * (PI / 180.0)


class Camera:
    """src/scene/camera.rs:34-127 (matrices only; the mouse manipulators are out of scope)."""

    def __init__(self, window_size):
        # Camera::new, camera.rs:54-71
        self.position = np.array([10, 10, 10], dtype=F32)
        self.center = np.zeros(3, dtype=F32)
        self.up = np.array([0, -1, 0], dtype=F32)
        self.vfov = F32(35.0)
        self.z_near = F32(0.1)
        self.z_far = F32(1000.0)
        self.view = np.eye(4, dtype=F32)
        self.persp = np.eye(4, dtype=F32)
        self.window_size = (F32(window_size[0]), F32(window_size[1]))
        self._update_persp()

    @classmethod
    def from_view(cls, view, yfov, z_near, z_far):
        # camera.rs:73-94: persp stays IDENTITY until set_window_size; window 1920x1080
        cam = cls.__new__(cls)
        vi = mat4_inverse(np.array(view, dtype=F32))
        cam.position = vi[3, 0:3].copy()
        cam.up = vi[1, 0:3].copy()
        cam.center = (cam.position + vi[2, 0:3] * F32(-4.0)).astype(F32)
        cam.vfov = F32(yfov)
        cam.z_near = F32(z_near)
        cam.z_far = F32(z_far)
        cam.view = np.array(view, dtype=F32)
        cam.persp = np.eye(4, dtype=F32)
        cam.window_size = (F32(1920.0), F32(1080.0))
        return cam

    def _update_persp(self):
        aspect = self.window_size[0] / self.window_size[1]
        self.persp = perspective_rh(to_radians(self.vfov), aspect, self.z_near, self.z_far)  # camera.rs:102-106

    def look_at(self, eye, center, up):
        self.position = np.array(eye, dtype=F32)
        self.center = np.array(center, dtype=F32)
        self.up = np.array(up, dtype=F32)
        self.view = look_at_rh(self.position, self.center, self.up)  # camera.rs:98-100

    def set_window_size(self, size):
        self.window_size = (F32(size[0]), F32(size[1]))
        self._update_persp()

    def set_vfov(self, vfov):
        self.vfov = F32(vfov)
        self._update_persp()


def scene_uniforms(camera, width, height, frame):
    """SceneUniforms::from (examples/5-pathtrace.rs:19-31) -> 400-byte block as bytes."""
    buf = np.zeros(100, dtype=F32)
    vp = mat4_mul(camera.persp, camera.view)
    mats = [np.eye(4, dtype=F32), camera.view, mat4_inverse(camera.view), camera.persp, mat4_inverse(camera.persp), vp]
    for i, m in enumerate(mats):
        buf[16 * i : 16 * i + 16] = np.asarray(m, dtype=F32).reshape(16)
    raw = bytearray(buf.tobytes())
    raw[384:396] = np.array([width, height, frame], dtype=np.uint32).tobytes()
    return bytes(raw)
