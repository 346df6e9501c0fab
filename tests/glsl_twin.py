"""A second, independent reading of the reference's path-tracing shaders in plain Python / numpy float32 — TEST-ONLY.

oracle/oracle.c is the oracle; this file exists to catch TRANSCRIPTION errors in it: the same GLSL
(/root/reference/assets/glsl/pathtrace.rgen:39-104, pathtrace.rchit:56-113, pathtrace.rmiss:8-21, sampling.glsl:18-97) restated a
second time, statement by statement, with none of oracle.c's helper structure.  The closest hit itself (the Vulkan driver's part) is
taken from oracle.Scene.trace_rays, so what is cross-checked is ray generation, RNG consumption order, hit shading, BRDF sampling,
sky, the bounce cap and the per-frame mean — everything the reference computes in shader code.  render_ao_frame does the same for
ao.rgen:36-83 / ao.rchit:43-89 / ao.rmiss.  Far too slow for anything but a few hundred pixels.
"""
import numpy as np

import oracle
from oracle import gltf_flatten as gf

F = np.float32
MISS = 0xFFFFFFFF


def _v3(x, y, z):
    return np.array([x, y, z], dtype=F)


def _dot(a, b):
    return F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2]))


def _length(a):
    return F(np.sqrt(_dot(a, a)))


def _normalize(a):
    return (a / _length(a)).astype(F)


def _mat_vec4(m, v):
    """column-major 16 floats (glam / GLSL mat4) times vec4"""
    out = np.zeros(4, dtype=F)
    for row in range(4):
        acc = F(0)
        for col in range(4):
            acc = F(acc + F(m[4 * col + row] * v[col]))
        out[row] = acc
    return out


def _saturate(x):
    return F(min(max(x, F(0)), F(1)))


class Rng:
    def __init__(self, state):
        self.s = state & 0xFFFFFFFF

    def next(self):  # sampling.glsl:36-43
        self.s = (self.s * 747796405 + 1) & 0xFFFFFFFF
        word = (((self.s >> ((self.s >> 28) + 4)) ^ self.s) * 277803737) & 0xFFFFFFFF
        word = ((word >> 22) ^ word) & 0xFFFFFFFF
        return F(F(word) / F(4294967295.0))


def tea(v0, v1):  # sampling.glsl:18-33
    s0 = 0
    for _ in range(16):
        s0 = (s0 + 0x9E3779B9) & 0xFFFFFFFF
        v0 = (v0 + ((((v1 << 4) + 0xA341316C) & 0xFFFFFFFF) ^ ((v1 + s0) & 0xFFFFFFFF) ^ (((v1 >> 5) + 0xC8013EA4) & 0xFFFFFFFF))) & 0xFFFFFFFF
        v1 = (v1 + ((((v0 << 4) + 0xAD90777D) & 0xFFFFFFFF) ^ ((v0 + s0) & 0xFFFFFFFF) ^ (((v0 >> 5) + 0x7E95761E) & 0xFFFFFFFF))) & 0xFFFFFFFF
    return v0


def fresnel_dielectric(i, m, eta):  # sampling.glsl:53-65
    result = F(1)
    cos_i = F(abs(_dot(i, m)))
    sin_o2 = F(F(eta * eta) * F(F(1) - F(cos_i * cos_i)))
    if sin_o2 <= F(1):
        cos_o = F(np.sqrt(_saturate(F(F(1) - sin_o2))))
        rs = F(F(cos_i - F(eta * cos_o)) / F(cos_i + F(eta * cos_o)))
        rp = F(F(F(eta * cos_i) - cos_o) / F(F(eta * cos_i) + cos_o))
        result = F(F(0.5) * F(F(rs * rs) + F(rp * rp)))
    return result


def align_to_direction(n, cos_theta, phi):  # sampling.glsl:67-86
    sin_theta = F(np.sqrt(_saturate(F(F(1) - F(cos_theta * cos_theta)))))
    s = F(-1) if n[2] < 0 else F(1)
    a = F(F(-1) / F(s + n[2]))
    b = F(F(n[0] * n[1]) * a)
    u = _v3(F(F(1) + F(F(F(s * n[0]) * n[0]) * a)), F(s * b), F(-s * n[0]))
    v = _v3(b, F(s + F(F(n[1] * n[1]) * a)), F(-n[1]))
    c, sn = F(np.cos(phi)), F(np.sin(phi))
    return ((u * c + v * sn).astype(F) * sin_theta + n * cos_theta).astype(F)


TWO_PI = F(6.28318530718)


def sample_ggx(n, xi, alpha2):  # sampling.glsl:88-92
    cos_theta = F(np.sqrt(_saturate(F(F(F(1) - xi[0]) / F(F(xi[0] * F(alpha2 - F(1))) + F(1))))))
    return align_to_direction(n, cos_theta, F(xi[1] * TWO_PI))


def sample_cosine(n, xi):  # sampling.glsl:94-98
    return align_to_direction(n, F(np.sqrt(xi[0])), F(xi[1] * TWO_PI))


def reflect(i, n):
    return (i - n * F(F(2) * _dot(n, i))).astype(F)


def shade_miss(enable_sky, d):  # pathtrace.rmiss:8-21
    if not enable_sky:
        return _v3(0, 0, 0)
    wi = _normalize(d)
    x = _saturate(F(F(F(F(0.5) * F(wi[1] + F(1))) - F(0.35)) / F(F(0.65) - F(0.35))))
    t = F(F(x * x) * F(F(3) - F(F(2) * x)))  # smoothstep
    sky = (_v3(0.58, 0.45, 0.25) * F(F(1) - t) + _v3(0.3, 0.4, 0.5) * t).astype(F)
    is_sun = _dot(wi, _normalize(_v3(0.0, 1.0, -0.25))) > F(0.99)
    return _v3(120.0, 100.0, 50.0) if is_sun else sky


def render_frame(fs, osc, uniforms, w, h, enable_sky, spp, max_bounces):
    """One frame of pathtrace.rgen into a fresh accumulation image (frame == accum_start_frame, so alpha = 1)."""
    u = np.frombuffer(uniforms, dtype=F)
    view_inv, proj_inv = u[32:48], u[64:80]
    frame = int(np.frombuffer(uniforms, dtype=np.uint32)[98])
    verts, indices = np.asarray(fs.vertices, dtype=F), np.asarray(fs.indices)
    inst_data = []
    for inst in fs.instances:
        t = np.asarray(inst["transform"], dtype=F).reshape(16)
        t_it = np.asarray(gf.mat4_inverse(np.asarray(inst["transform"], dtype=F).reshape(4, 4)).T, dtype=F).reshape(16)
        inst_data.append((inst, t, t_it, np.asarray(fs.materials[inst["material"]], dtype=F)))
    out = np.zeros((h, w, 4), dtype=F)
    n_rays = 0
    for y in range(h):
        for x in range(w):
            rng = Rng(tea(x + y * w, frame))  # rgen:47
            pixel = _v3(0, 0, 0)
            for _ in range(spp):
                jx = rng.next()
                jy = rng.next()
                in_uv = np.array([F(F(x) + jx) / F(w), F(F(y) + jy) / F(h)], dtype=F)
                d = (in_uv * F(2) - F(1)).astype(F)
                origin = _mat_vec4(view_inv, np.array([0, 0, 0, 1], dtype=F))
                target = _mat_vec4(proj_inv, np.array([d[0], d[1], 1, 1], dtype=F))
                tn = _normalize(target[:3])
                direction = _mat_vec4(view_inv, np.array([tn[0], tn[1], tn[2], 0], dtype=F))
                ray_o, ray_d = origin[:3].copy(), direction[:3].copy()
                tmin, tmax = F(max(F(1), _length(origin[:3])) * F(1e-3)), F(10000.0)  # rgen:35
                depth, acc = 0, _v3(1, 1, 1)
                while True:
                    ray = np.array([[ray_o[0], ray_o[1], ray_o[2], tmin, ray_d[0], ray_d[1], ray_d[2], tmax]], dtype=F)
                    hits, _t, _ = osc.trace_rays(ray)
                    n_rays += 1
                    inst_id, prim = int(hits[0, 0]), int(hits[0, 1])
                    done = False
                    if inst_id == MISS:
                        hit_value, done = shade_miss(enable_sky, ray_d), True
                    else:
                        inst, t, t_it, mat = inst_data[inst_id]
                        bu, bv = hits[0, 2:4].view(F)
                        if mat[4] >= 1 or mat[5] >= 1 or mat[6] >= 1:  # rchit:72-77
                            hit_value, done = mat[4:7].copy(), True
                            depth += 1
                        else:
                            tri = [verts[inst["first_vertex"] + int(indices[inst["first_index"] + 3 * prim + k])] for k in range(3)]
                            bary = _v3(F(F(F(1) - bu) - bv), bu, bv)

                            def interp(lo):
                                return ((tri[0][lo:lo + 3] * bary[0]).astype(F) + (tri[1][lo:lo + 3] * bary[1]).astype(F)
                                        + (tri[2][lo:lo + 3] * bary[2]).astype(F)).astype(F)

                            nrm = interp(8)
                            nrm = _normalize(_mat_vec4(t_it, np.array([nrm[0], nrm[1], nrm[2], 0], dtype=F))[:3])
                            pos = interp(0)
                            world = _mat_vec4(t, np.array([pos[0], pos[1], pos[2], 1], dtype=F))[:3]
                            vcol = interp(4)
                            wi = _normalize(ray_d)
                            sgn = F(np.sign(_dot(nrm, -wi)))
                            n_o = (nrm * sgn).astype(F)
                            alpha2 = F(mat[9] * mat[9])
                            xi = (rng.next(), rng.next())
                            rnd = rng.next()
                            ray_o_new = (world + n_o * F(0.0001)).astype(F)
                            base = (mat[0:3] * vcol).astype(F)
                            if rnd < mat[8]:
                                ray_d_new = sample_ggx(reflect(ray_d, n_o), xi, alpha2)
                                hit_value = base
                            else:
                                m = sample_ggx(n_o, xi, alpha2)
                                if rnd < fresnel_dielectric(n_o, m, F(F(1.0) / F(1.5))):
                                    ray_d_new = reflect(ray_d, m)
                                    hit_value = _v3(1, 1, 1)
                                else:
                                    ray_d_new = sample_cosine(n_o, xi)
                                    hit_value = base
                            ray_o, ray_d = ray_o_new, ray_d_new
                            depth += 1
                    acc = (acc * hit_value).astype(F)
                    if done:
                        break
                    if depth > max_bounces:
                        acc = _v3(0, 0, 0)
                        break
                pixel = (pixel + acc).astype(F)
            pixel = (pixel * F(F(1) / F(spp))).astype(F)
            out[y, x, :3] = pixel
            out[y, x, 3] = 1
    return out, n_rays


def render_ao_frame(fs, osc, uniforms, w, h, blue):
    """One frame of ao.rgen:36-83 + ao.rchit:43-89 + ao.rmiss into a fresh image (frame == accum_start_frame, a = 1).
    blue = the blue-noise texture as uploaded (rgba8, [y][x][c])."""
    u = np.frombuffer(uniforms, dtype=F)
    view_inv, proj_inv = u[32:48], u[64:80]
    frame = int(np.frombuffer(uniforms, dtype=np.uint32)[98])
    verts, indices = np.asarray(fs.vertices, dtype=F), np.asarray(fs.indices)
    inst_data = []
    for inst in fs.instances:
        t = np.asarray(inst["transform"], dtype=F).reshape(16)
        t_it = np.asarray(gf.mat4_inverse(np.asarray(inst["transform"], dtype=F).reshape(4, 4)).T, dtype=F).reshape(16)
        inst_data.append((inst, t, t_it))
    th, tw = blue.shape[0], blue.shape[1]
    max_samples, sample_count = 4, 4
    out = np.zeros((h, w, 4), dtype=F)
    n_rays = 0
    for y in range(h):
        for x in range(w):
            rng = Rng(tea(x + y * w, frame))
            ao = _v3(0, 0, 0)
            for s in range(sample_count):
                jx = rng.next()
                jy = rng.next()
                in_uv = np.array([F(F(x) + jx) / F(w), F(F(y) + jy) / F(h)], dtype=F)
                d = (in_uv * F(2) - F(1)).astype(F)
                origin = _mat_vec4(view_inv, np.array([0, 0, 0, 1], dtype=F))
                target = _mat_vec4(proj_inv, np.array([d[0], d[1], 1, 1], dtype=F))
                tn = _normalize(target[:3])
                direction = _mat_vec4(view_inv, np.array([tn[0], tn[1], tn[2], 0], dtype=F))
                ray_o, ray_d = origin[:3].copy(), direction[:3].copy()
                tmin, tmax = F(max(F(1), _length(origin[:3])) * F(1e-3)), F(10000.0)
                depth, hit_value = 0, _v3(0, 0, 0)
                while True:
                    ray = np.array([[ray_o[0], ray_o[1], ray_o[2], tmin, ray_d[0], ray_d[1], ray_d[2], tmax]], dtype=F)
                    hits, _t, _ = osc.trace_rays(ray)
                    n_rays += 1
                    inst_id, prim = int(hits[0, 0]), int(hits[0, 1])
                    if inst_id == MISS:
                        break  # ao.rmiss: done = 1
                    inst, t, t_it = inst_data[inst_id]
                    bu, bv = hits[0, 2:4].view(F)
                    tri = [verts[inst["first_vertex"] + int(indices[inst["first_index"] + 3 * prim + k])] for k in range(3)]
                    bary = _v3(F(F(F(1) - bu) - bv), bu, bv)

                    def interp(lo):
                        return ((tri[0][lo:lo + 3] * bary[0]).astype(F) + (tri[1][lo:lo + 3] * bary[1]).astype(F)
                                + (tri[2][lo:lo + 3] * bary[2]).astype(F)).astype(F)

                    nrm = interp(8)
                    nrm = _normalize(_mat_vec4(t_it, np.array([nrm[0], nrm[1], nrm[2], 0], dtype=F))[:3])
                    pos = interp(0)
                    world = _mat_vec4(t, np.array([pos[0], pos[1], pos[2], 1], dtype=F))[:3]
                    ray_o_new = (world + ray_d * F(0.00001)).astype(F)
                    # getBlueRand2(depth + depth * sampleId)
                    i = depth + depth * s
                    rx = rng.next()
                    ry = rng.next()
                    fx, fy = F(F(x) + F(rx * F(tw))), F(F(y) + F(ry * F(th)))
                    cx = int(F(fx - F(F(tw) * F(np.floor(F(fx / F(tw)))))))
                    cy = int(F(fy - F(F(th) * F(np.floor(F(fy / F(th)))))))
                    texel = blue[cy, cx].astype(F) / F(255.0)
                    xi = (F(texel[i % 4]), F(texel[(i + 1) % 4]))
                    hit_norm = (nrm * F(np.sign(_dot(-ray_d, nrm)))).astype(F)
                    ray_d_new = sample_cosine(hit_norm, xi)
                    tmin, tmax = F(0.001), F(10.0)
                    if depth > 0:
                        hit_value = (hit_value + _v3(1, 1, 1)).astype(F)
                    depth += 1
                    ray_o, ray_d = ray_o_new, ray_d_new
                    if depth > max_samples:
                        break
                ao = (ao + hit_value * F(F(1) / F(max_samples))).astype(F)
            color = (_v3(1, 1, 1) - ao / F(sample_count)).astype(F)
            out[y, x, :3] = color
            out[y, x, 3] = 1
    return out, n_rays
