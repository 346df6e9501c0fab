"""Shared helpers: build the same scene/camera on the oracle side and on the product side."""
import os

import numpy as np

import oracle
from oracle import camera as ocam
from oracle import gltf_flatten as gf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODELS = os.path.join(ROOT, "assets", "models")


def model_path(name):
    return os.path.join(MODELS, name + ".gltf")


def oracle_scene(name):
    fs = gf.load_scene(model_path(name))
    return fs, oracle.Scene(fs)


def oracle_camera(fs, name, w, h):
    """cornell/tunnel: the file camera (5-pathtrace); Duck: the 3-ray-debug camera; Duck_ao: the 4-ray-ao camera."""
    if name == "Duck":
        c = ocam.Camera((w, h))
        c.look_at((5, 5, 5), (0, 0, 0), (0, -1, 0))  # examples/3-ray-debug.rs:78-79
    elif name == "Duck_ao":
        c = ocam.Camera((w, h))
        c.look_at((4, 1, 4), (0, 0.5, 0), (0, -1, 0))  # examples/4-ray-ao.rs:89-90
    else:
        c = ocam.Camera.from_view(fs.camera["view"], fs.camera["yfov"], fs.camera["znear"], fs.camera["zfar"])
        c.set_window_size((w, h))
    return c


def product_camera(scene_obj, name, w, h):
    from sol_rs_b200 import scene

    if name == "Duck":
        c = scene.Camera((w, h))
        c.look_at((5, 5, 5), (0, 0, 0), (0, -1, 0))
    elif name == "Duck_ao":
        c = scene.Camera((w, h))
        c.look_at((4, 1, 4), (0, 0.5, 0), (0, -1, 0))
    else:
        c = scene_obj.camera
        c.set_window_size((w, h))
    return c


def pathtrace_pipeline(ctx, enable_sky):
    from sol_rs_b200 import ray

    pipe = ray.Pipeline(ctx, ray.PipelineInfo().shader("glsl/pathtrace.rgen", ray.RAYGEN_KHR)
                        .shader("glsl/pathtrace.rmiss", ray.MISS_KHR).shader("glsl/pathtrace.rchit", ray.CLOSEST_HIT_KHR)
                        .specialization([1 if enable_sky else 0], 0).name("pathtrace"))
    return ray.ShaderBindingTable(ctx, pipe, ray.ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2))


def simple_pipeline(ctx, kind):
    from sol_rs_b200 import ray

    pipe = ray.Pipeline(ctx, ray.PipelineInfo().shader("glsl/%s.rgen" % kind, ray.RAYGEN_KHR)
                        .shader("glsl/%s.rmiss" % kind, ray.MISS_KHR).shader("glsl/%s.rchit" % kind, ray.CLOSEST_HIT_KHR))
    return ray.ShaderBindingTable(ctx, pipe, ray.ShaderBindingTableInfo().raygen(0).miss(1).hitgroup(2))


def load_blue_noise():
    """assets/textures/HDR_RGBA_0.png as the reference uploads it: image::open -> flipv -> to_rgba8
    (src/texture.rs:490-493); 16-bit -> 8-bit by (v + 128) / 257 (image 0.24 crate rule)."""
    import cv2

    im = cv2.imread(os.path.join(ROOT, "assets", "textures", "HDR_RGBA_0.png"), cv2.IMREAD_UNCHANGED)
    assert im is not None and im.dtype == np.uint16 and im.shape[2] == 4
    im = im[:, :, [2, 1, 0, 3]]  # BGRA -> RGBA
    im = im[::-1]                # flipv
    return ((im.astype(np.uint32) + 128) // 257).astype(np.uint8).copy()


def image_metrics(a, b):
    """mean relative error on float images and PSNR on their gamma-2.2 rgba8 versions (north_star gates)."""
    a3, b3 = a[..., :3].astype(np.float64), b[..., :3].astype(np.float64)
    mre = np.abs(a3 - b3).sum() / max(np.abs(b3).sum(), 1e-12)
    ga = np.clip(np.power(np.clip(a3, 0, None), 1 / 2.2), 0, 1) * 255
    gb = np.clip(np.power(np.clip(b3, 0, None), 1 / 2.2), 0, 1) * 255
    mse = np.mean((np.rint(ga) - np.rint(gb)) ** 2)
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    return mre, psnr


def trs(translate=(0, 0, 0), axis=(0, 1, 0), angle=0.0, scale=(1, 1, 1)):
    """column-major-by-convention 4x4 (stored like glam::Mat4::to_cols_array_2d, i.e. m[col][row]) of T * R * S."""
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    c, s = np.cos(angle), np.sin(angle)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + s * K + (1 - c) * (K @ K)
    M = np.eye(4)
    M[:3, :3] = R @ np.diag(scale)
    M[:3, 3] = translate
    return np.ascontiguousarray(M.T, dtype=np.float32)  # rows of the array = columns of the matrix


def instanced_variant(fs, extras):
    """FlatScene copy with extra instances (source_instance, transform, material) appended, as
    solb_scene_add_instance does: same geometry, own transform and material."""
    import copy

    fs2 = copy.copy(fs)
    fs2.instances = list(fs.instances) + [dict(fs.instances[src], transform=np.asarray(t, dtype=np.float32), material=mat)
                                          for src, t, mat in extras]
    return fs2


def duck_extras(fs):
    """three more ducks: mirrored (det < 0), scaled + rotated, and one overlapping the original."""
    base = np.asarray(fs.instances[0]["transform"], dtype=np.float32).reshape(4, 4)

    def compose(m):
        return np.ascontiguousarray((m.T.astype(np.float64) @ base.T.astype(np.float64)).T, dtype=np.float32)  # world = m * base

    return [(0, compose(trs((2.0, 0.0, 0.5), (0, 1, 0), 0.7, (-1, 1, 1))), 0),
            (0, compose(trs((-1.5, 0.3, 1.0), (1, 0, 1), 1.1, (0.6, 0.6, 0.6))), 0),
            (0, compose(trs((0.2, 0.1, 0.0), (0, 0, 1), 0.2, (1, 1.3, 1))), 0)]


def write_glb(gltf_path, out_path):
    """Repack a .gltf (+ its external or data-URI buffers) as ONE .glb container: JSON chunk + a single BIN chunk
    holding all buffers back to back, bufferViews rebased onto buffer 0 (glTF 2.0 binary file format)."""
    import base64
    import json
    import struct

    with open(gltf_path) as f:
        doc = json.load(f)
    base = os.path.dirname(os.path.abspath(gltf_path))
    blob, offsets = b"", []
    for b in doc.get("buffers", []):
        uri = b["uri"]
        data = base64.b64decode(uri.split(",", 1)[1]) if uri.startswith("data:") else open(os.path.join(base, uri), "rb").read()
        offsets.append(len(blob))
        blob += data[: b["byteLength"]]
        blob += b"\0" * (-len(blob) % 4)
    for v in doc.get("bufferViews", []):
        v["byteOffset"] = v.get("byteOffset", 0) + offsets[v["buffer"]]
        v["buffer"] = 0
    doc["buffers"] = [{"byteLength": len(blob)}]
    js = json.dumps(doc, separators=(",", ":")).encode("utf-8")
    js += b" " * (-len(js) % 4)
    total = 12 + 8 + len(js) + 8 + len(blob)
    with open(out_path, "wb") as f:
        f.write(struct.pack("<4sII", b"glTF", 2, total))
        f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        f.write(struct.pack("<II", len(blob), 0x004E4942) + blob)
    return out_path


def write_instanced_gltf(gltf_path, out_path):
    """Copy of a .gltf whose mesh 0 is referenced by three more nodes: a sibling of its original node, a root node with a
    rotation + translation, and a grandchild under a scaled parent (glTF node-graph instancing)."""
    import json
    import shutil

    with open(gltf_path) as f:
        doc = json.load(f)
    base = os.path.dirname(os.path.abspath(gltf_path))
    for b in doc.get("buffers", []):
        if not b["uri"].startswith("data:"):
            shutil.copy(os.path.join(base, b["uri"]), os.path.join(os.path.dirname(out_path), b["uri"]))
    nodes = doc["nodes"]
    n0 = len(nodes)
    nodes.append({"mesh": 0, "translation": [150.0, 0.0, 40.0], "rotation": [0.0, 0.38268343, 0.0, 0.92387953]})   # n0: root-level
    nodes.append({"children": [n0 + 2], "scale": [0.5, 0.5, 0.5], "translation": [-120.0, 30.0, 0.0]})              # n0+1: scaled parent
    nodes.append({"children": [n0 + 3], "rotation": [0.70710678, 0.0, 0.0, 0.70710678]})                           # n0+2
    nodes.append({"mesh": 0, "translation": [0.0, 0.0, 200.0]})                                                    # n0+3: grandchild
    scene0 = doc["scenes"][doc.get("scene", 0)]
    # hang the new roots under the same parent chain as the original mesh node so they inherit its root scale
    first = next(i for i, n in enumerate(nodes) if n.get("mesh", None) == 0)
    parent = next((i for i, n in enumerate(nodes) if first in n.get("children", [])), None)
    if parent is None:
        scene0["nodes"] += [n0, n0 + 1]
    else:
        nodes[parent]["children"] += [n0, n0 + 1]
    with open(out_path, "w") as f:
        json.dump(doc, f)
    return out_path


def flat_from_product_scene(sc):
    """gltf_flatten.FlatScene (what the oracle consumes) from a sol_rs_b200.scene.Scene built in memory (synth.make_scene):
    meshes x sections -> instances in the reference's order (src/ray/mod.rs:78-134), buffers concatenated."""
    fs = gf.FlatScene()
    verts, inds = [], []
    nv, ni = 0, 0
    for mi, m in enumerate(sc.meshes):
        secs = []
        for s_ in m.primitive_sections:
            secs.append(dict(first_vertex=nv + s_.first_vertex, n_vertices=s_.n_vertices, first_index=ni + s_.first_index,
                             n_indices=s_.n_indices, material=s_.material_index))
        t = np.asarray(m.transform, dtype=np.float32).reshape(4, 4)
        fs.meshes.append(dict(name=m.name, transform=t, sections=secs))
        verts.append(m.vertices)
        inds.append(m.indices)
        nv += m.vertices.shape[0]
        ni += m.indices.shape[0]
        for sec in secs:
            fs.instances.append(dict(mesh=mi, transform=t, **sec))
    fs.vertices = np.concatenate(verts, axis=0)
    fs.indices = np.concatenate(inds)
    fs.materials = np.ascontiguousarray(sc.materials, dtype=np.float32).reshape(-1, 12)
    return fs
