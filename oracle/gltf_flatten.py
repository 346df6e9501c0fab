"""glTF -> flat reference-layout arrays.  ORACLE SIDE (test infrastructure only).

Restates, in numpy f32, what the reference does between a .gltf file and the buffers its
ray-tracing shaders read:

* ``load_scene``                      /root/reference  src/scene/mod.rs:138-295
* ``find_mesh`` / global transform    src/scene/mod.rs:106-136
* ``SceneDescription::from_meshes``   src/ray/mod.rs:59-156   (one BLAS + one instance per section)
* ``ModelVertex`` / ``MaterialInfo``  src/scene/mesh.rs:9-14, src/scene/mod.rs:19-29

Out-of-tree dependency restated from its published behaviour: the ``gltf`` crate 1.0.0
(Cargo.lock:518-519) — ``Node::transform().matrix()`` (matrix verbatim, or T*R*S from the
decomposed form with the cgmath-style quaternion->matrix formula), accessor readers
(``into_u32``, ``into_f32``, ``into_rgba_f32``), material defaults.  PARITY UNPINNED: the
reference holds no fixture for any of this.

The product has its own, independent loader in C++ (sol_rs_b200/csrc/host/gltf.cpp); tests
compare the two.  Nothing in the product imports this module.
"""
import base64
import json
import os
import urllib.parse

import numpy as np

F32 = np.float32

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}


class Gltf:
    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            raw = f.read()
        glb_bin = None
        if raw[:4] == b"glTF":
            # GLB container (what gltf::import also accepts; ToyCar.glb of examples/4-ray-ao.rs:76): header
            # magic | version | length, then {u32 length, u32 type, data} chunks: JSON (0x4E4F534A), BIN (0x004E4942)
            import struct

            version, total = struct.unpack_from("<II", raw, 4)
            assert version == 2, "GLB container version must be 2"
            off, doc = 12, None
            while off + 8 <= min(total, len(raw)):
                clen, ctype = struct.unpack_from("<II", raw, off)
                chunk = raw[off + 8: off + 8 + clen]
                if ctype == 0x4E4F534A and doc is None:
                    doc = json.loads(chunk.decode("utf-8"))
                elif ctype == 0x004E4942 and glb_bin is None:
                    glb_bin = chunk
                off += 8 + ((clen + 3) & ~3)
            assert doc is not None, "GLB without a JSON chunk"
            self.doc = doc
        else:
            self.doc = json.loads(raw.decode("utf-8"))
        self.buffers = []
        base = os.path.dirname(os.path.abspath(path))
        for bi, b in enumerate(self.doc.get("buffers", [])):
            uri = b.get("uri")
            if uri is None:
                assert glb_bin is not None and bi == 0, "buffer without uri outside a GLB container"
                data = glb_bin
            elif uri.startswith("data:"):
                data = base64.b64decode(uri.split(",", 1)[1])
            else:
                with open(os.path.join(base, urllib.parse.unquote(uri)), "rb") as f:
                    data = f.read()
            self.buffers.append(data[: b["byteLength"]])

    def accessor(self, idx):
        """Accessor -> ndarray [count, ncomp] in its stored component type (no sparse support:
        no shipped asset uses it)."""
        acc = self.doc["accessors"][idx]
        dt = np.dtype(_COMPONENT[acc["componentType"]])
        nc = _NCOMP[acc["type"]]
        count = acc["count"]
        if "bufferView" not in acc:  # an accessor without bufferView is all zeros (glTF 2.0 5.1.1)
            return np.zeros((count, nc), dtype=dt), acc
        view = self.doc["bufferViews"][acc["bufferView"]]
        start = view.get("byteOffset", 0) + acc.get("byteOffset", 0)
        stride = view.get("byteStride", 0) or dt.itemsize * nc
        strided = np.ndarray(shape=(count, nc), dtype=dt, buffer=self.buffers[view["buffer"]], offset=start,
                             strides=(stride, dt.itemsize))  # honours byteStride
        return np.array(strided), acc


def _normalized_to_f32(arr):
    """gltf crate `into_f32` / `into_rgba_f32` casts: u8 -> x/255, u16 -> x/65535."""
    if arr.dtype == np.float32:
        return arr
    if arr.dtype == np.uint8:
        return (arr.astype(F32) / F32(255.0)).astype(F32)
    if arr.dtype == np.uint16:
        return (arr.astype(F32) / F32(65535.0)).astype(F32)
    if arr.dtype == np.int8:  # signed normalised (KHR_mesh_quantization): max(v / 127, -1)
        return np.maximum(arr.astype(F32) / F32(127.0), F32(-1.0)).astype(F32)
    if arr.dtype == np.int16:
        return np.maximum(arr.astype(F32) / F32(32767.0), F32(-1.0)).astype(F32)
    raise ValueError("unsupported normalized type %s" % arr.dtype)


def node_matrix(node):
    """gltf::scene::Transform::matrix() -> column-major 4x4 as a [4,4] array indexed [col][row].

    `matrix` property is taken verbatim; otherwise T * R * S with the crate's (cgmath-derived)
    quaternion formula evaluated in f32."""
    if "matrix" in node:
        return np.array(node["matrix"], dtype=F32).reshape(4, 4)
    t = [F32(v) for v in node.get("translation", [0.0, 0.0, 0.0])]
    r = [F32(v) for v in node.get("rotation", [0.0, 0.0, 0.0, 1.0])]
    s = [F32(v) for v in node.get("scale", [1.0, 1.0, 1.0])]
    x, y, z, w = r
    x2, y2, z2 = x + x, y + y, z + z
    xx2, xy2, xz2 = x2 * x, x2 * y, x2 * z
    yy2, yz2, zz2 = y2 * y, y2 * z, z2 * z
    sy2, sz2, sx2 = y2 * w, z2 * w, x2 * w
    one = F32(1.0)
    R = np.array(
        [
            [one - yy2 - zz2, xy2 + sz2, xz2 - sy2, 0],
            [xy2 - sz2, one - xx2 - zz2, yz2 + sx2, 0],
            [xz2 + sy2, yz2 - sx2, one - xx2 - yy2, 0],
            [0, 0, 0, 1],
        ],
        dtype=F32,
    )
    T = np.eye(4, dtype=F32)
    T[3, 0:3] = t
    S = np.diag(np.array([s[0], s[1], s[2], one], dtype=F32)).astype(F32)
    return mat4_mul(mat4_mul(T, R), S)


def mat4_mul(a, b):
    """Column-major product a*b with arrays indexed [col][row]; f32, terms summed in k order."""
    out = np.zeros((4, 4), dtype=F32)
    for c in range(4):
        for r in range(4):
            acc = F32(0.0)
            for k in range(4):
                acc = F32(acc + a[k, r] * b[c, k])
            out[c, r] = acc
    return out


def _find_mesh(doc, node_idx, transforms, mesh_index):
    """src/scene/mod.rs:106-122"""
    node = doc["nodes"][node_idx]
    transforms.append(node_matrix(node))
    if node.get("mesh", None) == mesh_index:
        return True
    for child in node.get("children", []):
        if _find_mesh(doc, child, transforms, mesh_index):
            return True
    transforms.pop()
    return False


def mesh_first_node(doc, mesh_index):
    """index of the node the reference's search (src/scene/mod.rs:106-136) stops at for this mesh, or -1"""
    def walk(i):
        node = doc["nodes"][i]
        if node.get("mesh", None) == mesh_index:
            return i
        for c in node.get("children", []):
            r = walk(c)
            if r >= 0:
                return r
        return -1

    for i in range(len(doc.get("nodes", []))):
        r = walk(i)
        if r >= 0:
            return r
    return -1


def mesh_global_transform(doc, mesh_index):
    """src/scene/mod.rs:124-136: scan nodes in index order, first subtree holding the mesh wins."""
    g = np.eye(4, dtype=F32)
    transforms = []
    for i in range(len(doc.get("nodes", []))):
        if _find_mesh(doc, i, transforms, mesh_index):
            for t in transforms:
                g = mat4_mul(g, t)
            break
    return g


def mat4_inverse(m):
    """glam 0.20.2 Mat4::inverse (scalar-math path, the GLM cofactor scheme), f32.
    m indexed [col][row].  Call sites: src/ray/mod.rs:29,116; examples/5-pathtrace.rs:24,26."""
    m = m.astype(F32)
    m00, m01, m02, m03 = m[0]
    m10, m11, m12, m13 = m[1]
    m20, m21, m22, m23 = m[2]
    m30, m31, m32, m33 = m[3]
    coef00 = m22 * m33 - m32 * m23
    coef02 = m12 * m33 - m32 * m13
    coef03 = m12 * m23 - m22 * m13
    coef04 = m21 * m33 - m31 * m23
    coef06 = m11 * m33 - m31 * m13
    coef07 = m11 * m23 - m21 * m13
    coef08 = m21 * m32 - m31 * m22
    coef10 = m11 * m32 - m31 * m12
    coef11 = m11 * m22 - m21 * m12
    coef12 = m20 * m33 - m30 * m23
    coef14 = m10 * m33 - m30 * m13
    coef15 = m10 * m23 - m20 * m13
    coef16 = m20 * m32 - m30 * m22
    coef18 = m10 * m32 - m30 * m12
    coef19 = m10 * m22 - m20 * m12
    coef20 = m20 * m31 - m30 * m21
    coef22 = m10 * m31 - m30 * m11
    coef23 = m10 * m21 - m20 * m11
    v = lambda a, b, c, d: np.array([a, b, c, d], dtype=F32)
    fac0 = v(coef00, coef00, coef02, coef03)
    fac1 = v(coef04, coef04, coef06, coef07)
    fac2 = v(coef08, coef08, coef10, coef11)
    fac3 = v(coef12, coef12, coef14, coef15)
    fac4 = v(coef16, coef16, coef18, coef19)
    fac5 = v(coef20, coef20, coef22, coef23)
    vec0 = v(m10, m00, m00, m00)
    vec1 = v(m11, m01, m01, m01)
    vec2 = v(m12, m02, m02, m02)
    vec3 = v(m13, m03, m03, m03)
    inv0 = vec1 * fac0 - vec2 * fac1 + vec3 * fac2
    inv1 = vec0 * fac0 - vec2 * fac3 + vec3 * fac4
    inv2 = vec0 * fac1 - vec1 * fac3 + vec3 * fac5
    inv3 = vec0 * fac2 - vec1 * fac4 + vec2 * fac5
    sign_a = v(1, -1, 1, -1)
    sign_b = v(-1, 1, -1, 1)
    inverse = np.stack([inv0 * sign_a, inv1 * sign_b, inv2 * sign_a, inv3 * sign_b]).astype(F32)
    col0 = v(inverse[0, 0], inverse[1, 0], inverse[2, 0], inverse[3, 0])
    dot0 = m[0] * col0
    dot1 = F32(F32(F32(dot0[0] + dot0[1]) + dot0[2]) + dot0[3])
    rcp_det = F32(1.0) / dot1
    return (inverse * rcp_det).astype(F32)


class FlatScene:
    """What the RT shaders see: per-instance sections over concatenated reference-layout buffers."""

    def __init__(self):
        self.vertices = np.zeros((0, 16), dtype=F32)  # ModelVertex: pos, color, normal, uv
        self.indices = np.zeros((0,), dtype=np.uint32)  # section-relative
        self.materials = np.zeros((0, 12), dtype=F32)  # MaterialInfo
        self.instances = []  # dicts: mesh, first_vertex, n_vertices, first_index, n_indices, material, transform
        self.meshes = []  # dicts: name, transform, sections
        self.camera = None  # dict(view=[4,4], yfov, znear, zfar) or None
        # EXTENSION shared with the product (the reference loads no images): base-colour textures of the
        # materials, (rgba8 [h, w, 4] rows top first, wrap_s, wrap_t), and per material an index into them or None
        self.textures = []
        self.material_textures = []


def decode_png(data):
    """PNG -> uint8 [h, w, 4]: 8 / 16-bit grey, rgb, palette (8-bit), grey-alpha, rgba; non-interlaced.  zlib does the inflate;
    16-bit samples -> 8-bit by (v + 128) // 257 (the image 0.24 crate's to_rgba8 rule)."""
    import struct
    import zlib

    assert data[:8] == b"\x89PNG\r\n\x1a\n", "not a PNG"
    off, idat, plte, trns, hdr = 8, b"", b"", b"", None
    while off + 12 <= len(data):
        n, tag = struct.unpack_from(">I4s", data, off)
        body = data[off + 8: off + 8 + n]
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body[:13])
        elif tag == b"PLTE":
            plte = body
        elif tag == b"tRNS":
            trns = body
        elif tag == b"IDAT":
            idat += body
        elif tag == b"IEND":
            break
        off += 12 + n
    w, h, depth, ctype, _, _, interlace = hdr
    assert interlace == 0 and (depth == 8 or (depth == 16 and ctype != 3))
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    bpp = ch * depth // 8
    stride = w * bpp
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8)[: (stride + 1) * h].reshape(h, stride + 1)
    rows = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int32)
    for y in range(h):
        f, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if f == 0:
            cur = line
        elif f == 2:
            cur = (line + prev) & 255
        else:  # filters that look left: byte-serial
            cur = np.zeros(stride, dtype=np.int32)
            for i in range(stride):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if f == 1:
                    p = a
                elif f == 3:
                    p = (a + b) >> 1
                else:
                    q = a + b - c
                    pa, pb, pc = abs(q - a), abs(q - b), abs(q - c)
                    p = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (line[i] + p) & 255
        rows[y] = cur
        prev = cur
    if depth == 16:
        v = rows.reshape(h, w, ch, 2).astype(np.uint32)
        s8 = (((v[..., 0] << 8) | v[..., 1]) + 128) // 257
    else:
        s8 = rows.reshape(h, w, ch).astype(np.uint32)
    out = np.full((h, w, 4), 255, dtype=np.uint8)
    if ctype == 0:
        out[..., :3] = s8[..., :1]
    elif ctype == 2:
        out[..., :3] = s8
    elif ctype == 4:
        out[..., :3] = s8[..., :1]
        out[..., 3] = s8[..., 1]
    elif ctype == 6:
        out[...] = s8
    else:
        pal = np.frombuffer(plte, dtype=np.uint8).reshape(-1, 3)
        out[..., :3] = pal[s8[..., 0]]
        alpha = np.full(256, 255, dtype=np.uint8)
        alpha[: len(trns)] = np.frombuffer(trns, dtype=np.uint8)
        out[..., 3] = alpha[s8[..., 0]]
    return out


def _load_textures(g, fs):
    doc = g.doc
    base = os.path.dirname(os.path.abspath(g.path))
    slot = {}
    for mat in doc.get("materials", []):
        ref = mat.get("pbrMetallicRoughness", {}).get("baseColorTexture")
        bound = None
        if ref is not None and ref.get("texCoord", 0) == 0 and 0 <= ref.get("index", 0) < len(doc.get("textures", [])):
            ti = ref.get("index", 0)
            if ti not in slot:
                slot[ti] = None
                try:
                    tex = doc["textures"][ti]
                    img = doc["images"][tex["source"]]
                    if "uri" in img:
                        uri = img["uri"]
                        if uri.startswith("data:"):
                            data = base64.b64decode(uri.split(",", 1)[1])
                        else:
                            with open(os.path.join(base, urllib.parse.unquote(uri)), "rb") as f:
                                data = f.read()
                    else:
                        view = doc["bufferViews"][img["bufferView"]]
                        o = view.get("byteOffset", 0)
                        data = bytes(g.buffers[view["buffer"]][o: o + view["byteLength"]])
                    rgba = decode_png(data)
                    smp = doc.get("samplers", [])[tex["sampler"]] if "sampler" in tex else {}
                    slot[ti] = len(fs.textures)
                    fs.textures.append((rgba, smp.get("wrapS", 10497), smp.get("wrapT", 10497)))
                except Exception:
                    pass  # not a PNG / missing file: the material stays untextured
            bound = slot[ti]
        fs.material_textures.append(bound)


def _other_node_transforms(doc, mesh_index, first_node):
    """glTF node-graph instancing (beyond the reference): the global transform root -> node of every node that
    references the mesh except the one the reference's own search stopped at, node-index order."""
    nodes = doc.get("nodes", [])
    parent = [-1] * len(nodes)
    for i, n in enumerate(nodes):
        for c in n.get("children", []):
            if 0 <= c < len(nodes):
                parent[c] = i
    out = []
    for i, n in enumerate(nodes):
        if i == first_node or n.get("mesh", None) != mesh_index:
            continue
        chain, k = [], i
        while k >= 0 and len(chain) <= len(nodes):
            chain.append(k)
            k = parent[k]
        gm = np.eye(4, dtype=F32)
        for k in reversed(chain):
            gm = mat4_mul(gm, node_matrix(nodes[k]))
        out.append(gm)
    return out


def load_scene(path, instancing=False):
    """instancing=True appends, after the reference's instances, one instance per (further node of a mesh) x (section)."""
    g = Gltf(path)
    doc = g.doc
    fs = FlatScene()
    mats = []
    for mat in doc.get("materials", []):
        pbr = mat.get("pbrMetallicRoughness", {})
        base = pbr.get("baseColorFactor", [1.0, 1.0, 1.0, 1.0])
        em = mat.get("emissiveFactor", [0.0, 0.0, 0.0])
        mats.append(
            [base[0], base[1], base[2], base[3], em[0], em[1], em[2], 0.0,
             pbr.get("metallicFactor", 1.0), pbr.get("roughnessFactor", 1.0), 0.0, 0.0]
        )
    fs.materials = np.array(mats, dtype=F32).reshape(-1, 12)
    _load_textures(g, fs)
    verts, inds = [], []
    nv_total, ni_total = 0, 0
    for mi, mesh in enumerate(doc.get("meshes", [])):
        sections = []
        transform = mesh_global_transform(doc, mi)
        for prim in mesh["primitives"]:
            attrs = prim["attributes"]
            n = 0
            if "POSITION" in attrs:
                pos, _ = g.accessor(attrs["POSITION"])
                n = pos.shape[0]
                v = np.zeros((n, 16), dtype=F32)
                v[:, 0:3] = pos
                v[:, 3] = 1.0
                v[:, 4:8] = 1.0  # colour default (1,1,1,1): src/scene/mod.rs:186
                if "COLOR_0" in attrs:
                    col, _ = g.accessor(attrs["COLOR_0"])
                    col = _normalized_to_f32(col)
                    k = min(n, col.shape[0])
                    v[:k, 4 : 4 + col.shape[1]] = col[:k]  # vec3 -> alpha stays 1 (into_rgba_f32)
                v[:, 8:12] = [0.0, 1.0, 0.0, 1.0]  # normal default (0,1,0), w = 1: mod.rs:184,190
                if "NORMAL" in attrs:
                    nrm, _ = g.accessor(attrs["NORMAL"])
                    nrm = _normalized_to_f32(nrm)
                    k = min(n, nrm.shape[0])
                    v[:k, 8:11] = nrm[:k]
                if "TEXCOORD_0" in attrs:
                    uv, _ = g.accessor(attrs["TEXCOORD_0"])
                    uv = _normalized_to_f32(uv)
                    k = min(n, uv.shape[0])
                    v[:k, 12:14] = uv[:k]
                verts.append(v)
            sec = dict(first_vertex=nv_total, n_vertices=n, first_index=None, n_indices=0,
                       material=prim.get("material", None))
            nv_total += n
            if "indices" in prim:
                idx, _ = g.accessor(prim["indices"])
                idx = idx[:, 0].astype(np.uint32)
                sec["first_index"] = ni_total
                sec["n_indices"] = int(idx.shape[0])
                ni_total += int(idx.shape[0])
                inds.append(idx)
            sections.append(sec)
        fs.meshes.append(dict(name=mesh.get("name", ""), transform=transform, sections=sections))
    fs.vertices = np.concatenate(verts, axis=0) if verts else np.zeros((0, 16), dtype=F32)
    fs.indices = np.concatenate(inds) if inds else np.zeros((0,), dtype=np.uint32)
    # src/ray/mod.rs:78-134: instance id = running count over meshes x sections
    for mi, mesh in enumerate(fs.meshes):
        for sec in mesh["sections"]:
            if sec["first_index"] is None:
                raise ValueError("non-indexed primitive: unsupported on the RT path (SURVEY App.A 4)")
            if sec["material"] is None:
                raise ValueError("primitive without material: the reference unwrap()s (src/scene/mod.rs:65)")
            fs.instances.append(dict(mesh=mi, transform=mesh["transform"], **sec))
    if instancing:
        for mi, mesh in enumerate(fs.meshes):
            for t in _other_node_transforms(doc, mi, mesh_first_node(doc, mi)):
                for sec in mesh["sections"]:
                    fs.instances.append(dict(mesh=mi, transform=t, **sec))
    # camera: first camera only, perspective only (src/scene/mod.rs:261-287)
    cams = doc.get("cameras", [])
    if cams and cams[0].get("type") == "perspective":
        persp = cams[0]["perspective"]
        for node in doc.get("nodes", []):
            if node.get("camera", None) == 0:
                fs.camera = dict(view=node_matrix(node), yfov=F32(persp["yfov"]), znear=F32(persp["znear"]),
                                 zfar=F32(persp.get("zfar", 100.0)))
                break
    return fs
