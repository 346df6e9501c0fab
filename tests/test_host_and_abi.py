"""CPU tier: the C-ABI library loads and exports every symbol include/solb.h declares; the C++ host mirror
(glTF loader, Camera, SceneUniforms) reproduces the oracle-side restatement bit for bit; the product refuses to
run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from oracle import camera as ocam
from oracle import gltf_flatten as gf

from helpers import ROOT, model_path

import __graft_entry__ as entry


@pytest.fixture(scope="module", autouse=True)
def built():
    entry.build()


def test_abi_exports_every_declared_symbol():
    from sol_rs_b200 import _native as N

    hdr = open(os.path.join(ROOT, "include", "solb.h")).read()
    declared = set(re.findall(r"SOLB_API\s+[\w\s\*]+?\b(solb_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(N.SYMBOLS), "bindings and header disagree: %s" % (declared ^ set(N.SYMBOLS))
    L = ctypes.CDLL(N.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), "libsolb.so does not export %s" % name
    assert N.lib().solb_version() == 0x000100


def test_pod_layouts_match_reference_byte_contracts():
    from sol_rs_b200 import _native as N

    # SURVEY Appendix B
    assert ctypes.sizeof(N.ModelVertex) == 64 and N.ModelVertex.color.offset == 16 and N.ModelVertex.normal.offset == 32
    assert ctypes.sizeof(N.MaterialInfo) == 48 and N.MaterialInfo.emissive.offset == 16 and N.MaterialInfo.metallic.offset == 32
    assert ctypes.sizeof(N.SceneInstance) == 144 and N.SceneInstance.transform.offset == 16 and N.SceneInstance.transform_it.offset == 80
    assert ctypes.sizeof(N.SceneUniforms) == 400 and N.SceneUniforms.view_inverse.offset == 128
    assert N.SceneUniforms.projection_inverse.offset == 256 and N.SceneUniforms.frame.offset == 384


def test_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sol_rs_b200 as sol

    with pytest.raises(sol.SolbError) as e:
        sol.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under sol_rs_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sol_rs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src and "oracle/" not in src.replace("oracle/_ref", ""), f


@pytest.mark.parametrize("name", ["cornell", "tunnel", "Duck"])
def test_cpp_loader_matches_oracle_loader(name):
    from sol_rs_b200 import scene

    s = scene.load_scene(None, model_path(name))
    fs = gf.load_scene(model_path(name))
    assert len(s.meshes) == len(fs.meshes)
    v = np.concatenate([m.vertices for m in s.meshes])
    i = np.concatenate([m.indices for m in s.meshes])
    assert np.array_equal(v, fs.vertices) and np.array_equal(i, fs.indices)
    assert np.array_equal(s.materials, fs.materials)
    k = 0
    for m, fm in zip(s.meshes, fs.meshes):
        assert np.array_equal(m.transform, fm["transform"].reshape(16))
        assert len(m.primitive_sections) == len(fm["sections"])
        first_v = fm["sections"][0]["first_vertex"]
        for ps, fsec in zip(m.primitive_sections, fm["sections"]):
            assert ps.n_vertices == fsec["n_vertices"] and ps.n_indices == fsec["n_indices"]
            assert ps.first_vertex == fsec["first_vertex"] - first_v  # mesh-relative on the product side
            assert ps.material_index == fsec["material"]
            k += 1
    assert k == len(fs.instances)
    assert (s.camera is not None) == (fs.camera is not None)


@pytest.mark.parametrize("name,w,h", [("cornell", 512, 512), ("tunnel", 1920, 1080), ("tunnel", 3840, 2160)])
def test_cpp_camera_uniforms_match_oracle(name, w, h):
    from sol_rs_b200 import scene

    s = scene.load_scene(None, model_path(name))
    fs = gf.load_scene(model_path(name))
    s.camera.set_window_size((w, h))
    oc = ocam.Camera.from_view(fs.camera["view"], fs.camera["yfov"], fs.camera["znear"], fs.camera["zfar"])
    oc.set_window_size((w, h))
    for frame in (0, 7, 511):
        assert bytes(scene.scene_uniforms(s.camera, w, h, frame)) == ocam.scene_uniforms(oc, w, h, frame)


def test_cpp_look_at_camera_matches_oracle():
    from sol_rs_b200 import scene

    for eye, center in (((5, 5, 5), (0, 0, 0)), ((4, 1, 4), (0, 0.5, 0))):  # 3-ray-debug / 4-ray-ao cameras
        c = scene.Camera((900, 600))
        c.look_at(eye, center, (0, -1, 0))
        o = ocam.Camera((900, 600))
        o.look_at(eye, center, (0, -1, 0))
        assert bytes(scene.scene_uniforms(c, 900, 600, 3)) == ocam.scene_uniforms(o, 900, 600, 3)
        c.set_vfov(50.0)
        o.set_vfov(50.0)
        np.testing.assert_array_equal(c.perspective_matrix(), o.persp.reshape(16))


def test_loader_error_behaviour(tmp_path):
    """The reference unwrap()s on a missing / malformed file (src/scene/mod.rs:140): the mirror raises."""
    from sol_rs_b200 import scene
    import sol_rs_b200 as sol

    with pytest.raises(sol.SolbError):
        scene.load_scene(None, str(tmp_path / "missing.gltf"))
    bad = tmp_path / "bad.gltf"
    bad.write_text("{ not json")
    with pytest.raises(sol.SolbError):
        scene.load_scene(None, str(bad))
    # minimal valid document with no meshes / cameras
    ok = tmp_path / "empty.gltf"
    ok.write_text('{"asset": {"version": "2.0"}}')
    s = scene.load_scene(None, str(ok))
    assert s.meshes == [] and s.camera is None and s.materials.shape == (0, 12)


def test_find_asset():
    from sol_rs_b200 import util

    assert util.find_asset("models/cornell.gltf").endswith("assets/models/cornell.gltf")
    assert util.find_asset("models/ToyCar.glb") is None  # missing upstream too (.MISSING_LARGE_BLOBS)


def test_pipeline_selects_kernel_family():
    from sol_rs_b200 import ray
    import sol_rs_b200 as sol

    mk = lambda stem, spec=None: ray.Pipeline(None, (ray.PipelineInfo().shader("glsl/%s.rgen" % stem, ray.RAYGEN_KHR)
                                                     .shader("glsl/%s.rmiss" % stem, ray.MISS_KHR)
                                                     .shader("glsl/%s.rchit" % stem, ray.CLOSEST_HIT_KHR)
                                                     .specialization(spec or [0], 0)))
    assert mk("pathtrace").kind == ray.PATHTRACE and not mk("pathtrace").enable_sky
    assert mk("pathtrace", [1]).enable_sky
    assert mk("ao").kind == ray.AO and mk("debug").kind == ray.DEBUG
    with pytest.raises(sol.SolbError):
        mk("cube")
    with pytest.raises(sol.SolbError):
        ray.ShaderBindingTable(None, mk("debug"), ray.ShaderBindingTableInfo().raygen(0).miss(1))


def test_png_writer_roundtrip(tmp_path):
    """Offscreen image writer (replaces blit-to-present): decodes back to the same pixels."""
    import cv2
    from sol_rs_b200 import io

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(37, 53, 4), dtype=np.uint8)
    p = str(tmp_path / "f.png")
    io.write_png(p, img)
    back = cv2.imread(p, cv2.IMREAD_UNCHANGED)
    assert back is not None and back.shape == (37, 53, 4)
    assert np.array_equal(back[..., [2, 1, 0, 3]], img)
    io.write_ppm(str(tmp_path / "f.ppm"), img)
    raw = open(str(tmp_path / "f.ppm"), "rb").read()
    assert raw.startswith(b"P6\n53 37\n255\n") and len(raw) == 13 + 37 * 53 * 3


@pytest.mark.parametrize("name", ["cornell", "tunnel", "Duck"])
def test_glb_container_loads_like_its_gltf(name, tmp_path):
    """the .glb container (the form the reference's 4-ray-ao default asset ToyCar.glb ships in,
    examples/4-ray-ao.rs:76) goes through both loaders and yields exactly the scene of the equivalent .gltf."""
    from helpers import write_glb
    from sol_rs_b200 import scene

    glb = write_glb(model_path(name), str(tmp_path / (name + ".glb")))
    a, b = gf.load_scene(model_path(name)), gf.load_scene(glb)
    assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.indices, b.indices) and np.array_equal(a.materials, b.materials)
    assert len(a.instances) == len(b.instances) and all(np.array_equal(x["transform"], y["transform"]) for x, y in zip(a.instances, b.instances))
    s = scene.load_scene(None, glb)
    assert np.array_equal(np.concatenate([m.vertices for m in s.meshes]), a.vertices)
    assert np.array_equal(np.concatenate([m.indices for m in s.meshes]), a.indices)
    assert np.array_equal(s.materials, a.materials) and (s.camera is not None) == (a.camera is not None)
    # malformed containers fail like the reference's gltf::import(...).unwrap(): an error, never a crash
    raw = open(glb, "rb").read()
    bad = tmp_path / "bad.glb"
    bad.write_bytes(raw[:20] + b"\xff\xff\xff\x7f" + raw[24:])  # JSON chunk type corrupted
    with pytest.raises(Exception):
        scene.load_scene(None, str(bad))
    bad.write_bytes(raw[: len(raw) // 2])                        # truncated
    with pytest.raises(Exception):
        scene.load_scene(None, str(bad))


def test_node_graph_instancing_loader(tmp_path):
    """a mesh referenced by several glTF nodes.  The reference keeps the first node's transform only
    (src/scene/mod.rs:106-136, App.A item 5) — unchanged; the further nodes' global transforms are exposed separately and
    agree between the C++ loader and the oracle-side flatten."""
    from helpers import write_instanced_gltf
    from sol_rs_b200 import scene

    path = write_instanced_gltf(model_path("Duck"), str(tmp_path / "Duck_inst.gltf"))
    ref, ref_plain = gf.load_scene(path), gf.load_scene(model_path("Duck"))
    assert len(ref.instances) == 1 and np.array_equal(ref.instances[0]["transform"], ref_plain.instances[0]["transform"])
    inst = gf.load_scene(path, instancing=True)
    assert len(inst.instances) == 3 and np.array_equal(inst.instances[0]["transform"], ref.instances[0]["transform"])
    s = scene.load_scene(None, path)
    assert len(s.meshes) == 1 and len(s.meshes[0].extra_instance_transforms) == 2
    assert np.array_equal(s.meshes[0].transform, ref.meshes[0]["transform"].reshape(16))
    for t, oi in zip(s.meshes[0].extra_instance_transforms, inst.instances[1:]):
        np.testing.assert_allclose(t, oi["transform"].reshape(16), rtol=0, atol=0)
    # independent point check of the grandchild: origin -> T(0,0,200) -> rotX(+90 deg): (0,-200,0) -> scale .5 ->
    # T(-120,30,0): (-120,-70,0) in the frame of the node the new roots were hung under
    import json

    doc = json.load(open(path))
    first = next(i for i, n in enumerate(doc["nodes"]) if n.get("mesh", None) == 0)
    par = next((i for i, n in enumerate(doc["nodes"]) if first in n.get("children", [])), None)
    assert par is not None and not any(par in n.get("children", []) for n in doc["nodes"]), "Duck: mesh node under one root node"
    P = gf.node_matrix(doc["nodes"][par]).astype(np.float64).T  # arrays are [col][row]
    want = P @ np.array([-120.0, -70.0, 0.0, 1.0])
    got = inst.instances[2]["transform"].astype(np.float64).T @ np.array([0.0, 0.0, 0.0, 1.0])
    np.testing.assert_allclose(got[:3], want[:3], rtol=1e-5, atol=1e-5)
    assert len(scene.load_scene(None, model_path("Duck")).meshes[0].extra_instance_transforms) == 0


def test_rust_crate_binds_every_declared_symbol_with_matching_arity():
    from sol_rs_b200 import _native as N

    """rust/sol is the reference-side binding of include/solb.h (the build image has no rustc, so the crate is checked
    textually): ffi.rs declares exactly the header's entry points, each with the header's number of parameters; the crate
    has the modules lib.rs names; the POD mirrors have the header's sizes."""
    rust = os.path.join(ROOT, "rust", "sol", "src")
    hdr = open(os.path.join(ROOT, "include", "solb.h")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decl = {m.group(1): m.group(2) for m in re.finditer(r"SOLB_API\s+[\w\s\*]+?\b(solb_\w+)\s*\(([^;]*?)\)\s*;", hdr_nc, flags=re.S)}
    assert len(decl) >= 45
    ffi = open(os.path.join(rust, "ffi.rs")).read()
    ffi_nc = re.sub(r"//.*", "", ffi)
    bound = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (solb_\w+)\s*\(([^)]*)\)", ffi_nc, flags=re.S)}
    assert set(bound) == set(decl), "ffi.rs and solb.h disagree: %s" % (set(bound) ^ set(decl))

    def arity(params):
        p = params.strip()
        return 0 if p in ("", "void") else len([x for x in p.split(",") if x.strip()])

    for name in decl:
        assert arity(decl[name]) == arity(bound[name]), name

    # struct sizes: add up the fields of the #[repr(C)] mirrors
    def rust_struct_bytes(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, ffi, flags=re.S).group(1)
        total = 0
        for ty in re.findall(r"pub \w+: ([^,\n]+),", body):
            m = re.match(r"\[(\w+); (\d+)\]", ty.strip())
            base, n = (m.group(1), int(m.group(2))) if m else (ty.strip(), 1)
            total += {"f32": 4, "u32": 4, "i32": 4, "u64": 8}[base] * n
        return total

    assert rust_struct_bytes("SolbModelVertex") == 64 and rust_struct_bytes("SolbMaterialInfo") == 48
    assert rust_struct_bytes("SolbSceneInstance") == 144 and rust_struct_bytes("SolbSceneUniforms") == 400
    assert rust_struct_bytes("SolbTraceParams") == ctypes.sizeof(N.TraceParams)
    assert rust_struct_bytes("SolbStats") == ctypes.sizeof(N.Stats) and rust_struct_bytes("SolbAccelInfo") == ctypes.sizeof(N.AccelInfo)
    # module layout: everything lib.rs declares exists, and the names the reference's examples import are defined
    lib = open(os.path.join(rust, "lib.rs")).read()
    mods = set(re.findall(r"^(?:pub )?mod (\w+);", lib, flags=re.M))
    for mod in mods:
        assert os.path.exists(os.path.join(rust, mod + ".rs")), mod
    ray_rs, scene_rs = open(os.path.join(rust, "ray.rs")).read(), open(os.path.join(rust, "scene.rs")).read()
    for item in ("pub struct SceneDescription", "pub struct PipelineInfo", "pub struct Pipeline", "pub struct ShaderBindingTableInfo",
                 "pub struct ShaderBindingTable", "pub fn from_scene(", "pub fn from_meshes(", "pub fn blas_transform(",
                 "pub fn blas_transforms(", "pub fn tlas_regenerate<", "pub fn update(", "pub fn cmd_trace_rays(",
                 "pub fn new(context: Arc<Context>, pipeline: &Pipeline, info: ShaderBindingTableInfo)"):
        assert item in ray_rs, item
    for item in ("pub fn load_scene(", "pub struct Scene", "pub struct Mesh", "pub struct PrimitiveSection", "pub struct Camera",
                 "pub fn from_view(", "pub fn look_at(", "pub fn set_window_size(", "pub fn view_matrix(", "pub fn perspective_matrix("):
        assert item in scene_rs, item
    # every `crate::x` path in the sources names a module of the crate or an item re-exported at its root
    for f in os.listdir(rust):
        for m in re.findall(r"crate::(\w+)", open(os.path.join(rust, f)).read()):
            assert m in mods or m in ("Context", "Image2d", "SceneUniforms", "HostBuffer", "Fence"), (f, m)


def test_loaders_handle_snorm_zero_filled_and_percent_encoded_uris(tmp_path):
    """Valid glTF the shipped assets do not exercise: normalised signed BYTE normals, an accessor without
    bufferView (= zeros), a percent-encoded buffer URI.  The C++ loader and the oracle's loader must agree byte for byte."""
    import json
    import struct

    from sol_rs_b200 import scene

    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32)
    nrm = np.array([[0, 0, 127, 0], [0, 127, 0, 0], [-128, 0, 0, 0]], dtype=np.int8)  # VEC3 of BYTE, stride 4
    idx = np.array([0, 1, 2], dtype=np.uint16)
    blob = pos.tobytes() + nrm.tobytes() + idx.tobytes() + b"\x00\x00"
    (tmp_path / "my mesh.bin").write_bytes(blob)
    doc = {"asset": {"version": "2.0"}, "buffers": [{"uri": "my%20mesh.bin", "byteLength": len(blob)}],
           "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 36}, {"buffer": 0, "byteOffset": 36, "byteLength": 12, "byteStride": 4},
                           {"buffer": 0, "byteOffset": 48, "byteLength": 6}],
           "accessors": [{"bufferView": 0, "componentType": 5126, "count": 3, "type": "VEC3", "min": [0, 0, 0], "max": [1, 1, 0]},
                         {"bufferView": 1, "componentType": 5120, "normalized": True, "count": 3, "type": "VEC3"},
                         {"bufferView": 2, "componentType": 5123, "count": 3, "type": "SCALAR"},
                         {"componentType": 5126, "count": 3, "type": "VEC2"}],  # no bufferView: zero-filled texture coordinates
           "materials": [{"pbrMetallicRoughness": {"baseColorFactor": [0.5, 0.5, 0.5, 1.0]}}],
           "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "NORMAL": 1, "TEXCOORD_0": 3}, "indices": 2, "material": 0}]}],
           "nodes": [{"mesh": 0}], "scenes": [{"nodes": [0]}], "scene": 0}
    path = tmp_path / "edge.gltf"
    path.write_text(json.dumps(doc))
    s = scene.load_scene(None, str(path))
    fs = gf.load_scene(str(path))
    v = s.meshes[0].vertices
    assert np.array_equal(v, fs.vertices) and np.array_equal(s.meshes[0].indices, fs.indices)
    np.testing.assert_array_equal(v[:, 8:11], np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0]], dtype=np.float32))  # -128 clamps to -1
    assert np.all(v[:, 12:14] == 0.0) and np.array_equal(v[:, 0:3], pos)
    assert struct.calcsize("f") == 4


def _encode_png(img, ctype, depth, filt, level, palette=None, trns=None):
    """minimal PNG writer for the decoder tests: img [h, w, channels] of uint8 / uint16 samples, one filter type for all rows"""
    import struct
    import zlib

    h, w, ch = img.shape
    raw = img.astype(">u2").tobytes() if depth == 16 else img.astype(np.uint8).tobytes()
    bpp = ch * depth // 8
    stride = w * bpp
    rows = np.frombuffer(raw, dtype=np.uint8).reshape(h, stride).astype(np.int32)
    out = bytearray()
    prev = np.zeros(stride, dtype=np.int32)
    for y in range(h):
        cur = rows[y]
        a = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]]) if stride > bpp else np.zeros(stride, np.int32)
        c = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]]) if stride > bpp else np.zeros(stride, np.int32)
        if filt == 0:
            pred = np.zeros(stride, np.int32)
        elif filt == 1:
            pred = a
        elif filt == 2:
            pred = prev
        elif filt == 3:
            pred = (a + prev) >> 1
        else:
            p = a + prev - c
            pa, pb, pc = np.abs(p - a), np.abs(p - prev), np.abs(p - c)
            pred = np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, prev, c))
        out.append(filt)
        out += ((cur - pred) & 255).astype(np.uint8).tobytes()
        prev = cur

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if palette is not None:
        png += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    if trns is not None:
        png += chunk(b"tRNS", trns.astype(np.uint8).tobytes())
    comp = zlib.compress(bytes(out), level)
    half = len(comp) // 2
    return png + chunk(b"IDAT", comp[:half]) + chunk(b"IDAT", comp[half:]) + chunk(b"IEND", b"")


def test_png_decoder_all_colour_types_filters_and_block_types():
    """sol::image::decode_png (csrc/host/png.cpp: own inflate) against the oracle loader's decoder (Python zlib) and against the
    pixels that went in: every colour type at 8 and 16 bits, every row filter, stored / fixed / dynamic deflate blocks, IDAT
    split over two chunks; plus the two PNGs the repo ships."""
    from oracle import gltf_flatten as gf
    from sol_rs_b200 import _native as N
    from sol_rs_b200 import scene

    rng = np.random.default_rng(3)
    w, h = 13, 7
    n = 0
    for ctype, ch in ((0, 1), (2, 3), (3, 1), (4, 2), (6, 4)):
        for depth in ((8,) if ctype == 3 else (8, 16)):
            for filt in range(5):
                for level in (0, 1, 9):
                    smooth = (np.add.outer(np.arange(h) * 9, np.arange(w) * 5)[..., None] + np.arange(ch) * 40) % (1 << depth)
                    img = smooth if level == 9 else rng.integers(0, 1 << depth, size=(h, w, ch))
                    pal = rng.integers(0, 256, size=(256, 3)) if ctype == 3 else None
                    trns = rng.integers(0, 256, size=(200,)) if ctype == 3 else None
                    if ctype == 3:
                        img = img % 256
                    data = _encode_png(img, ctype, depth, filt, level, pal, trns)
                    got = scene.decode_png(data)
                    assert np.array_equal(got, gf.decode_png(data)), (ctype, depth, filt, level)
                    s8 = ((img.astype(np.uint32) + 128) // 257 if depth == 16 else img).astype(np.uint8)
                    want = np.full((h, w, 4), 255, dtype=np.uint8)
                    if ctype == 0:
                        want[..., :3] = s8[..., :1]
                    elif ctype == 2:
                        want[..., :3] = s8
                    elif ctype == 4:
                        want[..., :3], want[..., 3] = s8[..., :1], s8[..., 1]
                    elif ctype == 6:
                        want[...] = s8
                    else:
                        want[..., :3] = pal[s8[..., 0]]
                        want[..., 3] = np.concatenate([trns, np.full(56, 255)])[s8[..., 0]]
                    assert np.array_equal(got, want), (ctype, depth, filt, level)
                    n += 1
    assert n == 135
    for f in (os.path.join(ROOT, "assets", "models", "DuckCM.png"), os.path.join(ROOT, "assets", "textures", "HDR_RGBA_0.png")):
        data = open(f, "rb").read()
        assert np.array_equal(scene.decode_png(data), gf.decode_png(data))
    with pytest.raises(N.SolbError):
        scene.decode_png(b"\x89PNG\r\n\x1a\n" + b"\x00" * 40)
    with pytest.raises(N.SolbError):
        scene.decode_png(data[: len(data) // 2])


def test_loaders_agree_on_base_colour_textures(tmp_path):
    """Duck.gltf names DuckCM.png as its material's baseColorTexture: both loaders decode the same pixels, wrap modes and
    material -> texture table; a material whose image is missing or not a PNG simply stays untextured (the reference loads no
    images at all); scenes without textures report none."""
    import json
    import shutil

    from oracle import gltf_flatten as gf
    from sol_rs_b200 import scene

    duck = os.path.join(ROOT, "assets", "models", "Duck.gltf")
    sc, fs = scene.load_scene(None, duck), gf.load_scene(duck)
    assert len(sc.textures) == len(fs.textures) == 1 and sc.material_textures == fs.material_textures == [0]
    assert np.array_equal(sc.textures[0].rgba8, fs.textures[0][0]) and (sc.textures[0].wrap_s, sc.textures[0].wrap_t) == fs.textures[0][1:]
    assert sc.textures[0].rgba8.shape == (512, 512, 4) and sc.textures[0].rgba8[..., 3].min() == 255
    for name in ("cornell", "tunnel"):
        p = os.path.join(ROOT, "assets", "models", name + ".gltf")
        a, b = scene.load_scene(None, p), gf.load_scene(p)
        assert a.textures == [] and b.textures == [] and all(t is None for t in a.material_textures + b.material_textures)
    # clamp / mirror samplers, an embedded (data URI) image, a second material sharing the texture, one with a broken image
    doc = json.load(open(duck))
    import base64
    doc["images"] = [{"uri": "data:image/png;base64," + base64.b64encode(open(os.path.join(ROOT, "assets", "models", "DuckCM.png"), "rb").read()).decode()},
                     {"uri": "missing.png"}]
    doc["samplers"] = [{"wrapS": 33071, "wrapT": 33648}]
    doc["textures"] = [{"sampler": 0, "source": 0}, {"source": 1}]
    doc["materials"] = [doc["materials"][0], {"pbrMetallicRoughness": {"baseColorTexture": {"index": 1}}},
                        {"pbrMetallicRoughness": {"baseColorTexture": {"index": 0}}}, {"pbrMetallicRoughness": {"baseColorTexture": {"index": 0, "texCoord": 1}}}]
    shutil.copy(os.path.join(ROOT, "assets", "models", "Duck0.bin"), tmp_path / "Duck0.bin")
    json.dump(doc, open(tmp_path / "d.gltf", "w"))
    a, b = scene.load_scene(None, str(tmp_path / "d.gltf")), gf.load_scene(str(tmp_path / "d.gltf"))
    assert a.material_textures == b.material_textures == [0, None, 0, None]
    assert len(a.textures) == len(b.textures) == 1 and np.array_equal(a.textures[0].rgba8, b.textures[0][0])
    assert (a.textures[0].wrap_s, a.textures[0].wrap_t) == b.textures[0][1:] == (33071, 33648)


def test_exr_writer_round_trip_and_independent_reader(tmp_path):
    """sol_rs_b200.io.write_exr (EXR dump of the float accumulation target): bit-exact round trip through the
    module's own reader and, where OpenCV was built with OpenEXR, through an independent decoder."""
    from sol_rs_b200 import io as sio

    a = np.random.default_rng(7).normal(scale=50.0, size=(9, 17, 4)).astype(np.float32)
    a[0, 0, :3] = (0.0, np.float32(1e-30), np.float32(6.5e4))
    for channels in ("RGB", "RGBA"):
        p = str(tmp_path / ("t_%s.exr" % channels))
        sio.write_exr(p, a, channels)
        b, names = sio.read_exr(p)
        assert names == sorted(channels)
        assert np.array_equal(b, a[..., ["RGBA".index(c) for c in names]])
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    try:
        import cv2
    except ImportError:
        return
    im = cv2.imread(str(tmp_path / "t_RGB.exr"), cv2.IMREAD_UNCHANGED)
    if im is not None:  # (None: this OpenCV build has no OpenEXR codec)
        assert np.array_equal(im, a[..., [2, 1, 0]])
