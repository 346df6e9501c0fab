// gltf.cpp — sol::scene::load_scene: glTF 2.0 (.gltf + external/embedded buffers) -> Scene.
//
// Mirrors src/scene/mod.rs:106-295 of the reference, including its quirks (SURVEY App.A 1-6):
//   * one Mesh per glTF mesh; all primitives concatenated, each a PrimitiveSection; indices stay
//     relative to the section's own vertex range (mod.rs:163-216)
//   * ModelVertex = {pos.xyz,1 ; color rgba (default 1) ; normal.xyz,1 (default 0,1,0) ; uv,0,0} (mod.rs:182-193)
//   * mesh transform = product of node matrices root->node of the FIRST node (index order) whose subtree
//     holds the mesh (mod.rs:106-136)
//   * camera: first camera only, perspective only; the node's LOCAL matrix is stored as the view
//     matrix; yfov passed through unchanged (Camera applies to_radians()) (mod.rs:261-287)
// The `gltf` crate (1.0.0, Cargo.lock:518-519) behaviour it relies on is restated: material
// defaults, Node::transform().matrix() for decomposed TRS, accessor readers into_u32 / into_f32 /
// into_rgba_f32.  .glb containers are accepted (JSON + BIN chunks).  Unsupported (no shipped asset uses them): sparse
// accessors, Draco.  Beyond the reference: the materials' base-colour textures are decoded (PNG) into
// Scene::textures / material_textures; the reference loads none on this path (texture_offset is reserved, src/ray/mod.rs:20),
// and nothing changes for a caller that does not bind them.
#include <cstdio>
#include <fstream>
#include <sstream>

#include "json.hpp"
#include "sol.hpp"
#include <algorithm>

namespace sol {
namespace scene {

namespace {

std::string read_file(const std::string &path, bool binary) {
    std::ifstream f(path, binary ? std::ios::binary : std::ios::in);
    if (!f) throw Error(SOLB_ERR_INVALID, "load_scene: cannot open '" + path + "'");
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

std::string base64_decode(const std::string &in, size_t begin) {
    static int8_t lut[256];
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 256; i++) lut[i] = -1;
        const char *abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; i++) lut[(unsigned char)abc[i]] = (int8_t)i;
        init = true;
    }
    std::string out;
    out.reserve((in.size() - begin) * 3 / 4);
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = begin; i < in.size(); i++) {
        const int8_t v = lut[(unsigned char)in[i]];
        if (v < 0) continue;  // '=', whitespace
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((char)((acc >> bits) & 0xff)); }
    }
    return out;
}

struct Doc {
    json::Value root;
    std::vector<std::string> buffers;
};

struct AccessorView {
    const uint8_t *base = nullptr;  // null: an accessor without bufferView = all zeros (glTF 2.0 5.1.1)
    size_t count = 0, stride = 0;
    int component = 0, ncomp = 0;
};

// RFC 3986 percent-decoding of a relative buffer URI ("my%20mesh.bin")
std::string uri_decode(const std::string &s) {
    auto hex = [](char c) { return c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'f' ? c - 'a' + 10 : (c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1)); };
    std::string out;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '%' && i + 2 < s.size() + 0 && hex(s[i + 1]) >= 0 && hex(s[i + 2]) >= 0) {
            out.push_back((char)(hex(s[i + 1]) * 16 + hex(s[i + 2])));
            i += 2;
        } else out.push_back(s[i]);
    }
    return out;
}

int comp_size(int c) { return (c == 5120 || c == 5121) ? 1 : ((c == 5122 || c == 5123) ? 2 : 4); }
int type_ncomp(const std::string &t) {
    if (t == "SCALAR") return 1;
    if (t == "VEC2") return 2;
    if (t == "VEC3") return 3;
    if (t == "VEC4") return 4;
    throw Error(SOLB_ERR_UNSUPPORTED, "load_scene: accessor type '" + t + "'");
}

AccessorView accessor(const Doc &d, size_t index) {
    const json::Value &acc = d.root["accessors"][index];
    if (acc.has("sparse")) throw Error(SOLB_ERR_UNSUPPORTED, "load_scene: sparse accessors are not supported");
    AccessorView a;
    a.component = (int)acc["componentType"].integer(0);
    a.ncomp = type_ncomp(acc["type"].string());
    if (!acc.has("bufferView")) {  // zero-filled accessor: no storage to point at
        a.count = (size_t)acc["count"].integer(0);
        a.stride = (size_t)comp_size(a.component) * a.ncomp;
        return a;
    }
    const json::Value &view = d.root["bufferViews"][(size_t)acc["bufferView"].integer(0)];
    a.count = (size_t)acc["count"].integer(0);
    const size_t elem = (size_t)comp_size(a.component) * a.ncomp;
    const size_t bs = (size_t)view["byteStride"].integer(0);
    a.stride = bs ? bs : elem;
    const std::string &buf = d.buffers.at((size_t)view["buffer"].integer(0));
    const size_t off = (size_t)view["byteOffset"].integer(0) + (size_t)acc["byteOffset"].integer(0);
    if (a.count && off + (a.count - 1) * a.stride + elem > buf.size()) throw Error(SOLB_ERR_INVALID, "load_scene: accessor outside its buffer");
    a.base = (const uint8_t *)buf.data() + off;
    return a;
}

float read_f32(const AccessorView &a, size_t i, int c) {  // into_f32 / into_rgba_f32 casts
    if (!a.base) return 0.0f;
    const uint8_t *p = a.base + i * a.stride + (size_t)c * comp_size(a.component);
    switch (a.component) {
        case 5126: { float f; std::memcpy(&f, p, 4); return f; }
        case 5121: return (float)p[0] / 255.0f;
        case 5123: { uint16_t v; std::memcpy(&v, p, 2); return (float)v / 65535.0f; }
        // signed normalised components (KHR_mesh_quantization normals / texture coordinates): max(v / 127, -1), max(v / 32767, -1)
        case 5120: return std::max((float)(int8_t)p[0] / 127.0f, -1.0f);
        case 5122: { int16_t v; std::memcpy(&v, p, 2); return std::max((float)v / 32767.0f, -1.0f); }
        default: throw Error(SOLB_ERR_UNSUPPORTED, "load_scene: unsupported float attribute component type");
    }
}
uint32_t read_u32(const AccessorView &a, size_t i) {  // into_u32
    if (!a.base) return 0u;
    const uint8_t *p = a.base + i * a.stride;
    switch (a.component) {
        case 5121: return p[0];
        case 5123: { uint16_t v; std::memcpy(&v, p, 2); return v; }
        case 5125: { uint32_t v; std::memcpy(&v, p, 4); return v; }
        default: throw Error(SOLB_ERR_UNSUPPORTED, "load_scene: unsupported index component type");
    }
}

// gltf::scene::Transform::matrix(): `matrix` verbatim, else T * R * S (cgmath-style quaternion formula), f32
Mat4 node_matrix(const json::Value &node) {
    if (node.has("matrix")) {
        Mat4 m{};
        for (size_t i = 0; i < 16; i++) m[i] = (float)node["matrix"][i].number(0.0);
        return m;
    }
    float t[3] = { 0, 0, 0 }, r[4] = { 0, 0, 0, 1 }, s[3] = { 1, 1, 1 };
    if (node.has("translation")) for (size_t i = 0; i < 3; i++) t[i] = (float)node["translation"][i].number(0.0);
    if (node.has("rotation")) for (size_t i = 0; i < 4; i++) r[i] = (float)node["rotation"][i].number(0.0);
    if (node.has("scale")) for (size_t i = 0; i < 3; i++) s[i] = (float)node["scale"][i].number(1.0);
    const float x = r[0], y = r[1], z = r[2], w = r[3];
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float xx2 = x2 * x, xy2 = x2 * y, xz2 = x2 * z, yy2 = y2 * y, yz2 = y2 * z, zz2 = z2 * z;
    const float sy2 = y2 * w, sz2 = z2 * w, sx2 = x2 * w;
    Mat4 R = { 1.0f - yy2 - zz2, xy2 + sz2, xz2 - sy2, 0, xy2 - sz2, 1.0f - xx2 - zz2, yz2 + sx2, 0,
               xz2 + sy2, yz2 - sx2, 1.0f - xx2 - yy2, 0, 0, 0, 0, 1 };
    Mat4 T = math::identity();
    T[12] = t[0]; T[13] = t[1]; T[14] = t[2];
    const Mat4 S = math::from_scale({ s[0], s[1], s[2] });
    return math::mul(math::mul(T, R), S);
}

bool find_mesh(const json::Value &nodes, size_t node_index, std::vector<Mat4> &transforms, long mesh_index, long *found_node) {  // mod.rs:106-122
    const json::Value &node = nodes[node_index];
    transforms.push_back(node_matrix(node));
    if (node.has("mesh") && node["mesh"].integer(-1) == mesh_index) { *found_node = (long)node_index; return true; }
    const json::Value &children = node["children"];
    for (size_t i = 0; i < children.size(); i++)
        if (find_mesh(nodes, (size_t)children[i].integer(0), transforms, mesh_index, found_node)) return true;
    transforms.pop_back();
    return false;
}

Mat4 calc_mesh_global_transform(const Doc &d, long mesh_index, long *found_node) {  // mod.rs:124-136
    Mat4 g = math::identity();
    std::vector<Mat4> transforms;
    const json::Value &nodes = d.root["nodes"];
    *found_node = -1;
    for (size_t i = 0; i < nodes.size(); i++) {
        if (find_mesh(nodes, i, transforms, mesh_index, found_node)) {
            for (const Mat4 &t : transforms) g = math::mul(g, t);
            break;
        }
    }
    return g;
}

// glTF node-graph instancing (beyond the reference): the true global transform (root -> node) of every node that
// references the mesh, except the node the reference's own search stopped at
std::vector<Mat4> other_node_transforms(const Doc &d, long mesh_index, long first_node) {
    const json::Value &nodes = d.root["nodes"];
    std::vector<long> parent(nodes.size(), -1);
    for (size_t i = 0; i < nodes.size(); i++) {
        const json::Value &children = nodes[i]["children"];
        for (size_t k = 0; k < children.size(); k++) {
            const long c = children[k].integer(-1);
            if (c >= 0 && (size_t)c < nodes.size()) parent[(size_t)c] = (long)i;
        }
    }
    std::vector<Mat4> out;
    for (size_t i = 0; i < nodes.size(); i++) {
        if ((long)i == first_node || !nodes[i].has("mesh") || nodes[i]["mesh"].integer(-1) != mesh_index) continue;
        std::vector<long> chain;
        for (long n = (long)i; n >= 0 && chain.size() <= nodes.size(); n = parent[(size_t)n]) chain.push_back(n);
        Mat4 g = math::identity();
        for (size_t k = chain.size(); k-- > 0;) g = math::mul(g, node_matrix(nodes[(size_t)chain[k]]));
        out.push_back(g);
    }
    return out;
}

}  // namespace

// GLB container (glTF 2.0 binary, the form ToyCar.glb of examples/4-ray-ao.rs:76 ships in; gltf::import of the gltf crate
// accepts both): 12-byte header "glTF" | version 2 | total length, then chunks {u32 length, u32 type, data}: the first is
// JSON (0x4E4F534A), the optional second is BIN (0x004E4942) = the buffer without a uri.
static bool split_glb(const std::string &file, std::string &json_out, std::string &bin_out) {
    auto u32 = [&](size_t off) {
        uint32_t v;
        std::memcpy(&v, file.data() + off, 4);
        return v;
    };
    if (file.size() < 12 || std::memcmp(file.data(), "glTF", 4) != 0) return false;
    if (u32(4) != 2) throw Error(SOLB_ERR_UNSUPPORTED, "load_scene: GLB container version must be 2");
    const size_t total = std::min<size_t>(u32(8), file.size());
    size_t off = 12;
    bool have_json = false;
    while (off + 8 <= total) {
        const size_t len = u32(off);
        const uint32_t type = u32(off + 4);
        if (off + 8 + len > total) throw Error(SOLB_ERR_INVALID, "load_scene: GLB chunk runs past the end of the file");
        if (type == 0x4E4F534Au && !have_json) { json_out.assign(file, off + 8, len); have_json = true; }
        else if (type == 0x004E4942u && bin_out.empty()) bin_out.assign(file, off + 8, len);
        off += 8 + ((len + 3) & ~(size_t)3);
    }
    if (!have_json) throw Error(SOLB_ERR_INVALID, "load_scene: GLB without a JSON chunk");
    return true;
}

Scene load_scene(std::shared_ptr<Context>, const std::string &filepath) {
    Doc d;
    std::string glb_bin;
    bool is_glb = false;
    try {
        const std::string file = read_file(filepath, true);
        std::string json_text;
        is_glb = split_glb(file, json_text, glb_bin);
        d.root = json::parse(is_glb ? json_text : file);
    } catch (const std::runtime_error &e) {
        throw Error(SOLB_ERR_INVALID, std::string("load_scene: ") + e.what());
    }
    std::string dir = ".";
    const size_t slash = filepath.find_last_of('/');
    if (slash != std::string::npos) dir = filepath.substr(0, slash);
    const json::Value &buffers = d.root["buffers"];
    for (size_t i = 0; i < buffers.size(); i++) {
        const std::string uri = buffers[i].has("uri") ? buffers[i]["uri"].string() : std::string();
        std::string data;
        if (uri.empty()) {  // GLB-stored buffer: only buffer 0 may omit its uri
            if (!is_glb || i != 0) throw Error(SOLB_ERR_INVALID, "load_scene: buffer without uri outside a GLB container");
            data = glb_bin;
        } else if (uri.compare(0, 5, "data:") == 0) {
            const size_t comma = uri.find(',');
            if (comma == std::string::npos) throw Error(SOLB_ERR_INVALID, "load_scene: malformed data URI");
            data = base64_decode(uri, comma + 1);
        } else {
            data = read_file(dir + "/" + uri_decode(uri), true);
        }
        const size_t want = (size_t)buffers[i]["byteLength"].integer(0);
        if (data.size() < want) throw Error(SOLB_ERR_INVALID, "load_scene: buffer shorter than byteLength");
        data.resize(want);
        d.buffers.push_back(std::move(data));
    }

    Scene scene;
    const json::Value &mats = d.root["materials"];
    for (size_t i = 0; i < mats.size(); i++) {  // mod.rs:144-156 with the gltf crate's defaults
        const json::Value &pbr = mats[i]["pbrMetallicRoughness"];
        MaterialInfo m;
        std::memset(&m, 0, sizeof(m));
        for (size_t k = 0; k < 4; k++) m.base_color[k] = pbr.has("baseColorFactor") ? (float)pbr["baseColorFactor"][k].number(1.0) : 1.0f;
        for (size_t k = 0; k < 3; k++) m.emissive[k] = mats[i].has("emissiveFactor") ? (float)mats[i]["emissiveFactor"][k].number(0.0) : 0.0f;
        m.metallic = (float)pbr["metallicFactor"].number(1.0);
        m.roughness = (float)pbr["roughnessFactor"].number(1.0);
        scene.materials.push_back(m);
    }
    // base-colour textures: material -> texture -> (image, sampler).  One Scene::textures entry per glTF texture in use.
    {
        const json::Value &textures = d.root["textures"], &images = d.root["images"], &samplers = d.root["samplers"];
        std::vector<uint32_t> slot(textures.size(), SOLB_NO_TEXTURE - 1u);  // SOLB_NO_TEXTURE - 1: not tried yet
        scene.material_textures.assign(mats.size(), SOLB_NO_TEXTURE);
        for (size_t i = 0; i < mats.size(); i++) {
            const json::Value &pbr = mats[i]["pbrMetallicRoughness"];
            if (!pbr.has("baseColorTexture")) continue;
            const json::Value &ref = pbr["baseColorTexture"];
            const size_t ti = (size_t)ref["index"].integer(0);
            if (ti >= textures.size() || ref["texCoord"].integer(0) != 0) continue;
            if (slot[ti] == SOLB_NO_TEXTURE - 1u) {
                slot[ti] = SOLB_NO_TEXTURE;
                try {
                    const json::Value &tex = textures[ti];
                    if (!tex.has("source")) throw Error(SOLB_ERR_UNSUPPORTED, "texture without source");
                    const json::Value &img = images[(size_t)tex["source"].integer(0)];
                    std::string bytes;
                    if (img.has("uri")) {
                        const std::string uri = img["uri"].string();
                        if (uri.compare(0, 5, "data:") == 0) {
                            const size_t comma = uri.find(',');
                            if (comma == std::string::npos) throw Error(SOLB_ERR_INVALID, "malformed data URI");
                            bytes = base64_decode(uri, comma + 1);
                        } else bytes = read_file(dir + "/" + uri_decode(uri), true);
                    } else {
                        const json::Value &view = d.root["bufferViews"][(size_t)img["bufferView"].integer(0)];
                        const std::string &buf = d.buffers.at((size_t)view["buffer"].integer(0));
                        const size_t off = (size_t)view["byteOffset"].integer(0), len = (size_t)view["byteLength"].integer(0);
                        if (off + len > buf.size()) throw Error(SOLB_ERR_INVALID, "image bufferView outside its buffer");
                        bytes = buf.substr(off, len);
                    }
                    const image::Rgba8 im = image::decode_png((const uint8_t *)bytes.data(), bytes.size());
                    Texture t;
                    t.width = im.width;
                    t.height = im.height;
                    t.rgba8 = im.pixels;
                    if (tex.has("sampler")) {
                        const json::Value &smp = samplers[(size_t)tex["sampler"].integer(0)];
                        t.wrap_s = (uint32_t)smp["wrapS"].integer(10497);
                        t.wrap_t = (uint32_t)smp["wrapT"].integer(10497);
                    }
                    slot[ti] = (uint32_t)scene.textures.size();
                    scene.textures.push_back(std::move(t));
                } catch (const std::exception &) {
                    // not a PNG / missing file / broken stream: the material stays untextured, as it is for the reference
                }
            }
            scene.material_textures[i] = slot[ti];
        }
    }

    const json::Value &meshes = d.root["meshes"];
    for (size_t mi = 0; mi < meshes.size(); mi++) {
        Mesh mesh;
        mesh.name = meshes[mi]["name"].string();
        const json::Value &prims = meshes[mi]["primitives"];
        for (size_t pi = 0; pi < prims.size(); pi++) {
            const json::Value &prim = prims[pi];
            const json::Value &attrs = prim["attributes"];
            const size_t offset = mesh.vertices.size();
            if (attrs.has("POSITION")) {
                const AccessorView pos = accessor(d, (size_t)attrs["POSITION"].integer(0));
                AccessorView nrm, uv, col;
                const bool has_n = attrs.has("NORMAL"), has_uv = attrs.has("TEXCOORD_0"), has_c = attrs.has("COLOR_0");
                if (has_n) nrm = accessor(d, (size_t)attrs["NORMAL"].integer(0));
                if (has_uv) uv = accessor(d, (size_t)attrs["TEXCOORD_0"].integer(0));
                if (has_c) col = accessor(d, (size_t)attrs["COLOR_0"].integer(0));
                for (size_t i = 0; i < pos.count; i++) {
                    ModelVertex v;
                    v.pos[0] = read_f32(pos, i, 0); v.pos[1] = read_f32(pos, i, 1); v.pos[2] = read_f32(pos, i, 2); v.pos[3] = 1.0f;
                    v.color[0] = v.color[1] = v.color[2] = v.color[3] = 1.0f;
                    if (has_c && i < col.count) {
                        for (int c = 0; c < col.ncomp && c < 4; c++) v.color[c] = read_f32(col, i, c);  // vec3 -> alpha 1
                    }
                    v.normal[0] = 0.0f; v.normal[1] = 1.0f; v.normal[2] = 0.0f; v.normal[3] = 1.0f;
                    if (has_n && i < nrm.count) { v.normal[0] = read_f32(nrm, i, 0); v.normal[1] = read_f32(nrm, i, 1); v.normal[2] = read_f32(nrm, i, 2); }
                    v.uv[0] = v.uv[1] = v.uv[2] = v.uv[3] = 0.0f;
                    if (has_uv && i < uv.count) { v.uv[0] = read_f32(uv, i, 0); v.uv[1] = read_f32(uv, i, 1); }
                    mesh.vertices.push_back(v);
                }
            }
            PrimitiveSection sec;
            sec.index = pi;
            sec.vertices = { offset, mesh.vertices.size() - offset };
            if (prim.has("material")) sec.material_index = (size_t)prim["material"].integer(0);
            if (prim.has("indices")) {
                const AccessorView idx = accessor(d, (size_t)prim["indices"].integer(0));
                const size_t ioff = mesh.indices.size();
                for (size_t i = 0; i < idx.count; i++) mesh.indices.push_back(read_u32(idx, i));
                sec.indices = BufferPart{ ioff, mesh.indices.size() - ioff };
            }
            mesh.primitive_sections.push_back(sec);
        }
        long first_node = -1;
        mesh.transform = calc_mesh_global_transform(d, (long)mi, &first_node);
        mesh.extra_instance_transforms = other_node_transforms(d, (long)mi, first_node);
        scene.meshes.push_back(std::move(mesh));
    }

    // first camera only; orthographic -> none (mod.rs:261-287)
    const json::Value &cams = d.root["cameras"];
    if (cams.size() > 0 && cams[0]["type"].string() == "perspective") {
        const json::Value &persp = cams[0]["perspective"];
        const json::Value &nodes = d.root["nodes"];
        for (size_t i = 0; i < nodes.size(); i++) {
            if (nodes[i].has("camera") && nodes[i]["camera"].integer(-1) == 0) {
                scene.camera = Camera::from_view(node_matrix(nodes[i]), (float)persp["yfov"].number(0.0),
                                                 (float)persp["znear"].number(0.0), (float)persp["zfar"].number(100.0));
                break;
            }
        }
    }
    return scene;
}

}  // namespace scene
}  // namespace sol
