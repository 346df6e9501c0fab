//! sol::scene for the ray-tracing path: glTF -> host arrays in the reference's layouts (src/scene/mod.rs:138-295,
//! src/scene/mesh.rs:9-61) and the camera matrices (src/scene/camera.rs:54-127).  The reference uploads every array into a
//! Vulkan buffer as it goes (Buffer::from_data); here they stay in `Vec`s and ray::SceneDescription hands them to libsolb.
//! Mouse manipulators of the camera are out of scope (interactive input).
use crate::ffi::{SolbMaterialInfo, SolbModelVertex};
use crate::Context;
use glam::{Mat4, Vec2, Vec3, Vec4Swizzles};
use std::path::PathBuf;
use std::sync::Arc;

/// src/scene/mesh.rs:9-14: pos.xyz 1 | color rgba | normal.xyz 1 | uv 0 0  (64 bytes)
pub type ModelVertex = SolbModelVertex;
/// src/scene/mod.rs:19-29 (48 bytes)
pub type MaterialInfo = SolbMaterialInfo;

/// src/scene/mod.rs:31-35
#[derive(Clone, Copy, Debug, Default)]
pub struct BufferPart {
    pub offset: usize,
    pub element_count: usize,
}

/// src/scene/mod.rs:37-44,46-97: one glTF primitive of a mesh = one BLAS = one instance
#[derive(Clone, Debug, Default)]
pub struct PrimitiveSection {
    pub index: usize,
    pub vertices: BufferPart,
    pub indices: Option<BufferPart>,
    pub material_index: Option<usize>,
}

impl PrimitiveSection {
    pub fn get_index(&self) -> usize {
        self.index
    }
    pub fn get_vertices(&self) -> &BufferPart {
        &self.vertices
    }
    pub fn get_indices(&self) -> &Option<BufferPart> {
        &self.indices
    }
    pub fn get_vertex_count(&self) -> u32 {
        self.vertices.element_count as u32
    }
    pub fn get_vertex_offset(&self) -> u32 {
        self.vertices.offset as u32
    }
    /// panics on a non-indexed primitive, like the reference's unwrap()
    pub fn get_index_count(&self) -> u32 {
        self.indices.unwrap().element_count as u32
    }
    pub fn get_index_offset(&self) -> u32 {
        self.indices.unwrap().offset as u32
    }
    /// panics on a primitive without material (src/scene/mod.rs:65)
    pub fn get_material_index(&self) -> usize {
        self.material_index.unwrap()
    }
}

/// src/scene/mesh.rs:53-61 with host storage
pub struct Mesh {
    pub name: String,
    pub vertices: Vec<ModelVertex>,
    /// section-relative u32 indices (the u64 copy of src/scene/mod.rs:228-233 is not reproduced)
    pub indices: Vec<u32>,
    pub transform: Mat4,
    pub primitive_sections: Vec<PrimitiveSection>,
    /// beyond the reference: global transforms of the OTHER nodes referencing this mesh
    pub extra_instance_transforms: Vec<Mat4>,
}

/// Base-colour texture of a material (beyond the reference): rgba8 rows top first + the sampler's wrap modes
/// (glTF codes, 10497 = REPEAT).
pub struct Texture {
    pub width: u32,
    pub height: u32,
    pub rgba8: Vec<u8>,
    pub wrap_s: u32,
    pub wrap_t: u32,
}

/// src/scene/mod.rs:99-104 (+ `textures` / `material_textures`, beyond the reference: bound with
/// `SceneDescription::set_textures`; nothing samples them otherwise)
pub struct Scene {
    pub meshes: Vec<Mesh>,
    pub materials: Vec<MaterialInfo>,
    pub camera: Option<Camera>,
    pub textures: Vec<Texture>,
    pub material_textures: Vec<Option<u32>>,
}

fn wrap_code(w: gltf::texture::WrappingMode) -> u32 {
    match w {
        gltf::texture::WrappingMode::ClampToEdge => 33071,
        gltf::texture::WrappingMode::MirroredRepeat => 33648,
        gltf::texture::WrappingMode::Repeat => 10497,
    }
}

fn local_matrix(node: &gltf::Node) -> Mat4 {
    Mat4::from_cols_array_2d(&node.transform().matrix())
}

/// Transform the reference gives a mesh (src/scene/mod.rs:106-136): the document's nodes are tried in index order — every
/// node, not only scene roots — and the first one whose subtree (depth first, children in order) holds the mesh wins; the
/// result is the product of the local matrices from THAT node down to the mesh node.  (A parent listed after its child is
/// therefore dropped: SURVEY a1.)  Also returns the index of the mesh node that was found.
fn reference_mesh_transform(doc: &gltf::Document, mesh_index: usize) -> (Mat4, Option<usize>) {
    // depth-first walk with an explicit stack of (node, matrix accumulated from the start node)
    for start in doc.nodes() {
        let mut pending = vec![(start.clone(), local_matrix(&start))];
        while let Some((node, accumulated)) = pending.pop() {
            if node.mesh().map(|m| m.index()) == Some(mesh_index) {
                return (accumulated, Some(node.index()));
            }
            // push in reverse so the first child is visited first
            let children: Vec<gltf::Node> = node.children().collect();
            for child in children.into_iter().rev() {
                let m = accumulated * local_matrix(&child);
                pending.push((child, m));
            }
        }
    }
    (Mat4::IDENTITY, None)
}

/// glTF node-graph instancing: true global transform (root -> node) of every other node that references the mesh.
fn other_instance_transforms(doc: &gltf::Document, mesh_index: usize, first_node: Option<usize>) -> Vec<Mat4> {
    let n = doc.nodes().count();
    let mut parent = vec![usize::MAX; n];
    for node in doc.nodes() {
        for child in node.children() {
            parent[child.index()] = node.index();
        }
    }
    let locals: Vec<Mat4> = doc.nodes().map(|node| local_matrix(&node)).collect();
    let mut out = Vec::new();
    for node in doc.nodes() {
        if Some(node.index()) == first_node || node.mesh().map(|m| m.index()) != Some(mesh_index) {
            continue;
        }
        let mut chain = Vec::new();
        let mut k = node.index();
        while k != usize::MAX && chain.len() <= n {
            chain.push(k);
            k = parent[k];
        }
        let mut global = Mat4::IDENTITY;
        for &k in chain.iter().rev() {
            global = global * locals[k];
        }
        out.push(global);
    }
    out
}

/// src/scene/mod.rs:138-295.  `context` is accepted for source compatibility (the reference uploads here).
pub fn load_scene(_context: Arc<Context>, filepath: &PathBuf) -> Scene {
    let (doc, buffers, images) = gltf::import(filepath).unwrap();

    // materials with the gltf crate's defaults (base 1,1,1,1; metallic 1; roughness 1; emissive 0)
    let materials: Vec<MaterialInfo> = doc
        .materials()
        .map(|m| {
            let pbr = m.pbr_metallic_roughness();
            MaterialInfo {
                base_color: pbr.base_color_factor(),
                emissive: m.emissive_factor(),
                padding0: 0.0,
                metallic: pbr.metallic_factor(),
                roughness: pbr.roughness_factor(),
                padding1: 0.0,
                padding2: 0.0,
            }
        })
        .collect();

    let mut meshes = Vec::new();
    for mesh in doc.meshes() {
        let mut vertices: Vec<ModelVertex> = Vec::new();
        let mut indices: Vec<u32> = Vec::new();
        let mut sections = Vec::new();
        for (section_index, primitive) in mesh.primitives().enumerate() {
            let reader = primitive.reader(|b| Some(&buffers[b.index()]));
            let first_vertex = vertices.len();
            if let Some(positions) = reader.read_positions() {
                let mut normals = reader.read_normals();
                let mut colors = reader.read_colors(0).map(|c| c.into_rgba_f32());
                let mut uvs = reader.read_tex_coords(0).map(|t| t.into_f32());
                for p in positions {
                    // defaults of src/scene/mod.rs:182-193: colour 1,1,1,1; normal 0,1,0; uv 0,0
                    let n = normals.as_mut().and_then(|it| it.next()).unwrap_or([0.0, 1.0, 0.0]);
                    let c = colors.as_mut().and_then(|it| it.next()).unwrap_or([1.0; 4]);
                    let t = uvs.as_mut().and_then(|it| it.next()).unwrap_or([0.0, 0.0]);
                    vertices.push(ModelVertex {
                        pos: [p[0], p[1], p[2], 1.0],
                        color: c,
                        normal: [n[0], n[1], n[2], 1.0],
                        uv: [t[0], t[1], 0.0, 0.0],
                    });
                }
            }
            let index_part = reader.read_indices().map(|it| {
                let first_index = indices.len();
                indices.extend(it.into_u32()); // relative to the primitive's own vertex range (src/scene/mod.rs:207-215)
                BufferPart { offset: first_index, element_count: indices.len() - first_index }
            });
            sections.push(PrimitiveSection {
                index: section_index,
                vertices: BufferPart { offset: first_vertex, element_count: vertices.len() - first_vertex },
                indices: index_part,
                material_index: primitive.material().index(),
            });
        }
        let (transform, first_node) = reference_mesh_transform(&doc, mesh.index());
        meshes.push(Mesh {
            name: mesh.name().unwrap_or("").to_string(),
            vertices,
            indices,
            transform,
            primitive_sections: sections,
            extra_instance_transforms: other_instance_transforms(&doc, mesh.index(), first_node),
        });
    }

    // first camera only, perspective only; the node's LOCAL matrix is passed as "view" (src/scene/mod.rs:261-287)
    let mut camera = None;
    if let Some(first) = doc.cameras().next() {
        if let gltf::camera::Projection::Perspective(p) = first.projection() {
            for node in doc.nodes() {
                if node.camera().map(|c| c.index()) == Some(first.index()) {
                    camera = Some(Camera::from_view(local_matrix(&node), p.yfov(), p.znear(), p.zfar().unwrap_or(100.0)));
                    break;
                }
            }
        }
    }
    // base-colour textures (texture-coordinate set 0 only), one entry per glTF texture in use
    let mut textures: Vec<Texture> = Vec::new();
    let mut slot: Vec<Option<u32>> = vec![None; doc.textures().count()];
    let mut material_textures: Vec<Option<u32>> = Vec::new();
    for m in doc.materials() {
        let mut bound = None;
        if let Some(info) = m.pbr_metallic_roughness().base_color_texture() {
            if info.tex_coord() == 0 {
                let tex = info.texture();
                if slot[tex.index()].is_none() {
                    let data = &images[tex.source().index()];
                    let rgba8 = match data.format {
                        gltf::image::Format::R8G8B8A8 => Some(data.pixels.clone()),
                        gltf::image::Format::R8G8B8 => Some(data.pixels.chunks(3).flat_map(|p| [p[0], p[1], p[2], 255]).collect()),
                        gltf::image::Format::R8 => Some(data.pixels.iter().flat_map(|&v| [v, v, v, 255]).collect()),
                        gltf::image::Format::R8G8 => Some(data.pixels.chunks(2).flat_map(|p| [p[0], p[0], p[0], p[1]]).collect()),
                        _ => None,
                    };
                    if let Some(rgba8) = rgba8 {
                        slot[tex.index()] = Some(textures.len() as u32);
                        textures.push(Texture {
                            width: data.width,
                            height: data.height,
                            rgba8,
                            wrap_s: wrap_code(tex.sampler().wrap_s()),
                            wrap_t: wrap_code(tex.sampler().wrap_t()),
                        });
                    }
                }
                bound = slot[tex.index()];
            }
        }
        material_textures.push(bound);
    }
    Scene { meshes, materials, camera, textures, material_textures }
}

/// src/scene/camera.rs:34-127 (matrices only)
#[derive(Clone, Copy, Debug)]
pub struct Camera {
    position: Vec3,
    center: Vec3,
    up: Vec3,
    vfov: f32,
    z_near: f32,
    z_far: f32,
    view_matrix: Mat4,
    persp_matrix: Mat4,
    window_size: Vec2,
}

impl Camera {
    /// camera.rs:54-72: eye (10,10,10), centre 0, up -Y, vfov 35, near 0.1, far 1000; only the projection is computed
    pub fn new(window_size: Vec2) -> Camera {
        let mut c = Camera {
            position: Vec3::splat(10.0),
            center: Vec3::ZERO,
            up: -Vec3::Y,
            vfov: 35.0,
            z_near: 0.1,
            z_far: 1000.0,
            view_matrix: Mat4::IDENTITY,
            persp_matrix: Mat4::IDENTITY,
            window_size,
        };
        c.update_persp();
        c
    }

    /// camera.rs:74-97: the matrix is taken as the view matrix as is; the projection stays identity until
    /// set_window_size / set_vfov
    pub fn from_view(view: Mat4, yfov: f32, z_near: f32, z_far: f32) -> Camera {
        let inv = view.inverse();
        let position = inv * glam::vec4(0.0, 0.0, 0.0, 1.0);
        let up = inv * glam::vec4(0.0, 1.0, 0.0, 0.0);
        let center = position + inv * glam::vec4(0.0, 0.0, -4.0, 0.0);
        Camera {
            position: position.xyz(),
            center: center.xyz(),
            up: up.xyz(),
            vfov: yfov,
            z_near,
            z_far,
            view_matrix: view,
            persp_matrix: Mat4::IDENTITY,
            window_size: glam::vec2(1920.0, 1080.0),
        }
    }

    fn update_view(&mut self) {
        self.view_matrix = Mat4::look_at_rh(self.position, self.center, self.up);
    }

    fn update_persp(&mut self) {
        let aspect = self.window_size.x / self.window_size.y;
        self.persp_matrix = Mat4::perspective_rh(self.vfov.to_radians(), aspect, self.z_near, self.z_far);
    }

    pub fn look_at(&mut self, eye: Vec3, center: Vec3, up: Vec3) {
        self.position = eye;
        self.center = center;
        self.up = up;
        self.update_view();
    }

    pub fn set_window_size(&mut self, window_size: Vec2) {
        self.window_size = window_size;
        self.update_persp();
    }

    pub fn set_vfov(&mut self, vfov: f32) {
        self.vfov = vfov;
        self.update_persp();
    }

    pub fn view_matrix(&self) -> Mat4 {
        self.view_matrix
    }

    pub fn perspective_matrix(&self) -> Mat4 {
        self.persp_matrix
    }

    pub fn position(&self) -> Vec3 {
        self.position
    }
}
