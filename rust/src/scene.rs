//! Scene-construction API with the reference's names and signatures, plus `flatten`.
use crate::ffi::*;
use crate::host_rng::{random_double, random_int, random_range};
use std::sync::Arc;

#[repr(C)]
#[derive(Debug, Clone, Copy)]
pub struct Vec3 {
    x: f64,
    y: f64,
    z: f64,
}
pub type Point3 = Vec3;
pub type Color = Vec3;

impl Vec3 {
    pub fn new(x: f64, y: f64, z: f64) -> Self { Vec3 { x, y, z } }
    pub fn new_zero() -> Self { Vec3 { x: 0., y: 0., z: 0. } }
    pub fn x(&self) -> f64 { self.x }
    pub fn y(&self) -> f64 { self.y }
    pub fn z(&self) -> f64 { self.z }
    fn arr(&self) -> [f64; 3] { [self.x, self.y, self.z] }
    pub fn length_squared(&self) -> f64 { self.x * self.x + self.y * self.y + self.z * self.z }
    pub fn length(&self) -> f64 { self.length_squared().sqrt() }
}
impl std::ops::Add for Vec3 { type Output = Vec3; fn add(self, o: Vec3) -> Vec3 { Vec3::new(self.x + o.x, self.y + o.y, self.z + o.z) } }
impl std::ops::Sub for Vec3 { type Output = Vec3; fn sub(self, o: Vec3) -> Vec3 { Vec3::new(self.x - o.x, self.y - o.y, self.z - o.z) } }
impl std::ops::Neg for Vec3 { type Output = Vec3; fn neg(self) -> Vec3 { Vec3::new(-self.x, -self.y, -self.z) } }
impl std::ops::Mul<Vec3> for f64 { type Output = Vec3; fn mul(self, v: Vec3) -> Vec3 { Vec3::new(self * v.x, self * v.y, self * v.z) } }
impl std::ops::Mul<Vec3> for Vec3 { type Output = Vec3; fn mul(self, v: Vec3) -> Vec3 { Vec3::new(self.x * v.x, self.y * v.y, self.z * v.z) } }
impl std::ops::Div<f64> for Vec3 { type Output = Vec3; fn div(self, t: f64) -> Vec3 { Vec3::new(self.x / t, self.y / t, self.z / t) } }
impl std::ops::Mul<f64> for Vec3 { type Output = Vec3; fn mul(self, t: f64) -> Vec3 { t * self } }              // vec3.rs:113-119 (`Color::new(..) * 10.`, main.rs:86)
impl std::ops::AddAssign for Vec3 { fn add_assign(&mut self, o: Vec3) { *self = *self + o; } }                     // vec3.rs:74-80
impl std::ops::MulAssign<f64> for Vec3 { fn mul_assign(&mut self, t: f64) { *self = t * *self; } }                 // vec3.rs:121-127
impl std::ops::DivAssign<f64> for Vec3 { fn div_assign(&mut self, t: f64) { *self = *self / t; } }                 // vec3.rs:153-157
pub fn dot(u: &Vec3, v: &Vec3) -> f64 { u.x * v.x + u.y * v.y + u.z * v.z }                                        // vec3.rs:167-169
pub fn cross(u: &Vec3, v: &Vec3) -> Vec3 { Vec3::new(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x) }  // vec3.rs:171-177
pub fn unit_vector(v: &Vec3) -> Vec3 { *v / v.length() }
pub fn random_vec3() -> Vec3 { Vec3::new(random_double(), random_double(), random_double()) }                      // vec3.rs:193-195
pub fn random_vec3_range(min: f64, max: f64) -> Vec3 { Vec3::new(random_range(min, max), random_range(min, max), random_range(min, max)) }

// ---- textures -----------------------------------------------------------------------------------
pub enum Texture {
    Solid(Color),
    Checker { inv_scale: f64, even: Arc<Texture>, odd: Arc<Texture> },
    Image { width: i32, height: i32, rgb: Arc<Vec<u8>> },
    Noise { scale: f64, perlin: Box<RtbPerlin> },
}
pub struct SolidColor;
impl SolidColor { pub fn new(c: Color) -> Texture { Texture::Solid(c) } }
pub struct CheckerTexture;
impl CheckerTexture {
    pub fn from_color(scale: f64, c1: Color, c2: Color) -> Texture {
        Texture::Checker { inv_scale: 1. / scale, even: Arc::new(Texture::Solid(c1)), odd: Arc::new(Texture::Solid(c2)) }
    }
}
pub struct ImageTexture;
impl ImageTexture {
    /// RGB8, top row first (what `image::open(..).to_rgb8()` yields). A crate user with the `image`
    /// dependency decodes the file and passes the bytes; the library never touches the file system.
    pub fn from_rgb8(width: i32, height: i32, rgb: Vec<u8>) -> Texture { Texture::Image { width, height, rgb: Arc::new(rgb) } }

    /// reference src/texture.rs:89 + src/rt_image.rs:13-27: decode `filename` to RGB8.  With the `image` feature the
    /// file goes through the same `image` crate call the reference makes (`image::open(..).to_rgb8()`); without it only
    /// binary PPM (P6, maxval 255) is read.  Panics like the reference when the file cannot be opened.
    pub fn new(filename: &str) -> Texture {
        #[cfg(feature = "image")]
        {
            let img = image::open(filename).expect("Could not open image.").to_rgb8();
            let (w, h) = (img.width() as i32, img.height() as i32);
            return Self::from_rgb8(w, h, img.into_raw());
        }
        #[cfg(not(feature = "image"))]
        {
            let bytes = std::fs::read(filename).expect("Could not open image.");
            let mut fields = Vec::new();
            let mut pos = 0usize;
            while fields.len() < 4 {
                while pos < bytes.len() && bytes[pos].is_ascii_whitespace() { pos += 1; }
                let start = pos;
                while pos < bytes.len() && !bytes[pos].is_ascii_whitespace() { pos += 1; }
                fields.push(String::from_utf8_lossy(&bytes[start..pos]).into_owned());
            }
            pos += 1;
            let (w, h): (usize, usize) = (fields[1].parse().expect("Could not open image."), fields[2].parse().expect("Could not open image."));
            assert!(fields[0] == "P6" && fields[3] == "255" && bytes.len() >= pos + 3 * w * h, "Could not open image.");
            Self::from_rgb8(w as i32, h as i32, bytes[pos..pos + 3 * w * h].to_vec())
        }
    }
}
pub struct NoiseTexture;
impl NoiseTexture {
    pub fn new(scale: f64) -> Texture {
        let mut p = Box::new(RtbPerlin { ranvec: [[0.; 3]; 256], perm_x: [0; 256], perm_y: [0; 256], perm_z: [0; 256] });
        for i in 0..256 {
            p.ranvec[i] = unit_vector(&random_vec3_range(-1., 1.)).arr();
        }
        for perm in [&mut p.perm_x, &mut p.perm_y, &mut p.perm_z] {
            for (i, v) in perm.iter_mut().enumerate() { *v = i as i32; }
            for i in (0..256usize).rev() {
                let target = random_int(0, i as i64) as usize;
                perm.swap(i, target);
            }
        }
        Texture::Noise { scale, perlin: p }
    }
}

// ---- materials ----------------------------------------------------------------------------------
#[derive(Clone)]
pub enum Material {
    Lambertian(Arc<Texture>),
    Metal { albedo: Color, fuzz: f64 },
    Dielectric { tint: Color, ir: f64 },
    DiffuseLight(Arc<Texture>),
    Isotropic(Arc<Texture>),
}
pub struct Lambertian;
impl Lambertian {
    pub fn new(albedo: Color) -> Material { Material::Lambertian(Arc::new(Texture::Solid(albedo))) }
    pub fn from_texture(t: Arc<Texture>) -> Material { Material::Lambertian(t) }
}
pub struct Metal;
impl Metal { pub fn new(albedo: Color, f: f64) -> Material { Material::Metal { albedo, fuzz: if f < 1. { f } else { 1. } } } }
pub struct Dielectric;
impl Dielectric {
    pub fn new(ir: f64, tint: Color) -> Material { Material::Dielectric { tint, ir } }
    pub fn new_clear(ir: f64) -> Material { Self::new(ir, Color::new(1., 1., 1.)) }
}
pub struct DiffuseLight;
impl DiffuseLight { pub fn new(c: Color) -> Material { Material::DiffuseLight(Arc::new(Texture::Solid(c))) } }
pub struct Isotropic;
impl Isotropic { pub fn new(c: Color) -> Material { Material::Isotropic(Arc::new(Texture::Solid(c))) } }

// ---- objects ------------------------------------------------------------------------------------
#[derive(Clone)]
pub enum Object {
    Sphere { center: Point3, radius: f64, center_vec: Option<Vec3>, mat: Material },
    Quad { q: Point3, u: Vec3, v: Vec3, mat: Material },
    List(Arc<HittableList>),
    Node(Arc<HittableList>), // create_bvh marker: the device builds its own BVH
    Translate { object: Arc<Object>, offset: Vec3 },
    RotateY { object: Arc<Object>, angle: f64 },
    Volume { boundary: Arc<Object>, density: f64, phase: Material },
}
pub struct Sphere;
impl Sphere {
    pub fn new(center: Point3, radius: f64, mat: Material) -> Object { Object::Sphere { center, radius, center_vec: None, mat } }
    pub fn new_moving(c1: Point3, c2: Point3, radius: f64, mat: Material) -> Object { Object::Sphere { center: c1, radius, center_vec: Some(c2 - c1), mat } }
}
pub struct Quad;
impl Quad { pub fn new(q: Point3, u: Vec3, v: Vec3, mat: Material) -> Object { Object::Quad { q, u, v, mat } } }
pub struct Translate;
impl Translate { pub fn new(p: Arc<Object>, displacement: Vec3) -> Object { Object::Translate { object: p, offset: displacement } } }
pub struct RotateY;
impl RotateY { pub fn new(p: Arc<Object>, angle: f64) -> Object { Object::RotateY { object: p, angle } } }
pub struct ConstantMedium;
impl ConstantMedium { pub fn new(boundary: Arc<Object>, density: f64, c: Color) -> Object { Object::Volume { boundary, density, phase: Isotropic::new(c) } } }
/// reference src/object.rs:216-241.  Accepted and ignored by the integrator at HEAD (the sun term of ray_color is
/// commented out, src/render.rs:300-308); the records still reach the library, which adds the term back under
/// RTB_FLAG_SUN_LIGHT.
pub struct Sun { pub direction: Vec3, pub albedo: Color, pub angular_diameter: f64 }
impl Sun {
    pub fn new(direction: Vec3, albedo: Color, angular_diameter: f64) -> Sun { Sun { direction: unit_vector(&direction), albedo, angular_diameter } }
}

pub struct HittableList { pub objects: Vec<Object> }
impl HittableList {
    pub fn new() -> HittableList { HittableList { objects: vec![] } }
    pub fn from_object(obj: Object) -> HittableList { HittableList { objects: vec![obj] } }
    pub fn add(&mut self, object: Object) { self.objects.push(object); }
    pub fn create_bvh(&mut self) -> HittableList {
        HittableList::from_object(Object::Node(Arc::new(HittableList { objects: self.objects.clone() })))
    }
}

pub fn make_box(a: &Point3, b: &Point3, mat: &Material) -> Object {
    let mn = Point3::new(a.x().min(b.x()), a.y().min(b.y()), a.z().min(b.z()));
    let mx = Point3::new(a.x().max(b.x()), a.y().max(b.y()), a.z().max(b.z()));
    let dx = Vec3::new(mx.x() - mn.x(), 0., 0.);
    let dy = Vec3::new(0., mx.y() - mn.y(), 0.);
    let dz = Vec3::new(0., 0., mx.z() - mn.z());
    let mut sides = HittableList::new();
    sides.add(Quad::new(Point3::new(mn.x(), mn.y(), mx.z()), dx, dy, mat.clone())); // front
    sides.add(Quad::new(Point3::new(mx.x(), mn.y(), mx.z()), -dz, dy, mat.clone())); // right
    sides.add(Quad::new(Point3::new(mx.x(), mn.y(), mn.z()), -dx, dy, mat.clone())); // back
    sides.add(Quad::new(Point3::new(mn.x(), mn.y(), mn.z()), dz, dy, mat.clone())); // left
    sides.add(Quad::new(Point3::new(mn.x(), mx.y(), mx.z()), dx, -dz, mat.clone())); // top
    sides.add(Quad::new(Point3::new(mn.x(), mn.y(), mn.z()), dx, dz, mat.clone())); // bottom
    Object::List(Arc::new(sides))
}

// ---- camera -------------------------------------------------------------------------------------
pub struct Camera {
    pub aspect_ratio: f64,
    pub image_width: i32,
    pub samples_per_pixel: i32,
    pub max_depth: i32,
    pub vfov: f64,
    pub lookfrom: Point3,
    pub lookat: Point3,
    pub vup: Vec3,
    pub defocus_angle: f64,
    pub focus_dist: f64,
    pub background: Color,
    pub auto_exposure: bool,
}
impl Camera {
    #[allow(clippy::too_many_arguments)]
    pub fn new(aspect_ratio: f64, image_width: i32, samples_per_pixel: i32, max_depth: i32, vfov: f64, lookfrom: Point3,
               lookat: Point3, vup: Vec3, defocus_angle: f64, focus_dist: f64, background: Color) -> Self {
        Camera { aspect_ratio, image_width, samples_per_pixel, max_depth, vfov, lookfrom, lookat, vup, defocus_angle, focus_dist, background, auto_exposure: false }
    }
    pub fn image_height(&self) -> i32 { std::cmp::max(1, (self.image_width as f64 / self.aspect_ratio) as i32) }
}

// ---- flatten ------------------------------------------------------------------------------------
#[derive(Default)]
pub struct FlatScene {
    objects: Vec<RtbObject>,
    children: Vec<i32>,
    lights: Vec<i32>,
    materials: Vec<RtbMaterial>,
    textures: Vec<RtbTexture>,
    images: Vec<RtbImage>,
    image_bytes: Vec<Arc<Vec<u8>>>,
    perlins: Vec<RtbPerlin>,
    suns: Vec<RtbSun>,
    tex_seen: Vec<(*const Texture, i32)>,     // textures are interned by Arc pointer, materials by value:
    mat_seen: Vec<(RtbMaterial, i32)>,        // final_scene's 1000 spheres share ONE white Lambertian record
    world: i32,
    camera: Option<RtbCamera>,
    pub flags: u32,
    pub seed: u64,
}

impl FlatScene {
    pub fn new(cam: &Camera, world: &HittableList, lights: &Arc<Object>) -> Self {
        let mut f = FlatScene { seed: 20240001, ..Default::default() };
        f.camera = Some(RtbCamera {
            aspect_ratio: cam.aspect_ratio, image_width: cam.image_width, samples_per_pixel: cam.samples_per_pixel,
            max_depth: cam.max_depth, reserved: 0, vfov: cam.vfov, lookfrom: cam.lookfrom.arr(), lookat: cam.lookat.arr(),
            vup: cam.vup.arr(), defocus_angle: cam.defocus_angle, focus_dist: cam.focus_dist, background: cam.background.arr(),
        });
        f.world = f.emit_list(&world.objects, OBJ_LIST);
        // `lights: Arc<Object>` (src/render.rs:149): pdf_value / random dispatch on ANY object (src/object.rs:53-69) --
        // a list contributes its members, a bare Quad or Sphere is a one-element list, any other kind falls to the
        // Hittable defaults (pdf 0, direction (1,0,0): src/hittable.rs:46-52), which the library reproduces
        match &**lights {
            Object::List(l) | Object::Node(l) => {
                for o in l.objects.iter() {
                    let id = f.emit(o);
                    f.lights.push(id);
                }
            }
            other => {
                let id = f.emit(other);
                f.lights.push(id);
            }
        }
        f
    }

    pub fn set_suns(&mut self, suns: &[Sun]) {
        self.suns = suns.iter().map(|s| RtbSun { direction: s.direction.arr(), albedo: s.albedo.arr(), angular_diameter: s.angular_diameter }).collect();
    }

    pub fn desc(&self) -> RtbSceneDesc {
        RtbSceneDesc {
            abi_version: RTB_ABI_VERSION, flags: self.flags, seed: self.seed,
            objects: self.objects.as_ptr(), n_objects: self.objects.len() as i32,
            children: self.children.as_ptr(), n_children: self.children.len() as i32, world: self.world,
            lights: self.lights.as_ptr(), n_lights: self.lights.len() as i32,
            materials: self.materials.as_ptr(), n_materials: self.materials.len() as i32,
            textures: self.textures.as_ptr(), n_textures: self.textures.len() as i32,
            images: self.images.as_ptr(), n_images: self.images.len() as i32,
            perlins: self.perlins.as_ptr(), n_perlins: self.perlins.len() as i32,
            camera: self.camera.unwrap(),
            suns: self.suns.as_ptr(), n_suns: self.suns.len() as i32, reserved: 0,
        }
    }

    fn emit_texture(&mut self, t: &Texture) -> i32 {
        if let Some((_, id)) = self.tex_seen.iter().find(|(p, _)| std::ptr::eq(*p, t)) { return *id; }
        let id = self.emit_texture_new(t);
        self.tex_seen.push((t as *const Texture, id));
        id
    }

    fn emit_texture_new(&mut self, t: &Texture) -> i32 {
        let mut r = RtbTexture { kind: TEX_SOLID, a: -1, b: -1, reserved: 0, color: [0.; 3], scale: 1. };
        match t {
            Texture::Solid(c) => { r.color = c.arr(); }
            Texture::Checker { inv_scale, even, odd } => { r.kind = TEX_CHECKER; r.scale = *inv_scale; r.a = self.emit_texture(even); r.b = self.emit_texture(odd); }
            Texture::Image { width, height, rgb } => {
                r.kind = TEX_IMAGE;
                self.image_bytes.push(rgb.clone());
                self.images.push(RtbImage { width: *width, height: *height, rgb: rgb.as_ptr() });
                r.a = self.images.len() as i32 - 1;
            }
            Texture::Noise { scale, perlin } => { r.kind = TEX_NOISE; r.scale = *scale; self.perlins.push(**perlin); r.a = self.perlins.len() as i32 - 1; }
        }
        self.textures.push(r);
        self.textures.len() as i32 - 1
    }

    fn emit_material(&mut self, m: &Material) -> i32 {
        let r = match m {
            Material::Lambertian(t) => RtbMaterial { kind: MAT_LAMBERTIAN, texture: self.emit_texture(t), color: [0.; 3], param: 0. },
            Material::Metal { albedo, fuzz } => RtbMaterial { kind: MAT_METAL, texture: -1, color: albedo.arr(), param: *fuzz },
            Material::Dielectric { tint, ir } => RtbMaterial { kind: MAT_DIELECTRIC, texture: -1, color: tint.arr(), param: *ir },
            Material::DiffuseLight(t) => RtbMaterial { kind: MAT_DIFFUSE_LIGHT, texture: self.emit_texture(t), color: [0.; 3], param: 0. },
            Material::Isotropic(t) => RtbMaterial { kind: MAT_ISOTROPIC, texture: self.emit_texture(t), color: [0.; 3], param: 0. },
        };
        let same = |a: &RtbMaterial, b: &RtbMaterial| a.kind == b.kind && a.texture == b.texture && a.param.to_bits() == b.param.to_bits()
            && a.color.iter().zip(b.color.iter()).all(|(x, y)| x.to_bits() == y.to_bits());
        if let Some((_, id)) = self.mat_seen.iter().find(|(m2, _)| same(m2, &r)) { return *id; }
        self.materials.push(r);
        let id = self.materials.len() as i32 - 1;
        self.mat_seen.push((r, id));
        id
    }

    fn emit_list(&mut self, objs: &[Object], kind: i32) -> i32 {
        let me = self.objects.len();
        self.objects.push(RtbObject { kind, material: -1, first: 0, count: 0, v: [0.; 10] });
        let ids: Vec<i32> = objs.iter().map(|o| self.emit(o)).collect();
        self.objects[me].first = self.children.len() as i32;
        self.objects[me].count = ids.len() as i32;
        self.children.extend(ids);
        me as i32
    }

    fn emit(&mut self, o: &Object) -> i32 {
        match o {
            Object::List(l) => self.emit_list(&l.objects, OBJ_LIST),
            Object::Node(l) => self.emit_list(&l.objects, OBJ_BVH),
            _ => {
                let me = self.objects.len();
                self.objects.push(RtbObject { kind: 0, material: -1, first: -1, count: 0, v: [0.; 10] });
                let mut r = self.objects[me];
                match o {
                    Object::Sphere { center, radius, center_vec, mat } => {
                        r.kind = OBJ_SPHERE;
                        r.material = self.emit_material(mat);
                        r.v[..3].copy_from_slice(&center.arr());
                        r.v[3] = *radius;
                        if let Some(cv) = center_vec { r.v[4..7].copy_from_slice(&cv.arr()); r.v[7] = 1.; }
                    }
                    Object::Quad { q, u, v, mat } => {
                        r.kind = OBJ_QUAD;
                        r.material = self.emit_material(mat);
                        r.v[..3].copy_from_slice(&q.arr());
                        r.v[3..6].copy_from_slice(&u.arr());
                        r.v[6..9].copy_from_slice(&v.arr());
                    }
                    Object::Translate { object, offset } => { r.kind = OBJ_TRANSLATE; r.v[..3].copy_from_slice(&offset.arr()); r.first = self.emit(object); r.count = 1; }
                    Object::RotateY { object, angle } => { r.kind = OBJ_ROTATE_Y; r.v[0] = *angle; r.first = self.emit(object); r.count = 1; }
                    Object::Volume { boundary, density, phase } => { r.kind = OBJ_MEDIUM; r.v[0] = *density; r.material = self.emit_material(phase); r.first = self.emit(boundary); r.count = 1; }
                    Object::List(_) | Object::Node(_) => unreachable!(),
                }
                self.objects[me] = r;
                me as i32
            }
        }
    }
}
