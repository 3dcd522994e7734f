//! `extern "C"` mirror of include/rtb200.h (ABI version 2). Field order and types must match the
//! header exactly; tests/test_abi.py pins the struct sizes on the C side.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const RTB_ABI_VERSION: i32 = 2;
pub const RTB_FLAG_ISO_PDF_ZERO: u32 = 1;
pub const RTB_FLAG_PROPAGATE_NAN: u32 = 2;
pub const RTB_FLAG_BVH4: u32 = 0x10;
pub const RTB_FLAG_QNODES: u32 = 0x20;
pub const RTB_FLAG_BVH_LEAF4: u32 = 0x40;
pub const RTB_FLAG_NO_BOX_SCAN: u32 = 0x80;
/// switches the sun term HEAD comments out (src/render.rs:300-308) back on
pub const RTB_FLAG_SUN_LIGHT: u32 = 0x100;
pub const RTB_FLAG_RUSSIAN_ROULETTE: u32 = 0x200;
pub const RTB_FLAG_NO_BOX_LEAVES: u32 = 0x400;

pub const OBJ_SPHERE: i32 = 0;
pub const OBJ_QUAD: i32 = 1;
pub const OBJ_LIST: i32 = 2;
pub const OBJ_BVH: i32 = 3;
pub const OBJ_TRANSLATE: i32 = 4;
pub const OBJ_ROTATE_Y: i32 = 5;
pub const OBJ_MEDIUM: i32 = 6;

pub const MAT_LAMBERTIAN: i32 = 0;
pub const MAT_METAL: i32 = 1;
pub const MAT_DIELECTRIC: i32 = 2;
pub const MAT_DIFFUSE_LIGHT: i32 = 3;
pub const MAT_ISOTROPIC: i32 = 4;

pub const TEX_SOLID: i32 = 0;
pub const TEX_CHECKER: i32 = 1;
pub const TEX_IMAGE: i32 = 2;
pub const TEX_NOISE: i32 = 3;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RtbObject {
    pub kind: i32,
    pub material: i32,
    pub first: i32,
    pub count: i32,
    pub v: [f64; 10],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RtbMaterial {
    pub kind: i32,
    pub texture: i32,
    pub color: [f64; 3],
    pub param: f64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RtbTexture {
    pub kind: i32,
    pub a: i32,
    pub b: i32,
    pub reserved: i32,
    pub color: [f64; 3],
    pub scale: f64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RtbImage {
    pub width: i32,
    pub height: i32,
    pub rgb: *const u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RtbPerlin {
    pub ranvec: [[f64; 3]; 256],
    pub perm_x: [i32; 256],
    pub perm_y: [i32; 256],
    pub perm_z: [i32; 256],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RtbCamera {
    pub aspect_ratio: f64,
    pub image_width: i32,
    pub samples_per_pixel: i32,
    pub max_depth: i32,
    pub reserved: i32,
    pub vfov: f64,
    pub lookfrom: [f64; 3],
    pub lookat: [f64; 3],
    pub vup: [f64; 3],
    pub defocus_angle: f64,
    pub focus_dist: f64,
    pub background: [f64; 3],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RtbSun {
    pub direction: [f64; 3],
    pub albedo: [f64; 3],
    pub angular_diameter: f64,
}

#[repr(C)]
pub struct RtbSceneDesc {
    pub abi_version: i32,
    pub flags: u32,
    pub seed: u64,
    pub objects: *const RtbObject,
    pub n_objects: i32,
    pub children: *const i32,
    pub n_children: i32,
    pub world: i32,
    pub lights: *const i32,
    pub n_lights: i32,
    pub materials: *const RtbMaterial,
    pub n_materials: i32,
    pub textures: *const RtbTexture,
    pub n_textures: i32,
    pub images: *const RtbImage,
    pub n_images: i32,
    pub perlins: *const RtbPerlin,
    pub n_perlins: i32,
    pub camera: RtbCamera,
    pub suns: *const RtbSun,
    pub n_suns: i32,
    pub reserved: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct RtbSceneInfo {
    pub image_width: i32,
    pub image_height: i32,
    pub spp_used: i32,
    pub sqrt_spp: i32,
    pub max_depth: i32,
    pub n_surface_prims: i32,
    pub n_boundary_prims: i32,
    pub n_media: i32,
    pub n_bvh_nodes: i32,
    pub n_lights: i32,
    pub bvh_depth: i32,
    pub device: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct RtbRenderParams {
    pub sample_begin: i64,
    pub sample_end: i64,
    pub pipeline: i32,
    pub collect_stats: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct RtbStats {
    pub paths: u64,
    pub segments: u64,
    pub node_visits: u64,
    pub prim_tests: u64,
    pub medium_probes: u64,
    pub nonfinite_samples: u64,
    pub kernel_launches: u64,
    pub device_ms: f64,
    pub exact_tests: u64,
    pub overflow_rays: u64,
    pub stage_ms: [f64; 3],
}

#[repr(C)]
pub struct rtb_scene {
    _private: [u8; 0],
}

extern "C" {
    pub fn rtb_version() -> c_int;
    pub fn rtb_device_count() -> c_int;
    pub fn rtb_last_error() -> *const c_char;
    pub fn rtb_scene_create(desc: *const RtbSceneDesc, device: c_int, out: *mut *mut rtb_scene) -> c_int;
    pub fn rtb_scene_destroy(scene: *mut rtb_scene);
    pub fn rtb_scene_info(scene: *const rtb_scene, info: *mut RtbSceneInfo) -> c_int;
    pub fn rtb_render(scene: *mut rtb_scene, params: *const RtbRenderParams, pixels_rgb: *mut f64, stats: *mut RtbStats) -> c_int;
    /// d_accum: w*h x 4 u64 {r, g, b, count}, 2^-32 fixed-point sums (order-independent: bit-reproducible, split-invariant)
    pub fn rtb_render_device(scene: *mut rtb_scene, params: *const RtbRenderParams, d_accum_u64x4: *mut c_void, cuda_stream: *mut c_void) -> c_int;
    pub fn rtb_render_stats(scene: *mut rtb_scene, stats: *mut RtbStats) -> c_int;
    pub fn rtb_accum_to_pixels(scene: *mut rtb_scene, d_accum_u64x4: *const c_void, pixels_rgb: *mut f64) -> c_int;
    /// the reference seam on the GPUs of one box: threads + streams per GPU, one NCCL int64 sum-reduce, one D2H
    pub fn rtb_render_multi(desc: *const RtbSceneDesc, n_devices: c_int, devices: *const c_int, params: *const RtbRenderParams,
                            pixels_rgb: *mut f64, stats: *mut RtbStats) -> c_int;
    pub fn rtb_scene_set_option(scene: *mut rtb_scene, option: c_int, value: i64) -> c_int;
    pub fn rtb_trim_cache() -> i64;
    /// auto_expose (src/render.rs:325-339) in the reference's own sequential order
    pub fn rtb_auto_expose(pixels_rgb: *const f64, n_pixels: i64, spp: f64, exposure_out: *mut f64) -> c_int;
    pub fn rtb_write_color(scene: *mut rtb_scene, pixels_rgb: *const f64, n_pixels: i64, spp: f64, exposure: f64, rgb8_out: *mut u8) -> c_int;
}
