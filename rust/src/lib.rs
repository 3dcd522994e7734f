//! Host crate: the reference's scene-construction API (same names and argument order as
//! carlosconley/surely-raytracing: Sphere::new, Quad::new, make_box, Translate::new, RotateY::new,
//! ConstantMedium::new, Lambertian::new, ..., HittableList::{new, add, create_bvh}, Camera::new,
//! init_pixels, render_par, render_par_lights) on top of librtb200.so.
//!
//! The types keep what the constructors receive; `flatten()` walks the object tree in `add` order
//! and emits the plain arrays of include/rtb200.h (one node per occurrence, so canonical primitive
//! ids follow insertion order).  Derived fields (bounding boxes, quad normal/d/w, sin/cos, the
//! camera frame) are recomputed inside the library exactly as the reference constructors do.
//!
//! SOURCE ONLY in this repository: the CI image has no cargo/rustc.  The C++ mirror
//! (surely_raytracing_b200/host/rtb/*.hpp) is the same code shape and is what the tests exercise.
pub mod ffi;
pub mod host_rng;
pub mod scene;

pub use host_rng::{random_double, random_range, seed_host_rng};
pub use scene::*;

use std::ffi::CStr;
use std::sync::Arc;

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::rtb_last_error()).to_string_lossy().into_owned() }
}

/// reference src/render.rs:136-138
pub fn init_pixels(cam: &Camera) -> Vec<Color> {
    vec![Color::new_zero(); (cam.image_height() * cam.image_width) as usize]
}

/// reference src/render.rs:140-142: an EMPTY light list (the library then samples the material pdf alone)
pub fn render_par(cam: &Camera, world: &HittableList, pixels: &mut Vec<Color>, suns: &Vec<Sun>) {
    render_par_lights(cam, world, pixels, suns, Arc::new(Object::List(Arc::new(HittableList::new()))))
}

/// reference src/render.rs:144-216.  Same signature; the rayon loop is one call into librtb200.so.
/// Panics on error, like the reference does (`expect`/`panic!`).
pub fn render_par_lights(cam: &Camera, world: &HittableList, pixels: &mut Vec<Color>, _suns: &Vec<Sun>, lights: Arc<Object>) {
    println!("P3\n{} {}\n255", cam.image_width, cam.image_height());
    let flat = FlatScene::new(cam, world, &lights);
    let desc = flat.desc();
    let mut scene: *mut ffi::rtb_scene = std::ptr::null_mut();
    let rc = unsafe { ffi::rtb_scene_create(&desc, 0, &mut scene) };
    if rc != 0 {
        panic!("rtb_scene_create: {}", last_error());
    }
    let mut info = ffi::RtbSceneInfo::default();
    unsafe { ffi::rtb_scene_info(scene, &mut info) };
    assert_eq!(pixels.len(), (info.image_width * info.image_height) as usize, "use init_pixels");
    let params = ffi::RtbRenderParams { sample_begin: 0, sample_end: info.spp_used as i64, pipeline: 0, collect_stats: 0 };
    // Color is #[repr(C)] {x, y, z: f64}: the pixel vector is the f64 rgb-sum buffer the ABI wants
    let rc = unsafe { ffi::rtb_render(scene, &params, pixels.as_mut_ptr() as *mut f64, std::ptr::null_mut()) };
    if rc != 0 {
        let msg = last_error();
        unsafe { ffi::rtb_scene_destroy(scene) };
        panic!("rtb_render: {}", msg);
    }
    eprintln!("\rWriting...            ");
    let mut rgb8 = vec![0u8; pixels.len() * 3];
    let rc = unsafe { ffi::rtb_write_color(scene, pixels.as_ptr() as *const f64, pixels.len() as i64, info.spp_used as f64, 0.0, rgb8.as_mut_ptr()) };
    if rc != 0 {
        panic!("rtb_write_color: {}", last_error());
    }
    let mut text = String::with_capacity(pixels.len() * 12);
    for px in rgb8.chunks(3) {
        text.push_str(&format!("{} {} {}\n", px[0], px[1], px[2]));
    }
    print!("{}", text);
    unsafe { ffi::rtb_scene_destroy(scene) };
    eprintln!("\rDone!                           ");
}
