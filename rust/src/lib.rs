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

/// reference src/render.rs:144-216.  Same signature; the rayon loop is one call into librtb200.so -- rtb_render on one
/// GPU, rtb_render_multi (one thread + stream per GPU, one NCCL int64 sum-reduce, one read-back) when the job keeps more
/// than one of the box's GPUs busy (>= 128 M paths each; RTB_GPUS overrides).  Panics on error, like the reference does.
pub fn render_par_lights(cam: &Camera, world: &HittableList, pixels: &mut Vec<Color>, suns: &Vec<Sun>, lights: Arc<Object>) {
    println!("P3\n{} {}\n255", cam.image_width, cam.image_height());
    let mut flat = FlatScene::new(cam, world, &lights);
    flat.set_suns(suns);
    let desc = flat.desc();
    let root = (cam.samples_per_pixel as f64).sqrt() as i64;
    let spp_used = root * root; // nearest_square, src/render.rs:38-41, 108
    assert_eq!(pixels.len(), (cam.image_height() * cam.image_width) as usize, "use init_pixels");
    let visible = unsafe { ffi::rtb_device_count() } as i64;
    let wanted = std::env::var("RTB_GPUS").ok().and_then(|v| v.parse::<i64>().ok())
        .unwrap_or((spp_used * pixels.len() as i64) / (128 << 20));
    let n_dev = wanted.clamp(1, visible.max(1)) as i32;
    let params = ffi::RtbRenderParams { sample_begin: 0, sample_end: spp_used, pipeline: 0, collect_stats: 0 };
    // Color is #[repr(C)] {x, y, z: f64}: the pixel vector is the f64 rgb-sum buffer the ABI wants
    let px = pixels.as_mut_ptr() as *mut f64;
    let rc = if n_dev > 1 {
        unsafe { ffi::rtb_render_multi(&desc, n_dev, std::ptr::null(), &params, px, std::ptr::null_mut()) }
    } else {
        let mut scene: *mut ffi::rtb_scene = std::ptr::null_mut();
        let rc = unsafe { ffi::rtb_scene_create(&desc, 0, &mut scene) };
        if rc != 0 {
            panic!("rtb_scene_create: {}", last_error());
        }
        let rc = unsafe { ffi::rtb_render(scene, &params, px, std::ptr::null_mut()) };
        unsafe { ffi::rtb_scene_destroy(scene) };
        rc
    };
    if rc != 0 {
        panic!("render: {}", last_error());
    }
    eprintln!("\rWriting...            ");
    // src/render.rs:201-213: `if cam.auto_exposure { Some(auto_expose(cam, pixels)) } else { None }`, then write_color per pixel
    let mut exposure = 0.0f64; // <= 0: None
    if cam.auto_exposure && unsafe { ffi::rtb_auto_expose(px, pixels.len() as i64, spp_used as f64, &mut exposure) } != 0 {
        panic!("rtb_auto_expose: {}", last_error());
    }
    let mut rgb8 = vec![0u8; pixels.len() * 3];
    let rc = unsafe { ffi::rtb_write_color(std::ptr::null_mut(), px, pixels.len() as i64, spp_used as f64, exposure, rgb8.as_mut_ptr()) };
    if rc != 0 {
        panic!("rtb_write_color: {}", last_error());
    }
    let mut text = String::with_capacity(pixels.len() * 12);
    for p in rgb8.chunks(3) {
        text.push_str(&format!("{} {} {}\n", p[0], p[1], p[2]));
    }
    print!("{}", text);
    eprintln!("\rDone!                           ");
}
