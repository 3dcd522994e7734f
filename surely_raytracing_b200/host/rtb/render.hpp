// rtb/render.hpp -- drop-in replacements for the reference's render entry points
//     init_pixels        reference src/render.rs:136-138
//     render_par         reference src/render.rs:140-142
//     render_par_lights  reference src/render.rs:144-216
// with the same argument order.  The rayon loop (render.rs:171-197) is replaced by one call into
// librtb200.so (rtb_scene_create + rtb_render, or rtb_render_multi over the GPUs of the box); the PPM emission
// (render.rs:151, 201-213) is kept, with auto_expose (render.rs:325-339) through rtb_auto_expose and write_color
// (color.rs:8-33) evaluated on the device through rtb_write_color.
// Errors: the reference panics; these throw std::runtime_error carrying rtb_last_error().
#pragma once

#include <algorithm>
#include <cmath>
#include <iostream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "scene.hpp"

namespace rtb {

struct RenderOptions {
  int device = 0;
  // GPUs of this box to spread the stratum range over (rtb_render_multi: one thread + stream per GPU, one NCCL
  // sum-reduce): 1 = `device` only; 0 = as many of the visible GPUs as the job keeps busy (>= 128 M paths each)
  int n_devices = 0;
  std::ostream* ppm_out = &std::cout;  // nullptr: do not emit the P3 image
  std::ostream* log = &std::cerr;      // nullptr: quiet
  uint32_t flags = 0;                  // RTB_FLAG_*
  uint64_t seed = 20240001ull;
  int pipeline = RTB_PIPELINE_DEFAULT;
  RtbStats* stats = nullptr;
};
inline RenderOptions& default_render_options() { static RenderOptions o; return o; }

inline std::vector<Color> init_pixels(const Camera& cam) {
  return std::vector<Color>((size_t)cam.image_height() * cam.image_width, Color::new_zero());
}

inline void rtb_check(int rc, const char* what) {
  if (rc != RTB_OK) throw std::runtime_error(std::string(what) + ": " + rtb_last_error());
}

namespace detail {
inline void render_flat(const Camera& cam, FlatScene& flat, std::vector<Color>& pixels, const std::vector<Sun>& suns,
                        const RenderOptions& opt) {
  if (opt.ppm_out) *opt.ppm_out << "P3\n" << cam.image_width << " " << cam.image_height() << "\n255\n";  // render.rs:151
  flat.flags = opt.flags;
  flat.seed = opt.seed;
  flat.set_suns(suns);
  const RtbSceneDesc desc = flat.desc();
  const int root = (int)std::sqrt((double)cam.samples_per_pixel);
  const int64_t spp_used = (int64_t)root * root;  // nearest_square, render.rs:38-41, 108
  if ((size_t)cam.image_height() * cam.image_width != pixels.size()) throw std::runtime_error("pixels has the wrong length (use init_pixels)");
  int n_dev = opt.n_devices;
  if (n_dev <= 0) {
    const int64_t paths = spp_used * (int64_t)pixels.size();
    n_dev = (int)std::max<int64_t>(1, std::min<int64_t>(rtb_device_count(), paths / (128ll << 20)));
  }
  RtbRenderParams params{};
  params.sample_begin = 0;
  params.sample_end = spp_used;
  params.pipeline = opt.pipeline;
  params.collect_stats = opt.stats ? 1 : 0;
  static_assert(sizeof(Color) == 3 * sizeof(double), "Color must be three packed f64");
  double* px = reinterpret_cast<double*>(pixels.data());
  if (n_dev > 1) {
    if (opt.log) *opt.log << "Rendering on " << n_dev << " CUDA devices\n";
    std::vector<int> devs(n_dev);
    for (int k = 0; k < n_dev; k++) devs[k] = (opt.device + k) % rtb_device_count();
    rtb_check(rtb_render_multi(&desc, n_dev, devs.data(), &params, px, opt.stats), "rtb_render_multi");
  } else {
    rtb_scene* scene = nullptr;
    rtb_check(rtb_scene_create(&desc, opt.device, &scene), "rtb_scene_create");
    if (opt.log) *opt.log << "Rendering on CUDA device " << opt.device << "\n";
    const int rc = rtb_render(scene, &params, px, opt.stats);
    rtb_scene_destroy(scene);
    rtb_check(rc, "rtb_render");
  }
  if (opt.log) *opt.log << "\rWriting...            \n";
  if (opt.ppm_out) {
    // render.rs:201-213: `let exposure = if cam.auto_exposure { Some(auto_expose(..)) } else { None }`, then write_color per pixel
    double exposure = 0.;  // <= 0: None
    if (cam.auto_exposure) rtb_check(rtb_auto_expose(px, (int64_t)pixels.size(), (double)spp_used, &exposure), "rtb_auto_expose");
    std::vector<uint8_t> rgb8(pixels.size() * 3);
    // (cam.samples_per_pixel of the reference is the value Camera::new has already rounded to the square, render.rs:108-120)
    rtb_check(rtb_write_color(nullptr, px, (int64_t)pixels.size(), (double)spp_used, exposure, rgb8.data()), "rtb_write_color");
    std::string text;
    text.reserve(pixels.size() * 12);
    for (size_t i = 0; i < pixels.size(); i++) {
      text += std::to_string((int)rgb8[3 * i]); text += ' ';
      text += std::to_string((int)rgb8[3 * i + 1]); text += ' ';
      text += std::to_string((int)rgb8[3 * i + 2]); text += '\n';
    }
    *opt.ppm_out << text;
  }
  if (opt.log) *opt.log << "\rDone!                           \n";
}
}  // namespace detail

// render.rs:144-150: `lights: Arc<Object>` -- any object; a HittableList overload for the usual `Object::List(..)` call
inline void render_par_lights(const Camera& cam, const HittableList& world, std::vector<Color>& pixels, const std::vector<Sun>& suns,
                              const Object& lights, const RenderOptions& opt = default_render_options()) {
  FlatScene flat(cam, world, FlatScene::LightObject{lights});
  detail::render_flat(cam, flat, pixels, suns, opt);
}
inline void render_par_lights(const Camera& cam, const HittableList& world, std::vector<Color>& pixels, const std::vector<Sun>& suns,
                              const HittableList& lights, const RenderOptions& opt = default_render_options()) {
  FlatScene flat(cam, world, &lights);
  detail::render_flat(cam, flat, pixels, suns, opt);
}

inline void render_par(const Camera& cam, const HittableList& world, std::vector<Color>& pixels,
                       const std::vector<Sun>& suns, const RenderOptions& opt = default_render_options()) {
  // render.rs:140-142 forwards an EMPTY light list; the library then samples the material pdf
  // alone (HEAD would panic here).
  render_par_lights(cam, world, pixels, suns, HittableList::new_(), opt);
}

}  // namespace rtb
