// rtb/render.hpp -- drop-in replacements for the reference's render entry points
//     init_pixels        reference src/render.rs:136-138
//     render_par         reference src/render.rs:140-142
//     render_par_lights  reference src/render.rs:144-216
// with the same argument order.  The rayon loop (render.rs:171-197) is replaced by one call into
// librtb200.so (rtb_scene_create + rtb_render); the PPM emission (render.rs:151, 201-213) is kept,
// with write_color (color.rs:8-33) evaluated on the device through rtb_write_color.
// Errors: the reference panics; these throw std::runtime_error carrying rtb_last_error().
#pragma once

#include <iostream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "scene.hpp"

namespace rtb {

struct RenderOptions {
  int device = 0;
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

inline void render_par_lights(const Camera& cam, const HittableList& world, std::vector<Color>& pixels,
                              const std::vector<Sun>& /*suns: accepted and ignored, Q23*/,
                              const HittableList& lights, const RenderOptions& opt = default_render_options()) {
  if (opt.ppm_out) *opt.ppm_out << "P3\n" << cam.image_width << " " << cam.image_height() << "\n255\n";
  FlatScene flat(cam, world, &lights);
  flat.flags = opt.flags;
  flat.seed = opt.seed;
  RtbSceneDesc desc = flat.desc();
  rtb_scene* scene = nullptr;
  rtb_check(rtb_scene_create(&desc, opt.device, &scene), "rtb_scene_create");
  RtbSceneInfo info{};
  rtb_scene_info(scene, &info);
  if ((size_t)info.image_width * info.image_height != pixels.size()) {
    rtb_scene_destroy(scene);
    throw std::runtime_error("pixels has the wrong length (use init_pixels)");
  }
  if (opt.log) *opt.log << "Rendering on CUDA device " << info.device << "\n";
  RtbRenderParams params{};
  params.sample_begin = 0;
  params.sample_end = info.spp_used;
  params.pipeline = opt.pipeline;
  params.collect_stats = opt.stats ? 1 : 0;
  static_assert(sizeof(Color) == 3 * sizeof(double), "Color must be three packed f64");
  int rc = rtb_render(scene, &params, reinterpret_cast<double*>(pixels.data()), opt.stats);
  if (rc != RTB_OK) { rtb_scene_destroy(scene); rtb_check(rc, "rtb_render"); }
  if (opt.log) *opt.log << "\rWriting...            \n";
  if (opt.ppm_out) {
    std::vector<uint8_t> rgb8(pixels.size() * 3);
    rc = rtb_write_color(scene, reinterpret_cast<const double*>(pixels.data()), (int64_t)pixels.size(),
                         (double)info.spp_used, /*exposure: auto_exposure is out of scope*/ 0., rgb8.data());
    if (rc != RTB_OK) { rtb_scene_destroy(scene); rtb_check(rc, "rtb_write_color"); }
    std::string text;
    text.reserve(pixels.size() * 12);
    for (size_t i = 0; i < pixels.size(); i++) {
      text += std::to_string((int)rgb8[3 * i]); text += ' ';
      text += std::to_string((int)rgb8[3 * i + 1]); text += ' ';
      text += std::to_string((int)rgb8[3 * i + 2]); text += '\n';
    }
    *opt.ppm_out << text;
  }
  rtb_scene_destroy(scene);
  if (opt.log) *opt.log << "\rDone!                           \n";
}

inline void render_par(const Camera& cam, const HittableList& world, std::vector<Color>& pixels,
                       const std::vector<Sun>& suns, const RenderOptions& opt = default_render_options()) {
  // render.rs:140-142 forwards an EMPTY light list; the library then samples the material pdf
  // alone (HEAD would panic here).
  render_par_lights(cam, world, pixels, suns, HittableList::new_(), opt);
}

}  // namespace rtb
