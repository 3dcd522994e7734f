// scenes.cpp -- the reference's scene functions (reference src/main.rs) rebuilt on the host API
// mirror (rtb/scene.hpp), exported as flat descriptions for tests/ and bench.py.
//
// Each function follows the object order and the literals of the cited main.rs function, because
// canonical primitive ids are defined by `add` order.  Scene-layout
// randomness (box heights, sphere centres, Perlin tables) comes from the seeded host stream and is
// drawn in the reference's order.  BASELINE.json's five configs are "c1".."c5".
#include <cstring>
#include <memory>
#include <string>

#include "rtb/scene.hpp"

using namespace rtb;

namespace {

thread_local std::string g_err;

struct Built {
  std::unique_ptr<FlatScene> flat;
  RtbSceneDesc desc;
};

enum { VARIANT_LIGHTS = 1u };  // pass the scene's light quad as `lights` (c3/c4/simple_light)

int pick(int asked, int dflt) { return asked > 0 ? asked : dflt; }

// A deterministic stand-in for earthmap.jpg, which the reference repo does not ship (F7):
// latitude colour bands + "continents" from a value-noise threshold + hashed speckle.
Texture synthetic_earth(uint64_t seed) {
  const int W = 1024, H = 512;
  std::vector<uint8_t> px((size_t)W * H * 3);
  auto hash = [&](uint32_t x, uint32_t y) {
    uint64_t h = seed ^ (0x9E3779B97F4A7C15ull * (x + 1)) ^ (0xC2B2AE3D27D4EB4Full * (y + 1));
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return (uint32_t)h;
  };
  auto lattice = [&](int x, int y) { return (hash((uint32_t)(x & 31), (uint32_t)(y & 15)) & 0xFFFF) / 65535.0; };
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      double fx = x / 32.0, fy = y / 32.0;
      int ix = (int)fx, iy = (int)fy;
      double tx = fx - ix, ty = fy - iy;
      tx = tx * tx * (3 - 2 * tx); ty = ty * ty * (3 - 2 * ty);
      double n = (1 - tx) * (1 - ty) * lattice(ix, iy) + tx * (1 - ty) * lattice(ix + 1, iy) +
                 (1 - tx) * ty * lattice(ix, iy + 1) + tx * ty * lattice(ix + 1, iy + 1);
      double lat = std::fabs((y + 0.5) / H - 0.5) * 2.0;  // 0 equator .. 1 pole
      double r, g, b;
      if (lat > 0.88) { r = g = b = 235; }
      else if (n > 0.55) { r = 60 + 120 * lat; g = 140 - 40 * lat; b = 50; }
      else { r = 20; g = 60 + 40 * n; b = 150 + 60 * (1 - lat); }
      int speck = (int)(hash((uint32_t)x + 7919u, (uint32_t)y + 104729u) & 15) - 8;
      auto cl = [](double v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); };
      uint8_t* p = &px[3 * ((size_t)y * W + x)];
      p[0] = cl(r + speck); p[1] = cl(g + speck); p[2] = cl(b + speck);
    }
  return ImageTexture::from_rgb8(W, H, std::move(px));
}

// ---- c1: scene_random_balls  (main.rs:135-210) ---------------------------------------------------
FlatScene* scene_random_balls(int width, int spp, int depth) {
  HittableList world;
  Texture checker = CheckerTexture::from_color(0.32, Color(0.2, 0.3, 0.1), Color(0.9, 0.9, 0.9));
  world.add(Sphere::new_(Point3(0., -2000., 0.), 2000., Lambertian::from_texture(checker)));
  for (int ia = -11; ia < 11; ia++)
    for (int ib = -11; ib < 11; ib++) {
      double a = ia, b = ib;
      double choose_mat = random_double();
      double cx = a + 0.9 * random_double();
      double cz = b + 0.9 * random_double();
      Point3 center(cx, 0.2, cz);
      Point3 center2 = center + Vec3(0., random_range(0., 0.5), 0.);
      if ((center - Point3(4., 0.2, 0.)).length_squared() > 0.9 * 0.9) {
        if (choose_mat < 0.8) {
          Vec3 a1 = random_vec3();
          Vec3 a2 = random_vec3();
          world.add(Sphere::new_moving(center, center2, 0.2, Lambertian::new_(a1 * a2)));
        } else if (choose_mat < 0.95) {
          Color albedo = random_vec3_range(0.5, 1.);
          double fuzz = random_range(0., 0.5);
          world.add(Sphere::new_(center, 0.2, Metal::new_(albedo, fuzz)));
        } else {
          double ir = random_range(1.2, 1.6);
          world.add(Sphere::new_(center, 0.2, Dielectric::new_(ir, Color(1., 1., 1.))));
        }
      }
    }
  world.add(Sphere::new_(Point3(0., 1., 0.), 1.0, Dielectric::new_(1.5, Color(1., 1., 1.))));
  world.add(Sphere::new_(Point3(-4., 1., 0.), 1.0, Lambertian::new_(Color(0.4, 0.2, 0.1))));
  world.add(Sphere::new_(Point3(4., 1., 0.), 1.0, Metal::new_(Color(0.7, 0.6, 0.5), 0.0)));
  Camera cam = Camera::new_(16. / 9., pick(width, 400), pick(spp, 50), pick(depth, 50), 20., Point3(13., 2., 3.),
                            Point3(0., 0., 0.), Vec3(0., 1., 0.), 0.6, 10., Color(0.7, 0.8, 1.));
  HittableList accel = world.create_bvh();
  return new FlatScene(cam, accel, nullptr);
}

// ---- shared Cornell shell (main.rs:420-460 / 517-557) ----------------------------------------------
void cornell_walls(HittableList& world, const Material& light, Point3 lq, Vec3 lu, Vec3 lv) {
  Material red = Lambertian::new_(Color(0.65, 0.05, 0.05));
  Material white = Lambertian::new_(Color(0.73, 0.73, 0.73));
  Material green = Lambertian::new_(Color(0.12, 0.45, 0.15));
  world.add(Quad::new_(Point3(555., 0., 0.), Vec3(0., 555., 0.), Vec3(0., 0., 555.), green));
  world.add(Quad::new_(Point3(0., 0., 0.), Vec3(0., 555., 0.), Vec3(0., 0., 555.), red));
  world.add(Quad::new_(lq, lu, lv, light));
  world.add(Quad::new_(Point3(0., 0., 0.), Vec3(555., 0., 0.), Vec3(0., 0., 555.), white));
  world.add(Quad::new_(Point3(555., 555., 555.), Vec3(-555., 0., 0.), Vec3(0., 0., -555.), white));
  world.add(Quad::new_(Point3(0., 0., 555.), Vec3(555., 0., 0.), Vec3(0., 555., 0.), white));
}
Object cornell_box1(const Material& white) {
  Object b = make_box(Point3::new_zero(), Point3(165., 330., 165.), white);
  return Translate::new_(RotateY::new_(b, 15.), Vec3(265., 0., 295.));
}
Object cornell_box2(const Material& white) {
  Object b = make_box(Point3::new_zero(), Point3(165., 165., 165.), white);
  return Translate::new_(RotateY::new_(b, -18.), Vec3(130., 0., 65.));
}
Camera cornell_camera(int width, int spp, int depth) {
  return Camera::new_(1., width, spp, depth, 40., Point3(278., 278., -800.), Point3(278., 278., 0.), Vec3(0., 1., 0.),
                      0., 0., Color::new_zero());
}

// ---- c5: cornell_box as at HEAD (main.rs:417-512); c2: both boxes, no sphere, lights=[quad] --------
FlatScene* cornell_box(int width, int spp, int depth, bool head) {
  HittableList world;
  Material white = Lambertian::new_(Color(0.73, 0.73, 0.73));
  Material light = DiffuseLight::new_(Color(15., 15., 15.));
  Point3 lq(343., 554., 332.);
  Vec3 lu(-130., 0., 0.), lv(0., 0., -105.);
  cornell_walls(world, light, lq, lu, lv);
  world.add(cornell_box1(white));
  HittableList lights;
  lights.add(Quad::new_(lq, lu, lv, light));
  if (head) {
    world.add(Sphere::new_(Point3(190., 90., 190.), 90., Dielectric::new_clear(1.5)));
    lights.add(Sphere::new_(Point3(190., 90., 190.), 90., light));
  } else {
    world.add(cornell_box2(white));
  }
  Camera cam = cornell_camera(pick(width, 600), pick(spp, 1000), pick(depth, 50));
  return new FlatScene(cam, world, &lights);
}

// ---- c3: cornell_smoke (main.rs:514-601) -------------------------------------------------------------
FlatScene* cornell_smoke(int width, int spp, int depth, uint32_t variant) {
  HittableList world;
  Material white = Lambertian::new_(Color(0.73, 0.73, 0.73));
  Material light = DiffuseLight::new_(Color(7., 7., 7.));
  Point3 lq(113., 554., 127.);
  Vec3 lu(330., 0., 0.), lv(0., 0., 305.);
  cornell_walls(world, light, lq, lu, lv);
  world.add(ConstantMedium::new_(cornell_box1(white), 0.01, Color::new_zero()));
  world.add(ConstantMedium::new_(cornell_box2(white), 0.01, Color(1., 1., 1.)));
  HittableList lights;
  if (variant & VARIANT_LIGHTS) lights.add(Quad::new_(lq, lu, lv, light));
  Camera cam = cornell_camera(pick(width, 600), pick(spp, 2000), pick(depth, 10));
  return new FlatScene(cam, world, &lights);
}

// ---- c4: final_scene (main.rs:603-712) -----------------------------------------------------------------
FlatScene* final_scene(int width, int spp, int depth, uint32_t variant, uint64_t seed) {
  HittableList boxes1;
  Material ground = Lambertian::new_(Color(0.48, 0.83, 0.53));
  const int boxes_per_side = 20;
  for (int i = 0; i < boxes_per_side; i++)
    for (int j = 0; j < boxes_per_side; j++) {
      double w = 100.;
      double x0 = -1000. + i * w, z0 = -1000. + j * w, y0 = 0.;
      double x1 = x0 + w, y1 = random_range(1., 101.), z1 = z0 + w;
      boxes1.add(make_box(Point3(x0, y0, z0), Point3(x1, y1, z1), ground));
    }
  HittableList world;
  world.add(ObjectList(boxes1.create_bvh()));
  Material light = DiffuseLight::new_(Color(7., 7., 7.));
  Point3 lq(123., 554., 147.);
  Vec3 lu(300., 0., 0.), lv(0., 0., 265.);
  world.add(Quad::new_(lq, lu, lv, light));
  Point3 center1(400., 400., 200.);
  Point3 center2 = center1 + Vec3(30., 0., 0.);
  world.add(Sphere::new_moving(center1, center2, 50., Lambertian::new_(Color(0.7, 0.3, 0.1))));
  world.add(Sphere::new_(Point3(260., 150., 45.), 50., Dielectric::new_clear(1.5)));
  world.add(Sphere::new_(Point3(0., 150., 145.), 50., Metal::new_(Color(0.8, 0.8, 0.9), 1.0)));
  world.add(Sphere::new_(Point3(360., 150., 145.), 70., Dielectric::new_clear(1.5)));
  world.add(ConstantMedium::new_(Sphere::new_(Point3(360., 150., 145.), 70., Dielectric::new_clear(1.5)), 0.2,
                                 Color(0.2, 0.4, 0.9)));
  world.add(ConstantMedium::new_(Sphere::new_(Point3::new_zero(), 5000., Dielectric::new_clear(1.5)), 0.0001,
                                 Color(1., 1., 1.)));
  world.add(Sphere::new_(Point3(400., 200., 400.), 100., Lambertian::from_texture(synthetic_earth(seed))));
  Texture pertext = NoiseTexture::new_(0.1);
  world.add(Sphere::new_(Point3(220., 280., 300.), 80., Lambertian::from_texture(pertext)));
  HittableList boxes2;
  Material white = Lambertian::new_(Color(0.73, 0.73, 0.73));
  for (int k = 0; k < 1000; k++) boxes2.add(Sphere::new_(random_vec3_range(0., 165.), 10., white));
  world.add(Translate::new_(RotateY::new_(ObjectList(boxes2.create_bvh()), 15.), Vec3(-100., 270., 395.)));
  HittableList lights;
  if (variant & VARIANT_LIGHTS) lights.add(Quad::new_(lq, lu, lv, light));
  Camera cam = Camera::new_(1.0, pick(width, 800), pick(spp, 10000), pick(depth, 40), 40., Point3(478., 278., -600.),
                            Point3(278., 278., 0.), Vec3(0., 1., 0.), 0., 0., Color::new_zero());
  return new FlatScene(cam, world, &lights);
}

// ---- remaining main.rs scenes ----------------------------------------------------------
FlatScene* scene_three_spheres(int width, int spp, int depth) {  // main.rs:92-133
  HittableList world;
  Material left = Dielectric::new_(1.5, Color(1.0, 0.9, 0.8));
  world.add(Sphere::new_(Point3(0., 0., -1.), 0.5, Lambertian::new_(Color(0.1, 0.2, 0.5))));
  world.add(Sphere::new_(Point3(-1., 0., -1.), 0.5, left));
  world.add(Sphere::new_(Point3(-1., 0., -1.), -0.4, left));
  world.add(Sphere::new_(Point3(0., -100.5, -1.), 100., Lambertian::new_(Color(0.8, 0.8, 0.0))));
  world.add(Sphere::new_(Point3(1., 0., -1.), 0.5, Metal::new_(Color(0.8, 0.6, 0.2), 0.)));
  Camera cam = Camera::new_(16. / 9., pick(width, 800), pick(spp, 1000), pick(depth, 50), 90., Point3(0., 0., 0.),
                            Point3(0., 0., -1.), Vec3(0., 1., 0.), 2., 1., Color(0.7, 0.8, 1.));
  return new FlatScene(cam, world.create_bvh(), nullptr);
}
FlatScene* scene_sun_spheres(int width, int spp, int depth) {  // main.rs:32-90 (scene -2: suns + auto_exposure)
  HittableList world;
  Material left = Dielectric::new_(1.5, Color(1., 1., 1.));
  world.add(Sphere::new_(Point3(0., 0., -1.), 0.5, Lambertian::new_(Color(0.1, 0.2, 0.5))));
  world.add(Sphere::new_(Point3(-1., 0., -1.25), 0.5, left));
  world.add(Sphere::new_(Point3(-1., 0., -1.25), -0.4, left));
  world.add(Sphere::new_(Point3(0., -100.5, -1.), 100., Lambertian::new_(Color(0.8, 0.8, 0.0))));
  world.add(Sphere::new_(Point3(1., 0., -0.75), 0.5, Metal::new_(Color(0.8, 0.6, 0.2), 0.)));
  Camera cam = Camera::new_(16. / 9., pick(width, 640), pick(spp, 1000), pick(depth, 50), 90., Point3(0., 0., 0.),
                            Point3(0., 0., -1.), Vec3(0., 1., 0.), 0., 1., Color(0.02, 0.05, 0.1));
  cam.auto_exposure = true;
  FlatScene* f = new FlatScene(cam, world, nullptr);
  f->set_suns({Sun::new_(Vec3(-1., 1., 1.), Color(1., 1., 1.) * 10., 2.)});
  return f;
}
FlatScene* two_spheres(int width, int spp, int depth) {  // main.rs:212-250
  HittableList world;
  Texture checker = CheckerTexture::from_color(0.3, Color(0.2, 0.3, 0.1), Color(0.9, 0.9, 0.9));
  world.add(Sphere::new_(Point3(0., -10., 0.), 10., Lambertian::from_texture(checker)));
  world.add(Sphere::new_(Point3(0., 10., 0.), 10., Lambertian::from_texture(checker)));
  Camera cam = Camera::new_(16. / 9., pick(width, 400), pick(spp, 100), pick(depth, 50), 20., Point3(13., 2., 3.),
                            Point3(0., 0., 0.), Vec3(0., 1., 0.), 0., 0., Color(0.7, 0.8, 1.));
  return new FlatScene(cam, world, nullptr);
}
FlatScene* earth(int width, int spp, int depth, uint64_t seed) {  // main.rs:252-277
  Object globe = Sphere::new_(Point3::new_zero(), 2., Lambertian::from_texture(synthetic_earth(seed)));
  Camera cam = Camera::new_(16. / 9., pick(width, 1000), pick(spp, 1000), pick(depth, 50), 20., Point3(13., 3., 2.),
                            Point3(0., 0., 0.), Vec3(0., 1., 0.), 0., 0., Color(0.7, 0.8, 1.));
  return new FlatScene(cam, HittableList::from_object(globe), nullptr);
}
FlatScene* two_perlin_spheres(int width, int spp, int depth) {  // main.rs:279-313
  HittableList world;
  Texture pertext = NoiseTexture::new_(4.);
  world.add(Sphere::new_(Point3(0., -1000., 0.), 1000., Lambertian::from_texture(pertext)));
  world.add(Sphere::new_(Point3(0., 2., 0.), 2., Lambertian::from_texture(pertext)));
  Camera cam = Camera::new_(16. / 9., pick(width, 400), pick(spp, 100), pick(depth, 50), 20., Point3(13., 2., 3.),
                            Point3(0., 0., 0.), Vec3(0., 1., 0.), 0., 0., Color(0.6, 0.7, 1.));
  return new FlatScene(cam, world, nullptr);
}
FlatScene* quads(int width, int spp, int depth) {  // main.rs:315-372
  HittableList world;
  world.add(Quad::new_(Point3(-3., -2., 5.), Vec3(0., 0., -4.), Vec3(0., 4., 0.), Lambertian::new_(Color(1., 0.2, 0.2))));
  world.add(Quad::new_(Point3(-2., -2., 0.), Vec3(4., 0., 0.), Vec3(0., 4., 0.), Lambertian::new_(Color(0.2, 1.0, 0.2))));
  world.add(Quad::new_(Point3(3., -2., 1.), Vec3(0., 0., 4.), Vec3(0., 4., 0.), Lambertian::new_(Color(0.2, 0.2, 1.0))));
  world.add(Quad::new_(Point3(-2., 3., 1.), Vec3(4., 0., 0.), Vec3(0., 0., 4.), Lambertian::new_(Color(1.0, 0.5, 0.))));
  world.add(Quad::new_(Point3(-2., -3., 5.), Vec3(4., 0., 0.), Vec3(0., 0., -4.), Lambertian::new_(Color(0.2, 0.8, 0.8))));
  Camera cam = Camera::new_(1.0, pick(width, 400), pick(spp, 100), pick(depth, 50), 80., Point3(0., 0., 9.),
                            Point3(0., 0., 0.), Vec3(0., 1., 0.), 0., 0., Color(0.6, 0.7, 1.));
  return new FlatScene(cam, world, nullptr);
}
FlatScene* simple_light(int width, int spp, int depth, uint32_t variant) {  // main.rs:374-415
  HittableList world;
  Texture pertex = NoiseTexture::new_(4.);
  world.add(Sphere::new_(Point3(0., -1000., 0.), 1000., Lambertian::from_texture(pertex)));
  world.add(Sphere::new_(Point3(0., 2., 0.), 2., Lambertian::from_texture(pertex)));
  Material difflight = DiffuseLight::new_(Color(4., 4., 4.));
  world.add(Quad::new_(Point3(3., 1., -2.), Vec3(2., 0., 0.), Vec3(0., 2., 0.), difflight));
  world.add(Sphere::new_(Point3(0., 7., 0.), 2., difflight));
  HittableList lights;
  if (variant & VARIANT_LIGHTS) {
    lights.add(Quad::new_(Point3(3., 1., -2.), Vec3(2., 0., 0.), Vec3(0., 2., 0.), difflight));
    lights.add(Sphere::new_(Point3(0., 7., 0.), 2., difflight));
  }
  Camera cam = Camera::new_(16. / 9., pick(width, 400), pick(spp, 400), pick(depth, 50), 20., Point3(26., 3., 6.),
                            Point3(0., 2., 0.), Vec3(0., 1., 0.), 0., 0., Color::new_zero());
  return new FlatScene(cam, world, &lights);
}

// A closed white-furnace test scene (not in the reference): a Lambertian sphere of albedo `a`
// inside a uniformly emitting background; every pixel must converge to the analytic value.
FlatScene* furnace(int width, int spp, int depth) {
  HittableList world;
  world.add(Sphere::new_(Point3(0., 0., 0.), 1., Lambertian::new_(Color(0.5, 0.5, 0.5))));
  Camera cam = Camera::new_(1., pick(width, 64), pick(spp, 256), pick(depth, 50), 40., Point3(0., 0., 4.),
                            Point3(0., 0., 0.), Vec3(0., 1., 0.), 0., 0., Color(1., 1., 1.));
  return new FlatScene(cam, world, nullptr);
}

// Not a reference scene: a stress case for the box leaves of the traversal (csrc/rtb_device.cuh, prefilter_box) built with the
// reference's own make_box -- boxes that share whole faces and edges, a box inside a box, spheres inside and through boxes, a
// translated box (its baked corners need not agree bit for bit: it must quietly stay six leaves), a rotated one, the camera
// INSIDE a large glass box, a light box.
FlatScene* box_city(int width, int spp, int depth) {
  HittableList world;
  Material grey = Lambertian::new_(Color(0.6, 0.6, 0.6)), red = Lambertian::new_(Color(0.7, 0.2, 0.2));
  Material glass = Dielectric::new_(1.5, Color(1., 1., 1.)), steel = Metal::new_(Color(0.8, 0.8, 0.9), 0.1);
  Material lamp = DiffuseLight::new_(Color(6., 6., 5.));
  HittableList blocks;
  for (int i = -3; i < 3; i++)
    for (int k = -3; k < 3; k++) {  // a 6 x 6 block of towers that touch along whole faces, heights on a fixed pattern
      const double h = 1. + (double)((i * 7 + k * 13 + 80) % 5);
      blocks.add(make_box(Point3(2. * i, 0., 2. * k), Point3(2. * i + 2., h, 2. * k + 2.), ((i + k) & 1) ? grey : red));
    }
  world.add(ObjectList(blocks.create_bvh()));
  world.add(make_box(Point3(-6., -1., -6.), Point3(6., 0., 6.), grey));                     // a slab under all of them (shared plane y = 0)
  world.add(make_box(Point3(-0.5, 5.5, -0.5), Point3(0.5, 6.5, 0.5), lamp));                // a light box above the tallest tower
  world.add(make_box(Point3(-9., 0., -9.), Point3(9., 12., 9.), glass));                    // everything, camera included, inside a glass box
  world.add(make_box(Point3(6.5, 0., -1.), Point3(8.5, 2., 1.), steel));
  world.add(make_box(Point3(7., 0.5, -0.5), Point3(8., 1.5, 0.5), red));                     // a box inside a box
  world.add(Sphere::new_(Point3(7.5, 1., 0.), 0.3, lamp));                                  // a lamp inside both
  world.add(Sphere::new_(Point3(-7.5, 1., 0.), 1.2, glass));
  world.add(make_box(Point3(-8., 0., -0.5), Point3(-7., 1., 0.5), steel));                   // a box through that sphere
  world.add(Translate::new_(make_box(Point3(0., 0., 0.), Point3(1.1, 0.7, 1.3), red), Vec3(0.1, 6.2, 7.3)));
  world.add(Translate::new_(RotateY::new_(make_box(Point3(0., 0., 0.), Point3(1., 2., 1.), grey), 30.), Vec3(-7.5, 0., 6.5)));
  Camera cam = Camera::new_(1., pick(width, 200), pick(spp, 64), pick(depth, 12), 70., Point3(5.2, 7.5, 8.1),
                            Point3(0., 2., 0.), Vec3(0., 1., 0.), 0., 10., Color(0.3, 0.4, 0.6));
  return new FlatScene(cam, world, nullptr);
}

}  // namespace

extern "C" {

const char* rtbs_last_error(void) { return g_err.c_str(); }

// name: c1..c5 or a main.rs scene name; width/spp/depth <= 0 pick the config's own value.
void* rtbs_build(const char* name, int width, int spp, int depth, uint64_t seed, uint32_t variant, uint32_t flags) {
  try {
    seed_host_rng(seed);
    std::string n(name ? name : "");
    FlatScene* f = nullptr;
    if (n == "c1" || n == "scene_random_balls") f = scene_random_balls(width, spp, depth);
    else if (n == "c2" || n == "mixed_pdf") f = cornell_box(width, spp, depth, false);
    else if (n == "c3" || n == "cornell_smoke") f = cornell_smoke(width, spp, depth, variant);
    else if (n == "c4" || n == "final_scene") f = final_scene(width, spp, depth, variant, seed);
    else if (n == "c5" || n == "cornell_box") f = cornell_box(width, spp, depth, true);
    else if (n == "scene_three_spheres") f = scene_three_spheres(width, spp, depth);
    else if (n == "scene_sun_spheres") f = scene_sun_spheres(width, spp, depth);
    else if (n == "two_spheres") f = two_spheres(width, spp, depth);
    else if (n == "earth") f = earth(width, spp, depth, seed);
    else if (n == "two_perlin_spheres") f = two_perlin_spheres(width, spp, depth);
    else if (n == "quads") f = quads(width, spp, depth);
    else if (n == "simple_light") f = simple_light(width, spp, depth, variant);
    else if (n == "furnace") f = furnace(width, spp, depth);
    else if (n == "box_city") f = box_city(width, spp, depth);
    else { g_err = "unknown scene '" + n + "'"; return nullptr; }
    f->seed = seed;
    f->flags = flags;
    Built* b = new Built();
    b->flat.reset(f);
    b->desc = f->desc();
    return b;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

RtbSceneDesc* rtbs_desc(void* h) { return h ? &static_cast<Built*>(h)->desc : nullptr; }
void rtbs_free(void* h) { delete static_cast<Built*>(h); }

}  // extern "C"
