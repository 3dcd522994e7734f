// example_cornell.cpp -- a scene written against the host API mirror exactly as the reference writes
// it (cornell_box, reference src/main.rs:417-512), rendered through the drop-in render_par_lights.
// Usage: example_cornell [width] [spp] > out.ppm     (PPM "P3" on stdout like `make run file=x`)
#include <cstdlib>

#include "rtb/render.hpp"

using namespace rtb;

int main(int argc, char** argv) {
  const int width = argc > 1 ? std::atoi(argv[1]) : 600;
  const int spp = argc > 2 ? std::atoi(argv[2]) : 1000;
  HittableList world = HittableList::new_();

  Material red = Lambertian::new_(Color::new_(0.65, 0.05, 0.05));
  Material white = Lambertian::new_(Color::new_(0.73, 0.73, 0.73));
  Material green = Lambertian::new_(Color::new_(0.12, 0.45, 0.15));
  Material light = DiffuseLight::new_(Color::new_(15., 15., 15.));

  world.add(Quad::new_(Point3::new_(555., 0., 0.), Vec3::new_(0., 555., 0.), Vec3::new_(0., 0., 555.), green));
  world.add(Quad::new_(Point3::new_(0., 0., 0.), Vec3::new_(0., 555., 0.), Vec3::new_(0., 0., 555.), red));
  world.add(Quad::new_(Point3::new_(343., 554., 332.), Vec3::new_(-130., 0., 0.), Vec3::new_(0., 0., -105.), light));
  world.add(Quad::new_(Point3::new_(0., 0., 0.), Vec3::new_(555., 0., 0.), Vec3::new_(0., 0., 555.), white));
  world.add(Quad::new_(Point3::new_(555., 555., 555.), Vec3::new_(-555., 0., 0.), Vec3::new_(0., 0., -555.), white));
  world.add(Quad::new_(Point3::new_(0., 0., 555.), Vec3::new_(555., 0., 0.), Vec3::new_(0., 555., 0.), white));

  Object box1 = make_box(Point3::new_zero(), Point3::new_(165., 330., 165.), white);
  box1 = RotateY::new_(box1, 15.);
  box1 = Translate::new_(box1, Vec3::new_(265., 0., 295.));
  world.add(box1);

  Material glass = Dielectric::new_clear(1.5);
  world.add(Sphere::new_(Point3::new_(190., 90., 190.), 90., glass));

  HittableList lights = HittableList::new_();
  lights.add(Quad::new_(Point3::new_(343., 554., 332.), Vec3::new_(-130., 0., 0.), Vec3::new_(0., 0., -105.), light));
  lights.add(Sphere::new_(Point3::new_(190., 90., 190.), 90., light));

  Camera cam = Camera::new_(1., width, spp, 50, 40., Point3::new_(278., 278., -800.), Point3::new_(278., 278., 0.),
                            Vec3::new_(0., 1., 0.), 0., 0., Color::new_zero());

  std::vector<Color> pixels = init_pixels(cam);
  try {
    render_par_lights(cam, world, pixels, {}, lights);
  } catch (const std::exception& e) {
    std::cerr << "render failed: " << e.what() << "\n";  // the reference would panic here
    return 1;
  }
  return 0;
}
