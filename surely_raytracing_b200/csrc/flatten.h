// flatten.h -- host-side scene compiler: RtbSceneDesc (object graph, f64) -> flat device arrays.
//   * validates the description (indices, tree shape, depth);
//   * assigns canonical primitive ids in DFS `add` order;
//   * bakes Translate / RotateY chains into world-space primitives (src/transform.rs:57-135);
//   * derives the camera frame exactly like Camera::new (src/render.rs:62-134);
//   * builds one SAH BVH2 over all surface primitives (replaces BvhNode::new, src/hittable.rs:147-187,
//     whose randomised median split is not reproduced -- closest hits do not depend on tree shape).
#pragma once
#include <string>
#include <vector>

#include "../../include/rtb200.h"
#include "device_scene.h"

namespace rtb {

struct HostScene {
  std::vector<float4> nodes;        // 4 per inner node
  std::vector<uint4> qnodes;        // 2 per inner node (16-bit grid boxes), same indices as `nodes`
  double grid_base[3] = {0., 0., 0.}, grid_inv_cell[3] = {1., 1., 1.};
  float grid_cell[3] = {1.f, 1.f, 1.f};
  std::vector<float4> nodes4;       // 8 per BVH4 node (collapsed from `nodes`)
  int use_bvh4 = 0;
  int use_qnodes = 0;
  int defer_ok = 0;                 // non-solid textures are reached through Lambertian surface materials only
  int multi_leaf = 0;               // some BVH leaf holds more than one primitive
  int spec_bits = 0;                // SPEC_* features in use (device_scene.h)               // the grid inflates the boxes by < 3 % (surface-area measure): extend uses qnodes
  std::vector<double> prims;        // PRIM_DOUBLES per primitive (BVH order; surfaces, then boundaries)
  std::vector<int4> prim_info;
  std::vector<DPre> pre;            // one per primitive, same order
  float scene_mag = 1.f;
  std::vector<double2> xforms;
  std::vector<DMedium> media;
  std::vector<DMaterial> materials;
  std::vector<DTexture> textures;
  std::vector<uint8_t> texels;
  std::vector<float4> perlin_vec;
  std::vector<uint8_t> perlin_perm;
  std::vector<DLight> lights;
  DCamera cam;
  int n_suns = 0;
  DSun suns[MAX_SUNS] = {};
  int n_surface_prims = 0;
  int bvh_depth = 0;
  uint32_t flags = 0;
  uint64_t seed = 0;
};

// returns RTB_OK or a negative RTB_ERR_* with `err` set
int flatten_scene(const RtbSceneDesc& d, HostScene& out, std::string& err);

}  // namespace rtb
