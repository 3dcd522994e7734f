/*
 * rtb200.h -- C ABI of librtb200.so, the B200 (sm_100a) backend for the per-pixel
 * integration hot path of carlosconley/surely-raytracing.
 *
 * The reference has no FFI of its own: the seam this library replaces is the
 * Rust function pair
 *     render_par        (cam, world, pixels, suns)          reference src/render.rs:140
 *     render_par_lights (cam, world, pixels, suns, lights)  reference src/render.rs:144-216
 * i.e. the rayon loop `pixels[idx] += ray_color(get_ray(..))` (src/render.rs:179-197).
 * A host crate keeps the reference's scene-construction API and flattens the object graph
 * (enum Object, src/object.rs:18-26) into the plain arrays below; INTEGRATION.md shows the
 * Rust `extern "C"` block a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative RTB_ERR_* otherwise; the message of the
 *     most recent failure on the calling thread is available from rtb_last_error()
 *     (the reference panics instead: src/hittable.rs:121, src/transform.rs:45,148);
 *   - all input arrays are borrowed for the duration of the call and copied;
 *   - no torch / C++ types cross this boundary: plain pointers, sizes and PODs only;
 *   - scene quantities are f64 exactly as the reference holds them (everything there is f64).
 *
 * The same RtbSceneDesc is consumed by the test oracle (oracle/oracle.cpp), which is NOT part of
 * this library.
 */
#ifndef RTB200_H
#define RTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_ABI_VERSION 2

/* ---- error codes ------------------------------------------------------------------------- */
#define RTB_OK 0
#define RTB_ERR_INVALID (-1)     /* malformed description (bad index, cycle, empty world, ...)   */
#define RTB_ERR_UNSUPPORTED (-2) /* legal for the reference but outside this backend            */
#define RTB_ERR_CUDA (-3)        /* CUDA runtime failure; rtb_last_error() carries the string   */
#define RTB_ERR_NO_DEVICE (-4)   /* no sm_100 device visible: there is no CPU fallback          */

/* ---- object graph: mirrors `enum Object` (reference src/object.rs:18-26) ----------------- */
enum RtbObjKind {
  RTB_OBJ_SPHERE = 0,    /* Sphere::new / new_moving      src/object.rs:83-105                  */
  RTB_OBJ_QUAD = 1,      /* Quad::new                     src/object.rs:428-445                 */
  RTB_OBJ_LIST = 2,      /* Object::List(HittableList)    src/hittable.rs:55-85                 */
  RTB_OBJ_BVH = 3,       /* Object::Node from create_bvh  src/hittable.rs:82-84 (a list marker) */
  RTB_OBJ_TRANSLATE = 4, /* Translate::new                src/transform.rs:43-53                */
  RTB_OBJ_ROTATE_Y = 5,  /* RotateY::new                  src/transform.rs:143-186              */
  RTB_OBJ_MEDIUM = 6     /* ConstantMedium::new           src/constant_medium.rs:23-29          */
};

typedef struct RtbObject {
  int32_t kind;     /* RtbObjKind                                                               */
  int32_t material; /* SPHERE/QUAD: material index; MEDIUM: its Isotropic material; else -1     */
  int32_t first;    /* LIST/BVH: first slot in RtbSceneDesc.children;                           */
                    /* TRANSLATE/ROTATE_Y/MEDIUM: index of the wrapped object                   */
  int32_t count;    /* LIST/BVH: number of children (insertion order = `add` order)             */
  double v[10];     /* SPHERE : center[3], radius, center_vec[3], moving(0/1)                   */
                    /* QUAD   : q[3], u[3], v[3]                                                */
                    /* TRANSLATE: offset[3]     ROTATE_Y: angle in degrees     MEDIUM: density  */
} RtbObject;

/* ---- materials: `enum Material` (reference src/material.rs:26-32) ------------------------ */
enum RtbMatKind {
  RTB_MAT_LAMBERTIAN = 0,    /* texture                      src/material.rs:75-110             */
  RTB_MAT_METAL = 1,         /* color = albedo, param = fuzz (<=1)   src/material.rs:112-139    */
  RTB_MAT_DIELECTRIC = 2,    /* color = tint,   param = ir           src/material.rs:141-192    */
  RTB_MAT_DIFFUSE_LIGHT = 3, /* texture = emit               src/material.rs:194-222            */
  RTB_MAT_ISOTROPIC = 4      /* texture = albedo             src/material.rs:224-248            */
};

typedef struct RtbMaterial {
  int32_t kind;    /* RtbMatKind */
  int32_t texture; /* texture index or -1 */
  double color[3];
  double param;
} RtbMaterial;

/* ---- textures: `enum Texture` (reference src/texture.rs:10-15) --------------------------- */
enum RtbTexKind {
  RTB_TEX_SOLID = 0,   /* color                                        src/texture.rs:28-46     */
  RTB_TEX_CHECKER = 1, /* scale = inv_scale, a = even tex, b = odd tex src/texture.rs:48-82     */
  RTB_TEX_IMAGE = 2,   /* a = image index                              src/texture.rs:84-108    */
  RTB_TEX_NOISE = 3    /* a = perlin index, scale                      src/texture.rs:110-131   */
};

typedef struct RtbTexture {
  int32_t kind; /* RtbTexKind */
  int32_t a;
  int32_t b;
  int32_t reserved;
  double color[3];
  double scale;
} RtbTexture;

/* RGB8, row-major, top row first: what `image::open(..).to_rgb8()` yields (src/rt_image.rs:13-27) */
typedef struct RtbImage {
  int32_t width;
  int32_t height;
  const uint8_t* rgb;
} RtbImage;

/* Perlin tables (reference src/perlin.rs:7-12); generated on the host (src/perlin.rs:15-28) */
typedef struct RtbPerlin {
  double ranvec[256][3];
  int32_t perm_x[256];
  int32_t perm_y[256];
  int32_t perm_z[256];
} RtbPerlin;

/* Arguments of Camera::new (reference src/render.rs:62-74); the derived frame is recomputed
 * inside the library exactly as src/render.rs:75-133 does (incl. spp -> nearest square). */
typedef struct RtbCamera {
  double aspect_ratio;
  int32_t image_width;
  int32_t samples_per_pixel;
  int32_t max_depth;
  int32_t reserved;
  double vfov;
  double lookfrom[3];
  double lookat[3];
  double vup[3];
  double defocus_angle;
  double focus_dist;
  double background[3];
} RtbCamera;

/* flags */
#define RTB_FLAG_ISO_PDF_ZERO 1u /* HEAD-literal Isotropic::scattering_pdf == 0;    */
                                 /* default (flag clear) is the intended 1/(4*pi)               */
#define RTB_FLAG_PROPAGATE_NAN 2u /* HEAD-literal: non-finite samples poison the pixel (Q22);   */
                                  /* default: a non-finite / zero-pdf sample contributes 0      */

/* builder arms of the device tree (measured alternatives to the default 64-byte BVH2 with one-primitive leaves) */
#define RTB_FLAG_BVH4 0x10u        /* traverse the tree with every other level collapsed (4-wide nodes)          */
#define RTB_FLAG_QNODES 0x20u      /* 32-byte nodes with 16-bit quantised boxes (vetoed where the grid is coarse) */
#define RTB_FLAG_BVH_LEAF4 0x40u   /* allow leaves of up to 4 primitives                                          */
#define RTB_FLAG_NO_BOX_SCAN 0x80u /* quad-bounded media: two boundary probes as written instead of one scan      */
#define RTB_FLAG_NO_BOX_LEAVES 0x400u /* axis-aligned make_box lists stay six one-quad leaves (A/B arm)          */
/* The sun term of ray_color is commented out at HEAD (src/render.rs:300-308: `cam.background //+ sun_light`), so
 * `suns` is accepted and ignored (Q23).  This flag switches the term back on: a miss adds, for every sun,
 * Sun::_hit (src/object.rs:232-239): albedo if dot(unit(d), direction) > 1 - angular_diameter/180. */
#define RTB_FLAG_SUN_LIGHT 0x100u
/* Opt-in Russian roulette (NOT in the reference, whose only terminations are max_depth, a miss and an emitter): from the
 * 4th bounce on a path survives with probability q = clamp(max(throughput), 0.05, 1) and its throughput is divided by q.
 * Unbiased, fewer segments per path, more variance per path; the default stays the reference's estimator. */
#define RTB_FLAG_RUSSIAN_ROULETTE 0x200u

/* Sun::new (reference src/object.rs:223-231) */
typedef struct RtbSun {
  double direction[3]; /* as given; normalised by the library like Sun::new does */
  double albedo[3];
  double angular_diameter;
} RtbSun;

typedef struct RtbSceneDesc {
  int32_t abi_version; /* RTB_ABI_VERSION */
  uint32_t flags;
  uint64_t seed;

  const RtbObject* objects;
  int32_t n_objects;
  const int32_t* children; /* child object indices of every LIST/BVH, concatenated */
  int32_t n_children;
  int32_t world; /* index of the root LIST (the `world: &HittableList` argument) */

  const int32_t* lights; /* objects of the `lights` list, in order; n_lights == 0 reproduces */
  int32_t n_lights;      /* render_par's empty list: the material pdf is used alone (F2)     */

  const RtbMaterial* materials;
  int32_t n_materials;
  const RtbTexture* textures;
  int32_t n_textures;
  const RtbImage* images;
  int32_t n_images;
  const RtbPerlin* perlins;
  int32_t n_perlins;

  RtbCamera camera;

  const RtbSun* suns; /* the `suns: &Vec<Sun>` argument of render_par (src/render.rs:140); see RTB_FLAG_SUN_LIGHT */
  int32_t n_suns;
  int32_t reserved;
} RtbSceneDesc;

/* ---- deterministic-parity harness types -------------------------------------------------- */
typedef struct RtbRay {
  double origin[3];
  double direction[3]; /* not normalised (Q3) */
  double time;
  double t_min; /* 1e-4 for radiance rays (src/render.rs:267) */
} RtbRay;

/* canonical primitive id: DFS order of the leaf spheres/quads reachable from `world`, as added,
 * recursing through LIST/BVH/TRANSLATE/ROTATE_Y and MEDIUM boundaries.  -1 = miss. */
typedef struct RtbHit {
  int32_t prim;
  int32_t front_face;
  int32_t material;
  int32_t reserved;
  double t;
  double p[3];
  double normal[3];
  double u, v;
} RtbHit;

#define RTB_TRACE_BRUTE_FORCE 1u /* linear scan over all primitives instead of the BVH */
/* Trace through the kernels rtb_render itself runs: the rays are packed into a ray queue, traversed by the
 * wavefront extend kernel (conservative classification, <= 2 candidates), re-traced exactly where the candidates
 * overflowed, and resolved by the exact tests as the shade stage does.  t_min must be 1e-4 (radiance rays).
 * Default: primary records (f64 directions, what get_ray produces).  With RTB_TRACE_SECONDARY the rays travel
 * as scattered-ray records: the library rounds their directions and time to fp32, as the shade stage's output is. */
#define RTB_TRACE_WAVEFRONT 2u
#define RTB_TRACE_SECONDARY 4u

typedef struct RtbSceneInfo {
  int32_t image_width, image_height;
  int32_t spp_used; /* nearest square below the request (Q1) */
  int32_t sqrt_spp;
  int32_t max_depth;
  int32_t n_surface_prims, n_boundary_prims, n_media, n_bvh_nodes, n_lights;
  int32_t bvh_depth;
  int32_t device;
} RtbSceneInfo;

enum RtbPipeline {
  RTB_PIPELINE_DEFAULT = 0,
  RTB_PIPELINE_MEGAKERNEL = 1, /* one persistent kernel, per-lane path regeneration (A/B arm) */
  RTB_PIPELINE_WAVEFRONT = 2   /* ray-gen / extend / shade / accumulate kernels over SoA queues */
};

typedef struct RtbRenderParams {
  int64_t sample_begin; /* flat stratum index s = s_j*sqrt_spp + s_i, range [begin, end)     */
  int64_t sample_end;   /* (0, spp_used) renders the whole image; ranks split this range     */
  int32_t pipeline;     /* RtbPipeline */
  int32_t collect_stats; /* non-zero: fill the counters in RtbStats (slower)                 */
} RtbRenderParams;

typedef struct RtbStats {
  uint64_t paths;
  uint64_t segments;
  uint64_t node_visits;
  uint64_t prim_tests;
  uint64_t medium_probes;
  uint64_t nonfinite_samples;
  uint64_t kernel_launches;
  double device_ms; /* CUDA-event time of the kernels of this call */
  uint64_t exact_tests;   /* f64 reference-order primitive tests (wavefront: on candidates, in the shade stage) */
  uint64_t overflow_rays; /* rays re-traced by the exact kernel because their candidates did not fit the slots  */
  double stage_ms[3];     /* RTB_OPT_PROFILE only: CUDA-event totals of generate / extend / shade              */
} RtbStats;

/* per-scene tuning knobs (rtb_scene_set_option); every default is the measured best */
enum RtbOption {
  RTB_OPT_WF_CAPACITY = 1,      /* path slots of the wavefront queues (0 = sized by the call; >= 1024)           */
  RTB_OPT_EXACT_LEAVES = 2,     /* 1: f64 primitive tests inside the traversal (the round-1 extend kernel)       */
  RTB_OPT_SMEM_TOP = 3,         /* 1: top BVH levels staged in shared memory                                     */
  RTB_OPT_NO_DEFER_RARE = 4,    /* 1: textured Lambertian items shaded in place                                  */
  RTB_OPT_EXTEND_BLOCKS = 5,    /* cap of persistent extend blocks per SM (0 = what fits)                        */
  RTB_OPT_FINISH_BELOW = 6,     /* rays left at which the finishing kernel takes over (-1 default, 0 never)      */
  RTB_OPT_PROFILE = 7,          /* 1: per-stage CUDA-event totals on stderr; 2: per iteration                    */
  RTB_OPT_MEGA_BELOW = 8        /* RTB_PIPELINE_DEFAULT renders calls of fewer paths with the megakernel (-1 default) */
};

typedef struct rtb_scene rtb_scene;

/* ---- entry points ------------------------------------------------------------------------ */
int rtb_version(void);
int rtb_device_count(void);
const char* rtb_last_error(void);

/* validate + flatten + bake instance transforms + build the BVH + upload to `device`.
 * Replaces nothing in the reference (it holds its objects by value); this is the `flatten()` sink. */
int rtb_scene_create(const RtbSceneDesc* desc, int device, rtb_scene** out);
void rtb_scene_destroy(rtb_scene* scene);
int rtb_scene_info(const rtb_scene* scene, RtbSceneInfo* info);

/* The hot path. Replaces the body of render_par_lights (src/render.rs:171-197) for the stratum
 * range in `params`: pixels[3*(y*w+x)+c] += sum over samples of ray_color (f64 sums, caller-zeroed,
 * accumulated INTO like the reference's `row[i] = row[i] + color`, Q24). Blocking. */
int rtb_render(rtb_scene* scene, const RtbRenderParams* params, double* pixels_rgb, RtbStats* stats);

/* Same, but accumulates into a caller-owned DEVICE buffer of w*h x 4 uint64 {r, g, b, count} on
 * `cuda_stream` (a cudaStream_t, may be NULL) without synchronising.  r, g, b are two's-complement
 * fixed-point sums in units of 2^-32 (RTB_ACCUM_SCALE) added with integer atomics: the buffer does not
 * depend on the order paths finish in, so a render is bit-reproducible and buffers of disjoint stratum
 * ranges -- other calls, other GPUs -- add up exactly (reduce them as int64 sums).  count = strata
 * accumulated per pixel; bit 62 marks a pixel poisoned under RTB_FLAG_PROPAGATE_NAN.
 * Stats are valid after the stream has been synchronised and rtb_render_stats() has been called. */
#define RTB_ACCUM_SCALE 4294967296.0
int rtb_render_device(rtb_scene* scene, const RtbRenderParams* params, void* d_accum_u64x4,
                      void* cuda_stream);
int rtb_render_stats(rtb_scene* scene, RtbStats* stats);
/* device accumulation buffer -> host f64 sums, accumulated INTO pixels_rgb like rtb_render does (blocking) */
int rtb_accum_to_pixels(rtb_scene* scene, const void* d_accum_u64x4, double* pixels_rgb);

/* The reference seam on one multi-GPU box: render_par_lights (src/render.rs:144-216) for the stratum range in
 * `params`, split into contiguous slices over `n_devices` GPUs (devices[] or, if NULL, 0..n_devices-1): one
 * host thread, scene copy and stream per GPU, ONE NCCL sum-reduce of the int64 accumulation buffers to the
 * first device over NVLink, ONE device-to-host copy.  pixels_rgb as in rtb_render; the image does not depend
 * on n_devices (bit-identical).  stats (may be NULL): sums over devices, device_ms = the slowest device. */
int rtb_render_multi(const RtbSceneDesc* desc, int n_devices, const int* devices, const RtbRenderParams* params,
                     double* pixels_rgb, RtbStats* stats);

/* On-disk checkpoints of a progressive render: the accumulation buffer (fixed-point sums + strata counts) of `scene`'s
 * image size, written to / read from `path` (64-byte header: magic, abi, width, height, seed, flags; then w*h*4 u64).
 * Because the sums are exact integers, "render [0,k), save, ... load, render [k,n)" leaves the very bits of one
 * uninterrupted render of [0,n).  load OVERWRITES d_accum_u64x4 and fails if the file belongs to another image size or seed. */
int rtb_checkpoint_save(rtb_scene* scene, const void* d_accum_u64x4, const char* path);
int rtb_checkpoint_load(rtb_scene* scene, void* d_accum_u64x4, const char* path);

int rtb_scene_set_option(rtb_scene* scene, int option /* RtbOption */, int64_t value);
/* frees the idle blocks of the process-wide buffer cache (queues, staging); returns the bytes released */
int64_t rtb_trim_cache(void);

/* Deterministic closest-hit of `n` host rays against the scene's SURFACES (media are stochastic
 * and ignored here): the parity harness of SURVEY 8(d). Mirrors HittableList::hit
 * (src/hittable.rs:88-109) over Interval{t_min, INF}. */
int rtb_trace(rtb_scene* scene, const RtbRay* rays, int64_t n, uint32_t flags, RtbHit* hits);

/* Pixel-centre primary rays (no jitter, time 0, t_min 1e-4) in row-major order: the
 * deterministic part of get_ray (src/render.rs:221-232). rays must hold w*h entries. */
int rtb_camera_rays(const rtb_scene* scene, RtbRay* rays);

/* Boundary interval of medium `medium` for each ray: t_enter = first boundary hit over the whole
 * line, t_exit = first hit after t_enter + 1e-4 (src/constant_medium.rs:46-55); NaN = no hit. */
int rtb_medium_interval(rtb_scene* scene, int32_t medium, const RtbRay* rays, int64_t n,
                        double* t_enter, double* t_exit);

/* KAT hooks: Texture::value (src/texture.rs:18-25) and the light-list pdf
 * HittableList::pdf_value (src/hittable.rs:115-124), evaluated on the device. */
int rtb_eval_texture(rtb_scene* scene, int32_t texture, const double* uvp /* n x 5: u v px py pz */,
                     int64_t n, double* rgb_out /* n x 3 */);
int rtb_eval_light_pdf(rtb_scene* scene, const double* origin_dir /* n x 6 */, int64_t n,
                       double* pdf_out);

/* Output stage: write_color (src/color.rs:8-33) on the device:
 * divide by spp, optional exposure (exposure <= 0: none), sRGB OETF, clamp, (256*x) as u8.
 * `scene` may be NULL (the image of rtb_render_multi): the current CUDA device is used. */
int rtb_write_color(rtb_scene* scene, const double* pixels_rgb, int64_t n_pixels, double spp,
                    double exposure, uint8_t* rgb8_out);
/* auto_expose (src/render.rs:325-339): exposure value for write_color from the f64 sums, evaluated on the host
 * in the reference's own sequential order (a serial f64 sum: any other order moves the last bit). */
int rtb_auto_expose(const double* pixels_rgb, int64_t n_pixels, double spp, double* exposure_out);
/* KAT hook: Dielectric::scatter (src/material.rs:167-191: Schlick reflectance :156-163, reflect / refract src/vec3.rs:219-229)
 * on the device.  in9 = n x {direction[3], face normal[3], front_face (0/1), ir, uniform draw}; dir_out = n x 3. */
int rtb_eval_dielectric(rtb_scene* scene, const double* in9, int64_t n, double* dir_out);
/* Random123 known-answer hook for the DEVICE copy of Philox4x32-10: ctr_key = n x {c0,c1,c2,c3,k0,k1}, out = n x 4 */
int rtb_philox(rtb_scene* scene, const uint32_t* ctr_key, int64_t n, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */
