/*
 * rayito_b200 -- C ABI of the B200 render core (librayito_b200.so).
 *
 * This is the drop-in boundary for Rayito's render hot path.  The reference has no
 * FFI; its only seams are C++ calls, and each entry point below names the reference
 * call it stands behind (file:line relative to the reference repository, Stage 7
 * unless noted).  Every signature uses plain pointers and sizes.  All functions
 * return 0 on success or a negative RtStatus; rt_last_error_string() describes the
 * last failure on the calling thread.  There is no CPU fallback: without a CUDA
 * device every compute entry point fails with RT_ERR_CUDA.
 *
 * Host buffers passed in are borrowed for the duration of the call only.
 */
#ifndef RAYITO_B200_H
#define RAYITO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 2

typedef enum RtStatus
{
    RT_OK = 0,
    RT_ERR_ARG = -1,        /* bad argument / inconsistent scene description */
    RT_ERR_CUDA = -2,       /* CUDA runtime failure (including "no device") */
    RT_ERR_DEPTH = -3,      /* a BVH is deeper than 49: the reference's 50-entry
                               traversal stack (RAccel.h:379,414,502) would overflow */
    RT_ERR_UNSUPPORTED = -4,
    RT_ERR_COMM = -5        /* NCCL failure, or no NCCL library to load */
} RtStatus;

/* Mirrors Rayito::Ray (RRay.h:31-36; sizeof == 32). */
typedef struct RtRay
{
    float origin[3];
    float direction[3];
    float tmax;
    float time;
} RtRay;

/* Result of scene.intersect(): Intersection::m_t and the identity of the winner
 * (RRay.h:98-105).  shape is the index in "finite shapes in insertion order, then
 * infinite shapes" or -1; face / tri are the mesh face and fan triangle, else -1.
 * On a miss t is the ray's tmax. */
typedef struct RtHit
{
    float t;
    int32_t shape;
    int32_t face;
    int32_t tri;
} RtHit;

/* RtHit plus the shading inputs the reference stores in Intersection
 * (m_normal, m_colorModifier -- always a grey value: 1 or 0.2). */
typedef struct RtHitEx
{
    float t;
    int32_t shape;
    int32_t face;
    int32_t tri;
    float normal[3];
    float color_modifier;
} RtHitEx;

enum { RT_SHAPE_PLANE = 0, RT_SHAPE_SPHERE = 1, RT_SHAPE_RECT = 2, RT_SHAPE_MESH = 3 };
enum { RT_BRDF_NONE = 0 /* Emitter */, RT_BRDF_LAMBERT = 1, RT_BRDF_GLOSSY = 2, RT_BRDF_MIRROR = 3 };
enum { RT_NO_INDEX = 0xffffffffu };

/* One member of the ShapeSet (RScene.h:246-252).  A ShapeLight (RLight.h:250-332)
 * is flattened to the geometry and transform of the shape it wraps, with the
 * light's Emitter as material and light >= 0. */
typedef struct RtShape
{
    uint32_t type;       /* RT_SHAPE_* */
    uint32_t geom;       /* index into planes / spheres / rects / meshes */
    uint32_t xform;      /* index into xforms */
    uint32_t material;   /* index into materials */
    int32_t light;       /* index into lights, or -1 */
} RtShape;

/* Keyed scale-rotate-translate transform (RMath.h:619-941); num_keys == 0 is the
 * keyless identity.  Rotation keys must already be normalised (Transform::prepare). */
typedef struct RtXform
{
    uint32_t first_key;
    uint32_t num_keys;
} RtXform;

typedef struct RtPlane  { float position[3]; float normal[3]; uint32_t bullseye; } RtPlane;
typedef struct RtSphere { float position[3]; float radius; } RtSphere;
typedef struct RtRect   { float position[3]; float side1[3]; float side2[3]; } RtRect;

/* Polygon mesh (RMesh.h:39-60) with its face BVH in the reference node format. */
typedef struct RtMesh
{
    uint32_t first_vertex, num_vertices;     /* into vertices (xyz triples) */
    uint32_t first_normal, num_normals;      /* into normals */
    uint32_t first_face, num_faces;          /* into face_start / face_flags */
    uint32_t first_node, num_nodes;          /* into mesh_nodes */
    uint32_t first_cdf;                      /* into face_area_cdf (num_faces + 1 floats) */
    float total_area;
} RtMesh;

/* Rayito::BvhNode as laid out by the reference (RAccel.h:136-145, 32 bytes):
 * flags bits 0-1 = split axis, bit 2 = leaf (RAccel.h:119-124). */
typedef struct RtBvhNode
{
    float bbox_min[3];
    float bbox_max[3];
    uint32_t first_child_or_prim;
    uint32_t flags;
} RtBvhNode;

typedef struct RtMaterial
{
    float color[3];       /* reflectance (Diffuse/Glossy/Reflection) */
    float emittance[3];   /* Emitter: color * power (RMaterial.h:537) */
    float exponent;       /* Glossy: 1 / roughness^2 (RMaterial.h:211) */
    uint32_t brdf;        /* RT_BRDF_* */
} RtMaterial;

/* Flat scene: what ShapeSet::prepare() leaves behind (RScene.h:186-205). */
typedef struct RtSceneDesc
{
    uint32_t abi_version;            /* RT_ABI_VERSION */
    uint32_t set_xform;              /* transform of the ShapeSet itself */

    uint32_t num_finite;             /* ShapeSet::m_shapes, insertion order */
    uint32_t num_infinite;           /* ShapeSet::m_infiniteShapes (planes) */
    const RtShape* shapes;           /* num_finite + num_infinite entries */

    uint32_t num_top_nodes;          /* 0 => linear list (<= 2 finite shapes, RScene.h:135) */
    const RtBvhNode* top_nodes;

    uint32_t num_xforms;
    const RtXform* xforms;
    uint32_t num_keys;
    const float* key_time;           /* num_keys */
    const float* key_scale;          /* num_keys * 3 */
    const float* key_rotation;       /* num_keys * 4, (w, x, y, z) */
    const float* key_translation;    /* num_keys * 3 */

    uint32_t num_planes;  const RtPlane* planes;
    uint32_t num_spheres; const RtSphere* spheres;
    uint32_t num_rects;   const RtRect* rects;
    uint32_t num_meshes;  const RtMesh* meshes;

    uint32_t num_vertices;   const float* vertices;      /* xyz */
    uint32_t num_normals;    const float* normals;       /* xyz */
    uint32_t num_faces;
    const uint32_t* face_start;      /* num_faces + 1 offsets into the index arrays */
    const uint32_t* face_has_normals;/* num_faces flags */
    uint32_t num_indices;
    const uint32_t* vertex_index;    /* mesh-local vertex indices */
    const uint32_t* normal_index;    /* mesh-local normal indices (ignored if !has_normals) */
    uint32_t num_mesh_nodes; const RtBvhNode* mesh_nodes;
    uint32_t num_cdf;        const float* face_area_cdf;

    uint32_t num_materials;  const RtMaterial* materials;
    uint32_t num_lights;     const uint32_t* lights;     /* shape indices, findLights() order */

    /* Which stage's rules apply (RT_SEMANTICS_*).  Stage 6 (Rayito_Stage6_QT) has no
     * transforms at all (every xform must be keyless and is not applied, so a -0.0
     * stays -0.0), a face claims a hit at its FIRST fan triangle (S6 RMesh.h:204-209),
     * flat-shaded faces keep the un-normalised geometric normal (S6 RMesh.h:298), and
     * the renderer loops over every light and counts emission at bounce 0 only
     * (S6 RaytraceMain.cpp:250, 274-384). */
    uint32_t semantics;
} RtSceneDesc;

enum { RT_SEMANTICS_STAGE7 = 0, RT_SEMANTICS_STAGE6 = 6 };

/* PerspectiveCamera after its constructor ran (RaytraceMain.cpp:205-222). */
typedef struct RtCamera
{
    float origin[3];
    float forward[3];
    float right[3];
    float up[3];
    float tan_fov;
    float focal_distance;
    float lens_radius;
    float shutter_open;
    float shutter_close;
} RtCamera;

/* Arguments of raytrace() (rayito.h:138-144) plus the screen-tile partition used
 * to shard one image over ranks. */
typedef struct RtRenderParams
{
    uint32_t width, height;
    uint32_t pixel_samples_hint;     /* spp = hint^2 */
    uint32_t light_samples_hint;     /* light samples per bounce = hint^2 */
    uint32_t max_ray_depth;
    uint32_t tile_size;              /* square tiles; 0 = library default */
    uint32_t rank, world;            /* this call renders tiles with index % world == rank */
    uint32_t max_batch_samples;      /* wavefront size cap; 0 = library default */
    uint32_t flags;                  /* RT_RENDER_* */
} RtRenderParams;

enum
{
    RT_RENDER_COUNT_WORK = 1u,  /* also count node pops / triangle tests (slower) */
    RT_RENDER_TIME_TRACE = 2u,  /* CUDA-event pairs around every traversal stage -> trace_ms */
    RT_RENDER_UNIFIED_TRAVERSAL = 4u, /* one kernel walks both BVH levels (default: split top-level / mesh passes) */
    RT_RENDER_DYNAMIC_TOP = 8u        /* split mode: per-lane top-level pass instead of the tabulated walk (A/B testing) */
};

typedef struct RtRenderStats
{
    uint64_t samples;             /* pixel samples traced */
    uint64_t closest_rays;        /* scene.intersect calls (RaytraceMain.cpp:294,423) */
    uint64_t any_rays;            /* scene.doesIntersect calls (RaytraceMain.cpp:395) */
    uint64_t node_pops;           /* BVH nodes popped, both levels (RT_RENDER_COUNT_WORK) */
    uint64_t tri_tests;           /* triangles tested */
    uint64_t shape_tests;         /* analytic shapes tested */
    uint64_t xform_evals;         /* Ray::transformToLocal calls (RRay.h:78-81): the set and every shape entered */
    uint64_t xform_keyed;         /* ... of which on a transform with >= 1 key (the reference reads key data, RMath.h:681-715) */
    uint64_t xform_pairs;         /* ... of which on a transform with >= 2 keys (a key pair may be read) */
    uint64_t kernel_launches;     /* CUDA kernels launched by this call */
    uint64_t trace_launches;      /* of which traversal kernels (closest / any hit) */
    float render_ms;              /* device time of the render (CUDA events) */
    float trace_ms;               /* device time inside the traversal kernels (RT_RENDER_TIME_TRACE) */
    float upload_ms;              /* host->device scene/camera copies inside this call */
    float download_ms;            /* device->host image copy inside this call */
} RtRenderStats;

typedef struct RtScene RtScene;

const char* rt_last_error_string(void);
int rt_abi_version(void);
/* Number of visible CUDA devices (0 when there is none; never fails). */
int rt_device_count(void);

/* Upload a prepared scene to `device`.  Stands behind scene.prepare() at
 * RaytraceMain.cpp:497.  Fails with RT_ERR_DEPTH if any BVH is deeper than 49. */
int rt_scene_create(const RtSceneDesc* desc, int device, RtScene** out_scene);
/* The same with options.  RT_SCENE_BUILD_MESH_BVH: Bvh<Mesh>::build (RAccel.h:262-374, called from
 * Mesh::prepare, RMesh.h:128) runs ON THE DEVICE, out of the uploaded faces, for every mesh that comes without
 * nodes (RtMesh.num_nodes == 0: a mesh of F faces then gets 2F-1 nodes; meshes that bring their nodes keep
 * them), and no node of those trees crosses PCIe.  The tree is the reference's, node for node -- element order of std::partition,
 * slot numbering of the recursion, boxes down to the sign of a zero (rayito_b200/csrc/rt_build.cuh) --
 * so hit records stay bit-equal.  Stage 7 semantics only. */
enum { RT_SCENE_BUILD_MESH_BVH = 1u };
int rt_scene_create_ex(const RtSceneDesc* desc, int device, uint32_t flags, RtScene** out_scene);
int rt_scene_destroy(RtScene* scene);
/* Read a mesh's face BVH back in the reference's node format (RtBvhNode: leaves name their face), whoever built
 * it: `nodes` receives min(capacity, 2F-1) nodes.  depth (may be NULL): deepest leaf, root = 0; build_ms (may
 * be NULL): device time of the device build of all meshes, 0 for a host-built scene.  For tests and tools. */
int rt_scene_mesh_nodes(RtScene* scene, uint32_t mesh, RtBvhNode* nodes, uint32_t capacity, uint32_t* depth, float* build_ms);

/* ShapeSet::intersect (RScene.h:120-156) for n rays; host buffers. */
int rt_trace_closest(RtScene* scene, const RtRay* rays, size_t n, RtHit* hits);
int rt_trace_closest_ex(RtScene* scene, const RtRay* rays, size_t n, RtHitEx* hits);
/* ShapeSet::doesIntersect (RScene.h:158-184) for n rays; hits[i] is 0 or 1. */
int rt_trace_any(RtScene* scene, const RtRay* rays, size_t n, uint8_t* hits);

/* Same, on buffers already resident on the scene's device, enqueued on `stream`
 * (a cudaStream_t passed as void*; NULL = the legacy default stream).  These do
 * not synchronise; up to 16 calls per scene may be in flight on different streams at once
 * (each takes its own work cursor), and rt_scene_destroy() waits for the device first.  work, if not NULL, is a device array of 6 uint64 counters
 * {node_pops, tri_tests, shape_tests, xform_evals, xform_keyed, xform_pairs} (meanings as in
 * RtRenderStats) that the kernel adds to. */
int rt_trace_closest_device(RtScene* scene, const RtRay* d_rays, size_t n, RtHit* d_hits,
                            uint64_t* d_work, void* stream);
int rt_trace_any_device(RtScene* scene, const RtRay* d_rays, size_t n, uint8_t* d_hits,
                        uint64_t* d_work, void* stream);

/* rt_trace_closest / rt_trace_any that also return the work of the batch: work[0..5] =
 * {node_pops, tri_tests, shape_tests, xform_evals, xform_keyed, xform_pairs}, the numbers
 * the roofline's algorithmic bytes are made of.  Host buffers. */
int rt_trace_closest_counted(RtScene* scene, const RtRay* rays, size_t n, RtHit* hits, uint64_t* work);
int rt_trace_any_counted(RtScene* scene, const RtRay* rays, size_t n, uint8_t* hits, uint64_t* work);

/* raytrace() (RaytraceMain.cpp:485-579): render the tiles of this rank into rgb
 * (width*height*3 floats, row-major, rows top-down like Image::pixel).  Pixels
 * of other ranks' tiles are left untouched.  Host output buffer. */
int rt_render(RtScene* scene, const RtCamera* camera, const RtRenderParams* params,
              float* rgb, RtRenderStats* stats);
/* Same with a device output buffer; enqueues on `stream` and synchronises it
 * before returning so that stats are final. */
int rt_render_device(RtScene* scene, const RtCamera* camera, const RtRenderParams* params,
                     float* d_rgb, RtRenderStats* stats, void* stream);

/* ---- one frame over several GPUs (one process per GPU) ------------------------------
 * The reference shards a frame over 16 threads by image chunks that all write one Image
 * (RaytraceMain.cpp:504-568).  Here ranks own screen tiles (rt_tile_owners): each renders
 * its tiles into a PACKED buffer -- its tiles in ascending tile index one after the other,
 * pixel (row r, column q) of the k-th tile at ((k * tile + r) * tile + q) * 3 -- and the
 * packed buffers are gathered on one rank and scattered into the frame.  That gather is the
 * only collective of the path (NCCL send / receive over NVLink; the library loads NCCL at
 * run time, libnccl.so.2 or RAYITO_B200_NCCL_LIB, and reuses a copy the process already
 * holds).  One RtComm per process and device. */
#define RT_COMM_ID_BYTES 128
typedef struct RtComm RtComm;
/* ncclGetUniqueId: call on one rank, hand the 128 bytes to all the others (MPI, a file, ...). */
int rt_comm_unique_id(uint8_t* id);
/* ncclCommInitRank on `device`: collective over all `world` ranks.  world == 1 needs no NCCL. */
int rt_comm_create(const uint8_t* id, int rank, int world, int device, RtComm** out_comm);
/* Wrap an ncclComm_t the application already has (borrowed: not destroyed with the handle). */
int rt_comm_from_nccl(void* nccl_comm, int device, RtComm** out_comm);
int rt_comm_destroy(RtComm* comm);
/* raytrace() over all ranks of `comm`: every rank renders the tiles params->rank / world name
 * (which must equal the communicator's), the root receives everybody's packed tiles and
 * assembles the frame in d_rgb (device, width*height*3 floats; ignored on the other ranks).
 * stats = this rank's render; assemble_ms (may be NULL) = device time of gather + scatter.
 * Enqueues on `stream` and synchronises it before returning. */
int rt_render_multi(RtScene* scene, const RtCamera* camera, const RtRenderParams* params, RtComm* comm, int root,
                    float* d_rgb, RtRenderStats* stats, float* assemble_ms, void* stream);
/* The same with the assembled frame copied to HOST memory on the root (rgb: width*height*3
 * floats there, ignored elsewhere): what Rayito::raytrace() hands back.  Legacy stream. */
int rt_render_multi_host(RtScene* scene, const RtCamera* camera, const RtRenderParams* params, RtComm* comm, int root,
                         float* rgb, RtRenderStats* stats, float* assemble_ms);
/* Rank, size and device of a communicator (any pointer may be NULL). */
int rt_comm_rank(const RtComm* comm, int* rank, int* world, int* device);
/* The two halves, for applications that move the tiles themselves: floats in the packed buffer
 * of a rank (0 if it owns no tile); render this rank's tiles packed into a device buffer;
 * scatter a rank's packed tiles into a device frame. */
size_t rt_packed_floats(uint32_t width, uint32_t height, uint32_t tile_size, uint32_t world, uint32_t rank);
int rt_render_tiles_packed(RtScene* scene, const RtCamera* camera, const RtRenderParams* params, float* d_packed,
                           size_t packed_floats, RtRenderStats* stats, void* stream);
int rt_unpack_tiles(int device, const float* d_packed, uint32_t width, uint32_t height, uint32_t tile_size,
                    uint32_t world, uint32_t rank, float* d_rgb, void* stream);

/* Generate only the camera rays of pixel-sample `psi` for every pixel of this
 * rank's tiles (RenderThread::run, RaytraceMain.cpp:112-142): rays[y*width+x]. */
int rt_generate_camera_rays(RtScene* scene, const RtCamera* camera, const RtRenderParams* params,
                            uint32_t psi, RtRay* rays);

/* displayImage() (MainWindow.cpp:37-91): exposure, gamma, clamp, truncate to
 * 8 bits.  out is width*height*4 bytes B,G,R,A.  Host buffers. */
int rt_tonemap_bgra8(int device, const float* rgb, size_t num_pixels, float exposure_stops, float gamma,
                     uint8_t* bgra);
/* The same on DEVICE buffers -- the frame rt_render_device / rt_render_multi left in HBM goes to
 * display bytes without a trip through host memory.  Enqueued on `stream`, not synchronised. */
int rt_tonemap_bgra8_device(int device, const float* d_rgb, size_t num_pixels, float exposure_stops, float gamma,
                            uint8_t* d_bgra, void* stream);

/* ---- Stage 1 (BASELINE.json configs[0]) ------------------------------------------ */

/* Stage 1 plane (Rayito_Stage1/rayito.h:465-512): one-sided, flat colour, normal
 * already normalised by the constructor. */
typedef struct RtStage1Plane
{
    float position[3];
    float normal[3];
    float color[3];
} RtStage1Plane;

/* The whole Stage 1 program (Rayito_Stage1/main.cpp:65-135): one ray per pixel
 * through the pixel CORNERS (x/(W-1), 1 - y/(H-1)), closest plane in list order with
 * kRayTMin = 1e-5 (rayito.h:301), colour clamped and truncated to 8 bits.  `camera`
 * holds makeCameraRay's basis (main.cpp:28-52: forward, right and up all normalised,
 * tan of the full field of view).  rgb8 = width*height*3 bytes, the P6 payload. */
int rt_stage1_render(int device, const RtStage1Plane* planes, uint32_t num_planes, const RtCamera* camera,
                     uint32_t width, uint32_t height, uint8_t* rgb8);

/* The same, also returning pixelColor before clamp() (rgb: width*height*3 floats), which is what
 * the program streams out when built with WRITE_PFM (main.cpp:57,122-123). */
int rt_stage1_render_float(int device, const RtStage1Plane* planes, uint32_t num_planes, const RtCamera* camera,
                           uint32_t width, uint32_t height, float* rgb, uint8_t* rgb8);

/* Device memory the library keeps between calls for speed (one wavefront state block per device,
 * parked when a scene is destroyed; the Stage 2/3 working set) is freed here.  Optional. */
void rt_release_cached_memory(void);

/* ---- Stage 2 / Stage 3: the serial-Rng programs ------------------------------- */
/* Rayito_Stage2/main.cpp and Rayito_Stage3/main.cpp (BASELINE config C2 is the Stage 3
 * pixel-sample sweep).  A handful of analytic shapes in a linear list, closest hit also
 * for shadow rays, kRayTMin = 1e-5, ONE Rng for the whole image in scan order.  The
 * library finds every sample's position in that stream on the device (see
 * csrc/rt_stage23.cuh) and reproduces the programs' images bit for bit. */
enum { RT_S23_PLANE = 0, RT_S23_SPHERE = 1, RT_S23_RECT = 2 };
enum { RT_S23_MAT_LAMBERT = 0, RT_S23_MAT_PHONG = 1, RT_S23_MAT_EMITTER = 2 };
enum { RT_S23_MAX_SHAPES = 16, RT_S23_MAX_LIGHTS = 4 };

typedef struct RtS23Shape
{
    uint32_t type;          /* RT_S23_* */
    uint32_t material;      /* index into RtS23Scene.materials; a light's is its Emitter */
    float position[3];      /* plane point / sphere centre / rectangle corner */
    float normal[3];        /* plane: normalised by the constructor (Rayito_Stage3/rayito.h:735) */
    float side1[3], side2[3];   /* rectangle light */
    float radius;
    uint32_t bullseye;      /* plane: colour modifier 0.2 on alternate rings (rayito.h:773) */
} RtS23Shape;

typedef struct RtS23Material
{
    uint32_t kind;          /* RT_S23_MAT_* (Stage 2 surfaces: LAMBERT with the shape's colour) */
    float color[3];
    float exponent;         /* Phong */
    float emittance[3];     /* Emitter: colour * power (rayito.h:490); zero otherwise */
} RtS23Material;

typedef struct RtS23Scene
{
    uint32_t num_shapes;    const RtS23Shape* shapes;        /* ShapeSet list order */
    uint32_t num_materials; const RtS23Material* materials;
    uint32_t num_lights;    const uint32_t* lights;          /* shape indices, findLights() order */
} RtS23Scene;

typedef struct RtS23Params
{
    uint32_t stage;                             /* 2 or 3 */
    uint32_t width, height;
    uint32_t pixel_samples_u, pixel_samples_v;  /* Stage 3: strata (reference 4 x 4); Stage 2: u = random samples per pixel (64), v ignored */
    uint32_t light_samples_u, light_samples_v;  /* Stage 3: strata per light (reference 4 x 4); Stage 2: ignored (one sample) */
    uint32_t seed_z, seed_w;                    /* Rng() defaults: 362436069, 521288629 (main.cpp:35) */
} RtS23Params;

/* Renders the whole program.  `camera` holds makeCameraRay's basis (main.cpp:55-79:
 * forward, right, up normalised; tan of the FULL field of view).  rgb (may be NULL):
 * width*height*3 floats, pixelColor after the box-filter division and before clamp();
 * rgb8 (may be NULL): the P6 payload the program streams into out.ppm.  stats (may be
 * NULL): samples, closest_rays (= every ShapeSet::intersect call), render_ms = stream
 * pre-pass (upload_ms) + shading (trace_ms), trace_launches = pre-pass rounds. */
int rt_stage23_render(int device, const RtS23Scene* scene, const RtCamera* camera, const RtS23Params* params,
                      float* rgb, uint8_t* rgb8, RtRenderStats* stats);

/* ---- host-side pieces of the path (no device needed) ------------------------- */

/* Tile partition used to shard one image over `world` ranks: owners[ty*tiles_x+tx] is
 * the rank that renders that tile (tile_size 0 = library default).  Either count
 * pointer may be NULL; owners may be NULL to query the counts only. */
int rt_tile_owners(uint32_t width, uint32_t height, uint32_t tile_size, uint32_t world,
                   uint32_t* owners, uint32_t* tiles_x, uint32_t* tiles_y, uint32_t* tile_size_used);

/* The 5*depth+3 sampler permutations the reference draws for pixel (x, y)
 * (RaytraceMain.cpp:69-108, 159-169), computed by MWC jump-ahead exactly as the
 * device does.  out[5b+0..4] = bounce, light selection, light element, light, brdf of
 * bounce b; out[5*depth+0..2] = time, lens, subpixel. */
int rt_sample_permutations(uint32_t width, uint32_t height, uint32_t depth, uint32_t x, uint32_t y, uint32_t* out);

/* CorrelatedMultiJitterSampler::sample1D / sample2D (RSampling.h:272-306), the same
 * code the device compiles. */
float rt_cmj_sample1d(uint32_t index, uint32_t samples, uint32_t permutation);
void rt_cmj_sample2d(uint32_t index, uint32_t x_samples, uint32_t y_samples, uint32_t permutation, float* u, float* v);

/* The device's sinf / cosf / powf (rt_libm.cuh: the C library's algorithms restated,
 * see there), evaluated on the host from the same source, for n arguments.
 * kind 0 = sinf(x), 1 = cosf(x), 2 = powf(x, y), 3 / 4 = the sine / cosine of the
 * shared-reduction sincos.  y may be NULL unless kind is 2. */
int rt_libm_eval(int kind, const float* x, const float* y, size_t n, float* out);

#ifdef __cplusplus
}
#endif

#endif /* RAYITO_B200_H */
