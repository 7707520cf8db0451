/* imba_b200.h -- C ABI of the B200-native hot path of imbatracer.
 *
 * Plain C, plain pointers and sizes, no torch / CUDA types in any signature (streams are
 * passed as void*).  All functions are exported from libimba_b200.so.
 *
 * Citations are relative to the reference tree (microcompunics/imbatracer).
 *
 * Level 1 (L1) is a symbol-for-symbol replacement of the four entry points the reference
 * binds from the external AnyDSL/traversal library (core/traversal_interface.h:4-10, call
 * sites render/scheduling/ray_queue.h:254-265, 279-290, 304-315, 329-340).
 * Level 2 (L2) is the device-resident API behind the reference's RayQueue / RayScheduler /
 * Integrator plug-in points (ray_scheduler.h:15-38, integrator.h:19-39): the scene is
 * flattened once, queues live in HBM, a whole progressive iteration runs without leaving
 * the device.
 *
 * Error model: the reference's entry points return void and have no error channel
 * (ray_queue.h:248,270-271 only assert).  L1 keeps that signature; failures (no CUDA device,
 * launch error) are fatal: message on stderr + abort().  There is NO CPU fallback.
 * L2 functions return 0 on success or a negative imba_status; imba_last_error() gives text.
 */
#ifndef IMBA_B200_H
#define IMBA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * ABI data types (what the AnyDSL-generated traversal_{cpu,gpu}.h declare)
 * ---------------------------------------------------------------------------------------- */
typedef struct { float x, y, z, w; } imba_vec4;
typedef struct { float x, y; } imba_vec2;
/* org.w = tmin, dir.w = tmax (render/ray_gen/camera.h:61-64) */
typedef struct { imba_vec4 org; imba_vec4 dir; } imba_ray;
/* tri_id < 0 = miss (integrator.h:57); tmax = t along the un-normalised dir; u = weight of v1 (integrator.h:90-101) */
typedef struct { int32_t tri_id; int32_t inst_id; float tmax; float u; } imba_hit;
/* MBVH4 node of the CPU layout (core/cpu_adapter.cpp:14-25, 80-100) */
typedef struct { float min_x[4], min_y[4], min_z[4], max_x[4], max_y[4], max_z[4]; int32_t children[4]; } imba_node4;
/* BVH2 node of the GPU layout (core/gpu_adapter.cpp:14-31, 79-98) */
typedef struct { float lo_x, lo_y, lo_z, hi_x, hi_y, hi_z; } imba_bbox;
typedef struct { imba_bbox left_bb, right_bb; int32_t left, right; } imba_node2;
/* core/cpu_adapter.cpp:356-365 */
typedef struct { float transf[3][4]; int32_t id; int32_t next; int32_t pad[2]; } imba_instance_node;
/* core/mask.h:9-20 */
typedef struct { int32_t width, height, offset; } imba_mask;

typedef enum {
    IMBA_OK = 0,
    IMBA_ERR_NO_DEVICE = -1,
    IMBA_ERR_CUDA = -2,
    IMBA_ERR_INVALID = -3,
    IMBA_ERR_OVERFLOW = -4
} imba_status;

const char* imba_last_error(void);
/* Returns the number of CUDA kernels this library has launched since load (bench.py "gpu_launches"). */
uint64_t imba_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * L1: drop-in traversal entry points
 * ---------------------------------------------------------------------------------------- */
/* Replaces traversal_gpu::intersect_gpu_masked_instanced / occluded_gpu_masked_instanced
 * (ray_queue.h:279-290, 329-340).  Every pointer is a DEVICE pointer on the current CUDA
 * device (the reference allocates them with anydsl::Array(Platform::Cuda, Device(0)),
 * ray_queue.h:78-79, scene.cpp:14-26).  nodes/tris are in the reference's BVH2 GPU layout.
 * Synchronous: hits are complete on return (the caller D2H-copies them right away,
 * ray_queue.h:292).  Re-entrant for disjoint ray/hit buffers (tile_scheduler.h:69-78).
 * count may include padding rays (ray_queue.h:273); they are traced like any other ray.
 * The compressed 8-wide BVH is built on first use per (nodes, tris, instances, root) and cached. */
#ifndef IMBA_B200_NO_L1_DECLS
void intersect_gpu_masked_instanced(int root, imba_node2* nodes, imba_instance_node* instances, imba_vec4* tris,
                                    imba_ray* rays, imba_hit* hits, int* indices, imba_vec2* texcoords,
                                    imba_mask* masks, char* mask_buffer, int count);
void occluded_gpu_masked_instanced(int root, imba_node2* nodes, imba_instance_node* instances, imba_vec4* tris,
                                   imba_ray* rays, imba_hit* hits, int* indices, imba_vec2* texcoords,
                                   imba_mask* masks, char* mask_buffer, int count);
/* Replaces traversal_cpu::intersect_cpu_masked_instanced / occluded_cpu_masked_instanced
 * (ray_queue.h:254-265, 304-315).  Every pointer is a HOST pointer; nodes/tris are in the
 * reference's MBVH4 CPU layout.  Rays are staged through pinned memory in chunks and copied,
 * traced and copied back on two streams so that PCIe and the SMs overlap.  Lets an unmodified
 * `--cpu` build of the reference run its traversal on the B200 by swapping the library. */
void intersect_cpu_masked_instanced(int root, imba_node4* nodes, imba_instance_node* instances, imba_vec4* tris,
                                    imba_ray* rays, imba_hit* hits, int* indices, imba_vec2* texcoords,
                                    imba_mask* masks, char* mask_buffer, int count);
void occluded_cpu_masked_instanced(int root, imba_node4* nodes, imba_instance_node* instances, imba_vec4* tris,
                                   imba_ray* rays, imba_hit* hits, int* indices, imba_vec2* texcoords,
                                   imba_mask* masks, char* mask_buffer, int count);
#endif /* IMBA_B200_NO_L1_DECLS */
/* Drops every cached acceleration structure built by the L1 entry points (call after the host
 * rebuilt or freed its node arrays, e.g. Scene::build_top_level_accel, scene.cpp:114). */
void imba_l1_invalidate(void);

/* ------------------------------------------------------------------------------------------
 * L2: flattened scene (what render/scene.h keeps host-side for shading)
 * ---------------------------------------------------------------------------------------- */
typedef enum { IMBA_ACCEL_MBVH4 = 0, IMBA_ACCEL_BVH2 = 1,
               /* no reference-built structure is supplied (nodes / tris / inst_nodes are ignored): imba_scene_create builds the
                * per-mesh trees on the device (linear BVH + SAH-optimal 8-wide collapse) and the small top level on the host;
                * replaces SplitBvhBuilder (core/sbvh_builder.h) + the adapters for scenes loaded without one */
               IMBA_ACCEL_DEVICE_BUILD = 2 } imba_accel_layout;
/* IMBA_MAT_PHONG: a BSDF holding the reference's Phong BxDF (render/materials/brdfs.h:48-92; coefficient = ks, exponent).  No material of
 * the reference's loader instantiates it (build_scene.cpp:125-161), a host program that does builds BSDF(isect, Phong, nullptr).
 * IMBA_MAT_OREN_NAYAR: likewise for OrenNayar (brdfs.h:94-134; reflectance = kd, roughness in degrees = exponent field). */
typedef enum { IMBA_MAT_DIFFUSE = 0, IMBA_MAT_MIRROR = 1, IMBA_MAT_GLASS = 2, IMBA_MAT_GLOSSY = 3, IMBA_MAT_PHONG = 4, IMBA_MAT_OREN_NAYAR = 5 } imba_material_kind;
typedef enum { IMBA_LIGHT_TRIANGLE = 0, IMBA_LIGHT_DIRECTIONAL = 1, IMBA_LIGHT_POINT = 2, IMBA_LIGHT_SPOT = 3,
               IMBA_LIGHT_ENV = 4 } imba_light_kind;

/* render/materials/materials.h:58-191 flattened to one tagged record (64 bytes). */
typedef struct {
    int32_t kind;          /* imba_material_kind */
    int32_t diffuse_tex;   /* texture index or -1 (DiffuseMaterial::sampler_, GlossyMaterial::diff_sampler_) */
    int32_t bump_tex;      /* texture index or -1 (Material::bump_) */
    int32_t is_emitter;    /* Material::emit_ != nullptr */
    float kd[3];           /* DiffuseMaterial::color_ / GlossyMaterial::diffuse_color_ */
    float eta;             /* MirrorMaterial fresnel eta / GlassMaterial::eta_ */
    float ks[3];           /* MirrorMaterial::scale_ / GlassMaterial::reflectance_ / GlossyMaterial::specular_color_ */
    float kappa;           /* MirrorMaterial fresnel kappa */
    float tf[3];           /* GlassMaterial::transmittance_ */
    float exponent;        /* GlossyMaterial::exponent_ */
    float emit[3];         /* AreaEmitter::intensity (light.h:14-47) */
    float emit_area;       /* AreaEmitter::area */
} imba_material;

/* render/light.h flattened (80 bytes). */
typedef struct {
    int32_t kind;          /* imba_light_kind */
    float   angle;         /* SpotLight::angle_ (radians) */
    float   pad0[2];
    float   intensity[4];  /* rgb (+pad) */
    float   p0[4];         /* triangle v0 / position / direction(dir light) */
    float   p1[4];         /* triangle v1 / spot direction */
    float   p2[4];         /* triangle v2 */
} imba_light;

typedef struct { int32_t mesh_id; float mat[12]; float inv_mat[12]; } imba_instance; /* core/mesh.h:20-27 (3x4 row-major) */
typedef struct { int32_t width, height; int64_t offset; } imba_texture; /* offset in float4 texels into the texel pool */

typedef struct {
    /* geometry of all meshes, concatenated in mesh order (render/scene.cpp:74-97) */
    int32_t        mesh_count;
    const int32_t* mesh_vertex_base;  /* [mesh_count + 1] */
    const int32_t* mesh_tri_base;     /* [mesh_count + 1]  == Scene::tri_layout_ */
    const float*   vertices;          /* 4 floats per vertex (core/mesh.h:66) */
    const float*   normals;           /* 4 floats per vertex (FLOAT3 attribute stride 16, mesh.h:128) */
    const float*   texcoords;         /* 2 floats per vertex */
    const float*   geom_normals;      /* 4 floats per triangle */
    const int32_t* indices;           /* 4 ints per triangle: mesh-local i0,i1,i2, material id */
    int32_t        instance_count;
    const imba_instance* instances;
    int32_t        material_count;
    const imba_material* materials;
    int32_t        light_count;
    const imba_light* lights;
    int32_t        texture_count;
    const imba_texture* textures;
    const float*   texels;            /* float4 RGBA pool */
    int64_t        texel_count;
    /* environment map (render/light.h:361-485); env_tex = -1 if none */
    int32_t        env_tex;
    float          env_intensity;
    /* scene bounding sphere (render/scene.cpp:190-202) */
    float          bsphere_center[3];
    float          bsphere_radius;
    /* acceleration structure in one of the reference's layouts (core/{cpu,gpu}_adapter.cpp) */
    int32_t        accel_layout;      /* imba_accel_layout */
    int32_t        root;
    const void*    nodes;             /* imba_node4[] or imba_node2[] */
    int32_t        node_count;
    const imba_vec4* tris;
    int64_t        tri_vec4_count;
    const imba_instance_node* inst_nodes;
    int32_t        inst_node_count;
    /* opacity masks (core/mask.h), optional: mask_count == 0 -> everything opaque */
    int32_t        mask_count;
    const imba_mask* masks;
    const uint8_t* mask_buffer;
    int64_t        mask_buffer_size;
    /* global per-triangle index records used by the mask test: 4 ints per global triangle with
     * texcoord indices offset per mesh (scene.cpp:89-93); may be NULL when mask_count == 0 */
    const int32_t* mask_indices;
    const imba_vec2* mask_texcoords;
    int64_t        mask_texcoord_count;
} imba_scene_desc;

typedef struct imba_scene imba_scene;

/* Uploads the scene to the current CUDA device and converts the acceleration structure to the
 * compressed 8-wide layout (replaces Scene::upload_* , scene.cpp:132-188). */
imba_scene* imba_scene_create(const imba_scene_desc* desc);
void        imba_scene_destroy(imba_scene* scene);
/* Statistics of the converted structure: out[0]=bvh8 nodes, [1]=triangles (48-byte blocks), [2]=instances,
 * [3]=bytes of node+triangle storage, [4]=max depth. */
int imba_scene_stats(const imba_scene* scene, int64_t out[5]);
/* IMBA_ACCEL_DEVICE_BUILD scenes: out[0] = device seconds of the mesh builds (sort, hierarchy, refit, collapse, emission),
 * out[1] = levels of the deepest 8-wide mesh tree; zeros for converted scenes. */
int imba_scene_build_info(const imba_scene* scene, float out[2]);

/* Closest-hit (any_hit = 0) or occlusion (any_hit = 1) traversal of `count` rays.
 * d_rays / d_hits are device pointers; asynchronous on `stream` (a cudaStream_t, NULL = default stream).
 * d_counters (optional, device, 4 uint32 per ray) receives per-ray {nodes fetched, triangle blocks fetched,
 * instances entered, stack high-water}: the N_node / N_tri of the roofline model. */
int imba_trace(imba_scene* scene, const imba_ray* d_rays, imba_hit* d_hits, int count, int any_hit,
               uint32_t* d_counters, void* stream);
/* Same with HOST buffers: pinned staging, chunked H2D / trace / D2H overlap; synchronous.
 * This is RayQueue::traverse_gpu / traverse_occluded_gpu (ray_queue.h:269-293, 319-343). */
int imba_trace_host(imba_scene* scene, const imba_ray* h_rays, imba_hit* h_hits, int count, int any_hit);

/* Page-locks / unlocks a host buffer that is passed to imba_trace_host repeatedly (a RayQueue's ray and hit arrays,
 * ray_queue.h:78-79) so that the chunked copies overlap with traversal. */
int imba_host_register(void* ptr, size_t bytes);
int imba_host_unregister(void* ptr);

/* Host-only conversion of a reference-layout structure to the device layout (no CUDA device needed): the arrays that
 * imba_scene_create uploads.  Node = 80 bytes, triangle = 48 bytes, instance = 64 bytes (imbatracer_b200/csrc/bvh8.h). */
typedef struct imba_bvh8 imba_bvh8;
imba_bvh8* imba_bvh8_build(int accel_layout, int root, const void* nodes, const imba_vec4* tris,
                           const imba_instance_node* inst_nodes);
void imba_bvh8_free(imba_bvh8* h);
int  imba_bvh8_arrays(const imba_bvh8* h, const void** nodes, int64_t* node_count, const void** tris, int64_t* tri_count,
                      const void** insts, int64_t* inst_count, uint32_t* top_root, int32_t* max_depth);

/* Host builder of the two-level structure in the reference's BVH2 layout for scenes that come without one (procedural
 * scenes): binned SAH per mesh and over the instance boxes, emitted exactly as core/gpu_adapter.cpp / render/scene.cpp:
 * 168-176 lay it out, so it feeds imba_scene_create (and the CPU oracle) like a reference-built structure.  Uses
 * mesh_*, vertices, indices, instances of `desc`; imba_accel_fill points desc's accel fields at the result. */
typedef struct imba_accel imba_accel;
imba_accel* imba_accel_build(const imba_scene_desc* desc);
/* Same result type with the per-mesh trees built on the current CUDA device (linear BVH: Morton codes, radix sort, parallel
 * hierarchy, bottom-up refit; csrc/lbvh.cu); out_seconds[0] (optional) receives the device time of the mesh builds. */
imba_accel* imba_accel_build_gpu(const imba_scene_desc* desc, float* out_seconds);
int  imba_accel_fill(const imba_accel* a, imba_scene_desc* desc);
int  imba_accel_arrays(const imba_accel* a, const void** nodes, int64_t* node_count, const void** tris, int64_t* tri_vec4_count,
                       const void** inst_nodes, int64_t* inst_node_count, int32_t* root);
void imba_accel_free(imba_accel* a);

/* The reference's per-mesh `.bvh` cache files (loaders/load_bvh.cpp:12-43, 144-302): magic 0x313F1A57, blocks of
 * {u64 size, u32 type, u32 node_count, u32 prim_count, nodes, prims}; type 1 = BVH2 (imba_node2 + Vec4 triples), type 2 =
 * MBVH4 (imba_node4 + Vec4 packets), indices file-local; type 3 (new) = this library's Node8 (80 B) + Tri48 (48 B).
 * prim_count counts Vec4 for types 1/2 and triangles for type 3.  Unknown block types are skipped on read. */
typedef struct imba_bvh_file imba_bvh_file;
imba_bvh_file* imba_bvh_cache_open(const char* path);
int  imba_bvh_cache_block(const imba_bvh_file* f, int type, const void** nodes, uint32_t* node_count, const void** prims, uint32_t* prim_count);
int  imba_bvh_cache_append(const char* path, int type, const void* nodes, uint32_t node_count, const void* prims, uint32_t prim_count);
void imba_bvh_cache_close(imba_bvh_file* f);

/* ------------------------------------------------------------------------------------------
 * L2: ray-queue primitives on device arrays (render/scheduling/ray_queue.h:161-244)
 * ---------------------------------------------------------------------------------------- */
/* compact_hits (ray_queue.h:161-185): stable partition, hits first; rays/hits/states are permuted
 * into the *_out arrays (ping-pong, no in-place swap).  Returns the hit count in *d_hit_count (device int). */
int imba_queue_compact_hits(imba_scene* scene, const imba_ray* d_rays, const imba_hit* d_hits, const void* d_states,
                            int state_bytes, imba_ray* d_rays_out, imba_hit* d_hits_out, void* d_states_out, int count,
                            int* d_hit_count, void* stream);
/* sort_by_material (ray_queue.h:211-244): fills d_sorted_indices[0..hit_count) with ray indices grouped by
 * material id, stable within a material (the reference's scatter order is racy, any grouping is valid); misses of the
 * `count` input hits are skipped (compact_hits is implied) and *d_hit_count (device int) receives the number of hits.
 * All three queue calls are synchronous; state_bytes must be a multiple of 8 (16 / 40 / 48 in the reference). */
int imba_queue_sort_by_material(imba_scene* scene, const imba_hit* d_hits, int count, int* d_sorted_indices,
                                int* d_hit_count, void* stream);
/* compact_rays (ray_queue.h:188-207): stable removal of states whose pixel_id (first int) is < 0. */
int imba_queue_compact_rays(const imba_ray* d_rays, const void* d_states, int state_bytes, imba_ray* d_rays_out,
                            void* d_states_out, int count, int* d_out_count, void* stream);
/* The same three primitives for queues kept in HOST memory, as RayQueue keeps them (ray_queue.h:161-244): the arrays are staged
 * through device memory, partitioned by the kernels above and copied back IN PLACE.  compact_hits: hits first, both parts in
 * their original order (the reference's swap loop keeps only the hits in order; callers drop the misses right after,
 * pt.cpp:119).  compact_rays: states whose pixel_id (first int) is < 0 are removed, *kept = new size.  sort_keys: counting sort
 * of the element indices 0..count-1 by h_keys[i] in [0, num_bins) -- the keys are what the caller's GetMatIDFn returned
 * (ray_queue.h:209-244) -- stable inside a bin. */
int imba_queue_compact_hits_host(imba_ray* h_rays, imba_hit* h_hits, void* h_states, int state_bytes, int count, int* hit_count);
int imba_queue_compact_rays_host(imba_ray* h_rays, void* h_states, int state_bytes, int count, int* kept);
int imba_queue_sort_keys_host(const int* h_keys, int count, int num_bins, int* h_sorted_indices);

/* ------------------------------------------------------------------------------------------
 * L2: integrators (render/integrators/{pt,vcm}.*) behind Integrator::{render,reset,preprocess,pixel_size}
 * ---------------------------------------------------------------------------------------- */
typedef enum { IMBA_ALGO_PT = 0, IMBA_ALGO_BPT = 1, IMBA_ALGO_VCM = 2, IMBA_ALGO_PPM = 3, IMBA_ALGO_VCM_PT = 4,
               IMBA_ALGO_LT = 5 } imba_algorithm; /* frontend/cmd_line.h:38-45 */

typedef struct {
    int32_t algorithm;        /* imba_algorithm */
    int32_t width, height;    /* -w -h (512x512) */
    float   fov;              /* -f (60) */
    float   cam_pos[3], cam_dir[3], cam_up[3];
    int32_t spp;              /* --spp: samples per pixel within one iteration (1) */
    int32_t max_path_len;     /* --max-path-len (10) */
    int32_t light_path_count; /* --light-path-count (w*h/2) */
    int32_t num_connections;  /* -c (1) */
    int32_t num_knn;          /* -k (10) */
    float   radius_factor;    /* -r (2) */
    uint64_t seed;            /* replaces std::random_device (ray_gen.h:49-50) so runs are reproducible */
    /* image-space sharding for multi-GPU: this rank renders sample slices
     * iteration_index*rank_count + rank of the progressive sequence. */
    int32_t rank, rank_count;
    /* PT and LT: imba_integrator_render(n) processes up to this many consecutive progressive iterations as ONE wavefront pass
     * of that many times the camera paths (PT) / light paths (LT) (0 or 1 = one iteration per pass, the reference's behaviour;
     * at most 16; ignored by BPT / VCM / PPM, whose iterations are pipelined inside one render(n) call instead).  Every path keeps the
     * seed of its own iteration, so the accumulated image is the same; larger wavefronts amortise the small late waves. */
    int32_t iterations_per_pass;
    int32_t reserved;
} imba_render_settings;

typedef struct imba_integrator imba_integrator;

imba_integrator* imba_integrator_create(imba_scene* scene, const imba_render_settings* settings);
void  imba_integrator_destroy(imba_integrator* it);
/* Integrator::preprocess (integrator.cpp:14-71 pixel-size probe; light_vertices.cpp:15-129 cache sizing). */
int   imba_integrator_preprocess(imba_integrator* it);
float imba_integrator_pixel_size(const imba_integrator* it);
/* Light-vertex slots the preprocess step allocated: LightVertices::compute_cache_size (light_vertices.cpp:15-129) -- 10 000 probe
 * light paths, margin * ceil(average stored vertices per path) * light_path_count, margin 1.1 (10 below 1000 paths); 0 for PT / LT. */
int   imba_integrator_vertex_cache_size(const imba_integrator* it);
/* Integrator::reset (vcm.h:69-72). */
int   imba_integrator_reset(imba_integrator* it);
/* Camera moved (PerspectiveCamera::move, camera.h:27-50, followed by Integrator::reset as render_window.cpp:97-101 does):
 * waits for pending work, installs the new position / direction / up and restarts the progressive schedule. */
int   imba_integrator_set_camera(imba_integrator* it, const float pos[3], const float dir[3], const float up[3]);
/* Integrator::render: runs `iterations` progressive iterations, ADDING un-normalised radiance into the
 * device accumulation buffer (the caller divides by iterations*spp, render_window.cpp:105,178). Asynchronous. */
int   imba_integrator_render(imba_integrator* it, int iterations);
/* Device pointer of the W*H*3 float accumulation buffer (for NCCL reduction by the host program). */
float* imba_integrator_framebuffer(imba_integrator* it);
/* Makes the integrator accumulate into a caller-owned device buffer (W*H*3 floats), e.g. one registered with NCCL. */
int   imba_integrator_set_framebuffer(imba_integrator* it, float* d_rgb);
/* Multi-GPU (one process per GPU, scene replicated, progressive iterations sharded over the ranks by rank / rank_count): sums the
 * W*H*3 float framebuffer over all ranks of `nccl_comm` (an ncclComm_t of the host program's NCCL) in place, on `stream` (a
 * cudaStream_t) or, if null, on the integrator's own stream after the rendering enqueued so far.  The single collective of the path. */
int   imba_integrator_allreduce(imba_integrator* it, void* nccl_comm, void* stream);
/* The cudaStream_t all work of this integrator is enqueued on; imba_integrator_sync waits for it. */
void* imba_integrator_stream(imba_integrator* it);
int   imba_integrator_sync(imba_integrator* it);
/* Synchronises and copies the accumulation buffer to host (W*H*3 floats). */
int   imba_integrator_download(imba_integrator* it, float* h_rgb);
int   imba_integrator_clear(imba_integrator* it);
/* Ray statistics like TileScheduler's counters (tile_scheduler.h:59-60,117-127): out[0]=primary, out[1]=shadow rays. */
int   imba_integrator_ray_counts(imba_integrator* it, uint64_t out[2]);

/* PT shading stage on an injected batch (parity harness for calculate_intersection, integrator.h:71-123, and
 * PathTracer::process_primary_rays, pt.cpp:82-160).  States are the reference's 40-byte PTState.
 * Outputs: continuation rays/states (compacted, order = input order), shadow rays/states (16-byte ShadowState)
 * and contributions added to the integrator's accumulation buffer.  Host pointers; synchronous. */
/* Parity harnesses for PerspectiveCamera::generate_ray (camera.h:57-65) and the primary-ray seeding (ray_gen.h:57-72). */
int imba_integrator_camera_rays(imba_integrator* it, int count, const float* raster_xy, imba_ray* h_rays);
int imba_seed_states(uint32_t seed_base, int first, int count, int spp, uint64_t* out);

/* One VCM stage on an injected batch (parity harness for vcm.cpp:193-258 and 323-439, light_vertices.h:77-84,
 * rangesearch.h:25-129).  stage 0 = process_light_rays: returns continuation queue, camera-connection shadow queue and
 * the vertex cache (112-byte records: pos|path_length, out_dir|material, normal|dVC, geom_normal|dVCM, u_tangent|dVM,
 * v_tangent|uv.x, throughput|uv.y).  stage 1 = process_camera_rays with `h_cache_in` installed as the vertex cache (photon
 * grid rebuilt for PPM / VCM).  States are the reference's 48-byte VCMState; `iteration` is 1-based.  Host pointers. */
int imba_vcm_stage_batch(imba_integrator* it, int stage, int iteration, float base_radius, int count, const imba_ray* h_rays,
                         const imba_hit* h_hits, const void* h_states, const void* h_cache_in, int cache_in_count,
                         imba_ray* h_rays_out, void* h_states_out, int* out_count, imba_ray* h_shadow_rays, void* h_shadow_states,
                         int* shadow_count, void* h_cache_out, int* cache_out_count);

/* Rays and 48-byte states produced by the light (light = 1) or camera (0) ray generation of a VCM-family integrator. */
int imba_vcm_raygen_batch(imba_integrator* it, int light, int iteration, float base_radius, uint32_t seed_base, int count, imba_ray* h_rays,
                          void* h_states);

int imba_pt_shade_batch(imba_integrator* it, int count, const imba_ray* h_rays, const imba_hit* h_hits,
                        const void* h_states, imba_ray* h_rays_out, void* h_states_out, int* out_count,
                        imba_ray* h_shadow_rays, void* h_shadow_states, int* shadow_count);

#ifdef __cplusplus
}
#endif
#endif /* IMBA_B200_H */
