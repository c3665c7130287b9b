/* rt_b200.h — C ABI of the B200-native replacement for raytracer_lib's per-pixel render loop.
 *
 * This is the drop-in boundary: exactly the calls a Rust `RayTracer` shim would bind to
 * replace raytracer_lib/src/raytracer/mod.rs. Every entry point cites the reference interface it replaces
 * (paths relative to the upstream repository Andreas-Edling/raytracer-rs).
 *
 * Conventions
 *   - return value 0 (RT_OK) = success, negative = error; the message is available from rt_last_error().
 *   - every input buffer is caller-owned and copied during the call; every output buffer is caller-allocated.
 *   - a handle may be moved between host threads (the reference moves RayTracer into its render thread,
 *     raytracer/src/main.rs:194-196) but must not be called concurrently. No global mutable state.
 *   - all arithmetic on the path is IEEE binary32 with the reference's operation order; there is no CPU
 *     fallback: every render entry point fails with RT_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_OK 0
#define RT_ERR_INVALID -1  /* bad argument */
#define RT_ERR_LOAD -2     /* Collada / texture loading failed (message mirrors SceneLoadError's Display) */
#define RT_ERR_CUDA -3     /* CUDA runtime error or no usable device */
#define RT_ERR_UNSUPPORTED -4

/* raytracer_lib/src/raytracer/accel_intersect/oct_tree_intersector.rs:12 (re-exported lib.rs:7) */
#define RT_DEFAULT_TRIANGLES_PER_LEAF 70
#define RT_DEVICE_NONE (-2)

typedef struct rt_scene rt_scene;         /* flattened scene on the host  (scene/mod.rs:24-29 `Scene`) */
typedef struct rt_raytracer rt_raytracer; /* device-resident renderer     (raytracer/mod.rs:32-47 `RayTracer`) */

/* ---- flattened scene description (structure-of-arrays) ------------------------------------------------ */

enum { RT_DIFFUSE_COLOR = 0, RT_DIFFUSE_TEXTURE = 1 }; /* scene/color.rs:98-101 `Diffuse` */

typedef struct rt_material { /* scene/mod.rs:63-69 `Material`; only `diffuse` is used by the path */
    int32_t kind;            /* RT_DIFFUSE_COLOR / RT_DIFFUSE_TEXTURE */
    float rgb[3];            /* Diffuse::Color */
    uint32_t texture_id;     /* Diffuse::TextureId */
} rt_material;

typedef struct rt_light { /* scene/mod.rs:12-22 `Light` */
    float pos[3];
    float color[3];
} rt_light;

typedef struct rt_texture { /* scene/texture.rs:6-10 `Texture`, texels already byte/256 (texture.rs:40-46) */
    uint32_t width, height;
    const float* rgb; /* width*height*3 */
} rt_texture;

typedef struct rt_scene_desc {
    uint32_t num_triangles;
    const float* vertices;    /* num_triangles*9: v0 v1 v2 of every triangle, geometry order then triangle
                                 order (Geometry::transformed_vertices, scene/mod.rs:46-61) */
    const uint32_t* tri_geom; /* num_triangles: geometry index per triangle, non-decreasing */
    uint32_t num_geometries;
    const rt_material* materials; /* num_geometries */
    uint32_t num_lights;
    const rt_light* lights;
    uint32_t num_textures;
    const rt_texture* textures;
    float camera_orientation[16]; /* cameras[0]: ColladaMatrix::to_vecmath_matrix of its node (row-vector convention) */
    float camera_fov_deg;         /* xfov */
    /* Reserved per-vertex attributes, both may be NULL. The reference parses the NORMAL and TEXCOORD inputs of a
       <triangles> element but consumes only the positions (scene/loaders/colladaloader.rs:587-593): normals are
       recomputed per hit from the vertices (raytracer/mod.rs:198-205) and texture lookups use the hit's barycentrics
       (raytracer/mod.rs:246). The path therefore IGNORES these arrays, exactly as the reference does; they are part of
       the structure so that a host which keeps them (the Rust loader does, in its Collada DOM) has a slot to pass them
       and a future shading model needs no ABI change. Layout when present: normals num_triangles*9 (one xyz per
       vertex, same order as `vertices`), uvs num_triangles*6 (one uv per vertex). */
    const float* normals;
    const float* uvs;
} rt_scene_desc;

/* ---- configuration ------------------------------------------------------------------------------------ */

enum { RT_ACCEL_OCTREE = 0, /* flattened reference octree, traversal bit-identical to oct_tree_intersector.rs:148-272 */
       RT_ACCEL_BVH = 1,    /* binary SAH BVH, closest hit + lowest-index tie break + root-cube acceptance */
       RT_ACCEL_CWBVH = 2,  /* compressed 8-wide BVH (same hit rules as RT_ACCEL_BVH; 80-byte nodes, costlier box decode) */
       RT_ACCEL_BVH4 = 3,   /* 4-wide BVH, full-precision boxes, one 128-byte node per visit (same hit rules) */
       RT_ACCEL_LBVH = 4 }; /* binary BVH built ON THE GPU (Morton codes, radix sort, Karras hierarchy, refit); same
                               traversal kernel and hit rules as RT_ACCEL_BVH */
enum { RT_JITTER_FIXED_HALF = 0, /* xi = (0.5, 0.5): the pinned parity mode */
       RT_JITTER_HASHED = 1 };   /* xi = hash(seed, pixel, sample, axis) * 2^-24: stands in for StdRng::from_os_rng
                                    (raytracer/mod.rs:84, scene/camera.rs:82-84) */

typedef struct rt_config {
    uint32_t width, height;
    uint32_t triangles_per_leaf; /* create_raytracer's `triangles_per_leaf` (lib.rs:15); 0 = default 70 */
    uint32_t rows_per_call;      /* rows traced by rt_trace_frame_additive; 0 = 50 (raytracer/mod.rs:87) */
    int32_t recursions;          /* RECURSIONS (mod.rs:81): reference value 2; 0 = primary + shadow only */
    uint32_t sub_spread;         /* SUB_SPREAD (mod.rs:82): reference value 1 */
    int32_t jitter_mode;         /* RT_JITTER_* */
    uint32_t seed;
    int32_t accel;               /* RT_ACCEL_* */
    int32_t device;              /* CUDA device ordinal; -1 = current device; RT_DEVICE_NONE = host-side handle
                                    (construction, camera, acceleration-structure introspection; render calls
                                    fail with RT_ERR_CUDA — used by CPU-only tests of the host logic) */
    /* image sharding for one-process-per-GPU rendering: this instance owns the bands b (of `band_rows` rows)
       with b % shard_count == shard_index. shard_count = 0 or 1 = whole image. */
    uint32_t shard_index, shard_count, band_rows;
} rt_config;

/* Fills *cfg with the reference's defaults (recursions 2, spread 1, hashed jitter, 50 rows, 70 triangles/leaf). */
void rt_config_default(rt_config* cfg, uint32_t width, uint32_t height);

/* ---- scene loading: replaces scene/loaders (ColladaLoader::from_file / from_str, loaders/colladaloader.rs:22-46) -- */

int rt_scene_load_file(const char* collada_filename, rt_scene** out, char* err, size_t err_len);
int rt_scene_load_str(const char* collada_doc, const char* data_dir /* may be NULL */, rt_scene** out, char* err,
                      size_t err_len);
/* Pointers inside *desc stay valid until rt_scene_free. */
int rt_scene_get_desc(const rt_scene* scene, rt_scene_desc* desc);
void rt_scene_free(rt_scene* scene);

/* ---- construction: replaces lib.rs:15-44 ------------------------------------------------------------- */

/* create_raytracer(collada_doc, triangles_per_leaf, width, height) -> Result<RayTracer, String>   (lib.rs:15-20) */
int rt_create_raytracer(const char* collada_doc, size_t triangles_per_leaf, size_t width, size_t height,
                        rt_raytracer** out, char* err, size_t err_len);
/* create_raytracer_from_file(collada_filename, triangles_per_leaf, width, height)                  (lib.rs:22-27) */
int rt_create_raytracer_from_file(const char* collada_filename, size_t triangles_per_leaf, size_t width, size_t height,
                                  rt_raytracer** out, char* err, size_t err_len);
/* build_raytracer(scene, ...) (lib.rs:29-44) on an already flattened scene: the FFI entry a Rust host that keeps
   its own Collada loader calls. */
int rt_create(const rt_scene_desc* scene, const rt_config* cfg, rt_raytracer** out, char* err, size_t err_len);
void rt_destroy(rt_raytracer* rt);
const char* rt_last_error(const rt_raytracer* rt);

/* Change pinned-mode switches after construction (rebuilds nothing except when `accel` changes). */
int rt_configure(rt_raytracer* rt, int32_t recursions, uint32_t sub_spread, int32_t jitter_mode, uint32_t seed,
                 int32_t accel);
int rt_set_rows_per_call(rt_raytracer* rt, uint32_t rows);

/* ---- the render loop: replaces raytracer/mod.rs:80-128 ------------------------------------------------ */

/* RayTracer::trace_frame_additive(&mut self) -> u32   (mod.rs:80-117): traces `rows_per_call` rows starting at the
   internal current row (wrapping modulo height), accumulates into the device film, returns rows*width. */
int rt_trace_frame_additive(rt_raytracer* rt, uint32_t* num_primary_rays);
/* Same loop over an explicit row range, `spp` passes: the whole-frame / multi-sample launch used by the benchmark.
   n_primary / n_shadow (may be NULL) receive the rays issued by this call (shadow: one per (hit, light) with
   n.l >= 0, mod.rs:218-226). */
int rt_trace_rows(rt_raytracer* rt, uint32_t first_row, uint32_t n_rows, uint32_t spp, uint64_t* n_primary,
                  uint64_t* n_shadow);
/* RayTracer::get_tonemapped_pixels(&self) -> Vec<u32>   (mod.rs:120-128): width*height 0xAARRGGBB, caller-allocated. */
int rt_get_tonemapped_pixels(rt_raytracer* rt, uint32_t* out);
/* Incremental form of the same readback for a host that keeps ONE frame buffer across calls, which is what the reference's
   render loop amounts to (trace_frame_additive traces 50 rows, get_tonemapped_pixels converts the whole frame,
   raytracer/src/main.rs:200-201: 2.07 M pixel conversions and 8.3 MB for 96 000 traced pixels at 1080p). `out` must be the
   buffer passed to the previous rt_get_tonemapped_pixels_delta call on this handle, unmodified since; then only the rows
   whose pixels can have changed since that call (rows traced, film cleared or replaced) are copied — 384 KB instead of
   8.3 MB after a 50-row call — and the buffer again equals what rt_get_tonemapped_pixels would deliver. A different (or
   first) buffer receives the whole frame. The contract is the caller's: a buffer that was freed and reallocated at the
   same address, or scribbled over, keeps its stale rows. */
int rt_get_tonemapped_pixels_delta(rt_raytracer* rt, uint32_t* out);
/* Pipelined form of the same readback, for a host that double-buffers frames the way the reference's render and GUI
   threads do (raytracer/src/main.rs:200-209): takes a device-side snapshot of the packed frame on the render stream
   (a few microseconds) and copies it to `pinned_out` (page-locked host memory) on a separate copy stream, so the
   device -> host transfer overlaps the NEXT trace call. Returns at once; the pixels are valid after rt_wait_pixels.
   Two copies may be in flight per handle (two snapshots): a third call waits, on the device, for the first to leave its snapshot. */
int rt_get_tonemapped_pixels_async(rt_raytracer* rt, uint32_t* pinned_out);
int rt_wait_pixels(rt_raytracer* rt);
/* Waits until at most `keep` of the copies started by rt_get_tonemapped_pixels_async are still in flight (they complete in order).
   With keep = 1 and two alternating host buffers the host hands frame k to the copy engine BEFORE it waits for frame k-1, so the
   PCIe link never idles while the host gets around to its next call (the library holds two snapshots for this). */
int rt_wait_pixels_keep(rt_raytracer* rt, uint32_t keep);
/* Film::clear (film.rs:37-41) through the pub field `film` (raytracer/src/main.rs:126). */
int rt_film_clear(rt_raytracer* rt);
/* Film contents: width*height*7 floats per pixel: sum rgb, sum of squares rgb, num_samples (film.rs:3-7). */
int rt_get_film(rt_raytracer* rt, float* out);
/* Overwrites the film (`pub film: Film` with `pub pixel_datas`, film.rs:27-29, is writable in the reference too) from the
   layout rt_get_film returns; the packed frame is recomputed from it. num_samples must be an integer in [0, 2^32). */
int rt_set_film(rt_raytracer* rt, const float* in);
/* Film::get_estimated_variances(&self) -> Vec<RGB>   (film.rs:50-67): width*height*3 floats, per pixel and channel
   (sum_sq / (n (n-1)) - sum^2 / (n * n (n-1))) * 50, with n (n-1) evaluated in wrapping u32 arithmetic before the
   conversion to f32 as the reference's release build does. Pixels with n <= 1 are NaN (x/0 - y/0), as in the reference. */
int rt_get_estimated_variances(rt_raytracer* rt, float* out);
/* Parity hook: global triangle index (geometry order, then triangle order) of the last primary hit per pixel,
   0xFFFFFFFF for a miss. */
int rt_get_primary_ids(rt_raytracer* rt, uint32_t* out);

/* ---- camera: the pub field `camera` (scene/camera.rs:63-78, used raytracer/src/main.rs:125-161) -------- */

int rt_camera_move_rel(rt_raytracer* rt, float x, float y, float z);
int rt_camera_add_x_angle(rt_raytracer* rt, float radians);
int rt_camera_add_y_angle(rt_raytracer* rt, float radians);
/* out34: rotation_matrix[16], orientation_matrix[16], max_x, max_y */
int rt_camera_get(const rt_raytracer* rt, float* out34);
/* Direct state injection for hosts that keep Camera in their own language: angles + position, then
   update_matrices() (camera.rs:92-98). Uploads 128 bytes to the device on the next trace call. */
int rt_camera_set_state(rt_raytracer* rt, float x_angle, float y_angle, const float pos[3]);

/* The map behind the perspective grid of the camera rays (RT_TUNE_CAMERA_GRID), i.e. the inverse of Camera::get_ray
   (camera.rs:80-90) for the handle's current view: out12[0..9) = A row major, out12[9..12) = the ray origin. A point p = origin + t * dir
   of the camera ray of pixel (u, v) with sub-pixel offsets (xi1, xi2) has (X, Y, Z) = A (p - origin) with X / Z = u + xi1, Y / Z = v + xi2
   and Z = t. Needs no device. */
int rt_get_camera_plane_matrix(rt_raytracer* rt, double* out12);

/* ---- device-side access for multi-GPU gather and for timing on a caller-owned stream ------------------- */

/* All later launches / copies of this handle are issued on `cuda_stream` (a cudaStream_t; NULL = default stream). */
int rt_set_stream(rt_raytracer* rt, void* cuda_stream);
/* Device pointer to the width*height packed LDR frame kept current by the trace kernel's epilogue. */
int rt_get_ldr_device_ptr(rt_raytracer* rt, void** dev_ptr);
/* Redirect the LDR stores of the rows this instance owns into `dev_ptr` (e.g. a peer-mapped framebuffer of
   rank 0, or a torch tensor): the fused trace+gather path. NULL restores the internal buffer. */
int rt_set_ldr_target(rt_raytracer* rt, void* dev_ptr);
/* Zero-copy readback: register a page-locked (cudaHostAlloc / cudaHostRegister, mapped) host buffer of width*height
   u32. From then on the trace kernel's epilogue also stores every packed pixel straight into that buffer over PCIe
   while it renders, and rt_get_tonemapped_pixels(rt, same pointer) only synchronises the stream (no device->host
   copy after the fact). NULL unregisters. Rows this handle never traces keep whatever the buffer held. */
int rt_set_host_frame(rt_raytracer* rt, uint32_t* pinned_host_frame);
/* Copies only the rows owned by this shard, compacted (owned rows in ascending order), into dev_out. */
int rt_get_owned_ldr_rows_device(rt_raytracer* rt, void* dev_out, uint32_t* n_rows);
/* Device buffers owned by the library (plain cudaMalloc, therefore exportable over CUDA IPC), and IPC plumbing for
   the one-process-per-GPU gather: rank 0 exports its frame buffer, the other ranks map it and pass the mapped
   pointer to rt_set_ldr_target, so their trace kernel stores packed pixels straight into rank 0's memory over
   NVLink. handle64 is a cudaIpcMemHandle_t (64 bytes). */
int rt_device_alloc(rt_raytracer* rt, size_t bytes, void** dev_ptr);
int rt_device_free(rt_raytracer* rt, void* dev_ptr);
int rt_ipc_export(rt_raytracer* rt, void* dev_ptr, uint8_t* handle64);
int rt_ipc_open(rt_raytracer* rt, const uint8_t* handle64, void** dev_ptr);
int rt_ipc_close(rt_raytracer* rt, void* dev_ptr);
/* Cross-GPU frame fence for the fused gather (all stream-ordered on the handle's stream, no host synchronisation):
   rt_stream_signal_flag stores `value` into *dev_flag (system scope) once everything enqueued before it has finished;
   rt_stream_wait_flags holds the stream until every one of dev_flags[0..n_flags) (uint32, typically
   peer-mapped memory of rank 0) has reached `target` (wrap-safe >=); in the same launch it can first store `target`
   into dev_flags[signal_slot] and afterwards into dev_flags[release_slot] (-1 = none), which is rank 0's whole
   per-frame fence. rt_stream_signal_then_wait is the per-frame fence of every other rank in one launch: store `value`
   into *dev_signal_flag ("my stores of this frame are done"), then hold the stream until *dev_wait_flag has reached
   `target` ("the frame that last used the next buffer has been read"). A wait gives up after ~2 s and counts a timeout
   (rt_sync_timeouts) instead of hanging the GPU. */
int rt_stream_signal_flag(rt_raytracer* rt, void* dev_flag, uint32_t value);
int rt_stream_signal_then_wait(rt_raytracer* rt, void* dev_signal_flag, uint32_t value, void* dev_wait_flag, uint32_t target);
int rt_stream_wait_flags(rt_raytracer* rt, void* dev_flags, uint32_t n_flags, uint32_t target, int32_t signal_slot,
                         int32_t release_slot);
int rt_sync_timeouts(rt_raytracer* rt, uint32_t* count);
/* The same fence without kernel launches: stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32, fetched from the
   driver at run time). rt_stream_write_value stores `value` into *dev_flag once everything enqueued before it has finished, behind a
   system-wide memory barrier; rt_stream_wait_value holds the stream until (int32)(*dev_flag - value) >= 0. The stream's front end
   executes them, no SM is occupied. Unlike the flag kernels a wait has NO timeout: a peer that dies leaves the stream waiting until
   the process is torn down. RT_ERR_UNSUPPORTED when the driver does not offer them for this address. */
int rt_stream_write_value(rt_raytracer* rt, void* dev_flag, uint32_t value);
int rt_stream_wait_value(rt_raytracer* rt, void* dev_flag, uint32_t value);
/* Fuses the "my stores of this frame are done" half of that fence into the NEXT trace call (one-shot): the last warp out of
   the call's last kernel stores `value` into *dev_flag at system scope, after every warp has fenced its own stores into
   the peer-mapped frame — no separate signal launch, and the flag is on its way while the launch drains. Calls whose last
   kernel does not finish the pixels itself (sample planes, bounce wavefront, the one-thread-per-pixel variant) append
   the signal as a launch of its own; either way the flag is published exactly once per armed call. */
int rt_set_done_signal(rt_raytracer* rt, void* dev_flag, uint32_t value);
/* Host-side frame delivery for one process per GPU: every rank copies the rows IT owns over its OWN PCIe link into one frame in
   host memory that all processes share (POSIX shared memory, page-locked in every process with rt_host_register), instead of
   funnelling the whole frame through rank 0's link. rt_copy_owned_rows copies the owned bands of `src_frame` (a device frame this
   handle stored its pixels into, e.g. an rt_set_ldr_target buffer) to the same rows of `dst_frame` (one strided 2-D copy) on
   `cuda_stream` (NULL = the handle's stream); rt_signal_flag_on_stream stores `value` into a flag (device address of mapped host
   memory or device memory) behind it in stream order, which is how a rank tells the consumer that its rows of frame k have arrived.
   rt_host_register returns the device address of the registered range through *dev_ptr (may be NULL). */
int rt_host_register(rt_raytracer* rt, void* host_ptr, size_t bytes, void** dev_ptr);
int rt_host_unregister(rt_raytracer* rt, void* host_ptr);
int rt_copy_owned_rows(rt_raytracer* rt, const void* src_frame, void* dst_frame, void* cuda_stream);
int rt_signal_flag_on_stream(rt_raytracer* rt, void* dev_flag, uint32_t value, void* cuda_stream);
/* Device pointer to the 4 uint64 ray counters of the LAST trace call: shadow rays, primary hits, bounce rays, blocked
   shadow rays. Two sets alternate from call to call (the kernel of one call zeroes the set of the next, so a call needs no
   memset): ask again after every trace call. */
int rt_get_counters_device_ptr(rt_raytracer* rt, void** dev_ptr);
/* Bytes of per-launch parameters (camera + pointers) that travel host -> device with every trace launch. */
uint32_t rt_launch_param_bytes(void);
/* Developer tuning knobs (results never change, only the schedule). RT_TUNE_KERNEL_VARIANT: 1 (default) persistent
   warps pulling 8x4 pixel tiles from an atomic queue, while-while traversal; 2 ray pool — lanes are decoupled from
   pixels through per-warp shared-memory ray rings (binary BVH, recursions 0, one light; other configurations run
   variant 1); 0 one thread per pixel, single-loop. */
#define RT_TUNE_KERNEL_VARIANT 0
/* RT_TUNE_TILE_SCHEDULE: 1 (default) the persistent kernel hands out tiles heaviest-first using the cycle counts
   recorded by the previous launch of the same view (longest-processing-time-first); 0 image order. */
#define RT_TUNE_TILE_SCHEDULE 1
/* RT_TUNE_POOL_REFILL: idle lanes of a warp that trigger a refill from the ray rings (1..32, default 16).
   RT_TUNE_POOL_BLOCKS: resident 256-thread blocks per SM of the ray-pool kernel (0 = as many as fit). */
#define RT_TUNE_POOL_REFILL 2
#define RT_TUNE_POOL_BLOCKS 3
/* RT_TUNE_POOL_MIN_INNER: the ray-pool kernel leaves its inner-node loop when fewer lanes than this are still
   descending while other lanes wait at a leaf or with a finished ray (0..32, default 8; 0 = classic while-while). */
#define RT_TUNE_POOL_MIN_INNER 4
/* RT_TUNE_MULTI_SAMPLE_LAUNCH: how rt_trace_rows handles spp > 1. 1 (default): even spp on the persistent kernel use
   sample lanes — the 32 lanes of a warp item trace 8, 4 or 2 samples of 4, 8 or 16 pixels and add them to the film in
   sample order (spp/8, spp/4 or spp/2 launches, no intermediate buffer); everything else traces all samples in one launch
   into per-sample radiance planes followed by one ordered accumulation pass. 2: always planes. 0: one launch per
   sample. Same film in every mode. */
#define RT_TUNE_MULTI_SAMPLE_LAUNCH 5
/* RT_TUNE_BOUNCE_WAVEFRONT: 1 (default) bounce rays (recursions > 0) run as a wavefront — the hits of every level are
   compacted into a dense list and the next level's rays fill whole warps; 0 every pixel walks its bounce tree depth
   first inside the trace kernel. Same rays, same film. */
#define RT_TUNE_BOUNCE_WAVEFRONT 6
/* RT_TUNE_SPLIT_QUARTERS (variant 1, BVH kernels): tiles whose recorded cost exceeds T = (balanced launch time) * q / 4 are
   handed out as 4 (cost <= 4T), 8 (<= 8T) or 16 items; 0 = never split (default 4). */
#define RT_TUNE_SPLIT_QUARTERS 7
/* RT_TUNE_QUEUE_BATCH / RT_TUNE_QUEUE_BATCH_FROM (variant 1): in the cheap tail of the cost-sorted tile queue — from
   `from` percent of its length on — a warp claims `batch` slots with one atomic (defaults 4 and 33; batch 1 = off). */
#define RT_TUNE_QUEUE_BATCH 8
#define RT_TUNE_QUEUE_BATCH_FROM 9
/* RT_TUNE_TIME_LAUNCHES: 1 (default) every trace call records two CUDA events around its kernels so that
   rt_launch_stats.trace_kernel_ms is available; 0 skips them (trace_kernel_ms reads 0) — two stream operations less per
   call for hosts that time frames themselves. */
#define RT_TUNE_TIME_LAUNCHES 10
/* RT_TUNE_MIN_SCHEDULE_TILES: launches with fewer 32-lane tiles than this (default 4096) get no cost-feedback schedule of their own and
   run in image order (measured on the reference's 50-row bands, which walk over the image with a period of lcm(50, height) rows —
   108 launch geometries at 1080p: per-band schedules 2.0 ms per 22 bands, image order 1.5 ms, every tile handed out in parts
   1.6 ms; such a launch is bound by its chain of cold-cache node fetches, not by its heaviest tile — hence RT_TUNE_BAND_LOOKAHEAD).
   RT_TUNE_MAX_SPLIT_LEVEL: finest split of a heavy tile in a scheduled launch, 0 never, 1 / 2 / 3 = up to 4 / 8 / 16 items of
   8 / 4 / 2 pixels (default 3; the finer levels only engage when a tile alone costs more than 4x / 8x the balanced launch time,
   i.e. when the launch cannot fill the GPU). */
#define RT_TUNE_MIN_SCHEDULE_TILES 11
#define RT_TUNE_MAX_SPLIT_LEVEL 12
/* RT_TUNE_BOUNCE_STREAM: 1 (default) every level of the bounce wavefront runs as a ray stream on the binary BVH — bounce rays and the
   shadow rays of their hits share the lanes of a warp, and lanes whose ray ended are refilled RT_TUNE_STREAM_REFILL (1..32, default
   16) at a time; 0 the lockstep form (one kernel traces 32 bounce rays per warp, a second shades the compacted hits). Same rays, same
   film. Other structures always use the lockstep form. */
#define RT_TUNE_BOUNCE_STREAM 13
#define RT_TUNE_STREAM_REFILL 14
/* RT_TUNE_STREAM_MIN_INNER: the ray-stream kernel leaves its inner-node loop when fewer lanes than this are still descending while
   other lanes wait at a leaf (0..32, default 8; 0 = classic while-while rounds). */
#define RT_TUNE_STREAM_MIN_INNER 15
/* RT_TUNE_STREAM_BLOCKS: resident 256-thread blocks per SM the ray-stream kernel is compiled and launched for (3, 4 or 5: a register
   budget of 80, 64 or 48; default 4). RT_TUNE_WF_BLOCKS: cap on the resident blocks per SM of the lockstep wavefront kernels (0 = as many as fit). */
#define RT_TUNE_STREAM_BLOCKS 16
#define RT_TUNE_WF_BLOCKS 17
/* RT_TUNE_STREAM_CHAIN: 1 (default) when every bounce level from the second on sends one ray per hit (the reference's RECURSIONS = 2,
   SUB_SPREAD = 1), a lane of the ray-stream kernel that completes a hit continues in place with that hit's bounce ray, so one launch
   walks the whole bounce tree; 0 one launch per level. */
#define RT_TUNE_STREAM_CHAIN 18
/* RT_TUNE_BAND_LOOKAHEAD: 1 (default) rt_trace_frame_additive traces a lap ahead — the first band call of a lap traces the next
   sample of every row down to the bottom of the image in one launch (a 50-row launch cannot fill the GPU and lasts as long as its
   slowest chain of cold-cache node fetches), every band call commits its rows from that plane; film, ids, frame and per-call ray
   counts after every call equal band-by-band tracing, and anything that would make the plane stale (camera, film, configuration,
   rt_trace_rows) drops it. 0: every call traces its own rows. */
#define RT_TUNE_BAND_LOOKAHEAD 19
/* RT_TUNE_FILM_PREFETCH: L2 prefetches of the trace kernel, for frames that start with a cold cache. 0 off; 1 a pixel's film sums
   are requested before its rays are traced (they are read when the sample is added, after the traversal; with a cold film the add
   otherwise waits for DRAM); 2 also the sums of squares of a pixel whose camera ray hit; 3 (default) the launch also requests the
   whole binary BVH (nodes, triangles, shading records) and the tile queue up front instead of discovering them level by level, one
   DRAM round trip per level of the first rays. Measured with the L2 flushed between frames: 0.1535 / 0.1500 / 0.1491 / 0.1464 ms
   per frame (requesting the whole film up front as well: 0.1522 ms, not kept). */
#define RT_TUNE_FILM_PREFETCH 20
/* RT_TUNE_FILM_PREFETCH_ROWS_MB (with RT_TUNE_FILM_PREFETCH = 3): launches that need a pixel's sample number before they can form its rays
   (hashed sub-pixel offsets, bounce rays) read the film record first thing; such a launch requests the records of all its rows up front
   when they are at most this many MB (default 48: a 1080p frame is 33 MB; 0 = never). */
#define RT_TUNE_FILM_PREFETCH_ROWS_MB 21
/* RT_TUNE_CAMERA_GRID (binary BVH structures, persistent kernel): camera rays all leave one point, so the sample plane of Camera::get_ray
   is itself an index. 2..5 (default 3): the plane is cut into cells of 4..32 pixels, every cell lists the triangles whose projection (plus
   a margin of one pixel) can reach it — rebuilt on the device whenever the camera, the resolution or the tree changes (three small
   kernels and one 4-byte readback) — and a camera ray tests its cell's list instead of walking the tree (same Moller-Trumbore test, same
   tie rule, same root-cube acceptance: the same hit). Shadow and bounce rays walk the tree. 0 = camera rays walk the tree as well. */
#define RT_TUNE_CAMERA_GRID 22
/* RT_TUNE_CAMERA_GRID_AFTER: launches a view must have seen without changing before its grid is built (default 1: the second frame of a
   view builds it; 0 = at once). A build costs about as much as a frame, so a camera that moves every frame keeps walking the tree. */
#define RT_TUNE_CAMERA_GRID_AFTER 23
/* RT_TUNE_LIGHT_GRID (launches that use the camera grid): shadow rays all end at their light, so the direction in which the light sees
   the shaded point is an index as well. 6..9 (default 8): a cube of grids of 64..512 cells per face edge around every point light (up to
   four; built once per tree) lists per cell the triangles that direction can meet, and a shadow ray tests that list (same test, same
   decision as the tree walk: blocked iff the closest hit has 0.01 < t < 1). A ray long enough to reach a surface lying BEYOND the light
   within its last hundredth walks the tree. 0 = all shadow rays walk the tree. */
#define RT_TUNE_LIGHT_GRID 24
int rt_set_tuning(rt_raytracer* rt, int32_t key, int32_t value);
/* Launch statistics of the last rt_trace_rows / rt_trace_frame_additive call. */
typedef struct rt_launch_stats {
    uint32_t kernels_launched; /* CUDA kernels launched by the call */
    float trace_kernel_ms;     /* device time of the trace kernel(s), CUDA events on the launching stream */
    uint64_t n_primary, n_shadow, n_bounce;
} rt_launch_stats;
int rt_get_launch_stats(const rt_raytracer* rt, rt_launch_stats* out);
/* Schedule introspection: SM cycles the last scheduled launch spent per 32-lane tile (a split tile reports parts x its slowest part),
   for the most recently used launch geometry; *n_items = entries of its sorted queue (0 before the first sort). Any pointer may be NULL.
   What the critical-path figures of DESIGN.md section 7 are computed from. */
int rt_get_tile_costs(rt_raytracer* rt, uint32_t* out, uint32_t capacity, uint32_t* n_tiles, uint32_t* n_items);
/* Rays issued by this handle since creation: out3 = primary, shadow, bounce (synchronises the stream). Exact totals
   over any number of asynchronous trace calls. */
int rt_get_ray_totals(rt_raytracer* rt, uint64_t* out3);
/* Total kernels launched by this handle since creation. */
uint64_t rt_kernels_launched(const rt_raytracer* rt);

/* ---- acceleration-structure introspection (parity of the build, oct_tree_intersector.rs:66-146) -------- */

/* out[6]: nodes, inner nodes, leaves, empty leaves, triangle references, depth */
int rt_octree_stats(const rt_raytracer* rt, uint64_t* out);
/* cubes: nodes*6 (min xyz, max xyz); first_child: nodes (-1 = leaf); leaf_offset: nodes+1; leaf_tris: references.
   Any pointer may be NULL. Returns the number of triangle references through *n_refs. */
int rt_octree_export(const rt_raytracer* rt, float* cubes, int32_t* first_child, uint32_t* leaf_offset,
                     uint32_t* leaf_tris, uint64_t* n_refs);
/* out[4]: bvh nodes, leaves, max leaf size, depth */
int rt_bvh_stats(const rt_raytracer* rt, uint64_t* out);
/* boxes: nodes*12 (child0 lo xyz, hi xyz, child1 lo xyz, hi xyz); children: nodes*2 (>= 0 inner node, < 0 leaf with
   ~child = first triangle slot); counts: nodes*2 (triangles of a leaf child); tri_order: slot -> global triangle. */
int rt_bvh_export(const rt_raytracer* rt, float* boxes, int32_t* children, int32_t* counts, uint32_t* tri_order);
/* GPU tree build (RT_ACCEL_LBVH): (re)builds the tree on the device and reports out3 = nodes, depth, triangles and
   the device time of the build in milliseconds (CUDA events). Either output may be NULL. */
int rt_lbvh_build(rt_raytracer* rt, uint64_t* out3, float* build_ms);
/* The GPU-built tree in the layout of rt_bvh_export (boxes: nodes*12, children/counts: nodes*2, tri_order: slot ->
   global triangle); nodes = max(triangles - 1, 1). Nodes that were folded into a leaf of their parent stay in the
   array but are unreachable. */
int rt_lbvh_export(rt_raytracer* rt, float* boxes, int32_t* children, int32_t* counts, uint32_t* tri_order);
/* 4-wide BVH (RT_ACCEL_BVH4). out[4]: nodes, leaves, max leaf size, depth */
int rt_bvh4_stats(const rt_raytracer* rt, uint64_t* out);
/* boxes: nodes*24 (per child: lo xyz, hi xyz; an empty slot is lo = hi = +inf); children: nodes*4 (>= 0 inner node,
   < 0 leaf with ~child = first triangle slot); counts: nodes*4; tri_order: slot -> global triangle. */
int rt_bvh4_export(const rt_raytracer* rt, float* boxes, int32_t* children, int32_t* counts, uint32_t* tri_order);
/* compressed 8-wide BVH (RT_ACCEL_CWBVH). out[4]: nodes, leaf children, triangle slots, depth */
int rt_cwbvh_stats(const rt_raytracer* rt, uint64_t* out);
/* node_words: nodes*20 u32 (five 16-byte words per node, layout in csrc/cwbvh_build.cpp); tri_order: slot -> global
   triangle. Either pointer may be NULL. */
int rt_cwbvh_export(const rt_raytracer* rt, uint32_t* node_words, uint32_t* tri_order);

/* ---- stats.rs / timing crate stand-ins ---------------------------------------------------------------- */

typedef struct rt_stats rt_stats;             /* raytracer_lib/src/stats.rs:1-40 `Stats` */
rt_stats* rt_stats_new(void);                 /* Stats::new */
void rt_stats_free(rt_stats* s);
/* Stats::stats(num_primary_rays) -> "fps: {}  primary rays/s: {}" */
int rt_stats_stats(rt_stats* s, uint32_t num_primary_rays, char* out, size_t out_len);
/* Stats::mean_stats() -> "mean fps: {}  mean primary rays/s: {}" */
int rt_stats_mean_stats(const rt_stats* s, char* out, size_t out_len);

typedef struct rt_benchmark rt_benchmark;     /* timing/src/lib.rs:6-59 `BenchMark` */
rt_benchmark* rt_benchmark_new(void);
void rt_benchmark_free(rt_benchmark* b);
int rt_benchmark_start(rt_benchmark* b, const char* name);
int rt_benchmark_stop(rt_benchmark* b, const char* name); /* RT_ERR_INVALID for an unknown name (the reference panics) */
/* Display impl (timing/src/lib.rs:95-109): "{name} total: {}ms, mean: {}ms, samples: {}\n", sorted by total, descending */
int rt_benchmark_report(const rt_benchmark* b, char* out, size_t out_len);

/* Library identification: "rt_b200 <version> sm_100a". */
const char* rt_version(void);
/* First 16 hex digits of the sha256 of the kernel sources (csrc/kernels.cu + csrc/device_types.h) this library was built from.
   Profiler counts quoted by the benchmark (profiles/traffic.json) carry the digest of the build they were measured on. */
const char* rt_kernels_hash(void);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
