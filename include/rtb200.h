/*
 * rtb200.h — C ABI of the B200-native replacement for the per-pixel hot path of
 * termuxinator/html5-canvas-raytracer (reference: main.js, build '741').
 *
 * The reference exports nothing (main.js is a module script without `export`); the replaceable
 * unit is the body of the row/column loops of `spanish(y)` (main.js:183-199), whose only
 * observable effect is filling `colorbuf.data` (a Uint8ClampedArray(W*H*4), main.js:83) before
 * `context.putImageData` (main.js:200).  Everything that loop captures from `main()`'s closure
 * becomes an explicit field below; closures (`mtl.sampler`) become enumerated sampler kinds.
 *
 * Ownership : the caller owns every pointer it passes and the output buffer; the library copies
 *             what it needs during the call and keeps no caller pointer.
 * Errors    : 0 = RTB_OK, negative = error; message via rtb_last_error() (thread-local).
 *             The library never aborts and never throws across this boundary.
 * Threading : entry points serialise on a process-wide mutex (the reference's caller is the single JS
 *             thread).  rtb_render is synchronous; rtb_render_device only enqueues, and with
 *             rtb_set_frames_in_flight(n) up to n such frames overlap on the GPU; rtb_render_batch keeps
 *             several frames in flight internally.  Streams are library-owned unless one is passed in.
 * No CPU fallback: every entry point that computes needs a CUDA device and fails with
 *             RTB_ERR_NO_DEVICE / RTB_ERR_CUDA otherwise.
 *
 * Plain C, no CUDA/torch types: this header is what a cgo / ctypes / N-API binding includes.
 */
#ifndef RTB200_H
#define RTB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_ABI_VERSION 1u
#define RTB_MAX_LIGHTS 4      /* main.js:283 lists 2 live + 2 commented-out lights */
#define RTB_MAX_SEGS 16       /* recursion depth literal is 8 at main.js:194 */
#define RTB_MAX_GROUP 8       /* frames rendered by one pass of the wavefront pipeline (rtb_render_device_group) */

/* ---- error codes --------------------------------------------------------------------------- */
enum {
  RTB_OK = 0,
  RTB_ERR_INVALID = -1,        /* bad argument / malformed scene (message says which field)      */
  RTB_ERR_NO_DEVICE = -2,      /* no CUDA device visible — there is no CPU fallback              */
  RTB_ERR_CUDA = -3,           /* a CUDA runtime call failed (message carries cudaGetErrorString) */
  RTB_ERR_OOM = -4,            /* device or host allocation failed                               */
  RTB_ERR_NOT_INIT = -5,       /* rtb_init() not called                                          */
  RTB_ERR_OVERFLOW = -6,       /* internal ray queue overflow that band splitting could not fix  */
  RTB_ERR_UNSUPPORTED = -7     /* valid request this build does not implement                    */
};

/* ---- sampler kinds: what `mtl.sampler(hit)` (main.js:320) was at scene-build time ------------ */
enum {
  RTB_SAMPLER_COLOR = 0,    /* default `return hit.m.color`            main.js:404              */
  RTB_SAMPLER_TEXTURE = 1,  /* `sampleTexture(unit, hit.u, hit.v)`     main.js:143-157, 343-351 */
  RTB_SAMPLER_CHECKER = 2,  /* home-sphere procedural checker on hit.n main.js:126-133          */
  RTB_SAMPLER_STARS = 3     /* skybox `Math.random()` stars            main.js:135-139          */
};

/* ---- precision / kernel selection ------------------------------------------------------------ */
enum {
  RTB_PRECISION_FP32 = 0,   /* throughput kernels (graded: +-2 LSB on >= 99.9 % of pixels)       */
  RTB_PRECISION_FP64 = 1    /* validation kernel, reference operation order, no FMA contraction  */
};
enum {
  RTB_KERNEL_AUTO = 0,      /* fp64, or fp32 with <= 32 spheres -> per-pixel kernel; else wavefront */
  RTB_KERNEL_PIXEL = 1,     /* one thread per pixel, explicit post-order stack                   */
  RTB_KERNEL_WAVEFRONT = 2  /* fp32 only: ballot-compacted ray queues per recursion level        */
};
enum {
  RTB_STARS_OFF = 0,        /* Math.random() stubbed to a constant >= threshold: sky is black    */
  RTB_STARS_HASH = 1,       /* deterministic per-(pixel,node) counter hash in place of Math.random;
                               NOT the reference's raster-order stream                           */
  RTB_STARS_STREAM = 2      /* exact replay: RtbRenderOpts.star_stream[n] is the value
                               of the n-th Math.random() call in the reference's own call order (rows top
                               to bottom, pixels left to right, post-order inside a pixel's tree).  Whole
                               frames on the fp64 validation kernel only: the replay needs the exact
                               ray tree (one differing hit/miss would shift every later index)     */
};
enum {
  RTB_CAMERA_REFERENCE = 0, /* main.js:187-191 verbatim: per-COMPONENT dist (bug-compatible)     */
  RTB_CAMERA_BASIS = 1      /* target = origin + axisX*dist[0] + axisY*dist[1] + axisZ*dist[2]   */
};

/* ---- scene description ----------------------------------------------------------------------- */

/* createMaterial(color, albedo, se, ri) main.js:397-406, plus the sampler closure as a kind. */
typedef struct RtbMaterial {
  double color[3];
  double albedo[5];            /* ambient, diffuse, specular, reflect, refract (main.js:400)      */
  double specular_exponent;
  double refract_index;
  int32_t sampler_kind;        /* RTB_SAMPLER_*                                                  */
  int32_t texture_unit;        /* RTB_SAMPLER_TEXTURE: index into RtbSceneDesc.textures          */
  double checker_freq[2];      /* RTB_SAMPLER_CHECKER: 5000, 2500 (main.js:129-130)              */
  double checker_color[2][3];  /* RTB_SAMPLER_CHECKER: [1,1,0],[1,0,1] (main.js:131)             */
  double star_threshold;       /* RTB_SAMPLER_STARS: 0.001 (main.js:137)                         */
  double star_gain;            /* RTB_SAMPLER_STARS: 1000  (main.js:138)                         */
} RtbMaterial;

/* createSphere(o, r, m) main.js:408-418.  The reference stores r*r, never r (main.js:412-413).
 * surface_area is only used by the host-side sort (main.js:159-163) and is not part of the ABI:
 * spheres arrive ALREADY in post-sort order — array order fixes closest-hit ties (main.js:229),
 * the self-shadow skip index (main.js:294) and the occluder loop order (main.js:293-304). */
typedef struct RtbSphere {
  double origin[3];
  double r2;
  RtbMaterial mtl;
} RtbSphere;

/* createTexture() main.js:339-341: {width,height,texels}.  texels = RGBA8, row-major, w*h*4 bytes
 * (what getImageData returns at main.js:388-391, or checkerTexture writes at main.js:353-373). */
typedef struct RtbTexture {
  int32_t width;
  int32_t height;
  const uint8_t* texels;
} RtbTexture;

typedef struct RtbSceneDesc {
  uint32_t abi_version;                 /* RTB_ABI_VERSION                                       */
  int32_t n_spheres;
  const RtbSphere* spheres;             /* post-sort order                                       */
  int32_t n_textures;
  const RtbTexture* textures;
  int32_t n_lights;                     /* 2 in the reference (main.js:283)                      */
  double lights[RTB_MAX_LIGHTS][3];     /* [[5,10,5],[5,10,0]]                                   */
  double light_intensity;               /* 50, declared once OUTSIDE the light loop (main.js:284)*/
  double eps;                           /* 0.001 (main.js:430-436, 445)                          */
} RtbSceneDesc;

/* Camera + projection: the closure variables of main() that spanish() reads (main.js:85-105). */
typedef struct RtbCamera {
  int32_t width, height;                /* canvas.width / canvas.height (main.js:79-80)          */
  double origin[3];
  double axisX[3], axisY[3], axisZ[3];
  double projW, projH, projD;           /* W/2, H/2, projW / tan(30 deg) (main.js:103-105)       */
  int32_t camera_mode;                  /* RTB_CAMERA_*                                          */
  int32_t _pad;
} RtbCamera;

typedef struct RtbRenderOpts {
  int32_t segs;          /* recursion budget passed to intersectWorld; 8 at main.js:194          */
  int32_t spp_grid;      /* g: g*g regular stratified sub-samples per pixel; 1 = reference, 4 = 16 spp */
  int32_t precision;     /* RTB_PRECISION_*                                                      */
  int32_t kernel;        /* RTB_KERNEL_*                                                         */
  int32_t star_mode;     /* RTB_STARS_*                                                          */
  int32_t y0, y1;        /* row range [y0,y1) to render; y1 <= 0 means height.  Rows outside are
                            left untouched (the reference hands back row by row, main.js:200)    */
  int32_t stripe_rows;   /* multi-GPU interleaving: rows are grouped in stripes of stripe_rows;   */
  int32_t stripe_rank;   /* this call renders stripes with (stripe_index % stripe_world) ==       */
  int32_t stripe_world;  /* stripe_rank.  stripe_world <= 1 disables striping.                    */
  int32_t out_packed;    /* 0: rows land at their natural position of a full W*H frame (others
                            untouched); 1: the rows this call renders are packed contiguously, in
                            ascending y, from the start of the buffer (the all-gather send layout) */
  uint64_t star_seed;    /* RTB_STARS_HASH seed                                                   */
  const double* star_stream;   /* RTB_STARS_STREAM: host array of Math.random() values, caller-owned   */
  uint64_t star_stream_len;    /* must cover every draw of the frame, else RTB_ERR_INVALID             */
} RtbRenderOpts;

/* Filled by every render call when non-NULL.  Device-side counts, not the oracle's. */
typedef struct RtbStats {
  double kernel_ms;          /* CUDA-event time of all kernels of the call                       */
  double h2d_ms, d2h_ms;     /* copies inside the call (0 for the device entry point)            */
  uint64_t nodes;            /* intersectWorld invocations with segs > 0 (tree nodes)            */
  uint64_t shadow_rays;      /* light iterations that passed shadow_dot > 0 (main.js:292)        */
  uint64_t sphere_tests;     /* ray-sphere tests the DEVICE executed (<= reference's count)      */
  uint32_t launches;         /* kernel launches issued                                           */
  uint32_t bands;            /* row bands the frame was split into (wavefront queues)            */
  /* CUDA-event time per kernel class of the wavefront pipeline (0 for the per-pixel kernels):   */
  double ms_trace_primary;   /* level-1 closest-hit sweep + hit processing                       */
  double ms_trace_secondary; /* levels >= 2 (reflect / refract rays)                             */
  double ms_shadow;          /* occluder sweeps, all lights                                      */
  double ms_compose;         /* post-order compose + RGBA8 resolve                               */
} RtbStats;

typedef struct RtbScene_* rtb_scene_t;

/* ---- entry points ---------------------------------------------------------------------------- */

/* Select the CUDA device (a negative value picks the current device) and create library streams.
 * Fails with RTB_ERR_NO_DEVICE when no device exists. Idempotent per process. */
int rtb_init(int device);
void rtb_shutdown(void);

/* Validate and upload a scene (spheres, materials, textures, lights).  Immutable afterwards.
 * Replaces: the `objects` array + `texture_unit*` closures of main() (main.js:107-163). */
int rtb_scene_create(const RtbSceneDesc* desc, rtb_scene_t* out);
void rtb_scene_destroy(rtb_scene_t scene);
/* Bytes rtb_scene_create copied host -> device for this scene (geometry, materials, textures, tables). */
uint64_t rtb_scene_h2d_bytes(rtb_scene_t scene);

/* Render one frame into a caller-owned HOST buffer of width*height*4 bytes (RGBA8, row-major,
 * A = 255): exactly the bytes `colorbuf.data` holds after the last `spanish` tick.
 * Replaces: redraw()/spanish() pixel loops, main.js:180-199.  H2D of the camera and D2H of the
 * framebuffer happen inside the call. */
int rtb_render(rtb_scene_t scene, const RtbCamera* cam, const RtbRenderOpts* opts,
               uint8_t* out_rgba8, RtbStats* stats_or_null);

/* Same, but `d_out_rgba8` is a DEVICE pointer on the library's device and `cuda_stream` is a
 * cudaStream_t (as void*, NULL = library stream).  With stats == NULL the call only ENQUEUES the frame
 * (no host synchronisation), so that the caller can queue its collective behind it; the one thing that
 * can still go wrong afterwards - the ray queues of the wavefront path overflowing - is reported by
 * rtb_finish().  Used by the multi-GPU host layer, which all-gathers stripes with NCCL. */
int rtb_render_device(rtb_scene_t scene, const RtbCamera* cam, const RtbRenderOpts* opts,
                      void* d_out_rgba8, void* cuda_stream, RtbStats* stats_or_null);

/* A frame GROUP on the device: nframes (1 .. RTB_MAX_GROUP) frames of a sequence - same scene, canvas size
 * and options; one camera and one device framebuffer (or packed stripe buffer) each - rendered by ONE pass
 * of the wavefront pipeline: the primary rays of all frames form level 1 and every later recursion level,
 * shadow queue and compose pass is shared.  A rank of an N-GPU row-striped sequence renders 1/N of every
 * frame; grouping N frames makes its kernels as large as those of a whole frame (no short launches, N
 * times fewer of them).  Otherwise exactly rtb_render_device: enqueue-only without stats, one verdict
 * (rtb_finish / rtb_copy_status_device) for the whole group.  Per-pixel kernels (RTB_KERNEL_PIXEL, or AUTO
 * on small scenes) render the frames one after the other.
 * Replaces: nframes consecutive redraw() calls of an animation loop around main.js:180-199. */
int rtb_render_device_group(rtb_scene_t scene, const RtbCamera* cams, int32_t nframes, const RtbRenderOpts* opts,
                            void* const* d_out_rgba8, void* cuda_stream, RtbStats* stats_or_null);

/* Wait for the scene's asynchronous frame (if any).  RTB_OK: the frame is valid.  RTB_ERR_OVERFLOW: its
 * ray queues overflowed and the frame is INVALID; the library has already reduced the band size, so
 * issuing the same rtb_render_device call again succeeds (deterministic).  A verdict that was never
 * collected is returned by the next render call on the scene instead. */
int rtb_finish(rtb_scene_t scene);

/* How many asynchronous wavefront frames (rtb_render_device with stats == NULL) the library keeps in
 * flight: 1 (default) .. 8.  With n > 1, consecutive calls rotate through n private workspaces, and the
 * call for frame k+n - not the one for frame k+1 - is what collects frame k.  Issued on different
 * streams, the short per-level kernels of neighbouring frames then run concurrently and fill each
 * other's tails (sequences; a striped multi-GPU frame is a sequence of small frames).  rtb_finish() waits
 * for every frame in flight; rtb_copy_status_device() refers to the frame enqueued last. */
int rtb_set_frames_in_flight(int32_t n);

/* Enqueue, on `cuda_stream`, a 4-byte device-to-device copy of the overflow flag of the frame just
 * enqueued for `scene` (0 = valid, non-zero = ray queues overflowed) to `d_dst_u32`.  Lets a multi-GPU
 * caller ship every rank's verdict inside the all-gather payload, so that all ranks take the same
 * retry decision without a second collective. */
int rtb_copy_status_device(rtb_scene_t scene, void* d_dst_u32, void* cuda_stream);

/* Frame batch (config C5: orbit animation): nframes cameras, nframes*W*H*4 bytes out (host). */
int rtb_render_batch(rtb_scene_t scene, const RtbCamera* cams, int32_t nframes,
                     const RtbRenderOpts* opts, uint8_t* out_rgba8, RtbStats* stats_or_null);

/* De-interleave an all-gathered stripe buffer ([rank][rank_stride_rows][W*4]) into a row-major
 * framebuffer; both device pointers.  Used after ncclAllGather in the multi-GPU path.
 * rank_stride_rows = rows each rank contributes to the gather: 0 means rtb_stripe_padded_rows(); a larger
 * value leaves room for trailer rows (the multi-GPU layer ships its status word in one). */
int rtb_unstripe_device(const void* d_gathered, void* d_frame, int32_t width, int32_t height,
                        int32_t stripe_rows, int32_t stripe_world, int32_t rank_stride_rows, void* cuda_stream);

/* ---- framebuffer exchange over peer memory (NVLink / NVSwitch), multi-GPU mode without a collective ----
 * One process per GPU.  Buffers are cudaMalloc'ed by the library and shared with CUDA IPC handles, which
 * the host layer exchanges once (torch.distributed all_gather_object, MPI, a pipe ...).
 *   rtb_peer_alloc   allocate + zero a shareable device buffer, return its 64-byte IPC handle
 *   rtb_peer_open    map another process's buffer (peer access is enabled on first use)
 *   rtb_peer_close / rtb_peer_free
 * Per rank: K framebuffers (W*H*4 bytes) and one control block of uint32 words, all opened by all peers.
 *   rtb_peer_scatter enqueue ONE kernel that stores this rank's packed stripe rows (rtb_render_device with
 *                    out_packed = 1) into every rank's framebuffer at their final rows, then writes this
 *                    rank's status word to ctl[status_offset + rank] and `seq` to ctl[ready_offset + rank]
 *                    of every peer (release order: a peer that sees seq also sees the rows and the status).
 *                    peer_frames / peer_ctl are HOST arrays of `stripe_world` device pointers, own included.
 *   rtb_peer_wait    hold the stream until n consecutive uint32 words have all reached `value` (signed
 *                    distance; sequence numbers only grow).  timeout_ms > 0: a one-warp kernel that gives
 *                    up after timeout_ms and sets *d_timeout_flag_u32 = 1 instead of hanging the GPU.
 *                    timeout_ms == 0: stream memory operations of the driver (cuStreamWaitValue32) - no
 *                    kernel, no SM slot, no timeout; falls back to the kernel (20 s) if the driver has none.
 *   rtb_peer_signal  enqueue: ctl[word_offset + rank] = value on every peer (e.g. "I am done reading
 *                    framebuffer k": the credit a producer waits for before overwriting it).
 * Replaces, for the multi-GPU mode, the ncclAllGather + rtb_unstripe_device pair. */
#define RTB_MAX_PEERS 16
#define RTB_PEER_HANDLE_BYTES 64
int rtb_peer_alloc(uint64_t bytes, void** d_ptr, uint8_t* handle64);
int rtb_peer_open(const uint8_t* handle64, void** d_ptr);
int rtb_peer_close(void* d_ptr);
int rtb_peer_free(void* d_ptr);
int rtb_peer_scatter(const void* d_local_packed, void* const* peer_frames, void* const* peer_ctl, int32_t width,
                     int32_t height, int32_t stripe_rows, int32_t stripe_rank, int32_t stripe_world,
                     const void* d_status_u32, uint32_t ready_offset, uint32_t status_offset, uint32_t seq,
                     void* cuda_stream);
/* The same for the n_frames frames of a group in ONE kernel: frame f's packed rows start at d_local_packed +
 * f * local_frame_stride_bytes and go to peer_frames[p] + f * peer_frame_stride_bytes on every peer p (both
 * strides multiples of 16); one status word and one sequence number for the whole group. */
int rtb_peer_scatter_group(const void* d_local_packed, uint64_t local_frame_stride_bytes, int32_t n_frames,
                           void* const* peer_frames, uint64_t peer_frame_stride_bytes, void* const* peer_ctl,
                           int32_t width, int32_t height, int32_t stripe_rows, int32_t stripe_rank,
                           int32_t stripe_world, const void* d_status_u32, uint32_t ready_offset,
                           uint32_t status_offset, uint32_t seq, void* cuda_stream);
int rtb_peer_signal(void* const* peer_ctl, int32_t world, int32_t rank, uint32_t word_offset, uint32_t value,
                    void* cuda_stream);
int rtb_peer_wait(const void* d_words_u32, int32_t n, uint32_t value, uint32_t timeout_ms, void* d_timeout_flag_u32,
                  void* cuda_stream);

/* Multi-GPU host delivery: copy this rank's PACKED stripe rows (device) to their natural rows of a full
 * width*height*4 HOST frame, asynchronously on `cuda_stream` (one strided 2-D copy).  With a frame in host memory
 * that every rank's process has mapped and page-locked (POSIX shared memory + rtb_host_register), the N ranks of
 * a box deliver a frame over their N PCIe links at once instead of funnelling it through one GPU's link.
 * rtb_host_register / rtb_host_unregister page-lock caller-owned host memory (cudaHostRegister, portable). */
int rtb_stripe_rows_to_host(const void* d_packed, void* h_frame, int32_t width, int32_t height, int32_t stripe_rows,
                            int32_t stripe_rank, int32_t stripe_world, void* cuda_stream);
int rtb_host_register(void* p, uint64_t bytes);
int rtb_host_unregister(void* p);

/* Number of rows rank `stripe_rank` owns / the padded per-rank row count used by the gather. */
int32_t rtb_stripe_local_rows(int32_t height, int32_t stripe_rows, int32_t stripe_rank,
                              int32_t stripe_world);
int32_t rtb_stripe_padded_rows(int32_t height, int32_t stripe_rows, int32_t stripe_world);

/* Host-side inputs that fix the result, for hosts that bind this header directly (no device work):
 *   rtb_look_at      main.js's closure-private lookAt(org, tgt, up) + the projection constants
 *                    (main.js:92-105): fills origin, axisX/Y/Z, projW = W/2, projH = H/2, projD = projW/tan(30 deg).
 *   rtb_sort_spheres objects.sort(...) of main.js:159-163 in place: ascending surface_area / distance to the
 *                    camera origin (surface_area = 4*PI*r2), stable.  rtb_scene_create expects this order. */
int rtb_look_at(const double org[3], const double tgt[3], const double up[3], int32_t width, int32_t height,
                int32_t camera_mode, RtbCamera* out);
int rtb_sort_spheres(RtbSphere* spheres, int32_t n, const double camera_origin[3]);

/* loadTexture without the browser (main.js:375-395: Image -> offscreen canvas -> getImageData), host only: PNG
 * bytes -> the RGBA8 texels an RtbTexture wants.  Call once with out_rgba8 == NULL for the size, then with a
 * buffer of width*height*4 bytes.  Non-interlaced 8-bit greyscale / RGB / palette (+ tRNS) / grey+alpha / RGBA;
 * the reference's earth.png / mars.png are 8-bit palette images (PLTE only).  zlib is bound at run time
 * (libz.so.1); without it RTB_ERR_UNSUPPORTED. */
int rtb_decode_png(const uint8_t* png, uint64_t len, int32_t* width, int32_t* height, uint8_t* out_rgba8,
                   uint64_t out_capacity);

/* Diagnostics, host only: the packed sweep tables of the wavefront kernels for rays leaving `origin` (camera or
 * light) and for general rays, exactly as the library builds them - lets a test check their safety margins
 * against float64 on the CPU.  Per sphere pair p: tab[8p..8p+7] = x0 x1 y0 y1 z0 z1 w0 w1; every buffer holds
 * rtb_debug_table_pad(n_spheres) * 4 floats. */
int32_t rtb_debug_table_pad(int32_t n_spheres);
int rtb_debug_sweep_tables(const RtbSceneDesc* desc, const double origin[3], float* origin_tab, float* origin_exact,
                           float* general_tab);

/* Reference defaults (the literals of main.js) for the option/scene scalars. */
void rtb_default_opts(RtbRenderOpts* opts);

const char* rtb_last_error(void);
const char* rtb_version(void);          /* "rtb200 <abi> sm_100a" */
uint32_t rtb_abi_version(void);
/* Device properties the bench needs without importing anything else: SM count, clock (kHz). */
int rtb_device_info(int32_t* sm_count, int32_t* clock_khz, char* name_buf, int32_t name_len);

/* Diagnostics: enqueue a probe that stores, for every SM that runs one of its blocks, the SM's cycle counter
 * and the global nanosecond timer into d_out[2*sm], d_out[2*sm + 1] (device memory, 2*max_sms u64, zero it
 * first).  Two probes around a region give the SM clock the region ran at (delta cycles / delta ns) without
 * any driver query inside the region. */
int rtb_clock_probe(uint64_t* d_out, int32_t max_sms, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* RTB200_H */
