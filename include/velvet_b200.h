/*
 * velvet_b200.h -- C ABI of the B200-native XPBD cloth solver (drop-in for the hot path of
 * vitalight/Velvet: VtClothSolverGPU::Simulate + SpatialHashGPU::Hash + VtBuffer).
 *
 * Two layers, both plain C (pointers + sizes + PODs, no C++/torch types):
 *
 *  1. The KERNEL SEAM: one entry point per free function that the reference's host code
 *     (VtClothSolverGPU.hpp, SpatialHashGPU.hpp) calls into its .cu files.  Same argument order
 *     and meaning as the reference declarations cited on each prototype; glm::vec3* becomes
 *     float* (packed xyz, 12-byte stride), glm::mat4 becomes const float[16] column-major.
 *     All pointers must be device-accessible (cudaMalloc / cudaMallocManaged).  Launches are
 *     asynchronous on the seam stream (default: the legacy default stream, as in the reference).
 *
 *  2. The OBJECT SURFACE: an opaque handle mirroring class VtClothSolverGPU
 *     (VtClothSolverGPU.hpp L23-231) and SpatialHashGPU (SpatialHashGPU.hpp L15-60): cloth
 *     registration, constraint append, collider update, Simulate(), public sim buffers.
 *     Simulate() on a handle runs the fused sm_100a pipeline (SoA float4 state, tile-fused
 *     deterministic Jacobi iteration, own radix sort, one CUDA graph per frame).
 *
 * Error convention (reference: print + exit(EXIT_FAILURE), helper_cuda.h L566-579): every
 * function returns 0 on success or a negative VelvetStatus and never exits; the message is
 * available from velvet_last_error() (thread-local).
 */
#ifndef VELVET_B200_H
#define VELVET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VELVET_API __declspec(dllexport)
#else
#define VELVET_API __attribute__((visibility("default")))
#endif

/* ------------------------------------------------------------------ PODs (layout == reference) */

/* Common.hpp L19-47, sizeof == 80.  A translation unit that already sees the reference's own (global-namespace) struct
 * VtSimParams -- the kernel-level drop-in shim, velvet_b200/csrc/dropin/VelvetB200Shim.cpp -- defines
 * VELVET_B200_USE_REFERENCE_SIMPARAMS before including this header; the prototypes below then take the reference's type,
 * whose layout is the same (static_assert in the shim). */
#ifndef VELVET_B200_USE_REFERENCE_SIMPARAMS
typedef struct VtSimParams {
    int32_t numSubsteps;           /* 0  */
    int32_t numIterations;         /* 4  */
    int32_t maxNumNeighbors;       /* 8  */
    float maxSpeed;                /* 12 */
    float gravity[3];              /* 16 */
    float bendCompliance;          /* 28 */
    float damping;                 /* 32 */
    float relaxationFactor;        /* 36 */
    float longRangeStretchiness;   /* 40 */
    float collisionMargin;         /* 44 */
    float friction;                /* 48 */
    uint8_t enableSelfCollision;   /* 52 (C++ bool) */
    uint8_t _pad[3];
    int32_t interleavedHash;       /* 56 */
    uint32_t numParticles;         /* 60 */
    float particleDiameter;        /* 64 */
    float deltaTime;               /* 68 */
    float particleDiameterScalar;  /* 72 */
    float hashCellSizeScalar;      /* 76 */
} VtSimParams;
#endif

/* Common.hpp L113-118 */
typedef enum VtColliderType { VT_COLLIDER_SPHERE = 0, VT_COLLIDER_PLANE = 1, VT_COLLIDER_CUBE = 2 } VtColliderType;

/* VtClothSolverGPU.cuh L8-18, sizeof == 196. */
typedef struct VtSDFCollider {
    int32_t type;              /* 0   */
    float position[3];         /* 4   */
    float scale[3];            /* 16  */
    float deltaTime;           /* 28  */
    float curTransform[9];     /* 32  mat3 column-major (upper-left of the model matrix) */
    float invCurTransform[16]; /* 68  mat4 column-major */
    float lastTransform[16];   /* 132 mat4 column-major */
} VtSDFCollider;

/* SpatialHashGPU.cuh L7-15, sizeof == 24. */
typedef struct VtHashParams {
    uint32_t numObjects;
    uint32_t maxNumNeighbors;
    float cellSpacing;
    float cellSpacing2;
    int32_t tableSize;
    float particleDiameter2;
} VtHashParams;

typedef enum VelvetStatus {
    VELVET_OK = 0,
    VELVET_ERR_INVALID_ARGUMENT = -1,
    VELVET_ERR_CUDA = -2,
    VELVET_ERR_STATE = -3,
    VELVET_ERR_UNSUPPORTED = -4
} VelvetStatus;

VELVET_API const char* velvet_last_error(void);
VELVET_API int velvet_version(void);
/* Fills *p with the reference's host-side defaults (Common.hpp L21-46). */
VELVET_API int velvet_default_params(VtSimParams* p);

/* ------------------------------------------------------------------ 1. kernel seam */

/* VtClothSolverGPU.cuh L99 / .cu L23-28.  Copies the params; later seam calls use the copy. */
VELVET_API int velvet_SetSimulationParams(const VtSimParams* hostParams);
/* .cuh L101 / .cu L30-40 */
VELVET_API int velvet_InitializePositions(float* positions, int start, int count, const float* modelMatrix16);
/* .cuh L103-107 / .cu L42-63 */
VELVET_API int velvet_PredictPositions(float* predicted, float* velocities, const float* positions, float deltaTime);
/* .cuh L109-116 / .cu L65-115 */
VELVET_API int velvet_SolveStretch(float* predicted, float* deltas, int* deltaCounts, const int* stretchIndices,
                                   const float* stretchLengths, const float* invMasses, unsigned numConstraints);
/* .cuh L120-128 / .cu L117-203 */
VELVET_API int velvet_SolveBending(float* predicted, float* deltas, int* deltaCounts, const unsigned* bendingIndices,
                                   const float* bendingAngles, const float* invMass, unsigned numConstraints,
                                   float deltaTime);
/* .cuh L130-139 / .cu L205-251 */
VELVET_API int velvet_SolveAttachment(float* predicted, float* deltas, int* deltaCounts, const float* invMass,
                                      const int* attachParticleIDs, const int* attachSlotIDs,
                                      const float* attachSlotPositions, const float* attachDistances,
                                      int numConstraints);
/* .cuh L141 / .cu L253-270 */
VELVET_API int velvet_ApplyDeltas(float* predicted, float* deltas, int* deltaCounts);
/* .cuh L143-148 / .cu L289-327.  predicted may alias positions (pre-stabilisation pass). */
VELVET_API int velvet_CollideSDF(float* predicted, const VtSDFCollider* colliders, const float* positions,
                                 unsigned numColliders, float deltaTime);
/* .cuh L150-156 / .cu L329-386 (includes the trailing ApplyDeltas) */
VELVET_API int velvet_CollideParticles(float* deltas, int* deltaCounts, float* predicted, const float* invMasses,
                                       const unsigned* neighbors, const float* positions);
/* .cuh L158-162 / .cu L388-417 */
VELVET_API int velvet_Finalize(float* velocities, float* positions, const float* predicted, float deltaTime);
/* .cuh L164-168 / .cu L419-465 */
VELVET_API int velvet_ComputeNormal(float* normals, const float* positions, const unsigned* indices,
                                    unsigned numTriangles);
/* SpatialHashGPU.cuh L17-25 / .cu L159-196.  cellStart/cellEnd hold tableSize entries. */
VELVET_API int velvet_HashObjects(unsigned* particleHash, unsigned* particleIndex, unsigned* cellStart,
                                  unsigned* cellEnd, unsigned* neighbors, const float* positions,
                                  const float* originalPositions, VtHashParams params);

/* Stand-alone stable LSD radix sort of (key,value) pairs on bits [0,endBit) -- the replacement for the
 * reference's cub::DeviceRadixSort::SortPairs call (SpatialHashGPU.cu L133-157); sorts in place. */
VELVET_API int velvet_SortPairs(unsigned* keys, unsigned* values, unsigned numItems, int endBit);

/* Stream used by the seam functions (a cudaStream_t); NULL = legacy default stream. */
VELVET_API int velvet_seam_set_stream(void* cudaStream);
/* Self-test of the library's IEEE division (vt_div and vec3 / scalar, vt_math.cuh): for device arrays x, y of n floats writes
 * the bits of vt_div(x[i], y[i]), of (x[i], x[i+1], x[i+2]) / y[i] (indices mod n, 3 words per i) and of the compiler's own
 * x[i] / y[i].  All three must agree bit for bit with IEEE-754 division (tests/test_seam_gpu.py). */
VELVET_API int velvet_selftest_division(const float* x, const float* y, unsigned n, unsigned* outDiv, unsigned* outVec3,
                                        unsigned* outPlain);
/* Self-test of the checked-fast constraint evaluators (stretch_eval_u / bend_eval_u / vt_sqrt_u, vt_math.cuh) against the
 * branchy evaluators: `operands` holds n records of 18 floats (p0..p3 xyz, w0..w3, rest, xpbd compliance term; the stretch
 * test uses p0, p1, w0, w1, rest).  mismatches3 / fast3 are device arrays of 3 counters {stretch, bend, sqrt}: results that
 * differ while the validity predicate held (must be 0) and how often it held; the sqrt test sweeps all 2^32 operands. */
VELVET_API int velvet_selftest_constraints(const float* operands, unsigned n, unsigned long long* mismatches3,
                                           unsigned long long* fast3);
VELVET_API int velvet_device_synchronize(void);

/* VtAllocBuffer / VtFreeBuffer (Common.cuh L66-78): managed memory, plus explicit copies. */
VELVET_API int velvet_alloc(void** devPtr, size_t bytes);
VELVET_API int velvet_free(void* devPtr);
VELVET_API int velvet_copy(void* dst, const void* src, size_t bytes); /* cudaMemcpyDefault, synchronous */

/* ------------------------------------------------------------------ 2. object surface */

typedef struct VelvetSolver VelvetSolver;

/* Names follow the public members of VtClothSolverGPU (hpp L209-231) and SpatialHashGPU (hpp L54-60). */
typedef enum VelvetBufferId {
    VELVET_BUF_POSITIONS = 0,       /* float3  */
    VELVET_BUF_NORMALS,             /* float3  */
    VELVET_BUF_INDICES,             /* uint    */
    VELVET_BUF_VELOCITIES,          /* float3  */
    VELVET_BUF_PREDICTED,           /* float3  */
    VELVET_BUF_DELTAS,              /* float3  */
    VELVET_BUF_DELTACOUNTS,         /* int     */
    VELVET_BUF_INVMASSES,           /* float   */
    VELVET_BUF_STRETCHINDICES,      /* int[2S] */
    VELVET_BUF_STRETCHLENGTHS,      /* float   */
    VELVET_BUF_BENDINDICES,         /* uint[4B]*/
    VELVET_BUF_BENDANGLES,          /* float   */
    VELVET_BUF_ATTACHPARTICLEIDS,   /* int     */
    VELVET_BUF_ATTACHSLOTIDS,       /* int     */
    VELVET_BUF_ATTACHDISTANCES,     /* float   */
    VELVET_BUF_ATTACHSLOTPOSITIONS, /* float3  */
    VELVET_BUF_NEIGHBORS,           /* uint, column-major [i + N*k] */
    VELVET_BUF_INITIALPOSITIONS,    /* float3  */
    VELVET_BUF_PARTICLEHASH,        /* uint    */
    VELVET_BUF_PARTICLEINDEX,       /* uint    */
    VELVET_BUF_CELLSTART,           /* uint    */
    VELVET_BUF_CELLEND,             /* uint    */
    VELVET_BUF_SDFCOLLIDERS,        /* VtSDFCollider */
    VELVET_BUF_COUNT
} VelvetBufferId;

typedef enum VelvetPipeline {
    VELVET_PIPELINE_FUSED = 0, /* SoA float4 + tile-fused Jacobi + CUDA graph (default)           */
    VELVET_PIPELINE_SEAM = 1   /* the reference's launch sequence over the seam kernels (A/B, debug) */
} VelvetPipeline;

typedef enum VelvetMathMode {
    VELVET_MATH_EXACT = 0, /* default: no FMA contraction, IEEE division / sqrt -- bit-identical to the CPU oracle (oracle/)
                              at any horizon, hence within north_star's 1-frame and 60-frame tolerances              */
    VELVET_MATH_FAST = 1   /* opt-in: FMA contraction + approximate division / sqrt in the float kernels; deterministic,
                              ~1e-5 x extent from the oracle after a frame (tolerance 1e-4), ~35 % faster Jacobi
                              iteration; like the reference's own build it is not comparable beyond ~20 frames of a
                              contact-rich scene                                                                      */
} VelvetMathMode;

typedef enum VelvetIterateMode {
    VELVET_ITERATE_AUTO = 0,  /* default: the implicit-grid Jacobi kernel when every registered cloth is a grid carrying exactly
                                 the constraints VtClothObjectGPU generates (VtClothObjectGPU.hpp L75-132), else the tile kernel */
    VELVET_ITERATE_TILES = 1, /* always the record-driven tile kernel (any mesh)                                                */
    VELVET_ITERATE_GRID = 2   /* reported by velvet_solver_iterate_kernel only                                                  */
} VelvetIterateMode;

/* VtClothSolverGPU::Start (hpp L27-33): numParticles = 0.  params may be NULL (defaults).
 * device < 0 keeps the current CUDA device. */
VELVET_API int velvet_solver_create(VelvetSolver** out, int device, const VtSimParams* params);
VELVET_API int velvet_solver_destroy(VelvetSolver* s);
/* Global::simParams of this instance; the caller may edit it between frames (ImGui sliders / ModifyParameter). */
VELVET_API VtSimParams* velvet_solver_params(VelvetSolver* s);
VELVET_API int velvet_solver_set_pipeline(VelvetSolver* s, int pipeline);
/* Fused pipeline only: VelvetMathMode.  The spatial hash is bit-exact in both modes. */
VELVET_API int velvet_solver_set_math_mode(VelvetSolver* s, int mode);
/* Fused pipeline only: VELVET_ITERATE_AUTO or VELVET_ITERATE_TILES.  Both kernels are bit-identical. */
VELVET_API int velvet_solver_set_iterate_mode(VelvetSolver* s, int mode);
/* Writes the Jacobi kernel the next frame will run (VELVET_ITERATE_TILES or VELVET_ITERATE_GRID) to *kernel. */
VELVET_API int velvet_solver_iterate_kernel(VelvetSolver* s, int* kernel);
/* Fused pipeline only: particles per Jacobi tile (0 = default, else 128 / 256 / 512). */
VELVET_API int velvet_solver_set_tile_size(VelvetSolver* s, int particlesPerTile);

/* AddCloth (hpp L114-156): vertices = host float[3*numVertices] in model space, indices = host uint[numIndices].
 * Writes the particle offset of the new cloth to *offset. */
VELVET_API int velvet_solver_add_cloth(VelvetSolver* s, const float* vertices, int numVertices,
                                       const unsigned* indices, int numIndices, const float* modelMatrix16,
                                       float particleDiameter, int* offset);
VELVET_API int velvet_solver_add_stretch(VelvetSolver* s, int idx1, int idx2, float distance);     /* L158-163 */
VELVET_API int velvet_solver_add_attach_slot(VelvetSolver* s, const float* slotPos3);              /* L165-168 */
VELVET_API int velvet_solver_add_attach(VelvetSolver* s, int particleIndex, int slotIndex, float distance); /* L170-176 */
VELVET_API int velvet_solver_add_bend(VelvetSolver* s, unsigned idx1, unsigned idx2, unsigned idx3,
                                      unsigned idx4, float angle);                                  /* L178-185 */
/* UpdateColliders (hpp L187-205) with the SDFCollider structs already marshalled by the caller. */
VELVET_API int velvet_solver_update_colliders(VelvetSolver* s, const VtSDFCollider* colliders, int numColliders);
/* Convenience marshal of one collider exactly as hpp L195-203 does (invCurTransform = glm::inverse(cur)). */
VELVET_API int velvet_make_collider(int type, const float* position3, const float* scale3, const float* curTransform16,
                                    const float* lastTransform16, float deltaTime, VtSDFCollider* out);
/* Simulate (hpp L56-111): one frame of 1/60 s.  Asynchronous unless sync != 0 (the reference always syncs). */
VELVET_API int velvet_solver_simulate(VelvetSolver* s, int sync);
/* New overload named by the spec: Simulate(dt). */
VELVET_API int velvet_solver_simulate_dt(VelvetSolver* s, float frameTime, int sync);
VELVET_API int velvet_solver_synchronize(VelvetSolver* s);
/* SpatialHashGPU::Hash(predicted) on the solver's own hash (hpp L83). */
VELVET_API int velvet_solver_hash(VelvetSolver* s);

/* The same rebuild through the fused pipeline's own hash kernels (float4 state, own radix sort, reordered tag-filtered
 * neighbour cache) -- what velvet_solver_simulate runs internally -- on the public `predicted` buffer; results in the same
 * public hash buffers, bit-identical to velvet_solver_hash.  Synchronous. */
VELVET_API int velvet_solver_hash_fused(VelvetSolver* s);

/* Device pointer + element count (elements of the type listed at VelvetBufferId) of a public buffer. */
VELVET_API int velvet_solver_buffer(VelvetSolver* s, int bufferId, void** devPtr, size_t* count);
/* Copy a public buffer to / from host memory (count elements of the buffer's type, synchronous). */
VELVET_API int velvet_solver_download(VelvetSolver* s, int bufferId, void* host, size_t bytes);
VELVET_API int velvet_solver_upload(VelvetSolver* s, int bufferId, const void* host, size_t bytes);
/* Renderer hand-off, VtClothSolverGPU.hpp L107-110 (positions.sync(); normals.sync() into the cloths' GL vertex buffers,
 * VtBuffer.hpp L122-236).  The caller registers, per cloth (in AddCloth order), device arrays it owns -- the pointers it got
 * from cudaGraphicsResourceGetMappedPointer for that cloth's VBOs, or any device allocation of 3 floats per vertex;
 * sync_render_targets mirrors the cloth's range of positions / normals into them on the solver stream.  NULL detaches. */
VELVET_API int velvet_solver_set_render_targets(VelvetSolver* s, int clothIndex, float* positionsDev, float* normalsDev);
VELVET_API int velvet_solver_sync_render_targets(VelvetSolver* s);
/* The five SpatialHashGPU arrays in managed memory, host-indexable like the reference's VtBuffers (SpatialHashGPU.hpp
 * L54-60), instead of plain device memory.  Call before velvet_solver_add_cloth. */
VELVET_API int velvet_solver_set_hash_host_readable(VelvetSolver* s, int on);
/* Debug guard (cf. VtClothSolverCPU::CheckNAN, VtClothSolverCPU.hpp L407-418): number of non-finite components in positions /
 * velocities / predicted, and the first offending particle (numParticles when none).  Synchronous. */
VELVET_API int velvet_solver_check_nan(VelvetSolver* s, unsigned* nonFiniteCount, unsigned* firstParticle);
/* MouseGrabber (MouseGrabber.hpp L31-110) on the device.  The reference walks every position on the HOST to find the vertex
 * under the mouse ray and edits positions / velocities / invMass through managed memory; these three calls do the same work
 * in kernels on the solver stream (the camera math that turns a mouse position into a ray, L112-130, stays with the caller).
 *   grab:    FindClosestVertexToRay (L92-110) + pin (L46-52): *grabbedIndex = picked particle or -1, *distanceToOrigin its
 *            distance along the ray (FLT_MAX when none); the particle's inverse mass becomes 0.  Synchronous (returns the pick).
 *   drag:    UpdateGrappedVertex (L66-79) for the ray of this frame; no-op when nothing is grabbed.  Asynchronous.
 *   release: mouse-up branch (L57-62): the inverse mass is restored.  Asynchronous. */
VELVET_API int velvet_solver_grab(VelvetSolver* s, const float* rayOrigin3, const float* rayDirection3, int* grabbedIndex,
                                  float* distanceToOrigin);
VELVET_API int velvet_solver_drag(VelvetSolver* s, const float* rayOrigin3, const float* rayDirection3);
VELVET_API int velvet_solver_release(VelvetSolver* s);
/* Asynchronous read-back of positions+normals on the solver stream into pinned host memory (headless
 * replacement of positions.sync()/normals.sync(), hpp L109-110). */
VELVET_API int velvet_solver_readback_async(VelvetSolver* s, float* hostPositions, float* hostNormals);
/* Double-buffered variant: the results are snapshot on the solver stream into one of two device staging buffers and copied
 * to pinned host memory on a separate copy stream, so the transfer of frame k overlaps the simulation of frame k+1 (the
 * next velvet_solver_simulate may be issued at once).  *ticket (0 or 1) identifies the transfer for
 * velvet_solver_readback_wait, which blocks the host until that transfer has landed.  At most two may be outstanding:
 * wait for ticket t before the host buffers passed with it are reused. */
VELVET_API int velvet_solver_readback_pipelined(VelvetSolver* s, float* hostPositions, float* hostNormals, int* ticket);
VELVET_API int velvet_solver_readback_wait(VelvetSolver* s, int ticket);
/* The cudaStream_t the solver launches on. */
VELVET_API void* velvet_solver_stream(VelvetSolver* s);
/* Number of kernel launches (graph kernel nodes included) issued by the last Simulate call. */
VELVET_API int velvet_solver_last_launch_count(VelvetSolver* s);
/* Per-stage GPU milliseconds of the last *timed* frame under the reference's labels (GUI.cpp L32-51);
 * labels/ms arrays of capacity cap; returns the number of stages written. Timing runs outside the graph. */
VELVET_API int velvet_solver_simulate_timed(VelvetSolver* s, const char** labels, float* ms, int cap);

/* ---- inputs either side of the path: Scene.hpp / VtClothObjectGPU.hpp / Transform.hpp */
/* GenerateClothMesh (Scene.hpp L131-168): vertices float[3*(R+1)^2], indices uint[6*R^2]. */
VELVET_API int velvet_generate_cloth_mesh(int resolution, float* vertices, unsigned* indices);
/* Transform::matrix() (Transform.hpp L22-29, Helper.cpp L8-15): T * Ry * Rz * Rx * S, degrees. */
VELVET_API int velvet_transform_matrix(const float* position3, const float* rotationDeg3, const float* scale3,
                                       float* out16);
/* VtClothObjectGPU::Start (VtClothObjectGPU.hpp L43-58): AddCloth + stretch + attach + bend generation. */
VELVET_API int velvet_cloth_object_start(VelvetSolver* s, int resolution, const float* vertices,
                                         const unsigned* indices, const float* modelMatrix16,
                                         const int* attachedIndices, int numAttached, int* offset);

/* Batched independent cloths (BASELINE config 4, "batched instances shard with no communication"): numInstances copies
 * of one grid cloth of `resolution`, instance i placed by modelMatrices16 + 16*i.  Instances never interact (own hash-table
 * rows, own neighbour lists, own attach slots) and share one constraint set built from instance 0.  Must be the only
 * registration call on the handle; fused pipeline only.  Particle p of instance i is global particle i*(R+1)^2 + p. */
VELVET_API int velvet_solver_add_cloth_instances(VelvetSolver* s, int resolution, const float* vertices,
                                                 const unsigned* indices, const float* modelMatrices16, int numInstances,
                                                 const int* attachedIndices, int numAttached);

/* ---- One large cloth decomposed over several GPUs (north_star mode 2).  Every rank (one process per GPU) registers the
 * same cloth on its own handle and calls velvet_solver_dd_setup(rank, world).
 *
 * Strip form (a single grid cloth, peer transport below): rank r owns a contiguous range of particle rows -- a contiguous
 * index range, "decomposed by particle index range" -- and the Jacobi kernel exchanges the boundary rows itself; dd_info then
 * reports tile ROWS in tileBegin / tileEnd / numTiles and no staging buffers.
 *
 * Tile form (any mesh; the stepped schedule): rank r owns a contiguous range of the Morton-ordered Jacobi tiles.  It is set
 * up when first needed (velvet_solver_dd_prepare_stepped, dd_offsets or the first dd_step).  A frame is driven step by step
 * with velvet_solver_dd_step; the caller moves the bytes:
 *   ITERATE_OWNED  -> send sendBuf[sendOffsets[q] .. sendOffsets[q+1]) to rank q, receive recvBuf[recvOffsets[q] ..) from q
 *   ITERATE_FINISH
 *   GATHER_PACK    -> all-gather gatherSend (maxOwnedCount float4 per rank) into gatherRecv -> GATHER_UNPACK
 * (torch.distributed/NCCL in velvet_b200/decomposed.py).  Results are bit-identical to the single-GPU solver. */
typedef enum VelvetDDOp {
    VELVET_DD_FRAME_BEGIN = 0,   /* arg unused, farg = frame time                                   */
    VELVET_DD_SUBSTEP_BEGIN = 1, /* arg = substep: [hash rebuild] + collide of the owned particles + pack; follow with
                                    the same exchange as an iteration and ITERATE_FINISH              */
    VELVET_DD_ITERATE_OWNED = 2, /* Jacobi iteration on the owned tiles + pack of the boundary      */
    VELVET_DD_ITERATE_FINISH = 3,/* unpack of the received halo                                      */
    VELVET_DD_GATHER_PACK = 4,
    VELVET_DD_GATHER_UNPACK = 5,
    VELVET_DD_SUBSTEP_END = 6,   /* arg = substep: Finalize (+ next Predict / export)                */
    VELVET_DD_FRAME_END = 7      /* normals                                                          */
} VelvetDDOp;
typedef struct VelvetDDInfo {
    int rank, world;
    unsigned tileBegin, tileEnd, numTiles;
    unsigned ownedCount, maxOwnedCount; /* particles owned by this rank / the largest such count        */
    unsigned sendTotal, recvTotal;      /* float4 elements in sendBuf / recvBuf                          */
    void *sendBuf, *recvBuf, *gatherSend, *gatherRecv; /* device pointers, float4 elements              */
} VelvetDDInfo;
VELVET_API int velvet_solver_dd_setup(VelvetSolver* s, int rank, int world);
VELVET_API int velvet_solver_dd_info(VelvetSolver* s, VelvetDDInfo* out);
/* per-peer element offsets into sendBuf / recvBuf: arrays of world + 1 entries */
VELVET_API int velvet_solver_dd_offsets(VelvetSolver* s, unsigned* sendOffsets, unsigned* recvOffsets);
/* Sets up the tile form (exchange lists, staging buffers) if dd_setup chose the strip form; dd_info is valid for it afterwards. */
VELVET_API int velvet_solver_dd_prepare_stepped(VelvetSolver* s);
VELVET_API int velvet_solver_dd_step(VelvetSolver* s, int op, int arg, float farg);
/* NVLink peer-memory transport (the fast path; velvet_b200/csrc/dd_peer.cuh).  After dd_setup every rank exports a blob of
 * CUDA IPC handles (its two predicted-position arrays + a flag array), the caller all-gathers the blobs (any transport:
 * they are `velvet_dd_peer_blob_bytes()` plain bytes) and hands all `world` of them, rank-major, to dd_peer_import.  From
 * then on velvet_solver_dd_simulate runs a whole frame as ONE CUDA graph per rank: boundary particles are stored straight
 * into the peers' arrays by this library's kernels and ordered by release/acquire epoch flags over NVLink -- no host, no
 * NCCL in the loop.  Every rank must call dd_simulate the same number of times.  A peer that never arrives makes the wait
 * kernels time out (20 s): the next synchronising call returns VELVET_ERR_STATE.  dd_peer_close unmaps the peers; all ranks
 * must have closed before any of them destroys its solver.  Results are bit-identical to velvet_solver_simulate. */
VELVET_API size_t velvet_dd_peer_blob_bytes(void);
VELVET_API int velvet_solver_dd_peer_export(VelvetSolver* s, void* blob);
VELVET_API int velvet_solver_dd_peer_import(VelvetSolver* s, const void* blobs, size_t blobBytes);
VELVET_API int velvet_solver_dd_peer_close(VelvetSolver* s);
VELVET_API int velvet_solver_dd_simulate(VelvetSolver* s, float deltaTime, int sync);
/* Host-only: shape of the Jacobi tile plan of a grid cloth (what the fused pipeline builds at registration), for tests and
 * tuning without a GPU.  perTile4 (may be NULL) receives {nOwned, nHalo, nStretch, nBend} for the first capacityTiles tiles;
 * globals4 (may be NULL) = {maxLocals, maxKS, maxKB, maxBendPerTile}. */
VELVET_API int velvet_plan_grid_tiles(int resolution, int tileSize, unsigned* numTiles, unsigned* perTile4, unsigned capacityTiles,
                                      unsigned* globals4);
/* Host-only: FNV-1a digest of every array of the tile plan of a grid cloth (with two attach slots on vertices {0, R} when
 * withAttach != 0).  The plan is built by worker threads (VELVET_PLAN_THREADS, default: the hardware concurrency); the digest
 * must not depend on their number (tests/test_capi_cpu.py). */
VELVET_API int velvet_plan_grid_digest(int resolution, int tileSize, int withAttach, unsigned long long* digest);
/* Host-only: the grid-cloth recognition the solver runs before choosing its Jacobi kernel (grid_plan.hpp).  clothCounts =
 * particles of every AddCloth call, in order.  *recognised = 1 when the stretch / bending lists are exactly the pattern
 * VtClothObjectGPU generates for square grids of those sizes (VtClothObjectGPU.hpp L75-132); then *numTiles = tiles of the
 * implicit-grid kernel and rest4 (may be NULL; 4 floats per particle) receives, per vertex, the rest lengths of the
 * (vertical, horizontal, diagonal, anti-diagonal) stretch constraints generated there.  whyNot (may be NULL, 128 bytes)
 * receives the reason otherwise. */
VELVET_API int velvet_grid_plan_check(const unsigned* clothCounts, int numCloths, const int* stretchIndices, const float* stretchLengths,
                                      size_t numStretch, const unsigned* bendIndices, const float* bendAngles, size_t numBend,
                                      int* recognised, unsigned* numTiles, float* rest4, char* whyNot);
/* Host-only: shared-memory wavefronts per Jacobi iteration of the constraint threads' 16-byte accesses (position loads and
 * slot stores) for the tile plan of a grid cloth: out3 = {minimum, with records in constraint-id order, with the emitted
 * bank-conflict-avoiding order}. */
VELVET_API int velvet_plan_grid_smem_wavefronts(int resolution, int tileSize, unsigned long long* out3);
/* Host-only: the exchange lists of `rank` for a grid cloth of `resolution` cut into `world` ranks with tiles of
 * `tileSize` particles (what dd_setup computes), for tests without a GPU.  ids arrays may be NULL to query counts only:
 * counts[q] / counts[world + q] = number of particle ids sent to / received from rank q. */
VELVET_API int velvet_dd_plan_grid(int resolution, int tileSize, int rank, int world, unsigned* counts, unsigned* sendIds,
                                   unsigned* recvIds, unsigned* ownedRange2);

/* ---- SpatialHashGPU as its own object (SpatialHashGPU.hpp L15-60) */
typedef struct VelvetSpatialHash VelvetSpatialHash;
VELVET_API int velvet_hash_create(VelvetSpatialHash** out, float particleDiameter, int maxNumObjects,
                                  float hashCellSizeScalar, int maxNumNeighbors);
VELVET_API int velvet_hash_destroy(VelvetSpatialHash* h);
/* SetInitialPositions: device-accessible float3 array of count elements. */
VELVET_API int velvet_hash_set_initial_positions(VelvetSpatialHash* h, const float* positions, size_t count);
/* Hash(positions): device-accessible float3 array of count elements. */
VELVET_API int velvet_hash_hash(VelvetSpatialHash* h, const float* positions, size_t count);
VELVET_API int velvet_hash_buffer(VelvetSpatialHash* h, int bufferId, void** devPtr, size_t* count);

#ifdef __cplusplus
}
#endif
#endif /* VELVET_B200_H */
