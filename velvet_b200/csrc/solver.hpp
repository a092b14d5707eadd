// solver.hpp -- headless C++17 mirror of the reference's host classes for the hot path:
//   velvet::SpatialHashGPU     <-> Velvet::SpatialHashGPU     (SpatialHashGPU.hpp L15-60)
//   velvet::VtClothSolverGPU   <-> Velvet::VtClothSolverGPU   (VtClothSolverGPU.hpp L23-231)
//   velvet::VtClothObjectGPU   <-> Velvet::VtClothObjectGPU   (VtClothObjectGPU.hpp L12-149, constraint generation)
// Same member names, argument meaning and buffer layouts; no GL, no ECS, no globals: each solver owns its
// VtSimParams (the reference's process-global Global::simParams), its CUDA stream and its device.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "dd_peer.cuh"
#include "fused_kernels.cuh"
#include "grid_plan.hpp"
#include "input_kernels.cuh"
#include "radix_sort.cuh"
#include "seam.hpp"
#include "tile_plan.hpp"
#include "vt_buffer.hpp"

namespace velvet {

using uint = unsigned int;

void default_sim_params(VtSimParams& p);  // Common.hpp L21-46
constexpr float kFixedDeltaTime = 1.0f / 60.0f;  // Timer.hpp L235

class SpatialHashGPU {
public:
    // SpatialHashGPU.hpp L18-28 (hashCellSizeScalar / maxNumNeighbors come from Global::simParams there)
    // hostReadable: the five hash arrays in managed memory like the reference's VtBuffers (SpatialHashGPU.hpp L54-60) instead
    // of plain device memory (the default: 4.3 GB of neighbour slots per 16.7M particles stay out of the unified-memory pool)
    SpatialHashGPU(float particleDiameter, int maxNumObjects, float hashCellSizeScalar, int maxNumNeighbors, bool hostReadable = false);

    // L32-39: snapshot of `count` packed float3 positions (device-accessible or host pointer)
    void SetInitialPositions(const float* positions, size_t count);
    // L41-52: runs the seam HashObjects on `count` packed float3 positions
    void Hash(const float* positions, size_t count, float particleDiameter, cudaStream_t stream);
    VtHashParams MakeParams(size_t count, float particleDiameter) const;

    VtBuffer<uint> neighbors;
    VtBuffer<vec3> initialPositions;
    VtBuffer<uint> particleHash;
    VtBuffer<uint> particleIndex;
    VtBuffer<uint> cellStart;
    VtBuffer<uint> cellEnd;

    float spacing() const { return m_spacing; }
    int tableSize() const { return m_tableSize; }
    int maxNumNeighbors() const { return m_maxNumNeighbors; }

    // debug helpers of the reference (L63-90), host side
    int ComputeIntCoord(float value) const;
    int HashCoords(int x, int y, int z) const;
    int HashPosition(const float* p3) const;

private:
    float m_spacing;
    int m_tableSize;
    int m_maxNumNeighbors;
};

struct StageTiming {
    std::vector<std::string> labels;
    std::vector<float> ms;
};

class VtClothSolverGPU {
public:
    explicit VtClothSolverGPU(int device = -1, const VtSimParams* params = nullptr);
    ~VtClothSolverGPU();
    VtClothSolverGPU(const VtClothSolverGPU&) = delete;
    VtClothSolverGPU& operator=(const VtClothSolverGPU&) = delete;

    // ---- reference surface
    void Simulate();                 // hpp L56-111, frame time = 1/60
    void Simulate(float frameTime);  // overload named by the spec
    int AddCloth(const float* vertices, int numVertices, const uint* meshIndices, int numIndices,
                 const float* modelMatrix16, float particleDiameter);                   // L114-156
    void AddStretch(int idx1, int idx2, float distance);                               // L158-163
    void AddAttachSlot(const float* attachSlotPos3);                                   // L165-168
    void AddAttach(int particleIndex, int slotIndex, float distance);                  // L170-176
    void AddBend(uint idx1, uint idx2, uint idx3, uint idx4, float angle);             // L178-185
    void UpdateColliders(const VtSDFCollider* colliders, int numColliders);            // L187-205
    void Synchronize();
    void OnDestroy();  // L50-54

    // Double-buffered read-back of positions + normals (headless replacement of VtMergedBuffer::sync, VtBuffer.hpp
    // L180-236): the frame's results are snapshot into one of two device staging buffers on the solver stream (a few
    // microseconds) and travel to pinned host memory on a separate copy stream, so the PCIe transfer of frame k overlaps
    // the simulation of frame k+1.  Returns a ticket for ReadbackWait; at most two read-backs may be outstanding.
    // To be called after the host (or velvet_solver_upload) rewrote a public buffer in place: constraint / topology buffers
    // make the fused pipeline rebuild its tile plan, initialPositions refreshes its float4 copy; state buffers (positions,
    // velocities, predicted, invMasses, attachSlotPositions) are re-imported every frame anyway.
    void NotifyBufferEdited(int bufferId);
    // The fused pipeline's spatial-hash stage (hash -> sort -> cell table -> reordered, tag-filtered neighbour cache) run
    // stand-alone on the public `predicted` buffer and the hash's current initialPositions; results land in the public hash
    // buffers.  Simulate() runs exactly these kernels on its internal float4 state; this entry exists so that they can be
    // checked on arbitrary inputs (tests/test_hash_gpu.py).  Synchronous.
    void HashFused();
    // Renderer hand-off (VtClothSolverGPU.hpp L107-110: positions.sync(); normals.sync()): device arrays owned by the caller
    // -- the mapped pointers of the cloth's GL vertex buffers, or any device allocation of 3 floats per particle of that cloth
    // -- registered per cloth; SyncRenderTargets() mirrors the cloth ranges into them in stream order.  NULL detaches.
    void SetRenderTargets(int clothIndex, float* positionsDev, float* normalsDev);
    void SyncRenderTargets();
    // Debug guard (the reference's unused VtClothSolverCPU::CheckNAN, L407-418): counts non-finite components of positions,
    // velocities and predicted on the device; returns the count and the first offending particle (or numParticles).
    unsigned CheckNaN(unsigned* firstParticle);
    // MouseGrabber.hpp L31-110 as device operations (input_kernels.cuh); Grab synchronises to return the pick
    struct GrabResult {
        int index;
        float distanceToOrigin;
    };
    GrabResult Grab(const float* rayOrigin3, const float* rayDirection3);
    void Drag(const float* rayOrigin3, const float* rayDirection3);
    void Release();
    // must be called before AddCloth: hash arrays host-readable (managed) like the reference's
    void setHashHostReadable(bool on) { m_hashHostReadable = on; }
    int ReadbackPipelined(float* hostPositions, float* hostNormals);
    void ReadbackWait(int ticket);

    // Batched independent cloths (north_star mode 1 / BASELINE config 4): `numInstances` copies of one grid cloth of
    // `resolution`, instance i placed by modelMatrices16[i].  Instances never interact: each has its own rows of the
    // hash table, its own neighbour lists and attach-slot positions; they share ONE constraint set / tile plan, built
    // from instance 0 (rest lengths are those of instance 0's world-space mesh).  Must be the only registration call
    // on the solver; fused pipeline only.  Particle i of instance k is global particle k * (R+1)^2 + i.
    void AddClothInstances(int resolution, const float* vertices, const uint* meshIndices, const float* modelMatrices16,
                           int numInstances, const int* attachedIndices, int numAttached);
    int numInstances() const { return (int)m_instancing.count; }

    // ---- one cloth decomposed over `world` ranks (north_star mode 2).  Every rank registers the same cloth; rank r owns
    // the tile range ddPlan().tileBegin..tileEnd of the (identical) Morton-ordered plan and runs the Jacobi iterations of
    // those tiles only.  The frame is driven step by step; between ddIterateOwned and ddIterateFinish the caller moves
    // ddSendBuf -> peers' ddRecvBuf (per-peer ranges from ddPeerRanges), and between ddGatherPack and ddGatherUnpack it
    // all-gathers ddGatherSend into ddGatherRecv.  Stages other than the iterations are replicated on every rank (for now).
    struct DDBuffers {
        float4 *sendBuf, *recvBuf, *gatherSend, *gatherRecv;
        unsigned sendTotal, recvTotal, ownedCount, maxOwnedCount;
    };
    void ddSetup(int rank, int world);
    const ExchangePlan& ddPlan() const { return m_dd; }
    DDBuffers ddBuffers();      // (sets up the tile form if only the strip form exists so far)
    void ddEnsureTiles();
    // strip form of a single grid cloth (peer transport): true once ddSetup chose it; the tile form may not exist yet
    bool ddIsStrip() const { return m_ddStrip; }
    bool ddTilesReady() const { return m_ddTilesReady; }
    // {first owned tile row, end tile row, tile rows of the cloth, owned particles, largest share, rows exchanged per iteration}
    void ddStripInfo(unsigned out[6]) const;
    void ddFrameBegin(float frameTime);
    void ddSubstepBegin(int substep);  // hash (keys/sort replicated, lists of owned particles) + collide of owned particles + pack
    void ddIterateOwned();
    void ddIterateFinish();
    void ddGatherPack();
    void ddGatherUnpack();
    void ddSubstepEnd(int substep);
    void ddFrameEnd();
    // NVLink peer-memory transport (dd_peer.cuh): export this rank's IPC handles, map the peers', then run whole frames
    // as one CUDA graph with no host or NCCL in the loop.
    static size_t ddPeerBlobBytes();
    void ddPeerExport(void* blob);
    void ddPeerImport(const void* blobs, size_t blobBytes);  // `world` blobs, rank-major (own entry ignored)
    void ddPeerClose();
    void ddSimulate(float frameTime);
    bool ddPeerError();  // true when a wait kernel timed out (peer missing); synchronises the stream

    // VtClothObjectGPU::Start for a grid cloth whose particles AddCloth just registered at `base`: the stretch / attach /
    // bending lists are written by kernels (setup_kernels.cuh), bit-identical to GenerateGridConstraints' and in its order.
    void GenerateGridClothOnDevice(int resolution, int base, const float* vertices, const float* modelMatrix16,
                                   const std::vector<int>& attachedIndices);
    // false when VELVET_HOST_GENERATE is set (the host loops of round 1: kept for A/B tests of the lists)
    static bool deviceRegistration();

    // bulk variants (one memcpy instead of a managed-memory push_back per element)
    void AddStretchBulk(const int* idxPairs, const float* distances, size_t n);
    void AddBendBulk(const uint* idxQuads, const float* angles, size_t n);
    void AddAttachBulk(const int* particleIds, const int* slotIds, const float* distances, size_t n);

    // per-stage timing of one frame, labels as in GUI.cpp L32-51 (runs the frame un-graphed)
    StageTiming SimulateTimed();

    VtSimParams simParams;  // Global::simParams of this solver

    // ---- public sim buffers, names as VtClothSolverGPU.hpp L209-231
    VtMergedBuffer<vec3> positions;
    VtMergedBuffer<vec3> normals;
    VtBuffer<uint> indices;
    VtBuffer<vec3> velocities;
    VtBuffer<vec3> predicted;
    VtBuffer<vec3> deltas;
    VtBuffer<int> deltaCounts;
    VtBuffer<float> invMasses;
    VtBuffer<int> stretchIndices;
    VtBuffer<float> stretchLengths;
    VtBuffer<uint> bendIndices;
    VtBuffer<float> bendAngles;
    VtBuffer<int> attachParticleIDs;
    VtBuffer<int> attachSlotIDs;
    VtBuffer<float> attachDistances;
    VtBuffer<vec3> attachSlotPositions;
    VtBuffer<VtSDFCollider> sdfColliders;

    std::shared_ptr<SpatialHashGPU> spatialHash() const { return m_spatialHash; }

    // ---- new controls
    void setPipeline(int pipeline);
    int pipeline() const { return m_pipeline; }
    void setTileSize(int particlesPerTile);
    void setMathMode(int mode);  // VELVET_MATH_EXACT (default) or VELVET_MATH_FAST
    // Jacobi kernel selection: VELVET_ITERATE_AUTO (default) runs the implicit-grid kernel when every registered cloth is a
    // grid with the reference's constraint pattern (grid_plan.hpp) and the record-driven tile kernel otherwise;
    // VELVET_ITERATE_TILES always runs the tile kernel.  Both give bit-identical results.
    void setIterateMode(int mode);
    int iterateKernel();  // the kernel the next frame will run: VELVET_ITERATE_TILES or VELVET_ITERATE_GRID (builds the plans if needed)
    const GridPlan& gridPlan() const { return m_gridPlan; }
    int mathMode() const { return m_mathMode; }
    cudaStream_t stream() const { return m_stream; }
    int device() const { return m_device; }
    int lastLaunchCount() const { return m_lastLaunches; }
    const TilePlan& tilePlan()  // builds the record-driven plan if the solver has been running the grid kernel only
    {
        ensureTilePlan();
        return m_plan;
    }
    const std::string& fusedFallbackReason() const { return m_fallbackReason; }

private:
    struct Stage;  // timing helper
    void simulateSeam(float frameTime, Stage* timing);
    void recordFusedFrame(Stage* timing);
    void ensureFusedResources();
    bool buildTilePlan();
    void ensureTilePlan();
    bool m_tilePlanBuilt = false;
    void invalidate() { m_topologyDirty = true; }
    void clampNeighborBound(VtSimParams& P) const;
    void quiesce();  // drains an asynchronous frame before a registration call touches managed buffers
    bool m_mayBeBusy = false;
    unsigned long long topologyKey() const;

    int m_device = 0;
    cudaStream_t m_stream = nullptr;
    cudaStream_t m_copyStream = nullptr;
    cudaEvent_t m_staged[2] = {nullptr, nullptr}, m_copyDone[2] = {nullptr, nullptr};
    bool m_copyPending[2] = {false, false};
    unsigned m_readbackSeq = 0;
    DeviceBuffer<float> m_stagePos[2], m_stageNrm[2];
    int m_pipeline = 0;
    int m_tileSize = 0;
    int m_mathMode = VELVET_MATH_EXACT;
    int m_iterateMode = VELVET_ITERATE_AUTO;
    bool m_hashHostReadable = false;
    DeviceBuffer<unsigned> m_nanScratch;
    std::vector<ClothRange> m_clothRanges;  // particle range of every AddCloth call
    // what GenerateGridClothOnDevice appended for each cloth: while the lists consist of exactly these ranges, the grid plan
    // follows from the generator (no entry-by-entry check on the host, which would fault the lists back out of the device)
    struct GeneratedCloth {
        uint base;
        int R;
        size_t stretchBegin, bendBegin, attachBegin;
        uint numSlots, firstSlot;
    };
    std::vector<GeneratedCloth> m_generated;
    bool generatedListsIntact() const;
    unsigned walkBandParticles() const;
    unsigned gridIterationsPerLaunch() const;
    DeviceBuffer<unsigned> m_gridBarrier;  // arrival counter of the grid-wide barriers of a multi-iteration launch
    bool buildGridPlanOnDevice(uint planN, cudaStream_t st);
    DeviceBuffer<input::GrabState> m_grab;
    DeviceBuffer<int> m_setupFlags;  // [0] mesh index out of range, [1] bending quads differ from the grid pattern
    Instancing m_instancing{1, 0, 0};
    // domain decomposition state
    ExchangePlan m_dd;
    bool m_ddReady = false;
    std::vector<unsigned> m_ddSendOff, m_ddRecvOff, m_ddOwnedBegin, m_ddOwnedCount;  // per peer / per rank
    unsigned m_ddMaxOwned = 0;
    DeviceBuffer<uint> m_ddSendIds, m_ddRecvIds;
    DeviceBuffer<unsigned char> m_ddOwnedMask;
    DeviceBuffer<float4> m_ddSendBuf, m_ddRecvBuf, m_ddGatherSend, m_ddGatherRecv;
    float4 *m_ddCur = nullptr, *m_ddOther = nullptr;
    void recordDDFrame();
    void recordDDStripFrame(Stage* timing);
    void ddSetupTiles();
    bool m_ddTilesReady = false;
    bool m_ddStrip = false;                 // single grid cloth: strips of tile rows, exchange fused into iterate_grid_kernel
    std::vector<unsigned> m_ddTileRow;      // [world + 1] tile rows of every rank
    DeviceBuffer<unsigned char> m_ddStripMask;
    void ddValidateFrame() const;
    DeviceBuffer<unsigned> m_ddFlags;
    DeviceBuffer<ddpeer::Control> m_ddCtl;
    DeviceBuffer<unsigned char> m_ddSendPeer;
    ddpeer::PeerTable m_ddPeers{};
    std::vector<void*> m_ddOpened;
    bool m_ddPeersReady = false;
    unsigned long long m_ddGraphKey = 0;
    cudaGraph_t m_ddGraph = nullptr;
    cudaGraphExec_t m_ddGraphExec = nullptr;
    int m_ddGraphLaunches = 0;
    bool m_instanced = false;
    int m_lastLaunches = 0;
    std::shared_ptr<SpatialHashGPU> m_spatialHash;

    // fused pipeline state
    bool m_topologyDirty = true;
    bool m_initDirty = false;  // initialPositions edited: m_init4 must be re-packed
    bool m_fusedUsable = false;
    std::string m_fallbackReason;
    unsigned long long m_graphKey = 0;
    cudaGraph_t m_graph = nullptr;
    cudaGraphExec_t m_graphExec = nullptr;
    int m_graphLaunches = 0;

    DeviceBuffer<float4> m_pos4, m_predA, m_predB, m_init4, m_sorted;
    DeviceBuffer<uint> m_keysAlt, m_valsAlt;
    DeviceBuffer<VtSDFCollider> m_collidersDev;  // stream-ordered device copy of the UpdateColliders block
    DeviceBuffer<PreparedCollider> m_prepared;
    DeviceBuffer<FrameParams> m_frameParams;
    DeviceBuffer<uint> m_vtxTriOff, m_vtxTris;
    DeviceBuffer<float> m_slotsDev;  // per-frame device copy of attachSlotPositions
    RadixSorter m_sorter;
    TilePlan m_plan;
    TilePlanDev m_planDev{};
    DeviceBuffer<TileDesc> m_dTiles;
    DeviceBuffer<uint> m_dOwned, m_dHalo, m_dAttOff;
    DeviceBuffer<uint16_t> m_dCnt16;
    DeviceBuffer<uint2> m_dStretchRec, m_dAttachRec;
    DeviceBuffer<uint4> m_dBendRec;
    // implicit-grid plan (valid when every cloth is a grid with the reference's constraint pattern)
    GridPlan m_gridPlan;
    GridPlanDev m_gridDev{};
    bool m_gridUsable = false;
    bool m_squareTilesOnly = false;  // set by the strip decomposition, which counts the cloth in 15-row tiles
    DeviceBuffer<GridCloth> m_gCloths;
    DeviceBuffer<float4> m_gRest4;
    DeviceBuffer<float> m_gAngle;
    DeviceBuffer<uint> m_gAttOff;
    DeviceBuffer<uint2> m_gAttachRec;
};

// Constraint generation of the reference's cloth component (VtClothObjectGPU.hpp L43-148) for grid meshes
// produced by GenerateClothMesh (Scene.hpp L131-168).
class VtClothObjectGPU {
public:
    VtClothObjectGPU(int resolution, VtClothSolverGPU* solver) : m_resolution(resolution), m_solver(solver) {}
    void SetAttachedIndices(std::vector<int> indices) { m_attachedIndices = std::move(indices); }
    float particleDiameter() const { return m_particleDiameter; }
    int indexOffset() const { return m_indexOffset; }
    // Start(): vertices in model space (host), mesh indices (host), model matrix
    void Start(const float* vertices, const uint* meshIndices, const float* modelMatrix16);

private:
    int m_resolution;
    int m_indexOffset = 0;
    VtClothSolverGPU* m_solver;
    std::vector<int> m_attachedIndices;
    float m_particleDiameter = 0;
};

// The constraint lists VtClothObjectGPU::Start generates for a grid cloth (VtClothObjectGPU.hpp L75-148), as arrays.
struct GridConstraints {
    float particleDiameter = 0;
    std::vector<int> stretchIdx;
    std::vector<float> stretchLen;
    std::vector<uint> bendIdx;
    std::vector<float> bendAngle;
    std::vector<float> slotPositions;  // 3 per attach slot
    std::vector<int> attachPid, attachSlot;
    std::vector<float> attachDist;
};
GridConstraints GenerateGridConstraints(int resolution, const float* vertices, const uint* meshIndices, const float* modelMatrix16,
                                        const std::vector<int>& attachedIndices, float particleDiameterScalar, int indexOffset);

// Scene.hpp L131-168, Transform.hpp L22-29 (+ Helper.cpp L8-15), glm::inverse, VtClothSolverGPU.hpp L195-203
void GenerateClothMesh(int resolution, float* vertices, uint* meshIndices);
void TransformMatrix(const float* position3, const float* rotationDeg3, const float* scale3, float* out16);
void Mat4Inverse(const float* m16, float* out16);
void MakeCollider(int type, const float* position3, const float* scale3, const float* cur16, const float* last16,
                  float deltaTime, VtSDFCollider* out);

}  // namespace velvet
