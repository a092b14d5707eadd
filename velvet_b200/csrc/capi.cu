// capi.cu -- extern "C" object surface of include/velvet_b200.h over velvet::VtClothSolverGPU / SpatialHashGPU.
#include <cstring>

#include "capi_util.hpp"
#include "solver.hpp"

namespace velvet {

namespace {
thread_local std::string t_lastError;
}

int set_error(int status, const std::string& msg)
{
    t_lastError = msg;
    return status;
}
void clear_error() { t_lastError.clear(); }

}  // namespace velvet

using namespace velvet;

struct VelvetSolver {
    VtClothSolverGPU impl;
    VelvetSolver(int device, const VtSimParams* p) : impl(device, p) {}
};

struct VelvetSpatialHash {
    SpatialHashGPU impl;
    float particleDiameter;
    cudaStream_t stream = 0;
    VelvetSpatialHash(float d, int n, float scalar, int k) : impl(d, n, scalar, k), particleDiameter(d) {}
};

#define VT_REQUIRE(cond, msg) \
    if (!(cond)) return set_error(VELVET_ERR_INVALID_ARGUMENT, msg)

namespace {

struct BufView {
    void* ptr;
    size_t count;     // elements
    size_t elemSize;  // bytes per element
};

bool solver_buffer(VtClothSolverGPU& s, int id, BufView& v)
{
    SpatialHashGPU* h = s.spatialHash().get();
    switch (id) {
    case VELVET_BUF_POSITIONS: v = {s.positions.data(), s.positions.size(), 12}; return true;
    case VELVET_BUF_NORMALS: v = {s.normals.data(), s.normals.size(), 12}; return true;
    case VELVET_BUF_INDICES: v = {s.indices.data(), s.indices.size(), 4}; return true;
    case VELVET_BUF_VELOCITIES: v = {s.velocities.data(), s.velocities.size(), 12}; return true;
    case VELVET_BUF_PREDICTED: v = {s.predicted.data(), s.predicted.size(), 12}; return true;
    case VELVET_BUF_DELTAS: v = {s.deltas.data(), s.deltas.size(), 12}; return true;
    case VELVET_BUF_DELTACOUNTS: v = {s.deltaCounts.data(), s.deltaCounts.size(), 4}; return true;
    case VELVET_BUF_INVMASSES: v = {s.invMasses.data(), s.invMasses.size(), 4}; return true;
    case VELVET_BUF_STRETCHINDICES: v = {s.stretchIndices.data(), s.stretchIndices.size(), 4}; return true;
    case VELVET_BUF_STRETCHLENGTHS: v = {s.stretchLengths.data(), s.stretchLengths.size(), 4}; return true;
    case VELVET_BUF_BENDINDICES: v = {s.bendIndices.data(), s.bendIndices.size(), 4}; return true;
    case VELVET_BUF_BENDANGLES: v = {s.bendAngles.data(), s.bendAngles.size(), 4}; return true;
    case VELVET_BUF_ATTACHPARTICLEIDS: v = {s.attachParticleIDs.data(), s.attachParticleIDs.size(), 4}; return true;
    case VELVET_BUF_ATTACHSLOTIDS: v = {s.attachSlotIDs.data(), s.attachSlotIDs.size(), 4}; return true;
    case VELVET_BUF_ATTACHDISTANCES: v = {s.attachDistances.data(), s.attachDistances.size(), 4}; return true;
    case VELVET_BUF_ATTACHSLOTPOSITIONS: v = {s.attachSlotPositions.data(), s.attachSlotPositions.size(), 12}; return true;
    case VELVET_BUF_SDFCOLLIDERS: v = {s.sdfColliders.data(), s.sdfColliders.size(), sizeof(VtSDFCollider)}; return true;
    default: break;
    }
    if (!h) {
        v = {nullptr, 0, 4};
        return id >= VELVET_BUF_NEIGHBORS && id <= VELVET_BUF_CELLEND;
    }
    switch (id) {
    case VELVET_BUF_NEIGHBORS: v = {h->neighbors.data(), h->neighbors.size(), 4}; return true;
    case VELVET_BUF_INITIALPOSITIONS: v = {h->initialPositions.data(), h->initialPositions.size(), 12}; return true;
    case VELVET_BUF_PARTICLEHASH: v = {h->particleHash.data(), h->particleHash.size(), 4}; return true;
    case VELVET_BUF_PARTICLEINDEX: v = {h->particleIndex.data(), h->particleIndex.size(), 4}; return true;
    case VELVET_BUF_CELLSTART: v = {h->cellStart.data(), h->cellStart.size(), 4}; return true;
    case VELVET_BUF_CELLEND: v = {h->cellEnd.data(), h->cellEnd.size(), 4}; return true;
    default: return false;
    }
}

}  // namespace

extern "C" {

const char* velvet_last_error(void) { return t_lastError.c_str(); }
int velvet_version(void) { return 100; }

int velvet_default_params(VtSimParams* p)
{
    VT_REQUIRE(p, "params is NULL");
    default_sim_params(*p);
    return VELVET_OK;
}

int velvet_solver_create(VelvetSolver** out, int device, const VtSimParams* params)
{
    VT_API_BEGIN
    VT_REQUIRE(out, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_error(VELVET_ERR_CUDA, std::string("no CUDA device available: ") + cudaGetErrorString(e));
    *out = new VelvetSolver(device, params);
    VT_API_END
}

int velvet_solver_destroy(VelvetSolver* s)
{
    VT_API_BEGIN
    delete s;
    VT_API_END
}

VtSimParams* velvet_solver_params(VelvetSolver* s) { return s ? &s->impl.simParams : nullptr; }

int velvet_solver_set_pipeline(VelvetSolver* s, int pipeline)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.setPipeline(pipeline);
    VT_API_END
}

int velvet_solver_set_math_mode(VelvetSolver* s, int mode)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.setMathMode(mode);
    VT_API_END
}

int velvet_solver_set_iterate_mode(VelvetSolver* s, int mode)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.setIterateMode(mode);
    VT_API_END
}

int velvet_solver_iterate_kernel(VelvetSolver* s, int* kernel)
{
    VT_API_BEGIN
    VT_REQUIRE(s && kernel, "NULL argument");
    *kernel = s->impl.iterateKernel();
    VT_API_END
}

int velvet_solver_set_tile_size(VelvetSolver* s, int n)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.setTileSize(n);
    VT_API_END
}

int velvet_solver_add_cloth(VelvetSolver* s, const float* vertices, int numVertices, const unsigned* indices, int numIndices,
                            const float* modelMatrix16, float particleDiameter, int* offset)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    const int off = s->impl.AddCloth(vertices, numVertices, indices, numIndices, modelMatrix16, particleDiameter);
    if (offset) *offset = off;
    VT_API_END
}

int velvet_solver_add_stretch(VelvetSolver* s, int idx1, int idx2, float distance)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.AddStretch(idx1, idx2, distance);
    VT_API_END
}

int velvet_solver_add_attach_slot(VelvetSolver* s, const float* slotPos3)
{
    VT_API_BEGIN
    VT_REQUIRE(s && slotPos3, "bad argument");
    s->impl.AddAttachSlot(slotPos3);
    VT_API_END
}

int velvet_solver_add_attach(VelvetSolver* s, int particleIndex, int slotIndex, float distance)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.AddAttach(particleIndex, slotIndex, distance);
    VT_API_END
}

int velvet_solver_add_bend(VelvetSolver* s, unsigned idx1, unsigned idx2, unsigned idx3, unsigned idx4, float angle)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.AddBend(idx1, idx2, idx3, idx4, angle);
    VT_API_END
}

int velvet_solver_update_colliders(VelvetSolver* s, const VtSDFCollider* colliders, int numColliders)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.UpdateColliders(colliders, numColliders);
    VT_API_END
}

int velvet_make_collider(int type, const float* position3, const float* scale3, const float* cur16, const float* last16,
                         float deltaTime, VtSDFCollider* out)
{
    VT_REQUIRE(position3 && scale3 && cur16 && last16 && out, "make_collider: NULL argument");
    MakeCollider(type, position3, scale3, cur16, last16, deltaTime, out);
    return VELVET_OK;
}

int velvet_solver_simulate(VelvetSolver* s, int sync)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.Simulate();
    if (sync) s->impl.Synchronize();
    VT_API_END
}

int velvet_solver_simulate_dt(VelvetSolver* s, float frameTime, int sync)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    VT_REQUIRE(frameTime > 0, "frameTime must be positive");
    s->impl.Simulate(frameTime);
    if (sync) s->impl.Synchronize();
    VT_API_END
}

int velvet_solver_synchronize(VelvetSolver* s)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.Synchronize();
    VT_API_END
}

int velvet_solver_hash(VelvetSolver* s)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    VT_REQUIRE(s->impl.spatialHash(), "no cloth registered");
    s->impl.spatialHash()->Hash(reinterpret_cast<const float*>(s->impl.predicted.data()), s->impl.predicted.size(),
                                s->impl.simParams.particleDiameter, s->impl.stream());
    VT_API_END
}

int velvet_solver_hash_fused(VelvetSolver* s)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    VT_REQUIRE(s->impl.spatialHash(), "no cloth registered");
    s->impl.HashFused();
    VT_API_END
}

int velvet_solver_buffer(VelvetSolver* s, int bufferId, void** devPtr, size_t* count)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    BufView v;
    if (!solver_buffer(s->impl, bufferId, v)) return set_error(VELVET_ERR_INVALID_ARGUMENT, "unknown buffer id");
    if (devPtr) *devPtr = v.ptr;
    if (count) *count = v.count;
    VT_API_END
}

int velvet_solver_download(VelvetSolver* s, int bufferId, void* host, size_t bytes)
{
    VT_API_BEGIN
    VT_REQUIRE(s && host, "bad argument");
    BufView v;
    if (!solver_buffer(s->impl, bufferId, v)) return set_error(VELVET_ERR_INVALID_ARGUMENT, "unknown buffer id");
    VT_REQUIRE(bytes <= v.count * v.elemSize, "download: more bytes than the buffer holds");
    s->impl.Synchronize();
    if (bytes) VT_CUDA(cudaMemcpy(host, v.ptr, bytes, cudaMemcpyDefault));
    VT_API_END
}

int velvet_solver_upload(VelvetSolver* s, int bufferId, const void* host, size_t bytes)
{
    VT_API_BEGIN
    VT_REQUIRE(s && host, "bad argument");
    BufView v;
    if (!solver_buffer(s->impl, bufferId, v)) return set_error(VELVET_ERR_INVALID_ARGUMENT, "unknown buffer id");
    VT_REQUIRE(bytes <= v.count * v.elemSize, "upload: more bytes than the buffer holds");
    s->impl.Synchronize();
    if (bytes) VT_CUDA(cudaMemcpy(v.ptr, host, bytes, cudaMemcpyDefault));
    s->impl.NotifyBufferEdited(bufferId);
    VT_API_END
}

int velvet_solver_set_render_targets(VelvetSolver* s, int clothIndex, float* positionsDev, float* normalsDev)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.SetRenderTargets(clothIndex, positionsDev, normalsDev);
    VT_API_END
}

int velvet_solver_sync_render_targets(VelvetSolver* s)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.SyncRenderTargets();
    VT_API_END
}

int velvet_solver_set_hash_host_readable(VelvetSolver* s, int on)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    VT_REQUIRE(s->impl.simParams.numParticles == 0, "set_hash_host_readable must precede AddCloth");
    s->impl.setHashHostReadable(on != 0);
    VT_API_END
}

int velvet_solver_check_nan(VelvetSolver* s, unsigned* nonFiniteCount, unsigned* firstParticle)
{
    VT_API_BEGIN
    VT_REQUIRE(s && nonFiniteCount, "bad argument");
    *nonFiniteCount = s->impl.CheckNaN(firstParticle);
    VT_API_END
}

int velvet_solver_grab(VelvetSolver* s, const float* rayOrigin3, const float* rayDirection3, int* grabbedIndex, float* distanceToOrigin)
{
    VT_API_BEGIN
    VT_REQUIRE(s && rayOrigin3 && rayDirection3, "grab: bad argument");
    const VtClothSolverGPU::GrabResult r = s->impl.Grab(rayOrigin3, rayDirection3);
    if (grabbedIndex) *grabbedIndex = r.index;
    if (distanceToOrigin) *distanceToOrigin = r.distanceToOrigin;
    VT_API_END
}

int velvet_solver_drag(VelvetSolver* s, const float* rayOrigin3, const float* rayDirection3)
{
    VT_API_BEGIN
    VT_REQUIRE(s && rayOrigin3 && rayDirection3, "drag: bad argument");
    s->impl.Drag(rayOrigin3, rayDirection3);
    VT_API_END
}

int velvet_solver_release(VelvetSolver* s)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.Release();
    VT_API_END
}

int velvet_solver_readback_async(VelvetSolver* s, float* hostPositions, float* hostNormals)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    const size_t bytes = s->impl.positions.size() * 12;
    if (hostPositions && bytes)
        VT_CUDA(cudaMemcpyAsync(hostPositions, s->impl.positions.data(), bytes, cudaMemcpyDeviceToHost, s->impl.stream()));
    if (hostNormals && bytes)
        VT_CUDA(cudaMemcpyAsync(hostNormals, s->impl.normals.data(), bytes, cudaMemcpyDeviceToHost, s->impl.stream()));
    VT_API_END
}

int velvet_solver_readback_pipelined(VelvetSolver* s, float* hostPositions, float* hostNormals, int* ticket)
{
    VT_API_BEGIN
    VT_REQUIRE(s && ticket, "readback_pipelined: bad argument");
    *ticket = s->impl.ReadbackPipelined(hostPositions, hostNormals);
    VT_API_END
}

int velvet_solver_readback_wait(VelvetSolver* s, int ticket)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.ReadbackWait(ticket);
    VT_API_END
}

void* velvet_solver_stream(VelvetSolver* s) { return s ? (void*)s->impl.stream() : nullptr; }
int velvet_solver_last_launch_count(VelvetSolver* s) { return s ? s->impl.lastLaunchCount() : 0; }

int velvet_solver_simulate_timed(VelvetSolver* s, const char** labels, float* ms, int cap)
{
    static thread_local StageTiming keep;  // owns the label strings handed back to the caller
    try {
        clear_error();
        if (!s) return set_error(VELVET_ERR_INVALID_ARGUMENT, "solver is NULL");
        keep = s->impl.SimulateTimed();
        int n = (int)keep.labels.size();
        if (n > cap) n = cap;
        for (int i = 0; i < n; i++) {
            if (labels) labels[i] = keep.labels[i].c_str();
            if (ms) ms[i] = keep.ms[i];
        }
        return n;
    } catch (const Error& e) {
        return set_error(e.status, e.what());
    } catch (const std::exception& e) {
        return set_error(VELVET_ERR_STATE, e.what());
    }
}

int velvet_generate_cloth_mesh(int resolution, float* vertices, unsigned* indices)
{
    VT_REQUIRE(resolution > 0 && vertices && indices, "generate_cloth_mesh: bad argument");
    GenerateClothMesh(resolution, vertices, indices);
    return VELVET_OK;
}

int velvet_transform_matrix(const float* position3, const float* rotationDeg3, const float* scale3, float* out16)
{
    VT_REQUIRE(position3 && rotationDeg3 && scale3 && out16, "transform_matrix: NULL argument");
    TransformMatrix(position3, rotationDeg3, scale3, out16);
    return VELVET_OK;
}

int velvet_cloth_object_start(VelvetSolver* s, int resolution, const float* vertices, const unsigned* indices,
                              const float* modelMatrix16, const int* attachedIndices, int numAttached, int* offset)
{
    VT_API_BEGIN
    VT_REQUIRE(s && resolution > 0 && vertices && indices && modelMatrix16 && numAttached >= 0, "cloth_object_start: bad argument");
    VT_REQUIRE(numAttached == 0 || attachedIndices, "cloth_object_start: attachedIndices is NULL");
    const int nv = (resolution + 1) * (resolution + 1);
    for (int i = 0; i < numAttached; i++) VT_REQUIRE(attachedIndices[i] >= 0 && attachedIndices[i] < nv, "attached index out of range");
    VtClothObjectGPU obj(resolution, &s->impl);
    obj.SetAttachedIndices(std::vector<int>(attachedIndices, attachedIndices + numAttached));
    obj.Start(vertices, indices, modelMatrix16);
    if (offset) *offset = obj.indexOffset();
    VT_API_END
}

int velvet_solver_add_cloth_instances(VelvetSolver* s, int resolution, const float* vertices, const unsigned* indices,
                                      const float* modelMatrices16, int numInstances, const int* attachedIndices, int numAttached)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.AddClothInstances(resolution, vertices, indices, modelMatrices16, numInstances, attachedIndices, numAttached);
    VT_API_END
}

int velvet_solver_dd_setup(VelvetSolver* s, int rank, int world)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.ddSetup(rank, world);
    VT_API_END
}

int velvet_solver_dd_info(VelvetSolver* s, VelvetDDInfo* out)
{
    VT_API_BEGIN
    VT_REQUIRE(s && out, "bad argument");
    if (s->impl.ddIsStrip() && !s->impl.ddTilesReady()) {  // strip form only: no staging buffers, rows travel by peer stores
        unsigned si[6];
        s->impl.ddStripInfo(si);
        const ExchangePlan& x = s->impl.ddPlan();
        std::memset(out, 0, sizeof(*out));
        out->rank = x.rank;
        out->world = x.world;
        out->tileBegin = si[0];
        out->tileEnd = si[1];
        out->numTiles = si[2];
        out->ownedCount = si[3];
        out->maxOwnedCount = si[4];
        out->sendTotal = si[5];
        out->recvTotal = si[5];
        return VELVET_OK;
    }
    const ExchangePlan& x = s->impl.ddPlan();
    const VtClothSolverGPU::DDBuffers b = s->impl.ddBuffers();
    out->rank = x.rank;
    out->world = x.world;
    out->tileBegin = x.tileBegin;
    out->tileEnd = x.tileEnd;
    out->numTiles = (unsigned)s->impl.tilePlan().tiles.size();
    out->ownedCount = b.ownedCount;
    out->maxOwnedCount = b.maxOwnedCount;
    out->sendTotal = b.sendTotal;
    out->recvTotal = b.recvTotal;
    out->sendBuf = b.sendBuf;
    out->recvBuf = b.recvBuf;
    out->gatherSend = b.gatherSend;
    out->gatherRecv = b.gatherRecv;
    VT_API_END
}

int velvet_solver_dd_offsets(VelvetSolver* s, unsigned* sendOffsets, unsigned* recvOffsets)
{
    VT_API_BEGIN
    VT_REQUIRE(s && sendOffsets && recvOffsets, "bad argument");
    s->impl.ddBuffers();  // throws unless set up; brings the tile form (and its exchange lists) into being
    const ExchangePlan& x = s->impl.ddPlan();
    unsigned so = 0, ro = 0;
    for (int q = 0; q < x.world; q++) {
        sendOffsets[q] = so;
        recvOffsets[q] = ro;
        so += (unsigned)x.sendIds[q].size();
        ro += (unsigned)x.recvIds[q].size();
    }
    sendOffsets[x.world] = so;
    recvOffsets[x.world] = ro;
    VT_API_END
}

int velvet_solver_dd_prepare_stepped(VelvetSolver* s)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.ddEnsureTiles();
    VT_API_END
}

int velvet_solver_dd_step(VelvetSolver* s, int op, int arg, float farg)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    switch (op) {
    case VELVET_DD_FRAME_BEGIN: s->impl.ddFrameBegin(farg > 0 ? farg : kFixedDeltaTime); break;
    case VELVET_DD_SUBSTEP_BEGIN: s->impl.ddSubstepBegin(arg); break;
    case VELVET_DD_ITERATE_OWNED: s->impl.ddIterateOwned(); break;
    case VELVET_DD_ITERATE_FINISH: s->impl.ddIterateFinish(); break;
    case VELVET_DD_GATHER_PACK: s->impl.ddGatherPack(); break;
    case VELVET_DD_GATHER_UNPACK: s->impl.ddGatherUnpack(); break;
    case VELVET_DD_SUBSTEP_END: s->impl.ddSubstepEnd(arg); break;
    case VELVET_DD_FRAME_END: s->impl.ddFrameEnd(); break;
    default: return set_error(VELVET_ERR_INVALID_ARGUMENT, "unknown decomposition step");
    }
    VT_API_END
}

size_t velvet_dd_peer_blob_bytes(void) { return VtClothSolverGPU::ddPeerBlobBytes(); }

int velvet_solver_dd_peer_export(VelvetSolver* s, void* blob)
{
    VT_API_BEGIN
    VT_REQUIRE(s && blob, "dd_peer_export: NULL argument");
    s->impl.ddPeerExport(blob);
    VT_API_END
}

int velvet_solver_dd_peer_import(VelvetSolver* s, const void* blobs, size_t blobBytes)
{
    VT_API_BEGIN
    VT_REQUIRE(s && blobs, "dd_peer_import: NULL argument");
    s->impl.ddPeerImport(blobs, blobBytes);
    VT_API_END
}

int velvet_solver_dd_peer_close(VelvetSolver* s)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.ddPeerClose();
    VT_API_END
}

int velvet_solver_dd_simulate(VelvetSolver* s, float deltaTime, int sync)
{
    VT_API_BEGIN
    VT_REQUIRE(s, "solver is NULL");
    s->impl.ddSimulate(deltaTime > 0 ? deltaTime : kFixedDeltaTime);
    if (sync && s->impl.ddPeerError())
        return set_error(VELVET_ERR_STATE, "dd_simulate: a peer did not arrive within the time-out (exchange flags never advanced)");
    VT_API_END
}

int velvet_dd_plan_grid(int resolution, int tileSize, int rank, int world, unsigned* counts, unsigned* sendIds, unsigned* recvIds,
                        unsigned* ownedRange2)
{
    VT_API_BEGIN
    VT_REQUIRE(resolution > 0 && world > 0 && rank >= 0 && rank < world && counts, "dd_plan_grid: bad argument");
    const int R = resolution;
    const size_t n = (size_t)(R + 1) * (R + 1);
    std::vector<float> v(3 * n);
    std::vector<unsigned> idx((size_t)6 * R * R);
    GenerateClothMesh(R, v.data(), idx.data());
    const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const GridConstraints g = GenerateGridConstraints(R, v.data(), idx.data(), identity, {}, 1.5f, 0);
    const TilePlan plan = build_tile_plan((unsigned)n, v.data(), g.stretchIdx.data(), g.stretchLen.data(), g.stretchLen.size(),
                                          g.bendIdx.data(), g.bendAngle.data(), g.bendAngle.size(), nullptr, nullptr, nullptr, 0,
                                          tileSize ? tileSize : 256);
    if (!plan.valid) return set_error(VELVET_ERR_UNSUPPORTED, plan.whyInvalid);
    VT_REQUIRE((size_t)world <= plan.tiles.size(), "dd_plan_grid: more ranks than tiles");
    const ExchangePlan x = build_exchange_plan(plan, (unsigned)n, rank, world);
    size_t so = 0, ro = 0;
    for (int q = 0; q < world; q++) {
        counts[q] = (unsigned)x.sendIds[q].size();
        counts[world + q] = (unsigned)x.recvIds[q].size();
        if (sendIds) std::memcpy(sendIds + so, x.sendIds[q].data(), 4 * x.sendIds[q].size());
        if (recvIds) std::memcpy(recvIds + ro, x.recvIds[q].data(), 4 * x.recvIds[q].size());
        so += x.sendIds[q].size();
        ro += x.recvIds[q].size();
    }
    if (ownedRange2) {
        ownedRange2[0] = x.tileBegin * (unsigned)plan.tileSize;
        ownedRange2[1] = (unsigned)std::min<size_t>((size_t)x.tileEnd * plan.tileSize, n);
    }
    VT_API_END
}

int velvet_plan_grid_tiles(int resolution, int tileSize, unsigned* numTiles, unsigned* perTile4, unsigned capacityTiles, unsigned* globals4)
{
    VT_API_BEGIN
    VT_REQUIRE(resolution > 0 && numTiles, "plan_grid_tiles: bad argument");
    const int R = resolution;
    const size_t n = (size_t)(R + 1) * (R + 1);
    std::vector<float> v(3 * n);
    std::vector<unsigned> idx((size_t)6 * R * R);
    GenerateClothMesh(R, v.data(), idx.data());
    const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const GridConstraints g = GenerateGridConstraints(R, v.data(), idx.data(), identity, {}, 1.5f, 0);
    const TilePlan plan = build_tile_plan((unsigned)n, v.data(), g.stretchIdx.data(), g.stretchLen.data(), g.stretchLen.size(),
                                          g.bendIdx.data(), g.bendAngle.data(), g.bendAngle.size(), nullptr, nullptr, nullptr, 0,
                                          tileSize ? tileSize : 256);
    if (!plan.valid) return set_error(VELVET_ERR_UNSUPPORTED, plan.whyInvalid);
    *numTiles = (unsigned)plan.tiles.size();
    if (perTile4)
        for (size_t t = 0; t < plan.tiles.size() && t < capacityTiles; t++) {
            perTile4[4 * t + 0] = plan.tiles[t].nOwned;
            perTile4[4 * t + 1] = plan.tiles[t].nHalo;
            perTile4[4 * t + 2] = plan.tiles[t].nStretch;
            perTile4[4 * t + 3] = plan.tiles[t].nBend;
        }
    if (globals4) {
        globals4[0] = plan.maxLocals;
        globals4[1] = plan.maxKS;
        globals4[2] = plan.maxKB;
        globals4[3] = plan.maxBendPerTile;
    }
    VT_API_END
}

int velvet_grid_plan_check(const unsigned* clothCounts, int numCloths, const int* stretchIndices, const float* stretchLengths,
                           size_t numStretch, const unsigned* bendIndices, const float* bendAngles, size_t numBend, int* recognised,
                           unsigned* numTiles, float* rest4, char* whyNot)
{
    VT_API_BEGIN
    VT_REQUIRE(clothCounts && numCloths > 0 && recognised, "grid_plan_check: bad argument");
    std::vector<ClothRange> ranges;
    unsigned total = 0;
    for (int c = 0; c < numCloths; c++) {
        ranges.push_back(ClothRange{total, clothCounts[c]});
        total += clothCounts[c];
    }
    const GridPlan g = build_grid_plan(total, ranges, stretchIndices, stretchLengths, numStretch, bendIndices, bendAngles, numBend,
                                       nullptr, nullptr, nullptr, 0);
    *recognised = g.valid ? 1 : 0;
    if (numTiles) *numTiles = g.valid ? g.numTiles : 0u;
    if (g.valid && rest4) std::memcpy(rest4, g.rest4.data(), sizeof(float) * g.rest4.size());
    if (whyNot) {
        std::memset(whyNot, 0, 128);
        std::strncpy(whyNot, g.why.c_str(), 127);
    }
    VT_API_END
}

int velvet_plan_grid_digest(int resolution, int tileSize, int withAttach, unsigned long long* digest)
{
    VT_API_BEGIN
    VT_REQUIRE(resolution > 0 && digest, "plan_grid_digest: bad argument");
    const int R = resolution;
    const size_t n = (size_t)(R + 1) * (R + 1);
    std::vector<float> v(3 * n);
    std::vector<unsigned> idx((size_t)6 * R * R);
    GenerateClothMesh(R, v.data(), idx.data());
    const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::vector<int> attached;
    if (withAttach) attached = {0, R};
    const GridConstraints g = GenerateGridConstraints(R, v.data(), idx.data(), identity, attached, 1.5f, 0);
    const TilePlan plan = build_tile_plan((unsigned)n, v.data(), g.stretchIdx.data(), g.stretchLen.data(), g.stretchLen.size(),
                                          g.bendIdx.data(), g.bendAngle.data(), g.bendAngle.size(), g.attachPid.data(),
                                          g.attachSlot.data(), g.attachDist.data(), g.attachDist.size(), tileSize ? tileSize : 256);
    if (!plan.valid) return set_error(VELVET_ERR_UNSUPPORTED, plan.whyInvalid);
    unsigned long long h = 1469598103934665603ull;  // FNV-1a over every array of the plan
    auto mix = [&](const void* p, size_t bytes) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < bytes; i++) {
            h ^= b[i];
            h *= 1099511628211ull;
        }
    };
    mix(plan.tiles.data(), plan.tiles.size() * sizeof(TileDesc));
    mix(plan.ownedIds.data(), plan.ownedIds.size() * 4);
    mix(plan.haloIds.data(), plan.numHalo * 4);
    mix(plan.sCnt.data(), plan.sCnt.size());
    mix(plan.bCnt.data(), plan.bCnt.size());
    mix(plan.attOff.data(), plan.attOff.size() * 4);
    mix(plan.stretchRec.data(), plan.stretchRec.size() * sizeof(Rec2));
    mix(plan.bendRec.data(), plan.bendRec.size() * sizeof(Rec4));
    mix(plan.attachRec.data(), plan.attachRec.size() * sizeof(Rec2));
    const unsigned globals[5] = {plan.maxLocals, plan.maxBendPerTile, plan.maxStretchPerTile, plan.maxKS, plan.maxKB};
    mix(globals, sizeof(globals));
    *digest = h;
    VT_API_END
}

int velvet_plan_grid_smem_wavefronts(int resolution, int tileSize, unsigned long long* out3)
{
    VT_API_BEGIN
    VT_REQUIRE(resolution > 0 && out3, "plan_grid_smem_wavefronts: bad argument");
    const int R = resolution;
    const size_t n = (size_t)(R + 1) * (R + 1);
    std::vector<float> v(3 * n);
    std::vector<unsigned> idx((size_t)6 * R * R);
    GenerateClothMesh(R, v.data(), idx.data());
    const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    const GridConstraints g = GenerateGridConstraints(R, v.data(), idx.data(), identity, {}, 1.5f, 0);
    const TilePlan plan = build_tile_plan((unsigned)n, v.data(), g.stretchIdx.data(), g.stretchLen.data(), g.stretchLen.size(),
                                          g.bendIdx.data(), g.bendAngle.data(), g.bendAngle.size(), nullptr, nullptr, nullptr, 0,
                                          tileSize ? tileSize : 256);
    if (!plan.valid) return set_error(VELVET_ERR_UNSUPPORTED, plan.whyInvalid);
    out3[0] = plan.smemWavefrontsIdeal;
    out3[1] = plan.smemWavefrontsIdOrder;
    out3[2] = plan.smemWavefronts;
    VT_API_END
}

int velvet_hash_create(VelvetSpatialHash** out, float particleDiameter, int maxNumObjects, float hashCellSizeScalar,
                       int maxNumNeighbors)
{
    VT_API_BEGIN
    VT_REQUIRE(out && maxNumObjects > 0 && maxNumNeighbors > 0 && particleDiameter > 0 && hashCellSizeScalar > 0,
               "hash_create: bad argument");
    *out = new VelvetSpatialHash(particleDiameter, maxNumObjects, hashCellSizeScalar, maxNumNeighbors);
    VT_API_END
}

int velvet_hash_destroy(VelvetSpatialHash* h)
{
    VT_API_BEGIN
    delete h;
    VT_API_END
}

int velvet_hash_set_initial_positions(VelvetSpatialHash* h, const float* positions, size_t count)
{
    VT_API_BEGIN
    VT_REQUIRE(h && (positions || !count), "bad argument");
    h->impl.SetInitialPositions(positions, count);
    VT_API_END
}

int velvet_hash_hash(VelvetSpatialHash* h, const float* positions, size_t count)
{
    VT_API_BEGIN
    VT_REQUIRE(h && (positions || !count), "bad argument");
    h->impl.Hash(positions, count, h->particleDiameter, h->stream);
    VT_CUDA(cudaStreamSynchronize(h->stream));
    VT_API_END
}

int velvet_hash_buffer(VelvetSpatialHash* h, int bufferId, void** devPtr, size_t* count)
{
    VT_API_BEGIN
    VT_REQUIRE(h, "hash is NULL");
    void* p = nullptr;
    size_t n = 0;
    switch (bufferId) {
    case VELVET_BUF_NEIGHBORS: p = h->impl.neighbors.data(); n = h->impl.neighbors.size(); break;
    case VELVET_BUF_INITIALPOSITIONS: p = h->impl.initialPositions.data(); n = h->impl.initialPositions.size(); break;
    case VELVET_BUF_PARTICLEHASH: p = h->impl.particleHash.data(); n = h->impl.particleHash.size(); break;
    case VELVET_BUF_PARTICLEINDEX: p = h->impl.particleIndex.data(); n = h->impl.particleIndex.size(); break;
    case VELVET_BUF_CELLSTART: p = h->impl.cellStart.data(); n = h->impl.cellStart.size(); break;
    case VELVET_BUF_CELLEND: p = h->impl.cellEnd.data(); n = h->impl.cellEnd.size(); break;
    default: return set_error(VELVET_ERR_INVALID_ARGUMENT, "unknown hash buffer id");
    }
    if (devPtr) *devPtr = p;
    if (count) *count = n;
    VT_API_END
}

}  // extern "C"
