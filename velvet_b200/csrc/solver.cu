// solver.cu -- VtClothSolverGPU / SpatialHashGPU / VtClothObjectGPU host orchestration (see solver.hpp).
//
// Reference flow being replaced (VtClothSolverGPU.hpp L56-111): ~250 default-stream launches per frame, each
// bracketed by two cudaEventCreate + two cudaEventRecord (Timer.hpp L95-119), a cudaMemcpyToSymbol of the
// params, then cudaDeviceSynchronize.  Here a frame is one cudaMemcpyAsync of a 100-byte parameter block plus
// one cudaGraphLaunch on the solver's own stream; the graph is re-captured only when the topology changes
// (particle/constraint counts, substeps, iterations, hash interleave, buffer reallocation).
#include "solver.hpp"
#include "setup_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <cstdio>

namespace velvet {

// VELVET_SETUP_TIMING=1: wall time of the phases of registration and of the first Simulate, on stderr (diagnostic)
struct SetupClock {
    bool on;
    std::chrono::steady_clock::time_point t0;
    SetupClock() : on(getenv("VELVET_SETUP_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char* what)
    {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[velvet setup] %-44s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

void default_sim_params(VtSimParams& p)
{
    std::memset(&p, 0, sizeof(p));
    p.numSubsteps = 2;
    p.numIterations = 4;
    p.maxNumNeighbors = 64;
    p.maxSpeed = 50.0f;
    p.gravity[0] = 0.0f;
    p.gravity[1] = -9.8f;
    p.gravity[2] = 0.0f;
    p.bendCompliance = 0.0f;
    p.damping = 0.25f;
    p.relaxationFactor = 1.0f;
    p.longRangeStretchiness = 1.2f;
    p.collisionMargin = 0.06f;
    p.friction = 0.1f;
    p.enableSelfCollision = 1;
    p.interleavedHash = 3;
    p.particleDiameterScalar = 1.5f;
    p.hashCellSizeScalar = 1.5f;
}

// ------------------------------------------------------------------------------------------------ SpatialHashGPU

SpatialHashGPU::SpatialHashGPU(float particleDiameter, int maxNumObjects, float hashCellSizeScalar, int maxNumNeighbors, bool hostReadable)
{
    m_spacing = particleDiameter * hashCellSizeScalar;
    m_tableSize = 2 * maxNumObjects;
    m_maxNumNeighbors = maxNumNeighbors;
    // only kernels touch these five (the reference keeps them managed but never indexes them on the host): plain device
    // memory keeps 4.3 GB of neighbour slots per 16.7M particles out of the unified-memory pool
    if (!hostReadable) {
        neighbors.setDeviceOnly();
        particleHash.setDeviceOnly();
        particleIndex.setDeviceOnly();
        cellStart.setDeviceOnly();
        cellEnd.setDeviceOnly();
    }
    neighbors.resize((size_t)maxNumObjects * (size_t)maxNumNeighbors);
    particleHash.resize((size_t)maxNumObjects);
    particleIndex.resize((size_t)maxNumObjects);
    cellStart.resize((size_t)m_tableSize);
    cellEnd.resize((size_t)m_tableSize);
}

void SpatialHashGPU::SetInitialPositions(const float* positions, size_t count)
{
    initialPositions.resize(count);
    if (count) VT_CUDA(cudaMemcpy(initialPositions.data(), positions, count * sizeof(vec3), cudaMemcpyDefault));
}

VtHashParams SpatialHashGPU::MakeParams(size_t count, float particleDiameter) const
{
    VtHashParams hp;
    hp.numObjects = (uint)count;
    hp.cellSpacing = m_spacing;
    hp.cellSpacing2 = m_spacing * m_spacing;
    hp.tableSize = m_tableSize;
    hp.maxNumNeighbors = (uint)m_maxNumNeighbors;
    hp.particleDiameter2 = particleDiameter * particleDiameter;
    return hp;
}

void SpatialHashGPU::Hash(const float* positions, size_t count, float particleDiameter, cudaStream_t stream)
{
    if (count > particleHash.size()) throw Error(VELVET_ERR_INVALID_ARGUMENT, "SpatialHashGPU::Hash: more objects than maxNumObjects");
    if (initialPositions.size() < count) throw Error(VELVET_ERR_STATE, "SpatialHashGPU::Hash: SetInitialPositions not called");
    seam::HashObjects(particleHash, particleIndex, cellStart, cellEnd, neighbors, positions,
                      reinterpret_cast<const float*>(initialPositions.data()), MakeParams(count, particleDiameter), stream);
}

int SpatialHashGPU::ComputeIntCoord(float value) const { return int_coord(value, m_spacing); }
int SpatialHashGPU::HashCoords(int x, int y, int z) const { return hash_coords(x, y, z, m_tableSize); }
int SpatialHashGPU::HashPosition(const float* p) const
{
    return HashCoords(ComputeIntCoord(p[0]), ComputeIntCoord(p[1]), ComputeIntCoord(p[2]));
}

// ------------------------------------------------------------------------------------------------ timing helper

struct VtClothSolverGPU::Stage {
    struct Span {
        std::string label;
        cudaEvent_t a, b;
    };
    std::vector<Span> spans;
    cudaStream_t stream;
    explicit Stage(cudaStream_t st) : stream(st) {}
    void begin(const char* label)
    {
        Span s;
        s.label = label;
        VT_CUDA(cudaEventCreate(&s.a));
        VT_CUDA(cudaEventCreate(&s.b));
        VT_CUDA(cudaEventRecord(s.a, stream));
        spans.push_back(s);
    }
    void end() { VT_CUDA(cudaEventRecord(spans.back().b, stream)); }
    StageTiming collect()
    {
        StageTiming out;
        VT_CUDA(cudaStreamSynchronize(stream));
        for (auto& s : spans) {
            float ms = 0;
            VT_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
            auto it = std::find(out.labels.begin(), out.labels.end(), s.label);
            if (it == out.labels.end()) {
                out.labels.push_back(s.label);
                out.ms.push_back(ms);
            } else {
                out.ms[it - out.labels.begin()] += ms;
            }
            cudaEventDestroy(s.a);
            cudaEventDestroy(s.b);
        }
        spans.clear();
        return out;
    }
};

#define STAGE_BEGIN(t, label) \
    if (t) (t)->begin(label)
#define STAGE_END(t) \
    if (t) (t)->end()

// ------------------------------------------------------------------------------------------------ VtClothSolverGPU

VtClothSolverGPU::VtClothSolverGPU(int device, const VtSimParams* params)
{
    if (device >= 0) VT_CUDA(cudaSetDevice(device));
    VT_CUDA(cudaGetDevice(&m_device));
    VT_CUDA(cudaStreamCreateWithFlags(&m_stream, cudaStreamNonBlocking));
    if (params) simParams = *params;
    else default_sim_params(simParams);
    simParams.numParticles = 0;  // VtClothSolverGPU.hpp L29
    m_collidersDev.allocate(VT_MAX_COLLIDERS);
}

VtClothSolverGPU::~VtClothSolverGPU()
{
    cudaSetDevice(m_device);
    if (m_stream) cudaStreamSynchronize(m_stream);
    try { ddPeerClose(); } catch (...) {}
    if (m_graphExec) cudaGraphExecDestroy(m_graphExec);
    if (m_graph) cudaGraphDestroy(m_graph);
    if (m_copyStream) {
        cudaStreamSynchronize(m_copyStream);
        cudaStreamDestroy(m_copyStream);
    }
    for (int i = 0; i < 2; i++) {
        if (m_staged[i]) cudaEventDestroy(m_staged[i]);
        if (m_copyDone[i]) cudaEventDestroy(m_copyDone[i]);
    }
    if (m_stream) cudaStreamDestroy(m_stream);
}

void VtClothSolverGPU::NotifyBufferEdited(int bufferId)
{
    switch (bufferId) {
    case VELVET_BUF_INDICES:
    case VELVET_BUF_STRETCHINDICES:
    case VELVET_BUF_STRETCHLENGTHS:
    case VELVET_BUF_BENDINDICES:
    case VELVET_BUF_BENDANGLES:
    case VELVET_BUF_ATTACHPARTICLEIDS:
    case VELVET_BUF_ATTACHSLOTIDS:
    case VELVET_BUF_ATTACHDISTANCES:
        // the lists are no longer what the device generator wrote (rest angles, say, may now differ from 0): the plan is
        // derived from the lists themselves again (grid_plan.cpp), whatever they now hold
        m_generated.clear();
        invalidate();
        break;
    case VELVET_BUF_INITIALPOSITIONS:
        m_initDirty = true;
        break;
    default:
        break;
    }
}

void VtClothSolverGPU::HashFused()
{
    VT_CUDA(cudaSetDevice(m_device));
    const uint N = simParams.numParticles;
    if (!N) return;
    ensureFusedResources();
    if (!m_fusedUsable) throw Error(VELVET_ERR_UNSUPPORTED, "HashFused: the fused pipeline is not usable (" + m_fallbackReason + ")");
    SpatialHashGPU& H = *m_spatialHash;
    FusedLaunch L{m_stream, N};
    exact_math::launch_pack_float4(L, reinterpret_cast<const float*>(H.initialPositions.data()), m_init4, N);
    exact_math::launch_pack_float4(L, reinterpret_cast<const float*>(predicted.data()), m_predA, N);
    const int maxBit = (int)std::ceil(std::log2((double)H.tableSize()));
    const bool odd = RadixSorter::numPasses(maxBit) & 1;
    uint* k0 = odd ? m_keysAlt.data() : H.particleHash.data();
    uint* v0 = odd ? m_valsAlt.data() : H.particleIndex.data();
    uint* k1 = odd ? H.particleHash.data() : m_keysAlt.data();
    uint* v1 = odd ? H.particleIndex.data() : m_valsAlt.data();
    exact_math::launch_hash_particles(L, k0, v0, m_predA, H.spacing(), H.tableSize() / (int)m_instancing.count, m_instancing, H.cellStart, H.tableSize());
    m_sorter.sort(k0, v0, k1, v1, N, maxBit, m_stream);
    VtHashParams hp = H.MakeParams(N, simParams.particleDiameter);
    hp.tableSize = H.tableSize() / (int)m_instancing.count;
    if (!exact_math::launch_cache_neighbors_sorted(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, m_predA, m_init4, m_sorted, hp,
                                                   m_instancing, nullptr, 0, H.particleHash)) {
        exact_math::launch_find_cell_start(L, H.cellStart, H.cellEnd, H.particleHash);
        exact_math::launch_cache_neighbors(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, m_predA, m_init4, hp);
    }
    VT_CUDA(cudaGetLastError());
    Synchronize();
}

void VtClothSolverGPU::SetRenderTargets(int clothIndex, float* positionsDev, float* normalsDev)
{
    if (clothIndex < 0 || (size_t)clothIndex >= positions.numRanges()) throw Error(VELVET_ERR_INVALID_ARGUMENT, "SetRenderTargets: no such cloth");
    positions.attachRegistered((size_t)clothIndex, reinterpret_cast<vec3*>(positionsDev));
    normals.attachRegistered((size_t)clothIndex, reinterpret_cast<vec3*>(normalsDev));
}

void VtClothSolverGPU::SyncRenderTargets()
{
    VT_CUDA(cudaSetDevice(m_device));
    positions.sync(m_stream);
    normals.sync(m_stream);
}

namespace {
__global__ void __launch_bounds__(256) count_nonfinite_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                              const float* __restrict__ c, unsigned n, unsigned* out)
{
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    unsigned bad = 0;
    for (int k = 0; k < 3; k++) {
        bad += isfinite(a[3 * (size_t)id + k]) ? 0u : 1u;
        bad += isfinite(b[3 * (size_t)id + k]) ? 0u : 1u;
        bad += isfinite(c[3 * (size_t)id + k]) ? 0u : 1u;
    }
    if (bad) {
        atomicAdd(out, bad);
        atomicMin(out + 1, id);
    }
}
}  // namespace

VtClothSolverGPU::GrabResult VtClothSolverGPU::Grab(const float* o, const float* d)
{
    VT_CUDA(cudaSetDevice(m_device));
    if (!m_grab.data()) {
        m_grab.allocate(1);
        const input::GrabState none{~0ull, -1, 0.0f, 0.0f, 0};
        VT_CUDA(cudaMemcpyAsync(m_grab.data(), &none, sizeof(none), cudaMemcpyHostToDevice, m_stream));
    }
    input::grab(m_grab, reinterpret_cast<const float*>(positions.data()), invMasses.data(), simParams.numParticles, V3(o[0], o[1], o[2]),
                V3(d[0], d[1], d[2]), simParams.particleDiameter, m_stream);
    input::GrabState st;
    VT_CUDA(cudaMemcpyAsync(&st, m_grab.data(), sizeof(st), cudaMemcpyDeviceToHost, m_stream));
    VT_CUDA(cudaStreamSynchronize(m_stream));
    return GrabResult{st.index, st.distanceToOrigin};
}

void VtClothSolverGPU::Drag(const float* o, const float* d)
{
    if (!m_grab.data()) return;  // nothing was ever grabbed
    VT_CUDA(cudaSetDevice(m_device));
    input::drag(m_grab, reinterpret_cast<float*>(positions.data()), reinterpret_cast<float*>(velocities.data()), V3(o[0], o[1], o[2]),
                V3(d[0], d[1], d[2]), kFixedDeltaTime, m_stream);
    m_mayBeBusy = true;
}

void VtClothSolverGPU::Release()
{
    if (!m_grab.data()) return;
    VT_CUDA(cudaSetDevice(m_device));
    input::release(m_grab, invMasses.data(), m_stream);
    m_mayBeBusy = true;
}

unsigned VtClothSolverGPU::CheckNaN(unsigned* firstParticle)
{
    VT_CUDA(cudaSetDevice(m_device));
    const unsigned n = simParams.numParticles;
    unsigned host[2] = {0u, n};
    if (n) {
        m_nanScratch.allocate(2);
        VT_CUDA(cudaMemcpyAsync(m_nanScratch.data(), host, sizeof(host), cudaMemcpyHostToDevice, m_stream));
        count_nonfinite_kernel<<<(n + 255) / 256, 256, 0, m_stream>>>(reinterpret_cast<const float*>(positions.data()),
                                                                      reinterpret_cast<const float*>(velocities.data()),
                                                                      reinterpret_cast<const float*>(predicted.data()), n, m_nanScratch.data());
        VT_CUDA(cudaMemcpyAsync(host, m_nanScratch.data(), sizeof(host), cudaMemcpyDeviceToHost, m_stream));
        Synchronize();
    }
    if (firstParticle) *firstParticle = host[1];
    return host[0];
}

int VtClothSolverGPU::ReadbackPipelined(float* hostPositions, float* hostNormals)
{
    if (!m_copyStream) {
        VT_CUDA(cudaStreamCreateWithFlags(&m_copyStream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            VT_CUDA(cudaEventCreateWithFlags(&m_staged[i], cudaEventDisableTiming));
            VT_CUDA(cudaEventCreateWithFlags(&m_copyDone[i], cudaEventDisableTiming));
        }
    }
    const int slot = (int)(m_readbackSeq++ & 1u);
    const size_t floats = positions.size() * 3;
    // the staging buffer of this slot may still be feeding the copy issued two calls ago
    if (m_copyPending[slot]) VT_CUDA(cudaStreamWaitEvent(m_stream, m_copyDone[slot], 0));
    if (hostPositions && floats) {
        m_stagePos[slot].allocate(floats);
        exact_math::launch_copy_words(m_stream, positions.data(), m_stagePos[slot].data(), floats);
    }
    if (hostNormals && floats) {
        m_stageNrm[slot].allocate(floats);
        exact_math::launch_copy_words(m_stream, normals.data(), m_stageNrm[slot].data(), floats);
    }
    VT_CUDA(cudaEventRecord(m_staged[slot], m_stream));
    VT_CUDA(cudaStreamWaitEvent(m_copyStream, m_staged[slot], 0));
    if (hostPositions && floats)
        VT_CUDA(cudaMemcpyAsync(hostPositions, m_stagePos[slot].data(), floats * 4, cudaMemcpyDeviceToHost, m_copyStream));
    if (hostNormals && floats)
        VT_CUDA(cudaMemcpyAsync(hostNormals, m_stageNrm[slot].data(), floats * 4, cudaMemcpyDeviceToHost, m_copyStream));
    VT_CUDA(cudaEventRecord(m_copyDone[slot], m_copyStream));
    m_copyPending[slot] = true;
    return slot;
}

void VtClothSolverGPU::ReadbackWait(int ticket)
{
    if (ticket < 0 || ticket > 1) throw Error(VELVET_ERR_INVALID_ARGUMENT, "ReadbackWait: unknown ticket");
    if (m_copyPending[ticket]) VT_CUDA(cudaEventSynchronize(m_copyDone[ticket]));
}

void VtClothSolverGPU::Synchronize()
{
    VT_CUDA(cudaStreamSynchronize(m_stream));
    m_mayBeBusy = false;
}

// The hash buffers were sized with the maxNumNeighbors of AddCloth time (SpatialHashGPU.hpp L18-28); a larger value written
// into simParams afterwards must not make the collide kernels walk past the neighbour columns.
void VtClothSolverGPU::clampNeighborBound(VtSimParams& P) const
{
    if (m_spatialHash && P.maxNumNeighbors > m_spatialHash->maxNumNeighbors()) P.maxNumNeighbors = m_spatialHash->maxNumNeighbors();
}

// Registration calls reallocate managed buffers: the device must be idle first, whichever device is current in the caller
void VtClothSolverGPU::quiesce()
{
    if (!m_mayBeBusy) return;
    VT_CUDA(cudaSetDevice(m_device));
    Synchronize();
}

void VtClothSolverGPU::OnDestroy()
{
    Synchronize();
    positions.destroy();
    normals.destroy();
    invalidate();
}

void VtClothSolverGPU::setPipeline(int pipeline)
{
    if (pipeline != VELVET_PIPELINE_FUSED && pipeline != VELVET_PIPELINE_SEAM)
        throw Error(VELVET_ERR_INVALID_ARGUMENT, "unknown pipeline");
    m_pipeline = pipeline;
}

void VtClothSolverGPU::setMathMode(int mode)
{
    if (mode != VELVET_MATH_FAST && mode != VELVET_MATH_EXACT) throw Error(VELVET_ERR_INVALID_ARGUMENT, "unknown math mode");
    m_mathMode = mode;  // the graph key includes the mode: the next Simulate re-captures, nothing else is rebuilt
}

void VtClothSolverGPU::setIterateMode(int mode)
{
    if (mode != VELVET_ITERATE_AUTO && mode != VELVET_ITERATE_TILES) throw Error(VELVET_ERR_INVALID_ARGUMENT, "unknown iterate mode");
    if (mode == m_iterateMode) return;
    m_iterateMode = mode;
    invalidate();
}

int VtClothSolverGPU::iterateKernel()
{
    VT_CUDA(cudaSetDevice(m_device));
    ensureFusedResources();
    return m_gridUsable ? VELVET_ITERATE_GRID : VELVET_ITERATE_TILES;
}

void VtClothSolverGPU::setTileSize(int particlesPerTile)
{
    if (particlesPerTile != 0 && (particlesPerTile < 32 || particlesPerTile > VT_MAX_TILE || particlesPerTile % 32))
        throw Error(VELVET_ERR_INVALID_ARGUMENT, "tile size must be 0 (default) or a multiple of 32 in [32, 512]");
    m_tileSize = particlesPerTile;
    invalidate();
}

int VtClothSolverGPU::AddCloth(const float* vertices, int numVertices, const uint* meshIndices, int numIndices,
                               const float* modelMatrix16, float particleDiameter)
{
    if (!vertices || numVertices <= 0 || !modelMatrix16 || numIndices < 0 || (numIndices && !meshIndices))
        throw Error(VELVET_ERR_INVALID_ARGUMENT, "AddCloth: bad argument");
    if (m_instanced) throw Error(VELVET_ERR_STATE, "AddCloth: the solver holds batched instances (AddClothInstances)");
    if ((unsigned long long)simParams.numParticles + (unsigned long long)numVertices >= (1ull << 30))
        throw Error(VELVET_ERR_INVALID_ARGUMENT, "AddCloth: at most 2^30 - 1 particles (the hash table has 2 * numParticles int rows)");
    VT_CUDA(cudaSetDevice(m_device));
    Synchronize();
    SetupClock clk;
    const int prevNumParticles = (int)simParams.numParticles;
    const int newParticles = numVertices;
    m_clothRanges.push_back(ClothRange{(uint)prevNumParticles, (uint)newParticles});

    // global parameters, hpp L122-125
    simParams.numParticles += (uint)newParticles;
    simParams.particleDiameter = particleDiameter;
    simParams.deltaTime = kFixedDeltaTime;
    simParams.maxSpeed = 2 * particleDiameter / kFixedDeltaTime * simParams.numSubsteps;

    if (deviceRegistration()) {
        // every per-particle array is filled on the device, in pages created there (VtBuffer::extendOnDevice)
        const size_t n = (size_t)newParticles;
        cudaStream_t st = m_stream;
        vec3* pos = positions.registerNewBufferOnDevice(n, m_device, st);
        VT_CUDA(cudaMemcpyAsync(pos, vertices, n * sizeof(vec3), cudaMemcpyHostToDevice, st));
        VT_CUDA(cudaMemsetAsync(normals.registerNewBufferOnDevice(n, m_device, st), 0, n * sizeof(vec3), st));
        clk.lap("AddCloth: positions + normals buffers");
        uint* idx = indices.extendOnDevice((size_t)numIndices, m_device, st);
        if (numIndices) {
            VT_CUDA(cudaMemcpyAsync(idx, meshIndices, (size_t)numIndices * sizeof(uint), cudaMemcpyHostToDevice, st));
            if (prevNumParticles) setup::offset_indices(idx, idx, (size_t)numIndices, (uint)prevNumParticles, st);
        }
        clk.lap("AddCloth: mesh indices");
        VT_CUDA(cudaMemsetAsync(velocities.extendOnDevice(n, m_device, st), 0, n * sizeof(vec3), st));
        VT_CUDA(cudaMemsetAsync(predicted.extendOnDevice(n, m_device, st), 0, n * sizeof(vec3), st));
        VT_CUDA(cudaMemsetAsync(deltas.extendOnDevice(n, m_device, st), 0, n * sizeof(vec3), st));
        VT_CUDA(cudaMemsetAsync(deltaCounts.extendOnDevice(n, m_device, st), 0, n * sizeof(int), st));
        const float one = 1.0f;
        uint oneBits;
        std::memcpy(&oneBits, &one, 4);
        setup::fill_words(invMasses.extendOnDevice(n, m_device, st), n, oneBits, st);
        clk.lap("AddCloth: velocity/predicted/delta/mass buffers");
    } else {
        positions.registerNewBuffer(reinterpret_cast<const vec3*>(vertices), (size_t)newParticles);
        normals.registerNewBuffer(nullptr, (size_t)newParticles);
        clk.lap("AddCloth: positions + normals buffers");

        std::vector<uint> shifted((size_t)numIndices);
        for (int i = 0; i < numIndices; i++) shifted[i] = meshIndices[i] + (uint)prevNumParticles;
        indices.push_back(shifted);
        clk.lap("AddCloth: mesh indices");

        velocities.push_back((size_t)newParticles, V3(0, 0, 0));
        predicted.push_back((size_t)newParticles, V3(0, 0, 0));
        deltas.push_back((size_t)newParticles, V3(0, 0, 0));
        deltaCounts.push_back((size_t)newParticles, 0);
        invMasses.push_back((size_t)newParticles, 1.0f);
        clk.lap("AddCloth: velocity/predicted/delta/mass buffers");
    }

    // world transform on the device, hpp L143-145
    seam::InitializePositions(reinterpret_cast<float*>(positions.data()), prevNumParticles, newParticles, modelMatrix16, m_stream);
    Synchronize();
    clk.lap("AddCloth: world transform");

    // hash sized to the total particle count; snapshot taken after the transform, hpp L148-149
    m_spatialHash = std::make_shared<SpatialHashGPU>(particleDiameter, (int)simParams.numParticles,
                                                     simParams.hashCellSizeScalar, simParams.maxNumNeighbors, m_hashHostReadable);
    m_spatialHash->SetInitialPositions(reinterpret_cast<const float*>(positions.data()), positions.size());
    clk.lap("AddCloth: spatial hash");
    invalidate();
    return prevNumParticles;
}

void VtClothSolverGPU::AddClothInstances(int R, const float* vertices, const uint* meshIndices, const float* models,
                                         int numInstances, const int* attachedIndices, int numAttached)
{
    if (R <= 0 || !vertices || !meshIndices || !models || numInstances <= 0 || numAttached < 0 || (numAttached && !attachedIndices))
        throw Error(VELVET_ERR_INVALID_ARGUMENT, "AddClothInstances: bad argument");
    if (numInstances > 65535) throw Error(VELVET_ERR_INVALID_ARGUMENT, "AddClothInstances: at most 65535 instances");
    if (simParams.numParticles != 0) throw Error(VELVET_ERR_STATE, "AddClothInstances must be the only registration on a solver");
    VT_CUDA(cudaSetDevice(m_device));
    Synchronize();
    const size_t n = (size_t)(R + 1) * (R + 1);
    const size_t ni = (size_t)6 * R * R;
    const size_t total = n * (size_t)numInstances;
    if (total >= (1ull << 30)) throw Error(VELVET_ERR_INVALID_ARGUMENT, "AddClothInstances: at most 2^30 - 1 particles in total");
    const std::vector<int> attached(attachedIndices, attachedIndices + numAttached);
    for (int a : attached)
        if (a < 0 || (size_t)a >= n) throw Error(VELVET_ERR_INVALID_ARGUMENT, "AddClothInstances: attached index out of range");

    const GridConstraints g = GenerateGridConstraints(R, vertices, meshIndices, models, attached, simParams.particleDiameterScalar, 0);
    simParams.numParticles = (uint)total;
    simParams.particleDiameter = g.particleDiameter;
    simParams.deltaTime = kFixedDeltaTime;
    simParams.maxSpeed = 2 * g.particleDiameter / kFixedDeltaTime * simParams.numSubsteps;

    // state for every instance; mesh indices / constraints once (per-instance local indices)
    const bool onDevice = deviceRegistration();
    if (onDevice) {
        // per-particle arrays of all instances written on the device, in pages created there (VtBuffer::extendOnDevice)
        cudaStream_t st = m_stream;
        DeviceBuffer<float> vtx, mdl;
        vtx.upload(vertices, 3 * n, st);
        mdl.upload(models, 16 * (size_t)numInstances, st);
        vec3* pos = positions.registerNewBuffersOnDevice(n, (size_t)numInstances, m_device, st);
        setup::instance_positions(reinterpret_cast<float*>(pos), vtx, mdl, (uint)n, (uint)numInstances, st);
        VT_CUDA(cudaMemsetAsync(normals.registerNewBuffersOnDevice(n, (size_t)numInstances, m_device, st), 0, total * sizeof(vec3), st));
        VT_CUDA(cudaMemsetAsync(velocities.extendOnDevice(total, m_device, st), 0, total * sizeof(vec3), st));
        VT_CUDA(cudaMemsetAsync(predicted.extendOnDevice(total, m_device, st), 0, total * sizeof(vec3), st));
        VT_CUDA(cudaMemsetAsync(deltas.extendOnDevice(total, m_device, st), 0, total * sizeof(vec3), st));
        VT_CUDA(cudaMemsetAsync(deltaCounts.extendOnDevice(total, m_device, st), 0, total * sizeof(int), st));
        const float one = 1.0f;
        uint oneBits;
        std::memcpy(&oneBits, &one, 4);
        float* w = invMasses.extendOnDevice(total, m_device, st);
        setup::fill_words(w, total, oneBits, st);
        std::vector<int> pinned;  // hpp L172, in every instance
        for (size_t c = 0; c < g.attachDist.size(); c++)
            if (g.attachDist[c] == 0) pinned.push_back(g.attachPid[c]);
        DeviceBuffer<int> pinnedDev;
        pinnedDev.upload(pinned, st);
        setup::instance_pin(w, pinnedDev, (uint)pinned.size(), (uint)n, (uint)numInstances, st);
        Synchronize();  // the staging arrays go out of scope
    } else {
        for (int k = 0; k < numInstances; k++) {
            positions.registerNewBuffer(reinterpret_cast<const vec3*>(vertices), n);
            normals.registerNewBuffer(nullptr, n);
        }
        velocities.push_back(total, V3(0, 0, 0));
        predicted.push_back(total, V3(0, 0, 0));
        deltas.push_back(total, V3(0, 0, 0));
        deltaCounts.push_back(total, 0);
        invMasses.push_back(total, 1.0f);
        for (int k = 0; k < numInstances; k++)
            seam::InitializePositions(reinterpret_cast<float*>(positions.data()), (int)(k * n), (int)n, models + 16 * (size_t)k, m_stream);
        Synchronize();
    }
    indices.append(meshIndices, ni);

    stretchIndices.append(g.stretchIdx.data(), g.stretchIdx.size());
    stretchLengths.append(g.stretchLen.data(), g.stretchLen.size());
    bendIndices.append(g.bendIdx.data(), g.bendIdx.size());
    bendAngles.append(g.bendAngle.data(), g.bendAngle.size());
    attachParticleIDs.append(g.attachPid.data(), g.attachPid.size());
    attachSlotIDs.append(g.attachSlot.data(), g.attachSlot.size());
    attachDistances.append(g.attachDist.data(), g.attachDist.size());
    // slot positions of every instance: the world position of the attached vertex under that instance's matrix
    // (computed here with the function the device transform applies: indexing positions[] would fault device pages back)
    for (int k = 0; k < numInstances; k++)
        for (int a : attached) attachSlotPositions.push_back(mul_point(models + 16 * (size_t)k, load3(vertices, (size_t)a), 1.0f));
    if (!onDevice)
        for (size_t c = 0; c < g.attachDist.size(); c++)
            if (g.attachDist[c] == 0)  // hpp L172, in every instance
                for (int k = 0; k < numInstances; k++) invMasses[k * n + (size_t)g.attachPid[c]] = 0;

    m_spatialHash = std::make_shared<SpatialHashGPU>(g.particleDiameter, (int)total, simParams.hashCellSizeScalar,
                                                     simParams.maxNumNeighbors, m_hashHostReadable);
    m_spatialHash->SetInitialPositions(reinterpret_cast<const float*>(positions.data()), positions.size());
    m_instancing = Instancing{(uint)numInstances, (uint)n, (uint)numAttached};
    m_clothRanges.assign(1, ClothRange{0u, (uint)n});  // one topology, shared by every instance
    m_instanced = true;
    invalidate();
}

// Band size of the candidate walk: ids of a grid cloth are row-major, so index bands are compact pieces of cloth (hash_kernels.cuh)
// 0 / 1: one Jacobi iteration per launch; otherwise all iterations of a substep in one launch (VELVET_GRID_MULTI=0 for A/B runs)
unsigned VtClothSolverGPU::gridIterationsPerLaunch() const
{
    const char* e = getenv("VELVET_GRID_MULTI");
    return (e && e[0] == '0') ? 1u : 2u;
}

unsigned VtClothSolverGPU::walkBandParticles() const
{
    return (m_gridUsable && !m_instanced && simParams.numParticles > VT_WALK_BAND_MIN_PARTICLES) ? VT_WALK_BAND_PARTICLES : 0u;
}

bool VtClothSolverGPU::deviceRegistration()
{
    const char* e = getenv("VELVET_HOST_GENERATE");
    return !(e && *e && std::strcmp(e, "0") != 0);
}

void VtClothSolverGPU::GenerateGridClothOnDevice(int R, int base, const float* vertices, const float* M,
                                                 const std::vector<int>& attachedIndices)
{
    if (R <= 0 || base < 0 || !vertices || !M) throw Error(VELVET_ERR_INVALID_ARGUMENT, "GenerateGridClothOnDevice: bad argument");
    const size_t nv = (size_t)(R + 1) * (R + 1);
    if ((size_t)base + nv > positions.size() || indices.size() < (size_t)6 * R * R)
        throw Error(VELVET_ERR_STATE, "GenerateGridClothOnDevice: the cloth is not registered (AddCloth first)");
    VT_CUDA(cudaSetDevice(m_device));
    quiesce();
    SetupClock clk;
    cudaStream_t st = m_stream;
    const float* world = reinterpret_cast<const float*>(positions.data());  // AddCloth has applied the model matrix
    GeneratedCloth g;
    g.base = (uint)base;
    g.R = R;
    g.stretchBegin = stretchLengths.size();
    g.bendBegin = bendAngles.size();
    g.attachBegin = attachParticleIDs.size();
    g.numSlots = (uint)attachedIndices.size();
    g.firstSlot = (uint)attachSlotPositions.size();

    // GenerateStretch, VtClothObjectGPU.hpp L75-116
    const size_t numStretch = 4 * (size_t)R * R + 2 * (size_t)R;
    int* sIdx = stretchIndices.extendOnDevice(2 * numStretch, m_device, st);
    float* sLen = stretchLengths.extendOnDevice(numStretch, m_device, st);
    setup::generate_stretch(sIdx, sLen, world, (uint)base, R, st);

    // GenerateAttach, L134-148: AddAttachSlot, then that slot's constraint for every particle, slot by slot
    for (size_t slot = 0; slot < attachedIndices.size(); slot++) {
        const int a = attachedIndices[slot];
        if (a < 0 || (size_t)a >= nv) throw Error(VELVET_ERR_INVALID_ARGUMENT, "attached index out of range");
        const vec3 slotPos = mul_point(M, load3(vertices, (size_t)a), 1.0f);  // the bits the device transform produced
        attachSlotPositions.push_back(slotPos);
        int* pid = attachParticleIDs.extendOnDevice(nv, m_device, st);
        int* sid = attachSlotIDs.extendOnDevice(nv, m_device, st);
        float* dist = attachDistances.extendOnDevice(nv, m_device, st);
        setup::generate_attach(pid, sid, dist, world, invMasses.data(), (uint)base, (uint)nv, (int)(g.firstSlot + slot), slotPos, st);
    }

    // GenerateBending, L118-132: the triangles' own indices (i, i+5, i+2, i+1), rest angle 0
    const size_t numQuads = (size_t)R * R;
    const size_t meshBegin = indices.size() - 6 * numQuads;  // this cloth's (shifted) triangles: the last AddCloth's
    uint* bIdx = bendIndices.extendOnDevice(4 * numQuads, m_device, st);
    float* bAng = bendAngles.extendOnDevice(numQuads, m_device, st);
    setup::generate_bend(bIdx, bAng, indices.data() + meshBegin, numQuads, st);
    m_generated.push_back(g);
    invalidate();
    VT_CUDA(cudaStreamSynchronize(st));  // the lists are public: the host may index them as soon as Start returns
    clk.lap("Start: constraint lists (device)");
}

bool VtClothSolverGPU::generatedListsIntact() const
{
    if (m_instanced || m_generated.size() != m_clothRanges.size() || m_generated.empty()) return false;
    size_t s = 0, b = 0, a = 0;
    uint slots = 0;
    for (size_t k = 0; k < m_generated.size(); k++) {
        const GeneratedCloth& g = m_generated[k];
        const size_t nv = (size_t)(g.R + 1) * (g.R + 1);
        if (g.base != m_clothRanges[k].base || nv != m_clothRanges[k].count) return false;
        if (g.stretchBegin != s || g.bendBegin != b || g.attachBegin != a || g.firstSlot != slots) return false;
        s += 4 * (size_t)g.R * g.R + 2 * (size_t)g.R;
        b += (size_t)g.R * g.R;
        a += nv * g.numSlots;
        slots += g.numSlots;
    }
    return s == stretchLengths.size() && b == bendAngles.size() && a == attachParticleIDs.size() &&
           slots == attachSlotPositions.size();
}

// The implicit-grid plan of generated cloths, on the device (grid_plan.cpp is the host builder for everything else).
// Leaves m_setupFlags[1] != 0 when the mesh was not triangulated like the reference's grid.
bool VtClothSolverGPU::buildGridPlanOnDevice(uint planN, cudaStream_t st)
{
    if (m_generated.size() > GRID_MAX_CLOTHS) return false;
    GridPlan plan;
    std::vector<uint> sides;
    for (const GeneratedCloth& g : m_generated) {
        GridCloth gc;
        gc.base = g.base;
        gc.side = (uint)g.R + 1;
        gc.tilesY = gc.firstTile = 0;
        plan.cloths.push_back(gc);
        sides.push_back(gc.side);
    }
    choose_grid_tile_shape(sides, m_squareTilesOnly, plan.tileX, plan.tileY);
    plan.numTiles = lay_out_grid_tiles(plan.cloths, plan.tileX, plan.tileY);
    plan.valid = true;
    const size_t numAttach = attachParticleIDs.size();
    m_gCloths.upload(plan.cloths, st);
    m_gRest4.allocate(planN);
    m_gAngle.release();  // every generated quad rests flat: the kernel takes the angle as a scalar
    m_gAttOff.allocate((size_t)planN + 1);
    m_gAttachRec.allocate(numAttach + 1);
    uint attBase = 0;
    for (const GeneratedCloth& g : m_generated) {
        const uint nv = (uint)((g.R + 1) * (g.R + 1));
        setup::grid_plan_from_lists(m_gRest4, stretchLengths.data() + g.stretchBegin, bendIndices.data() + 4 * g.bendBegin, g.base, g.R,
                                    m_setupFlags.data() + 1, st);
        setup::grid_attach_records(m_gAttOff, m_gAttachRec, attachDistances.data() + g.attachBegin, g.base, nv, g.numSlots, g.firstSlot,
                                   attBase, st);
        attBase += nv * g.numSlots;
    }
    setup::fill_words(m_gAttOff.data() + planN, 1, attBase, st);
    m_gridPlan = std::move(plan);
    return true;
}

void VtClothSolverGPU::AddStretch(int idx1, int idx2, float distance)
{
    quiesce();
    stretchIndices.push_back(idx1);
    stretchIndices.push_back(idx2);
    stretchLengths.push_back(distance);
    invalidate();
}

void VtClothSolverGPU::AddAttachSlot(const float* p)
{
    quiesce();
    attachSlotPositions.push_back(V3(p[0], p[1], p[2]));
    invalidate();
}

void VtClothSolverGPU::AddAttach(int particleIndex, int slotIndex, float distance)
{
    quiesce();
    if (particleIndex < 0 || (size_t)particleIndex >= invMasses.size())
        throw Error(VELVET_ERR_INVALID_ARGUMENT, "AddAttach: particle index out of range");
    if (distance == 0) invMasses[(size_t)particleIndex] = 0;  // hpp L172
    attachParticleIDs.push_back(particleIndex);
    attachSlotIDs.push_back(slotIndex);
    attachDistances.push_back(distance);
    invalidate();
}

void VtClothSolverGPU::AddBend(uint idx1, uint idx2, uint idx3, uint idx4, float angle)
{
    quiesce();
    bendIndices.push_back(idx1);
    bendIndices.push_back(idx2);
    bendIndices.push_back(idx3);
    bendIndices.push_back(idx4);
    bendAngles.push_back(angle);
    invalidate();
}

void VtClothSolverGPU::AddStretchBulk(const int* idxPairs, const float* distances, size_t n)
{
    quiesce();
    stretchIndices.append(idxPairs, 2 * n);
    stretchLengths.append(distances, n);
    invalidate();
}

void VtClothSolverGPU::AddBendBulk(const uint* idxQuads, const float* angles, size_t n)
{
    quiesce();
    bendIndices.append(idxQuads, 4 * n);
    bendAngles.append(angles, n);
    invalidate();
}

void VtClothSolverGPU::AddAttachBulk(const int* particleIds, const int* slotIds, const float* distances, size_t n)
{
    quiesce();
    for (size_t i = 0; i < n; i++) {
        if (particleIds[i] < 0 || (size_t)particleIds[i] >= invMasses.size())
            throw Error(VELVET_ERR_INVALID_ARGUMENT, "AddAttach: particle index out of range");
        if (distances[i] == 0) invMasses[(size_t)particleIds[i]] = 0;
    }
    attachParticleIDs.append(particleIds, n);
    attachSlotIDs.append(slotIds, n);
    attachDistances.append(distances, n);
    invalidate();
}

void VtClothSolverGPU::UpdateColliders(const VtSDFCollider* colliders, int numColliders)
{
    if (numColliders < 0 || (numColliders && !colliders)) throw Error(VELVET_ERR_INVALID_ARGUMENT, "UpdateColliders: bad argument");
    VT_CUDA(cudaSetDevice(m_device));
    // The seam pipeline reads the managed block directly (like the reference): drain the previous frame first.
    // The fused pipeline reads a device copy that is refreshed in stream order, so no host sync is needed.
    if (m_ddReady && (unsigned)numColliders > VT_MAX_COLLIDERS)
        throw Error(VELVET_ERR_UNSUPPORTED, "a decomposed cloth takes at most 64 SDF colliders");
    if (m_pipeline != VELVET_PIPELINE_FUSED) Synchronize();
    sdfColliders.resize((size_t)numColliders);
    if (numColliders) {
        std::memcpy(sdfColliders.data(), colliders, sizeof(VtSDFCollider) * (size_t)numColliders);
        if ((unsigned)numColliders <= VT_MAX_COLLIDERS)
            VT_CUDA(cudaMemcpyAsync(m_collidersDev.data(), colliders, sizeof(VtSDFCollider) * (size_t)numColliders,
                                    cudaMemcpyHostToDevice, m_stream));
    }
}

// ---- the reference's launch order over the seam kernels (pipeline = SEAM), VtClothSolverGPU.hpp L62-101
void VtClothSolverGPU::simulateSeam(float frameTime, Stage* t)
{
    const VtSimParams& P = simParams;
    const float substepTime = frameTime / (float)P.numSubsteps;
    cudaStream_t st = m_stream;
    float* pos = reinterpret_cast<float*>(positions.data());
    float* pred = reinterpret_cast<float*>(predicted.data());
    float* vel = reinterpret_cast<float*>(velocities.data());
    float* dlt = reinterpret_cast<float*>(deltas.data());
    const float* slots = reinterpret_cast<const float*>(attachSlotPositions.data());
    const uint numColliders = (uint)sdfColliders.size();
    int launches = 0;

    STAGE_BEGIN(t, "Solver_CollideSDFs");
    seam::CollideSDF(P, pos, sdfColliders, pos, numColliders, frameTime, st);
    launches += seam::LastLaunchCount();
    STAGE_END(t);
    for (int substep = 0; substep < P.numSubsteps; substep++) {
        STAGE_BEGIN(t, "Solver_Predict");
        seam::PredictPositions(P, pred, vel, pos, substepTime, st);
        launches += seam::LastLaunchCount();
        STAGE_END(t);
        if (P.enableSelfCollision) {
            if (substep % P.interleavedHash == 0) {
                STAGE_BEGIN(t, "Solver_Hash");
                m_spatialHash->Hash(pred, predicted.size(), P.particleDiameter, st);
                launches += seam::LastLaunchCount();
                STAGE_END(t);
            }
            STAGE_BEGIN(t, "Solver_CollideParticles");
            seam::CollideParticles(P, dlt, deltaCounts, pred, invMasses, m_spatialHash->neighbors, pos, st);
            launches += seam::LastLaunchCount();
            STAGE_END(t);
        }
        STAGE_BEGIN(t, "Solver_CollideSDFs");
        seam::CollideSDF(P, pred, sdfColliders, pos, numColliders, substepTime, st);
        launches += seam::LastLaunchCount();
        STAGE_END(t);
        for (int iteration = 0; iteration < P.numIterations; iteration++) {
            STAGE_BEGIN(t, "Solver_SolveStretch");
            seam::SolveStretch(pred, dlt, deltaCounts, stretchIndices, stretchLengths, invMasses, (uint)stretchLengths.size(), st);
            launches += seam::LastLaunchCount();
            STAGE_END(t);
            STAGE_BEGIN(t, "Solver_SolveAttach");
            seam::SolveAttachment(P, pred, dlt, deltaCounts, invMasses, attachParticleIDs, attachSlotIDs, slots,
                                  attachDistances, (int)attachParticleIDs.size(), st);
            launches += seam::LastLaunchCount();
            STAGE_END(t);
            STAGE_BEGIN(t, "Solver_SolveBending");
            seam::SolveBending(P, pred, dlt, deltaCounts, bendIndices, bendAngles, invMasses, (uint)bendAngles.size(),
                               substepTime, st);
            launches += seam::LastLaunchCount();
            STAGE_END(t);
            STAGE_BEGIN(t, "Solver_ApplyDeltas");
            seam::ApplyDeltas(P, pred, dlt, deltaCounts, st);
            launches += seam::LastLaunchCount();
            STAGE_END(t);
        }
        STAGE_BEGIN(t, "Solver_Finalize");
        seam::Finalize(P, vel, pos, pred, substepTime, st);
        launches += seam::LastLaunchCount();
        STAGE_END(t);
    }
    STAGE_BEGIN(t, "Solver_UpdateNormals");
    seam::ComputeNormal(P, reinterpret_cast<float*>(normals.data()), pos, indices, (uint)(indices.size() / 3), st);
    launches += seam::LastLaunchCount();
    STAGE_END(t);
    m_lastLaunches = launches;
}

// ---- the record-driven tile plan and its device copy (tile_plan.hpp); false + m_fallbackReason when the mesh cannot be tiled
bool VtClothSolverGPU::buildTilePlan()
{
    const uint planN = m_instancing.particles;
    cudaStream_t st = m_stream;
    const int tileSize = m_tileSize ? m_tileSize : 256;  // power of two: the kernel is specialised on log2(tile)
    m_plan = build_tile_plan(planN, reinterpret_cast<const float*>(m_spatialHash->initialPositions.data()), stretchIndices.data(),
                             stretchLengths.data(), stretchLengths.size(), bendIndices.data(), bendAngles.data(),
                             bendAngles.size(), attachParticleIDs.data(), attachSlotIDs.data(), attachDistances.data(),
                             attachParticleIDs.size(), tileSize);
    if (!m_plan.valid) {
        m_fallbackReason = m_plan.whyInvalid;
        return false;
    }
    m_dTiles.upload(m_plan.tiles, st);
    m_dOwned.upload(m_plan.ownedIds, st);
    m_dHalo.upload(m_plan.haloIds, st);
    {
        std::vector<uint16_t> cnt16(m_plan.sCnt.size());  // stretch | bend << 8 per owned particle: one load in the kernel
        for (size_t i = 0; i < cnt16.size(); i++) cnt16[i] = (uint16_t)(m_plan.sCnt[i] | (m_plan.bCnt[i] << 8));
        m_dCnt16.upload(cnt16, st);
    }
    m_dAttOff.upload(m_plan.attOff, st);
    m_dStretchRec.upload(reinterpret_cast<const uint2*>(m_plan.stretchRec.data()), m_plan.stretchRec.size(), st);
    m_dBendRec.upload(reinterpret_cast<const uint4*>(m_plan.bendRec.data()), m_plan.bendRec.size(), st);
    m_dAttachRec.upload(reinterpret_cast<const uint2*>(m_plan.attachRec.data()), m_plan.attachRec.size(), st);
    m_planDev.tiles = m_dTiles;
    m_planDev.ownedIds = m_dOwned;
    m_planDev.haloIds = m_dHalo;
    m_planDev.cnt16 = m_dCnt16;
    m_planDev.attOff = m_dAttOff;
    m_planDev.stretchRec = m_dStretchRec;
    m_planDev.bendRec = m_dBendRec;
    m_planDev.attachRec = m_dAttachRec;
    m_planDev.numTiles = (uint)m_plan.tiles.size();
    m_planDev.maxLocals = m_plan.maxLocals;
    m_planDev.maxKS = m_plan.maxKS;
    m_planDev.maxKB = m_plan.maxKB;
    m_planDev.maxBendPerTile = m_plan.maxBendPerTile;
    m_planDev.maxStretchPerTile = m_plan.maxStretchPerTile;
    m_planDev.tileSize = (uint)m_plan.tileSize;
    m_planDev.threads = m_plan.tileSize <= 128 ? 128u : (m_plan.tileSize <= 256 ? 256u : 512u);
    m_planDev.hasAttach = attachParticleIDs.size() ? 1u : 0u;
    // Experiment knob (VELVET_ITERATE_WIDE=1): a quarter more threads per CTA, so that a tile touched by more bending
    // constraints than it has particles (289 vs 256 on a grid) needs one bend trip per warp.  Measured on B200 at 1M
    // particles: 75.9 us per iteration against 70.0 us with T threads -- three resident CTAs per SM instead of four cost
    // more than the shorter bend phase saves -- so it stays off.
    m_planDev.ctaThreads = m_planDev.threads;
    if (const char* e = getenv("VELVET_ITERATE_WIDE"))
        if (atoi(e) != 0) m_planDev.ctaThreads += m_planDev.threads / 4;
    const size_t smem = exact_math::iterate_smem_bytes(m_planDev);
    if (smem > 200 * 1024) {
        m_fallbackReason = "tile needs more than 200 KB of shared memory";
        return false;
    }
    m_planDev.residentCtas = std::min(exact_math::configure_iterate_kernel(smem, m_planDev.threads, m_planDev.ctaThreads),
                                      fast_math::configure_iterate_kernel(smem, m_planDev.threads, m_planDev.ctaThreads));

    m_tilePlanBuilt = true;
    return true;
}

void VtClothSolverGPU::ensureTilePlan()
{
    VT_CUDA(cudaSetDevice(m_device));
    ensureFusedResources();
    if (!m_fusedUsable) throw Error(VELVET_ERR_UNSUPPORTED, "the fused pipeline is unavailable (" + m_fallbackReason + ")");
    if (m_tilePlanBuilt) return;
    Synchronize();
    if (!buildTilePlan()) throw Error(VELVET_ERR_UNSUPPORTED, "the mesh cannot be tiled (" + m_fallbackReason + ")");
    VT_CUDA(cudaStreamSynchronize(m_stream));
}

// ---- fused pipeline resources: SoA state, tile plan, vertex->triangle CSR (rebuilt when the topology changes)
void VtClothSolverGPU::ensureFusedResources()
{
    if (!m_topologyDirty) {
        if (m_initDirty && m_fusedUsable) {  // initialPositions were edited in place: refresh the float4 copy the hash filters with
            exact_math::launch_pack_float4(FusedLaunch{m_stream, simParams.numParticles},
                                           reinterpret_cast<const float*>(m_spatialHash->initialPositions.data()), m_init4,
                                           simParams.numParticles);
            m_initDirty = false;
        }
        return;
    }
    m_initDirty = false;  // the rebuild below packs the current initialPositions
    Synchronize();
    SetupClock clk;
    const uint N = simParams.numParticles;
    if (!m_instanced) m_instancing = Instancing{1u, N, (uint)attachSlotPositions.size()};
    const uint planN = m_instancing.particles;  // the tile plan / vertex CSR cover one instance
    m_fusedUsable = false;
    m_fallbackReason.clear();
    if (m_graphExec) {
        cudaGraphExecDestroy(m_graphExec);
        m_graphExec = nullptr;
    }
    if (m_graph) {
        cudaGraphDestroy(m_graph);
        m_graph = nullptr;
    }
    m_topologyDirty = false;
    if (N == 0) {
        m_fallbackReason = "no particles";
        return;
    }

    cudaStream_t st = m_stream;
    // lists that are exactly what GenerateGridClothOnDevice wrote stay on the device: the checks below run there too
    const bool generated = generatedListsIntact();
    if (!generated)
        for (size_t i = 0; i < attachSlotIDs.size(); i++)
            if (attachSlotIDs[i] < 0 || (size_t)attachSlotIDs[i] >= (size_t)m_instancing.slots) {
                m_fallbackReason = "attach slot index out of range";
                return;
            }

    // vertex -> incident triangles (ascending triangle id) and the mesh index range check, on the device
    m_setupFlags.allocate(2);
    VT_CUDA(cudaMemsetAsync(m_setupFlags.data(), 0, 2 * sizeof(int), st));
    {
        const size_t numIdx = indices.size() - indices.size() % 3;
        m_vtxTriOff.allocate((size_t)planN + 1);
        m_vtxTris.allocate(std::max<size_t>(numIdx, 1));
        DeviceBuffer<uint> cursor;
        cursor.allocate(planN);
        clk.lap("  vtx-tri: allocations");
        setup::vertex_triangles(indices.data(), numIdx, planN, m_vtxTriOff, m_vtxTris, cursor, m_setupFlags.data(), st);
        clk.lap("  vtx-tri: launches");
        VT_CUDA(cudaStreamSynchronize(st));  // `cursor` goes out of scope
        clk.lap("  vtx-tri: sync");
    }
    clk.lap("resources: vertex -> triangle lists (device)");

    m_tilePlanBuilt = false;
    // grid cloths carrying exactly the reference's constraint pattern get the implicit-grid kernel (grid_plan.hpp)
    m_gridUsable = false;
    m_gridPlan = GridPlan{};
    int mode = m_iterateMode;
    if (const char* e = getenv("VELVET_ITERATE"))
        if (std::string(e) == "tiles") mode = VELVET_ITERATE_TILES;
    bool planOnDevice = false;
    if (mode == VELVET_ITERATE_AUTO && generated) planOnDevice = buildGridPlanOnDevice(planN, st);
    int flags[2] = {0, 0};
    VT_CUDA(cudaMemcpyAsync(flags, m_setupFlags.data(), sizeof(flags), cudaMemcpyDeviceToHost, st));
    VT_CUDA(cudaStreamSynchronize(st));
    if (flags[0]) {
        m_fallbackReason = "triangle index out of range";
        return;
    }
    if (planOnDevice && flags[1]) {  // a mesh triangulated differently: the host builder below says why and the tile plan takes over
        planOnDevice = false;
        m_gridPlan = GridPlan{};
    }
    if (mode == VELVET_ITERATE_AUTO) {
        if (!planOnDevice) {
            m_gridPlan = build_grid_plan(planN, m_clothRanges, stretchIndices.data(), stretchLengths.data(), stretchLengths.size(),
                                         bendIndices.data(), bendAngles.data(), bendAngles.size(), attachParticleIDs.data(),
                                         attachSlotIDs.data(), attachDistances.data(), attachParticleIDs.size(), m_squareTilesOnly);
            if (m_gridPlan.valid) {
                m_gCloths.upload(m_gridPlan.cloths, st);
                m_gRest4.upload(reinterpret_cast<const float4*>(m_gridPlan.rest4.data()), planN, st);
                m_gAngle.upload(m_gridPlan.restAngle, st);
                m_gAttOff.upload(m_gridPlan.attOff, st);
                m_gAttachRec.upload(reinterpret_cast<const uint2*>(m_gridPlan.attachRec.data()), m_gridPlan.attachRec.size() / 2, st);
            }
        }
        if (m_gridPlan.valid) {
            m_gridDev.cloths = m_gCloths;
            m_gridDev.rest4 = m_gRest4;
            m_gridDev.restAngle = m_gAngle;
            m_gridDev.uniformAngle = 0.0f;
            if (planOnDevice) {
                m_gridDev.restAngle = nullptr;  // generated quads rest flat
            } else {  // one rest angle for every quad (what the reference registers): the kernel takes it as a scalar
                bool uniform = true;
                for (size_t i = 1; i < bendAngles.size() && uniform; i++) uniform = bendAngles[i] == bendAngles[0];
                if (uniform && bendAngles.size()) {
                    m_gridDev.restAngle = nullptr;
                    m_gridDev.uniformAngle = bendAngles[0];
                }
            }
            m_gridDev.attOff = m_gAttOff;
            m_gridDev.attachRec = m_gAttachRec;
            m_gridDev.numCloths = (uint)m_gridPlan.cloths.size();
            m_gridDev.numTiles = m_gridPlan.numTiles;
            m_gridDev.tilesY0 = m_gridPlan.cloths[0].tilesY;
            m_gridDev.tileX = m_gridPlan.tileX;
            m_gridDev.tileY = m_gridPlan.tileY;
            m_gridDev.hasAttach = attachParticleIDs.size() ? 1u : 0u;
            m_gridDev.residentCtas = std::min(exact_math::configure_iterate_grid_kernel(), fast_math::configure_iterate_grid_kernel());
            m_gridUsable = true;
        }
    }
    clk.lap(planOnDevice ? "resources: grid plan (device)" : "resources: grid plan (host) + upload");

    // the record-driven tile plan (0.35 s of host work per million particles) only when the grid kernel cannot run; the
    // decomposed tile form and the plan accessors build it on demand (ensureTilePlan)
    if (!m_gridUsable && !buildTilePlan()) return;
    clk.lap("resources: tile plan");

    clk.lap("  (before allocations)");
    m_pos4.allocate(N);
    // at least 2 MB each: the decomposed mode exports them over CUDA IPC and must not share a driver slab with other arrays
    m_predA.allocate(std::max<size_t>(N, 131072));
    m_predB.allocate(std::max<size_t>(N, 131072));
    m_init4.allocate(N);
    m_gridBarrier.allocate(4);
    m_sorted.allocate(exact_math::cache_neighbors_scratch_float4(N));
    m_keysAlt.allocate(N);
    m_valsAlt.allocate(N);
    m_prepared.allocate(VT_MAX_COLLIDERS);
    m_frameParams.allocate(1);
    m_slotsDev.allocate(std::max<size_t>(3 * attachSlotPositions.size(), 3));
    clk.lap("  allocations: cudaMalloc x 11");
    m_sorter.reserve(N);
    clk.lap("  allocations: sorter");
    FusedLaunch L{st, N};
    exact_math::launch_pack_float4(L, reinterpret_cast<const float*>(m_spatialHash->initialPositions.data()), m_init4, N);
    clk.lap("resources: device allocations");

    // the big public buffers live in managed memory (reference contract): make them device-resident now
    auto prefetch = [&](const void* p, size_t bytes) {
        if (p && bytes) cudaMemPrefetchAsync(p, bytes, m_device, st);
    };
    prefetch(positions.data(), positions.size() * sizeof(vec3));
    prefetch(normals.data(), normals.size() * sizeof(vec3));
    prefetch(velocities.data(), velocities.size() * sizeof(vec3));
    prefetch(predicted.data(), predicted.size() * sizeof(vec3));
    prefetch(invMasses.data(), invMasses.size() * sizeof(float));
    prefetch(indices.data(), indices.size() * sizeof(uint));
    (void)cudaGetLastError();  // prefetch is best effort
    VT_CUDA(cudaStreamSynchronize(st));
    clk.lap("resources: prefetch + sync");
    m_fusedUsable = true;
}

unsigned long long VtClothSolverGPU::topologyKey() const
{
    // everything the captured launch sequence depends on (values inside FrameParams are read at run time)
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) {
        h ^= v;
        h *= 1099511628211ull;
    };
    mix(simParams.numParticles);
    mix((unsigned)simParams.numSubsteps);
    mix((unsigned)simParams.numIterations);
    mix(simParams.enableSelfCollision ? 1u : 0u);
    mix((unsigned)simParams.interleavedHash);
    mix((unsigned)simParams.maxNumNeighbors);
    mix(positions.generation());
    mix(normals.generation());
    mix(velocities.generation());
    mix(predicted.generation());
    mix(invMasses.generation());
    mix(indices.generation());
    mix(attachSlotPositions.generation());
    mix(attachSlotPositions.size());
    mix((unsigned long long)(uintptr_t)m_spatialHash.get());
    mix((unsigned)m_mathMode);
    mix((unsigned)m_iterateMode);
    return h;
}

// ---- one fused frame on m_stream (captured into the graph, or run directly when timing)
void VtClothSolverGPU::recordFusedFrame(Stage* t)
{
    const VtSimParams& P = simParams;
    const uint N = P.numParticles;
    FusedLaunch L{m_stream, N};
    const FrameParams* fp = m_frameParams;
    SpatialHashGPU& H = *m_spatialHash;
    const FusedOps ops = fused_ops(m_mathMode == VELVET_MATH_FAST);
    int launches = 0;

    STAGE_BEGIN(t, "Solver_SetParams");
    ops.prepare_inputs(L, m_collidersDev, m_prepared, reinterpret_cast<const float*>(attachSlotPositions.data()),
                          m_slotsDev, (uint)(3 * attachSlotPositions.size()), fp);
    launches++;
    STAGE_END(t);

    float4* cur = m_predA;
    float4* other = m_predB;
    STAGE_BEGIN(t, "Solver_Predict");  // import + pre-stabilisation + predict(0)
    ops.begin_frame(L, reinterpret_cast<const float*>(positions.data()), reinterpret_cast<const float*>(velocities.data()),
                       invMasses, m_pos4, cur, m_prepared, fp);
    launches++;
    STAGE_END(t);

    const int maxBit = (int)std::ceil(std::log2((double)H.tableSize()));
    const bool odd = RadixSorter::numPasses(maxBit) & 1;
    for (int substep = 0; substep < P.numSubsteps; substep++) {
        if (P.enableSelfCollision && substep % P.interleavedHash == 0) {
            uint* k0 = odd ? m_keysAlt.data() : H.particleHash.data();
            uint* v0 = odd ? m_valsAlt.data() : H.particleIndex.data();
            uint* k1 = odd ? H.particleHash.data() : m_keysAlt.data();
            uint* v1 = odd ? H.particleIndex.data() : m_valsAlt.data();
            STAGE_BEGIN(t, "Solver_HashParticle");
            exact_math::launch_hash_particles(L, k0, v0, cur, H.spacing(), H.tableSize() / (int)m_instancing.count, m_instancing, H.cellStart, H.tableSize());
            launches++;
            STAGE_END(t);
            STAGE_BEGIN(t, "Solver_HashSort");
            m_sorter.sort(k0, v0, k1, v1, N, maxBit, m_stream);
            launches += m_sorter.lastLaunchCount();
            STAGE_END(t);
            STAGE_BEGIN(t, "Solver_HashCache");  // + HashBuildCell: the cell table is built by the reorder launch
            VtHashParams hp = H.MakeParams(N, P.particleDiameter);
            hp.tableSize = H.tableSize() / (int)m_instancing.count;  // rows per instance
            if (const int nl = exact_math::launch_cache_neighbors_sorted(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, cur,
                                                                         m_init4, m_sorted, hp, m_instancing, nullptr, 0, H.particleHash,
                                                                         0xffffffffu, walkBandParticles())) {
                launches += nl;
            } else {
                exact_math::launch_find_cell_start(L, H.cellStart, H.cellEnd, H.particleHash);
                exact_math::launch_cache_neighbors(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, cur, m_init4, hp);
                launches += 2;
            }
            STAGE_END(t);
        }
        STAGE_BEGIN(t, "Solver_CollideParticles");  // + ApplyDeltas + CollideSDFs
        ops.collide(L, cur, other, m_pos4, H.neighbors, m_prepared, fp, P.enableSelfCollision != 0, nullptr, 0);
        launches++;
        std::swap(cur, other);
        STAGE_END(t);

        STAGE_BEGIN(t, "Solver_Iterate");  // SolveStretch + SolveAttach + SolveBending + ApplyDeltas
        if (m_gridUsable && gridIterationsPerLaunch() > 1 && P.numIterations > 1 && (uint)P.numIterations <= VT_GRID_MAX_ITERATIONS_PER_LAUNCH &&
            (size_t)m_gridDev.numTiles * m_instancing.count >= 2 * (size_t)m_gridDev.residentCtas) {
            // all iterations of the substep in one launch of the (one-wave, fully resident) grid kernel, grid barriers in between
            // (with less than a few tiles per CTA a kernel boundary is the cheaper barrier: 256^2 0.526 against 0.545 ms per frame)
            ops.iterate_grid(L, cur, other, m_gridDev, m_slotsDev, fp, m_instancing, nullptr, (unsigned)P.numIterations, m_gridBarrier.data());
            launches++;
            if (P.numIterations & 1) std::swap(cur, other);
        } else {
            for (int iteration = 0; iteration < P.numIterations; iteration++) {
                if (m_gridUsable)
                    ops.iterate_grid(L, cur, other, m_gridDev, m_slotsDev, fp, m_instancing, nullptr, 1u, nullptr);
                else
                    ops.iterate(L, cur, other, m_planDev, m_slotsDev, fp, m_instancing);
                launches++;
                std::swap(cur, other);
            }
        }
        STAGE_END(t);

        STAGE_BEGIN(t, "Solver_Finalize");  // + Predict of the next substep / export on the last one
        const bool last = substep == P.numSubsteps - 1;
        ops.end_substep(L, cur, m_pos4, other, last, reinterpret_cast<float*>(positions.data()),
                           reinterpret_cast<float*>(velocities.data()), reinterpret_cast<float*>(predicted.data()), fp);
        launches++;
        if (!last) std::swap(cur, other);
        STAGE_END(t);
    }
    STAGE_BEGIN(t, "Solver_UpdateNormals");
    ops.normals(L, m_pos4, indices, m_vtxTriOff, m_vtxTris, reinterpret_cast<float*>(normals.data()), m_instancing);
    launches++;
    STAGE_END(t);
    VT_CUDA(cudaGetLastError());
    m_graphLaunches = launches;
}

// ------------------------------------------------------------------------------------------------ domain decomposition

void VtClothSolverGPU::ddSetup(int rank, int world)
{
    if (world < 1 || rank < 0 || rank >= world) throw Error(VELVET_ERR_INVALID_ARGUMENT, "ddSetup: bad rank/world");
    if (m_instanced) throw Error(VELVET_ERR_UNSUPPORTED, "ddSetup: batched instances are sharded, not decomposed");
    VT_CUDA(cudaSetDevice(m_device));
    if (simParams.numParticles == 0) throw Error(VELVET_ERR_STATE, "ddSetup: no cloth registered");
    if (!m_squareTilesOnly) {  // strips are cut in rows of 15-particle tiles
        m_squareTilesOnly = true;
        invalidate();
    }
    ensureFusedResources();
    if (!m_fusedUsable) throw Error(VELVET_ERR_UNSUPPORTED, "ddSetup: the fused pipeline is unavailable (" + m_fallbackReason + ")");
    const uint N = simParams.numParticles;
    m_dd = ExchangePlan{};
    m_dd.rank = rank;
    m_dd.world = world;
    m_ddTilesReady = false;
    // A single grid cloth is decomposed into strips of tile rows for the peer-memory transport (contiguous particle ranges,
    // the implicit-grid Jacobi kernel with the exchange fused in); the tile-plan decomposition (ddSetupTiles) serves the
    // stepped / NCCL transport and every other mesh, and is set up only when one of those is used.
    m_ddStrip = false;
    {
        const char* e = getenv("VELVET_DD");
        const bool forceTiles = e && std::string(e) == "tiles";
        if (!forceTiles && m_gridUsable && m_gridPlan.cloths.size() == 1 && (unsigned)world <= m_gridPlan.cloths[0].tilesY) {
            const unsigned rows = m_gridPlan.cloths[0].tilesY, side = m_gridPlan.cloths[0].side;
            m_ddTileRow.assign(world + 1, 0);
            for (int r = 0; r < world; r++) m_ddTileRow[r + 1] = m_ddTileRow[r] + rows / world + ((unsigned)r < rows % world ? 1u : 0u);
            const unsigned rowFirst = m_ddTileRow[rank] * GRID_TILE;
            const unsigned rowEnd = std::min(m_ddTileRow[rank + 1] * GRID_TILE, side);
            std::vector<unsigned char> mask(N, 0);
            std::fill(mask.begin() + (size_t)rowFirst * side, mask.begin() + (size_t)rowEnd * side, (unsigned char)1);
            m_ddStripMask.upload(mask, m_stream);
            m_ddStrip = true;
        }
    }
    if (!m_ddStrip) ddSetupTiles();
    ddPeerClose();  // mappings of an earlier setup refer to buffers that may have moved
    Synchronize();
    m_ddReady = true;
}

// The tile-plan form: rank r owns a contiguous range of the Morton-ordered tiles of the record-driven plan.  Needed by the
// stepped ABI (NCCL transport, LocalShards) and by every mesh that is not a single grid cloth; a strip-decomposed cloth sets
// it up lazily, so that a 16.7 M-particle cloth over NVLink never builds the host-side tile plan at all.
void VtClothSolverGPU::ddSetupTiles()
{
    ensureTilePlan();
    const int rank = m_dd.rank, world = m_dd.world;
    if ((unsigned)world > m_plan.tiles.size()) throw Error(VELVET_ERR_INVALID_ARGUMENT, "ddSetup: more ranks than tiles");
    const uint N = simParams.numParticles;
    m_dd = build_exchange_plan(m_plan, N, rank, world);
    const unsigned T = (unsigned)m_plan.tileSize;
    m_ddOwnedBegin.assign(world, 0);
    m_ddOwnedCount.assign(world, 0);
    m_ddMaxOwned = 0;
    for (int r = 0; r < world; r++) {
        m_ddOwnedBegin[r] = m_dd.tileBeginOf[r] * T;
        const unsigned end = std::min<unsigned long long>((unsigned long long)m_dd.tileBeginOf[r + 1] * T, N);
        m_ddOwnedCount[r] = end - m_ddOwnedBegin[r];
        m_ddMaxOwned = std::max(m_ddMaxOwned, m_ddOwnedCount[r]);
    }
    std::vector<uint> sendIds, recvIds;
    m_ddSendOff.assign(world + 1, 0);
    m_ddRecvOff.assign(world + 1, 0);
    for (int q = 0; q < world; q++) {
        sendIds.insert(sendIds.end(), m_dd.sendIds[q].begin(), m_dd.sendIds[q].end());
        recvIds.insert(recvIds.end(), m_dd.recvIds[q].begin(), m_dd.recvIds[q].end());
        m_ddSendOff[q + 1] = (unsigned)sendIds.size();
        m_ddRecvOff[q + 1] = (unsigned)recvIds.size();
    }
    if (sendIds.empty()) sendIds.push_back(0);
    if (recvIds.empty()) recvIds.push_back(0);
    m_ddSendIds.upload(sendIds, m_stream);
    m_ddRecvIds.upload(recvIds, m_stream);
    m_ddSendBuf.allocate(std::max<size_t>(m_ddSendOff[world], 1));
    m_ddRecvBuf.allocate(std::max<size_t>(m_ddRecvOff[world], 1));
    {
        std::vector<unsigned char> mask(N, 0);
        for (unsigned i = 0; i < m_ddOwnedCount[rank]; i++) mask[m_plan.ownedIds[m_ddOwnedBegin[rank] + i]] = 1;
        m_ddOwnedMask.upload(mask, m_stream);
    }
    {
        std::vector<unsigned char> peerOf(sendIds.size());
        for (int q = 0; q < world; q++)
            for (unsigned i = m_ddSendOff[q]; i < m_ddSendOff[q + 1]; i++) peerOf[i] = (unsigned char)q;
        m_ddSendPeer.upload(peerOf.empty() ? std::vector<unsigned char>(1, 0) : peerOf, m_stream);
    }
    m_ddGatherSend.allocate(m_ddMaxOwned);
    m_ddGatherRecv.allocate((size_t)m_ddMaxOwned * world);
    Synchronize();
    m_ddTilesReady = true;
}

void VtClothSolverGPU::ddStripInfo(unsigned out[6]) const
{
    if (!m_ddReady || !m_ddStrip) throw Error(VELVET_ERR_STATE, "no strip decomposition");
    const unsigned side = m_gridPlan.cloths[0].side;
    const int r = m_dd.rank, w = m_dd.world;
    auto owned = [&](int q) { return (std::min(m_ddTileRow[q + 1] * GRID_TILE, side) - m_ddTileRow[q] * GRID_TILE) * side; };
    unsigned most = 0;
    for (int q = 0; q < w; q++) most = std::max(most, owned(q));
    out[0] = m_ddTileRow[r];
    out[1] = m_ddTileRow[r + 1];
    out[2] = m_ddTileRow[w];
    out[3] = owned(r);
    out[4] = most;
    out[5] = ((r > 0 ? 1u : 0u) + (r < w - 1 ? 1u : 0u)) * side;  // particles sent (= received) per iteration
}

void VtClothSolverGPU::ddEnsureTiles()
{
    if (!m_ddReady) throw Error(VELVET_ERR_STATE, "ddSetup has not been called");
    if (!m_ddTilesReady) {
        VT_CUDA(cudaSetDevice(m_device));
        ddSetupTiles();
    }
}

VtClothSolverGPU::DDBuffers VtClothSolverGPU::ddBuffers()
{
    ddEnsureTiles();
    return DDBuffers{m_ddSendBuf.data(), m_ddRecvBuf.data(), m_ddGatherSend.data(), m_ddGatherRecv.data(),
                     m_ddSendOff[m_dd.world], m_ddRecvOff[m_dd.world], m_ddOwnedCount[m_dd.rank], m_ddMaxOwned};
}

// Checks shared by every entry point that starts a decomposed frame.  The decomposed mode has no seam fallback, so what
// the fused kernels cannot take (more colliders than their shared-memory stage holds) is an error here, not a silent
// out-of-bounds read.
void VtClothSolverGPU::ddValidateFrame() const
{
    if (!m_ddReady) throw Error(VELVET_ERR_STATE, "ddSetup has not been called");
    if (m_topologyDirty) throw Error(VELVET_ERR_STATE, "the cloth changed after ddSetup: call ddSetup again");
    if (simParams.numSubsteps <= 0 || simParams.interleavedHash <= 0) throw Error(VELVET_ERR_INVALID_ARGUMENT, "bad substep parameters");
    if (sdfColliders.size() > VT_MAX_COLLIDERS)
        throw Error(VELVET_ERR_UNSUPPORTED, "a decomposed cloth takes at most 64 SDF colliders");
}

void VtClothSolverGPU::ddFrameBegin(float frameTime)
{
    ddValidateFrame();
    ddEnsureTiles();  // the stepped schedule runs the tile form
    VT_CUDA(cudaSetDevice(m_device));
    FrameParams hp;
    hp.P = simParams;
    clampNeighborBound(hp.P);
    hp.frameTime = frameTime;
    hp.substepTime = frameTime / (float)simParams.numSubsteps;
    hp.xpbdBend = simParams.bendCompliance / hp.substepTime / hp.substepTime;
    hp.numColliders = (uint)sdfColliders.size();
    VT_CUDA(cudaMemcpyAsync(m_frameParams.data(), &hp, sizeof(hp), cudaMemcpyHostToDevice, m_stream));
    const FusedOps ops = fused_ops(m_mathMode == VELVET_MATH_FAST);
    FusedLaunch L{m_stream, simParams.numParticles};
    ops.prepare_inputs(L, m_collidersDev, m_prepared, reinterpret_cast<const float*>(attachSlotPositions.data()), m_slotsDev,
                       (uint)(3 * attachSlotPositions.size()), m_frameParams);
    m_ddCur = m_predA;
    m_ddOther = m_predB;
    ops.begin_frame(L, reinterpret_cast<const float*>(positions.data()), reinterpret_cast<const float*>(velocities.data()), invMasses,
                    m_pos4, m_ddCur, m_prepared, m_frameParams);
    VT_CUDA(cudaGetLastError());
}

void VtClothSolverGPU::ddSubstepBegin(int substep)
{
    const VtSimParams& P = simParams;
    const uint N = P.numParticles;
    const FusedOps ops = fused_ops(m_mathMode == VELVET_MATH_FAST);
    FusedLaunch L{m_stream, N};
    SpatialHashGPU& H = *m_spatialHash;
    if (P.enableSelfCollision && substep % P.interleavedHash == 0) {  // replicated on every rank (for now)
        const int maxBit = (int)std::ceil(std::log2((double)H.tableSize()));
        const bool odd = RadixSorter::numPasses(maxBit) & 1;
        uint* k0 = odd ? m_keysAlt.data() : H.particleHash.data();
        uint* v0 = odd ? m_valsAlt.data() : H.particleIndex.data();
        uint* k1 = odd ? H.particleHash.data() : m_keysAlt.data();
        uint* v1 = odd ? H.particleIndex.data() : m_valsAlt.data();
        exact_math::launch_hash_particles(L, k0, v0, m_ddCur, H.spacing(), H.tableSize(), m_instancing, H.cellStart, H.tableSize());
        m_sorter.sort(k0, v0, k1, v1, N, maxBit, m_stream);
        VtHashParams hp = H.MakeParams(N, P.particleDiameter);
        // keys / sort / cell table are replicated; the expensive candidate walk only for the particles this rank owns
        if (!exact_math::launch_cache_neighbors_sorted(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, m_ddCur, m_init4,
                                                       m_sorted, hp, m_instancing, m_ddOwnedMask, m_ddOwnedCount[m_dd.rank], H.particleHash)) {
            exact_math::launch_find_cell_start(L, H.cellStart, H.cellEnd, H.particleHash);
            exact_math::launch_cache_neighbors(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, m_ddCur, m_init4, hp);
        }
    }
    // collide only the owned particles, then hand the boundary to the peers exactly like after an iteration
    ops.collide(L, m_ddCur, m_ddOther, m_pos4, H.neighbors, m_prepared, m_frameParams, P.enableSelfCollision != 0,
                m_dOwned.data() + m_ddOwnedBegin[m_dd.rank], m_ddOwnedCount[m_dd.rank]);
    exact_math::launch_gather_by_id(L, m_ddOther, m_ddSendIds, m_ddSendOff[m_dd.world], m_ddSendBuf);
    VT_CUDA(cudaGetLastError());
}

void VtClothSolverGPU::ddIterateOwned()
{
    const FusedOps ops = fused_ops(m_mathMode == VELVET_MATH_FAST);
    FusedLaunch L{m_stream, simParams.numParticles};
    TilePlanDev sub = m_planDev;  // this rank's tile range of the shared plan
    sub.tiles = m_planDev.tiles + m_dd.tileBegin;
    sub.numTiles = m_dd.tileEnd - m_dd.tileBegin;
    if (sub.numTiles) ops.iterate(L, m_ddCur, m_ddOther, sub, m_slotsDev, m_frameParams, m_instancing);
    // boundary particles this rank owns and a peer reads next iteration
    exact_math::launch_gather_by_id(L, m_ddOther, m_ddSendIds, m_ddSendOff[m_dd.world], m_ddSendBuf);
    VT_CUDA(cudaGetLastError());
}

void VtClothSolverGPU::ddIterateFinish()
{
    FusedLaunch L{m_stream, simParams.numParticles};
    exact_math::launch_scatter_by_id(L, m_ddRecvBuf, m_ddRecvIds, m_ddRecvOff[m_dd.world], m_ddOther);
    std::swap(m_ddCur, m_ddOther);
    VT_CUDA(cudaGetLastError());
}

void VtClothSolverGPU::ddGatherPack()
{
    FusedLaunch L{m_stream, simParams.numParticles};
    exact_math::launch_gather_by_id(L, m_ddCur, m_dOwned.data() + m_ddOwnedBegin[m_dd.rank], m_ddOwnedCount[m_dd.rank], m_ddGatherSend);
    VT_CUDA(cudaGetLastError());
}

void VtClothSolverGPU::ddGatherUnpack()
{
    FusedLaunch L{m_stream, simParams.numParticles};
    for (int q = 0; q < m_dd.world; q++)
        exact_math::launch_scatter_by_id(L, m_ddGatherRecv.data() + (size_t)q * m_ddMaxOwned, m_dOwned.data() + m_ddOwnedBegin[q],
                                         m_ddOwnedCount[q], m_ddCur);
    VT_CUDA(cudaGetLastError());
}

void VtClothSolverGPU::ddSubstepEnd(int substep)
{
    const FusedOps ops = fused_ops(m_mathMode == VELVET_MATH_FAST);
    FusedLaunch L{m_stream, simParams.numParticles};
    const bool last = substep == simParams.numSubsteps - 1;
    ops.end_substep(L, m_ddCur, m_pos4, m_ddOther, last, reinterpret_cast<float*>(positions.data()),
                    reinterpret_cast<float*>(velocities.data()), reinterpret_cast<float*>(predicted.data()), m_frameParams);
    if (!last) std::swap(m_ddCur, m_ddOther);
    VT_CUDA(cudaGetLastError());
}

void VtClothSolverGPU::ddFrameEnd()
{
    const FusedOps ops = fused_ops(m_mathMode == VELVET_MATH_FAST);
    FusedLaunch L{m_stream, simParams.numParticles};
    ops.normals(L, m_pos4, indices, m_vtxTriOff, m_vtxTris, reinterpret_cast<float*>(normals.data()), m_instancing);
    VT_CUDA(cudaGetLastError());
}


// ---- NVLink peer-memory transport ------------------------------------------------------------------------------------

namespace {
struct DDPeerBlob {
    cudaIpcMemHandle_t pred[2];
    cudaIpcMemHandle_t flags;
    unsigned long long particles;
    int rank, world, device, pad;
};
constexpr size_t kFlagWords = 524288;  // 2 MB: an allocation of its own
}  // namespace

size_t VtClothSolverGPU::ddPeerBlobBytes() { return sizeof(DDPeerBlob); }

void VtClothSolverGPU::ddPeerExport(void* out)
{
    if (!m_ddReady) throw Error(VELVET_ERR_STATE, "ddSetup has not been called");
    if (m_dd.world > ddpeer::kMaxWorld) throw Error(VELVET_ERR_UNSUPPORTED, "peer transport: at most 16 ranks");
    VT_CUDA(cudaSetDevice(m_device));
    ddPeerClose();
    m_ddFlags.allocate(kFlagWords);
    m_ddCtl.allocate(1);
    VT_CUDA(cudaMemsetAsync(m_ddFlags.data(), 0, m_ddFlags.bytes(), m_stream));
    VT_CUDA(cudaMemsetAsync(m_ddCtl.data(), 0, m_ddCtl.bytes(), m_stream));
    VT_CUDA(cudaStreamSynchronize(m_stream));  // zeroed before any peer can map and signal
    DDPeerBlob b{};
    VT_CUDA(cudaIpcGetMemHandle(&b.pred[0], m_predA.data()));
    VT_CUDA(cudaIpcGetMemHandle(&b.pred[1], m_predB.data()));
    VT_CUDA(cudaIpcGetMemHandle(&b.flags, m_ddFlags.data()));
    b.particles = simParams.numParticles;
    b.rank = m_dd.rank;
    b.world = m_dd.world;
    b.device = m_device;
    std::memcpy(out, &b, sizeof(b));
}

void VtClothSolverGPU::ddPeerImport(const void* blobs, size_t blobBytes)
{
    if (!m_ddReady || !m_ddFlags.data()) throw Error(VELVET_ERR_STATE, "ddPeerImport: call ddSetup and ddPeerExport first");
    if (blobBytes != sizeof(DDPeerBlob)) throw Error(VELVET_ERR_INVALID_ARGUMENT, "ddPeerImport: blob size mismatch");
    VT_CUDA(cudaSetDevice(m_device));
    ddpeer::PeerTable T{};
    T.rank = m_dd.rank;
    T.world = m_dd.world;
    std::vector<void*> opened;
    auto open = [&](const cudaIpcMemHandle_t& h) {
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            for (void* o : opened) cudaIpcCloseMemHandle(o);
            throw Error(VELVET_ERR_CUDA, std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e));
        }
        opened.push_back(p);
        return p;
    };
    for (int q = 0; q < m_dd.world; q++) {
        if (q == m_dd.rank) {
            T.pred[0][q] = m_predA.data();
            T.pred[1][q] = m_predB.data();
            T.flags[q] = m_ddFlags.data();
            continue;
        }
        DDPeerBlob b;
        std::memcpy(&b, (const char*)blobs + (size_t)q * blobBytes, sizeof(b));
        if (b.rank != q || b.world != m_dd.world || b.particles != simParams.numParticles) {
            for (void* o : opened) cudaIpcCloseMemHandle(o);
            throw Error(VELVET_ERR_INVALID_ARGUMENT, "ddPeerImport: blob " + std::to_string(q) + " does not describe rank " + std::to_string(q) +
                                                         " of the same cloth");
        }
        T.pred[0][q] = (float4*)open(b.pred[0]);
        T.pred[1][q] = (float4*)open(b.pred[1]);
        T.flags[q] = (unsigned*)open(b.flags);
    }
    m_ddPeers = T;
    m_ddOpened = std::move(opened);
    m_ddPeersReady = true;
    m_ddGraphKey = 0;  // pointers are baked into the graph
}

void VtClothSolverGPU::ddPeerClose()
{
    if (m_stream) cudaStreamSynchronize(m_stream);
    if (m_ddGraphExec) cudaGraphExecDestroy(m_ddGraphExec);
    if (m_ddGraph) cudaGraphDestroy(m_ddGraph);
    m_ddGraphExec = nullptr;
    m_ddGraph = nullptr;
    m_ddGraphKey = 0;
    for (void* p : m_ddOpened) cudaIpcCloseMemHandle(p);
    m_ddOpened.clear();
    m_ddPeersReady = false;
}

bool VtClothSolverGPU::ddPeerError()
{
    if (!m_ddCtl.data()) return false;
    ddpeer::Control c{};
    VT_CUDA(cudaMemcpyAsync(&c, m_ddCtl.data(), sizeof(c), cudaMemcpyDeviceToHost, m_stream));
    VT_CUDA(cudaStreamSynchronize(m_stream));
    return c.error != 0;
}

// One decomposed frame on m_stream.  Buffer roles: index 0 = predA, 1 = predB on every rank (all ranks swap in lock step).
void VtClothSolverGPU::recordDDFrame()
{
    const VtSimParams& P = simParams;
    const uint N = P.numParticles;
    if (m_ddStrip) {
        recordDDStripFrame(nullptr);
        return;
    }
    const FusedOps ops = fused_ops(m_mathMode == VELVET_MATH_FAST);
    FusedLaunch L{m_stream, N};
    SpatialHashGPU& H = *m_spatialHash;
    const FrameParams* fp = m_frameParams;
    const ddpeer::PeerTable* T = &m_ddPeers;
    ddpeer::Control* ctl = m_ddCtl.data();
    const unsigned long long timeoutNs = 20ull * 1000000000ull;
    float4* buf[2] = {m_predA.data(), m_predB.data()};
    int cur = 0, other = 1;
    int launches = 0;
    const int rank = m_dd.rank, world = m_dd.world;
    const uint* ownedIds = m_dOwned.data() + m_ddOwnedBegin[rank];
    const uint ownedCount = m_ddOwnedCount[rank];
    const uint sendTotal = m_ddSendOff[world];
    auto wait = [&] {
        ddpeer::launch_wait(m_stream, T, ctl, m_ddFlags.data(), timeoutNs);
        launches++;
    };
    auto push_halo = [&](int which) {  // boundary particles this rank owns and a peer reads next
        ddpeer::launch_push_halo(m_stream, T, ctl, which, buf[which], m_ddSendIds, m_ddSendPeer, sendTotal);
        launches++;
    };

    ops.prepare_inputs(L, m_collidersDev, m_prepared, reinterpret_cast<const float*>(attachSlotPositions.data()), m_slotsDev,
                       (uint)(3 * attachSlotPositions.size()), fp);
    ops.begin_frame(L, reinterpret_cast<const float*>(positions.data()), reinterpret_cast<const float*>(velocities.data()), invMasses,
                    m_pos4, buf[cur], m_prepared, fp);
    launches += 2;
    TilePlanDev sub = m_planDev;  // this rank's tile range of the shared plan
    sub.tiles = m_planDev.tiles + m_dd.tileBegin;
    sub.numTiles = m_dd.tileEnd - m_dd.tileBegin;
    const int maxBit = (int)std::ceil(std::log2((double)H.tableSize()));
    const bool odd = RadixSorter::numPasses(maxBit) & 1;
    for (int substep = 0; substep < P.numSubsteps; substep++) {
        // "I no longer read the buffer the peers are about to push into" (previous end_substep / begin_frame are done)
        ddpeer::launch_signal(m_stream, T, ctl);
        launches++;
        if (P.enableSelfCollision && substep % P.interleavedHash == 0) {
            uint* k0 = odd ? m_keysAlt.data() : H.particleHash.data();
            uint* v0 = odd ? m_valsAlt.data() : H.particleIndex.data();
            uint* k1 = odd ? H.particleHash.data() : m_keysAlt.data();
            uint* v1 = odd ? H.particleIndex.data() : m_valsAlt.data();
            exact_math::launch_hash_particles(L, k0, v0, buf[cur], H.spacing(), H.tableSize(), m_instancing, H.cellStart, H.tableSize());
            m_sorter.sort(k0, v0, k1, v1, N, maxBit, m_stream);
            launches += 1 + m_sorter.lastLaunchCount();
            VtHashParams hp = H.MakeParams(N, P.particleDiameter);
            if (const int nl = exact_math::launch_cache_neighbors_sorted(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, buf[cur],
                                                                         m_init4, m_sorted, hp, m_instancing, m_ddOwnedMask, ownedCount,
                                                                         H.particleHash)) {
                launches += nl;
            } else {
                exact_math::launch_find_cell_start(L, H.cellStart, H.cellEnd, H.particleHash);
                exact_math::launch_cache_neighbors(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, buf[cur], m_init4, hp);
                launches += 2;
            }
        }
        ops.collide(L, buf[cur], buf[other], m_pos4, H.neighbors, m_prepared, fp, P.enableSelfCollision != 0, ownedIds, ownedCount);
        launches++;
        wait();  // every peer has signalled: their `other` is free to be written
        push_halo(other);
        wait();
        std::swap(cur, other);
        for (int iteration = 0; iteration < P.numIterations; iteration++) {
            if (sub.numTiles) {
                ops.iterate(L, buf[cur], buf[other], sub, m_slotsDev, fp, m_instancing);
                launches++;
            }
            push_halo(other);
            wait();
            std::swap(cur, other);
        }
        // all-gather of the substep result: every owned particle into every peer's array at the same index
        ddpeer::launch_push_owned(m_stream, T, ctl, cur, buf[cur], ownedIds, ownedCount);
        launches++;
        wait();
        const bool last = substep == P.numSubsteps - 1;
        ops.end_substep(L, buf[cur], m_pos4, buf[other], last, reinterpret_cast<float*>(positions.data()),
                        reinterpret_cast<float*>(velocities.data()), reinterpret_cast<float*>(predicted.data()), fp);
        launches++;
        if (!last) std::swap(cur, other);
    }
    ops.normals(L, m_pos4, indices, m_vtxTriOff, m_vtxTris, reinterpret_cast<float*>(normals.data()), m_instancing);
    launches++;
    VT_CUDA(cudaGetLastError());
    m_ddGraphLaunches = launches;
}

// One frame of a strip-decomposed grid cloth (ddSetup: m_ddStrip).  Rank r owns the particle rows of its tile rows, i.e. the
// contiguous index range [begin, begin + count).  Per substep:
//   hash keys / sort / cell table on every rank (replicated: bit-identical lists need the global order); the candidate walk
//   and collide only for the owned range; the two outermost owned rows to the neighbours;
//   numIterations x iterate_grid_kernel over the owned tile rows -- the per-iteration row exchange happens INSIDE the kernel
//   (boundary tiles first, peer stores from the epilogue, publish from the last boundary tile; the next launch waits for the
//   neighbours' rows when it starts): no exchange launches between iterations;
//   all-gather of the owned range (peer stores), then Finalize + Predict on every rank for every particle, which keeps
//   pos4 / the public buffers identical everywhere (collide reads pos4 of arbitrary neighbours).
// The signal / wait pairs are the "I have stopped reading the buffer you are about to write" handshakes.
void VtClothSolverGPU::recordDDStripFrame(Stage* t)
{
    const VtSimParams& P = simParams;
    const uint N = P.numParticles;
    const FusedOps ops = fused_ops(m_mathMode == VELVET_MATH_FAST);
    FusedLaunch L{m_stream, N};
    SpatialHashGPU& H = *m_spatialHash;
    const FrameParams* fp = m_frameParams;
    const ddpeer::PeerTable* T = &m_ddPeers;
    ddpeer::Control* ctl = m_ddCtl.data();
    const unsigned long long timeoutNs = 20ull * 1000000000ull;
    const int rank = m_dd.rank, world = m_dd.world;
    const unsigned side = m_gridPlan.cloths[0].side;
    ddpeer::StripArgs A{};
    A.T = m_ddPeers;
    A.ctl = ctl;
    A.localFlags = m_ddFlags.data();
    A.timeoutNs = timeoutNs;
    A.up = rank > 0 ? rank - 1 : -1;
    A.down = rank < world - 1 ? rank + 1 : -1;
    A.tileRowBegin = m_ddTileRow[rank];
    A.tileRowEnd = m_ddTileRow[rank + 1];
    A.rowFirst = A.tileRowBegin * GRID_TILE;
    A.rowLast = std::min(A.tileRowEnd * GRID_TILE, side) - 1;
    const unsigned begin = A.rowFirst * side, count = (A.rowLast + 1 - A.rowFirst) * side;
    float4* buf[2] = {m_predA.data(), m_predB.data()};
    int cur = 0, other = 1;
    int launches = 0;
    auto wait_all = [&] {
        ddpeer::launch_strip_wait_all(m_stream, T, ctl, m_ddFlags.data(), timeoutNs);
        launches++;
    };
    auto signal = [&] {
        ddpeer::launch_strip_signal(m_stream, T, ctl);
        launches++;
    };

    STAGE_BEGIN(t, "DD_BeginFrame(replicated)");
    ops.prepare_inputs(L, m_collidersDev, m_prepared, reinterpret_cast<const float*>(attachSlotPositions.data()), m_slotsDev,
                       (uint)(3 * attachSlotPositions.size()), fp);
    ops.begin_frame(L, reinterpret_cast<const float*>(positions.data()), reinterpret_cast<const float*>(velocities.data()), invMasses,
                    m_pos4, buf[cur], m_prepared, fp);
    launches += 2;
    STAGE_END(t);
    const int maxBit = (int)std::ceil(std::log2((double)H.tableSize()));
    const bool odd = RadixSorter::numPasses(maxBit) & 1;
    for (int substep = 0; substep < P.numSubsteps; substep++) {
        STAGE_BEGIN(t, "DD_Sync");
        signal();  // begin_frame / end_substep are done with `other` (and with `cur` as an output): the neighbours may write rows
        STAGE_END(t);
        if (P.enableSelfCollision && substep % P.interleavedHash == 0) {
            STAGE_BEGIN(t, "DD_HashKeysSortCells(replicated)");
            uint* k0 = odd ? m_keysAlt.data() : H.particleHash.data();
            uint* v0 = odd ? m_valsAlt.data() : H.particleIndex.data();
            uint* k1 = odd ? H.particleHash.data() : m_keysAlt.data();
            uint* v1 = odd ? H.particleIndex.data() : m_valsAlt.data();
            exact_math::launch_hash_particles(L, k0, v0, buf[cur], H.spacing(), H.tableSize(), m_instancing, H.cellStart, H.tableSize());
            m_sorter.sort(k0, v0, k1, v1, N, maxBit, m_stream);
            launches += 1 + m_sorter.lastLaunchCount();
            STAGE_END(t);
            STAGE_BEGIN(t, "DD_NeighborCache(reorder replicated, walk owned)");
            VtHashParams hp = H.MakeParams(N, P.particleDiameter);
            if (const int nl = exact_math::launch_cache_neighbors_sorted(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, buf[cur],
                                                                         m_init4, m_sorted, hp, m_instancing, m_ddStripMask, count,
                                                                         H.particleHash, begin, walkBandParticles())) {
                launches += nl;
            } else {
                exact_math::launch_find_cell_start(L, H.cellStart, H.cellEnd, H.particleHash);
                exact_math::launch_cache_neighbors(L, H.neighbors, H.particleIndex, H.cellStart, H.cellEnd, buf[cur], m_init4, hp);
                launches += 2;
            }
            STAGE_END(t);
        }
        STAGE_BEGIN(t, "DD_Collide(owned)");
        ops.collide_range(L, buf[cur], buf[other], m_pos4, H.neighbors, m_prepared, fp, P.enableSelfCollision != 0, begin, count);
        launches++;
        STAGE_END(t);
        STAGE_BEGIN(t, "DD_Sync");
        wait_all();  // every peer is past its end_substep: its `other` may take rows
        A.which = other;
        ddpeer::launch_strip_push_rows(m_stream, A, buf[other], side);
        launches++;
        STAGE_END(t);
        std::swap(cur, other);
        STAGE_BEGIN(t, "DD_Iterate(owned, exchange fused)");
        for (int iteration = 0; iteration < P.numIterations; iteration++) {
            A.which = other;
            // The last iteration stores every result into every peer: the per-substep all-gather overlaps the Jacobi math.
            // Nobody reads those rows of that buffer meanwhile: iterations only read the rows next to their own strip, and a
            // neighbour has published (i.e. finished reading them) before this launch passes its wait.
            A.gatherAll = iteration == P.numIterations - 1 ? 1 : 0;
            ops.iterate_grid(L, buf[cur], buf[other], m_gridDev, m_slotsDev, fp, m_instancing, &A, 1u, nullptr);
            launches++;
            std::swap(cur, other);
        }
        A.gatherAll = 0;
        STAGE_END(t);
        STAGE_BEGIN(t, "DD_Sync");
        if (P.numIterations <= 0) {  // nothing to ride on: gather the collide output with a launch of its own
            signal();
            wait_all();
            ddpeer::launch_strip_push_range(m_stream, T, ctl, cur, buf[cur], begin, count);
            launches++;
        } else {
            signal();  // stream order: the whole last iteration, all its peer stores included, is behind this
        }
        wait_all();
        STAGE_END(t);
        STAGE_BEGIN(t, "DD_Finalize(replicated)");
        const bool last = substep == P.numSubsteps - 1;
        ops.end_substep(L, buf[cur], m_pos4, buf[other], last, reinterpret_cast<float*>(positions.data()),
                        reinterpret_cast<float*>(velocities.data()), reinterpret_cast<float*>(predicted.data()), fp);
        launches++;
        STAGE_END(t);
        if (!last) std::swap(cur, other);
    }
    STAGE_BEGIN(t, "DD_Normals(replicated)");
    ops.normals(L, m_pos4, indices, m_vtxTriOff, m_vtxTris, reinterpret_cast<float*>(normals.data()), m_instancing);
    launches++;
    STAGE_END(t);
    VT_CUDA(cudaGetLastError());
    m_ddGraphLaunches = launches;
}

void VtClothSolverGPU::ddSimulate(float frameTime)
{
    ddValidateFrame();
    if (!m_ddPeersReady) throw Error(VELVET_ERR_STATE, "ddSimulate: peers are not mapped (ddPeerExport / ddPeerImport)");
    VT_CUDA(cudaSetDevice(m_device));
    FrameParams hp;
    hp.P = simParams;
    clampNeighborBound(hp.P);
    hp.frameTime = frameTime;
    hp.substepTime = frameTime / (float)simParams.numSubsteps;
    hp.xpbdBend = simParams.bendCompliance / hp.substepTime / hp.substepTime;
    hp.numColliders = (uint)sdfColliders.size();
    VT_CUDA(cudaMemcpyAsync(m_frameParams.data(), &hp, sizeof(hp), cudaMemcpyHostToDevice, m_stream));
    if (m_ddStrip && getenv("VELVET_DD_TIMING")) {  // diagnostic: the frame un-graphed with an event pair per stage (every rank alike)
        Stage timing(m_stream);
        recordDDStripFrame(&timing);
        const StageTiming r = timing.collect();
        if (m_dd.rank == 0) {
            std::string line = "[velvet dd timing, rank 0, ms]";
            for (size_t i = 0; i < r.labels.size(); i++) line += " " + r.labels[i] + "=" + std::to_string(r.ms[i]);
            fprintf(stderr, "%s\n", line.c_str());
        }
        m_lastLaunches = m_ddGraphLaunches;
        return;
    }
    const unsigned long long key = topologyKey() | 1ull;
    if (!m_ddGraphExec || key != m_ddGraphKey) {
        if (m_ddGraphExec) cudaGraphExecDestroy(m_ddGraphExec);
        if (m_ddGraph) cudaGraphDestroy(m_ddGraph);
        m_ddGraphExec = nullptr;
        m_ddGraph = nullptr;
        VT_CUDA(cudaStreamBeginCapture(m_stream, cudaStreamCaptureModeThreadLocal));
        try {
            recordDDFrame();
        } catch (...) {
            cudaGraph_t broken = nullptr;
            cudaStreamEndCapture(m_stream, &broken);
            if (broken) cudaGraphDestroy(broken);
            throw;
        }
        VT_CUDA(cudaStreamEndCapture(m_stream, &m_ddGraph));
        VT_CUDA(cudaGraphInstantiate(&m_ddGraphExec, m_ddGraph, 0));
        m_ddGraphKey = key;
    }
    VT_CUDA(cudaGraphLaunch(m_ddGraphExec, m_stream));
    m_mayBeBusy = true;
    m_lastLaunches = m_ddGraphLaunches;
}

void VtClothSolverGPU::Simulate() { Simulate(kFixedDeltaTime); }

void VtClothSolverGPU::Simulate(float frameTime)
{
    VT_CUDA(cudaSetDevice(m_device));
    m_mayBeBusy = true;  // until the next Synchronize()
    if (simParams.numSubsteps <= 0) throw Error(VELVET_ERR_INVALID_ARGUMENT, "numSubsteps must be positive");
    if (simParams.interleavedHash <= 0) throw Error(VELVET_ERR_INVALID_ARGUMENT, "interleavedHash must be positive");
    if (simParams.numParticles == 0) {
        m_lastLaunches = 0;
        return;
    }
    bool fused = (m_pipeline == VELVET_PIPELINE_FUSED);
    if (fused) {
        ensureFusedResources();
        if (!m_fusedUsable || sdfColliders.size() > VT_MAX_COLLIDERS) fused = false;
    }
    if (!fused) {
        if (m_instanced)
            throw Error(VELVET_ERR_UNSUPPORTED, "batched instances need the fused pipeline (" +
                                                    (m_fallbackReason.empty() ? std::string("seam pipeline selected") : m_fallbackReason) + ")");
        simulateSeam(frameTime, nullptr);
        return;
    }

    FrameParams hp;
    hp.P = simParams;
    clampNeighborBound(hp.P);
    hp.frameTime = frameTime;
    hp.substepTime = frameTime / (float)simParams.numSubsteps;
    hp.xpbdBend = simParams.bendCompliance / hp.substepTime / hp.substepTime;
    hp.numColliders = (uint)sdfColliders.size();
    VT_CUDA(cudaMemcpyAsync(m_frameParams.data(), &hp, sizeof(hp), cudaMemcpyHostToDevice, m_stream));

    const unsigned long long key = topologyKey();
    if (!m_graphExec || key != m_graphKey) {
        if (m_graphExec) {
            cudaGraphExecDestroy(m_graphExec);
            m_graphExec = nullptr;
        }
        if (m_graph) {
            cudaGraphDestroy(m_graph);
            m_graph = nullptr;
        }
        SetupClock clk;
        VT_CUDA(cudaStreamBeginCapture(m_stream, cudaStreamCaptureModeThreadLocal));
        try {
            recordFusedFrame(nullptr);
        } catch (...) {
            cudaGraph_t broken = nullptr;
            cudaStreamEndCapture(m_stream, &broken);
            if (broken) cudaGraphDestroy(broken);
            throw;
        }
        VT_CUDA(cudaStreamEndCapture(m_stream, &m_graph));
        VT_CUDA(cudaGraphInstantiate(&m_graphExec, m_graph, 0));
        m_graphKey = key;
        clk.lap("Simulate: graph capture + instantiate");
    }
    VT_CUDA(cudaGraphLaunch(m_graphExec, m_stream));
    m_mayBeBusy = true;
    m_lastLaunches = m_graphLaunches;
}

StageTiming VtClothSolverGPU::SimulateTimed()
{
    VT_CUDA(cudaSetDevice(m_device));
    Stage timing(m_stream);
    if (simParams.numParticles == 0) return StageTiming{};
    bool fused = (m_pipeline == VELVET_PIPELINE_FUSED);
    if (fused) {
        ensureFusedResources();
        if (!m_fusedUsable || sdfColliders.size() > VT_MAX_COLLIDERS) fused = false;
    }
    if (!fused && m_instanced) throw Error(VELVET_ERR_UNSUPPORTED, "batched instances need the fused pipeline");
    timing.begin("Solver_Total");
    if (!fused) {
        simulateSeam(kFixedDeltaTime, &timing);
    } else {
        FrameParams hp;
        hp.P = simParams;
        clampNeighborBound(hp.P);
    clampNeighborBound(hp.P);
        hp.frameTime = kFixedDeltaTime;
        hp.substepTime = kFixedDeltaTime / (float)simParams.numSubsteps;
        hp.xpbdBend = simParams.bendCompliance / hp.substepTime / hp.substepTime;
        hp.numColliders = (uint)sdfColliders.size();
        VT_CUDA(cudaMemcpyAsync(m_frameParams.data(), &hp, sizeof(hp), cudaMemcpyHostToDevice, m_stream));
        recordFusedFrame(&timing);
        m_lastLaunches = m_graphLaunches;
    }
    // close Solver_Total (it is the first span)
    VT_CUDA(cudaEventRecord(timing.spans.front().b, m_stream));
    return timing.collect();
}

// ------------------------------------------------------------------------------------------------ inputs

// Scene.hpp L131-168
void GenerateClothMesh(int resolution, float* vertices, uint* meshIndices)
{
    const float clothSize = 2.0f;
    const float r = (float)resolution;
    float* v = vertices;
    for (int y = 0; y <= resolution; y++)
        for (int x = 0; x <= resolution; x++) {
            *v++ = clothSize * ((float)x / r - 0.5f);
            *v++ = clothSize * (-(float)y / r);
            *v++ = clothSize * 0.0f;
        }
    const uint side = (uint)resolution + 1;
    auto at = [side](uint x, uint y) { return x * side + y; };
    uint* o = meshIndices;
    for (uint x = 0; x < (uint)resolution; x++)
        for (uint y = 0; y < (uint)resolution; y++) {
            const uint quad[6] = {at(x, y), at(x + 1, y), at(x, y + 1), at(x, y + 1), at(x + 1, y), at(x + 1, y + 1)};
            std::memcpy(o, quad, sizeof(quad));
            o += 6;
        }
}

namespace {

struct Mat4 {
    float c[4][4];  // c[col][row]
};

Mat4 identity4()
{
    Mat4 m{};
    for (int i = 0; i < 4; i++) m.c[i][i] = 1.0f;
    return m;
}

// glm::rotate(m, angle, unit axis): rotation block first, then m * R column by column (left-to-right sums)
Mat4 rotated(const Mat4& m, float angle, float ax, float ay, float az)
{
    const float c = std::cos(angle), s = std::sin(angle);
    const float axis[3] = {ax, ay, az};
    const float temp[3] = {(1.0f - c) * ax, (1.0f - c) * ay, (1.0f - c) * az};
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = temp[0] * axis[1] + s * axis[2];
    R[0][2] = temp[0] * axis[2] - s * axis[1];
    R[1][0] = temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = temp[1] * axis[2] + s * axis[0];
    R[2][0] = temp[2] * axis[0] + s * axis[1];
    R[2][1] = temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    Mat4 out = m;
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 4; row++)
            out.c[col][row] = (m.c[0][row] * R[col][0] + m.c[1][row] * R[col][1]) + m.c[2][row] * R[col][2];
    return out;
}

}  // namespace

// Transform::matrix(): translate, RotateWithDegree (y, z, x), scale -- Transform.hpp L22-29, Helper.cpp L8-15
void TransformMatrix(const float* position3, const float* rotationDeg3, const float* scale3, float* out16)
{
    Mat4 m = identity4();
    for (int row = 0; row < 4; row++)  // glm::translate: m[3] = m[0]*v.x + m[1]*v.y + m[2]*v.z + m[3]
        m.c[3][row] = ((m.c[0][row] * position3[0] + m.c[1][row] * position3[1]) + m.c[2][row] * position3[2]) + m.c[3][row];
    const float toRad = 0.01745329251994329576923690768489f;  // glm::radians
    m = rotated(m, rotationDeg3[1] * toRad, 0, 1, 0);
    m = rotated(m, rotationDeg3[2] * toRad, 0, 0, 1);
    m = rotated(m, rotationDeg3[0] * toRad, 1, 0, 0);
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 4; row++) m.c[col][row] *= scale3[col];
    std::memcpy(out16, m.c, sizeof(float) * 16);
}

// glm::inverse(mat4): the cofactor expansion published in glm/detail/func_matrix.inl
void Mat4Inverse(const float* in16, float* out16)
{
    float m[4][4];
    std::memcpy(m, in16, sizeof(m));  // m[col][row]
    const float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    const float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    const float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    const float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    const float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    const float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    const float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    const float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    const float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    const float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    const float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    const float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    const float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    const float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    const float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    const float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    const float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];

    const float Fac0[4] = {Coef00, Coef00, Coef02, Coef03};
    const float Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
    const float Fac2[4] = {Coef08, Coef08, Coef10, Coef11};
    const float Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
    const float Fac4[4] = {Coef16, Coef16, Coef18, Coef19};
    const float Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
    const float Vec0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]};
    const float Vec1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    const float Vec2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]};
    const float Vec3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    const float SignA[4] = {+1, -1, +1, -1};
    const float SignB[4] = {-1, +1, -1, +1};
    float inv[4][4];
    for (int k = 0; k < 4; k++) {
        inv[0][k] = ((Vec1[k] * Fac0[k] - Vec2[k] * Fac1[k]) + Vec3[k] * Fac2[k]) * SignA[k];
        inv[1][k] = ((Vec0[k] * Fac0[k] - Vec2[k] * Fac3[k]) + Vec3[k] * Fac4[k]) * SignB[k];
        inv[2][k] = ((Vec0[k] * Fac1[k] - Vec1[k] * Fac3[k]) + Vec3[k] * Fac5[k]) * SignA[k];
        inv[3][k] = ((Vec0[k] * Fac2[k] - Vec1[k] * Fac4[k]) + Vec2[k] * Fac5[k]) * SignB[k];
    }
    const float Dot0[4] = {m[0][0] * inv[0][0], m[0][1] * inv[1][0], m[0][2] * inv[2][0], m[0][3] * inv[3][0]};
    const float Dot1 = (Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]);
    const float OneOverDeterminant = 1.0f / Dot1;
    for (int col = 0; col < 4; col++)
        for (int row = 0; row < 4; row++) out16[4 * col + row] = inv[col][row] * OneOverDeterminant;
}

// UpdateColliders body for one collider, VtClothSolverGPU.hpp L195-203
void MakeCollider(int type, const float* position3, const float* scale3, const float* cur16, const float* last16,
                  float deltaTime, VtSDFCollider* out)
{
    std::memset(out, 0, sizeof(*out));
    out->type = type;
    std::memcpy(out->position, position3, 12);
    std::memcpy(out->scale, scale3, 12);
    out->deltaTime = deltaTime;
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 3; row++) out->curTransform[3 * col + row] = cur16[4 * col + row];
    Mat4Inverse(cur16, out->invCurTransform);
    std::memcpy(out->lastTransform, last16, 64);
}

// ------------------------------------------------------------------------------------------------ VtClothObjectGPU

GridConstraints GenerateGridConstraints(int R, const float* vertices, const uint* meshIndices, const float* M,
                                        const std::vector<int>& attachedIndices, float particleDiameterScalar, int off)
{
    GridConstraints g;
    const int side = R + 1;
    const size_t nv = (size_t)side * side;
    const size_t ni = (size_t)6 * R * R;
    auto vtx = [&](size_t i) { return load3(vertices, i); };
    g.particleDiameter = length_plain(vtx(0) - vtx(1)) * particleDiameterScalar;  // VtClothObjectGPU.hpp L49

    // ApplyTransform (L67-73): host copy of the world-space positions, used for rest lengths only
    std::vector<vec3> world(nv);
    for (size_t i = 0; i < nv; i++) world[i] = mul_point(M, vtx(i), 1.0f);
    auto dist = [&](int a, int b) { return length_plain(world[(size_t)a] - world[(size_t)b]); };
    auto at = [side](int x, int y) { return x * side + y; };

    // GenerateStretch (L75-116): structural (y, x) then the two shear diagonals, in that emission order
    g.stretchIdx.reserve(2 * (size_t)(4 * R * R + 2 * R));
    g.stretchLen.reserve((size_t)(4 * R * R + 2 * R));
    auto emit = [&](int a, int b) {
        g.stretchIdx.push_back(off + a);
        g.stretchIdx.push_back(off + b);
        g.stretchLen.push_back(dist(a, b));
    };
    for (int x = 0; x < side; x++)
        for (int y = 0; y < side; y++) {
            if (y != R) emit(at(x, y), at(x, y + 1));
            if (x != R) emit(at(x, y), at(x + 1, y));
            if (y != R && x != R) {
                emit(at(x, y), at(x + 1, y + 1));
                emit(at(x, y + 1), at(x + 1, y));
            }
        }

    // GenerateAttach (L134-148): every particle gets a long-range attachment to every slot
    for (size_t slot = 0; slot < attachedIndices.size(); slot++) {
        const vec3 slotPos = world[(size_t)attachedIndices[slot]];
        g.slotPositions.push_back(slotPos.x);
        g.slotPositions.push_back(slotPos.y);
        g.slotPositions.push_back(slotPos.z);
        for (size_t i = 0; i < nv; i++) {
            g.attachPid.push_back(off + (int)i);
            g.attachSlot.push_back((int)slot);
            g.attachDist.push_back(length_plain(slotPos - world[i]));
        }
    }

    // GenerateBending (L118-132): one dihedral per quad, indices (i, i+5, i+2, i+1), rest angle 0
    g.bendIdx.reserve(4 * (size_t)R * R);
    g.bendAngle.reserve((size_t)R * R);
    for (size_t i = 0; i + 5 < ni; i += 6) {
        g.bendIdx.push_back((uint)off + meshIndices[i]);
        g.bendIdx.push_back((uint)off + meshIndices[i + 5]);
        g.bendIdx.push_back((uint)off + meshIndices[i + 2]);
        g.bendIdx.push_back((uint)off + meshIndices[i + 1]);
        g.bendAngle.push_back(0.0f);
    }
    return g;
}

void VtClothObjectGPU::Start(const float* vertices, const uint* meshIndices, const float* M)
{
    const int R = m_resolution;
    const size_t nv = (size_t)(R + 1) * (R + 1);
    const size_t ni = (size_t)6 * R * R;
    m_particleDiameter = length_plain(load3(vertices, 0) - load3(vertices, 1)) * m_solver->simParams.particleDiameterScalar;  // L49
    m_indexOffset = m_solver->AddCloth(vertices, (int)nv, meshIndices, (int)ni, M, m_particleDiameter);
    if (VtClothSolverGPU::deviceRegistration()) {
        m_solver->GenerateGridClothOnDevice(R, m_indexOffset, vertices, M, m_attachedIndices);
        return;
    }
    SetupClock clk;
    const GridConstraints g = GenerateGridConstraints(R, vertices, meshIndices, M, m_attachedIndices,
                                                      m_solver->simParams.particleDiameterScalar, m_indexOffset);
    clk.lap("Start: GenerateGridConstraints (host)");
    m_solver->AddStretchBulk(g.stretchIdx.data(), g.stretchLen.data(), g.stretchLen.size());
    // AddAttachSlot then that slot's AddAttach calls, slot by slot (L134-148)
    for (size_t slot = 0; slot < m_attachedIndices.size(); slot++) {
        m_solver->AddAttachSlot(&g.slotPositions[3 * slot]);
        m_solver->AddAttachBulk(&g.attachPid[slot * nv], &g.attachSlot[slot * nv], &g.attachDist[slot * nv], nv);
    }
    m_solver->AddBendBulk(g.bendIdx.data(), g.bendAngle.data(), g.bendAngle.size());
    clk.lap("Start: bulk registration");
}

}  // namespace velvet
