// ref_driver.cu -- C entry points around the REFERENCE's own kernels (O3).  TEST INFRASTRUCTURE ONLY.
//
// This file is ours; everything it calls (Velvet::PredictPositions, SolveStretch, ..., HashObjects, VtBuffer,
// Timer/ScopedTimerGPU) is the unmodified reference source compiled next to it (see build_ref_cuda.sh).  The host
// class VtClothSolverGPU.hpp cannot be compiled headless (Component/Mesh/GL/ImGui), so its orchestration is restated
// here line by line: AddCloth's buffer setup (hpp L114-156), SpatialHashGPU's constructor / SetInitialPositions / Hash
// (SpatialHashGPU.hpp L18-52) and Simulate (hpp L56-111).
#include "VtClothSolverGPU.cuh"
#include "SpatialHashGPU.cuh"
#include "Timer.hpp"
#include "VtBuffer.hpp"

namespace Velvet {
Timer* Timer::s_timer = nullptr;  // Timer.cpp
}

using namespace Velvet;

namespace {

struct RefSolver {
    VtSimParams P;
    VtBuffer<glm::vec3> positions, normals, velocities, predicted, deltas;
    VtBuffer<uint> indices;
    VtBuffer<int> deltaCounts;
    VtBuffer<float> invMasses;
    VtBuffer<int> stretchIndices;
    VtBuffer<float> stretchLengths;
    VtBuffer<uint> bendIndices;
    VtBuffer<float> bendAngles;
    VtBuffer<int> attachParticleIDs, attachSlotIDs;
    VtBuffer<float> attachDistances;
    VtBuffer<glm::vec3> attachSlotPositions;
    VtBuffer<SDFCollider> sdfColliders;
    // SpatialHashGPU members
    VtBuffer<uint> neighbors, particleHash, particleIndex, cellStart, cellEnd;
    VtBuffer<glm::vec3> initialPositions;
    float spacing = 0;
    int tableSize = 0;
};

Timer* g_timer = nullptr;

const char* kLabels[] = {"Solver_Total", "Solver_SetParams", "Solver_Initialize", "Solver_Predict", "Solver_SolveStretch",
                         "Solver_SolveBending", "Solver_SolveAttach", "Solver_ApplyDeltas", "Solver_CollideSDFs",
                         "Solver_CollideParticles", "Solver_Finalize", "Solver_UpdateNormals", "Solver_HashParticle",
                         "Solver_HashSort", "Solver_HashBuildCell", "Solver_HashCache"};
double g_lastMs[16];

void run_hash(RefSolver* s, const VtBuffer<glm::vec3>& positions)
{  // SpatialHashGPU::Hash, SpatialHashGPU.hpp L41-52
    HashParams params;
    params.numObjects = (uint)positions.size();
    params.cellSpacing = s->spacing;
    params.cellSpacing2 = s->spacing * s->spacing;
    params.tableSize = s->tableSize;
    params.maxNumNeighbors = s->P.maxNumNeighbors;
    params.particleDiameter2 = s->P.particleDiameter * s->P.particleDiameter;
    HashObjects(s->particleHash, s->particleIndex, s->cellStart, s->cellEnd, s->neighbors, positions, s->initialPositions, params);
}

}  // namespace

extern "C" {

void* refcuda_create(const VtSimParams* params)
{
    if (!g_timer) g_timer = new Timer();
    RefSolver* s = new RefSolver();
    s->P = *params;
    s->P.numParticles = 0;
    return s;
}

void refcuda_destroy(void* h)
{
    cudaDeviceSynchronize();
    delete (RefSolver*)h;
}

VtSimParams* refcuda_params(void* h) { return &((RefSolver*)h)->P; }

// AddCloth (hpp L114-156): vertices are model-space; the world transform runs on the device (reference kernel).
int refcuda_add_cloth(void* h, const float* vertices, int numVertices, const unsigned* meshIndices, int numIndices,
                      const float* model16, float particleDiameter)
{
    RefSolver* s = (RefSolver*)h;
    const float fixedDt = 1.0f / 60.0f;
    int prev = (int)s->P.numParticles;
    s->P.numParticles += numVertices;
    s->P.particleDiameter = particleDiameter;
    s->P.deltaTime = fixedDt;
    s->P.maxSpeed = 2 * particleDiameter / fixedDt * s->P.numSubsteps;

    size_t off = s->positions.size();
    s->positions.resize(off + numVertices);
    memcpy(s->positions.data() + off, vertices, sizeof(float) * 3 * (size_t)numVertices);
    s->normals.resize(off + numVertices);
    for (int i = 0; i < numIndices; i++) s->indices.push_back(meshIndices[i] + prev);
    s->velocities.push_back(numVertices, glm::vec3(0));
    s->predicted.push_back(numVertices, glm::vec3(0));
    s->deltas.push_back(numVertices, glm::vec3(0));
    s->deltaCounts.push_back(numVertices, 0);
    s->invMasses.push_back(numVertices, 1.0f);

    glm::mat4 M;
    memcpy(&M, model16, 64);
    InitializePositions(s->positions, prev, numVertices, M);
    cudaDeviceSynchronize();

    // SpatialHashGPU ctor + SetInitialPositions (SpatialHashGPU.hpp L18-39)
    const int N = (int)s->P.numParticles;
    s->spacing = particleDiameter * s->P.hashCellSizeScalar;
    s->tableSize = 2 * N;
    s->neighbors.resize((size_t)N * s->P.maxNumNeighbors);
    s->particleHash.resize(N);
    s->particleIndex.resize(N);
    s->cellStart.resize(s->tableSize);
    s->cellEnd.resize(s->tableSize);
    s->initialPositions.resize(s->positions.size());
    for (size_t i = 0; i < s->positions.size(); i++) s->initialPositions[i] = s->positions[i];
    return prev;
}

void refcuda_add_stretch_bulk(void* h, const int* pairs, const float* lengths, size_t n)
{
    RefSolver* s = (RefSolver*)h;
    for (size_t i = 0; i < n; i++) {  // AddStretch, hpp L158-163
        s->stretchIndices.push_back(pairs[2 * i]);
        s->stretchIndices.push_back(pairs[2 * i + 1]);
        s->stretchLengths.push_back(lengths[i]);
    }
}

void refcuda_add_bend_bulk(void* h, const unsigned* quads, const float* angles, size_t n)
{
    RefSolver* s = (RefSolver*)h;
    for (size_t i = 0; i < n; i++) {  // AddBend, hpp L178-185
        for (int k = 0; k < 4; k++) s->bendIndices.push_back(quads[4 * i + k]);
        s->bendAngles.push_back(angles[i]);
    }
}

void refcuda_add_attach_slot(void* h, const float* p)
{
    ((RefSolver*)h)->attachSlotPositions.push_back(glm::vec3(p[0], p[1], p[2]));
}

void refcuda_add_attach_bulk(void* h, const int* pids, const int* slots, const float* dists, size_t n)
{
    RefSolver* s = (RefSolver*)h;
    for (size_t i = 0; i < n; i++) {  // AddAttach, hpp L170-176
        if (dists[i] == 0) s->invMasses[pids[i]] = 0;
        s->attachParticleIDs.push_back(pids[i]);
        s->attachSlotIDs.push_back(slots[i]);
        s->attachDistances.push_back(dists[i]);
    }
}

void refcuda_set_colliders(void* h, const SDFCollider* colliders, int n)
{  // UpdateColliders with pre-marshalled structs, hpp L187-205
    RefSolver* s = (RefSolver*)h;
    cudaDeviceSynchronize();
    s->sdfColliders.resize(n);
    for (int i = 0; i < n; i++) s->sdfColliders[i] = colliders[i];
}

// Simulate, hpp L56-111 (without the GL VBO sync)
void refcuda_simulate(void* h)
{
    RefSolver* s = (RefSolver*)h;
    Timer::StartTimerGPU("Solver_Total");
    float frameTime = 1.0f / 60.0f;
    float substepTime = frameTime / s->P.numSubsteps;
    SetSimulationParams(&s->P);
    CollideSDF(s->positions, s->sdfColliders, s->positions, (uint)s->sdfColliders.size(), frameTime);
    for (int substep = 0; substep < s->P.numSubsteps; substep++) {
        PredictPositions(s->predicted, s->velocities, s->positions, substepTime);
        if (s->P.enableSelfCollision) {
            if (substep % s->P.interleavedHash == 0) run_hash(s, s->predicted);
            CollideParticles(s->deltas, s->deltaCounts, s->predicted, s->invMasses, s->neighbors, s->positions);
        }
        CollideSDF(s->predicted, s->sdfColliders, s->positions, (uint)s->sdfColliders.size(), substepTime);
        for (int iteration = 0; iteration < s->P.numIterations; iteration++) {
            SolveStretch(s->predicted, s->deltas, s->deltaCounts, s->stretchIndices, s->stretchLengths, s->invMasses,
                         (uint)s->stretchLengths.size());
            SolveAttachment(s->predicted, s->deltas, s->deltaCounts, s->invMasses, s->attachParticleIDs, s->attachSlotIDs,
                            s->attachSlotPositions, s->attachDistances, (uint)s->attachParticleIDs.size());
            SolveBending(s->predicted, s->deltas, s->deltaCounts, s->bendIndices, s->bendAngles, s->invMasses,
                         (uint)s->bendAngles.size(), substepTime);
            ApplyDeltas(s->predicted, s->deltas, s->deltaCounts);
        }
        Finalize(s->velocities, s->positions, s->predicted, substepTime);
    }
    ComputeNormal(s->normals, s->positions, s->indices, (uint)(s->indices.size() / 3));
    Timer::EndTimerGPU("Solver_Total");
    cudaDeviceSynchronize();
    // what the GUI does every frame (GUI.cpp L111-121): drain the per-call event pairs
    for (int i = 0; i < 16; i++) g_lastMs[i] = Timer::GetTimerGPU(kLabels[i]);
}

void refcuda_hash_predicted(void* h)
{
    RefSolver* s = (RefSolver*)h;
    run_hash(s, s->predicted);
    cudaDeviceSynchronize();
    for (int i = 0; i < 16; i++) g_lastMs[i] = Timer::GetTimerGPU(kLabels[i]);
}

int refcuda_num_labels() { return 16; }
const char* refcuda_label(int i) { return kLabels[i]; }
double refcuda_label_ms(int i) { return g_lastMs[i]; }

// buffer ids follow include/velvet_b200.h VelvetBufferId
void* refcuda_buffer(void* h, int id, size_t* count, size_t* elemSize)
{
    RefSolver* s = (RefSolver*)h;
#define RET(b, es) { *count = s->b.size(); *elemSize = es; return (void*)s->b.data(); }
    switch (id) {
    case 0: RET(positions, 12) case 1: RET(normals, 12) case 2: RET(indices, 4) case 3: RET(velocities, 12)
    case 4: RET(predicted, 12) case 5: RET(deltas, 12) case 6: RET(deltaCounts, 4) case 7: RET(invMasses, 4)
    case 8: RET(stretchIndices, 4) case 9: RET(stretchLengths, 4) case 10: RET(bendIndices, 4) case 11: RET(bendAngles, 4)
    case 12: RET(attachParticleIDs, 4) case 13: RET(attachSlotIDs, 4) case 14: RET(attachDistances, 4)
    case 15: RET(attachSlotPositions, 12) case 16: RET(neighbors, 4) case 17: RET(initialPositions, 12)
    case 18: RET(particleHash, 4) case 19: RET(particleIndex, 4) case 20: RET(cellStart, 4) case 21: RET(cellEnd, 4)
    }
#undef RET
    *count = 0; *elemSize = 0;
    return nullptr;
}

void refcuda_sync() { cudaDeviceSynchronize(); }

}  // extern "C"
