#include "CudaConstVKernels.h"

#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "CudaIntegrationUtilities.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

using namespace OpenMM;

namespace {

// C-ABI status -> the exception type the reference throws (messages are produced by the library;
// for CONSTV_ERR_SYSTEM they are the reference's own texts).
void check(int status) {
    if (status != CONSTV_OK)
        throw OpenMMException(constv_last_error());
}

void* devicePointer(CudaArray& array) {
    return reinterpret_cast<void*>(array.getDevicePointer());
}

// System checks of the reference (CudaConstVKernels.cpp:76-96 and, verbatim again, :184-205), performed by
// the library so the Python host and these classes cannot drift apart; then the slab on OpenMM's context.
constv_slab* createSlab(CudaContext& cu, const System& system, int numCells, double cellSize, int randomNumberSeed, double& zmax) {
    const int numParticles = system.getNumParticles();
    std::vector<double> masses(numParticles);
    for (int i = 0; i < numParticles; i++)
        masses[i] = system.getParticleMass(i);
    Vec3 a, b, c;
    system.getDefaultPeriodicBoxVectors(a, b, c);
    check(constv_validate_system(numParticles, numCells, masses.data(), c[2], cellSize, &zmax));

    constv_config config;
    std::memset(&config, 0, sizeof config);
    config.struct_size = sizeof config;
    config.device = -1;  // stay in the CUDA context OpenMM made current
    config.num_atoms = cu.getNumAtoms();
    config.padded_num_atoms = cu.getPaddedNumAtoms();
    config.num_cells = numCells;
    config.precision = cu.getUseDoublePrecision() ? CONSTV_PRECISION_DOUBLE
                       : cu.getUseMixedPrecision() ? CONSTV_PRECISION_MIXED : CONSTV_PRECISION_SINGLE;
    const char* mirror = std::getenv("CONSTV_MIRROR");
    config.mirror_mode = (mirror != NULL && std::strcmp(mirror, "negate") == 0) ? CONSTV_MIRROR_NEGATE : CONSTV_MIRROR_REFERENCE;
    config.flags = CONSTV_FLAGS_DEFAULT;
    config.zmax = zmax;
    unsigned long long seed = (unsigned int) randomNumberSeed;
    if (seed == 0) {  // "a unique seed is chosen when a Context is created" (ConstVLangevinIntegrator.h:138-140)
        std::random_device entropy;
        seed = ((unsigned long long) entropy() << 32) | entropy() | 1ull;
    }
    config.seed = seed;
    config.global_atom_offset = 0;
    config.boltz = BOLTZ;
    constv_slab* slab = NULL;
    check(constv_create(&config, &slab));
    return slab;
}

void bindSlabArrays(CudaContext& cu, constv_slab* slab, bool slotsPermuted) {
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    void* correction = cu.getUseMixedPrecision() ? devicePointer(cu.getPosqCorrection()) : NULL;
    const int32_t* order = slotsPermuted ? reinterpret_cast<const int32_t*>(devicePointer(cu.getAtomIndexArray())) : NULL;
    check(constv_bind_arrays(slab, devicePointer(cu.getPosq()), correction, devicePointer(cu.getVelm()),
                             devicePointer(cu.getForce()), devicePointer(integration.getPosDelta()), order));
}

}  // namespace

CudaIntegrateConstVLangevinStepKernel::~CudaIntegrateConstVLangevinStepKernel() {
    cu.setAsCurrent();
    if (slab != NULL)
        constv_destroy(slab);
}

void CudaIntegrateConstVLangevinStepKernel::initialize(const System& system, const ConstVLangevinIntegrator& integrator) {
    cu.getPlatformData().initializeContexts(system);
    cu.setAsCurrent();
    slab = createSlab(cu, system, integrator.getNumCells(), integrator.getCellSize(), integrator.getRandomNumberSeed(), zmax);
    hasConstraints = system.getNumConstraints() > 0;
    slotsPermuted = false;
    bindArrays();

    // The constraint hand-off (reference :153).  When every constraint belongs to a rigid 3-site molecule --
    // the reference's example: SPC/E water with constraints=AllBonds -- the library solves them inside the
    // step kernel (SETTLE, as OpenMM's own solver would) and nothing is handed over.
    fusedConstraints = false;
    numRigidMolecules = 0;
    if (hasConstraints) {
        const int numConstraints = system.getNumConstraints();
        std::vector<int32_t> atom1(numConstraints), atom2(numConstraints);
        std::vector<double> distance(numConstraints);
        for (int i = 0; i < numConstraints; i++) {
            int p1, p2;
            system.getConstraintParameters(i, p1, p2, distance[i]);
            atom1[i] = p1;
            atom2[i] = p2;
        }
        std::vector<int32_t> clusterAtoms(numConstraints + 3);
        std::vector<float> clusterParams(numConstraints + 2);
        int32_t found = 0, other = 0;
        check(constv_settle_find_clusters(system.getNumParticles(), numConstraints, atom1.data(), atom2.data(), distance.data(),
                                          clusterAtoms.data(), clusterParams.data(), &found, &other));
        numRigidMolecules = found;
        if (found > 0 && other == 0 && std::getenv("CONSTV_NO_FUSED_CONSTRAINTS") == NULL) {
            check(constv_settle_configure(slab, found, clusterAtoms.data(), clusterParams.data()));
            fusedConstraints = true;
        }
    }
}

void CudaIntegrateConstVLangevinStepKernel::bindArrays() {
    bindSlabArrays(cu, slab, slotsPermuted);
}

void CudaIntegrateConstVLangevinStepKernel::execute(ContextImpl& context, const ConstVLangevinIntegrator& integrator) {
    cu.setAsCurrent();
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    const double stepSize = integrator.getStepSize();
    integration.setNextStepSize(stepSize);
    check(constv_set_parameters(slab, integrator.getTemperature(), integrator.getFriction(), stepSize));
    void* stream = reinterpret_cast<void*>(cu.getCurrentStream());

    // OpenMM re-sorts the atoms from time to time; from then on slots are addressed through
    // atomIndex / invAtomOrder (CudaConstVKernels.cpp:139-142).
    if (cu.getAtomsWereReordered()) {
        slotsPermuted = true;
        bindArrays();
        check(constv_update_inverse_order(slab, stream));
    }

    if (!hasConstraints) {
        // nothing to hand to the constraint solver: the whole step is one kernel
        check(constv_step_fused(slab, NULL, 0, stream));
    } else if (fusedConstraints) {
        // rigid 3-site molecules only: part 1, SETTLE, part 2 and the image rewrite in one kernel
        check(constv_step_fused_settle(slab, NULL, 0, stream));
    } else {
        check(constv_step_part1(slab, NULL, 0, stream));
        integration.applyConstraints(integrator.getConstraintTolerance());
        check(constv_step_part2_mirror(slab, stream));
    }
    integration.computeVirtualSites();

    cu.setTime(cu.getTime() + stepSize);
    cu.setStepCount(cu.getStepCount() + 1);
    cu.reorderAtoms();
}

double CudaIntegrateConstVLangevinStepKernel::computeKineticEnergy(ContextImpl& context, const ConstVLangevinIntegrator& integrator) {
    cu.setAsCurrent();
    const double timeShift = 0.5 * integrator.getStepSize();
    if (hasConstraints)  // OpenMM also projects the shifted velocities onto the constraints
        return cu.getIntegrationUtilities().computeKineticEnergy(timeShift);
    double energy = 0;
    check(constv_kinetic_energy(slab, timeShift, &energy, reinterpret_cast<void*>(cu.getCurrentStream())));
    return energy;
}

// ---- ConstVDrudeLangevinIntegrator ------------------------------------------------------------------

CudaIntegrateConstVDrudeLangevinStepKernel::~CudaIntegrateConstVDrudeLangevinStepKernel() {
    cu.setAsCurrent();
    if (slab != NULL)
        constv_destroy(slab);
}

void CudaIntegrateConstVDrudeLangevinStepKernel::initialize(const System& system, const ConstVDrudeLangevinIntegrator& integrator,
                                                            const DrudeForce& force) {
    cu.getPlatformData().initializeContexts(system);
    cu.setAsCurrent();
    slab = createSlab(cu, system, integrator.getNumCells(), integrator.getCellSize(), integrator.getRandomNumberSeed(), zmax);

    // (Drude particle, parent) of every DrudeForce entry; every other real atom is an ordinary particle
    // (CudaConstVKernels.cpp:213-230)
    std::vector<int32_t> pairs;
    for (int i = 0; i < force.getNumParticles(); i++) {
        int p, p1, p2, p3, p4;
        double charge, polarizability, aniso12, aniso34;
        force.getParticleParameters(i, p, p1, p2, p3, p4, charge, polarizability, aniso12, aniso34);
        pairs.push_back(p);
        pairs.push_back(p1);
    }
    check(constv_drude_configure(slab, force.getNumParticles(), pairs.empty() ? NULL : pairs.data()));
    hasConstraints = system.getNumConstraints() > 0;
    slotsPermuted = false;
    bindArrays();
}

void CudaIntegrateConstVDrudeLangevinStepKernel::bindArrays() {
    bindSlabArrays(cu, slab, slotsPermuted);
}

void CudaIntegrateConstVDrudeLangevinStepKernel::execute(ContextImpl& context, const ConstVDrudeLangevinIntegrator& integrator) {
    cu.setAsCurrent();
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    const double stepSize = integrator.getStepSize();
    integration.setNextStepSize(stepSize);  // the reference uploads dt itself (CudaConstVKernels.cpp:271-281)
    check(constv_drude_set_parameters(slab, integrator.getTemperature(), integrator.getFriction(), integrator.getDrudeTemperature(),
                                      integrator.getDrudeFriction(), integrator.getMaxDrudeDistance(), stepSize));
    void* stream = reinterpret_cast<void*>(cu.getCurrentStream());

    if (cu.getAtomsWereReordered()) {  // CudaConstVKernels.cpp:303-306
        slotsPermuted = true;
        bindArrays();
        check(constv_update_inverse_order(slab, stream));
    }

    if (!hasConstraints) {
        // part 1, part 2, hard wall and image rewrite of CudaConstVKernels.cpp:307-344 in one kernel
        check(constv_drude_step_fused(slab, NULL, 0, stream));
    } else {
        check(constv_drude_step_part1(slab, NULL, 0, stream));
        integration.applyConstraints(integrator.getConstraintTolerance());
        check(constv_drude_step_part2(slab, stream));
    }
    integration.computeVirtualSites();

    cu.setTime(cu.getTime() + stepSize);
    cu.setStepCount(cu.getStepCount() + 1);
    cu.reorderAtoms();
}

double CudaIntegrateConstVDrudeLangevinStepKernel::computeKineticEnergy(ContextImpl& context, const ConstVDrudeLangevinIntegrator& integrator) {
    cu.setAsCurrent();
    const double timeShift = 0.5 * integrator.getStepSize();  // CudaConstVKernels.cpp:355-357
    if (hasConstraints)
        return cu.getIntegrationUtilities().computeKineticEnergy(timeShift);
    double energy = 0;
    check(constv_kinetic_energy(slab, timeShift, &energy, reinterpret_cast<void*>(cu.getCurrentStream())));
    return energy;
}
