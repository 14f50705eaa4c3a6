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

}  // namespace

CudaIntegrateConstVLangevinStepKernel::~CudaIntegrateConstVLangevinStepKernel() {
    cu.setAsCurrent();
    if (slab != NULL)
        constv_destroy(slab);
}

void CudaIntegrateConstVLangevinStepKernel::initialize(const System& system, const ConstVLangevinIntegrator& integrator) {
    cu.getPlatformData().initializeContexts(system);
    cu.setAsCurrent();

    // System checks of the reference (CudaConstVKernels.cpp:76-96), performed by the library so the
    // Python host and this class cannot drift apart.
    const int numParticles = system.getNumParticles();
    std::vector<double> masses(numParticles);
    for (int i = 0; i < numParticles; i++)
        masses[i] = system.getParticleMass(i);
    Vec3 a, b, c;
    system.getDefaultPeriodicBoxVectors(a, b, c);
    check(constv_validate_system(numParticles, integrator.getNumCells(), masses.data(), c[2], integrator.getCellSize(), &zmax));

    constv_config config;
    std::memset(&config, 0, sizeof config);
    config.struct_size = sizeof config;
    config.device = -1;  // stay in the CUDA context OpenMM made current
    config.num_atoms = cu.getNumAtoms();
    config.padded_num_atoms = cu.getPaddedNumAtoms();
    config.num_cells = integrator.getNumCells();
    config.precision = cu.getUseDoublePrecision() ? CONSTV_PRECISION_DOUBLE
                       : cu.getUseMixedPrecision() ? CONSTV_PRECISION_MIXED : CONSTV_PRECISION_SINGLE;
    const char* mirror = std::getenv("CONSTV_MIRROR");
    config.mirror_mode = (mirror != NULL && std::strcmp(mirror, "negate") == 0) ? CONSTV_MIRROR_NEGATE : CONSTV_MIRROR_REFERENCE;
    config.flags = CONSTV_FLAGS_DEFAULT;
    config.zmax = zmax;
    unsigned long long seed = (unsigned int) integrator.getRandomNumberSeed();
    if (seed == 0) {  // "a unique seed is chosen when a Context is created" (ConstVLangevinIntegrator.h:138-140)
        std::random_device entropy;
        seed = ((unsigned long long) entropy() << 32) | entropy() | 1ull;
    }
    config.seed = seed;
    config.global_atom_offset = 0;
    config.boltz = BOLTZ;
    check(constv_create(&config, &slab));
    hasConstraints = system.getNumConstraints() > 0;
    slotsPermuted = false;
    bindArrays();
}

void CudaIntegrateConstVLangevinStepKernel::bindArrays() {
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    void* correction = cu.getUseMixedPrecision() ? devicePointer(cu.getPosqCorrection()) : NULL;
    const int32_t* order = slotsPermuted ? reinterpret_cast<const int32_t*>(devicePointer(cu.getAtomIndexArray())) : NULL;
    check(constv_bind_arrays(slab, devicePointer(cu.getPosq()), correction, devicePointer(cu.getVelm()),
                             devicePointer(cu.getForce()), devicePointer(integration.getPosDelta()), order));
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
