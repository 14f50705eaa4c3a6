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

int envInt(const char* name, int fallback) {
    const char* v = std::getenv(name);
    return (v != NULL && *v != 0) ? std::atoi(v) : fallback;
}

}  // namespace

// ---- the slab behind both kernel classes -------------------------------------------------------------

// OpenMM calls this after every re-sort of the atoms, whoever triggered it (a step, setPositions,
// loadCheckpoint, the minimizer): polling cu.getAtomsWereReordered() inside execute(), as the reference does
// (CudaConstVKernels.cpp:139-142), misses a re-sort whose flag a later reorderAtoms() call has already cleared.
class ConstVSlabBinding::Listener : public CudaContext::ReorderListener {
public:
    explicit Listener(std::shared_ptr<OrderWatch> watch) : watch(watch) {}
    void execute() { watch->dirty = true; }
private:
    std::shared_ptr<OrderWatch> watch;
};

ConstVSlabBinding::~ConstVSlabBinding() {
    if (slab != NULL) {
        cu.setAsCurrent();
        constv_destroy(slab);
    }
}

void* ConstVSlabBinding::stream() {
#ifdef CONSTV_OPENMM_HAS_CURRENT_STREAM
    return reinterpret_cast<void*>(cu.getCurrentStream());
#else
    return NULL;  // OpenMM <= 7.4 launches everything on the default stream of its context
#endif
}

// System checks of the reference (CudaConstVKernels.cpp:76-96 and, verbatim again, :184-205), performed by
// the library so the Python host and these classes cannot drift apart; then the slab on OpenMM's context.
void ConstVSlabBinding::create(const System& system, int numCells, double cellSize, int randomNumberSeed) {
    const int numParticles = system.getNumParticles();
    std::vector<double> masses(numParticles);
    hasVirtualSites = false;
    for (int i = 0; i < numParticles; i++) {
        masses[i] = system.getParticleMass(i);
        hasVirtualSites = hasVirtualSites || system.isVirtualSite(i);
    }
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
    // The image charges (posq.w of the image slots, re-read every step by constVLangevin.cu:153) are kept in a
    // side array and re-read from posq after every re-sort and every CONSTV_IMAGE_CHARGE_REFRESH steps
    // (default 64; 1 = every step, the reference's behaviour to the letter for a host that rewrites charges
    // with updateParametersInContext while stepping; CONSTV_CACHE_IMAGE_CHARGE=0 turns the side array off).
    config.flags = CONSTV_FLAGS_DEFAULT;
    if (envInt("CONSTV_CACHE_IMAGE_CHARGE", 1) != 0)
        config.flags |= CONSTV_FLAG_CACHE_IMAGE_CHARGE;
    chargeRefreshInterval = envInt("CONSTV_IMAGE_CHARGE_REFRESH", 64);
    if (chargeRefreshInterval < 1) chargeRefreshInterval = 1;
    if (!(config.flags & CONSTV_FLAG_CACHE_IMAGE_CHARGE)) chargeRefreshInterval = 0x7fffffff;  // nothing to refresh
    config.zmax = zmax;
    unsigned long long seed = (unsigned int) randomNumberSeed;
    if (seed == 0) {  // "a unique seed is chosen when a Context is created" (ConstVLangevinIntegrator.h:138-140)
        std::random_device entropy;
        seed = ((unsigned long long) entropy() << 32) | entropy() | 1ull;
    }
    config.seed = seed;
    config.global_atom_offset = 0;
    config.boltz = BOLTZ;
    check(constv_create(&config, &slab));
    watch = std::make_shared<OrderWatch>();
    cu.addReorderListener(new Listener(watch));  // owned and deleted by the CudaContext
    slotsPermuted = false;
    bindArrays();
}

void ConstVSlabBinding::bindArrays() {
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    void* correction = cu.getUseMixedPrecision() ? devicePointer(cu.getPosqCorrection()) : NULL;
    const int32_t* order = slotsPermuted ? reinterpret_cast<const int32_t*>(devicePointer(cu.getAtomIndexArray())) : NULL;
    check(constv_bind_arrays(slab, devicePointer(cu.getPosq()), correction, devicePointer(cu.getVelm()),
                             devicePointer(cu.getForce()), devicePointer(integration.getPosDelta()), order));
    stepsSinceChargeRefresh = 0;  // binding invalidates the image-charge snapshot: the next step retakes it
}

void ConstVSlabBinding::beforeStep(void* stream) {
    // OpenMM re-sorts the atoms from time to time; from then on slots are addressed through atomIndex /
    // invAtomOrder (CudaConstVKernels.cpp:139-142).  The check runs on the first step and after every re-sort
    // the listener saw -- including those made between two steps -- and compares OpenMM's host copy of the order
    // with the identity, so that an unpermuted context keeps the coalesced identity-order kernels.
    if (watch->dirty || cu.getAtomsWereReordered()) {
        const std::vector<int>& order = cu.getAtomIndex();
        bool permuted = false;
        const int n = cu.getNumAtoms();
        for (int i = 0; i < n && !permuted; i++) permuted = order[i] != i;
        slotsPermuted = permuted;
        bindArrays();
        if (slotsPermuted) check(constv_update_inverse_order(slab, stream));
        watch->dirty = false;
    }
    // The Philox counter is (seed; atom, step) with step = OpenMM's own step count, so a checkpoint restored
    // into a new Context (cu.setStepCount) continues the noise sequence instead of replaying it from 0.
    if ((long long) cu.getStepCount() != lastStepCount) {
        check(constv_set_step_count(slab, (uint64_t) cu.getStepCount()));
        lastStepCount = cu.getStepCount();
    }
    if (stepsSinceChargeRefresh >= chargeRefreshInterval) {
        check(constv_refresh_image_charges(slab, stream));
        stepsSinceChargeRefresh = 0;
    }
    stepsSinceChargeRefresh++;
}

void ConstVSlabBinding::finishStep(void* stream) {
    cu.getIntegrationUtilities().computeVirtualSites();
    // The reference rewrites the images AFTER the virtual sites have been placed (part 2, computeVirtualSites,
    // updateImageParticlePositions: CudaConstVKernels.cpp:158-166 and :333-347), so the image of a charged
    // virtual site (the M site of a 4-site water) mirrors the site's NEW position.  The step kernels above have
    // already rewritten every image from the pre-site positions; with virtual sites in the System the rewrite is
    // run once more (it is idempotent for every atom that is not a site).
    if (hasVirtualSites) check(constv_mirror(slab, stream));
    lastStepCount = cu.getStepCount() + 1;
}

// ---- ConstVLangevinIntegrator --------------------------------------------------------------------------

CudaIntegrateConstVLangevinStepKernel::~CudaIntegrateConstVLangevinStepKernel() {
}

void CudaIntegrateConstVLangevinStepKernel::initialize(const System& system, const ConstVLangevinIntegrator& integrator) {
    cu.getPlatformData().initializeContexts(system);
    cu.setAsCurrent();
    binding.create(system, integrator.getNumCells(), integrator.getCellSize(), integrator.getRandomNumberSeed());
    hasConstraints = system.getNumConstraints() > 0;

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
            check(constv_settle_configure(binding.get(), found, clusterAtoms.data(), clusterParams.data()));
            fusedConstraints = true;
        }
    }
}

void CudaIntegrateConstVLangevinStepKernel::execute(ContextImpl& context, const ConstVLangevinIntegrator& integrator) {
    cu.setAsCurrent();
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    constv_slab* slab = binding.get();
    const double stepSize = integrator.getStepSize();
    integration.setNextStepSize(stepSize);
    check(constv_set_parameters(slab, integrator.getTemperature(), integrator.getFriction(), stepSize));
    void* stream = binding.stream();
    binding.beforeStep(stream);

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
    binding.finishStep(stream);

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
    check(constv_kinetic_energy(binding.get(), timeShift, &energy, binding.stream()));
    return energy;
}

// ---- ConstVDrudeLangevinIntegrator ------------------------------------------------------------------

CudaIntegrateConstVDrudeLangevinStepKernel::~CudaIntegrateConstVDrudeLangevinStepKernel() {
}

void CudaIntegrateConstVDrudeLangevinStepKernel::initialize(const System& system, const ConstVDrudeLangevinIntegrator& integrator,
                                                            const DrudeForce& force) {
    cu.getPlatformData().initializeContexts(system);
    cu.setAsCurrent();
    binding.create(system, integrator.getNumCells(), integrator.getCellSize(), integrator.getRandomNumberSeed());

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
    check(constv_drude_configure(binding.get(), force.getNumParticles(), pairs.empty() ? NULL : pairs.data()));
    hasConstraints = system.getNumConstraints() > 0;
}

void CudaIntegrateConstVDrudeLangevinStepKernel::execute(ContextImpl& context, const ConstVDrudeLangevinIntegrator& integrator) {
    cu.setAsCurrent();
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    constv_slab* slab = binding.get();
    const double stepSize = integrator.getStepSize();
    integration.setNextStepSize(stepSize);  // the reference uploads dt itself (CudaConstVKernels.cpp:271-281)
    check(constv_drude_set_parameters(slab, integrator.getTemperature(), integrator.getFriction(), integrator.getDrudeTemperature(),
                                      integrator.getDrudeFriction(), integrator.getMaxDrudeDistance(), stepSize));
    void* stream = binding.stream();
    binding.beforeStep(stream);  // CudaConstVKernels.cpp:303-306

    if (!hasConstraints) {
        // part 1, part 2, hard wall and image rewrite of CudaConstVKernels.cpp:307-344 in one kernel
        check(constv_drude_step_fused(slab, NULL, 0, stream));
    } else {
        check(constv_drude_step_part1(slab, NULL, 0, stream));
        integration.applyConstraints(integrator.getConstraintTolerance());
        check(constv_drude_step_part2(slab, stream));
    }
    binding.finishStep(stream);  // computeVirtualSites, then the image rewrite (:342-347)

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
    check(constv_kinetic_energy(binding.get(), timeShift, &energy, binding.stream()));
    return energy;
}
