#ifndef OPENMM_CONSTVLANGEVININTEGRATOR_H_
#define OPENMM_CONSTVLANGEVININTEGRATOR_H_

// ConstVLangevinIntegrator: Langevin dynamics for a slab between two planar electrodes whose
// polarisation is represented by image charges (constant-potential, "ConstV").  Drop-in for the
// class of the same name in scychon/openmm_constV
// (openmmapi/include/openmm/ConstVLangevinIntegrator.h:46-176): identical constructor, accessors,
// defaults and kernel name, so user scripts, the SWIG wrapper and serialized XML keep working.  The
// time step itself runs in the B200 kernels behind IntegrateConstVLangevinStepKernel.

#include <string>
#include <vector>

#include "openmm/Integrator.h"
#include "openmm/Kernel.h"
#include "openmm/internal/windowsExport.h"

namespace OpenMM {

class OPENMM_EXPORT ConstVLangevinIntegrator : public Integrator {
public:
    /**
     * @param temperature    heat-bath temperature (K)
     * @param frictionCoeff  coupling to the heat bath (1/ps)
     * @param stepSize       integration step (ps)
     * @param numCells       number of cells, real + image (2 = one real cell and its mirror image);
     *                       particles [N/numCells, N) are the images and must be massless
     * @param zmax           cell size along z (nm); negative = box z length / numCells
     */
    ConstVLangevinIntegrator(double temperature, double frictionCoeff, double stepSize, int numCells = 2, double zmax = -1);

    int getNumCells() const { return numCellCopies; }
    /** Must be even; values above 2 describe several real/image pairs. */
    void setNumCells(int cells) { numCellCopies = cells; }

    double getCellSize() const { return cellsize; }
    void setCellSize(double zmax) { cellsize = zmax; }

    double getTemperature() const { return temperature; }
    void setTemperature(double temp) { temperature = temp; }

    double getFriction() const { return friction; }
    void setFriction(double coeff) { friction = coeff; }

    int getRandomNumberSeed() const { return randomNumberSeed; }
    /** 0 (the default) lets every Context pick its own seed; equal non-zero seeds reproduce the
     *  noise stream exactly (counter-based Philox keyed by seed, atom and step). */
    void setRandomNumberSeed(int seed) { randomNumberSeed = seed; }

    /** Advance the simulation by `steps` time steps. */
    void step(int steps);

protected:
    void initialize(ContextImpl& context);
    void cleanup();
    std::vector<std::string> getKernelNames();
    double computeKineticEnergy();

private:
    double temperature, friction, cellsize;
    int randomNumberSeed;
    int numCellCopies;
    Kernel kernel;
};

}  // namespace OpenMM

#endif /* OPENMM_CONSTVLANGEVININTEGRATOR_H_ */
