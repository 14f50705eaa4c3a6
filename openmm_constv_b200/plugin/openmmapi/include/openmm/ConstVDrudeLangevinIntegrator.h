#ifndef OPENMM_CONSTVDRUDELANGEVININTEGRATOR_H_
#define OPENMM_CONSTVDRUDELANGEVININTEGRATOR_H_

// ConstVDrudeLangevinIntegrator: the image-charge slab integrator for polarizable (Drude
// oscillator) force fields.  Ordinary particles and the centre of mass of every (Drude particle,
// parent) pair are coupled to a heat bath at `temperature`; the motion of each Drude particle
// relative to its parent is coupled to a second, cold bath at `drudeTemperature`; an optional hard
// wall bounds the Drude displacement.  Drop-in for the class of the same name in
// scychon/openmm_constV (openmmapi/include/openmm/ConstVDrudeLangevinIntegrator.h,
// openmmapi/src/ConstVDrudeLangevinIntegrator.cpp:44-129): identical constructor, accessors,
// defaults, error messages and kernel name.  The System must contain exactly one DrudeForce.

#include <string>
#include <vector>

#include "openmm/Integrator.h"
#include "openmm/Kernel.h"
#include "openmm/State.h"
#include "openmm/internal/windowsExport.h"

namespace OpenMM {

class OPENMM_EXPORT ConstVDrudeLangevinIntegrator : public Integrator {
public:
    /**
     * @param temperature         heat bath of ordinary particles and pair centres of mass (K)
     * @param frictionCoeff       its coupling (1/ps)
     * @param drudeTemperature    heat bath of the Drude-parent relative motion (K)
     * @param drudeFrictionCoeff  its coupling (1/ps)
     * @param stepSize            integration step (ps)
     * @param numCells            number of cells, real + image; particles [N/numCells, N) are massless images
     * @param zmax                cell size along z (nm); negative = box z length / numCells
     */
    ConstVDrudeLangevinIntegrator(double temperature, double frictionCoeff, double drudeTemperature,
                                  double drudeFrictionCoeff, double stepSize, int numCells = 2, double zmax = -1);

    int getNumCells() const { return numCellCopies; }
    void setNumCells(int cells) { numCellCopies = cells; }
    double getCellSize() const { return cellsize; }
    void setCellSize(double zmax) { cellsize = zmax; }
    double getTemperature() const { return temperature; }
    void setTemperature(double temp) { temperature = temp; }
    double getFriction() const { return friction; }
    void setFriction(double coeff) { friction = coeff; }
    double getDrudeTemperature() const { return drudeTemperature; }
    void setDrudeTemperature(double temp) { drudeTemperature = temp; }
    double getDrudeFriction() const { return drudeFriction; }
    void setDrudeFriction(double coeff) { drudeFriction = coeff; }
    /** Largest allowed Drude-parent distance (nm); 0 (the default) disables the hard wall. */
    double getMaxDrudeDistance() const { return maxDrudeDistance; }
    /** Throws "setMaxDrudeDistance: Distance cannot be negative". */
    void setMaxDrudeDistance(double distance);
    int getRandomNumberSeed() const { return randomNumberSeed; }
    void setRandomNumberSeed(int seed) { randomNumberSeed = seed; }

    void step(int steps);

protected:
    void initialize(ContextImpl& context);
    void cleanup();
    void stateChanged(State::DataType changed) {}
    std::vector<std::string> getKernelNames();
    double computeKineticEnergy();

private:
    double temperature, friction, drudeTemperature, drudeFriction, maxDrudeDistance, cellsize;
    int randomNumberSeed;
    int numCellCopies;
    Kernel kernel;
};

}  // namespace OpenMM

#endif /* OPENMM_CONSTVDRUDELANGEVININTEGRATOR_H_ */
