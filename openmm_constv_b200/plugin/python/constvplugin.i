/* SWIG interface of the `constvplugin` Python module for the B200 ConstV Langevin plugin.
 *
 * Same module name, class, constructor defaults, accessors and unit decoration as the reference's
 * wrapper (scychon/openmm_constV python/constvplugin.i:1,35-61,65-100), so `from constvplugin import ConstVLangevinIntegrator` in user
 * scripts (example/nacl_1m_spce/nacl_constv.py:5,36) keeps working.  Built by setup.py against a
 * real OpenMM install (needs SWIG and OpenMM's swig headers; neither exists in the build container,
 * where the ctypes module constvplugin.py in this directory is what the tests exercise).
 */
%module constvplugin

%import(module="simtk.openmm") "swig/OpenMMSwigHeaders.i"
%include "swig/typemaps.i"

/* std::vector<double> / std::vector<int> wrappers under the reference's names (constvplugin.i:12-16) */
%include "std_vector.i"
namespace std {
  %template(vectord) vector<double>;
  %template(vectori) vector<int>;
};

%{
#include "OpenMM.h"
#include "OpenMMDrude.h"
#include "OpenMMConstV.h"
%}

%pythoncode %{
import simtk.openmm as mm
import simtk.unit as unit
%}

/* getters that carry units on the Python side: exactly the reference's seven (constvplugin.i:35-61);
 * getCellSize / getNumCells / getRandomNumberSeed return bare numbers there and here */
%pythonappend OpenMM::ConstVLangevinIntegrator::getTemperature() const %{
   val = unit.Quantity(val, unit.kelvin)
%}
%pythonappend OpenMM::ConstVLangevinIntegrator::getFriction() const %{
   val = unit.Quantity(val, 1/unit.picosecond)
%}

%pythonappend OpenMM::ConstVDrudeLangevinIntegrator::getTemperature() const %{
   val = unit.Quantity(val, unit.kelvin)
%}
%pythonappend OpenMM::ConstVDrudeLangevinIntegrator::getFriction() const %{
   val = unit.Quantity(val, 1/unit.picosecond)
%}
%pythonappend OpenMM::ConstVDrudeLangevinIntegrator::getDrudeTemperature() const %{
   val = unit.Quantity(val, unit.kelvin)
%}
%pythonappend OpenMM::ConstVDrudeLangevinIntegrator::getDrudeFriction() const %{
   val = unit.Quantity(val, 1/unit.picosecond)
%}
%pythonappend OpenMM::ConstVDrudeLangevinIntegrator::getMaxDrudeDistance() const %{
   val = unit.Quantity(val, unit.nanometer)
%}

namespace OpenMM {

class ConstVLangevinIntegrator : public Integrator {
public:
    ConstVLangevinIntegrator(double temperature, double frictionCoeff, double stepSize, int numCells = 2, double zmax = -1);

    int getNumCells() const;
    void setNumCells(int cells);
    double getCellSize() const;
    void setCellSize(double zmax);
    double getTemperature() const;
    void setTemperature(double temp);
    double getFriction() const;
    void setFriction(double coeff);
    int getRandomNumberSeed() const;
    void setRandomNumberSeed(int seed);
    virtual void step(int steps);
};

class ConstVDrudeLangevinIntegrator : public Integrator {
public:
    ConstVDrudeLangevinIntegrator(double temperature, double friction, double drudeTemperature, double drudeFriction, double stepSize, int numCells = 2, double zmax = -1);

    int getNumCells() const;
    void setNumCells(int cells);
    double getCellSize() const;
    void setCellSize(double zmax);
    double getTemperature() const;
    void setTemperature(double temp);
    double getFriction() const;
    void setFriction(double coeff);
    double getDrudeTemperature() const;
    void setDrudeTemperature(double temp);
    double getDrudeFriction() const;
    void setDrudeFriction(double coeff);
    double getMaxDrudeDistance() const;
    void setMaxDrudeDistance(double distance);
    int getRandomNumberSeed() const;
    void setRandomNumberSeed(int seed);
    virtual void step(int steps);
};

}
