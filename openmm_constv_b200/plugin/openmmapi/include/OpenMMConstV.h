#ifndef OPENMM_CONSTV_H_
#define OPENMM_CONSTV_H_
// Umbrella header (as openmmapi/include/OpenMMConstV.h of the reference).
#include "openmm/Integrator.h"
#include "openmm/ConstVLangevinIntegrator.h"
#include "openmm/ConstVDrudeLangevinIntegrator.h"
#endif /* OPENMM_CONSTV_H_ */
