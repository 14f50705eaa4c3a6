// Forwarding header: the OpenMM shim keeps its declarations in openmm/ShimCore.h
#include "openmm/ShimCore.h"
