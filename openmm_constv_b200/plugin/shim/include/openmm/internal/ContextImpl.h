#include "openmm/ShimCore.h"
