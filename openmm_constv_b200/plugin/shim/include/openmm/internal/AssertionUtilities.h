// The assertion macros OpenMM's tests use (shim).
#ifndef OPENMM_SHIM_ASSERTIONS_H_
#define OPENMM_SHIM_ASSERTIONS_H_
#include <cmath>
#include <sstream>
#include "openmm/ShimCore.h"
namespace OpenMM {
inline void throwException(const char* file, int line, const std::string& details) {
    std::stringstream msg;
    msg << "Assertion failure at " << file << ":" << line << ".  " << details;
    throw OpenMMException(msg.str());
}
}
#define ASSERT(cond) { if (!(cond)) OpenMM::throwException(__FILE__, __LINE__, #cond); }
#define ASSERT_EQUAL(expected, found) { if (!((expected) == (found))) { std::stringstream d; d << "Expected " << (expected) << ", found " << (found); OpenMM::throwException(__FILE__, __LINE__, d.str()); } }
#define ASSERT_EQUAL_TOL(expected, found, tol) { double s_ = std::max(1.0, std::fabs((double) (expected))); if (!(std::fabs((double) (expected) - (double) (found)) / s_ <= (tol))) { std::stringstream d; d << "Expected " << (expected) << ", found " << (found); OpenMM::throwException(__FILE__, __LINE__, d.str()); } }
#define ASSERT_USUALLY_EQUAL_TOL(expected, found, tol) ASSERT_EQUAL_TOL(expected, found, tol)
#define ASSERT_EQUAL_VEC(expected, found, tol) { ASSERT_EQUAL_TOL((expected)[0], (found)[0], tol); ASSERT_EQUAL_TOL((expected)[1], (found)[1], tol); ASSERT_EQUAL_TOL((expected)[2], (found)[2], tol); }
#endif
