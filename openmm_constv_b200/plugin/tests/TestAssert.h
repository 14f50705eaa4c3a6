// Tiny assertion helpers for the plugin's C++ tests (OpenMM's AssertionUtilities.h is not available
// without OpenMM).  A failed check throws; main() reports and returns 1.
#ifndef CONSTV_TEST_ASSERT_H_
#define CONSTV_TEST_ASSERT_H_

#include <cmath>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>

#define CHECK(cond)                                                                             \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            std::ostringstream m__;                                                             \
            m__ << __FILE__ << ":" << __LINE__ << ": check failed: " #cond;                     \
            throw std::runtime_error(m__.str());                                                \
        }                                                                                       \
    } while (0)

#define CHECK_EQ(expected, found)                                                               \
    do {                                                                                        \
        if (!((expected) == (found))) {                                                         \
            std::ostringstream m__;                                                             \
            m__.precision(17);                                                                  \
            m__ << __FILE__ << ":" << __LINE__ << ": expected " << (expected) << ", found " << (found); \
            throw std::runtime_error(m__.str());                                                \
        }                                                                                       \
    } while (0)

// |expected - found| <= tol * max(1, |expected|), the scaling OpenMM's ASSERT_EQUAL_TOL uses
#define CHECK_CLOSE(expected, found, tol)                                                       \
    do {                                                                                        \
        const double e__ = (expected), f__ = (found);                                           \
        const double s__ = std::fabs(e__) > 1.0 ? std::fabs(e__) : 1.0;                          \
        if (!(std::fabs(e__ - f__) / s__ <= (tol))) {                                           \
            std::ostringstream m__;                                                             \
            m__.precision(17);                                                                  \
            m__ << __FILE__ << ":" << __LINE__ << ": expected " << e__ << ", found " << f__ << " (tol " << (tol) << ")"; \
            throw std::runtime_error(m__.str());                                                \
        }                                                                                       \
    } while (0)

// the statement must throw an exception whose what() is exactly `text`
#define CHECK_THROWS(statement, text)                                                           \
    do {                                                                                        \
        bool threw__ = false;                                                                   \
        try {                                                                                   \
            statement;                                                                          \
        } catch (const std::exception& ex__) {                                                  \
            threw__ = true;                                                                     \
            if (std::string(ex__.what()) != (text)) {                                           \
                std::ostringstream m__;                                                         \
                m__ << __FILE__ << ":" << __LINE__ << ": expected exception \"" << (text) << "\", got \"" << ex__.what() << "\""; \
                throw std::runtime_error(m__.str());                                            \
            }                                                                                   \
        }                                                                                       \
        if (!threw__) {                                                                         \
            std::ostringstream m__;                                                             \
            m__ << __FILE__ << ":" << __LINE__ << ": no exception; expected \"" << (text) << "\""; \
            throw std::runtime_error(m__.str());                                                \
        }                                                                                       \
    } while (0)

inline unsigned long long bitsOf(double v) {
    unsigned long long b;
    std::memcpy(&b, &v, sizeof b);
    return b;
}

#endif
