/* Default user-configuration header #1 for zsc-b200 (sized types).
 *
 * The reference expects the integrator to supply this file (reference README.md:28-33,
 * reference test/zsc_test_global_types.h is its test instance).  An integrator's own copy placed
 * earlier on the include path replaces this one; it must provide the same names.
 */
#ifndef ZSC_CONF_GLOBAL_TYPES_H
#define ZSC_CONF_GLOBAL_TYPES_H

#include <stddef.h>
#include <stdint.h>

typedef uint8_t  U8;
typedef uint16_t U16;
typedef uint32_t U32;
typedef int32_t  I32;
typedef uint32_t z_crc_t;
typedef size_t   z_size_t;

#define U32_MAX ((U32)0xFFFFFFFFu)

/* compile-time check usable at file scope in C and C++ */
#define ZSC_COMPILE_ASSERT(test, msg) typedef U8(msg)[(test) ? 1 : -1]

#endif
