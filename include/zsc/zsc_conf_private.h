/* Default user-configuration header #2 for zsc-b200 (assert / warning / memory hooks used by the
 * host C layer).  Same macro names as the reference's private config
 * (reference test/zsc_test_private.h:38-90) so a flight build can redirect them unchanged.
 */
#ifndef ZSC_CONF_PRIVATE_H
#define ZSC_CONF_PRIVATE_H

#include <assert.h>
#include <stdio.h>
#include <string.h>
#include "zsc/zsc_conf_global_types.h"

#ifndef ZSC_PRIVATE
#define ZSC_PRIVATE static
#endif

#define ZSC_ASSERT(t)               assert(t)
#define ZSC_ASSERT1(t, a)           assert(t)
#define ZSC_ASSERT2(t, a, b)        assert(t)
#define ZSC_ASSERT3(t, a, b, c)     assert(t)

#ifndef ZSC_WARN_STREAM
#define ZSC_WARN_STREAM stdout
#endif
#define ZSC_WARN(f)                 fprintf(ZSC_WARN_STREAM, "ZSC WARNING " f "\n")
#define ZSC_WARN1(f, a)             fprintf(ZSC_WARN_STREAM, "ZSC WARNING " f "\n", a)
#define ZSC_WARN2(f, a, b)          fprintf(ZSC_WARN_STREAM, "ZSC WARNING " f "\n", a, b)
#define ZSC_WARN3(f, a, b, c)       fprintf(ZSC_WARN_STREAM, "ZSC WARNING " f "\n", a, b, c)
#define ZSC_WARN4(f, a, b, c, d)    fprintf(ZSC_WARN_STREAM, "ZSC WARNING " f "\n", a, b, c, d)
#define ZSC_WARN5(f, a, b, c, d, e) fprintf(ZSC_WARN_STREAM, "ZSC WARNING " f "\n", a, b, c, d, e)

#define zmemcpy memcpy
#define zmemcmp memcmp
#define zmemzero(d, n) memset((d), 0, (n))

#endif
