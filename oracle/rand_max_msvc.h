/* oracle/rand_max_msvc.h -- TEST INFRASTRUCTURE ONLY, force-included (-include) when the
 * reference's own test program is compiled by oracle/Makefile (target ref_tests).
 *
 * The reference's tests scale their inputs with `(float)mrand() / (float)(RAND_MAX / max_float)`
 * where mrand() <= 32767 (reference test/testmain.c:87-92, test/testsamecpugpuresultmany.h:70).
 * With MSVC's RAND_MAX = 32767 -- the author's platform -- that is uniform [0, 10); with glibc's
 * 2^31 - 1 every coordinate lands in [0, 1.53e-4].  This header gives the
 * program the RAND_MAX it was written for.  No reference source is modified or copied. */
#include <stdlib.h>
#undef RAND_MAX
#define RAND_MAX 32767
