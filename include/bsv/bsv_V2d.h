#ifndef CVTX_B200_BSV_COMPAT_V2d_H
#define CVTX_B200_BSV_COMPAT_V2d_H
/* Forwarding header: the reference includes <bsv/bsv_V2d.h> in a few places
 * (e.g. reference src/GridParticleOcttree.h:35); everything lives in bsv.h. */
#include "bsv.h"
#endif
