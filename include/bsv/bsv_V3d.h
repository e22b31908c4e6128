#ifndef CVTX_B200_BSV_COMPAT_V3d_H
#define CVTX_B200_BSV_COMPAT_V3d_H
/* Forwarding header: the reference includes <bsv/bsv_V3d.h> in a few places
 * (e.g. reference src/GridParticleOcttree.h:35); everything lives in bsv.h. */
#include "bsv.h"
#endif
