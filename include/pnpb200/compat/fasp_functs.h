/* pnp-b200 — forwarding header for `#include "fasp_functs.h"`: the prototypes live in fasp_compat.h */
#ifndef PNPB200_COMPAT_FASP_FUNCTS_H
#define PNPB200_COMPAT_FASP_FUNCTS_H
#include "../fasp_compat.h"
#endif
